// TEST HARNESS, NOT PRODUCT CODE.
//
// The reference's own callers of the hot path, restated so that the tests can drive the product
// (include/glsns_solver.hpp -> the C ABI -> the CUDA kernels) exactly as the reference does:
//   NewtonNonLinearSolver::solve       include/core/newton_non_linear_solver.h:76-139
//   SkipNewtonNonLinearSolver::solve   include/core/skip_newton_non_linear_solver.h:53-133
//   PhysicsSolver's choice of driver   include/core/physics_solver.h:126-145
//   NavierStokesBase::iterate / first_iteration / finish_time_step
//                                      source/solvers/navier_stokes_base.cc:428-590
//   SimulationControl::add_time_step   source/core/simulation_control.cc:29-38
// These loops are the reference's ("stays on the host unchanged"): a
// maintainer who adopts the product keeps using Lethe's own headers.  They are transcribed here
// only because deal.II is not in this image, and they are compiled into
// tests/mirror/libglsns_mirror.so, never into libglsns.so.
#ifndef GLSNS_TESTS_REFERENCE_DRIVERS_HPP
#define GLSNS_TESTS_REFERENCE_DRIVERS_HPP

#include "../../include/glsns_solver.hpp"

namespace glsns
{
  // Newton with backtracking line search (newton_non_linear_solver.h:76-139)
  template <typename VectorType>
  class NewtonNonLinearSolver : public NonLinearSolver<VectorType>
  {
  public:
    using NonLinearSolver<VectorType>::NonLinearSolver;
    void
    solve(const TimeSteppingMethod time_stepping_method, const bool is_initial_step,
          const bool = true) override
    {
      double       current_res = 1.0, last_res = 1.0;
      const bool   first_step      = is_initial_step;
      unsigned int outer_iteration = 0;
      auto        *solver          = this->physics_solver;
      while ((current_res > this->params.tolerance) &&
             outer_iteration < this->params.max_iterations)
        {
          solver->evaluation_point = solver->present_solution;
          solver->assemble_matrix_and_rhs(time_stepping_method);
          if (outer_iteration == 0)
            {
              current_res = solver->system_rhs.l2_norm();
              last_res    = current_res;
            }
          if (this->params.verbosity != Parameters::Verbosity::quiet)
            solver->pcout << "Newton iteration: " << outer_iteration
                          << "  - Residual:  " << current_res << std::endl;
          solver->solve_linear_system(first_step);
          line_search(solver, time_stepping_method, current_res, last_res, this->params);
          solver->present_solution = solver->evaluation_point;
          last_res                 = current_res;
          ++outer_iteration;
        }
    }

    static void
    line_search(PhysicsSolver<VectorType> *solver, const TimeSteppingMethod method,
                double &current_res, const double last_res, const Parameters::NonLinearSolver &params)
    {
      for (double alpha = 1.0; alpha > 1e-3; alpha *= 0.5)
        {
          solver->local_evaluation_point = solver->present_solution;
          solver->local_evaluation_point.add(alpha, solver->newton_update);
          solver->apply_constraints();
          solver->evaluation_point = solver->local_evaluation_point;
          solver->assemble_rhs(method);
          current_res = solver->system_rhs.l2_norm();
          if (params.verbosity != Parameters::Verbosity::quiet)
            solver->pcout << "\t\talpha = " << std::setw(6) << alpha << std::setw(0)
                          << " res = " << std::setprecision(params.display_precision)
                          << current_res << std::endl;
          if (current_res < 0.9 * last_res || last_res < params.tolerance)
            break;
        }
    }
  };

  // Same loop; Jacobian + preconditioner rebuilt every `skip iterations` calls
  // (skip_newton_non_linear_solver.h:53-133)
  template <typename VectorType>
  class SkipNewtonNonLinearSolver : public NonLinearSolver<VectorType>
  {
  public:
    using NonLinearSolver<VectorType>::NonLinearSolver;
    void
    solve(const TimeSteppingMethod time_stepping_method, const bool is_initial_step,
          const bool force_matrix_renewal = true) override
    {
      double       current_res = 1.0, last_res = 1.0;
      const bool   first_step      = is_initial_step;
      unsigned int outer_iteration = 0;
      bool assembly_needed = consecutive_iters == 0 || is_initial_step || force_matrix_renewal;
      auto *solver         = this->physics_solver;
      while ((current_res > this->params.tolerance) &&
             outer_iteration < this->params.max_iterations)
        {
          solver->evaluation_point = solver->present_solution;
          if (assembly_needed)
            solver->assemble_matrix_and_rhs(time_stepping_method);
          else if (outer_iteration == 0)
            solver->assemble_rhs(time_stepping_method);
          if (outer_iteration == 0)
            {
              current_res = solver->system_rhs.l2_norm();
              last_res    = current_res;
            }
          if (this->params.verbosity != Parameters::Verbosity::quiet)
            solver->pcout << "Newton iteration: " << outer_iteration
                          << "  - Residual:  " << current_res << std::endl;
          solver->solve_linear_system(first_step, assembly_needed);
          NewtonNonLinearSolver<VectorType>::line_search(solver, time_stepping_method,
                                                         current_res, last_res, this->params);
          solver->present_solution = solver->evaluation_point;
          last_res                 = current_res;
          ++outer_iteration;
          assembly_needed = false;
        }
      if (!force_matrix_renewal)
        {
          consecutive_iters++;
          consecutive_iters = consecutive_iters % this->params.skip_iterations;
        }
    }

  private:
    unsigned int consecutive_iters = 0;
  };

  // PhysicsSolver's constructor from Parameters::NonLinearSolver (physics_solver.h:126-145) picks
  // the driver; here as a factory whose result is attached to solver->non_linear_solver
  template <typename VectorType>
  NonLinearSolver<VectorType> *
  make_non_linear_solver(PhysicsSolver<VectorType> *solver, const Parameters::NonLinearSolver &p)
  {
    switch (p.solver)
      {
        case Parameters::NonLinearSolver::SolverType::newton:
          return new NewtonNonLinearSolver<VectorType>(solver, p);
        case Parameters::NonLinearSolver::SolverType::skip_newton:
          return new SkipNewtonNonLinearSolver<VectorType>(solver, p);
        default:
          return nullptr;
      }
  }

  // ------------------------------------------------------------------------------------------
  // Time-stepping glue of NavierStokesBase around solve_non_linear_system (the caller of the hot
  // path in a transient run, source/solvers/navier_stokes_base.cc:428-590), for any solver that
  // has present_solution, solution_m1..m3, time_steps_vector and solve_non_linear_system.
  // ------------------------------------------------------------------------------------------
  inline bool
  is_bdf(const TimeSteppingMethod method)
  {
    return method == TimeSteppingMethod::bdf1 || method == TimeSteppingMethod::bdf2 ||
           method == TimeSteppingMethod::bdf3;
  }

  // SimulationControl::add_time_step (source/core/simulation_control.cc:29-38): the vector that
  // get_time_steps_vector() hands to the assembly, newest first
  inline void
  add_time_step(std::vector<double> &time_steps_vector, const double dt)
  {
    for (size_t i = time_steps_vector.size() - 1; i > 0; --i)
      time_steps_vector[i] = time_steps_vector[i - 1];
    time_steps_vector[0] = dt;
  }

  // NavierStokesBase::iterate (navier_stokes_base.cc:461-505): the SDIRK stages of one time step
  // (stage results become solution_m2 / solution_m3), or one solve for steady / BDF
  template <class Solver>
  void
  iterate(Solver &s, const TimeSteppingMethod method)
  {
    if (method == TimeSteppingMethod::sdirk2)
      {
        s.solve_non_linear_system(TimeSteppingMethod::sdirk2_1, false, false);
        s.solution_m2 = s.present_solution;
        s.solve_non_linear_system(TimeSteppingMethod::sdirk2_2, false, false);
      }
    else if (method == TimeSteppingMethod::sdirk3)
      {
        s.solve_non_linear_system(TimeSteppingMethod::sdirk3_1, false, false);
        s.solution_m2 = s.present_solution;
        s.solve_non_linear_system(TimeSteppingMethod::sdirk3_2, false, false);
        s.solution_m3 = s.present_solution;
        s.solve_non_linear_system(TimeSteppingMethod::sdirk3_3, false, false);
      }
    else
      s.solve_non_linear_system(method, false, false);
  }

  // NavierStokesBase::first_iteration (navier_stokes_base.cc:511-590): BDF2 / BDF3 start with
  // Euler steps of dt * startup_timestep_scaling (`startup time scaling`, default 0.4) and finish
  // the step with the rest; `dt` has already been added to time_steps_vector by integrate().
  template <class Solver>
  void
  first_iteration(Solver &s, const TimeSteppingMethod method, const double dt,
                  const double startup_timestep_scaling)
  {
    if (!is_bdf(method) || method == TimeSteppingMethod::bdf1)
      iterate(s, method);
    else if (method == TimeSteppingMethod::bdf2)
      {
        add_time_step(s.time_steps_vector, dt * startup_timestep_scaling);
        s.solve_non_linear_system(TimeSteppingMethod::bdf1, false, true);
        s.solution_m2 = s.solution_m1;
        s.solution_m1 = s.present_solution;
        add_time_step(s.time_steps_vector, dt * (1. - startup_timestep_scaling));
        s.solve_non_linear_system(TimeSteppingMethod::bdf2, false, true);
      }
    else // bdf3
      {
        const double time_step = dt * startup_timestep_scaling;
        add_time_step(s.time_steps_vector, time_step);
        s.solve_non_linear_system(TimeSteppingMethod::bdf1, false, true);
        s.solution_m2 = s.solution_m1;
        s.solution_m1 = s.present_solution;
        add_time_step(s.time_steps_vector, time_step);
        s.solve_non_linear_system(TimeSteppingMethod::bdf1, false, true);
        s.solution_m3 = s.solution_m2;
        s.solution_m2 = s.solution_m1;
        s.solution_m1 = s.present_solution;
        add_time_step(s.time_steps_vector, dt * (1. - 2. * startup_timestep_scaling));
        s.solve_non_linear_system(TimeSteppingMethod::bdf3, false, true);
      }
  }

  // NavierStokesBase::finish_time_step (navier_stokes_base.cc:428-442), the vectors (the CFL number
  // of the new solution is GLSNavierStokesSolver::calculate_CFL; checkpoints stay with the host)
  template <class Solver>
  void
  finish_time_step(Solver &s, const TimeSteppingMethod method)
  {
    if (method != TimeSteppingMethod::steady)
      {
        s.solution_m3 = s.solution_m2;
        s.solution_m2 = s.solution_m1;
        s.solution_m1 = s.present_solution;
      }
  }

} // namespace glsns

#endif
