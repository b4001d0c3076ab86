// TEST HARNESS (tests/mirror/libglsns_mirror.so), not part of libglsns.so.
// Compiled front end of the host-side mirror (include/glsns_solver.hpp) driven by the transcribed
// reference drivers (reference_drivers.hpp), for the tests:
//   * the reference's fake physics backend (tests/core/non_linear_test_system_01.h:50-129: the 2x2
//     system x0^2 + x1 = 0, 2 x1 + 3 = 0) plugged into the mirrored Newton drivers — host only;
//   * GLSNavierStokesSolver over the C ABI, built from a glsnsh_mesh and a .prm text.
#include <sstream>

#include "reference_drivers.hpp"

extern "C" {
typedef struct glsnsh_mesh glsnsh_mesh;
void glsnsh_mesh_fill_desc(const glsnsh_mesh *m, glsns_fe_desc *fe, glsns_mesh_desc *md);
}

namespace
{
  // tests/core/non_linear_test_system_01.h: Jacobian [[2 x0, 1], [0, 2]], residual -(F(x))
  class TestClass : public glsns::PhysicsSolver<glsns::Vector>
  {
  public:
    explicit TestClass(const glsns::Parameters::NonLinearSolver &params)
      : PhysicsSolver(nullptr)
    {
      non_linear_solver = glsns::make_non_linear_solver<glsns::Vector>(this, params);
      evaluation_point.reinit(2), system_rhs.reinit(2), local_evaluation_point.reinit(2);
      present_solution.reinit(2), newton_update.reinit(2);
      present_solution[0] = 1, present_solution[1] = 0;
    }
    void
    assemble_matrix_and_rhs(const glsns::TimeSteppingMethod m) override
    {
      J[0] = 2 * evaluation_point[0], J[1] = 1, J[2] = 0, J[3] = 2;
      assemble_rhs(m);
      ++n_matrix;
    }
    void
    assemble_rhs(const glsns::TimeSteppingMethod) override
    {
      const double x0 = evaluation_point[0], x1 = evaluation_point[1];
      system_rhs[0] = -(x0 * x0 + x1), system_rhs[1] = -(2 * x1 + 3);
    }
    void
    solve_linear_system(const bool, const bool) override
    {
      const double det = J[0] * J[3] - J[1] * J[2];
      newton_update[0] = (J[3] * system_rhs[0] - J[1] * system_rhs[1]) / det;
      newton_update[1] = (-J[2] * system_rhs[0] + J[0] * system_rhs[1]) / det;
    }
    void
    apply_constraints() override
    {}
    double J[4]     = {0, 0, 0, 0};
    int    n_matrix = 0;
  };

  // records what the time-stepping glue asks of a solver: (method, first, renew, time_steps[0..2],
  // tags of m1, m2, m3 at the time of the call); "solving" gives present_solution a fresh tag
  struct RecordingSolver
  {
    glsns::Vector       present_solution, solution_m1, solution_m2, solution_m3;
    std::vector<double> time_steps_vector = {0, 0, 0, 0};
    std::vector<double> log;
    int                 next_tag = 1;
    RecordingSolver()
    {
      present_solution.reinit(1), solution_m1.reinit(1), solution_m2.reinit(1), solution_m3.reinit(1);
    }
    void
    solve_non_linear_system(const glsns::TimeSteppingMethod m, const bool first, const bool renew)
    {
      log.insert(log.end(), {(double)(int)m, (double)first, (double)renew, time_steps_vector[0],
                             time_steps_vector[1], time_steps_vector[2], solution_m1[0],
                             solution_m2[0], solution_m3[0]});
      present_solution[0] = next_tag++;
    }
  };

  struct SolverHandle
  {
    glsns::GLSNavierStokesSolver *solver = nullptr;
    std::ostringstream            log;
    std::string                   log_copy, error;
  };
} // namespace

extern "C" {

// newton_non_linear_solver_01.cc / skip_newton_non_linear_solver_01.cc: tolerance 1e-8, 10 iterations.
// Returns the number of Jacobian assemblies; x_out[2] receives present_solution.
int
glsnsh_newton_toy(int use_skip_newton, int skip_iterations, double *x_out)
{
  glsns::Parameters::NonLinearSolver params;
  params.verbosity = glsns::Parameters::Verbosity::quiet;
  params.solver    = use_skip_newton ? glsns::Parameters::NonLinearSolver::SolverType::skip_newton :
                                       glsns::Parameters::NonLinearSolver::SolverType::newton;
  params.tolerance = 1e-8, params.max_iterations = 10, params.display_precision = 4;
  params.skip_iterations = skip_iterations;
  TestClass solver(params);
  solver.solve_non_linear_system(glsns::TimeSteppingMethod::steady, true, true);
  x_out[0] = solver.present_solution[0], x_out[1] = solver.present_solution[1];
  return solver.n_matrix;
}

// The time-stepping glue on the recording solver: n_steps time steps of size dt with `method`
// (TimeSteppingMethod value), the first one through first_iteration.  out receives 9 doubles per
// solve_non_linear_system call (see RecordingSolver); returns the number of calls.
int
glsnsh_time_stepping_trace(int method, int n_steps, double dt, double startup_scaling, double *out,
                           int out_len)
{
  RecordingSolver                s;
  const glsns::TimeSteppingMethod m = static_cast<glsns::TimeSteppingMethod>(method);
  glsns::finish_time_step(s, m); // after set_initial_condition (tag 0 = the initial condition)
  for (int k = 0; k < n_steps; ++k)
    {
      glsns::add_time_step(s.time_steps_vector, dt); // SimulationControlTransient::integrate
      if (k == 0)
        glsns::first_iteration(s, m, dt, startup_scaling);
      else
        glsns::iterate(s, m);
      glsns::finish_time_step(s, m);
    }
  const int n = (int)s.log.size();
  for (int i = 0; i < n && i < out_len; ++i)
    out[i] = s.log[i];
  return n / 9;
}

// Parse a .prm text; out[0..17] (18 doubles) = non-linear {tolerance, max iterations, skip iterations, solver},
// linear {relative, minimum, max iters, fill, atol}; returns 0 or 1 (+ message in err).
int
glsnsh_parse_prm(const char *text, double *out, char *err, int err_len)
{
  try
    {
      glsns::NavierStokesSolverParameters p;
      p.parse(glsns::ParameterFile(text));
      out[0] = p.non_linear_solver.tolerance, out[1] = p.non_linear_solver.max_iterations;
      out[2] = p.non_linear_solver.skip_iterations, out[3] = (int)p.non_linear_solver.solver;
      out[4] = p.linear_solver.relative_residual, out[5] = p.linear_solver.minimum_residual;
      out[6] = p.linear_solver.max_iterations, out[7] = p.linear_solver.ilu_precond_fill;
      out[8] = p.linear_solver.ilu_precond_atol, out[9] = p.linear_solver.ilu_precond_rtol;
      out[10] = (int)p.linear_solver.solver, out[11] = p.physical_properties.viscosity;
      out[12] = p.fem_parameters.velocity_order, out[13] = p.fem_parameters.pressure_order;
      out[14] = (int)p.velocitySource.type, out[15] = p.velocitySource.omega_z;
      out[16] = (int)p.initial_condition.type, out[17] = p.initial_condition.viscosity;
      return 0;
    }
  catch (const std::exception &e)
    {
      snprintf(err, err_len, "%s", e.what());
      return 1;
    }
}

void *
glsnsh_solver_create(const glsnsh_mesh *mesh, const char *prm_text, const double *forcing_at_q,
                     int cuda_device)
{
  SolverHandle *h = new SolverHandle();
  try
    {
      glsns::NavierStokesSolverParameters p;
      p.parse(glsns::ParameterFile(prm_text ? prm_text : ""));
      glsns_fe_desc   fe;
      glsns_mesh_desc md;
      glsnsh_mesh_fill_desc(mesh, &fe, &md);
      h->solver = new glsns::GLSNavierStokesSolver(p, fe, md, forcing_at_q, cuda_device);
      h->solver->non_linear_solver =
        glsns::make_non_linear_solver<glsns::Vector>(h->solver, p.non_linear_solver);
      h->solver->pcout.set_stream(h->log);
    }
  catch (const std::exception &e)
    {
      h->error = e.what();
    }
  return h;
}

const char *
glsnsh_solver_error(void *handle)
{
  return ((SolverHandle *)handle)->error.c_str();
}

void
glsnsh_solver_destroy(void *handle)
{
  SolverHandle *h = (SolverHandle *)handle;
  delete h->solver;
  delete h;
}

// which: 0 present_solution, 1 solution_m1, 2 m2, 3 m3
void
glsnsh_solver_set_vector(void *handle, int which, const double *data, int64_t n)
{
  SolverHandle  *h = (SolverHandle *)handle;
  glsns::Vector *v = which == 0 ? &h->solver->present_solution :
                     which == 1 ? &h->solver->solution_m1 :
                     which == 2 ? &h->solver->solution_m2 :
                                  &h->solver->solution_m3;
  for (int64_t i = 0; i < n; ++i)
    (*v)[i] = data[i];
}

void
glsnsh_solver_get_present(void *handle, double *data, int64_t n)
{
  SolverHandle *h = (SolverHandle *)handle;
  for (int64_t i = 0; i < n; ++i)
    data[i] = h->solver->present_solution[i];
}

void
glsnsh_solver_set_time_steps(void *handle, const double *dts, int n)
{
  SolverHandle *h = (SolverHandle *)handle;
  for (int i = 0; i < n && i < 4; ++i)
    h->solver->time_steps_vector[i] = dts[i];
}

// solve_non_linear_system(method, first_iteration, force_matrix_renewal).
// Returns 0 ok, 3 NoConvergence (SolverControl::NoConvergence), 1 any other exception.
int
glsnsh_solver_solve_non_linear_system(void *handle, int method, int first_iteration,
                                      int force_matrix_renewal)
{
  SolverHandle *h = (SolverHandle *)handle;
  try
    {
      h->solver->solve_non_linear_system(static_cast<glsns::TimeSteppingMethod>(method),
                                         first_iteration != 0, force_matrix_renewal != 0);
      return 0;
    }
  catch (const glsns::NoConvergence &e)
    {
      h->error = e.what();
      return 3;
    }
  catch (const std::exception &e)
    {
      h->error = e.what();
      return 1;
    }
}

// set_initial_condition(L2projection) / calculate_CFL of the mirror.  Return 0 ok, 3
// NoConvergence, 1 any other exception.
int
glsnsh_solver_set_initial_condition_l2(void *handle, const double *initial_at_q)
{
  SolverHandle *h = (SolverHandle *)handle;
  try
    {
      h->solver->set_initial_condition_L2projection(initial_at_q);
      return 0;
    }
  catch (const glsns::NoConvergence &e)
    {
      h->error = e.what();
      return 3;
    }
  catch (const std::exception &e)
    {
      h->error = e.what();
      return 1;
    }
}

// set_initial_condition(type): type < 0 takes the `initial conditions` subsection's, otherwise
// Parameters::InitialConditionType (0 none, 1 L2projection, 2 viscous, 3 nodal).
int
glsnsh_solver_set_initial_condition(void *handle, int type, const double *initial_nodal,
                                    const double *initial_at_q)
{
  SolverHandle *h = (SolverHandle *)handle;
  try
    {
      h->solver->set_initial_condition(
        type < 0 ? h->solver->parameters().initial_condition.type :
                   static_cast<glsns::Parameters::InitialConditionType>(type),
        initial_nodal, initial_at_q);
      return 0;
    }
  catch (const glsns::NoConvergence &e)
    {
      h->error = e.what();
      return 3;
    }
  catch (const std::exception &e)
    {
      h->error = e.what();
      return 1;
    }
}

int
glsnsh_solver_calculate_cfl(void *handle, const double *shape_u_at_centre, double time_step,
                            double *cfl)
{
  SolverHandle *h = (SolverHandle *)handle;
  try
    {
      *cfl = h->solver->calculate_CFL(shape_u_at_centre, time_step);
      return 0;
    }
  catch (const std::exception &e)
    {
      h->error = e.what();
      return 1;
    }
}

// NavierStokesBase's time-stepping glue on the device-backed solver (navier_stokes_base.cc:428-590):
// what = 0 integrate (add_time_step(dt)), 1 first_iteration, 2 iterate, 3 finish_time_step.
// Returns 0 ok, 3 NoConvergence, 1 any other exception.
int
glsnsh_solver_time_step(void *handle, int what, int method, double dt, double startup_scaling)
{
  SolverHandle *h = (SolverHandle *)handle;
  try
    {
      const glsns::TimeSteppingMethod m = static_cast<glsns::TimeSteppingMethod>(method);
      if (what == 0)
        glsns::add_time_step(h->solver->time_steps_vector, dt);
      else if (what == 1)
        glsns::first_iteration(*h->solver, m, dt, startup_scaling);
      else if (what == 2)
        glsns::iterate(*h->solver, m);
      else
        glsns::finish_time_step(*h->solver, m);
      return 0;
    }
  catch (const glsns::NoConvergence &e)
    {
      h->error = e.what();
      return 3;
    }
  catch (const std::exception &e)
    {
      h->error = e.what();
      return 1;
    }
}

// everything the solver wrote to pcout so far
const char *
glsnsh_solver_log(void *handle)
{
  SolverHandle *h = (SolverHandle *)handle;
  h->log_copy     = h->log.str();
  return h->log_copy.c_str();
}

} // extern "C"
