// glsns_solver.hpp — host-side C++ mirror of the reference's solver interface for the hot path,
// written above the C ABI (glsns.h).  Header only, C++17, no deal.II.
//
// Mirrors (same names, argument meaning and error behaviour; reference paths relative to the
// reference root):
//   Parameters::{Verbosity, SimulationControl::TimeSteppingMethod, NonLinearSolver, LinearSolver,
//               FEM, PhysicalProperties, VelocitySource}   include/core/parameters.h,
//                                                           source/core/parameters.cc
//   PhysicsSolver<VectorType>                               include/core/physics_solver.h:38-158
//   NonLinearSolver (the abstract driver interface)         include/core/non_linear_solver.h
//   GLSNavierStokesSolver::{assemble_matrix_and_rhs, assemble_rhs, solve_linear_system,
//                           setup_ILU, solve_system_GMRES}  source/solvers/gls_navier_stokes.cc:916-1289
//
// A maintainer of the reference drops the three overrides of glsns::GLSNavierStokesSolver into
// GLSNavierStokesSolver<dim> (INTEGRATION.md shows the diff); here VectorType is glsns::Vector, a
// plain host vector with the handful of operations the Newton drivers use.  The Newton drivers
// themselves (NewtonNonLinearSolver, SkipNewtonNonLinearSolver) and NavierStokesBase's
// time-stepping glue are the reference's and are NOT part of this product: the tests drive this
// class with a labelled transcription of them that lives under tests/mirror/.
#ifndef GLSNS_SOLVER_HPP
#define GLSNS_SOLVER_HPP

#include <cmath>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "glsns.h"

namespace glsns
{
  // ------------------------------------------------------------------------------------------
  // .prm files: the subset of deal.II's ParameterHandler syntax the shipped files use
  // ------------------------------------------------------------------------------------------
  class ParameterFile
  {
  public:
    ParameterFile() = default;
    explicit ParameterFile(const std::string &text)
    {
      parse(text);
    }

    void
    parse(const std::string &text)
    {
      std::istringstream       in(text);
      std::string              line;
      std::vector<std::string> path;
      while (std::getline(in, line))
        {
          const auto hash = line.find('#');
          if (hash != std::string::npos)
            line.erase(hash);
          line = trim(line);
          if (line.empty())
            continue;
          if (line.rfind("subsection", 0) == 0)
            path.push_back(trim(line.substr(10)));
          else if (line == "end")
            {
              if (path.empty())
                throw std::runtime_error("unbalanced 'end' in parameter file");
              path.pop_back();
            }
          else if (line.rfind("set", 0) == 0)
            {
              const auto eq = line.find('=');
              if (eq == std::string::npos)
                throw std::runtime_error("malformed 'set' line: " + line);
              std::string key = trim(line.substr(3, eq - 3));
              std::string sec;
              for (const auto &p : path)
                sec += p + "/";
              values[sec + key] = trim(line.substr(eq + 1));
            }
          else
            throw std::runtime_error("unrecognised parameter line: " + line);
        }
      if (!path.empty())
        throw std::runtime_error("subsection '" + path.back() + "' is not closed");
    }

    std::string
    get(const std::string &section, const std::string &key, const std::string &def) const
    {
      const auto it = values.find(section + "/" + key);
      return it == values.end() ? def : it->second;
    }
    double
    get_double(const std::string &section, const std::string &key, double def) const
    {
      const auto it = values.find(section + "/" + key);
      return it == values.end() ? def : std::stod(it->second);
    }
    long
    get_integer(const std::string &section, const std::string &key, long def) const
    {
      const auto it = values.find(section + "/" + key);
      return it == values.end() ? def : std::stol(it->second);
    }

  private:
    static std::string
    trim(const std::string &s)
    {
      const auto a = s.find_first_not_of(" \t\r\n");
      if (a == std::string::npos)
        return "";
      const auto b = s.find_last_not_of(" \t\r\n");
      return s.substr(a, b - a + 1);
    }
    std::map<std::string, std::string> values;
  };

  namespace Parameters
  {
    enum class Verbosity
    {
      quiet,
      verbose
    };

    inline Verbosity
    parse_verbosity(const std::string &op)
    {
      if (op == "verbose")
        return Verbosity::verbose;
      if (op == "quiet")
        return Verbosity::quiet;
      throw std::runtime_error("Invalid verbosity level");
    }

    struct SimulationControl
    {
      // include/core/parameters.h:56-69
      enum class TimeSteppingMethod
      {
        steady,
        bdf1,
        bdf2,
        bdf3,
        sdirk2,
        sdirk2_1,
        sdirk2_2,
        sdirk3,
        sdirk3_1,
        sdirk3_2,
        sdirk3_3
      } method = TimeSteppingMethod::steady;
    };

    // `non-linear solver` subsection, source/core/parameters.cc:373-447
    struct NonLinearSolver
    {
      enum class SolverType
      {
        newton,
        skip_newton
      };
      Verbosity    verbosity         = Verbosity::verbose;
      SolverType   solver            = SolverType::newton;
      double       tolerance         = 1e-6;
      unsigned int max_iterations    = 10;
      unsigned int skip_iterations   = 1;
      unsigned int display_precision = 4;

      void
      parse_parameters(const ParameterFile &prm)
      {
        const std::string s = "non-linear solver";
        verbosity           = parse_verbosity(prm.get(s, "verbosity", "verbose"));
        const std::string m = prm.get(s, "solver", "newton");
        if (m == "newton")
          solver = SolverType::newton;
        else if (m == "skip_newton")
          solver = SolverType::skip_newton;
        else
          throw std::runtime_error("Invalid non-linear solver ");
        tolerance         = prm.get_double(s, "tolerance", 1e-6);
        max_iterations    = (unsigned)prm.get_integer(s, "max iterations", 10);
        skip_iterations   = (unsigned)prm.get_integer(s, "skip iterations", 1);
        display_precision = (unsigned)prm.get_integer(s, "residual precision", 4);
      }
    };

    // `linear solver` subsection, source/core/parameters.cc:502-648
    struct LinearSolver
    {
      enum class SolverType
      {
        gmres,
        bicgstab,
        amg
      };
      Verbosity    verbosity          = Verbosity::verbose;
      SolverType   solver             = SolverType::gmres;
      unsigned int residual_precision = 6;
      double       relative_residual  = 1e-3;
      double       minimum_residual   = 1e-8;
      int          max_iterations     = 1000;
      double       ilu_precond_fill   = 0;
      double       ilu_precond_atol   = 1e-8;
      double       ilu_precond_rtol   = 1.00;

      void
      parse_parameters(const ParameterFile &prm)
      {
        const std::string s = "linear solver";
        verbosity           = parse_verbosity(prm.get(s, "verbosity", "verbose"));
        const std::string m = prm.get(s, "method", "gmres");
        if (m == "gmres")
          solver = SolverType::gmres;
        else if (m == "bicgstab")
          solver = SolverType::bicgstab;
        else if (m == "amg")
          solver = SolverType::amg;
        else
          throw std::runtime_error(
            "Error, invalid iterative solver type. Choices are amg, gmres or bicgstab");
        residual_precision = (unsigned)prm.get_integer(s, "residual precision", 6);
        relative_residual  = prm.get_double(s, "relative residual", 1e-3);
        minimum_residual   = prm.get_double(s, "minimum residual", 1e-8);
        max_iterations     = (int)prm.get_integer(s, "max iters", 1000);
        ilu_precond_fill   = prm.get_double(s, "ilu preconditioner fill", 0);
        ilu_precond_atol   = prm.get_double(s, "ilu preconditioner absolute tolerance", 1e-8);
        ilu_precond_rtol   = prm.get_double(s, "ilu preconditioner relative tolerance", 1.00);
      }
    };

    // `initial conditions` (include/solvers/initial_conditions.h:36-121).  The function `uvwp` is
    // evaluated on the host (deal.II's ParsedFunction); the solver takes its values.
    enum class InitialConditionType
    {
      none,
      L2projection,
      viscous,
      nodal
    };
    struct InitialConditions
    {
      InitialConditionType type      = InitialConditionType::nodal;
      double               viscosity = 1;
      void
      parse_parameters(const ParameterFile &prm)
      {
        const std::string op = prm.get("initial conditions", "type", "nodal");
        if (op == "L2projection")
          type = InitialConditionType::L2projection;
        else if (op == "viscous")
          type = InitialConditionType::viscous;
        else if (op == "nodal")
          type = InitialConditionType::nodal;
        else // Patterns::Selection("L2projection|viscous|nodal") rejects anything else
          throw std::runtime_error("initial conditions: type must be one of <L2projection|viscous|nodal>");
        viscosity = prm.get_double("initial conditions", "viscosity", 1);
      }
    };

    // `FEM`, `physical properties`, `velocity source` (parameters.cc:169-210, 803-831)
    struct FEM
    {
      unsigned int velocity_order = 1, pressure_order = 1, quadrature_points = 0;
      void
      parse_parameters(const ParameterFile &prm)
      {
        velocity_order    = (unsigned)prm.get_integer("FEM", "velocity order", 1);
        pressure_order    = (unsigned)prm.get_integer("FEM", "pressure order", 1);
        quadrature_points = (unsigned)prm.get_integer("FEM", "quadrature points", 0);
      }
    };
    struct PhysicalProperties
    {
      double viscosity = 1;
      void
      parse_parameters(const ParameterFile &prm)
      {
        viscosity = prm.get_double("physical properties", "kinematic viscosity", 1);
      }
    };
    struct VelocitySource
    {
      enum class VelocitySourceType
      {
        none,
        srf
      } type         = VelocitySourceType::none;
      double omega_x = 0, omega_y = 0, omega_z = 0;
      void
      parse_parameters(const ParameterFile &prm)
      {
        const std::string op = prm.get("velocity source", "type", "none");
        if (op == "none")
          type = VelocitySourceType::none;
        else if (op == "srf")
          type = VelocitySourceType::srf;
        else
          throw std::runtime_error("Error, invalid velocity source type");
        omega_x = prm.get_double("velocity source", "omega_x", 0);
        omega_y = prm.get_double("velocity source", "omega_y", 0);
        omega_z = prm.get_double("velocity source", "omega_z", 0);
      }
    };
  } // namespace Parameters

  using TimeSteppingMethod = Parameters::SimulationControl::TimeSteppingMethod;

  // ------------------------------------------------------------------------------------------
  // Host vector with the operations the Newton drivers use on TrilinosWrappers::MPI::Vector
  // ------------------------------------------------------------------------------------------
  class Vector
  {
  public:
    Vector() = default;
    explicit Vector(std::size_t n)
      : v(n, 0.0)
    {}
    void
    reinit(std::size_t n)
    {
      v.assign(n, 0.0);
    }
    std::size_t
    size() const
    {
      return v.size();
    }
    double &
    operator[](std::size_t i)
    {
      return v[i];
    }
    double
    operator[](std::size_t i) const
    {
      return v[i];
    }
    double *
    data()
    {
      return v.data();
    }
    const double *
    data() const
    {
      return v.data();
    }
    Vector &
    operator=(double s)
    {
      for (auto &x : v)
        x = s;
      return *this;
    }
    void
    add(double a, const Vector &o)
    {
      for (std::size_t i = 0; i < v.size(); ++i)
        v[i] += a * o.v[i];
    }
    double
    l2_norm() const
    {
      if (cached_norm >= 0)
        return cached_norm;
      double s = 0;
      for (double x : v)
        s += x * x;
      return std::sqrt(s);
    }
    // the device computes system_rhs.l2_norm(); the mirror can carry it along with the data
    double cached_norm = -1;

  private:
    std::vector<double> v;
  };

  // what deal.II throws from SolverGMRES::solve when max iters is reached
  class NoConvergence : public std::runtime_error
  {
  public:
    NoConvergence(unsigned int last_step, double last_residual)
      : std::runtime_error("Iterative method reported convergence failure in step " +
                           std::to_string(last_step) + ". The residual in the last step was " +
                           std::to_string(last_residual) + ".")
      , last_step(last_step)
      , last_residual(last_residual)
    {}
    unsigned int last_step;
    double       last_residual;
  };

  class ConditionalOStream
  {
  public:
    explicit ConditionalOStream(std::ostream &o = std::cout, bool active = true)
      : out(&o)
      , active(active)
    {}
    template <typename T>
    const ConditionalOStream &
    operator<<(const T &t) const
    {
      if (active)
        *out << t;
      return *this;
    }
    const ConditionalOStream &
    operator<<(std::ostream &(*p)(std::ostream &)) const
    {
      if (active)
        *out << p;
      return *this;
    }
    void
    set_stream(std::ostream &o)
    {
      out = &o;
    }

  private:
    std::ostream *out;
    bool          active;
  };

  template <typename VectorType>
  class NonLinearSolver;

  // ------------------------------------------------------------------------------------------
  // PhysicsSolver: the plug-in boundary (include/core/physics_solver.h:38-158)
  // ------------------------------------------------------------------------------------------
  template <typename VectorType>
  class PhysicsSolver
  {
  public:
    // The Newton drivers (include/core/newton_non_linear_solver.h, skip_newton_non_linear_solver.h)
    // are the reference's own and stay with it: the driver is handed in (the reference's
    // constructor picks it from Parameters::NonLinearSolver, physics_solver.h:126-145) and
    // owned, as there (physics_solver.h:54-57).
    explicit PhysicsSolver(NonLinearSolver<VectorType> *non_linear_solver = nullptr)
      : non_linear_solver(non_linear_solver)
    {}
    virtual ~PhysicsSolver();

    virtual void
    assemble_matrix_and_rhs(const TimeSteppingMethod time_stepping_method) = 0;
    virtual void
    assemble_rhs(const TimeSteppingMethod time_stepping_method) = 0;
    virtual void
    solve_linear_system(const bool initial_step, const bool renewed_matrix = true) = 0;

    void
    solve_non_linear_system(const TimeSteppingMethod time_stepping_method,
                            const bool first_iteration, const bool force_matrix_renewal)
    {
      if (!non_linear_solver)
        throw std::runtime_error("PhysicsSolver: no non-linear solver attached");
      non_linear_solver->solve(time_stepping_method, first_iteration, force_matrix_renewal);
    }

    // nonzero_constraints.distribute(local_evaluation_point)
    virtual void
    apply_constraints()
    {
      for (std::size_t i = 0; i < constrained.size(); ++i)
        if (constrained[i])
          local_evaluation_point[i] = constraint_values.empty() ? 0.0 : constraint_values[i];
    }

    NonLinearSolver<VectorType> *non_linear_solver = nullptr;
    VectorType                   system_rhs;
    VectorType                   evaluation_point;
    VectorType                   local_evaluation_point;
    VectorType                   present_solution;
    VectorType                   newton_update;
    // AffineConstraints stand-in: Dirichlet lines of nonzero_constraints
    std::vector<unsigned char> constrained;
    std::vector<double>        constraint_values;
    ConditionalOStream         pcout;
  };

  template <typename VectorType>
  class NonLinearSolver
  {
  public:
    NonLinearSolver(PhysicsSolver<VectorType> *physics_solver, const Parameters::NonLinearSolver &params)
      : physics_solver(physics_solver)
      , params(params)
    {}
    virtual ~NonLinearSolver() = default;
    virtual void
    solve(const TimeSteppingMethod time_stepping_method, const bool is_initial_step,
          const bool force_matrix_renewal = true) = 0;

  protected:
    PhysicsSolver<VectorType> *physics_solver;
    Parameters::NonLinearSolver params;
  };

  template <typename VectorType>
  PhysicsSolver<VectorType>::~PhysicsSolver()
  {
    delete non_linear_solver;
  }

  // ------------------------------------------------------------------------------------------
  // GLSNavierStokesSolver: the three overrides forwarded to the device through the C ABI
  // ------------------------------------------------------------------------------------------
  struct NavierStokesSolverParameters
  {
    Parameters::NonLinearSolver    non_linear_solver;
    Parameters::LinearSolver       linear_solver;
    Parameters::FEM                fem_parameters;
    Parameters::PhysicalProperties physical_properties;
    Parameters::VelocitySource     velocitySource;
    Parameters::InitialConditions  initial_condition;
    void
    parse(const ParameterFile &prm)
    {
      initial_condition.parse_parameters(prm);
      non_linear_solver.parse_parameters(prm);
      linear_solver.parse_parameters(prm);
      fem_parameters.parse_parameters(prm);
      physical_properties.parse_parameters(prm);
      velocitySource.parse_parameters(prm);
    }
  };

  class GLSNavierStokesSolver : public PhysicsSolver<Vector>
  {
  public:
    // fe / mesh: the host arrays setup_dofs produced (borrowed during construction only)
    GLSNavierStokesSolver(const NavierStokesSolverParameters &nsparam, const glsns_fe_desc &fe,
                          const glsns_mesh_desc &mesh, const double *forcing_at_q = nullptr,
                          int cuda_device = 0)
      : PhysicsSolver<Vector>(nullptr)
      , nsparam(nsparam)
      , n_dofs(mesh.n_dofs)
      , n_owned(mesh.n_owned)
    {
      check(glsns_create(cuda_device, &ctx), "glsns_create");
      check(glsns_set_fe(ctx, &fe), "glsns_set_fe");
      check(glsns_set_mesh(ctx, &mesh), "glsns_set_mesh");
      set_physics();
      check(glsns_set_forcing(ctx, forcing_at_q), "glsns_set_forcing");
      constrained.assign(mesh.constrained, mesh.constrained + mesh.n_dofs);
      if (mesh.constraint_values)
        constraint_values.assign(mesh.constraint_values, mesh.constraint_values + mesh.n_dofs);
      system_rhs.reinit(n_owned);
      newton_update.reinit(n_owned);
      evaluation_point.reinit(n_dofs);
      local_evaluation_point.reinit(n_owned);
      present_solution.reinit(n_dofs);
      solution_m1.reinit(n_dofs), solution_m2.reinit(n_dofs), solution_m3.reinit(n_dofs);
    }
    ~GLSNavierStokesSolver() override
    {
      glsns_destroy(ctx);
    }
    GLSNavierStokesSolver(const GLSNavierStokesSolver &) = delete;

    // gls_navier_stokes.cc:916-1022
    void
    assemble_matrix_and_rhs(const TimeSteppingMethod time_stepping_method) override
    {
      assembleGLS(true, time_stepping_method);
    }
    // gls_navier_stokes.cc:1023-1128
    void
    assemble_rhs(const TimeSteppingMethod time_stepping_method) override
    {
      assembleGLS(false, time_stepping_method);
    }
    // gls_navier_stokes.cc:1130-1159
    void
    solve_linear_system(const bool initial_step, const bool renewed_matrix = true) override
    {
      const double absolute_residual = nsparam.linear_solver.minimum_residual;
      const double relative_residual = nsparam.linear_solver.relative_residual;
      if (nsparam.linear_solver.solver == Parameters::LinearSolver::SolverType::gmres)
        solve_system_GMRES(initial_step, absolute_residual, relative_residual, renewed_matrix);
      else if (nsparam.linear_solver.solver == Parameters::LinearSolver::SolverType::bicgstab)
        solve_system_BiCGStab(initial_step, absolute_residual, relative_residual, renewed_matrix);
      else
        // amg (Trilinos ML) is not built on the device; the reference's own message for an
        // unknown method (:1158)
        throw std::runtime_error("This solver is not allowed");
    }

    // set_initial_condition(Parameters::InitialConditionType::L2projection)
    // (gls_navier_stokes.cc:795-803): assemble_L2_projection(); solve_system_GMRES(true, 1e-15,
    // 1e-15, true); present_solution = newton_update.  initial_at_q: the initial-condition
    // function (u, p) at the quadrature points, [n_cells][n_q][dim+1] (what
    // initial_condition->uvwp.vector_value_list delivers, :871-872).  finish_time_step() and
    // postprocess() stay with the caller.
    void
    set_initial_condition_L2projection(const double *initial_at_q)
    {
      check(glsns_assemble_l2_projection(ctx, initial_at_q), "assemble_L2_projection");
      check(glsns_get_vector(ctx, GLSNS_VEC_SYSTEM_RHS, system_rhs.data(), n_owned), "get rhs");
      system_rhs.cached_norm = -1;
      solve_system_GMRES(true, 1e-15, 1e-15, true);
      for (int64_t i = 0; i < n_owned; ++i)
        present_solution[i] = newton_update[i];
      check(glsns_set_vector(ctx, GLSNS_VEC_PRESENT_SOLUTION, present_solution.data(), n_dofs),
            "set present_solution");
      check(glsns_update_ghosts(ctx, GLSNS_VEC_PRESENT_SOLUTION), "ghost import");
      check(glsns_get_vector(ctx, GLSNS_VEC_PRESENT_SOLUTION, present_solution.data(), n_dofs),
            "get present_solution");
    }

    // set_nodal_values (source/solvers/navier_stokes_base.cc:926-944): VectorTools::interpolate
    // of the initial-condition function (done by the host: `initial_nodal`, [n_dofs]), then
    // nonzero_constraints.distribute, present_solution = newton_update.
    void
    set_nodal_values(const double *initial_nodal)
    {
      for (int64_t i = 0; i < n_dofs; ++i)
        present_solution[i] =
          (constrained[i] && !constraint_values.empty()) ? constraint_values[i] :
          constrained[i]                                  ? 0.0 :
                                                            initial_nodal[i];
      for (int64_t i = 0; i < n_owned; ++i)
        newton_update[i] = present_solution[i];
    }

    // set_initial_condition(initial_condition_type, restart = false)
    // (gls_navier_stokes.cc:784-828); finish_time_step() / postprocess() stay with the caller,
    // restart (read_checkpoint) is out of scope.
    void
    set_initial_condition(const Parameters::InitialConditionType initial_condition_type,
                          const double *initial_nodal, const double *initial_at_q)
    {
      if (initial_condition_type == Parameters::InitialConditionType::L2projection)
        set_initial_condition_L2projection(initial_at_q);
      else if (initial_condition_type == Parameters::InitialConditionType::nodal)
        set_nodal_values(initial_nodal);
      else if (initial_condition_type == Parameters::InitialConditionType::viscous)
        { // a steady solve at the artificial viscosity of the `initial conditions` subsection
          set_nodal_values(initial_nodal);
          const double viscosity                = nsparam.physical_properties.viscosity;
          nsparam.physical_properties.viscosity = nsparam.initial_condition.viscosity;
          set_physics();
          try
            {
              PhysicsSolver<Vector>::solve_non_linear_system(TimeSteppingMethod::steady, false, true);
            }
          catch (...)
            {
              nsparam.physical_properties.viscosity = viscosity;
              set_physics();
              throw;
            }
          nsparam.physical_properties.viscosity = viscosity;
          set_physics();
        }
      else
        throw std::runtime_error("GLSNS - Initial condition could not be set"); // :825
    }

    // calculate_CFL(dof_handler, present_solution, fem_parameters, time_step, communicator)
    // (source/solvers/postprocessing_cfl.cc:34-87; called from navier_stokes_base.cc:436-441)
    double
    calculate_CFL(const double *shape_u_at_centre, const double time_step)
    {
      check(glsns_set_vector(ctx, GLSNS_VEC_PRESENT_SOLUTION, present_solution.data(), n_dofs),
            "set present_solution");
      double    cfl = 0;
      const int deg = (int)std::max(nsparam.fem_parameters.velocity_order,
                                    nsparam.fem_parameters.pressure_order);
      check(glsns_calculate_cfl(ctx, GLSNS_VEC_PRESENT_SOLUTION, shape_u_at_centre, deg, time_step,
                                &cfl),
            "calculate_CFL");
      return cfl;
    }

    // state the time-stepping glue sets (navier_stokes_base.h:291-293, simulation_control.h:239)
    Vector              solution_m1, solution_m2, solution_m3;
    std::vector<double> time_steps_vector = {1.0, 1.0, 1.0, 1.0};
    glsns_solve_info    last_solve        = {0, 0, 0, 0};

    glsns_context *
    context()
    {
      return ctx;
    }
    const NavierStokesSolverParameters &
    parameters() const
    {
      return nsparam;
    }

  private:
    // viscosity and velocity source of nsparam -> the device context
    void
    set_physics()
    {
      const double omega[3] = {nsparam.velocitySource.omega_x, nsparam.velocitySource.omega_y,
                               nsparam.velocitySource.omega_z};
      check(glsns_set_physics(ctx, nsparam.physical_properties.viscosity,
                              nsparam.velocitySource.type ==
                                  Parameters::VelocitySource::VelocitySourceType::srf ?
                                GLSNS_SOURCE_SRF :
                                GLSNS_SOURCE_NONE,
                              omega),
            "glsns_set_physics");
    }

    void
    check(glsns_status s, const char *what) const
    {
      if (s != GLSNS_OK)
        throw std::runtime_error(std::string(what) + ": " + (ctx ? glsns_last_error(ctx) : "failed"));
    }

    static glsns_scheme
    to_abi(const TimeSteppingMethod m)
    {
      return static_cast<glsns_scheme>(static_cast<int>(m)); // same enumerator order
    }

    // assembleGLS<assemble_matrix, scheme, velocity_source> (:231-777) on the device
    void
    assembleGLS(const bool assemble_matrix, const TimeSteppingMethod scheme)
    {
      check(glsns_set_vector(ctx, GLSNS_VEC_EVALUATION_POINT, evaluation_point.data(), n_dofs),
            "set evaluation_point");
      if (scheme != TimeSteppingMethod::steady)
        {
          check(glsns_set_vector(ctx, GLSNS_VEC_SOLUTION_M1, solution_m1.data(), n_dofs), "m1");
          check(glsns_set_vector(ctx, GLSNS_VEC_SOLUTION_M2, solution_m2.data(), n_dofs), "m2");
          check(glsns_set_vector(ctx, GLSNS_VEC_SOLUTION_M3, solution_m3.data(), n_dofs), "m3");
        }
      check(glsns_assemble(ctx, assemble_matrix ? 1 : 0, to_abi(scheme), time_steps_vector.data()),
            "glsns_assemble");
      check(glsns_get_vector(ctx, GLSNS_VEC_SYSTEM_RHS, system_rhs.data(), n_owned), "get rhs");
      system_rhs.cached_norm = -1;
    }

    // setup_ILU (:1161-1176)
    void
    setup_ILU()
    {
      check(glsns_setup_ilu(ctx, (int)nsparam.linear_solver.ilu_precond_fill,
                            nsparam.linear_solver.ilu_precond_atol,
                            nsparam.linear_solver.ilu_precond_rtol),
            "setup_ILU");
    }

    // solve_system_GMRES (:1242-1289)
    void
    solve_system_GMRES(const bool initial_step, const double absolute_residual,
                       const double relative_residual, const bool renewed_matrix)
    {
      solve_system_Krylov(GLSNS_SOLVER_GMRES, "solve_system_GMRES", initial_step,
                          absolute_residual, relative_residual, renewed_matrix);
    }

    // solve_system_BiCGStab (:1291-1340)
    void
    solve_system_BiCGStab(const bool initial_step, const double absolute_residual,
                          const double relative_residual, const bool renewed_matrix)
    {
      solve_system_Krylov(GLSNS_SOLVER_BICGSTAB, "solve_system_BiCGStab", initial_step,
                          absolute_residual, relative_residual, renewed_matrix);
    }

    // the body the two share in the reference (tolerance rule, ILU renewal, printing, distribute)
    void
    solve_system_Krylov(const glsns_solver_method method, const char *what,
                        const bool initial_step, const double absolute_residual,
                        const double relative_residual, const bool renewed_matrix)
    {
      const double linear_solver_tolerance =
        std::max(relative_residual * system_rhs.l2_norm(), absolute_residual);
      if (nsparam.linear_solver.verbosity != Parameters::Verbosity::quiet)
        pcout << "  -Tolerance of iterative solver is : "
              << std::setprecision(nsparam.linear_solver.residual_precision)
              << linear_solver_tolerance << std::endl;
      if (renewed_matrix || !have_ilu)
        {
          setup_ILU();
          have_ilu = true;
        }
      glsns_linear_solver_params p;
      p.relative_residual = relative_residual;
      p.minimum_residual  = absolute_residual;
      p.max_iterations    = nsparam.linear_solver.max_iterations;
      p.restart           = 30;
      p.ilu_fill          = (int)nsparam.linear_solver.ilu_precond_fill;
      p.ilu_atol          = nsparam.linear_solver.ilu_precond_atol;
      p.ilu_rtol          = nsparam.linear_solver.ilu_precond_rtol;
      p.method            = (int32_t)method;
      const glsns_status s =
        glsns_solve_linear_system(ctx, &p, 0, newton_update.data(), &last_solve);
      if (s == GLSNS_ERR_NO_CONVERGENCE)
        throw NoConvergence(last_solve.iterations, last_solve.true_residual);
      check(s, what);
      if (initial_step && !constraint_values.empty()) // nonzero_constraints.distribute (:1287)
        for (int64_t i = 0; i < n_owned; ++i)
          if (constrained[i])
            newton_update[i] = constraint_values[i];
      if (nsparam.linear_solver.verbosity != Parameters::Verbosity::quiet)
        pcout << "  -Iterative solver took : " << last_solve.iterations << " steps " << std::endl;
    }

    NavierStokesSolverParameters nsparam;
    glsns_context               *ctx = nullptr;
    int64_t                      n_dofs, n_owned;
    bool                         have_ilu = false;
  };
} // namespace glsns

#endif // GLSNS_SOLVER_HPP
