/*
 * glsns.h — C ABI of the B200-native GLS Navier–Stokes hot path.
 *
 * This is the drop-in boundary behind Lethe's GLSNavierStokesSolver<dim>
 * (reference paths relative to the reference root):
 *
 *   assemble_matrix_and_rhs(method)   source/solvers/gls_navier_stokes.cc:916-1022
 *   assemble_rhs(method)              source/solvers/gls_navier_stokes.cc:1023-1128
 *        -> assembleGLS<...>()        source/solvers/gls_navier_stokes.cc:231-777
 *   solve_linear_system(...)          source/solvers/gls_navier_stokes.cc:1130-1159
 *        -> setup_ILU()               source/solvers/gls_navier_stokes.cc:1161-1176
 *        -> solve_system_GMRES(...)   source/solvers/gls_navier_stokes.cc:1242-1289
 *   system_rhs.l2_norm()              include/core/newton_non_linear_solver.h:99,119
 *   the PhysicsSolver vectors         include/core/physics_solver.h:107-111
 *
 * Mesh, refinement, DoF numbering, constraints and FEValues precomputation stay
 * on the host (deal.II); the host hands over plain arrays.  Every pointer in a
 * signature is a HOST pointer that is only borrowed for the duration of the
 * call; the context owns all device memory.  No exception crosses this ABI:
 * every function returns a glsns_status, glsns_last_error() gives the text.
 * A context is not thread safe; one context per GPU (one per MPI rank in the
 * reference's terms); calls are synchronous at return.  In a multi-rank run
 * every rank calls each entry point in the same order (they are collective,
 * like the reference's).
 *
 * Local dof layout of a cell (n = dim*n_su + n_sp entries):
 *   k = c*n_su + a   velocity component c < dim, scalar shape a < n_su
 *   k = dim*n_su + a pressure shape a < n_sp
 * A deal.II adapter fills cell_dofs in this order with
 * fe.system_to_component_index.
 */
#ifndef GLSNS_H
#define GLSNS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct glsns_context glsns_context;

typedef enum
{
  GLSNS_OK                 = 0,
  GLSNS_ERR_BAD_ARGUMENT   = 1,
  GLSNS_ERR_CUDA           = 2,
  /* GMRES hit max iters: what deal.II reports by throwing
     SolverControl::NoConvergence */
  GLSNS_ERR_NO_CONVERGENCE = 3,
  GLSNS_ERR_ZERO_PIVOT     = 4, /* ILU breakdown */
  GLSNS_ERR_STATE          = 5, /* call order: e.g. solve before assemble */
  /* std::runtime_error("This solver is not allowed"), gls_navier_stokes.cc:1158,
     and what is refused: curved cells without mapping_laplacian, hanging-node lines on more
     than one rank */
  GLSNS_ERR_UNSUPPORTED    = 6,
  GLSNS_ERR_COMM           = 7
} glsns_status;

/* Parameters::SimulationControl::TimeSteppingMethod, include/core/parameters.h:56-69
   (same order). sdirk2 / sdirk3 are the umbrella values the reference never
   assembles with; they are rejected with GLSNS_ERR_BAD_ARGUMENT. */
typedef enum
{
  GLSNS_STEADY   = 0,
  GLSNS_BDF1     = 1,
  GLSNS_BDF2     = 2,
  GLSNS_BDF3     = 3,
  GLSNS_SDIRK2   = 4,
  GLSNS_SDIRK2_1 = 5,
  GLSNS_SDIRK2_2 = 6,
  GLSNS_SDIRK3   = 7,
  GLSNS_SDIRK3_1 = 8,
  GLSNS_SDIRK3_2 = 9,
  GLSNS_SDIRK3_3 = 10
} glsns_scheme;

/* Parameters::VelocitySource::VelocitySourceType, include/core/parameters.h */
typedef enum
{
  GLSNS_SOURCE_NONE = 0,
  GLSNS_SOURCE_SRF  = 1
} glsns_velocity_source;

/* The PhysicsSolver / NavierStokesBase vectors (include/core/physics_solver.h:107-111,
   include/solvers/navier_stokes_base.h:291-293). */
typedef enum
{
  GLSNS_VEC_EVALUATION_POINT = 0, /* in : ghosted linearisation point          */
  GLSNS_VEC_SOLUTION_M1      = 1, /* in : u^n     (transient schemes)          */
  GLSNS_VEC_SOLUTION_M2      = 2, /* in : u^(n-1) / SDIRK stage                */
  GLSNS_VEC_SOLUTION_M3      = 3, /* in : u^(n-2) / SDIRK stage                */
  GLSNS_VEC_SYSTEM_RHS       = 4, /* out: owned residual                       */
  GLSNS_VEC_NEWTON_UPDATE    = 5, /* out: owned GMRES solution (constrained=0) */
  GLSNS_VEC_PRESENT_SOLUTION = 6  /* in/out: used by glsns_line_search_point   */
} glsns_vector;

/* FESystem(FE_Q(pu)^dim, FE_Q(pp)) evaluated by FEValues on the REFERENCE cell
   at the n_q points of QGauss (gls_navier_stokes.cc:244-252). Row-major. */
typedef struct
{
  int32_t       dim;             /* 2 or 3                                     */
  int32_t       velocity_degree; /* divides h, gls_navier_stokes.cc:340-345    */
  int32_t       n_su;            /* scalar shape functions per velocity comp.  */
  int32_t       n_sp;            /* pressure shape functions                   */
  int32_t       n_q;
  const double *shape_u;         /* [n_q][n_su]                                */
  const double *grad_u;          /* [n_q][n_su][dim]      d/dxi                */
  const double *hess_u;          /* [n_q][n_su][dim][dim] d2/dxi dxi           */
  const double *shape_p;         /* [n_q][n_sp]                                */
  const double *grad_p;          /* [n_q][n_sp][dim]                           */
  const double *weights;         /* [n_q] weights on the unit cell             */
} glsns_fe_desc;

/* What setup_dofs (gls_navier_stokes.cc:57-228) produces, restricted to this
   rank.  Local dof indices: owned dofs 0..n_owned-1 (the rank's contiguous
   locally_owned range, in global order), then ghosts n_owned..n_dofs-1 grouped
   by owner rank.  Serial: n_owned = n_dofs. */
typedef struct
{
  int64_t        n_dofs;
  int64_t        n_owned;
  int64_t        n_cells;        /* cells this rank assembles: every cell that
                                    touches an owned row (ghost-layer cells are
                                    assembled redundantly; only owned rows are
                                    written, so no compress(add) exchange)      */
  const int32_t *cell_dofs;      /* [n_cells][n] local dof indices              */
  /* geometry.  geometry_per_q = 0: affine cells (parallelograms / parallelepipeds), one
     inverse Jacobian and one determinant per cell.  geometry_per_q = 1: per (cell,q)
     entries (MappingQ, `qmapping all`, or Q1 mapping of a non-parallelogram cell); then
     mapping_laplacian (below) is REQUIRED: without it the real-space Laplacians of the
     shape functions that FEValues delivers with update_hessians
     (gls_navier_stokes.cc:417-422) cannot be formed, and glsns_set_mesh returns
     GLSNS_ERR_UNSUPPORTED rather than assemble wrong strong residuals. */
  int32_t        geometry_per_q;
  const double  *inv_jacobian;   /* [n_cells]([n_q])[dim][dim], [r][d]=dxi_r/dx_d */
  const double  *det_jacobian;   /* [n_cells]([n_q]); JxW = det * weight        */
  const double  *cell_measure;   /* [n_cells] cell->measure()                   */
  const double  *q_points;       /* [n_cells][n_q][dim] or NULL (needed for SRF) */
  /* zero_constraints: 1 = homogeneous-Dirichlet-constrained dof; 2 = hanging-node line
     (then constraint_ptr/idx/weight below hold its entries).                           */
  const uint8_t *constrained;    /* [n_dofs]                                    */
  /* nonzero_constraints inhomogeneities (values the constrained dofs take in
     apply_constraints, physics_solver.h:98-102); NULL = all zero.             */
  const double  *constraint_values; /* [n_dofs] or NULL                         */
  /* sparsity (gls_navier_stokes.cc:204-213, keep_constrained_dofs = false):
     CSR of the owned rows, columns are local dof indices, sorted per row.     */
  const int64_t *row_ptr;        /* [n_owned+1]                                 */
  const int32_t *col_idx;        /* [row_ptr[n_owned]]                          */
  /* cell colouring for the atomic-free scatter: cells of one colour share no
     dof. color_cells lists cell indices colour by colour.                     */
  int32_t        n_colors;
  const int32_t *color_ptr;      /* [n_colors+1]                                */
  const int32_t *color_cells;    /* [n_cells]                                   */
  /* halo (multi-rank only; n_neighbors = 0 in serial) */
  int32_t        n_neighbors;
  const int32_t *neighbor_rank;  /* [n_neighbors]                               */
  const int64_t *send_ptr;       /* [n_neighbors+1] into send_idx               */
  const int32_t *send_idx;       /* owned local indices to send                 */
  const int64_t *recv_ptr;       /* [n_neighbors+1]; ghosts of neighbour i are
                                    n_owned+recv_ptr[i] .. n_owned+recv_ptr[i+1] */
  /* geometry_per_q = 1 only: [n_cells][n_q][dim], the mapping's second derivatives
     contracted with the metric,
        c_k = sum_{r,s} d2x_k/dxi_r dxi_s (J^-1 J^-T)_{rs}
     (from fe_values.jacobian_grad(q) and inverse_jacobian(q), update_jacobian_grads).
     The real-space Laplacian of a shape function is then
        lap N = H_ref(N) : (J^-1 J^-T) - grad_x N . c
     which is what trace(fe_values[velocities].hessian(k, q)) is on a curved cell.
     NULL when geometry_per_q = 0. */
  const double  *mapping_laplacian;
  /* Hanging-node lines (DoFTools::make_hanging_node_constraints in setup_dofs; Kelly-refined
     meshes, navier_stokes_base.cc:610-729), in the CLOSED form of zero_constraints: dof i with
     constrained[i] == 2 is x_i = sum_k constraint_weight[k] x_{constraint_idx[k]},
     k in [constraint_ptr[i], constraint_ptr[i+1]); the masters are unconstrained dofs (Dirichlet
     masters have dropped out; a line may be empty).  constraint_inhomogeneity[i] is what the
     Dirichlet masters contribute in nonzero_constraints (apply_constraints,
     physics_solver.h:98-102), NULL = 0.  The scatter resolves the lines as
     AffineConstraints::distribute_local_to_global does (gls_navier_stokes.cc:755-771); the solver
     distributes them after the solve (:1287).  The cell colouring must then separate cells that
     share a MASTER, not only a dof.  All NULL when there are no hanging nodes.  One rank only
     (GLSNS_ERR_UNSUPPORTED with n_ranks > 1); glsns_assemble_l2_projection does not take them.  */
  const int64_t *constraint_ptr;          /* [n_dofs+1] or NULL                 */
  const int32_t *constraint_idx;
  const double  *constraint_weight;
  const double  *constraint_inhomogeneity; /* [n_dofs] or NULL                  */
} glsns_mesh_desc;

/* `linear solver` subsection (source/core/parameters.cc:507-559). */
typedef enum
{
  GLSNS_SOLVER_GMRES    = 0, /* method = gmres:    solve_system_GMRES,    gls_navier_stokes.cc:1242-1289 */
  GLSNS_SOLVER_BICGSTAB = 1  /* method = bicgstab: solve_system_BiCGStab, gls_navier_stokes.cc:1291-1340 */
  /* method = amg (solve_system_AMG, Trilinos ML) is not built: the host mirror throws the
     reference's "This solver is not allowed" */
} glsns_solver_method;

typedef struct
{
  double  relative_residual; /* default 1e-3  */
  double  minimum_residual;  /* default 1e-8  */
  int32_t max_iterations;    /* default 1000  */
  int32_t restart;           /* deal.II SolverGMRES default: 30 */
  int32_t ilu_fill;          /* default 0; k > 0: level-of-fill ILU(k) as Ifpack builds it */
  double  ilu_atol;          /* default 1e-8  */
  double  ilu_rtol;          /* default 1.0   */
  int32_t method;            /* glsns_solver_method, default GLSNS_SOLVER_GMRES */
} glsns_linear_solver_params;

typedef struct
{
  int32_t iterations;      /* solver_control.last_step()                        */
  double  tolerance;       /* max(rel*||rhs||, abs), gls_navier_stokes.cc:1251  */
  double  true_residual;   /* ||b - A x||_2 (what SolverControl logs)           */
  double  estimated_residual; /* |g_{j+1}| of the Givens recurrence             */
} glsns_solve_info;

/* Accumulated device time per phase, named after the reference's TimerOutput
   sections, plus the kernels inside GMRES. Milliseconds. */
typedef struct
{
  double  assemble_system_ms, assemble_rhs_ms, setup_ilu_ms, solve_linear_system_ms;
  double  spmv_ms, trsv_ms, orthog_ms;
  int64_t assemble_system_calls, assemble_rhs_calls, setup_ilu_calls, solve_calls;
  int64_t spmv_calls, trsv_calls, orthog_calls;
  int64_t kernel_launches;   /* every kernel this library launched */
} glsns_timers;

/* ---- life cycle ----------------------------------------------------------- */
glsns_status glsns_create(int32_t cuda_device, glsns_context **out);
void         glsns_destroy(glsns_context *ctx);
const char  *glsns_last_error(const glsns_context *ctx);
const char  *glsns_version(void);

/* Multi-rank: one NCCL communicator per context. unique_id is the 128-byte
   ncclUniqueId created by glsns_comm_unique_id on rank 0 and broadcast by the
   host (MPI_Bcast in the reference's world). */
glsns_status glsns_comm_unique_id(uint8_t unique_id[128]);
glsns_status glsns_comm_init(glsns_context *ctx, int32_t n_ranks, int32_t rank,
                             const uint8_t unique_id[128]);

/* ---- setup (after every setup_dofs) --------------------------------------- */
glsns_status glsns_set_fe(glsns_context *ctx, const glsns_fe_desc *fe);
glsns_status glsns_set_mesh(glsns_context *ctx, const glsns_mesh_desc *mesh);
/* `physical properties` / `velocity source` (parameters.cc:169-176, 803-831) */
glsns_status glsns_set_physics(glsns_context *ctx, double kinematic_viscosity,
                               glsns_velocity_source source, const double omega[3]);
/* forcing_function->vector_value_list at the q-points, gls_navier_stokes.cc:364-369.
   [n_cells][n_q][dim]; NULL = NoForce. */
glsns_status glsns_set_forcing(glsns_context *ctx, const double *force_at_q);

/* ---- vectors --------------------------------------------------------------- */
/* n must be n_dofs for the ghosted inputs and n_owned for rhs / update. */
glsns_status glsns_set_vector(glsns_context *ctx, glsns_vector which,
                              const double *host, int64_t n);
glsns_status glsns_get_vector(glsns_context *ctx, glsns_vector which, double *host,
                              int64_t n);

/* ---- hot path -------------------------------------------------------------- */
/* assemble_matrix != 0: assemble_matrix_and_rhs(scheme); else assemble_rhs(scheme).
   time_steps = simulationControl->get_time_steps_vector() (dt_n, dt_n-1, dt_n-2);
   may be NULL for GLSNS_STEADY. */
glsns_status glsns_assemble(glsns_context *ctx, int32_t assemble_matrix,
                            glsns_scheme scheme, const double *time_steps);
/* system_rhs.l2_norm() without downloading the vector (all-reduced over ranks). */
glsns_status glsns_rhs_norm(glsns_context *ctx, double *norm);
/* setup_ILU(): factorises the rank-local diagonal block (Ifpack, overlap 0). */
glsns_status glsns_setup_ilu(glsns_context *ctx, int32_t fill, double atol, double rtol);
/* solve_system_GMRES(): right-preconditioned GMRES(restart), zero initial guess;
   calls glsns_setup_ilu first when renewed_matrix != 0 or no factors exist.
   The constrained entries of the solution are zeroed (zero_constraints.distribute)
   and the result is kept as GLSNS_VEC_NEWTON_UPDATE; newton_update_out may be
   NULL.  Returns GLSNS_ERR_NO_CONVERGENCE at max_iterations (info is filled). */
glsns_status glsns_solve_linear_system(glsns_context *ctx,
                                       const glsns_linear_solver_params *params,
                                       int32_t renewed_matrix, double *newton_update_out,
                                       glsns_solve_info *info);
/* assemble_L2_projection (source/solvers/gls_navier_stokes.cc:829-914): the mass system of
   set_initial_condition(L2projection) into system_matrix / system_rhs, scattered with the
   non-zero constraints.  initial_at_q: [n_cells][n_q][dim+1], the initial-condition function
   (u, p) at the quadrature points (initial_condition->uvwp.vector_value_list, :871-872).
   Follow with glsns_solve_linear_system (the reference uses tolerances 1e-15, :800) and
   glsns_distribute_constraints(GLSNS_VEC_NEWTON_UPDATE). */
glsns_status glsns_assemble_l2_projection(glsns_context *ctx, const double *initial_at_q);
/* nonzero_constraints.distribute(v) (solve_system_GMRES with initial_step = true, :1287):
   constrained entries of a device vector take their constraint values. */
glsns_status glsns_distribute_constraints(glsns_context *ctx, glsns_vector which);
/* calculate_CFL (source/solvers/postprocessing_cfl.cc:34-87): max over the cells of
   |u(cell centre)| / h * time_step, h from the cell measure and fe_degree (= fe.degree of the
   FESystem), all-reduced (max) over the ranks.  `which`: a ghosted vector (present_solution,
   evaluation_point, ...); shape_u_at_centre: [n_su], the scalar velocity shape functions at the
   point of QGauss(1). */
glsns_status glsns_calculate_cfl(glsns_context *ctx, glsns_vector which,
                                 const double *shape_u_at_centre, int32_t fe_degree,
                                 double time_step, double *cfl);
/* Device-resident line-search trial (newton_non_linear_solver.h:113-116):
   evaluation_point = present_solution + alpha * newton_update, then
   apply_constraints (constrained dofs take constraint_values), ghosts updated. */
glsns_status glsns_line_search_point(glsns_context *ctx, double alpha);
/* Ghost import of a ghosted input vector whose owned part was just set (what assigning to a
   ghosted TrilinosWrappers::MPI::Vector does, newton_non_linear_solver.h:93,116). No-op on 1 rank. */
glsns_status glsns_update_ghosts(glsns_context *ctx, glsns_vector which);
/* present_solution = evaluation_point (newton_non_linear_solver.h:135). */
glsns_status glsns_accept_evaluation_point(glsns_context *ctx);

/* ---- inspection (parity tests, benchmarks) -------------------------------- */
glsns_status glsns_get_matrix_values(glsns_context *ctx, double *values, int64_t nnz);
glsns_status glsns_set_matrix_values(glsns_context *ctx, const double *values, int64_t nnz);
/* L\U on the pattern of the installed fill level (glsns_get_ilu_pattern); for fill = 0 that is
   the mesh's own pattern. */
glsns_status glsns_get_ilu_values(glsns_context *ctx, double *values, int64_t nnz);
/* The factor pattern: the mesh's CSR for fill = 0, the level-of-fill pattern of the rank-local
   block (a superset, explicit zeros in the matrix) once glsns_setup_ilu ran with fill > 0.
   Any of nnz / row_ptr [n_owned+1] / col_idx [*nnz] may be NULL. */
glsns_status glsns_get_ilu_pattern(glsns_context *ctx, int64_t *nnz, int64_t *row_ptr,
                                   int32_t *col_idx);
/* y = A x on the device matrix; x is [n_dofs] (ghosted), y is [n_owned]. */
glsns_status glsns_spmv(glsns_context *ctx, const double *x, double *y);
/* z = (LU)^-1 r, both [n_owned]. */
glsns_status glsns_ilu_apply(glsns_context *ctx, const double *r, double *z);
/* glsns_ilu_apply with a trace, for tuning the triangular solves: t_publish[0..n) /
   [n..2n) = device time (ns) at which the lower / upper sweep published each row;
   [2n..3n) / [3n..4n) = lower / upper: per resident warp w, at [8w..8w+7), the clock
   cycles it spent per pipeline stage and its item count (see tools/trsv_trace.py);
   [4n..5n) / [5n..6n) = lower / upper: when each row's totals reached the solver's mailbox;
   [6n..8n), [8n..10n) likewise: when the helper began the last item of the row's group and
   when it had all the inputs (t_publish is [10 n]);
   row_warp (may be NULL) = resident warp each row was scheduled on in the two
   sweeps (bit 30: solved behind its predecessor on the same warp, -1: diagonal row). */
glsns_status glsns_ilu_apply_trace(glsns_context *ctx, const double *r, double *z,
                                   uint64_t *t_publish, int32_t *row_warp);
/* number of dependency levels of the lower / upper triangular solves */
glsns_status glsns_ilu_levels(glsns_context *ctx, int32_t *lower, int32_t *upper);
glsns_status glsns_get_timers(glsns_context *ctx, glsns_timers *out);
glsns_status glsns_reset_timers(glsns_context *ctx);
/* Times `reps` back-to-back launches of one kernel on device-resident data with
   CUDA events on the context's stream (kernel: 0 spmv, 1 ilu apply (L+U),
   2 one CGS2 orthogonalisation against `nvec` basis vectors, 3 matrix assembly,
   4 rhs-only assembly, 5 ilu factorisation). */
glsns_status glsns_time_kernel(glsns_context *ctx, int32_t kernel, int32_t reps,
                               int32_t nvec, double *avg_ms);

#ifdef __cplusplus
}
#endif
#endif /* GLSNS_H */
