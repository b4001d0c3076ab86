// Internal state of a glsns_context (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/glsns.h"

namespace glsns
{
  struct Error
  {
    glsns_status code;
    std::string  msg;
  };

  // Device allocation owned by the context.
  template <typename T>
  struct DevBuf
  {
    T     *p = nullptr;
    size_t n = 0;
    void
    release()
    {
      if (p)
        cudaFree(p);
      p = nullptr;
      n = 0;
    }
  };

  // One item of the chain triangular solve: <= 128 entries of one group (trsv.cu).
  struct TrsvItem
  {
    int64_t rs0;   // CSR offset of the group's first row
    int32_t r0;    // first row
    int32_t len;   // row length (same for every row of the group)
    int32_t e_off; // first entry of this item inside the row
    int32_t flags; // rows in the group (block) | IT_LAST / IT_SOLVER ... | entries << 16
    int32_t nlow;  // offset of the in-group block inside the row
    int32_t fmask; // group descriptor: bit d = couples to the chain row at distance d;
                   // helper item: where the totals go (block position << 4 | row offset)
    int32_t fmask2; // group descriptor: the same for the distances 32 ...
    int32_t pad_;
  };

  // One sweep (lower or upper) of the ILU application in stream form (trsv.cu).
  struct TrsvSweep
  {
    DevBuf<TrsvItem>      items;    // per-warp item lists, concatenated
    DevBuf<TrsvItem>      gdesc;    // the groups of every block (a solver item's rs0 / len index it)
    DevBuf<int64_t>       blob_off; // byte offset of each item's blob in the stream (helper item: its indices)
    DevBuf<int64_t>       blob_off_v; // helper item: byte offset of its values
    DevBuf<int32_t>       next16;   // bytes/16 of the blob NSLOT items ahead (same warp)
    DevBuf<unsigned char> dir;      // per-warp directory
    DevBuf<unsigned char> stream;   // headers, indices and factor values in consumption order
    int64_t               n_items = 0, stream_bytes = 0;
    void
    release()
    {
      items.release(), gdesc.release(), blob_off.release(), blob_off_v.release(), next16.release(), dir.release(), stream.release();
      n_items = stream_bytes = 0;
    }
  };

  struct EventPair
  {
    cudaEvent_t a = nullptr, b = nullptr;
  };

  // One timer class: a pool of event pairs recorded around launches, drained
  // (elapsed times summed) whenever the stream is known to be idle.
  struct KernelTimer
  {
    std::vector<EventPair> pool;
    size_t                 used = 0;
    double                 ms   = 0;
    int64_t                calls = 0;
  };

  enum TimerId
  {
    T_ASSEMBLE_SYSTEM = 0,
    T_ASSEMBLE_RHS,
    T_SETUP_ILU,
    T_SOLVE,
    T_SPMV,
    T_TRSV,
    T_ORTHOG,
    T_COUNT
  };
} // namespace glsns

struct glsns_context
{
  int          device = 0;
  cudaStream_t stream = nullptr;
  std::string  err;

  // ---- finite element tables ----
  bool    have_fe = false;
  int32_t dim = 0, vel_degree = 0, n_su = 0, n_sp = 0, n_q = 0, n_loc = 0;
  glsns::DevBuf<double> shape_u, grad_u, hess_u, shape_p, grad_p, weights;

  // ---- mesh ----
  bool    have_mesh = false;
  int64_t n_dofs = 0, n_owned = 0, n_cells = 0, nnz = 0;
  // ilu preconditioner fill = k > 0: the device pattern is the level-of-fill pattern of the
  // rank-local block (the entries of A plus explicit zeros at the fill positions), so that
  // ILU(k) is ILU(0) on it; the pattern the host handed over (what the inspection calls and
  // the host see) is kept here, `a2p` maps its entries into the padded one
  int32_t               ilu_fill = 0;
  int64_t               nnz_base = 0;
  std::vector<int64_t>  base_rowptr;
  std::vector<int32_t>  base_col;
  glsns::DevBuf<int64_t> a2p;
  int32_t geometry_per_q = 0, n_colors = 0;
  glsns::DevBuf<int32_t> cell_dofs, col, color_cells, order_l, diag_rows;
  glsns::DevBuf<int32_t> grp_first; // per row: first row of its group (-1: diagonal-only row)
  glsns::TrsvSweep                trsv_l, trsv_u;
  int32_t                         trsv_grid = 0;
  std::vector<int32_t>            trsv_row_warp_l, trsv_row_warp_u; // schedule, for the trace
  glsns::DevBuf<int64_t> rowptr, diag_pos;
  glsns::DevBuf<double>  inv_jac, det_jac, measure, q_points, force, cvalues, map_lap;
  glsns::DevBuf<uint8_t> constrained;
  // hanging-node lines (glsns_mesh_desc::constraint_*): CSR over the dofs, the list of the
  // hanging dofs, their inhomogeneities (may be empty)
  int64_t                n_hanging = 0;
  glsns::DevBuf<int64_t> hang_ptr;
  glsns::DevBuf<int32_t> hang_idx, hang_list;
  glsns::DevBuf<double>  hang_w, hang_inhom;
  glsns::DevBuf<int4>    rowdesc; // per row: start, diagonal offset, length | rows left in its group << 16 (ILU factorisation)
  glsns::DevBuf<int2>    fgroups; // (first row, rows) of the row groups, by lower-sweep level
  glsns::DevBuf<int2>    sgroups; // every row in a group (diagonal-only rows too), by row: SpMV
  int64_t                n_sgroups = 0;
  std::vector<int32_t>   color_ptr;
  int32_t                levels_l = 0, levels_u = 0, levels_rows = 0, n_groups = 0;
  int32_t                max_row_len = 0, n_diag_rows = 0;
  double                 avg_row_len = 0;

  // ---- physics ----
  double  viscosity = 1.0;
  int32_t srf       = 0;
  double  omega[3]  = {0, 0, 0};
  bool    have_force = false;

  // ---- matrix / factors / vectors ----
  glsns::DevBuf<double> val, lu, dinv;
  bool                  have_matrix = false, have_ilu = false, have_rhs = false;
  glsns::DevBuf<double> vec[7];
  bool                  vec_set[7] = {false, false, false, false, false, false, false};

  // ---- solver workspace ----
  int32_t               krylov_m = 0;
  glsns::DevBuf<double> V, w, zg, ytmp, tvec, partials, hbuf, ycoef;
  glsns::DevBuf<int32_t> counters; // [0] ticket, [1] error flag
  glsns::DevBuf<int32_t> row_done; // factorisation flags
  int32_t               epoch = 0;
  double               *h_pinned = nullptr; // 3 x 192 doubles: Hessenberg columns of two steps in flight, y
  cudaEvent_t           step_event[2] = {nullptr, nullptr}; // GMRES: a step's column has reached the host
  bool                  gmres_fused = true;     // GLSNS_GMRES_FUSED=0: separate axpy / dot / scale kernels (same bits)
  bool                  gmres_lookahead = true; // GLSNS_GMRES_LOOKAHEAD=0: one step on the stream at a time
  int32_t               n_sm = 148;

  // ---- communication ----
  int32_t n_ranks = 1, rank = 0;
  void   *nccl_comm = nullptr;
  int32_t n_neighbors = 0;
  std::vector<int32_t>   neighbor_rank;
  std::vector<int64_t>   send_ptr, recv_ptr;
  glsns::DevBuf<int32_t> send_idx;
  glsns::DevBuf<double>  send_buf;

  // ---- timers ----
  glsns::KernelTimer timers[glsns::T_COUNT];
  int64_t            kernel_launches = 0;
  bool               timing_enabled  = true;
};

namespace glsns
{
  // error helpers -----------------------------------------------------------
  glsns_status fail(glsns_context *ctx, glsns_status code, const std::string &msg);
  glsns_status cuda_fail(glsns_context *ctx, cudaError_t e, const char *what);

#define GLSNS_CUDA(ctx, call)                                         \
  do                                                                  \
    {                                                                 \
      cudaError_t e__ = (call);                                       \
      if (e__ != cudaSuccess)                                         \
        return glsns::cuda_fail(ctx, e__, #call);                     \
    }                                                                 \
  while (0)

#define GLSNS_TRY(expr)                \
  do                                   \
    {                                  \
      glsns_status s__ = (expr);       \
      if (s__ != GLSNS_OK)             \
        return s__;                    \
    }                                  \
  while (0)

  template <typename T>
  glsns_status
  dev_alloc(glsns_context *ctx, DevBuf<T> &b, size_t n)
  {
    if (b.n == n && b.p)
      return GLSNS_OK;
    b.release();
    if (n == 0)
      return GLSNS_OK;
    cudaError_t e = cudaMalloc((void **)&b.p, n * sizeof(T));
    if (e != cudaSuccess)
      return cuda_fail(ctx, e, "cudaMalloc");
    b.n = n;
    return GLSNS_OK;
  }

  template <typename T>
  glsns_status
  dev_upload(glsns_context *ctx, DevBuf<T> &b, const T *host, size_t n)
  {
    GLSNS_TRY(dev_alloc(ctx, b, n));
    if (n)
      GLSNS_CUDA(ctx, cudaMemcpyAsync(b.p, host, n * sizeof(T), cudaMemcpyHostToDevice,
                                      ctx->stream));
    return GLSNS_OK;
  }

  // timers ------------------------------------------------------------------
  void timer_begin(glsns_context *ctx, TimerId id);
  void timer_end(glsns_context *ctx, TimerId id);
  void timers_drain(glsns_context *ctx); // requires an idle stream

  // kernels (host launchers) ---------------------------------------------------
  // assembly.cu
  glsns_status launch_assembly(glsns_context *ctx, bool assemble_matrix, bool transient,
                               double sdt, const double coefs[4]);
  // sparse.cu
  glsns_status launch_spmv(glsns_context *ctx, const double *x, double *y);
  glsns_status ilu_analyse(glsns_context *ctx, const int64_t *rowptr, const int32_t *col);
  glsns_status ilu_install_fill(glsns_context *ctx, int32_t fill);
  void         iluk_symbolic_host(int64_t n, const int64_t *rowptr, const int32_t *col, int fill,
                                  std::vector<int64_t> &prow, std::vector<int32_t> &pcol);
  glsns_status matrix_values_to_host(glsns_context *ctx, const double *dev_padded, double *host_base);
  glsns_status matrix_values_from_host(glsns_context *ctx, const double *host_base);
  glsns_status launch_ilu_factor(glsns_context *ctx, double atol, double rtol);
  glsns_status launch_ilu_apply(glsns_context *ctx, const double *r, double *z,
                                unsigned long long *trace = nullptr);
  // reads the device error flag (zero pivot / dependency wait timed out); synchronises
  glsns_status check_counters(glsns_context *ctx, const char *what);
  // trsv.cu
  glsns_status trsv_analyse(glsns_context *ctx, const int64_t *rowptr, const int32_t *col,
                            const int64_t *diag);
  glsns_status trsv_prepare(glsns_context *ctx);
  // krylov.cu
  glsns_status ensure_workspace(glsns_context *ctx, int restart);
  glsns_status device_norm2(glsns_context *ctx, const double *x, double *out);
  glsns_status gmres_solve(glsns_context *ctx, const glsns_linear_solver_params *p,
                           glsns_solve_info *info);
  glsns_status bicgstab_solve(glsns_context *ctx, const glsns_linear_solver_params *p,
                           glsns_solve_info *info);
  glsns_status time_orthog(glsns_context *ctx, int nvec);
  glsns_status launch_axpy_constraints(glsns_context *ctx, double alpha);
  glsns_status launch_zero_constrained(glsns_context *ctx, double *x);
  // comm.cu
  glsns_status comm_unique_id(uint8_t out[128]);
  glsns_status comm_init(glsns_context *ctx, int32_t n_ranks, int32_t rank,
                         const uint8_t unique_id[128]);
  void         comm_destroy(glsns_context *ctx);
  glsns_status allreduce_sum(glsns_context *ctx, double *dev, int n);
  glsns_status allreduce_max(glsns_context *ctx, double *dev, int n);
  glsns_status launch_distribute_constraints(glsns_context *ctx, double *x, int64_t n);
  glsns_status launch_l2_projection(glsns_context *ctx, const double *init_dev);
  glsns_status launch_cfl(glsns_context *ctx, const double *shape_centre_dev, const double *U,
                          double dt, double degree, double *out_dev);
  glsns_status halo_exchange(glsns_context *ctx, double *ghosted);
} // namespace glsns
