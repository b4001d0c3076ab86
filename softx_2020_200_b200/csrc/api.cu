// C ABI of the library (include/glsns.h): argument checking, state machine,
// host<->device transfers, timers.  The kernels live in assembly.cu, sparse.cu,
// krylov.cu and comm.cu.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "context.h"

namespace glsns
{
  glsns_status
  fail(glsns_context *ctx, glsns_status code, const std::string &msg)
  {
    if (ctx)
      ctx->err = msg;
    return code;
  }

  glsns_status
  cuda_fail(glsns_context *ctx, cudaError_t e, const char *what)
  {
    return fail(ctx, GLSNS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }

  void
  timer_begin(glsns_context *ctx, TimerId id)
  {
    if (!ctx->timing_enabled)
      return;
    KernelTimer &t = ctx->timers[id];
    if (t.used == t.pool.size())
      {
        EventPair p;
        cudaEventCreate(&p.a);
        cudaEventCreate(&p.b);
        t.pool.push_back(p);
      }
    cudaEventRecord(t.pool[t.used].a, ctx->stream);
  }

  void
  timer_end(glsns_context *ctx, TimerId id)
  {
    if (!ctx->timing_enabled)
      return;
    KernelTimer &t = ctx->timers[id];
    cudaEventRecord(t.pool[t.used].b, ctx->stream);
    t.used++;
    t.calls++;
  }

  void
  timers_drain(glsns_context *ctx)
  {
    for (int i = 0; i < T_COUNT; ++i)
      {
        KernelTimer &t = ctx->timers[i];
        for (size_t k = 0; k < t.used; ++k)
          {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, t.pool[k].a, t.pool[k].b) == cudaSuccess)
              t.ms += ms;
          }
        t.used = 0;
      }
  }

  namespace
  {
    // Variable-step BDF coefficients: derivative at t_0 of the Lagrange interpolant
    // through t_0 > t_1 > ... > t_p, t_i = -sum_{j<i} dt_j.  Same numbers as the
    // reference's divided-difference recursion (source/core/bdf.cc:46-75, pinned by
    // tests/core/bdf_01.output).
    void
    bdf_coefficients(int p, const double *dt, double *alpha)
    {
      double t[5] = {0, 0, 0, 0, 0};
      for (int i = 1; i <= p; ++i)
        t[i] = t[i - 1] - dt[i - 1];
      alpha[0] = 0;
      for (int k = 1; k <= p; ++k)
        alpha[0] += 1.0 / (t[0] - t[k]);
      for (int i = 1; i <= p; ++i)
        {
          double num = 1, den = 1;
          for (int k = 0; k <= p; ++k)
            {
              if (k != i && k != 0)
                num *= t[0] - t[k];
              if (k != i)
                den *= t[i] - t[k];
            }
          alpha[i] = num / den;
        }
    }

    // (transient, 1/dt, c[4]) of a scheme: the compile-time dispatch of
    // assemble_matrix_and_rhs / assemble_rhs (gls_navier_stokes.cc:916-1128) and the
    // coefficient set-up of assembleGLS (:295-329); SDIRK tables from
    // source/core/sdirk.cc:11-44.
    glsns_status
    scheme_coefficients(glsns_context *ctx, glsns_scheme scheme, const double *dts,
                        bool &transient, double &sdt, double c[4])
    {
      c[0] = c[1] = c[2] = c[3] = 0;
      transient = scheme != GLSNS_STEADY;
      sdt       = 0;
      if (!transient)
        return GLSNS_OK;
      if (!dts || !(dts[0] > 0))
        return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "transient scheme needs time_steps[0] > 0");
      sdt = 1.0 / dts[0];
      switch (scheme)
        {
          case GLSNS_BDF1:
            bdf_coefficients(1, dts, c);
            break;
          case GLSNS_BDF2:
            bdf_coefficients(2, dts, c);
            break;
          case GLSNS_BDF3:
            bdf_coefficients(3, dts, c);
            break;
          case GLSNS_SDIRK2_1:
          case GLSNS_SDIRK2_2:
            {
              const double a = (2. - sqrt(2.)) / 2.;
              if (scheme == GLSNS_SDIRK2_1)
                {
                  c[0] = 1. / a * sdt;
                  c[1] = -1. / a * sdt;
                }
              else
                {
                  c[0] = 1. / a * sdt;
                  c[1] = -(2 * a - 1) / a / a * sdt;
                  c[2] = -(1 - a) / a / a * sdt;
                }
              break;
            }
          case GLSNS_SDIRK3_1:
            c[0] = 2.29428036027904 * sdt;
            c[1] = -2.29428036027904 * sdt;
            break;
          case GLSNS_SDIRK3_2:
            c[0] = 2.29428036027904 * sdt;
            c[1] = -0.809559354637498 * sdt;
            c[2] = -1.48472100564154 * sdt;
            break;
          case GLSNS_SDIRK3_3:
            c[0] = 2.29428036027904 * sdt;
            c[1] = 2.87009860433106 * sdt;
            c[2] = -8.55612780155264 * sdt;
            c[3] = 3.39174883694255 * sdt;
            break;
          default:
            // the reference's dispatcher has no branch for the umbrella values
            return fail(ctx, GLSNS_ERR_BAD_ARGUMENT,
                        "scheme is not an assembly stage (sdirk2/sdirk3 umbrella value)");
        }
      return GLSNS_OK;
    }

    bool
    is_ghosted_input(glsns_vector v)
    {
      return v == GLSNS_VEC_EVALUATION_POINT || v == GLSNS_VEC_SOLUTION_M1 ||
             v == GLSNS_VEC_SOLUTION_M2 || v == GLSNS_VEC_SOLUTION_M3 ||
             v == GLSNS_VEC_PRESENT_SOLUTION;
    }
  } // namespace
} // namespace glsns

using namespace glsns;

#define CHECK_CTX(ctx)               \
  if (!(ctx))                        \
    return GLSNS_ERR_BAD_ARGUMENT;   \
  cudaSetDevice((ctx)->device)

extern "C" {

const char *
glsns_version(void)
{
  return "glsns-b200 0.1 (sm_100a)";
}

glsns_status
glsns_create(int32_t cuda_device, glsns_context **out)
{
  if (!out)
    return GLSNS_ERR_BAD_ARGUMENT;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return GLSNS_ERR_CUDA; // no CUDA device: there is no CPU fallback
  if (cuda_device < 0 || cuda_device >= ndev)
    return GLSNS_ERR_BAD_ARGUMENT;
  glsns_context *ctx = new (std::nothrow) glsns_context();
  if (!ctx)
    return GLSNS_ERR_CUDA;
  ctx->device = cuda_device;
  if (cudaSetDevice(cuda_device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess)
    {
      delete ctx;
      return GLSNS_ERR_CUDA;
    }
  int sm = 0;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, cuda_device);
  ctx->n_sm = sm > 0 ? sm : 148;
  if (const char *e = getenv("GLSNS_GMRES_FUSED"))
    ctx->gmres_fused = atoi(e) != 0;
  if (const char *e = getenv("GLSNS_GMRES_LOOKAHEAD"))
    ctx->gmres_lookahead = atoi(e) != 0;
  if (dev_alloc(ctx, ctx->counters, 4) != GLSNS_OK)
    {
      delete ctx;
      return GLSNS_ERR_CUDA;
    }
  *out = ctx;
  return GLSNS_OK;
}

void
glsns_destroy(glsns_context *ctx)
{
  if (!ctx)
    return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_destroy(ctx);
  DevBuf<double> *dbl[] = {&ctx->shape_u, &ctx->grad_u, &ctx->hess_u, &ctx->shape_p,
                           &ctx->grad_p, &ctx->weights, &ctx->inv_jac, &ctx->det_jac,
                           &ctx->measure, &ctx->q_points, &ctx->force, &ctx->cvalues, &ctx->map_lap,
                           &ctx->val, &ctx->lu, &ctx->V, &ctx->w, &ctx->zg, &ctx->ytmp,
                           &ctx->tvec, &ctx->partials, &ctx->hbuf, &ctx->ycoef, &ctx->send_buf};
  for (auto *b : dbl)
    b->release();
  for (auto &v : ctx->vec)
    v.release();
  DevBuf<int32_t> *i32[] = {&ctx->cell_dofs, &ctx->col, &ctx->color_cells, &ctx->order_l,
                            &ctx->counters, &ctx->row_done, &ctx->send_idx};
  for (auto *b : i32)
    b->release();
  ctx->trsv_l.release();
  ctx->trsv_u.release();
  ctx->dinv.release();
  ctx->a2p.release();
  ctx->diag_rows.release();
  ctx->grp_first.release();
  ctx->hang_ptr.release(), ctx->hang_idx.release(), ctx->hang_w.release(), ctx->hang_list.release();
  ctx->hang_inhom.release();
  ctx->fgroups.release();
  ctx->rowdesc.release();
  ctx->sgroups.release();
  ctx->rowptr.release();
  ctx->diag_pos.release();
  ctx->constrained.release();
  if (ctx->h_pinned)
    cudaFreeHost(ctx->h_pinned);
  for (int k = 0; k < 2; ++k)
    if (ctx->step_event[k])
      cudaEventDestroy(ctx->step_event[k]);
  for (auto &t : ctx->timers)
    for (auto &p : t.pool)
      {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
      }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *
glsns_last_error(const glsns_context *ctx)
{
  return ctx ? ctx->err.c_str() : "null context";
}

glsns_status
glsns_comm_unique_id(uint8_t unique_id[128])
{
  if (!unique_id)
    return GLSNS_ERR_BAD_ARGUMENT;
  return comm_unique_id(unique_id);
}

glsns_status
glsns_comm_init(glsns_context *ctx, int32_t n_ranks, int32_t rank, const uint8_t unique_id[128])
{
  CHECK_CTX(ctx);
  if (n_ranks > 1 && !unique_id)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "unique_id is null");
  return comm_init(ctx, n_ranks, rank, unique_id);
}

glsns_status
glsns_set_fe(glsns_context *ctx, const glsns_fe_desc *fe)
{
  CHECK_CTX(ctx);
  if (!fe || (fe->dim != 2 && fe->dim != 3) || fe->n_su < 1 || fe->n_sp < 1 || fe->n_q < 1 ||
      fe->velocity_degree < 1 || !fe->shape_u || !fe->grad_u || !fe->hess_u || !fe->shape_p ||
      !fe->grad_p || !fe->weights)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "bad glsns_fe_desc");
  if (fe->n_q > 256)
    return fail(ctx, GLSNS_ERR_UNSUPPORTED, "more than 256 quadrature points per cell");
  ctx->dim = fe->dim, ctx->vel_degree = fe->velocity_degree;
  ctx->n_su = fe->n_su, ctx->n_sp = fe->n_sp, ctx->n_q = fe->n_q;
  ctx->n_loc     = fe->dim * fe->n_su + fe->n_sp;
  const size_t d = fe->dim, q = fe->n_q, su = fe->n_su, sp = fe->n_sp;
  GLSNS_TRY(dev_upload(ctx, ctx->shape_u, fe->shape_u, q * su));
  GLSNS_TRY(dev_upload(ctx, ctx->grad_u, fe->grad_u, q * su * d));
  GLSNS_TRY(dev_upload(ctx, ctx->hess_u, fe->hess_u, q * su * d * d));
  GLSNS_TRY(dev_upload(ctx, ctx->shape_p, fe->shape_p, q * sp));
  GLSNS_TRY(dev_upload(ctx, ctx->grad_p, fe->grad_p, q * sp * d));
  GLSNS_TRY(dev_upload(ctx, ctx->weights, fe->weights, q));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->have_fe   = true;
  ctx->have_mesh = false; // a mesh is always set against the current element
  return GLSNS_OK;
}

glsns_status
glsns_set_mesh(glsns_context *ctx, const glsns_mesh_desc *m)
{
  CHECK_CTX(ctx);
  if (!ctx->have_fe)
    return fail(ctx, GLSNS_ERR_STATE, "glsns_set_fe must be called before glsns_set_mesh");
  if (!m || m->n_dofs < 0 || m->n_owned < 0 || m->n_owned > m->n_dofs || m->n_cells < 0 ||
      m->n_dofs >= (int64_t)INT32_MAX)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "bad glsns_mesh_desc sizes");
  if (!m->row_ptr || (m->n_cells && (!m->cell_dofs || !m->inv_jacobian || !m->det_jacobian ||
                                     !m->cell_measure || !m->color_ptr || !m->color_cells)) ||
      (m->n_dofs && !m->constrained) || m->n_colors < 0)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "glsns_mesh_desc has a null array");
  if (m->n_colors && m->color_ptr[m->n_colors] != m->n_cells)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "color_ptr does not cover all cells");
  const int64_t nnz = m->row_ptr[m->n_owned];
  if (nnz && !m->col_idx)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "col_idx is null");
  if (m->geometry_per_q && m->n_cells && !m->mapping_laplacian)
    return fail(ctx, GLSNS_ERR_UNSUPPORTED,
                "geometry_per_q = 1 needs mapping_laplacian: without the mapping's second derivatives "
                "the shape-function Laplacians of gls_navier_stokes.cc:417-422 would be wrong on "
                "non-affine cells");
  ctx->have_mesh = ctx->have_matrix = ctx->have_ilu = ctx->have_rhs = false;
  for (bool &b : ctx->vec_set)
    b = false;
  ctx->n_dofs = m->n_dofs, ctx->n_owned = m->n_owned, ctx->n_cells = m->n_cells, ctx->nnz = nnz;
  ctx->nnz_base = nnz, ctx->ilu_fill = 0;
  ctx->base_rowptr.clear(), ctx->base_col.clear();
  ctx->a2p.release();
  ctx->geometry_per_q = m->geometry_per_q ? 1 : 0;
  ctx->n_colors       = m->n_colors;
  ctx->color_ptr.assign(m->color_ptr, m->color_ptr + (m->n_cells ? m->n_colors + 1 : 0));
  if (!m->n_cells)
    ctx->n_colors = 0;
  const size_t nc = (size_t)m->n_cells, d = ctx->dim, q = ctx->n_q;
  const size_t geo = ctx->geometry_per_q ? nc * q : nc;
  GLSNS_TRY(dev_upload(ctx, ctx->cell_dofs, m->cell_dofs, nc * ctx->n_loc));
  GLSNS_TRY(dev_upload(ctx, ctx->inv_jac, m->inv_jacobian, geo * d * d));
  GLSNS_TRY(dev_upload(ctx, ctx->det_jac, m->det_jacobian, geo));
  GLSNS_TRY(dev_upload(ctx, ctx->measure, m->cell_measure, nc));
  if (ctx->geometry_per_q && nc)
    GLSNS_TRY(dev_upload(ctx, ctx->map_lap, m->mapping_laplacian, nc * q * d));
  else
    ctx->map_lap.release();
  if (m->q_points)
    GLSNS_TRY(dev_upload(ctx, ctx->q_points, m->q_points, nc * q * d));
  else
    ctx->q_points.release();
  GLSNS_TRY(dev_upload(ctx, ctx->constrained, m->constrained, (size_t)m->n_dofs));
  { // hanging-node lines
    std::vector<int32_t> hang;
    for (int64_t i = 0; i < m->n_dofs; ++i)
      if (m->constrained[i] == 2)
        hang.push_back((int32_t)i);
    ctx->n_hanging = (int64_t)hang.size();
    ctx->hang_ptr.release(), ctx->hang_idx.release(), ctx->hang_w.release(), ctx->hang_list.release();
    ctx->hang_inhom.release();
    if (ctx->n_hanging)
      {
        if (!m->constraint_ptr || (m->constraint_ptr[m->n_dofs] && (!m->constraint_idx || !m->constraint_weight)))
          return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "constrained[i] == 2 without constraint_ptr / idx / weight");
        if (ctx->n_ranks > 1)
          return fail(ctx, GLSNS_ERR_UNSUPPORTED, "hanging-node constraints on more than one rank");
        const int64_t ne = m->constraint_ptr[m->n_dofs];
        for (int64_t k = 0; k < ne; ++k)
          if (m->constraint_idx[k] < 0 || m->constraint_idx[k] >= m->n_dofs ||
              m->constrained[m->constraint_idx[k]] != 0)
            return fail(ctx, GLSNS_ERR_BAD_ARGUMENT,
                        "a hanging-node line refers to a constrained dof: pass the closed constraints");
        GLSNS_TRY(dev_upload(ctx, ctx->hang_ptr, m->constraint_ptr, (size_t)m->n_dofs + 1));
        GLSNS_TRY(dev_alloc(ctx, ctx->hang_idx, (size_t)ne + 1));
        GLSNS_TRY(dev_alloc(ctx, ctx->hang_w, (size_t)ne + 1));
        if (ne)
          {
            GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->hang_idx.p, m->constraint_idx, sizeof(int32_t) * ne,
                                            cudaMemcpyHostToDevice, ctx->stream));
            GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->hang_w.p, m->constraint_weight, sizeof(double) * ne,
                                            cudaMemcpyHostToDevice, ctx->stream));
          }
        GLSNS_TRY(dev_upload(ctx, ctx->hang_list, hang.data(), hang.size()));
        if (m->constraint_inhomogeneity)
          GLSNS_TRY(dev_upload(ctx, ctx->hang_inhom, m->constraint_inhomogeneity, (size_t)m->n_dofs));
        GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // `hang` goes out of scope
      }
  }
  if (m->constraint_values)
    GLSNS_TRY(dev_upload(ctx, ctx->cvalues, m->constraint_values, (size_t)m->n_dofs));
  else
    ctx->cvalues.release();
  GLSNS_TRY(dev_upload(ctx, ctx->rowptr, m->row_ptr, (size_t)m->n_owned + 1));
  // (a few elements of padding: bulk copies of the last row's entries round up to 16 bytes)
  GLSNS_TRY(dev_alloc(ctx, ctx->col, (size_t)nnz + 8));
  if (nnz)
    GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->col.p, m->col_idx, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice,
                                    ctx->stream));
  GLSNS_TRY(dev_upload(ctx, ctx->color_cells, m->color_cells, nc));
  ctx->force.release();
  ctx->have_force = false;
  // halo
  ctx->n_neighbors = m->n_neighbors;
  ctx->neighbor_rank.clear(), ctx->send_ptr.clear(), ctx->recv_ptr.clear();
  if (m->n_neighbors > 0)
    {
      if (!m->neighbor_rank || !m->send_ptr || !m->recv_ptr)
        return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "halo arrays are null");
      ctx->neighbor_rank.assign(m->neighbor_rank, m->neighbor_rank + m->n_neighbors);
      ctx->send_ptr.assign(m->send_ptr, m->send_ptr + m->n_neighbors + 1);
      ctx->recv_ptr.assign(m->recv_ptr, m->recv_ptr + m->n_neighbors + 1);
      const size_t ns = (size_t)ctx->send_ptr.back();
      if (ns && !m->send_idx)
        return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "send_idx is null");
      GLSNS_TRY(dev_upload(ctx, ctx->send_idx, m->send_idx, ns));
      GLSNS_TRY(dev_alloc(ctx, ctx->send_buf, std::max<size_t>(ns, 1)));
      if (ctx->recv_ptr.back() != m->n_dofs - m->n_owned)
        return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "recv_ptr does not cover the ghost range");
    }
  else if (m->n_dofs != m->n_owned && ctx->n_ranks == 1)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "ghost dofs without neighbours");
  // matrix storage, vectors
  GLSNS_TRY(dev_alloc(ctx, ctx->val, (size_t)nnz + 8));
  ctx->lu.release();
  for (int v = 0; v < 7; ++v)
    {
      const size_t len =
        is_ghosted_input((glsns_vector)v) ? (size_t)m->n_dofs : (size_t)m->n_owned;
      GLSNS_TRY(dev_alloc(ctx, ctx->vec[v], std::max<size_t>(len, 1)));
    }
  // triangular-solve schedule (also validates the diagonal and uploads diag_pos)
  GLSNS_TRY(ilu_analyse(ctx, m->row_ptr, m->col_idx));
  GLSNS_TRY(ensure_workspace(ctx, ctx->krylov_m ? ctx->krylov_m : 30));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->epoch     = 0;
  ctx->have_mesh = true;
  return GLSNS_OK;
}

glsns_status
glsns_set_physics(glsns_context *ctx, double nu, glsns_velocity_source source,
                  const double omega[3])
{
  CHECK_CTX(ctx);
  if (!(nu > 0) || (source != GLSNS_SOURCE_NONE && source != GLSNS_SOURCE_SRF))
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "bad physics");
  ctx->viscosity = nu;
  ctx->srf       = source == GLSNS_SOURCE_SRF;
  for (int i = 0; i < 3; ++i)
    ctx->omega[i] = omega ? omega[i] : 0.0;
  return GLSNS_OK;
}

glsns_status
glsns_set_forcing(glsns_context *ctx, const double *force_at_q)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh)
    return fail(ctx, GLSNS_ERR_STATE, "no mesh");
  if (!force_at_q)
    {
      ctx->have_force = false;
      return GLSNS_OK;
    }
  GLSNS_TRY(dev_upload(ctx, ctx->force, force_at_q,
                       (size_t)ctx->n_cells * ctx->n_q * ctx->dim));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->have_force = true;
  return GLSNS_OK;
}

glsns_status
glsns_set_vector(glsns_context *ctx, glsns_vector which, const double *host, int64_t n)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh)
    return fail(ctx, GLSNS_ERR_STATE, "no mesh");
  if ((int)which < 0 || (int)which > 6 || !host)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "bad vector id or null pointer");
  const int64_t len = is_ghosted_input(which) ? ctx->n_dofs : ctx->n_owned;
  if (n != len)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "vector length mismatch");
  GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->vec[which].p, host, sizeof(double) * n,
                                  cudaMemcpyHostToDevice, ctx->stream));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->vec_set[which] = true;
  if (which == GLSNS_VEC_SYSTEM_RHS)
    ctx->have_rhs = true;
  return GLSNS_OK;
}

glsns_status
glsns_get_vector(glsns_context *ctx, glsns_vector which, double *host, int64_t n)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh)
    return fail(ctx, GLSNS_ERR_STATE, "no mesh");
  if ((int)which < 0 || (int)which > 6 || !host)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "bad vector id or null pointer");
  const int64_t len = is_ghosted_input(which) ? ctx->n_dofs : ctx->n_owned;
  if (n != len)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "vector length mismatch");
  if (!ctx->vec_set[which])
    return fail(ctx, GLSNS_ERR_STATE, "vector has not been produced yet");
  GLSNS_CUDA(ctx, cudaMemcpyAsync(host, ctx->vec[which].p, sizeof(double) * n,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GLSNS_OK;
}

glsns_status
glsns_assemble(glsns_context *ctx, int32_t assemble_matrix, glsns_scheme scheme,
               const double *time_steps)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh)
    return fail(ctx, GLSNS_ERR_STATE, "no mesh");
  if (!ctx->vec_set[GLSNS_VEC_EVALUATION_POINT])
    return fail(ctx, GLSNS_ERR_STATE, "evaluation_point has not been set");
  bool   transient;
  double sdt, c[4];
  GLSNS_TRY(scheme_coefficients(ctx, scheme, time_steps, transient, sdt, c));
  const glsns_vector hist[3] = {GLSNS_VEC_SOLUTION_M1, GLSNS_VEC_SOLUTION_M2,
                                GLSNS_VEC_SOLUTION_M3};
  for (int k = 0; k < 3; ++k)
    if (c[k + 1] != 0.0 && !ctx->vec_set[hist[k]])
      return fail(ctx, GLSNS_ERR_STATE, "a previous-solution vector the scheme needs is not set");
  const TimerId id = assemble_matrix ? T_ASSEMBLE_SYSTEM : T_ASSEMBLE_RHS;
  timer_begin(ctx, id);
  GLSNS_TRY(launch_assembly(ctx, assemble_matrix != 0, transient, sdt, c));
  timer_end(ctx, id);
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  timers_drain(ctx);
  ctx->vec_set[GLSNS_VEC_SYSTEM_RHS] = true;
  ctx->have_rhs                      = true;
  if (assemble_matrix)
    {
      ctx->have_matrix = true;
      ctx->have_ilu    = false;
    }
  return GLSNS_OK;
}

glsns_status
glsns_rhs_norm(glsns_context *ctx, double *norm)
{
  CHECK_CTX(ctx);
  if (!norm)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "null pointer");
  if (!ctx->have_mesh || !ctx->have_rhs)
    return fail(ctx, GLSNS_ERR_STATE, "no right-hand side");
  return device_norm2(ctx, ctx->vec[GLSNS_VEC_SYSTEM_RHS].p, norm);
}

glsns_status
glsns_assemble_l2_projection(glsns_context *ctx, const double *initial_at_q)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh)
    return fail(ctx, GLSNS_ERR_STATE, "no mesh");
  if (!initial_at_q && ctx->n_cells)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "null pointer");
  if (ctx->n_hanging)
    return fail(ctx, GLSNS_ERR_UNSUPPORTED, "assemble_L2_projection with hanging-node constraints");
  glsns::DevBuf<double> init;
  GLSNS_TRY(dev_upload(ctx, init, initial_at_q,
                       (size_t)ctx->n_cells * ctx->n_q * (ctx->dim + 1)));
  timer_begin(ctx, T_ASSEMBLE_SYSTEM);
  glsns_status s = launch_l2_projection(ctx, init.p);
  timer_end(ctx, T_ASSEMBLE_SYSTEM);
  cudaStreamSynchronize(ctx->stream);
  timers_drain(ctx);
  init.release();
  GLSNS_TRY(s);
  ctx->have_matrix = ctx->have_rhs = true;
  ctx->have_ilu                    = false;
  ctx->vec_set[GLSNS_VEC_SYSTEM_RHS] = true;
  return GLSNS_OK;
}

glsns_status
glsns_distribute_constraints(glsns_context *ctx, glsns_vector which)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh)
    return fail(ctx, GLSNS_ERR_STATE, "no mesh");
  if ((int)which < 0 || (int)which > 6 || !ctx->vec_set[which])
    return fail(ctx, GLSNS_ERR_STATE, "vector has not been produced yet");
  const int64_t len = is_ghosted_input(which) ? ctx->n_dofs : ctx->n_owned;
  GLSNS_TRY(launch_distribute_constraints(ctx, ctx->vec[which].p, len));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GLSNS_OK;
}

glsns_status
glsns_calculate_cfl(glsns_context *ctx, glsns_vector which, const double *shape_u_at_centre,
                    int32_t fe_degree, double time_step, double *cfl)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh)
    return fail(ctx, GLSNS_ERR_STATE, "no mesh");
  if (!shape_u_at_centre || !cfl || fe_degree < 1)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "bad calculate_CFL arguments");
  if ((int)which < 0 || (int)which > 6 || !is_ghosted_input(which) || !ctx->vec_set[which])
    return fail(ctx, GLSNS_ERR_STATE, "calculate_CFL needs a ghosted solution vector that is set");
  glsns::DevBuf<double> tab, out;
  GLSNS_TRY(dev_upload(ctx, tab, shape_u_at_centre, (size_t)ctx->n_su));
  GLSNS_TRY(dev_alloc(ctx, out, 1));
  glsns_status s = launch_cfl(ctx, tab.p, ctx->vec[which].p, time_step, (double)fe_degree, out.p);
  if (s == GLSNS_OK)
    {
      cudaMemcpyAsync(cfl, out.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        s = fail(ctx, GLSNS_ERR_CUDA, "calculate_CFL");
    }
  else
    cudaStreamSynchronize(ctx->stream);
  tab.release(), out.release();
  return s;
}

glsns_status
glsns_setup_ilu(glsns_context *ctx, int32_t fill, double atol, double rtol)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh || !ctx->have_matrix)
    return fail(ctx, GLSNS_ERR_STATE, "no matrix to factorise");
  // fill = k > 0: ILU(0) on the level-of-fill pattern (installed when the level changes)
  GLSNS_TRY(ilu_install_fill(ctx, fill));
  timer_begin(ctx, T_SETUP_ILU);
  glsns_status s = launch_ilu_factor(ctx, atol, rtol);
  timer_end(ctx, T_SETUP_ILU);
  cudaStreamSynchronize(ctx->stream);
  timers_drain(ctx);
  ctx->have_ilu = s == GLSNS_OK;
  return s;
}

glsns_status
glsns_solve_linear_system(glsns_context *ctx, const glsns_linear_solver_params *p,
                          int32_t renewed_matrix, double *newton_update_out,
                          glsns_solve_info *info)
{
  CHECK_CTX(ctx);
  glsns_solve_info local;
  if (!info)
    info = &local;
  memset(info, 0, sizeof(*info));
  if (!p || p->max_iterations < 0 || !(p->relative_residual >= 0) ||
      !(p->minimum_residual >= 0))
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "bad linear solver parameters");
  if (p->method != GLSNS_SOLVER_GMRES && p->method != GLSNS_SOLVER_BICGSTAB)
    return fail(ctx, GLSNS_ERR_UNSUPPORTED, "This solver is not allowed");
  if (!ctx->have_mesh || !ctx->have_matrix || !ctx->have_rhs)
    return fail(ctx, GLSNS_ERR_STATE, "assemble_matrix_and_rhs must precede solve_linear_system");
  // gls_navier_stokes.cc:1270-1271
  if (renewed_matrix || !ctx->have_ilu)
    GLSNS_TRY(glsns_setup_ilu(ctx, p->ilu_fill, p->ilu_atol, p->ilu_rtol));
  timer_begin(ctx, T_SOLVE);
  glsns_status s = p->method == GLSNS_SOLVER_BICGSTAB ? bicgstab_solve(ctx, p, info) :
                                                        gmres_solve(ctx, p, info);
  timer_end(ctx, T_SOLVE);
  cudaStreamSynchronize(ctx->stream);
  timers_drain(ctx);
  if (s != GLSNS_OK && s != GLSNS_ERR_NO_CONVERGENCE)
    return s;
  if (newton_update_out)
    {
      GLSNS_CUDA(ctx, cudaMemcpyAsync(newton_update_out, ctx->vec[GLSNS_VEC_NEWTON_UPDATE].p,
                                      sizeof(double) * ctx->n_owned, cudaMemcpyDeviceToHost,
                                      ctx->stream));
      GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
  return s;
}

glsns_status
glsns_line_search_point(glsns_context *ctx, double alpha)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh || !ctx->vec_set[GLSNS_VEC_PRESENT_SOLUTION] ||
      !ctx->vec_set[GLSNS_VEC_NEWTON_UPDATE])
    return fail(ctx, GLSNS_ERR_STATE, "present_solution and newton_update are needed");
  GLSNS_TRY(launch_axpy_constraints(ctx, alpha));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->vec_set[GLSNS_VEC_EVALUATION_POINT] = true;
  return GLSNS_OK;
}

glsns_status
glsns_update_ghosts(glsns_context *ctx, glsns_vector which)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh)
    return fail(ctx, GLSNS_ERR_STATE, "no mesh");
  if ((int)which < 0 || (int)which > 6 || !is_ghosted_input(which))
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "not a ghosted vector");
  if (!ctx->vec_set[which])
    return fail(ctx, GLSNS_ERR_STATE, "vector has not been set");
  GLSNS_TRY(halo_exchange(ctx, ctx->vec[which].p));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GLSNS_OK;
}

glsns_status
glsns_accept_evaluation_point(glsns_context *ctx)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh || !ctx->vec_set[GLSNS_VEC_EVALUATION_POINT])
    return fail(ctx, GLSNS_ERR_STATE, "no evaluation point");
  GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->vec[GLSNS_VEC_PRESENT_SOLUTION].p,
                                  ctx->vec[GLSNS_VEC_EVALUATION_POINT].p,
                                  sizeof(double) * ctx->n_dofs, cudaMemcpyDeviceToDevice,
                                  ctx->stream));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->vec_set[GLSNS_VEC_PRESENT_SOLUTION] = true;
  return GLSNS_OK;
}

glsns_status
glsns_get_matrix_values(glsns_context *ctx, double *values, int64_t nnz)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh || !ctx->have_matrix)
    return fail(ctx, GLSNS_ERR_STATE, "no matrix");
  if (!values || nnz != ctx->nnz_base)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "nnz mismatch");
  return matrix_values_to_host(ctx, ctx->val.p, values);
}

glsns_status
glsns_set_matrix_values(glsns_context *ctx, const double *values, int64_t nnz)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh)
    return fail(ctx, GLSNS_ERR_STATE, "no mesh");
  if (!values || nnz != ctx->nnz_base)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "nnz mismatch");
  GLSNS_TRY(matrix_values_from_host(ctx, values));
  ctx->have_matrix = true;
  ctx->have_ilu    = false;
  return GLSNS_OK;
}

glsns_status
glsns_get_ilu_values(glsns_context *ctx, double *values, int64_t nnz)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh || !ctx->have_ilu)
    return fail(ctx, GLSNS_ERR_STATE, "no ILU factors");
  if (!values || nnz != ctx->nnz)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "nnz mismatch");
  GLSNS_CUDA(ctx, cudaMemcpyAsync(values, ctx->lu.p, sizeof(double) * nnz,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GLSNS_OK;
}

glsns_status
glsns_get_ilu_pattern(glsns_context *ctx, int64_t *nnz, int64_t *row_ptr, int32_t *col_idx)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh)
    return fail(ctx, GLSNS_ERR_STATE, "no mesh");
  if (nnz)
    *nnz = ctx->nnz;
  if (row_ptr)
    GLSNS_CUDA(ctx, cudaMemcpyAsync(row_ptr, ctx->rowptr.p, sizeof(int64_t) * (ctx->n_owned + 1),
                                    cudaMemcpyDeviceToHost, ctx->stream));
  if (col_idx)
    GLSNS_CUDA(ctx, cudaMemcpyAsync(col_idx, ctx->col.p, sizeof(int32_t) * ctx->nnz,
                                    cudaMemcpyDeviceToHost, ctx->stream));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GLSNS_OK;
}

glsns_status
glsns_spmv(glsns_context *ctx, const double *x, double *y)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh || !ctx->have_matrix)
    return fail(ctx, GLSNS_ERR_STATE, "no matrix");
  if (!x || !y)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "null pointer");
  GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->zg.p, x, sizeof(double) * ctx->n_dofs,
                                  cudaMemcpyHostToDevice, ctx->stream));
  GLSNS_TRY(launch_spmv(ctx, ctx->zg.p, ctx->w.p));
  GLSNS_CUDA(ctx, cudaMemcpyAsync(y, ctx->w.p, sizeof(double) * ctx->n_owned,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GLSNS_OK;
}

glsns_status
glsns_ilu_apply(glsns_context *ctx, const double *r, double *z)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh || !ctx->have_ilu)
    return fail(ctx, GLSNS_ERR_STATE, "no ILU factors");
  if (!r || !z)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "null pointer");
  GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->tvec.p, r, sizeof(double) * ctx->n_owned,
                                  cudaMemcpyHostToDevice, ctx->stream));
  GLSNS_TRY(launch_ilu_apply(ctx, ctx->tvec.p, ctx->zg.p));
  GLSNS_CUDA(ctx, cudaMemcpyAsync(z, ctx->zg.p, sizeof(double) * ctx->n_owned,
                                  cudaMemcpyDeviceToHost, ctx->stream));
  return check_counters(ctx, "ILU apply");
}

glsns_status
glsns_ilu_apply_trace(glsns_context *ctx, const double *r, double *z, uint64_t *t_publish,
                      int32_t *row_warp)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh || !ctx->have_ilu)
    return fail(ctx, GLSNS_ERR_STATE, "no ILU factors");
  if (!r || !z || !t_publish)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "null pointer");
  const int64_t                 n = ctx->n_owned;
  glsns::DevBuf<unsigned long long> tr;
  GLSNS_TRY(dev_alloc(ctx, tr, (size_t)10 * n));
  cudaMemsetAsync(tr.p, 0, sizeof(unsigned long long) * 10 * n, ctx->stream);
  GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->tvec.p, r, sizeof(double) * n, cudaMemcpyHostToDevice,
                                  ctx->stream));
  glsns_status s = launch_ilu_apply(ctx, ctx->tvec.p, ctx->zg.p, tr.p);
  if (s == GLSNS_OK)
    {
      cudaMemcpyAsync(z, ctx->zg.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
      cudaMemcpyAsync(t_publish, tr.p, sizeof(uint64_t) * 10 * n, cudaMemcpyDeviceToHost,
                      ctx->stream);
      s = check_counters(ctx, "ILU apply");
    }
  tr.release();
  if (row_warp && s == GLSNS_OK)
    {
      memcpy(row_warp, ctx->trsv_row_warp_l.data(), sizeof(int32_t) * n);
      memcpy(row_warp + n, ctx->trsv_row_warp_u.data(), sizeof(int32_t) * n);
    }
  return s;
}

glsns_status
glsns_ilu_levels(glsns_context *ctx, int32_t *lower, int32_t *upper)
{
  CHECK_CTX(ctx);
  if (!ctx->have_mesh)
    return fail(ctx, GLSNS_ERR_STATE, "no mesh");
  if (lower)
    *lower = ctx->levels_l;
  if (upper)
    *upper = ctx->levels_u;
  return GLSNS_OK;
}

glsns_status
glsns_get_timers(glsns_context *ctx, glsns_timers *out)
{
  CHECK_CTX(ctx);
  if (!out)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "null pointer");
  cudaStreamSynchronize(ctx->stream);
  timers_drain(ctx);
  out->assemble_system_ms = ctx->timers[T_ASSEMBLE_SYSTEM].ms;
  out->assemble_rhs_ms = ctx->timers[T_ASSEMBLE_RHS].ms;
  out->setup_ilu_ms = ctx->timers[T_SETUP_ILU].ms;
  out->solve_linear_system_ms = ctx->timers[T_SOLVE].ms;
  out->spmv_ms = ctx->timers[T_SPMV].ms;
  out->trsv_ms = ctx->timers[T_TRSV].ms;
  out->orthog_ms = ctx->timers[T_ORTHOG].ms;
  out->assemble_system_calls = ctx->timers[T_ASSEMBLE_SYSTEM].calls;
  out->assemble_rhs_calls = ctx->timers[T_ASSEMBLE_RHS].calls;
  out->setup_ilu_calls = ctx->timers[T_SETUP_ILU].calls;
  out->solve_calls = ctx->timers[T_SOLVE].calls;
  out->spmv_calls = ctx->timers[T_SPMV].calls;
  out->trsv_calls = ctx->timers[T_TRSV].calls;
  out->orthog_calls = ctx->timers[T_ORTHOG].calls;
  out->kernel_launches = ctx->kernel_launches;
  return GLSNS_OK;
}

glsns_status
glsns_reset_timers(glsns_context *ctx)
{
  CHECK_CTX(ctx);
  cudaStreamSynchronize(ctx->stream);
  timers_drain(ctx);
  for (auto &t : ctx->timers)
    {
      t.ms    = 0;
      t.calls = 0;
    }
  ctx->kernel_launches = 0;
  return GLSNS_OK;
}

glsns_status
glsns_time_kernel(glsns_context *ctx, int32_t kernel, int32_t reps, int32_t nvec, double *avg_ms)
{
  CHECK_CTX(ctx);
  if (!avg_ms || reps < 1)
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "bad arguments");
  if (!ctx->have_mesh || !ctx->have_matrix)
    return fail(ctx, GLSNS_ERR_STATE, "no matrix");
  if ((kernel == 1) && !ctx->have_ilu)
    return fail(ctx, GLSNS_ERR_STATE, "no ILU factors");
  if (kernel == 2 && (nvec < 1 || nvec > ctx->krylov_m))
    return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "nvec out of range");
  cudaEvent_t a, b;
  GLSNS_CUDA(ctx, cudaEventCreate(&a));
  GLSNS_CUDA(ctx, cudaEventCreate(&b));
  double      *rhs = ctx->vec[GLSNS_VEC_SYSTEM_RHS].p;
  glsns_status s   = GLSNS_OK;
  const double c[4] = {0, 0, 0, 0};
  // inputs: the right-hand side as the vector (ghost part of zg is whatever it was)
  cudaMemcpyAsync(ctx->zg.p, rhs, sizeof(double) * ctx->n_owned, cudaMemcpyDeviceToDevice,
                  ctx->stream);
  if (kernel == 2)
    { // an orthonormal-ish basis is not needed for timing; fill V with the rhs
      for (int v = 0; v <= nvec; ++v)
        cudaMemcpyAsync(ctx->V.p + (int64_t)v * ctx->n_owned, rhs, sizeof(double) * ctx->n_owned,
                        cudaMemcpyDeviceToDevice, ctx->stream);
    }
  cudaStreamSynchronize(ctx->stream);
  cudaEventRecord(a, ctx->stream);
  for (int r = 0; r < reps && s == GLSNS_OK; ++r)
    switch (kernel)
      {
        case 0:
          s = launch_spmv(ctx, ctx->zg.p, ctx->w.p);
          break;
        case 1:
          s = launch_ilu_apply(ctx, rhs, ctx->zg.p);
          break;
        case 2:
          cudaMemcpyAsync(ctx->w.p, rhs, sizeof(double) * ctx->n_owned, cudaMemcpyDeviceToDevice,
                          ctx->stream);
          s = time_orthog(ctx, nvec);
          break;
        case 3:
        case 4:
          if (!ctx->vec_set[GLSNS_VEC_EVALUATION_POINT])
            s = fail(ctx, GLSNS_ERR_STATE, "no evaluation point");
          else
            s = launch_assembly(ctx, kernel == 3, false, 0.0, c);
          break;
        case 5:
          s = launch_ilu_factor(ctx, 1e-12, 1.0);
          if (s == GLSNS_OK)
            ctx->have_ilu = true;
          break;
        default:
          s = fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "unknown kernel id");
      }
  cudaEventRecord(b, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *avg_ms = ms / reps;
  if (kernel == 3)
    ctx->have_ilu = false;
  return s;
}

} // extern "C"
