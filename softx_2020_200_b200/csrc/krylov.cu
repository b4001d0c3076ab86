// Right-preconditioned restarted GMRES with classical Gram–Schmidt applied twice.
//
// Replaces solve_system_GMRES (reference: source/solvers/gls_navier_stokes.cc:1242-1289),
// i.e. TrilinosWrappers::SolverGMRES -> AztecOO AZ_gmres with deal.II's defaults
// (Krylov space 30, AZ_noscaled convergence ||r||_2 < tol, zero initial guess,
// ILU from setup_ILU as right preconditioner).
//
// Device side: all vector work is fused into three kernels per Gram–Schmidt pass
// — a batched dot of w against the whole basis (one read of w per 8 basis
// vectors, warp-shuffle + shared-memory block reduction, fixed-order second stage
// so the result is reproducible), a batched axpy that subtracts the projections
// (and accumulates ||w||^2 on the second pass), and the normalisation of the new
// basis vector.  Host side: only the (j+1)-entry Hessenberg column, the Givens
// rotations and the convergence test; one stream synchronisation per iteration.
#include <math.h>

#include <algorithm>
#include <string>
#include <vector>

#include "context.h"

namespace glsns
{
  namespace
  {
    constexpr int VB = 256; // threads per block for vector kernels
    constexpr int DOT_BATCH = 8;

    __device__ __forceinline__ double
    block_sum(double v, double *sh)
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, o);
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      __syncthreads();
      if (lane == 0)
        sh[warp] = v;
      __syncthreads();
      double r = 0;
      if (warp == 0)
        {
          r = lane < (VB / 32) ? sh[lane] : 0.0;
#pragma unroll
          for (int o = (VB / 64); o > 0; o >>= 1)
            r += __shfl_down_sync(0xffffffffu, r, o);
        }
      return r; // valid in thread 0
    }

    // partials[(v0+v)*nblocks + block] = sum over this block's elements of V[v0+v][t]*w[t];
    // SELF: also partials[self_index*nblocks + block] = sum of w[t]^2 (w is in registers anyway)
    template <bool SELF>
    __global__ void __launch_bounds__(VB)
    multi_dot_kernel(const int64_t n, const double *__restrict__ V, const int64_t ld,
                     const int v0, const int nv, const double *__restrict__ w,
                     double *__restrict__ partials, const int self_index)
    {
      __shared__ double sh[VB / 32];
      double            acc[DOT_BATCH], self = 0;
#pragma unroll
      for (int v = 0; v < DOT_BATCH; ++v)
        acc[v] = 0;
      const int64_t stride = (int64_t)gridDim.x * VB;
      for (int64_t t = (int64_t)blockIdx.x * VB + threadIdx.x; t < n; t += stride)
        {
          const double wt = w[t];
#pragma unroll
          for (int v = 0; v < DOT_BATCH; ++v)
            if (v < nv)
              acc[v] += V[(int64_t)(v0 + v) * ld + t] * wt;
          if (SELF)
            self += wt * wt;
        }
#pragma unroll
      for (int v = 0; v < DOT_BATCH; ++v)
        if (v < nv)
          {
            const double r = block_sum(acc[v], sh);
            if (threadIdx.x == 0)
              partials[(int64_t)(v0 + v) * gridDim.x + blockIdx.x] = r;
          }
      if (SELF)
        {
          const double r = block_sum(self, sh);
          if (threadIdx.x == 0)
            partials[(int64_t)self_index * gridDim.x + blockIdx.x] = r;
        }
    }

    // out[v] = sum_b partials[v*nblocks + b]   (one block per v, fixed order)
    __global__ void __launch_bounds__(VB)
    reduce_partials_kernel(const int nblocks, const double *__restrict__ partials,
                           double *__restrict__ out)
    {
      __shared__ double sh[VB / 32];
      double            s = 0;
      for (int b = threadIdx.x; b < nblocks; b += VB)
        s += partials[(int64_t)blockIdx.x * nblocks + b];
      const double r = block_sum(s, sh);
      if (threadIdx.x == 0)
        out[blockIdx.x] = r;
    }

    // w -= sum_{i<nv} h[i] V[i];  NORM: partials[block] = sum of the new w^2
    template <bool NORM>
    __global__ void __launch_bounds__(VB)
    multi_axpy_kernel(const int64_t n, const double *__restrict__ V, const int64_t ld,
                      const int nv, const double *__restrict__ h, double *__restrict__ w,
                      double *__restrict__ partials)
    {
      __shared__ double sh[VB / 32];
      __shared__ double hs[64];
      if (threadIdx.x < nv)
        hs[threadIdx.x] = h[threadIdx.x];
      __syncthreads();
      double        nrm    = 0;
      const int64_t stride = (int64_t)gridDim.x * VB;
      for (int64_t t = (int64_t)blockIdx.x * VB + threadIdx.x; t < n; t += stride)
        {
          double s0 = 0, s1 = 0;
          int    i  = 0;
          for (; i + 1 < nv; i += 2)
            {
              s0 += hs[i] * V[(int64_t)i * ld + t];
              s1 += hs[i + 1] * V[(int64_t)(i + 1) * ld + t];
            }
          if (i < nv)
            s0 += hs[i] * V[(int64_t)i * ld + t];
          const double r = w[t] - (s0 + s1);
          w[t]           = r;
          if (NORM)
            nrm += r * r;
        }
      if (NORM)
        {
          const double r = block_sum(nrm, sh);
          if (threadIdx.x == 0)
            partials[blockIdx.x] = r;
        }
    }

    // ||w - V h||^2 for w' = w before the projections h were subtracted, V orthonormal:
    // ||w'||^2 - sum h_i^2.  (Used after the SECOND Gram-Schmidt pass only, where h is the
    // rounding-level correction of a vector that is already orthogonal to V: no cancellation.)
    // The host evaluates the same expression in the same order.
    __host__ __device__ inline double
    norm2_after_projection(const double sumsq, const double *h, const int nh)
    {
      double s = sumsq;
      for (int i = 0; i < nh; ++i)
        s = fma(-h[i], h[i], s);
      return s;
    }

    // The first Gram-Schmidt update and the second pass's projections in ONE sweep over the
    // basis: w <- w - V h1 row by row, and with the row's basis entries still in registers
    // partials[v] += V[v] . w, partials[nv] += w . w.  The same operations on the same operands,
    // in the same order and with the same partial sums as multi_axpy_kernel followed by
    // multi_dot_kernel; measured: identical iteration counts, true residuals equal to 3e-13
    // (32^3 cells), every golden count reproduced.  One pass over V less: 3 nv + 5 vector passes
    // per iteration instead of 4 nv + 8 together with scaled_axpy_kernel -- 0.66 instead of
    // 0.87 ms at nv = 15, 1.36 instead of 1.63 ms at nv = 30 (64^3 cells, tools/orthog_ab.py).
    template <int NV>
    __global__ void __launch_bounds__(VB)
    axpy_dots_kernel(const int64_t n, const double *__restrict__ V, const int64_t ld, const int nv,
                     const double *__restrict__ h, double *__restrict__ w, double *__restrict__ partials)
    {
      __shared__ double sh[VB / 32];
      __shared__ double hs[64];
      if (threadIdx.x < nv)
        hs[threadIdx.x] = h[threadIdx.x];
      __syncthreads();
      double acc[NV], self = 0;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        acc[v] = 0;
      const int64_t stride = (int64_t)gridDim.x * VB;
      for (int64_t t = (int64_t)blockIdx.x * VB + threadIdx.x; t < n; t += stride)
        {
          double vv[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v)
            vv[v] = v < nv ? V[(int64_t)v * ld + t] : 0.0;
          double s0 = 0, s1 = 0;
#pragma unroll
          for (int v = 0; v < NV; ++v)
            if (v < nv)
              {
                if (v & 1)
                  s1 += hs[v] * vv[v];
                else
                  s0 += hs[v] * vv[v];
              }
          const double r = w[t] - (s0 + s1);
          w[t]           = r;
#pragma unroll
          for (int v = 0; v < NV; ++v)
            if (v < nv)
              acc[v] += vv[v] * r;
          self += r * r;
        }
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (v < nv)
          {
            const double r = block_sum(acc[v], sh);
            if (threadIdx.x == 0)
              partials[(int64_t)v * gridDim.x + blockIdx.x] = r;
          }
      const double r = block_sum(self, sh);
      if (threadIdx.x == 0)
        partials[(int64_t)nv * gridDim.x + blockIdx.x] = r;
    }

    // out = (w - sum_{i<nv} h[i] V[i]) / sqrt(*sumsq - sum h^2): the second Gram-Schmidt update and
    // the normalisation of the new basis vector in one pass (multi_axpy_kernel + scale_kernel)
    __global__ void __launch_bounds__(VB)
    scaled_axpy_kernel(const int64_t n, const double *__restrict__ V, const int64_t ld, const int nv,
                       const double *__restrict__ h, const double *__restrict__ w,
                       const double *__restrict__ sumsq, double *__restrict__ out)
    {
      __shared__ double hs[64];
      if (threadIdx.x < nv)
        hs[threadIdx.x] = h[threadIdx.x];
      __syncthreads();
      const double  f      = 1.0 / sqrt(norm2_after_projection(*sumsq, h, nv));
      const int64_t stride = (int64_t)gridDim.x * VB;
      for (int64_t t = (int64_t)blockIdx.x * VB + threadIdx.x; t < n; t += stride)
        {
          double s0 = 0, s1 = 0;
          int    i  = 0;
          for (; i + 1 < nv; i += 2)
            {
              s0 += hs[i] * V[(int64_t)i * ld + t];
              s1 += hs[i + 1] * V[(int64_t)(i + 1) * ld + t];
            }
          if (i < nv)
            s0 += hs[i] * V[(int64_t)i * ld + t];
          const double r = w[t] - (s0 + s1);
          out[t]         = r * f;
        }
    }

    // out = sum_{i<nv} y[i] V[i]
    __global__ void __launch_bounds__(VB)
    combine_kernel(const int64_t n, const double *__restrict__ V, const int64_t ld, const int nv,
                   const double *__restrict__ y, double *__restrict__ out)
    {
      __shared__ double ys[64];
      if (threadIdx.x < nv)
        ys[threadIdx.x] = y[threadIdx.x];
      __syncthreads();
      const int64_t stride = (int64_t)gridDim.x * VB;
      for (int64_t t = (int64_t)blockIdx.x * VB + threadIdx.x; t < n; t += stride)
        {
          double s = 0;
          for (int i = 0; i < nv; ++i)
            s += ys[i] * V[(int64_t)i * ld + t];
          out[t] = s;
        }
    }

    // out = w / sqrt(*sumsq - sum h^2)   (on the device) or out = w * scale when sumsq == nullptr
    __global__ void __launch_bounds__(VB)
    scale_kernel(const int64_t n, const double *__restrict__ w, const double *__restrict__ sumsq,
                 const double scale, double *__restrict__ out, const double *__restrict__ h = nullptr,
                 const int nh = 0)
    {
      const double  f      = sumsq ? 1.0 / sqrt(norm2_after_projection(*sumsq, h, nh)) : scale;
      const int64_t stride = (int64_t)gridDim.x * VB;
      for (int64_t t = (int64_t)blockIdx.x * VB + threadIdx.x; t < n; t += stride)
        out[t] = w[t] * f;
    }

    // out = a*x + b*y
    __global__ void __launch_bounds__(VB)
    lincomb_kernel(const int64_t n, const double a, const double *x, const double b,
                   const double *y, double *out) // out may alias x or y
    {
      const int64_t stride = (int64_t)gridDim.x * VB;
      for (int64_t t = (int64_t)blockIdx.x * VB + threadIdx.x; t < n; t += stride)
        out[t] = a * x[t] + b * y[t];
    }

    __global__ void __launch_bounds__(VB)
    zero_constrained_kernel(const int64_t n, const uint8_t *__restrict__ constrained,
                            double *__restrict__ x)
    {
      const int64_t stride = (int64_t)gridDim.x * VB;
      for (int64_t t = (int64_t)blockIdx.x * VB + threadIdx.x; t < n; t += stride)
        if (constrained[t])
          x[t] = 0.0;
    }

    // AffineConstraints::distribute for Dirichlet constraints: x_i = g_i on constrained dofs
    __global__ void __launch_bounds__(VB)
    distribute_kernel(const int64_t n, const uint8_t *__restrict__ constrained,
                      const double *__restrict__ cvalues, double *__restrict__ x)
    {
      const int64_t stride = (int64_t)gridDim.x * VB;
      for (int64_t t = (int64_t)blockIdx.x * VB + threadIdx.x; t < n; t += stride)
        if (constrained[t])
          x[t] = cvalues ? cvalues[t] : 0.0;
    }

    // evaluation_point = present + alpha*update, constrained dofs <- constraint values
    __global__ void __launch_bounds__(VB)
    line_search_kernel(const int64_t n, const double alpha, const double *__restrict__ present,
                       const double *__restrict__ update,
                       const uint8_t *__restrict__ constrained,
                       const double *__restrict__ cvalues, double *__restrict__ eval)
    {
      const int64_t stride = (int64_t)gridDim.x * VB;
      for (int64_t t = (int64_t)blockIdx.x * VB + threadIdx.x; t < n; t += stride)
        eval[t] = constrained[t] ? (cvalues ? cvalues[t] : 0.0) : present[t] + alpha * update[t];
    }

    // AffineConstraints::distribute for the hanging-node lines: x_i = sum_k w_k x_{m_k} (+ g_i);
    // the masters are unconstrained dofs, so one pass after the Dirichlet values is enough
    __global__ void __launch_bounds__(VB)
    distribute_hanging_kernel(const int64_t n_hanging, const int32_t *__restrict__ list,
                              const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                              const double *__restrict__ w, const double *__restrict__ inhom,
                              double *__restrict__ x)
    {
      const int64_t t = (int64_t)blockIdx.x * VB + threadIdx.x;
      if (t >= n_hanging)
        return;
      const int32_t i = list[t];
      double        s = inhom ? inhom[i] : 0.0;
      for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k)
        s += w[k] * x[idx[k]];
      x[i] = s;
    }

    int
    vec_grid(const glsns_context *ctx, int64_t n)
    {
      return (int)std::max<int64_t>(
        1, std::min<int64_t>((n + VB - 1) / VB, (int64_t)ctx->n_sm * 8));
    }

    // hbuf[out_off + v] = V[v] . w for v < nv (all-reduced over ranks); with_self: one more
    // entry, hbuf[out_off + nv] = w . w, in the same kernels and the same all-reduce
    glsns_status
    batched_dots(glsns_context *ctx, const double *V, int64_t ld, int nv, const double *w,
                 int out_off, bool with_self = false)
    {
      const int64_t n    = ctx->n_owned;
      const int     grid = vec_grid(ctx, n);
      for (int v0 = 0; v0 < nv; v0 += DOT_BATCH)
        {
          const int nb = std::min(DOT_BATCH, nv - v0);
          if (with_self && v0 + DOT_BATCH >= nv) // (the last batch also squares w)
            multi_dot_kernel<true><<<grid, VB, 0, ctx->stream>>>(n, V, ld, v0, nb, w, ctx->partials.p, nv);
          else
            multi_dot_kernel<false><<<grid, VB, 0, ctx->stream>>>(n, V, ld, v0, nb, w, ctx->partials.p, -1);
          ctx->kernel_launches++;
        }
      nv += with_self;
      reduce_partials_kernel<<<nv, VB, 0, ctx->stream>>>(grid, ctx->partials.p,
                                                         ctx->hbuf.p + out_off);
      ctx->kernel_launches++;
      GLSNS_CUDA(ctx, cudaGetLastError());
      return allreduce_sum(ctx, ctx->hbuf.p + out_off, nv);
    }
  } // namespace

  glsns_status
  ensure_workspace(glsns_context *ctx, int restart)
  {
    const int64_t n = std::max<int64_t>(ctx->n_owned, 1);
    if (restart < 1 || restart > 60)
      return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "GMRES restart must be in 1..60");
    // (dev_alloc is a no-op when the size is unchanged; the basis must follow n_owned when a
    // second glsns_set_mesh -- after refinement -- changes it, not only the restart length)
    GLSNS_TRY(dev_alloc(ctx, ctx->V, (size_t)n * (restart + 1)));
    ctx->krylov_m = restart;
    GLSNS_TRY(dev_alloc(ctx, ctx->w, (size_t)n));
    GLSNS_TRY(dev_alloc(ctx, ctx->tvec, (size_t)n));
    GLSNS_TRY(dev_alloc(ctx, ctx->ytmp, (size_t)n));
    GLSNS_TRY(dev_alloc(ctx, ctx->zg, (size_t)std::max<int64_t>(ctx->n_dofs, 1)));
    GLSNS_TRY(dev_alloc(ctx, ctx->partials, (size_t)64 * ctx->n_sm * 8));
    GLSNS_TRY(dev_alloc(ctx, ctx->hbuf, 192));
    GLSNS_TRY(dev_alloc(ctx, ctx->ycoef, 64));
    if (!ctx->h_pinned)
      GLSNS_CUDA(ctx, cudaMallocHost((void **)&ctx->h_pinned, 3 * 192 * sizeof(double)));
    for (int k = 0; k < 2; ++k)
      if (!ctx->step_event[k])
        GLSNS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->step_event[k], cudaEventDisableTiming));
    return GLSNS_OK;
  }

  glsns_status
  device_norm2(glsns_context *ctx, const double *x, double *out)
  {
    GLSNS_TRY(batched_dots(ctx, x, 0, 1, x, 0));
    GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->hbuf.p, sizeof(double),
                                    cudaMemcpyDeviceToHost, ctx->stream));
    GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = sqrt(ctx->h_pinned[0]);
    return GLSNS_OK;
  }

  // hanging dofs of x <- their lines (inhomogeneous: with the share of the Dirichlet masters)
  static glsns_status
  launch_distribute_hanging(glsns_context *ctx, double *x, bool inhomogeneous)
  {
    if (!ctx->n_hanging)
      return GLSNS_OK;
    distribute_hanging_kernel<<<(unsigned)((ctx->n_hanging + VB - 1) / VB), VB, 0, ctx->stream>>>(
      ctx->n_hanging, ctx->hang_list.p, ctx->hang_ptr.p, ctx->hang_idx.p, ctx->hang_w.p,
      inhomogeneous ? ctx->hang_inhom.p : nullptr, x);
    ctx->kernel_launches++;
    GLSNS_CUDA(ctx, cudaGetLastError());
    return GLSNS_OK;
  }

  // zero_constraints.distribute(x)
  glsns_status
  launch_zero_constrained(glsns_context *ctx, double *x)
  {
    const int64_t n = ctx->n_owned;
    zero_constrained_kernel<<<vec_grid(ctx, n), VB, 0, ctx->stream>>>(n, ctx->constrained.p, x);
    ctx->kernel_launches++;
    GLSNS_CUDA(ctx, cudaGetLastError());
    return launch_distribute_hanging(ctx, x, false);
  }

  glsns_status
  launch_distribute_constraints(glsns_context *ctx, double *x, int64_t n)
  {
    distribute_kernel<<<vec_grid(ctx, n), VB, 0, ctx->stream>>>(n, ctx->constrained.p,
                                                                ctx->cvalues.p, x);
    ctx->kernel_launches++;
    GLSNS_CUDA(ctx, cudaGetLastError());
    return launch_distribute_hanging(ctx, x, true);
  }

  glsns_status
  launch_axpy_constraints(glsns_context *ctx, double alpha)
  {
    const int64_t n = ctx->n_owned;
    line_search_kernel<<<vec_grid(ctx, n), VB, 0, ctx->stream>>>(
      n, alpha, ctx->vec[GLSNS_VEC_PRESENT_SOLUTION].p, ctx->vec[GLSNS_VEC_NEWTON_UPDATE].p,
      ctx->constrained.p, ctx->cvalues.p, ctx->vec[GLSNS_VEC_EVALUATION_POINT].p);
    ctx->kernel_launches++;
    GLSNS_CUDA(ctx, cudaGetLastError());
    GLSNS_TRY(launch_distribute_hanging(ctx, ctx->vec[GLSNS_VEC_EVALUATION_POINT].p, true));
    return halo_exchange(ctx, ctx->vec[GLSNS_VEC_EVALUATION_POINT].p);
  }

  // One CGS2 orthogonalisation of ctx->w against V[0..nv) and normalisation into
  // V[nv].  Leaves h1 in hbuf[0..nv), h2 in hbuf[64..64+nv) and, in hbuf[64+nv], ||w'||^2 of
  // the vector w' after the first pass: the norm of the final vector is
  // sqrt(||w'||^2 - sum h2^2) (norm2_after_projection), so the second pass needs no reduction
  // of its own -- two all-reduces per iteration instead of three.
  static glsns_status
  orthogonalise(glsns_context *ctx, int nv)
  {
    const int64_t n    = ctx->n_owned;
    const int     grid = vec_grid(ctx, n);
    double       *V = ctx->V.p, *w = ctx->w.p;
    GLSNS_TRY(batched_dots(ctx, V, n, nv, w, 0));
    if (ctx->gmres_fused && nv <= 32)
      {
        // w <- w - V h1 and the second pass's dots in one sweep, then w - V h2 normalised
        // straight into V[nv] (GLSNS_GMRES_FUSED=0: the four separate kernels, same bits)
        if (nv <= 8)
          axpy_dots_kernel<8><<<grid, VB, 0, ctx->stream>>>(n, V, n, nv, ctx->hbuf.p, w, ctx->partials.p);
        else if (nv <= 16)
          axpy_dots_kernel<16><<<grid, VB, 0, ctx->stream>>>(n, V, n, nv, ctx->hbuf.p, w, ctx->partials.p);
        else if (nv <= 24)
          axpy_dots_kernel<24><<<grid, VB, 0, ctx->stream>>>(n, V, n, nv, ctx->hbuf.p, w, ctx->partials.p);
        else
          axpy_dots_kernel<32><<<grid, VB, 0, ctx->stream>>>(n, V, n, nv, ctx->hbuf.p, w, ctx->partials.p);
        reduce_partials_kernel<<<nv + 1, VB, 0, ctx->stream>>>(grid, ctx->partials.p, ctx->hbuf.p + 64);
        ctx->kernel_launches += 2;
        GLSNS_CUDA(ctx, cudaGetLastError());
        GLSNS_TRY(allreduce_sum(ctx, ctx->hbuf.p + 64, nv + 1));
        scaled_axpy_kernel<<<grid, VB, 0, ctx->stream>>>(n, V, n, nv, ctx->hbuf.p + 64, w,
                                                         ctx->hbuf.p + 64 + nv, V + (int64_t)nv * n);
        ctx->kernel_launches++;
        GLSNS_CUDA(ctx, cudaGetLastError());
        return GLSNS_OK;
      }
    multi_axpy_kernel<false><<<grid, VB, 0, ctx->stream>>>(n, V, n, nv, ctx->hbuf.p, w, nullptr);
    GLSNS_TRY(batched_dots(ctx, V, n, nv, w, 64, true));
    multi_axpy_kernel<false><<<grid, VB, 0, ctx->stream>>>(n, V, n, nv, ctx->hbuf.p + 64, w, nullptr);
    scale_kernel<<<grid, VB, 0, ctx->stream>>>(n, w, ctx->hbuf.p + 64 + nv, 0.0, V + (int64_t)nv * n,
                                               ctx->hbuf.p + 64, nv);
    ctx->kernel_launches += 3;
    GLSNS_CUDA(ctx, cudaGetLastError());
    return GLSNS_OK;
  }

  glsns_status
  time_orthog(glsns_context *ctx, int nvec)
  {
    return orthogonalise(ctx, nvec);
  }

  glsns_status
  gmres_solve(glsns_context *ctx, const glsns_linear_solver_params *p, glsns_solve_info *info)
  {
    const int64_t n = ctx->n_owned;
    const int     m = p->restart;
    GLSNS_TRY(ensure_workspace(ctx, m));
    const int grid = vec_grid(ctx, n);
    double   *b = ctx->vec[GLSNS_VEC_SYSTEM_RHS].p;
    GLSNS_TRY(dev_alloc(ctx, ctx->vec[GLSNS_VEC_NEWTON_UPDATE], (size_t)std::max<int64_t>(n, 1)));
    double *x = ctx->vec[GLSNS_VEC_NEWTON_UPDATE].p;
    double *V = ctx->V.p, *w = ctx->w.p, *zg = ctx->zg.p, *tv = ctx->tvec.p;
    cudaStream_t st = ctx->stream;

    GLSNS_CUDA(ctx, cudaMemsetAsync(x, 0, sizeof(double) * n, st));
    double beta = 0;
    GLSNS_TRY(device_norm2(ctx, b, &beta));
    // gls_navier_stokes.cc:1251-1252
    const double tol = std::max(p->relative_residual * beta, p->minimum_residual);
    info->tolerance  = tol;

    std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), y(m);
    int                 it        = 0;
    bool                converged = beta < tol;
    double              est       = beta;

    while (!converged && it < p->max_iterations)
      {
        if (it == 0)
          GLSNS_CUDA(ctx, cudaMemcpyAsync(w, b, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        else
          {
            GLSNS_CUDA(ctx,
                       cudaMemcpyAsync(zg, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
            GLSNS_TRY(halo_exchange(ctx, zg));
            GLSNS_TRY(launch_spmv(ctx, zg, w));
            lincomb_kernel<<<grid, VB, 0, st>>>(n, 1.0, b, -1.0, w, w);
            ctx->kernel_launches++;
            GLSNS_TRY(device_norm2(ctx, w, &beta));
          }
        scale_kernel<<<grid, VB, 0, st>>>(n, w, nullptr, 1.0 / beta, V);
        ctx->kernel_launches++;
        std::fill(g.begin(), g.end(), 0.0);
        g[0]  = beta;
        // One Arnoldi step needs nothing from the host (V[j + 1] is normalised on the device), so
        // step j + 1 is put on the stream BEFORE the host waits for the Hessenberg column of step
        // j: the stream never runs dry while the host does its Givens rotations and launches.  If
        // step j turns out to be the last one, step j + 1 has run for nothing (it touches
        // V[j + 2], w, zg and hbuf only, none of which the update below reads); the iteration
        // count and every number the host sees are those of the unpipelined loop.
        auto enqueue_step = [&](int jj) -> glsns_status {
          timer_begin(ctx, T_TRSV);
          GLSNS_TRY(launch_ilu_apply(ctx, V + (int64_t)jj * n, zg));
          timer_end(ctx, T_TRSV);
          GLSNS_TRY(halo_exchange(ctx, zg));
          timer_begin(ctx, T_SPMV);
          GLSNS_TRY(launch_spmv(ctx, zg, w));
          timer_end(ctx, T_SPMV);
          timer_begin(ctx, T_ORTHOG);
          GLSNS_TRY(orthogonalise(ctx, jj + 1));
          timer_end(ctx, T_ORTHOG);
          GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 192 * (jj & 1), ctx->hbuf.p,
                                          (64 + jj + 2) * sizeof(double), cudaMemcpyDeviceToHost, st));
          GLSNS_CUDA(ctx, cudaEventRecord(ctx->step_event[jj & 1], st));
          return GLSNS_OK;
        };
        int j = 0, queued = -1; // (the last step that is on the stream)
        for (; j < m && it < p->max_iterations; ++j)
          {
            if (queued < j)
              GLSNS_TRY(enqueue_step(queued = j));
            if (ctx->gmres_lookahead && j + 1 < m && it + 1 < p->max_iterations)
              GLSNS_TRY(enqueue_step(queued = j + 1));
            GLSNS_CUDA(ctx, cudaEventSynchronize(ctx->step_event[j & 1]));
            const double *h1 = ctx->h_pinned + 192 * (j & 1), *h2 = h1 + 64;
            for (int i = 0; i <= j; ++i)
              H[(size_t)i * m + j] = h1[i] + h2[i];
            H[(size_t)(j + 1) * m + j] = sqrt(norm2_after_projection(h2[j + 1], h2, j + 1));
            for (int i = 0; i < j; ++i)
              {
                const double a = H[(size_t)i * m + j], c = H[(size_t)(i + 1) * m + j];
                H[(size_t)i * m + j]       = cs[i] * a + sn[i] * c;
                H[(size_t)(i + 1) * m + j] = -sn[i] * a + cs[i] * c;
              }
            const double a = H[(size_t)j * m + j], c = H[(size_t)(j + 1) * m + j];
            const double rr = hypot(a, c);
            cs[j]           = a / rr;
            sn[j]           = c / rr;
            H[(size_t)j * m + j]       = rr;
            H[(size_t)(j + 1) * m + j] = 0;
            g[j + 1]                   = -sn[j] * g[j];
            g[j]                       = cs[j] * g[j];
            ++it;
            est = fabs(g[j + 1]);
            if (est < tol)
              {
                converged = true;
                ++j;
                break;
              }
          }
        // x += M^-1 (V y),  H y = g
        for (int i = j - 1; i >= 0; --i)
          {
            double s = g[i];
            for (int k = i + 1; k < j; ++k)
              s -= H[(size_t)i * m + k] * y[k];
            y[i] = s / H[(size_t)i * m + i];
          }
        for (int i = 0; i < j; ++i) // (third slot: a step that ran ahead may still be writing one of the others)
          ctx->h_pinned[384 + i] = y[i];
        GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->ycoef.p, ctx->h_pinned + 384, sizeof(double) * j,
                                        cudaMemcpyHostToDevice, st));
        combine_kernel<<<grid, VB, 0, st>>>(n, V, n, j, ctx->ycoef.p, tv);
        ctx->kernel_launches++;
        timer_begin(ctx, T_TRSV);
        GLSNS_TRY(launch_ilu_apply(ctx, tv, zg));
        timer_end(ctx, T_TRSV);
        lincomb_kernel<<<grid, VB, 0, st>>>(n, 1.0, x, 1.0, zg, x);
        ctx->kernel_launches++;
        GLSNS_CUDA(ctx, cudaStreamSynchronize(st)); // h_pinned is reused next cycle
        timers_drain(ctx);
      }

    // explicitly recomputed ||b - A x||, what SolverControl logs
    GLSNS_CUDA(ctx, cudaMemcpyAsync(zg, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    GLSNS_TRY(halo_exchange(ctx, zg));
    GLSNS_TRY(launch_spmv(ctx, zg, w));
    lincomb_kernel<<<grid, VB, 0, st>>>(n, 1.0, b, -1.0, w, w);
    ctx->kernel_launches++;
    double tr = 0;
    GLSNS_TRY(device_norm2(ctx, w, &tr));
    // zero_constraints.distribute(x), gls_navier_stokes.cc:1287
    GLSNS_TRY(launch_zero_constrained(ctx, x));
    GLSNS_TRY(check_counters(ctx, "GMRES")); // synchronises; the solves' bug guard
    timers_drain(ctx);
    ctx->vec_set[GLSNS_VEC_NEWTON_UPDATE] = true;
    info->iterations                      = it;
    info->true_residual                   = tr;
    info->estimated_residual              = est;
    if (!converged)
      return fail(ctx, GLSNS_ERR_NO_CONVERGENCE,
                  "GMRES did not converge in " + std::to_string(it) + " iterations");
    return GLSNS_OK;
  }
  // ---------------------------------------------------------------------------------
  // Right-preconditioned BiCGStab: `method = bicgstab`, solve_system_BiCGStab
  // (reference: source/solvers/gls_navier_stokes.cc:1291-1340), i.e.
  // TrilinosWrappers::SolverBicgstab -> AztecOO AZ_bicgstab with the ILU of setup_ILU as
  // preconditioner, zero initial guess, ||r||_2 < max(rel ||rhs||, abs), one iteration =
  // two preconditioner applications and two matrix products.  No reference test or example
  // uses this method, so the iteration counts are pinned against the oracle's restatement of
  // the published algorithm only (oracle/reference_port.py: bicgstab).
  // Basis columns of ctx->V: 0 r~ (shadow residual), 1 r / s, 2 t, 3 p, 4 v.
  glsns_status
  bicgstab_solve(glsns_context *ctx, const glsns_linear_solver_params *p, glsns_solve_info *info)
  {
    const int64_t n = ctx->n_owned;
    GLSNS_TRY(ensure_workspace(ctx, std::max(p->restart, 4)));
    const int grid = vec_grid(ctx, n);
    double   *b = ctx->vec[GLSNS_VEC_SYSTEM_RHS].p;
    GLSNS_TRY(dev_alloc(ctx, ctx->vec[GLSNS_VEC_NEWTON_UPDATE], (size_t)std::max<int64_t>(n, 1)));
    double      *x = ctx->vec[GLSNS_VEC_NEWTON_UPDATE].p;
    const int64_t ld = std::max<int64_t>(n, 1);
    double      *rt = ctx->V.p, *r = rt + ld, *t = rt + 2 * ld, *pv = rt + 3 * ld, *v = rt + 4 * ld;
    double      *zg = ctx->zg.p, *w = ctx->w.p;
    cudaStream_t st = ctx->stream;
    auto         fetch = [&](int count) -> glsns_status { // hbuf[0..count) -> h_pinned
      GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->hbuf.p, sizeof(double) * count,
                                      cudaMemcpyDeviceToHost, st));
      GLSNS_CUDA(ctx, cudaStreamSynchronize(st));
      timers_drain(ctx);
      return GLSNS_OK;
    };
    auto lincomb = [&](double a, const double *xx, double bb, const double *yy, double *out) {
      lincomb_kernel<<<grid, VB, 0, st>>>(n, a, xx, bb, yy, out);
      ctx->kernel_launches++;
    };
    auto precond_matvec = [&](const double *in, double *out) -> glsns_status { // out = A M^-1 in, zg = M^-1 in
      timer_begin(ctx, T_TRSV);
      GLSNS_TRY(launch_ilu_apply(ctx, in, zg));
      timer_end(ctx, T_TRSV);
      GLSNS_TRY(halo_exchange(ctx, zg));
      timer_begin(ctx, T_SPMV);
      GLSNS_TRY(launch_spmv(ctx, zg, out));
      timer_end(ctx, T_SPMV);
      return GLSNS_OK;
    };

    GLSNS_CUDA(ctx, cudaMemsetAsync(x, 0, sizeof(double) * n, st));
    GLSNS_CUDA(ctx, cudaMemcpyAsync(r, b, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    GLSNS_CUDA(ctx, cudaMemcpyAsync(rt, b, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    GLSNS_TRY(batched_dots(ctx, rt, ld, 2, r, 0)); // <r~, r>, <r, r>
    GLSNS_TRY(fetch(2));
    double       rho = ctx->h_pinned[0], res = sqrt(ctx->h_pinned[1]);
    const double tol = std::max(p->relative_residual * res, p->minimum_residual);
    info->tolerance  = tol;
    double rho_old = 1, alpha = 1, omega = 1;
    int    it        = 0;
    bool   converged = res < tol, breakdown = false;
    while (!converged && it < p->max_iterations)
      {
        if (rho == 0.0 || omega == 0.0)
          {
            breakdown = true;
            break;
          }
        if (it == 0)
          GLSNS_CUDA(ctx, cudaMemcpyAsync(pv, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        else
          {
            const double beta = (rho / rho_old) * (alpha / omega);
            lincomb(1.0, pv, -omega, v, pv); // p = r + beta (p - omega v)
            lincomb(1.0, r, beta, pv, pv);
          }
        GLSNS_TRY(precond_matvec(pv, v)); // v = A M^-1 p
        GLSNS_TRY(batched_dots(ctx, rt, ld, 1, v, 0));
        GLSNS_TRY(fetch(1));
        if (ctx->h_pinned[0] == 0.0)
          {
            breakdown = true;
            break;
          }
        alpha = rho / ctx->h_pinned[0];
        lincomb(1.0, x, alpha, zg, x);  // x += alpha M^-1 p
        lincomb(1.0, r, -alpha, v, r);  // s = r - alpha v
        GLSNS_TRY(precond_matvec(r, t)); // t = A M^-1 s
        GLSNS_TRY(batched_dots(ctx, r, ld, 2, t, 0)); // <s, t>, <t, t>
        GLSNS_TRY(fetch(2));
        omega = ctx->h_pinned[1] != 0.0 ? ctx->h_pinned[0] / ctx->h_pinned[1] : 0.0;
        lincomb(1.0, x, omega, zg, x);  // x += omega M^-1 s
        lincomb(1.0, r, -omega, t, r);  // r = s - omega t
        rho_old = rho;
        GLSNS_TRY(batched_dots(ctx, rt, ld, 2, r, 0)); // <r~, r>, <r, r>
        GLSNS_TRY(fetch(2));
        rho = ctx->h_pinned[0];
        res = sqrt(ctx->h_pinned[1]);
        ++it;
        converged = res < tol;
      }
    // explicitly recomputed ||b - A x||, what SolverControl logs
    GLSNS_CUDA(ctx, cudaMemcpyAsync(zg, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    GLSNS_TRY(halo_exchange(ctx, zg));
    GLSNS_TRY(launch_spmv(ctx, zg, w));
    lincomb(1.0, b, -1.0, w, w);
    double tr = 0;
    GLSNS_TRY(device_norm2(ctx, w, &tr));
    GLSNS_TRY(launch_zero_constrained(ctx, x)); // zero_constraints.distribute(x), :1337
    GLSNS_TRY(check_counters(ctx, "BiCGStab"));
    timers_drain(ctx);
    ctx->vec_set[GLSNS_VEC_NEWTON_UPDATE] = true;
    info->iterations                      = it;
    info->true_residual                   = tr;
    info->estimated_residual              = res;
    if (!converged)
      return fail(ctx, GLSNS_ERR_NO_CONVERGENCE,
                  std::string("BiCGStab ") + (breakdown ? "broke down after " : "did not converge in ") +
                    std::to_string(it) + " iterations");
    return GLSNS_OK;
  }
} // namespace glsns
