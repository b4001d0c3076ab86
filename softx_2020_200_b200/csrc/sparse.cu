// Sparse kernels of the Krylov solve: CSR SpMV, ILU(0) factorisation and the two
// triangular solves of its application.
//
// Replace what the reference reaches through Trilinos
// (source/solvers/gls_navier_stokes.cc:1161-1176 setup_ILU -> Ifpack ILU(0),
//  :1276-1279 SolverGMRES::solve -> Epetra_CrsMatrix::Multiply and
//  Ifpack_ILU::ApplyInverse once per iteration).
//
// All three triangular kernels are level scheduled: the host sorts the rows by
// dependency level once per sparsity pattern (ilu_analyse); warps then take rows
// in that order through an atomic ticket and wait on the rows they depend on
// (point-to-point, no grid barrier between levels).  A warp's dependencies
// always hold smaller tickets, i.e. are already owned by a running warp, so the
// wait cannot deadlock whatever the residency of the grid.
//   * factorisation: a per-row "done" flag with release/acquire semantics;
//   * solves: the solution vector itself carries readiness — it is pre-filled
//     with an all-ones NaN pattern and a consumer spins until the 8-byte value
//     it needs has been overwritten (no flag traffic, no fences).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "context.h"

namespace glsns
{
  namespace
  {
    constexpr unsigned long long SENTINEL = 0xFFFFFFFFFFFFFFFFull;
    constexpr long long          SPIN_LIMIT = 1ll << 22; // ~10 s of polling: a bug guard, never reached in a correct run

    // ------------------------------------------------------------------ SpMV
    // TPR threads cooperate on one row; consecutive lanes read consecutive
    // nonzeros (coalesced 8 B + 4 B streams), x is gathered through L2/L1.
    template <int TPR>
    __global__ void __launch_bounds__(256)
    spmv_kernel(const int64_t n_rows, const int64_t *__restrict__ rowptr,
                const int32_t *__restrict__ col, const double *__restrict__ val,
                const double *__restrict__ x, double *__restrict__ y)
    {
      const int64_t gtid   = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      const int     lane   = threadIdx.x & (TPR - 1);
      const int64_t stride = ((int64_t)gridDim.x * blockDim.x) / TPR;
      for (int64_t row = gtid / TPR; row < n_rows; row += stride)
        {
          const int64_t rs = rowptr[row], re = rowptr[row + 1];
          double        s0 = 0, s1 = 0, s2 = 0, s3 = 0;
          int64_t       k = rs + lane;
          for (; k + 3 * TPR < re; k += 4 * TPR)
            {
              const int32_t c0 = __ldcs(col + k), c1 = __ldcs(col + k + TPR),
                            c2 = __ldcs(col + k + 2 * TPR), c3 = __ldcs(col + k + 3 * TPR);
              const double v0 = __ldcs(val + k), v1 = __ldcs(val + k + TPR),
                           v2 = __ldcs(val + k + 2 * TPR), v3 = __ldcs(val + k + 3 * TPR);
              s0 += v0 * __ldg(x + c0);
              s1 += v1 * __ldg(x + c1);
              s2 += v2 * __ldg(x + c2);
              s3 += v3 * __ldg(x + c3);
            }
          for (; k < re; k += TPR)
            s0 += __ldcs(val + k) * __ldg(x + __ldcs(col + k));
          double s = (s0 + s1) + (s2 + s3);
#pragma unroll
          for (int o = TPR / 2; o > 0; o >>= 1)
            s += __shfl_down_sync(0xffffffffu, s, o, TPR);
          if (lane == 0)
            y[row] = s;
        }
    }

    // ------------------------------------------------- ILU(0) factorisation
    __global__ void
    ilu_prepare_kernel(const int64_t n, const int64_t *__restrict__ diag_pos,
                       const double atol, const double rtol, double *__restrict__ lu)
    {
      // Ifpack: d <- rtol*d + sgn(d)*atol before factorising
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i < n)
        {
          const double d  = lu[diag_pos[i]];
          lu[diag_pos[i]] = rtol * d + (d >= 0 ? atol : -atol);
        }
    }

    __device__ __forceinline__ int
    ld_acquire(const int *p)
    {
      int v;
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
      return v;
    }
    __device__ __forceinline__ void
    st_release(int *p, int v)
    {
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
    __device__ __forceinline__ unsigned long long
    ld_volatile_u64(const double *p)
    {
      unsigned long long v;
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
      return v;
    }
    __device__ __forceinline__ void
    st_result(double *p, double v)
    {
      unsigned long long b = (unsigned long long)__double_as_longlong(v);
      if (b == SENTINEL) // a NaN that happens to carry the sentinel payload
        b = 0x7FF8000000000000ull;
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(b) : "memory");
    }

    constexpr int FACTOR_WARPS   = 4;   // warps per CTA
    constexpr int FACTOR_MAX_ROW = 640; // staged row length (3D Q2-Q2 vertex rows: 500)

    // IKJ ILU(0) restricted to the rank-local diagonal block (columns < n_owned).
    // One warp per row; the row is staged in shared memory, pivot rows stream
    // from L2/HBM.  Rows longer than FACTOR_MAX_ROW are updated in global memory.
    // Pivot rows are read with ld.cg: an L1 line fetched while a neighbouring row
    // was staged may hold pre-factorisation values of the pivot row's tail.
    __global__ void __launch_bounds__(FACTOR_WARPS * 32)
    ilu_factor_kernel(const int64_t n, const int32_t *__restrict__ order,
                      const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                      const int64_t *__restrict__ diag_pos, double *lu, int *row_done,
                      const int epoch, int *counters)
    {
      __shared__ double  s_val[FACTOR_WARPS][FACTOR_MAX_ROW];
      __shared__ int32_t s_col[FACTOR_WARPS][FACTOR_MAX_ROW];
      const int          warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      double            *sv = s_val[warp];
      int32_t           *sc = s_col[warp];
      for (;;)
        {
          int t = 0;
          if (lane == 0)
            t = atomicAdd(&counters[0], 1);
          t = __shfl_sync(0xffffffffu, t, 0);
          if (t >= n)
            break;
          const int32_t i  = order[t];
          const int64_t rs = rowptr[i], re = rowptr[i + 1], dp = diag_pos[i];
          const int     len = (int)(re - rs), nl = (int)(dp - rs);
          const bool    staged = len <= FACTOR_MAX_ROW;
          // columns beyond the diagonal block (ghost columns) take no part
          int nblk = len;
          if (staged)
            {
              for (int k = lane; k < len; k += 32)
                {
                  sc[k] = col[rs + k];
                  sv[k] = lu[rs + k];
                }
              __syncwarp();
            }
          for (int kk = 0; kk < nl; ++kk)
            {
              const int32_t k = staged ? sc[kk] : col[rs + kk];
              if (lane == 0)
                {
                  long long spins = 0;
                  while (ld_acquire(row_done + k) != epoch)
                    if (++spins > SPIN_LIMIT)
                      {
                        atomicExch(&counters[1], 2);
                        break;
                      }
                }
              __syncwarp();
              const int64_t kd = diag_pos[k], ke = rowptr[k + 1];
              const double  lik = (staged ? sv[kk] : __ldcg(lu + rs + kk)) / __ldcg(lu + kd);
              __syncwarp();
              if (lane == 0)
                {
                  if (staged)
                    sv[kk] = lik;
                  else
                    __stcg(lu + rs + kk, lik);
                }
              for (int64_t q = kd + 1 + lane; q < ke; q += 32)
                {
                  const int32_t j = col[q];
                  if (j >= n)
                    break; // ghost column: outside the block
                  // binary search j in the row, right of kk
                  int lo = kk + 1, hi = nblk;
                  if (staged)
                    {
                      while (lo < hi)
                        {
                          const int mid = (lo + hi) >> 1;
                          if (sc[mid] < j)
                            lo = mid + 1;
                          else
                            hi = mid;
                        }
                      if (lo < nblk && sc[lo] == j)
                        sv[lo] -= lik * __ldcg(lu + q);
                    }
                  else
                    {
                      while (lo < hi)
                        {
                          const int mid = (lo + hi) >> 1;
                          if (col[rs + mid] < j)
                            lo = mid + 1;
                          else
                            hi = mid;
                        }
                      if (lo < nblk && col[rs + lo] == j)
                        __stcg(lu + rs + lo, __ldcg(lu + rs + lo) - lik * __ldcg(lu + q));
                    }
                }
              __syncwarp();
            }
          if (staged)
            for (int k = lane; k < len; k += 32)
              lu[rs + k] = sv[k];
          __syncwarp();
          if (lane == 0)
            {
              const double piv = staged ? sv[nl] : __ldcg(lu + dp);
              if (piv == 0.0)
                atomicExch(&counters[1], 1);
              __threadfence();
              st_release(row_done + i, epoch);
            }
          __syncwarp();
        }
    }

    // ------------------------------------------------------ triangular solves
    // Rows are handled in GROUPS: up to TRSV_G consecutive rows with identical
    // column patterns (the dim+1 dofs of one mesh node under an interleaved
    // numbering).  One warp owns a group: the column indices and the awaited x
    // values of the part outside the group are loaded once for all its rows, the
    // small dense triangle inside the group is solved in registers.  This divides
    // the length of the dependency chain (and the index traffic) by the group size.
    //
    // Critical path per dependency level = one L2 store->load hop (0.36 us measured
    // on B200, tools/hop_latency.cu) + one L2 read + the warp reduction.  To stay
    // there (a) everything a group needs is described by one 32-byte record in
    // ticket order (no chain of dependent index loads), (b) the matrix entries and
    // the already-available x values are fetched as soon as the ticket is taken,
    // (c) while a group waits, ONE lane polls ONE address — the dependency the host
    // found to sit on the highest level — so waiting warps put no load on L2.
    constexpr int TRSV_G   = 4;
    constexpr int TRSV_UNR = 8;

    template <bool UPPER>
    __global__ void __launch_bounds__(128)
    trsv_kernel(const int32_t n_groups, const TrsvGroup *__restrict__ desc,
                const int32_t *__restrict__ col, const double *__restrict__ lu,
                const double *__restrict__ rhs_vec, double *x, int *counters,
                const unsigned flags)
    {
      const int lane = threadIdx.x & 31;
      for (;;)
        {
          int t = 0;
          if (lane == 0)
            t = atomicAdd(&counters[0], 1);
          t = __shfl_sync(0xffffffffu, t, 0);
          if (t >= n_groups)
            break;
          const TrsvGroup d   = desc[t];
          const int64_t   rs0 = d.rs0;
          const int       len = d.len_m & 0x0FFFFFFF, m = d.len_m >> 28, r0 = d.r0;
          const int       off = UPPER ? d.nlow + m : 0; // first entry outside the group
          const int       cnt = d.cnt;                  // entries outside the group (in block)
          // in-group triangle and right-hand side, fetched early by lane 0
          double tri[TRSV_G][TRSV_G], rhs[TRSV_G];
          if (lane == 0)
            {
#pragma unroll
              for (int a = 0; a < TRSV_G; ++a)
                if (a < m)
                  {
                    rhs[a] = rhs_vec[r0 + a];
#pragma unroll
                    for (int b = 0; b < TRSV_G; ++b)
                      if (UPPER ? (b >= a && b < m) : (b < a))
                        tri[a][b] = __ldcs(lu + rs0 + (int64_t)a * len + d.nlow + b);
                    if (UPPER) // Ifpack stores and applies the inverted diagonal as well
                      tri[a][a] = 1.0 / tri[a][a];
                  }
            }
          double s[TRSV_G];
#pragma unroll
          for (int a = 0; a < TRSV_G; ++a)
            s[a] = 0;
          bool waited = false;
          for (int k0 = 0; k0 < cnt; k0 += 32 * TRSV_UNR)
            {
              // phase A: matrix entries and the x values that are already there
              int32_t            c[TRSV_UNR];
              double             v[TRSV_UNR][TRSV_G];
              unsigned long long b[TRSV_UNR];
              unsigned           pending = 0;
#pragma unroll
              for (int u = 0; u < TRSV_UNR; ++u)
                {
                  const int k = k0 + u * 32 + lane;
                  if (k < cnt)
                    {
                      pending |= 1u << u;
                      c[u] = __ldcs(col + rs0 + off + k);
#pragma unroll
                      for (int a = 0; a < TRSV_G; ++a)
                        if (a < m)
                          v[u][a] = __ldcs(lu + rs0 + (int64_t)a * len + off + k);
                    }
                }
              long long spins = 0;
              bool      first = true;
              while (__any_sync(0xffffffffu, pending != 0))
                {
#pragma unroll
                  for (int u = 0; u < TRSV_UNR; ++u)
                    if (pending & (1u << u))
                      b[u] = ld_volatile_u64(x + c[u]);
#pragma unroll
                  for (int u = 0; u < TRSV_UNR; ++u)
                    if ((pending & (1u << u)) && (b[u] != SENTINEL || (flags & 1)))
                      {
                        pending &= ~(1u << u);
                        const double xv = __longlong_as_double((long long)b[u]);
#pragma unroll
                        for (int a = 0; a < TRSV_G; ++a)
                          if (a < m)
                            s[a] += v[u][a] * xv;
                      }
                  if (first && !waited && d.crit2 >= 0 && !(flags & 1))
                    {
                      // phase W: far from the front the whole warp parks behind one lane
                      // polling one address (the critical dependency of our critical
                      // dependency); once that is solved we are one level from the front
                      // and every lane polls its own missing entries directly.
                      if (lane == 0)
                        {
                          long long w = 0;
                          while (ld_volatile_u64(x + d.crit2) == SENTINEL)
                            if (++w > SPIN_LIMIT)
                              {
                                atomicExch(&counters[1], 2);
                                break;
                              }
                        }
                      __syncwarp();
                      waited = true;
                    }
                  first = false;
                  if (++spins > SPIN_LIMIT)
                    {
                      atomicExch(&counters[1], 2);
                      break;
                    }
                }
            }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int a = 0; a < TRSV_G; ++a)
              s[a] += __shfl_down_sync(0xffffffffu, s[a], o);
          if (lane == 0)
            {
              double out[TRSV_G];
              if (UPPER)
                {
#pragma unroll
                  for (int a = TRSV_G - 1; a >= 0; --a)
                    if (a < m)
                      {
                        double acc = rhs[a] - s[a];
#pragma unroll
                        for (int b = TRSV_G - 1; b >= 0; --b)
                          if (b > a && b < m)
                            acc -= tri[a][b] * out[b];
                        acc *= tri[a][a];
                        out[a] = acc;
                        st_result(x + r0 + a, acc);
                      }
                }
              else
                {
#pragma unroll
                  for (int a = 0; a < TRSV_G; ++a)
                    if (a < m)
                      {
                        double acc = rhs[a] - s[a];
#pragma unroll
                        for (int b = 0; b < TRSV_G; ++b)
                          if (b < a)
                            acc -= tri[a][b] * out[b];
                        out[a] = acc;
                        st_result(x + r0 + a, acc);
                      }
                }
            }
        }
    }
  } // namespace

  glsns_status
  launch_spmv(glsns_context *ctx, const double *x, double *y)
  {
    const int64_t n = ctx->n_owned;
    if (n == 0)
      return GLSNS_OK;
    const int block = 256;
    const int tpr   = ctx->avg_row_len >= 96 ? 32 : ctx->avg_row_len >= 40 ? 16 : 8;
    int64_t   want  = (n * tpr + block - 1) / block;
    const int grid  = (int)std::min<int64_t>(want, (int64_t)ctx->n_sm * 8 * 4);
    if (tpr == 32)
      spmv_kernel<32><<<grid, block, 0, ctx->stream>>>(n, ctx->rowptr.p, ctx->col.p,
                                                       ctx->val.p, x, y);
    else if (tpr == 16)
      spmv_kernel<16><<<grid, block, 0, ctx->stream>>>(n, ctx->rowptr.p, ctx->col.p,
                                                       ctx->val.p, x, y);
    else
      spmv_kernel<8><<<grid, block, 0, ctx->stream>>>(n, ctx->rowptr.p, ctx->col.p, ctx->val.p,
                                                      x, y);
    ctx->kernel_launches++;
    GLSNS_CUDA(ctx, cudaGetLastError());
    return GLSNS_OK;
  }

  // Dependency levels of the triangular solves on the diagonal block, rows
  // counting-sorted by level.  Host work, once per sparsity pattern.
  glsns_status
  ilu_analyse(glsns_context *ctx, const int64_t *rowptr, const int32_t *col)
  {
    const int64_t        n = ctx->n_owned;
    std::vector<int64_t> diag(n);
    std::vector<int32_t> lev(n), order(n);
    int32_t              max_len = 0;
    for (int64_t i = 0; i < n; ++i)
      {
        const int32_t *b = col + rowptr[i], *e = col + rowptr[i + 1];
        const int32_t *d = std::lower_bound(b, e, (int32_t)i);
        if (d == e || *d != i)
          return fail(ctx, GLSNS_ERR_BAD_ARGUMENT,
                      "sparsity pattern has no diagonal entry in row " + std::to_string(i));
        diag[i] = d - col;
        max_len = std::max<int32_t>(max_len, (int32_t)(e - b));
      }
    ctx->max_row_len = max_len;
    ctx->avg_row_len = n ? (double)rowptr[n] / (double)n : 0;
    GLSNS_TRY(dev_upload(ctx, ctx->diag_pos, diag.data(), (size_t)n));

    // ---- row-level schedule (factorisation) ----
    auto sort_by_level = [&](int64_t count, int32_t nlev, DevBuf<int32_t> &dst) -> glsns_status {
      std::vector<int64_t> start(nlev + 2, 0);
      for (int64_t i = 0; i < count; ++i)
        start[lev[i] + 1]++;
      for (int32_t l = 0; l <= nlev; ++l)
        start[l + 1] += start[l];
      for (int64_t i = 0; i < count; ++i)
        order[start[lev[i]]++] = (int32_t)i;
      return dev_upload(ctx, dst, order.data(), (size_t)count);
    };
    int32_t nl = 0;
    for (int64_t i = 0; i < n; ++i)
      {
        int32_t l = 0;
        for (int64_t k = rowptr[i]; k < diag[i]; ++k)
          l = std::max(l, lev[col[k]] + 1);
        lev[i] = l;
        nl     = std::max(nl, l);
      }
    ctx->levels_rows = n ? nl + 1 : 0;
    GLSNS_TRY(sort_by_level(n, nl, ctx->order_l));

    // ---- groups of consecutive rows with identical column patterns ----
    std::vector<int32_t> grp_ptr, grp_of(n);
    grp_ptr.reserve(n / 2 + 2);
    for (int64_t i = 0; i < n;)
      {
        grp_ptr.push_back((int32_t)i);
        const int64_t len = rowptr[i + 1] - rowptr[i];
        int64_t       j   = i + 1;
        // a diagonal-only row never groups (its pattern {i} differs from {i+1})
        while (j < n && j - i < TRSV_G && rowptr[j + 1] - rowptr[j] == len &&
               memcmp(col + rowptr[i], col + rowptr[j], sizeof(int32_t) * len) == 0)
          ++j;
        for (int64_t r = i; r < j; ++r)
          grp_of[r] = (int32_t)(grp_ptr.size() - 1);
        i = j;
      }
    const int64_t ng = (int64_t)grp_ptr.size();
    grp_ptr.push_back((int32_t)n);
    ctx->n_groups = (int32_t)ng;
    std::vector<int32_t>   glev(ng);
    std::vector<TrsvGroup> desc(ng), sorted(ng);
    auto upload_sorted = [&](int32_t nlev, DevBuf<TrsvGroup> &dst) -> glsns_status {
      std::vector<int64_t> start(nlev + 2, 0);
      for (int64_t g = 0; g < ng; ++g)
        start[glev[g] + 1]++;
      for (int32_t l = 0; l <= nlev; ++l)
        start[l + 1] += start[l];
      for (int64_t g = 0; g < ng; ++g)
        sorted[start[glev[g]]++] = desc[g];
      return dev_upload(ctx, dst, sorted.data(), (size_t)ng);
    };
    // lower: level(g) = 1 + max level(group(k)), k left of the group; the dependency on
    // the highest level (largest column on ties) is the one a waiting warp polls
    nl = 0;
    for (int64_t g = 0; g < ng; ++g)
      {
        const int64_t i = grp_ptr[g];
        int32_t       l = 0, crit = -1, crit_lev = -1;
        for (int64_t k = rowptr[i]; k < diag[i]; ++k)
          {
            const int32_t dl = glev[grp_of[col[k]]];
            l                = std::max(l, dl + 1);
            if (dl >= crit_lev)
              crit_lev = dl, crit = col[k];
          }
        glev[g] = l;
        nl      = std::max(nl, l);
        TrsvGroup &d = desc[g];
        d.rs0 = rowptr[i], d.r0 = (int32_t)i;
        d.len_m = (int32_t)(rowptr[i + 1] - rowptr[i]) | ((grp_ptr[g + 1] - grp_ptr[g]) << 28);
        d.nlow = (int32_t)(diag[i] - rowptr[i]), d.cnt = d.nlow, d.crit = crit;
        d.crit2 = crit >= 0 ? desc[grp_of[crit]].crit : -1;
      }
    ctx->levels_l = ng ? nl + 1 : 0;
    GLSNS_TRY(upload_sorted(nl, ctx->desc_l));
    // upper: level(g) = 1 + max level(group(j)), j right of the group, j < n_owned
    int32_t nu = 0;
    for (int64_t g = ng - 1; g >= 0; --g)
      {
        const int64_t i = grp_ptr[g + 1] - 1; // last row of the group
        int32_t       l = 0, crit = -1, crit_lev = -1, cnt = 0;
        for (int64_t k = diag[i] + 1; k < rowptr[i + 1]; ++k)
          {
            if (col[k] >= n)
              break;
            ++cnt;
            const int32_t dl = glev[grp_of[col[k]]];
            l                = std::max(l, dl + 1);
            if (dl > crit_lev)
              crit_lev = dl, crit = col[k];
          }
        glev[g] = l;
        nu      = std::max(nu, l);
        desc[g].cnt = cnt, desc[g].crit = crit;
        desc[g].crit2 = crit >= 0 ? desc[grp_of[crit]].crit : -1;
      }
    ctx->levels_u = ng ? nu + 1 : 0;
    GLSNS_TRY(upload_sorted(nu, ctx->desc_u));
    GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // host vectors go out of scope
    return GLSNS_OK;
  }

  static glsns_status
  check_counters(glsns_context *ctx, const char *what)
  {
    int32_t h[2];
    GLSNS_CUDA(ctx, cudaMemcpyAsync(h, ctx->counters.p, sizeof(h), cudaMemcpyDeviceToHost,
                                    ctx->stream));
    GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[1] == 1)
      return fail(ctx, GLSNS_ERR_ZERO_PIVOT, std::string(what) + ": zero pivot");
    if (h[1] == 2)
      return fail(ctx, GLSNS_ERR_CUDA, std::string(what) + ": dependency wait timed out");
    return GLSNS_OK;
  }

  glsns_status
  launch_ilu_factor(glsns_context *ctx, double atol, double rtol)
  {
    const int64_t n = ctx->n_owned;
    GLSNS_TRY(dev_alloc(ctx, ctx->lu, (size_t)ctx->nnz));
    GLSNS_TRY(dev_alloc(ctx, ctx->row_done, (size_t)std::max<int64_t>(n, 1)));
    if (ctx->epoch == 0)
      GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->row_done.p, 0, sizeof(int32_t) * n, ctx->stream));
    ctx->epoch++;
    GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->lu.p, ctx->val.p, sizeof(double) * ctx->nnz,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->counters.p, 0, 2 * sizeof(int32_t), ctx->stream));
    if (n)
      {
        ilu_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
          n, ctx->diag_pos.p, atol, rtol, ctx->lu.p);
        int per_sm = 0;
        GLSNS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                          &per_sm, ilu_factor_kernel, FACTOR_WARPS * 32, 0));
        const int grid = (int)std::min<int64_t>((n + FACTOR_WARPS - 1) / FACTOR_WARPS,
                                                (int64_t)ctx->n_sm * std::max(per_sm, 1));
        ilu_factor_kernel<<<grid, FACTOR_WARPS * 32, 0, ctx->stream>>>(
          n, ctx->order_l.p, ctx->rowptr.p, ctx->col.p, ctx->diag_pos.p, ctx->lu.p,
          ctx->row_done.p, ctx->epoch, ctx->counters.p);
        ctx->kernel_launches += 2;
      }
    GLSNS_CUDA(ctx, cudaGetLastError());
    return check_counters(ctx, "ILU(0) factorisation");
  }

  // z = (LU)^-1 r; ctx->ytmp is the intermediate.  Asynchronous on ctx->stream.
  glsns_status
  launch_ilu_apply(glsns_context *ctx, const double *r, double *z)
  {
    const int64_t n = ctx->n_owned;
    if (n == 0)
      return GLSNS_OK;
    static const int block = getenv("GLSNS_TRSV_BLOCK") ? atoi(getenv("GLSNS_TRSV_BLOCK")) : 128;
    const int32_t ng    = ctx->n_groups;
    const int64_t want  = ((int64_t)ng * 32 + block - 1) / block;
    static const int      ctas_per_sm = getenv("GLSNS_TRSV_CTAS") ? atoi(getenv("GLSNS_TRSV_CTAS")) : 2;
    static const unsigned flags = getenv("GLSNS_TRSV_FLAGS") ? atoi(getenv("GLSNS_TRSV_FLAGS")) : 0;
    static const int only = getenv("GLSNS_TRSV_ONLY") ? atoi(getenv("GLSNS_TRSV_ONLY")) : 0;
    static const int grid_override = getenv("GLSNS_TRSV_GRID") ? atoi(getenv("GLSNS_TRSV_GRID")) : 0;
    const int grid = grid_override ? grid_override : (int)std::min<int64_t>(want, (int64_t)ctx->n_sm * ctas_per_sm);
    GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->ytmp.p, 0xFF, sizeof(double) * n, ctx->stream));
    GLSNS_CUDA(ctx, cudaMemsetAsync(z, 0xFF, sizeof(double) * n, ctx->stream));
    GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->counters.p, 0, sizeof(int32_t), ctx->stream));
    if (only != 2)
      trsv_kernel<false><<<grid, block, 0, ctx->stream>>>(ng, ctx->desc_l.p, ctx->col.p, ctx->lu.p, r,
                                                          ctx->ytmp.p, ctx->counters.p, flags);
    GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->counters.p, 0, sizeof(int32_t), ctx->stream));
    if (only != 1)
      trsv_kernel<true><<<grid, block, 0, ctx->stream>>>(ng, ctx->desc_u.p, ctx->col.p, ctx->lu.p,
                                                         ctx->ytmp.p, z, ctx->counters.p, flags);
    ctx->kernel_launches += 2;
    GLSNS_CUDA(ctx, cudaGetLastError());
    return GLSNS_OK;
  }
} // namespace glsns
