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
#include <type_traits>

#include "context.h"

namespace glsns
{
  namespace
  {
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

    // Rows of one mesh node (up to 4 consecutive rows with identical column patterns,
    // found by trsv_analyse) are multiplied together: the column indices and the x
    // entries are read once per group instead of once per row, i.e. 9 instead of 12
    // bytes per nonzero for 4-row groups.  TPG threads per group, consecutive lanes on
    // consecutive nonzeros.  Measured and rejected in round 2 (64^3 cells, 4.70 ms as is):
    // requesting the next group's first 64 entries before the reduction (80 registers, three
    // CTAs per SM instead of five: 5.89 ms); the four totals reduced together in 6 shuffles
    // with the register count held at 48 by __launch_bounds__(256, 5) (7.54 ms).  What did
    // help is further down: spmv_stream_kernel, used for the large matrices.
    template <int TPG>
    __global__ void __launch_bounds__(256)
    spmv_groups_kernel(const int32_t n_groups, const int2 *__restrict__ groups,
                       const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                       const double *__restrict__ val, const double *__restrict__ x,
                       double *__restrict__ y)
    {
      const int64_t gtid   = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      const int     lane   = threadIdx.x & (TPG - 1);
      const int64_t stride = ((int64_t)gridDim.x * blockDim.x) / TPG;
      // the group descriptor and the row pointers of the NEXT group are requested before the
      // current group's entries: three dependent loads (group -> row pointer -> entries -> x)
      // per group become one
      int64_t g = gtid / TPG;
      int2    gr_n = make_int2(0, 0);
      int64_t rs_n = 0, re_n = 0;
      if (g < n_groups)
        {
          gr_n = groups[g];
          rs_n = rowptr[gr_n.x], re_n = rowptr[gr_n.x + 1];
        }
      for (; g < n_groups; g += stride)
        {
          const int2    gr = gr_n;
          const int     r0 = gr.x, m = gr.y;
          const int64_t rs = rs_n;
          const int     len = (int)(re_n - rs);
          if (g + stride < n_groups)
            {
              gr_n = groups[g + stride];
              rs_n = rowptr[gr_n.x], re_n = rowptr[gr_n.x + 1];
            }
          double        s[4][2];
#pragma unroll
          for (int a = 0; a < 4; ++a)
            s[a][0] = s[a][1] = 0;
          int k = lane;
          if (m == 4)
            { // the common case, fully unrolled
              for (; k + TPG < len; k += 2 * TPG)
                {
                  const int32_t c0 = __ldcs(col + rs + k), c1 = __ldcs(col + rs + k + TPG);
                  double        v0[4], v1[4];
#pragma unroll
                  for (int a = 0; a < 4; ++a)
                    {
                      v0[a] = __ldcs(val + rs + (int64_t)a * len + k);
                      v1[a] = __ldcs(val + rs + (int64_t)a * len + k + TPG);
                    }
                  const double x0 = __ldg(x + c0), x1 = __ldg(x + c1);
#pragma unroll
                  for (int a = 0; a < 4; ++a)
                    {
                      s[a][0] += v0[a] * x0;
                      s[a][1] += v1[a] * x1;
                    }
                }
            }
          for (; k < len; k += TPG)
            {
              const double xv = __ldg(x + __ldcs(col + rs + k));
#pragma unroll
              for (int a = 0; a < 4; ++a)
                if (a < m)
                  s[a][0] += __ldcs(val + rs + (int64_t)a * len + k) * xv;
            }
#pragma unroll
          for (int a = 0; a < 4; ++a)
            {
              double t = s[a][0] + s[a][1];
#pragma unroll
              for (int o = TPG / 2; o > 0; o >>= 1)
                t += __shfl_down_sync(0xffffffffu, t, o, TPG);
              if (lane == 0 && a < m)
                y[r0 + a] = t;
            }
        }
    }

    // ---- the same product with the matrix STREAMED through shared memory ----
    // spmv_groups_kernel keeps its loads in registers: two entries per lane and row in flight,
    // 40 warps per SM, and nothing in flight while a group's totals are reduced -- 81 % of the
    // measured HBM peak on the algorithmic bytes at 64^3 cells, 66 % on the bytes it really
    // moves.  Here one lane per warp moves whole pieces of a group (<= 256 entries per row: the
    // indices once and the entries of its <= 4 rows, five bulk copies completing on one
    // mbarrier) into a ring of two slots per warp, 12 warps per SM: up to 24 pieces of 9 KB in
    // flight per SM whatever the other lanes are doing.  The x entries of a piece are gathered as
    // soon as its indices have arrived, one piece ahead of the products; group descriptors are
    // read 32 at a time, one per lane.  The stream is marked evict-first in L2, which is left
    // to x.  Bulk copies need 16-byte alignment and rows start anywhere, so a copy starts at the
    // aligned address below its first entry and the reader skips the slack (the value and index
    // arrays are allocated with a few elements of padding for the last row).
    // Measured at 64^3 cells (2.06e9 non-zeros), ms per product: 8 warps x 3 slots 4.42,
    // 8 x 4 slots of 192 entries 5.97, 16 x 3 slots of 128 entries 4.50, **12 x 2 slots 4.17**
    // (spmv_groups_kernel 4.72): 0.91 of the HBM peak on the algorithmic bytes.  At 32^3 cells
    // it is 3 % slower than spmv_groups_kernel (0.56 vs 0.54 ms), so launch_spmv takes it for
    // matrices of 5e8 non-zeros and more (GLSNS_SPMV_STREAM=0/1 forces either).
#ifndef GLSNS_SS_WARPS
#define GLSNS_SS_WARPS 12
#endif
#ifndef GLSNS_SS_SLOTS
#define GLSNS_SS_SLOTS 2
#endif
#ifndef GLSNS_SS_CH
#define GLSNS_SS_CH 256
#endif
    constexpr int SS_WARPS = GLSNS_SS_WARPS, SS_SLOTS = GLSNS_SS_SLOTS, SS_CH = GLSNS_SS_CH;
    constexpr int SS_COLB  = 4 * SS_CH + 16;       // bytes of a slot's index part
    constexpr int SS_ROWB  = 8 * SS_CH + 16;       // bytes of one row of a slot's value part
    constexpr int SS_SLOTB = SS_COLB + 4 * SS_ROWB; // 9296

    __device__ __forceinline__ void
    ss_mbar_init(void *bar, int count)
    {
      const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    }
    __device__ __forceinline__ void
    ss_mbar_expect_tx(void *bar, unsigned bytes)
    {
      const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ bool
    ss_mbar_try_wait(void *bar, unsigned parity)
    {
      const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
      unsigned       ok;
      asm volatile("{\n\t.reg .pred p;\n\t"
                   "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(a), "r"(parity)
                   : "memory");
      return ok != 0;
    }
    __device__ __forceinline__ void
    ss_bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, void *bar, unsigned long long policy)
    {
      const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
      const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
                   "[%0], [%1], %2, [%3], %4;" ::"r"(d),
                   "l"(gsrc), "r"(bytes), "r"(b), "l"(policy)
                   : "memory");
    }

    struct SsPiece // what the consumer needs to know about the piece in a slot (32 bytes)
    {
      int64_t rs;
      int32_t r0, m, len, off, cnt, last;
    };

    __global__ void __launch_bounds__(SS_WARPS * 32, 1)
    spmv_stream_kernel(const int32_t n_groups, const int2 *__restrict__ groups,
                       const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                       const double *__restrict__ val, const double *__restrict__ x,
                       double *__restrict__ y)
    {
      extern __shared__ __align__(128) unsigned char ss_smem[];
      const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      unsigned char *ring = ss_smem + (size_t)warp * SS_SLOTS * SS_SLOTB;
      unsigned char *tail = ss_smem + (size_t)SS_WARPS * SS_SLOTS * SS_SLOTB;
      unsigned long long *bars = reinterpret_cast<unsigned long long *>(tail) + warp * SS_SLOTS;
      SsPiece *info = reinterpret_cast<SsPiece *>(tail + 8 * SS_WARPS * SS_SLOTS) + warp * SS_SLOTS;
      const int64_t W = (int64_t)gridDim.x * SS_WARPS, wid = (int64_t)blockIdx.x * SS_WARPS + warp;
      unsigned long long policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      // the warp's groups are wid, wid + W, ...; their descriptors are read 32 at a time, one
      // per lane (two dependent loads per batch instead of per group)
      int64_t batch0 = wid;       // group of lane 0 in the batch
      int     bi     = 32;        // next entry of the batch to hand out (32: batch used up)
      int     d_r0 = 0, d_m = 0, d_len = 0;
      int64_t d_rs = 0;
      int64_t n_mine = wid < n_groups ? (n_groups - 1 - wid) / W + 1 : 0; // groups of this warp
      int64_t gi     = 0;         // groups handed out so far
      auto next_group = [&](int &r0, int &m, int64_t &rs, int &len) {
        if (bi == 32)
          {
            const int64_t g = batch0 + (int64_t)lane * W;
            if (g < n_groups)
              {
                const int2 gr = groups[g];
                d_r0 = gr.x, d_m = gr.y;
                d_rs  = rowptr[gr.x];
                d_len = (int)(rowptr[gr.x + 1] - d_rs);
              }
            batch0 += 32 * W;
            bi = 0;
          }
        r0  = __shfl_sync(0xffffffffu, d_r0, bi);
        m   = __shfl_sync(0xffffffffu, d_m, bi);
        rs  = __shfl_sync(0xffffffffu, d_rs, bi);
        len = __shfl_sync(0xffffffffu, d_len, bi);
        ++bi, ++gi;
      };
      // producer position: the current group and the chunk of it to issue next
      int     p_r0 = 0, p_m = 0, p_len = 0, p_ch = 0, p_nch = 0;
      int64_t p_rs = 0;
      bool    p_valid = false;
      auto producer_step = [&](const int s) -> bool { // issue the next piece into slot s
        if (!p_valid || p_ch == p_nch)
          {
            if (gi >= n_mine)
              return false;
            next_group(p_r0, p_m, p_rs, p_len);
            p_nch = max(1, (p_len + SS_CH - 1) / SS_CH), p_ch = 0, p_valid = true;
          }
        if (lane == 0)
          {
            unsigned char *S   = ring + (size_t)s * SS_SLOTB;
            const int      off = p_ch * SS_CH, cnt = min(SS_CH, p_len - off);
            const int64_t  c0 = p_rs + off, c0a = c0 & ~(int64_t)3;
            const unsigned cb = (unsigned)(((c0 - c0a + cnt) * 4 + 15) & ~15);
            unsigned       total = cb, vb[4];
            int64_t        v0a[4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
              {
                const int64_t v0 = p_rs + (int64_t)a * p_len + off;
                v0a[a]           = v0 & ~(int64_t)1;
                vb[a]            = a < p_m ? (unsigned)(((v0 - v0a[a] + cnt) * 8 + 15) & ~15) : 0u;
                total += vb[a];
              }
            SsPiece pc;
            pc.rs = p_rs, pc.r0 = p_r0, pc.m = p_m, pc.len = p_len, pc.off = off, pc.cnt = cnt;
            pc.last = p_ch + 1 == p_nch;
            info[s] = pc;
            ss_mbar_expect_tx(bars + s, total);
            ss_bulk_load(S, col + c0a, cb, bars + s, policy);
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (a < p_m)
                ss_bulk_load(S + SS_COLB + a * SS_ROWB, val + v0a[a], vb[a], bars + s, policy);
          }
        ++p_ch;
        return true;
      };
      if (lane == 0)
        {
          for (int s = 0; s < SS_SLOTS; ++s)
            ss_mbar_init(bars + s, 1);
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
      __syncwarp();
      int64_t n_issued = 0, n_done = 0;
      for (int s = 0; s < SS_SLOTS; ++s)
        if (producer_step(s))
          ++n_issued;
      __syncwarp();
      int      slot  = 0;
      unsigned phase = 0;
      double   acc[4] = {0, 0, 0, 0};
      // the x entries of a piece are requested one piece ahead (as soon as its indices have
      // arrived), so that their trip through L2 overlaps the products of the piece before it
      double xb[2][SS_CH / 32]; // two register sets, selected at compile time by the unrolled loop
      bool   have_x = false;
      auto gather = [&](const int s, double (&out)[SS_CH / 32]) {
        const SsPiece        pc = info[s];
        const unsigned char *S  = ring + (size_t)s * SS_SLOTB;
        const int32_t *sc = reinterpret_cast<const int32_t *>(S) + (int)((pc.rs + pc.off) & 3);
#pragma unroll
        for (int u = 0; u < SS_CH / 32; ++u)
          {
            const int k = lane + 32 * u;
            out[u]      = k < pc.cnt ? __ldg(x + sc[k]) : 0.0;
          }
      };
      auto piece = [&](auto CUR) {
        constexpr int cur = decltype(CUR)::value, nxt = 1 - cur;
        while (!ss_mbar_try_wait(bars + slot, phase))
          ;
        const SsPiece        pc = info[slot];
        const unsigned char *S  = ring + (size_t)slot * SS_SLOTB;
        const double        *sv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
          sv[a] = reinterpret_cast<const double *>(S + SS_COLB + a * SS_ROWB) +
                  (int)((pc.rs + (int64_t)a * pc.len + pc.off) & 1);
        if (!have_x)
          gather(slot, xb[cur]);
        const int      ns = slot + 1 == SS_SLOTS ? 0 : slot + 1;
        const unsigned np = phase ^ (unsigned)(ns == 0);
        have_x            = n_done + 1 < n_issued && ss_mbar_try_wait(bars + ns, np);
        if (have_x)
          gather(ns, xb[nxt]);
        if (pc.m == 4)
          {
#pragma unroll
            for (int u = 0; u < SS_CH / 32; ++u)
              {
                const int k = lane + 32 * u;
                if (k < pc.cnt)
                  {
                    acc[0] += sv[0][k] * xb[cur][u];
                    acc[1] += sv[1][k] * xb[cur][u];
                    acc[2] += sv[2][k] * xb[cur][u];
                    acc[3] += sv[3][k] * xb[cur][u];
                  }
              }
          }
        else
          {
#pragma unroll
            for (int u = 0; u < SS_CH / 32; ++u)
              {
                const int k = lane + 32 * u;
                if (k < pc.cnt)
#pragma unroll
                  for (int a = 0; a < 4; ++a)
                    if (a < pc.m)
                      acc[a] += sv[a][k] * xb[cur][u];
              }
          }
        __syncwarp(); // every lane is done with the slot (and its record) before it is refilled
        if (producer_step(slot))
          ++n_issued;
        if (pc.last)
          { // the group is complete: its totals
            if (pc.m == 4)
              { // four totals over 32 lanes in 6 shuffles (lanes 8 a .. 8 a + 7 end with row a's)
                const bool h16 = lane & 16, h8 = lane & 8;
                double     k0 = h16 ? acc[2] : acc[0], k1 = h16 ? acc[3] : acc[1];
                k0 += __shfl_xor_sync(0xffffffffu, h16 ? acc[0] : acc[2], 16);
                k1 += __shfl_xor_sync(0xffffffffu, h16 ? acc[1] : acc[3], 16);
                double tot = (h8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, h8 ? k0 : k1, 8);
                tot += __shfl_xor_sync(0xffffffffu, tot, 4);
                tot += __shfl_xor_sync(0xffffffffu, tot, 2);
                tot += __shfl_xor_sync(0xffffffffu, tot, 1);
                if ((lane & 7) == 0)
                  y[pc.r0 + (lane >> 3)] = tot;
              }
            else
              {
#pragma unroll
                for (int a = 0; a < 4; ++a)
                  {
                    double t = acc[a];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
                      t += __shfl_down_sync(0xffffffffu, t, o);
                    if (lane == 0 && a < pc.m)
                      y[pc.r0 + a] = t;
                  }
              }
#pragma unroll
            for (int a = 0; a < 4; ++a)
              acc[a] = 0;
          }
        ++n_done;
        slot  = ns;
        phase = np;
        __syncwarp();
      };
      while (n_done < n_issued)
        {
          piece(std::integral_constant<int, 0>());
          if (n_done < n_issued)
            piece(std::integral_constant<int, 1>());
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


    // IKJ ILU(0) for a GROUP of up to 4 consecutive rows with identical column
    // patterns (the dim+1 dofs of a mesh node; same groups as the triangular solves,
    // trsv.cu).  The arithmetic per row is exactly the scalar IKJ elimination (pivots
    // ascending, l = a_ik / u_kk, a_ij -= l u_kj on the pattern) — the grouping only
    // shares the work around it: one warp stages the group's rows in shared memory,
    // waits once per pivot row instead of once per (row, pivot), reads each pivot row
    // once for all rows of the group and locates each of its entries in the common
    // pattern with ONE hash look-up (column -> position, built in shared memory when
    // the group is staged).  Four pivot rows are in flight: the first 128 entries of
    // row kk+3 are requested while row kk is applied.  Rows of the same group eliminate
    // each other from shared memory at the end.  Measured (B200, 3D Q2-Q2, 64^3 cells):
    // 0.9 s -> the kernel is bound by the ~410 GB of pivot-row reads it issues as
    // 1.5 KB pieces (each group re-reads the ~128 pivot rows its neighbours also
    // read); sharing pivot rows between the consecutive groups of a chain is the
    // next step.
    constexpr int FG_PRE   = 4; // 32-entry chunks of the next pivot row requested early

    __global__ void __launch_bounds__(512, 1)
    ilu_factor_groups_kernel(const int32_t n_groups, const int2 *__restrict__ groups,
                             const int64_t n, const int64_t *__restrict__ rowptr,
                             const int32_t *__restrict__ col, const int64_t *__restrict__ diag_pos,
                             double *lu, int *row_done, const int epoch, int *counters,
                             const int maxlen, const int hbits)
    {
      extern __shared__ __align__(16) unsigned char fg_smem[];
      const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const int      hsize = 1 << hbits, hmask = hsize - 1;
      const size_t   per_warp = (size_t)maxlen * 36 + 64 + (size_t)hsize * 6;
      unsigned char *W  = fg_smem + warp * per_warp;
      double        *sv = reinterpret_cast<double *>(W);                    // [4][maxlen]
      double        *l4 = reinterpret_cast<double *>(W + (size_t)maxlen * 32); // [4]
      int32_t       *sc = reinterpret_cast<int32_t *>(W + (size_t)maxlen * 32 + 64); // [maxlen]
      // column -> position in the group's rows: open addressing, linear probing
      int32_t       *hkey = sc + maxlen;                                    // [hsize], -1 = empty
      int16_t       *hpos = reinterpret_cast<int16_t *>(hkey + hsize);      // [hsize]
      auto hash = [&](int32_t j) { return (int)(((unsigned)j * 2654435761u) >> (32 - hbits)); };
      for (;;)
        {
          int t = 0;
          if (lane == 0)
            t = atomicAdd(&counters[0], 1);
          t = __shfl_sync(0xffffffffu, t, 0);
          if (t >= n_groups)
            break;
          const int2    g  = groups[t];
          const int     r0 = g.x, m = g.y;
          const int64_t rs = rowptr[r0];
          const int     len = (int)(rowptr[r0 + 1] - rs), nl = (int)(diag_pos[r0] - rs);
          for (int k = lane; k < hsize; k += 32)
            hkey[k] = -1;
          __syncwarp();
          for (int k = lane; k < len; k += 32)
            {
              const int32_t j = col[rs + k];
              sc[k]           = j;
#pragma unroll
              for (int a = 0; a < 4; ++a)
                if (a < m)
                  sv[a * maxlen + k] = lu[rs + (int64_t)a * len + k];
              int s = hash(j);
              while (atomicCAS(hkey + s, -1, j) != -1)
                s = (s + 1) & hmask;
              hpos[s] = (int16_t)k;
            }
          __syncwarp();
          // ---- pivot rows outside the group ----
          // CSR extents of the pivot rows: 32 at a time, one per lane, a batch ahead
          int64_t bkd = 0, bke = 0, nkd = 0, nke = 0; // pivots [b0, b0+32) and [b0+32, b0+64)
          int     b0  = 0;
          auto load_extents = [&](int first, int64_t &kd_, int64_t &ke_) {
            if (first + lane < nl)
              {
                const int32_t k = sc[first + lane];
                kd_             = diag_pos[k];
                ke_             = rowptr[k + 1];
              }
          };
          load_extents(0, bkd, bke);
          load_extents(32, nkd, nke);
          auto extents = [&](int kk, int64_t &kd_, int64_t &ke_) { // kk in [b0, b0+64)
            const int64_t a = __shfl_sync(0xffffffffu, bkd, kk & 31), b = __shfl_sync(0xffffffffu, bke, kk & 31);
            const int64_t c = __shfl_sync(0xffffffffu, nkd, kk & 31), d = __shfl_sync(0xffffffffu, nke, kk & 31);
            kd_             = kk < b0 + 32 ? a : c;
            ke_             = kk < b0 + 32 ? b : d;
          };
          // Pivot rows that are final already (nearly all of them: everything but the
          // last few levels) are covered by ONE acquire fence; whenever a pivot row had
          // to be waited for, the next fence covers every row found final before it.
          int  first_unready = 0;
          auto scan_ready    = [&]() { // advance first_unready over rows that are final; then fence
            while (first_unready < nl)
              {
                const int  kk   = first_unready + lane;
                const bool nope = kk >= nl ? false : *(volatile int *)(row_done + sc[kk]) != epoch;
                const unsigned bal = __ballot_sync(0xffffffffu, nope);
                if (bal)
                  {
                    first_unready += __ffs(bal) - 1;
                    break;
                  }
                first_unready = min(nl, first_unready + 32);
              }
            __threadfence();
          };
          scan_ready();
          // four pivot rows in flight: row kk+3 is requested while row kk is applied
          int64_t kdS[4], keS[4];
          double  ukkS[4];
          int32_t cjS[4][FG_PRE];
          double  uvS[4][FG_PRE];
          auto fetch = [&](auto SET, int kk) {
            constexpr int st = decltype(SET)::value;
            if (kk >= nl)
              return;
            if (kk >= first_unready)
              { // not final when last looked: wait for it
                const int32_t k = sc[kk];
                if (lane == 0)
                  {
                    long long spins = 0;
                    while (*(volatile int *)(row_done + k) != epoch)
                      if (++spins > SPIN_LIMIT)
                        {
                          atomicExch(&counters[1], 2);
                          break;
                        }
                  }
                __syncwarp();
                first_unready = kk + 1;
                scan_ready();
              }
            extents(kk, kdS[st], keS[st]);
            ukkS[st] = __ldcg(lu + kdS[st]);
#pragma unroll
            for (int u = 0; u < FG_PRE; ++u)
              {
                const int64_t q = kdS[st] + 1 + lane + 32 * u;
                cjS[st][u]      = q < keS[st] ? __ldg(col + q) : 0x7fffffff;
                uvS[st][u]      = q < keS[st] ? __ldcg(lu + q) : 0.0;
              }
          };
          auto apply_row = [&](auto SET, int kk) {
            constexpr int st = decltype(SET)::value;
            if (kk >= nl)
              return;
            // multipliers of the group's rows
            if (lane < m)
              {
                const double l         = sv[lane * maxlen + kk] / ukkS[st];
                sv[lane * maxlen + kk] = l;
                l4[lane]               = l;
              }
            __syncwarp();
            double lm[4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
              lm[a] = a < m ? l4[a] : 0.0;
            auto apply = [&](const int32_t j, const double u) {
              if (j >= n)
                return; // ghost column (or past the end): outside the diagonal block
              int s = hash(j);
              for (;;)
                {
                  const int32_t key = hkey[s];
                  if (key == j)
                    {
                      const int p = hpos[s];
#pragma unroll
                      for (int a = 0; a < 4; ++a)
                        if (a < m)
                          sv[a * maxlen + p] -= lm[a] * u;
                      return;
                    }
                  if (key == -1)
                    return; // not in the pattern: ILU(0) drops the fill
                  s = (s + 1) & hmask;
                }
            };
#pragma unroll
            for (int u = 0; u < FG_PRE; ++u)
              apply(cjS[st][u], uvS[st][u]);
            for (int64_t q = kdS[st] + 1 + lane + 32 * FG_PRE; q < keS[st]; q += 32)
              apply(__ldg(col + q), __ldcg(lu + q));
            __syncwarp();
          };
          auto shift_extents = [&](int kk_next_fetch) {
            if (kk_next_fetch >= b0 + 64 - 1)
              { // the first batch of extents is used up: shift, load the one after next
                bkd = nkd, bke = nke;
                b0 += 32;
                load_extents(b0 + 32, nkd, nke);
              }
          };
          fetch(std::integral_constant<int, 0>(), 0);
          fetch(std::integral_constant<int, 1>(), 1);
          fetch(std::integral_constant<int, 2>(), 2);
          for (int kk = 0; kk < nl; kk += 4)
            {
              shift_extents(kk + 6);
              fetch(std::integral_constant<int, 3>(), kk + 3);
              apply_row(std::integral_constant<int, 0>(), kk);
              fetch(std::integral_constant<int, 0>(), kk + 4);
              apply_row(std::integral_constant<int, 1>(), kk + 1);
              fetch(std::integral_constant<int, 1>(), kk + 5);
              apply_row(std::integral_constant<int, 2>(), kk + 2);
              fetch(std::integral_constant<int, 2>(), kk + 6);
              apply_row(std::integral_constant<int, 3>(), kk + 3);
            }
          // ---- rows of the group eliminate each other ----
          for (int b = 0; b + 1 < m; ++b)
            {
              const int pb = nl + b;
              if (lane > b && lane < m)
                {
                  const double l      = sv[lane * maxlen + pb] / sv[b * maxlen + pb];
                  sv[lane * maxlen + pb] = l;
                  l4[lane]            = l;
                }
              __syncwarp();
              for (int p = pb + 1 + lane; p < len; p += 32)
                if (sc[p] < n)
                  {
                    const double ub = sv[b * maxlen + p];
#pragma unroll
                    for (int a = 1; a < 4; ++a)
                      if (a > b && a < m)
                        sv[a * maxlen + p] -= l4[a] * ub;
                  }
              __syncwarp();
            }
          for (int k = lane; k < len; k += 32)
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (a < m)
                lu[rs + (int64_t)a * len + k] = sv[a * maxlen + k];
          if (lane < m && sv[lane * maxlen + nl + lane] == 0.0)
            atomicExch(&counters[1], 1);
          __syncwarp();
          __threadfence(); // the rows, then (one fence for all of them) their flags
          if (lane < m)
            *(volatile int *)(row_done + r0 + lane) = epoch;
          __syncwarp();
        }
    }

    // The same factorisation with the PIVOT rows taken group-wise as well.  The rows that
    // eliminate a group come in runs of up to 4 consecutive rows of one group (the dofs of a
    // neighbouring mesh node): one column pattern, so one read of the indices and ONE hash
    // look-up per column serve the whole run, and the run is one step of the warp's
    // dependent loop instead of four.  ilu_factor_groups_kernel issued ~410 GB of pivot-row
    // reads in 1.5 KB pieces, one piece per (group, pivot row) with a hash probe per entry;
    // here a piece is a run (<= 4 rows x 1 KB of values + 0.5 KB of indices) and the probes
    // drop fourfold.  The arithmetic per entry is unchanged -- the pivots of a run are
    // applied in ascending order, one fused multiply-add each, exactly the scalar IKJ
    // sequence -- so the factors are bitwise those of the row-wise kernels
    // (tests/test_gpu_parity.py compares them).
    constexpr int FR_PRE = 4; // 32-entry chunks of a run requested early

    // Per row, everything a warp needs to know about a pivot row in ONE 16-byte load: where the
    // row starts, where its diagonal is, how long it is and how many rows of its group follow
    // it (a run of pivot rows never leaves a group; diagonal-only rows are runs of one).
    // Without it a run cost up to six dependent round trips to L2 (grp_first of up to four
    // rows one after the other, then rowptr / diag_pos, then the entries); now two.
    __global__ void __launch_bounds__(256)
    ilu_rowdesc_kernel(const int64_t n, const int64_t *__restrict__ rowptr,
                       const int64_t *__restrict__ diag_pos, const int32_t *__restrict__ grp_first,
                       int4 *__restrict__ rowdesc)
    {
      const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (k >= n)
        return;
      const int64_t rs0 = rowptr[k];
      const int32_t gf  = grp_first[k];
      int           rem = 1;
      if (gf >= 0)
        while (rem < 4 && k + rem < n && grp_first[k + rem] == gf)
          ++rem;
      rowdesc[k] = make_int4((int)(unsigned)(rs0 & 0xffffffffll), (int)(rs0 >> 32), (int)(diag_pos[k] - rs0),
                             (int)(rowptr[k + 1] - rs0) | (rem << 16));
    }

    __global__ void __launch_bounds__(320, 1)
    ilu_factor_runs_kernel(const int32_t n_groups, const int2 *__restrict__ groups,
                           const int64_t n, const int64_t *__restrict__ rowptr,
                           const int32_t *__restrict__ col, const int64_t *__restrict__ diag_pos,
                           const int4 *__restrict__ rowdesc, double *lu, int *row_done,
                           const int epoch, int *counters, const int maxlen, const int hbits)
    {
      extern __shared__ __align__(16) unsigned char fg_smem[];
      const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const int      hsize = 1 << hbits, hmask = hsize - 1;
      const size_t   per_warp = (size_t)maxlen * 36 + 192 + (size_t)hsize * 6;
      unsigned char *W  = fg_smem + warp * per_warp;
      double        *sv = reinterpret_cast<double *>(W);                       // [4][maxlen]
      double        *l4 = reinterpret_cast<double *>(W + (size_t)maxlen * 32); // [4 rows][4 pivots]
      int32_t       *sc = reinterpret_cast<int32_t *>(W + (size_t)maxlen * 32 + 192); // [maxlen]
      int32_t       *hkey = sc + maxlen;                                       // [hsize], -1 = empty
      int16_t       *hpos = reinterpret_cast<int16_t *>(hkey + hsize);         // [hsize]
      auto hash = [&](int32_t j) { return (int)(((unsigned)j * 2654435761u) >> (32 - hbits)); };
      for (;;)
        {
          int t = 0;
          if (lane == 0)
            t = atomicAdd(&counters[0], 1);
          t = __shfl_sync(0xffffffffu, t, 0);
          if (t >= n_groups)
            break;
          const int2    g  = groups[t];
          const int     r0 = g.x, m = g.y;
          const int64_t rs = rowptr[r0];
          const int     len = (int)(rowptr[r0 + 1] - rs), nl = (int)(diag_pos[r0] - rs);
          for (int k = lane; k < hsize; k += 32)
            hkey[k] = -1;
          __syncwarp();
          for (int k = lane; k < len; k += 32)
            {
              const int32_t j = col[rs + k];
              sc[k]           = j;
#pragma unroll
              for (int a = 0; a < 4; ++a)
                if (a < m)
                  sv[a * maxlen + k] = lu[rs + (int64_t)a * len + k];
              int s = hash(j);
              while (atomicCAS(hkey + s, -1, j) != -1)
                s = (s + 1) & hmask;
              hpos[s] = (int16_t)k;
            }
          __syncwarp();
          // readiness of the pivot rows, as in ilu_factor_groups_kernel
          int  first_unready = 0;
          auto scan_ready    = [&]() {
            while (first_unready < nl)
              {
                const int  kk   = first_unready + lane;
                const bool nope = kk >= nl ? false : *(volatile int *)(row_done + sc[kk]) != epoch;
                const unsigned bal = __ballot_sync(0xffffffffu, nope);
                if (bal)
                  {
                    first_unready += __ffs(bal) - 1;
                    break;
                  }
                first_unready = min(nl, first_unready + 32);
              }
            __threadfence();
          };
          scan_ready();
          // two runs in flight: the next one is requested while the current one is applied
          int     rc[2] = {0, 0}, rkk[2] = {0, 0}, rbase[2] = {0, 0}, rlen[2] = {0, 0};
          int64_t rrs[2][4];
          double  rdiag[2][4], rin[2][6];
          int32_t cjS[2][FR_PRE];
          double  uvS[2][4][FR_PRE];
          auto fetch = [&](auto SET, const int kk) -> int { // returns the pivot index after the run
            constexpr int st = decltype(SET)::value;
            rc[st]           = 0;
            if (kk >= nl)
              return nl;
            const int32_t k   = sc[kk];
            const int4    dsc = __ldg(rowdesc + k);
            const int     rem = dsc.w >> 16;
            int           c   = 1;
            while (c < rem && kk + c < nl && sc[kk + c] == k + c)
              ++c;
            if (kk + c > first_unready)
              { // some row of the run was not final when last looked: wait for it
                if (lane == 0)
                  for (int b = 0; b < c; ++b)
                    {
                      long long spins = 0;
                      while (*(volatile int *)(row_done + k + b) != epoch)
                        if (++spins > SPIN_LIMIT)
                          {
                            atomicExch(&counters[1], 2);
                            break;
                          }
                    }
                __syncwarp();
                first_unready = kk + c;
                scan_ready();
              }
            rc[st] = c, rkk[st] = kk;
            const int64_t rs0 = (int64_t)(unsigned)dsc.x | ((int64_t)dsc.y << 32), d0 = dsc.z;
            rlen[st]  = dsc.w & 0xffff;
            rbase[st] = (int)d0 + c;
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (b < c)
                {
                  rrs[st][b]   = rs0 + (int64_t)b * rlen[st];
                  rdiag[st][b] = __ldcg(lu + rrs[st][b] + d0 + b);
                }
            { // couplings inside the run: u(b, b') for b < b', packed 01 02 03 12 13 23
              int q = 0;
#pragma unroll
              for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int b2 = b + 1; b2 < 4; ++b2, ++q)
                  rin[st][q] = b2 < c ? __ldcg(lu + rrs[st][b] + d0 + b2) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < FR_PRE; ++u)
              {
                const int o = rbase[st] + lane + 32 * u;
                cjS[st][u]  = o < rlen[st] ? __ldg(col + rs0 + o) : 0x7fffffff;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                  uvS[st][b][u] = (b < c && o < rlen[st]) ? __ldcg(lu + rrs[st][b] + o) : 0.0;
              }
            return kk + c;
          };
          auto apply_run = [&](auto SET) {
            constexpr int st = decltype(SET)::value;
            const int     c  = rc[st];
            if (c == 0)
              return;
            // multipliers of the group's rows, pivot by pivot (scalar IKJ order)
            if (lane < m)
              {
                double *row = sv + lane * maxlen + rkk[st];
                int     q   = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                  {
                    if (b < c)
                      {
                        const double l = row[b] / rdiag[st][b];
                        row[b]         = l;
                        l4[lane * 4 + b] = l;
#pragma unroll
                        for (int b2 = b + 1; b2 < 4; ++b2)
                          if (b2 < c)
                            row[b2] = fma(-l, rin[st][q + b2 - b - 1], row[b2]);
                      }
                    q += 3 - b;
                  }
              }
            __syncwarp();
            double lm[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b)
                lm[a][b] = (a < m && b < c) ? l4[a * 4 + b] : 0.0;
            auto apply = [&](const int32_t j, const double u0, const double u1, const double u2,
                             const double u3) {
              if (j >= n)
                return; // ghost column (or past the end): outside the diagonal block
              int s = hash(j);
              for (;;)
                {
                  const int32_t key = hkey[s];
                  if (key == j)
                    {
                      const int p = hpos[s];
#pragma unroll
                      for (int a = 0; a < 4; ++a)
                        if (a < m)
                          {
                            double v = sv[a * maxlen + p];
                            v        = fma(-lm[a][0], u0, v);
                            if (c > 1)
                              v = fma(-lm[a][1], u1, v);
                            if (c > 2)
                              v = fma(-lm[a][2], u2, v);
                            if (c > 3)
                              v = fma(-lm[a][3], u3, v);
                            sv[a * maxlen + p] = v;
                          }
                      return;
                    }
                  if (key == -1)
                    return; // not in the pattern: ILU(0) drops the fill
                  s = (s + 1) & hmask;
                }
            };
#pragma unroll
            for (int u = 0; u < FR_PRE; ++u)
              apply(cjS[st][u], uvS[st][0][u], uvS[st][1][u], uvS[st][2][u], uvS[st][3][u]);
            for (int o = rbase[st] + lane + 32 * FR_PRE; o < rlen[st]; o += 32)
              apply(__ldg(col + rrs[st][0] + o), __ldcg(lu + rrs[st][0] + o),
                    c > 1 ? __ldcg(lu + rrs[st][1] + o) : 0.0, c > 2 ? __ldcg(lu + rrs[st][2] + o) : 0.0,
                    c > 3 ? __ldcg(lu + rrs[st][3] + o) : 0.0);
            __syncwarp();
          };
          int next = fetch(std::integral_constant<int, 0>(), 0);
          while (rc[0])
            {
              next = fetch(std::integral_constant<int, 1>(), next);
              apply_run(std::integral_constant<int, 0>());
              if (!rc[1])
                break;
              next = fetch(std::integral_constant<int, 0>(), next);
              apply_run(std::integral_constant<int, 1>());
            }
          // ---- rows of the group eliminate each other ----
          for (int b = 0; b + 1 < m; ++b)
            {
              const int pb = nl + b;
              if (lane > b && lane < m)
                {
                  const double l      = sv[lane * maxlen + pb] / sv[b * maxlen + pb];
                  sv[lane * maxlen + pb] = l;
                  l4[lane]            = l;
                }
              __syncwarp();
              for (int p = pb + 1 + lane; p < len; p += 32)
                if (sc[p] < n)
                  {
                    const double ub = sv[b * maxlen + p];
#pragma unroll
                    for (int a = 1; a < 4; ++a)
                      if (a > b && a < m)
                        sv[a * maxlen + p] -= l4[a] * ub;
                  }
              __syncwarp();
            }
          for (int k = lane; k < len; k += 32)
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (a < m)
                lu[rs + (int64_t)a * len + k] = sv[a * maxlen + k];
          if (lane < m && sv[lane * maxlen + nl + lane] == 0.0)
            atomicExch(&counters[1], 1);
          __syncwarp();
          __threadfence(); // the rows, then (one fence for all of them) their flags
          if (lane < m)
            *(volatile int *)(row_done + r0 + lane) = epoch;
          __syncwarp();
        }
    }


    // The same factorisation, one CTA of FC_T threads per group.  ilu_factor_runs_kernel gives a
    // group to one warp, and a warp is a serial machine: 30-60 pivot runs one after the other,
    // each a chain of dependent loads, a hash probe and a handful of FMAs per lane -- 8.3e10 warp
    // instructions at 0.17 per scheduler and cycle with the 9 warps per SM that the staged
    // groups (24 KB of shared memory each) leave room for (ncu, round 2).  Here the columns of
    // a pivot run are dealt to FC_T threads (a run is then one or two entries per thread), the
    // multipliers are computed by the first rows' threads between two CTA barriers, and eight
    // CTAs of four warps are resident per SM.  Arithmetic per entry: unchanged (every entry of the
    // group is updated by one thread per pivot, pivots in ascending order, one fused multiply-add
    // each), so the factors are bitwise those of the other kernels.
#ifndef GLSNS_ILU_CTA_THREADS
#define GLSNS_ILU_CTA_THREADS 128
#endif
    constexpr int FC_T   = GLSNS_ILU_CTA_THREADS; // (64: 9 CTAs per SM)
    constexpr int FC_PRE = 2; // entries of a run per thread requested a run ahead

    __global__ void __launch_bounds__(FC_T, FC_T >= 128 ? 7 : 9)
    ilu_factor_cta_kernel(const int32_t n_groups, const int2 *__restrict__ groups,
                          const int64_t n, const int64_t *__restrict__ rowptr,
                          const int32_t *__restrict__ col, const int64_t *__restrict__ diag_pos,
                          const int4 *__restrict__ rowdesc, double *lu, int *row_done,
                          const int epoch, int *counters, const int maxlen, const int hbits)
    {
      extern __shared__ __align__(16) unsigned char fc_smem[];
      __shared__ int s_ticket;
      const int tid = threadIdx.x;
      const int hsize = 1 << hbits, hmask = hsize - 1;
      double   *sv = reinterpret_cast<double *>(fc_smem);                          // [4][maxlen]
      double   *l4 = reinterpret_cast<double *>(fc_smem + (size_t)maxlen * 32);    // [4 rows][4 pivots]
      int32_t  *sc = reinterpret_cast<int32_t *>(fc_smem + (size_t)maxlen * 32 + 192); // [maxlen]
      int32_t  *hkey = sc + maxlen;                                                // [hsize], -1 = empty
      int16_t  *hpos = reinterpret_cast<int16_t *>(hkey + hsize);                  // [hsize]
      auto hash = [&](int32_t j) { return (int)(((unsigned)j * 2654435761u) >> (32 - hbits)); };
      // one pivot run: what every thread knows about it, its own entries, and (threads < 4: the
      // rows of the group) the run's diagonal entries and inner couplings
      struct Run
      {
        int     c, kk, rbase, rlen;
        int64_t rs0;
        int32_t cj[FC_PRE];
        double  uv[4][FC_PRE];
      };
      for (;;)
        {
          if (tid == 0)
            s_ticket = atomicAdd(&counters[0], 1);
          __syncthreads();
          const int t = s_ticket;
          if (t >= n_groups)
            break;
          const int2    g  = groups[t];
          const int     r0 = g.x, m = g.y;
          const int64_t rs = rowptr[r0];
          const int     len = (int)(rowptr[r0 + 1] - rs), nl = (int)(diag_pos[r0] - rs);
          for (int k = tid; k < hsize; k += FC_T)
            hkey[k] = -1;
          __syncthreads();
          for (int k = tid; k < len; k += FC_T)
            {
              const int32_t j = col[rs + k];
              sc[k]           = j;
#pragma unroll
              for (int a = 0; a < 4; ++a)
                if (a < m)
                  sv[a * maxlen + k] = lu[rs + (int64_t)a * len + k];
              int s = hash(j);
              while (atomicCAS(hkey + s, -1, j) != -1)
                s = (s + 1) & hmask;
              hpos[s] = (int16_t)k;
            }
          __syncthreads();
          // the run that starts at pivot kk: wait until its rows are final (the first threads
          // look, the barrier that follows tells the others), then request it
          auto describe = [&](Run &R, const int kk) {
            R.kk = kk, R.c = 0;
            if (kk >= nl)
              return;
            const int32_t k   = sc[kk];
            const int4    dsc = __ldg(rowdesc + k); // (staging all descriptors of a group in shared memory up front: measured slower, 385 -> 462 ms)
            const int     rem = dsc.w >> 16;
            int           c   = 1;
            while (c < rem && kk + c < nl && sc[kk + c] == k + c)
              ++c;
            R.c    = c;
            R.rs0  = (int64_t)(unsigned)dsc.x | ((int64_t)dsc.y << 32);
            R.rlen = dsc.w & 0xffff;
            R.rbase = dsc.z + c;
          };
          auto wait_ready = [&](const Run &R) { // threads 32 .. 35 (a warp that computes no multipliers)
            if (tid >= 32 && tid < 32 + R.c)
              {
                const int32_t k = sc[R.kk] + (tid - 32);
                long long spins = 0;
                while (*(volatile int *)(row_done + k) != epoch)
                  if (++spins > SPIN_LIMIT)
                    {
                      atomicExch(&counters[1], 2);
                      break;
                    }
                __threadfence();
              }
          };
          // (the run's diagonal entries and inner couplings, for the threads that compute the
          // multipliers: one set -- they are used up before the next run is requested)
          double rdiag[4], rin[6];
          auto request = [&](Run &R) {
            if (R.c == 0)
              return;
            const int d0 = R.rbase - R.c;
            if (tid < 4)
              {
#pragma unroll
                for (int b = 0; b < 4; ++b)
                  rdiag[b] = b < R.c ? __ldcg(lu + R.rs0 + (int64_t)b * R.rlen + d0 + b) : 1.0;
                int q = 0;
#pragma unroll
                for (int b = 0; b < 3; ++b)
#pragma unroll
                  for (int b2 = b + 1; b2 < 4; ++b2, ++q)
                    rin[q] = b2 < R.c ? __ldcg(lu + R.rs0 + (int64_t)b * R.rlen + d0 + b2) : 0.0;
              }
#pragma unroll
            for (int u = 0; u < FC_PRE; ++u)
              {
                const int o = R.rbase + tid + FC_T * u;
                R.cj[u]     = o < R.rlen ? __ldg(col + R.rs0 + o) : 0x7fffffff;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                  R.uv[b][u] = (b < R.c && o < R.rlen) ? __ldcg(lu + R.rs0 + (int64_t)b * R.rlen + o) : 0.0;
              }
          };
          auto apply = [&](const int c, const int32_t j, const double u0, const double u1, const double u2,
                           const double u3) {
            if (j >= n)
              return; // ghost column (or past the end): outside the diagonal block
            int s = hash(j);
            for (;;)
              {
                const int32_t key = hkey[s];
                if (key == j)
                  {
                    const int p = hpos[s];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                      if (a < m)
                        {
                          double v = sv[a * maxlen + p];
                          v        = fma(-l4[a * 4 + 0], u0, v);
                          if (c > 1)
                            v = fma(-l4[a * 4 + 1], u1, v);
                          if (c > 2)
                            v = fma(-l4[a * 4 + 2], u2, v);
                          if (c > 3)
                            v = fma(-l4[a * 4 + 3], u3, v);
                          sv[a * maxlen + p] = v;
                        }
                    return;
                  }
                if (key == -1)
                  return; // not in the pattern: ILU(0) drops the fill
                s = (s + 1) & hmask;
              }
          };
          Run cur, nxt;
          describe(cur, 0);
          wait_ready(cur);
          __syncthreads();
          request(cur);
          while (cur.c)
            {
              describe(nxt, cur.kk + cur.c);
              // multipliers of the group's rows, pivot by pivot (scalar IKJ order)
              if (tid < m)
                {
                  double *row = sv + tid * maxlen + cur.kk;
                  int     q   = 0;
#pragma unroll
                  for (int b = 0; b < 4; ++b)
                    {
                      if (b < cur.c)
                        {
                          const double l = row[b] / rdiag[b];
                          row[b]         = l;
                          l4[tid * 4 + b] = l;
#pragma unroll
                          for (int b2 = b + 1; b2 < 4; ++b2)
                            if (b2 < cur.c)
                              row[b2] = fma(-l, rin[q + b2 - b - 1], row[b2]);
                        }
                      q += 3 - b;
                    }
                }
              wait_ready(nxt);
              __syncthreads();
              request(nxt);
              {
                const int c = cur.c;
#pragma unroll
                for (int u = 0; u < FC_PRE; ++u)
                  apply(c, cur.cj[u], cur.uv[0][u], cur.uv[1][u], cur.uv[2][u], cur.uv[3][u]);
                for (int o = cur.rbase + tid + FC_T * FC_PRE; o < cur.rlen; o += FC_T)
                  apply(c, __ldg(col + cur.rs0 + o), __ldcg(lu + cur.rs0 + o),
                        c > 1 ? __ldcg(lu + cur.rs0 + (int64_t)cur.rlen + o) : 0.0,
                        c > 2 ? __ldcg(lu + cur.rs0 + 2 * (int64_t)cur.rlen + o) : 0.0,
                        c > 3 ? __ldcg(lu + cur.rs0 + 3 * (int64_t)cur.rlen + o) : 0.0);
              }
              __syncthreads();
              cur = nxt;
            }
          // ---- rows of the group eliminate each other ----
          for (int b = 0; b + 1 < m; ++b)
            {
              const int pb = nl + b;
              if (tid > b && tid < m)
                {
                  const double l      = sv[tid * maxlen + pb] / sv[b * maxlen + pb];
                  sv[tid * maxlen + pb] = l;
                  l4[tid]             = l;
                }
              __syncthreads();
              for (int p = pb + 1 + tid; p < len; p += FC_T)
                if (sc[p] < n)
                  {
                    const double ub = sv[b * maxlen + p];
#pragma unroll
                    for (int a = 1; a < 4; ++a)
                      if (a > b && a < m)
                        sv[a * maxlen + p] -= l4[a] * ub;
                  }
              __syncthreads();
            }
          for (int k = tid; k < len; k += FC_T)
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (a < m)
                lu[rs + (int64_t)a * len + k] = sv[a * maxlen + k];
          if (tid < m && sv[tid * maxlen + nl + tid] == 0.0)
            atomicExch(&counters[1], 1);
          __threadfence(); // the rows, then (one barrier for all of them) their flags
          __syncthreads();
          if (tid < m)
            *(volatile int *)(row_done + r0 + tid) = epoch;
        }
    }

    // flags of the rows no group covers (diagonal-only rows): final from the start
    __global__ void __launch_bounds__(256)
    ilu_mark_rows_kernel(const int32_t n_rows, const int32_t *__restrict__ rows, int *row_done,
                         const int epoch)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i < n_rows)
        row_done[rows[i]] = epoch;
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
    static const bool by_rows = getenv("GLSNS_SPMV_BY_ROWS") != nullptr;
    static const int stream_env = getenv("GLSNS_SPMV_STREAM") ? atoi(getenv("GLSNS_SPMV_STREAM")) : -1;
    const bool       streamed   = stream_env >= 0 ? stream_env != 0 : ctx->nnz >= 500000000ll;
    if (ctx->n_sgroups > 0 && !by_rows && streamed)
      {
        const size_t smem = (size_t)SS_WARPS * SS_SLOTS * SS_SLOTB + (size_t)SS_WARPS * SS_SLOTS * (8 + sizeof(SsPiece));
        GLSNS_CUDA(ctx, cudaFuncSetAttribute(spmv_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
        const int grid = (int)std::min<int64_t>((ctx->n_sgroups + SS_WARPS - 1) / SS_WARPS, ctx->n_sm);
        spmv_stream_kernel<<<grid, SS_WARPS * 32, smem, ctx->stream>>>(
          (int32_t)ctx->n_sgroups, ctx->sgroups.p, ctx->rowptr.p, ctx->col.p, ctx->val.p, x, y);
      }
    else if (ctx->n_sgroups > 0 && !by_rows)
      {
        const int64_t ng   = ctx->n_sgroups;
        const int64_t want = (ng * tpr + block - 1) / block;
        const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->n_sm * 8 * 4);
        if (tpr == 32)
          spmv_groups_kernel<32><<<grid, block, 0, ctx->stream>>>(
            (int32_t)ng, ctx->sgroups.p, ctx->rowptr.p, ctx->col.p, ctx->val.p, x, y);
        else if (tpr == 16)
          spmv_groups_kernel<16><<<grid, block, 0, ctx->stream>>>(
            (int32_t)ng, ctx->sgroups.p, ctx->rowptr.p, ctx->col.p, ctx->val.p, x, y);
        else
          spmv_groups_kernel<8><<<grid, block, 0, ctx->stream>>>(
            (int32_t)ng, ctx->sgroups.p, ctx->rowptr.p, ctx->col.p, ctx->val.p, x, y);
      }
    else
      {
        int64_t   want = (n * tpr + block - 1) / block;
        const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->n_sm * 8 * 4);
        if (tpr == 32)
          spmv_kernel<32><<<grid, block, 0, ctx->stream>>>(n, ctx->rowptr.p, ctx->col.p,
                                                           ctx->val.p, x, y);
        else if (tpr == 16)
          spmv_kernel<16><<<grid, block, 0, ctx->stream>>>(n, ctx->rowptr.p, ctx->col.p,
                                                           ctx->val.p, x, y);
        else
          spmv_kernel<8><<<grid, block, 0, ctx->stream>>>(n, ctx->rowptr.p, ctx->col.p,
                                                          ctx->val.p, x, y);
      }
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

    GLSNS_TRY(trsv_analyse(ctx, rowptr, col, diag.data()));
    GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // host vectors go out of scope
    return GLSNS_OK;
  }

  // ------------------------------------------------- ILU(k), k > 0
  namespace
  {
    // dst[map[k]] = src[k] / dst[k] = src[map[k]]
    __global__ void __launch_bounds__(256)
    scatter_values_kernel(const int64_t n, const int64_t *__restrict__ map,
                          const double *__restrict__ src, double *__restrict__ dst)
    {
      const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (k < n)
        dst[map[k]] = src[k];
    }
    __global__ void __launch_bounds__(256)
    gather_values_kernel(const int64_t n, const int64_t *__restrict__ map,
                         const double *__restrict__ src, double *__restrict__ dst)
    {
      const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (k < n)
        dst[k] = src[map[k]];
    }

    // Level-of-fill pattern of the rank-local diagonal block: what Ifpack_IlukGraph builds
    // for TrilinosWrappers::PreconditionILU::AdditionalData(ilu_fill = k, ..., overlap = 0)
    // (reference call site source/solvers/gls_navier_stokes.cc:1166-1175).  Entries of A
    // have level 0; eliminating row i with row k < i creates (i, j) for every (k, j), j > k,
    // with level lev(i,k) + lev(k,j) + 1, kept if <= k (the minimum over all k).  Columns
    // >= n (ghosts, outside the block) stay in their rows and take no part.
    void
    iluk_symbolic(const int64_t n, const int64_t *rowptr, const int32_t *col, const int fill,
                  std::vector<int64_t> &prow, std::vector<int32_t> &pcol)
    {
      constexpr int32_t    END = INT32_MAX;
      std::vector<uint8_t> plev;
      std::vector<int64_t> pdiag((size_t)n, -1), pghost((size_t)n, 0); // diagonal / first ghost entry
      std::vector<int32_t> next((size_t)n + 1), lev((size_t)n, 0), ghosts;
      std::vector<int64_t> mark((size_t)n, -1);
      prow.assign((size_t)n + 1, 0);
      pcol.clear();
      pcol.reserve((size_t)rowptr[n] * (fill + 1));
      plev.reserve((size_t)rowptr[n] * (fill + 1));
      for (int64_t i = 0; i < n; ++i)
        {
          ghosts.clear();
          int32_t prev = (int32_t)n; // head of the sorted list
          for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
            {
              const int32_t c = col[k];
              if (c >= n)
                {
                  ghosts.push_back(c);
                  continue;
                }
              next[prev] = c, prev = c;
              lev[c] = 0, mark[c] = i;
            }
          next[prev] = END;
          for (int32_t k = next[n]; k < i; k = next[k])
            {
              const int lk = lev[k];
              int32_t   p  = k;
              for (int64_t t = pdiag[k] + 1; t < pghost[k]; ++t)
                {
                  const int nl = lk + plev[(size_t)t] + 1;
                  if (nl > fill)
                    continue;
                  const int32_t j = pcol[(size_t)t];
                  if (mark[j] == i)
                    lev[j] = std::min(lev[j], nl);
                  else
                    {
                      while (next[p] < j)
                        p = next[p];
                      next[j] = next[p], next[p] = j;
                      mark[j] = i, lev[j] = nl;
                    }
                  p = j;
                }
            }
          for (int32_t c = next[n]; c != END; c = next[c])
            {
              if (c == i)
                pdiag[i] = (int64_t)pcol.size();
              pcol.push_back(c), plev.push_back((uint8_t)lev[c]);
            }
          pghost[i] = (int64_t)pcol.size();
          for (int32_t c : ghosts)
            pcol.push_back(c), plev.push_back(0);
          prow[i + 1] = (int64_t)pcol.size();
        }
    }
  } // namespace

  // setup_ILU with `ilu preconditioner fill` = fill != the installed level: switch the device
  // pattern (and with it the SpMV groups, the factorisation order and the triangular-solve
  // schedule); the matrix values move along.
  glsns_status
  ilu_install_fill(glsns_context *ctx, const int32_t fill)
  {
    const int64_t n = ctx->n_owned;
    if (fill < 0 || fill > 200)
      return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "ilu preconditioner fill must be in 0..200");
    if (fill == ctx->ilu_fill)
      return GLSNS_OK;
    if (ctx->base_rowptr.empty())
      { // the installed pattern is the host's: fetch it (host arrays were only borrowed)
        ctx->base_rowptr.resize((size_t)n + 1);
        ctx->base_col.resize((size_t)ctx->nnz);
        ctx->nnz_base = ctx->nnz;
        GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->base_rowptr.data(), ctx->rowptr.p,
                                        sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost,
                                        ctx->stream));
        GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->base_col.data(), ctx->col.p,
                                        sizeof(int32_t) * ctx->nnz, cudaMemcpyDeviceToHost,
                                        ctx->stream));
        GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      }
    const int64_t        nnz_b = ctx->nnz_base;
    std::vector<int64_t> prow;
    std::vector<int32_t> pcol;
    if (fill > 0)
      {
        for (int64_t i = 0; i < n; ++i)
          if (!std::binary_search(ctx->base_col.begin() + ctx->base_rowptr[i],
                                  ctx->base_col.begin() + ctx->base_rowptr[i + 1], (int32_t)i))
            return fail(ctx, GLSNS_ERR_BAD_ARGUMENT, "sparsity pattern lacks a diagonal entry");
        iluk_symbolic(n, ctx->base_rowptr.data(), ctx->base_col.data(), fill, prow, pcol);
      }
    const int64_t *nrow = fill > 0 ? prow.data() : ctx->base_rowptr.data();
    const int32_t *ncol = fill > 0 ? pcol.data() : ctx->base_col.data();
    const int64_t  nnz_p = nrow[n];
    // values: installed pattern -> host pattern -> new pattern
    DevBuf<double> base_val, new_val;
    GLSNS_TRY(dev_alloc(ctx, base_val, (size_t)std::max<int64_t>(nnz_b, 1)));
    GLSNS_TRY(dev_alloc(ctx, new_val, (size_t)nnz_p + 8));
    const unsigned gb = (unsigned)((nnz_b + 255) / 256);
    if (nnz_b)
      {
        if (ctx->a2p.p)
          gather_values_kernel<<<gb, 256, 0, ctx->stream>>>(nnz_b, ctx->a2p.p, ctx->val.p, base_val.p);
        else
          GLSNS_CUDA(ctx, cudaMemcpyAsync(base_val.p, ctx->val.p, sizeof(double) * nnz_b,
                                          cudaMemcpyDeviceToDevice, ctx->stream));
      }
    if (fill > 0)
      {
        std::vector<int64_t> a2p((size_t)nnz_b);
        for (int64_t i = 0; i < n; ++i)
          {
            int64_t t = prow[i];
            for (int64_t k = ctx->base_rowptr[i]; k < ctx->base_rowptr[i + 1]; ++k)
              {
                while (pcol[(size_t)t] != ctx->base_col[(size_t)k])
                  ++t; // (both sorted over the block, ghosts in the same order behind it)
                a2p[(size_t)k] = t++;
              }
          }
        GLSNS_TRY(dev_upload(ctx, ctx->a2p, a2p.data(), a2p.size()));
        GLSNS_CUDA(ctx, cudaMemsetAsync(new_val.p, 0, sizeof(double) * nnz_p, ctx->stream));
        if (nnz_b)
          scatter_values_kernel<<<gb, 256, 0, ctx->stream>>>(nnz_b, ctx->a2p.p, base_val.p, new_val.p);
        GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // a2p goes out of scope
      }
    else
      {
        ctx->a2p.release();
        if (nnz_b)
          GLSNS_CUDA(ctx, cudaMemcpyAsync(new_val.p, base_val.p, sizeof(double) * nnz_b,
                                          cudaMemcpyDeviceToDevice, ctx->stream));
      }
    GLSNS_CUDA(ctx, cudaGetLastError());
    GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    base_val.release();
    ctx->val.release();
    ctx->val = new_val; // (plain handle: pointer + size)
    ctx->lu.release();
    ctx->nnz = nnz_p;
    GLSNS_TRY(dev_upload(ctx, ctx->rowptr, nrow, (size_t)n + 1));
    GLSNS_TRY(dev_alloc(ctx, ctx->col, (size_t)nnz_p + 8)); // (padding: see glsns_set_mesh)
    if (nnz_p)
      GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->col.p, ncol, sizeof(int32_t) * nnz_p, cudaMemcpyHostToDevice,
                                      ctx->stream));
    GLSNS_TRY(ilu_analyse(ctx, nrow, ncol));
    GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->epoch    = 0;
    ctx->ilu_fill = fill;
    ctx->have_ilu = false;
    if (fill == 0)
      {
        ctx->base_rowptr.clear(), ctx->base_rowptr.shrink_to_fit();
        ctx->base_col.clear(), ctx->base_col.shrink_to_fit();
      }
    return GLSNS_OK;
  }

  // matrix values as the host sees them (its own pattern), whatever pattern is installed
  glsns_status
  matrix_values_to_host(glsns_context *ctx, const double *dev_padded, double *host_base)
  {
    const int64_t nb = ctx->a2p.p ? ctx->nnz_base : ctx->nnz;
    if (!ctx->a2p.p)
      {
        GLSNS_CUDA(ctx, cudaMemcpyAsync(host_base, dev_padded, sizeof(double) * nb,
                                        cudaMemcpyDeviceToHost, ctx->stream));
        GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return GLSNS_OK;
      }
    DevBuf<double> tmp;
    GLSNS_TRY(dev_alloc(ctx, tmp, (size_t)std::max<int64_t>(nb, 1)));
    if (nb)
      gather_values_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(nb, ctx->a2p.p,
                                                                                   dev_padded, tmp.p);
    GLSNS_CUDA(ctx, cudaMemcpyAsync(host_base, tmp.p, sizeof(double) * nb, cudaMemcpyDeviceToHost,
                                    ctx->stream));
    GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    tmp.release();
    return GLSNS_OK;
  }

  glsns_status
  matrix_values_from_host(glsns_context *ctx, const double *host_base)
  {
    const int64_t nb = ctx->a2p.p ? ctx->nnz_base : ctx->nnz;
    if (!ctx->a2p.p)
      {
        GLSNS_CUDA(ctx, cudaMemcpyAsync(ctx->val.p, host_base, sizeof(double) * nb,
                                        cudaMemcpyHostToDevice, ctx->stream));
        GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return GLSNS_OK;
      }
    DevBuf<double> tmp;
    GLSNS_TRY(dev_upload(ctx, tmp, host_base, (size_t)nb));
    GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->val.p, 0, sizeof(double) * ctx->nnz, ctx->stream));
    if (nb)
      scatter_values_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(nb, ctx->a2p.p,
                                                                                    tmp.p, ctx->val.p);
    GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    tmp.release();
    return GLSNS_OK;
  }

  glsns_status
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
        const int    maxlen   = (ctx->max_row_len + 3) & ~3;
        int          hbits    = 4;
        while ((1 << hbits) < 2 * maxlen)
          ++hbits;
        const size_t per_warp = (size_t)maxlen * 36 + 64 + ((size_t)6 << hbits);
        const int    warps    = (int)std::min<size_t>(16, (size_t)(227 * 1024) / per_warp);
        static const bool by_rows = getenv("GLSNS_ILU_BY_ROWS") != nullptr;
        static const bool by_pivot_rows = getenv("GLSNS_ILU_BY_PIVOT_ROWS") != nullptr;
        const size_t per_warp_r = per_warp + 128;
        const int    warps_r    = (int)std::min<size_t>(10, (size_t)(227 * 1024) / per_warp_r);
        static const bool by_warp_runs = getenv("GLSNS_ILU_BY_WARP_RUNS") != nullptr;
        const size_t      smem_cta     = per_warp + 128;
        if (smem_cta <= (size_t)(227 * 1024) / 2 && ctx->n_groups > 0 && !by_rows && !by_pivot_rows && !by_warp_runs &&
            ctx->grp_first.p && maxlen < 32768)
          {
            // one CTA of four warps per group, a pivot run's columns dealt to its threads
            if (ctx->n_diag_rows)
              ilu_mark_rows_kernel<<<(ctx->n_diag_rows + 255) / 256, 256, 0, ctx->stream>>>(
                ctx->n_diag_rows, ctx->diag_rows.p, ctx->row_done.p, ctx->epoch);
            GLSNS_CUDA(ctx, cudaFuncSetAttribute(ilu_factor_cta_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
            int per_sm = 0;
            GLSNS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ilu_factor_cta_kernel, FC_T, smem_cta));
            const int grid = (int)std::min<int64_t>(ctx->n_groups, (int64_t)ctx->n_sm * std::max(per_sm, 1));
            GLSNS_TRY(dev_alloc(ctx, ctx->rowdesc, (size_t)n));
            ilu_rowdesc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
              n, ctx->rowptr.p, ctx->diag_pos.p, ctx->grp_first.p, ctx->rowdesc.p);
            ilu_factor_cta_kernel<<<grid, FC_T, smem_cta, ctx->stream>>>(
              ctx->n_groups, ctx->fgroups.p, n, ctx->rowptr.p, ctx->col.p, ctx->diag_pos.p,
              ctx->rowdesc.p, ctx->lu.p, ctx->row_done.p, ctx->epoch, ctx->counters.p, maxlen, hbits);
            ctx->kernel_launches += 4;
          }
        else if (warps_r >= 2 && ctx->n_groups > 0 && !by_rows && !by_pivot_rows && ctx->grp_first.p && maxlen < 65536)
          {
            // rows of a group share one warp, and so do the pivot rows of a run
            if (ctx->n_diag_rows)
              ilu_mark_rows_kernel<<<(ctx->n_diag_rows + 255) / 256, 256, 0, ctx->stream>>>(
                ctx->n_diag_rows, ctx->diag_rows.p, ctx->row_done.p, ctx->epoch);
            const size_t smem = warps_r * per_warp_r;
            GLSNS_CUDA(ctx, cudaFuncSetAttribute(ilu_factor_runs_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem));
            const int grid = (int)std::min<int64_t>((ctx->n_groups + warps_r - 1) / warps_r, ctx->n_sm);
            GLSNS_TRY(dev_alloc(ctx, ctx->rowdesc, (size_t)n));
            ilu_rowdesc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
              n, ctx->rowptr.p, ctx->diag_pos.p, ctx->grp_first.p, ctx->rowdesc.p);
            ilu_factor_runs_kernel<<<grid, warps_r * 32, smem, ctx->stream>>>(
              ctx->n_groups, ctx->fgroups.p, n, ctx->rowptr.p, ctx->col.p, ctx->diag_pos.p,
              ctx->rowdesc.p, ctx->lu.p, ctx->row_done.p, ctx->epoch, ctx->counters.p, maxlen, hbits);
            ctx->kernel_launches += 4;
          }
        else if (warps >= 2 && ctx->n_groups > 0 && !by_rows)
          {
            // rows of a group share one warp (the general case)
            if (ctx->n_diag_rows)
              ilu_mark_rows_kernel<<<(ctx->n_diag_rows + 255) / 256, 256, 0, ctx->stream>>>(
                ctx->n_diag_rows, ctx->diag_rows.p, ctx->row_done.p, ctx->epoch);
            const size_t smem = warps * per_warp;
            GLSNS_CUDA(ctx, cudaFuncSetAttribute(ilu_factor_groups_kernel,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem));
            const int grid = (int)std::min<int64_t>((ctx->n_groups + warps - 1) / warps, ctx->n_sm);
            ilu_factor_groups_kernel<<<grid, warps * 32, smem, ctx->stream>>>(
              ctx->n_groups, ctx->fgroups.p, n, ctx->rowptr.p, ctx->col.p, ctx->diag_pos.p,
              ctx->lu.p, ctx->row_done.p, ctx->epoch, ctx->counters.p, maxlen, hbits);
            ctx->kernel_launches += 3;
          }
        else
          {
            // rows too long to stage a group in shared memory: one row per warp
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
      }
    GLSNS_CUDA(ctx, cudaGetLastError());
    GLSNS_TRY(trsv_prepare(ctx));
    return check_counters(ctx, "ILU(0) factorisation");
  }

  void
  iluk_symbolic_host(const int64_t n, const int64_t *rowptr, const int32_t *col, const int fill,
                     std::vector<int64_t> &prow, std::vector<int32_t> &pcol)
  {
    iluk_symbolic(n, rowptr, col, fill, prow, pcol);
  }
} // namespace glsns

// Host-only entry point (no device needed): the level-of-fill pattern glsns_setup_ilu installs
// for `ilu preconditioner fill` = fill, for tests of the symbolic phase.  Returns the number of
// entries; out_row_ptr [n+1] and out_col_idx [returned count] may be NULL to query the size.
extern "C" int64_t
glsnsh_iluk_pattern(int64_t n, const int64_t *row_ptr, const int32_t *col_idx, int32_t fill,
                    int64_t *out_row_ptr, int32_t *out_col_idx)
{
  if (n < 0 || !row_ptr || (row_ptr[n] && !col_idx) || fill < 0 || fill > 200)
    return -1;
  std::vector<int64_t> prow;
  std::vector<int32_t> pcol;
  glsns::iluk_symbolic_host(n, row_ptr, col_idx, fill, prow, pcol);
  if (out_row_ptr)
    memcpy(out_row_ptr, prow.data(), sizeof(int64_t) * (size_t)(n + 1));
  if (out_col_idx)
    memcpy(out_col_idx, pcol.data(), sizeof(int32_t) * pcol.size());
  return (int64_t)pcol.size();
}
