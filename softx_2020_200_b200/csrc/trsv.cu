// ILU(0) application z = U^-1 L^-1 r: the two triangular solves of every GMRES
// iteration (what Ifpack_ILU::ApplyInverse does behind TrilinosWrappers::
// PreconditionILU, reference call site source/solvers/gls_navier_stokes.cc:1276-1279).
//
// The solves keep the host's row ordering exactly (the preconditioner, and with it
// the GMRES iteration count, must be the reference's), so the parallelism is what
// the dependency DAG of that ordering offers.  Under Cuthill-McKee that DAG is
// deep and narrow (3D Q2-Q2, 32^3 cells: 2742 levels of ~100 mesh nodes each) and
// ~88 % of its critical edges join a node to the node numbered just before it.  A
// level-synchronous or row-per-warp solve pays one L2 store->load hop (0.36 us on
// B200, tools/hop_latency.cu) plus a warp reduction per level; this kernel is built
// to take both HBM and that hop off the critical path:
//
//   * GROUPS.  Up to 4 consecutive rows with identical column patterns (the dim+1
//     dofs of a mesh node) are solved together: one index stream, a 4x4 triangle.
//   * CHAINS.  The host schedules the groups on the resident warps level by level
//     (trsv_analyse): a group whose predecessor in the numbering is one of its
//     dependencies goes to the warp that solves that predecessor, right behind it.
//     The solutions of the last 16 rows of the chain stay in the warp's registers
//     (one per lane, a shift register over the row distance), so the entries that
//     couple a group to its recent predecessors never wait for L2; only the
//     dependencies on other chains travel through L2, and most of those have
//     several levels of slack.  Every warp's list is sorted by level, which is what makes the
//     waiting deadlock free (the blocked group of lowest level would wait on a group
//     of lower level that is some warp's current or earlier item) provided all warps
//     are resident — hence the cooperative launch, which refuses instead of hanging.
//   * RING.  Each warp streams the column indices and factor entries of its next
//     items from HBM into a private shared-memory ring with cp.async, several items
//     ahead of the one it is solving (~200 KB x 148 SMs in flight), and gathers the
//     solution entries of item i+1 while it reduces item i and finishes item i-1.
//   * The solution vector itself carries readiness: it is pre-filled with an
//     all-ones NaN pattern and a consumer re-reads an entry until it has been
//     overwritten (no flags, no fences).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <type_traits>

#include "context.h"

namespace glsns
{
  namespace
  {
    constexpr unsigned long long SENTINEL   = 0xFFFFFFFFFFFFFFFFull;
    constexpr long long          SPIN_LIMIT = 1ll << 21; // a bug guard (seconds), never reached in a correct run

    constexpr int TRSV_G = 4;          // rows per group
    constexpr int TS_CH  = 128;        // entries per ring slot
    constexpr int TS_U   = TS_CH / 32; // entries per lane and slot
    constexpr int TS_WIN = 16;         // rows of the chain kept in registers
    // slot layout (bytes): header 16 | rhs 32 | dinv 32 | tri 128 | fwd 8*G*WIN | col 4*CH | val 8*G*CH
    constexpr int TS_OFF_RHS  = 16;
    constexpr int TS_OFF_DINV = 48;
    constexpr int TS_OFF_TRI  = 80;
    constexpr int TS_OFF_FWD  = 208;
    constexpr int TS_OFF_COL  = TS_OFF_FWD + 8 * TRSV_G * TS_WIN;
    constexpr int TS_OFF_VAL  = TS_OFF_COL + 4 * TS_CH;
    constexpr int TS_SLOT     = TS_OFF_VAL + 8 * TRSV_G * TS_CH; // 5328
    constexpr int TS_REC      = 32; // item records staged per warp (1 KB)
    constexpr int TS_SMEM_MAX = 227 * 1024;

    // item flags
    constexpr int IT_LAST = 1 << 8; // last item of its group: finish and store
    // bits 0-2: rows in the group (m); bits 4-7: diagonal-only rows between the chain
    // predecessor and this group; bits 16..: entries.  TrsvItem::fmask, last item
    // only: bit d = the group couples to the chain row at distance d (in registers)

    __device__ __forceinline__ unsigned long long
    ld_relaxed_u64(const double *p)
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
    __device__ __forceinline__ void
    cp_async4(void *smem_dst, const void *gsrc)
    {
      const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
    }
    __device__ __forceinline__ void
    cp_async8(void *smem_dst, const void *gsrc)
    {
      const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    }
    __device__ __forceinline__ void
    cp_async16(void *smem_dst, const void *gsrc)
    {
      const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    }
    __device__ __forceinline__ void
    cp_async_commit()
    {
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    template <int N>
    __device__ __forceinline__ void
    cp_async_wait()
    {
      asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
    }

    // rows whose pattern is the diagonal alone (constrained dofs) depend on nothing:
    // they are solved up front, before the sweeps start
    __global__ void __launch_bounds__(256)
    trsv_diag_rows_kernel(const int32_t n_rows, const int32_t *__restrict__ rows,
                          const double *__restrict__ dinv, const double *__restrict__ r,
                          double *__restrict__ y, double *__restrict__ z)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i < n_rows)
        {
          const int32_t row = rows[i];
          const double  v   = r[row];
          y[row]            = v;
          z[row]            = v * dinv[row];
        }
    }

    // One warp = one list of items (TrsvItem, <= TS_CH entries of one group each), in
    // the order the host scheduled them.  Three-stage software pipeline over the items:
    //   G(i+3): column indices from the ring, solution entries requested from L2
    //   B(i)  : last item of a group: solve the in-group triangle, publish the
    //           solution, shift it into the register window
    //   R(i+1): entries that were not there yet are re-read until they are; multiply;
    //           last item of a group: add the coupling to the register window,
    //           warp-reduce
    template <bool UPPER, int NSLOT>
    __global__ void __launch_bounds__(256, 1)
    trsv_chain_kernel(const int64_t *__restrict__ warp_ptr, const TrsvItem *__restrict__ items,
                      const int32_t *__restrict__ col, const double *__restrict__ lu,
                      const double *__restrict__ dinv, const double *__restrict__ rhs_vec,
                      double *x, int *counters, unsigned long long *trace, const int64_t trace_n)
    {
      static_assert(NSLOT >= 5, "the pipeline holds four items besides the ones in flight");
      extern __shared__ __align__(16) unsigned char ring_all[];
      const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const int64_t  w    = (int64_t)warp * gridDim.x + blockIdx.x; // consecutive lists on different SMs
      unsigned char *ring = ring_all + (size_t)warp * NSLOT * TS_SLOT;
      const int64_t  i0 = warp_ptr[w], n_items = warp_ptr[w + 1] - i0;
      if (n_items == 0)
        return;
      const TrsvItem *my = items + i0;

      // ---- issue side ----
      // The item records themselves are prefetched too: 32 of them live in shared
      // memory, refilled by halves with cp.async 16 items before they are needed.
      int64_t n_iss = 0;
      int4   *recs  = reinterpret_cast<int4 *>(ring_all + (size_t)(blockDim.x >> 5) * NSLOT * TS_SLOT) +
                    warp * (2 * TS_REC);
      auto fetch_recs = [&](int64_t first) { // 16 records = 32 x 16 bytes, one per lane
        const int64_t q = first * 2 + lane;  // index in int4 units
        if (q < n_items * 2)
          cp_async16(recs + (q & (2 * TS_REC - 1)), reinterpret_cast<const int4 *>(my) + q);
      };
      fetch_recs(0);
      fetch_recs(TS_REC / 2);
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
      auto issue = [&](int slot) {
        if (n_iss < n_items)
          {
            unsigned char *S     = ring + (size_t)slot * TS_SLOT;
            const int4     ra = recs[(n_iss & (TS_REC - 1)) * 2], rb = recs[(n_iss & (TS_REC - 1)) * 2 + 1];
            if ((n_iss & (TS_REC / 2 - 1)) == 0 && n_iss > 0)
              fetch_recs(n_iss + TS_REC / 2); // the half just left; joins this item's group
            const int64_t  rs0   = ((int64_t)(unsigned)ra.x) | ((int64_t)ra.y << 32);
            const int      r0    = ra.z, len = ra.w;
            const int      e_off = rb.x, flags = rb.y, nlow = rb.z;
            const unsigned fmask = (unsigned)rb.w;
            const int      m = flags & 7, cntc = flags >> 16;
            const int64_t  e0   = rs0 + e_off;
            int32_t       *scol = reinterpret_cast<int32_t *>(S + TS_OFF_COL);
            double        *sval = reinterpret_cast<double *>(S + TS_OFF_VAL);
#pragma unroll
            for (int u = 0; u < TS_U; ++u)
              {
                const int k = lane + 32 * u;
                if (k < cntc)
                  {
                    cp_async4(scol + k, col + e0 + k);
#pragma unroll
                    for (int a = 0; a < TRSV_G; ++a)
                      if (a < m)
                        cp_async8(sval + a * TS_CH + k, lu + e0 + (int64_t)a * len + k);
                  }
              }
            if (flags & IT_LAST)
              {
                if (lane < 16)
                  {
                    const int a = lane >> 2, b = lane & 3;
                    if (a < m && b < m && (UPPER ? b > a : b < a))
                      cp_async8(S + TS_OFF_TRI + 8 * lane, lu + rs0 + (int64_t)a * len + nlow + b);
                  }
                else if (lane < 16 + m)
                  cp_async8(S + TS_OFF_RHS + 8 * (lane - 16), rhs_vec + r0 + (lane - 16));
                else if (UPPER && lane >= 24 && lane < 24 + m)
                  cp_async8(S + TS_OFF_DINV + 8 * (lane - 24), dinv + r0 + (lane - 24));
                // coupling to the chain rows held in registers: lane (h, d) copies rows
                // a = h (mod 2) of the entry at distance d
                const int d = lane & 15, hh = lane >> 4;
                if (fmask & (1u << d))
                  {
                    const int before = __popc(fmask & ((1u << d) - 1u));
                    const int pos    = UPPER ? nlow + m + before : nlow - 1 - before;
#pragma unroll
                    for (int a = 0; a < TRSV_G; ++a)
                      if ((a & 1) == hh && a < m)
                        cp_async8(S + TS_OFF_FWD + 8 * (a * TS_WIN + d),
                                  lu + rs0 + (int64_t)a * len + pos);
                  }
              }
            if (lane == 0)
              {
                int4 h;
                h.x = r0, h.y = flags, h.z = rb.w /* fmask */, h.w = 0;
                *reinterpret_cast<int4 *>(S) = h;
              }
            ++n_iss;
          }
        cp_async_commit();
      };
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        issue(s);

      // ---- pipeline registers: three rotating sets (no copies: a copy would wait for
      //      the loads in flight), selected at compile time by the unrolled loop ----
      int32_t            cS[3][TS_U];
      unsigned long long bS[3][TS_U];
      unsigned           pS[3] = {0, 0, 0};
      double             acc[TRSV_G], accB[TRSV_G];
      double             win = 0; // solution of the chain row at distance (lane & 15)
#pragma unroll
      for (int a = 0; a < TRSV_G; ++a)
        acc[a] = accB[a] = 0;
      int slotG = 0, slotR = 0, slotB = 0;

      // one pipeline step: G(it) into set KG, B(it-3), R(it-2) from set (KG+1)%3
      auto step = [&](auto KG, const int64_t it) {
        constexpr int      kg = decltype(KG)::value, kr = (kg + 1) % 3;
        int32_t(&cN)[TS_U]            = cS[kg];
        unsigned long long(&bN)[TS_U] = bS[kg];
        unsigned &pendN               = pS[kg];
        int32_t(&cG)[TS_U]            = cS[kr];
        unsigned long long(&bG)[TS_U] = bS[kr];
        unsigned &pendG               = pS[kr];
          // ================= G(it) =================
          pendN = 0;
          if (it < n_items)
            {
              cp_async_wait<NSLOT - 4>();
              __syncwarp();
              const unsigned char *S    = ring + (size_t)slotG * TS_SLOT;
              const int            cntc = reinterpret_cast<const int4 *>(S)->y >> 16;
              const int32_t       *scol = reinterpret_cast<const int32_t *>(S + TS_OFF_COL);
#pragma unroll
              for (int u = 0; u < TS_U; ++u)
                {
                  const int k = lane + 32 * u;
                  if (k < cntc)
                    {
                      cN[u] = scol[k];
                      pendN |= 1u << u;
                    }
                }
#pragma unroll
              for (int u = 0; u < TS_U; ++u)
                if (pendN & (1u << u))
                  bN[u] = ld_relaxed_u64(x + cN[u]);
              slotG = slotG + 1 == NSLOT ? 0 : slotG + 1;
            }
          // ================= B(it-3) =================
          // (before R: the item R waits for may depend, through other warps, on the
          //  group this stage publishes)
          if (it >= 3)
            {
              const unsigned char *S     = ring + (size_t)slotB * TS_SLOT;
              const int4           h     = *reinterpret_cast<const int4 *>(S);
              const int            flags = h.y;
              if (flags & IT_LAST)
                {
                  const int     m = flags & 7, r0 = h.x;
                  const double *rhs = reinterpret_cast<const double *>(S + TS_OFF_RHS);
                  const double *tri = reinterpret_cast<const double *>(S + TS_OFF_TRI);
                  const double *di  = reinterpret_cast<const double *>(S + TS_OFF_DINV);
                  double        out[TRSV_G];
#pragma unroll
                  for (int a = 0; a < TRSV_G; ++a)
                    out[a] = a < m ? rhs[a] - accB[a] : 0.0;
                  if (UPPER)
                    {
#pragma unroll
                      for (int a = TRSV_G - 1; a >= 0; --a)
                        if (a < m)
                          {
                            double v = out[a];
#pragma unroll
                            for (int b = TRSV_G - 1; b >= 0; --b)
                              if (b > a && b < m)
                                v -= tri[a * 4 + b] * out[b];
                            out[a] = v * di[a]; // Ifpack stores and applies the inverted diagonal
                          }
                    }
                  else
                    {
#pragma unroll
                      for (int a = 0; a < TRSV_G; ++a)
                        if (a < m)
                          {
                            double v = out[a];
#pragma unroll
                            for (int b = 0; b < TRSV_G; ++b)
                              if (b < a)
                                v -= tri[a * 4 + b] * out[b];
                            out[a] = v;
                          }
                    }
                  if (lane < m)
                    {
                      double v = out[0];
#pragma unroll
                      for (int a = 1; a < TRSV_G; ++a)
                        if (lane == a)
                          v = out[a];
                      st_result(x + r0 + lane, v);
                      if (trace) // debugging aid (glsns_ilu_apply_trace): when was the row published
                        {
                          unsigned long long tns;
                          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
                          trace[r0 + lane] = tns;
                        }
                    }
                  // shift the solved rows into the register window: distance d of the
                  // next group of the chain is m-1-a (lower sweep) or a (upper sweep)
                  {
                    const int    d  = lane & 15;
                    const double up = __shfl_up_sync(0xffffffffu, win, m, 16);
                    double       nv = up;
#pragma unroll
                    for (int a = 0; a < TRSV_G; ++a)
                      if (a < m && d == (UPPER ? a : m - 1 - a))
                        nv = out[a];
                    win = nv;
                  }
                }
              __syncwarp(); // every lane is done with the slot before it is refilled
              issue(slotB);
              slotB = slotB + 1 == NSLOT ? 0 : slotB + 1;
            }
          // ================= R(it-2) =================
          bool lastR = false;
          if (it >= 2 && it <= n_items + 1)
            {
              const unsigned char *S     = ring + (size_t)slotR * TS_SLOT;
              const int            flags = reinterpret_cast<const int4 *>(S)->y;
              const int            m     = flags & 7;
              const double        *sval  = reinterpret_cast<const double *>(S + TS_OFF_VAL);
              lastR                      = flags & IT_LAST;
              long long spins            = 0;
              int       dbg_near         = 0x7fffffff;
              for (;;)
                {
#pragma unroll
                  for (int u = 0; u < TS_U; ++u)
                    if ((pendG & (1u << u)) && bG[u] != SENTINEL)
                      {
                        pendG &= ~(1u << u);
                        const double xv = __longlong_as_double((long long)bG[u]);
                        const int    k  = lane + 32 * u;
#pragma unroll
                        for (int a = 0; a < TRSV_G; ++a)
                          if (a < m)
                            acc[a] += sval[a * TS_CH + k] * xv;
                      }
                  if (!__any_sync(0xffffffffu, pendG != 0))
                    break;
                  if (trace && spins == 0)
                    { // debugging aid: which entries made this item wait (row distance)
                      const int r0 = reinterpret_cast<const int4 *>(S)->x;
#pragma unroll
                      for (int u = 0; u < TS_U; ++u)
                        if (pendG & (1u << u))
                          dbg_near = min(dbg_near, abs(cG[u] - r0));
                    }
#pragma unroll
                  for (int u = 0; u < TS_U; ++u)
                    if (pendG & (1u << u))
                      bG[u] = ld_relaxed_u64(x + cG[u]);
                  // bug guard: give up after seconds of waiting, or as soon as another
                  // warp has given up (the host reports GLSNS_ERR_CUDA)
                  if ((++spins & 1023) == 0 &&
                      (spins > SPIN_LIMIT || *(volatile int *)(counters + 1) != 0))
                    {
                      atomicExch(&counters[1], 2);
                      break;
                    }
                }
              if (trace)
                {
#pragma unroll
                  for (int o = 16; o > 0; o >>= 1)
                    dbg_near = min(dbg_near, __shfl_xor_sync(0xffffffffu, dbg_near, o));
                  if (lane == 0 && lastR)
                    {
                      const int r0         = reinterpret_cast<const int4 *>(S)->x;
                      trace[2 * trace_n + r0] = ((unsigned long long)spins << 32) | (unsigned)dbg_near;
                    }
                }
              slotR = slotR + 1 == NSLOT ? 0 : slotR + 1;
            }
          // ---- hand-over R -> B: totals of a finished group in every lane ----
          if (lastR)
            {
              // item it-2 closes its group: add the coupling to the chain rows in
              // registers (all solved by now: their B stages ran at or before this
              // iteration), then total over the warp
              {
                const unsigned char *S  = ring + (size_t)(slotR == 0 ? NSLOT - 1 : slotR - 1) * TS_SLOT;
                const int4           h  = *reinterpret_cast<const int4 *>(S);
                const int            m  = h.y & 7, d = lane & 15, hh = lane >> 4;
                const double        *fw = reinterpret_cast<const double *>(S + TS_OFF_FWD);
                const int            gap = (h.y >> 4) & 15;
                if (gap) // diagonal-only rows between the predecessor and this group
                  win = __shfl_up_sync(0xffffffffu, win, gap, 16);
                if ((unsigned)h.z & (1u << d))
                  {
#pragma unroll
                    for (int a = 0; a < TRSV_G; ++a)
                      if ((a & 1) == hh && a < m)
                        acc[a] += fw[a * TS_WIN + d] * win;
                  }
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int a = 0; a < TRSV_G; ++a)
                  acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
#pragma unroll
              for (int a = 0; a < TRSV_G; ++a)
                {
                  accB[a] = acc[a];
                  acc[a]  = 0;
                }
            }
      };
      for (int64_t it = 0; it < n_items + 3; it += 3)
        {
          step(std::integral_constant<int, 0>(), it);
          if (it + 1 < n_items + 3)
            step(std::integral_constant<int, 1>(), it + 1);
          if (it + 2 < n_items + 3)
            step(std::integral_constant<int, 2>(), it + 2);
        }
      cp_async_wait<0>();
    }

    __global__ void __launch_bounds__(256)
    inv_diag_kernel(const int64_t n, const int64_t *__restrict__ diag_pos,
                    const double *__restrict__ lu, double *__restrict__ dinv)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i < n)
        dinv[i] = 1.0 / lu[diag_pos[i]];
    }

    struct TrsvConfig
    {
      int warps = 6, nslot = 7;
    };

    TrsvConfig
    trsv_config()
    {
      static TrsvConfig c = [] {
        TrsvConfig t;
        if (getenv("GLSNS_TRSV_WARPS"))
          t.warps = atoi(getenv("GLSNS_TRSV_WARPS"));
        if (getenv("GLSNS_TRSV_NSLOT"))
          t.nslot = atoi(getenv("GLSNS_TRSV_NSLOT"));
        t.warps = std::max(1, std::min(8, t.warps));
        if (t.nslot < 5 || t.nslot > 9)
          t.nslot = 7;
        while ((size_t)t.warps * (t.nslot * TS_SLOT + TS_REC * 32) > (size_t)TS_SMEM_MAX)
          --t.warps;
        return t;
      }();
      return c;
    }

    template <bool UPPER>
    glsns_status
    launch_chain(glsns_context *ctx, const int64_t *warp_ptr, const TrsvItem *items,
                 const double *rhs, double *x, unsigned long long *trace)
    {
      const TrsvConfig cfg  = trsv_config();
      const size_t     smem = (size_t)cfg.warps * (cfg.nslot * TS_SLOT + TS_REC * 32);
      void (*kern)(const int64_t *, const TrsvItem *, const int32_t *, const double *,
                   const double *, const double *, double *, int *, unsigned long long *,
                   const int64_t) = nullptr;
      switch (cfg.nslot)
        {
          case 5: kern = trsv_chain_kernel<UPPER, 5>; break;
          case 6: kern = trsv_chain_kernel<UPPER, 6>; break;
          case 7: kern = trsv_chain_kernel<UPPER, 7>; break;
          case 8: kern = trsv_chain_kernel<UPPER, 8>; break;
          default: kern = trsv_chain_kernel<UPPER, 9>; break;
        }
      GLSNS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
      const int32_t *col = ctx->col.p;
      const double  *lu = ctx->lu.p, *dinv = ctx->dinv.p;
      int           *counters = ctx->counters.p;
      void *args[] = {(void *)&warp_ptr, (void *)&items, (void *)&col, (void *)&lu,
                      (void *)&dinv,     (void *)&rhs,   (void *)&x,   (void *)&counters,
                      (void *)&trace,    (void *)&ctx->n_owned};
      GLSNS_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kern, dim3(ctx->trsv_grid),
                                                  dim3(cfg.warps * 32), args, smem, ctx->stream));
      ctx->kernel_launches++;
      return GLSNS_OK;
    }
  } // namespace

  // Groups of consecutive rows with identical column patterns, their dependency
  // levels in the lower and the upper sweep, and the per-warp item lists of both
  // sweeps.  Host work, once per sparsity pattern.
  glsns_status
  trsv_analyse(glsns_context *ctx, const int64_t *rowptr, const int32_t *col,
               const int64_t *diag)
  {
    const int64_t        n = ctx->n_owned;
    std::vector<int32_t> grp_ptr, grp_of(n), diag_rows;
    grp_ptr.reserve(n / 2 + 2);
    for (int64_t i = 0; i < n;)
      {
        const int64_t len = rowptr[i + 1] - rowptr[i];
        if (len == 1)
          { // diagonal-only row: solved by the elementwise kernel
            diag_rows.push_back((int32_t)i);
            grp_of[i] = -1;
            ++i;
            continue;
          }
        grp_ptr.push_back((int32_t)i);
        int64_t j = i + 1;
        while (j < n && j - i < TRSV_G && rowptr[j + 1] - rowptr[j] == len &&
               memcmp(col + rowptr[i], col + rowptr[j], sizeof(int32_t) * len) == 0)
          ++j;
        for (int64_t r = i; r < j; ++r)
          grp_of[r] = (int32_t)(grp_ptr.size() - 1);
        i = j;
      }
    const int64_t        ng = (int64_t)grp_ptr.size();
    std::vector<int32_t> grp_m(ng);
    for (int64_t g = 0; g < ng; ++g)
      {
        int64_t e = grp_ptr[g] + 1;
        while (e < n && grp_of[e] == g)
          ++e;
        grp_m[g] = (int32_t)(e - grp_ptr[g]);
      }
    ctx->n_groups    = (int32_t)ng;
    ctx->n_diag_rows = (int32_t)diag_rows.size();
    GLSNS_TRY(dev_upload(ctx, ctx->diag_rows, diag_rows.data(), diag_rows.size()));

    const TrsvConfig cfg = trsv_config();
    ctx->trsv_grid       = ctx->n_sm;
    const int64_t NW     = (int64_t)ctx->trsv_grid * cfg.warps;

    std::vector<int32_t> glev(ng), fmask(ng), gapv(ng), cnt(ng), e_off(ng), order(ng), warp_of(ng);
    std::vector<uint8_t> link(ng), has_succ(ng);

    // One sweep: levels, chain links, level-ordered schedule on NW warps, item lists.
    auto schedule = [&](const bool upper, DevBuf<TrsvItem> &d_items, DevBuf<int64_t> &d_ptr,
                        int32_t &n_levels) -> glsns_status {
      // entries of group g this sweep reads, [kb, ke) in CSR offsets of its first row
      auto range = [&](int64_t g, int64_t &kb, int64_t &ke) {
        const int64_t i = grp_ptr[g];
        kb              = upper ? diag[i] + grp_m[g] : rowptr[i];
        ke              = upper ? rowptr[i + 1] : diag[i];
        if (upper)
          while (ke > kb && col[ke - 1] >= n)
            --ke; // ghost columns: outside the diagonal block
      };
      int32_t nlev = 0;
      std::fill(has_succ.begin(), has_succ.end(), 0);
      for (int64_t gi = 0; gi < ng; ++gi)
        {
          const int64_t g = upper ? ng - 1 - gi : gi;
          int64_t       kb, ke;
          range(g, kb, ke);
          int32_t l = 0;
          for (int64_t k = kb; k < ke; ++k)
            {
              const int32_t dg = grp_of[col[k]];
              if (dg >= 0) // (diagonal-only rows are solved before the sweep starts)
                l = std::max(l, glev[dg] + 1);
            }
          glev[g] = l;
          nlev    = std::max(nlev, l);
          // chain link: the group depends on its neighbour in the numbering, with
          // nothing but diagonal-only rows (and fewer than a window of them) in between
          const int64_t p = upper ? g + 1 : g - 1;
          link[g]         = 0;
          if (p >= 0 && p < ng)
            {
              const int64_t gap = upper ? grp_ptr[p] - (grp_ptr[g] + grp_m[g]) :
                                          grp_ptr[g] - (grp_ptr[p] + grp_m[p]);
              if (gap < TS_WIN - 1)
                {
                  if (!upper)
                    for (int64_t k = ke - 1; k >= kb; --k)
                      if (grp_of[col[k]] >= 0)
                        {
                          link[g] = grp_of[col[k]] == p;
                          break;
                        }
                  if (upper)
                    for (int64_t k = kb; k < ke; ++k)
                      if (grp_of[col[k]] >= 0)
                        {
                          link[g] = grp_of[col[k]] == p;
                          break;
                        }
                }
            }
          if (link[g])
            has_succ[p] = 1;
        }
      n_levels = ng ? nlev + 1 : 0;
      // groups by level (within a level in sweep order)
      {
        std::vector<int64_t> start(nlev + 2, 0);
        for (int64_t g = 0; g < ng; ++g)
          start[glev[g] + 1]++;
        for (int32_t l = 0; l <= nlev; ++l)
          start[l + 1] += start[l];
        for (int64_t gi = 0; gi < ng; ++gi)
          {
            const int64_t g         = upper ? ng - 1 - gi : gi;
            order[start[glev[g]]++] = (int32_t)g;
          }
      }
      // list scheduling: a chained group follows its predecessor on the same warp;
      // a chain head takes the warp that has been free the longest
      std::vector<int32_t> last_of(NW, -1), chain_edge(NW, 0);
      std::vector<uint8_t> live(NW, 0), pooled(NW, 1);
      std::vector<int64_t> n_it(NW + 1, 0);
      std::deque<int32_t>  pool;
      for (int64_t w = 0; w < NW; ++w)
        pool.push_back((int32_t)w);
      int64_t rr = 0;
      for (int64_t t = 0; t < ng; ++t)
        {
          const int32_t g = order[t];
          const int64_t p = upper ? (int64_t)g + 1 : (int64_t)g - 1;
          const int32_t r0 = grp_ptr[g], m = grp_m[g];
          int32_t       w;
          bool          chained = false;
          if (link[g])
            {
              w       = warp_of[p];
              chained = last_of[w] == p; // else: interrupted (more chains than warps)
            }
          else if (!pool.empty())
            {
              w = pool.front();
              pool.pop_front();
              pooled[w] = 0;
            }
          else
            w = (int32_t)(rr++ % NW);
          if (!chained) // a new chain starts here: first row (lower) / end row (upper)
            chain_edge[w] = upper ? r0 + m : r0;
          // entries that couple to the rows of this chain still held in registers: the
          // run next to the in-group block, at most TS_WIN rows away
          int64_t kb, ke;
          range(g, kb, ke);
          uint32_t fm = 0;
          int32_t  nf = 0;
          if (chained)
            {
              if (!upper)
                for (int64_t k = ke - 1; k >= kb; --k)
                  {
                    const int32_t d = r0 - 1 - col[k];
                    if (d >= TS_WIN || col[k] < chain_edge[w] || grp_of[col[k]] < 0)
                      break;
                    fm |= 1u << d;
                    ++nf;
                  }
              else
                for (int64_t k = kb; k < ke; ++k)
                  {
                    const int32_t d = col[k] - (r0 + m);
                    if (d >= TS_WIN || col[k] >= chain_edge[w] || grp_of[col[k]] < 0)
                      break;
                    fm |= 1u << d;
                    ++nf;
                  }
              // rows between the predecessor and this group (diagonal-only ones): the
              // kernel shifts its register window by that much first
              gapv[g] = (int32_t)(upper ? grp_ptr[p] - (r0 + m) : r0 - (grp_ptr[p] + grp_m[p]));
            }
          else
            gapv[g] = 0;
          fmask[g] = (int32_t)fm;
          e_off[g] = (int32_t)((upper ? kb + nf : kb) - rowptr[grp_ptr[g]]);
          cnt[g]   = (int32_t)(ke - kb - nf);
          warp_of[g] = w;
          last_of[w] = g;
          live[w]    = has_succ[g];
          n_it[w + 1] += std::max<int64_t>(1, (cnt[g] + TS_CH - 1) / TS_CH);
          if (!live[w] && !pooled[w])
            {
              pool.push_back(w);
              pooled[w] = 1;
            }
        }
      for (int64_t w = 0; w < NW; ++w)
        n_it[w + 1] += n_it[w];
      std::vector<TrsvItem> items((size_t)n_it[NW]);
      std::vector<int64_t>  fill(n_it.begin(), n_it.end() - 1);
      for (int64_t t = 0; t < ng; ++t)
        {
          const int32_t g = order[t];
          const int64_t i = grp_ptr[g];
          const int32_t m = grp_m[g], len = (int32_t)(rowptr[i + 1] - rowptr[i]);
          const int32_t nchunk = std::max(1, (cnt[g] + TS_CH - 1) / TS_CH);
          for (int32_t c = 0; c < nchunk; ++c)
            {
              TrsvItem  &it   = items[(size_t)fill[warp_of[g]]++];
              const int  cntc = std::max(0, std::min(TS_CH, cnt[g] - c * TS_CH));
              const bool last = c == nchunk - 1;
              it.rs0   = rowptr[i];
              it.r0    = (int32_t)i;
              it.len   = len;
              it.e_off = e_off[g] + c * TS_CH;
              it.flags = m | (last ? IT_LAST | (gapv[g] << 4) : 0) | (cntc << 16);
              it.nlow  = (int32_t)(diag[i] - rowptr[i]);
              it.fmask = last ? fmask[g] : 0;
            }
        }
      std::vector<int32_t> &row_warp = upper ? ctx->trsv_row_warp_u : ctx->trsv_row_warp_l;
      row_warp.assign((size_t)n, -1);
      for (int64_t g = 0; g < ng; ++g)
        for (int32_t a = 0; a < grp_m[g]; ++a)
          row_warp[grp_ptr[g] + a] = warp_of[g] | (fmask[g] ? 1 << 30 : 0);
      GLSNS_TRY(dev_upload(ctx, d_items, items.data(), items.size()));
      GLSNS_TRY(dev_upload(ctx, d_ptr, n_it.data(), n_it.size()));
      GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // host vectors go out of scope
      return GLSNS_OK;
    };
    GLSNS_TRY(schedule(false, ctx->items_l, ctx->wptr_l, ctx->levels_l));
    GLSNS_TRY(schedule(true, ctx->items_u, ctx->wptr_u, ctx->levels_u));
    return GLSNS_OK;
  }

  // after every factorisation: the inverted diagonal of U (Ifpack keeps it too)
  glsns_status
  trsv_prepare(glsns_context *ctx)
  {
    const int64_t n = ctx->n_owned;
    GLSNS_TRY(dev_alloc(ctx, ctx->dinv, (size_t)std::max<int64_t>(n, 1)));
    if (n)
      {
        inv_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, ctx->diag_pos.p,
                                                                             ctx->lu.p,
                                                                             ctx->dinv.p);
        ctx->kernel_launches++;
      }
    GLSNS_CUDA(ctx, cudaGetLastError());
    return GLSNS_OK;
  }

  // z = (LU)^-1 r; ctx->ytmp is the intermediate.  Asynchronous on ctx->stream.
  glsns_status
  launch_ilu_apply(glsns_context *ctx, const double *r, double *z, unsigned long long *trace)
  {
    const int64_t n = ctx->n_owned;
    if (n == 0)
      return GLSNS_OK;
    GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->ytmp.p, 0xFF, sizeof(double) * n, ctx->stream));
    GLSNS_CUDA(ctx, cudaMemsetAsync(z, 0xFF, sizeof(double) * n, ctx->stream));
    if (ctx->n_diag_rows)
      {
        trsv_diag_rows_kernel<<<(ctx->n_diag_rows + 255) / 256, 256, 0, ctx->stream>>>(
          ctx->n_diag_rows, ctx->diag_rows.p, ctx->dinv.p, r, ctx->ytmp.p, z);
        ctx->kernel_launches++;
      }
    if (ctx->n_groups)
      {
        GLSNS_TRY(launch_chain<false>(ctx, ctx->wptr_l.p, ctx->items_l.p, r, ctx->ytmp.p, trace));
        GLSNS_TRY(launch_chain<true>(ctx, ctx->wptr_u.p, ctx->items_u.p, ctx->ytmp.p, z,
                                    trace ? trace + n : nullptr));
      }
    GLSNS_CUDA(ctx, cudaGetLastError());
    return GLSNS_OK;
  }
} // namespace glsns
