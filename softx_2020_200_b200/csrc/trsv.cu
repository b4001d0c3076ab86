// ILU(0) application z = U^-1 L^-1 r: the two triangular solves of every GMRES
// iteration (what Ifpack_ILU::ApplyInverse does behind TrilinosWrappers::
// PreconditionILU, reference call site source/solvers/gls_navier_stokes.cc:1276-1279).
//
// The solves keep the host's row ordering exactly (the preconditioner, and with it
// the GMRES iteration count, must be the reference's), so the parallelism is what
// the dependency DAG of that ordering offers.  Under Cuthill-McKee that DAG is
// deep and narrow (3D Q2-Q2, 32^3 cells: 2742 levels of ~100 mesh nodes each) and
// ~90 % of its critical edges join a node to the node numbered just before it.  A
// level-synchronous or row-per-warp solve pays one L2 store->load hop (0.36 us on
// B200, tools/hop_latency.cu) plus a warp reduction per level; this kernel is built
// to take HBM, the hop and most instructions off that critical path:
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
//     several levels of slack.  Every warp's list is sorted by level, which makes the
//     waiting deadlock free (the blocked group of lowest level would wait on a group
//     of lower level that is some warp's current or earlier item) provided all warps
//     are resident — hence the cooperative launch, which refuses instead of hanging.
//   * STREAMS.  After every factorisation the factor entries are re-packed, per
//     sweep, into one contiguous byte stream per warp in exactly the order that warp
//     consumes them: per item (<= 128 entries of one group) a 16-byte header, the
//     inverted diagonal, the in-group triangle, the window couplings, the column
//     indices and the values.  The solve then reads HBM strictly sequentially, and
//     one lane moves a whole item into the warp's shared-memory ring with a single
//     bulk copy (cp.async.bulk, completion on an mbarrier) several items ahead of
//     the one being solved: ~220 KB x 148 SMs of factor data in flight, ~5 issue
//     slots per item instead of ~150 for per-entry copies.
//   * The solution vector itself carries readiness: it is pre-filled with an
//     all-ones NaN pattern and a consumer re-reads an entry until it has been
//     overwritten (no flags, no fences).
#include <stdio.h>
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
    constexpr int TS_CH  = 128;        // entries per item
    constexpr int TS_U   = TS_CH / 32; // entries per lane and item
    constexpr int TS_WIN = 16;         // rows of the chain kept in registers
    // item blob (bytes, 16-byte aligned):
    //   header 16: r0 | flags | fmask | bytes/16 of the blob NSLOT items ahead
    //   last item of a group only: dinv 32 | tri 48 | fwd 8*G*WIN
    //   col 4*pad4(entries) | val 8*m*pad4(entries)
    constexpr int TS_OFF_M    = 16;                                  // solver blob: T^-1, [4][4]
    constexpr int TS_OFF_G    = TS_OFF_M + 8 * TRSV_G * TRSV_G;      // solver blob: T^-1 F, [4][WIN]
    constexpr int TS_OFF_COL1 = TS_OFF_G + 8 * TRSV_G * TS_WIN;      // 656 = size of a solver blob
    constexpr int TS_OFF_COL0 = 16;                                 // (other items)
    constexpr int TS_MAX_SLOTS = 9;
    constexpr int TS_SMEM_MAX  = 227 * 1024;

    // item flags: bits 0-2 rows in the group (m); bit 8 last helper item of its group;
    // bit 9 solver item (one per group: inverted diagonal, in-group triangle, couplings
    // to the chain window; fmask bit d = couples to the chain row at distance d);
    // bits 16.. entries of a helper item
    constexpr int IT_LAST   = 1 << 8;
    constexpr int IT_SOLVER = 1 << 9;
    constexpr int IT_GUEST  = 1 << 10; // solver item of a group that runs through no window
    // bits 12-13 of a solver item: which of the team's windows its chain runs through
    constexpr int TS_NWIN   = 4;

    __host__ __device__ inline int
    pad4(int v)
    {
      return (v + 3) & ~3;
    }
    __host__ __device__ inline int
    blob_bytes(int flags)
    {
      const int m = flags & 7, c = pad4(flags >> 16);
      return (flags & IT_SOLVER) ? TS_OFF_COL1 : TS_OFF_COL0 + 4 * c + 8 * m * c;
    }
    // per-warp entry of the stream directory (64 bytes)
    struct TrsvWarpDir
    {
      int64_t offset;  // byte offset of the warp's first blob in the stream
      int32_t n_items;
      int32_t first16[TS_MAX_SLOTS]; // bytes/16 of the first NSLOT blobs
      int32_t pad[4];
    };
    static_assert(sizeof(TrsvWarpDir) == 64, "directory entry");

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
    mbar_init(void *bar, int count)
    {
      const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    }
    __device__ __forceinline__ void
    mbar_expect_tx(void *bar, unsigned bytes)
    {
      const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes)
                   : "memory");
    }
    __device__ __forceinline__ bool
    mbar_try_wait(void *bar, unsigned parity)
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
    // one whole item: global -> shared, completion counted in bytes on the mbarrier;
    // the stream is read once per sweep, so it is marked evict-first in L2
    __device__ __forceinline__ void
    bulk_load(void *smem_dst, const void *gsrc, unsigned bytes, void *bar, unsigned long long policy)
    {
      const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
      const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
                   "[%0], [%1], %2, [%3], %4;" ::"r"(d),
                   "l"(gsrc), "r"(bytes), "r"(b), "l"(policy)
                   : "memory");
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

    // ---- stream packing: one warp per item -------------------------------------
    // static part (once per sparsity pattern): header and column indices
    __global__ void __launch_bounds__(256)
    trsv_pack_static_kernel(const int64_t n_items, const TrsvItem *__restrict__ items,
                            const int64_t *__restrict__ blob_off, const int32_t *__restrict__ next16,
                            const int32_t *__restrict__ col, unsigned char *__restrict__ stream)
    {
      const int64_t it   = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
      const int     lane = threadIdx.x & 31;
      if (it >= n_items)
        return;
      const TrsvItem d = items[it];
      unsigned char *B = stream + blob_off[it];
      if (lane == 0)
        *reinterpret_cast<int4 *>(B) = make_int4(d.r0, d.flags, d.fmask, next16[it]);
      if (d.flags & IT_SOLVER)
        return;
      const int cntc = d.flags >> 16, cp = pad4(cntc);
      int32_t  *bc   = reinterpret_cast<int32_t *>(B + TS_OFF_COL0);
      for (int k = lane; k < cp; k += 32)
        bc[k] = k < cntc ? col[d.rs0 + d.e_off + k] : d.r0; // padding: any valid index
    }

    // values (after every factorisation)
    template <bool UPPER>
    __global__ void __launch_bounds__(256)
    trsv_pack_values_kernel(const int64_t n_items, const TrsvItem *__restrict__ items,
                            const int64_t *__restrict__ blob_off, const double *__restrict__ lu,
                            unsigned char *__restrict__ stream)
    {
      const int64_t it   = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
      const int     lane = threadIdx.x & 31;
      if (it >= n_items)
        return;
      const TrsvItem d = items[it];
      unsigned char *B = stream + blob_off[it];
      const int      m = d.flags & 7, cntc = d.flags >> 16, cp = pad4(cntc);
      if (!(d.flags & IT_SOLVER))
        {
          double *bv = reinterpret_cast<double *>(B + TS_OFF_COL0 + 4 * cp);
          for (int a = 0; a < m; ++a)
            for (int k = lane; k < cp; k += 32)
              bv[a * cp + k] = k < cntc ? __ldcs(lu + d.rs0 + (int64_t)a * d.len + d.e_off + k) : 0.0;
          return;
        }
      // Solver blob.  With T the group's own 4x4 triangle (unit lower / upper with the
      // diagonal) and F its couplings to the 16 chain rows of the window, the recurrence
      //     T out = -(totals + F w)
      // is stored solved for out:  M = T^-1 and G = T^-1 F, so that the solver warp does
      // one 4 x 20 product per group and no substitution (the explicit inverse of a 4x4
      // block costs a few ulps times its condition number; Ifpack substitutes).
      // Lanes 0..3 solve for the columns of M, lanes 4..19 for those of G.
      double T[TRSV_G][TRSV_G];
#pragma unroll
      for (int a = 0; a < TRSV_G; ++a)
#pragma unroll
        for (int b = 0; b < TRSV_G; ++b)
          {
            const bool in = a < m && b < m && (UPPER ? b >= a : b < a);
            T[a][b]       = in ? lu[d.rs0 + (int64_t)a * d.len + d.nlow + b] : (a == b ? 1.0 : 0.0);
          }
      if (!UPPER)
#pragma unroll
        for (int a = 0; a < TRSV_G; ++a)
          T[a][a] = 1.0;
      const unsigned fmask = (unsigned)d.fmask;
      double         rhs[TRSV_G];
      const int      dd = lane - TRSV_G; // window distance of this lane's G column
#pragma unroll
      for (int a = 0; a < TRSV_G; ++a)
        {
          double v = 0;
          if (lane < TRSV_G)
            v = lane == a ? 1.0 : 0.0;
          else if (dd < TS_WIN && a < m && (fmask & (1u << dd)))
            {
              const int before = __popc(fmask & ((1u << dd) - 1u));
              const int pos    = UPPER ? d.nlow + m + before : d.nlow - 1 - before;
              v                = lu[d.rs0 + (int64_t)a * d.len + pos];
            }
          rhs[a] = v;
        }
      double sol[TRSV_G];
      if (UPPER)
        {
#pragma unroll
          for (int a = TRSV_G - 1; a >= 0; --a)
            {
              double v = rhs[a];
#pragma unroll
              for (int b = TRSV_G - 1; b > a; --b)
                v -= T[a][b] * sol[b];
              sol[a] = v / T[a][a];
            }
        }
      else
        {
#pragma unroll
          for (int a = 0; a < TRSV_G; ++a)
            {
              double v = rhs[a];
#pragma unroll
              for (int b = 0; b < a; ++b)
                v -= T[a][b] * sol[b];
              sol[a] = v;
            }
        }
      double *M = reinterpret_cast<double *>(B + TS_OFF_M);
      double *G = reinterpret_cast<double *>(B + TS_OFF_G);
#pragma unroll
      for (int a = 0; a < TRSV_G; ++a)
        {
          if (lane < TRSV_G)
            M[a * TRSV_G + lane] = sol[a];
          else if (dd < TS_WIN)
            G[a * TS_WIN + dd] = sol[a];
        }
    }

    // ---- the solve ---------------------------------------------------------------
    // A TEAM of 1 + K warps owns one list of groups (chains, in the order the host
    // scheduled them):
    //   * K helper warps take the groups round-robin.  A helper streams the items of
    //     its groups (column indices + factor entries) through its own ring, gathers
    //     the solution entries they refer to (re-reading until they are there),
    //     multiplies, folds in the right-hand side, reduces over the warp and posts
    //     the four totals in the team's mailbox.  None of this depends on the chain,
    //     so it runs ahead of it, on several groups at once.
    //   * the solver warp is the chain's recurrence and nothing else: mailbox totals
    //     minus the couplings to the last 16 rows of the chain (kept in a tiny
    //     shared-memory window, coefficients from the solver's own stream), the 4x4
    //     triangle, publish.  ~300 issue slots per group.
    constexpr int TS_NSH  = 5;  // helper ring slots
    constexpr int TS_NSS  = 3;  // solver ring slots
    constexpr int TS_SCH  = 3;  // groups per solver ring slot (one bulk copy, one barrier wait)
    constexpr int TS_MBOX = 8;  // mailbox entries per team
    constexpr int TS_HSLOT = TS_OFF_COL0 + 4 * TS_CH + 8 * TRSV_G * TS_CH; // 4624
    constexpr int TS_SSLOT = TS_OFF_COL1;                                   // 608
    constexpr int TS_TEAM_AREA = 976; // windows 4 x 128 | mailbox 256 | solved counter 16 | barriers
    static_assert(TS_MBOX == 8 && TS_NWIN == 4 && 128 * TS_NWIN + 32 * TS_MBOX + 16 + 8 * (4 * TS_NSH + TS_NSS) <= TS_TEAM_AREA,
                  "team area");

    template <bool UPPER>
    __global__ void __launch_bounds__(512, 1)
    trsv_team_kernel(const TrsvWarpDir *__restrict__ dir, const unsigned char *__restrict__ stream,
                     const double *__restrict__ rhs_vec, double *x, int *counters,
                     unsigned long long *trace, const int64_t trace_n, const int K)
    {
      extern __shared__ __align__(128) unsigned char smem_all[];
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const int team_in_cta = warp / (K + 1), role = warp - team_in_cta * (K + 1);
      const int n_teams_cta = (blockDim.x >> 5) / (K + 1);
      if (team_in_cta >= n_teams_cta)
        return;
      const int64_t team      = (int64_t)team_in_cta * gridDim.x + blockIdx.x; // consecutive lists on different SMs
      const size_t  team_smem = (size_t)K * TS_NSH * TS_HSLOT + (size_t)TS_NSS * TS_SCH * TS_SSLOT + TS_TEAM_AREA;
      unsigned char *T0   = smem_all + (size_t)team_in_cta * team_smem;
      unsigned char *area = T0 + (size_t)K * TS_NSH * TS_HSLOT + (size_t)TS_NSS * TS_SCH * TS_SSLOT;
      double        *wsm_all = reinterpret_cast<double *>(area);         // [TS_NWIN][16] chain windows by row & 15
      double        *mbox = reinterpret_cast<double *>(area + 128 * TS_NWIN); // [TS_MBOX][4], all-ones = empty
      volatile int  *done = reinterpret_cast<volatile int *>(area + 128 * TS_NWIN + 32 * TS_MBOX); // groups solved
      unsigned long long *bars_all =
        reinterpret_cast<unsigned long long *>(area + 128 * TS_NWIN + 32 * TS_MBOX + 16);
      const TrsvWarpDir  *D        = dir + team * (K + 1) + role;
      const int64_t       n_items  = D->n_items;
      unsigned long long  policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      // team-wide initialisation by the solver warp (window zero, mailbox empty), made
      // visible to the helpers by the one CTA barrier of the kernel
      if (role == 0)
        {
          wsm_all[lane] = wsm_all[lane + 32] = 0.0; // TS_NWIN * 16 = 64 entries
          reinterpret_cast<unsigned long long *>(mbox)[lane] = SENTINEL; // TS_MBOX * 4 = 32 entries, all empty
          if (lane == 0)
            *done = 0;
        }
      __syncthreads();
      if (n_items == 0)
        return;
      const unsigned char *src = stream + D->offset;

      if (role == 0)
        {
          // =============================== solver ===============================
          unsigned char      *ring = T0 + (size_t)K * TS_NSH * TS_HSLOT;
          unsigned long long *bars = bars_all + K * TS_NSH;
          int64_t             n_iss = 0;
          if (lane == 0)
            {
              for (int s = 0; s < TS_NSS; ++s)
                mbar_init(bars + s, 1);
              asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              for (int s = 0; s < TS_NSS && n_iss < n_items; ++s)
                {
                  const unsigned bytes = (unsigned)min((int64_t)TS_SCH, n_items - n_iss) * TS_SSLOT;
                  mbar_expect_tx(bars + s, bytes);
                  bulk_load(ring + (size_t)s * TS_SCH * TS_SSLOT, src, bytes, bars + s, policy);
                  src += bytes;
                  n_iss += TS_SCH;
                }
            }
          __syncwarp();
          int      slot = 0;
          unsigned phase = 0;
          long long tstage[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64(); // debugging aid (trace)
#define TS_TICK(k)                              \
  if (trace)                                    \
    {                                           \
      const long long now_ = clock64();         \
      tstage[k] += now_ - tlast;                \
      tlast = now_;                             \
    }
          int sub = 0; // group inside the slot
          for (int64_t g = 0; g < n_items; ++g)
            {
              if (sub == 0)
                while (!mbar_try_wait(bars + slot, phase))
                  ;
              TS_TICK(0)
              const unsigned char *S  = ring + ((size_t)slot * TS_SCH + sub) * TS_SSLOT;
              const int4           h  = *reinterpret_cast<const int4 *>(S);
              const int            r0 = h.x, m = h.y & 7;
              double              *wsm = wsm_all + 16 * ((h.y >> 12) & (TS_NWIN - 1));
              // out = -(M totals + G w): lane (a, j) = (lane >> 3, lane & 7) takes columns j
              // and j+8 of G (window rows r0-1-d / r0+m+d at distances d = j, j+8) and, for
              // j < 4, column j of M.  Coefficients of absent couplings are exactly zero
              // and the window only ever holds finite numbers: no masking.
              const int     a = lane >> 3, j = lane & 7;
              const double *G = reinterpret_cast<const double *>(S + TS_OFF_G) + a * TS_WIN;
              const double  mm = reinterpret_cast<const double *>(S + TS_OFF_M)[a * TRSV_G + (j & 3)];
              const double  w0 = wsm[(UPPER ? r0 + m + j : r0 - 1 - j) & 15];
              const double  w1 = wsm[(UPPER ? r0 + m + j + 8 : r0 - 9 - j) & 15];
              double        p  = G[j] * w0 + G[j + 8] * w1;
              TS_TICK(1)
              // totals of everything else (minus the right-hand side), from a helper: the
              // mailbox entry carries its own readiness (all-ones pattern = empty)
              const int mb = (int)(g & (TS_MBOX - 1));
              volatile unsigned long long *mv =
                reinterpret_cast<volatile unsigned long long *>(mbox + mb * 4);
              unsigned long long tv;
              {
                long long spins = 0;
                while (!__all_sync(0xffffffffu, (tv = mv[lane & 3]) != SENTINEL))
                  if ((++spins & 4095) == 0 &&
                      (spins > 64 * SPIN_LIMIT || *(volatile int *)(counters + 1) != 0))
                    {
                      atomicExch(&counters[1], 2);
                      break;
                    }
              }
              TS_TICK(2)
              if (j < 4)
                p += mm * __longlong_as_double((long long)tv);
              __syncwarp();
              if (lane < TRSV_G) // hand the entry back: empty it, then let group g + TS_MBOX in
                mv[lane] = SENTINEL;
              if (lane == 0)
                *done = (int)g + 1;
              p += __shfl_xor_sync(0xffffffffu, p, 4);
              p += __shfl_xor_sync(0xffffffffu, p, 2);
              p += __shfl_xor_sync(0xffffffffu, p, 1);
              if (j == 0 && a < m)
                {
                  const double v = -p;
                  st_result(x + r0 + a, v);
                  if (!(h.y & IT_GUEST))
                    wsm[(r0 + a) & 15] = v;
                  if (trace) // debugging aid (glsns_ilu_apply_trace): when was the row published
                    {
                      unsigned long long tns;
                      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
                      trace[r0 + a] = tns;
                    }
                }
              __syncwarp();
              TS_TICK(3)
              if (++sub == TS_SCH)
                { // the slot is used up: refill it with the groups TS_NSS slots ahead
                  sub = 0;
                  if (lane == 0 && n_iss < n_items)
                    {
                      const unsigned bytes = (unsigned)min((int64_t)TS_SCH, n_items - n_iss) * TS_SSLOT;
                      mbar_expect_tx(bars + slot, bytes);
                      bulk_load(ring + (size_t)slot * TS_SCH * TS_SSLOT, src, bytes, bars + slot, policy);
                      src += bytes;
                      n_iss += TS_SCH;
                    }
                  slot = slot + 1 == TS_NSS ? 0 : slot + 1;
                  phase ^= slot == 0;
                }
              TS_TICK(4)
            }
          if (trace && lane == 0 && (team + 1) * 8 <= trace_n)
            { // cycles in: ring wait, window product, mailbox wait, triangle+publish, release+refill
              for (int k = 0; k < 6; ++k)
                trace[2 * trace_n + team * 8 + k] = (unsigned long long)tstage[k];
              trace[2 * trace_n + team * 8 + 6] = (unsigned long long)n_items;
            }
#undef TS_TICK
          return;
        }

      // =============================== helper ===============================
      const int           hid  = role - 1;
      unsigned char      *ring = T0 + (size_t)hid * TS_NSH * TS_HSLOT;
      unsigned long long *bars = bars_all + hid * TS_NSH;
      int64_t             n_iss = 0;
      if (lane == 0)
        {
          for (int s = 0; s < TS_NSH; ++s)
            mbar_init(bars + s, 1);
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          for (int s = 0; s < TS_NSH && s < n_items; ++s)
            {
              const unsigned bytes = 16u * (unsigned)D->first16[s];
              mbar_expect_tx(bars + s, bytes);
              bulk_load(ring + (size_t)s * TS_HSLOT, src, bytes, bars + s, policy);
              src += bytes;
              ++n_iss;
            }
        }
      __syncwarp();
      // pipeline registers: three rotating sets (no copies: a copy would wait for the
      // loads in flight), selected at compile time by the unrolled loop
      int32_t            cS[3][TS_U];
      unsigned long long bS[3][TS_U];
      unsigned           pS[3] = {0, 0, 0};
      double             rS[3] = {0, 0, 0}; // right-hand side of row r0 + lane
      double             acc[TRSV_G];
#pragma unroll
      for (int a = 0; a < TRSV_G; ++a)
        acc[a] = 0;
      int      slotG = 0, slotR = 0;
      unsigned phaseG = 0;
      int64_t  seq    = hid; // position of the helper's current group in the team's list
      // one pipeline step: G(it) into set KG, R(it-2) from set (KG+1)%3
      auto step = [&](auto KG, const int64_t it) {
        constexpr int      kg = decltype(KG)::value, kr = (kg + 1) % 3;
        int32_t(&cN)[TS_U]            = cS[kg];
        unsigned long long(&bN)[TS_U] = bS[kg];
        unsigned &pendN               = pS[kg];
        int32_t(&cG)[TS_U]            = cS[kr];
        unsigned long long(&bG)[TS_U] = bS[kr];
        unsigned &pendG               = pS[kr];
        // ---- G(it): wait for the item, request the solution entries it refers to ----
        pendN = 0;
        if (it < n_items)
          {
            while (!mbar_try_wait(bars + slotG, phaseG))
              ;
            const unsigned char *S     = ring + (size_t)slotG * TS_HSLOT;
            const int4           h     = *reinterpret_cast<const int4 *>(S);
            const int            flags = h.y, cntc = flags >> 16;
            const int32_t       *scol  = reinterpret_cast<const int32_t *>(S + TS_OFF_COL0);
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
            if ((flags & IT_LAST) && lane < (flags & 7))
              rS[kg] = rhs_vec[h.x + lane];
            slotG = slotG + 1 == TS_NSH ? 0 : slotG + 1;
            phaseG ^= slotG == 0;
          }
        // ---- R(it-2): entries not there yet are re-read until they are; multiply ----
        if (it >= 2)
          {
            const unsigned char *S     = ring + (size_t)slotR * TS_HSLOT;
            const int4           h     = *reinterpret_cast<const int4 *>(S);
            const int            flags = h.y, m = flags & 7, cp = pad4(flags >> 16);
            const double        *sval  = reinterpret_cast<const double *>(S + TS_OFF_COL0 + 4 * cp);
            long long            spins = 0;
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
                          acc[a] += sval[a * cp + k] * xv;
                    }
                if (!__any_sync(0xffffffffu, pendG != 0))
                  break;
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
            if (flags & IT_LAST)
              {
                // the group is complete: fold in the right-hand side, total over the
                // warp, post in the mailbox once the solver has freed the entry
#pragma unroll
                for (int a = 0; a < TRSV_G; ++a)
                  if (lane == a && a < m)
                    acc[a] -= rS[kr];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                  for (int a = 0; a < TRSV_G; ++a)
                    acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
                const int mb = (int)(seq & (TS_MBOX - 1));
                volatile unsigned long long *mv =
                  reinterpret_cast<volatile unsigned long long *>(mbox + mb * 4);
                spins = 0;
                // the entry is this group's once the solver is within TS_MBOX groups of it
                while ((int64_t)*done + TS_MBOX <= seq)
                  if ((++spins & 4095) == 0 &&
                      (spins > 64 * SPIN_LIMIT || *(volatile int *)(counters + 1) != 0))
                    {
                      atomicExch(&counters[1], 2);
                      break;
                    }
                if (lane < TRSV_G)
                  {
                    double v = acc[0];
#pragma unroll
                    for (int a = 1; a < TRSV_G; ++a)
                      if (lane == a)
                        v = acc[a];
                    unsigned long long b = (unsigned long long)__double_as_longlong(v);
                    if (b == SENTINEL)
                      b = 0x7FF8000000000000ull;
                    mv[lane] = b;
                  }
#pragma unroll
                for (int a = 0; a < TRSV_G; ++a)
                  acc[a] = 0;
                seq += K;
              }
            __syncwarp(); // every lane is done with the slot before it is refilled
            if (lane == 0 && n_iss < n_items)
              {
                const unsigned bytes = 16u * (unsigned)h.w; // size of the item TS_NSH ahead
                mbar_expect_tx(bars + slotR, bytes);
                bulk_load(ring + (size_t)slotR * TS_HSLOT, src, bytes, bars + slotR, policy);
                src += bytes;
                ++n_iss;
              }
            slotR = slotR + 1 == TS_NSH ? 0 : slotR + 1;
          }
      };
      for (int64_t it = 0; it < n_items + 2; it += 3)
        {
          step(std::integral_constant<int, 0>(), it);
          if (it + 1 < n_items + 2)
            step(std::integral_constant<int, 1>(), it + 1);
          if (it + 2 < n_items + 2)
            step(std::integral_constant<int, 2>(), it + 2);
        }
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
      int teams = 3, helpers = 3; // per SM
    };

    size_t
    team_smem_bytes(int helpers)
    {
      return (size_t)helpers * TS_NSH * TS_HSLOT + (size_t)TS_NSS * TS_SCH * TS_SSLOT + TS_TEAM_AREA;
    }

    TrsvConfig
    trsv_config()
    {
      static TrsvConfig c = [] {
        TrsvConfig t;
        if (getenv("GLSNS_TRSV_TEAMS"))
          t.teams = atoi(getenv("GLSNS_TRSV_TEAMS"));
        if (getenv("GLSNS_TRSV_HELPERS"))
          t.helpers = atoi(getenv("GLSNS_TRSV_HELPERS"));
        t.helpers = std::max(1, std::min(4, t.helpers));
        t.teams   = std::max(1, std::min(16 / (t.helpers + 1), t.teams));
        while (t.teams > 1 && t.teams * team_smem_bytes(t.helpers) > (size_t)TS_SMEM_MAX)
          --t.teams;
        return t;
      }();
      return c;
    }

    template <bool UPPER>
    glsns_status
    launch_team(glsns_context *ctx, const TrsvWarpDir *dir, const unsigned char *stream,
                const double *rhs, double *x, unsigned long long *trace)
    {
      const TrsvConfig cfg  = trsv_config();
      const size_t     smem = cfg.teams * team_smem_bytes(cfg.helpers);
      auto             kern = trsv_team_kernel<UPPER>;
      GLSNS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
      int  *counters = ctx->counters.p;
      int   K        = cfg.helpers;
      void *args[]   = {(void *)&dir,      (void *)&stream, (void *)&rhs,          (void *)&x,
                        (void *)&counters, (void *)&trace,  (void *)&ctx->n_owned, (void *)&K};
      GLSNS_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kern, dim3(ctx->trsv_grid),
                                                  dim3(cfg.teams * (cfg.helpers + 1) * 32), args,
                                                  smem, ctx->stream));
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
    { // the SpMV takes every row, group by group, in row order (sparse.cu)
      std::vector<int2> sg;
      sg.reserve((size_t)ng + diag_rows.size());
      size_t dr = 0;
      for (int64_t g = 0; g < ng; ++g)
        {
          while (dr < diag_rows.size() && diag_rows[dr] < grp_ptr[g])
            sg.push_back(make_int2(diag_rows[dr++], 1));
          sg.push_back(make_int2(grp_ptr[g], grp_m[g]));
        }
      while (dr < diag_rows.size())
        sg.push_back(make_int2(diag_rows[dr++], 1));
      ctx->n_sgroups = (int64_t)sg.size();
      GLSNS_TRY(dev_upload(ctx, ctx->sgroups, sg.data(), sg.size()));
      GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    GLSNS_TRY(dev_upload(ctx, ctx->diag_rows, diag_rows.data(), diag_rows.size()));

    const TrsvConfig cfg = trsv_config();
    ctx->trsv_grid       = ctx->n_sm;
    const int64_t NW     = (int64_t)ctx->trsv_grid * cfg.teams; // teams: one chain list each
    const int32_t max_level_gap = getenv("GLSNS_TRSV_GAP") ? atoi(getenv("GLSNS_TRSV_GAP")) : 2;
    const int     K      = cfg.helpers;

    std::vector<int32_t> glev(ng), fmask(ng), gapv(ng), cnt(ng), e_off(ng), order(ng), warp_of(ng),
      slot_of(ng);
    std::vector<uint8_t> link(ng), has_succ(ng), is_primary(ng);

    // One sweep: levels, chain links, level-ordered schedule on NW warps, item lists.
    auto schedule = [&](const bool upper, TrsvSweep &sw, int32_t &n_levels) -> glsns_status {
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
          // a predecessor solved long before this group needs it is no link: its rows
          // reach the group through L2 in time, and the team is free for another chain
          if (link[g] && l - glev[p] > max_level_gap)
            link[g] = 0;
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
      if (!upper)
        { // the factorisation takes the groups in the same order (sparse.cu)
          std::vector<int2> fg((size_t)ng);
          for (int64_t t = 0; t < ng; ++t)
            fg[(size_t)t] = make_int2(grp_ptr[order[t]], grp_m[order[t]]);
          GLSNS_TRY(dev_upload(ctx, ctx->fgroups, fg.data(), fg.size()));
          GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
      // list scheduling: a chained group follows its predecessor on the same warp;
      // a chain head takes the warp that has been free the longest
      // Every team has TS_NWIN chain slots (one window each).  A chain head takes a slot
      // of the least loaded team: `bucket[k]` lists teams believed to run k chains
      // (entries are validated when popped).
      std::vector<int32_t> last_of(NW * TS_NWIN, -1), chain_edge(NW * TS_NWIN, 0);
      std::vector<uint8_t> n_live(NW, 0);
      std::deque<int32_t>  bucket[TS_NWIN];
      for (int64_t w = 0; w < NW; ++w)
        bucket[0].push_back((int32_t)w);
      auto take_slot = [&]() -> int32_t { // team * TS_NWIN + slot, or -1
        for (int k = 0; k < TS_NWIN; ++k)
          while (!bucket[k].empty())
            {
              const int32_t w = bucket[k].front();
              bucket[k].pop_front();
              if (n_live[w] != k)
                continue; // stale entry
              for (int sl = 0; sl < TS_NWIN; ++sl)
                if (last_of[(int64_t)w * TS_NWIN + sl] == -1)
                  {
                    ++n_live[w];
                    if (n_live[w] < TS_NWIN)
                      bucket[n_live[w]].push_back(w);
                    return w * TS_NWIN + sl;
                  }
            }
        return -1;
      };
      int64_t rr = 0, n_heads = 0, n_steals = 0, n_interrupted = 0;
      for (int64_t t = 0; t < ng; ++t)
        {
          const int32_t g = order[t];
          const int64_t p = upper ? (int64_t)g + 1 : (int64_t)g - 1;
          const int32_t r0 = grp_ptr[g], m = grp_m[g];
          // A chain runs through one window of its team.  When every window of every
          // team is taken, a group is placed on a busy team as a GUEST: solved in list
          // order without touching any window, so the chains there keep their forwarding.
          int32_t ws; // team * TS_NWIN + slot
          bool    chained = false, primary = true;
          if (link[g] && is_primary[p] && last_of[slot_of[p]] == p)
            {
              ws      = slot_of[p];
              chained = true;
            }
          else if ((ws = take_slot()) >= 0)
            n_interrupted += link[g]; // chain head, or the successor of a guest: a fresh window
          else
            {
              ws      = (link[g] ? warp_of[p] : (int32_t)(rr++ % NW)) * TS_NWIN;
              primary = false;
              ++n_steals;
            }
          const int32_t w = ws / TS_NWIN;
          n_heads += !link[g];
          is_primary[g] = primary;
          if (!chained && primary) // a new chain starts here: first row (lower) / end row (upper)
            chain_edge[ws] = upper ? r0 + m : r0;
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
                    if (d >= TS_WIN || col[k] < chain_edge[ws] || grp_of[col[k]] < 0)
                      break;
                    fm |= 1u << d;
                    ++nf;
                  }
              else
                for (int64_t k = kb; k < ke; ++k)
                  {
                    const int32_t d = col[k] - (r0 + m);
                    if (d >= TS_WIN || col[k] >= chain_edge[ws] || grp_of[col[k]] < 0)
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
          slot_of[g] = ws;
          if (primary)
            {
              if (has_succ[g])
                last_of[ws] = g; // (last group of the chain in this window so far)
              else
                { // the chain ends here: the window is free again
                  last_of[ws] = -1;
                  --n_live[w];
                  bucket[n_live[w]].push_back(w);
                }
            }
        }
      if (getenv("GLSNS_TRSV_DEBUG"))
        fprintf(stderr,
                "trsv_analyse %s: %lld groups, %d levels, %lld teams, %lld chain heads, %lld taken "
                "as guests on busy teams, %lld chains restarted\n",
                upper ? "upper" : "lower", (long long)ng, nlev + 1, (long long)NW, (long long)n_heads,
                (long long)n_steals, (long long)n_interrupted);
      // item lists: per team one solver list (one item per group) and K helper lists
      // (the groups round-robin, <= TS_CH entries per item)
      const int64_t        NWARP = NW * (K + 1);
      std::vector<int64_t> n_it(NWARP + 1, 0), team_cnt(NW, 0);
      std::vector<int32_t> seq_in_team(ng);
      for (int64_t t = 0; t < ng; ++t)
        {
          const int32_t g  = order[t];
          const int64_t tm = warp_of[g];
          const int64_t j  = team_cnt[tm]++;
          seq_in_team[g]   = (int32_t)j;
          n_it[tm * (K + 1) + 0 + 1] += 1;
          n_it[tm * (K + 1) + 1 + (j % K) + 1] += std::max<int64_t>(1, (cnt[g] + TS_CH - 1) / TS_CH);
        }
      for (int64_t w = 0; w < NWARP; ++w)
        n_it[w + 1] += n_it[w];
      std::vector<TrsvItem> items((size_t)n_it[NWARP]);
      std::vector<int64_t>  fill(n_it.begin(), n_it.end() - 1);
      for (int64_t t = 0; t < ng; ++t)
        {
          const int32_t g = order[t];
          const int64_t i = grp_ptr[g];
          const int32_t m = grp_m[g], len = (int32_t)(rowptr[i + 1] - rowptr[i]);
          const int64_t tm = warp_of[g], j = seq_in_team[g];
          {
            TrsvItem &it = items[(size_t)fill[tm * (K + 1)]++];
            it.rs0 = rowptr[i], it.r0 = (int32_t)i, it.len = len, it.e_off = 0;
            it.flags = m | IT_SOLVER | (is_primary[g] ? 0 : IT_GUEST) | ((slot_of[g] % TS_NWIN) << 12);
            it.nlow  = (int32_t)(diag[i] - rowptr[i]);
            it.fmask = fmask[g];
          }
          const int32_t nchunk = std::max(1, (cnt[g] + TS_CH - 1) / TS_CH);
          for (int32_t c = 0; c < nchunk; ++c)
            {
              TrsvItem  &it   = items[(size_t)fill[tm * (K + 1) + 1 + (j % K)]++];
              const int  cntc = std::max(0, std::min(TS_CH, cnt[g] - c * TS_CH));
              const bool last = c == nchunk - 1;
              it.rs0   = rowptr[i];
              it.r0    = (int32_t)i;
              it.len   = len;
              it.e_off = e_off[g] + c * TS_CH;
              it.flags = m | (last ? IT_LAST : 0) | (cntc << 16);
              it.nlow  = (int32_t)(diag[i] - rowptr[i]);
              it.fmask = 0;
            }
        }
      std::vector<int32_t> &row_warp = upper ? ctx->trsv_row_warp_u : ctx->trsv_row_warp_l;
      row_warp.assign((size_t)n, -1);
      for (int64_t g = 0; g < ng; ++g)
        for (int32_t a = 0; a < grp_m[g]; ++a)
          row_warp[grp_ptr[g] + a] = warp_of[g] | (fmask[g] ? 1 << 30 : 0);
      // stream layout: the blobs of one warp back to back, warps one after another
      const int64_t        nit = n_it[NWARP];
      std::vector<int64_t> blob_off((size_t)nit);
      std::vector<int32_t> next16((size_t)nit, 0);
      std::vector<TrsvWarpDir> dirv((size_t)NWARP);
      int64_t              off = 0;
      for (int64_t w = 0; w < NWARP; ++w)
        {
          TrsvWarpDir &D = dirv[w];
          memset(&D, 0, sizeof(D));
          D.offset  = off;
          D.n_items = (int32_t)(n_it[w + 1] - n_it[w]);
          for (int64_t k = n_it[w]; k < n_it[w + 1]; ++k)
            {
              const int32_t b16 = blob_bytes(items[(size_t)k].flags) / 16;
              blob_off[(size_t)k] = off;
              off += 16 * (int64_t)b16;
              const int64_t j = k - n_it[w];
              if (j < TS_NSH)
                D.first16[j] = b16;
              else
                next16[(size_t)(k - TS_NSH)] = b16; // (helper lists; the solver's blobs have one size)
            }
        }
      sw.n_items      = nit;
      sw.stream_bytes = off;
      GLSNS_TRY(dev_upload(ctx, sw.items, items.data(), items.size()));
      GLSNS_TRY(dev_upload(ctx, sw.blob_off, blob_off.data(), blob_off.size()));
      GLSNS_TRY(dev_upload(ctx, sw.next16, next16.data(), next16.size()));
      GLSNS_TRY(dev_upload(ctx, sw.dir, reinterpret_cast<const unsigned char *>(dirv.data()),
                           dirv.size() * sizeof(TrsvWarpDir)));
      GLSNS_TRY(dev_alloc(ctx, sw.stream, (size_t)std::max<int64_t>(off, 16)));
      if (nit)
        {
          trsv_pack_static_kernel<<<(unsigned)((nit * 32 + 255) / 256), 256, 0, ctx->stream>>>(
            nit, sw.items.p, sw.blob_off.p, sw.next16.p, ctx->col.p, sw.stream.p);
          ctx->kernel_launches++;
          GLSNS_CUDA(ctx, cudaGetLastError());
        }
      GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // host vectors go out of scope
      return GLSNS_OK;
    };
    GLSNS_TRY(schedule(false, ctx->trsv_l, ctx->levels_l));
    GLSNS_TRY(schedule(true, ctx->trsv_u, ctx->levels_u));
    return GLSNS_OK;
  }

  // after every factorisation: the inverted diagonal of U (Ifpack keeps it too) and
  // the factor values in stream order
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
    if (ctx->trsv_l.n_items)
      {
        const int64_t nit = ctx->trsv_l.n_items;
        trsv_pack_values_kernel<false><<<(unsigned)((nit * 32 + 255) / 256), 256, 0, ctx->stream>>>(
          nit, ctx->trsv_l.items.p, ctx->trsv_l.blob_off.p, ctx->lu.p, ctx->trsv_l.stream.p);
        ctx->kernel_launches++;
      }
    if (ctx->trsv_u.n_items)
      {
        const int64_t nit = ctx->trsv_u.n_items;
        trsv_pack_values_kernel<true><<<(unsigned)((nit * 32 + 255) / 256), 256, 0, ctx->stream>>>(
          nit, ctx->trsv_u.items.p, ctx->trsv_u.blob_off.p, ctx->lu.p, ctx->trsv_u.stream.p);
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
        GLSNS_TRY(launch_team<false>(ctx, reinterpret_cast<const TrsvWarpDir *>(ctx->trsv_l.dir.p),
                                      ctx->trsv_l.stream.p, r, ctx->ytmp.p, trace));
        GLSNS_TRY(launch_team<true>(ctx, reinterpret_cast<const TrsvWarpDir *>(ctx->trsv_u.dir.p),
                                     ctx->trsv_u.stream.p, ctx->ytmp.p, z,
                                    trace ? trace + n : nullptr));
      }
    GLSNS_CUDA(ctx, cudaGetLastError());
    return GLSNS_OK;
  }
} // namespace glsns
