// ILU application z = U^-1 L^-1 r: the two triangular solves of every Krylov
// iteration (what Ifpack_ILU::ApplyInverse does behind TrilinosWrappers::
// PreconditionILU, reference call site source/solvers/gls_navier_stokes.cc:1276-1279).
//
// The solves keep the host's row ordering exactly (the preconditioner, and with it
// the GMRES iteration count, must be the reference's), so the parallelism is what
// the dependency DAG of that ordering offers.  Under Cuthill-McKee that DAG is
// deep and narrow (3D Q2-Q2, 64^3 cells: 5558 levels of ~390 mesh nodes each) and
// ~90 % of its critical edges join a node to the node numbered just before it.  A
// level-synchronous or row-per-warp solve pays one L2 store->load hop (0.36 us on
// B200, tools/hop_latency.cu) plus a warp reduction per level; this kernel is built
// to take HBM, the hop and most instructions off that critical path:
//
//   * GROUPS.  Up to 4 consecutive rows with identical column patterns (the dim+1
//     dofs of a mesh node) share one index stream.
//   * BLOCKS.  Up to 4 consecutive linked groups (<= 16 rows) are the unit of the
//     recurrence.  Everything a block needs from outside was numbered before its
//     first row, so solving its rows in one step cannot dead-lock; its own triangle T
//     and its couplings F to the 48 chain rows before it are packed SOLVED after every
//     factorisation (M = T^-1, G = T^-1 F), so a block is one 16 x 64 product
//     out = -(M totals + G w) and the chain advances 16 rows per step.  (48 rows, not
//     16: a block couples to the last two or three blocks of its chain, and what is not
//     in the window goes through L2 and a helper -- with a 16-row window the helpers'
//     totals arrived 0.8 us after the chain predecessor in the median block, which tripled
//     the chain step: 8.66 -> 7.89 ms per application at 64^3 cells.)
//   * CHAINS.  The host schedules the blocks on the resident TEAMS level by level
//     (trsv_analyse): a block whose predecessor in the numbering is one of its
//     dependencies goes to the team that solves that predecessor, right behind it,
//     and reads the last 48 rows of the chain from a window in shared memory (the chain's
//     last 64 rows, by row number modulo 64) instead of through L2; only the dependencies
//     on other chains travel through L2.  Every
//     team's list is sorted by a key that grows along every dependency (the block
//     level, or, after the profile-guided pass, the time its inputs were published in
//     a traced run), which makes the waiting dead-lock free (the blocked block with
//     the smallest key waits on a block with a smaller one, which is some team's
//     current or earlier item) provided all teams are resident -- hence the
//     cooperative launch, which refuses instead of hanging.
//   * TEAMS.  1 solver warp (the recurrence and nothing else) + K helper warps (they
//     stream the groups' factor entries, gather the solution entries, reduce, and
//     post the totals in the block's mailbox entry); 2 teams x (1 + 7) warps per SM.
//   * STREAMS.  After every factorisation the factor entries are re-packed, per
//     sweep, into contiguous byte streams per warp in exactly the order that warp
//     consumes them: per helper item (<= 64 entries of one group) a 16-byte header with
//     the column indices in one stream and the values of the group's rows in another
//     (the indices are needed when the solution entries are requested, the values two
//     items later when they are used: two rings, 8 x 272 B and 4 x 2 KB per helper); per
//     solver item (one block) the solved recurrence, lane by lane, with the header of the
//     NEXT block.  The solve then reads HBM strictly sequentially, and one lane moves a
//     whole blob into the warp's shared-memory ring with a single bulk copy
//     (cp.async.bulk, completion on an mbarrier) several items ahead of the one in work.
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
#ifndef GLSNS_TRSV_ITEM
#define GLSNS_TRSV_ITEM 64
#endif
    constexpr int TS_CH  = GLSNS_TRSV_ITEM; // entries per item (a multiple of 32)
    constexpr int TS_U   = TS_CH / 32; // entries per lane and item
#ifndef GLSNS_TRSV_WIN
#define GLSNS_TRSV_WIN 48
#endif
    constexpr int TS_WIN  = GLSNS_TRSV_WIN; // rows of the chain behind a block that its recurrence couples to
    constexpr int TS_HIST = 64;        // rows of a chain kept in its shared-memory window
    constexpr int TS_BG  = 4;          // groups per block (the solver's unit)
    constexpr int TS_BR  = 16;         // rows per block
    constexpr int TS_WH  = TS_WIN / 2; // window rows per lane of a row's lane pair
    constexpr int TS_NC  = TS_WH + 8;  // coefficients per lane: TS_WH of G, 8 of M
    static_assert(TS_BR == TRSV_G * TS_BG && TS_WIN % 8 == 0 && TS_WIN >= 16 && TS_WIN + TS_BR <= TS_HIST,
                  "blocks of 16 rows; the window and the block being written share the history");
    // item blob (bytes, 16-byte aligned), first 16 bytes = header
    //   helper item, two blobs in two streams:
    //                indices: r0 | flags | block position in the team's list << 4 | row offset in the
    //                block | bytes/16 of the index blob TS_NCS items ahead | bytes/16 of the value
    //                blob TS_NVS items ahead << 8; then col 4*pad4(entries)
    //                values: val 8*m*pad4(entries) (at least 16 bytes)
    //   solver item (one per block): r0 of the NEXT block of the list | its flags | rows of the
    //                block TS_MBOX ahead | bytes/16 of the blob TS_SNSLOT items ahead; then the block's
    //                solved recurrence lane by lane: 2R lanes (R = rows padded to 4) x TS_CSTR bytes
    constexpr int TS_OFF_COL0 = 16;
    constexpr int TS_OFF_C    = 16;
    constexpr int TS_CSTR     = 8 * TS_NC + 16; // bytes of one solver lane's coefficients (+ 16: conflict-free 128-bit loads)
#ifndef GLSNS_TRSV_NSLOT
#define GLSNS_TRSV_NSLOT 4
#endif
#ifndef GLSNS_TRSV_THREADS
#define GLSNS_TRSV_THREADS 512
#endif
    constexpr int TS_NCS       = 8;                // ring slots of a helper warp: column indices (+ header)
    constexpr int TS_NVS       = GLSNS_TRSV_NSLOT; // ring slots of a helper warp: factor entries
    constexpr int TS_MAXWARPS  = GLSNS_TRSV_THREADS / 32;
#ifndef GLSNS_TRSV_DEPTH
#define GLSNS_TRSV_DEPTH 2
#endif
    constexpr int TS_D = GLSNS_TRSV_DEPTH; // items a helper has in flight between requesting and using the solution entries
    static_assert(TS_D >= 2 && TS_D <= 4 && TS_D < TS_NCS, "pipeline depth of the helpers");
    constexpr int TS_SNSLOT    = 4; // ring slots of a solver warp
    constexpr int TS_MAX_SLOTS = 9;
    constexpr int TS_SMEM_MAX  = 227 * 1024;

    // item flags: bits 0-4 rows (m: of the group, <= 4, in a helper item; of the block, <= 16,
    // in a solver item); bit 8 last helper item of its group; bit 9 solver item; bit 10 solver
    // item of a block that runs through no window; bits 12-13 of a solver item: which of the
    // team's windows its chain runs through; bits 16.. entries of a helper item
    constexpr int IT_LAST   = 1 << 8;
    constexpr int IT_SOLVER = 1 << 9;
    constexpr int IT_GUEST  = 1 << 10;
    constexpr int TS_NWIN   = 4;

    __host__ __device__ inline int
    pad4(int v)
    {
      return (v + 3) & ~3;
    }
    __host__ __device__ inline int
    blob_bytes(int flags)
    {
      const int m = flags & 31, c = pad4(flags >> 16);
      return (flags & IT_SOLVER) ? TS_OFF_C + TS_CSTR * 2 * pad4(m) : TS_OFF_COL0 + 4 * c;
    }
    __host__ __device__ inline int
    value_blob_bytes(int flags) // helper items only
    {
      const int m = flags & 31, c = pad4(flags >> 16);
      return 8 * m * c > 16 ? 8 * m * c : 16;
    }
    // per-warp entry of the stream directory (64 bytes)
    struct TrsvWarpDir
    {
      int64_t offset;  // byte offset of the warp's first blob in the stream
      int32_t n_items;
      int32_t first16[TS_MAX_SLOTS]; // bytes/16 of the first NSLOT blobs
      int32_t first_m[2];            // solver: rows of its first 8 blocks, 8 bits each
      int32_t first_r0, first_flags; // solver: header of its first block
      int64_t offset_v;              // helper: byte offset of its first value blob
      int32_t firstv16[4];           // helper: bytes/16 of its first TS_NVS value blobs
      int32_t pad[10];
    };
    static_assert(TS_NVS <= 4 && TS_NCS <= TS_MAX_SLOTS, "directory entry");
    static_assert(sizeof(TrsvWarpDir) == 128, "directory entry");

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
        *reinterpret_cast<int4 *>(B) = (d.flags & IT_SOLVER) ? make_int4(d.e_off, d.nlow, d.pad_, next16[it]) :
                                                              make_int4(d.r0, d.flags, d.fmask, next16[it]);
      if (d.flags & IT_SOLVER)
        return;
      const int cntc = d.flags >> 16, cp = pad4(cntc);
      int32_t  *bc   = reinterpret_cast<int32_t *>(B + TS_OFF_COL0);
      for (int k = lane; k < cp; k += 32)
        bc[k] = k < cntc ? col[d.rs0 + d.e_off + k] : d.r0; // padding: any valid index
    }

    // values (after every factorisation)
    template <bool UPPER>
    __global__ void __launch_bounds__(128)
    trsv_pack_values_kernel(const int64_t n_items, const TrsvItem *__restrict__ items,
                            const TrsvItem *__restrict__ gdesc,
                            const int64_t *__restrict__ blob_off, const int64_t *__restrict__ blob_off_v,
                            const double *__restrict__ lu, unsigned char *__restrict__ stream)
    {
      __shared__ double TF[4][TS_BR * TS_BR + TS_BR * TS_WIN]; // per warp: T [16][16], F [16][TS_WIN]
      const int64_t it   = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
      const int     lane = threadIdx.x & 31;
      if (it >= n_items)
        return;
      const TrsvItem d = items[it];
      unsigned char *B = stream + blob_off[it];
      const int      m = d.flags & 31, cntc = d.flags >> 16, cp = pad4(cntc);
      if (!(d.flags & IT_SOLVER))
        {
          double *bv = reinterpret_cast<double *>(stream + blob_off_v[it]);
          if (m * cp == 0 && lane < 2)
            bv[lane] = 0.0;
          for (int a = 0; a < m; ++a)
            for (int k = lane; k < cp; k += 32)
              bv[a * cp + k] = k < cntc ? __ldcs(lu + d.rs0 + (int64_t)a * d.len + d.e_off + k) : 0.0;
          return;
        }
      // Solver blob of a BLOCK (<= 4 consecutive groups of one chain, rows [r0, r0 + m)).
      // With T the block's own triangle (unit lower / upper with the diagonal: the
      // in-group triangles and the couplings between the groups of the block) and F its
      // couplings to the 16 chain rows next to it (the window), the recurrence
      //     T out = -(totals + F w)
      // is stored solved for out:  M = T^-1 and G = T^-1 F, so that the solver warp does
      // one 16 x 32 product per block and no substitution (the explicit inverse of a
      // 16 x 16 triangle costs a few ulps times its condition number; Ifpack substitutes).
      double *T = TF[(threadIdx.x >> 5)], *F = T + TS_BR * TS_BR;
      for (int k = lane; k < TS_BR * TS_BR; k += 32)
        T[k] = (k / TS_BR == k % TS_BR) ? 1.0 : 0.0;
      for (int k = lane; k < TS_BR * TS_WIN; k += 32)
        F[k] = 0.0;
      __syncwarp();
      const int r0b = d.r0;
      for (int q = 0; q < d.len; ++q)
        {
          const TrsvItem gq = gdesc[d.rs0 + q];
          const int      mg = gq.flags & 31, off = gq.r0 - r0b;
          const unsigned long long fm =
            (unsigned long long)(unsigned)gq.fmask | ((unsigned long long)(unsigned)gq.fmask2 << 32);
          for (int a = 0; a < mg; ++a)
            {
              const double *row = lu + gq.rs0 + (int64_t)a * gq.len;
              if (lane < mg && (UPPER ? lane >= a : lane < a))
                T[(off + a) * TS_BR + off + lane] = row[gq.nlow + lane];
              for (int dd = lane; dd < TS_WIN; dd += 32)
                if (fm & (1ull << dd))
                  {
                    const int    before = __popcll(fm & ((1ull << dd) - 1ull));
                    const double v      = row[UPPER ? gq.nlow + mg + before : gq.nlow - 1 - before];
                    const int    tr     = UPPER ? gq.r0 + mg + dd : gq.r0 - 1 - dd; // coupled row
                    if (UPPER ? tr < r0b + m : tr >= r0b)
                      T[(off + a) * TS_BR + tr - r0b] = v;
                    else
                      F[(off + a) * TS_WIN + (UPPER ? tr - (r0b + m) : r0b - 1 - tr)] = v;
                  }
            }
        }
      __syncwarp();
      // one column per lane and pass: columns 0 .. 15 are those of M, 16 .. 16 + TS_WIN - 1 of G.
      // Layout: the TS_NC coefficients of solver lane (a, hh) = 2 a + hh are contiguous (the lane
      // reads them with 128-bit loads at fixed offsets), lanes TS_CSTR bytes apart; jj < TS_WH:
      // G[a][TS_WH hh + jj] (the chain row at distance TS_WH hh + jj), jj >= TS_WH:
      // M[a][8 hh + jj - TS_WH].  Rows the block does not have (padding to 4) are zero.
      const int R2 = 2 * pad4(m);
      double   *C  = reinterpret_cast<double *>(B + TS_OFF_C);
      for (int c = lane; c < TS_BR + TS_WIN; c += 32)
        {
          double y[TS_BR];
          if (UPPER)
            {
#pragma unroll
              for (int a = TS_BR - 1; a >= 0; --a)
                {
                  double v = c < TS_BR ? (c == a ? 1.0 : 0.0) : F[a * TS_WIN + c - TS_BR];
#pragma unroll
                  for (int b = TS_BR - 1; b > a; --b)
                    v -= T[a * TS_BR + b] * y[b];
                  y[a] = v / T[a * TS_BR + a];
                }
            }
          else
            {
#pragma unroll
              for (int a = 0; a < TS_BR; ++a)
                {
                  double v = c < TS_BR ? (c == a ? 1.0 : 0.0) : F[a * TS_WIN + c - TS_BR];
#pragma unroll
                  for (int b = 0; b < a; ++b)
                    v -= T[a * TS_BR + b] * y[b];
                  y[a] = v;
                }
            }
          const int hh = c < TS_BR ? c >> 3 : (c - TS_BR) / TS_WH;
          const int jj = c < TS_BR ? TS_WH + (c & 7) : (c - TS_BR) % TS_WH;
#pragma unroll
          for (int a = 0; a < TS_BR; ++a)
            if (2 * a < R2)
              C[(2 * a + hh) * (TS_CSTR / 8) + jj] = a < m ? y[a] : 0.0;
        }
    }

    // ---- the solve ---------------------------------------------------------------
    // A TEAM of 1 + K warps owns one list of BLOCKS (chains, in the order the host
    // scheduled them); a block is up to 4 consecutive groups of one chain (<= 16 rows):
    //   * K helper warps take the groups round-robin.  A helper streams the items of
    //     its groups (column indices + factor entries) through its own ring, gathers
    //     the solution entries they refer to (re-reading until they are there),
    //     multiplies, folds in the right-hand side, reduces over the warp and posts
    //     the group's totals in the block's mailbox entry.  None of this depends on the
    //     chain, so it runs ahead of it, on several groups at once.
    //   * the solver warp is the chain's recurrence and nothing else, a whole block per
    //     step: out = -(M totals + G w) with w the last 16 rows of the chain before the
    //     block (kept in a tiny shared-memory window) and M, G the block's recurrence
    //     solved ahead of time; publish.  The chain advances 16 rows per step.
    constexpr int TS_MBOX  = 8;  // mailbox entries (blocks) per team
    constexpr int TS_CSLOT  = TS_OFF_COL0 + 4 * TS_CH;                      // 272 (64-entry items)
    constexpr int TS_VSLOT  = 8 * TRSV_G * TS_CH;                           // 2048
    constexpr int TS_HBYTES = TS_NCS * TS_CSLOT + TS_NVS * TS_VSLOT;        // rings of one helper
    constexpr int TS_SSLOT = TS_OFF_C + TS_CSTR * 2 * TS_BR;                // 8720 (48-row window)
    // team area: windows 4 x 1024 (every row twice, TS_HIST apart: a lane reads its part of a
    // window at fixed offsets from one base, no wrap-around) | mailbox 8 x 128 | solved counter
    // 16 | barriers
    constexpr int TS_MENT      = TS_BR; // doubles per mailbox entry
    constexpr int TS_OFF_MBOX  = 2 * 8 * TS_HIST * TS_NWIN;
    constexpr int TS_TEAM_AREA = TS_OFF_MBOX + 8 * TS_MENT * TS_MBOX + 16 + 8 * (11 * (TS_NCS + TS_NVS) + TS_SNSLOT) + 16;
    static_assert(TS_MBOX == 8 && TS_NWIN == 4 && TS_MENT == 16 && TS_TEAM_AREA % 16 == 0, "team area");

    template <bool UPPER, bool TRACE>
    __global__ void __launch_bounds__(GLSNS_TRSV_THREADS, 1)
    trsv_team_kernel(const TrsvWarpDir *__restrict__ dir, const unsigned char *__restrict__ stream,
                     const double *__restrict__ rhs_vec, double *x, int *counters,
                     unsigned long long *trace, const int64_t trace_n, const int K_gate /* helpers per team */)
    {
      extern __shared__ __align__(128) unsigned char smem_all[];
      // K_gate: helpers per team | warp layout << 8 | teams per CTA << 16.  A warp's scheduler is
      // its index modulo 4, and a solver warp that shares a scheduler with the other solver and
      // two polling helpers (layout 0: team t = warps t (K+1) ..., both solvers on scheduler 0 when
      // K = 7) gets a fraction of the issue slots -- its 350 instructions per block took 1400
      // cycles.  Layout 1: the solvers are warps 0 .. teams-1 (one scheduler each), the helpers
      // follow, dealt round-robin to the teams.  Layout 2: as 1, and warps 4 .. 4+teams-1 stay
      // idle, so a solver's scheduler carries one warp less.
      const int K = K_gate & 255, layout = (K_gate >> 8) & 255, n_teams_cta = K_gate >> 16;
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      int       team_in_cta, role;
      if (layout == 0)
        {
          team_in_cta = warp / (K + 1), role = warp - team_in_cta * (K + 1);
        }
      else if (warp < n_teams_cta)
        {
          team_in_cta = warp, role = 0;
        }
      else
        {
          if (layout == 2 && warp >= 4 && warp < 4 + n_teams_cta)
            return;
          const int h = warp - n_teams_cta - (layout == 2 && warp >= 4 ? n_teams_cta : 0);
          team_in_cta = h % n_teams_cta, role = 1 + h / n_teams_cta;
        }
      if (team_in_cta >= n_teams_cta || role > K)
        return;
      const int64_t team      = (int64_t)team_in_cta * gridDim.x + blockIdx.x; // consecutive lists on different SMs
      const size_t  team_smem = (size_t)K * TS_HBYTES + (size_t)TS_SNSLOT * TS_SSLOT + TS_TEAM_AREA;
      unsigned char *T0   = smem_all + (size_t)team_in_cta * team_smem;
      unsigned char *area = T0 + (size_t)K * TS_HBYTES + (size_t)TS_SNSLOT * TS_SSLOT;
      double        *wsm_all = reinterpret_cast<double *>(area);         // [TS_NWIN][2][64] chain windows by row & 63
      double        *mbox = reinterpret_cast<double *>(area + TS_OFF_MBOX); // [TS_MBOX][16], all-ones = empty
      volatile int  *done = reinterpret_cast<volatile int *>(area + TS_OFF_MBOX + 8 * TS_MENT * TS_MBOX); // blocks solved
      unsigned long long *bars_all =
        reinterpret_cast<unsigned long long *>(area + TS_OFF_MBOX + 8 * TS_MENT * TS_MBOX + 16);
      const TrsvWarpDir  *D        = dir + team * (K + 1) + role;
      const int64_t       n_items  = D->n_items;
      unsigned long long  policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      // team-wide initialisation by the solver warp (windows zero, mailbox empty), made
      // visible to the helpers by the one CTA barrier of the kernel
      if (role == 0)
        {
          for (int k = lane; k < 2 * TS_NWIN * TS_HIST; k += 32)
            wsm_all[k] = 0.0;
          // (entries of rows a block does not have hold 0, so that the solver needs no mask)
          const TrsvWarpDir *D0 = dir + ((int64_t)team_in_cta * gridDim.x + blockIdx.x) * (K + 1);
#pragma unroll
          for (int k = 0; k < TS_MBOX * TS_BR / 32; ++k)
            {
              const int e = 2 * k + (lane >> 4), me = (D0->first_m[e >> 2] >> (8 * (e & 3))) & 255;
              reinterpret_cast<unsigned long long *>(mbox)[lane + 32 * k] = (lane & 15) < me ? SENTINEL : 0ull;
            }
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
          // One item per BLOCK (<= 16 rows): out = -(M totals + G w), the recurrence of the
          // whole block solved ahead of time (trsv_pack_values_kernel).  Lane (a, hh) =
          // (lane >> 1, lane & 1) takes row a and, of G, the chain rows at distance TS_WH hh ..
          // TS_WH hh + TS_WH - 1 from the block (window rows r0-1-d / r0+m+d), of M the columns
          // 8 hh .. 8 hh + 7 (totals of the block's own rows).  Coefficients of absent couplings
          // and rows are exactly zero, a window only ever holds finite numbers and the mailbox
          // entries of rows a block does not have hold 0: nothing is masked.
          //
          // This loop is the critical path of a sweep (~1700 of its ~2000 hops are chain steps)
          // and one warp's instruction stream: what counts is the NUMBER of instructions between
          // two publications.  Round 2 started with 485 SASS instructions per block, 32 of them
          // DFMA and 66 loads (the rest: addresses of loads with a run-time stride, modulo
          // arithmetic of the circular window, zeroing of predicated loads, masks).  Now every
          // load is base + immediate: the coefficients lie lane by lane (8 x 128-bit loads of G, 4
          // of M), the window keeps every row twice (no wrap-around), a block's header arrives
          // with the block before it (the window loads go out before the ring is even looked at).
          const unsigned      ring0 = (unsigned)__cvta_generic_to_shared(T0 + (size_t)K * TS_HBYTES);
          unsigned char      *ring  = T0 + (size_t)K * TS_HBYTES;
          unsigned long long *bars  = bars_all + K * (TS_NCS + TS_NVS);
          int64_t             n_iss = 0;
          if (lane == 0)
            {
              for (int s = 0; s < TS_SNSLOT; ++s)
                mbar_init(bars + s, 1);
              asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              for (int s = 0; s < TS_SNSLOT && s < n_items; ++s)
                {
                  const unsigned bytes = 16u * (unsigned)D->first16[s];
                  mbar_expect_tx(bars + s, bytes);
                  bulk_load(ring + (size_t)s * TS_SSLOT, src, bytes, bars + s, policy);
                  src += bytes;
                  ++n_iss;
                }
            }
          __syncwarp();
          int      slot = 0;
          unsigned phase = 0;
          long long tstage[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64(); // debugging aid (trace)
#define TS_TICK(k)                              \
  if (TRACE)                                    \
    {                                           \
      const long long now_ = clock64();         \
      tstage[k] += now_ - tlast;                \
      tlast = now_;                             \
    }
          const int      a = lane >> 1, hh = lane & 1;
          const unsigned wsm0  = (unsigned)__cvta_generic_to_shared(wsm_all);
          const unsigned mbox0 = (unsigned)__cvta_generic_to_shared(mbox);
          int            r0 = D->first_r0, fl = D->first_flags; // header of the block at hand
          for (int64_t g = 0; g < n_items; ++g)
            {
              const int m = fl & 31;
              // the window: what the chain waits for (loads at fixed offsets from one base)
              const unsigned wwin = wsm0 + 16 * TS_HIST * ((fl >> 12) & (TS_NWIN - 1));
              const double *wp = wsm_all + 2 * TS_HIST * ((fl >> 12) & (TS_NWIN - 1)) +
                                 (UPPER ? ((r0 + m) & (TS_HIST - 1)) + TS_WH * hh :
                                          ((r0 - 1) & (TS_HIST - 1)) + TS_HIST - TS_WH * hh);
              double w[TS_WH];
#pragma unroll
              for (int j = 0; j < TS_WH; ++j)
                w[j] = wp[UPPER ? j : -j];
              while (!mbar_try_wait(bars + slot, phase))
                ;
              TS_TICK(0)
              const unsigned S = ring0 + (unsigned)slot * TS_SSLOT;
              int4           h; // r0 and flags of the next block | rows of the block TS_MBOX ahead | next blob
              asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(h.x), "=r"(h.y), "=r"(h.z), "=r"(h.w) : "r"(S));
              const double2 *cl = reinterpret_cast<const double2 *>(ring + (size_t)slot * TS_SSLOT + TS_OFF_C +
                                                                    (size_t)lane * TS_CSTR);
              double         c[TS_NC];
#pragma unroll
              for (int j = 0; j < TS_NC; j += 2)
                {
                  const double2 v = cl[j >> 1];
                  c[j] = v.x, c[j + 1] = v.y;
                }
              double p0 = 0, p1 = 0, p2 = 0, p3 = 0;
#pragma unroll
              for (int j = 0; j < TS_WH; j += 4)
                {
                  p0 += c[j] * w[j];
                  p1 += c[j + 1] * w[j + 1];
                  p2 += c[j + 2] * w[j + 2];
                  p3 += c[j + 3] * w[j + 3];
                }
              TS_TICK(1)
              // totals of everything else (minus the right-hand side), from the helpers: a
              // mailbox entry carries its own readiness (all-ones pattern = empty)
              const unsigned mv = mbox0 + 8 * TS_MENT * (unsigned)(g & (TS_MBOX - 1));
              {
                long long spins = 0;
                for (;;)
                  {
                    unsigned long long b;
                    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(b) : "r"(mv + 8 * (lane & 15)));
                    if (__all_sync(0xffffffffu, b != SENTINEL))
                      break;
                    if ((++spins & 4095) == 0 &&
                        (spins > 64 * SPIN_LIMIT || *(volatile int *)(counters + 1) != 0))
                      {
                        atomicExch(&counters[1], 2);
                        break;
                      }
                  }
              }
              TS_TICK(2)
              {
                double t[8];
                const unsigned tv = mv + 64 * hh;
                asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(t[0]), "=d"(t[1]) : "r"(tv));
                asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2+16];" : "=d"(t[2]), "=d"(t[3]) : "r"(tv));
                asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2+32];" : "=d"(t[4]), "=d"(t[5]) : "r"(tv));
                asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2+48];" : "=d"(t[6]), "=d"(t[7]) : "r"(tv));
#pragma unroll
                for (int j = 0; j < 8; j += 4)
                  {
                    p0 += c[TS_WH + j] * t[j];
                    p1 += c[TS_WH + j + 1] * t[j + 1];
                    p2 += c[TS_WH + j + 2] * t[j + 2];
                    p3 += c[TS_WH + j + 3] * t[j + 3];
                  }
              }
              double p = (p0 + p1) + (p2 + p3);
              p += __shfl_xor_sync(0xffffffffu, p, 1);
              if (hh == 0 && a < m)
                {
                  const double v = -p;
                  st_result(x + r0 + a, v);
                  if (!(fl & IT_GUEST))
                    {
                      const unsigned wa = wwin + 8 * ((r0 + a) & (TS_HIST - 1));
                      asm volatile("st.shared.f64 [%0], %1;\n\tst.shared.f64 [%0+512], %1;" ::"r"(wa), "d"(v) : "memory");
                      static_assert(8 * TS_HIST == 512, "second copy of a window row");
                    }
                  if (TRACE) // debugging aid (glsns_ilu_apply_trace): when was the row published
                    {
                      unsigned long long tns;
                      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
                      trace[r0 + a] = tns;
                    }
                }
              __syncwarp(); // the window rows are written, the mailbox entry is read
              // hand the entry back: empty it for the block TS_MBOX ahead (0 in the rows that
              // block does not have), then let that block in
              if (lane < TS_BR)
                asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(mv + 8 * lane), "l"(lane < h.z ? SENTINEL : 0ull) : "memory");
              if (lane == 0)
                *done = (int)g + 1;
              TS_TICK(3)
              // the slot is used up: refill it with the block TS_SNSLOT ahead
              if (lane == 0 && n_iss < n_items)
                {
                  const unsigned bytes = 16u * (unsigned)h.w;
                  mbar_expect_tx(bars + slot, bytes);
                  bulk_load(ring + (size_t)slot * TS_SSLOT, src, bytes, bars + slot, policy);
                  src += bytes;
                  ++n_iss;
                }
              r0   = h.x;
              fl   = h.y;
              slot = slot + 1 == TS_SNSLOT ? 0 : slot + 1;
              phase ^= slot == 0;
              TS_TICK(4)
            }
          if (TRACE && lane == 0 && (team + 1) * 8 <= trace_n)
            { // cycles in: ring wait (with the window loads), product, mailbox wait, totals+publish, release+refill
              for (int k = 0; k < 6; ++k)
                trace[2 * trace_n + team * 8 + k] = (unsigned long long)tstage[k];
              trace[2 * trace_n + team * 8 + 6] = (unsigned long long)n_items;
            }
#undef TS_TICK
          return;
        }

      // =============================== helper ===============================
      // A helper's rate is set by the latency of its gathers: an item (<= 64 entries of a group)
      // is requested in stage G and consumed TS_D steps later in stage R, so a helper has TS_D
      // items' worth of solution entries in flight and finishes one item every latency / TS_D.
      // With the whole blob (indices + values) in one ring the depth was 2 -- 0.5 us per item,
      // 0.7 us per block with 7 helpers, which is what a chain step took (trace, p10) however
      // lean the solver warp was.  So the column indices (272 B) and the factor entries (2 KB)
      // travel in rings of their own: the indices are needed in G only (their slot is refilled
      // at once, 8 slots), the values in R only (4 slots), and the depth is a matter of
      // registers (one set per item in flight), not of shared memory.
      const int           hid   = role - 1;
      unsigned char      *cring = T0 + (size_t)hid * TS_HBYTES, *vring = cring + TS_NCS * TS_CSLOT;
      unsigned long long *cbar  = bars_all + hid * (TS_NCS + TS_NVS), *vbar = cbar + TS_NCS;
      const unsigned char *srcv = stream + D->offset_v;
      int64_t             n_issc = 0, n_issv = 0;
      if (lane == 0)
        {
          for (int s = 0; s < TS_NCS + TS_NVS; ++s)
            mbar_init(cbar + s, 1);
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          for (int s = 0; s < TS_NCS && s < n_items; ++s)
            {
              const unsigned bytes = 16u * (unsigned)D->first16[s];
              mbar_expect_tx(cbar + s, bytes);
              bulk_load(cring + (size_t)s * TS_CSLOT, src, bytes, cbar + s, policy);
              src += bytes;
              ++n_issc;
            }
          for (int s = 0; s < TS_NVS && s < n_items; ++s)
            {
              const unsigned bytes = 16u * (unsigned)D->firstv16[s];
              mbar_expect_tx(vbar + s, bytes);
              bulk_load(vring + (size_t)s * TS_VSLOT, srcv, bytes, vbar + s, policy);
              srcv += bytes;
              ++n_issv;
            }
        }
      __syncwarp();
      // pipeline registers: TS_D + 1 rotating sets (no copies: a copy would wait for the loads
      // in flight), selected at compile time by the unrolled loop
      constexpr int      NS = TS_D + 1;
      int32_t            cS[NS][TS_U];
      unsigned long long bS[NS][TS_U];
      unsigned           pS[NS];
      int                fS[NS], zS[NS], wS[NS]; // header of the item: flags, where its totals go, sizes of the blobs ahead
      int                xS[NS];                 // its first row (trace only)
      double             rS[NS];                 // right-hand side of row r0 + lane
#pragma unroll
      for (int k = 0; k < NS; ++k)
        pS[k] = 0, fS[k] = 0, zS[k] = 0, wS[k] = 0, xS[k] = 0, rS[k] = 0;
      double acc[TRSV_G];
#pragma unroll
      for (int a = 0; a < TRSV_G; ++a)
        acc[a] = 0;
      int      slotG = 0, slotR = 0;
      unsigned phaseG = 0, phaseR = 0;
      unsigned long long n_rounds = 0, n_polled = 0, n_waited = 0; // debugging aid (trace)
      long long          hst[4] = {0, 0, 0, 0}; // trace: cycles waiting for indices, values, solution entries, a free mailbox entry
      const long long    h_begin = TRACE ? clock64() : 0;
#define HS_WAIT(k, stmt)                 \
  {                                      \
    long long c0_ = TRACE ? clock64() : 0; \
    stmt;                                \
    if (TRACE)                           \
      hst[k] += clock64() - c0_;         \
  }
      // one pipeline step: G(it) into set KG, R(it - TS_D) from set (KG + 1) % NS
      auto step = [&](auto KG, const int64_t it) {
        constexpr int      kg = decltype(KG)::value, kr = (kg + 1) % NS, k1 = (kg + 2) % NS, k2 = (kg + 3) % NS;
        int32_t(&cN)[TS_U]            = cS[kg];
        unsigned long long(&bN)[TS_U] = bS[kg];
        unsigned &pendN               = pS[kg];
        int32_t(&cG)[TS_U]            = cS[kr];
        unsigned long long(&bG)[TS_U] = bS[kr];
        unsigned &pendG               = pS[kr];
        // ---- G(it): wait for the indices, request the solution entries they refer to ----
        pendN = 0;
        if (it < n_items)
          {
            HS_WAIT(0, while (!mbar_try_wait(cbar + slotG, phaseG));)
            unsigned char *S     = cring + (size_t)slotG * TS_CSLOT;
            const int4     h     = *reinterpret_cast<const int4 *>(S);
            const int      flags = h.y, cntc = flags >> 16;
            const int32_t *scol  = reinterpret_cast<const int32_t *>(S + TS_OFF_COL0);
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
            rS[kg] = 0.0; // (lane 8 a ends up with the total of row a: it takes that row's right-hand side)
            if ((flags & IT_LAST) && (lane & 7) == 0 && (lane >> 3) < (flags & 31))
              rS[kg] = rhs_vec[h.x + (lane >> 3)];
            fS[kg] = flags, zS[kg] = h.z, wS[kg] = h.w;
            if (TRACE)
              xS[kg] = h.x;
            __syncwarp(); // every lane has its indices: the slot takes the indices TS_NCS items ahead
            if (lane == 0 && n_issc < n_items)
              {
                const unsigned bytes = 16u * (unsigned)(h.w & 255);
                mbar_expect_tx(cbar + slotG, bytes);
                bulk_load(S, src, bytes, cbar + slotG, policy);
                src += bytes;
                ++n_issc;
              }
            slotG = slotG + 1 == TS_NCS ? 0 : slotG + 1;
            phaseG ^= slotG == 0;
          }
        // ---- R(it - TS_D): entries not there yet are re-read until they are; multiply ----
        if (it >= TS_D)
          {
            const int    flags = fS[kr], m = flags & 31, cp = pad4(flags >> 16);
            unsigned char *SV  = vring + (size_t)slotR * TS_VSLOT;
            const double *sval = reinterpret_cast<const double *>(SV);
            long long            spins = 0;
            unsigned long long   t_rs = 0, t_det = 0; // debugging aid (trace)
            if (TRACE)
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_rs));
            HS_WAIT(1, while (!mbar_try_wait(vbar + slotR, phaseR));)
            const long long cp0 = TRACE ? clock64() : 0;
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
                if (TRACE)
                  {
                    n_waited += spins == 0;
                    ++n_rounds;
                    n_polled += __popc(pendG);
                  }
#pragma unroll
                for (int u = 0; u < TS_U; ++u)
                  if (pendG & (1u << u))
                    bG[u] = ld_relaxed_u64(x + cG[u]);
                // the entries of the two items behind this one that were not there when they
                // were first read are refreshed in the same round trip: when this item is
                // done, they need no round trip of their own
#pragma unroll
                for (int u = 0; u < TS_U; ++u)
                  {
                    if ((pS[k1] & (1u << u)) && bS[k1][u] == SENTINEL)
                      bS[k1][u] = ld_relaxed_u64(x + cS[k1][u]);
                    if (NS > 3 && (pS[k2] & (1u << u)) && bS[k2][u] == SENTINEL)
                      bS[k2][u] = ld_relaxed_u64(x + cS[k2][u]);
                  }
                // bug guard: give up after seconds of waiting, or as soon as another
                // warp has given up (the host reports GLSNS_ERR_CUDA)
                if ((++spins & 1023) == 0 &&
                    (spins > SPIN_LIMIT || *(volatile int *)(counters + 1) != 0))
                  {
                    atomicExch(&counters[1], 2);
                    break;
                  }
              }
            if (TRACE)
              {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_det));
                if (spins)
                  hst[2] += clock64() - cp0;
              }
            if (flags & IT_LAST)
              {
                // the group is complete: total over the warp, fold in the right-hand side, post
                // in the mailbox once the solver has freed the entry.  The four rows are
                // transposed through shared memory (the item's value slot: its entries are
                // used up) so that eight lanes share a row: lane (a, s) = (lane >> 3, lane & 7)
                // adds four partial sums of row a, three shuffles do the rest -- a third of
                // the instructions of reducing four values over 32 lanes with shuffles alone;
                // lane 8 a ends up with the total of row a
                double tot;
                {
                  double *scr = reinterpret_cast<double *>(SV);
                  __syncwarp();
#pragma unroll
                  for (int a = 0; a < TRSV_G; ++a)
                    scr[a * 32 + lane] = acc[a];
                  __syncwarp();
                  const double2 v0 = reinterpret_cast<const double2 *>(scr)[2 * lane],
                                v1 = reinterpret_cast<const double2 *>(scr)[2 * lane + 1];
                  tot = (v0.x + v0.y) + (v1.x + v1.y);
                  tot += __shfl_xor_sync(0xffffffffu, tot, 4);
                  tot += __shfl_xor_sync(0xffffffffu, tot, 2);
                  tot += __shfl_xor_sync(0xffffffffu, tot, 1);
                  tot -= rS[kr];
                }
                // (header: position of the group's block in the team's list << 4 | row offset)
                const int bseq = zS[kr] >> 4, boff = zS[kr] & 15;
                volatile unsigned long long *mv = reinterpret_cast<volatile unsigned long long *>(
                  mbox + (bseq & (TS_MBOX - 1)) * TS_MENT + boff);
                spins = 0;
                // the entry is this block's once the solver is within TS_MBOX blocks of it
                HS_WAIT(3, while (*done + TS_MBOX <= bseq) if ((++spins & 4095) == 0 &&
                                                               (spins > 64 * SPIN_LIMIT || *(volatile int *)(counters + 1) != 0)) {
                  atomicExch(&counters[1], 2);
                  break;
                })
                if ((lane & 7) == 0 && (lane >> 3) < m)
                  {
                    unsigned long long b = (unsigned long long)__double_as_longlong(tot);
                    if (b == SENTINEL)
                      b = 0x7FF8000000000000ull;
                    mv[lane >> 3] = b;
                    if (TRACE) // debugging aid: when were the row's totals posted
                      {
                        unsigned long long tns;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
                        trace[4 * trace_n + xS[kr] + (lane >> 3)] = tns;
                        trace[6 * trace_n + xS[kr] + (lane >> 3)] = t_rs;  // began the group's last item
                        trace[8 * trace_n + xS[kr] + (lane >> 3)] = t_det; // had all its inputs
                      }
                  }
#pragma unroll
                for (int a = 0; a < TRSV_G; ++a)
                  acc[a] = 0;
              }
            __syncwarp(); // every lane is done with the values before the slot is refilled
            if (lane == 0 && n_issv < n_items)
              {
                const unsigned bytes = 16u * (unsigned)((wS[kr] >> 8) & 255); // the values TS_NVS items ahead
                if (flags & IT_LAST) // (the slot served as scratch for the reduction: generic writes before the bulk copy)
                  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(vbar + slotR, bytes);
                bulk_load(SV, srcv, bytes, vbar + slotR, policy);
                srcv += bytes;
                ++n_issv;
              }
            slotR = slotR + 1 == TS_NVS ? 0 : slotR + 1;
            phaseR ^= slotR == 0;
          }
      };
      for (int64_t it = 0; it < n_items + TS_D; it += NS)
        {
          step(std::integral_constant<int, 0>(), it);
          if (it + 1 < n_items + TS_D)
            step(std::integral_constant<int, 1>(), it + 1);
          if (it + 2 < n_items + TS_D)
            step(std::integral_constant<int, 2>(), it + 2);
          if constexpr (NS > 3)
            if (it + 3 < n_items + TS_D)
              step(std::integral_constant<int, 3>(), it + 3);
          if constexpr (NS > 4)
            if (it + 4 < n_items + TS_D)
              step(std::integral_constant<int, 4>(), it + 4);
        }
      if (TRACE && 20 * (int64_t)gridDim.x * n_teams_cta + (team + 1) * 8 <= trace_n)
        { // per helper lane: items that had to wait, polling rounds, entries re-read
          for (int o = 16; o > 0; o >>= 1)
            n_polled += __shfl_xor_sync(0xffffffffu, n_polled, o);
          if (lane == 0)
            {
              unsigned long long *hc = trace + 2 * trace_n + 8 * (int64_t)gridDim.x * n_teams_cta + team * 4;
              atomicAdd(hc + 0, (unsigned long long)n_items);
              atomicAdd(hc + 1, n_waited);
              atomicAdd(hc + 2, n_rounds);
              atomicAdd(hc + 3, n_polled);
              // cycles: waiting for indices, values, solution entries, a mailbox entry; all
              unsigned long long *hs = trace + 2 * trace_n + 12 * (int64_t)gridDim.x * n_teams_cta + team * 8;
              for (int k = 0; k < 4; ++k)
                atomicAdd(hs + k, (unsigned long long)hst[k]);
              atomicAdd(hs + 4, (unsigned long long)(clock64() - h_begin));
            }
#undef HS_WAIT
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
      // per SM.  Helpers are the servers of a team's queue of groups: tools/trsv_trace.py shows
      // the critical group waiting for its helper to get to it, not for its inputs, so many
      // helpers per solver (and small items, which make their rings small) win: 3 x (1+3)
      // with 128-entry items 10.8 ms per application at 64^3 cells, 2 x (1+5) 8.75 ms,
      // 2 x (1+6) with 96-entry items 8.51 ms, 2 x (1+7) with 64-entry items 8.33 ms.
      int teams = 2, helpers = 7;
      int layout = 1; // which warps are the solvers (see trsv_team_kernel)
      int
      warps() const
      {
        return teams * (helpers + 1) + (layout == 2 ? teams : 0);
      }
    };

    size_t
    team_smem_bytes(int helpers)
    {
      return (size_t)helpers * TS_HBYTES + (size_t)TS_SNSLOT * TS_SSLOT + TS_TEAM_AREA;
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
        if (getenv("GLSNS_TRSV_LAYOUT"))
          t.layout = std::max(0, std::min(2, atoi(getenv("GLSNS_TRSV_LAYOUT"))));
        t.helpers = std::max(1, std::min(11, t.helpers));
        t.teams   = std::max(1, std::min(TS_MAXWARPS / (t.helpers + 1), t.teams));
        if (t.warps() > TS_MAXWARPS || (t.layout == 2 && t.teams > 4))
          t.layout = 1;
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
      auto             kern = trace ? trsv_team_kernel<UPPER, true> : trsv_team_kernel<UPPER, false>;
      GLSNS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
      int  *counters = ctx->counters.p;
      int   K        = cfg.helpers | (cfg.layout << 8) | (cfg.teams << 16);
      void *args[]   = {(void *)&dir,      (void *)&stream, (void *)&rhs,          (void *)&x,
                        (void *)&counters, (void *)&trace,  (void *)&ctx->n_owned, (void *)&K};
      GLSNS_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kern, dim3(ctx->trsv_grid),
                                                  dim3(cfg.warps() * 32), args,
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
    { // per row: the first row of its group (the factorisation takes its pivot rows group-wise)
      std::vector<int32_t> gf((size_t)std::max<int64_t>(n, 1), -1);
      for (int64_t i = 0; i < n; ++i)
        gf[i] = grp_of[i] >= 0 ? grp_ptr[grp_of[i]] : -1;
      GLSNS_TRY(dev_upload(ctx, ctx->grp_first, gf.data(), gf.size()));
      GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }

    const TrsvConfig cfg = trsv_config();
    ctx->trsv_grid       = ctx->n_sm;
    const int64_t NW     = (int64_t)ctx->trsv_grid * cfg.teams; // teams: one chain list each
    const int32_t max_level_gap = getenv("GLSNS_TRSV_GAP") ? atoi(getenv("GLSNS_TRSV_GAP")) : 1;
    const int32_t max_block     = std::max(1, std::min(TS_BG, getenv("GLSNS_TRSV_BLOCK") ? atoi(getenv("GLSNS_TRSV_BLOCK")) : TS_BG));
    const int     K      = cfg.helpers;

    std::vector<int32_t> glev(ng), cnt(ng), e_off(ng), order(ng), blk_of(ng);
    std::vector<uint64_t> fmask(ng);
    std::vector<uint8_t> link(ng);

    // One sweep: levels, blocks, chain links, level-ordered schedule on NW teams, item lists.
    // `when` (may be null): for every row, when a traced application published it (ns)
    auto schedule = [&](const bool upper, TrsvSweep &sw, int32_t &n_levels,
                        const unsigned long long *when) -> glsns_status {
      // entries of group g this sweep reads, [kb, ke) in CSR offsets of its first row
      auto range = [&](int64_t g, int64_t &kb, int64_t &ke) {
        const int64_t i = grp_ptr[g];
        kb              = upper ? diag[i] + grp_m[g] : rowptr[i];
        ke              = upper ? rowptr[i + 1] : diag[i];
        if (upper)
          while (ke > kb && col[ke - 1] >= n)
            --ke; // ghost columns: outside the diagonal block
      };
      // ---- group levels (the depth of the ordering's dependency DAG) and links ----
      int32_t nlev = 0;
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
          // link: the group depends on its neighbour in the numbering, with nothing but
          // diagonal-only rows (and fewer than a window of them) in between
          const int64_t p = upper ? g + 1 : g - 1;
          link[g]         = 0;
          if (p >= 0 && p < ng)
            {
              const int64_t gap = upper ? grp_ptr[p] - (grp_ptr[g] + grp_m[g]) :
                                          grp_ptr[g] - (grp_ptr[p] + grp_m[p]);
              if (gap < TS_BR - 1)
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
              // bit 1: the neighbour's rows are adjacent and it was solved just before this
              // group can be (no other input of the group is much younger): the two are
              // worth solving as one block
              if (link[g] && gap == 0 && l - glev[p] <= max_level_gap)
                link[g] |= 2;
            }
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
      if (!upper && !when)
        { // the factorisation takes the groups in the same order (sparse.cu)
          std::vector<int2> fg((size_t)ng);
          for (int64_t t = 0; t < ng; ++t)
            fg[(size_t)t] = make_int2(grp_ptr[order[t]], grp_m[order[t]]);
          GLSNS_TRY(dev_upload(ctx, ctx->fgroups, fg.data(), fg.size()));
          GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
      // ---- blocks: runs of <= TS_BG linked groups with adjacent rows, in sweep order ----
      // Everything a block needs from outside was numbered before its first group (in sweep
      // order), so nothing outside it can wait for part of it: solving its rows together
      // (one step of the solver warp) cannot dead-lock, and it removes the hops between them
      // from the critical path.
      std::vector<int32_t> b_first, b_ng, b_r0, b_m;
      b_first.reserve(ng / 2 + 1), b_ng.reserve(ng / 2 + 1), b_r0.reserve(ng / 2 + 1), b_m.reserve(ng / 2 + 1);
      for (int64_t gi = 0; gi < ng; ++gi)
        {
          const int64_t g = upper ? ng - 1 - gi : gi;
          if (gi > 0 && (link[g] & 2) && b_ng.back() < max_block && b_m.back() + grp_m[g] <= TS_BR)
            {
              ++b_ng.back();
              b_m.back() += grp_m[g];
              if (upper)
                b_r0.back() = grp_ptr[g];
            }
          else
            {
              b_first.push_back((int32_t)g), b_ng.push_back(1), b_r0.push_back(grp_ptr[g]),
                b_m.push_back(grp_m[g]);
            }
          blk_of[g] = (int32_t)b_first.size() - 1;
        }
      const int64_t nb = (int64_t)b_first.size();
      auto          group_of_block = [&](int64_t b, int32_t q) -> int64_t { // q-th group in sweep order
        return upper ? (int64_t)b_first[b] - q : (int64_t)b_first[b] + q;
      };
      // block levels and block links (block ids run in sweep order: the predecessor of the
      // first group of block b is the last group of block b - 1)
      std::vector<int32_t> blev(nb), border(nb), team_of(nb), slot_of(nb), bseq(nb);
      std::vector<int64_t> best(nb);
      std::vector<uint8_t> blink(nb), has_succ(nb, 0), is_primary(nb);
      int32_t              nblev = 0;
      // `best`, the key the lists are ordered by.  With a trace (`when`): the moment the block's
      // last input was published in that run, i.e. when the block could have been solved -- a
      // team that orders its list by level makes a chain that is ready wait behind one that
      // is late, and how late is only known from a run.  (A block is published strictly after
      // its inputs, so this key grows along every dependency, as the dead-lock argument needs.)
      // Without a trace: when a block can be solved at the earliest, in units of one chain step, if a
      // value that travels to another team (L2, a helper's gather and reduction, the mailbox)
      // takes cost_cross of them; the lists are ordered by it.  With both costs 1 (the
      // default) it is the block level.  Weighted orders (cross = 3: where the DAG is mostly
      // chains the front advances several levels in the time a level takes elsewhere) were
      // measured: -4 % at 32^3 cells, +5 % at 64^3.
      const int64_t cost_chain = getenv("GLSNS_TRSV_COST_CHAIN") ? atoi(getenv("GLSNS_TRSV_COST_CHAIN")) : 1;
      // Second session of round 2 (lean solver warp): cross = 3 is 9 % faster at 32^3 cells (2.17 ->
      // 1.97 ms), 5 % slower at 64^3 (7.37 -> 7.72).  The difference is the width of the DAG against
      // the number of teams: where a level has fewer groups than half the teams (the 32^3 mesh,
      // the per-rank blocks of a strong-scaled run) the sweep is a few long chains and the order
      // should follow the time a value needs to cross to another team; where every team has several
      // groups per level the level order keeps the lists balanced.  A function of the sparsity
      // pattern alone, like the rest of the schedule.
      const int64_t cost_cross = getenv("GLSNS_TRSV_COST_CROSS") ? atoi(getenv("GLSNS_TRSV_COST_CROSS")) :
                                 (ng < (int64_t)(nlev + 1) * (NW / 2) ? 3 : 1);
      for (int64_t b = 0; b < nb; ++b)
        {
          int32_t l = 0;
          int64_t e = 0;
          for (int32_t q = 0; q < b_ng[b]; ++q)
            {
              int64_t kb, ke;
              range(group_of_block(b, q), kb, ke);
              for (int64_t k = kb; k < ke; ++k)
                {
                  const int32_t dg = grp_of[col[k]];
                  if (dg >= 0 && blk_of[dg] != b)
                    {
                      const int32_t db = blk_of[dg];
                      l                = std::max(l, blev[db] + 1);
                      e = when ? std::max<int64_t>(e, (int64_t)when[col[k]]) :
                                 std::max(e, best[db] + (db == b - 1 && (link[b_first[b]] & 1) ? cost_chain : cost_cross));
                    }
                }
            }
          blev[b]  = l;
          best[b]  = e;
          nblev    = std::max(nblev, l);
          blink[b] = b > 0 && (link[b_first[b]] & 1) && l - blev[b - 1] <= max_level_gap;
          if (blink[b])
            has_succ[b - 1] = 1;
        }
      // (a key that grows along every dependency keeps the lists dead-lock free: the blocked
      // block with the smallest key waits on a block with a smaller one, which is some team's
      // current or earlier item)
      for (int64_t b = 0; b < nb; ++b)
        border[b] = (int32_t)b;
      std::stable_sort(border.begin(), border.end(),
                       [&](int32_t x, int32_t y) { return best[x] < best[y]; });
      // ---- list scheduling: a chained block follows its predecessor on the same team ----
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
      for (int64_t t = 0; t < nb; ++t)
        {
          const int32_t b = border[t];
          // A chain runs through one window of its team.  When every window of every
          // team is taken, a block is placed on a busy team as a GUEST: solved in list
          // order without touching any window, so the chains there keep their forwarding.
          int32_t ws; // team * TS_NWIN + slot
          bool    chained = false, primary = true;
          if (blink[b] && is_primary[b - 1] && last_of[slot_of[b - 1]] == b - 1)
            {
              ws      = slot_of[b - 1];
              chained = true;
            }
          else if ((ws = take_slot()) >= 0)
            n_interrupted += blink[b]; // chain head, or the successor of a guest: a fresh window
          else
            {
              ws      = (blink[b] ? team_of[b - 1] : (int32_t)(rr++ % NW)) * TS_NWIN;
              primary = false;
              ++n_steals;
            }
          n_heads += !blink[b];
          is_primary[b] = primary;
          team_of[b]    = ws / TS_NWIN;
          slot_of[b]    = ws;
          if (!chained && primary) // a new chain starts here: first row (lower) / end row (upper)
            chain_edge[ws] = upper ? b_r0[b] + b_m[b] : b_r0[b];
          // Entries that couple a group to the rows of its own block and, on a chain, to the
          // chain rows still in the window: the run next to the in-group block, at most
          // TS_WIN rows away.  They are the solver's (T and F of the block); the helpers
          // take the rest.
          const int32_t edge = chained ? chain_edge[ws] : (upper ? b_r0[b] + b_m[b] : b_r0[b]);
          for (int32_t q = 0; q < b_ng[b]; ++q)
            {
              const int64_t g  = group_of_block(b, q);
              const int32_t r0 = grp_ptr[g], m = grp_m[g];
              int64_t       kb, ke;
              range(g, kb, ke);
              uint64_t fm = 0;
              int32_t  nf = 0;
              if (!upper)
                for (int64_t k = ke - 1; k >= kb; --k)
                  {
                    const int32_t d = r0 - 1 - col[k];
                    if (d >= TS_WIN || col[k] < edge || grp_of[col[k]] < 0)
                      break;
                    fm |= 1ull << d;
                    ++nf;
                  }
              else
                for (int64_t k = kb; k < ke; ++k)
                  {
                    const int32_t d = col[k] - (r0 + m);
                    if (d >= TS_WIN || col[k] >= edge || grp_of[col[k]] < 0)
                      break;
                    fm |= 1ull << d;
                    ++nf;
                  }
              fmask[g] = fm;
              e_off[g] = (int32_t)((upper ? kb + nf : kb) - rowptr[grp_ptr[g]]);
              cnt[g]   = (int32_t)(ke - kb - nf);
            }
          if (primary)
            {
              if (has_succ[b])
                last_of[ws] = b; // (last block of the chain in this window so far)
              else
                { // the chain ends here: the window is free again
                  last_of[ws] = -1;
                  --n_live[team_of[b]];
                  bucket[n_live[team_of[b]]].push_back(team_of[b]);
                }
            }
        }
      if (getenv("GLSNS_TRSV_DEBUG"))
        fprintf(stderr,
                "trsv_analyse %s: %lld groups, %d levels; %lld blocks, %d block levels, %lld teams, "
                "%lld chain heads, %lld taken as guests on busy teams, %lld chains restarted\n",
                upper ? "upper" : "lower", (long long)ng, nlev + 1, (long long)nb, nblev + 1,
                (long long)NW, (long long)n_heads, (long long)n_steals, (long long)n_interrupted);
      // ---- item lists: per team one solver list (one item per block) and K helper lists
      // (the team's groups round-robin, <= TS_CH entries per item) ----
      // ---- item lists: per team one solver list (one item per block) and K helper lists
      // (the team's groups round-robin, <= TS_CH entries per item) ----
      const int64_t        NWARP = NW * (K + 1);
      std::vector<int64_t> n_it(NWARP + 1, 0), team_blocks(NW, 0), team_groups(NW, 0);
      std::vector<int32_t> helper_of(ng);
      for (int64_t t = 0; t < nb; ++t)
        {
          const int32_t b  = border[t];
          const int64_t tm = team_of[b];
          bseq[b]          = (int32_t)team_blocks[tm]++;
          n_it[tm * (K + 1) + 0 + 1] += 1;
          for (int32_t q = 0; q < b_ng[b]; ++q)
            {
              const int64_t g = group_of_block(b, q);
              helper_of[g]    = (int32_t)(team_groups[tm]++ % K);
              n_it[tm * (K + 1) + 1 + helper_of[g] + 1] += std::max<int64_t>(1, (cnt[g] + TS_CH - 1) / TS_CH);
            }
        }
      for (int64_t w = 0; w < NWARP; ++w)
        n_it[w + 1] += n_it[w];
      std::vector<TrsvItem> items((size_t)n_it[NWARP]), gdesc((size_t)ng);
      std::vector<int64_t>  fill(n_it.begin(), n_it.end() - 1);
      int64_t               n_gdesc = 0;
      for (int64_t t = 0; t < nb; ++t)
        {
          const int32_t b  = border[t];
          const int64_t tm = team_of[b];
          {
            TrsvItem &it = items[(size_t)fill[tm * (K + 1)]++];
            it.rs0 = n_gdesc, it.r0 = b_r0[b], it.len = b_ng[b], it.e_off = 0, it.nlow = 0, it.fmask = 0;
            it.flags = b_m[b] | IT_SOLVER | (is_primary[b] ? 0 : IT_GUEST) | ((slot_of[b] % TS_NWIN) << 12);
          }
          for (int32_t q = 0; q < b_ng[b]; ++q)
            {
              const int64_t g = group_of_block(b, q);
              const int64_t i = grp_ptr[g];
              const int32_t m = grp_m[g], len = (int32_t)(rowptr[i + 1] - rowptr[i]);
              {
                TrsvItem &gd = gdesc[(size_t)n_gdesc++];
                gd.rs0 = rowptr[i], gd.r0 = (int32_t)i, gd.len = len, gd.e_off = 0, gd.flags = m;
                gd.nlow = (int32_t)(diag[i] - rowptr[i]);
                gd.fmask = (int32_t)(uint32_t)fmask[g], gd.fmask2 = (int32_t)(uint32_t)(fmask[g] >> 32);
              }
              const int32_t nchunk = std::max(1, (cnt[g] + TS_CH - 1) / TS_CH);
              for (int32_t c = 0; c < nchunk; ++c)
                {
                  TrsvItem  &it   = items[(size_t)fill[tm * (K + 1) + 1 + helper_of[g]]++];
                  // (the entries solved last come last: the upper sweep takes the chunks of a
                  // row from the far end, so that a helper waits on the group's final item only)
                  const int  ch   = upper ? nchunk - 1 - c : c;
                  const int  cntc = std::max(0, std::min(TS_CH, cnt[g] - ch * TS_CH));
                  const bool last = c == nchunk - 1;
                  it.rs0   = rowptr[i];
                  it.r0    = (int32_t)i;
                  it.len   = len;
                  it.e_off = e_off[g] + ch * TS_CH;
                  it.flags = m | (last ? IT_LAST : 0) | (cntc << 16);
                  it.nlow  = (int32_t)(diag[i] - rowptr[i]);
                  it.fmask = (bseq[b] << 4) | (int32_t)(i - b_r0[b]); // where its totals go
                }
            }
        }
      std::vector<int32_t> &row_warp = upper ? ctx->trsv_row_warp_u : ctx->trsv_row_warp_l;
      row_warp.assign((size_t)n, -1);
      for (int64_t g = 0; g < ng; ++g)
        for (int32_t a = 0; a < grp_m[g]; ++a)
          row_warp[grp_ptr[g] + a] = team_of[blk_of[g]] | (fmask[g] ? 1 << 30 : 0);
      // stream layout: the blobs of one warp back to back, warps one after another
      const int64_t        nit = n_it[NWARP];
      std::vector<int64_t> blob_off((size_t)nit), blob_off_v((size_t)nit, 0);
      std::vector<int32_t> next16((size_t)nit, 0);
      std::vector<TrsvWarpDir> dirv((size_t)NWARP);
      int64_t              off = 0;
      for (int64_t w = 0; w < NWARP; ++w)
        {
          TrsvWarpDir &D = dirv[w];
          memset(&D, 0, sizeof(D));
          D.offset  = off;
          D.n_items = (int32_t)(n_it[w + 1] - n_it[w]);
          if (w % (K + 1) == 0)
            { // a solver's list: every block's header travels with the block before it
              D.first_m[0] = D.first_m[1] = 0x10101010;
              for (int64_t k = n_it[w]; k < n_it[w + 1]; ++k)
                {
                  TrsvItem &it = items[(size_t)k];
                  const int64_t j = k - n_it[w];
                  if (j < TS_MBOX)
                    D.first_m[j >> 2] = (D.first_m[j >> 2] & ~(255 << (8 * (j & 3)))) | ((it.flags & 31) << (8 * (j & 3)));
                  if (j == 0)
                    D.first_r0 = it.r0, D.first_flags = it.flags;
                  it.e_off = k + 1 < n_it[w + 1] ? items[(size_t)k + 1].r0 : 0;
                  it.nlow  = k + 1 < n_it[w + 1] ? items[(size_t)k + 1].flags : 0;
                  it.pad_  = k + TS_MBOX < n_it[w + 1] ? (items[(size_t)k + TS_MBOX].flags & 31) : TS_BR;
                }
            }
          const bool solver = w % (K + 1) == 0; // (a team's first list is its solver's)
          for (int64_t k = n_it[w]; k < n_it[w + 1]; ++k)
            {
              const int32_t b16 = blob_bytes(items[(size_t)k].flags) / 16;
              blob_off[(size_t)k] = off;
              off += 16 * (int64_t)b16;
              const int64_t j  = k - n_it[w];
              const int     ns = solver ? TS_SNSLOT : TS_NCS;
              if (j < ns)
                D.first16[j] = b16;
              else
                next16[(size_t)(k - ns)] |= b16;
            }
          if (!solver)
            { // a helper's factor entries follow its indices, in a stream of their own
              D.offset_v = off;
              for (int64_t k = n_it[w]; k < n_it[w + 1]; ++k)
                {
                  const int32_t v16 = value_blob_bytes(items[(size_t)k].flags) / 16;
                  blob_off_v[(size_t)k] = off;
                  off += 16 * (int64_t)v16;
                  const int64_t j = k - n_it[w];
                  if (j < TS_NVS)
                    D.firstv16[j] = v16;
                  else
                    next16[(size_t)(k - TS_NVS)] |= v16 << 8;
                }
            }
        }
      sw.n_items      = nit;
      sw.stream_bytes = off;
      GLSNS_TRY(dev_upload(ctx, sw.items, items.data(), items.size()));
      GLSNS_TRY(dev_upload(ctx, sw.gdesc, gdesc.data(), gdesc.size()));
      GLSNS_TRY(dev_upload(ctx, sw.blob_off, blob_off.data(), blob_off.size()));
      GLSNS_TRY(dev_upload(ctx, sw.blob_off_v, blob_off_v.data(), blob_off_v.size()));
      GLSNS_TRY(dev_upload(ctx, sw.next16, next16.data(), next16.size()));
      GLSNS_TRY(dev_upload(ctx, sw.dir, reinterpret_cast<const unsigned char *>(dirv.data()),
                           dirv.size() * sizeof(TrsvWarpDir)));
      GLSNS_TRY(dev_alloc(ctx, sw.stream, (size_t)std::max<int64_t>(off, 16)));
      // (all-zero factor entries until the first factorisation: the schedule can be run, and
      // timed, on them -- every solution is 0, the waiting and the traffic are the real ones)
      GLSNS_CUDA(ctx, cudaMemsetAsync(sw.stream.p, 0, (size_t)std::max<int64_t>(off, 16), ctx->stream));
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
    GLSNS_TRY(schedule(false, ctx->trsv_l, ctx->levels_l, nullptr));
    GLSNS_TRY(schedule(true, ctx->trsv_u, ctx->levels_u, nullptr));

    // ---- profile-guided list order (opt-in: GLSNS_TRSV_TUNE=<rounds>) ----
    // The level order assumes that every level takes the same time everywhere; it does not
    // (chains advance several levels in the time a hop between teams takes), and a team then
    // keeps a chain that is ready waiting behind one that is late.  So: run the schedule once on
    // zero factors with the trace on, re-order the lists by when each block's inputs were there,
    // keep the new schedule if it is faster, repeat.  Measured: -4.6 % at 32^3 cells with the
    // round-1 kernel, but at 64^3 the first round is slower (7.85 -> 8.43 ms) and is discarded
    // after seconds of set-up time, and a tuned order depends on measured times, so the last
    // bits of an ILU application differed between two contexts.  Off by default: the schedule
    // is then a function of the sparsity pattern alone and results are reproducible run to run.
    const int rounds = getenv("GLSNS_TRSV_TUNE") ? atoi(getenv("GLSNS_TRSV_TUNE")) : 0;
    if (rounds > 0 && ng > 0)
      {
        DevBuf<double>             r, z;
        DevBuf<unsigned long long> tr;
        GLSNS_TRY(dev_alloc(ctx, r, (size_t)n));
        GLSNS_TRY(dev_alloc(ctx, z, (size_t)n));
        GLSNS_TRY(dev_alloc(ctx, tr, (size_t)10 * n));
        GLSNS_TRY(dev_alloc(ctx, ctx->ytmp, (size_t)n));
        GLSNS_TRY(dev_alloc(ctx, ctx->dinv, (size_t)n));
        GLSNS_CUDA(ctx, cudaMemsetAsync(r.p, 0, sizeof(double) * n, ctx->stream));
        GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->dinv.p, 0, sizeof(double) * n, ctx->stream));
        GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->counters.p, 0, 2 * sizeof(int32_t), ctx->stream));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        auto time_apply = [&](float &ms) -> glsns_status {
          ms = 1e30f;
          for (int rep = 0; rep < 3; ++rep)
            {
              cudaEventRecord(e0, ctx->stream);
              GLSNS_TRY(launch_ilu_apply(ctx, r.p, z.p, nullptr));
              cudaEventRecord(e1, ctx->stream);
              GLSNS_TRY(check_counters(ctx, "ILU apply (schedule tuning)"));
              float t = 0;
              cudaEventElapsedTime(&t, e0, e1);
              ms = std::min(ms, t);
            }
          return GLSNS_OK;
        };
        std::vector<unsigned long long> when((size_t)2 * n);
        float                           t_cur = 0;
        glsns_status                    st    = time_apply(t_cur);
        for (int round = 0; round < rounds && st == GLSNS_OK; ++round)
          {
            GLSNS_CUDA(ctx, cudaMemsetAsync(tr.p, 0, sizeof(unsigned long long) * 10 * n, ctx->stream));
            if ((st = launch_ilu_apply(ctx, r.p, z.p, tr.p)) != GLSNS_OK ||
                (st = check_counters(ctx, "ILU apply (schedule tuning)")) != GLSNS_OK)
              break;
            GLSNS_CUDA(ctx, cudaMemcpyAsync(when.data(), tr.p, sizeof(unsigned long long) * 2 * n,
                                            cudaMemcpyDeviceToHost, ctx->stream));
            GLSNS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            for (int sweep = 0; sweep < 2; ++sweep)
              { // times relative to the sweep's first publication (0 = not a row of the sweep)
                unsigned long long *w = when.data() + (size_t)sweep * n, t0 = ~0ull;
                for (int64_t i = 0; i < n; ++i)
                  if (w[i])
                    t0 = std::min(t0, w[i]);
                for (int64_t i = 0; i < n; ++i)
                  w[i] = w[i] ? w[i] - t0 + 1 : 0;
              }
            TrsvSweep keep_l, keep_u;
            std::swap(keep_l, ctx->trsv_l), std::swap(keep_u, ctx->trsv_u);
            std::vector<int32_t> rw_l = ctx->trsv_row_warp_l, rw_u = ctx->trsv_row_warp_u;
            int32_t              dummy;
            float                t_new = 1e30f;
            // a candidate that cannot be built (no memory for a second set of streams) or does
            // not run is simply not taken: the schedule in use stays
            const bool cand_ok = schedule(false, ctx->trsv_l, dummy, when.data()) == GLSNS_OK &&
                                 schedule(true, ctx->trsv_u, dummy, when.data() + n) == GLSNS_OK &&
                                 time_apply(t_new) == GLSNS_OK;
            if (!cand_ok)
              {
                cudaStreamSynchronize(ctx->stream);
                cudaGetLastError();
                cudaMemsetAsync(ctx->counters.p, 0, 2 * sizeof(int32_t), ctx->stream);
                t_new = 1e30f;
              }
            if (getenv("GLSNS_TRSV_DEBUG"))
              fprintf(stderr, "trsv_analyse: tuning round %d: %.3f ms -> %.3f ms\n", round, t_cur, t_new);
            if (t_new < t_cur)
              {
                t_cur = t_new;
                keep_l.release(), keep_u.release();
              }
            else
              { // no better: back to the previous schedule, done
                std::swap(keep_l, ctx->trsv_l), std::swap(keep_u, ctx->trsv_u);
                keep_l.release(), keep_u.release();
                ctx->trsv_row_warp_l = rw_l, ctx->trsv_row_warp_u = rw_u;
                break;
              }
          }
        cudaEventDestroy(e0), cudaEventDestroy(e1);
        r.release(), z.release(), tr.release();
        GLSNS_TRY(st);
      }
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
        trsv_pack_values_kernel<false><<<(unsigned)((nit * 32 + 127) / 128), 128, 0, ctx->stream>>>(
          nit, ctx->trsv_l.items.p, ctx->trsv_l.gdesc.p, ctx->trsv_l.blob_off.p, ctx->trsv_l.blob_off_v.p, ctx->lu.p,
          ctx->trsv_l.stream.p);
        ctx->kernel_launches++;
      }
    if (ctx->trsv_u.n_items)
      {
        const int64_t nit = ctx->trsv_u.n_items;
        trsv_pack_values_kernel<true><<<(unsigned)((nit * 32 + 127) / 128), 128, 0, ctx->stream>>>(
          nit, ctx->trsv_u.items.p, ctx->trsv_u.gdesc.p, ctx->trsv_u.blob_off.p, ctx->trsv_u.blob_off_v.p, ctx->lu.p,
          ctx->trsv_u.stream.p);
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
    // (GLSNS_TRSV_NODEPS=1, a measuring aid: every entry counts as published from the start, so
    // nothing ever waits for another team -- wrong results, and the time the machinery takes
    // when the dependencies cost nothing)
    static const int fill = getenv("GLSNS_TRSV_NODEPS") && atoi(getenv("GLSNS_TRSV_NODEPS")) ? 0 : 0xFF;
    GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->ytmp.p, fill, sizeof(double) * n, ctx->stream));
    GLSNS_CUDA(ctx, cudaMemsetAsync(z, fill, sizeof(double) * n, ctx->stream));
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
