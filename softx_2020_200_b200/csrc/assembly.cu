// Cell assembly of the GLS-stabilised Navier–Stokes Jacobian and residual.
//
// Replaces the cell loop of GLSNavierStokesSolver<dim>::assembleGLS
// (reference: source/solvers/gls_navier_stokes.cc:334-773) and the
// constraint-aware scatter distribute_local_to_global (:755-771).
//
// One cell per CTA.  The element's nodal values, the real-space shape features
// {N, grad N, lap N, u.grad N, L} at every quadrature point and the per-point
// field data {u, grad u, R, tau, JxW, ...} are staged in shared memory.  The
// local matrix is never materialised: thread (a,b) owns the (dim+1)x(dim+1)
// block that couples scalar shapes a and b, accumulates it over the quadrature
// points in registers in the structured form
//      uu: d_ij [nu gNa.gNb + adv_b Na + c0 Na Nb + tau L_b adv_a]
//          + (G+W)_ij (Na Nb + tau Nb adv_a) + tau R_i d_j Na Nb
//      up: -d_i Na Nb + tau d_i Nb adv_a
//      pu:  Na d_j Nb + tau [Nb (G+W)_:j . gNa + d_j Na L_b]
//      pp:  tau gNa.gNb
// (the B^T D B contraction of the reference's q x j x i tensor loop, :519-625)
// and adds it straight into the device CSR.  Cells are launched colour by
// colour, so no two CTAs of a launch touch the same matrix row or RHS entry:
// the scatter is a plain read-modify-write, atomic free and deterministic.
#include <math.h>

#include <algorithm>

#include "context.h"

namespace glsns
{
  namespace
  {
    constexpr int QS = 56; // doubles per quadrature-point record

    // offsets in the quadrature-point record
    constexpr int Q_U = 0, Q_G = 3, Q_LAP = 12, Q_P = 15, Q_GP = 16, Q_UDOT = 19, Q_R = 22,
                  Q_TAU = 25, Q_JXW = 26, Q_F = 27, Q_COR = 30, Q_CEN = 33, Q_GU = 36,
                  Q_DIVU = 39, Q_CN = 40, Q_GT = 43;

    struct AsmArgs
    {
      // fe
      int           n_su, n_sp, n_q, vel_degree;
      const double *shape_u, *grad_u, *hess_u, *shape_p, *grad_p, *weights;
      // cells
      const int32_t *cell_list; // cells of this colour
      const int32_t *cell_dofs;
      int            geometry_per_q;
      const double  *inv_jac, *det_jac, *measure, *q_points, *force;
      const double  *map_lap; // [n_cells][n_q][DIM] (geometry_per_q only): see glsns_mesh_desc
      // dofs
      int64_t        n_owned;
      const uint8_t *constrained;
      const int64_t *rowptr, *diag_pos;
      const int32_t *col;
      // hanging-node lines (null: none): dof i with constrained[i] == 2 stands for its masters
      const int64_t *hang_ptr;
      const int32_t *hang_idx;
      const double  *hang_w;
      // state
      const double *U, *U1, *U2, *U3;
      // parameters
      double nu, sdt, c[4], omega[3];
      int    transient, srf;
      // output
      double *val, *rhs;
    };

    __device__ __forceinline__ int64_t
    find_col(const int32_t *__restrict__ col, int64_t lo, int64_t hi, int32_t c)
    {
      // first position in [lo,hi) with col >= c (columns are sorted and c exists)
      while (lo < hi)
        {
          const int64_t mid = (lo + hi) >> 1;
          if (__ldg(col + mid) < c)
            lo = mid + 1;
          else
            hi = mid;
        }
      return lo;
    }

    template <int DIM, bool MATRIX>
    __global__ void __launch_bounds__(256)
    assemble_cells(const AsmArgs A)
    {
      extern __shared__ double smem[];
      const int n_su = A.n_su, n_sp = A.n_sp, nq = A.n_q;
      const int n    = DIM * n_su + n_sp;
      const int tid = threadIdx.x, nt = blockDim.x;
      const int64_t cell = A.cell_list[blockIdx.x];

      // ---- shared memory carve-up ----
      double *sNu  = smem;                  // [nq][n_su]
      double *sGu  = sNu + nq * n_su;       // [nq][n_su][DIM]
      double *sLap = sGu + nq * n_su * DIM; // [nq][n_su]
      double *sAdv = sLap + nq * n_su;      // [nq][n_su]
      double *sL   = sAdv + nq * n_su;      // [nq][n_su]
      double *sNp  = sL + nq * n_su;        // [nq][n_sp]
      double *sGp  = sNp + nq * n_sp;       // [nq][n_sp][DIM]
      double *sU   = sGp + nq * n_sp * DIM; // [n]
      double *sUd  = sU + n;                // [n]  c0 U + c1 U1 + c2 U2 + c3 U3
      double *sQ   = sUd + n;               // [nq][QS]
      int64_t *sRow = (int64_t *)(sQ + nq * QS); // [n][2] row start / end
      int32_t *sDof = (int32_t *)(sRow + 2 * n); // [n]
      int32_t *sCon = sDof + n;                  // [n]

      // ---- phase 0: element dofs and nodal values ----
      __shared__ int sHang; // does the cell have a dof with a hanging-node line?
      if (tid == 0)
        sHang = 0;
      __syncthreads();
      for (int k = tid; k < n; k += nt)
        {
          const int32_t g = A.cell_dofs[cell * n + k];
          sDof[k]         = g;
          sCon[k]         = A.constrained[g];
          if (sCon[k] == 2)
            sHang = 1;
          sU[k]           = A.U[g];
          double ud       = 0;
          if (A.transient)
            {
              ud = A.c[0] * A.U[g];
              if (A.U1)
                ud += A.c[1] * A.U1[g];
              if (A.U2)
                ud += A.c[2] * A.U2[g];
              if (A.U3)
                ud += A.c[3] * A.U3[g];
            }
          sUd[k] = ud;
          if (g < A.n_owned)
            {
              sRow[2 * k]     = A.rowptr[g];
              sRow[2 * k + 1] = A.rowptr[g + 1];
            }
          else
            sRow[2 * k] = sRow[2 * k + 1] = 0;
        }

      // ---- phase 1: real-space shape features (what FEValues::reinit provides) ----
      for (int idx = tid; idx < nq * n_su; idx += nt)
        {
          const int     q  = idx / n_su;
          const double *iJ = A.inv_jac + (A.geometry_per_q ? (cell * nq + q) : cell) * DIM * DIM;
          double        J[DIM][DIM];
#pragma unroll
          for (int r = 0; r < DIM; ++r)
#pragma unroll
            for (int d = 0; d < DIM; ++d)
              J[r][d] = iJ[r * DIM + d];
          const double *gr = A.grad_u + (size_t)idx * DIM;
          const double *hr = A.hess_u + (size_t)idx * DIM * DIM;
          double        lap = 0;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              double g = 0;
#pragma unroll
              for (int r = 0; r < DIM; ++r)
                g += gr[r] * J[r][d];
              sGu[idx * DIM + d] = g;
              // curved cell: the part of the real-space Hessian that comes from the mapping's
              // own second derivatives, -grad_x N . c (c from the host, see glsns_mesh_desc)
              if (A.geometry_per_q)
                lap -= g * A.map_lap[((size_t)cell * nq + q) * DIM + d];
            }
#pragma unroll
          for (int r = 0; r < DIM; ++r)
#pragma unroll
            for (int s = 0; s < DIM; ++s)
              {
                double k = 0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                  k += J[r][d] * J[s][d];
                lap += hr[r * DIM + s] * k;
              }
          sLap[idx] = lap;
          sNu[idx]  = A.shape_u[idx];
        }
      for (int idx = tid; idx < nq * n_sp; idx += nt)
        {
          const int     q  = idx / n_sp;
          const double *iJ = A.inv_jac + (A.geometry_per_q ? (cell * nq + q) : cell) * DIM * DIM;
          const double *gr = A.grad_p + (size_t)idx * DIM;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              double g = 0;
#pragma unroll
              for (int r = 0; r < DIM; ++r)
                g += gr[r] * iJ[r * DIM + d];
              sGp[idx * DIM + d] = g;
            }
          sNp[idx] = A.shape_p[idx];
        }
      __syncthreads();

      // ---- phase 2: fields at the quadrature points (:351-384) ----
      {
        constexpr int NF = DIM + DIM * DIM + DIM + 1 + DIM + DIM;
        for (int idx = tid; idx < nq * NF; idx += nt)
          {
            const int q = idx / NF, f = idx - q * NF;
            double    s = 0;
            double   *Q = sQ + q * QS;
            if (f < DIM) // u_c
              {
                const double *Uc = sU + f * n_su, *N = sNu + q * n_su;
                for (int a = 0; a < n_su; ++a)
                  s += Uc[a] * N[a];
                Q[Q_U + f] = s;
              }
            else if (f < DIM + DIM * DIM) // G[c][d]
              {
                const int     c = (f - DIM) / DIM, d = (f - DIM) - c * DIM;
                const double *Uc = sU + c * n_su, *G = sGu + q * n_su * DIM + d;
                for (int a = 0; a < n_su; ++a)
                  s += Uc[a] * G[a * DIM];
                Q[Q_G + c * 3 + d] = s;
              }
            else if (f < 2 * DIM + DIM * DIM) // laplacian of u_c
              {
                const int     c  = f - DIM - DIM * DIM;
                const double *Uc = sU + c * n_su, *Lp = sLap + q * n_su;
                for (int a = 0; a < n_su; ++a)
                  s += Uc[a] * Lp[a];
                Q[Q_LAP + c] = s;
              }
            else if (f == 2 * DIM + DIM * DIM) // p
              {
                const double *Up = sU + DIM * n_su, *N = sNp + q * n_sp;
                for (int a = 0; a < n_sp; ++a)
                  s += Up[a] * N[a];
                Q[Q_P] = s;
              }
            else if (f < 3 * DIM + DIM * DIM + 1) // grad p
              {
                const int     d  = f - (2 * DIM + DIM * DIM + 1);
                const double *Up = sU + DIM * n_su, *G = sGp + q * n_sp * DIM + d;
                for (int a = 0; a < n_sp; ++a)
                  s += Up[a] * G[a * DIM];
                Q[Q_GP + d] = s;
              }
            else // time-derivative combination sum_k c_k u^(k)_c
              {
                const int c = f - (3 * DIM + DIM * DIM + 1);
                if (A.transient)
                  {
                    const double *Uc = sUd + c * n_su, *N = sNu + q * n_su;
                    for (int a = 0; a < n_su; ++a)
                      s += Uc[a] * N[a];
                  }
                Q[Q_UDOT + c] = s;
              }
          }
      }
      __syncthreads();

      // ---- phase 3: tau, strong residual and RHS coefficients per point (:391-516) ----
      if (tid < nq)
        {
          const int q = tid;
          double   *Q = sQ + q * QS;
          double    u[DIM], R[DIM], f[DIM], cor[DIM], cen[DIM];
          double    un = 0;
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            {
              u[c] = Q[Q_U + c];
              un += u[c] * u[c];
              f[c]   = A.force ? A.force[((size_t)cell * nq + q) * DIM + c] : 0.0;
              cor[c] = cen[c] = 0;
            }
          const double u_mag = fmax(sqrt(un), 1e-12);
          const double meas  = A.measure[cell];
          const double h     = (DIM == 2 ? sqrt(4. * meas / M_PI) : cbrt(6. * meas / M_PI)) /
                           A.vel_degree;
          const double a1 = 2. * u_mag / h, a2 = 4. * A.nu / (h * h);
          const double tau =
            1. / sqrt((A.transient ? A.sdt * A.sdt : 0.0) + a1 * a1 + 9. * a2 * a2);
          const double det = A.det_jac[A.geometry_per_q ? (cell * nq + q) : cell];
          if (A.srf)
            {
              const double *x = A.q_points + ((size_t)cell * nq + q) * DIM;
              const double *w = A.omega;
              if (DIM == 2)
                {
                  cor[0] = -2. * w[2] * u[1];
                  cor[1] = 2. * w[2] * u[0];
                  cen[0] = -w[2] * w[2] * x[0];
                  cen[1] = -w[2] * w[2] * x[1];
                }
              else
                {
                  cor[0] = 2. * (w[1] * u[2] - w[2] * u[1]);
                  cor[1] = 2. * (w[2] * u[0] - w[0] * u[2]);
                  cor[DIM - 1] = 2. * (w[0] * u[1] - w[1] * u[0]);
                  const double t0 = w[1] * x[DIM - 1] - w[2] * x[1],
                               t1 = w[2] * x[0] - w[0] * x[DIM - 1],
                               t2 = w[0] * x[1] - w[1] * x[0];
                  cen[0]       = w[1] * t2 - w[2] * t1;
                  cen[1]       = w[2] * t0 - w[0] * t2;
                  cen[DIM - 1] = w[0] * t1 - w[1] * t0;
                }
            }
          double divu = 0;
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            {
              double gu = 0;
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                gu += Q[Q_G + c * 3 + d] * u[d];
              divu += Q[Q_G + c * 3 + c];
              R[c] = gu + Q[Q_GP + c] - A.nu * Q[Q_LAP + c] - f[c] + cor[c] + cen[c] +
                     Q[Q_UDOT + c];
              Q[Q_GU + c]  = gu;
              Q[Q_R + c]   = R[c];
              Q[Q_F + c]   = f[c];
              Q[Q_COR + c] = cor[c];
              Q[Q_CEN + c] = cen[c];
              Q[Q_CN + c]  = -gu + f[c] - Q[Q_UDOT + c] - cor[c] - cen[c];
            }
          Q[Q_DIVU] = divu;
          Q[Q_TAU]  = tau;
          Q[Q_JXW]  = det * A.weights[q];
          // G + W, W = matrix of phi -> 2 omega x phi (:535-544)
#pragma unroll
          for (int c = 0; c < DIM; ++c)
#pragma unroll
            for (int d = 0; d < DIM; ++d)
              Q[Q_GT + c * 3 + d] = Q[Q_G + c * 3 + d];
          if (A.srf)
            {
              const double *w = A.omega;
              if (DIM == 2)
                {
                  Q[Q_GT + 0 * 3 + 1] += -2. * w[2];
                  Q[Q_GT + 1 * 3 + 0] += 2. * w[2];
                }
              else
                {
                  Q[Q_GT + 0 * 3 + 1] += -2. * w[2];
                  Q[Q_GT + 0 * 3 + (DIM - 1)] += 2. * w[1];
                  Q[Q_GT + 1 * 3 + 0] += 2. * w[2];
                  Q[Q_GT + 1 * 3 + (DIM - 1)] += -2. * w[0];
                  Q[Q_GT + (DIM - 1) * 3 + 0] += -2. * w[1];
                  Q[Q_GT + (DIM - 1) * 3 + 1] += 2. * w[0];
                }
            }
        }
      __syncthreads();

      // ---- phase 4: u.grad N and L = u.grad N - nu lap N + c0 N ----
      const double c0 = A.transient ? A.c[0] : 0.0;
      for (int idx = tid; idx < nq * n_su; idx += nt)
        {
          const int     q = idx / n_su;
          const double *Q = sQ + q * QS;
          double        adv = 0;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            adv += Q[Q_U + d] * sGu[idx * DIM + d];
          sAdv[idx] = adv;
          sL[idx]   = adv - A.nu * sLap[idx] + c0 * sNu[idx];
        }
      __syncthreads();

      // ---- phase 5: residual (:628-748) ----
      for (int i = tid; i < n; i += nt)
        {
          const int32_t gi = sDof[i];
          if (sCon[i] == 1 || (sCon[i] == 2 && !A.hang_ptr) || gi >= A.n_owned)
            continue;
          double s = 0;
          if (i < DIM * n_su)
            {
              const int c = i / n_su, a = i - c * n_su;
              for (int q = 0; q < nq; ++q)
                {
                  const double *Q  = sQ + q * QS;
                  const double *ga = sGu + (q * n_su + a) * DIM;
                  double        Gg = 0;
#pragma unroll
                  for (int d = 0; d < DIM; ++d)
                    Gg += Q[Q_G + c * 3 + d] * ga[d];
                  s += Q[Q_JXW] * (-A.nu * Gg + Q[Q_CN + c] * sNu[q * n_su + a] +
                                   Q[Q_P] * ga[c] - Q[Q_TAU] * Q[Q_R + c] * sAdv[q * n_su + a]);
                }
            }
          else
            {
              const int a = i - DIM * n_su;
              for (int q = 0; q < nq; ++q)
                {
                  const double *Q  = sQ + q * QS;
                  const double *ga = sGp + (q * n_sp + a) * DIM;
                  double        Rg = 0;
#pragma unroll
                  for (int d = 0; d < DIM; ++d)
                    Rg += Q[Q_R + d] * ga[d];
                  s += Q[Q_JXW] * (-Q[Q_DIVU] * sNp[q * n_sp + a] - Q[Q_TAU] * Rg);
                }
            }
          if (!sHang)
            A.rhs[gi] += s; // (cells of one colour share no dof)
          else if (sCon[i] == 0)
            atomicAdd(A.rhs + gi, s); // (a master of a hanging dof of this very cell may be gi)
          else
            for (int64_t k = A.hang_ptr[gi]; k < A.hang_ptr[gi + 1]; ++k)
              atomicAdd(A.rhs + A.hang_idx[k], A.hang_w[k] * s);
        }

      // ---- phase 6: Jacobian blocks (:519-625) and scatter (:755-771) ----
      if (MATRIX)
        {
          const int nmax = n_su > n_sp ? n_su : n_sp;
          for (int pidx = tid; pidx < nmax * nmax; pidx += nt)
            {
              const int  a = pidx / nmax, b = pidx - a * nmax;
              const bool au = a < n_su, bu = b < n_su, ap = a < n_sp, bp = b < n_sp;
              double     uu[DIM][DIM], up[DIM], pu[DIM], pp = 0;
#pragma unroll
              for (int c = 0; c < DIM; ++c)
                {
                  up[c] = pu[c] = 0;
#pragma unroll
                  for (int d = 0; d < DIM; ++d)
                    uu[c][d] = 0;
                }
              for (int q = 0; q < nq; ++q)
                {
                  const double *Q   = sQ + q * QS;
                  const double  tau = Q[Q_TAU], JxW = Q[Q_JXW];
                  double        Na = 0, adva = 0, ga[DIM], Nb = 0, Lb = 0, advb = 0, gb[DIM];
                  double        Npa = 0, gpa[DIM], Npb = 0, gpb[DIM];
#pragma unroll
                  for (int d = 0; d < DIM; ++d)
                    ga[d] = gb[d] = gpa[d] = gpb[d] = 0;
                  if (au)
                    {
                      Na   = sNu[q * n_su + a];
                      adva = sAdv[q * n_su + a];
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        ga[d] = sGu[(q * n_su + a) * DIM + d];
                    }
                  if (bu)
                    {
                      Nb   = sNu[q * n_su + b];
                      advb = sAdv[q * n_su + b];
                      Lb   = sL[q * n_su + b];
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        gb[d] = sGu[(q * n_su + b) * DIM + d];
                    }
                  if (ap)
                    {
                      Npa = sNp[q * n_sp + a];
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        gpa[d] = sGp[(q * n_sp + a) * DIM + d];
                    }
                  if (bp)
                    {
                      Npb = sNp[q * n_sp + b];
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        gpb[d] = sGp[(q * n_sp + b) * DIM + d];
                    }
                  if (au && bu)
                    {
                      double gg = 0;
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        gg += ga[d] * gb[d];
                      const double diag = A.nu * gg + advb * Na + c0 * Na * Nb + tau * Lb * adva;
                      const double s    = Na * Nb + tau * Nb * adva;
                      const double t    = tau * Nb;
#pragma unroll
                      for (int ci = 0; ci < DIM; ++ci)
                        {
                          const double tr = t * Q[Q_R + ci];
#pragma unroll
                          for (int cj = 0; cj < DIM; ++cj)
                            {
                              double v = Q[Q_GT + ci * 3 + cj] * s + tr * ga[cj];
                              if (ci == cj)
                                v += diag;
                              uu[ci][cj] += JxW * v;
                            }
                        }
                    }
                  if (au && bp)
                    {
#pragma unroll
                      for (int ci = 0; ci < DIM; ++ci)
                        up[ci] += JxW * (-ga[ci] * Npb + tau * gpb[ci] * adva);
                    }
                  if (ap && bu)
                    {
#pragma unroll
                      for (int cj = 0; cj < DIM; ++cj)
                        {
                          double gtg = 0;
#pragma unroll
                          for (int c = 0; c < DIM; ++c)
                            gtg += Q[Q_GT + c * 3 + cj] * gpa[c];
                          pu[cj] += JxW * (Npa * gb[cj] + tau * (Nb * gtg + Lb * gpa[cj]));
                        }
                    }
                  if (ap && bp)
                    {
                      double gg = 0;
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        gg += gpa[d] * gpb[d];
                      pp += JxW * tau * gg;
                    }
                }

              // scatter (AffineConstraints::distribute_local_to_global, :755-771): a constrained
              // row keeps |local(i,i)| on its diagonal, couplings to Dirichlet columns are
              // dropped; a dof with a hanging-node line stands for its masters, row and column
              auto put = [&](const int i, const int j, const double v) {
                const int32_t gi = sDof[i];
                if (gi >= A.n_owned)
                  return;
                if (!sHang)
                  { // no hanging node in the cell: plain read-modify-write (coloured cells)
                    if (sCon[i])
                      {
                        if (i == j)
                          A.val[A.diag_pos[gi]] += fabs(v);
                        return;
                      }
                    if (!sCon[j])
                      A.val[find_col(A.col, sRow[2 * i], sRow[2 * i + 1], sDof[j])] += v;
                    return;
                  }
                // a cell with hanging nodes: two of its entries may land on one matrix entry
                // (a master is often a dof of the same cell), hence atomics
                if (sCon[i] && i == j)
                  atomicAdd(A.val + A.diag_pos[gi], fabs(v));
                if (sCon[i] == 1 || sCon[j] == 1)
                  return;
                const int32_t gj = sDof[j];
                const int64_t ib = sCon[i] == 2 ? A.hang_ptr[gi] : 0, ie = sCon[i] == 2 ? A.hang_ptr[gi + 1] : 1;
                const int64_t jb = sCon[j] == 2 ? A.hang_ptr[gj] : 0, je = sCon[j] == 2 ? A.hang_ptr[gj + 1] : 1;
                for (int64_t ki = ib; ki < ie; ++ki)
                  {
                    const int32_t ri = sCon[i] == 2 ? A.hang_idx[ki] : gi;
                    const double  wi = sCon[i] == 2 ? A.hang_w[ki] : 1.0;
                    const int64_t rs = A.rowptr[ri], re = A.rowptr[ri + 1];
                    for (int64_t kj = jb; kj < je; ++kj)
                      {
                        const int32_t cj = sCon[j] == 2 ? A.hang_idx[kj] : gj;
                        const double  wj = sCon[j] == 2 ? A.hang_w[kj] : 1.0;
                        atomicAdd(A.val + find_col(A.col, rs, re, cj), wi * wj * v);
                      }
                  }
              };
              if (au)
                {
#pragma unroll
                  for (int ci = 0; ci < DIM; ++ci)
                    {
                      const int i = ci * n_su + a;
                      if (bu)
                        {
#pragma unroll
                          for (int cj = 0; cj < DIM; ++cj)
                            put(i, cj * n_su + b, uu[ci][cj]);
                        }
                      if (bp)
                        put(i, DIM * n_su + b, up[ci]);
                    }
                }
              if (ap)
                {
                  const int i = DIM * n_su + a;
                  if (bu)
                    {
#pragma unroll
                      for (int cj = 0; cj < DIM; ++cj)
                        put(i, cj * n_su + b, pu[cj]);
                    }
                  if (bp)
                    put(i, DIM * n_su + b, pp);
                }
            }
        }
    }

    // ---- assemble_L2_projection (reference: source/solvers/gls_navier_stokes.cc:829-914) ----
    // The mass system of set_initial_condition(L2projection): local(i,j) = (phi_u_j . phi_u_i +
    // phi_p_j phi_p_i) JxW (non-zero only inside one solution component), local_rhs(i) =
    // (phi_u_i . u0 + phi_p_i p0) JxW with (u0, p0) the initial-condition function at the
    // quadrature points, scattered with the NONZERO constraints (:902-909,
    // AffineConstraints::distribute_local_to_global with its default
    // use_inhomogeneities_for_rhs = false): couplings to a constrained column j move to the
    // right-hand side as -local(i,j) g_j, a constrained row keeps |local(i,i)| on its diagonal
    // and a zero right-hand side (solve_system_GMRES(initial_step = true) distributes the
    // constraint values afterwards).  One cell per CTA, cells colour by colour.
    struct L2Args
    {
      int            n_su, n_sp, n_q;
      const double  *shape_u, *shape_p, *weights;
      const int32_t *cell_list, *cell_dofs;
      int            geometry_per_q;
      const double  *det_jac, *init; // init: [n_cells][n_q][dim + 1]
      int64_t        n_owned;
      const uint8_t *constrained;
      const double  *cvalues; // may be null (all constraints homogeneous)
      const int64_t *rowptr, *diag_pos;
      const int32_t *col;
      double        *val, *rhs;
    };

    template <int DIM>
    __global__ void __launch_bounds__(256)
    l2_projection_cells(const L2Args A)
    {
      extern __shared__ double smem[];
      const int     n_su = A.n_su, n_sp = A.n_sp, nq = A.n_q;
      const int     n    = DIM * n_su + n_sp;
      const int     tid = threadIdx.x, nt = blockDim.x;
      const int64_t cell = A.cell_list[blockIdx.x];
      double  *sJxW = smem;     // [nq]
      double  *sG   = sJxW + nq; // [n] constraint values
      int64_t *sRow = (int64_t *)(sG + n);
      int32_t *sDof = (int32_t *)(sRow + 2 * n);
      int32_t *sCon = sDof + n;
      for (int q = tid; q < nq; q += nt)
        sJxW[q] = A.det_jac[A.geometry_per_q ? (cell * nq + q) : cell] * A.weights[q];
      for (int k = tid; k < n; k += nt)
        {
          const int32_t g = A.cell_dofs[cell * n + k];
          sDof[k]         = g;
          sCon[k]         = A.constrained[g];
          sG[k]           = (A.constrained[g] && A.cvalues) ? A.cvalues[g] : 0.0;
          if (g < A.n_owned)
            sRow[2 * k] = A.rowptr[g], sRow[2 * k + 1] = A.rowptr[g + 1];
          else
            sRow[2 * k] = sRow[2 * k + 1] = 0;
        }
      __syncthreads();
      // matrix: thread (a, b) owns the scalar mass entries of shapes a and b
      const int nmax = n_su > n_sp ? n_su : n_sp;
      for (int pair = tid; pair < nmax * nmax; pair += nt)
        {
          const int  a = pair / nmax, b = pair - a * nmax;
          const bool uu = a < n_su && b < n_su, pp = a < n_sp && b < n_sp;
          double     mu = 0, mp = 0;
          for (int q = 0; q < nq; ++q)
            {
              if (uu)
                mu += A.shape_u[q * n_su + a] * A.shape_u[q * n_su + b] * sJxW[q];
              if (pp)
                mp += A.shape_p[q * n_sp + a] * A.shape_p[q * n_sp + b] * sJxW[q];
            }
          for (int c = 0; c <= DIM; ++c)
            {
              if (c < DIM ? !uu : !pp)
                continue;
              const int     i = c < DIM ? c * n_su + a : DIM * n_su + a;
              const int     j = c < DIM ? c * n_su + b : DIM * n_su + b;
              const double  m = c < DIM ? mu : mp;
              const int32_t gi = sDof[i];
              if (gi >= A.n_owned)
                continue;
              if (sCon[i])
                {
                  if (a == b)
                    A.val[A.diag_pos[gi]] += fabs(m);
                  continue;
                }
              if (sCon[j])
                continue;
              A.val[find_col(A.col, sRow[2 * i], sRow[2 * i + 1], sDof[j])] += m;
            }
        }
      // right-hand side: one thread per local dof (fixed summation order)
      for (int i = tid; i < n; i += nt)
        {
          const int32_t gi = sDof[i];
          if (gi >= A.n_owned || sCon[i])
            continue;
          const int     c  = i < DIM * n_su ? i / n_su : DIM;
          const int     ns = c < DIM ? n_su : n_sp, base = c < DIM ? c * n_su : DIM * n_su;
          const int     a  = i - base;
          const double *N  = c < DIM ? A.shape_u : A.shape_p;
          double        r  = 0;
          for (int q = 0; q < nq; ++q)
            r += N[q * ns + a] * A.init[(cell * nq + q) * (DIM + 1) + c] * sJxW[q];
          for (int b = 0; b < ns; ++b)
            if (sCon[base + b])
              {
                double m = 0;
                for (int q = 0; q < nq; ++q)
                  m += N[q * ns + a] * N[q * ns + b] * sJxW[q];
                r -= m * sG[base + b];
              }
          A.rhs[gi] += r;
        }
    }

    // ---- calculate_CFL (reference: source/solvers/postprocessing_cfl.cc:34-87) ----
    // max over cells of |u(cell centre)| / h * dt, h from the cell measure (:69-72), the
    // velocity at the one point of QGauss(1).  One thread per cell, block maxima in `partial`.
    template <int DIM>
    __global__ void __launch_bounds__(256)
    cfl_cells(const int64_t n_cells, const int n_su, const int n_loc, const double degree,
              const int32_t *__restrict__ cell_dofs, const double *__restrict__ measure,
              const double *__restrict__ shape_centre, const double *__restrict__ U,
              const double dt, double *__restrict__ partial)
    {
      __shared__ double sh[256];
      double            best = 0;
      for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells;
           c += (int64_t)gridDim.x * blockDim.x)
        {
          const double h = (DIM == 2 ? sqrt(4. * measure[c] / M_PI) : pow(6 * measure[c] / M_PI, 1. / 3.)) / degree;
          double       u2 = 0;
          for (int d = 0; d < DIM; ++d)
            {
              double u = 0;
              for (int a = 0; a < n_su; ++a)
                u += shape_centre[a] * U[cell_dofs[c * n_loc + d * n_su + a]];
              u2 += u * u;
            }
          best = fmax(best, sqrt(u2) / h * dt);
        }
      sh[threadIdx.x] = best;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1)
        {
          if (threadIdx.x < o)
            sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]);
          __syncthreads();
        }
      if (threadIdx.x == 0)
        partial[blockIdx.x] = sh[0];
    }
    __global__ void __launch_bounds__(256)
    max_partials_kernel(const int n, const double *__restrict__ partial, double *__restrict__ out)
    {
      __shared__ double sh[256];
      double            best = 0;
      for (int i = threadIdx.x; i < n; i += 256)
        best = fmax(best, partial[i]);
      sh[threadIdx.x] = best;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1)
        {
          if (threadIdx.x < o)
            sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]);
          __syncthreads();
        }
      if (threadIdx.x == 0)
        *out = sh[0];
    }

    // Constrained rows that no cell contributed to get a unit diagonal.  A Dirichlet row always
    // receives |local(i,i)| > 0 from its cells; a row stays empty when the host has identified its
    // dof with another one in the cell -> dof table (periodic faces: the cells refer to the master
    // dofs, the slave rows are only carried along), and ILU needs a pivot there.
    __global__ void __launch_bounds__(256)
    unit_diagonal_on_empty_constrained_rows(const int64_t n, const uint8_t *__restrict__ constrained,
                                            const int64_t *__restrict__ diag_pos, double *val)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i < n && constrained[i] && val[diag_pos[i]] == 0.0)
        val[diag_pos[i]] = 1.0;
    }

    size_t
    assembly_smem_bytes(int dim, int n_su, int n_sp, int nq)
    {
      const int n = dim * n_su + n_sp;
      size_t    d = (size_t)nq * n_su * (4 + dim) + (size_t)nq * n_sp * (1 + dim) + 2 * n +
                 (size_t)nq * QS;
      return d * sizeof(double) + (size_t)2 * n * sizeof(int64_t) + (size_t)2 * n * sizeof(int32_t);
    }
  } // namespace

  glsns_status
  launch_assembly(glsns_context *ctx, bool assemble_matrix, bool transient, double sdt,
                  const double coefs[4])
  {
    AsmArgs A;
    A.n_su = ctx->n_su, A.n_sp = ctx->n_sp, A.n_q = ctx->n_q, A.vel_degree = ctx->vel_degree;
    A.shape_u = ctx->shape_u.p, A.grad_u = ctx->grad_u.p, A.hess_u = ctx->hess_u.p;
    A.shape_p = ctx->shape_p.p, A.grad_p = ctx->grad_p.p, A.weights = ctx->weights.p;
    A.cell_dofs      = ctx->cell_dofs.p;
    A.geometry_per_q = ctx->geometry_per_q;
    A.inv_jac = ctx->inv_jac.p, A.det_jac = ctx->det_jac.p, A.measure = ctx->measure.p;
    A.q_points    = ctx->q_points.p;
    A.map_lap     = ctx->map_lap.p;
    A.force       = ctx->have_force ? ctx->force.p : nullptr;
    A.n_owned     = ctx->n_owned;
    A.constrained = ctx->constrained.p;
    A.rowptr = ctx->rowptr.p, A.diag_pos = ctx->diag_pos.p, A.col = ctx->col.p;
    A.hang_ptr = ctx->n_hanging ? ctx->hang_ptr.p : nullptr;
    A.hang_idx = ctx->hang_idx.p, A.hang_w = ctx->hang_w.p;
    A.U  = ctx->vec[GLSNS_VEC_EVALUATION_POINT].p;
    A.U1 = ctx->vec_set[GLSNS_VEC_SOLUTION_M1] ? ctx->vec[GLSNS_VEC_SOLUTION_M1].p : nullptr;
    A.U2 = ctx->vec_set[GLSNS_VEC_SOLUTION_M2] ? ctx->vec[GLSNS_VEC_SOLUTION_M2].p : nullptr;
    A.U3 = ctx->vec_set[GLSNS_VEC_SOLUTION_M3] ? ctx->vec[GLSNS_VEC_SOLUTION_M3].p : nullptr;
    A.nu = ctx->viscosity, A.sdt = sdt;
    for (int i = 0; i < 4; ++i)
      A.c[i] = transient ? coefs[i] : 0.0;
    // history vectors that carry a zero coefficient are not read
    if (A.c[1] == 0.0) A.U1 = nullptr;
    if (A.c[2] == 0.0) A.U2 = nullptr;
    if (A.c[3] == 0.0) A.U3 = nullptr;
    for (int i = 0; i < 3; ++i)
      A.omega[i] = ctx->omega[i];
    A.transient = transient ? 1 : 0;
    A.srf       = ctx->srf;
    A.val = ctx->val.p, A.rhs = ctx->vec[GLSNS_VEC_SYSTEM_RHS].p;

    if (A.srf && !A.q_points)
      return fail(ctx, GLSNS_ERR_STATE, "velocity source srf needs mesh.q_points");

    const size_t smem = assembly_smem_bytes(ctx->dim, A.n_su, A.n_sp, A.n_q);
    if (smem > 227 * 1024)
      return fail(ctx, GLSNS_ERR_UNSUPPORTED, "element too large for the shared-memory staging");
    auto kern = ctx->dim == 2 ? (assemble_matrix ? assemble_cells<2, true> : assemble_cells<2, false>) :
                                (assemble_matrix ? assemble_cells<3, true> : assemble_cells<3, false>);
    GLSNS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    if (assemble_matrix)
      GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->val.p, 0, sizeof(double) * ctx->nnz, ctx->stream));
    GLSNS_CUDA(ctx, cudaMemsetAsync(A.rhs, 0, sizeof(double) * ctx->n_owned, ctx->stream));
    for (int c = 0; c < ctx->n_colors; ++c)
      {
        const int32_t n_in = ctx->color_ptr[c + 1] - ctx->color_ptr[c];
        if (n_in == 0)
          continue;
        A.cell_list = ctx->color_cells.p + ctx->color_ptr[c];
        kern<<<n_in, 256, smem, ctx->stream>>>(A);
        ctx->kernel_launches++;
      }
    if (assemble_matrix && ctx->n_owned)
      {
        unit_diagonal_on_empty_constrained_rows<<<(unsigned)((ctx->n_owned + 255) / 256), 256, 0,
                                                  ctx->stream>>>(ctx->n_owned, ctx->constrained.p,
                                                                 ctx->diag_pos.p, ctx->val.p);
        ctx->kernel_launches++;
      }
    GLSNS_CUDA(ctx, cudaGetLastError());
    return GLSNS_OK;
  }
  glsns_status
  launch_l2_projection(glsns_context *ctx, const double *init_dev)
  {
    L2Args A;
    A.n_su = ctx->n_su, A.n_sp = ctx->n_sp, A.n_q = ctx->n_q;
    A.shape_u = ctx->shape_u.p, A.shape_p = ctx->shape_p.p, A.weights = ctx->weights.p;
    A.cell_dofs      = ctx->cell_dofs.p;
    A.geometry_per_q = ctx->geometry_per_q;
    A.det_jac = ctx->det_jac.p, A.init = init_dev;
    A.n_owned     = ctx->n_owned;
    A.constrained = ctx->constrained.p;
    A.cvalues     = ctx->cvalues.p;
    A.rowptr = ctx->rowptr.p, A.diag_pos = ctx->diag_pos.p, A.col = ctx->col.p;
    A.val = ctx->val.p, A.rhs = ctx->vec[GLSNS_VEC_SYSTEM_RHS].p;
    const int    n    = ctx->dim * A.n_su + A.n_sp;
    const size_t smem = sizeof(double) * (A.n_q + n) + sizeof(int64_t) * 2 * n + sizeof(int32_t) * 2 * n;
    auto         kern = ctx->dim == 2 ? l2_projection_cells<2> : l2_projection_cells<3>;
    GLSNS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GLSNS_CUDA(ctx, cudaMemsetAsync(ctx->val.p, 0, sizeof(double) * ctx->nnz, ctx->stream));
    GLSNS_CUDA(ctx, cudaMemsetAsync(A.rhs, 0, sizeof(double) * ctx->n_owned, ctx->stream));
    for (int c = 0; c < ctx->n_colors; ++c)
      {
        const int32_t n_in = ctx->color_ptr[c + 1] - ctx->color_ptr[c];
        if (n_in == 0)
          continue;
        A.cell_list = ctx->color_cells.p + ctx->color_ptr[c];
        kern<<<n_in, 256, smem, ctx->stream>>>(A);
        ctx->kernel_launches++;
      }
    if (ctx->n_owned)
      {
        unit_diagonal_on_empty_constrained_rows<<<(unsigned)((ctx->n_owned + 255) / 256), 256, 0,
                                                  ctx->stream>>>(ctx->n_owned, ctx->constrained.p,
                                                                 ctx->diag_pos.p, ctx->val.p);
        ctx->kernel_launches++;
      }
    GLSNS_CUDA(ctx, cudaGetLastError());
    return GLSNS_OK;
  }

  // *out_dev = max over the local cells (all-reduced over ranks)
  glsns_status
  launch_cfl(glsns_context *ctx, const double *shape_centre_dev, const double *U, double dt,
             double degree, double *out_dev)
  {
    const int64_t nc = ctx->n_cells;
    const int     grid = (int)std::max<int64_t>(1, std::min<int64_t>((nc + 255) / 256, (int64_t)ctx->n_sm * 8));
    GLSNS_TRY(dev_alloc(ctx, ctx->partials, (size_t)64 * ctx->n_sm * 8));
    if (ctx->dim == 2)
      cfl_cells<2><<<grid, 256, 0, ctx->stream>>>(nc, ctx->n_su, ctx->n_loc, degree, ctx->cell_dofs.p,
                                                   ctx->measure.p, shape_centre_dev, U, dt, ctx->partials.p);
    else
      cfl_cells<3><<<grid, 256, 0, ctx->stream>>>(nc, ctx->n_su, ctx->n_loc, degree, ctx->cell_dofs.p,
                                                   ctx->measure.p, shape_centre_dev, U, dt, ctx->partials.p);
    max_partials_kernel<<<1, 256, 0, ctx->stream>>>(grid, ctx->partials.p, out_dev);
    ctx->kernel_launches += 2;
    GLSNS_CUDA(ctx, cudaGetLastError());
    return allreduce_max(ctx, out_dev, 1);
  }
} // namespace glsns
