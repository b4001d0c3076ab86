/*
 * oracle/gls_oracle.c — CPU restatement of Lethe's GLS Navier–Stokes hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product path (softx_2020_200_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  The chain mesh -> assembly -> constraint scatter ->
 * ILU(0) -> GMRES(30) -> Newton built from these functions reproduces the
 * reference's golden files tests/solvers/restart_01.output (GMRES iteration
 * counts 8/6/10, true residuals, L2 error 0.0343628) and
 * applications_tests/gls_navier_stokes_3d/mms3d_gls.output (see
 * tests/test_oracle_golden.py).
 *
 * The reference itself (deal.II 9.2 + Trilinos + p4est + MPI) cannot be built in
 * this image, so there is no oracle/_ref; GMRES and ILU live in Trilinos
 * (AztecOO / Ifpack, reached through deal.II 9.2.0 TrilinosWrappers, version
 * pinned only by the docker image named in the reference's .travis.yml) and are
 * restated here from their published algorithms.
 *
 * Each function cites the reference file:line it follows
 * (paths relative to the reference root).
 *
 * Local dof layout used everywhere in this repo (not deal.II's, which is an
 * internal detail of FESystem): k = c*n_su + a for velocity component c < dim
 * and scalar shape a < n_su;  k = dim*n_su + a for pressure shape a < n_sp.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 3

typedef struct
{
  int           dim, n_su, n_sp, nq, vel_degree;
  const double *Nu;   /* [nq][n_su]            velocity scalar shape values      */
  const double *dNu;  /* [nq][n_su][dim]       reference-cell gradients          */
  const double *d2Nu; /* [nq][n_su][dim][dim]  reference-cell hessians           */
  const double *Np;   /* [nq][n_sp]                                               */
  const double *dNp;  /* [nq][n_sp][dim]                                          */
  const double *wq;   /* [nq] quadrature weights on the unit cell                 */
} glso_fe;

typedef struct
{
  int64_t        ncell;
  const int32_t *cell_dofs;    /* [ncell][n]                                       */
  const double  *cell_invJ;    /* [ncell][dim][dim]: invJ[r][d] = d xi_r / d x_d  */
  const double  *cell_detJ;    /* [ncell]                                          */
  const double  *cell_measure; /* [ncell]                                          */
  const double  *qpoints;      /* [ncell][nq][dim] or NULL (needed for SRF only)  */
  const double  *force;        /* [ncell][nq][dim] or NULL (NoForce)              */
  /* MappingQ on curved cells (`qmapping all`, gls_navier_stokes.cc:244-252): cell_invJ and
     cell_detJ are then per (cell, q), and map_lap [ncell][nq][dim] holds
     c_k = sum_rs d2x_k/dxi_r dxi_s (J^-1 J^-T)_rs, the part of the real-space Hessian trace that
     comes from the mapping:  lap N = H_ref(N) : (J^-1 J^-T) - grad_x N . c  (what FEValues
     delivers with update_hessians, :417-422) */
  int            geometry_per_q;
  const double  *map_lap;
} glso_cells;

typedef struct
{
  double viscosity;
  int    transient; /* 0: steady tau; 1: tau includes (1/dt)^2                    */
  double sdt;       /* 1/dt                                                        */
  double coefs[4];  /* c0 multiplies the present solution, c1..c3 solution_m1..m3 */
  int    srf;       /* velocity source: 0 none, 1 rotating frame                  */
  double omega[3];
} glso_params;

/* ---- literal cell kernel ----------------------------------------------------
 * Follows source/solvers/gls_navier_stokes.cc:338-749 statement by statement:
 * loop order q, then j (column) outer / i (row) inner, full Tensor<1,dim> /
 * Tensor<2,dim> algebra for every (i,j) pair although every velocity shape
 * function has a single non-zero component.
 * The BDF/SDIRK terms (:477-516, :563-569, :644-701) are written with the one
 * coefficient vector coefs[]: for every scheme the reference adds
 *   sum_k c_k u^(k)  to the strong residual,  c_0 phi_j phi_i to the Jacobian and
 *   -sum_k c_k (u^(k) . phi_i) to the RHS
 * (bdf1's RHS is spelled -c0 (u - u1) there, identical because c1 = -c0).
 */
static void
cell_literal(const glso_fe *fe, const glso_cells *cs, const glso_params *pr,
             int64_t cell, const double *U, const double *U1, const double *U2,
             const double *U3, int assemble_matrix, double *M, double *b,
             double *scratch)
{
  const int dim = fe->dim, n_su = fe->n_su, n_sp = fe->n_sp, nq = fe->nq;
  const int n   = dim * n_su + n_sp;
  const int32_t *dofs = cs->cell_dofs + cell * n;
  const double  *iJ   = cs->cell_invJ + cell * dim * dim;
  const double   nu   = pr->viscosity;
  const double   zero3[MAXD] = {0, 0, 0};

  /* per-dof shape data at one q point (:282-288) */
  double *phi_u      = scratch;               /* [n][MAXD]       */
  double *grad_phi_u = phi_u + n * MAXD;      /* [n][MAXD][MAXD] */
  double *lap_phi_u  = grad_phi_u + n * 9;    /* [n][MAXD]       */
  double *div_phi_u  = lap_phi_u + n * MAXD;  /* [n]             */
  double *phi_p      = div_phi_u + n;         /* [n]             */
  double *grad_phi_p = phi_p + n;             /* [n][MAXD]       */

  if (assemble_matrix)
    memset(M, 0, sizeof(double) * n * n);
  memset(b, 0, sizeof(double) * n);

  /* :340-345 */
  double h;
  if (dim == 2)
    h = sqrt(4. * cs->cell_measure[cell] / M_PI) / fe->vel_degree;
  else
    h = pow(6 * cs->cell_measure[cell] / M_PI, 1. / 3.) / fe->vel_degree;

  for (int q = 0; q < nq; ++q)
    {
      /* :412-423 shape functions in real space (affine cell: grad = invJ^T grad_ref,
         hessian = invJ^T H_ref invJ; what FEValues::reinit does on such a cell; on a curved
         cell the mapping's second derivatives add -grad . c to the Hessian trace) */
      const double *mlap = zero3;
      if (cs->geometry_per_q)
        {
          iJ   = cs->cell_invJ + ((size_t)cell * nq + q) * dim * dim;
          mlap = cs->map_lap + ((size_t)cell * nq + q) * dim;
        }
      memset(phi_u, 0, sizeof(double) * n * (MAXD + 9 + MAXD + 1 + 1 + MAXD));
      for (int c = 0; c < dim; ++c)
        for (int a = 0; a < n_su; ++a)
          {
            const int     k  = c * n_su + a;
            const double *gr = fe->dNu + ((size_t)q * n_su + a) * dim;
            const double *hr = fe->d2Nu + ((size_t)q * n_su + a) * dim * dim;
            double        g[MAXD] = {0, 0, 0}, lap = 0;
            for (int d = 0; d < dim; ++d)
              for (int r = 0; r < dim; ++r)
                g[d] += gr[r] * iJ[r * dim + d];
            for (int d = 0; d < dim; ++d)
              for (int r = 0; r < dim; ++r)
                for (int s = 0; s < dim; ++s)
                  lap += hr[r * dim + s] * iJ[r * dim + d] * iJ[s * dim + d];
            for (int d = 0; d < dim; ++d)
              lap -= g[d] * mlap[d];
            phi_u[k * MAXD + c] = fe->Nu[(size_t)q * n_su + a];
            for (int d = 0; d < dim; ++d)
              grad_phi_u[k * 9 + c * MAXD + d] = g[d];
            lap_phi_u[k * MAXD + c] = lap;
            div_phi_u[k]            = g[c];
          }
      for (int a = 0; a < n_sp; ++a)
        {
          const int     k  = dim * n_su + a;
          const double *gr = fe->dNp + ((size_t)q * n_sp + a) * dim;
          phi_p[k]         = fe->Np[(size_t)q * n_sp + a];
          for (int d = 0; d < dim; ++d)
            {
              double g = 0;
              for (int r = 0; r < dim; ++r)
                g += gr[r] * iJ[r * dim + d];
              grad_phi_p[k * MAXD + d] = g;
            }
        }

      /* :351-362, :373-384 function values at q */
      double u[MAXD] = {0, 0, 0}, G[MAXD][MAXD] = {{0}}, lap_u[MAXD] = {0, 0, 0};
      double p = 0, gp[MAXD] = {0, 0, 0};
      double u1[MAXD] = {0, 0, 0}, u2[MAXD] = {0, 0, 0}, u3[MAXD] = {0, 0, 0};
      for (int k = 0; k < n; ++k)
        {
          const double v = U[dofs[k]];
          for (int c = 0; c < dim; ++c)
            {
              u[c] += v * phi_u[k * MAXD + c];
              lap_u[c] += v * lap_phi_u[k * MAXD + c];
              for (int d = 0; d < dim; ++d)
                G[c][d] += v * grad_phi_u[k * 9 + c * MAXD + d];
            }
          p += v * phi_p[k];
          for (int d = 0; d < dim; ++d)
            gp[d] += v * grad_phi_p[k * MAXD + d];
          if (U1)
            for (int c = 0; c < dim; ++c)
              u1[c] += U1[dofs[k]] * phi_u[k * MAXD + c];
          if (U2)
            for (int c = 0; c < dim; ++c)
              u2[c] += U2[dofs[k]] * phi_u[k * MAXD + c];
          if (U3)
            for (int c = 0; c < dim; ++c)
              u3[c] += U3[dofs[k]] * phi_u[k * MAXD + c];
        }

      /* :391-408 */
      double unorm = 0;
      for (int c = 0; c < dim; ++c)
        unorm += u[c] * u[c];
      const double u_mag = fmax(sqrt(unorm), 1e-12 * 1.0 /*GLS_u_scale*/);
      const double JxW =
        cs->cell_detJ[cs->geometry_per_q ? (size_t)cell * nq + q : (size_t)cell] * fe->wq[q];
      const double tau =
        !pr->transient ?
          1. / sqrt(pow(2. * u_mag / h, 2) + 9 * pow(4 * nu / (h * h), 2)) :
          1. / sqrt(pow(pr->sdt, 2) + pow(2. * u_mag / h, 2) +
                    9 * pow(4 * nu / (h * h), 2));

      /* :425-431 */
      double force[MAXD] = {0, 0, 0};
      if (cs->force)
        for (int c = 0; c < dim; ++c)
          force[c] = cs->force[((size_t)cell * nq + q) * dim + c];

      const double div_u = (dim == 2) ? G[0][0] + G[1][1] : G[0][0] + G[1][1] + G[2][2];

      /* :438-441 */
      double R[MAXD] = {0, 0, 0};
      for (int c = 0; c < dim; ++c)
        {
          double conv = 0;
          for (int d = 0; d < dim; ++d)
            conv += G[c][d] * u[d];
          R[c] = conv + gp[c] - nu * lap_u[c] - force[c];
        }

      /* :443-467 rotating frame: Coriolis + centrifugal */
      double cor[MAXD] = {0, 0, 0}, cen[MAXD] = {0, 0, 0};
      const double *om = pr->omega;
      if (pr->srf)
        {
          const double *x = cs->qpoints + ((size_t)cell * nq + q) * dim;
          if (dim == 2)
            {
              const double wz = om[2];
              /* 2*wz*(-1)*cross_product_2d(u), cross_product_2d(v) = (v1,-v0) */
              cor[0] = 2 * wz * (-1.) * u[1];
              cor[1] = 2 * wz * (-1.) * (-u[0]);
              /* wz*(-1)*cp2d(wz*(-1)*cp2d(x)) */
              double t0 = wz * (-1.) * x[1], t1 = wz * (-1.) * (-x[0]);
              cen[0] = wz * (-1.) * t1;
              cen[1] = wz * (-1.) * (-t0);
            }
          else
            {
              cor[0] = 2 * (om[1] * u[2] - om[2] * u[1]);
              cor[1] = 2 * (om[2] * u[0] - om[0] * u[2]);
              cor[2] = 2 * (om[0] * u[1] - om[1] * u[0]);
              double t[3] = {om[1] * x[2] - om[2] * x[1], om[2] * x[0] - om[0] * x[2],
                             om[0] * x[1] - om[1] * x[0]};
              cen[0] = om[1] * t[2] - om[2] * t[1];
              cen[1] = om[2] * t[0] - om[0] * t[2];
              cen[2] = om[0] * t[1] - om[1] * t[0];
            }
          for (int c = 0; c < dim; ++c)
            R[c] += cor[c] + cen[c];
        }

      /* :477-516 */
      double udot[MAXD] = {0, 0, 0};
      if (pr->transient)
        for (int c = 0; c < dim; ++c)
          {
            udot[c] = pr->coefs[0] * u[c] + pr->coefs[1] * u1[c] +
                      pr->coefs[2] * u2[c] + pr->coefs[3] * u3[c];
            R[c] += udot[c];
          }

      /* :519-625 */
      if (assemble_matrix)
        for (int j = 0; j < n; ++j)
          {
            const double *pj = phi_u + j * MAXD, *gj = grad_phi_u + j * 9;
            double        sj[MAXD] = {0, 0, 0}; /* strong_jac :525-544 */
            double        Gphi[MAXD] = {0, 0, 0}, gphiu[MAXD] = {0, 0, 0};
            for (int c = 0; c < dim; ++c)
              {
                for (int d = 0; d < dim; ++d)
                  {
                    Gphi[c] += G[c][d] * pj[d];
                    gphiu[c] += gj[c * MAXD + d] * u[d];
                  }
                sj[c] = Gphi[c] + gphiu[c] + grad_phi_p[j * MAXD + c] -
                        nu * lap_phi_u[j * MAXD + c];
                if (pr->transient)
                  sj[c] += pj[c] * pr->coefs[0];
              }
            double corj[MAXD] = {0, 0, 0};
            if (pr->srf)
              {
                if (dim == 2)
                  {
                    corj[0] = 2 * om[2] * (-1.) * pj[1];
                    corj[1] = 2 * om[2] * (-1.) * (-pj[0]);
                  }
                else
                  {
                    corj[0] = 2 * (om[1] * pj[2] - om[2] * pj[1]);
                    corj[1] = 2 * (om[2] * pj[0] - om[0] * pj[2]);
                    corj[2] = 2 * (om[0] * pj[1] - om[1] * pj[0]);
                  }
                for (int c = 0; c < dim; ++c)
                  sj[c] += corj[c];
              }
            for (int i = 0; i < n; ++i)
              {
                const double *pi = phi_u + i * MAXD, *gi = grad_phi_u + i * 9;
                double        gg = 0, a1 = 0, a2 = 0, mass = 0;
                for (int c = 0; c < dim; ++c)
                  {
                    for (int d = 0; d < dim; ++d)
                      gg += gj[c * MAXD + d] * gi[c * MAXD + d];
                    a1 += Gphi[c] * pi[c];
                    a2 += gphiu[c] * pi[c];
                    mass += pj[c] * pi[c];
                  }
                /* :548-560 */
                double m = (nu * gg + a1 + a2 - div_phi_u[i] * phi_p[j] +
                            phi_p[i] * div_phi_u[j]) *
                           JxW;
                /* :563-569 */
                if (pr->transient)
                  m += mass * pr->coefs[0] * JxW;
                /* :572-573 PSPG */
                double pspg = 0;
                for (int c = 0; c < dim; ++c)
                  pspg += sj[c] * grad_phi_p[i * MAXD + c];
                m += tau * pspg * JxW;
                /* :575-587 */
                if (pr->srf)
                  {
                    double cc = 0;
                    for (int c = 0; c < dim; ++c)
                      cc += corj[c] * pi[c];
                    m += cc * JxW;
                  }
                /* :599-606 SUPG (tau-derivative terms are disabled in the
                   reference, :590-596 and :608-621) */
                double s1 = 0, s2 = 0;
                for (int c = 0; c < dim; ++c)
                  {
                    double giu = 0, gipj = 0;
                    for (int d = 0; d < dim; ++d)
                      {
                        giu += gi[c * MAXD + d] * u[d];
                        gipj += gi[c * MAXD + d] * pj[d];
                      }
                    s1 += sj[c] * giu;
                    s2 += R[c] * gipj;
                  }
                m += tau * (s1 + s2) * JxW;
                M[(size_t)i * n + j] += m;
              }
          }

      /* :628-748 */
      for (int i = 0; i < n; ++i)
        {
          const double *pi = phi_u + i * MAXD, *gi = grad_phi_u + i * 9;
          double        Ggi = 0, conv = 0, fphi = 0;
          for (int c = 0; c < dim; ++c)
            {
              double Gu = 0;
              for (int d = 0; d < dim; ++d)
                {
                  Ggi += G[c][d] * gi[c * MAXD + d];
                  Gu += G[c][d] * u[d];
                }
              conv += Gu * pi[c];
              fphi += force[c] * pi[c];
            }
          double r = (-nu * Ggi - conv + p * div_phi_u[i] + fphi - div_u * phi_p[i]) * JxW;
          if (pr->transient)
            {
              double t = 0;
              for (int c = 0; c < dim; ++c)
                t += udot[c] * pi[c];
              r -= t * JxW;
            }
          if (pr->srf)
            {
              double t = 0, t2 = 0;
              for (int c = 0; c < dim; ++c)
                {
                  t += cor[c] * pi[c];
                  t2 += cen[c] * pi[c];
                }
              r += -t * JxW;
              r += -t2 * JxW;
            }
          double pspg = 0, supg = 0;
          for (int c = 0; c < dim; ++c)
            {
              double giu = 0;
              for (int d = 0; d < dim; ++d)
                giu += gi[c * MAXD + d] * u[d];
              pspg += R[c] * grad_phi_p[i * MAXD + c];
              supg += R[c] * giu;
            }
          r += -tau * pspg * JxW;
          r += -tau * supg * JxW;
          b[i] += r;
        }
    }
}

/* ---- structured cell kernel ("Mode B" of BASELINE.md section 3) ------------------
 * The same local matrix and right-hand side as cell_literal, computed the way a tuned CPU
 * code would: every velocity shape function is N_a e_c, so the n x n matrix is a
 * (dim+1) x (dim+1) grid of scalar n_s x n_s blocks built from the per-point features
 * {N, grad N, lap N, u.grad N} (derived from
 * source/solvers/gls_navier_stokes.cc:525-606, :628-748).  About ten times fewer flops than
 * the literal q x j x i loop.  Not the reference's arithmetic order: it is the honest
 * best-CPU baseline of bench.py, and tests/test_oracle_golden.py checks it against
 * cell_literal to rounding.  Needs n_su == n_sp or not: pressure has its own tables. */
static void
cell_structured(const glso_fe *fe, const glso_cells *cs, const glso_params *pr,
                int64_t cell, const double *U, const double *U1, const double *U2,
                const double *U3, int assemble_matrix, double *M, double *b,
                double *scratch)
{
  const int dim = fe->dim, n_su = fe->n_su, n_sp = fe->n_sp, nq = fe->nq;
  const int n   = dim * n_su + n_sp;
  const int32_t *dofs = cs->cell_dofs + cell * n;
  const double  *iJ   = cs->cell_invJ + cell * dim * dim;
  const double   nu   = pr->viscosity;
  const double   c0   = pr->transient ? pr->coefs[0] : 0.0;
  double *N   = scratch;             /* [n_su]       */
  double *dN  = N + n_su;            /* [n_su][MAXD] */
  double *lap = dN + n_su * MAXD;    /* [n_su]       */
  double *adv = lap + n_su;          /* [n_su]       */
  double *L   = adv + n_su;          /* [n_su]       */
  double *Np  = L + n_su;            /* [n_sp]       */
  double *dNp = Np + n_sp;           /* [n_sp][MAXD] */
  if (assemble_matrix)
    memset(M, 0, sizeof(double) * n * n);
  memset(b, 0, sizeof(double) * n);
  double h;
  if (dim == 2)
    h = sqrt(4. * cs->cell_measure[cell] / M_PI) / fe->vel_degree;
  else
    h = pow(6 * cs->cell_measure[cell] / M_PI, 1. / 3.) / fe->vel_degree;
  double W[MAXD][MAXD] = {{0}};
  if (pr->srf)
    {
      const double *om = pr->omega;
      if (dim == 2)
        W[0][1] = -2 * om[2], W[1][0] = 2 * om[2];
      else
        {
          W[0][1] = -2 * om[2], W[0][2] = 2 * om[1], W[1][0] = 2 * om[2];
          W[1][2] = -2 * om[0], W[2][0] = -2 * om[1], W[2][1] = 2 * om[0];
        }
    }
  const double zero3[MAXD] = {0, 0, 0};
  for (int q = 0; q < nq; ++q)
    {
      const double *mlap = zero3;
      if (cs->geometry_per_q)
        {
          iJ   = cs->cell_invJ + ((size_t)cell * nq + q) * dim * dim;
          mlap = cs->map_lap + ((size_t)cell * nq + q) * dim;
        }
      for (int a = 0; a < n_su; ++a)
        {
          const double *gr = fe->dNu + ((size_t)q * n_su + a) * dim;
          const double *hr = fe->d2Nu + ((size_t)q * n_su + a) * dim * dim;
          double        l  = 0;
          for (int d = 0; d < dim; ++d)
            {
              double g = 0;
              for (int r = 0; r < dim; ++r)
                g += gr[r] * iJ[r * dim + d];
              dN[a * MAXD + d] = g;
              for (int r = 0; r < dim; ++r)
                for (int t = 0; t < dim; ++t)
                  l += hr[r * dim + t] * iJ[r * dim + d] * iJ[t * dim + d];
              l -= g * mlap[d];
            }
          lap[a] = l;
          N[a]   = fe->Nu[(size_t)q * n_su + a];
        }
      for (int a = 0; a < n_sp; ++a)
        {
          const double *gr = fe->dNp + ((size_t)q * n_sp + a) * dim;
          Np[a]            = fe->Np[(size_t)q * n_sp + a];
          for (int d = 0; d < dim; ++d)
            {
              double g = 0;
              for (int r = 0; r < dim; ++r)
                g += gr[r] * iJ[r * dim + d];
              dNp[a * MAXD + d] = g;
            }
        }
      double u[MAXD] = {0, 0, 0}, G[MAXD][MAXD] = {{0}}, lap_u[MAXD] = {0, 0, 0};
      double p = 0, gp[MAXD] = {0, 0, 0}, udot[MAXD] = {0, 0, 0};
      for (int c = 0; c < dim; ++c)
        for (int a = 0; a < n_su; ++a)
          {
            const int32_t g = dofs[c * n_su + a];
            const double  v = U[g];
            u[c] += v * N[a];
            lap_u[c] += v * lap[a];
            for (int d = 0; d < dim; ++d)
              G[c][d] += v * dN[a * MAXD + d];
            if (pr->transient)
              udot[c] += (pr->coefs[0] * v + (U1 ? pr->coefs[1] * U1[g] : 0.0) +
                          (U2 ? pr->coefs[2] * U2[g] : 0.0) + (U3 ? pr->coefs[3] * U3[g] : 0.0)) *
                         N[a];
          }
      for (int a = 0; a < n_sp; ++a)
        {
          const double v = U[dofs[dim * n_su + a]];
          p += v * Np[a];
          for (int d = 0; d < dim; ++d)
            gp[d] += v * dNp[a * MAXD + d];
        }
      double unorm = 0, div_u = 0;
      for (int c = 0; c < dim; ++c)
        unorm += u[c] * u[c], div_u += G[c][c];
      const double u_mag = fmax(sqrt(unorm), 1e-12);
      const double JxW =
        cs->cell_detJ[cs->geometry_per_q ? (size_t)cell * nq + q : (size_t)cell] * fe->wq[q];
      const double tau =
        !pr->transient ?
          1. / sqrt(pow(2. * u_mag / h, 2) + 9 * pow(4 * nu / (h * h), 2)) :
          1. / sqrt(pow(pr->sdt, 2) + pow(2. * u_mag / h, 2) + 9 * pow(4 * nu / (h * h), 2));
      double force[MAXD] = {0, 0, 0}, R[MAXD], Gu[MAXD], body[MAXD] = {0, 0, 0};
      if (cs->force)
        for (int c = 0; c < dim; ++c)
          force[c] = cs->force[((size_t)cell * nq + q) * dim + c];
      if (pr->srf)
        { /* Coriolis + centrifugal (:443-467) */
          const double *x = cs->qpoints + ((size_t)cell * nq + q) * dim, *om = pr->omega;
          double        cen[MAXD] = {0, 0, 0};
          if (dim == 2)
            cen[0] = -om[2] * om[2] * x[0], cen[1] = -om[2] * om[2] * x[1];
          else
            {
              const double t[3] = {om[1] * x[2] - om[2] * x[1], om[2] * x[0] - om[0] * x[2],
                                   om[0] * x[1] - om[1] * x[0]};
              cen[0] = om[1] * t[2] - om[2] * t[1];
              cen[1] = om[2] * t[0] - om[0] * t[2];
              cen[2] = om[0] * t[1] - om[1] * t[0];
            }
          for (int c = 0; c < dim; ++c)
            {
              body[c] = cen[c];
              for (int d = 0; d < dim; ++d)
                body[c] += W[c][d] * u[d];
            }
        }
      for (int c = 0; c < dim; ++c)
        {
          Gu[c] = 0;
          for (int d = 0; d < dim; ++d)
            Gu[c] += G[c][d] * u[d];
          R[c] = Gu[c] + gp[c] - nu * lap_u[c] - force[c] + body[c] + udot[c];
        }
      for (int a = 0; a < n_su; ++a)
        {
          double s = 0;
          for (int d = 0; d < dim; ++d)
            s += u[d] * dN[a * MAXD + d];
          adv[a] = s;
          L[a]   = s - nu * lap[a] + c0 * N[a];
        }
      if (assemble_matrix)
        {
          /* velocity rows */
          for (int a = 0; a < n_su; ++a)
            {
              const double Na = N[a], adva = adv[a], *ga = dN + a * MAXD;
              for (int b2 = 0; b2 < n_su; ++b2)
                {
                  const double Nb = N[b2], *gb = dN + b2 * MAXD;
                  double       gg = 0;
                  for (int d = 0; d < dim; ++d)
                    gg += ga[d] * gb[d];
                  const double diag = nu * gg + adv[b2] * Na + c0 * Na * Nb + tau * L[b2] * adva;
                  const double NN = Na * Nb, tNa = tau * Nb * adva;
                  for (int ci = 0; ci < dim; ++ci)
                    for (int cj = 0; cj < dim; ++cj)
                      M[(size_t)(ci * n_su + a) * n + cj * n_su + b2] +=
                        ((ci == cj ? diag : 0.0) + (G[ci][cj] + W[ci][cj]) * (NN + tNa) +
                         tau * R[ci] * ga[cj] * Nb) *
                        JxW;
                }
              for (int b2 = 0; b2 < n_sp; ++b2)
                for (int ci = 0; ci < dim; ++ci)
                  M[(size_t)(ci * n_su + a) * n + dim * n_su + b2] +=
                    (-ga[ci] * Np[b2] + tau * dNp[b2 * MAXD + ci] * adva) * JxW;
            }
          /* pressure rows */
          for (int a = 0; a < n_sp; ++a)
            {
              const double *ga = dNp + a * MAXD;
              for (int b2 = 0; b2 < n_su; ++b2)
                for (int cj = 0; cj < dim; ++cj)
                  {
                    double t = 0;
                    for (int c = 0; c < dim; ++c)
                      t += (G[c][cj] + W[c][cj]) * ga[c];
                    M[(size_t)(dim * n_su + a) * n + cj * n_su + b2] +=
                      (Np[a] * dN[b2 * MAXD + cj] + tau * (N[b2] * t + ga[cj] * L[b2])) * JxW;
                  }
              for (int b2 = 0; b2 < n_sp; ++b2)
                {
                  double gg = 0;
                  for (int d = 0; d < dim; ++d)
                    gg += ga[d] * dNp[b2 * MAXD + d];
                  M[(size_t)(dim * n_su + a) * n + dim * n_su + b2] += tau * gg * JxW;
                }
            }
        }
      for (int a = 0; a < n_su; ++a)
        for (int ci = 0; ci < dim; ++ci)
          {
            double Gg = 0;
            for (int d = 0; d < dim; ++d)
              Gg += G[ci][d] * dN[a * MAXD + d];
            b[ci * n_su + a] += (-nu * Gg + p * dN[a * MAXD + ci] +
                                 (force[ci] - Gu[ci] - udot[ci] - body[ci]) * N[a] -
                                 tau * R[ci] * adv[a]) *
                                JxW;
          }
      for (int a = 0; a < n_sp; ++a)
        {
          double t = 0;
          for (int d = 0; d < dim; ++d)
            t += R[d] * dNp[a * MAXD + d];
          b[dim * n_su + a] += (-div_u * Np[a] - tau * t) * JxW;
        }
    }
}

static inline int64_t
csr_find(const int64_t *rowptr, const int32_t *col, int32_t row, int32_t c)
{
  int64_t lo = rowptr[row], hi = rowptr[row + 1] - 1;
  while (lo <= hi)
    {
      int64_t mid = (lo + hi) >> 1;
      if (col[mid] == c)
        return mid;
      if (col[mid] < c)
        lo = mid + 1;
      else
        hi = mid - 1;
    }
  return -1;
}

/* Constraint-aware scatter — the behaviour of
 * AffineConstraints::distribute_local_to_global for homogeneous Dirichlet
 * constraints as used at source/solvers/gls_navier_stokes.cc:755-771 (sparsity
 * built with keep_constrained_dofs = false, :204-208): a constrained row keeps
 * only its diagonal, which receives |local(i,i)|; its RHS entry stays 0;
 * couplings to constrained columns are dropped.  */
/* Hanging-node lines of the constraints (Kelly-refined meshes, navier_stokes_base.cc:610-729;
 * DoFTools::make_hanging_node_constraints in setup_dofs): dof i with constrained[i] == 2 is
 * x_i = sum_k w_k x_{m_k} over the entries hptr[i] .. hptr[i+1] (masters unconstrained: the closed
 * form of zero_constraints, Dirichlet masters dropped).  Set by glso_set_hanging, NULL = none. */
static const int64_t *g_hptr = NULL;
static const int32_t *g_hidx = NULL;
static const double  *g_hw   = NULL;
void
glso_set_hanging(const int64_t *ptr, const int32_t *idx, const double *w)
{
  g_hptr = ptr, g_hidx = idx, g_hw = w;
}

static void
scatter_cell(int n, const int32_t *dofs, const uint8_t *constrained,
             const int64_t *rowptr, const int32_t *col, int assemble_matrix,
             const double *M, const double *b, double *val, double *rhs)
{
  if (g_hptr)
    { /* distribute_local_to_global with hanging-node lines: every local dof stands for its
         masters (itself with weight 1 when unconstrained, nothing when Dirichlet); a constrained
         row -- Dirichlet or hanging -- keeps |local(i,i)| on its own diagonal */
      for (int i = 0; i < n; ++i)
        {
          const int32_t gi = dofs[i];
          if (constrained[gi] && assemble_matrix)
            val[csr_find(rowptr, col, gi, gi)] += fabs(M[(size_t)i * n + i]);
          const int64_t ib = constrained[gi] == 2 ? g_hptr[gi] : 0,
                        ie = constrained[gi] == 2 ? g_hptr[gi + 1] : (constrained[gi] ? 0 : 1);
          for (int64_t ki = ib; ki < ie; ++ki)
            {
              const int32_t ri = constrained[gi] == 2 ? g_hidx[ki] : gi;
              const double  wi = constrained[gi] == 2 ? g_hw[ki] : 1.0;
              rhs[ri] += wi * b[i];
              if (!assemble_matrix)
                continue;
              for (int j = 0; j < n; ++j)
                {
                  const int32_t gj = dofs[j];
                  const int64_t jb = constrained[gj] == 2 ? g_hptr[gj] : 0,
                                je = constrained[gj] == 2 ? g_hptr[gj + 1] : (constrained[gj] ? 0 : 1);
                  for (int64_t kj = jb; kj < je; ++kj)
                    {
                      const int32_t cj = constrained[gj] == 2 ? g_hidx[kj] : gj;
                      const double  wj = constrained[gj] == 2 ? g_hw[kj] : 1.0;
                      val[csr_find(rowptr, col, ri, cj)] += wi * wj * M[(size_t)i * n + j];
                    }
                }
            }
        }
      return;
    }
  for (int i = 0; i < n; ++i)
    {
      const int32_t gi = dofs[i];
      if (constrained[gi])
        {
          if (assemble_matrix)
            val[csr_find(rowptr, col, gi, gi)] += fabs(M[(size_t)i * n + i]);
          continue;
        }
      rhs[gi] += b[i];
      if (assemble_matrix)
        for (int j = 0; j < n; ++j)
          if (!constrained[dofs[j]])
            val[csr_find(rowptr, col, gi, dofs[j])] += M[(size_t)i * n + j];
    }
}

/* 0: the reference-faithful literal loop (default, what every parity test uses);
 * 1: the structured form (bench.py's "Mode B" CPU figure only) */
static int glso_cell_mode = 0;
void
glso_set_cell_mode(int structured)
{
  glso_cell_mode = structured ? 1 : 0;
}
static void
cell_kernel(const glso_fe *fe, const glso_cells *cs, const glso_params *pr, int64_t cell,
            const double *U, const double *U1, const double *U2, const double *U3,
            int assemble_matrix, double *M, double *b, double *scratch)
{
  if (glso_cell_mode)
    cell_structured(fe, cs, pr, cell, U, U1, U2, U3, assemble_matrix, M, b, scratch);
  else
    cell_literal(fe, cs, pr, cell, U, U1, U2, U3, assemble_matrix, M, b, scratch);
}

static size_t
scratch_doubles(int n)
{
  return (size_t)n * (MAXD + 9 + MAXD + 1 + 1 + MAXD);
}

/* assembleGLS<assemble_matrix, scheme, velocity_source>,
 * source/solvers/gls_navier_stokes.cc:231-777, serial cell order 0..ncell-1
 * (the reference's cell loop :334). U1..U3 may be NULL.  If local_out != NULL the
 * un-scattered local matrices/rhs are also stored ([ncell][n*n], [ncell][n]). */
int
glso_assemble(const glso_fe *fe, const glso_cells *cs, const glso_params *pr,
              int64_t ndof, const double *U, const double *U1, const double *U2,
              const double *U3, const uint8_t *constrained, const int64_t *rowptr,
              const int32_t *col, int assemble_matrix, double *val, double *rhs,
              double *localM_out, double *localb_out)
{
  const int n = fe->dim * fe->n_su + fe->n_sp;
  double   *M = (double *)malloc(sizeof(double) * n * n);
  double   *b = (double *)malloc(sizeof(double) * n);
  double   *s = (double *)malloc(sizeof(double) * scratch_doubles(n));
  if (assemble_matrix)
    memset(val, 0, sizeof(double) * rowptr[ndof]); /* :237-238 */
  memset(rhs, 0, sizeof(double) * ndof);           /* :239 */
  for (int64_t c = 0; c < cs->ncell; ++c)
    {
      cell_kernel(fe, cs, pr, c, U, U1, U2, U3, assemble_matrix, M, b, s);
      if (localM_out && assemble_matrix)
        memcpy(localM_out + (size_t)c * n * n, M, sizeof(double) * n * n);
      if (localb_out)
        memcpy(localb_out + (size_t)c * n, b, sizeof(double) * n);
      scatter_cell(n, cs->cell_dofs + c * n, constrained, rowptr, col,
                   assemble_matrix, M, b, val, rhs);
    }
  free(M);
  free(b);
  free(s);
  return 0;
}

/* Same arithmetic, threaded the way the reference is parallel: the cell range
 * is cut into `nthreads` contiguous blocks ("one MPI rank per core"); colours
 * (cell_color[ncell], ncolor) serialise cells that share dofs so the scatter
 * needs no atomics.  Used only for the timed CPU baseline. */
int
glso_assemble_mt(const glso_fe *fe, const glso_cells *cs, const glso_params *pr,
                 int64_t ndof, const double *U, const double *U1, const double *U2,
                 const double *U3, const uint8_t *constrained,
                 const int64_t *rowptr, const int32_t *col, int assemble_matrix,
                 double *val, double *rhs, const int32_t *color_ptr,
                 const int32_t *color_cells, int ncolor)
{
  const int n = fe->dim * fe->n_su + fe->n_sp;
  if (assemble_matrix)
    {
      const int64_t nnz = rowptr[ndof];
#pragma omp parallel for schedule(static)
      for (int64_t k = 0; k < nnz; ++k)
        val[k] = 0;
    }
  memset(rhs, 0, sizeof(double) * ndof);
#pragma omp parallel
  {
    double *M = (double *)malloc(sizeof(double) * n * n);
    double *b = (double *)malloc(sizeof(double) * n);
    double *s = (double *)malloc(sizeof(double) * scratch_doubles(n));
    for (int col_i = 0; col_i < ncolor; ++col_i)
      {
#pragma omp for schedule(static)
        for (int32_t t = color_ptr[col_i]; t < color_ptr[col_i + 1]; ++t)
          {
            const int64_t c = color_cells[t];
            cell_kernel(fe, cs, pr, c, U, U1, U2, U3, assemble_matrix, M, b, s);
            scatter_cell(n, cs->cell_dofs + c * n, constrained, rowptr, col,
                         assemble_matrix, M, b, val, rhs);
          }
      }
    free(M);
    free(b);
    free(s);
  }
  return 0;
}

/* y = A x (what Epetra_CrsMatrix::Multiply does for AztecOO). */
void
glso_spmv(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val,
          const double *x, double *y)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    {
      double s = 0;
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
        s += val[k] * x[col[k]];
      y[i] = s;
    }
}

/* setup_ILU, source/solvers/gls_navier_stokes.cc:1161-1176 with fill = 0:
 * Ifpack "ILU" of the rank-local matrix, overlap 0.  Restated from Ifpack_ILU's
 * published algorithm: perturb the diagonal d <- rtol*d + sgn(d)*atol, then an
 * IKJ incomplete factorisation restricted to the matrix pattern.  `block_ptr`
 * ([nblock+1] row offsets) gives the rank-local diagonal blocks: entries whose
 * column is outside the row's block are ignored (block-Jacobi, what Ifpack does
 * with one MPI rank per block).  nblock = 1, block_ptr = {0,n} is the serial
 * case.  lu has the pattern of A; diag_pos[i] receives the diagonal position.
 * Returns 0, or 1+row of the first zero pivot.  */
int
glso_ilu0(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val,
          double atol, double rtol, int nblock, const int64_t *block_ptr,
          double *lu, int64_t *diag_pos)
{
  int status = 0;
  memcpy(lu, val, sizeof(double) * rowptr[n]);
#pragma omp parallel for schedule(dynamic, 1)
  for (int blk = 0; blk < nblock; ++blk)
    {
      const int64_t r0 = block_ptr[blk], r1 = block_ptr[blk + 1];
      for (int64_t i = r0; i < r1; ++i)
        {
          int64_t dp = csr_find(rowptr, col, (int32_t)i, (int32_t)i);
          diag_pos[i] = dp;
          double d    = lu[dp];
          lu[dp]      = rtol * d + (d >= 0 ? 1.0 : -1.0) * atol;
        }
      for (int64_t i = r0; i < r1; ++i)
        {
          for (int64_t kk = rowptr[i]; kk < diag_pos[i]; ++kk)
            {
              const int32_t k = col[kk];
              if (k < r0)
                continue; /* outside the diagonal block */
              const double piv = lu[diag_pos[k]];
              const double lik = lu[kk] / piv;
              lu[kk]           = lik;
              /* a_ij -= l_ik u_kj for j > k, (i,j) in pattern */
              int64_t p = kk + 1;
              for (int64_t q = diag_pos[k] + 1; q < rowptr[k + 1]; ++q)
                {
                  const int32_t j = col[q];
                  if (j >= r1)
                    break;
                  while (p < rowptr[i + 1] && col[p] < j)
                    ++p;
                  if (p == rowptr[i + 1])
                    break;
                  if (col[p] == j)
                    lu[p] -= lik * lu[q];
                }
            }
          if (lu[diag_pos[i]] == 0.0)
            {
#pragma omp critical
              if (!status)
                status = (int)(1 + i);
            }
        }
    }
  return status;
}

/* z = (LU)^-1 r on each diagonal block: unit-lower solve then upper solve. */
void
glso_ilu_apply(int64_t n, const int64_t *rowptr, const int32_t *col,
               const double *lu, const int64_t *diag_pos, int nblock,
               const int64_t *block_ptr, const double *r, double *z)
{
  (void)n;
#pragma omp parallel for schedule(dynamic, 1)
  for (int blk = 0; blk < nblock; ++blk)
    {
      const int64_t r0 = block_ptr[blk], r1 = block_ptr[blk + 1];
      for (int64_t i = r0; i < r1; ++i)
        {
          double s = r[i];
          for (int64_t k = rowptr[i]; k < diag_pos[i]; ++k)
            if (col[k] >= r0)
              s -= lu[k] * z[col[k]];
          z[i] = s;
        }
      for (int64_t i = r1 - 1; i >= r0; --i)
        {
          double s = z[i];
          for (int64_t k = diag_pos[i] + 1; k < rowptr[i + 1]; ++k)
            if (col[k] < r1)
              s -= lu[k] * z[col[k]];
          z[i] = s / lu[diag_pos[i]];
        }
    }
}

static double
dot(int64_t n, const double *a, const double *b)
{
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i)
    s += a[i] * b[i];
  return s;
}

/* solve_system_GMRES, source/solvers/gls_navier_stokes.cc:1242-1289:
 * TrilinosWrappers::SolverGMRES = AztecOO AZ_gmres, Krylov space `restart`
 * (deal.II default 30), right preconditioning with the ILU above, AZ_noscaled
 * convergence (||r||_2 < tol), classical Gram–Schmidt applied twice, zero
 * initial guess (:1261-1262).  Restated from the published algorithm; the
 * recurrence residual |g_{j+1}| decides convergence, the explicitly recomputed
 * ||b - A x||_2 is what SolverControl logs (tests/solvers/restart_01.output).
 * Returns 0 converged, 1 max_iters reached.  */
int
glso_gmres(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val,
           const double *lu, const int64_t *diag_pos, int nblock,
           const int64_t *block_ptr, const double *b, double tol, int max_iters,
           int restart, double *x, int *iters_out, double *true_res_out,
           double *res_hist /* [max_iters+1] or NULL */)
{
  const int m  = restart;
  double   *V  = (double *)malloc(sizeof(double) * (size_t)n * (m + 1));
  double   *w  = (double *)malloc(sizeof(double) * n);
  double   *z  = (double *)malloc(sizeof(double) * n);
  double   *H  = (double *)calloc((size_t)(m + 1) * m, sizeof(double));
  double   *cs = (double *)malloc(sizeof(double) * m);
  double   *sn = (double *)malloc(sizeof(double) * m);
  double   *g  = (double *)malloc(sizeof(double) * (m + 1));
  double   *y  = (double *)malloc(sizeof(double) * m);
  double   *hc = (double *)malloc(sizeof(double) * (m + 1));
  int       it = 0, status = 1;
  memset(x, 0, sizeof(double) * n);
  double beta = sqrt(dot(n, b, b));
  if (res_hist)
    res_hist[0] = beta;
  if (beta < tol)
    status = 0;
  while (status && it < max_iters)
    {
      /* r = b - A x  (x = 0 on the first cycle) */
      if (it == 0)
        memcpy(w, b, sizeof(double) * n);
      else
        {
          glso_spmv(n, rowptr, col, val, x, w);
          for (int64_t i = 0; i < n; ++i)
            w[i] = b[i] - w[i];
          beta = sqrt(dot(n, w, w));
        }
      for (int64_t i = 0; i < n; ++i)
        V[i] = w[i] / beta;
      memset(g, 0, sizeof(double) * (m + 1));
      g[0]  = beta;
      int j = 0;
      for (; j < m && it < max_iters; ++j)
        {
          double *vj = V + (size_t)j * n, *vn = V + (size_t)(j + 1) * n;
          glso_ilu_apply(n, rowptr, col, lu, diag_pos, nblock, block_ptr, vj, z);
          glso_spmv(n, rowptr, col, val, z, w);
          for (int i = 0; i <= j; ++i)
            H[i * m + j] = 0;
          for (int pass = 0; pass < 2; ++pass)
            {
              for (int i = 0; i <= j; ++i)
                hc[i] = dot(n, V + (size_t)i * n, w);
              for (int i = 0; i <= j; ++i)
                {
                  const double  hi = hc[i];
                  const double *vi = V + (size_t)i * n;
#pragma omp parallel for schedule(static)
                  for (int64_t t = 0; t < n; ++t)
                    w[t] -= hi * vi[t];
                  H[i * m + j] += hi;
                }
            }
          const double hn = sqrt(dot(n, w, w));
          H[(j + 1) * m + j] = hn;
          if (hn != 0)
            for (int64_t t = 0; t < n; ++t)
              vn[t] = w[t] / hn;
          for (int i = 0; i < j; ++i)
            {
              const double t  = cs[i] * H[i * m + j] + sn[i] * H[(i + 1) * m + j];
              H[(i + 1) * m + j] = -sn[i] * H[i * m + j] + cs[i] * H[(i + 1) * m + j];
              H[i * m + j]       = t;
            }
          const double a = H[j * m + j], bb = H[(j + 1) * m + j];
          const double rr = hypot(a, bb);
          cs[j]           = a / rr;
          sn[j]           = bb / rr;
          H[j * m + j]    = rr;
          H[(j + 1) * m + j] = 0;
          g[j + 1]        = -sn[j] * g[j];
          g[j]            = cs[j] * g[j];
          ++it;
          if (res_hist)
            res_hist[it] = fabs(g[j + 1]);
          if (fabs(g[j + 1]) < tol)
            {
              status = 0;
              ++j;
              break;
            }
        }
      /* x += M^-1 V y */
      for (int i = j - 1; i >= 0; --i)
        {
          double s = g[i];
          for (int k = i + 1; k < j; ++k)
            s -= H[i * m + k] * y[k];
          y[i] = s / H[i * m + i];
        }
      memset(w, 0, sizeof(double) * n);
      for (int i = 0; i < j; ++i)
        {
          const double *vi = V + (size_t)i * n;
          for (int64_t t = 0; t < n; ++t)
            w[t] += y[i] * vi[t];
        }
      glso_ilu_apply(n, rowptr, col, lu, diag_pos, nblock, block_ptr, w, z);
      for (int64_t t = 0; t < n; ++t)
        x[t] += z[t];
    }
  glso_spmv(n, rowptr, col, val, x, w);
  for (int64_t i = 0; i < n; ++i)
    w[i] = b[i] - w[i];
  *true_res_out = sqrt(dot(n, w, w));
  *iters_out    = it;
  free(V); free(w); free(z); free(H); free(cs); free(sn); free(g); free(y); free(hc);
  return status;
}

/* bdf_coefficients(p, dt[]), source/core/bdf.cc:24-75: divided differences on
 * the time table t_i = -sum_{j<i} dt_j.  out[p+1].  Pinned by
 * tests/core/bdf_01.output. */
static void
bdf_delta(int p, int n, int j, const double *times, double *out)
{
  if (j == 0)
    {
      for (int i = 0; i <= p; ++i)
        out[i] = 0;
      out[n] = 1;
      return;
    }
  double d1[8], d2[8];
  bdf_delta(p, n, j - 1, times, d1);
  bdf_delta(p, n + 1, j - 1, times, d2);
  for (int i = 0; i <= p; ++i)
    out[i] = (d1[i] - d2[i]) / (times[n] - times[n + j]);
}

void
glso_bdf_coefficients(int p, const double *dt, double *alpha)
{
  double times[8];
  for (int i = 0; i <= p; ++i)
    {
      times[i] = 0;
      for (int j = 0; j < i; ++j)
        times[i] -= dt[j];
    }
  for (int i = 0; i <= p; ++i)
    alpha[i] = 0;
  for (int j = 1; j <= p; ++j)
    {
      double factor = 1;
      for (int i = 1; i < j; ++i)
        factor *= times[0] - times[i];
      double term[8];
      bdf_delta(p, 0, j, times, term);
      for (int i = 0; i <= p; ++i)
        alpha[i] += factor * term[i];
    }
}

int
glso_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void
glso_set_num_threads(int t)
{
#ifdef _OPENMP
  omp_set_num_threads(t);
#else
  (void)t;
#endif
}
