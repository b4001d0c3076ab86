// Host-side stand-in for the deal.II objects the hot path consumes.
//
// In Lethe these come from GridGenerator::hyper_cube / subdivided_hyper_cube
// (source/core/grids.cc:12-78), FESystem(FE_Q(pu)^dim, FE_Q(pp)) + QGauss +
// FEValues (source/solvers/navier_stokes_base.cc:62,70,93-94;
// gls_navier_stokes.cc:244-252) and setup_dofs (gls_navier_stokes.cc:57-228:
// distribute_dofs, DoFRenumbering::Cuthill_McKee, boundary constraints, sparsity
// pattern with keep_constrained_dofs=false, owned/ghost index sets).  deal.II is
// not available in this image, so this file builds the same arrays for uniform box
// meshes and fills the glsns_fe_desc / glsns_mesh_desc a deal.II adapter would
// fill.  Pure host code, no CUDA.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/glsns.h"

namespace
{
  struct HostMesh
  {
    int    dim = 0, pu = 0, pp = 0, nq1 = 0;
    int    ncd[3] = {1, 1, 1};
    double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    // element tables
    int                 n_su = 0, n_sp = 0, n_q = 0, n_loc = 0;
    std::vector<double> Nu, dNu, d2Nu, Np, dNp, wq, xq;
    // dofs
    int64_t              n_dofs = 0, n_owned = 0, n_cells = 0, n_global = 0, owned_begin = 0;
    std::vector<int32_t> cell_dofs, col, color_ptr, color_cells, dof_comp;
    std::vector<int32_t> periodic_slave, periodic_master; // dofs identified by make_periodic
    std::vector<int64_t> rowptr, cell_ids, local_to_global;
    std::vector<uint8_t> constrained;
    std::vector<double>  cvalues, inv_jac, det_jac, measure, q_points, dof_coords;
    // halo
    std::vector<int32_t> neighbor_rank, send_idx;
    std::vector<int64_t> send_ptr, recv_ptr;
    std::string          error;
  };

  // ---- 1D Lagrange basis on equispaced nodes of [0,1] (FE_Q support points, p<=2
  //      coincide with Gauss-Lobatto) and Gauss-Legendre points ----
  void
  lagrange1d(int p, double x, double *v, double *d, double *d2)
  {
    std::vector<double> nodes(p + 1);
    for (int i = 0; i <= p; ++i)
      nodes[i] = p ? (double)i / p : 0.5;
    for (int a = 0; a <= p; ++a)
      {
        double den = 1;
        for (int b = 0; b <= p; ++b)
          if (b != a)
            den *= nodes[a] - nodes[b];
        double val = 1, der = 0, der2 = 0;
        for (int b = 0; b <= p; ++b)
          if (b != a)
            val *= x - nodes[b];
        for (int b = 0; b <= p; ++b)
          if (b != a)
            {
              double t = 1;
              for (int c = 0; c <= p; ++c)
                if (c != a && c != b)
                  t *= x - nodes[c];
              der += t;
              for (int c = 0; c <= p; ++c)
                if (c != a && c != b)
                  {
                    double s = 1;
                    for (int e = 0; e <= p; ++e)
                      if (e != a && e != b && e != c)
                        s *= x - nodes[e];
                    der2 += s;
                  }
            }
        v[a]  = val / den;
        d[a]  = der / den;
        d2[a] = der2 / den;
      }
  }

  void
  gauss01(int n, std::vector<double> &x, std::vector<double> &w)
  {
    x.resize(n), w.resize(n);
    for (int i = 0; i < n; ++i)
      {
        // Newton on P_n, start from the Chebyshev guess
        double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0;
        for (int it = 0; it < 100; ++it)
          {
            double p1 = 1, p2 = 0;
            for (int j = 1; j <= n; ++j)
              {
                const double p3 = p2;
                p2              = p1;
                p1              = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
              }
            pp               = n * (z * p1 - p2) / (z * z - 1.0);
            const double dz = p1 / pp;
            z -= dz;
            if (fabs(dz) < 1e-16)
              break;
          }
        x[n - 1 - i] = 0.5 * (z + 1.0);
        w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
      }
  }

  void
  tensor_tables(int dim, int p, const std::vector<double> &xq1, std::vector<double> &N,
                std::vector<double> &dN, std::vector<double> *d2N)
  {
    const int n1 = p + 1, q1 = (int)xq1.size();
    int       ns = 1, nq = 1;
    for (int d = 0; d < dim; ++d)
      ns *= n1, nq *= q1;
    std::vector<double> V(n1 * q1), D(n1 * q1), D2(n1 * q1), v(n1), dd(n1), d2(n1);
    for (int q = 0; q < q1; ++q)
      {
        lagrange1d(p, xq1[q], v.data(), dd.data(), d2.data());
        for (int a = 0; a < n1; ++a)
          V[a * q1 + q] = v[a], D[a * q1 + q] = dd[a], D2[a * q1 + q] = d2[a];
      }
    N.assign((size_t)nq * ns, 0), dN.assign((size_t)nq * ns * dim, 0);
    if (d2N)
      d2N->assign((size_t)nq * ns * dim * dim, 0);
    for (int q = 0; q < nq; ++q)
      for (int a = 0; a < ns; ++a)
        {
          int qi[3], ai[3], qq = q, aa = a;
          for (int d = 0; d < dim; ++d)
            qi[d] = qq % q1, qq /= q1, ai[d] = aa % n1, aa /= n1;
          double val = 1;
          for (int d = 0; d < dim; ++d)
            val *= V[ai[d] * q1 + qi[d]];
          N[(size_t)q * ns + a] = val;
          for (int d = 0; d < dim; ++d)
            {
              double g = 1;
              for (int e = 0; e < dim; ++e)
                g *= (e == d ? D : V)[ai[e] * q1 + qi[e]];
              dN[((size_t)q * ns + a) * dim + d] = g;
              if (d2N)
                for (int e = 0; e < dim; ++e)
                  {
                    double h = 1;
                    for (int f = 0; f < dim; ++f)
                      {
                        const std::vector<double> &T =
                          (d == e) ? (f == d ? D2 : V) : ((f == d || f == e) ? D : V);
                        h *= T[ai[f] * q1 + qi[f]];
                      }
                    (*d2N)[(((size_t)q * ns + a) * dim + d) * dim + e] = h;
                  }
            }
        }
  }

  // node-level helper for the structured grid
  struct Grid
  {
    int     dim, pu, gu[3];
    int64_t nnode;
    // neighbour index range of node coordinate i along direction d: every node that
    // shares a cell with it
    inline void
    range(int d, int i, int &a, int &b) const
    {
      if (i % pu == 0)
        a = std::max(0, i - pu), b = std::min(gu[d] - 1, i + pu);
      else
        a = (i / pu) * pu, b = a + pu;
    }
    inline void
    split(int64_t node, int idx[3]) const
    {
      idx[0] = (int)(node % gu[0]);
      idx[1] = (int)((node / gu[0]) % gu[1]);
      idx[2] = dim == 3 ? (int)(node / ((int64_t)gu[0] * gu[1])) : 0;
    }
    inline int64_t
    join(int i, int j, int k) const
    {
      return i + (int64_t)gu[0] * (j + (int64_t)gu[1] * k);
    }
  };

  // node order: natural or Cuthill-McKee on the node graph weighted by dofs per node
  // (every dof of a node has the same neighbours, so this equals CM on the dof graph,
  // DoFRenumbering::Cuthill_McKee at gls_navier_stokes.cc:70); node_order[position] = node
  void
  order_nodes(const Grid &G, const std::vector<int32_t> &ndof_node, const int renumber,
              std::vector<int64_t> &node_order)
  {
    const int dim = G.dim;
    node_order.resize(G.nnode);
  if (!renumber)
      std::iota(node_order.begin(), node_order.end(), (int64_t)0);
    else
      {
        std::vector<int64_t> degree(G.nnode);
#pragma omp parallel for schedule(static)
        for (int64_t v = 0; v < G.nnode; ++v)
          {
            int idx[3], a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
            G.split(v, idx);
            for (int d = 0; d < dim; ++d)
              G.range(d, idx[d], a[d], b[d]);
            int64_t deg = 0;
            for (int k = a[2]; k <= b[2]; ++k)
              for (int j = a[1]; j <= b[1]; ++j)
                for (int i = a[0]; i <= b[0]; ++i)
                  deg += ndof_node[G.join(i, j, k)];
            degree[v] = deg;
          }
        std::vector<uint8_t> seen(G.nnode, 0);
        int64_t              N = 0;
        // seed: lowest-index node of minimal degree (a corner), as a stable argsort gives
        std::vector<int64_t> by_deg(G.nnode);
        std::iota(by_deg.begin(), by_deg.end(), (int64_t)0);
        std::stable_sort(by_deg.begin(), by_deg.end(),
                         [&](int64_t x, int64_t y) { return degree[x] < degree[y]; });
        std::vector<int64_t> fresh;
        for (int64_t z = 0; z < G.nnode && N < G.nnode; ++z)
          {
            if (seen[by_deg[z]])
              continue;
            node_order[N++]  = by_deg[z];
            seen[by_deg[z]]  = 1;
            int64_t level_lo = N - 1;
            while (level_lo < N)
              {
                const int64_t level_hi = N;
                for (int64_t t = level_lo; t < level_hi; ++t)
                  {
                    const int64_t v = node_order[t];
                    int           idx[3], a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
                    G.split(v, idx);
                    for (int d = 0; d < dim; ++d)
                      G.range(d, idx[d], a[d], b[d]);
                    fresh.clear();
                    for (int k = a[2]; k <= b[2]; ++k)
                      for (int j = a[1]; j <= b[1]; ++j)
                        for (int i = a[0]; i <= b[0]; ++i)
                          {
                            const int64_t u = G.join(i, j, k);
                            if (!seen[u])
                              seen[u] = 1, fresh.push_back(u);
                          }
                    // neighbours of one parent by increasing degree, ties in index order
                    std::stable_sort(fresh.begin(), fresh.end(), [&](int64_t x, int64_t y) {
                      return degree[x] < degree[y];
                    });
                    for (int64_t u : fresh)
                      node_order[N++] = u;
                  }
                level_lo = level_hi;
              }
          }
      }
  }
} // namespace

extern "C" {

typedef struct glsnsh_mesh glsnsh_mesh; // opaque: HostMesh

// bc_type[2*dim]: 0 none, 1 noslip, 2 constant Dirichlet value bc_value[face][3]
// (face ids as deal.II colorize: 0 x=lo, 1 x=hi, 2 y=lo, 3 y=hi, 4 z=lo, 5 z=hi);
// bc_order[2*dim]: faces in the order the constraints are created (first wins).
// renumber: 0 none, 1 Cuthill-McKee (gls_navier_stokes.cc:70).
glsnsh_mesh *
glsnsh_mesh_create(int dim, const int *n_cells_dir, int pu, int pp, const double *lo,
                   const double *hi, int nq1, const int *bc_type, const double *bc_value,
                   const int *bc_order, int renumber, int with_q_points)
{
  HostMesh *M = new HostMesh();
  if ((dim != 2 && dim != 3) || pu < 1 || pp < 1 || pu % pp != 0)
    {
      M->error = "bad dim / degrees";
      return (glsnsh_mesh *)M;
    }
  M->dim = dim, M->pu = pu, M->pp = pp, M->nq1 = nq1 > 0 ? nq1 : pu + 1;
  for (int d = 0; d < dim; ++d)
    M->ncd[d] = n_cells_dir[d], M->lo[d] = lo[d], M->hi[d] = hi[d];
  // ---- element ----
  std::vector<double> x1, w1;
  gauss01(M->nq1, x1, w1);
  tensor_tables(dim, pu, x1, M->Nu, M->dNu, &M->d2Nu);
  tensor_tables(dim, pp, x1, M->Np, M->dNp, nullptr);
  M->n_su = 1, M->n_sp = 1, M->n_q = 1;
  for (int d = 0; d < dim; ++d)
    M->n_su *= pu + 1, M->n_sp *= pp + 1, M->n_q *= M->nq1;
  M->n_loc = dim * M->n_su + M->n_sp;
  M->wq.assign(M->n_q, 1.0), M->xq.assign((size_t)M->n_q * dim, 0.0);
  for (int q = 0; q < M->n_q; ++q)
    {
      int qq = q;
      for (int d = 0; d < dim; ++d)
        {
          const int i = qq % M->nq1;
          qq /= M->nq1;
          M->wq[q] *= w1[i];
          M->xq[(size_t)q * dim + d] = x1[i];
        }
    }
  // ---- nodes and provisional dof numbering (node-major, components interleaved) ----
  Grid G;
  G.dim = dim, G.pu = pu;
  G.gu[0] = G.gu[1] = G.gu[2] = 1;
  G.nnode = 1;
  for (int d = 0; d < dim; ++d)
    G.gu[d] = pu * M->ncd[d] + 1, G.nnode *= G.gu[d];
  const int ratio = pu / pp;
  auto      has_p = [&](const int idx[3]) {
    for (int d = 0; d < dim; ++d)
      if (idx[d] % ratio)
        return false;
    return true;
  };
  // node order: natural or Cuthill-McKee on the node graph weighted by dofs per node
  // (every dof of a node has the same neighbours, so this equals CM on the dof graph)
  std::vector<int64_t> node_order; // node_order[position] = node
  std::vector<int32_t> ndof_node(G.nnode);
  for (int64_t v = 0; v < G.nnode; ++v)
    {
      int idx[3];
      G.split(v, idx);
      ndof_node[v] = dim + (has_p(idx) ? 1 : 0);
    }
  order_nodes(G, ndof_node, renumber, node_order);
  std::vector<int64_t> first_dof(G.nnode + 1, 0); // first new dof id of a node
  {
    int64_t next = 0;
    for (int64_t t = 0; t < G.nnode; ++t)
      {
        first_dof[node_order[t]] = next;
        next += ndof_node[node_order[t]];
      }
    M->n_dofs = M->n_owned = M->n_global = next;
  }
  if (M->n_dofs >= (int64_t)INT32_MAX)
    {
      M->error = "more than 2^31 dofs";
      return (glsnsh_mesh *)M;
    }
  const double hx[3] = {(M->hi[0] - M->lo[0]) / M->ncd[0], (M->hi[1] - M->lo[1]) / M->ncd[1],
                        dim == 3 ? (M->hi[2] - M->lo[2]) / M->ncd[2] : 1.0};
  // ---- dof meta + boundary constraints ----
  M->constrained.assign(M->n_dofs, 0), M->cvalues.assign(M->n_dofs, 0.0);
  M->dof_comp.assign(M->n_dofs, 0), M->dof_coords.assign((size_t)M->n_dofs * dim, 0.0);
  M->local_to_global.resize(M->n_dofs);
  std::iota(M->local_to_global.begin(), M->local_to_global.end(), (int64_t)0);
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < G.nnode; ++v)
    {
      int idx[3];
      G.split(v, idx);
      for (int c = 0; c < ndof_node[v]; ++c)
        {
          const int64_t g = first_dof[v] + c;
          M->dof_comp[g]  = c;
          for (int d = 0; d < dim; ++d)
            M->dof_coords[(size_t)g * dim + d] = M->lo[d] + idx[d] * hx[d] / pu;
        }
      for (int f = 0; f < 2 * dim; ++f)
        {
          const int face = bc_order ? bc_order[f] : f;
          if (!bc_type || bc_type[face] == 0)
            continue;
          const int d = face / 2, side = face % 2;
          if (idx[d] != (side ? G.gu[d] - 1 : 0))
            continue;
          bool any = false;
          for (int c = 0; c < dim; ++c)
            {
              const int64_t g = first_dof[v] + c;
              if (!M->constrained[g])
                {
                  M->constrained[g] = 1;
                  M->cvalues[g]     = bc_type[face] == 2 ? bc_value[face * 3 + c] : 0.0;
                  any               = true;
                }
            }
          if (any)
            break; // first listed boundary wins on edges and corners
        }
    }
  // ---- cells ----
  M->n_cells = 1;
  for (int d = 0; d < dim; ++d)
    M->n_cells *= M->ncd[d];
  M->cell_dofs.resize((size_t)M->n_cells * M->n_loc);
  M->cell_ids.resize(M->n_cells);
  M->inv_jac.assign((size_t)M->n_cells * dim * dim, 0.0);
  M->det_jac.resize(M->n_cells), M->measure.resize(M->n_cells);
  if (with_q_points)
    M->q_points.resize((size_t)M->n_cells * M->n_q * dim);
  std::vector<int32_t>              color(M->n_cells);
  const int                         nu1 = pu + 1, np1 = pp + 1;
  double                            vol = 1;
  for (int d = 0; d < dim; ++d)
    vol *= hx[d];
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < M->n_cells; ++c)
    {
      int ci[3] = {(int)(c % M->ncd[0]), (int)((c / M->ncd[0]) % M->ncd[1]),
                   dim == 3 ? (int)(c / ((int64_t)M->ncd[0] * M->ncd[1])) : 0};
      M->cell_ids[c] = c;
      int32_t *cd    = M->cell_dofs.data() + (size_t)c * M->n_loc;
      for (int a = 0; a < M->n_su; ++a)
        {
          int ai[3] = {a % nu1, (a / nu1) % nu1, dim == 3 ? a / (nu1 * nu1) : 0};
          const int64_t v =
            G.join(ci[0] * pu + ai[0], ci[1] * pu + ai[1], dim == 3 ? ci[2] * pu + ai[2] : 0);
          for (int comp = 0; comp < dim; ++comp)
            cd[comp * M->n_su + a] = (int32_t)(first_dof[v] + comp);
        }
      for (int a = 0; a < M->n_sp; ++a)
        {
          int ai[3] = {a % np1, (a / np1) % np1, dim == 3 ? a / (np1 * np1) : 0};
          const int64_t v = G.join(ci[0] * pu + ai[0] * ratio, ci[1] * pu + ai[1] * ratio,
                                   dim == 3 ? ci[2] * pu + ai[2] * ratio : 0);
          cd[dim * M->n_su + a] = (int32_t)(first_dof[v] + dim);
        }
      for (int d = 0; d < dim; ++d)
        M->inv_jac[(size_t)c * dim * dim + d * dim + d] = 1.0 / hx[d];
      M->det_jac[c] = M->measure[c] = vol;
      if (with_q_points)
        for (int q = 0; q < M->n_q; ++q)
          for (int d = 0; d < dim; ++d)
            M->q_points[((size_t)c * M->n_q + q) * dim + d] =
              M->lo[d] + (ci[d] + M->xq[(size_t)q * dim + d]) * hx[d];
      color[c] = (ci[0] & 1) + 2 * (ci[1] & 1) + (dim == 3 ? 4 * (ci[2] & 1) : 0);
    }
  const int ncolor = 1 << dim;
  M->color_ptr.assign(ncolor + 1, 0);
  for (int64_t c = 0; c < M->n_cells; ++c)
    M->color_ptr[color[c] + 1]++;
  for (int k = 0; k < ncolor; ++k)
    M->color_ptr[k + 1] += M->color_ptr[k];
  M->color_cells.resize(M->n_cells);
  {
    std::vector<int32_t> pos(M->color_ptr.begin(), M->color_ptr.end() - 1);
    for (int64_t c = 0; c < M->n_cells; ++c)
      M->color_cells[pos[color[c]]++] = (int32_t)c;
  }
  // ---- sparsity: couplings between unconstrained dofs sharing a cell + the diagonal ----
  std::vector<int64_t> node_of_dof(M->n_dofs);
  for (int64_t v = 0; v < G.nnode; ++v)
    for (int c = 0; c < ndof_node[v]; ++c)
      node_of_dof[first_dof[v] + c] = v;
  M->rowptr.assign(M->n_dofs + 1, 0);
  auto row_entries = [&](int64_t r, int32_t *out) -> int64_t {
    if (M->constrained[r])
      {
        if (out)
          out[0] = (int32_t)r;
        return 1;
      }
    int idx[3], a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
    G.split(node_of_dof[r], idx);
    for (int d = 0; d < dim; ++d)
      G.range(d, idx[d], a[d], b[d]);
    int64_t cnt = 0;
    for (int k = a[2]; k <= b[2]; ++k)
      for (int j = a[1]; j <= b[1]; ++j)
        for (int i = a[0]; i <= b[0]; ++i)
          {
            const int64_t u = G.join(i, j, k);
            for (int c = 0; c < ndof_node[u]; ++c)
              {
                const int64_t g = first_dof[u] + c;
                if (M->constrained[g])
                  continue;
                if (out)
                  out[cnt] = (int32_t)g;
                ++cnt;
              }
          }
    if (out)
      std::sort(out, out + cnt);
    return cnt;
  };
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < M->n_dofs; ++r)
    M->rowptr[r + 1] = row_entries(r, nullptr);
  for (int64_t r = 0; r < M->n_dofs; ++r)
    M->rowptr[r + 1] += M->rowptr[r];
  M->col.resize((size_t)M->rowptr[M->n_dofs]);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t r = 0; r < M->n_dofs; ++r)
    row_entries(r, M->col.data() + M->rowptr[r]);
  return (glsnsh_mesh *)M;
}

// `type = periodic` boundary pairs (boundary_conditions.h; DoFTools::make_periodicity_constraints
// in setup_dofs): bit d of dir_mask identifies the dofs of face 2d+1 with those of face 2d.  The
// cells then refer to the lo-face dofs directly -- the same global system as resolving x_hi = x_lo
// in distribute_local_to_global -- while the hi-face dofs stay in the numbering (deal.II counts
// them too) as constrained rows no cell touches (the device assembly gives those a unit diagonal).
// Cell colours and the sparsity pattern are rebuilt.  Serial meshes only; returns 0 or 1 (+ error).
int
glsnsh_mesh_make_periodic(glsnsh_mesh *m, int dir_mask)
{
  HostMesh *M = (HostMesh *)m;
  if (!M->error.empty())
    return 1;
  if (M->n_owned != M->n_dofs || !M->periodic_slave.empty())
    {
      M->error = "make_periodic: serial, not yet periodic meshes only";
      return 1;
    }
  const int dim = M->dim, pu = M->pu;
  int       gu[3] = {1, 1, 1};
  double    hn[3] = {1, 1, 1};
  for (int d = 0; d < dim; ++d)
    gu[d] = pu * M->ncd[d] + 1, hn[d] = (M->hi[d] - M->lo[d]) / M->ncd[d] / pu;
  const int64_t nnode = (int64_t)gu[0] * gu[1] * gu[2];
  // dof of (node, component)
  std::vector<int32_t> dof_at((size_t)nnode * (dim + 1), -1);
  std::vector<int64_t> node_of(M->n_dofs);
  for (int64_t g = 0; g < M->n_dofs; ++g)
    {
      int64_t v = 0, stride = 1;
      for (int d = 0; d < dim; ++d)
        {
          const int i = (int)llround((M->dof_coords[(size_t)g * dim + d] - M->lo[d]) / hn[d]);
          v += stride * i, stride *= gu[d];
        }
      node_of[g]                                     = v;
      dof_at[(size_t)v * (dim + 1) + M->dof_comp[g]] = (int32_t)g;
    }
  std::vector<int32_t> master_of(M->n_dofs);
  for (int64_t g = 0; g < M->n_dofs; ++g)
    {
      int64_t v = node_of[g], stride = 1, mv = v;
      for (int d = 0; d < dim; ++d)
        {
          const int i = (int)((v / stride) % gu[d]);
          if (((dir_mask >> d) & 1) && i == gu[d] - 1)
            mv -= stride * i;
          stride *= gu[d];
        }
      master_of[g] = dof_at[(size_t)mv * (dim + 1) + M->dof_comp[g]];
      if (master_of[g] != (int32_t)g)
        {
          M->periodic_slave.push_back((int32_t)g), M->periodic_master.push_back(master_of[g]);
          M->constrained[g] = 1, M->cvalues[g] = 0.0;
        }
    }
  for (int32_t &g : M->cell_dofs)
    g = master_of[g];
  // colours: parity per direction; across a periodic wrap with an odd cell count the last cell
  // of the direction takes a third colour
  int radix = 2;
  for (int d = 0; d < dim; ++d)
    if (((dir_mask >> d) & 1) && (M->ncd[d] & 1))
      radix = 3;
  int ncolor = 1;
  for (int d = 0; d < dim; ++d)
    ncolor *= radix;
  std::vector<int32_t> color(M->n_cells);
  for (int64_t c = 0; c < M->n_cells; ++c)
    {
      const int ci[3] = {(int)(c % M->ncd[0]), (int)((c / M->ncd[0]) % M->ncd[1]),
                         dim == 3 ? (int)(c / ((int64_t)M->ncd[0] * M->ncd[1])) : 0};
      int       col = 0, mul = 1;
      for (int d = 0; d < dim; ++d)
        {
          int k = ci[d] & 1;
          if (((dir_mask >> d) & 1) && (M->ncd[d] & 1) && ci[d] == M->ncd[d] - 1)
            k = 2;
          col += mul * k, mul *= radix;
        }
      color[c] = col;
    }
  M->color_ptr.assign(ncolor + 1, 0);
  for (int64_t c = 0; c < M->n_cells; ++c)
    M->color_ptr[color[c] + 1]++;
  for (int k = 0; k < ncolor; ++k)
    M->color_ptr[k + 1] += M->color_ptr[k];
  {
    std::vector<int32_t> pos(M->color_ptr.begin(), M->color_ptr.end() - 1);
    for (int64_t c = 0; c < M->n_cells; ++c)
      M->color_cells[pos[color[c]]++] = (int32_t)c;
  }
  // sparsity from the cell -> dof table: couplings between unconstrained dofs of a cell + diagonal
  std::vector<int64_t> cptr(M->n_dofs + 1, 0);
  for (int32_t g : M->cell_dofs)
    cptr[g + 1]++;
  for (int64_t g = 0; g < M->n_dofs; ++g)
    cptr[g + 1] += cptr[g];
  std::vector<int32_t> cells_of((size_t)cptr[M->n_dofs]);
  {
    std::vector<int64_t> pos(cptr.begin(), cptr.end() - 1);
    for (int64_t c = 0; c < M->n_cells; ++c)
      for (int k = 0; k < M->n_loc; ++k)
        cells_of[(size_t)pos[M->cell_dofs[(size_t)c * M->n_loc + k]]++] = (int32_t)c;
  }
  std::vector<std::vector<int32_t>> rows((size_t)M->n_dofs);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t r = 0; r < M->n_dofs; ++r)
    {
      std::vector<int32_t> &out = rows[(size_t)r];
      out.push_back((int32_t)r);
      if (!M->constrained[r])
        for (int64_t t = cptr[r]; t < cptr[r + 1]; ++t)
          {
            const int32_t *cd = M->cell_dofs.data() + (size_t)cells_of[(size_t)t] * M->n_loc;
            for (int k = 0; k < M->n_loc; ++k)
              if (!M->constrained[cd[k]])
                out.push_back(cd[k]);
          }
      std::sort(out.begin(), out.end());
      out.erase(std::unique(out.begin(), out.end()), out.end());
    }
  M->rowptr.assign(M->n_dofs + 1, 0);
  for (int64_t r = 0; r < M->n_dofs; ++r)
    M->rowptr[r + 1] = M->rowptr[r] + (int64_t)rows[(size_t)r].size();
  M->col.resize((size_t)M->rowptr[M->n_dofs]);
  for (int64_t r = 0; r < M->n_dofs; ++r)
    std::copy(rows[(size_t)r].begin(), rows[(size_t)r].end(), M->col.begin() + M->rowptr[r]);
  return 0;
}

// The rank-local view of a (serial) mesh: contiguous blocks of the global dof order
// balanced by nonzeros, local numbering owned-first then ghosts by global id (which
// groups them by owner), every cell that touches an owned row.
glsnsh_mesh *
glsnsh_mesh_partition(const glsnsh_mesh *serial, int n_ranks, int rank)
{
  const HostMesh *S = (const HostMesh *)serial;
  HostMesh       *M = new HostMesh();
  if (!S->error.empty() || n_ranks < 1 || rank < 0 || rank >= n_ranks || S->n_owned != S->n_dofs)
    {
      M->error = "bad partition request";
      return (glsnsh_mesh *)M;
    }
  M->dim = S->dim, M->pu = S->pu, M->pp = S->pp, M->nq1 = S->nq1;
  memcpy(M->ncd, S->ncd, sizeof(M->ncd)), memcpy(M->lo, S->lo, sizeof(M->lo)),
    memcpy(M->hi, S->hi, sizeof(M->hi));
  M->n_su = S->n_su, M->n_sp = S->n_sp, M->n_q = S->n_q, M->n_loc = S->n_loc;
  M->Nu = S->Nu, M->dNu = S->dNu, M->d2Nu = S->d2Nu, M->Np = S->Np, M->dNp = S->dNp;
  M->wq = S->wq, M->xq = S->xq;
  M->n_global = S->n_dofs;
  // block boundaries: equal share of the nonzeros
  std::vector<int64_t> begin(n_ranks + 1, 0);
  const int64_t        nnz = S->rowptr[S->n_dofs];
  for (int r = 1; r < n_ranks; ++r)
    {
      const int64_t target = nnz / n_ranks * r;
      begin[r] = std::lower_bound(S->rowptr.begin(), S->rowptr.end(), target) - S->rowptr.begin();
      begin[r] = std::min<int64_t>(std::max(begin[r], begin[r - 1]), S->n_dofs);
    }
  begin[n_ranks] = S->n_dofs;
  const int64_t r0 = begin[rank], r1 = begin[rank + 1];
  M->owned_begin = r0;
  M->n_owned     = r1 - r0;
  // cells touching an owned dof; ghost dofs = their other dofs
  std::vector<int64_t> ghosts;
  for (int64_t c = 0; c < S->n_cells; ++c)
    {
      const int32_t *cd = S->cell_dofs.data() + (size_t)c * S->n_loc;
      bool           touches = false;
      for (int k = 0; k < S->n_loc && !touches; ++k)
        touches = cd[k] >= r0 && cd[k] < r1;
      if (!touches)
        continue;
      M->cell_ids.push_back(c);
      for (int k = 0; k < S->n_loc; ++k)
        if (cd[k] < r0 || cd[k] >= r1)
          ghosts.push_back(cd[k]);
    }
  // owned rows may also couple to dofs of cells listed above only; all are covered
  std::sort(ghosts.begin(), ghosts.end());
  ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
  M->n_cells = (int64_t)M->cell_ids.size();
  M->n_dofs  = M->n_owned + (int64_t)ghosts.size();
  auto to_local = [&](int64_t g) -> int32_t {
    if (g >= r0 && g < r1)
      return (int32_t)(g - r0);
    return (int32_t)(M->n_owned +
                     (std::lower_bound(ghosts.begin(), ghosts.end(), g) - ghosts.begin()));
  };
  M->local_to_global.resize(M->n_dofs);
  for (int64_t i = 0; i < M->n_owned; ++i)
    M->local_to_global[i] = r0 + i;
  for (size_t i = 0; i < ghosts.size(); ++i)
    M->local_to_global[M->n_owned + i] = ghosts[i];
  M->constrained.resize(M->n_dofs), M->cvalues.resize(M->n_dofs), M->dof_comp.resize(M->n_dofs);
  M->dof_coords.resize((size_t)M->n_dofs * M->dim);
  for (int64_t i = 0; i < M->n_dofs; ++i)
    {
      const int64_t g   = M->local_to_global[i];
      M->constrained[i] = S->constrained[g], M->cvalues[i] = S->cvalues[g];
      M->dof_comp[i]    = S->dof_comp[g];
      for (int d = 0; d < M->dim; ++d)
        M->dof_coords[(size_t)i * M->dim + d] = S->dof_coords[(size_t)g * M->dim + d];
    }
  // cells
  const int dim = M->dim;
  M->cell_dofs.resize((size_t)M->n_cells * M->n_loc);
  M->inv_jac.resize((size_t)M->n_cells * dim * dim), M->det_jac.resize(M->n_cells);
  M->measure.resize(M->n_cells);
  if (!S->q_points.empty())
    M->q_points.resize((size_t)M->n_cells * M->n_q * dim);
  std::vector<int32_t> color(M->n_cells);
  std::vector<int32_t> color_of_serial(S->n_cells);
  for (size_t k = 0; k + 1 < S->color_ptr.size(); ++k)
    for (int32_t t = S->color_ptr[k]; t < S->color_ptr[k + 1]; ++t)
      color_of_serial[S->color_cells[t]] = (int32_t)k;
  for (int64_t lc = 0; lc < M->n_cells; ++lc)
    {
      const int64_t c = M->cell_ids[lc];
      for (int k = 0; k < M->n_loc; ++k)
        M->cell_dofs[(size_t)lc * M->n_loc + k] = to_local(S->cell_dofs[(size_t)c * S->n_loc + k]);
      memcpy(&M->inv_jac[(size_t)lc * dim * dim], &S->inv_jac[(size_t)c * dim * dim],
             sizeof(double) * dim * dim);
      M->det_jac[lc] = S->det_jac[c], M->measure[lc] = S->measure[c];
      if (!S->q_points.empty())
        memcpy(&M->q_points[(size_t)lc * M->n_q * dim], &S->q_points[(size_t)c * M->n_q * dim],
               sizeof(double) * M->n_q * dim);
      color[lc] = color_of_serial[c];
    }
  const int ncolor = (int)S->color_ptr.size() - 1;
  M->color_ptr.assign(ncolor + 1, 0);
  for (int64_t c = 0; c < M->n_cells; ++c)
    M->color_ptr[color[c] + 1]++;
  for (int k = 0; k < ncolor; ++k)
    M->color_ptr[k + 1] += M->color_ptr[k];
  M->color_cells.resize(M->n_cells);
  {
    std::vector<int32_t> pos(M->color_ptr.begin(), M->color_ptr.end() - 1);
    for (int64_t c = 0; c < M->n_cells; ++c)
      M->color_cells[pos[color[c]]++] = (int32_t)c;
  }
  // local CSR of the owned rows
  M->rowptr.resize(M->n_owned + 1);
  for (int64_t i = 0; i <= M->n_owned; ++i)
    M->rowptr[i] = S->rowptr[r0 + i] - S->rowptr[r0];
  M->col.resize((size_t)M->rowptr[M->n_owned]);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < M->n_owned; ++i)
    {
      int32_t *out = M->col.data() + M->rowptr[i];
      int64_t  n   = 0;
      for (int64_t k = S->rowptr[r0 + i]; k < S->rowptr[r0 + i + 1]; ++k)
        out[n++] = to_local(S->col[k]);
      std::sort(out, out + n);
    }
  // halo lists: what each neighbour needs from us is what we would need from it by
  // symmetry of the cell coupling; computed without communication from the serial mesh
  std::map<int, std::vector<int64_t>> recv_from, send_to;
  auto                                 owner = [&](int64_t g) {
    return (int)(std::upper_bound(begin.begin(), begin.end(), g) - begin.begin()) - 1;
  };
  for (int64_t g : ghosts)
    recv_from[owner(g)].push_back(g);
  for (int64_t c = 0; c < S->n_cells; ++c)
    {
      const int32_t *cd = S->cell_dofs.data() + (size_t)c * S->n_loc;
      bool           mine = false;
      for (int k = 0; k < S->n_loc && !mine; ++k)
        mine = cd[k] >= r0 && cd[k] < r1;
      if (!mine)
        continue;
      // every other rank owning a dof of this cell assembles it too and needs our dofs
      for (int k = 0; k < S->n_loc; ++k)
        {
          const int o = owner(cd[k]);
          if (o == rank)
            continue;
          for (int l = 0; l < S->n_loc; ++l)
            if (cd[l] >= r0 && cd[l] < r1)
              send_to[o].push_back(cd[l]);
        }
    }
  std::vector<int> nb;
  for (auto &kv : recv_from)
    nb.push_back(kv.first);
  for (auto &kv : send_to)
    nb.push_back(kv.first);
  std::sort(nb.begin(), nb.end());
  nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
  M->send_ptr.assign(1, 0), M->recv_ptr.assign(1, 0);
  for (int o : nb)
    {
      M->neighbor_rank.push_back(o);
      std::vector<int64_t> &s = send_to[o];
      std::sort(s.begin(), s.end());
      s.erase(std::unique(s.begin(), s.end()), s.end());
      for (int64_t g : s)
        M->send_idx.push_back((int32_t)(g - r0));
      M->send_ptr.push_back((int64_t)M->send_idx.size());
      M->recv_ptr.push_back(M->recv_ptr.back() + (int64_t)recv_from[o].size());
    }
  return (glsnsh_mesh *)M;
}

// The rank-local mesh built directly, without ever holding the global dof-level arrays
// (BASELINE.json configs[3]: 116^3 cells, 50.6 M dofs, 1.3e10 non-zeros -- the global CSR is
// 52 GB of column indices): the same arrays, entry for entry, as glsnsh_mesh_create followed by
// glsnsh_mesh_partition (tests/test_host_mirror.py holds the comparison), from node-level global
// data only -- the Cuthill-McKee order of the node graph (12.6 M nodes at 116^3), the first dof of
// every node, and row lengths per node (every dof of a node has the same row).
glsnsh_mesh *
glsnsh_mesh_create_local(int dim, const int *n_cells_dir, int pu, int pp, const double *lo,
                         const double *hi, int nq1, const int *bc_type, const double *bc_value,
                         const int *bc_order, int renumber, int with_q_points, int n_ranks, int rank)
{
  HostMesh *M = new HostMesh();
  if ((dim != 2 && dim != 3) || pu < 1 || pp < 1 || pu % pp != 0 || n_ranks < 1 || rank < 0 ||
      rank >= n_ranks)
    {
      M->error = "bad dim / degrees / rank";
      return (glsnsh_mesh *)M;
    }
  M->dim = dim, M->pu = pu, M->pp = pp, M->nq1 = nq1 > 0 ? nq1 : pu + 1;
  for (int d = 0; d < dim; ++d)
    M->ncd[d] = n_cells_dir[d], M->lo[d] = lo[d], M->hi[d] = hi[d];
  std::vector<double> x1, w1;
  gauss01(M->nq1, x1, w1);
  tensor_tables(dim, pu, x1, M->Nu, M->dNu, &M->d2Nu);
  tensor_tables(dim, pp, x1, M->Np, M->dNp, nullptr);
  M->n_su = 1, M->n_sp = 1, M->n_q = 1;
  for (int d = 0; d < dim; ++d)
    M->n_su *= pu + 1, M->n_sp *= pp + 1, M->n_q *= M->nq1;
  M->n_loc = dim * M->n_su + M->n_sp;
  M->wq.assign(M->n_q, 1.0), M->xq.assign((size_t)M->n_q * dim, 0.0);
  for (int q = 0; q < M->n_q; ++q)
    {
      int qq = q;
      for (int d = 0; d < dim; ++d)
        {
          const int i = qq % M->nq1;
          qq /= M->nq1;
          M->wq[q] *= w1[i];
          M->xq[(size_t)q * dim + d] = x1[i];
        }
    }
  Grid G;
  G.dim = dim, G.pu = pu;
  G.gu[0] = G.gu[1] = G.gu[2] = 1;
  G.nnode = 1;
  for (int d = 0; d < dim; ++d)
    G.gu[d] = pu * M->ncd[d] + 1, G.nnode *= G.gu[d];
  const int ratio = pu / pp;
  auto      has_p = [&](const int idx[3]) {
    for (int d = 0; d < dim; ++d)
      if (idx[d] % ratio)
        return false;
    return true;
  };
  // the boundary the node's velocity dofs are constrained by (first listed wins), or -1
  auto bc_face = [&](const int idx[3]) -> int {
    for (int f = 0; f < 2 * dim; ++f)
      {
        const int face = bc_order ? bc_order[f] : f;
        if (!bc_type || bc_type[face] == 0)
          continue;
        const int d = face / 2, side = face % 2;
        if (idx[d] == (side ? G.gu[d] - 1 : 0))
          return face;
      }
    return -1;
  };
  std::vector<int64_t> node_order;
  std::vector<int32_t> ndof_node(G.nnode);
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < G.nnode; ++v)
    {
      int idx[3];
      G.split(v, idx);
      ndof_node[v] = dim + (has_p(idx) ? 1 : 0);
    }
  order_nodes(G, ndof_node, renumber, node_order);
  // position -> first dof (monotone), node -> first dof
  std::vector<int64_t> pos_first(G.nnode + 1), first_dof(G.nnode);
  {
    int64_t next = 0;
    for (int64_t t = 0; t < G.nnode; ++t)
      {
        pos_first[t]             = next;
        first_dof[node_order[t]] = next;
        next += ndof_node[node_order[t]];
      }
    pos_first[G.nnode] = next;
    M->n_global        = next;
  }
  if (M->n_global >= (int64_t)INT32_MAX)
    {
      M->error = "more than 2^31 dofs";
      return (glsnsh_mesh *)M;
    }
  // unconstrained dofs of every node, and of its neighbourhood (= the row length of its
  // unconstrained dofs)
  std::vector<int8_t> n_free(G.nnode);
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < G.nnode; ++v)
    {
      int idx[3];
      G.split(v, idx);
      n_free[v] = (int8_t)(ndof_node[v] - (bc_face(idx) >= 0 ? dim : 0));
    }
  auto row_len_free = [&](int64_t v) -> int64_t {
    int idx[3], a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
    G.split(v, idx);
    for (int d = 0; d < dim; ++d)
      G.range(d, idx[d], a[d], b[d]);
    int64_t cnt = 0;
    for (int k = a[2]; k <= b[2]; ++k)
      for (int j = a[1]; j <= b[1]; ++j)
        for (int i = a[0]; i <= b[0]; ++i)
          cnt += n_free[G.join(i, j, k)];
    return cnt;
  };
  // non-zeros before every position of the order
  std::vector<int64_t> pos_nnz(G.nnode + 1);
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < G.nnode; ++t)
    {
      const int64_t v  = node_order[t];
      const int     nf = n_free[v], nc = ndof_node[v] - nf;
      pos_nnz[t + 1]   = nc + (nf ? nf * row_len_free(v) : 0);
    }
  pos_nnz[0] = 0;
  for (int64_t t = 0; t < G.nnode; ++t)
    pos_nnz[t + 1] += pos_nnz[t];
  const int64_t nnz_global = pos_nnz[G.nnode];
  // rowptr of a global dof (constrained dofs of a node come first: velocity components)
  auto rowptr_of = [&](int64_t g) -> int64_t {
    const int64_t t = std::upper_bound(pos_first.begin(), pos_first.end(), g) - pos_first.begin() - 1;
    if (t >= G.nnode)
      return nnz_global;
    const int64_t v  = node_order[t];
    const int     nf = n_free[v], nc = ndof_node[v] - nf, c = (int)(g - pos_first[t]);
    return pos_nnz[t] + (c < nc ? c : nc + (int64_t)(c - nc) * row_len_free(v));
  };
  // block boundaries: equal share of the non-zeros, exactly as glsnsh_mesh_partition
  std::vector<int64_t> begin(n_ranks + 1, 0);
  for (int r = 1; r < n_ranks; ++r)
    {
      const int64_t target = nnz_global / n_ranks * r;
      int64_t       a = 0, b = M->n_global; // smallest g with rowptr_of(g) >= target
      while (a < b)
        {
          const int64_t mid = (a + b) / 2;
          if (rowptr_of(mid) >= target)
            b = mid;
          else
            a = mid + 1;
        }
      begin[r] = std::min<int64_t>(std::max(a, begin[r - 1]), M->n_global);
    }
  begin[n_ranks]   = M->n_global;
  const int64_t r0 = begin[rank], r1 = begin[rank + 1];
  M->owned_begin   = r0;
  M->n_owned       = r1 - r0;
  auto node_pos_of = [&](int64_t g) -> int64_t {
    return std::upper_bound(pos_first.begin(), pos_first.end(), g) - pos_first.begin() - 1;
  };
  // ---- cells touching an owned dof, ghosts ----
  const int64_t t0 = r1 > r0 ? node_pos_of(r0) : 0, t1 = r1 > r0 ? node_pos_of(r1 - 1) : -1;
  int64_t       ncell_glob = 1;
  for (int d = 0; d < dim; ++d)
    ncell_glob *= M->ncd[d];
  for (int64_t t = t0; t <= t1; ++t)
    {
      int idx[3], ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
      G.split(node_order[t], idx);
      for (int d = 0; d < dim; ++d)
        {
          ca[d] = idx[d] % pu == 0 ? std::max(0, idx[d] / pu - 1) : idx[d] / pu;
          cb[d] = std::min(M->ncd[d] - 1, idx[d] / pu);
        }
      for (int k = ca[2]; k <= cb[2]; ++k)
        for (int j = ca[1]; j <= cb[1]; ++j)
          for (int i = ca[0]; i <= cb[0]; ++i)
            M->cell_ids.push_back(i + (int64_t)M->ncd[0] * (j + (int64_t)M->ncd[1] * k));
    }
  std::sort(M->cell_ids.begin(), M->cell_ids.end());
  M->cell_ids.erase(std::unique(M->cell_ids.begin(), M->cell_ids.end()), M->cell_ids.end());
  M->n_cells = (int64_t)M->cell_ids.size();
  const int nu1 = pu + 1, np1 = pp + 1;
  // global dofs of a cell, in the local dof layout of the ABI
  auto cell_global_dofs = [&](int64_t c, int64_t *out) {
    int ci[3] = {(int)(c % M->ncd[0]), (int)((c / M->ncd[0]) % M->ncd[1]),
                 dim == 3 ? (int)(c / ((int64_t)M->ncd[0] * M->ncd[1])) : 0};
    for (int a = 0; a < M->n_su; ++a)
      {
        int ai[3] = {a % nu1, (a / nu1) % nu1, dim == 3 ? a / (nu1 * nu1) : 0};
        const int64_t v =
          G.join(ci[0] * pu + ai[0], ci[1] * pu + ai[1], dim == 3 ? ci[2] * pu + ai[2] : 0);
        for (int comp = 0; comp < dim; ++comp)
          out[comp * M->n_su + a] = first_dof[v] + comp;
      }
    for (int a = 0; a < M->n_sp; ++a)
      {
        int ai[3] = {a % np1, (a / np1) % np1, dim == 3 ? a / (np1 * np1) : 0};
        const int64_t v = G.join(ci[0] * pu + ai[0] * ratio, ci[1] * pu + ai[1] * ratio,
                                 dim == 3 ? ci[2] * pu + ai[2] * ratio : 0);
        out[dim * M->n_su + a] = first_dof[v] + dim;
      }
  };
  std::vector<int64_t> cg((size_t)M->n_cells * M->n_loc);
#pragma omp parallel for schedule(static)
  for (int64_t lc = 0; lc < M->n_cells; ++lc)
    cell_global_dofs(M->cell_ids[lc], cg.data() + (size_t)lc * M->n_loc);
  std::vector<int64_t> ghosts;
  for (int64_t g : cg)
    if (g < r0 || g >= r1)
      ghosts.push_back(g);
  std::sort(ghosts.begin(), ghosts.end());
  ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
  M->n_dofs     = M->n_owned + (int64_t)ghosts.size();
  auto to_local = [&](int64_t g) -> int32_t {
    if (g >= r0 && g < r1)
      return (int32_t)(g - r0);
    return (int32_t)(M->n_owned +
                     (std::lower_bound(ghosts.begin(), ghosts.end(), g) - ghosts.begin()));
  };
  M->local_to_global.resize(M->n_dofs);
  for (int64_t i = 0; i < M->n_owned; ++i)
    M->local_to_global[i] = r0 + i;
  for (size_t i = 0; i < ghosts.size(); ++i)
    M->local_to_global[M->n_owned + i] = ghosts[i];
  const double hx[3] = {(M->hi[0] - M->lo[0]) / M->ncd[0], (M->hi[1] - M->lo[1]) / M->ncd[1],
                        dim == 3 ? (M->hi[2] - M->lo[2]) / M->ncd[2] : 1.0};
  M->constrained.resize(M->n_dofs), M->cvalues.resize(M->n_dofs), M->dof_comp.resize(M->n_dofs);
  M->dof_coords.resize((size_t)M->n_dofs * dim);
  std::vector<int64_t> node_of_local(M->n_dofs);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < M->n_dofs; ++i)
    {
      const int64_t g = M->local_to_global[i], t = node_pos_of(g), v = node_order[t];
      const int     c = (int)(g - pos_first[t]);
      int           idx[3];
      G.split(v, idx);
      node_of_local[i]  = v;
      const int face    = c < dim ? bc_face(idx) : -1;
      M->constrained[i] = face >= 0;
      M->cvalues[i]     = face >= 0 && bc_type[face] == 2 ? bc_value[face * 3 + c] : 0.0;
      M->dof_comp[i]    = c;
      for (int d = 0; d < dim; ++d)
        M->dof_coords[(size_t)i * dim + d] = M->lo[d] + idx[d] * hx[d] / pu;
    }
  // ---- cells: local dofs, geometry, colours ----
  M->cell_dofs.resize((size_t)M->n_cells * M->n_loc);
  M->inv_jac.assign((size_t)M->n_cells * dim * dim, 0.0);
  M->det_jac.resize(M->n_cells), M->measure.resize(M->n_cells);
  if (with_q_points)
    M->q_points.resize((size_t)M->n_cells * M->n_q * dim);
  std::vector<int32_t> color(M->n_cells);
  double               vol = 1;
  for (int d = 0; d < dim; ++d)
    vol *= hx[d];
#pragma omp parallel for schedule(static)
  for (int64_t lc = 0; lc < M->n_cells; ++lc)
    {
      const int64_t c = M->cell_ids[lc];
      int ci[3] = {(int)(c % M->ncd[0]), (int)((c / M->ncd[0]) % M->ncd[1]),
                   dim == 3 ? (int)(c / ((int64_t)M->ncd[0] * M->ncd[1])) : 0};
      for (int k = 0; k < M->n_loc; ++k)
        M->cell_dofs[(size_t)lc * M->n_loc + k] = to_local(cg[(size_t)lc * M->n_loc + k]);
      for (int d = 0; d < dim; ++d)
        M->inv_jac[(size_t)lc * dim * dim + d * dim + d] = 1.0 / hx[d];
      M->det_jac[lc] = M->measure[lc] = vol;
      if (with_q_points)
        for (int q = 0; q < M->n_q; ++q)
          for (int d = 0; d < dim; ++d)
            M->q_points[((size_t)lc * M->n_q + q) * dim + d] =
              M->lo[d] + (ci[d] + M->xq[(size_t)q * dim + d]) * hx[d];
      color[lc] = (ci[0] & 1) + 2 * (ci[1] & 1) + (dim == 3 ? 4 * (ci[2] & 1) : 0);
    }
  const int ncolor = 1 << dim;
  M->color_ptr.assign(ncolor + 1, 0);
  for (int64_t c = 0; c < M->n_cells; ++c)
    M->color_ptr[color[c] + 1]++;
  for (int k = 0; k < ncolor; ++k)
    M->color_ptr[k + 1] += M->color_ptr[k];
  M->color_cells.resize(M->n_cells);
  {
    std::vector<int32_t> pos(M->color_ptr.begin(), M->color_ptr.end() - 1);
    for (int64_t c = 0; c < M->n_cells; ++c)
      M->color_cells[pos[color[c]]++] = (int32_t)c;
  }
  // ---- CSR of the owned rows ----
  M->rowptr.assign(M->n_owned + 1, 0);
  auto row_entries = [&](int64_t i, int32_t *out) -> int64_t {
    if (M->constrained[i])
      {
        if (out)
          out[0] = (int32_t)i;
        return 1;
      }
    const int64_t v = node_of_local[i];
    if (!out)
      return row_len_free(v);
    int idx[3], a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
    G.split(v, idx);
    for (int d = 0; d < dim; ++d)
      G.range(d, idx[d], a[d], b[d]);
    int64_t cnt = 0;
    for (int k = a[2]; k <= b[2]; ++k)
      for (int j = a[1]; j <= b[1]; ++j)
        for (int ii = a[0]; ii <= b[0]; ++ii)
          {
            const int64_t u  = G.join(ii, j, k);
            const int     nf = n_free[u], nc = ndof_node[u] - nf;
            for (int c = nc; c < ndof_node[u]; ++c)
              out[cnt++] = to_local(first_dof[u] + c);
          }
    std::sort(out, out + cnt);
    return cnt;
  };
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < M->n_owned; ++i)
    M->rowptr[i + 1] = row_entries(i, nullptr);
  for (int64_t i = 0; i < M->n_owned; ++i)
    M->rowptr[i + 1] += M->rowptr[i];
  M->col.resize((size_t)M->rowptr[M->n_owned]);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < M->n_owned; ++i)
    row_entries(i, M->col.data() + M->rowptr[i]);
  // ---- halo lists (as glsnsh_mesh_partition, from the local cells) ----
  std::map<int, std::vector<int64_t>> recv_from, send_to;
  auto                                 owner = [&](int64_t g) {
    return (int)(std::upper_bound(begin.begin(), begin.end(), g) - begin.begin()) - 1;
  };
  for (int64_t g : ghosts)
    recv_from[owner(g)].push_back(g);
  for (int64_t lc = 0; lc < M->n_cells; ++lc)
    {
      const int64_t *cd = cg.data() + (size_t)lc * M->n_loc;
      for (int k = 0; k < M->n_loc; ++k)
        {
          if (cd[k] >= r0 && cd[k] < r1)
            continue;
          const int            o = owner(cd[k]);
          std::vector<int64_t> &s = send_to[o];
          for (int l = 0; l < M->n_loc; ++l)
            if (cd[l] >= r0 && cd[l] < r1)
              s.push_back(cd[l]);
          if (s.size() > (size_t)(1 << 22))
            { // keep the lists small while they are built
              std::sort(s.begin(), s.end());
              s.erase(std::unique(s.begin(), s.end()), s.end());
            }
        }
    }
  std::vector<int> nb;
  for (auto &kv : recv_from)
    nb.push_back(kv.first);
  for (auto &kv : send_to)
    nb.push_back(kv.first);
  std::sort(nb.begin(), nb.end());
  nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
  M->send_ptr.assign(1, 0), M->recv_ptr.assign(1, 0);
  for (int o : nb)
    {
      M->neighbor_rank.push_back(o);
      std::vector<int64_t> &s = send_to[o];
      std::sort(s.begin(), s.end());
      s.erase(std::unique(s.begin(), s.end()), s.end());
      for (int64_t g : s)
        M->send_idx.push_back((int32_t)(g - r0));
      M->send_ptr.push_back((int64_t)M->send_idx.size());
      M->recv_ptr.push_back(M->recv_ptr.back() + (int64_t)recv_from[o].size());
    }
  return (glsnsh_mesh *)M;
}

void
glsnsh_mesh_destroy(glsnsh_mesh *m)
{
  delete (HostMesh *)m;
}

const char *
glsnsh_mesh_error(const glsnsh_mesh *m)
{
  return ((const HostMesh *)m)->error.c_str();
}

// out[0..11]: n_dofs, n_owned, n_cells, nnz, n_loc, n_q, n_su, n_sp, n_colors,
//             n_neighbors, n_global, owned_begin
void
glsnsh_mesh_info(const glsnsh_mesh *m, int64_t *out)
{
  const HostMesh *M = (const HostMesh *)m;
  out[0] = M->n_dofs, out[1] = M->n_owned, out[2] = M->n_cells;
  out[3] = M->rowptr.empty() ? 0 : M->rowptr[M->n_owned];
  out[4] = M->n_loc, out[5] = M->n_q, out[6] = M->n_su, out[7] = M->n_sp;
  out[8] = (int64_t)M->color_ptr.size() - 1, out[9] = (int64_t)M->neighbor_rank.size();
  out[10] = M->n_global, out[11] = M->owned_begin;
}

// Borrowed pointer to a named array (valid until glsnsh_mesh_destroy); *count
// receives the number of elements.
const void *
glsnsh_mesh_array(const glsnsh_mesh *m, const char *name, int64_t *count)
{
  const HostMesh *M = (const HostMesh *)m;
  const std::string n(name);
#define ARR(key, vec)                   \
  if (n == key)                         \
    {                                   \
      if (count)                        \
        *count = (int64_t)M->vec.size(); \
      return M->vec.data();             \
    }
  ARR("cell_dofs", cell_dofs)
  ARR("col_idx", col)
  ARR("row_ptr", rowptr)
  ARR("constrained", constrained)
  ARR("constraint_values", cvalues)
  ARR("inv_jacobian", inv_jac)
  ARR("det_jacobian", det_jac)
  ARR("cell_measure", measure)
  ARR("q_points", q_points)
  ARR("color_ptr", color_ptr)
  ARR("color_cells", color_cells)
  ARR("dof_component", dof_comp)
  ARR("dof_coords", dof_coords)
  ARR("local_to_global", local_to_global)
  ARR("cell_ids", cell_ids)
  ARR("periodic_slave", periodic_slave)
  ARR("periodic_master", periodic_master)
  ARR("neighbor_rank", neighbor_rank)
  ARR("send_ptr", send_ptr)
  ARR("send_idx", send_idx)
  ARR("recv_ptr", recv_ptr)
  ARR("shape_u", Nu)
  ARR("grad_u", dNu)
  ARR("hess_u", d2Nu)
  ARR("shape_p", Np)
  ARR("grad_p", dNp)
  ARR("weights", wq)
  ARR("unit_q_points", xq)
#undef ARR
  if (count)
    *count = -1;
  return nullptr;
}

// Fill the ABI descriptors (pointers borrowed from the mesh object).
void
glsnsh_mesh_fill_desc(const glsnsh_mesh *m, glsns_fe_desc *fe, glsns_mesh_desc *md)
{
  const HostMesh *M = (const HostMesh *)m;
  if (fe)
    {
      fe->dim = M->dim, fe->velocity_degree = M->pu, fe->n_su = M->n_su, fe->n_sp = M->n_sp;
      fe->n_q = M->n_q;
      fe->shape_u = M->Nu.data(), fe->grad_u = M->dNu.data(), fe->hess_u = M->d2Nu.data();
      fe->shape_p = M->Np.data(), fe->grad_p = M->dNp.data(), fe->weights = M->wq.data();
    }
  if (md)
    {
      memset(md, 0, sizeof(*md));
      md->n_dofs = M->n_dofs, md->n_owned = M->n_owned, md->n_cells = M->n_cells;
      md->cell_dofs      = M->cell_dofs.data();
      md->geometry_per_q = 0;
      md->inv_jacobian = M->inv_jac.data(), md->det_jacobian = M->det_jac.data();
      md->cell_measure = M->measure.data();
      md->q_points     = M->q_points.empty() ? nullptr : M->q_points.data();
      md->constrained = M->constrained.data(), md->constraint_values = M->cvalues.data();
      md->row_ptr = M->rowptr.data(), md->col_idx = M->col.data();
      md->n_colors  = (int32_t)M->color_ptr.size() - 1;
      md->color_ptr = M->color_ptr.data(), md->color_cells = M->color_cells.data();
      md->n_neighbors   = (int32_t)M->neighbor_rank.size();
      md->neighbor_rank = M->neighbor_rank.data();
      md->send_ptr = M->send_ptr.data(), md->send_idx = M->send_idx.data();
      md->recv_ptr = M->recv_ptr.data();
    }
}

} // extern "C"
