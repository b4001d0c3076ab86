// glsns_dealii_adapter.hpp — fills glsns_fe_desc / glsns_mesh_desc from deal.II objects.
//
// This is the reference-side half of the drop-in boundary: what GLSNavierStokesSolver<dim> calls
// at the end of setup_dofs() (source/solvers/gls_navier_stokes.cc:57-228, after every refinement)
// to hand the mesh-dependent data of the hot path to the library.  It needs deal.II >= 9.2 with
// Trilinos and MPI, which this repository's build image does not have: the file is compiled only
// where <deal.II/base/config.h> is found (everything is inside `#ifdef DEAL_II_VERSION_MAJOR`),
// it has NOT been compiled or run by this repository's tests, and the arrays it produces are
// specified by -- and tested through -- the stand-in that fills the same descriptors for box
// meshes (softx_2020_200_b200/csrc/host_mesh.cpp, tests/test_host_mirror.py).  INTEGRATION.md
// section 3 shows the three overrides that use it.
//
// Conventions it implements (include/glsns.h):
//   * local dof order of a cell: k = c * n_su + a (velocity component c), dim * n_su + a (pressure),
//     through fe.system_to_component_index;
//   * local numbering of a rank: owned dofs in global order, then ghosts grouped by owner rank;
//   * cells: every active cell with at least one owned dof (owned cells and the ghost layer;
//     the library assembles ghost-layer cells redundantly instead of compress(add));
//   * geometry: affine cells -> one inverse Jacobian / determinant per cell; otherwise per
//     quadrature point, with mapping_laplacian from the Jacobian gradients;
//   * constraints: homogeneous Dirichlet lines of zero_constraints as a mask (1), inhomogeneities
//     of nonzero_constraints as values; hanging-node lines (mask 2) as a CSR of (master, weight)
//     taken from the CLOSED zero_constraints, their inhomogeneities from the closed
//     nonzero_constraints; the cell colouring separates cells that share a master.  The library
//     takes hanging-node lines on one rank only.
#ifndef GLSNS_DEALII_ADAPTER_HPP
#define GLSNS_DEALII_ADAPTER_HPP

#if __has_include(<deal.II/base/config.h>)
#  include <deal.II/base/config.h>
#endif

#ifdef DEAL_II_VERSION_MAJOR

#  include <deal.II/base/index_set.h>
#  include <deal.II/base/mpi.h>
#  include <deal.II/base/quadrature_lib.h>
#  include <deal.II/base/utilities.h>
#  include <deal.II/dofs/dof_handler.h>
#  include <deal.II/dofs/dof_tools.h>
#  include <deal.II/fe/fe_system.h>
#  include <deal.II/fe/fe_values.h>
#  include <deal.II/fe/mapping_q.h>
#  include <deal.II/grid/grid_generator.h>
#  include <deal.II/grid/tria.h>
#  include <deal.II/lac/affine_constraints.h>
#  include <deal.II/lac/dynamic_sparsity_pattern.h>

#  include <algorithm>
#  include <map>
#  include <stdexcept>
#  include <vector>

#  include "glsns.h"

namespace glsns
{
  namespace dealii_adapter
  {
    using namespace dealii;

    // Owns the host arrays the descriptors point to (they are only borrowed by glsns_set_fe /
    // glsns_set_mesh, so this object may be destroyed right after those calls).
    template <int dim>
    struct HostArrays
    {
      // fe
      std::vector<double> shape_u, grad_u, hess_u, shape_p, grad_p, weights;
      std::vector<unsigned int> lib_index; // deal.II local dof i -> library local dof
      // mesh
      std::vector<int32_t> cell_dofs, col_idx, color_ptr, color_cells, neighbor_rank, send_idx;
      std::vector<int64_t> row_ptr, send_ptr, recv_ptr;
      std::vector<double>  inv_jacobian, det_jacobian, cell_measure, q_points, constraint_values,
        mapping_laplacian, constraint_weight, constraint_inhomogeneity;
      std::vector<int64_t> constraint_ptr;
      std::vector<int32_t> constraint_idx;
      std::vector<uint8_t> constrained;
      // numbering
      std::vector<types::global_dof_index> local_to_global; // owned first, then ghosts by owner
      std::map<types::global_dof_index, int32_t> ghost_local;
      types::global_dof_index                    owned_begin = 0;
      int64_t                                    n_owned = 0, n_dofs = 0;

      int32_t
      to_local(const types::global_dof_index g) const
      {
        if (g >= owned_begin && g < owned_begin + (types::global_dof_index)n_owned)
          return (int32_t)(g - owned_begin);
        return ghost_local.at(g);
      }
      // evaluation_point etc. in library order: v[to_local(g)] = ghosted_vector[g]
      template <typename VectorType>
      void
      gather(const VectorType &ghosted, std::vector<double> &out) const
      {
        out.resize(local_to_global.size());
        for (std::size_t i = 0; i < local_to_global.size(); ++i)
          out[i] = ghosted[local_to_global[i]];
      }
    };

    // FESystem(FE_Q(pu)^dim, FE_Q(pp)) on the reference cell at the points of QGauss
    // (gls_navier_stokes.cc:244-252; navier_stokes_base.cc:62,70,93-94)
    template <int dim>
    void
    fill_fe_desc(const FESystem<dim> &fe, const Quadrature<dim> &quadrature,
                 const unsigned int velocity_degree, HostArrays<dim> &h, glsns_fe_desc &out)
    {
      const unsigned int n_q = quadrature.size(), n = fe.dofs_per_cell;
      unsigned int       n_su = 0, n_sp = 0;
      for (unsigned int i = 0; i < n; ++i)
        {
          const auto ci = fe.system_to_component_index(i);
          if (ci.first == 0)
            n_su = std::max(n_su, ci.second + 1);
          if (ci.first == dim)
            n_sp = std::max(n_sp, ci.second + 1);
        }
      h.lib_index.resize(n);
      for (unsigned int i = 0; i < n; ++i)
        {
          const auto ci  = fe.system_to_component_index(i);
          h.lib_index[i] = ci.first < (unsigned int)dim ? ci.first * n_su + ci.second : dim * n_su + ci.second;
        }
      h.shape_u.assign(n_q * n_su, 0), h.grad_u.assign(n_q * n_su * dim, 0);
      h.hess_u.assign(n_q * n_su * dim * dim, 0);
      h.shape_p.assign(n_q * n_sp, 0), h.grad_p.assign(n_q * n_sp * dim, 0);
      h.weights.resize(n_q);
      for (unsigned int q = 0; q < n_q; ++q)
        {
          const Point<dim> &xi = quadrature.point(q);
          h.weights[q]         = quadrature.weight(q);
          for (unsigned int i = 0; i < n; ++i)
            {
              const auto ci = fe.system_to_component_index(i);
              if (ci.first == 0)
                {
                  const unsigned int a = ci.second;
                  h.shape_u[q * n_su + a] = fe.shape_value_component(i, xi, 0);
                  const Tensor<1, dim> g  = fe.shape_grad_component(i, xi, 0);
                  const Tensor<2, dim> H  = fe.shape_grad_grad_component(i, xi, 0);
                  for (unsigned int d = 0; d < dim; ++d)
                    {
                      h.grad_u[(q * n_su + a) * dim + d] = g[d];
                      for (unsigned int e = 0; e < dim; ++e)
                        h.hess_u[((q * n_su + a) * dim + d) * dim + e] = H[d][e];
                    }
                }
              else if (ci.first == (unsigned int)dim)
                {
                  const unsigned int a = ci.second;
                  h.shape_p[q * n_sp + a] = fe.shape_value_component(i, xi, dim);
                  const Tensor<1, dim> g  = fe.shape_grad_component(i, xi, dim);
                  for (unsigned int d = 0; d < dim; ++d)
                    h.grad_p[(q * n_sp + a) * dim + d] = g[d];
                }
            }
        }
      out.dim = dim, out.velocity_degree = (int32_t)velocity_degree;
      out.n_su = (int32_t)n_su, out.n_sp = (int32_t)n_sp, out.n_q = (int32_t)n_q;
      out.shape_u = h.shape_u.data(), out.grad_u = h.grad_u.data(), out.hess_u = h.hess_u.data();
      out.shape_p = h.shape_p.data(), out.grad_p = h.grad_p.data(), out.weights = h.weights.data();
    }

    // What setup_dofs() leaves behind, restricted to this rank.  `dsp` is the sparsity pattern of
    // gls_navier_stokes.cc:204-213 (make_sparsity_pattern(..., zero_constraints,
    // keep_constrained_dofs = false) after distribute_sparsity_pattern), rows in global indices.
    template <int dim>
    void
    fill_mesh_desc(const DoFHandler<dim> &dof_handler, const Mapping<dim> &mapping,
                   const FESystem<dim> &fe, const Quadrature<dim> &quadrature,
                   const AffineConstraints<double> &zero_constraints,
                   const AffineConstraints<double> &nonzero_constraints,
                   const IndexSet &locally_owned_dofs, const DynamicSparsityPattern &dsp,
                   const bool need_q_points, const MPI_Comm comm, HostArrays<dim> &h,
                   glsns_mesh_desc &out)
    {
      if (!locally_owned_dofs.is_contiguous())
        throw std::runtime_error("glsns: the locally owned dofs must be one contiguous range");
      const unsigned int n = fe.dofs_per_cell, n_q = quadrature.size();
      h.n_owned     = locally_owned_dofs.n_elements();
      h.owned_begin = h.n_owned ? locally_owned_dofs.nth_index_in_set(0) : 0;
      // ---- cells with at least one owned dof; ghosts = their other dofs ----
      std::vector<typename DoFHandler<dim>::active_cell_iterator> cells;
      std::vector<types::global_dof_index>                        dofs(n), ghosts;
      for (const auto &cell : dof_handler.active_cell_iterators())
        if (cell->is_locally_owned() || cell->is_ghost())
          {
            cell->get_dof_indices(dofs);
            bool touches = false;
            for (const auto g : dofs)
              touches = touches || locally_owned_dofs.is_element(g);
            if (!touches)
              continue;
            cells.push_back(cell);
            for (const auto g : dofs)
              if (!locally_owned_dofs.is_element(g))
                ghosts.push_back(g);
          }
      std::sort(ghosts.begin(), ghosts.end());
      ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
      // global order = grouped by owner, because every rank owns one contiguous range
      h.n_dofs = h.n_owned + (int64_t)ghosts.size();
      h.local_to_global.resize(h.n_dofs);
      for (int64_t i = 0; i < h.n_owned; ++i)
        h.local_to_global[i] = h.owned_begin + i;
      h.ghost_local.clear();
      for (std::size_t i = 0; i < ghosts.size(); ++i)
        {
          h.local_to_global[h.n_owned + i] = ghosts[i];
          h.ghost_local[ghosts[i]]         = (int32_t)(h.n_owned + i);
        }
      // ---- constraints: Dirichlet lines (1) and hanging-node lines (2) of the closed objects ----
      h.constrained.assign(h.n_dofs, 0), h.constraint_values.assign(h.n_dofs, 0.0);
      h.constraint_ptr.assign(h.n_dofs + 1, 0), h.constraint_idx.clear(), h.constraint_weight.clear();
      h.constraint_inhomogeneity.assign(h.n_dofs, 0.0);
      bool any_hanging = false;
      for (int64_t i = 0; i < h.n_dofs; ++i)
        {
          const auto g = h.local_to_global[i];
          if (zero_constraints.is_constrained(g))
            {
              // a line that had entries BEFORE close() is a hanging-node line even if all its
              // masters were Dirichlet dofs and dropped out; a boundary-value line never has any
              const auto *entries  = zero_constraints.get_constraint_entries(g);
              const auto *entries_n = nonzero_constraints.get_constraint_entries(g);
              const bool  hanging  = (entries && !entries->empty()) || (entries_n && !entries_n->empty());
              h.constrained[i]     = hanging ? 2 : 1;
              any_hanging          = any_hanging || hanging;
              if (hanging && entries)
                for (const auto &e : *entries)
                  {
                    h.constraint_idx.push_back(h.to_local(e.first));
                    h.constraint_weight.push_back(e.second);
                  }
              if (hanging)
                h.constraint_inhomogeneity[i] = nonzero_constraints.get_inhomogeneity(g);
            }
          h.constraint_ptr[i + 1] = (int64_t)h.constraint_idx.size();
          if (nonzero_constraints.is_constrained(g) && h.constrained[i] != 2)
            h.constraint_values[i] = nonzero_constraints.get_inhomogeneity(g);
        }
      // ---- cells: dofs in library order, geometry ----
      const std::size_t nc = cells.size();
      h.cell_dofs.resize(nc * n), h.cell_measure.resize(nc);
      FEValues<dim> fe_values(mapping, fe, quadrature,
                              update_inverse_jacobians | update_JxW_values | update_quadrature_points |
                                update_jacobian_grads);
      // first pass: is every cell affine (inverse Jacobian constant over the points)?
      std::vector<double> invJ(nc * n_q * dim * dim), det(nc * n_q), lap(nc * n_q * dim, 0.0), xq;
      if (need_q_points)
        xq.resize(nc * n_q * dim);
      bool affine = true;
      for (std::size_t c = 0; c < nc; ++c)
        {
          const auto &cell = cells[c];
          cell->get_dof_indices(dofs);
          for (unsigned int i = 0; i < n; ++i)
            h.cell_dofs[c * n + h.lib_index[i]] = h.to_local(dofs[i]);
          h.cell_measure[c] = cell->measure(); // what h is computed from, :340-345
          fe_values.reinit(cell);
          for (unsigned int q = 0; q < n_q; ++q)
            {
              const DerivativeForm<1, dim, dim> K = fe_values.inverse_jacobian(q); // K[r][d] = dxi_r/dx_d
              const DerivativeForm<2, dim, dim> G = fe_values.jacobian_grad(q);    // G[k][r][s] = d2x_k/dxi_r dxi_s
              for (unsigned int r = 0; r < dim; ++r)
                for (unsigned int d = 0; d < dim; ++d)
                  {
                    invJ[((c * n_q + q) * dim + r) * dim + d] = K[r][d];
                    affine = affine && std::abs(K[r][d] - invJ[((c * n_q) * dim + r) * dim + d]) <=
                                         1e-14 * (std::abs(K[r][d]) + 1e-300);
                  }
              det[c * n_q + q] = fe_values.JxW(q) / quadrature.weight(q);
              for (unsigned int k = 0; k < dim; ++k)
                {
                  double s = 0;
                  for (unsigned int r = 0; r < dim; ++r)
                    for (unsigned int t = 0; t < dim; ++t)
                      {
                        double kkt = 0;
                        for (unsigned int d = 0; d < dim; ++d)
                          kkt += K[r][d] * K[t][d];
                        s += G[k][r][t] * kkt;
                      }
                  lap[(c * n_q + q) * dim + k] = s;
                }
              if (need_q_points)
                for (unsigned int d = 0; d < dim; ++d)
                  xq[(c * n_q + q) * dim + d] = fe_values.quadrature_point(q)[d];
            }
        }
      // every rank must take the same branch: the kernels are compiled for one layout per mesh
      affine = Utilities::MPI::min((int)affine, comm) != 0;
      if (affine)
        {
          h.inv_jacobian.resize(nc * dim * dim), h.det_jacobian.resize(nc);
          for (std::size_t c = 0; c < nc; ++c)
            {
              std::copy(invJ.begin() + c * n_q * dim * dim, invJ.begin() + (c * n_q + 1) * dim * dim,
                        h.inv_jacobian.begin() + c * dim * dim);
              h.det_jacobian[c] = det[c * n_q];
            }
          h.mapping_laplacian.clear();
        }
      else
        {
          h.inv_jacobian.swap(invJ), h.det_jacobian.swap(det), h.mapping_laplacian.swap(lap);
        }
      h.q_points.swap(xq);
      // ---- sparsity of the owned rows, local column indices, sorted ----
      h.row_ptr.assign(h.n_owned + 1, 0);
      for (int64_t i = 0; i < h.n_owned; ++i)
        h.row_ptr[i + 1] = h.row_ptr[i] + dsp.row_length(h.owned_begin + i);
      h.col_idx.resize(h.row_ptr[h.n_owned]);
      for (int64_t i = 0; i < h.n_owned; ++i)
        {
          int32_t *o = h.col_idx.data() + h.row_ptr[i];
          for (unsigned int k = 0; k < dsp.row_length(h.owned_begin + i); ++k)
            o[k] = h.to_local(dsp.column_number(h.owned_begin + i, k));
          std::sort(o, o + (h.row_ptr[i + 1] - h.row_ptr[i]));
        }
      // ---- colouring: greedy, cells of one colour share no dof and no master of a dof ----
      {
        std::vector<int32_t>              color(nc, -1);
        std::vector<std::vector<int32_t>> dof_colors(h.n_dofs); // colours already used at a dof
        int32_t                           ncolor = 0;
        std::vector<int32_t>              touched;
        for (std::size_t c = 0; c < nc; ++c)
          {
            touched.clear();
            for (unsigned int i = 0; i < n; ++i)
              {
                const int32_t d = h.cell_dofs[c * n + i];
                touched.push_back(d);
                for (int64_t k = h.constraint_ptr[d]; k < h.constraint_ptr[d + 1]; ++k)
                  touched.push_back(h.constraint_idx[k]);
              }
            int32_t pick = 0;
            for (;; ++pick)
              {
                bool clash = false;
                for (const int32_t d : touched)
                  {
                    for (const int32_t used : dof_colors[d])
                      if (used == pick)
                        {
                          clash = true;
                          break;
                        }
                    if (clash)
                      break;
                  }
                if (!clash)
                  break;
              }
            color[c] = pick;
            ncolor   = std::max(ncolor, pick + 1);
            for (const int32_t d : touched)
              dof_colors[d].push_back(pick);
          }
        h.color_ptr.assign(ncolor + 1, 0);
        for (std::size_t c = 0; c < nc; ++c)
          h.color_ptr[color[c] + 1]++;
        for (int32_t k = 0; k < ncolor; ++k)
          h.color_ptr[k + 1] += h.color_ptr[k];
        h.color_cells.resize(nc);
        std::vector<int32_t> pos(h.color_ptr.begin(), h.color_ptr.end() - 1);
        for (std::size_t c = 0; c < nc; ++c)
          h.color_cells[pos[color[c]]++] = (int32_t)c;
      }
      // ---- halo: receive our ghosts from their owners; send what the neighbours' assembled cells need ----
      {
        const unsigned int n_ranks = Utilities::MPI::n_mpi_processes(comm);
        const auto         owned_per_rank =
          Utilities::MPI::all_gather(comm, std::make_pair(h.owned_begin, (types::global_dof_index)h.n_owned));
        auto owner = [&](const types::global_dof_index g) {
          for (unsigned int r = 0; r < n_ranks; ++r)
            if (g >= owned_per_rank[r].first && g < owned_per_rank[r].first + owned_per_rank[r].second)
              return r;
          throw std::runtime_error("glsns: dof without an owner");
        };
        std::map<unsigned int, std::vector<types::global_dof_index>> recv_from, send_to;
        for (const auto g : ghosts)
          recv_from[owner(g)].push_back(g);
        // a neighbour that owns a dof of one of our cells assembles that cell too (it is in its
        // ghost layer) and reads OUR owned dofs of it
        for (std::size_t c = 0; c < nc; ++c)
          for (unsigned int i = 0; i < n; ++i)
            {
              const auto gi = h.local_to_global[h.cell_dofs[c * n + i]];
              if (locally_owned_dofs.is_element(gi))
                continue;
              auto &s = send_to[owner(gi)];
              for (unsigned int j = 0; j < n; ++j)
                {
                  const auto gj = h.local_to_global[h.cell_dofs[c * n + j]];
                  if (locally_owned_dofs.is_element(gj))
                    s.push_back(gj);
                }
            }
        std::vector<unsigned int> nb;
        for (const auto &kv : recv_from)
          nb.push_back(kv.first);
        for (const auto &kv : send_to)
          nb.push_back(kv.first);
        std::sort(nb.begin(), nb.end());
        nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
        h.neighbor_rank.clear(), h.send_idx.clear();
        h.send_ptr.assign(1, 0), h.recv_ptr.assign(1, 0);
        for (const unsigned int o : nb)
          {
            h.neighbor_rank.push_back((int32_t)o);
            auto &s = send_to[o];
            std::sort(s.begin(), s.end());
            s.erase(std::unique(s.begin(), s.end()), s.end());
            for (const auto g : s)
              h.send_idx.push_back((int32_t)(g - h.owned_begin));
            h.send_ptr.push_back((int64_t)h.send_idx.size());
            h.recv_ptr.push_back(h.recv_ptr.back() + (int64_t)recv_from[o].size());
          }
      }
      // ---- the descriptor ----
      out                   = glsns_mesh_desc();
      out.n_dofs            = h.n_dofs, out.n_owned = h.n_owned, out.n_cells = (int64_t)nc;
      out.cell_dofs         = h.cell_dofs.data();
      out.geometry_per_q    = affine ? 0 : 1;
      out.inv_jacobian      = h.inv_jacobian.data(), out.det_jacobian = h.det_jacobian.data();
      out.cell_measure      = h.cell_measure.data();
      out.q_points          = h.q_points.empty() ? nullptr : h.q_points.data();
      out.constrained       = h.constrained.data();
      out.constraint_values = h.constraint_values.data();
      out.row_ptr = h.row_ptr.data(), out.col_idx = h.col_idx.data();
      out.n_colors  = (int32_t)h.color_ptr.size() - 1;
      out.color_ptr = h.color_ptr.data(), out.color_cells = h.color_cells.data();
      out.n_neighbors   = (int32_t)h.neighbor_rank.size();
      out.neighbor_rank = h.neighbor_rank.data();
      out.send_ptr = h.send_ptr.data(), out.send_idx = h.send_idx.data();
      out.recv_ptr          = h.recv_ptr.data();
      out.mapping_laplacian = affine ? nullptr : h.mapping_laplacian.data();
      if (any_hanging)
        {
          out.constraint_ptr           = h.constraint_ptr.data();
          out.constraint_idx           = h.constraint_idx.data();
          out.constraint_weight        = h.constraint_weight.data();
          out.constraint_inhomogeneity = h.constraint_inhomogeneity.data();
        }
    }
  } // namespace dealii_adapter
} // namespace glsns

#endif // DEAL_II_VERSION_MAJOR
#endif // GLSNS_DEALII_ADAPTER_HPP
