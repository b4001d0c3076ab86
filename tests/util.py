"""Shared helpers of the parity tests: feed an oracle mesh through the C ABI."""
import numpy as np


def hotpath_from_oracle_mesh(mesh, viscosity=1.0, force=None, srf=False, omega=(0, 0, 0), device=0):
    """A GLSHotPath context set up with the host-side arrays of an oracle BoxMesh (the arrays a
    deal.II adapter would hand over)."""
    from softx_2020_200_b200 import GLSHotPath
    fe = mesh.fe
    hp = GLSHotPath(device)
    hp.set_fe(mesh.dim, mesh.pu, fe.Nu, fe.dNu, fe.d2Nu, fe.Np, fe.dNp, fe.wq)
    ptr, order, _ = mesh.color_lists()
    hp.set_mesh(mesh.ndof, mesh.cell_dofs, mesh.cell_invJ, mesh.cell_detJ, mesh.cell_measure,
                mesh.constrained, mesh.rowptr, mesh.col, ptr, order, q_points=mesh.qpoints,
                constraint_values=mesh.constraint_value,
                geometry_per_q=getattr(mesh, "geometry_per_q", False),
                mapping_laplacian=getattr(mesh, "map_lap", None),
                constraint_ptr=getattr(mesh, "hang_ptr", None),
                constraint_idx=getattr(mesh, "hang_idx", None),
                constraint_weight=getattr(mesh, "hang_w", None),
                constraint_inhomogeneity=getattr(mesh, "hang_inhomogeneity", None))
    hp.set_physics(viscosity, srf, omega)
    hp.set_forcing(force)
    return hp


def row_scaled_error(mesh, a_gpu, a_ref):
    """max_i max_j |A_gpu - A_ref|(i,j) / max_j |A_ref(i,j)|  (SURVEY.md §8a note 5)."""
    rows = np.repeat(np.arange(mesh.ndof), np.diff(mesh.rowptr))
    rowmax = np.zeros(mesh.ndof)
    np.maximum.at(rowmax, rows, np.abs(a_ref))
    rowmax[rowmax == 0] = 1.0
    return float(np.max(np.abs(a_gpu - a_ref) / rowmax[rows]))


def random_state(mesh, seed=1234, scale=1.0):
    rng = np.random.default_rng(seed)
    U = scale * rng.uniform(-1.0, 1.0, mesh.ndof)
    return mesh.apply_nonzero_constraints(U)
