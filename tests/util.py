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


def rank_local_reference(oracle, n_cells, world, nu=0.005, atol=1e-12):
    """The partitioned 3D Q2-Q2 cavity of tests/multi_gpu_check.py without a communicator: the
    global C++ mesh, every rank's part of it, and the oracle's matrix, right-hand side, block-Jacobi
    ILU(0) factors (the ranks' row blocks: Ifpack with overlap 0), SpMV and ILU application on a
    seeded state -- what each rank's block must reproduce on its own."""
    from softx_2020_200_b200.mesh import BoxMesh
    bcs = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"),
           (3, "function", (1.0, 0.0, 0.0))]
    g = BoxMesh(3, n_cells, 2, 2, bcs=bcs)
    parts = [g.partition(world, r) for r in range(world)]
    lid = lambda x: np.stack([np.ones(len(x)), 0 * x[:, 0], 0 * x[:, 0]], axis=1)
    obcs = {0: ("noslip",), 1: ("noslip",), 2: ("noslip",), 4: ("noslip",), 5: ("noslip",),
            3: ("function", lid)}
    prov = oracle.BoxMesh(3, n_cells, 2, 2, bcs=obcs, renumber="none")
    key = lambda c, k: np.lexsort(tuple(np.round(c[:, d] * 1e6).astype(np.int64) for d in range(3)) + (k,))
    new_of_old = np.empty(prov.ndof, dtype=np.int64)
    new_of_old[key(prov.dof_coords, prov.dof_comp)] = key(g.array("dof_coords").reshape(-1, 3),
                                                          g.array("dof_component"))
    om = oracle.BoxMesh(3, n_cells, 2, 2, bcs=obcs, renumber=new_of_old)
    xyz = g.array("dof_coords").reshape(-1, 3)
    U = 0.05 * np.sin(np.pi * xyz[:, 0] + 0.3 * g.array("dof_component")) * np.cos(np.pi * xyz[:, 1]) \
        * np.cos(0.5 * np.pi * xyz[:, 2])
    con = g.array("constrained").astype(bool)
    U[con] = g.array("constraint_values")[con]
    a, b = oracle.assemble(om, U, oracle.scheme_params("steady", None, nu), True)
    bp = np.concatenate([[0], np.cumsum([p.n_owned for p in parts])]).astype(np.int64)
    lu, dp = oracle.ilu0(om, a, atol, 1.0, block_ptr=bp)
    x = np.random.default_rng(11).standard_normal(om.ndof)
    return dict(global_mesh=g, parts=parts, oracle_mesh=om, state=U, matrix=a, rhs=b, block_ptr=bp,
                ilu=lu, x=x, spmv=oracle.spmv(om, a, x), ilu_apply=oracle.ilu_apply(om, lu, dp, x, block_ptr=bp))
