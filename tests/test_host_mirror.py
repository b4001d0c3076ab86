"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol,
the transcribed reference Newton drivers (tests/mirror, a test harness) reproduce the reference's
fake-backend tests on the product's PhysicsSolver interface, the
.prm reader honours the reference's option names and defaults, and the C++ box-mesh stand-in
produces exactly the arrays of the oracle's deal.II restatement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_of_the_header():
    from softx_2020_200_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "glsns.h")).read()
    declared = set(re.findall(r"\b(glsns_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert b"sm_100a" in L.glsns_version()


def test_no_cpu_fallback_without_a_device():
    """glsns_create fails loudly when there is no CUDA device (no CPU path exists)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from softx_2020_200_b200 import GLSHotPath, GlsnsError
    with pytest.raises(GlsnsError):
        GLSHotPath(0)


@pytest.mark.parametrize("skip,skip_iterations", [(0, 1), (1, 1), (1, 3)])
def test_newton_drivers_on_the_reference_fake_backend(skip, skip_iterations):
    """tests/core/newton_non_linear_solver_01.output and skip_newton_non_linear_solver_01.output:
    'The final solution is : 1.22474 -1.50000' for x0^2+x1=0, 2x1+3=0 from (1,0)."""
    from tests.mirror.solver import mirror_lib
    L = mirror_lib()
    x = (C.c_double * 2)()
    n_matrix = L.glsnsh_newton_toy(skip, skip_iterations, x)
    assert "%.5f %.5f" % (x[0], x[1]) == "1.22474 -1.50000"
    # newton rebuilds the Jacobian every iteration, skip_newton once per call
    assert n_matrix == (1 if skip else 4)


def _parse(text):
    from tests.mirror.solver import mirror_lib
    L = mirror_lib()
    out = (C.c_double * 18)()
    err = C.create_string_buffer(256)
    rc = L.glsnsh_parse_prm(text.encode(), out, err, 256)
    return rc, list(out), err.value.decode()


def test_prm_defaults_match_the_reference():
    """source/core/parameters.cc:373-447 (non-linear solver), :502-559 (linear solver)."""
    rc, v, _ = _parse("")
    assert rc == 0
    assert v[:4] == [1e-6, 10, 1, 0]                      # tolerance, max it, skip it, newton
    assert v[4:11] == [1e-3, 1e-8, 1000, 0, 1e-8, 1.0, 0]  # rel, abs, max iters, fill, atol, rtol, gmres
    assert v[11:16] == [1.0, 1, 1, 0, 0.0]
    assert v[16:18] == [3, 1.0]                            # initial conditions: nodal, viscosity 1


def test_prm_reads_the_shipped_cavity_file_syntax():
    text = """
# Listing of Parameters
subsection physical properties
    set kinematic viscosity            = 0.005
end
subsection FEM
    set velocity order            = 2
    set pressure order            = 2
end
subsection non-linear solver
  set solver                  = skip_newton
  set tolerance               = 1e-8
  set max iterations          = 10
  set skip iterations         = 3
  set verbosity               = quiet
end
subsection linear solver
  set method                                 = gmres
  set max iters                              = 5000
  set relative residual                      = 1e-4
  set minimum residual                       = 1e-9
  set ilu preconditioner fill                = 0
  set ilu preconditioner absolute tolerance  = 1e-12
  set ilu preconditioner relative tolerance  = 1.00
  set verbosity               = quiet
end
subsection velocity source
  set type = srf
  set omega_z = -6.28318
end
"""
    rc, v, _ = _parse(text)
    assert rc == 0
    assert v[:4] == [1e-8, 10, 3, 1]
    assert v[4:11] == [1e-4, 1e-9, 5000, 0, 1e-12, 1.0, 0]
    assert v[11:16] == [0.005, 2, 2, 1, -6.28318]
    rc, _, err = _parse("subsection linear solver\n set method = direct\nend\n")
    assert rc == 1 and "invalid iterative solver type" in err
    rc, _, err = _parse("subsection non-linear solver\n set verbosity = loud\nend\n")
    assert rc == 1 and "Invalid verbosity level" in err
    # include/solvers/initial_conditions.h:69-121 (as in taylor-green-vortex_gls_*.prm)
    rc, v, _ = _parse("subsection initial conditions\n set type = L2projection\n set viscosity = 0.1\nend\n")
    assert rc == 0 and v[16:18] == [1, 0.1]
    rc, v, _ = _parse("subsection initial conditions\n set type = viscous\nend\n")
    assert rc == 0 and v[16:18] == [2, 1.0]
    rc, _, err = _parse("subsection initial conditions\n set type = random\nend\n")
    assert rc == 1 and "L2projection|viscous|nodal" in err


def _match_numbering(A, B, dim):
    key = lambda m: np.lexsort(tuple(np.round(m.array("dof_coords").reshape(-1, dim)[:, d] * 1e6)
                                     .astype(np.int64) for d in range(dim))
                               + (m.array("dof_component"),))
    new_of_old = np.empty(A.n_dofs, dtype=np.int64)
    new_of_old[key(A)] = key(B)
    return new_of_old


@pytest.mark.parametrize("dim,n,pu,pp", [(2, 4, 1, 1), (2, 3, 2, 2), (2, 3, 2, 1), (3, 3, 1, 1),
                                          (3, 2, 2, 2), (3, 2, 2, 1)])
def test_box_mesh_equals_the_oracle_mesh(oracle, dim, n, pu, pp):
    """The C++ stand-in for setup_dofs against the numpy restatement: dof tables, constraints,
    sparsity (keep_constrained_dofs=false), FE tables, geometry; then the Cuthill–McKee numbering
    (node-level CM == scipy's dof-level CM whenever the seed corner is the same)."""
    from softx_2020_200_b200.mesh import BoxMesh
    A = BoxMesh(dim, n, pu, pp, renumber=False, with_q_points=True)
    O = oracle.BoxMesh(dim, n, pu, pp, renumber="none")
    assert A.n_dofs == O.ndof and A.n_cells == O.ncell
    for name, ref in [("cell_dofs", O.cell_dofs.ravel()), ("row_ptr", O.rowptr), ("col_idx", O.col),
                      ("constrained", O.constrained), ("constraint_values", O.constraint_value),
                      ("dof_component", O.dof_comp)]:
        assert np.array_equal(A.array(name), ref), name
    for name, ref in [("shape_u", O.fe.Nu), ("grad_u", O.fe.dNu), ("hess_u", O.fe.d2Nu),
                      ("shape_p", O.fe.Np), ("grad_p", O.fe.dNp), ("weights", O.fe.wq),
                      ("inv_jacobian", O.cell_invJ), ("det_jacobian", O.cell_detJ),
                      ("q_points", O.qpoints), ("dof_coords", O.dof_coords)]:
        assert np.allclose(A.array(name), ref.ravel(), rtol=0, atol=2e-14), name
    B = BoxMesh(dim, n, pu, pp, renumber=True)
    new_of_old = _match_numbering(A, B, dim)
    O2 = oracle.BoxMesh(dim, n, pu, pp, renumber=new_of_old)
    for name, ref in [("cell_dofs", O2.cell_dofs.ravel()), ("row_ptr", O2.rowptr),
                      ("col_idx", O2.col), ("constrained", O2.constrained)]:
        assert np.array_equal(B.array(name), ref), name
    O3 = oracle.BoxMesh(dim, n, pu, pp, renumber="cm")
    if int(np.argmin(O3.new_of_old)) == int(np.argmin(new_of_old)):      # same seed corner
        assert np.array_equal(O3.new_of_old, new_of_old)
    # colouring: cells of one colour share no dof
    ptr, cells, cd = B.array("color_ptr"), B.array("color_cells"), B.array("cell_dofs").reshape(B.n_cells, -1)
    for c in range(len(ptr) - 1):
        d = cd[cells[ptr[c]:ptr[c + 1]]].ravel()
        assert len(np.unique(d)) == d.size


@pytest.mark.parametrize("dim,n,lo,hi,pu,pp", [(2, (5, 3), (0.0, 0.0), (10.0, 1.0), 1, 1),
                                                (2, (3, 4), (-1.0, 0.5), (2.0, 1.5), 2, 1),
                                                (3, (2, 3, 4), (0.0, 0.0, 0.0), (1.0, 2.0, 0.5), 2, 2)])
def test_anisotropic_box_mesh_equals_the_oracle_mesh(oracle, dim, n, lo, hi, pu, pp):
    """Boxes with different cell counts and extents per direction (subdivided_hyper_rectangle-like;
    the channel of poiseuille_gls.prm is one): the C++ stand-in against the numpy restatement."""
    from softx_2020_200_b200.mesh import BoxMesh
    A = BoxMesh(dim, n, pu, pp, lo=lo, hi=hi, renumber=False, with_q_points=True)
    O = oracle.BoxMesh(dim, n, pu, pp, lo=lo, hi=hi, renumber="none")
    assert A.n_dofs == O.ndof and A.n_cells == O.ncell
    for name, ref in [("cell_dofs", O.cell_dofs.ravel()), ("row_ptr", O.rowptr), ("col_idx", O.col),
                      ("constrained", O.constrained), ("dof_component", O.dof_comp)]:
        assert np.array_equal(A.array(name), ref), name
    for name, ref in [("inv_jacobian", O.cell_invJ), ("det_jacobian", O.cell_detJ),
                      ("cell_measure", O.cell_measure), ("q_points", O.qpoints),
                      ("dof_coords", O.dof_coords)]:
        assert np.allclose(A.array(name), ref.ravel(), rtol=0, atol=2e-14), name
    B = BoxMesh(dim, n, pu, pp, lo=lo, hi=hi, renumber=True)
    O2 = oracle.BoxMesh(dim, n, pu, pp, lo=lo, hi=hi, renumber=_match_numbering(A, B, dim))
    for name, ref in [("cell_dofs", O2.cell_dofs.ravel()), ("row_ptr", O2.rowptr),
                      ("col_idx", O2.col), ("constrained", O2.constrained)]:
        assert np.array_equal(B.array(name), ref), name


@pytest.mark.parametrize("dim,n,lo,hi,pu,pp,periodic,walls", [
    (2, 4, 0.0, 6.28318530718, 1, 1, (0, 1), ()),           # taylor-green-vortex_gls_bdf1.prm
    (2, 4, 0.0, 6.28318530718, 2, 1, (0, 1), ()),           # taylor-green-vortex_gls_sdirk*.prm
    (2, (5, 3), (0.0, 0.0), (10.0, 1.0), 1, 1, (0,), (2, 3)),  # poiseuille_gls.prm (odd count)
    (3, (2, 3, 2), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), 2, 2, (2,), (0, 1, 2, 3))])
def test_periodic_box_mesh_equals_the_oracle_mesh(oracle, dim, n, lo, hi, pu, pp, periodic, walls):
    """`type = periodic` boundary pairs in the C++ stand-in (glsnsh_mesh_make_periodic) against the
    numpy restatement: identified cell -> dof table, slave dofs as constrained rows, sparsity; the
    colouring stays conflict free across the wrap (odd cell counts take a third colour)."""
    from softx_2020_200_b200.mesh import BoxMesh
    bcs = [(f, "noslip") for f in walls]
    obcs = {f: ("noslip",) for f in walls}
    A = BoxMesh(dim, n, pu, pp, lo=lo, hi=hi, bcs=bcs, renumber=False, periodic=periodic)
    O = oracle.BoxMesh(dim, n, pu, pp, lo=lo, hi=hi, bcs=obcs, renumber="none", periodic=periodic)
    assert A.n_dofs == O.ndof and A.n_cells == O.ncell
    for name, ref in [("cell_dofs", O.cell_dofs.ravel()), ("row_ptr", O.rowptr), ("col_idx", O.col),
                      ("constrained", O.constrained), ("periodic_slave", O.periodic_slave),
                      ("periodic_master", O.periodic_master)]:
        assert np.array_equal(A.array(name), ref), name
    assert np.all(A.array("constraint_values")[A.array("periodic_slave")] == 0)
    B = BoxMesh(dim, n, pu, pp, lo=lo, hi=hi, bcs=bcs, renumber=True, periodic=periodic)
    nat = BoxMesh(dim, n, pu, pp, lo=lo, hi=hi, bcs=bcs, renumber=False)
    full = BoxMesh(dim, n, pu, pp, lo=lo, hi=hi, bcs=bcs, renumber=True)
    O2 = oracle.BoxMesh(dim, n, pu, pp, lo=lo, hi=hi, bcs=obcs, renumber=_match_numbering(nat, full, dim),
                        periodic=periodic)
    for name, ref in [("cell_dofs", O2.cell_dofs.ravel()), ("row_ptr", O2.rowptr), ("col_idx", O2.col),
                      ("constrained", O2.constrained)]:
        assert np.array_equal(B.array(name), ref), name
    ptr, cells, cd = B.array("color_ptr"), B.array("color_cells"), B.array("cell_dofs").reshape(B.n_cells, -1)
    assert ptr[-1] == B.n_cells
    for c in range(len(ptr) - 1):
        d = cd[cells[ptr[c]:ptr[c + 1]]].ravel()
        assert len(np.unique(d)) == d.size


def test_cavity_boundary_conditions_first_listed_wins():
    from softx_2020_200_b200.mesh import BoxMesh
    bcs = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"),
           (3, "function", (1.0, 0.0, 0.0))]
    m = BoxMesh(3, 2, 2, 2, bcs=bcs)
    xyz = m.array("dof_coords").reshape(-1, 3)
    comp, con, val = m.array("dof_component"), m.array("constrained"), m.array("constraint_values")
    lid_interior = (np.abs(xyz[:, 1] - 1) < 1e-12) & (np.abs(xyz[:, 0]) < 1 - 1e-12) & \
        (np.abs(xyz[:, 2]) < 1 - 1e-12)
    assert np.all(val[lid_interior & (comp == 0)] == 1.0) and np.all(con[lid_interior & (comp < 3)] == 1)
    lid_edge = (np.abs(xyz[:, 1] - 1) < 1e-12) & (np.abs(np.abs(xyz[:, 0]) - 1) < 1e-12)
    assert np.all(val[lid_edge] == 0.0)          # walls are listed before the lid
    assert np.all(con[comp == 3] == 0)           # pressure is never constrained


@pytest.mark.parametrize("dim,n,pu,pp,fill", [(2, 8, 1, 1, 1), (2, 8, 1, 1, 4), (2, 4, 2, 2, 2),
                                               (3, 3, 1, 1, 1), (3, 2, 2, 2, 1), (2, 6, 2, 1, 0)])
def test_level_of_fill_pattern_equals_the_oracle(oracle, dim, n, pu, pp, fill):
    """The symbolic phase of `ilu preconditioner fill = k` (host code of libglsns.so, no device):
    the pattern glsns_setup_ilu installs against the oracle's independent restatement of Ifpack's
    level-of-fill graph."""
    import ctypes as C
    from softx_2020_200_b200 import _lib
    L = _lib.lib()
    L.glsnsh_iluk_pattern.restype = C.c_int64
    L.glsnsh_iluk_pattern.argtypes = [C.c_int64, _lib.c_i64_p, _lib.c_i32_p, C.c_int32, _lib.c_i64_p,
                                      _lib.c_i32_p]
    mesh = oracle.BoxMesh(dim, n, pu, pp)
    rp = np.ascontiguousarray(mesh.rowptr, dtype=np.int64)
    col = np.ascontiguousarray(mesh.col, dtype=np.int32)
    p = lambda a, t: a.ctypes.data_as(t)
    nnz = L.glsnsh_iluk_pattern(mesh.ndof, p(rp, _lib.c_i64_p), p(col, _lib.c_i32_p), fill, None, None)
    orp, ocol = np.empty(mesh.ndof + 1, dtype=np.int64), np.empty(nnz, dtype=np.int32)
    assert L.glsnsh_iluk_pattern(mesh.ndof, p(rp, _lib.c_i64_p), p(col, _lib.c_i32_p), fill,
                                 p(orp, _lib.c_i64_p), p(ocol, _lib.c_i32_p)) == nnz
    pm, _ = oracle.iluk_pattern(mesh, fill)
    assert np.array_equal(orp, pm.rowptr) and np.array_equal(ocol, pm.col)
    assert L.glsnsh_iluk_pattern(mesh.ndof, p(rp, _lib.c_i64_p), p(col, _lib.c_i32_p), -1, None, None) == -1


def _time_stepping(method, n_steps, dt=0.1, scaling=0.4):
    from softx_2020_200_b200 import _lib
    from tests.mirror.solver import mirror_lib
    L = mirror_lib()
    L.glsnsh_time_stepping_trace.restype = C.c_int
    L.glsnsh_time_stepping_trace.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double,
                                             C.POINTER(C.c_double), C.c_int]
    out = (C.c_double * 900)()
    n = L.glsnsh_time_stepping_trace(_lib.SCHEMES[method], n_steps, dt, scaling, out, 900)
    names = {v: k for k, v in _lib.SCHEMES.items()}
    return [(names[int(out[9 * i])], bool(out[9 * i + 1]), bool(out[9 * i + 2]),
             tuple(round(out[9 * i + j], 12) for j in (3, 4, 5)),
             tuple(int(out[9 * i + j]) for j in (6, 7, 8))) for i in range(n)]


def test_time_stepping_glue_of_navier_stokes_base():
    """NavierStokesBase::first_iteration / iterate / finish_time_step
    (source/solvers/navier_stokes_base.cc:428-590) of the mirror on a recording solver: which
    method each solve runs with, the time-step vector it sees (newest first) and which earlier
    solutions sit in solution_m1 / m2 / m3 (tag 0 = the initial condition, tag k = result of the
    k-th solve)."""
    # sdirk3: three stage solves per step, stage results in m2 / m3, m1 = previous step
    calls = _time_stepping("sdirk3", 2)
    assert [c[0] for c in calls] == ["sdirk3_1", "sdirk3_2", "sdirk3_3"] * 2
    assert [c[4] for c in calls[:3]] == [(0, 0, 0), (0, 1, 0), (0, 1, 2)]
    assert [c[4][0] for c in calls[3:]] == [3, 3, 3] and calls[5][4] == (3, 4, 5)
    assert all(c[1:3] == (False, False) and c[3][0] == 0.1 for c in calls)
    # sdirk2
    calls = _time_stepping("sdirk2", 1)
    assert [c[0] for c in calls] == ["sdirk2_1", "sdirk2_2"] and calls[1][4][:2] == (0, 1)
    # bdf1 / steady: one solve per step
    assert [c[0] for c in _time_stepping("bdf1", 3)] == ["bdf1"] * 3
    assert [c[4][0] for c in _time_stepping("bdf1", 3)] == [0, 1, 2]
    # bdf2 starts with an Euler step of 0.4 dt and completes the step with a bdf2 solve of 0.6 dt
    calls = _time_stepping("bdf2", 2)
    assert [c[0] for c in calls] == ["bdf1", "bdf2", "bdf2"]
    assert calls[0][3][0] == 0.04 and calls[1][3][:2] == (0.06, 0.04) and calls[2][3][:2] == (0.1, 0.06)
    assert calls[0][2] and calls[1][2] and not calls[2][2]          # start-up solves renew the matrix
    assert calls[1][4][:2] == (1, 0) and calls[2][4][:2] == (2, 1)
    # bdf3: two Euler steps of 0.4 dt, then a bdf3 solve of 0.2 dt
    calls = _time_stepping("bdf3", 2)
    assert [c[0] for c in calls] == ["bdf1", "bdf1", "bdf3", "bdf3"]
    assert [c[3][0] for c in calls] == [0.04, 0.04, 0.02, 0.1]
    assert calls[2][4] == (2, 1, 0) and calls[3][4] == (3, 2, 1)


@pytest.mark.parametrize("dim,n,pu,pp,ranks", [(3, (4, 3, 5), 2, 2, 3), (3, 4, 1, 1, 2), (2, (7, 5), 2, 1, 4),
                                               (3, 5, 2, 2, 8), (2, 6, 2, 2, 1)])
def test_rank_local_mesh_equals_the_partition_of_the_global_one(dim, n, pu, pp, ranks):
    """glsnsh_mesh_create_local (what bench.py uses for N > 1 and for the 50 M dof mesh of
    BASELINE.json configs[3], where no rank can hold the global CSR) against glsnsh_mesh_create +
    glsnsh_mesh_partition, array by array."""
    from softx_2020_200_b200.mesh import BoxMesh, _DTYPES
    bcs = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (3, "function", (1.0, 0.5, 0.25))]
    if dim == 3:
        bcs += [(4, "noslip"), (5, "noslip")]
    glob = BoxMesh(dim, n, pu, pp, bcs=bcs, with_q_points=True)
    for rank in range(ranks):
        a = glob.partition(ranks, rank)
        b = BoxMesh(dim, n, pu, pp, bcs=bcs, with_q_points=True, local=(ranks, rank))
        for attr in ("n_dofs", "n_owned", "n_cells", "nnz", "n_colors", "n_neighbors", "n_global",
                     "owned_begin"):
            assert getattr(a, attr) == getattr(b, attr), (attr, rank)
        for name in _DTYPES:
            if name in ("periodic_slave", "periodic_master"):
                continue
            assert np.array_equal(a.array(name), b.array(name)), (name, rank)


@pytest.mark.parametrize("n,world", [(3, 3), (4, 2)])
def test_rank_local_arrays_reproduce_the_block_jacobi_oracle(oracle, n, world):
    """The reference side of tests/test_gpu_parity.py::test_rank_local_blocks_on_one_gpu, on the
    CPU: a rank's local arrays (owned rows, sorted local columns with the ghosts behind the owned
    ones) carry exactly the oracle's rows of its block, and the oracle's SpMV / triangular sweeps
    over those local arrays equal its block-Jacobi results on the global matrix."""
    import types
    from tests.util import rank_local_reference
    ref = rank_local_reference(oracle, n, world)
    om, bp, x = ref["oracle_mesh"], ref["block_ptr"], ref["x"]
    assert bp[-1] == om.ndof
    for r, m in enumerate(ref["parts"]):
        l2g, rp, col = m.array("local_to_global"), m.array("row_ptr"), m.array("col_idx")
        assert np.array_equal(l2g[:m.n_owned], np.arange(bp[r], bp[r + 1]))
        rows, col_g = slice(bp[r], bp[r + 1]), l2g[col]
        a_loc, lu_loc = np.empty(m.nnz), np.empty(m.nnz)
        for i in range(m.n_owned):
            o = np.argsort(col_g[rp[i]:rp[i + 1]])
            gr = slice(om.rowptr[bp[r] + i], om.rowptr[bp[r] + i + 1])
            assert np.all(np.diff(col[rp[i]:rp[i + 1]]) > 0)
            assert np.array_equal(col_g[rp[i]:rp[i + 1]][o], om.col[gr])
            a_loc[rp[i]:rp[i + 1]][o] = ref["matrix"][gr]
            lu_loc[rp[i]:rp[i + 1]][o] = ref["ilu"][gr]
        loc = types.SimpleNamespace(ndof=m.n_owned, rowptr=rp, col=col)
        # (the ghosts sit behind the owned columns: another summation order than the global row's)
        assert np.max(np.abs(oracle.spmv(loc, a_loc, x[l2g]) - ref["spmv"][rows])) <= 1e-13 * np.max(np.abs(ref["spmv"]))
        dp = np.array([rp[i] + np.searchsorted(col[rp[i]:rp[i + 1]], i) for i in range(m.n_owned)],
                      dtype=np.int64)
        z = oracle.ilu_apply(loc, lu_loc, dp, x[rows])
        assert np.max(np.abs(z - ref["ilu_apply"][rows])) <= 1e-13 * np.max(np.abs(ref["ilu_apply"]))
    ref["global_mesh"].close()
