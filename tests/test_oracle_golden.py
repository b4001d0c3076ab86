"""Pins the CPU oracle against the reference's own golden files.

The expected numbers below are copied from the reference's checked-in test outputs; the inputs
(mesh, forcing, solver settings) are those of the corresponding reference test."""
import json
import os

import numpy as np
import pytest

from tests import mms

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN_DIR, "reference_golden.json")) as _f:
    REF = json.load(_f)      # the reference's own numbers, file:line in each entry's "source"


def test_bdf_coefficients_bdf_01_output(oracle):
    # tests/core/bdf_01.output: dt = 0.1, 0.2, 0.3
    dts = [0.1, 0.2, 0.3]
    assert np.allclose(oracle.bdf_coefficients(1, dts), [10.0, -10.0], rtol=0, atol=5e-5)
    assert np.allclose(oracle.bdf_coefficients(2, dts), [13.3333, -15.0, 1.66667], rtol=0, atol=5e-5)
    assert np.allclose(oracle.bdf_coefficients(3, dts), [15.0, -18.0, 3.33333, -0.333333],
                       rtol=0, atol=5e-6)


def test_restart_01_output(oracle):
    """tests/solvers/restart_01.output:2-8 — GMRES(30)+ILU(0) iteration count and true residual per
    Newton step, then the velocity L2 error. 2D Q1-Q1, hyper_cube(-1,1) refined 4x, no-slip, nu=1,
    all .prm defaults (Newton 1e-6/10; GMRES rel 1e-3, abs 1e-8, 1000; ILU 0, 1e-8, 1)."""
    m = oracle.BoxMesh(2, 16, 1, 1)
    assert m.ndof == 867
    f = m.evaluate_force(mms.forcing_2d)
    pr = oracle.scheme_params("steady", [1.0], 1.0)
    log = []
    U, it, res = oracle.newton_solve(m, np.zeros(m.ndof), pr, f, log=log)
    assert [k for k, _ in log] == REF["restart_01"]["gmres_iterations"] == [8, 6, 10]
    golden = REF["restart_01"]["true_residuals"]
    for (_, r), g in zip(log, golden):
        assert float("%.6g" % r) == g          # every printed digit
    err_u, _ = oracle.l2_error(m, U, mms.exact_2d)
    assert float("%.6g" % err_u) == REF["restart_01"]["l2_error_velocity"] == 0.0343628
    # "Error after zeroing the solution: 0.612372" = ||u_exact||
    assert float("%.6g" % oracle.l2_error(m, np.zeros(m.ndof), mms.exact_2d)[0]) == 0.612372


@pytest.mark.parametrize("n,ndof,eu,ep", [(4, 500, 5.4021e-01, 4.5537e-02),
                                          (8, 2916, 1.3126e-01, 1.7717e-01)])
def test_mms3d_gls_output(oracle, n, ndof, eu, ep):
    """applications_tests/gls_navier_stokes_3d/mms3d_gls.output:13-18 (3D Q1-Q1, nu=1, Newton 1e-8).
    The reference solves the linear systems with AMG; the converged Newton solution is
    solver-independent, so GMRES+ILU(0) with that .prm's tolerances is used here."""
    m = oracle.BoxMesh(3, n, 1, 1)
    assert m.ndof == ndof and m.ncell == n ** 3
    f = m.evaluate_force(mms.forcing_3d)
    pr = oracle.scheme_params("steady", [1.0], 1.0)
    U, it, res = oracle.newton_solve(m, np.zeros(m.ndof), pr, f, tol=1e-8,
                                     lin=dict(rel=1e-4, abs_=1e-9, max_iters=5000, ilu_atol=1e-10))
    assert res < 1e-8
    err_u, err_p = oracle.l2_error(m, U, mms.exact_3d)
    assert float("%.5g" % err_u) == eu
    assert float("%.5g" % err_p) == ep


@pytest.mark.parametrize("n,ndof,eu,ep", [(8, 243, 1.3284e-01, 1.7844e-01),
                                          (16, 867, 3.4363e-02, 9.7118e-02),
                                          (32, 3267, 8.7362e-03, 3.0300e-02)])
def test_mms2d_gls_output(oracle, n, ndof, eu, ep):
    """applications_tests/gls_navier_stokes_2d/mms2d_gls.output:24-26. The reference uses ILU(4) as
    the preconditioner; the Newton-converged solution does not depend on it."""
    m = oracle.BoxMesh(2, n, 1, 1)
    assert m.ndof == ndof
    f = m.evaluate_force(mms.forcing_mms2d)
    pr = oracle.scheme_params("steady", [1.0], 1.0)
    U, it, res = oracle.newton_solve(m, np.zeros(m.ndof), pr, f, tol=1e-8,
                                     lin=dict(rel=1e-4, abs_=1e-9, max_iters=5000))
    err_u, err_p = oracle.l2_error(m, U, mms.exact_mms2d)
    assert float("%.5g" % err_u) == eu
    assert float("%.5g" % err_p) == ep


@pytest.mark.parametrize("name", ["case_2d_q2q1_bdf2", "case_3d_q1q1_steady", "case_3d_q2q2_steady"])
def test_oracle_reproduces_committed_fixtures(oracle, name):
    """tests/golden/case_*.npz (made by tests/golden/make_fixtures.py) are what the GPU parity tests
    meet; the oracle must still produce them bit for bit (same compiler flags: no FMA contraction)."""
    fx = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    mesh = oracle.BoxMesh(int(fx["dim"]), int(fx["n"]), int(fx["pu"]), int(fx["pp"]))
    assert np.array_equal(mesh.rowptr, fx["row_ptr"]) and np.array_equal(mesh.col, fx["col_idx"])
    dts = list(fx["dts"]) or None
    pr = oracle.scheme_params(str(fx["scheme"]), dts, float(fx["viscosity"]))
    U1 = fx["U1"] if fx["U1"].size else None
    U2 = fx["U2"] if fx["U2"].size else None
    val, rhs = oracle.assemble(mesh, fx["U"], pr, True, None, U1, U2, None)
    assert np.array_equal(val, fx["matrix"]) and np.array_equal(rhs, fx["rhs"])
    lu, dp = oracle.ilu0(mesh, val, 1e-8, 1.0)
    assert np.array_equal(lu, fx["ilu"])
    assert np.array_equal(oracle.ilu_apply(mesh, lu, dp, fx["x"]), fx["ilu_apply"])
    assert np.array_equal(oracle.spmv(mesh, val, fx["x"]), fx["spmv"])


def test_oracle_bicgstab_agrees_with_pinned_gmres(oracle):
    """solve_system_BiCGStab has no golden vector in the reference (no test uses it): the oracle's
    restatement is checked against the pinned GMRES path on restart_01's first Newton system."""
    from tests import mms
    mesh = oracle.BoxMesh(2, 16, 1, 1)
    force = mesh.evaluate_force(mms.forcing_2d)
    val, rhs = oracle.assemble(mesh, np.zeros(mesh.ndof), oracle.scheme_params("steady", None, 1.0),
                               True, force)
    xg, itg, _ = oracle.solve_linear_system(mesh, val, rhs, rel=1e-10, abs_=1e-14)
    xb, itb, res = oracle.solve_linear_system(mesh, val, rhs, rel=1e-10, abs_=1e-14, method="bicgstab")
    assert res < 1e-9 * np.linalg.norm(rhs)
    assert np.linalg.norm(xb - xg) <= 1e-7 * np.linalg.norm(xg)
    assert itb < itg            # two products per iteration
    with pytest.raises(RuntimeError, match="This solver is not allowed"):
        oracle.solve_linear_system(mesh, val, rhs, method="amg")


def test_oracle_ilu_fill_levels(oracle):
    """ILU(k) restatement: k = 0 is the matrix pattern, the pattern grows with k up to the one of the
    complete LU factorisation, and there the factors ARE the (pivot-free) LU factors."""
    from tests import mms
    mesh = oracle.BoxMesh(2, 6, 1, 1)
    val, _ = oracle.assemble(mesh, np.zeros(mesh.ndof), oracle.scheme_params("steady", None, 1.0),
                             True, mesh.evaluate_force(mms.forcing_2d))
    p0, a2p0 = oracle.iluk_pattern(mesh, 0)
    assert np.array_equal(p0.rowptr, mesh.rowptr) and np.array_equal(p0.col, mesh.col)
    assert np.array_equal(a2p0, np.arange(mesh.rowptr[-1]))
    nnz = [int(oracle.iluk_pattern(mesh, k)[0].rowptr[-1]) for k in (0, 1, 2, 4, 200)]
    assert nnz == sorted(nnz) and nnz[1] > nnz[0]
    pm, a2p = oracle.iluk_pattern(mesh, 200)
    lu, dp = oracle.ilu0(pm, oracle.pad_values(pm, a2p, val), 0.0, 1.0)
    n = mesh.ndof
    D = np.zeros((n, n))
    rows = np.repeat(np.arange(n), np.diff(mesh.rowptr))
    D[rows, mesh.col] = val
    for i in range(n):                       # dense IKJ elimination without pivoting
        for k in range(i):
            if D[i, k] != 0.0:
                D[i, k] /= D[k, k]
                D[i, k + 1:] -= D[i, k] * D[k, k + 1:]
    assert np.count_nonzero(D) == nnz[-1]
    prow = np.repeat(np.arange(n), np.diff(pm.rowptr))
    ref = D[prow, pm.col]
    assert np.max(np.abs(lu - ref)) <= 1e-9 * np.max(np.abs(ref))
    # block-Jacobi: the fill never crosses a block boundary
    bp = np.array([0, n // 2, n])
    pb, _ = oracle.iluk_pattern(mesh, 2, bp)
    brow = np.repeat(np.arange(n), np.diff(pb.rowptr))
    base = set(zip(rows.tolist(), mesh.col.tolist()))
    for i, j in zip(brow.tolist(), pb.col.tolist()):
        assert (i, j) in base or (i < n // 2) == (j < n // 2)


def _linear_field(dim):
    """A field inside every FE_Q(p >= 1) space: its L2 projection is its nodal interpolant."""
    def f(x):
        out = np.zeros((len(x), dim + 1))
        for c in range(dim + 1):
            out[:, c] = 0.3 * (c + 1) + sum((0.2 + 0.1 * c + 0.05 * d) * x[:, d] for d in range(dim))
        return out
    return f


@pytest.mark.parametrize("dim,n,pu,pp", [(2, 4, 1, 1), (2, 3, 2, 2), (2, 3, 2, 1), (3, 2, 2, 2)])
def test_oracle_l2_projection_and_cfl(oracle, dim, n, pu, pp):
    """assemble_L2_projection + set_initial_condition(L2projection): a field of the FE space comes
    back as its nodal values (constrained dofs carry the constraint values); calculate_CFL of a
    uniform velocity is |u| dt / h with h from the cell measure."""
    f = _linear_field(dim)
    lid = lambda x: f(x)[:, :dim]                     # Dirichlet data consistent with the field
    bcs = {b: ("function", lid) for b in range(2 * dim)}
    mesh = oracle.BoxMesh(dim, n, pu, pp, bcs=bcs)
    U, it, ok = oracle.l2_projection(mesh, f, rel=1e-13, abs_=1e-14)
    assert ok
    exact = f(mesh.dof_coords)[np.arange(mesh.ndof), mesh.dof_comp]
    assert np.max(np.abs(U - exact)) <= 1e-10
    # uniform velocity (1, 2[, 2]): CFL = |u| dt / h
    V = np.zeros(mesh.ndof)
    vel = np.array([1.0, 2.0, 2.0])[:dim]
    for c in range(dim):
        V[mesh.dof_comp == c] = vel[c]
    meas = mesh.cell_measure[0]
    h = (np.sqrt(4 * meas / np.pi) if dim == 2 else (6 * meas / np.pi) ** (1 / 3)) / max(pu, pp)
    assert abs(oracle.calculate_cfl(mesh, V, 0.01) - np.linalg.norm(vel) * 0.01 / h) <= 1e-14


# ---- Taylor-Green vortex: the reference's transient goldens -----------------------------------
TWO_PI = 6.28318530718      # `set grid arguments = 0 : 6.28318530718 : true`


def _taylor_green(t, nu=1.0):
    """initial conditions / analytical solution of taylor-green-vortex_gls_*.prm (:33-36, :44-48)."""
    def f(x):
        e = np.exp(-2 * nu * t)
        return np.stack([e * np.cos(x[:, 0]) * np.sin(x[:, 1]), -e * np.sin(x[:, 0]) * np.cos(x[:, 1]),
                         -0.25 * (np.cos(2 * x[:, 0]) + np.cos(2 * x[:, 1]))], axis=1)
    return f


TG_LIN = dict(rel=1e-4, abs_=1e-9, max_iters=5000, ilu_atol=1e-5)   # linear solver :108-116 (ILU(0)
# here instead of the file's fill 1 on 2 ranks: only the path to the Newton-converged state differs)


@pytest.mark.parametrize("order", [3, 2])
def test_taylor_green_vortex_sdirk_output(oracle, order):
    """applications_tests/gls_navier_stokes_2d/taylor-green-vortex_gls_sdirk{3,2}.mpirun=2.output:
    4096 periodic Q2-Q1 cells, 37507 dofs, L2-projected initial condition, ONE sdirk step of 0.1.
    Pins, digit for digit as printed: calculate_CFL of the projected field (1.80106), enstrophy and
    kinetic energy before (0.5, 0.25) and after the step, and the velocity L2 error to the three
    digits the solver tolerances leave stable (tightening GMRES / Newton moves the 4th:
    1.38187e-4 ... 1.38256e-4 around the reference's 1.38223e-4)."""
    g = REF["taylor_green_vortex_sdirk%d" % order]
    mesh = oracle.BoxMesh(2, 64, 2, 1, lo=0.0, hi=TWO_PI, bcs={}, periodic=(0, 1))
    assert (mesh.ncell, mesh.ndof) == (g["cells"], g["dofs"])
    U0, _, ok = oracle.l2_projection(mesh, _taylor_green(0.0), ilu_atol=1e-5)
    assert ok
    assert "%.6g" % oracle.enstrophy(mesh, U0) == g["enstrophy_0"]
    assert "%.6g" % oracle.kinetic_energy(mesh, U0) == g["kinetic_energy_0"]
    assert "%.6g" % oracle.calculate_cfl(mesh, U0, 0.1) == g["cfl"][0]
    U1 = oracle.sdirk_step(mesh, order, U0, 0.1, 1.0, None, tol=1e-6, max_it=5, lin=TG_LIN)
    assert "%.6g" % oracle.enstrophy(mesh, U1) == g["enstrophy"][0]
    assert "%.6g" % oracle.kinetic_energy(mesh, U1) == g["kinetic_energy"][0]
    err = oracle.l2_error(mesh, U1, _taylor_green(0.1))[0]
    assert abs(err - float(g["l2_error_velocity"][0])) <= 5e-4 * err


def test_taylor_green_vortex_bdf1_output(oracle):
    """taylor-green-vortex_gls_bdf1.mpirun=2.output: 1024 periodic Q1-Q1 cells, 3267 dofs, 100 BDF1
    steps of 0.01.  Every printed CFL number, enstrophy, kinetic energy and velocity L2 error of the
    100 steps, as printed (6 significant digits; the error table's 5)."""
    g = REF["taylor_green_vortex_bdf1"]
    mesh = oracle.BoxMesh(2, 32, 1, 1, lo=0.0, hi=TWO_PI, bcs={}, periodic=(0, 1))
    assert (mesh.ncell, mesh.ndof) == (g["cells"], g["dofs"])
    U, _, ok = oracle.l2_projection(mesh, _taylor_green(0.0), ilu_atol=1e-5)
    assert ok
    assert "%.6g" % oracle.enstrophy(mesh, U) == g["enstrophy_0"]
    assert "%.6g" % oracle.kinetic_energy(mesh, U) == g["kinetic_energy_0"]
    pr = oracle.scheme_params("bdf1", [0.01] * 4, 1.0)
    exact = {k: 0 for k in ("cfl", "enstrophy", "kinetic_energy", "l2_error_velocity", "error_table")}
    for k in range(100):
        cfl = oracle.calculate_cfl(mesh, U, 0.01)
        U, _, _ = oracle.newton_solve(mesh, U, pr, None, tol=1e-6, max_it=5, lin=TG_LIN,
                                      hist=(U, None, None))
        vals = dict(cfl=cfl, enstrophy=oracle.enstrophy(mesh, U),
                    kinetic_energy=oracle.kinetic_energy(mesh, U),
                    l2_error_velocity=oracle.l2_error(mesh, U, _taylor_green(0.01 * (k + 1)))[0])
        for name, v in vals.items():
            ref = float(g[name][k])
            assert abs(v - ref) <= 2e-5 * abs(ref), (name, k, v, ref)
            exact[name] += "%.6g" % v == g[name][k]
        t, e = g["error_table"][k]
        assert t == "%.4f" % (0.01 * (k + 1))
        exact["error_table"] += "%.4e" % vals["l2_error_velocity"] == e
    # digit for digit, up to the rare last-digit flip that the solver tolerances allow
    assert all(c >= 95 for c in exact.values()), exact


def test_poiseuille_gls_output(oracle):
    """applications_tests/gls_navier_stokes_2d/poiseuille_gls.output:24-26: plane Poiseuille flow
    driven by the source term (1, 0) in the x-periodic channel [0,10] x [0,1] (2d_channel.msh is the
    structured 49 x 9 mesh, refined uniformly twice), Q1-Q1, nu = 1: cells, dofs (deal.II counts the
    periodic duplicates) and the velocity error against u = y (1 - y) / 2, as printed.  The exact
    pressure is 0: the printed pressure errors are solver noise (1e-9), ours must be too."""
    g = REF["poiseuille_gls"]

    def exact(x):
        return np.stack([0.5 * x[:, 1] * (1 - x[:, 1]), 0 * x[:, 0], 0 * x[:, 0]], axis=1)
    force = lambda x: np.stack([np.ones(len(x)), 0 * x[:, 0], 0 * x[:, 0]], axis=1)
    lin = dict(rel=1e-4, abs_=1e-9, max_iters=5000, ilu_atol=1e-5)          # poiseuille_gls.prm:108-116
    for k in range(3):
        mesh = oracle.BoxMesh(2, (49 * 2 ** k, 9 * 2 ** k), 1, 1, lo=(0.0, 0.0), hi=(10.0, 1.0),
                              bcs={2: ("noslip",), 3: ("noslip",)}, periodic=(0,))
        assert (mesh.ncell, mesh.ndof) == (g["cells"][k], g["dofs"][k])
        U, _, res = oracle.newton_solve(mesh, mesh.apply_nonzero_constraints(np.zeros(mesh.ndof)),
                                        oracle.scheme_params("steady", None, 1.0),
                                        mesh.evaluate_force(force), tol=1e-6, max_it=3, lin=lin)
        eu, ep = oracle.l2_error(mesh, U, exact)
        assert "%.4e" % eu == g["error_velocity"][k]
        assert ep < 1e-6


@pytest.mark.parametrize("dim,n,pu,pp,scheme,srf", [
    (2, 3, 1, 1, "steady", False), (2, 3, 2, 1, "bdf2", False), (3, 2, 2, 2, "steady", False),
    (3, 2, 2, 2, "sdirk3_2", True), (2, 3, 2, 2, "bdf1", True), (3, 2, 1, 1, "bdf3", False)])
def test_structured_cell_kernel_equals_the_literal_loop(oracle, dim, n, pu, pp, scheme, srf):
    """the structured block form of the local matrix (cell_structured, the
    "Mode B" CPU baseline of bench.py and the form the CUDA kernel computes) against the literal
    q x j x i loop of gls_navier_stokes.cc:387-748 on random states, entry by entry."""
    from tests import mms
    mesh = oracle.BoxMesh(dim, n, pu, pp)
    rng = np.random.default_rng(5)
    U, U1, U2, U3 = (rng.uniform(-1, 1, mesh.ndof) for _ in range(4))
    force = mesh.evaluate_force(mms.forcing_2d if dim == 2 else mms.forcing_3d)
    pr = oracle.scheme_params(scheme, None if scheme == "steady" else [0.1, 0.2, 0.3], 0.05, srf,
                              (0.3, -0.4, 1.1))
    _, _, m_lit, b_lit = oracle.assemble(mesh, U, pr, True, force, U1, U2, U3, return_local=True)
    _, _, m_str, b_str = oracle.assemble(mesh, U, pr, True, force, U1, U2, U3, return_local=True,
                                         structured=True)
    assert np.max(np.abs(m_str - m_lit)) <= 1e-13 * np.max(np.abs(m_lit))
    assert np.max(np.abs(b_str - b_lit)) <= 1e-13 * np.max(np.abs(b_lit))


def test_taylor_couette_golden_curved_q2_cells(oracle):
    """applications_tests/gls_navier_stokes_2d/taylorcouette_gls.output:45-47 -- the reference's only
    pin of the Q2 Laplacian terms on curved (MappingQ) cells: velocity L2 errors 7.7383e-04,
    1.1452e-04, 1.5173e-05 at 64 / 256 / 1024 cells, and the straight-sided volumes it prints.
    (Newton-converged to 1e-10, so independent of the reference's AMG solver.)"""
    from tests import mms
    g = REF["taylorcouette_gls"]
    pr = oracle.scheme_params("steady", None, 1.0)
    for level, (cells, err) in enumerate(zip(g["cells"], g["error_velocity"])):
        mesh = mms.couette_mesh(oracle, level + 2)
        assert mesh.ncell == cells
        # deal.II's shell has no duplicated nodes across theta = 0; here they stay as identified rows
        assert mesh.ndof - mesh.periodic_slave.size == g["dofs"][level]
        if level:
            assert "%.6g" % mesh.cell_measure.sum() == g["volume_q1"][level - 1]
        U0 = mesh.apply_nonzero_constraints(np.zeros(mesh.ndof))
        U, it, res = oracle.newton_solve(mesh, U0, pr, None, tol=1e-10, max_it=10,
                                         lin=dict(rel=1e-8, abs_=1e-13, max_iters=4000, ilu_atol=1e-10))
        assert res <= 1e-10
        err_u, _ = oracle.l2_error(mesh, U, mms.couette_exact)
        assert "%.4e" % err_u == err, (cells, err_u, err)


def test_mapping_laplacian_formula_against_finite_differences(oracle):
    """lap N = H_ref(N) : (J^-1 J^-T) - grad_x N . c with c_k = sum_rs d2x_k/dxi_r dxi_s (J^-1 J^-T)_rs
    (include/glsns.h: mapping_laplacian) against a finite-difference Laplacian in real space of
    x -> N(xi(x)) on a curved Q2 cell of the shell mesh."""
    from tests import mms
    mesh = mms.couette_mesh(oracle, 2)
    fe, c, q = mesh.fe, 5, 4
    X = mesh.cell_X[c]
    K, lapc = mesh.cell_invJ[c, q], mesh.map_lap[c, q]
    grad = fe.dNu[q] @ K                                                  # [ns, dim] real-space gradients
    lap = np.einsum("ars,rs->a", fe.d2Nu[q], K @ K.T) - grad @ lapc

    def shape_at(x):                                                     # N(xi(x)) by Newton on the mapping
        xi = fe.xq[q].copy()
        for _ in range(30):
            N, dN, _ = oracle._tensor_tables(2, 2, [0.5])               # placeholder shapes (overwritten below)
            V0, D0, _ = oracle.lagrange1d(2, [xi[0]])
            V1, D1, _ = oracle.lagrange1d(2, [xi[1]])
            N = np.array([V0[a % 3, 0] * V1[a // 3, 0] for a in range(9)])
            dN = np.array([[D0[a % 3, 0] * V1[a // 3, 0], V0[a % 3, 0] * D1[a // 3, 0]] for a in range(9)])
            r = N @ X - x
            if np.linalg.norm(r) < 1e-15:
                break
            xi -= np.linalg.solve((X.T @ dN), r)
        return N
    x0, h = mesh.qpoints[c, q], 2e-4
    fd = np.zeros(9)
    for d in range(2):
        e = np.zeros(2)
        e[d] = h
        fd += (shape_at(x0 + e) - 2 * shape_at(x0) + shape_at(x0 - e)) / (h * h)
    assert np.max(np.abs(fd - lap)) <= 2e-5 * np.max(np.abs(lap))
    # and the correction is not small on this cell: without it the test above would fail
    assert np.max(np.abs(grad @ lapc)) >= 1e-2 * np.max(np.abs(lap))


@pytest.mark.parametrize("dim,n,pu,pp", [(2, 4, 1, 1), (2, 4, 2, 1), (2, 4, 2, 2), (3, 3, 1, 1), (3, 2, 2, 2)])
def test_hanging_node_patch_test(oracle, dim, n, pu, pp):
    """RefinedBoxMesh (one refined region, hanging nodes on faces and edges): parity unpinned by the
    reference (no reproducible golden with hanging nodes), so the restatement of
    make_hanging_node_constraints + distribute_local_to_global + distribute is checked by a patch
    test -- a flow that lies in the finite element space solves the discrete stabilised equations
    exactly on the non-conforming mesh (the GLS residual terms vanish for it)."""
    nu = 0.7
    if pu == 1:
        shear = (lambda x: x[:, 1]) if dim == 2 else (lambda x: x[:, 1] + 0.5 * x[:, 2])
    else:
        shear = lambda x: 1 - x[:, 1] ** 2
    bc = lambda x: np.stack([shear(x)] + [0 * x[:, 0]] * (dim - 1), axis=1)
    refine = (lambda c: (c[:, 0] > 0) & (c[:, 1] < 0.5)) if dim == 2 else \
        (lambda c: (c[:, 0] > 0) & (c[:, 1] < 0.3) & (c[:, 2] > -0.4))
    mesh = oracle.RefinedBoxMesh(dim, n, pu, pp, refine, bcs={None: ("function", bc)})
    assert (mesh.constrained == 2).sum() > 0
    # every hanging line interpolates: its weights sum to 1 once the Dirichlet masters' share is counted
    U0 = mesh.apply_nonzero_constraints(np.zeros(mesh.ndof))
    U, it, res = oracle.newton_solve(mesh, U0, oracle.scheme_params("steady", None, nu), None, tol=1e-11,
                                     max_it=15, lin=dict(rel=1e-10, abs_=1e-14, max_iters=3000,
                                                         ilu_atol=1e-10))
    assert res <= 1e-11
    vel = mesh.dof_comp < dim
    exact = np.where(mesh.dof_comp == 0, shear(mesh.dof_coords), 0.0)
    assert np.max(np.abs(U[vel] - exact[vel])) <= 1e-10
