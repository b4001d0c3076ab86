"""GPU parity tests: the CUDA hot path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Tolerances: matrix entries 1e-12 relative to the row max-abs, RHS 1e-12
relative to its inf-norm (BASELINE.json north_star, SURVEY.md §8a note 5); GMRES iteration counts
within +-2; Newton solution 1e-10 relative L2."""
import os
import types

import numpy as np
import pytest

from tests import mms
from tests.util import hotpath_from_oracle_mesh, random_state, row_scaled_error

pytestmark = pytest.mark.gpu

TOL_ENTRY = 1e-12

CASES = [  # dim, n, pu, pp
    (2, 6, 1, 1), (2, 5, 2, 2), (2, 5, 2, 1), (3, 3, 1, 1), (3, 3, 2, 2), (3, 2, 2, 1)]


def _check_assembly(oracle, mesh, hp, U, scheme, dts, hist, force, nu, srf=False, omega=(0, 0, 0)):
    pr = oracle.scheme_params(scheme, dts, nu, srf, omega)
    a_ref, b_ref = oracle.assemble(mesh, U, pr, True, force, *hist)
    hp.set_vector("evaluation_point", U)
    for name, h in zip(("solution_m1", "solution_m2", "solution_m3"), hist):
        if h is not None:
            hp.set_vector(name, h)
    hp.assemble(True, scheme, dts)
    a_gpu, b_gpu = hp.get_matrix_values(), hp.get_vector("system_rhs")
    assert row_scaled_error(mesh, a_gpu, a_ref) <= TOL_ENTRY
    assert np.max(np.abs(b_gpu - b_ref)) <= TOL_ENTRY * np.max(np.abs(b_ref))
    assert abs(hp.rhs_norm() - np.linalg.norm(b_ref)) <= 1e-12 * np.linalg.norm(b_ref)
    # rhs-only assembly must give the same residual and leave the matrix untouched
    hp.assemble(False, scheme, dts)
    assert np.max(np.abs(hp.get_vector("system_rhs") - b_ref)) <= TOL_ENTRY * np.max(np.abs(b_ref))
    assert np.array_equal(hp.get_matrix_values(), a_gpu)
    return a_ref, b_ref


@pytest.mark.parametrize("dim,n,pu,pp", CASES)
def test_assembly_steady(oracle, dim, n, pu, pp):
    mesh = oracle.BoxMesh(dim, n, pu, pp)
    force = mesh.evaluate_force(mms.forcing_2d if dim == 2 else mms.forcing_3d)
    hp = hotpath_from_oracle_mesh(mesh, 0.37, force)
    _check_assembly(oracle, mesh, hp, random_state(mesh), "steady", None, (None,) * 3, force, 0.37)
    hp.close()


@pytest.mark.parametrize("scheme", ["bdf1", "bdf2", "bdf3", "sdirk2_1", "sdirk2_2", "sdirk3_1",
                                    "sdirk3_2", "sdirk3_3"])
@pytest.mark.parametrize("dim,n,pu,pp", [(2, 4, 2, 1), (3, 2, 2, 2)])
def test_assembly_transient(oracle, scheme, dim, n, pu, pp):
    mesh = oracle.BoxMesh(dim, n, pu, pp)
    hp = hotpath_from_oracle_mesh(mesh, 0.05, None)
    hist = tuple(random_state(mesh, seed=s) for s in (11, 12, 13))
    _check_assembly(oracle, mesh, hp, random_state(mesh), scheme, [0.1, 0.2, 0.3], hist, None, 0.05)
    hp.close()


@pytest.mark.parametrize("dim,n,pu,pp,omega", [(2, 4, 1, 1, (0, 0, 1.7)), (3, 2, 2, 2, (0.3, -0.4, 1.1)),
                                               (3, 3, 1, 1, (0, 0, -6.28318))])
def test_assembly_rotating_frame(oracle, dim, n, pu, pp, omega):
    mesh = oracle.BoxMesh(dim, n, pu, pp)
    hp = hotpath_from_oracle_mesh(mesh, 0.2, None, True, omega)
    _check_assembly(oracle, mesh, hp, random_state(mesh), "steady", None, (None,) * 3, None, 0.2,
                    True, omega)
    _check_assembly(oracle, mesh, hp, random_state(mesh, 5), "bdf2", [0.05, 0.05, 0.05],
                    (random_state(mesh, 6), random_state(mesh, 7), None), None, 0.2, True, omega)
    hp.close()


def test_assembly_lid_driven_cavity_bcs(oracle):
    """Inhomogeneous (lid) + homogeneous Dirichlet rows, 3D Q2-Q2 cavity (the bench workload)."""
    lid = lambda x: np.stack([np.ones(len(x)), 0 * x[:, 0], 0 * x[:, 0]], axis=1)
    bcs = {0: ("noslip",), 1: ("noslip",), 2: ("noslip",), 4: ("noslip",), 5: ("noslip",),
           3: ("function", lid)}
    mesh = oracle.BoxMesh(3, 3, 2, 2, bcs=bcs)
    hp = hotpath_from_oracle_mesh(mesh, 0.005, None)
    U = mesh.apply_nonzero_constraints(np.zeros(mesh.ndof))
    _check_assembly(oracle, mesh, hp, U, "steady", None, (None,) * 3, None, 0.005)
    _check_assembly(oracle, mesh, hp, random_state(mesh, 3, 0.3), "steady", None, (None,) * 3, None,
                    0.005)
    hp.close()


@pytest.mark.parametrize("dim,n,pu,pp", CASES)
def test_spmv_ilu_against_oracle(oracle, dim, n, pu, pp):
    mesh = oracle.BoxMesh(dim, n, pu, pp)
    hp = hotpath_from_oracle_mesh(mesh, 0.1, None)
    U = random_state(mesh, scale=0.5)
    a_ref, b_ref = oracle.assemble(mesh, U, oracle.scheme_params("steady", None, 0.1), True)
    hp.set_matrix_values(a_ref)
    x = np.random.default_rng(7).standard_normal(mesh.ndof)
    y_ref = oracle.spmv(mesh, a_ref, x)
    assert np.max(np.abs(hp.spmv(x) - y_ref)) <= 1e-13 * np.max(np.abs(y_ref))
    lu_ref, dp = oracle.ilu0(mesh, a_ref, 1e-8, 1.0)
    hp.setup_ilu(0, 1e-8, 1.0)
    lu = hp.get_ilu_values()
    assert row_scaled_error(mesh, lu, lu_ref) <= 1e-11
    z_ref = oracle.ilu_apply(mesh, lu_ref, dp, x)
    z = hp.ilu_apply(x)
    assert np.max(np.abs(z - z_ref)) <= 1e-10 * np.max(np.abs(z_ref))
    lo, up = hp.ilu_levels()
    assert 1 <= lo <= mesh.ndof and 1 <= up <= mesh.ndof
    hp.close()


def test_ilu_diagonal_perturbation_and_zero_pivot(oracle):
    from softx_2020_200_b200 import GlsnsError
    mesh = oracle.BoxMesh(2, 3, 1, 1)
    hp = hotpath_from_oracle_mesh(mesh)
    a_ref, _ = oracle.assemble(mesh, random_state(mesh), oracle.scheme_params("steady", None, 1.0), True)
    hp.set_matrix_values(a_ref)
    for atol, rtol in ((1e-3, 1.0), (0.0, 1.5), (1e-12, 1.0)):
        hp.setup_ilu(0, atol, rtol)
        lu_ref, _ = oracle.ilu0(mesh, a_ref, atol, rtol)
        assert row_scaled_error(mesh, hp.get_ilu_values(), lu_ref) <= 1e-11
    hp.set_matrix_values(np.zeros_like(a_ref))
    with pytest.raises(GlsnsError) as e:
        hp.setup_ilu(0, 0.0, 1.0)
    assert e.value.status == 4  # GLSNS_ERR_ZERO_PIVOT
    with pytest.raises(GlsnsError) as e:
        hp.setup_ilu(-1, 1e-8, 1.0)
    assert e.value.status == 1  # GLSNS_ERR_BAD_ARGUMENT
    hp.close()


@pytest.mark.parametrize("dim,n,pu,pp,fills", [(2, 8, 1, 1, (1, 4, 2, 0)), (2, 5, 2, 2, (1, 2)),
                                               (3, 3, 1, 1, (1, 2)), (3, 2, 2, 2, (1,))])
def test_ilu_fill_levels_against_oracle(oracle, dim, n, pu, pp, fills):
    """`ilu preconditioner fill = k` (setup_ILU, gls_navier_stokes.cc:1161-1176; the shipped cavity
    and cylinder examples use 1, mms2d_gls.prm 4): the level-of-fill pattern, the factors on it and
    their application against the oracle; the matrix the host sees keeps its own pattern; levels
    can be switched back and forth."""
    mesh = oracle.BoxMesh(dim, n, pu, pp)
    hp = hotpath_from_oracle_mesh(mesh, 0.1, None)
    a_ref, b_ref = oracle.assemble(mesh, random_state(mesh, scale=0.5),
                                   oracle.scheme_params("steady", None, 0.1), True)
    hp.set_matrix_values(a_ref)
    x = np.random.default_rng(7).standard_normal(mesh.ndof)
    y_ref = oracle.spmv(mesh, a_ref, x)
    for fill in fills:
        hp.setup_ilu(fill, 1e-8, 1.0)
        pm, a2p = oracle.iluk_pattern(mesh, fill)
        rp, col = hp.get_ilu_pattern()
        assert np.array_equal(rp, pm.rowptr) and np.array_equal(col, pm.col)
        lu_ref, dp = oracle.ilu0(pm, oracle.pad_values(pm, a2p, a_ref), 1e-8, 1.0)
        assert row_scaled_error(pm, hp.get_ilu_values(), lu_ref) <= 1e-10
        z_ref = oracle.ilu_apply(pm, lu_ref, dp, x)
        assert np.max(np.abs(hp.ilu_apply(x) - z_ref)) <= 1e-9 * np.max(np.abs(z_ref))
        # the matrix is unchanged, on the host's pattern and in the product
        assert np.array_equal(hp.get_matrix_values(), a_ref)
        assert np.max(np.abs(hp.spmv(x) - y_ref)) <= 1e-13 * np.max(np.abs(y_ref))
    hp.close()


def test_gmres_with_ilu_fill_and_reassembly(oracle):
    """GMRES + ILU(k) through solve_linear_system, including a re-assembly on the padded pattern
    (second Newton iteration): iteration counts against the oracle, and fewer than with ILU(0)."""
    mesh = oracle.BoxMesh(2, 16, 1, 1)
    force = mesh.evaluate_force(mms.forcing_2d)
    pr = oracle.scheme_params("steady", None, 1.0)
    hp = hotpath_from_oracle_mesh(mesh, 1.0, force)
    U = np.zeros(mesh.ndof)
    its = {}
    for fill in (0, 1, 4):
        hp.set_vector("evaluation_point", U)
        hp.assemble(True)
        val, rhs = oracle.assemble(mesh, U, pr, True, force)
        x_ref, it_ref, _ = oracle.solve_linear_system(mesh, val, rhs, rel=1e-8, abs_=1e-12,
                                                      ilu_fill=fill)
        x, info = hp.solve_linear_system(relative_residual=1e-8, minimum_residual=1e-12,
                                         ilu_fill=fill)
        assert abs(info["iterations"] - it_ref) <= 1
        assert np.linalg.norm(x - x_ref) <= 1e-6 * np.linalg.norm(x_ref)
        assert row_scaled_error(mesh, hp.get_matrix_values(), val) <= TOL_ENTRY
        its[fill] = info["iterations"]
        U = mesh.apply_nonzero_constraints(U + x)        # next linearisation point
    assert its[4] < its[1] <= its[0] + 2
    hp.close()


def test_mms2d_gls_prm_with_its_shipped_ilu_fill(oracle):
    """applications_tests/gls_navier_stokes_2d/mms2d_gls.prm runs GMRES with `ilu preconditioner
    fill = 4` (:84); its output table (mms2d_gls.output:24-26) through the mirrored C++ interface
    with that setting: 256 cells -> velocity L2 error 3.4363e-02."""
    from softx_2020_200_b200.mesh import BoxMesh
    from tests.mirror.solver import GLSNavierStokesSolver
    from tests.test_host_mirror import _match_numbering
    mesh = BoxMesh(2, 16, 1, 1, with_q_points=True)
    force = mms.forcing_2d(mesh.array("q_points").reshape(-1, 2)).reshape(mesh.n_cells, mesh.n_q, 2)
    prm = ("subsection linear solver\n set method = gmres\n set ilu preconditioner fill = 4\n"
           " set ilu preconditioner absolute tolerance = 1e-12\n set relative residual = 1e-4\n"
           " set minimum residual = 1e-9\n set verbosity = verbose\nend\n")
    s = GLSNavierStokesSolver(mesh, prm, force)
    s.set_vector("present_solution", np.zeros(mesh.n_dofs))
    s.solve_non_linear_system("steady", False, True)
    nat = BoxMesh(2, 16, 1, 1, renumber=False)
    om = oracle.BoxMesh(2, 16, 1, 1, renumber=_match_numbering(nat, mesh, 2))
    err_u, _ = oracle.l2_error(om, s.present_solution, mms.exact_2d)
    assert float("%.5g" % err_u) == 3.4363e-02
    s.close()


def test_geometry_per_q_without_mapping_laplacian_is_refused(oracle):
    """glsns.h: per-point geometry needs the mapping's second derivatives; without them the
    Laplacian terms would be silently wrong, so glsns_set_mesh refuses (GLSNS_ERR_UNSUPPORTED)."""
    from softx_2020_200_b200 import GLSHotPath, GlsnsError
    mesh = mms.couette_mesh(oracle, 2)
    fe = mesh.fe
    hp = GLSHotPath(0)
    hp.set_fe(mesh.dim, mesh.pu, fe.Nu, fe.dNu, fe.d2Nu, fe.Np, fe.dNp, fe.wq)
    ptr, order, _ = mesh.color_lists()
    with pytest.raises(GlsnsError) as e:
        hp.set_mesh(mesh.ndof, mesh.cell_dofs, mesh.cell_invJ, mesh.cell_detJ, mesh.cell_measure,
                    mesh.constrained, mesh.rowptr, mesh.col, ptr, order, q_points=mesh.qpoints,
                    constraint_values=mesh.constraint_value, geometry_per_q=True)
    assert e.value.status == 6
    hp.close()


def _newton_gpu(hp, mesh, U0, scheme="steady", dts=None, tol=1e-6, max_it=10, lin=None, log=None):
    """NewtonNonLinearSolver::solve (newton_non_linear_solver.h:76-139) over the C ABI with the
    device-resident line search."""
    lin = lin or {}
    hp.set_vector("present_solution", U0)
    hp.set_vector("evaluation_point", U0)
    current_res = last_res = 1.0
    it = 0
    while current_res > tol and it < max_it:
        hp.assemble(True, scheme, dts)
        if it == 0:
            current_res = last_res = hp.rhs_norm()
        _, info = hp.solve_linear_system(download=False, **lin)
        if log is not None:
            log.append((info["iterations"], info["true_residual"]))
        alpha = 1.0
        while alpha > 1e-3:
            hp.line_search_point(alpha)
            hp.assemble(False, scheme, dts)
            current_res = hp.rhs_norm()
            if current_res < 0.9 * last_res or last_res < tol:
                break
            alpha *= 0.5
        hp.accept_evaluation_point()
        last_res = current_res
        it += 1
    return hp.get_vector("present_solution"), it, current_res


def test_restart_01_golden_on_gpu(oracle):
    """The reference's tests/solvers/restart_01.output reproduced by the CUDA path: GMRES iteration
    counts 8/6/10 and true residuals per Newton step, velocity L2 error 0.0343628."""
    mesh = oracle.BoxMesh(2, 16, 1, 1)
    force = mesh.evaluate_force(mms.forcing_2d)
    hp = hotpath_from_oracle_mesh(mesh, 1.0, force)
    log = []
    U, it, res = _newton_gpu(hp, mesh, np.zeros(mesh.ndof), log=log)
    assert [k for k, _ in log] == [8, 6, 10]
    for (_, r), g in zip(log, [0.00204885, 9.85227e-05, 3.32384e-08]):
        assert abs(r - g) <= 2e-6 * g
    err_u, _ = oracle.l2_error(mesh, U, mms.exact_2d)
    assert float("%.6g" % err_u) == 0.0343628
    hp.close()


HANGING = [  # dim, n, pu, pp, refined region
    (2, 4, 1, 1, lambda c: (c[:, 0] > 0) & (c[:, 1] < 0.5)),
    (2, 4, 2, 2, lambda c: (c[:, 0] > 0) & (c[:, 1] < 0.5)),
    (2, 5, 2, 1, lambda c: np.abs(c[:, 0]) + np.abs(c[:, 1]) < 0.7),
    (3, 3, 1, 1, lambda c: (c[:, 0] > 0) & (c[:, 1] < 0.3) & (c[:, 2] > -0.4)),
    (3, 2, 2, 2, lambda c: (c[:, 0] > 0) & (c[:, 1] < 0.3) & (c[:, 2] > -0.4))]


@pytest.mark.parametrize("dim,n,pu,pp,refine", HANGING)
def test_hanging_node_constraints_against_oracle(oracle, dim, n, pu, pp, refine):
    """A box with a once-refined region (what Kelly refinement produces, examples/03-cylinder and
    04-ribbon_mixer; navier_stokes_base.cc:610-729): the hanging-node lines resolved in the scatter
    as AffineConstraints::distribute_local_to_global does (gls_navier_stokes.cc:755-771) -- matrix
    and right-hand side entry by entry against the oracle, steady and BDF2 -- and distributed after
    the solve and in the line search: a Newton solve of a flow that lies in the finite element
    space (Couette for Q1, Poiseuille for Q2) is reproduced to rounding on the non-conforming mesh."""
    nu = 0.7
    if pu == 1:
        shear = (lambda x: x[:, 1]) if dim == 2 else (lambda x: x[:, 1] + 0.5 * x[:, 2])
    else:
        shear = lambda x: 1 - x[:, 1] ** 2
    bc = lambda x: np.stack([shear(x)] + [0 * x[:, 0]] * (dim - 1), axis=1)
    mesh = oracle.RefinedBoxMesh(dim, n, pu, pp, refine, bcs={None: ("function", bc)})
    assert (mesh.constrained == 2).sum() > 0
    hp = hotpath_from_oracle_mesh(mesh, nu, None)
    _check_assembly(oracle, mesh, hp, random_state(mesh, scale=0.4), "steady", None, (None,) * 3, None, nu)
    _check_assembly(oracle, mesh, hp, random_state(mesh, 5, 0.4), "bdf2", [0.05, 0.05, 0.05],
                    (random_state(mesh, 6, 0.4), random_state(mesh, 7, 0.4), None), None, nu)
    U0 = mesh.apply_nonzero_constraints(np.zeros(mesh.ndof))
    lin = dict(rel=1e-10, abs_=1e-14, max_iters=3000, ilu_atol=1e-10)
    log_ref, log = [], []
    U_ref, it_ref, _ = oracle.newton_solve(mesh, U0, oracle.scheme_params("steady", None, nu), None,
                                           tol=1e-11, max_it=15, lin=lin, log=log_ref)
    U, it, res = _newton_gpu(hp, mesh, U0, tol=1e-11, max_it=15, log=log,
                             lin=dict(relative_residual=1e-10, minimum_residual=1e-14,
                                      max_iterations=3000, ilu_atol=1e-10))
    assert res <= 1e-11 and it == it_ref
    for (k, _), (k_ref, _) in zip(log, log_ref):
        assert abs(k - k_ref) <= 2
    vel = mesh.dof_comp < dim
    exact = np.where(mesh.dof_comp == 0, shear(mesh.dof_coords), 0.0)
    assert np.max(np.abs(U[vel] - exact[vel])) <= 1e-10          # the flow is in the FE space
    assert np.max(np.abs(U - U_ref)) <= 1e-9 * np.max(np.abs(U_ref))
    hp.close()


def test_taylor_couette_curved_q2_cells_on_gpu(oracle):
    """examples/02-taylor-couette's geometry (MappingQ(2) on every cell of a hyper_shell,
    taylorcouette_gls.prm with `qmapping all = true`): Jacobian and residual entries of the CUDA
    assembly on curved cells against the oracle (steady, BDF2 and rotating-frame terms), then the
    Newton solution through the CUDA path reproduces taylorcouette_gls.output:45 -- velocity L2
    error 7.7383e-04 at 64 cells, the reference's pin of the Q2 Laplacian terms on curved cells."""
    mesh = mms.couette_mesh(oracle, 2)
    assert mesh.geometry_per_q
    hp = hotpath_from_oracle_mesh(mesh, 1.0, None)
    _check_assembly(oracle, mesh, hp, random_state(mesh, scale=0.3), "steady", None, (None,) * 3,
                    None, 1.0)
    _check_assembly(oracle, mesh, hp, random_state(mesh, 5, 0.3), "bdf2", [0.05, 0.05, 0.05],
                    (random_state(mesh, 6, 0.3), random_state(mesh, 7, 0.3), None), None, 1.0)
    hp.close()
    hp = hotpath_from_oracle_mesh(mesh, 0.2, None, True, (0, 0, 1.7))
    _check_assembly(oracle, mesh, hp, random_state(mesh, 8, 0.3), "steady", None, (None,) * 3, None,
                    0.2, True, (0, 0, 1.7))
    hp.close()
    hp = hotpath_from_oracle_mesh(mesh, 1.0, None)
    U0 = mesh.apply_nonzero_constraints(np.zeros(mesh.ndof))
    U, it, res = _newton_gpu(hp, mesh, U0, tol=1e-10,
                             lin=dict(relative_residual=1e-8, minimum_residual=1e-13,
                                      max_iterations=4000, ilu_atol=1e-10))
    assert res <= 1e-10
    err_u, _ = oracle.l2_error(mesh, U, mms.couette_exact)
    assert "%.4e" % err_u == "7.7383e-04", err_u
    hp.close()


@pytest.mark.parametrize("dim,n,pu,pp,nu", [(3, 4, 1, 1, 1.0), (3, 3, 2, 2, 1.0), (2, 8, 2, 2, 0.1)])
def test_newton_solution_matches_oracle(oracle, dim, n, pu, pp, nu):
    mesh = oracle.BoxMesh(dim, n, pu, pp)
    force = mesh.evaluate_force(mms.forcing_2d if dim == 2 else mms.forcing_3d)
    lin = dict(rel=1e-4, abs_=1e-9, max_iters=5000, ilu_atol=1e-10)
    log_ref, log = [], []
    U_ref, it_ref, _ = oracle.newton_solve(mesh, np.zeros(mesh.ndof),
                                           oracle.scheme_params("steady", None, nu), force,
                                           tol=1e-8, lin=lin, log=log_ref)
    hp = hotpath_from_oracle_mesh(mesh, nu, force)
    U, it, res = _newton_gpu(hp, mesh, np.zeros(mesh.ndof), tol=1e-8, log=log,
                             lin=dict(relative_residual=1e-4, minimum_residual=1e-9,
                                      max_iterations=5000, ilu_atol=1e-10))
    assert it == it_ref
    for (k, _), (k_ref, _) in zip(log, log_ref):
        assert abs(k - k_ref) <= 2
    # the two runs stop GMRES at slightly different iterates; both are Newton-converged to 1e-8
    assert np.linalg.norm(U - U_ref) <= 1e-7 * np.linalg.norm(U_ref)
    hp.close()


@pytest.mark.parametrize("dim,n,pu,pp,nu", [(2, 16, 1, 1, 1.0), (3, 4, 2, 2, 1.0), (2, 8, 2, 2, 0.1)])
def test_bicgstab_matches_oracle(oracle, dim, n, pu, pp, nu):
    """`method = bicgstab` (solve_system_BiCGStab, gls_navier_stokes.cc:1291-1340): same matrix, same
    ILU, iteration count within +-1 of the oracle's restatement, solution to the solver tolerance,
    logged residual = the true one."""
    mesh = oracle.BoxMesh(dim, n, pu, pp)
    force = mesh.evaluate_force(mms.forcing_2d if dim == 2 else mms.forcing_3d)
    U = 0.1 * random_state(mesh)
    pr = oracle.scheme_params("steady", None, nu)
    val, rhs = oracle.assemble(mesh, U, pr, True, force)
    tol = max(1e-8 * np.linalg.norm(rhs), 1e-12)
    lu, dp = oracle.ilu0(mesh, val, 1e-10, 1.0)
    x_ref, it_ref, tr_ref, ok = oracle.bicgstab(mesh, val, lu, dp, rhs, tol, 2000)
    assert ok
    hp = hotpath_from_oracle_mesh(mesh, nu, force)
    hp.set_vector("evaluation_point", U)
    hp.assemble(True)
    x, info = hp.solve_linear_system(relative_residual=1e-8, minimum_residual=1e-12,
                                     max_iterations=2000, ilu_atol=1e-10, method="bicgstab")
    assert abs(info["iterations"] - it_ref) <= 1
    assert info["true_residual"] < 10 * tol
    x_ref[mesh.constrained != 0] = 0.0
    assert np.linalg.norm(x - x_ref) <= 1e-6 * np.linalg.norm(x_ref)
    # and against GMRES on the same system
    xg, _ = hp.solve_linear_system(relative_residual=1e-10, minimum_residual=1e-13,
                                   max_iterations=2000, ilu_atol=1e-10)
    assert np.linalg.norm(x - xg) <= 1e-6 * np.linalg.norm(xg)
    from softx_2020_200_b200 import GlsnsError
    with pytest.raises(GlsnsError, match="This solver is not allowed"):
        hp.solve_linear_system(method="amg")
    hp.close()


def test_bicgstab_through_the_cpp_mirror(oracle):
    """`set method = bicgstab` in the .prm reaches solve_system_BiCGStab and Newton converges to the
    same discrete solution as with GMRES (restart_01's problem: velocity L2 error 0.0343628)."""
    import re
    from softx_2020_200_b200.mesh import BoxMesh
    from tests.mirror.solver import GLSNavierStokesSolver
    from tests.test_host_mirror import _match_numbering
    mesh = BoxMesh(2, 16, 1, 1, with_q_points=True)
    force = mms.forcing_2d(mesh.array("q_points").reshape(-1, 2)).reshape(mesh.n_cells, mesh.n_q, 2)
    s = GLSNavierStokesSolver(mesh, "subsection linear solver\n set method = bicgstab\nend\n", force)
    s.set_vector("present_solution", np.zeros(mesh.n_dofs))
    s.solve_non_linear_system("steady", False, True)
    assert len(re.findall(r"-Iterative solver took : (\d+) steps", s.log)) == 3
    nat = BoxMesh(2, 16, 1, 1, renumber=False)
    om = oracle.BoxMesh(2, 16, 1, 1, renumber=_match_numbering(nat, mesh, 2))
    err_u, _ = oracle.l2_error(om, s.present_solution, mms.exact_2d)
    assert abs(err_u - 0.0343628) < 2e-6
    s.close()


@pytest.mark.parametrize("dim,n,pu,pp", [(2, 6, 1, 1), (2, 4, 2, 2), (2, 4, 2, 1), (3, 3, 2, 2), (3, 3, 1, 1)])
def test_l2_projection_and_cfl_against_oracle(oracle, dim, n, pu, pp):
    """set_initial_condition(L2projection) (gls_navier_stokes.cc:795-803, assemble_L2_projection
    :829-914) and calculate_CFL (postprocessing_cfl.cc:34-87) against their restatements: mass
    matrix and right-hand side entry by entry, the projected field, the CFL number."""
    def init(x):
        out = np.zeros((len(x), dim + 1))
        for c in range(dim + 1):
            out[:, c] = np.sin(0.7 * (c + 1) * x[:, 0]) * np.cos(0.4 * x[:, 1] + 0.2 * c) + 0.1 * c
        if dim == 3:
            out *= (1.0 + 0.3 * x[:, 2:3])
        return out
    lid = lambda x: init(x)[:, :dim]
    bcs = {b: ("function", lid) for b in range(2 * dim - 1)}    # one face stays free
    mesh = oracle.BoxMesh(dim, n, pu, pp, bcs=bcs)
    hp = hotpath_from_oracle_mesh(mesh, 1.0, None)
    v_ref, b_ref = oracle.assemble_l2_projection(mesh, init)
    fe = mesh.fe
    hp.assemble_l2_projection(init(mesh.qpoints.reshape(-1, dim)))
    assert row_scaled_error(mesh, hp.get_matrix_values(), v_ref) <= TOL_ENTRY
    assert np.max(np.abs(hp.get_vector("system_rhs") - b_ref)) <= TOL_ENTRY * np.max(np.abs(b_ref))
    U_ref, it_ref, ok = oracle.l2_projection(mesh, init, rel=1e-12, abs_=1e-14)
    assert ok
    _, info = hp.solve_linear_system(relative_residual=1e-12, minimum_residual=1e-14)
    assert abs(info["iterations"] - it_ref) <= 1
    hp.distribute_constraints("newton_update")
    U = hp.get_vector("newton_update")
    assert np.max(np.abs(U - U_ref)) <= 1e-9 * np.max(np.abs(U_ref))
    con = mesh.constrained != 0
    assert np.array_equal(U[con], mesh.constraint_value[con])
    hp.set_vector("present_solution", U)
    cfl = hp.calculate_cfl("present_solution", oracle.shape_at_centre(dim, pu), max(pu, pp), 0.02)
    assert abs(cfl - oracle.calculate_cfl(mesh, U, 0.02)) <= 1e-13 * cfl
    hp.close()


def test_l2_projection_initial_condition_through_the_cpp_mirror(oracle):
    """set_initial_condition(L2projection) + calculate_CFL of the mirrored GLSNavierStokesSolver with
    the reference's own tolerances (solve_system_GMRES(true, 1e-15, 1e-15, true), :800): the
    Taylor-Green field of applications_tests/.../taylor-green-vortex_gls_*.prm's initial condition
    (u = cos x sin y, v = -sin x cos y) on a periodic-free box, against the oracle."""
    from softx_2020_200_b200.mesh import BoxMesh
    from tests.mirror.solver import GLSNavierStokesSolver
    from tests.test_host_mirror import _match_numbering

    def tg(x):
        return np.stack([np.cos(x[:, 0]) * np.sin(x[:, 1]), -np.sin(x[:, 0]) * np.cos(x[:, 1]),
                         -0.25 * (np.cos(2 * x[:, 0]) + np.cos(2 * x[:, 1]))], axis=1)
    mesh = BoxMesh(2, 8, 2, 1, with_q_points=True)
    s = GLSNavierStokesSolver(mesh, "subsection FEM\n set velocity order = 2\n set pressure order = 1\nend\n"
                                    "subsection linear solver\n set max iters = 200\nend\n", None)
    init = tg(mesh.array("q_points").reshape(-1, 2))
    try:
        s.set_initial_condition_l2_projection(init)
    except Exception as e:            # 1e-15 can be below what fp64 GMRES reaches: the reference would
        pytest.skip("tolerance 1e-15 not reached: %s" % e)   # throw NoConvergence there too
    nat = BoxMesh(2, 8, 2, 1, renumber=False)
    om = oracle.BoxMesh(2, 8, 2, 1, renumber=_match_numbering(nat, mesh, 2))
    U_ref, _, _ = oracle.l2_projection(om, tg, rel=1e-13, abs_=1e-14)
    U = s.present_solution
    assert np.max(np.abs(U - U_ref)) <= 1e-9
    cfl = s.calculate_cfl(oracle.shape_at_centre(2, 2), 0.05)
    assert abs(cfl - oracle.calculate_cfl(om, U_ref, 0.05)) <= 1e-9
    s.close()


CAVITY_PRM = """
# examples/01-cavity/cavity.prm as shipped (physical properties :15-17, FEM :63-66,
# non-linear solver :78-83, linear solver :88-97); BASELINE.json configs[0]
subsection physical properties
    set kinematic viscosity            = %g
end
subsection FEM
    set velocity order            = 1
    set pressure order            = 1
end
subsection non-linear solver
  set tolerance               = 1e-8
  set max iterations          = 10
  set residual precision      = 2
  set verbosity               = verbose
end
subsection linear solver
  set method                                 = gmres
  set max iters                              = 5000
  set relative residual                      = 1e-9
  set minimum residual                       = 1e-9
  set ilu preconditioner fill                = 1
  set ilu preconditioner absolute tolerance  = 1e-12
  set ilu preconditioner relative tolerance  = 1.00
  set verbosity               = verbose
end
"""


@pytest.mark.parametrize("nu,n", [(1.0, 64), (0.005, 32)])
def test_example_01_cavity_prm_as_shipped(oracle, nu, n):
    """BASELINE.json configs[0], the reference's CPU-runnable case: 2D lid-driven cavity
    (hyper_cube -1:1 colorized, bc 0-2 noslip, bc 3 u = 1; examples/01-cavity/cavity.prm:20-60),
    Q1-Q1, steady Newton + GMRES/ILU(1) with the file's own solver settings, at its shipped
    viscosity (initial refinement 6 = 64^2 cells) and at Re = 400 (nu = 0.005), through the mirrored
    GLSNavierStokesSolver against the oracle's Newton solve: same Newton iteration count, GMRES
    counts within +-2, same discrete solution."""
    import re
    from softx_2020_200_b200.mesh import BoxMesh
    from tests.mirror.solver import GLSNavierStokesSolver
    from tests.test_host_mirror import _match_numbering
    bcs = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (3, "function", (1.0, 0.0))]
    mesh = BoxMesh(2, n, 1, 1, bcs=bcs)
    s = GLSNavierStokesSolver(mesh, CAVITY_PRM % nu, None)
    s.set_vector("present_solution", mesh.initial_state())
    s.solve_non_linear_system("steady", False, True)
    its = [int(k) for k in re.findall(r"-Iterative solver took : (\d+) steps", s.log)]
    lid = lambda x: np.stack([np.ones(len(x)), 0 * x[:, 0]], axis=1)
    obcs = {0: ("noslip",), 1: ("noslip",), 2: ("noslip",), 3: ("function", lid)}
    nat = BoxMesh(2, n, 1, 1, bcs=bcs, renumber=False)
    om = oracle.BoxMesh(2, n, 1, 1, bcs=obcs, renumber=_match_numbering(nat, mesh, 2))
    log = []
    U_ref, it_ref, res_ref = oracle.newton_solve(
        om, om.apply_nonzero_constraints(np.zeros(om.ndof)), oracle.scheme_params("steady", None, nu),
        None, tol=1e-8, max_it=10, log=log,
        lin=dict(rel=1e-9, abs_=1e-9, max_iters=5000, ilu_atol=1e-12, ilu_fill=1))
    assert len(its) == it_ref and res_ref < 1e-8
    for k, (k_ref, _) in zip(its, log):
        assert abs(k - k_ref) <= 2
    U = s.present_solution
    assert np.linalg.norm(U - U_ref) <= 1e-7 * np.linalg.norm(U_ref)
    s.close()


def test_set_initial_condition_viscous_and_nodal_through_the_cpp_mirror(oracle):
    """set_initial_condition (gls_navier_stokes.cc:784-828) with `initial conditions: type = viscous`
    (a steady solve at the subsection's artificial viscosity, :811-822, the physical viscosity
    restored afterwards) and `nodal` (set_nodal_values, navier_stokes_base.cc:926-944) on the 2D
    cavity, against the oracle's Newton solves."""
    from softx_2020_200_b200.mesh import BoxMesh
    from tests.mirror.solver import GLSNavierStokesSolver
    from tests.test_host_mirror import _match_numbering
    n, nu, nu_ic = 16, 0.02, 1.0
    bcs = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (3, "function", (1.0, 0.0))]
    mesh = BoxMesh(2, n, 1, 1, bcs=bcs)
    prm = CAVITY_PRM % nu + "subsection initial conditions\n set type = viscous\n set viscosity = %g\nend\n" % nu_ic
    s = GLSNavierStokesSolver(mesh, prm, None)
    s.set_initial_condition(initial_nodal=np.zeros(mesh.n_dofs))        # type from the .prm
    lid = lambda x: np.stack([np.ones(len(x)), 0 * x[:, 0]], axis=1)
    obcs = {0: ("noslip",), 1: ("noslip",), 2: ("noslip",), 3: ("function", lid)}
    nat = BoxMesh(2, n, 1, 1, bcs=bcs, renumber=False)
    om = oracle.BoxMesh(2, n, 1, 1, bcs=obcs, renumber=_match_numbering(nat, mesh, 2))
    lin = dict(rel=1e-9, abs_=1e-9, max_iters=5000, ilu_atol=1e-12, ilu_fill=1)
    U0 = om.apply_nonzero_constraints(np.zeros(om.ndof))
    U_ic, _, _ = oracle.newton_solve(om, U0, oracle.scheme_params("steady", None, nu_ic), None,
                                     tol=1e-8, max_it=10, lin=lin)
    assert np.linalg.norm(s.present_solution - U_ic) <= 1e-7 * np.linalg.norm(U_ic)
    # the physical viscosity is back: the steady solve from that start matches the oracle's at nu
    s.solve_non_linear_system("steady", False, True)
    U_ref, _, res = oracle.newton_solve(om, U_ic, oracle.scheme_params("steady", None, nu), None,
                                        tol=1e-8, max_it=10, lin=lin)
    assert res < 1e-8
    assert np.linalg.norm(s.present_solution - U_ref) <= 1e-6 * np.linalg.norm(U_ref)
    # nodal: the interpolated values with the non-zero constraints distributed
    vals = np.random.default_rng(3).uniform(-1, 1, mesh.n_dofs)
    s.set_initial_condition(initial_nodal=vals, type="nodal")
    con = mesh.array("constrained") != 0
    expect = np.where(con, mesh.array("constraint_values"), vals)
    assert np.array_equal(s.present_solution, expect)
    with pytest.raises(RuntimeError, match="Initial condition could not be set"):
        s.set_initial_condition(initial_nodal=vals, type="none")
    s.close()


def test_taylor_green_vortex_bdf1_golden_on_gpu(oracle):
    """The reference's own transient golden through the CUDA path:
    applications_tests/gls_navier_stokes_2d/taylor-green-vortex_gls_bdf1.mpirun=2.output (1024
    periodic Q1-Q1 cells, 3267 dofs, nu = 1, BDF1 steps of 0.01).  Initial condition by
    glsns_assemble_l2_projection + GMRES, then 30 Newton-converged BDF1 steps; the CFL number of
    every step comes from glsns_calculate_cfl, enstrophy / kinetic energy / velocity L2 error from
    the host post-processors on the downloaded solution: all as printed by the reference (6
    significant digits; 2e-5 relative where the solver tolerances flip the last one).  Periodic
    faces: the host identifies the dofs in the cell -> dof table (oracle BoxMesh(periodic=...))."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                           "reference_golden.json")) as f:
        g = json.load(f)["taylor_green_vortex_bdf1"]

    def tg(t):
        def f(x):
            e = np.exp(-2.0 * t)
            return np.stack([e * np.cos(x[:, 0]) * np.sin(x[:, 1]), -e * np.sin(x[:, 0]) * np.cos(x[:, 1]),
                             -0.25 * (np.cos(2 * x[:, 0]) + np.cos(2 * x[:, 1]))], axis=1)
        return f
    mesh = oracle.BoxMesh(2, 32, 1, 1, lo=0.0, hi=6.28318530718, bcs={}, periodic=(0, 1))
    assert (mesh.ncell, mesh.ndof) == (g["cells"], g["dofs"])
    hp = hotpath_from_oracle_mesh(mesh, 1.0, None)
    hp.assemble_l2_projection(tg(0.0)(mesh.qpoints.reshape(-1, 2)))
    hp.solve_linear_system(relative_residual=1e-14, minimum_residual=1e-14, ilu_atol=1e-5,
                           download=False)
    hp.distribute_constraints("newton_update")
    U = hp.get_vector("newton_update")
    assert "%.6g" % oracle.enstrophy(mesh, U) == g["enstrophy_0"]
    assert "%.6g" % oracle.kinetic_energy(mesh, U) == g["kinetic_energy_0"]
    lin = dict(relative_residual=1e-4, minimum_residual=1e-9, max_iterations=5000, ilu_atol=1e-5)
    centre = oracle.shape_at_centre(2, 1)
    exact = 0
    for k in range(30):
        hp.set_vector("solution_m1", U)
        hp.set_vector("present_solution", U)
        vals = {"cfl": hp.calculate_cfl("present_solution", centre, 1, 0.01)}
        U, _, res = _newton_gpu(hp, mesh, U, "bdf1", [0.01] * 4, tol=1e-6, max_it=5, lin=lin)
        assert res < 1e-6
        vals.update(enstrophy=oracle.enstrophy(mesh, U), kinetic_energy=oracle.kinetic_energy(mesh, U),
                    l2_error_velocity=oracle.l2_error(mesh, U, tg(0.01 * (k + 1)))[0])
        for name, v in vals.items():
            ref = float(g[name][k])
            assert abs(v - ref) <= 2e-5 * abs(ref), (name, k, v, ref)
            exact += "%.6g" % v == g[name][k]
    assert exact >= 0.95 * 4 * 30, exact
    hp.close()


def test_taylor_green_vortex_sdirk3_through_the_cpp_mirror(oracle):
    """taylor-green-vortex_gls_sdirk3.prm end to end through the product's own host side: periodic
    C++ BoxMesh, GLSNavierStokesSolver with the file's solver subsections, set_initial_condition
    (L2projection), one sdirk3 time step by the mirrored time-stepping glue; CFL, enstrophy, kinetic
    energy and velocity L2 error against taylor-green-vortex_gls_sdirk3.mpirun=2.output."""
    import json
    from softx_2020_200_b200.mesh import BoxMesh
    from tests.mirror.solver import GLSNavierStokesSolver
    from tests.test_host_mirror import _match_numbering
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                           "reference_golden.json")) as f:
        g = json.load(f)["taylor_green_vortex_sdirk3"]

    def tg(t):
        def f(x):
            e = np.exp(-2.0 * t)
            return np.stack([e * np.cos(x[:, 0]) * np.sin(x[:, 1]), -e * np.sin(x[:, 0]) * np.cos(x[:, 1]),
                             -0.25 * (np.cos(2 * x[:, 0]) + np.cos(2 * x[:, 1]))], axis=1)
        return f
    L = 6.28318530718
    mesh = BoxMesh(2, 64, 2, 1, lo=0.0, hi=L, bcs=[], with_q_points=True, periodic=(0, 1))
    assert (mesh.n_cells, mesh.n_dofs) == (g["cells"], g["dofs"])
    prm = ("subsection FEM\n set velocity order = 2\n set pressure order = 1\nend\n"
           "subsection physical properties\n set kinematic viscosity = 1.000\nend\n"
           "subsection initial conditions\n set type = L2projection\nend\n"
           "subsection non-linear solver\n set verbosity = quiet\n set tolerance = 1e-6\n"
           " set max iterations = 5\nend\n"
           "subsection linear solver\n set verbosity = quiet\n set method = gmres\n"
           " set max iters = 5000\n set relative residual = 1e-4\n set minimum residual = 1e-9\n"
           " set ilu preconditioner fill = 1\n set ilu preconditioner absolute tolerance = 1e-5\n"
           " set ilu preconditioner relative tolerance = 1.00\nend\n")
    s = GLSNavierStokesSolver(mesh, prm, None)
    s.set_initial_condition(initial_at_q=tg(0.0)(mesh.array("q_points").reshape(-1, 2)))
    s.finish_time_step("sdirk3")
    nat = BoxMesh(2, 64, 2, 1, lo=0.0, hi=L, bcs=[], renumber=False)
    full = BoxMesh(2, 64, 2, 1, lo=0.0, hi=L, bcs=[], renumber=True)
    om = oracle.BoxMesh(2, 64, 2, 1, lo=0.0, hi=L, bcs={}, renumber=_match_numbering(nat, full, 2),
                        periodic=(0, 1))
    assert np.array_equal(om.cell_dofs.ravel(), mesh.array("cell_dofs"))
    U0 = s.present_solution
    assert "%.6g" % oracle.enstrophy(om, U0) == g["enstrophy_0"]
    assert "%.6g" % oracle.kinetic_energy(om, U0) == g["kinetic_energy_0"]
    assert "%.6g" % s.calculate_cfl(oracle.shape_at_centre(2, 2), 0.1) == g["cfl"][0]
    s.advance("sdirk3", 0.1, first=True)
    U1 = s.present_solution
    assert "%.6g" % oracle.enstrophy(om, U1) == g["enstrophy"][0]
    assert "%.6g" % oracle.kinetic_energy(om, U1) == g["kinetic_energy"][0]
    err = oracle.l2_error(om, U1, tg(0.1))[0]
    assert abs(err - float(g["l2_error_velocity"][0])) <= 5e-4 * err
    s.close()


def test_gmres_no_convergence_and_state_errors(oracle):
    from softx_2020_200_b200 import GlsnsError, NoConvergence
    mesh = oracle.BoxMesh(2, 8, 1, 1)
    force = mesh.evaluate_force(mms.forcing_2d)
    hp = hotpath_from_oracle_mesh(mesh, 1.0, force)
    with pytest.raises(GlsnsError) as e:        # solve before assemble
        hp.solve_linear_system()
    assert e.value.status == 5
    hp.set_vector("evaluation_point", np.zeros(mesh.ndof))
    hp.assemble(True)
    with pytest.raises(NoConvergence) as e:     # SolverControl::NoConvergence
        hp.solve_linear_system(relative_residual=1e-14, minimum_residual=1e-14, max_iterations=3)
    assert e.value.info["iterations"] == 3
    with pytest.raises(GlsnsError):             # umbrella scheme value
        hp.assemble(True, "sdirk2", [0.1])
    with pytest.raises(GlsnsError):             # transient without history
        hp.assemble(True, "bdf1", [0.1])
    hp.close()


def test_restart_01_through_the_cpp_mirror(oracle):
    """The reference's own solver-level test (tests/solvers/restart_01.cc) read through the mirrored
    C++ interface: GLSNavierStokesSolver + NewtonNonLinearSolver with all .prm defaults, host
    vectors in and out; the log lines of solve_system_GMRES carry the golden iteration counts."""
    import re
    from softx_2020_200_b200.mesh import BoxMesh
    from tests.mirror.solver import GLSNavierStokesSolver
    from tests.test_host_mirror import _match_numbering
    mesh = BoxMesh(2, 16, 1, 1, with_q_points=True)
    force = mms.forcing_2d(mesh.array("q_points").reshape(-1, 2)).reshape(mesh.n_cells, mesh.n_q, 2)
    s = GLSNavierStokesSolver(mesh, "", force)          # every .prm default
    s.set_vector("present_solution", np.zeros(mesh.n_dofs))
    s.solve_non_linear_system("steady", False, True)
    assert re.findall(r"-Iterative solver took : (\d+) steps", s.log) == ["8", "6", "10"]
    assert len(re.findall(r"-Tolerance of iterative solver is : ", s.log)) == 3
    assert len(re.findall(r"Newton iteration: \d+  - Residual:", s.log)) == 3
    nat = BoxMesh(2, 16, 1, 1, renumber=False)
    om = oracle.BoxMesh(2, 16, 1, 1, renumber=_match_numbering(nat, mesh, 2))
    err_u, _ = oracle.l2_error(om, s.present_solution, mms.exact_2d)
    assert float("%.6g" % err_u) == 0.0343628
    s.close()


def test_cpp_mirror_error_behaviour():
    """std::runtime_error('This solver is not allowed') for methods that are not built;
    SolverControl::NoConvergence when max iters is hit."""
    from softx_2020_200_b200 import NoConvergence
    from softx_2020_200_b200.mesh import BoxMesh
    from tests.mirror.solver import GLSNavierStokesSolver
    mesh = BoxMesh(2, 8, 1, 1, with_q_points=True)
    force = mms.forcing_2d(mesh.array("q_points").reshape(-1, 2)).reshape(mesh.n_cells, mesh.n_q, 2)
    s = GLSNavierStokesSolver(mesh, "subsection linear solver\n set method = amg\nend\n", force)
    s.set_vector("present_solution", np.zeros(mesh.n_dofs))
    with pytest.raises(RuntimeError, match="This solver is not allowed"):
        s.solve_non_linear_system("steady")
    s.close()
    s = GLSNavierStokesSolver(mesh, "subsection linear solver\n set max iters = 2\n"
                                    " set relative residual = 1e-14\nend\n", force)
    s.set_vector("present_solution", np.zeros(mesh.n_dofs))
    with pytest.raises(NoConvergence):
        s.solve_non_linear_system("steady")
    s.close()


def test_skip_newton_and_transient_through_the_cpp_mirror(oracle):
    """skip_newton reuses Jacobian + ILU (renewed_matrix=false path, gls_navier_stokes.cc:1270);
    one bdf1 step of the MMS problem matches the oracle's Newton solve of the same step."""
    from softx_2020_200_b200.mesh import BoxMesh
    from tests.mirror.solver import GLSNavierStokesSolver
    from tests.test_host_mirror import _match_numbering
    mesh = BoxMesh(2, 8, 2, 1, with_q_points=True)
    force = mms.forcing_2d(mesh.array("q_points").reshape(-1, 2)).reshape(mesh.n_cells, mesh.n_q, 2)
    prm = ("subsection non-linear solver\n set solver = skip_newton\n set skip iterations = 2\n"
           " set tolerance = 1e-9\n set max iterations = 30\nend\n"
           "subsection linear solver\n set relative residual = 1e-6\n set minimum residual = 1e-12\nend\n")
    s = GLSNavierStokesSolver(mesh, prm, force)
    s.set_time_steps([0.1, 0.1, 0.1, 0.1])
    z = np.zeros(mesh.n_dofs)
    s.set_vector("present_solution", z)
    s.set_vector("solution_m1", z)
    s.solve_non_linear_system("bdf1", False, True)
    nat = BoxMesh(2, 8, 2, 1, renumber=False)
    om = oracle.BoxMesh(2, 8, 2, 1, renumber=_match_numbering(nat, mesh, 2))
    pr = oracle.scheme_params("bdf1", [0.1], 1.0)
    U_ref, _, _ = oracle.newton_solve(om, z, pr, om.evaluate_force(mms.forcing_2d), tol=1e-9,
                                      max_it=30, lin=dict(rel=1e-6, abs_=1e-12), hist=(z, None, None))
    # all-Dirichlet problem: the pressure is defined up to a constant that no solver pins (the
    # reference compares it mean-free too, navier_stokes_base.cc:288-379), so shift it out
    U, is_p = s.present_solution.copy(), mesh.array("dof_component") == 2
    U[is_p] -= U[is_p].mean() - U_ref[is_p].mean()
    assert np.linalg.norm(U - U_ref) <= 1e-7 * np.linalg.norm(U_ref)
    s.close()


@pytest.mark.parametrize("name", ["case_2d_q2q1_bdf2", "case_3d_q1q1_steady", "case_3d_q2q2_steady"])
def test_against_committed_golden_fixtures(oracle, name):
    """The CUDA path against tests/golden/case_*.npz (oracle outputs frozen by make_fixtures.py): the
    oracle only rebuilds the mesh arrays here, no oracle arithmetic runs."""
    import os
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    mesh = oracle.BoxMesh(int(fx["dim"]), int(fx["n"]), int(fx["pu"]), int(fx["pp"]))
    assert np.array_equal(mesh.col, fx["col_idx"])
    nu, scheme = float(fx["viscosity"]), str(fx["scheme"])
    dts = list(fx["dts"]) or None
    hp = hotpath_from_oracle_mesh(mesh, nu, None)
    hp.set_vector("evaluation_point", fx["U"])
    if fx["U1"].size:
        hp.set_vector("solution_m1", fx["U1"])
        hp.set_vector("solution_m2", fx["U2"])
    hp.assemble(True, scheme, dts)
    assert row_scaled_error(mesh, hp.get_matrix_values(), fx["matrix"]) <= TOL_ENTRY
    b = hp.get_vector("system_rhs")
    assert np.max(np.abs(b - fx["rhs"])) <= TOL_ENTRY * np.max(np.abs(fx["rhs"]))
    assert np.max(np.abs(hp.spmv(fx["x"]) - fx["spmv"])) <= 1e-13 * np.max(np.abs(fx["spmv"]))
    hp.setup_ilu(0, 1e-8, 1.0)
    assert row_scaled_error(mesh, hp.get_ilu_values(), fx["ilu"]) <= 1e-11
    z = hp.ilu_apply(fx["x"])
    assert np.max(np.abs(z - fx["ilu_apply"])) <= 1e-10 * np.max(np.abs(fx["ilu_apply"]))
    dx, info = hp.solve_linear_system(relative_residual=1e-6, minimum_residual=1e-12,
                                      max_iterations=2000, ilu_atol=1e-8)
    assert abs(info["iterations"] - int(fx["gmres_iterations"])) <= 2
    assert info["true_residual"] <= info["tolerance"] * 1.0000001
    hp.close()


@pytest.mark.parametrize("n,world", [(3, 3), (4, 2)])
def test_rank_local_blocks_on_one_gpu(oracle, n, world):
    """What a rank of an N-rank run does between two exchanges, on one GPU and without a
    communicator: every rank's part of the partitioned cavity (owned rows + ghost columns) is
    attached to a context of its own and its assembled rows, its block of the block-Jacobi ILU(0)
    (Ifpack with overlap 0: ghost columns are outside the block), its SpMV with the ghost values
    supplied and both triangular sweeps are compared with the oracle on the same row blocks.
    (The exchanges themselves need N GPUs: tests/multi_gpu_check.py, bench.py's parity block.)"""
    from softx_2020_200_b200 import GLSHotPath
    from tests.util import rank_local_reference
    ref = rank_local_reference(oracle, n, world)
    om, bp, U, x = ref["oracle_mesh"], ref["block_ptr"], ref["state"], ref["x"]
    for r, m in enumerate(ref["parts"]):
        hp = GLSHotPath(0)
        m.attach(hp)
        hp.set_physics(0.005)
        l2g = m.array("local_to_global")
        hp.set_vector("present_solution", U[l2g])
        hp.set_vector("evaluation_point", U[l2g])
        hp.assemble(True)
        rows = slice(bp[r], bp[r + 1])
        b = hp.get_vector("system_rhs")
        assert np.max(np.abs(b - ref["rhs"][rows])) <= TOL_ENTRY * np.max(np.abs(ref["rhs"]))
        a_loc = hp.get_matrix_values()
        hp.setup_ilu(0, 1e-12, 1.0)
        lu_loc = hp.get_ilu_values()
        col_g, rp = l2g[m.array("col_idx")], m.array("row_ptr")
        err_a = err_lu = 0.0
        for i in range(m.n_owned):
            o = np.argsort(col_g[rp[i]:rp[i + 1]])
            gi = bp[r] + i
            ref_rows = slice(om.rowptr[gi], om.rowptr[gi + 1])
            a_ref, lu_ref, c = ref["matrix"][ref_rows], ref["ilu"][ref_rows], om.col[ref_rows]
            assert np.array_equal(col_g[rp[i]:rp[i + 1]][o], c)
            scale = max(np.max(np.abs(a_ref)), 1e-300)
            err_a = max(err_a, np.max(np.abs(a_loc[rp[i]:rp[i + 1]][o] - a_ref)) / scale)
            inb = (c >= bp[r]) & (c < bp[r + 1])  # the diagonal block: what the factorisation touches
            err_lu = max(err_lu, np.max(np.abs(lu_loc[rp[i]:rp[i + 1]][o][inb] - lu_ref[inb])) /
                         max(np.max(np.abs(lu_ref[inb])), 1e-300))
        assert err_a <= TOL_ENTRY
        # (nu = 0.005 and a 1e-12 shift: the factors of this matrix carry 1e-10 of rounding where
        # those of the nu = 0.1 cases above carry 1e-12; measured 1.4e-10 on the B200)
        assert err_lu <= 1e-8
        # SpMV and the triangular sweeps on the DEVICE's own values, by the oracle's loops over
        # the rank's local arrays (ghost columns sit behind the owned ones and are outside the
        # block), as tools/trsv_sweep.py does for the single-rank case
        loc = types.SimpleNamespace(ndof=m.n_owned, rowptr=rp, col=m.array("col_idx"))
        y_ref = oracle.spmv(loc, a_loc, x[l2g])
        y = hp.spmv(x[l2g])
        assert np.max(np.abs(y - y_ref)) <= 1e-13 * np.max(np.abs(y_ref))
        assert np.max(np.abs(y - ref["spmv"][rows])) <= 1e-9 * np.max(np.abs(ref["spmv"]))
        dp = np.array([rp[i] + np.searchsorted(loc.col[rp[i]:rp[i + 1]], i) for i in range(m.n_owned)],
                      dtype=np.int64)
        z_ref = oracle.ilu_apply(loc, lu_loc, dp, x[rows])
        z = hp.ilu_apply(x[rows])
        assert np.max(np.abs(z - z_ref)) <= 1e-10 * np.max(np.abs(z_ref))
        hp.close()
    ref["global_mesh"].close()


def test_two_gpu_newton_step_matches_block_jacobi_oracle():
    """N = 2 ranks over NCCL (skipped on a 1-GPU box; run with `gpurun --gpus 2`)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                          "29731", os.path.join(root, "tests", "multi_gpu_check.py"), "4"],
                         capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
