"""Parity at BASELINE.json's full size (3D Q2-Q2 cavity, 64^3 cells, 8.59 M dofs, 2.06e9 non-zeros)
through size-independent properties, where the CPU oracle is too slow to be the checker:

  * SpMV: linearity, and agreement with a host CSR product on sampled rows;
  * ILU(0): the defining property (L U)_ij = A'_ij on the pattern (A' = A with the Ifpack diagonal
    perturbation), on sampled rows;
  * ILU application: x -> r = L (U x) (two device SpMVs with the factors loaded as matrices)
    -> z = (LU)^-1 r gives x back (the round trip of the triangular solves);
  * GMRES: the logged residual is the true one, recomputed on the host from sampled rows.

Runs at 32^3 cells (1.10 M dofs, 2.8e8 non-zeros, 11 s on a B200) by default;
GLSNS_FULL_SIZE_CELLS=64 runs the bench mesh itself (several 16 GB host arrays, about two
minutes; profiles/r2_full_size_n64.log keeps the log of such a run)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CELLS = int(os.environ.get("GLSNS_FULL_SIZE_CELLS", "32"))
CAVITY = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"),
          (3, "function", (1.0, 0.0, 0.0))]


def _rows_product(rp, col, val, x, rows):
    return np.array([val[rp[i]:rp[i + 1]] @ x[col[rp[i]:rp[i + 1]]] for i in rows])


def test_full_size_properties():
    from softx_2020_200_b200 import GLSHotPath
    from softx_2020_200_b200.mesh import BoxMesh
    rng = np.random.default_rng(2024)
    mesh = BoxMesh(3, CELLS, 2, 2, bcs=CAVITY)
    hp = GLSHotPath(0)
    mesh.attach(hp)
    hp.set_physics(0.005)
    N = mesh.n_dofs
    rp, col = mesh.array("row_ptr"), mesh.array("col_idx")
    xyz = mesh.array("dof_coords").reshape(-1, 3)
    U = mesh.initial_state() + np.where(mesh.array("constrained") != 0, 0.0,
                                        0.05 * np.sin(np.pi * xyz[:, 0]) * np.cos(np.pi * xyz[:, 1]))
    hp.set_vector("present_solution", U)
    hp.set_vector("evaluation_point", U)
    hp.assemble(True)
    a = hp.get_matrix_values()
    rows = rng.choice(N, 2000, replace=False)

    # ---- SpMV ----
    x, y = rng.standard_normal(N), rng.standard_normal(N)
    ax, ay = hp.spmv(x), hp.spmv(y)
    lin = hp.spmv(0.7 * x - 1.3 * y)
    scale = np.max(np.abs(ax)) + np.max(np.abs(ay))
    assert np.max(np.abs(lin - (0.7 * ax - 1.3 * ay))) <= 1e-12 * scale
    ref = _rows_product(rp, col, a, x, rows)
    assert np.max(np.abs(ax[rows] - ref)) <= 1e-12 * np.max(np.abs(ref))

    # ---- ILU(0): (L U)_ij = A'_ij on the pattern ----
    atol = 1e-12
    hp.setup_ilu(0, atol, 1.0)
    lu = hp.get_ilu_values()
    worst = 0.0
    for i in rows[:300]:
        ci, li = col[rp[i]:rp[i + 1]], lu[rp[i]:rp[i + 1]]
        ai = a[rp[i]:rp[i + 1]].copy()
        d = int(np.searchsorted(ci, i))
        ai[d] += atol if ai[d] >= 0 else -atol            # Ifpack: d <- rtol d + sign(d) atol
        prod = np.zeros(len(ci))
        for t in range(len(ci)):                          # k = ci[t]: L_ik (k < i, unit diagonal at k = i)
            k = ci[t]
            if k > i:
                break
            lik = li[t] if k < i else 1.0
            ck, lk = col[rp[k]:rp[k + 1]], lu[rp[k]:rp[k + 1]]
            up = ck >= k                                  # row k of U
            pos = np.searchsorted(ci, ck[up])
            ok = (pos < len(ci)) & (ci[np.minimum(pos, len(ci) - 1)] == ck[up])
            prod[pos[ok]] += lik * lk[up][ok]
        worst = max(worst, float(np.max(np.abs(prod - ai)) / np.max(np.abs(ai))))
    assert worst <= 1e-10

    # ---- ILU application: z = (LU)^-1 (L (U x)) = x ----
    row_of = np.repeat(np.arange(N, dtype=np.int32), np.diff(rp))
    upper = col >= row_of
    diag = col == row_of
    del row_of
    hp.set_matrix_values(np.where(upper, lu, 0.0))
    w = hp.spmv(x)
    hp.set_matrix_values(np.where(diag, 1.0, np.where(upper, 0.0, lu)))
    r = hp.spmv(w)
    del upper, diag
    hp.set_matrix_values(a)                               # (loading values drops the factors:
    hp.setup_ilu(0, atol, 1.0)                            #  factorise A again, same factors)
    assert np.array_equal(hp.get_ilu_values(), lu), "the factorisation is not reproducible"
    z = hp.ilu_apply(r)
    assert np.linalg.norm(z - x) <= 1e-8 * np.linalg.norm(x)

    # ---- GMRES: the logged residual is the true one ----
    hp.assemble(True)                                     # (set_matrix_values invalidated the ILU)
    b = hp.get_vector("system_rhs")
    dx, info = hp.solve_linear_system(relative_residual=1e-4, minimum_residual=1e-9,
                                      max_iterations=5000, ilu_atol=atol)
    assert info["true_residual"] <= info["tolerance"] * 1.01
    # the sampled rows of b - A x (host product; constrained rows are diagonal with a zero
    # right-hand side, so zero_constraints.distribute changed nothing) stay below the global norm
    res_rows = b[rows] - _rows_product(rp, col, a, dx, rows)
    assert np.linalg.norm(res_rows) <= info["true_residual"] * 1.01 + 1e-14
    hp.close()


def test_bench_mesh_n16_newton_step_against_the_oracle(oracle):
    """The bench workload at 16^3 cells (143 748 dofs, 3.4e7 non-zeros, a 1334-level triangular
    solve) against the oracle: matrix and right-hand side entries, ILU(0) factors and their
    application, the GMRES iteration count of the bench's own solver settings (+-2, north_star),
    and the solution of the linear system solved tightly by both."""
    import time
    from tests.util import hotpath_from_oracle_mesh, row_scaled_error
    lid = lambda x: np.stack([np.ones(len(x)), 0 * x[:, 0], 0 * x[:, 0]], axis=1)
    bcs = {0: ("noslip",), 1: ("noslip",), 2: ("noslip",), 4: ("noslip",), 5: ("noslip",),
           3: ("function", lid)}
    mesh = oracle.BoxMesh(3, 16, 2, 2, bcs=bcs)
    hp = hotpath_from_oracle_mesh(mesh, 0.005, None)
    x = mesh.dof_coords
    U = 0.05 * np.sin(np.pi * x[:, 0] + 0.3 * mesh.dof_comp) * np.cos(np.pi * x[:, 1]) * \
        np.cos(0.5 * np.pi * x[:, 2])
    U = mesh.apply_nonzero_constraints(np.where(mesh.constrained != 0, 0.0, U))
    pr = oracle.scheme_params("steady", None, 0.005)
    threads = os.cpu_count() or 1
    a_ref, b_ref = oracle.assemble(mesh, U, pr, True, threads=threads)   # (coloured: thread-count independent)
    oracle.lib().glso_set_num_threads(1)
    hp.set_vector("present_solution", U)
    hp.set_vector("evaluation_point", U)
    hp.assemble(True)
    assert row_scaled_error(mesh, hp.get_matrix_values(), a_ref) <= 1e-12
    b = hp.get_vector("system_rhs")
    assert np.max(np.abs(b - b_ref)) <= 1e-12 * np.max(np.abs(b_ref))
    lu_ref, dp = oracle.ilu0(mesh, a_ref, 1e-12, 1.0)
    hp.setup_ilu(0, 1e-12, 1.0)
    assert row_scaled_error(mesh, hp.get_ilu_values(), lu_ref) <= 1e-11
    r = np.random.default_rng(3).standard_normal(mesh.ndof)
    z_ref = oracle.ilu_apply(mesh, lu_ref, dp, r)
    assert np.max(np.abs(hp.ilu_apply(r) - z_ref)) <= 1e-10 * np.max(np.abs(z_ref))
    # the bench's solver settings (examples/01-cavity/cavity.prm:88-94)
    dx_ref, its_ref, _ = oracle.solve_linear_system(mesh, a_ref, b_ref, rel=1e-4, abs_=1e-9,
                                                    max_iters=5000, ilu_atol=1e-12)
    dx, info = hp.solve_linear_system(relative_residual=1e-4, minimum_residual=1e-9,
                                      max_iterations=5000, ilu_atol=1e-12)
    assert abs(info["iterations"] - its_ref) <= 2, (info["iterations"], its_ref)
    assert info["true_residual"] <= info["tolerance"] * 1.01
    # and solved tightly: the two solutions of the same system
    dx_ref, its_ref, _ = oracle.solve_linear_system(mesh, a_ref, b_ref, rel=1e-12, abs_=1e-30,
                                                    max_iters=5000, ilu_atol=1e-12)
    dx, info = hp.solve_linear_system(relative_residual=1e-12, minimum_residual=1e-30,
                                      max_iterations=5000, ilu_atol=1e-12)
    assert abs(info["iterations"] - its_ref) <= 2, (info["iterations"], its_ref)
    dx_ref[mesh.constrained != 0] = 0.0
    err = np.linalg.norm(dx - dx_ref) / np.linalg.norm(dx_ref)
    print("n=16 Newton step: GMRES %d / oracle %d iterations, update err %.2e" %
          (info["iterations"], its_ref, err))
    assert err <= 1e-8
    hp.close()


def test_second_set_mesh_on_one_context_with_a_larger_mesh(oracle):
    """glsns_set_mesh "after every setup_dofs" (include/glsns.h): a context that has solved on one
    mesh is given a larger one (refinement) and must size every work vector -- the Krylov basis
    included -- to the new mesh."""
    from tests import mms
    from tests.util import hotpath_from_oracle_mesh
    small, large = oracle.BoxMesh(3, 2, 2, 2), oracle.BoxMesh(3, 4, 2, 2)
    hp = hotpath_from_oracle_mesh(small, 0.1, small.evaluate_force(mms.forcing_3d))
    hp.set_vector("evaluation_point", np.zeros(small.ndof))
    hp.assemble(True)
    hp.solve_linear_system(relative_residual=1e-8, minimum_residual=1e-12)
    # the same context, the larger mesh
    fe = large.fe
    hp.set_fe(large.dim, large.pu, fe.Nu, fe.dNu, fe.d2Nu, fe.Np, fe.dNp, fe.wq)
    ptr, order, _ = large.color_lists()
    hp.set_mesh(large.ndof, large.cell_dofs, large.cell_invJ, large.cell_detJ, large.cell_measure,
                large.constrained, large.rowptr, large.col, ptr, order, q_points=large.qpoints,
                constraint_values=large.constraint_value)
    hp.set_physics(0.1)
    force = large.evaluate_force(mms.forcing_3d)
    hp.set_forcing(force)
    U = np.zeros(large.ndof)
    hp.set_vector("evaluation_point", U)
    hp.assemble(True)
    a_ref, b_ref = oracle.assemble(large, U, oracle.scheme_params("steady", None, 0.1), True, force)
    x_ref, its_ref, _ = oracle.solve_linear_system(large, a_ref, b_ref, rel=1e-8, abs_=1e-12)
    x, info = hp.solve_linear_system(relative_residual=1e-8, minimum_residual=1e-12)
    assert abs(info["iterations"] - its_ref) <= 2
    x_ref[large.constrained != 0] = 0.0
    assert np.linalg.norm(x - x_ref) <= 1e-6 * np.linalg.norm(x_ref)
    hp.close()
