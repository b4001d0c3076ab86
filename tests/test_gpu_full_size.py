"""Parity at BASELINE.json's full size (3D Q2-Q2 cavity, 64^3 cells, 8.59 M dofs, 2.06e9 non-zeros)
through size-independent properties, where the CPU oracle is too slow to be the checker:

  * SpMV: linearity, and agreement with a host CSR product on sampled rows;
  * ILU(0): the defining property (L U)_ij = A'_ij on the pattern (A' = A with the Ifpack diagonal
    perturbation), on sampled rows;
  * ILU application: x -> r = L (U x) (two device SpMVs with the factors loaded as matrices)
    -> z = (LU)^-1 r gives x back (the round trip of the triangular solves);
  * GMRES: the logged residual is the true one, recomputed on the host from sampled rows.

OPT-IN (GLSNS_FULL_SIZE_CELLS=64, or 32 for a quick run): it holds several 16 GB host arrays and
takes about two minutes, and it has not yet been run on a GPU box (written at the end of round 1
after the GPU budget was spent; its host-side checks were exercised on the CPU with the oracle's
factors in place of the device's) -- enable it in the default run once it has passed there."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CELLS = int(os.environ.get("GLSNS_FULL_SIZE_CELLS", "0"))
CAVITY = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"),
          (3, "function", (1.0, 0.0, 0.0))]


def _rows_product(rp, col, val, x, rows):
    return np.array([val[rp[i]:rp[i + 1]] @ x[col[rp[i]:rp[i + 1]]] for i in rows])


@pytest.mark.skipif(CELLS == 0, reason="opt-in: set GLSNS_FULL_SIZE_CELLS=64")
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
