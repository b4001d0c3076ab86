"""Generates the committed golden fixtures of tests/golden/ (run from the repo root:
    python tests/golden/make_fixtures.py).

* reference_golden.json — numbers copied from the reference's own checked-in test outputs
  (file:line given per entry); tests/test_oracle_golden.py pins the CPU oracle against them.
* case_*.npz — inputs and oracle outputs of small seeded cases (state, assembled matrix and
  right-hand side, ILU(0) factors, one ILU application, the GMRES solve), so that the GPU parity
  tests have fixed vectors to meet that do not depend on re-running the oracle.
The oracle is test infrastructure (oracle/gls_oracle.c header); nothing here ships."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference_port as R  # noqa: E402

REFERENCE_GOLDEN = {
    "restart_01": {"source": "tests/solvers/restart_01.output:2-8",
                   "gmres_iterations": [8, 6, 10],
                   "true_residuals": [0.00204885, 9.85227e-05, 3.32384e-08],
                   "l2_error_velocity": 0.0343628, "l2_error_after_zeroing": 0.612372},
    "mms3d_gls": {"source": "applications_tests/gls_navier_stokes_3d/mms3d_gls.output:13-18",
                  "cells": [64, 512], "dofs": [500, 2916],
                  "error_velocity": [5.4021e-01, 1.3126e-01], "error_pressure": [4.5537e-02, 1.7717e-01]},
    "mms2d_gls": {"source": "applications_tests/gls_navier_stokes_2d/mms2d_gls.output:24-26",
                  "cells": [64, 256, 1024], "dofs": [243, 867, 3267],
                  "error_velocity": [1.3284e-01, 3.4363e-02, 8.7362e-03],
                  "error_pressure": [1.7844e-01, 9.7118e-02, 3.0300e-02]},
    "bdf_01": {"source": "tests/core/bdf_01.output", "time_steps": [0.1, 0.2, 0.3],
               "order1": [10.0, -10.0], "order2": [13.3333, -15.0, 1.66667],
               "order3": [15.0, -18.0, 3.33333, -0.333333]},
}

CASES = {  # name: (dim, n, pu, pp, viscosity, scheme, dts)
    "case_2d_q2q1_bdf2": (2, 4, 2, 1, 0.05, "bdf2", [0.1, 0.2, 0.3]),
    "case_3d_q1q1_steady": (3, 3, 1, 1, 0.37, "steady", None),
    "case_3d_q2q2_steady": (3, 2, 2, 2, 0.2, "steady", None),
}


def main():
    R.lib().glso_set_num_threads(1)
    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(REFERENCE_GOLDEN, f, indent=1)
    for name, (dim, n, pu, pp, nu, scheme, dts) in CASES.items():
        mesh = R.BoxMesh(dim, n, pu, pp)
        rng = np.random.default_rng(20201)
        U = mesh.apply_nonzero_constraints(0.4 * rng.uniform(-1, 1, mesh.ndof))
        U[mesh.constrained != 0] = 0.0
        hist = [0.4 * rng.uniform(-1, 1, mesh.ndof) for _ in range(2)] if scheme != "steady" else [None, None]
        pr = R.scheme_params(scheme, dts, nu)
        val, rhs = R.assemble(mesh, U, pr, True, None, hist[0], hist[1], None)
        lu, dp = R.ilu0(mesh, val, 1e-8, 1.0)
        x = rng.standard_normal(mesh.ndof)
        z = R.ilu_apply(mesh, lu, dp, x)
        y = R.spmv(mesh, val, x)
        dx, its, true_res = R.solve_linear_system(mesh, val, rhs, rel=1e-6, abs_=1e-12, max_iters=2000,
                                                  ilu_atol=1e-8)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), dim=dim, n=n, pu=pu, pp=pp, viscosity=nu,
                            scheme=scheme, dts=np.array(dts if dts else []), U=U,
                            U1=hist[0] if hist[0] is not None else np.zeros(0),
                            U2=hist[1] if hist[1] is not None else np.zeros(0),
                            matrix=val, rhs=rhs, ilu=lu, x=x, ilu_apply=z, spmv=y, update=dx,
                            gmres_iterations=its, true_residual=true_res,
                            row_ptr=mesh.rowptr, col_idx=mesh.col)
        print(name, "ndof", mesh.ndof, "nnz", len(val), "gmres", its)


if __name__ == "__main__":
    main()
