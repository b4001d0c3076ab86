"""ILU-apply kernel check + timing on the 3D Q2-Q2 cavity (run under gpurun).
    [GLSNS_TRSV_TEAMS=t GLSNS_TRSV_HELPERS=k GLSNS_TRSV_TUNE=r] python tools/trsv_sweep.py N [check]
`check` compares the device triangular solves with the CPU oracle's on the device's own factors."""
import json, os, sys, types
sys.path.insert(0, ".")
import numpy as np
from softx_2020_200_b200 import GLSHotPath
from softx_2020_200_b200.mesh import BoxMesh

n = int(sys.argv[1])
check = len(sys.argv) > 2 and sys.argv[2] == "check"
CAVITY = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"), (3, "function", (1, 0, 0))]
m = BoxMesh(3, n, 2, 2, bcs=CAVITY); hp = GLSHotPath(0); m.attach(hp); hp.set_physics(0.005)
U0 = m.initial_state(); hp.set_vector("evaluation_point", U0); hp.assemble(True); hp.setup_ilu(0, 1e-12, 1.0)
out = dict(n=n, teams=os.environ.get("GLSNS_TRSV_TEAMS"), helpers=os.environ.get("GLSNS_TRSV_HELPERS"),
           levels=hp.ilu_levels())
if check:
    from oracle import reference_port as R
    R.lib().glso_set_num_threads(1)
    om = types.SimpleNamespace(ndof=m.n_dofs, rowptr=m.array("row_ptr"), col=m.array("col_idx"))
    lu = hp.get_ilu_values()
    rows = np.arange(m.n_dofs)
    dp = np.array([om.rowptr[i] + np.searchsorted(om.col[om.rowptr[i]:om.rowptr[i + 1]], i) for i in rows],
                  dtype=np.int64)
    x = np.random.default_rng(5).standard_normal(m.n_dofs)
    z_ref = R.ilu_apply(om, lu, dp, x)
    z = hp.ilu_apply(x)
    out["apply_err"] = float(np.max(np.abs(z - z_ref)) / np.max(np.abs(z_ref)))
ms = hp.time_kernel("ilu_apply", reps=10)
by = 12 * m.nnz + 40 * m.n_dofs
out.update(ilu_apply_ms=ms, GBs=by / ms / 1e6, frac=by / ms / 1e6 / 6546.2, spmv_ms=hp.time_kernel("spmv", reps=10))
print(json.dumps(out))
