"""ILU(0) factorisation kernel check on the 3D Q2-Q2 cavity: time and a digest of the factors, so
that two kernel variants (run in separate processes: GLSNS_ILU_BY_PIVOT_ROWS=1 selects the
row-wise pivot kernel) can be compared bit for bit.   python tools/ilu_factor_check.py N"""
import hashlib, json, os, sys
sys.path.insert(0, ".")
import numpy as np
from softx_2020_200_b200 import GLSHotPath
from softx_2020_200_b200.mesh import BoxMesh
n = int(sys.argv[1])
CAVITY = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"), (3, "function", (1, 0, 0))]
m = BoxMesh(3, n, 2, 2, bcs=CAVITY); hp = GLSHotPath(0); m.attach(hp); hp.set_physics(0.005)
xyz = m.array("dof_coords").reshape(-1, 3)
U = m.initial_state() + np.where(m.array("constrained") != 0, 0.0, 0.05 * np.sin(np.pi * xyz[:, 0]) * np.cos(np.pi * xyz[:, 1]))
hp.set_vector("evaluation_point", U); hp.assemble(True); hp.setup_ilu(0, 1e-12, 1.0)
lu = hp.get_ilu_values()
print(json.dumps(dict(n=n, variant=os.environ.get("GLSNS_ILU_BY_PIVOT_ROWS", "runs"),
                      sha=hashlib.sha256(lu.tobytes()).hexdigest()[:16], finite=bool(np.isfinite(lu).all()),
                      ilu_factor_ms=hp.time_kernel("ilu_factor", reps=3))))
