import json, sys, time, os
sys.path.insert(0, ".")
import numpy as np
from softx_2020_200_b200 import GLSHotPath
from softx_2020_200_b200.mesh import BoxMesh
n = int(sys.argv[1])
CAVITY = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"), (3, "function", (1, 0, 0))]
m = BoxMesh(3, n, 2, 2, bcs=CAVITY); hp = GLSHotPath(0); m.attach(hp); hp.set_physics(0.005)
U0 = m.initial_state(); hp.set_vector("evaluation_point", U0); hp.assemble(True); hp.setup_ilu(0, 1e-12, 1.0)
print(os.environ.get("GLSNS_TRSV_CTAS"), os.environ.get("GLSNS_TRSV_SLEEP"), "levels", hp.ilu_levels(), "ilu_apply ms", hp.time_kernel("ilu_apply", reps=5), "spmv", hp.time_kernel("spmv", reps=5))
