"""A short program for ncu: every hot kernel of the path launched a few times on the 3D Q2-Q2 cavity.
    python tools/profile_kernels.py N [kernels...]      (kernels: spmv ilu_apply ilu_factor assemble_system assemble_rhs orthog)
Prints the CUDA-event time of each (not a bench number when run under ncu)."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
from softx_2020_200_b200 import GLSHotPath
from softx_2020_200_b200.mesh import BoxMesh
n = int(sys.argv[1])
which = sys.argv[2:] or ["spmv", "ilu_apply", "ilu_factor", "assemble_system", "assemble_rhs", "orthog"]
CAVITY = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"), (3, "function", (1, 0, 0))]
m = BoxMesh(3, n, 2, 2, bcs=CAVITY); hp = GLSHotPath(0); m.attach(hp); hp.set_physics(0.005)
xyz = m.array("dof_coords").reshape(-1, 3)
U = m.initial_state() + np.where(m.array("constrained") != 0, 0.0, 0.05 * np.sin(np.pi * xyz[:, 0]) * np.cos(np.pi * xyz[:, 1]))
hp.set_vector("present_solution", U); hp.set_vector("evaluation_point", U); hp.assemble(True); hp.setup_ilu(0, 1e-12, 1.0)
out = dict(n=n, ndof=m.n_dofs, nnz=m.nnz, n_cells=m.n_cells)
for k in which:
    out[k + "_ms"] = hp.time_kernel(k, reps=2, nvec=15)
print(json.dumps(out))
