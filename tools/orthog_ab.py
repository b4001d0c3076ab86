"""CGS2 orthogonalisation, fused (axpy_dots_kernel + scaled_axpy_kernel) against the four separate kernels,
in one process on one mesh:   python tools/orthog_ab.py N"""
import json, os, sys
sys.path.insert(0, ".")
from softx_2020_200_b200 import GLSHotPath
from softx_2020_200_b200.mesh import BoxMesh
n = int(sys.argv[1])
CAVITY = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"), (3, "function", (1, 0, 0))]
m = BoxMesh(3, n, 2, 2, bcs=CAVITY)
out = dict(n=n, ndof=m.n_dofs)
for fused in ("1", "0"):
    os.environ["GLSNS_GMRES_FUSED"] = fused
    hp = GLSHotPath(0); m.attach(hp); hp.set_physics(0.005)
    hp.set_vector("evaluation_point", m.initial_state()); hp.assemble(True)
    for nv in (8, 15, 30):
        ms = hp.time_kernel("orthog", reps=5, nvec=nv)
        passes = (3 * nv + 5) if fused == "1" else (4 * nv + 8)
        out["fused%s_nv%d" % (fused, nv)] = dict(ms=ms, GBs=passes * 8 * m.n_dofs / ms / 1e6)
    del hp
print(json.dumps(out))
