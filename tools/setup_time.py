"""Host/device setup cost by mesh size: python tools/setup_time.py N"""
import sys, time, json
sys.path.insert(0, ".")
from softx_2020_200_b200 import GLSHotPath
from softx_2020_200_b200.mesh import BoxMesh
n = int(sys.argv[1])
CAVITY = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"), (3, "function", (1, 0, 0))]
t = time.time(); m = BoxMesh(3, n, 2, 2, bcs=CAVITY); t_mesh = time.time() - t
hp = GLSHotPath(0)
t = time.time(); m.attach(hp); t_attach = time.time() - t
hp.set_physics(0.005)
U0 = m.initial_state(); hp.set_vector("evaluation_point", U0); hp.assemble(True)
t = time.time(); hp.setup_ilu(0, 1e-12, 1.0); t_ilu = time.time() - t
ms = hp.time_kernel("ilu_apply", reps=5)
by = 12 * m.nnz + 40 * m.n_dofs
print(json.dumps(dict(n=n, ndof=m.n_dofs, t_mesh=t_mesh, t_attach=t_attach, t_setup_ilu=t_ilu, levels=hp.ilu_levels(),
                      ilu_apply_ms=ms, frac=by / ms / 1e6 / 6546.2, spmv_ms=hp.time_kernel("spmv", reps=5))))
print(json.dumps(dict(ilu_factor_ms=hp.time_kernel("ilu_factor", reps=2))))
