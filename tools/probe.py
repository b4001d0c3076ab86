"""Kernel-level probe on the 3D Q2-Q2 cavity: per-kernel times, roofline fractions, one Newton step.
    python tools/probe.py N [max_iters]"""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
from softx_2020_200_b200 import GLSHotPath, NoConvergence
from softx_2020_200_b200.mesh import BoxMesh

n = int(sys.argv[1]); max_it = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
CAVITY = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"), (3, "function", (1, 0, 0))]
t = time.time(); m = BoxMesh(3, n, 2, 2, bcs=CAVITY); t_mesh = time.time() - t
hp = GLSHotPath(0)
t = time.time(); m.attach(hp); t_attach = time.time() - t
hp.set_physics(0.005)
U0 = m.initial_state()
hp.set_vector("present_solution", U0); hp.set_vector("evaluation_point", U0)
N, nnz = m.n_dofs, m.nnz
out = dict(n=n, ndof=N, nnz=nnz, t_mesh=t_mesh, t_attach=t_attach, levels=hp.ilu_levels())
hp.assemble(True)
out["rhs_norm0"] = hp.rhs_norm()
t = time.time(); hp.setup_ilu(0, 1e-12, 1.0); out["t_ilu_factor_first"] = time.time() - t
HBM = 6546.2
for k, by in (("spmv", 12 * nnz + 24 * N), ("ilu_apply", 12 * nnz + 40 * N), ("orthog", (4 * 16 + 6) * 8 * N),
              ("assemble_system", None), ("assemble_rhs", None), ("ilu_factor", None)):
    ms = hp.time_kernel(k, reps=5 if k != "ilu_factor" else 2, nvec=15)
    out[k + "_ms"] = ms
    if by: out[k + "_GBs"] = by / ms / 1e6; out[k + "_frac"] = by / ms / 1e6 / HBM
hp.assemble(True)
hp.reset_timers()
t = time.time()
try:
    _, info = hp.solve_linear_system(1e-4, 1e-9, max_it, 30, 0, 1e-12, 1.0, download=False)
except NoConvergence as e:
    info = e.info; info["no_convergence"] = True
out["t_solve_wall"] = time.time() - t
out["solve"] = info
out["timers"] = hp.timers()
print(json.dumps(out, indent=1))
