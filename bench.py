#!/usr/bin/env python
"""bench.py — Newton-step throughput of the GLS Navier–Stokes hot path on the 3D Q2-Q2 cavity.

Metric (BASELINE.json): MDoF/s per Newton step (assembly + GMRES).  A "step" is ONE Newton
iteration of NewtonNonLinearSolver::solve (reference: include/core/newton_non_linear_solver.h:90-138)
at a fixed, non-trivial linearisation point (the state after the first Newton update from rest):
  assemble_matrix_and_rhs -> setup_ILU -> GMRES(30) solve -> line search (assemble_rhs + l2_norm
  per trial, alpha = 1, 1/2, ... until the residual drops below 0.9 x the previous one).
The state is reset (device copy, untimed) before every step so all K steps do identical work.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--cells n] [--impl ours|reference]

`value` is measured with the state resident in HBM, timed with CUDA events on the library's stream
(the per-phase device timers of the C ABI); `e2e` is the same step driven through host buffers
(H2D of evaluation_point for every assembly, D2H of newton_update, host-side line-search update).
`--impl reference` times the CPU restatement of the reference path (oracle/, the reference's
deal.II/Trilinos build is impossible in this image) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CAVITY = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"),
          (3, "function", (1.0, 0.0, 0.0))]
NU = 0.005                                   # Re = 400 with U = 1, L = 2
LIN = dict(relative_residual=1e-4, minimum_residual=1e-9, max_iterations=5000, restart=30,
           ilu_fill=0, ilu_atol=1e-12, ilu_rtol=1.0)   # examples/01-cavity/cavity.prm:88-94
METRIC = "MDoF/s per Newton step (assembly+GMRES), 3D cavity Q2-Q2"
CPU_SAMPLE_CELLS = 16


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    except OSError:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(r) > 2 + i and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference path, all host threads, one MPI-rank-like
# block per thread (block-Jacobi ILU as Ifpack with overlap 0 gives on that many ranks)
# --------------------------------------------------------------------------------------------
def cpu_newton_step_setup(n_cells, threads):
    import numpy as np
    from oracle import reference_port as R
    R.lib().glso_set_num_threads(threads)
    lid = lambda x: np.stack([np.ones(len(x)), 0 * x[:, 0], 0 * x[:, 0]], axis=1)
    bcs = {0: ("noslip",), 1: ("noslip",), 2: ("noslip",), 4: ("noslip",), 5: ("noslip",),
           3: ("function", lid)}
    mesh = R.BoxMesh(3, n_cells, 2, 2, bcs=bcs)
    pr = R.scheme_params("steady", None, NU)
    # contiguous row blocks with equal nonzeros: one per thread
    bp = np.searchsorted(mesh.rowptr, np.arange(threads + 1) * (mesh.rowptr[-1] // threads))
    bp[0], bp[-1] = 0, mesh.ndof
    state = dict(R=R, mesh=mesh, pr=pr, bp=bp.astype(np.int64), threads=threads)
    U0 = mesh.apply_nonzero_constraints(np.zeros(mesh.ndof))
    state["U1"], _ = cpu_newton_step(state, U0)          # the linearisation point of the step
    return state


def cpu_newton_step(st, U):
    import numpy as np
    R, mesh, pr = st["R"], st["mesh"], st["pr"]
    val, rhs = R.assemble(mesh, U, pr, True, threads=st["threads"])
    last = float(np.linalg.norm(rhs))
    tol = max(LIN["relative_residual"] * last, LIN["minimum_residual"])
    lu, dp = R.ilu0(mesh, val, LIN["ilu_atol"], LIN["ilu_rtol"], st["bp"])
    dx, its, res, ok, _ = R.gmres(mesh, val, lu, dp, rhs, tol, LIN["max_iterations"],
                                  LIN["restart"], st["bp"])
    dx[mesh.constrained != 0] = 0.0
    alpha = 1.0
    while alpha > 1e-3:
        Un = mesh.apply_nonzero_constraints(U + alpha * dx)
        _, r2 = R.assemble(mesh, Un, pr, False, threads=st["threads"])
        if float(np.linalg.norm(r2)) < 0.9 * last:
            break
        alpha *= 0.5
    return Un, its


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    st = cpu_newton_step_setup(args.cpu_cells, threads)
    its = 0
    for _ in range(args.warmup):
        _, its = cpu_newton_step(st, st["U1"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, its = cpu_newton_step(st, st["U1"])
    dt = (time.perf_counter() - t0) / args.steps
    ndof = st["mesh"].ndof
    v = ndof / dt / 1e6
    sample = ("same cavity at n=%d (%d DoFs), Newton iteration 1, %d GMRES iterations, %d threads "
              "= %d block-Jacobi ILU blocks" % (args.cpu_cells, ndof, its, threads, threads))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "MDoF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "3D lid-driven cavity Q2-Q2 Re=400 steady, one Newton iteration "
                               "(CPU restatement of the reference path; Trilinos unavailable)",
                   "cells_per_dir": args.cpu_cells, "n_dofs": ndof, "gmres_iterations": its},
        "cpu_baseline": {"value": v, "unit": "MDoF/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": "MDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def newton_step_device(hp, U1_set):
    """One Newton iteration, state resident in HBM. Returns (gmres iterations, line-search trials)."""
    U1_set()
    hp.assemble(True)
    last = hp.rhs_norm()
    _, info = hp.solve_linear_system(download=False, **LIN)
    alpha, trials = 1.0, 0
    while alpha > 1e-3:
        hp.line_search_point(alpha)
        hp.assemble(False)
        trials += 1
        if hp.rhs_norm() < 0.9 * last:
            break
        alpha *= 0.5
    return info["iterations"], trials


def newton_step_host(hp, U1, constrained, cvalues, pin):
    """The same iteration through host buffers, as the reference's Newton driver moves its vectors:
    evaluation_point uploaded for every assembly, newton_update downloaded, update done on the host."""
    import numpy as np
    h2d = d2h = 0
    ev = pin("ev", U1.size)
    ev[:] = U1
    hp.set_vector("evaluation_point", ev)
    h2d += ev.nbytes
    hp.assemble(True)
    last = hp.rhs_norm()
    d2h += 8
    dx, info = hp.solve_linear_system(download=True, **LIN)
    d2h += dx.nbytes
    alpha = 1.0
    no = dx.size
    while alpha > 1e-3:
        np.multiply(dx, alpha, out=ev[:no])
        ev[:no] += U1[:no]
        ev[constrained] = cvalues[constrained]
        hp.set_vector("evaluation_point", ev)
        hp.update_ghosts("evaluation_point")      # ghost import (no-op on one rank)
        h2d += ev.nbytes
        hp.assemble(False)
        d2h += 8
        if hp.rhs_norm() < 0.9 * last:
            break
        alpha *= 0.5
    return info["iterations"], h2d, d2h


def run_ours(args):
    import numpy as np
    import torch
    from softx_2020_200_b200 import GLSHotPath
    from softx_2020_200_b200.mesh import BoxMesh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.cells
    t0 = time.perf_counter()
    gmesh = BoxMesh(3, n, 2, 2, bcs=CAVITY)
    n_global = gmesh.n_dofs
    mesh = gmesh if world == 1 else gmesh.partition(world, rank)
    hp = GLSHotPath(local_rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(GLSHotPath.comm_unique_id().copy())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        hp.comm_init(world, rank, uid.cpu().numpy())
    mesh.attach(hp)
    hp.set_physics(NU)
    t_setup = time.perf_counter() - t0
    if world > 1:
        gmesh.close()

    constrained = mesh.array("constrained").astype(bool)
    cvalues = mesh.array("constraint_values").copy()
    U0 = mesh.initial_state()
    # Newton iteration 0 from rest -> U1, the linearisation point of every timed step
    hp.set_vector("present_solution", U0)
    hp.set_vector("evaluation_point", U0)
    its0, _ = newton_step_device(hp, lambda: None)
    hp.accept_evaluation_point()
    U1 = hp.get_vector("present_solution")

    def reset():
        hp.set_vector("present_solution", U1)
        hp.set_vector("evaluation_point", U1)

    pins = {}

    def pin(name, size):
        if name not in pins:
            pins[name] = torch.empty(size, dtype=torch.float64, pin_memory=True)
        return pins[name].numpy()

    # ---- device-resident arm ----
    for _ in range(max(args.warmup - 1, 0)):
        reset()
        newton_step_device(hp, lambda: None)
    clocks = ClockSampler(local_rank)
    barrier()
    hp.reset_timers()
    clocks.start()
    wall0 = time.perf_counter()
    its = trials = 0
    for _ in range(args.steps):
        reset()
        its, trials = newton_step_device(hp, lambda: None)
    barrier()
    wall = time.perf_counter() - wall0
    clk = clocks.stop()
    tm = hp.timers()
    dev_ms = (tm["assemble_system_ms"] + tm["assemble_rhs_ms"] + tm["setup_ilu_ms"] +
              tm["solve_linear_system_ms"]) / args.steps
    dev_ms = max_over_ranks(dev_ms)
    value = n_global / (dev_ms * 1e-3) / 1e6

    # ---- end-to-end arm: host buffers in, host buffers out ----
    newton_step_host(hp, U1, constrained, cvalues, pin)              # warm-up
    barrier()
    e0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        _, h2d, d2h = newton_step_host(hp, U1, constrained, cvalues, pin)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - e0) / args.steps)
    e2e = n_global / e2e_s / 1e6

    # ---- roofline of the dominant kernel, timed live with CUDA events on the library stream ----
    hbm, hbm_src = peaks()
    N, nnz = mesh.n_owned, mesh.nnz
    per = {"spmv": tm["spmv_ms"] / max(tm["spmv_calls"], 1),
           "ilu_apply": tm["trsv_ms"] / max(tm["trsv_calls"], 1),
           "orthog": tm["orthog_ms"] / max(tm["orthog_calls"], 1)}
    algo = {"spmv": 12 * nnz + 24 * N, "ilu_apply": 12 * nnz + 40 * N}
    share = {k: tm[k2] / args.steps / dev_ms for k, k2 in
             (("spmv", "spmv_ms"), ("ilu_apply", "trsv_ms"), ("orthog", "orthog_ms"),
              ("assemble_system", "assemble_system_ms"), ("assemble_rhs", "assemble_rhs_ms"),
              ("setup_ilu", "setup_ilu_ms"))}
    dom = "ilu_apply" if tm["trsv_ms"] >= tm["spmv_ms"] else "spmv"
    achieved = algo[dom] / (per[dom] * 1e-3) / 1e9
    # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    # (profiles/), only when it was taken on this very workload
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            cap = json.load(f)
        key = "%s@n%d" % (dom, n)
        if world == 1 and key in cap:
            traffic = cap[key]["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm, "traffic": traffic, "peak_source": hbm_src,
                "algorithmic_bytes_per_launch": algo[dom], "avg_launch_ms": per[dom],
                "share_of_step": share[dom],
                "spmv": {"achieved": algo["spmv"] / (per["spmv"] * 1e-3) / 1e9,
                         "frac": algo["spmv"] / (per["spmv"] * 1e-3) / 1e9 / hbm,
                         "avg_launch_ms": per["spmv"], "share_of_step": share["spmv"]}}

    if rank != 0:
        return
    out = {
        "metric": METRIC, "value": value, "unit": "MDoF/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D lid-driven cavity Q2-Q2, Re=400 steady, one Newton iteration at "
                               "the state after the first Newton update from rest; GMRES(30)+ILU(0), "
                               "rel 1e-4 / abs 1e-9, ILU atol 1e-12",
                   "cells_per_dir": n, "n_dofs": n_global, "nnz": int(gmesh.nnz) if world == 1 else None,
                   "gmres_iterations": its, "line_search_trials": trials,
                   "parallelism": "1 rank per GPU, contiguous row blocks, block-Jacobi ILU per rank"
                   if world > 1 else "1 GPU",
                   "cache": "inputs larger than L2 (matrix+factors %.1f GB)" % (20 * nnz / 1e9)},
        "gpu_launches": int(tm["kernel_launches"]),
        "wall_ms_per_step": wall / args.steps * 1e3,
        "phases_ms_per_step": {k: tm[k] / args.steps for k in
                               ("assemble_system_ms", "assemble_rhs_ms", "setup_ilu_ms",
                                "solve_linear_system_ms", "spmv_ms", "trsv_ms", "orthog_ms")},
        "setup_s": t_setup, "first_newton_iteration_gmres_iterations": its0,
        "e2e": {"value": e2e, "unit": "MDoF/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3},
        "roofline": roofline, "clocks": clk,
    }
    # ---- CPU baseline on this box's host cores (bounded sample), rank 0, N = 1 only ----
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        st = cpu_newton_step_setup(args.cpu_cells, threads)
        c0 = time.perf_counter()
        _, cits = cpu_newton_step(st, st["U1"])
        cdt = time.perf_counter() - c0
        out["cpu_baseline"] = {
            "value": st["mesh"].ndof / cdt / 1e6, "unit": "MDoF/s", "cores": threads,
            "kind": "port",
            "sample": "same cavity at n=%d (%d DoFs), Newton iteration 1, %d GMRES iterations, %d "
                      "threads = %d block-Jacobi ILU blocks, %.1f s" %
                      (args.cpu_cells, st["mesh"].ndof, cits, threads, threads, cdt)}
    print(json.dumps(out))
    hp.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cells", type=int, default=64,
                    help="cells per direction (64 -> 8.59 M DoFs, the largest 1-GPU configuration)")
    ap.add_argument("--cpu-cells", type=int, default=CPU_SAMPLE_CELLS)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
