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
CPU_SAMPLE_CELLS = 24     # 470 596 DoFs: about 10-15 s of CPU work per Newton step on 16 cores


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


def cpu_newton_step(st, U, structured=False):
    """structured=False: the reference's literal q x j x i cell loop (gls_navier_stokes.cc:387-748),
    "Mode A" of BASELINE.md section 3; True: the structured block form, the best-CPU "Mode B"."""
    import numpy as np
    R, mesh, pr = st["R"], st["mesh"], st["pr"]
    val, rhs = R.assemble(mesh, U, pr, True, threads=st["threads"], structured=structured)
    last = float(np.linalg.norm(rhs))
    tol = max(LIN["relative_residual"] * last, LIN["minimum_residual"])
    lu, dp = R.ilu0(mesh, val, LIN["ilu_atol"], LIN["ilu_rtol"], st["bp"])
    dx, its, res, ok, _ = R.gmres(mesh, val, lu, dp, rhs, tol, LIN["max_iterations"],
                                  LIN["restart"], st["bp"])
    dx[mesh.constrained != 0] = 0.0
    alpha = 1.0
    while alpha > 1e-3:
        Un = mesh.apply_nonzero_constraints(U + alpha * dx)
        _, r2 = R.assemble(mesh, Un, pr, False, threads=st["threads"], structured=structured)
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
              "= %d block-Jacobi ILU blocks, the reference's literal cell loop (Mode A)"
              % (args.cpu_cells, ndof, its, threads, threads))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "MDoF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "3D lid-driven cavity Q2-Q2 Re=400 steady, one Newton iteration "
                               "(CPU restatement of the reference path; Trilinos unavailable)",
                   "cells_per_dir": args.cpu_cells, "n_dofs": ndof, "gmres_iterations": its,
                   "ilu_blocks": threads,
                   "same_config_as_gpu_arm": False,
                   "note": "bounded sample: the GPU arm's default mesh (n=64) would take hours "
                           "here; the GPU arm prints its own value at this size under same_config"},
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
    return info["iterations"], trials, info


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


class GpuCase:
    """The cavity at n cells per direction on this rank's GPU, set up through the C ABI, with the
    linearisation point of the timed Newton step (the state after the first update from rest)."""

    def __init__(self, n, world, rank, local_rank, dist, torch):
        import numpy as np
        from softx_2020_200_b200 import GLSHotPath
        from softx_2020_200_b200.mesh import BoxMesh
        self.n, self.world, self.dist, self.torch = n, world, dist, torch
        t0 = time.perf_counter()
        # N > 1: every rank builds its own part only (owned rows + ghost layer,
        # glsnsh_mesh_create_local): no rank ever holds the global CSR
        self.mesh = mesh = BoxMesh(3, n, 2, 2, bcs=CAVITY, local=(world, rank) if world > 1 else None)
        self.n_global, self.nnz_global = mesh.n_global, (int(mesh.nnz) if world == 1 else None)
        self.hp = hp = GLSHotPath(local_rank)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                uid = torch.from_numpy(GLSHotPath.comm_unique_id().copy())
            uid = uid.cuda()
            dist.broadcast(uid, 0)
            hp.comm_init(world, rank, uid.cpu().numpy())
        mesh.attach(hp)
        hp.set_physics(NU)
        self.setup_s = time.perf_counter() - t0
        self.constrained = mesh.array("constrained").astype(bool)
        self.cvalues = mesh.array("constraint_values").copy()
        U0 = mesh.initial_state()
        hp.set_vector("present_solution", U0)
        hp.set_vector("evaluation_point", U0)
        self.its0, _, _ = newton_step_device(hp, lambda: None)
        hp.accept_evaluation_point()
        self.U1 = hp.get_vector("present_solution")
        self.pins = {}

    def reset(self):
        self.hp.set_vector("present_solution", self.U1)
        self.hp.set_vector("evaluation_point", self.U1)

    def pin(self, name, size):
        if name not in self.pins:
            self.pins[name] = self.torch.empty(size, dtype=self.torch.float64, pin_memory=True)
        return self.pins[name].numpy()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_device(self, steps, warmup, clocks=None):
        """K Newton steps with the state resident in HBM, timed by the library's CUDA-event phase
        timers on its stream; max over ranks."""
        hp = self.hp
        for _ in range(warmup):
            self.reset()
            newton_step_device(hp, lambda: None)
        self.barrier()
        hp.reset_timers()
        if clocks:
            clocks.start()
        wall0 = time.perf_counter()
        its = trials = 0
        info = None
        for _ in range(steps):
            self.reset()
            its, trials, info = newton_step_device(hp, lambda: None)
        self.barrier()
        wall = time.perf_counter() - wall0
        tm = hp.timers()
        dev_ms = (tm["assemble_system_ms"] + tm["assemble_rhs_ms"] + tm["setup_ilu_ms"] +
                  tm["solve_linear_system_ms"]) / steps
        dev_ms = self.max_over_ranks(dev_ms)
        return dict(ms=dev_ms, value=self.n_global / (dev_ms * 1e-3) / 1e6, wall_ms=wall / steps * 1e3,
                    its=its, trials=trials, info=info, timers=tm)

    def time_e2e(self, steps):
        """The same steps through host buffers (pinned): H2D of every evaluation point, D2H of the
        update and of every residual norm, inside the timed region; max over ranks."""
        newton_step_host(self.hp, self.U1, self.constrained, self.cvalues, self.pin)   # warm-up
        self.barrier()
        e0 = time.perf_counter()
        h2d = d2h = 0
        for _ in range(steps):
            _, h2d, d2h = newton_step_host(self.hp, self.U1, self.constrained, self.cvalues, self.pin)
        self.barrier()
        s = self.max_over_ranks((time.perf_counter() - e0) / steps)
        return dict(value=self.n_global / s / 1e6, ms=s * 1e3, h2d=int(h2d), d2h=int(d2h))

    def close(self):
        self.hp.close()


def run_ours(args):
    import torch

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

    # ---- N > 1: parity of the partitioned path against the block-Jacobi oracle, before timing ----
    parity = None
    if world > 1 and not args.no_parity_check:
        from tests.multi_gpu_check import check as multi_gpu_check
        parity = multi_gpu_check(args.parity_cells, world, rank, local_rank, dist, verbose=False)
        if not parity["ok"]:
            raise SystemExit("multi-GPU parity check failed: %r" % (parity,))

    n = args.cells
    case = GpuCase(n, world, rank, local_rank, dist, torch)
    hp, mesh = case.hp, case.mesh
    clocks = ClockSampler(local_rank)
    dev = case.time_device(args.steps, max(args.warmup - 1, 0), clocks)
    clk = clocks.stop()
    tm, dev_ms, value, its, trials = dev["timers"], dev["ms"], dev["value"], dev["its"], dev["trials"]
    info = dev["info"]
    if not (info["true_residual"] <= info["tolerance"] * 1.01):
        raise SystemExit("GMRES' logged residual %.3e exceeds its tolerance %.3e: the step is invalid"
                         % (info["true_residual"], info["tolerance"]))
    e2e = case.time_e2e(args.steps)

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
    # DRAM bytes per launch from the committed ncu --set full captures (profiles/traffic.json),
    # only when they were taken on this very workload
    cap = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            cap = json.load(f)
    except (OSError, ValueError):
        pass

    def traffic_of(kernel):
        key = "%s@n%d" % (kernel, n)
        return cap[key]["dram_bytes_per_launch"] if world == 1 and key in cap else None

    def line(kernel):
        t = traffic_of(kernel)
        a = algo[kernel] / (per[kernel] * 1e-3) / 1e9
        return {"achieved": a, "frac": a / hbm, "avg_launch_ms": per[kernel],
                "share_of_step": share[kernel], "traffic": t,
                "dram_frac": (t / (per[kernel] * 1e-3) / 1e9 / hbm) if t else None}

    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm, "traffic": traffic_of(dom), "peak_source": hbm_src,
                "algorithmic_bytes_per_launch": algo[dom], "avg_launch_ms": per[dom],
                "share_of_step": share[dom], "spmv": line("spmv"), "ilu_apply": line("ilu_apply")}
    fp64 = cap.get("fp64_peak")
    if fp64 and world == 1:
        # assembly: algorithmic flops of the structured form (2.5e6 per Q2-Q2 hex
        # for matrix + rhs) against the fp64 FMA peak measured on this pool (tools/fp64_peak.cu)
        ms_asm = tm["assemble_system_ms"] / max(tm["assemble_system_calls"], 1)
        tf = 2.5e6 * mesh.n_cells / (ms_asm * 1e-3) / 1e12
        roofline["assembly"] = {"bound": "fp64", "achieved": tf, "peak": fp64["tflops"],
                                "unit": "TFLOP/s", "frac": tf / fp64["tflops"], "avg_launch_ms": ms_asm,
                                "peak_source": fp64.get("source")}

    # ---- the GPU arm once more at the CPU sample size: a same-configuration ratio ----
    same = None
    if world == 1 and not args.no_cpu_baseline and args.cpu_cells != n:
        small = GpuCase(args.cpu_cells, 1, 0, local_rank, None, torch)
        sd = small.time_device(3, 1)
        se = small.time_e2e(3)
        same = {"cells_per_dir": args.cpu_cells, "n_dofs": small.n_global,
                "gpu": {"value": sd["value"], "e2e": se["value"], "ms_per_step": sd["ms"],
                        "gmres_iterations": sd["its"], "ilu_blocks": 1}}
        small.close()

    if rank != 0:
        return
    out = {
        "metric": METRIC, "value": value, "unit": "MDoF/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "3D lid-driven cavity Q2-Q2, Re=400 steady, one Newton iteration at "
                               "the state after the first Newton update from rest; GMRES(30)+ILU(0), "
                               "rel 1e-4 / abs 1e-9, ILU atol 1e-12",
                   "cells_per_dir": n, "n_dofs": case.n_global, "nnz": case.nnz_global,
                   "n_dofs_per_gpu_rank0": int(mesh.n_owned), "nnz_per_gpu_rank0": int(mesh.nnz),
                   "gmres_iterations": its, "line_search_trials": trials,
                   "true_residual": info["true_residual"], "tolerance": info["tolerance"],
                   "ilu_blocks": world,
                   "parallelism": "1 rank per GPU, contiguous row blocks, block-Jacobi ILU per rank"
                   if world > 1 else "1 GPU",
                   "cache": "inputs larger than L2 (matrix+factors %.1f GB)" % (20 * nnz / 1e9)},
        "gpu_launches": int(tm["kernel_launches"]),
        "wall_ms_per_step": dev["wall_ms"],
        "phases_ms_per_step": {k: tm[k] / args.steps for k in
                               ("assemble_system_ms", "assemble_rhs_ms", "setup_ilu_ms",
                                "solve_linear_system_ms", "spmv_ms", "trsv_ms", "orthog_ms")},
        "setup_s": case.setup_s, "first_newton_iteration_gmres_iterations": case.its0,
        "e2e": {"value": e2e["value"], "unit": "MDoF/s", "h2d_bytes_per_step": e2e["h2d"],
                "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": e2e["ms"]},
        "roofline": roofline, "clocks": clk,
    }
    if parity is not None:
        out["parity_check"] = parity
    # ---- CPU baseline on this box's host cores (bounded sample), rank 0, N = 1 only ----
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        st = cpu_newton_step_setup(args.cpu_cells, threads)
        c0 = time.perf_counter()
        _, cits = cpu_newton_step(st, st["U1"])
        cdt = time.perf_counter() - c0
        ndof_c = st["mesh"].ndof
        out["cpu_baseline"] = {
            "value": ndof_c / cdt / 1e6, "unit": "MDoF/s", "cores": threads,
            "kind": "port",
            "sample": "same cavity at n=%d (%d DoFs), Newton iteration 1, %d GMRES iterations, %d "
                      "threads = %d block-Jacobi ILU blocks, %.1f s; the reference's literal cell loop "
                      "(Mode A)" % (args.cpu_cells, ndof_c, cits, threads, threads, cdt)}
        c0 = time.perf_counter()
        _, cits_b = cpu_newton_step(st, st["U1"], structured=True)
        cdt_b = time.perf_counter() - c0
        out["cpu_baseline_structured"] = {
            "value": ndof_c / cdt_b / 1e6, "unit": "MDoF/s", "cores": threads, "kind": "port",
            "sample": "as cpu_baseline with the structured (block-form) cell kernel instead of the "
                      "reference's literal loop (Mode B, best CPU), %d GMRES iterations, %.1f s"
                      % (cits_b, cdt_b)}
        if same is not None:
            same["cpu"] = {"value": ndof_c / cdt / 1e6, "gmres_iterations": cits,
                           "ilu_blocks": threads, "cores": threads}
            same["cpu_structured"] = {"value": ndof_c / cdt_b / 1e6}
            same["ratio_e2e"] = same["gpu"]["e2e"] / same["cpu"]["value"]
            same["ratio_e2e_vs_structured_cpu"] = same["gpu"]["e2e"] / same["cpu_structured"]["value"]
            same["note"] = ("both arms on the same mesh and tolerances; the CPU arm runs one "
                            "block-Jacobi ILU block per thread (what Ifpack with overlap 0 gives on "
                            "that many MPI ranks), the GPU one block: the iteration counts differ")
            out["same_config"] = same
    print(json.dumps(out))
    case.close()
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
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--parity-cells", type=int, default=8,
                    help="N > 1: mesh of the parity check against the oracle that precedes the timing")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
