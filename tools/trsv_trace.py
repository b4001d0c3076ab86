"""Where does the time of the triangular solves go?  Traces one ILU apply (publish time of every
row in both sweeps, tools: glsns_ilu_apply_trace) and walks the dependency DAG on the host:
for every group, the wait between its last-arriving input and its own publication, split by
whether that input came from the same warp (register window) or through L2.
    python tools/trsv_trace.py N"""
import json, sys
sys.path.insert(0, ".")
import numpy as np
from softx_2020_200_b200 import GLSHotPath
from softx_2020_200_b200.mesh import BoxMesh

n = int(sys.argv[1])
CAVITY = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"), (3, "function", (1, 0, 0))]
m = BoxMesh(3, n, 2, 2, bcs=CAVITY); hp = GLSHotPath(0); m.attach(hp); hp.set_physics(0.005)
U0 = m.initial_state(); hp.set_vector("evaluation_point", U0); hp.assemble(True); hp.setup_ilu(0, 1e-12, 1.0)
x = np.random.default_rng(5).standard_normal(m.n_dofs)
for _ in range(3):
    z, tl, tu, wl, wu = hp.ilu_apply_trace(x)
rp, col, N = m.array("row_ptr"), m.array("col_idx"), m.n_dofs
lens = np.diff(rp)
out = {"n": n}
for name, t, w, upper in (("lower", tl, wl, False), ("upper", tu, wu, True)):
    ok = w >= 0
    t = t.astype(np.int64); t0 = t[ok].min(); t = t - t0
    warp = w & 0xFFFFFF; chained = (w >> 30) & 1
    # group heads: first row of each group = rows whose predecessor row has a different publish... use patterns
    rows = np.nonzero(ok)[0]
    # per row: last-arriving dependency
    dt_same, dt_other, crit_is_same = [], [], []
    other_chained, other_head, other_dist = [], [], []
    pure = {1000: [], 3000: [], 10000: []}  # chain step when every other-warp input is older than .. ns
    step = max(1, len(rows) // 60000)
    for i in rows[::step]:
        c = col[rp[i]:rp[i + 1]]
        c = c[(c > i) & (c < N)] if upper else c[c < i]
        c = c[w[c] >= 0]
        # drop own group (rows published at the same instant by the same warp adjacent)
        c = c[np.abs(c - i) >= 4] if len(c) else c
        if not len(c):
            continue
        j = c[np.argmax(t[c])]
        d = t[i] - t[j]
        same = warp[j] == warp[i]
        (dt_same if same else dt_other).append(d)
        if not same:
            (other_chained if chained[i] else other_head).append(d)
            other_dist.append(abs(int(j) - int(i)))
        if same:
            oth = c[warp[c] != warp[i]]
            age = t[j] - (t[oth].max() if len(oth) else -10**9)   # how long before the same-warp input
            for k in pure:
                if age > k:
                    pure[k].append(d)
    ds, do = np.array(dt_same), np.array(dt_other)
    out[name] = {"span_us": float(t[ok].max() / 1e3), "rows_sampled": len(ds) + len(do),
                 "frac_last_input_same_warp": len(ds) / max(1, len(ds) + len(do)),
                 "frac_rows_chained": float(chained[ok].mean()),
                 "wait_after_last_input_same_warp_ns": [float(np.percentile(ds, p)) for p in (10, 50, 90)] if len(ds) else None,
                 "pure_chain_step_ns_p10_p50_p90_by_min_age_of_other_inputs":
                     {k: [float(np.percentile(v, p)) for p in (10, 50, 90)] + [len(v)] for k, v in pure.items() if len(v)},
                 "other_team_last_input:_chained_rows_(n,p50_ns)_vs_chain_heads_(n,p50_ns)":
                     [len(other_chained), float(np.median(other_chained)) if other_chained else None,
                      len(other_head), float(np.median(other_head)) if other_head else None],
                 "other_team_last_input_row_distance_p10_p50_p90":
                     [float(np.percentile(other_dist, p)) for p in (10, 50, 90)] if other_dist else None,
                 "wait_after_last_input_other_warp_ns": [float(np.percentile(do, p)) for p in (10, 50, 90)] if len(do) else None}
    # critical path by time: walk back from the last published row through last-arriving inputs
    i = int(np.argmax(np.where(ok, t, -1))); hops_same = hops_other = 0; time_same = time_other = 0
    while True:
        c = col[rp[i]:rp[i + 1]]
        c = c[(c > i) & (c < N)] if upper else c[c < i]
        c = c[w[c] >= 0]
        if not len(c):
            break
        j = int(c[np.argmax(t[c])])
        if warp[j] == warp[i]:
            hops_same += 1; time_same += t[i] - t[j]
        else:
            hops_other += 1; time_other += t[i] - t[j]
        i = j
    out[name]["critical_path"] = {"same_warp_hops": hops_same, "same_warp_us": time_same / 1e3,
                                  "other_warp_hops": hops_other, "other_warp_us": time_other / 1e3}
for name, pl in zip(("lower", "upper"), hp.last_trace_polls):
    nw = min(len(pl) // 8, 148 * 4)
    st = pl[:nw * 8].reshape(nw, 8).astype(np.float64)
    busy = st[st[:, 6] > 0]
    tot = busy[:, :5].sum(axis=0)
    out[name]["solver_cycles_per_group_[ring_wait,window,mailbox_wait,triangle_publish,release_refill]"] = [float(v) for v in tot / busy[:, 6].sum()]
    k = int(np.argmax(st[:, 6]))
    out[name]["busiest_solver_groups_and_cycles_per_group"] = [float(st[k, 6])] + [float(v / st[k, 6]) for v in st[k, :5]]
print(json.dumps(out, indent=1))
