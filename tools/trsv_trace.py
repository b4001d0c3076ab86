"""Where does the time of the triangular solves go?  Traces one ILU apply (tools:
glsns_ilu_apply_trace: when every row was published, when its totals reached the team's mailbox)
and walks the dependency DAG on the host, block by block (a block = the rows one solver step
publishes together): the wait between a block's last-arriving input and its publication, split
into the helper part (input published -> totals posted) and the solver part (posted -> published),
by whether that input came from the same team (shared-memory window) or through L2.
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
if len(sys.argv) > 2:  # raw trace for offline analysis: ns offsets from the first publication of each sweep
    def rel(t, ref):
        t = t.astype(np.int64); return np.where(t > 0, t - ref, -1).astype(np.int32)
    r0l, r0u = int(tl[wl >= 0].min()), int(tu[wu >= 0].min())
    np.savez_compressed(sys.argv[2], tl=rel(tl, r0l), tu=rel(tu, r0u), pl=rel(hp.last_trace_posts[0], r0l),
                        pu=rel(hp.last_trace_posts[1], r0u), wl=wl, wu=wu,
                        bl=rel(hp.last_trace_begin[0], r0l), bu=rel(hp.last_trace_begin[1], r0u),
                        il=rel(hp.last_trace_inputs[0], r0l), iu=rel(hp.last_trace_inputs[1], r0u))
pct = lambda v: [float(np.percentile(v, p)) for p in (10, 50, 90)] + [len(v)] if len(v) else None
out = {"n": n}
for name, t, tp, w, upper in (("lower", tl, hp.last_trace_posts[0], wl, False),
                              ("upper", tu, hp.last_trace_posts[1], wu, True)):
    ok = w >= 0
    t = t.astype(np.int64); t0 = t[ok].min(); t = t - t0; tp = tp.astype(np.int64) - t0
    team = w & 0xFFFFFF
    # blocks: runs of adjacent rows of one team published within 100 ns of each other
    rows = np.nonzero(ok)[0]
    new = np.ones(len(rows), dtype=bool)
    new[1:] = (np.diff(rows) != 1) | (team[rows[1:]] != team[rows[:-1]]) | (np.abs(np.diff(t[rows])) > 100)
    bid = np.cumsum(new) - 1
    blk = -np.ones(N, dtype=np.int64); blk[rows] = bid
    nb = int(bid[-1]) + 1
    bstart = rows[new]; bend = np.append(rows[np.nonzero(new)[0][1:] - 1], rows[-1]) + 1
    info = {}

    def block_inputs(b):
        if b in info:
            return info[b]
        r0, r1 = bstart[b], bend[b]
        c = col[rp[r0]:rp[r1]]
        c = c[(c >= r1) & (c < N)] if upper else c[c < r0]
        c = c[w[c] >= 0]
        res = None
        if len(c):
            same = team[c] == team[r0]
            cs, co = c[same], c[~same]
            js = int(cs[np.argmax(t[cs])]) if len(cs) else -1
            jo = int(co[np.argmax(t[co])]) if len(co) else -1
            res = (js, jo, int(t[r0:r1].max()), int(tp[r0:r1].max()))
        info[b] = res
        return res

    A = {k: [] for k in ("chain_total", "chain_after_post", "cross_total", "cross_helper", "cross_solver",
                         "cross_input_age_when_chain_last", "rows_per_block")}
    for b in range(0, nb, max(1, nb // 40000)):
        A["rows_per_block"].append(bend[b] - bstart[b])
        r = block_inputs(b)
        if r is None:
            continue
        js, jo, tpub, tpost = r
        ts_, to_ = (t[js] if js >= 0 else -10**12), (t[jo] if jo >= 0 else -10**12)
        if ts_ >= to_:
            A["chain_total"].append(tpub - ts_); A["chain_after_post"].append(tpub - tpost)
            if jo >= 0:
                A["cross_input_age_when_chain_last"].append(ts_ - to_)
        else:
            A["cross_total"].append(tpub - to_); A["cross_helper"].append(tpost - to_); A["cross_solver"].append(tpub - tpost)
    o = {"span_us": float(t[ok].max() / 1e3), "blocks": nb, "mean_rows_per_block": float(np.mean(A["rows_per_block"]))}
    for k in A:
        if k != "rows_per_block":
            o[k + "_ns_p10_p50_p90_n"] = pct(np.array(A[k]))
    # critical path by time: walk back from the last published block through last-arriving inputs
    b = int(blk[int(np.argmax(np.where(ok, t, -1)))])
    hs = ho = 0; ts = to = th = tsol = 0
    while True:
        r = block_inputs(b)
        if r is None:
            break
        js, jo, tpub, tpost = r
        if jo < 0 or (js >= 0 and t[js] >= t[jo]):
            hs += 1; ts += tpub - t[js]; b = int(blk[js])
        else:
            ho += 1; to += tpub - t[jo]; th += tpost - t[jo]; tsol += tpub - tpost; b = int(blk[jo])
    o["critical_path"] = {"chain_hops": hs, "chain_us": ts / 1e3, "cross_hops": ho, "cross_us": to / 1e3,
                          "cross_helper_part_us": th / 1e3, "cross_solver_part_us": tsol / 1e3}
    out[name] = o
for name, pl in zip(("lower", "upper"), hp.last_trace_polls):
    nw = min(len(pl) // 20, int(__import__('os').environ.get('GLSNS_TRACE_TEAMS', 296)))
    st = pl[:nw * 8].reshape(nw, 8).astype(np.float64)
    busy = st[st[:, 6] > 0]
    tot = busy[:, :5].sum(axis=0)
    out[name]["solver_cycles_per_block_[ring_wait,window,mailbox_wait,totals_publish,release_refill]"] = [float(v) for v in tot / busy[:, 6].sum()]
    hc = pl[nw * 8:nw * 12].reshape(nw, 4).astype(np.float64).sum(axis=0)
    out[name]["helpers_[items,items_that_waited,poll_rounds,entries_re_read]"] = [float(v) for v in hc]
    hs = pl[nw * 12:nw * 20].reshape(nw, 8).astype(np.float64).sum(axis=0)
    out[name]["helper_cycle_shares_[indices,values,solution_entries,mailbox_full,other]"] = \
        [float(v / max(hs[4], 1)) for v in hs[:4]] + [float(1 - hs[:4].sum() / max(hs[4], 1))]
    out[name]["helper_cycles_per_item"] = float(hs[4] / max(hc[0], 1))
    k = int(np.argmax(st[:, 6]))
    out[name]["busiest_solver_blocks_and_cycles_per_block"] = [float(st[k, 6])] + [float(v / st[k, 6]) for v in st[k, :5]]
print(json.dumps(out, indent=1))
