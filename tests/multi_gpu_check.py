"""Run under torchrun on N >= 2 GPUs: one Newton iteration of the 3D Q2-Q2 cavity on the partitioned
mesh (NCCL halo exchange + all-reduce, block-Jacobi ILU per rank) against the CPU oracle run with
the same row blocks (the reference's behaviour on N MPI ranks: Ifpack ILU with overlap 0)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def check(n, world, rank, local, dist, verbose=True):
    """One Newton iteration of the 3D Q2-Q2 cavity at n cells per direction on `world` ranks (the
    process group `dist` is up, the CUDA device is set) against the oracle with the same row
    blocks.  Returns the parity record (identical on every rank)."""
    from oracle import reference_port as R
    from softx_2020_200_b200 import GLSHotPath
    from softx_2020_200_b200.mesh import BoxMesh
    bcs = [(0, "noslip"), (1, "noslip"), (2, "noslip"), (4, "noslip"), (5, "noslip"),
           (3, "function", (1.0, 0.0, 0.0))]
    g = BoxMesh(3, n, 2, 2, bcs=bcs)
    m = g.partition(world, rank)
    hp = GLSHotPath(local)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.from_numpy(GLSHotPath.comm_unique_id().copy())
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    hp.comm_init(world, rank, uid.cpu().numpy())
    m.attach(hp)
    hp.set_physics(0.005)
    l2g = m.array("local_to_global")
    # a non-trivial global state, same on every rank
    xyz = g.array("dof_coords").reshape(-1, 3)
    Ug = 0.05 * np.sin(np.pi * xyz[:, 0] + 0.3 * g.array("dof_component")) * np.cos(np.pi * xyz[:, 1]) \
        * np.cos(0.5 * np.pi * xyz[:, 2])
    con = g.array("constrained").astype(bool)
    Ug[con] = g.array("constraint_values")[con]
    hp.set_vector("present_solution", Ug[l2g])
    hp.set_vector("evaluation_point", Ug[l2g])
    hp.assemble(True)
    a_loc, b_loc = hp.get_matrix_values(), hp.get_vector("system_rhs")
    norm = hp.rhs_norm()
    dx_loc, info = hp.solve_linear_system(relative_residual=1e-6, minimum_residual=1e-12,
                                          max_iterations=2000, ilu_atol=1e-12)
    hp.line_search_point(1.0)
    ev = hp.get_vector("evaluation_point")
    hp.assemble(False)
    res1 = hp.rhs_norm()

    # gather owned pieces on rank 0
    def gather(x):
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([x.size], dtype=torch.int64, device="cuda"))
        sizes = [int(s.item()) for s in sizes]
        bufs = [torch.zeros(s, dtype=torch.float64, device="cuda") for s in sizes]
        dist.all_gather(bufs, torch.from_numpy(np.ascontiguousarray(x)).cuda())
        return [b.cpu().numpy() for b in bufs], sizes

    b_all, sizes = gather(b_loc)
    dx_all, _ = gather(dx_loc)
    ev_owned, _ = gather(ev[:m.n_owned])
    # ghost values of the updated evaluation point must equal their owners' values
    ev_glob = np.concatenate(ev_owned)
    assert np.array_equal(ev, ev_glob[l2g]), "halo exchange mismatch on rank %d" % rank
    ok, err_a, err_b, err_x, its_ref = True, 0.0, 0.0, 0.0, 0
    if rank == 0:
        R.lib().glso_set_num_threads(1)
        lid = lambda x: np.stack([np.ones(len(x)), 0 * x[:, 0], 0 * x[:, 0]], axis=1)
        obcs = {0: ("noslip",), 1: ("noslip",), 2: ("noslip",), 4: ("noslip",), 5: ("noslip",),
                3: ("function", lid)}
        prov = R.BoxMesh(3, n, 2, 2, bcs=obcs, renumber="none")
        key = lambda c, k: np.lexsort(tuple(np.round(c[:, d] * 1e6).astype(np.int64)
                                            for d in range(3)) + (k,))
        new_of_old = np.empty(prov.ndof, dtype=np.int64)
        new_of_old[key(prov.dof_coords, prov.dof_comp)] = key(
            g.array("dof_coords").reshape(-1, 3), g.array("dof_component"))
        om = R.BoxMesh(3, n, 2, 2, bcs=obcs, renumber=new_of_old)
        pr = R.scheme_params("steady", None, 0.005)
        a_ref, b_ref = R.assemble(om, Ug, pr, True)
        bp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        b_gpu, dx_gpu = np.concatenate(b_all), np.concatenate(dx_all)
        err_b = np.max(np.abs(b_gpu - b_ref)) / np.max(np.abs(b_ref))
        dx_ref, its_ref, _ = R.solve_linear_system(om, a_ref, b_ref, rel=1e-6, abs_=1e-12,
                                                   max_iters=2000, ilu_atol=1e-12, block_ptr=bp)
        err_x = np.linalg.norm(dx_gpu - dx_ref) / np.linalg.norm(dx_ref)
        # matrix rows of rank 0
        rows = slice(om.rowptr[0], om.rowptr[sizes[0]])
        col_g = l2g[m.array("col_idx")]
        rp = m.array("row_ptr")
        err_a = 0.0
        for i in range(sizes[0]):
            o = np.argsort(col_g[rp[i]:rp[i + 1]])
            ref = a_ref[om.rowptr[i]:om.rowptr[i + 1]]
            err_a = max(err_a, np.max(np.abs(a_loc[rp[i]:rp[i + 1]][o] - ref)) / max(np.max(np.abs(ref)), 1e-300))
        (print if verbose else (lambda *a: None))("multi_gpu_check world=%d n=%d: matrix err %.2e rhs err %.2e | GMRES %d (oracle, %d "
              "blocks: %d) update err %.2e | ||rhs|| %.6e (oracle %.6e) -> %.3e"
              % (world, n, err_a, err_b, info["iterations"], world, its_ref, err_x, norm,
                 np.linalg.norm(b_ref), res1))
        ok = err_a <= 1e-12 and err_b <= 1e-12 and abs(info["iterations"] - its_ref) <= 2 and \
            err_x <= 1e-5 and abs(norm - np.linalg.norm(b_ref)) <= 1e-12 * norm
    rec = torch.zeros(6, dtype=torch.float64, device="cuda")
    if rank == 0:
        rec = torch.tensor([1.0 if ok else 0.0, info["iterations"], its_ref, err_x, err_a, err_b],
                           dtype=torch.float64, device="cuda")
    dist.broadcast(rec, 0)
    hp.close()
    g.close()
    r = rec.cpu().numpy()
    return {"ok": bool(r[0]), "cells_per_dir": n, "n_ranks": world, "iterations": int(r[1]),
            "oracle_iterations": int(r[2]), "update_err": float(r[3]), "matrix_err": float(r[4]),
            "rhs_err": float(r[5]), "oracle": "block-Jacobi ILU with the ranks' row blocks"}


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    rec = check(n, world, rank, local, dist)
    dist.destroy_process_group()
    if not rec["ok"]:
        sys.exit(1)
    if rank == 0:
        print("MULTI_GPU_CHECK_OK", rec)


if __name__ == "__main__":
    main()
