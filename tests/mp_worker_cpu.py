"""Worker for tests/test_multiprocess.py (CPU, gloo): every rank holds ONE part and the halo data
really crosses processes.  It re-enacts the N>1 data path on the host from the PRODUCT's own
split blocks and halo plans (C ABI queries): consistent! = gloo isend/irecv along the plan,
mul! = A_oo x_own + A_og x_ghost, dot = all_reduce of own-value partials."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from parallel_amg_b200 import _lib as L
    from util import det_vector
    import amg_oracle as O
    dims = (18, 16, 12)
    pp = {2: (2, 1, 1), 4: (2, 2, 1)}[world]
    c = L.Context(world)
    c.gallery_poisson(dims, pp)
    c.setup()
    # 1. the replicated host setup is deterministic: every rank holds the same hierarchy
    import hashlib
    sha = hashlib.sha256()
    for l in range(c.num_levels()):
        for p in range(world):
            for a in c.index_maps(l, p):
                sha.update(a.tobytes())
            ip, ix, d = c.block(l, p, L.A_OO)
            sha.update(ip.tobytes() + ix.tobytes() + d.tobytes())
    digs = [None] * world
    dist.all_gather_object(digs, sha.hexdigest())
    assert len(set(digs)) == 1, "hierarchies differ between ranks"

    # 1b. rank 0 builds / the others load (the hand-off bench.py uses for N > 1): a rank that only LOADED the
    #     hierarchy holds exactly the same own part and the same metadata for the parts of the other ranks
    from parallel_amg_b200.distributed import shared_setup

    def build(ctx):
        ctx.gallery_poisson(dims, pp)
        ctx.setup()

    cl = L.Context(world)
    n_nnz = shared_setup(cl, build, rank, world, tag="mpcpu")
    assert tuple(n_nnz) == tuple(c.global_size())
    assert cl.num_levels() == c.num_levels()
    for l in range(c.num_levels()):
        for b in range(6):
            if l == c.num_levels() - 1 and b >= L.P_OO:
                continue
            assert all(np.array_equal(u, v) for u, v in zip(cl.block(l, rank, b), c.block(l, rank, b)))
        assert all(np.array_equal(u, v) for u, v in zip(cl.index_maps(l, rank), c.index_maps(l, rank)))
        pa, pb = cl.halo_plan(l, rank), c.halo_plan(l, rank)
        assert all(np.array_equal(pa[k], pb[k]) for k in pa)
        for p in range(world):
            ia, ib = cl.level_info(l, p), c.level_info(l, p)
            assert (ia.n_own, ia.n_ghost, list(ia.nnz), ia.n_send) == (ib.n_own, ib.n_ghost, list(ib.nnz), ib.n_send)
    cl.close()

    # 2. distributed mul!/dot on every level with ONLY this rank's part + its halo plan
    A = O.poisson_fd(dims)
    owner = O.uniform_partition(pp, dims)
    h = O.build(A, owner, world)
    for l in range(c.num_levels()):
        info = c.level_info(l, rank)
        own, gh, gho = c.index_maps(l, rank)
        plan = c.halo_plan(l, rank)
        n_glob = info.n_global
        xg = det_vector(n_glob, 3 + l)
        x_loc = np.concatenate([xg[own], np.full(len(gh), np.nan)])
        reqs, recv_bufs = [], []
        off = 0
        for q, cnt in zip(plan["send_part"], plan["send_count"]):
            buf = torch.from_numpy(x_loc[plan["send_idx"][off:off + cnt]].copy())
            reqs.append(dist.isend(buf, int(q), tag=l))
            off += cnt
        for q, s0, cnt in zip(plan["recv_part"], plan["recv_slot0"], plan["recv_count"]):
            buf = torch.empty(int(cnt), dtype=torch.float64)
            reqs.append(dist.irecv(buf, int(q), tag=l))
            recv_bufs.append((int(s0), int(cnt), buf))
        for r in reqs:
            r.wait()
        for s0, cnt, buf in recv_bufs:
            x_loc[len(own) + s0:len(own) + s0 + cnt] = buf.numpy()
        assert np.array_equal(x_loc[len(own):], xg[gh]), "consistent! delivered wrong ghost values"
        blocks = []
        for which, ncol in ((L.A_OO, len(own)), (L.A_OG, len(gh))):
            ip, ix, d = c.block(l, rank, which)
            blocks.append(sp.csr_matrix((d, ix, ip), shape=(len(own), ncol)))
        y = blocks[0] @ x_loc[:len(own)]
        if len(gh):
            y = y + blocks[1] @ x_loc[len(own):]
        Ag = h["global"]["levels"][l]["A"]
        ref = (Ag @ xg)[own]
        assert np.abs(y - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), f"level {l}: distributed mul! mismatch"
        part = torch.tensor([float(np.dot(y, xg[own]))], dtype=torch.float64)
        dist.all_reduce(part)
        assert abs(part.item() - float(xg @ (Ag @ xg))) <= 1e-10 * abs(part.item())
    dist.barrier()
    if rank == 0:
        print("MP_CPU_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
