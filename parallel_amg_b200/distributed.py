"""One process per GPU: wiring of the parts over torch.distributed (plumbing only).

`shared_setup`: rank 0 runs the host setup once with every host core and hands the hierarchy to the other
ranks through a shared-memory file; each of them loads only its own part in full (pamg_hierarchy_save /
pamg_hierarchy_load).  `connect_parts`: every rank uploads its part and exchanges the opaque peer-memory
handle of its arena with all other ranks; afterwards the halo exchange / all-reduces run inside the CUDA
kernels over NVLink peer mappings — torch.distributed is not on the data path."""
from __future__ import annotations

import os
import shutil


def _scratch_dir(need_bytes: int) -> str:
    for cand in ("/dev/shm", "/tmp"):
        try:
            if shutil.disk_usage(cand).free > need_bytes:
                return cand
        except OSError:
            pass
    return "/tmp"


def shared_setup(ctx, build, rank: int, world: int, tag: str = "h", need_bytes: int = 1 << 30):
    """ctx: parallel_amg_b200._lib.Context(world).  build(ctx) sets the matrix and runs ctx.setup(...) — it is
    called on rank 0 only.  Returns (n_global, nnz_global).  After the call every rank holds the hierarchy: rank 0
    everything (it also keeps the global matrix), rank r > 0 its own part plus metadata."""
    import torch.distributed as dist
    if ctx.nparts != world:
        raise ValueError(f"nparts={ctx.nparts} but world size={world}: one part per rank")
    path = os.path.join(_scratch_dir(need_bytes), f"pamg_hier_{os.environ.get('MASTER_PORT', '0')}_{tag}_{world}.bin")
    meta = [None]
    if rank == 0:
        build(ctx)
        ctx.hierarchy_save(path)
        meta = [ctx.global_size()]
    dist.broadcast_object_list(meta, src=0)
    if rank != 0:
        ctx.hierarchy_load(path, keep_part=rank)
    dist.barrier()
    if rank == 0:
        os.remove(path)
    return meta[0]


def connect_parts(ctx, rank: int, world: int, local_rank: int):
    """ctx: parallel_amg_b200._lib.Context with the hierarchy set up; nparts must equal world."""
    import torch.distributed as dist
    if ctx.nparts != world:
        raise ValueError(f"nparts={ctx.nparts} but world size={world}: one part per rank")
    ctx.device_init([rank], [local_rank])
    if world == 1:
        return
    blobs = [None] * world
    dist.all_gather_object(blobs, ctx.comm_export(rank))
    for p, blob in enumerate(blobs):
        if p != rank:
            ctx.comm_import(p, blob)
    ctx.comm_connect()
    dist.barrier()
