"""One process per GPU: wiring of the parts over torch.distributed (plumbing only).

Every rank builds the same hierarchy metadata (the host setup is deterministic), uploads its own
part, and exchanges the opaque peer-memory handle of its arena with all other ranks; after
`connect_parts` the halo exchange / all-reduces run inside the CUDA kernels over NVLink peer
mappings — torch.distributed is not on the data path."""
from __future__ import annotations


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
