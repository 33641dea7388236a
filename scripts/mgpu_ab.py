"""torchrun worker: A/B of the engine's environment switches on N GPUs with ONE host setup.
  python -m torch.distributed.run --nproc-per-node N scripts/mgpu_ab.py <tag> [workload] -- "ENV1=a ENV2=b" "ENV1=c" ...
For every configuration: re-upload (device_init reads the switches), 3 warm-up + 5 timed resident solves (device time,
max over ranks), iteration count, and rank 0's per-kernel trace in gpurun_out/<tag>_trace_<k>.txt."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from parallel_amg_b200 import _lib as L  # noqa: E402
from parallel_amg_b200.distributed import connect_parts, shared_setup  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tag = sys.argv[1]
    workload = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "--" else "poisson3d-256"
    configs = sys.argv[sys.argv.index("--") + 1:] if "--" in sys.argv else [""]
    wl = bench.WORKLOADS[workload]
    L.set_num_threads(os.cpu_count() or 1)
    c = L.Context(world)

    def build(ctx):
        bench.make_problem(ctx, wl, world)
        ctx.setup(ctx.default_options(**wl.get("opts", {})))

    n, nnz = shared_setup(c, build, rank, world, tag=tag, need_bytes=400 * int(np.prod(wl["dims"])))
    tb = torch.from_numpy(c.host_matvec_global(bench.xstar(n))).cuda() if rank == 0 else torch.empty(n, dtype=torch.float64, device="cuda")
    dist.broadcast(tb, src=0)
    b = tb.cpu().numpy()
    del tb
    rows = []
    seen = set()
    for k, cfg in enumerate(configs):
        for name in seen:
            os.environ.pop(name, None)
        for kv in cfg.split():
            name, val = kv.split("=", 1)
            os.environ[name] = val
            seen.add(name)
        connect_parts(c, rank, world, local)
        own = c.index_maps(0, rank)[0]
        b_parts = [b[own] if p == rank else None for p in range(world)]
        c.load_rhs(b_parts)
        for _ in range(3):
            it, hist, ok = c.pcg_resident(bench.RTOL, bench.MAXITER, True)
        dist.barrier()
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(5):
            it, hist, ok = c.pcg_resident(bench.RTOL, bench.MAXITER, True)
            ms += c.stats().solve_ms
        t = torch.tensor([ms / 5], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vc = float(np.mean(c.time_kernel(5, 0, 13, False)[3:]))
        if rank == 0:
            bench.write_trace(c, rank, os.path.join(ROOT, "gpurun_out", f"{tag}_trace_{k}.txt"))
        else:
            c.pcg_resident(bench.RTOL, bench.MAXITER, True)   # the traced solve is collective
        dist.barrier()
        if rank == 0:
            row = dict(k=k, config=cfg, solve_ms=float(t.item()), iters=int(it), us_per_iter=1e3 * float(t.item()) / max(it, 1), vcycle_ms=vc, ok=bool(ok))
            rows.append(row)
            print(json.dumps(row), flush=True)
    if rank == 0:
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"{tag}_ab.json"), "w"), indent=1)
    dist.barrier()
    c.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
