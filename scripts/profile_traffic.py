"""Level-0/1 SpMV-family launches of several BASELINE workloads in ONE process, for one `ncu --set full` pass
(DRAM bytes per launch vs algorithmic bytes: is there anything for an x-vector staging to recover?).
  python scripts/profile_traffic.py [reps=2] [workloads=poisson3d-256,elasticity3d-96,diffusion-jump-3d-256]
Prints one line per timed kind with the algorithmic bytes of DESIGN.md section 4, in launch order."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from parallel_amg_b200 import _lib as L  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["poisson3d-256", "elasticity3d-96", "diffusion-jump-3d-256"]
L.set_num_threads(os.cpu_count() or 1)
for name in names:
    wl = bench.WORKLOADS[name]
    c = L.Context(1)
    bench.make_problem(c, wl, 1)
    c.setup(c.default_options(**wl.get("opts", {})))
    c.device_init()
    st = c.stats()
    for lvl in (0, 1):
        if lvl >= c.num_levels() - 1:
            continue
        info = c.level_info(lvl, 0)
        nr, nc = info.n_own, info.n_own_coarse
        a_b = 12 * info.nnz[0] + 4 * (nr + 1) + 16 * nr
        p_b = 12 * info.nnz[2] + 4 * (nr + 1) + 8 * nc + 16 * nr
        for kind, kn, nb in ((0, "spmv", a_b), (1, "jacobi", a_b + 16 * nr), (3, "prolong", p_b)):
            ms = c.time_kernel(kind, lvl, reps, True)
            print(json.dumps(dict(workload=name, level=lvl, kind=kn, rows=nr, nnz=info.nnz[0] if kind != 3 else info.nnz[2],
                                  algorithmic_bytes=nb, ms=[round(float(m), 4) for m in ms], format=int(st.format[lvl]),
                                  fill=round(st.sell_fill[lvl], 3))), flush=True)
    c.close()
