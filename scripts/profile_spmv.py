"""Few launches of the level-0/1 SpMV-family kernels of the 256^3 Poisson hierarchy, for `ncu --set full`
(each kind is launched `reps` times with the L2 flushed before it):
  python scripts/profile_spmv.py [n=256] [reps=2]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallel_amg_b200 import _lib as L  # noqa: E402

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = L.Context(1)
c.gallery_poisson((n1, n1, n1), (1, 1, 1))
c.setup()
c.device_init()
for lvl in (0, 1):
    for kind in (0, 1, 2, 3, 6, 7) if lvl == 0 else (0, 1, 2, 3):
        ms = c.time_kernel(kind, lvl, reps, True)
        print(lvl, kind, [round(float(m), 4) for m in ms], flush=True)
