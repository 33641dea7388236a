"""2 (or N) parts on ONE GPU in split mode (pack / main / boundary as separate launches): lets ncu time
the halo roles at the real per-part sizes, which a multi-rank run cannot be profiled for.
  python scripts/split_probe.py [n=128] [nparts=2] [solves=2]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallel_amg_b200 import _lib as L  # noqa: E402

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nparts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
solves = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pp = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[nparts]
c = L.Context(nparts)
c.gallery_poisson((n1, n1, n1), pp)
c.setup(c.default_options(use_graph=0))
n, nnz = c.global_size()
b = c.host_matvec_global(np.random.default_rng(1).uniform(-1, 1, n))
c.device_init(list(range(nparts)), [0] * nparts)
bs = [b[c.index_maps(0, p)[0]] for p in range(nparts)]
c.load_rhs(bs)
for _ in range(solves):
    it, hist, ok = c.pcg_resident(1e-8, 200, True)
st = c.stats()
print("iters", it, "solve_ms", st.solve_ms, "launches", st.kernel_launches, "fused", st.fused_halo)
