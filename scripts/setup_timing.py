"""Wall time of pamg_setup per phase (PAMG_SETUP_TIMING=1 goes to stderr): python scripts/setup_timing.py [n=256] [parts=1]
Runs the default path (device chain incl. aggregation when a GPU is present) and PAMG_GPU_AGG=0 (host walk) in one process."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PAMG_SETUP_TIMING"] = "1"
from parallel_amg_b200 import _lib as L  # noqa: E402

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 1
PP = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[parts]
for label, env in (("device aggregation", {}), ("host aggregation", {"PAMG_GPU_AGG": "0"}), ("device aggregation (2nd run)", {})):
    for k in ("PAMG_GPU_AGG",):
        os.environ.pop(k, None)
    os.environ.update(env)
    c = L.Context(parts)
    t0 = time.perf_counter()
    c.gallery_poisson((n1, n1, n1), PP)
    t1 = time.perf_counter()
    c.setup()
    t2 = time.perf_counter()
    print(f"[{label}] gallery {t1 - t0:.2f} s, setup {t2 - t1:.2f} s, levels {c.num_levels()}", file=sys.stderr, flush=True)
    c.close()
