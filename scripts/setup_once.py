"""One pamg_setup of Poisson n^3 on one part, wall times of gallery and setup: python scripts/setup_once.py [n=256]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parallel_amg_b200 import _lib as L  # noqa: E402

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
c = L.Context(1)
t0 = time.perf_counter()
c.gallery_poisson((n1, n1, n1), (1, 1, 1))
t1 = time.perf_counter()
c.setup()
t2 = time.perf_counter()
print(f"PAMG_HUGE_PAGES={os.environ.get('PAMG_HUGE_PAGES', '(default)')}: gallery {t1 - t0:.2f} s, setup {t2 - t1:.2f} s, levels {c.num_levels()}", flush=True)
