import sys, time, os, resource
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
from parallel_amg_b200 import _lib as L
n1 = int(sys.argv[1]); pp = tuple(int(x) for x in sys.argv[2].split(',')) if len(sys.argv) > 2 else (1,1,1)
P = int(np.prod(pp))
c = L.Context(P)
t=time.time(); c.gallery_poisson((n1,)*3, pp); print("gallery %.1fs"%(time.time()-t))
t=time.time(); c.setup(); print("setup %.1fs"%(time.time()-t), "maxrss %.1f GB"%(resource.getrusage(resource.RUSAGE_SELF).ru_maxrss/1e6))
nl = c.num_levels()
for l in range(nl):
    i = [c.level_info(l,p) for p in range(P)]
    print(l, "n", i[0].n_global, "nnzA", sum(x.nnz[0]+x.nnz[1] for x in i), "nnzP", sum(x.nnz[2]+x.nnz[3] for x in i), "ghost", [x.n_ghost for x in i][:4], "rho %.3f"%i[0].rho)
