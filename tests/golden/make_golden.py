"""Generates tests/golden/oracle_golden.json from oracle/amg_oracle.py.

The reference snapshot has no golden vectors, so these fixtures freeze the
oracle's own outputs: any later change to a tie-break, the partition rule or the smoother shows up
as a diff here.  Run:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import amg_oracle as O  # noqa: E402

CASES = [
    dict(dims=[200, 200], parts=[2, 2], opts={}),          # BASELINE.json configs[0]
    dict(dims=[200, 200], parts=[1, 1], opts={}),
    dict(dims=[32, 32, 32], parts=[2, 2, 2], opts={}),
    dict(dims=[32, 32, 32], parts=[2, 1, 1], opts={}),
    dict(dims=[48, 48, 48], parts=[1, 1, 1], opts={}),
    dict(dims=[24, 24, 24], parts=[2, 2, 1], opts={"smoother": "l1jacobi"}),
    dict(dims=[24, 24, 24], parts=[2, 2, 1], opts={"smoother": "chebyshev", "cheb_degree": 3}),
    # BASELINE.json configs[3] / configs[4] at fixture size
    dict(dims=[10, 9, 8], parts=[2, 2, 1], opts={"coarse_size": 60}, kind="elasticity"),
    dict(dims=[16, 16, 16], parts=[2, 2, 2], opts={"eps_strength": 0.0831}, kind="jump"),
]


def run_case(dims, parts, opts, kind="poisson"):
    dims, parts = tuple(dims), tuple(parts)
    P = int(np.prod(parts))
    opts = dict(opts)
    if kind == "elasticity":
        A, coords = O.elasticity_q1(dims)
        owner = np.repeat(O.uniform_partition(parts, dims), 3).astype(np.int32)
        opts.update(block_size=3, nullspace=O.rigid_body_modes(coords))
    elif kind == "jump":
        A = O.diffusion_fv(dims, O.jump_coefficient_k(dims, blocks=4, kmax=1.0e4, eps_z=1.0e-3))
        owner = O.uniform_partition(parts, dims)
    else:
        A = O.poisson_fd(dims)
        owner = O.uniform_partition(parts, dims)
    h = O.build(A, owner, P, opts)
    gh = h["global"]
    n = A.shape[0]
    # rhs: b = A*1 plus a deterministic rough component (b = A*1 alone can converge in one step)
    i = np.arange(n, dtype=np.uint64)
    rough = ((i * np.uint64(2654435761)) % np.uint64(2 ** 32)).astype(np.float64) / 2.0 ** 31 - 1.0
    b = A @ np.ones(n) + rough
    xs, it, hist = O.pcg(h, O.pvector_from_global(h["levels"][0], b))
    sha = hashlib.sha256()
    for lev in gh["levels"][:-1]:
        sha.update(np.ascontiguousarray(lev["agg"], dtype=np.int64).tobytes())
    return dict(
        level_sizes=[int(l["A"].shape[0]) for l in gh["levels"]],
        level_nnz=[int(l["A"].nnz) for l in gh["levels"]],
        agg_sha256=sha.hexdigest(),
        ghost_counts=[[int(len(d["ghost_to_global"])) for d in lev["parts"]] for lev in h["levels"]],
        iters=int(it),
        hist=[float(v) for v in hist],
    )


if __name__ == "__main__":
    out = dict(generator="tests/golden/make_golden.py", cases=[])
    for c in CASES:
        r = run_case(c["dims"], c["parts"], c["opts"], c.get("kind", "poisson"))
        r.update(c)
        out["cases"].append(r)
        print(c, "->", r["level_sizes"], r["iters"])
    with open(os.path.join(HERE, "oracle_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
