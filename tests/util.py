"""Shared helpers for the parity tests: problems, oracle hierarchies, vector plumbing."""
import functools

import numpy as np

import amg_oracle as O
from parallel_amg_b200 import _lib as L

SMOOTHERS = {"jacobi": L.SMOOTHER_JACOBI, "l1jacobi": L.SMOOTHER_L1JACOBI, "chebyshev": L.SMOOTHER_CHEBYSHEV}

# oracle option names -> pamg_options names
def product_options(ctx, oopts, **extra):
    kw = {}
    for k, v in (oopts or {}).items():
        if k == "smoother":
            kw[k] = SMOOTHERS[v]
        elif k == "cycle":
            kw[k] = {"v": L.CYCLE_V, "w": L.CYCLE_W}[v]
        else:
            kw[k] = v
    kw.update(extra)
    return ctx.default_options(**kw)


@functools.lru_cache(maxsize=16)
def oracle_problem(dims, pp, opts_items=()):
    """(A, owner, h) with h = oracle hierarchy (global + localised)."""
    A = O.poisson_fd(dims)
    owner = O.uniform_partition(pp, dims)
    h = O.build(A, owner, int(np.prod(pp)), dict(opts_items))
    return A, owner, h


def det_vector(n, seed=1):
    """deterministic pseudo-random vector in (-1, 1) (same bits on every platform)."""
    i = np.arange(n, dtype=np.uint64)
    x = (i * np.uint64(2654435761) + np.uint64(seed) * np.uint64(40503)) % np.uint64(2 ** 32)
    return x.astype(np.float64) / 2.0 ** 31 - 1.0


def own_parts(level, v_global):
    return [v_global[d["own_to_global"]].copy() for d in level["parts"]]


def own_of(level, vs):
    return [v[: len(d["own_to_global"])].copy() for d, v in zip(level["parts"], vs)]


def product_context_from_oracle(h, oopts=None, **extra):
    """Upload the ORACLE-built hierarchy through pamg_level_upload so that device results can be
    compared against the oracle on bit-identical operators."""
    c = L.Context(h["nparts"])
    opts = product_options(c, oopts, **extra)
    levels = [lev["parts"] for lev in h["levels"]]
    c.upload_hierarchy(levels, h["coarse_inv"], opts, rho=h["rho_dinv_a"])
    return c


def rel_err(a, b):
    a = np.concatenate([np.ravel(x) for x in a])
    b = np.concatenate([np.ravel(x) for x in b])
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
