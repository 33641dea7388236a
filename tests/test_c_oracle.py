"""The C/OpenMP oracle (oracle/pamg_oracle.c, the CPU baseline that is timed in bench.py) against
the normative Python oracle: V-cycle <= 1e-12, identical PCG iteration counts."""
import numpy as np
import pytest

import amg_oracle as O
import c_oracle
from util import det_vector, oracle_problem, own_of, own_parts, rel_err


@pytest.mark.parametrize("dims,pp,oopts", [
    ((200, 200), (2, 2), {}),
    ((20, 20, 20), (2, 2, 2), {}),
    ((33, 31, 17), (3, 2, 1), {"nu_pre": 2, "nu_post": 2}),
    ((28, 28, 28), (1, 1, 1), {"smoother": "l1jacobi"}),
    ((7, 5), (1, 1), {}),
    ((20, 20, 20), (2, 2, 1), {"cycle": "w", "coarse_size": 40}),
    ((60, 60), (2, 2), {"cycle": "w", "coarse_size": 40, "nu_pre": 2, "nu_post": 1}),
])
def test_c_oracle_matches_python_oracle(dims, pp, oopts):
    A, owner, h = oracle_problem(dims, pp, tuple(sorted(oopts.items())))
    co = c_oracle.COracle.from_oracle_hierarchy(h)
    lev = h["levels"][0]
    n = A.shape[0]
    b = det_vector(n, 7)
    z_ref = own_of(lev, O.vcycle(h, O.pvector_from_global(lev, b)))
    assert rel_err(co.vcycle(own_parts(lev, b)), z_ref) <= 1e-12
    for rhs in (A @ np.ones(n) + det_vector(n, 81), b):
        xs, it_ref, hist_ref = O.pcg(h, O.pvector_from_global(lev, rhs))
        x, it, hist = co.pcg(own_parts(lev, rhs))
        assert it == it_ref
        assert np.allclose(hist, hist_ref, rtol=1e-8)
        assert rel_err(x, own_of(lev, xs)) <= 1e-10
    xs, it_ref, hist_ref = O.pcg(h, O.pvector_from_global(lev, b), flexible=True)      # flexible CG (Polak-Ribiere beta)
    x, it, hist = co.pcg(own_parts(lev, b), flexible=True)
    assert it == it_ref and np.allclose(hist, hist_ref, rtol=1e-8) and rel_err(x, own_of(lev, xs)) <= 1e-10
    x, it, hist = co.pcg(own_parts(lev, b), precond=False, maxiter=400)
    xs, it_ref, _ = O.pcg(h, O.pvector_from_global(lev, b), precond=False, maxiter=400)
    assert it == it_ref


@pytest.mark.parametrize("dims,pp,oopts", [
    ((40, 40), (2, 2), {}),
    ((14, 14, 14), (2, 2, 1), {}),
    ((30, 30), (2, 1), {"cycle": "w", "coarse_size": 40}),
])
def test_c_oracle_fgmres_matches_python_oracle(dims, pp, oopts):
    """orc_fgmres against amg_oracle.fgmres: identical inner iteration counts, residual estimates, solutions; with a restart short
    enough to be taken, without preconditioner (plain restarted GMRES), and truncated by maxiter."""
    A, owner, h = oracle_problem(dims, pp, tuple(sorted(oopts.items())))
    co = c_oracle.COracle.from_oracle_hierarchy(h)
    lev = h["levels"][0]
    n = A.shape[0]
    b = A @ det_vector(n, 5)
    for kw in ({}, {"restart": 4}, {"precond": False, "restart": 10, "maxiter": 300}, {"maxiter": 3}, {"maxiter": 0}):
        xs, it_ref, hist_ref = O.fgmres(h, O.pvector_from_global(lev, b), **kw)
        x, it, hist = co.fgmres(own_parts(lev, b), **kw)
        assert it == it_ref, kw
        assert np.allclose(hist, hist_ref, rtol=1e-7), kw
        assert rel_err(x, own_of(lev, xs)) <= 1e-9, kw
    xs, it, hist = O.fgmres(h, O.pvector_from_global(lev, b))
    x = np.zeros(n)
    for d, v in zip(lev["parts"], xs):
        x[d["own_to_global"]] = v[: len(d["own_to_global"])]
    assert np.linalg.norm(b - A @ x) <= 1.0001e-8 * np.linalg.norm(b)          # the estimate is the true residual norm
