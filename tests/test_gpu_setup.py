"""the sparse products of the AMG setup (A_F*P0, A*P, R*(A*P)) on the GPU (csrc/setup_gpu.cu).
They keep the host product's accumulation order, so the hierarchy must be (i) bit-exact in structure and within 1e-12 in
values against the ORACLE, like the host setup, and (ii) bit-IDENTICAL to the host setup in every value."""
import time

import numpy as np
import pytest

import amg_oracle as O
from parallel_amg_b200 import _lib as L
from test_setup_parity import check_structure
from util import oracle_problem, product_options

pytestmark = pytest.mark.gpu


def _hierarchy_arrays(c):
    out = []
    for l in range(c.num_levels()):
        for p in range(c.nparts):
            own, gh, gho = c.index_maps(l, p)
            out += [own, gh, gho]
            for b in range(6):
                if l == c.num_levels() - 1 and b >= L.P_OO:
                    continue
                out += list(c.block(l, p, b))
    out.append(c.coarse_inverse())
    return out


@pytest.mark.parametrize("dims,pp,oopts", [((200, 200), (2, 2), {}), ((33, 31, 17), (3, 2, 1), {}), ((28, 28, 28), (1, 1, 1), {}),
                                          ((24, 24, 24), (2, 2, 1), {"eps_strength": 0.0831}), ((64, 64), (4, 1), {"coarse_size": 50})])
@pytest.mark.parametrize("budget", [None, "70000"], ids=["one-chunk", "many-chunks"])
def test_gpu_products_match_oracle_and_host_bitwise(dims, pp, oopts, budget, monkeypatch):
    A, owner, h = oracle_problem(dims, pp, tuple(sorted(oopts.items())))
    nparts = h["nparts"]
    monkeypatch.setenv("PAMG_GPU_SETUP", "1")
    if budget:
        monkeypatch.setenv("PAMG_GPU_SETUP_BUDGET", budget)   # products per chunk: forces the row-chunk loop
    cg = L.Context(nparts)
    cg.gallery_poisson(dims, pp)
    cg.setup(product_options(cg, oopts))
    check_structure(cg, h)
    monkeypatch.setenv("PAMG_GPU_SETUP", "0")
    ch = L.Context(nparts)
    ch.gallery_poisson(dims, pp)
    ch.setup(product_options(ch, oopts))
    for a, b in zip(_hierarchy_arrays(cg), _hierarchy_arrays(ch)):
        assert np.array_equal(a, b)


def test_gpu_products_elasticity_bitwise(monkeypatch):
    dims, pp = (10, 9, 8), (2, 2, 1)
    A, coords = O.elasticity_q1(dims)
    h = O.build(A, np.repeat(O.uniform_partition(pp, dims), 3).astype(np.int32), 4,
                dict(block_size=3, nullspace=O.rigid_body_modes(coords), coarse_size=60))
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("PAMG_GPU_SETUP", flag)
        c = L.Context(4)
        c.gallery_elasticity(dims, pp)
        c.setup(c.default_options(coarse_size=60))
        check_structure(c, h)
        res.append(_hierarchy_arrays(c))
    for a, b in zip(*res):
        assert np.array_equal(a, b)


def test_gpu_products_at_config_size_and_timing(monkeypatch, capsys):
    """128^3 (BASELINE config 2): same hierarchy bit for bit from both paths; the wall times go to the test log."""
    dims, pp = (128, 128, 128), (1, 1, 1)
    res, secs = [], []
    for flag in ("1", "0"):
        monkeypatch.setenv("PAMG_GPU_SETUP", flag)
        c = L.Context(1)
        c.gallery_poisson(dims, pp)
        t0 = time.perf_counter()
        c.setup()
        secs.append(time.perf_counter() - t0)
        info = [c.level_info(l, 0) for l in range(c.num_levels())]
        res.append([np.concatenate([np.ravel(x) for x in c.block(l, 0, b)]) for l in range(c.num_levels() - 1) for b in (0, 2, 4)])
        assert info[0].n_own == 128 ** 3
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    with capsys.disabled():
        print(f"\n[setup 128^3] GPU products {secs[0]:.2f} s, host products {secs[1]:.2f} s")
