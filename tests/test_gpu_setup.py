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


# ---- the greedy aggregation on the device (setup_gpu.cu gpu_aggregate) ---------------------------------------------
def _setup(kind, dims, pp, opts, monkeypatch, agg_mode):
    monkeypatch.setenv("PAMG_GPU_SETUP", "1")
    monkeypatch.setenv("PAMG_GPU_AGG", agg_mode)          # "2": the device or an error (no silent fallback); "0": the host walk
    c = L.Context(int(np.prod(pp)))
    if kind == "poisson":
        c.gallery_poisson(dims, pp)
    elif kind == "jump":
        c.gallery_diffusion_jump(dims, pp, blocks=4, kmax=1.0e4, eps_z=1.0e-3)
    else:
        c.gallery_elasticity(dims, pp)
    c.setup(c.default_options(**opts))
    return c


AGG_CASES = [("poisson", (200, 200), (2, 2), {}), ("poisson", (33, 31, 17), (3, 2, 1), {}), ("poisson", (48, 48, 48), (1, 1, 1), {}),
             ("poisson", (5000,), (1,), {}),                                   # 1-D: the longest dependency chain per vertex
             ("jump", (32, 32, 32), (2, 2, 2), dict(eps_strength=0.08)), ("jump", (40, 36, 28), (1, 2, 1), dict(eps_strength=0.25)),
             ("elasticity", (10, 9, 8), (2, 2, 1), dict(coarse_size=60))]


@pytest.mark.parametrize("kind,dims,pp,opts", AGG_CASES, ids=[f"{k}-{'x'.join(map(str, d))}-p{int(np.prod(p))}" for k, d, p, o in AGG_CASES])
def test_gpu_aggregation_is_the_host_walk_bit_for_bit(kind, dims, pp, opts, monkeypatch):
    """Aggregates of every (level, part) and the whole hierarchy behind them: device fixed point == host walk."""
    cg = _setup(kind, dims, pp, opts, monkeypatch, "2")
    ch = _setup(kind, dims, pp, opts, monkeypatch, "0")
    assert cg.num_levels() == ch.num_levels() and cg.num_levels() >= 2
    for l in range(cg.num_levels() - 1):
        for p in range(cg.nparts):
            assert np.array_equal(cg.aggregates(l, p), ch.aggregates(l, p)), (l, p)
    for a, b in zip(_hierarchy_arrays(cg), _hierarchy_arrays(ch)):
        assert np.array_equal(a, b)


def test_gpu_aggregation_matches_oracle(monkeypatch):
    """... and the oracle's walk directly (aggregates are part of check_structure)."""
    monkeypatch.setenv("PAMG_GPU_SETUP", "1")
    monkeypatch.setenv("PAMG_GPU_AGG", "2")
    for dims, pp, oopts in (((33, 31, 17), (3, 2, 1), {}), ((24, 24, 24), (2, 2, 1), {"eps_strength": 0.0831})):
        A, owner, h = oracle_problem(dims, pp, tuple(sorted(oopts.items())))
        c = L.Context(h["nparts"])
        c.gallery_poisson(dims, pp)
        c.setup(product_options(c, oopts))
        check_structure(c, h)


def test_gpu_aggregation_at_config_size(monkeypatch, capsys):
    """128^3 on 1 and 8 parts (config 2 / the shape of config 3's parts): identical aggregates; wall times to the log."""
    for pp in ((1, 1, 1), (2, 2, 2)):
        secs, aggs = [], []
        for mode in ("2", "0"):
            t0 = time.perf_counter()
            c = _setup("poisson", (128, 128, 128), pp, {}, monkeypatch, mode)
            secs.append(time.perf_counter() - t0)
            aggs.append([c.aggregates(l, p) for l in range(c.num_levels() - 1) for p in range(c.nparts)])
            c.close()
        for a, b in zip(*aggs):
            assert np.array_equal(a, b)
        with capsys.disabled():
            print(f"\n[setup 128^3, {int(np.prod(pp))} parts] device aggregation {secs[0]:.2f} s, host walk {secs[1]:.2f} s")
