"""the walk-free restatement of the greedy aggregation that csrc/setup_gpu.cu runs on the device (lexicographically first
independent sets as fixed points) gives the oracle's aggregates exactly -- on the gallery graphs, with strength filtering, on the
elasticity node graph, and on random symmetric graphs where passes 2 and 3 have real work."""
import numpy as np
import pytest
import scipy.sparse as sp

import amg_oracle as O
from agg_parallel_emulation import aggregate_parallel


def _check(S, owner, nparts):
    _, counts_ref, loc_ref = O.aggregate_parts(S, owner, nparts)
    loc, counts = aggregate_parallel(S, owner, nparts)
    assert np.array_equal(counts, counts_ref)
    assert np.array_equal(loc, loc_ref)


@pytest.mark.parametrize("dims,pp", [((30, 30), (2, 2)), ((12, 12, 12), (2, 2, 2)), ((17, 9, 11), (3, 1, 2)), ((40,), (1,)), ((25, 13), (1, 1))])
def test_poisson_graphs(dims, pp):
    A = O.poisson_fd(dims)
    owner = O.uniform_partition(pp, dims)
    _check(O.strength_graph(A, owner, 0.0), owner, int(np.prod(pp)))


def test_filtered_strength_jump_coefficients():
    dims, pp = (16, 16, 16), (2, 2, 1)
    A = O.diffusion_fv(dims, O.jump_coefficient_k(dims, blocks=4, kmax=1.0e4, eps_z=1.0e-3))
    owner = O.uniform_partition(pp, dims)
    for eps in (0.08, 0.25):
        _check(O.strength_graph(A, owner, eps), owner, 4)


def test_elasticity_node_graph():
    dims, pp = (7, 6, 5), (2, 1, 1)
    A, coords = O.elasticity_q1(dims)
    owner = O.uniform_partition(pp, dims)
    N = O.node_graph(A, 3)
    _check(O.strength_graph(N, owner, 0.0), owner, 2)


@pytest.mark.parametrize("seed", range(6))
def test_random_symmetric_graphs(seed):
    rng = np.random.default_rng(seed)
    n, nparts = 400, 3
    M = sp.random(n, n, density=0.012 + 0.004 * seed, random_state=rng, format="csr")
    M = (M + M.T).tocsr()
    M.data[:] = 1.0
    owner = rng.integers(0, nparts, n).astype(np.int32)        # interleaved ownership: local order != global order gaps
    A = (M + sp.identity(n)).tocsr()
    _check(O.strength_graph(A, owner, 0.0), owner, nparts)
