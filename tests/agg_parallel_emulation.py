"""numpy restatement of csrc/setup_gpu.cu gpu_aggregate: the greedy three-pass aggregation WITHOUT the sequential
walk, as fixed points over the whole graph (all parts at once, global ids).  The CUDA kernels cannot run without a GPU; this file
states the same rules statement by statement so that the CPU suite pins the restatement against the oracle's walk
(amg_oracle._aggregate_py) -- the -m gpu tests then pin the kernels against the host walk bit for bit.  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import scipy.sparse as sp

U, ROOT, NOT = 0, 1, 2


def _first_mis(S, state, dist):
    """lexicographically first maximal independent set of S^dist:
    i -> NOT once an earlier vertex within `dist` is a ROOT, ROOT once all of them are NOT (k_mis_round)."""
    n = S.shape[0]
    G = S.astype(np.int64)
    if dist == 2:
        G = (G @ G + G).tocsr()
        G.setdiag(0)
        G.eliminate_zeros()
    low = sp.tril(G, -1).tocsr()
    low.data[:] = 1
    rounds = 0
    while (state == U).any():
        rounds += 1
        has_root = (low @ (state == ROOT).astype(np.int64)) > 0
        blocked = (low @ (state == U).astype(np.int64)) > 0
        u = state == U
        state[u & has_root] = NOT
        state[u & ~has_root & ~blocked] = ROOT
    return rounds


def _root_of(S, state):
    """for every vertex: itself when it is a ROOT, else its adjacent ROOT (at most one: closed neighbourhoods of roots are disjoint), else -1
    (k_agg_pass1)."""
    n = S.shape[0]
    r = np.where(state == ROOT, np.arange(n), -1)
    rows = np.repeat(np.arange(n), np.diff(S.indptr))
    hit = state[S.indices] == ROOT
    first = np.full(n, -1, np.int64)
    # ascending columns inside a row: the first hit of a row is its smallest adjacent root
    rr, cc = rows[hit][::-1], S.indices[hit][::-1]
    first[rr] = cc
    return np.where(r >= 0, r, first)


def aggregate_parallel(S, owner, nparts):
    """S: symmetric boolean CSR of the same-part strong off-diagonals (amg_oracle.strength_graph).  Returns (agg_loc, counts)."""
    S = S.tocsr()
    S.sort_indices()
    n = S.shape[0]
    assert (abs(S - S.T)).nnz == 0, "the restatement needs a symmetric strength graph (the device falls back otherwise)"
    order = np.lexsort((np.arange(n), owner))                 # part-major position
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n)
    part_off = np.concatenate([[0], np.cumsum(np.bincount(owner, minlength=nparts))])

    def local_rank(state):
        flags = np.zeros(n + 1, np.int64)
        flags[pos] = state == ROOT
        rank = np.concatenate([[0], np.cumsum(flags)[:-1]])   # exclusive scan
        base = rank[part_off]
        return rank, base

    state = np.zeros(n, np.int8)
    _first_mis(S, state, 2)                                   # pass 1
    rank, base1 = local_rank(state)
    r = _root_of(S, state)
    agg1 = np.where(r >= 0, rank[pos[np.maximum(r, 0)]] - base1[owner[np.maximum(r, 0)]], -1)
    rows = np.repeat(np.arange(n), np.diff(S.indptr))         # pass 2: first neighbour that pass 1 aggregated
    hit = agg1[S.indices] != -1
    first = np.full(n, -1, np.int64)
    first[rows[hit][::-1]] = agg1[S.indices[hit][::-1]]
    agg2 = np.where(agg1 != -1, agg1, first)
    # pass 3 is empty for a symmetric S: the root set is maximal in S^2, so every vertex has a root within distance 2 and passes 1 / 2
    # reach it (the device returns "not done" otherwise and the host walks the rows)
    assert (agg2 >= 0).all()
    return agg2, np.diff(base1)
