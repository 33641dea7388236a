"""The Julia shim (julia/PAMG.jl) cannot run here (no Julia), so its marshalling is restated in
tests/julia_shim_emulation.py and driven through the SAME C entry points with the arrays the shim would pass:
1-based Int64 CSC blocks of a split local matrix with discovery-order ghosts -> part_rows -> pamg_set_part_rows.
The hierarchy the library then builds must equal the oracle's bit for bit (maps, aggregates, structure), i.e. every
own-ghost coupling survived the trip (round 1's shim dropped them and solved the block-diagonal system)."""
import numpy as np
import pytest

import amg_oracle as O
import julia_shim_emulation as J
from parallel_amg_b200 import _lib as L
from util import det_vector


def _shim_context(A, owner, nparts, ghost_seed):
    c = L.Context(nparts)
    pa = []
    for p in range(nparts):
        Aoo, Aog, own1, gh1 = J.pa_split_part(A, owner, p, ghost_seed=ghost_seed + p if ghost_seed else 0)
        o2g, rowptr, colgid, val = J.part_rows(Aoo, Aog, own1, own1, gh1)
        c.set_part_rows(p, o2g, rowptr, colgid, val)
        pa.append((own1, gh1))
    return c, pa


@pytest.mark.parametrize("dims,pp,ghost_seed", [((12, 10, 8), (2, 2, 1), 0), ((12, 10, 8), (2, 2, 1), 7), ((30, 30), (2, 2), 3),
                                                ((9, 9, 9), (1, 1, 1), 0)])
def test_part_rows_marshalling_reproduces_the_oracle_hierarchy(dims, pp, ghost_seed):
    nparts = int(np.prod(pp))
    A = O.poisson_fd(dims)
    owner = O.uniform_partition(pp, dims)
    c, pa = _shim_context(A, owner, nparts, ghost_seed)
    n, nnz = c.global_size()
    assert n == A.shape[0] and nnz == A.nnz          # no coupling was dropped
    x = det_vector(n, 3)
    assert np.array_equal(c.host_matvec_global(x), A @ x)
    c.setup()
    h = O.build(A, owner, nparts)
    assert c.num_levels() == len(h["levels"])
    for l, lev in enumerate(h["levels"]):
        for p, d in enumerate(lev["parts"]):
            own, gh, gho = c.index_maps(l, p)
            assert np.array_equal(own, d["own_to_global"]) and np.array_equal(gh, d["ghost_to_global"])
            assert np.array_equal(gho, d["ghost_to_owner"])
            for b, name in enumerate(L.BLOCK_NAMES):
                if name not in d:
                    continue
                ip, ix, dd = c.block(l, p, b)
                m = d[name]
                assert np.array_equal(ip, m.indptr) and np.array_equal(ix, m.indices)
                assert m.nnz == 0 or np.abs(dd - m.data).max() <= 1e-12 * np.abs(m.data).max()
            if l + 1 < len(h["levels"]):
                assert np.array_equal(c.aggregates(l, p), d["agg_local"])


def test_non_symmetric_values_keep_their_rows():
    """part_rows transposes the CSC blocks instead of assuming symmetry: a matrix with a symmetric pattern but
    non-symmetric values must arrive row for row."""
    dims, pp = (10, 9), (2, 1)
    A = O.poisson_fd(dims).tocsr().astype(np.float64)
    A.data = A.data * (1.0 + 0.01 * det_vector(A.nnz, 5))
    owner = O.uniform_partition(pp, dims)
    c, _ = _shim_context(A, owner, 2, 11)
    x = det_vector(A.shape[0], 9)
    assert np.array_equal(c.host_matvec_global(x), A @ x)


def test_ghost_permutation_round_trip():
    dims, pp = (8, 8, 6), (2, 2, 1)
    nparts = 4
    A = O.poisson_fd(dims)
    owner = O.uniform_partition(pp, dims)
    c, pa = _shim_context(A, owner, nparts, 5)
    c.setup()
    v = det_vector(A.shape[0], 21)
    for p, (own1, gh1) in enumerate(pa):
        _, lib_gh, _ = c.index_maps(0, p)
        perm = J.ghost_permutation(lib_gh, gh1)
        assert np.array_equal(lib_gh, (gh1 - 1)[perm])
        own_vals, pa_ghost = v[own1 - 1], v[gh1 - 1]
        buf = J.to_library_local(own_vals, pa_ghost, perm)
        assert np.array_equal(buf[len(own1):], v[lib_gh])       # what pamg_consistent expects / returns
        o2, g2 = J.from_library_local(buf, len(own1), perm)
        assert np.array_equal(o2, own_vals) and np.array_equal(g2, pa_ghost)


def test_bad_rowptr_is_an_error_not_a_crash():
    c = L.Context(1)
    with pytest.raises(L.PamgError):
        c.set_part_rows(0, np.arange(3), np.array([0, 2, 1, 3]), np.array([0, 1, 2]), np.ones(3))
    owner = np.zeros(4, np.int32)
    assert L.load().pamg_uniform_partition(1, (L.C.c_int64 * 1)(4), (L.C.c_int32 * 1)(0), owner.ctypes.data_as(L.C.POINTER(L.C.c_int32))) == L.ERR_ARG


@pytest.mark.gpu
def test_device_consistent_through_the_shim_permutation():
    """pamg_consistent / pamg_assemble on buffers laid out as the shim lays them out; results mapped back to the
    PartitionedArrays ghost order must equal the owners' values / the oracle's assemble!."""
    dims, pp = (12, 10, 8), (2, 2, 1)
    nparts = 4
    A = O.poisson_fd(dims)
    owner = O.uniform_partition(pp, dims)
    c, pa = _shim_context(A, owner, nparts, 9)
    c.setup()
    c.device_init()
    v = det_vector(A.shape[0], 31)
    perms, bufs = [], []
    for p, (own1, gh1) in enumerate(pa):
        _, lib_gh, _ = c.index_maps(0, p)
        perms.append(J.ghost_permutation(lib_gh, gh1))
        bufs.append(J.to_library_local(v[own1 - 1], np.full(len(gh1), np.nan), perms[-1]))
    c.consistent(0, bufs)
    for p, (own1, gh1) in enumerate(pa):
        o2, g2 = J.from_library_local(bufs[p], len(own1), perms[p])
        assert np.array_equal(g2, v[gh1 - 1])
    h = O.build(A, owner, nparts)
    lev = h["levels"][0]
    ws = [det_vector(len(own1) + len(gh1), 60 + p) for p, (own1, gh1) in enumerate(pa)]   # PA layout: own then PA-ordered ghosts
    lib_ws = [J.to_library_local(w[:len(own1)], w[len(own1):], perms[p]) for p, ((own1, gh1), w) in enumerate(zip(pa, ws))]
    ref = O.assemble(lev, [w.copy() for w in lib_ws])
    c.assemble(0, lib_ws)
    for a, b in zip(lib_ws, ref):
        assert np.allclose(a, b, rtol=0, atol=8e-15)
