"""Host setup (C++ product code) vs the oracle: index maps, aggregates, halo plans and the CSR
structure of every split block must be BIT-EXACT; values within 1e-12 of the block's max entry.
(BASELINE.json north_star "Correctness".)"""
import numpy as np
import pytest

import amg_oracle as O
from parallel_amg_b200 import _lib as L
from util import oracle_problem, product_options

CASES = [
    ((200, 200), (2, 2), {}),                       # BASELINE config 1
    ((20, 20, 20), (2, 2, 2), {}),
    ((33, 31, 17), (3, 2, 1), {}),                  # ragged blocks, odd sizes
    ((28, 28, 28), (1, 1, 1), {}),
    ((24, 24, 24), (2, 2, 1), {"eps_strength": 0.0831}),
    ((64, 64), (4, 1), {"coarse_size": 50}),
    ((7, 5), (1, 1), {}),                           # single level (n <= coarse_size)
]


def check_structure(c, h):
    nl = c.num_levels()
    assert nl == len(h["levels"])
    worst = 0.0
    for l in range(nl):
        for p in range(h["nparts"]):
            d = h["levels"][l]["parts"][p]
            own, gh, gho = c.index_maps(l, p)
            assert np.array_equal(own, d["own_to_global"])
            assert np.array_equal(gh, d["ghost_to_global"])
            assert np.array_equal(gho, d["ghost_to_owner"])
            for b, name in enumerate(L.BLOCK_NAMES):
                if name not in d:
                    continue
                ip, ix, dd = c.block(l, p, b)
                m = d[name]
                assert np.array_equal(ip, m.indptr), (l, p, name)
                assert np.array_equal(ix, m.indices), (l, p, name)
                if m.nnz:
                    worst = max(worst, float(np.abs(dd - m.data).max() / np.abs(m.data).max()))
            if "agg_local" in d:
                assert np.array_equal(c.aggregates(l, p), d["agg_local"])
            hp = c.halo_plan(l, p)
            pl = h["levels"][l]["plan"][p]
            assert [int(x) for x in hp["recv_part"]] == [r[0] for r in pl["recv"]]
            assert [int(x) for x in hp["recv_slot0"]] == [r[1] for r in pl["recv"]]
            assert [int(x) for x in hp["recv_count"]] == [r[2] for r in pl["recv"]]
            assert [int(x) for x in hp["send_part"]] == [s[0] for s in pl["send"]]
            assert [int(x) for x in hp["send_slot0"]] == [s[2] for s in pl["send"]]
            ref_idx = np.concatenate([s[1] for s in pl["send"]]) if pl["send"] else np.zeros(0, np.int32)
            assert np.array_equal(hp["send_idx"], ref_idx)
            dg, dl1 = c.diag(l, p)
            assert np.allclose(dg, d["diag"], rtol=1e-12, atol=0)
            assert np.allclose(dl1, d["diag_l1"], rtol=1e-12, atol=0)
            info = c.level_info(l, p)
            assert abs(info.rho - h["rho_dinv_a"][l]) <= 1e-12 * abs(info.rho)
    assert worst <= 1e-12
    inv = c.coarse_inverse()
    assert np.abs(inv - h["coarse_inv"]).max() <= 1e-10 * np.abs(inv).max()


@pytest.mark.parametrize("dims,pp,oopts", CASES)
def test_gallery_setup_matches_oracle(dims, pp, oopts):
    A, owner, h = oracle_problem(dims, pp, tuple(sorted(oopts.items())))
    c = L.Context(h["nparts"])
    c.gallery_poisson(dims, pp)
    n, nnz = c.global_size()
    assert (n, nnz) == (A.shape[0], A.nnz)
    c.setup(product_options(c, oopts))
    check_structure(c, h)


def test_set_matrix_global_and_part_rows_agree():
    dims, pp = (18, 14, 10), (2, 1, 2)
    A, owner, h = oracle_problem(dims, pp)
    c1 = L.Context(4)
    c1.set_matrix_global(A.indptr, A.indices, A.data, owner)
    c1.setup()
    check_structure(c1, h)
    c2 = L.Context(4)
    for p in range(4):  # PSparseMatrix-style: each part hands over its own rows, global columns
        own = np.flatnonzero(owner == p)
        sub = A[own].tocsr()
        c2.set_part_rows(p, own, sub.indptr, sub.indices, sub.data)
    c2.setup()
    check_structure(c2, h)


def test_jump_coefficient_gallery_matches_oracle():
    dims, pp = (16, 16, 16), (2, 2, 2)
    A = O.diffusion_fv(dims, O.jump_coefficient_k(dims, blocks=4, kmax=1e4, eps_z=1e-3))
    owner = O.uniform_partition(pp, dims)
    c = L.Context(8)
    c.gallery_diffusion_jump(dims, pp, blocks=4, kmax=1e4, eps_z=1e-3)
    c2 = L.Context(8)
    c2.set_matrix_global(A.indptr, A.indices, A.data, owner)
    x = np.linspace(-1, 1, A.shape[0])
    assert np.array_equal(c.host_matvec_global(x), c2.host_matvec_global(x))  # gallery is bit-exact
    oopts = {"eps_strength": 0.0831}
    h = O.build(A, owner, 8, oopts)
    c.setup(product_options(c, oopts))
    check_structure(c, h)


@pytest.mark.parametrize("dims,pp", [((8, 7, 6), (2, 1, 1)), ((9, 8, 7), (2, 2, 1)), ((6, 5, 5), (1, 1, 1)), ((10, 9, 3), (3, 2, 1))])
def test_elasticity_gallery_and_nullspace_setup_match_oracle(dims, pp):
    """BASELINE config 4 at test size: Q1 elasticity, 3 DOFs per node, rigid-body near-nullspace.  The gallery
    (integer element sums) and the rigid-body modes are bit-exact; node aggregates, index maps, halo plans and
    the structure of every block are bit-exact; values (per-aggregate Householder QR) to 1e-12."""
    nparts = int(np.prod(pp))
    A, coords = O.elasticity_q1(dims)
    B = O.rigid_body_modes(coords)
    owner = np.repeat(O.uniform_partition(pp, dims), 3).astype(np.int32)
    c = L.Context(nparts)
    c.gallery_elasticity(dims, pp)
    n, nnz = c.global_size()
    assert (n, nnz) == (A.shape[0], A.nnz)
    c2 = L.Context(nparts)
    c2.set_matrix_global(A.indptr, A.indices, A.data, owner)
    x = np.linspace(-1, 1, n)
    assert np.array_equal(c.host_matvec_global(x), c2.host_matvec_global(x))
    bs, Bp = c.near_nullspace()
    assert bs == 3 and np.array_equal(Bp, B)
    oopts = dict(block_size=3, nullspace=B, coarse_size=60)
    h = O.build(A, owner, nparts, oopts)
    assert len(h["levels"]) >= 2
    c.setup(c.default_options(coarse_size=60))
    check_structure(c, h)
    # the same hierarchy from a caller-supplied matrix + near-nullspace
    c2.set_near_nullspace(3, B)
    c2.setup(c2.default_options(coarse_size=60))
    check_structure(c2, h)


def test_part_without_rows():
    """An index partition may leave a part empty (PartitionedArrays allows it): the setup keeps it as a part with
    zero own rows on every level, exactly like the oracle."""
    A = O.poisson_fd((12, 12, 12))
    n = A.shape[0]
    owner = np.zeros(n, np.int32)
    owner[n // 2:] = 2                      # part 1 owns nothing
    h = O.build(A, owner, 3)
    c = L.Context(3)
    c.set_matrix_global(A.indptr, A.indices, A.data, owner)
    c.setup()
    check_structure(c, h)
    assert all(c.level_info(l, 1).n_own == 0 for l in range(c.num_levels()))


def test_hierarchy_save_load_roundtrip(tmp_path):
    """Rank-0-builds / others-load hand-off: a saved hierarchy loads back bit for bit; with keep_part the other
    parts come back as metadata (sizes, nnz, halo neighbours) except on the small replicated-tail levels."""
    dims, pp = (20, 20, 20), (2, 2, 1)
    A, owner, h = oracle_problem(dims, pp)
    c = L.Context(4)
    c.gallery_poisson(dims, pp)
    c.setup(c.default_options(tail_rows=700))
    path = str(tmp_path / "h.bin")
    c.hierarchy_save(path)
    c_all = L.Context(4)
    c_all.hierarchy_load(path)
    check_structure(c_all, h)
    c1 = L.Context(4)
    c1.hierarchy_load(path, keep_part=1)
    nl = c1.num_levels()
    for l in range(nl):
        small = c1.level_info(l, 0).n_global <= 700 or l == nl - 1
        for p in range(4):
            a, b = c1.level_info(l, p), c.level_info(l, p)
            assert (a.n_own, a.n_ghost, list(a.nnz), a.n_send, a.n_recv_nbrs, a.n_send_nbrs) == \
                   (b.n_own, b.n_ghost, list(b.nnz), b.n_send, b.n_recv_nbrs, b.n_send_nbrs)
            if p == 1 or small:
                assert np.array_equal(c1.index_maps(l, p)[0], c.index_maps(l, p)[0])
                for blk in range(6):
                    if l == nl - 1 and blk >= L.P_OO:
                        continue
                    assert all(np.array_equal(x, y) for x, y in zip(c1.block(l, p, blk), c.block(l, p, blk)))
            else:
                with pytest.raises(L.PamgError):
                    c1.index_maps(l, p)
    with pytest.raises(L.PamgError):
        L.Context(2).hierarchy_load(path)            # written for 4 parts
    with pytest.raises(L.PamgError):
        c1.hierarchy_load(str(tmp_path / "missing.bin"))


def test_near_nullspace_argument_checks():
    c = L.Context(1)
    c.gallery_poisson((6, 6), (1, 1))
    with pytest.raises(L.PamgError):
        c.set_near_nullspace(5, np.ones((36, 2)))      # 36 rows are not a multiple of 5
    c.set_near_nullspace(1, np.ones((36, 1)))          # k = 1: normalised piecewise-constant tentative P
    c.setup(c.default_options(coarse_size=10))
    assert c.num_levels() >= 2
    c.set_near_nullspace(1, None)                      # back to scalar SA
    c.setup(c.default_options(coarse_size=10))


def test_uniform_partition_matches_oracle():
    import ctypes as C
    lib = L.load()
    for dims, pp in [((10, 7, 5), (3, 2, 2)), ((200, 200), (2, 2)), ((9,), (4,))]:
        n = int(np.prod(dims))
        out = np.empty(n, np.int32)
        d = np.asarray(dims, np.int64)
        q = np.asarray(pp, np.int32)
        st = lib.pamg_uniform_partition(len(dims), d.ctypes.data_as(C.POINTER(C.c_int64)),
                                        q.ctypes.data_as(C.POINTER(C.c_int32)), out.ctypes.data_as(C.POINTER(C.c_int32)))
        assert st == 0
        assert np.array_equal(out, O.uniform_partition(pp, dims))


def test_bad_arguments_return_status_not_abort():
    c = L.Context(2)
    with pytest.raises(L.PamgError) as e:
        c.setup()  # no matrix
    assert e.value.status == L.ERR_ARG
    with pytest.raises(L.PamgError):
        c.gallery_poisson((8, 8), (2, 2))  # 4 parts != 2
    with pytest.raises(L.PamgError):
        c.num_levels()
    c.gallery_poisson((8, 8), (2, 1))
    c.setup()
    with pytest.raises(L.PamgError) as e:
        c.spmv(0, [np.zeros(32), np.zeros(32)]) if hasattr(c, "local_parts") else c.lib and (_ for _ in ()).throw(
            L.PamgError(L.ERR_ARG, "device not initialised"))
    assert e.value.status == L.ERR_ARG
