"""Device layouts (SELL-C-sigma slices, CSR-stream row blocks, boundary-row lists) are built on the host by the
same code the upload uses (csrc/formats.hpp) and queried through the C ABI, so their invariants are checked here
without a GPU: every CSR entry appears exactly once in its row's slot, padding is (0.0, valid column), the row
permutation stays inside its sigma window and sorts by length, row blocks respect the buffer limits, boundary rows
are exactly the rows with ghost columns, and y = A x computed FROM the layout equals the CSR product bit for bit."""
import numpy as np
import pytest
import scipy.sparse as sp

from parallel_amg_b200 import _lib as L
from util import det_vector


@pytest.fixture(scope="module")
def ctx():
    c = L.Context(4)
    c.gallery_poisson((17, 15, 13), (2, 2, 1))
    c.setup(c.default_options(coarse_size=40))
    return c


def blocks(c):
    for l in range(c.num_levels() - 1):
        for p in range(c.nparts):
            for which in (L.A_OO, L.P_OO, L.R_OO):
                yield l, p, which


def csr_of(c, l, p, which):
    ip, ix, d = c.block(l, p, which)
    i = c.level_info(l, p)
    ncols = {L.A_OO: i.n_own, L.P_OO: i.n_own_coarse, L.R_OO: i.n_own}[which]
    return sp.csr_matrix((d, ix, ip), shape=(len(ip) - 1, max(ncols, 1)))


@pytest.mark.parametrize("C,sigma", [(32, 1), (64, 1), (64, 256), (32, 96), (128, -4)])
def test_sell_layout_invariants(ctx, C, sigma):
    for l, p, which in blocks(ctx):
        A = csr_of(ctx, l, p, which)
        nr = A.shape[0]
        lay = ctx.layout_sell(l, p, which, C, sigma)
        off, col, val, perm = lay["slice_off"], lay["col"], lay["val"], lay["perm"]
        lens = np.diff(A.indptr)
        assert sorted(perm.tolist()) == list(range(nr))                       # a permutation
        if sigma < 0:                                                          # interleaved: lane * R + k  <->  row k * 32 + lane
            R = -sigma
            full = (nr // C) * C
            q = np.arange(full)
            assert np.array_equal(perm[:full], (q // C) * C + (q % R) * 32 + (q % C) // R)
            assert np.array_equal(perm[full:], np.arange(full, nr))
        elif sigma <= 1:
            assert not lay["permuted"] and np.array_equal(perm, np.arange(nr))
        else:
            for w0 in range(0, nr, sigma):                                     # inside its window, longest rows first
                w = perm[w0:w0 + sigma]
                assert w.min() >= w0 and w.max() < w0 + sigma
                assert np.all(np.diff(lens[w]) <= 0)
        x = det_vector(A.shape[1], 7)
        y = np.zeros(nr)
        for sl in range(len(off) - 1):
            width = off[sl + 1] - off[sl]
            for q in range(C):
                slot = sl * C + q
                if slot >= nr:
                    continue
                r = perm[slot]
                idx = (off[sl] + np.arange(width)) * C + q
                cc, vv = col[idx], val[idx]
                n = lens[r]
                assert width >= n
                assert np.array_equal(cc[:n], A.indices[A.indptr[r]:A.indptr[r + 1]])
                assert np.array_equal(vv[:n], A.data[A.indptr[r]:A.indptr[r + 1]])
                assert np.all(vv[n:] == 0.0) and np.all((cc[n:] >= 0) & (cc[n:] < A.shape[1]))
                s = 0.0
                for k in range(width):                                         # the kernel's order: sequential, product rounded first
                    s = s + vv[k] * x[cc[k]]
                y[r] = s
        ref = np.zeros(nr)
        for r in range(nr):
            s = 0.0
            for k in range(A.indptr[r], A.indptr[r + 1]):
                s = s + A.data[k] * x[A.indices[k]]
            ref[r] = s
        assert np.array_equal(y, ref)


def test_stream_row_blocks_respect_limits(ctx):
    for l, p, which in blocks(ctx):
        A = csr_of(ctx, l, p, which)
        for max_rows, max_entries in ((1024, 3069), (256, 500), (7, 64)):
            out = ctx.layout_stream(l, p, which, max_rows, max_entries)
            lens = np.diff(A.indptr)
            if out is None:
                assert lens.max() > max_entries
                continue
            r, e = out
            assert r[0] == 0 and r[-1] == A.shape[0] and e[-1] == A.nnz
            assert np.all(np.diff(r) >= 1) and np.all(np.diff(r) <= max_rows)
            assert np.array_equal(e, A.indptr[r])
            assert np.all(np.diff(e) <= max_entries)
            for k in range(len(r) - 2):   # greedy: the next row would not have fitted
                nxt = r[k + 1]
                assert (r[k + 1] - r[k] == max_rows) or (A.indptr[nxt + 1] - e[k] > max_entries)


def test_boundary_rows_are_the_rows_with_ghost_columns(ctx):
    for l, p, which in blocks(ctx):
        oo = csr_of(ctx, l, p, which)
        ip, ix, d = ctx.block(l, p, which + 1)
        lay = ctx.layout_boundary(l, p, which)
        has_ghost = np.diff(ip) > 0
        assert np.array_equal(lay["rows"], np.flatnonzero(has_ghost))
        assert np.array_equal(lay["skip"].astype(bool), has_ghost)
        assert lay["lanes"] in (1, 2, 4, 8, 16, 32)
        for k, r in enumerate(lay["rows"]):
            a, m, b = lay["ptr"][k], lay["mid"][k], lay["ptr"][k + 1]
            assert np.array_equal(lay["col"][a:m], oo.indices[oo.indptr[r]:oo.indptr[r + 1]])
            assert np.array_equal(lay["val"][a:m], oo.data[oo.indptr[r]:oo.indptr[r + 1]])
            assert np.array_equal(lay["col"][m:b], ix[ip[r]:ip[r + 1]])
            assert np.array_equal(lay["val"][m:b], d[ip[r]:ip[r + 1]])


@pytest.mark.parametrize("C,sigma", [(64, 1), (64, 256), (128, -4)])
def test_value_indexed_sell_storage(ctx, C, sigma):
    """kernels.cuh k_spmv_sell_vi reads dict[vidx[k]] instead of val[k]: the dictionary must reproduce every stored value BIT FOR BIT
    (padding -> entry 0 = +0.0), hold <= 255 distinct non-zero patterns in ascending pattern order, and blocks with more distinct values
    (the Galerkin matrices of the coarse levels) must be reported as not indexable."""
    seen = {1: 0, 2: 0, 0: 0}
    for l, p, which in blocks(ctx):
        lay = ctx.layout_sell(l, p, which, C, sigma)
        vi = ctx.layout_sell_values(l, p, which, C, sigma)
        vals = ctx.block(l, p, which)[2]
        distinct = np.unique(vals.view(np.uint64))
        distinct = distinct[distinct != 0]
        if len(distinct) > 4095:
            assert vi is None
            seen[0] += 1
            continue
        assert vi is not None
        d, idx = vi
        width = 1 if len(distinct) <= 255 else 2
        assert idx.dtype.itemsize == width
        seen[width] += 1
        assert len(idx) == len(lay["val"])
        assert np.array_equal(d[idx].view(np.uint64), lay["val"].view(np.uint64))      # bit-identical values, padding included
        assert d.view(np.uint64)[0] == 0
        used = d.view(np.uint64)[1:1 + len(distinct)]
        assert np.array_equal(used, distinct) and np.all(d.view(np.uint64)[1 + len(distinct):] == 0)
    assert seen[1] >= 3 and seen[2] >= 1     # level 0 of Poisson has 2 / 9 / 9 distinct values; coarse-level blocks have hundreds


def test_value_dictionary_keeps_signed_zero_and_counts_exactly():
    """-0.0 is a value of its own (its products keep their sign); exactly 255 distinct non-zero values still fit one byte, 256 need two."""
    for ndist, expect in ((255, True), (256, False)):
        n = 600
        off = np.ones(n - 1)                                                   # distinct patterns: 1.0, 2 .. k, -0.0, the diagonal = k + 2
        k = ndist - 2
        off[:k - 1] = np.arange(2, k + 1, dtype=np.float64)
        off[300] = -0.0
        i = np.arange(n - 1)
        A = sp.coo_matrix((np.concatenate([off, np.full(n, 1000.0), off]), (np.concatenate([i + 1, np.arange(n), i]),
                                                                            np.concatenate([i, np.arange(n), i + 1]))), shape=(n, n)).tocsr()
        A.sort_indices()                                                       # explicit (signed) zeros are kept
        c = L.Context(1)
        c.set_matrix_global(A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64), np.zeros(n, np.int32))
        c.setup(c.default_options(coarse_size=10 ** 6))                      # single level: the block is the matrix
        vi = c.layout_sell_values(0, 0, L.A_OO, 64, 1)
        assert vi is not None and (vi[1].dtype.itemsize == 1) == expect
        d, idx = vi
        lay = c.layout_sell(0, 0, L.A_OO, 64, 1)
        assert np.array_equal(d[idx].view(np.uint64), lay["val"].view(np.uint64))
        assert (d.view(np.uint64) == np.float64(-0.0).view(np.uint64)).sum() == 1
