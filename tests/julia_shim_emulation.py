"""Line-by-line numpy restatement of the marshalling in julia/PAMG.jl (which cannot run here: no Julia).

Inputs are what PartitionedArrays.jl holds per part of an assembled, split-format PSparseMatrix [RECALL-UNVERIFIED,
SURVEY.md App. A]: the own-own and own-ghost blocks as 1-based Int64 CSC (SparseMatrixCSC: colptr / rowval / nzval),
own_to_global (1-based, ascending) and ghost_to_global (1-based) in DISCOVERY order.  Outputs are exactly the arrays
the shim passes across the C ABI, so tests/test_julia_shim_marshalling.py can feed them to libpamg.so."""
import numpy as np
import scipy.sparse as sp


def pa_split_part(A, owner, p, ghost_seed=0):
    """What PartitionedArrays would hold for part p of the global CSR matrix A: (Aoo, Aog, own_to_global,
    ghost_to_global) with 1-based Int64 CSC blocks and ghosts in discovery order (order of first appearance in a
    row-major sweep of the own rows), optionally shuffled further (psparse's COO input order is arbitrary)."""
    A = A.tocsr()
    own = np.flatnonzero(owner == p).astype(np.int64)             # ascending gid
    rows = A[own]
    gcols = rows.indices.astype(np.int64)
    is_ghost = owner[gcols] != p
    seen, ghosts = set(), []
    for g in gcols[is_ghost]:                                     # discovery order
        if g not in seen:
            seen.add(int(g))
            ghosts.append(int(g))
    ghosts = np.array(ghosts, dtype=np.int64)
    if ghost_seed:
        ghosts = ghosts[np.random.default_rng(ghost_seed).permutation(len(ghosts))]
    lid_own = -np.ones(A.shape[0], np.int64)
    lid_own[own] = np.arange(len(own))
    lid_gh = -np.ones(A.shape[0], np.int64)
    lid_gh[ghosts] = np.arange(len(ghosts))
    r = np.repeat(np.arange(len(own)), np.diff(rows.indptr))
    Aoo = sp.csc_matrix((rows.data[~is_ghost], (r[~is_ghost], lid_own[gcols[~is_ghost]])), shape=(len(own), len(own)))
    Aog = sp.csc_matrix((rows.data[is_ghost], (r[is_ghost], lid_gh[gcols[is_ghost]])), shape=(len(own), len(ghosts)))

    def julia(M):  # SparseMatrixCSC fields, 1-based Int64
        M.sort_indices()
        return dict(colptr=M.indptr.astype(np.int64) + 1, rowval=M.indices.astype(np.int64) + 1, nzval=M.data.astype(np.float64),
                    m=M.shape[0], n=M.shape[1])

    return julia(Aoo), julia(Aog), own + 1, ghosts + 1


def _transpose_csc(M):
    """sparse(transpose(M)) for a 1-based CSC dict: again a 1-based CSC dict (rows sorted within each column)."""
    S = sp.csc_matrix((M["nzval"], M["rowval"] - 1, M["colptr"] - 1), shape=(M["m"], M["n"]))
    T = S.T.tocsc()
    T.sort_indices()
    return dict(colptr=T.indptr.astype(np.int64) + 1, rowval=T.indices.astype(np.int64) + 1, nzval=T.data, m=T.shape[0], n=T.shape[1])


def part_rows(Aoo, Aog, own_to_global_rows1, own_to_global_cols1, ghost_to_global_cols1):
    """julia/PAMG.jl part_rows: (o2g_rows0, rowptr0, colgid0, val) for pamg_set_part_rows.  Loops mirror the Julia
    source statement by statement (1-based indices kept in variables ending in 1)."""
    o2g_rows = own_to_global_rows1.astype(np.int64) - 1
    o2g = own_to_global_cols1.astype(np.int64) - 1
    g2g = ghost_to_global_cols1.astype(np.int64) - 1
    nown = len(o2g_rows)
    Too, Tog = _transpose_csc(Aoo), _transpose_csc(Aog)
    rowptr = np.zeros(nown + 1, np.int64)
    for i1 in range(1, nown + 1):
        rowptr[i1] = rowptr[i1 - 1] + (Too["colptr"][i1] - Too["colptr"][i1 - 1]) + (Tog["colptr"][i1] - Tog["colptr"][i1 - 1])
    colgid = np.empty(rowptr[-1], np.int64)
    val = np.empty(rowptr[-1], np.float64)
    for i1 in range(1, nown + 1):
        q = rowptr[i1 - 1]
        for k1 in range(Too["colptr"][i1 - 1], Too["colptr"][i1]):
            colgid[q] = o2g[Too["rowval"][k1 - 1] - 1]
            val[q] = Too["nzval"][k1 - 1]
            q += 1
        for k1 in range(Tog["colptr"][i1 - 1], Tog["colptr"][i1]):
            colgid[q] = g2g[Tog["rowval"][k1 - 1] - 1]
            val[q] = Tog["nzval"][k1 - 1]
            q += 1
    return o2g_rows, rowptr, colgid, val


def ghost_permutation(lib_ghost_gid0, pa_ghost_gid1):
    """perm (0-based here) with lib_ghost[k] = pa_ghost[perm[k]]."""
    pos = {int(g) - 1: k for k, g in enumerate(pa_ghost_gid1)}
    return np.array([pos[int(g)] for g in lib_ghost_gid0], dtype=np.int64)


def to_library_local(own_vals, pa_ghost_vals, perm):
    """device_halo!: the buffer handed to pamg_consistent / pamg_assemble = [own ; ghosts in library order]."""
    return np.concatenate([own_vals, pa_ghost_vals[perm]])


def from_library_local(buf, n_own, perm):
    """device_halo!: back to (own, ghosts in PartitionedArrays order)."""
    g = np.empty(len(buf) - n_own)
    g[perm] = buf[n_own:]
    return buf[:n_own].copy(), g
