"""Brute-force symbolic Cholesky (TEST INFRASTRUCTURE): the independent recomputation that the
host symbolic analysis (tulip.jl_b200/csrc/symbolic.cpp) must match bit-exactly.

The reference delegates this to CHOLMOD's analyse phase (src/KKT/Cholmod/spd.jl:17, sqd.jl:19)
and pins none of it in its tests, so "bit-exact" here means: for the permutation
the product chose, its elimination tree, column counts and supernodal row structure equal the
ones obtained by literally eliminating the permuted pattern column by column.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def kkt_pattern(A, system):
    """Symmetric pattern (full, boolean CSC) of the matrix factored for system 'K1' / 'K2'."""
    A = sp.csc_matrix(A)
    m, n = A.shape
    B = sp.csc_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape)
    if system == "K1":
        S = (B @ B.T + sp.identity(m)).tocsc()          # spd.jl:14  A*A' + I
    else:
        S = sp.bmat([[sp.identity(n), B.T], [B, sp.identity(m)]], format="csc")   # sqd.jl:13-16
    S.data[:] = 1.0
    return S


def symbolic_bruteforce(S, perm):
    """Eliminate P S P' column by column.  Returns (parent, colcount, cols) where cols[j] is the
    sorted array of row indices of L(:, j) (diagonal included), all in permuted numbering."""
    N = S.shape[0]
    perm = np.asarray(perm)
    Sp = sp.csc_matrix(S)[perm][:, perm].tocsc()
    Sp.sort_indices()
    struct = [None] * N
    children = [[] for _ in range(N)]
    parent = np.full(N, -1, np.int64)
    cc = np.zeros(N, np.int64)
    for j in range(N):
        rows = Sp.indices[Sp.indptr[j]:Sp.indptr[j + 1]]
        s = set(int(r) for r in rows if r >= j)
        s.add(j)
        for c in children[j]:
            s.update(r for r in struct[c] if r > c and r >= j)
        arr = np.array(sorted(s), dtype=np.int64)
        struct[j] = arr
        cc[j] = len(arr)
        if len(arr) > 1:
            parent[j] = arr[1]
            children[arr[1]].append(j)
    return parent, cc, struct


def supernode_rows_bruteforce(struct, sn_first):
    """Row list of each supernode (own columns first, then sorted union of the member columns'
    below rows) -- what sn_rows must equal."""
    out = []
    for s in range(len(sn_first) - 1):
        f, l = int(sn_first[s]), int(sn_first[s + 1])
        below = set()
        for j in range(f, l):
            below.update(int(r) for r in struct[j] if r >= l)
        out.append(np.array(list(range(f, l)) + sorted(below), dtype=np.int64))
    return out


def is_postordered(parent):
    """Every node's descendants form the contiguous range ending at the node."""
    N = len(parent)
    size = np.ones(N, np.int64)
    for j in range(N):
        if parent[j] >= 0:
            if parent[j] <= j:
                return False
            size[parent[j]] += size[j]
    first = np.arange(N) - size + 1
    for j in range(N):
        p = parent[j]
        if p >= 0 and not (first[p] <= first[j]):
            return False
    return True
