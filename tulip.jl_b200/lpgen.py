"""Deterministic synthetic LP generators for the BASELINE.json configs.

All Float64, ``numpy.random.default_rng(PCG64(seed))``, seed = 20260925 + config#.
LPs are primal-dual feasible by construction: x0~U(0.5,1.5), b=A x0, y0~N(0,1), z0~U(0.5,1.5),
c=A'y0+z0, l=0, u=+inf unless stated.  Each generator returns a ``StdLP`` already in the
standard form the KKT boundary sees (``A x = b, l <= x <= u`` -- src/IPM/ipmdata.jl:6-12),
i.e. what ``IPMData`` would hold with Presolve_Level=0.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

BASE_SEED = 20260925


@dataclass
class StdLP:
    name: str
    A: sp.csc_matrix
    b: np.ndarray
    c: np.ndarray
    l: np.ndarray
    u: np.ndarray
    c0: float = 0.0
    objsense: bool = True
    meta: dict | None = None

    @property
    def shape(self):
        return self.A.shape


def _finish(name, A, rng, u=None, meta=None):
    A = sp.csc_matrix(A, dtype=np.float64)
    A.sum_duplicates()
    A.sort_indices()
    m, n = A.shape
    x0 = rng.uniform(0.5, 1.5, n)
    y0 = rng.standard_normal(m)
    z0 = rng.uniform(0.5, 1.5, n)
    b = A @ x0
    c = A.T @ y0 + z0
    l = np.zeros(n)
    if u is None:
        u = np.full(n, np.inf)
    else:
        u = np.asarray(u, float)
        u = np.maximum(u, x0 + 0.5)
    md = dict(m=m, n=n, nnz=int(A.nnz))
    md.update(meta or {})
    return StdLP(name, A, b, c, l, u, meta=md)


def _rows_without_replacement(rng, ncols, k, lo, hi):
    """(ncols, k) int array; row r of it holds k distinct ints in [lo[r], hi[r])."""
    lo = np.broadcast_to(np.asarray(lo, np.int64), (ncols,))
    hi = np.broadcast_to(np.asarray(hi, np.int64), (ncols,))
    span = (hi - lo)
    assert np.all(span >= k)
    R = lo[:, None] + (rng.random((ncols, k)) * span[:, None]).astype(np.int64)
    for _ in range(64):
        R.sort(axis=1)
        dup = np.zeros_like(R, dtype=bool)
        dup[:, 1:] = R[:, 1:] == R[:, :-1]
        nd = int(dup.sum())
        if nd == 0:
            break
        rr, cc = np.nonzero(dup)
        R[rr, cc] = lo[rr] + (rng.random(nd) * span[rr]).astype(np.int64)
    else:  # pragma: no cover
        raise RuntimeError("could not draw distinct rows")
    R.sort(axis=1)
    return R


def random_sparse(m, n, nnz_per_col, seed=BASE_SEED + 2, name="random_sparse"):
    """Configs 2 / T: rows uniform w/o replacement, values N(0,1); A[i,i] += 1 for i < m."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    k = nnz_per_col
    R = _rows_without_replacement(rng, n, k, 0, m)
    V = rng.standard_normal((n, k))
    cols = np.repeat(np.arange(n, dtype=np.int64), k)
    A = sp.coo_matrix((V.ravel(), (R.ravel(), cols)), shape=(m, n))
    A = A + sp.coo_matrix((np.ones(m), (np.arange(m), np.arange(m))), shape=(m, n))
    return _finish(name, A, rng, meta=dict(kind="uniform-random", nnz_per_col=k))


def banded_random(m, n, nnz_per_col, width, seed=BASE_SEED + 12, name="banded_random"):
    """Bounded-fill variant (SURVEY 8d row T option B): rows of column j within a window of
    ``width`` around floor(j*m/n); plus A[i,i] += 1."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    k = nnz_per_col
    centre = (np.arange(n, dtype=np.int64) * m) // n
    lo = np.clip(centre - width // 2, 0, max(m - width, 0))
    hi = np.minimum(lo + width, m)
    R = _rows_without_replacement(rng, n, k, lo, hi)
    V = rng.standard_normal((n, k))
    cols = np.repeat(np.arange(n, dtype=np.int64), k)
    A = sp.coo_matrix((V.ravel(), (R.ravel(), cols)), shape=(m, n))
    A = A + sp.coo_matrix((np.ones(m), (np.arange(m), np.arange(m))), shape=(m, n))
    return _finish(name, A, rng, meta=dict(kind="banded-random", nnz_per_col=k, width=width))


def staircase(stages=128, nodes=820, arcs=1200, couple=0.10, seed=BASE_SEED + 3, name="staircase"):
    """Config 3 (ken-18-shaped): per stage a node-arc incidence block (+1/-1) of a random digraph;
    ``couple`` of the arcs also get one entry in the next stage's rows; all columns u~U(2,10)."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    rows, cols, vals = [], [], []
    for s in range(stages):
        tail = rng.integers(0, nodes, arcs)
        head = (tail + 1 + rng.integers(0, nodes - 1, arcs)) % nodes
        j = s * arcs + np.arange(arcs)
        rows += [s * nodes + tail, s * nodes + head]
        cols += [j, j]
        vals += [np.ones(arcs), -np.ones(arcs)]
        if s + 1 < stages:
            pick = np.nonzero(rng.random(arcs) < couple)[0]
            rows.append((s + 1) * nodes + rng.integers(0, nodes, len(pick)))
            cols.append(j[pick])
            vals.append(rng.uniform(0.5, 1.5, len(pick)))
    m, n = stages * nodes, stages * arcs
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(m, n))
    # node-arc incidence blocks are rank deficient by one per connected component; that is what
    # the dual regularisation Rd is for (reference: src/IPM/HSD/step.jl:29-31).
    u = rng.uniform(2.0, 10.0, n)
    return _finish(name, A, rng, u=u, meta=dict(kind="staircase", stages=stages, nodes=nodes, arcs=arcs))


def block_angular(blocks=64, mb=1536, nb=3072, nnz_per_col=4, width=128, link=512, plink=0.25,
                  seed=BASE_SEED + 4, name="block_angular"):
    """Config 4: ``blocks`` independent banded-random blocks + ``link`` linking rows (last rows);
    each column has one linking entry with probability ``plink``."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    rows, cols, vals = [], [], []
    for k in range(blocks):
        centre = (np.arange(nb, dtype=np.int64) * mb) // nb
        lo = np.clip(centre - width // 2, 0, max(mb - width, 0))
        hi = np.minimum(lo + width, mb)
        R = _rows_without_replacement(rng, nb, nnz_per_col, lo, hi)
        V = rng.standard_normal((nb, nnz_per_col))
        rows.append((k * mb + R).ravel())
        cols.append(np.repeat(k * nb + np.arange(nb, dtype=np.int64), nnz_per_col))
        vals.append(V.ravel())
        d = np.arange(mb)
        rows.append(k * mb + d); cols.append(k * nb + d); vals.append(np.ones(mb))
    m, n = blocks * mb + link, blocks * nb
    pick = np.nonzero(rng.random(n) < plink)[0]
    rows.append(blocks * mb + rng.integers(0, link, len(pick)))
    cols.append(pick)
    vals.append(rng.standard_normal(len(pick)))
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(m, n))
    return _finish(name, A, rng, meta=dict(kind="block-angular", blocks=blocks, mb=mb, nb=nb, link=link))


def dense_columns(m=50_000, n=100_000, ndense=8, dense_nnz=25_000, sparse_nnz=300_000, width=256,
                  seed=BASE_SEED + 5, name="dense_columns"):
    """Config 5: ``ndense`` columns with ``dense_nnz`` non-zeros each + banded-random remainder."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    ns = n - ndense
    k = max(1, sparse_nnz // ns)
    centre = (np.arange(ns, dtype=np.int64) * m) // ns
    lo = np.clip(centre - width // 2, 0, max(m - width, 0))
    hi = np.minimum(lo + width, m)
    R = _rows_without_replacement(rng, ns, k, lo, hi)
    V = rng.standard_normal((ns, k))
    rows = [R.ravel(), np.arange(m)]
    cols = [np.repeat(np.arange(ns, dtype=np.int64), k), np.arange(m)]
    vals = [V.ravel(), np.ones(m)]
    for d in range(ndense):
        r = rng.choice(m, size=dense_nnz, replace=False)
        rows.append(r); cols.append(np.full(dense_nnz, ns + d)); vals.append(rng.standard_normal(dense_nnz))
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(m, n))
    return _finish(name, A, rng, meta=dict(kind="dense-columns", ndense=ndense, dense_nnz=dense_nnz))


# ---- the BASELINE.json configs by number (full size) and scaled-down "mini" versions for tests
def config(num, mini=False):
    if num == 2:
        return random_sparse(10_000, 20_000, 10, name="cfg2_random_1e4") if not mini else \
            random_sparse(300, 600, 6, name="cfg2_mini")
    if num == "T" or num == 6:
        return random_sparse(100_000, 200_000, 5, seed=BASE_SEED + 6, name="cfgT_random_1e5") if not mini else \
            random_sparse(500, 1000, 4, seed=BASE_SEED + 6, name="cfgT_mini")
    if num == 3:
        return staircase(name="cfg3_staircase") if not mini else \
            staircase(stages=6, nodes=40, arcs=60, name="cfg3_mini")
    if num == 4:
        return block_angular(name="cfg4_block_angular") if not mini else \
            block_angular(blocks=4, mb=96, nb=192, width=32, link=16, name="cfg4_mini")
    if num == 5:
        return dense_columns(name="cfg5_dense_cols") if not mini else \
            dense_columns(m=400, n=800, ndense=3, dense_nnz=200, sparse_nnz=2400, width=48, name="cfg5_mini")
    raise ValueError(f"unknown config {num}")
