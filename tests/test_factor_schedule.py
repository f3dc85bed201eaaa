"""CPU emulation of the numeric factorisation SCHEDULE (integer plan + NumPy arithmetic): replay the level plan exactly as
solver.cu enqueues it -- one-CTA supernodes, diagonal blocks, trsm, urgent tiles at their level; the lazy FP64 tiles as
late as the schedule allows (just before level L + 2) and the tcgen05 tasks just before level L + 3 -- on a dense copy of
the permuted matrix, and compare with a dense Cholesky.  A missing / duplicated contribution, a task that reads a column
piece before it is final, or a dependency rule that is too weak all show up as a wrong factor.
Reference: cholesky!(F, Symmetric(K)) of /root/reference/src/KKT/Cholmod/spd.jl:46."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

import tlpb200_loader

pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen  # noqa: E402


def _signed_chol(B, sg):
    """B = L diag(sg) L' without pivoting (the signed Cholesky of DESIGN 1; plain Cholesky when sg = +1)."""
    w = B.shape[0]
    L = np.zeros_like(B)
    for j in range(w):
        dj = B[j, j] - (L[j, :j] ** 2) @ sg[:j]
        assert dj * sg[j] > 0
        L[j, j] = np.sqrt(dj * sg[j])
        if j + 1 < w:
            L[j + 1:, j] = (B[j + 1:, j] - (L[j + 1:, :j] * sg[:j]) @ L[j, :j]) / (sg[j] * L[j, j])
    return L


def _emulate(A, backend_kwargs, theta_seed=0, system="K1"):
    A = sp.csc_matrix(A)
    m, n = A.shape
    k = pkg.setup(A, pkg.K1() if system == "K1" else pkg.K2(), pkg.Backend(analyze_only=True, **backend_kwargs))
    plan, sym = k.update_plan(), k.symbolic()
    LF = {f: i for i, f in enumerate(k.LEVEL_FIELDS)}
    assert plan["levels"].shape[1] == len(LF)
    rng = np.random.default_rng(theta_seed)
    d = 1.0 / np.exp(rng.uniform(-2, 2, n))
    p = sym["perm"]
    if system == "K1":
        K = (A @ sp.diags(d) @ A.T + sp.diags(np.full(m, 1e-3))).toarray()
        sg = np.ones(m)
    else:                                                   # systems.jl:8-32: [-(Theta^-1 + Rp)  A'; A  Rd]
        K = sp.bmat([[sp.diags(-(1.0 / d + 1e-3)), A.T], [A, sp.diags(np.full(m, 1e-3))]]).toarray()
        sg = np.where(p < n, -1.0, 1.0)                     # expected pivot signs in permuted order
    Kp = K[np.ix_(p, p)]
    W = np.tril(Kp).copy()
    first, rp, rows_all = sym["sn_first"], sym["sn_rowptr"], sym["sn_rows"]
    rows_of = lambda s: rows_all[rp[s]:rp[s + 1]]
    pieces, views = plan["pieces"], plan["views"]

    def apply_tile(T):                       # UpdTask: C[rows I, rows K] -= L[I, piece] L[K, piece]'
        piece, i0, ni, k0, nk, tgt, diag, _ = (int(x) for x in T)
        s, c0, c1, _lvl = (int(x) for x in pieces[piece])
        r = rows_of(s)
        I, Kc = r[i0:i0 + ni], r[k0:k0 + nk]
        upd = (W[np.ix_(I, np.arange(c0, c1))] * sg[c0:c1]) @ W[np.ix_(Kc, np.arange(c0, c1))].T
        if diag:
            upd = np.tril(upd)
        W[np.ix_(I, Kc)] -= upd

    def apply_oz(T, tile):                   # OzTask: same-supernode, panel row blocks, K chunk range
        view, rbA, rbB, half, k0, k1, _, _ = (int(x) for x in T)
        s = int(views[view, 0])
        f, nc = int(first[s]), int(first[s + 1] - first[s])
        r = rows_of(s)
        ia = np.arange(rbA * 128, min(len(r), rbA * 128 + 128))
        jb = np.arange(rbB * 128 + half * 64, min(nc, rbB * 128 + (128 if tile == 128 else half * 64 + 64)))
        cols = np.arange(f + 32 * k0, f + min(nc, 32 * k1))
        upd = W[np.ix_(r[ia], cols)] @ W[np.ix_(r[jb], cols)].T
        upd[ia[:, None] < jb[None, :]] = 0.0          # lower part only
        W[np.ix_(r[ia], r[jb])] -= upd

    nlev = len(plan["levels"])
    pending_lazy, pending_oz = {}, {}
    for L in range(nlev):
        lv = plan["levels"][L]
        g = lambda f: int(lv[LF[f]])
        for q in [q for q in pending_lazy if q <= L - 2]:
            for T in pending_lazy.pop(q):
                apply_tile(T)
        for q in [q for q in pending_oz if q <= L - 3]:
            tasks, tile = pending_oz.pop(q)
            for T in tasks:
                apply_oz(T, tile)
        for s in plan["small_list"][g("small_begin"):g("small_end")]:      # k_small_factor: factor + all its updates
            s = int(s)
            f, l = int(first[s]), int(first[s + 1])
            r = rows_of(s)
            below = r[l - f:]
            L11 = _signed_chol(W[f:l, f:l] + np.tril(W[f:l, f:l], -1).T, sg[f:l])
            W[f:l, f:l] = L11
            if len(below):                                   # X = A21 L11^-T S
                L21 = sla.solve_triangular(L11, W[np.ix_(below, np.arange(f, l))].T, lower=True).T * sg[f:l]
                W[np.ix_(below, np.arange(f, l))] = L21
                W[np.ix_(below, below)] -= np.tril((L21 * sg[f:l]) @ L21.T)
        for pc in plan["level_pieces"][g("piece_begin"):g("piece_end")]:   # k_diag_factor + k_trsm
            s, c0, c1, lvl = (int(x) for x in pieces[int(pc)])
            assert lvl == L
            f = int(first[s])
            r = rows_of(s)
            below = r[c1 - f:]
            blk = W[c0:c1, c0:c1]
            L11 = _signed_chol(blk + np.tril(blk, -1).T, sg[c0:c1])
            W[c0:c1, c0:c1] = L11
            if len(below):
                W[np.ix_(below, np.arange(c0, c1))] = sla.solve_triangular(L11, W[np.ix_(below, np.arange(c0, c1))].T, lower=True).T * sg[c0:c1]
        for T in plan["upd"][g("ext_begin"):g("ext_end")]:                 # urgent tiles (k_update)
            apply_tile(T)
        if g("lazy_end") > g("lazy_begin"):
            pending_lazy[L] = plan["upd128"][g("lazy_begin"):g("lazy_end")]
        if g("oz_end") > g("oz_begin"):
            pending_oz[L] = (plan["oz"][g("oz_begin"):g("oz_end")], g("oz_tile"))
    assert not pending_lazy or max(pending_lazy) >= nlev - 2                # whatever is left is joined at the end ...
    for q in sorted(pending_lazy):
        for T in pending_lazy[q]:
            apply_tile(T)
    for q in sorted(pending_oz):
        for T in pending_oz[q][0]:
            apply_oz(T, pending_oz[q][1])
    Lref = _signed_chol(Kp, sg)
    return W, Lref, k.stats()


@pytest.mark.parametrize("kwargs", [dict(ozaki_ncol=512), dict(ozaki_ncol=-1)], ids=["tcgen05+fp64", "fp64 only"])
def test_level_schedule_reproduces_the_cholesky_factor(kwargs):
    lp = lpgen.random_sparse(1800, 3600, 8, seed=424242, name="oz_medium")
    W, Lref, st = _emulate(lp.A, kwargs)
    assert (st["oz_tasks"] > 0) == (kwargs["ozaki_ncol"] > 0)
    assert np.abs(W - Lref).max() <= 1e-9 * np.abs(Lref).max()


def test_level_schedule_sparse_staircase_k1():
    """many small supernodes, ancestors updated through the segment lists (no dense root)"""
    lp = lpgen.banded_random(900, 1800, 4, 64, seed=99, name="band")
    W, Lref, _ = _emulate(lp.A, dict(ozaki_ncol=-1))
    assert np.abs(W - Lref).max() <= 1e-9 * np.abs(Lref).max()


def test_level_schedule_k2_signed_factor():
    """K2 (sqd.jl:24-55): quasi-definite augmented matrix, signed Cholesky L S L' with the x-block pivots negative"""
    lp = lpgen.config(3, mini=True)
    W, Lref, _ = _emulate(lp.A, dict(), system="K2")
    assert np.abs(W - Lref).max() <= 1e-9 * np.abs(Lref).max()
    lp = lpgen.banded_random(300, 600, 4, 48, seed=5, name="band2")
    W, Lref, _ = _emulate(lp.A, dict(), system="K2")
    assert np.abs(W - Lref).max() <= 1e-9 * np.abs(Lref).max()
