"""CPU check of the dense-solve plan (plan.cpp: BigPack / BigTask lists) by emulating, in NumPy, exactly what
k_pack_big / k_fwd_big / k_bwd_big (csrc/kernels_dense_solve.cu) do with it: tile numbering, the order the
tiles of a task are streamed in, the exchange-slot indices, ragged edge blocks and the unit-block-diagonal
algebra  L u = b  <=>  Lhat w = b,  x = Lhat^{-T} (L_kk^{-T} S L_kk^{-1} w).  The panels come from the CPU
port of the factorisation (oracle/cpu_supernodal.c); the answer is compared with its own triangular solves,
which stand for  F \\ xi  (spd.jl:61, sqd.jl:66)."""
import numpy as np
import pytest
import scipy.linalg as sla

import tlpb200_loader
from oracle import cpu_kkt

pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen  # noqa: E402

B = 128


def _emulate(cpu, plan, rhs_perm):
    first, rp, rows, xptr = cpu.sn_first, cpu.sn_rowptr, cpu.sn_rows, cpu.sn_xptr
    ns = cpu.nsuper
    sign = cpu.sign.astype(float)
    big = set(int(s) for s in plan["fwd"]["sn"])
    panel = {}
    for s in range(ns):
        nc = first[s + 1] - first[s]
        nr = rp[s + 1] - rp[s]
        panel[s] = cpu.Lx[xptr[s]:xptr[s + 1]].reshape(nc, nr).T          # nrow x ncol (column-major storage)
    # ---- k_pack_big -------------------------------------------------------------------------
    Ft = np.full((plan["n_ftiles"], B * B), np.nan)
    Bt = np.full((plan["n_btiles"], B * B), np.nan)
    dinv = {}
    for s in big:
        nc = first[s + 1] - first[s]
        for k in range((nc + B - 1) // B):
            nb = min(B, nc - k * B)
            X = np.zeros((B, B))
            X[:nb, :nb] = np.linalg.inv(np.tril(panel[s][k * B:k * B + nb, k * B:k * B + nb]))
            dinv[(s, k)] = X
    for p in plan["pack"]:
        s, r0, nr, j = int(p["sn"]), int(p["r0"]), int(p["nr"]), int(p["j"])
        nc = first[s + 1] - first[s]
        nbj = min(B, nc - j * B)
        T = np.zeros((B, B))
        T[:nr, :nbj] = panel[s][r0:r0 + nr, j * B:j * B + nbj]
        That = T @ dinv[(s, j)]
        assert np.isnan(Ft[p["fdst"]][0]), "tile written twice"
        Ft[p["fdst"]] = That.T.ravel()      # [c*128 + r]
        if p["bdst"] >= 0:
            assert r0 < nc and np.isnan(Bt[p["bdst"]][0]), "tile written twice"
            Bt[p["bdst"]] = That.ravel()    # [r*128 + c]
        else:
            assert r0 >= nc                 # rows below the columns: no backward tile (k_bwd_below)
    assert not np.isnan(Ft).any() and not np.isnan(Bt).any(), "tile never written"
    tile = lambda buf, i: buf[i].reshape(B, B).T   # M[r, c] = flat[c*128 + r]  (the kernels' load_tile pattern)

    wk = rhs_perm.copy()
    xq = np.full(plan["xq_slots"], np.nan)
    # ---- forward ----------------------------------------------------------------------------
    ftasks = {s: [t for t in plan["fwd"] if t["sn"] == s] for s in big}
    for s in range(ns):
        f, l = first[s], first[s + 1]
        nc = l - f
        r = rows[rp[s]:rp[s + 1]]
        P = panel[s]
        if s not in big:
            u = sla.solve_triangular(np.tril(P[:nc, :nc]), wk[f:l], lower=True)
            wk[f:l] = u
            wk[r[nc:]] -= P[nc:, :] @ u
            continue
        for t in ftasks[s]:
            acc = np.zeros(B)
            for j in range(t["ntile"]):
                x = xq[t["xq0"] + j * B: t["xq0"] + (j + 1) * B]
                assert not np.isnan(x).any(), "forward task reads a block that was not published yet"
                acc -= tile(Ft, t["tile0"] + j) @ x
            nr, r0 = int(t["nr"]), int(t["r0"])
            if t["kind"] == 0:
                assert t["ntile"] == t["blk"] and r0 == t["blk"] * B
                w = acc.copy()
                w[:nr] += wk[f + r0:f + r0 + nr]
                assert np.all(w[nr:] == 0.0)
                xq[t["xq0"] + t["blk"] * B: t["xq0"] + (t["blk"] + 1) * B] = w
                wk[f + r0:f + r0 + nr] = w[:nr]
            else:
                wk[r[r0:r0 + nr]] += acc[:nr]
    # ---- backward ---------------------------------------------------------------------------
    xq[:] = np.nan
    btasks = {s: [t for t in plan["bwd"] if t["sn"] == s] for s in big}
    for s in range(ns - 1, -1, -1):
        f, l = first[s], first[s + 1]
        nc = l - f
        r = rows[rp[s]:rp[s + 1]]
        P = panel[s]
        if s not in big:
            t_ = sign[f:l] * wk[f:l] - P[nc:, :].T @ wk[r[nc:]]
            wk[f:l] = sla.solve_triangular(np.tril(P[:nc, :nc]), t_, lower=True, trans="T")
            continue
        ncb = (nc + B - 1) // B
        for t in btasks[s]:
            k, nr = int(t["blk"]), int(t["nr"])
            w = np.zeros(B)
            w[:nr] = wk[f + k * B:f + k * B + nr]
            y = dinv[(s, k)] @ w
            y[:nr] *= sign[f + k * B:f + k * B + nr]
            # k_bwd_below: rows below the columns on the unscaled panel (every big supernode with such rows is split)
            y[:nr] -= P[nc:, k * B:k * B + nr].T @ wk[r[nc:]]
            z = dinv[(s, k)].T @ y
            acc = np.zeros(B)
            assert t["ntile"] == ncb - 1 - k and t["nbelow"] == 0
            for j in range(t["ntile"]):
                cb = ncb - 1 - j
                x = xq[t["xq0"] + cb * B: t["xq0"] + (cb + 1) * B]
                assert not np.isnan(x).any(), "backward task reads a block that was not published yet"
                acc -= tile(Bt, t["tile0"] + j) @ x
            xk = z + acc
            xq[t["xq0"] + k * B: t["xq0"] + (k + 1) * B] = xk
            wk[f + k * B:f + k * B + nr] = xk[:nr]
    return wk


CASES = [("cfg2-mini", lambda: lpgen.config(2, mini=True), "K1", 1),
         ("cfg4-mini", lambda: lpgen.config(4, mini=True), "K1", 1),
         ("staircase-K2", lambda: lpgen.staircase(stages=8, nodes=150, arcs=260, name="st"), "K2", 1),
         ("random-700", lambda: lpgen.random_sparse(700, 1400, 6, name="r700"), "K1", 1),
         ("random-700-K2", lambda: lpgen.random_sparse(400, 800, 5, name="r400"), "K2", 200),
         ("cfg4-small", lambda: lpgen.block_angular(blocks=3, mb=300, nb=600, width=64, link=150, name="ba"), "K1", 130)]


@pytest.mark.parametrize("name,gen,sysname,ncol", CASES, ids=[c[0] for c in CASES])
def test_dense_solve_plan_emulation(name, gen, sysname, ncol):
    lp = gen()
    A = lp.A
    m, n = A.shape
    sy = pkg.K1() if sysname == "K1" else pkg.K2()
    k = pkg.setup(A, sy, pkg.Backend(analyze_only=True, dense_solve_ncol=ncol))
    plan = k.big_plan()
    assert len(plan["fwd"]) > 0, "case has no big supernode: adjust the generator"
    cpu = cpu_kkt.CpuSupernodalKKT(A, sysname, nthreads=1, symbolic_from=k)
    rng = np.random.default_rng(5)
    theta = np.exp(rng.uniform(-2, 2, n)); regP = np.full(n, 1e-4); regD = np.full(m, 1e-4)
    cpu.update(theta, regP, regD)
    rhs = rng.standard_normal(cpu.N)
    ref = cpu._fsolve(rhs)[cpu.perm]            # permuted solution of the CPU port's own sweeps
    got = _emulate(cpu, plan, rhs[cpu.perm])
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 1e-9, err


def test_default_threshold_keeps_small_problems_off_the_dense_path():
    lp = lpgen.config(2, mini=True)
    k = pkg.setup(lp.A, pkg.K1(), pkg.Backend(analyze_only=True))
    plan = k.big_plan()
    assert len(plan["fwd"]) == 0 and plan["n_ftiles"] == 0
