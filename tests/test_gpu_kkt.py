"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA path, called through the C ABI,
against the oracle on the same seeded inputs; bit-level structure checks; reference error
behaviour; end-to-end IPM parity; size-independent properties at BASELINE sizes.

Tolerances: BASELINE.json north_star asks 1e-8 relative on objectives/residuals; the reference's
own KKT test asks residuals <= sqrt(eps) (src/KKT/Test/test.jl:39-44)."""
import numpy as np
import pytest
import scipy.sparse as sp

import tlpb200_loader
from golden.lpex import KKT_CONFORMANCE, LPEX
from oracle import hsd_ref, kkt_ref

pkg = tlpb200_loader.load()
from tulip_jl_b200 import hsd, lpgen  # noqa: E402

pytestmark = pytest.mark.gpu
SQRT_EPS = float(np.sqrt(np.finfo(float).eps))
SYSTEMS = {"K1": pkg.K1, "K2": pkg.K2}


def _oracle(A, sysname):
    return kkt_ref.SparseK1(A) if sysname == "K1" else kkt_ref.SparseK2(A)


def _dense_from_lx(k, lx, xptr):
    """Rebuild the dense lower-triangular array stored in the supernodal panels."""
    sym = k.symbolic()
    N = len(sym["perm"])
    L = np.zeros((N, N))
    first, rp, rows = sym["sn_first"], sym["sn_rowptr"], sym["sn_rows"]
    for s in range(len(first) - 1):
        f, l = first[s], first[s + 1]
        r = rows[rp[s]:rp[s + 1]]
        P = lx[xptr[s]:xptr[s + 1]].reshape(l - f, len(r)).T      # column-major nrow x ncol
        for c in range(l - f):
            L[r[c:], f + c] = P[c:, c]
    return L, sym


def _kkt_matrix(A, sysname, theta, regP, regD):
    A = sp.csc_matrix(A)
    if sysname == "K1":
        return (A @ sp.diags(1.0 / (theta + regP)) @ A.T + sp.diags(regD)).toarray()
    return sp.bmat([[sp.diags(-(theta + regP)), A.T], [A, sp.diags(regD)]]).toarray()


@pytest.mark.parametrize("sysname", ["K1", "K2"])
def test_reference_conformance(sysname):
    """test/KKT/Cholmod/cholmod.jl:8-16 with the new backend: KKT.run_ls_tests(A, kkt)."""
    A = KKT_CONFORMANCE["A"]
    kkt = pkg.setup(sp.csc_matrix(A), SYSTEMS[sysname](), pkg.Backend())
    rp, rd, dx, dy = kkt_ref.run_ls_tests(A, kkt)
    assert rp <= SQRT_EPS and rd <= SQRT_EPS
    np.testing.assert_allclose(dx, KKT_CONFORMANCE["dx"], atol=1e-14)
    np.testing.assert_allclose(dy, KKT_CONFORMANCE["dy"], atol=1e-14)


@pytest.mark.parametrize("cfg", [2, 3, 4, 5, "T"])
@pytest.mark.parametrize("sysname", ["K1", "K2"])
def test_assemble_and_factor_parity(cfg, sysname):
    """assemble kernel == A D A' + Rd (spd.jl:43) / K2 matrix (sqd.jl:44-51) to 1e-14;
    numeric factor: |L S L' - P K P'| <= 1e-12 |L||L'| componentwise (backward-error bound)."""
    lp = lpgen.config(cfg, mini=True)
    m, n = lp.A.shape
    rng = np.random.default_rng(11)
    theta = np.exp(rng.uniform(-5, 5, n)); regP = np.full(n, 1e-7); regD = np.full(m, 1e-7)
    k = pkg.setup(lp.A, SYSTEMS[sysname](), pkg.Backend())
    Aeff = lp.A
    dc = k.dense_cols()
    if len(dc):      # K1 dense-column path: the sparse factor holds A_s D_s A_s' + Rd only
        keep = np.ones(n); keep[dc] = 0.0
        Aeff = (lp.A @ sp.diags(keep)).tocsc()
    K = _kkt_matrix(Aeff, sysname, theta, regP, regD)
    lx, xptr = k.debug_assembled(theta, regP, regD)
    Lasm, sym = _dense_from_lx(k, lx, xptr)
    p = sym["perm"]
    Kp = K[np.ix_(p, p)]
    np.testing.assert_allclose(Lasm, np.tril(Kp), rtol=1e-13, atol=1e-14 * np.abs(K).max())
    k.update(theta, regP, regD)
    lx, xptr = k.debug_lx()
    L, _ = _dense_from_lx(k, lx, xptr)
    sgn = np.ones(len(p)) if sysname == "K1" else np.where(p < n, -1.0, 1.0)
    R = (L * sgn[None, :]) @ L.T
    # componentwise backward-error bound of a Cholesky/LDL' without pivoting: |K - L S L'| <= c*N*u*|L||L'|
    bound = 1e-12 * (np.abs(L) @ np.abs(L).T) + 1e-300
    assert np.all(np.abs(R - Kp) <= bound), float((np.abs(R - Kp) / bound).max())


@pytest.mark.parametrize("cfg", [2, 3, 4, 5, "T"])
@pytest.mark.parametrize("sysname", ["K1", "K2"])
def test_solve_parity_vs_oracle(cfg, sysname):
    """per-call parity (SURVEY 8d protocol i): same (θ, regP, regD, ξp, ξd) -> same (dx, dy)."""
    lp = lpgen.config(cfg, mini=True)
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(3)
    k = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend())
    o = _oracle(A, sysname)
    for spread, reg in ((3.0, 1e-4), (8.0, 1e-6)):
        theta = np.exp(rng.uniform(-spread, spread, n))
        theta[rng.random(n) < 0.05] = 0.0                       # free variables: θinv_j = 0
        regP = np.full(n, reg); regD = np.full(m, reg)
        xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
        k.update(theta, regP, regD)
        o.update(theta, regP, regD)
        dx = np.zeros(n); dy = np.zeros(m); dx0 = np.zeros(n); dy0 = np.zeros(m)
        k.solve(dx, dy, xi_p, xi_d)
        o.solve(dx0, dy0, xi_p, xi_d)
        rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, dx, dy, xi_p, xi_d)
        rp0, rd0 = kkt_ref.kkt_residuals(A, theta, regP, regD, dx0, dy0, xi_p, xi_d)
        scale = max(1.0, np.abs(dx0).max(), np.abs(dy0).max())
        # device residuals no worse than 10x the oracle's (both are backward-stable factorizations)
        assert rp <= max(10 * rp0, SQRT_EPS * scale) and rd <= max(10 * rd0, SQRT_EPS * scale)
        ex = np.abs(dx - dx0).max() / max(np.abs(dx0).max(), 1e-300)
        ey = np.abs(dy - dy0).max() / max(np.abs(dy0).max(), 1e-300)
        # The bar is 1e-8 (north_star).  Two backward-stable factorisations of the same matrix can only agree to
        # ~cond(K) * u in the forward error, so a case whose condition number makes 1e-8 unreachable is held to
        # 10 * cond_2(K) * u instead (computed exactly on the dense matrix: these are the mini configs) and is listed by
        # name with its condition number in ILL_CONDITIONED_CASES (printed with -rP / on failure).  No floating yardstick.
        kappa = float(np.linalg.cond(_kkt_matrix(A, sysname, theta, regP, regD)))
        tol = 1e-8
        if 10 * kappa * np.finfo(float).eps > tol:
            tol = min(10 * kappa * np.finfo(float).eps, 1e-5)      # never looser than 1e-5, whatever the conditioning
            ILL_CONDITIONED_CASES[(str(cfg), sysname, spread, reg)] = (kappa, tol, max(ex, ey))
            print(f"ill-conditioned case cfg{cfg}-mini {sysname} spread={spread} reg={reg}: cond_2(K)={kappa:.3e}, "
                  f"bar {tol:.2e}, measured {max(ex, ey):.2e}")
        assert ex <= tol and ey <= tol, (ex, ey, kappa, tol)


# (config, system, theta spread, regularisation) -> (cond_2(K), bar used, measured error) for the per-call cases above
# whose conditioning makes a 1e-8 forward-error agreement between two correct factorisations impossible
ILL_CONDITIONED_CASES = {}


DENSE_SOLVE_CASES = [
    ("cfg2-mini", lambda: lpgen.config(2, mini=True), "K1", 1),
    ("cfg4-mini", lambda: lpgen.config(4, mini=True), "K1", 1),
    ("staircase-K2", lambda: lpgen.staircase(stages=8, nodes=150, arcs=260, name="st"), "K2", 1),
    ("random-700", lambda: lpgen.random_sparse(700, 1400, 6, name="r700"), "K1", 1),
    ("random-400-K2", lambda: lpgen.random_sparse(400, 800, 5, name="r400"), "K2", 200),
    ("block-angular", lambda: lpgen.block_angular(blocks=3, mb=300, nb=600, width=64, link=150, name="ba"), "K1", 130),
    ("random-2000", lambda: lpgen.random_sparse(2000, 4000, 8, name="r2000"), "K1", 0),   # default threshold: root only
]


@pytest.mark.parametrize("name,gen,sysname,ncol", DENSE_SOLVE_CASES, ids=[c[0] for c in DENSE_SOLVE_CASES])
def test_dense_solve_path_parity(name, gen, sysname, ncol):
    """The dense-solve path (repacked unit-block-diagonal tiles, flag-in-data hand-over; ragged edge blocks, rows
    below the columns, K2 signs) against the oracle and against the block path on the same factor."""
    lp = gen()
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(11)
    kd = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend(dense_solve_ncol=ncol))
    kb = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend(dense_solve_ncol=10 ** 9))
    assert len(kd.big_plan()["fwd"]) > 0 and len(kb.big_plan()["fwd"]) == 0
    o = _oracle(A, sysname)
    for rep in range(3):                                  # repeated sweeps: the exchange slots are reused with a new key
        theta = np.exp(rng.uniform(-2, 2, n)); regP = np.full(n, 1e-4); regD = np.full(m, 1e-4)
        xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
        for k in (kd, kb, o):
            k.update(theta, regP, regD)
        sols = []
        for k in (kd, kb, o):
            dx = np.zeros(n); dy = np.zeros(m)
            k.solve(dx, dy, xi_p, xi_d)
            if k is kd:                                   # same rhs again on the same factor: the exchange slots are reused
                dx2 = np.zeros(n); dy2 = np.zeros(m)        # (equal up to the rounding of the atomic accumulations)
                k.solve(dx2, dy2, xi_p, xi_d)
                assert np.abs(dx2 - dx).max() <= 1e-9 * np.abs(dx).max(), np.abs(dx2 - dx).max() / np.abs(dx).max()
                assert np.abs(dy2 - dy).max() <= 1e-9 * np.abs(dy).max(), np.abs(dy2 - dy).max() / np.abs(dy).max()
            sols.append(np.concatenate([dx, dy]))
        ref = np.abs(sols[2]).max()
        assert np.abs(sols[0] - sols[2]).max() / ref < 1e-8, np.abs(sols[0] - sols[2]).max() / ref
        assert np.abs(sols[0] - sols[1]).max() / ref < 1e-9, np.abs(sols[0] - sols[1]).max() / ref
        rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, sols[0][:n], sols[0][n:], xi_p, xi_d)
        assert rp <= SQRT_EPS * max(1.0, ref) and rd <= SQRT_EPS * max(1.0, ref)


def test_errors_match_reference():
    A = lpgen.config(2, mini=True).A
    m, n = A.shape
    for sysname in ("K1", "K2"):
        k = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend())
        with pytest.raises(pkg.DimensionMismatch):                 # spd.jl:26-34
            k.update(np.ones(n - 1), np.ones(n), np.ones(m))
        with pytest.raises(pkg.DimensionMismatch):
            k.update(np.ones(n), np.ones(n), np.ones(m + 2))
        with pytest.raises(pkg.PosDefException):                   # spd.jl:47 / ldlfact.jl:112-117
            k.update(np.ones(n), np.ones(n), -1e3 * np.ones(m))
        with pytest.raises(pkg.PosDefException):
            k.update(np.full(n, np.nan), np.ones(n), np.ones(m))
        # the solver stays usable after a failed update (regularisation bump path, step.jl:34-51)
        k.update(np.ones(n), np.ones(n), np.ones(m))
        dx = np.zeros(n); dy = np.zeros(m)
        k.solve(dx, dy, np.ones(m), np.ones(n))
        rp, rd = kkt_ref.kkt_residuals(A, np.ones(n), np.ones(n), np.ones(m), dx, dy, np.ones(m), np.ones(n))
        assert rp <= SQRT_EPS and rd <= SQRT_EPS
    k2 = pkg.setup(A, pkg.K2(), pkg.Backend())
    with pytest.raises(pkg.PosDefException):                       # wrong sign in the (1,1) block
        k2.update(-5 * np.ones(n), np.ones(n), np.ones(m))


@pytest.mark.parametrize("name", list(LPEX))
@pytest.mark.parametrize("sysname", ["K1", "K2"])
def test_example_lps_end_to_end(name, sysname):
    """test/examples.jl with KKT_Backend = TlpB200: statuses and values the reference asserts."""
    lp = LPEX[name]
    dat = hsd_ref.standard_form(**{k: v for k, v in lp.items() if k != "expect"})
    kkt = pkg.setup(dat.A, SYSTEMS[sysname](), pkg.Backend())
    h = hsd.HSD(dat.A, dat.b, dat.c, dat.l, dat.u, kkt, c0=dat.c0, objsense=dat.objsense)
    status = h.optimize()
    exp = lp["expect"]
    tol = 100 * SQRT_EPS
    assert status == exp["status"]
    if "obj" in exp:
        assert abs(h.primal_objective - exp["obj"]) <= tol * (1 + abs(exp["obj"]))
    if "x" in exp:
        np.testing.assert_allclose(h.x[:lp["nvar"]] / h.tau, exp["x"], atol=tol, rtol=tol)
    if "y" in exp:
        np.testing.assert_allclose(h.y / h.tau, exp["y"], atol=tol, rtol=tol)


@pytest.mark.parametrize("cfg,sysname", [(2, "K1"), (3, "K2"), (4, "K1"), (5, "K2"), ("T", "K1")])
def test_ipm_end_to_end_parity(cfg, sysname):
    """SURVEY 8d protocol ii: the restated HSD run with the oracle KKT and with the device KKT,
    tolerances tightened to 1e-10, objectives and residual norms agree to 1e-8 relative."""
    lp = lpgen.config(cfg, mini=True)
    P = dict(TolerancePFeas=1e-10, ToleranceDFeas=1e-10, ToleranceRGap=1e-10)
    dat = hsd_ref.IPMData(lp.A, lp.b, True, lp.c, 0.0, lp.l, lp.u)
    ref = hsd_ref.HSDRef(dat, _oracle(lp.A, sysname), hsd_ref.IPMOptions(**P))
    ref.optimize()
    kkt = pkg.setup(lp.A, SYSTEMS[sysname](), pkg.Backend())
    h = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, kkt, params=hsd.IPMOptions(**P))
    h.optimize()
    assert h.status == ref.status == "Trm_Optimal"
    assert abs(h.niter - ref.niter) <= 1
    for a, b in ((h.primal_objective, ref.primal_objective), (h.dual_objective, ref.dual_objective)):
        assert abs(a - b) <= 1e-8 * (1 + abs(b))
    assert h.rp_nrm <= 1e-8 * (1 + np.abs(lp.b).max()) and h.rd_nrm <= 1e-8 * (1 + np.abs(lp.c).max())


def test_multi_rhs_and_device_pointer_api():
    torch = pytest.importorskip("torch")
    lp = lpgen.config(2, mini=True)
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(9)
    theta = np.exp(rng.uniform(-4, 4, n)); regP = np.full(n, 1e-6); regD = np.full(m, 1e-6)
    for sysname in ("K1", "K2"):
        k = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend())
        k.update(theta, regP, regD)
        XP = rng.standard_normal((3, m)); XD = rng.standard_normal((3, n))
        DX = np.zeros((3, n)); DY = np.zeros((3, m))
        k.solve_multi(DX, DY, XP, XD)
        for r in range(3):
            dx = np.zeros(n); dy = np.zeros(m)
            k.solve(dx, dy, XP[r], XD[r])
            # not bitwise: the forward sweep reduces into ancestors with floating-point atomics
            np.testing.assert_allclose(dx, DX[r], rtol=1e-7, atol=1e-10)
            np.testing.assert_allclose(dy, DY[r], rtol=1e-7, atol=1e-10)
        # device-resident inputs on torch's current stream
        dev = torch.device("cuda:0")
        k.set_stream(torch.cuda.current_stream().cuda_stream)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        k.update_dev(t(theta), t(regP), t(regD))
        k.update_status()
        ddx = torch.zeros(n, dtype=torch.float64, device=dev); ddy = torch.zeros(m, dtype=torch.float64, device=dev)
        k.solve_dev(ddx, ddy, t(XP[0]), t(XD[0]))
        torch.cuda.synchronize()
        # a second factorisation of the same data: update tiles reach the ancestors through RED.ADD.F64 (and the critical tiles
        # are split over K across CTAs), so two factors differ by summation order -- O(kappa u) in the solution, not bitwise
        np.testing.assert_allclose(ddx.cpu().numpy(), DX[0], rtol=1e-7, atol=1e-10)
        np.testing.assert_allclose(ddy.cpu().numpy(), DY[0], rtol=1e-7, atol=1e-10)


def test_graph_and_plain_launch_agree():
    lp = lpgen.config(4, mini=True)
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(2)
    theta = np.exp(rng.uniform(-4, 4, n)); regP = np.full(n, 1e-6); regD = np.full(m, 1e-6)
    xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
    out = []
    for graph in (True, False):
        k = pkg.setup(A, pkg.K1(), pkg.Backend(use_graph=graph))
        k.update(theta, regP, regD)
        dx = np.zeros(n); dy = np.zeros(m)
        k.solve(dx, dy, xi_p, xi_d)
        out.append(np.concatenate([dx, dy]))
        st = k.stats()
        assert st["launches_update"] > 0 and st["launches_solve"] > 0
    # two factorisations (atomic accumulation order differs from run to run): equal to O(kappa u), theta spans e^-4 .. e^4
    np.testing.assert_allclose(out[0], out[1], rtol=1e-7, atol=1e-10)


def test_empty_and_ragged_inputs():
    """empty columns, a column of explicit zeros, a single row."""
    rng = np.random.default_rng(4)
    A = sp.csc_matrix((np.array([1.0, 2.0, 0.0, -1.0]), (np.array([0, 2, 1, 1]), np.array([1, 3, 4, 5]))), shape=(3, 7))
    A = (A + sp.eye(3, 7)).tocsc()
    for sysname in ("K1", "K2"):
        k = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend())
        m, n = A.shape
        th = np.exp(rng.uniform(-2, 2, n)); rP = np.full(n, 1e-3); rD = np.full(m, 1e-3)
        k.update(th, rP, rD)
        dx = np.zeros(n); dy = np.zeros(m); xp = rng.standard_normal(m); xd = rng.standard_normal(n)
        k.solve(dx, dy, xp, xd)
        rp, rd = kkt_ref.kkt_residuals(A, th, rP, rD, dx, dy, xp, xd)
        assert rp <= 1e-10 and rd <= 1e-10


@pytest.mark.parametrize("cfg,sysname", [(2, "K1"), (3, "K2")])
def test_full_size_properties(cfg, sysname):
    """BASELINE sizes: KKT residual (test.jl:39-40 formulas) and linearity of solve! in the rhs."""
    lp = lpgen.config(cfg)
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(8)
    theta = np.exp(rng.uniform(-3, 3, n)); regP = np.full(n, 1e-6); regD = np.full(m, 1e-6)
    k = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend())
    k.update(theta, regP, regD)
    xs = []
    rhs = [(rng.standard_normal(m), rng.standard_normal(n)) for _ in range(2)]
    rhs.append((2.0 * rhs[0][0] - 3.0 * rhs[1][0], 2.0 * rhs[0][1] - 3.0 * rhs[1][1]))
    for xp, xd in rhs:
        dx = np.zeros(n); dy = np.zeros(m)
        k.solve(dx, dy, xp, xd)
        rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, dx, dy, xp, xd)
        scale = max(1.0, np.abs(dx).max(), np.abs(dy).max())
        assert rp <= SQRT_EPS * scale and rd <= SQRT_EPS * scale
        xs.append(np.concatenate([dx, dy]))
    comb = 2.0 * xs[0] - 3.0 * xs[1]
    assert np.abs(xs[2] - comb).max() <= 1e-7 * max(1.0, np.abs(comb).max())


def test_dense_column_schur_path():
    """BASELINE config 5: K1 with dense columns handled by the low-rank Schur correction == oracle on the FULL
    matrix (the reference itself would form a dense A*D*A', spd.jl:43) == the same backend with the path disabled."""
    lp = lpgen.config(5, mini=True)
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(12)
    theta = np.exp(rng.uniform(-3, 3, n)); regP = np.full(n, 1e-5); regD = np.full(m, 1e-5)
    xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
    k = pkg.setup(A, pkg.K1(), pkg.Backend())
    assert len(k.dense_cols()) == 3
    k_off = pkg.setup(A, pkg.K1(), pkg.Backend(dense_col_threshold=-1))
    assert len(k_off.dense_cols()) == 0
    assert k.stats()["nnzL"] < 0.5 * k_off.stats()["nnzL"]          # the point of the exercise: far less fill
    o = kkt_ref.DenseK1(A)
    sols = []
    for kk in (k, k_off, o):
        kk.update(theta, regP, regD)
        dx = np.zeros(n); dy = np.zeros(m)
        kk.solve(dx, dy, xi_p, xi_d)
        rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, dx, dy, xi_p, xi_d)
        assert rp <= 1e-9 and rd <= 1e-9
        sols.append(np.concatenate([dx, dy]))
    for s_ in sols[:2]:
        assert np.abs(s_ - sols[2]).max() <= 1e-8 * np.abs(sols[2]).max()
    # end to end: the IPM converges to the same objective with and without the path
    objs = []
    for be in (pkg.Backend(), pkg.Backend(dense_col_threshold=-1)):
        kkt = pkg.setup(A, pkg.K1(), be)
        h = hsd.HSD(A, lp.b, lp.c, lp.l, lp.u, kkt)
        assert h.optimize() == "Trm_Optimal"
        objs.append(h.primal_objective)
    assert abs(objs[0] - objs[1]) <= 1e-7 * (1 + abs(objs[1]))


def test_dense_columns_full_size():
    """config 5 at BASELINE size (m=5e4, n=1e5, 8 columns of 25 000 non-zeros): KKT residuals of the full system."""
    lp = lpgen.config(5)
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(13)
    theta = np.exp(rng.uniform(-3, 3, n)); regP = np.full(n, 1e-6); regD = np.full(m, 1e-6)
    k = pkg.setup(A, pkg.K1(), pkg.Backend())
    assert len(k.dense_cols()) == 8
    k.update(theta, regP, regD)
    xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
    dx = np.zeros(n); dy = np.zeros(m)
    k.solve(dx, dy, xi_p, xi_d)
    rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, dx, dy, xi_p, xi_d)
    scale = max(1.0, np.abs(dx).max(), np.abs(dy).max())
    assert rp <= SQRT_EPS * scale and rd <= SQRT_EPS * scale


# ---- parity at BASELINE sizes against the independent CPU floating point (oracle/cpu_kkt.py) -----------------------
def _ipm_like_inputs(lp, which, rng):
    """(theta_inv, regP, regD): 'start' = what the first update! of the reference sees (HSD.jl:238-247 start point,
    step.jl:24-31: theta_inv in {0,1,2}, regP = regD = 0.1), 'mid' = a mid-run iterate (theta spread e^+-3, reg 1e-5)."""
    m, n = lp.A.shape
    if which == "start":
        theta = np.isfinite(lp.l).astype(float) + np.isfinite(lp.u).astype(float)
        return theta, np.full(n, 0.1), np.full(m, 0.1)
    return np.exp(rng.uniform(-3, 3, n)), np.full(n, 1e-5), np.full(m, 1e-5)


@pytest.mark.parametrize("cfg,sysname", [(2, "K1"), (3, "K2"), (4, "K1")])
@pytest.mark.parametrize("which", ["start", "mid"])
def test_full_size_parity_vs_cpu_port(cfg, sysname, which):
    """update!/solve! at BASELINE size (tcgen05 path on where the plan uses it) against oracle/cpu_kkt.CpuSupernodalKKT --
    independent floating point (OpenBLAS left-looking supernodal factorisation) -- to 1e-8 relative, same inputs."""
    import os
    from oracle import cpu_kkt
    lp = lpgen.config(cfg)
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(21)
    theta, regP, regD = _ipm_like_inputs(lp, which, rng)
    k = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend())
    if cfg == 2:
        assert k.stats()["oz_tasks"] > 0, "cfg2 is expected to run its root supernode on the tcgen05 path"
    an = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend(analyze_only=True))
    ck = cpu_kkt.CpuSupernodalKKT(A, sysname, nthreads=os.cpu_count(), symbolic_from=an)
    k.update(theta, regP, regD)
    ck.update(theta, regP, regD)
    for _ in range(2):
        xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
        dx = np.zeros(n); dy = np.zeros(m); dx0 = np.zeros(n); dy0 = np.zeros(m)
        k.solve(dx, dy, xi_p, xi_d)
        ck.solve(dx0, dy0, xi_p, xi_d)
        ex = np.abs(dx - dx0).max() / np.abs(dx0).max()
        ey = np.abs(dy - dy0).max() / np.abs(dy0).max()
        assert ex <= 1e-8 and ey <= 1e-8, (cfg, sysname, which, ex, ey)
        rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, dx, dy, xi_p, xi_d)
        scale = max(1.0, np.abs(dx).max(), np.abs(dy).max())
        assert rp <= SQRT_EPS * scale and rd <= SQRT_EPS * scale


def test_full_size_parity_dense_columns_vs_cpu_k2():
    """config 5 at BASELINE size: the K1 dense-column Schur path on the device against the CPU port's K2 factorisation of
    the SAME linear system ("parity vs oracle K2 on the same LP"; a CPU K1 would form a dense A D A')."""
    import os
    from oracle import cpu_kkt
    lp = lpgen.config(5)
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(22)
    k = pkg.setup(A, pkg.K1(), pkg.Backend())
    assert len(k.dense_cols()) == 8
    an = pkg.setup(A, pkg.K2(), pkg.Backend(analyze_only=True))
    ck = cpu_kkt.CpuSupernodalKKT(A, "K2", nthreads=os.cpu_count(), symbolic_from=an)
    for which in ("start", "mid"):
        theta, regP, regD = _ipm_like_inputs(lp, which, rng)
        k.update(theta, regP, regD)
        ck.update(theta, regP, regD)
        xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
        dx = np.zeros(n); dy = np.zeros(m); dx0 = np.zeros(n); dy0 = np.zeros(m)
        k.solve(dx, dy, xi_p, xi_d)
        ck.solve(dx0, dy0, xi_p, xi_d)
        ex = np.abs(dx - dx0).max() / np.abs(dx0).max()
        ey = np.abs(dy - dy0).max() / np.abs(dy0).max()
        assert ex <= 1e-8 and ey <= 1e-8, (which, ex, ey)


def test_solve_timeout_flag_is_reported_and_cleared():
    """a sweep-kernel hand-over time-out must surface from solve! itself.  The flag is raised artificially here
    (tlpb200_debug_raise_timeout): the next solve! returns TLPB200_INTERNAL, the one after works again."""
    lp = lpgen.config(2, mini=True)
    A = lp.A
    m, n = A.shape
    k = pkg.setup(A, pkg.K1(), pkg.Backend())
    k.update(np.ones(n), np.ones(n), np.ones(m))
    dx = np.zeros(n); dy = np.zeros(m)
    k.solve(dx, dy, np.ones(m), np.ones(n))
    k.debug_raise_timeout()
    with pytest.raises(pkg.TlpB200Error):
        k.solve(dx, dy, np.ones(m), np.ones(n))
    k.solve(dx, dy, np.ones(m), np.ones(n))
    rp, rd = kkt_ref.kkt_residuals(A, np.ones(n), np.ones(n), np.ones(m), dx, dy, np.ones(m), np.ones(n))
    assert rp <= SQRT_EPS and rd <= SQRT_EPS


@pytest.mark.parametrize("sysname", ["K1", "K2"])
def test_iterative_refinement_option(sysname):
    """refine_steps > 0 re-solves on the FP64 residual of the factored KKT system; on ill-conditioned data the
    KKT residuals (test.jl:39-40 formulas) must not get worse and the solution must stay within the per-call parity bar."""
    lp = lpgen.config(3 if sysname == "K2" else 2, mini=True)
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(31)
    theta = np.exp(rng.uniform(-7, 7, n)); regP = np.full(n, 1e-7); regD = np.full(m, 1e-7)
    xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
    res = []
    sols = []
    for steps in (0, 2):
        k = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend(refine_steps=steps))
        k.update(theta, regP, regD)
        dx = np.zeros(n); dy = np.zeros(m)
        k.solve(dx, dy, xi_p, xi_d)
        res.append(max(kkt_ref.kkt_residuals(A, theta, regP, regD, dx, dy, xi_p, xi_d)))
        sols.append(np.concatenate([dx, dy]))
        if steps:
            assert k.stats()["launches_solve"] > launches0      # the refinement sweeps were really enqueued
        launches0 = k.stats()["launches_solve"]
    scale = max(1.0, np.abs(sols[0]).max())
    assert res[1] <= max(2.0 * res[0], 1e-12 * scale), res
    assert res[1] <= SQRT_EPS * scale
    assert np.abs(sols[1] - sols[0]).max() <= 1e-5 * np.abs(sols[0]).max()


def test_dense_columns_full_size_ipm_converges():
    """config 5 at BASELINE size through the whole IPM.  Round-2 regression: with the explicit Sherman-Morrison-Woodbury
    formula the solves lost all accuracy once the dense columns became basic (K_s singular up to the regularisation) and the
    IPM stalled at pfeas ~ 5 until the iteration limit; the factorised Schur form converges like the CPU port's K2
    (`python bench.py --impl reference --config 5`: Trm_Optimal after 17 iterations, objective 50833.48157246)."""
    lp = lpgen.config(5)
    k = pkg.setup(lp.A, pkg.K1(), pkg.Backend())
    assert len(k.dense_cols()) == 8
    h = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, k)
    assert h.optimize() == "Trm_Optimal"
    assert h.niter <= 20
    assert abs(h.primal_objective - 50833.48157246) <= 1e-7 * 50833.48157246
    assert abs(h.primal_objective - h.dual_objective) <= 1e-8 * (1 + abs(h.dual_objective))


@pytest.mark.parametrize("cfg,sysname", [(3, "K2"), (4, "K1"), (2, "K1")])
def test_merged_level_sweeps_match_per_level_sweeps(cfg, sysname, monkeypatch):
    """Round 2: the block-solve items of consecutive levels share one persistent launch and synchronise through per-supernode
    counters (Plan::SolveOp).  At BASELINE size (config 3: a chain of ~95 levels) the merged sweeps must give the solutions
    of the per-level launches (TLPB200_MERGE_LEVELS=0) up to the rounding of the atomic accumulations, for many right-hand
    sides in a row on the same factor (the counters are reset per solve), and use fewer launches."""
    lp = lpgen.config(cfg)
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(41)
    theta = np.exp(rng.uniform(-3, 3, n)); regP = np.full(n, 1e-5); regD = np.full(m, 1e-5)
    monkeypatch.setenv("TLPB200_MERGE_LEVELS", "0")
    k0 = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend())
    monkeypatch.setenv("TLPB200_MERGE_LEVELS", "1")
    k1 = pkg.setup(A, SYSTEMS[sysname](), pkg.Backend())
    k0.update(theta, regP, regD); k1.update(theta, regP, regD)
    worst = 0.0
    for rep in range(12):
        xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
        a = [np.zeros(n), np.zeros(m)]; b = [np.zeros(n), np.zeros(m)]
        k0.solve(a[0], a[1], xi_p, xi_d)
        k1.solve(b[0], b[1], xi_p, xi_d)
        ref = max(np.abs(a[0]).max(), np.abs(a[1]).max())
        worst = max(worst, np.abs(a[0] - b[0]).max() / ref, np.abs(a[1] - b[1]).max() / ref)
        rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, b[0], b[1], xi_p, xi_d)
        assert rp <= SQRT_EPS * max(1.0, ref) and rd <= SQRT_EPS * max(1.0, ref)
    assert worst <= 1e-9, worst
    assert k1.stats()["launches_solve"] < k0.stats()["launches_solve"]
