"""CPU tests: C-ABI library loads and exports every declared symbol; host symbolic analysis is
bit-exact against the brute-force recomputation in oracle/symbolic_ref.py; the product refuses
to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import tlpb200_loader
from oracle import symbolic_ref as sr

pkg = tlpb200_loader.load()
from tulip_jl_b200 import _lib, lpgen  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "tlpb200.h")).read()
    declared = sorted(set(re.findall(r"\b(tlpb200_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tlpb200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared


def test_binding_struct_layouts_match_the_library():
    """the ctypes mirrors (and, by the same field lists, the Julia structs of tulip.jl_b200/julia/TlpB200.jl) must have the
    sizes the compiled library uses for tlpb200_options / tlpb200_stats"""
    import ctypes as C
    lib = _lib.load()
    out = (C.c_int32 * 3)()
    lib.tlpb200_abi_sizes(out)
    assert out[0] == C.sizeof(_lib.Options) == 16 * 4
    assert out[1] == C.sizeof(_lib.Stats)
    assert out[2] == 24 and len(_lib.KERNEL_CLASSES) <= out[2]
    jl = open(os.path.join(ROOT, "tulip.jl_b200", "julia", "TlpB200.jl")).read()
    body = jl[jl.index("struct COptions"):jl.index("end", jl.index("struct COptions"))]
    n32 = len(re.findall(r"::Int32", body)) + sum(int(x) for x in re.findall(r"NTuple\{(\d+),Int32\}", body))
    assert n32 * 4 == out[0], "Julia COptions does not match sizeof(tlpb200_options)"


def test_backend_strings():
    assert "TlpB200" in pkg.backend(None)


def _check(A, sysname):
    sy = pkg.K1() if sysname == "K1" else pkg.K2()
    k = pkg.setup(A, sy, pkg.Backend(analyze_only=True))
    st, sym = k.stats(), k.symbolic()
    N = st["order"]
    assert N == (A.shape[0] if sysname == "K1" else sum(A.shape))
    assert sorted(sym["perm"].tolist()) == list(range(N))
    dc = k.dense_cols()                      # K1: columns kept out of the sparse factor (low-rank Schur path)
    assert sysname == "K1" or len(dc) == 0
    Aeff = sp.csc_matrix(A)
    if len(dc):
        keep = np.ones(A.shape[1]); keep[dc] = 0.0
        Aeff = (Aeff @ sp.diags(keep)).tocsc(); Aeff.eliminate_zeros()
        nnzcol = np.diff(sp.csc_matrix(A).indptr)
        assert nnzcol[dc].min() > max(32, 0.05 * A.shape[0]) and len(dc) <= 64
    S = sr.kkt_pattern(Aeff, sysname)
    parent, cc, struct = sr.symbolic_bruteforce(S, sym["perm"])
    assert np.array_equal(parent, sym["parent"]), "elimination tree differs"
    assert np.array_equal(cc, sym["colcount"]), "column counts differ"
    assert sr.is_postordered(sym["parent"])
    assert st["nnzL"] == int(cc.sum())
    assert st["flops"] == float((cc.astype(float) ** 2).sum())
    first = sym["sn_first"]
    assert first[0] == 0 and first[-1] == N and np.all(np.diff(first) > 0)
    rows = sr.supernode_rows_bruteforce(struct, first)
    for s in range(st["nsuper"]):
        mine = sym["sn_rows"][sym["sn_rowptr"][s]:sym["sn_rowptr"][s + 1]]
        assert np.array_equal(mine, rows[s]), f"row structure of supernode {s} differs"
    assert pkg.linear_system(k).endswith("(K1)" if sysname == "K1" else "(K2)")
    return k, st


@pytest.mark.parametrize("cfg", [2, 3, 4, 5, "T"])
@pytest.mark.parametrize("sysname", ["K1", "K2"])
def test_symbolic_bit_exact_mini_configs(cfg, sysname):
    _check(lpgen.config(cfg, mini=True).A, sysname)


@pytest.mark.parametrize("sysname", ["K1", "K2"])
def test_symbolic_edge_cases(sysname):
    # reference conformance matrix (test/KKT/Cholmod/cholmod.jl:3-6)
    _check(sp.csc_matrix(np.array([[1.0, 0, 1, 0], [0, 1, 0, 1]])), sysname)
    # single row / single column / empty columns / duplicate-free ragged
    _check(sp.csc_matrix(np.array([[1.0, 2.0, 0.0, 3.0]])), sysname)
    _check(sp.csc_matrix(np.array([[1.0], [0.0], [2.0]])), sysname)
    _check(sp.csc_matrix((np.array([1.0, 2.0]), (np.array([0, 2]), np.array([1, 3]))), shape=(4, 6)), sysname)
    # one dense column + identity (arrow), and one dense row
    A = sp.hstack([sp.identity(30), sp.csc_matrix(np.ones((30, 1)))]).tocsc()
    _check(A, sysname)
    B = sp.vstack([sp.identity(30), sp.csc_matrix(np.ones((1, 30)))]).tocsc()
    _check(B, sysname)


def test_ordering_reduces_fill_on_grid():
    """AMD quality: 40x40 grid Laplacian (as A = node-arc incidence); fill must be far below the
    natural ordering's and within 25% of SuperLU's MMD (independent implementation)."""
    import scipy.sparse.linalg as spla
    k = 40
    idx = np.arange(k * k).reshape(k, k)
    rows, cols, vals, e = [], [], [], 0
    for a, b in ((idx[:, :-1].ravel(), idx[:, 1:].ravel()), (idx[:-1, :].ravel(), idx[1:, :].ravel())):
        ne = len(a)
        rows += [a, b]; cols += [e + np.arange(ne)] * 2; vals += [np.ones(ne), -np.ones(ne)]; e += ne
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(k * k, e))
    kk, st = _check(A, "K1")
    S = sr.kkt_pattern(A, "K1")
    _, cc_nat, _ = sr.symbolic_bruteforce(S, np.arange(k * k))
    lu = spla.splu((A @ A.T + sp.identity(k * k)).tocsc(), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0,
                   options=dict(SymmetricMode=True))
    assert st["nnzL"] < 0.5 * cc_nat.sum()
    assert st["nnzL"] < 1.25 * lu.L.nnz


def test_natural_ordering_option():
    A = lpgen.config(2, mini=True).A
    k = pkg.setup(A, pkg.K1(), pkg.Backend(analyze_only=True, ordering=0))
    sym = k.symbolic()
    S = sr.kkt_pattern(A, "K1")
    parent, cc, _ = sr.symbolic_bruteforce(S, sym["perm"])
    assert np.array_equal(cc, sym["colcount"])


def test_no_cpu_fallback():
    """analyze_only solvers (and any box without a GPU) must refuse numeric work loudly."""
    A = lpgen.config(2, mini=True).A
    k = pkg.setup(A, pkg.K1(), pkg.Backend(analyze_only=True))
    m, n = A.shape
    with pytest.raises(pkg.TlpB200Error):
        k.update(np.ones(n), np.ones(n), np.ones(m))
    with pytest.raises(pkg.TlpB200Error):
        k.solve(np.zeros(n), np.zeros(m), np.ones(m), np.ones(n))


def test_create_rejects_bad_arguments():
    lib = _lib.load()
    h = C.c_void_p()
    colptr = (C.c_int64 * 3)(0, 1, 5)           # claims 5 entries but rows hold an out-of-range index
    rowval = (C.c_int64 * 5)(0, 7, 0, 0, 0)
    nz = (C.c_double * 5)(1, 1, 1, 1, 1)
    opt = _lib.Options(); lib.tlpb200_default_options(C.byref(opt)); opt.analyze_only = 1
    rc = lib.tlpb200_create(C.byref(h), 2, 2, colptr, rowval, nz, 0, 1, C.byref(opt))
    assert rc == _lib.BAD_ARG
    assert b"row index" in lib.tlpb200_last_error(h)
    lib.tlpb200_destroy(h)
    rc = lib.tlpb200_create(C.byref(h), 2, 2, colptr, rowval, nz, 0, 7, C.byref(opt))
    assert rc == _lib.BAD_ARG
    lib.tlpb200_destroy(h)


def test_julia_index_base():
    """index_base=1 (what the Julia glue passes) gives the same analysis as 0-based input."""
    lib = _lib.load()
    A = lpgen.config(3, mini=True).A
    m, n = A.shape
    opt = _lib.Options(); lib.tlpb200_default_options(C.byref(opt)); opt.analyze_only = 1
    perms = []
    for base in (0, 1):
        cp = np.ascontiguousarray(A.indptr, np.int64) + base
        ri = np.ascontiguousarray(A.indices, np.int64) + base
        h = C.c_void_p()
        rc = lib.tlpb200_create(C.byref(h), m, n, cp.ctypes.data_as(C.POINTER(C.c_int64)),
                                ri.ctypes.data_as(C.POINTER(C.c_int64)),
                                A.data.ctypes.data_as(C.POINTER(C.c_double)), base, 2, C.byref(opt))
        assert rc == _lib.OK
        perm = np.zeros(m + n, np.int32)
        lib.tlpb200_get_symbolic(h, C.c_void_p(perm.ctypes.data), None, None, None)
        perms.append(perm)
        lib.tlpb200_destroy(h)
    assert np.array_equal(perms[0], perms[1])
