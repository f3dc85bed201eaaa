"""The product's fill-reducing ordering against an INDEPENDENT one: the CPU baseline borrows the
product's symbolic analysis, so a poor ordering would inflate both arms alike and stay invisible.  Here nnz(L) of the
product's approximate-minimum-degree ordering is compared with SuperLU's multiple-minimum-degree ordering
(``splu(permc_spec="MMD_AT_PLUS_A")``, symmetric mode, no pivoting) on scaled-down instances of the BASELINE configs.
Measured here: 1.000 (cfg2-like), 1.008 (cfg4-like), 0.935 (cfg3-like, K2), 1.068 (cfg5-like, K2)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

import tlpb200_loader

pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen  # noqa: E402

CASES = {
    "cfg2-like": (lambda: lpgen.random_sparse(2500, 5000, 10, name="r"), "K1"),
    "cfg4-like": (lambda: lpgen.block_angular(blocks=8, mb=768, nb=1536, width=128, link=128, name="b"), "K1"),
    "cfg3-like": (lambda: lpgen.staircase(stages=16, nodes=410, arcs=600, name="s"), "K2"),
    "cfg5-like": (lambda: lpgen.dense_columns(m=4000, n=8000, ndense=4, dense_nnz=2000, sparse_nnz=24000, width=128, name="d"), "K2"),
}


@pytest.mark.parametrize("name", list(CASES))
def test_amd_fill_is_within_10_percent_of_superlu_mmd(name):
    gen, sysname = CASES[name]
    A = gen().A
    m, n = A.shape
    if sysname == "K1":
        K = (A @ A.T + sp.eye(m)).tocsc()                                   # pattern of spd.jl:14
    else:
        K = sp.bmat([[-sp.eye(n), A.T], [A, sp.eye(m)]], format="csc")      # pattern of sqd.jl:13-16
    k = pkg.setup(A, pkg.K1() if sysname == "K1" else pkg.K2(), pkg.Backend(analyze_only=True, dense_col_threshold=-1))
    mine = k.stats()["nnzL"]
    # same pattern made strictly diagonally dominant: SuperLU then never pivots off the diagonal, nnz(L) is the symbolic fill
    Kd = (abs(K) + 10.0 * float(abs(K).sum(axis=1).max()) * sp.eye(K.shape[0])).tocsc()
    lu = sla.splu(Kd, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    assert np.array_equal(lu.perm_r, lu.perm_c), "SuperLU pivoted off the diagonal: nnz(L) is not the symbolic fill"
    ratio = mine / lu.L.nnz
    print(f"{name}: product AMD nnz(L) = {mine}, SuperLU MMD nnz(L) = {lu.L.nnz}, ratio {ratio:.3f}")
    assert ratio <= 1.10, ratio
