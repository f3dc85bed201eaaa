"""CPU replay of the critical-chain kernels' index logic (scripts/chain_emu.py mirrors k_diag_factor2 / k_trsm2 of
tulip.jl_b200/csrc/kernels_factor.cu lane by lane): ragged widths, signed pivots (K2), every trailing tile updated exactly
once (asserted inside the mirror), never-loaded shared memory (1e300 marks) must not reach the factor, bad pivots reported."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import chain_emu as ce  # noqa: E402


def _problem(w, signed, seed):
    rng = np.random.default_rng(seed)
    sign = np.ones(w)
    if signed:
        sign[rng.random(w) < 0.4] = -1
    L0 = np.tril(rng.standard_normal((w, w))) * 0.3
    L0[np.diag_indices(w)] = 1 + rng.random(w)
    return rng, sign, L0, L0 @ np.diag(sign) @ L0.T


@pytest.mark.parametrize("w", [1, 5, 8, 17, 64, 100, 121, 128])
@pytest.mark.parametrize("signed", [False, True])
def test_diag_block_mirror_factors_to_rounding(w, signed):
    _, sign, L0, K = _problem(w, signed, 100 + w)
    L, bad = ce.diag_factor4(K, sign)
    assert not bad and np.isfinite(L).all()
    assert np.abs(L @ np.diag(sign) @ L.T - K).max() <= 1e-12 * np.abs(K).max()
    assert np.abs(L - L0).max() <= 1e-10 * (1 + np.abs(L0).max())


def test_diag_block_mirror_reports_the_first_bad_pivot():
    K = np.eye(24)
    K[13, 13] = -1.0
    K[20, 20] = 0.0
    _, bad = ce.diag_factor4(K, np.ones(24))
    assert min(bad) == 13 and 20 in bad


@pytest.mark.parametrize("w,nr", [(5, 3), (17, 64), (100, 37), (128, 64)])
@pytest.mark.parametrize("signed", [False, True])
def test_trsm_mirror_matches_a_dense_solve(w, nr, signed):
    rng, sign, L0, _ = _problem(w, signed, 200 + w)
    A = rng.standard_normal((nr, w))
    X = ce.trsm3(L0, sign, A)
    ref = np.linalg.solve(L0, A.T).T @ np.diag(sign)
    assert np.abs(X - ref).max() <= 1e-9 * np.abs(ref).max()


@pytest.mark.parametrize("nb", [1, 5, 16, 17, 100, 128])
def test_block_doubling_inverse_mirror(nb):
    """k_invert_diag2: 16x16 substitution + three levels of DMMA block doubling; every output entry written once per phase,
    upper triangle exactly zero (asserted inside the mirror)"""
    rng = np.random.default_rng(300 + nb)
    L = np.tril(rng.standard_normal((nb, nb))) * 0.3
    L[np.diag_indices(nb)] = 1 + rng.random(nb)
    X = ce.invert2(L)
    assert np.abs(X @ L - np.eye(nb)).max() <= 1e-12
