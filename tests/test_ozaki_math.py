"""CPU statement of the arithmetic of the tcgen05 int8 update path (tulip.jl_b200/csrc/kernels_ozaki.cu): the digit
planes written by k_oz_slice reconstruct the 62-bit fixed-point value exactly, every int32 accumulator stays in range for
K <= 4096, and recombining the 36 plane pairs with p + q <= 7 the way the epilogue does gives the FP64 product  L L'
(the dense SYRK of /root/reference/src/KKT/Dense/lapack.jl:85-88) to better than FP64-GEMM accuracy.  NumPy integers only:
this is the specification the GPU tests (tests/test_gpu_ozaki.py) hold the kernels to."""
import numpy as np
import pytest

S = 8


def row_exponent(d):
    """k_oz_rowexp: sqrt(d) < 2^E from the diagonal entry d > 0."""
    _, e = np.frexp(d)
    return (e + 1) >> 1


def digit_planes(X, E):
    """k_oz_slice: q = rint(x 2^(62-E)) and its 8 balanced base-256 digits, most significant first."""
    q = np.rint(np.ldexp(X, (62 - E)[:, None])).astype(np.int64)
    planes = np.zeros((S,) + X.shape, np.int64)
    for p in range(S - 1, 0, -1):
        d = ((q & 0xFF) ^ 0x80) - 0x80            # sign-extended low byte
        q = (q - d) >> 8
        planes[p] = d
    planes[0] = q
    return planes


@pytest.mark.parametrize("K,spread", [(64, 0.0), (512, 6.0), (4096, 2.0)])
def test_digit_planes_and_recombination(K, spread):
    rng = np.random.default_rng(K)
    R = 96
    X = rng.standard_normal((R, K))
    if spread:
        X *= np.exp(rng.uniform(-spread, spread, (R, 1))) * np.exp(rng.uniform(-spread, 0, (R, K)))
    E = row_exponent((X * X).sum(1))
    assert np.all(np.abs(X) < np.ldexp(1.0, E)[:, None])                      # |L_rk| <= sqrt(K_rr) < 2^E
    P = digit_planes(X, E)
    assert P[1:].min() >= -128 and P[1:].max() <= 127 and np.abs(P[0]).max() <= 65        # all planes fit int8
    q = sum(P[p] << (8 * (S - 1 - p)) for p in range(S))
    assert np.array_equal(q, np.rint(np.ldexp(X, (62 - E)[:, None])).astype(np.int64))    # exact reconstruction
    # TMEM accumulators: level t = p + q, exact integer sums
    acc = [sum(P[p] @ P[t - p].T for p in range(t + 1)) for t in range(S)]
    assert max(int(np.abs(a).max()) for a in acc) < 2 ** 31                   # int32 range (K <= 4096)
    hi = ((acc[0] * 256 + acc[1]) * 256 + acc[2]) * 256 + acc[3]
    lo = ((acc[4] * 256 + acc[5]) * 256 + acc[6]) * 256 + acc[7]
    assert int(np.abs(hi).max()) < 2 ** 62 and int(np.abs(lo).max()) < 2 ** 62
    val = hi.astype(np.float64) * 2.0 ** 32 + lo.astype(np.float64)            # the epilogue's one conversion
    got = np.ldexp(val, (E[:, None] + E[None, :] - 68))
    Xl = X.astype(np.longdouble)
    ref = Xl @ Xl.T
    nrm = np.sqrt(np.outer((X * X).sum(1), (X * X).sum(1)))
    err = np.abs(got.astype(np.longdouble) - ref).astype(np.float64)
    # one rounding of the value + the dropped plane pairs (p + q >= 8): measured 2^-49.5 sqrt(K / 4096) of sqrt(K_ii K_jj)
    assert np.all(err <= 2.0 ** -51 * np.abs(ref).astype(np.float64) + 2.0 ** -49 * np.sqrt(max(K, 64) / 4096.0) * nrm)
    err64 = np.abs((X @ X.T).astype(np.longdouble) - ref).astype(np.float64)
    assert err.max() / nrm.max() <= max(4.0 * (err64 / nrm).max(), 2.0 ** -52)   # at least FP64-GEMM grade
