"""GPU probe of the tcgen05 int8 (Ozaki) Schur-update path: exactness vs NumPy FP64 and throughput."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import tlpb200_loader

pkg = tlpb200_loader.load()
lib = pkg._lib.load()


def run(P, C0, ksplit=0, reps=0):
    R, K = P.shape
    Pf = np.asfortranarray(P)
    Cf = np.asfortranarray(C0.copy())
    ms = (C.c_float * 3)()
    err = C.c_int32(0)
    rc = lib.tlpb200_debug_ozaki(Pf.ctypes.data_as(C.POINTER(C.c_double)), R, K, Cf.ctypes.data_as(C.POINTER(C.c_double)),
                                 ksplit, reps, ms, C.byref(err))
    return rc, err.value, Cf, list(ms)


def check(R, K, ksplit, seed, spread=0.0, zero_c=True):
    rng = np.random.default_rng(seed)
    P = rng.standard_normal((R, K))
    if spread:
        P *= np.exp(rng.uniform(-spread, spread, (R, 1)))          # rows of very different magnitude
        P *= np.exp(rng.uniform(-spread / 2, 0, (R, K)))           # and a wide range inside each row
    C0 = np.zeros((R, R)) if zero_c else rng.standard_normal((R, R))
    rc, err, Cg, ms = run(P, C0, ksplit)
    Pl = P.astype(np.longdouble)
    ref = C0.astype(np.longdouble) - Pl @ Pl.T                       # 64-bit mantissa reference
    low = np.tril(np.ones((R, R), bool))
    d = np.abs(Cg.astype(np.longdouble) - ref)[low].astype(np.float64)
    nrm = np.sqrt(np.outer((P * P).sum(1), (P * P).sum(1)))[low]     # sqrt(K_ii K_jj): the Cholesky error scale
    mag = np.abs(ref)[low].astype(np.float64) + np.abs(C0)[low]
    nsplit = max(1, -(-K // ksplit)) if ksplit else 1
    # one rounding of the product (+ one per RED) relative to the values, plus the dropped digit pairs relative to the row scales
    tol = 2.0 ** -52 * nsplit * mag + 64.0 * np.sqrt(7.0 * K) * 5476.0 * 2.0 ** -76 * nrm + (2.0 ** -51 * nrm if K > 2048 else 0.0)
    worst = float(np.max(d / tol))
    up_untouched = np.array_equal(Cg[~low], C0[~low])
    ref64 = C0 - P @ P.T
    d64 = np.abs(ref64.astype(np.longdouble) - ref)[low].astype(np.float64)
    print(f"R={R} K={K} ksplit={ksplit} spread={spread} zero_c={zero_c}: rc={rc} err={err} max err/tol={worst:.3f} "
          f"max|d|/sqrt(KiiKjj)={np.max(d / nrm):.3e} (NumPy FP64 GEMM: {np.max(d64 / nrm):.3e}) upper untouched={up_untouched} "
          f"slice ms={ms[0]:.3f} tasks={int(ms[2])}", flush=True)
    return rc == 0 and err == 0 and worst <= 1.0 and up_untouched


def bench(R, K, ksplit, reps=5):
    rng = np.random.default_rng(1)
    P = rng.standard_normal((R, K))
    C0 = np.zeros((R, R))
    rc, err, Cg, ms = run(P, C0, ksplit, reps)
    flops = float(R) * (R + 128) * K        # lower triangle incl. the full diagonal tiles, 2 flops per MAC
    print(f"bench R={R} K={K} ksplit={ksplit}: rc={rc} err={err} update {ms[1]:.3f} ms = {flops / ms[1] / 1e9:.1f} TFLOP/s FP64-equivalent; "
          f"slicing {ms[0]:.3f} ms = {R * K * 16 / ms[0] / 1e6:.0f} GB/s; tasks={int(ms[2])}", flush=True)


if __name__ == "__main__":
    ok = True
    ok &= check(128, 32, 0, 0)
    ok &= check(128, 128, 0, 1)
    ok &= check(256, 256, 0, 2)
    ok &= check(300, 200, 64, 3)
    ok &= check(640, 1024, 256, 4, spread=8.0)
    ok &= check(640, 1024, 0, 5, spread=8.0, zero_c=False)
    ok &= check(1000, 4096, 0, 6, spread=3.0)
    print("ALL OK" if ok else "FAILED", flush=True)
    if ok or "--force" in sys.argv:
        bench(4096, 1024, 0)
        bench(8192, 2048, 0)
        bench(8192, 2048, 512)
        bench(8192, 512, 0)
        bench(8192, 128, 0)
