#!/usr/bin/env python
"""bench.py -- IPM iterations/sec (KKT factor+solve) on B200, next to a CPU baseline.

Contract: ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line (rank 0).

A *step* is the KKT work of one iteration of Tulip's homogeneous self-dual IPM on the synthetic LP:
1 ``update!`` (assemble + numeric factorisation) + that iteration's ``solve!`` calls (3-6), with
the (theta, regP, regD, xi) the real algorithm produces -- the driver is the host-side mirror of
the reference's caller (tulip.jl_b200/hsd.py <- src/IPM/HSD/step.jl).  The reference reads the
same quantity out of its TimerOutputs sections "Factorization" + "KKT" (plan A):
iterations/s = niter / (sum Factorization + sum KKT).

* ``e2e``  : the steps timed through the reference-facing host-pointer API (``KKT.update!`` /
             ``KKT.solve!`` with host vectors; H2D/D2H copies and syncs inside the timed region).
* ``value``: the same steps replayed with all inputs resident in HBM (``*_dev`` entry points),
             CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
* N > 1    : config 2 does not shard ("replicas only"): every rank runs an independent
             replica, value = N * steps / max-over-ranks time, scaling = "weak".
* ``--impl reference``: the oracle's CPU port of the reference path (oracle/cpu_kkt.py: SciPy
             SpGEMM assemble as in spd.jl:43 + own supernodal Cholesky on OpenBLAS, all host
             threads) driven by the oracle's HSD restatement -- NOT CHOLMOD (no Julia/SuiteSparse
             in the image).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "IPM iterations/sec (KKT factor+solve)"
UNIT = "iter/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_lp(cfg):
    import tlpb200_loader
    pkg = tlpb200_loader.load()
    from tulip_jl_b200 import lpgen
    if cfg == "2":
        return pkg, lpgen.config(2), "K1"
    if cfg == "3":
        return pkg, lpgen.config(3), "K2"
    if cfg == "4":
        return pkg, lpgen.config(4), "K1"
    if cfg == "4mini":
        return pkg, lpgen.config(4, mini=True), "K1"
    if cfg == "T":
        return pkg, lpgen.config("T"), "K1"
    if cfg == "mini":
        return pkg, lpgen.config(2, mini=True), "K1"
    raise SystemExit(f"unknown --config {cfg}")


def workload_name(lp, sysname, nsolve):
    md = lp.meta
    return (f"{lp.name}: {md.get('kind')} LP m={md['m']} n={md['n']} nnz(A)={md['nnz']}, {sysname} "
            f"({'normal equations Cholesky' if sysname == 'K1' else 'augmented LDLt'}); step = 1 update! + "
            f"{nsolve:.2f} solve! (mean over timed HSD iterations)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def record_hsd(pkg, lp, kkt, niter):
    """Run the HSD mirror for `niter` iterations through the host API, recording every KKT input
    and timing every KKT call.  Returns per-iteration records."""
    from tulip_jl_b200 import hsd
    recs = []
    cur = {}

    class Rec:
        m, n = kkt.m, kkt.n

        def update(self, th, rp, rd):
            cur.clear()
            cur.update(theta=th.copy(), regP=rp.copy(), regD=rd.copy(), rhs=[], t_update=0.0, t_solve=0.0)
            t0 = time.perf_counter()
            kkt.update(th, rp, rd)
            cur["t_update"] = time.perf_counter() - t0

        def solve(self, dx, dy, xp, xd):
            cur["rhs"].append((np.array(xp, copy=True), np.array(xd, copy=True)))
            t0 = time.perf_counter()
            kkt.solve(dx, dy, xp, xd)
            cur["t_solve"] += time.perf_counter() - t0

    h = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, Rec())
    h.optimize(max_iter=niter, callback=lambda hh: recs.append(dict(cur)))
    return h, recs


def gpu_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    pkg, lp, sysname = build_lp(args.config)
    A = lp.A
    m, n = A.shape
    sy = pkg.K1() if sysname == "K1" else pkg.K2()
    K, W = args.steps, args.warmup
    if world > 1 and args.config.startswith("4"):
        return sharded_arm(args, pkg, lp, sysname, sy, dist, rank, world, local)
    t0 = time.time()
    kkt = pkg.setup(A, sy, pkg.Backend(device=local))
    t_setup = time.time() - t0
    # ---- pass 1: the real IPM through the host API (e2e) ------------------------------------
    sampler = ClockSampler(local)
    h, recs = record_hsd(pkg, lp, kkt, W + K)
    if len(recs) < W + 1:
        raise SystemExit("IPM terminated during warm-up; lower --warmup")
    base = len(recs)
    while len(recs) < W + K:                # converged early: keep cycling over the recorded iterations
        recs.append(recs[W + (len(recs) - base) % (base - W)])
    timed = recs[W:W + K]
    nsolve = float(np.mean([len(r["rhs"]) for r in timed]))
    e2e_time = sum(r["t_update"] + r["t_solve"] for r in timed)
    h2d = float(np.mean([(2 * n + m) * 8 + len(r["rhs"]) * (n + m) * 8 for r in timed]))
    d2h = float(np.mean([4 + len(r["rhs"]) * (n + m) * 8 for r in timed]))
    st0 = kkt.stats()
    # ---- pass 2: device-resident replay (value) ---------------------------------------------
    dev = torch.device(f"cuda:{local}")
    stream = torch.cuda.current_stream()
    kkt.set_stream(stream.cuda_stream)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    dev_recs = [dict(theta=tt(r["theta"]), regP=tt(r["regP"]), regD=tt(r["regD"]),
                     rhs=[(tt(xp), tt(xd)) for xp, xd in r["rhs"]]) for r in recs[:W + K]]
    ddx = torch.zeros(n, dtype=torch.float64, device=dev)
    ddy = torch.zeros(m, dtype=torch.float64, device=dev)

    def replay(r):
        kkt.update_dev(r["theta"], r["regP"], r["regD"])
        kkt.update_status()          # the caller must know the factorisation succeeded (step.jl:34-51)
        for xp, xd in r["rhs"]:
            kkt.solve_dev(ddx, ddy, xp, xd)

    for r in dev_recs[:W]:
        replay(r)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for r in dev_recs[W:W + K]:
        replay(r)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    if dist:
        t = torch.tensor([ms_total, e2e_time * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms = float(t[0]), float(t[1])
        e2e_time = e2e_ms / 1e3
    launches = int(sum(st0["launches_update"] + len(r["rhs"]) * st0["launches_solve"] for r in timed))
    # ---- pass 3: per-kernel-class profile of one step (roofline numerators) -------------------
    kkt.set_stream(0)
    kkt.set_profiling(True)
    r = recs[W]
    kkt.update(r["theta"], r["regP"], r["regD"])
    dx = np.zeros(n); dy = np.zeros(m)
    kkt.solve(dx, dy, r["rhs"][0][0], r["rhs"][0][1])
    sp = kkt.stats()
    kkt.set_profiling(False)
    cls = dict(zip(pkg._lib.KERNEL_CLASSES, zip(sp["ms_class"], sp["n_class"])))
    peaks, peak_src = measured_peaks()
    ms_upd = cls["update"][0] + cls["update128"][0]
    n_upd = cls["update"][1] + cls["update128"][1]
    flops_upd = sp["flops_update_ext"]
    # measured FP64 GEMM ceiling (cuBLAS DGEMM through torch), same spirit as MEASURED_PEAKS' bf16 number
    a = torch.randn(4096, 4096, dtype=torch.float64, device=dev); b = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    best = 1e9
    for _ in range(4):
        s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
        s0.record(); torch.matmul(a, b); s1.record(); torch.cuda.synchronize()
        best = min(best, s0.elapsed_time(s1))
    dgemm_tf = 2 * 4096 ** 3 / (best * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k_update_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    ach = flops_upd / (ms_upd * 1e-3) / 1e12 if ms_upd > 0 else 0.0
    roofline_dmma = {"kernel": "k_update (FP64 DMMA m8n8k4 tile update: supernode SYRK/GEMM + scatter)",
                     "bound": "tensor", "achieved": round(ach, 3), "peak": round(dgemm_tf, 2), "unit": "TFLOP/s",
                     "frac": round(ach / dgemm_tf, 4) if dgemm_tf > 0 else None, "traffic": traffic,
                     "peak_source": "cuBLAS DGEMM 4096^3 measured live in this run (MEASURED_PEAKS.json has no FP64 figure; nominal B200 FP64 = 40 TFLOP/s)",
                     "launches_per_step": int(n_upd), "avg_launch_ms": round(ms_upd / max(1, n_upd), 4),
                     "algorithmic_flops_per_step": flops_upd, "share_of_update_ms": round(ms_upd / max(1e-9, sp["ms_assemble"] + sp["ms_factor"]), 3)}
    ms_oz, n_oz = cls["oz_update"]
    if n_oz > 0 and sp["flops_update_oz"] > 0:
        # dominant kernel: the tcgen05 int8 (Ozaki) update.  One FP64 multiply-add = 36 int8 digit-plane multiply-adds, so
        # the algorithmic tensor work is 36 x the FP64 flops; peak = dense int8 rate = 2 x the measured dense bf16 rate
        # (B200: 4.5 vs 2.25 POP/s nominal), burst figure because profiling mode times every launch alone.
        int8_peak = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
        tops = 36.0 * sp["flops_update_oz"] / (ms_oz * 1e-3) / 1e12
        tp2 = os.path.join(ROOT, "profiles", "k_oz_update_traffic.json")
        traffic_oz = None
        if os.path.exists(tp2):
            try:
                traffic_oz = json.load(open(tp2)).get("dram_bytes_per_launch")
            except Exception:
                traffic_oz = None
        roofline = {"kernel": "k_oz_update (tcgen05.mma kind::i8, 36 digit-plane products per FP64 product, accumulators in TMEM)",
                    "bound": "tensor", "achieved": round(tops, 1), "peak": round(int8_peak, 1), "unit": "TOP/s (int8)",
                    "frac": round(tops / int8_peak, 4), "traffic": traffic_oz,
                    "peak_source": f"2 x dense bf16 {peak_src} (no int8 figure in MEASURED_PEAKS.json; int8 dense = 2 x bf16 dense on B200)",
                    "fp64_equivalent_tflops": round(sp["flops_update_oz"] / (ms_oz * 1e-3) / 1e12, 2),
                    "fp64_dgemm_tflops_measured": round(dgemm_tf, 2),
                    "launches_per_step": int(n_oz), "avg_launch_ms": round(ms_oz / n_oz, 4),
                    "algorithmic_flops_per_step": sp["flops_update_oz"], "tasks_per_step": sp["oz_tasks"],
                    "share_of_update_ms": round(ms_oz / max(1e-9, sp["ms_assemble"] + sp["ms_factor"]), 3),
                    "note": "CUDA-event bracket around every launch of the class, kernels serialised (profiling mode)"}
    else:
        roofline = roofline_dmma
    ms_tri = sum(cls[k][0] for k in ("fwd_small", "fwd_large", "bwd_large", "bwd_small", "fwd_big", "bwd_big"))
    # dominant solve kernels: the dense sweeps over the big supernodes (k_fwd_big + k_bwd_big); their algorithmic
    # bytes are 16 B per non-zero of the big supernodes' panels (L read once forward, once backward: SURVEY 8d)
    bp = kkt.big_plan()
    sym = kkt.symbolic()
    big_sn = np.unique(bp["fwd"]["sn"]) if len(bp["fwd"]) else np.zeros(0, np.int64)
    # structural non-zeros of those supernodes' columns (exact column counts: the panels also hold the explicit zeros
    # of relaxed amalgamation, which are streamed but are not algorithmic bytes)
    cc = np.asarray(sym["colcount"], dtype=np.float64)
    csum = np.concatenate([[0.0], np.cumsum(cc)])
    nnz_big = float(np.sum(csum[sym["sn_first"][big_sn + 1]] - csum[sym["sn_first"][big_sn]]))
    ms_big = cls["fwd_big"][0] + cls["bwd_big"][0]
    ach_big = 16.0 * nnz_big / (ms_big * 1e-3) / 1e9 if ms_big > 0 else None
    roofline_solve = {"kernel": "k_fwd_big + k_bwd_big (dense sweeps over the big supernodes, one rhs)", "bound": "hbm",
                      "achieved": round(ach_big, 1) if ach_big else None, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                      "frac": round(ach_big / peaks["hbm_gbs"], 4) if ach_big else None,
                      "algorithmic_bytes": 16.0 * nnz_big, "ms": round(ms_big, 4), "launches_per_solve": int(cls["fwd_big"][1] + cls["bwd_big"][1]),
                      "streamed_tile_bytes": (bp["n_ftiles"] + bp["n_btiles"]) * 128 * 128 * 8,
                      "share_of_L": round(nnz_big / max(1.0, float(sp["nnzL"])), 4),
                      "whole_sweep": {"ms": round(ms_tri, 4), "algorithmic_bytes": 16.0 * sp["nnzL"],
                                      "achieved": round(16.0 * sp["nnzL"] / (ms_tri * 1e-3) / 1e9, 1) if ms_tri > 0 else None,
                                      "frac": round(16.0 * sp["nnzL"] / (ms_tri * 1e-3) / 1e9 / peaks["hbm_gbs"], 4) if ms_tri > 0 else None,
                                      "note": "all sweep kernels of one solve (small / medium / below / big), CUDA-event brackets per launch"},
                      "peak_source": peak_src}
    phases = {k: {"ms": round(v[0], 4), "launches": int(v[1])} for k, v in cls.items()}
    out = {
        "metric": METRIC, "value": round(world * K / (ms_total * 1e-3), 4), "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": round(ms_total / K, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(lp, sysname, nsolve),
                   "parallelism": "replicas only (this config's elimination tree does not shard)" if world > 1 else "single GPU",
                   "l2": f"inputs larger than L2: factor panels {sp['nnzL_stored'] * 8 / 1e6:.0f} MB streamed every step (L2 126 MB); no explicit flush",
                   "nnzL": sp["nnzL"], "factor_flops": sp["flops"], "nsuper": sp["nsuper"], "levels": sp["nlevels"],
                   "setup_s": round(t_setup, 2), "ipm_status_after": h.status, "ipm_iters_run": h.niter},
        "clocks": clocks,
        "e2e": {"value": round(world * K / e2e_time, 4), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": round(e2e_time * 1e3 / K, 4)},
        "gpu_launches": launches,
        "update_ms_host_api": round(float(np.mean([r["t_update"] for r in timed])) * 1e3, 3),
        "solve_ms_host_api": round(float(np.sum([r["t_solve"] for r in timed]) / np.sum([len(r["rhs"]) for r in timed])) * 1e3, 3),
        "roofline": roofline, "roofline_dmma": roofline_dmma, "roofline_solve": roofline_solve, "phases_one_step": phases,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_sample(pkg, lp, sysname, recs[W])
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


def sharded_arm(args, pkg, lp, sysname, sy, dist, rank, world, local):
    """BASELINE configs[3]: block-angular LP, elimination-tree subtrees sharded across the ranks, NCCL all-reduce of
    the separator front (tulip.jl_b200/parallel.py).  Every rank runs the same IPM (SPMD); timing = sum of the
    KKT calls of the timed iterations through the host API, max over ranks.  Strong scaling: the job is fixed."""
    import torch
    from tulip_jl_b200 import parallel
    K, W = args.steps, args.warmup
    t0 = time.time()
    kkt = parallel.DistB200KKT(lp.A, sy, pkg.Backend(device=local))
    t_setup = time.time() - t0
    sampler = ClockSampler(local)
    dist.barrier(); torch.cuda.synchronize()
    sampler.start()
    h, recs = record_hsd(pkg, lp, kkt, W + K)
    torch.cuda.synchronize(); dist.barrier()
    clocks = sampler.stop()
    base = len(recs)
    if base < W + 1:
        raise SystemExit("IPM terminated during warm-up; lower --warmup")
    timed = recs[W:min(base, W + K)]
    k_done = len(timed)
    t = torch.tensor([sum(r["t_update"] + r["t_solve"] for r in timed)], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = float(t[0])
    nsolve = float(np.mean([len(r["rhs"]) for r in timed]))
    st = kkt.stats()
    owner, off, cnt = kkt.dist_info()
    m, n = lp.A.shape
    val = k_done / tot
    out = {"metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": world, "steps": k_done, "warmup": W,
           "ms_per_step": round(tot * 1e3 / k_done, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(lp, sysname, nsolve),
                      "parallelism": f"etree subtrees sharded over {world} ranks + NCCL all-reduce of the separator panels "
                                     f"({cnt * 8 / 1e6:.2f} MB per update!, {2 * 8 * st['order'] / 1e6:.2f} MB per solve!)",
                      "l2": "timed through the host-pointer API (H2D/D2H inside); factor panels exceed L2",
                      "nnzL": st["nnzL"], "factor_flops": st["flops"], "setup_s": round(t_setup, 2),
                      "ipm_status_after": h.status, "ipm_iters_run": h.niter},
           "clocks": clocks,
           "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": int((2 * n + m) * 8 + nsolve * (n + m) * 8),
                   "d2h_bytes_per_step": int(4 + nsolve * (n + m) * 8), "ms_per_step": round(tot * 1e3 / k_done, 4)},
           "gpu_launches": int(k_done * (st["launches_update"] + nsolve * st["launches_solve"])),
           "roofline": {"bound": "tensor", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                        "note": "per-kernel roofline is reported by the single-GPU run (--gpus 1)"},
           "update_ms_host_api": round(float(np.mean([r["t_update"] for r in timed])) * 1e3, 3),
           "solve_ms_host_api": round(float(np.sum([r["t_solve"] for r in timed]) / np.sum([len(r["rhs"]) for r in timed])) * 1e3, 3)}
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


EXTRAPOLATE_ABOVE_FLOPS = 5e12     # one CPU factorisation beyond this does not fit a bounded sample


def cpu_sample_extrapolated(st, nsolve):
    """Bounded CPU sample for a dense-dominated config whose CPU factorisation takes minutes to hours (config T:
    SURVEY 8d option A): time the two dense kernels that carry > 99 % of the CPU port's work on this workload -- LAPACK
    dpotrf and the triangular solves, SciPy's OpenBLAS, all host cores -- on an n0 x n0 block, and scale by the
    factorisation's flop count sum_j c_j^2 and the solves' 16 nnz(L) bytes.  Labelled as an extrapolation."""
    import scipy.linalg as sla
    cores = os.cpu_count() or 1
    n0 = 12000
    rng = np.random.default_rng(0)
    M = rng.standard_normal((n0, 64))
    S = np.asfortranarray(M @ M.T)
    S[np.diag_indices(n0)] += 64.0
    t0 = time.perf_counter()
    L, info = sla.lapack.dpotrf(S, lower=1, overwrite_a=1)
    t_f = time.perf_counter() - t0
    rate = n0 ** 3 / 3.0 / t_f                      # flops/s in the c_j^2 convention (sum_j c_j^2 = n^3/3 when dense)
    x = rng.standard_normal(n0)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        y = sla.solve_triangular(L, x, lower=True, check_finite=False)
        y = sla.solve_triangular(L, y, lower=True, trans=1, check_finite=False)
    t_s = (time.perf_counter() - t0) / reps
    bw = 16.0 * (n0 * (n0 + 1) / 2) / t_s           # algorithmic bytes/s of one forward+backward solve
    t_iter = st["flops"] / rate + nsolve * 16.0 * st["nnzL"] / bw
    return {"value": round(1.0 / t_iter, 6), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"EXTRAPOLATED (a full CPU factorisation of this config takes ~{st['flops'] / rate / 60:.0f} min): dense dpotrf "
                      f"{n0}^2 in {t_f:.1f} s = {rate / 1e9:.0f} GF/s (c_j^2 convention) and forward+backward solve at {bw / 1e9:.1f} GB/s, "
                      f"scaled to sum c_j^2 = {st['flops']:.3e} flops + {nsolve:.2f} solves x 16 nnz(L) = {16.0 * st['nnzL'] / 1e9:.1f} GB",
            "label": "CPU port kernels (SciPy-OpenBLAS dpotrf / dtrsv, all host cores) -- NOT Tulip/CHOLMOD; not run to completion"}


def cpu_sample(pkg, lp, sysname, rec):
    """oracle CPU port on a bounded sample: ONE recorded IPM iteration (1 update! + its solves)."""
    from oracle import cpu_kkt
    A = lp.A
    m, n = A.shape
    sy = pkg.K1() if sysname == "K1" else pkg.K2()
    an = pkg.setup(A, sy, pkg.Backend(analyze_only=True))
    cores = os.cpu_count() or 1
    st = an.stats()
    if st["flops"] > EXTRAPOLATE_ABOVE_FLOPS:
        return cpu_sample_extrapolated(st, float(len(rec["rhs"])))
    ck = cpu_kkt.CpuSupernodalKKT(A, sysname, nthreads=cores, symbolic_from=an)
    t0 = time.perf_counter()
    ck.update(rec["theta"], rec["regP"], rec["regD"])
    dx = np.zeros(n); dy = np.zeros(m)
    for xp, xd in rec["rhs"]:
        ck.solve(dx, dy, xp, xd)
    dt = time.perf_counter() - t0
    return {"value": round(1.0 / dt, 5), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 recorded HSD iteration of the same workload (1 update! + {len(rec['rhs'])} solve!), {dt:.2f} s",
            "label": "CPU port: SciPy SpGEMM assemble + own left-looking supernodal Cholesky on SciPy-OpenBLAS -- NOT Tulip/CHOLMOD"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    pkg, lp, sysname = build_lp(args.config)
    from oracle import cpu_kkt, hsd_ref
    # the reference runs its own vector work with BLAS threads = 1 (model.jl:73); the factorisation's BLAS
    # (SciPy's OpenBLAS, a separate library instance) gets every core below.
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=1, user_api="blas")
    A = lp.A
    sy = pkg.K1() if sysname == "K1" else pkg.K2()
    an = pkg.setup(A, sy, pkg.Backend(analyze_only=True))       # integer analysis only, no device
    cores = os.cpu_count() or 1
    st = an.stats()
    if st["flops"] > EXTRAPOLATE_ABOVE_FLOPS:
        nsolve = 5.0
        cb = cpu_sample_extrapolated(st, nsolve)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 / cb["value"], 1),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": workload_name(lp, sysname, nsolve)}, "cpu_baseline": cb,
                          "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    ck = cpu_kkt.CpuSupernodalKKT(A, sysname, nthreads=cores, symbolic_from=an)
    K, W = args.steps, args.warmup
    marks = []

    class T:
        def update(self, *a):
            ck.update(*a)

        def solve(self, *a):
            ck.solve(*a)

    dat = hsd_ref.IPMData(A, lp.b, True, lp.c, 0.0, lp.l, lp.u)
    hs = hsd_ref.HSDRef(dat, T(), hsd_ref.IPMOptions(IterationsLimit=W + K))
    hs.optimize(callback=lambda s: marks.append((s.t_factor + s.t_solve, s.n_solve)))
    if len(marks) <= W:
        raise SystemExit("reference arm: IPM ended during warm-up")
    t_w, ns_w = marks[W - 1] if W > 0 else (0.0, 0)
    t_e, ns_e = marks[-1]
    k_done = len(marks) - W
    dt = t_e - t_w
    val = k_done / dt
    nsolve = (ns_e - ns_w) / k_done
    out = {"impl": "reference", "metric": METRIC, "value": round(val, 5), "unit": UNIT, "n_gpus": world,
           "steps": k_done, "warmup": W, "ms_per_step": round(dt * 1e3 / k_done, 3), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(lp, sysname, nsolve)},
           "cpu_baseline": {"value": round(val, 5), "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{k_done} HSD iterations (after {W} warm-up) of the same workload, {dt:.1f} s of KKT time",
                            "label": "CPU port: SciPy SpGEMM assemble + own left-looking supernodal Cholesky on SciPy-OpenBLAS -- NOT Tulip/CHOLMOD (no Julia/SuiteSparse in the image)"},
           "e2e": {"value": round(val, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="2", help="2 (default, BASELINE configs[1]) | 3 | T | mini")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("note: timing rules ask for >= 3 warm-up steps")
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
