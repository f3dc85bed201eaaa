#!/usr/bin/env python
"""bench.py -- IPM iterations/sec (KKT factor+solve) on B200, next to a CPU baseline.

Contract: ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line (rank 0).

A *step* is the KKT work of one iteration of Tulip's homogeneous self-dual IPM on the synthetic LP:
1 ``update!`` (assemble + numeric factorisation) + that iteration's ``solve!`` calls (3-6), with
the (theta, regP, regD, xi) the real algorithm produces -- the driver is the host-side mirror of
the reference's caller (tulip.jl_b200/hsd.py <- src/IPM/HSD/step.jl).  The reference reads the
same quantity out of its TimerOutputs sections "Factorization" + "KKT" (plan A):
iterations/s = niter / (sum Factorization + sum KKT).

* default workload (``--config auto``): N = 1 -> config T, the north-star LP of BASELINE.json (m=1e5, n=2e5,
             nnz 1.1e6, K1; it fits one GPU); N > 1 -> config 4 (block-angular), the only BASELINE config whose
             elimination tree shards: subtrees over the ranks + separator reduce, scaling = "strong"; rank 0 first
             times the same workload on one GPU (``n1_same_workload``).  The other configs (2, 3, 5, mini) are
             selected with ``--config``; with N > 1 they run as independent replicas (scaling = "weak").
* ``e2e``  : the steps timed through the reference-facing host-pointer API (``KKT.update!`` /
             ``KKT.solve!`` with host vectors; H2D/D2H copies and syncs inside the timed region).
* ``value``: the same steps replayed with all inputs resident in HBM (``*_dev`` entry points),
             CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
* the IPM runs to convergence (reference default IterationsLimit = 100); when it needs fewer than W + K
             iterations the timed window cycles over the recorded post-warm-up iterations (``steps_real``).
* ``--impl reference``: the oracle's CPU port of the reference path (oracle/cpu_kkt.py: SciPy
             SpGEMM assemble as in spd.jl:43 + own supernodal Cholesky on OpenBLAS, all host
             threads) driven by the oracle's HSD restatement -- NOT CHOLMOD (no Julia/SuiteSparse
             in the image).  On config T the CPU arm runs ONE real iteration (1 update! + its solves: minutes of
             host time), on the other configs the whole IPM.  Its result is cached in /tmp so that the GPU arm, run
             afterwards on the same box, quotes it as ``cpu_baseline`` and compares objectives with it.
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "IPM iterations/sec (KKT factor+solve)"
UNIT = "iter/s"
CPU_LABEL = ("CPU port: SciPy SpGEMM assemble + own left-looking supernodal Cholesky on SciPy-OpenBLAS -- NOT Tulip/CHOLMOD "
             "(no Julia/SuiteSparse in the image)")
ONE_ITER_ABOVE_FLOPS = 5e12     # CPU arm: beyond this a factorisation takes minutes -> ONE real iteration instead of the whole IPM


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout when the
# library's communicator is created), so the real stdout is kept aside for the JSON line and fd 1 is pointed at stderr.
_OUT = None


def claim_stdout():
    global _OUT
    if _OUT is None:
        sys.stdout.flush()
        _OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    print(json.dumps(obj), file=_OUT if _OUT is not None else sys.stdout, flush=True)


CONFIGS = {"2": (2, False, "K1"), "3": (3, False, "K2"), "4": (4, False, "K1"), "5": (5, False, "K1"), "T": ("T", False, "K1"),
           "mini": (2, True, "K1"), "3mini": (3, True, "K2"), "4mini": (4, True, "K1"), "5mini": (5, True, "K1"), "Tmini": ("T", True, "K1")}


def build_lp(cfg):
    import tlpb200_loader
    pkg = tlpb200_loader.load()
    from tulip_jl_b200 import lpgen
    if cfg not in CONFIGS:
        raise SystemExit(f"unknown --config {cfg}")
    num, mini, sysname = CONFIGS[cfg]
    return pkg, lpgen.config(num, mini=mini), sysname


def workload_name(lp, sysname):
    md = lp.meta
    return (f"{lp.name}: {md.get('kind')} LP m={md['m']} n={md['n']} nnz(A)={md['nnz']}, {sysname} "
            f"({'normal equations Cholesky' if sysname == 'K1' else 'augmented LDLt'}); step = 1 update! + its solve! calls "
            f"of one HSD iteration")


# config 5 (8 dense columns): the reference has no dense-column handling -- its K1 (spd.jl:43) would form a fully dense
# A D A' (10 GB factor, ~1 min per CPU factorisation) -- so the CPU arm runs this LP the way Tulip runs it by default,
# with the augmented system K2 (KKT.jl:134-137), which is the faster CPU choice (conservative for the speed-up quoted).
CPU_SYSTEM = {"5": "K2", "5mini": "K2"}


def cache_path(cfg):
    return os.path.join("/tmp", f"tlpb200_cpu_port_cfg{cfg}.json")


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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_median": float(np.median(pw)) if pw else None}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def ipm_summary(h):
    """final objective / residual figures of an HSD run (either the host mirror or the oracle restatement):
    log rows are (iter, pobj, dobj, pfeas, dfeas, gfeas, mu), the last one belongs to the final iterate."""
    it, pobj, dobj, pf, df, gf, mu = h.log[-1]
    out = {"status": h.status, "iters": int(h.niter), "pobj": float(pobj), "dobj": float(dobj),
           "rel_gap": float(abs(pobj - dobj) / (1.0 + abs(dobj))), "pfeas": float(pf), "dfeas": float(df), "mu": float(mu)}
    if len(h.log) > 1:
        _, p1, d1, _, _, _, mu1 = h.log[1]
        out["after_iter1"] = {"pobj": float(p1), "dobj": float(d1), "mu": float(mu1)}
    return out


def objective_parity(mine, ref):
    """relative differences between this arm's objectives and the CPU arm's (same LP, same algorithm, different KKT backend)"""
    if ref is None:
        return None
    rel = lambda a, b: float(abs(a - b) / max(1.0, abs(b)))
    out = {}
    if ref.get("status") == mine.get("status") == "Trm_Optimal":
        out["final_pobj_rel_diff"] = rel(mine["pobj"], ref["pobj"])
        out["final_dobj_rel_diff"] = rel(mine["dobj"], ref["dobj"])
        out["final_within_1e-8"] = bool(out["final_pobj_rel_diff"] <= 1e-8 and out["final_dobj_rel_diff"] <= 1e-8)
    if "after_iter1" in ref and "after_iter1" in mine:
        out["iter1_pobj_rel_diff"] = rel(mine["after_iter1"]["pobj"], ref["after_iter1"]["pobj"])
        out["iter1_dobj_rel_diff"] = rel(mine["after_iter1"]["dobj"], ref["after_iter1"]["dobj"])
        out["iter1_within_1e-8"] = bool(out["iter1_pobj_rel_diff"] <= 1e-8 and out["iter1_dobj_rel_diff"] <= 1e-8)
    out["reference_ipm"] = {k: ref.get(k) for k in ("status", "iters", "pobj", "dobj", "after_iter1")}
    out["note"] = ("both arms stop at the reference's default tolerances (sqrt(eps) = 1.5e-8 relative, options.jl:10-13), so two "
                   "correct runs may differ by ~1e-8 in the FINAL objective (SURVEY 8d); the iterate after iteration 1 is "
                   "deterministic given the same KKT answers and is the sharp comparison; tests/test_gpu_kkt.py::"
                   "test_ipm_end_to_end_parity repeats the final comparison with tolerances tightened to 1e-10")
    return out


def record_hsd(pkg, lp, kkt, niter):
    """Run the HSD mirror for at most `niter` iterations through the host API, recording every KKT input
    and timing every KKT call.  Returns the driver and the per-iteration records."""
    from tulip_jl_b200 import hsd
    recs = []
    cur = {}

    class Rec:
        m, n = kkt.m, kkt.n

        def update(self, th, rp, rd):
            # a regularisation bump (step.jl:34-51) repeats update!: the record keeps the successful inputs and ALL the time
            t_prev = cur.get("t_update", 0.0) if cur.get("open") else 0.0
            cur.clear()
            cur.update(theta=th.copy(), regP=rp.copy(), regD=rd.copy(), rhs=[], t_update=t_prev, t_solve=0.0, open=True)
            t0 = time.perf_counter()
            try:
                kkt.update(th, rp, rd)
            finally:
                cur["t_update"] += time.perf_counter() - t0

        def solve(self, dx, dy, xp, xd):
            cur["rhs"].append((np.array(xp, copy=True), np.array(xd, copy=True)))
            t0 = time.perf_counter()
            kkt.solve(dx, dy, xp, xd)
            cur["t_solve"] += time.perf_counter() - t0

    def done(hh):
        cur["open"] = False
        recs.append(dict(cur))

    h = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, Rec())
    t0 = time.perf_counter()
    h.optimize(max_iter=niter, callback=done)
    h.wall_s = time.perf_counter() - t0
    return h, recs


def device_ipm_figures(pkg, kkt_local, lp, limit, reduce_max=None):
    """the whole HSD loop resident on the device (tlpb200_hsd_*): KKT-only event time per iteration and wall clock"""
    from tulip_jl_b200 import hsd as hsd_mod
    d = pkg.DeviceHSD(kkt_local, lp.b, lp.c, lp.l, lp.u, params=hsd_mod.IPMOptions(IterationsLimit=limit))
    d.optimize()                      # warm-up of the IPM kernels
    d.optimize()
    kk = d.t_factor + d.t_solve
    wall = d.info["seconds_total"]
    if reduce_max is not None:
        kk, wall = reduce_max([kk, wall])
    return {"status": d.status, "iters": d.niter, "pobj": d.primal_objective, "dobj": d.dual_objective,
            "kkt_iter_per_s": round(d.niter / kk, 4) if kk > 0 else None, "kkt_ms_per_iter": round(kk * 1e3 / max(1, d.niter), 4),
            "whole_ipm_wall_s": round(wall, 4), "whole_ipm_iter_per_s": round(d.niter / wall, 4) if wall > 0 else None}


def timed_window(recs, W, K):
    """K records after W warm-up ones; cycles over the post-warm-up iterations when the IPM converged earlier"""
    base = len(recs)
    if base < W + 1:
        raise SystemExit(f"IPM terminated after {base} iterations, inside the warm-up; lower --warmup")
    out = list(recs)
    while len(out) < W + K:
        out.append(recs[W + (len(out) - base) % (base - W)])
    return out[:W + K], min(K, base - W)


def read_cpu_cache(cfg):
    """the CPU arm's result if `bench.py --impl reference` ran on this box within the last 6 hours"""
    try:
        d = json.load(open(cache_path(cfg)))
        if d.get("host") == socket.gethostname() and time.time() - d.get("when", 0) < 6 * 3600:
            return d
    except Exception:
        pass
    return None


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
    cfg = args.config
    if cfg == "auto":
        cfg = "T" if world == 1 else "4"
    pkg, lp, sysname = build_lp(cfg)
    A = lp.A
    m, n = A.shape
    sy = pkg.K1() if sysname == "K1" else pkg.K2()
    K, W = args.steps, args.warmup
    if world > 1 and cfg.startswith("4"):
        return sharded_arm(args, cfg, pkg, lp, sysname, sy, dist, rank, world, local)
    t0 = time.time()
    kkt = pkg.setup(A, sy, pkg.Backend(device=local))
    t_setup = time.time() - t0
    # ---- pass 1: the real IPM through the host API (e2e), run to convergence -------------------
    sampler = ClockSampler(local)
    h, recs = record_hsd(pkg, lp, kkt, max(args.ipm_limit, W + K) if args.ipm_limit > 0 else W + K)
    win, k_real = timed_window(recs, W, K)
    timed = win[W:W + K]
    nsolve = float(np.mean([len(r["rhs"]) for r in timed]))
    e2e_time = sum(r["t_update"] + r["t_solve"] for r in timed)
    h2d = float(np.mean([(2 * n + m) * 8 + len(r["rhs"]) * (n + m) * 8 for r in timed]))
    d2h = float(np.mean([4 + len(r["rhs"]) * (n + m) * 8 for r in timed]))
    st0 = kkt.stats()
    ipm = ipm_summary(h)
    # ---- pass 1b: the same IPM with the iteration resident on the device (tlpb200_hsd_*) ---------------
    ipm_dev = None
    if args.ipm_device == "on" or (args.ipm_device == "auto" and st0["flops"] <= ONE_ITER_ABOVE_FLOPS):
        try:
            ipm_dev = device_ipm_figures(pkg, kkt, lp, max(args.ipm_limit, W + K) if args.ipm_limit > 0 else W + K)
            ipm_dev.update({"host_mirror_whole_ipm_wall_s": round(h.wall_s, 4),
                            "host_mirror_whole_ipm_iter_per_s": round(h.niter / h.wall_s, 4),
                            "pobj_rel_diff_vs_host_mirror": float(abs(ipm_dev["pobj"] - h.primal_objective) / max(1.0, abs(h.primal_objective))),
                            "note": "the whole HSD loop (residuals, status tests, theta / rhs / recovery / step-length kernels, update!, "
                                    "solve!) with every vector resident in HBM: per-iteration host traffic = a 312-byte scalar block per "
                                    "decision; kkt_* = CUDA-event time of the update!/solve! sequences only (the metric's numerator), "
                                    "whole_ipm_* = wall clock of the complete solve next to the host-mirror driver's"})
        except Exception as e:          # the headline must not depend on the extra measurement
            ipm_dev = {"error": f"{type(e).__name__}: {e}"}
    # ---- pass 2: device-resident replay (value) ---------------------------------------------
    dev = torch.device(f"cuda:{local}")
    stream = torch.cuda.current_stream()
    kkt.set_stream(stream.cuda_stream)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cache = {}

    def dev_rec(r):
        if id(r) not in cache:
            cache[id(r)] = dict(theta=tt(r["theta"]), regP=tt(r["regP"]), regD=tt(r["regD"]), rhs=[(tt(xp), tt(xd)) for xp, xd in r["rhs"]])
        return cache[id(r)]

    dev_recs = [dev_rec(r) for r in win]
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
    r = win[W]
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
    del a, b

    def traffic_of(name):
        """DRAM bytes per launch from the committed ncu --set full capture -- only when it was taken on THIS config (a launch of
        another config moves a different amount of data); null otherwise"""
        tp = os.path.join(ROOT, "profiles", name)
        try:
            d = json.load(open(tp))
            return d.get("dram_bytes_per_launch") if str(d.get("config")) == cfg else None
        except Exception:
            return None

    ms_fac = max(1e-9, sp["ms_assemble"] + sp["ms_factor"])
    ach = flops_upd / (ms_upd * 1e-3) / 1e12 if ms_upd > 0 else 0.0
    roofline_dmma = {"kernel": "k_update (FP64 DMMA m8n8k4 tile update: supernode SYRK/GEMM + scatter)",
                     "bound": "tensor", "achieved": round(ach, 3), "peak": round(dgemm_tf, 2), "unit": "TFLOP/s",
                     "frac": round(ach / dgemm_tf, 4) if dgemm_tf > 0 else None, "traffic": traffic_of("k_update_traffic.json"),
                     "peak_source": "cuBLAS DGEMM 4096^3 measured live in this run (MEASURED_PEAKS.json has no FP64 figure; nominal B200 FP64 = 40 TFLOP/s)",
                     "launches_per_step": int(n_upd), "avg_launch_ms": round(ms_upd / max(1, n_upd), 4),
                     "algorithmic_flops_per_step": flops_upd, "share_of_update_ms": round(ms_upd / ms_fac, 3)}
    ms_oz, n_oz = cls["oz_update"]
    if n_oz > 0 and sp["flops_update_oz"] > 0:
        # dominant kernel: the tcgen05 int8 (Ozaki) update.  One FP64 multiply-add = 36 int8 digit-plane multiply-adds, so
        # the algorithmic tensor work is 36 x the FP64 flops of the tasks' tiles; peak = dense int8 rate = 2 x the measured
        # dense bf16 rate (B200: 4.5 vs 2.25 POP/s nominal).  `frac` uses the BURST figure (conservative); a launch of a
        # multi-second step runs at the sustained rate, reported as frac_vs_sustained.  frac_structural counts only the
        # structural flops sum_j c_j^2 share of the tiles (relaxed amalgamation pads the big supernodes with explicit zeros).
        int8_burst = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
        int8_sust = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
        tops = 36.0 * sp["flops_update_oz"] / (ms_oz * 1e-3) / 1e12
        tile_flops = sp["flops_update_oz"] + sp["flops_update_ext"] + sp["flops_update_inner"]
        structural_share = min(1.0, sp["flops"] / tile_flops) if tile_flops > 0 else 1.0
        roofline = {"kernel": "k_oz_update (tcgen05.mma kind::i8, 36 digit-plane products per FP64 product, accumulators in TMEM)",
                    "bound": "tensor", "achieved": round(tops, 1), "peak": round(int8_burst, 1), "unit": "TOP/s (int8)",
                    "frac": round(tops / int8_burst, 4), "traffic": traffic_of("k_oz_update_traffic.json"),
                    "frac_vs_sustained": round(tops / int8_sust, 4), "peak_sustained": round(int8_sust, 1),
                    "frac_structural": round(tops * structural_share / int8_burst, 4),
                    "structural_share_of_tile_flops": round(structural_share, 4),
                    "peak_source": f"2 x dense bf16 {peak_src} (no int8 figure in MEASURED_PEAKS.json; int8 dense = 2 x bf16 dense on B200)",
                    "fp64_equivalent_tflops": round(sp["flops_update_oz"] / (ms_oz * 1e-3) / 1e12, 2),
                    "fp64_dgemm_tflops_measured": round(dgemm_tf, 2),
                    "launches_per_step": int(n_oz), "avg_launch_ms": round(ms_oz / n_oz, 4),
                    "algorithmic_flops_per_step": sp["flops_update_oz"], "tasks_per_step": sp["oz_tasks"],
                    "share_of_update_ms": round(ms_oz / ms_fac, 3),
                    "note": "CUDA-event bracket around every launch of the class, kernels serialised (profiling mode)"}
        if roofline["traffic"] is None:
            roofline["traffic_note"] = ("no ncu --set full capture at this config's launch size (kernel replay would have to save / restore "
                                        "the 145 GB working set); captures of the same kernels on smaller configs: profiles/r01_k_oz_update2_ncu.md "
                                        "(128x128 two-pass tiles: 1.67x the algorithmic DRAM bytes), profiles/r01_k_oz_update_ncu.md (1.07x)")
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
    whole = {"ms": round(ms_tri, 4), "algorithmic_bytes": 16.0 * sp["nnzL"],
             "achieved": round(16.0 * sp["nnzL"] / (ms_tri * 1e-3) / 1e9, 1) if ms_tri > 0 else None,
             "frac": round(16.0 * sp["nnzL"] / (ms_tri * 1e-3) / 1e9 / peaks["hbm_gbs"], 4) if ms_tri > 0 else None,
             "note": "all sweep kernels of one solve (small / medium / below / big), CUDA-event brackets per launch"}
    roofline_solve = {"kernel": "k_fwd_big + k_bwd_big (dense sweeps over the big supernodes, one rhs)", "bound": "hbm",
                      "achieved": round(ach_big, 1) if ach_big else None, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                      "frac": round(ach_big / peaks["hbm_gbs"], 4) if ach_big else None,
                      "algorithmic_bytes": 16.0 * nnz_big, "ms": round(ms_big, 4), "launches_per_solve": int(cls["fwd_big"][1] + cls["bwd_big"][1]),
                      "streamed_tile_bytes": (bp["n_ftiles"] + bp["n_btiles"]) * 128 * 128 * 8,
                      "share_of_L": round(nnz_big / max(1.0, float(sp["nnzL"])), 4),
                      "whole_sweep": whole, "peak_source": peak_src}
    phases = {k: {"ms": round(v[0], 4), "launches": int(v[1])} for k, v in cls.items()}
    cfg_out = {"workload": workload_name(lp, sysname),
               "parallelism": "replicas only (this config's elimination tree does not shard)" if world > 1 else "single GPU",
               "l2": f"inputs larger than L2: factor panels {sp['nnzL_stored'] * 8 / 1e6:.0f} MB streamed every step (L2 126 MB); no explicit flush",
               "solves_per_step": round(nsolve, 3), "nnzL": sp["nnzL"], "factor_flops": sp["flops"], "nsuper": sp["nsuper"],
               "levels": sp["nlevels"], "setup_s": round(t_setup, 2), "bytes_device": sp["bytes_device"]}
    out = {
        "metric": METRIC, "value": round(world * K / (ms_total * 1e-3), 4), "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": round(ms_total / K, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg_out, "steps_real": int(k_real),
        "steps_note": (None if k_real >= K else f"the IPM converged after {len(recs)} iterations: the {K} timed steps cycle over the "
                       f"{k_real} real post-warm-up iterations"),
        "clocks": clocks,
        "e2e": {"value": round(world * K / e2e_time, 4), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": round(e2e_time * 1e3 / K, 4)},
        "gpu_launches": launches,
        "update_ms_host_api": round(float(np.mean([r["t_update"] for r in timed])) * 1e3, 3),
        "solve_ms_host_api": round(float(np.sum([r["t_solve"] for r in timed]) / np.sum([len(r["rhs"]) for r in timed])) * 1e3, 3),
        "ipm": ipm, "ipm_device_resident": ipm_dev,
        "roofline": roofline, "roofline_dmma": roofline_dmma, "roofline_solve": roofline_solve, "phases_one_step": phases,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del dev_recs, cache
        kkt.close()
        torch.cuda.empty_cache()
        cached = read_cpu_cache(cfg)
        if cached is not None:
            out["cpu_baseline"] = dict(cached["cpu_baseline"], source="measured by `bench.py --impl reference` on this box "
                                       f"{(time.time() - cached['when']) / 60:.0f} min earlier (cached in {cache_path(cfg)})")
            out["objective_parity"] = objective_parity(ipm, cached.get("ipm"))
        else:
            out["cpu_baseline"] = cpu_sample(pkg, lp, sysname, win[W], cfg)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


def sharded_arm(args, cfg, pkg, lp, sysname, sy, dist, rank, world, local):
    """BASELINE configs[3]: block-angular LP, elimination-tree subtrees sharded across the ranks + reduce of the separator
    front (tulip.jl_b200/parallel.py).  Every rank runs the same IPM (SPMD); timing = sum of the KKT calls of the timed
    iterations through the host API, max over ranks.  Strong scaling: the job is fixed; rank 0 first times the same
    workload on one GPU so that the line carries its own N = 1 figure."""
    import torch
    from tulip_jl_b200 import parallel
    K, W = args.steps, args.warmup
    m, n = lp.A.shape
    dev = torch.device(f"cuda:{local}")
    limit = max(args.ipm_limit, W + K) if args.ipm_limit > 0 else W + K
    n1 = None
    if rank == 0 and not args.no_n1:
        k1 = pkg.setup(lp.A, sy, pkg.Backend(device=local))
        h1, recs1 = record_hsd(pkg, lp, k1, limit)
        win1, kr1 = timed_window(recs1, W, K)
        t1 = sum(r["t_update"] + r["t_solve"] for r in win1[W:W + K])
        n1 = {"value": round(K / t1, 4), "unit": UNIT, "ms_per_step": round(t1 * 1e3 / K, 4), "steps_real": int(kr1),
              "update_ms_host_api": round(float(np.mean([r["t_update"] for r in win1[W:W + K]])) * 1e3, 3),
              "solve_ms_host_api": round(float(np.sum([r["t_solve"] for r in win1[W:W + K]]) / np.sum([len(r["rhs"]) for r in win1[W:W + K]])) * 1e3, 3),
              "ipm": ipm_summary(h1), "note": "same workload, same host API, one GPU (rank 0), timed in this run before the sharded solver"}
        n1["e2e_value"] = n1["value"]; n1["e2e_ms_per_step"] = n1["ms_per_step"]
        try:
            n1["ipm_device_resident"] = device_ipm_figures(pkg, k1, lp, limit)
            if n1["ipm_device_resident"].get("kkt_iter_per_s"):      # same convention as the sharded line: value = device loop, e2e = host API
                n1["value"] = n1["ipm_device_resident"]["kkt_iter_per_s"]
                n1["ms_per_step"] = n1["ipm_device_resident"]["kkt_ms_per_iter"]
        except Exception as e:
            n1["ipm_device_resident"] = {"error": f"{type(e).__name__}: {e}"}
        k1.close()
        del k1, recs1, win1
        torch.cuda.empty_cache()
    dist.barrier()
    t0 = time.time()
    kkt = parallel.DistB200KKT(lp.A, sy, pkg.Backend(device=local))
    t_setup = time.time() - t0
    sampler = ClockSampler(local)
    dist.barrier(); torch.cuda.synchronize()
    sampler.start()
    h, recs = record_hsd(pkg, lp, kkt, limit)
    torch.cuda.synchronize(); dist.barrier()
    clocks = sampler.stop()
    win, k_real = timed_window(recs, W, K)
    timed = win[W:W + K]
    t = torch.tensor([sum(r["t_update"] + r["t_solve"] for r in timed)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = float(t[0])
    nsolve = float(np.mean([len(r["rhs"]) for r in timed]))
    st = kkt.stats()
    comm = kkt.comm_profile() if hasattr(kkt, "comm_profile") else None
    val = K / tot
    out = {"metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": round(tot * 1e3 / K, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(lp, sysname),
                      "parallelism": f"elimination-tree subtrees sharded over {world} ranks, separator part replicated; "
                                     + (kkt.describe() if hasattr(kkt, "describe") else "NCCL all-reduce between the phases"),
                      "l2": "timed through the host-pointer API (H2D/D2H inside); factor panels exceed L2",
                      "solves_per_step": round(nsolve, 3), "nnzL": st["nnzL"], "factor_flops": st["flops"], "setup_s": round(t_setup, 2),
                      "bytes_device": st["bytes_device"]},
           "steps_real": int(k_real), "clocks": clocks, "comm_nranks_seen": [world],
           "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": int((2 * n + m) * 8 + nsolve * (n + m) * 8),
                   "d2h_bytes_per_step": int(4 + nsolve * (n + m) * 8), "ms_per_step": round(tot * 1e3 / K, 4)},
           "gpu_launches": int(K * (st["launches_update"] + nsolve * st["launches_solve"])),
           "ipm": ipm_summary(h), "n1_same_workload": n1, "comm": comm,
           "roofline": {"bound": "tensor", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                        "note": "per-kernel roofline is reported by the single-GPU run (--gpus 1)"},
           "update_ms_host_api": round(float(np.mean([r["t_update"] for r in timed])) * 1e3, 3),
           "solve_ms_host_api": round(float(np.sum([r["t_solve"] for r in timed]) / np.sum([len(r["rhs"]) for r in timed])) * 1e3, 3)}
    def reduce_max(v):
        tt = torch.tensor(v, dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return [float(x) for x in tt]

    # `value` (inputs resident in HBM): the same sharded solver under the device-resident IPM loop (tlpb200_hsd_*): no host
    # vector work between the KKT calls, so the ranks stay in lock-step; CUDA-event time of the update!/solve! sequences, max
    # over ranks.  `e2e` stays the host-pointer API, which also contains the skew of N Python IPM drivers waiting for each
    # other inside the collectives.
    out["e2e_strong_speedup_vs_n1"] = round(val / n1["e2e_value"], 3) if n1 is not None else None
    try:
        dv = device_ipm_figures(pkg, kkt.local, lp, limit, reduce_max)
        out["ipm_device_resident"] = dv
        if dv.get("kkt_iter_per_s"):
            out["value"] = dv["kkt_iter_per_s"]
            out["ms_per_step"] = dv["kkt_ms_per_iter"]
            out["value_note"] = ("value = KKT iterations/s of the device-resident HSD loop on the sharded solver (inputs resident in HBM, "
                                 f"CUDA events, max over ranks, {dv['iters']} real iterations to {dv['status']}); e2e = host-pointer API")
            n1d = (n1 or {}).get("ipm_device_resident") or {}
            if n1d.get("kkt_iter_per_s"):
                out["strong_speedup_vs_n1"] = round(dv["kkt_iter_per_s"] / n1d["kkt_iter_per_s"], 3)
                out["limiter"] = ("per-level kernel chains of the factorisation and of the sweeps, whose length does not depend on N, "
                                  "plus the replicated separator part and the full-vector rhs / recovery on every rank; the collectives "
                                  f"cost {comm['per_update_ms']:.3f} ms per update! and {comm['per_solve_ms']:.3f} ms per solve!" if comm else None)
    except Exception as e:
        out["ipm_device_resident"] = {"error": f"{type(e).__name__}: {e}"}
    kkt.close()
    # weak-scaling variant of the same generator: 64 blocks PER RANK (the BASELINE config is the fixed 64-block LP above;
    # this line shows what the sharding does when the tree has work for every GPU)
    if not args.no_weak:
        try:
            from tulip_jl_b200 import lpgen
            lpw = lpgen.block_angular(blocks=64 * world, name=f"cfg4_weak_x{world}")
            kw = parallel.DistB200KKT(lpw.A, sy, pkg.Backend(device=local))
            dist.barrier()
            hw, recw = record_hsd(pkg, lpw, kw, W + K)
            winw, krw = timed_window(recw, W, K)
            tw = reduce_max([sum(r["t_update"] + r["t_solve"] for r in winw[W:W + K])])[0]
            stw = kw.stats()
            out["weak_variant"] = {"workload": workload_name(lpw, sysname), "blocks": 64 * world, "ms_per_step": round(tw * 1e3 / K, 4),
                                   "value": round(K / tw, 4), "nnzL": stw["nnzL"], "factor_flops": stw["flops"],
                                   "update_ms_host_api": round(float(np.mean([r["t_update"] for r in winw[W:W + K]])) * 1e3, 3),
                                   "solve_ms_host_api": round(float(np.sum([r["t_solve"] for r in winw[W:W + K]]) / np.sum([len(r["rhs"]) for r in winw[W:W + K]])) * 1e3, 3),
                                   "weak_efficiency_vs_n1": round(n1["e2e_ms_per_step"] / (tw * 1e3 / K), 3) if n1 is not None else None,
                                   "note": "per-rank work fixed (64 blocks per GPU + the 512 linking rows), host-pointer API, max over ranks; "
                                           "efficiency = N=1 ms/step on the 64-block LP / this ms/step"}
            kw.close()
        except Exception as e:
            out["weak_variant"] = {"error": f"{type(e).__name__}: {e}"}
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        emit(out)


def cpu_sample_extrapolated(st, nsolve, why):
    """Fallback CPU sample for a dense-dominated config whose CPU factorisation takes minutes (config T) when the real
    CPU arm has not run on this box: time the two dense kernels that carry > 99 % of the CPU port's work on this workload --
    LAPACK dpotrf and the triangular solves, SciPy's OpenBLAS, all host cores -- on an n0 x n0 block, and scale by the
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
            "sample": f"EXTRAPOLATED ({why}): dense dpotrf {n0}^2 in {t_f:.1f} s = {rate / 1e9:.0f} GF/s (c_j^2 convention) and "
                      f"forward+backward solve at {bw / 1e9:.1f} GB/s, scaled to sum c_j^2 = {st['flops']:.3e} flops + {nsolve:.2f} solves "
                      f"x 16 nnz(L) = {16.0 * st['nnzL'] / 1e9:.1f} GB",
            "label": "CPU port kernels (SciPy-OpenBLAS dpotrf / dtrsv, all host cores) -- NOT Tulip/CHOLMOD; not run to completion"}


def cpu_sample(pkg, lp, sysname, rec, cfg=""):
    """oracle CPU port on a bounded sample: ONE recorded IPM iteration (1 update! + its solves)."""
    sysname = CPU_SYSTEM.get(cfg, sysname)
    from oracle import cpu_kkt
    A = lp.A
    m, n = A.shape
    sy = pkg.K1() if sysname == "K1" else pkg.K2()
    an = pkg.setup(A, sy, pkg.Backend(analyze_only=True, dense_col_threshold=-1))
    cores = os.cpu_count() or 1
    st = an.stats()
    if st["flops"] > ONE_ITER_ABOVE_FLOPS:
        return cpu_sample_extrapolated(st, float(len(rec["rhs"])), "one CPU factorisation of this config takes minutes; run "
                                       "`bench.py --impl reference` first on the same box for the measured figure")
    ck = cpu_kkt.CpuSupernodalKKT(A, sysname, nthreads=cores, symbolic_from=an)
    t0 = time.perf_counter()
    ck.update(rec["theta"], rec["regP"], rec["regD"])
    dx = np.zeros(n); dy = np.zeros(m)
    for xp, xd in rec["rhs"]:
        ck.solve(dx, dy, xp, xd)
    dt = time.perf_counter() - t0
    return {"value": round(1.0 / dt, 5), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 recorded HSD iteration of the same workload (1 update! + {len(rec['rhs'])} solve!), {dt:.2f} s",
            "label": CPU_LABEL + (f"; CPU system {sysname}" if cfg in CPU_SYSTEM else "")}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cfg = args.config
    if cfg == "auto":
        cfg = "T" if world == 1 else "4"
    pkg, lp, sysname = build_lp(cfg)
    gpu_sysname = sysname
    sysname = CPU_SYSTEM.get(cfg, sysname)
    from oracle import cpu_kkt, hsd_ref
    # the reference runs its own vector work with BLAS threads = 1 (model.jl:73); the factorisation's BLAS
    # (SciPy's OpenBLAS, a separate library instance) gets every core below.
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=1, user_api="blas")
    A = lp.A
    sy = pkg.K1() if sysname == "K1" else pkg.K2()
    an = pkg.setup(A, sy, pkg.Backend(analyze_only=True, dense_col_threshold=-1))       # integer analysis only, no device
    cores = os.cpu_count() or 1
    st = an.stats()
    K, W = args.steps, args.warmup
    one_iter = st["flops"] > ONE_ITER_ABOVE_FLOPS
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": world, "higher_is_better": True,
            "scaling": "strong" if (world > 1 and cfg.startswith("4")) else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    if one_iter:
        # panels + the largest descendant-update buffer of the left-looking factorisation (oracle/cpu_supernodal.c)
        need = 8.0 * (st["nnzL_stored"] + float(st["max_nrow"]) ** 2) * 1.05
        try:
            import psutil
            avail = float(psutil.virtual_memory().available)
        except Exception:
            avail = float("inf")
        if avail < need:
            cb = cpu_sample_extrapolated(st, 5.0, f"host RAM {avail / 1e9:.0f} GB < {need / 1e9:.0f} GB needed by the CPU port for this config")
            out = dict(base, value=cb["value"], steps=0, warmup=0, ms_per_step=round(1e3 / cb["value"], 1),
                       config={"workload": workload_name(lp, gpu_sysname), "solves_per_step": 5.0}, cpu_baseline=cb,
                       e2e={"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
            emit(out)
            return
        K, W = 1, 0
    ck = cpu_kkt.CpuSupernodalKKT(A, sysname, nthreads=cores, symbolic_from=an)
    marks = []

    class T:
        def update(self, *a):
            ck.update(*a)

        def solve(self, *a):
            ck.solve(*a)

    dat = hsd_ref.IPMData(A, lp.b, True, lp.c, 0.0, lp.l, lp.u)
    limit = 1 if one_iter else (max(args.ipm_limit, W + K) if args.ipm_limit > 0 else W + K)
    hs = hsd_ref.HSDRef(dat, T(), hsd_ref.IPMOptions(IterationsLimit=limit))
    hs.optimize(callback=lambda s: marks.append((s.t_factor + s.t_solve, s.n_solve)))
    if len(marks) <= W:
        raise SystemExit("reference arm: IPM ended during warm-up")
    t_w, ns_w = marks[W - 1] if W > 0 else (0.0, 0)
    hi = min(len(marks), W + K)
    t_e, ns_e = marks[hi - 1]
    k_done = hi - W
    dt = t_e - t_w
    val = k_done / dt
    nsolve = (ns_e - ns_w) / k_done
    ipm = ipm_summary(hs)
    cb = {"value": round(val, 6), "unit": UNIT, "cores": cores, "kind": "port",
          "sample": (f"ONE real HSD iteration of the same workload from the reference's start point (1 update! + {ns_e - ns_w} solve!), "
                     f"{dt:.1f} s of KKT time; not extrapolated" if one_iter else
                     f"{k_done} HSD iterations (after {W} warm-up) of the same workload, {dt:.3g} s of KKT time"),
          "label": CPU_LABEL + (f"; CPU system {sysname} (Tulip's default; the reference's K1 would form a dense A D A' here)" if cfg in CPU_SYSTEM else "")}
    out = dict(base, value=round(val, 6), steps=k_done, warmup=W, ms_per_step=round(dt * 1e3 / k_done, 3),
               config={"workload": workload_name(lp, gpu_sysname), "solves_per_step": round(nsolve, 3),
                       "cpu_system": sysname}, cpu_baseline=cb, ipm=ipm,
               e2e={"value": round(val, 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    try:
        json.dump({"host": socket.gethostname(), "when": time.time(), "cpu_baseline": cb, "ipm": ipm}, open(cache_path(cfg), "w"))
    except Exception as e:          # pragma: no cover
        log(f"could not cache the CPU arm's result: {e}")
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="auto", help="auto (T at N=1, 4 sharded at N>1) | T | 2 | 3 | 4 | 5 | mini | 3mini | 4mini | 5mini | Tmini")
    ap.add_argument("--ipm-limit", type=int, default=100, help="IPM IterationsLimit (reference default 100); 0 = stop after warmup+steps iterations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ipm-device", default="auto", choices=["auto", "on", "off"],
                    help="also run the device-resident HSD loop (auto: every config whose factorisation is below 5e12 flops, i.e. not T)")
    ap.add_argument("--no-weak", action="store_true", help="sharded run: skip the weak-scaling variant (64 blocks per rank)")
    ap.add_argument("--no-n1", action="store_true", help="sharded run: skip the single-GPU timing of the same workload on rank 0")
    args = ap.parse_args()
    claim_stdout()
    if args.warmup < 3 and args.impl == "b200":
        log("note: timing rules ask for >= 3 warm-up steps")
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
