"""torchrun target: multi-GPU sharded KKT solve vs the oracle (and timing), NCCL backend.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py [mini|full]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import tlpb200_loader

pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen, parallel  # noqa: E402
from oracle import kkt_ref  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    which = sys.argv[1] if len(sys.argv) > 1 else "mini"
    ok = True
    modes = [m_ for m_ in os.environ.get("TLPB200_DIST_MODES", "library,phases").split(",") if m_]
    for sysname, mode in [(s_, m_) for s_ in ("K1", "K2") for m_ in modes]:
        lp = lpgen.config(4, mini=(which == "mini"))
        A = lp.A
        m, n = A.shape
        rng = np.random.default_rng(23)
        theta = np.exp(rng.uniform(-4, 4, n)); regP = np.full(n, 1e-6); regD = np.full(m, 1e-6)
        xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
        sy = pkg.K1() if sysname == "K1" else pkg.K2()
        k = parallel.DistB200KKT(A, sy, pkg.Backend(device=local), mode=mode)
        owner, off, cnt = k.dist_info()
        k.update(theta, regP, regD)
        dx = np.zeros(n); dy = np.zeros(m)
        k.solve(dx, dy, xi_p, xi_d)
        # timing (device work + collectives + host copies), max over ranks
        t = []
        for _ in range(3):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter(); k.update(theta, regP, regD); t1 = time.perf_counter()
            k.solve(dx, dy, xi_p, xi_d); t2 = time.perf_counter()
            t.append((t1 - t0, t2 - t1))
        tt = torch.tensor(min(t), dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, dx, dy, xi_p, xi_d)
        msg = f"[{sysname}/{mode}] world={world} m={m} n={n} top={cnt * 8 / 1e6:.2f} MB shards={np.bincount(owner[owner >= 0], minlength=world).tolist()} " \
              f"update {tt[0].item() * 1e3:.2f} ms solve {tt[1].item() * 1e3:.2f} ms |rp|={rp:.2e} |rd|={rd:.2e}"
        scale = max(1.0, np.abs(dx).max(), np.abs(dy).max())
        good = rp <= 1.5e-8 * scale and rd <= 1.5e-8 * scale
        if which == "mini":
            o = kkt_ref.SparseK1(A) if sysname == "K1" else kkt_ref.SparseK2(A)
            o.update(theta, regP, regD)
            dx0 = np.zeros(n); dy0 = np.zeros(m); o.solve(dx0, dy0, xi_p, xi_d)
            ex = np.abs(dx - dx0).max() / np.abs(dx0).max(); ey = np.abs(dy - dy0).max() / np.abs(dy0).max()
            msg += f" relerr dx={ex:.2e} dy={ey:.2e}"
            good = good and ex < 1e-7 and ey < 1e-7
        # every rank must hold the same solution
        chk = torch.tensor([float(np.abs(dx).sum()), float(np.abs(dy).sum())], dtype=torch.float64, device=f"cuda:{local}")
        lo = chk.clone(); hi = chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        good = good and bool(torch.all((hi - lo) <= 1e-9 * hi.abs().clamp_min(1e-300)).item())
        # PosDefException must surface on every rank
        try:
            k.update(theta, regP, -1e3 * np.ones(m))
            raised = False
        except pkg.PosDefException:
            raised = True
        good = good and raised
        k.update(theta, regP, regD)       # still usable
        if mode == "library":
            cp = k.comm_profile(20)
            msg += f" | collectives/update {cp['per_update_ms']:.3f} ms, /solve {cp['per_solve_ms']:.3f} ms"
        if rank == 0:
            print(msg, "posdef-consistent" if raised else "POSDEF-NOT-RAISED", flush=True)
        k.close()
        ok = ok and good
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST-GPU-OK" if int(flag.item()) == 1 else "DIST-GPU-FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
