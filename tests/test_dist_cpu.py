"""CPU tests of the multi-GPU (N > 1) path's host logic:

* the subtree partition is deterministic, ancestor-closed and gives independent shards;
* the sharded algorithm (partial separator Schur complements -> all-reduce(sum) -> redundant top
  factorisation; forward -> all-reduce -> top -> backward -> all-reduce) reproduces the oracle's solution,
  run with world_size = 2 over the `gloo` backend (dense NumPy stands in for the device kernels here).
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import tlpb200_loader

pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _partition(A, system, nranks, rank=0):
    k = pkg.setup(A, system, pkg.Backend(analyze_only=True, rank=rank, nranks=nranks))
    owner, top_off, top_cnt = k.dist_info()
    return k, owner, top_off, top_cnt


@pytest.mark.parametrize("nranks", [2, 4])
@pytest.mark.parametrize("sysname", ["K1", "K2"])
def test_partition_properties(nranks, sysname):
    lp = lpgen.config(4, mini=True)
    sy = pkg.K1() if sysname == "K1" else pkg.K2()
    k0, owner0, off0, cnt0 = _partition(lp.A, sy, nranks, rank=0)
    k1, owner1, off1, cnt1 = _partition(lp.A, sy, nranks, rank=nranks - 1)
    assert np.array_equal(owner0, owner1) and off0 == off1 and cnt0 == cnt1      # same plan on every rank
    assert set(np.unique(owner0)).issubset(set(range(-1, nranks)))
    sym = k0.symbolic()
    first, rp, rows = sym["sn_first"], sym["sn_rowptr"], sym["sn_rows"]
    ns = len(first) - 1
    col2sn = np.repeat(np.arange(ns), np.diff(first))
    work = np.zeros(nranks)
    for s in range(ns):
        below = rows[rp[s] + (first[s + 1] - first[s]):rp[s + 1]]
        tg = np.unique(owner0[col2sn[below]]) if len(below) else np.array([], int)
        if owner0[s] == -1:
            assert np.all(tg == -1), "the top part must be ancestor-closed"
        else:
            assert set(tg.tolist()).issubset({owner0[s], -1}), "a subtree may only update itself or the top part"
            work[owner0[s]] += float((sym["colcount"][first[s]:first[s + 1]].astype(float) ** 2).sum())
    assert work.min() > 0, "every rank gets work on the block-angular config"
    assert work.max() <= 0.75 * work.sum(), "grossly unbalanced partition"
    st = k0.stats()
    assert off0 + cnt0 == st["nnzL_stored"] and cnt0 > 0


def test_single_rank_is_unsharded():
    lp = lpgen.config(4, mini=True)
    _, owner, off, cnt = _partition(lp.A, pkg.K1(), 1)
    assert np.all(owner == 0) and cnt == 0


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
import tlpb200_loader
pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen
from oracle import kkt_ref
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lp = lpgen.config(4, mini=True)
A = lp.A; m, n = A.shape
rng = np.random.default_rng(17)
theta = np.exp(rng.uniform(-3, 3, n)); regP = np.full(n, 1e-5); regD = np.full(m, 1e-5)
xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
k = pkg.setup(A, pkg.K1(), pkg.Backend(analyze_only=True, rank=rank, nranks=world))
owner, _, _ = k.dist_info()
sym = k.symbolic(); perm = sym["perm"]; first = sym["sn_first"]
col_owner = np.repeat(owner, np.diff(first))
D = 1.0 / (theta + regP)
K = (A @ sp.diags(D) @ A.T + sp.diags(regD)).toarray()[np.ix_(perm, perm)]          # spd.jl:43, permuted
b = (xi_p + A @ (D * xi_d))[perm]
T = np.nonzero(col_owner == -1)[0]; S = np.nonzero(col_owner == rank)[0]
others = np.nonzero((col_owner != -1) & (col_owner != rank))[0]
assert np.abs(K[np.ix_(S, others)]).max() == 0.0, "shards are not independent"
# update!: partial separator Schur complement of this rank, all-reduce(sum), redundant top factor
Kss = K[np.ix_(S, S)]; Kts = K[np.ix_(T, S)]
part = -Kts @ np.linalg.solve(Kss, Kts.T)
if rank == 0: part = part + K[np.ix_(T, T)]
part_t = torch.from_numpy(part); dist.all_reduce(part_t)
Lt = np.linalg.cholesky(part_t.numpy())
# solve!: forward on own shard, all-reduce, top, backward on own shard, all-reduce
wk = np.zeros(len(perm))
ys = np.linalg.solve(Kss, b[S])
wk[T] = -Kts @ ys + (b[T] if rank == 0 else 0.0)
wk_t = torch.from_numpy(wk); dist.all_reduce(wk_t)
xt = np.linalg.solve(Lt.T, np.linalg.solve(Lt, wk_t.numpy()[T]))
xs = ys - np.linalg.solve(Kss, Kts.T @ xt)
out = np.zeros(len(perm)); out[S] = xs
if rank == 0: out[T] = xt
out_t = torch.from_numpy(out); dist.all_reduce(out_t)
dy = np.empty(m); dy[perm] = out_t.numpy()
dx = D * (A.T @ dy - xi_d)
o = kkt_ref.SparseK1(A); o.update(theta, regP, regD)
dx0 = np.zeros(n); dy0 = np.zeros(m); o.solve(dx0, dy0, xi_p, xi_d)
ex = np.abs(dx - dx0).max() / np.abs(dx0).max(); ey = np.abs(dy - dy0).max() / np.abs(dy0).max()
assert ex < 1e-8 and ey < 1e-8, (ex, ey)
dist.barrier()
if rank == 0: print("DIST-CPU-OK", ex, ey)
dist.destroy_process_group()
"""


def test_sharded_algorithm_world2_gloo(tmp_path):
    import subprocess
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST-CPU-OK" in r.stdout
