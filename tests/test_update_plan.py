"""CPU test of the update-task plan (integer work, bit-exact): inside a supernode that uses the tcgen05 path, every
contribution  C[i, k] -= L[i, piece j] L[k, piece j]'  (k in a later column block, i >= k over all panel rows) must be
applied exactly once -- by an FP64 tile task (column blocks j+1, j+2) or by a tcgen05 task whose K range contains piece j
(column blocks >= j+3).  Reference: the supernodal update loop of cholesky! (/root/reference/src/KKT/Cholmod/spd.jl:46)."""
import numpy as np
import pytest

import tlpb200_loader

pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen  # noqa: E402


@pytest.mark.parametrize("tile", [64, 128, 0])
def test_fp64_tiles_and_tcgen05_tasks_cover_every_contribution_once(tile, monkeypatch):
    monkeypatch.setenv("TLPB200_OZAKI_TILE", str(tile))
    monkeypatch.setenv("TLPB200_OZAKI_KSPLIT", "256")           # several K splits even on this small problem
    lp = lpgen.random_sparse(1800, 3600, 8, seed=424242, name="oz_medium")
    k = pkg.setup(lp.A, pkg.K1(), pkg.Backend(analyze_only=True, ozaki_ncol=512))
    plan = k.update_plan()
    sym = k.symbolic()
    assert len(plan["views"]) >= 1 and len(plan["oz"]) > 0
    pieces = plan["pieces"]
    for sn, nrb, ncb, _base in plan["views"]:
        f = int(sym["sn_first"][sn]); nc = int(sym["sn_first"][sn + 1]) - f
        nrow = int(sym["sn_rowptr"][sn + 1] - sym["sn_rowptr"][sn])
        assert nrb == -(-nrow // 128) and ncb == -(-nc // 128)
        view_idx = int(np.where(plan["views"][:, 0] == sn)[0][0])
        mine = [p for p in range(len(pieces)) if pieces[p, 0] == sn]
        for p in mine:
            j = (int(pieces[p, 1]) - f) // 128
            c1 = int(pieces[p, 2]) - f
            dm = np.zeros((nrow, nc), np.int16)
            for name in ("upd", "upd128"):
                T = plan[name]
                for piece, i0, ni, k0, nk, tgt, diag, _ in T[(T[:, 0] == p) & (T[:, 5] == sn)]:
                    blk = np.ones((ni, nk), np.int16)
                    if diag:
                        blk = np.tril(blk)
                    dm[i0:i0 + ni, k0:k0 + nk] += blk
            want = np.zeros((nrow, nc), np.int16)
            ii, kk = np.meshgrid(np.arange(nrow), np.arange(nc), indexing="ij")
            want[(kk >= c1) & (ii >= kk)] = 1
            oz = plan["oz"][plan["oz"][:, 0] == view_idx]
            for g in range(4 * j, 4 * j + 4):                      # the four 32-column K chunks of the piece
                cnt = dm.copy()
                for view, rbA, rbB, half, k0, k1, _, _ in oz[(oz[:, 4] <= g) & (g < oz[:, 5])]:
                    # half tiles are 64 wide; a column block whose tasks all have half == 0 is covered by 128-wide tasks
                    wide = not np.any((oz[:, 2] == rbB) & (oz[:, 3] == 1)) and (tile != 64)
                    cols = np.arange(rbB * 128 + half * 64, min(nc, rbB * 128 + (128 if wide else half * 64 + 64)))
                    rows = np.arange(rbA * 128, min(nrow, rbA * 128 + 128))
                    cnt[np.ix_(rows, cols)] += (rows[:, None] >= cols[None, :]).astype(np.int16)
                assert np.array_equal(cnt, want), (sn, p, j, g)
