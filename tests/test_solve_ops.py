"""CPU test of the merged-level launch sequences of the triangular sweeps (csrc/plan.cpp SolveOp; kernels k_fwd_large /
k_bwd_large with merged = 1): a replay of the persistent kernels' schedule -- G CTAs walking a merged op's items with a grid
stride, an item finishing only when the flags / counters it waits on are set -- must terminate (no dead-lock for any G) and
must respect the elimination-tree order: a supernode's right-hand side is read only after every contribution to it has been
added (forward), its ancestors' solution only after it is final (backward)."""
import numpy as np
import pytest

import tlpb200_loader

pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen  # noqa: E402

CASES = [("cfg3-mini", lambda: lpgen.staircase(stages=12, nodes=200, arcs=320, name="s"), "K2"),
         ("cfg4-mini", lambda: lpgen.block_angular(blocks=6, mb=400, nb=800, width=64, link=150, name="b"), "K1"),
         ("cfg2-mini", lambda: lpgen.random_sparse(900, 1800, 6, name="r"), "K1"),
         ("tall-border", lambda: lpgen.block_angular(blocks=4, mb=600, nb=1200, width=96, link=500, name="b2"), "K1")]


def _replay(items, begin, end, G, ready_fn, done_fn):
    """G CTAs, CTA g owns items begin+g, begin+g+G, ...; returns False on dead-lock"""
    pos = [begin + g for g in range(G)]
    progressed = True
    while progressed:
        progressed = False
        for g in range(G):
            while pos[g] < end and ready_fn(pos[g]):
                done_fn(pos[g])
                pos[g] += G
                progressed = True
    return all(p >= end for p in pos)


@pytest.mark.parametrize("name,gen,sysname", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("G", [1, 3, 148])
@pytest.mark.parametrize("ncol", [10 ** 9, 200])          # without / with dense-solve ("big") supernodes breaking the runs
def test_merged_sweeps_replay(name, gen, sysname, G, ncol):
    A = gen().A
    k = pkg.setup(A, pkg.K1() if sysname == "K1" else pkg.K2(), pkg.Backend(analyze_only=True, dense_solve_ncol=ncol))
    ops = k.solve_ops()
    sym = k.symbolic()
    first, rp, rows = sym["sn_first"], sym["sn_rowptr"], sym["sn_rows"]
    ns = len(first) - 1
    col2sn = np.repeat(np.arange(ns), np.diff(first))
    par = ops["sn_parent"]
    fit, bit = ops["fwd_items"], ops["bwd_seq"]
    assert len(fit) > 0, "the case must have block-solve supernodes"
    # every item is covered exactly once by the merged ops, and merged ops really span several levels somewhere
    cover = np.zeros(len(fit), int)
    for kind, b, e, lvl in ops["fwd_ops"]:
        if kind == 1:
            cover[b:e] += 1
    assert np.all(cover == 1)
    assert np.array_equal(np.bincount(bit["sn"][bit["kind"] == 2], minlength=ns), ops["bwd_nbelow"])
    cover = np.zeros(len(bit), int)
    for kind, b, e, lvl in ops["bwd_ops"]:
        if kind == 1:
            cover[b:e] += 1
    assert np.all(cover == 1)
    nlev_items = len({(int(i["sn"])) for i in fit})
    assert sum(1 for o in ops["fwd_ops"] if o[0] == 1) < nlev_items or nlev_items <= 1
    ncb = lambda s: (first[s + 1] - first[s] + 127) // 128
    # ---- forward replay -------------------------------------------------------------------------------------------------
    done_sn = np.zeros(ns, bool)          # supernode fully processed (any kind of op)
    fin_items = np.zeros(ns, int)         # finished forward items per supernode
    n_items = np.zeros(ns, int)
    for it in fit:
        n_items[it["sn"]] += 1
    flag = {}
    cnt = np.zeros(ns, int)
    children = [[] for _ in range(ns)]
    for s in range(ns):
        if par[s] >= 0:
            children[par[s]].append(s)
    small_list = k.update_plan()["small_list"]
    big_fwd = k.big_plan()["fwd"]
    for kind, b, e, lvl in ops["fwd_ops"]:
        if kind == 0:                      # one-CTA supernodes of one level: children must be complete, then they are
            for s in small_list[b:e]:
                assert all(done_sn[c] for c in children[s]), ("small supernode before its child", s)
            done_sn[small_list[b:e]] = True
            continue
        if kind == 2:
            for s in np.unique(big_fwd["sn"][b:e]):
                assert all(done_sn[c] for c in children[s]), ("dense-solve supernode before its child", s)
                done_sn[s] = True
            continue

        def ready(x):
            it = fit[x]; s = int(it["sn"])
            if it["kind"] == 0:
                if cnt[s] < ops["fwd_need"][s]:
                    return False
                return all(flag.get((s, j), False) for j in range(int(it["blk"])))
            return all(flag.get((s, j), False) for j in range(ncb(s)))

        def done(x):
            it = fit[x]; s = int(it["sn"])
            if it["kind"] == 0:
                # the right-hand side is read here: every child must be complete (in-launch: all its items; else: earlier launch)
                for c in children[s]:
                    assert done_sn[c] or fin_items[c] == n_items[c] > 0, ("rhs read before child finished", s, c)
                flag[(s, int(it["blk"]))] = True
            fin_items[s] += 1
            if ops["fwd_parent"][s] >= 0:
                cnt[ops["fwd_parent"][s]] += 1
        assert _replay(fit, b, e, G, ready, done), f"forward dead-lock in op [{b},{e}) with {G} CTAs"
        for x in range(b, e):
            if fin_items[fit[x]["sn"]] == n_items[fit[x]["sn"]]:
                done_sn[fit[x]["sn"]] = True
    assert done_sn.all(), "a supernode was never processed by the forward sweep"
    # ---- backward replay ------------------------------------------------------------------------------------------------
    bflag = {}
    bdone = np.zeros(ns, int)
    bdone_sn = np.zeros(ns, bool)
    below_done = np.zeros(ns, int)
    big_bwd = k.big_plan()["bwd"]
    for kind, b, e, lvl in ops["bwd_ops"]:
        if kind == 3:
            continue                       # below items only read ancestors' x: checked through the ops' level order below
        if kind == 0:
            for s in small_list[b:e]:
                assert par[s] < 0 or bdone_sn[par[s]], ("small supernode before its parent (backward)", s)
            bdone_sn[small_list[b:e]] = True
            continue
        if kind == 2:
            for s in np.unique(big_bwd["sn"][b:e]):
                assert par[s] < 0 or bdone_sn[par[s]], ("dense-solve supernode before its parent (backward)", s)
                bdone_sn[s] = True
            continue

        def ready(x):
            it = bit[x]; s = int(it["sn"])
            w = ops["bwd_wait"][s]
            if w >= 0 and bdone[w] < ops["bwd_nitems"][w]:
                return False
            if it["kind"] == 2:
                return True
            if below_done[s] < ops["bwd_nbelow"][s]:
                return False
            return all(bflag.get((s, j), False) for j in range(int(it["blk"]) + 1, ncb(s)))

        def done(x):
            it = bit[x]; s = int(it["sn"])
            p = par[s]
            if it["kind"] == 2:            # rows below the columns: reads the ancestors' solution
                assert p < 0 or bdone_sn[p] or (ops["bwd_nitems"][p] > 0 and bdone[p] == ops["bwd_nitems"][p]), ("below item early", s, p)
                below_done[s] += 1
                return
            if p >= 0:                     # the ancestors' solution is read here: the parent must be final
                assert bdone_sn[p] or (ops["bwd_nitems"][p] > 0 and bdone[p] == ops["bwd_nitems"][p]), ("ancestor solution read early", s, p)
            bflag[(s, int(it["blk"]))] = True
            bdone[s] += 1
        assert _replay(bit, b, e, G, ready, done), f"backward dead-lock in op [{b},{e}) with {G} CTAs"
        for x in range(b, e):
            if bdone[bit[x]["sn"]] == ops["bwd_nitems"][bit[x]["sn"]]:
                bdone_sn[bit[x]["sn"]] = True
    assert np.array_equal(bdone, ops["bwd_nitems"]) and bdone_sn.all() and np.array_equal(below_done, ops["bwd_nbelow"])
    # a below op of level L sits after every op of a higher level in the sequence
    seen_lower = -1
    for kind, b, e, lvl in ops["bwd_ops"]:
        if kind == 3:
            assert all(not (k2 in (0, 2) and l2 > lvl) for k2, _, _, l2 in ops["bwd_ops"][list(map(tuple, ops["bwd_ops"])).index((kind, b, e, lvl)) + 1:])


@pytest.mark.parametrize("nranks", [2, 4])
@pytest.mark.parametrize("G", [1, 148])
def test_merged_sweeps_replay_inside_sharded_phases(nranks, G):
    """N > 1: every rank walks the same merged launch sequence twice -- phase 0 on the subtrees it owns, phase 1 on the
    replicated top part -- with per-phase dependency targets (tlpb200_debug_phase_deps).  Replay of the forward order
    (phase 0, then phase 1) and of the backward order (phase 1, then phase 0) on every rank: no dead-lock, a supernode's
    right-hand side is read only after ALL its children finished (in this phase or an earlier one), an ancestor's solution only
    after it is final, and a top supernode's counter is never touched by a subtree child (it runs in another phase)."""
    lp = lpgen.block_angular(blocks=8, mb=300, nb=600, width=64, link=200, name="b")
    for rank in range(nranks):
        k = pkg.setup(lp.A, pkg.K1(), pkg.Backend(analyze_only=True, rank=rank, nranks=nranks))
        ops = k.solve_ops()
        owner, _, _ = k.dist_info()
        sym = k.symbolic()
        first = sym["sn_first"]
        ns = len(first) - 1
        par = ops["sn_parent"]
        fit, bit = ops["fwd_items"], ops["bwd_seq"]
        small_list = k.update_plan()["small_list"]
        bp = k.big_plan()
        ncb = lambda s: (first[s + 1] - first[s] + 127) // 128
        children = [[] for _ in range(ns)]
        for s in range(ns):
            if par[s] >= 0:
                children[par[s]].append(s)
        live = [owner == rank, owner == -1]
        assert live[0].any() and live[1].any()
        n_items = np.bincount(fit["sn"], minlength=ns)
        # ---- forward: phase 0 then phase 1 ------------------------------------------------------------------------
        done_sn = np.zeros(ns, bool)
        cnt = np.zeros(ns, int)
        fin = np.zeros(ns, int)
        flag = {}
        for ph in (0, 1):
            need, fpar, _ = k.phase_deps(ph)
            assert np.all(need[~live[ph]] == 0) and np.all(fpar[~live[ph]] == -1)
            for kind, b, e, lvl in ops["fwd_ops"]:
                if kind == 0:
                    for s in small_list[b:e]:
                        if live[ph][s]:
                            assert all(done_sn[c] for c in children[s] if owner[c] in (rank, -1))
                            done_sn[s] = True
                    continue
                if kind == 2:
                    for s in np.unique(bp["fwd"]["sn"][b:e]):
                        if live[ph][s]:
                            assert all(done_sn[c] for c in children[s] if owner[c] in (rank, -1))
                            done_sn[s] = True
                    continue

                def ready(x):
                    it = fit[x]; s = int(it["sn"])
                    if not live[ph][s]:
                        return True                      # skipped item: `continue`
                    if it["kind"] == 0:
                        return cnt[s] >= need[s] and all(flag.get((s, j), False) for j in range(int(it["blk"])))
                    return all(flag.get((s, j), False) for j in range(ncb(s)))

                def done(x):
                    it = fit[x]; s = int(it["sn"])
                    if not live[ph][s]:
                        return
                    if it["kind"] == 0:
                        for c in children[s]:
                            if owner[c] in (rank, -1):      # children owned by other ranks arrive through the all-reduce
                                assert done_sn[c] or fin[c] == n_items[c] > 0, ("rhs read before child finished", ph, s, c)
                        flag[(s, int(it["blk"]))] = True
                    fin[s] += 1
                    if fpar[s] >= 0:
                        assert live[ph][fpar[s]], "notification across phases"
                        cnt[fpar[s]] += 1
                assert _replay(fit, b, e, G, ready, done), f"forward dead-lock, rank {rank} phase {ph}"
                for x in range(b, e):
                    s = fit[x]["sn"]
                    if live[ph][s] and fin[s] == n_items[s]:
                        done_sn[s] = True
        assert np.all(done_sn[(owner == rank) | (owner == -1)])
        # ---- backward: phase 1 (top) then phase 0 (own subtrees) ---------------------------------------------------------
        bdone = np.zeros(ns, int)
        bdone_sn = np.zeros(ns, bool)
        below_done = np.zeros(ns, int)
        bflag = {}
        for ph in (1, 0):
            _, _, bwait = k.phase_deps(ph)
            for kind, b, e, lvl in ops["bwd_ops"]:
                if kind == 3:
                    continue
                if kind == 0:
                    for s in small_list[b:e]:
                        if live[ph][s]:
                            assert par[s] < 0 or bdone_sn[par[s]]
                            bdone_sn[s] = True
                    continue
                if kind == 2:
                    for s in np.unique(bp["bwd"]["sn"][b:e]):
                        if live[ph][s]:
                            assert par[s] < 0 or bdone_sn[par[s]]
                            bdone_sn[s] = True
                    continue

                def ready(x):
                    it = bit[x]; s = int(it["sn"])
                    if not live[ph][s]:
                        return True
                    w = bwait[s]
                    if w >= 0 and bdone[w] < ops["bwd_nitems"][w]:
                        return False
                    if it["kind"] == 2:
                        return True
                    if below_done[s] < ops["bwd_nbelow"][s]:
                        return False
                    return all(bflag.get((s, j), False) for j in range(int(it["blk"]) + 1, ncb(s)))

                def done(x):
                    it = bit[x]; s = int(it["sn"])
                    if not live[ph][s]:
                        return
                    p = par[s]
                    assert p < 0 or bdone_sn[p] or (ops["bwd_nitems"][p] > 0 and bdone[p] == ops["bwd_nitems"][p]), ("ancestor read early", ph, s, p)
                    if it["kind"] == 2:
                        below_done[s] += 1
                        return
                    bflag[(s, int(it["blk"]))] = True
                    bdone[s] += 1
                assert _replay(bit, b, e, G, ready, done), f"backward dead-lock, rank {rank} phase {ph}"
                for x in range(b, e):
                    s = bit[x]["sn"]
                    if live[ph][s] and bdone[s] == ops["bwd_nitems"][s]:
                        bdone_sn[s] = True
        assert np.all(bdone_sn[(owner == rank) | (owner == -1)])
