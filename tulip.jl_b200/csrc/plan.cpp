// Static plan + assembly maps (host).  See plan.hpp.
#include "plan.hpp"

#include <algorithm>
#include <stdexcept>

namespace tlp {

namespace {
inline int32_t sn_ncol(const Symbolic& S, int32_t s) { return S.sn_first[s + 1] - S.sn_first[s]; }
inline int32_t sn_nrow(const Symbolic& S, int32_t s) { return (int32_t)(S.sn_rowptr[s + 1] - S.sn_rowptr[s]); }
inline const int32_t* rows_of(const Symbolic& S, int32_t s) { return S.sn_rows.data() + S.sn_rowptr[s]; }
}  // namespace

int64_t lx_position(const Symbolic& S, int32_t gi, int32_t gk) {
    const int32_t t = S.col2sn[gk];
    const int32_t f = S.sn_first[t], l = S.sn_first[t + 1];
    const int64_t nrow = sn_nrow(S, t);
    int64_t pos;
    if (gi < l) pos = gi - f;
    else {
        const int32_t* b = S.sn_rows.data() + S.sn_rowptr[t] + (l - f);
        const int32_t* e = S.sn_rows.data() + S.sn_rowptr[t + 1];
        const int32_t* it = std::lower_bound(b, e, gi);
        if (it == e || *it != gi) throw std::runtime_error("lx_position: entry outside the factor pattern");
        pos = (l - f) + (it - b);
    }
    return S.sn_xptr[t] + (int64_t)(gk - f) * nrow + pos;
}

// tiles of the update  C[rows >= k, cols kbeg..kend) -= L[rows, piece cols] S L[cols, piece cols]'
static void emit_update_tiles(std::vector<UpdTask>& out, int32_t piece, int32_t kbeg, int32_t kend, int32_t nrow,
                              int32_t tgt, int32_t TILE = tlp::TILE) {
    for (int32_t k0 = kbeg; k0 < kend; k0 += TILE) {
        const int32_t nk = std::min(TILE, kend - k0);
        for (int32_t i0 = k0; i0 < nrow; i0 += TILE) {
            UpdTask t;
            t.piece = piece;
            t.i0 = i0;
            t.ni = std::min(TILE, nrow - i0);
            t.k0 = k0;
            t.nk = nk;
            t.tgt = tgt;
            t.diag = (i0 == k0) ? 1 : 0;
            t.pad = 0;
            out.push_back(t);
        }
    }
}

void build_plan(const Symbolic& S, const PlanOptions& opt, Plan& P) {
    P.opt = opt;
    const int32_t ns = S.nsuper;
    P.pieces.clear();
    P.sn_small.assign(ns, 0);
    P.sn_level.assign(ns, 0);
    P.sn_dblk.assign(ns, -1);
    P.sn_big.assign(ns, 0);
    P.sn_split.assign(ns, 0);
    P.bwd_below.clear();
    P.big_pack.clear();
    P.fwd_big.clear();
    P.bwd_big.clear();
    P.n_ftiles = P.n_btiles = 0;
    P.xq_slots = 0;
    P.dblk_sn.clear();
    P.dblk_idx.clear();
    P.seg_ptr.assign(ns + 1, 0);
    P.seg_k0.clear();
    P.seg_tgt.clear();
    P.max_small_elems = 0;
    P.max_small_nrow = 0;
    P.sn_oz.assign(ns, -1);
    P.oz_views.clear();
    P.oz_rb_off.clear();
    P.oz_slots = 0;
    P.oz_rows = 0;
    P.oz_tasks.clear();
    P.oz_slices.clear();
    P.flops_oz = 0.0;

    // target segments
    for (int32_t s = 0; s < ns; ++s) {
        const int32_t nc = sn_ncol(S, s), nr = sn_nrow(S, s);
        const int32_t* rows = S.sn_rows.data() + S.sn_rowptr[s];
        int32_t cur = -1;
        for (int32_t q = nc; q < nr; ++q) {
            const int32_t t = S.col2sn[rows[q]];
            if (t != cur) { P.seg_k0.push_back(q); P.seg_tgt.push_back(t); cur = t; }
        }
        P.seg_ptr[s + 1] = (int64_t)P.seg_k0.size();
    }

    // items and levels
    std::vector<int32_t> childmax(ns, -1), base_level(ns, 0);
    int32_t maxlevel = -1;
    for (int32_t s = 0; s < ns; ++s) {
        const int32_t nc = sn_ncol(S, s), nr = sn_nrow(S, s);
        const int32_t base = childmax[s] + 1;
        base_level[s] = base;
        const bool small = ((int64_t)nc * nr <= opt.small_elems) && nc <= opt.small_ncol && nr <= opt.small_nrow;
        if (small) {
            P.sn_small[s] = 1;
            P.sn_level[s] = base;
            P.max_small_elems = std::max(P.max_small_elems, nc * nr);
            P.max_small_nrow = std::max(P.max_small_nrow, nr);
        } else {
            const int32_t np = (nc + PIECE - 1) / PIECE;
            if (nc >= opt.big_ncol) P.sn_big[s] = 1;
            if (opt.oz_ncol > 0 && nc >= opt.oz_ncol && np >= 4) {
                bool allpos = true;
                for (int32_t j = S.sn_first[s]; j < S.sn_first[s + 1] && allpos; ++j) allpos = S.sign[j] > 0;
                for (int32_t q = nc; q < nr && allpos; ++q) allpos = S.sign[rows_of(S, s)[q]] > 0;
                if (allpos) {
                    OzViewPlan v;
                    v.sn = s;
                    v.nrb = (nr + 127) / 128;
                    v.ncb = np;
                    v.off0 = (int64_t)P.oz_rb_off.size();
                    v.row0 = P.oz_rows;
                    v.base_level = base;
                    v.pad = 0;
                    P.oz_rows += nr;
                    for (int32_t rb = 0; rb < v.nrb; ++rb) {
                        P.oz_rb_off.push_back(P.oz_slots);
                        P.oz_slots += 4 * (int64_t)std::min(rb, np);      // row block rb holds the planes of pieces 0 .. min(rb, ncb) - 1
                    }
                    P.oz_rb_off.push_back(P.oz_slots);
                    P.sn_oz[s] = (int32_t)P.oz_views.size();
                    P.oz_views.push_back(v);
                }
            }
            P.sn_dblk[s] = (int32_t)P.dblk_sn.size();
            for (int32_t k = 0; k < np; ++k) {
                Piece pc;
                pc.sn = s;
                pc.c0 = S.sn_first[s] + k * PIECE;
                pc.c1 = std::min(S.sn_first[s + 1], pc.c0 + PIECE);
                pc.level = base + k;
                P.pieces.push_back(pc);
                P.dblk_sn.push_back(s);
                P.dblk_idx.push_back(k);
            }
            P.sn_level[s] = base + np - 1;
        }
        maxlevel = std::max(maxlevel, P.sn_level[s]);
        const int32_t par = S.sn_parent[s];
        if (par >= 0) childmax[par] = std::max(childmax[par], P.sn_level[s]);
    }
    P.ndblk = (int32_t)P.dblk_sn.size();
    const int32_t nlev = maxlevel + 1;
    P.levels.assign(nlev, LevelPlan());

    // level of the item (small supernode / column piece) every column belongs to
    std::vector<int32_t> col_level(S.N, 0);
    for (int32_t s = 0; s < ns; ++s)
        for (int32_t j = S.sn_first[s]; j < S.sn_first[s + 1]; ++j)
            col_level[j] = P.sn_small[s] ? P.sn_level[s] : base_level[s] + (j - S.sn_first[s]) / PIECE;

    std::vector<std::vector<int32_t>> lsmall(nlev), lpiece(nlev), llarge(nlev), lbig(nlev);
    for (int32_t s = 0; s < ns; ++s) {
        if (P.sn_small[s]) lsmall[P.sn_level[s]].push_back(s);
        else if (P.sn_big[s]) lbig[base_level[s]].push_back(s);
        else llarge[base_level[s]].push_back(s);
    }
    for (int32_t p = 0; p < (int32_t)P.pieces.size(); ++p) lpiece[P.pieces[p].level].push_back(p);

    P.small_list.clear();
    P.level_pieces.clear();
    P.upd.clear();
    P.upd128.clear();
    P.panel.clear();
    P.fwd_items.clear();
    P.bwd_items.clear();
    P.flops_update = P.flops_panel = 0.0;
    for (int32_t L = 0; L < nlev; ++L) {
        LevelPlan& lp = P.levels[L];
        lp.small_begin = (int32_t)P.small_list.size();
        P.small_list.insert(P.small_list.end(), lsmall[L].begin(), lsmall[L].end());
        lp.small_end = (int32_t)P.small_list.size();
        lp.piece_begin = (int32_t)P.level_pieces.size();
        P.level_pieces.insert(P.level_pieces.end(), lpiece[L].begin(), lpiece[L].end());
        lp.piece_end = (int32_t)P.level_pieces.size();

        // trsm row tiles below each piece's diagonal block
        lp.panel_begin = (int32_t)P.panel.size();
        for (int32_t p : lpiece[L]) {
            const Piece& pc = P.pieces[p];
            const int32_t f = S.sn_first[pc.sn];
            const int32_t nrow = sn_nrow(S, pc.sn);
            const int32_t w = pc.c1 - pc.c0;
            for (int32_t r0 = pc.c1 - f; r0 < nrow; r0 += PIECE) {
                PanelTask pt;
                pt.piece = p;
                pt.r0 = r0;
                pt.nr = std::min(PIECE, nrow - r0);
                pt.pad = 0;
                P.panel.push_back(pt);
            }
            P.flops_panel += (double)w * w * w / 3.0 + (double)(nrow - (pc.c1 - f)) * w * w;
        }
        lp.panel_end = (int32_t)P.panel.size();

        // updates from each piece: rest of the own supernode, then ancestors
        // Column blocks of <= 128 target columns (never straddling two targets).  A block whose first
        // column belongs to an item of the next level is "urgent" (it gates that level's factorisation):
        // 64x64 tiles on the main stream.  Everything else is "lazy": 128x128 tiles for the persistent
        // side-stream kernel that runs underneath the critical chain.
        lp.ext_begin = (int32_t)P.upd.size();
        lp.lazy_begin = (int32_t)P.upd128.size();
        for (int32_t p : lpiece[L]) {
            const Piece& pc = P.pieces[p];
            const int32_t s = pc.sn;
            const int32_t f = S.sn_first[s], l = S.sn_first[s + 1];
            const int32_t nrow = sn_nrow(S, s);
            const int32_t* rows = S.sn_rows.data() + S.sn_rowptr[s];
            const size_t before = P.upd.size(), before128 = P.upd128.size();
            const int32_t ozv = P.sn_oz[s];
            auto emit_range = [&](int32_t kb, int32_t ke, int32_t tgt) {
                for (int32_t k0 = kb; k0 < ke; k0 += TILE128) {
                    const int32_t k1 = std::min(ke, k0 + TILE128);
                    if (col_level[rows[k0]] <= L + 1) emit_update_tiles(P.upd, p, k0, k1, nrow, tgt, TILE);
                    else if (tgt == s && ozv >= 0 && col_level[rows[k0]] >= L + 3) continue;   // tcgen05 path (below)
                    else emit_update_tiles(P.upd128, p, k0, k1, nrow, tgt, TILE);
                }
            };
            if (pc.c1 < l) emit_range(pc.c1 - f, l - f, s);
            for (int64_t g = P.seg_ptr[s]; g < P.seg_ptr[s + 1]; ++g) {
                const int32_t kb = P.seg_k0[g];
                const int32_t ke = (g + 1 < P.seg_ptr[s + 1]) ? P.seg_k0[g + 1] : nrow;
                emit_range(kb, ke, P.seg_tgt[g]);
            }
            const double w = pc.c1 - pc.c0;
            auto tile_flops = [&](const UpdTask& t) {
                double ent = (double)t.ni * t.nk;
                if (t.diag) ent -= (double)t.nk * (t.nk - 1) / 2.0;
                return 2.0 * ent * w;
            };
            for (size_t x = before; x < P.upd.size(); ++x) P.flops_update += tile_flops(P.upd[x]);
            for (size_t x = before128; x < P.upd128.size(); ++x) P.flops_update += tile_flops(P.upd128[x]);
        }
        lp.ext_end = (int32_t)P.upd.size();
        lp.urgent_end = lp.ext_end;
        lp.lazy_end = (int32_t)P.upd128.size();

        // tcgen05 tasks: piece j of an oz supernode completes the K range [0, 128 (j+1)) of column block c = j + 3
        lp.oz_begin = (int32_t)P.oz_tasks.size();
        lp.ozs_begin = (int32_t)P.oz_slices.size();
        {
            int64_t n128 = 0;     // 128x128 tasks this level's launch would have
            const int32_t kmax0 = std::max(1, std::min(opt.oz_ksplit, 4096) / 32);
            for (int32_t p : lpiece[L]) {
                const Piece& pc = P.pieces[p];
                const int32_t vi = P.sn_oz[pc.sn];
                if (vi < 0) continue;
                const OzViewPlan& v = P.oz_views[vi];
                const int32_t j = (pc.c0 - S.sn_first[pc.sn]) / PIECE;
                if (j + 3 >= v.ncb) continue;
                n128 += (int64_t)(v.nrb - (j + 3)) * ((4 * (j + 1) + kmax0 - 1) / kmax0);
            }
            lp.oz_tile = opt.oz_tile_n == 128 ? 128 : (opt.oz_tile_n == 64 ? 64 : (n128 >= 264 ? 128 : 64));
        }
        for (int32_t p : lpiece[L]) {
            const Piece& pc = P.pieces[p];
            const int32_t s = pc.sn;
            const int32_t vi = P.sn_oz[s];
            if (vi < 0) continue;
            const OzViewPlan& v = P.oz_views[vi];
            const int32_t f = S.sn_first[s], nc = sn_ncol(S, s), nrow = sn_nrow(S, s);
            const int32_t j = (pc.c0 - f) / PIECE;
            const int32_t c = j + 3;
            if (c >= v.ncb) continue;
            OzSlice sl;
            sl.view = vi; sl.piece = p; sl.j = j; sl.pad = 0;
            P.oz_slices.push_back(sl);
            const int32_t nk32 = 4 * (j + 1);
            const int32_t ncolc = std::min(128, nc - c * 128);
            const int32_t nhalf = (lp.oz_tile == 128) ? 1 : (ncolc > 64 ? 2 : 1);
            const int32_t tn = lp.oz_tile;
            const int32_t nrbt = v.nrb - c;
            const int32_t kmax = std::max(1, std::min(opt.oz_ksplit, 4096) / 32);
            int32_t nsplit = (nk32 + kmax - 1) / kmax;
            const int32_t want = (132 + nrbt * nhalf - 1) / (nrbt * nhalf);      // enough tasks for one wave ...
            nsplit = std::max(nsplit, std::min(want, std::max(1, nk32 / 16)));  // ... but at least 512 columns each
            const int32_t per = (nk32 + nsplit - 1) / nsplit;
            for (int32_t k0 = 0; k0 < nk32; k0 += per)
                for (int32_t a = c; a < v.nrb; ++a)
                    for (int32_t h = 0; h < nhalf; ++h) {
                        OzTask t;
                        t.view = vi; t.rbA = a; t.rbB = c; t.half = h; t.k0 = k0; t.k1 = std::min(nk32, k0 + per);
                        t.pad[0] = t.pad[1] = 0;
                        P.oz_tasks.push_back(t);
                        const int32_t ni = std::min(128, nrow - a * 128), nk = std::min(tn, ncolc - h * 64);
                        double ent = (double)ni * nk;
                        if (a == c) {
                            ent = 0.0;
                            for (int32_t jj = h * 64; jj < h * 64 + nk; ++jj) ent += std::max(0, ni - jj);
                        }
                        P.flops_oz += 2.0 * ent * 32.0 * (t.k1 - t.k0);
                    }
        }
        lp.oz_end = (int32_t)P.oz_tasks.size();
        lp.ozs_end = (int32_t)P.oz_slices.size();
        lp.ext_atomic = 1;

        // Critical subset: the update tiles that land in a diagonal block of a next-level piece or anywhere in a
        // next-level small supernode are all that the first kernels of level L+1 (one-CTA supernodes, diagonal
        // blocks) wait for; the row tiles of the trsm they read are the critical part of the trsm.  Both lists
        // are reordered critical-first so that the chain can run them ahead of the rest (solver.cu).
        {
            std::vector<char> ucrit(lp.ext_end - lp.ext_begin, 0);
            std::vector<std::vector<char>> rowmark(lpiece[L].size());
            auto piece_slot = [&](int32_t p) { return (size_t)(std::lower_bound(lpiece[L].begin(), lpiece[L].end(), p) - lpiece[L].begin()); };
            for (int32_t x = lp.ext_begin; x < lp.ext_end; ++x) {
                const UpdTask& T = P.upd[x];
                const int32_t s = P.pieces[T.piece].sn;
                const int32_t* rows = S.sn_rows.data() + S.sn_rowptr[s];
                bool cr = false;
                int32_t last_pf = -1;                 // piece already tested (and missed) by an earlier column of the tile
                for (int32_t k = T.k0; k < T.k0 + T.nk && !cr; ++k) {
                    const int32_t gk = rows[k];
                    if (col_level[gk] != L + 1) continue;
                    const int32_t ts = S.col2sn[gk];
                    if (P.sn_small[ts]) { cr = true; break; }
                    const int32_t pf = S.sn_first[ts] + ((gk - S.sn_first[ts]) / PIECE) * PIECE;
                    if (pf == last_pf) continue;
                    last_pf = pf;
                    const int32_t pl = std::min(pf + PIECE, S.sn_first[ts + 1]);
                    // the row list is ascending: does the tile's row range meet [pf, pl)?
                    const int32_t* rb = rows + T.i0;
                    const int32_t* re = rb + T.ni;
                    const int32_t* it = std::lower_bound(rb, re, pf);
                    if (it != re && *it < pl) cr = true;
                }
                if (!cr) continue;
                ucrit[x - lp.ext_begin] = 1;
                std::vector<char>& mk = rowmark[piece_slot(T.piece)];
                if (mk.empty()) mk.assign(sn_nrow(S, s), 0);
                for (int32_t i = T.i0; i < T.i0 + T.ni; ++i) mk[i] = 1;
                for (int32_t k = T.k0; k < T.k0 + T.nk; ++k) mk[k] = 1;
            }
            std::vector<UpdTask> a, b;
            for (int32_t x = lp.ext_begin; x < lp.ext_end; ++x) (ucrit[x - lp.ext_begin] ? a : b).push_back(P.upd[x]);
            std::copy(a.begin(), a.end(), P.upd.begin() + lp.ext_begin);
            std::copy(b.begin(), b.end(), P.upd.begin() + lp.ext_begin + (int32_t)a.size());
            lp.ext_crit_end = lp.ext_begin + (int32_t)a.size();
            std::vector<PanelTask> pa, pb;
            for (int32_t x = lp.panel_begin; x < lp.panel_end; ++x) {
                const PanelTask& pt = P.panel[x];
                const std::vector<char>& mk = rowmark[piece_slot(pt.piece)];
                bool cr = false;
                if (!mk.empty())
                    for (int32_t r = pt.r0; r < pt.r0 + pt.nr; ++r)
                        if (mk[r]) { cr = true; break; }
                (cr ? pa : pb).push_back(pt);
            }
            std::copy(pa.begin(), pa.end(), P.panel.begin() + lp.panel_begin);
            std::copy(pb.begin(), pb.end(), P.panel.begin() + lp.panel_begin + (int32_t)pa.size());
            lp.panel_crit_end = lp.panel_begin + (int32_t)pa.size();
        }

        // dense block-solve items of the supernodes that *start* at this level, in wavefront order
        // (block index major) so that every dependency of an item precedes it in the list
        lp.fwd_begin = (int32_t)P.fwd_items.size();
        lp.bwd_begin = (int32_t)P.bwd_items.size();
        {
            int32_t maxblk = 0;
            for (int32_t s : llarge[L]) {
                const int32_t nc = sn_ncol(S, s), nr = sn_nrow(S, s);
                maxblk = std::max(maxblk, (nc + SBLK - 1) / SBLK + (nr - nc + SBLK - 1) / SBLK);
            }
            for (int32_t b = 0; b < maxblk; ++b)
                for (int32_t s : llarge[L]) {
                    const int32_t nc = sn_ncol(S, s), nr = sn_nrow(S, s);
                    const int32_t ncb = (nc + SBLK - 1) / SBLK, nbb = (nr - nc + SBLK - 1) / SBLK;
                    if (b >= ncb + nbb) continue;
                    SolveItem it;
                    it.sn = s;
                    it.pad = 0;
                    if (b < ncb) { it.kind = 0; it.blk = b; it.r0 = b * SBLK; it.nr = std::min(SBLK, nc - it.r0); }
                    else { it.kind = 1; it.blk = b - ncb; it.r0 = nc + (b - ncb) * SBLK; it.nr = std::min(SBLK, nr - it.r0); }
                    P.fwd_items.push_back(it);
                }
            int32_t maxcb = 0;
            for (int32_t s : llarge[L]) maxcb = std::max(maxcb, (sn_ncol(S, s) + SBLK - 1) / SBLK);
            for (int32_t d = 0; d < maxcb; ++d)      // d-th block from the end
                for (int32_t s : llarge[L]) {
                    const int32_t nc = sn_ncol(S, s);
                    const int32_t ncb = (nc + SBLK - 1) / SBLK;
                    if (d >= ncb) continue;
                    SolveItem it;
                    it.sn = s;
                    it.pad = 0;
                    it.kind = 0;
                    it.blk = ncb - 1 - d;
                    it.r0 = it.blk * SBLK;
                    it.nr = std::min(SBLK, nc - it.r0);
                    P.bwd_items.push_back(it);
                }
        }
        lp.fwd_end = (int32_t)P.fwd_items.size();
        lp.bwd_end = (int32_t)P.bwd_items.size();

        // dense-solve tasks of the big supernodes that start at this level: same wavefront order; the tiles
        // of a task are laid out contiguously in the order the task streams them
        lp.fbig_begin = (int32_t)P.fwd_big.size();
        lp.bbig_begin = (int32_t)P.bwd_big.size();
        if (!lbig[L].empty()) {
            const size_t nb_ = lbig[L].size();
            std::vector<int32_t> xq0(nb_);
            std::vector<std::vector<int64_t>> ftile0(nb_), btile0(nb_);   // per supernode: first tile of every fwd / bwd task
            int32_t maxblk = 0, maxcb = 0;
            for (size_t x = 0; x < nb_; ++x) {
                const int32_t s = lbig[L][x];
                const int32_t nc = sn_ncol(S, s), nr = sn_nrow(S, s);
                const int32_t ncb = (nc + SBLK - 1) / SBLK, nbb = (nr - nc + SBLK - 1) / SBLK;
                maxblk = std::max(maxblk, ncb + nbb);
                maxcb = std::max(maxcb, ncb);
                xq0[x] = P.xq_slots;
                P.xq_slots += ncb * SBLK;
                ftile0[x].resize(ncb + nbb);
                btile0[x].resize(ncb);
                for (int32_t b = 0; b < ncb + nbb; ++b) { ftile0[x][b] = P.n_ftiles; P.n_ftiles += (b < ncb) ? b : ncb; }
                for (int32_t k = 0; k < ncb; ++k) { btile0[x][k] = P.n_btiles; P.n_btiles += ncb - 1 - k; }
                // one pack task per tile: block row b (column block or below block) x column block j
                for (int32_t b = 0; b < ncb + nbb; ++b) {
                    const int32_t rr0 = (b < ncb) ? b * SBLK : nc + (b - ncb) * SBLK;   // below blocks start at row nc
                    const int32_t rnr = (b < ncb) ? std::min(SBLK, nc - rr0) : std::min(SBLK, nr - rr0);
                    const int32_t nj = (b < ncb) ? b : ncb;
                    for (int32_t j = 0; j < nj; ++j) {
                        BigPack pk;
                        pk.sn = s;
                        pk.r0 = rr0;
                        pk.nr = rnr;
                        pk.j = j;
                        pk.fdst = ftile0[x][b] + j;
                        pk.bdst = (b < ncb) ? btile0[x][j] + (ncb - 1 - b) : -1;
                        P.big_pack.push_back(pk);
                    }
                }
            }
            for (int32_t b = 0; b < maxblk; ++b)
                for (size_t x = 0; x < nb_; ++x) {
                    const int32_t s = lbig[L][x];
                    const int32_t nc = sn_ncol(S, s), nr = sn_nrow(S, s);
                    const int32_t ncb = (nc + SBLK - 1) / SBLK, nbb = (nr - nc + SBLK - 1) / SBLK;
                    if (b >= ncb + nbb) continue;
                    BigTask t;
                    t.sn = s;
                    t.nbelow = 0;
                    t.xq0 = xq0[x];
                    t.tile0 = ftile0[x][b];
                    if (b < ncb) { t.kind = 0; t.blk = b; t.r0 = b * SBLK; t.nr = std::min(SBLK, nc - t.r0); t.ntile = b; }
                    else { t.kind = 1; t.blk = b - ncb; t.r0 = nc + (b - ncb) * SBLK; t.nr = std::min(SBLK, nr - t.r0); t.ntile = ncb; }
                    P.fwd_big.push_back(t);
                }
            for (int32_t d = 0; d < maxcb; ++d)
                for (size_t x = 0; x < nb_; ++x) {
                    const int32_t s = lbig[L][x];
                    const int32_t nc = sn_ncol(S, s);
                    const int32_t ncb = (nc + SBLK - 1) / SBLK;
                    if (d >= ncb) continue;
                    BigTask t;
                    t.sn = s;
                    t.kind = 0;
                    t.blk = ncb - 1 - d;
                    t.r0 = t.blk * SBLK;
                    t.nr = std::min(SBLK, nc - t.r0);
                    t.nbelow = 0;
                    t.ntile = d;
                    t.xq0 = xq0[x];
                    t.tile0 = btile0[x][t.blk];
                    P.bwd_big.push_back(t);
                }
        }
        lp.fbig_end = (int32_t)P.fwd_big.size();
        lp.bbig_end = (int32_t)P.bwd_big.size();

        // backward "below" items of the tall supernodes that start at this level
        lp.below_begin = (int32_t)P.bwd_below.size();
        for (int pass = 0; pass < 2; ++pass)
            for (int32_t s : (pass == 0 ? lbig[L] : llarge[L])) {
                const int32_t nc = sn_ncol(S, s), nr = sn_nrow(S, s);
                if (nr == nc || (pass == 1 && nr - nc < SPLIT_ROWS)) continue;
                P.sn_split[s] = 1;
                const int32_t ncb = (nc + SBLK - 1) / SBLK;
                for (int32_t k = 0; k < ncb; ++k)
                    for (int32_t r0 = nc; r0 < nr; r0 += BELOW_ROWS) {
                        BelowItem bi;
                        bi.sn = s;
                        bi.blk = k;
                        bi.r0 = r0;
                        bi.nr = std::min(BELOW_ROWS, nr - r0);
                        P.bwd_below.push_back(bi);
                    }
            }
        lp.below_end = (int32_t)P.bwd_below.size();
    }

    // ---- launch sequences of the triangular sweeps with the block-solve items of consecutive levels merged (SolveOp) ----
    {
        const int32_t nlev = (int32_t)P.levels.size();
        P.fwd_ops.clear(); P.bwd_ops.clear(); P.bwd_seq.clear();
        P.fwd_need.assign(S.nsuper, 0); P.fwd_parent.assign(S.nsuper, -1);
        P.bwd_wait.assign(S.nsuper, -1); P.bwd_nitems.assign(S.nsuper, 0);
        std::vector<int32_t> frun(S.nsuper, -1), brun(S.nsuper, -1), fitems(S.nsuper, 0);
        // forward: levels ascending; per level the launch order is small, large, big
        int32_t pend_b = -1, pend_e = -1, pend_l = 0;
        auto flush_f = [&]() {
            if (pend_b >= 0 && pend_e > pend_b) P.fwd_ops.push_back(SolveOp{1, pend_b, pend_e, pend_l});
            pend_b = pend_e = -1;
        };
        for (int32_t L = 0; L < nlev; ++L) {
            const LevelPlan& lp = P.levels[L];
            if (lp.small_end > lp.small_begin) { flush_f(); P.fwd_ops.push_back(SolveOp{0, lp.small_begin, lp.small_end, L}); }
            if (lp.fwd_end > lp.fwd_begin) {
                if (pend_b >= 0 && pend_e != lp.fwd_begin) flush_f();
                if (pend_b < 0) { pend_b = lp.fwd_begin; pend_l = L; }
                pend_e = lp.fwd_end;
                for (int32_t x = lp.fwd_begin; x < lp.fwd_end; ++x) {
                    frun[P.fwd_items[x].sn] = (int32_t)P.fwd_ops.size();     // id of the op this run will become
                    fitems[P.fwd_items[x].sn]++;
                }
            }
            if (lp.fbig_end > lp.fbig_begin) { flush_f(); P.fwd_ops.push_back(SolveOp{2, lp.fbig_begin, lp.fbig_end, L}); }
        }
        flush_f();
        for (int32_t s2 = 0; s2 < S.nsuper; ++s2) {
            const int32_t par = S.sn_parent[s2];
            if (frun[s2] >= 0 && par >= 0 && frun[par] == frun[s2]) {
                P.fwd_parent[s2] = par;
                P.fwd_need[par] += fitems[s2];
            }
        }
        // backward: levels descending; per level the launch order is below, big, large, small.  The below items of the
        // dense-solve supernodes stay a launch of their own (the dense-solve launch that follows breaks the run anyway);
        // those of the block-solve supernodes join the merged sequence as items of kind 2, ahead of their level's blocks.
        P.bwd_nbelow.assign(S.nsuper, 0);
        std::vector<int32_t> seq_b(nlev, 0), seq_e(nlev, 0), below_big_end(nlev, 0);
        for (int32_t L = nlev - 1; L >= 0; --L) {
            const LevelPlan& lp = P.levels[L];
            seq_b[L] = (int32_t)P.bwd_seq.size();
            below_big_end[L] = lp.below_begin;
            for (int32_t x = lp.below_begin; x < lp.below_end; ++x) {
                const BelowItem& bi = P.bwd_below[x];
                if (P.sn_big[bi.sn]) {
                    if (x != below_big_end[L]) throw std::logic_error("below items: dense-solve supernodes are expected first");
                    below_big_end[L] = x + 1;
                    continue;
                }
                SolveItem it;
                it.sn = bi.sn; it.blk = bi.blk; it.kind = 2; it.r0 = bi.r0; it.nr = bi.nr; it.pad = 0;
                P.bwd_seq.push_back(it);
                P.bwd_nbelow[bi.sn]++;
            }
            for (int32_t x = lp.bwd_begin; x < lp.bwd_end; ++x) { P.bwd_seq.push_back(P.bwd_items[x]); P.bwd_nitems[P.bwd_items[x].sn]++; }
            seq_e[L] = (int32_t)P.bwd_seq.size();
        }
        pend_b = pend_e = -1;
        auto flush_b = [&]() {
            if (pend_b >= 0 && pend_e > pend_b) P.bwd_ops.push_back(SolveOp{1, pend_b, pend_e, pend_l});
            pend_b = pend_e = -1;
        };
        for (int32_t L = nlev - 1; L >= 0; --L) {
            const LevelPlan& lp = P.levels[L];
            if (below_big_end[L] > lp.below_begin) { flush_b(); P.bwd_ops.push_back(SolveOp{3, lp.below_begin, below_big_end[L], L}); }
            if (lp.bbig_end > lp.bbig_begin) { flush_b(); P.bwd_ops.push_back(SolveOp{2, lp.bbig_begin, lp.bbig_end, L}); }
            if (seq_e[L] > seq_b[L]) {
                if (pend_b >= 0 && pend_e != seq_b[L]) flush_b();
                if (pend_b < 0) { pend_b = seq_b[L]; pend_l = L; }
                pend_e = seq_e[L];
                for (int32_t x = seq_b[L]; x < seq_e[L]; ++x) brun[P.bwd_seq[x].sn] = (int32_t)P.bwd_ops.size();
            }
            if (lp.small_end > lp.small_begin) { flush_b(); P.bwd_ops.push_back(SolveOp{0, lp.small_begin, lp.small_end, L}); }
        }
        flush_b();
        for (int32_t s2 = 0; s2 < S.nsuper; ++s2) {
            const int32_t par = S.sn_parent[s2];
            if (brun[s2] >= 0 && par >= 0 && brun[par] == brun[s2]) P.bwd_wait[s2] = par;
        }
    }

    // diagonal blocks and pack tiles in level order: solver.cu starts inverting / repacking the columns that are
    // final while the tail of the factorisation (which leaves most SMs idle) is still running
    P.inv_order.resize(P.pieces.size());
    for (int32_t p = 0; p < (int32_t)P.pieces.size(); ++p) P.inv_order[p] = p;
    std::stable_sort(P.inv_order.begin(), P.inv_order.end(),
                     [&](int32_t a, int32_t b) { return P.pieces[a].level < P.pieces[b].level; });
    auto pack_level = [&](const BigPack& k) { return P.pieces[P.sn_dblk[k.sn] + k.j].level; };
    std::stable_sort(P.big_pack.begin(), P.big_pack.end(),
                     [&](const BigPack& a, const BigPack& b) { return pack_level(a) < pack_level(b); });
    {
        size_t ii = 0, pp = 0;
        for (int32_t L = 0; L < nlev; ++L) {
            while (ii < P.inv_order.size() && P.pieces[P.inv_order[ii]].level <= L) ++ii;
            while (pp < P.big_pack.size() && pack_level(P.big_pack[pp]) <= L) ++pp;
            P.levels[L].inv_end = (int32_t)ii;
            P.levels[L].pack_end = (int32_t)pp;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Assembly maps
// ------------------------------------------------------------------------------------------
void build_assembly_k1(const Symbolic& S, int64_t m, int64_t n, const int64_t* colptr, const int32_t* rowidx,
                       const double* val, AssemblyMaps& M) {
    const int64_t nnz = colptr[n];
    // CSR of A with permuted row numbering: row q (permuted) -> (column j, value)
    std::vector<int64_t> rp(m + 1, 0);
    for (int64_t p = 0; p < nnz; ++p) rp[S.iperm[rowidx[p]] + 1]++;
    for (int64_t q = 0; q < m; ++q) rp[q + 1] += rp[q];
    std::vector<int32_t> rc(nnz);
    std::vector<double> rv(nnz);
    {
        std::vector<int64_t> nxt(rp.begin(), rp.end() - 1);
        for (int64_t j = 0; j < n; ++j)
            for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                const int64_t q = S.iperm[rowidx[p]];
                rc[nxt[q]] = (int32_t)j;
                rv[nxt[q]] = val[p];
                nxt[q]++;
            }
    }
    struct Prod { int64_t dest; int32_t col; double w; };
    std::vector<Prod> buf;
    std::vector<int32_t> posmap(m, -1), postag(m, -1);
    M.w_ptr.clear();
    M.w_dest.clear();
    M.w_col.clear();
    M.w_val.clear();
    int32_t cur_t = -1;
    for (int32_t c = 0; c < (int32_t)m; ++c) {
        const int32_t t = S.col2sn[c];
        if (t != cur_t) {
            const int64_t b = S.sn_rowptr[t], e = S.sn_rowptr[t + 1];
            for (int64_t x = b; x < e; ++x) { posmap[S.sn_rows[x]] = (int32_t)(x - b); postag[S.sn_rows[x]] = t; }
            cur_t = t;
        }
        const int64_t nrow = S.sn_rowptr[t + 1] - S.sn_rowptr[t];
        const int64_t colbase = S.sn_xptr[t] + (int64_t)(c - S.sn_first[t]) * nrow;
        buf.clear();
        for (int64_t x = rp[c]; x < rp[c + 1]; ++x) {
            const int32_t j = rc[x];
            const double acj = rv[x];
            for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                const int32_t q = S.iperm[rowidx[p]];
                if (q < c) continue;
                if (postag[q] != t) throw std::runtime_error("build_assembly_k1: A*A' entry outside the factor pattern");
                buf.push_back(Prod{colbase + posmap[q], j, val[p] * acj});
            }
        }
        std::sort(buf.begin(), buf.end(), [](const Prod& a, const Prod& b) {
            return a.dest != b.dest ? a.dest < b.dest : a.col < b.col;
        });
        for (size_t x = 0; x < buf.size(); ++x) {
            if (x == 0 || buf[x].dest != buf[x - 1].dest) {
                M.w_ptr.push_back((int64_t)M.w_col.size());
                M.w_dest.push_back(buf[x].dest);
            }
            M.w_col.push_back(buf[x].col);
            M.w_val.push_back(buf[x].w);
        }
    }
    M.w_ptr.push_back((int64_t)M.w_col.size());
}

void build_assembly_k2(const Symbolic& S, int64_t m, int64_t n, const int64_t* colptr, const int32_t* rowidx,
                       AssemblyMaps& M) {
    (void)m;
    const int64_t nnz = colptr[n];
    M.a_dest.resize(nnz);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
            const int32_t a = S.iperm[j], b = S.iperm[n + rowidx[p]];
            M.a_dest[p] = lx_position(S, std::max(a, b), std::min(a, b));
        }
}

}  // namespace tlp

// ------------------------------------------------------------------------------------------
// Multi-GPU: shard independent elimination-tree subtrees across ranks.  Every rank runs
// this on the same symbolic analysis and gets the same answer.  owner[s] = rank that factors
// supernode s, or -1 for the replicated top part (the separator: ancestors of the cut).
// ------------------------------------------------------------------------------------------
namespace tlp {

void partition_subtrees(const Symbolic& S, int nranks, std::vector<int32_t>& owner, std::vector<double>* rank_work) {
    const int32_t ns = S.nsuper;
    owner.assign(ns, 0);
    if (rank_work) rank_work->assign(std::max(nranks, 1), 0.0);
    if (nranks <= 1 || ns == 0) return;
    // work of a supernode ~ sum of squared column counts of its columns; subtree sums bottom-up
    std::vector<double> work(ns, 0.0), sub(ns, 0.0);
    for (int32_t s = 0; s < ns; ++s) {
        for (int32_t j = S.sn_first[s]; j < S.sn_first[s + 1]; ++j) work[s] += (double)S.colcount[j] * S.colcount[j];
        sub[s] += work[s];
        if (S.sn_parent[s] >= 0) sub[S.sn_parent[s]] += sub[s];
    }
    std::vector<std::vector<int32_t>> kids(ns);
    std::vector<int32_t> cand;                  // current subtree roots
    for (int32_t s = 0; s < ns; ++s) {
        if (S.sn_parent[s] >= 0) kids[S.sn_parent[s]].push_back(s);
        else cand.push_back(s);
    }
    std::vector<char> top(ns, 0);
    for (;;) {
        // heaviest candidate (ties: lowest index)
        int32_t best = -1;
        double total = 0.0;
        for (int32_t r : cand) {
            total += sub[r];
            if (best < 0 || sub[r] > sub[best]) best = r;
        }
        if (best < 0) break;
        const bool enough = (int)cand.size() >= 2 * nranks && sub[best] <= 0.6 * total / nranks;
        if (enough || kids[best].empty()) break;
        top[best] = 1;
        cand.erase(std::find(cand.begin(), cand.end(), best));
        for (int32_t k : kids[best]) cand.push_back(k);
    }
    // longest-processing-time assignment of the subtrees
    std::vector<int32_t> order(cand);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return sub[a] != sub[b] ? sub[a] > sub[b] : a < b; });
    std::vector<double> load(nranks, 0.0);
    std::vector<int32_t> root_owner(ns, -1);
    for (int32_t r : order) {
        int g = 0;
        for (int q = 1; q < nranks; ++q) if (load[q] < load[g]) g = q;
        root_owner[r] = g;
        load[g] += sub[r];
    }
    // propagate down: supernodes are postordered, so parents come after children -> walk top-down
    for (int32_t s = ns - 1; s >= 0; --s) {
        if (top[s]) owner[s] = -1;
        else if (root_owner[s] >= 0) owner[s] = root_owner[s];
        else owner[s] = owner[S.sn_parent[s]];
    }
    if (rank_work) *rank_work = load;
}

// Panel storage order: supernodes of the subtrees first, the replicated top part last (so that the
// top panels form one contiguous range for the all-reduce).  Recomputes sn_xptr and diagpos.
int64_t relayout_panels(Symbolic& S, const std::vector<int32_t>& owner, int32_t nranks, std::vector<int64_t>* rank_begin) {
    const int32_t ns = S.nsuper;
    int64_t off = 0, top_begin = 0;
    if (rank_begin) rank_begin->assign((size_t)nranks + 1, 0);
    // panels grouped by owner: [rank 0 | rank 1 | ... | replicated top part], so that a rank clears / touches one contiguous range
    for (int32_t pass = 0; pass <= nranks; ++pass) {
        if (pass == nranks) top_begin = off;
        if (rank_begin && pass < nranks) (*rank_begin)[pass] = off;
        for (int32_t s = 0; s < ns; ++s) {
            if (pass < nranks ? (owner[s] != pass) : (owner[s] >= 0)) continue;
            const int64_t ncol = S.sn_first[s + 1] - S.sn_first[s];
            const int64_t nrow = S.sn_rowptr[s + 1] - S.sn_rowptr[s];
            S.sn_xptr[s] = off;
            off += nrow * ncol;
        }
    }
    if (rank_begin) (*rank_begin)[nranks] = top_begin;
    S.sn_xptr[ns] = off;     // == lx_size (the last entry is no longer "end of supernode ns-1")
    for (int32_t s = 0; s < ns; ++s) {
        const int64_t nrow = S.sn_rowptr[s + 1] - S.sn_rowptr[s];
        for (int32_t j = S.sn_first[s]; j < S.sn_first[s + 1]; ++j) {
            const int64_t lc = j - S.sn_first[s];
            S.diagpos[j] = S.sn_xptr[s] + lc * nrow + lc;
        }
    }
    return top_begin;
}

}  // namespace tlp
