// Static plan + assembly maps (host).  See plan.hpp.
#include "plan.hpp"

#include <algorithm>
#include <stdexcept>

namespace tlp {

namespace {
inline int32_t sn_ncol(const Symbolic& S, int32_t s) { return S.sn_first[s + 1] - S.sn_first[s]; }
inline int32_t sn_nrow(const Symbolic& S, int32_t s) { return (int32_t)(S.sn_rowptr[s + 1] - S.sn_rowptr[s]); }
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

static void emit_update_tiles(std::vector<UpdTask>& out, int32_t piece, int32_t kdim, int32_t kbeg, int32_t kend,
                              int32_t nrow, int32_t tgt, int32_t diagflag) {
    for (int32_t k0 = kbeg; k0 < kend; k0 += TILE) {
        const int32_t nk = std::min(TILE, kend - k0);
        for (int32_t i0 = k0; i0 < nrow; i0 += TILE) {
            if (kdim == 0 && i0 != k0) break;   // nothing to subtract: only the diagonal tile (potrf) is needed
            UpdTask t;
            t.piece = piece;
            t.kdim = kdim;
            t.i0 = i0;
            t.ni = std::min(TILE, nrow - i0);
            t.k0 = k0;
            t.nk = nk;
            t.tgt = tgt;
            t.diag = (i0 == k0) ? diagflag : 0;
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
    P.seg_ptr.assign(ns + 1, 0);
    P.seg_k0.clear();
    P.seg_tgt.clear();
    P.max_small_elems = 0;
    P.max_small_nrow = 0;

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
    std::vector<int32_t> childmax(ns, -1);
    std::vector<int32_t> first_piece(ns, -1), npieces(ns, 0);
    int32_t maxlevel = -1;
    for (int32_t s = 0; s < ns; ++s) {
        const int32_t nc = sn_ncol(S, s), nr = sn_nrow(S, s);
        const int32_t base = childmax[s] + 1;
        const bool small = ((int64_t)nc * nr <= opt.small_elems) && nc <= opt.small_ncol;
        if (small) {
            P.sn_small[s] = 1;
            P.sn_level[s] = base;
            P.max_small_elems = std::max(P.max_small_elems, nc * nr);
            P.max_small_nrow = std::max(P.max_small_nrow, nr);
        } else {
            const int32_t W = std::max(TILE, (opt.piece_width / TILE) * TILE);
            const int32_t np = (nc + W - 1) / W;
            first_piece[s] = (int32_t)P.pieces.size();
            npieces[s] = np;
            for (int32_t k = 0; k < np; ++k) {
                Piece pc;
                pc.sn = s;
                pc.c0 = S.sn_first[s] + k * W;
                pc.c1 = std::min(S.sn_first[s + 1], pc.c0 + W);
                pc.level = base + k;
                P.pieces.push_back(pc);
            }
            P.sn_level[s] = base + np - 1;
        }
        maxlevel = std::max(maxlevel, P.sn_level[s]);
        const int32_t par = S.sn_parent[s];
        if (par >= 0) childmax[par] = std::max(childmax[par], P.sn_level[s]);
    }
    const int32_t nlev = maxlevel + 1;
    P.levels.assign(nlev, LevelPlan());

    // group by level
    std::vector<std::vector<int32_t>> lsmall(nlev), lpiece(nlev);
    for (int32_t s = 0; s < ns; ++s)
        if (P.sn_small[s]) lsmall[P.sn_level[s]].push_back(s);
    for (int32_t p = 0; p < (int32_t)P.pieces.size(); ++p) lpiece[P.pieces[p].level].push_back(p);

    P.small_list.clear();
    P.level_pieces.clear();
    P.upd.clear();
    P.panel.clear();
    P.solve.clear();
    for (int32_t L = 0; L < nlev; ++L) {
        LevelPlan& lp = P.levels[L];
        lp.small_begin = (int32_t)P.small_list.size();
        P.small_list.insert(P.small_list.end(), lsmall[L].begin(), lsmall[L].end());
        lp.small_end = (int32_t)P.small_list.size();
        lp.piece_begin = (int32_t)P.level_pieces.size();
        P.level_pieces.insert(P.level_pieces.end(), lpiece[L].begin(), lpiece[L].end());
        lp.piece_end = (int32_t)P.level_pieces.size();

        int32_t nsteps = 0;
        for (int32_t p : lpiece[L]) nsteps = std::max(nsteps, (P.pieces[p].c1 - P.pieces[p].c0 + TILE - 1) / TILE);
        lp.nsteps = nsteps;
        lp.inner_begin.assign(nsteps, 0);
        lp.inner_end.assign(nsteps, 0);
        lp.panel_begin.assign(nsteps, 0);
        lp.panel_end.assign(nsteps, 0);
        for (int32_t t = 0; t < nsteps; ++t) {
            lp.inner_begin[t] = (int32_t)P.upd.size();
            for (int32_t p : lpiece[L]) {
                const Piece& pc = P.pieces[p];
                const int32_t w = pc.c1 - pc.c0;
                if (t * TILE >= w) continue;
                const int32_t f = S.sn_first[pc.sn];
                const int32_t nrow = sn_nrow(S, pc.sn);
                const int32_t kbeg = (pc.c0 - f) + t * TILE;
                const int32_t kend = (pc.c0 - f) + std::min(w, (t + 1) * TILE);
                // block column t of the piece: update with the piece's first t*TILE columns; the
                // diagonal tile additionally factors itself (diag = 2)
                emit_update_tiles(P.upd, p, t * TILE, kbeg, kend, nrow, pc.sn, 2);
            }
            lp.inner_end[t] = (int32_t)P.upd.size();
            lp.panel_begin[t] = (int32_t)P.panel.size();
            for (int32_t p : lpiece[L]) {
                const Piece& pc = P.pieces[p];
                const int32_t w = pc.c1 - pc.c0;
                if (t * TILE >= w) continue;
                const int32_t f = S.sn_first[pc.sn];
                const int32_t nrow = sn_nrow(S, pc.sn);
                const int32_t kend = (pc.c0 - f) + std::min(w, (t + 1) * TILE);
                for (int32_t r0 = kend; r0 < nrow; r0 += 2 * TILE) {
                    PanelTask pt;
                    pt.piece = p;
                    pt.step = t;
                    pt.r0 = r0;
                    pt.nr = std::min(2 * TILE, nrow - r0);
                    P.panel.push_back(pt);
                }
            }
            lp.panel_end[t] = (int32_t)P.panel.size();
        }
        // external updates: rows below the piece (rest of own supernode, then ancestors)
        lp.ext_begin = (int32_t)P.upd.size();
        for (int32_t p : lpiece[L]) {
            const Piece& pc = P.pieces[p];
            const int32_t s = pc.sn;
            const int32_t f = S.sn_first[s], l = S.sn_first[s + 1];
            const int32_t nrow = sn_nrow(S, s);
            const int32_t w = pc.c1 - pc.c0;
            if (pc.c1 < l) emit_update_tiles(P.upd, p, w, pc.c1 - f, l - f, nrow, s, 1);
            for (int64_t g = P.seg_ptr[s]; g < P.seg_ptr[s + 1]; ++g) {
                const int32_t kb = P.seg_k0[g];
                const int32_t ke = (g + 1 < P.seg_ptr[s + 1]) ? P.seg_k0[g + 1] : nrow;
                emit_update_tiles(P.upd, p, w, kb, ke, nrow, P.seg_tgt[g], 1);
            }
        }
        lp.ext_end = (int32_t)P.upd.size();
        lp.ext_atomic = (lpiece[L].size() > 1) ? 1 : 0;
        // solve tasks
        lp.solve_begin = (int32_t)P.solve.size();
        for (int32_t p : lpiece[L]) {
            const Piece& pc = P.pieces[p];
            const int32_t f = S.sn_first[pc.sn];
            const int32_t nrow = sn_nrow(S, pc.sn);
            for (int32_t r0 = pc.c1 - f; r0 < nrow; r0 += SOLVE_ROWS) {
                SolveTask st;
                st.piece = p;
                st.r0 = r0;
                st.nr = std::min(SOLVE_ROWS, nrow - r0);
                P.solve.push_back(st);
            }
        }
        lp.solve_end = (int32_t)P.solve.size();
    }
    // algorithmic flops of the tile updates: 2*kdim per structurally needed output entry
    auto tile_flops = [](const UpdTask& t) {
        double ent = (double)t.ni * t.nk;
        if (t.diag) ent -= (double)t.nk * (t.nk - 1) / 2.0;
        return 2.0 * ent * (double)t.kdim;
    };
    P.flops_update_inner = P.flops_update_ext = 0.0;
    for (const LevelPlan& lp : P.levels) {
        for (int32_t t = 0; t < lp.nsteps; ++t)
            for (int32_t x = lp.inner_begin[t]; x < lp.inner_end[t]; ++x) {
                P.flops_update_inner += tile_flops(P.upd[x]);
                if (P.upd[x].diag == 2) P.flops_update_inner += (double)P.upd[x].nk * P.upd[x].nk * P.upd[x].nk / 3.0;
            }
        for (int32_t x = lp.ext_begin; x < lp.ext_end; ++x) P.flops_update_ext += tile_flops(P.upd[x]);
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
