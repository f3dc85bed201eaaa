// Static execution plan for the numeric phase: because the sparsity structure is fixed between
// IPM iterations (the reference re-runs only the numeric part in update!,
// /root/reference/src/KKT/Cholmod/spd.jl:46, sqd.jl:53), everything that decides *what* runs
// *where* is computed once here on the host and uploaded; update!/solve! replay it.
#pragma once
#include <cstdint>
#include <vector>

#include "symbolic.hpp"

namespace tlp {

constexpr int TILE = 64;          // update tile edge / panel inner block width
constexpr int SOLVE_ROWS = 256;   // rows per CTA in the solve gemv kernels

struct PlanOptions {
    int small_elems = 4096;   // supernodes with nrow*ncol <= small_elems (and ncol <= small_ncol) run in one CTA
    int small_ncol = 32;
    int piece_width = 256;    // wide supernodes are processed in column pieces of at most this width
};

// One column piece [c0,c1) of supernode sn (global permuted column indices).
struct Piece {
    int32_t sn, c0, c1, level;
};

// C_tgt[pos(rows I), cols K] -= L_piece[I, 0:kdim] * S * L_piece[K, 0:kdim]'
struct UpdTask {
    int32_t piece;   // source piece
    int32_t kdim;    // number of leading columns of the piece that take part
    int32_t i0, ni;  // I block: indices into the source supernode's row list
    int32_t k0, nk;  // K block: indices into the source supernode's row list (all rows map to columns of tgt)
    int32_t tgt;     // target supernode
    int32_t diag;    // 1 = I and K blocks overlap (keep only row >= col)
};

// potrf of the diagonal block of inner step `step` of a piece + trsm of one row tile below it
struct PanelTask {
    int32_t piece, step;
    int32_t r0, nr;   // row tile: indices into the supernode's row list (rows strictly below the diagonal block)
};

// solve: one row tile of the rows below a piece (gemv / gemv-transposed)
struct SolveTask {
    int32_t piece;
    int32_t r0, nr;
};

struct LevelPlan {
    // ranges into the flat arrays of Plan
    int32_t small_begin = 0, small_end = 0;        // Plan::small_list
    int32_t piece_begin = 0, piece_end = 0;        // Plan::level_pieces
    int32_t nsteps = 0;                            // max inner steps over this level's pieces
    std::vector<int32_t> inner_begin, inner_end;   // [nsteps] ranges into Plan::upd (step 0 empty)
    std::vector<int32_t> panel_begin, panel_end;   // [nsteps] ranges into Plan::panel
    int32_t ext_begin = 0, ext_end = 0;            // external update tasks (Plan::upd)
    int32_t ext_atomic = 1;
    int32_t solve_begin = 0, solve_end = 0;        // Plan::solve tasks of this level's pieces
};

struct Plan {
    PlanOptions opt;
    std::vector<Piece> pieces;
    std::vector<int32_t> sn_small;        // [nsuper] 1 = handled by the one-CTA kernels
    std::vector<int32_t> sn_level;        // [nsuper] level of the supernode's last item
    // target segments of each supernode's below rows: consecutive below rows that are columns of the same target
    std::vector<int64_t> seg_ptr;         // [nsuper+1]
    std::vector<int32_t> seg_k0;          // first index (into the supernode's row list) of the segment
    std::vector<int32_t> seg_tgt;         // target supernode
    std::vector<int32_t> small_list;      // supernode ids grouped by level
    std::vector<int32_t> level_pieces;    // piece ids grouped by level
    std::vector<UpdTask> upd;
    std::vector<PanelTask> panel;
    std::vector<SolveTask> solve;
    std::vector<LevelPlan> levels;
    int32_t max_small_elems = 0;          // largest nrow*ncol among small supernodes
    int32_t max_small_nrow = 0;
    double flops_update_inner = 0.0, flops_update_ext = 0.0;   // algorithmic (lower-triangle) flops of the tile updates
};

void build_plan(const Symbolic& S, const PlanOptions& opt, Plan& P);

// ---- assembly maps ----------------------------------------------------------------------
// K1:  Lx[w_dest[e]] = sum_{p in [w_ptr[e], w_ptr[e+1])} w_val[p] * d[w_col[p]]
//      (one "entry" e per structural non-zero of lower(A*A'), permuted), then Lx[diagpos[q]] += regD[perm[q]].
// K2:  Lx[a_dest[p]] = Ax[p]; Lx[diagpos[q]] = v<n ? -(theta[v]+regP[v]) : regD[v-n]  with v = perm[q].
struct AssemblyMaps {
    std::vector<int64_t> w_ptr, w_dest;
    std::vector<int32_t> w_col;
    std::vector<double> w_val;
    std::vector<int64_t> a_dest;
};

void build_assembly_k1(const Symbolic& S, int64_t m, int64_t n, const int64_t* colptr, const int32_t* rowidx,
                       const double* val, AssemblyMaps& M);
void build_assembly_k2(const Symbolic& S, int64_t m, int64_t n, const int64_t* colptr, const int32_t* rowidx,
                       AssemblyMaps& M);

// position of permuted entry (row gi, column gk), gi >= gk, inside Lx
int64_t lx_position(const Symbolic& S, int32_t gi, int32_t gk);

}  // namespace tlp
