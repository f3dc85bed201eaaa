// Static execution plan for the numeric phase: because the sparsity structure is fixed between
// IPM iterations (the reference re-runs only the numeric part in update!,
// /root/reference/src/KKT/Cholmod/spd.jl:46, sqd.jl:53), everything that decides *what* runs
// *where* is computed once here on the host and uploaded; update!/solve! replay it.
#pragma once
#include <cstdint>
#include <vector>

#include "symbolic.hpp"

namespace tlp {

constexpr int TILE = 64;      // update tile edge (urgent tiles, main stream)
constexpr int TILE128 = 128;  // update tile edge of the persistent side-stream kernel
constexpr int PIECE = 128;    // column-piece width of wide supernodes = solve block size
constexpr int SBLK = 128;     // block size of the dense triangular-solve kernels (== PIECE)
constexpr int BELOW_ROWS = 512;   // rows per backward "below" item
constexpr int SPLIT_ROWS = 384;   // non-big supernodes with at least this many rows below their columns are split

struct PlanOptions {
    int small_elems = 4096;   // supernodes with nrow*ncol <= small_elems, ncol <= small_ncol and
    int small_ncol = 32;      // nrow <= small_nrow run in the one-CTA kernels
    int small_nrow = 256;
    int big_ncol = 384;       // non-small supernodes with >= big_ncol columns use the dense-solve path (BigTask)
    int oz_ncol = 1024;       // all-positive supernodes with >= oz_ncol columns: far Schur updates on the tcgen05 int8 path; <= 0 = off
    int oz_tile_n = 0;        // 64: 128x64 tiles, one pass | 128: 128x128 tiles, two passes over K (k_oz_update2) | 0: per level,
                              // 128 when the launch has at least two waves of 128x128 tasks (measured: 68 vs 57 TF on large
                              // launches, but coarser tasks lose on small ones)
    int oz_ksplit = 2048;     // columns of K accumulated per tcgen05 task (multiple of 32, <= 4096: exact int32 accumulation)
};

// One column piece [c0,c1) of supernode sn (global permuted column indices), c1-c0 <= PIECE.
struct Piece {
    int32_t sn, c0, c1, level;
};

// C_tgt[pos(rows I), cols K] -= L_piece[I, :] * S * L_piece[K, :]'
struct UpdTask {
    int32_t piece;   // source piece (all of its columns take part)
    int32_t i0, ni;  // I block: indices into the source supernode's row list
    int32_t k0, nk;  // K block: indices into the source supernode's row list (all rows map to columns of tgt)
    int32_t tgt;     // target supernode
    int32_t diag;    // 1 = I and K blocks start at the same row (keep only row >= col)
    int32_t pad;
};

// triangular solve of one row tile (<= 128 rows) below a piece's factored diagonal block
struct PanelTask {
    int32_t piece;
    int32_t r0, nr;   // indices into the supernode's row list
    int32_t pad;
};

// dense block solve work item: block `blk` of supernode sn; kind 0 = diagonal block (blk-th
// 128-block of the columns), kind 1 = block of rows below the columns (forward sweep only)
struct SolveItem {
    int32_t sn, blk, kind;
    int32_t r0, nr;   // row range of the block inside the supernode's row list
    int32_t pad;
};

// ---- dense-solve ("big") supernodes -------------------------------------------------------------
// After the factorisation the panel of a big supernode is repacked into 128x128 tiles of
// Lhat = L * blockdiag(L_kk)^{-1} (unit block diagonal), each tile contiguous (16384 doubles, zero padded):
//   Ft: tile (I, j) column-major   [c*128 + r] = Lhat[I.r0 + r, j*128 + c]   (forward sweep, block row I)
//   Bt: the same tile row-major    [r*128 + c]                               (backward sweep, block column j;
//       only the tiles inside the supernode's columns: the rows below are BelowItems on the unscaled panel)
// The tiles of one solve task are contiguous in the order the task streams them.
struct BigPack {
    int32_t sn;
    int32_t r0, nr;   // row range of the tile inside the supernode's row list
    int32_t j;        // column block
    int64_t fdst;     // tile index inside Ft
    int64_t bdst;     // tile index inside Bt (-1: none)
};

struct BigTask {
    int32_t sn;
    int32_t kind;     // forward: 0 = block row inside the columns, 1 = block of rows below.  backward: 0
    int32_t blk;      // column block (kind 0) / below block (kind 1)
    int32_t r0, nr;   // row range inside the supernode's row list
    int32_t ntile;    // tiles streamed: forward kind 0: blk, kind 1: ncb; backward: ncb-1-blk
    int32_t nbelow;   // unused (the rows below the columns are BelowItems in the backward sweep)
    int32_t xq0;      // first exchange slot of the supernode (slot = xq0 + column offset inside the supernode)
    int64_t tile0;    // first tile inside Ft / Bt
};

// ---- tcgen05 int8 (Ozaki) path of the far same-supernode updates (kernels_ozaki.cu) ---------------
// Column block c of an "oz" supernode receives the contributions of its pieces 0 .. c-3 in ONE left-looking pass
// (launched when piece c-3 is done, needed at level of piece c); pieces c-2 and c-1 stay on the FP64 DMMA path.
// C[rbA*128 .. +128, rbB*128 + half*64 .. +64] -= L[rows A, 32*k0 .. 32*k1) * L[rows B, same]'   (lower part only)
struct OzTask {
    int32_t view;
    int32_t rbA, rbB, half;
    int32_t k0, k1;           // K-chunk range (32 columns each)
    int32_t pad[2];
};
struct OzViewPlan {
    int32_t sn;
    int32_t nrb, ncb;         // 128-row blocks of the whole panel / 128-column blocks
    int64_t off0;             // first entry of this view inside Plan::oz_rb_off
    int64_t row0;             // first entry of this view's rows inside the E / scale arrays
    int32_t base_level;       // level of the supernode's first piece (row scales are taken there)
    int32_t pad;
};
struct OzSlice {
    int32_t view, piece;      // piece id (Plan::pieces)
    int32_t j;                // index of the piece inside its supernode
    int32_t pad;
};

// backward sweep, rows below the columns of a tall supernode: p = L[rows, block]' x[rows] is accumulated into
// DevCtx::bacc by a separate, fully parallel launch (k_bwd_below) before the level's block solves run
struct BelowItem {
    int32_t sn, blk;   // supernode, column block
    int32_t r0, nr;    // row range inside the supernode's row list (r0 >= ncol)
};

struct LevelPlan {
    int32_t small_begin = 0, small_end = 0;    // Plan::small_list
    int32_t piece_begin = 0, piece_end = 0;    // Plan::level_pieces (diag-factor CTAs)
    int32_t panel_begin = 0, panel_end = 0;    // Plan::panel (trsm row tiles)
    int32_t panel_crit_end = 0;                // [panel_begin, panel_crit_end): row tiles read by the critical update tiles
    int32_t ext_begin = 0, ext_end = 0;        // Plan::upd
    int32_t ext_crit_end = 0;                  // [ext_begin, ext_crit_end): tiles that land in a diagonal block / small
                                               // supernode of the next level (they gate its first kernels)
    int32_t urgent_end = 0;                    // == ext_end (all 64x64 tiles are urgent)
    int32_t lazy_begin = 0, lazy_end = 0;      // Plan::upd128 (128x128 tiles, side stream)
    int32_t ext_atomic = 1;
    int32_t fwd_begin = 0, fwd_end = 0;        // Plan::fwd_items (supernodes whose first piece is at this level)
    int32_t bwd_begin = 0, bwd_end = 0;        // Plan::bwd_items
    int32_t fbig_begin = 0, fbig_end = 0;      // Plan::fwd_big (big supernodes that start at this level)
    int32_t bbig_begin = 0, bbig_end = 0;      // Plan::bwd_big
    int32_t below_begin = 0, below_end = 0;    // Plan::bwd_below
    int32_t oz_begin = 0, oz_end = 0;          // Plan::oz_tasks launched when this level's pieces are done (needed at level + 3)
    int32_t ozs_begin = 0, ozs_end = 0;        // Plan::oz_slices: pieces of this level whose digit planes are needed
    int32_t oz_tile = 64;                      // output tile width of this level's tcgen05 launch (64 | 128)
    int32_t inv_end = 0;                       // Plan::inv_order[0, inv_end): diagonal blocks of pieces at levels <= this one
    int32_t pack_end = 0;                      // Plan::big_pack[0, pack_end): tiles whose column block is at a level <= this one
};

// One launch of a triangular sweep (single-GPU path).  The block-solve items of consecutive levels that are not separated
// by a launch of another kind are merged into ONE persistent launch: inside it a supernode's items wait on counters instead
// of on a kernel boundary (forward: all items of its in-launch children have finished; backward: all items of its in-launch
// parent have finished), so a chain of L levels costs L flag hand-overs instead of L launches.
struct SolveOp {
    int32_t kind;          // 0 small, 1 large (merged), 2 big, 3 below
    int32_t begin, end;    // range in small_list / fwd_items or bwd_seq / fwd_big or bwd_big / bwd_below
    int32_t level;         // level of the op (first level of a merged run)
};

struct Plan {
    PlanOptions opt;
    std::vector<SolveOp> fwd_ops, bwd_ops;   // launch sequences of the forward / backward sweep
    std::vector<SolveItem> bwd_seq;          // bwd_items re-ordered: levels descending (the order the backward sweep walks them)
    std::vector<int32_t> fwd_need;           // [nsuper] forward: items of in-launch children that must finish first
    std::vector<int32_t> fwd_parent;         // [nsuper] forward: in-launch parent to notify (-1: none)
    std::vector<int32_t> bwd_wait;           // [nsuper] backward: in-launch parent to wait for (-1: none)
    std::vector<int32_t> bwd_nitems;         // [nsuper] backward items of the supernode (= its column blocks)
    std::vector<int32_t> bwd_nbelow;         // [nsuper] below items (kind 2) of a block-solve supernode inside the merged sequence
    std::vector<Piece> pieces;
    std::vector<int32_t> sn_small;        // [nsuper] 1 = handled by the one-CTA kernels
    std::vector<int32_t> sn_level;        // [nsuper] level of the supernode's last item
    std::vector<int32_t> sn_dblk;         // [nsuper] index of the supernode's first diagonal block (-1 if small)
    int32_t ndblk = 0;                    // total diagonal blocks of non-small supernodes
    std::vector<int32_t> dblk_sn, dblk_idx;  // [ndblk] owner supernode / block index inside it
    // target segments of each supernode's below rows: consecutive below rows that are columns of the same target
    std::vector<int64_t> seg_ptr;         // [nsuper+1]
    std::vector<int32_t> seg_k0;          // first index (into the supernode's row list) of the segment
    std::vector<int32_t> seg_tgt;         // target supernode
    std::vector<int32_t> small_list;      // supernode ids grouped by level
    std::vector<int32_t> level_pieces;    // piece ids grouped by level
    std::vector<UpdTask> upd;
    std::vector<UpdTask> upd128;
    std::vector<PanelTask> panel;
    std::vector<SolveItem> fwd_items, bwd_items;
    std::vector<int32_t> sn_big;          // [nsuper] 1 = dense-solve path
    std::vector<BigPack> big_pack;
    std::vector<BigTask> fwd_big, bwd_big;
    std::vector<int32_t> inv_order;       // diagonal blocks (== piece ids) sorted by level
    std::vector<int32_t> sn_split;        // [nsuper] 1 = the rows below the columns are handled by BelowItems in the backward sweep
    std::vector<BelowItem> bwd_below;
    int64_t n_ftiles = 0, n_btiles = 0;   // tiles of Ft / Bt
    int32_t xq_slots = 0;                 // exchange slots (sum over big supernodes of ncb*128)
    std::vector<int32_t> sn_oz;           // [nsuper] index into oz_views or -1
    std::vector<OzViewPlan> oz_views;
    std::vector<int64_t> oz_rb_off;       // per view nrb+1 entries: first K-chunk slot (32 KiB each) of every row block
    int64_t oz_slots = 0;                 // total K-chunk slots of the digit planes
    int64_t oz_rows = 0;                  // total panel rows of the oz views
    std::vector<OzTask> oz_tasks;
    std::vector<OzSlice> oz_slices;
    double flops_oz = 0.0;                // algorithmic (lower-triangle) flops of the tcgen05 tasks
    std::vector<LevelPlan> levels;
    int32_t max_small_elems = 0;          // largest nrow*ncol among small supernodes
    int32_t max_small_nrow = 0;
    double flops_update = 0.0;            // algorithmic (lower-triangle) flops of the tile updates
    double flops_panel = 0.0;             // diag-block factor + trsm flops
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

// multi-GPU sharding of independent elimination-tree subtrees (see plan.cpp)
void partition_subtrees(const Symbolic& S, int nranks, std::vector<int32_t>& owner, std::vector<double>* rank_work = nullptr);
int64_t relayout_panels(Symbolic& S, const std::vector<int32_t>& owner, int32_t nranks, std::vector<int64_t>* rank_begin = nullptr);   // returns the offset of the top part in Lx

// position of permuted entry (row gi, column gk), gi >= gk, inside Lx
int64_t lx_position(const Symbolic& S, int32_t gi, int32_t gk);

}  // namespace tlp
