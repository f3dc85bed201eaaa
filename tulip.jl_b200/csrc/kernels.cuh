// Device-side context and kernel launchers of the numeric phase (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "plan.hpp"

namespace tlp {

struct DevCtx {
    // symbolic structure
    const int32_t* sn_first;
    const int64_t* sn_rowptr;
    const int32_t* sn_rows;
    const int64_t* sn_xptr;
    const int32_t* col2sn;
    const int8_t* sign;
    const int64_t* diagpos;
    const int32_t* perm;
    const int32_t* iperm;
    // plan
    const Piece* pieces;
    const int64_t* seg_ptr;
    const int32_t* seg_k0;
    const int32_t* seg_tgt;
    const int32_t* small_list;
    const int32_t* level_pieces;
    const UpdTask* upd;
    const UpdTask* upd_lazy;
    const PanelTask* panel;
    const SolveItem* fwd_items;
    const SolveItem* bwd_items;
    const int32_t* sn_dblk;    // [nsuper] first diagonal block of a non-small supernode
    const int32_t* dblk_sn;    // [ndblk]
    const int32_t* dblk_idx;   // [ndblk]
    const int32_t* inv_order;  // [ndblk] diagonal blocks sorted by level
    // numeric state
    double* Lx;
    double* Dinv;    // [ndblk][128*128] explicit inverses of the diagonal blocks (column-major, lower)
    double* DinvT;   // [ndblk][128*128] their transposes (column-major, upper)
    double* LsubT;   // [ndblk][128*128] transposed sub-diagonal tile of each diagonal block (backward sweep)
    int32_t* flags;  // [2*ndblk] forward / backward "block solved" flags of the dense solve kernels
    int32_t* info;   // info[0] = smallest permuted column with a bad pivot (INT_MAX if none)
    double* wk;      // [N] work vector of the triangular solves (permuted order)
    int32_t N;
    int32_t ndblk;
    int32_t has_neg;   // 1 when some pivots are expected negative (K2)
    // dense-solve path of the big supernodes (kernels_dense_solve.cu)
    const BigPack* big_pack;
    const BigTask* fwd_big;
    const BigTask* bwd_big;
    double* Ft;                    // forward tiles  (column-major 128x128, contiguous)
    double* Bt;                    // backward tiles (row-major 128x128, contiguous)
    unsigned long long* xq;        // exchange slots: {bits(x), bits(x) ^ key(epoch)} per entry
    unsigned long long* epoch;     // [0] sweep counter (bumped by the last CTA of every dense sweep), [1] exit counter
    unsigned long long* dbg_ts;    // [2 * nxblk] globaltimer at every block publish (forward, then backward); nullptr = off
    int32_t nxblk;                 // xq_slots / 128
    unsigned long long* trace_min; // [nlevels][4] first CTA start of {diag, trsm, urgent update, lazy update} (nullptr = off)
    unsigned long long* trace_max; // [nlevels][4] last CTA end
    const BelowItem* bwd_below;
    const int32_t* sn_split;       // [nsuper] 1 = rows below the columns are accumulated into bacc by k_bwd_below
    double* bacc;                  // [N] backward accumulator sum_rows L[row, c] x[row]; consumers re-zero their entries
    // merged-level sweeps (Plan::SolveOp): per-supernode dependency counters and their static targets
    int32_t* dep_cnt;              // [3 * nsuper] forward: finished items of in-launch children / backward: finished items /
                                   // backward: finished below items; zeroed per solve
    const int32_t* bwd_nbelow;     // [nsuper]
    const int32_t* fwd_need;       // [nsuper]
    const int32_t* fwd_parent;     // [nsuper]
    const int32_t* bwd_wait;       // [nsuper]
    const int32_t* bwd_nitems;     // [nsuper]
    const SolveItem* bwd_seq;      // backward items, levels descending
    int32_t nsuper;
    const int8_t* skip;   // multi-GPU: skip[s] != 0 -> supernode s is not processed in this phase on this rank (nullptr: none)
};

#ifdef __CUDACC__
// factorisation timeline (TLPB200_TRACE_FACTOR): per level and kernel class, globaltimer of the first start / last end
__device__ __forceinline__ void trace_mark(const DevCtx& c, int level, int cls, bool end) {
    if (!c.trace_min) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    if (end) atomicMax(c.trace_max + 4 * level + cls, t);
    else atomicMin(c.trace_min + 4 * level + cls, t);
}
#endif

// matrix A on the device (CSC + CSR copies) and the assemble maps
struct DevMat {
    int64_t m, n, nnz;
    const int64_t* colptr;  // CSC
    const int32_t* rowidx;
    const double* val;
    const int64_t* rowptr;  // CSR
    const int32_t* colidx;
    const double* rval;
    // K1 assemble
    int64_t nentries;
    const int64_t* w_ptr;
    const int64_t* w_dest;
    const int32_t* w_col;
    const double* w_val;
    // K2 assemble
    const int64_t* a_dest;
    // columns with more than LONG_COL non-zeros (the dense columns of BASELINE config 5): the per-column kernels leave them
    // to one CTA each instead of one thread
    const int32_t* long_cols;   // [nlong] sorted
    int32_t nlong;
};
constexpr int LONG_COL = 64;

// dense columns of A handled outside the sparse factorisation (K1 only, see kernels_dense_cols.cu)
struct DenseCols {
    int32_t nd;
    const int64_t* colptr;   // [nd+1] into prow / val
    const int32_t* prow;     // permuted row index of every entry
    const double* val;
    const int32_t* col_id;   // [nd] original column index
    double* Wt;              // [nd][N]   forward-swept dense columns as the sweep kernels leave them: G L^{-1} P A_d
    double* Wh;              // [nd][N]   (G G')^{-1} Wt, so that Wh' Wt = W'W with W = L^{-1} P A_d (see kernels_dense_cols.cu)
    double* C;               // [nd*nd]   D_d^{-1} + W'W, then its Cholesky factor (row-major lower)
    double* g;               // [nd]
    const int32_t* gblk;     // [3 * ngblk] {diagonal-block index, first permuted column, width} of the dense-solve supernodes' blocks
    int32_t ngblk;
};
void launch_dc_ginv(const DevCtx& c, const DenseCols& dc, const double* in, double* out, cudaStream_t st);
void launch_dc_scatter(const DevCtx& c, const DenseCols& dc, int j, int64_t colnnz, cudaStream_t st);
void launch_dc_gram_chol(const DevCtx& c, const DenseCols& dc, const double* theta, const double* regP, cudaStream_t st);
void launch_dc_apply(const DevCtx& c, const DenseCols& dc, cudaStream_t st);
struct DevMat;
void launch_dc_residual(const DevCtx& c, const DevMat& A, const double* d, const double* regD, const double* xi, const double* y,
                        double* tn, cudaStream_t st);
void launch_dc_axpy(const DevCtx& c, const double* y, cudaStream_t st);
void launch_k2_residual(const DevCtx& c, const DevMat& A, const double* theta, const double* regP, const double* regD, const double* xi,
                        const double* y, cudaStream_t st);

// ---- launchers (all asynchronous on `st`) ---------------------------------------------------
void launch_compute_d(const double* theta, const double* regP, double* d, int64_t n, cudaStream_t st);
void launch_assemble_k1(const DevCtx& c, const DevMat& A, const double* d, const double* regD, cudaStream_t st);
void launch_assemble_k2(const DevCtx& c, const DevMat& A, const double* theta, const double* regP, const double* regD,
                        cudaStream_t st);
void launch_small_factor(const DevCtx& c, int32_t begin, int32_t end, size_t smem, cudaStream_t st);
void launch_diag_factor(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);
void launch_trsm(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);
void launch_update(const DevCtx& c, int32_t begin, int32_t end, int atomic, cudaStream_t st, int ksplit = 1);
void launch_update_lazy(const DevCtx& c, int32_t begin, int32_t end, int32_t* counter, int nsm, int reserve, cudaStream_t st);
void launch_invert_diag(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);   // range of DevCtx::inv_order

void launch_fwd_small(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);
void launch_bwd_small(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);
// merged = 1: [begin, end) spans several levels of fwd_items / bwd_seq and the items synchronise through the dependency
// counters; merged = 0: one level of fwd_items / bwd_items (the kernel boundary is the synchronisation; sharded phases)
void launch_fwd_large(const DevCtx& c, int32_t begin, int32_t end, int nsm, int merged, cudaStream_t st);
void launch_bwd_large(const DevCtx& c, int32_t begin, int32_t end, int nsm, int merged, cudaStream_t st);
void launch_bwd_below(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);

void launch_pack_big(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);
void launch_fwd_big(const DevCtx& c, int32_t begin, int32_t end, int nsm, cudaStream_t st);
void launch_bwd_big(const DevCtx& c, int32_t begin, int32_t end, int nsm, cudaStream_t st);
cudaError_t dense_solve_static_init();

void launch_gather_top(const double* wk, const int32_t* cols, int32_t ntop, double* buf, cudaStream_t st);
void launch_scatter_top(double* wk, const int32_t* cols, int32_t ntop, const double* buf, cudaStream_t st);
void launch_pack_info(const int32_t* info, int32_t* tmp, cudaStream_t st);
void launch_unpack_info(int32_t* info, const int32_t* tmp, cudaStream_t st);
void launch_zero_unowned(const DevCtx& c, const int8_t* keep, cudaStream_t st);   // wk[q] = 0 where keep[q] == 0
void launch_k1_rhs(const DevCtx& c, const DevMat& A, const double* d, const double* xi_p, const double* xi_d,
                   cudaStream_t st);
void launch_k1_recover(const DevCtx& c, const DevMat& A, const double* d, const double* xi_d, double* dx, double* dy,
                       cudaStream_t st);
void launch_k2_rhs(const DevCtx& c, const DevMat& A, const double* xi_p, const double* xi_d, cudaStream_t st);
void launch_k2_recover(const DevCtx& c, const DevMat& A, double* dx, double* dy, cudaStream_t st);

// ---- Ozaki / tcgen05 int8 path of the big supernodes' Schur updates (kernels_ozaki.cu) -----------
constexpr int OZ_S = 8;   // 7-bit digit planes per FP64 value

// one dense square part (first `nrows` rows x columns of a big supernode's panel, or a test matrix)
struct OzView {
    const uint8_t* planes;    // digit planes: [row block][K chunk of 32][plane][4096 B]
    const int64_t* rb_off;    // [row blocks + 1] first K-chunk slot of each row block (units of OZ_S * 4096 B)
    double* C;                // target panel, column-major
    int64_t ldc;
    const double* scl;        // [nrows] 2^E_r of every panel row
    int32_t nrows;            // rows of the panel
    int32_t ncols;            // columns of the supernode (targets are columns < ncols)
};

cudaError_t ozaki_static_init();
void launch_oz_rowexp(const double* Lx, const int64_t* diagpos, const int32_t* grow, int32_t nrows, int32_t* E, double* scl,
                      cudaStream_t st);
void launch_oz_slice(const double* panel, int64_t ld, int32_t nrows, int32_t c0, int32_t w, int32_t rb0, int32_t nrb, int32_t kchunk0,
                     const int32_t* E, const int64_t* rb_off, uint8_t* planes, cudaStream_t st);
void launch_oz_update(const OzView* views, const OzTask* tasks, int32_t begin, int32_t end, int32_t* counter, int nsm, int reserve,
                      int32_t* err, int tile_n, cudaStream_t st);   // tile_n: 64 (one pass, 128x64 tiles) | 128 (two passes, 128x128)

size_t small_factor_smem(int32_t max_elems, int32_t max_nrow);
cudaError_t kernels_static_init();   // cudaFuncSetAttribute calls
cudaError_t factor_kernels_static_init();

}  // namespace tlp
