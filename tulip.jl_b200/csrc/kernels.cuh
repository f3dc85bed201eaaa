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
    const PanelTask* panel;
    const SolveTask* solve;
    // numeric state
    double* Lx;
    int32_t* info;   // info[0] = smallest permuted column with a bad pivot (INT_MAX if none)
    double* wk;      // [N] work vector of the triangular solves (permuted order)
    double* acc;     // [N] accumulator for the backward gemv of wide pieces (kept zero between uses)
    int32_t N;
};

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
};

// ---- launchers (all asynchronous on `st`) ---------------------------------------------------
void launch_compute_d(const double* theta, const double* regP, double* d, int64_t n, cudaStream_t st);
void launch_assemble_k1(const DevCtx& c, const DevMat& A, const double* d, const double* regD, cudaStream_t st);
void launch_assemble_k2(const DevCtx& c, const DevMat& A, const double* theta, const double* regP, const double* regD,
                        cudaStream_t st);
void launch_small_factor(const DevCtx& c, int32_t begin, int32_t end, size_t smem, cudaStream_t st);
void launch_update(const DevCtx& c, int32_t begin, int32_t end, int atomic, cudaStream_t st);
void launch_trsm(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);

void launch_fwd_small(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);
void launch_bwd_small(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);
void launch_fwd_trsv(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);
void launch_fwd_gemv(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);
void launch_bwd_gemv(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);
void launch_bwd_trsv(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st);

void launch_k1_rhs(const DevCtx& c, const DevMat& A, const double* d, const double* xi_p, const double* xi_d,
                   cudaStream_t st);
void launch_k1_recover(const DevCtx& c, const DevMat& A, const double* d, const double* xi_d, double* dx, double* dy,
                       cudaStream_t st);
void launch_k2_rhs(const DevCtx& c, const DevMat& A, const double* xi_p, const double* xi_d, cudaStream_t st);
void launch_k2_recover(const DevCtx& c, const DevMat& A, double* dx, double* dy, cudaStream_t st);

size_t small_factor_smem(int32_t max_elems, int32_t max_nrow);
void kernels_static_init();   // cudaFuncSetAttribute calls

}  // namespace tlp
