// Private definition of the opaque solver handle (shared by solver.cu and ipm.cu; not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/tlpb200.h"
#include "kernels.cuh"
#include "plan.hpp"
#include "symbolic.hpp"

struct tlpb200_ipm;   // device-resident HSD state (ipm.cu)

using namespace tlp;

struct tlpb200_solver {
    tlpb200_options opt;
    int system = 1;
    int64_t m = 0, n = 0, nnz = 0;
    // canonical 0-based CSC copy of A
    std::vector<int64_t> colptr;
    std::vector<int32_t> rowidx;
    std::vector<double> val;

    Symbolic sym;
    Plan plan;
    AssemblyMaps maps;

    bool on_device = false;
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, side_stream = nullptr;
    cudaStream_t side2[3] = {nullptr, nullptr, nullptr};   // further side streams: consecutive lazy batches overlap their ramp-down/up
    int nside = 1;                                         // TLPB200_SIDE_STREAMS (1..4)
    std::vector<cudaEvent_t> ev_f, ev_lazy;   // per level: chain done / lazy update batch done
    std::vector<cudaEvent_t> ev_d, ev_tr, ev_ur;   // per level: diagonal blocks done / rest of the trsm done / rest of the urgent tiles done
    cudaStream_t aux_stream = nullptr;        // non-critical part of the chain kernels (TLPB200_SPLIT_CHAIN)
    bool split_chain = true;
    int crit_ksplit = 4;          // CTAs per critical update tile (K range shared, RED accumulation; TLPB200_CRIT_KSPLIT)
    cudaEvent_t ev_pack = nullptr;
    std::vector<cudaEvent_t> ev_stage;   // one event per 256 KiB chunk of a device -> host result copy (pipelined staging)
    int pack_slice = 96;          // TLPB200_PACK_SLICE: repack tiles issued per level from the split level on
    double pack_split = 0.0;      // TLPB200_PACK_SPLIT: level fraction from which invert/repack slices are issued early (0 = off:
                                  // measured no gain on cfg2 -- the slices queue behind the 250 us diagonal-block inversions)

    std::vector<void*> allocs;
    size_t bytes_device = 0;
    DevCtx ctx{};
    DevCtx ctxA{}, ctxB{};        // multi-GPU phases: own subtrees / replicated top part
    const DevCtx* cur = nullptr;  // context the enqueue_* helpers launch with
    int rank = 0, nranks = 1;
    std::vector<int32_t> owner;   // [nsuper] rank owning each supernode, -1 = replicated top part
    int64_t top_begin = 0;        // offset of the top panels inside Lx
    std::vector<int64_t> rank_begin;   // [nranks + 1] panels are grouped by owner: rank r owns Lx[rank_begin[r], rank_begin[r + 1])
    int8_t* d_keep = nullptr;     // [N] 1 = this rank contributes wk[q] to the all-reduce
    // sharded phases: prefix counts of the items a phase really processes (phase 0 = own subtrees, 1 = replicated top part),
    // per item list, so that launches whose whole range is skipped on this rank are not issued at all
    enum WorkList { WL_SMALL = 0, WL_PIECE, WL_PANEL, WL_EXT, WL_LAZY, WL_FWD, WL_BWD, WL_FBIG, WL_BBIG, WL_BELOW, WL_INV, WL_PACK, WL_BSEQ, WL_COUNT };
    std::vector<int32_t> work_prefix[2][WL_COUNT];
    std::vector<int32_t> phase_need[2], phase_fpar[2], phase_bwait[2];   // merged-level sweeps: per-phase dependency targets
    // in-library collectives (tlpb200_comm_init): NCCL on the solver's stream, so that a sharded update!/solve! is ONE
    // stream-ordered (graph-captured) sequence without host synchronisation between its phases
    ncclComm_t comm = nullptr;
    int32_t ntop = 0;             // columns of the replicated top (separator) part
    const int32_t* d_top_cols = nullptr;   // [ntop] their permuted indices
    double* d_tbuf = nullptr;     // [ntop] packed top entries of wk for the small all-reduce of a solve
    int32_t* d_info_tmp = nullptr;   // [4] status words arranged for one max-all-reduce
    bool dist_graph = true;       // TLPB200_DIST_GRAPH=0: launch the sharded sequences without CUDA graphs
    DevMat mat{};
    double *d_theta = nullptr, *d_regP = nullptr, *d_regD = nullptr, *d_d = nullptr;
    double *d_xip = nullptr, *d_xid = nullptr, *d_dx = nullptr, *d_dy = nullptr;
    double* h_pin = nullptr;   // pinned staging: 2n + m (update) / (n+m) in + (n+m) out (solve)
    int32_t* h_info = nullptr; // pinned
    size_t small_smem = 0;
    int nsm = 148;
    int32_t* lazy_ctr = nullptr;   // [2*nlevels] work-queue / exit counters of the lazy update launches
    int chain_sms = 16;   // SMs kept free of the bulk-update work queue for the critical chain (TLPB200_CHAIN_SMS)
    int chain_sms_late = -1;      // same, from level chain_switch * nlevels on (TLPB200_CHAIN_SMS_LATE; -1 = same as chain_sms)
    double chain_switch = 0.5;    // TLPB200_CHAIN_SWITCH

    // tcgen05 int8 (Ozaki) path of the far Schur updates inside big all-positive supernodes (kernels_ozaki.cu)
    bool oz_on = false;
    uint8_t* oz_planes = nullptr;
    const int64_t* d_oz_rb_off = nullptr;
    OzView* d_oz_views = nullptr;
    const OzTask* d_oz_tasks = nullptr;
    int32_t* oz_E = nullptr;
    double* oz_scl = nullptr;
    int32_t* oz_ctr = nullptr;                 // [2*nlevels] work-queue counters of the tcgen05 launches
    cudaStream_t oz_slice_stream = nullptr, oz_stream = nullptr;
    std::vector<cudaEvent_t> ev_ozt, ev_ozs, ev_oz;   // per level: trsm done (main stream) / digit planes written / tasks done
    int oz_sms_free = 16;                      // SMs the tcgen05 work queue leaves to the chain kernels (TLPB200_OZAKI_FREE_SMS)

    cudaGraphExec_t g_update = nullptr, g_solve = nullptr;
    bool profiling = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};

    // per-kernel-class profiling (profiling mode only)
    std::vector<cudaEvent_t> pool;
    std::vector<int> pool_cls;
    size_t pool_used = 0;
    bool merge_levels = true;     // single GPU: block-solve items of consecutive levels share one launch (TLPB200_MERGE_LEVELS=0: per level)
    int scope_depth = 0;          // nesting depth of the profiling brackets (only the outermost records)
    double ms_class[TLPB200_NCLASS] = {0};
    int64_t n_class[TLPB200_NCLASS] = {0};

    int64_t launches_update = 0, launches_solve = 0;
    double ms_assemble = 0, ms_factor = 0, ms_solve = 0;
    int64_t bad_pivot = -1, n_update = 0, n_solve = 0;
    // dense columns (K1): kept out of the sparse factor, applied as a low-rank Schur correction
    std::vector<int32_t> dense_cols;
    std::vector<int64_t> dc_colptr;
    DenseCols dc{};
    double *dc_xi = nullptr, *dc_y = nullptr, *dc_tn = nullptr;   // refinement work vectors
    int dc_refine = 2;
    int refine = 0;               // general iterative-refinement steps inside solve! (tlpb200_options::refine_steps; 0 = off)
    std::string err;
    tlpb200_ipm* ipm = nullptr;   // device-resident IPM state (tlpb200_hsd_create); owned, freed by tlpb200_destroy
};

namespace tlp_internal {
void run_update(tlpb200_solver* s);                            // enqueue update! from d_theta / d_regP / d_regD (throws std::runtime_error)
int finish_update(tlpb200_solver* s, int64_t* bad_pivot);      // synchronise + status of the last update!
void run_solve(tlpb200_solver* s);                             // enqueue solve! d_xip / d_xid -> d_dx / d_dy
int check_timeouts(tlpb200_solver* s, const char* where);      // after a D2H of info into h_info + sync
int set_error(tlpb200_solver* s, int code, const std::string& msg);
void* device_alloc(tlpb200_solver* s, size_t bytes);           // freed with the solver
}  // namespace tlp_internal
