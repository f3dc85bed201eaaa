/* tlpb200.h -- C ABI of the B200-native KKT backend for Tulip.jl's IPM.
 *
 * The reference has no FFI: its KKT plug-in boundary is Julia multiple dispatch
 * (/root/reference/src/KKT/KKT.jl:59 setup, :83 update!, :100 solve!, :114 backend,
 * :121 linear_system).  Each entry point below is what a Julia `ccall` (see
 * tulip.jl_b200/julia/TlpB200.jl and INTEGRATION.md) binds to implement one of those methods
 * for a new `TlpB200.Backend <: AbstractKKTBackend`.
 *
 * Conventions
 *   - A is CSC: colptr[n+1], rowval[nnz], nzval[nnz]; 64-bit indices as in Julia's
 *     SparseMatrixCSC{Float64,Int} (src/LinearAlgebra/LinearAlgebra.jl:16-32); `index_base`
 *     is 1 when called from Julia, 0 from C/Python.
 *   - All vectors are Float64.  Host-pointer calls copy in/out through pinned staging inside the
 *     library and are fully synchronised on return (results are read on the host right away,
 *     src/IPM/HSD/step.jl:225-237).  `_dev` variants take device pointers and only enqueue work
 *     on the solver's stream.
 *   - theta_inv / regP / regD are copied on entry (src/KKT/Cholmod/spd.jl:36-38); xi_p, xi_d are
 *     read-only (the caller passes dat.b itself, src/IPM/HSD/step.jl:63).
 *   - No CPU fallback: every numeric entry point fails with TLPB200_CUDA if no sm_100 device
 *     is usable.
 */
#ifndef TLPB200_H
#define TLPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tlpb200_solver tlpb200_solver; /* opaque */

enum {
    TLPB200_OK = 0,
    TLPB200_NOT_POSDEF = 1, /* wrong-sign / zero / NaN pivot -> Julia PosDefException (spd.jl:47, ldlfact.jl:115) */
    TLPB200_OOM = 2,        /* host or device allocation failed -> Julia OutOfMemoryError (HSD.jl:327) */
    TLPB200_BAD_ARG = 3,    /* dimension / argument error -> Julia DimensionMismatch (spd.jl:26-34) */
    TLPB200_CUDA = 4,       /* CUDA runtime error or no usable device -> ErrorException */
    TLPB200_INTERNAL = 5,
    TLPB200_NCCL = 6        /* NCCL error (multi-GPU collectives) -> ErrorException */
};

enum { TLPB200_K1 = 1, TLPB200_K2 = 2 }; /* src/KKT/systems.jl:32 (K2), :54 (K1) */

typedef struct tlpb200_options {
    int32_t ordering;      /* 0 natural, 1 approximate minimum degree (default 1) */
    int32_t device;        /* CUDA device ordinal (default 0) */
    int32_t piece_width;   /* column-piece width of wide supernodes (fixed: 128) */
    int32_t small_elems;   /* supernodes with nrow*ncol <= this run in the one-CTA kernels (default 4096) */
    int32_t relax_always;  /* amalgamation: always merge when merged width <= this (default 8) */
    int32_t use_graph;     /* 1 = replay update!/solve! as CUDA graphs (default 1) */
    int32_t analyze_only;  /* 1 = host symbolic analysis only, no device is touched (tests, no GPU) */
    int32_t rank, nranks;  /* multi-GPU subtree sharding: this process's rank and the world size (default 0, 1) */
    int32_t dense_col_threshold; /* K1: columns with more non-zeros are handled by a low-rank Schur correction;
                                    0 = auto (max(32, 5% of m)), < 0 = off */
    int32_t dense_solve_ncol;    /* non-small supernodes with at least this many columns use the dense-solve path
                                    (repacked unit-block-diagonal tiles, flag-in-data hand-over); 0 = default (384) */
    int32_t ozaki_ncol;          /* K1, single GPU: supernodes with at least this many columns run their far Schur updates on
                                    the tcgen05 int8 tensor-core path (exact digit-plane products, FP64-grade result);
                                    0 = default (1024), < 0 = off (FP64 DMMA path everywhere) */
    int32_t refine_steps;        /* iterative-refinement steps inside solve! on the factored KKT system (the reference
                                    has only TODOs, spd.jl:68 / sqd.jl:72): 0 = off (default, the reference's single solve), <= 8;
                                    single GPU.  The dense-column path always refines (its Woodbury solve is the preconditioner). */
    int32_t reserved[3];
} tlpb200_options;

#define TLPB200_NCLASS 24
typedef struct tlpb200_stats {
    int64_t m, n, nnzA;
    int64_t order;          /* N: m for K1, n+m for K2 */
    int64_t nnzL;           /* sum of column counts (structural, before amalgamation) */
    int64_t nnzL_stored;    /* doubles of panel storage on the device */
    double flops;           /* sum of squared column counts = factorisation flops */
    int64_t nsuper, npieces, nlevels;
    int64_t max_ncol, max_nrow;
    int64_t nproducts;      /* K1: scalar products a_ij*a_kj streamed by the assemble kernel */
    int64_t nentries;       /* K1: structural non-zeros of lower(A*A') */
    int64_t launches_update, launches_solve; /* kernels of one update! / one solve! */
    double ms_assemble, ms_factor, ms_solve;  /* CUDA-event times of the last update!/solve! (profiling mode) */
    int64_t bad_pivot;      /* permuted column of the first bad pivot of the last update!, -1 if none */
    int64_t n_update, n_solve;
    int64_t bytes_device;
    /* algorithmic flops of one update! (2 per multiply-add, lower triangle only): diagonal-block
     * factor + trsm of the column pieces / tile-update kernel (supernode SYRK-GEMM + scatter) */
    double flops_update_inner, flops_update_ext;
    /* profiling mode: CUDA-event time and launch count per kernel class of the last update!/solve!
     * 0 assemble  1 small_factor  2 diag_factor  3 trsm  4 update  5 rhs+recover
     * 6 fwd_small 7 fwd_large 8 -  9 bwd_large 10 invert_diag 11 bwd_small */
    double ms_class[TLPB200_NCLASS];   /* ... 12 dense_cols 13 pack_big 14 fwd_big 15 bwd_big 16 oz_slice 17 oz_update 18 collectives 19 refinement */
    int64_t n_class[TLPB200_NCLASS];
    double flops_update_oz; /* algorithmic flops of one update! done by the tcgen05 int8 tasks (not part of flops_update_ext) */
    int64_t oz_tasks;       /* tcgen05 tasks per update! */
    int64_t oz_bytes;       /* device bytes of the digit planes */
} tlpb200_stats;

void tlpb200_default_options(tlpb200_options* opt);
/* out[0] = sizeof(tlpb200_options), out[1] = sizeof(tlpb200_stats), out[2] = TLPB200_NCLASS: lets a binding (the ctypes mirror in
 * tulip.jl_b200/_lib.py, the Julia structs in tulip.jl_b200/julia/TlpB200.jl) check its struct layouts against the library. */
void tlpb200_abi_sizes(int32_t* out);

/* KKT.setup(A, system, backend)  (KKT.jl:59; Cholmod/spd.jl:5-20, sqd.jl:5-22):
 * symbolic analysis on the host, plan + matrix upload to the device. */
int tlpb200_create(tlpb200_solver** out, int64_t m, int64_t n, const int64_t* colptr, const int64_t* rowval,
                   const double* nzval, int index_base, int system, const tlpb200_options* opt);

/* KKT.update!(kkt, theta_inv, regP, regD)  (KKT.jl:83; spd.jl:22-50, sqd.jl:24-55):
 * assemble + numeric factorisation on the device.  *bad_pivot (may be NULL) receives the
 * permuted column index of the first bad pivot or -1. */
int tlpb200_update(tlpb200_solver* s, const double* theta_inv, const double* regP, const double* regD,
                   int64_t* bad_pivot);
int tlpb200_update_dev(tlpb200_solver* s, const double* d_theta_inv, const double* d_regP, const double* d_regD);
/* result of the last update_dev: synchronises, returns TLPB200_OK / TLPB200_NOT_POSDEF */
int tlpb200_update_status(tlpb200_solver* s, int64_t* bad_pivot);

/* KKT.solve!(dx, dy, kkt, xi_p, xi_d)  (KKT.jl:100; spd.jl:52-70, sqd.jl:57-74).
 * nrhs right-hand sides stored with leading dimensions ldx (dx, xi_d) and ldy (dy, xi_p). */
int tlpb200_solve(tlpb200_solver* s, double* dx, double* dy, const double* xi_p, const double* xi_d,
                  int32_t nrhs, int64_t ldx, int64_t ldy);
int tlpb200_solve_dev(tlpb200_solver* s, double* d_dx, double* d_dy, const double* d_xi_p, const double* d_xi_d,
                      int32_t nrhs, int64_t ldx, int64_t ldy);

/* result of the solve_dev calls enqueued so far: synchronises; TLPB200_INTERNAL if a sweep kernel's bounded hand-over wait ran
 * out (results invalid).  tlpb200_solve itself performs this check before returning. */
int tlpb200_solve_status(tlpb200_solver* s);

/* test hook: raise the device-side "sweep hand-over timed out" flag as a kernel would (the next solve must report it) */
int tlpb200_debug_raise_timeout(tlpb200_solver* s);

/* stream the _dev calls enqueue on (a cudaStream_t); default: a stream owned by the solver */
int tlpb200_set_stream(tlpb200_solver* s, void* cuda_stream);
int tlpb200_synchronize(tlpb200_solver* s);
/* 1 = bracket assemble / factor / solve with CUDA events and fill ms_* in the stats (adds syncs) */
int tlpb200_set_profiling(tlpb200_solver* s, int on);

int tlpb200_stats_get(const tlpb200_solver* s, tlpb200_stats* out);

/* Integer results of the symbolic analysis, for the bit-exact tests.  Each pointer may be NULL.
 * perm[N] (perm[new]=old, 0-based), parent[N] (etree of the permuted matrix, -1 root),
 * colcount[N], sn_first[nsuper+1]. */
int tlpb200_get_symbolic(const tlpb200_solver* s, int32_t* perm, int32_t* parent, int32_t* colcount,
                         int32_t* sn_first);
/* supernodal row structure: rowptr[nsuper+1], rows[rowptr[nsuper]] (pass NULL to query sizes via stats) */
int tlpb200_get_structure(const tlpb200_solver* s, int64_t* rowptr, int32_t* rows);

/* Debug/parity: copy the assembled-and-factored panels (or, with factor=0 in the last
 * tlpb200_debug_assemble call, the assembled matrix) back to the host: Lx[nnzL_stored]. */
int tlpb200_debug_assemble(tlpb200_solver* s, const double* theta_inv, const double* regP, const double* regD);
int tlpb200_debug_get_lx(tlpb200_solver* s, double* lx, int64_t* xptr /* nsuper+1, may be NULL */);
/* dense-solve plan of the big supernodes (host data; works on analyze_only handles): counts[6] = #pack tasks,
   #forward tasks, #backward tasks, #forward tiles, #backward tiles, #exchange slots; the arrays receive the raw
   32-byte pack records {i32 sn, r0, nr, j; i64 fdst, bdst} and 40-byte task records
   {i32 sn, kind, blk, r0, nr, ntile, nbelow, xq0; i64 tile0}.  Any pointer may be NULL. */
/* globaltimer (ns) of every block publish of the last dense sweeps: out[0..nblk) forward, out[nblk..2 nblk) backward;
   recorded only when TLPB200_CHAIN_TIMES was set in the environment at tlpb200_create */
int tlpb200_debug_chain_times(tlpb200_solver* s, uint64_t* out, int64_t* nblk);
/* factorisation timeline of the last update! (needs TLPB200_TRACE_FACTOR in the environment at tlpb200_create):
   out[level][cls][2] = globaltimer ns of the first CTA start / last CTA end of cls = {diagonal blocks, trsm, urgent update
   tiles, lazy update tiles}; zeros where a class is absent */
int tlpb200_debug_factor_trace(tlpb200_solver* s, uint64_t* out, int64_t* nlevels);
/* Stand-alone check of the tcgen05 int8 (Ozaki) Schur-update path on a dense matrix: C (R x R, column-major, lower part)
 * -= P P' with P R x K column-major; `ksplit` = K columns per task (multiple of 32, 0 = all).  ms[0] = digit-plane
 * kernels, ms[1] = one update pass (mean of `reps`), ms[2] = tasks; *err != 0 = pipeline time-out. */
int tlpb200_debug_ozaki(const double* P, int64_t R, int64_t K, double* C, int32_t ksplit, int32_t reps, float* ms, int32_t* err);
/* Update-task plan as int32 records (host data, also on analyze_only handles): upd / upd128 = FP64 tile tasks
 * {piece, i0, ni, k0, nk, tgt, diag, pad}, oz = tcgen05 tasks {view, rbA, rbB, half, k0, k1, pad, pad}, pieces =
 * {sn, c0, c1, level}, views = {sn, nrb, ncb, base_level}, panel = trsm row tiles {piece, r0, nr, pad}, levels = the
 * per-level ranges (LevelPlan, counts[7] int32 each), small_list / level_pieces = the one-CTA supernodes / pieces grouped by
 * level; counts[0..6], counts[8], counts[9] = their lengths (counts has 10 entries).  NULL arrays are skipped. */
int tlpb200_debug_update_plan(const tlpb200_solver* s, int64_t* counts, int32_t* upd, int32_t* upd128, int32_t* oz, int32_t* pieces,
                              int32_t* views, int32_t* panel, int32_t* levels, int32_t* small_list, int32_t* level_pieces);
int tlpb200_debug_big_plan(const tlpb200_solver* s, int64_t* counts, void* pack, void* fwd, void* bwd);
/* Launch sequences of the triangular sweeps (host data, also on analyze_only handles): the block-solve items of consecutive
 * levels share one launch and synchronise through per-supernode counters.  counts[4] = #fwd ops, #bwd ops, #fwd items,
 * #bwd items; ops = int32 records {kind (0 small, 1 merged block-solve, 2 dense-solve, 3 below), begin, end, level};
 * fwd_need / fwd_parent / bwd_wait / bwd_nitems / sn_parent / bwd_nbelow = per-supernode int32 arrays; items = 24-byte
 * records {sn, blk, kind, r0, nr, pad} in the order the sweeps walk them (backward kind 2 = rows below the columns of a tall
 * block-solve supernode, accumulated before its blocks).  NULL pointers are skipped. */
int tlpb200_debug_solve_ops(const tlpb200_solver* s, int64_t* counts, int32_t* fwd_ops, int32_t* bwd_ops, int32_t* fwd_need,
                            int32_t* fwd_parent, int32_t* bwd_wait, int32_t* bwd_nitems, void* fwd_items, void* bwd_seq,
                            int32_t* sn_parent, int32_t* bwd_nbelow);

/* sharded solver (nranks > 1): dependency targets of the merged-level sweeps inside phase 0 (own subtrees) / 1 (replicated top
 * part) -- per-supernode int32 arrays like fwd_need / fwd_parent / bwd_wait above, restricted to supernodes of the same phase */
int tlpb200_debug_phase_deps(const tlpb200_solver* s, int32_t phase, int32_t* fwd_need, int32_t* fwd_parent, int32_t* bwd_wait);

/* ---- multi-GPU, one process per GPU (no counterpart in the reference, NEWS.md:31) -------------
 * Created with opt.nranks > 1 every rank analyses the same matrix, owns the elimination-tree subtrees
 * assigned to it (owner[s] == rank) and replicates the top part (owner[s] == -1).  The caller performs the
 * collectives on the exposed device buffers between the phases (tulip.jl_b200/parallel.py uses NCCL):
 *   update_begin -> all_reduce(sum, top_panels) -> update_end (+ all_reduce(max) of the return codes)
 *   solve_begin  -> all_reduce(sum, work_vector) -> solve_mid -> all_reduce(sum, work_vector) -> solve_end */
int tlpb200_dist_info(const tlpb200_solver* s, int32_t* owner /* nsuper */, int64_t* top_offset, int64_t* top_count);
int tlpb200_update_begin(tlpb200_solver* s, const double* theta_inv, const double* regP, const double* regD);
int tlpb200_top_panels(tlpb200_solver* s, void** dptr, int64_t* count);
int tlpb200_update_end(tlpb200_solver* s, int64_t* bad_pivot);
int tlpb200_solve_begin(tlpb200_solver* s, const double* xi_p, const double* xi_d);
int tlpb200_work_vector(tlpb200_solver* s, void** dptr, int64_t* count);
int tlpb200_solve_mid(tlpb200_solver* s);
int tlpb200_solve_end(tlpb200_solver* s, double* dx, double* dy);

/* In-library collectives (preferred over the phase API above): after tlpb200_comm_init the ordinary entry points
 * tlpb200_update / tlpb200_update_dev / tlpb200_solve / tlpb200_solve_dev run the whole sharded sequence -- own subtrees,
 * NCCL all-reduce of the top panels / of the separator entries of the work vector, replicated top part -- as ONE
 * stream-ordered (CUDA-graph) sequence on the solver's stream, without host synchronisation between the phases; return
 * codes are identical on every rank (the status words are all-reduced).  One rank calls tlpb200_comm_unique_id (128 bytes,
 * an ncclUniqueId), the caller's own plumbing broadcasts it (torch.distributed in tulip.jl_b200/parallel.py), then EVERY
 * rank calls tlpb200_comm_init (collective).  NCCL is resolved at run time (dlopen of libnccl.so.2). */
int tlpb200_comm_unique_id(void* out128);
int tlpb200_comm_init(tlpb200_solver* s, const void* id128);
/* mean time [ms] of each collective of a sharded update!/solve!, timed alone (`reps` back-to-back calls): [0] top panels,
 * [1] separator entries, [2] solution vector, [3] status words; bytes[] = their payloads.  Collective call. */
int tlpb200_comm_profile(tlpb200_solver* s, int32_t reps, float* ms /* 4 */, int64_t* bytes /* 4 */);

/* ---- device-resident IPM iteration ------------------------------------------------------------------
 * The caller of the KKT path -- Tulip's homogeneous self-dual IPM, /root/reference/src/IPM/HSD/HSD.jl:203-350 (main loop,
 * residuals :77-128, status tests :136-196) and src/IPM/HSD/step.jl:10-401 (compute_step!) -- with every vector resident in
 * HBM: theta / regularisation schedule / rhs builds / recoveries / step lengths / corrector targets are fused elementwise
 * kernels, update! / solve! are the same device sequences as above, the host reads back one block of scalars per decision.
 * The LP is the standard form the KKT boundary sees (A x = b, l <= x <= u; ipmdata.jl:6-12) with the A given to
 * tlpb200_create.  With nranks > 1 (after tlpb200_comm_init) every rank runs the same loop on the sharded factorisation. */
enum {
    TLPB200_TRM_UNKNOWN = 0,           /* keep iterating                       (status.jl: Trm_Unknown) */
    TLPB200_TRM_OPTIMAL = 1,           /* HSD.jl:161-166 */
    TLPB200_TRM_PRIMAL_INFEASIBLE = 2, /* HSD.jl:181-192 */
    TLPB200_TRM_DUAL_INFEASIBLE = 3,   /* HSD.jl:170-179 */
    TLPB200_TRM_ITERATION_LIMIT = 4,   /* HSD.jl:303-305 */
    TLPB200_TRM_TIME_LIMIT = 5,        /* HSD.jl:306-308 */
    TLPB200_TRM_NUMERICAL_PROBLEM = 6, /* HSD.jl:321-326 (factorisation could not be saved, step.jl:51) */
    TLPB200_TRM_MEMORY_LIMIT = 7       /* HSD.jl:327 */
};
typedef struct tlpb200_hsd_options {   /* src/IPM/options.jl:1-25 */
    int32_t iterations_limit;          /* 100 */
    int32_t correction_limit;          /* 3 */
    double time_limit;                 /* seconds, inf */
    double tol_pfeas, tol_dfeas, tol_rgap, tol_ifeas;   /* sqrt(eps) */
    double step_damp;                  /* 0.9995 */
    double gamma_min;                  /* 0.1 */
    double centrality_outlier;         /* 0.1 */
    double preg_min, dreg_min;         /* sqrt(eps) */
} tlpb200_hsd_options;
typedef struct tlpb200_hsd_info {
    int32_t status, niter;
    double pobj, dobj;                 /* primal / dual objective of the current iterate (HSD.jl:126-127) */
    double rp_nrm, rl_nrm, ru_nrm, rd_nrm, rg_nrm;   /* residual infinity norms (HSD.jl:119-123) */
    double mu, tau, kappa;
    int64_t n_update, n_solve;         /* KKT calls so far */
    double ms_update, ms_solve;        /* CUDA-event time of those calls: the reference's "Factorization" / "KKT" timer sections */
    double seconds_total;              /* wall time of tlpb200_hsd_optimize */
} tlpb200_hsd_info;
void tlpb200_hsd_default_options(tlpb200_hsd_options* opt);
int tlpb200_hsd_create(tlpb200_solver* s, const double* b /* m */, const double* c /* n */, const double* l /* n, -inf allowed */,
                       const double* u /* n, +inf allowed */, double c0);
int tlpb200_hsd_reset(tlpb200_solver* s);   /* next iterate call starts again from the start point (HSD.jl:238-247) */
/* one pass of the main-loop body: residuals + status tests, then compute_step! unless the run is over */
int tlpb200_hsd_iterate(tlpb200_solver* s, const tlpb200_hsd_options* opt, tlpb200_hsd_info* info);
/* ipm_optimize!: from the start point until a termination status */
int tlpb200_hsd_optimize(tlpb200_solver* s, const tlpb200_hsd_options* opt, tlpb200_hsd_info* info);
/* copy the iterate to the host (any pointer may be NULL); tau_kappa[2] */
int tlpb200_hsd_get_point(tlpb200_solver* s, double* x, double* xl, double* xu, double* y, double* zl, double* zu, double* tau_kappa);
/* rows of 8 doubles {iter, pobj, dobj, pfeas, dfeas, gfeas, mu, tau}, one per pass (the reference's log line, HSD.jl:266-287) */
int tlpb200_hsd_get_log(tlpb200_solver* s, double* out, int64_t* rows);

/* K1 dense-column path: number and (0-based) indices of the columns kept out of the sparse factor */
int tlpb200_get_dense_cols(const tlpb200_solver* s, int32_t* count, int64_t* cols /* may be NULL */);

const char* tlpb200_last_error(const tlpb200_solver* s);
const char* tlpb200_backend_name(void);            /* KKT.backend(kkt)       (KKT.jl:114) */
const char* tlpb200_linear_system(const tlpb200_solver* s); /* KKT.linear_system(kkt) (KKT.jl:121) */
void tlpb200_destroy(tlpb200_solver* s);

#ifdef __cplusplus
}
#endif
#endif /* TLPB200_H */
