// Host orchestration + C ABI of the B200 KKT backend (see include/tlpb200.h).
//
// Mirrors the life-cycle of the reference's CholmodSolver (/root/reference/src/KKT/Cholmod/
// cholmod.jl:46-60): setup once (spd.jl:5-20 / sqd.jl:5-22), then per IPM iteration one update!
// (spd.jl:22-50 / sqd.jl:24-55) and 3-6 solve! calls (spd.jl:52-70 / sqd.jl:57-74).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only: the library is resolved at run time (dlopen), see NcclApi below

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <climits>
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/tlpb200.h"
#include "kernels.cuh"
#include "plan.hpp"
#include "symbolic.hpp"
#include "solver_internal.hpp"

using namespace tlp;

namespace {

struct CudaFail {
    cudaError_t e;
    const char* what;
};

#define CK(call)                                              \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) throw CudaFail{_e, #call};     \
    } while (0)

template <typename T>
T* dalloc(tlpb200_solver* s, size_t count) {
    void* p = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CK(cudaMalloc(&p, bytes));
    s->allocs.push_back(p);
    s->bytes_device += bytes;
    return (T*)p;
}

template <typename T>
const T* upload(tlpb200_solver* s, const std::vector<T>& v) {
    T* p = dalloc<T>(s, v.size());
    if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}

int fail(tlpb200_solver* s, int code, const std::string& msg) {
    if (s) s->err = msg;
    return code;
}

// RAII event bracket around one launch (only active in profiling mode).  Brackets nest (the dense-column phase wraps whole
// sweeps that bracket their own launches): only the outermost one records, so that the event pool stays consistent --
// the inner ones used to advance pool_used under the outer bracket, whose destructor then wrote past the pool (segfault of
// the first profiled config-5 run, round 2).
struct Scope {
    tlpb200_solver* s;
    bool on;
    Scope(tlpb200_solver* s_, int cls) : s(s_), on(s_->profiling && s_->scope_depth == 0) {
        if (s->profiling) s->scope_depth++;
        if (!on) return;
        if (s->pool_used + 2 > s->pool.size()) {
            for (int k = 0; k < 2; ++k) { cudaEvent_t e; cudaEventCreate(&e); s->pool.push_back(e); }
        }
        s->pool_cls.resize(s->pool.size() / 2);
        s->pool_cls[s->pool_used / 2] = cls;
        cudaEventRecord(s->pool[s->pool_used], s->stream);
    }
    ~Scope() {
        if (s->profiling && s->scope_depth > 0) s->scope_depth--;
        if (!on) return;
        cudaEventRecord(s->pool[s->pool_used + 1], s->stream);
        s->pool_used += 2;
    }
};

void collect_profile(tlpb200_solver* s, bool reset_update_classes) {
    // called after a stream sync
    static const bool is_update_class[TLPB200_NCLASS] = {1, 1, 1, 1, 1, 0, 0, 0, 1, 0, 1, 0, 0, 1, 0, 0, 1, 1};   // class 12 (dense cols) is left cumulative
    for (int c = 0; c < TLPB200_NCLASS; ++c)
        if (is_update_class[c] == reset_update_classes) { s->ms_class[c] = 0; s->n_class[c] = 0; }
    for (size_t i = 0; i + 1 < s->pool_used; i += 2) {
        float ms = 0;
        cudaEventElapsedTime(&ms, s->pool[i], s->pool[i + 1]);
        const int c = s->pool_cls[i / 2];
        s->ms_class[c] += ms;
        s->n_class[c] += 1;
    }
    s->pool_used = 0;
}

int cuda_fail(tlpb200_solver* s, const CudaFail& f) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s", (int)f.e, cudaGetErrorString(f.e), f.what);
    cudaGetLastError();
    return fail(s, f.e == cudaErrorMemoryAllocation ? TLPB200_OOM : TLPB200_CUDA, buf);
}

// Host <-> device through the pinned staging area, in chunks: the CPU copy of chunk k + 1 overlaps the DMA of chunk k (host ->
// device), the DMA of chunk k + 1 overlaps the CPU copy of chunk k (device -> host).  On the mid-size configs the single
// memcpy + single DMA of round 1 was a quarter of the host-API time of a solve! (2.4 MB each way at ~10 GB/s + ~25 GB/s).
constexpr size_t STAGE_CHUNK = 32 * 1024;   // doubles (256 KiB)

void stage_h2d(tlpb200_solver* s, double* dev, double* pin, const double* host, size_t count) {
    for (size_t o = 0; o < count; o += STAGE_CHUNK) {
        const size_t c = std::min(STAGE_CHUNK, count - o);
        std::memcpy(pin + o, host + o, c * 8);
        CK(cudaMemcpyAsync(dev + o, pin + o, c * 8, cudaMemcpyHostToDevice, s->stream));
    }
}

struct StageOut {
    double* host;
    const double* pin;
    size_t count;
    size_t ev0;     // first event of this transfer inside s->ev_stage
};

// enqueue the chunked device -> host copies (one event per chunk); finish_d2h then drains them chunk by chunk
StageOut stage_d2h_begin(tlpb200_solver* s, double* host, double* pin, const double* dev, size_t count, size_t ev0) {
    size_t k = ev0;
    for (size_t o = 0; o < count; o += STAGE_CHUNK, ++k) {
        const size_t c = std::min(STAGE_CHUNK, count - o);
        CK(cudaMemcpyAsync(pin + o, dev + o, c * 8, cudaMemcpyDeviceToHost, s->stream));
        while (s->ev_stage.size() <= k) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s->ev_stage.push_back(e);
        }
        CK(cudaEventRecord(s->ev_stage[k], s->stream));
    }
    return StageOut{host, pin, count, ev0};
}
size_t stage_chunks(size_t count) { return (count + STAGE_CHUNK - 1) / STAGE_CHUNK; }
void stage_d2h_finish(tlpb200_solver* s, const StageOut& t) {
    size_t k = t.ev0;
    for (size_t o = 0; o < t.count; o += STAGE_CHUNK, ++k) {
        const size_t c = std::min(STAGE_CHUNK, t.count - o);
        CK(cudaEventSynchronize(s->ev_stage[k]));
        std::memcpy(t.host + o, t.pin + o, c * 8);
    }
}

// NCCL is taken from the process at run time: torch has already loaded its bundled libnccl.so.2 when the Python host
// mirror drives the library (RTLD_NOLOAD finds that copy); a stand-alone caller gets the system library.  Nothing is
// linked, so the single-GPU product and the CPU-only tests do not depend on NCCL being installed.
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string err;
};

NcclApi& nccl_api() {
    static NcclApi api;
    if (api.h || !api.err.empty()) return api;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return api; }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))dlsym(h, "ncclGetVersion");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) { api.err = "libnccl.so.2 lacks the expected symbols"; return api; }
    api.h = h;
    return api;
}

struct NcclFail {
    ncclResult_t e;
    const char* what;
};
#define NK(call)                                               \
    do {                                                       \
        ncclResult_t _e = (call);                              \
        if (_e != ncclSuccess) throw NcclFail{_e, #call};      \
    } while (0)

int nccl_fail(tlpb200_solver* s, const NcclFail& f) {
    char buf[512];
    const NcclApi& api = nccl_api();
    snprintf(buf, sizeof buf, "NCCL error %d (%s) in %s", (int)f.e, api.GetErrorString ? api.GetErrorString(f.e) : "?", f.what);
    return fail(s, TLPB200_NCCL, buf);
}

inline bool sharded(const tlpb200_solver* s) { return s->nranks > 1; }

// does [begin, end) of item list `wl` contain anything the current phase processes?  (always, unless a sharded phase is active)
inline bool has_work(const tlpb200_solver* s, int wl, int64_t begin, int64_t end) {
    if (end <= begin) return false;
    const int ph = (s->cur == &s->ctxA) ? 0 : ((s->cur == &s->ctxB) ? 1 : -1);
    if (ph < 0) return true;
    const std::vector<int32_t>& pre = s->work_prefix[ph][wl];
    if (pre.empty()) return true;
    return pre[(size_t)end] - pre[(size_t)begin] > 0;
}

// Merged-level sweeps inside a sharded phase (phase 0 = own subtrees, 1 = replicated top part): a supernode only waits for /
// notifies supernodes that run in the SAME phase.  A subtree root's parent belongs to the top part: in the forward sweep it
// runs in a later phase (no notification: its counter must only see top children), in the backward sweep it is complete
// before the phase starts (no wait).
void build_phase_deps(tlpb200_solver* s) {
    const Plan& P = s->plan;
    const int32_t ns = s->sym.nsuper;
    std::vector<int32_t> fitems(ns, 0);
    for (const SolveItem& it : P.fwd_items) fitems[it.sn]++;
    for (int ph = 0; ph < 2; ++ph) {
        auto live = [&](int32_t sn) { return ph == 0 ? (s->owner[sn] == s->rank) : (s->owner[sn] == -1); };
        s->phase_need[ph].assign(ns, 0);
        s->phase_fpar[ph].assign(ns, -1);
        s->phase_bwait[ph].assign(ns, -1);
        for (int32_t sn = 0; sn < ns; ++sn) {
            if (!live(sn)) continue;
            const int32_t fp = P.fwd_parent[sn];
            if (fp >= 0 && live(fp)) { s->phase_fpar[ph][sn] = fp; s->phase_need[ph][fp] += fitems[sn]; }
            const int32_t bp = P.bwd_wait[sn];
            if (bp >= 0 && live(bp)) s->phase_bwait[ph][sn] = bp;
        }
    }
}

void build_work_prefix(tlpb200_solver* s) {
    const Plan& P = s->plan;
    for (int ph = 0; ph < 2; ++ph) {
        auto live = [&](int32_t sn) { return ph == 0 ? (s->owner[sn] == s->rank) : (s->owner[sn] == -1); };
        auto fill = [&](int wl, size_t n, auto sn_of) {
            std::vector<int32_t>& pre = s->work_prefix[ph][wl];
            pre.assign(n + 1, 0);
            for (size_t i = 0; i < n; ++i) pre[i + 1] = pre[i] + (live(sn_of(i)) ? 1 : 0);
        };
        fill(tlpb200_solver::WL_SMALL, P.small_list.size(), [&](size_t i) { return P.small_list[i]; });
        fill(tlpb200_solver::WL_PIECE, P.level_pieces.size(), [&](size_t i) { return P.pieces[P.level_pieces[i]].sn; });
        fill(tlpb200_solver::WL_PANEL, P.panel.size(), [&](size_t i) { return P.pieces[P.panel[i].piece].sn; });
        fill(tlpb200_solver::WL_EXT, P.upd.size(), [&](size_t i) { return P.pieces[P.upd[i].piece].sn; });
        fill(tlpb200_solver::WL_LAZY, P.upd128.size(), [&](size_t i) { return P.pieces[P.upd128[i].piece].sn; });
        fill(tlpb200_solver::WL_FWD, P.fwd_items.size(), [&](size_t i) { return P.fwd_items[i].sn; });
        fill(tlpb200_solver::WL_BWD, P.bwd_items.size(), [&](size_t i) { return P.bwd_items[i].sn; });
        fill(tlpb200_solver::WL_FBIG, P.fwd_big.size(), [&](size_t i) { return P.fwd_big[i].sn; });
        fill(tlpb200_solver::WL_BBIG, P.bwd_big.size(), [&](size_t i) { return P.bwd_big[i].sn; });
        fill(tlpb200_solver::WL_BELOW, P.bwd_below.size(), [&](size_t i) { return P.bwd_below[i].sn; });
        fill(tlpb200_solver::WL_INV, P.inv_order.size(), [&](size_t i) { return P.dblk_sn[P.inv_order[i]]; });
        fill(tlpb200_solver::WL_PACK, P.big_pack.size(), [&](size_t i) { return P.big_pack[i].sn; });
        fill(tlpb200_solver::WL_BSEQ, P.bwd_seq.size(), [&](size_t i) { return P.bwd_seq[i].sn; });
    }
}

// ---- the numeric phases, enqueued on s->stream; `count` accumulates kernel launches -------------
void enqueue_assemble(tlpb200_solver* s, int64_t& count, bool own_panels_only = false) {
    cudaStream_t st = s->stream;
    Scope sc(s, 0);
    if (own_panels_only && s->nranks > 1 && s->rank_begin.size() == (size_t)s->nranks + 1) {
        // sharded: only this rank's subtrees and the top part are ever read here (panels are grouped by owner)
        const int64_t b = s->rank_begin[s->rank], e = s->rank_begin[s->rank + 1];
        if (e > b) CK(cudaMemsetAsync(s->ctx.Lx + b, 0, (size_t)(e - b) * sizeof(double), st));
        if (s->sym.lx_size > s->top_begin) CK(cudaMemsetAsync(s->ctx.Lx + s->top_begin, 0, (size_t)(s->sym.lx_size - s->top_begin) * sizeof(double), st));
    } else {
        CK(cudaMemsetAsync(s->ctx.Lx, 0, (size_t)s->sym.lx_size * sizeof(double), st));
    }
    CK(cudaMemsetAsync(s->ctx.info, 0x7f, sizeof(int32_t), st));
    CK(cudaMemsetAsync(s->lazy_ctr, 0, (2 * s->plan.levels.size() + 2) * sizeof(int32_t), st));
    if (s->ctx.trace_min) {
        CK(cudaMemsetAsync(s->ctx.trace_min, 0xff, 4 * s->plan.levels.size() * sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(s->ctx.trace_max, 0, 4 * s->plan.levels.size() * sizeof(unsigned long long), st));
    }
    if (s->system == TLPB200_K1) {
        launch_compute_d(s->d_theta, s->d_regP, s->d_d, s->n, st);
        launch_assemble_k1(s->ctx, s->mat, s->d_d, s->d_regD, st);
        count += 3;
        if (s->oz_on) CK(cudaMemsetAsync(s->oz_ctr, 0, (2 * s->plan.levels.size() + 2) * sizeof(int32_t), st));
    } else {
        launch_assemble_k2(s->ctx, s->mat, s->d_theta, s->d_regP, s->d_regD, st);
        count += 2;
    }
}

// Numeric factorisation.  Critical chain per level L on the main stream: one-CTA supernodes, diagonal
// blocks, trsm, then the "urgent" update tiles (those that feed level L+1).  The bulk of the updates
// ("lazy" tiles, needed from level L+2 on) runs on a lower-priority side stream underneath the chain.
// Concurrent updates into the same ancestor entries are resolved by RED.ADD.F64, which is also the
// faster epilogue on its own: 12.7 ms vs 17.7 ms per cfg2 factorisation for read-modify-write.
// Profiling mode and TLPB200_NO_OVERLAP=1 use the single-stream order.
void enqueue_fwd(tlpb200_solver* s, int64_t& count);
void enqueue_bwd(tlpb200_solver* s, int64_t& count);
void reset_sweep_state(tlpb200_solver* s, cudaStream_t st);

void enqueue_factor(tlpb200_solver* s, int64_t& count) {
    cudaStream_t st = s->stream;
    static const bool no_overlap_env = getenv("TLPB200_NO_OVERLAP") != nullptr;
    const bool overlap = !s->profiling && !no_overlap_env && s->side_stream != nullptr;
    const auto& L = s->plan.levels;
    const size_t nlev = L.size();
    if (overlap && s->ev_f.size() < nlev) {
        while (s->ev_f.size() < nlev) {
            cudaEvent_t a, b;
            CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            s->ev_f.push_back(a);
            s->ev_lazy.push_back(b);
            cudaEvent_t c3[3];
            for (auto& e : c3) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s->ev_d.push_back(c3[0]);
            s->ev_tr.push_back(c3[1]);
            s->ev_ur.push_back(c3[2]);
        }
    }
    const bool oz = s->oz_on && s->cur == &s->ctx;
    if (oz && overlap && s->ev_oz.size() < nlev) {
        while (s->ev_oz.size() < nlev) {
            cudaEvent_t e3[3];
            for (auto& e : e3) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s->ev_ozt.push_back(e3[0]);
            s->ev_ozs.push_back(e3[1]);
            s->ev_oz.push_back(e3[2]);
        }
    }
    std::vector<char> has_oz(nlev, 0);
    long oz_waited = -1;     // every tcgen05 batch of a level <= oz_waited has been joined by the main stream
    // columns that are final early are inverted / repacked for the solves underneath the tail of the factorisation
    const long pack_level = (overlap && s->pack_split > 0.0 && s->pack_split < 1.0 && !s->plan.big_pack.empty()) ? (long)(s->pack_split * (double)nlev) : -1;
    int32_t inv_done = 0, pack_done = 0;
    bool early_pack = false;
    std::vector<char> has_lazy(nlev, 0);
    long waited = -1;        // every lazy batch of a level <= waited has been joined by the main stream
    int nbatch = 0;
    for (size_t l = 0; l < nlev; ++l) {
        const LevelPlan& lp = L[l];
        if (overlap && l >= 2) {     // level l needs every lazy batch of levels <= l-2 (RED updates commute: no order among batches)
            for (long q = waited + 1; q <= (long)l - 2; ++q)
                if (has_lazy[q]) CK(cudaStreamWaitEvent(st, s->ev_lazy[q], 0));
            waited = (long)l - 2;
        }
        if (oz && overlap && l >= 3) {   // column block c of an oz supernode was completed by the batch of level c - 3
            for (long q = oz_waited + 1; q <= (long)l - 3; ++q)
                if (has_oz[q]) CK(cudaStreamWaitEvent(st, s->ev_oz[q], 0));
            oz_waited = (long)l - 3;
        }
        if (oz)      // row scales of the oz supernodes that start at this level (before their first piece is factored)
            for (const OzViewPlan& v : s->plan.oz_views)
                if (v.base_level == (int32_t)l) {
                    launch_oz_rowexp(s->ctx.Lx, s->ctx.diagpos, s->ctx.sn_rows + s->sym.sn_rowptr[v.sn],
                                     (int32_t)(s->sym.sn_rowptr[v.sn + 1] - s->sym.sn_rowptr[v.sn]), s->oz_E + v.row0, s->oz_scl + v.row0, st);
                    count++;
                }
        if (has_work(s, tlpb200_solver::WL_SMALL, lp.small_begin, lp.small_end)) {
            Scope sc(s, 1);
            launch_small_factor((*s->cur), lp.small_begin, lp.small_end, s->small_smem, st);
            count++;
        }
        if (has_work(s, tlpb200_solver::WL_PIECE, lp.piece_begin, lp.piece_end)) { Scope sc(s, 2); launch_diag_factor((*s->cur), lp.piece_begin, lp.piece_end, st); count++; }
        const bool split = overlap && s->split_chain;
        bool ev_f_recorded = false;
        if (!split) {
            if (has_work(s, tlpb200_solver::WL_PANEL, lp.panel_begin, lp.panel_end)) { Scope sc(s, 3); launch_trsm((*s->cur), lp.panel_begin, lp.panel_end, st); count++; }
        } else {
            // Chain (main stream): diagonal blocks -> critical trsm row tiles -> critical update tiles -> next level's
            // diagonal blocks.  The rest of the trsm and of the urgent tiles runs on the aux stream underneath the next
            // diagonal-block factorisation; the next level's trsm joins it.
            CK(cudaEventRecord(s->ev_d[l], st));
            CK(cudaStreamWaitEvent(s->aux_stream, s->ev_d[l], 0));
            if (has_work(s, tlpb200_solver::WL_PANEL, lp.panel_crit_end, lp.panel_end)) { launch_trsm((*s->cur), lp.panel_crit_end, lp.panel_end, s->aux_stream); count++; }
            CK(cudaEventRecord(s->ev_tr[l], s->aux_stream));
            if (l > 0) CK(cudaStreamWaitEvent(st, s->ev_ur[l - 1], 0));   // rest of the previous level's urgent tiles
            if (has_work(s, tlpb200_solver::WL_PANEL, lp.panel_begin, lp.panel_crit_end)) { launch_trsm((*s->cur), lp.panel_begin, lp.panel_crit_end, st); count++; }
        }
        // the persistent bulk kernel leaves `chain_sms` SMs to the critical-chain kernels
        if (has_work(s, tlpb200_solver::WL_LAZY, lp.lazy_begin, lp.lazy_end)) {
            if (!overlap) {
                Scope sc(s, 8);
                launch_update_lazy((*s->cur), lp.lazy_begin, lp.lazy_end, s->lazy_ctr + 2 * l, s->nsm, 0, st);
            } else {
                cudaStream_t side = (nbatch % s->nside == 0) ? s->side_stream : s->side2[nbatch % s->nside - 1];
                nbatch++;
                CK(cudaEventRecord(s->ev_f[l], st));
                CK(cudaStreamWaitEvent(side, s->ev_f[l], 0));
                if (split) CK(cudaStreamWaitEvent(side, s->ev_tr[l], 0));
                const bool late = s->chain_sms_late >= 0 && (double)l >= s->chain_switch * (double)nlev;
                launch_update_lazy((*s->cur), lp.lazy_begin, lp.lazy_end, s->lazy_ctr + 2 * l, s->nsm, late ? s->chain_sms_late : s->chain_sms, side);
                CK(cudaEventRecord(s->ev_lazy[l], side));
                has_lazy[l] = 1;
            }
            count++;
        }
        if (oz && lp.ozs_end > lp.ozs_begin) {
            // digit planes of this level's pieces (rows from column block j + 3 down), then the left-looking tcgen05 tasks
            // of column block j + 3 over the pieces 0 .. j
            cudaStream_t ss = st, us = st;
            if (overlap) {
                ss = s->oz_slice_stream;
                us = s->oz_stream;      // one stream: at most nsm - oz_sms_free tcgen05 CTAs are ever resident
                CK(cudaEventRecord(s->ev_ozt[l], st));
                CK(cudaStreamWaitEvent(ss, s->ev_ozt[l], 0));
                if (split) CK(cudaStreamWaitEvent(ss, s->ev_tr[l], 0));
            }
            {
                Scope sc(s, 16);
                for (int32_t x = lp.ozs_begin; x < lp.ozs_end; ++x) {
                    const OzSlice& sl = s->plan.oz_slices[x];
                    const OzViewPlan& v = s->plan.oz_views[sl.view];
                    const Piece& pc = s->plan.pieces[sl.piece];
                    const int32_t sn = v.sn;
                    const int64_t ld = s->sym.sn_rowptr[sn + 1] - s->sym.sn_rowptr[sn];
                    launch_oz_slice(s->ctx.Lx + s->sym.sn_xptr[sn], ld, (int32_t)ld, pc.c0 - s->sym.sn_first[sn], pc.c1 - pc.c0, sl.j + 3,
                                    v.nrb - (sl.j + 3), 4 * sl.j, s->oz_E + v.row0, s->d_oz_rb_off + v.off0, s->oz_planes, ss);
                    count++;
                }
            }
            if (overlap) {
                CK(cudaEventRecord(s->ev_ozs[l], ss));
                CK(cudaStreamWaitEvent(us, s->ev_ozs[l], 0));
            }
            {
                Scope sc(s, 17);
                launch_oz_update(s->d_oz_views, s->d_oz_tasks, lp.oz_begin, lp.oz_end, s->oz_ctr + 2 * l, s->nsm, overlap ? s->oz_sms_free : 0,
                                 s->ctx.info + 3, lp.oz_tile, us);
                count++;
            }
            if (overlap) {
                CK(cudaEventRecord(s->ev_oz[l], us));
                has_oz[l] = 1;
            }
        }
        if (pack_level >= 0 && (long)l >= pack_level && (lp.pack_end > pack_done || lp.inv_end > inv_done)) {
            // a slice of the invert / repack work per level: short-lived CTAs that fill idle SMs of the tail without
            // holding them against the chain kernels
            cudaStream_t ps = s->side2[2];
            if (!has_work(s, tlpb200_solver::WL_LAZY, lp.lazy_begin, lp.lazy_end)) CK(cudaEventRecord(s->ev_f[l], st));
            CK(cudaStreamWaitEvent(ps, s->ev_f[l], 0));
            if (split) CK(cudaStreamWaitEvent(ps, s->ev_tr[l], 0));
            if (lp.inv_end > inv_done) { launch_invert_diag((*s->cur), inv_done, lp.inv_end, ps); inv_done = lp.inv_end; count++; }
            const int32_t pe = std::min(lp.pack_end, pack_done + s->pack_slice);
            if (pe > pack_done) { launch_pack_big((*s->cur), pack_done, pe, ps); pack_done = pe; count++; }
            CK(cudaEventRecord(s->ev_pack, ps));
            early_pack = true;
            ev_f_recorded = true;
        }
        if (!split) {
            if (has_work(s, tlpb200_solver::WL_EXT, lp.ext_begin, lp.ext_end)) { Scope sc(s, 4); launch_update((*s->cur), lp.ext_begin, lp.ext_end, 1, st); count++; }
        } else {
            if (!has_work(s, tlpb200_solver::WL_LAZY, lp.lazy_begin, lp.lazy_end) && !ev_f_recorded) CK(cudaEventRecord(s->ev_f[l], st));   // critical trsm done
            CK(cudaStreamWaitEvent(s->aux_stream, s->ev_f[l], 0));
            if (has_work(s, tlpb200_solver::WL_EXT, lp.ext_crit_end, lp.ext_end)) { launch_update((*s->cur), lp.ext_crit_end, lp.ext_end, 1, s->aux_stream); count++; }
            CK(cudaEventRecord(s->ev_ur[l], s->aux_stream));
            if (has_work(s, tlpb200_solver::WL_EXT, lp.ext_begin, lp.ext_crit_end)) { launch_update((*s->cur), lp.ext_begin, lp.ext_crit_end, 1, st, s->crit_ksplit); count++; }
        }
    }
    if (overlap && s->split_chain && nlev > 0) CK(cudaStreamWaitEvent(st, s->ev_ur[nlev - 1], 0));
    if (overlap)
        for (long q = waited + 1; q < (long)nlev; ++q)
            if (has_lazy[q]) CK(cudaStreamWaitEvent(st, s->ev_lazy[q], 0));   // join
    if (oz && overlap)
        for (long q = oz_waited + 1; q < (long)nlev; ++q)
            if (has_oz[q]) CK(cudaStreamWaitEvent(st, s->ev_oz[q], 0));   // join
    if (early_pack) CK(cudaStreamWaitEvent(st, s->ev_pack, 0));
    if (has_work(s, tlpb200_solver::WL_INV, inv_done, (*s->cur).ndblk)) { Scope sc(s, 10); launch_invert_diag((*s->cur), inv_done, (*s->cur).ndblk, st); count++; }
    if (has_work(s, tlpb200_solver::WL_PACK, pack_done, (int64_t)s->plan.big_pack.size())) { Scope sc(s, 13); launch_pack_big((*s->cur), pack_done, (int32_t)s->plan.big_pack.size(), st); count++; }
    if (s->dc.nd > 0) {
        // W = L^{-1} A_d : one forward sweep per dense column (kept as the sweeps leave it, plus its (G G')^{-1} image), then
        // the nd x nd Schur matrix D_d^{-1} + W'W and its Cholesky factor (kernels_dense_cols.cu)
        Scope sc(s, 12);
        for (int j = 0; j < s->dc.nd; ++j) {
            CK(cudaMemsetAsync(s->ctx.wk, 0, (size_t)s->sym.N * 8, st));
            reset_sweep_state(s, st);
            launch_dc_scatter(s->ctx, s->dc, j, s->dc_colptr[j + 1] - s->dc_colptr[j], st);
            enqueue_fwd(s, count);
            CK(cudaMemcpyAsync(s->dc.Wt + (size_t)j * s->sym.N, s->ctx.wk, (size_t)s->sym.N * 8, cudaMemcpyDeviceToDevice, st));
            launch_dc_ginv(s->ctx, s->dc, s->dc.Wt + (size_t)j * s->sym.N, s->dc.Wh + (size_t)j * s->sym.N, st);
            count += 3;
        }
        launch_dc_gram_chol(s->ctx, s->dc, s->d_theta, s->d_regP, st);
        count += 2;
    }
    CK(cudaGetLastError());
}

// hand-over state of one forward + backward sweep: "block solved" flags and the dependency counters of the merged levels
// (allocated back to back: flags [2 * ndblk], then dep_cnt [2 * nsuper])
void reset_sweep_state(tlpb200_solver* s, cudaStream_t st) {
    const size_t n = (size_t)2 * s->ctx.ndblk + (size_t)3 * s->sym.nsuper;
    if (n > 0) CK(cudaMemsetAsync(s->ctx.flags, 0, n * sizeof(int32_t), st));
}

void enqueue_rhs(tlpb200_solver* s, const double* xip, const double* xid, int64_t& count) {
    cudaStream_t st = s->stream;
    {
        Scope sc(s, 5);
        if (s->system == TLPB200_K1) launch_k1_rhs((*s->cur), s->mat, s->d_d, xip, xid, st);
        else launch_k2_rhs((*s->cur), s->mat, xip, xid, st);
    }
    count++;
    reset_sweep_state(s, st);
}

void enqueue_fwd(tlpb200_solver* s, int64_t& count) {
    cudaStream_t st = s->stream;
    if (s->merge_levels) {
        // the plan's launch sequence with the block-solve items of consecutive levels merged (Plan::SolveOp); a sharded phase
        // walks the same sequence with its own dependency targets (ctxA / ctxB) and drops the ops it has no item in
        for (const SolveOp& op : s->plan.fwd_ops) {
            const int wl = op.kind == 0 ? tlpb200_solver::WL_SMALL : (op.kind == 1 ? tlpb200_solver::WL_FWD : tlpb200_solver::WL_FBIG);
            if (!has_work(s, wl, op.begin, op.end)) continue;
            if (op.kind == 0) { Scope sc(s, 6); launch_fwd_small((*s->cur), op.begin, op.end, st); }
            else if (op.kind == 1) { Scope sc(s, 7); launch_fwd_large((*s->cur), op.begin, op.end, s->nsm, 1, st); }
            else { Scope sc(s, 14); launch_fwd_big((*s->cur), op.begin, op.end, s->nsm, st); }
            count++;
        }
        return;
    }
    const auto& L = s->plan.levels;
    for (size_t l = 0; l < L.size(); ++l) {
        const LevelPlan& lp = L[l];
        if (has_work(s, tlpb200_solver::WL_SMALL, lp.small_begin, lp.small_end)) { Scope sc(s, 6); launch_fwd_small((*s->cur), lp.small_begin, lp.small_end, st); count++; }
        if (has_work(s, tlpb200_solver::WL_FWD, lp.fwd_begin, lp.fwd_end)) { Scope sc(s, 7); launch_fwd_large((*s->cur), lp.fwd_begin, lp.fwd_end, s->nsm, 0, st); count++; }
        if (has_work(s, tlpb200_solver::WL_FBIG, lp.fbig_begin, lp.fbig_end)) { Scope sc(s, 14); launch_fwd_big((*s->cur), lp.fbig_begin, lp.fbig_end, s->nsm, st); count++; }
    }
}

void enqueue_bwd(tlpb200_solver* s, int64_t& count) {
    cudaStream_t st = s->stream;
    if (s->merge_levels) {
        for (const SolveOp& op : s->plan.bwd_ops) {
            const int wl = op.kind == 3 ? tlpb200_solver::WL_BELOW : (op.kind == 2 ? tlpb200_solver::WL_BBIG
                         : (op.kind == 1 ? tlpb200_solver::WL_BSEQ : tlpb200_solver::WL_SMALL));
            if (!has_work(s, wl, op.begin, op.end)) continue;
            if (op.kind == 3) { Scope sc(s, 9); launch_bwd_below((*s->cur), op.begin, op.end, st); }
            else if (op.kind == 2) { Scope sc(s, 15); launch_bwd_big((*s->cur), op.begin, op.end, s->nsm, st); }
            else if (op.kind == 1) { Scope sc(s, 9); launch_bwd_large((*s->cur), op.begin, op.end, s->nsm, 1, st); }
            else { Scope sc(s, 11); launch_bwd_small((*s->cur), op.begin, op.end, st); }
            count++;
        }
        return;
    }
    const auto& L = s->plan.levels;
    for (size_t l = L.size(); l-- > 0;) {
        const LevelPlan& lp = L[l];
        if (has_work(s, tlpb200_solver::WL_BELOW, lp.below_begin, lp.below_end)) { Scope sc(s, 9); launch_bwd_below((*s->cur), lp.below_begin, lp.below_end, st); count++; }
        if (has_work(s, tlpb200_solver::WL_BBIG, lp.bbig_begin, lp.bbig_end)) { Scope sc(s, 15); launch_bwd_big((*s->cur), lp.bbig_begin, lp.bbig_end, s->nsm, st); count++; }
        if (has_work(s, tlpb200_solver::WL_BWD, lp.bwd_begin, lp.bwd_end)) { Scope sc(s, 9); launch_bwd_large((*s->cur), lp.bwd_begin, lp.bwd_end, s->nsm, 0, st); count++; }
        if (has_work(s, tlpb200_solver::WL_SMALL, lp.small_begin, lp.small_end)) { Scope sc(s, 11); launch_bwd_small((*s->cur), lp.small_begin, lp.small_end, st); count++; }
    }
}

void enqueue_recover(tlpb200_solver* s, const double* xid, double* dx, double* dy, int64_t& count) {
    cudaStream_t st = s->stream;
    {
        Scope sc(s, 5);
        if (s->system == TLPB200_K1) launch_k1_recover((*s->cur), s->mat, s->d_d, xid, dx, dy, st);
        else launch_k2_recover((*s->cur), s->mat, dx, dy, st);
    }
    count++;
    CK(cudaGetLastError());
}

void enqueue_solve(tlpb200_solver* s, const double* xip, const double* xid, double* dx, double* dy, int64_t& count) {
    enqueue_rhs(s, xip, xid, count);
    const size_t nb = (size_t)s->sym.N * 8;
    if (s->dc.nd > 0 || s->refine > 0) CK(cudaMemcpyAsync(s->dc_xi, s->ctx.wk, nb, cudaMemcpyDeviceToDevice, s->stream));
    enqueue_fwd(s, count);
    if (s->dc.nd > 0) {      // Schur correction of the dense columns, half-way between the sweeps
        Scope sc(s, 12);
        launch_dc_apply(s->ctx, s->dc, s->stream);
        count += 2;
    }
    enqueue_bwd(s, count);
    if (s->dc.nd == 0 && s->refine > 0) {
        // iterative refinement inside solve! (the reference only has TODOs: spd.jl:68, sqd.jl:72): residual of the
        // KKT system actually factored (K1: A D A' + Rd, K2: the augmented matrix) in FP64 from A itself, corrected through
        // the same factor.  Off by default (refine_steps = 0 reproduces the reference's single solve).
        Scope sc(s, 19);
        for (int it = 0; it < s->refine; ++it) {
            CK(cudaMemcpyAsync(s->dc_y, s->ctx.wk, nb, cudaMemcpyDeviceToDevice, s->stream));
            if (s->system == TLPB200_K1) launch_dc_residual(s->ctx, s->mat, s->d_d, s->d_regD, s->dc_xi, s->dc_y, s->dc_tn, s->stream);
            else launch_k2_residual(s->ctx, s->mat, s->d_theta, s->d_regP, s->d_regD, s->dc_xi, s->dc_y, s->stream);
            reset_sweep_state(s, s->stream);
            enqueue_fwd(s, count);
            enqueue_bwd(s, count);
            launch_dc_axpy(s->ctx, s->dc_y, s->stream);
            count += 4;
        }
    }
    if (s->dc.nd > 0) {
        // iterative refinement on the full system with the Schur-corrected solve as the preconditioner
        Scope sc(s, 12);
        for (int it = 0; it < s->dc_refine; ++it) {
            CK(cudaMemcpyAsync(s->dc_y, s->ctx.wk, nb, cudaMemcpyDeviceToDevice, s->stream));
            launch_dc_residual(s->ctx, s->mat, s->d_d, s->d_regD, s->dc_xi, s->dc_y, s->dc_tn, s->stream);
            reset_sweep_state(s, s->stream);
            enqueue_fwd(s, count);
            launch_dc_apply(s->ctx, s->dc, s->stream);
            enqueue_bwd(s, count);
            launch_dc_axpy(s->ctx, s->dc_y, s->stream);
            count += 5;
        }
    }
    enqueue_recover(s, xid, dx, dy, count);
}

// Sharded (nranks > 1) update!: own subtrees -> all-reduce(sum) of the partial top panels over NVLink -> replicated top
// part -> one max-all-reduce of the status words (so that PosDefException / time-outs are raised on EVERY rank, step.jl:34-51).
// Everything is enqueued on the solver's stream: no host synchronisation between the phases.
void enqueue_update_all(tlpb200_solver* s, int64_t& cnt) {
    if (!sharded(s)) {
        enqueue_assemble(s, cnt);
        enqueue_factor(s, cnt);
        return;
    }
    const NcclApi& api = nccl_api();
    const int64_t top_cnt = s->sym.lx_size - s->top_begin;
    s->cur = &s->ctx;
    enqueue_assemble(s, cnt, true);
    // original entries of the replicated top part are contributed by rank 0 only
    if (s->rank != 0 && top_cnt > 0) CK(cudaMemsetAsync(s->ctx.Lx + s->top_begin, 0, (size_t)top_cnt * 8, s->stream));
    s->cur = &s->ctxA;
    enqueue_factor(s, cnt);
    if (top_cnt > 0) {
        Scope sc(s, 18);
        NK(api.AllReduce(s->ctx.Lx + s->top_begin, s->ctx.Lx + s->top_begin, (size_t)top_cnt, ncclFloat64, ncclSum, s->comm, s->stream));
        cnt++;
    }
    CK(cudaMemsetAsync(s->lazy_ctr, 0, (2 * s->plan.levels.size() + 2) * sizeof(int32_t), s->stream));
    s->cur = &s->ctxB;
    enqueue_factor(s, cnt);
    s->cur = &s->ctx;
    {
        Scope sc(s, 18);
        launch_pack_info(s->ctx.info, s->d_info_tmp, s->stream);
        NK(api.AllReduce(s->d_info_tmp, s->d_info_tmp, 4, ncclInt32, ncclMax, s->comm, s->stream));
        launch_unpack_info(s->ctx.info, s->d_info_tmp, s->stream);
        cnt += 3;
    }
}

// Sharded solve!: forward sweep on own subtrees -> all-reduce(sum) of the separator entries only (ntop doubles)
// -> replicated top forward + backward -> backward sweep on own subtrees -> all-reduce of the zero-padded solution (every
// rank's host IPM needs the whole vector) -> recovery.  One stream-ordered sequence.
void enqueue_solve_all(tlpb200_solver* s, const double* xip, const double* xid, double* dx, double* dy, int64_t& cnt) {
    if (!sharded(s)) {
        enqueue_solve(s, xip, xid, dx, dy, cnt);
        return;
    }
    const NcclApi& api = nccl_api();
    if (!s->d_keep || !s->ctxA.skip || !s->ctxB.skip || (s->ntop > 0 && (!s->d_top_cols || !s->d_tbuf)))
        throw std::runtime_error("sharded solver: device state of the phases is incomplete");
    s->cur = &s->ctx;
    enqueue_rhs(s, xip, xid, cnt);
    launch_zero_unowned(s->ctx, s->d_keep, s->stream);
    s->cur = &s->ctxA;
    enqueue_fwd(s, cnt);
    if (s->ntop > 0) {
        Scope sc(s, 18);
        launch_gather_top(s->ctx.wk, s->d_top_cols, s->ntop, s->d_tbuf, s->stream);
        NK(api.AllReduce(s->d_tbuf, s->d_tbuf, (size_t)s->ntop, ncclFloat64, ncclSum, s->comm, s->stream));
        launch_scatter_top(s->ctx.wk, s->d_top_cols, s->ntop, s->d_tbuf, s->stream);
        cnt += 3;
    }
    s->cur = &s->ctxB;
    enqueue_fwd(s, cnt);
    enqueue_bwd(s, cnt);
    s->cur = &s->ctxA;
    enqueue_bwd(s, cnt);
    s->cur = &s->ctx;
    launch_zero_unowned(s->ctx, s->d_keep, s->stream);
    {
        Scope sc(s, 18);
        NK(api.AllReduce(s->ctx.wk, s->ctx.wk, (size_t)s->sym.N, ncclFloat64, ncclSum, s->comm, s->stream));
        cnt += 3;
    }
    enqueue_recover(s, xid, dx, dy, cnt);
}

void destroy_graphs(tlpb200_solver* s) {
    if (s->g_update) { cudaGraphExecDestroy(s->g_update); s->g_update = nullptr; }
    if (s->g_solve) { cudaGraphExecDestroy(s->g_solve); s->g_solve = nullptr; }
}

void build_graphs(tlpb200_solver* s) {
    destroy_graphs(s);
    cudaGraph_t g = nullptr;
    int64_t cnt = 0;
    CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    try {
        enqueue_update_all(s, cnt);
    } catch (...) {
        s->cur = &s->ctx;
        cudaStreamEndCapture(s->stream, &g);
        if (g) cudaGraphDestroy(g);
        throw;
    }
    CK(cudaStreamEndCapture(s->stream, &g));
    s->launches_update = cnt;
    CK(cudaGraphInstantiate(&s->g_update, g, 0));
    CK(cudaGraphDestroy(g));
    g = nullptr;
    cnt = 0;
    CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    try {
        enqueue_solve_all(s, s->d_xip, s->d_xid, s->d_dx, s->d_dy, cnt);
    } catch (...) {
        s->cur = &s->ctx;
        cudaStreamEndCapture(s->stream, &g);
        if (g) cudaGraphDestroy(g);
        throw;
    }
    CK(cudaStreamEndCapture(s->stream, &g));
    s->launches_solve = cnt;
    CK(cudaGraphInstantiate(&s->g_solve, g, 0));
    CK(cudaGraphDestroy(g));
}

void run_update(tlpb200_solver* s) {
    if (sharded(s) && !s->comm) throw std::runtime_error("solver was created with nranks > 1: call tlpb200_comm_init first (or drive the phase API: update_begin / update_end)");
    const bool graph = s->opt.use_graph && (!sharded(s) || s->dist_graph);
    if (s->profiling && !sharded(s)) {
        int64_t cnt = 0;
        CK(cudaEventRecord(s->ev[0], s->stream));
        enqueue_assemble(s, cnt);
        CK(cudaEventRecord(s->ev[1], s->stream));
        enqueue_factor(s, cnt);
        CK(cudaEventRecord(s->ev[2], s->stream));
        s->launches_update = cnt;
    } else if (s->profiling) {
        int64_t cnt = 0;
        CK(cudaEventRecord(s->ev[0], s->stream));
        CK(cudaEventRecord(s->ev[1], s->stream));
        enqueue_update_all(s, cnt);
        CK(cudaEventRecord(s->ev[2], s->stream));
        s->launches_update = cnt;
    } else if (graph) {
        if (!s->g_update) build_graphs(s);
        CK(cudaGraphLaunch(s->g_update, s->stream));
    } else {
        int64_t cnt = 0;
        enqueue_update_all(s, cnt);
        s->launches_update = cnt;
    }
    CK(cudaMemcpyAsync(s->h_info, s->ctx.info, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    s->n_update++;
}

// info[2] (hand-over of a triangular sweep timed out) / info[3] (tcgen05 pipeline timed out) are raised by the kernels and
// reported by the call that observes them; the device flags are cleared so that the handle's state stays well defined.
int check_kernel_timeouts(tlpb200_solver* s, const char* where) {
    const int32_t sweep = s->h_info[2], oz = s->h_info[3];
    if (sweep == 0 && oz == 0) return TLPB200_OK;
    cudaMemsetAsync(s->ctx.info + 2, 0, 2 * sizeof(int32_t), s->stream);
    s->h_info[2] = s->h_info[3] = 0;
    std::string msg = std::string(where) + ": ";
    if (sweep) msg += "hand-over of a triangular sweep timed out (a block solution was never published; results are invalid)";
    if (oz) msg += std::string(sweep ? "; " : "") + "tcgen05 update pipeline timed out (mbarrier never completed)";
    return fail(s, TLPB200_INTERNAL, msg);
}

int finish_update(tlpb200_solver* s, int64_t* bad_pivot) {
    CK(cudaStreamSynchronize(s->stream));
    if (s->profiling) {
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, s->ev[0], s->ev[1]));
        CK(cudaEventElapsedTime(&b, s->ev[1], s->ev[2]));
        s->ms_assemble = a;
        s->ms_factor = b;
        collect_profile(s, true);
    }
    const int32_t info = *s->h_info;
    if (const int rc = check_kernel_timeouts(s, "update!")) return rc;
    s->bad_pivot = (info >= 0 && info < s->sym.N) ? info : -1;
    if (bad_pivot) *bad_pivot = s->bad_pivot;
    if (s->bad_pivot >= 0) {
        char buf[160];
        snprintf(buf, sizeof buf, "factorisation breakdown: wrong-sign, zero or NaN pivot at permuted column %lld",
                 (long long)s->bad_pivot);
        return fail(s, TLPB200_NOT_POSDEF, buf);
    }
    return TLPB200_OK;
}

// solve one right-hand side held in the internal buffers d_xip/d_xid -> d_dx/d_dy
void run_solve_internal(tlpb200_solver* s) {
    if (sharded(s) && !s->comm) throw std::runtime_error("solver was created with nranks > 1: call tlpb200_comm_init first (or drive the phase API: solve_begin / solve_mid / solve_end)");
    const bool graph = s->opt.use_graph && (!sharded(s) || s->dist_graph);
    if (s->profiling) {
        int64_t cnt = 0;
        CK(cudaEventRecord(s->ev[0], s->stream));
        enqueue_solve_all(s, s->d_xip, s->d_xid, s->d_dx, s->d_dy, cnt);
        CK(cudaEventRecord(s->ev[3], s->stream));
        s->launches_solve = cnt;
    } else if (graph) {
        if (!s->g_solve) build_graphs(s);
        CK(cudaGraphLaunch(s->g_solve, s->stream));
    } else {
        int64_t cnt = 0;
        enqueue_solve_all(s, s->d_xip, s->d_xid, s->d_dx, s->d_dy, cnt);
        s->launches_solve = cnt;
    }
    s->n_solve++;
}

void canonicalize(tlpb200_solver* s, const int64_t* colptr, const int64_t* rowval, const double* nzval, int base) {
    const int64_t n = s->n, m = s->m;
    s->colptr.assign(n + 1, 0);
    std::vector<std::pair<int32_t, double>> col;
    for (int64_t j = 0; j < n; ++j) {
        const int64_t b = colptr[j] - base, e = colptr[j + 1] - base;
        if (b < 0 || e < b) throw std::invalid_argument("colptr is not monotone");
        col.clear();
        for (int64_t p = b; p < e; ++p) {
            const int64_t r = rowval[p] - base;
            if (r < 0 || r >= m) throw std::invalid_argument("row index out of range");
            col.emplace_back((int32_t)r, nzval[p]);
        }
        std::stable_sort(col.begin(), col.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        for (size_t x = 0; x < col.size(); ++x) {
            if (x > 0 && col[x].first == col[x - 1].first) s->val.back() += col[x].second;
            else { s->rowidx.push_back(col[x].first); s->val.push_back(col[x].second); }
        }
        s->colptr[j + 1] = (int64_t)s->rowidx.size();
    }
    s->nnz = s->colptr[n];
}

void setup_device(tlpb200_solver* s) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) throw CudaFail{e == cudaSuccess ? cudaErrorNoDevice : e, "cudaGetDeviceCount (no CUDA device: this backend has no CPU fallback)"};
    if (s->opt.device < 0 || s->opt.device >= ndev) throw std::invalid_argument("device ordinal out of range");
    CK(cudaSetDevice(s->opt.device));
    s->device = s->opt.device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, s->device));
    if (prop.major != 10) throw std::runtime_error("tlpb200 is built for sm_100a (B200) only; found compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor));
    {
        int lo = 0, hi = 0;   // numerically lower = higher priority
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&s->own_stream, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&s->side_stream, cudaStreamNonBlocking, lo));
        for (auto& x : s->side2) CK(cudaStreamCreateWithPriority(&x, cudaStreamNonBlocking, lo));
        CK(cudaStreamCreateWithPriority(&s->aux_stream, cudaStreamNonBlocking, std::min(lo, hi + 1)));
        if (const char* e = getenv("TLPB200_SPLIT_CHAIN")) s->split_chain = atoi(e) != 0;
        if (const char* e = getenv("TLPB200_CRIT_KSPLIT")) s->crit_ksplit = std::max(1, std::min(8, atoi(e)));
        if (const char* e = getenv("TLPB200_PACK_SPLIT")) s->pack_split = atof(e);
        if (const char* e = getenv("TLPB200_PACK_SLICE")) s->pack_slice = std::max(1, atoi(e));
        CK(cudaEventCreateWithFlags(&s->ev_pack, cudaEventDisableTiming));
        if (const char* e = getenv("TLPB200_SIDE_STREAMS")) s->nside = std::max(1, std::min(4, atoi(e)));
    }
    s->stream = s->own_stream;
    CK(kernels_static_init());
    CK(dense_solve_static_init());
    CK(ozaki_static_init());
    for (auto& ev : s->ev) CK(cudaEventCreate(&ev));

    const Symbolic& S = s->sym;
    const Plan& P = s->plan;
    DevCtx& c = s->ctx;
    c.N = S.N;
    c.sn_first = upload(s, S.sn_first);
    c.sn_rowptr = upload(s, S.sn_rowptr);
    c.sn_rows = upload(s, S.sn_rows);
    c.sn_xptr = upload(s, S.sn_xptr);
    c.col2sn = upload(s, S.col2sn);
    c.sign = upload(s, S.sign);
    c.diagpos = upload(s, S.diagpos);
    c.perm = upload(s, S.perm);
    c.iperm = upload(s, S.iperm);
    c.pieces = upload(s, P.pieces);
    c.seg_ptr = upload(s, P.seg_ptr);
    c.seg_k0 = upload(s, P.seg_k0);
    c.seg_tgt = upload(s, P.seg_tgt);
    c.small_list = upload(s, P.small_list);
    c.level_pieces = upload(s, P.level_pieces);
    c.upd = upload(s, P.upd);
    c.upd_lazy = upload(s, P.upd128);
    s->lazy_ctr = dalloc<int32_t>(s, 2 * P.levels.size() + 2);
    c.panel = upload(s, P.panel);
    c.fwd_items = upload(s, P.fwd_items);
    c.bwd_items = upload(s, P.bwd_items);
    c.sn_dblk = upload(s, P.sn_dblk);
    c.dblk_sn = upload(s, P.dblk_sn);
    c.dblk_idx = upload(s, P.dblk_idx);
    c.inv_order = upload(s, P.inv_order);
    c.ndblk = P.ndblk;
    c.has_neg = (s->system == TLPB200_K2) ? 1 : 0;
    c.Lx = dalloc<double>(s, (size_t)S.lx_size);
    c.Dinv = dalloc<double>(s, (size_t)P.ndblk * SBLK * SBLK);
    c.DinvT = dalloc<double>(s, (size_t)P.ndblk * SBLK * SBLK);
    c.LsubT = dalloc<double>(s, (size_t)P.ndblk * SBLK * SBLK);
    c.flags = dalloc<int32_t>(s, (size_t)2 * P.ndblk + (size_t)3 * S.nsuper);
    CK(cudaMemset(c.flags, 0, std::max<size_t>((size_t)2 * P.ndblk + (size_t)3 * S.nsuper, 1) * sizeof(int32_t)));
    c.dep_cnt = c.flags + (size_t)2 * P.ndblk;
    c.fwd_need = upload(s, P.fwd_need);
    c.fwd_parent = upload(s, P.fwd_parent);
    c.bwd_wait = upload(s, P.bwd_wait);
    c.bwd_nitems = upload(s, P.bwd_nitems);
    c.bwd_nbelow = upload(s, P.bwd_nbelow);
    c.bwd_seq = upload(s, P.bwd_seq);
    c.nsuper = S.nsuper;
    if (const char* e = getenv("TLPB200_MERGE_LEVELS")) s->merge_levels = atoi(e) != 0;
    c.info = dalloc<int32_t>(s, 4);
    CK(cudaMemset(c.info, 0, 4 * sizeof(int32_t)));
    c.big_pack = upload(s, P.big_pack);
    c.fwd_big = upload(s, P.fwd_big);
    c.bwd_big = upload(s, P.bwd_big);
    c.Ft = dalloc<double>(s, (size_t)P.n_ftiles * SBLK * SBLK);
    c.Bt = dalloc<double>(s, (size_t)P.n_btiles * SBLK * SBLK);
    c.bwd_below = upload(s, P.bwd_below);
    c.sn_split = upload(s, P.sn_split);
    c.bacc = dalloc<double>(s, (size_t)S.N);
    CK(cudaMemset(c.bacc, 0, std::max<size_t>(S.N, 1) * sizeof(double)));
    c.xq = dalloc<unsigned long long>(s, (size_t)2 * P.xq_slots);
    CK(cudaMemset(c.xq, 0, std::max<size_t>(2 * (size_t)P.xq_slots, 1) * sizeof(unsigned long long)));
    c.trace_min = c.trace_max = nullptr;
    if (getenv("TLPB200_TRACE_FACTOR")) {
        c.trace_min = dalloc<unsigned long long>(s, 4 * P.levels.size());
        c.trace_max = dalloc<unsigned long long>(s, 4 * P.levels.size());
    }
    c.nxblk = P.xq_slots / SBLK;
    c.dbg_ts = nullptr;
    if (getenv("TLPB200_CHAIN_TIMES")) {
        c.dbg_ts = dalloc<unsigned long long>(s, (size_t)2 * c.nxblk);
        CK(cudaMemset(c.dbg_ts, 0, std::max<size_t>(2 * (size_t)c.nxblk, 1) * sizeof(unsigned long long)));
    }
    c.epoch = dalloc<unsigned long long>(s, 2);
    CK(cudaMemset(c.epoch, 0, 2 * sizeof(unsigned long long)));
    c.wk = dalloc<double>(s, (size_t)S.N);
    CK(cudaMemset(c.wk, 0, std::max<size_t>(S.N, 1) * sizeof(double)));
    c.skip = nullptr;
    s->cur = &s->ctx;
    if (s->nranks > 1) {
        std::vector<int8_t> skipA(S.nsuper), skipB(S.nsuper), keep(S.N);
        for (int32_t sn = 0; sn < S.nsuper; ++sn) {
            skipA[sn] = (s->owner[sn] != s->rank);
            skipB[sn] = (s->owner[sn] != -1);
            const int8_t k = (s->owner[sn] == s->rank) || (s->owner[sn] == -1 && s->rank == 0);
            for (int32_t j = S.sn_first[sn]; j < S.sn_first[sn + 1]; ++j) keep[j] = k;
        }
        s->ctxA = s->ctx; s->ctxA.skip = upload(s, skipA);
        s->ctxB = s->ctx; s->ctxB.skip = upload(s, skipB);
        // merged-level sweeps inside a phase: per-phase dependency targets (build_phase_deps, host)
        for (int ph = 0; ph < 2; ++ph) {
            DevCtx& cx = ph == 0 ? s->ctxA : s->ctxB;
            cx.fwd_need = upload(s, s->phase_need[ph]);
            cx.fwd_parent = upload(s, s->phase_fpar[ph]);
            cx.bwd_wait = upload(s, s->phase_bwait[ph]);
        }
        s->d_keep = const_cast<int8_t*>(upload(s, keep));
        std::vector<int32_t> top_cols;
        for (int32_t sn = 0; sn < S.nsuper; ++sn)
            if (s->owner[sn] == -1)
                for (int32_t j = S.sn_first[sn]; j < S.sn_first[sn + 1]; ++j) top_cols.push_back(j);
        s->ntop = (int32_t)top_cols.size();
        s->d_top_cols = upload(s, top_cols);
        s->d_tbuf = dalloc<double>(s, std::max<size_t>(top_cols.size(), 1));
        s->d_info_tmp = dalloc<int32_t>(s, 4);
        if (const char* e = getenv("TLPB200_DIST_GRAPH")) s->dist_graph = atoi(e) != 0;
    }
    if (!P.oz_views.empty()) {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&s->oz_slice_stream, cudaStreamNonBlocking, std::min(lo, hi + 1)));
        CK(cudaStreamCreateWithPriority(&s->oz_stream, cudaStreamNonBlocking, lo));
        s->oz_planes = dalloc<uint8_t>(s, (size_t)P.oz_slots * OZ_S * 4096);
        CK(cudaMemset(s->oz_planes, 0, (size_t)P.oz_slots * OZ_S * 4096));
        s->d_oz_rb_off = upload(s, P.oz_rb_off);
        s->d_oz_tasks = upload(s, P.oz_tasks);
        s->oz_E = dalloc<int32_t>(s, (size_t)P.oz_rows);
        s->oz_scl = dalloc<double>(s, (size_t)P.oz_rows);
        s->oz_ctr = dalloc<int32_t>(s, 2 * P.levels.size() + 2);
        std::vector<OzView> hv(P.oz_views.size());
        for (size_t i = 0; i < hv.size(); ++i) {
            const OzViewPlan& v = P.oz_views[i];
            OzView& o = hv[i];
            o.planes = s->oz_planes;
            o.rb_off = s->d_oz_rb_off + v.off0;
            o.C = c.Lx + S.sn_xptr[v.sn];
            o.ldc = S.sn_rowptr[v.sn + 1] - S.sn_rowptr[v.sn];
            o.scl = s->oz_scl + v.row0;
            o.nrows = (int32_t)o.ldc;
            o.ncols = S.sn_first[v.sn + 1] - S.sn_first[v.sn];
        }
        s->d_oz_views = const_cast<OzView*>(upload(s, hv));
        s->oz_on = true;
        if (const char* e = getenv("TLPB200_OZAKI_FREE_SMS")) s->oz_sms_free = std::max(0, std::min(prop.multiProcessorCount - 1, atoi(e)));
    }
    s->nsm = prop.multiProcessorCount;
    if (const char* e = getenv("TLPB200_CHAIN_SMS")) s->chain_sms = std::max(0, std::min(s->nsm - 1, atoi(e)));
    if (const char* e = getenv("TLPB200_CHAIN_SMS_LATE")) s->chain_sms_late = std::max(0, std::min(s->nsm - 1, atoi(e)));
    if (const char* e = getenv("TLPB200_CHAIN_SWITCH")) s->chain_switch = atof(e);

    DevMat& A = s->mat;
    A.m = s->m; A.n = s->n; A.nnz = s->nnz;
    A.colptr = upload(s, s->colptr);
    A.rowidx = upload(s, s->rowidx);
    A.val = upload(s, s->val);
    {   // CSR copy for the K1 right-hand side product
        std::vector<int64_t> rp(s->m + 1, 0);
        for (int64_t p = 0; p < s->nnz; ++p) rp[s->rowidx[p] + 1]++;
        for (int64_t i = 0; i < s->m; ++i) rp[i + 1] += rp[i];
        std::vector<int32_t> ci(s->nnz);
        std::vector<double> rv(s->nnz);
        std::vector<int64_t> nxt(rp.begin(), rp.end() - 1);
        for (int64_t j = 0; j < s->n; ++j)
            for (int64_t p = s->colptr[j]; p < s->colptr[j + 1]; ++p) {
                const int64_t q = nxt[s->rowidx[p]]++;
                ci[q] = (int32_t)j;
                rv[q] = s->val[p];
            }
        A.rowptr = upload(s, rp);
        A.colidx = upload(s, ci);
        A.rval = upload(s, rv);
    }
    {
        std::vector<int32_t> lc;
        for (int64_t j = 0; j < s->n; ++j)
            if (s->colptr[j + 1] - s->colptr[j] > LONG_COL) lc.push_back((int32_t)j);
        A.nlong = (int32_t)lc.size();
        A.long_cols = upload(s, lc);
    }
    A.nentries = (int64_t)s->maps.w_dest.size();
    A.w_ptr = upload(s, s->maps.w_ptr);
    A.w_dest = upload(s, s->maps.w_dest);
    A.w_col = upload(s, s->maps.w_col);
    A.w_val = upload(s, s->maps.w_val);
    A.a_dest = upload(s, s->maps.a_dest);
    if (!s->dense_cols.empty()) {
        const int nd = (int)s->dense_cols.size();
        std::vector<int32_t> prow;
        std::vector<double> dval;
        s->dc_colptr.assign(nd + 1, 0);
        for (int i = 0; i < nd; ++i) {
            const int32_t j = s->dense_cols[i];
            for (int64_t p = s->colptr[j]; p < s->colptr[j + 1]; ++p) { prow.push_back(S.iperm[s->rowidx[p]]); dval.push_back(s->val[p]); }
            s->dc_colptr[i + 1] = (int64_t)prow.size();
        }
        s->dc.nd = nd;
        s->dc.colptr = upload(s, s->dc_colptr);
        s->dc.prow = upload(s, prow);
        s->dc.val = upload(s, dval);
        s->dc.col_id = upload(s, s->dense_cols);
        s->dc.Wt = dalloc<double>(s, (size_t)nd * S.N);
        s->dc.Wh = dalloc<double>(s, (size_t)nd * S.N);
        {   // diagonal blocks of the dense-solve supernodes: the forward sweeps leave those entries scaled by L_kk
            std::vector<int32_t> gb;
            for (int32_t sn = 0; sn < S.nsuper; ++sn) {
                if (sn >= (int32_t)P.sn_big.size() || !P.sn_big[sn]) continue;
                const int32_t f = S.sn_first[sn], nc = S.sn_first[sn + 1] - f;
                for (int32_t k = 0; k * SBLK < nc; ++k) {
                    gb.push_back(P.sn_dblk[sn] + k);
                    gb.push_back(f + k * SBLK);
                    gb.push_back(std::min<int32_t>(SBLK, nc - k * SBLK));
                }
            }
            s->dc.ngblk = (int32_t)(gb.size() / 3);
            s->dc.gblk = upload(s, gb);
        }
        s->dc.C = dalloc<double>(s, (size_t)nd * nd);
        s->dc.g = dalloc<double>(s, nd);
        s->dc_xi = dalloc<double>(s, S.N);
        s->dc_y = dalloc<double>(s, S.N);
        s->dc_tn = dalloc<double>(s, s->n);
        if (const char* e = getenv("TLPB200_DC_REFINE")) s->dc_refine = std::max(0, atoi(e));
    }

    s->refine = (s->nranks == 1) ? std::max(0, std::min(8, (int)s->opt.refine_steps)) : 0;
    if (const char* e = getenv("TLPB200_REFINE")) s->refine = (s->nranks == 1) ? std::max(0, std::min(8, atoi(e))) : 0;
    if (s->refine > 0 && s->dense_cols.empty()) {
        s->dc_xi = dalloc<double>(s, S.N);
        s->dc_y = dalloc<double>(s, S.N);
        s->dc_tn = dalloc<double>(s, s->n);
    }
    s->d_theta = dalloc<double>(s, s->n);
    s->d_regP = dalloc<double>(s, s->n);
    s->d_regD = dalloc<double>(s, s->m);
    s->d_d = dalloc<double>(s, s->n);
    s->d_xip = dalloc<double>(s, s->m);
    s->d_xid = dalloc<double>(s, s->n);
    s->d_dx = dalloc<double>(s, s->n);
    s->d_dy = dalloc<double>(s, s->m);
    CK(cudaMallocHost((void**)&s->h_pin, std::max<size_t>(2 * (size_t)(s->n + s->m) + (size_t)s->n, 1) * sizeof(double)));
    CK(cudaMallocHost((void**)&s->h_info, 4 * sizeof(int32_t)));
    s->small_smem = small_factor_smem(P.max_small_elems, P.max_small_nrow);
    s->on_device = true;
}

}  // namespace

// ---- internal entry points used by the device-resident IPM loop (ipm.cu) -----------------------------------------------------
namespace tlp_internal {
void run_update(tlpb200_solver* s) {
    try { ::run_update(s); }
    catch (const CudaFail& f) { cuda_fail(s, f); throw std::runtime_error(s->err); }
    catch (const NcclFail& f) { nccl_fail(s, f); throw std::runtime_error(s->err); }
}
int finish_update(tlpb200_solver* s, int64_t* bad_pivot) {
    try { return ::finish_update(s, bad_pivot); }
    catch (const CudaFail& f) { return cuda_fail(s, f); }
}
void run_solve(tlpb200_solver* s) {
    try { ::run_solve_internal(s); }
    catch (const CudaFail& f) { cuda_fail(s, f); throw std::runtime_error(s->err); }
    catch (const NcclFail& f) { nccl_fail(s, f); throw std::runtime_error(s->err); }
}
int check_timeouts(tlpb200_solver* s, const char* where) { return ::check_kernel_timeouts(s, where); }
int set_error(tlpb200_solver* s, int code, const std::string& msg) { return ::fail(s, code, msg); }
void* device_alloc(tlpb200_solver* s, size_t bytes) {
    try { return (void*)::dalloc<char>(s, bytes); }
    catch (const CudaFail& f) { cuda_fail(s, f); throw std::runtime_error(s->err); }
}
}  // namespace tlp_internal

void tlpb200_ipm_free(tlpb200_ipm* ip);   // ipm.cu

extern "C" {

void tlpb200_default_options(tlpb200_options* o) {
    std::memset(o, 0, sizeof *o);
    o->ordering = 1;
    o->device = 0;
    o->piece_width = 128;
    o->small_elems = 4096;
    o->relax_always = 8;
    o->use_graph = 1;
    o->analyze_only = 0;
    o->rank = 0;
    o->nranks = 1;
}

int tlpb200_create(tlpb200_solver** out, int64_t m, int64_t n, const int64_t* colptr, const int64_t* rowval,
                   const double* nzval, int index_base, int system, const tlpb200_options* opt) {
    if (!out) return TLPB200_BAD_ARG;
    *out = nullptr;
    tlpb200_solver* s = new (std::nothrow) tlpb200_solver();
    if (!s) return TLPB200_OOM;
    *out = s;   // returned even on failure so that tlpb200_last_error() can be read; caller destroys
    if (opt) s->opt = *opt; else tlpb200_default_options(&s->opt);
    if (m < 0 || n < 0 || !colptr || (system != TLPB200_K1 && system != TLPB200_K2) || (index_base != 0 && index_base != 1))
        return fail(s, TLPB200_BAD_ARG, "tlpb200_create: bad argument");
    if ((system == TLPB200_K1 ? m : m + n) > (int64_t)INT32_MAX - 1)
        return fail(s, TLPB200_BAD_ARG, "tlpb200_create: system order exceeds 32-bit indexing");
    s->opt.piece_width = PIECE;
    if (s->opt.small_elems < 64) s->opt.small_elems = 4096;
    if (s->opt.small_elems > 8192) s->opt.small_elems = 8192;
    s->system = system;
    s->m = m;
    s->n = n;
    try {
        static const bool trace_setup = getenv("TLPB200_TRACE") != nullptr;
        auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        double t_mark = now();
        auto lap = [&](const char* what) {
            if (!trace_setup) return;
            const double t = now();
            fprintf(stderr, "tlpb200 setup: %-22s %8.3f s\n", what, t - t_mark);
            t_mark = t;
        };
        canonicalize(s, colptr, rowval, nzval, index_base);
        lap("canonicalize");
        SymOptions so;
        so.ordering = s->opt.ordering;
        if (s->opt.relax_always > 0) so.relax_always = s->opt.relax_always;
        s->sym.system = system;
        s->sym.m = m;
        s->sym.n = n;
        // dense columns (K1 only): more than max(32, 5% of m) non-zeros, at most 64 of them
        std::vector<int64_t> fcolptr;       // A with the dense columns emptied (pattern + assemble use this)
        std::vector<int32_t> frowidx;
        std::vector<double> fval;
        if (system == TLPB200_K1 && s->opt.dense_col_threshold >= 0 && std::max(1, s->opt.nranks) == 1) {
            const int64_t thr = s->opt.dense_col_threshold > 0 ? s->opt.dense_col_threshold
                                                               : std::max<int64_t>(32, (int64_t)(0.05 * (double)m));
            for (int64_t j = 0; j < n && s->dense_cols.size() < 64; ++j)
                if (s->colptr[j + 1] - s->colptr[j] > thr) s->dense_cols.push_back((int32_t)j);
        }
        const int64_t* cp_f = s->colptr.data();
        const int32_t* ri_f = s->rowidx.data();
        const double* va_f = s->val.data();
        if (!s->dense_cols.empty()) {
            std::vector<char> isd(n, 0);
            for (int32_t j : s->dense_cols) isd[j] = 1;
            fcolptr.assign(n + 1, 0);
            for (int64_t j = 0; j < n; ++j) {
                if (!isd[j])
                    for (int64_t p = s->colptr[j]; p < s->colptr[j + 1]; ++p) { frowidx.push_back(s->rowidx[p]); fval.push_back(s->val[p]); }
                fcolptr[j + 1] = (int64_t)frowidx.size();
            }
            cp_f = fcolptr.data(); ri_f = frowidx.data(); va_f = fval.data();
        }
        if (system == TLPB200_K1) {
            SymPattern P = pattern_k1(m, n, cp_f, ri_f);
            lap("pattern");
            analyze_pattern(P, so, nullptr, s->sym);
        } else {
            SymPattern P = pattern_k2(m, n, s->colptr.data(), s->rowidx.data());
            std::vector<int8_t> sg(n + m, 1);
            for (int64_t j = 0; j < n; ++j) sg[j] = -1;   // first n pivots < 0, last m > 0 (systems.jl:10-32)
            analyze_pattern(P, so, sg.data(), s->sym);
        }
        lap("analyze_pattern");
        s->rank = s->opt.rank;
        s->nranks = std::max(1, s->opt.nranks);
        if (s->rank < 0 || s->rank >= s->nranks) throw std::invalid_argument("rank out of range");
        partition_subtrees(s->sym, s->nranks, s->owner);
        if (s->nranks > 1) s->top_begin = relayout_panels(s->sym, s->owner, s->nranks, &s->rank_begin);
        PlanOptions po;
        po.small_elems = s->opt.small_elems;
        if (s->opt.dense_solve_ncol > 0) po.big_ncol = s->opt.dense_solve_ncol;
        if (const char* e = getenv("TLPB200_DENSE_SOLVE_NCOL")) po.big_ncol = std::max(1, atoi(e));
        // tcgen05 int8 path: single-GPU K1 only (the K2 panels carry negative pivots; sharded runs keep the FP64 path)
        po.oz_ncol = (system == TLPB200_K1 && s->nranks == 1) ? (s->opt.ozaki_ncol != 0 ? s->opt.ozaki_ncol : po.oz_ncol) : -1;
        if (const char* e = getenv("TLPB200_OZAKI_NCOL")) { if (po.oz_ncol > 0 || atoi(e) <= 0) po.oz_ncol = atoi(e); }
        if (const char* e = getenv("TLPB200_OZAKI_TILE")) po.oz_tile_n = atoi(e) == 128 ? 128 : (atoi(e) == 64 ? 64 : 0);
        if (const char* e = getenv("TLPB200_OZAKI_KSPLIT")) po.oz_ksplit = std::max(32, (atoi(e) / 32) * 32);
        build_plan(s->sym, po, s->plan);
        if (s->nranks > 1) { build_work_prefix(s); build_phase_deps(s); }
        lap("build_plan");
        if (system == TLPB200_K1)
            build_assembly_k1(s->sym, m, n, cp_f, ri_f, va_f, s->maps);
        else
            build_assembly_k2(s->sym, m, n, s->colptr.data(), s->rowidx.data(), s->maps);
        lap("assembly maps");
        if (!s->opt.analyze_only) setup_device(s);
        lap("setup_device");
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const std::bad_alloc&) {
        return fail(s, TLPB200_OOM, "host allocation failed during setup");
    } catch (const std::invalid_argument& e) {
        return fail(s, TLPB200_BAD_ARG, e.what());
    } catch (const std::exception& e) {
        return fail(s, TLPB200_INTERNAL, e.what());
    }
    return TLPB200_OK;
}

#define REQUIRE_DEVICE(s)                                                                              \
    if (!(s)) return TLPB200_BAD_ARG;                                                                  \
    if (!(s)->on_device) return fail((s), TLPB200_CUDA, "solver has no device state (analyze_only or failed setup); there is no CPU fallback")

int tlpb200_update(tlpb200_solver* s, const double* theta_inv, const double* regP, const double* regD,
                   int64_t* bad_pivot) {
    REQUIRE_DEVICE(s);
    if (!theta_inv || !regP || !regD) return fail(s, TLPB200_BAD_ARG, "tlpb200_update: null vector");
    try {
        static const bool trace = getenv("TLPB200_TRACE") != nullptr;
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t0 = trace ? now() : 0;
        CK(cudaSetDevice(s->device));
        const size_t n = (size_t)s->n, m = (size_t)s->m;
        const double t1 = trace ? now() : 0;
        stage_h2d(s, s->d_theta, s->h_pin, theta_inv, n);
        stage_h2d(s, s->d_regP, s->h_pin + n, regP, n);
        stage_h2d(s, s->d_regD, s->h_pin + 2 * n, regD, m);
        const double t2 = trace ? now() : 0;
        run_update(s);
        const double t3 = trace ? now() : 0;
        const int rc = finish_update(s, bad_pivot);
        if (trace) fprintf(stderr, "[tlpb200 trace] update: stage %.3f  h2d-enqueue %.3f  launch %.3f  sync %.3f ms\n", t1 - t0, t2 - t1, t3 - t2, now() - t3);
        return rc;
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const NcclFail& f) {
        s->cur = &s->ctx;
        return nccl_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

int tlpb200_update_dev(tlpb200_solver* s, const double* d_theta_inv, const double* d_regP, const double* d_regD) {
    REQUIRE_DEVICE(s);
    try {
        CK(cudaSetDevice(s->device));
        CK(cudaMemcpyAsync(s->d_theta, d_theta_inv, (size_t)s->n * 8, cudaMemcpyDeviceToDevice, s->stream));
        CK(cudaMemcpyAsync(s->d_regP, d_regP, (size_t)s->n * 8, cudaMemcpyDeviceToDevice, s->stream));
        CK(cudaMemcpyAsync(s->d_regD, d_regD, (size_t)s->m * 8, cudaMemcpyDeviceToDevice, s->stream));
        run_update(s);
        return TLPB200_OK;
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const NcclFail& f) {
        s->cur = &s->ctx;
        return nccl_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

int tlpb200_update_status(tlpb200_solver* s, int64_t* bad_pivot) {
    REQUIRE_DEVICE(s);
    try {
        return finish_update(s, bad_pivot);
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const NcclFail& f) {
        s->cur = &s->ctx;
        return nccl_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

int tlpb200_solve(tlpb200_solver* s, double* dx, double* dy, const double* xi_p, const double* xi_d, int32_t nrhs,
                  int64_t ldx, int64_t ldy) {
    REQUIRE_DEVICE(s);
    if (!dx || !dy || !xi_p || !xi_d || nrhs < 1 || (nrhs > 1 && (ldx < s->n || ldy < s->m)))
        return fail(s, TLPB200_BAD_ARG, "tlpb200_solve: bad argument");
    try {
        CK(cudaSetDevice(s->device));
        const size_t n = (size_t)s->n, m = (size_t)s->m;
        double* hin = s->h_pin;
        double* hout = s->h_pin + (n + m);
        for (int32_t r = 0; r < nrhs; ++r) {
            stage_h2d(s, s->d_xip, hin, xi_p + (size_t)r * ldy, m);
            stage_h2d(s, s->d_xid, hin + m, xi_d + (size_t)r * ldx, n);
            run_solve_internal(s);
            // status words first (tiny), then the results chunk by chunk: a timed-out sweep is reported before anything is
            // handed to the caller, and the CPU copy of a chunk overlaps the DMA of the next one
            CK(cudaMemcpyAsync(s->h_info, s->ctx.info, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
            const StageOut ox = stage_d2h_begin(s, dx + (size_t)r * ldx, hout, s->d_dx, n, 0);
            const StageOut oy = stage_d2h_begin(s, dy + (size_t)r * ldy, hout + n, s->d_dy, m, stage_chunks(n));
            if (n + m > 0) CK(cudaEventSynchronize(s->ev_stage[0])); else CK(cudaStreamSynchronize(s->stream));
            if (const int rc = check_kernel_timeouts(s, "solve!")) { cudaStreamSynchronize(s->stream); return rc; }
            stage_d2h_finish(s, ox);
            stage_d2h_finish(s, oy);
            CK(cudaStreamSynchronize(s->stream));
            if (s->profiling) {
                float a = 0;
                CK(cudaEventElapsedTime(&a, s->ev[0], s->ev[3]));
                s->ms_solve = a;
                collect_profile(s, false);
            }
        }
        return TLPB200_OK;
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const NcclFail& f) {
        s->cur = &s->ctx;
        return nccl_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

int tlpb200_solve_dev(tlpb200_solver* s, double* d_dx, double* d_dy, const double* d_xi_p, const double* d_xi_d,
                      int32_t nrhs, int64_t ldx, int64_t ldy) {
    REQUIRE_DEVICE(s);
    if (nrhs < 1) return fail(s, TLPB200_BAD_ARG, "tlpb200_solve_dev: nrhs < 1");
    try {
        CK(cudaSetDevice(s->device));
        const size_t n = (size_t)s->n, m = (size_t)s->m;
        for (int32_t r = 0; r < nrhs; ++r) {
            CK(cudaMemcpyAsync(s->d_xip, d_xi_p + (size_t)r * ldy, m * 8, cudaMemcpyDeviceToDevice, s->stream));
            CK(cudaMemcpyAsync(s->d_xid, d_xi_d + (size_t)r * ldx, n * 8, cudaMemcpyDeviceToDevice, s->stream));
            run_solve_internal(s);
            CK(cudaMemcpyAsync(d_dx + (size_t)r * ldx, s->d_dx, n * 8, cudaMemcpyDeviceToDevice, s->stream));
            CK(cudaMemcpyAsync(d_dy + (size_t)r * ldy, s->d_dy, m * 8, cudaMemcpyDeviceToDevice, s->stream));
        }
        return TLPB200_OK;
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const NcclFail& f) {
        s->cur = &s->ctx;
        return nccl_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

int tlpb200_solve_status(tlpb200_solver* s) {
    REQUIRE_DEVICE(s);
    try {
        CK(cudaSetDevice(s->device));
        CK(cudaMemcpyAsync(s->h_info, s->ctx.info, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        return check_kernel_timeouts(s, "solve! (device-pointer variant)");
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const NcclFail& f) {
        s->cur = &s->ctx;
        return nccl_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

int tlpb200_debug_raise_timeout(tlpb200_solver* s) {
    REQUIRE_DEVICE(s);
    const int32_t one = 1;
    cudaError_t e = cudaMemcpyAsync(s->ctx.info + 2, &one, sizeof one, cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) return cuda_fail(s, CudaFail{e, "tlpb200_debug_raise_timeout"});
    return TLPB200_OK;
}

int tlpb200_set_stream(tlpb200_solver* s, void* cuda_stream) {
    REQUIRE_DEVICE(s);
    cudaStreamSynchronize(s->stream);
    s->stream = cuda_stream ? (cudaStream_t)cuda_stream : s->own_stream;
    destroy_graphs(s);   // graphs are captured per stream
    return TLPB200_OK;
}

int tlpb200_synchronize(tlpb200_solver* s) {
    REQUIRE_DEVICE(s);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) return cuda_fail(s, CudaFail{e, "cudaStreamSynchronize"});
    return TLPB200_OK;
}

int tlpb200_set_profiling(tlpb200_solver* s, int on) {
    REQUIRE_DEVICE(s);
    s->profiling = on != 0;
    return TLPB200_OK;
}

int tlpb200_stats_get(const tlpb200_solver* s, tlpb200_stats* o) {
    if (!s || !o) return TLPB200_BAD_ARG;
    std::memset(o, 0, sizeof *o);
    o->m = s->m; o->n = s->n; o->nnzA = s->nnz;
    o->order = s->sym.N;
    o->nnzL = s->sym.nnzL;
    o->nnzL_stored = s->sym.lx_size;
    o->flops = s->sym.flops;
    o->nsuper = s->sym.nsuper;
    o->npieces = (int64_t)s->plan.pieces.size();
    o->nlevels = (int64_t)s->plan.levels.size();
    o->max_ncol = s->sym.max_ncol;
    o->max_nrow = s->sym.max_nrow;
    o->nproducts = (int64_t)s->maps.w_col.size();
    o->nentries = (int64_t)s->maps.w_dest.size();
    o->launches_update = s->launches_update;
    o->launches_solve = s->launches_solve;
    o->ms_assemble = s->ms_assemble;
    o->ms_factor = s->ms_factor;
    o->ms_solve = s->ms_solve;
    o->bad_pivot = s->bad_pivot;
    o->n_update = s->n_update;
    o->n_solve = s->n_solve;
    o->bytes_device = (int64_t)s->bytes_device;
    o->flops_update_inner = s->plan.flops_panel;
    o->flops_update_ext = s->plan.flops_update;
    for (int c = 0; c < TLPB200_NCLASS; ++c) { o->ms_class[c] = s->ms_class[c]; o->n_class[c] = s->n_class[c]; }
    o->flops_update_oz = s->plan.flops_oz;
    o->oz_tasks = (int64_t)s->plan.oz_tasks.size();
    o->oz_bytes = s->plan.oz_slots * (int64_t)(OZ_S * 4096);
    return TLPB200_OK;
}

int tlpb200_get_symbolic(const tlpb200_solver* s, int32_t* perm, int32_t* parent, int32_t* colcount, int32_t* sn_first) {
    if (!s) return TLPB200_BAD_ARG;
    const Symbolic& S = s->sym;
    if (perm) std::copy(S.perm.begin(), S.perm.end(), perm);
    if (parent) std::copy(S.parent.begin(), S.parent.end(), parent);
    if (colcount) std::copy(S.colcount.begin(), S.colcount.end(), colcount);
    if (sn_first) std::copy(S.sn_first.begin(), S.sn_first.end(), sn_first);
    return TLPB200_OK;
}

int tlpb200_get_structure(const tlpb200_solver* s, int64_t* rowptr, int32_t* rows) {
    if (!s) return TLPB200_BAD_ARG;
    const Symbolic& S = s->sym;
    if (rowptr) std::copy(S.sn_rowptr.begin(), S.sn_rowptr.end(), rowptr);
    if (rows) std::copy(S.sn_rows.begin(), S.sn_rows.end(), rows);
    return TLPB200_OK;
}

int tlpb200_debug_assemble(tlpb200_solver* s, const double* theta_inv, const double* regP, const double* regD) {
    REQUIRE_DEVICE(s);
    try {
        CK(cudaSetDevice(s->device));
        CK(cudaMemcpyAsync(s->d_theta, theta_inv, (size_t)s->n * 8, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(s->d_regP, regP, (size_t)s->n * 8, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(s->d_regD, regD, (size_t)s->m * 8, cudaMemcpyHostToDevice, s->stream));
        int64_t cnt = 0;
        enqueue_assemble(s, cnt);
        CK(cudaStreamSynchronize(s->stream));
        return TLPB200_OK;
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const NcclFail& f) {
        s->cur = &s->ctx;
        return nccl_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

int tlpb200_debug_get_lx(tlpb200_solver* s, double* lx, int64_t* xptr) {
    REQUIRE_DEVICE(s);
    try {
        CK(cudaSetDevice(s->device));
        CK(cudaStreamSynchronize(s->stream));
        if (lx) CK(cudaMemcpy(lx, s->ctx.Lx, (size_t)s->sym.lx_size * 8, cudaMemcpyDeviceToHost));
        if (xptr) std::copy(s->sym.sn_xptr.begin(), s->sym.sn_xptr.end(), xptr);
        return TLPB200_OK;
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const NcclFail& f) {
        s->cur = &s->ctx;
        return nccl_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

/* globaltimer (ns) of every block publish of the last dense sweeps: out[0..nblk) forward, out[nblk..2 nblk) backward;
   only recorded when the solver was created with TLPB200_CHAIN_TIMES set in the environment */
int tlpb200_debug_chain_times(tlpb200_solver* s, uint64_t* out, int64_t* nblk) {
    REQUIRE_DEVICE(s);
    if (nblk) *nblk = s->ctx.nxblk;
    if (!s->ctx.dbg_ts) return fail(s, TLPB200_BAD_ARG, "chain times were not recorded (TLPB200_CHAIN_TIMES unset at setup)");
    cudaStreamSynchronize(s->stream);
    if (out) cudaMemcpy(out, s->ctx.dbg_ts, (size_t)2 * s->ctx.nxblk * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    return TLPB200_OK;
}

/* factorisation timeline of the last update! (TLPB200_TRACE_FACTOR set at create): out[level][cls][0/1] = globaltimer ns of
   the first CTA start / last CTA end of cls = {diag, trsm, urgent update, lazy update}; 0 = class absent at that level */
int tlpb200_debug_factor_trace(tlpb200_solver* s, uint64_t* out, int64_t* nlevels) {
    REQUIRE_DEVICE(s);
    const size_t nl = s->plan.levels.size();
    if (nlevels) *nlevels = (int64_t)nl;
    if (!s->ctx.trace_min) return fail(s, TLPB200_BAD_ARG, "factor trace was not recorded (TLPB200_TRACE_FACTOR unset at setup)");
    if (!out) return TLPB200_OK;
    cudaStreamSynchronize(s->stream);
    std::vector<unsigned long long> mn(4 * nl), mx(4 * nl);
    cudaMemcpy(mn.data(), s->ctx.trace_min, mn.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(mx.data(), s->ctx.trace_max, mx.size() * 8, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < 4 * nl; ++i) {
        out[2 * i] = (mx[i] == 0) ? 0 : mn[i];
        out[2 * i + 1] = mx[i];
    }
    return TLPB200_OK;
}

int tlpb200_debug_big_plan(const tlpb200_solver* s, int64_t* counts, void* pack, void* fwd, void* bwd) {
    if (!s) return TLPB200_BAD_ARG;
    const Plan& P = s->plan;
    if (counts) {
        counts[0] = (int64_t)P.big_pack.size();
        counts[1] = (int64_t)P.fwd_big.size();
        counts[2] = (int64_t)P.bwd_big.size();
        counts[3] = P.n_ftiles;
        counts[4] = P.n_btiles;
        counts[5] = P.xq_slots;
    }
    if (pack && !P.big_pack.empty()) std::memcpy(pack, P.big_pack.data(), P.big_pack.size() * sizeof(BigPack));
    if (fwd && !P.fwd_big.empty()) std::memcpy(fwd, P.fwd_big.data(), P.fwd_big.size() * sizeof(BigTask));
    if (bwd && !P.bwd_big.empty()) std::memcpy(bwd, P.bwd_big.data(), P.bwd_big.size() * sizeof(BigTask));
    return TLPB200_OK;
}

// update-task plan (host data; tests/test_update_plan.py checks that the FP64 tiles and the tcgen05 tasks cover every
// (piece, target entry) pair exactly once).  Arrays are int32 records: upd / upd128 = UpdTask (8), oz = OzTask (8),
// pieces = Piece (4), views = {sn, nrb, ncb, base_level}.
int tlpb200_debug_update_plan(const tlpb200_solver* s, int64_t* counts, int32_t* upd, int32_t* upd128, int32_t* oz, int32_t* pieces,
                              int32_t* views, int32_t* panel, int32_t* levels, int32_t* small_list, int32_t* level_pieces) {
    if (!s) return TLPB200_BAD_ARG;
    const Plan& P = s->plan;
    static_assert(sizeof(UpdTask) == 32 && sizeof(OzTask) == 32 && sizeof(Piece) == 16 && sizeof(PanelTask) == 16, "record sizes");
    static_assert(sizeof(LevelPlan) % sizeof(int32_t) == 0, "LevelPlan is a record of int32");
    if (counts) {
        counts[0] = (int64_t)P.upd.size();
        counts[1] = (int64_t)P.upd128.size();
        counts[2] = (int64_t)P.oz_tasks.size();
        counts[3] = (int64_t)P.pieces.size();
        counts[4] = (int64_t)P.oz_views.size();
        counts[5] = (int64_t)P.panel.size();
        counts[6] = (int64_t)P.levels.size();
        counts[7] = (int64_t)(sizeof(LevelPlan) / sizeof(int32_t));
        counts[8] = (int64_t)P.small_list.size();
        counts[9] = (int64_t)P.level_pieces.size();
    }
    if (small_list && !P.small_list.empty()) std::memcpy(small_list, P.small_list.data(), P.small_list.size() * sizeof(int32_t));
    if (level_pieces && !P.level_pieces.empty()) std::memcpy(level_pieces, P.level_pieces.data(), P.level_pieces.size() * sizeof(int32_t));
    if (panel && !P.panel.empty()) std::memcpy(panel, P.panel.data(), P.panel.size() * sizeof(PanelTask));
    if (levels && !P.levels.empty()) std::memcpy(levels, P.levels.data(), P.levels.size() * sizeof(LevelPlan));
    if (upd && !P.upd.empty()) std::memcpy(upd, P.upd.data(), P.upd.size() * sizeof(UpdTask));
    if (upd128 && !P.upd128.empty()) std::memcpy(upd128, P.upd128.data(), P.upd128.size() * sizeof(UpdTask));
    if (oz && !P.oz_tasks.empty()) std::memcpy(oz, P.oz_tasks.data(), P.oz_tasks.size() * sizeof(OzTask));
    if (pieces && !P.pieces.empty()) std::memcpy(pieces, P.pieces.data(), P.pieces.size() * sizeof(Piece));
    if (views)
        for (size_t i = 0; i < P.oz_views.size(); ++i) {
            views[4 * i + 0] = P.oz_views[i].sn;
            views[4 * i + 1] = P.oz_views[i].nrb;
            views[4 * i + 2] = P.oz_views[i].ncb;
            views[4 * i + 3] = P.oz_views[i].base_level;
        }
    return TLPB200_OK;
}

// launch sequences of the triangular sweeps with merged levels (host data; tests/test_dense_solve_plan.py replays them):
// counts[0..3] = #fwd ops, #bwd ops, #fwd items, #bwd items; ops as int32 records {kind, begin, end, level}; per-supernode
// arrays of length nsuper; items as 24-byte SolveItem records {sn, blk, kind, r0, nr, pad}.  Any pointer may be NULL.
int tlpb200_debug_solve_ops(const tlpb200_solver* s, int64_t* counts, int32_t* fwd_ops, int32_t* bwd_ops, int32_t* fwd_need,
                            int32_t* fwd_parent, int32_t* bwd_wait, int32_t* bwd_nitems, void* fwd_items, void* bwd_seq, int32_t* sn_parent,
                            int32_t* bwd_nbelow) {
    if (!s) return TLPB200_BAD_ARG;
    const Plan& P = s->plan;
    static_assert(sizeof(SolveOp) == 16 && sizeof(SolveItem) == 24, "record sizes");
    if (counts) {
        counts[0] = (int64_t)P.fwd_ops.size(); counts[1] = (int64_t)P.bwd_ops.size();
        counts[2] = (int64_t)P.fwd_items.size(); counts[3] = (int64_t)P.bwd_seq.size();
    }
    auto cp = [](void* dst, const void* src, size_t bytes) { if (dst && bytes) std::memcpy(dst, src, bytes); };
    cp(fwd_ops, P.fwd_ops.data(), P.fwd_ops.size() * sizeof(SolveOp));
    cp(bwd_ops, P.bwd_ops.data(), P.bwd_ops.size() * sizeof(SolveOp));
    cp(fwd_need, P.fwd_need.data(), P.fwd_need.size() * 4);
    cp(fwd_parent, P.fwd_parent.data(), P.fwd_parent.size() * 4);
    cp(bwd_wait, P.bwd_wait.data(), P.bwd_wait.size() * 4);
    cp(bwd_nitems, P.bwd_nitems.data(), P.bwd_nitems.size() * 4);
    cp(fwd_items, P.fwd_items.data(), P.fwd_items.size() * sizeof(SolveItem));
    cp(bwd_seq, P.bwd_seq.data(), P.bwd_seq.size() * sizeof(SolveItem));
    cp(sn_parent, s->sym.sn_parent.data(), s->sym.sn_parent.size() * 4);
    cp(bwd_nbelow, P.bwd_nbelow.data(), P.bwd_nbelow.size() * 4);
    return TLPB200_OK;
}

// per-phase dependency targets of the merged-level sweeps of a sharded solver (host data; phase 0 = own subtrees, 1 = top part)
int tlpb200_debug_phase_deps(const tlpb200_solver* s, int32_t phase, int32_t* fwd_need, int32_t* fwd_parent, int32_t* bwd_wait) {
    if (!s || phase < 0 || phase > 1) return TLPB200_BAD_ARG;
    if (s->nranks < 2) return TLPB200_BAD_ARG;
    auto cp = [](int32_t* dst, const std::vector<int32_t>& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * 4); };
    cp(fwd_need, s->phase_need[phase]);
    cp(fwd_parent, s->phase_fpar[phase]);
    cp(bwd_wait, s->phase_bwait[phase]);
    return TLPB200_OK;
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU (one process per GPU): subtree-sharded factorisation.  The caller (tulip.jl_b200/parallel.py,
// torch.distributed / NCCL) performs the collectives between the phases on the exposed device buffers.
// ---------------------------------------------------------------------------------------------
int tlpb200_dist_info(const tlpb200_solver* s, int32_t* owner, int64_t* top_offset, int64_t* top_count) {
    if (!s) return TLPB200_BAD_ARG;
    if (owner) std::copy(s->owner.begin(), s->owner.end(), owner);
    if (top_offset) *top_offset = s->nranks > 1 ? s->top_begin : s->sym.lx_size;
    if (top_count) *top_count = s->nranks > 1 ? s->sym.lx_size - s->top_begin : 0;
    return TLPB200_OK;
}

// assemble + factorisation of this rank's subtrees; leaves the partial top panels in Lx[top range]
int tlpb200_update_begin(tlpb200_solver* s, const double* theta_inv, const double* regP, const double* regD) {
    REQUIRE_DEVICE(s);
    if (s->nranks < 2) return fail(s, TLPB200_BAD_ARG, "tlpb200_update_begin: solver was not created with nranks > 1");
    try {
        CK(cudaSetDevice(s->device));
        const size_t n = (size_t)s->n, m = (size_t)s->m;
        std::memcpy(s->h_pin, theta_inv, n * 8);
        std::memcpy(s->h_pin + n, regP, n * 8);
        std::memcpy(s->h_pin + 2 * n, regD, m * 8);
        CK(cudaMemcpyAsync(s->d_theta, s->h_pin, n * 8, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(s->d_regP, s->h_pin + n, n * 8, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(s->d_regD, s->h_pin + 2 * n, m * 8, cudaMemcpyHostToDevice, s->stream));
        int64_t cnt = 0;
        s->cur = &s->ctx;
        enqueue_assemble(s, cnt);
        // original entries of the replicated top part are contributed by rank 0 only
        if (s->rank != 0 && s->sym.lx_size > s->top_begin)
            CK(cudaMemsetAsync(s->ctx.Lx + s->top_begin, 0, (size_t)(s->sym.lx_size - s->top_begin) * 8, s->stream));
        s->cur = &s->ctxA;
        enqueue_factor(s, cnt);
        s->cur = &s->ctx;
        s->launches_update = cnt;
        CK(cudaStreamSynchronize(s->stream));
        return TLPB200_OK;
    } catch (const CudaFail& f) {
        s->cur = &s->ctx;
        return cuda_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

int tlpb200_top_panels(tlpb200_solver* s, void** dptr, int64_t* count) {
    REQUIRE_DEVICE(s);
    if (dptr) *dptr = s->ctx.Lx + (s->nranks > 1 ? s->top_begin : s->sym.lx_size);
    if (count) *count = s->nranks > 1 ? s->sym.lx_size - s->top_begin : 0;
    return TLPB200_OK;
}

// (after the all-reduce of the top panels) factorisation of the replicated top part
int tlpb200_update_end(tlpb200_solver* s, int64_t* bad_pivot) {
    REQUIRE_DEVICE(s);
    if (s->nranks < 2) return fail(s, TLPB200_BAD_ARG, "tlpb200_update_end: solver was not created with nranks > 1");
    try {
        CK(cudaSetDevice(s->device));
        int64_t cnt = 0;
        CK(cudaMemsetAsync(s->lazy_ctr, 0, (2 * s->plan.levels.size() + 2) * sizeof(int32_t), s->stream));
        s->cur = &s->ctxB;
        enqueue_factor(s, cnt);
        s->cur = &s->ctx;
        s->launches_update += cnt;
        CK(cudaMemcpyAsync(s->h_info, s->ctx.info, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
        s->n_update++;
        return finish_update(s, bad_pivot);
    } catch (const CudaFail& f) {
        s->cur = &s->ctx;
        return cuda_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

// rhs build + forward sweep over this rank's subtrees; wk then holds this rank's contribution
int tlpb200_solve_begin(tlpb200_solver* s, const double* xi_p, const double* xi_d) {
    REQUIRE_DEVICE(s);
    if (s->nranks < 2) return fail(s, TLPB200_BAD_ARG, "tlpb200_solve_begin: solver was not created with nranks > 1");
    try {
        CK(cudaSetDevice(s->device));
        const size_t n = (size_t)s->n, m = (size_t)s->m;
        std::memcpy(s->h_pin, xi_p, m * 8);
        std::memcpy(s->h_pin + m, xi_d, n * 8);
        CK(cudaMemcpyAsync(s->d_xip, s->h_pin, m * 8, cudaMemcpyHostToDevice, s->stream));
        CK(cudaMemcpyAsync(s->d_xid, s->h_pin + m, n * 8, cudaMemcpyHostToDevice, s->stream));
        int64_t cnt = 0;
        s->cur = &s->ctx;
        enqueue_rhs(s, s->d_xip, s->d_xid, cnt);
        launch_zero_unowned(s->ctx, s->d_keep, s->stream);
        s->cur = &s->ctxA;
        enqueue_fwd(s, cnt);
        s->cur = &s->ctx;
        s->launches_solve = cnt + 1;
        CK(cudaStreamSynchronize(s->stream));
        return TLPB200_OK;
    } catch (const CudaFail& f) {
        s->cur = &s->ctx;
        return cuda_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

int tlpb200_work_vector(tlpb200_solver* s, void** dptr, int64_t* count) {
    REQUIRE_DEVICE(s);
    if (dptr) *dptr = s->ctx.wk;
    if (count) *count = s->sym.N;
    return TLPB200_OK;
}

// (after the all-reduce of wk) top part forward + backward, then backward over this rank's subtrees;
// wk again holds this rank's contribution (its own columns; the top part on rank 0)
int tlpb200_solve_mid(tlpb200_solver* s) {
    REQUIRE_DEVICE(s);
    if (s->nranks < 2) return fail(s, TLPB200_BAD_ARG, "tlpb200_solve_mid: solver was not created with nranks > 1");
    try {
        CK(cudaSetDevice(s->device));
        int64_t cnt = 0;
        s->cur = &s->ctxB;
        enqueue_fwd(s, cnt);
        enqueue_bwd(s, cnt);
        s->cur = &s->ctxA;
        enqueue_bwd(s, cnt);
        s->cur = &s->ctx;
        launch_zero_unowned(s->ctx, s->d_keep, s->stream);
        s->launches_solve += cnt + 1;
        CK(cudaStreamSynchronize(s->stream));
        return TLPB200_OK;
    } catch (const CudaFail& f) {
        s->cur = &s->ctx;
        return cuda_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

// (after the second all-reduce of wk) recovery of dx, dy on every rank
int tlpb200_solve_end(tlpb200_solver* s, double* dx, double* dy) {
    REQUIRE_DEVICE(s);
    if (s->nranks < 2) return fail(s, TLPB200_BAD_ARG, "tlpb200_solve_end: solver was not created with nranks > 1");
    try {
        CK(cudaSetDevice(s->device));
        const size_t n = (size_t)s->n, m = (size_t)s->m;
        int64_t cnt = 0;
        s->cur = &s->ctx;
        enqueue_recover(s, s->d_xid, s->d_dx, s->d_dy, cnt);
        double* hout = s->h_pin + (n + m);
        CK(cudaMemcpyAsync(hout, s->d_dx, n * 8, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaMemcpyAsync(hout + n, s->d_dy, m * 8, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaMemcpyAsync(s->h_info, s->ctx.info, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        if (const int rc = check_kernel_timeouts(s, "solve! (sharded)")) return rc;
        std::memcpy(dx, hout, n * 8);
        std::memcpy(dy, hout + n, m * 8);
        s->launches_solve += cnt;
        s->n_solve++;
        return TLPB200_OK;
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const NcclFail& f) {
        s->cur = &s->ctx;
        return nccl_fail(s, f);
    } catch (const std::exception& e) {
        s->cur = &s->ctx;
        return fail(s, TLPB200_INTERNAL, e.what());
    }
}

// ---- in-library collectives: NCCL over NVLink on the solver's stream ------------------------------------------------------
int tlpb200_comm_unique_id(void* out128) {
    if (!out128) return TLPB200_BAD_ARG;
    NcclApi& api = nccl_api();
    if (!api.h) return TLPB200_NCCL;
    ncclUniqueId id;
    if (api.GetUniqueId(&id) != ncclSuccess) return TLPB200_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out128, &id, sizeof id);
    return TLPB200_OK;
}

int tlpb200_comm_init(tlpb200_solver* s, const void* id128) {
    REQUIRE_DEVICE(s);
    if (!id128) return fail(s, TLPB200_BAD_ARG, "tlpb200_comm_init: null id");
    if (s->nranks < 2) return fail(s, TLPB200_BAD_ARG, "tlpb200_comm_init: solver was not created with nranks > 1");
    if (s->comm) return TLPB200_OK;
    NcclApi& api = nccl_api();
    if (!api.h) return fail(s, TLPB200_NCCL, api.err);
    try {
        CK(cudaSetDevice(s->device));
        ncclUniqueId id;
        std::memcpy(&id, id128, sizeof id);
        NK(api.CommInitRank(&s->comm, s->nranks, id, s->rank));
        // first collective outside any stream capture: NCCL sets up its channels / connections lazily
        CK(cudaMemsetAsync(s->d_info_tmp, 0, 4 * sizeof(int32_t), s->stream));
        NK(api.AllReduce(s->d_info_tmp, s->d_info_tmp, 4, ncclInt32, ncclMax, s->comm, s->stream));
        if (s->ntop > 0) {
            CK(cudaMemsetAsync(s->d_tbuf, 0, (size_t)s->ntop * 8, s->stream));
            NK(api.AllReduce(s->d_tbuf, s->d_tbuf, (size_t)s->ntop, ncclFloat64, ncclSum, s->comm, s->stream));
        }
        NK(api.AllReduce(s->ctx.wk, s->ctx.wk, (size_t)s->sym.N, ncclFloat64, ncclSum, s->comm, s->stream));
        const int64_t top_cnt = s->sym.lx_size - s->top_begin;
        if (top_cnt > 0) NK(api.AllReduce(s->ctx.Lx + s->top_begin, s->ctx.Lx + s->top_begin, (size_t)top_cnt, ncclFloat64, ncclSum, s->comm, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        destroy_graphs(s);
        return TLPB200_OK;
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const NcclFail& f) {
        return nccl_fail(s, f);
    }
}

/* mean time [ms] of the collectives of one sharded update! / solve!, each timed alone over `reps` back-to-back calls with CUDA
   events on the solver's stream: ms[0] = all-reduce of the top panels (update!), ms[1] = all-reduce of the separator entries
   (solve!), ms[2] = all-reduce of the solution vector (solve!), ms[3] = status all-reduce (update!); bytes[0..3] = their payloads */
int tlpb200_comm_profile(tlpb200_solver* s, int32_t reps, float* ms, int64_t* bytes) {
    REQUIRE_DEVICE(s);
    if (!s->comm) return fail(s, TLPB200_BAD_ARG, "tlpb200_comm_profile: no communicator (tlpb200_comm_init)");
    if (reps < 1) reps = 1;
    NcclApi& api = nccl_api();
    try {
        CK(cudaSetDevice(s->device));
        const int64_t top_cnt = s->sym.lx_size - s->top_begin;
        // scratch copies so that the factor / work vector are left alone
        double* scratch = nullptr;
        const size_t cnt = (size_t)std::max<int64_t>(std::max<int64_t>(top_cnt, s->sym.N), 4);
        CK(cudaMalloc((void**)&scratch, cnt * 8));
        CK(cudaMemsetAsync(scratch, 0, cnt * 8, s->stream));
        const size_t sizes[4] = {(size_t)top_cnt, (size_t)s->ntop, (size_t)s->sym.N, 4};
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int k = 0; k < 4; ++k) {
            if (bytes) bytes[k] = (int64_t)sizes[k] * (k == 3 ? 4 : 8);
            if (ms) ms[k] = 0.f;
            if (sizes[k] == 0) continue;
            for (int w = 0; w < 2; ++w) {
                if (k == 3) NK(api.AllReduce(scratch, scratch, 4, ncclInt32, ncclMax, s->comm, s->stream));
                else NK(api.AllReduce(scratch, scratch, sizes[k], ncclFloat64, ncclSum, s->comm, s->stream));
            }
            CK(cudaEventRecord(e0, s->stream));
            for (int r = 0; r < reps; ++r) {
                if (k == 3) NK(api.AllReduce(scratch, scratch, 4, ncclInt32, ncclMax, s->comm, s->stream));
                else NK(api.AllReduce(scratch, scratch, sizes[k], ncclFloat64, ncclSum, s->comm, s->stream));
            }
            CK(cudaEventRecord(e1, s->stream));
            CK(cudaStreamSynchronize(s->stream));
            float t = 0;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (ms) ms[k] = t / (float)reps;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaFree(scratch);
        return TLPB200_OK;
    } catch (const CudaFail& f) {
        return cuda_fail(s, f);
    } catch (const NcclFail& f) {
        return nccl_fail(s, f);
    }
}

int tlpb200_get_dense_cols(const tlpb200_solver* s, int32_t* count, int64_t* cols) {
    if (!s) return TLPB200_BAD_ARG;
    if (count) *count = (int32_t)s->dense_cols.size();
    if (cols) for (size_t i = 0; i < s->dense_cols.size(); ++i) cols[i] = s->dense_cols[i];
    return TLPB200_OK;
}

void tlpb200_abi_sizes(int32_t* out) {
    if (!out) return;
    out[0] = (int32_t)sizeof(tlpb200_options);
    out[1] = (int32_t)sizeof(tlpb200_stats);
    out[2] = TLPB200_NCLASS;
}

const char* tlpb200_last_error(const tlpb200_solver* s) { return s ? s->err.c_str() : "null solver"; }
const char* tlpb200_backend_name(void) { return "TlpB200 (supernodal signed Cholesky, CUDA sm_100a)"; }
const char* tlpb200_linear_system(const tlpb200_solver* s) {
    if (!s) return "Unknown";
    return s->system == TLPB200_K1 ? "Normal equations (K1)" : "Augmented system (K2)";
}

void tlpb200_destroy(tlpb200_solver* s) {
    if (!s) return;
    if (s->ipm) { tlpb200_ipm_free(s->ipm); s->ipm = nullptr; }
    if (s->on_device) {
        cudaSetDevice(s->device);
        cudaStreamSynchronize(s->stream);
        destroy_graphs(s);
        if (s->comm && nccl_api().h) { nccl_api().CommDestroy(s->comm); s->comm = nullptr; }
        for (void* p : s->allocs) cudaFree(p);
        if (s->h_pin) cudaFreeHost(s->h_pin);
        if (s->h_info) cudaFreeHost(s->h_info);
        for (auto& ev : s->ev) if (ev) cudaEventDestroy(ev);
        for (auto& ev : s->pool) cudaEventDestroy(ev);
        for (auto& ev : s->ev_stage) cudaEventDestroy(ev);
        for (auto& ev : s->ev_f) cudaEventDestroy(ev);
        for (auto& ev : s->ev_lazy) cudaEventDestroy(ev);
        if (s->side_stream) cudaStreamDestroy(s->side_stream);
        for (auto& x : s->side2) if (x) cudaStreamDestroy(x);
        if (s->aux_stream) cudaStreamDestroy(s->aux_stream);
        if (s->ev_pack) cudaEventDestroy(s->ev_pack);
        for (auto& ev : s->ev_d) cudaEventDestroy(ev);
        for (auto& ev : s->ev_tr) cudaEventDestroy(ev);
        for (auto& ev : s->ev_ur) cudaEventDestroy(ev);
        for (auto& ev : s->ev_ozt) cudaEventDestroy(ev);
        for (auto& ev : s->ev_ozs) cudaEventDestroy(ev);
        for (auto& ev : s->ev_oz) cudaEventDestroy(ev);
        if (s->oz_slice_stream) cudaStreamDestroy(s->oz_slice_stream);
        if (s->oz_stream) cudaStreamDestroy(s->oz_stream);
        if (s->own_stream) cudaStreamDestroy(s->own_stream);
    }
    delete s;
}

}  // extern "C"
