// ct_ctx.h -- internal: the context behind the C ABI (device scratch, pinned staging, streams) and small helpers shared by
// the translation units that implement the entry points (ct_api.cu, ct_fast.cu, ct_dist.cu).
#pragma once
#include "../../include/contrack_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "ct_host.h"
#include "ct_kernels.h"
#include "ct_tables.h"

namespace cti {

std::string& last_error();                     // thread-local message behind ct_last_error()

inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define CT_CUDA(expr)                                                                                                \
    do {                                                                                                             \
        cudaError_t e__ = (expr);                                                                                    \
        if (e__ != cudaSuccess)                                                                                      \
            return cti::fail(CT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);   \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // grow to at least `bytes`, preserving the first `keep` bytes (copied on `st`); `hint` = expected final size
    cudaError_t grow(size_t bytes, size_t keep, cudaStream_t st, size_t hint = 0) {
        if (bytes <= cap) return cudaSuccess;
        if (!p || keep == 0) {
            cudaError_t e = ensure(hint > bytes ? hint : bytes);
            return e == cudaSuccess ? e : ensure(bytes);
        }
        size_t want = std::max(bytes + bytes / 4 + 256, hint);
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, want);
        if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&q, want); }
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(q, p, std::min(keep, cap), cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(p);
        p = q; cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace cti

using cti::DevBuf;
using cti::PinBuf;

struct ct_ctx {
    int device = 0, sm_count = 148;
    long opt_tma = 1;                         // threshold kernel: rows staged by cp.async.bulk (0: plain coalesced loads)
    // geometry of the last run
    long T = 0; int H = 0, W = 0, Ww = 0;
    long nruns = 0, ncomp = 0, npair = 0, nseam = 0, novr = 0;
    // device scratch
    DevBuf bits, row_cnt, seam_flag, row_ptr, seam_pos, scan_tmp, counters, slots;
    DevBuf run_x, run_row, parent, root_flag, rank, run_comp, run_val;
    DevBuf c_t, c_y0, c_y1, c_x0, c_x1, c_E, c_S, c_nsp, c_cls, c_val;
    DevBuf s_row, s_a, s_b;
    DevBuf h_key, h_npix, h_nsp, h_E, h_S;
    DevBuf p_b, p_npix, p_nsp, p_E, p_S;
    DevBuf k_conE, k_conS, k_fE, k_fS, k_nsp, k_fnsp, pcnt, pfill, pptr;
    DevBuf seg_start, seg_pos, g_t, g_y0, g_y1, g_a, g_b;
    DevBuf o_t, o_y, o_x0, o_x1, o_val;
    DevBuf w_dev, special_dev, thr_dev;
    DevBuf chunk_in[2], chunk_out[2];
    DevBuf a_gptr, a_gidx, a_gmean, a_group;
    // pinned host staging
    PinBuf hp_counters, hp_tables, hp_val, hp_ovr;
    cth::Result host_result;
    cth::FastTables host_tb;                 // tables of the last tables_gpu() call (pointers into hp_tables)
    long nseg = 0, halo_comps = 0;
    int has_prev = 0;                        // sharded run: plane 0 of the scratch is the previous rank's last plane
    int special_uniform = 0;
    long opt_gpu_tables = 1;                 // step 3 + 3-D labels on the device (single-GPU path)
    long opt_chunks = 4;                     // time chunks of the pipelined run (tables of chunk k under threshold k+1)
    long opt_chunk_min_planes = 1024;        // ... but never fewer planes per chunk than this (launch latency of ~45 small
                                             // kernels and 3 host round trips per chunk)
    long tb_planes = 0, tb_runs = 0, tb_comps = 0, tb_seams = 0, tb_segs = 0, tb_pairs = 0;   // tables built so far
    std::vector<cudaEvent_t> ev_chunk;
    long opt_host_sparse = 1;                // host-buffer call: flag travels back as row-runs, not as a dense cube
    long opt_host_zero_threads = 0;          // threads of the zeroing pass only (0: host_threads / automatic)
    long opt_host_out_zeroed = 0;            // the caller guarantees that flag_host is all zero (fresh calloc / np.zeros pages)
    long opt_host_threads = 0;               // host threads that zero / paint the host flag cube (0 = automatic)
    DevBuf lc_st, lc_t, lc_label, lc_npix, lc_roll, lc_out, lc_bitmaps;    // run_lifecycle scratch
    PinBuf hp_lc;
    long lc_rows = 0;
    PinBuf hp_runs;                          // row-run table of the last host-buffer call
    DevBuf l_parent, l_flag, l_rank, l_label, l_kept, l_accE, l_accS, l_accN;
    DevBuf b_t0, b_t1, b_y0, b_y1, b_x0, b_x1, b_fin;     // label boxes, value per label
    cudaStream_t side_stream = nullptr;      // zero fill (lowest priority)
    cudaStream_t tbl_stream = nullptr;       // table phase (highest priority): must get onto the SMs between fill blocks
    cudaEvent_t ev_tbl[2] = {nullptr, nullptr};
    cudaEvent_t ev_side[2] = {nullptr, nullptr};
    long opt_overlap_zero = 1;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t copy_stream = nullptr, work_stream = nullptr;
    std::map<std::string, double> stats;
    std::vector<double> w_host, thr_cached;
    long nspecial_cached = 0;
    long launches = 0;
    // time-sharded run: packed tables of all ranks -> global tables (ct_global_merge), plane runs served by the caller
    DevBuf x_q, x_qscratch, x_idx, ovf_rows;
    ct_plane_runs_fn fetch_fn = nullptr;
    void* fetch_user = nullptr;
    // ct_shard_begin: thresholding of the own planes is deferred to ct_shard_tables_dev (pipelined with the table kernels)
    long opt_fill_ctas = -1;                  // resident blocks per SM of the zero fill (room for the table kernels beside it); -1 = automatic, 0 = uncapped
    long opt_p2p = 1;                         // sharded run: 1 = tables reach the other ranks through peer windows (NVLink stores), 0 = ncclAllGather
    long opt_fill_late = 0;                   // plane-kernel path: 1 = the zero fill starts after the plane kernel (0: beside it)
    int32_t* pend_fill = nullptr;             // ... the fill ctf::finish() has to start
    size_t pend_fill_cells = 0;
    long opt_profile_tables = 0;              // debug: CUDA-event time of every group of table kernels -> stats "ms_t_*"
    std::vector<std::pair<std::string, cudaEvent_t>> prof;
    size_t prof_used = 0;
    // ---- fast path (ct_fast.cu): per-plane table kernel + cooperative global kernel, no host round trip between them ----
    long opt_plane_kernel = 2;               // tables: 0 = global-memory kernels of ct_kernels.cu, 1 = plane kernel, 2 = by size
    long opt_plane_max_planes = 4096;        // ... plane kernel up to this many planes per context
    long opt_fast_chunks = 1;                // time chunks of the fast path (1: the plane kernel runs beside the zero fill)
    long opt_max_sweeps = 32;                // Jacobi sweeps of step 3 before the plane-ordered wavefront takes over
    long opt_plane_smem = 0;                 // shared-memory budget of the plane kernel in bytes (0 = automatic)
    size_t pl_budget = 0;                    // budget in use (grows from 40 KB to 200 KB when a plane does not fit)
    DevBuf pl_chain, pl_done, pl_ctl;        // look-back descriptors, "plane written" flags, ticket / status / totals / results
    PinBuf hp_ctl, hp_ev, hp_ev2, hp_patch;
    DevBuf g_dirty, g_blocksum, g_evflag, g_ev, g_lrec, g_patch;
    long pl_planes = 0;                      // planes the chain was set up for
    int coop_grid = 0;
    struct D2H { void* dst; const void* src; size_t bytes; } extra_d2h[2] = {{nullptr, nullptr, 0}, {nullptr, nullptr, 0}};
                                             // small device-to-host copies ctf::global() queues behind its kernel
    int plane_timed = 0;
    int fast_tables = 0;                     // the value of every component is in c_val on the device (paint by component)
    // ---- time-sharded run (ct_dist.cu) ----
    ct_ctx* gctx = nullptr;                  // the merged global tables live in a second context on the same device (owned)
    DevBuf sh_export, sh_gathered, sh_mdesc; // packed tables of this rank / of all ranks, per-rank descriptors
    DevBuf sh_lastplane;                     // host-buffer variant: staging of the shard's last plane
    cudaStream_t host_stream = nullptr;      // ... and the stream it runs on
    PinBuf hp_hdr;
    cudaEvent_t ev_x[2] = {nullptr, nullptr};
    cudaEvent_t ev_p[4] = {nullptr, nullptr, nullptr, nullptr};    // plane kernel begin / end, global kernel begin / end
};

namespace cti {
// Resident blocks per SM of the zero fill.  Beside the plane kernel (and the exchange / merge / global kernels of a short
// shard) one block per SM: the fill still finishes before the paint needs it, and the latency-bound kernels beside it run
// almost as if alone (8 GPUs, 1370 planes per rank: 2.60 ms per step against 3.01 ms with two blocks).  Beside the
// global-memory table kernels of a long cube the fill itself is on the critical path: two blocks.
inline int fill_ctas(const ct_ctx* c, bool beside_plane_kernel) {
    return c->opt_fill_ctas >= 0 ? (int)c->opt_fill_ctas : (beside_plane_kernel ? 1 : 2);
}
}  // namespace cti


