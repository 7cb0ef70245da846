// ct_api.cu -- C-ABI of the library (include/contrack_b200.h): context, device scratch, orchestration.
//
// ct_run_contrack = threshold (bits + row-runs in row slots) -> scans -> compact runs -> ccl -> component / pair / date-line
// tables -> step 3 sweeps + 3-D labels + label boxes (all CUDA, ct_kernels.cu) -> label boxes to the host -> date-line
// merge + persistence per label (ct_tables.cpp; exact per-component / all-host replays as fallbacks, ct_host.cpp) -> value
// per label / component / run back on the device -> zero fill + sparse paint.  Reference: contrack/contrack.py:646-791.
// The time-sharded entry points (ct_shard_*, ct_global_*) run the same kernels per rank and merge the rank tables on the
// device (ct_shard.cu); ct_extras.cu / ct_anom.cu / ct_lifecycle.cu hold the callers either side of the path.
#include "../../include/contrack_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "ct_ctx.h"
#include "ct_fast.h"
#include "ct_internal.h"
#include "ct_shard.h"

namespace cta {
cudaError_t group_mean(const void* z, int f64, long HW, int G, const int32_t* gptr_dev, const int32_t* gidx_dev, void* gmean,
                       cudaStream_t st);
cudaError_t clim_smooth(const void* gmean, int f64, long HW, int G, int window, void* clim, cudaStream_t st);
cudaError_t anom(const void* z, int f64, long HW, long T, const int32_t* group_dev, const void* clim, int smooth, void* out,
                 cudaStream_t st);
cudaError_t anom_chunks(const void* z, int f64, long HW, long T, const int32_t* chunk_start_dev, int nchunks,
                        const int32_t* group_dev, const void* clim, int smooth, void* out, cudaStream_t st);
}  // namespace cta

namespace cte {
cudaError_t flag_count(const int32_t* flag, long T, int H, int W, int v, int32_t* count, int sm_count, cudaStream_t st);
cudaError_t divide_f32(const float* in, size_t n, float g, float* out, cudaStream_t st);
cudaError_t gather_planes(const void* src, int f64, int G, int Hs, int Ws, const int32_t* iy_dev, const int32_t* ix_dev, int H,
                          int W, void* dst, cudaStream_t st);
}  // namespace cte

using cti::fail;
using cti::now_ms;

namespace cti {
std::string& last_error() {
    thread_local std::string e;
    return e;
}
}  // namespace cti

namespace {

struct DeviceRunSource : cth::RunSource {
    ct_ctx* c;
    cudaStream_t st;
    bool plane_runs(long t, std::vector<cth::PlaneRun>& out) override {
        out.clear();
        const int H = c->H;
        std::vector<uint32_t> rp(H + 1);
        if (cudaMemcpyAsync(rp.data(), c->row_ptr.as<uint32_t>() + t * H, (H + 1) * sizeof(uint32_t),
                            cudaMemcpyDeviceToHost, st) != cudaSuccess) return false;
        if (cudaStreamSynchronize(st) != cudaSuccess) return false;
        const uint32_t r0 = rp[0], n = rp[H] - rp[0];
        if (n == 0) return true;
        std::vector<uint32_t> rx(n), rc(n);
        if (cudaMemcpyAsync(rx.data(), c->run_x.as<uint32_t>() + r0, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st)
            != cudaSuccess) return false;
        if (cudaMemcpyAsync(rc.data(), c->run_comp.as<uint32_t>() + r0, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st)
            != cudaSuccess) return false;
        if (cudaStreamSynchronize(st) != cudaSuccess) return false;
        out.resize(n);
        int y = 0;
        for (uint32_t i = 0; i < n; ++i) {
            while (rp[y + 1] - r0 <= i) ++y;
            out[i] = cth::PlaneRun{y, (int)(rx[i] & 0xffff), (int)(rx[i] >> 16), rc[i]};
        }
        return true;
    }
};

// Plane runs served by the caller (time-sharded run: the plane may live on another rank); component ids are global.
struct CallbackRunSource : cth::RunSource {
    ct_plane_runs_fn fn; void* user;
    bool plane_runs(long t, std::vector<cth::PlaneRun>& out) override {
        out.clear();
        long n = 0;
        const int32_t *y = nullptr, *x0 = nullptr, *x1 = nullptr;
        const uint32_t* comp = nullptr;
        if (!fn || fn(user, t, &n, &y, &x0, &x1, &comp) != 0) return false;
        out.resize((size_t)n);
        for (long i = 0; i < n; ++i) out[i] = cth::PlaneRun{y[i], x0[i], x1[i], comp[i]};
        return true;
    }
};

// Rows whose weights can be summed exactly in float64 in ANY order (so atomics give numpy's value bit for bit):
// all partial sums are multiples of 2^lb_min and bounded by W * sum|w| < 2^(53 + lb_min).  Rows with the lowest
// set bits (the pole rows, cos(float32(pi/2)) ~ -4.4e-8) are taken out until that holds; their pixels are
// accumulated separately and any decision that involves them is guarded (ct_host.cpp).
void classify_rows(const double* w, int H, int W, std::vector<uint8_t>& special) {
    special.assign(H, 0);
    std::vector<int> lb(H, INT32_MAX);
    std::vector<int> order;
    for (int y = 0; y < H; ++y) {
        const double a = std::fabs(w[y]);
        if (a == 0.0) continue;
        if (!std::isfinite(a)) { special[y] = 1; continue; }
        int e;
        const double m = std::frexp(a, &e);                           // a = m * 2^e, m in [0.5, 1)
        unsigned long long mi = (unsigned long long)std::ldexp(m, 53);
        lb[y] = e - 53 + __builtin_ctzll(mi);
        order.push_back(y);
    }
    std::sort(order.begin(), order.end(), [&](int a, int b) { return lb[a] < lb[b]; });
    size_t first = 0;                                                 // order[first..] is the exact set
    while (first < order.size()) {
        double bound = 0.0;
        for (size_t i = first; i < order.size(); ++i) bound += std::fabs(w[order[i]]) * (double)W;
        const int lbmin = lb[order[first]];
        if (bound < std::ldexp(1.0, 53 + lbmin)) break;
        special[order[first]] = 1;
        ++first;
    }
}

// 1 if every special row carries the same weight (e.g. the two pole rows of a regular grid)
int special_rows_uniform(const double* w, const std::vector<uint8_t>& special) {
    bool have = false;
    double v = 0.0;
    for (size_t y = 0; y < special.size(); ++y) {
        if (!special[y]) continue;
        if (!have) { v = w[y]; have = true; }
        else if (memcmp(&v, &w[y], sizeof v) != 0) return 0;
    }
    return 1;
}

int special_rows_uniform(const double* w, int H, int W) {
    std::vector<uint8_t> sp;
    classify_rows(w, H, W, sp);
    return special_rows_uniform(w, sp);
}

uint32_t next_pow2(uint64_t v) {
    uint64_t p = 1024;
    while (p < v && p < (1ull << 31)) p <<= 1;
    return (uint32_t)p;
}


int check_args(long T, int H, int W, const double* w_host, const double* thr_host, long thr_n, int in_dtype, int op) {
    if (T < 0 || H <= 0 || W <= 0) return fail(CT_ERR_ARG, "bad shape T=%ld H=%d W=%d", T, H, W);
    if (H > 65535 || W > 65535) return fail(CT_ERR_CAPACITY, "H and W must be <= 65535 (got %d x %d)", H, W);
    if ((double)T * H >= 2147483647.0) return fail(CT_ERR_CAPACITY, "T*H must be < 2^31");
    if (!w_host || !thr_host) return fail(CT_ERR_ARG, "null weight / threshold pointer");
    if (thr_n != 1 && thr_n != T) return fail(CT_ERR_ARG, "thr_n must be 1 or T");
    if (in_dtype != CT_F32 && in_dtype != CT_F64) return fail(CT_ERR_ARG, "in_dtype must be CT_F32 or CT_F64");
    if (op < CT_GE || op > CT_LT) return fail(CT_ERR_ARG, " Please select from [>, >=, <, >=] for gorl");
    return CT_OK;
}

// upload weights / special-row map / thresholds; size the row-indexed scratch
int prepare(ct_ctx* c, long T, int H, int W, const double* w_host, const double* thr_host, long thr_n,
            cudaStream_t st) {
    // weights / special-row map / thresholds of the previous call are still on the device: a repeated call with the same
    // grid and thresholds (time stepping, benchmarks) skips the classification, three uploads and a synchronize
    const bool same = c->H == H && c->W == W && (long)c->w_host.size() == H && (long)c->thr_cached.size() == thr_n &&
                      c->w_dev.p && c->special_dev.p && c->thr_dev.p &&
                      memcmp(c->w_host.data(), w_host, (size_t)H * 8) == 0 &&
                      memcmp(c->thr_cached.data(), thr_host, (size_t)thr_n * 8) == 0;
    c->T = T; c->H = H; c->W = W; c->Ww = (W + 31) / 32;
    const long nrows = T * H;
    if (same) {
        c->stats["special_rows"] = (double)c->nspecial_cached;
    } else {
        c->w_host.assign(w_host, w_host + H);
        c->thr_cached.assign(thr_host, thr_host + thr_n);
        std::vector<uint8_t> special;
        classify_rows(w_host, H, W, special);
        long nspecial = 0;
        for (uint8_t s : special) nspecial += s;
        c->nspecial_cached = nspecial;
        c->stats["special_rows"] = (double)nspecial;
        c->special_uniform = special_rows_uniform(w_host, special);
        CT_CUDA(c->w_dev.ensure(H * sizeof(double)));
        CT_CUDA(c->special_dev.ensure(H));
        CT_CUDA(c->thr_dev.ensure(thr_n * sizeof(double)));
        CT_CUDA(cudaMemcpyAsync(c->w_dev.p, w_host, H * sizeof(double), cudaMemcpyHostToDevice, st));
        CT_CUDA(cudaMemcpyAsync(c->special_dev.p, special.data(), H, cudaMemcpyHostToDevice, st));
        CT_CUDA(cudaMemcpyAsync(c->thr_dev.p, thr_host, thr_n * sizeof(double), cudaMemcpyHostToDevice, st));
        CT_CUDA(cudaStreamSynchronize(st));                           // `special` is a local
    }
    CT_CUDA(c->bits.ensure((size_t)nrows * c->Ww * sizeof(uint32_t)));
    CT_CUDA(c->row_cnt.ensure((size_t)nrows * sizeof(uint32_t)));
    CT_CUDA(c->seam_flag.ensure((size_t)nrows * sizeof(uint32_t)));
    CT_CUDA(c->row_ptr.ensure((size_t)(nrows + 1) * sizeof(uint32_t)));
    CT_CUDA(c->seam_pos.ensure((size_t)(nrows + 1) * sizeof(uint32_t)));
    CT_CUDA(c->slots.ensure((size_t)nrows * ctk::RUN_SLOTS_PER_ROW * sizeof(uint32_t)));
    CT_CUDA(c->counters.ensure(128));
    CT_CUDA(c->hp_counters.ensure(128));
    CT_CUDA(cudaMemsetAsync(c->counters.as<uint32_t>() + 16, 0, 12, st));           // 16: "a row has more runs than slots",
                                                                                    // 17: overflow-row count, 18: run index wrapped
    for (auto& e : c->ev) if (!e) CT_CUDA(cudaEventCreate(&e));
    c->launches = 0;
    return CT_OK;
}

int launch_threshold(ct_ctx* c, const void* anom_dev, int in_dtype, long t0, long nt, long thr_n, int thr_is_f32,
                     int op, cudaStream_t st, int all_bits = 1) {
    ctk::ThresholdArgs a;
    // the bit rows are read by: the dense paint, the boundary export of a shard and the host-buffer call's dense fallback --
    // callers that need none of these pass all_bits = 0 and only rows with more runs than slots get their bit row
    a.bits_overflow_only = (!all_bits && c->opt_overlap_zero) ? 1 : 0;
    a.anom = anom_dev; a.in_dtype = in_dtype; a.T = nt; a.H = c->H; a.W = c->W; a.Ww = c->Ww;
    a.thr_dev = thr_n == 1 ? c->thr_dev.as<double>() : c->thr_dev.as<double>() + t0;
    a.thr_n = thr_n; a.thr_is_f32 = thr_is_f32; a.op = op;
    const long r0 = t0 * c->H;
    a.bits = c->bits.as<uint32_t>() + (size_t)r0 * c->Ww;
    a.row_cnt = c->row_cnt.as<uint32_t>() + r0;
    a.seam_flag = c->seam_flag.as<uint32_t>() + r0;
    a.slots = c->slots.as<uint32_t>() + (size_t)r0 * ctk::RUN_SLOTS_PER_ROW;
    a.overflow = c->counters.as<uint32_t>() + 16;
    a.variant = (int)c->opt_tma;
    CT_CUDA(ctk::threshold_bits(a, c->sm_count, st));
    c->launches += 1;
    return CT_OK;
}

// GPU half of the table phase: bit rows -> runs -> components -> tables (left on the device), for the planes [p0, p1).
// Chunks must be processed in time order; every table is indexed globally (rows, runs, components, pairs, date-line rows
// and segments continue where the previous chunk stopped), so the result is identical to one pass over all planes.
// Three host round trips per chunk (run count, component count, pair count) size the tables.
// debug profiling of the table kernels (option "profile_tables"): an event after every kernel group
void prof_mark(ct_ctx* c, const char* name, cudaStream_t st) {
    if (!c->opt_profile_tables) return;
    if (c->prof_used == c->prof.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        c->prof.emplace_back(std::string(), e);
    }
    c->prof[c->prof_used].first = name;
    cudaEventRecord(c->prof[c->prof_used].second, st);
    c->prof_used++;
}
void prof_collect(ct_ctx* c) {                // after a synchronize: elapsed time between successive marks, summed by name
    for (size_t i = 1; i < c->prof_used; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->prof[i - 1].second, c->prof[i].second) == cudaSuccess)
            c->stats["ms_t_" + c->prof[i].first] += ms;
    }
    c->prof_used = 0;
}

void tables_begin(ct_ctx* c) {
    c->tb_planes = c->tb_runs = c->tb_comps = c->tb_seams = c->tb_segs = c->tb_pairs = 0;
}

int tables_chunk(ct_ctx* c, long p0, long p1, cudaStream_t st) {
    const int H = c->H, W = c->W;
    const long r0 = p0 * H, r1 = p1 * H, n = r1 - r0;
    if (p0 != c->tb_planes || p1 <= p0 || p1 > c->T) return fail(CT_ERR_INTERNAL, "table chunks out of order");
    uint32_t* cnt_dev = c->counters.as<uint32_t>();
    uint32_t* cnt_host = c->hp_counters.as<uint32_t>();
    auto U = [](DevBuf& b) { return b.as<uint32_t>(); };
    // the first chunk of several sizes the tables for the whole cube (its density x all planes + 25 %)
    const double scale = p0 == 0 && p1 < c->T ? 1.25 * (double)c->T / (double)(p1 - p0) : 0.0;
    auto hint = [&](long elems, size_t elt) { return (size_t)((double)elems * scale) * elt; };

    CT_CUDA(c->scan_tmp.ensure(ctk::scan_tmp_elems(n > 1024 ? n : 1024) * sizeof(uint32_t)));
    prof_mark(c, "start", st);
    // counter 18: "the run total of the cube no longer fits the 32-bit run index" (dense / noisy masks on large cubes)
    CT_CUDA(ctk::exclusive_scan_u32(U(c->row_cnt) + r0, U(c->row_ptr) + r0, n, U(c->scan_tmp), st, (uint32_t)c->tb_runs,
                                    cnt_dev + 18));
    CT_CUDA(ctk::exclusive_scan_u32(U(c->seam_flag) + r0, U(c->seam_pos) + r0, n, U(c->scan_tmp), st,
                                    (uint32_t)c->tb_seams));
    c->launches += 6;
    CT_CUDA(cudaMemcpyAsync(cnt_host + 0, U(c->row_ptr) + r1, 4, cudaMemcpyDeviceToHost, st));
    CT_CUDA(cudaMemcpyAsync(cnt_host + 1, U(c->seam_pos) + r1, 4, cudaMemcpyDeviceToHost, st));
    CT_CUDA(cudaMemcpyAsync(cnt_host + 16, cnt_dev + 16, 12, cudaMemcpyDeviceToHost, st));
    prof_mark(c, "row_scans", st);
    CT_CUDA(cudaStreamSynchronize(st));
    if (cnt_host[18] || cnt_host[0] >= 0xfffffff0u)
        return fail(CT_ERR_CAPACITY, "more than 2^32 - 16 row-runs in the cube: the run tables are indexed with 32 bits");
    const long Rb = c->tb_runs, Re = cnt_host[0], Sb = c->tb_seams, Se = cnt_host[1];

    // ---- runs + 2-D components ----
    {
        DevBuf* rb[] = {&c->run_x, &c->run_row, &c->parent, &c->root_flag, &c->rank, &c->run_comp};
        for (DevBuf* b : rb) CT_CUDA(b->grow((size_t)(Re + 2) * 4, (size_t)Rb * 4, st, hint(Re + 2, 4)));
    }
    CT_CUDA(c->scan_tmp.ensure(ctk::scan_tmp_elems(std::max(Re - Rb, n)) * sizeof(uint32_t)));
    {
        uint32_t* ovf = nullptr;
        if (cnt_host[16]) {                                      // some row of the cube has more runs than slots
            CT_CUDA(c->ovf_rows.ensure((size_t)n * 4));
            ovf = U(c->ovf_rows);
            c->launches += 1; c->stats["slot_overflow"] = 1.0;
        }
        CT_CUDA(ctk::compact_runs(U(c->slots), U(c->bits), U(c->row_ptr), r0, n, c->Ww, ovf, cnt_dev + 17, U(c->run_x),
                                  U(c->run_row), st));
    }
    prof_mark(c, "extract_runs", st);
    CT_CUDA(ctk::ccl_init(U(c->parent), Rb, Re, st));
    CT_CUDA(ctk::ccl_union(U(c->row_ptr), U(c->run_x), U(c->run_row), Rb, Re, H, U(c->parent), st));
    prof_mark(c, "ccl_union", st);
    CT_CUDA(ctk::ccl_flatten(U(c->parent), U(c->root_flag), Rb, Re, st));
    prof_mark(c, "ccl_flatten", st);
    CT_CUDA(ctk::exclusive_scan_u32(U(c->root_flag) + Rb, U(c->rank) + Rb, Re - Rb, U(c->scan_tmp), st,
                                    (uint32_t)c->tb_comps));
    c->launches += 7;
    CT_CUDA(cudaMemcpyAsync(cnt_host + 2, U(c->rank) + Re, 4, cudaMemcpyDeviceToHost, st));
    prof_mark(c, "root_scan", st);
    CT_CUDA(cudaStreamSynchronize(st));
    const long Cb = c->tb_comps, Ce = cnt_host[2];

    // ---- component tables, date-line rows and classes ----
    CT_CUDA(ctk::ccl_assign(U(c->parent), U(c->rank), U(c->run_comp), Rb, Re, st));
    {
        DevBuf* b4[] = {&c->c_t, &c->c_y0, &c->c_y1, &c->c_x0, &c->c_x1, &c->c_nsp, &c->c_cls, &c->c_val,
                        &c->k_nsp, &c->k_fnsp, &c->pcnt, &c->pfill, &c->pptr};
        DevBuf* b8[] = {&c->c_E, &c->c_S, &c->k_conE, &c->k_conS, &c->k_fE, &c->k_fS};
        for (DevBuf* b : b4) CT_CUDA(b->grow((size_t)(Ce + 3) * 4, (size_t)Cb * 4, st, hint(Ce + 3, 4)));
        for (DevBuf* b : b8) CT_CUDA(b->grow((size_t)(Ce + 3) * 8, (size_t)Cb * 8, st, hint(Ce + 3, 8)));
    }
    ctk::CompTables ct;
    ct.t = c->c_t.as<int32_t>(); ct.y0 = c->c_y0.as<int32_t>(); ct.y1 = c->c_y1.as<int32_t>();
    ct.x0 = c->c_x0.as<int32_t>(); ct.x1 = c->c_x1.as<int32_t>(); ct.areaE = c->c_E.as<double>();
    ct.areaS = c->c_S.as<double>(); ct.nsp = U(c->c_nsp); ct.cls = U(c->c_cls);
    ctk::ClassTables kt;
    kt.conE = c->k_conE.as<double>(); kt.conS = c->k_conS.as<double>(); kt.fE = c->k_fE.as<double>();
    kt.fS = c->k_fS.as<double>(); kt.nsp = U(c->k_nsp); kt.fnsp = U(c->k_fnsp);
    CT_CUDA(ctk::comp_init(ct, Cb, Ce, W, st));
    CT_CUDA(ctk::comp_accumulate(U(c->run_x), U(c->run_row), U(c->run_comp), Rb, Re, H, c->w_dev.as<double>(),
                                 c->special_dev.as<uint8_t>(), ct, st));
    prof_mark(c, "comp_accumulate", st);
    {
        DevBuf* sb[] = {&c->s_row, &c->s_a, &c->s_b, &c->seg_start, &c->seg_pos, &c->g_t, &c->g_y0, &c->g_y1, &c->g_a,
                        &c->g_b};
        for (DevBuf* b : sb) CT_CUDA(b->grow((size_t)(Se + 3) * 4, (size_t)Sb * 4, st, hint(Se + 3, 4)));
    }
    CT_CUDA(ctk::seam_rows(U(c->seam_flag), U(c->seam_pos), U(c->row_ptr), U(c->run_comp), r0, r1, U(c->s_row),
                           U(c->s_a), U(c->s_b), ct.cls, st));
    CT_CUDA(ctk::cls_flatten(ct.cls, Cb, Ce, st));
    ctk::SegTables sg;
    sg.t = c->g_t.as<int32_t>(); sg.y0 = c->g_y0.as<int32_t>(); sg.y1 = c->g_y1.as<int32_t>();
    sg.a = U(c->g_a); sg.b = U(c->g_b);
    CT_CUDA(c->scan_tmp.ensure(ctk::scan_tmp_elems(std::max(std::max(Re - Rb, n), std::max(Ce - Cb, Se - Sb)) + 1) * 4));
    CT_CUDA(ctk::seg_flags(U(c->s_row), U(c->s_a), U(c->s_b), Sb, Se, H, U(c->seg_start), st));
    CT_CUDA(ctk::exclusive_scan_u32(U(c->seg_start) + Sb, U(c->seg_pos) + Sb, Se - Sb, U(c->scan_tmp), st,
                                    (uint32_t)c->tb_segs));
    CT_CUDA(ctk::seg_write(U(c->s_row), U(c->s_a), U(c->s_b), U(c->seg_start), U(c->seg_pos), Sb, Se, H, sg, st));
    CT_CUDA(cudaMemcpyAsync(cnt_host + 3, U(c->seg_pos) + Se, 4, cudaMemcpyDeviceToHost, st));
    prof_mark(c, "seam_segs", st);
    c->launches += 10;

    // ---- adjacent-plane pairs of this chunk's planes (the earlier plane may belong to the previous chunk): hash
    // accumulate -> CSR over the plane-t component; class sums ----
    const size_t o8 = (size_t)Cb * 8, o4 = (size_t)Cb * 4, n8 = (size_t)(Ce - Cb) * 8, n4 = (size_t)(Ce - Cb) * 4;
    if (Ce > Cb) {
        CT_CUDA(cudaMemsetAsync((char*)c->k_conE.p + o8, 0, n8, st)); CT_CUDA(cudaMemsetAsync((char*)c->k_conS.p + o8, 0, n8, st));
        CT_CUDA(cudaMemsetAsync((char*)c->k_fE.p + o8, 0, n8, st)); CT_CUDA(cudaMemsetAsync((char*)c->k_fS.p + o8, 0, n8, st));
        CT_CUDA(cudaMemsetAsync((char*)c->k_nsp.p + o4, 0, n4, st)); CT_CUDA(cudaMemsetAsync((char*)c->k_fnsp.p + o4, 0, n4, st));
        CT_CUDA(cudaMemsetAsync((char*)c->pcnt.p + o4, 0, n4, st)); CT_CUDA(cudaMemsetAsync((char*)c->pfill.p + o4, 0, n4, st));
    }
    CT_CUDA(ctk::class_sums(ct, kt, Cb, Ce, st));
    prof_mark(c, "class_sums", st);
    ctk::PairTable pt;
    uint64_t want = (uint64_t)(Ce - Cb) * 4;
    for (int attempt = 0;; ++attempt) {
        pt.cap = next_pow2(want);
        CT_CUDA(c->h_key.ensure((size_t)pt.cap * 8)); CT_CUDA(c->h_npix.ensure((size_t)pt.cap * 4));
        CT_CUDA(c->h_nsp.ensure((size_t)pt.cap * 4)); CT_CUDA(c->h_E.ensure((size_t)pt.cap * 8));
        CT_CUDA(c->h_S.ensure((size_t)pt.cap * 8));
        pt.key = c->h_key.as<unsigned long long>(); pt.npix = U(c->h_npix);
        pt.nsp = U(c->h_nsp); pt.areaE = c->h_E.as<double>(); pt.areaS = c->h_S.as<double>();
        pt.overflow = cnt_dev + 4; pt.count = cnt_dev + 5;
        CT_CUDA(ctk::pairs_init(pt, st));
        CT_CUDA(ctk::pairs_accumulate(U(c->row_ptr), U(c->run_x), U(c->run_row), U(c->run_comp), Rb, Re, H,
                                      c->w_dev.as<double>(), c->special_dev.as<uint8_t>(), pt, st));
        c->launches += 3;
        prof_mark(c, "pairs_accumulate", st);
        CT_CUDA(cudaMemcpyAsync(cnt_host + 4, cnt_dev + 4, 8, cudaMemcpyDeviceToHost, st));
        CT_CUDA(cudaStreamSynchronize(st));
        if (cnt_host[4] == 0 && (uint64_t)cnt_host[5] * 10 <= (uint64_t)pt.cap * 7) break;
        if (attempt >= 6 || pt.cap >= (1u << 31)) return fail(CT_ERR_CAPACITY, "pair table overflow");
        want = (uint64_t)pt.cap * 4;                                  // too full: probing would crawl
    }
    const long Pb = c->tb_pairs, Pe = Pb + cnt_host[5];
    {
        DevBuf* b4[] = {&c->p_b, &c->p_npix, &c->p_nsp};
        DevBuf* b8[] = {&c->p_E, &c->p_S};
        for (DevBuf* b : b4) CT_CUDA(b->grow((size_t)(Pe + 2) * 4, (size_t)Pb * 4, st, hint(Pe + 2, 4)));
        for (DevBuf* b : b8) CT_CUDA(b->grow((size_t)(Pe + 2) * 8, (size_t)Pb * 8, st, hint(Pe + 2, 8)));
    }
    ctk::PairCsr q;
    q.b = U(c->p_b); q.npix = U(c->p_npix); q.nsp = U(c->p_nsp); q.E = c->p_E.as<double>(); q.S = c->p_S.as<double>();
    CT_CUDA(ctk::pairs_count(pt, ct.cls, kt, U(c->pcnt), st));
    CT_CUDA(ctk::exclusive_scan_u32(U(c->pcnt) + Cb, U(c->pptr) + Cb, Ce - Cb, U(c->scan_tmp), st, (uint32_t)Pb));
    CT_CUDA(ctk::pairs_fill(pt, U(c->pptr), U(c->pfill), q, st));
    prof_mark(c, "pairs_csr", st);
    if (c->opt_profile_tables) { CT_CUDA(cudaStreamSynchronize(st)); prof_collect(c); }
    c->launches += 5;
    c->tb_planes = p1; c->tb_runs = Re; c->tb_comps = Ce; c->tb_seams = Se; c->tb_pairs = Pe;
    c->tb_segs = Se > Sb ? cnt_host[3] : c->tb_segs;
    return CT_OK;
}

int tables_finish(ct_ctx* c, cudaStream_t st) {
    if (c->tb_planes != c->T) return fail(CT_ERR_INTERNAL, "tables cover %ld of %ld planes", c->tb_planes, c->T);
    c->nruns = c->tb_runs; c->ncomp = c->tb_comps; c->npair = c->tb_pairs; c->nseam = c->tb_seams; c->nseg = c->tb_segs;
    CT_CUDA(c->run_val.ensure((size_t)(c->nruns + 1) * 4));
    (void)st;
    c->stats["runs"] = (double)c->nruns; c->stats["comps2d"] = (double)c->ncomp; c->stats["pairs"] = (double)c->npair;
    c->stats["seam_rows"] = (double)c->nseam; c->stats["seam_segments"] = (double)c->nseg;
    return CT_OK;
}

int tables_build(ct_ctx* c, cudaStream_t st) {
    tables_begin(c);
    if (c->T > 0) {
        int rc = tables_chunk(c, 0, c->T, st);
        if (rc != CT_OK) return rc;
    }
    return tables_finish(c, st);
}

// Tables -> pinned host memory (c->host_tb).  full = 0 copies only what steps 4c/4d need (component boxes, classes,
// date-line segments); full = 1 adds the class sums and the pair CSR for the ordered host phase / the sharded gather.
int tables_d2h(ct_ctx* c, int full, cudaStream_t st) {
    const long nc = c->ncomp, np = full ? c->npair : 0, nseg = c->nseg;
    const int H = c->H, W = c->W;
    // ---- tables -> pinned host memory (8-byte arrays first) ----
    const size_t ncp = (size_t)nc + 2, npp = (size_t)np + 2, ngp = (size_t)nseg + 2;
    const size_t bytes = ncp * (4 * 8 + 10 * 4) + npp * (2 * 8 + 3 * 4) + ngp * 5 * 4 + 512;
    CT_CUDA(c->hp_tables.ensure(bytes));
    char* base = c->hp_tables.as<char>();
    size_t off = 0;
    auto take = [&](size_t n, size_t elt) { void* p = base + off; off += ((n * elt + 15) / 16) * 16; return p; };
    double* h_conE = (double*)take(ncp, 8); double* h_conS = (double*)take(ncp, 8);
    double* h_fE = (double*)take(ncp, 8); double* h_fS = (double*)take(ncp, 8);
    double* h_pE = (double*)take(npp, 8); double* h_pS = (double*)take(npp, 8);
    int32_t* h_t = (int32_t*)take(ncp, 4); int32_t* h_y0 = (int32_t*)take(ncp, 4); int32_t* h_y1 = (int32_t*)take(ncp, 4);
    int32_t* h_x0 = (int32_t*)take(ncp, 4); int32_t* h_x1 = (int32_t*)take(ncp, 4);
    uint32_t* h_cls = (uint32_t*)take(ncp, 4); uint32_t* h_knsp = (uint32_t*)take(ncp, 4);
    uint32_t* h_kfnsp = (uint32_t*)take(ncp, 4);
    uint32_t* h_pptr = (uint32_t*)take(ncp, 4);
    uint32_t* h_pb = (uint32_t*)take(npp, 4); uint32_t* h_pn = (uint32_t*)take(npp, 4);
    uint32_t* h_pnsp = (uint32_t*)take(npp, 4);
    int32_t* h_gt = (int32_t*)take(ngp, 4); int32_t* h_gy0 = (int32_t*)take(ngp, 4); int32_t* h_gy1 = (int32_t*)take(ngp, 4);
    uint32_t* h_ga = (uint32_t*)take(ngp, 4); uint32_t* h_gb = (uint32_t*)take(ngp, 4);
    if (off > c->hp_tables.cap) return fail(CT_ERR_INTERNAL, "staging layout overflow");
#define CT_D2H(dst, src, n, elt) \
    if ((n) > 0) CT_CUDA(cudaMemcpyAsync(dst, (src).p, (size_t)(n) * (elt), cudaMemcpyDeviceToHost, st))
    if (full) {
        CT_D2H(h_conE, c->k_conE, nc, 8); CT_D2H(h_conS, c->k_conS, nc, 8); CT_D2H(h_fE, c->k_fE, nc, 8);
        CT_D2H(h_fS, c->k_fS, nc, 8); CT_D2H(h_knsp, c->k_nsp, nc, 4); CT_D2H(h_kfnsp, c->k_fnsp, nc, 4);
        CT_D2H(h_pptr, c->pptr, nc + 1, 4);
    }
    CT_D2H(h_t, c->c_t, nc, 4); CT_D2H(h_y0, c->c_y0, nc, 4); CT_D2H(h_y1, c->c_y1, nc, 4);
    CT_D2H(h_x0, c->c_x0, nc, 4); CT_D2H(h_x1, c->c_x1, nc, 4); CT_D2H(h_cls, c->c_cls, nc, 4);
    CT_D2H(h_pE, c->p_E, np, 8); CT_D2H(h_pS, c->p_S, np, 8); CT_D2H(h_pb, c->p_b, np, 4);
    CT_D2H(h_pn, c->p_npix, np, 4); CT_D2H(h_pnsp, c->p_nsp, np, 4);
    CT_D2H(h_gt, c->g_t, nseg, 4); CT_D2H(h_gy0, c->g_y0, nseg, 4); CT_D2H(h_gy1, c->g_y1, nseg, 4);
    CT_D2H(h_ga, c->g_a, nseg, 4); CT_D2H(h_gb, c->g_b, nseg, 4);
#undef CT_D2H
    CT_CUDA(cudaStreamSynchronize(st));
    if (nc == 0) h_pptr[0] = 0;

    const char* dump = full ? getenv("CT_DUMP_TABLES") : nullptr;
    if (dump) {              // debugging aid: tables of this run as raw binary
        if (FILE* f = fopen(dump, "wb")) {
            long hdr[8] = {c->T, (long)H, (long)W, nc, np, nseg, 1, 0};
            fwrite(hdr, sizeof(long), 8, f);
            fwrite(c->w_host.data(), 8, H, f);
            fwrite(h_t, 4, nc, f); fwrite(h_y0, 4, nc, f); fwrite(h_y1, 4, nc, f); fwrite(h_x0, 4, nc, f);
            fwrite(h_x1, 4, nc, f); fwrite(h_cls, 4, nc, f); fwrite(h_conE, 8, nc, f); fwrite(h_conS, 8, nc, f);
            for (long i = 0; i < nc; ++i) h_knsp[i] += h_kfnsp[i];   // dump format: one counter
            fwrite(h_fE, 8, nc, f); fwrite(h_fS, 8, nc, f); fwrite(h_knsp, 4, nc, f); fwrite(h_pptr, 4, nc + 1, f);
            for (long i = 0; i < nc; ++i) h_knsp[i] -= h_kfnsp[i];
            fwrite(h_pb, 4, np, f); fwrite(h_pn, 4, np, f); fwrite(h_pnsp, 4, np, f);
            fwrite(h_pE, 8, np, f); fwrite(h_pS, 8, np, f);
            fwrite(h_gt, 4, nseg, f); fwrite(h_gy0, 4, nseg, f); fwrite(h_gy1, 4, nseg, f); fwrite(h_ga, 4, nseg, f);
            fwrite(h_gb, 4, nseg, f);
            fclose(f);
        }
    }

    cth::FastTables& tb = c->host_tb;
    tb = cth::FastTables();
    tb.T = c->T; tb.H = H; tb.W = W; tb.ncomp = nc;
    tb.comp_t = h_t; tb.comp_y0 = h_y0; tb.comp_y1 = h_y1; tb.comp_x0 = h_x0; tb.comp_x1 = h_x1; tb.comp_cls = h_cls;
    tb.cls_conE = h_conE; tb.cls_conS = h_conS; tb.cls_fE = h_fE; tb.cls_fS = h_fS; tb.cls_nsp = h_knsp;
    tb.cls_fnsp = h_kfnsp;
    tb.pair_ptr = h_pptr; tb.pair_b = h_pb; tb.pair_npix = h_pn; tb.pair_nsp = h_pnsp; tb.pair_E = h_pE; tb.pair_S = h_pS;
    tb.nseg = nseg; tb.seg_t = h_gt; tb.seg_y0 = h_gy0; tb.seg_y1 = h_gy1; tb.seg_a = h_ga; tb.seg_b = h_gb;
    tb.w = c->w_host.data();
    tb.special_uniform = c->special_uniform;
    return CT_OK;
}

// Upload the value of every component (and the override sub-runs) and resolve them to a value per row-run.
int upload_values(ct_ctx* c, const int32_t* comp_val_pinned, const std::vector<ctb::Override>& overrides, cudaStream_t st) {
    const long nc = c->ncomp, R = c->nruns;
    const long novr = (long)overrides.size();
    c->novr = novr;
    if (nc && comp_val_pinned)                                   // nullptr: c_val was already computed on the device
        CT_CUDA(cudaMemcpyAsync(c->c_val.p, comp_val_pinned, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
    if (novr) {
        CT_CUDA(c->hp_ovr.ensure((size_t)novr * 5 * 4));
        int32_t* ho = c->hp_ovr.as<int32_t>();
        for (long i = 0; i < novr; ++i) {
            const ctb::Override& o = overrides[i];
            ho[i] = o.t; ho[novr + i] = o.y; ho[2 * novr + i] = o.x0; ho[3 * novr + i] = o.x1; ho[4 * novr + i] = o.val;
        }
        const size_t ob = (size_t)novr * 4;
        CT_CUDA(c->o_t.ensure(ob)); CT_CUDA(c->o_y.ensure(ob)); CT_CUDA(c->o_x0.ensure(ob));
        CT_CUDA(c->o_x1.ensure(ob)); CT_CUDA(c->o_val.ensure(ob));
        CT_CUDA(cudaMemcpyAsync(c->o_t.p, ho, ob, cudaMemcpyHostToDevice, st));
        CT_CUDA(cudaMemcpyAsync(c->o_y.p, ho + novr, ob, cudaMemcpyHostToDevice, st));
        CT_CUDA(cudaMemcpyAsync(c->o_x0.p, ho + 2 * novr, ob, cudaMemcpyHostToDevice, st));
        CT_CUDA(cudaMemcpyAsync(c->o_x1.p, ho + 3 * novr, ob, cudaMemcpyHostToDevice, st));
        CT_CUDA(cudaMemcpyAsync(c->o_val.p, ho + 4 * novr, ob, cudaMemcpyHostToDevice, st));
    }
    CT_CUDA(ctk::run_values(c->run_comp.as<uint32_t>(), c->c_val.as<int32_t>(), c->run_val.as<int32_t>(), R, st));
    c->launches += 1;
    c->stats["override_runs"] = (double)novr;
    return CT_OK;
}

// The ordered phase on the HOST (ct_host.cpp: the reference's loops replayed on the tables, exact near-tie resolver, stale-box
// splits at run granularity): what the device path (cooperative global kernel + event replay, ct_fast.cu) falls back to when
// a verdict sits within rounding distance of `overlap` on rows that do not sum exactly or a label straddles a stale box, and
// what the debug stages and "gpu_tables" = 0 use.  On return the value per row-run and the override sub-runs are on the device.
int table_phase(ct_ctx* c, double overlap, int persistence, int twosided, int stage, long* n_features, cudaStream_t st,
                bool tables_built = false) {
    int rc0 = tables_built ? CT_OK : tables_build(c, st);
    if (rc0 != CT_OK) return rc0;
    const long nc = c->ncomp;
    cth::FastTables& tb = c->host_tb;
    CT_CUDA(cudaEventRecord(c->ev[2], st));
    if ((rc0 = tables_d2h(c, 1, st)) != CT_OK) return rc0;

    // ---- ordered table phase on the host ----
    const double t_host0 = now_ms();
    cth::Params pr;
    pr.overlap = overlap; pr.persistence = persistence; pr.twosided = twosided; pr.stage = stage;
    DeviceRunSource dsrc;
    dsrc.c = c; dsrc.st = st;
    CallbackRunSource csrc;
    csrc.fn = c->fetch_fn; csrc.user = c->fetch_user;
    cth::RunSource* src = c->fetch_fn ? static_cast<cth::RunSource*>(&csrc) : static_cast<cth::RunSource*>(&dsrc);
    CT_CUDA(c->hp_val.ensure((size_t)(nc + 1) * 4));
    int32_t* hv = c->hp_val.as<int32_t>();
    cth::Result& res = c->host_result;
    std::string err;
    int rc = cth::host_phase_fast(tb, pr, src, hv, res, err);
    if (rc != 0) return fail(rc, "%s", err.c_str());
    c->stats["ms_host_tables"] = now_ms() - t_host0;

    // ---- values back to the device ----
    if ((rc = upload_values(c, hv, res.overrides, st)) != CT_OK) return rc;

    c->stats["kept_comps"] = (double)res.n_kept;
    c->stats["labels3d"] = (double)res.n_labels3d; c->stats["features"] = (double)res.n_features;
    c->stats["seam_events"] = (double)res.n_seam_events; c->stats["seam_splits"] = (double)res.n_seam_splits;
    c->stats["neartie_resolved"] = (double)res.n_neartie;
    if (n_features) *n_features = res.n_features;
    return CT_OK;
}


// Which table builder for T planes: the plane kernel (one launch, no host round trip; per-plane work is latency-bound, so its
// throughput is ~T / 740 x 0.2 ms) for small T, e.g. the shard of a multi-GPU run; the global-memory table kernels (every
// step fully parallel over all planes, pipelined under the threshold kernel) for long cubes on one GPU.
bool use_plane_kernel(const ct_ctx* c, long T) {
    if (!c->opt_gpu_tables) return false;
    return c->opt_plane_kernel == 1 || (c->opt_plane_kernel == 2 && T <= c->opt_plane_max_planes);
}

// Tables of the planes [0, T) of the context's scratch and the ordered phase on them.  `wait_chunk(k)` makes `ts` wait for
// the thresholding of time chunk k (chunks of `cp` planes).  Tables: plane kernel (ct_plane.cu) or global-memory table
// kernels; ordered phase: cooperative global kernel + event replay (ct_global.cu, ct_fast.cu), with the host replays of
// table_phase() behind it (debug stages, near-ties on non-exact rows, labels that straddle a stale box).
template <typename WaitFn>
int solve_tables(ct_ctx* c, long T, long nchunk, long cp, WaitFn wait_chunk, double overlap, int persistence, int twosided,
                 int stage, long* n_features, cudaStream_t ts) {
    int rc = CT_OK;
    c->fast_tables = 0;
    bool waited = false;
    if (use_plane_kernel(c, T)) {
        for (int attempt = 0; attempt < 4; ++attempt) {
            if ((rc = ctf::begin(c, T, ts)) != CT_OK) return rc;
            for (long k = 0; k < nchunk; ++k) {
                const long t0 = k * cp, nt = std::min(cp, T - t0);
                if (!waited && (rc = wait_chunk(k)) != CT_OK) return rc;
                if ((rc = ctf::chunk(c, t0, t0 + nt, ts)) != CT_OK) return rc;
            }
            waited = true;
            if ((rc = ctf::finish(c, ts)) != CT_OK) return rc;
            int outcome = ctf::FAST_SLOW;
            if (stage == CT_STAGE_FINAL) {
                if ((rc = ctf::global(c, T, overlap, persistence, twosided, n_features, ts, &outcome, nullptr)) != CT_OK) return rc;
            } else {
                if ((rc = ctf::totals_to_host(c, ts, &outcome)) != CT_OK) return rc;
            }
            c->stats["plane_attempts"] = (double)(attempt + 1);
            if (outcome == ctf::FAST_OK) { c->fast_tables = 1; c->stats["fast_path"] = 1.0; return CT_OK; }
            if (outcome == ctf::FAST_SLOW) {
                // valid tables, but the decision needs an ordered host replay (debug stage, near-tie on non-exact rows, a
                // label that straddles a stale box)
                CT_CUDA(c->run_val.ensure((size_t)(c->nruns + 1) * 4));
                c->nseam = 0;
                c->stats["fast_path"] = 0.5;
                return table_phase(c, overlap, persistence, twosided, stage, n_features, ts, true);
            }
            if (outcome == ctf::FAST_RETRY) continue;
            if (outcome == ctf::FAST_FALLBACK && ctf::next_budget(c)) continue;
            break;
        }
    }
    // ---- global-memory table kernels ----
    c->stats["fast_path"] = 0.0;
    tables_begin(c);
    for (long k = 0; k < nchunk && rc == CT_OK; ++k) {
        const long t0 = k * cp, nt = std::min(cp, T - t0);
        if (!waited && (rc = wait_chunk(k)) != CT_OK) return rc;
        rc = tables_chunk(c, t0, t0 + nt, ts);
    }
    if (rc == CT_OK) rc = tables_finish(c, ts);
    if (rc != CT_OK) return rc;
    if (stage == CT_STAGE_FINAL && c->opt_gpu_tables) {
        // the cooperative global kernel on these tables: their counts are known to the host, the control block is written here
        if ((rc = ctf::ensure_global_scratch(c, (size_t)c->ncomp, (size_t)c->nseg)) != CT_OK) return rc;
        struct { uint32_t w[4]; unsigned long long tot[4]; } ctl = {{0, 0, 0, 0},
            {(unsigned long long)c->ncomp, (unsigned long long)c->nseg, (unsigned long long)c->nruns, (unsigned long long)c->npair}};
        memcpy(c->hp_ctl.p, &ctl, sizeof ctl);
        CT_CUDA(cudaMemcpyAsync(c->pl_ctl.p, c->hp_ctl.p, sizeof ctl, cudaMemcpyHostToDevice, ts));
        int outcome = ctf::FAST_SLOW;
        if ((rc = ctf::global(c, T, overlap, persistence, twosided, n_features, ts, &outcome, nullptr)) != CT_OK) return rc;
        if (outcome == ctf::FAST_OK) { c->fast_tables = 1; return CT_OK; }
        if (outcome != ctf::FAST_SLOW) return fail(CT_ERR_INTERNAL, "global phase: unexpected outcome %d", outcome);
    }
    return table_phase(c, overlap, persistence, twosided, stage, n_features, ts, true);
}

// side streams: zero fill at the lowest priority, table phase at the highest
int ensure_streams(ct_ctx* c) {
    int lo = 0, hi = 0;
    CT_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    if (!c->side_stream) CT_CUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, lo));
    if (!c->tbl_stream) CT_CUDA(cudaStreamCreateWithPriority(&c->tbl_stream, cudaStreamNonBlocking, hi));
    for (auto& e : c->ev_side) if (!e) CT_CUDA(cudaEventCreate(&e));
    for (auto& e : c->ev_tbl) if (!e) CT_CUDA(cudaEventCreate(&e));
    return CT_OK;
}

// paint planes [t0, t0+nt) into `flag_dev` (which starts at plane t0)
int launch_paint(ct_ctx* c, long t0, long nt, int32_t* flag_dev, int sparse, cudaStream_t st) {
    ctk::PaintArgs a;
    const long r0 = t0 * c->H;
    a.bits = c->bits.as<uint32_t>() + (size_t)r0 * c->Ww;
    a.row_ptr = c->row_ptr.as<uint32_t>() + r0;
    a.run_val = c->run_val.as<int32_t>();
    a.nrows = nt * c->H; a.W = c->W; a.Ww = c->Ww; a.flag = flag_dev;
    a.sparse = sparse ? 1 : 0;
    a.run_x = c->run_x.as<uint32_t>(); a.run_row = c->run_row.as<uint32_t>(); a.row0 = r0;
    if (c->fast_tables && a.sparse) { a.run_comp = c->run_comp.as<uint32_t>(); a.comp_val = c->c_val.as<int32_t>(); }
    else if (c->fast_tables == 1) {
        // the row-wise / dense paints go by a value per run: materialise it once
        CT_CUDA(c->run_val.ensure((size_t)(c->nruns + 1) * 4));
        a.run_val = c->run_val.as<int32_t>();
        CT_CUDA(ctk::run_values(c->run_comp.as<uint32_t>(), c->c_val.as<int32_t>(), c->run_val.as<int32_t>(), c->nruns, st));
        c->launches += 1;
        c->fast_tables = 2;                                           // (run_val is valid from here on)
    }
    CT_CUDA(ctk::paint(a, c->sm_count, st));
    c->launches += 1;
    return CT_OK;
}

}  // namespace

// ---- host side of the host-buffer entry points: zeroing and run expansion by host threads ----
namespace cti {
// zeroing threads: enough to finish under the host-to-device copy, few enough not to take host memory bandwidth away from it
// (measured on a 16-core box: 4 threads leave the copy at its 55 GB/s but finish 65 ms late, 8 finish in time and slow the
// copy to 51 GB/s); `share` = processes of this host doing the same at the same time (ranks of a sharded run)
void api_host_zero_start(ct_ctx* c, int32_t* flag_host, size_t cells, int share, std::vector<std::thread>& threads) {
    int nthreads = (int)(c->opt_host_zero_threads > 0 ? c->opt_host_zero_threads : c->opt_host_threads);
    if (nthreads <= 0) {
        const unsigned hc = std::thread::hardware_concurrency();
        nthreads = hc >= 16 ? 6 : (hc >= 4 ? (int)hc / 2 : 1);
        // several ranks on this host: their PCIe links work in parallel, so the copies are no longer the bound -- the zeroing
        // is (host memory writes): all cores, shared among the ranks
        if (share > 1) nthreads = std::max(1, (int)hc / share);
    }
    const size_t per = ((cells + nthreads - 1) / nthreads + 1023) / 1024 * 1024;
    for (int i = 0; i < nthreads; ++i) {
        const size_t b = std::min(cells, per * i), e = std::min(cells, per * (i + 1));
        if (e > b) threads.emplace_back([=] { memset(flag_host + b, 0, (e - b) * sizeof(int32_t)); });
    }
}
// row-runs (x0 | x1 << 16, row, value) -> cells of flag_host; rows are shifted by -row_shift (rows below it are skipped: the
// halo plane of a shard).  The painters are bound by cache / TLB misses on the 4 B/cell host cube (every run lands on another
// page): all cores, and the destination of the run 16 ahead is prefetched while the current one is written.  Returns the
// number of threads used.
int api_host_expand_runs(ct_ctx* c, const uint32_t* h_x, const uint32_t* h_row, const int32_t* h_val, long R, long row_shift,
                         int W, int32_t* flag_host, int share) {
    int npaint = (int)c->opt_host_threads;
    if (npaint <= 0) npaint = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency() / (unsigned)std::max(1, share)));
    const int np = (int)std::max(1L, std::min((long)npaint, R / 4096 + 1));
    std::vector<std::thread> painters;
    const long per = (R + np - 1) / np;
    for (int i = 0; i < np; ++i) {
        const long b = std::min(R, per * i), e = std::min(R, per * (i + 1));
        if (e <= b) continue;
        painters.emplace_back([=] {
            constexpr long AHEAD = 16;
            for (long r = b; r < e; ++r) {
                if (r + AHEAD < e && h_val[r + AHEAD] != 0 && (long)h_row[r + AHEAD] >= row_shift) {
                    const int32_t* p = flag_host + (size_t)((long)h_row[r + AHEAD] - row_shift) * W + (h_x[r + AHEAD] & 0xffff);
                    __builtin_prefetch(p, 1, 0);
                    __builtin_prefetch(p + 16, 1, 0);
                }
                const int32_t v = h_val[r];
                if (v == 0 || (long)h_row[r] < row_shift) continue;
                const uint32_t x = h_x[r];
                int32_t* out = flag_host + (size_t)((long)h_row[r] - row_shift) * W;
                for (uint32_t xx = x & 0xffff, x1 = x >> 16; xx < x1; ++xx) out[xx] = v;
            }
        });
    }
    for (auto& t : painters) t.join();
    return np;
}
}  // namespace cti

// ---- what the other translation units of the library use (ct_internal.h) ----
namespace cti {
int api_check_args(long T, int H, int W, const double* w_host, const double* thr_host, long thr_n, int in_dtype, int op) {
    return check_args(T, H, W, w_host, thr_host, thr_n, in_dtype, op);
}
int api_prepare(ct_ctx* c, long T, int H, int W, const double* w_host, const double* thr_host, long thr_n, cudaStream_t st) {
    return prepare(c, T, H, W, w_host, thr_host, thr_n, st);
}
int api_launch_threshold(ct_ctx* c, const void* anom_dev, int in_dtype, long t0, long nt, long thr_n, int thr_is_f32, int op,
                         cudaStream_t st, int all_bits) {
    return launch_threshold(c, anom_dev, in_dtype, t0, nt, thr_n, thr_is_f32, op, st, all_bits);
}
int api_launch_paint(ct_ctx* c, long t0, long nt, int32_t* flag_dev, int sparse, cudaStream_t st) {
    return launch_paint(c, t0, nt, flag_dev, sparse, st);
}
int api_ensure_streams(ct_ctx* c) { return ensure_streams(c); }
int api_table_phase(ct_ctx* c, double overlap, int persistence, int twosided, int stage, long* n_features, cudaStream_t st) {
    return table_phase(c, overlap, persistence, twosided, stage, n_features, st, true);
}
int api_classic_tables(ct_ctx* c, cudaStream_t st) {                // planes [0, c->T) with the global-memory table kernels
    tables_begin(c);
    int rc = c->T > 0 ? tables_chunk(c, 0, c->T, st) : CT_OK;
    return rc != CT_OK ? rc : tables_finish(c, st);
}
bool api_plane_runs(ct_ctx* c, long plane, std::vector<cth::PlaneRun>& out, cudaStream_t st) {
    DeviceRunSource src;
    src.c = c; src.st = st;
    return src.plane_runs(plane, out);
}
}  // namespace cti

// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

int ct_version(void) { return 100; }

const char* ct_last_error(void) { return cti::last_error().c_str(); }

int ct_create(int device, ct_ctx** out) {
    if (!out) return fail(CT_ERR_ARG, "null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(CT_ERR_CUDA, "no usable CUDA device (%s)", e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(CT_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    CT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CT_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(CT_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);
    ct_ctx* c = new ct_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return CT_OK;
}

void ct_destroy(ct_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    DevBuf* bufs[] = {&c->bits, &c->row_cnt, &c->seam_flag, &c->row_ptr, &c->seam_pos, &c->scan_tmp, &c->counters,
                      &c->run_x, &c->run_row, &c->parent, &c->root_flag, &c->rank, &c->run_comp, &c->run_val,
                      &c->c_t, &c->c_y0, &c->c_y1, &c->c_x0, &c->c_x1, &c->c_E, &c->c_S, &c->c_nsp, &c->c_cls, &c->c_val,
                      &c->s_row, &c->s_a, &c->s_b, &c->h_key, &c->h_npix, &c->h_nsp, &c->h_E, &c->h_S,
                      &c->p_b, &c->p_npix, &c->p_nsp, &c->p_E, &c->p_S,
                      &c->k_conE, &c->k_conS, &c->k_fE, &c->k_fS, &c->k_nsp, &c->k_fnsp, &c->pcnt, &c->pfill, &c->pptr,
                      &c->seg_start, &c->seg_pos, &c->g_t, &c->g_y0, &c->g_y1, &c->g_a, &c->g_b,
                      &c->o_t, &c->o_y, &c->o_x0, &c->o_x1, &c->o_val, &c->w_dev, &c->special_dev, &c->thr_dev,
                      &c->chunk_in[0], &c->chunk_in[1], &c->chunk_out[0], &c->chunk_out[1],
                      &c->a_gptr, &c->a_gidx, &c->a_gmean, &c->a_group,
                      &c->l_parent, &c->l_flag, &c->l_rank, &c->l_label, &c->l_kept, &c->l_accE, &c->l_accS, &c->l_accN,
                      &c->b_t0, &c->b_t1, &c->b_y0, &c->b_y1, &c->b_x0, &c->b_x1, &c->b_fin,
                      &c->lc_st, &c->lc_t, &c->lc_label, &c->lc_npix, &c->lc_roll, &c->lc_out, &c->lc_bitmaps, &c->slots, &c->x_q, &c->x_qscratch, &c->x_idx, &c->ovf_rows,
                      &c->pl_chain, &c->pl_done, &c->pl_ctl, &c->g_dirty, &c->g_blocksum, &c->g_evflag, &c->g_ev, &c->g_lrec, &c->g_patch,
                      &c->sh_export, &c->sh_gathered, &c->sh_mdesc, &c->sh_lastplane};
    for (DevBuf* b : bufs) b->release();
    c->hp_counters.release(); c->hp_tables.release(); c->hp_val.release(); c->hp_ovr.release(); c->hp_runs.release(); c->hp_lc.release();
    c->hp_ctl.release(); c->hp_ev.release(); c->hp_ev2.release(); c->hp_patch.release(); c->hp_hdr.release();
    for (auto& e : c->ev_x) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_p) if (e) cudaEventDestroy(e);
    if (c->gctx) ct_destroy(c->gctx);
    if (c->host_stream) cudaStreamDestroy(c->host_stream);
    for (auto& e : c->ev_side) if (e) cudaEventDestroy(e);
    if (c->side_stream) cudaStreamDestroy(c->side_stream);
    for (auto& e : c->ev_tbl) if (e) cudaEventDestroy(e);
    if (c->tbl_stream) cudaStreamDestroy(c->tbl_stream);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_chunk) if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->work_stream) cudaStreamDestroy(c->work_stream);
    delete c;
}

int ct_set_option(ct_ctx* c, const char* key, long value) {
    if (!c || !key) return fail(CT_ERR_ARG, "null argument");
    if (!strcmp(key, "tma")) { c->opt_tma = value; return CT_OK; }
    if (!strcmp(key, "overlap_zero")) { c->opt_overlap_zero = value; return CT_OK; }
    if (!strcmp(key, "gpu_tables")) { c->opt_gpu_tables = value; return CT_OK; }
    if (!strcmp(key, "chunks")) { c->opt_chunks = value < 1 ? 1 : value; return CT_OK; }
    if (!strcmp(key, "chunk_min_planes")) { c->opt_chunk_min_planes = value < 1 ? 1 : value; return CT_OK; }
    if (!strcmp(key, "host_sparse")) { c->opt_host_sparse = value; return CT_OK; }
    if (!strcmp(key, "host_threads")) { c->opt_host_threads = value; return CT_OK; }
    if (!strcmp(key, "host_zero_threads")) { c->opt_host_zero_threads = value; return CT_OK; }
    if (!strcmp(key, "host_out_zeroed")) { c->opt_host_out_zeroed = value; return CT_OK; }
    if (!strcmp(key, "fill_late")) { c->opt_fill_late = value; return CT_OK; }
    if (!strcmp(key, "p2p")) { c->opt_p2p = value; return CT_OK; }
    if (!strcmp(key, "fill_ctas")) { c->opt_fill_ctas = value; return CT_OK; }
    if (!strcmp(key, "profile_tables")) { c->opt_profile_tables = value; return CT_OK; }
    if (!strcmp(key, "plane_kernel")) { c->opt_plane_kernel = value; return CT_OK; }
    if (!strcmp(key, "plane_max_planes")) { c->opt_plane_max_planes = value; return CT_OK; }
    if (!strcmp(key, "fast_chunks")) { c->opt_fast_chunks = value < 1 ? 1 : value; return CT_OK; }
    if (!strcmp(key, "max_sweeps")) { c->opt_max_sweeps = value < 1 ? 1 : value; return CT_OK; }
    if (!strcmp(key, "plane_smem")) { c->opt_plane_smem = value; c->pl_budget = 0; return CT_OK; }
    return fail(CT_ERR_ARG, "unknown option '%s'", key);
}

int ct_run_contrack(ct_ctx* c, const void* anom_dev, int in_dtype, long T, int H, int W, const double* w_host,
                    const double* thr_host, long thr_n, int thr_is_f32, int op, double overlap, int persistence,
                    int twosided, int32_t* flag_dev, long* n_features, int stage, void* stream) {
    if (!c) return fail(CT_ERR_ARG, "null context");
    int rc = check_args(T, H, W, w_host, thr_host, thr_n, in_dtype, op);
    if (rc != CT_OK) return rc;
    if (stage < CT_STAGE_FINAL || stage > CT_STAGE_LABEL3D) return fail(CT_ERR_ARG, "bad stage %d", stage);
    if (n_features) *n_features = 0;
    if (T == 0) return CT_OK;
    if (!anom_dev || !flag_dev) return fail(CT_ERR_ARG, "null device pointer");
    CT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    c->stats.clear();
    if ((rc = prepare(c, T, H, W, w_host, thr_host, thr_n, st)) != CT_OK) return rc;
    c->has_prev = 0;
    // Pipeline: the cube is thresholded in time chunks on `st`; the table kernels of chunk k (latency-bound, small) run
    // on a high-priority stream while chunk k+1 is being thresholded (HBM-bound).  After the last chunk the zero fill of
    // the flag cube -- most of the 4 B/cell the path has to write -- runs on a low-priority stream under the global part
    // of the table phase (overlap filter, 3-D labels, date-line merge, persistence); afterwards only the cells of
    // row-runs are painted.
    const int sparse = c->opt_overlap_zero ? 1 : 0;
    // fast path: ONE threshold launch (it runs alone, at full speed); the plane kernel and the global kernel then run beside
    // the zero fill, which is bound by HBM writes and leaves the issue slots to them ("fast_chunks" > 1 pipelines the plane
    // kernel under the threshold kernel instead)
    const bool fast = use_plane_kernel(c, T);
    const long want_chunks = fast ? c->opt_fast_chunks : c->opt_chunks;
    long nchunk = sparse ? std::min<long>(want_chunks, std::max<long>(1, T / c->opt_chunk_min_planes)) : 1;
    const long cp = (T + nchunk - 1) / nchunk;
    nchunk = (T + cp - 1) / cp;
    cudaStream_t ts = st;
    if (sparse) {
        if ((rc = ensure_streams(c)) != CT_OK) return rc;
        ts = c->tbl_stream;
        while ((long)c->ev_chunk.size() < nchunk) {
            cudaEvent_t e;
            CT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->ev_chunk.push_back(e);
        }
    }
    const size_t plane_bytes = (size_t)H * W * (in_dtype == CT_F64 ? 8 : 4);
    CT_CUDA(cudaEventRecord(c->ev[0], st));
    for (long k = 0; k < nchunk; ++k) {
        const long t0 = k * cp, nt = std::min(cp, T - t0);
        if ((rc = launch_threshold(c, (const char*)anom_dev + (size_t)t0 * plane_bytes, in_dtype, t0, nt, thr_n, thr_is_f32,
                                   op, st, /*all_bits=*/0)) != CT_OK) return rc;
        if (sparse) CT_CUDA(cudaEventRecord(c->ev_chunk[k], st));
    }
    CT_CUDA(cudaEventRecord(c->ev[1], st));
    // Zero fill of the flag cube on the low-priority stream.  Beside the plane kernel both slow down (the fill saturates HBM,
    // the plane kernel lives on memory latency): with "fill_late" the fill starts when the plane kernel has finished and runs
    // beside the global kernel and the host replay instead.
    const bool fill_late = sparse && fast && c->opt_fill_late;
    c->pend_fill = nullptr; c->pend_fill_cells = 0;
    if (sparse) {
        CT_CUDA(cudaEventRecord(c->ev_side[0], st));
        if (fill_late) {
            c->pend_fill = flag_dev; c->pend_fill_cells = (size_t)T * H * W;      // enqueued by ctf::finish()
        } else {
            CT_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_side[0], 0));
            CT_CUDA(ctk::zero_fill(flag_dev, (size_t)T * H * W, c->sm_count, c->side_stream, cti::fill_ctas(c, fast)));
            CT_CUDA(cudaEventRecord(c->ev_side[1], c->side_stream));
            c->launches += 1;
        }
    }
    const double t_h0 = now_ms();
    auto wait_chunk = [&](long k) -> int {
        if (sparse) CT_CUDA(cudaStreamWaitEvent(ts, c->ev_chunk[k], 0));
        return CT_OK;
    };
    rc = solve_tables(c, T, nchunk, cp, wait_chunk, overlap, persistence, twosided, stage, n_features, ts);
    const double t_h1 = now_ms();
    if (rc != CT_OK) {
        if (sparse) { cudaStreamSynchronize(c->side_stream); cudaStreamSynchronize(c->tbl_stream); }
        cudaStreamSynchronize(st);
        return rc;
    }
    c->stats["chunks"] = (double)nchunk;
    c->stats["ms_h_tables"] = t_h1 - t_h0;                           // host wall clock of the table phase (ends in a sync)
    if (sparse) {
        if (c->pend_fill) {                                            // (the tables came from the fallback path: fill now)
            CT_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_side[0], 0));
            CT_CUDA(ctk::zero_fill(c->pend_fill, c->pend_fill_cells, c->sm_count, c->side_stream, cti::fill_ctas(c, false)));
            CT_CUDA(cudaEventRecord(c->ev_side[1], c->side_stream));
            c->launches += 1;
            c->pend_fill = nullptr;
        }
        CT_CUDA(cudaEventRecord(c->ev_tbl[0], ts));
        CT_CUDA(cudaStreamWaitEvent(st, c->ev_tbl[0], 0));
        CT_CUDA(cudaStreamWaitEvent(st, c->ev_side[1], 0));          // the cube is zero
    }
    CT_CUDA(cudaEventRecord(c->ev[3], st));
    if ((rc = launch_paint(c, 0, T, flag_dev, sparse, st)) != CT_OK) return rc;
    if (c->novr) {
        CT_CUDA(ctk::paint_overrides(c->o_t.as<int32_t>(), c->o_y.as<int32_t>(), c->o_x0.as<int32_t>(),
                                     c->o_x1.as<int32_t>(), c->o_val.as<int32_t>(), c->novr, H, W, 0, T, flag_dev, st));
        c->launches += 1;
    }
    CT_CUDA(cudaEventRecord(c->ev[4], st));
    CT_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); c->stats["ms_threshold"] = ms;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); c->stats["ms_tables_gpu"] = ms;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3])); c->stats["ms_tables_host_roundtrip"] = ms;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev[3], c->ev[4])); c->stats["ms_paint"] = ms;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[4])); c->stats["ms_total"] = ms;
    if (sparse) {
        CT_CUDA(cudaEventElapsedTime(&ms, c->ev_side[0], c->ev_side[1])); c->stats["ms_zero_fill"] = ms;
        CT_CUDA(cudaEventElapsedTime(&ms, c->ev[1], c->ev_tbl[0])); c->stats["ms_tables_after_threshold"] = ms;
    }
    c->stats["kernel_launches"] = (double)c->launches;
    return CT_OK;
}

int ct_run_contrack_host(ct_ctx* c, const void* anom_host, int in_dtype, long T, int H, int W, const double* w_host,
                         const double* thr_host, long thr_n, int thr_is_f32, int op, double overlap, int persistence,
                         int twosided, int32_t* flag_host, long* n_features, long chunk_planes) {
    if (!c) return fail(CT_ERR_ARG, "null context");
    int rc = check_args(T, H, W, w_host, thr_host, thr_n, in_dtype, op);
    if (rc != CT_OK) return rc;
    if (n_features) *n_features = 0;
    if (T == 0) return CT_OK;
    if (!anom_host || !flag_host) return fail(CT_ERR_ARG, "null host pointer");
    CT_CUDA(cudaSetDevice(c->device));
    if (!c->copy_stream) CT_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->work_stream) CT_CUDA(cudaStreamCreateWithFlags(&c->work_stream, cudaStreamNonBlocking));
    cudaStream_t cs = c->copy_stream, ws = c->work_stream;
    c->stats.clear();
    if ((rc = prepare(c, T, H, W, w_host, thr_host, thr_n, ws)) != CT_OK) return rc;
    c->has_prev = 0;
    const size_t esz = in_dtype == CT_F64 ? 8 : 4;
    const size_t plane = (size_t)H * W;
    if (chunk_planes <= 0) {
        chunk_planes = (long)((256u << 20) / (plane * esz));
        if (chunk_planes < 1) chunk_planes = 1;
    }
    if (chunk_planes > T) chunk_planes = T;
    const long nchunks = (T + chunk_planes - 1) / chunk_planes;
    cudaEvent_t in_ready[2], in_free[2];
    for (int i = 0; i < 2; ++i) {
        CT_CUDA(cudaEventCreateWithFlags(&in_ready[i], cudaEventDisableTiming));
        CT_CUDA(cudaEventCreateWithFlags(&in_free[i], cudaEventDisableTiming));
        CT_CUDA(c->chunk_in[i].ensure((size_t)chunk_planes * plane * esz));
    }
    const double t0_ms = now_ms();
    // The flag cube is ~96 % zeros on fields like Z500 anomalies and PCIe is the bound of this entry point: instead of a
    // dense cube (4 B/cell) the result travels back as the row-run table (12 B per run, ~1 % of the dense bytes) and host
    // threads expand it into `flag_host`, which they zero-fill while the input chunks are still streaming in.
    const size_t cells = (size_t)T * plane;
    const bool want_sparse = c->opt_host_sparse != 0;
    std::vector<std::thread> zero_threads;
    if (want_sparse && !c->opt_host_out_zeroed) cti::api_host_zero_start(c, flag_host, cells, 1, zero_threads);
    struct Joiner {
        std::vector<std::thread>& v;
        ~Joiner() { for (auto& t : v) if (t.joinable()) t.join(); }
    } joiner{zero_threads};
    // ---- stream the cube in: copy chunk k+1 while chunk k is thresholded; the float cube is never resident ----
    for (long k = 0; k < nchunks; ++k) {
        const int s = (int)(k & 1);
        const long t0 = k * chunk_planes, nt = (t0 + chunk_planes <= T) ? chunk_planes : T - t0;
        if (k >= 2) CT_CUDA(cudaStreamWaitEvent(cs, in_free[s], 0));
        CT_CUDA(cudaMemcpyAsync(c->chunk_in[s].p, (const char*)anom_host + (size_t)t0 * plane * esz,
                                (size_t)nt * plane * esz, cudaMemcpyHostToDevice, cs));
        CT_CUDA(cudaEventRecord(in_ready[s], cs));
        CT_CUDA(cudaStreamWaitEvent(ws, in_ready[s], 0));
        if ((rc = launch_threshold(c, c->chunk_in[s].p, in_dtype, t0, nt, thr_n, thr_is_f32, op, ws)) != CT_OK) return rc;
        CT_CUDA(cudaEventRecord(in_free[s], ws));
    }
    CT_CUDA(cudaStreamSynchronize(ws));
    const double t1_ms = now_ms();
    auto no_wait = [](long) -> int { return CT_OK; };
    if ((rc = solve_tables(c, T, 1, T, no_wait, overlap, persistence, twosided, CT_STAGE_FINAL, n_features, ws)) != CT_OK) return rc;
    if (c->fast_tables) {                                             // the sparse export ships a value per run
        CT_CUDA(c->run_val.ensure((size_t)(c->nruns + 1) * 4));
        CT_CUDA(ctk::run_values(c->run_comp.as<uint32_t>(), c->c_val.as<int32_t>(), c->run_val.as<int32_t>(), c->nruns, ws));
        c->launches += 1;
        c->fast_tables = 2;
    }
    CT_CUDA(cudaStreamSynchronize(ws));
    const double t2_ms = now_ms();
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(in_ready[i]); cudaEventDestroy(in_free[i]); }
    const long R = c->nruns;
    size_t d2h_bytes = 0;
    if (want_sparse && (size_t)R * 12 <= cells * 2) {
        // ---- row-runs (x0|x1<<16, row, value) -> pinned host memory -> host threads paint the non-zero runs ----
        CT_CUDA(c->hp_runs.ensure((size_t)(R + 1) * 12));
        uint32_t* h_x = c->hp_runs.as<uint32_t>();
        uint32_t* h_row = h_x + R;
        int32_t* h_val = reinterpret_cast<int32_t*>(h_row + R);
        if (R) {
            CT_CUDA(cudaMemcpyAsync(h_x, c->run_x.p, (size_t)R * 4, cudaMemcpyDeviceToHost, ws));
            CT_CUDA(cudaMemcpyAsync(h_row, c->run_row.p, (size_t)R * 4, cudaMemcpyDeviceToHost, ws));
            CT_CUDA(cudaMemcpyAsync(h_val, c->run_val.p, (size_t)R * 4, cudaMemcpyDeviceToHost, ws));
        }
        for (auto& t : zero_threads) t.join();                    // (normally long finished)
        zero_threads.clear();
        CT_CUDA(cudaStreamSynchronize(ws));
        d2h_bytes = (size_t)R * 12;
        const int np = cti::api_host_expand_runs(c, h_x, h_row, h_val, R, 0, W, flag_host, 1);
        for (const ctb::Override& o : c->host_result.overrides) {  // pieces of components split at a stale box
            int32_t* out = flag_host + ((size_t)o.t * H + o.y) * W;
            for (int xx = o.x0; xx < o.x1; ++xx) out[xx] = o.val;
        }
        c->stats["host_sparse"] = 1.0;
        c->stats["host_threads"] = (double)np;
    } else {
        for (auto& t : zero_threads) t.join();
        zero_threads.clear();
        // ---- dense: paint chunk k+1 while chunk k travels back ----
        const long out_planes = std::max(1L, std::min(T, (long)((256u << 20) / (plane * 4))));
        const long nout = (T + out_planes - 1) / out_planes;
        cudaEvent_t out_ready[2], out_free[2];
        for (int i = 0; i < 2; ++i) {
            CT_CUDA(cudaEventCreateWithFlags(&out_ready[i], cudaEventDisableTiming));
            CT_CUDA(cudaEventCreateWithFlags(&out_free[i], cudaEventDisableTiming));
            CT_CUDA(c->chunk_out[i].ensure((size_t)out_planes * plane * 4));
        }
        for (long k = 0; k < nout; ++k) {
            const int s = (int)(k & 1);
            const long t0 = k * out_planes, nt = (t0 + out_planes <= T) ? out_planes : T - t0;
            if (k >= 2) CT_CUDA(cudaStreamWaitEvent(ws, out_free[s], 0));
            int32_t* dst = c->chunk_out[s].as<int32_t>();
            if ((rc = launch_paint(c, t0, nt, dst, 0, ws)) != CT_OK) return rc;
            if (c->novr) {
                CT_CUDA(ctk::paint_overrides(c->o_t.as<int32_t>(), c->o_y.as<int32_t>(), c->o_x0.as<int32_t>(),
                                             c->o_x1.as<int32_t>(), c->o_val.as<int32_t>(), c->novr, H, W, t0, t0 + nt,
                                             dst, ws));
                c->launches += 1;
            }
            CT_CUDA(cudaEventRecord(out_ready[s], ws));
            CT_CUDA(cudaStreamWaitEvent(cs, out_ready[s], 0));
            CT_CUDA(cudaMemcpyAsync(flag_host + (size_t)t0 * plane, dst, (size_t)nt * plane * 4, cudaMemcpyDeviceToHost, cs));
            CT_CUDA(cudaEventRecord(out_free[s], cs));
        }
        CT_CUDA(cudaStreamSynchronize(cs));
        CT_CUDA(cudaStreamSynchronize(ws));
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(out_ready[i]); cudaEventDestroy(out_free[i]); }
        d2h_bytes = cells * 4;
        c->stats["host_sparse"] = 0.0;
    }
    const double t3_ms = now_ms();
    c->stats["h2d_bytes"] = (double)(cells * esz);
    c->stats["d2h_bytes"] = (double)d2h_bytes;
    c->stats["ms_h2d_threshold"] = t1_ms - t0_ms;
    c->stats["ms_tables"] = t2_ms - t1_ms;
    c->stats["ms_paint_d2h"] = t3_ms - t2_ms;
    c->stats["ms_total"] = t3_ms - t0_ms;
    c->stats["kernel_launches"] = (double)c->launches;
    return CT_OK;
}

double ct_get_stat(ct_ctx* c, const char* key) {
    if (!c || !key) return -1;
    auto it = c->stats.find(key);
    return it == c->stats.end() ? -1 : it->second;
}

int ct_track_tables(long T, int H, int W, int persistence, long ncomp, const int32_t* comp_t, const int32_t* comp_y0,
                    const int32_t* comp_y1, const int32_t* comp_x0, const int32_t* comp_x1, const int32_t* comp_label,
                    long nseg, const int32_t* seg_t, const int32_t* seg_y0, const int32_t* seg_y1, const int32_t* seg_a,
                    const int32_t* seg_b, const int64_t* run_ptr, const int32_t* run_y, const int32_t* run_x0,
                    const int32_t* run_x1, int32_t* comp_val, long ovr_cap, int32_t* ovr_t, int32_t* ovr_y,
                    int32_t* ovr_x0, int32_t* ovr_x1, int32_t* ovr_val, long* n_ovr, long* n_features, long* n_events,
                    long* n_splits) {
    struct CsrFetcher : ctb::RunFetcher {
        const int64_t* ptr; const int32_t *y, *x0, *x1;
        bool fetch(long comp, std::vector<ctb::SubRun>& out) override {
            out.clear();
            for (int64_t i = ptr[comp]; i < ptr[comp + 1]; ++i) out.push_back(ctb::SubRun{y[i], x0[i], x1[i]});
            return !out.empty();
        }
    } fetcher;
    fetcher.ptr = run_ptr; fetcher.y = run_y; fetcher.x0 = run_x0; fetcher.x1 = run_x1;
    std::vector<ctb::Override> ovr;
    ctb::TrackStats stats;
    int rc = ctb::track_tables(T, H, W, persistence, ncomp, comp_t, comp_y0, comp_y1, comp_x0, comp_x1, comp_label, nseg,
                               seg_t, seg_y0, seg_y1, seg_a, seg_b, run_ptr ? &fetcher : nullptr, comp_val, ovr, stats);
    if (rc != 0) return fail(CT_ERR_INTERNAL, "a component must be split and no run table was supplied");
    if (n_ovr) *n_ovr = (long)ovr.size();
    if ((long)ovr.size() > ovr_cap) return fail(CT_ERR_CAPACITY, "override capacity %ld < %zu", ovr_cap, ovr.size());
    for (size_t i = 0; i < ovr.size(); ++i) {
        ovr_t[i] = ovr[i].t; ovr_y[i] = ovr[i].y; ovr_x0[i] = ovr[i].x0; ovr_x1[i] = ovr[i].x1; ovr_val[i] = ovr[i].val;
    }
    if (n_features) *n_features = stats.n_features;
    if (n_events) *n_events = stats.n_events;
    if (n_splits) *n_splits = stats.n_splits;
    return CT_OK;
}

void ct_classify_rows(const double* w_host, int H, int W, uint8_t* special_out) {
    std::vector<uint8_t> sp;
    classify_rows(w_host, H, W, sp);
    memcpy(special_out, sp.data(), (size_t)H);
}

int ct_host_tables(long T, int H, int W, const double* w_host, double overlap, int persistence, int twosided, int stage,
                   long ncomp, const int32_t* comp_t, const int32_t* comp_y0, const int32_t* comp_y1,
                   const int32_t* comp_x0, const int32_t* comp_x1, const uint32_t* comp_cls, const double* comp_areaE,
                   const double* comp_areaS, const uint32_t* comp_nsp,
                   long npair, const uint32_t* pair_a, const uint32_t* pair_b, const uint32_t* pair_npix,
                   const uint32_t* pair_nsp, const double* pair_areaE, const double* pair_areaS,
                   long nseam, const uint32_t* seam_row, const uint32_t* seam_a, const uint32_t* seam_b,
                   const int64_t* plane_run_ptr, const int32_t* run_y, const int32_t* run_x0, const int32_t* run_x1,
                   const uint32_t* run_comp,
                   int32_t* comp_val, long ovr_cap, int32_t* ovr_t, int32_t* ovr_y, int32_t* ovr_x0, int32_t* ovr_x1,
                   int32_t* ovr_val, long* n_ovr, long* stats8) {
    struct ArraySource : cth::RunSource {
        const int64_t* ptr; const int32_t *y, *x0, *x1; const uint32_t* comp;
        bool plane_runs(long t, std::vector<cth::PlaneRun>& out) override {
            out.clear();
            for (int64_t i = ptr[t]; i < ptr[t + 1]; ++i) out.push_back(cth::PlaneRun{y[i], x0[i], x1[i], comp[i]});
            return true;
        }
    } src;
    src.ptr = plane_run_ptr; src.y = run_y; src.x0 = run_x0; src.x1 = run_x1; src.comp = run_comp;
    cth::Tables tb;
    tb.T = T; tb.H = H; tb.W = W; tb.w = w_host; tb.special_uniform = special_rows_uniform(w_host, H, W);
    tb.ncomp = ncomp; tb.comp_t = comp_t; tb.comp_y0 = comp_y0; tb.comp_y1 = comp_y1; tb.comp_x0 = comp_x0;
    tb.comp_x1 = comp_x1; tb.comp_cls = comp_cls; tb.comp_areaE = comp_areaE; tb.comp_areaS = comp_areaS;
    tb.comp_nsp = comp_nsp;
    tb.npair = npair; tb.pair_a = pair_a; tb.pair_b = pair_b; tb.pair_npix = pair_npix; tb.pair_nsp = pair_nsp;
    tb.pair_areaE = pair_areaE; tb.pair_areaS = pair_areaS;
    tb.nseam = nseam; tb.seam_row = seam_row; tb.seam_a = seam_a; tb.seam_b = seam_b;
    cth::Params pr;
    pr.overlap = overlap; pr.persistence = persistence; pr.twosided = twosided; pr.stage = stage;
    cth::Result res;
    std::string err;
    int rc = cth::host_phase(tb, pr, plane_run_ptr ? &src : nullptr, res, err);
    if (rc != 0) return fail(rc, "%s", err.c_str());
    if (ncomp) memcpy(comp_val, res.comp_val.data(), (size_t)ncomp * 4);
    if (n_ovr) *n_ovr = (long)res.overrides.size();
    if ((long)res.overrides.size() > ovr_cap)
        return fail(CT_ERR_CAPACITY, "override capacity %ld < %zu", ovr_cap, res.overrides.size());
    for (size_t i = 0; i < res.overrides.size(); ++i) {
        const ctb::Override& o = res.overrides[i];
        ovr_t[i] = o.t; ovr_y[i] = o.y; ovr_x0[i] = o.x0; ovr_x1[i] = o.x1; ovr_val[i] = o.val;
    }
    if (stats8) {
        stats8[0] = res.n_features; stats8[1] = res.n_kept; stats8[2] = res.n_labels3d; stats8[3] = res.n_seam_events;
        stats8[4] = res.n_seam_splits; stats8[5] = res.n_neartie; stats8[6] = 0; stats8[7] = 0;
    }
    return CT_OK;
}

int ct_host_tables_fast(long T, int H, int W, const double* w_host, double overlap, int persistence, int twosided,
                        int stage, long ncomp, const int32_t* comp_t, const int32_t* comp_y0, const int32_t* comp_y1,
                        const int32_t* comp_x0, const int32_t* comp_x1, const uint32_t* comp_cls,
                        const double* cls_conE, const double* cls_conS, const double* cls_fE, const double* cls_fS,
                        const uint32_t* cls_nsp, const uint32_t* pair_ptr, const uint32_t* pair_b,
                        const uint32_t* pair_npix, const uint32_t* pair_nsp, const double* pair_E, const double* pair_S,
                        long nseg, const int32_t* seg_t, const int32_t* seg_y0, const int32_t* seg_y1,
                        const uint32_t* seg_a, const uint32_t* seg_b, ct_plane_runs_fn fetch, void* user,
                        int32_t* comp_val, long ovr_cap, int32_t* ovr_t, int32_t* ovr_y, int32_t* ovr_x0, int32_t* ovr_x1,
                        int32_t* ovr_val, long* n_ovr, long* stats8) {
    struct CallbackSource : cth::RunSource {
        ct_plane_runs_fn fn; void* user;
        bool plane_runs(long t, std::vector<cth::PlaneRun>& out) override {
            long n = 0;
            const int32_t *y = nullptr, *x0 = nullptr, *x1 = nullptr;
            const uint32_t* comp = nullptr;
            if (fn(user, t, &n, &y, &x0, &x1, &comp) != 0) return false;
            out.resize(n);
            for (long i = 0; i < n; ++i) out[i] = cth::PlaneRun{y[i], x0[i], x1[i], comp[i]};
            return true;
        }
    } src;
    src.fn = fetch; src.user = user;
    cth::FastTables tb;
    tb.T = T; tb.H = H; tb.W = W; tb.w = w_host; tb.ncomp = ncomp; tb.special_uniform = special_rows_uniform(w_host, H, W);
    tb.comp_t = comp_t; tb.comp_y0 = comp_y0; tb.comp_y1 = comp_y1; tb.comp_x0 = comp_x0; tb.comp_x1 = comp_x1;
    tb.comp_cls = comp_cls; tb.cls_conE = cls_conE; tb.cls_conS = cls_conS; tb.cls_fE = cls_fE; tb.cls_fS = cls_fS;
    tb.cls_nsp = cls_nsp; tb.pair_ptr = pair_ptr; tb.pair_b = pair_b; tb.pair_npix = pair_npix; tb.pair_nsp = pair_nsp;
    tb.pair_E = pair_E; tb.pair_S = pair_S; tb.nseg = nseg; tb.seg_t = seg_t; tb.seg_y0 = seg_y0; tb.seg_y1 = seg_y1;
    tb.seg_a = seg_a; tb.seg_b = seg_b;
    cth::Params pr;
    pr.overlap = overlap; pr.persistence = persistence; pr.twosided = twosided; pr.stage = stage;
    cth::Result res;
    std::string err;
    int rc = cth::host_phase_fast(tb, pr, fetch ? &src : nullptr, comp_val, res, err);
    if (rc != 0) return fail(rc, "%s", err.c_str());
    if (n_ovr) *n_ovr = (long)res.overrides.size();
    if ((long)res.overrides.size() > ovr_cap)
        return fail(CT_ERR_CAPACITY, "override capacity %ld < %zu", ovr_cap, res.overrides.size());
    for (size_t i = 0; i < res.overrides.size(); ++i) {
        const ctb::Override& o = res.overrides[i];
        ovr_t[i] = o.t; ovr_y[i] = o.y; ovr_x0[i] = o.x0; ovr_x1[i] = o.x1; ovr_val[i] = o.val;
    }
    if (stats8) {
        stats8[0] = res.n_features; stats8[1] = res.n_kept; stats8[2] = res.n_labels3d; stats8[3] = res.n_seam_events;
        stats8[4] = res.n_seam_splits; stats8[5] = res.n_neartie; stats8[6] = 0; stats8[7] = 0;
    }
    return CT_OK;
}

// ---- run_lifecycle (contrack.py:799-907) ---------------------------------------------------------------------------
int ct_run_lifecycle(ct_ctx* c, const int32_t* flag_dev, const void* var_dev, int var_dtype, long T, int H, int W,
                     const double* w_host, long* n_rows, void* stream) {
    if (!c || !n_rows) return fail(CT_ERR_ARG, "null argument");
    *n_rows = 0;
    c->lc_rows = 0;
    if (T < 0 || H <= 0 || W <= 0) return fail(CT_ERR_ARG, "bad shape T=%ld H=%d W=%d", T, H, W);
    if (H > 65535 || W > 65535) return fail(CT_ERR_CAPACITY, "H and W must be <= 65535 (got %d x %d)", H, W);
    if ((double)T * H >= 2147483647.0) return fail(CT_ERR_CAPACITY, "T*H must be < 2^31");
    if (var_dtype != CT_F32 && var_dtype != CT_F64) return fail(CT_ERR_ARG, "var_dtype must be CT_F32 or CT_F64");
    if (T == 0) return CT_OK;
    if (!flag_dev || !var_dev || !w_host) return fail(CT_ERR_ARG, "null pointer");
    CT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    c->stats.clear();
    c->T = T; c->H = H; c->W = W; c->Ww = (W + 31) / 32;
    c->nruns = c->ncomp = c->npair = 0;                              // the run_contrack tables of this context are gone
    const long nrows = T * H;
    const int Ww = c->Ww;
    auto U = [](DevBuf& b) { return b.as<uint32_t>(); };
    CT_CUDA(c->w_dev.ensure(H * sizeof(double)));
    CT_CUDA(cudaMemcpyAsync(c->w_dev.p, w_host, H * sizeof(double), cudaMemcpyHostToDevice, st));
    CT_CUDA(c->bits.ensure((size_t)nrows * Ww * 4)); CT_CUDA(c->lc_st.ensure((size_t)nrows * Ww * 4));
    CT_CUDA(c->row_cnt.ensure((size_t)nrows * 4)); CT_CUDA(c->row_ptr.ensure((size_t)(nrows + 1) * 4));
    CT_CUDA(c->counters.ensure(64)); CT_CUDA(c->hp_counters.ensure(64));
    CT_CUDA(c->scan_tmp.ensure(ctk::scan_tmp_elems(nrows > 1024 ? nrows : 1024) * 4));
    uint32_t* cnt_dev = c->counters.as<uint32_t>();
    uint32_t* cnt_host = c->hp_counters.as<uint32_t>();
    const double t0_ms = now_ms();
    for (auto& e : c->ev) if (!e) CT_CUDA(cudaEventCreate(&e));
    CT_CUDA(cudaEventRecord(c->ev[0], st));
    CT_CUDA(ctl::lc_rows(flag_dev, nrows, W, Ww, U(c->bits), U(c->lc_st), U(c->row_cnt), c->sm_count, st));
    CT_CUDA(cudaEventRecord(c->ev[1], st));
    CT_CUDA(ctk::exclusive_scan_u32(U(c->row_cnt), U(c->row_ptr), nrows, U(c->scan_tmp), st));
    CT_CUDA(cudaMemcpyAsync(cnt_host, U(c->row_ptr) + nrows, 4, cudaMemcpyDeviceToHost, st));
    CT_CUDA(cudaStreamSynchronize(st));                              // also: w_host may go away after the call
    const long R = cnt_host[0];
    long launches = 4;
    CT_CUDA(c->run_x.ensure((size_t)(R + 1) * 4)); CT_CUDA(c->run_row.ensure((size_t)(R + 1) * 4));
    CT_CUDA(c->run_val.ensure((size_t)(R + 1) * 4));
    CT_CUDA(ctl::lc_extract(flag_dev, U(c->bits), U(c->lc_st), U(c->row_ptr), nrows, W, Ww, U(c->run_x), U(c->run_row),
                            c->run_val.as<int32_t>(), st));
    launches += 1;
    ctl::EntryTable et;
    uint64_t want = (uint64_t)R / 8 + 1024;
    for (int attempt = 0;; ++attempt) {
        et.cap = next_pow2(want);
        CT_CUDA(c->h_key.ensure((size_t)et.cap * 8)); CT_CUDA(c->h_npix.ensure((size_t)et.cap * 4));
        CT_CUDA(c->h_nsp.ensure((size_t)et.cap * 4));
        et.key = c->h_key.as<unsigned long long>(); et.npix = U(c->h_npix); et.flags = U(c->h_nsp);
        et.overflow = cnt_dev + 4; et.count = cnt_dev + 5; et.nroll = cnt_dev + 6;
        CT_CUDA(ctl::lc_entries(U(c->run_x), U(c->run_row), c->run_val.as<int32_t>(), R, H, W, et, st));
        launches += 2;
        CT_CUDA(cudaMemcpyAsync(cnt_host + 4, cnt_dev + 4, 8, cudaMemcpyDeviceToHost, st));
        CT_CUDA(cudaStreamSynchronize(st));
        if (cnt_host[4] == 0 && (uint64_t)cnt_host[5] * 10 <= (uint64_t)et.cap * 7) break;
        if (attempt >= 8 || et.cap >= (1u << 31)) return fail(CT_ERR_CAPACITY, "lifecycle entry table overflow");
        want = (uint64_t)et.cap * 4;
    }
    const long ne = cnt_host[5];
    CT_CUDA(c->lc_t.ensure((size_t)(ne + 1) * 4)); CT_CUDA(c->lc_label.ensure((size_t)(ne + 1) * 4));
    CT_CUDA(c->lc_npix.ensure((size_t)(ne + 1) * 4)); CT_CUDA(c->lc_roll.ensure((size_t)(ne + 1) * 4));
    CT_CUDA(c->lc_out.ensure((size_t)(ne + 1) * 5 * 8));
    CT_CUDA(cudaMemsetAsync(cnt_dev + 7, 0, 4, st));
    CT_CUDA(ctl::lc_compact(et, c->lc_t.as<int32_t>(), c->lc_label.as<int32_t>(), U(c->lc_npix), c->lc_roll.as<int32_t>(),
                            cnt_dev + 7, st));
    CT_CUDA(cudaMemcpyAsync(cnt_host + 6, cnt_dev + 6, 4, cudaMemcpyDeviceToHost, st));
    CT_CUDA(cudaStreamSynchronize(st));
    const long nroll = cnt_host[6];
    launches += 1;
    if (nroll) {
        CT_CUDA(c->lc_bitmaps.ensure((size_t)nroll * Ww * 4));
        CT_CUDA(cudaMemsetAsync(c->lc_bitmaps.p, 0, (size_t)nroll * Ww * 4, st));
        CT_CUDA(ctl::lc_roll(U(c->row_ptr), U(c->run_x), c->run_val.as<int32_t>(), c->lc_t.as<int32_t>(),
                             c->lc_label.as<int32_t>(), c->lc_roll.as<int32_t>(), ne, H, W, Ww, U(c->lc_bitmaps), st));
        launches += 1;
    }
    double* o = c->lc_out.as<double>();
    const size_t ne1 = (size_t)ne + 1;
    CT_CUDA(ctl::lc_sums(U(c->row_ptr), U(c->run_x), U(c->run_row), c->run_val.as<int32_t>(), var_dev, var_dtype == CT_F64,
                         c->w_dev.as<double>(), c->lc_t.as<int32_t>(), c->lc_label.as<int32_t>(), U(c->lc_npix),
                         c->lc_roll.as<int32_t>(), ne, H, W, o, o + ne1, o + 2 * ne1, o + 3 * ne1, o + 4 * ne1, st));
    launches += ne ? 1 : 0;
    CT_CUDA(cudaEventRecord(c->ev[2], st));
    // results -> pinned host memory: 5 doubles, then t, label, npix, roll
    CT_CUDA(c->hp_lc.ensure(ne1 * (5 * 8 + 4 * 4)));
    char* hp = c->hp_lc.as<char>();
    if (ne) {
        CT_CUDA(cudaMemcpyAsync(hp, o, ne1 * 5 * 8, cudaMemcpyDeviceToHost, st));
        CT_CUDA(cudaMemcpyAsync(hp + ne1 * 40, c->lc_t.p, (size_t)ne * 4, cudaMemcpyDeviceToHost, st));
        CT_CUDA(cudaMemcpyAsync(hp + ne1 * 44, c->lc_label.p, (size_t)ne * 4, cudaMemcpyDeviceToHost, st));
        CT_CUDA(cudaMemcpyAsync(hp + ne1 * 48, c->lc_npix.p, (size_t)ne * 4, cudaMemcpyDeviceToHost, st));
        CT_CUDA(cudaMemcpyAsync(hp + ne1 * 52, c->lc_roll.p, (size_t)ne * 4, cudaMemcpyDeviceToHost, st));
    }
    CT_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); c->stats["ms_lc_rows"] = ms;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); c->stats["ms_lc_tables_sums"] = ms;
    c->stats["ms_total"] = now_ms() - t0_ms;
    c->stats["runs"] = (double)R; c->stats["lc_rows"] = (double)ne; c->stats["lc_rolled"] = (double)nroll;
    c->stats["kernel_launches"] = (double)launches;
    c->lc_rows = ne;
    *n_rows = ne;
    return CT_OK;
}

int ct_lifecycle_fetch(ct_ctx* c, long cap, int32_t* t, int32_t* label, int32_t* npix, int32_t* roll, double* area,
                       double* wsum, double* norm, double* sy, double* sx) {
    if (!c) return fail(CT_ERR_ARG, "null context");
    const long ne = c->lc_rows;
    if (cap < ne) return fail(CT_ERR_CAPACITY, "capacity %ld < %ld rows", cap, ne);
    if (ne == 0) return CT_OK;
    if (!t || !label || !npix || !roll || !area || !wsum || !norm || !sy || !sx) return fail(CT_ERR_ARG, "null pointer");
    const size_t ne1 = (size_t)ne + 1;
    const char* hp = c->hp_lc.as<char>();
    const double* o = reinterpret_cast<const double*>(hp);
    memcpy(area, o, (size_t)ne * 8); memcpy(wsum, o + ne1, (size_t)ne * 8); memcpy(norm, o + 2 * ne1, (size_t)ne * 8);
    memcpy(sy, o + 3 * ne1, (size_t)ne * 8); memcpy(sx, o + 4 * ne1, (size_t)ne * 8);
    memcpy(t, hp + ne1 * 40, (size_t)ne * 4); memcpy(label, hp + ne1 * 44, (size_t)ne * 4);
    memcpy(npix, hp + ne1 * 48, (size_t)ne * 4); memcpy(roll, hp + ne1 * 52, (size_t)ne * 4);
    return CT_OK;
}

double ct_numpy_pairwise_sum_rle(const double* value, const int64_t* count, long n) {
    std::vector<double> v;
    for (long i = 0; i < n; ++i) v.insert(v.end(), (size_t)count[i], value[i]);
    return ctb::numpy_pairwise_sum(v.data(), (long)v.size());
}

int ct_calc_clim(ct_ctx* c, const float* z_dev, long T, int H, int W, const int32_t* group_host, int G, int window,
                 float* clim_dev, void* stream) {
    return ct_calc_clim_t(c, z_dev, CT_F32, T, H, W, group_host, G, window, clim_dev, stream);
}

int ct_calc_clim_t(ct_ctx* c, const void* z_dev, int dtype, long T, int H, int W, const int32_t* group_host, int G, int window,
                   void* clim_dev, void* stream) {
    if (!c || !z_dev || !group_host || !clim_dev) return fail(CT_ERR_ARG, "null argument");
    if (dtype != CT_F32 && dtype != CT_F64) return fail(CT_ERR_ARG, "dtype must be CT_F32 or CT_F64");
    const int f64 = dtype == CT_F64;
    if (T <= 0 || H <= 0 || W <= 0 || G <= 0 || window <= 0) return fail(CT_ERR_ARG, "bad shape / window");
    CT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long HW = (long)H * W;
    // members of every group, in time order
    std::vector<int32_t> gptr(G + 1, 0), gidx(T);
    for (long t = 0; t < T; ++t) {
        if (group_host[t] < 0 || group_host[t] >= G) return fail(CT_ERR_ARG, "group index out of range at t=%ld", t);
        gptr[group_host[t] + 1]++;
    }
    for (int g = 0; g < G; ++g) gptr[g + 1] += gptr[g];
    {
        std::vector<int32_t> pos(gptr.begin(), gptr.end() - 1);
        for (long t = 0; t < T; ++t) gidx[pos[group_host[t]]++] = (int32_t)t;
    }
    CT_CUDA(c->a_gptr.ensure((size_t)(G + 1) * 4)); CT_CUDA(c->a_gidx.ensure((size_t)T * 4));
    CT_CUDA(c->a_gmean.ensure((size_t)G * HW * (f64 ? 8 : 4)));
    CT_CUDA(cudaMemcpyAsync(c->a_gptr.p, gptr.data(), (size_t)(G + 1) * 4, cudaMemcpyHostToDevice, st));
    CT_CUDA(cudaMemcpyAsync(c->a_gidx.p, gidx.data(), (size_t)T * 4, cudaMemcpyHostToDevice, st));
    CT_CUDA(cudaStreamSynchronize(st));                               // gptr / gidx are locals
    CT_CUDA(cta::group_mean(z_dev, f64, HW, G, c->a_gptr.as<int32_t>(), c->a_gidx.as<int32_t>(), c->a_gmean.p, st));
    CT_CUDA(cta::clim_smooth(c->a_gmean.p, f64, HW, G, window, clim_dev, st));
    return CT_OK;
}

int ct_calc_anom(ct_ctx* c, const float* z_dev, long T, int H, int W, const int32_t* group_host, int G,
                 const float* clim_dev, int smooth, float* anom_dev, void* stream) {
    return ct_calc_anom_t(c, z_dev, CT_F32, T, H, W, group_host, G, clim_dev, smooth, anom_dev, stream);
}

int ct_calc_anom_t(ct_ctx* c, const void* z_dev, int dtype, long T, int H, int W, const int32_t* group_host, int G,
                   const void* clim_dev, int smooth, void* anom_dev, void* stream) {
    if (!c || !z_dev || !group_host || !clim_dev || !anom_dev) return fail(CT_ERR_ARG, "null argument");
    if (dtype != CT_F32 && dtype != CT_F64) return fail(CT_ERR_ARG, "dtype must be CT_F32 or CT_F64");
    const int f64 = dtype == CT_F64;
    if (T <= 0 || H <= 0 || W <= 0 || G <= 0 || smooth <= 0) return fail(CT_ERR_ARG, "bad shape / smooth");
    for (long t = 0; t < T; ++t)
        if (group_host[t] < 0 || group_host[t] >= G) return fail(CT_ERR_ARG, "group index out of range at t=%ld", t);
    CT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    // time chunks: a new chunk where the group index wraps (a new year of a day-of-year climatology); long chunks are cut
    // so that there are enough of them to fill the machine, at most 4096
    std::vector<int32_t> cs;
    cs.push_back(0);
    for (long t = 1; t < T; ++t) if (group_host[t] < group_host[t - 1]) cs.push_back((int32_t)t);
    cs.push_back((int32_t)T);
    if (T >= 2147483647L) return fail(CT_ERR_CAPACITY, "T must be < 2^31");
    while ((long)cs.size() - 1 < 24 && (long)cs.size() - 1 < T) {       // few cycles: halve every chunk
        std::vector<int32_t> h;
        for (size_t i = 0; i + 1 < cs.size(); ++i) {
            h.push_back(cs[i]);
            if (cs[i + 1] - cs[i] > 1) h.push_back(cs[i] + (cs[i + 1] - cs[i]) / 2);
        }
        h.push_back((int32_t)T);
        if (h.size() == cs.size()) break;
        cs.swap(h);
    }
    const int nchunks = (int)cs.size() - 1;
    CT_CUDA(c->a_group.ensure((size_t)(T + nchunks + 2) * 4));
    int32_t* gdev = c->a_group.as<int32_t>();
    CT_CUDA(cudaMemcpyAsync(gdev, group_host, (size_t)T * 4, cudaMemcpyHostToDevice, st));
    CT_CUDA(cudaMemcpyAsync(gdev + T, cs.data(), (size_t)(nchunks + 1) * 4, cudaMemcpyHostToDevice, st));
    CT_CUDA(cudaStreamSynchronize(st));                               // caller may free group_host after return; cs is local
    if (nchunks <= 65535)
        CT_CUDA(cta::anom_chunks(z_dev, f64, (long)H * W, T, gdev + T, nchunks, gdev, clim_dev, smooth, anom_dev, st));
    else
        CT_CUDA(cta::anom(z_dev, f64, (long)H * W, T, gdev, clim_dev, smooth, anom_dev, st));
    return CT_OK;
}

// ---- callers either side of the path: ct_extras.cu -------------------------------------------------------
int ct_quantile_time(ct_ctx* c, const float* x_dev, long T, int H, int W, int y0, int y1, const double* q_host, int nq,
                     double* out_dev, void* stream) {
    return ct_quantile_time_t(c, nullptr, x_dev, CT_F32, T, H, W, y0, y1, q_host, nq, out_dev, stream);
}

int ct_flag_count(ct_ctx* c, const int32_t* flag_dev, long T, int H, int W, int greater_than, int32_t* count_dev,
                  void* stream) {
    if (!c || !flag_dev || !count_dev) return fail(CT_ERR_ARG, "null argument");
    if (T < 0 || H <= 0 || W <= 0) return fail(CT_ERR_ARG, "bad shape");
    CT_CUDA(cudaSetDevice(c->device));
    CT_CUDA(cte::flag_count(flag_dev, T, H, W, greater_than, count_dev, c->sm_count, (cudaStream_t)stream));
    return CT_OK;
}

int ct_divide_f32(ct_ctx* c, const float* in_dev, size_t n, float divisor, float* out_dev, void* stream) {
    if (!c || !in_dev || !out_dev) return fail(CT_ERR_ARG, "null argument");
    CT_CUDA(cudaSetDevice(c->device));
    CT_CUDA(cte::divide_f32(in_dev, n, divisor, out_dev, (cudaStream_t)stream));
    return CT_OK;
}

int ct_gather_planes(ct_ctx* c, const float* src_dev, int G, int Hs, int Ws, const int32_t* iy_host, const int32_t* ix_host,
                     int H, int W, float* dst_dev, void* stream) {
    return ct_gather_planes_t(c, src_dev, CT_F32, G, Hs, Ws, iy_host, ix_host, H, W, dst_dev, stream);
}

int ct_gather_planes_t(ct_ctx* c, const void* src_dev, int dtype, int G, int Hs, int Ws, const int32_t* iy_host,
                       const int32_t* ix_host, int H, int W, void* dst_dev, void* stream) {
    if (!c || !src_dev || !iy_host || !ix_host || !dst_dev) return fail(CT_ERR_ARG, "null argument");
    if (dtype != CT_F32 && dtype != CT_F64) return fail(CT_ERR_ARG, "dtype must be CT_F32 or CT_F64");
    if (G <= 0 || Hs <= 0 || Ws <= 0 || H <= 0 || W <= 0 || H > 65535) return fail(CT_ERR_ARG, "bad shape");
    for (int y = 0; y < H; ++y) if (iy_host[y] < 0 || iy_host[y] >= Hs) return fail(CT_ERR_ARG, "row index out of range");
    for (int x = 0; x < W; ++x) if (ix_host[x] < 0 || ix_host[x] >= Ws) return fail(CT_ERR_ARG, "column index out of range");
    CT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    CT_CUDA(c->x_idx.ensure((size_t)(H + W) * 4));
    int32_t* d = c->x_idx.as<int32_t>();
    CT_CUDA(cudaMemcpyAsync(d, iy_host, (size_t)H * 4, cudaMemcpyHostToDevice, st));
    CT_CUDA(cudaMemcpyAsync(d + H, ix_host, (size_t)W * 4, cudaMemcpyHostToDevice, st));
    CT_CUDA(cudaStreamSynchronize(st));
    CT_CUDA(cte::gather_planes(src_dev, dtype == CT_F64, G, Hs, Ws, d, d + H, H, W, dst_dev, st));
    return CT_OK;
}

}  // extern "C"
