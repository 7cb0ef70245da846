// ct_fast.cu -- host side of the fast table pipeline: per-plane table kernel (ct_plane.cu) -> cooperative global kernel
// (ct_global.cu) -> O(events) host replay of the stale-box date-line merge (ct_tables.cpp: track_events_fast) -> patches.
// Between the threshold kernel and the paint there is ONE host synchronisation (the read-back of the control block and the
// date-line events); table sizes are never read by the host to size a launch: capacities are checked on the device.
#include "ct_fast.h"

#include <cstring>

namespace ctf {

namespace {

using cti::fail;

__global__ void k_plane_init(unsigned long long* chain, long stride, uint32_t* ctl) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        chain[0] = 2ull; chain[stride] = 2ull; chain[2 * stride] = 2ull;      // slot 0: inclusive prefix 0
        ctl[0] = 0; ctl[1] = 0; ctl[2] = 0;                                     // ticket, status, info
    }
}

// totals {components, segments, runs, pairs} from the last chain slot
__global__ void k_plane_totals(const unsigned long long* chain, long stride, long planes, unsigned long long* totals) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const unsigned long long a = chain[planes] >> 2, r = chain[stride + planes] >> 2, p = chain[2 * stride + planes] >> 2;
        totals[0] = a >> 31; totals[1] = a & 0x7fffffffull; totals[2] = r; totals[3] = p;
    }
}

__global__ void __launch_bounds__(256) k_apply_patches(const int32_t* __restrict__ lab, const int32_t* __restrict__ val, long n,
                                                       int32_t* __restrict__ fin) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fin[lab[i]] = val[i];
}

__global__ void __launch_bounds__(256) k_comp_values(const int32_t* __restrict__ label, const int32_t* __restrict__ fin,
                                                     const unsigned long long* __restrict__ totals, int32_t* __restrict__ val) {
    const long nc = (long)totals[0];
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += (long)gridDim.x * blockDim.x) val[i] = fin[label[i]];
}

constexpr size_t SMALL_BUDGET = 40 * 1024;       // fits beside a threshold block (185 KB of the SM's 228 KB)
constexpr size_t LARGE_BUDGET = 200 * 1024;

inline uint32_t cap_of(const DevBuf& b, size_t elt, size_t slack) {
    const size_t n = b.cap / elt;
    return (uint32_t)std::min<size_t>(n > slack ? n - slack : 0, 0xfffffff0u);
}

struct Caps { uint32_t runs, comps, pairs, segs; };

Caps current_caps(ct_ctx* c) {
    Caps k;
    k.runs = std::min(std::min(cap_of(c->run_x, 4, 2), cap_of(c->run_row, 4, 2)), cap_of(c->run_comp, 4, 2));
    uint32_t cc = 0xfffffff0u;
    DevBuf* b4[] = {&c->c_t, &c->c_y0, &c->c_y1, &c->c_x0, &c->c_x1, &c->c_nsp, &c->c_cls, &c->c_val, &c->k_nsp, &c->k_fnsp,
                    &c->pcnt, &c->pfill, &c->pptr, &c->l_parent, &c->l_rank, &c->l_label, &c->l_accN,
                    &c->b_t0, &c->b_t1, &c->b_y0, &c->b_y1, &c->b_x0, &c->b_x1, &c->b_fin};
    DevBuf* b8[] = {&c->c_E, &c->c_S, &c->k_conE, &c->k_conS, &c->k_fE, &c->k_fS, &c->l_accE, &c->l_accS};
    for (DevBuf* b : b4) cc = std::min(cc, cap_of(*b, 4, 4));
    for (DevBuf* b : b8) cc = std::min(cc, cap_of(*b, 8, 4));
    cc = std::min(cc, cap_of(c->l_kept, 1, 16));
    k.comps = cc;
    uint32_t pp = 0xfffffff0u;
    DevBuf* p4[] = {&c->p_b, &c->p_npix, &c->p_nsp};
    DevBuf* p8[] = {&c->p_E, &c->p_S};
    for (DevBuf* b : p4) pp = std::min(pp, cap_of(*b, 4, 2));
    for (DevBuf* b : p8) pp = std::min(pp, cap_of(*b, 8, 2));
    k.pairs = pp;
    uint32_t ss = 0xfffffff0u;
    DevBuf* s4[] = {&c->g_t, &c->g_y0, &c->g_y1, &c->g_a, &c->g_b, &c->g_evflag};
    for (DevBuf* b : s4) ss = std::min(ss, cap_of(*b, 4, 3));
    ss = std::min(ss, cap_of(c->g_ev, 2 * 4, 1));
    ss = std::min(ss, cap_of(c->g_lrec, 2 * 7 * 4, 1));           // at most two labels per event segment
    ss = std::min(ss, cap_of(c->l_flag, 4, 2));                 // (the root flags double as event positions)
    k.segs = ss;
    return k;
}

}  // namespace

// every table / scratch buffer of the fast path at (at least) the given element counts
int ensure_tables(ct_ctx* c, size_t runs, size_t comps, size_t pairs, size_t segs) {
    DevBuf* r4[] = {&c->run_x, &c->run_row, &c->run_comp};
    for (DevBuf* b : r4) CT_CUDA(b->ensure((runs + 2) * 4));
    DevBuf* b4[] = {&c->c_t, &c->c_y0, &c->c_y1, &c->c_x0, &c->c_x1, &c->c_nsp, &c->c_cls, &c->c_val, &c->k_nsp, &c->k_fnsp,
                    &c->pcnt, &c->pfill, &c->pptr, &c->l_parent, &c->l_rank, &c->l_label, &c->l_accN,
                    &c->b_t0, &c->b_t1, &c->b_y0, &c->b_y1, &c->b_x0, &c->b_x1, &c->b_fin};
    DevBuf* b8[] = {&c->c_E, &c->c_S, &c->k_conE, &c->k_conS, &c->k_fE, &c->k_fS, &c->l_accE, &c->l_accS};
    for (DevBuf* b : b4) CT_CUDA(b->ensure((comps + 4) * 4));
    for (DevBuf* b : b8) CT_CUDA(b->ensure((comps + 4) * 8));
    CT_CUDA(c->l_kept.ensure(comps + 16));
    CT_CUDA(c->l_flag.ensure((std::max(comps, segs) + 4) * 4));
    DevBuf* p4[] = {&c->p_b, &c->p_npix, &c->p_nsp};
    DevBuf* p8[] = {&c->p_E, &c->p_S};
    for (DevBuf* b : p4) CT_CUDA(b->ensure((pairs + 2) * 4));
    for (DevBuf* b : p8) CT_CUDA(b->ensure((pairs + 2) * 8));
    DevBuf* s4[] = {&c->g_t, &c->g_y0, &c->g_y1, &c->g_a, &c->g_b, &c->g_evflag};
    for (DevBuf* b : s4) CT_CUDA(b->ensure((segs + 3) * 4));
    CT_CUDA(c->g_ev.ensure((segs + 1) * 2 * 4));
    CT_CUDA(c->g_lrec.ensure((segs + 1) * 2 * 7 * 4));
    return CT_OK;
}

// scratch of the global kernel only (the tables themselves are left alone: they may hold data)
int ensure_global_scratch(ct_ctx* c, size_t comps, size_t segs) {
    DevBuf* b4[] = {&c->c_val, &c->l_parent, &c->l_rank, &c->l_label, &c->l_accN, &c->b_t0, &c->b_t1, &c->b_y0, &c->b_y1,
                    &c->b_x0, &c->b_x1, &c->b_fin};
    DevBuf* b8[] = {&c->l_accE, &c->l_accS};
    for (DevBuf* b : b4) CT_CUDA(b->ensure((comps + 4) * 4));
    for (DevBuf* b : b8) CT_CUDA(b->ensure((comps + 4) * 8));
    CT_CUDA(c->l_kept.ensure(comps + 16));
    CT_CUDA(c->l_flag.ensure((std::max(comps, segs) + 4) * 4));
    CT_CUDA(c->g_evflag.ensure((segs + 3) * 4));
    CT_CUDA(c->g_ev.ensure((segs + 1) * 2 * 4));
    CT_CUDA(c->g_lrec.ensure((segs + 1) * 2 * 7 * 4));
    return ensure_control(c);
}

int ensure_control(ct_ctx* c) {
    CT_CUDA(c->pl_ctl.ensure(256));
    CT_CUDA(c->hp_ctl.ensure(256));
    if (!c->coop_grid) {
        c->coop_grid = ctp::global_grid(c->sm_count);
        if (c->coop_grid <= 0) return fail(CT_ERR_CUDA, "the cooperative global kernel does not fit this device");
    }
    CT_CUDA(c->g_blocksum.ensure((size_t)(c->coop_grid + 2) * 4));
    return CT_OK;
}

int begin(ct_ctx* c, long planes, cudaStream_t st) {
    const int H = c->H;
    const size_t nrows = (size_t)planes * H;
    // first call on this context: sizes from the row count (typical anomaly fields: 1.4 runs per row, 2 % as many components
    // as runs); a table that turns out too small is reported by the kernels and rebuilt with the exact totals
    const size_t estR = nrows * 2 + 4096, estC = estR / 8 + 4096, estP = estC * 2, estS = nrows / 8 + 4096;
    int rc = ensure_tables(c, estR, estC, estP, estS);
    if (rc != CT_OK) return rc;
    if ((rc = ensure_control(c)) != CT_OK) return rc;
    CT_CUDA(c->pl_chain.ensure((size_t)3 * (planes + 1) * 8));
    CT_CUDA(c->pl_done.ensure((size_t)(planes + 1) * 4));
    CT_CUDA(cudaMemsetAsync(c->pl_chain.p, 0, (size_t)3 * (planes + 1) * 8, st));
    CT_CUDA(cudaMemsetAsync(c->pl_done.p, 0, (size_t)(planes + 1) * 4, st));
    k_plane_init<<<1, 32, 0, st>>>(c->pl_chain.as<unsigned long long>(), planes + 1, c->pl_ctl.as<uint32_t>());
    CT_CUDA(cudaGetLastError());
    c->pl_planes = planes;
    c->launches += 1;
    if (!c->pl_budget) c->pl_budget = c->opt_plane_smem > 0 ? (size_t)c->opt_plane_smem : SMALL_BUDGET;
    c->tb_planes = 0;
    return CT_OK;
}

int chunk(ct_ctx* c, long p0, long p1, cudaStream_t st) {
    if (p0 != c->tb_planes || p1 <= p0 || p1 > c->pl_planes) return fail(CT_ERR_INTERNAL, "plane chunks out of order");
    ctp::PlaneArgs a;
    auto U = [](DevBuf& b) { return b.as<uint32_t>(); };
    a.row_cnt = U(c->row_cnt); a.seam_flag = U(c->seam_flag); a.slots = U(c->slots); a.bits = U(c->bits);
    a.H = c->H; a.W = c->W; a.Ww = c->Ww; a.p0 = p0; a.np = p1 - p0;
    a.w = c->w_dev.as<double>(); a.special = c->special_dev.as<uint8_t>();
    a.row_ptr = U(c->row_ptr); a.run_x = U(c->run_x); a.run_row = U(c->run_row); a.run_comp = U(c->run_comp);
    a.ct.t = c->c_t.as<int32_t>(); a.ct.y0 = c->c_y0.as<int32_t>(); a.ct.y1 = c->c_y1.as<int32_t>();
    a.ct.x0 = c->c_x0.as<int32_t>(); a.ct.x1 = c->c_x1.as<int32_t>(); a.ct.areaE = c->c_E.as<double>();
    a.ct.areaS = c->c_S.as<double>(); a.ct.nsp = U(c->c_nsp); a.ct.cls = U(c->c_cls);
    a.kt.conE = c->k_conE.as<double>(); a.kt.conS = c->k_conS.as<double>(); a.kt.fE = c->k_fE.as<double>();
    a.kt.fS = c->k_fS.as<double>(); a.kt.nsp = U(c->k_nsp); a.kt.fnsp = U(c->k_fnsp);
    a.pcnt = U(c->pcnt); a.pfill = U(c->pfill); a.pptr = U(c->pptr);
    a.pc.b = U(c->p_b); a.pc.npix = U(c->p_npix); a.pc.nsp = U(c->p_nsp); a.pc.E = c->p_E.as<double>(); a.pc.S = c->p_S.as<double>();
    a.sg.t = c->g_t.as<int32_t>(); a.sg.y0 = c->g_y0.as<int32_t>(); a.sg.y1 = c->g_y1.as<int32_t>();
    a.sg.a = U(c->g_a); a.sg.b = U(c->g_b);
    const Caps k = current_caps(c);
    a.cap_runs = k.runs; a.cap_comps = k.comps; a.cap_pairs = k.pairs; a.cap_segs = k.segs;
    a.chain = c->pl_chain.as<unsigned long long>(); a.chain_stride = c->pl_planes + 1;
    a.done = U(c->pl_done); a.ticket = U(c->pl_ctl); a.status = U(c->pl_ctl) + 1; a.info = U(c->pl_ctl) + 2;
    if (!ctp::plane_config(c->H, c->pl_budget, &a.smem_runs, &a.hash_cap))
        return fail(CT_ERR_CAPACITY, "H = %d rows do not fit the plane kernel's shared memory", c->H);
    const size_t smem = ctp::plane_smem_bytes(c->H, a.smem_runs, &a.smem_scan_off);
    CT_CUDA(cudaMemsetAsync(c->pl_ctl.p, 0, 4, st));                 // ticket of this launch
    for (auto& e : c->ev_p) if (!e) CT_CUDA(cudaEventCreate(&e));
    if (p0 == 0) CT_CUDA(cudaEventRecord(c->ev_p[0], st));
    CT_CUDA(ctp::plane_tables(a, smem, st));
    CT_CUDA(cudaEventRecord(c->ev_p[1], st));
    c->plane_timed = 1;
    c->launches += 1;
    c->tb_planes = p1;
    return CT_OK;
}

int finish(ct_ctx* c, cudaStream_t st) {
    if (c->tb_planes != c->pl_planes) return fail(CT_ERR_INTERNAL, "tables cover %ld of %ld planes", c->tb_planes, c->pl_planes);
    k_plane_totals<<<1, 32, 0, st>>>(c->pl_chain.as<unsigned long long>(), c->pl_planes + 1, c->pl_planes,
                                     reinterpret_cast<unsigned long long*>(c->pl_ctl.as<char>() + 16));
    CT_CUDA(cudaGetLastError());
    c->launches += 1;
    if (c->pend_fill) {
        // the zero fill of the flag planes was held back until the plane kernel is done ("fill_late"): from here on it runs
        // on the low-priority stream beside the global kernel and the host replay
        CT_CUDA(cudaEventRecord(c->ev_tbl[1], st));
        CT_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_tbl[1], 0));
        CT_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_side[0], 0));
        CT_CUDA(ctk::zero_fill(c->pend_fill, c->pend_fill_cells, c->sm_count, c->side_stream, cti::fill_ctas(c, true)));
        CT_CUDA(cudaEventRecord(c->ev_side[1], c->side_stream));
        c->launches += 1;
        c->pend_fill = nullptr;
    }
    return CT_OK;
}

// larger shared-memory budget after ST_FALLBACK; false when the large one was already in use
bool next_budget(ct_ctx* c) {
    if (c->pl_budget >= LARGE_BUDGET || c->opt_plane_smem > 0) return false;
    c->pl_budget = LARGE_BUDGET;
    return true;
}

// control block -> host; FAST_SLOW when the tables are valid (their counts are in the context afterwards)
int totals_to_host(ct_ctx* c, cudaStream_t st, int* outcome) {
    uint32_t* hctl = c->hp_ctl.as<uint32_t>();
    CT_CUDA(cudaMemcpyAsync(hctl, c->pl_ctl.p, 128, cudaMemcpyDeviceToHost, st));
    CT_CUDA(cudaStreamSynchronize(st));
    const uint32_t status = hctl[1];
    const unsigned long long* tot = reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(hctl) + 16);
    c->ncomp = (long)tot[0]; c->nseg = (long)tot[1]; c->nruns = (long)tot[2]; c->npair = (long)tot[3];
    c->stats["runs"] = (double)c->nruns; c->stats["comps2d"] = (double)c->ncomp; c->stats["pairs"] = (double)c->npair;
    c->stats["seam_segments"] = (double)c->nseg;
    if (hctl[2] & 1u) c->stats["slot_overflow"] = 1.0;
    if (status & ctp::ST_TIMEOUT) return fail(CT_ERR_INTERNAL, "plane kernel: a predecessor plane never published its tables");
    if (status & ctp::ST_FALLBACK) { *outcome = FAST_FALLBACK; return CT_OK; }
    if (status & ctp::ST_CAPACITY) {
        if (tot[2] >= 0xfffffff0ull)
            return fail(CT_ERR_CAPACITY, "more than 2^32 - 16 row-runs in the cube: the run tables are indexed with 32 bits");
        if (tot[0] >= 0x7ffffff0ull || tot[3] >= 0x7ffffff0ull) return fail(CT_ERR_CAPACITY, "tables exceed 2^31 entries");
        auto grow = [](unsigned long long v) { return (size_t)(v + v / 8 + 1024); };
        int rc = ensure_tables(c, grow(tot[2]), grow(tot[0]), grow(tot[3]), grow(tot[1]));
        if (rc != CT_OK) return rc;
        *outcome = FAST_RETRY;
        return CT_OK;
    }
    *outcome = FAST_SLOW;
    return CT_OK;
}

int global(ct_ctx* c, long T, double overlap, int persistence, int twosided, long* n_features, cudaStream_t st, int* outcome,
           uint32_t* raw_status) {
    *outcome = FAST_SLOW;
    int rc = ensure_control(c);
    if (rc != CT_OK) return rc;
    const Caps k = current_caps(c);
    CT_CUDA(c->g_dirty.ensure((size_t)2 * (T + 2) + 16));
    auto U = [](DevBuf& b) { return b.as<uint32_t>(); };
    uint32_t* ctl = U(c->pl_ctl);
    ctp::GlobalArgs a;
    a.totals = reinterpret_cast<const unsigned long long*>(c->pl_ctl.as<char>() + 16);
    a.T = T;
    a.status = ctl + 1; a.cap_comps = k.comps; a.cap_segs = k.segs;
    a.comp_t = c->c_t.as<int32_t>(); a.comp_y0 = c->c_y0.as<int32_t>(); a.comp_y1 = c->c_y1.as<int32_t>();
    a.comp_x0 = c->c_x0.as<int32_t>(); a.comp_x1 = c->c_x1.as<int32_t>(); a.cls = U(c->c_cls);
    a.conE = c->k_conE.as<double>(); a.conS = c->k_conS.as<double>(); a.fE = c->k_fE.as<double>(); a.fS = c->k_fS.as<double>();
    a.nsp = U(c->k_nsp); a.fnsp = U(c->k_fnsp);
    a.pair_ptr = U(c->pptr); a.pair_b = U(c->p_b); a.pair_npix = U(c->p_npix); a.pair_nsp = U(c->p_nsp);
    a.pair_E = c->p_E.as<double>(); a.pair_S = c->p_S.as<double>();
    a.seg_a = U(c->g_a); a.seg_b = U(c->g_b);
    a.overlap = overlap; a.twosided = twosided; a.special_uniform = c->special_uniform; a.persistence = persistence;
    a.max_sweeps = (int)std::max<long>(1, c->opt_max_sweeps);
    a.kept = c->l_kept.as<uint8_t>(); a.accE = c->l_accE.as<double>(); a.accS = c->l_accS.as<double>(); a.accN = U(c->l_accN);
    a.dirty = c->g_dirty.as<uint8_t>();
    a.parent = U(c->l_parent); a.rootflag = U(c->l_flag); a.rank = U(c->l_rank); a.label = c->l_label.as<int32_t>();
    a.bt0 = c->b_t0.as<int32_t>(); a.bt1 = c->b_t1.as<int32_t>(); a.by0 = c->b_y0.as<int32_t>(); a.by1 = c->b_y1.as<int32_t>();
    a.bx0 = c->b_x0.as<int32_t>(); a.bx1 = c->b_x1.as<int32_t>(); a.fin = c->b_fin.as<int32_t>();
    a.blocksum = U(c->g_blocksum); a.evflag = U(c->g_evflag); a.ev = c->g_ev.as<int32_t>(); a.lrec = c->g_lrec.as<int32_t>();
    a.out8 = ctl + 12;
    const double t_g0 = cti::now_ms();
    for (auto& e : c->ev_p) if (!e) CT_CUDA(cudaEventCreate(&e));
    CT_CUDA(cudaEventRecord(c->ev_p[2], st));
    CT_CUDA(ctp::global_phase(a, c->coop_grid, st));
    CT_CUDA(cudaEventRecord(c->ev_p[3], st));
    c->launches += 1;
    // ---- control block + (speculatively) the first events and label records in one round trip ----
    constexpr long EV_FIRST = 32768, REC_FIRST = 8192;
    uint32_t* hctl = c->hp_ctl.as<uint32_t>();
    const long ev_first = std::min<long>(EV_FIRST, (long)k.segs), rec_first = std::min<long>(REC_FIRST, 2 * (long)k.segs);
    CT_CUDA(c->hp_ev.ensure((size_t)EV_FIRST * 8 + (size_t)REC_FIRST * 28));
    CT_CUDA(cudaMemcpyAsync(hctl, ctl, 128, cudaMemcpyDeviceToHost, st));
    for (auto& x : c->extra_d2h) {
        if (x.dst) CT_CUDA(cudaMemcpyAsync(x.dst, x.src, x.bytes, cudaMemcpyDeviceToHost, st));
        x.dst = nullptr;
    }
    if (ev_first > 0) {
        CT_CUDA(cudaMemcpyAsync(c->hp_ev.p, c->g_ev.p, (size_t)ev_first * 8, cudaMemcpyDeviceToHost, st));
        CT_CUDA(cudaMemcpyAsync(c->hp_ev.as<char>() + (size_t)EV_FIRST * 8, c->g_lrec.p, (size_t)rec_first * 28,
                                cudaMemcpyDeviceToHost, st));
    }
    CT_CUDA(cudaEventRecord(c->ev[2], st));
    CT_CUDA(cudaStreamSynchronize(st));
    const double t_g1 = cti::now_ms();
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->ev_p[2], c->ev_p[3]) == cudaSuccess) c->stats["ms_global_kernel"] = ms;
        if (c->plane_timed && cudaEventElapsedTime(&ms, c->ev_p[0], c->ev_p[1]) == cudaSuccess) c->stats["ms_plane_kernel"] = ms;
        c->plane_timed = 0;
    }
    const uint32_t status = hctl[1];
    const unsigned long long* tot = reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(hctl) + 16);
    const uint32_t* out8 = hctl + 12;
    c->ncomp = (long)tot[0]; c->nseg = (long)tot[1]; c->nruns = (long)tot[2]; c->npair = (long)tot[3];
    c->stats["runs"] = (double)c->nruns; c->stats["comps2d"] = (double)c->ncomp; c->stats["pairs"] = (double)c->npair;
    c->stats["seam_segments"] = (double)c->nseg;
    if (raw_status) {                                                // (sharded run: the caller interprets the merged status)
        *raw_status = status;
        if (status) { *outcome = FAST_STATUS; return CT_OK; }
    }
    if (status & ctp::ST_TIMEOUT) return fail(CT_ERR_INTERNAL, "plane kernel: a predecessor plane never published its tables");
    if (status & ctp::ST_FALLBACK) { *outcome = FAST_FALLBACK; return CT_OK; }
    if (status & ctp::ST_CAPACITY) {
        if (tot[2] >= 0xfffffff0ull)
            return fail(CT_ERR_CAPACITY, "more than 2^32 - 16 row-runs in the cube: the run tables are indexed with 32 bits");
        if (tot[0] >= 0x7ffffff0ull || tot[3] >= 0x7ffffff0ull) return fail(CT_ERR_CAPACITY, "tables exceed 2^31 entries");
        auto grow = [](unsigned long long v) { return (size_t)(v + v / 8 + 1024); };
        if ((rc = ensure_tables(c, grow(tot[2]), grow(tot[0]), grow(tot[3]), grow(tot[1]))) != CT_OK) return rc;
        *outcome = FAST_RETRY;
        return CT_OK;
    }
    if (hctl[2] & 1u) c->stats["slot_overflow"] = 1.0;
    c->stats["sweeps"] = (double)out8[0];
    c->stats["wavefront_planes"] = (double)out8[5];
    c->stats["neartie_flagged"] = (double)out8[1];
    c->stats["ms_g_kernel"] = t_g1 - t_g0;
    if (out8[1]) { *outcome = FAST_SLOW; return CT_OK; }            // near-tie on non-exact rows: the exact host resolver
    const long nlab = out8[2], nev = out8[3], nrec = out8[6];
    // ---- date-line events -> host replay at label granularity ----
    const int32_t* ev = c->hp_ev.as<int32_t>();
    const int32_t* lrec = reinterpret_cast<const int32_t*>(c->hp_ev.as<char>() + (size_t)EV_FIRST * 8);
    if (nev > ev_first || nrec > rec_first) {
        CT_CUDA(c->hp_ev2.ensure((size_t)nev * 8 + (size_t)nrec * 28 + 64));
        CT_CUDA(cudaMemcpyAsync(c->hp_ev2.p, c->g_ev.p, (size_t)nev * 8, cudaMemcpyDeviceToHost, st));
        CT_CUDA(cudaMemcpyAsync(c->hp_ev2.as<char>() + (size_t)nev * 8, c->g_lrec.p, (size_t)nrec * 28, cudaMemcpyDeviceToHost, st));
        CT_CUDA(cudaStreamSynchronize(st));
        ev = c->hp_ev2.as<int32_t>();
        lrec = reinterpret_cast<const int32_t*>(c->hp_ev2.as<char>() + (size_t)nev * 8);
    }
    static thread_local std::vector<int32_t> plab, pval;
    ctb::TrackStats ts;
    long feat_delta = 0;
    if (ctb::track_events_fast(persistence, nev, ev, nrec, lrec, plab, pval, &feat_delta, ts) != 0) {
        c->stats["label_fast"] = 0.0;
        *outcome = FAST_SLOW;                                        // a label straddles a stale box: per-component replay
        return CT_OK;
    }
    const long np = (long)plab.size();
    if (np) {
        CT_CUDA(c->hp_patch.ensure((size_t)np * 8));
        CT_CUDA(c->g_patch.ensure((size_t)np * 8));
        int32_t* hp = c->hp_patch.as<int32_t>();
        memcpy(hp, plab.data(), (size_t)np * 4); memcpy(hp + np, pval.data(), (size_t)np * 4);
        CT_CUDA(cudaMemcpyAsync(c->g_patch.p, hp, (size_t)np * 8, cudaMemcpyHostToDevice, st));
        k_apply_patches<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(c->g_patch.as<int32_t>(), c->g_patch.as<int32_t>() + np, np,
                                                                      c->b_fin.as<int32_t>());
        CT_CUDA(cudaGetLastError());
        c->launches += 1;
    }
    k_comp_values<<<std::max(1, std::min(c->sm_count * 4, (int)((c->ncomp + 255) / 256))), 256, 0, st>>>(
        c->l_label.as<int32_t>(), c->b_fin.as<int32_t>(), a.totals, c->c_val.as<int32_t>());
    CT_CUDA(cudaGetLastError());
    c->launches += 1;
    cth::Result& res = c->host_result;
    res.overrides.clear();
    res.n_neartie = 0; res.n_labels3d = nlab; res.n_features = (long)out8[4] + feat_delta;
    res.n_seam_events = ts.n_events; res.n_seam_splits = 0;
    c->novr = 0;
    c->stats["ms_host_tables"] = cti::now_ms() - t_g1;
    c->stats["labels3d"] = (double)nlab; c->stats["features"] = (double)res.n_features;
    c->stats["seam_events"] = (double)res.n_seam_events; c->stats["seam_splits"] = 0.0;
    c->stats["neartie_resolved"] = 0.0; c->stats["moved_comps"] = 0.0; c->stats["override_runs"] = 0.0;
    c->stats["label_fast"] = 1.0; c->stats["event_segments"] = (double)nev; c->stats["ht_walked"] = (double)ts.n_walked;
    if (n_features) *n_features = res.n_features;
    *outcome = FAST_OK;
    return CT_OK;
}

}  // namespace ctf
