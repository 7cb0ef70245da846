// ct_global.cu -- the global part of the path on the tables, as ONE cooperative kernel (grid-wide barriers instead of ~40
// launches and 4 host round trips):
//
//   contrack.py:706-742  time-sequential overlap filter.  The verdict of a class at plane t depends on the FINAL verdicts of
//                        plane t-1 only, so the recurrence has exactly one solution.  Jacobi sweeps (all classes at once)
//                        reach it; only planes whose predecessor changed in the previous sweep are re-evaluated.  After
//                        `max_sweeps` sweeps one block finishes the job as a plane-ordered wavefront (the reference's own
//                        order), so adversarial chains cost O(planes) barrier steps inside this kernel and no host trips.
//   contrack.py:747-751  3-D labels: union-find over kept components that share a pixel in adjacent planes, numbered by
//                        first pixel (= smallest component id), grid-wide scan of the root flags
//   contrack.py:753, 765-772  label boxes, persistence verdict per label, and the list of date-line segments whose two ends
//                        carry different labels ("events", with the boxes of both labels): all the host needs to replay the
//                        stale-box merge (ct_tables.cpp) in O(events)
#include "ct_plane.h"

#include <cooperative_groups.h>

#include <climits>

namespace cg = cooperative_groups;

namespace ctp {

namespace {

constexpr int GLOBAL_THREADS = 512;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
    volatile uint32_t* p = parent;
    while (true) {
        const uint32_t q = p[x];
        if (q == x) return x;
        x = q;
    }
}
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }
        const uint32_t old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

__device__ __forceinline__ bool kill_decision(double areacon, double fwd, double bwd, double ov, bool twosided, double* fb_out,
                                              double* ff_out) {
    const double inv = __ddiv_rn(1.0, areacon);       // contrack.py:721-722: reciprocal, then multiply -- two roundings
    const double fb = __dmul_rn(inv, bwd);
    const double ff = __dmul_rn(inv, fwd);
    *fb_out = fb; *ff_out = ff;
    bool kill = false;
    if (twosided) {
        if (fb != 0 && ff != 0) { if ((fb < ov) || (ff < ov)) kill = true; }
        if (fb != 0 && ff == 0) { if (fb < ov) kill = true; }
        if (fb == 0 && ff != 0) { if (ff < ov) kill = true; }
    } else {
        if (ff < ov) kill = true;
    }
    return kill;
}

// backward overlap of component c with the kept components of the plane before it, added to its class
__device__ __forceinline__ void acc_one(const GlobalArgs& a, long c) {
    double e = 0.0, s2 = 0.0;
    uint32_t n = 0;
    const uint32_t k1 = a.pair_ptr[c + 1];
    for (uint32_t k = a.pair_ptr[c]; k < k1; ++k) {
        if (!__ldcg(a.kept + a.cls[a.pair_b[k]])) continue;
        e += a.pair_E[k]; s2 += a.pair_S[k]; n += a.pair_nsp[k];
    }
    const uint32_t rep = a.cls[c];
    if (e != 0.0) atomicAdd(&a.accE[rep], e);
    if (n) { atomicAdd(&a.accS[rep], s2); atomicAdd(&a.accN[rep], n); }
}

// verdict of class representative c from the accumulated sums; returns true if it changed
__device__ __forceinline__ bool decide_one(const GlobalArgs& a, long c, uint32_t* nflag) {
    const double bE = __ldcg(a.accE + c), bS = __ldcg(a.accS + c);
    const uint32_t bn = __ldcg(a.accN + c);
    a.accE[c] = 0.0; a.accS[c] = 0.0; a.accN[c] = 0;
    const double areacon = __dadd_rn(a.conE[c], a.conS[c]), fwd = __dadd_rn(a.fE[c], a.fS[c]), bwd = __dadd_rn(bE, bS);
    double fb, ff;
    const bool kill = kill_decision(areacon, fwd, bwd, a.overlap, a.twosided != 0, &fb, &ff);
    if (nflag && a.nsp[c] + a.fnsp[c] + bn > 0) {
        // sums with special-row weights are not exactly summable in general: a fraction within rounding distance of
        // `overlap` must be decided in numpy's summation order (host).  Exception: a class that lies entirely in special
        // rows of one common weight -- every sum is a small integer multiple of that weight, exact in any order.
        const bool near = (fabs(ff - a.overlap) <= 1e-9) || (a.twosided && fabs(fb - a.overlap) <= 1e-9);
        const bool exact = a.special_uniform && a.conE[c] == 0.0 && a.fE[c] == 0.0 && bE == 0.0;
        if (near && !exact) atomicAdd(nflag, 1u);
    }
    const uint8_t nk = kill ? 0 : 1;
    if (__ldcg(a.kept + c) != nk) { a.kept[c] = nk; return true; }
    return false;
}

__device__ __forceinline__ long lower_bound_t(const int32_t* t, long n, int v) {      // first component with plane >= v
    long lo = 0, hi = n;
    while (lo < hi) { const long m = (lo + hi) >> 1; if (t[m] < v) lo = m + 1; else hi = m; }
    return lo;
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* s_w) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    if (lane == 0) s_w[wid] = v;
    __syncthreads();
    uint32_t x = 0;
    if (wid == 0) {
        x = lane < nw ? s_w[lane] : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(FULL, x, d);
        if (lane == 0) s_w[32] = x;
    }
    __syncthreads();
    x = s_w[32];
    __syncthreads();
    return x;
}
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_w, uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += n; }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const uint32_t x = lane < nw ? s_w[lane] : 0;
        uint32_t xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(FULL, xi, d); if (lane >= d) xi += n; }
        if (lane < nw) s_w[lane] = xi - x;
        if (lane == 31) s_w[32] = xi;
    }
    __syncthreads();
    const uint32_t ex = inc - v + s_w[wid];
    *total = s_w[32];
    __syncthreads();
    return ex;
}

// grid-wide exclusive scan of flag[0..n): pos[i] = number of set flags before i; returns the total (same in every thread).
// Each block owns a contiguous tile.  Two grid barriers.
__device__ uint32_t grid_excl_scan(cg::grid_group& grid, const uint32_t* flag, long n, uint32_t* pos, uint32_t* blocksum,
                                   uint32_t* s_w) {
    const long tile = (n + gridDim.x - 1) / gridDim.x;
    const long b0 = (long)blockIdx.x * tile, b1 = b0 + tile < n ? b0 + tile : n;
    uint32_t s = 0;
    for (long i = b0 + threadIdx.x; i < b1; i += blockDim.x) s += __ldcg(flag + i);
    const uint32_t bs = block_sum(s, s_w);
    if (threadIdx.x == 0) blocksum[blockIdx.x] = bs;
    grid.sync();
    uint32_t before = 0, all = 0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        const uint32_t v = __ldcg(blocksum + b);
        all += v;
        if (b < blockIdx.x) before += v;
    }
    before = block_sum(before, s_w);
    all = block_sum(all, s_w);
    uint32_t run = before;
    for (long i0 = b0; i0 < b1; i0 += blockDim.x) {
        const long i = i0 + threadIdx.x;
        const uint32_t f = i < b1 ? __ldcg(flag + i) : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(f, s_w, &tot);
        if (i < b1) pos[i] = run + ex;
        run += tot;
    }
    grid.sync();
    return all;
}

__global__ void __launch_bounds__(GLOBAL_THREADS) k_global_phase(GlobalArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t s_w[40];
    const long NC = (long)a.totals[0], NS = (long)a.totals[1];
    const long T = a.T;
    // incomplete tables (a plane fell back / a table was too small): nothing to do, the host rebuilds them.  Every thread
    // reads the same words, so the whole grid leaves before the first barrier.
    if (*a.status != 0u || NC > (long)a.cap_comps || NS > (long)a.cap_segs) return;
    const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gsize = (long)gridDim.x * blockDim.x;
    uint32_t* ctr = a.out8 + 8;                                   // [2] near-tie flags, [3] features, [4] wavefront planes,
                                                                  // [5..7] "changed" of sweep k at 5 + k % 3: a flag is
                                                                  // cleared two barriers after its last reader
    uint8_t* dirty0 = a.dirty;
    uint8_t* dirty1 = a.dirty + (T + 2);

    // ---- init ----
    for (long c = gtid; c < NC; c += gsize) { a.kept[c] = 1; a.accE[c] = 0.0; a.accS[c] = 0.0; a.accN[c] = 0; a.parent[c] = (uint32_t)c; }
    for (long t = gtid; t < T + 2; t += gsize) { dirty0[t] = 1; dirty1[t] = 0; }
    if (gtid == 0) { for (int i = 0; i < 8; ++i) ctr[i] = 0; for (int i = 0; i < 8; ++i) a.out8[i] = 0; }
    grid.sync();

    // ---- step 3: keep / kill fixpoint ----
    uint32_t sweeps = 0, wave_planes = 0;
    bool full = true;                                             // the running sweep evaluates every plane
    int cur = 0;
    while (true) {
        uint8_t* dc = cur ? dirty1 : dirty0;
        uint8_t* dn = cur ? dirty0 : dirty1;
        for (long t = gtid; t < T + 2; t += gsize) dn[t] = 0;
        if (gtid == 0) { ctr[5 + (sweeps + 1) % 3] = 0; ctr[2] = 0; }
        for (long c = gtid; c < NC; c += gsize) {
            const int tt = a.comp_t[c];
            if (tt < 1 || tt + 1 >= T || !dc[tt]) continue;
            acc_one(a, c);
        }
        grid.sync();
        for (long c = gtid; c < NC; c += gsize) {
            if (a.cls[c] != (uint32_t)c) continue;
            const int tt = a.comp_t[c];
            if (tt < 1 || tt + 1 >= T || !dc[tt]) continue;
            if (decide_one(a, c, full ? ctr + 2 : nullptr)) { dn[tt + 1] = 1; ctr[5 + sweeps % 3] = 1u; }
        }
        grid.sync();
        const bool changed = __ldcg(ctr + 5 + sweeps % 3) != 0u;
        ++sweeps;
        cur ^= 1;
        if (!changed) {
            if (full) break;                                      // a full sweep changed nothing: the fixpoint, flags counted
            for (long t = gtid; t < T + 2; t += gsize) (cur ? dirty1 : dirty0)[t] = 1;     // verify with one full sweep
            full = true;
            grid.sync();
            continue;
        }
        full = false;
        if ((int)sweeps >= a.max_sweeps) {
            // ---- plane-ordered wavefront from the first dirty plane (block 0; the other blocks wait at the barrier) ----
            if (blockIdx.x == 0) {
                uint8_t* dd = cur ? dirty1 : dirty0;
                __shared__ long s_first;
                if (threadIdx.x == 0) {
                    long f = T;
                    for (long t = 1; t < T; ++t) if (dd[t]) { f = t; break; }
                    s_first = f;
                }
                __syncthreads();
                long lo = lower_bound_t(a.comp_t, NC, (int)s_first);
                for (long t = s_first; t + 1 < T; ++t) {
                    long hi = lo;
                    while (hi < NC && a.comp_t[hi] == (int)t) ++hi;     // (every thread walks the same few entries)
                    for (long c = lo + threadIdx.x; c < hi; c += blockDim.x) acc_one(a, c);
                    __syncthreads();
                    for (long c = lo + threadIdx.x; c < hi; c += blockDim.x)
                        if (a.cls[c] == (uint32_t)c) decide_one(a, c, nullptr);
                    __threadfence_block();
                    __syncthreads();
                    lo = hi;
                    if (threadIdx.x == 0) ++wave_planes;
                }
                if (threadIdx.x == 0) ctr[4] = wave_planes;
            }
            for (long t = gtid; t < T + 2; t += gsize) { dirty0[t] = 1; dirty1[t] = 1; }      // then verify: one full sweep
            full = true;
            grid.sync();
        }
    }
    const uint32_t nflag = __ldcg(ctr + 2);
    if (nflag) {                                                  // a verdict needs the exact host resolver
        if (gtid == 0) { a.out8[0] = sweeps; a.out8[1] = nflag; a.out8[5] = __ldcg(ctr + 4); }
        return;
    }

    // ---- step 4a/b: 3-D labels ----
    for (long c = gtid; c < NC; c += gsize) {
        if (!a.kept[a.cls[c]]) continue;
        const uint32_t k1 = a.pair_ptr[c + 1];
        for (uint32_t k = a.pair_ptr[c]; k < k1; ++k) {
            const uint32_t b = a.pair_b[k];
            if (a.pair_npix[k] && a.kept[a.cls[b]]) uf_union(a.parent, (uint32_t)c, b);
        }
    }
    grid.sync();
    for (long c = gtid; c < NC; c += gsize) {
        uint32_t rf = 0;
        if (a.kept[a.cls[c]]) {
            const uint32_t root = uf_find(a.parent, (uint32_t)c);
            rf = root == (uint32_t)c ? 1u : 0u;
        }
        a.rootflag[c] = rf;
    }
    grid.sync();
    const uint32_t nlab = grid_excl_scan(grid, a.rootflag, NC, a.rank, a.blocksum, s_w);
    for (long c = gtid; c < NC; c += gsize)
        a.label[c] = a.kept[a.cls[c]] ? (int32_t)(__ldcg(a.rank + uf_find(a.parent, (uint32_t)c)) + 1u) : 0;
    for (long v = gtid; v <= (long)nlab; v += gsize) {
        a.bt0[v] = INT_MAX; a.bt1[v] = 0; a.by0[v] = INT_MAX; a.by1[v] = 0; a.bx0[v] = INT_MAX; a.bx1[v] = 0;
    }
    grid.sync();
    // ---- label boxes (find_objects before the date-line merge, contrack.py:753) ----
    for (long c = gtid; c < NC; c += gsize) {
        const int v = a.label[c];
        if (v == 0) continue;
        atomicMin(&a.bt0[v], a.comp_t[c]); atomicMax(&a.bt1[v], a.comp_t[c] + 1);
        atomicMin(&a.by0[v], a.comp_y0[c]); atomicMax(&a.by1[v], a.comp_y1[c]);
        atomicMin(&a.bx0[v], a.comp_x0[c]); atomicMax(&a.bx1[v], a.comp_x1[c]);
    }
    for (long s = gtid; s < NS; s += gsize) {
        const int la = a.label[a.seg_a[s]], lb = a.label[a.seg_b[s]];
        a.evflag[s] = (la != 0 && lb != 0 && la != lb) ? 1u : 0u;
    }
    grid.sync();
    // ---- persistence verdict of every label as if no date-line event touched it (contrack.py:765-772) ----
    uint32_t feats = 0;
    for (long v = gtid; v <= (long)nlab; v += gsize) {
        const bool keep = v > 0 && a.bt1[v] > a.bt0[v] && (a.bt1[v] - a.bt0[v]) >= a.persistence;
        a.fin[v] = keep ? (int32_t)v : 0;
        feats += keep;
    }
    feats = block_sum(feats, s_w);
    if (threadIdx.x == 0 && feats) atomicAdd(ctr + 3, feats);
    // ---- labels that occur in events -> compact records (label, box); events -> pairs of record indices, (t, y) order ----
    uint32_t* mark = a.rank;                                      // (rank and parent are free once the labels are known)
    uint32_t* lpos = a.parent;
    for (long v = gtid; v <= (long)nlab; v += gsize) mark[v] = 0;
    grid.sync();
    for (long s = gtid; s < NS; s += gsize) {
        if (!a.evflag[s]) continue;
        mark[a.label[a.seg_a[s]]] = 1u; mark[a.label[a.seg_b[s]]] = 1u;
    }
    grid.sync();
    const uint32_t nrec = grid_excl_scan(grid, mark, (long)nlab + 1, lpos, a.blocksum, s_w);
    for (long v = gtid; v <= (long)nlab; v += gsize) {
        if (!mark[v]) continue;
        int32_t* r = a.lrec + (size_t)lpos[v] * 7;
        r[0] = (int32_t)v; r[1] = a.bt0[v]; r[2] = a.bt1[v]; r[3] = a.by0[v]; r[4] = a.by1[v]; r[5] = a.bx0[v]; r[6] = a.bx1[v];
    }
    uint32_t* evpos = a.rootflag;                                  // (the root flags are no longer needed; NS <= capacity)
    const uint32_t nev = grid_excl_scan(grid, a.evflag, NS, evpos, a.blocksum, s_w);
    for (long s = gtid; s < NS; s += gsize) {
        if (!a.evflag[s]) continue;
        const uint32_t e = evpos[s];
        a.ev[2 * (size_t)e] = (int32_t)lpos[a.label[a.seg_a[s]]];
        a.ev[2 * (size_t)e + 1] = (int32_t)lpos[a.label[a.seg_b[s]]];
    }
    grid.sync();
    if (gtid == 0) {
        a.out8[0] = sweeps; a.out8[1] = 0; a.out8[2] = nlab; a.out8[3] = nev; a.out8[4] = __ldcg(ctr + 3);
        a.out8[5] = __ldcg(ctr + 4); a.out8[6] = nrec;
    }
}

}  // namespace

int global_grid(int sm_count) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_global_phase, GLOBAL_THREADS, 0) != cudaSuccess || per_sm < 1)
        return 0;
    return sm_count;                                               // one block per SM is plenty for table-sized work
}

cudaError_t global_phase(const GlobalArgs& a, int grid, cudaStream_t st) {
    GlobalArgs args = a;
    void* params[] = {&args};
    return cudaLaunchCooperativeKernel((const void*)k_global_phase, dim3((unsigned)grid), dim3(GLOBAL_THREADS), params, 0, st);
}

}  // namespace ctp
