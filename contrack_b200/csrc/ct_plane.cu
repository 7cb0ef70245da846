// ct_plane.cu -- the per-plane table kernel: ONE thread block owns one time plane and turns the row-runs the threshold kernel
// left behind into every table the ordered phase needs, in shared memory, in one launch per time chunk.
//
//   contrack.py:684-687   2-D 8-connected components = union-find over the plane's row-runs in SHARED memory
//                         (16-bit parents, smallest run index is the root, so components come out in first-pixel order)
//   contrack.py:691-698   same-row date-line classes (a second union-find over the plane's components) and the date-line
//                         rows grouped into segments
//   contrack.py:703-704, 717-719   area per component / class and the (component at t, component at t-1) pair table with
//                         the overlap areas: a shared-memory hash per plane, written out as CSR
//
// Global numbering without a second pass: run / component / segment / pair indices of a plane start where the previous
// plane's end, obtained by a decoupled look-back over per-plane descriptors (aggregate first, inclusive prefix as soon as a
// predecessor's is known).  Blocks take their plane from an atomic ticket, so every predecessor a block waits for is
// already running.  Nothing is sized by the host between kernels: capacities are checked on the device and reported in
// `status`; the true totals always reach the chain, so a retry knows the exact sizes.
//
// A plane whose runs do not fit the shared-memory budget (or whose pair hash overflows) sets ST_FALLBACK; the caller then
// rebuilds the tables with the global-memory kernels of ct_kernels.cu (same tables, bit for bit).
#include "ct_plane.h"

#include <climits>

namespace ctp {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned long long HASH_EMPTY = ~0ull;

struct HashEntry {                 // 32 bytes
    unsigned long long key;        // (local component of this plane << 32) | global component of the previous plane
    uint32_t npix, nsp;
    double E, S;
};

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- block-wide helpers (blockDim.x threads, a multiple of 32, at most 1024) ----
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += n;
    }
    return v;
}
// exclusive prefix of v over the block; *total = block sum.  s_w: 33 words of shared scratch.  Ends with a barrier.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_w, uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const uint32_t x = lane < nw ? s_w[lane] : 0;
        const uint32_t xi = warp_incl_scan(x, lane);
        if (lane < nw) s_w[lane] = xi - x;
        if (lane == 31) s_w[32] = xi;
    }
    __syncthreads();
    const uint32_t ex = inc - v + s_w[wid];
    *total = s_w[32];
    __syncthreads();
    return ex;
}

// ---- union-find over 16-bit parents in shared memory (lock-free, smallest index is the root) ----
__device__ __forceinline__ uint32_t uf_find(const volatile uint16_t* par, uint32_t x) {
    while (true) {
        const uint32_t q = par[x];
        if (q == x) return x;
        x = q;
    }
}
__device__ __forceinline__ void uf_union(uint16_t* par, uint32_t a, uint32_t b) {
    while (true) {
        a = uf_find(par, a);
        b = uf_find(par, b);
        if (a == b) return;
        if (a < b) { const uint32_t t = a; a = b; b = t; }          // a > b: hang a below b
        const unsigned short old = atomicCAS(reinterpret_cast<unsigned short*>(par) + a, (unsigned short)a, (unsigned short)b);
        if (old == (unsigned short)a) return;                       // else: somebody re-parented a meanwhile, try again
    }
}

// ---- decoupled look-back (one warp): publishes `agg` for slot `slot`, returns the sum of all earlier slots ----
// word = value << 2 | flag; flag 1 = aggregate of this slot only, 2 = inclusive prefix.  Slot 0 is the sentinel the
// host wrote (inclusive 0, or the totals the previous chunks reached: slots of earlier launches are all inclusive).
// A predecessor that never publishes (which would be a bug: blocks take their plane from a ticket, so every predecessor is
// already running) is reported in *status after SPIN_LIMIT polls instead of hanging the device.
constexpr uint32_t SPIN_LIMIT = 1u << 25;
__device__ __forceinline__ unsigned long long lookback(unsigned long long* words, long slot, unsigned long long agg, int lane,
                                                       uint32_t* status) {
    if (lane == 0) st_release(words + slot, (agg << 2) | 1ull);
    unsigned long long excl = 0;
    long idx = slot - 1;
    while (true) {
        const long j = idx - lane;
        unsigned long long w = 2ull;                                // before slot 0: inclusive zero
        if (j >= 0) {
            uint32_t spins = 0;
            do {
                w = ld_acquire(words + j);
                if (++spins == SPIN_LIMIT) { atomicOr(status, ST_TIMEOUT); w = 2ull; }
            } while ((w & 3ull) == 0ull);
        }
        const unsigned inc_mask = __ballot_sync(FULL, (w & 3ull) == 2ull);
        const int first = inc_mask ? __ffs(inc_mask) - 1 : 32;     // nearest slot that already knows its inclusive prefix
        unsigned long long v = (lane <= first) ? (w >> 2) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
        excl += v;
        if (inc_mask) break;
        idx -= 32;
    }
    if (lane == 0) st_release(words + slot, ((excl + agg) << 2) | 2ull);
    return excl;
}

// row-runs of one bit row, serially (rows with more runs than the threshold kernel's slots: rare)
__device__ void runs_from_bits(const uint32_t* __restrict__ b, int W, int Ww, uint32_t* __restrict__ out, uint32_t cap) {
    bool in_run = false;
    uint32_t x0 = 0, n = 0;
    for (int k = 0; k < Ww; ++k) {
        const uint32_t m = __ldcg(b + k);
        int pos = 0;
        while (pos < 32) {
            if (!in_run) {
                const uint32_t r = m >> pos;
                if (r == 0) break;
                pos += __ffs(r) - 1;
                x0 = (uint32_t)(k * 32 + pos);
                in_run = true;
            } else {
                const uint32_t r = (~m) >> pos;                      // zeros shifted in from the top never look like an end
                const uint32_t valid = pos == 0 ? FULL : ((1u << (32 - pos)) - 1u);
                const uint32_t z = r & valid;
                if (z == 0) break;                                    // the run reaches the end of this word
                pos += __ffs(z) - 1;
                if (n < cap) out[n] = x0 | ((uint32_t)(k * 32 + pos) << 16);
                ++n;
                in_run = false;
            }
        }
    }
    if (in_run && n < cap) out[n] = x0 | ((uint32_t)W << 16);
}

// row of local run i: last y with roff[y] <= i
__device__ __forceinline__ int row_of(const uint16_t* roff, int H, uint32_t i) {
    int lo = 0, hi = H - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (roff[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(PLANE_THREADS) k_plane_tables(PlaneArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;
    const int H = a.H, W = a.W;
    const uint32_t Rcap = a.smem_runs;
    // ---- shared-memory carve-up ----
    uint32_t* sx = reinterpret_cast<uint32_t*>(smem);                               // [Rcap]   x0 | x1 << 16
    uint16_t* par = reinterpret_cast<uint16_t*>(sx + Rcap);                         // [Rcap]   union-find parents; later cls
    uint16_t* kid = par + Rcap;                                                     // [Rcap]   local component of a run
    uint32_t* sab = reinterpret_cast<uint32_t*>(kid + Rcap);                        // [H]      date-line row: a | b << 16
    uint16_t* roff = reinterpret_cast<uint16_t*>(sab + H);                          // [H + 1]  first run of every row
    uint16_t* spos = roff + (H + 2);                                                // [H]      segment index of a row
    uint32_t* s_w = reinterpret_cast<uint32_t*>(smem + a.smem_scan_off);            // [40]     scan scratch + broadcasts
    HashEntry* hash = reinterpret_cast<HashEntry*>(par);                            // pair hash re-uses the parents' space
    const uint32_t HC = a.hash_cap;                                                 // power of two, HC * 32 <= Rcap * 2

    __shared__ long s_plane;
    if (tid == 0) s_plane = a.p0 + (long)atomicAdd(a.ticket, 1u);
    __syncthreads();
    const long plane = s_plane;                                   // plane of the context's scratch (a halo plane counts)
    const long row0 = plane * H;
    const long slot = plane + 1;                                  // chain slot (slot 0 = sentinel)
    unsigned long long* chainA = a.chain;                          // value = components << 31 | segments
    unsigned long long* chainR = a.chain + a.chain_stride;         // runs
    unsigned long long* chainP = a.chain + 2 * a.chain_stride;     // pairs

    // ---- A. runs per row -> local offsets ----
    uint32_t carry = 0;
    for (int y0 = 0; y0 < H; y0 += nthr) {
        const int y = y0 + tid;
        const uint32_t c = y < H ? __ldcg(a.row_cnt + row0 + y) : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(c, s_w, &tot);
        if (y < H) roff[y] = (uint16_t)min(carry + ex, 0xffffu);
        carry += tot;
    }
    const uint32_t n_true = carry;                                 // runs of this plane
    const bool too_big = n_true > Rcap || n_true > 0xfff0u;
    const uint32_t n = too_big ? 0u : n_true;
    if (tid == 0) { roff[H] = (uint16_t)n; if (too_big) atomicOr(a.status, ST_FALLBACK); }
    __syncthreads();

    // ---- B. runs into shared memory (slots of the threshold kernel; bit rows for rows with more runs than slots) ----
    if (!too_big) {
        for (int y = tid; y < H; y += nthr) {
            const uint32_t off = roff[y], c = roff[y + 1] - off;
            if (c == 0) continue;
            if (c <= (uint32_t)RUN_SLOTS) {
                const uint4* s4 = reinterpret_cast<const uint4*>(a.slots + (row0 + y) * (long)RUN_SLOTS);
                const uint4 u = __ldcg(s4);
                const uint32_t v0[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) if ((uint32_t)i < c) sx[off + i] = v0[i];
                if (c > 4) {
                    const uint4 v = __ldcg(s4 + 1);
                    const uint32_t v1[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) if ((uint32_t)(4 + i) < c) sx[off + 4 + i] = v1[i];
                }
            } else {
                runs_from_bits(a.bits + (row0 + y) * (long)a.Ww, W, a.Ww, sx + off, c);
                *a.info = 1u;
            }
        }
    }
    for (uint32_t i = tid; i < n; i += nthr) par[i] = (uint16_t)i;
    __syncthreads();

    // ---- C. 8-connectivity with the row above: [x0 - 1, x1 + 1) must meet [px0, px1) ----
    for (uint32_t i = tid; i < n; i += nthr) {
        const int y = row_of(roff, H, i);
        if (y == 0) continue;
        const uint32_t x = sx[i];
        const int x0 = x & 0xffff, x1 = x >> 16;
        uint32_t lo = roff[y - 1], hi = roff[y];
        const uint32_t end = hi;
        while (lo < hi) {                                            // first run above with px1 >= x0
            const uint32_t mid = (lo + hi) >> 1;
            if ((int)(sx[mid] >> 16) >= x0) hi = mid; else lo = mid + 1;
        }
        for (uint32_t p = lo; p < end; ++p) {
            if ((int)(sx[p] & 0xffff) > x1) break;
            uf_union(par, i, p);
        }
    }
    __syncthreads();
    // ---- D. roots -> local component index in raster order ----
    // The roots go to `kid` (a read-only walk over `par`: flattening `par` in place while other threads still walk it would be
    // harmless -- every value ever stored is an ancestor -- but it is a data race by the letter, and racecheck says so);
    // the root positions of `par`, free from then on, take the component indices.
    for (uint32_t i = tid; i < n; i += nthr) kid[i] = (uint16_t)uf_find(par, i);
    __syncthreads();
    uint32_t nC = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += nthr) {
        const uint32_t i = i0 + tid;
        const uint32_t isroot = (i < n && kid[i] == i) ? 1u : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(isroot, s_w, &tot);
        if (isroot) par[i] = (uint16_t)(nC + ex);
        nC += tot;
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += nthr) kid[i] = par[kid[i]];  // (own entry of `kid`, root entries of `par`)
    __syncthreads();

    // ---- E. date-line rows: classes over the plane's components, segments of consecutive rows ----
    uint16_t* cls = par;                                            // [nC], the run parents are no longer needed
    for (uint32_t k = tid; k < nC; k += nthr) cls[k] = (uint16_t)k;
    __syncthreads();
    for (int y = tid; y < H; y += nthr) {
        uint32_t ab = 0xffffffffu;
        if (!too_big && roff[y + 1] > roff[y] && __ldcg(a.seam_flag + row0 + y)) {
            const uint32_t ca = kid[roff[y]], cb = kid[roff[y + 1] - 1];
            ab = ca | (cb << 16);
            if (ca != cb) uf_union(cls, ca, cb);
        }
        sab[y] = ab;
    }
    __syncthreads();
    uint32_t nS = 0;                                                 // (class representatives: uf_find(cls, k) where they are used)
    for (int y0 = 0; y0 < H; y0 += nthr) {
        const int y = y0 + tid;
        uint32_t st = 0;
        if (y < H && sab[y] != 0xffffffffu) st = (y == 0 || sab[y - 1] != sab[y]) ? 1u : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(st, s_w, &tot);
        if (y < H) spos[y] = (uint16_t)(nS + ex + st - 1u);         // segment of row y (meaningful on date-line rows)
        nS += tot;
    }

    // ---- chain 1: global base of this plane's runs / components / segments ----
    if (tid < 32) {                                               // (two warps, two chains: the look-backs run side by side)
        const unsigned long long eA = lookback(chainA, slot, ((unsigned long long)nC << 31) | nS, lane, a.status);
        if (lane == 0) { s_w[34] = (uint32_t)(eA >> 31); s_w[35] = (uint32_t)(eA & 0x7fffffffu); }
    } else if (tid < 64) {
        const unsigned long long eR = lookback(chainR, slot, n_true, lane, a.status);
        if (lane == 0) { s_w[36] = (uint32_t)eR; s_w[37] = (uint32_t)(eR >> 32); }
    }
    __syncthreads();
    const uint32_t baseC = s_w[34], baseS = s_w[35];
    const unsigned long long baseR64 = (unsigned long long)s_w[36] | ((unsigned long long)s_w[37] << 32);
    const uint32_t baseR = (uint32_t)baseR64;
    bool ok = !too_big;
    if (baseR64 + n_true > (unsigned long long)a.cap_runs || (unsigned long long)baseC + nC > a.cap_comps ||
        (unsigned long long)baseS + nS > a.cap_segs) {
        ok = false;
        if (tid == 0) atomicOr(a.status, ST_CAPACITY);
    }

    // ---- F. this plane's part of the global tables ----
    if (ok) {
        for (int y = tid; y <= H; y += nthr) a.row_ptr[row0 + y] = baseR + roff[y];
        for (uint32_t i = tid; i < n; i += nthr) {
            a.run_x[baseR + i] = sx[i];
            a.run_row[baseR + i] = (uint32_t)(row0 + row_of(roff, H, i));
            a.run_comp[baseR + i] = baseC + kid[i];
        }
        for (uint32_t k = tid; k < nC; k += nthr) {
            const uint32_t c = baseC + k;
            a.ct.t[c] = (int32_t)plane; a.ct.y0[c] = INT_MAX; a.ct.y1[c] = 0; a.ct.x0[c] = W; a.ct.x1[c] = 0;
            a.ct.areaE[c] = 0.0; a.ct.areaS[c] = 0.0; a.ct.nsp[c] = 0; a.ct.cls[c] = baseC + uf_find(cls, k);
            a.kt.conE[c] = 0.0; a.kt.conS[c] = 0.0; a.kt.fE[c] = 0.0; a.kt.fS[c] = 0.0; a.kt.nsp[c] = 0; a.kt.fnsp[c] = 0;
            a.pcnt[c] = 0; a.pfill[c] = 0;
        }
        for (int y = tid; y < H; y += nthr) {
            const uint32_t ab = sab[y];
            if (ab == 0xffffffffu) continue;
            const uint32_t s = baseS + spos[y];
            if (y == 0 || sab[y - 1] != ab) {
                a.sg.t[s] = (int32_t)plane; a.sg.y0[s] = y; a.sg.a[s] = baseC + (ab & 0xffff); a.sg.b[s] = baseC + (ab >> 16);
            }
            if (y == H - 1 || sab[y + 1] != ab) a.sg.y1[s] = y + 1;
        }
    }
    __syncthreads();
    if (ok) {
        for (uint32_t i = tid; i < n; i += nthr) {
            const int y = row_of(roff, H, i);
            const uint32_t x = sx[i], c = baseC + kid[i];
            const int x0 = x & 0xffff, x1 = x >> 16;
            atomicMin(&a.ct.y0[c], y); atomicMax(&a.ct.y1[c], y + 1);
            atomicMin(&a.ct.x0[c], x0); atomicMax(&a.ct.x1[c], x1);
            const double area = (double)(x1 - x0) * a.w[y];          // exact: < 2^16 times a float32-valued double
            if (a.special[y]) { atomicAdd(&a.ct.areaS[c], area); atomicAdd(&a.ct.nsp[c], (uint32_t)(x1 - x0)); }
            else atomicAdd(&a.ct.areaE[c], area);
        }
    }
    __syncthreads();
    if (ok) {
        for (uint32_t k = tid; k < nC; k += nthr) {
            const uint32_t c = baseC + k, rep = baseC + uf_find(cls, k);
            atomicAdd(&a.kt.conE[rep], __ldcg(&a.ct.areaE[c]));
            const uint32_t ns = __ldcg(&a.ct.nsp[c]);
            if (ns) { atomicAdd(&a.kt.conS[rep], __ldcg(&a.ct.areaS[c])); atomicAdd(&a.kt.nsp[rep], ns); }
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release32(a.done + slot, 1u);                  // runs, components and classes of this plane are out

    // ---- G. pairs with the previous plane: shared-memory hash -> CSR over this plane's components ----
    const bool has_prev = plane > 0;
    uint32_t nP = 0;
    if (has_prev) {
        if (tid == 0) {
            uint32_t spins = 0;
            while (ld_acquire32(a.done + slot - 1) == 0u)
                if (++spins == SPIN_LIMIT) { atomicOr(a.status, ST_TIMEOUT | ST_FALLBACK); break; }
        }
        for (uint32_t h = tid; h < HC; h += nthr) { hash[h].key = HASH_EMPTY; hash[h].npix = 0; hash[h].nsp = 0; hash[h].E = 0.0; hash[h].S = 0.0; }
        if (tid == 0) { s_w[38] = 0; s_w[39] = 0; }
        __syncthreads();
        const bool prev_ok = (__ldcg(a.status) & (ST_FALLBACK | ST_CAPACITY | ST_TIMEOUT)) == 0u;   // (tables of earlier planes exist)
        if (ok && prev_ok) {
            for (uint32_t i = tid; i < n; i += nthr) {
                const int y = row_of(roff, H, i);
                const long prow = row0 - H + y;
                const uint32_t x = sx[i];
                const int x0 = x & 0xffff, x1 = x >> 16;
                uint32_t lo = __ldcg(a.row_ptr + prow), hi = __ldcg(a.row_ptr + prow + 1);
                const uint32_t end = hi;
                while (lo < hi) {                                    // first run of the earlier plane with px1 > x0
                    const uint32_t mid = (lo + hi) >> 1;
                    if ((int)(__ldcg(a.run_x + mid) >> 16) > x0) hi = mid; else lo = mid + 1;
                }
                const double wy = a.w[y];
                const bool sp = a.special[y] != 0;
                const unsigned long long ka = (unsigned long long)kid[i] << 32;
                for (uint32_t p = lo; p < end; ++p) {
                    const uint32_t px = __ldcg(a.run_x + p);
                    const int px0 = px & 0xffff, px1 = px >> 16;
                    if (px0 >= x1) break;
                    const int npx = min(x1, px1) - max(x0, px0);
                    const unsigned long long key = ka | __ldcg(a.run_comp + p);
                    uint32_t h = (uint32_t)((key * 0x9e3779b97f4a7c15ull) >> 40) & (HC - 1);
                    bool found = false;
                    for (uint32_t probe = 0; probe < HC; ++probe) {
                        unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&hash[h].key);
                        if (k == HASH_EMPTY) {
                            k = atomicCAS(&hash[h].key, HASH_EMPTY, key);
                            if (k == HASH_EMPTY) atomicAdd(&s_w[38], 1u);
                        }
                        if (k == HASH_EMPTY || k == key) { found = true; break; }
                        h = (h + 1) & (HC - 1);
                    }
                    if (!found) { s_w[39] = 1u; break; }
                    atomicAdd(&hash[h].npix, (uint32_t)npx);
                    const double area = (double)npx * wy;
                    if (sp) { atomicAdd(&hash[h].S, area); atomicAdd(&hash[h].nsp, (uint32_t)npx); }
                    else atomicAdd(&hash[h].E, area);
                }
            }
        }
        __syncthreads();
        nP = s_w[38];
        if (s_w[39] || nP * 4 > HC * 3) {                            // hash full (or crawling): rebuild with the global path
            if (tid == 0) atomicOr(a.status, ST_FALLBACK);
            nP = 0; ok = false;
        }
    }
    // ---- chain 2: global base of this plane's pairs; CSR ----
    __syncthreads();
    if (tid < 32) {
        const unsigned long long eP = lookback(chainP, slot, nP, lane, a.status);
        if (lane == 0) { s_w[36] = (uint32_t)eP; s_w[37] = (uint32_t)(eP >> 32); }
    }
    __syncthreads();
    const unsigned long long baseP64 = (unsigned long long)s_w[36] | ((unsigned long long)s_w[37] << 32);
    if (baseP64 + nP > (unsigned long long)a.cap_pairs) {
        ok = false;
        if (tid == 0) atomicOr(a.status, ST_CAPACITY);
    }
    const uint32_t baseP = (uint32_t)baseP64;
    if (ok) {
        if (has_prev) {
            for (uint32_t h = tid; h < HC; h += nthr)
                if (hash[h].key != HASH_EMPTY) atomicAdd(&a.pcnt[baseC + (uint32_t)(hash[h].key >> 32)], 1u);
        }
        __syncthreads();
        uint32_t run = 0;
        for (uint32_t k0 = 0; k0 < nC; k0 += nthr) {
            const uint32_t k = k0 + tid;
            const uint32_t c = k < nC ? __ldcg(&a.pcnt[baseC + k]) : 0u;
            uint32_t tot;
            const uint32_t ex = block_excl_scan(c, s_w, &tot);
            if (k < nC) a.pptr[baseC + k] = baseP + run + ex;
            run += tot;
        }
        if (tid == 0) a.pptr[baseC + nC] = baseP + nP;               // (the next plane writes the same value)
        __syncthreads();
        if (has_prev) {
            for (uint32_t h = tid; h < HC; h += nthr) {
                const HashEntry e = hash[h];
                if (e.key == HASH_EMPTY) continue;
                const uint32_t ca = baseC + (uint32_t)(e.key >> 32), cb = (uint32_t)e.key;
                const uint32_t pos = __ldcg(&a.pptr[ca]) + atomicAdd(&a.pfill[ca], 1u);
                a.pc.b[pos] = cb; a.pc.npix[pos] = e.npix; a.pc.nsp[pos] = e.nsp; a.pc.E[pos] = e.E; a.pc.S[pos] = e.S;
                // forward overlap of the earlier plane's classes (contrack.py:718: all of plane t+1, unfiltered)
                const uint32_t rep = __ldcg(&a.ct.cls[cb]);
                atomicAdd(&a.kt.fE[rep], e.E);
                if (e.nsp) { atomicAdd(&a.kt.fS[rep], e.S); atomicAdd(&a.kt.fnsp[rep], e.nsp); }
            }
        }
    }
}

}  // namespace

size_t plane_smem_bytes(int H, uint32_t smem_runs, size_t* scan_off) {
    size_t o = (size_t)smem_runs * 4 + (size_t)smem_runs * 2 * 2;       // sx, par, kid
    o += (size_t)H * 4;                                                 // sab
    o += (size_t)(H + 2) * 2 + (size_t)H * 2;                           // roff, spos
    o = (o + 15) / 16 * 16;
    if (scan_off) *scan_off = o;
    return o + 40 * 4;
}

// largest run capacity (a multiple of 16 with a power-of-two hash in the parents' space) that fits `budget` bytes
bool plane_config(int H, size_t budget, uint32_t* smem_runs, uint32_t* hash_cap) {
    const size_t fixed = plane_smem_bytes(H, 0, nullptr) + 64;
    if (fixed + 8 * 512 > budget) return false;
    size_t r = (budget - fixed) / 8;
    if (r > 0xfff0u) r = 0xfff0u;
    r = r / 16 * 16;
    uint32_t hc = 1;
    while ((size_t)hc * 2 * 32 <= r * 2) hc <<= 1;                     // hc * 32 bytes <= r * 2 bytes
    if (hc < 16) return false;
    *smem_runs = (uint32_t)r; *hash_cap = hc;
    return true;
}

cudaError_t plane_tables(const PlaneArgs& a, size_t smem_bytes, cudaStream_t st) {
    if (a.np <= 0) return cudaSuccess;
    static size_t configured = 0;
    if (smem_bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_plane_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        configured = smem_bytes;
    }
    k_plane_tables<<<(unsigned)a.np, PLANE_THREADS, smem_bytes, st>>>(a);
    return cudaGetLastError();
}

}  // namespace ctp
