// ct_lifecycle.cu -- sm_100a kernels of run_lifecycle (reference contrack/contrack.py:799-907): per (time step, flag id)
// area, area-weighted intensity and centre of mass of every flagged contour, from the int32 flag cube and the variable.
//
//   lc_rows      flag cube -> per row: non-zero bits, run-start bits (value changes), run count      4 B/cell read, HBM-bound
//   lc_extract   bit rows  -> row-runs (x0, x1, row, flag id) in raster order
//   lc_entries   row-runs  -> one entry per (time step, flag id): pixel count, touches-column-0 / column-W-1 bits
//   lc_roll      entries that touch both date-line columns: western edge = column after the widest gap (contrack.py:880-883)
//   lc_sums      per entry, one thread, in the REFERENCE'S summation orders (float64 adds are not associative):
//                  area = np.sum(w[mask]), np.sum(w[mask] * v[mask])       numpy pairwise order        (contrack.py:874-875)
//                  sum(w*v), sum(w*v*y), sum(w*v*x')                       sequential raster order of the rolled plane
//                                                                           (scipy.ndimage.center_of_mass -> np.bincount)
// Only cells of flagged contours are read from the variable (a few % of the cube).
#include "ct_kernels.h"

#include <algorithm>
#include <climits>

namespace ctl {

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t* total) {
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += n;
    }
    *total = __shfl_sync(FULL, inc, 31);
    return inc - v;
}

// One warp per row, 8 coalesced 128-byte loads in flight per lane.  A run starts where the value is non-zero and differs
// from its left neighbour (two different ids may touch: pieces of a feature that the date-line merge split).
__global__ void __launch_bounds__(256) k_lc_rows(const int32_t* __restrict__ flag, long nrows, int W, int Ww,
                                                 uint32_t* __restrict__ nz, uint32_t* __restrict__ st,
                                                 uint32_t* __restrict__ row_cnt) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long row = warp0; row < nrows; row += nwarps) {
        const int32_t* f = flag + row * (long)W;
        int32_t carry = 0;
        uint32_t cnt = 0;
        for (int k0 = 0; k0 < Ww; k0 += 32) {
            uint32_t mynz = 0, myst = 0;
            const int kend = min(32, Ww - k0);
            for (int j0 = 0; j0 < kend; j0 += 8) {
                int32_t v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int x = (k0 + j0 + j) * 32 + lane;
                    v[j] = (x < W) ? __ldcs(f + x) : 0;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    int32_t prev = __shfl_up_sync(FULL, v[j], 1);
                    if (lane == 0) prev = carry;
                    carry = __shfl_sync(FULL, v[j], 31);
                    const uint32_t a = __ballot_sync(FULL, v[j] != 0);
                    const uint32_t b = __ballot_sync(FULL, v[j] != 0 && v[j] != prev);
                    if (lane == j0 + j) { mynz = a; myst = b; }
                    cnt += __popc(b);
                }
            }
            if (lane < kend) { nz[row * (long)Ww + k0 + lane] = mynz; st[row * (long)Ww + k0 + lane] = myst; }
        }
        if (lane == 0) row_cnt[row] = cnt;
    }
}

// One warp per row with runs: a run ends at the next cell that is zero or starts another run.
__global__ void __launch_bounds__(256) k_lc_extract(const int32_t* __restrict__ flag, const uint32_t* __restrict__ nz,
                                                    const uint32_t* __restrict__ st,
                                                    const uint32_t* __restrict__ row_ptr, long nrows, int W, int Ww,
                                                    uint32_t* __restrict__ run_x, uint32_t* __restrict__ run_row,
                                                    int32_t* __restrict__ run_label) {
    extern __shared__ uint32_t lc_smem[];
    const int lane = threadIdx.x & 31;
    uint32_t* E = lc_smem + (size_t)(threadIdx.x >> 5) * Ww;       // end mask of the row this warp works on
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long row = warp0; row < nrows; row += nwarps) {
        const uint32_t base = row_ptr[row];
        if (row_ptr[row + 1] == base) continue;
        const uint32_t* rn = nz + row * (long)Ww;
        const uint32_t* rs = st + row * (long)Ww;
        __syncwarp();
        for (int k = lane; k < Ww; k += 32) E[k] = ~rn[k] | rs[k];
        __syncwarp();
        uint32_t sbase = base;
        for (int k0 = 0; k0 < Ww; k0 += 32) {
            const int k = k0 + lane;
            uint32_t s = k < Ww ? rs[k] : 0u;
            uint32_t tot;
            uint32_t idx = sbase + warp_excl_scan(__popc(s), lane, &tot);
            while (s) {
                const int b = __ffs(s) - 1;
                s &= s - 1;
                const int p = k * 32 + b;
                uint32_t m = b == 31 ? 0u : (E[k] & ~((2u << b) - 1u));
                int kk = k;
                while (!m && ++kk < Ww) m = E[kk];
                int x1 = m ? kk * 32 + __ffs(m) - 1 : W;
                if (x1 > W) x1 = W;
                run_x[idx] = (uint32_t)p | ((uint32_t)x1 << 16);
                run_row[idx] = (uint32_t)row;
                run_label[idx] = flag[row * (long)W + p];
                ++idx;
            }
            sbase += tot;
        }
    }
}

__device__ __forceinline__ uint32_t hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (uint32_t)k;
}

__global__ void __launch_bounds__(256) k_lc_table_init(EntryTable e) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.cap) return;
    e.key[i] = ctk::PAIR_EMPTY; e.npix[i] = 0; e.flags[i] = 0;
    if (i == 0) { *e.overflow = 0; *e.count = 0; *e.nroll = 0; }
}

__global__ void __launch_bounds__(256) k_lc_entries(const uint32_t* __restrict__ run_x, const uint32_t* __restrict__ run_row,
                                                    const int32_t* __restrict__ run_label, long nruns, int H, int W,
                                                    EntryTable e) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    const uint32_t x = run_x[r];
    const uint32_t x0 = x & 0xffff, x1 = x >> 16;
    const unsigned long long key = ((unsigned long long)(run_row[r] / (uint32_t)H) << 32) | (uint32_t)run_label[r];
    const uint32_t mask = e.cap - 1;
    uint32_t slot = hash64(key) & mask;
    for (uint32_t probe = 0;; ++probe) {
        if (probe > mask) { *e.overflow = 1; return; }
        unsigned long long k = *((volatile unsigned long long*)&e.key[slot]);
        if (k == ctk::PAIR_EMPTY) {
            k = atomicCAS(&e.key[slot], ctk::PAIR_EMPTY, key);
            if (k == ctk::PAIR_EMPTY) atomicAdd(e.count, 1u);
        }
        if (k == ctk::PAIR_EMPTY || k == key) break;
        slot = (slot + 1) & mask;
    }
    atomicAdd(&e.npix[slot], x1 - x0);
    const uint32_t fl = (x0 == 0 ? 1u : 0u) | (x1 == (uint32_t)W ? 2u : 0u);
    if (fl) atomicOr(&e.flags[slot], fl);
}

__global__ void __launch_bounds__(256) k_lc_compact(EntryTable e, int32_t* __restrict__ out_t, int32_t* __restrict__ out_label,
                                                    uint32_t* __restrict__ out_npix, int32_t* __restrict__ out_roll,
                                                    uint32_t* __restrict__ fill) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.cap) return;
    const unsigned long long k = e.key[i];
    if (k == ctk::PAIR_EMPTY) return;
    const uint32_t idx = atomicAdd(fill, 1u);
    out_t[idx] = (int32_t)(k >> 32); out_label[idx] = (int32_t)(uint32_t)k; out_npix[idx] = e.npix[i];
    // entries that touch both date-line columns get a slot for their column bitmap; out_roll holds -(slot + 2) until
    // k_lc_roll replaces it by the western edge; -1 = not rolled
    out_roll[idx] = e.flags[i] == 3u ? -(int32_t)(atomicAdd(e.nroll, 1u) + 2u) : -1;
}

// lon_roll = np.unique(xloc)[np.argmax(np.diff(np.unique(xloc))) + 1]   (contrack.py:882-883)
__global__ void __launch_bounds__(128) k_lc_roll(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ run_x,
                                                 const int32_t* __restrict__ run_label, const int32_t* __restrict__ ent_t,
                                                 const int32_t* __restrict__ ent_label, int32_t* __restrict__ ent_roll,
                                                 long nent, int H, int W, int Ww, uint32_t* __restrict__ bitmaps) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nent) return;
    const int code = ent_roll[i];
    if (code >= -1) return;
    if (W < 2) { ent_roll[i] = -1; return; }
    uint32_t* bm = bitmaps + (size_t)(-(code + 2)) * Ww;           // zeroed by the caller
    const int label = ent_label[i];
    const long t = ent_t[i];
    const uint32_t rb = row_ptr[t * H], re = row_ptr[(t + 1) * H];
    for (uint32_t r = rb; r < re; ++r) {
        if (run_label[r] != label) continue;
        const uint32_t x = run_x[r];
        const int x0 = x & 0xffff, x1 = x >> 16;                    // set bits [x0, x1)
        const int k0 = x0 >> 5, k1 = (x1 - 1) >> 5;
        const uint32_t m0 = ~0u << (x0 & 31), m1 = ~0u >> (31 - ((x1 - 1) & 31));
        if (k0 == k1) bm[k0] |= m0 & m1;
        else {
            bm[k0] |= m0;
            for (int k = k0 + 1; k < k1; ++k) bm[k] = ~0u;
            bm[k1] |= m1;
        }
    }
    int prev = -1, best = -1, bestcol = 0;
    for (int k = 0; k < Ww; ++k) {
        uint32_t word = bm[k];
        while (word) {
            const int col = k * 32 + __ffs(word) - 1;
            word &= word - 1;
            if (prev >= 0 && col - prev > best) { best = col - prev; bestcol = col; }
            prev = col;
        }
    }
    ent_roll[i] = bestcol;
}

// ---- per-entry sums ------------------------------------------------------------------------------------------------------
struct D2 { double a, b; };

template <typename TV>
struct PixelWalk {                       // cells of one (time step, id) in raster order
    const uint32_t* run_x; const int32_t* run_label; const uint32_t* run_row;
    const TV* var; const double* w;
    long r, rend; int label, H, W;
    int x, x1; long rowbase; double wy;
    __device__ __forceinline__ void next(double& wo, double& wv) {
        while (x >= x1) {                                          // the caller never asks for more cells than there are
            while (run_label[r] != label) ++r;
            const uint32_t xx = run_x[r];
            x = xx & 0xffff; x1 = xx >> 16;
            const uint32_t row = run_row[r];
            rowbase = (long)row * W; wy = w[row % (uint32_t)H];
            ++r;
        }
        wo = wy;
        wv = __dmul_rn(wy, (double)var[rowbase + x]);             // weight_grid[m] * variable[m]: float64 product
        ++x;
    }
};

template <typename TV>
__device__ D2 pw_leaf(PixelWalk<TV>& it, long n) {                // numpy pairwise_sum, n <= 128 (loops_utils.h.src)
    double w, wv;
    if (n < 8) {
        D2 res{0., 0.};
        for (long i = 0; i < n; ++i) { it.next(w, wv); res.a = __dadd_rn(res.a, w); res.b = __dadd_rn(res.b, wv); }
        return res;
    }
    D2 r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { it.next(w, wv); r[j].a = w; r[j].b = wv; }
    long i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { it.next(w, wv); r[j].a = __dadd_rn(r[j].a, w); r[j].b = __dadd_rn(r[j].b, wv); }
    }
    D2 res;
    res.a = __dadd_rn(__dadd_rn(__dadd_rn(r[0].a, r[1].a), __dadd_rn(r[2].a, r[3].a)),
                      __dadd_rn(__dadd_rn(r[4].a, r[5].a), __dadd_rn(r[6].a, r[7].a)));
    res.b = __dadd_rn(__dadd_rn(__dadd_rn(r[0].b, r[1].b), __dadd_rn(r[2].b, r[3].b)),
                      __dadd_rn(__dadd_rn(r[4].b, r[5].b), __dadd_rn(r[6].b, r[7].b)));
    for (; i < n; ++i) { it.next(w, wv); res.a = __dadd_rn(res.a, w); res.b = __dadd_rn(res.b, wv); }
    return res;
}

__device__ __forceinline__ long pw_half(long n) { long n2 = n / 2; return n2 - (n2 % 8); }

// the recursion pairwise(a, n2) + pairwise(a + n2, n - n2) with an explicit stack; leaves consume the cells in order
template <typename TV>
__device__ D2 pw_sum(PixelWalk<TV>& it, long n) {
    long fn[48]; int fstage[48]; D2 fleft[48];
    int sp = 0;
    fn[0] = n; fstage[0] = 0;
    D2 ret{0., 0.};
    while (sp >= 0) {
        const long m = fn[sp];
        if (m <= 128) {
            ret = pw_leaf(it, m);
            --sp;
            while (sp >= 0) {
                if (fstage[sp] == 1) {                             // left half done: run the right half
                    fleft[sp] = ret; fstage[sp] = 2;
                    const long right = fn[sp] - pw_half(fn[sp]);
                    ++sp; fn[sp] = right; fstage[sp] = 0;
                    break;
                }
                ret.a = __dadd_rn(fleft[sp].a, ret.a); ret.b = __dadd_rn(fleft[sp].b, ret.b);
                --sp;
            }
        } else {
            fstage[sp] = 1;
            const long left = pw_half(m);
            ++sp; fn[sp] = left; fstage[sp] = 0;
        }
    }
    ret.a = __dadd_rn(0., ret.a); ret.b = __dadd_rn(0., ret.b);   // np.add.reduce starts from the identity
    return ret;
}

template <typename TV>
__global__ void __launch_bounds__(128) k_lc_sums(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ run_x,
                                                 const uint32_t* __restrict__ run_row,
                                                 const int32_t* __restrict__ run_label, const TV* __restrict__ var,
                                                 const double* __restrict__ w, const int32_t* __restrict__ ent_t,
                                                 const int32_t* __restrict__ ent_label,
                                                 const uint32_t* __restrict__ ent_npix,
                                                 const int32_t* __restrict__ ent_roll, long nent, int H, int W,
                                                 double* __restrict__ out_area, double* __restrict__ out_int,
                                                 double* __restrict__ out_norm, double* __restrict__ out_sy,
                                                 double* __restrict__ out_sx) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nent) return;
    const int label = ent_label[i];
    const long t = ent_t[i];
    const uint32_t rb = row_ptr[t * H], re = row_ptr[(t + 1) * H];
    {   // contrack.py:874-875 -- np.sum over the boolean-masked cells in raster order
        PixelWalk<TV> it;
        it.run_x = run_x; it.run_label = run_label; it.run_row = run_row; it.var = var; it.w = w;
        it.r = rb; it.rend = re; it.label = label; it.H = H; it.W = W; it.x = 0; it.x1 = 0; it.rowbase = 0; it.wy = 0.;
        const D2 s = pw_sum(it, (long)ent_npix[i]);
        out_area[i] = s.a; out_int[i] = s.b;
    }
    // contrack.py:886 / 892 -- ndimage.center_of_mass(variable * weight_grid, flag, [label]) = three np.bincount sums,
    // i.e. sequential float64 accumulation over the cells in raster order of the plane rolled by -lon_roll columns
    const int roll = ent_roll[i] > 0 ? ent_roll[i] : 0;
    double sn = 0., sy = 0., sx = 0.;
    for (int y = 0; y < H; ++y) {
        const uint32_t r0 = row_ptr[t * H + y], r1 = row_ptr[t * H + y + 1];
        if (r0 == r1) continue;
        const double wy = w[y], fy = (double)y;
        const TV* vrow = var + (t * H + y) * (long)W;
        for (int phase = 0; phase < (roll ? 2 : 1); ++phase) {
            for (uint32_t r = r0; r < r1; ++r) {
                if (run_label[r] != label) continue;
                const uint32_t xx = run_x[r];
                int x0 = xx & 0xffff, x1 = xx >> 16, shift;
                if (phase == 0) { x0 = max(x0, roll); shift = -roll; }        // columns roll .. W-1 come first
                else { x1 = min(x1, roll); shift = W - roll; }                 // then columns 0 .. roll-1
                for (int x = x0; x < x1; ++x) {
                    const double in = __dmul_rn((double)vrow[x], wy);
                    sn = __dadd_rn(sn, in);
                    sy = __dadd_rn(sy, __dmul_rn(in, fy));
                    sx = __dadd_rn(sx, __dmul_rn(in, (double)(x + shift)));
                }
            }
        }
    }
    out_norm[i] = sn; out_sy[i] = sy; out_sx[i] = sx;
}

inline unsigned blocks_for(long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

cudaError_t lc_rows(const int32_t* flag, long nrows, int W, int Ww, uint32_t* nz, uint32_t* st, uint32_t* row_cnt,
                    int sm_count, cudaStream_t stream) {
    if (nrows == 0) return cudaSuccess;
    const long want = (nrows + 7) / 8;
    k_lc_rows<<<(unsigned)std::min<long>(want, (long)sm_count * 8), 256, 0, stream>>>(flag, nrows, W, Ww, nz, st, row_cnt);
    return cudaGetLastError();
}

cudaError_t lc_extract(const int32_t* flag, const uint32_t* nz, const uint32_t* st, const uint32_t* row_ptr, long nrows,
                       int W, int Ww, uint32_t* run_x, uint32_t* run_row, int32_t* run_label, cudaStream_t stream) {
    if (nrows == 0) return cudaSuccess;
    const size_t smem = (size_t)8 * Ww * sizeof(uint32_t);
    cudaError_t e = cudaFuncSetAttribute(k_lc_extract, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long want = (nrows + 7) / 8;
    k_lc_extract<<<(unsigned)std::min<long>(want, 148L * 16), 256, smem, stream>>>(flag, nz, st, row_ptr, nrows, W, Ww, run_x,
                                                                                 run_row, run_label);
    return cudaGetLastError();
}

cudaError_t lc_entries(const uint32_t* run_x, const uint32_t* run_row, const int32_t* run_label, long nruns, int H, int W,
                       const EntryTable& e, cudaStream_t stream) {
    k_lc_table_init<<<blocks_for(e.cap, 256), 256, 0, stream>>>(e);
    if (nruns) k_lc_entries<<<blocks_for(nruns, 256), 256, 0, stream>>>(run_x, run_row, run_label, nruns, H, W, e);
    return cudaGetLastError();
}

cudaError_t lc_compact(const EntryTable& e, int32_t* out_t, int32_t* out_label, uint32_t* out_npix, int32_t* out_roll,
                       uint32_t* fill_zeroed, cudaStream_t stream) {
    k_lc_compact<<<blocks_for(e.cap, 256), 256, 0, stream>>>(e, out_t, out_label, out_npix, out_roll, fill_zeroed);
    return cudaGetLastError();
}

cudaError_t lc_roll(const uint32_t* row_ptr, const uint32_t* run_x, const int32_t* run_label, const int32_t* ent_t,
                    const int32_t* ent_label, int32_t* ent_roll, long nent, int H, int W, int Ww, uint32_t* bitmaps,
                    cudaStream_t stream) {
    if (nent == 0) return cudaSuccess;
    k_lc_roll<<<blocks_for(nent, 128), 128, 0, stream>>>(row_ptr, run_x, run_label, ent_t, ent_label, ent_roll, nent, H, W,
                                                        Ww, bitmaps);
    return cudaGetLastError();
}

cudaError_t lc_sums(const uint32_t* row_ptr, const uint32_t* run_x, const uint32_t* run_row, const int32_t* run_label,
                    const void* var, int var_is_f64, const double* w, const int32_t* ent_t, const int32_t* ent_label,
                    const uint32_t* ent_npix, const int32_t* ent_roll, long nent, int H, int W, double* out_area,
                    double* out_int, double* out_norm, double* out_sy, double* out_sx, cudaStream_t stream) {
    if (nent == 0) return cudaSuccess;
    if (var_is_f64)
        k_lc_sums<double><<<blocks_for(nent, 128), 128, 0, stream>>>(row_ptr, run_x, run_row, run_label, (const double*)var, w,
                                                                     ent_t, ent_label, ent_npix, ent_roll, nent, H, W,
                                                                     out_area, out_int, out_norm, out_sy, out_sx);
    else
        k_lc_sums<float><<<blocks_for(nent, 128), 128, 0, stream>>>(row_ptr, run_x, run_row, run_label, (const float*)var, w,
                                                                    ent_t, ent_label, ent_npix, ent_roll, nent, H, W,
                                                                    out_area, out_int, out_norm, out_sy, out_sx);
    return cudaGetLastError();
}

}  // namespace ctl
