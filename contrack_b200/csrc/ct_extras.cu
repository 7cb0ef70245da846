// ct_extras.cu -- the callers either side of run_contrack that README.rst shows:
//   quantile_time     README.rst:150-151  ds[var].sel(lat band).quantile(q, dim='time')  (numpy nanquantile, 'linear'); float32
//                                         or float64 cubes, optionally time-sharded over several ranks (exact: the per-point
//                                         tallies of the radix select are summed over the ranks between the passes)
//   flag_count        README.rst:161      xr.where(flag > v, 1, 0).sum(dim='time')       (blocking frequency numerator)
//   divide_f32        contrack.py:417-419 geopotential -> geopotential height, float32 division by g
//   gather_planes     contrack.py:565     clim.reindex(lat, lon, method='nearest') with index maps computed by the caller
// All four are HBM-bound streaming kernels: a thread owns one grid point (or four) and walks the time axis; neighbouring
// threads read neighbouring cells, so every warp load is one coalesced 128-byte (or 512-byte) request.
#include "ct_extras.h"

#include <cmath>

namespace cte {

namespace {

// order-preserving maps float -> unsigned key (ascending; -0.0 sorts before +0.0)
template <typename S> struct Key;
template <> struct Key<float> {
    using K = uint32_t;
    static constexpr int BITS = 32;
    static __device__ __forceinline__ K of(float v) { const uint32_t b = __float_as_uint(v); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
    static __device__ __forceinline__ float back(K k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
};
template <> struct Key<double> {
    using K = unsigned long long;
    static constexpr int BITS = 64;
    static __device__ __forceinline__ K of(double v) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v);
        return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    }
    static __device__ __forceinline__ double back(K k) {
        return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
    }
};

// ---- quantile over time as a sequence of passes over the (local) time steps; between the passes the per-point tallies can
// be summed over the ranks of a time-sharded cube, so every rank selects the same order statistics of the WHOLE cube:
//   count     n[p]         = values that are not NaN
//   init      numpy's virtual index  (n - 1) * q  ->  lo = floor, gamma = fraction; k = lo (rank still to find)
//   hist      h[iq][d][p]  = values whose key matches the prefix found so far and whose next 4-bit digit is d
//   select    the digit that holds rank k; k -= values in smaller digits; prefix |= digit      (8 / 16 rounds: exact radix select)
//   neigh     le = values <= the selected one, best = smallest key above it    ->  the (lo+1)-th smallest value
//   finish    numpy's linear interpolation  a + (b - a) * gamma  /  b - (b - a) * (1 - gamma) for gamma >= 0.5, in numpy's
//             operation order (b - a in the input precision, the rest in float64: numpy/lib/_function_base_impl.py:_lerp)
// One thread per grid point of rows [y0, y1); the reads of a warp are contiguous in x.
template <typename S>
__global__ void __launch_bounds__(128) k_q_count(const S* __restrict__ x, long T, long HW, int W, int y0, long npts,
                                                 uint32_t* __restrict__ cnt) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const S* col = x + (long)y0 * W + p;
    uint32_t n = 0;
    long t = 0;
    for (; t + 4 <= T; t += 4) {
        const S a = __ldg(col + t * HW), b = __ldg(col + (t + 1) * HW), c = __ldg(col + (t + 2) * HW), d = __ldg(col + (t + 3) * HW);
        n += (a == a) + (b == b) + (c == c) + (d == d);
    }
    for (; t < T; ++t) { const S a = __ldg(col + t * HW); n += (a == a); }
    cnt[p] = n;
}

template <typename KT>
__global__ void __launch_bounds__(128) k_q_init(const uint32_t* __restrict__ cnt, long npts, const double* __restrict__ q, int nq,
                                                uint32_t* __restrict__ lo_out, double* __restrict__ gamma_out,
                                                uint32_t* __restrict__ k_out, KT* __restrict__ prefix) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const long n = cnt[p];
    for (int iq = 0; iq < nq; ++iq) {
        const long i = (long)iq * npts + p;
        long lo = 0;
        double gamma = 0.0;
        if (n > 0) {
            const double virt = __dmul_rn((double)(n - 1), q[iq]);
            if (!(virt < (double)(n - 1))) { lo = n - 1; gamma = 0.0; }           // numpy: index -1 for both neighbours
            else if (virt < 0.0) { lo = 0; gamma = 0.0; }
            else { const double fl = floor(virt); lo = (long)fl; gamma = __dsub_rn(virt, fl); }
        }
        lo_out[i] = (uint32_t)lo; gamma_out[i] = gamma; k_out[i] = (uint32_t)lo; prefix[i] = 0;
    }
}

template <typename S>
__global__ void __launch_bounds__(128) k_q_hist(const S* __restrict__ x, long T, long HW, int W, int y0, long npts, int nq, int shift,
                                                const typename Key<S>::K* __restrict__ prefix, uint32_t* __restrict__ hist) {
    using KT = typename Key<S>::K;
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const S* col = x + (long)y0 * W + p;
    for (int iq = 0; iq < nq; ++iq) {
        const KT pre = prefix[(long)iq * npts + p];
        const KT mask = shift + 4 >= Key<S>::BITS ? (KT)0 : (~(KT)0) << (shift + 4);      // digits already fixed
        uint32_t c[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) c[j] = 0;
        long t = 0;
        for (; t + 4 <= T; t += 4) {
            S v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(col + (t + u) * HW);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const KT key = Key<S>::of(v[u]);
                const bool ok = (v[u] == v[u]) && ((key & mask) == pre);
                const uint32_t dg = (uint32_t)(key >> shift) & 15u;
#pragma unroll
                for (int j = 0; j < 16; ++j) c[j] += (ok && dg == (uint32_t)j);
            }
        }
        for (; t < T; ++t) {
            const S v = __ldg(col + t * HW);
            const KT key = Key<S>::of(v);
            const bool ok = (v == v) && ((key & mask) == pre);
            const uint32_t dg = (uint32_t)(key >> shift) & 15u;
#pragma unroll
            for (int j = 0; j < 16; ++j) c[j] += (ok && dg == (uint32_t)j);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) hist[((long)iq * 16 + j) * npts + p] = c[j];
    }
}

template <typename KT>
__global__ void __launch_bounds__(128) k_q_select(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ cnt, long npts,
                                                  int nq, int shift, uint32_t* __restrict__ k_io, KT* __restrict__ prefix) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts || cnt[p] == 0) return;
    for (int iq = 0; iq < nq; ++iq) {
        const long i = (long)iq * npts + p;
        uint32_t k = k_io[i], dsel = 15;
        bool found = false;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t c = hist[((long)iq * 16 + j) * npts + p];
            if (!found) {
                if (k < c) { dsel = j; found = true; }
                else k -= c;
            }
        }
        k_io[i] = k;
        prefix[i] |= (KT)dsel << shift;
    }
}

template <typename S>
__global__ void __launch_bounds__(128) k_q_neigh(const S* __restrict__ x, long T, long HW, int W, int y0, long npts, int nq,
                                                 const typename Key<S>::K* __restrict__ prefix, uint32_t* __restrict__ le_out,
                                                 long long* __restrict__ best_out) {
    using KT = typename Key<S>::K;
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const S* col = x + (long)y0 * W + p;
    for (int iq = 0; iq < nq; ++iq) {
        const KT pre = prefix[(long)iq * npts + p];
        uint32_t le = 0;
        KT best = ~(KT)0;
        for (long t = 0; t < T; ++t) {
            const S v = __ldg(col + t * HW);
            if (v == v) {
                const KT key = Key<S>::of(v);
                le += key <= pre;
                if (key > pre && key < best) best = key;
            }
        }
        le_out[(long)iq * npts + p] = le;
        // keys travel as signed 64-bit integers of the same order (a MIN all-reduce over the ranks follows)
        best_out[(long)iq * npts + p] = sizeof(KT) == 4 ? (long long)best : (long long)(best ^ (KT)0x8000000000000000ull);
    }
}

template <typename S>
__global__ void __launch_bounds__(128) k_q_finish(const uint32_t* __restrict__ cnt, long npts, int nq, const uint32_t* __restrict__ lo_in,
                                                  const double* __restrict__ gamma_in, const typename Key<S>::K* __restrict__ prefix,
                                                  const uint32_t* __restrict__ le_in, const long long* __restrict__ best_in,
                                                  double* __restrict__ out) {
    using KT = typename Key<S>::K;
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    for (int iq = 0; iq < nq; ++iq) {
        const long i = (long)iq * npts + p;
        if (cnt[p] == 0) { out[i] = __longlong_as_double(0x7ff8000000000000LL); continue; }
        const double gamma = gamma_in[i];
        const S a = Key<S>::back(prefix[i]);
        S b = a;
        if (gamma != 0.0) {
            // (lo+1)-th smallest: a again if a occurs beyond position lo, else the smallest value above a
            const KT best = sizeof(KT) == 4 ? (KT)best_in[i] : (KT)((unsigned long long)best_in[i] ^ 0x8000000000000000ull);
            b = ((long)le_in[i] >= (long)lo_in[i] + 2) ? a : Key<S>::back(best);
        }
        double r;
        if (sizeof(S) == 4) {
            const float d = __fsub_rn((float)b, (float)a);
            r = __dadd_rn((double)a, __dmul_rn((double)d, gamma));
            if (gamma >= 0.5) r = __dsub_rn((double)b, __dmul_rn((double)d, __dsub_rn(1.0, gamma)));
        } else {
            const double d = __dsub_rn((double)b, (double)a);
            r = __dadd_rn((double)a, __dmul_rn(d, gamma));
            if (gamma >= 0.5) r = __dsub_rn((double)b, __dmul_rn(d, __dsub_rn(1.0, gamma)));
        }
        out[i] = r;
    }
}

// count[cell] += number of time steps in this block's time slice with flag > v; four cells per thread
template <int V>
__global__ void __launch_bounds__(256) k_flag_count(const int32_t* __restrict__ flag, long T, long HW, int tsplit, int v,
                                                    int32_t* __restrict__ count) {
    const long cell = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (cell >= HW) return;
    const long per = (T + tsplit - 1) / tsplit;
    const long t0 = (long)blockIdx.y * per, t1 = t0 + per < T ? t0 + per : T;
    int acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0;
    for (long t = t0; t < t1; ++t) {
        if (V == 4) {
            const int4 f = __ldcs(reinterpret_cast<const int4*>(flag + t * HW + cell));
            acc[0] += f.x > v; acc[1] += f.y > v; acc[2] += f.z > v; acc[3] += f.w > v;
        } else {
            acc[0] += __ldcs(flag + t * HW + cell) > v;
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) if (acc[i]) atomicAdd(count + cell + i, acc[i]);
}

__global__ void __launch_bounds__(256) k_divide_f32(const float* __restrict__ in, size_t n, float g, float* __restrict__ out) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 4 <= n) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(in + i));
        __stcs(reinterpret_cast<float4*>(out + i),
               make_float4(__fdiv_rn(v.x, g), __fdiv_rn(v.y, g), __fdiv_rn(v.z, g), __fdiv_rn(v.w, g)));
    } else {
        for (size_t j = i; j < n; ++j) out[j] = __fdiv_rn(in[j], g);
    }
}

// dst[g, y, x] = src[g, iy[y], ix[x]]
template <typename S>
__global__ void __launch_bounds__(256) k_gather_planes(const S* __restrict__ src, int Hs, int Ws,
                                                       const int32_t* __restrict__ iy, const int32_t* __restrict__ ix, int H,
                                                       int W, S* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const long g = blockIdx.z;
    if (x >= W) return;
    dst[(g * H + y) * (long)W + x] = __ldg(src + (g * Hs + iy[y]) * (long)Ws + ix[x]);
}

}  // namespace

size_t quantile_scratch_bytes(long npts, int nq, int f64) {
    const size_t n = (size_t)npts, q = (size_t)nq, key = f64 ? 8 : 4;
    return n * 4 + q * n * (4 + 8 + 4 + key + 4 + 8) + q * 16 * n * 4 + 256;
}

template <typename S>
static cudaError_t quantile_passes(const S* x, long T, int H, int W, int y0, int y1, const double* q_dev, int nq, double* out,
                                   char* scratch, const QuantileReduce* red, cudaStream_t st) {
    using KT = typename Key<S>::K;
    const long npts = (long)(y1 - y0) * W, HW = (long)H * W;
    const size_t n = (size_t)npts, q = (size_t)nq;
    // carve-up (8-byte arrays first)
    double* gamma = reinterpret_cast<double*>(scratch);
    long long* best = reinterpret_cast<long long*>(gamma + q * n);
    KT* prefix = reinterpret_cast<KT*>(best + q * n);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(prefix + q * n);
    uint32_t* lo = cnt + n;
    uint32_t* k = lo + q * n;
    uint32_t* le = k + q * n;
    uint32_t* hist = le + q * n;
    const unsigned blocks = (unsigned)((npts + 127) / 128);
    k_q_count<S><<<blocks, 128, 0, st>>>(x, T, HW, W, y0, npts, cnt);
    if (red && red->sum_u32(red->user, cnt, n, st)) return cudaErrorUnknown;
    k_q_init<KT><<<blocks, 128, 0, st>>>(cnt, npts, q_dev, nq, lo, gamma, k, prefix);
    for (int shift = Key<S>::BITS - 4; shift >= 0; shift -= 4) {
        k_q_hist<S><<<blocks, 128, 0, st>>>(x, T, HW, W, y0, npts, nq, shift, prefix, hist);
        if (red && red->sum_u32(red->user, hist, q * 16 * n, st)) return cudaErrorUnknown;
        k_q_select<KT><<<blocks, 128, 0, st>>>(hist, cnt, npts, nq, shift, k, prefix);
    }
    k_q_neigh<S><<<blocks, 128, 0, st>>>(x, T, HW, W, y0, npts, nq, prefix, le, best);
    if (red && (red->sum_u32(red->user, le, q * n, st) || red->min_i64(red->user, best, q * n, st))) return cudaErrorUnknown;
    k_q_finish<S><<<blocks, 128, 0, st>>>(cnt, npts, nq, lo, gamma, prefix, le, best, out);
    return cudaGetLastError();
}

cudaError_t quantile_time(const void* x, int f64, long T, int H, int W, int y0, int y1, const double* q_dev, int nq, double* out,
                          void* scratch, const QuantileReduce* red, cudaStream_t st) {
    const long npts = (long)(y1 - y0) * W;
    if (npts <= 0 || nq <= 0) return cudaSuccess;
    return f64 ? quantile_passes((const double*)x, T, H, W, y0, y1, q_dev, nq, out, (char*)scratch, red, st)
               : quantile_passes((const float*)x, T, H, W, y0, y1, q_dev, nq, out, (char*)scratch, red, st);
}

cudaError_t flag_count(const int32_t* flag, long T, int H, int W, int v, int32_t* count, int sm_count, cudaStream_t st) {
    const long HW = (long)H * W;
    cudaError_t e = cudaMemsetAsync(count, 0, (size_t)HW * 4, st);
    if (e != cudaSuccess || T == 0) return e;
    const bool v4 = HW % 4 == 0 && ((reinterpret_cast<uintptr_t>(flag) | reinterpret_cast<uintptr_t>(count)) & 15) == 0;
    const long groups = v4 ? HW / 4 : HW;
    const unsigned bx = (unsigned)((groups + 255) / 256);
    // enough time slices to fill the machine a few times over, never more than T
    long ts = ((long)sm_count * 16 + bx - 1) / bx;
    if (ts < 1) ts = 1;
    if (ts > T) ts = T;
    if (ts > 65535) ts = 65535;
    if (v4) k_flag_count<4><<<dim3(bx, (unsigned)ts), 256, 0, st>>>(flag, T, HW, (int)ts, v, count);
    else k_flag_count<1><<<dim3(bx, (unsigned)ts), 256, 0, st>>>(flag, T, HW, (int)ts, v, count);
    return cudaGetLastError();
}

cudaError_t divide_f32(const float* in, size_t n, float g, float* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return cudaErrorMisalignedAddress;
    k_divide_f32<<<(unsigned)((n / 4 + 1 + 255) / 256), 256, 0, st>>>(in, n, g, out);
    return cudaGetLastError();
}

template <typename S>
static cudaError_t gather_planes_t(const S* src, int G, int Hs, int Ws, const int32_t* iy_dev, const int32_t* ix_dev, int H, int W,
                                   S* dst, cudaStream_t st) {
    for (int g0 = 0; g0 < G; g0 += 65535) {
        const int ng = G - g0 < 65535 ? G - g0 : 65535;
        k_gather_planes<S><<<dim3((unsigned)((W + 255) / 256), (unsigned)H, (unsigned)ng), 256, 0, st>>>(
            src + (size_t)g0 * Hs * Ws, Hs, Ws, iy_dev, ix_dev, H, W, dst + (size_t)g0 * H * W);
    }
    return cudaGetLastError();
}

cudaError_t gather_planes(const void* src, int f64, int G, int Hs, int Ws, const int32_t* iy_dev, const int32_t* ix_dev, int H,
                          int W, void* dst, cudaStream_t st) {
    if (G <= 0 || H <= 0 || W <= 0) return cudaSuccess;
    return f64 ? gather_planes_t((const double*)src, G, Hs, Ws, iy_dev, ix_dev, H, W, (double*)dst, st)
               : gather_planes_t((const float*)src, G, Hs, Ws, iy_dev, ix_dev, H, W, (float*)dst, st);
}

}  // namespace cte
