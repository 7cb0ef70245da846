// ct_extras.cu -- the callers either side of run_contrack that README.rst shows:
//   quantile_time     README.rst:150-151  ds[var].sel(lat band).quantile(q, dim='time')  (numpy nanquantile, 'linear')
//   flag_count        README.rst:161      xr.where(flag > v, 1, 0).sum(dim='time')       (blocking frequency numerator)
//   divide_f32        contrack.py:417-419 geopotential -> geopotential height, float32 division by g
//   gather_planes     contrack.py:565     clim.reindex(lat, lon, method='nearest') with index maps computed by the caller
// All four are HBM-bound streaming kernels: a thread owns one grid point (or four) and walks the time axis; neighbouring
// threads read neighbouring cells, so every warp load is one coalesced 128-byte (or 512-byte) request.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>

namespace cte {

namespace {

// order-preserving map float32 -> uint32 (ascending; -0.0 sorts before +0.0)
__device__ __forceinline__ uint32_t f2key(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// One thread per grid point of rows [y0, y1): exact k-th order statistics of the T values along time by an 8-pass radix
// select (4 bits per pass, 16 counters in registers), NaN skipped, then numpy's linear interpolation
//   virtual = (n - 1) * q;  lo = floor(virtual);  gamma = virtual - lo
//   r = a + (b - a) * gamma            (b - a in float32, the rest in float64: numpy/lib/_function_base_impl.py:_lerp)
//   r = b - (b - a) * (1 - gamma)      where gamma >= 0.5
// T values x 9 passes per quantile are read; the reads of a warp are contiguous in x.
__global__ void __launch_bounds__(128) k_quantile_time(const float* __restrict__ x, long T, long HW, int W, int y0, long npts,
                                                       const double* __restrict__ q, int nq, double* __restrict__ out) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const float* col = x + (long)y0 * W + p;
    long n = 0;
    {
        long t = 0;
        for (; t + 4 <= T; t += 4) {
            const float a = __ldg(col + t * HW), b = __ldg(col + (t + 1) * HW), c = __ldg(col + (t + 2) * HW),
                        d = __ldg(col + (t + 3) * HW);
            n += (a == a) + (b == b) + (c == c) + (d == d);
        }
        for (; t < T; ++t) { const float a = __ldg(col + t * HW); n += (a == a); }
    }
    for (int iq = 0; iq < nq; ++iq) {
        double* dst = out + (long)iq * npts + p;
        if (n == 0) { *dst = __longlong_as_double(0x7ff8000000000000LL); continue; }
        const double qq = q[iq];
        const double virt = __dmul_rn((double)(n - 1), qq);
        long lo;
        double gamma;
        if (!(virt < (double)(n - 1))) { lo = n - 1; gamma = 0.0; }           // numpy: index -1 for both neighbours
        else if (virt < 0.0) { lo = 0; gamma = 0.0; }
        else { const double fl = floor(virt); lo = (long)fl; gamma = __dsub_rn(virt, fl); }
        // ---- radix select of the lo-th smallest key ----
        uint32_t prefix = 0, mask = 0;
        long k = lo;
        for (int shift = 28; shift >= 0; shift -= 4) {
            uint32_t c[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) c[j] = 0;
            long t = 0;
            for (; t + 4 <= T; t += 4) {
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg(col + (t + u) * HW);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t key = f2key(v[u]);
                    const bool ok = (v[u] == v[u]) && ((key & mask) == prefix);
                    const uint32_t dg = (key >> shift) & 15u;
#pragma unroll
                    for (int j = 0; j < 16; ++j) c[j] += (ok && dg == (uint32_t)j);
                }
            }
            for (; t < T; ++t) {
                const float v = __ldg(col + t * HW);
                const uint32_t key = f2key(v);
                const bool ok = (v == v) && ((key & mask) == prefix);
                const uint32_t dg = (key >> shift) & 15u;
#pragma unroll
                for (int j = 0; j < 16; ++j) c[j] += (ok && dg == (uint32_t)j);
            }
            uint32_t dsel = 15;
            bool found = false;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (!found) {
                    if (k < (long)c[j]) { dsel = j; found = true; }
                    else k -= c[j];
                }
            }
            prefix |= dsel << shift;
            mask |= 15u << shift;
        }
        const float a = key2f(prefix);
        float b = a;
        if (gamma != 0.0) {
            // (lo+1)-th smallest: a again if a occurs beyond position lo, else the smallest value above a
            long le = 0;
            uint32_t best = 0xffffffffu;
            for (long t = 0; t < T; ++t) {
                const float v = __ldg(col + t * HW);
                if (v == v) {
                    const uint32_t key = f2key(v);
                    le += key <= prefix;
                    if (key > prefix && key < best) best = key;
                }
            }
            b = (le >= lo + 2) ? a : key2f(best);
        }
        const float d = __fsub_rn(b, a);
        double r = __dadd_rn((double)a, __dmul_rn((double)d, gamma));
        if (gamma >= 0.5) r = __dsub_rn((double)b, __dmul_rn((double)d, __dsub_rn(1.0, gamma)));
        *dst = r;
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

cudaError_t quantile_time(const float* x, long T, int H, int W, int y0, int y1, const double* q_dev, int nq, double* out,
                          cudaStream_t st) {
    const long npts = (long)(y1 - y0) * W;
    if (npts <= 0 || nq <= 0) return cudaSuccess;
    k_quantile_time<<<(unsigned)((npts + 127) / 128), 128, 0, st>>>(x, T, (long)H * W, W, y0, npts, q_dev, nq, out);
    return cudaGetLastError();
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
