// ct_anom.cu -- calc_clim / calc_anom (reference contrack/contrack.py:458-491, 494-581) as float32 streaming kernels.
//
//   group_mean   contrack.py:483      mean over the time steps of each group (e.g. day of year), NaN skipped
//   clim_smooth  contrack.py:487-489  centred rolling mean over the group axis (window w, all w values required),
//                                     incomplete windows filled with the mean of the LAST w unsmoothed group means
//   anom         contrack.py:568-570  centred rolling mean over time (window `smooth`) of z[t] - clim[group[t]]
//
// Each thread owns four consecutive cells of a plane (one 16-byte load/store per plane it touches); planes are read
// once from HBM (re-reads of neighbouring planes inside a window hit L2).  Sums are accumulated in float64 and rounded to
// float32 once; the deviation z - clim itself is a float32 subtraction as in the reference.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>

namespace cta {

namespace {

// V consecutive cells of a plane per thread = one 16-byte access (4 float32 or 2 float64), or one scalar when the plane
// size / alignment does not allow it
template <typename S, int V> struct IO;
template <> struct IO<float, 4> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p)); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void load_stream(const float* p, float (&v)[4]) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(p)); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    static __device__ __forceinline__ void store_stream(float* p, const float (&v)[4]) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    }
};
template <> struct IO<double, 2> {
    static __device__ __forceinline__ void load(const double* p, double (&v)[2]) {
        const double2 q = __ldg(reinterpret_cast<const double2*>(p)); v[0] = q.x; v[1] = q.y;
    }
    static __device__ __forceinline__ void load_stream(const double* p, double (&v)[2]) {
        const double2 q = __ldcs(reinterpret_cast<const double2*>(p)); v[0] = q.x; v[1] = q.y;
    }
    static __device__ __forceinline__ void store(double* p, const double (&v)[2]) {
        *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    }
    static __device__ __forceinline__ void store_stream(double* p, const double (&v)[2]) {
        __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1]));
    }
};
template <typename S> struct IO<S, 1> {
    static __device__ __forceinline__ void load(const S* p, S (&v)[1]) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void load_stream(const S* p, S (&v)[1]) { v[0] = __ldcs(p); }
    static __device__ __forceinline__ void store(S* p, const S (&v)[1]) { *p = v[0]; }
    static __device__ __forceinline__ void store_stream(S* p, const S (&v)[1]) { __stcs(p, v[0]); }
};
template <typename S> struct Wide { static constexpr int V = 16 / sizeof(S); };

// grid (cell groups, G): mean of z[t] over the members t of group blockIdx.y
template <typename S, int V>
__global__ void __launch_bounds__(256) k_group_mean(const S* __restrict__ z, long HW, const int32_t* __restrict__ gptr,
                                                    const int32_t* __restrict__ gidx, S* __restrict__ gmean) {
    const long cell = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (cell >= HW) return;
    const int g = blockIdx.y;
    double sum[V];
    int cnt[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { sum[i] = 0.0; cnt[i] = 0; }
    const int e = gptr[g + 1];
    int k = gptr[g];
    for (; k + 4 <= e; k += 4) {                       // four planes in flight per thread (same summation order)
        S v[4][V];
#pragma unroll
        for (int u = 0; u < 4; ++u) IO<S, V>::load_stream(z + (long)gidx[k + u] * HW + cell, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < V; ++i) if (v[u][i] == v[u][i]) { sum[i] += (double)v[u][i]; ++cnt[i]; }
    }
    for (; k < e; ++k) {
        S v[V];
        IO<S, V>::load_stream(z + (long)gidx[k] * HW + cell, v);
#pragma unroll
        for (int i = 0; i < V; ++i) if (v[i] == v[i]) { sum[i] += (double)v[i]; ++cnt[i]; }
    }
    S out[V];
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = cnt[i] ? (S)(sum[i] / cnt[i]) : (S)NAN;
    IO<S, V>::store(gmean + (long)g * HW + cell, out);
}

// thread per cell group: walks the group axis with a sliding window
template <typename S, int V>
__global__ void __launch_bounds__(256) k_clim_smooth(const S* __restrict__ gmean, long HW, int G, int window,
                                                     S* __restrict__ clim) {
    const long cell = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (cell >= HW) return;
    // fill value: mean (NaN skipped) of the last `window` unsmoothed entries (contrack.py:488: clim[-window:])
    double fs[V];
    int fc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { fs[i] = 0.0; fc[i] = 0; }
    for (int g = (window < G ? G - window : 0); g < G; ++g) {
        S v[V];
        IO<S, V>::load(gmean + (long)g * HW + cell, v);
#pragma unroll
        for (int i = 0; i < V; ++i) if (v[i] == v[i]) { fs[i] += (double)v[i]; ++fc[i]; }
    }
    S fill[V];
#pragma unroll
    for (int i = 0; i < V; ++i) fill[i] = fc[i] ? (S)(fs[i] / fc[i]) : (S)NAN;
    const int left = window / 2;
    double sum[V];
    int nan[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { sum[i] = 0.0; nan[i] = 0; }
    // window of output g covers [g - left, g - left + window)
    for (int g = -(window - 1 - left); g < G; ++g) {
        const int enter = g - left + window - 1;
        if (enter >= 0 && enter < G) {
            S v[V];
            IO<S, V>::load(gmean + (long)enter * HW + cell, v);
#pragma unroll
            for (int i = 0; i < V; ++i) { if (v[i] == v[i]) sum[i] += (double)v[i]; else ++nan[i]; }
        }
        if (g >= 0) {
            const bool complete = (g - left >= 0) && (enter < G);
            S out[V];
#pragma unroll
            for (int i = 0; i < V; ++i) out[i] = (complete && nan[i] == 0) ? (S)(sum[i] / window) : fill[i];
            IO<S, V>::store(clim + (long)g * HW + cell, out);
        }
        const int leave = g - left;
        if (leave >= 0 && leave < G) {
            S v[V];
            IO<S, V>::load(gmean + (long)leave * HW + cell, v);
#pragma unroll
            for (int i = 0; i < V; ++i) { if (v[i] == v[i]) sum[i] -= (double)v[i]; else --nan[i]; }
        }
    }
}

// grid (cell groups, T): anom[t] = mean over the centred window of (z - clim[group])
template <typename S, int V>
__global__ void __launch_bounds__(256) k_anom(const S* __restrict__ z, long HW, long T, long t_off,
                                              const int32_t* __restrict__ group, const S* __restrict__ clim,
                                              int smooth, S* __restrict__ anom) {
    const long cell = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (cell >= HW) return;
    const long t = t_off + blockIdx.y;
    const long a = t - smooth / 2, b = a + smooth;
    S out[V];
    if (a < 0 || b > T) {
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = (S)NAN;
    } else if (smooth == 1) {
        S zv[V], cv[V];
        IO<S, V>::load_stream(z + t * HW + cell, zv);
        IO<S, V>::load(clim + (long)group[t] * HW + cell, cv);
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = zv[i] - cv[i];
    } else {
        double sum[V];
#pragma unroll
        for (int i = 0; i < V; ++i) sum[i] = 0.0;
        for (long k = a; k < b; ++k) {
            S zv[V], cv[V];
            IO<S, V>::load(z + k * HW + cell, zv);
            IO<S, V>::load(clim + (long)group[k] * HW + cell, cv);
#pragma unroll
            for (int i = 0; i < V; ++i) sum[i] += (double)(zv[i] - cv[i]);       // NaN propagates: all values required
        }
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = (S)(sum[i] / smooth);
    }
    IO<S, V>::store_stream(anom + t * HW + cell, out);
}

// The same arithmetic with the time axis cut into chunks that start where the group index wraps (one chunk per year for a
// day-of-year climatology): grid (chunk, cell group), chunk fastest, so the blocks that are resident together are the SAME
// cells in all years, walking the groups in step -- clim[g] is fetched from HBM once and then served from L2 to the other
// years, instead of once per year (4 B/cell saved); z[t-1] .. of the window was read by the same thread one step earlier.
template <typename S, int V>
__global__ void __launch_bounds__(256) k_anom_chunks(const S* __restrict__ z, long HW, long T,
                                                     const int32_t* __restrict__ chunk_start, const int32_t* __restrict__ group,
                                                     const S* __restrict__ clim, int smooth, S* __restrict__ anom) {
    const long cell = ((long)blockIdx.y * blockDim.x + threadIdx.x) * V;
    if (cell >= HW) return;
    const long t0 = chunk_start[blockIdx.x], t1 = chunk_start[blockIdx.x + 1];
    const int left = smooth / 2;
    if (smooth == 1) {
#pragma unroll 4
        for (long t = t0; t < t1; ++t) {
            S zv[V], cv[V], out[V];
            IO<S, V>::load_stream(z + t * HW + cell, zv);
            IO<S, V>::load(clim + (long)group[t] * HW + cell, cv);
#pragma unroll
            for (int i = 0; i < V; ++i) out[i] = zv[i] - cv[i];
            IO<S, V>::store_stream(anom + t * HW + cell, out);
        }
        return;
    }
    if (smooth == 2) {
        // window {t-1, t}: the previous deviation stays in registers, every plane is loaded once; same summation order as
        // the generic loop below (0 + dev[t-1] + dev[t] in float64, divided by 2, rounded once)
        S prev[V];
        if (t0 >= 1) {
            S zv[V], cv[V];
            IO<S, V>::load(z + (t0 - 1) * HW + cell, zv);
            IO<S, V>::load(clim + (long)group[t0 - 1] * HW + cell, cv);
#pragma unroll
            for (int i = 0; i < V; ++i) prev[i] = zv[i] - cv[i];
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) prev[i] = (S)NAN;                  // out[0] = NaN: the window is incomplete
        }
#pragma unroll 4
        for (long t = t0; t < t1; ++t) {
            S zv[V], cv[V], out[V];
            IO<S, V>::load_stream(z + t * HW + cell, zv);
            IO<S, V>::load(clim + (long)group[t] * HW + cell, cv);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const S d = zv[i] - cv[i];
                out[i] = (t == 0) ? NAN : (S)((0.0 + (double)prev[i] + (double)d) / 2);
                prev[i] = d;
            }
            IO<S, V>::store_stream(anom + t * HW + cell, out);
        }
        return;
    }
    // deviations of the window in a register ring would need a compile-time size: re-read instead (L1 / L2 hits)
    for (long t = t0; t < t1; ++t) {
        const long a = t - left, b = a + smooth;
        S out[V];
        if (a < 0 || b > T) {
#pragma unroll
            for (int i = 0; i < V; ++i) out[i] = (S)NAN;
        } else {
            double sum[V];
#pragma unroll
            for (int i = 0; i < V; ++i) sum[i] = 0.0;
            for (long k = a; k < b; ++k) {
                S zv[V], cv[V];
                IO<S, V>::load(z + k * HW + cell, zv);
                IO<S, V>::load(clim + (long)group[k] * HW + cell, cv);
#pragma unroll
                for (int i = 0; i < V; ++i) sum[i] += (double)(zv[i] - cv[i]);
            }
#pragma unroll
            for (int i = 0; i < V; ++i) out[i] = (S)(sum[i] / smooth);
        }
        IO<S, V>::store_stream(anom + t * HW + cell, out);
    }
}

inline unsigned blocks_for(long n, int per) { return (unsigned)((n + per - 1) / per); }

template <typename S>
cudaError_t group_mean_t(const S* z, long HW, int G, const int32_t* gptr_dev, const int32_t* gidx_dev, S* gmean, cudaStream_t st) {
    constexpr int V = Wide<S>::V;
    const bool wide = HW % V == 0 && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(gmean)) & 15) == 0;
    if (wide) k_group_mean<S, V><<<dim3(blocks_for(HW, 256 * V), G), 256, 0, st>>>(z, HW, gptr_dev, gidx_dev, gmean);
    else k_group_mean<S, 1><<<dim3(blocks_for(HW, 256), G), 256, 0, st>>>(z, HW, gptr_dev, gidx_dev, gmean);
    return cudaGetLastError();
}

template <typename S>
cudaError_t clim_smooth_t(const S* gmean, long HW, int G, int window, S* clim, cudaStream_t st) {
    constexpr int V = Wide<S>::V;
    const bool wide = HW % V == 0 && ((reinterpret_cast<uintptr_t>(clim) | reinterpret_cast<uintptr_t>(gmean)) & 15) == 0;
    if (wide) k_clim_smooth<S, V><<<blocks_for(HW, 256 * V), 256, 0, st>>>(gmean, HW, G, window, clim);
    else k_clim_smooth<S, 1><<<blocks_for(HW, 256), 256, 0, st>>>(gmean, HW, G, window, clim);
    return cudaGetLastError();
}

template <typename S>
cudaError_t anom_t(const S* z, long HW, long T, const int32_t* group_dev, const S* clim, int smooth, S* out, cudaStream_t st) {
    constexpr int V = Wide<S>::V;
    const bool wide = HW % V == 0 &&
                      ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(clim) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    for (long t0 = 0; t0 < T; t0 += 32768) {                        // grid.y is limited to 65535
        const unsigned nt = (unsigned)((t0 + 32768 <= T) ? 32768 : T - t0);
        if (wide) k_anom<S, V><<<dim3(blocks_for(HW, 256 * V), nt), 256, 0, st>>>(z, HW, T, t0, group_dev, clim, smooth, out);
        else k_anom<S, 1><<<dim3(blocks_for(HW, 256), nt), 256, 0, st>>>(z, HW, T, t0, group_dev, clim, smooth, out);
    }
    return cudaGetLastError();
}

template <typename S>
cudaError_t anom_chunks_t(const S* z, long HW, long T, const int32_t* chunk_start_dev, int nchunks, const int32_t* group_dev,
                          const S* clim, int smooth, S* out, cudaStream_t st) {
    constexpr int V = Wide<S>::V;
    const bool wide = HW % V == 0 &&
                      ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(clim) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (nchunks <= 0) return cudaSuccess;
    if (wide) k_anom_chunks<S, V><<<dim3((unsigned)nchunks, blocks_for(HW, 256 * V)), 256, 0, st>>>(z, HW, T, chunk_start_dev,
                                                                                               group_dev, clim, smooth, out);
    else k_anom_chunks<S, 1><<<dim3((unsigned)nchunks, blocks_for(HW, 256)), 256, 0, st>>>(z, HW, T, chunk_start_dev, group_dev,
                                                                                        clim, smooth, out);
    return cudaGetLastError();
}

}  // namespace

// `f64` selects the element type of z / clim / anomaly (float32 or float64: xarray keeps the input's precision)
cudaError_t group_mean(const void* z, int f64, long HW, int G, const int32_t* gptr_dev, const int32_t* gidx_dev, void* gmean,
                       cudaStream_t st) {
    return f64 ? group_mean_t((const double*)z, HW, G, gptr_dev, gidx_dev, (double*)gmean, st)
               : group_mean_t((const float*)z, HW, G, gptr_dev, gidx_dev, (float*)gmean, st);
}
cudaError_t clim_smooth(const void* gmean, int f64, long HW, int G, int window, void* clim, cudaStream_t st) {
    return f64 ? clim_smooth_t((const double*)gmean, HW, G, window, (double*)clim, st)
               : clim_smooth_t((const float*)gmean, HW, G, window, (float*)clim, st);
}
cudaError_t anom(const void* z, int f64, long HW, long T, const int32_t* group_dev, const void* clim, int smooth, void* out,
                 cudaStream_t st) {
    return f64 ? anom_t((const double*)z, HW, T, group_dev, (const double*)clim, smooth, (double*)out, st)
               : anom_t((const float*)z, HW, T, group_dev, (const float*)clim, smooth, (float*)out, st);
}
cudaError_t anom_chunks(const void* z, int f64, long HW, long T, const int32_t* chunk_start_dev, int nchunks,
                        const int32_t* group_dev, const void* clim, int smooth, void* out, cudaStream_t st) {
    return f64 ? anom_chunks_t((const double*)z, HW, T, chunk_start_dev, nchunks, group_dev, (const double*)clim, smooth,
                               (double*)out, st)
               : anom_chunks_t((const float*)z, HW, T, chunk_start_dev, nchunks, group_dev, (const float*)clim, smooth,
                               (float*)out, st);
}

}  // namespace cta
