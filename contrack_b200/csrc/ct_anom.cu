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

template <int V> struct Vec;
template <> struct Vec<4> { using type = float4; };
template <> struct Vec<1> { using type = float; };

template <int V> __device__ __forceinline__ void load(const float* p, float (&v)[V]);
template <> __device__ __forceinline__ void load<4>(const float* p, float (&v)[4]) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <> __device__ __forceinline__ void load<1>(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
template <int V> __device__ __forceinline__ void load_stream(const float* p, float (&v)[V]);
template <> __device__ __forceinline__ void load_stream<4>(const float* p, float (&v)[4]) {
    const float4 q = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <> __device__ __forceinline__ void load_stream<1>(const float* p, float (&v)[1]) { v[0] = __ldcs(p); }
template <int V> __device__ __forceinline__ void store(float* p, const float (&v)[V]);
template <> __device__ __forceinline__ void store<4>(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store<1>(float* p, const float (&v)[1]) { *p = v[0]; }
template <int V> __device__ __forceinline__ void store_stream(float* p, const float (&v)[V]);
template <> __device__ __forceinline__ void store_stream<4>(float* p, const float (&v)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
}
template <> __device__ __forceinline__ void store_stream<1>(float* p, const float (&v)[1]) { __stcs(p, v[0]); }

// grid (cell groups, G): mean of z[t] over the members t of group blockIdx.y
template <int V>
__global__ void __launch_bounds__(256) k_group_mean(const float* __restrict__ z, long HW, const int32_t* __restrict__ gptr,
                                                    const int32_t* __restrict__ gidx, float* __restrict__ gmean) {
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
        float v[4][V];
#pragma unroll
        for (int u = 0; u < 4; ++u) load_stream<V>(z + (long)gidx[k + u] * HW + cell, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < V; ++i) if (v[u][i] == v[u][i]) { sum[i] += (double)v[u][i]; ++cnt[i]; }
    }
    for (; k < e; ++k) {
        float v[V];
        load_stream<V>(z + (long)gidx[k] * HW + cell, v);
#pragma unroll
        for (int i = 0; i < V; ++i) if (v[i] == v[i]) { sum[i] += (double)v[i]; ++cnt[i]; }
    }
    float out[V];
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = cnt[i] ? (float)(sum[i] / cnt[i]) : NAN;
    store<V>(gmean + (long)g * HW + cell, out);
}

// thread per cell group: walks the group axis with a sliding window
template <int V>
__global__ void __launch_bounds__(256) k_clim_smooth(const float* __restrict__ gmean, long HW, int G, int window,
                                                     float* __restrict__ clim) {
    const long cell = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (cell >= HW) return;
    // fill value: mean (NaN skipped) of the last `window` unsmoothed entries (contrack.py:488: clim[-window:])
    double fs[V];
    int fc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { fs[i] = 0.0; fc[i] = 0; }
    for (int g = (window < G ? G - window : 0); g < G; ++g) {
        float v[V];
        load<V>(gmean + (long)g * HW + cell, v);
#pragma unroll
        for (int i = 0; i < V; ++i) if (v[i] == v[i]) { fs[i] += (double)v[i]; ++fc[i]; }
    }
    float fill[V];
#pragma unroll
    for (int i = 0; i < V; ++i) fill[i] = fc[i] ? (float)(fs[i] / fc[i]) : NAN;
    const int left = window / 2;
    double sum[V];
    int nan[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { sum[i] = 0.0; nan[i] = 0; }
    // window of output g covers [g - left, g - left + window)
    for (int g = -(window - 1 - left); g < G; ++g) {
        const int enter = g - left + window - 1;
        if (enter >= 0 && enter < G) {
            float v[V];
            load<V>(gmean + (long)enter * HW + cell, v);
#pragma unroll
            for (int i = 0; i < V; ++i) { if (v[i] == v[i]) sum[i] += (double)v[i]; else ++nan[i]; }
        }
        if (g >= 0) {
            const bool complete = (g - left >= 0) && (enter < G);
            float out[V];
#pragma unroll
            for (int i = 0; i < V; ++i) out[i] = (complete && nan[i] == 0) ? (float)(sum[i] / window) : fill[i];
            store<V>(clim + (long)g * HW + cell, out);
        }
        const int leave = g - left;
        if (leave >= 0 && leave < G) {
            float v[V];
            load<V>(gmean + (long)leave * HW + cell, v);
#pragma unroll
            for (int i = 0; i < V; ++i) { if (v[i] == v[i]) sum[i] -= (double)v[i]; else --nan[i]; }
        }
    }
}

// grid (cell groups, T): anom[t] = mean over the centred window of (z - clim[group])
template <int V>
__global__ void __launch_bounds__(256) k_anom(const float* __restrict__ z, long HW, long T, long t_off,
                                              const int32_t* __restrict__ group, const float* __restrict__ clim,
                                              int smooth, float* __restrict__ anom) {
    const long cell = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (cell >= HW) return;
    const long t = t_off + blockIdx.y;
    const long a = t - smooth / 2, b = a + smooth;
    float out[V];
    if (a < 0 || b > T) {
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = NAN;
    } else if (smooth == 1) {
        float zv[V], cv[V];
        load_stream<V>(z + t * HW + cell, zv);
        load<V>(clim + (long)group[t] * HW + cell, cv);
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = zv[i] - cv[i];
    } else {
        double sum[V];
#pragma unroll
        for (int i = 0; i < V; ++i) sum[i] = 0.0;
        for (long k = a; k < b; ++k) {
            float zv[V], cv[V];
            load<V>(z + k * HW + cell, zv);
            load<V>(clim + (long)group[k] * HW + cell, cv);
#pragma unroll
            for (int i = 0; i < V; ++i) sum[i] += (double)(zv[i] - cv[i]);       // NaN propagates: all values required
        }
#pragma unroll
        for (int i = 0; i < V; ++i) out[i] = (float)(sum[i] / smooth);
    }
    store_stream<V>(anom + t * HW + cell, out);
}

// The same arithmetic with the time axis cut into chunks that start where the group index wraps (one chunk per year for a
// day-of-year climatology): grid (chunk, cell group), chunk fastest, so the blocks that are resident together are the SAME
// cells in all years, walking the groups in step -- clim[g] is fetched from HBM once and then served from L2 to the other
// years, instead of once per year (4 B/cell saved); z[t-1] .. of the window was read by the same thread one step earlier.
template <int V>
__global__ void __launch_bounds__(256) k_anom_chunks(const float* __restrict__ z, long HW, long T,
                                                     const int32_t* __restrict__ chunk_start, const int32_t* __restrict__ group,
                                                     const float* __restrict__ clim, int smooth, float* __restrict__ anom) {
    const long cell = ((long)blockIdx.y * blockDim.x + threadIdx.x) * V;
    if (cell >= HW) return;
    const long t0 = chunk_start[blockIdx.x], t1 = chunk_start[blockIdx.x + 1];
    const int left = smooth / 2;
    if (smooth == 1) {
#pragma unroll 4
        for (long t = t0; t < t1; ++t) {
            float zv[V], cv[V], out[V];
            load_stream<V>(z + t * HW + cell, zv);
            load<V>(clim + (long)group[t] * HW + cell, cv);
#pragma unroll
            for (int i = 0; i < V; ++i) out[i] = zv[i] - cv[i];
            store_stream<V>(anom + t * HW + cell, out);
        }
        return;
    }
    if (smooth == 2) {
        // window {t-1, t}: the previous deviation stays in registers, every plane is loaded once; same summation order as
        // the generic loop below (0 + dev[t-1] + dev[t] in float64, divided by 2, rounded once)
        float prev[V];
        if (t0 >= 1) {
            float zv[V], cv[V];
            load<V>(z + (t0 - 1) * HW + cell, zv);
            load<V>(clim + (long)group[t0 - 1] * HW + cell, cv);
#pragma unroll
            for (int i = 0; i < V; ++i) prev[i] = zv[i] - cv[i];
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) prev[i] = NAN;                  // out[0] = NaN: the window is incomplete
        }
#pragma unroll 4
        for (long t = t0; t < t1; ++t) {
            float zv[V], cv[V], out[V];
            load_stream<V>(z + t * HW + cell, zv);
            load<V>(clim + (long)group[t] * HW + cell, cv);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float d = zv[i] - cv[i];
                out[i] = (t == 0) ? NAN : (float)((0.0 + (double)prev[i] + (double)d) / 2);
                prev[i] = d;
            }
            store_stream<V>(anom + t * HW + cell, out);
        }
        return;
    }
    // deviations of the window in a register ring would need a compile-time size: re-read instead (L1 / L2 hits)
    for (long t = t0; t < t1; ++t) {
        const long a = t - left, b = a + smooth;
        float out[V];
        if (a < 0 || b > T) {
#pragma unroll
            for (int i = 0; i < V; ++i) out[i] = NAN;
        } else {
            double sum[V];
#pragma unroll
            for (int i = 0; i < V; ++i) sum[i] = 0.0;
            for (long k = a; k < b; ++k) {
                float zv[V], cv[V];
                load<V>(z + k * HW + cell, zv);
                load<V>(clim + (long)group[k] * HW + cell, cv);
#pragma unroll
                for (int i = 0; i < V; ++i) sum[i] += (double)(zv[i] - cv[i]);
            }
#pragma unroll
            for (int i = 0; i < V; ++i) out[i] = (float)(sum[i] / smooth);
        }
        store_stream<V>(anom + t * HW + cell, out);
    }
}

inline unsigned blocks_for(long n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

cudaError_t group_mean(const float* z, long HW, int G, const int32_t* gptr_dev, const int32_t* gidx_dev, float* gmean,
                       cudaStream_t st) {
    const bool v4 = HW % 4 == 0 && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(gmean)) & 15) == 0;
    if (v4) k_group_mean<4><<<dim3(blocks_for(HW, 1024), G), 256, 0, st>>>(z, HW, gptr_dev, gidx_dev, gmean);
    else k_group_mean<1><<<dim3(blocks_for(HW, 256), G), 256, 0, st>>>(z, HW, gptr_dev, gidx_dev, gmean);
    return cudaGetLastError();
}

cudaError_t clim_smooth(const float* gmean, long HW, int G, int window, float* clim, cudaStream_t st) {
    const bool v4 = HW % 4 == 0 && ((reinterpret_cast<uintptr_t>(clim) | reinterpret_cast<uintptr_t>(gmean)) & 15) == 0;
    if (v4) k_clim_smooth<4><<<blocks_for(HW, 1024), 256, 0, st>>>(gmean, HW, G, window, clim);
    else k_clim_smooth<1><<<blocks_for(HW, 256), 256, 0, st>>>(gmean, HW, G, window, clim);
    return cudaGetLastError();
}

cudaError_t anom(const float* z, long HW, long T, const int32_t* group_dev, const float* clim, int smooth, float* out,
                 cudaStream_t st) {
    const bool v4 = HW % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(clim) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    for (long t0 = 0; t0 < T; t0 += 32768) {                        // grid.y is limited to 65535
        const unsigned nt = (unsigned)((t0 + 32768 <= T) ? 32768 : T - t0);
        if (v4) k_anom<4><<<dim3(blocks_for(HW, 1024), nt), 256, 0, st>>>(z, HW, T, t0, group_dev, clim, smooth, out);
        else k_anom<1><<<dim3(blocks_for(HW, 256), nt), 256, 0, st>>>(z, HW, T, t0, group_dev, clim, smooth, out);
    }
    return cudaGetLastError();
}

cudaError_t anom_chunks(const float* z, long HW, long T, const int32_t* chunk_start_dev, int nchunks, const int32_t* group_dev,
                        const float* clim, int smooth, float* out, cudaStream_t st) {
    const bool v4 = HW % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(clim) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (nchunks <= 0) return cudaSuccess;
    if (v4) k_anom_chunks<4><<<dim3((unsigned)nchunks, blocks_for(HW, 1024)), 256, 0, st>>>(z, HW, T, chunk_start_dev, group_dev,
                                                                                         clim, smooth, out);
    else k_anom_chunks<1><<<dim3((unsigned)nchunks, blocks_for(HW, 256)), 256, 0, st>>>(z, HW, T, chunk_start_dev, group_dev,
                                                                                      clim, smooth, out);
    return cudaGetLastError();
}

}  // namespace cta
