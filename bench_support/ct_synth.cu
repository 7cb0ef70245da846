// ct_synth.cu -- benchmark tooling (NOT part of the product library): synthetic Z500-anomaly-like cubes generated on the
// device, SURVEY.md section 8(d): white N(0,1) noise keyed on (seed, t, y, x) so that any time chunk can be regenerated
// on any GPU, separable Gaussian smoothing (nearest, nearest, wrap-in-longitude), scaled to a global std of `target_std`
// (analytic, from the filter weights), float32.  An optional seasonal cycle turns the anomaly into a "z" field for the
// calc_anom configuration.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <vector>

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

// noise for planes tb..tb+np-1 where plane index is clamped to [0, T-1] ("nearest" boundary in time)
__global__ void k_noise(float* out, uint64_t seed, long tb, long np, long T, int H, int W) {
    const long n = np * (long)H * W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        long p = i / ((long)H * W), r = i % ((long)H * W);
        long t = tb + p;
        t = t < 0 ? 0 : (t >= T ? T - 1 : t);
        const uint64_t h = mix64(mix64(seed ^ 0x5851f42d4c957f2dULL) + (uint64_t)t * (uint64_t)H * W + (uint64_t)r);
        const float u1 = ((uint32_t)(h >> 32) + 1.0f) * 2.3283064365386963e-10f;      // (0, 1]
        const float u2 = (uint32_t)h * 2.3283064365386963e-10f;
        out[i] = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    }
}

// along x, periodic: one block per row
__global__ void k_smooth_x(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ g, int r,
                           int W) {
    extern __shared__ float s[];
    float* row = s;
    float* gw = s + W;
    const long base = (long)blockIdx.x * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) row[x] = in[base + x];
    for (int k = threadIdx.x; k <= 2 * r; k += blockDim.x) gw[k] = g[k];
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        float acc = 0.f;
        int xx = x - r;
        xx %= W; if (xx < 0) xx += W;
        for (int k = 0; k <= 2 * r; ++k) {
            acc = fmaf(gw[k], row[xx], acc);
            if (++xx == W) xx = 0;
        }
        out[base + x] = acc;
    }
}

// along y, clamped: block = 32 columns of one plane
__global__ void k_smooth_y(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ g, int r,
                           int H, int W) {
    extern __shared__ float s[];
    float* tile = s;                  // [H][32]
    float* gw = s + (size_t)H * 32;
    const long plane = (long)blockIdx.y * H * W;
    const int x0 = blockIdx.x * 32, lane = threadIdx.x & 31, wy = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int y = wy; y < H; y += nw) tile[y * 32 + lane] = (x0 + lane < W) ? in[plane + (long)y * W + x0 + lane] : 0.f;
    for (int k = threadIdx.x; k <= 2 * r; k += blockDim.x) gw[k] = g[k];
    __syncthreads();
    for (int y = wy; y < H; y += nw) {
        float acc = 0.f;
        for (int k = 0; k <= 2 * r; ++k) {
            int yy = y + k - r;
            yy = yy < 0 ? 0 : (yy >= H ? H - 1 : yy);
            acc = fmaf(gw[k], tile[yy * 32 + lane], acc);
        }
        if (x0 + lane < W) out[plane + (long)y * W + x0 + lane] = acc;
    }
}

// along t over the halo'd chunk; adds the optional seasonal cycle and scales
__global__ void k_smooth_t(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ g, int r,
                           long nt, long HW, float scale, float season_amp, float season_period, long t0, int H, int W) {
    const long n = nt * HW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long p = i / HW, c = i % HW;
        float acc = 0.f;
        for (int k = 0; k <= 2 * r; ++k) acc = fmaf(g[k], in[(p + k) * HW + c], acc);
        float v = acc * scale;
        if (season_amp != 0.f) {
            const int y = (int)(c / W);
            const float lat = 1.5707963f * (1.f - 2.f * y / (float)(H - 1));
            v += 5500.f + season_amp * sinf(lat) * cospif(2.f * (float)(t0 + p) / season_period);
        }
        out[i] = v;
    }
}

// position-weighted 64-bit checksum of an int32 cube: sum over cells of value * mix64(global cell index), modulo 2^64.
// Additive over disjoint parts of a cube, so time shards can be summed (all-reduce) to the checksum of the whole cube.
__global__ void k_checksum(const int32_t* __restrict__ p, size_t n, uint64_t index0, unsigned long long* out) {
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int32_t v = __ldcs(p + i);
        if (v) acc += (unsigned long long)(uint32_t)v * (mix64(index0 + i) | 1ULL);
    }
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

std::vector<float> gauss(double sigma, int* r_out) {
    int r = (int)(4.0 * sigma + 0.5);
    std::vector<double> w(2 * r + 1);
    double s = 0;
    for (int k = -r; k <= r; ++k) { w[k + r] = std::exp(-0.5 * k * k / (sigma * sigma)); s += w[k + r]; }
    std::vector<float> f(2 * r + 1);
    for (int k = 0; k <= 2 * r; ++k) f[k] = (float)(w[k] / s);
    *r_out = r;
    return f;
}

}  // namespace

extern "C" {

// Fill out_dev[nt, H, W] with planes [t0, t0+nt) of the cube (seed, T, H, W, sigma_t, sigma_y, sigma_x).
// season_amp != 0 adds 5500 + amp*sin(lat)*cos(2 pi t / period) (a smooth "geopotential height" with a seasonal cycle).
// Returns 0 or a cudaError_t value.
int ct_synth_fill(float* out_dev, unsigned long long seed, long t0, long nt, long T, int H, int W, double sigma_t,
                  double sigma_y, double sigma_x, double target_std, double season_amp, double season_period,
                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rt, ry, rx;
    std::vector<float> gt = gauss(sigma_t, &rt), gy = gauss(sigma_y, &ry), gx = gauss(sigma_x, &rx);
    double st2 = 0, sy2 = 0, sx2 = 0;
    for (float v : gt) st2 += (double)v * v;
    for (float v : gy) sy2 += (double)v * v;
    for (float v : gx) sx2 += (double)v * v;
    const float scale = (float)(target_std / std::sqrt(st2 * sy2 * sx2));
    const long HW = (long)H * W;
    float *g_dev = nullptr, *a = nullptr, *b = nullptr;
    cudaError_t e;
    const size_t gn = gt.size() + gy.size() + gx.size();
    if ((e = cudaMalloc(&g_dev, gn * 4)) != cudaSuccess) return (int)e;
    cudaMemcpyAsync(g_dev, gt.data(), gt.size() * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(g_dev + gt.size(), gy.data(), gy.size() * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(g_dev + gt.size() + gy.size(), gx.data(), gx.size() * 4, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    const long chunk = 48;
    const size_t tmp = (size_t)(chunk + 2 * rt) * HW * 4;
    if ((e = cudaMalloc(&a, tmp)) != cudaSuccess) { cudaFree(g_dev); return (int)e; }
    if ((e = cudaMalloc(&b, tmp)) != cudaSuccess) { cudaFree(g_dev); cudaFree(a); return (int)e; }
    cudaFuncSetAttribute(k_smooth_y, cudaFuncAttributeMaxDynamicSharedMemorySize, (H * 32 + 2 * ry + 1) * 4);
    for (long c0 = 0; c0 < nt; c0 += chunk) {
        const long n = (c0 + chunk <= nt) ? chunk : nt - c0;
        const long np = n + 2 * rt;
        k_noise<<<148 * 8, 256, 0, st>>>(a, seed, t0 + c0 - rt, np, T, H, W);
        k_smooth_x<<<(unsigned)(np * H), 256, (W + 2 * rx + 1) * 4, st>>>(a, b, g_dev + gt.size() + gy.size(), rx, W);
        k_smooth_y<<<dim3((W + 31) / 32, (unsigned)np), 512, (H * 32 + 2 * ry + 1) * 4, st>>>(b, a, g_dev + gt.size(), ry,
                                                                                            H, W);
        k_smooth_t<<<148 * 8, 256, 0, st>>>(a, out_dev + c0 * HW, g_dev, rt, n, HW, scale, (float)season_amp,
                                             (float)season_period, t0 + c0, H, W);
    }
    e = cudaStreamSynchronize(st);
    cudaFree(g_dev); cudaFree(a); cudaFree(b);
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}

// *out_dev (zeroed here) = checksum of p[0..n) as cells index0 .. index0+n-1 of a larger cube (see k_checksum)
int ct_checksum_i32(const int32_t* p, size_t n, unsigned long long index0, unsigned long long* out_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out_dev, 0, 8, st);
    if (e != cudaSuccess) return (int)e;
    if (n) k_checksum<<<148 * 8, 256, 0, st>>>(p, n, index0, out_dev);
    return (int)cudaGetLastError();
}

}  // extern "C"
