// ct_shard.cu -- device side of the time-sharded run: the rank-local tables of every rank, all-gathered
// into one device buffer, are renumbered and concatenated into GLOBAL tables in the layout the single-GPU table phase
// works on (component boxes / classes, class sums, pair CSR, date-line segments).  Every rank then runs the same global
// phase (contrack.py:706-772 on tables) on its own copy and obtains the same global ids.
//
// Local component i of rank r has the global id i + off_r, off_r = comp_base_r - halo_r: own components are numbered
// consecutively in rank order (ranks are ordered in time, so this is the global first-pixel order) and the halo components
// of rank r -- the components of rank r-1's last plane, built from the same bit rows in the same raster order -- fall onto
// the ids rank r-1 gave them.
#include "ct_shard.h"

namespace cts {

namespace {

__device__ __forceinline__ int find_rank(const long* base, int nranks, long i) {
    int lo = 0, hi = nranks - 1;                       // last r with base[r] <= i (base is non-decreasing, base[0] = 0)
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (base[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <typename T> __device__ __forceinline__ const T* arr(const char* base, const RankDesc& d, int k) {
    return reinterpret_cast<const T*>(base + d.src + d.off[k]);
}

__global__ void __launch_bounds__(256) k_merge_comps(const char* __restrict__ gathered, const RankDesc* __restrict__ desc,
                                                     const long* __restrict__ comp_base, int nranks, long NC, long NP,
                                                     GlobalTables g) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > NC) return;
    if (i == NC) { g.pptr[NC] = (uint32_t)NP; return; }
    const int r = find_rank(comp_base, nranks, i);
    const RankDesc d = desc[r];
    const long j = i - d.comp_base + d.nh;             // local index
    g.t[i] = arr<int32_t>(gathered, d, A_T)[j] + (int32_t)d.t_shift;
    g.y0[i] = arr<int32_t>(gathered, d, A_Y0)[j];
    g.y1[i] = arr<int32_t>(gathered, d, A_Y1)[j];
    g.x0[i] = arr<int32_t>(gathered, d, A_X0)[j];
    g.x1[i] = arr<int32_t>(gathered, d, A_X1)[j];
    g.cls[i] = (uint32_t)((long)arr<uint32_t>(gathered, d, A_CLS)[j] + d.comp_base - d.nh);
    g.conE[i] = arr<double>(gathered, d, A_CONE)[j];
    g.conS[i] = arr<double>(gathered, d, A_CONS)[j];
    g.nsp[i] = arr<uint32_t>(gathered, d, A_NSP)[j];
    double fE = arr<double>(gathered, d, A_FE)[j], fS = arr<double>(gathered, d, A_FS)[j];
    uint32_t fn = arr<uint32_t>(gathered, d, A_FNSP)[j];
    if (r + 1 < nranks) {
        // forward overlap of this rank's last-plane classes with the next rank's first plane was accumulated over there,
        // on the halo copies of these components
        const RankDesc e = desc[r + 1];
        const long h = i - (e.comp_base - e.nh);
        if (h >= 0 && h < e.nh) {
            fE += arr<double>(gathered, e, A_FE)[h];
            fS += arr<double>(gathered, e, A_FS)[h];
            fn += arr<uint32_t>(gathered, e, A_FNSP)[h];
        }
    }
    g.fE[i] = fE; g.fS[i] = fS; g.fnsp[i] = fn;
    g.pptr[i] = (uint32_t)(d.pair_base + (long)arr<uint32_t>(gathered, d, A_PPTR)[j] - d.e0);
}

__global__ void __launch_bounds__(256) k_merge_pairs(const char* __restrict__ gathered, const RankDesc* __restrict__ desc,
                                                     const long* __restrict__ pair_base, int nranks, long NP, GlobalTables g) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    const int r = find_rank(pair_base, nranks, p);
    const RankDesc d = desc[r];
    const long q = p - d.pair_base + d.e0;
    g.p_b[p] = (uint32_t)((long)arr<uint32_t>(gathered, d, A_PB)[q] + d.comp_base - d.nh);
    g.p_npix[p] = arr<uint32_t>(gathered, d, A_PNPIX)[q];
    g.p_nsp[p] = arr<uint32_t>(gathered, d, A_PNSP)[q];
    g.p_E[p] = arr<double>(gathered, d, A_PE)[q];
    g.p_S[p] = arr<double>(gathered, d, A_PS)[q];
}

__global__ void __launch_bounds__(256) k_merge_segs(const char* __restrict__ gathered, const RankDesc* __restrict__ desc,
                                                    const long* __restrict__ seg_base, int nranks, long NS, GlobalTables g) {
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= NS) return;
    const int r = find_rank(seg_base, nranks, s);
    const RankDesc d = desc[r];
    const long q = s - d.seg_base + d.ns_h;
    const long off = d.comp_base - d.nh;
    g.g_t[s] = arr<int32_t>(gathered, d, A_GT)[q] + (int32_t)d.t_shift;
    g.g_y0[s] = arr<int32_t>(gathered, d, A_GY0)[q];
    g.g_y1[s] = arr<int32_t>(gathered, d, A_GY1)[q];
    g.g_a[s] = (uint32_t)((long)arr<uint32_t>(gathered, d, A_GA)[q] + off);
    g.g_b[s] = (uint32_t)((long)arr<uint32_t>(gathered, d, A_GB)[q] + off);
}

// out[0] = components of plane 0 (nh), out[1] = pptr[nh], out[2] = segments of plane 0, out[3] = components of the last plane
__global__ void k_shard_counts(const int32_t* __restrict__ comp_t, long nc, const uint32_t* __restrict__ pptr,
                               const int32_t* __restrict__ seg_t, long ns, int has_prev, int last_plane,
                               uint32_t* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    auto lower = [](const int32_t* a, long n, int v) {  // first index with a[i] >= v
        long lo = 0, hi = n;
        while (lo < hi) { const long m = (lo + hi) >> 1; if (a[m] < v) lo = m + 1; else hi = m; }
        return lo;
    };
    const long nh = has_prev ? lower(comp_t, nc, 1) : 0;
    out[0] = (uint32_t)nh;
    out[1] = nc ? pptr[nh] : 0u;
    out[2] = (uint32_t)(has_prev ? lower(seg_t, ns, 1) : 0);
    out[3] = (uint32_t)(nc - lower(comp_t, nc, last_plane));
}

inline unsigned blocks_for(long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

size_t layout(long nc, long np, long ns, size_t off[A_COUNT]) {
    static const int elt[A_COUNT] = {4, 4, 4, 4, 4, 4, 8, 8, 8, 8, 4, 4, 4, 4, 4, 4, 8, 8, 4, 4, 4, 4, 4};
    size_t o = 0;
    for (int k = 0; k < A_COUNT; ++k) {
        const long n = k <= A_FNSP ? nc : k == A_PPTR ? nc + 1 : k <= A_PS ? np : ns;
        off[k] = o;
        o += (((size_t)n * elt[k] + 15) / 16) * 16;
    }
    return o;
}

cudaError_t shard_counts(const int32_t* comp_t, long nc, const uint32_t* pptr, const int32_t* seg_t, long ns, int has_prev,
                         long last_plane, uint32_t* out4_dev, cudaStream_t st) {
    k_shard_counts<<<1, 32, 0, st>>>(comp_t, nc, pptr, seg_t, ns, has_prev, (int)last_plane, out4_dev);
    return cudaGetLastError();
}

cudaError_t merge(const char* gathered, const RankDesc* desc_dev, const long* bases_dev, int nranks, long NC, long NP,
                  long NS, const GlobalTables& g, cudaStream_t st) {
    k_merge_comps<<<blocks_for(NC + 1), 256, 0, st>>>(gathered, desc_dev, bases_dev, nranks, NC, NP, g);
    if (NP) k_merge_pairs<<<blocks_for(NP), 256, 0, st>>>(gathered, desc_dev, bases_dev + nranks, nranks, NP, g);
    if (NS) k_merge_segs<<<blocks_for(NS), 256, 0, st>>>(gathered, desc_dev, bases_dev + 2 * nranks, nranks, NS, g);
    return cudaGetLastError();
}

}  // namespace cts
