// ct_dist.cu -- the time-sharded run as ONE C-ABI call per rank (contrack.py:646-772 across time shards).
//
// Rank r owns planes [t_begin, t_begin + T_local) of the cube; ranks are ordered in time.  Per rank, on the device:
//   1. the LAST own plane is thresholded first and its bit rows go to rank r+1 (the one halo exchange: send/recv of
//      H * ceil(W/32) words), while the other planes are being thresholded;
//   2. the plane kernel (ct_plane.cu) builds the tables of halo + own planes beside the zero fill of the flag planes;
//   3. the tables are packed behind a 128-byte header into a slot of fixed stride and reach every rank in one step: the pack
//      kernel STORES each element straight into the gathered buffers of all ranks (peer memory over NVLink / NVSwitch, mapped
//      once per communicator) and raises a flag word at every peer when its last block is through -- pack and all-gather are
//      one kernel, there is no collective launch and no staging copy; without peer memory (or with option p2p=0) the slot is
//      packed locally and all-gathered with one ncclAllGather.  The stride is negotiated once per problem size and kept; a
//      rank whose tables outgrow it says so in its header and every rank repeats the exchange with a larger one;
//   4. a merge kernel renumbers the gathered tables into global tables -- local component i of rank r becomes
//      i + comp_base_r - halo_r; halo components fall onto the ids the previous rank gave its last-plane components; forward
//      sums accumulated on halo copies are added to their owners -- with the per-rank bases computed on the device;
//   5. every rank runs the cooperative global kernel (ct_global.cu) and the O(events) host replay on its copy and gets the
//      same global ids (the "global relabel"); 6. it paints its own planes.
// The host synchronises once (control block + date-line events), exactly like the single-GPU call.  The collectives go
// through ctc::Comm (ct_comm.h): NCCL over NVLink in production, an in-process group on single-GPU test boxes.
#include <cstring>
#include <thread>

#include "ct_comm.h"
#include "ct_extras.h"
#include "ct_fast.h"
#include "ct_internal.h"
#include "ct_shard.h"

namespace {

using cti::fail;

constexpr int HDR_WORDS = 16;                 // 128-byte header in front of a rank's packed tables
constexpr size_t HDR_BYTES = HDR_WORDS * 8;
enum Hdr { H_NC = 0, H_NH, H_NP, H_NS, H_E0, H_NSH, H_NLAST, H_TSHIFT, H_STATUS, H_NRUNS, H_CBASE, H_PBASE, H_SBASE, H_COFF };
constexpr uint32_t ST_EXCHANGE = 8u;          // a rank's tables do not fit the negotiated stride
constexpr uint32_t ST_MISMATCH = 16u;         // halo components of rank r != last-plane components of rank r-1 (internal)
constexpr int MAX_PEERS = 16;                 // ranks that can exchange through peer windows (more: the collective)
constexpr size_t WIN_DATA = 4096;             // window: [0, 256) flag words [parity][rank], [1024] pack counter, data from 4096
// A rank may legitimately enter the call long after the others (host work, I/O): the wait for its flag is as patient as a
// collective would be; only a rank that never arrives (it failed) ends the wait, with an error instead of a hang.
constexpr unsigned long long FLAG_WAIT_NS = 600ull * 1000000000ull;

struct Offsets { size_t off[cts::A_COUNT]; };

struct PackArgs {
    const int32_t *t, *y0, *y1, *x0, *x1; const uint32_t* cls;
    const double *conE, *conS, *fE, *fS; const uint32_t *nsp, *fnsp, *pptr;
    const uint32_t *pb, *pnpix, *pnsp; const double *pE, *pS;
    const int32_t *gt, *gy0, *gy1; const uint32_t *ga, *gb;
    const unsigned long long* totals;          // {components, segments, runs, pairs}
    const uint32_t* status;
    unsigned long long capC, capP, capS;
    int has_prev; int last_plane; long t_shift;
    // destinations: this rank's slot in the gathered buffer of every rank (peer windows), or one local export buffer
    char* dst[MAX_PEERS]; int ndst;
    unsigned long long* flag[MAX_PEERS];       // peer windows: this rank's flag word at every rank (null = no signalling)
    unsigned long long epoch; uint32_t* done_ctr;
    Offsets o;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ long lower_i32(const int32_t* a, long n, int v) {       // first index with a[i] >= v
    long lo = 0, hi = n;
    while (lo < hi) { const long m = (lo + hi) >> 1; if (a[m] < v) lo = m + 1; else hi = m; }
    return lo;
}

__global__ void __launch_bounds__(256) k_pack_tables(PackArgs a) {
    const long nc = (long)a.totals[0], ns = (long)a.totals[1], np = (long)a.totals[3];
    const uint32_t st = *a.status;
    const bool fits = st == 0u && (unsigned long long)nc <= a.capC && (unsigned long long)np <= a.capP &&
                      (unsigned long long)ns <= a.capS;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long h[HDR_WORDS];
        const long nh = (a.has_prev && st == 0u) ? lower_i32(a.t, nc, 1) : 0;
        h[H_NC] = nc; h[H_NH] = nh; h[H_NP] = np; h[H_NS] = ns;
        h[H_E0] = (st == 0u && nc) ? a.pptr[nh] : 0;
        h[H_NSH] = (a.has_prev && st == 0u) ? lower_i32(a.gt, ns, 1) : 0;
        h[H_NLAST] = st == 0u ? nc - lower_i32(a.t, nc, a.last_plane) : 0;
        h[H_TSHIFT] = (unsigned long long)a.t_shift;
        h[H_STATUS] = st | (fits || st != 0u ? 0u : ST_EXCHANGE);
        h[H_NRUNS] = a.totals[2];
        for (int i = H_CBASE; i < HDR_WORDS; ++i) h[i] = 0;
        for (int q = 0; q < a.ndst; ++q)
            for (int i = 0; i < HDR_WORDS; ++i) reinterpret_cast<unsigned long long*>(a.dst[q])[i] = h[i];
    }
    if (fits) {
        const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gsize = (long)gridDim.x * blockDim.x;
        const int nd = a.ndst;
        // every value is read once and stored to all destinations (NVLink stores are fire-and-forget: no round trip)
#define CT_PUT(T, k, i, v)                                                                          \
    do {                                                                                            \
        const T v_ = (v);                                                                           \
        for (int q = 0; q < nd; ++q) reinterpret_cast<T*>(a.dst[q] + HDR_BYTES + a.o.off[cts::k])[i] = v_; \
    } while (0)
        for (long i = gtid; i < nc; i += gsize) {
            CT_PUT(int32_t, A_T, i, a.t[i]); CT_PUT(int32_t, A_Y0, i, a.y0[i]); CT_PUT(int32_t, A_Y1, i, a.y1[i]);
            CT_PUT(int32_t, A_X0, i, a.x0[i]); CT_PUT(int32_t, A_X1, i, a.x1[i]); CT_PUT(uint32_t, A_CLS, i, a.cls[i]);
            CT_PUT(double, A_CONE, i, a.conE[i]); CT_PUT(double, A_CONS, i, a.conS[i]); CT_PUT(double, A_FE, i, a.fE[i]);
            CT_PUT(double, A_FS, i, a.fS[i]); CT_PUT(uint32_t, A_NSP, i, a.nsp[i]); CT_PUT(uint32_t, A_FNSP, i, a.fnsp[i]);
        }
        for (long i = gtid; i <= nc; i += gsize) CT_PUT(uint32_t, A_PPTR, i, nc ? a.pptr[i] : 0u);
        for (long i = gtid; i < np; i += gsize) {
            CT_PUT(uint32_t, A_PB, i, a.pb[i]); CT_PUT(uint32_t, A_PNPIX, i, a.pnpix[i]); CT_PUT(uint32_t, A_PNSP, i, a.pnsp[i]);
            CT_PUT(double, A_PE, i, a.pE[i]); CT_PUT(double, A_PS, i, a.pS[i]);
        }
        for (long i = gtid; i < ns; i += gsize) {
            CT_PUT(int32_t, A_GT, i, a.gt[i]); CT_PUT(int32_t, A_GY0, i, a.gy0[i]); CT_PUT(int32_t, A_GY1, i, a.gy1[i]);
            CT_PUT(uint32_t, A_GA, i, a.ga[i]); CT_PUT(uint32_t, A_GB, i, a.gb[i]);
        }
#undef CT_PUT
    }
    if (a.done_ctr) {
        // the last block through raises this rank's flag at every peer: stores of all blocks are ordered before it
        // (fence by every thread, block barrier, counter; then fence + release store by the last block)
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(a.done_ctr, 1u) == gridDim.x - 1) {
            *a.done_ctr = 0;                                          // (the next launch follows in stream order)
            __threadfence_system();
            for (int q = 0; q < a.ndst; ++q)
                if (a.flag[q]) st_release_sys(a.flag[q], a.epoch);
        }
    }
}

// one block: headers of all ranks -> per-rank descriptors (header + bases) and the control block of the global context
// (one warp).  With peer windows the kernel first waits until every rank's flag word carries this exchange's epoch.
__global__ void k_merge_desc(const char* gathered, size_t stride, int nranks, unsigned long long* mdesc /*[nranks * 16]*/,
                             uint32_t* gctl /*ticket, status, info, -, totals u64 x 4 at word 4*/, unsigned long long* need4,
                             unsigned long long capCg, unsigned long long capPg, unsigned long long capSg,
                             const unsigned long long* flags /*[nranks] or null*/, unsigned long long epoch) {
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    int late = 0;
    if (flags && (int)threadIdx.x < nranks) {
        const unsigned long long t0 = global_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
            __nanosleep(spins < 4096u ? 100u : 2000u);                // (one warp per GPU: cheap, and gentle when it is long)
            if ((++spins & 255u) == 0u && global_ns() - t0 > FLAG_WAIT_NS) { late = 1; break; }
        }
    }
    late = __any_sync(0xffffffffu, late);
    if (threadIdx.x != 0) return;
    __threadfence_system();
    unsigned long long NC = 0, NP = 0, NS = 0, NR = 0, mc = 0, mp = 0, ms = 0;
    uint32_t status = 0;
    if (late) {                                   // a rank never delivered its tables (it failed, or left the call early)
        unsigned long long* totals = reinterpret_cast<unsigned long long*>(gctl + 4);
        totals[0] = totals[1] = totals[2] = totals[3] = 0;
        gctl[0] = 0; gctl[1] = ctp::ST_TIMEOUT; gctl[2] = 0;
        for (int i = 0; i < nranks * HDR_WORDS; ++i) mdesc[i] = 0;
        need4[0] = need4[1] = need4[2] = 0; need4[3] = ctp::ST_TIMEOUT;
        return;
    }
    for (int r = 0; r < nranks; ++r) {
        const unsigned long long* h = reinterpret_cast<const unsigned long long*>(gathered + (size_t)r * stride);
        unsigned long long* m = mdesc + (size_t)r * HDR_WORDS;
        for (int i = 0; i < H_CBASE; ++i) m[i] = h[i];
        status |= (uint32_t)h[H_STATUS];
        mc = h[H_NC] > mc ? h[H_NC] : mc; mp = h[H_NP] > mp ? h[H_NP] : mp; ms = h[H_NS] > ms ? h[H_NS] : ms;
        if (r == 0 && h[H_NH] != 0) status |= ST_MISMATCH;
        if (r > 0 && h[H_STATUS] == 0 && mdesc[(size_t)(r - 1) * HDR_WORDS + H_STATUS] == 0 &&
            mdesc[(size_t)(r - 1) * HDR_WORDS + H_NLAST] != h[H_NH]) status |= ST_MISMATCH;
        if (h[H_NH] > h[H_NC] || h[H_E0] > h[H_NP] || h[H_NSH] > h[H_NS]) status |= ST_MISMATCH;
        m[H_CBASE] = NC; m[H_PBASE] = NP; m[H_SBASE] = NS; m[H_COFF] = NC - h[H_NH];
        m[14] = 0; m[15] = 0;
        if (h[H_STATUS] == 0) { NC += h[H_NC] - h[H_NH]; NP += h[H_NP] - h[H_E0]; NS += h[H_NS] - h[H_NSH]; }
        NR += h[H_NRUNS];
    }
    if (NC > capCg || NP > capPg || NS > capSg) status |= ctp::ST_CAPACITY;
    unsigned long long* totals = reinterpret_cast<unsigned long long*>(gctl + 4);
    totals[0] = NC; totals[1] = NS; totals[2] = NR; totals[3] = NP;
    gctl[0] = 0; gctl[1] = status; gctl[2] = 0;
    need4[0] = mc; need4[1] = mp; need4[2] = ms; need4[3] = status;
}

__device__ __forceinline__ int find_rank(const unsigned long long* mdesc, int nranks, int which, unsigned long long i) {
    int lo = 0, hi = nranks - 1;                  // last r with base[r] <= i (bases are non-decreasing, base[0] = 0)
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (mdesc[(size_t)mid * HDR_WORDS + which] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <typename T> __device__ __forceinline__ const T* arr(const char* gathered, size_t stride, int r, const Offsets& o, int k) {
    return reinterpret_cast<const T*>(gathered + (size_t)r * stride + HDR_BYTES + o.off[k]);
}

// upper-bound grids (nranks x capacity): every kernel reads the true totals from the global control block
__global__ void __launch_bounds__(256) k_merge_comps(const char* __restrict__ gathered, size_t stride, int nranks,
                                                     const unsigned long long* __restrict__ mdesc, const uint32_t* gctl, Offsets o,
                                                     cts::GlobalTables g) {
    if (gctl[1] != 0u) return;
    const unsigned long long* totals = reinterpret_cast<const unsigned long long*>(gctl + 4);
    const long NC = (long)totals[0], NP = (long)totals[3];
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > NC) return;
    if (i == NC) { g.pptr[NC] = (uint32_t)NP; return; }
    const int r = find_rank(mdesc, nranks, H_CBASE, (unsigned long long)i);
    const unsigned long long* d = mdesc + (size_t)r * HDR_WORDS;
    const long nh = (long)d[H_NH], cbase = (long)d[H_CBASE];
    const long j = i - cbase + nh;                     // local index
    g.t[i] = arr<int32_t>(gathered, stride, r, o, cts::A_T)[j] + (int32_t)(long)d[H_TSHIFT];
    g.y0[i] = arr<int32_t>(gathered, stride, r, o, cts::A_Y0)[j];
    g.y1[i] = arr<int32_t>(gathered, stride, r, o, cts::A_Y1)[j];
    g.x0[i] = arr<int32_t>(gathered, stride, r, o, cts::A_X0)[j];
    g.x1[i] = arr<int32_t>(gathered, stride, r, o, cts::A_X1)[j];
    g.cls[i] = (uint32_t)((long)arr<uint32_t>(gathered, stride, r, o, cts::A_CLS)[j] + cbase - nh);
    g.conE[i] = arr<double>(gathered, stride, r, o, cts::A_CONE)[j];
    g.conS[i] = arr<double>(gathered, stride, r, o, cts::A_CONS)[j];
    g.nsp[i] = arr<uint32_t>(gathered, stride, r, o, cts::A_NSP)[j];
    double fE = arr<double>(gathered, stride, r, o, cts::A_FE)[j], fS = arr<double>(gathered, stride, r, o, cts::A_FS)[j];
    uint32_t fn = arr<uint32_t>(gathered, stride, r, o, cts::A_FNSP)[j];
    if (r + 1 < nranks) {
        // forward overlap of this rank's last-plane classes with the next rank's first plane was accumulated over there,
        // on the halo copies of these components
        const unsigned long long* e = mdesc + (size_t)(r + 1) * HDR_WORDS;
        const long h = i - ((long)e[H_CBASE] - (long)e[H_NH]);
        if (h >= 0 && h < (long)e[H_NH]) {
            fE += arr<double>(gathered, stride, r + 1, o, cts::A_FE)[h];
            fS += arr<double>(gathered, stride, r + 1, o, cts::A_FS)[h];
            fn += arr<uint32_t>(gathered, stride, r + 1, o, cts::A_FNSP)[h];
        }
    }
    g.fE[i] = fE; g.fS[i] = fS; g.fnsp[i] = fn;
    g.pptr[i] = (uint32_t)((long)d[H_PBASE] + (long)arr<uint32_t>(gathered, stride, r, o, cts::A_PPTR)[j] - (long)d[H_E0]);
}

__global__ void __launch_bounds__(256) k_merge_pairs(const char* __restrict__ gathered, size_t stride, int nranks,
                                                     const unsigned long long* __restrict__ mdesc, const uint32_t* gctl, Offsets o,
                                                     cts::GlobalTables g) {
    if (gctl[1] != 0u) return;
    const long NP = (long)reinterpret_cast<const unsigned long long*>(gctl + 4)[3];
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    const int r = find_rank(mdesc, nranks, H_PBASE, (unsigned long long)p);
    const unsigned long long* d = mdesc + (size_t)r * HDR_WORDS;
    const long q = p - (long)d[H_PBASE] + (long)d[H_E0];
    g.p_b[p] = (uint32_t)((long)arr<uint32_t>(gathered, stride, r, o, cts::A_PB)[q] + (long)d[H_CBASE] - (long)d[H_NH]);
    g.p_npix[p] = arr<uint32_t>(gathered, stride, r, o, cts::A_PNPIX)[q];
    g.p_nsp[p] = arr<uint32_t>(gathered, stride, r, o, cts::A_PNSP)[q];
    g.p_E[p] = arr<double>(gathered, stride, r, o, cts::A_PE)[q];
    g.p_S[p] = arr<double>(gathered, stride, r, o, cts::A_PS)[q];
}

__global__ void __launch_bounds__(256) k_merge_segs(const char* __restrict__ gathered, size_t stride, int nranks,
                                                    const unsigned long long* __restrict__ mdesc, const uint32_t* gctl, Offsets o,
                                                    cts::GlobalTables g) {
    if (gctl[1] != 0u) return;
    const long NS = (long)reinterpret_cast<const unsigned long long*>(gctl + 4)[1];
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= NS) return;
    const int r = find_rank(mdesc, nranks, H_SBASE, (unsigned long long)s);
    const unsigned long long* d = mdesc + (size_t)r * HDR_WORDS;
    const long q = s - (long)d[H_SBASE] + (long)d[H_NSH];
    const long off = (long)d[H_CBASE] - (long)d[H_NH];
    g.g_t[s] = arr<int32_t>(gathered, stride, r, o, cts::A_GT)[q] + (int32_t)(long)d[H_TSHIFT];
    g.g_y0[s] = arr<int32_t>(gathered, stride, r, o, cts::A_GY0)[q];
    g.g_y1[s] = arr<int32_t>(gathered, stride, r, o, cts::A_GY1)[q];
    g.g_a[s] = (uint32_t)((long)arr<uint32_t>(gathered, stride, r, o, cts::A_GA)[q] + off);
    g.g_b[s] = (uint32_t)((long)arr<uint32_t>(gathered, stride, r, o, cts::A_GB)[q] + off);
}

inline unsigned blocks_for(size_t n) { return (unsigned)std::max<size_t>(1, (n + 255) / 256); }

// plane runs for the exact host replays (near-ties on non-exact rows, stale-box splits): the owner of the plane serves
// them to everybody.  Every rank replays the same ordered phase, so every rank asks for the same planes in the same order.
struct DistFetch {
    ct_ctx* c; ctc::Comm* comm; cudaStream_t st;
    const unsigned long long* mdesc;            // host copy [nranks * 16]
    std::vector<int32_t> y, x0, x1; std::vector<uint32_t> comp;
    DevBuf dbuf;
    std::string err;
};

int dist_fetch_cb(void* user, long t, long* n, const int32_t** y, const int32_t** x0, const int32_t** x1, const uint32_t** comp) {
    DistFetch* f = static_cast<DistFetch*>(user);
    const int nranks = f->comm->size(), rank = f->comm->rank();
    // owner: the last rank whose first own plane (t_shift + has_prev) is <= t
    int owner = 0;
    for (int r = 0; r < nranks; ++r) {
        const long first = (long)f->mdesc[(size_t)r * HDR_WORDS + H_TSHIFT] + (r > 0 ? 1 : 0);
        if (first <= t) owner = r;
    }
    std::vector<cth::PlaneRun> runs;
    long cnt = 0;
    if (rank == owner) {
        const long local = t - (long)f->mdesc[(size_t)rank * HDR_WORDS + H_TSHIFT];
        if (!cti::api_plane_runs(f->c, local, runs, f->st)) return -1;
        cnt = (long)runs.size();
    }
    // count, then (y, x0, x1, comp) as 4 x int32 per run
    if (f->dbuf.ensure(64) != cudaSuccess) return -1;
    if (rank == owner && cudaMemcpyAsync(f->dbuf.p, &cnt, 8, cudaMemcpyHostToDevice, f->st) != cudaSuccess) return -1;
    if (f->comm->bcast(f->dbuf.p, 8, owner, f->st)) return -1;
    if (cudaMemcpyAsync(&cnt, f->dbuf.p, 8, cudaMemcpyDeviceToHost, f->st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(f->st) != cudaSuccess) return -1;
    std::vector<int32_t> flat((size_t)cnt * 4);
    if (cnt) {
        if (f->dbuf.ensure((size_t)cnt * 16) != cudaSuccess) return -1;
        if (rank == owner) {
            const long off = (long)f->mdesc[(size_t)rank * HDR_WORDS + H_COFF];
            for (long i = 0; i < cnt; ++i) {
                flat[4 * i] = runs[i].y; flat[4 * i + 1] = runs[i].x0; flat[4 * i + 2] = runs[i].x1;
                flat[4 * i + 3] = (int32_t)((long)runs[i].comp + off);
            }
            if (cudaMemcpyAsync(f->dbuf.p, flat.data(), (size_t)cnt * 16, cudaMemcpyHostToDevice, f->st) != cudaSuccess) return -1;
        }
        if (f->comm->bcast(f->dbuf.p, (size_t)cnt * 16, owner, f->st)) return -1;
        if (cudaMemcpyAsync(flat.data(), f->dbuf.p, (size_t)cnt * 16, cudaMemcpyDeviceToHost, f->st) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(f->st) != cudaSuccess) return -1;
    }
    f->y.resize(cnt); f->x0.resize(cnt); f->x1.resize(cnt); f->comp.resize(cnt);
    for (long i = 0; i < cnt; ++i) {
        f->y[i] = flat[4 * i]; f->x0[i] = flat[4 * i + 1]; f->x1[i] = flat[4 * i + 2]; f->comp[i] = (uint32_t)flat[4 * i + 3];
    }
    *n = cnt; *y = f->y.data(); *x0 = f->x0.data(); *x1 = f->x1.data(); *comp = f->comp.data();
    return 0;
}

int q_sum_u32(void* user, uint32_t* buf, size_t n, cudaStream_t st) {
    return static_cast<ctc::Comm*>(user)->allreduce(buf, n, ctc::SUM_U32, st);
}
int q_min_i64(void* user, long long* buf, size_t n, cudaStream_t st) {
    return static_cast<ctc::Comm*>(user)->allreduce(buf, n, ctc::MIN_I64, st);
}

int comm_fail(ctc::Comm* comm, const char* what) { return fail(CT_ERR_COMM, "%s: %s", what, comm->err.c_str()); }

}  // namespace

extern "C" {

// ---- communicators ----------------------------------------------------------------------------------------------------
int ct_nccl_unique_id(unsigned char id[128]) {
    std::string err;
    if (ctc::nccl_unique_id(id, err)) return fail(CT_ERR_COMM, "%s", err.c_str());
    return CT_OK;
}

int ct_comm_init_nccl(const unsigned char id[128], int rank, int nranks, int device, ct_comm** out) {
    if (!id || !out || nranks < 1 || rank < 0 || rank >= nranks) return fail(CT_ERR_ARG, "bad communicator arguments");
    *out = nullptr;
    CT_CUDA(cudaSetDevice(device));
    std::string err;
    ctc::Comm* c = ctc::nccl_create(id, rank, nranks, err);
    if (!c) return fail(CT_ERR_COMM, "%s", err.c_str());
    ct_comm* h = new ct_comm();
    h->impl = c;
    *out = h;
    return CT_OK;
}

int ct_comm_from_nccl(void* nccl_comm, int rank, int nranks, ct_comm** out) {
    if (!nccl_comm || !out || nranks < 1 || rank < 0 || rank >= nranks) return fail(CT_ERR_ARG, "bad communicator arguments");
    std::string err;
    ctc::Comm* c = ctc::nccl_wrap(nccl_comm, rank, nranks, err);
    if (!c) return fail(CT_ERR_COMM, "%s", err.c_str());
    ct_comm* h = new ct_comm();
    h->impl = c;
    *out = h;
    return CT_OK;
}

int ct_comm_init_local(int nranks, ct_comm** out /* [nranks] */) {
    if (!out || nranks < 1 || nranks > 64) return fail(CT_ERR_ARG, "an in-process group has 1..64 ranks");
    ctc::LocalGroup* g = ctc::local_group_create(nranks);
    for (int r = 0; r < nranks; ++r) {
        ct_comm* h = new ct_comm();
        h->impl = ctc::local_comm(g, r);
        h->group = r == 0 ? g : nullptr;
        out[r] = h;
    }
    return CT_OK;
}

void ct_comm_destroy(ct_comm* comm) {
    if (!comm) return;
    if (comm->window) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (comm->window_device >= 0) cudaSetDevice(comm->window_device);
        cudaDeviceSynchronize();
        delete comm->impl;                     // (closes its mappings of the other ranks' windows first)
        comm->impl = nullptr;
        cudaFree(comm->window);
        cudaSetDevice(dev);
    }
    delete comm->impl;
    if (comm->group) ctc::local_group_destroy(comm->group);
    delete comm;
}

int ct_comm_rank(ct_comm* comm) { return comm && comm->impl ? comm->impl->rank() : -1; }
int ct_comm_size(ct_comm* comm) { return comm && comm->impl ? comm->impl->size() : -1; }

// ---- quantile over time, optionally over the time shards of all ranks (README.rst:150-151) -------------------------------
int ct_quantile_time_t(ct_ctx* c, ct_comm* comm_h, const void* x_dev, int dtype, long T_local, int H, int W, int y0, int y1,
                       const double* q_host, int nq, double* out_dev, void* stream) {
    if (!c || !x_dev || !q_host || !out_dev) return fail(CT_ERR_ARG, "null argument");
    if (dtype != CT_F32 && dtype != CT_F64) return fail(CT_ERR_ARG, "dtype must be CT_F32 or CT_F64");
    if (T_local <= 0 || H <= 0 || W <= 0 || y0 < 0 || y1 > H || y0 >= y1 || nq <= 0) return fail(CT_ERR_ARG, "bad shape / row range");
    for (int i = 0; i < nq; ++i)
        if (!(q_host[i] >= 0.0 && q_host[i] <= 1.0)) return fail(CT_ERR_ARG, "Quantiles must be in the range [0, 1]");
    CT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long npts = (long)(y1 - y0) * W;
    CT_CUDA(c->x_q.ensure((size_t)nq * 8));
    CT_CUDA(c->x_qscratch.ensure(cte::quantile_scratch_bytes(npts, nq, dtype == CT_F64)));
    CT_CUDA(cudaMemcpyAsync(c->x_q.p, q_host, (size_t)nq * 8, cudaMemcpyHostToDevice, st));
    CT_CUDA(cudaStreamSynchronize(st));                               // q_host may be a temporary of the caller
    cte::QuantileReduce red{q_sum_u32, q_min_i64, comm_h ? comm_h->impl : nullptr};
    const cudaError_t e = cte::quantile_time(x_dev, dtype == CT_F64, T_local, H, W, y0, y1, c->x_q.as<double>(), nq, out_dev,
                                             c->x_qscratch.p, comm_h && comm_h->impl && comm_h->impl->size() > 1 ? &red : nullptr, st);
    if (e != cudaSuccess) {
        if (comm_h && comm_h->impl && !comm_h->impl->err.empty()) return comm_fail(comm_h->impl, "quantile all-reduce");
        return fail(CT_ERR_CUDA, "quantile_time failed: %s", cudaGetErrorString(e));
    }
    return CT_OK;
}

// ---- the sharded run ----------------------------------------------------------------------------------------------------
}  // extern "C"

// Peer windows of the table exchange: one allocation per rank = flag words + two gathered buffers (double buffer by exchange
// parity: a fast rank may already be packing exchange k+1 into a peer that is still merging exchange k).  Mapped once per
// communicator, re-made (collectively) when the stride outgrows it.  All ranks take the same branches: the sizes derive from
// the negotiated capacities, which are equal everywhere.
static int window_prepare(ct_ctx* c, ct_comm* h, size_t stride, cudaStream_t ts) {
    ctc::Comm* comm = h->impl;
    const int nranks = comm->size();
    const size_t need = WIN_DATA + 2 * (size_t)nranks * stride;
    if (h->window_mode == 2) return CT_OK;
    if (h->window_mode == 1 && need <= h->window_bytes) return CT_OK;
    if (h->window_mode == 1) {
        CT_CUDA(cudaStreamSynchronize(ts));
        if (comm->window_unmap(ts)) return fail(CT_ERR_COMM, "peer windows: %s", comm->err.c_str());
        CT_CUDA(cudaFree(h->window));
        h->window = nullptr; h->window_bytes = 0; h->window_mode = 0;
    }
    const size_t bytes = need + need / 4;
    CT_CUDA(cudaMalloc(&h->window, bytes));
    CT_CUDA(cudaMemsetAsync(h->window, 0, WIN_DATA, ts));
    CT_CUDA(cudaStreamSynchronize(ts));
    h->window_device = c->device;
    h->peers.assign((size_t)nranks, nullptr);
    const int r = comm->window_map(h->window, bytes, h->peers.data(), ts);
    if (r < 0) return fail(CT_ERR_COMM, "peer windows: %s", comm->err.c_str());
    if (r == 1) {                                                      // no peer memory between some pair of ranks
        CT_CUDA(cudaFree(h->window));
        h->window = nullptr; h->window_mode = 2;
        return CT_OK;
    }
    h->window_bytes = bytes; h->window_mode = 1; h->epoch = 0;
    return CT_OK;
}

// Device buffers (anom_dev / flag_dev) or host buffers (anom_host / flag_host): with host buffers the shard is streamed in
// time chunks host -> device under the threshold kernel (the float shard is never resident) and the result leaves as the
// row-run table, expanded by host threads into flag_host, which they zero while the input streams in.
static int sharded_run(ct_ctx* c, ct_comm* comm_h, const void* anom_dev, const void* anom_host, int in_dtype, long T_local,
                       long t_begin, long T_total, int H, int W, const double* w_host, const double* thr_host, long thr_n,
                       int thr_is_f32, int op, double overlap, int persistence, int twosided, int32_t* flag_dev,
                       int32_t* flag_host, long chunk_planes, long* n_features, void* stream) {
    if (!c || !comm_h || !comm_h->impl) return fail(CT_ERR_ARG, "null context / communicator");
    const bool host_io = anom_host != nullptr;
    ctc::Comm* comm = comm_h->impl;
    const int rank = comm->rank(), nranks = comm->size();
    int rc = cti::api_check_args(T_local, H, W, w_host, thr_host, thr_n, in_dtype, op);
    if (rc != CT_OK) return rc;
    if (T_local <= 0 || (host_io ? (!anom_host || !flag_host) : (!anom_dev || !flag_dev)))
        return fail(CT_ERR_ARG, "a rank needs at least one plane and both cubes");
    if (t_begin < 0 || t_begin + T_local > T_total) return fail(CT_ERR_ARG, "planes [%ld, %ld) outside the cube of %ld", t_begin,
                                                                t_begin + T_local, T_total);
    if (n_features) *n_features = 0;
    const int hp = rank > 0 ? 1 : 0;
    CT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    c->stats.clear();
    std::vector<double> thr(thr_host, thr_host + thr_n);
    if (thr_n != 1 && hp) thr.insert(thr.begin(), 0.0);               // thresholds are indexed by scratch plane
    const long planes = T_local + hp;
    if ((rc = cti::api_prepare(c, planes, H, W, w_host, thr.data(), (long)thr.size(), st)) != CT_OK) return rc;
    c->has_prev = hp;
    if ((rc = cti::api_ensure_streams(c)) != CT_OK) return rc;
    if (!c->copy_stream) {
        int lo = 0, hi = 0;
        CT_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CT_CUDA(cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, hi));
    }
    for (auto& e : c->ev_x) if (!e) CT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (!c->gctx) {
        if ((rc = ct_create(c->device, &c->gctx)) != CT_OK) return rc;
    }
    ct_ctx* g = c->gctx;
    g->opt_max_sweeps = c->opt_max_sweeps;

    cudaStream_t ts = c->tbl_stream, aux = c->copy_stream, side = c->side_stream;
    const size_t plane_bytes = (size_t)H * W * (in_dtype == CT_F64 ? 8 : 4);
    const size_t words = (size_t)H * c->Ww;
    const double t_h0 = cti::now_ms();

    // ---- 1. last own plane first; its bit rows travel to the next rank while the other planes are thresholded ----
    std::vector<std::thread> zero_threads;
    struct Joiner {
        std::vector<std::thread>& v;
        ~Joiner() { for (auto& t : v) if (t.joinable()) t.join(); }
    } joiner{zero_threads};
    const size_t cells = (size_t)T_local * H * W;
    if (host_io) {
        if (!c->work_stream) CT_CUDA(cudaStreamCreateWithFlags(&c->work_stream, cudaStreamNonBlocking));
        if (!c->opt_host_out_zeroed) cti::api_host_zero_start(c, flag_host, cells, nranks, zero_threads);
        CT_CUDA(c->sh_lastplane.ensure(plane_bytes));
    }
    CT_CUDA(cudaEventRecord(c->ev[0], st));
    {
        const void* last = (const char*)(host_io ? anom_host : anom_dev) + (size_t)(T_local - 1) * plane_bytes;
        if (host_io) {
            CT_CUDA(cudaMemcpyAsync(c->sh_lastplane.p, last, plane_bytes, cudaMemcpyHostToDevice, st));
            last = c->sh_lastplane.p;
        }
        if ((rc = cti::api_launch_threshold(c, last, in_dtype, planes - 1, 1, (long)thr.size(), thr_is_f32, op, st, 1)) != CT_OK)
            return rc;
    }
    CT_CUDA(cudaEventRecord(c->ev_x[0], st));
    CT_CUDA(cudaStreamWaitEvent(aux, c->ev_x[0], 0));
    if (nranks > 1) {
        uint32_t* bits = c->bits.as<uint32_t>();
        if (comm->sendrecv(bits + (size_t)(planes - 1) * words, rank + 1 < nranks ? rank + 1 : -1, bits, hp ? rank - 1 : -1,
                           words * 4, aux)) return comm_fail(comm, "halo exchange");
        c->launches += 1;
    }
    if (hp) {
        CT_CUDA(ctk::row_stats(c->bits.as<uint32_t>(), H, W, c->Ww, c->row_cnt.as<uint32_t>(), c->seam_flag.as<uint32_t>(),
                               c->slots.as<uint32_t>(), c->counters.as<uint32_t>() + 16, aux));
        c->launches += 1;
    }
    CT_CUDA(cudaEventRecord(c->ev_x[1], aux));
    if (T_local > 1 && !host_io) {
        if ((rc = cti::api_launch_threshold(c, anom_dev, in_dtype, hp, T_local - 1, (long)thr.size(), thr_is_f32, op, st, 0)) != CT_OK)
            return rc;
    } else if (T_local > 1) {
        // host shard: copy chunk k + 1 while chunk k is thresholded (two staging buffers)
        const long Tm = T_local - 1;
        if (chunk_planes <= 0) chunk_planes = std::max<long>(1, (long)((256u << 20) / plane_bytes));
        chunk_planes = std::min(chunk_planes, Tm);
        const long nchunks = (Tm + chunk_planes - 1) / chunk_planes;
        cudaStream_t cs = c->work_stream;
        cudaEvent_t in_ready[2], in_free[2];
        for (int i = 0; i < 2; ++i) {
            CT_CUDA(cudaEventCreateWithFlags(&in_ready[i], cudaEventDisableTiming));
            CT_CUDA(cudaEventCreateWithFlags(&in_free[i], cudaEventDisableTiming));
            CT_CUDA(c->chunk_in[i].ensure((size_t)chunk_planes * plane_bytes));
        }
        CT_CUDA(cudaStreamWaitEvent(cs, c->ev_x[0], 0));                 // (orders the copies after whatever `st` did before)
        for (long k = 0; k < nchunks; ++k) {
            const int b = (int)(k & 1);
            const long t0 = k * chunk_planes, nt = std::min(chunk_planes, Tm - t0);
            if (k >= 2) CT_CUDA(cudaStreamWaitEvent(cs, in_free[b], 0));
            CT_CUDA(cudaMemcpyAsync(c->chunk_in[b].p, (const char*)anom_host + (size_t)t0 * plane_bytes, (size_t)nt * plane_bytes,
                                    cudaMemcpyHostToDevice, cs));
            CT_CUDA(cudaEventRecord(in_ready[b], cs));
            CT_CUDA(cudaStreamWaitEvent(st, in_ready[b], 0));
            if ((rc = cti::api_launch_threshold(c, c->chunk_in[b].p, in_dtype, hp + t0, nt, (long)thr.size(), thr_is_f32, op, st, 0)) != CT_OK)
                return rc;
            CT_CUDA(cudaEventRecord(in_free[b], st));
        }
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(in_ready[i]); cudaEventDestroy(in_free[i]); }
    }
    CT_CUDA(cudaEventRecord(c->ev[1], st));
    // ---- zero fill of the own flag planes on the low-priority stream, beside everything that follows ----
    CT_CUDA(cudaEventRecord(c->ev_side[0], st));
    CT_CUDA(cudaStreamWaitEvent(side, c->ev_side[0], 0));
    c->pend_fill = nullptr; c->pend_fill_cells = 0;
    const bool plane_first = c->opt_gpu_tables &&
                             (c->opt_plane_kernel == 1 || (c->opt_plane_kernel == 2 && planes <= c->opt_plane_max_planes));
    if (!host_io) {
        if (c->opt_fill_late && plane_first) {
            c->pend_fill = flag_dev; c->pend_fill_cells = cells;      // started by ctf::finish(), after the plane kernel
        } else {
            CT_CUDA(ctk::zero_fill(flag_dev, cells, c->sm_count, side, cti::fill_ctas(c, plane_first)));
            c->launches += 1;
        }
    }
    CT_CUDA(cudaEventRecord(c->ev_side[1], side));
    CT_CUDA(cudaStreamWaitEvent(ts, c->ev_side[0], 0));
    CT_CUDA(cudaStreamWaitEvent(ts, c->ev_x[1], 0));

    // ---- 2..5, repeated when a table or the exchange stride turns out too small ----
    // table builder by shard size, like the single-GPU call: the plane kernel for short shards, the global-memory kernels for
    // long ones (they run beside the zero fill, which is long enough to hide them there)
    bool local_ok = false;
    bool classic = !(c->opt_gpu_tables &&
                     (c->opt_plane_kernel == 1 || (c->opt_plane_kernel == 2 && planes <= c->opt_plane_max_planes)));
    int outcome = ctf::FAST_SLOW;
    std::vector<unsigned long long> mdesc_host((size_t)nranks * HDR_WORDS);
    CT_CUDA(c->hp_hdr.ensure((size_t)nranks * HDR_BYTES + 256));
    int attempts = 0;
    for (;; ++attempts) {
        if (attempts >= 6) return fail(CT_ERR_INTERNAL, "sharded run: tables still do not fit after %d attempts", attempts);
        // ---- local tables ----
        if (!local_ok) {
            if (classic) {
                if ((rc = ctf::ensure_control(c)) != CT_OK) return rc;
                if ((rc = cti::api_classic_tables(c, ts)) != CT_OK) return rc;          // (synchronises; counts in the context)
                if ((rc = ctf::ensure_tables(c, 0, (size_t)c->ncomp, (size_t)c->npair, (size_t)c->nseg)) != CT_OK) return rc;
                unsigned long long tot[4] = {(unsigned long long)c->ncomp, (unsigned long long)c->nseg, (unsigned long long)c->nruns,
                                             (unsigned long long)c->npair};
                uint32_t zero3[3] = {0, 0, 0};
                CT_CUDA(cudaMemcpyAsync(c->pl_ctl.p, zero3, 12, cudaMemcpyHostToDevice, ts));
                CT_CUDA(cudaMemcpyAsync(c->pl_ctl.as<char>() + 16, tot, 32, cudaMemcpyHostToDevice, ts));
                CT_CUDA(cudaStreamSynchronize(ts));
                c->fast_tables = 0;
            } else {
                if ((rc = ctf::begin(c, planes, ts)) != CT_OK) return rc;
                if ((rc = ctf::chunk(c, 0, planes, ts)) != CT_OK) return rc;
                if ((rc = ctf::finish(c, ts)) != CT_OK) return rc;
                c->fast_tables = 1;
            }
        }
        // ---- exchange stride: negotiated once per problem (the only extra round trip, first call only) ----
        if (!comm_h->capC) {
            int oc = ctf::FAST_SLOW;
            if (!classic && (rc = ctf::totals_to_host(c, ts, &oc)) != CT_OK) return rc;
            unsigned long long mine[4] = {(unsigned long long)c->ncomp, (unsigned long long)c->npair, (unsigned long long)c->nseg, 0};
            CT_CUDA(c->sh_mdesc.ensure((size_t)(nranks + 1) * HDR_BYTES + 256));
            unsigned long long* d = c->sh_mdesc.as<unsigned long long>();
            CT_CUDA(cudaMemcpyAsync(d + 4 * nranks, mine, 32, cudaMemcpyHostToDevice, ts));
            if (comm->allgather(d + 4 * nranks, d, 32, ts)) return comm_fail(comm, "size exchange");
            std::vector<unsigned long long> all((size_t)4 * nranks);
            CT_CUDA(cudaMemcpyAsync(all.data(), d, (size_t)32 * nranks, cudaMemcpyDeviceToHost, ts));
            CT_CUDA(cudaStreamSynchronize(ts));
            unsigned long long mc = 0, mp = 0, ms = 0;
            for (int r = 0; r < nranks; ++r) { mc = std::max(mc, all[4 * r]); mp = std::max(mp, all[4 * r + 1]); ms = std::max(ms, all[4 * r + 2]); }
            comm_h->capC = (long)(mc + mc / 8 + 512); comm_h->capP = (long)(mp + mp / 8 + 512); comm_h->capS = (long)(ms + ms / 8 + 512);
            c->stats["exchange_negotiated"] = 1.0;
            // (a local capacity retry / fallback is reported through the header and handled below like any other)
        }
        Offsets o;
        const size_t stride = (HDR_BYTES + cts::layout(comm_h->capC, comm_h->capP, comm_h->capS, o.off) + 255) / 256 * 256;
        CT_CUDA(c->sh_mdesc.ensure((size_t)(nranks + 1) * HDR_BYTES + 256));
        // peer windows (pack kernel stores into every rank's gathered buffer) or local slot + all-gather
        if (c->opt_p2p && nranks > 1 && nranks <= MAX_PEERS && (rc = window_prepare(c, comm_h, stride, ts)) != CT_OK) return rc;
        const bool p2p = c->opt_p2p && nranks > 1 && nranks <= MAX_PEERS && comm_h->window_mode == 1;
        const char* gathered = nullptr;
        const unsigned long long* wait_flags = nullptr;
        unsigned long long epoch = 0;
        if (!p2p) {
            CT_CUDA(c->sh_export.ensure(stride));
            CT_CUDA(c->sh_gathered.ensure(stride * nranks));
            gathered = c->sh_gathered.as<char>();
        }
        // ---- 3. pack + all-gather ----
        {
            PackArgs a;
            auto U = [](DevBuf& b) { return b.as<uint32_t>(); };
            a.t = c->c_t.as<int32_t>(); a.y0 = c->c_y0.as<int32_t>(); a.y1 = c->c_y1.as<int32_t>(); a.x0 = c->c_x0.as<int32_t>();
            a.x1 = c->c_x1.as<int32_t>(); a.cls = U(c->c_cls); a.conE = c->k_conE.as<double>(); a.conS = c->k_conS.as<double>();
            a.fE = c->k_fE.as<double>(); a.fS = c->k_fS.as<double>(); a.nsp = U(c->k_nsp); a.fnsp = U(c->k_fnsp); a.pptr = U(c->pptr);
            a.pb = U(c->p_b); a.pnpix = U(c->p_npix); a.pnsp = U(c->p_nsp); a.pE = c->p_E.as<double>(); a.pS = c->p_S.as<double>();
            a.gt = c->g_t.as<int32_t>(); a.gy0 = c->g_y0.as<int32_t>(); a.gy1 = c->g_y1.as<int32_t>(); a.ga = U(c->g_a); a.gb = U(c->g_b);
            a.totals = reinterpret_cast<const unsigned long long*>(c->pl_ctl.as<char>() + 16);
            a.status = U(c->pl_ctl) + 1;
            a.capC = (unsigned long long)comm_h->capC; a.capP = (unsigned long long)comm_h->capP; a.capS = (unsigned long long)comm_h->capS;
            a.has_prev = hp; a.last_plane = (int)(planes - 1); a.t_shift = t_begin - hp;
            a.o = o;
            for (int q = 0; q < MAX_PEERS; ++q) { a.dst[q] = nullptr; a.flag[q] = nullptr; }
            if (p2p) {
                epoch = ++comm_h->epoch;
                const size_t par = (size_t)(epoch & 1);
                const size_t buf = WIN_DATA + par * (size_t)nranks * stride;
                const bool flags = comm->window_device_flags();
                for (int q = 0; q < nranks; ++q) {
                    char* w = static_cast<char*>(comm_h->peers[(size_t)q]);
                    a.dst[q] = w + buf + (size_t)rank * stride;
                    if (flags) a.flag[q] = reinterpret_cast<unsigned long long*>(w) + par * MAX_PEERS + rank;
                }
                a.ndst = nranks; a.epoch = epoch;
                a.done_ctr = flags ? reinterpret_cast<uint32_t*>(static_cast<char*>(comm_h->window) + 1024) : nullptr;
                gathered = static_cast<char*>(comm_h->window) + buf;
                if (flags) wait_flags = reinterpret_cast<unsigned long long*>(comm_h->window) + par * MAX_PEERS;
            } else {
                a.dst[0] = c->sh_export.as<char>(); a.ndst = 1; a.epoch = 0; a.done_ctr = nullptr;
            }
            k_pack_tables<<<c->sm_count * 2, 256, 0, ts>>>(a);
            CT_CUDA(cudaGetLastError());
            c->launches += 1;
        }
        if (p2p) {
            if (!comm->window_device_flags() && comm->window_fence(ts)) return comm_fail(comm, "peer window fence");
        } else {
            if (comm->allgather(c->sh_export.p, c->sh_gathered.p, stride, ts)) return comm_fail(comm, "table all-gather");
            c->launches += 1;
        }
        c->stats["p2p"] = p2p ? 1.0 : 0.0;
        // ---- 4. merge into the global context ----
        const size_t gC = (size_t)nranks * comm_h->capC, gP = (size_t)nranks * comm_h->capP, gS = (size_t)nranks * comm_h->capS;
        if ((rc = ctf::ensure_tables(g, 0, gC, gP, gS)) != CT_OK) return rc;
        if ((rc = ctf::ensure_control(g)) != CT_OK) return rc;
        g->T = T_total; g->H = H; g->W = W; g->Ww = c->Ww; g->special_uniform = c->special_uniform;
        g->w_host = c->w_host;
        for (auto& e : g->ev) if (!e) CT_CUDA(cudaEventCreate(&e));
        g->launches = 0; g->stats.clear();
        {
            cts::GlobalTables gt;
            auto U = [](DevBuf& b) { return b.as<uint32_t>(); };
            gt.t = g->c_t.as<int32_t>(); gt.y0 = g->c_y0.as<int32_t>(); gt.y1 = g->c_y1.as<int32_t>(); gt.x0 = g->c_x0.as<int32_t>();
            gt.x1 = g->c_x1.as<int32_t>(); gt.cls = U(g->c_cls); gt.conE = g->k_conE.as<double>(); gt.conS = g->k_conS.as<double>();
            gt.fE = g->k_fE.as<double>(); gt.fS = g->k_fS.as<double>(); gt.nsp = U(g->k_nsp); gt.fnsp = U(g->k_fnsp); gt.pptr = U(g->pptr);
            gt.p_b = U(g->p_b); gt.p_npix = U(g->p_npix); gt.p_nsp = U(g->p_nsp); gt.p_E = g->p_E.as<double>(); gt.p_S = g->p_S.as<double>();
            gt.g_t = g->g_t.as<int32_t>(); gt.g_y0 = g->g_y0.as<int32_t>(); gt.g_y1 = g->g_y1.as<int32_t>(); gt.g_a = U(g->g_a); gt.g_b = U(g->g_b);
            unsigned long long* md = c->sh_mdesc.as<unsigned long long>();
            unsigned long long* need4 = md + (size_t)nranks * HDR_WORDS;
            k_merge_desc<<<1, 32, 0, ts>>>(gathered, stride, nranks, md, U(g->pl_ctl), need4, gC, gP, gS, wait_flags, epoch);
            k_merge_comps<<<blocks_for(gC + 1), 256, 0, ts>>>(gathered, stride, nranks, md, U(g->pl_ctl), o, gt);
            k_merge_pairs<<<blocks_for(gP), 256, 0, ts>>>(gathered, stride, nranks, md, U(g->pl_ctl), o, gt);
            k_merge_segs<<<blocks_for(gS), 256, 0, ts>>>(gathered, stride, nranks, md, U(g->pl_ctl), o, gt);
            CT_CUDA(cudaGetLastError());
            g->launches += 4;
            // descriptors (+ need) and this rank's own control block come back with the global phase's one synchronisation
            // (queued behind the global kernel, not in front of it)
            g->extra_d2h[0] = {c->hp_hdr.p, md, (size_t)nranks * HDR_BYTES + 32};
            g->extra_d2h[1] = {c->hp_ctl.p, c->pl_ctl.p, 128};
        }
        // ---- 5. global phase on the merged tables (one synchronisation) ----
        uint32_t gstatus = 0;
        if ((rc = ctf::global(g, T_total, overlap, persistence, twosided, n_features, ts, &outcome, &gstatus)) != CT_OK) return rc;
        memcpy(mdesc_host.data(), c->hp_hdr.p, (size_t)nranks * HDR_BYTES);
        const unsigned long long* need4 = reinterpret_cast<const unsigned long long*>(c->hp_hdr.as<char>() + (size_t)nranks * HDR_BYTES);
        if (outcome != ctf::FAST_STATUS) break;
        // ---- somebody's tables were incomplete: every rank sees the same status word and takes the same decision ----
        if (gstatus & (ctp::ST_TIMEOUT | ST_MISMATCH))
            return fail(CT_ERR_INTERNAL, "sharded run: inconsistent rank tables (status %u)", gstatus);
        const uint32_t* myctl = c->hp_ctl.as<uint32_t>();
        const uint32_t mystatus = c->fast_tables ? myctl[1] : 0u;
        if (mystatus & ctp::ST_CAPACITY) {                          // my plane kernel ran out of table space: exact sizes now
            const unsigned long long* tot = reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(myctl) + 16);
            auto grow = [](unsigned long long v) { return (size_t)(v + v / 8 + 1024); };
            if ((rc = ctf::ensure_tables(c, grow(tot[2]), grow(tot[0]), grow(tot[3]), grow(tot[1]))) != CT_OK) return rc;
            local_ok = false;
        } else if (mystatus & ctp::ST_FALLBACK) {                    // a plane of mine does not fit shared memory
            if (!ctf::next_budget(c)) classic = true;
            local_ok = false;
        } else {
            local_ok = true;
        }
        if (gstatus & (ST_EXCHANGE | ctp::ST_CAPACITY)) {            // the stride (or the global tables) must grow
            comm_h->capC = (long)std::max<unsigned long long>(comm_h->capC, need4[0] + need4[0] / 4 + 1024);
            comm_h->capP = (long)std::max<unsigned long long>(comm_h->capP, need4[1] + need4[1] / 4 + 1024);
            comm_h->capS = (long)std::max<unsigned long long>(comm_h->capS, need4[2] + need4[2] / 4 + 1024);
        }
    }
    c->stats["shard_attempts"] = (double)(attempts + 1);
    const long comp_off = (long)mdesc_host[(size_t)rank * HDR_WORDS + H_COFF];
    c->nruns = (long)mdesc_host[(size_t)rank * HDR_WORDS + H_NRUNS];
    c->ncomp = (long)mdesc_host[(size_t)rank * HDR_WORDS + H_NC];
    if (outcome == ctf::FAST_SLOW) {
        // near-tie on non-exact rows / a label straddling a stale box: the ordered host replay on the merged tables, plane
        // runs served collectively by their owners
        DistFetch f;
        f.c = c; f.comm = comm; f.st = ts; f.mdesc = mdesc_host.data();
        g->fetch_fn = dist_fetch_cb; g->fetch_user = &f;
        g->fast_tables = 0; g->has_prev = 0; g->nruns = 0; g->nseam = 0; g->novr = 0;
        CT_CUDA(g->counters.ensure(128)); CT_CUDA(g->hp_counters.ensure(128)); CT_CUDA(g->run_val.ensure(16));
        rc = cti::api_table_phase(g, overlap, persistence, twosided, CT_STAGE_FINAL, n_features, ts);
        g->fetch_fn = nullptr; g->fetch_user = nullptr;
        f.dbuf.release();
        if (rc != CT_OK) return rc;
    } else if (outcome != ctf::FAST_OK) {
        return fail(CT_ERR_INTERNAL, "sharded run: unexpected outcome %d of the global phase", outcome);
    }
    c->stats["ms_h_tables"] = cti::now_ms() - t_h0;
    // ---- 6. paint the own planes: value of local component i = value of global component i + comp_off ----
    std::vector<ctb::Override> ovr;
    for (const ctb::Override& ov : g->host_result.overrides)
        if (ov.t >= t_begin && ov.t < t_begin + T_local) ovr.push_back(ctb::Override{(int32_t)(ov.t - t_begin), ov.y, ov.x0, ov.x1, ov.val});
    if (c->pend_fill) {                                                // (local tables came from the fallback kernels)
        CT_CUDA(ctk::zero_fill(c->pend_fill, c->pend_fill_cells, c->sm_count, side, cti::fill_ctas(c, false)));
        CT_CUDA(cudaEventRecord(c->ev_side[1], side));
        c->launches += 1;
        c->pend_fill = nullptr;
    }
    CT_CUDA(cudaEventRecord(c->ev_tbl[0], ts));
    CT_CUDA(cudaStreamWaitEvent(st, c->ev_tbl[0], 0));
    CT_CUDA(cudaStreamWaitEvent(st, c->ev_side[1], 0));
    CT_CUDA(cudaEventRecord(c->ev[3], st));
    if (!host_io) {
        ctk::PaintArgs a;
        const long r0 = (long)hp * H;
        a.bits = nullptr; a.row_ptr = c->row_ptr.as<uint32_t>() + r0; a.run_val = nullptr;
        a.nrows = T_local * H; a.W = W; a.Ww = c->Ww; a.flag = flag_dev; a.sparse = 1;
        a.run_x = c->run_x.as<uint32_t>(); a.run_row = c->run_row.as<uint32_t>(); a.row0 = r0;
        a.run_comp = c->run_comp.as<uint32_t>(); a.comp_val = g->c_val.as<int32_t>() + comp_off;
        CT_CUDA(ctk::paint(a, c->sm_count, st));
        c->launches += 1;
        if (!ovr.empty()) {
            const long novr = (long)ovr.size();
            CT_CUDA(c->hp_ovr.ensure((size_t)novr * 5 * 4));
            int32_t* ho = c->hp_ovr.as<int32_t>();
            for (long i = 0; i < novr; ++i) {
                ho[i] = ovr[i].t; ho[novr + i] = ovr[i].y; ho[2 * novr + i] = ovr[i].x0; ho[3 * novr + i] = ovr[i].x1; ho[4 * novr + i] = ovr[i].val;
            }
            DevBuf* ob5[] = {&c->o_t, &c->o_y, &c->o_x0, &c->o_x1, &c->o_val};
            for (int k = 0; k < 5; ++k) {
                CT_CUDA(ob5[k]->ensure((size_t)novr * 4));
                CT_CUDA(cudaMemcpyAsync(ob5[k]->p, ho + (size_t)k * novr, (size_t)novr * 4, cudaMemcpyHostToDevice, st));
            }
            CT_CUDA(ctk::paint_overrides(c->o_t.as<int32_t>(), c->o_y.as<int32_t>(), c->o_x0.as<int32_t>(), c->o_x1.as<int32_t>(),
                                         c->o_val.as<int32_t>(), novr, H, W, 0, T_local, flag_dev, st));
            c->launches += 1;
        }
    } else {
        // ---- the result leaves as the row-run table (12 B per run); host threads expand the surviving runs ----
        const long R = c->nruns;
        CT_CUDA(c->run_val.ensure((size_t)(R + 1) * 4));
        CT_CUDA(ctk::run_values(c->run_comp.as<uint32_t>(), g->c_val.as<int32_t>() + comp_off, c->run_val.as<int32_t>(), R, st));
        c->launches += 1;
        CT_CUDA(c->hp_runs.ensure((size_t)(R + 1) * 12));
        uint32_t* h_x = c->hp_runs.as<uint32_t>();
        uint32_t* h_row = h_x + R;
        int32_t* h_val = reinterpret_cast<int32_t*>(h_row + R);
        if (R) {
            CT_CUDA(cudaMemcpyAsync(h_x, c->run_x.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
            CT_CUDA(cudaMemcpyAsync(h_row, c->run_row.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
            CT_CUDA(cudaMemcpyAsync(h_val, c->run_val.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st));
        }
        for (auto& t : zero_threads) t.join();
        zero_threads.clear();
        CT_CUDA(cudaStreamSynchronize(st));
        const int np = cti::api_host_expand_runs(c, h_x, h_row, h_val, R, (long)hp * H, W, flag_host, nranks);
        for (const ctb::Override& o : ovr) {                           // pieces of components split at a stale box
            int32_t* out = flag_host + ((size_t)o.t * H + o.y) * W;
            for (int xx = o.x0; xx < o.x1; ++xx) out[xx] = o.val;
        }
        c->stats["host_threads"] = (double)np;
        c->stats["h2d_bytes"] = (double)((size_t)T_local * plane_bytes);
        c->stats["d2h_bytes"] = (double)((size_t)R * 12);
    }
    CT_CUDA(cudaEventRecord(c->ev[4], st));
    CT_CUDA(cudaStreamSynchronize(st));
    CT_CUDA(cudaStreamSynchronize(aux));
    float ms = 0;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); c->stats["ms_threshold"] = ms;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev[3], c->ev[4])); c->stats["ms_paint"] = ms;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev[0], c->ev[4])); c->stats["ms_total"] = ms;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev_side[0], c->ev_side[1])); c->stats["ms_zero_fill"] = ms;
    CT_CUDA(cudaEventElapsedTime(&ms, c->ev[1], c->ev_tbl[0])); c->stats["ms_tables_after_threshold"] = ms;
    if (c->plane_timed && cudaEventElapsedTime(&ms, c->ev_p[0], c->ev_p[1]) == cudaSuccess) c->stats["ms_plane_kernel"] = ms;
    if (c->plane_timed && cudaEventElapsedTime(&ms, c->ev_p[1], g->ev_p[2]) == cudaSuccess) c->stats["ms_exchange"] = ms;   // pack + all-gather + merge
    c->plane_timed = 0;
    c->stats["kernel_launches"] = (double)(c->launches + g->launches);
    c->stats["runs"] = (double)c->nruns; c->stats["comps2d"] = (double)c->ncomp;
    c->stats["exchange_bytes"] = (double)((HDR_BYTES + [&] { size_t off[cts::A_COUNT]; return cts::layout(comm_h->capC, comm_h->capP, comm_h->capS, off); }() + 255) / 256 * 256);
    for (const char* k : {"labels3d", "features", "seam_events", "seam_splits", "neartie_resolved", "neartie_flagged", "sweeps",
                          "wavefront_planes", "ms_host_tables", "ms_g_kernel", "label_fast", "event_segments", "ms_global_kernel"})
        if (g->stats.count(k)) c->stats[k] = g->stats[k];
    c->stats["fast_path"] = outcome == ctf::FAST_OK ? (classic ? 0.5 : 1.0) : 0.25;
    return CT_OK;
}

extern "C" {

int ct_run_contrack_sharded(ct_ctx* c, ct_comm* comm, const void* anom_dev, int in_dtype, long T_local, long t_begin,
                            long T_total, int H, int W, const double* w_host, const double* thr_host, long thr_n,
                            int thr_is_f32, int op, double overlap, int persistence, int twosided, int32_t* flag_dev,
                            long* n_features, void* stream) {
    if (!anom_dev || !flag_dev) return fail(CT_ERR_ARG, "null device pointer");
    return sharded_run(c, comm, anom_dev, nullptr, in_dtype, T_local, t_begin, T_total, H, W, w_host, thr_host, thr_n, thr_is_f32,
                       op, overlap, persistence, twosided, flag_dev, nullptr, 0, n_features, stream);
}

int ct_run_contrack_sharded_host(ct_ctx* c, ct_comm* comm, const void* anom_host, int in_dtype, long T_local, long t_begin,
                                 long T_total, int H, int W, const double* w_host, const double* thr_host, long thr_n,
                                 int thr_is_f32, int op, double overlap, int persistence, int twosided, int32_t* flag_host,
                                 long* n_features, long chunk_planes) {
    if (!anom_host || !flag_host) return fail(CT_ERR_ARG, "null host pointer");
    if (!c) return fail(CT_ERR_ARG, "null context");
    CT_CUDA(cudaSetDevice(c->device));
    if (!c->host_stream) CT_CUDA(cudaStreamCreateWithFlags(&c->host_stream, cudaStreamNonBlocking));
    return sharded_run(c, comm, nullptr, anom_host, in_dtype, T_local, t_begin, T_total, H, W, w_host, thr_host, thr_n, thr_is_f32,
                       op, overlap, persistence, twosided, nullptr, flag_host, chunk_planes, n_features, c->host_stream);
}

}  // extern "C"
