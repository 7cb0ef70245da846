// ct_kernels.cu -- sm_100a kernels of the run_contrack path ("label once, then tables").
//
//   threshold_bits   contrack.py:648-674   anomaly cube -> 1 bit per cell, row-run counts, date-line flags and the first 8
//                                          row-runs of every row (x0 | x1 << 16) in the row's slots
//   compact_runs                           row slots -> row-runs (y, x0, x1) in raster order (extract_runs: from bit rows,
//                                          for rows with more runs than slots)
//   ccl_*            contrack.py:684-687   8-connected components of each plane = union-find over row-runs
//   seam_rows/cls_*  contrack.py:691-698   same-row date-line classes on top of the components
//   comp_*/pairs_*   contrack.py:717-719   area tables: per component, and per (component at t, component at t-1)
//   step3_* / link_* / label_*  contrack.py:706-751  overlap filter (Jacobi sweeps), 3-D labels, label boxes, on the tables
//   paint            contrack.py:776-791   per-run value -> int32 flag cube (zero fill + cells of runs, or dense)
//
// The two cube-sized kernels (threshold_bits: 4 B/cell read; paint: 4 B/cell write) are HBM-bound streaming kernels;
// everything between them touches only bit rows (1/32 B/cell... 4/32 B per cell) and run/component tables.
#include "ct_kernels.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>

namespace ctk {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int RUN_SLOTS = ctk::RUN_SLOTS_PER_ROW;

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

// ---------------------------------------------------------------------------------------------------------------
// row-runs straight from the mask words a warp holds (lane j = word k of the row, `left` = the word before it): the i-th
// run of the row goes to the row's slot i as two uint16 halves (x0 | x1 << 16, x1 exclusive).  A run ends where a 0 follows
// a 1, so only the LEFT neighbour is needed and the ends come out in the order of the starts.  Rows with more than
// RUN_SLOTS runs keep only the first RUN_SLOTS here; the table phase re-extracts those rows from the bit rows.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_runs(uint32_t w, uint32_t left, int k, int lane, uint32_t& sbase, uint32_t& ebase,
                                          uint16_t* __restrict__ slot16) {
    const uint32_t edge = (w << 1) | (left >> 31);
    uint32_t starts = w & ~edge, falls = ~w & edge;
    if (!__any_sync(FULL, (starts | falls) != 0u)) return;
    uint32_t ts, te;
    uint32_t is = sbase + warp_excl_scan(__popc(starts), lane, &ts);
    uint32_t ie = ebase + warp_excl_scan(__popc(falls), lane, &te);
    while (starts) {
        const int bit = __ffs(starts) - 1;
        starts &= starts - 1;
        if (is < (uint32_t)RUN_SLOTS) slot16[2 * is] = (uint16_t)(k * 32 + bit);
        ++is;
    }
    while (falls) {
        const int bit = __ffs(falls) - 1;
        falls &= falls - 1;
        if (ie < (uint32_t)RUN_SLOTS) slot16[2 * ie + 1] = (uint16_t)(k * 32 + bit);
        ++ie;
    }
    sbase += ts; ebase += te;
}
// a run that reaches the end of a row whose last word fills lane 31 of the last group has no word behind it: close it
__device__ __forceinline__ void emit_close(uint32_t carry_word, int W, int lane, uint32_t ebase, uint16_t* __restrict__ slot16) {
    if ((carry_word >> 31) && lane == 0 && ebase < (uint32_t)RUN_SLOTS) slot16[2 * ebase + 1] = (uint16_t)W;
}

// ---------------------------------------------------------------------------------------------------------------
// threshold -> bits
// ---------------------------------------------------------------------------------------------------------------
template <typename TIn, bool F32CMP, int OP>
__device__ __forceinline__ bool cmp_thr(TIn v, float thr_f, double thr_d) {
    if (F32CMP) {
        const float x = (float)v;
        if (OP == 0) return x >= thr_f;
        if (OP == 1) return x <= thr_f;
        if (OP == 2) return x > thr_f;
        return x < thr_f;
    } else {
        const double x = (double)v;
        if (OP == 0) return x >= thr_d;
        if (OP == 1) return x <= thr_d;
        if (OP == 2) return x > thr_d;
        return x < thr_d;
    }
}

__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }

// One warp per row.  A batch is 8 independent 128-byte warp loads (4 B per lane, fully coalesced); one ballot packs the
// 32 compared cells into a mask word that lane (k mod 32) keeps.  Run starts, the row's run count and the two date-line
// bits are derived once per 32 words from the kept words (lane-parallel), not per word: the inner loop is
// LDG + FSETP + VOTE + SEL per 32 cells, so the kernel stays bound by HBM rather than by instruction issue.
template <typename TIn, bool F32CMP, int OP, int NB>
__global__ void __launch_bounds__(256) k_threshold(const TIn* __restrict__ anom, long nrows, int H, int W, int Ww,
                                                   const double* __restrict__ thr, long thr_n,
                                                   uint32_t* __restrict__ bits, uint32_t* __restrict__ row_cnt,
                                                   uint32_t* __restrict__ seam_flag, uint32_t* __restrict__ slots,
                                                   uint32_t* __restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int last_word = (W - 1) >> 5, last_bit = (W - 1) & 31;
    const int nfull = W >> 5;                                       // words whose 32 cells are all inside the row
    for (long row = warp0; row < nrows; row += nwarps) {
        const double thr_d = thr[thr_n == 1 ? 0 : row / H];
        const float thr_f = (float)thr_d;
        const TIn* a = anom + row * (long)W;
        uint32_t cnt = 0, carry_word = 0, first = 0, last = 0, sb = 0, eb = 0;
        uint16_t* slot16 = reinterpret_cast<uint16_t*>(slots + row * (long)RUN_SLOTS);
        for (int k0 = 0; k0 < Ww; k0 += 32) {
            uint32_t myword = 0;
            const int kend = min(32, Ww - k0);
            for (int j0 = 0; j0 < kend; j0 += NB) {
                TIn v[NB];
                if (k0 + j0 + NB <= nfull) {
#pragma unroll
                    for (int j = 0; j < NB; ++j) v[j] = ld_stream(a + (k0 + j0 + j) * 32 + lane);
                } else {
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        const int x = (k0 + j0 + j) * 32 + lane;
                        v[j] = (x < W) ? ld_stream(a + x) : (TIn)NAN;
                    }
                }
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const uint32_t m = __ballot_sync(FULL, cmp_thr<TIn, F32CMP, OP>(v[j], thr_f, thr_d));
                    if (lane == j0 + j) myword = m;
                }
            }
            if (lane < kend) bits[row * (long)Ww + k0 + lane] = myword;
            uint32_t prev = __shfl_up_sync(FULL, myword, 1);
            if (lane == 0) prev = carry_word;
            cnt += __popc(myword & ~((myword << 1) | (prev >> 31)));
            emit_runs(myword, prev, k0 + lane, lane, sb, eb, slot16);
            carry_word = __shfl_sync(FULL, myword, 31);
            if (k0 == 0) first = __shfl_sync(FULL, myword, 0) & 1u;
            if (last_word >= k0 && last_word < k0 + 32) last = (__shfl_sync(FULL, myword, last_word - k0) >> last_bit) & 1u;
        }
        emit_close(carry_word, W, lane, eb, slot16);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
        if (lane == 0) {
            row_cnt[row] = cnt; seam_flag[row] = first & last;
            if (cnt > (uint32_t)RUN_SLOTS) *overflow = 1u;
        }
    }
}

// ---- variant 1: rows staged through shared memory by the bulk-copy engine (cp.async.bulk, SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the cube is read exactly once: mark its lines evict-first so that the (small, re-used) tables of the table kernels that
// run beside this kernel stay in L2
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// Each warp runs its own ring of NS row buffers: lane 0 arms the stage's mbarrier and issues one bulk copy per row
// (W * sizeof(TIn) bytes, a multiple of 16); the warp waits on the barrier's phase, compares the row out of shared memory
// (lane-contiguous 4-byte reads, conflict free), and re-issues the buffer for the row NS steps ahead.  No load
// instruction touches global memory, no register staging: NS * 8 rows per SM are in flight.
template <typename TIn, bool F32CMP, int OP>
__global__ void __launch_bounds__(512) k_threshold_bulk(const TIn* __restrict__ anom, long nrows, int H, int W, int Ww,
                                                        const double* __restrict__ thr, long thr_n,
                                                        uint32_t* __restrict__ bits, uint32_t* __restrict__ row_cnt,
                                                        uint32_t* __restrict__ seam_flag, uint32_t* __restrict__ slots,
                                                        uint32_t* __restrict__ overflow, int NS, int stage_bytes, int bits_mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + wid * NS;                   // [nw][NS]
    unsigned char* stage0 = smem + ((nw * NS * 8 + 127) / 128) * 128 + (size_t)wid * NS * stage_bytes;
    const long warp0 = (long)blockIdx.x * nw + wid;
    const long nwarps = (long)gridDim.x * nw;
    const uint32_t row_bytes = (uint32_t)(W * sizeof(TIn));
    const int last_word = (W - 1) >> 5, last_bit = (W - 1) & 31;
    const int nfull = W >> 5;
    const uint64_t pol = l2_evict_first_policy();
    if (lane == 0) {
        for (int s2 = 0; s2 < NS; ++s2) mbar_init(&bars[s2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        for (int s2 = 0; s2 < NS; ++s2) {
            const long r = warp0 + (long)s2 * nwarps;
            if (r < nrows) {
                mbar_expect_tx(&bars[s2], row_bytes);
                bulk_g2s_hint(stage0 + (size_t)s2 * stage_bytes, anom + r * (long)W, row_bytes, &bars[s2], pol);
            }
        }
    }
    int st = 0;
    uint32_t phase = 0;
    const double thr_one = thr[0];                                   // the usual case: one threshold for the whole cube
    for (long row = warp0; row < nrows; row += nwarps) {
        const double thr_d = thr_n == 1 ? thr_one : thr[row / H];
        const float thr_f = (float)thr_d;
        mbar_wait(&bars[st], phase);
        const TIn* a = reinterpret_cast<const TIn*>(stage0 + (size_t)st * stage_bytes);
        uint32_t carry_word = 0, first = 0, last = 0, sb = 0, eb = 0;
        uint16_t* slot16 = reinterpret_cast<uint16_t*>(slots + row * (long)RUN_SLOTS);
        uint32_t* brow = bits + row * (long)Ww;
        uint32_t kept0 = 0, kept1 = 0;                               // bits_mode 1: the row's words wait for the run count
        for (int k0 = 0; k0 < Ww; k0 += 32) {
            uint32_t myword = 0;
            const int kend = min(32, Ww - k0);
            for (int j0 = 0; j0 < kend; j0 += 8) {
                TIn v[8];
                if (k0 + j0 + 8 <= nfull) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = a[(k0 + j0 + j) * 32 + lane];
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int x = (k0 + j0 + j) * 32 + lane;
                        v[j] = (x < W) ? a[x] : (TIn)NAN;
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t m = __ballot_sync(FULL, cmp_thr<TIn, F32CMP, OP>(v[j], thr_f, thr_d));
                    if (lane == j0 + j) myword = m;
                }
            }
            if (bits_mode == 0) { if (lane < kend) brow[k0 + lane] = myword; }
            else if (k0 == 0) kept0 = myword;
            else kept1 = myword;
            uint32_t prev = __shfl_up_sync(FULL, myword, 1);
            if (lane == 0) prev = carry_word;
            emit_runs(myword, prev, k0 + lane, lane, sb, eb, slot16);   // sb: run starts of the row so far (warp-uniform)
            carry_word = __shfl_sync(FULL, myword, 31);
            if (k0 == 0) first = __shfl_sync(FULL, myword, 0) & 1u;
            if (last_word >= k0 && last_word < k0 + 32) last = (__shfl_sync(FULL, myword, last_word - k0) >> last_bit) & 1u;
        }
        // every lane has read its part of the buffer (the ballots above are warp-synchronous): hand it back
        __syncwarp();
        if (lane == 0) {
            const long r = row + (long)NS * nwarps;
            if (r < nrows) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&bars[st], row_bytes);
                bulk_g2s_hint(stage0 + (size_t)st * stage_bytes, anom + r * (long)W, row_bytes, &bars[st], pol);
            }
        }
        emit_close(carry_word, W, lane, eb, slot16);
        if (bits_mode != 0 && sb > (uint32_t)RUN_SLOTS) {            // only rows the slots cannot hold need their bit row
            if (lane < Ww) brow[lane] = kept0;
            if (32 + lane < Ww) brow[32 + lane] = kept1;
        }
        if (lane == 0) {
            row_cnt[row] = sb; seam_flag[row] = first & last;
            if (sb > (uint32_t)RUN_SLOTS) *overflow = 1u;
        }
        if (++st == NS) { st = 0; phase ^= 1u; }
    }
}

template <typename TIn, bool F32CMP>
cudaError_t launch_threshold_bulk(const ThresholdArgs& a, int sm_count, cudaStream_t st) {
    const long nrows = a.T * a.H;
    const int stage_bytes = (int)(((size_t)a.W * sizeof(TIn) + 127) / 128 * 128);
    // 16 warps per CTA when two row buffers per warp still fit (more warps to hide the LDS / ballot latency), else 8
    int nw = (16 * 2 * stage_bytes <= 199 * 1024) ? 16 : 8;
    int NS = (200 * 1024 - 1024) / (nw * stage_bytes);
    if (NS > 8) NS = 8;
    const size_t smem = ((size_t)nw * NS * 8 + 127) / 128 * 128 + (size_t)nw * NS * stage_bytes;
    int blocks = (int)std::min<long>((nrows + nw - 1) / nw, (long)sm_count);
#define CT_LAUNCH_THRB(OPV)                                                                                         \
    do {                                                                                                            \
        cudaError_t e = cudaFuncSetAttribute(k_threshold_bulk<TIn, F32CMP, OPV>,                                    \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
        if (e != cudaSuccess) return e;                                                                             \
        k_threshold_bulk<TIn, F32CMP, OPV><<<blocks, nw * 32, smem, st>>>((const TIn*)a.anom, nrows, a.H, a.W, a.Ww, \
                                                                          a.thr_dev, a.thr_n, a.bits, a.row_cnt,    \
                                                                          a.seam_flag, a.slots, a.overflow, NS,     \
                                                                          stage_bytes,                              \
                                                                          (a.bits_overflow_only && a.Ww <= 64) ? 1 : 0); \
    } while (0)
    switch (a.op) {
        case 0: CT_LAUNCH_THRB(0); break;
        case 1: CT_LAUNCH_THRB(1); break;
        case 2: CT_LAUNCH_THRB(2); break;
        case 3: CT_LAUNCH_THRB(3); break;
        default: return cudaErrorInvalidValue;
    }
#undef CT_LAUNCH_THRB
    return cudaGetLastError();
}

// bulk staging needs 16-byte aligned rows and at least two row buffers per warp in shared memory
template <typename TIn>
bool bulk_ok(const ThresholdArgs& a) {
    const size_t rb = (size_t)a.W * sizeof(TIn);
    const size_t stage = (rb + 127) / 128 * 128;
    return (rb % 16 == 0) && ((reinterpret_cast<uintptr_t>(a.anom) & 15) == 0) && (8 * 2 * stage <= 199 * 1024);
}

template <typename TIn, bool F32CMP>
cudaError_t launch_threshold(const ThresholdArgs& a, int blocks, cudaStream_t st) {
    const long nrows = a.T * a.H;
#define CT_LAUNCH_THR(OPV)                                                                                          \
    k_threshold<TIn, F32CMP, OPV, 16><<<blocks, 256, 0, st>>>((const TIn*)a.anom, nrows, a.H, a.W, a.Ww, a.thr_dev,  \
                                                              a.thr_n, a.bits, a.row_cnt, a.seam_flag, a.slots,     \
                                                              a.overflow)
    switch (a.op) {
        case 0: CT_LAUNCH_THR(0); break;
        case 1: CT_LAUNCH_THR(1); break;
        case 2: CT_LAUNCH_THR(2); break;
        case 3: CT_LAUNCH_THR(3); break;
        default: return cudaErrorInvalidValue;
    }
#undef CT_LAUNCH_THR
    return cudaGetLastError();
}

// run count and date-line flag of bit rows that arrived from elsewhere (the halo plane of a time-sharded run)
__global__ void __launch_bounds__(256) k_row_stats(const uint32_t* __restrict__ bits, long nrows, int W, int Ww,
                                                   uint32_t* __restrict__ row_cnt, uint32_t* __restrict__ seam_flag,
                                                   uint32_t* __restrict__ slots, uint32_t* __restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const uint32_t* b = bits + row * (long)Ww;
    uint32_t cnt = 0, carry_word = 0, sb = 0, eb = 0;
    uint16_t* slot16 = reinterpret_cast<uint16_t*>(slots + row * (long)RUN_SLOTS);
    for (int k0 = 0; k0 < Ww; k0 += 32) {
        const uint32_t m = (k0 + lane < Ww) ? b[k0 + lane] : 0u;
        uint32_t prev = __shfl_up_sync(FULL, m, 1);
        if (lane == 0) prev = carry_word;
        cnt += __popc(m & ~((m << 1) | (prev >> 31)));
        emit_runs(m, prev, k0 + lane, lane, sb, eb, slot16);
        carry_word = __shfl_sync(FULL, m, 31);
    }
    emit_close(carry_word, W, lane, eb, slot16);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
    if (lane == 0) {
        row_cnt[row] = cnt;
        seam_flag[row] = (b[0] & 1u) & ((b[(W - 1) >> 5] >> ((W - 1) & 31)) & 1u);
        if (cnt > (uint32_t)RUN_SLOTS) *overflow = 1u;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// exclusive scan (three small kernels; inputs are row/run-sized tables, a few tens of MB at most)
// ---------------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t wtot;
    uint32_t ex = warp_excl_scan(v, lane, &wtot);
    if (lane == 31) s_warp[wid] = wtot;
    __syncthreads();
    if (wid == 0) {
        uint32_t x = lane < nw ? s_warp[lane] : 0, t;
        uint32_t e = warp_excl_scan(x, lane, &t);
        if (lane < nw) s_warp[lane] = e;
        if (lane == 0) s_warp[32] = t;
    }
    __syncthreads();
    ex += s_warp[wid];
    *total = s_warp[32];
    __syncthreads();
    return ex;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, long n,
                                                              uint32_t* __restrict__ sums) {
    __shared__ uint32_t s_warp[33];
    const long base = (long)blockIdx.x * SCAN_TILE + (long)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) s += in[base + i];
    uint32_t tot;
    block_excl_scan(s, s_warp, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// wide_flag (may be null): set to 1 when base + total does not fit in 32 bits (the scanned values have wrapped)
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* sums, long nb, uint32_t* total_out, uint32_t base,
                                                    uint32_t* wide_flag) {
    __shared__ uint32_t s_warp[33];
    uint32_t running = base;
    unsigned long long running64 = base;
    for (long b0 = 0; b0 < nb; b0 += 1024) {
        const long i = b0 + threadIdx.x;
        uint32_t v = i < nb ? sums[i] : 0, tot;
        uint32_t ex = block_excl_scan(v, s_warp, &tot);
        if (i < nb) sums[i] = running + ex;
        running += tot;
        running64 += tot;
    }
    if (threadIdx.x == 0) {
        *total_out = running;
        if (wide_flag && (running64 >> 32)) *wide_flag = 1u;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                             long n, const uint32_t* __restrict__ sums) {
    __shared__ uint32_t s_warp[33];
    const long base = (long)blockIdx.x * SCAN_TILE + (long)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
    uint32_t tot;
    uint32_t ex = block_excl_scan(s, s_warp, &tot) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
}

// ---------------------------------------------------------------------------------------------------------------
// bit rows -> row-runs
// ---------------------------------------------------------------------------------------------------------------
// rows with at most RUN_SLOTS runs: the threshold kernel left their runs in the row's slots -> dense run tables
__global__ void __launch_bounds__(256) k_compact_runs(const uint4* __restrict__ slots, const uint32_t* __restrict__ row_ptr,
                                                      long nrows, uint32_t* __restrict__ run_x,
                                                      uint32_t* __restrict__ run_row, long row0,
                                                      uint32_t* __restrict__ ovf_rows, uint32_t* __restrict__ ovf_count) {
    static_assert(RUN_SLOTS == 8, "two uint4 per row");
    const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const uint32_t base = row_ptr[row], n = row_ptr[row + 1] - base;
    if (n == 0) return;
    if (n > (uint32_t)RUN_SLOTS) {                     // the slots hold only the first runs: list the row for re-extraction
        if (ovf_rows) ovf_rows[atomicAdd(ovf_count, 1u)] = (uint32_t)row;
        return;
    }
    const uint4 a = slots[2 * row];
    const uint32_t v0[4] = {a.x, a.y, a.z, a.w};
    const uint32_t r = (uint32_t)(row0 + row);
#pragma unroll
    for (int i = 0; i < 4; ++i) if ((uint32_t)i < n) { run_x[base + i] = v0[i]; run_row[base + i] = r; }
    if (n > 4) {
        const uint4 b = slots[2 * row + 1];
        const uint32_t v1[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) if ((uint32_t)(4 + i) < n) { run_x[base + 4 + i] = v1[i]; run_row[base + 4 + i] = r; }
    }
}

// rows listed in ovf_rows[0 .. *ovf_count): the ones with more runs than the threshold kernel's slots hold
__global__ void __launch_bounds__(256) k_extract_runs(const uint32_t* __restrict__ bits,
                                                      const uint32_t* __restrict__ row_ptr, long nrows, int Ww,
                                                      uint32_t* __restrict__ run_x, uint32_t* __restrict__ run_row,
                                                      long row0, const uint32_t* __restrict__ ovf_rows,
                                                      const uint32_t* __restrict__ ovf_count) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    uint16_t* rx = reinterpret_cast<uint16_t*>(run_x);
    const long nwork = (long)*ovf_count;
    for (long wi = warp0; wi < nwork; wi += nwarps) {
        const long row = (long)ovf_rows[wi];
        // everything a 64-word row needs is requested before anything is consumed: one memory latency per row
        const uint32_t* b = bits + row * (long)Ww;
        const uint32_t base = row_ptr[row], next = row_ptr[row + 1];
        uint32_t m = lane < Ww ? b[lane] : 0u;
        uint32_t m_ahead = (32 + lane < Ww) ? b[32 + lane] : 0u;
        if (next == base) continue;
        uint32_t sbase = base, ebase = base, prev = 0;
        for (int k0 = 0; k0 < Ww; k0 += 32) {
            const int k = k0 + lane;
            uint32_t left = __shfl_up_sync(FULL, m, 1);
            if (lane == 0) left = prev;
            uint32_t right = __shfl_down_sync(FULL, m, 1);
            const uint32_t ahead0 = __shfl_sync(FULL, m_ahead, 0);
            if (lane == 31) right = ahead0;
            uint32_t starts = m & ~((m << 1) | (left >> 31));
            uint32_t ends = m & ~((m >> 1) | (right << 31));
            uint32_t ts, te;
            uint32_t is = sbase + warp_excl_scan(__popc(starts), lane, &ts);
            uint32_t ie = ebase + warp_excl_scan(__popc(ends), lane, &te);
            while (starts) {
                const int bit = __ffs(starts) - 1;
                starts &= starts - 1;
                rx[2 * (size_t)is] = (uint16_t)(k * 32 + bit);
                run_row[is] = (uint32_t)(row0 + row);
                ++is;
            }
            while (ends) {
                const int bit = __ffs(ends) - 1;
                ends &= ends - 1;
                rx[2 * (size_t)ie + 1] = (uint16_t)(k * 32 + bit + 1);
                ++ie;
            }
            sbase += ts; ebase += te;
            prev = __shfl_sync(FULL, m, 31);
            m = m_ahead;
            m_ahead = (k0 + 64 + lane < Ww) ? b[k0 + 64 + lane] : 0u;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// union-find helpers (lock-free, smallest index is the root)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t uf_find(const uint32_t* parent, uint32_t x) {
    const volatile uint32_t* p = parent;
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
        if (a < b) { const uint32_t t = a; a = b; b = t; }      // a > b: hang a below b
        const uint32_t old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;                                                 // somebody re-parented a meanwhile: join that with b
    }
}

__global__ void k_iota(uint32_t* p, long begin, long end) {
    const long i = begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < end) p[i] = (uint32_t)i;
}

// 8-connectivity between a run and the runs of the row above: [x0-1, x1+1) must meet [px0, px1).
__global__ void __launch_bounds__(256) k_ccl_union(const uint32_t* __restrict__ row_ptr,
                                                   const uint32_t* __restrict__ run_x,
                                                   const uint32_t* __restrict__ run_row, long begin, long rend,
                                                   int H, uint32_t* parent) {
    const long r = begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rend) return;
    const uint32_t row = run_row[r];
    if (row % (uint32_t)H == 0) return;
    const uint32_t x = run_x[r];
    const int x0 = x & 0xffff, x1 = x >> 16;
    uint32_t lo = row_ptr[row - 1], hi = row_ptr[row];
    const uint32_t end = hi;
    while (lo < hi) {                                            // first run above with px1 >= x0
        const uint32_t mid = (lo + hi) >> 1;
        if ((int)(run_x[mid] >> 16) >= x0) hi = mid; else lo = mid + 1;
    }
    for (uint32_t p = lo; p < end; ++p) {
        if ((int)(run_x[p] & 0xffff) > x1) break;
        uf_union(parent, (uint32_t)r, p);
    }
}

__global__ void __launch_bounds__(256) k_ccl_flatten(uint32_t* parent, uint32_t* __restrict__ root_flag, long begin,
                                                     long end) {
    const long r = begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= end) return;
    const uint32_t root = uf_find(parent, (uint32_t)r);
    root_flag[r] = root == (uint32_t)r ? 1u : 0u;
    if (root != (uint32_t)r) parent[r] = root;
}

__global__ void __launch_bounds__(256) k_ccl_assign(const uint32_t* __restrict__ parent,
                                                    const uint32_t* __restrict__ rank, uint32_t* __restrict__ run_comp,
                                                    long begin, long end) {
    const long r = begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= end) return;
    // after k_ccl_flatten parent[r] is the root or one hop from it (a concurrent flatten may have left a short chain)
    run_comp[r] = rank[uf_find(parent, (uint32_t)r)];
}

// ---------------------------------------------------------------------------------------------------------------
// component tables
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_comp_init(CompTables c, long begin, long end, int W) {
    const long i = begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    c.t[i] = 0; c.y0[i] = INT_MAX; c.y1[i] = 0; c.x0[i] = W; c.x1[i] = 0;
    c.areaE[i] = 0.0; c.areaS[i] = 0.0; c.nsp[i] = 0; c.cls[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_comp_accumulate(const uint32_t* __restrict__ run_x,
                                                         const uint32_t* __restrict__ run_row,
                                                         const uint32_t* __restrict__ run_comp, long begin, long end,
                                                         int H, const double* __restrict__ w,
                                                         const uint8_t* __restrict__ special, CompTables c) {
    const long r = begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= end) return;
    const uint32_t row = run_row[r], comp = run_comp[r];
    const int t = (int)(row / (uint32_t)H), y = (int)(row % (uint32_t)H);
    const uint32_t x = run_x[r];
    const int x0 = x & 0xffff, x1 = x >> 16;
    c.t[comp] = t;
    atomicMin(&c.y0[comp], y);
    atomicMax(&c.y1[comp], y + 1);
    atomicMin(&c.x0[comp], x0);
    atomicMax(&c.x1[comp], x1);
    const double area = (double)(x1 - x0) * w[y];                // exact: < 2^16 times a float32-valued double
    if (special[y]) { atomicAdd(&c.areaS[comp], area); atomicAdd(&c.nsp[comp], (uint32_t)(x1 - x0)); }
    else atomicAdd(&c.areaE[comp], area);
}

__global__ void __launch_bounds__(256) k_seam_rows(const uint32_t* __restrict__ seam_flag,
                                                   const uint32_t* __restrict__ seam_pos,
                                                   const uint32_t* __restrict__ row_ptr,
                                                   const uint32_t* __restrict__ run_comp, long row_begin,
                                                   long row_end, uint32_t* __restrict__ seam_row,
                                                   uint32_t* __restrict__ seam_a, uint32_t* __restrict__ seam_b,
                                                   uint32_t* cls_parent) {
    const long row = row_begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= row_end || !seam_flag[row]) return;
    const uint32_t a = run_comp[row_ptr[row]], b = run_comp[row_ptr[row + 1] - 1];
    const uint32_t pos = seam_pos[row];
    seam_row[pos] = (uint32_t)row; seam_a[pos] = a; seam_b[pos] = b;
    if (a != b) uf_union(cls_parent, a, b);
}

__global__ void __launch_bounds__(256) k_cls_flatten(uint32_t* cls_parent, long begin, long end) {
    const long c = begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= end) return;
    const uint32_t root = uf_find(cls_parent, (uint32_t)c);
    if (root != (uint32_t)c) cls_parent[c] = root;
}

// ---------------------------------------------------------------------------------------------------------------
// pair table
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pairs_init(PairTable p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.cap) return;
    p.key[i] = PAIR_EMPTY; p.npix[i] = 0; p.nsp[i] = 0; p.areaE[i] = 0.0; p.areaS[i] = 0.0;
    if (i == 0) { *p.overflow = 0; *p.count = 0; }
}

__device__ __forceinline__ uint32_t hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (uint32_t)k;
}

__global__ void __launch_bounds__(256) k_pairs_accumulate(const uint32_t* __restrict__ row_ptr,
                                                          const uint32_t* __restrict__ run_x,
                                                          const uint32_t* __restrict__ run_row,
                                                          const uint32_t* __restrict__ run_comp, long begin, long rend,
                                                          int H, const double* __restrict__ w,
                                                          const uint8_t* __restrict__ special, PairTable pt) {
    const long r = begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rend) return;
    const uint32_t row = run_row[r];
    if (row < (uint32_t)H) return;                               // plane 0 has no predecessor
    const uint32_t prow = row - (uint32_t)H;
    const uint32_t x = run_x[r];
    const int x0 = x & 0xffff, x1 = x >> 16;
    uint32_t lo = row_ptr[prow], hi = row_ptr[prow + 1];
    const uint32_t end = hi;
    while (lo < hi) {                                            // first run of the earlier plane with px1 > x0
        const uint32_t mid = (lo + hi) >> 1;
        if ((int)(run_x[mid] >> 16) > x0) hi = mid; else lo = mid + 1;
    }
    const int y = (int)(row % (uint32_t)H);
    const double wy = w[y];
    const bool sp = special[y] != 0;
    const unsigned long long ka = (unsigned long long)run_comp[r] << 32;
    const uint32_t mask = pt.cap - 1;
    for (uint32_t p = lo; p < end; ++p) {
        const uint32_t px = run_x[p];
        const int px0 = px & 0xffff, px1 = px >> 16;
        if (px0 >= x1) break;
        const int n = min(x1, px1) - max(x0, px0);
        const unsigned long long key = ka | run_comp[p];
        uint32_t slot = hash64(key) & mask;
        bool found = false;
        for (uint32_t probe = 0; probe <= mask; ++probe) {
            unsigned long long k = *((volatile unsigned long long*)&pt.key[slot]);
            if (k == PAIR_EMPTY) {
                k = atomicCAS(&pt.key[slot], PAIR_EMPTY, key);
                if (k == PAIR_EMPTY) atomicAdd(pt.count, 1u);       // a new pair
            }
            if (k == PAIR_EMPTY || k == key) { found = true; break; }
            slot = (slot + 1) & mask;
        }
        if (!found) { *pt.overflow = 1; return; }
        atomicAdd(&pt.npix[slot], (uint32_t)n);
        const double area = (double)n * wy;
        if (sp) { atomicAdd(&pt.areaS[slot], area); atomicAdd(&pt.nsp[slot], (uint32_t)n); }
        else atomicAdd(&pt.areaE[slot], area);
    }
}

__global__ void __launch_bounds__(256) k_pairs_compact(PairTable p, uint32_t* __restrict__ out_a,
                                                       uint32_t* __restrict__ out_b, uint32_t* __restrict__ out_npix,
                                                       uint32_t* __restrict__ out_nsp, double* __restrict__ out_E,
                                                       double* __restrict__ out_S, uint32_t out_cap, uint32_t* count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.cap) return;
    const unsigned long long k = p.key[i];
    if (k == PAIR_EMPTY) return;
    const uint32_t idx = atomicAdd(count, 1u);
    if (idx >= out_cap) return;
    out_a[idx] = (uint32_t)(k >> 32); out_b[idx] = (uint32_t)k;
    out_npix[idx] = p.npix[i]; out_nsp[idx] = p.nsp[i]; out_E[idx] = p.areaE[i]; out_S[idx] = p.areaS[i];
}

// ---------------------------------------------------------------------------------------------------------------
// tables in the layout the ordered host phase reads sequentially (ct_host.h: FastTables)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_class_sums(CompTables c, ClassTables k, long begin, long end) {
    const long i = begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const uint32_t rep = c.cls[i];
    atomicAdd(&k.conE[rep], c.areaE[i]);
    if (c.nsp[i]) { atomicAdd(&k.conS[rep], c.areaS[i]); atomicAdd(&k.nsp[rep], c.nsp[i]); }
}

// per occupied slot: count the pair for its plane-t component, add its area to the forward sums of the class of its
// plane-(t-1) component
__global__ void __launch_bounds__(256) k_pairs_count(PairTable p, const uint32_t* __restrict__ cls, ClassTables k,
                                                     uint32_t* __restrict__ pcnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.cap) return;
    const unsigned long long key = p.key[i];
    if (key == PAIR_EMPTY) return;
    atomicAdd(&pcnt[(uint32_t)(key >> 32)], 1u);
    const uint32_t rep = cls[(uint32_t)key];
    atomicAdd(&k.fE[rep], p.areaE[i]);
    if (p.nsp[i]) { atomicAdd(&k.fS[rep], p.areaS[i]); atomicAdd(&k.fnsp[rep], p.nsp[i]); }
}

__global__ void __launch_bounds__(256) k_pairs_fill(PairTable p, const uint32_t* __restrict__ pptr,
                                                    uint32_t* __restrict__ pfill, PairCsr o) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.cap) return;
    const unsigned long long key = p.key[i];
    if (key == PAIR_EMPTY) return;
    const uint32_t a = (uint32_t)(key >> 32);
    const uint32_t pos = pptr[a] + atomicAdd(&pfill[a], 1u);
    o.b[pos] = (uint32_t)key; o.npix[pos] = p.npix[i]; o.nsp[pos] = p.nsp[i]; o.E[pos] = p.areaE[i]; o.S[pos] = p.areaS[i];
}

// date-line rows -> segments of consecutive rows of one plane with the same two components
__global__ void __launch_bounds__(256) k_seg_flags(const uint32_t* __restrict__ srow, const uint32_t* __restrict__ sa,
                                                   const uint32_t* __restrict__ sb, long begin, long end, int H,
                                                   uint32_t* __restrict__ start) {
    const long i = begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    bool st = i == begin;                                          // a range starts at a plane boundary
    if (!st) st = srow[i] != srow[i - 1] + 1 || (srow[i] % (uint32_t)H) == 0 || sa[i] != sa[i - 1] || sb[i] != sb[i - 1];
    start[i] = st ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_seg_write(const uint32_t* __restrict__ srow, const uint32_t* __restrict__ sa,
                                                   const uint32_t* __restrict__ sb, const uint32_t* __restrict__ start,
                                                   const uint32_t* __restrict__ segpos, long begin, long end, int H,
                                                   SegTables o) {
    const long i = begin + (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const uint32_t seg = segpos[i] + start[i] - 1;                 // inclusive scan - 1
    const uint32_t row = srow[i];
    if (start[i]) {
        o.t[seg] = (int32_t)(row / (uint32_t)H); o.y0[seg] = (int32_t)(row % (uint32_t)H); o.a[seg] = sa[i]; o.b[seg] = sb[i];
    }
    if (i == end - 1 || start[i + 1]) o.y1[seg] = (int32_t)(row % (uint32_t)H) + 1;
}

// ---------------------------------------------------------------------------------------------------------------
// paint: bit rows + value per run -> int32 cube.  One warp per row, 1024 cells (32 mask words) per pass; a lane owns
// four consecutive cells and stores them as one 16-byte streaming store.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_run_values(const uint32_t* __restrict__ run_comp,
                                                    const int32_t* __restrict__ comp_val, int32_t* __restrict__ run_val,
                                                    long nruns) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nruns) run_val[r] = comp_val[run_comp[r]];
}

// streaming zero fill (16-byte stores); runs on a side stream while the table phase is busy
// Short-lived blocks (64 KB each) on purpose: the fill runs on a low-priority stream underneath the latency-bound
// table kernels, which can only get onto an SM when a block of the fill retires.
constexpr int ZERO_PER_THREAD = 16;
__global__ void __launch_bounds__(256) k_zero_fill(int4* __restrict__ p, size_t n16, int32_t* __restrict__ tail, int ntail) {
    const int4 z = make_int4(0, 0, 0, 0);
    const size_t base = (size_t)blockIdx.x * (256 * ZERO_PER_THREAD) + threadIdx.x;
#pragma unroll
    for (int i = 0; i < ZERO_PER_THREAD; ++i) {
        const size_t j = base + (size_t)i * 256;
        if (j < n16) __stcs(p + j, z);
    }
    if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] = 0;
}

template <bool VEC4>
__global__ void __launch_bounds__(256) k_paint(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ row_ptr,
                                               const int32_t* __restrict__ run_val, long nrows, int W, int Ww,
                                               int32_t* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long row = warp0; row < nrows; row += nwarps) {
        const uint32_t rbase = row_ptr[row];
        const bool empty = row_ptr[row + 1] == rbase;
        int32_t* out = flag + row * (long)W;
        const uint32_t* b = bits + row * (long)Ww;
        uint32_t prev = 0, running = rbase;
        for (int k0 = 0; k0 < Ww; k0 += 32) {
            const int k = k0 + lane;
            const uint32_t m = (!empty && k < Ww) ? b[k] : 0u;
            const bool any = __ballot_sync(FULL, m != 0) != 0;
            uint32_t starts = 0, base = 0;
            if (any) {
                uint32_t left = __shfl_up_sync(FULL, m, 1);
                if (lane == 0) left = prev;
                starts = m & ~((m << 1) | (left >> 31));
                uint32_t tot;
                base = running + warp_excl_scan(__popc(starts), lane, &tot);
                running += tot;
                prev = __shfl_sync(FULL, m, 31);
            } else {
                prev = 0;
            }
            const int xbase = k0 * 32;
            if (VEC4) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int x = xbase + (i * 32 + lane) * 4;
                    int4 v = make_int4(0, 0, 0, 0);
                    if (any) {
                        const int src = 4 * i + (lane >> 3);
                        const uint32_t mw = __shfl_sync(FULL, m, src);
                        const uint32_t sw = __shfl_sync(FULL, starts, src);
                        const uint32_t bw = __shfl_sync(FULL, base, src);
                        const int off = (lane & 7) * 4;
                        const uint32_t nib = (mw >> off) & 0xfu;
                        if (nib) {
                            int vv[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                vv[j] = 0;
                                if ((nib >> j) & 1u) {
                                    const uint32_t rank = bw + __popc(sw & ((2u << (off + j)) - 1u)) - 1u;
                                    vv[j] = run_val[rank];
                                }
                            }
                            v = make_int4(vv[0], vv[1], vv[2], vv[3]);
                        }
                    }
                    if (x < W) __stcs(reinterpret_cast<int4*>(out + x), v);
                }
            } else {
                for (int i = 0; i < 32; ++i) {
                    const int x = xbase + i * 32 + lane;
                    if (xbase + i * 32 >= W) break;
                    int v = 0;
                    if (any) {
                        const uint32_t mw = __shfl_sync(FULL, m, i);
                        const uint32_t sw = __shfl_sync(FULL, starts, i);
                        const uint32_t bw = __shfl_sync(FULL, base, i);
                        if ((mw >> lane) & 1u) v = run_val[bw + __popc(sw & ((2u << lane) - 1u)) - 1u];
                    }
                    if (x < W) __stcs(out + x, v);
                }
            }
        }
    }
}

// Sparse paint by runs (the cube is already zero): a warp takes 32 consecutive runs -- their (x0, x1), row and value come
// in with coalesced loads -- and writes them one after the other, 32 cells per store instruction.  No bit rows, no
// dependent gathers.  Runs of removed components (value 0) are skipped.
// by_comp: the value of a run is comp_val[run_comp[r]] (no per-run value table is materialised).
// (Measured and dropped: every lane painting its own run with 16-byte stores -- 32 runs in flight per warp, 15x fewer
// instructions -- is slower, 1.59 ms against 1.20 ms at 10957 planes: a store instruction then touches 32 different rows, and
// the kernel is bound by how the scattered partial lines reach DRAM, not by instruction issue.)
__global__ void __launch_bounds__(256) k_paint_runs(const uint32_t* __restrict__ row_ptr, long r0, long nrows,
                                                    const uint32_t* __restrict__ run_x,
                                                    const uint32_t* __restrict__ run_row,
                                                    const int32_t* __restrict__ run_val,
                                                    const uint32_t* __restrict__ run_comp,
                                                    const int32_t* __restrict__ comp_val, int W,
                                                    int32_t* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const long run_begin = row_ptr[r0], run_end = row_ptr[r0 + nrows];
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long base = run_begin + warp0 * 32; base < run_end; base += nwarps * 32) {
        const long r = base + lane;
        int x0 = 0, x1 = 0, v = 0;
        long dst = 0;
        if (r < run_end) {
            const uint32_t x = run_x[r];
            x0 = x & 0xffff; x1 = x >> 16; v = run_comp ? comp_val[run_comp[r]] : run_val[r];
            dst = ((long)run_row[r] - r0) * (long)W;
        }
        uint32_t todo = __ballot_sync(FULL, v != 0);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const int jx0 = __shfl_sync(FULL, x0, j), jx1 = __shfl_sync(FULL, x1, j), jv = __shfl_sync(FULL, v, j);
            const long jdst = __shfl_sync(FULL, dst, j);
            for (int xx = jx0 + lane; xx < jx1; xx += 32) flag[jdst + xx] = jv;
        }
    }
}

__global__ void k_paint_overrides(const int32_t* __restrict__ t, const int32_t* __restrict__ y,
                                  const int32_t* __restrict__ x0, const int32_t* __restrict__ x1,
                                  const int32_t* __restrict__ val, long n, int H, int W, long t_begin, long t_end,
                                  int32_t* __restrict__ flag) {
    const long i = blockIdx.x;                                   // one block per sub-run
    if (i >= n || t[i] < t_begin || t[i] >= t_end) return;
    int32_t* out = flag + ((long)(t[i] - t_begin) * H + y[i]) * (long)W;
    for (int x = x0[i] + threadIdx.x; x < x1[i]; x += blockDim.x) out[x] = val[i];
}

inline unsigned blocks_for(long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
cudaError_t threshold_bits(const ThresholdArgs& a, int sm_count, cudaStream_t st) {
    const long nrows = a.T * a.H;
    if (nrows == 0) return cudaSuccess;
    long want = (nrows + 7) / 8;
    int blocks = (int)(want < (long)sm_count * 8 ? want : (long)sm_count * 8);
    if (a.variant != 0) {
        if (a.in_dtype == 1 && bulk_ok<double>(a)) return launch_threshold_bulk<double, false>(a, sm_count, st);
        if (a.in_dtype == 0 && bulk_ok<float>(a)) {
            if (a.thr_is_f32) return launch_threshold_bulk<float, true>(a, sm_count, st);
            return launch_threshold_bulk<float, false>(a, sm_count, st);
        }
    }
    // no bulk staging possible (row bytes not a multiple of 16, or rows too long for shared memory): deep plain loads
    if (a.in_dtype == 1) return launch_threshold<double, false>(a, blocks, st);
    if (a.thr_is_f32) return launch_threshold<float, true>(a, blocks, st);
    return launch_threshold<float, false>(a, blocks, st);
}

cudaError_t row_stats(const uint32_t* bits, long nrows, int W, int Ww, uint32_t* row_cnt, uint32_t* seam_flag,
                      uint32_t* slots, uint32_t* overflow, cudaStream_t st) {
    if (nrows == 0) return cudaSuccess;
    k_row_stats<<<(unsigned)((nrows + 7) / 8), 256, 0, st>>>(bits, nrows, W, Ww, row_cnt, seam_flag, slots, overflow);
    return cudaGetLastError();
}

size_t scan_tmp_elems(long n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 2; }

cudaError_t exclusive_scan_u32(const uint32_t* in, uint32_t* out, long n, uint32_t* tmp, cudaStream_t st, uint32_t base,
                               uint32_t* wide_flag) {
    if (n == 0) {
        k_scan_sums<<<1, 1024, 0, st>>>(tmp, 0, out, base, wide_flag);
        return cudaGetLastError();
    }
    const long nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, tmp);
    k_scan_sums<<<1, 1024, 0, st>>>(tmp, nb, out + n, base, wide_flag);
    k_scan_final<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, tmp);
    return cudaGetLastError();
}

cudaError_t compact_runs(const uint32_t* slots, const uint32_t* bits, const uint32_t* row_ptr, long row0, long nrows, int Ww,
                         uint32_t* ovf_rows, uint32_t* ovf_count, uint32_t* run_x, uint32_t* run_row, cudaStream_t st) {
    if (nrows == 0) return cudaSuccess;
    if (ovf_rows) {
        cudaError_t e = cudaMemsetAsync(ovf_count, 0, 4, st);
        if (e != cudaSuccess) return e;
    }
    k_compact_runs<<<blocks_for(nrows, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(slots + row0 * (long)RUN_SLOTS),
                                                           row_ptr + row0, nrows, run_x, run_row, row0, ovf_rows, ovf_count);
    if (ovf_rows)
        k_extract_runs<<<148 * 2, 256, 0, st>>>(bits + row0 * (long)Ww, row_ptr + row0, nrows, Ww, run_x, run_row,
                                                           row0, ovf_rows, ovf_count);
    return cudaGetLastError();
}

cudaError_t ccl_init(uint32_t* parent, long begin, long end, cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    k_iota<<<blocks_for(end - begin, 256), 256, 0, st>>>(parent, begin, end);
    return cudaGetLastError();
}

cudaError_t ccl_union(const uint32_t* row_ptr, const uint32_t* run_x, const uint32_t* run_row, long begin, long end,
                      int H, uint32_t* parent, cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    k_ccl_union<<<blocks_for(end - begin, 256), 256, 0, st>>>(row_ptr, run_x, run_row, begin, end, H, parent);
    return cudaGetLastError();
}

cudaError_t ccl_flatten(uint32_t* parent, uint32_t* root_flag, long begin, long end, cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    k_ccl_flatten<<<blocks_for(end - begin, 256), 256, 0, st>>>(parent, root_flag, begin, end);
    return cudaGetLastError();
}

cudaError_t ccl_assign(const uint32_t* parent, const uint32_t* rank, uint32_t* run_comp, long begin, long end,
                       cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    k_ccl_assign<<<blocks_for(end - begin, 256), 256, 0, st>>>(parent, rank, run_comp, begin, end);
    return cudaGetLastError();
}

cudaError_t comp_init(const CompTables& c, long begin, long end, int W, cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    k_comp_init<<<blocks_for(end - begin, 256), 256, 0, st>>>(c, begin, end, W);
    return cudaGetLastError();
}

cudaError_t comp_accumulate(const uint32_t* run_x, const uint32_t* run_row, const uint32_t* run_comp, long begin, long end,
                            int H, const double* w_dev, const uint8_t* special_dev, const CompTables& c, cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    k_comp_accumulate<<<blocks_for(end - begin, 256), 256, 0, st>>>(run_x, run_row, run_comp, begin, end, H, w_dev,
                                                                    special_dev, c);
    return cudaGetLastError();
}

cudaError_t seam_rows(const uint32_t* seam_flag, const uint32_t* seam_pos, const uint32_t* row_ptr,
                      const uint32_t* run_comp, long row_begin, long row_end, uint32_t* seam_row, uint32_t* seam_a,
                      uint32_t* seam_b, uint32_t* cls_parent, cudaStream_t st) {
    if (row_end <= row_begin) return cudaSuccess;
    k_seam_rows<<<blocks_for(row_end - row_begin, 256), 256, 0, st>>>(seam_flag, seam_pos, row_ptr, run_comp, row_begin,
                                                                      row_end, seam_row, seam_a, seam_b, cls_parent);
    return cudaGetLastError();
}

cudaError_t cls_flatten(uint32_t* cls_parent, long begin, long end, cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    k_cls_flatten<<<blocks_for(end - begin, 256), 256, 0, st>>>(cls_parent, begin, end);
    return cudaGetLastError();
}

cudaError_t pairs_init(const PairTable& p, cudaStream_t st) {
    k_pairs_init<<<blocks_for(p.cap, 256), 256, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t pairs_accumulate(const uint32_t* row_ptr, const uint32_t* run_x, const uint32_t* run_row,
                             const uint32_t* run_comp, long begin, long end, int H, const double* w_dev,
                             const uint8_t* special_dev, const PairTable& p, cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    k_pairs_accumulate<<<blocks_for(end - begin, 256), 256, 0, st>>>(row_ptr, run_x, run_row, run_comp, begin, end, H,
                                                                     w_dev, special_dev, p);
    return cudaGetLastError();
}

cudaError_t pairs_compact(const PairTable& p, uint32_t* out_a, uint32_t* out_b, uint32_t* out_npix, uint32_t* out_nsp,
                          double* out_E, double* out_S, uint32_t out_cap, uint32_t* count_dev, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(count_dev, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    k_pairs_compact<<<blocks_for(p.cap, 256), 256, 0, st>>>(p, out_a, out_b, out_npix, out_nsp, out_E, out_S, out_cap,
                                                            count_dev);
    return cudaGetLastError();
}

cudaError_t run_values(const uint32_t* run_comp, const int32_t* comp_val, int32_t* run_val, long nruns,
                       cudaStream_t st) {
    if (nruns == 0) return cudaSuccess;
    k_run_values<<<blocks_for(nruns, 256), 256, 0, st>>>(run_comp, comp_val, run_val, nruns);
    return cudaGetLastError();
}

cudaError_t paint(const PaintArgs& a, int sm_count, cudaStream_t st) {
    if (a.nrows == 0) return cudaSuccess;
    long want = (a.nrows + 7) / 8;
    int blocks = (int)(want < (long)sm_count * 8 ? want : (long)sm_count * 8);
    const bool vec = (a.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.flag) & 15) == 0);
    if (a.sparse) {
        k_paint_runs<<<sm_count * 8, 256, 0, st>>>(a.row_ptr - a.row0, a.row0, a.nrows, a.run_x, a.run_row, a.run_val,
                                                   a.run_comp, a.comp_val, a.W, a.flag);
    } else {
        if (vec) k_paint<true><<<blocks, 256, 0, st>>>(a.bits, a.row_ptr, a.run_val, a.nrows, a.W, a.Ww, a.flag);
        else k_paint<false><<<blocks, 256, 0, st>>>(a.bits, a.row_ptr, a.run_val, a.nrows, a.W, a.Ww, a.flag);
    }
    return cudaGetLastError();
}

// The fill runs beside the latency-bound table kernels.  Left alone it fills every thread slot of every SM (8 blocks of 256
// threads), and a table kernel's blocks then have to wait for fill blocks to retire -- measured: the plane kernel takes 1.11 ms
// instead of 0.43 ms, the cooperative global kernel 0.82 ms instead of 0.14 ms (1370 planes).  HBM writes saturate with far
// fewer threads, so the fill is capped at `ctas_per_sm` blocks per SM through an (unused) dynamic shared-memory request.
cudaError_t zero_fill(int32_t* p, size_t n, int sm_count, cudaStream_t st, int ctas_per_sm) {
    if (n == 0) return cudaSuccess;
    // head up to 16-byte alignment is handled by shifting into the tail path only when misaligned (rare: cudaMalloc
    // and torch allocations are 256/512-byte aligned)
    if (reinterpret_cast<uintptr_t>(p) & 15) return cudaMemsetAsync(p, 0, n * 4, st);
    const size_t n16 = n / 4;
    const size_t per_block = 256 * ZERO_PER_THREAD;
    (void)sm_count;
    size_t smem = 0;
    if (ctas_per_sm >= 1 && ctas_per_sm < 8) {
        smem = (size_t)(227 * 1024) / (size_t)ctas_per_sm - 1024;            // (1 KB per block is reserved by the system)
        static size_t configured = 0;
        if (smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(k_zero_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured = smem;
        }
    }
    k_zero_fill<<<(unsigned)((n16 + per_block - 1) / per_block + (n16 == 0)), 256, smem, st>>>(
        reinterpret_cast<int4*>(p), n16, p + n16 * 4, (int)(n - n16 * 4));
    return cudaGetLastError();
}

cudaError_t class_sums(const CompTables& c, const ClassTables& k, long begin, long end, cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    k_class_sums<<<blocks_for(end - begin, 256), 256, 0, st>>>(c, k, begin, end);
    return cudaGetLastError();
}

cudaError_t pairs_count(const PairTable& p, const uint32_t* cls, const ClassTables& k, uint32_t* pcnt, cudaStream_t st) {
    k_pairs_count<<<blocks_for(p.cap, 256), 256, 0, st>>>(p, cls, k, pcnt);
    return cudaGetLastError();
}

cudaError_t pairs_fill(const PairTable& p, const uint32_t* pptr, uint32_t* pfill, const PairCsr& o, cudaStream_t st) {
    k_pairs_fill<<<blocks_for(p.cap, 256), 256, 0, st>>>(p, pptr, pfill, o);
    return cudaGetLastError();
}

cudaError_t seg_flags(const uint32_t* srow, const uint32_t* sa, const uint32_t* sb, long begin, long end, int H,
                      uint32_t* start, cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    k_seg_flags<<<blocks_for(end - begin, 256), 256, 0, st>>>(srow, sa, sb, begin, end, H, start);
    return cudaGetLastError();
}

cudaError_t seg_write(const uint32_t* srow, const uint32_t* sa, const uint32_t* sb, const uint32_t* start,
                      const uint32_t* segpos, long begin, long end, int H, const SegTables& o, cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    k_seg_write<<<blocks_for(end - begin, 256), 256, 0, st>>>(srow, sa, sb, start, segpos, begin, end, H, o);
    return cudaGetLastError();
}

cudaError_t paint_overrides(const int32_t* t, const int32_t* y, const int32_t* x0, const int32_t* x1,
                            const int32_t* val, long n, int H, int W, long t_begin, long t_end, int32_t* flag,
                            cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_paint_overrides<<<(unsigned)n, 128, 0, st>>>(t, y, x0, x1, val, n, H, W, t_begin, t_end, flag);
    return cudaGetLastError();
}

}  // namespace ctk
