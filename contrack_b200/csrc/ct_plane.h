// ct_plane.h -- per-plane table kernel (ct_plane.cu) and the cooperative global-phase kernel (ct_global.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "ct_kernels.h"

namespace ctp {

constexpr int PLANE_THREADS = 256;
constexpr int RUN_SLOTS = ctk::RUN_SLOTS_PER_ROW;

// bits of *status
constexpr uint32_t ST_FALLBACK = 1u;     // a plane does not fit the shared-memory budget / its pair hash overflowed
constexpr uint32_t ST_CAPACITY = 2u;     // a table is smaller than the totals the chain reports: grow and run again
constexpr uint32_t ST_TIMEOUT = 4u;      // a look-back / predecessor wait gave up (internal error, never expected)

struct PlaneArgs {
    // what the threshold kernel left per row
    const uint32_t *row_cnt, *seam_flag, *slots, *bits;
    int H, W, Ww;
    long p0, np;                             // planes [p0, p0 + np) of the context's scratch (a halo plane is plane 0)
    const double* w; const uint8_t* special;
    // global tables (indices continue from plane to plane)
    uint32_t* row_ptr;                       // [planes * H + 1]
    uint32_t *run_x, *run_row, *run_comp;    // [cap_runs]
    ctk::CompTables ct; ctk::ClassTables kt; // [cap_comps]
    uint32_t *pcnt, *pfill, *pptr;           // [cap_comps + 1]
    ctk::PairCsr pc;                         // [cap_pairs]
    ctk::SegTables sg;                       // [cap_segs]
    uint32_t cap_runs, cap_comps, cap_pairs, cap_segs;
    // chain: three arrays of (planes + 1) descriptor words (components << 31 | segments; runs; pairs), slot 0 = sentinel
    unsigned long long* chain; long chain_stride;
    uint32_t* done;                          // [planes + 1] "runs / components / classes of this plane are written"
    uint32_t *ticket, *status;
    uint32_t* info;                          // bit 0: some row had more runs than slots (its runs came from the bit row)
    // shared-memory configuration (plane_config)
    uint32_t smem_runs, hash_cap; size_t smem_scan_off;
};

size_t plane_smem_bytes(int H, uint32_t smem_runs, size_t* scan_off);
bool plane_config(int H, size_t budget, uint32_t* smem_runs, uint32_t* hash_cap);
cudaError_t plane_tables(const PlaneArgs& a, size_t smem_bytes, cudaStream_t st);

// ---- cooperative global phase (ct_global.cu): contrack.py:706-751 + label boxes + date-line events, one launch ----
struct GlobalArgs {
    // tables (device); the counts are read from the chain totals at `totals` = {components, segments, runs, pairs} (u64 x 4)
    const unsigned long long* totals;
    const uint32_t* status;                  // tables are garbage when any bit is set: the kernel returns at once
    uint32_t cap_comps, cap_segs;            // what the scratch arrays hold
    long T;                                  // planes of the cube (components of plane 0 and T-1 are never filtered)
    const int32_t *comp_t, *comp_y0, *comp_y1, *comp_x0, *comp_x1;
    const uint32_t* cls;
    const double *conE, *conS, *fE, *fS; const uint32_t *nsp, *fnsp;
    const uint32_t *pair_ptr, *pair_b, *pair_npix, *pair_nsp; const double *pair_E, *pair_S;
    const uint32_t *seg_a, *seg_b;
    double overlap; int twosided, special_uniform, persistence;
    int max_sweeps;                          // Jacobi sweeps before the plane-ordered wavefront takes over
    // scratch [components + 2] unless noted
    uint8_t* kept; double *accE, *accS; uint32_t* accN;
    uint8_t* dirty;                          // [2 * (T + 2)] planes whose classes must be re-evaluated (ping-pong)
    uint32_t *parent, *rootflag, *rank;
    int32_t* label;                          // 3-D label of every component (0 = removed)
    int32_t *bt0, *bt1, *by0, *by1, *bx0, *bx1;   // label boxes [components + 2]
    int32_t* fin;                            // [components + 2] value painted for every label (persistence applied)
    uint32_t* blocksum;                      // [gridDim.x + 1] grid-wide scan scratch
    // date-line events (segments whose two ends carry different labels), in (t, y) order, as pairs of indices into the
    // records of the labels that occur in events: lrec = {label, t0, t1, y0, y1, x0, x1} sorted by label
    uint32_t* evflag;                        // [segments + 1]
    int32_t* ev;                             // [2 * segments]
    int32_t* lrec;                           // [7 * min(2 * segments, components + 1)]
    // results: {sweeps, near-tie flags, labels, events, features before the host pass, wavefront planes, records, 0}
    uint32_t* out8;
};
int global_grid(int sm_count);
cudaError_t global_phase(const GlobalArgs& a, int grid, cudaStream_t st);

}  // namespace ctp
