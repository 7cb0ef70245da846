// ct_kernels.h -- launchers of the sm_100a kernels behind ct_run_contrack (definitions in ct_kernels.cu).
// Every launcher enqueues on `st` and returns the cudaError_t of the launch (cudaGetLastError()).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ctk {

constexpr uint64_t PAIR_EMPTY = ~0ull;
constexpr int RUN_SLOTS_PER_ROW = 8;       // runs per row the threshold kernel leaves in the row's slots (32 bytes per row)

struct ThresholdArgs {
    const void* anom; int in_dtype;          // ct_dtype
    long T; int H, W, Ww;                    // Ww = words of 32 mask bits per row
    const double* thr_dev; long thr_n;       // thresholds on the device (1 or T values)
    int thr_is_f32, op;                      // ct_op
    uint32_t* bits;                          // [T*H*Ww] out
    uint32_t* row_cnt;                       // [T*H] out: row-runs per row
    uint32_t* seam_flag;                     // [T*H] out: 1 if the pixels at x=0 and x=W-1 are both set
    uint32_t* slots;                         // [T*H*RUN_SLOTS_PER_ROW] out: first runs of every row, x0 | x1 << 16
    uint32_t* overflow;                      // device flag, set if a row has more runs than slots
    int bits_overflow_only = 0;              // 1: bit rows are written only for rows with more runs than slots (nothing else
                                             // reads them when the runs come from the slots and the paint goes by runs)
    int variant;                             // 0: 4-byte loads + ballot, 16 loads in flight per lane; otherwise cp.async.bulk
                                             // row staging (needs 16-byte aligned rows that fit shared memory, else as 0)
};
cudaError_t threshold_bits(const ThresholdArgs& a, int sm_count, cudaStream_t st);

// row_cnt / seam_flag of existing bit rows (halo plane of a time-sharded run)
cudaError_t row_stats(const uint32_t* bits, long nrows, int W, int Ww, uint32_t* row_cnt, uint32_t* seam_flag,
                      uint32_t* slots, uint32_t* overflow, cudaStream_t st);

// The table kernels work on index RANGES [begin, end) of the global row / run / component tables, so that the planes of a
// cube can be processed in time chunks while later planes are still being thresholded (ct_api.cu: tables_chunk).
//
// out[0..n] = base + exclusive prefix sums of in[0..n-1] (out[n] = base + total).  `tmp` needs scan_tmp_elems(n) uint32.
size_t scan_tmp_elems(long n);
// wide_flag (device, may be null) is set to 1 if base + total needs more than 32 bits (the output has wrapped).
cudaError_t exclusive_scan_u32(const uint32_t* in, uint32_t* out, long n, uint32_t* tmp, cudaStream_t st,
                               uint32_t base = 0, uint32_t* wide_flag = nullptr);

// rows [row0, row0 + nrows): row-runs from the row slots the threshold kernel filled -> run_x, run_row at row_ptr[row]...; when ovf_rows is given ([nrows] scratch + a device counter), rows
// with more runs than slots are listed there and re-extracted from their bit rows
cudaError_t compact_runs(const uint32_t* slots, const uint32_t* bits, const uint32_t* row_ptr, long row0, long nrows, int Ww,
                         uint32_t* ovf_rows, uint32_t* ovf_count, uint32_t* run_x, uint32_t* run_row, cudaStream_t st);

// 2-D 8-connected components over row-runs (contrack.py:684-687): union-find with the smallest run index as root.
cudaError_t ccl_init(uint32_t* parent, long begin, long end, cudaStream_t st);
cudaError_t ccl_union(const uint32_t* row_ptr, const uint32_t* run_x, const uint32_t* run_row, long begin, long end,
                      int H, uint32_t* parent, cudaStream_t st);
cudaError_t ccl_flatten(uint32_t* parent, uint32_t* root_flag, long begin, long end, cudaStream_t st);
// run_comp[r] = rank[parent[r]]  (rank = base + exclusive scan of root_flag)
cudaError_t ccl_assign(const uint32_t* parent, const uint32_t* rank, uint32_t* run_comp, long begin, long end,
                       cudaStream_t st);

struct CompTables {
    int32_t *t, *y0, *y1, *x0, *x1;          // [ncomp]
    double *areaE, *areaS;
    uint32_t *nsp, *cls;
};
cudaError_t comp_init(const CompTables& c, long begin, long end, int W, cudaStream_t st);
cudaError_t comp_accumulate(const uint32_t* run_x, const uint32_t* run_row, const uint32_t* run_comp, long begin, long end,
                            int H, const double* w_dev, const uint8_t* special_dev, const CompTables& c, cudaStream_t st);

// date-line rows (contrack.py:691-698): list (row, comp at x=0, comp at x=W-1) at position seam_pos[row]; class union
cudaError_t seam_rows(const uint32_t* seam_flag, const uint32_t* seam_pos, const uint32_t* row_ptr,
                      const uint32_t* run_comp, long row_begin, long row_end, uint32_t* seam_row, uint32_t* seam_a,
                      uint32_t* seam_b, uint32_t* cls_parent, cudaStream_t st);
cudaError_t cls_flatten(uint32_t* cls_parent, long begin, long end, cudaStream_t st);

struct PairTable {
    unsigned long long* key;                 // [cap] (a << 32) | b, PAIR_EMPTY when free
    uint32_t *npix, *nsp;                    // [cap]
    double *areaE, *areaS;                   // [cap]
    uint32_t cap;                            // power of two
    uint32_t* overflow;                      // device flag
    uint32_t* count;                         // device counter: occupied slots (= pairs)
};
cudaError_t pairs_init(const PairTable& p, cudaStream_t st);
// two-timestep intersection (contrack.py:717-719 on tables): for every run of plane t >= 1 the overlapping runs of
// plane t-1 in the same row; pixel counts and areas accumulate per (comp_t, comp_t-1)
cudaError_t pairs_accumulate(const uint32_t* row_ptr, const uint32_t* run_x, const uint32_t* run_row,
                             const uint32_t* run_comp, long begin, long end, int H, const double* w_dev,
                             const uint8_t* special_dev, const PairTable& p, cudaStream_t st);
// compact occupied slots: out arrays sized >= number of pairs; *count_dev receives the number
cudaError_t pairs_compact(const PairTable& p, uint32_t* out_a, uint32_t* out_b, uint32_t* out_npix, uint32_t* out_nsp,
                          double* out_E, double* out_S, uint32_t out_cap, uint32_t* count_dev, cudaStream_t st);

// ---- tables in the layout of cth::FastTables ----
struct ClassTables { double *conE, *conS, *fE, *fS; uint32_t *nsp, *fnsp; };   // [ncomp], zero-initialised by caller
struct PairCsr { uint32_t *b, *npix, *nsp; double *E, *S; };                   // [npair]
struct SegTables { int32_t *t, *y0, *y1; uint32_t *a, *b; };                   // [nseg <= nseam]
cudaError_t class_sums(const CompTables& c, const ClassTables& k, long begin, long end, cudaStream_t st);
// pcnt[a]++ per pair, forward sums into the class of b (pcnt zeroed by caller)
cudaError_t pairs_count(const PairTable& p, const uint32_t* cls, const ClassTables& k, uint32_t* pcnt, cudaStream_t st);
// pptr = exclusive scan of pcnt; pfill zeroed by caller
cudaError_t pairs_fill(const PairTable& p, const uint32_t* pptr, uint32_t* pfill, const PairCsr& o, cudaStream_t st);
cudaError_t seg_flags(const uint32_t* srow, const uint32_t* sa, const uint32_t* sb, long begin, long end, int H,
                      uint32_t* start, cudaStream_t st);
// segpos = base + exclusive scan of start
cudaError_t seg_write(const uint32_t* srow, const uint32_t* sa, const uint32_t* sb, const uint32_t* start,
                      const uint32_t* segpos, long begin, long end, int H, const SegTables& o, cudaStream_t st);

// ctas_per_sm: 1..7 caps the resident blocks per SM (room for the table kernels beside the fill); 0 / 8 = no cap
cudaError_t zero_fill(int32_t* p, size_t n, int sm_count, cudaStream_t st, int ctas_per_sm = 0);

// run_val[r] = comp_val[run_comp[r]]
cudaError_t run_values(const uint32_t* run_comp, const int32_t* comp_val, int32_t* run_val, long nruns,
                       cudaStream_t st);

struct PaintArgs {
    const uint32_t* bits; const uint32_t* row_ptr; const int32_t* run_val;
    long nrows; int W, Ww;
    int32_t* flag;                           // [nrows * W] out
    int sparse;                              // 0: dense (every cell written, from the bit rows); 1: `flag` is already zero, only
                                             // the cells of row-runs with a value are written
    const uint32_t* run_x; const uint32_t* run_row;   // sparse
    const uint32_t* run_comp = nullptr;      // sparse, optional: value of run r = comp_val[run_comp[r]] instead of run_val[r]
    const int32_t* comp_val = nullptr;
    long row0;                               // first row painted (row_ptr points at it)
};
cudaError_t paint(const PaintArgs& a, int sm_count, cudaStream_t st);
// sub-runs of planes [t_begin, t_end) only; `flag` starts at plane t_begin
cudaError_t paint_overrides(const int32_t* t, const int32_t* y, const int32_t* x0, const int32_t* x1,
                            const int32_t* val, long n, int H, int W, long t_begin, long t_end, int32_t* flag,
                            cudaStream_t st);

}  // namespace ctk

// ---- run_lifecycle kernels (ct_lifecycle.cu; reference contrack.py:799-907) -------------------------------------------
namespace ctl {

struct EntryTable {                          // hash set of (time step << 32 | flag id)
    unsigned long long* key;                 // [cap], ctk::PAIR_EMPTY when free
    uint32_t *npix, *flags;                  // [cap] cells; bit 0: touches column 0, bit 1: touches column W-1
    uint32_t cap;                            // power of two
    uint32_t *overflow, *count, *nroll;      // device counters
};
// flag cube -> non-zero bits / run-start bits [nrows * Ww] and runs per row
cudaError_t lc_rows(const int32_t* flag, long nrows, int W, int Ww, uint32_t* nz, uint32_t* st, uint32_t* row_cnt,
                    int sm_count, cudaStream_t stream);
// -> row-runs in raster order: run_x = x0 | x1 << 16 (x1 exclusive), run_row, run_label (row_ptr = scan of row_cnt)
cudaError_t lc_extract(const int32_t* flag, const uint32_t* nz, const uint32_t* st, const uint32_t* row_ptr, long nrows,
                       int W, int Ww, uint32_t* run_x, uint32_t* run_row, int32_t* run_label, cudaStream_t stream);
cudaError_t lc_entries(const uint32_t* run_x, const uint32_t* run_row, const int32_t* run_label, long nruns, int H, int W,
                       const EntryTable& e, cudaStream_t stream);
// occupied slots -> entry arrays (any order); out_roll = -1, or -(bitmap slot + 2) for entries that lc_roll must resolve
cudaError_t lc_compact(const EntryTable& e, int32_t* out_t, int32_t* out_label, uint32_t* out_npix, int32_t* out_roll,
                       uint32_t* fill_zeroed, cudaStream_t stream);
cudaError_t lc_roll(const uint32_t* row_ptr, const uint32_t* run_x, const int32_t* run_label, const int32_t* ent_t,
                    const int32_t* ent_label, int32_t* ent_roll, long nent, int H, int W, int Ww, uint32_t* bitmaps_zeroed,
                    cudaStream_t stream);
cudaError_t lc_sums(const uint32_t* row_ptr, const uint32_t* run_x, const uint32_t* run_row, const int32_t* run_label,
                    const void* var, int var_is_f64, const double* w, const int32_t* ent_t, const int32_t* ent_label,
                    const uint32_t* ent_npix, const int32_t* ent_roll, long nent, int H, int W, double* out_area,
                    double* out_int, double* out_norm, double* out_sy, double* out_sx, cudaStream_t stream);

}  // namespace ctl
