// ct_tables.h -- host-side, strictly ordered table phase of run_contrack (no CUDA in this translation unit).
//
// Implements, on component tables instead of pixels:
//   * the step-4 date-line merge that relabels through bounding boxes taken BEFORE merging
//     (reference contrack/contrack.py:753-763), including the rare case where a merged feature is only partly
//     inside the stale box and has to be split at row-run / pixel granularity;
//   * the persistence filter on the merged values (contrack.py:765-772);
//   * numpy's pairwise float64 summation order (used by the near-tie resolver for contrack.py:717-719).
#pragma once
#include <cstdint>
#include <vector>

namespace ctb {

struct SubRun { int y, x0, x1; };                 // x1 exclusive
struct Override { int t, y, x0, x1, val; };

// Supplies the row-runs of one kept 2-D component (only called when a component must be split).
struct RunFetcher {
    virtual bool fetch(long comp, std::vector<SubRun>& out) = 0;
    virtual ~RunFetcher() {}
};

struct TrackStats {
    long n_features = 0, n_events = 0, n_splits = 0;
    long n_walked = 0;                       // members visited by date-line events (track_tables_sparse)
    double ms_init = 0, ms_events = 0, ms_persist = 0;
};

// Returns 0 on success, -1 if a split was needed but the fetcher was missing/failed.
int track_tables(long T, int H, int W, int persistence,
                 long ncomp, const int32_t* comp_t, const int32_t* comp_y0, const int32_t* comp_y1,
                 const int32_t* comp_x0, const int32_t* comp_x1, const int32_t* comp_label,
                 long nseg, const int32_t* seg_t, const int32_t* seg_y0, const int32_t* seg_y1,
                 const int32_t* seg_a, const int32_t* seg_b,
                 RunFetcher* fetcher, int32_t* comp_val, std::vector<Override>& overrides, TrackStats& stats);

// The two steps at LABEL granularity, from the device's event list alone (the product path, ct_fast.cu): a date-line event
// relabels the members of value `hi` that lie inside the stale box of label hi; when every 3-D label that currently carries
// `hi` is either completely inside or completely outside that box, whole labels move and no component has to be looked at.
// `lrec`: one record {label, t0, t1, y0, y1, x0, x1}
// per label that occurs in an event, sorted by label; `ev`: for every date-line segment whose two ends carry different 3-D
// labels, in (t, y) order, the record indices of the label at x = 0 and at x = W-1.  Only labels that occur in events can
// ever change value, so nothing else is needed.  Outputs: for every label an event touched, its final value (patch_label /
// patch_value; 0 = removed by the persistence filter), and *feat_delta = the correction to the number of features the
// device counted under the assumption that no event happened.  Returns 0, or 1 if a label straddles a stale box
// (per-component replay needed; nothing written).
int track_events_fast(int persistence, long nev, const int32_t* ev, long nrec, const int32_t* lrec,
                      std::vector<int32_t>& patch_label, std::vector<int32_t>& patch_value, long* feat_delta,
                      TrackStats& stats);

// np.sum order on a contiguous float64 vector: 0 + pairwise(a, n) with 128-element blocks and 8 accumulators.
double numpy_pairwise_sum(const double* a, long n);

}  // namespace ctb
