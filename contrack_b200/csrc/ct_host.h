// ct_host.h -- the strictly ordered part of run_contrack, on component tables (pure host C++, no CUDA).
//
// The CUDA kernels reduce the cube to tables (2-D components, date-line classes, adjacent-plane pair areas, date-line
// rows).  This translation unit replays, on those tables and in the reference's order:
//   step 3   contrack/contrack.py:706-742   time-sequential forward/backward overlap filter
//   step 4a/b contrack.py:747-751           3-D labelling of the kept mask, ids in scipy's first-pixel order
//   step 4c/d contrack.py:753-772           date-line merge through stale boxes + persistence (ct_tables.cpp)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "ct_tables.h"

namespace cth {

struct PlaneRun { int y, x0, x1; uint32_t comp; };          // x1 exclusive; comp = global 2-D component id

// Row-runs of one time plane in raster order (only needed for near-tie decisions and stale-box splits: rare).
struct RunSource {
    virtual bool plane_runs(long t, std::vector<PlaneRun>& out) = 0;
    virtual ~RunSource() {}
};

struct Tables {
    long T = 0; int H = 0, W = 0;
    // 2-D components (8-connected, no date-line wrap) in global first-pixel order: id ascending <=> (t, y, x) ascending
    long ncomp = 0;
    const int32_t* comp_t = nullptr;
    const int32_t *comp_y0 = nullptr, *comp_y1 = nullptr, *comp_x0 = nullptr, *comp_x1 = nullptr;   // half-open bbox
    const uint32_t* comp_cls = nullptr;      // date-line class (contrack.py:691-698): smallest component id of the class
    const double* comp_areaE = nullptr;      // sum of w[y] over pixels in exactly-summable rows
    const double* comp_areaS = nullptr;      // ... over the other ("special", e.g. pole) rows
    const uint32_t* comp_nsp = nullptr;      // pixel count in special rows
    // pairs (a at plane t, b at plane t-1) with at least one common pixel
    long npair = 0;
    const uint32_t *pair_a = nullptr, *pair_b = nullptr, *pair_npix = nullptr, *pair_nsp = nullptr;
    const double *pair_areaE = nullptr, *pair_areaS = nullptr;
    // date-line rows: row = t*H + y, pixel x=0 belongs to component a, pixel x=W-1 to component b; sorted by row
    long nseam = 0;
    const uint32_t *seam_row = nullptr, *seam_a = nullptr, *seam_b = nullptr;
    const double* w = nullptr;               // [H] area weight per row (contrack.py:703-704)
    int special_uniform = 0;                 // see FastTables
};

// The layout the CUDA kernels deliver (everything the ordered phase reads is sequential in memory):
// class sums indexed by the representative component, pairs in CSR form over their plane-t component, date-line rows
// already grouped into segments of consecutive rows with the same two components.
struct FastTables {
    long T = 0; int H = 0, W = 0;
    long ncomp = 0;
    const int32_t* comp_t = nullptr;
    const int32_t *comp_y0 = nullptr, *comp_y1 = nullptr, *comp_x0 = nullptr, *comp_x1 = nullptr;
    const uint32_t* comp_cls = nullptr;
    // per class, stored at the representative's index: own area and forward overlap (all of plane t+1), E / S parts,
    // and the number of special-row pixels in those two sums
    const double *cls_conE = nullptr, *cls_conS = nullptr, *cls_fE = nullptr, *cls_fS = nullptr;
    const uint32_t* cls_nsp = nullptr;       // special-row pixels in the own-area sum (+ forward sum if cls_fnsp is null)
    const uint32_t* cls_fnsp = nullptr;      // special-row pixels in the forward sum (optional)
    // pairs of component c with components of plane t-1: entries pair_ptr[c] .. pair_ptr[c+1]
    const uint32_t* pair_ptr = nullptr;
    const uint32_t *pair_b = nullptr, *pair_npix = nullptr, *pair_nsp = nullptr;
    const double *pair_E = nullptr, *pair_S = nullptr;
    long nseg = 0;
    const int32_t *seg_t = nullptr, *seg_y0 = nullptr, *seg_y1 = nullptr;
    const uint32_t *seg_a = nullptr, *seg_b = nullptr;
    const double* w = nullptr;
    int special_uniform = 0;                 // 1: all special rows carry the same weight (the two poles of a regular grid)
};

struct Params {
    double overlap = 0.5;
    int persistence = 1;
    int twosided = 1;
    int stage = 0;                           // ct_stage
};

struct Result {
    std::vector<int32_t> comp_val;           // value painted for each component (host_phase only)
    std::vector<ctb::Override> overrides;    // sub-runs of split components (stage FINAL only)
    long n_features = 0, n_kept = 0, n_labels3d = 0, n_seam_events = 0, n_seam_splits = 0, n_neartie = 0;
};

// returns 0, or a negative ct_status with `err` filled.  `comp_val` [ncomp] receives the value painted per component.
int host_phase_fast(const FastTables& tb, const Params& pr, RunSource* runs, int32_t* comp_val, Result& out,
                    std::string& err);
// Steps 4c/4d only (date-line merge through stale boxes + persistence) for 3-D labels computed elsewhere (on the device):
// uses tb's component boxes and segments; label[c] = 0 marks removed components.
int track_phase(const FastTables& tb, const int32_t* label, int persistence, RunSource* runs, int32_t* comp_val,
                Result& out, std::string& err);
// same from the unsorted tables (converts, then calls host_phase_fast); result in out.comp_val
int host_phase(const Tables& tb, const Params& pr, RunSource* runs, Result& out, std::string& err);

}  // namespace cth
