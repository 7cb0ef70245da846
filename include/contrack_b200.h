/* contrack_b200.h -- C-ABI of the B200-native ConTrack tracking path.
 *
 * The reference (steidani/ConTrack) has no FFI/plugin layer: its boundary is the Python class method
 * contrack.run_contrack (contrack/contrack.py:583-796) plus calc_clim / calc_anom (contrack.py:458-581).
 * Every entry point below names the reference lines it replaces.  Python owns every caller buffer, the library
 * never frees caller memory, no exception crosses the ABI: functions return 0 on success or a negative
 * ct_status; ct_last_error() gives the message (thread-local).
 *
 * Conventions: cubes are C-order (time, lat, lon); "dev" pointers are CUDA device pointers on the context's
 * device, "host" pointers are ordinary host memory.  A context is bound to one GPU, is not re-entrant and must be
 * used from one host thread at a time.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 */
#ifndef CONTRACK_B200_H
#define CONTRACK_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ct_ctx ct_ctx;
typedef struct ct_comm ct_comm;     /* collectives of the time-sharded entry points, see ct_comm_init_nccl */

enum ct_status {
    CT_OK = 0,
    CT_ERR_ARG = -1,        /* bad argument (also: unknown `op`, the reference's ValueError at contrack.py:658/673) */
    CT_ERR_CUDA = -2,       /* CUDA runtime error, or no usable GPU */
    CT_ERR_CAPACITY = -3,   /* a size exceeds what the index types support: H, W > 65535; T*H >= 2^31; more than 2^32 - 16
                               row-runs in the cube (a row holds up to W/2 runs: dense or noisy masks on very large cubes);
                               more than 2^31 - 1 components or pairs */
    CT_ERR_NEARTIE = -4,    /* overlap decision within rounding distance of the threshold on rows whose area weights
                               are not exactly summable and the exact resolver could not decide */
    CT_ERR_INTERNAL = -5,
    CT_ERR_COMM = -6        /* a collective of a time-sharded run failed (NCCL error, libnccl not found) */
};

enum ct_dtype { CT_F32 = 0, CT_F64 = 1 };
/* gorl of run_contrack (contrack.py:649-656, 664-671): '>=' / 'ge', '<=' / 'le', '>' / 'gt', '<' / 'lt' */
enum ct_op { CT_GE = 0, CT_LE = 1, CT_GT = 2, CT_LT = 3 };

/* Debug/parity stages for ct_run_contrack(): which intermediate array is painted into `flag`.
 * Values painted for stages 1-4 are (index of a representative row-run + 1): compare PARTITIONS with the oracle. */
enum ct_stage {
    CT_STAGE_FINAL = 0,     /* ds['flag'] of contrack.py:776-791, ids identical to the reference */
    CT_STAGE_LABEL2D = 1,   /* after contrack.py:684-687 (per-plane 8-connected components, no seam) */
    CT_STAGE_SEAM2D = 2,    /* after contrack.py:691-698 (same-row date-line merge) */
    CT_STAGE_FILTERED = 3,  /* after contrack.py:706-742 (overlap filter), class id or 0 */
    CT_STAGE_LABEL3D = 4    /* after contrack.py:748-751: the scipy 3-D label ids themselves (exact ids) */
};

int ct_version(void);
const char* ct_last_error(void);

/* One context per GPU; owns all scratch (bit planes, run tables, component tables). */
int ct_create(int device, ct_ctx** out);
void ct_destroy(ct_ctx* ctx);
/* Runtime switches (defaults in brackets).  Returns CT_ERR_ARG for unknown keys.  Every combination gives the same bytes
 * (tests/test_gpu_parity.py, test_gpu_fastpath.py); they select between the product path and the fallbacks it keeps.
 *   "plane_kernel" [2] who builds the tables: 1 = the plane kernel (one thread block per time plane: union-find over the
 *                  plane's row-runs in shared memory, look-back numbering, pair hash; one launch, no host round trip),
 *                  0 = the global-memory table kernels (every step parallel over all planes, pipelined in time chunks under
 *                  the threshold kernel), 2 = by size: the plane kernel up to "plane_max_planes" [4096] planes per context
 *                  (the shards of a multi-GPU run), the global-memory kernels for longer cubes.  A plane that does not fit the
 *                  plane kernel's shared memory ("plane_smem" [0 = 40 KB, then 200 KB]) falls back automatically.
 *   "max_sweeps"   [32] Jacobi sweeps of the overlap filter before the plane-ordered wavefront takes over (bounded cost on
 *                  adversarial keep/kill chains)
 *   "gpu_tables"   [1] ordered phase (contrack.py:706-751 + label boxes + date-line events) as ONE cooperative kernel and an
 *                  O(events) host replay; 0 = the whole ordered phase on the host (the reference's loops on the tables: what
 *                  near-ties on non-exact rows and labels that straddle a stale box fall back to)
 *   "tma"          [1] threshold kernel: rows staged by cp.async.bulk + mbarrier; 0 = plain coalesced loads (also used
 *                  automatically for rows that are not 16-byte aligned or too long for shared memory)
 *   "overlap_zero" [1] zero fill of the flag cube on a side stream beside the table phase + sparse paint; 0 = dense paint
 *   "fill_ctas"    [-1] resident blocks per SM of the zero fill (room for the table kernels beside it); -1 = automatic
 *                  (1 beside the plane kernel, 2 beside the global-memory table kernels), 0 = uncapped
 *   "fill_late"    [0] plane-kernel path: 1 = the zero fill starts after the plane kernel instead of beside it
 *   "p2p"          [1] sharded run: the pack kernel stores every rank's tables straight into the gathered buffers of all
 *                  ranks (peer memory over NVLink, mapped with CUDA IPC once per communicator) and signals with flag words;
 *                  0 = local pack + one ncclAllGather (also used when peer memory is unavailable or with > 16 ranks).
 *                  Must be the same on all ranks.
 *   "chunks"       [4] time chunks of the global-memory table pipeline ("chunk_min_planes" [1024] = smallest chunk);
 *                  "fast_chunks" [1] the same for the plane kernel (1: it runs beside the zero fill)
 *   "host_sparse"  [1] ct_run_contrack_host returns the result as row-runs expanded by host threads; 0 = dense copy
 *   "host_zero_threads" [0 = automatic], "host_threads" [0 = automatic] host threads of the host-buffer calls;
 *   "host_out_zeroed" [0] 1 = the caller guarantees that flag_host is all zero on entry (the zeroing pass is skipped)
 *   "profile_tables" [0] debug: CUDA-event time of every group of global-memory table kernels -> stats "ms_t_*" */
int ct_set_option(ct_ctx* ctx, const char* key, long value);

/* ---- run_contrack, contrack.py:646-772 -------------------------------------------------------------------------
 * anom       [T,H,W] float32/float64 on the device (ds[variable] transposed to time,lat,lon: contrack.py:677-681)
 * w_host     [H] float64 area weight per latitude row, computed by the caller with the reference's expression
 *            (contrack.py:703-704; float32 values widened to float64)
 * thr_host   threshold(s): thr_n == 1 (numeric branch, contrack.py:663-674) or thr_n == T (DataArray/dayofyear
 *            branch already gathered per time step, contrack.py:648-661)
 * thr_is_f32 1: compare in float32 against (float)thr (numpy weak-scalar rule for a Python number vs a float32
 *            array); 0: compare in float64
 * flag_dev   [T,H,W] int32 out.  n_features (may be NULL) receives contrack.py:793's count.
 */
int ct_run_contrack(ct_ctx* ctx, const void* anom_dev, int in_dtype, long T, int H, int W,
                    const double* w_host, const double* thr_host, long thr_n, int thr_is_f32, int op,
                    double overlap, int persistence, int twosided,
                    int32_t* flag_dev, long* n_features, int stage, void* stream);

/* Same call with HOST buffers: streams time chunks host->device (the threshold kernel consumes them; the float cube is
 * never resident) and runs the table phases.  PCIe is the bound of this call, so the result does not travel back as a
 * dense cube: host threads zero `flag_host` while the input streams in, the row-run table (x0|x1<<16, row, value: 12 B
 * per run, ~1 % of the dense bytes on Z500-like fields) is copied device->host and the threads expand the runs of the
 * surviving features into `flag_host` ("host_sparse" = 0, or a run table larger than half the cube, selects the dense
 * paint + copy instead).  `chunk_planes` <= 0 picks a default.  Host buffers should be page-locked for full PCIe rate
 * (the call works with pageable memory).  Stats "h2d_bytes" / "d2h_bytes" report what crossed PCIe. */
int ct_run_contrack_host(ct_ctx* ctx, const void* anom_host, int in_dtype, long T, int H, int W,
                         const double* w_host, const double* thr_host, long thr_n, int thr_is_f32, int op,
                         double overlap, int persistence, int twosided,
                         int32_t* flag_host, long* n_features, long chunk_planes);

/* ---- statistics of the last run on this context (for bench.py / tests) ------------------------------------------
 * keys: "runs", "comps2d", "pairs", "labels3d", "features", "seam_segments", "seam_events", "seam_splits", "sweeps",
 *       "wavefront_planes", "neartie_resolved", "kernel_launches", "fast_path" (1 = plane kernel + cooperative kernel,
 *       0 = global-memory table kernels, 0.5 / 0.25 = an ordered host replay took over), "plane_attempts",
 *       "ms_threshold", "ms_zero_fill", "ms_plane_kernel", "ms_global_kernel", "ms_host_tables", "ms_paint", "ms_total"
 *       (CUDA-event / host times of the last run in milliseconds)
 * returns the value, or -1 for an unknown key. */
double ct_get_stat(ct_ctx* ctx, const char* key);

/* ---- host-only table phase (no GPU needed): seam merge with the bounding boxes taken BEFORE merging
 * (contrack.py:753-763) followed by the persistence filter (contrack.py:765-772), on component tables.
 * comp_* arrays describe the kept 2-D components in first-pixel (t,y,x) order; comp_label[] is the scipy 3-D label
 * (1..N) of each.  seg_* are the date-line contact segments in (t,y) order: rows y0..y1-1 of plane t whose pixel at
 * x=0 belongs to component seg_a and whose pixel at x=W-1 belongs to component seg_b.
 * run_ptr/run_y/run_x0/run_x1 (CSR over components, x1 exclusive) may be NULL if no component needs splitting;
 * then a needed split returns CT_ERR_INTERNAL.
 * Outputs: comp_val[ncomp] final id per component (0 = removed; components that were split get 0 here and their
 * pieces are reported as overrides); ovr_* (capacity ovr_cap) receive (t,y,x0,x1,value) sub-runs; *n_ovr their count.
 */
int ct_track_tables(long T, int H, int W, int persistence,
                    long ncomp, const int32_t* comp_t, const int32_t* comp_y0, const int32_t* comp_y1,
                    const int32_t* comp_x0, const int32_t* comp_x1, const int32_t* comp_label,
                    long nseg, const int32_t* seg_t, const int32_t* seg_y0, const int32_t* seg_y1,
                    const int32_t* seg_a, const int32_t* seg_b,
                    const int64_t* run_ptr, const int32_t* run_y, const int32_t* run_x0, const int32_t* run_x1,
                    int32_t* comp_val, long ovr_cap, int32_t* ovr_t, int32_t* ovr_y, int32_t* ovr_x0,
                    int32_t* ovr_x1, int32_t* ovr_val, long* n_ovr, long* n_features, long* n_events, long* n_splits);

/* ---- host-only: the whole ordered table phase (no GPU needed) ------------------------------------------------------
 * Replays contrack.py:706-772 on the tables the CUDA kernels produce: the time-sequential overlap filter (706-742), the
 * 3-D labelling of the kept mask with scipy's first-pixel numbering (747-751), the date-line merge through stale boxes
 * and the persistence filter (753-772).  Exposed so the ordered logic can be checked against the oracle without a GPU.
 *   comp_*    2-D components (8-connected, no wrap) in global first-pixel order; comp_cls = smallest id of the same-row
 *             date-line class (contrack.py:691-698); areaE / areaS / nsp: see ct_classify_rows
 *   pair_*    (a at plane t, b at plane t-1) with common pixels: pixel count, areas, special-row pixel count
 *   seam_*    rows (row = t*H + y, sorted) whose pixels at x=0 (component a) and x=W-1 (component b) are both set
 *   plane_run_ptr/run_*  optional (may be NULL): all row-runs, raster order, CSR over planes [T+1]; needed only for
 *             near-tie decisions on special rows and for stale-box splits
 *   stats8    {features, kept 2-D comps, 3-D labels, seam events, seam splits, near-ties resolved, 0, 0}
 */
int ct_host_tables(long T, int H, int W, const double* w_host, double overlap, int persistence, int twosided, int stage,
                   long ncomp, const int32_t* comp_t, const int32_t* comp_y0, const int32_t* comp_y1,
                   const int32_t* comp_x0, const int32_t* comp_x1, const uint32_t* comp_cls, const double* comp_areaE,
                   const double* comp_areaS, const uint32_t* comp_nsp,
                   long npair, const uint32_t* pair_a, const uint32_t* pair_b, const uint32_t* pair_npix,
                   const uint32_t* pair_nsp, const double* pair_areaE, const double* pair_areaS,
                   long nseam, const uint32_t* seam_row, const uint32_t* seam_a, const uint32_t* seam_b,
                   const int64_t* plane_run_ptr, const int32_t* run_y, const int32_t* run_x0, const int32_t* run_x1,
                   const uint32_t* run_comp,
                   int32_t* comp_val, long ovr_cap, int32_t* ovr_t, int32_t* ovr_y, int32_t* ovr_x0, int32_t* ovr_x1,
                   int32_t* ovr_val, long* n_ovr, long* stats8);

/* The same ordered phase on tables that are already in the sequential layout the kernels deliver (class sums at the
 * representative's index, pairs in CSR form over their plane-t component, date-line rows grouped into segments).  This is
 * what a time-sharded run calls after gathering every rank's tables.  `fetch` (may be NULL) supplies the row-runs of one
 * plane when a near-tie decision or a stale-box split needs them: it sets *n and four pointers that must stay valid until
 * the next call; component ids in `comp` are the global ones; return 0 on success. */
typedef int (*ct_plane_runs_fn)(void* user, long t, long* n, const int32_t** y, const int32_t** x0, const int32_t** x1,
                                const uint32_t** comp);
int ct_host_tables_fast(long T, int H, int W, const double* w_host, double overlap, int persistence, int twosided,
                        int stage, long ncomp, const int32_t* comp_t, const int32_t* comp_y0, const int32_t* comp_y1,
                        const int32_t* comp_x0, const int32_t* comp_x1, const uint32_t* comp_cls,
                        const double* cls_conE, const double* cls_conS, const double* cls_fE, const double* cls_fS,
                        const uint32_t* cls_nsp, const uint32_t* pair_ptr, const uint32_t* pair_b,
                        const uint32_t* pair_npix, const uint32_t* pair_nsp, const double* pair_E, const double* pair_S,
                        long nseg, const int32_t* seg_t, const int32_t* seg_y0, const int32_t* seg_y1,
                        const uint32_t* seg_a, const uint32_t* seg_b, ct_plane_runs_fn fetch, void* user,
                        int32_t* comp_val, long ovr_cap, int32_t* ovr_t, int32_t* ovr_y, int32_t* ovr_x0, int32_t* ovr_x1,
                        int32_t* ovr_val, long* n_ovr, long* stats8);

/* ---- time-sharded run: one context and one communicator per rank / GPU -------------------------------------
 * Rank r of nranks owns planes [t_begin, t_begin + T_local) of the cube; ranks are ordered in time.  ct_run_contrack_sharded
 * is ONE collective call per rank (contrack.py:646-772 across the shards): the last own plane's bit rows go to the next rank
 * (the one halo exchange, send/recv of H * ceil(W/32) words), the rank tables (a few MB, slots of fixed stride) reach every
 * rank in one step -- stored straight into the peers' buffers over NVLink by the kernel that packs them (option "p2p"; CUDA
 * IPC windows set up once per communicator), or with one ncclAllGather -- are merged by a kernel into global tables on every
 * rank, and every rank replays the global part (overlap
 * filter, 3-D numbering, stale-box date-line merge, persistence) on its copy: all ranks obtain the same global ids ("global
 * relabel") and paint their own planes.  The cube never moves.  ids, dtype and semantics are those of ct_run_contrack on the
 * concatenated cube, bit for bit.
 *
 * Communicator: NCCL, bound at run time (dlopen of the libnccl the process already uses, else the system one).
 *   ct_nccl_unique_id    rank 0 makes the 128-byte ncclUniqueId; the caller distributes it (MPI_Bcast, a file, torch.distributed)
 *   ct_comm_init_nccl    every rank joins (ncclCommInitRank on `device`)
 *   ct_comm_from_nccl    ... or wraps a ncclComm_t the caller already owns (not destroyed by ct_comm_destroy)
 *   ct_comm_init_local   nranks communicators of an in-process group (one host thread per rank, contexts on one or several GPUs
 *                        of this process): the same stream-ordered semantics with events and peer copies -- for tests on a
 *                        single-GPU box, not a product transport
 * thr_host: one value, or the T_local values of this rank's planes.  `stream`: the caller's stream; the call returns after this
 * rank's flag planes are written (it synchronises once for the date-line events, like ct_run_contrack).
 * Runtime options of the context apply ("plane_kernel", "max_sweeps", "p2p", ...) and must be equal on all ranks; stats as for
 * ct_run_contrack plus "exchange_bytes", "shard_attempts", "ms_exchange", "p2p" (1 = peer windows were used).
 * The negotiated slot sizes and the peer windows belong to the communicator: the ranks of a communicator must make the same
 * sequence of sharded calls (as with any collective).  A rank that enters the call late is waited for; a rank that never
 * arrives ends the others' wait with CT_ERR_INTERNAL after 10 minutes (peer windows) or blocks like any NCCL collective. */
int ct_nccl_unique_id(unsigned char id[128]);
int ct_comm_init_nccl(const unsigned char id[128], int rank, int nranks, int device, ct_comm** out);
int ct_comm_from_nccl(void* nccl_comm, int rank, int nranks, ct_comm** out);
int ct_comm_init_local(int nranks, ct_comm** out /* [nranks] */);
void ct_comm_destroy(ct_comm* comm);
int ct_comm_rank(ct_comm* comm);
int ct_comm_size(ct_comm* comm);
int ct_run_contrack_sharded(ct_ctx* ctx, ct_comm* comm, const void* anom_dev, int in_dtype, long T_local, long t_begin,
                            long T_total, int H, int W, const double* w_host, const double* thr_host, long thr_n,
                            int thr_is_f32, int op, double overlap, int persistence, int twosided, int32_t* flag_dev,
                            long* n_features, void* stream);
/* The same collective with HOST buffers (this rank's planes in anom_host, its flag planes out in flag_host; page-locked memory
 * for full PCIe rate): the shard streams host -> device in time chunks under the threshold kernel, the result leaves as the
 * row-run table (12 B per run) and host threads expand the surviving runs into flag_host, which they zero while the input is
 * streaming in -- see ct_run_contrack_host.  With N ranks on one host the N PCIe links work in parallel; host memory bandwidth
 * (reading the shards, zeroing the results) is the shared bound.  Stats "h2d_bytes" / "d2h_bytes" as for ct_run_contrack_host. */
int ct_run_contrack_sharded_host(ct_ctx* ctx, ct_comm* comm, const void* anom_host, int in_dtype, long T_local, long t_begin,
                                 long T_total, int H, int W, const double* w_host, const double* thr_host, long thr_n,
                                 int thr_is_f32, int op, double overlap, int persistence, int twosided, int32_t* flag_host,
                                 long* n_features, long chunk_planes);

/* ---- run_lifecycle, contrack.py:799-907 -------------------------------------------------------------------------------
 * flag_dev [T,H,W] int32 (ds[flag]), var_dev [T,H,W] float32/float64 (ds[variable]), w_host [H] the float32-valued area
 * weights of contrack.py:847-848.  One result row per (time step, flag id != 0) that occurs in the cube, in no particular
 * order (the caller sorts like contrack.py:907).  Every sum is accumulated in the order the reference's numpy / scipy calls
 * use, so the float64 values are the reference's bit for bit:
 *   area  = np.sum(weight_grid[flag[t] == id])                       (contrack.py:874, numpy pairwise order)
 *   wsum  = np.sum(weight_grid[m] * variable[t][m])                   (contrack.py:875; intensity = wsum / area)
 *   norm, sy, sx = the three sums of scipy.ndimage.center_of_mass(variable * weight_grid, flag, [id]) (contrack.py:886,
 *           892): sequential np.bincount order over the plane rolled by -roll columns; centre = (sy / norm, sx / norm)
 *   roll  = lon_roll of contrack.py:882-883 if the id touches both the first and the last longitude column, else -1
 * Results stay in the context until the next call; ct_lifecycle_fetch copies them out (cap >= *n_rows). */
int ct_run_lifecycle(ct_ctx* ctx, const int32_t* flag_dev, const void* var_dev, int var_dtype, long T, int H, int W,
                     const double* w_host, long* n_rows, void* stream);
int ct_lifecycle_fetch(ct_ctx* ctx, long cap, int32_t* t, int32_t* label, int32_t* npix, int32_t* roll, double* area,
                       double* wsum, double* norm, double* sy, double* sx);

/* special_out[y] = 0 if row y belongs to the set of rows whose weights sum exactly in float64 in any order (areas of
 * those rows accumulate in areaE and equal numpy's np.sum bit for bit), 1 otherwise (pole rows: areaS, nsp). */
void ct_classify_rows(const double* w_host, int H, int W, uint8_t* special_out);

/* numpy's float64 pairwise summation (the order np.sum uses on a contiguous 1-D array) over a run-length encoded
 * sequence: value[i] repeated count[i] times.  Used by the near-tie resolver; exported for the parity test. */
double ct_numpy_pairwise_sum_rle(const double* value, const int64_t* count, long n);

/* ---- calc_clim / calc_anom, contrack.py:458-581 (float32, tolerance parity) --------------------------------------
 * z_dev [T,H,W] float32; group_host [T] int32 group index 0..G-1 of each time step (dayofyear rank);
 * clim_dev [G,H,W] float32 out: group mean (skip-NaN) -> centred rolling mean (window, min_periods = window) ->
 * NaN edges filled with the mean of the last `window` unsmoothed group means (contrack.py:483-489). */
int ct_calc_clim(ct_ctx* ctx, const float* z_dev, long T, int H, int W, const int32_t* group_host, int G,
                 int window, float* clim_dev, void* stream);
/* anom_dev [T,H,W] float32 out = centred rolling mean over time (window `smooth`) of z[t] - clim[group[t]]
 * (contrack.py:568-570). */
int ct_calc_anom(ct_ctx* ctx, const float* z_dev, long T, int H, int W, const int32_t* group_host, int G,
                 const float* clim_dev, int smooth, float* anom_dev, void* stream);

/* The same two calls for float32 OR float64 cubes (dtype = ct_dtype of z, the climatology and the anomaly alike): xarray
 * keeps the precision of its input -- a float64 `z` (packed ERA5 decoded with scale/offset) gives a float64 climatology and
 * anomaly, every mean accumulated in float64 (contrack.py:483-489, 568-570). */
int ct_calc_clim_t(ct_ctx* ctx, const void* z_dev, int dtype, long T, int H, int W, const int32_t* group_host, int G,
                   int window, void* clim_dev, void* stream);
int ct_calc_anom_t(ct_ctx* ctx, const void* z_dev, int dtype, long T, int H, int W, const int32_t* group_host, int G,
                   const void* clim_dev, int smooth, void* anom_dev, void* stream);

/* ---- callers either side of the path -------------------------------------------------------------------
 * ct_quantile_time   README.rst:150-151: ds[var].sel(latitude=band).quantile(q, dim='time') -- numpy's nanquantile with the
 *                    'linear' method along time for every grid point of rows [y0, y1): out_dev [nq, y1-y0, W] float64, bit
 *                    for bit what numpy returns for a float32 input and float64 q (exact order statistics by radix select;
 *                    the interpolation in numpy's operation order).  NaN are skipped; all-NaN points give NaN.
 * ct_flag_count      README.rst:161: xr.where(flag > greater_than, 1, 0).sum(dim='time') -> count_dev [H, W] int32
 * ct_divide_f32      contrack.py:417-419 (calculate_gph_from_gp): out = in / divisor in float32 (IEEE division)
 * ct_gather_planes   contrack.py:565 (clim.reindex(lat, lon, method='nearest')): dst[g, y, x] = src[g, iy[y], ix[x]]; the
 *                    nearest-neighbour index maps come from the caller (pandas Index.get_indexer, as xarray does) */
int ct_quantile_time(ct_ctx* ctx, const float* x_dev, long T, int H, int W, int y0, int y1, const double* q_host, int nq,
                     double* out_dev, void* stream);
/* ... for float32 or float64 cubes (dtype = ct_dtype; float64 in -> numpy's float64 arithmetic throughout) and, when `comm` is
 * given, over ALL time steps of a time-sharded cube: x_dev holds this rank's T_local steps, the per-point tallies of the
 * radix select are summed over the ranks between the passes (all-reduce), every rank receives the quantiles of the whole cube,
 * bit-identical to np.nanquantile on the gathered cube.  Collective over `comm` (NULL: this rank alone). */
int ct_quantile_time_t(ct_ctx* ctx, ct_comm* comm, const void* x_dev, int dtype, long T_local, int H, int W, int y0, int y1,
                       const double* q_host, int nq, double* out_dev, void* stream);
int ct_flag_count(ct_ctx* ctx, const int32_t* flag_dev, long T, int H, int W, int greater_than, int32_t* count_dev,
                  void* stream);
int ct_divide_f32(ct_ctx* ctx, const float* in_dev, size_t n, float divisor, float* out_dev, void* stream);
int ct_gather_planes(ct_ctx* ctx, const float* src_dev, int G, int Hs, int Ws, const int32_t* iy_host, const int32_t* ix_host,
                     int H, int W, float* dst_dev, void* stream);
/* ... for float32 or float64 planes (dtype = ct_dtype) */
int ct_gather_planes_t(ct_ctx* ctx, const void* src_dev, int dtype, int G, int Hs, int Ws, const int32_t* iy_host,
                       const int32_t* ix_host, int H, int W, void* dst_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CONTRACK_B200_H */
