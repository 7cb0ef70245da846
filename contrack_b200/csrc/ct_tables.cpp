// ct_tables.cpp -- see ct_tables.h.  Pure host C++ (no CUDA), exact integer logic.
#include "ct_tables.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <unordered_map>

namespace ctb {

namespace {

struct Box3 { int t0, t1, y0, y1, x0, x1; };      // half-open on every axis

struct Piece {
    long comp;                  // kept 2-D component this piece belongs to
    int t, y0, y1, x0, x1;      // plane and tight 2-D bounding box (half-open)
    int value;                  // current id (the reference's flag value of these pixels)
    bool has_runs;              // runs[] materialised
    std::vector<SubRun> runs;
};

enum Rel { OUTSIDE = 0, INSIDE = 1, PARTIAL = 2 };

inline Rel classify(const Piece& p, const Box3& b) {
    if (p.t < b.t0 || p.t >= b.t1) return OUTSIDE;
    if (p.y0 >= b.y0 && p.y1 <= b.y1 && p.x0 >= b.x0 && p.x1 <= b.x1) return INSIDE;
    if (p.y1 <= b.y0 || p.y0 >= b.y1 || p.x1 <= b.x0 || p.x0 >= b.x1) return OUTSIDE;
    return PARTIAL;
}

inline void tight_box(Piece& p) {
    int y0 = INT_MAX, y1 = 0, x0 = INT_MAX, x1 = 0;
    for (const SubRun& r : p.runs) {
        y0 = std::min(y0, r.y); y1 = std::max(y1, r.y + 1);
        x0 = std::min(x0, r.x0); x1 = std::max(x1, r.x1);
    }
    p.y0 = y0; p.y1 = y1; p.x0 = x0; p.x1 = x1;
}

}  // namespace

int track_tables(long T, int H, int W, int persistence,
                 long ncomp, const int32_t* comp_t, const int32_t* comp_y0, const int32_t* comp_y1,
                 const int32_t* comp_x0, const int32_t* comp_x1, const int32_t* comp_label,
                 long nseg, const int32_t* seg_t, const int32_t* seg_y0, const int32_t* seg_y1,
                 const int32_t* seg_a, const int32_t* seg_b,
                 RunFetcher* fetcher, int32_t* comp_val, std::vector<Override>& overrides, TrackStats& stats) {
    (void)H; (void)T; (void)seg_t;
    overrides.clear();
    stats = TrackStats();

    // Buffers live across calls (one set per host thread): a fresh multi-megabyte vector per call costs more in page
    // faults than the work done on it.
    struct Workspace {
        std::vector<int> value, tmin, tmax, fin, head;
        std::vector<long> nxt;
        std::vector<Box3> box;
    };
    static thread_local Workspace tls_ws;
    Workspace& ws = tls_ws;                      // one TLS lookup, not one per access
    std::vector<int>&value = ws.value, &tmin = ws.tmin, &tmax = ws.tmax, &fin = ws.fin, &head = ws.head;
    std::vector<long>& nxt = ws.nxt;
    std::vector<Box3>& box = ws.box;

    int nlabel = 0;
    for (long c = 0; c < ncomp; ++c) nlabel = std::max(nlabel, (int)comp_label[c]);
    // value of each whole component (0 = removed before this stage); pieces beyond ncomp only exist after a split
    value.assign(comp_label, comp_label + ncomp);
    std::vector<Piece> extra;                               // pieces created by splits (index ncomp + k)
    std::vector<Piece> whole_runs;                          // whole components whose runs were materialised
    std::unordered_map<long, int> whole_runs_idx;           // comp -> index in whole_runs
    std::unordered_map<long, std::vector<long>> comp_pieces; // split comps -> piece ids (comp id itself = remainder)

    auto make_piece = [&](long c) {
        Piece p;
        p.comp = c; p.t = comp_t[c]; p.y0 = comp_y0[c]; p.y1 = comp_y1[c]; p.x0 = comp_x0[c]; p.x1 = comp_x1[c];
        p.value = value[c]; p.has_runs = false;
        return p;
    };

    // 3-D boxes of the ORIGINAL labels (find_objects before merging, contrack.py:753); their t-extent is also the
    // persistence extent of every label no date-line event touches
    box.assign(nlabel + 1, Box3{INT_MAX, 0, INT_MAX, 0, INT_MAX, 0});
    for (long c = 0; c < ncomp; ++c) {
        const int v = comp_label[c];
        if (v == 0) continue;
        Box3& b = box[v];
        b.t0 = std::min(b.t0, (int)comp_t[c]); b.t1 = std::max(b.t1, (int)comp_t[c] + 1);
        b.y0 = std::min(b.y0, (int)comp_y0[c]); b.y1 = std::max(b.y1, (int)comp_y1[c]);
        b.x0 = std::min(b.x0, (int)comp_x0[c]); b.x1 = std::max(b.x1, (int)comp_x1[c]);
    }
    tmin.resize(nlabel + 1); tmax.resize(nlabel + 1);
    for (int v = 0; v <= nlabel; ++v) { tmin[v] = box[v].t0; tmax[v] = box[v].t1 - 1; }

    // does any date-line row join two different labels at all?  (usually only a few per cent of the segments do)
    bool any_event = false;
    for (long s = 0; s < nseg && !any_event; ++s) {
        const int la = comp_label[seg_a[s]], lb = comp_label[seg_b[s]];
        any_event = la != 0 && lb != 0 && la != lb;
    }

    if (any_event) {
        // member lists of the labels: intrusive singly linked lists (head per value, next per piece); moving a piece from
        // one value to another is an unlink + push, no allocation.  Order inside a list does not matter: every member is
        // classified against the same stale box independently.
        head.assign(nlabel + 1, -1);
        nxt.assign(ncomp, -1);
        for (long c = ncomp - 1; c >= 0; --c) {
            const int v = comp_label[c];
            if (v == 0) continue;
            nxt[c] = head[v]; head[v] = (int)c;
        }
        auto piece_ref = [&](long pid) -> Piece* {           // only valid until `extra` / `whole_runs` grow
            if (pid >= ncomp) return &extra[pid - ncomp];
            auto it = whole_runs_idx.find(pid);
            return it == whole_runs_idx.end() ? nullptr : &whole_runs[it->second];
        };
        auto piece_value = [&](long pid) -> int { return pid >= ncomp ? extra[pid - ncomp].value : value[pid]; };
        auto piece_of = [&](long c, int y, int x) -> long {  // piece holding pixel (y, x) of component c
            if (comp_pieces.empty()) return c;
            auto it = comp_pieces.find(c);
            if (it == comp_pieces.end()) return c;
            for (long pid : it->second) {
                const Piece* p = piece_ref(pid);
                for (const SubRun& r : p->runs)
                    if (r.y == y && r.x0 <= x && x < r.x1) return pid;
            }
            return c;                                         // unreachable for consistent tables
        };

        int rc = 0;
        std::vector<long> moved;                              // heads are ints; extra pieces may exceed int only in theory
        // after an event on `hi` all of its remaining members lie outside box[hi]: a repeat cannot move anything until new
        // members arrive
        std::vector<uint8_t> settled(nlabel + 1, 0);
        auto do_event = [&](int hi, int lo) {
            if (settled[hi]) return;
            const Box3 b = box[hi];
            long pid = head[hi];
            long prev = -1;
            moved.clear();
            while (pid >= 0) {
                const long next = nxt[pid];
                Piece tmp;
                Piece* pp = piece_ref(pid);
                if (!pp) { tmp = make_piece(pid); pp = &tmp; }
                Rel rel = classify(*pp, b);
                bool move_whole = false;
                if (rel == PARTIAL) {
                    // materialise the row-runs of this piece
                    if (!pp->has_runs) {
                        Piece np_ = make_piece(pid);
                        if (!fetcher || !fetcher->fetch(pid, np_.runs)) { rc = -1; prev = pid; pid = next; continue; }
                        np_.has_runs = true;
                        whole_runs_idx[pid] = (int)whole_runs.size();
                        whole_runs.push_back(std::move(np_));
                        pp = &whole_runs.back();
                    }
                    std::vector<SubRun> in, out;
                    for (const SubRun& r : pp->runs) {
                        if (r.y < b.y0 || r.y >= b.y1 || r.x1 <= b.x0 || r.x0 >= b.x1) { out.push_back(r); continue; }
                        int a = std::max(r.x0, b.x0), e = std::min(r.x1, b.x1);
                        if (r.x0 < a) out.push_back(SubRun{r.y, r.x0, a});
                        in.push_back(SubRun{r.y, a, e});
                        if (e < r.x1) out.push_back(SubRun{r.y, e, r.x1});
                    }
                    if (in.empty()) rel = OUTSIDE;
                    else if (out.empty()) rel = INSIDE;
                    else {
                        // split: the inside part becomes a new piece with value lo, the rest keeps hi
                        Piece q;
                        q.comp = pp->comp; q.t = pp->t; q.value = lo; q.has_runs = true; q.runs.swap(in);
                        tight_box(q);
                        pp->runs.swap(out);
                        tight_box(*pp);
                        const long comp = pp->comp;
                        const long qid = ncomp + (long)extra.size();
                        extra.push_back(std::move(q));        // may invalidate pp
                        nxt.push_back(-1);
                        std::vector<long>& cp = comp_pieces[comp];
                        if (cp.empty()) cp.push_back(comp);   // the whole-comp id now denotes the remainder
                        cp.push_back(qid);
                        moved.push_back(qid);
                        stats.n_splits++;
                        prev = pid; pid = next;
                        continue;
                    }
                }
                if (rel == INSIDE) move_whole = true;
                if (move_whole) {
                    if (pid >= ncomp) extra[pid - ncomp].value = lo; else value[pid] = lo;
                    if (prev < 0) head[hi] = (int)next; else nxt[prev] = next;      // unlink
                    moved.push_back(pid);
                } else {
                    prev = pid;
                }
                pid = next;
            }
            for (long m : moved) { nxt[m] = head[lo]; head[lo] = (int)m; }
            settled[hi] = 1;
            if (!moved.empty()) settled[lo] = 0;
            // the t-extents of both values are stale now: recomputed after all events
            tmax[hi] = -2; tmax[lo] = -2;
        };

        for (long s = 0; s < nseg && rc == 0; ++s) {
            const long a = seg_a[s], b = seg_b[s];
            if (comp_label[a] == 0 || comp_label[b] == 0) continue;          // a removed component: pixel value 0
            for (int y = seg_y0[s]; y < seg_y1[s]; ++y) {
                long pa = piece_of(a, y, 0), pb = piece_of(b, y, W - 1);
                int va = piece_value(pa), vb = piece_value(pb);
                if (va != vb) {
                    do_event(std::max(va, vb), std::min(va, vb));
                    stats.n_events++;
                }
                // With both components unsplit every later row of the segment sees the same two pieces and no other
                // event intervenes, so re-applying the (idempotent) merge changes nothing.
                if (comp_pieces.empty() ||
                    (comp_pieces.find(a) == comp_pieces.end() && comp_pieces.find(b) == comp_pieces.end())) break;
            }
        }
        if (rc != 0) return rc;

        // t-extent of the values touched by events, from their current members
        for (int v = 1; v <= nlabel; ++v) {
            if (tmax[v] != -2) continue;
            int lo = INT_MAX, hi = -1;
            for (long pid = head[v]; pid >= 0; pid = nxt[pid]) {
                const int t = pid >= ncomp ? extra[pid - ncomp].t : comp_t[pid];
                lo = std::min(lo, t); hi = std::max(hi, t);
            }
            tmin[v] = lo; tmax[v] = hi;
        }
    }

    // persistence on the merged values (contrack.py:765-772): t-extent of all pixels that carry value v
    fin.assign(nlabel + 1, 0);
    for (int v = 1; v <= nlabel; ++v) {
        if (tmax[v] < 0) continue;
        if ((tmax[v] + 1 - tmin[v]) < persistence) continue;
        fin[v] = v;
        stats.n_features++;
    }
    for (long c = 0; c < ncomp; ++c) comp_val[c] = fin[value[c]];
    for (auto& kv : comp_pieces) {
        comp_val[kv.first] = 0;
        for (long pid : kv.second) {
            const Piece* p = pid >= ncomp ? &extra[pid - ncomp] : &whole_runs[whole_runs_idx[pid]];
            int v = fin[pid >= ncomp ? p->value : value[pid]];
            if (v == 0) continue;
            for (const SubRun& r : p->runs) overrides.push_back(Override{p->t, r.y, r.x0, r.x1, v});
        }
    }
    return 0;
}

int track_events_fast(int persistence, long nev, const int32_t* ev, long nrec, const int32_t* lrec,
                      std::vector<int32_t>& patch_label, std::vector<int32_t>& patch_value, long* feat_delta,
                      TrackStats& stats) {
    stats = TrackStats();
    patch_label.clear(); patch_value.clear();
    *feat_delta = 0;
    if (nev == 0) return 0;
    // the records are sorted by label, so the order of record indices is the order of the label values
    struct Workspace { std::vector<int> value, head, nxt, touched; std::vector<uint8_t> built, settled; };
    static thread_local Workspace tls_ws;
    Workspace& ws = tls_ws;
    const int m = (int)nrec;
    auto box = [&](int L) { const int32_t* r = lrec + 7 * (size_t)L; return Box3{r[1], r[2], r[3], r[4], r[5], r[6]}; };
    std::vector<int>&value = ws.value, &head = ws.head, &nxt = ws.nxt, &touched = ws.touched;
    std::vector<uint8_t>&built = ws.built, &settled = ws.settled;
    value.resize(m); head.resize(m); nxt.resize(m);
    built.assign(m, 0); settled.assign(m, 0);
    touched.clear();
    for (int L = 0; L < m; ++L) value[L] = L;             // every label starts as its own value
    auto ensure_built = [&](int v) {                      // class v starts with label v alone
        if (built[v]) return;
        built[v] = 1; head[v] = v; nxt[v] = -1;
        touched.push_back(v);
    };
    const int* val = value.data();
    for (long e = 0; e < nev; ++e) {
        // most segments repeat a merge that has already happened (a contour crosses the date line for many rows and days):
        // two loads and a compare
        const uint32_t la = (uint32_t)ev[2 * e], lb = (uint32_t)ev[2 * e + 1];
        if (la >= (uint32_t)m || lb >= (uint32_t)m) return 1;
        const int va = val[la], vb = val[lb];
        if (va == vb) continue;
        stats.n_events++;
        const int hi = std::max(va, vb), lo = std::min(va, vb);
        ensure_built(hi); ensure_built(lo);
        if (settled[hi]) continue;
        const Box3 b = box(hi);
        int L = head[hi], prev = -1;
        bool moved = false;
        while (L >= 0) {
            const int next = nxt[L];
            stats.n_walked++;
            const Box3 q = box(L);
            const bool inside = q.t0 >= b.t0 && q.t1 <= b.t1 && q.y0 >= b.y0 && q.y1 <= b.y1 && q.x0 >= b.x0 && q.x1 <= b.x1;
            const bool outside = q.t1 <= b.t0 || q.t0 >= b.t1 || q.y1 <= b.y0 || q.y0 >= b.y1 || q.x1 <= b.x0 || q.x0 >= b.x1;
            if (inside) {
                if (prev < 0) head[hi] = next; else nxt[prev] = next;
                nxt[L] = head[lo]; head[lo] = L; value[L] = lo;
                moved = true;
            } else if (outside) {
                prev = L;
            } else {
                return 1;                                 // members of L may be on both sides of the box
            }
            L = next;
        }
        settled[hi] = 1;
        if (moved) settled[lo] = 0;
    }
    // persistence (contrack.py:765-772) of the touched values from their current members; the device counted every label
    // as a feature of its own box
    for (int v : touched) {
        const Box3 o = box(v);
        if (o.t1 > o.t0 && (o.t1 - o.t0) >= persistence) --*feat_delta;
    }
    for (int v : touched) {
        int lo = INT_MAX, hi = 0;
        for (int L = head[v]; L >= 0; L = nxt[L]) { const Box3 q = box(L); lo = std::min(lo, q.t0); hi = std::max(hi, q.t1); }
        const bool keep = hi > lo && (hi - lo) >= persistence;
        *feat_delta += keep;
        for (int L = head[v]; L >= 0; L = nxt[L]) {
            patch_label.push_back(lrec[7 * (size_t)L]); patch_value.push_back(keep ? lrec[7 * (size_t)v] : 0);
        }
    }
    return 0;
}

// ---- numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src: @TYPE@_pairwise_sum, PW_BLOCKSIZE 128) ----
static double pairwise(const double* a, long n) {
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        long i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise(a, n2) + pairwise(a + n2, n - n2);
    }
}

double numpy_pairwise_sum(const double* a, long n) { return 0. + pairwise(a, n); }

}  // namespace ctb
