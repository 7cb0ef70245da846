// ct_host.cpp -- see ct_host.h.  Pure host C++; compiled with -ffp-contract=off: the float64 decisions of step 3 must
// round exactly like the reference's numpy scalar arithmetic (contrack/contrack.py:721-722).
#include "ct_host.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace cth {

namespace {

// contrack.py:721-742 -- reciprocal, then multiply (two roundings), then the keep/kill cascade.
inline bool kill_decision(double areacon, double fwd, double bwd, double ov, bool twosided, double* fb_out,
                          double* ff_out) {
    const double inv = 1.0 / areacon;      // three separately rounded operations (-ffp-contract=off, no fast-math)
    const double fb = inv * bwd;
    const double ff = inv * fwd;
    *fb_out = fb; *ff_out = ff;
    bool kill = false;
    if (twosided) {
        if (fb != 0 && ff != 0) { if ((fb < ov) || (ff < ov)) kill = true; }
        if (fb != 0 && ff == 0) { if (fb < ov) kill = true; }
        if (fb == 0 && ff != 0) { if (ff < ov) kill = true; }
    } else {
        if (ff < ov) kill = true;
    }
    return kill;
}

// One fetched plane: runs in raster order + row offsets.
struct Plane {
    long t = -1;
    std::vector<PlaneRun> runs;
    std::vector<int> row_ptr;       // [H+1]
};

class PlaneCache {
public:
    PlaneCache(RunSource* src, int H) : src_(src), H_(H) {}
    const Plane* get(long t) {
        for (auto& p : slots_) if (p && p->t == t) return p.get();
        if (!src_) return nullptr;
        std::unique_ptr<Plane> p(new Plane);
        p->t = t;
        if (!src_->plane_runs(t, p->runs)) return nullptr;
        p->row_ptr.assign(H_ + 1, 0);
        for (const PlaneRun& r : p->runs) p->row_ptr[r.y + 1]++;
        for (int y = 0; y < H_; ++y) p->row_ptr[y + 1] += p->row_ptr[y];
        slots_[next_] = std::move(p);
        const Plane* ret = slots_[next_].get();
        next_ = (next_ + 1) % 4;
        return ret;
    }
private:
    RunSource* src_;
    int H_;
    std::unique_ptr<Plane> slots_[4];
    int next_ = 0;
};

// numpy-order sums for one date-line class at plane t (contrack.py:717-719): np.sum over the boolean-gathered weights in
// raster order.  `kept` holds the already-final keep state of plane t-1.
// Cold path (near-tie decisions only): kept out of line.
__attribute__((noinline)) bool exact_class_sums(PlaneCache& pc, const uint32_t* comp_cls, const double* w, const std::vector<uint8_t>& kept,
                      long t, uint32_t rep, double* areacon, double* fwd, double* bwd) {
    const Plane* cur = pc.get(t);
    const Plane* nxt = pc.get(t + 1);
    const Plane* prv = pc.get(t - 1);
    if (!cur || !nxt || !prv) return false;
    std::vector<double> con, f, b;
    for (const PlaneRun& r : cur->runs) {
        if (comp_cls[r.comp] != rep) continue;
        const double wy = w[r.y];
        con.insert(con.end(), (size_t)(r.x1 - r.x0), wy);
        for (int i = nxt->row_ptr[r.y]; i < nxt->row_ptr[r.y + 1]; ++i) {
            const PlaneRun& q = nxt->runs[i];
            int a = std::max(r.x0, q.x0), e = std::min(r.x1, q.x1);
            if (e > a) f.insert(f.end(), (size_t)(e - a), wy);
        }
        for (int i = prv->row_ptr[r.y]; i < prv->row_ptr[r.y + 1]; ++i) {
            const PlaneRun& q = prv->runs[i];
            if (!kept[q.comp]) continue;
            int a = std::max(r.x0, q.x0), e = std::min(r.x1, q.x1);
            if (e > a) b.insert(b.end(), (size_t)(e - a), wy);
        }
    }
    *areacon = ctb::numpy_pairwise_sum(con.data(), (long)con.size());
    *fwd = ctb::numpy_pairwise_sum(f.data(), (long)f.size());
    *bwd = ctb::numpy_pairwise_sum(b.data(), (long)b.size());
    return true;
}

struct SplitFetcher : ctb::RunFetcher {
    PlaneCache* pc; const int32_t* comp_t;
    bool fetch(long c, std::vector<ctb::SubRun>& out) override {
        const Plane* p = pc->get(comp_t[c]);
        if (!p) return false;
        out.clear();
        for (const PlaneRun& r : p->runs)
            if ((long)r.comp == c) out.push_back(ctb::SubRun{r.y, r.x0, r.x1});
        return !out.empty();
    }
};

}  // namespace

int host_phase_fast(const FastTables& tb, const Params& pr, RunSource* runs, int32_t* comp_val, Result& out,
                    std::string& err) {
    const long T = tb.T, nc = tb.ncomp;
    out.overrides.clear();
    out.n_features = out.n_kept = out.n_labels3d = out.n_seam_events = out.n_seam_splits = out.n_neartie = 0;
    if (pr.stage == 1) { for (long c = 0; c < nc; ++c) comp_val[c] = (int32_t)(c + 1); return 0; }
    if (pr.stage == 2) { for (long c = 0; c < nc; ++c) comp_val[c] = (int32_t)(tb.comp_cls[c] + 1); return 0; }

    PlaneCache pc(runs, tb.H);
    const bool timing = getenv("CT_HOST_TIMING") != nullptr;
    auto tnow = [] {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    double tm0 = tnow();
    auto lap = [&](const char* what) {
        if (timing) { double t = tnow(); fprintf(stderr, "  host_phase %-10s %.3f ms\n", what, t - tm0); tm0 = t; }
    };

    // buffers live across calls (one set per host thread): see ct_tables.cpp
    struct Workspace {
        std::vector<uint8_t> kept, verdict;
        std::vector<double> bE, bS;
        std::vector<uint32_t> bn, parent;
        std::vector<int32_t> label, seg_a32, seg_b32;
    };
    static thread_local Workspace tls_ws;
    Workspace& ws = tls_ws;                      // one TLS lookup; raw pointers below keep it out of the loops
    ws.kept.assign(nc, 1);
    ws.verdict.resize(nc); ws.bE.resize(nc); ws.bS.resize(nc); ws.bn.resize(nc);
    ws.parent.resize(nc); ws.label.assign(nc, 0);
    std::vector<uint8_t>& kept = ws.kept;
    uint8_t* const verdict = ws.verdict.data();
    double* const bE = ws.bE.data();
    double* const bS = ws.bS.data();
    uint32_t* const bn = ws.bn.data();
    uint32_t* const parent = ws.parent.data();
    int32_t* const label = ws.label.data();

    // ---- step 3: contrack.py:706-742.  Plane t reads plane t-1 AFTER it was filtered and plane t+1 BEFORE. ----
    // (local copies of the table pointers: the byte-typed `kept` / `verdict` stores may alias anything otherwise)
    const int32_t* const comp_t = tb.comp_t;
    const uint32_t* const comp_cls = tb.comp_cls;
    const uint32_t* const pair_ptr = tb.pair_ptr;
    const uint32_t* const pair_b = tb.pair_b;
    const uint32_t* const pair_nsp = tb.pair_nsp;
    const double* const pair_E = tb.pair_E;
    const double* const pair_S = tb.pair_S;
    uint8_t* const keptp = kept.data();
    const bool two = pr.twosided != 0;
    const double band = 1e-9, ov = pr.overlap;
    long c0 = 0;
    while (c0 < nc && comp_t[c0] < 1) ++c0;                          // plane 0 is never filtered
    while (c0 < nc) {
        const long t = comp_t[c0];
        long c1 = c0;
        while (c1 < nc && comp_t[c1] == t) ++c1;
        if (c1 < nc && comp_t[c1] < t) { err = "component table not sorted by time"; return -5; }
        if (t + 1 >= T) break;                                       // neither is the last plane
        for (long c = c0; c < c1; ++c) {
            const long rep = (long)comp_cls[c];
            // rep must lie in [c0, c].  Written as ONE unsigned comparison on purpose: g++ 13.3 folds the natural
            // `rep > c || rep < c0` to `true` in the peeled first iteration (c == c0), where it reads
            // (rep > c0) | (c0 > rep) -- dom2: "Relation adjustment ... combine to produce [1, 1]" (forgets rep == c0).
            if ((unsigned long)(rep - c0) > (unsigned long)(c - c0)) {
                err = "class representative is not the smallest id of its class"; return -5;
            }
            double e = 0.0, s2 = 0.0;
            uint32_t n = 0;
            for (uint32_t k = pair_ptr[c]; k < pair_ptr[c + 1]; ++k) {
                if (!keptp[pair_b[k]]) continue;
                e += pair_E[k]; s2 += pair_S[k]; n += pair_nsp[k];
            }
            if (rep == c) { bE[c] = e; bS[c] = s2; bn[c] = n; }
            else { bE[rep] += e; bS[rep] += s2; bn[rep] += n; }
        }
        for (long c = c0; c < c1; ++c) {
            if ((long)comp_cls[c] != c) continue;
            const double areacon = tb.cls_conE[c] + tb.cls_conS[c], fwd = tb.cls_fE[c] + tb.cls_fS[c];
            const double bwd = bE[c] + bS[c];
            double fb, ff;
            bool kill = kill_decision(areacon, fwd, bwd, ov, two, &fb, &ff);
            if (tb.cls_nsp[c] + (tb.cls_fnsp ? tb.cls_fnsp[c] : 0u) + bn[c] > 0) {
                // Sums that include special-row weights are not exactly summable: numpy's pairwise order decides the
                // last bits.  Only a fraction within rounding distance of `overlap` can flip the decision.
                // (a class that lies entirely in special rows of one common weight sums exactly in any order)
                const bool exact = tb.special_uniform && tb.cls_conE[c] == 0.0 && tb.cls_fE[c] == 0.0 && bE[c] == 0.0;
                const bool near = !exact && ((std::fabs(ff - ov) <= band) || (two && std::fabs(fb - ov) <= band));
                if (near) {
                    double ac2, f2, b2;
                    if (!exact_class_sums(pc, comp_cls, tb.w, kept, t, (uint32_t)c, &ac2, &f2, &b2)) {
                        err = "near-tie overlap decision on special rows and no run source to resolve it"; return -4;
                    }
                    kill = kill_decision(ac2, f2, b2, ov, two, &fb, &ff);
                    out.n_neartie++;
                }
            }
            verdict[c] = kill ? 1 : 0;
        }
        for (long c = c0; c < c1; ++c) keptp[c] = verdict[comp_cls[c]] ? 0 : 1;
        c0 = c1;
    }
    long nkept = 0;
    for (long c = 0; c < nc; ++c) nkept += kept[c];
    out.n_kept = nkept;
    lap("step3");
    if (pr.stage == 3) {
        for (long c = 0; c < nc; ++c) comp_val[c] = kept[c] ? (int32_t)(tb.comp_cls[c] + 1) : 0;
        return 0;
    }

    // ---- step 4a/b: contrack.py:747-751.  3-D components = kept 2-D components joined by common pixels in adjacent
    // planes; scipy numbers them by first pixel in (t, y, x) order = rank of the smallest member id.  Partners of c are in
    // the plane before, so they have smaller ids and were processed already. ----
    auto find = [&](uint32_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    for (long c = 0; c < nc; ++c) {
        parent[c] = (uint32_t)c;
        if (!kept[c]) continue;
        uint32_t m = (uint32_t)c;
        for (uint32_t k = tb.pair_ptr[c]; k < tb.pair_ptr[c + 1]; ++k) {
            const uint32_t b = tb.pair_b[k];
            if (b >= (uint32_t)c) { err = "pair table: partner is not in an earlier plane"; return -5; }
            if (!kept[b] || tb.pair_npix[k] == 0) continue;
            const uint32_t r = find(b);
            if (r == m) continue;
            if (r < m) { parent[m] = r; m = r; } else parent[r] = m;
        }
    }
    int32_t nlab = 0;
    for (long c = 0; c < nc; ++c) {
        if (!kept[c]) continue;
        const uint32_t r = find((uint32_t)c);
        label[c] = (r == (uint32_t)c) ? ++nlab : label[r];
    }
    out.n_labels3d = nlab;
    lap("link3d");
    if (pr.stage == 4) { memcpy(comp_val, label, (size_t)nc * 4); return 0; }

    int rc = track_phase(tb, label, pr.persistence, runs, comp_val, out, err);
    lap("track");
    return rc;
}

int track_phase(const FastTables& tb, const int32_t* label, int persistence, RunSource* runs, int32_t* comp_val,
                Result& out, std::string& err) {
    // ---- step 4c/d (ct_tables.cpp): removed components carry label 0 and are skipped there ----
    const long nc = tb.ncomp;
    static thread_local std::vector<int32_t> seg_a32, seg_b32;
    seg_a32.resize(tb.nseg); seg_b32.resize(tb.nseg);
    for (long s = 0; s < tb.nseg; ++s) {
        if (tb.seg_a[s] >= (uint32_t)nc || tb.seg_b[s] >= (uint32_t)nc) { err = "seam table out of range"; return -5; }
        seg_a32[s] = (int32_t)tb.seg_a[s]; seg_b32[s] = (int32_t)tb.seg_b[s];
    }
    PlaneCache pc(runs, tb.H);
    SplitFetcher fetcher;
    fetcher.pc = &pc; fetcher.comp_t = tb.comp_t;
    ctb::TrackStats stats;
    int rc;
    if (getenv("CT_TRACK_EVENTS")) {
        // test hook: what the cooperative global kernel + track_events_fast do in the product path -- label boxes, persistence
        // of every label from its own box, the event list (segments whose ends carry different labels, as indices into the
        // records of the labels that occur in events), the replay on the events alone, patches on top.  Falls through to the
        // per-component replay when a label straddles a stale box.
        int nlabel = 0;
        for (long c = 0; c < nc; ++c) nlabel = std::max(nlabel, (int)label[c]);
        std::vector<int32_t> t0(nlabel + 1, INT32_MAX), t1(nlabel + 1, 0), y0(nlabel + 1, INT32_MAX), y1(nlabel + 1, 0),
            x0(nlabel + 1, INT32_MAX), x1(nlabel + 1, 0);
        for (long c = 0; c < nc; ++c) {
            const int v = label[c];
            if (!v) continue;
            t0[v] = std::min(t0[v], tb.comp_t[c]); t1[v] = std::max(t1[v], tb.comp_t[c] + 1);
            y0[v] = std::min(y0[v], tb.comp_y0[c]); y1[v] = std::max(y1[v], tb.comp_y1[c]);
            x0[v] = std::min(x0[v], tb.comp_x0[c]); x1[v] = std::max(x1[v], tb.comp_x1[c]);
        }
        std::vector<int32_t> ev, lrec, lidx(nlabel + 1, -1), lab_fin;
        for (long s = 0; s < tb.nseg; ++s) {
            const int la = label[seg_a32[s]], lb = label[seg_b32[s]];
            if (la == 0 || lb == 0 || la == lb) continue;
            lidx[la] = 0; lidx[lb] = 0;
        }
        for (int v = 1; v <= nlabel; ++v) {
            if (lidx[v] < 0) continue;
            lidx[v] = (int32_t)(lrec.size() / 7);
            const int32_t r[7] = {v, t0[v], t1[v], y0[v], y1[v], x0[v], x1[v]};
            lrec.insert(lrec.end(), r, r + 7);
        }
        for (long s = 0; s < tb.nseg; ++s) {
            const int la = label[seg_a32[s]], lb = label[seg_b32[s]];
            if (la == 0 || lb == 0 || la == lb) continue;
            ev.push_back(lidx[la]); ev.push_back(lidx[lb]);
        }
        std::vector<int32_t> pl, pv;
        long delta = 0, feats = 0;
        if (ctb::track_events_fast(persistence, (long)ev.size() / 2, ev.data(), (long)lrec.size() / 7, lrec.data(), pl, pv, &delta,
                                   stats) == 0) {
            lab_fin.assign(nlabel + 1, 0);
            for (int v = 1; v <= nlabel; ++v) {
                const bool keep = t1[v] > t0[v] && (t1[v] - t0[v]) >= persistence;
                lab_fin[v] = keep ? v : 0; feats += keep;
            }
            for (size_t i = 0; i < pl.size(); ++i) lab_fin[pl[i]] = pv[i];
            for (long c = 0; c < nc; ++c) comp_val[c] = lab_fin[label[c]];
            out.overrides.clear();
            out.n_features = feats + delta; out.n_seam_events = stats.n_events; out.n_seam_splits = 0;
            out.n_neartie += 1000000;           // marker for the test: the event replay produced this result
            return 0;
        }
    }
    {
        rc = ctb::track_tables(tb.T, tb.H, tb.W, persistence, nc, tb.comp_t, tb.comp_y0, tb.comp_y1, tb.comp_x0,
                               tb.comp_x1, label, tb.nseg, tb.seg_t, tb.seg_y0, tb.seg_y1, seg_a32.data(),
                               seg_b32.data(), runs ? &fetcher : nullptr, comp_val, out.overrides, stats);
    }
    if (rc != 0) { err = "date-line merge needs to split a component and no run source is available"; return -5; }
    out.n_features = stats.n_features; out.n_seam_events = stats.n_events; out.n_seam_splits = stats.n_splits;
    return 0;
}

// Unsorted tables (tests, ct_host_tables) -> the fast layout.
int host_phase(const Tables& tb, const Params& pr, RunSource* runs, Result& out, std::string& err) {
    const long nc = tb.ncomp, np = tb.npair;
    out = Result();
    out.comp_val.assign(nc, 0);
    for (long c = 0; c < nc; ++c)
        if (tb.comp_cls[c] >= (uint32_t)nc) { err = "class table out of range"; return -5; }
    std::vector<double> conE(nc, 0.0), conS(nc, 0.0), fE(nc, 0.0), fS(nc, 0.0);
    std::vector<uint32_t> cnsp(nc, 0), pptr(nc + 1, 0), pb(np), pn(np), pnsp(np);
    std::vector<double> pE(np), pS(np);
    for (long c = 0; c < nc; ++c) {
        const uint32_t rep = tb.comp_cls[c];
        conE[rep] += tb.comp_areaE[c]; conS[rep] += tb.comp_areaS[c]; cnsp[rep] += tb.comp_nsp[c];
    }
    for (long p = 0; p < np; ++p) {
        if (tb.pair_a[p] >= nc || tb.pair_b[p] >= nc) { err = "pair table out of range"; return -5; }
        pptr[tb.pair_a[p] + 1]++;
        const uint32_t rep = tb.comp_cls[tb.pair_b[p]];
        fE[rep] += tb.pair_areaE[p]; fS[rep] += tb.pair_areaS[p]; cnsp[rep] += tb.pair_nsp[p];
    }
    for (long c = 0; c < nc; ++c) pptr[c + 1] += pptr[c];
    {
        std::vector<uint32_t> pos(pptr.begin(), pptr.end() - 1);
        for (long p = 0; p < np; ++p) {
            const uint32_t k = pos[tb.pair_a[p]]++;
            pb[k] = tb.pair_b[p]; pn[k] = tb.pair_npix[p]; pnsp[k] = tb.pair_nsp[p];
            pE[k] = tb.pair_areaE[p]; pS[k] = tb.pair_areaS[p];
        }
    }
    std::vector<int32_t> st, sy0, sy1;
    std::vector<uint32_t> sa, sb;
    for (long s = 0; s < tb.nseam; ++s) {
        const int32_t t = (int32_t)(tb.seam_row[s] / tb.H), y = (int32_t)(tb.seam_row[s] % tb.H);
        if (!st.empty() && st.back() == t && sy1.back() == y && sa.back() == tb.seam_a[s] && sb.back() == tb.seam_b[s]) {
            sy1.back() = y + 1;
        } else {
            st.push_back(t); sy0.push_back(y); sy1.push_back(y + 1); sa.push_back(tb.seam_a[s]); sb.push_back(tb.seam_b[s]);
        }
    }
    FastTables ft;
    ft.T = tb.T; ft.H = tb.H; ft.W = tb.W; ft.ncomp = nc; ft.comp_t = tb.comp_t; ft.comp_y0 = tb.comp_y0;
    ft.comp_y1 = tb.comp_y1; ft.comp_x0 = tb.comp_x0; ft.comp_x1 = tb.comp_x1; ft.comp_cls = tb.comp_cls;
    ft.cls_conE = conE.data(); ft.cls_conS = conS.data(); ft.cls_fE = fE.data(); ft.cls_fS = fS.data();
    ft.cls_nsp = cnsp.data(); ft.pair_ptr = pptr.data(); ft.pair_b = pb.data(); ft.pair_npix = pn.data();
    ft.pair_nsp = pnsp.data(); ft.pair_E = pE.data(); ft.pair_S = pS.data();
    ft.nseg = (long)st.size(); ft.seg_t = st.data(); ft.seg_y0 = sy0.data(); ft.seg_y1 = sy1.data();
    ft.seg_a = sa.data(); ft.seg_b = sb.data(); ft.w = tb.w; ft.special_uniform = tb.special_uniform;
    for (long c = 0; c < nc; ++c)
        if (c > 0 && tb.comp_t[c] < tb.comp_t[c - 1]) { err = "component table not sorted by time"; return -5; }
    return host_phase_fast(ft, pr, runs, out.comp_val.data(), out, err);
}

}  // namespace cth
