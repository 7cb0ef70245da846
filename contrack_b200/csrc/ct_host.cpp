// ct_host.cpp -- see ct_host.h.  Pure host C++; compiled with -ffp-contract=off: the float64 decisions of step 3 must
// round exactly like the reference's numpy scalar arithmetic (contrack/contrack.py:721-722).
#include "ct_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>

namespace cth {

namespace {

// contrack.py:721-742 -- reciprocal, then multiply (two roundings), then the keep/kill cascade.
inline bool kill_decision(double areacon, double fwd, double bwd, double ov, bool twosided, double* fb_out,
                          double* ff_out) {
    volatile double inv = 1.0 / areacon;
    volatile double fb = inv * bwd;
    volatile double ff = inv * fwd;
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

struct Acc { double conE, conS, fE, fS, bE, bS; uint64_t nsp; };

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
bool exact_class_sums(PlaneCache& pc, const Tables& tb, const std::vector<uint8_t>& kept, long t, uint32_t rep,
                      double* areacon, double* fwd, double* bwd) {
    const Plane* cur = pc.get(t);
    const Plane* nxt = pc.get(t + 1);
    const Plane* prv = pc.get(t - 1);
    if (!cur || !nxt || !prv) return false;
    std::vector<double> con, f, b;
    for (const PlaneRun& r : cur->runs) {
        if (tb.comp_cls[r.comp] != rep) continue;
        const double wy = tb.w[r.y];
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
    PlaneCache* pc; const Tables* tb; const std::vector<long>* kept_ids;
    bool fetch(long k, std::vector<ctb::SubRun>& out) override {
        long c = (*kept_ids)[k];
        const Plane* p = pc->get(tb->comp_t[c]);
        if (!p) return false;
        out.clear();
        for (const PlaneRun& r : p->runs)
            if ((long)r.comp == c) out.push_back(ctb::SubRun{r.y, r.x0, r.x1});
        return !out.empty();
    }
};

}  // namespace

int host_phase(const Tables& tb, const Params& pr, RunSource* runs, Result& out, std::string& err) {
    const long T = tb.T, nc = tb.ncomp, np = tb.npair;
    out = Result();
    out.comp_val.assign(nc, 0);
    if (pr.stage == 1) { for (long c = 0; c < nc; ++c) out.comp_val[c] = (int32_t)(c + 1); return 0; }
    if (pr.stage == 2) { for (long c = 0; c < nc; ++c) out.comp_val[c] = (int32_t)(tb.comp_cls[c] + 1); return 0; }

    PlaneCache pc(runs, tb.H);

    // components per plane
    std::vector<long> plane_ptr(T + 2, 0);
    for (long c = 0; c < nc; ++c) {
        if (tb.comp_t[c] < 0 || tb.comp_t[c] >= T || (c > 0 && tb.comp_t[c] < tb.comp_t[c - 1])) {
            err = "component table not sorted by time"; return -5;
        }
        plane_ptr[tb.comp_t[c] + 1]++;
    }
    for (long t = 0; t < T; ++t) plane_ptr[t + 1] += plane_ptr[t];

    // pairs bucketed by their plane-t component (a); forward sums per plane-(t-1) component (b)
    std::vector<uint32_t> pa_ptr(nc + 1, 0), pa_idx(np);
    std::vector<double> fE(nc, 0.0), fS(nc, 0.0);
    std::vector<uint32_t> fnsp(nc, 0);
    for (long p = 0; p < np; ++p) {
        if (tb.pair_a[p] >= nc || tb.pair_b[p] >= nc) { err = "pair table out of range"; return -5; }
        pa_ptr[tb.pair_a[p] + 1]++;
        const uint32_t b = tb.pair_b[p];
        fE[b] += tb.pair_areaE[p]; fS[b] += tb.pair_areaS[p]; fnsp[b] += tb.pair_nsp[p];
    }
    for (long c = 0; c < nc; ++c) pa_ptr[c + 1] += pa_ptr[c];
    {
        std::vector<uint32_t> pos(pa_ptr.begin(), pa_ptr.end() - 1);
        for (long p = 0; p < np; ++p) pa_idx[pos[tb.pair_a[p]]++] = (uint32_t)p;
    }

    for (long c = 0; c < nc; ++c) {
        const long rep = (long)tb.comp_cls[c];
        if (rep > c || rep < plane_ptr[tb.comp_t[c]]) {
            err = "class representative is not the smallest id of its class (component " + std::to_string(c) + ")";
            return -5;
        }
    }

    // ---- step 3: contrack.py:706-742.  Plane t reads plane t-1 AFTER it was filtered and plane t+1 BEFORE. ----
    std::vector<uint8_t> kept(nc, 1);
    std::vector<Acc> acc(nc);
    const bool two = pr.twosided != 0;
    const double band = 1e-9;
    for (long t = 1; t + 1 < T; ++t) {
        const long c0 = plane_ptr[t], c1 = plane_ptr[t + 1];
        for (long c = c0; c < c1; ++c) {
            const uint32_t rep = tb.comp_cls[c];
            if (rep == c) acc[rep] = Acc{0, 0, 0, 0, 0, 0, 0};
            Acc& a = acc[rep];
            a.conE += tb.comp_areaE[c]; a.conS += tb.comp_areaS[c]; a.nsp += tb.comp_nsp[c];
            a.fE += fE[c]; a.fS += fS[c]; a.nsp += fnsp[c];
            for (uint32_t k = pa_ptr[c]; k < pa_ptr[c + 1]; ++k) {
                const uint32_t p = pa_idx[k];
                if (!kept[tb.pair_b[p]]) continue;
                a.bE += tb.pair_areaE[p]; a.bS += tb.pair_areaS[p]; a.nsp += tb.pair_nsp[p];
            }
        }
        for (long c = c0; c < c1; ++c) {
            if (tb.comp_cls[c] != c) continue;
            Acc& a = acc[c];
            double areacon = a.conE + a.conS, fwd = a.fE + a.fS, bwd = a.bE + a.bS;
            double fb, ff;
            bool kill = kill_decision(areacon, fwd, bwd, pr.overlap, two, &fb, &ff);
            if (a.nsp > 0) {
                // Sums that include special-row weights are not exactly summable: numpy's pairwise order decides the
                // last bits.  Only a fraction within rounding distance of `overlap` can flip the decision.
                bool near = (std::fabs(ff - pr.overlap) <= band) || (two && std::fabs(fb - pr.overlap) <= band);
                if (near) {
                    double ac2, f2, b2;
                    if (!exact_class_sums(pc, tb, kept, t, (uint32_t)c, &ac2, &f2, &b2)) {
                        err = "near-tie overlap decision on special rows and no run source to resolve it"; return -4;
                    }
                    kill = kill_decision(ac2, f2, b2, pr.overlap, two, &fb, &ff);
                    out.n_neartie++;
                }
            }
            a.nsp = kill ? 1 : 0;                      // reuse as the class verdict
        }
        for (long c = c0; c < c1; ++c) kept[c] = acc[tb.comp_cls[c]].nsp ? 0 : 1;
    }
    for (long c = 0; c < nc; ++c) out.n_kept += kept[c];
    if (pr.stage == 3) {
        for (long c = 0; c < nc; ++c) out.comp_val[c] = kept[c] ? (int32_t)(tb.comp_cls[c] + 1) : 0;
        return 0;
    }

    // ---- step 4a/b: contrack.py:747-751.  3-D components = kept 2-D components joined by common pixels in adjacent
    // planes; scipy numbers them by first pixel in (t, y, x) order = rank of the smallest member id. ----
    std::vector<uint32_t> parent(nc);
    for (long c = 0; c < nc; ++c) parent[c] = (uint32_t)c;
    auto find = [&](uint32_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    for (long p = 0; p < np; ++p) {
        const uint32_t a = tb.pair_a[p], b = tb.pair_b[p];
        if (!kept[a] || !kept[b] || tb.pair_npix[p] == 0) continue;
        uint32_t ra = find(a), rb = find(b);
        if (ra == rb) continue;
        if (ra < rb) parent[rb] = ra; else parent[ra] = rb;
    }
    std::vector<int32_t> label(nc, 0);
    int32_t nlab = 0;
    for (long c = 0; c < nc; ++c) if (kept[c] && parent[c] == (uint32_t)c) label[c] = ++nlab;
    for (long c = 0; c < nc; ++c) if (kept[c]) label[c] = label[find((uint32_t)c)];
    out.n_labels3d = nlab;
    if (pr.stage == 4) { for (long c = 0; c < nc; ++c) out.comp_val[c] = label[c]; return 0; }

    // ---- step 4c/d on the kept components (ct_tables.cpp) ----
    std::vector<long> kept_ids;
    kept_ids.reserve(out.n_kept);
    std::vector<int32_t> kidx(nc, -1);
    for (long c = 0; c < nc; ++c) if (kept[c]) { kidx[c] = (int32_t)kept_ids.size(); kept_ids.push_back(c); }
    const long nk = (long)kept_ids.size();
    std::vector<int32_t> kt(nk), ky0(nk), ky1(nk), kx0(nk), kx1(nk), klab(nk), kval(nk, 0);
    for (long k = 0; k < nk; ++k) {
        const long c = kept_ids[k];
        kt[k] = tb.comp_t[c]; ky0[k] = tb.comp_y0[c]; ky1[k] = tb.comp_y1[c]; kx0[k] = tb.comp_x0[c];
        kx1[k] = tb.comp_x1[c]; klab[k] = label[c];
    }
    std::vector<int32_t> st, sy0, sy1, sa, sb;
    for (long s = 0; s < tb.nseam; ++s) {
        const uint32_t a = tb.seam_a[s], b = tb.seam_b[s];
        if (a >= nc || b >= nc) { err = "seam table out of range"; return -5; }
        if (!kept[a] || !kept[b]) continue;
        const int32_t t = (int32_t)(tb.seam_row[s] / tb.H), y = (int32_t)(tb.seam_row[s] % tb.H);
        if (!st.empty() && st.back() == t && sy1.back() == y && sa.back() == kidx[a] && sb.back() == kidx[b]) {
            sy1.back() = y + 1;
        } else {
            st.push_back(t); sy0.push_back(y); sy1.push_back(y + 1); sa.push_back(kidx[a]); sb.push_back(kidx[b]);
        }
    }
    SplitFetcher fetcher;
    fetcher.pc = &pc; fetcher.tb = &tb; fetcher.kept_ids = &kept_ids;
    ctb::TrackStats stats;
    int rc = ctb::track_tables(T, tb.H, tb.W, pr.persistence, nk, kt.data(), ky0.data(), ky1.data(), kx0.data(),
                               kx1.data(), klab.data(), (long)st.size(), st.data(), sy0.data(), sy1.data(), sa.data(),
                               sb.data(), runs ? &fetcher : nullptr, kval.data(), out.overrides, stats);
    if (rc != 0) { err = "date-line merge needs to split a component and no run source is available"; return -5; }
    for (long k = 0; k < nk; ++k) out.comp_val[kept_ids[k]] = kval[k];
    out.n_features = stats.n_features; out.n_seam_events = stats.n_events; out.n_seam_splits = stats.n_splits;
    return 0;
}

}  // namespace cth
