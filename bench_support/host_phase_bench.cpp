// Timing harness for the ordered host phase on tables dumped by a GPU run (CT_DUMP_TABLES=file, fast layout).
//   g++ -O2 -std=c++17 -ffp-contract=off -Icontrack_b200/csrc bench_support/host_phase_bench.cpp \
//       contrack_b200/csrc/ct_host.cpp contrack_b200/csrc/ct_tables.cpp -o /tmp/hb && /tmp/hb tables.bin [overlap]
#include "ct_host.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace std;
template <class T> vector<T> rd(FILE* f, long n) { vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != (size_t)n) abort(); return v; }
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    long h[8];
    if (!f || fread(h, 8, 8, f) != 8 || h[6] != 1) { fprintf(stderr, "need a fast-layout dump\n"); return 1; }
    const long T = h[0], H = h[1], W = h[2], nc = h[3], np = h[4], ng = h[5];
    auto w = rd<double>(f, H);
    auto t = rd<int32_t>(f, nc), y0 = rd<int32_t>(f, nc), y1 = rd<int32_t>(f, nc), x0 = rd<int32_t>(f, nc), x1 = rd<int32_t>(f, nc);
    auto cls = rd<uint32_t>(f, nc);
    auto conE = rd<double>(f, nc), conS = rd<double>(f, nc), fE = rd<double>(f, nc), fS = rd<double>(f, nc);
    auto knsp = rd<uint32_t>(f, nc), pptr = rd<uint32_t>(f, nc + 1);
    auto pb = rd<uint32_t>(f, np), pn = rd<uint32_t>(f, np), pnsp = rd<uint32_t>(f, np);
    auto pE = rd<double>(f, np), pS = rd<double>(f, np);
    auto gt = rd<int32_t>(f, ng), gy0 = rd<int32_t>(f, ng), gy1 = rd<int32_t>(f, ng);
    auto ga = rd<uint32_t>(f, ng), gb = rd<uint32_t>(f, ng);
    cth::FastTables tb;
    tb.T = T; tb.H = (int)H; tb.W = (int)W; tb.ncomp = nc; tb.comp_t = t.data(); tb.comp_y0 = y0.data(); tb.comp_y1 = y1.data();
    tb.comp_x0 = x0.data(); tb.comp_x1 = x1.data(); tb.comp_cls = cls.data(); tb.cls_conE = conE.data(); tb.cls_conS = conS.data();
    tb.cls_fE = fE.data(); tb.cls_fS = fS.data(); tb.cls_nsp = knsp.data(); tb.pair_ptr = pptr.data(); tb.pair_b = pb.data();
    tb.pair_npix = pn.data(); tb.pair_nsp = pnsp.data(); tb.pair_E = pE.data(); tb.pair_S = pS.data();
    tb.nseg = ng; tb.seg_t = gt.data(); tb.seg_y0 = gy0.data(); tb.seg_y1 = gy1.data(); tb.seg_a = ga.data(); tb.seg_b = gb.data();
    tb.w = w.data();
    cth::Params pr;
    pr.overlap = argc > 2 ? atof(argv[2]) : 0.51; pr.persistence = 5; pr.twosided = 1;
    vector<int32_t> val(nc + 1);
    printf("T %ld comps %ld pairs %ld segments %ld\n", T, nc, np, ng);
    {   // steps 4c/4d alone, on the 3-D labels (what the host still does when step 3 / linking run on the device)
        cth::Params p4 = pr; p4.stage = 4;
        vector<int32_t> label(nc + 1);
        cth::Result r; string err;
        cth::host_phase_fast(tb, p4, nullptr, label.data(), r, err);
        for (int i = 0; i < 6; i++) {
            auto a = chrono::steady_clock::now();
            int rc = cth::track_phase(tb, label.data(), pr.persistence, nullptr, val.data(), r, err);
            auto b = chrono::steady_clock::now();
            printf("track_phase rc %d features %ld events %ld: %.3f ms\n", rc, r.n_features, r.n_seam_events,
                   chrono::duration<double, milli>(b - a).count());
        }
    }
    for (int i = 0; i < 6; i++) {
        cth::Result r; string err;
        auto a = chrono::steady_clock::now();
        int rc = cth::host_phase_fast(tb, pr, nullptr, val.data(), r, err);
        auto b = chrono::steady_clock::now();
        printf("rc %d %s features %ld kept %ld labels %ld events %ld: %.3f ms\n", rc, err.c_str(), r.n_features, r.n_kept,
               r.n_labels3d, r.n_seam_events, chrono::duration<double, milli>(b - a).count());
    }
}
