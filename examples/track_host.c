/* examples/track_host.c -- the tracking path from plain C through the C ABI, with HOST buffers: no CUDA header, no CUDA
 * call in this file.  This is what a C / Fortran / Julia host program (or a cgo / JNI stub) binds; the Python class in
 * contrack_b200/contrack.py does the same through ctypes.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/track_host.c -Lcontrack_b200/lib -lcontrack_b200 -lm -o track_host
 *   LD_LIBRARY_PATH=contrack_b200/lib ./track_host [anom.f32 flag.i32]      (optional: raw dumps of the two cubes)
 *
 * The cube: three anomaly blobs on a 2-degree grid, 24 time steps; one blob drifts across the date line, one lives for
 * two steps only (removed by persistence = 4), one jumps too far between steps 11 and 12 (cut by overlap = 0.5).
 * Equivalent reference call: block.run_contrack(variable='anom', threshold=150, gorl='>=', overlap=0.5, persistence=4,
 * twosided=True)  (contrack.py:583-796). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "contrack_b200.h"

enum { T = 24, H = 91, W = 180 };

static void blob(float* plane, double cy, double cx, double ry, double rx, float amp) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            double dx = fabs(x - cx);
            if (dx > W / 2.0) dx = W - dx;                       /* periodic in longitude */
            const double d = (y - cy) * (y - cy) / (ry * ry) + dx * dx / (rx * rx);
            if (d < 1.0) plane[(size_t)y * W + x] += amp * (float)(1.0 - d);
        }
}

static int dump(const char* path, const void* p, size_t bytes) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    const size_t n = fwrite(p, 1, bytes, f);
    fclose(f);
    return n == bytes ? 0 : -1;
}

int main(int argc, char** argv) {
    float* anom = calloc((size_t)T * H * W, sizeof(float));
    int32_t* flag = malloc((size_t)T * H * W * sizeof(int32_t));
    double w[H];
    if (!anom || !flag) return 1;
    for (int t = 0; t < T; ++t) {
        float* p = anom + (size_t)t * H * W;
        blob(p, 25.0, fmod(165.0 + 2.0 * t, W), 8.0, 14.0, 400.f);              /* crosses x = W - 1 -> 0 */
        if (t == 5 || t == 6) blob(p, 60.0, 60.0, 6.0, 9.0, 400.f);             /* too short-lived */
        blob(p, 70.0, t < 12 ? 100.0 + t : 150.0 + t, 5.0, 8.0, 400.f);         /* jumps at t = 12 */
    }
    /* area weight per latitude row, the reference's expression (contrack.py:703-704): float32 values */
    for (int y = 0; y < H; ++y) {
        const double lat = 90.0 - 2.0 * y;
        w[y] = (double)(float)(111 * 2.0 * 111 * 2.0 * cos(lat * 3.14159265358979323846 / 180.0));
    }
    if (argc > 1 && dump(argv[1], anom, (size_t)T * H * W * sizeof(float))) return 1;
    ct_ctx* ctx = NULL;
    if (ct_create(0, &ctx) != CT_OK) {
        fprintf(stderr, "ct_create: %s\n", ct_last_error());
        return 2;
    }
    const double thr = 150.0;
    long nfeat = 0;
    const int rc = ct_run_contrack_host(ctx, anom, CT_F32, T, H, W, w, &thr, 1, /*thr_is_f32*/ 1, CT_GE, /*overlap*/ 0.5,
                                        /*persistence*/ 4, /*twosided*/ 1, flag, &nfeat, /*chunk_planes*/ 0);
    if (rc != CT_OK) {
        fprintf(stderr, "ct_run_contrack_host: %s\n", ct_last_error());
        ct_destroy(ctx);
        return 3;
    }
    if (argc > 2 && dump(argv[2], flag, (size_t)T * H * W * sizeof(int32_t))) return 1;
    long cells = 0;
    int32_t maxid = 0;
    for (size_t i = 0; i < (size_t)T * H * W; ++i) {
        cells += flag[i] != 0;
        if (flag[i] > maxid) maxid = flag[i];
    }
    printf("features %ld  flagged cells %ld  largest id %d  (3-D labels before filtering: %.0f)\n", nfeat, cells, (int)maxid,
           ct_get_stat(ctx, "labels3d"));
    ct_destroy(ctx);
    free(anom);
    free(flag);
    return 0;
}
