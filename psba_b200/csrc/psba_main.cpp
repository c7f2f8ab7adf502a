// psba_main.cpp -- the host program of PSBA (PSBA/main.cpp:70-231) written against include/psba_b200.h only:
// a plain C++ consumer of the C ABI (compiled by g++, linked to libpsba_b200.so; no CUDA headers, no torch).
//
//   psba_main <cams.txt> <pts.txt> [origin_cnp=11] [--out-cams f] [--out-pts f] [--ply f] [--verbose]
//
// Same sequence as the reference: readInitialSBAEstimate (+ quat2vec filter, local rotation := 0, K / extrinsics
// split, main.cpp:102-149) -> setup_cl -> fill_initBuffer2 -> generate_idxs / fill_idxBuffer (the dense tables are
// built by the engine) -> while(true){ levmar(); trust_region(); } (main.cpp:192-209) -> report (main.cpp:214-218).
// The dataset is an argument instead of an edit-and-recompile path (main.cpp:40-65), the result can be saved
// (the reference never writes it back), and nothing is written to e:\psba_debug.txt.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "psba_b200.h"

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <cams.txt> <pts.txt> [origin_cnp] [--out-cams f --out-pts f] [--ply f] [--verbose]\n", argv[0]);
        return 2;
    }
    const char *cams_file = argv[1], *pts_file = argv[2];
    int origin_cnp = 11;                                     // main.cpp:73
    const int pnp = 3, mnp = 2;                              // main.cpp:74-75
    std::string out_cams, out_pts, out_ply;
    int verbose = 0;
    for (int a = 3; a < argc; ++a) {
        if (!strcmp(argv[a], "--verbose")) verbose = 1;
        else if (!strcmp(argv[a], "--out-cams") && a + 1 < argc) out_cams = argv[++a];
        else if (!strcmp(argv[a], "--out-pts") && a + 1 < argc) out_pts = argv[++a];
        else if (!strcmp(argv[a], "--ply") && a + 1 < argc) out_ply = argv[++a];
        else origin_cnp = atoi(argv[a]);
    }
    // intrinsics of the 7-column demo sets (data/7camsvarK.txt:1; hard-coded in the reference's main_bak.cpp:31-32)
    const double Kdefault[5] = {851.57945, 330.24755, 262.19500, 1.00169, 0.0};

    int nCams = 0, n3Dpts = 0, n2Dprojs = 0;
    double *Kparas, *initrot, *camsExParas, *pts3D, *impts_data;
    int *iidx, *jidx;
    if (psba_readInitialSBAEstimate(cams_file, pts_file, origin_cnp, Kdefault, &nCams, &n3Dpts, &n2Dprojs, &Kparas, &initrot,
                                    &camsExParas, &pts3D, &impts_data, &iidx, &jidx) != 0)
        return 1;
    printf("cameras: %d\n3D points: %d\n2D projections: %d\n", nCams, n3Dpts, n2Dprojs);

    const int final_cnp = 6;                                 // main.cpp:140 (origin_cnp - 5 for the K|q|t files)
    psba_ctx *ctx = psba_setup_cl(final_cnp, pnp, mnp, nCams, n3Dpts, n2Dprojs);                       // main.cpp:177
    psba_fill_initBuffer2(ctx, final_cnp, pnp, mnp, nCams, n3Dpts, n2Dprojs, Kparas, impts_data, initrot, camsExParas, pts3D);
    psba_fill_idxBuffer(ctx, nCams, n3Dpts, n2Dprojs, iidx, jidx);                                     // main.cpp:188-189
    psba_set_option(ctx, "verbose", verbose);

    double initErr = 0.0, finalErr = 0.0;
    int itno = 0;
    const auto t0 = std::chrono::steady_clock::now();
    const int iter_flag = psba_solve(ctx, &initErr, &finalErr, &itno);                                  // main.cpp:192-209
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("iter_flag=%d\n", iter_flag);                                                               // main.cpp:211
    printf("time eclipse %lf s\n", secs);
    printf("initial error: %.15E \n", sqrt(initErr) / n2Dprojs);                                       // main.cpp:215-217
    printf("final error: %.15E \n", sqrt(finalErr) / n2Dprojs);
    printf("total iteration: %d\n", itno);
    printf("initial cost: %.15E\nfinal cost: %.15E\n", initErr, finalErr);

    if (!out_cams.empty() || !out_ply.empty()) {
        std::vector<double> cams((size_t)nCams * 6), pts((size_t)n3Dpts * 3);
        psba_get_params(ctx, PSBA_PARAMS_CUR, cams.data(), pts.data());
        if (!out_cams.empty() && !out_pts.empty() &&
            psba_write_sba_result(out_cams.c_str(), out_pts.c_str(), nCams, n3Dpts, n2Dprojs, Kparas, initrot, cams.data(), pts.data(),
                                  impts_data, iidx, jidx) != 0)
            return 1;
        if (!out_ply.empty() && psba_write_ply(out_ply.c_str(), nCams, n3Dpts, initrot, cams.data(), pts.data()) != 0) return 1;
    }
    psba_release_buffer(ctx);
    psba_free(Kparas); psba_free(initrot); psba_free(camsExParas); psba_free(pts3D); psba_free(impts_data);
    psba_free(iidx); psba_free(jidx);
    return 0;
}
