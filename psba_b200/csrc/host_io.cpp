// host_io.cpp -- SBA-format text loaders: the host-side I/O PSBA keeps (north star).
//   psba_readInitialSBAEstimate   <- readInitialSBAEstimate, PSBA/readparams.cpp:444-518
//   psba_quat2vec                 <- quat2vec, PSBA/misc.cpp:21-49
// plus the K | extrinsics split and the zeroing of the local rotation of PSBA/main.cpp:131-149 and
// the observation index lists of generate_idxs (PSBA/misc.cpp:178-218; iidx/jidx only -- the dense
// blk_idx / comm3DIdx tables are replaced by CSR lists built inside the engine).
// The files are parsed in one pass from memory instead of the reference's two fscanf passes; the
// dense visibility mask vmask[n*m] (2 GB at 1M points x 2000 cameras) is never built.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/psba_b200.h"

extern "C" void psba_free(void *p) { free(p); }

extern "C" void psba_quat2vec(const double *inp, int nin, double *outp, int nout)
{
    int i = 0;
    if (nin > 7) for (; i < nin - 7; ++i) outp[i] = inp[i];       // intrinsics (+ distortion)
    // normalise; q and -q are the same rotation: make the scalar part non-negative, keep the vector part
    double mag = std::sqrt(inp[i] * inp[i] + inp[i + 1] * inp[i + 1] + inp[i + 2] * inp[i + 2] + inp[i + 3] * inp[i + 3]);
    const double sg = (inp[i] >= 0.0) ? 1.0 : -1.0;
    mag = sg / mag;
    outp[i] = inp[i + 1] * mag;
    outp[i + 1] = inp[i + 2] * mag;
    outp[i + 2] = inp[i + 3] * mag;
    i += 3;
    for (; i < nout; ++i) outp[i] = inp[i + 1];                   // translation
}

static bool slurp(const char *name, std::string &out)
{
    FILE *f = fopen(name, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize((size_t)sz);
    size_t got = sz ? fread(&out[0], 1, (size_t)sz, f) : 0;
    fclose(f);
    return got == (size_t)sz;
}

// split into non-comment, non-empty lines (a line whose first character is '#' is a comment,
// readparams.cpp:36-40)
static void data_lines(std::string &buf, std::vector<char *> &lines)
{
    char *p = &buf[0], *end = p + buf.size();
    while (p < end) {
        char *nl = (char *)memchr(p, '\n', end - p);
        char *e = nl ? nl : end;
        if (nl) *nl = 0;
        if (*p != '#') {
            char *q = p;
            while (q < e && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
            if (q < e) lines.push_back(p);
        }
        p = e + 1;
    }
}

static int count_doubles(const char *s)
{
    int n = 0;
    char *e;
    for (;;) { strtod(s, &e); if (e == s) break; ++n; s = e; }
    return n;
}

extern "C" int psba_readInitialSBAEstimate(const char *camsfname, const char *ptsfname, int origin_cnp,
                                           const double *Kdefault, int *ncams, int *n3Dpts, int *n2Dprojs,
                                           double **Kparas, double **initrot, double **camsEx, double **pts,
                                           double **imgpts, int **iidx, int **jidx)
{
    if (origin_cnp != 6 && origin_cnp != 11 && origin_cnp != 16) {
        fprintf(stderr, "psba_readInitialSBAEstimate: origin_cnp must be 6, 11 or 16\n");
        return 1;
    }
    const int filecnp = origin_cnp + 1, pnp = 3, mnp = 2;
    std::string cbuf, pbuf;
    if (!slurp(camsfname, cbuf)) { fprintf(stderr, "cannot open file %s\n", camsfname); return 2; }
    if (!slurp(ptsfname, pbuf)) { fprintf(stderr, "cannot open file %s\n", ptsfname); return 2; }
    std::vector<char *> cl, pl;
    data_lines(cbuf, cl);
    data_lines(pbuf, pl);
    const int m = (int)cl.size(), n = (int)pl.size();
    if (m == 0 || n == 0) { fprintf(stderr, "psba_readInitialSBAEstimate: empty input\n"); return 3; }
    // readparams.cpp:189-192: only the first line's column count is validated
    {
        const int nf = count_doubles(cl[0]);
        if (nf != filecnp) {
            fprintf(stderr, "readCameraParams(): expected %d camera parameters, first line contains %d!\n", filecnp, nf);
            return 4;
        }
    }
    double *K = (double *)malloc(sizeof(double) * m * 5), *rot = (double *)malloc(sizeof(double) * m * 4);
    double *ex = (double *)malloc(sizeof(double) * m * 6);
    std::vector<double> raw(filecnp), filt(origin_cnp);
    for (int j = 0; j < m; ++j) {
        const char *s = cl[j];
        char *e;
        for (int k = 0; k < filecnp; ++k) {
            raw[k] = strtod(s, &e);
            if (e == s) { fprintf(stderr, "readCameraParams(): line %d contains %d parameters, expected %d!\n", j + 1, k, filecnp); return 5; }
            s = e;
        }
        psba_quat2vec(raw.data(), filecnp, filt.data(), origin_cnp);
        const double *qv = &filt[origin_cnp - 6];
        // readparams.cpp:222-226: the scalar part is RECOMPUTED from the filtered vector part
        rot[j * 4 + 1] = qv[0]; rot[j * 4 + 2] = qv[1]; rot[j * 4 + 3] = qv[2];
        rot[j * 4] = std::sqrt(1.0 - qv[0] * qv[0] - qv[1] * qv[1] - qv[2] * qv[2]);
        // main.cpp:131-149: K | (local rotation := 0) | t ; distortion columns (varKD) are dropped,
        // the reference has no distortion model
        for (int k = 0; k < 5; ++k) K[j * 5 + k] = origin_cnp >= 11 ? filt[k] : (Kdefault ? Kdefault[k] : 0.0);
        ex[j * 6] = ex[j * 6 + 1] = ex[j * 6 + 2] = 0.0;
        for (int k = 0; k < 3; ++k) ex[j * 6 + 3 + k] = filt[origin_cnp - 3 + k];
    }
    if (origin_cnp == 6 && !Kdefault) { fprintf(stderr, "psba_readInitialSBAEstimate: 7-column camera file needs Kdefault\n"); return 6; }

    // points: X Y Z nframes (frame x y [cov])*   (readparams.cpp:247-290, 332-423)
    int covvals = 0;
    {
        const int nfirst = count_doubles(pl[0]);
        char *e; const char *s = pl[0];
        for (int k = 0; k < pnp; ++k) { strtod(s, &e); s = e; }
        const int nframes = (int)strtol(s, &e, 10);
        const int rest = nfirst - (pnp + 1);
        if (rest == nframes * (mnp + 1 + mnp * mnp)) covvals = mnp * mnp;
        else if (rest == nframes * (mnp + 1 + mnp * (mnp + 1) / 2)) covvals = mnp * (mnp + 1) / 2;
    }
    double *P = (double *)malloc(sizeof(double) * (size_t)n * 3);
    std::vector<double> im; std::vector<int> ii, jj;
    im.reserve((size_t)n * 10); ii.reserve((size_t)n * 5); jj.reserve((size_t)n * 5);
    bool warned = false;
    for (int i = 0; i < n; ++i) {
        const char *s = pl[i];
        char *e;
        for (int k = 0; k < pnp; ++k) {
            P[(size_t)i * 3 + k] = strtod(s, &e);
            if (e == s) { fprintf(stderr, "readPointParamsAndProjections(): line %d: expecting %d parameters for 3D point\n", i, pnp); return 7; }
            s = e;
        }
        const long nframes = strtol(s, &e, 10);
        if (e == s) { fprintf(stderr, "readPointParamsAndProjections(): line %d: expecting number of frames\n", i); return 8; }
        s = e;
        const size_t first = jj.size();
        for (long f = 0; f < nframes; ++f) {
            const long frameno = strtol(s, &e, 10);
            if (e == s) { fprintf(stderr, "readPointParamsAndProjections(): line %d has fewer than %ld projections\n", i + 1, nframes); return 9; }
            s = e;
            if (frameno >= m || frameno < 0) {
                fprintf(stderr, "readPointParamsAndProjections(): line %d contains an image projection for frame %ld "
                                "but only %d cameras have been specified!\n", i + 1, frameno, m);
                return 10;
            }
            for (int k = 0; k < mnp; ++k) {
                const double v = strtod(s, &e);
                if (e == s) { fprintf(stderr, "readPointParamsAndProjections(): error reading image projections from line %d\n", i + 1); return 11; }
                im.push_back(v); s = e;
            }
            for (int k = 0; k < covvals; ++k) { strtod(s, &e); s = e; }   // covariances are parsed and dropped (never used by any kernel)
            ii.push_back(i); jj.push_back((int)frameno);
        }
        // generate_idxs scans the visibility mask with cameras ascending (misc.cpp:191-196) while the
        // image points stay in file order; identical when the frames of a point are listed ascending
        if (!std::is_sorted(jj.begin() + first, jj.end())) {
            if (!warned) { fprintf(stderr, "psba: point %d lists its frames out of order; indices follow generate_idxs (ascending)\n", i); warned = true; }
            std::sort(jj.begin() + first, jj.end());
        }
    }
    const size_t o = jj.size();
    double *IM = (double *)malloc(sizeof(double) * o * 2);
    int *I = (int *)malloc(sizeof(int) * o), *J = (int *)malloc(sizeof(int) * o);
    memcpy(IM, im.data(), sizeof(double) * o * 2);
    memcpy(I, ii.data(), sizeof(int) * o);
    memcpy(J, jj.data(), sizeof(int) * o);
    *ncams = m; *n3Dpts = n; *n2Dprojs = (int)o;
    *Kparas = K; *initrot = rot; *camsEx = ex; *pts = P; *imgpts = IM; *iidx = I; *jidx = J;
    return 0;
}
