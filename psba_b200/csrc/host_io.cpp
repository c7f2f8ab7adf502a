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

static int read_sba_impl(const char *camsfname, const char *ptsfname, int origin_cnp,
                         const double *Kdefault, int *ncams, int *n3Dpts, int *n2Dprojs,
                         double **Kparas, double **initrot, double **camsEx, double **pts,
                         double **imgpts, int **iidx, int **jidx, double **kc_out, double **cov_out, int *covsz_out)
{
    if (kc_out) *kc_out = nullptr;
    if (cov_out) *cov_out = nullptr;
    if (covsz_out) *covsz_out = 0;
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
    if (origin_cnp == 6 && !Kdefault) { fprintf(stderr, "psba_readInitialSBAEstimate: 7-column camera file needs Kdefault\n"); return 6; }
    double *K = (double *)malloc(sizeof(double) * m * 5), *rot = (double *)malloc(sizeof(double) * m * 4);
    double *ex = (double *)malloc(sizeof(double) * m * 6);
    double *P = nullptr;
    double *kcv = (kc_out && origin_cnp == 16) ? (double *)malloc(sizeof(double) * m * 5) : nullptr;
    // every error return below gives back what was allocated so far (the reference exits instead)
    auto fail = [&](int code) { free(K); free(rot); free(ex); free(P); free(kcv); return code; };
    std::vector<double> raw(filecnp), filt(origin_cnp);
    for (int j = 0; j < m; ++j) {
        const char *s = cl[j];
        char *e;
        for (int k = 0; k < filecnp; ++k) {
            raw[k] = strtod(s, &e);
            if (e == s) { fprintf(stderr, "readCameraParams(): line %d contains %d parameters, expected %d!\n", j + 1, k, filecnp); return fail(5); }
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
        if (kcv) for (int k = 0; k < 5; ++k) kcv[j * 5 + k] = filt[5 + k];       // varKD: columns 6-10 (quat2vec copies nin - 7 leading values)
        ex[j * 6] = ex[j * 6 + 1] = ex[j * 6 + 2] = 0.0;
        for (int k = 0; k < 3; ++k) ex[j * 6 + 3 + k] = filt[origin_cnp - 3 + k];
    }

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
    P = (double *)malloc(sizeof(double) * (size_t)n * 3);
    std::vector<double> im, cv; std::vector<int> ii, jj;
    im.reserve((size_t)n * 10); ii.reserve((size_t)n * 5); jj.reserve((size_t)n * 5);
    bool warned = false;
    for (int i = 0; i < n; ++i) {
        const char *s = pl[i];
        char *e;
        for (int k = 0; k < pnp; ++k) {
            P[(size_t)i * 3 + k] = strtod(s, &e);
            if (e == s) { fprintf(stderr, "readPointParamsAndProjections(): line %d: expecting %d parameters for 3D point\n", i, pnp); return fail(7); }
            s = e;
        }
        const long nframes = strtol(s, &e, 10);
        if (e == s) { fprintf(stderr, "readPointParamsAndProjections(): line %d: expecting number of frames\n", i); return fail(8); }
        s = e;
        const size_t first = jj.size();
        for (long f = 0; f < nframes; ++f) {
            const long frameno = strtol(s, &e, 10);
            if (e == s) { fprintf(stderr, "readPointParamsAndProjections(): line %d has fewer than %ld projections\n", i + 1, nframes); return fail(9); }
            s = e;
            if (frameno >= m || frameno < 0) {
                fprintf(stderr, "readPointParamsAndProjections(): line %d contains an image projection for frame %ld "
                                "but only %d cameras have been specified!\n", i + 1, frameno, m);
                return fail(10);
            }
            for (int k = 0; k < mnp; ++k) {
                const double v = strtod(s, &e);
                if (e == s) { fprintf(stderr, "readPointParamsAndProjections(): error reading image projections from line %d\n", i + 1); return fail(11); }
                im.push_back(v); s = e;
            }
            for (int k = 0; k < covvals; ++k) { const double v = strtod(s, &e); s = e; if (cov_out) cv.push_back(v); }   // the reference parses them and no kernel uses them
            ii.push_back(i); jj.push_back((int)frameno);
        }
        // generate_idxs scans the visibility mask with cameras ascending (misc.cpp:191-196) while the reference leaves
        // the image points in file order: a point that lists its frames out of order gets its measurements attached to
        // the wrong cameras there.  Every shipped file is ascending; for anything else the (frame, x, y)
        // tuples are sorted TOGETHER here, so that a measurement stays with its camera (documented deviation).
        if (!std::is_sorted(jj.begin() + first, jj.end())) {
            if (!warned) { fprintf(stderr, "psba: point %d lists its frames out of order; (frame, x, y) tuples are sorted by frame\n", i); warned = true; }
            const size_t cnt = jj.size() - first;
            std::vector<size_t> ord(cnt);
            for (size_t q = 0; q < cnt; ++q) ord[q] = q;
            std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return jj[first + a] < jj[first + b]; });
            std::vector<int> js(cnt); std::vector<double> ms(cnt * mnp), cs(cnt * covvals);
            const bool hc = cov_out && covvals;
            for (size_t q = 0; q < cnt; ++q) {
                js[q] = jj[first + ord[q]];
                for (int k = 0; k < mnp; ++k) ms[q * mnp + k] = im[(first + ord[q]) * mnp + k];
                if (hc) for (int k = 0; k < covvals; ++k) cs[q * covvals + k] = cv[(first + ord[q]) * covvals + k];
            }
            for (size_t q = 0; q < cnt; ++q) {
                jj[first + q] = js[q];
                for (int k = 0; k < mnp; ++k) im[(first + q) * mnp + k] = ms[q * mnp + k];
                if (hc) for (int k = 0; k < covvals; ++k) cv[(first + q) * covvals + k] = cs[q * covvals + k];
            }
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
    if (kc_out) *kc_out = kcv;
    if (cov_out && covvals && cv.size() == o * (size_t)covvals) {
        double *CV = (double *)malloc(sizeof(double) * cv.size());
        memcpy(CV, cv.data(), sizeof(double) * cv.size());
        *cov_out = CV;
        if (covsz_out) *covsz_out = covvals;
    }
    return 0;
}

extern "C" int psba_readInitialSBAEstimate(const char *camsfname, const char *ptsfname, int origin_cnp,
                                           const double *Kdefault, int *ncams, int *n3Dpts, int *n2Dprojs,
                                           double **Kparas, double **initrot, double **camsEx, double **pts,
                                           double **imgpts, int **iidx, int **jidx)
{
    return read_sba_impl(camsfname, ptsfname, origin_cnp, Kdefault, ncams, n3Dpts, n2Dprojs, Kparas, initrot, camsEx, pts, imgpts, iidx, jidx,
                         nullptr, nullptr, nullptr);
}

// the same reader, also handing out what the reference parses and drops: the distortion coefficients of a varKD camera
// file (origin_cnp = 16: kc[m*5], else NULL) and the image-point covariances (cov[o*covsz], covsz = 4 or 3, else NULL)
extern "C" int psba_readInitialSBAEstimate_ext(const char *camsfname, const char *ptsfname, int origin_cnp,
                                               const double *Kdefault, int *ncams, int *n3Dpts, int *n2Dprojs,
                                               double **Kparas, double **initrot, double **camsEx, double **pts,
                                               double **imgpts, int **iidx, int **jidx, double **kc, double **cov, int *covsz)
{
    return read_sba_impl(camsfname, ptsfname, origin_cnp, Kdefault, ncams, n3Dpts, n2Dprojs, Kparas, initrot, camsEx, pts, imgpts, iidx, jidx,
                         kc, cov, covsz);
}


// ---------------------------------------------------------------------------------------------
// Result writers.  The reference never saves its result: the writers exist only as commented-out
// prototypes (PSBA/readparams.h:13-25, PSBA/misc.cpp:60-85).  vec2quat is the inverse of the set-up of
// PSBA/main.cpp:131-149: the refined local rotation (vector part v, scalar sqrt(1-|v|^2)) is composed with
// the initial rotation, q = q_local (x) q_init, in the operation order of CL_files/compute_exQT.cl:46-49.
extern "C" void psba_vec2quat(const double *initrot4, const double *local3, double *q4)
{
    const double s0 = initrot4[0], a1 = initrot4[1], a2 = initrot4[2], a3 = initrot4[3];
    const double v1 = local3[0], v2 = local3[1], v3 = local3[2];
    const double sl = std::sqrt(1 - v1 * v1 - v2 * v2 - v3 * v3);
    q4[0] = sl * s0 - (a1 * v1 + a2 * v2 + a3 * v3);
    q4[1] = s0 * v1 + sl * a1 + a3 * v2 - a2 * v3;
    q4[2] = s0 * v2 + sl * a2 + a1 * v3 - a3 * v1;
    q4[3] = s0 * v3 + sl * a3 + a2 * v1 - a1 * v2;
}

// SBA text files that psba_readInitialSBAEstimate (origin_cnp = 11) reads back: cameras
// "fu u0 v0 ar s  q0 qx qy qz  tx ty tz", points "X Y Z nframes (frame x y)*" with the measurements the solve used
extern "C" int psba_write_sba_result(const char *camsfname, const char *ptsfname, int ncams, int n3Dpts, int n2Dprojs,
                                     const double *Kparas, const double *initrot, const double *camsEx, const double *pts,
                                     const double *imgpts, const int *iidx, const int *jidx)
{
    FILE *f = fopen(camsfname, "w");
    if (!f) { fprintf(stderr, "psba_b200: cannot open %s for writing\n", camsfname); return 1; }
    for (int j = 0; j < ncams; ++j) {
        double q[4];
        psba_vec2quat(initrot + j * 4, camsEx + j * 6, q);
        fprintf(f, "%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", Kparas[j * 5], Kparas[j * 5 + 1],
                Kparas[j * 5 + 2], Kparas[j * 5 + 3], Kparas[j * 5 + 4], q[0], q[1], q[2], q[3], camsEx[j * 6 + 3], camsEx[j * 6 + 4],
                camsEx[j * 6 + 5]);
    }
    fclose(f);
    f = fopen(ptsfname, "w");
    if (!f) { fprintf(stderr, "psba_b200: cannot open %s for writing\n", ptsfname); return 1; }
    int k = 0;
    for (int i = 0; i < n3Dpts; ++i) {
        int e = k;
        while (e < n2Dprojs && iidx[e] == i) ++e;
        fprintf(f, "%.17g %.17g %.17g %d", pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], e - k);
        for (; k < e; ++k) fprintf(f, " %d %.17g %.17g", jidx[k], imgpts[k * 2], imgpts[k * 2 + 1]);
        fprintf(f, "\n");
    }
    fclose(f);
    return 0;
}

// ASCII PLY: the points as vertices, then the camera centres C = -R(q)^T t (red)
extern "C" int psba_write_ply(const char *fname, int ncams, int n3Dpts, const double *initrot, const double *camsEx, const double *pts)
{
    FILE *f = fopen(fname, "w");
    if (!f) { fprintf(stderr, "psba_b200: cannot open %s for writing\n", fname); return 1; }
    fprintf(f, "ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
               "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n", n3Dpts + ncams);
    for (int i = 0; i < n3Dpts; ++i) fprintf(f, "%.9g %.9g %.9g 255 255 255\n", pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]);
    for (int j = 0; j < ncams; ++j) {
        double q[4];
        psba_vec2quat(initrot + j * 4, camsEx + j * 6, q);
        const double s = q[0], x = q[1], y = q[2], z = q[3];
        const double R[9] = {s * s + x * x - y * y - z * z, 2 * (x * y - s * z), 2 * (x * z + s * y),
                             2 * (x * y + s * z), s * s - x * x + y * y - z * z, 2 * (y * z - s * x),
                             2 * (x * z - s * y), 2 * (y * z + s * x), s * s - x * x - y * y + z * z};
        const double *t = camsEx + j * 6 + 3;
        fprintf(f, "%.9g %.9g %.9g 255 0 0\n", -(R[0] * t[0] + R[3] * t[1] + R[6] * t[2]), -(R[1] * t[0] + R[4] * t[1] + R[7] * t[2]),
                -(R[2] * t[0] + R[5] * t[1] + R[8] * t[2]));
    }
    fclose(f);
    return 0;
}


// ---------------------------------------------------------------------------------------------
// Native BAL loader ("Bundle Adjustment in the Large", problem-*.txt):
//   <cams> <points> <observations>
//   <camera> <point> <x> <y>                      one line per observation
//   9 numbers per camera: Rodrigues r(3), t(3), f, k1, k2        (one number per line)
//   3 numbers per point
// BAL projects p = -P/P.z, p' = f (1 + k1 |p|^2 + k2 |p|^4) p with P = R X + t.  PSBA's pinhole
// (CL_files/compute_exQT.cl:68-69) has no minus sign: the camera frame is turned by 180 degrees about x,
// R' = diag(1,-1,-1) R, t' = diag(1,-1,-1) t, and the image y axis is flipped, (x, y) -> (x, -y); then
// x = fu Xc'/Zc' with fu = f, u0 = v0 = 0, ar = 1, s = 0 -- the form of the shipped *-cams.txt files
// (SURVEY App. C).  The radial terms are returned in kc[m*2]; the reference's model ignores distortion
//  and so does the solver.  Observations are sorted point-major with ascending cameras
// (generate_idxs order).  Outputs are malloc'd as in psba_readInitialSBAEstimate.
static void rot_to_quat(const double R[9], double q[4])
{
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        const double s = std::sqrt(tr + 1.0) * 2;
        q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        const double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
        q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        const double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
        q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
    } else {
        const double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
        q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
    }
}

extern "C" int psba_read_bal(const char *fname, int *ncams, int *n3Dpts, int *n2Dprojs, double **Kparas, double **initrot,
                             double **camsEx, double **pts, double **imgpts, int **iidx, int **jidx, double **kc)
{
    std::string buf;
    if (!slurp(fname, buf)) { fprintf(stderr, "psba_b200: cannot open file %s\n", fname); return 1; }
    char *p = &buf[0];
    auto next_d = [&](double &v) -> bool { char *e; v = strtod(p, &e); if (e == p) return false; p = e; return true; };
    double hm, hn, ho;
    if (!next_d(hm) || !next_d(hn) || !next_d(ho) || hm < 1 || hn < 1 || ho < 1) { fprintf(stderr, "psba_b200: %s: bad BAL header\n", fname); return 2; }
    const int m = (int)hm, n = (int)hn, o = (int)ho;
    struct obs { int i, j; double x, y; };
    std::vector<obs> ob((size_t)o);
    for (int k = 0; k < o; ++k) {
        double cj, pi, x, y;
        if (!next_d(cj) || !next_d(pi) || !next_d(x) || !next_d(y) || cj < 0 || cj >= m || pi < 0 || pi >= n) {
            fprintf(stderr, "psba_b200: %s: bad observation %d\n", fname, k); return 3;
        }
        ob[k] = {(int)pi, (int)cj, x, -y};
    }
    std::stable_sort(ob.begin(), ob.end(), [](const obs &a, const obs &b) { return a.i != b.i ? a.i < b.i : a.j < b.j; });
    for (int k = 1; k < o; ++k)
        if (ob[k].i == ob[k - 1].i && ob[k].j == ob[k - 1].j) { fprintf(stderr, "psba_b200: %s: point %d is observed twice by camera %d\n", fname, ob[k].i, ob[k].j); return 4; }
    double *K = (double *)malloc(sizeof(double) * m * 5), *rot = (double *)malloc(sizeof(double) * m * 4);
    double *ex = (double *)malloc(sizeof(double) * m * 6), *X = (double *)malloc(sizeof(double) * n * 3);
    double *im = (double *)malloc(sizeof(double) * o * 2), *dist = (double *)malloc(sizeof(double) * m * 2);
    int *ii = (int *)malloc(sizeof(int) * o), *jj = (int *)malloc(sizeof(int) * o);
    auto fail = [&](int code) { free(K); free(rot); free(ex); free(X); free(im); free(dist); free(ii); free(jj); return code; };
    for (int j = 0; j < m; ++j) {
        double c9[9];
        for (int q = 0; q < 9; ++q) if (!next_d(c9[q])) { fprintf(stderr, "psba_b200: %s: camera %d is truncated\n", fname, j); return fail(5); }
        // Rodrigues vector -> rotation matrix
        const double th = std::sqrt(c9[0] * c9[0] + c9[1] * c9[1] + c9[2] * c9[2]);
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (th > 1e-300) {
            const double kx = c9[0] / th, ky = c9[1] / th, kz = c9[2] / th, cs = std::cos(th), sn = std::sin(th), v = 1 - cs;
            R[0] = cs + kx * kx * v; R[1] = kx * ky * v - kz * sn; R[2] = kx * kz * v + ky * sn;
            R[3] = ky * kx * v + kz * sn; R[4] = cs + ky * ky * v; R[5] = ky * kz * v - kx * sn;
            R[6] = kz * kx * v - ky * sn; R[7] = kz * ky * v + kx * sn; R[8] = cs + kz * kz * v;
        }
        for (int q = 3; q < 9; ++q) R[q] = -R[q];                   // R' = diag(1,-1,-1) R
        double q4[4];
        rot_to_quat(R, q4);
        const double in[7] = {q4[0], q4[1], q4[2], q4[3], c9[3], -c9[4], -c9[5]};
        double out6[6];
        psba_quat2vec(in, 7, out6, 6);                              // normalise, scalar part >= 0, keep the vector part
        const double vv = out6[0] * out6[0] + out6[1] * out6[1] + out6[2] * out6[2];
        rot[j * 4] = std::sqrt(1.0 - vv); rot[j * 4 + 1] = out6[0]; rot[j * 4 + 2] = out6[1]; rot[j * 4 + 3] = out6[2];   // readparams.cpp:222-226
        ex[j * 6] = ex[j * 6 + 1] = ex[j * 6 + 2] = 0.0;            // local rotation := 0 (main.cpp:131-149)
        ex[j * 6 + 3] = out6[3]; ex[j * 6 + 4] = out6[4]; ex[j * 6 + 5] = out6[5];
        K[j * 5] = c9[6]; K[j * 5 + 1] = 0; K[j * 5 + 2] = 0; K[j * 5 + 3] = 1; K[j * 5 + 4] = 0;
        dist[j * 2] = c9[7]; dist[j * 2 + 1] = c9[8];
    }
    for (int i = 0; i < n * 3; ++i) if (!next_d(X[i])) { fprintf(stderr, "psba_b200: %s: point %d is truncated\n", fname, i / 3); return fail(6); }
    for (int k = 0; k < o; ++k) { ii[k] = ob[k].i; jj[k] = ob[k].j; im[k * 2] = ob[k].x; im[k * 2 + 1] = ob[k].y; }
    *ncams = m; *n3Dpts = n; *n2Dprojs = o;
    *Kparas = K; *initrot = rot; *camsEx = ex; *pts = X; *imgpts = im; *iidx = ii; *jidx = jj;
    if (kc) *kc = dist; else free(dist);
    return 0;
}
