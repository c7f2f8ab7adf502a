/*
 * orc_kernels.c -- ORACLE (test infrastructure): restatement of the per-observation,
 * block-assembly, Schur and back-substitution kernels in the reference directory CL_files/.
 *
 * Each function cites the kernel it follows.  Loop nests are the kernels' NDRanges
 * (PSBA/sba_func.cpp global work sizes) executed serially, or with OpenMP across
 * independent work-items when s->nthreads > 1 (every work-item owns its output element,
 * so the result does not depend on the thread count).
 *
 * The Jacobian is an independent analytic derivation (quaternion calculus) of the same
 * quantity that the machine-generated code in compute_jacobiQT.cl:7-141 evaluates; it is
 * checked against that code (through oracle/_ref) to ~1e-13 relative in tests/.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "psba_oracle.h"

#ifdef _OPENMP
#include <omp.h>
#define PFOR _Pragma("omp parallel for schedule(static) num_threads(nt)")
#else
#define PFOR
#endif

/* q = ql (x) q0 with ql = (sqrt(1-|v|^2), v): compute_exQT.cl:33-49 */
static void total_quat(const double *q0, const double *v, double *q)
{
    double si = sqrt(1 - v[0] * v[0] - v[1] * v[1] - v[2] * v[2]);
    q[0] = si * q0[0] - (q0[1] * v[0] + q0[2] * v[1] + q0[3] * v[2]);
    q[1] = q0[0] * v[0] + si * q0[1] + q0[3] * v[1] - q0[2] * v[2];
    q[2] = q0[0] * v[1] + si * q0[2] + q0[1] * v[2] - q0[3] * v[0];
    q[3] = q0[0] * v[2] + si * q0[3] + q0[2] * v[0] - q0[1] * v[1];
}

/* Xc = q X q* + t evaluated as two quaternion products, compute_exQT.cl:51-65 */
static void transform_point(const double *q, const double *t, const double *X, double *Xc)
{
    double s = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
    double b0 = -X[0] * w1 - w2 * X[1] - w3 * X[2];          /* scalar part of q (x) X  (t11) */
    double a1 = s * X[0] + w2 * X[2] - w3 * X[1];            /* vector part             (t17) */
    double a2 = X[1] * s + w3 * X[0] - X[2] * w1;            /*                         (t22) */
    double a3 = s * X[2] + X[1] * w1 - w2 * X[0];            /*                         (t27) */
    Xc[0] = -w1 * b0 + s * a1 - a2 * w3 + w2 * a3 + t[0];
    Xc[1] = -w2 * b0 + s * a2 - a3 * w1 + w3 * a1 + t[1];
    Xc[2] = -b0 * w3 + s * a3 - w2 * a1 + w1 * a2 + t[2];
}

/* Extended projection (not in the reference: SURVEY F7, 8(f) rank 3): the distortion of Lourakis' sba "varKD"
 * camera (Bouguet model, kc = k1 k2 p1 p2 k3) applied to the normalised point, then the reference's affine K:
 *   xn = Xc/Zc, yn = Yc/Zc, r2 = xn^2 + yn^2, c = 1 + k1 r2 + k2 r2^2 + k3 r2^3,
 *   xd = c xn + 2 p1 xn yn + p2 (r2 + 2 xn^2),   yd = c yn + p1 (r2 + 2 yn^2) + 2 p2 xn yn,
 *   x = fu xd + s yd + u0,   y = fu ar yd + v0        (kc = 0: compute_exQT.cl:68-69).
 * P (optional) receives d(x, y) / d Xc. */
static void ext_project(const double *K, const double *kc, const double *Xc, double *px, double *py, double P[2][3])
{
    static const double zero5[5] = {0, 0, 0, 0, 0};
    const double *k = kc ? kc : zero5;
    double iz = 1 / Xc[2], xn = Xc[0] * iz, yn = Xc[1] * iz, r2 = xn * xn + yn * yn;
    double c = 1 + r2 * (k[0] + r2 * (k[1] + r2 * k[4])), dc = k[0] + r2 * (2 * k[1] + 3 * k[4] * r2);
    double xd = c * xn + 2 * k[2] * xn * yn + k[3] * (r2 + 2 * xn * xn);
    double yd = c * yn + k[2] * (r2 + 2 * yn * yn) + 2 * k[3] * xn * yn;
    double fa = K[0] * K[3];
    *px = K[0] * xd + K[4] * yd + K[1];
    *py = fa * yd + K[2];
    if (P) {
        double j00 = c + 2 * dc * xn * xn + 2 * k[2] * yn + 6 * k[3] * xn, j01 = 2 * dc * xn * yn + 2 * k[2] * xn + 2 * k[3] * yn;
        double j10 = j01, j11 = c + 2 * dc * yn * yn + 6 * k[2] * yn + 2 * k[3] * xn;
        double g00 = K[0] * j00 + K[4] * j10, g01 = K[0] * j01 + K[4] * j11, g10 = fa * j10, g11 = fa * j11;
        P[0][0] = g00 * iz; P[0][1] = g01 * iz; P[0][2] = -(g00 * xn + g01 * yn) * iz;
        P[1][0] = g10 * iz; P[1][1] = g11 * iz; P[1][2] = -(g10 * xn + g11 * yn) * iz;
    }
}

/* compute_exQT.cl:18-71 ; NDRange {o} (sba_func.cpp:115) */
static void k_exQT(orc_state *s, const double *cams, const double *pts, double *ex)
{
    int o = s->o, idx, nt = s->nthreads; (void)nt;
    PFOR
    for (idx = 0; idx < o; ++idx) {
        int i = s->iidx[idx], j = s->jidx[idx];
        const double *K = s->K + j * 5;
        double q[4], Xc[3], inv;
        total_quat(s->initcams + j * 4, cams + j * 6, q);
        transform_point(q, cams + j * 6 + 3, pts + (size_t)i * 3, Xc);
        if (s->kc || s->wgt) {
            double px, py, e0, e1;
            ext_project(K, s->kc ? s->kc + j * 5 : 0, Xc, &px, &py, 0);
            e0 = s->impts[idx * 2] - px; e1 = s->impts[idx * 2 + 1] - py;
            if (s->wgt) { const double *w = s->wgt + (size_t)idx * 3; e1 = w[1] * e0 + w[2] * e1; e0 = w[0] * e0; }
            ex[idx * 2] = e0; ex[idx * 2 + 1] = e1;
            continue;
        }
        inv = 1 / Xc[2];
        /* x = (fu*Xc + s*Yc + u0*Zc)/Zc ; y = (fu*ar*Yc + v0*Zc)/Zc  (compute_exQT.cl:68-69) */
        ex[idx * 2] = s->impts[idx * 2] - (K[0] * Xc[0] + K[4] * Xc[1] + K[1] * Xc[2]) * inv;
        ex[idx * 2 + 1] = s->impts[idx * 2 + 1] - (K[0] * K[3] * Xc[1] + K[2] * Xc[2]) * inv;
    }
}

/* quaternion product r = a (x) b */
static void qmul(const double *a, const double *b, double *r)
{
    r[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    r[1] = a[0] * b[1] + b[0] * a[1] + a[2] * b[3] - a[3] * b[2];
    r[2] = a[0] * b[2] + b[0] * a[2] + a[3] * b[1] - a[1] * b[3];
    r[3] = a[0] * b[3] + b[0] * a[3] + a[1] * b[2] - a[2] * b[1];
}

/* compute_jacobiQT.cl:7-141 ; NDRange {o} (sba_func.cpp:187).
 * A_ij = d proj / d (v, t)  (2x6, row-major JA[idx*12 + r*6 + c]),
 * B_ij = d proj / d X       (2x3, row-major JB[idx*6  + r*3 + c]).
 * With Xc = q X q* + t and X pure, d(q X q*)/d theta = 2 vec(dq (x) X (x) q*);
 * dq/dv_k = (-v_k/sl, e_k) (x) q0;  d(q X q*)/dX = M(q) (no unit-norm assumption, as the
 * reference's JB: compute_jacobiQT.cl:117-140). */
static void k_jacobiQT(orc_state *s)
{
    int o = s->o, idx, nt = s->nthreads; (void)nt;
    PFOR
    for (idx = 0; idx < o; ++idx) {
        int i = s->iidx[idx], j = s->jidx[idx], k, c;
        const double *K = s->K + j * 5, *q0 = s->initcams + j * 4, *v = s->cams + j * 6;
        const double *X = s->pts + (size_t)i * 3;
        double *JA = s->JA + (size_t)idx * 12, *JB = s->JB + (size_t)idx * 6;
        double q[4], Xc[3], b[4], qc[4], Xq[4] = {0, X[0], X[1], X[2]};
        double sl = sqrt(1.0 - v[0] * v[0] - v[1] * v[1] - v[2] * v[2]);
        double P[2][3], iz, iz2, D[3][3], M[3][3];
        total_quat(q0, v, q);
        transform_point(q, v + 3, X, Xc);
        iz = 1 / Xc[2]; iz2 = 1 / (Xc[2] * Xc[2]);
        /* d proj / d Xc */
        P[0][0] = K[0] * iz; P[0][1] = K[4] * iz; P[0][2] = -(K[0] * Xc[0] + K[4] * Xc[1]) * iz2;
        P[1][0] = 0.0;       P[1][1] = K[0] * K[3] * iz; P[1][2] = -(K[0] * K[3] * Xc[1]) * iz2;
        if (s->kc || s->wgt) {
            double px, py;
            ext_project(K, s->kc ? s->kc + j * 5 : 0, Xc, &px, &py, P);
            if (s->wgt) {       /* rows of the weighted Jacobian: W P */
                const double *w = s->wgt + (size_t)idx * 3;
                for (c = 0; c < 3; ++c) { P[1][c] = w[1] * P[0][c] + w[2] * P[1][c]; P[0][c] = w[0] * P[0][c]; }
            }
        }
        /* b = X (x) q* */
        qc[0] = q[0]; qc[1] = -q[1]; qc[2] = -q[2]; qc[3] = -q[3];
        qmul(Xq, qc, b);
        for (k = 0; k < 3; ++k) {
            double dql[4] = {-v[k] / sl, 0, 0, 0}, dq[4], r[4];
            dql[1 + k] = 1.0;
            qmul(dql, q0, dq);
            qmul(dq, b, r);
            D[0][k] = 2 * r[1]; D[1][k] = 2 * r[2]; D[2][k] = 2 * r[3];
        }
        {
            double s0 = q[0], x = q[1], y = q[2], z = q[3];
            M[0][0] = s0 * s0 + x * x - y * y - z * z; M[0][1] = 2 * (x * y - s0 * z); M[0][2] = 2 * (x * z + s0 * y);
            M[1][0] = 2 * (x * y + s0 * z); M[1][1] = s0 * s0 - x * x + y * y - z * z; M[1][2] = 2 * (y * z - s0 * x);
            M[2][0] = 2 * (x * z - s0 * y); M[2][1] = 2 * (y * z + s0 * x); M[2][2] = s0 * s0 - x * x - y * y + z * z;
        }
        for (k = 0; k < 2; ++k) {
            for (c = 0; c < 3; ++c) {
                JA[k * 6 + c] = P[k][0] * D[0][c] + P[k][1] * D[1][c] + P[k][2] * D[2][c];
                JA[k * 6 + 3 + c] = P[k][c];
                JB[k * 3 + c] = P[k][0] * M[0][c] + P[k][1] * M[1][c] + P[k][2] * M[2][c];
            }
        }
    }
}

/* compute_U.cl:5-35 ; NDRange {6, 6m} (sba_func.cpp:283): sum over i ascending */
static void k_U(orc_state *s, double coeff)
{
    int m = s->m, j, nt = s->nthreads; (void)nt;
    PFOR
    for (j = 0; j < m; ++j) {
        int r, c, a;
        for (r = 0; r < 6; ++r) for (c = 0; c < 6; ++c) {
            double sum = 0;
            for (a = s->cam_ptr[j]; a < s->cam_ptr[j + 1]; ++a) {
                const double *JA = s->JA + (size_t)s->cam_obs[a] * 12;
                sum = sum + JA[r] * JA[c] + JA[6 + r] * JA[6 + c];
            }
            sum = coeff * sum;
            s->U[j * 36 + r * 6 + c] = sum;
            if (r == c) s->UVdiag[j * 6 + r] = sum;
        }
    }
}

/* compute_V.cl:6-38 ; NDRange {3, 3n} (sba_func.cpp:369): sum over j ascending */
static void k_V(orc_state *s, double coeff)
{
    int n = s->n, i, nt = s->nthreads; (void)nt;
    PFOR
    for (i = 0; i < n; ++i) {
        int r, c, a;
        for (r = 0; r < 3; ++r) for (c = 0; c < 3; ++c) {
            double sum = 0;
            for (a = s->pt_ptr[i]; a < s->pt_ptr[i + 1]; ++a) {
                const double *JB = s->JB + (size_t)a * 6;
                sum = sum + JB[r] * JB[c] + JB[3 + r] * JB[3 + c];
            }
            sum = coeff * sum;
            s->V[(size_t)i * 9 + r * 3 + c] = sum;
            if (r == c) s->UVdiag[s->N + i * 3 + r] = sum;
        }
    }
}

/* compute_Wblks.cl:7-34 ; NDRange {6, 3o} (sba_func.cpp:485) */
static void k_Wblks(orc_state *s, double coeff)
{
    int o = s->o, idx, nt = s->nthreads; (void)nt;
    PFOR
    for (idx = 0; idx < o; ++idx) {
        const double *JA = s->JA + (size_t)idx * 12, *JB = s->JB + (size_t)idx * 6;
        int r, c;
        for (r = 0; r < 6; ++r) for (c = 0; c < 3; ++c) {
            double sum = JA[r] * JB[c] + JA[6 + r] * JB[3 + c];
            s->W[(size_t)idx * 18 + r * 3 + c] = coeff * sum;
        }
    }
}

/* compute_g.cl:6-60 ; NDRange {T} (sba_func.cpp:569) */
static void k_g(orc_state *s, double coeff)
{
    int T = s->T, N = s->N, tr, nt = s->nthreads; (void)nt;
    PFOR
    for (tr = 0; tr < T; ++tr) {
        double sum = 0;
        int a;
        if (tr < N) {
            int j = tr / 6, k = tr - j * 6;
            for (a = s->cam_ptr[j]; a < s->cam_ptr[j + 1]; ++a) {
                int idx = s->cam_obs[a];
                sum = sum + s->JA[(size_t)idx * 12 + k] * s->ex[idx * 2] + s->JA[(size_t)idx * 12 + 6 + k] * s->ex[idx * 2 + 1];
            }
        } else {
            int i = (tr - N) / 3, k = tr - N - i * 3;
            for (a = s->pt_ptr[i]; a < s->pt_ptr[i + 1]; ++a)
                sum = sum + s->JB[(size_t)a * 6 + k] * s->ex[a * 2] + s->JB[(size_t)a * 6 + 3 + k] * s->ex[a * 2 + 1];
        }
        s->g[tr] = coeff * sum;
    }
}

/* update_UV.cl:5-31 ; NDRange {T} (sba_func.cpp:641) */
static void k_update_UV(orc_state *s, double mu)
{
    int j, r;
    for (j = 0; j < s->m; ++j) for (r = 0; r < 6; ++r) s->U[j * 36 + r * 7] = s->U[j * 36 + r * 7] + mu;
    for (j = 0; j < s->n; ++j) for (r = 0; r < 3; ++r) s->V[(size_t)j * 9 + r * 4] = s->V[(size_t)j * 9 + r * 4] + mu;
}

/* restore_UVdiag.cl:2-26 ; NDRange {T} (sba_func.cpp:709) */
static void k_restore_UVdiag(orc_state *s)
{
    int j, r;
    for (j = 0; j < s->m; ++j) for (r = 0; r < 6; ++r) s->U[j * 36 + r * 7] = s->UVdiag[j * 6 + r];
    for (j = 0; j < s->n; ++j) for (r = 0; r < 3; ++r) s->V[(size_t)j * 9 + r * 4] = s->UVdiag[s->N + j * 3 + r];
}

/* compute_Vinv.cl:6-90 ; NDRange {n} (sba_func.cpp:748).  Adjugate inverse of the symmetric
 * 3x3 block read from its upper triangle, written to the lower triangle + diagonal; the
 * strict upper triangle keeps V.  Returns 1.0 if the last-written status was
 * "singular" -- the reference's *ret is written by every work-item, the host reads whatever
 * was written last; we report "any". */
static double k_Vinv(orc_state *s)
{
    int n = s->n, i, any = 0;
    for (i = 0; i < n; ++i) {
        double *V = s->V + (size_t)i * 9;
        double a[3][3], tmp, T;
        double a11, a12, a13, a21, a22, a23, a31, a32, a33;
        a11 = a[0][0] = V[0]; a12 = a[0][1] = V[1]; a13 = a[0][2] = V[2];
        a21 = a[1][0] = V[3]; a22 = a[1][1] = V[4]; a23 = a[1][2] = V[5];
        a31 = a[2][0] = V[6]; a32 = a[2][1] = V[7]; a33 = a[2][2] = V[8];
        T = (a33 * a12 * a12 - 2 * a12 * a13 * a23 + a22 * a13 * a13 + a11 * a23 * a23 - a11 * a22 * a33);
        if (fabs(T) < 1e-16) {  /* compute_Vinv.cl:31-73, LU-pivot determinant fallback */
            int max_idx = 0, c;
            any = 1;
            if (a[0][0] < a[1][0]) max_idx = 1;
            if (a[max_idx][0] < a[2][0]) max_idx = 2;
            if (max_idx != 0) for (c = 0; c < 3; ++c) { tmp = a[0][c]; a[0][c] = a[max_idx][c]; a[max_idx][c] = tmp; }
            /* "if (a[0, 0] != 0)" is a comma expression: tests the row pointer, always true */
            a[1][0] = a[1][0] / a[0][0]; a[2][0] = a[2][0] / a[0][0];
            a[1][1] = a[1][1] - a[1][0] * a[0][1]; a[1][2] = a[1][2] - a[1][0] * a[0][2];
            a[2][1] = a[2][1] - a[2][0] * a[0][1]; a[2][2] = a[2][2] - a[2][0] * a[0][2];
            if (a[1][1] < a[2][1]) for (c = 0; c < 3; ++c) { tmp = a[1][c]; a[1][c] = a[2][c]; a[2][c] = tmp; }
            if (a[1][1] != 0.0) { a[2][1] = a[2][1] / a[1][1]; a[2][2] = a[2][2] - a[2][1] * a[1][2]; }
            T = a[0][0] * a[1][1] * a[2][2];
            V[0] = (a22 * a33 - a23 * a32) / T;
            V[3] = -(a21 * a33 - a23 * a31) / T;
            V[4] = (a11 * a33 - a13 * a31) / T;
            V[6] = (a21 * a32 - a22 * a31) / T;
            V[7] = -(a11 * a32 - a12 * a31) / T;
            V[8] = (a11 * a22 - a12 * a21) / T;
            continue;
        }
        V[0] = -(-a23 * a23 + a22 * a33) / T;
        V[3] = -(a13 * a23 - a12 * a33) / T;
        V[4] = -(-a13 * a13 + a11 * a33) / T;
        V[6] = -(a12 * a23 - a13 * a22) / T;
        V[7] = -(a12 * a13 - a11 * a23) / T;
        V[8] = -(-a12 * a12 + a11 * a22) / T;
    }
    s->ret = any ? 1.0 : 0.0;
    return s->ret;
}

/* compute_Yblks.cl:6-39 ; NDRange {6, 3o} (sba_func.cpp:822): Y = W * Vinv, mixed-triangle read */
static void k_Yblks(orc_state *s)
{
    int o = s->o, idx, nt = s->nthreads; (void)nt;
    PFOR
    for (idx = 0; idx < o; ++idx) {
        const double *Vi = s->V + (size_t)s->iidx[idx] * 9;
        int r, c;
        for (r = 0; r < 6; ++r) {
            const double *W = s->W + (size_t)idx * 18 + r * 3;
            double *Y = s->Y + (size_t)idx * 18 + r * 3;
            for (c = 0; c < 3; ++c) {
                if (c > 0) Y[c] = W[0] * Vi[c * 3] + W[1] * Vi[c * 3 + 1] + W[2] * Vi[2 * 3 + c];
                else Y[c] = W[0] * Vi[0] + W[1] * Vi[3] + W[2] * Vi[6];
            }
        }
    }
}

/* compute_S.cl:6-78 ; NDRange {N, N} (sba_func.cpp:898): every element of the dense S is an
 * independent sum over the common points of (k,l) in ascending order. */
static void k_S(orc_state *s)
{
    int m = s->m, N = s->N, kl, nt = s->nthreads; (void)nt;
    PFOR
    for (kl = 0; kl < m * m; ++kl) {
        int k = kl / m, l = kl - k * m, r, c;
        long long q0 = s->pair_ptr[kl], q1 = s->pair_ptr[kl + 1], q;
        for (r = 0; r < 6; ++r) for (c = 0; c < 6; ++c) {
            double sum = 0;
            for (q = q0; q < q1; ++q) {
                const double *Y = s->Y + (size_t)s->pair_oa[q] * 18 + r * 3;
                const double *W = s->W + (size_t)s->pair_ob[q] * 18 + c * 3;
                sum += Y[0] * W[0] + Y[1] * W[1] + Y[2] * W[2];   /* dot(double3,double3) */
            }
            if (k == l) sum = s->U[k * 36 + r * 6 + c] - sum;
            else sum = -sum;
            s->S[(size_t)(k * 6 + r) * N + l * 6 + c] = sum;
        }
    }
}

/* compute_ea.cl:6-37 ; NDRange {N} (sba_func.cpp:965) */
static void k_ea(orc_state *s)
{
    int N = s->N, tr, nt = s->nthreads; (void)nt;
    PFOR
    for (tr = 0; tr < N; ++tr) {
        int j = tr / 6, r = tr - j * 6, a;
        double sum = 0;
        for (a = s->cam_ptr[j]; a < s->cam_ptr[j + 1]; ++a) {
            int idx = s->cam_obs[a];
            const double *Y = s->Y + (size_t)idx * 18 + r * 3;
            const double *gb = s->g + N + (size_t)s->iidx[idx] * 3;
            sum = sum + Y[0] * gb[0] + Y[1] * gb[1] + Y[2] * gb[2];
        }
        s->eab[tr] = s->g[tr] - sum;
    }
}

/* matVec_mul.cl:7-17 ; NDRange {N} (cl_linearalg.cpp:31): dp[0..N) = S * eab[0..N) */
static void k_matVec(orc_state *s)
{
    int N = s->N, i, nt = s->nthreads; (void)nt;
    PFOR
    for (i = 0; i < N; ++i) {
        double sum = 0;
        int k;
        for (k = 0; k < N; ++k) sum = sum + s->S[(size_t)i * N + k] * s->eab[k];
        s->dp[i] = sum;
    }
}

/* compute_eb.cl:6-41 ; NDRange {3n} (sba_func.cpp:1030) */
static void k_eb(orc_state *s)
{
    int n = s->n, N = s->N, i, nt = s->nthreads; (void)nt;
    PFOR
    for (i = 0; i < n; ++i) {
        int c, a, k;
        for (c = 0; c < 3; ++c) {
            double sum = 0;
            for (a = s->pt_ptr[i]; a < s->pt_ptr[i + 1]; ++a) {
                const double *W = s->W + (size_t)a * 18;
                const double *dpa = s->dp + s->jidx[a] * 6;
                for (k = 0; k < 6; ++k) sum = sum + W[k * 3 + c] * dpa[k];
            }
            s->eab[N + i * 3 + c] = s->g[N + i * 3 + c] - sum;
        }
    }
}

/* compute_dpb.cl:6-35 ; NDRange {3n} (sba_func.cpp:1090): dpb = Vinv * eb, mixed-triangle read */
static void k_dpb(orc_state *s)
{
    int n = s->n, N = s->N, i, nt = s->nthreads; (void)nt;
    PFOR
    for (i = 0; i < n; ++i) {
        const double *Vi = s->V + (size_t)i * 9, *e = s->eab + N + i * 3;
        int r;
        for (r = 0; r < 3; ++r) {
            double sum;
            if (r < 2) sum = Vi[r * 3] * e[0] + Vi[3 + r] * e[1] + Vi[6 + r] * e[2];
            else sum = Vi[r * 3] * e[0] + Vi[r * 3 + 1] * e[1] + Vi[r * 3 + 2] * e[2];
            s->dp[N + i * 3 + r] = sum;
        }
    }
}

/* compute_newp.cl:6-26 ; NDRange {T} (sba_func.cpp:1142) */
static void k_newp(orc_state *s)
{
    int k;
    for (k = 0; k < s->N; ++k) s->newcams[k] = s->cams[k] + s->dp[k];
    for (k = 0; k < 3 * s->n; ++k) s->newpts[k] = s->pts[k] + s->dp[s->N + k];
}

/* update_p.cl:6-25 ; NDRange {T} (sba_func.cpp:1188) */
static void k_update_p(orc_state *s)
{
    memcpy(s->cams, s->newcams, (size_t)s->N * 8);
    memcpy(s->pts, s->newpts, (size_t)3 * s->n * 8);
}

/* compute_Jmultiply.cl:6-52 ; NDRange {2mn} (sba_func.cpp:45).  The reference writes a dense
 * vector out[(i*m+j)*2+k] that is 0 where (i,j) has no observation; the host then forms
 * left-to-right dot products over it (trust_region.cpp:126,174-176).  Zeros do not change a
 * left-to-right sum, and (i*m+j) ascending is exactly the observation order, so the sparse
 * vector out[idx*2+k] yields bit-identical dot products. */
static void k_Jmultiply(orc_state *s, const double *x, double *out)
{
    int o = s->o, N = s->N, idx, nt = s->nthreads; (void)nt;
    PFOR
    for (idx = 0; idx < o; ++idx) {
        const double *JA = s->JA + (size_t)idx * 12, *JB = s->JB + (size_t)idx * 6;
        const double *xa = x + s->jidx[idx] * 6, *xb = x + N + (size_t)s->iidx[idx] * 3;
        int k, t;
        for (k = 0; k < 2; ++k) {
            double sum = 0;
            for (t = 0; t < 6; ++t) sum += JA[k * 6 + t] * xa[t];
            for (t = 0; t < 3; ++t) sum += JB[k * 3 + t] * xb[t];
            out[idx * 2 + k] = sum;
        }
    }
}

static const orc_ops g_native = {
    k_exQT, k_jacobiQT, k_U, k_V, k_Wblks, k_g, k_update_UV, k_restore_UVdiag, k_Vinv,
    k_Yblks, k_S, k_ea, k_matVec, k_eb, k_dpb, k_newp, k_update_p, k_Jmultiply, "restatement"
};
const orc_ops *orc_native_ops(void) { return &g_native; }
