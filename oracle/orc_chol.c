/*
 * orc_chol.c -- ORACLE (test infrastructure): serial restatement of the camera-system
 * solvers.  These two reference files use OpenCL-2.0 Blocks + enqueue_kernel and cannot be
 * compiled by g++, so this restatement is their only executable form here:
 *
 *   CL_files/SPD_inv.cl:20-411   + PSBA/cl_spdinv.cpp:18-204  (3x3-block Cholesky, block
 *                                  triangular inverse, S^-1 = L^-T L^-1)
 *   CL_files/cholmod_blk.cl:87-847 + PSBA/cl_cholmod.cpp:25-202 (modified Cholesky)
 *
 * The device-side enqueue chains are sequential loops over block columns; every work-group
 * of one launch is independent, so a serial sweep reproduces the arithmetic exactly.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "psba_oracle.h"

static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* closed-form factor of a 3x3 block from its lower triangle T (row-major 3x3):
 * SPD_inv.cl:62-118.  L = {L00, L10, L11, L20, L21, L22}. */
static void chol3_closed(const double *T, double *L)
{
    double t1, t2, t3, t4, t5;
    L[0] = sqrt(T[0]);
    L[1] = T[3] / sqrt(T[0]);
    L[2] = sqrt(T[4] - T[3] * T[3] / T[0]);
    L[3] = T[6] / sqrt(T[0]);
    L[4] = sqrt(T[0] / (T[0] * T[4] - T[3] * T[3])) * (T[7] - T[3] * T[6] / T[0]);
    t1 = -T[8] * T[3] * T[3];
    t2 = 2 * T[7] * T[3] * T[6];
    t3 = -T[4] * T[6] * T[6];
    t4 = T[0] * (T[4] * T[8] - T[7] * T[7]);
    t5 = -T[3] * T[3] + T[0] * T[4];
    L[5] = sqrt((t1 + t2 + t3 + t4) / t5);
}

/* inverse of the 3x3 lower-triangular factor, SPD_inv.cl:121-163 / cholmod_blk.cl:220-262 */
static void tri3_inverse(const double *L, double *inv /* row-major 3x3 */)
{
    inv[0] = 1 / L[0]; inv[1] = 0; inv[2] = 0;
    inv[3] = -L[1] / (L[0] * L[2]); inv[4] = 1 / L[2]; inv[5] = 0;
    inv[6] = (L[1] * L[4] - L[2] * L[3]) / (L[0] * L[2] * L[5]);
    inv[7] = -L[4] / (L[2] * L[5]); inv[8] = 1 / L[5];
}

/* kern_cholesky + kern_cholesky_s2, SPD_inv.cl:20-239; host cl_spdinv.cpp:57-103.
 * Returns 0.0, or 1.0 as soon as a factor entry of a diagonal block is not finite. */
double orc_cholesky(double *mat, double *diagInv, int N)
{
    int nb = N / 3, i, j, k, u, v;
    for (j = 0; j < nb; ++j) {
        double T[9], L[6], inv[9];
        double ret = 0.0;
        /* step 1: T_ij = A_ij - sum_k<j L_ik L_jk^T for every block row i >= j */
        for (i = j; i < nb; ++i)
            for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v) {
                size_t a = (size_t)(i * 3 + u) * N + j * 3 + v;
                double sum = mat[a];
                for (k = 0; k < j; ++k)
                    sum -= dot3(&mat[(size_t)(i * 3 + u) * N + k * 3], &mat[(size_t)(j * 3 + v) * N + k * 3]);
                mat[a] = sum;
                if (i == j) T[u * 3 + v] = sum;
            }
        chol3_closed(T, L);
        for (k = 0; k < 6; ++k) if (!isfinite(L[k])) ret = 1.0;
        {
            double *d = &mat[(size_t)(j * 3) * N + j * 3];
            d[0] = L[0]; d[1] = 0; d[2] = 0;
            d[N] = L[1]; d[N + 1] = L[2]; d[N + 2] = 0;
            d[2 * N] = L[3]; d[2 * N + 1] = L[4]; d[2 * N + 2] = L[5];
        }
        tri3_inverse(L, inv);
        for (k = 0; k < 9; ++k) { diagInv[j * 9 + k] = inv[k]; if (!isfinite(inv[k])) ret = 1.0; }
        if (ret != 0.0) return ret;              /* chain stops, SPD_inv.cl:165 */
        /* step 2: L_ij = T_ij L_jj^-T, zero the mirrored block */
        for (i = j + 1; i < nb; ++i) {
            double row[3][3];
            for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v)
                row[u][v] = dot3(&mat[(size_t)(i * 3 + u) * N + j * 3], &inv[v * 3]);
            for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v) {
                mat[(size_t)(i * 3 + u) * N + j * 3 + v] = row[u][v];
                mat[(size_t)(j * 3 + v) * N + i * 3 + u] = 0;
            }
        }
    }
    return 0.0;
}

/* kern_trigMat_inv, SPD_inv.cl:248-328: block diagonal ii of X = L^-1, stored transposed in
 * the upper triangle; diagonal blocks are overwritten by L_ii^-T. */
void orc_trigMat_inv(double *in, const double *diagBlk, int N)
{
    int nb = N / 3, ii, i, j, k, u, v, t;
    for (ii = 0; ii < nb; ++ii)
        for (j = 0; j + ii < nb; ++j) {
            i = ii + j;
            if (i == j) {
                for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v)
                    in[(size_t)(i * 3 + u) * N + j * 3 + v] = diagBlk[i * 9 + v * 3 + u];
            } else {
                double Tt[3][3], Xn[3][3];
                for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v) {
                    double tt = 0.0;
                    for (k = j; k < i; ++k) {
                        const double *a = &in[(size_t)(i * 3 + u) * N + k * 3];
                        const double *b = &in[(size_t)(j * 3 + v) * N + k * 3];
                        tt += a[0] * b[0]; tt += a[1] * b[1]; tt += a[2] * b[2];
                    }
                    Tt[u][v] = tt;
                }
                for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v)
                    in[(size_t)(j * 3 + v) * N + i * 3 + u] = Tt[u][v];
                for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v) {
                    double tt = 0.0;
                    for (t = 0; t < 3; ++t)
                        tt += in[(size_t)(i * 3 + t) * N + i * 3 + u] * in[(size_t)(j * 3 + v) * N + i * 3 + t];
                    Xn[u][v] = tt;
                }
                for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v)
                    in[(size_t)(j * 3 + v) * N + i * 3 + u] = -Xn[u][v];
            }
        }
}

/* kern_trigMat_mul + kern_fill_rest, SPD_inv.cl:343-411 */
void orc_trigMat_mul(double *in, double *diag, int N)
{
    int u, v, k;
    for (u = 0; u < N; ++u)
        for (v = 0; v <= u; ++v) {
            double val = 0;
            for (k = u; k < N; ++k) val += in[(size_t)u * N + k] * in[(size_t)v * N + k];
            if (u == v) diag[u] = val; else in[(size_t)u * N + v] = val;
        }
    for (u = 0; u < N; ++u) {
        in[(size_t)u * N + u] = diag[u];
        for (v = 0; v < u; ++v) in[(size_t)v * N + u] = in[(size_t)u * N + v];
    }
}

/* SPDinv, cl_spdinv.cpp:18-40.  The success path of the reference falls off the end of the
 * function without a return; the intended value 0.0 is returned here. */
double orc_SPDinv(double *mat, double *diagAux, int N)
{
    double ret = orc_cholesky(mat, diagAux, N);
    if (ret != 0.0) return ret;
    orc_trigMat_inv(mat, diagAux, N);
    orc_trigMat_mul(mat, diagAux, N);
    return 0.0;
}

/* variant P of SURVEY App. B.2: plain scalar Cholesky + two triangular solves */
double orc_potrf_solve(double *A, const double *rhs, double *x, int N)
{
    int i, j, k;
    for (j = 0; j < N; ++j) {
        double d = A[(size_t)j * N + j];
        for (k = 0; k < j; ++k) d -= A[(size_t)j * N + k] * A[(size_t)j * N + k];
        if (!(d > 0.0) || !isfinite(d)) return 1.0;
        d = sqrt(d);
        A[(size_t)j * N + j] = d;
        for (i = j + 1; i < N; ++i) {
            double sum = A[(size_t)i * N + j];
            for (k = 0; k < j; ++k) sum -= A[(size_t)i * N + k] * A[(size_t)j * N + k];
            A[(size_t)i * N + j] = sum / d;
        }
    }
    for (i = 0; i < N; ++i) {
        double sum = rhs[i];
        for (k = 0; k < i; ++k) sum -= A[(size_t)i * N + k] * x[k];
        x[i] = sum / A[(size_t)i * N + i];
    }
    for (i = N - 1; i >= 0; --i) {
        double sum = x[i];
        for (k = i + 1; k < N; ++k) sum -= A[(size_t)k * N + i] * x[k];
        x[i] = sum / A[(size_t)i * N + i];
    }
    return 0.0;
}

/* get_delta_beta, cl_cholmod.cpp:109-167 + kern_mat_max, cholmod_blk.cl:796-825 */
void orc_get_delta_beta(const double *mat, int N, double *delta, double *beta)
{
    double xi = 0, gamma = 0;
    int r, k;
    for (r = 0; r < N; ++r) {
        double mx = 0;
        for (k = 0; k < N; ++k) {
            double t1;
            if (k == r) continue;
            t1 = fabs(mat[(size_t)r * N + k]);
            if (t1 > mx) mx = t1;
        }
        if (mx > xi) xi = mx;
        if (fabs(mat[(size_t)r * N + r]) > gamma) gamma = fabs(mat[(size_t)r * N + r]);
    }
    *delta = 1e-15 * fmax(xi + gamma, 1);
    *beta = fmax(gamma, 1e-15);
    *beta = fmax(*beta, xi / sqrt((double)N * N - 1));
    *beta = sqrt(*beta);
}

/* scalar Gill-Murray column x = 3j+col, cholmod_blk.cl:446-697 (steps 1-4) */
static void cholmod_scalar_col(double *mat, double *aux, int N, double beta, double delta, int x)
{
    size_t jj = (size_t)x * N + x;
    double sum = mat[jj], d_j;
    int k, i, flagged = 0;
    for (k = 0; k < x; ++k) { double L = mat[(size_t)x * N + k]; sum -= L * L; }
    sum = fabs(sum);
    d_j = fmax(sum, delta);
    aux[x] = d_j;
    mat[jj] = sqrt(d_j);
    for (i = x + 1; i < N; ++i) {                      /* step 2 */
        double C = mat[(size_t)i * N + x];
        for (k = 0; k < x; ++k) C = C - (mat[(size_t)i * N + k] * mat[(size_t)x * N + k]);
        aux[N + i] = C;
        mat[(size_t)i * N + x] = C / mat[jj];
        mat[(size_t)x * N + i] = 0;
        if (mat[(size_t)i * N + x] > beta) flagged = 1;  /* signed compare, SURVEY A.4 */
    }
    if (flagged) {                                     /* steps 3-4 */
        double theta = 0.0;
        for (k = N + x + 1; k < 2 * N; ++k) theta = fmax(theta, fabs(aux[k]));
        mat[jj] = theta / beta;
        aux[x] = mat[jj] * mat[jj];
        for (i = x + 1; i < N; ++i) mat[(size_t)i * N + x] = aux[N + i] / mat[jj];
    }
}

/* kern_cholmod_blk / _blk_step2 / _blk_step3 / _step1..4 / _diaginv, cholmod_blk.cl:87-784.
 * aux: >= 3N doubles (block back-ups at [i*9..], scalar scratch d at [0,N), C at [N,2N)).
 * diag receives the original diagonal of every block column that starts on the block path. */
void orc_cholmod_blk(double *mat, double *aux, double *diagInv, double *diag, int N,
                     double beta, double delta, int *n_scalar_blocks)
{
    int nb = N / 3, i, j, k, u, v, nscalar = 0;
    for (j = 0; j < nb; ++j) {
        double T[9], L[6], inv[9];
        int fail = 0, scalar = 0;
        double *d = &mat[(size_t)(j * 3) * N + j * 3];
        /* back up A_jj and its diagonal, cholmod_blk.cl:107-114 */
        for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v) {
            aux[j * 9 + u * 3 + v] = d[(size_t)u * N + v];
            if (u == v) diag[j * 3 + u] = d[(size_t)u * N + v];
        }
        for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v) {
            double sum = aux[j * 9 + u * 3 + v];
            for (k = 0; k < j; ++k)
                sum -= dot3(&mat[(size_t)(j * 3 + u) * N + k * 3], &mat[(size_t)(j * 3 + v) * N + k * 3]);
            T[u * 3 + v] = sum;
        }
        /* closed-form factor with pivot tests, cholmod_blk.cl:133-194 */
        {
            double t1, t2, t3, t4, t5;
            L[0] = T[0];
            if (!isfinite(L[0]) || L[0] <= 0) fail = 1; else L[0] = sqrt(L[0]);
            L[1] = T[3] / sqrt(T[0]);
            if (!isfinite(L[1])) fail = 1;
            L[2] = T[4] - T[3] * T[3] / T[0];
            if (!isfinite(L[2]) || L[2] <= 0) fail = 1; else L[2] = sqrt(L[2]);
            L[3] = T[6] / sqrt(T[0]);
            if (!isfinite(L[3])) fail = 1;
            L[4] = sqrt(T[0] / (T[0] * T[4] - T[3] * T[3])) * (T[7] - T[3] * T[6] / T[0]);
            if (!isfinite(L[4])) fail = 1;
            t1 = -T[8] * T[3] * T[3];
            t2 = 2 * T[7] * T[3] * T[6];
            t3 = -T[4] * T[6] * T[6];
            t4 = T[0] * (T[4] * T[8] - T[7] * T[7]);
            t5 = -T[3] * T[3] + T[0] * T[4];
            L[5] = (t1 + t2 + t3 + t4) / t5;
            if (!isfinite(L[5]) || L[5] <= 0) fail = 1; else L[5] = sqrt(L[5]);
        }
        if (fail) {
            /* restore A_jj (all 9 entries) and take the scalar path, cholmod_blk.cl:198-214 */
            for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v) d[(size_t)u * N + v] = aux[j * 9 + u * 3 + v];
            scalar = 1;
        } else {
            d[0] = L[0]; d[1] = 0; d[2] = 0;
            d[N] = L[1]; d[N + 1] = L[2]; d[N + 2] = 0;
            d[2 * N] = L[3]; d[2 * N + 1] = L[4]; d[2 * N + 2] = L[5];
            tri3_inverse(L, inv);
            for (k = 0; k < 9; ++k) diagInv[j * 9 + k] = inv[k];
            if (N - (j + 1) * 3 >= 3) {
                int over = 0;
                /* step 2, cholmod_blk.cl:290-360 */
                for (i = j + 1; i < nb; ++i) {
                    double Tij[3][3];
                    for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v) {
                        size_t a = (size_t)(i * 3 + u) * N + j * 3 + v;
                        double sum = mat[a];
                        aux[i * 9 + u * 3 + v] = sum;
                        for (k = 0; k < j; ++k)
                            sum -= dot3(&mat[(size_t)(i * 3 + u) * N + k * 3], &mat[(size_t)(j * 3 + v) * N + k * 3]);
                        Tij[u][v] = sum;
                    }
                    for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v) {
                        double sum = dot3(Tij[u], &inv[v * 3]);
                        mat[(size_t)(i * 3 + u) * N + j * 3 + v] = sum;
                        mat[(size_t)(j * 3 + v) * N + i * 3 + u] = 0;
                        if (sum > beta) over = 1;          /* signed compare */
                    }
                }
                if (over) {
                    /* step 3 failure branch, cholmod_blk.cl:386-414: restore the block column
                     * and the lower triangle of A_jj, then the scalar path */
                    for (i = j + 1; i < nb; ++i)
                        for (u = 0; u < 3; ++u) for (v = 0; v < 3; ++v)
                            mat[(size_t)(i * 3 + u) * N + j * 3 + v] = aux[i * 9 + u * 3 + v];
                    d[0] = aux[j * 9];
                    d[N] = aux[j * 9 + 3]; d[N + 1] = aux[j * 9 + 4];
                    d[2 * N] = aux[j * 9 + 6]; d[2 * N + 1] = aux[j * 9 + 7]; d[2 * N + 2] = aux[j * 9 + 8];
                    scalar = 1;
                }
            }
        }
        if (scalar) {
            int col;
            nscalar++;
            for (col = 0; col < 3; ++col) cholmod_scalar_col(mat, aux, N, beta, delta, j * 3 + col);
            /* kern_cholmod_diaginv, cholmod_blk.cl:703-763 */
            L[0] = d[0]; L[1] = d[N]; L[2] = d[N + 1]; L[3] = d[2 * N]; L[4] = d[2 * N + 1]; L[5] = d[2 * N + 2];
            tri3_inverse(L, inv);
            for (k = 0; k < 9; ++k) diagInv[j * 9 + k] = inv[k];
        }
    }
    if (n_scalar_blocks) *n_scalar_blocks = nscalar;
}

/* kern_cholmod_E, cholmod_blk.cl:830-847: E_i = sum_{k<=i} L_ik^2 - diag_i */
void orc_cholmod_E(const double *mat, double *diag, int N)
{
    int i, k;
    for (i = 0; i < N; ++i) {
        double sum = 0.0;
        for (k = 0; k <= i; ++k) sum += mat[(size_t)i * N + k] * mat[(size_t)i * N + k];
        diag[i] = sum - diag[i];
    }
}
