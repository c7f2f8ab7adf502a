// kernels_pcg.cu -- iterative camera solve: block-Jacobi preconditioned conjugate gradients on the tile pool.
//
// "a sparse/PCG camera solve for m >> 10^3" -- the reference has nothing of the kind (its only
// solver is the dense inverse of PSBA/cl_spdinv.cpp:18-204), so this is an OPTION (psba_set_option "camera_solver" = 1),
// never the default: the direct tiled Cholesky of kernels_chol.cu stays the path whose results are compared with the
// reference.  S dpa = ea is solved on the tiles of S itself (no fill-in is ever touched), preconditioned by the inverses
// of the 6x6 camera blocks on the diagonal (the classical choice for reduced camera systems).  Everything runs on the
// engine's stream; scalars (alpha, beta, residual norms) stay on the device, the host looks at the residual every few
// iterations only.  All sums have a fixed order: same bits run to run.
#include "dev_math.cuh"

#define PCG_NT 256

// Minv_k = inverse of the 6x6 diagonal block of camera position p (identity on padding); not positive definite -> status 1
__global__ void k_pcg_prep(int npos, int nt, const int *__restrict__ tile_index, const int *__restrict__ pos2cam,
                           const double *__restrict__ Stiles, double *__restrict__ Minv, int *__restrict__ status)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npos) return;
    double *o = Minv + (size_t)p * 36;
    if (pos2cam[p] < 0) { for (int q = 0; q < 36; ++q) o[q] = (q % 7 == 0) ? 1.0 : 0.0; return; }
    const int I = p / 8, r0 = (p % 8) * 6;
    const double *t = Stiles + (size_t)tile_index[I * nt + I] * TS * TS;
    double L[6][6], X[6][6];
    bool bad = false;
    for (int r = 0; r < 6; ++r)
        for (int cc = 0; cc <= r; ++cc) {
            double s = t[(r0 + r) * TS + r0 + cc];
            for (int k = 0; k < cc; ++k) s -= L[r][k] * L[cc][k];
            if (r == cc) { bad |= !(s > 0.0 && s < 1e300); L[r][r] = sqrt(s); } else L[r][cc] = s / L[cc][cc];
        }
    if (bad) { *status = 1; return; }
    for (int cc = 0; cc < 6; ++cc) {                          // X = L^-1 column by column
        for (int r = 0; r < 6; ++r) {
            double s = r == cc ? 1.0 : 0.0;
            for (int k = cc; k < r; ++k) s -= L[r][k] * X[k][cc];
            X[r][cc] = r < cc ? 0.0 : s / L[r][r];
        }
    }
    for (int r = 0; r < 6; ++r)
        for (int cc = 0; cc < 6; ++cc) {                      // (L L^T)^-1 = X^T X
            double s = 0.0;
            for (int k = (r > cc ? r : cc); k < 6; ++k) s += X[k][r] * X[k][cc];
            o[r * 6 + cc] = s;
        }
}

// y = S x in the solver's (padded) ordering.  One CTA per tile row I; thread (rr, part): row rr of the tile row, the
// tiles of the row and of the column dealt over `part` (fixed assignment), partial sums combined in a fixed order.
// Only the lower triangle of S is stored (tiles I >= J; inside a diagonal tile only entries with column <= row are
// used), so row I also collects the transposed tiles of column I.
__global__ void __launch_bounds__(PCG_NT) k_pcg_spmv(int nt, int n_tiles_S, const int *__restrict__ tile_index, const double *__restrict__ Stiles,
                                                   const double *__restrict__ x, double *__restrict__ y)
{
    __shared__ double part[PCG_NT / TS + 1][TS];
    const int I = blockIdx.x, rr = threadIdx.x % TS, pt = threadIdx.x / TS, NP = PCG_NT / TS;   // 5 parts of 48 threads (240 used)
    double s = 0.0;
    if (pt < NP) {
        for (int J = pt; J < nt; J += NP) {
            if (J <= I) {
                const int slot = tile_index[I * nt + J];
                if (slot < 0 || slot >= n_tiles_S) continue;
                const double *t = Stiles + (size_t)slot * TS * TS + rr * TS;
                const double *xv = x + J * TS;
                const int lim = J == I ? rr + 1 : TS;
                for (int cc = 0; cc < lim; ++cc) s += t[cc] * xv[cc];
                if (J == I) for (int cc = rr + 1; cc < TS; ++cc) s += Stiles[(size_t)slot * TS * TS + cc * TS + rr] * xv[cc];
            } else {
                const int slot = tile_index[J * nt + I];
                if (slot < 0 || slot >= n_tiles_S) continue;
                const double *t = Stiles + (size_t)slot * TS * TS + rr;      // column rr of tile (J, I)
                const double *xv = x + J * TS;
                for (int cc = 0; cc < TS; ++cc) s += t[cc * TS] * xv[cc];
            }
        }
        part[pt][rr] = s;
    }
    __syncthreads();
    if (threadIdx.x < TS) {
        double a = 0.0;
#pragma unroll
        for (int q = 0; q < NP; ++q) a += part[q][threadIdx.x];
        y[I * TS + threadIdx.x] = a;
    }
}

// z = Minv r (6x6 blocks); one thread per row
__global__ void k_pcg_precond(int n, const double *__restrict__ Minv, const double *__restrict__ r, double *__restrict__ z)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = i / 6, rr = i % 6;
    const double *mi = Minv + (size_t)p * 36 + rr * 6, *rv = r + p * 6;
    z[i] = mi[0] * rv[0] + mi[1] * rv[1] + mi[2] * rv[2] + mi[3] * rv[3] + mi[4] * rv[4] + mi[5] * rv[5];
}

// fixed-order dot products by ONE CTA (n <= a few 10^4): out[0] = <a,b>, out[1] = <c,d> (second pair optional)
__global__ void __launch_bounds__(1024) k_pcg_dots(int n, const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ c2,
                                                  const double *__restrict__ d2, double *__restrict__ out)
{
    __shared__ double sh[2][1024];
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) { s0 += a[i] * b[i]; if (c2) s1 += c2[i] * d2[i]; }
    sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) { sh[0][threadIdx.x] += sh[0][threadIdx.x + w]; sh[1][threadIdx.x] += sh[1][threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sh[0][0]; out[1] = sh[1][0]; }
}

// sc: [0] rz  [1] pAp  [2] rz_new  [3] rr  [4] bb  [5] breakdown flag
// x += alpha p, r -= alpha Ap with alpha = rz / pAp (device scalars); a non-positive pAp marks S as not positive definite
__global__ void k_pcg_update_xr(int n, const double *__restrict__ p, const double *__restrict__ Ap, double *__restrict__ x, double *__restrict__ r,
                                double *__restrict__ sc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double pAp = sc[1];
    if (!(pAp > 0.0)) { if (i == 0) sc[5] = 1.0; return; }
    if (i >= n) return;
    const double alpha = sc[0] / pAp;
    x[i] += alpha * p[i];
    r[i] -= alpha * Ap[i];
}
// p = z + beta p with beta = rz_new / rz; thread 0 of the LAST block rolls rz forward after everyone has read it: done by a
// second tiny kernel to stay race free
__global__ void k_pcg_update_p(int n, const double *__restrict__ z, double *__restrict__ p, const double *__restrict__ sc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double beta = sc[2] / sc[0];
    p[i] = z[i] + beta * p[i];
}
__global__ void k_pcg_roll(double *sc) { sc[0] = sc[2]; }

__global__ void k_pcg_gather_rhs(int npad, const int *__restrict__ pos2cam, const double *__restrict__ ea, double *__restrict__ b)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npad) return;
    const int cam = pos2cam[k / 6];
    b[k] = cam >= 0 ? ea[cam * 6 + k % 6] : 0.0;
}
__global__ void k_pcg_scatter_sol(int npad, const int *__restrict__ pos2cam, const double *__restrict__ x, double *__restrict__ sol)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npad) return;
    const int cam = pos2cam[k / 6];
    if (cam >= 0) sol[cam * 6 + k % 6] = x[k];
}

// dp[0..N) = S^-1 eab[0..N) by PCG on the S assembled by the last psba_launch_schur.  Sets d_status[0] (0 ok, 1 not positive
// definite / no convergence) like the factorisation does; returns the number of iterations.
int psba_launch_pcg(psba_ctx *c)
{
    const int nt = c->nt, npad = nt * TS, npos = nt * 8;
    cudaStream_t st = c->stream;
    if (!c->pcg_work) c->pcg_work = (double *)psba_dev_alloc(c, ((size_t)6 * npad + (size_t)npos * 36 + 16) * sizeof(double), true);
    double *b = c->pcg_work, *x = b + npad, *r = x + npad, *z = r + npad, *p = z + npad, *Ap = p + npad, *Minv = Ap + npad, *sc = Minv + (size_t)npos * 36;
    const int nb = cdiv(npad, 256);
    CUDA_CHECK(cudaMemsetAsync(c->d_status, 0, sizeof(int), st));
    CUDA_CHECK(cudaMemsetAsync(x, 0, (size_t)npad * 8, st));
    CUDA_CHECK(cudaMemsetAsync(sc, 0, 16 * 8, st));
    k_pcg_prep<<<cdiv(npos, 128), 128, 0, st>>>(npos, nt, c->tile_index, c->pos2cam, c->Stiles, Minv, c->d_status);
    k_pcg_gather_rhs<<<nb, 256, 0, st>>>(npad, c->pos2cam, c->eab, b);
    CUDA_CHECK(cudaMemcpyAsync(r, b, (size_t)npad * 8, cudaMemcpyDeviceToDevice, st));          // x0 = 0: r0 = b
    k_pcg_precond<<<nb, 256, 0, st>>>(npad, Minv, r, z);
    CUDA_CHECK(cudaMemcpyAsync(p, z, (size_t)npad * 8, cudaMemcpyDeviceToDevice, st));
    k_pcg_dots<<<1, 1024, 0, st>>>(npad, r, z, b, b, sc + 6);                                    // rz, bb
    CUDA_CHECK(cudaMemcpyAsync(sc, sc + 6, 8, cudaMemcpyDeviceToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(sc + 4, sc + 7, 8, cudaMemcpyDeviceToDevice, st));
    c->st_launches += 4;
    int it = 0, status = 0;
    bool done = false;
    while (!done && it < c->pcg_max_iter) {
        for (int q = 0; q < 8 && it < c->pcg_max_iter; ++q, ++it) {
            k_pcg_spmv<<<nt, PCG_NT, 0, st>>>(nt, c->n_tiles_S, c->tile_index, c->Stiles, p, Ap);
            k_pcg_dots<<<1, 1024, 0, st>>>(npad, p, Ap, nullptr, nullptr, sc + 8);
            CUDA_CHECK(cudaMemcpyAsync(sc + 1, sc + 8, 8, cudaMemcpyDeviceToDevice, st));      // pAp
            k_pcg_update_xr<<<nb, 256, 0, st>>>(npad, p, Ap, x, r, sc);
            k_pcg_precond<<<nb, 256, 0, st>>>(npad, Minv, r, z);
            k_pcg_dots<<<1, 1024, 0, st>>>(npad, r, z, r, r, sc + 2);                            // rz_new, rr
            k_pcg_update_p<<<nb, 256, 0, st>>>(npad, z, p, sc);
            k_pcg_roll<<<1, 1, 0, st>>>(sc);
            c->st_launches += 7;
        }
        LAUNCH_CHECK();
        double h[6];
        CUDA_CHECK(cudaMemcpyAsync(h, sc, sizeof(h), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(&status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (status != 0 || h[5] != 0.0 || !(h[3] == h[3])) { status = 1; break; }              // block not PD, p^T S p <= 0, NaN
        if (h[3] <= c->pcg_tol * c->pcg_tol * h[4]) done = true;                                // ||r|| <= tol ||b||
    }
    if (!done && status == 0) status = 1;                                                       // no convergence: treated like a failed factorisation
    if (status) { const int one = 1; CUDA_CHECK(cudaMemcpyAsync(c->d_status, &one, sizeof(int), cudaMemcpyHostToDevice, st)); CUDA_CHECK(cudaStreamSynchronize(st)); }
    k_pcg_scatter_sol<<<nb, 256, 0, st>>>(npad, c->pos2cam, x, c->dp);
    c->st_launches += 1;
    LAUNCH_CHECK();
    c->pcg_last_iters = it;
    c->factor_valid = false;
    return it;
}
