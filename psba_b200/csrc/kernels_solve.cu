// kernels_solve.cu -- camera solve: tiled FP64 Cholesky of the reduced camera system S, the two
// triangular solves, explicit inverse (ABI parity only) and the modified Cholesky of the
// trust-region fallback.
//
// Replaces kern_cholesky / kern_cholesky_s2 / kern_trigMat_inv / kern_trigMat_mul / kern_fill_rest
// (CL_files/SPD_inv.cl:20-411, host PSBA/cl_spdinv.cpp:18-204), kern_matVec_mul (matVec_mul.cl:7-17)
// and kern_cholmod_* / kern_mat_max / kern_cholmod_E (cholmod_blk.cl:87-847, PSBA/cl_cholmod.cpp).
//
// S is stored as a pool of 48x48 tiles (8 camera blocks per tile edge), only the tiles of the
// symbolic factor (lower triangle incl. fill-in) exist.  Right-looking tiled factorisation:
// per panel K  potrf(K,K) -> trsm of the tiles below -> rank-48 updates of the trailing tiles.
// The task lists are fixed by the camera-pair structure, so the whole factorisation is one CUDA
// graph.  Failure (pivot <= 0 or not finite) sets a status word; the reference reports the same
// event as "not finite factor entry" (SPD_inv.cl:66-107).  FP64 on CUDA cores: tcgen05 has no
// FP64 kind, and at N = 6m <= ~10^3 the panels are far too small for DMMA to matter.
#include "dev_math.cuh"
#include <algorithm>

#define LDT (TS + 1)   // padded leading dimension in shared memory

// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// tile pool -> dense N x N row-major.  cam2pos != null: rows/columns in the CALLERS' camera order (S of the
// ABI and of the modified Cholesky; mirror fills the upper triangle); cam2pos == null: the solver's own
// ordering (the factor L, lower triangle, diagonal tiles from Ldiag).
__global__ void k_tiles_to_dense(int N, int nt, const int *__restrict__ tile_index, const int *__restrict__ cam2pos,
                                 const double *__restrict__ Stiles, double *__restrict__ dense, int mirror, const double *__restrict__ Ldiag)
{
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)N * N) return;
    const int r = (int)(e / N), cc = (int)(e % N);
    if (cc > r && !mirror) { dense[e] = 0.0; return; }
    int rr = r, c2 = cc;
    if (cam2pos) { rr = cam2pos[r / 6] * 6 + r % 6; c2 = cam2pos[cc / 6] * 6 + cc % 6; }
    if (c2 > rr) { const int t = rr; rr = c2; c2 = t; }
    if (Ldiag && rr / TS == c2 / TS) { dense[e] = Ldiag[(size_t)(rr / TS) * TS * TS + (rr % TS) * TS + c2 % TS]; return; }
    const int slot = tile_index[(rr / TS) * nt + c2 / TS];
    dense[e] = slot < 0 ? 0.0 : Stiles[(size_t)slot * TS * TS + (rr % TS) * TS + c2 % TS];
}

void psba_tiles_to_dense(psba_ctx *c, double *dense_dev, bool mirror)
{
    long long tot = (long long)c->N * c->N;
    k_tiles_to_dense<<<cdiv(tot, 256), 256, 0, c->stream>>>(c->N, c->nt, c->tile_index, c->cam2pos, c->Stiles, dense_dev, mirror ? 1 : 0, nullptr);
    c->st_launches += 1;
    LAUNCH_CHECK();
}

// the factor L as a dense lower-triangular matrix of size npad = nt*TS in the solver's ordering
// (diagonal tiles live in their own pool)
static void psba_factor_to_dense(psba_ctx *c, double *dense_dev)
{
    const int npad = c->nt * TS;
    long long tot = (long long)npad * npad;
    k_tiles_to_dense<<<cdiv(tot, 256), 256, 0, c->stream>>>(npad, c->nt, c->tile_index, nullptr, c->Stiles, dense_dev, 0, c->Ldiag);
    c->st_launches += 1;
    LAUNCH_CHECK();
}

// ABI parity only (SPDinv's explicit inverse, cl_spdinv.cpp:18-40): column of S^-1 for (camera, component)
// cidx by one forward and one backward substitution on the dense factor (solver's ordering, padded size
// NP), written back in the callers' camera order.  O(NP^2) per thread; small N only.
// half = 1: only the forward substitution, i.e. the column of L^-1 (trigMat_inv, cl_spdinv.cpp:120-162)
__global__ void k_explicit_inverse(int N, int NP, const int *__restrict__ cam2pos, const double *__restrict__ L,
                                   double *__restrict__ work, double *__restrict__ out, int half)
{
    int cidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx >= N) return;
    const int pc = cam2pos[cidx / 6] * 6 + cidx % 6;
    double *x = work + (size_t)cidx * NP;
    for (int r = 0; r < NP; ++r) {
        double s = (r == pc) ? 1.0 : 0.0;
        for (int k = 0; k < r; ++k) s -= L[(size_t)r * NP + k] * x[k];
        x[r] = s / L[(size_t)r * NP + r];
    }
    if (half) {
        for (int r = 0; r < N; ++r) out[(size_t)r * N + cidx] = x[cam2pos[r / 6] * 6 + r % 6];
        return;
    }
    for (int r = NP - 1; r >= 0; --r) {
        double s = x[r];
        for (int k = r + 1; k < NP; ++k) s -= L[(size_t)k * NP + r] * x[k];
        x[r] = s / L[(size_t)r * NP + r];
    }
    for (int r = 0; r < N; ++r) out[(size_t)r * N + cidx] = x[cam2pos[r / 6] * 6 + r % 6];
}

// the factor itself in the callers' camera order: M = P^T L P (M M^T = S; lower triangular when the solver kept the natural order)
__global__ void k_factor_callers_order(int N, int NP, const int *__restrict__ cam2pos, const double *__restrict__ L, double *__restrict__ out)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)N * N) return;
    const int r = (int)(e / N), cc = (int)(e % N);
    out[e] = L[(size_t)(cam2pos[r / 6] * 6 + r % 6) * NP + cam2pos[cc / 6] * 6 + cc % 6];
}

// what = 0: S^-1 (trigMat_mul after trigMat_inv), 1: L^-1 (trigMat_inv), 2: L (cholesky); all in the callers' camera order
void psba_launch_factor_products(psba_ctx *c, double *out_dev, int what)
{
    const size_t NP = (size_t)c->nt * TS;
    double *Ld = (double *)psba_dev_alloc(c, NP * NP * sizeof(double), false);
    psba_factor_to_dense(c, Ld);
    if (what == 2) k_factor_callers_order<<<cdiv((long long)c->N * c->N, 256), 256, 0, c->stream>>>(c->N, (int)NP, c->cam2pos, Ld, out_dev);
    else {
        double *work = (double *)psba_dev_alloc(c, NP * NP * sizeof(double), false);
        k_explicit_inverse<<<cdiv(c->N, 64), 64, 0, c->stream>>>(c->N, (int)NP, c->cam2pos, Ld, work, out_dev, what == 1 ? 1 : 0);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        psba_dev_free(c, work);
    }
    c->st_launches += 2;
    LAUNCH_CHECK();
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    psba_dev_free(c, Ld);
}

void psba_launch_explicit_inverse(psba_ctx *c, double *out_dev)
{
    psba_launch_factor_products(c, out_dev, 0);
}

// ---------------------------------------------------------------------------------------------
// modified Cholesky of the trust-region fallback: exact control flow of cholmod_blk.cl (SURVEY
// A.4) on a dense copy of S; one CTA, every matrix entry is produced by one thread that runs the
// reference's sequential dot product, so the factor matches the reference's arithmetic order.
__device__ __forceinline__ double dot3g(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ void tri3_inverse_dev(const double *L, double *inv)
{
    inv[0] = 1 / L[0]; inv[1] = 0; inv[2] = 0;
    inv[3] = -L[1] / (L[0] * L[2]); inv[4] = 1 / L[2]; inv[5] = 0;
    inv[6] = (L[1] * L[4] - L[2] * L[3]) / (L[0] * L[2] * L[5]);
    inv[7] = -L[4] / (L[2] * L[5]); inv[8] = 1 / L[5];
}

// row-wise max |offdiag| and |diag| (kern_mat_max, cholmod_blk.cl:796-825; cl_cholmod.cpp:109-167)
__global__ void k_mat_max(int N, const double *__restrict__ mat, double *__restrict__ out2)
{
    __shared__ double sx[256], sg[256];
    double xi = 0.0, ga = 0.0;
    for (int r = threadIdx.x; r < N; r += 256) {
        for (int k = 0; k < N; ++k) {
            double t = fabs(mat[(size_t)k * N + r]);           // symmetric input: column r = row r, coalesced
            if (k == r) { if (t > ga) ga = t; } else if (t > xi) xi = t;
        }
    }
    sx[threadIdx.x] = xi; sg[threadIdx.x] = ga;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { sx[threadIdx.x] = fmax(sx[threadIdx.x], sx[threadIdx.x + w]); sg[threadIdx.x] = fmax(sg[threadIdx.x], sg[threadIdx.x + w]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = sx[0]; out2[1] = sg[0]; }
}

// The working matrix is addressed TRANSPOSED, element (r, c) at mat[c * N + r]: the input is symmetric (dense S with
// the mirrored upper triangle), the threads of a warp own consecutive rows r, so every access of the factor is
// coalesced (row-major addressing made each lane pull its own 32-byte sector: 2.5 ms on N = 312).  Only E and the
// number of scalar blocks leave this file.
#define M(r, c) mat[(size_t)(c) * N + (r)]
#define CHM_NT 512
__global__ void __launch_bounds__(CHM_NT) k_cholmod(int N, double *__restrict__ mat, double *__restrict__ aux, double *__restrict__ diagInv,
                                                  double *__restrict__ diag, double beta, double delta, int *__restrict__ nscalar_out)
{
    extern __shared__ double rows3[];                    // [3][N]: the finished factor rows j3 .. j3+2 (columns < j3) of the current block
    __shared__ double T[9], L[6], inv[9];
    __shared__ int fail, over, flagged;
    __shared__ double theta_s;
    const int tid = threadIdx.x, nb = N / 3;
    int nscalar = 0;
    for (int j = 0; j < nb; ++j) {
        const int j3 = j * 3;
        // the three pivot rows are read by nine threads running sequential dot products (T_jj) and by every row of the
        // block column: staged once by the whole CTA (the nine threads alone paid one L2 round trip per term: 24 us per block)
        for (int e = tid; e < 3 * j3; e += CHM_NT) { const int v = e / j3, cc = e - v * j3; rows3[v * N + cc] = M(j3 + v, cc); }
        __syncthreads();
        if (tid < 9) {   // back up A_jj, diag; T_jj (cholmod_blk.cl:107-129)
            const int u = tid / 3, v = tid % 3;
            double sum = M(j3 + u, j3 + v);
            aux[j * 9 + tid] = sum;
            if (u == v) diag[j * 3 + u] = sum;
            const double *ru = rows3 + u * N, *rv = rows3 + v * N;
            for (int k = 0; k < j; ++k) sum -= ru[k * 3] * rv[k * 3] + ru[k * 3 + 1] * rv[k * 3 + 1] + ru[k * 3 + 2] * rv[k * 3 + 2];
            T[tid] = sum;
        }
        if (tid == 0) { fail = 0; over = 0; }
        __syncthreads();
        if (tid == 0) {  // closed-form factor with pivot tests (cholmod_blk.cl:133-194)
            int f = 0;
            L[0] = T[0];
            if (!isfinite(L[0]) || L[0] <= 0) f = 1; else L[0] = sqrt(L[0]);
            L[1] = T[3] / sqrt(T[0]);
            if (!isfinite(L[1])) f = 1;
            L[2] = T[4] - T[3] * T[3] / T[0];
            if (!isfinite(L[2]) || L[2] <= 0) f = 1; else L[2] = sqrt(L[2]);
            L[3] = T[6] / sqrt(T[0]);
            if (!isfinite(L[3])) f = 1;
            L[4] = sqrt(T[0] / (T[0] * T[4] - T[3] * T[3])) * (T[7] - T[3] * T[6] / T[0]);
            if (!isfinite(L[4])) f = 1;
            const double t1 = -T[8] * T[3] * T[3];
            const double t2 = 2 * T[7] * T[3] * T[6];
            const double t3 = -T[4] * T[6] * T[6];
            const double t4 = T[0] * (T[4] * T[8] - T[7] * T[7]);
            const double t5 = -T[3] * T[3] + T[0] * T[4];
            L[5] = (t1 + t2 + t3 + t4) / t5;
            if (!isfinite(L[5]) || L[5] <= 0) f = 1; else L[5] = sqrt(L[5]);
            fail = f;
            if (!f) {
                M(j3, j3) = L[0]; M(j3, j3 + 1) = 0; M(j3, j3 + 2) = 0;
                M(j3 + 1, j3) = L[1]; M(j3 + 1, j3 + 1) = L[2]; M(j3 + 1, j3 + 2) = 0;
                M(j3 + 2, j3) = L[3]; M(j3 + 2, j3 + 1) = L[4]; M(j3 + 2, j3 + 2) = L[5];
                tri3_inverse_dev(L, inv);
                for (int k = 0; k < 9; ++k) diagInv[j * 9 + k] = inv[k];
            }
        }
        __syncthreads();
        bool scalar = false;
        if (fail) scalar = true;     // A_jj was never overwritten on this path (restore is a no-op)
        else if (N - (j + 1) * 3 >= 3) {
            // step 2 (cholmod_blk.cl:290-360): one thread per row (i,u) of the block column
            const int nrows = N - (j + 1) * 3;
            for (int rr = tid; rr < nrows; rr += CHM_NT) {
                const int row = (j + 1) * 3 + rr;
                const int i = row / 3, u = row % 3;
                double Tij[3];
                for (int v = 0; v < 3; ++v) { Tij[v] = M(row, j3 + v); aux[i * 9 + u * 3 + v] = Tij[v]; }
#pragma unroll 8
                for (int k = 0; k < j; ++k) {               // per entry the reference's order: ascending k (cholmod_blk.cl:318-331); eight steps of loads in flight
                    const double a0 = M(row, k * 3), a1 = M(row, k * 3 + 1), a2 = M(row, k * 3 + 2);
#pragma unroll
                    for (int v = 0; v < 3; ++v) Tij[v] -= a0 * rows3[v * N + k * 3] + a1 * rows3[v * N + k * 3 + 1] + a2 * rows3[v * N + k * 3 + 2];
                }
                for (int v = 0; v < 3; ++v) {
                    const double sum = Tij[0] * inv[v * 3] + Tij[1] * inv[v * 3 + 1] + Tij[2] * inv[v * 3 + 2];
                    M(row, j3 + v) = sum;
                    M(j3 + v, row) = 0;
                    if (sum > beta) over = 1;          // signed compare, SURVEY A.4
                }
            }
            __syncthreads();
            if (over) {   // step 3 failure branch (cholmod_blk.cl:386-414)
                for (int rr = tid; rr < nrows; rr += CHM_NT) {
                    const int row = (j + 1) * 3 + rr;
                    const int i = row / 3, u = row % 3;
                    for (int v = 0; v < 3; ++v) M(row, j3 + v) = aux[i * 9 + u * 3 + v];
                }
                if (tid == 0) {
                    M(j3, j3) = aux[j * 9];
                    M(j3 + 1, j3) = aux[j * 9 + 3]; M(j3 + 1, j3 + 1) = aux[j * 9 + 4];
                    M(j3 + 2, j3) = aux[j * 9 + 6]; M(j3 + 2, j3 + 1) = aux[j * 9 + 7]; M(j3 + 2, j3 + 2) = aux[j * 9 + 8];
                }
                scalar = true;
            }
            __syncthreads();
        }
        if (scalar) {   // scalar Gill-Murray path for the three columns (cholmod_blk.cl:446-697)
            ++nscalar;
            for (int col = 0; col < 3; ++col) {
                const int x = j * 3 + col;
                const size_t jj = (size_t)x * N + x;              // the diagonal is where it was
                __syncthreads();
                if (tid == 0) {
                    double sum = mat[jj];
                    for (int k = 0; k < x; ++k) { const double Lk = M(x, k); sum -= Lk * Lk; }
                    sum = fabs(sum);
                    const double dj = fmax(sum, delta);
                    aux[x] = dj;
                    mat[jj] = sqrt(dj);
                    flagged = 0;
                }
                __syncthreads();
                const double ljj = mat[jj];
                for (int i = x + 1 + tid; i < N; i += CHM_NT) {
                    double C = M(i, x);
#pragma unroll 8
                    for (int k = 0; k < x; ++k) C = C - (M(i, k) * M(x, k));
                    aux[N + i] = C;
                    const double lij = C / ljj;
                    M(i, x) = lij;
                    M(x, i) = 0;
                    if (lij > beta) flagged = 1;
                }
                __syncthreads();
                if (flagged) {
                    if (tid == 0) {
                        double theta = 0.0;
                        for (int k = N + x + 1; k < 2 * N; ++k) theta = fmax(theta, fabs(aux[k]));
                        theta_s = theta / beta;
                        mat[jj] = theta_s;
                        aux[x] = theta_s * theta_s;
                    }
                    __syncthreads();
                    for (int i = x + 1 + tid; i < N; i += CHM_NT) M(i, x) = aux[N + i] / theta_s;
                }
                __syncthreads();
            }
            if (tid == 0) {   // kern_cholmod_diaginv (cholmod_blk.cl:703-763)
                L[0] = M(j3, j3); L[1] = M(j3 + 1, j3); L[2] = M(j3 + 1, j3 + 1);
                L[3] = M(j3 + 2, j3); L[4] = M(j3 + 2, j3 + 1); L[5] = M(j3 + 2, j3 + 2);
                tri3_inverse_dev(L, inv);
                for (int k = 0; k < 9; ++k) diagInv[j * 9 + k] = inv[k];
            }
        }
        __syncthreads();
    }
    if (tid == 0) *nscalar_out = nscalar;
}

// E_i = sum_{k<=i} L_ik^2 - diag_i (kern_cholmod_E, cholmod_blk.cl:830-847)
__global__ void k_cholmod_E(int N, const double *__restrict__ mat, double *__restrict__ diag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double sum = 0.0;
    for (int k = 0; k <= i; ++k) sum += M(i, k) * M(i, k);
    diag[i] = sum - diag[i];
}

#undef M
// runs on a dense symmetric matrix (S incl. mirrored upper triangle) of size N on the device; aux / diagInv / diag / E
// are work arrays of 3N+2TS, 3N+2TS and N doubles.  Returns sum_i E_i (left-to-right, trust_region.cpp:358-362).
double psba_launch_cholmod_dense(psba_ctx *c, int N, double *mat, double *aux, double *diagInv, double *E,
                                 double *delta_out, double *beta_out, int *nscalar_out)
{
    k_mat_max<<<1, 256, 0, c->stream>>>(N, mat, c->d_scal + 8);
    CUDA_CHECK(cudaMemcpyAsync(c->h_scal + 8, c->d_scal + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    const double xi = c->h_scal[8], gamma = c->h_scal[9];
    double delta = 1e-15 * fmax(xi + gamma, 1.0);                 // cl_cholmod.cpp:161-164
    double beta = fmax(gamma, 1e-15);
    beta = fmax(beta, xi / sqrt((double)N * N - 1));
    beta = sqrt(beta);
    const int dyn = 3 * N * (int)sizeof(double);
    if (dyn > 200 * 1024) { fprintf(stderr, "psba_b200: dense modified Cholesky: N = %d is beyond the single-CTA kernel (the tile-pool version handles it)\n", N); exit(EXIT_FAILURE); }
    psba_set_smem((const void *)k_cholmod, dyn);
    PROF(c, KID_CHOLMOD) k_cholmod<<<1, CHM_NT, dyn, c->stream>>>(N, mat, aux, diagInv, E, beta, delta, c->d_status + 2);
    k_cholmod_E<<<cdiv(N, 128), 128, 0, c->stream>>>(N, mat, E);
    c->st_launches += 3;
    LAUNCH_CHECK();
    std::vector<double> Eh(N);
    int ns = 0;
    CUDA_CHECK(cudaMemcpyAsync(Eh.data(), E, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(&ns, c->d_status + 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    double sum = 0.0;
    for (int i = 0; i < N; ++i) sum += Eh[i];                     // trust_region.cpp:358-362
    if (delta_out) *delta_out = delta;
    if (beta_out) *beta_out = beta;
    if (nscalar_out) *nscalar_out = ns;
    return sum;
}

// on c->Sdense (the S of the last compute_S in the callers' camera order)
double psba_launch_cholmod(psba_ctx *c, double *delta_out, double *beta_out, int *nscalar_out)
{
    return psba_launch_cholmod_dense(c, c->N, c->Sdense, c->chol_aux, c->chol_diag, c->chol_E, delta_out, beta_out, nscalar_out);
}
