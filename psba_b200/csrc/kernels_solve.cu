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
void psba_build_tile_structure(psba_ctx *c, const std::vector<std::pair<int, int>> &pairs)
{
    const int nt = (c->m + 7) / 8;
    c->nt = nt;
    std::vector<char> present((size_t)nt * nt, 0);
    for (int I = 0; I < nt; ++I) present[(size_t)I * nt + I] = 1;
    for (auto &p : pairs) present[(size_t)(p.first / 8) * nt + p.second / 8] = 1;
    c->panel_row_ptr.assign(1, 0); c->panel_rows.clear();
    c->panel_upd_ptr.assign(1, 0);
    std::vector<int> updI, updJ;
    for (int K = 0; K < nt; ++K) {
        std::vector<int> rows;
        for (int I = K + 1; I < nt; ++I) if (present[(size_t)I * nt + K]) rows.push_back(I);
        for (size_t a = 0; a < rows.size(); ++a)
            for (size_t b = 0; b <= a; ++b) {
                present[(size_t)rows[a] * nt + rows[b]] = 1;
                updI.push_back(rows[a]); updJ.push_back(rows[b]);
            }
        c->panel_rows.insert(c->panel_rows.end(), rows.begin(), rows.end());
        c->panel_row_ptr.push_back((int)c->panel_rows.size());
        c->panel_upd_ptr.push_back((int)updI.size());
    }
    c->h_tile_index.assign((size_t)nt * nt, -1);
    int slot = 0;
    std::vector<int> rptr(1, 0), rcol, rslot;
    for (int I = 0; I < nt; ++I) {
        for (int J = 0; J <= I; ++J)
            if (present[(size_t)I * nt + J]) {
                c->h_tile_index[(size_t)I * nt + J] = slot;
                if (J < I) { rcol.push_back(J); rslot.push_back(slot); }
                ++slot;
            }
        rptr.push_back((int)rcol.size());
    }
    std::vector<int> cptr(1, 0), crow, cslot;
    for (int J = 0; J < nt; ++J) {
        for (int I = J + 1; I < nt; ++I)
            if (present[(size_t)I * nt + J]) { crow.push_back(I); cslot.push_back(c->h_tile_index[(size_t)I * nt + J]); }
        cptr.push_back((int)crow.size());
    }
    c->n_tiles = slot;
    auto up = [&](int **d, const std::vector<int> &h) {
        CUDA_CHECK(cudaMalloc(d, std::max<size_t>(1, h.size()) * sizeof(int)));
        if (!h.empty()) CUDA_CHECK(cudaMemcpy(*d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
    };
    up(&c->tile_index, c->h_tile_index);
    up(&c->d_panel_rows, c->panel_rows);
    up(&c->d_upd_I, updI); up(&c->d_upd_J, updJ);
    up(&c->d_rowtile_ptr, rptr); up(&c->d_rowtile_col, rcol); up(&c->d_rowtile_slot, rslot);
    up(&c->d_coltile_ptr, cptr); up(&c->d_coltile_row, crow); up(&c->d_coltile_slot, cslot);
    CUDA_CHECK(cudaMalloc(&c->Stiles, (size_t)c->n_tiles * TS * TS * sizeof(double)));
    CUDA_CHECK(cudaMalloc(&c->Linv, (size_t)nt * TS * TS * sizeof(double)));
    c->chol_graph_ok = false;
}

// ---------------------------------------------------------------------------------------------
// potrf of the diagonal tile K (in shared memory) + inverse of its factor
__global__ void __launch_bounds__(256) k_potrf_diag(int K, int nt, const int *__restrict__ tile_index,
                                                    double *__restrict__ Stiles, double *__restrict__ Linv, int *__restrict__ status)
{
    __shared__ double A[TS * LDT];
    __shared__ int bad;
    if (*status != 0) return;
    const int tid = threadIdx.x;
    double *tile = Stiles + (size_t)tile_index[K * nt + K] * TS * TS;
    for (int e = tid; e < TS * TS; e += 256) A[(e / TS) * LDT + (e % TS)] = tile[e];
    if (tid == 0) bad = 0;
    __syncthreads();
    for (int j = 0; j < TS; ++j) {
        if (tid == 0) {
            double d = A[j * LDT + j];
            if (!(d > 0.0) || !isfinite(d)) bad = 1;
            A[j * LDT + j] = sqrt(d);
        }
        __syncthreads();
        if (bad) break;
        const double djj = A[j * LDT + j];
        if (tid > j && tid < TS) A[tid * LDT + j] /= djj;
        __syncthreads();
        // trailing lower triangle: (r,c), j < c <= r
        const int rem = TS - 1 - j;
        for (int e = tid; e < rem * rem; e += 256) {
            const int r = j + 1 + e / rem, cc = j + 1 + e % rem;
            if (cc <= r) A[r * LDT + cc] -= A[r * LDT + j] * A[cc * LDT + j];
        }
        __syncthreads();
    }
    if (bad) { if (tid == 0) *status = 1; return; }
    for (int e = tid; e < TS * TS; e += 256) {
        const int r = e / TS, cc = e % TS;
        tile[e] = (cc <= r) ? A[r * LDT + cc] : 0.0;
    }
    // column cidx of L^-1 by forward substitution (one thread per column)
    double *inv = Linv + (size_t)K * TS * TS;
    if (tid < TS) {
        const int cidx = tid;
        double x[TS];
#pragma unroll 1
        for (int r = 0; r < TS; ++r) {
            double v;
            if (r < cidx) v = 0.0;
            else if (r == cidx) v = 1.0 / A[r * LDT + r];
            else {
                double s = 0.0;
                for (int k = cidx; k < r; ++k) s += A[r * LDT + k] * x[k];
                v = -s / A[r * LDT + r];
            }
            x[r] = v;
            inv[r * TS + cidx] = v;
        }
    }
}

// L_IK = A_IK * Linv_KK^T for the tiles below the diagonal of panel K
__global__ void __launch_bounds__(256) k_trsm_tiles(int K, int nt, const int *__restrict__ rows, const int *__restrict__ tile_index,
                                                    double *__restrict__ Stiles, const double *__restrict__ Linv, const int *__restrict__ status)
{
    __shared__ double A[TS * LDT];
    __shared__ double Li[TS * LDT];
    if (*status != 0) return;
    const int tid = threadIdx.x;
    const int I = rows[blockIdx.x];
    double *tile = Stiles + (size_t)tile_index[I * nt + K] * TS * TS;
    const double *inv = Linv + (size_t)K * TS * TS;
    for (int e = tid; e < TS * TS; e += 256) { A[(e / TS) * LDT + (e % TS)] = tile[e]; Li[(e / TS) * LDT + (e % TS)] = inv[e]; }
    __syncthreads();
    const int tr = tid / 16, tc = tid % 16;
    double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < TS; ++k) {
        double a[3], b[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) { a[q] = A[(tr * 3 + q) * LDT + k]; b[q] = Li[(tc * 3 + q) * LDT + k]; }
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) acc[p][q] += a[p] * b[q];
    }
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) tile[(tr * 3 + p) * TS + tc * 3 + q] = acc[p][q];
}

// A_IJ -= L_IK * L_JK^T for the trailing tiles of panel K
__global__ void __launch_bounds__(256) k_update_tiles(int K, int nt, const int *__restrict__ updI, const int *__restrict__ updJ,
                                                      const int *__restrict__ tile_index, double *__restrict__ Stiles,
                                                      const int *__restrict__ status)
{
    __shared__ double A[TS * LDT];
    __shared__ double B[TS * LDT];
    if (*status != 0) return;
    const int tid = threadIdx.x;
    const int I = updI[blockIdx.x], J = updJ[blockIdx.x];
    const double *ta = Stiles + (size_t)tile_index[I * nt + K] * TS * TS;
    const double *tb = Stiles + (size_t)tile_index[J * nt + K] * TS * TS;
    double *tcij = Stiles + (size_t)tile_index[I * nt + J] * TS * TS;
    for (int e = tid; e < TS * TS; e += 256) { A[(e / TS) * LDT + (e % TS)] = ta[e]; B[(e / TS) * LDT + (e % TS)] = tb[e]; }
    __syncthreads();
    const int tr = tid / 16, tc = tid % 16;
    double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < TS; ++k) {
        double a[3], b[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) { a[q] = A[(tr * 3 + q) * LDT + k]; b[q] = B[(tc * 3 + q) * LDT + k]; }
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) acc[p][q] += a[p] * b[q];
    }
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) tcij[(tr * 3 + p) * TS + tc * 3 + q] -= acc[p][q];
}

static void enqueue_factor(psba_ctx *c)
{
    for (int K = 0; K < c->nt; ++K) {
        k_potrf_diag<<<1, 256, 0, c->stream>>>(K, c->nt, c->tile_index, c->Stiles, c->Linv, c->d_status);
        const int nr = c->panel_row_ptr[K + 1] - c->panel_row_ptr[K];
        if (nr > 0)
            k_trsm_tiles<<<nr, 256, 0, c->stream>>>(K, c->nt, c->d_panel_rows + c->panel_row_ptr[K], c->tile_index, c->Stiles,
                                                   c->Linv, c->d_status);
        const int nu = c->panel_upd_ptr[K + 1] - c->panel_upd_ptr[K];
        if (nu > 0)
            k_update_tiles<<<nu, 256, 0, c->stream>>>(K, c->nt, c->d_upd_I + c->panel_upd_ptr[K], c->d_upd_J + c->panel_upd_ptr[K],
                                                     c->tile_index, c->Stiles, c->d_status);
    }
}

double psba_launch_factor(psba_ctx *c)
{
    CUDA_CHECK(cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream));
    if (!c->chol_graph_ok) {
        cudaGraph_t graph;
        CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        enqueue_factor(c);
        CUDA_CHECK(cudaStreamEndCapture(c->stream, &graph));
        CUDA_CHECK(cudaGraphInstantiate(&c->chol_graph, graph, 0));
        CUDA_CHECK(cudaGraphDestroy(graph));
        c->chol_graph_ok = true;
    }
    PROF(c, KID_FACTOR) CUDA_CHECK(cudaGraphLaunch(c->chol_graph, c->stream));
    c->st_launches += 3 * c->nt;
    int st = 0;
    CUDA_CHECK(cudaMemcpyAsync(&st, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->factor_valid = (st == 0);
    c->S_valid = false;      // the factor overwrote the tile pool
    return st ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------
// dpa = S^-1 ea with the tiled factor: forward then backward substitution, one persistent CTA.
// 16 warps; warp w owns rows w, w+16, w+32 of the current tile row, lanes stride the columns.
__global__ void __launch_bounds__(512) k_tri_solve(int N, int nt, const int *__restrict__ tile_index,
                                                   const int *__restrict__ rptr, const int *__restrict__ rcol, const int *__restrict__ rslot,
                                                   const int *__restrict__ cptr, const int *__restrict__ crow, const int *__restrict__ cslot,
                                                   const double *__restrict__ Stiles, const double *__restrict__ Linv,
                                                   const double *__restrict__ rhs, double *__restrict__ ywork, double *__restrict__ sol)
{
    __shared__ double acc[TS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // forward: y_I = Linv_II (b_I - sum_{J<I} L_IJ y_J)
    for (int I = 0; I < nt; ++I) {
        for (int r = warp; r < TS; r += 16) {
            double s = 0.0;
            for (int t = rptr[I]; t < rptr[I + 1]; ++t) {
                const double *L = Stiles + (size_t)rslot[t] * TS * TS + r * TS;
                const double *y = ywork + rcol[t] * TS;
                for (int cc = lane; cc < TS; cc += 32) s += L[cc] * y[cc];
            }
#pragma unroll
            for (int w = 16; w > 0; w >>= 1) s += __shfl_down_sync(0xffffffffu, s, w);
            if (lane == 0) { const int gr = I * TS + r; acc[r] = (gr < N ? rhs[gr] : 0.0) - s; }
        }
        __syncthreads();
        if (tid < TS) {
            const double *inv = Linv + (size_t)I * TS * TS + tid * TS;
            double s = 0.0;
            for (int cc = 0; cc <= tid; ++cc) s += inv[cc] * acc[cc];
            ywork[I * TS + tid] = s;
        }
        __syncthreads();
    }
    // backward: x_I = Linv_II^T (y_I - sum_{J>I} L_JI^T x_J); x overwrites ywork
    for (int I = nt - 1; I >= 0; --I) {
        if (tid < TS) acc[tid] = 0.0;
        __syncthreads();
        // thread (col = tid % 48, part = tid / 48): partial sums over tile rows r = part, part+10, ...
        {
            const int col = tid % TS, part = tid / TS;    // 512 threads -> parts 0..9 (+ 32 idle)
            if (part < 10) {
                double s = 0.0;
                for (int t = cptr[I]; t < cptr[I + 1]; ++t) {
                    const double *L = Stiles + (size_t)cslot[t] * TS * TS;
                    const double *x = ywork + crow[t] * TS;
                    for (int r = part; r < TS; r += 10) s += L[r * TS + col] * x[r];
                }
                // fixed-order combination of the 10 parts
                for (int p = 0; p < 10; ++p) {
                    if (part == p) acc[col] += s;
                    __syncthreads();
                }
            } else {
                for (int p = 0; p < 10; ++p) __syncthreads();
            }
        }
        if (tid < TS) acc[tid] = ywork[I * TS + tid] - acc[tid];
        __syncthreads();
        if (tid < TS) {
            const double *inv = Linv + (size_t)I * TS * TS;
            double s = 0.0;
            for (int r = tid; r < TS; ++r) s += inv[r * TS + tid] * acc[r];
            ywork[I * TS + tid] = s;
            const int gr = I * TS + tid;
            if (gr < N) sol[gr] = s;
        }
        __syncthreads();
    }
}

void psba_launch_solve(psba_ctx *c)
{
    PROF(c, KID_TRI_SOLVE) k_tri_solve<<<1, 512, 0, c->stream>>>(c->N, c->nt, c->tile_index, c->d_rowtile_ptr, c->d_rowtile_col, c->d_rowtile_slot,
                                         c->d_coltile_ptr, c->d_coltile_row, c->d_coltile_slot, c->Stiles, c->Linv,
                                         c->eab, c->chol_aux, c->dp);
    c->st_launches += 1;
}

// ---------------------------------------------------------------------------------------------
// tile pool -> dense N x N row-major (lower triangle; mirror fills the upper one)
__global__ void k_tiles_to_dense(int N, int nt, const int *__restrict__ tile_index, const double *__restrict__ Stiles,
                                 double *__restrict__ dense, int mirror)
{
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)N * N) return;
    const int r = (int)(e / N), cc = (int)(e % N);
    int rr = r, c2 = cc;
    if (cc > r) { if (!mirror) { dense[e] = 0.0; return; } rr = cc; c2 = r; }
    const int slot = tile_index[(rr / TS) * nt + c2 / TS];
    dense[e] = slot < 0 ? 0.0 : Stiles[(size_t)slot * TS * TS + (rr % TS) * TS + c2 % TS];
}

void psba_tiles_to_dense(psba_ctx *c, double *dense_dev, bool mirror)
{
    long long tot = (long long)c->N * c->N;
    k_tiles_to_dense<<<cdiv(tot, 256), 256, 0, c->stream>>>(c->N, c->nt, c->tile_index, c->Stiles, dense_dev, mirror ? 1 : 0);
    c->st_launches += 1;
}

// ABI parity only (SPDinv's explicit inverse, cl_spdinv.cpp:18-40): column cidx of S^-1 by one
// forward and one backward substitution on the dense factor.  O(N^2) per thread; small N only.
__global__ void k_explicit_inverse(int N, const double *__restrict__ L, double *__restrict__ work, double *__restrict__ out)
{
    int cidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx >= N) return;
    double *x = work + (size_t)cidx * N;
    for (int r = 0; r < N; ++r) {
        double s = (r == cidx) ? 1.0 : 0.0;
        for (int k = 0; k < r; ++k) s -= L[(size_t)r * N + k] * x[k];
        x[r] = s / L[(size_t)r * N + r];
    }
    for (int r = N - 1; r >= 0; --r) {
        double s = x[r];
        for (int k = r + 1; k < N; ++k) s -= L[(size_t)k * N + r] * x[k];
        x[r] = s / L[(size_t)r * N + r];
    }
    for (int r = 0; r < N; ++r) out[(size_t)r * N + cidx] = x[r];
}

void psba_launch_explicit_inverse(psba_ctx *c, double *out_dev)
{
    const size_t nn = (size_t)c->N * c->N;
    if (!c->Sdense) CUDA_CHECK(cudaMalloc(&c->Sdense, nn * sizeof(double)));
    if (!c->Sdense_aux) CUDA_CHECK(cudaMalloc(&c->Sdense_aux, nn * sizeof(double)));
    psba_tiles_to_dense(c, c->Sdense, false);
    k_explicit_inverse<<<cdiv(c->N, 64), 64, 0, c->stream>>>(c->N, c->Sdense, c->Sdense_aux, out_dev);
    c->st_launches += 1;
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
            double t = fabs(mat[(size_t)r * N + k]);
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

__global__ void __launch_bounds__(1024) k_cholmod(int N, double *__restrict__ mat, double *__restrict__ aux, double *__restrict__ diagInv,
                                                  double *__restrict__ diag, double beta, double delta, int *__restrict__ nscalar_out)
{
    __shared__ double T[9], L[6], inv[9];
    __shared__ int fail, over, flagged;
    __shared__ double theta_s;
    const int tid = threadIdx.x, nb = N / 3;
    int nscalar = 0;
    for (int j = 0; j < nb; ++j) {
        double *d = mat + (size_t)(j * 3) * N + j * 3;
        if (tid < 9) {   // back up A_jj, diag; T_jj (cholmod_blk.cl:107-129)
            const int u = tid / 3, v = tid % 3;
            double sum = d[(size_t)u * N + v];
            aux[j * 9 + tid] = sum;
            if (u == v) diag[j * 3 + u] = sum;
            for (int k = 0; k < j; ++k) sum -= dot3g(mat + (size_t)(j * 3 + u) * N + k * 3, mat + (size_t)(j * 3 + v) * N + k * 3);
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
                d[0] = L[0]; d[1] = 0; d[2] = 0;
                d[N] = L[1]; d[N + 1] = L[2]; d[N + 2] = 0;
                d[2 * (size_t)N] = L[3]; d[2 * (size_t)N + 1] = L[4]; d[2 * (size_t)N + 2] = L[5];
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
            for (int rr = tid; rr < nrows; rr += 1024) {
                const int row = (j + 1) * 3 + rr;
                const int i = row / 3, u = row % 3;
                double Tij[3];
                for (int v = 0; v < 3; ++v) {
                    double sum = mat[(size_t)row * N + j * 3 + v];
                    aux[i * 9 + u * 3 + v] = sum;
                    for (int k = 0; k < j; ++k) sum -= dot3g(mat + (size_t)row * N + k * 3, mat + (size_t)(j * 3 + v) * N + k * 3);
                    Tij[v] = sum;
                }
                for (int v = 0; v < 3; ++v) {
                    const double sum = Tij[0] * inv[v * 3] + Tij[1] * inv[v * 3 + 1] + Tij[2] * inv[v * 3 + 2];
                    mat[(size_t)row * N + j * 3 + v] = sum;
                    mat[(size_t)(j * 3 + v) * N + row] = 0;
                    if (sum > beta) over = 1;          // signed compare, SURVEY A.4
                }
            }
            __syncthreads();
            if (over) {   // step 3 failure branch (cholmod_blk.cl:386-414)
                for (int rr = tid; rr < nrows; rr += 1024) {
                    const int row = (j + 1) * 3 + rr;
                    const int i = row / 3, u = row % 3;
                    for (int v = 0; v < 3; ++v) mat[(size_t)row * N + j * 3 + v] = aux[i * 9 + u * 3 + v];
                }
                if (tid == 0) {
                    d[0] = aux[j * 9];
                    d[N] = aux[j * 9 + 3]; d[N + 1] = aux[j * 9 + 4];
                    d[2 * (size_t)N] = aux[j * 9 + 6]; d[2 * (size_t)N + 1] = aux[j * 9 + 7]; d[2 * (size_t)N + 2] = aux[j * 9 + 8];
                }
                scalar = true;
            }
            __syncthreads();
        }
        if (scalar) {   // scalar Gill-Murray path for the three columns (cholmod_blk.cl:446-697)
            ++nscalar;
            for (int col = 0; col < 3; ++col) {
                const int x = j * 3 + col;
                const size_t jj = (size_t)x * N + x;
                __syncthreads();
                if (tid == 0) {
                    double sum = mat[jj];
                    for (int k = 0; k < x; ++k) { const double Lk = mat[(size_t)x * N + k]; sum -= Lk * Lk; }
                    sum = fabs(sum);
                    const double dj = fmax(sum, delta);
                    aux[x] = dj;
                    mat[jj] = sqrt(dj);
                    flagged = 0;
                }
                __syncthreads();
                const double ljj = mat[jj];
                for (int i = x + 1 + tid; i < N; i += 1024) {
                    double C = mat[(size_t)i * N + x];
                    for (int k = 0; k < x; ++k) C = C - (mat[(size_t)i * N + k] * mat[(size_t)x * N + k]);
                    aux[N + i] = C;
                    const double lij = C / ljj;
                    mat[(size_t)i * N + x] = lij;
                    mat[(size_t)x * N + i] = 0;
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
                    for (int i = x + 1 + tid; i < N; i += 1024) mat[(size_t)i * N + x] = aux[N + i] / theta_s;
                }
                __syncthreads();
            }
            if (tid == 0) {   // kern_cholmod_diaginv (cholmod_blk.cl:703-763)
                L[0] = d[0]; L[1] = d[N]; L[2] = d[N + 1];
                L[3] = d[2 * (size_t)N]; L[4] = d[2 * (size_t)N + 1]; L[5] = d[2 * (size_t)N + 2];
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
    for (int k = 0; k <= i; ++k) sum += mat[(size_t)i * N + k] * mat[(size_t)i * N + k];
    diag[i] = sum - diag[i];
}

// runs on c->Sdense (dense S incl. mirrored upper triangle). Returns sum_i E_i (left-to-right).
double psba_launch_cholmod(psba_ctx *c, double *delta_out, double *beta_out, int *nscalar_out)
{
    const int N = c->N;
    k_mat_max<<<1, 256, 0, c->stream>>>(N, c->Sdense, c->d_scal + 8);
    CUDA_CHECK(cudaMemcpyAsync(c->h_scal + 8, c->d_scal + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    const double xi = c->h_scal[8], gamma = c->h_scal[9];
    double delta = 1e-15 * fmax(xi + gamma, 1.0);                 // cl_cholmod.cpp:161-164
    double beta = fmax(gamma, 1e-15);
    beta = fmax(beta, xi / sqrt((double)N * N - 1));
    beta = sqrt(beta);
    PROF(c, KID_CHOLMOD) k_cholmod<<<1, 1024, 0, c->stream>>>(N, c->Sdense, c->chol_aux, c->chol_diag, c->chol_E, beta, delta, c->d_status + 2);
    k_cholmod_E<<<cdiv(N, 128), 128, 0, c->stream>>>(N, c->Sdense, c->chol_E);
    c->st_launches += 3;
    std::vector<double> E(N);
    int ns = 0;
    CUDA_CHECK(cudaMemcpyAsync(E.data(), c->chol_E, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(&ns, c->d_status + 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    double sum = 0.0;
    for (int i = 0; i < N; ++i) sum += E[i];                      // trust_region.cpp:358-362
    if (delta_out) *delta_out = delta;
    if (beta_out) *beta_out = beta;
    if (nscalar_out) *nscalar_out = ns;
    return sum;
}
