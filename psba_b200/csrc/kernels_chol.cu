// kernels_chol.cu -- camera solve: tiled FP64 Cholesky of the reduced camera system S fused with
// the forward substitution, and the backward substitution.
//
// Replaces kern_cholesky / kern_cholesky_s2 / kern_trigMat_inv / kern_trigMat_mul / kern_fill_rest
// (CL_files/SPD_inv.cl:20-411, host PSBA/cl_spdinv.cpp:18-204) and kern_matVec_mul
// (matVec_mul.cl:7-17): dpa = S^-1 ea is obtained from S = L L^T by two triangular solves instead of
// the explicit inverse (parity-safe, LM costs agree to 4e-15).
//
// S lives in a pool of 48x48 tiles (8 camera blocks per tile edge); only the tiles of the symbolic
// factor exist.  The dependent chain of a Cholesky factorisation is its panel sequence, so the cost
// at these sizes (N = 6m: 42..828 for BAL, 12000 block-banded for the synthetic ring) is launch
// latency x panels.  One kernel per panel K, all panels in one CUDA graph:
//   critical CTAs (one per tile row I >= K with a tile (I,K)):
//       apply the DEFERRED update of panel K-1 to tile (I,K) and to the diagonal tile (K,K),
//       factor the diagonal tile (every CTA redundantly: no grid-wide dependency), invert the
//       factor, L_IK = A_IK L_KK^-T, and the forward substitution y_K = L_KK^-1 b_K,
//       b_I -= L_IK y_K  (the right-hand side rides along as an extra column);
//   deferred CTAs: A_IJ -= L_I,K-1 L_J,K-1^T for the trailing tiles J > K of panel K-1.
// Failure (pivot <= 0 or not finite) sets a status word; the reference reports the same event as a
// non-finite factor entry (SPD_inv.cl:66-107).  FP64 on CUDA cores: tcgen05 has no FP64 kind and
// the panels (48 wide) are latency-, not throughput-bound.
#include "dev_math.cuh"
#include <algorithm>

#define LDT (TS + 1)            // padded leading dimension in shared memory
#define TILE_SM (TS * LDT)      // doubles per shared tile
#define CHOL_SMEM (5 * TILE_SM * sizeof(double))

// ---------------------------------------------------------------------------------------------
void psba_build_tile_structure(psba_ctx *c, const std::vector<std::pair<int, int>> &pairs)
{
    const int nt = (c->m + 7) / 8;
    c->nt = nt;
    std::vector<char> present((size_t)nt * nt, 0);
    for (int I = 0; I < nt; ++I) present[(size_t)I * nt + I] = 1;
    for (auto &p : pairs) present[(size_t)(p.first / 8) * nt + p.second / 8] = 1;
    // symbolic factorisation at tile granularity
    std::vector<std::vector<int>> rows(nt);
    for (int K = 0; K < nt; ++K) {
        for (int I = K + 1; I < nt; ++I) if (present[(size_t)I * nt + K]) rows[K].push_back(I);
        for (size_t a = 0; a < rows[K].size(); ++a)
            for (size_t b = 0; b <= a; ++b) present[(size_t)rows[K][a] * nt + rows[K][b]] = 1;
    }
    std::vector<int> crit_rows, ncrI, ncrJ;
    c->crit_ptr.assign(1, 0); c->ncr_ptr.assign(1, 0);
    for (int K = 0; K < nt; ++K) {
        crit_rows.push_back(K);
        crit_rows.insert(crit_rows.end(), rows[K].begin(), rows[K].end());
        c->crit_ptr.push_back((int)crit_rows.size());
        if (K > 0)
            for (size_t a = 0; a < rows[K - 1].size(); ++a)
                for (size_t b = 0; b <= a; ++b)
                    if (rows[K - 1][b] > K) { ncrI.push_back(rows[K - 1][a]); ncrJ.push_back(rows[K - 1][b]); }
        c->ncr_ptr.push_back((int)ncrI.size());
    }
    c->h_tile_index.assign((size_t)nt * nt, -1);
    int slot = 0;
    for (int I = 0; I < nt; ++I)
        for (int J = 0; J <= I; ++J)
            if (present[(size_t)I * nt + J]) c->h_tile_index[(size_t)I * nt + J] = slot++;
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
    up(&c->d_crit_rows, crit_rows);
    up(&c->d_ncr_I, ncrI); up(&c->d_ncr_J, ncrJ);
    up(&c->d_coltile_ptr, cptr); up(&c->d_coltile_row, crow); up(&c->d_coltile_slot, cslot);
    c->d_rowtile_ptr = c->d_rowtile_col = c->d_rowtile_slot = nullptr;
    CUDA_CHECK(cudaMalloc(&c->Stiles, (size_t)c->n_tiles * TS * TS * sizeof(double)));
    CUDA_CHECK(cudaMalloc(&c->Linv, (size_t)nt * TS * TS * sizeof(double)));
    CUDA_CHECK(cudaMalloc(&c->Ldiag, (size_t)nt * TS * TS * sizeof(double)));
    c->chol_graph_ok = false;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_tile(double *sm, const double *__restrict__ g)
{
    for (int e = threadIdx.x; e < TS * TS / 2; e += 256) {
        const double2 v = reinterpret_cast<const double2 *>(g)[e];
        const int r = (2 * e) / TS, cc = (2 * e) % TS;
        sm[r * LDT + cc] = v.x; sm[r * LDT + cc + 1] = v.y;
    }
}

// C (48x48 in smem) -= A * B^T, A and B 48x48 in smem; 256 threads, 3x3 register blocks
__device__ __forceinline__ void tile_syrk_sub(double *C, const double *A, const double *B)
{
    const int tr = threadIdx.x / 16, tc = threadIdx.x % 16;
    double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll 4
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
        for (int q = 0; q < 3; ++q) C[(tr * 3 + p) * LDT + tc * 3 + q] -= acc[p][q];
}

// in-smem Cholesky of the 48x48 tile D (lower part), factor written to Lo; one barrier per column.
// Right-looking with the scaling folded into the update: D_ic -= D_ij D_cj / D_jj.  Returns false
// (in every thread) when a pivot is <= 0 or not finite.
__device__ __forceinline__ bool tile_potrf(double *D, double *Lo, int *bad)
{
    const int tid = threadIdx.x;
    if (tid == 0) *bad = 0;
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < TS; ++j) {
        const double djj = D[j * LDT + j];
        if (!(djj > 0.0) || !isfinite(djj)) { if (tid == 0) *bad = 1; break; }     // uniform: all threads read the same value
        const double inv = 1.0 / djj;
        if (tid >= j && tid < TS) Lo[tid * LDT + j] = (tid == j) ? sqrt(djj) : D[tid * LDT + j] * (sqrt(djj) * inv);
        if (tid < j) Lo[tid * LDT + j] = 0.0;
        const int rem = TS - 1 - j;
        // trailing lower triangle (j < c <= i): linear index over the rem x rem square, upper half skipped
        for (int e = tid; e < rem * rem; e += 256) {
            const int i = j + 1 + e / rem, cc = j + 1 + e % rem;
            if (cc <= i) D[i * LDT + cc] -= D[i * LDT + j] * D[cc * LDT + j] * inv;
        }
        __syncthreads();
    }
    __syncthreads();
    return *bad == 0;
}

// X = L^-1 for the lower-triangular 48x48 factor in smem (two 24x24 diagonal blocks inverted by
// substitution, one thread per column; off-diagonal block X21 = -X22 * L21 * X11 by two small GEMMs;
// T is a 24x24 scratch inside X's upper-right corner, which is zero in the result and cleared last).
__device__ __forceinline__ void tile_tri_inverse(const double *L, double *X)
{
    const int tid = threadIdx.x;
    constexpr int H = TS / 2;
    for (int e = tid; e < TS * LDT; e += 256) X[e] = 0.0;
    __syncthreads();
    if (tid < TS) {
        const int b0 = (tid / H) * H, cidx = tid % H;          // block origin, column inside the block
        const int gc = b0 + cidx;
        X[gc * LDT + gc] = 1.0 / L[gc * LDT + gc];
        for (int r = cidx + 1; r < H; ++r) {
            const int gr = b0 + r;
            double s0 = 0.0, s1 = 0.0;
            int k = cidx;
            for (; k + 1 < r; k += 2) { s0 += L[gr * LDT + b0 + k] * X[(b0 + k) * LDT + gc]; s1 += L[gr * LDT + b0 + k + 1] * X[(b0 + k + 1) * LDT + gc]; }
            if (k < r) s0 += L[gr * LDT + b0 + k] * X[(b0 + k) * LDT + gc];
            X[gr * LDT + gc] = -(s0 + s1) / L[gr * LDT + gr];
        }
    }
    __syncthreads();
    // T = L21 * X11  (24x24), stored in the upper-right corner X[0..H)[H..TS)
    for (int e = tid; e < H * H; e += 256) {
        const int r = e / H, cc = e % H;
        double s = 0.0;
        for (int k = cc; k < H; ++k) s += L[(H + r) * LDT + k] * X[k * LDT + cc];
        X[r * LDT + H + cc] = s;
    }
    __syncthreads();
    // X21 = -X22 * T
    for (int e = tid; e < H * H; e += 256) {
        const int r = e / H, cc = e % H;
        double s = 0.0;
        for (int k = 0; k <= r; ++k) s += X[(H + r) * LDT + H + k] * X[k * LDT + H + cc];
        X[(H + r) * LDT + cc] = -s;
    }
    __syncthreads();
    for (int e = tid; e < H * H; e += 256) X[(e / H) * LDT + H + e % H] = 0.0;
    __syncthreads();
}

__global__ void k_init_rhs(int N, int npad, const double *__restrict__ ea, double *__restrict__ b)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < npad) b[k] = k < N ? ea[k] : 0.0;
}

__global__ void __launch_bounds__(256) k_panel(int K, int nt, int ncrit, const int *__restrict__ crit_rows,
                                               const int *__restrict__ ncrI, const int *__restrict__ ncrJ,
                                               const int *__restrict__ tile_index, double *__restrict__ Stiles,
                                               double *__restrict__ Linv, double *__restrict__ Ldiag,
                                               double *__restrict__ bwork, double *__restrict__ ywork, int *__restrict__ status)
{
    extern __shared__ double smem[];
    __shared__ int bad;
    __shared__ double yk[TS];
    if (*status != 0) return;
    const int tid = threadIdx.x;
    double *B0 = smem, *B1 = smem + TILE_SM, *B2 = smem + 2 * TILE_SM, *B3 = smem + 3 * TILE_SM, *B4 = smem + 4 * TILE_SM;

    if ((int)blockIdx.x >= ncrit) {
        // ---- deferred trailing update of panel K-1:  A_IJ -= L_I,K-1 L_J,K-1^T
        const int t = blockIdx.x - ncrit;
        const int I = ncrI[t], J = ncrJ[t];
        double *tij = Stiles + (size_t)tile_index[I * nt + J] * TS * TS;
        load_tile(B0, Stiles + (size_t)tile_index[I * nt + (K - 1)] * TS * TS);
        load_tile(B1, Stiles + (size_t)tile_index[J * nt + (K - 1)] * TS * TS);
        load_tile(B2, tij);
        __syncthreads();
        tile_syrk_sub(B2, B0, B1);
        __syncthreads();
        for (int e = tid; e < TS * TS; e += 256) tij[e] = B2[(e / TS) * LDT + e % TS];
        return;
    }

    // ---- critical path of panel K for tile row I
    const int I = crit_rows[blockIdx.x];
    const bool have_prev = K > 0 && tile_index[K * nt + (K - 1)] >= 0;
    load_tile(B0, Stiles + (size_t)tile_index[K * nt + K] * TS * TS);          // D = A_KK (partially updated)
    if (have_prev) load_tile(B2, Stiles + (size_t)tile_index[K * nt + (K - 1)] * TS * TS);   // L_K,K-1
    __syncthreads();
    if (have_prev) { tile_syrk_sub(B0, B2, B2); __syncthreads(); }
    if (!tile_potrf(B0, B1, &bad)) { if (tid == 0) *status = 1; return; }     // B1 = L_KK
    tile_tri_inverse(B1, B0);                                                  // B0 = L_KK^-1
    // forward substitution piece: y_K = L_KK^-1 b_K
    if (tid < TS) {
        double s = 0.0;
        for (int cc = 0; cc <= tid; ++cc) s += B0[tid * LDT + cc] * bwork[K * TS + cc];
        yk[tid] = s;
    }
    __syncthreads();
    if (I == K) {
        double *ld = Ldiag + (size_t)K * TS * TS, *li = Linv + (size_t)K * TS * TS;
        for (int e = tid; e < TS * TS; e += 256) { ld[e] = B1[(e / TS) * LDT + e % TS]; li[e] = B0[(e / TS) * LDT + e % TS]; }
        if (tid < TS) ywork[K * TS + tid] = yk[tid];
        return;
    }
    // ---- off-diagonal tile: A_IK (deferred update from panel K-1), then L_IK = A_IK L_KK^-T
    double *tik = Stiles + (size_t)tile_index[I * nt + K] * TS * TS;
    load_tile(B3, tik);
    const bool upd = have_prev && tile_index[I * nt + (K - 1)] >= 0;
    if (upd) load_tile(B4, Stiles + (size_t)tile_index[I * nt + (K - 1)] * TS * TS);
    __syncthreads();
    if (upd) { tile_syrk_sub(B3, B4, B2); __syncthreads(); }
    {
        const int tr = tid / 16, tc = tid % 16;
        double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        const int kmax = tc * 3 + 3;                     // Linv is lower triangular: k <= column index
#pragma unroll 4
        for (int k = 0; k < kmax; ++k) {
            double a[3], b[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) { a[q] = B3[(tr * 3 + q) * LDT + k]; b[q] = B0[(tc * 3 + q) * LDT + k]; }
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int q = 0; q < 3; ++q) acc[p][q] += a[p] * b[q];
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) B4[(tr * 3 + p) * LDT + tc * 3 + q] = acc[p][q];
    }
    __syncthreads();
    for (int e = tid; e < TS * TS; e += 256) tik[e] = B4[(e / TS) * LDT + e % TS];
    if (tid < TS) {                                       // b_I -= L_IK y_K
        double s = 0.0;
        for (int cc = 0; cc < TS; ++cc) s += B4[tid * LDT + cc] * yk[cc];
        bwork[I * TS + tid] -= s;
    }
}

static void enqueue_factor(psba_ctx *c)
{
    const int npad = c->nt * TS;
    k_init_rhs<<<cdiv(npad, 256), 256, 0, c->stream>>>(c->N, npad, c->eab, c->chol_aux);
    for (int K = 0; K < c->nt; ++K) {
        const int ncrit = c->crit_ptr[K + 1] - c->crit_ptr[K];
        const int nncr = c->ncr_ptr[K + 1] - c->ncr_ptr[K];
        k_panel<<<ncrit + nncr, 256, CHOL_SMEM, c->stream>>>(K, c->nt, ncrit, c->d_crit_rows + c->crit_ptr[K],
                                                            c->d_ncr_I + c->ncr_ptr[K], c->d_ncr_J + c->ncr_ptr[K], c->tile_index,
                                                            c->Stiles, c->Linv, c->Ldiag, c->chol_aux, c->chol_diag, c->d_status);
    }
}

// factorise S (tile pool) and forward-substitute ea; returns 0.0 / 1.0 (synchronises)
double psba_launch_factor(psba_ctx *c)
{
    CUDA_CHECK(cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream));
    if (!c->chol_graph_ok) {
        CUDA_CHECK(cudaFuncSetAttribute(k_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_SMEM));
        cudaGraph_t graph;
        CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        enqueue_factor(c);
        CUDA_CHECK(cudaStreamEndCapture(c->stream, &graph));
        CUDA_CHECK(cudaGraphInstantiate(&c->chol_graph, graph, 0));
        CUDA_CHECK(cudaGraphDestroy(graph));
        c->chol_graph_ok = true;
    }
    PROF(c, KID_FACTOR) CUDA_CHECK(cudaGraphLaunch(c->chol_graph, c->stream));
    c->st_launches += c->nt + 1;
    int st = 0;
    CUDA_CHECK(cudaMemcpyAsync(&st, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->factor_valid = (st == 0);
    c->S_valid = false;      // the factor overwrote the tile pool
    return st ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------
// backward substitution x_I = L_II^-T (y_I - sum_{J>I} L_JI^T x_J), one persistent CTA.  Threads are
// laid out as 48 columns x 10 tile slots: every thread streams one tile column (48 independent
// coalesced loads in flight), partial sums are combined in a fixed order.
__global__ void __launch_bounds__(480) k_backward(int N, int nt, const int *__restrict__ cptr, const int *__restrict__ crow,
                                                  const int *__restrict__ cslot, const double *__restrict__ Stiles,
                                                  const double *__restrict__ Linv, double *__restrict__ ywork, double *__restrict__ sol)
{
    __shared__ double part[10][TS];
    __shared__ double acc[TS];
    const int tid = threadIdx.x, col = tid % TS, slotid = tid / TS;
    for (int I = nt - 1; I >= 0; --I) {
        double s = 0.0;
        for (int t = cptr[I] + slotid; t < cptr[I + 1]; t += 10) {
            const double *L = Stiles + (size_t)cslot[t] * TS * TS + col;
            const double *x = ywork + crow[t] * TS;
#pragma unroll 8
            for (int r = 0; r < TS; ++r) s += L[r * TS] * x[r];
        }
        part[slotid][col] = s;
        __syncthreads();
        if (tid < TS) {
            double a = 0.0;
#pragma unroll
            for (int p = 0; p < 10; ++p) a += part[p][tid];
            acc[tid] = ywork[I * TS + tid] - a;
        }
        __syncthreads();
        if (tid < TS) {
            const double *inv = Linv + (size_t)I * TS * TS + tid;     // column tid of L_II^-1 = row of its transpose
            double a = 0.0;
#pragma unroll 8
            for (int r = tid; r < TS; ++r) a += inv[r * TS] * acc[r];
            ywork[I * TS + tid] = a;
            const int gr = I * TS + tid;
            if (gr < N) sol[gr] = a;
        }
        __syncthreads();
    }
}

void psba_launch_solve(psba_ctx *c)
{
    PROF(c, KID_TRI_SOLVE) k_backward<<<1, 480, 0, c->stream>>>(c->N, c->nt, c->d_coltile_ptr, c->d_coltile_row, c->d_coltile_slot,
                                                               c->Stiles, c->Linv, c->chol_diag, c->dp);
    c->st_launches += 1;
}
