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
#define CHOL_SMEM (2 * TILE_SM * sizeof(double))   // two staged tiles (factor tiles of panel K-1, then D and A_IK)

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
// a 48x48 tile travels global -> registers (all loads in flight at once) -> padded shared memory
template <int NT> struct TileRegs { double2 v[(TS * TS / 2 + NT - 1) / NT]; };
template <int NT>
__device__ __forceinline__ void tile_ldg(TileRegs<NT> &t, const double *__restrict__ g)
{
    constexpr int Q = (TS * TS / 2 + NT - 1) / NT;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int e = threadIdx.x + q * NT;
        t.v[q] = e < TS * TS / 2 ? reinterpret_cast<const double2 *>(g)[e] : make_double2(0.0, 0.0);
    }
}
template <int NT>
__device__ __forceinline__ void tile_sts(double *sm, const TileRegs<NT> &t)
{
    constexpr int Q = (TS * TS / 2 + NT - 1) / NT;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int e = threadIdx.x + q * NT;
        if (e < TS * TS / 2) { const int r = (2 * e) / TS, cc = (2 * e) % TS; sm[r * LDT + cc] = t.v[q].x; sm[r * LDT + cc + 1] = t.v[q].y; }
    }
}
// lower_only: entries above the diagonal are written as zero (diagonal factor tiles)
template <int NT>
__device__ __forceinline__ void tile_stg(double *__restrict__ g, const double *sm, bool lower_only = false)
{
    constexpr int Q = (TS * TS / 2 + NT - 1) / NT;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int e = threadIdx.x + q * NT;
        if (e < TS * TS / 2) {
            const int r = (2 * e) / TS, cc = (2 * e) % TS;
            double x = sm[r * LDT + cc], y = sm[r * LDT + cc + 1];
            if (lower_only) { if (cc > r) x = 0.0; if (cc + 1 > r) y = 0.0; }
            reinterpret_cast<double2 *>(g)[e] = make_double2(x, y);
        }
    }
}

// 128 threads as an 8 x 16 grid: thread (tr = tid % 8, tc = tid / 8) owns rows 6tr..6tr+5, columns
// 3tc..3tc+2 of a tile (a warp covers four adjacent column triples, so warps retire as the sweep advances).
// acc = A * B^T (and acc2 = A2 * B^T when TWO) for 48x48 tiles in smem.
template <bool TWO>
__device__ __forceinline__ void tile_abt(const double *A, const double *A2, const double *B, double acc[6][3], double acc2[6][3])
{
    const int tr = threadIdx.x % 8, tc = threadIdx.x / 8;
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) { acc[p][q] = 0.0; if (TWO) acc2[p][q] = 0.0; }
#pragma unroll 4
    for (int k = 0; k < TS; ++k) {
        double a[6], a2[6], b[3];
#pragma unroll
        for (int p = 0; p < 6; ++p) { a[p] = A[(tr * 6 + p) * LDT + k]; if (TWO) a2[p] = A2[(tr * 6 + p) * LDT + k]; }
#pragma unroll
        for (int q = 0; q < 3; ++q) b[q] = B[(tc * 3 + q) * LDT + k];
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) { acc[p][q] += a[p] * b[q]; if (TWO) acc2[p][q] += a2[p] * b[q]; }
    }
}

// X = L^-1 for the lower-triangular 48x48 factor in smem, by recursive doubling: the eight 6x6 diagonal
// blocks are inverted by substitution (one thread per column, <= 5 dependent steps), then blocks are
// merged pairwise (h = 6, 12, 24):  inv [[A,0],[B,C]] = [[A^-1,0],[-C^-1 B A^-1, C^-1]].  The h x h
// scratch T = B A^-1 lives in the upper-right corner of the pair, which is zero in the result.
template <int H>
__device__ __forceinline__ void tri_inverse_merge(const double *L, double *X)
{
    const int tid = threadIdx.x;
    constexpr int NP = TS / (2 * H), PER = H * H;
    for (int e = tid; e < NP * PER; e += 256) {               // T = L21 * X11
        const int o = (e / PER) * 2 * H, r = (e % PER) / H, cc = e % H;
        double s = 0.0;
        for (int k = cc; k < H; ++k) s += L[(o + H + r) * LDT + o + k] * X[(o + k) * LDT + o + cc];
        X[(o + r) * LDT + o + H + cc] = s;
    }
    __syncthreads();
    for (int e = tid; e < NP * PER; e += 256) {               // X21 = -X22 * T
        const int o = (e / PER) * 2 * H, r = (e % PER) / H, cc = e % H;
        double s = 0.0;
        for (int k = 0; k <= r; ++k) s += X[(o + H + r) * LDT + o + H + k] * X[(o + k) * LDT + o + H + cc];
        X[(o + H + r) * LDT + o + cc] = -s;
    }
    __syncthreads();
    for (int e = tid; e < NP * PER; e += 256) {
        const int o = (e / PER) * 2 * H, r = (e % PER) / H, cc = e % H;
        X[(o + r) * LDT + o + H + cc] = 0.0;
    }
    __syncthreads();
}

__device__ __forceinline__ void tile_tri_inverse(const double *L, double *X)
{
    const int tid = threadIdx.x;
    for (int e = tid; e < TS * LDT; e += 256) X[e] = 0.0;
    __syncthreads();
    if (tid < TS) {
        const int b0 = (tid / 6) * 6, cidx = tid % 6, gc = b0 + cidx;
        double x[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double v = 0.0;
            if (r == cidx) v = 1.0 / L[gc * LDT + gc];
            else if (r > cidx) {
                double sum = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) if (k >= cidx && k < r) sum += L[(b0 + r) * LDT + b0 + k] * x[k];
                v = -sum / L[(b0 + r) * LDT + b0 + r];
            }
            x[r] = v;
            X[(b0 + r) * LDT + gc] = v;
        }
    }
    __syncthreads();
    tri_inverse_merge<6>(L, X);
    tri_inverse_merge<12>(L, X);
    tri_inverse_merge<24>(L, X);
}

__global__ void k_init_rhs(int N, int npad, const double *__restrict__ ea, double *__restrict__ b)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < npad) b[k] = k < N ? ea[k] : 0.0;
}

#define PANEL_NT 128
// One kernel per panel K (128 threads per CTA).
//  critical CTA for tile row I (I = K: the diagonal CTA): loads D = A_KK, A_IK and the factor tiles of
//  panel K-1, applies the deferred updates, then ONE column sweep factors the stacked panel
//  [D ; A_IK ; b_K^T] (97 x 48): l_ij = d_ij * rsqrt(d_jj), d_ic -= l_ij l_cj.  Rows of D become L_KK,
//  rows of A_IK become L_IK = A_IK L_KK^-T (no explicit inverse, no separate trsm) and the extra row
//  b_K^T becomes y_K^T = (L_KK^-1 b_K)^T, the forward substitution.  Every thread keeps one 6x3 block
//  of D and one of A_IK in registers; per column only the 97 column entries go through shared memory
//  (double-buffered, one barrier per column); the loop body is branch-free (dead entries are updated
//  too, a non-positive pivot poisons the panel with NaN and is reported once at the end).
//  deferred CTAs: A_IJ -= L_I,K-1 L_J,K-1^T for the trailing tiles (J > K) of panel K-1.
__global__ void __launch_bounds__(PANEL_NT) k_panel(int K, int nt, int ncrit, const int *__restrict__ crit_rows,
                                                    const int *__restrict__ ncrI, const int *__restrict__ ncrJ,
                                                    const int *__restrict__ tile_index, double *__restrict__ Stiles,
                                                    double *__restrict__ Ldiag, double *__restrict__ bwork,
                                                    double *__restrict__ ywork, int *__restrict__ status)
{
    extern __shared__ double smem[];
    __shared__ __align__(16) double colbuf[2][2 * TS + 4];
    if (*status != 0) return;
    const int tid = threadIdx.x, tr = tid % 8, tc = tid / 8;
    double *B0 = smem, *B1 = smem + TILE_SM;

    if ((int)blockIdx.x >= ncrit) {
        // ---- deferred trailing update of panel K-1:  A_IJ -= L_I,K-1 L_J,K-1^T
        const int t = blockIdx.x - ncrit;
        const int I = ncrI[t], J = ncrJ[t];
        double *tij = Stiles + (size_t)tile_index[I * nt + J] * TS * TS;
        TileRegs<PANEL_NT> ra, rb;
        tile_ldg(ra, Stiles + (size_t)tile_index[I * nt + (K - 1)] * TS * TS);
        tile_ldg(rb, Stiles + (size_t)tile_index[J * nt + (K - 1)] * TS * TS);
        double c18[6][3];
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) c18[p][q] = tij[(tr * 6 + p) * TS + tc * 3 + q];
        tile_sts(B0, ra); tile_sts(B1, rb);
        __syncthreads();
        double acc[6][3];
        tile_abt<false>(B0, B0, B1, acc, acc);
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) tij[(tr * 6 + p) * TS + tc * 3 + q] = c18[p][q] - acc[p][q];
        return;
    }

    // ---- critical path of panel K for tile row I
    const int I = crit_rows[blockIdx.x];
    const bool diagcta = I == K;
    const bool have_prev = K > 0 && tile_index[K * nt + (K - 1)] >= 0;
    const bool upd = !diagcta && have_prev && tile_index[I * nt + (K - 1)] >= 0;
    double *tik = Stiles + (size_t)tile_index[I * nt + K] * TS * TS;
    const double *tkk = Stiles + (size_t)tile_index[K * nt + K] * TS * TS;
    // all global loads are issued before anything waits on them
    TileRegs<PANEL_NT> rp, ri;
    if (have_prev) tile_ldg(rp, Stiles + (size_t)tile_index[K * nt + (K - 1)] * TS * TS);    // L_K,K-1
    if (upd) tile_ldg(ri, Stiles + (size_t)tile_index[I * nt + (K - 1)] * TS * TS);          // L_I,K-1
    double d[6][3], a[6][3];
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            d[p][q] = tkk[(tr * 6 + p) * TS + tc * 3 + q];
            a[p][q] = diagcta ? 0.0 : tik[(tr * 6 + p) * TS + tc * 3 + q];
        }
    if (have_prev) {
        tile_sts(B0, rp);
        if (upd) tile_sts(B1, ri);
        __syncthreads();
        double acc[6][3], acc2[6][3];
        if (upd) tile_abt<true>(B0, B1, B0, acc, acc2);
        else tile_abt<false>(B0, B0, B0, acc, acc2);
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) { d[p][q] -= acc[p][q]; if (upd) a[p][q] -= acc2[p][q]; }
    }
    // ---- hand the updated blocks over to the ROW-OWNER layout of the sweep: thread t < 48 owns row t of
    // D, thread 48+r owns row r of A_IK, thread 96 owns the right-hand-side row b_K^T
    __syncthreads();                                           // the factor tiles in B0/B1 are consumed
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) { B0[(tr * 6 + p) * LDT + tc * 3 + q] = d[p][q]; B1[(tr * 6 + p) * LDT + tc * 3 + q] = a[p][q]; }
    __syncthreads();
    const bool active = tid < TS || (tid < 2 * TS && !diagcta) || tid == 2 * TS;
    double row[TS];
    if (tid == 2 * TS) {
#pragma unroll
        for (int c = 0; c < TS; ++c) row[c] = bwork[K * TS + c];
    } else {
        const double *src = tid < TS ? B0 + tid * LDT : B1 + (tid - TS) * LDT;
#pragma unroll
        for (int c = 0; c < TS; ++c) row[c] = active ? src[c] : 0.0;
    }
    // ---- column sweep over the stacked panel [D ; A_IK ; b^T] (97 x 48), one row per thread in registers.
    // Step j: every thread publishes its entry of column j (ONE 8-byte shared store per thread: the
    // barrier drains pending stores at ~19 cycles each, see tools/microbench/sweep_bench.cu), then
    //   inv = 1/d_jj,  f = row[j]*inv,  row[c] -= f * col[c]  (c > j; the next column first),
    // and finally row[j] *= rsqrt(d_jj) turns the entry into the factor value (off the critical chain).
    bool bad = false;
    if (active) colbuf[0][tid] = row[0];
#pragma unroll
    for (int j = 0; j < TS; ++j) {
        __syncthreads();
        if (active) {
            const double *cb = colbuf[j & 1];
            const double piv = cb[j];
            bad |= !(piv > 0.0 && piv < 1e300);
            const double f = row[j] * __drcp_rn(piv);
            if (j + 1 < TS) {
                row[j + 1] -= f * cb[j + 1];
                colbuf[(j + 1) & 1][tid] = row[j + 1];         // publish the next column before the bulk update
            }
#pragma unroll
            for (int c = j + 2; c < TS; ++c) row[c] -= f * cb[c];
            row[j] *= rsqrt(piv);
        }
    }
    if (__syncthreads_or(bad)) { if (tid == 0) *status = 1; return; }
    // ---- results: factor rows to global, y_K to shared, then b_I -= L_IK y_K
    double *ysh = colbuf[0];
    if (tid == 2 * TS) {
#pragma unroll
        for (int c = 0; c < TS; ++c) ysh[c] = row[c];
        if (diagcta) {
#pragma unroll
            for (int c = 0; c < TS; ++c) ywork[K * TS + c] = row[c];
        }
    }
    if (diagcta) {
        if (tid < TS) {
            double2 *dst = reinterpret_cast<double2 *>(Ldiag + (size_t)K * TS * TS + tid * TS);
#pragma unroll
            for (int c = 0; c < TS; c += 2) dst[c / 2] = make_double2(c <= tid ? row[c] : 0.0, c + 1 <= tid ? row[c + 1] : 0.0);
        }
        return;
    }
    __syncthreads();
    if (tid >= TS && tid < 2 * TS) {
        double2 *dst = reinterpret_cast<double2 *>(tik + (tid - TS) * TS);
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < TS; c += 2) { dst[c / 2] = make_double2(row[c], row[c + 1]); s += row[c] * ysh[c] + row[c + 1] * ysh[c + 1]; }
        bwork[I * TS + (tid - TS)] -= s;
    }
}
// L_KK^-1 for every diagonal tile (needed by the backward substitution only): one CTA per tile,
// all tiles in parallel, off the critical path of the factorisation
__global__ void __launch_bounds__(256) k_diag_inverse(const double *__restrict__ Ldiag, double *__restrict__ Linv, const int *__restrict__ status)
{
    extern __shared__ double smem[];
    if (*status != 0) return;
    double *B0 = smem, *B1 = smem + TILE_SM;
    TileRegs<256> r;
    tile_ldg(r, Ldiag + (size_t)blockIdx.x * TS * TS);
    tile_sts(B0, r);
    __syncthreads();
    tile_tri_inverse(B0, B1);
    tile_stg<256>(Linv + (size_t)blockIdx.x * TS * TS, B1);
}

static void enqueue_factor(psba_ctx *c)
{
    const int npad = c->nt * TS;
    k_init_rhs<<<cdiv(npad, 256), 256, 0, c->stream>>>(c->N, npad, c->eab, c->chol_aux);
    for (int K = 0; K < c->nt; ++K) {
        const int ncrit = c->crit_ptr[K + 1] - c->crit_ptr[K];
        const int nncr = c->ncr_ptr[K + 1] - c->ncr_ptr[K];
        k_panel<<<ncrit + nncr, PANEL_NT, CHOL_SMEM, c->stream>>>(K, c->nt, ncrit, c->d_crit_rows + c->crit_ptr[K],
                                                                 c->d_ncr_I + c->ncr_ptr[K], c->d_ncr_J + c->ncr_ptr[K], c->tile_index,
                                                                 c->Stiles, c->Ldiag, c->chol_aux, c->chol_diag, c->d_status);
    }
    k_diag_inverse<<<c->nt, 256, 2 * TILE_SM * sizeof(double), c->stream>>>(c->Ldiag, c->Linv, c->d_status);
}

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
    c->st_launches += c->nt + 2;
    int st = 0;
    CUDA_CHECK(cudaMemcpyAsync(&st, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->factor_valid = (st == 0);
    c->S_valid = false;      // the factor overwrote the tile pool
    return st ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------
// backward substitution x_I = L_II^-T (y_I - sum_{J>I} L_JI^T x_J), one persistent CTA of 960 threads
// = 48 columns x 20 tile slots: every thread streams one tile column (48 independent coalesced loads in
// flight, four independent FMA chains), partial sums are combined in a fixed order.  L_II^-1 does not
// depend on x and is prefetched into shared memory while the tile column streams.
#define BW_SLOTS 20
__global__ void __launch_bounds__(TS * BW_SLOTS) k_backward(int N, int nt, const int *__restrict__ cptr, const int *__restrict__ crow,
                                                           const int *__restrict__ cslot, const double *__restrict__ Stiles,
                                                           const double *__restrict__ Linv, double *__restrict__ ywork, double *__restrict__ sol)
{
    __shared__ double part[BW_SLOTS][TS];
    __shared__ double acc[TS];
    __shared__ double invs[TS * TS];
    __shared__ double mv[4][TS];
    const int tid = threadIdx.x, col = tid % TS, slotid = tid / TS;
    // tile indices of a column do not depend on x: they are fetched one step ahead so that the only
    // dependent global accesses inside a step are the tile column itself and the x vectors
    int nbeg = cptr[nt - 1], nend = cptr[nt];
    int nslot = -1, nrow = 0;
    if (nbeg + slotid < nend) { nslot = cslot[nbeg + slotid]; nrow = crow[nbeg + slotid]; }
    for (int I = nt - 1; I >= 0; --I) {
        const int beg = nbeg, end = nend, myslot = nslot, myrow = nrow;
        if (I > 0) {
            nbeg = cptr[I - 1]; nend = beg;
            nslot = -1;
            if (nbeg + slotid < nend) { nslot = cslot[nbeg + slotid]; nrow = crow[nbeg + slotid]; }
        }
        for (int e = tid; e < TS * TS; e += TS * BW_SLOTS) invs[e] = __ldg(Linv + (size_t)I * TS * TS + e);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int t = beg + slotid; t < end; t += BW_SLOTS) {
            const bool first = t == beg + slotid;
            const double *L = Stiles + (size_t)(first ? myslot : cslot[t]) * TS * TS + col;
            const double *x = ywork + (first ? myrow : crow[t]) * TS;
            double lv[TS];
#pragma unroll
            for (int r = 0; r < TS; ++r) lv[r] = __ldg(L + r * TS);
#pragma unroll
            for (int r = 0; r < TS; r += 4) { s0 += lv[r] * x[r]; s1 += lv[r + 1] * x[r + 1]; s2 += lv[r + 2] * x[r + 2]; s3 += lv[r + 3] * x[r + 3]; }
        }
        part[slotid][col] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        if (tid < TS) {
            double a = 0.0;
#pragma unroll
            for (int p = 0; p < BW_SLOTS; ++p) a += part[p][tid];
            acc[tid] = ywork[I * TS + tid] - a;
        }
        __syncthreads();
        if (tid < 4 * TS) {          // x_I[c] = sum_{r >= c} Linv[r][c] acc[r], rows split over 4 partial sums
            const int cidx = tid % TS, q = tid / TS;
            double a = 0.0;
            for (int r = cidx + q; r < TS; r += 4) a += invs[r * TS + cidx] * acc[r];
            mv[q][cidx] = a;
        }
        __syncthreads();
        if (tid < TS) {
            const double a = (mv[0][tid] + mv[1][tid]) + (mv[2][tid] + mv[3][tid]);
            ywork[I * TS + tid] = a;
            const int gr = I * TS + tid;
            if (gr < N) sol[gr] = a;
        }
        __syncthreads();
    }
}

void psba_launch_solve(psba_ctx *c)
{
    PROF(c, KID_TRI_SOLVE) k_backward<<<1, TS * BW_SLOTS, 0, c->stream>>>(c->N, c->nt, c->d_coltile_ptr, c->d_coltile_row, c->d_coltile_slot,
                                                               c->Stiles, c->Linv, c->chol_diag, c->dp);
    c->st_launches += 1;
}
