// kernels_schur.cu -- damping + Schur complement build.
//
// Replaces kern_update_UV / kern_restore_UVdiag (CL_files/update_UV.cl, restore_UVdiag.cl: the
// damping term is a kernel argument, U and V are never modified), kern_compute_Vinv
// (compute_Vinv.cl:6-90), kern_compute_Yblks (compute_Yblks.cl:6-39: Y stays in registers),
// kern_compute_S (compute_S.cl:6-78) and kern_compute_ea (compute_ea.cl:6-37).
//
// S_kl = delta_kl (U_k + mu I) - sum_{i in common(k,l)} Y_ik W_il^T is built only for k >= l (the
// reference's solvers read only that part, SPD_inv.cl:43-57) from a list of (obs_k, obs_l) triples
// sorted by camera pair with ascending point index -- the order of comm3DIdx (misc.cpp:199-209).
// Partial sums per chunk of triples, then a fixed-order sum per pair block: no atomics.
#include "dev_math.cuh"

// packed symmetric storage: V = (v00,v01,v02,v11,v12,v22); Vinv = (i00,i10,i20,i11,i21,i22)
__global__ void k_vinv(int n, const double *__restrict__ V, double mu, double *__restrict__ Vinv, int *__restrict__ flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *v = V + (size_t)i * 6;
    const double a11 = v[0] + mu, a12 = v[1], a13 = v[2], a22 = v[3] + mu, a23 = v[4], a33 = v[5] + mu;
    double T = (a33 * a12 * a12 - 2 * a12 * a13 * a23 + a22 * a13 * a13 + a11 * a23 * a23 - a11 * a22 * a33);
    double *o = Vinv + (size_t)i * 6;
    if (fabs(T) < 1e-16) {
        // compute_Vinv.cl:31-73 : determinant by pivoted LU, adjugate of the (symmetric) block
        *flag = 1;
        double a[3][3] = {{a11, a12, a13}, {a12, a22, a23}, {a13, a23, a33}};
        int mx = 0;
        if (a[0][0] < a[1][0]) mx = 1;
        if (a[mx][0] < a[2][0]) mx = 2;
        if (mx != 0) for (int q = 0; q < 3; ++q) { double t = a[0][q]; a[0][q] = a[mx][q]; a[mx][q] = t; }
        a[1][0] = a[1][0] / a[0][0]; a[2][0] = a[2][0] / a[0][0];
        a[1][1] = a[1][1] - a[1][0] * a[0][1]; a[1][2] = a[1][2] - a[1][0] * a[0][2];
        a[2][1] = a[2][1] - a[2][0] * a[0][1]; a[2][2] = a[2][2] - a[2][0] * a[0][2];
        if (a[1][1] < a[2][1]) for (int q = 0; q < 3; ++q) { double t = a[1][q]; a[1][q] = a[2][q]; a[2][q] = t; }
        if (a[1][1] != 0.0) { a[2][1] = a[2][1] / a[1][1]; a[2][2] = a[2][2] - a[2][1] * a[1][2]; }
        T = a[0][0] * a[1][1] * a[2][2];
        o[0] = (a22 * a33 - a23 * a23) / T;
        o[1] = -(a12 * a33 - a23 * a13) / T;
        o[3] = (a11 * a33 - a13 * a13) / T;
        o[2] = (a12 * a23 - a22 * a13) / T;
        o[4] = -(a11 * a23 - a12 * a13) / T;
        o[5] = (a11 * a22 - a12 * a12) / T;
        return;
    }
    o[0] = -(-a23 * a23 + a22 * a33) / T;
    o[1] = -(a13 * a23 - a12 * a33) / T;
    o[3] = -(-a13 * a13 + a11 * a33) / T;
    o[2] = -(a12 * a23 - a13 * a22) / T;
    o[4] = -(a12 * a13 - a11 * a23) / T;
    o[5] = -(-a12 * a12 + a11 * a22) / T;
}

double psba_launch_vinv(psba_ctx *c, double mu)
{
    CUDA_CHECK(cudaMemsetAsync(c->d_status + 1, 0, sizeof(int), c->stream));
    if (c->n > 0) PROF(c, KID_VINV) k_vinv<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->n, c->V, mu, c->Vinv, c->d_status + 1);
    c->st_launches += 1;
    return 0.0;
}

// Pair pass.  A group of G lanes (G = 1..32, chosen on the host from the mean run length) owns one chunk
// of triples of ONE camera pair (k >= l); lanes stride the chunk.  Per triple: Y = W_a * Vinv_i row by
// row (compute_Yblks.cl:26-37), acc[r][c] += Y_r . W_b[c] (compute_S.cl:44-52); diagonal pairs also
// accumulate Y * gb_i (compute_ea.cl:27-33).  The group's 36 (+6) sums are combined by an xor butterfly
// (bit-identical in every lane, fixed order) and written as the chunk's partial.
template <bool DIAG, int G>
__device__ __forceinline__ void pair_accumulate(long long beg, long long end, int lane, const int *__restrict__ tri_oa,
                                                const int *__restrict__ tri_ob, const int *__restrict__ tri_pt,
                                                const double *__restrict__ W, const double *__restrict__ Vinv,
                                                const double *__restrict__ gb, double *acc)
{
    long long t = beg + lane;
    int a_n = 0, b_n = 0, i_n = 0;
    if (t < end) { a_n = __ldg(tri_oa + t); b_n = DIAG ? a_n : __ldg(tri_ob + t); i_n = __ldg(tri_pt + t); }
#pragma unroll 1
    for (; t < end; t += G) {
        const int a = a_n, b = b_n, i = i_n;
        if (t + G < end) {                              // indices of the next triple fly with this triple's blocks
            a_n = __ldg(tri_oa + t + G); b_n = DIAG ? a_n : __ldg(tri_ob + t + G); i_n = __ldg(tri_pt + t + G);
        }
        const double2 *vp = reinterpret_cast<const double2 *>(Vinv + (size_t)i * 6);
        const double2 v01 = __ldg(vp), v23 = __ldg(vp + 1), v45 = __ldg(vp + 2);
        const double i00 = v01.x, i10 = v01.y, i20 = v23.x, i11 = v23.y, i21 = v45.x, i22 = v45.y;
        double wb[18];
        const double2 *wbp = reinterpret_cast<const double2 *>(W + (size_t)b * 18);
#pragma unroll
        for (int q = 0; q < 9; ++q) { double2 w2 = __ldg(wbp + q); wb[2 * q] = w2.x; wb[2 * q + 1] = w2.y; }
        double g0 = 0, g1 = 0, g2 = 0;
        if (DIAG) { const double *gp = gb + (size_t)i * 3; g0 = __ldg(gp); g1 = __ldg(gp + 1); g2 = __ldg(gp + 2); }
        const double2 *wap = reinterpret_cast<const double2 *>(W + (size_t)a * 18);
#pragma unroll
        for (int rp = 0; rp < 3; ++rp) {              // two rows of W_a (6 doubles = 3 double2) at a time
            double wa[6];
            if (DIAG) {
#pragma unroll
                for (int q = 0; q < 6; ++q) wa[q] = wb[rp * 6 + q];
            } else {
#pragma unroll
                for (int q = 0; q < 3; ++q) { double2 w2 = __ldg(wap + rp * 3 + q); wa[2 * q] = w2.x; wa[2 * q + 1] = w2.y; }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = rp * 2 + h;
                const double w0 = wa[h * 3], w1 = wa[h * 3 + 1], w2 = wa[h * 3 + 2];
                const double y0 = w0 * i00 + w1 * i10 + w2 * i20;
                const double y1 = w0 * i10 + w1 * i11 + w2 * i21;
                const double y2 = w0 * i20 + w1 * i21 + w2 * i22;
#pragma unroll
                for (int cc = 0; cc < 6; ++cc)
                    acc[r * 6 + cc] += y0 * wb[cc * 3] + y1 * wb[cc * 3 + 1] + y2 * wb[cc * 3 + 2];
                if (DIAG) acc[36 + r] += y0 * g0 + y1 * g1 + y2 * g2;
            }
        }
    }
}

template <int G>
__global__ void __launch_bounds__(PAIR_CTA, 2) k_schur_pairs(int n_pchunk, const int *__restrict__ pchunk_pair,
                                                         const long long *__restrict__ pchunk_beg, const long long *__restrict__ pchunk_end,
                                                         const int *__restrict__ pair_k, const int *__restrict__ pair_l,
                                                         const int *__restrict__ tri_oa, const int *__restrict__ tri_ob,
                                                         const int *__restrict__ tri_pt, const double *__restrict__ W,
                                                         const double *__restrict__ Vinv, const double *__restrict__ gb,
                                                         double *__restrict__ part)
{
    const int lane = threadIdx.x % G;
    const int ch = blockIdx.x * (PAIR_CTA / G) + threadIdx.x / G;
    double acc[42];
#pragma unroll
    for (int q = 0; q < 42; ++q) acc[q] = 0.0;
    bool diag = false;
    if (ch < n_pchunk) {
        const int pr = pchunk_pair[ch];
        diag = pair_k[pr] == pair_l[pr];
        if (diag) pair_accumulate<true, G>(pchunk_beg[ch], pchunk_end[ch], lane, tri_oa, tri_ob, tri_pt, W, Vinv, gb, acc);
        else pair_accumulate<false, G>(pchunk_beg[ch], pchunk_end[ch], lane, tri_oa, tri_ob, tri_pt, W, Vinv, gb, acc);
    }
#pragma unroll
    for (int w = G / 2; w > 0; w >>= 1) {
#pragma unroll
        for (int q = 0; q < 42; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], w);
    }
    if (ch < n_pchunk) {
        double *out = part + (size_t)ch * 42;
        const int nv = diag ? 42 : 36;
#pragma unroll
        for (int q = 0; q < 42; ++q)
            if ((q % G) == lane && q < nv) out[q] = acc[q];
    }
}

template <int G>
static void launch_pairs(psba_ctx *c)
{
    const int per_cta = PAIR_CTA / G;
    k_schur_pairs<G><<<cdiv(c->n_pchunk, per_cta), PAIR_CTA, 0, c->stream>>>(c->n_pchunk, c->pchunk_pair, c->pchunk_beg, c->pchunk_end,
                                                                          c->pair_k, c->pair_l, c->tri_oa, c->tri_ob, c->tri_pt, c->W,
                                                                          c->Vinv, c->g + c->N, c->pair_part);
}

// position of entry (r,cc) of the camera block (k,l) inside the tile pool: the camera system is stored in
// the solver's camera ordering (cam2pos); a block that lands above the diagonal is stored transposed
__device__ __forceinline__ double *s_entry(double *Stiles, const int *__restrict__ tile_index, int nt, int pk, int pl, int r, int cc)
{
    if (pk < pl) { int t = pk; pk = pl; pl = t; t = r; r = cc; cc = t; }
    const int slot = tile_index[(pk / 8) * nt + pl / 8];
    return Stiles + (size_t)slot * TS * TS + ((pk % 8) * 6 + r) * TS + (pl % 8) * 6 + cc;
}

// per pair block: fixed-order sum of its chunk partials, S_kl = [k==l](U_k + mu I) - sum, written
// into the 48x48 tile pool; ea_k = ga_k - sum_e.  with_U=0 writes only the (negated) local sums
// (multi-GPU: all-reduce first, then k_add_U).
__global__ void k_S_finalize(int n_pair, const int *__restrict__ pair_k, const int *__restrict__ pair_l,
                             const int *__restrict__ pair_chunk_ptr, const double *__restrict__ part,
                             const double *__restrict__ U, const double *__restrict__ ga, double mu, int with_U,
                             const int *__restrict__ tile_index, const int *__restrict__ cam2pos, int nt,
                             double *__restrict__ Stiles, double *__restrict__ ea)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int pr = t / 42, v = t - pr * 42;
    if (pr >= n_pair) return;
    const int k = pair_k[pr], l = pair_l[pr];
    if (v >= 36 && k != l) return;
    double s = 0.0;
    for (int ch = pair_chunk_ptr[pr]; ch < pair_chunk_ptr[pr + 1]; ++ch) s += part[(size_t)ch * 42 + v];
    if (v < 36) {
        const int r = v / 6, cc = v - r * 6;
        double val = -s;
        if (k == l && with_U) { val = U[k * 36 + r * 6 + cc] - s; if (r == cc) val = (U[k * 36 + r * 6 + cc] + mu) - s; }
        *s_entry(Stiles, tile_index, nt, cam2pos[k], cam2pos[l], r, cc) = val;
    } else {
        const int r = v - 36;
        ea[k * 6 + r] = with_U ? ga[k * 6 + r] - s : -s;
    }
}

// multi-GPU second half: add U_k + mu I to the diagonal blocks and ga to ea (after the all-reduce)
__global__ void k_add_U(int m, const double *__restrict__ U, const double *__restrict__ ga, double mu,
                        const int *__restrict__ tile_index, const int *__restrict__ cam2pos, int nt,
                        double *__restrict__ Stiles, double *__restrict__ ea)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int k = t / 42, v = t - k * 42;
    if (k >= m) return;
    if (v < 36) {
        const int r = v / 6, cc = v - r * 6;
        double *p = s_entry(Stiles, tile_index, nt, cam2pos[k], cam2pos[k], r, cc);
        double u = U[k * 36 + v];
        if (r == cc) u += mu;
        *p = u + *p;
    } else ea[k * 6 + (v - 36)] += ga[k * 6 + (v - 36)];
}

// identity on the padding rows (block positions without a camera) so that the factorisation is well defined
__global__ void k_pad_diag(int nt, const int *__restrict__ pos2cam, const int *__restrict__ tile_index, double *__restrict__ Stiles)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nt * TS || pos2cam[r / 6] >= 0) return;
    int I = r / TS;
    Stiles[(size_t)tile_index[I * nt + I] * TS * TS + (r % TS) * TS + (r % TS)] = 1.0;
}

void psba_launch_schur(psba_ctx *c, double mu)
{
    psba_launch_vinv(c, mu);
    PROF(c, KID_MEMSET_S) CUDA_CHECK(cudaMemsetAsync(c->Stiles, 0, (size_t)c->n_tiles * TS * TS * sizeof(double), c->stream));
    if (c->n_pchunk > 0)
        PROF(c, KID_SCHUR_PAIRS) {
            switch (c->pair_G) {
            case 1: launch_pairs<1>(c); break;
            case 2: launch_pairs<2>(c); break;
            case 4: launch_pairs<4>(c); break;
            case 8: launch_pairs<8>(c); break;
            case 16: launch_pairs<16>(c); break;
            default: launch_pairs<32>(c); break;
            }
        }
    const int single = c->nranks == 1;
    PROF(c, KID_S_FINALIZE) k_S_finalize<<<cdiv((long long)c->n_pair * 42, 128), 128, 0, c->stream>>>(c->n_pair, c->pair_k, c->pair_l, c->pair_chunk_ptr,
                                                                             c->pair_part, c->U, c->g, mu, single, c->tile_index,
                                                                             c->cam2pos, c->nt, c->Stiles, c->eab);
    c->st_launches += 3;
    if (!single) {
        // the tile pool and ea are contiguous-by-construction only separately: two all-reduces
        psba_allreduce_sum(c, c->Stiles, (size_t)c->n_tiles * TS * TS);
        psba_allreduce_sum(c, c->eab, (size_t)c->N);
        k_add_U<<<cdiv(c->m * 42, 128), 128, 0, c->stream>>>(c->m, c->U, c->g, mu, c->tile_index, c->cam2pos, c->nt, c->Stiles, c->eab);
        c->st_launches += 1;
    }
    if (c->nt * TS > c->N) { k_pad_diag<<<cdiv(c->nt * TS, 256), 256, 0, c->stream>>>(c->nt, c->pos2cam, c->tile_index, c->Stiles); c->st_launches += 1; }
    c->S_valid = true; c->factor_valid = false;
}

// compat only: Y_ij = W_ij * Vinv_i materialised in the reference's layout (compute_Yblks.cl:6-39)
__global__ void k_Y_materialize(int o, const int *__restrict__ iidx, const double *__restrict__ W,
                                const double *__restrict__ Vinv, double *__restrict__ Y)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= o) return;
    const double *vi = Vinv + (size_t)iidx[k] * 6;
    const double i00 = vi[0], i10 = vi[1], i20 = vi[2], i11 = vi[3], i21 = vi[4], i22 = vi[5];
    for (int r = 0; r < 6; ++r) {
        const double *w = W + (size_t)k * 18 + r * 3;
        double *y = Y + (size_t)k * 18 + r * 3;
        y[0] = w[0] * i00 + w[1] * i10 + w[2] * i20;
        y[1] = w[0] * i10 + w[1] * i11 + w[2] * i21;
        y[2] = w[0] * i20 + w[1] * i21 + w[2] * i22;
    }
}

void psba_launch_Y_materialize(psba_ctx *c, double *Y)
{
    if (c->o > 0) k_Y_materialize<<<cdiv(c->o, 128), 128, 0, c->stream>>>(c->o, c->iidx, c->W, c->Vinv, Y);
    c->st_launches += 1;
}
