// kernels_schur.cu -- damping + Schur complement build.
//
// Replaces kern_update_UV / kern_restore_UVdiag (CL_files/update_UV.cl, restore_UVdiag.cl: the
// damping term is a kernel argument, U and V are never modified), kern_compute_Vinv
// (compute_Vinv.cl:6-90), kern_compute_Yblks (compute_Yblks.cl:6-39: Y never reaches HBM),
// kern_compute_S (compute_S.cl:6-78) and kern_compute_ea (compute_ea.cl:6-37).
//
// S_kl = delta_kl (U_k + mu I) - sum_{i in common(k,l)} Y_ik W_il^T is built only for k >= l (the
// reference's solvers read only that part, SPD_inv.cl:43-57) from a list of (obs_k, obs_l) triples
// sorted by camera pair with ascending point index -- the order of comm3DIdx (misc.cpp:199-209).
// Partial sums per chunk of triples, then a fixed-order sum per pair block: no atomics.
//
// Two kernels build the partial sums:
//   k_schur_segs  (default)   CTA = one SEGMENT of a camera row k (<= SEG_V consecutive visits of camera k, ascending
//                 point): phase 1 forms Y_ik = W_ik Vinv_i for every visit of the segment ONCE into shared memory and
//                 sums the diagonal block and ea_k on the way; phase 2 walks the off-diagonal triples (k, l, i) of the
//                 segment pair by pair: Y_ik comes from shared memory, only W_il is gathered.
//   k_schur_pairs (PSBA_PAIR_MODE=0)   the round-1 kernel: every triple gathers W_ik, W_il and Vinv_i.
// What bounds both is the number of L1 line look-ups (a divergent load costs one look-up per lane whatever its
// width): every block is therefore fetched with 256-bit loads (LDG.E.256: five look-ups per 144-byte block
// instead of nine) and the segment kernel needs one block per off-diagonal triple instead of two blocks + Vinv.
#include "dev_math.cuh"

// ---- 256-bit gathers ---------------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const double *p, double &a, double &b, double &c, double &d)
{
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// 144-byte block at a 16-byte aligned address (W of one observation: every second block starts in the middle of a
// 32-byte sector): four 32-byte loads + one 16-byte load, the short one in front when the block starts mid-sector
__device__ __forceinline__ void load_blk18(const double *__restrict__ p, double w[18])
{
    const bool odd = (reinterpret_cast<unsigned long long>(p) & 16ull) != 0;
    const double *q = odd ? p + 2 : p;
    double t[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) ldg256(q + 4 * j, t[4 * j], t[4 * j + 1], t[4 * j + 2], t[4 * j + 3]);
    const double2 e = __ldg(reinterpret_cast<const double2 *>(odd ? p : p + 16));
#pragma unroll
    for (int k = 0; k < 18; ++k) {
        const double lo = k < 2 ? (k == 0 ? e.x : e.y) : t[k >= 2 ? k - 2 : 0];
        const double hi = k < 16 ? t[k < 16 ? k : 0] : (k == 16 ? e.x : e.y);
        w[k] = odd ? lo : hi;
    }
}
// 48-byte record (Vinv of one point) at a 16-byte aligned address
__device__ __forceinline__ void load_sym6(const double *__restrict__ p, double v[6])
{
    const bool odd = (reinterpret_cast<unsigned long long>(p) & 16ull) != 0;
    double t0, t1, t2, t3;
    ldg256(odd ? p + 2 : p, t0, t1, t2, t3);
    const double2 e = __ldg(reinterpret_cast<const double2 *>(odd ? p : p + 4));
    v[0] = odd ? e.x : t0; v[1] = odd ? e.y : t1; v[2] = odd ? t0 : t2; v[3] = odd ? t1 : t3; v[4] = odd ? t2 : e.x; v[5] = odd ? t3 : e.y;
}
// 24-byte record (gb of one point) at an 8-byte aligned address: 16 + 8 or 8 + 16
__device__ __forceinline__ void load_vec3(const double *__restrict__ p, double &g0, double &g1, double &g2)
{
    const bool odd = (reinterpret_cast<unsigned long long>(p) & 8ull) != 0;
    const double2 e = __ldg(reinterpret_cast<const double2 *>(odd ? p + 1 : p));
    const double s = __ldg(odd ? p : p + 2);
    g0 = odd ? s : e.x; g1 = odd ? e.x : e.y; g2 = odd ? e.y : s;
}

// packed symmetric storage: V = (v00,v01,v02,v11,v12,v22); Vinv = (i00,i10,i20,i11,i21,i22)
__global__ void k_vinv(int n, const double *__restrict__ V, const double *__restrict__ mu_p, double *__restrict__ Vinv, int *__restrict__ flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mu = __ldg(mu_p);                               // damping term of this solve (device-resident: psba_set_scalars)
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
    psba_set_scalars(c, mu, 0.0, 0.0);
    CUDA_CHECK(cudaMemsetAsync(c->d_status + 1, 0, sizeof(int), c->stream));
    if (c->n > 0) PROF(c, KID_VINV) k_vinv<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->n, c->V, c->d_mu, c->Vinv, c->d_status + 1);
    c->st_launches += 1;
    LAUNCH_CHECK();
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
        double vi[6];
        load_sym6(Vinv + (size_t)i * 6, vi);
        const double i00 = vi[0], i10 = vi[1], i20 = vi[2], i11 = vi[3], i21 = vi[4], i22 = vi[5];
        double wb[18], wfull[18];
        load_blk18(W + (size_t)b * 18, wb);
        if (!DIAG) load_blk18(W + (size_t)a * 18, wfull);
        double g0 = 0, g1 = 0, g2 = 0;
        if (DIAG) load_vec3(gb + (size_t)i * 3, g0, g1, g2);
#pragma unroll
        for (int rp = 0; rp < 3; ++rp) {              // two rows of W_a at a time
            double wa[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) wa[q] = DIAG ? wb[rp * 6 + q] : wfull[rp * 6 + q];
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


// ---- Segment kernel -----------------------------------------------------------------------------------
// sum of 32 values per lane over the warp by recursive halving: lane L ends up with the total of value L (fixed tree,
// 31 exchanged values instead of 32 x 5 for a butterfly of every value)
__device__ __forceinline__ double warp_reduce_scatter32(double v[32])
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int j = 0; j < h; ++j) {
            const double send = up ? v[j] : v[j + h], keep = up ? v[j + h] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return v[0];
}

struct seg_desc { int row, v0, v1, diag_chunk, sched0, sched1; };

// CTA = one segment: visits [v0, v1) of camera `row` in camera-major order (ascending point), at most SEG_V of them.
//  phase 1  one thread per visit (strided): W_ik, Vinv_i, gb_i -> Y_ik = W_ik Vinv_i (compute_Yblks.cl:26-37) into the
//           shared tile; the lower triangle of Y_ik W_ik^T and Y_ik gb_i (compute_S.cl:44-52 on the diagonal block,
//           compute_ea.cl:27-33) are summed per thread in ascending visit order, over the CTA by a fixed tree, and
//           leave as the segment's partial of the diagonal pair;
//  phase 2  the off-diagonal chunks of the segment -- the triples of one pair (k, l), l < k, whose visit lies in the
//           segment: a contiguous piece of the pair's run in the pair-sorted triple list, ascending point -- are dealt
//           to groups of G lanes, largest first, a warp's worth at a time (shared counter; which group computes a
//           chunk does not change its sum).  Per triple: Y_ik from shared memory, W_il by 256-bit gathers, 108 FMA.
// Partial per chunk; k_S_finalize sums the partials of a pair in segment order.
template <int G, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_schur_segs(const seg_desc *__restrict__ segs, const int *__restrict__ cam_obs,
                                                          const int *__restrict__ cam_pt, const int *__restrict__ sched,
                                                          const int *__restrict__ ch_beg, const int *__restrict__ ch_end,
                                                          const unsigned short *__restrict__ tri_vr, const int *__restrict__ tri_ob,
                                                          const double *__restrict__ W, const double *__restrict__ Vinv,
                                                          const double *__restrict__ gb, double *__restrict__ part)
{
    extern __shared__ __align__(16) double Ysm[];                 // [visits of the segment][18]
    __shared__ double red[NT / 32][32];
    __shared__ int4 sdesc[SEG_MAXD];
    __shared__ int next;
    const seg_desc sd = segs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int nv = sd.v1 - sd.v0, nch = sd.sched1 - sd.sched0;
    if (tid == 0) next = 0;
    for (int j = tid; j < min(nch, SEG_MAXD); j += NT) {      // chunk descriptors of phase 2 fly under phase 1
        const int c = __ldg(sched + sd.sched0 + j);
        sdesc[j] = make_int4(__ldg(ch_beg + c), __ldg(ch_end + c), c, 0);
    }
    // ---- phase 1
    double acc[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) acc[q] = 0.0;
    int q_n = 0, i_n = 0;
    if (tid < nv) { q_n = __ldg(cam_obs + sd.v0 + tid); i_n = __ldg(cam_pt + sd.v0 + tid); }
#pragma unroll 1
    for (int r = tid; r < nv; r += NT) {
        double w[18], vi[6], g0, g1, g2;
        const int q = q_n, i = i_n;
        if (r + NT < nv) { q_n = __ldg(cam_obs + sd.v0 + r + NT); i_n = __ldg(cam_pt + sd.v0 + r + NT); }
        load_blk18(W + (size_t)q * 18, w);
        load_sym6(Vinv + (size_t)i * 6, vi);
        load_vec3(gb + (size_t)i * 3, g0, g1, g2);
        const double i00 = vi[0], i10 = vi[1], i20 = vi[2], i11 = vi[3], i21 = vi[4], i22 = vi[5];
        double2 *yd = reinterpret_cast<double2 *>(Ysm + (size_t)r * 18);
#pragma unroll
        for (int rp = 0; rp < 3; ++rp) {                         // two rows of Y at a time: three 16-byte stores
            double y[6];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double w0 = w[(rp * 2 + h) * 3], w1 = w[(rp * 2 + h) * 3 + 1], w2 = w[(rp * 2 + h) * 3 + 2];
                y[h * 3] = w0 * i00 + w1 * i10 + w2 * i20;
                y[h * 3 + 1] = w0 * i10 + w1 * i11 + w2 * i21;
                y[h * 3 + 2] = w0 * i20 + w1 * i21 + w2 * i22;
            }
            yd[rp * 3] = make_double2(y[0], y[1]); yd[rp * 3 + 1] = make_double2(y[2], y[3]); yd[rp * 3 + 2] = make_double2(y[4], y[5]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rr = rp * 2 + h;
#pragma unroll
                for (int cc = 0; cc <= rr; ++cc)
                    acc[rr * (rr + 1) / 2 + cc] += y[h * 3] * w[cc * 3] + y[h * 3 + 1] * w[cc * 3 + 1] + y[h * 3 + 2] * w[cc * 3 + 2];
                acc[21 + rr] += y[h * 3] * g0 + y[h * 3 + 1] * g1 + y[h * 3 + 2] * g2;
            }
        }
    }
    {
        const double tot = warp_reduce_scatter32(acc);           // lane L: value L of this warp
        red[wrp][lane] = tot;
    }
    __syncthreads();                                              // Y tile complete, warp sums published, `next` and sdesc visible
    if (tid < 27) {
        double sum = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < NT / 32; ++w8) sum += red[w8][tid];
        double *out = part + (size_t)sd.diag_chunk * 42;
        if (tid < 21) {
            int rr = 0, base = 0;
            while (base + rr + 1 <= tid) { base += rr + 1; ++rr; }   // tid = rr (rr + 1) / 2 + cc
            const int cc = tid - base;
            out[rr * 6 + cc] = sum;
            out[cc * 6 + rr] = sum;                              // the block is symmetric: W Vinv W^T
        } else out[36 + (tid - 21)] = sum;
    }
    // ---- phase 2
    constexpr int GPW = 32 / G;
    const int gl = lane % G;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&next, GPW);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= nch) break;
        const int j = base + lane / G;
        int beg = 0, end = 0, c = -1;
        if (j < nch) {
            if (j < SEG_MAXD) { const int4 d = sdesc[j]; beg = d.x; end = d.y; c = d.z; }
            else { c = __ldg(sched + sd.sched0 + j); beg = __ldg(ch_beg + c); end = __ldg(ch_end + c); }
        }
        double a36[36];
#pragma unroll
        for (int q = 0; q < 36; ++q) a36[q] = 0.0;
        int t = beg + gl;
        int r_n = 0, b_n = 0;
        if (t < end) { r_n = __ldg(tri_vr + t); b_n = __ldg(tri_ob + t); }
#pragma unroll 1
        for (; t < end; t += G) {
            const int r = r_n, b = b_n;
            if (t + G < end) { r_n = __ldg(tri_vr + t + G); b_n = __ldg(tri_ob + t + G); }
            double wb[18];
            load_blk18(W + (size_t)b * 18, wb);
            const double2 *yp = reinterpret_cast<const double2 *>(Ysm + (size_t)r * 18);
#pragma unroll
            for (int rp = 0; rp < 3; ++rp) {                     // two rows of Y_ik (three double2) at a time
                const double2 p0 = yp[rp * 3], p1 = yp[rp * 3 + 1], p2 = yp[rp * 3 + 2];
                const double ya[6] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y};
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc)
                        a36[(rp * 2 + h) * 6 + cc] += ya[h * 3] * wb[cc * 3] + ya[h * 3 + 1] * wb[cc * 3 + 1] + ya[h * 3 + 2] * wb[cc * 3 + 2];
            }
        }
#pragma unroll
        for (int w2 = G / 2; w2 > 0; w2 >>= 1) {
#pragma unroll
            for (int q = 0; q < 36; ++q) a36[q] += __shfl_xor_sync(0xffffffffu, a36[q], w2);
        }
        if (c >= 0) {
            double *out = part + (size_t)c * 42;
#pragma unroll
            for (int q = 0; q < 36; ++q)
                if ((q % G) == gl) out[q] = a36[q];
        }
    }
}

// Launch shape: 384 threads, ONE CTA per SM (the kernel needs 162 registers), Y tile of up to 1 280 visits (180 KB).  Measured
// on the headline workload (15 M triples; profiles/pair_pass_r02.md): 256 x 2 at 128 registers (spills) 1.49 ms, 192 x 2
// 1.30, 512 x 1 at 128 registers 1.11, 384 x 1 0.83; lanes per chunk G = 2 / 4 / 8 / 16 / 32: 1.28 / 0.95 / 0.83 / 1.00 / 1.34;
// segments of 640 instead of 1 280 visits 0.88; the W block of the next triple prefetched into registers (224 registers,
// 256 x 1) 0.84 -- no gain: the kernel is bound by the gather rate of the memory system, not by a lane's latency chain.
#define SEG_NT 384
template <int G>
static void launch_segs(psba_ctx *c)
{
    const int dyn = c->seg_v * 144;
    psba_set_smem((const void *)k_schur_segs<G, SEG_NT, 1>, dyn);
    k_schur_segs<G, SEG_NT, 1><<<c->n_seg, SEG_NT, dyn, c->stream>>>((const seg_desc *)c->seg_desc, c->cam_obs, c->cam_pt, c->sched_chunk, c->sch_beg,
                                                                   c->sch_end, c->tri_vr, c->tri_ob, c->W, c->Vinv, c->g + c->N, c->pair_part);
}


// ---- Ring kernel --------------------------------------------------------------------------------------
// Same segments, same chunks and the same partial slots as the segment kernel; what changes is how phase 2 gets its W_il
// blocks.  The off-diagonal triples of the segment are laid out as ROWS of 32 lane slots (structure.cu: k_ring_plan /
// k_ring_rows), the rows of a warp contiguous.  A warp copies the 32 blocks of a row COOPERATIVELY (nine consecutive lanes
// read the 144 bytes of one block: whole lines instead of one 32-byte sector per lane -- the segment kernel is bound by
// the wavefronts of its divergent loads in the L1 data pipe, profiles/ncu_full_r02.md) with cp.async into its own ring of
// `STAGES` row buffers, `STAGES` rows ahead of the row it multiplies; no CTA-wide barrier after phase 1.  Sums: the G
// lanes of a chunk by recursive halving (36 -> 18 -> 9 values) and a butterfly of the last nine: fixed order.
// the warp copies the blocks W[b] of its 32 lanes (b < 0: no block) into stage[lane * 18 ...]: nine consecutive lanes per
// block.  Measured and not kept: three whole blocks per instruction (27 lanes, no block straddles two instructions) 0.728
// against 0.731 ms; cp.async.ca (through L1) instead of .cg 0.99 ms
__device__ __forceinline__ void ring_issue(double *stage, const double *__restrict__ W, int b, int lane)
{
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int p = q * 32 + lane, sl = p / 9, part = p - sl * 9;
        const int bs = __shfl_sync(0xffffffffu, b, sl);
        if (bs >= 0) cp_async16(stage + p * 2, W + (size_t)bs * 18 + part * 2);
    }
}
// record of one row: 32 lane slots {observation of camera l, rank of the visit} + {task word, schedule position} + pad = 272 bytes,
// copied as seventeen 16-byte pieces that bypass L1 (an L1-allocating 8-byte copy per lane fetched 2.25 sectors per sector asked for)
#define RING_REC 34
// measured on the headline workload: records as 16-byte pieces past L1 0.731 -> 0.699 ms; not kept: the index loads by ld.global.cg
// too (0.706), gb by three 8-byte loads per lane past L1 instead of 8-byte copies through L1 (0.90 ms)
__device__ __forceinline__ void ring_issue_rec(int2 *slot, const int2 *__restrict__ rows, int row, int lane)
{
    if (lane < 17) cp_async16(slot + 2 * lane, rows + (size_t)row * RING_REC + 2 * lane);
}

// sum of the 36 block entries over the G = 2^LG lanes of every chunk of a task, partials out: recursive halving 36 -> 18 -> 9
// (the lane with the group bit set keeps the upper half), then a butterfly of the last nine; lanes whose low bits are zero write
template <int LG>
__device__ __forceinline__ void ring_task_reduce(double (&a36)[36], int lane, int nch, const int *__restrict__ sched_task, double *__restrict__ part)
{
    constexpr int G = 1 << LG;
    const int gl = lane & (G - 1);
    int start = 0;
    if (G >= 2) {
        constexpr int h = G >> 1;
        const bool up = (gl & h) != 0;
        if (up) start += 18;
#pragma unroll
        for (int j = 0; j < 18; ++j) {
            const double send = up ? a36[j] : a36[j + 18], keep = up ? a36[j + 18] : a36[j];
            a36[j] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    if (G >= 4) {
        constexpr int h = G >> 2;
        const bool up = (gl & h) != 0;
        if (up) start += 9;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const double send = up ? a36[j] : a36[j + 9], keep = up ? a36[j + 9] : a36[j];
            a36[j] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
#pragma unroll
    for (int h = G >> 3; h >= 1; h >>= 1) {
#pragma unroll
        for (int j = 0; j < 9; ++j) a36[j] += __shfl_xor_sync(0xffffffffu, a36[j], h);
    }
    constexpr int cnt = G >= 4 ? 9 : (G == 2 ? 18 : 36);
    const int q = lane >> LG;
    const bool writer = q < nch && (G < 8 || (gl & ((G >> 2) - 1)) == 0);
    if (writer) {
        const int c = __ldg(sched_task + q);
        double *out = part + (size_t)c * 42 + start;
#pragma unroll
        for (int j = 0; j < cnt; ++j) out[j] = a36[j];
    }
}

template <int NT, int STAGES, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_schur_ring(const seg_desc *__restrict__ segs, const int *__restrict__ cam_obs,
                                                          const int *__restrict__ cam_pt, const int *__restrict__ wrow_ptr,
                                                          const int2 *__restrict__ rows,
                                                          const int *__restrict__ sched, const double *__restrict__ W,
                                                          const double *__restrict__ Vinv, const double *__restrict__ gb,
                                                          double *__restrict__ part, int seg_v, long long *__restrict__ dbg)
{
    constexpr int NW = NT / 32, RS = 2 * STAGES + 1;
    long long stamp[6];
    stamp[0] = clock64(); stamp[1] = stamp[0];
    extern __shared__ __align__(16) double sm[];   // [seg_v][18] Y tile | [NW][STAGES][32][18] copy rings | [NW][RS][34] row records
    __shared__ double red[NW][32];
    const seg_desc sd = segs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int nv = sd.v1 - sd.v0;
    double *Ysm = sm;
    double *ring = sm + (size_t)seg_v * 18 + (size_t)wrp * STAGES * 576;
    int2 *recs = reinterpret_cast<int2 *>(sm + (size_t)seg_v * 18 + (size_t)NW * STAGES * 576) + (size_t)wrp * RS * RING_REC;
    const int r0 = __ldg(wrow_ptr + (size_t)blockIdx.x * NW + wrp), r1 = __ldg(wrow_ptr + (size_t)blockIdx.x * NW + wrp + 1);
    // records of the first 2 STAGES rows of this warp: they land under phase 1
#pragma unroll
    for (int s = 0; s < 2 * STAGES; ++s)
        if (r0 + s < r1) ring_issue_rec(recs + ((r0 + s) % RS) * RING_REC, rows, r0 + s, lane);
    cp_async_commit();
    // ---- phase 1: Y tile, diagonal block and ea of the segment.  A warp owns the visit rows wrp, wrp + NW, ... (32
    // consecutive visits each).  W_ik goes straight to its place in the Y tile, Vinv_i and gb_i into the (still idle) copy
    // ring of the warp, all by cooperative cp.async; then lane = visit: Y_ik = W_ik Vinv_i overwrites W_ik in place
    // (compute_Yblks.cl:26-37), the lower triangle of Y_ik W_ik^T and Y_ik gb_i are summed per lane in ascending visit order.
    double acc[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) acc[q] = 0.0;
    constexpr int RB = STAGES * 2;                                  // visit rows per batch: 2 304 B of Vinv | gb each
    const int nvr = (nv + 31) >> 5;
#pragma unroll 1
    for (int vb = wrp; vb < nvr; vb += NW * RB) {
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int r = (vb + j * NW) * 32 + lane;
            int q = -1, i = 0;
            if (r < nv) { q = __ldg(cam_obs + sd.v0 + r); i = __ldg(cam_pt + sd.v0 + r); }
            double *dstW = Ysm + (size_t)(vb + j * NW) * 32 * 18, *stg = ring + j * 288;
            ring_issue(dstW, W, q, lane);
#pragma unroll
            for (int qq = 0; qq < 3; ++qq) {
                const int p = qq * 32 + lane, sl = p / 3, part = p - sl * 3;
                const int qs = __shfl_sync(0xffffffffu, q, sl), is = __shfl_sync(0xffffffffu, i, sl);
                if (qs >= 0) {
                    cp_async16(stg + p * 2, Vinv + (size_t)is * 6 + part * 2);
                    cp_async8(stg + 192 + p, gb + (size_t)is * 3 + part);
                }
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        stamp[1] = clock64();
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int r = (vb + j * NW) * 32 + lane;
            if (r < nv) {
                double w[18];
                double2 *yd = reinterpret_cast<double2 *>(Ysm + (size_t)r * 18);
#pragma unroll
                for (int k = 0; k < 9; ++k) { const double2 v = yd[k]; w[2 * k] = v.x; w[2 * k + 1] = v.y; }
                const double *stg = ring + j * 288;
                const double2 *vp = reinterpret_cast<const double2 *>(stg + lane * 6);
                const double2 v0 = vp[0], v1 = vp[1], v2 = vp[2];
                const double i00 = v0.x, i10 = v0.y, i20 = v1.x, i11 = v1.y, i21 = v2.x, i22 = v2.y;
                const double g0 = stg[192 + lane * 3], g1 = stg[192 + lane * 3 + 1], g2 = stg[192 + lane * 3 + 2];
#pragma unroll
                for (int rp = 0; rp < 3; ++rp) {                     // two rows of Y at a time: three 16-byte stores
                    double y[6];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double w0 = w[(rp * 2 + h) * 3], w1 = w[(rp * 2 + h) * 3 + 1], w2 = w[(rp * 2 + h) * 3 + 2];
                        y[h * 3] = w0 * i00 + w1 * i10 + w2 * i20;
                        y[h * 3 + 1] = w0 * i10 + w1 * i11 + w2 * i21;
                        y[h * 3 + 2] = w0 * i20 + w1 * i21 + w2 * i22;
                    }
                    yd[rp * 3] = make_double2(y[0], y[1]); yd[rp * 3 + 1] = make_double2(y[2], y[3]); yd[rp * 3 + 2] = make_double2(y[4], y[5]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int rr = rp * 2 + h;
#pragma unroll
                        for (int cc = 0; cc <= rr; ++cc) {
                            double &a = acc[rr * (rr + 1) / 2 + cc];
                            a = fma(y[h * 3], w[cc * 3], a); a = fma(y[h * 3 + 1], w[cc * 3 + 1], a); a = fma(y[h * 3 + 2], w[cc * 3 + 2], a);
                        }
                        double &e = acc[21 + rr];
                        e = fma(y[h * 3], g0, e); e = fma(y[h * 3 + 1], g1, e); e = fma(y[h * 3 + 2], g2, e);
                    }
                }
            }
        }
        __syncwarp();                                               // the staging area is reused by the next batch / by phase 2
    }
    {
        const double tot = warp_reduce_scatter32(acc);
        red[wrp][lane] = tot;
    }
    stamp[2] = clock64();
    // ---- the first rows of phase 2 start their way before the barrier: group j carries the blocks of row j + STAGES and the
    // record of row j + 2 STAGES
    cp_async_wait<0>();
    __syncwarp();
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
        int b = -1;
        if (r0 + s < r1) b = recs[((r0 + s) % RS) * RING_REC + lane].x;
        ring_issue(ring + s * 576, W, b, lane);
        cp_async_commit();
    }
    __syncthreads();                                                // Y tile complete, warp sums published
    stamp[3] = clock64();
    if (tid < 27) {
        double sum = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < NW; ++w8) sum += red[w8][tid];
        double *out = part + (size_t)sd.diag_chunk * 42;
        if (tid < 21) {
            int rr = 0, base = 0;
            while (base + rr + 1 <= tid) { base += rr + 1; ++rr; }
            const int cc = tid - base;
            out[rr * 6 + cc] = sum;
            out[cc * 6 + rr] = sum;
        } else out[36 + (tid - 21)] = sum;
    }
    // ---- phase 2: the rows of this warp
    double a36[36];
#pragma unroll
    for (int q = 0; q < 36; ++q) a36[q] = 0.0;
    int st = 0, sl_cur = r0 % RS, sl_far = (r0 + STAGES) % RS, sl_new = (r0 + 2 * STAGES) % RS;
#pragma unroll 1
    for (int row = r0; row < r1; ++row) {
        cp_async_wait<STAGES - 1>();
        __syncwarp();
        const int2 cur = recs[sl_cur * RING_REC + lane], ci = recs[sl_cur * RING_REC + 32];
        const int b_far = row + STAGES < r1 ? recs[sl_far * RING_REC + lane].x : -1;
        double *stage = ring + st * 576;
        if (cur.x >= 0) {
            double wb[18];
            const double2 *wp = reinterpret_cast<const double2 *>(stage + lane * 18);
#pragma unroll
            for (int j = 0; j < 9; ++j) { const double2 v = wp[j]; wb[2 * j] = v.x; wb[2 * j + 1] = v.y; }
            const double2 *yp = reinterpret_cast<const double2 *>(Ysm + (size_t)cur.y * 18);
#pragma unroll
            for (int rp = 0; rp < 3; ++rp) {
                const double2 p0 = yp[rp * 3], p1 = yp[rp * 3 + 1], p2 = yp[rp * 3 + 2];
                const double ya[6] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y};
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc) {
                        double &a = a36[(rp * 2 + h) * 6 + cc];
                        a = fma(ya[h * 3], wb[cc * 3], a); a = fma(ya[h * 3 + 1], wb[cc * 3 + 1], a); a = fma(ya[h * 3 + 2], wb[cc * 3 + 2], a);
                    }
            }
        }
        __syncwarp();                                               // every lane has read its block: the stage can be refilled
        ring_issue(stage, W, b_far, lane);
        if (row + 2 * STAGES < r1) ring_issue_rec(recs + sl_new * RING_REC, rows, row + 2 * STAGES, lane);
        cp_async_commit();
        st = st + 1 == STAGES ? 0 : st + 1;
        sl_cur = sl_cur + 1 == RS ? 0 : sl_cur + 1; sl_far = sl_far + 1 == RS ? 0 : sl_far + 1; sl_new = sl_new + 1 == RS ? 0 : sl_new + 1;
        if (ci.x & 256) {                                           // last row of a task: sum over the G lanes of every chunk, partials out
            const int nch = ci.x >> 16;
            switch (ci.x & 255) {                                   // one straight-line reduction per group size (no register shuffling at merges)
            case 0: ring_task_reduce<0>(a36, lane, nch, sched + ci.y, part); break;
            case 1: ring_task_reduce<1>(a36, lane, nch, sched + ci.y, part); break;
            case 2: ring_task_reduce<2>(a36, lane, nch, sched + ci.y, part); break;
            case 3: ring_task_reduce<3>(a36, lane, nch, sched + ci.y, part); break;
            case 4: ring_task_reduce<4>(a36, lane, nch, sched + ci.y, part); break;
            default: ring_task_reduce<5>(a36, lane, nch, sched + ci.y, part); break;
            }
#pragma unroll
            for (int j = 0; j < 36; ++j) a36[j] = 0.0;
        }
    }
    cp_async_wait<0>();
    if (dbg && lane == 0) {
        stamp[4] = clock64();
        long long *d = dbg + ((size_t)blockIdx.x * NW + wrp) * 8;
#pragma unroll
        for (int k = 0; k < 5; ++k) d[k] = stamp[k];
        d[5] = r1 - r0; d[6] = nv;
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        d[7] = smid;
    }
}

template <int NT, int STAGES, int MINB>
static void launch_ring_shape(psba_ctx *c)
{
    const int dyn = (int)psba_ring_smem(c->ring_cfg, c->seg_v);
    psba_set_smem((const void *)k_schur_ring<NT, STAGES, MINB>, dyn);
    static int dbg_runs = getenv("PSBA_RING_DEBUG") ? 3 : 0;       // third launch: clock stamps of every warp (phase times)
    long long *dbg = nullptr;
    const size_t nd = (size_t)c->n_seg * (NT / 32) * 8;
    if (dbg_runs > 0 && --dbg_runs == 0) { dbg = (long long *)psba_dev_alloc(c, nd * 8, true); }
    k_schur_ring<NT, STAGES, MINB><<<c->n_seg, NT, dyn, c->stream>>>((const seg_desc *)c->seg_desc, c->cam_obs, c->cam_pt, c->ring_wrow_ptr, c->ring_rows,
                                                                    c->sched_chunk, c->W, c->Vinv, c->g + c->N, c->pair_part, c->seg_v, dbg);
    if (dbg) {
        std::vector<long long> h(nd);
        CUDA_CHECK(cudaMemcpyAsync(h.data(), dbg, nd * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        const int NW = NT / 32;
        double s_p1wait = 0, s_p1comp = 0, s_bar = 0, s_p2 = 0, s_cta = 0, s_rows = 0, s_maxrows = 0, s_p2max = 0, s_p2min = 0;
        for (int b = 0; b < c->n_seg; ++b) {
            long long t0 = h[(size_t)b * NW * 8], tend = 0, p2max = 0, p2min = 1ll << 60; int maxr = 0;
            for (int w = 0; w < NW; ++w) {
                const long long *d = &h[((size_t)b * NW + w) * 8];
                t0 = std::min(t0, d[0]); tend = std::max(tend, d[4]);
                s_p1wait += (double)(d[1] - d[0]) / NW; s_p1comp += (double)(d[2] - d[1]) / NW; s_bar += (double)(d[3] - d[2]) / NW; s_p2 += (double)(d[4] - d[3]) / NW;
                s_rows += (double)d[5] / NW; maxr = std::max(maxr, (int)d[5]);
                p2max = std::max(p2max, d[4] - d[3]); p2min = std::min(p2min, d[4] - d[3]);
            }
            s_cta += (double)(tend - t0); s_maxrows += maxr; s_p2max += (double)p2max; s_p2min += (double)p2min;
        }
        const double n = c->n_seg;
        fprintf(stderr, "ring debug: %d CTAs x %d warps, per CTA (cycles): life %.0f | start->phase-1 data %.0f, phase-1 compute %.0f, reduce+prologue+barrier %.0f, phase 2 mean %.0f (min %.0f max %.0f) | rows/warp mean %.1f max %.1f\n",
                c->n_seg, NW, s_cta / n, s_p1wait / n, s_p1comp / n, s_bar / n, s_p2 / n, s_p2min / n, s_p2max / n, s_rows / n, s_maxrows / n);
        psba_dev_free(c, dbg);
    }
}
static void launch_ring(psba_ctx *c)
{
    switch (c->ring_cfg) {
    case 0: launch_ring_shape<384, 2, 1>(c); break;
    case 1: launch_ring_shape<256, 2, 1>(c); break;
    case 2: launch_ring_shape<256, 3, 1>(c); break;
    case 3: launch_ring_shape<128, 2, 2>(c); break;
    case 4: launch_ring_shape<128, 3, 2>(c); break;
    default: launch_ring_shape<192, 2, 1>(c); break;
    }
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
                             const double *__restrict__ U, const double *__restrict__ ga, const double *__restrict__ mu_p, int with_U,
                             const int *__restrict__ tile_index, const int *__restrict__ cam2pos, int nt,
                             double *__restrict__ Stiles, double *__restrict__ ea)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int pr = t / 42, v = t - pr * 42;
    if (pr >= n_pair) return;
    const double mu = __ldg(mu_p);
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
__global__ void k_add_U(int m, const double *__restrict__ U, const double *__restrict__ ga, const double *__restrict__ mu_p,
                        const int *__restrict__ tile_index, const int *__restrict__ cam2pos, int nt,
                        double *__restrict__ Stiles, const double *__restrict__ ea_red, double *__restrict__ ea)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int k = t / 42, v = t - k * 42;
    if (k >= m) return;
    const double mu = __ldg(mu_p);
    if (v < 36) {
        const int r = v / 6, cc = v - r * 6;
        double *p = s_entry(Stiles, tile_index, nt, cam2pos[k], cam2pos[k], r, cc);
        double u = U[k * 36 + v];
        if (r == cc) u += mu;
        *p = u + *p;
    } else ea[k * 6 + (v - 36)] = ea_red[k * 6 + (v - 36)] + ga[k * 6 + (v - 36)];
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
    const int single = c->nranks == 1;
    // N > 1 GPUs: the local sums of ea go right behind the S tiles of the pool so that ONE all-reduce moves both
    double *ea_red = c->Stiles + (size_t)c->n_tiles_S * TS * TS;
    double *ea_out = single ? c->eab : ea_red;
    if (c->pair_mode == 6) {
        if (c->n_seg > 0) PROF(c, KID_SCHUR_PAIRS) launch_ring(c);
    } else if (c->pair_mode == 5) {
        if (c->n_seg > 0)
            PROF(c, KID_SCHUR_PAIRS) {
                switch (c->pair_G) {
                case 1: launch_segs<1>(c); break;
                case 2: launch_segs<2>(c); break;
                case 4: launch_segs<4>(c); break;
                case 8: launch_segs<8>(c); break;
                case 16: launch_segs<16>(c); break;
                default: launch_segs<32>(c); break;
                }
            }
    } else if (c->n_pchunk > 0) {
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
    }
    PROF(c, KID_S_FINALIZE) k_S_finalize<<<cdiv((long long)c->n_pair * 42, 128), 128, 0, c->stream>>>(c->n_pair, c->pair_k, c->pair_l, c->pair_chunk_ptr,
                                                                         c->pair_part, c->U, c->g, c->d_mu, single, c->tile_index,
                                                                         c->cam2pos, c->nt, c->Stiles, ea_out);
    c->st_launches += 3;
    LAUNCH_CHECK();
    if (!single) {
        psba_allreduce_sum(c, c->Stiles, (size_t)c->n_tiles_S * TS * TS + (size_t)c->N);   // S tiles + ea; fill-in tiles are zero on every rank
        k_add_U<<<cdiv(c->m * 42, 128), 128, 0, c->stream>>>(c->m, c->U, c->g, c->d_mu, c->tile_index, c->cam2pos, c->nt, c->Stiles, ea_red, c->eab);
        c->st_launches += 1;
        LAUNCH_CHECK();
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
    LAUNCH_CHECK();
}
