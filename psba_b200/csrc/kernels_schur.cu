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


// Quad variant of the pair pass (PSBA_PAIR_MODE=1; measured 1.23 ms against 1.03 ms: twice the sector look-ups).  The lane-per-triple kernel above keeps 42 sums and two 6x3
// blocks per thread (~170 registers, 8 warps per SM) and is bound by the latency of its gathers, not by their
// volume.  Here FOUR lanes share a triple: lane (qa, qb) owns the 3x3 quadrant rows 3qa.., columns 3qb.. of the
// block and reads only rows 3qa.. of W_a and rows 3qb.. of W_b (lanes of a quad that read the same rows are
// merged by the load unit), so a thread holds 12 sums and streams its operands row by row: three times the
// resident warps, three times the gathers in flight.  A group of G lanes (G/4 quads) owns one chunk of
// triples; the quads' sums are combined by an xor butterfly over the quad index (fixed order).
template <bool DIAG, int G>
__device__ __forceinline__ void pair_accumulate_q(long long beg, long long end, int quad, int qa, int qb, const int *__restrict__ tri_oa,
                                                  const int *__restrict__ tri_ob, const int *__restrict__ tri_pt,
                                                  const double *__restrict__ W, const double *__restrict__ Vinv,
                                                  const double *__restrict__ gb, double *acc)
{
    constexpr int Q = G / 4;
    long long t = beg + quad;
    int a_n = 0, b_n = 0, i_n = 0;
    if (t < end) { a_n = __ldg(tri_oa + t); b_n = DIAG ? a_n : __ldg(tri_ob + t); i_n = __ldg(tri_pt + t); }
#pragma unroll 1
    for (; t < end; t += Q) {
        const int a = a_n, b = b_n, i = i_n;
        if (t + Q < end) {                              // indices of the next triple fly with this triple's blocks
            a_n = __ldg(tri_oa + t + Q); b_n = DIAG ? a_n : __ldg(tri_ob + t + Q); i_n = __ldg(tri_pt + t + Q);
        }
        const double2 *vp = reinterpret_cast<const double2 *>(Vinv + (size_t)i * 6);
        const double2 v01 = __ldg(vp), v23 = __ldg(vp + 1), v45 = __ldg(vp + 2);
        const double *wbp = W + (size_t)b * 18 + 9 * qb, *wap = W + (size_t)a * 18 + 9 * qa;
        double wb[9], wa[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) wb[q] = __ldg(wbp + q);
#pragma unroll
        for (int q = 0; q < 9; ++q) wa[q] = __ldg(wap + q);
        double g0 = 0, g1 = 0, g2 = 0;
        if (DIAG) { const double *gp = gb + (size_t)i * 3; g0 = __ldg(gp); g1 = __ldg(gp + 1); g2 = __ldg(gp + 2); }
        const double i00 = v01.x, i10 = v01.y, i20 = v23.x, i11 = v23.y, i21 = v45.x, i22 = v45.y;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double w0 = wa[r * 3], w1 = wa[r * 3 + 1], w2 = wa[r * 3 + 2];
            const double y0 = w0 * i00 + w1 * i10 + w2 * i20;
            const double y1 = w0 * i10 + w1 * i11 + w2 * i21;
            const double y2 = w0 * i20 + w1 * i21 + w2 * i22;
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
                acc[r * 3 + cc] += y0 * wb[cc * 3] + y1 * wb[cc * 3 + 1] + y2 * wb[cc * 3 + 2];
            if (DIAG) acc[9 + r] += y0 * g0 + y1 * g1 + y2 * g2;
        }
    }
}

template <int G>
__global__ void __launch_bounds__(PAIR_CTA, 5) k_schur_pairs_q(int n_pchunk, const int *__restrict__ pchunk_pair,
                                                           const long long *__restrict__ pchunk_beg, const long long *__restrict__ pchunk_end,
                                                           const int *__restrict__ pair_k, const int *__restrict__ pair_l,
                                                           const int *__restrict__ tri_oa, const int *__restrict__ tri_ob,
                                                           const int *__restrict__ tri_pt, const double *__restrict__ W,
                                                           const double *__restrict__ Vinv, const double *__restrict__ gb,
                                                           double *__restrict__ part)
{
    const int lane = threadIdx.x % G, quad = lane >> 2, qa = (lane >> 1) & 1, qb = lane & 1;
    const int ch = blockIdx.x * (PAIR_CTA / G) + threadIdx.x / G;
    double acc[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) acc[q] = 0.0;
    bool diag = false;
    if (ch < n_pchunk) {
        const int pr = pchunk_pair[ch];
        diag = pair_k[pr] == pair_l[pr];
        if (diag) pair_accumulate_q<true, G>(pchunk_beg[ch], pchunk_end[ch], quad, qa, qb, tri_oa, tri_ob, tri_pt, W, Vinv, gb, acc);
        else pair_accumulate_q<false, G>(pchunk_beg[ch], pchunk_end[ch], quad, qa, qb, tri_oa, tri_ob, tri_pt, W, Vinv, gb, acc);
    }
#pragma unroll
    for (int w = G / 2; w >= 4; w >>= 1) {
#pragma unroll
        for (int q = 0; q < 12; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], w);
    }
    if (ch < n_pchunk && quad == 0) {
        double *out = part + (size_t)ch * 42;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) out[(3 * qa + r) * 6 + 3 * qb + cc] = acc[r * 3 + cc];
        if (diag && qb == 0) {
#pragma unroll
            for (int r = 0; r < 3; ++r) out[36 + 3 * qa + r] = acc[9 + r];
        }
    }
}

template <int G>
static void launch_pairs_q(psba_ctx *c)
{
    const int per_cta = PAIR_CTA / G;
    k_schur_pairs_q<G><<<cdiv(c->n_pchunk, per_cta), PAIR_CTA, 0, c->stream>>>(c->n_pchunk, c->pchunk_pair, c->pchunk_beg, c->pchunk_end,
                                                                            c->pair_k, c->pair_l, c->tri_oa, c->tri_ob, c->tri_pt, c->W,
                                                                            c->Vinv, c->g + c->N, c->pair_part);
}

// Staged variant of the pair-major pass (PSBA_PAIR_MODE=3; measured 1.19 ms against 1.03 ms: the line look-ups fall
// from 21 to ~6 per triple, but two stages of 10.75 KB per warp allow only 8 warps per SM and one round in flight).  The lane-per-triple kernel is bound by the number of
// cache lines its loads touch: every lane pulls its own 16-byte pieces (21 per triple), a warp instruction touches 32
// different lines and the L1 serves about one line per cycle (an L1 prefetch of the next triple, five more requests
// per triple, makes it 34 % slower; three times the occupancy does not help).  Here the warp fetches the operands
// of its 32 triples COOPERATIVELY: the 21 x 16-byte pieces of a triple are consecutive pieces of the warp's copy
// list, consecutive lanes take consecutive pieces (asynchronous copies, LDGSTS), so one instruction touches ~7
// lines instead of 32; the copies of the next round fly during the products of this one; every lane then reads its
// own 336 bytes from the warp's stage (stride 21 x 16 B: conflict-free).  No CTA-wide barrier anywhere.
#define STG_PIECES 21                                          // 9 (W_a) + 9 (W_b) + 3 (Vinv) pieces of 16 B
template <int G>
__global__ void __launch_bounds__(PAIR_CTA, 2) k_schur_pairs_s(int n_pchunk, const int *__restrict__ pchunk_pair,
                                                           const long long *__restrict__ pchunk_beg, const long long *__restrict__ pchunk_end,
                                                           const int *__restrict__ pair_k, const int *__restrict__ pair_l,
                                                           const int *__restrict__ tri_oa, const int *__restrict__ tri_ob,
                                                           const int *__restrict__ tri_pt, const double *__restrict__ W,
                                                           const double *__restrict__ Vinv, const double *__restrict__ gb,
                                                           double *__restrict__ part)
{
    extern __shared__ __align__(16) unsigned char stg_dyn[];   // per warp: two stages of 32 x 336 B, two index tables
    constexpr int NW = PAIR_CTA / 32;
    constexpr int STAGE = 32 * STG_PIECES * 16;
    const int tid = threadIdx.x, wrp = tid >> 5, ln = tid & 31;
    unsigned char *stage = stg_dyn + (size_t)wrp * 2 * STAGE;
    int *itab = reinterpret_cast<int *>(stg_dyn + (size_t)NW * 2 * STAGE) + wrp * 2 * 32 * 3;
    const int lane = tid % G;
    const int ch = blockIdx.x * (PAIR_CTA / G) + tid / G;
    double acc[42];
#pragma unroll
    for (int q = 0; q < 42; ++q) acc[q] = 0.0;
    bool diag = false;
    long long t = 0, end = 0;
    if (ch < n_pchunk) {
        const int pr = pchunk_pair[ch];
        diag = pair_k[pr] == pair_l[pr];
        t = pchunk_beg[ch] + lane; end = pchunk_end[ch];
    }
    // indices of this lane's triple of the round being issued (a < 0: none)
    auto load_idx = [&](long long tt, int &a, int &b, int &i) {
        a = -1; b = 0; i = 0;
        if (tt < end) { a = __ldg(tri_oa + tt); b = diag ? a : __ldg(tri_ob + tt); i = __ldg(tri_pt + tt); }
    };
    // the warp's copies of one round.  W blocks (9 pieces): eight lanes take the first eight pieces of a block (128
    // contiguous bytes: one line, two at most), four blocks per instruction, eight instructions for the 32 blocks; one
    // more instruction for the ninth piece of every block.  Vinv (3 pieces): eight blocks per instruction.
    auto issue = [&](int buf, int a, int b, int i) {
        int *tb = itab + buf * 96;
        tb[ln * 3] = a; tb[ln * 3 + 1] = b; tb[ln * 3 + 2] = i;
        __syncwarp();
        unsigned char *st = stage + buf * STAGE;
        const int sub8 = ln & 7, blk4 = ln >> 3;
#pragma unroll 2
        for (int u = 0; u < 8; ++u) {
            const int trip = blk4 + 4 * u;
            const int ta = tb[trip * 3], tbb = tb[trip * 3 + 1];
            if (ta >= 0) {
                unsigned char *dst = st + trip * (STG_PIECES * 16) + sub8 * 16;
                if (ta != tbb) cp_async16(dst, reinterpret_cast<const char *>(W + (size_t)ta * 18) + sub8 * 16);   // diagonal triples: W_a is W_b
                cp_async16(dst + 144, reinterpret_cast<const char *>(W + (size_t)tbb * 18) + sub8 * 16);
            }
        }
        if (a >= 0) {                                          // ninth piece of this lane's own blocks
            unsigned char *dst = st + ln * (STG_PIECES * 16) + 128;
            if (a != b) cp_async16(dst, reinterpret_cast<const char *>(W + (size_t)a * 18) + 128);
            cp_async16(dst + 144, reinterpret_cast<const char *>(W + (size_t)b * 18) + 128);
        }
        const int blk8 = ln / 3, sub3 = ln - blk8 * 3;          // lanes 0..23: eight Vinv blocks of three pieces
        if (ln < 24) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int trip = blk8 + 8 * u;
                const int ta = tb[trip * 3], ti = tb[trip * 3 + 2];
                if (ta >= 0) cp_async16(st + trip * (STG_PIECES * 16) + 288 + sub3 * 16, reinterpret_cast<const char *>(Vinv + (size_t)ti * 6) + sub3 * 16);
            }
        }
    };
    int a0, b0, i0, a1, b1, i1;
    load_idx(t, a0, b0, i0);
    load_idx(t + G, a1, b1, i1);
    issue(0, a0, b0, i0);
    cp_async_commit();
    for (int it = 0; __any_sync(0xffffffffu, a0 >= 0); ++it) {
        const int buf = it & 1;
        issue(buf ^ 1, a1, b1, i1);                            // next round's copies fly during this round's products
        cp_async_commit();
        int a2, b2, i2;
        load_idx(t + 2 * (long long)G, a2, b2, i2);            // indices two rounds ahead
        double g0 = 0, g1 = 0, g2 = 0;
        if (diag && a0 >= 0) { const double *gp = gb + (size_t)i0 * 3; g0 = __ldg(gp); g1 = __ldg(gp + 1); g2 = __ldg(gp + 2); }
        cp_async_wait<1>();
        __syncwarp();
        if (a0 >= 0) {
            const double2 *sp = reinterpret_cast<const double2 *>(stage + buf * STAGE + ln * STG_PIECES * 16);
            const double2 v01 = sp[18], v23 = sp[19], v45 = sp[20];
            const double i00 = v01.x, i10 = v01.y, i20 = v23.x, i11 = v23.y, i21 = v45.x, i22 = v45.y;
            double wb[18];
#pragma unroll
            for (int q = 0; q < 9; ++q) { const double2 w2 = sp[9 + q]; wb[2 * q] = w2.x; wb[2 * q + 1] = w2.y; }
            const int ao = a0 == b0 ? 9 : 0;                   // diagonal triples read W_b again
#pragma unroll
            for (int rp = 0; rp < 3; ++rp) {
                double wa[6];
#pragma unroll
                for (int q = 0; q < 3; ++q) { const double2 w2 = sp[ao + rp * 3 + q]; wa[2 * q] = w2.x; wa[2 * q + 1] = w2.y; }
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
                    acc[36 + r] += y0 * g0 + y1 * g1 + y2 * g2;
                }
            }
        }
        __syncwarp();                                          // the stage and its index table are free again
        t += G;
        a0 = a1; b0 = b1; i0 = i1; a1 = a2; b1 = b2; i1 = i2;
    }
    cp_async_wait<0>();
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
static void launch_pairs_s(psba_ctx *c)
{
    const int per_cta = PAIR_CTA / G;
    const int dyn = (PAIR_CTA / 32) * (2 * 32 * STG_PIECES * 16 + 2 * 96 * (int)sizeof(int));
    static bool attr_set = false;
    if (!attr_set) { CUDA_CHECK(cudaFuncSetAttribute(k_schur_pairs_s<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn)); attr_set = true; }
    k_schur_pairs_s<G><<<cdiv(c->n_pchunk, per_cta), PAIR_CTA, dyn, c->stream>>>(c->n_pchunk, c->pchunk_pair, c->pchunk_beg, c->pchunk_end,
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


// ---- Row sweep (PSBA_PAIR_MODE=2; measured 1.39 ms against 1.03 ms).  CTA = a segment of ONE camera row k; thread = one camera pair
// (k, l) of that row, l < k, with its 36 sums in registers for the whole segment; warp 0 owns the diagonal
// pair.  The segment is walked in chunks of visits (camera k looks at point i).  Per chunk:
//   1. every block the chunk needs -- for each visit the observations of point i with camera <= k, a
//      CONTIGUOUS prefix of the point's observations -- is copied into shared memory by 16-byte asynchronous
//      copies (coalesced: consecutive lanes, consecutive 16-byte pieces), each block exactly once per visit;
//   2. one thread per visit forms Y_ik = W_ik Vinv_i and Y_ik gb_i (compute_Yblks.cl:26-37, compute_ea.cl:27-33);
//   3. every pair thread consumes its triples of the chunk (they are a contiguous piece of the pair's run in the
//      pair-sorted triple list: ascending point, the order of comm3DIdx) reading Y_ik and W_il from shared memory
//      (compute_S.cl:44-52); the lanes of warp 0 stride the visits for the diagonal block and ea.
// The pair-major kernel gathers 336 B per triple from L2 in 16-byte pieces per lane (5.2 GB of sector traffic
// for 15 M triples); here a point's prefix is fetched once per visit (2.4 GB, whole lines) and the 108-FMA block
// product runs out of shared memory.  Fixed summation order: ascending point per pair, segments in order.
template <int NT, int B>
__global__ void __launch_bounds__(NT, (NT <= 256 && B <= 320 ? 2 : 1))
k_schur_rows(const int *__restrict__ seg_row, const int2 *__restrict__ seg_chunks, const int *__restrict__ seg_slot_base,
             const int *__restrict__ row_pair0, const int4 *__restrict__ chunk_desc, const int4 *__restrict__ vis_desc,
             const int2 *__restrict__ runs, const unsigned *__restrict__ tri_meta, const double *__restrict__ W,
             const double *__restrict__ Vinv, const double *__restrict__ gb, double *__restrict__ part)
{
    constexpr int POOL = ROW_POOL_BYTES(B);                    // blocks from the bottom (144 B), Y entries from the top (208 B)
    constexpr int MAXV = (B + ROW_MAXLEN + 2) / 3 + 1;         // visits per chunk (<= NT)
    constexpr int YENT = 26;                                   // doubles per Y entry: rows 0-2 at 0, rows 3-5 at 10, Y gb at 20
    static_assert(MAXV <= NT, "one visit per thread");
    extern __shared__ __align__(128) unsigned char pool[];     // two pools: chunk c+1 lands while chunk c is consumed
    __shared__ unsigned short own_slot[MAXV];
    __shared__ __align__(8) unsigned long long bar[2];         // bytes of the bulk copies of each pool
    const int seg = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = __ldg(seg_row + seg);
    const int2 cr = __ldg(seg_chunks + seg);
    const int pair0 = __ldg(row_pair0 + k), nslot = __ldg(row_pair0 + k + 1) - pair0;      // the diagonal is the last slot
    const int sbase = __ldg(seg_slot_base + seg);
    const bool diag_warp = warp == 0;
    const int ve = lane * (NT / 32) + warp;                    // the visit of a chunk this thread holds: every warp has some
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    // pair threads: a group of four lanes per pair, lane (qa, qb) owns the 3x3 quadrant rows 3qa.., columns 3qb.. of
    // the 6x6 block (a lane per pair leaves most of a warp idle: the runs of a chunk are 2-4 triples long and differ
    // from pair to pair); a group serves the pairs grp, grp + NG and grp + 2 NG of the row.  Three triple records are always in
    // flight per pair (they are the only global loads of the product phase).
    constexpr int NG = (NT - 32) / 4;
    const int grp = (tid - 32) >> 2, qa = (tid >> 1) & 1, qb = tid & 1;
    constexpr int NH = 3;                                      // pairs per group
    int cur[NH], rend[NH];
    unsigned m0[NH], m1[NH], m2[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) { cur[h] = rend[h] = 0; m0[h] = m1[h] = m2[h] = 0xffffffffu; }   // never a chunk number
    if (!diag_warp) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const int slot = grp + h * NG;
            if (slot < nslot - 1) {
                const int2 r = __ldg(runs + sbase + slot);
                cur[h] = r.x; rend[h] = r.y;
                if (cur[h] < rend[h]) m0[h] = __ldg(tri_meta + cur[h]);
                if (cur[h] + 1 < rend[h]) m1[h] = __ldg(tri_meta + cur[h] + 1);
                if (cur[h] + 2 < rend[h]) m2[h] = __ldg(tri_meta + cur[h] + 2);
            }
        }
    }
    // every copy of a chunk: the thread that holds a visit sends its prefix as ONE bulk copy (the prefix is contiguous
    // in W) and its Vinv_i / gb_i as 16- and 8-byte asynchronous copies into the visit's Y entry
    auto issue = [&](const int4 &cd, const int4 &d, int buf) {
        unsigned char *pl = pool + buf * POOL;
        if (tid == 0) mbar_expect_tx(&bar[buf], (unsigned)cd.z * 144u);
        if (ve < cd.y) {
            fence_proxy_async();
            bulk_g2s(pl + d.z * 144, W + (size_t)d.x * 18, (unsigned)d.y * 144u, &bar[buf]);
            double *ent = reinterpret_cast<double *>(pl + POOL) - (ve + 1) * YENT;
            const double *vp = Vinv + (size_t)d.w * 6, *gp = gb + (size_t)d.w * 3;
            cp_async16(ent, vp); cp_async16(ent + 2, vp + 2); cp_async16(ent + 4, vp + 4);
            cp_async8(ent + 6, gp); cp_async8(ent + 7, gp + 1); cp_async8(ent + 8, gp + 2);
        }
    };
    double acc[27];                                            // pair lanes: 3 x 9; diagonal warp: 21 + 6
#pragma unroll
    for (int q = 0; q < 27; ++q) acc[q] = 0.0;
    const int4 zero4 = make_int4(0, 0, 0, 0);
    // cdN / dN: chunk record and this thread's visit record of the chunk N steps ahead (records beyond the segment: 0)
    int4 cd0 = cr.x < cr.y ? __ldg(chunk_desc + cr.x) : zero4;
    int4 cd1 = cr.x + 1 < cr.y ? __ldg(chunk_desc + cr.x + 1) : zero4;
    int4 cd2 = cr.x + 2 < cr.y ? __ldg(chunk_desc + cr.x + 2) : zero4;
    int4 d0 = ve < cd0.y ? __ldg(vis_desc + cd0.x + ve) : zero4;
    int4 d1 = ve < cd1.y ? __ldg(vis_desc + cd1.x + ve) : zero4;
    __syncthreads();                                           // barriers initialised
    issue(cd0, d0, 0);
    cp_async_commit();
    for (int c = cr.x, crel = 0; c < cr.y; ++c, ++crel) {
        const int buf = crel & 1;
        unsigned char *pl = pool + buf * POOL;
        if (c + 1 < cr.y) issue(cd1, d1, buf ^ 1);
        cp_async_commit();
        // two chunks ahead: visit records (their chunk record arrived an iteration ago); three ahead: the chunk record
        const int4 d2 = ve < cd2.y ? __ldg(vis_desc + cd2.x + ve) : zero4;
        const int4 cd3 = c + 3 < cr.y ? __ldg(chunk_desc + c + 3) : zero4;
        const int nvc = cd0.y;
        cp_async_wait<1>();
        mbar_wait(&bar[buf], (crel >> 1) & 1);
        __syncthreads();
        double *stage = reinterpret_cast<double *>(pl);
        double *ytop = reinterpret_cast<double *>(pl + POOL);
        // Y entries: Y_ik = W_ik Vinv_i (compute_Yblks.cl:26-37), Y_ik gb_i (compute_ea.cl:27-33)
        if (ve < nvc) {
            const int own = d0.z + d0.y - 1;                                                // the visit's own block: last of its prefix
            own_slot[ve] = (unsigned short)own;
            const double2 *wp = reinterpret_cast<const double2 *>(stage + own * 18);
            double2 *yp = reinterpret_cast<double2 *>(ytop - (ve + 1) * YENT);
            double w[18];
#pragma unroll
            for (int q = 0; q < 9; ++q) { const double2 w2 = wp[q]; w[2 * q] = w2.x; w[2 * q + 1] = w2.y; }
            const double2 a01 = yp[0], a23 = yp[1], a45 = yp[2], g01 = yp[3];
            const double g2 = reinterpret_cast<const double *>(yp)[8];
            const double i00 = a01.x, i10 = a01.y, i20 = a23.x, i11 = a23.y, i21 = a45.x, i22 = a45.y;
            double y[26];
            y[9] = 0.0; y[19] = 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const int o3 = r < 3 ? r * 3 : 10 + (r - 3) * 3;
                const double w0 = w[r * 3], w1 = w[r * 3 + 1], w2 = w[r * 3 + 2];
                y[o3] = w0 * i00 + w1 * i10 + w2 * i20;
                y[o3 + 1] = w0 * i10 + w1 * i11 + w2 * i21;
                y[o3 + 2] = w0 * i20 + w1 * i21 + w2 * i22;
                y[20 + r] = y[o3] * g01.x + y[o3 + 1] * g01.y + y[o3 + 2] * g2;
            }
#pragma unroll
            for (int q = 0; q < 13; ++q) yp[q] = make_double2(y[2 * q], y[2 * q + 1]);
        }
        __syncthreads();
        // block products out of shared memory (compute_S.cl:44-52)
        if (!diag_warp) {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                while ((m0[h] >> 18) == (unsigned)crel) {
                    const unsigned mt = m0[h];
                    m0[h] = m1[h]; m1[h] = m2[h];
                    m2[h] = cur[h] + 3 < rend[h] ? __ldg(tri_meta + cur[h] + 3) : 0xffffffffu;
                    ++cur[h];
                    // rows 3qa..3qa+2 of Y (entry doubles 10qa..10qa+8: 16-byte aligned), rows 3qb..3qb+2 of W_il (doubles 9qb..)
                    const double2 *yp = reinterpret_cast<const double2 *>(ytop - (((mt >> 10) & 255u) + 1) * YENT + 10 * qa);
                    const double *wp = stage + (mt & 1023u) * 18 + 9 * qb;
                    double y[10], w[9];
#pragma unroll
                    for (int q = 0; q < 5; ++q) { const double2 a = yp[q]; y[2 * q] = a.x; y[2 * q + 1] = a.y; }
#pragma unroll
                    for (int q = 0; q < 9; ++q) w[q] = wp[q];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            double t = acc[h * 9 + r * 3 + cc];
                            t = fma(y[r * 3], w[cc * 3], t); t = fma(y[r * 3 + 1], w[cc * 3 + 1], t); t = fma(y[r * 3 + 2], w[cc * 3 + 2], t);
                            acc[h * 9 + r * 3 + cc] = t;
                        }
                }
            }
        } else {
            for (int e = lane; e < nvc; e += 32) {
                const double2 *yp = reinterpret_cast<const double2 *>(ytop - (e + 1) * YENT);
                const double2 *wp = reinterpret_cast<const double2 *>(stage + (int)own_slot[e] * 18);
                double ye[26], w[18];
#pragma unroll
                for (int q = 0; q < 13; ++q) { const double2 a = yp[q]; ye[2 * q] = a.x; ye[2 * q + 1] = a.y; }
#pragma unroll
                for (int q = 0; q < 9; ++q) { const double2 b = wp[q]; w[2 * q] = b.x; w[2 * q + 1] = b.y; }
                int q = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const int o3 = r < 3 ? r * 3 : 10 + (r - 3) * 3;
#pragma unroll
                    for (int cc = 0; cc <= r; ++cc, ++q)
                        acc[q] += ye[o3] * w[cc * 3] + ye[o3 + 1] * w[cc * 3 + 1] + ye[o3 + 2] * w[cc * 3 + 2];
                }
#pragma unroll
                for (int r = 0; r < 6; ++r) acc[21 + r] += ye[20 + r];
            }
        }
        __syncthreads();                                                                    // this pool is free again
        cd0 = cd1; cd1 = cd2; cd2 = cd3; d0 = d1; d1 = d2;
    }
    cp_async_wait<0>();
    if (!diag_warp) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const int slot = grp + h * NG;
            if (slot < nslot - 1) {
                double *out = part + (size_t)(sbase + slot) * 42;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) out[(3 * qa + r) * 6 + 3 * qb + cc] = acc[h * 9 + r * 3 + cc];
            }
        }
    } else {
#pragma unroll
        for (int w = 16; w > 0; w >>= 1) {
#pragma unroll
            for (int q = 0; q < 27; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], w);
        }
        double *out = part + (size_t)(sbase + nslot - 1) * 42;
        int q = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int cc = 0; cc <= r; ++cc, ++q)
                if (lane == (q & 31)) { out[r * 6 + cc] = acc[q]; out[cc * 6 + r] = acc[q]; }
#pragma unroll
        for (int r = 0; r < 6; ++r)
            if (lane == r) out[36 + r] = acc[21 + r];
    }
}

template <int NT, int B>
static void launch_rows(psba_ctx *c)
{
    static bool attr_set = false;
    const int dyn = 2 * ROW_POOL_BYTES(B);
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(k_schur_rows<NT, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
        if (getenv("PSBA_ROW_DEBUG")) {
            int nb = 0;
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_schur_rows<NT, B>, NT, dyn));
            fprintf(stderr, "psba: k_schur_rows<%d,%d>: %d CTAs/SM, %d segments, %d chunks\n", NT, B, nb, c->n_rseg, c->n_rchunk);
        }
        attr_set = true;
    }
    k_schur_rows<NT, B><<<c->n_rseg, NT, dyn, c->stream>>>(c->rseg_row, c->rseg_chunks, c->rseg_slot_base, c->row_pair0, c->rchunk_desc,
                                                          c->vis_desc, c->rseg_runs, c->tri_meta, c->W, c->Vinv, c->g + c->N, c->pair_part);
}


// ---- Row sweep WITHOUT CTA-wide barriers (PSBA_PAIR_MODE=4; measured 0.99 ms against 1.03 ms).  Same tables and the same order of sums as
// k_schur_rows; what changes is who waits for whom.  Warp 0 is the PRODUCER: it copies the blocks of the next chunk of
// visits (eight lanes per block: 128 contiguous bytes, then the ninth piece), Vinv_i and gb_i of every visit into
// the other half of a two-stage pool with asynchronous copies whose completion arrives on the stage's `full`
// mbarrier, and accumulates the diagonal pair of the current chunk.  The seven PAIR warps wait on `full`, consume
// their triples of the chunk at their own pace (Y_ik = W_ik Vinv_i is formed on the fly from the visit's own staged
// block: no Y pass, no shared Y entries) and arrive on the stage's `empty` mbarrier; a fast warp is up to a chunk
// ahead of a slow one.  Nothing in the loop is a __syncthreads.
#define FLOW_NT 256
#define FLOW_REC 10                                            // doubles per visit record at the top of a stage: Vinv (6), gb (3), pad
template <int FLOW_B, int FLOW_S, int FLOW_ML>
__global__ void __launch_bounds__(FLOW_NT, 2)
k_schur_flow(const int *__restrict__ seg_row, const int2 *__restrict__ seg_chunks, const int *__restrict__ seg_slot_base,
             const int *__restrict__ row_pair0, const int4 *__restrict__ chunk_desc, const int4 *__restrict__ vis_desc,
             const int *__restrict__ blk_src, const int2 *__restrict__ runs, const unsigned *__restrict__ tri_meta,
             const double *__restrict__ W, const double *__restrict__ Vinv, const double *__restrict__ gb, double *__restrict__ part)
{
    constexpr int MAXB = FLOW_B + FLOW_ML + 2;                 // staged blocks per chunk
    constexpr int POOL = MAXB * 144;                           // blocks from the bottom, visit records from the top
    constexpr int NBR = (MAXB + 31) / 32;                      // block-source registers per producer lane
    constexpr int NVR = (MAXB / 3 + 1 + 31) / 32;              // visits per producer lane
    constexpr int NPW = FLOW_NT / 32 - 2;                      // pair warps (warp 0: producer, warp 1: diagonal pair)
    extern __shared__ __align__(128) unsigned char pool[];
    __shared__ __align__(8) unsigned long long full[FLOW_S], empty[FLOW_S];
    const int seg = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = __ldg(seg_row + seg);
    const int2 cr = __ldg(seg_chunks + seg);
    const int nch = cr.y - cr.x;
    const int pair0 = __ldg(row_pair0 + k), nslot = __ldg(row_pair0 + k + 1) - pair0;      // the diagonal is the last slot
    const int sbase = __ldg(seg_slot_base + seg);
    if (tid == 0) {
        for (int q = 0; q < FLOW_S; ++q) { mbar_init(&full[q], 32); mbar_init(&empty[q], NPW + 1); }
        mbar_fence_init();
    }
    __syncthreads();
    double acc[27];
#pragma unroll
    for (int q = 0; q < 27; ++q) acc[q] = 0.0;
    const int4 zero4 = make_int4(0, 0, 0, 0);

    if (warp == 0) {
        // ================= producer =================
        int srcr[NBR];                                         // observation behind block lane + 32 r of the chunk to issue
        int pt[NVR];                                           // point of visit lane + 32 r of the chunk to issue
        auto fetch = [&](const int4 &cd) {                     // what the issue of a chunk needs, one chunk ahead
#pragma unroll
            for (int r = 0; r < NBR; ++r) srcr[r] = lane + 32 * r < cd.z ? __ldg(blk_src + cd.w + lane + 32 * r) : 0;
#pragma unroll
            for (int r = 0; r < NVR; ++r) pt[r] = lane + 32 * r < cd.y ? __ldg(vis_desc + cd.x + lane + 32 * r).w : 0;
        };
        auto issue = [&](const int4 &cd, int st) {
            unsigned char *pl = pool + st * POOL;
            const int sub8 = lane & 7, blk4 = lane >> 3;
#pragma unroll
            for (int r = 0; r < NBR; ++r) {                    // blocks 32 r .. 32 r + 31: eight instructions of four blocks
                if (32 * r < cd.z) {
#pragma unroll
                    for (int gq = 0; gq < 8; ++gq) {
                        const int b = 32 * r + 4 * gq + blk4;
                        const int so = __shfl_sync(0xffffffffu, srcr[r], 4 * gq + blk4);
                        if (b < cd.z) cp_async16(pl + b * 144 + sub8 * 16, reinterpret_cast<const char *>(W + (size_t)so * 18) + sub8 * 16);
                    }
                    const int b = 32 * r + lane;               // ninth piece of this lane's own block
                    if (b < cd.z) cp_async16(pl + b * 144 + 128, reinterpret_cast<const char *>(W + (size_t)srcr[r] * 18) + 128);
                }
            }
#pragma unroll
            for (int r = 0; r < NVR; ++r) {
                const int e = lane + 32 * r;
                if (e < cd.y) {
                    double *rec = reinterpret_cast<double *>(pl + POOL) - (e + 1) * FLOW_REC;
                    const double *vp = Vinv + (size_t)pt[r] * 6, *gp = gb + (size_t)pt[r] * 3;
                    cp_async16(rec, vp); cp_async16(rec + 2, vp + 2); cp_async16(rec + 4, vp + 4);
                    cp_async8(rec + 6, gp); cp_async8(rec + 7, gp + 1); cp_async8(rec + 8, gp + 2);
                }
            }
            cp_async_mbar_arrive(&full[st]);
        };
        int4 cd1 = nch > 0 ? __ldg(chunk_desc + cr.x) : zero4;
        int4 cd2 = nch > 1 ? __ldg(chunk_desc + cr.x + 1) : zero4;
        fetch(cd1);
        for (int j = 0; j < nch; ++j) {                        // runs up to FLOW_S chunks ahead of the slowest consumer
            const int st = j % FLOW_S;
            if (j >= FLOW_S) mbar_wait(&empty[st], (j / FLOW_S - 1) & 1);                    // chunk j - FLOW_S has been consumed by every warp
            issue(cd1, st);
            const int4 cd3 = j + 2 < nch ? __ldg(chunk_desc + cr.x + j + 2) : zero4;
            if (j + 1 < nch) fetch(cd2);
            cd1 = cd2; cd2 = cd3;
        }
        cp_async_wait<0>();
        return;
    }
    if (warp == 1) {
        // ================= diagonal pair: lanes stride the visits of a chunk =================
        int4 cdn = nch > 0 ? __ldg(chunk_desc + cr.x) : zero4;
        int own_n[NVR], nv_n = cdn.y;
#pragma unroll
        for (int r = 0; r < NVR; ++r) { const int4 d = lane + 32 * r < cdn.y ? __ldg(vis_desc + cdn.x + lane + 32 * r) : zero4; own_n[r] = d.z + d.y - 1; }
        for (int c = 0; c < nch; ++c) {
            const int st = c % FLOW_S;
            int own[NVR]; const int nv = nv_n;
#pragma unroll
            for (int r = 0; r < NVR; ++r) own[r] = own_n[r];
            if (c + 1 < nch) {                                 // next chunk's own-block slots fly during this chunk
                cdn = __ldg(chunk_desc + cr.x + c + 1); nv_n = cdn.y;
#pragma unroll
                for (int r = 0; r < NVR; ++r) { const int4 d = lane + 32 * r < cdn.y ? __ldg(vis_desc + cdn.x + lane + 32 * r) : zero4; own_n[r] = d.z + d.y - 1; }
            }
            mbar_wait(&full[st], (c / FLOW_S) & 1);
            const unsigned char *pl = pool + st * POOL;
            const double *stage = reinterpret_cast<const double *>(pl);
            const double *top = reinterpret_cast<const double *>(pl + POOL);
#pragma unroll 1
            for (int r = 0; r < NVR; ++r) {
                const int e = lane + 32 * r;
                if (e < nv) {
                    const double2 *rp = reinterpret_cast<const double2 *>(top - (e + 1) * FLOW_REC);
                    const double2 a01 = rp[0], a23 = rp[1], a45 = rp[2], g01 = rp[3];
                    const double g2 = reinterpret_cast<const double *>(rp)[8];
                    const double i00 = a01.x, i10 = a01.y, i20 = a23.x, i11 = a23.y, i21 = a45.x, i22 = a45.y;
                    const double2 *wp = reinterpret_cast<const double2 *>(stage + own[r] * 18);
                    double w[18], y[18];
#pragma unroll
                    for (int q = 0; q < 9; ++q) { const double2 b = wp[q]; w[2 * q] = b.x; w[2 * q + 1] = b.y; }
#pragma unroll
                    for (int rr = 0; rr < 6; ++rr) {
                        const double w0 = w[rr * 3], w1 = w[rr * 3 + 1], w2 = w[rr * 3 + 2];
                        y[rr * 3] = w0 * i00 + w1 * i10 + w2 * i20;
                        y[rr * 3 + 1] = w0 * i10 + w1 * i11 + w2 * i21;
                        y[rr * 3 + 2] = w0 * i20 + w1 * i21 + w2 * i22;
                        acc[21 + rr] += y[rr * 3] * g01.x + y[rr * 3 + 1] * g01.y + y[rr * 3 + 2] * g2;
                    }
                    int q = 0;
#pragma unroll
                    for (int rr = 0; rr < 6; ++rr)
#pragma unroll
                        for (int cc = 0; cc <= rr; ++cc, ++q)
                            acc[q] += y[rr * 3] * w[cc * 3] + y[rr * 3 + 1] * w[cc * 3 + 1] + y[rr * 3 + 2] * w[cc * 3 + 2];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
#pragma unroll
        for (int w = 16; w > 0; w >>= 1) {
#pragma unroll
            for (int q = 0; q < 27; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], w);
        }
        double *out = part + (size_t)(sbase + nslot - 1) * 42;
        int q = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int cc = 0; cc <= r; ++cc, ++q)
                if (lane == (q & 31)) { out[r * 6 + cc] = acc[q]; out[cc * 6 + r] = acc[q]; }
#pragma unroll
        for (int r = 0; r < 6; ++r)
            if (lane == r) out[36 + r] = acc[21 + r];
        return;
    }

    // ================= pair warps =================
    constexpr int NG = (FLOW_NT - 64) / 4, NH = 3;
    const int grp = (tid - 64) >> 2, qa = (tid >> 1) & 1, qb = tid & 1;
    int cur[NH], rend[NH];
    unsigned m0[NH], m1[NH], m2[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
        cur[h] = rend[h] = 0; m0[h] = m1[h] = m2[h] = 0xffffffffu;                            // never a chunk number
        const int slot = grp + h * NG;
        if (slot < nslot - 1) {
            const int2 r = __ldg(runs + sbase + slot);
            cur[h] = r.x; rend[h] = r.y;
            if (cur[h] < rend[h]) m0[h] = __ldg(tri_meta + cur[h]);
            if (cur[h] + 1 < rend[h]) m1[h] = __ldg(tri_meta + cur[h] + 1);
            if (cur[h] + 2 < rend[h]) m2[h] = __ldg(tri_meta + cur[h] + 2);
        }
    }
    for (int c = 0; c < nch; ++c) {
        const int st = c % FLOW_S;
        mbar_wait(&full[st], (c / FLOW_S) & 1);
        const unsigned char *pl = pool + st * POOL;
        const double *stage = reinterpret_cast<const double *>(pl);
        const double *top = reinterpret_cast<const double *>(pl + POOL);
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            while ((m0[h] >> 26) == (unsigned)c) {
                const unsigned mt = m0[h];
                m0[h] = m1[h]; m1[h] = m2[h];
                m2[h] = cur[h] + 3 < rend[h] ? __ldg(tri_meta + cur[h] + 3) : 0xffffffffu;
                ++cur[h];
                const double2 *rp = reinterpret_cast<const double2 *>(top - (((mt >> 18) & 255u) + 1) * FLOW_REC);
                const double2 a01 = rp[0], a23 = rp[1], a45 = rp[2];
                const double i00 = a01.x, i10 = a01.y, i20 = a23.x, i11 = a23.y, i21 = a45.x, i22 = a45.y;
                const double *wap = stage + (mt & 511u) * 18 + 9 * qa;                       // rows 3qa.. of the visit's own block
                const double *wbp = stage + ((mt >> 9) & 511u) * 18 + 9 * qb;                // rows 3qb.. of W_il
                double wa[9], wb[9];
#pragma unroll
                for (int q = 0; q < 9; ++q) { wa[q] = wap[q]; wb[q] = wbp[q]; }
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double w0 = wa[r * 3], w1 = wa[r * 3 + 1], w2 = wa[r * 3 + 2];
                    const double y0 = w0 * i00 + w1 * i10 + w2 * i20;
                    const double y1 = w0 * i10 + w1 * i11 + w2 * i21;
                    const double y2 = w0 * i20 + w1 * i21 + w2 * i22;
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        double t = acc[h * 9 + r * 3 + cc];
                        t = fma(y0, wb[cc * 3], t); t = fma(y1, wb[cc * 3 + 1], t); t = fma(y2, wb[cc * 3 + 2], t);
                        acc[h * 9 + r * 3 + cc] = t;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
#pragma unroll
    for (int h = 0; h < NH; ++h) {
        const int slot = grp + h * NG;
        if (slot < nslot - 1) {
            double *out = part + (size_t)(sbase + slot) * 42;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) out[(3 * qa + r) * 6 + 3 * qb + cc] = acc[h * 9 + r * 3 + cc];
        }
    }
}

template <int FB, int FS, int FML>
static void launch_flow_t(psba_ctx *c)
{
    static bool attr_set = false;
    const int dyn = FS * (FB + FML + 2) * 144;
    if (!attr_set) { CUDA_CHECK(cudaFuncSetAttribute(k_schur_flow<FB, FS, FML>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn)); attr_set = true; }
    k_schur_flow<FB, FS, FML><<<c->n_rseg, FLOW_NT, dyn, c->stream>>>(c->rseg_row, c->rseg_chunks, c->rseg_slot_base, c->row_pair0, c->rchunk_desc, c->vis_desc,
                                                       c->rblk_src, c->rseg_runs, c->tri_meta, c->W, c->Vinv, c->g + c->N, c->pair_part);
}
static void launch_flow(psba_ctx *c)
{
    static const int deep = getenv("PSBA_FLOW_DEEP") ? atoi(getenv("PSBA_FLOW_DEEP")) : 0;
    if (c->row_budget == 304 && deep) launch_flow_t<304, 4, 64>(c);
    else if (c->row_budget == 192 && deep) launch_flow_t<192, 6, 32>(c);
    else if (c->row_budget == 304) launch_flow_t<304, 2, 64>(c);
    else if (c->row_budget == 192) launch_flow_t<192, 3, 32>(c);
    else if (c->row_budget == 144) launch_flow_t<144, 4, 32>(c);
    else launch_flow_t<112, 5, 32>(c);
}

// per pair block of the row sweep: fixed-order sum over the segments of its row, then as k_S_finalize
__global__ void k_S_finalize_rows(int n_pair, const int *__restrict__ pair_k, const int *__restrict__ pair_l,
                                  const int *__restrict__ row_pair0, const int *__restrict__ row_seg_ptr,
                                  const int *__restrict__ seg_slot_base, const double *__restrict__ part,
                                  const double *__restrict__ U, const double *__restrict__ ga, double mu, int with_U,
                                  const int *__restrict__ tile_index, const int *__restrict__ cam2pos, int nt,
                                  double *__restrict__ Stiles, double *__restrict__ ea)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int pr = t / 42, v = t - pr * 42;
    if (pr >= n_pair) return;
    const int k = pair_k[pr], l = pair_l[pr];
    if (v >= 36 && k != l) return;
    const int slot = pr - row_pair0[k];
    double s = 0.0;
    for (int sg = row_seg_ptr[k]; sg < row_seg_ptr[k + 1]; ++sg) s += part[(size_t)(seg_slot_base[sg] + slot) * 42 + v];
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
                        double *__restrict__ Stiles, const double *__restrict__ ea_red, double *__restrict__ ea)
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
    if (c->rows_ok) {
        if (c->n_rseg > 0)
            PROF(c, KID_SCHUR_PAIRS) {
                if (c->pair_mode == 4) launch_flow(c);
                else if (c->rows_nt == 256 && c->row_budget == 304) launch_rows<256, 304>(c);
                else if (c->rows_nt == 256) launch_rows<256, 640>(c);
                else if (c->row_budget == 304) launch_rows<544, 304>(c);
                else launch_rows<544, 640>(c);
            }
        PROF(c, KID_S_FINALIZE) k_S_finalize_rows<<<cdiv((long long)c->n_pair * 42, 128), 128, 0, c->stream>>>(c->n_pair, c->pair_k, c->pair_l, c->row_pair0,
                                                                             c->row_seg_ptr, c->rseg_slot_base, c->pair_part, c->U, c->g, mu, single,
                                                                             c->tile_index, c->cam2pos, c->nt, c->Stiles, ea_out);
    } else {
        if (c->n_pchunk > 0) {
            PROF(c, KID_SCHUR_PAIRS) {
                if (c->pair_mode == 3) {
                    switch (c->pair_G) {
                    case 1: launch_pairs_s<1>(c); break;
                    case 2: launch_pairs_s<2>(c); break;
                    case 4: launch_pairs_s<4>(c); break;
                    case 8: launch_pairs_s<8>(c); break;
                    case 16: launch_pairs_s<16>(c); break;
                    default: launch_pairs_s<32>(c); break;
                    }
                } else if (c->pair_mode == 1) {
                    switch (c->pair_G) {
                    case 4: launch_pairs_q<4>(c); break;
                    case 8: launch_pairs_q<8>(c); break;
                    case 16: launch_pairs_q<16>(c); break;
                    default: launch_pairs_q<32>(c); break;
                    }
                } else {
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
        }
        PROF(c, KID_S_FINALIZE) k_S_finalize<<<cdiv((long long)c->n_pair * 42, 128), 128, 0, c->stream>>>(c->n_pair, c->pair_k, c->pair_l, c->pair_chunk_ptr,
                                                                             c->pair_part, c->U, c->g, mu, single, c->tile_index,
                                                                             c->cam2pos, c->nt, c->Stiles, ea_out);
    }
    c->st_launches += 3;
    LAUNCH_CHECK();
    if (!single) {
        psba_allreduce_sum(c, c->Stiles, (size_t)c->n_tiles_S * TS * TS + (size_t)c->N);   // S tiles + ea; fill-in tiles are zero on every rank
        k_add_U<<<cdiv(c->m * 42, 128), 128, 0, c->stream>>>(c->m, c->U, c->g, mu, c->tile_index, c->cam2pos, c->nt, c->Stiles, ea_red, c->eab);
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
