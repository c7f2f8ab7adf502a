// kernels_backsub.cu -- point back-substitution, parameter update and cost of the candidate, fused
// per point chunk; plus the small vector kernels of the trust-region driver.
//
// Replaces kern_compute_eb (CL_files/compute_eb.cl:6-41), kern_compute_dpb (compute_dpb.cl:6-35),
// kern_compute_newp (compute_newp.cl:6-26), the re-evaluation kern_compute_exQT
// (compute_exQT.cl:18-71) and the host reductions compute_L2_sq / compute_rho
// (PSBA/misc.cpp:151-157, PSBA/levmar.cpp:271-280).  kern_update_p (update_p.cl:6-25) is a
// pointer swap of the two parameter sets.
#include "dev_math.cuh"
#include <algorithm>

// candidate cameras = cams + dpa, plus the camera part of the step scalars
__global__ void __launch_bounds__(1024) k_newcams(int N, const double *__restrict__ cams, const double *__restrict__ dpa, const double *__restrict__ ga,
                                                  const double *__restrict__ mu_p, double *__restrict__ newcams, double *__restrict__ scal3)
{
    const double mu = __ldg(mu_p);                               // damping term of this solve (device-resident: psba_set_scalars)
    // one CTA on the critical path of every try: 1024 threads so that the N = 6m entries are a dozen independent
    // loads per thread (25 us with 256 threads on 2 000 cameras); fixed-order tree
    __shared__ double s0[1024], s1[1024], s2[1024];
    double a = 0.0, b = 0.0, p2 = 0.0;
    for (int k = threadIdx.x; k < N; k += 1024) {
        const double d = dpa[k], x = cams[k] + d;
        newcams[k] = x;
        a += d * d;
        b += d * (mu * d + ga[k]);
        p2 += x * x;
    }
    s0[threadIdx.x] = a; s1[threadIdx.x] = b; s2[threadIdx.x] = p2;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) { s0[threadIdx.x] += s0[threadIdx.x + w]; s1[threadIdx.x] += s1[threadIdx.x + w]; s2[threadIdx.x] += s2[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { scal3[0] = s0[0]; scal3[1] = s1[0]; scal3[2] = s2[0]; }
}

// CTA (128 threads) = one chunk of whole points (same chunks as k_lin_points).
//  phase A: the wave's W tile (contiguous in HBM) is copied to shared memory with coalesced 16-byte
//           loads; thread per observation: t = W_ij^T dpa_j -> shared
//  phase B (owner thread per point): eb = gb - sum_j t (ascending camera, compute_eb.cl:27-37),
//           dpb = Vinv eb (compute_dpb.cl:23-32), candidate point, |dpb|^2, dpb.(mu dpb + gb)
//  phase C: the candidate cameras' projection entries (96 B each) are staged the same way; thread per
//           observation: residual of the candidate -> ||e||^2
// EVAL=false stops after phase B (trust region: the step is formed on the host side first).
#define PROJ_LD 14         // doubles per staged projection entry (12 + pad: conflict-free LDS.128)
template <bool EVAL, bool EXT>
__global__ void __launch_bounds__(PT_CTA, 4) k_backsub(const int *__restrict__ chunk_list, const int4 *__restrict__ ptdesc, const int *__restrict__ pt_ptr,
                                                      const int *__restrict__ iidx, const int *__restrict__ jidx,
                                                      const double *__restrict__ impts, const double *__restrict__ W,
                                                      const double *__restrict__ Vinv, const double *__restrict__ gb,
                                                      const double *__restrict__ dpa, const double *__restrict__ pts,
                                                      const double *__restrict__ newcache, const double *__restrict__ mu_p,
                                                      double *__restrict__ eb, double *__restrict__ dpb, double *__restrict__ newpts,
                                                      double *__restrict__ part, psba_ext ext)
{
    const double mu = __ldg(mu_p);
    __shared__ __align__(16) double stage[PT_CTA * 18];      // W tile of the wave, then the projection entries (128*14 <= 128*18)
    double *pstage = stage;
    __shared__ double sh[3][PT_CTA];
    __shared__ double shx[3][PT_CTA];
    __shared__ double red[4][PT_CTA / 32];
    __shared__ int sj[PT_CTA];
    const int tid = threadIdx.x;
    // a CTA is a chain of dependent L2 / HBM round trips: one 16-byte chunk descriptor, then every
    // independent load of the chunk (W tile, indices, measurements, the owner's point data) at once, then
    // the loads that need a camera index (dpa, projection entries)
    const int cidx = chunk_list ? chunk_list[blockIdx.x] : blockIdx.x;
    const int4 ds = __ldg(ptdesc + cidx);
    const int p0 = ds.x, p1 = ds.y, o0 = ds.z, o1 = ds.w;
    const int np = p1 - p0;
    const bool single = o1 - o0 <= PT_CTA;                   // the common case: the whole chunk is one wave
    double acc0 = 0, acc1 = 0, acc2 = 0;
    int my_a = 0, my_b = 0;
    double g0 = 0, g1 = 0, g2 = 0, i00 = 0, i10 = 0, i20 = 0, i11 = 0, i21 = 0, i22 = 0, px = 0, py = 0, pz = 0;
    if (tid < np) {
        const int p = p0 + tid;
        my_a = __ldg(pt_ptr + p); my_b = __ldg(pt_ptr + p + 1);
        g0 = __ldg(gb + (size_t)p * 3); g1 = __ldg(gb + (size_t)p * 3 + 1); g2 = __ldg(gb + (size_t)p * 3 + 2);
        const double2 *vi = reinterpret_cast<const double2 *>(Vinv + (size_t)p * 6);
        const double2 v01 = __ldg(vi), v23 = __ldg(vi + 1), v45 = __ldg(vi + 2);
        i00 = v01.x; i10 = v01.y; i20 = v23.x; i11 = v23.y; i21 = v45.x; i22 = v45.y;
        if (EVAL) { px = __ldg(pts + (size_t)p * 3); py = __ldg(pts + (size_t)p * 3 + 1); pz = __ldg(pts + (size_t)p * 3 + 2); }
    }
    int lp1 = 0;
    double2 mm1 = make_double2(0.0, 0.0);

    for (int base = o0; base < o1; base += PT_CTA) {
        const int k = base + tid;
        const int cnt = min(PT_CTA, o1 - base);
        {
            const double2 *wg = reinterpret_cast<const double2 *>(W + (size_t)base * 18);
            double2 *ws = reinterpret_cast<double2 *>(stage);
            double2 wv[9];                                   // all nine 16-byte pieces in flight at once
#pragma unroll
            for (int q = 0; q < 9; ++q) { const int p = tid + q * PT_CTA; wv[q] = p < cnt * 9 ? __ldg(wg + p) : make_double2(0.0, 0.0); }
#pragma unroll
            for (int q = 0; q < 9; ++q) { const int p = tid + q * PT_CTA; if (p < cnt * 9) ws[p] = wv[q]; }
        }
        const int ji = k < o1 ? __ldg(jidx + k) : 0;
        if (EVAL && single) {
            sj[tid] = ji;
            if (k < o1) { lp1 = __ldg(iidx + k) - p0; mm1 = __ldg(reinterpret_cast<const double2 *>(impts) + k); }
        }
        double d[6];
        if (k < o1) {
            const double2 *dq = reinterpret_cast<const double2 *>(dpa + ji * 6);
#pragma unroll
            for (int q = 0; q < 3; ++q) { double2 d2 = __ldg(dq + q); d[2 * q] = d2.x; d[2 * q + 1] = d2.y; }
        }
        __syncthreads();
        double2 pv[6];                                       // 6 x 16 B = q, t, K of the candidate camera
        if (EVAL && single) {
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int p = tid + q * PT_CTA, ob = p / 6, piece = p - ob * 6;
                pv[q] = p < cnt * 6 ? __ldg(reinterpret_cast<const double2 *>(newcache + (size_t)sj[ob] * CAMC) + piece) : make_double2(0.0, 0.0);
            }
        }
        if (k < o1) {
            const double2 *wp = reinterpret_cast<const double2 *>(stage + tid * 18);
            double w[18];
#pragma unroll
            for (int q = 0; q < 9; ++q) { double2 w2 = wp[q]; w[2 * q] = w2.x; w[2 * q + 1] = w2.y; }
            double t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) { t0 += w[r * 3] * d[r]; t1 += w[r * 3 + 1] * d[r]; t2 += w[r * 3 + 2] * d[r]; }
            sh[0][tid] = t0; sh[1][tid] = t1; sh[2][tid] = t2;
        }
        __syncthreads();                                     // the W tile has been consumed
        if (EVAL && single) {
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int p = tid + q * PT_CTA, ob = p / 6, piece = p - ob * 6;
                if (p < cnt * 6) *reinterpret_cast<double2 *>(pstage + ob * PROJ_LD + piece * 2) = pv[q];
            }
        }
        if (tid < np) {
            const int a = max(my_a, base), b = min(my_b, base + PT_CTA);
            for (int q = a; q < b; ++q) { acc0 += sh[0][q - base]; acc1 += sh[1][q - base]; acc2 += sh[2][q - base]; }
        }
        __syncthreads();
    }
    double s_dp2 = 0.0, s_dpg = 0.0, s_e2 = 0.0, s_p2 = 0.0;
    if (tid < np) {
        const int p = p0 + tid;
        const double e0 = g0 - acc0, e1 = g1 - acc1, e2 = g2 - acc2;
        const double d0 = i00 * e0 + i10 * e1 + i20 * e2;
        const double d1 = i10 * e0 + i11 * e1 + i21 * e2;
        const double d2 = i20 * e0 + i21 * e1 + i22 * e2;
        eb[(size_t)p * 3] = e0; eb[(size_t)p * 3 + 1] = e1; eb[(size_t)p * 3 + 2] = e2;
        dpb[(size_t)p * 3] = d0; dpb[(size_t)p * 3 + 1] = d1; dpb[(size_t)p * 3 + 2] = d2;
        if (EVAL) {
            const double x0 = px + d0, x1 = py + d1, x2 = pz + d2;
            newpts[(size_t)p * 3] = x0; newpts[(size_t)p * 3 + 1] = x1; newpts[(size_t)p * 3 + 2] = x2;
            shx[0][tid] = x0; shx[1][tid] = x1; shx[2][tid] = x2;
            s_dp2 = d0 * d0 + d1 * d1 + d2 * d2;
            s_p2 = x0 * x0 + x1 * x1 + x2 * x2;
            s_dpg = d0 * (mu * d0 + g0) + d1 * (mu * d1 + g1) + d2 * (mu * d2 + g2);
        }
    }
    if (!EVAL) return;
    if (single) {
        __syncthreads();                                   // shx and pstage are complete
        const int k = o0 + tid;
        if (k < o1) {
            CamProj cam;
            load_cam_proj<false>(pstage + tid * PROJ_LD, cam);
            double e0, e1;
            if (EXT) residual_ext(cam, ext, jidx[k], k, shx[0][lp1], shx[1][lp1], shx[2][lp1], mm1.x, mm1.y, e0, e1);
            else residual(cam, shx[0][lp1], shx[1][lp1], shx[2][lp1], mm1.x, mm1.y, e0, e1);
            s_e2 += e0 * e0 + e1 * e1;
        }
    } else {
        for (int base = o0; base < o1; base += PT_CTA) {
            const int k = base + tid;
            const int cnt = min(PT_CTA, o1 - base);
            __syncthreads();                               // shx written / previous wave's entries consumed
            sj[tid] = k < o1 ? jidx[k] : 0;
            __syncthreads();
            for (int p = tid; p < cnt * 6; p += PT_CTA) {
                const int ob = p / 6, piece = p - ob * 6;
                const double2 v = __ldg(reinterpret_cast<const double2 *>(newcache + (size_t)sj[ob] * CAMC) + piece);
                *reinterpret_cast<double2 *>(pstage + ob * PROJ_LD + piece * 2) = v;
            }
            __syncthreads();
            if (k < o1) {
                CamProj cam;
                load_cam_proj<false>(pstage + tid * PROJ_LD, cam);
                const int lp = iidx[k] - p0;
                double2 mm = __ldg(reinterpret_cast<const double2 *>(impts) + k);
                double e0, e1;
                if (EXT) residual_ext(cam, ext, sj[tid], k, shx[0][lp], shx[1][lp], shx[2][lp], mm.x, mm.y, e0, e1);
                else residual(cam, shx[0][lp], shx[1][lp], shx[2][lp], mm.x, mm.y, e0, e1);
                s_e2 += e0 * e0 + e1 * e1;
            }
        }
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        s_e2 += __shfl_down_sync(0xffffffffu, s_e2, w);
        s_dp2 += __shfl_down_sync(0xffffffffu, s_dp2, w);
        s_dpg += __shfl_down_sync(0xffffffffu, s_dpg, w);
        s_p2 += __shfl_down_sync(0xffffffffu, s_p2, w);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = s_e2; red[1][tid >> 5] = s_dp2; red[2][tid >> 5] = s_dpg; red[3][tid >> 5] = s_p2; }
    __syncthreads();
    if (tid < 4) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < PT_CTA / 32; ++w) s += red[tid][w];
        part[(size_t)(chunk_list ? (int)blockIdx.x : cidx) * 4 + tid] = s;      // listed launch: partials by CTA
    }
}


// PERSISTENT, software-pipelined variant for the chunks that fit one wave (all of them unless a point has more
// than PT_CTA observations).  A one-shot CTA is a chain of dependent L2 / HBM round trips -- descriptor, then
// indices and per-point data, then the gathers that need a camera index -- and 3-4 resident CTAs per SM cannot
// hide ~5 us of chain behind ~1 us of work.  Here nothing is waited for in the iteration that asks for it:
//   three chunks ahead  the 16-byte chunk descriptor,
//   two chunks ahead    camera index, local point, measurement of every observation (registers),
//   one chunk ahead     the W tile by ONE TMA bulk copy (cp.async.bulk -> mbarrier), the candidate-camera entries
//                       (6 x 16 B per observation) by asynchronous copies (LDGSTS) into a double-buffered stage,
//                       the dpa rows and the per-point data (gb, Vinv, point) in registers.
template <int DUMMY>
#ifndef BSUB_MINB
#define BSUB_MINB (384 / PT_CTA)   // resident CTAs per SM: twelve warps (168 registers per thread)
#endif
__global__ void __launch_bounds__(PT_CTA, BSUB_MINB) k_backsub_pipe(int n_list, const int *__restrict__ chunk_list, const int4 *__restrict__ ptdesc,
                                                           const int *__restrict__ pt_ptr, const int *__restrict__ iidx,
                                                           const int *__restrict__ jidx, const double *__restrict__ impts,
                                                           const double *__restrict__ W, const double *__restrict__ Vinv,
                                                           const double *__restrict__ gb, const double *__restrict__ dpa,
                                                           const double *__restrict__ pts, const double *__restrict__ newcache, const double *__restrict__ mu_p,
                                                           double *__restrict__ eb, double *__restrict__ dpb, double *__restrict__ newpts,
                                                           double *__restrict__ part)
{
    const double mu = __ldg(mu_p);
    extern __shared__ __align__(128) double dyn[];             // two W tiles (TMA destinations), then two entry stages
    double *wtile = dyn, *pstage = dyn + 2 * PT_CTA * 18;
    __shared__ __align__(8) unsigned long long bar[2];
    __shared__ double sh[3][PT_CTA];
    __shared__ double shx[3][PT_CTA];
    __shared__ double red[4][PT_CTA / 32];
    __shared__ int sj[2][PT_CTA];
    const int tid = threadIdx.x, G = gridDim.x;
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    struct idx { int j, lp; double2 mm; };                     // observation: camera, local point, measurement
    struct own { int a, b; double g0, g1, g2, i00, i10, i20, i11, i21, i22, px, py, pz; };   // owner of a point
    auto chunk_of = [&](int q) { return chunk_list ? __ldg(chunk_list + q) : q; };
    const int4 zero4 = make_int4(0, 0, 0, 0);
    auto load_idx = [&](const int4 &d, idx &r) {
        const int k = d.z + tid;
        r.j = 0; r.lp = 0; r.mm = make_double2(0.0, 0.0);
        if (k < d.w) { r.j = __ldg(jidx + k); r.lp = __ldg(iidx + k) - d.x; r.mm = __ldg(reinterpret_cast<const double2 *>(impts) + k); }
    };
    // everything of a chunk that needs only its descriptor and (in sj[buf]) its camera indices
    auto issue = [&](const int4 &d, int buf, int myj, own &r, double (&dq)[6]) {
        const int np = d.y - d.x, cnt = d.w - d.z;
        if (tid == 0) {
            const unsigned bytes = (unsigned)cnt * 144u;
            mbar_expect_tx(&bar[buf], bytes);
            bulk_g2s(wtile + buf * PT_CTA * 18, W + (size_t)d.z * 18, bytes, &bar[buf]);
        }
        double *ps = pstage + buf * PT_CTA * PROJ_LD;
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            const int p = tid + u * PT_CTA, ob = p / 6, piece = p - ob * 6;
            if (p < cnt * 6) cp_async16(ps + ob * PROJ_LD + piece * 2, newcache + (size_t)sj[buf][ob] * CAMC + piece * 2);
        }
        if (tid < cnt) {
            const double2 *dp2 = reinterpret_cast<const double2 *>(dpa + myj * 6);
#pragma unroll
            for (int u = 0; u < 3; ++u) { const double2 d2 = __ldg(dp2 + u); dq[2 * u] = d2.x; dq[2 * u + 1] = d2.y; }
        }
        if (tid < np) {
            const int p = d.x + tid;
            r.a = __ldg(pt_ptr + p); r.b = __ldg(pt_ptr + p + 1);
            r.g0 = __ldg(gb + (size_t)p * 3); r.g1 = __ldg(gb + (size_t)p * 3 + 1); r.g2 = __ldg(gb + (size_t)p * 3 + 2);
            const double2 *vi = reinterpret_cast<const double2 *>(Vinv + (size_t)p * 6);
            const double2 v01 = __ldg(vi), v23 = __ldg(vi + 1), v45 = __ldg(vi + 2);
            r.i00 = v01.x; r.i10 = v01.y; r.i20 = v23.x; r.i11 = v23.y; r.i21 = v45.x; r.i22 = v45.y;
            r.px = __ldg(pts + (size_t)p * 3); r.py = __ldg(pts + (size_t)p * 3 + 1); r.pz = __ldg(pts + (size_t)p * 3 + 2);
        }
    };
    int q = blockIdx.x;
    if (q >= n_list) return;
    int4 ds = __ldg(ptdesc + chunk_of(q));
    int4 ds1 = q + G < n_list ? __ldg(ptdesc + chunk_of(q + G)) : zero4;
    int4 ds2 = q + 2 * G < n_list ? __ldg(ptdesc + chunk_of(q + 2 * G)) : zero4;
    idx cur, mid, far;
    own oc, on;
    double d[6], dn[6];
    load_idx(ds, cur);
    load_idx(ds1, mid);
    sj[0][tid] = cur.j;
    __syncthreads();                                           // barriers initialised, sj[0] visible
    issue(ds, 0, cur.j, on, dn);
    cp_async_commit();
    double cta_sum = 0.0;
    for (int it = 0; q < n_list; q += G, ++it) {
        const int st = it & 1;
        oc = on;
#pragma unroll
        for (int u = 0; u < 6; ++u) d[u] = dn[u];
        sj[st ^ 1][tid] = mid.j;
        __syncthreads();                                       // the other stage half is free, its camera indices are visible
        if (q + G < n_list) issue(ds1, st ^ 1, mid.j, on, dn);
        cp_async_commit();
        load_idx(ds2, far);
        const int4 ds3 = q + 3 * G < n_list ? __ldg(ptdesc + chunk_of(q + 3 * G)) : zero4;
        const int p0 = ds.x, o0 = ds.z, o1 = ds.w, np = ds.y - ds.x;
        const int k = o0 + tid;
        const double *wt = wtile + st * PT_CTA * 18;
        cp_async_wait<1>();
        mbar_wait(&bar[st], (it >> 1) & 1);                    // the TMA bytes of this chunk have landed
        if (k < o1) {
            const double2 *wp = reinterpret_cast<const double2 *>(wt + tid * 18);
            double w[18];
#pragma unroll
            for (int u = 0; u < 9; ++u) { double2 w2 = wp[u]; w[2 * u] = w2.x; w[2 * u + 1] = w2.y; }
            double t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) { t0 += w[r * 3] * d[r]; t1 += w[r * 3 + 1] * d[r]; t2 += w[r * 3 + 2] * d[r]; }
            sh[0][tid] = t0; sh[1][tid] = t1; sh[2][tid] = t2;
        }
        __syncthreads();                                       // sh complete; every thread's entry copies have landed
        double s_dp2 = 0.0, s_dpg = 0.0, s_e2 = 0.0, s_p2 = 0.0;
        if (tid < np) {
            double acc0 = 0, acc1 = 0, acc2 = 0;
            for (int u = oc.a; u < oc.b; ++u) { acc0 += sh[0][u - o0]; acc1 += sh[1][u - o0]; acc2 += sh[2][u - o0]; }
            const int p = p0 + tid;
            const double e0 = oc.g0 - acc0, e1 = oc.g1 - acc1, e2 = oc.g2 - acc2;
            const double d0 = oc.i00 * e0 + oc.i10 * e1 + oc.i20 * e2;
            const double d1 = oc.i10 * e0 + oc.i11 * e1 + oc.i21 * e2;
            const double d2 = oc.i20 * e0 + oc.i21 * e1 + oc.i22 * e2;
            eb[(size_t)p * 3] = e0; eb[(size_t)p * 3 + 1] = e1; eb[(size_t)p * 3 + 2] = e2;
            dpb[(size_t)p * 3] = d0; dpb[(size_t)p * 3 + 1] = d1; dpb[(size_t)p * 3 + 2] = d2;
            const double x0 = oc.px + d0, x1 = oc.py + d1, x2 = oc.pz + d2;
            newpts[(size_t)p * 3] = x0; newpts[(size_t)p * 3 + 1] = x1; newpts[(size_t)p * 3 + 2] = x2;
            shx[0][tid] = x0; shx[1][tid] = x1; shx[2][tid] = x2;
            s_dp2 = d0 * d0 + d1 * d1 + d2 * d2;
            s_p2 = x0 * x0 + x1 * x1 + x2 * x2;
            s_dpg = d0 * (mu * d0 + oc.g0) + d1 * (mu * d1 + oc.g1) + d2 * (mu * d2 + oc.g2);
        }
        __syncthreads();                                       // shx complete
        if (k < o1) {
            CamProj cam;
            load_cam_proj<false>(pstage + st * PT_CTA * PROJ_LD + tid * PROJ_LD, cam);
            double e0, e1;
            residual(cam, shx[0][cur.lp], shx[1][cur.lp], shx[2][cur.lp], cur.mm.x, cur.mm.y, e0, e1);
            s_e2 = e0 * e0 + e1 * e1;
        }
#pragma unroll
        for (int w = 16; w > 0; w >>= 1) {
            s_e2 += __shfl_down_sync(0xffffffffu, s_e2, w);
            s_dp2 += __shfl_down_sync(0xffffffffu, s_dp2, w);
            s_dpg += __shfl_down_sync(0xffffffffu, s_dpg, w);
            s_p2 += __shfl_down_sync(0xffffffffu, s_p2, w);
        }
        if ((tid & 31) == 0) { red[0][tid >> 5] = s_e2; red[1][tid >> 5] = s_dp2; red[2][tid >> 5] = s_dpg; red[3][tid >> 5] = s_p2; }
        __syncthreads();
        if (tid < 4) {                                         // the CTA's own running sums (its chunks in order): one partial per CTA
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < PT_CTA / 32; ++w) s += red[tid][w];
            cta_sum += s;
        }
        cur = mid; mid = far;
        ds = ds1; ds1 = ds2; ds2 = ds3;
    }
    cp_async_wait<0>();
    if (tid < 4) part[(size_t)blockIdx.x * 4 + tid] = cta_sum;
}

// out[v] = sum_p part[p*4+v], v<4, fixed order.  1024 threads, four independent partial sums per thread and
// value so that the loads of one CTA are in flight together (it is a single-CTA kernel on the critical path
// of every try: 44 us with 256 threads and one dependent chain)
__global__ void __launch_bounds__(1024) k_final_reduce4(const double *__restrict__ part, int nparts, double *__restrict__ out)
{
    __shared__ double sh[4][1024];
    double s[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) s[u][v] = 0.0;
    const double4 *p4 = reinterpret_cast<const double4 *>(part);
    int p = threadIdx.x;
    for (; p + 3 * 1024 < nparts; p += 4 * 1024) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double4 q = p4[p + u * 1024];
            s[u][0] += q.x; s[u][1] += q.y; s[u][2] += q.z; s[u][3] += q.w;
        }
    }
    for (; p < nparts; p += 1024) { const double4 q = p4[p]; s[0][0] += q.x; s[0][1] += q.y; s[0][2] += q.z; s[0][3] += q.w; }
#pragma unroll
    for (int v = 0; v < 4; ++v) sh[v][threadIdx.x] = (s[0][v] + s[1][v]) + (s[2][v] + s[3][v]);
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
#pragma unroll
            for (int v = 0; v < 4; ++v) sh[v][threadIdx.x] += sh[v][threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) out[threadIdx.x] = sh[threadIdx.x][0];
}

// evaluate=true : dp (cams+points), candidate parameters, candidate cost, step scalars (LM try)
// evaluate=false: eb and dpb only (trust region / compat)
// everything of the back-substitution up to (and with) the copies of the step scalars and the status word to the host:
// no host round trip (the caller synchronises; psba_finish_try reads)
void psba_enqueue_backsub(psba_ctx *c, double mu_value, bool evaluate)
{
    psba_set_scalars(c, mu_value, 0.0, 0.0);
    const double *mu = c->d_mu;
    const int cur = c->cur, nw = 1 - cur;
    double *gb = c->g + c->N, *ebp = c->eab + c->N, *dpbp = c->dp + c->N;
    if (evaluate) {
        PROF(c, KID_NEWCAMS) k_newcams<<<1, 1024, 0, c->stream>>>(c->N, c->cams[cur], c->dp, c->g, mu, c->cams[nw], c->d_scal + 4);
        psba_launch_cam_prep(c, nw);
        int n_part = 0;
        PROF(c, KID_BACKSUB) {
            const int dyn = 2 * PT_CTA * (18 + PROJ_LD) * (int)sizeof(double);
            psba_set_smem((const void *)k_backsub_pipe<0>, dyn);
            int gs = std::min(c->n_small, c->n_sm * BSUB_MINB);      // persistent CTAs: one partial each
            if (c->ext_on) {       // extended camera model: the one-shot kernel evaluates the candidate (one partial per chunk)
                gs = c->n_ptchunk;
                if (gs > 0)
                    k_backsub<true, true><<<gs, PT_CTA, 0, c->stream>>>(nullptr, c->ptdesc, c->pt_ptr, c->iidx, c->jidx, c->impts, c->W, c->Vinv, gb, c->dp,
                                                                       c->pts[cur], c->camcache[nw], mu, ebp, dpbp, c->pts[nw], c->d_part, c->ext);
            } else if (c->n_small > 0)
                k_backsub_pipe<0><<<gs, PT_CTA, dyn, c->stream>>>(c->n_small, c->d_small_list, c->ptdesc, c->pt_ptr, c->iidx,
                                                              c->jidx, c->impts, c->W, c->Vinv, gb, c->dp, c->pts[cur],
                                                              c->camcache[nw], mu, ebp, dpbp, c->pts[nw], c->d_part);
            if (c->n_big > 0 && !c->ext_on)      // points with more observations than one wave: one partial per CTA behind the others
                k_backsub<true, false><<<c->n_big, PT_CTA, 0, c->stream>>>(c->d_big_list, c->ptdesc, c->pt_ptr, c->iidx, c->jidx, c->impts, c->W, c->Vinv,
                                                                   gb, c->dp, c->pts[cur], c->camcache[nw], mu, ebp, dpbp, c->pts[nw], c->d_part + (size_t)gs * 4, c->ext);
            n_part = c->ext_on ? gs : gs + c->n_big;
        }
        PROF(c, KID_REDUCE) k_final_reduce4<<<1, 1024, 0, c->stream>>>(c->d_part, n_part, c->d_scal);
        c->st_launches += 3; c->st_exqt += 1;
        LAUNCH_CHECK();
        if (c->nranks > 1) psba_allreduce_sum(c, c->d_scal, 4);
        CUDA_CHECK(cudaMemcpyAsync(c->h_scal, c->d_scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    } else {
        if (c->n_ptchunk > 0)
            PROF(c, KID_BACKSUB) k_backsub<false, false><<<c->n_ptchunk, PT_CTA, 0, c->stream>>>(nullptr, c->ptdesc, c->pt_ptr, c->iidx, c->jidx, c->impts, c->W, c->Vinv,
                                                                    gb, c->dp, c->pts[cur], c->camcache[cur], mu, ebp, dpbp,
                                                                    c->pts[nw], c->d_part, c->ext);
        c->st_launches += 1;
        LAUNCH_CHECK();
    }
}

void psba_finish_try(psba_ctx *c, psba_try_result *res)
{
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (res) {
        res->cost_new = c->h_scal[0];
        res->dp_L2 = c->h_scal[4] + c->h_scal[1];      // cameras first, then points (misc.cpp:151-157 order)
        res->dp_dot = c->h_scal[5] + c->h_scal[2];
        res->p_new_L2 = c->h_scal[6] + c->h_scal[3];   // ||candidate parameters||^2
    }
}

void psba_launch_backsub(psba_ctx *c, double mu, bool evaluate, psba_try_result *res)
{
    psba_enqueue_backsub(c, mu, evaluate);
    if (evaluate) psba_finish_try(c, res);
}

// new = cur + dp for an externally supplied step in c->dp (compute_newp.cl:6-26)
__global__ void k_newp(int n, const double *__restrict__ a, const double *__restrict__ d, double *__restrict__ out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = a[k] + d[k];
}

// dp = a x + b y over [N | 3n] and the candidate parameters p + dp in the same pass (trust-region radius try: one launch
// instead of k_axpby + two k_newp); the products are formed exactly as k_axpby and k_newp form them
__global__ void k_step_newp(int N, int n3, const double *__restrict__ coef, const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ dp,
                            const double *__restrict__ cams, const double *__restrict__ pts, double *__restrict__ cams_new, double *__restrict__ pts_new)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N + n3) return;
    const double a = __ldg(coef + 1), b = __ldg(coef + 2);       // device-resident step coefficients (psba_set_scalars)
    const double d = a * x[k] + b * y[k];
    dp[k] = d;
    if (k < N) cams_new[k] = cams[k] + d; else pts_new[k - N] = pts[k - N] + d;
}

void psba_launch_step_newp(psba_ctx *c, double a, const double *x, double b, const double *y)
{
    const int cur = c->cur, nw = 1 - cur, tot = c->N + 3 * c->n;
    psba_set_scalars(c, 0.0, a, b);
    PROF(c, KID_VEC) k_step_newp<<<cdiv(tot, 256), 256, 0, c->stream>>>(c->N, 3 * c->n, c->d_mu, x, y, c->dp, c->cams[cur], c->pts[cur], c->cams[nw], c->pts[nw]);
    c->cache_valid[nw] = false;
    c->st_launches += 1;
    LAUNCH_CHECK();
}

void psba_launch_newp(psba_ctx *c)
{
    const int cur = c->cur, nw = 1 - cur;
    k_newp<<<cdiv(c->N, 256), 256, 0, c->stream>>>(c->N, c->cams[cur], c->dp, c->cams[nw]);
    if (c->n > 0) k_newp<<<cdiv(3 * c->n, 256), 256, 0, c->stream>>>(3 * c->n, c->pts[cur], c->dp + c->N, c->pts[nw]);
    c->cache_valid[nw] = false;
    c->st_launches += 2;
    LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------
// all pairwise dot products of up to three [N | 3n] vectors: out = {xx, xy, xz, yy, yz, zz}.
// Camera part (replicated across ranks) and point part (sharded) are reduced separately.
// camera part (block 0: the first N entries) and point part (the other blocks) of the three vectors in ONE launch;
// partials: part[0..6) camera, part[8 + 6 b ..) point block b
__global__ void __launch_bounds__(256) k_dots2(int N, int np, const double *__restrict__ x, const double *__restrict__ y,
                                               const double *__restrict__ z, double *__restrict__ part)
{
    __shared__ double sh[16 * (256 + 4)];
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (blockIdx.x == 0) {
        for (int k = threadIdx.x; k < N; k += 256) {
            const double a = x[k], b = y[k], cc = z[k];
            v[0] += a * a; v[1] += a * b; v[2] += a * cc; v[3] += b * b; v[4] += b * cc; v[5] += cc * cc;
        }
        block_reduce_to<6, 256, 16>(v, sh, part);
    } else {
        const int nb = gridDim.x - 1;
        for (int k = (blockIdx.x - 1) * 256 + threadIdx.x; k < np; k += nb * 256) {
            const double a = x[N + k], b = y[N + k], cc = z[N + k];
            v[0] += a * a; v[1] += a * b; v[2] += a * cc; v[3] += b * b; v[4] += b * cc; v[5] += cc * cc;
        }
        block_reduce_to<6, 256, 16>(v, sh, part + 8 + (size_t)(blockIdx.x - 1) * 6);
    }
}
// out[0..6) = camera partial, out[6..12) = fixed-order sum of the point partials
__global__ void k_final_reduce6x2(const double *__restrict__ part, int nparts, double *__restrict__ out)
{
    const int v = threadIdx.x;
    if (v < 6) out[v] = part[v];
    else if (v < 12) {
        double s = 0.0;
#pragma unroll 8
        for (int p = 0; p < nparts; ++p) s += part[8 + (size_t)p * 6 + (v - 6)];
        out[v] = s;
    }
}

void psba_enqueue_dots(psba_ctx *c, const double *x, const double *y, const double *z, int off)
{
    const int np = 3 * c->n;
    const int nb = np > 0 ? std::min(cdiv(np, 256), 296) : 0;
    k_dots2<<<1 + nb, 256, 0, c->stream>>>(c->N, np, x, y, z, c->d_part);
    k_final_reduce6x2<<<1, 32, 0, c->stream>>>(c->d_part, nb, c->d_scal + off);
    c->st_launches += 2;
    LAUNCH_CHECK();
    if (c->nranks > 1) psba_allreduce_sum(c, c->d_scal + off + 6, 6);
}

void psba_launch_dots(psba_ctx *c, const double *x, const double *y, const double *z, double out[6])
{
    psba_enqueue_dots(c, x, y, z, 0);
    CUDA_CHECK(cudaMemcpyAsync(c->h_scal, c->d_scal, 12 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (int v = 0; v < 6; ++v) out[v] = c->h_scal[v] + c->h_scal[6 + v];
}

__global__ void k_axpby(int n, double a, const double *__restrict__ x, double b, const double *__restrict__ y, double *__restrict__ out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = a * x[k] + b * y[k];
}

void psba_launch_axpby(psba_ctx *c, double a, const double *x, double b, const double *y, double *out)
{
    const int n = c->N + 3 * c->n;
    PROF(c, KID_VEC) k_axpby<<<cdiv(n, 256), 256, 0, c->stream>>>(n, a, x, b, y, out);
    c->st_launches += 1;
    LAUNCH_CHECK();
}

// max over the diagonals of U and V (maxElmOfUV, sba_func.cpp:422-444; only positive values count)
__global__ void k_maxdiag(int m, int n, const double *__restrict__ U, const double *__restrict__ V, double *__restrict__ part)
{
    __shared__ double sh[256];
    double mx = 0.0;
    const int tot = 6 * m + 3 * n;
    for (int k = blockIdx.x * 256 + threadIdx.x; k < tot; k += gridDim.x * 256) {
        double v;
        if (k < 6 * m) { const int j = k / 6, r = k - j * 6; v = U[j * 36 + r * 7]; }
        else { const int q = k - 6 * m, i = q / 3, r = q - i * 3; v = V[(size_t)i * 6 + (r == 0 ? 0 : (r == 1 ? 3 : 5))]; }
        if (v > mx) mx = v;
    }
    sh[threadIdx.x] = mx;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

__global__ void k_final_max(const double *__restrict__ part, int nparts, double *__restrict__ out)
{
    double mx = 0.0;
    for (int p = 0; p < nparts; ++p) mx = fmax(mx, part[p]);
    out[0] = mx;
}

double psba_launch_maxdiag(psba_ctx *c)
{
    const int tot = c->N + 3 * c->n;
    const int nb = std::min(cdiv(tot, 256), 296);
    k_maxdiag<<<nb, 256, 0, c->stream>>>(c->m, c->n, c->U, c->V, c->d_part);
    k_final_max<<<1, 1, 0, c->stream>>>(c->d_part, nb, c->d_scal);
    c->st_launches += 2;
    LAUNCH_CHECK();
    if (c->nranks > 1) psba_allreduce_max(c, c->d_scal, 1);
    CUDA_CHECK(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return c->h_scal[0];
}
