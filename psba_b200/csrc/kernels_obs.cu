// kernels_obs.cu -- per-observation kernels: camera cache, residual / cost, fused linearisation
// (point-major pass: W, V, gb; camera-major pass: U, ga), materialised Jacobian (compat only)
// and J*x products with fused dot products (trust region).
//
// Replaces kern_compute_exQT (CL_files/compute_exQT.cl:18-71), kern_compute_jacobiQT
// (compute_jacobiQT.cl:7-141), kern_compute_U/V/Wblks/g (compute_U.cl:5-35, compute_V.cl:6-38,
// compute_Wblks.cl:7-34, compute_g.cl:6-60) and kern_compute_Jmultiply (compute_Jmultiply.cl:6-52).
// The Jacobian never reaches HBM: both passes recompute it from a 192-byte cache entry per camera.
// All cross-thread sums have a fixed order (no atomics) => bit-reproducible run to run.
//
// Memory access: per-observation gathers (camera entry 192 B, W block 144 B) are turned into
// coalesced 16-byte-per-lane transfers staged through shared memory: a warp moves 512 contiguous
// bytes per instruction instead of 32 scattered sectors (ncu r01: 23 sectors/request before).
#include "dev_math.cuh"
#include <algorithm>

#define CAM_LD 26          // doubles per staged camera entry in shared memory (24 + pad: conflict-free LDS.128)
#define CAM_LD2 18         // pipelined point pass: compact records (16 doubles + pad; the W tile of the chunk, 18 doubles per observation, takes the place)

// ------------------------------------------------------------------------------------------------
// per-camera cache: q = ql (x) q0, t, K, D_k = d q / d v_k
__global__ void k_cam_prep(int m, const double *__restrict__ K, const double *__restrict__ initcams,
                           const double *__restrict__ cams, double *__restrict__ cache)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double s0 = initcams[j * 4], a1 = initcams[j * 4 + 1], a2 = initcams[j * 4 + 2], a3 = initcams[j * 4 + 3];
    const double v1 = cams[j * 6], v2 = cams[j * 6 + 1], v3 = cams[j * 6 + 2];
    const double sl = sqrt(1 - v1 * v1 - v2 * v2 - v3 * v3);
    double *c = cache + (size_t)j * CAMC;
    {   // compact record behind the m full entries (load_cam_compact, dev_math.cuh)
        double *r = cache + (size_t)m * CAMC + (size_t)j * CAMC2;
        r[0] = s0; r[1] = a1; r[2] = a2; r[3] = a3; r[4] = -v1 / sl; r[5] = -v2 / sl; r[6] = -v3 / sl; r[7] = sl;
        r[8] = cams[j * 6 + 3]; r[9] = cams[j * 6 + 4]; r[10] = cams[j * 6 + 5];
#pragma unroll
        for (int k = 0; k < 5; ++k) r[11 + k] = K[j * 5 + k];
    }
    // q = ql (x) q0, same operation order as compute_exQT.cl:46-49
    c[CC_Q + 0] = sl * s0 - (a1 * v1 + a2 * v2 + a3 * v3);
    c[CC_Q + 1] = s0 * v1 + sl * a1 + a3 * v2 - a2 * v3;
    c[CC_Q + 2] = s0 * v2 + sl * a2 + a1 * v3 - a3 * v1;
    c[CC_Q + 3] = s0 * v3 + sl * a3 + a2 * v1 - a1 * v2;
    c[CC_T + 0] = cams[j * 6 + 3]; c[CC_T + 1] = cams[j * 6 + 4]; c[CC_T + 2] = cams[j * 6 + 5];
#pragma unroll
    for (int k = 0; k < 5; ++k) c[CC_K + k] = K[j * 5 + k];
    const double vv[3] = {v1, v2, v3};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // d ql / d v_k = (-v_k/sl, e_k);  D_k = dql (x) q0
        const double ds_l = -vv[k] / sl;
        const double e1 = (k == 0), e2 = (k == 1), e3 = (k == 2);
        c[CC_D + 4 * k + 0] = ds_l * s0 - (a1 * e1 + a2 * e2 + a3 * e3);
        c[CC_D + 4 * k + 1] = s0 * e1 + ds_l * a1 + a3 * e2 - a2 * e3;
        c[CC_D + 4 * k + 2] = s0 * e2 + ds_l * a2 + a1 * e3 - a3 * e1;
        c[CC_D + 4 * k + 3] = s0 * e3 + ds_l * a3 + a2 * e1 - a1 * e2;
    }
}

void psba_launch_cam_prep(psba_ctx *c, int set)
{
    PROF(c, KID_CAM_PREP) k_cam_prep<<<cdiv(c->m, 128), 128, 0, c->stream>>>(c->m, c->K, c->initcams, c->cams[set], c->camcache[set]);
    c->cache_valid[set] = true;
    c->st_launches += 1;
    LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
// final deterministic reduction of per-CTA partials: out[v] = sum_p part[p*stride + v]
__global__ void k_final_reduce(const double *__restrict__ part, int nparts, int stride, int nv, double *__restrict__ out)
{
    __shared__ double sh[256];
    for (int v = 0; v < nv; ++v) {
        double s = 0.0;
        for (int p = threadIdx.x; p < nparts; p += 256) s += part[(size_t)p * stride + v];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[v] = sh[0];
        __syncthreads();
    }
}

// residual + cost. One thread per observation (coalesced impts / idx loads, cached gathers).
template <bool EXT>
__global__ void __launch_bounds__(256) k_cost(int o, const int *__restrict__ iidx, const int *__restrict__ jidx,
                                              const double *__restrict__ impts, const double *__restrict__ cache,
                                              const double *__restrict__ pts, double *__restrict__ ex,
                                              double *__restrict__ part, psba_ext ext)
{
    __shared__ double sh[8];
    int k = blockIdx.x * 256 + threadIdx.x;
    double e2 = 0.0;
    if (k < o) {
        CamProj cam;
        load_cam_proj<true>(cache + (size_t)jidx[k] * CAMC, cam);
        const double *X = pts + (size_t)iidx[k] * 3;
        double2 mm = __ldg(reinterpret_cast<const double2 *>(impts) + k);
        double e0, e1;
        if (EXT) residual_ext(cam, ext, jidx[k], k, __ldg(X), __ldg(X + 1), __ldg(X + 2), mm.x, mm.y, e0, e1);
        else residual(cam, __ldg(X), __ldg(X + 1), __ldg(X + 2), mm.x, mm.y, e0, e1);
        if (ex) reinterpret_cast<double2 *>(ex)[k] = make_double2(e0, e1);
        e2 = e0 * e0 + e1 * e1;
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) e2 += __shfl_down_sync(0xffffffffu, e2, w);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = e2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sh[w];
        part[blockIdx.x] = s;
    }
}

static void read_scalars(psba_ctx *c, int off, int n)
{
    CUDA_CHECK(cudaMemcpyAsync(c->h_scal + off, c->d_scal + off, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

double psba_enqueue_cost(psba_ctx *c, int set, double *ex_dev)
{
    if (!c->cache_valid[set]) psba_launch_cam_prep(c, set);
    int nb = cdiv(c->o, 256);
    if (nb > 0)
        PROF(c, KID_COST) {
            if (c->ext_on) k_cost<true><<<nb, 256, 0, c->stream>>>(c->o, c->iidx, c->jidx, c->impts, c->camcache[set], c->pts[set], ex_dev, c->d_part, c->ext);
            else k_cost<false><<<nb, 256, 0, c->stream>>>(c->o, c->iidx, c->jidx, c->impts, c->camcache[set], c->pts[set], ex_dev, c->d_part, c->ext);
        }
    PROF(c, KID_REDUCE) k_final_reduce<<<1, 256, 0, c->stream>>>(c->d_part, nb, 1, 1, c->d_scal);
    c->st_launches += 2; c->st_exqt += 1;
    LAUNCH_CHECK();
    if (c->nranks > 1) psba_allreduce_sum(c, c->d_scal, 1);
    return 0.0;
}

double psba_launch_cost(psba_ctx *c, int set, double *ex_dev)
{
    psba_enqueue_cost(c, set, ex_dev);
    read_scalars(c, 0, 1);
    return c->h_scal[0];
}

// ------------------------------------------------------------------------------------------------
// point-major linearisation pass.  CTA (128 threads) = one chunk of whole points (<= 128
// observations per wave).
//   1. the wave's camera entries are copied global -> shared cooperatively (12 x 16 B pieces per
//      entry, consecutive lanes take consecutive pieces: coalesced);
//   2. thread k: observation -> e, A, B in registers, W = c * A^T B into a shared tile;
//      B^T B (6) and B^T e (3) go to shared for the point's owner thread, which sums them in
//      ascending camera order (the order of compute_V.cl:24-31 / compute_g.cl:43-54);
//   3. the W tile (128 x 144 B, contiguous in HBM because observations are point-major) is written
//      with fully coalesced 16-byte stores.
template <bool EXT>
__global__ void __launch_bounds__(PT_CTA, 4) k_lin_points(const int *__restrict__ chunk_list, const int4 *__restrict__ ptdesc, const int *__restrict__ pt_ptr,
                                                         const int *__restrict__ iidx, const int *__restrict__ jidx,
                                                         const double *__restrict__ impts, const double *__restrict__ cache,
                                                         const double *__restrict__ pts, double coeff, double coeff_g,
                                                         double *__restrict__ W, double *__restrict__ V, double *__restrict__ gb, psba_ext ext)
{
    __shared__ __align__(16) double stage[PT_CTA * CAM_LD];      // camera entries, then the W tile (128*18 <= 128*26)
    __shared__ double sh[9][PT_CTA];
    __shared__ int sj[PT_CTA];
    const int tid = threadIdx.x;
    // one 16-byte descriptor per chunk {p0, p1, o0, o1}; every independent load of a wave is issued before
    // anything waits (a CTA is a chain of dependent L2 / HBM round trips: the fewer links, the better)
    const int4 ds = __ldg(ptdesc + (chunk_list ? chunk_list[blockIdx.x] : blockIdx.x));
    const int p0 = ds.x, p1 = ds.y, o0 = ds.z, o1 = ds.w;
    const int np = p1 - p0;
    double acc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = 0.0;
    int my_a = 0, my_b = 0;
    if (tid < np) { my_a = __ldg(pt_ptr + p0 + tid); my_b = __ldg(pt_ptr + p0 + tid + 1); }

    for (int base = o0; base < o1; base += PT_CTA) {
        const int k = base + tid;
        const int cnt = min(PT_CTA, o1 - base);
        int ji = 0, ii = 0;
        double2 mm = make_double2(0.0, 0.0);
        if (k < o1) { ji = __ldg(jidx + k); ii = __ldg(iidx + k); mm = __ldg(reinterpret_cast<const double2 *>(impts) + k); }
        sj[tid] = ji;
        double X0 = 0, X1 = 0, X2 = 0;
        if (k < o1) { const double *X = pts + (size_t)ii * 3; X0 = __ldg(X); X1 = __ldg(X + 1); X2 = __ldg(X + 2); }
        __syncthreads();
        // 1. stage camera entries: piece p = (observation p/12, 16-byte part p%12)
#pragma unroll
        for (int h = 0; h < 2; ++h) {                       // two batches of six loads in flight (register budget)
            double2 cv[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int p = tid + (h * 6 + q) * PT_CTA, ob = p / 12, part = p - ob * 12;
                cv[q] = p < cnt * 12 ? __ldg(reinterpret_cast<const double2 *>(cache + (size_t)sj[ob] * CAMC) + part) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int p = tid + (h * 6 + q) * PT_CTA, ob = p / 12, part = p - ob * 12;
                if (p < cnt * 12) *reinterpret_cast<double2 *>(stage + ob * CAM_LD + part * 2) = cv[q];
            }
        }
        __syncthreads();
        double w[18];
        if (k < o1) {
            CamReg cam;
            load_cam<false>(stage + tid * CAM_LD, cam);
            double e0, e1, A[12], B[6];
            if (EXT) residual_jac_ext(cam, ext, ji, k, X0, X1, X2, mm.x, mm.y, e0, e1, A, B);
            else residual_jac(cam, X0, X1, X2, mm.x, mm.y, e0, e1, A, B);
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) w[r * 3 + cc] = coeff * (A[r] * B[cc] + A[6 + r] * B[3 + cc]);
            sh[0][tid] = B[0] * B[0] + B[3] * B[3];
            sh[1][tid] = B[0] * B[1] + B[3] * B[4];
            sh[2][tid] = B[0] * B[2] + B[3] * B[5];
            sh[3][tid] = B[1] * B[1] + B[4] * B[4];
            sh[4][tid] = B[1] * B[2] + B[4] * B[5];
            sh[5][tid] = B[2] * B[2] + B[5] * B[5];
            sh[6][tid] = B[0] * e0 + B[3] * e1;
            sh[7][tid] = B[1] * e0 + B[4] * e1;
            sh[8][tid] = B[2] * e0 + B[5] * e1;
        }
        __syncthreads();                               // every thread has consumed its camera entry
        if (k < o1) {
            double2 *ws = reinterpret_cast<double2 *>(stage + tid * 18);
#pragma unroll
            for (int q = 0; q < 9; ++q) ws[q] = make_double2(w[2 * q], w[2 * q + 1]);
        }
        if (tid < np) {
            const int a = max(my_a, base), b = min(my_b, base + PT_CTA);
            for (int q = a; q < b; ++q) {
#pragma unroll
                for (int v = 0; v < 9; ++v) acc[v] += sh[v][q - base];
            }
        }
        __syncthreads();
        // 3. coalesced store of the W tile
        {
            double2 *wg = reinterpret_cast<double2 *>(W + (size_t)base * 18);
            const double2 *ws = reinterpret_cast<const double2 *>(stage);
            for (int p = tid; p < cnt * 9; p += PT_CTA) wg[p] = ws[p];
        }
        __syncthreads();
    }
    if (tid < np) {
        double *Vp = V + (size_t)(p0 + tid) * 6;
#pragma unroll
        for (int v = 0; v < 6; ++v) Vp[v] = coeff * acc[v];
        double *gp = gb + (size_t)(p0 + tid) * 3;
        gp[0] = coeff_g * acc[6]; gp[1] = coeff_g * acc[7]; gp[2] = coeff_g * acc[8];
    }
}


// PERSISTENT, software-pipelined variant for the chunks that fit one wave (all of them unless a point has more
// than PT_CTA observations).  Nothing a chunk needs is waited for in the iteration that asks for it:
//   three chunks ahead  the 16-byte chunk descriptor,
//   two chunks ahead    indices, measurements, point ranges (registers),
//   one chunk ahead     camera entries (12 x 16 B per observation) and the chunk's points as asynchronous copies
//                       (LDGSTS) into the other half of a double-buffered stage,
// and the W tile (contiguous in HBM: observations are point-major) leaves with ONE bulk store (TMA) out of the
// stage half its camera entries came in, instead of a copy-out loop through the load/store unit.
template <int DUMMY>
#ifndef LINP_MINB
#define LINP_MINB (512 / PT_CTA)   // resident CTAs per SM: sixteen warps (128 registers per thread, 26 KB of shared memory per CTA)
#endif
__global__ void __launch_bounds__(PT_CTA, LINP_MINB) k_lin_points_pipe(int n_list, const int *__restrict__ chunk_list, const int4 *__restrict__ ptdesc,
                                                              const int *__restrict__ pt_ptr, const int *__restrict__ iidx,
                                                              const int *__restrict__ jidx, const double *__restrict__ impts,
                                                              const double *__restrict__ cache, const double *__restrict__ pts,
                                                              double coeff, double coeff_g, double *__restrict__ W,
                                                              double *__restrict__ V, double *__restrict__ gb)
{
    extern __shared__ __align__(128) double stage_dyn[];       // 2 x PT_CTA*CAM_LD2: compact camera records, then the W tile
    __shared__ double sh[9][PT_CTA];
    __shared__ double px[2][3][PT_CTA];
    __shared__ int sj[2][PT_CTA];
    const int tid = threadIdx.x, G = gridDim.x;
    struct idx { int j, lp, a, b; double2 mm; };
    auto chunk_of = [&](int q) { return chunk_list ? __ldg(chunk_list + q) : q; };
    const int4 zero4 = make_int4(0, 0, 0, 0);
    auto load_idx = [&](const int4 &d, idx &r) {
        const int k = d.z + tid;
        r.j = 0; r.lp = 0; r.a = 0; r.b = 0; r.mm = make_double2(0.0, 0.0);
        if (k < d.w) { r.j = __ldg(jidx + k); r.lp = __ldg(iidx + k) - d.x; r.mm = __ldg(reinterpret_cast<const double2 *>(impts) + k); }
        if (tid < d.y - d.x) { r.a = __ldg(pt_ptr + d.x + tid); r.b = __ldg(pt_ptr + d.x + tid + 1); }
    };
    auto issue = [&](const int4 &d, int buf) {               // camera entries (sj[buf] is visible) and points of a chunk
        const int cnt = d.w - d.z, np = d.y - d.x;
        double *st = stage_dyn + buf * PT_CTA * CAM_LD2;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int p = tid + q * PT_CTA, ob = p >> 3, part = p & 7;
            if (p < cnt * 8) cp_async16(st + ob * CAM_LD2 + part * 2, cache + (size_t)sj[buf][ob] * CAMC2 + part * 2);
        }
        if (tid < np) {
            const double *X = pts + (size_t)(d.x + tid) * 3;
            cp_async8(&px[buf][0][tid], X); cp_async8(&px[buf][1][tid], X + 1); cp_async8(&px[buf][2][tid], X + 2);
        }
    };
    int q = blockIdx.x;
    if (q >= n_list) return;
    int4 ds = __ldg(ptdesc + chunk_of(q));
    int4 ds1 = q + G < n_list ? __ldg(ptdesc + chunk_of(q + G)) : zero4;
    int4 ds2 = q + 2 * G < n_list ? __ldg(ptdesc + chunk_of(q + 2 * G)) : zero4;
    idx cur, mid, far;
    load_idx(ds, cur);
    load_idx(ds1, mid);
    sj[0][tid] = cur.j;
    __syncthreads();
    issue(ds, 0);
    cp_async_commit();
    for (int it = 0; q < n_list; q += G, ++it) {
        const int buf = it & 1;
        // 1. the store of the previous chunk has read its tile: that stage half takes the next chunk's entries
        if (tid == 0) bulk_wait_read0();
        sj[buf ^ 1][tid] = mid.j;
        __syncthreads();
        if (q + G < n_list) issue(ds1, buf ^ 1);
        cp_async_commit();
        // 2. two chunks ahead: indices; three ahead: the descriptor
        load_idx(ds2, far);
        const int4 ds3 = q + 3 * G < n_list ? __ldg(ptdesc + chunk_of(q + 3 * G)) : zero4;
        cp_async_wait<1>();
        __syncthreads();                                     // entries and points of this chunk have landed
        const int p0 = ds.x, o0 = ds.z, o1 = ds.w, np = ds.y - ds.x, cnt = o1 - o0;
        const int k = o0 + tid;
        double *st = stage_dyn + buf * PT_CTA * CAM_LD2;
        CamReg cam;
        load_cam_compact(st + tid * CAM_LD2, cam);
        const double X0 = px[buf][0][cur.lp], X1 = px[buf][1][cur.lp], X2 = px[buf][2][cur.lp];
        __syncthreads();                                     // every thread holds its entry: the half is free for the W tile
        if (k < o1) {
            double e0, e1, A[12], B[6];
            residual_jac(cam, X0, X1, X2, cur.mm.x, cur.mm.y, e0, e1, A, B);
            double2 *ws = reinterpret_cast<double2 *>(st + tid * 18);
#pragma unroll
            for (int r = 0; r < 6; r += 2) {
                double w[6];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) w[h * 3 + cc] = coeff * (A[r + h] * B[cc] + A[6 + r + h] * B[3 + cc]);
                ws[r / 2 * 3] = make_double2(w[0], w[1]); ws[r / 2 * 3 + 1] = make_double2(w[2], w[3]); ws[r / 2 * 3 + 2] = make_double2(w[4], w[5]);
            }
            sh[0][tid] = B[0] * B[0] + B[3] * B[3];
            sh[1][tid] = B[0] * B[1] + B[3] * B[4];
            sh[2][tid] = B[0] * B[2] + B[3] * B[5];
            sh[3][tid] = B[1] * B[1] + B[4] * B[4];
            sh[4][tid] = B[1] * B[2] + B[4] * B[5];
            sh[5][tid] = B[2] * B[2] + B[5] * B[5];
            sh[6][tid] = B[0] * e0 + B[3] * e1;
            sh[7][tid] = B[1] * e0 + B[4] * e1;
            sh[8][tid] = B[2] * e0 + B[5] * e1;
        }
        fence_proxy_async();                                 // the tile was written through the generic proxy
        __syncthreads();
        if (tid == 0) { bulk_s2g(W + (size_t)o0 * 18, st, (unsigned)cnt * 144u); bulk_commit(); }
        if (tid < np) {                                      // owner: ascending camera order (compute_V.cl:24-31 / compute_g.cl:43-54)
            double acc[9];
#pragma unroll
            for (int v = 0; v < 9; ++v) acc[v] = 0.0;
            for (int u = cur.a; u < cur.b; ++u) {
#pragma unroll
                for (int v = 0; v < 9; ++v) acc[v] += sh[v][u - o0];
            }
            double *Vp = V + (size_t)(p0 + tid) * 6;
#pragma unroll
            for (int v = 0; v < 6; ++v) Vp[v] = coeff * acc[v];
            double *gp = gb + (size_t)(p0 + tid) * 3;
            gp[0] = coeff_g * acc[6]; gp[1] = coeff_g * acc[7]; gp[2] = coeff_g * acc[8];
        }
        cur = mid; mid = far;
        ds = ds1; ds1 = ds2; ds2 = ds3;
    }
    cp_async_wait<0>();
    if (tid == 0) bulk_wait0();
}

// camera-major pass: chunk = up to CAM_CTA*CAM_OPT observations of ONE camera (ascending point).
// The camera cache entry is uniform per CTA; points and measurements are gathered (L2-resident).
// Each thread accumulates A^T A (21 upper entries) and A^T e (6) over its observations, then one
// deterministic block reduction per chunk writes 27 partials.
#ifndef LINC_MINB
#define LINC_MINB 4
#endif
template <bool EXT>
__global__ void __launch_bounds__(CAM_CTA, LINC_MINB) k_lin_cams(const int *__restrict__ cchunk_cam, const int *__restrict__ cchunk_beg,
                                                        const int *__restrict__ cchunk_end, const int *__restrict__ cam_pt,
                                                        const double *__restrict__ cam_impts,
                                                        const double *__restrict__ cache, const double *__restrict__ pts,
                                                        double *__restrict__ part, const int *__restrict__ cam_obs, psba_ext ext)
{
    __shared__ double sh[16 * (CAM_CTA + 4)];
    const int ch = blockIdx.x;
    const int beg = cchunk_beg[ch], end = cchunk_end[ch];
    CamReg cam;
    load_cam<true>(cache + (size_t)cchunk_cam[ch] * CAMC, cam);
    double acc[27];
#pragma unroll
    for (int q = 0; q < 27; ++q) acc[q] = 0.0;
    // point index and measurement come from camera-major copies made at set-up (coalesced streams); only the
    // point itself, which changes every iteration, is gathered.  A chunk holds at most CAM_OPT observations per thread: all
    // their indices, then all their points are requested before the first Jacobian is formed (the pass is bound by the
    // latency of these gathers: one observation at a time 0.154 ms; four in flight 0.167 / 0.123 / 0.116 ms at 2 / 3 / 4 CTAs per SM)
    int pi[CAM_OPT];
    double2 mm[CAM_OPT];
    double X[CAM_OPT][3];
#pragma unroll
    for (int u = 0; u < CAM_OPT; ++u) {
        const int t = beg + threadIdx.x + u * CAM_CTA;
        pi[u] = t < end ? __ldg(cam_pt + t) : -1;
        mm[u] = t < end ? __ldg(reinterpret_cast<const double2 *>(cam_impts) + t) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < CAM_OPT; ++u) {
        const double *P = pts + (size_t)(pi[u] < 0 ? 0 : pi[u]) * 3;
        X[u][0] = __ldg(P); X[u][1] = __ldg(P + 1); X[u][2] = __ldg(P + 2);
    }
#pragma unroll
    for (int u = 0; u < CAM_OPT; ++u) {
        if (pi[u] < 0) continue;
        const int t = beg + threadIdx.x + u * CAM_CTA;
        double e0, e1, A[12], B[6];
        if (EXT) residual_jac_ext(cam, ext, cchunk_cam[ch], __ldg(cam_obs + t), X[u][0], X[u][1], X[u][2], mm[u].x, mm[u].y, e0, e1, A, B);
        else residual_jac(cam, X[u][0], X[u][1], X[u][2], mm[u].x, mm[u].y, e0, e1, A, B);
        int q = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int cc = r; cc < 6; ++cc) acc[q++] += A[r] * A[cc] + A[6 + r] * A[6 + cc];
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[21 + r] += A[r] * e0 + A[6 + r] * e1;
    }
    block_reduce_to<27, CAM_CTA, 16>(acc, sh, part + (size_t)ch * 27);
}

// per camera: sum chunk partials in order; expand the 21 upper entries to the full 6x6 block
__global__ void k_cam_reduce(int m, const int *__restrict__ cam_cchunk_ptr, const double *__restrict__ part,
                             double coeff, double coeff_g, double *__restrict__ U, double *__restrict__ ga)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int j = t / 27, v = t - j * 27;
    if (j >= m) return;
    double s = 0.0;
    for (int ch = cam_cchunk_ptr[j]; ch < cam_cchunk_ptr[j + 1]; ++ch) s += part[(size_t)ch * 27 + v];
    if (v < 21) {
        int r = 0, rem = v;
        while (rem >= 6 - r) { rem -= 6 - r; ++r; }
        int cc = r + rem;
        s *= coeff;
        U[j * 36 + r * 6 + cc] = s;
        U[j * 36 + cc * 6 + r] = s;
    } else ga[j * 6 + (v - 21)] = coeff_g * s;
}

void psba_launch_linearize(psba_ctx *c, double coeff_uvw, double coeff_g)
{
    const int set = c->cur;
    if (!c->cache_valid[set]) psba_launch_cam_prep(c, set);
    // the camera pass (FP64-bound) and the point pass (memory-bound) are independent: outside profile mode the
    // camera pass runs on a second stream under the point pass
    const bool fork = !c->profile && c->n_cchunk > 0 && c->n_ptchunk > 0;
    cudaStream_t cs = c->stream;
    if (fork) {
        CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
        CUDA_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
        cs = c->stream2;
    }
    if (c->n_cchunk > 0)
        PROF(c, KID_LIN_CAMS) {
            if (c->ext_on) k_lin_cams<true><<<c->n_cchunk, CAM_CTA, 0, cs>>>(c->cchunk_cam, c->cchunk_beg, c->cchunk_end, c->cam_pt, c->cam_impts,
                                                                          c->camcache[set], c->pts[set], c->cam_part, c->cam_obs, c->ext);
            else k_lin_cams<false><<<c->n_cchunk, CAM_CTA, 0, cs>>>(c->cchunk_cam, c->cchunk_beg, c->cchunk_end, c->cam_pt, c->cam_impts,
                                                                 c->camcache[set], c->pts[set], c->cam_part, c->cam_obs, c->ext);
        }
    // N > 1 GPUs: ga goes right behind U so that ONE all-reduce sums both
    double *ga_out = c->nranks > 1 ? c->U + (size_t)c->m * 36 : c->g;
    PROF(c, KID_CAM_REDUCE) k_cam_reduce<<<cdiv(c->m * 27, 128), 128, 0, cs>>>(c->m, c->cam_cchunk_ptr, c->cam_part, coeff_uvw, coeff_g,
                                                             c->U, ga_out);
    if (c->n_ptchunk > 0)
        PROF(c, KID_LIN_POINTS) {
            const int dyn = 2 * PT_CTA * CAM_LD2 * (int)sizeof(double);
            psba_set_smem((const void *)k_lin_points_pipe<0>, dyn);
            // the persistent kernel pays for its three-stage prologue only when a CTA sees enough chunks (measured on
            // Venice-52, 2 700 chunks: 57 us against 31 us for the one-shot kernel)
            static const int pipe_env = getenv("PSBA_LIN_PIPE") ? atoi(getenv("PSBA_LIN_PIPE")) : -1;
            const bool pipe = pipe_env >= 0 ? pipe_env != 0 : c->n_small >= 8 * c->n_sm * LINP_MINB;
            if (c->ext_on)     // extended camera model (distortion / residual weights): the one-shot kernel carries it
                k_lin_points<true><<<c->n_ptchunk, PT_CTA, 0, c->stream>>>(nullptr, c->ptdesc, c->pt_ptr, c->iidx, c->jidx, c->impts, c->camcache[set], c->pts[set],
                                                                          coeff_uvw, coeff_g, c->W, c->V, c->g + c->N, c->ext);
            else if (!pipe)
                k_lin_points<false><<<c->n_ptchunk, PT_CTA, 0, c->stream>>>(nullptr, c->ptdesc, c->pt_ptr, c->iidx, c->jidx, c->impts, c->camcache[set], c->pts[set],
                                                                    coeff_uvw, coeff_g, c->W, c->V, c->g + c->N, c->ext);
            else {
                if (c->n_small > 0)
                    k_lin_points_pipe<0><<<std::min(c->n_small, c->n_sm * LINP_MINB), PT_CTA, dyn, c->stream>>>(c->n_small, c->d_small_list, c->ptdesc, c->pt_ptr, c->iidx, c->jidx,
                                                                                                     c->impts, c->camcache[set] + (size_t)c->m * CAMC, c->pts[set], coeff_uvw, coeff_g,
                                                                                                     c->W, c->V, c->g + c->N);
                if (c->n_big > 0)      // points with more observations than one wave
                    k_lin_points<false><<<c->n_big, PT_CTA, 0, c->stream>>>(c->d_big_list, c->ptdesc, c->pt_ptr, c->iidx, c->jidx, c->impts, c->camcache[set], c->pts[set],
                                                                    coeff_uvw, coeff_g, c->W, c->V, c->g + c->N, c->ext);
            }
        }
    if (fork) {
        CUDA_CHECK(cudaEventRecord(c->ev_join, c->stream2));
        CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    }
    c->st_launches += 3; c->st_lin += 1;
    LAUNCH_CHECK();
    if (c->nranks > 1) {
        psba_allreduce_sum(c, c->U, (size_t)c->m * 42);
        CUDA_CHECK(cudaMemcpyAsync(c->g, c->U + (size_t)c->m * 36, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    c->coeff_uvw = coeff_uvw; c->coeff_g = coeff_g;
    c->lin_valid = true; c->S_valid = false; c->factor_valid = false;
}

// ------------------------------------------------------------------------------------------------
// compat only: materialise JA (o x 2 x 6) and JB (o x 2 x 3) in the reference's layout
__global__ void k_jac_materialize(int o, const int *__restrict__ iidx, const int *__restrict__ jidx,
                                  const double *__restrict__ impts, const double *__restrict__ cache,
                                  const double *__restrict__ pts, double *__restrict__ JA, double *__restrict__ JB, psba_ext ext)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= o) return;
    CamReg cam;
    load_cam<true>(cache + (size_t)jidx[k] * CAMC, cam);
    const double *X = pts + (size_t)iidx[k] * 3;
    double e0, e1, A[12], B[6];
    if (ext.kc || ext.wgt) residual_jac_ext(cam, ext, jidx[k], k, X[0], X[1], X[2], impts[2 * k], impts[2 * k + 1], e0, e1, A, B);
    else residual_jac(cam, X[0], X[1], X[2], impts[2 * k], impts[2 * k + 1], e0, e1, A, B);
    for (int q = 0; q < 12; ++q) JA[(size_t)k * 12 + q] = A[q];
    for (int q = 0; q < 6; ++q) JB[(size_t)k * 6 + q] = B[q];
}

void psba_launch_jac_materialize(psba_ctx *c, double *JA, double *JB)
{
    const int set = c->cur;
    if (!c->cache_valid[set]) psba_launch_cam_prep(c, set);
    if (c->o > 0)
        k_jac_materialize<<<cdiv(c->o, 128), 128, 0, c->stream>>>(c->o, c->iidx, c->jidx, c->impts, c->camcache[set],
                                                                 c->pts[set], JA, JB, c->ext);
    c->st_launches += 1;
    LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
// J*x and J*y per observation with fused <Jx,Jx>, <Jx,Jy>, <Jy,Jy> (trust_region.cpp:125-126,
// 166-176, 208-210).  x, y are [N | 3n] vectors; y may alias x.
__global__ void __launch_bounds__(256) k_Jdot(int o, int N, const int *__restrict__ iidx, const int *__restrict__ jidx,
                                              const double *__restrict__ impts, const double *__restrict__ cache,
                                              const double *__restrict__ pts, const double *__restrict__ x,
                                              const double *__restrict__ y, double *__restrict__ Jx_out,
                                              double *__restrict__ part, psba_ext ext)
{
    __shared__ double sh[3][8];
    int k = blockIdx.x * 256 + threadIdx.x;
    double r0 = 0, r1 = 0, r2 = 0;
    if (k < o) {
        const int i = iidx[k], j = jidx[k];
        CamReg cam;
        load_cam<true>(cache + (size_t)j * CAMC, cam);
        const double *X = pts + (size_t)i * 3;
        double e0, e1, A[12], B[6];
        if (ext.kc || ext.wgt) residual_jac_ext(cam, ext, j, k, __ldg(X), __ldg(X + 1), __ldg(X + 2), 0.0, 0.0, e0, e1, A, B);
        else residual_jac(cam, __ldg(X), __ldg(X + 1), __ldg(X + 2), 0.0, 0.0, e0, e1, A, B);
        const double *xa = x + j * 6, *xb = x + N + (size_t)i * 3;
        const double *ya = y + j * 6, *yb = y + N + (size_t)i * 3;
        double jx[2], jy[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            double sx = 0.0, sy = 0.0;
#pragma unroll
            for (int t = 0; t < 6; ++t) { sx += A[r * 6 + t] * xa[t]; sy += A[r * 6 + t] * ya[t]; }
#pragma unroll
            for (int t = 0; t < 3; ++t) { sx += B[r * 3 + t] * xb[t]; sy += B[r * 3 + t] * yb[t]; }
            jx[r] = sx; jy[r] = sy;
        }
        if (Jx_out) { Jx_out[2 * (size_t)k] = jx[0]; Jx_out[2 * (size_t)k + 1] = jx[1]; }
        r0 = jx[0] * jx[0] + jx[1] * jx[1];
        r1 = jx[0] * jy[0] + jx[1] * jy[1];
        r2 = jy[0] * jy[0] + jy[1] * jy[1];
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        r0 += __shfl_down_sync(0xffffffffu, r0, w);
        r1 += __shfl_down_sync(0xffffffffu, r1, w);
        r2 += __shfl_down_sync(0xffffffffu, r2, w);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = r0; sh[1][threadIdx.x >> 5] = r1; sh[2][threadIdx.x >> 5] = r2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sh[threadIdx.x][w];
        part[(size_t)blockIdx.x * 3 + threadIdx.x] = s;
    }
}

void psba_enqueue_Jdot(psba_ctx *c, const double *x, const double *y, double *Jx_out, int off)
{
    const int set = c->cur;
    if (!c->cache_valid[set]) psba_launch_cam_prep(c, set);
    int nb = cdiv(c->o, 256);
    if (nb > 0)
        PROF(c, KID_JDOT) k_Jdot<<<nb, 256, 0, c->stream>>>(c->o, c->N, c->iidx, c->jidx, c->impts, c->camcache[set], c->pts[set],
                                         x, y, Jx_out, c->d_part, c->ext);
    PROF(c, KID_REDUCE) k_final_reduce<<<1, 256, 0, c->stream>>>(c->d_part, nb, 3, 3, c->d_scal + off);
    c->st_launches += 2;
    LAUNCH_CHECK();
    if (c->nranks > 1) psba_allreduce_sum(c, c->d_scal + off, 3);
}

void psba_launch_Jdot(psba_ctx *c, const double *x, const double *y, double *Jx_out, double res[3])
{
    psba_enqueue_Jdot(c, x, y, Jx_out, 0);
    read_scalars(c, 0, 3);
    res[0] = c->h_scal[0]; res[1] = c->h_scal[1]; res[2] = c->h_scal[2];
}
