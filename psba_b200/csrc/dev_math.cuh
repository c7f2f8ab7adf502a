// dev_math.cuh -- device-side geometry of the PSBA camera model (FP64).
//
// Camera j: fixed K = {fu,u0,v0,ar,s}, fixed unit quaternion q0, optimised {v(3), t(3)} with the
// local rotation ql = (sqrt(1-|v|^2), v) and q = ql (x) q0 (CL_files/compute_exQT.cl:33-49).
// Point in the camera frame: Xc = q X q* + t (compute_exQT.cl:51-65);
// projection x = (fu*Xc + s*Yc + u0*Zc)/Zc, y = (fu*ar*Yc + v0*Zc)/Zc (compute_exQT.cl:68-69);
// residual e = measured - projected.
//
// Everything that depends on the camera only is hoisted into a 24-double cache entry (k_cam_prep):
// q (4), t (3), K (5) and the three quaternions D_k = d q / d v_k (12).  With b = X (x) q*,
//   Xc        = vec(q (x) b) + t,
//   dXc/dv_k  = 2 vec(D_k (x) b),        dXc/dX = M(q),
// so an observation costs ~150 FP64 operations instead of the ~455 of the machine-generated
// Jacobian in compute_jacobiQT.cl:7-141 (same quantity), and gathers 192 B of camera data.
#pragma once
#include "psba_internal.h"

// camera-cache layout (doubles); CAMC = 24
#define CC_Q 0     // 4  total quaternion (s, x, y, z)
#define CC_T 4     // 3  t
#define CC_K 7     // 5  fu,u0,v0,ar,s
#define CC_D 12    // 12 D_0, D_1, D_2 (each s,x,y,z)

struct CamReg { double q[4], t[3], K[5], D[12]; };      // full entry (Jacobian)
struct CamProj { double q[4], t[3], K[5]; };            // first 12 doubles (residual only)

// entry -> registers from any 16-byte aligned source (global through the read-only path, or shared)
template <bool GLOBAL>
__device__ __forceinline__ void load_cam(const double *__restrict__ cc, CamReg &c)
{
    const double2 *p = reinterpret_cast<const double2 *>(cc);
    double buf[24];
#pragma unroll
    for (int i = 0; i < 12; ++i) { double2 v = GLOBAL ? __ldg(p + i) : p[i]; buf[2 * i] = v.x; buf[2 * i + 1] = v.y; }
#pragma unroll
    for (int i = 0; i < 4; ++i) c.q[i] = buf[CC_Q + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) c.t[i] = buf[CC_T + i];
#pragma unroll
    for (int i = 0; i < 5; ++i) c.K[i] = buf[CC_K + i];
#pragma unroll
    for (int i = 0; i < 12; ++i) c.D[i] = buf[CC_D + i];
}
// COMPACT record (16 doubles = 128 B = four sectors instead of six): q0 (4), ds = -v / sl (3), sl = sqrt(1 - |v|^2), t (3), K (5).
// v = -ds sl (no division per observation), then q and the D_k with the expressions of k_cam_prep (kernels_obs.cu) -- the pipelined point pass fetches one record per
// observation, and what bounds it is the L2 -> SM traffic of these fetches (profiles/ncu_full_r02b.md)
#define CAMC2 16
__device__ __forceinline__ void load_cam_compact(const double *cc, CamReg &c)
{
    const double2 *p = reinterpret_cast<const double2 *>(cc);
    double b[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const double2 v = p[i]; b[2 * i] = v.x; b[2 * i + 1] = v.y; }
    const double s0 = b[0], a1 = b[1], a2 = b[2], a3 = b[3], sl = b[7];
    const double v1 = -b[4] * sl, v2 = -b[5] * sl, v3 = -b[6] * sl;
    c.q[0] = sl * s0 - (a1 * v1 + a2 * v2 + a3 * v3);
    c.q[1] = s0 * v1 + sl * a1 + a3 * v2 - a2 * v3;
    c.q[2] = s0 * v2 + sl * a2 + a1 * v3 - a3 * v1;
    c.q[3] = s0 * v3 + sl * a3 + a2 * v1 - a1 * v2;
    c.t[0] = b[8]; c.t[1] = b[9]; c.t[2] = b[10];
#pragma unroll
    for (int k = 0; k < 5; ++k) c.K[k] = b[11 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double ds_l = b[4 + k];
        const double e1 = (k == 0), e2 = (k == 1), e3 = (k == 2);
        c.D[4 * k + 0] = ds_l * s0 - (a1 * e1 + a2 * e2 + a3 * e3);
        c.D[4 * k + 1] = s0 * e1 + ds_l * a1 + a3 * e2 - a2 * e3;
        c.D[4 * k + 2] = s0 * e2 + ds_l * a2 + a1 * e3 - a3 * e1;
        c.D[4 * k + 3] = s0 * e3 + ds_l * a3 + a2 * e1 - a1 * e2;
    }
}
template <bool GLOBAL>
__device__ __forceinline__ void load_cam_proj(const double *__restrict__ cc, CamProj &c)
{
    const double2 *p = reinterpret_cast<const double2 *>(cc);
    double buf[12];
#pragma unroll
    for (int i = 0; i < 6; ++i) { double2 v = GLOBAL ? __ldg(p + i) : p[i]; buf[2 * i] = v.x; buf[2 * i + 1] = v.y; }
#pragma unroll
    for (int i = 0; i < 4; ++i) c.q[i] = buf[CC_Q + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) c.t[i] = buf[CC_T + i];
#pragma unroll
    for (int i = 0; i < 5; ++i) c.K[i] = buf[CC_K + i];
}

// b = X (x) q* = (w.X, sX + w x X);  Xc = vec(q (x) b) + t  (the two quaternion products of
// compute_exQT.cl:51-65)
template <class CAM>
__device__ __forceinline__ void cam_transform(const CAM &c, double X, double Y, double Z,
                                              double &b0, double &b1, double &b2, double &b3,
                                              double &xc, double &yc, double &zc)
{
    const double s = c.q[0], w1 = c.q[1], w2 = c.q[2], w3 = c.q[3];
    b0 = X * w1 + w2 * Y + w3 * Z;
    b1 = s * X + w2 * Z - w3 * Y;
    b2 = Y * s + w3 * X - Z * w1;
    b3 = s * Z + Y * w1 - w2 * X;
    xc = w1 * b0 + s * b1 - b2 * w3 + w2 * b3 + c.t[0];
    yc = w2 * b0 + s * b2 - b3 * w1 + w3 * b1 + c.t[1];
    zc = b0 * w3 + s * b3 - w2 * b1 + w1 * b2 + c.t[2];
}

// e = measured - projected
template <class CAM>
__device__ __forceinline__ void residual(const CAM &c, double X, double Y, double Z, double mx, double my, double &e0, double &e1)
{
    double b0, b1, b2, b3, xc, yc, zc;
    cam_transform(c, X, Y, Z, b0, b1, b2, b3, xc, yc, zc);
    const double iz = 1.0 / zc;
    e0 = mx - (c.K[0] * xc + c.K[4] * yc + c.K[1] * zc) * iz;
    e1 = my - (c.K[0] * c.K[3] * yc + c.K[2] * zc) * iz;
}

// residual + A (2x6 row-major: d proj / d(v,t)) + B (2x3 row-major: d proj / dX)
__device__ __forceinline__ void residual_jac(const CamReg &c, double X, double Y, double Z, double mx, double my,
                                             double &e0, double &e1, double *A, double *B)
{
    double b0, b1, b2, b3, xc, yc, zc;
    cam_transform(c, X, Y, Z, b0, b1, b2, b3, xc, yc, zc);
    const double iz = 1.0 / zc;
    const double fa = c.K[0] * c.K[3];
    e0 = mx - (c.K[0] * xc + c.K[4] * yc + c.K[1] * zc) * iz;
    e1 = my - (fa * yc + c.K[2] * zc) * iz;
    // P = d proj / d Xc
    const double p00 = c.K[0] * iz, p01 = c.K[4] * iz, p02 = -(c.K[0] * xc + c.K[4] * yc) * iz * iz;
    const double p11 = fa * iz, p12 = -(fa * yc) * iz * iz;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // dXc/dv_k = 2 vec(D_k (x) b)
        const double ds = c.D[4 * k], dx = c.D[4 * k + 1], dy = c.D[4 * k + 2], dz = c.D[4 * k + 3];
        const double d0 = 2.0 * (ds * b1 + b0 * dx + dy * b3 - dz * b2);
        const double d1 = 2.0 * (ds * b2 + b0 * dy + dz * b1 - dx * b3);
        const double d2 = 2.0 * (ds * b3 + b0 * dz + dx * b2 - dy * b1);
        A[k] = p00 * d0 + p01 * d1 + p02 * d2;
        A[6 + k] = p11 * d1 + p12 * d2;
    }
    A[3] = p00; A[4] = p01; A[5] = p02;
    A[9] = 0.0; A[10] = p11; A[11] = p12;
    // B = P * M(q)  (general, non-normalised form as compute_jacobiQT.cl:117-140)
    const double s = c.q[0], x = c.q[1], y = c.q[2], z = c.q[3];
    const double m00 = s * s + x * x - y * y - z * z, m01 = 2 * (x * y - s * z), m02 = 2 * (x * z + s * y);
    const double m10 = 2 * (x * y + s * z), m11 = s * s - x * x + y * y - z * z, m12 = 2 * (y * z - s * x);
    const double m20 = 2 * (x * z - s * y), m21 = 2 * (y * z + s * x), m22 = s * s - x * x - y * y + z * z;
    B[0] = p00 * m00 + p01 * m10 + p02 * m20;
    B[1] = p00 * m01 + p01 * m11 + p02 * m21;
    B[2] = p00 * m02 + p01 * m12 + p02 * m22;
    B[3] = p11 * m10 + p12 * m20;
    B[4] = p11 * m11 + p12 * m21;
    B[5] = p11 * m12 + p12 * m22;
}

// ---------------------------------------------------------------------------------------------
// Extended camera model (the reference's kernels have neither, F7): fixed per-camera
// distortion kc[5] = (k1, k2, p1, p2, k3) of Lourakis' sba "varKD" camera (Bouguet model) applied to the normalised
// image point before the reference's affine K (data/54camsvarKD.txt columns 6-10, PSBA/misc.cpp:27-29 copies them
// through quat2vec), and a per-observation lower-triangular residual weight (w00, w10, w11) with
// W^T W = inverse image-point covariance (the covimgpts the reader parses, PSBA/readparams.cpp:272-283, 380-413).
//   xn = Xc/Zc, yn = Yc/Zc, r2 = xn^2 + yn^2, c = 1 + k1 r2 + k2 r2^2 + k3 r2^3,
//   xd = c xn + 2 p1 xn yn + p2 (r2 + 2 xn^2),  yd = c yn + p1 (r2 + 2 yn^2) + 2 p2 xn yn,
//   x = fu xd + s yd + u0,  y = fu ar yd + v0.       kc = 0: compute_exQT.cl:68-69.
// Kernels take the model as a template flag: the default path (no distortion, no weights) compiles to the code above.

// projection (px, py) and, if P != nullptr, P[6] = d(px, py) / d(xc, yc, zc) row-major
template <class CAM>
__device__ __forceinline__ void project_ext(const CAM &c, const double *__restrict__ kc, double xc, double yc, double zc,
                                            double &px, double &py, double *P)
{
    double k0 = 0, k1 = 0, k2 = 0, k3 = 0, k4 = 0;
    if (kc) { k0 = __ldg(kc); k1 = __ldg(kc + 1); k2 = __ldg(kc + 2); k3 = __ldg(kc + 3); k4 = __ldg(kc + 4); }
    const double iz = 1.0 / zc, xn = xc * iz, yn = yc * iz, r2 = xn * xn + yn * yn;
    const double cd = 1.0 + r2 * (k0 + r2 * (k1 + r2 * k4)), dc = k0 + r2 * (2.0 * k1 + 3.0 * k4 * r2);
    const double xd = cd * xn + 2.0 * k2 * xn * yn + k3 * (r2 + 2.0 * xn * xn);
    const double yd = cd * yn + k2 * (r2 + 2.0 * yn * yn) + 2.0 * k3 * xn * yn;
    const double fa = c.K[0] * c.K[3];
    px = c.K[0] * xd + c.K[4] * yd + c.K[1];
    py = fa * yd + c.K[2];
    if (P) {
        const double j00 = cd + 2.0 * dc * xn * xn + 2.0 * k2 * yn + 6.0 * k3 * xn, j01 = 2.0 * dc * xn * yn + 2.0 * k2 * xn + 2.0 * k3 * yn;
        const double j11 = cd + 2.0 * dc * yn * yn + 6.0 * k2 * yn + 2.0 * k3 * xn;
        const double g00 = c.K[0] * j00 + c.K[4] * j01, g01 = c.K[0] * j01 + c.K[4] * j11, g10 = fa * j01, g11 = fa * j11;
        P[0] = g00 * iz; P[1] = g01 * iz; P[2] = -(g00 * xn + g01 * yn) * iz;
        P[3] = g10 * iz; P[4] = g11 * iz; P[5] = -(g10 * xn + g11 * yn) * iz;
    }
}

// e = W (measured - projected); j = camera, k = observation
template <class CAM>
__device__ __forceinline__ void residual_ext(const CAM &c, const psba_ext &x, int j, int k, double X, double Y, double Z,
                                             double mx, double my, double &e0, double &e1)
{
    double b0, b1, b2, b3, xc, yc, zc, px, py;
    cam_transform(c, X, Y, Z, b0, b1, b2, b3, xc, yc, zc);
    project_ext(c, x.kc ? x.kc + (size_t)j * 5 : nullptr, xc, yc, zc, px, py, nullptr);
    e0 = mx - px; e1 = my - py;
    if (x.wgt) { const double *w = x.wgt + (size_t)k * 3; const double w00 = __ldg(w), w10 = __ldg(w + 1), w11 = __ldg(w + 2); e1 = w10 * e0 + w11 * e1; e0 = w00 * e0; }
}

// weighted residual + A = W d proj / d(v,t) (2x6) + B = W d proj / dX (2x3)
__device__ __forceinline__ void residual_jac_ext(const CamReg &c, const psba_ext &x, int j, int k, double X, double Y, double Z,
                                                 double mx, double my, double &e0, double &e1, double *A, double *B)
{
    double b0, b1, b2, b3, xc, yc, zc, px, py, P[6];
    cam_transform(c, X, Y, Z, b0, b1, b2, b3, xc, yc, zc);
    project_ext(c, x.kc ? x.kc + (size_t)j * 5 : nullptr, xc, yc, zc, px, py, P);
    e0 = mx - px; e1 = my - py;
    if (x.wgt) {
        const double *w = x.wgt + (size_t)k * 3;
        const double w00 = __ldg(w), w10 = __ldg(w + 1), w11 = __ldg(w + 2);
        e1 = w10 * e0 + w11 * e1; e0 = w00 * e0;
#pragma unroll
        for (int q = 0; q < 3; ++q) { P[3 + q] = w10 * P[q] + w11 * P[3 + q]; P[q] = w00 * P[q]; }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const double ds = c.D[4 * q], dx = c.D[4 * q + 1], dy = c.D[4 * q + 2], dz = c.D[4 * q + 3];
        const double d0 = 2.0 * (ds * b1 + b0 * dx + dy * b3 - dz * b2);
        const double d1 = 2.0 * (ds * b2 + b0 * dy + dz * b1 - dx * b3);
        const double d2 = 2.0 * (ds * b3 + b0 * dz + dx * b2 - dy * b1);
        A[q] = P[0] * d0 + P[1] * d1 + P[2] * d2;
        A[6 + q] = P[3] * d0 + P[4] * d1 + P[5] * d2;
        A[3 + q] = P[q]; A[9 + q] = P[3 + q];
    }
    const double s = c.q[0], qx = c.q[1], qy = c.q[2], qz = c.q[3];
    const double m00 = s * s + qx * qx - qy * qy - qz * qz, m01 = 2 * (qx * qy - s * qz), m02 = 2 * (qx * qz + s * qy);
    const double m10 = 2 * (qx * qy + s * qz), m11 = s * s - qx * qx + qy * qy - qz * qz, m12 = 2 * (qy * qz - s * qx);
    const double m20 = 2 * (qx * qz - s * qy), m21 = 2 * (qy * qz + s * qx), m22 = s * s - qx * qx - qy * qy + qz * qz;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        B[r * 3] = P[r * 3] * m00 + P[r * 3 + 1] * m10 + P[r * 3 + 2] * m20;
        B[r * 3 + 1] = P[r * 3] * m01 + P[r * 3 + 1] * m11 + P[r * 3 + 2] * m21;
        B[r * 3 + 2] = P[r * 3] * m02 + P[r * 3 + 1] * m12 + P[r * 3 + 2] * m22;
    }
}

// ---------------------------------------------------------------------------------------------
// deterministic block reduction of NV per-thread values (fixed summation order, no atomics):
// values are staged through shared memory G at a time; 8 lanes sum interleaved slices of each
// value's column, then a 3-step shuffle tree combines the 8 slices.  sh must hold G*(NT+4) doubles.
// Result (valid in the threads that own it) is written to out[0..NV).
template <int NV, int NT, int G>
__device__ __forceinline__ void block_reduce_to(const double *v, double *sh, double *out)
{
    const int tid = threadIdx.x;
    constexpr int LD = NT + 4;
#pragma unroll
    for (int g0 = 0; g0 < NV; g0 += G) {
#pragma unroll
        for (int k = 0; k < G; ++k)
            if (g0 + k < NV) sh[k * LD + tid] = v[g0 + k];
        __syncthreads();
        if (tid < G * 8) {
            const int k = tid >> 3, seg = tid & 7;
            double sum = 0.0;
            if (g0 + k < NV) {
                const double *col = sh + k * LD;
#pragma unroll 4
                for (int i = seg; i < NT; i += 8) sum += col[i];
            }
            sum += __shfl_down_sync(0xffffffffu, sum, 4, 8);
            sum += __shfl_down_sync(0xffffffffu, sum, 2, 8);
            sum += __shfl_down_sync(0xffffffffu, sum, 1, 8);
            if (seg == 0 && g0 + k < NV) out[g0 + k] = sum;
        }
        __syncthreads();
    }
}


// ---- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) with mbarrier completion ----------------------
// One thread arms the barrier with the byte count and issues the copy; every consumer waits on the phase
// parity.  Addresses and sizes are multiples of 16 bytes.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- 16-byte asynchronous copies global -> shared (cp.async.cg, SASS LDGSTS: no registers, L1 bypass) ----
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- bulk store shared -> global (TMA, bulk async-group completion) ----------------------------------
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- mbarrier arrivals: a consumer's plain arrive, and the arrive that fires when the executing thread's earlier
// cp.async copies have landed (.noinc: the arrival is part of the count the barrier was initialised with)
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
