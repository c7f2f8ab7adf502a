// dev_math.cuh -- device-side geometry of the PSBA camera model (FP64).
//
// Camera j: fixed K = {fu,u0,v0,ar,s}, fixed unit quaternion q0, optimised {v(3), t(3)} with the
// local rotation ql = (sqrt(1-|v|^2), v) and q = ql (x) q0 (CL_files/compute_exQT.cl:33-49).
// Point in the camera frame: Xc = q X q* + t = M(q) X + t (compute_exQT.cl:51-65);
// projection x = (fu*Xc + s*Yc + u0*Zc)/Zc, y = (fu*ar*Yc + v0*Zc)/Zc (compute_exQT.cl:68-69);
// residual e = measured - projected.
//
// Everything that depends on the camera only is hoisted into a per-camera cache entry
// (k_cam_prep): M(q), t, K and the three matrices G_k = dM/dv_k, so that the per-observation
// work is 9 FMA for Xc, 27 FMA for dXc/dv and a 2x3 projection derivative -- about a third of
// the flops of the machine-generated Jacobian in compute_jacobiQT.cl:7-141, same quantity.
#pragma once
#include "psba_internal.h"

// camera-cache layout (doubles)
#define CC_R 0     // 9  M(q) row-major
#define CC_T 9     // 3  t
#define CC_K 12    // 5  fu,u0,v0,ar,s
#define CC_G 17    // 27 G_0,G_1,G_2 row-major

struct CamReg {            // camera cache entry held in registers
    double R[9], t[3], K[5], G[27];
};

__device__ __forceinline__ void load_cam(const double *__restrict__ cc, CamReg &c)
{
    const double2 *p = reinterpret_cast<const double2 *>(cc);
    double buf[44];
#pragma unroll
    for (int i = 0; i < 22; ++i) { double2 v = __ldg(p + i); buf[2 * i] = v.x; buf[2 * i + 1] = v.y; }
#pragma unroll
    for (int i = 0; i < 9; ++i) c.R[i] = buf[CC_R + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) c.t[i] = buf[CC_T + i];
#pragma unroll
    for (int i = 0; i < 5; ++i) c.K[i] = buf[CC_K + i];
#pragma unroll
    for (int i = 0; i < 27; ++i) c.G[i] = buf[CC_G + i];
}

// residual only needs R, t, K (17 doubles)
struct CamProj { double R[9], t[3], K[5]; };
__device__ __forceinline__ void load_cam_proj(const double *__restrict__ cc, CamProj &c)
{
    const double2 *p = reinterpret_cast<const double2 *>(cc);
    double buf[18];
#pragma unroll
    for (int i = 0; i < 9; ++i) { double2 v = __ldg(p + i); buf[2 * i] = v.x; buf[2 * i + 1] = v.y; }
#pragma unroll
    for (int i = 0; i < 9; ++i) c.R[i] = buf[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) c.t[i] = buf[9 + i];
#pragma unroll
    for (int i = 0; i < 5; ++i) c.K[i] = buf[12 + i];
}

template <class CAM>
__device__ __forceinline__ void cam_transform(const CAM &c, double X, double Y, double Z, double &xc, double &yc, double &zc)
{
    xc = c.R[0] * X + c.R[1] * Y + c.R[2] * Z + c.t[0];
    yc = c.R[3] * X + c.R[4] * Y + c.R[5] * Z + c.t[1];
    zc = c.R[6] * X + c.R[7] * Y + c.R[8] * Z + c.t[2];
}

// e = measured - projected
template <class CAM>
__device__ __forceinline__ void residual(const CAM &c, double X, double Y, double Z, double mx, double my, double &e0, double &e1)
{
    double xc, yc, zc;
    cam_transform(c, X, Y, Z, xc, yc, zc);
    double iz = 1.0 / zc;
    e0 = mx - (c.K[0] * xc + c.K[4] * yc + c.K[1] * zc) * iz;
    e1 = my - (c.K[0] * c.K[3] * yc + c.K[2] * zc) * iz;
}

// residual + A (2x6 row-major: d proj / d(v,t)) + B (2x3 row-major: d proj / dX)
__device__ __forceinline__ void residual_jac(const CamReg &c, double X, double Y, double Z, double mx, double my,
                                             double &e0, double &e1, double *A, double *B)
{
    double xc, yc, zc;
    cam_transform(c, X, Y, Z, xc, yc, zc);
    const double iz = 1.0 / zc;
    const double fa = c.K[0] * c.K[3];
    e0 = mx - (c.K[0] * xc + c.K[4] * yc + c.K[1] * zc) * iz;
    e1 = my - (fa * yc + c.K[2] * zc) * iz;
    // P = d proj / d Xc
    const double p00 = c.K[0] * iz, p01 = c.K[4] * iz, p02 = -(c.K[0] * xc + c.K[4] * yc) * iz * iz;
    const double p11 = fa * iz, p12 = -(fa * yc) * iz * iz;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double *G = c.G + 9 * k;
        double d0 = G[0] * X + G[1] * Y + G[2] * Z;
        double d1 = G[3] * X + G[4] * Y + G[5] * Z;
        double d2 = G[6] * X + G[7] * Y + G[8] * Z;
        A[k] = p00 * d0 + p01 * d1 + p02 * d2;
        A[6 + k] = p11 * d1 + p12 * d2;
    }
    A[3] = p00; A[4] = p01; A[5] = p02;
    A[9] = 0.0; A[10] = p11; A[11] = p12;
#pragma unroll
    for (int cidx = 0; cidx < 3; ++cidx) {
        B[cidx] = p00 * c.R[cidx] + p01 * c.R[3 + cidx] + p02 * c.R[6 + cidx];
        B[3 + cidx] = p11 * c.R[3 + cidx] + p12 * c.R[6 + cidx];
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
