/*
 * ref_kernels.cpp -- ORACLE PIN (test infrastructure): runs the reference's OWN OpenCL kernel
 * bodies on the host.  The 17 kernel files below are #included UNCHANGED from where they lie
 * under REF_ROOT (nothing is copied into this repository); a macro shim maps the OpenCL-C
 * qualifiers onto C++, and every kernel is executed by a loop over the NDRange that
 * PSBA/sba_func.cpp launches it with (dimension 0 fastest).  The output library lives in
 * oracle/_ref/ (git-ignored).
 *
 * Not compilable this way: CL_files/SPD_inv.cl and cholmod_blk.cl (OpenCL-2.0 Blocks +
 * enqueue_kernel) -- those are restated in orc_chol.c.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include "psba_oracle.h"

static thread_local size_t g_gid[3];
#define get_global_id(d) (g_gid[d])
struct double3 { double x, y, z; };
static inline double dot(double3 a, double3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
#define __kernel static
#define __global
#define dtype  double
#define dtype3 double3
#define cnp 6
#define pnp 3
#define mnp 2

#define RSTR2(x) #x
#define RSTR(x) RSTR2(x)
#define REF_CL(name) RSTR(REF_ROOT/CL_files/name)

#include REF_CL(compute_exQT.cl)
#undef QtX
#undef QtY
#undef QtZ
#undef X
#undef Y
#undef Z
#undef ax
#undef ay
#undef x0
#undef y0
#include REF_CL(compute_jacobiQT.cl)
#include REF_CL(compute_U.cl)
#include REF_CL(compute_V.cl)
#include REF_CL(compute_Wblks.cl)
#include REF_CL(compute_g.cl)
#include REF_CL(update_UV.cl)
#include REF_CL(compute_Vinv.cl)
#include REF_CL(restore_UVdiag.cl)
#include REF_CL(compute_Yblks.cl)
#include REF_CL(compute_S.cl)
#include REF_CL(compute_ea.cl)
#include REF_CL(matVec_mul.cl)
#include REF_CL(compute_eb.cl)
#include REF_CL(compute_dpb.cl)
#include REF_CL(compute_newp.cl)
#include REF_CL(update_p.cl)
#include REF_CL(compute_Jmultiply.cl)

#define NT int nt = s->nthreads > 0 ? s->nthreads : 1; (void)nt
#define PFOR _Pragma("omp parallel for schedule(static) num_threads(nt)")

/* sba_func.cpp:115 {o} */
static void r_exQT(orc_state *s, const double *cams, const double *pts, double *ex)
{
    NT; PFOR
    for (int idx = 0; idx < s->o; ++idx) {
        g_gid[0] = idx;
        kern_compute_exQT(s->K, s->impts, s->initcams, (double *)cams, (double *)pts, s->iidx, s->jidx, ex);
    }
}
/* sba_func.cpp:187 {o} */
static void r_jacobiQT(orc_state *s)
{
    NT; PFOR
    for (int idx = 0; idx < s->o; ++idx) {
        g_gid[0] = idx;
        kern_compute_jacobiQT(s->K, s->impts, s->initcams, s->cams, s->pts, s->iidx, s->jidx, s->JA, s->JB);
    }
}
/* sba_func.cpp:283 {6, 6m} */
static void r_U(orc_state *s, double coeff)
{
    NT; PFOR
    for (int tc = 0; tc < 6 * s->m; ++tc) for (int r = 0; r < 6; ++r) {
        g_gid[0] = r; g_gid[1] = tc;
        kern_compute_U(s->m, s->n, s->o, s->JA, s->blk_idx, s->U, s->UVdiag, coeff);
    }
}
/* sba_func.cpp:369 {3, 3n} */
static void r_V(orc_state *s, double coeff)
{
    NT; PFOR
    for (int tc = 0; tc < 3 * s->n; ++tc) for (int r = 0; r < 3; ++r) {
        g_gid[0] = r; g_gid[1] = tc;
        kern_compute_V(s->m, s->n, s->o, s->JB, s->blk_idx, s->V, s->UVdiag, coeff);
    }
}
/* sba_func.cpp:485 {6, 3o} */
static void r_Wblks(orc_state *s, double coeff)
{
    NT; PFOR
    for (int tc = 0; tc < 3 * s->o; ++tc) for (int r = 0; r < 6; ++r) {
        g_gid[0] = r; g_gid[1] = tc;
        kern_compute_Wblks(s->m, s->n, s->o, s->JA, s->JB, s->iidx, s->jidx, s->W, coeff);
    }
}
/* sba_func.cpp:569 {T} */
static void r_g(orc_state *s, double coeff)
{
    NT; PFOR
    for (int tr = 0; tr < s->T; ++tr) {
        g_gid[0] = tr;
        kern_compute_g(s->m, s->n, s->o, coeff, s->JA, s->JB, s->blk_idx, s->ex, s->g);
    }
}
/* sba_func.cpp:641 {T} */
static void r_update_UV(orc_state *s, double mu)
{
    for (int t = 0; t < s->T; ++t) { g_gid[0] = t; kern_update_UV(s->m, s->n, s->U, s->V, mu); }
}
/* sba_func.cpp:709 {T} */
static void r_restore_UVdiag(orc_state *s)
{
    for (int t = 0; t < s->T; ++t) { g_gid[0] = t; kern_restore_UVdiag(s->m, s->n, s->U, s->V, s->UVdiag); }
}
/* sba_func.cpp:748 {n}; *ret is written by every work-item, the last writer wins */
static double r_Vinv(orc_state *s)
{
    double any = 0.0;
    for (int i = 0; i < s->n; ++i) {
        double ret = 0.0;
        g_gid[0] = i;
        kern_compute_Vinv(s->n, s->V, &ret);
        if (ret != 0.0) any = 1.0;
    }
    s->ret = any;
    return any;
}
/* sba_func.cpp:822 {6, 3o} */
static void r_Yblks(orc_state *s)
{
    NT; PFOR
    for (int tc = 0; tc < 3 * s->o; ++tc) for (int r = 0; r < 6; ++r) {
        g_gid[0] = r; g_gid[1] = tc;
        kern_compute_Yblks(s->m, s->n, s->iidx, s->W, s->V, s->Y);
    }
}
/* sba_func.cpp:898 {N, N} */
static void r_S(orc_state *s)
{
    NT; PFOR
    for (int tc = 0; tc < s->N; ++tc) for (int tr = 0; tr < s->N; ++tr) {
        g_gid[0] = tr; g_gid[1] = tc;
        kern_compute_S(s->m, s->n, s->blk_idx, s->U, s->Y, s->W, s->comm3DIdx, s->comm3DIdxCnt, s->S);
    }
}
/* sba_func.cpp:965 {N} */
static void r_ea(orc_state *s)
{
    NT; PFOR
    for (int tr = 0; tr < s->N; ++tr) {
        g_gid[0] = tr;
        kern_compute_ea(s->m, s->n, s->o, s->blk_idx, s->Y, s->g, s->eab);
    }
}
/* cl_linearalg.cpp:31 {N} */
static void r_matVec(orc_state *s)
{
    NT; PFOR
    for (int i = 0; i < s->N; ++i) { g_gid[0] = i; kern_matVec_mul(s->N, s->S, s->eab, s->dp); }
}
/* sba_func.cpp:1030 {3n} */
static void r_eb(orc_state *s)
{
    NT; PFOR
    for (int tr = 0; tr < 3 * s->n; ++tr) {
        g_gid[0] = tr;
        kern_compute_eb(s->m, s->n, s->o, s->blk_idx, s->W, s->dp, s->g, s->eab);
    }
}
/* sba_func.cpp:1090 {3n} */
static void r_dpb(orc_state *s)
{
    NT; PFOR
    for (int tr = 0; tr < 3 * s->n; ++tr) { g_gid[0] = tr; kern_compute_dpb(s->m, s->n, s->V, s->eab, s->dp); }
}
/* sba_func.cpp:1142 {T} */
static void r_newp(orc_state *s)
{
    for (int t = 0; t < s->T; ++t) {
        g_gid[0] = t;
        kern_compute_newp(6 * s->m, 3 * s->n, s->cams, s->pts, s->dp, s->newcams, s->newpts);
    }
}
/* sba_func.cpp:1188 {T} */
static void r_update_p(orc_state *s)
{
    for (int t = 0; t < s->T; ++t) {
        g_gid[0] = t;
        kern_update_p(6 * s->m, 3 * s->n, s->cams, s->pts, s->newcams, s->newpts);
    }
}
/* sba_func.cpp:45 {2mn}: dense result, compacted to observation order (the dense vector is
 * zero wherever blk_idx < 0, compute_Jmultiply.cl:29-47) */
static void r_Jmultiply(orc_state *s, const double *x, double *out)
{
    size_t dense = (size_t)2 * s->m * s->n;
    double *tmp = (double *)malloc(dense * sizeof(double));
    NT; PFOR
    for (long long tr = 0; tr < (long long)dense; ++tr) {
        g_gid[0] = (size_t)tr;
        kern_compute_Jmultiply(s->m, s->n, s->o, s->JA, s->JB, s->blk_idx, (double *)x, tmp);
    }
    for (int idx = 0; idx < s->o; ++idx) {
        size_t b = ((size_t)s->iidx[idx] * s->m + s->jidx[idx]) * 2;
        out[idx * 2] = tmp[b]; out[idx * 2 + 1] = tmp[b + 1];
    }
    free(tmp);
}

static const orc_ops g_ref = {
    r_exQT, r_jacobiQT, r_U, r_V, r_Wblks, r_g, r_update_UV, r_restore_UVdiag, r_Vinv,
    r_Yblks, r_S, r_ea, r_matVec, r_eb, r_dpb, r_newp, r_update_p, r_Jmultiply, "reference-kernels"
};
extern "C" const orc_ops *ref_ops(void) { return &g_ref; }
