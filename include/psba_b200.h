/*
 * psba_b200.h -- C ABI of libpsba_b200.so: a B200-native (sm_100a CUDA) replacement for the
 * hot path of eglrp/PSBA.  Every entry point names the reference interface it replaces
 * (paths relative to the reference tree).  Plain pointers and sizes only; all arithmetic is
 * FP64; indices are int32.  There is NO CPU fallback: every call needs a CUDA device and
 * aborts with "psba_b200: CUDA error ..." + exit(1) otherwise (the reference's checkErr
 * behaviour, PSBA/cl_psba.cpp:275-283).
 *
 * Conventions kept from the reference:
 *   - parameter order  (ctx, cnp, pnp, mnp, n3Dpts, nCams, n2Dprojs, [coeff|mu], out)
 *     (PSBA/sba_func.h:10-138); cnp=6, pnp=3, mnp=2 are the only supported dimensions
 *     (CL_files/PSBA.cl:5-7) and are checked;
 *   - an `out` pointer that is non-NULL makes the call copy its result to HOST memory in the
 *     reference's layout (the per-stage comparison hook); NULL keeps everything
 *     device-resident;
 *   - numerical status is returned as double: 0.0 ok, 1.0 failed (compute_Vinv, SPDinv);
 *   - drivers return the ITER_* codes of PSBA/psba.h:12-18.
 *
 * Where the reference passes cl_mem handles (cams_buffer / newCams_buffer ...) this ABI
 * takes a PSBA_PARAMS_* selector.
 */
#ifndef PSBA_B200_H
#define PSBA_B200_H

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct psba_ctx psba_ctx;          /* replaces PSBA_struct, PSBA/cl_psba.h:9-89 */

/* PSBA/psba.h:12-18 */
#define PSBA_ITER_TURN_TO_LM        1
#define PSBA_ITER_TURN_TO_TR        2
#define PSBA_ITER_CONTINUE          3
#define PSBA_ITER_ERR               4
#define PSBA_ITER_DP_NO_CHANGE      5
#define PSBA_ITER_ERR_SMALL_ENOUGH  6
#define PSBA_ITER_PASS              7

#define PSBA_PARAMS_CUR 0                  /* cams_buffer / pts3D_buffer        */
#define PSBA_PARAMS_NEW 1                  /* newCams_buffer / newPts3D_buffer  */

/* which device vector psba_compute_Jmultiply / psba_upload_vec address */
#define PSBA_VEC_G   0                     /* g_buffer  */
#define PSBA_VEC_DP  1                     /* dp_buffer */

/* ------------------------------------------------------------------ runtime (L0) ------ */

/* setup_cl, PSBA/cl_psba.cpp:16-133.  Creates the context on the current CUDA device.
 * Multi-GPU: call psba_comm_init first on every rank; n3Dpts/n2Dprojs are then the GLOBAL
 * sizes and each rank uploads only its own point range (psba_local_range). */
psba_ctx *psba_setup_cl(int cnp, int pnp, int mnp, int nCams, int n3Dpts, int n2Dprojs);

/* fill_initBuffer2, PSBA/cl_psba.cpp:138-172 (host arrays, copied) */
void psba_fill_initBuffer2(psba_ctx *ctx, int cnp, int pnp, int mnp, int nCams, int n3Dpts, int n2Dprojs,
                           const double *Kparas, const double *impts_data, const double *initcams_data,
                           const double *camsExParas, const double *pts3Ds);

/* fill_idxBuffer, PSBA/cl_psba.cpp:176-207.  Only iidx/jidx are taken: the dense tables
 * blk_idx / comm3DIdx / comm3DIdxCnt (PSBA/misc.cpp:178-218) are replaced by CSR lists that the
 * engine derives itself (same entries, same ascending order).  Observations must be
 * point-major with cameras ascending, as generate_idxs produces them. */
void psba_fill_idxBuffer(psba_ctx *ctx, int nCams, int n3Dpts, int n2Dprojs, const int *iidx, const int *jidx);

/* release_buffer, PSBA/cl_psba.cpp:210-241 */
void psba_release_buffer(psba_ctx *ctx);

/* Extended camera model (neither is used by any reference kernel).
 * set_distortion: fixed per-camera lens distortion kc[nCams*5] = (k1, k2, p1, p2, k3) of the sba "varKD" camera -- the
 * columns 6-10 of data/54camsvarKD.txt that PSBA/misc.cpp:27-29 (quat2vec) copies through and CL_files/PSBA.cl:5-7
 * (cnp = 6) then drops.  NULL / all zeros = the reference's projection.
 * set_covariances: image-point covariances as readInitialSBAEstimate parses them (PSBA/readparams.cpp:272-283, 380-413;
 * covsz = 4 full 2x2, 3 upper triangle), n2Dprojs * covsz doubles; residuals become W e with W^T W = Sigma^-1.
 * Returns 1 if a covariance is not positive definite.  Both may be called any time after fill_idxBuffer. */
void psba_set_distortion(psba_ctx *ctx, const double *kc);
int psba_set_covariances(psba_ctx *ctx, const double *cov, int covsz);

/* ------------------------------------------------------------------ operators (L2) ---- */

/* compute_exQT, PSBA/sba_func.cpp:81-145 -> kern_compute_exQT, CL_files/compute_exQT.cl:18-71.
 * Returns ||e||^2 (the reference reads ex back and sums on the host, levmar.cpp:93-94). */
double psba_compute_exQT(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs,
                         int params, double *ex);
/* compute_jacobiQT, sba_func.cpp:153-245 -> compute_jacobiQT.cl:7-141.  J is never stored on the
 * hot path; it is materialised only when jac_A / jac_B are non-NULL. */
void psba_compute_jacobiQT(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs,
                           double *jac_A, double *jac_B);
/* compute_U / compute_V / compute_Wblks / compute_g, sba_func.cpp:252-332, 338-417, 450-529, 536-617 */
void psba_compute_U(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double coeff, double *out);
void psba_compute_V(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double coeff, double *out);
void psba_compute_Wblks(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs,
                        const int *iidx, const int *jidx, double coeff, double *Wblks);
void psba_compute_g(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double coeff, double *g);
/* maxElmOfUV, sba_func.cpp:422-444 (UVdiag may be NULL) */
double psba_maxElmOfUV(psba_ctx *ctx, int totalParas, double *UVdiag);
/* update_UV / restore_UVdiag, sba_func.cpp:624-688, 694-720.  The damping term is carried as a
 * kernel argument; U and V themselves are never modified. */
void psba_update_UV(psba_ctx *ctx, int cnp, int pnp, int n3Dpts, int nCams, double mu, double *U, double *V);
void psba_restore_UVdiag(psba_ctx *ctx, int cnp, int pnp, int n3Dpts, int nCams);
/* compute_Vinv, sba_func.cpp:727-788 -> compute_Vinv.cl:6-90.  V (if non-NULL) receives the
 * reference's mixed-triangle layout (inverse in lower triangle + diagonal). */
double psba_compute_Vinv(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *V);
/* compute_Yblks, sba_func.cpp:795-860 (Y is materialised only when Yblks is non-NULL) */
void psba_compute_Yblks(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs,
                        const int *iidx, const int *jidx, double *Yblks);
/* compute_S, sba_func.cpp:866-929 -> compute_S.cl:6-78 ; S (if non-NULL) is dense N x N row-major */
void psba_compute_S(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *S);
/* compute_ea, sba_func.cpp:936-995 */
void psba_compute_ea(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *ea);
/* SPDinv, PSBA/cl_spdinv.cpp:18-40 (cholesky 57-103, trigMat_inv 120-162, trigMat_mul 169-204).
 * Factorises S in place (blocked Cholesky); the explicit inverse is formed only when outMat is
 * non-NULL.  Returns 0.0, or 1.0 when S is not positive definite. */
double psba_SPDinv(psba_ctx *ctx, int matSize, double *outMat);
/* cholesky / trigMat_inv / trigMat_mul, PSBA/cl_spdinv.h:10-18 (cl_spdinv.cpp:57-103, 120-162, 169-204): the three stages of
 * SPDinv.  outMat (N x N, may be NULL) in the callers' camera order: M with M M^T = S (the lower-triangular factor whenever the
 * solver kept the natural camera order: every dense system), M^-1, S^-1.  cholesky returns 0.0 / 1.0 like SPDinv. */
double psba_cholesky(psba_ctx *ctx, int matSize, double *outMat);
double psba_trigMat_inv(psba_ctx *ctx, int matSize, double *outMat);
void psba_trigMat_mul(psba_ctx *ctx, int matSize, double *outMat);
/* get_delta_beta / compute_cholmod_E, PSBA/cl_cholmod.h:14-19 (cl_cholmod.cpp:109-167, 176-202): delta and beta of the S of the
 * last compute_S; E of the last cholmod_blk */
void psba_get_delta_beta(psba_ctx *ctx, int matSize, double *delta, double *beta);
void psba_compute_cholmod_E(psba_ctx *ctx, int matSize, double *Eout);
/* matVec_mul, PSBA/cl_linearalg.cpp:19-55: dp[0..N) = S^-1 * eab[0..N) (two triangular solves) */
void psba_matVec_mul(psba_ctx *ctx, int mat_rsize, int mat_csize, double *out);
/* compute_eb / compute_dpb, sba_func.cpp:1001-1062, 1067-1117 */
void psba_compute_eb(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *eab);
void psba_compute_dpb(psba_ctx *ctx, int cnp, int pnp, int nCams, int n3Dpts, double *dp);
/* compute_newp / update_p, sba_func.cpp:1122-1164, 1170-1215 */
void psba_compute_newp(psba_ctx *ctx, int nCamParas, int n3DptsParas, double *new_p);
void psba_update_p(psba_ctx *ctx, int nCamParas, int n3DptsParas, double *p);
/* compute_Jmultiply, sba_func.cpp:19-75 -> compute_Jmultiply.cl:6-52.  x selects g_buffer or
 * dp_buffer.  out (if non-NULL) receives the 2 entries per OBSERVATION (2*n2Dprojs doubles,
 * observation order) -- the non-zero entries of the reference's dense 2*m*n vector.
 * Returns sum((Jx)^2). */
double psba_compute_Jmultiply(psba_ctx *ctx, int mnp, int n3Dpts, int nCams, int n2Dprojs, int x, double *out);
/* clEnqueueWriteBuffer(dp_buffer, P) of trust_region.cpp:166-184 */
void psba_upload_vec(psba_ctx *ctx, int vec, const double *host, int n);
/* cholmod_blk + get_delta_beta + compute_cholmod_E, PSBA/cl_cholmod.cpp:25-202 ->
 * CL_files/cholmod_blk.cl:87-847.  Runs on the S assembled by the last psba_compute_S.
 * E (N doubles, may be NULL) receives E_i; returns sum_i E_i (trust_region.cpp:358-364). */
double psba_cholmod_blk(psba_ctx *ctx, int matSize, double *E, double *delta, double *beta, int *n_scalar_blocks);

/* the same on a caller-supplied HOST matrix (the reference's cholmod_blk takes matBuf itself, PSBA/cl_cholmod.h:10-12):
 * mat (matSize x matSize, symmetric, row-major, matSize % 3 == 0) is replaced by the factor L (zero above the diagonal). */
double psba_cholmod_blk_mat(psba_ctx *ctx, int matSize, double *mat, double *E, double *delta, double *beta, int *n_scalar_blocks);

/* ------------------------------------------------------------------ drivers (L3) ------ */

/* levmar, PSBA/levmar.cpp:45-256 ; trust_region, PSBA/trust_region.cpp:49-288 (blk_idx is not
 * needed and may be NULL).  Fused device-resident implementation of the same control flow. */
int psba_levmar(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *finalErr);
int psba_trust_region(psba_ctx *ctx, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs,
                      int *blk_idx, double *finalErr);
/* the LM <-> TR alternation of PSBA/main.cpp:192-209; returns the final ITER_* flag */
int psba_solve(psba_ctx *ctx, double *initErr, double *finalErr, int *itno);

/* one line of the reference's run log (levmar.cpp:197, trust_region.cpp:250) */
typedef struct {
    int phase;       /* 0 LM try, 1 TR radius try, 2 TR cholmod event */
    int itno;
    double err, rho, mu, delta, pnorm;
    int accepted;
} psba_trace_rec;
int  psba_trace_count(psba_ctx *ctx);
void psba_trace_get(psba_ctx *ctx, int k, psba_trace_rec *rec);
/* options: "verbose" (0/1), "max_iter" (50), "itno", "lm_only" (stop instead of handing to TR),
 * "camera_solver" (0: tiled Cholesky, the default and the path compared with the reference; 1: block-Jacobi preconditioned
 * conjugate gradients on the tiles of S) with "pcg_tol" (1e-10, relative residual) and "pcg_max_iter" (1000),
 * "tr_fused" (1: psba_trust_region takes the scalars of a step -- pUpU, pUg, pBpB, pBg, |P|, the dog-leg quadratic, g.P, |JP|^2,
 * PSBA/trust_region.cpp:125-130,166-176,208-212,520-595 -- from six inner products of g and P_B computed once per step, one host
 * round trip per step and one per radius try; 0: every scalar from the explicit vectors, as the reference does),
 * "seq_graphs" (-1 auto: on one GPU and up to 2 M observations; 0 / 1: an LM try, a linearisation, a trust-region step and a radius try --
 * chains of launches without a host round trip -- run as CUDA graphs from their third use on; the damping term travels in device memory),
 * "trace_reset" (forget the run log), "stats_reset", "profile" (per-kernel CUDA-event timing), "timer_start"; unknown names abort. */
void psba_set_option(psba_ctx *ctx, const char *name, double value);
double psba_get_stat(psba_ctx *ctx, const char *name);
/* index-structure readback (test hook, like the `out` pointers of the operators): copies the named
 * device-built table to `out` (at most max_count 64-bit-or-narrower elements) and returns its length.
 * int32: "pt_ptr" (n+1), "cam_obs" (o), "pair_k"/"pair_l" (n_pairs), "tri_oa"/"tri_ob"/"tri_pt" (ntriples),
 * "pchunk_pair" (n_pchunk), "cam2pos" (m); int64: "pchunk_beg"/"pchunk_end" (n_pchunk).  These replace
 * blk_idx / comm3DIdx / comm3DIdxCnt of generate_idxs (PSBA/misc.cpp:178-218). */
long long psba_get_index(psba_ctx *ctx, const char *name, void *out, long long max_count);
/* camera-system plan readback (test hook; HOST ONLY -- needs no GPU and no context).  Runs the set-up's host stage for the
 * camera system of `nCams` cameras whose coupled pairs are (pair_k[q], pair_l[q]), q < npairs (either triangle; the diagonal
 * is implied): nested-dissection ordering of the 8-camera tiles, symbolic factorisation, step schedule and task lists of the
 * tiled factorisation that replaces SPDinv / cholmod_blk (PSBA/cl_spdinv.cpp:57-103, CL_files/SPD_inv.cl:20-239) -- exactly
 * the tables fill_idxBuffer uploads.  psba_plan_get copies the named int32 table to `out` (at most max_count elements; out may be
 * NULL) and returns its length, -1 for an unknown name:
 *   "stats" = {nt, n_steps, n_tiles_S, n_tiles, one panel per step?}, "cam2pos" (m), "tile_index" (nt*nt, slot of factor tile
 *   (I,J), I >= J in the solver's order, -1 = structurally zero), "step_panels" + "step_panel_ptr" (panels of every step),
 *   "crit_I"/"crit_K" + "step_crit_ptr" (panel CTAs), "psrc_ptr"/"psrc" (source panels a panel's own CTAs apply),
 *   "def_I"/"def_J"/"def_sptr"/"def_src" + "step_def_ptr" (deferred trailing updates: target tile, source panels, step),
 *   "b_J"/"b_sptr"/"b_slot" + "step_b_ptr" (right-hand-side tasks), and the flat records the step kernels load, as consecutive ints:
 *   "crit_desc" (8 per panel CTA: I, K, slot(I,K), slot(K,K), first source, end source, 0, 0), "crit_src" (2 per source: slot(K,P),
 *   slot(I,P) or -1), "def_desc" (4 per deferred task: slot(I,J), first source, end source, 0), "def_srcs" (2 per source: slot(I,P),
 *   slot(J,P)); "bw_order" (panel of every CTA of the backward solve) with "coltile_ptr"/"coltile_row"/"coltile_slot" (tile column of a
 *   panel).  tests/test_tile_plan_cpu.py replays the plan in numpy. */
void *psba_plan_open(int nCams, long long npairs, const int *pair_k, const int *pair_l);
long long psba_plan_get(void *plan, const char *name, int *out, long long max_count);
void psba_plan_close(void *plan);
/* lambda-follow hook for parity runs: the k-th modified-Cholesky event uses lam[k] */
void psba_force_lambda(psba_ctx *ctx, const double *lam, int n);
/* copy current parameters to the host: cams[m*6], pts[n*3] (either may be NULL) */
void psba_get_params(psba_ctx *ctx, int params, double *cams, double *pts);

/* restart from host parameters: cams[m*6], pts[n3Dpts*3] (global arrays; either may be NULL) */
void psba_set_params(psba_ctx *ctx, const double *cams, const double *pts);

/* fused hot-path steps used by the drivers and by bench.py (device-resident, no host arrays):
 * linearise at the current parameters; one damped solve + candidate evaluation. */
typedef struct {
    double cost_new;      /* ||e(p+dp)||^2                              */
    double dp_L2;         /* ||dp||^2                                   */
    double dp_dot;        /* sum dp_i (mu dp_i + g_i)   (levmar.cpp:271) */
    double solve_status;  /* 0.0 ok, 1.0 S not positive definite        */
    double p_new_L2;      /* ||p + dp||^2 (levmar.cpp:214: p_L2 after an accepted step) */
} psba_try_result;
void psba_linearize(psba_ctx *ctx, double coeff_uvw, double coeff_g);
void psba_try_step(psba_ctx *ctx, double mu, psba_try_result *res);

/* ------------------------------------------------------------------ I/O (L0) ---------- */

/* readInitialSBAEstimate, PSBA/readparams.cpp:444-518, with quat2vec (PSBA/misc.cpp:21-49) as
 * input filter and the split of PSBA/main.cpp:131-149.  origin_cnp = 11 (K|q|t files),
 * 16 (K|kc|q|t, kc dropped) or 6 (q|t, K taken from Kdefault[5]).  Outputs are malloc'd:
 * Kparas[m*5], initrot[m*4], camsEx[m*6], pts[n*3], impts[o*2], iidx[o], jidx[o].
 * Returns 0 on success (nonzero + message on a malformed file; the reference exits). */
int psba_readInitialSBAEstimate(const char *camsfname, const char *ptsfname, int origin_cnp,
                                const double *Kdefault,
                                int *ncams, int *n3Dpts, int *n2Dprojs,
                                double **Kparas, double **initrot, double **camsEx, double **pts,
                                double **imgpts, int **iidx, int **jidx);
/* the same reader, also returning what the reference parses and then drops: kc[m*5] (origin_cnp = 16, else NULL) for
 * psba_set_distortion and the image-point covariances cov[o*covsz] (covsz 4 = FULLCOV, 3 = TRICOV, PSBA/readparams.cpp:272-283;
 * NULL if the points file has none) for psba_set_covariances */
int psba_readInitialSBAEstimate_ext(const char *camsfname, const char *ptsfname, int origin_cnp,
                                    const double *Kdefault,
                                    int *ncams, int *n3Dpts, int *n2Dprojs,
                                    double **Kparas, double **initrot, double **camsEx, double **pts,
                                    double **imgpts, int **iidx, int **jidx, double **kc, double **cov, int *covsz);
void psba_quat2vec(const double *inp, int nin, double *outp, int nout);   /* misc.cpp:21-49 */
void psba_free(void *p);

/* Result writers (the reference has only commented-out prototypes, PSBA/readparams.h:13-25,
 * PSBA/misc.cpp:60-85).  vec2quat: q = q_local(v) (x) q_init, the inverse of the set-up of PSBA/main.cpp:131-149.
 * write_sba_result writes the 12-column camera file and the points file that psba_readInitialSBAEstimate
 * (origin_cnp = 11) reads back; write_ply writes points (white) and camera centres (red).  Return 0 on success. */
/* Native BAL loader: problem-*.txt (Rodrigues r, t, f, k1, k2 per camera; p = -P/P.z) converted to
 * PSBA's form: R' = diag(1,-1,-1) R, t' likewise, image y flipped, K = {f, 0, 0, 1, 0}; observations sorted point-major
 * with ascending cameras.  kc[m*2] receives (k1, k2), which the reference's model ignores; may be NULL. */
int psba_read_bal(const char *fname, int *ncams, int *n3Dpts, int *n2Dprojs, double **Kparas, double **initrot,
                  double **camsEx, double **pts, double **imgpts, int **iidx, int **jidx, double **kc);
void psba_vec2quat(const double *initrot4, const double *local3, double *q4);
int psba_write_sba_result(const char *camsfname, const char *ptsfname, int ncams, int n3Dpts, int n2Dprojs,
                          const double *Kparas, const double *initrot, const double *camsEx, const double *pts,
                          const double *imgpts, const int *iidx, const int *jidx);
int psba_write_ply(const char *fname, int ncams, int n3Dpts, const double *initrot, const double *camsEx, const double *pts);

/* ------------------------------------------------------------------ multi-GPU --------- */

/* One process per GPU.  unique_id is the 128-byte ncclUniqueId created by rank 0
 * (psba_comm_unique_id) and distributed by the launcher (torch.distributed / MPI / file). */
void psba_comm_unique_id(char *out128);
void psba_comm_init(int rank, int nranks, const char *unique_id128);
void psba_comm_finalize(void);
/* contiguous point range [p0,p1) and observation range [o0,o1) owned by `rank` for a problem
 * whose per-point observation counts are given by iidx (balanced by observation count) */
void psba_local_range(int n3Dpts, int n2Dprojs, const int *iidx, int rank, int nranks,
                      int *p0, int *p1, int *o0, int *o1);

/* Page-locked host memory for the arrays handed to fill_initBuffer2 / fill_idxBuffer / get_params: the reference's
 * reader mallocs them (PSBA/readparams.cpp:470-500, emalloc in PSBA/misc.cpp:135-146); uploads from pageable memory run
 * at ~5 GB/s, from page-locked memory at the PCIe rate.  Any host pointer is accepted by the ABI; these two are a
 * convenience for callers that own the allocation. */
void *psba_host_alloc(size_t bytes);
void psba_host_free(void *p);

const char *psba_version(void);

#ifdef __cplusplus
}
#endif
#endif
