/*
 * psba_oracle.h -- CPU ORACLE for the PSBA bundle-adjustment hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (psba_b200/) never links, imports or calls anything in this directory.
 *
 * It is a plain-C restatement of the algorithm of eglrp/PSBA (reference tree at
 * /root/reference, file:line cited on every function).  Arithmetic is FP64; summation
 * orders follow the reference kernels (ascending point index for per-camera sums,
 * ascending camera index for per-point sums, left-to-right host reductions).  The only
 * structural change is that the dense index tables of the reference
 * (blk_idx[n*m], comm3DIdx[m*m*n], PSBA/misc.cpp:178-218) are replaced by CSR lists that
 * enumerate exactly the same entries in exactly the same order, so that large problems
 * fit in memory.
 *
 * Parity pin: oracle/_ref/libpsba_ref.so compiles the reference's OWN kernel bodies
 * (the .cl files of CL_files/) and loaders (PSBA/readparams.cpp, PSBA/misc.cpp) in place; tests/
 * check this restatement against it stage by stage and against the golden trajectories of
 * SURVEY.md App. B.3 (tests/golden/).  SPD_inv.cl and cholmod_blk.cl use OpenCL-2.0 device
 * enqueue and cannot be compiled by g++; for those two files the restatement in
 * orc_chol.c is the only executable form ("restated, pinned by trajectories only").
 */
#ifndef PSBA_ORACLE_H
#define PSBA_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* iteration flags, PSBA/psba.h:12-18 */
#define ORC_ITER_TURN_TO_LM       1
#define ORC_ITER_TURN_TO_TR       2
#define ORC_ITER_CONTINUE         3
#define ORC_ITER_ERR              4
#define ORC_ITER_DP_NO_CHANGE     5
#define ORC_ITER_ERR_SMALL_ENOUGH 6
#define ORC_ITER_PASS             7

/* constants, PSBA/psba.h:6-10, trust_region.cpp:18 */
#define ORC_INIT_MU      1e-03
#define ORC_STOP_THRESH  1e-12
#define ORC_EPSILON      1e-12
#define ORC_EPSILON2     1e-12
#define ORC_MAX_DELTA    10000.0
#define ORC_MAX_ITER     50

#define ORC_TRACE_MAX    4096

/* one printed line of the reference's run log (levmar.cpp:197, trust_region.cpp:250) */
typedef struct {
    int    phase;     /* 0 = LM try, 1 = TR radius try, 2 = TR cholmod event */
    int    itno;
    double err;       /* new / actual ||e||^2 of the try                       */
    double rho;
    double mu;        /* LM: mu used for the try. TR: lambda                   */
    double delta;     /* TR: radius after the update                           */
    double pnorm;     /* TR: ||p||                                             */
    int    accepted;  /* 1 accepted, 0 rejected / shrunk                       */
} orc_trace_rec;

struct orc_state;

/* operator table = the reference's L2 API (PSBA/sba_func.h:10-138).  Two back ends fill
 * it: the restatement (orc_kernels.c) and the reference's own kernel bodies
 * (ref_shim/ref_kernels.cpp -> oracle/_ref). */
typedef struct {
    void (*exQT)(struct orc_state *s, const double *cams, const double *pts, double *ex);
    void (*jacobiQT)(struct orc_state *s);
    void (*U)(struct orc_state *s, double coeff);
    void (*V)(struct orc_state *s, double coeff);
    void (*Wblks)(struct orc_state *s, double coeff);
    void (*g)(struct orc_state *s, double coeff);
    void (*update_UV)(struct orc_state *s, double mu);
    void (*restore_UVdiag)(struct orc_state *s);
    double (*Vinv)(struct orc_state *s);
    void (*Yblks)(struct orc_state *s);
    void (*S)(struct orc_state *s);
    void (*ea)(struct orc_state *s);
    void (*matVec)(struct orc_state *s);            /* dp[0..N) = S * eab[0..N) */
    void (*eb)(struct orc_state *s);
    void (*dpb)(struct orc_state *s);
    void (*newp)(struct orc_state *s);
    void (*update_p)(struct orc_state *s);
    void (*Jmultiply)(struct orc_state *s, const double *x, double *out); /* out[2*o] */
    const char *name;
} orc_ops;

typedef struct orc_state {
    int m, n, o;            /* nCams, n3Dpts, n2Dprojs */
    int N, T;               /* 6m, 6m+3n */
    /* parameters (cl_psba.cpp:40-84 buffer list) */
    double *K;              /* m*5  fu,u0,v0,ar,s */
    double *impts;          /* o*2 */
    double *initcams;       /* m*4 */
    /* extended camera model (NOT in the reference, whose kernels have neither): fixed
     * per-camera distortion kc[m*5] of the sba varKD model and a per-observation lower-triangular residual
     * weight wgt[o*3] = (w00, w10, w11) with W^T W = inverse image-point covariance.  NULL = absent. */
    double *kc, *wgt;
    double *cams, *newcams; /* m*6 */
    double *pts, *newpts;   /* n*3 */
    /* structure (point-major observation order) */
    int *iidx, *jidx;       /* o */
    int *pt_ptr;            /* n+1: observations of point i are [pt_ptr[i], pt_ptr[i+1]) */
    int *cam_ptr, *cam_obs; /* m+1, o: observations of camera j in ascending point order */
    long long *pair_ptr;    /* m*m+1: triples of camera pair (k,l), ascending point order */
    int *pair_oa, *pair_ob; /* obs id in camera k / camera l for each triple */
    long long ntriples;
    /* dense tables, only allocated for the _ref back end */
    int *blk_idx, *comm3DIdx, *comm3DIdxCnt;
    /* work buffers */
    double *ex, *JA, *JB, *U, *V, *UVdiag, *W, *Y;
    double *S, *Saux, *diagAux, *blkBackup, *E;
    double *g, *dp, *eab, *Jx1, *Jx2;
    double ret;
    /* driver state (globals of PSBA/main.cpp:22-37) */
    int itno;
    double initErr, finalErr;
    int use_explicit_inverse;    /* 1: SPDinv chain as the reference; 0: potrf+potrs (variant P) */
    /* lambda-follow hook: if n_force_lambda>0 the k-th cholmod event returns it */
    double force_lambda[64]; int n_force_lambda; int n_cholmod_events;
    /* trace */
    orc_trace_rec trace[ORC_TRACE_MAX]; int ntrace;
    int verbose;
    int nthreads;                /* OpenMP threads for the NDRange loops (1 = serial) */
    const orc_ops *ops;
    /* timers (seconds) */
    double t_total, t_lin, t_schur, t_solve, t_backsub, t_cost;
    long long n_tries, n_exqt, n_lin;
} orc_state;

/* ---- orc_io.c : PSBA/readparams.cpp + PSBA/misc.cpp:21-49 + PSBA/main.cpp:131-149 ---- */
/* returns 0 on success. Allocates all outputs with malloc. filecnp = origin_cnp+1. */
int orc_read_sba(const char *camsfname, const char *ptsfname, int origin_cnp,
                 int *ncams, int *n3Dpts, int *n2Dprojs,
                 double **motstruct, double **initrot, double **imgpts,
                 int **pt_nframes, int **frames);
void orc_quat2vec(const double *inp, int nin, double *outp, int nout);
/* main.cpp:131-149: zero local rotation, split K | extrinsics. origin_cnp in {6, 11, 16} */
void orc_split_motion(const double *motstruct, int origin_cnp, int m,
                      double *Kparas, double *camsEx);

/* ---- orc_state.c ---- */
void orc_set_ext(orc_state *s, const double *kc /* m*5 or NULL */, const double *wgt /* o*3 or NULL */);
orc_state *orc_create(int m, int n, int o, const double *K, const double *impts,
                      const double *initcams, const double *camsEx, const double *pts,
                      const int *iidx, const int *jidx, int want_dense_tables);
void orc_destroy(orc_state *s);
void orc_use_ops(orc_state *s, const orc_ops *ops);
const orc_ops *orc_native_ops(void);
/* generate_idxs (misc.cpp:178-218) from per-point frame lists (file order); fills iidx/jidx */
void orc_generate_idxs(int m, int n, int o, const int *pt_nframes, const int *frames,
                       int *iidx, int *jidx);

/* ---- orc_chol.c : CL_files/SPD_inv.cl, cholmod_blk.cl, PSBA/cl_spdinv.cpp, cl_cholmod.cpp ---- */
double orc_cholesky(double *mat, double *diagInv, int N);
void   orc_trigMat_inv(double *mat, const double *diagInv, int N);
void   orc_trigMat_mul(double *mat, double *diag, int N);
double orc_SPDinv(double *mat, double *diagAux, int N);
void   orc_get_delta_beta(const double *mat, int N, double *delta, double *beta);
void   orc_cholmod_blk(double *mat, double *aux, double *diagInv, double *diag, int N,
                       double beta, double delta, int *n_scalar_blocks);
void   orc_cholmod_E(const double *mat, double *diag, int N);
/* variant P: plain potrf + potrs, returns 0.0 ok / 1.0 not PD */
double orc_potrf_solve(double *mat, const double *rhs, double *x, int N);

/* ---- orc_driver.c : PSBA/levmar.cpp, trust_region.cpp, main.cpp:192-209 ---- */
double orc_L2_sq(int n, const double *x);                 /* misc.cpp:151-157 */
double orc_dot(const double *a, const double *b, int n);  /* misc.cpp:162-169 */
int orc_levmar(orc_state *s);
int orc_trust_region(orc_state *s);
int orc_solve(orc_state *s);       /* while(true){levmar; trust_region;} returns final flag */

#ifdef __cplusplus
}
#endif
#endif
