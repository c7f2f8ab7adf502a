// psba_internal.h -- internal state of libpsba_b200 (not part of the ABI).
//
// Data layout in HBM (all FP64 / int32, device-resident for the whole solve):
//   per camera  : K[5], initcams[4], cams[2 sets][6], camcache[2 sets][CAMC]  (replicated on every GPU)
//   per point   : pts[2 sets][3], V[6] (packed symmetric), Vinv[6], gb = g[N+3i..], dpb = dp[N+3i..]
//   per obs     : impts[2], iidx, jidx, W[18]  -- point-major, cameras ascending (misc.cpp:189-217)
//   camera sys  : S as a pool of 48x48 lower tiles (only tiles of the symbolic Cholesky factor),
//                 ea / dpa in eab[0..N) / dp[0..N)
// The two parameter sets (current / candidate) are swapped by pointer, never copied.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <map>
#include "../../include/psba_b200.h"

#define PSBA_CNP 6
#define PSBA_PNP 3
#define PSBA_MNP 2
#define CAMC 24            // doubles per camera-cache entry: q(4) t(3) K(5) dq/dv(12) = 192 B
#define TS 48              // tile size of the camera system (8 cameras)
#ifndef PT_CTA
#define PT_CTA 64          // observations per point-major CTA wave (measured 128 / 96 / 64 / 32: point pass 0.297 / 0.290 / 0.287 / 0.286 ms,
                           // back-substitution 0.283 / 0.263 / 0.259 / 0.271 ms: smaller CTAs, cheaper barriers, same warps per SM)
#endif
#define CAM_CTA 128        // threads per camera-major CTA (the block reduction of the pass assumes four warps)
#define CAM_OPT 4          // observations per thread in the camera-major pass
#define PAIR_CTA 128       // threads per pair-pass CTA
#define PAIR_TPL 24        // target triples per lane in the pair pass (lane-per-triple variant)
#define SEG_V_MAX 1400     // visits per segment: the Y tile (144 B per visit) of two resident CTAs fits one SM
#define SEG_MAXD 384       // chunk descriptors of a segment prefetched into shared memory
#define NSCAL 48           // size of the device scalar block ([16..) : the scalars of a fused trust-region step)

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "psba_b200: CUDA error %d (%s) at %s(%d)\n", (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); \
    exit(EXIT_FAILURE); } } while (0)

// every kernel launch is followed by this check (launch-configuration errors -- too much shared memory, too many
// registers for the block size -- are reported by cudaGetLastError, not by the launch statement itself)
#define LAUNCH_CHECK() CUDA_CHECK(cudaGetLastError())

struct psba_comm;   // NCCL communicator wrapper (comm.cu)

// per-kernel device timing (CUDA events on the launching stream), enabled by option "profile"
enum psba_kid { KID_CAM_PREP = 0, KID_COST, KID_LIN_POINTS, KID_LIN_CAMS, KID_CAM_REDUCE, KID_VINV, KID_MEMSET_S,
                KID_SCHUR_PAIRS, KID_S_FINALIZE, KID_FACTOR, KID_TRI_SOLVE, KID_NEWCAMS, KID_BACKSUB, KID_REDUCE,
                KID_JDOT, KID_VEC, KID_CHOLMOD, KID_ALLREDUCE, KID_COUNT };
static const char *const psba_kid_name[KID_COUNT] = {
    "k_cam_prep", "k_cost", "k_lin_points", "k_lin_cams", "k_cam_reduce", "k_vinv", "memset_S", "k_schur_pairs",
    "k_S_finalize", "chol_graph", "k_tri_solve", "k_newcams", "k_backsub", "k_reduce", "k_Jdot", "k_vec", "k_cholmod",
    "allreduce" };
struct psba_prof_rec { int id; cudaEvent_t e0, e1; };
// extended camera model (dev_math.cuh): per-camera distortion kc[m*5], per-observation residual weights wgt[o*3]; null = absent
struct psba_ext { const double *kc; const double *wgt; };

struct psba_ctx {
    // global sizes
    int m, n_glob, o_glob, N, T_glob;
    // local (this rank's) point / observation range; equal to global on one GPU
    int n, o, p_off, o_off;
    int rank, nranks;
    cudaStream_t stream;
    cudaStream_t stream2; cudaEvent_t ev_fork, ev_join;   // side stream of the linearisation (camera pass under the point pass)

    // ---- parameters
    double *K, *initcams, *impts;
    double *stage_impts, *stage_pts;  // whole-problem uploads of fill_initBuffer2 until fill_idxBuffer slices them
    double *cams[2], *pts[2], *camcache[2];
    psba_ext ext; bool ext_on;      // extended camera model (psba_set_distortion / psba_set_covariances); off by default
    int cur;                        // index of the "current" parameter set
    bool cache_valid[2];
    // ---- structure
    int *iidx, *jidx;               // local obs -> LOCAL point id, camera id
    int *pt_ptr;                    // n+1
    int *ptchunk; int n_ptchunk;    int4 *ptdesc;    // point-major CTA chunks (point boundaries)
    int n_small, n_big; int *d_small_list, *d_big_list;   // one-wave chunks (null list = all) / oversize chunks
    int n_sm;                       // multiprocessors of the device (persistent grids)
    int *cam_obs;                   // o: local obs ids in camera-major order (ascending point)
    int *cam_pt; double *cam_impts; // camera-major copies of the point index and of the measurements
    int *cchunk_cam, *cchunk_beg, *cchunk_end; int n_cchunk;   // camera-major chunks
    int *cam_cchunk_ptr;            // m+1: chunk range of each camera
    // pair structure (lower triangle k>=l), triples sorted by (k,l), ascending point
    long long ntri;
    int *tri_oa, *tri_ob, *tri_pt;   // observation of camera k, of camera l, local point
    int n_pair; int *pair_k, *pair_l;          // pair blocks present GLOBALLY (all ranks agree)
    int *pair_chunk_ptr;                        // n_pair+1
    int pair_G;                                 // lanes per chunk in the pair pass (1..32)
    int pair_mode;                              // 6: ring kernel, 5: segment kernel, 0: pair-major gather kernel of round 1
    int n_pchunk; int *pchunk_pair; long long *pchunk_beg, *pchunk_end;      // mode 0: fixed-size chunks of the pair runs
    // ---- segment kernel (k_schur_segs): CTA = <= seg_v consecutive visits of one camera row; chunk = the triples of one
    // camera pair inside one segment (pair-major ids: the partial slots k_S_finalize sums per pair, in segment order)
    int seg_v, n_seg, seg_cfg;
    void *seg_desc;                 // n_seg x {row, v0, v1, diag chunk, first / last+1 position in sched_chunk}
    int *sched_chunk;               // off-diagonal chunk ids segment by segment, largest first
    int *sch_beg, *sch_end;         // triple range of every chunk
    unsigned short *tri_vr;         // per triple: rank of the visit (observation of camera k) inside its segment
    // ---- ring kernel (k_schur_ring, pair_mode 6): the off-diagonal triples of a segment as ROWS of 32 lane slots, the rows of
    // one warp contiguous; a task = consecutive rows that sum the chunks of 32/G pairs with G lanes each
    int ring_nw, ring_stages, ring_rt, ring_cfg;   // warps per CTA, stages per warp, target rows per task, launch shape
    long long ring_n_rows;
    int *ring_wrow_ptr;             // n_seg * ring_nw + 1: first row of (segment, warp)
    int2 *ring_rows;                // n_rows * 32: {observation of camera l (-1: empty slot), rank of the visit in the segment}
    int2 *ring_info;                // n_rows: {log2 G | last row of its task << 8 | chunks of the task << 16, position of the task's first chunk in sched_chunk}
    // ---- linearisation products
    double *W, *V, *Vinv, *U, *g, *UVdiag_scr;
    double coeff_uvw, coeff_g;
    bool lin_valid;
    double *cam_part;               // n_cchunk * 27
    double *pair_part;              // n_pchunk * 42
    // ---- camera system
    int nt;                         // tiles per dimension
    int n_tiles; int *tile_index;   // nt*nt -> slot or -1 (device + host copy)
    int n_tiles_S;                  // the first n_tiles_S slots are the tiles of S itself, the rest is fill-in of the factor
    std::vector<int> h_tile_index;
    double *Stiles;                 // n_tiles * TS*TS  (factor overwrites it)
    double *Linv;                   // nt * TS*TS   inverse of the diagonal factor tiles
    double *eab, *dp;               // T_loc-sized vectors laid out [N | 3n]
    int *d_status;                  // device int: 0 ok, 1 not PD
    // camera ordering of S (nested dissection at tile granularity): cam2pos[j] = block position of camera j,
    // pos2cam[p] = camera at block position p (-1: padding).  tile_index is in permuted coordinates.
    int *cam2pos, *pos2cam;
    std::vector<int> h_cam2pos;
    // step schedule of the factorisation: panels whose dependencies are met run in the same launch.
    //   critical task (I,K): tile row I of panel K (I == K: the diagonal CTA)
    //   panel sources      : panels P of the previous step with a tile (K,P) (their updates are still pending)
    //   rhs task J         : b_J -= sum_P L_JP y_P over the panels P of the previous step (J in a later step)
    //   deferred task (I,J): trailing tile touched by panels of the previous step, J in a later step
    bool chol_pdl;                             // programmatic dependent launch between the step kernels
    int n_steps; bool chain_schedule;          // chain: one panel per step (dense S)
    double chol_flops;                         // algorithmic FP64 flops of one factorisation + solves (symbolic factor)
    std::vector<int> step_crit_ptr, step_def_ptr, step_panel_ptr, step_b_ptr;
    int *d_crit_I, *d_crit_K, *d_psrc_ptr, *d_psrc, *d_b_J, *d_b_sptr, *d_b_slot;
    int *d_def_I, *d_def_J, *d_def_sptr, *d_def_src, *d_step_panels;
    int4 *d_crit_desc, *d_def_desc; int2 *d_crit_src, *d_def_srcs;   // flat task descriptors (one load per CTA)
    void *tile_block; size_t tile_block_bytes;   // the schedule tables above live in ONE device block (psba_flush_tile_uploads)
    // dataflow factorisation (k_panel_flow): task table in step order, per-tile write counters and what a task waits for
    int n_flow_tasks; int2 *d_flow_tasks; int *d_flow_final, *d_flow_defseq, *d_flow_bseq, *d_flow_critneed, *d_flow_ver;
    int camera_solver; double *pcg_work; double pcg_tol; int pcg_max_iter, pcg_last_iters;   // optional PCG camera solve (kernels_pcg.cu)
    bool chol_flow;                 // one flag-driven launch instead of one kernel per step (PSBA_CHOL_FLOW=0 restores the step kernels)
    double *contrib;                // n_tiles * TS: L_IK y_K per factor tile
    double *Ldiag;                  // nt * TS*TS   factor of the diagonal tiles (kept out of the tile pool)
    int *d_coltile_ptr, *d_coltile_row, *d_coltile_slot;   // CSC of the factor tiles (backward solve)
    cudaGraphExec_t bw_graph; bool bw_graph_ok;
    int *d_bw_order; double *bw_xbuf;           // dataflow backward solve: panel order, solution buffer in the solver's ordering (sentinel = not there yet)
    cudaGraphExec_t chol_graph; bool chol_graph_ok;
    bool S_valid, factor_valid;
    double *Sdense, *Sdense_aux;    // N*N, only allocated on demand (compat / cholmod)
    double *chol_aux, *chol_diag, *chol_E;
    // ---- scalars
    double *d_part;                 // per-chunk partial sums (n_ptchunk * 4)
    double *d_scal; double *h_scal; // NSCAL doubles (h_scal pinned)
    // ---- device-resident scalars of a solve and whole-sequence CUDA graphs (psba_api.cu: psba_set_scalars, psba_seq_begin/end)
    double *d_mu;                   // device: {damping term, step coefficient a, step coefficient b, -}
    double *h_mu_ring; int h_mu_next;   // pinned: 64 slots of 4 doubles; slots 0..31 belong to captured sequences, the rest rotate
    int use_graphs;                 // -1 auto (small problems on one GPU), 0 off, 1 on
    struct seq_graph { cudaGraphExec_t exec; int seen, slot; bool ok; double d_launches, d_tries, d_exqt, d_lin; bool cv0, cv1, S_valid, factor_valid, lin_valid; double cu, cg; };
    std::map<unsigned long long, seq_graph> *seqs;
    seq_graph *capturing;           // sequence being captured (its scalar slot is the source of the captured copies)
    double snap_launches, snap_tries, snap_exqt, snap_lin;
    int tr_fused;                   // trust region: scalars of a step from six inner products, one read-back (host_drivers.cpp); 0: explicit vectors
    int *h_status;                  // pinned copy of d_status[0] read with the step scalars
    // ---- TR vectors (local layout [N | 3n])
    double *P_U, *P_B, *P;
    // ---- compat state
    double mu_pending;              // update_UV / restore_UVdiag
    double *tmpA, *tmpB;            // on-demand o*12 / o*18 scratch for materialised J / Y
    // ---- driver state (globals of PSBA/main.cpp:22-37)
    int itno, max_iter, verbose, lm_only;
    double initErr;
    std::vector<psba_trace_rec> trace;
    std::vector<double> force_lambda; int n_cholmod_events;
    double cholmod_max_l_over_beta;   // tile-pool modified Cholesky: largest factor entry / beta of the last run
    // stats
    double st_tries, st_exqt, st_lin, st_launches, st_seq_replays;
    bool profile;
    std::vector<psba_prof_rec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[KID_COUNT]; double prof_n[KID_COUNT];
    cudaEvent_t timer_e0, timer_e1; bool timer_init;
    psba_comm *comm;
};

// ---- psba_api.cu: stream-ordered pool allocator
void *psba_dev_alloc(psba_ctx *c, size_t bytes, bool zero);
void psba_dev_free(psba_ctx *c, void *p);
// ---- structure.cu
void psba_build_structure(psba_ctx *c, const int *iidx_host, const int *jidx_host);
void psba_build_camera_major_copies(psba_ctx *c);   // cam_pt, cam_impts (needs impts on the device)
// ---- kernels_obs.cu
void psba_launch_cam_prep(psba_ctx *c, int set);
double psba_launch_cost(psba_ctx *c, int set, double *ex_dev /*may be null*/);
void psba_launch_linearize(psba_ctx *c, double coeff_uvw, double coeff_g);
void psba_launch_jac_materialize(psba_ctx *c, double *JA, double *JB);
void psba_launch_Jdot(psba_ctx *c, const double *x, const double *y, double *Jx_out /*may be null*/, double res[3]);
// ---- kernels_schur.cu
double psba_launch_vinv(psba_ctx *c, double mu);
void psba_launch_schur(psba_ctx *c, double mu);
size_t psba_ring_smem(int cfg, int seg_v);   // dynamic shared memory of k_schur_ring (structure.cu)
void psba_launch_Y_materialize(psba_ctx *c, double *Y);
// ---- kernels_solve.cu
void psba_build_tile_structure(psba_ctx *c, const std::vector<std::pair<int,int>> &camera_pairs);   // host only; then:
void psba_flush_tile_uploads(psba_ctx *c);       // uploads / allocations the symbolic part recorded
double psba_launch_factor(psba_ctx *c, bool defer_status = false);      // returns 0.0 / 1.0 (syncs unless deferred)
void psba_launch_solve(psba_ctx *c);         // dp[0..N) = S^-1 eab[0..N)
void psba_tiles_to_dense(psba_ctx *c, double *dense_dev, bool mirror);
void psba_launch_explicit_inverse(psba_ctx *c, double *out_dev);
void psba_launch_factor_products(psba_ctx *c, double *out_dev, int what);   // 0 S^-1, 1 L^-1, 2 L (callers' camera order)
void psba_tile_delta_beta(psba_ctx *c, double *delta, double *beta);
double psba_launch_cholmod(psba_ctx *c, double *delta, double *beta, int *nscalar);
double psba_launch_cholmod_tiles(psba_ctx *c, double *delta, double *beta, int *nmod, double *E_host, double *max_l_over_beta);
bool psba_cholmod_use_tiles(psba_ctx *c);
bool psba_cholmod_dense_possible(psba_ctx *c);
double psba_launch_cholmod_dense(psba_ctx *c, int N, double *mat, double *aux, double *diagInv, double *E,
                                 double *delta, double *beta, int *nscalar);
// ---- kernels_pcg.cu
int psba_launch_pcg(psba_ctx *c);            // dp[0..N) = S^-1 eab[0..N) by block-Jacobi PCG; status word as the factorisation
// ---- kernels_backsub.cu
void psba_launch_backsub(psba_ctx *c, double mu, bool evaluate, psba_try_result *res);
void psba_enqueue_backsub(psba_ctx *c, double mu, bool evaluate);   // no host round trip; evaluate: scalars + status copies enqueued
void psba_finish_try(psba_ctx *c, psba_try_result *res);            // synchronise, read the step scalars
// damping term / step coefficients into device memory (kernels read c->d_mu): a 32-byte copy from a pinned slot
void psba_set_scalars(psba_ctx *c, double mu, double a, double b);
// A SEQUENCE is a fixed chain of launches without a host round trip (an LM try, a linearisation, a trust-region step, a radius
// try).  psba_seq_begin returns true when the caller has to enqueue the chain (first sighting: plainly; second: under stream
// capture) and false when an instantiated graph of the chain has been launched instead; psba_seq_end closes a begun chain.
// `kind` + the host state the chain depends on (parameter set, valid camera caches, camera model) is the key of the graph.
bool psba_seq_begin(psba_ctx *c, int kind, double mu, double a, double b);
void psba_seq_end(psba_ctx *c);
enum { SEQ_TRY = 1, SEQ_LIN_LM = 2, SEQ_LIN_TR = 3, SEQ_TR_STEP = 4, SEQ_TR_RADIUS = 5, SEQ_COST = 6 };
double psba_enqueue_cost(psba_ctx *c, int set, double *ex_dev);     // cost kernels + reduction into d_scal[0]; no read-back (returns 0)
void psba_launch_newp(psba_ctx *c);
void psba_launch_step_newp(psba_ctx *c, double a, const double *x, double b, const double *y);   // dp = a x + b y and p + dp -> candidate set
// ---- vector helpers (kernels_backsub.cu)
void psba_launch_dots(psba_ctx *c, const double *x, const double *y, const double *z, double out[6]);
// the same sums left in d_scal[off..off+12) (camera part, point part) / d_scal[off..off+3): no read-back, no host round trip
void psba_enqueue_dots(psba_ctx *c, const double *x, const double *y, const double *z, int off);
void psba_enqueue_Jdot(psba_ctx *c, const double *x, const double *y, double *Jx_out, int off);
void psba_launch_axpby(psba_ctx *c, double a, const double *x, double b, const double *y, double *out);
double psba_launch_maxdiag(psba_ctx *c);
// ---- comm.cu
void psba_allreduce_sum(psba_ctx *c, double *buf, size_t count);
void psba_allreduce_max(psba_ctx *c, double *buf, size_t count);
bool psba_comm_active();
int psba_comm_rank();
int psba_comm_size();

// scoped kernel timer: PROF(c, KID_x) { launch...; }
struct psba_prof_scope {
    psba_ctx *c; int id; cudaEvent_t e0, e1; bool on;
    psba_prof_scope(psba_ctx *c_, int id_) : c(c_), id(id_), on(c_->profile) {
        if (!on) return;
        auto get = [&]() { cudaEvent_t e; if (c->prof_pool.empty()) { CUDA_CHECK(cudaEventCreate(&e)); } else { e = c->prof_pool.back(); c->prof_pool.pop_back(); } return e; };
        e0 = get(); e1 = get();
        CUDA_CHECK(cudaEventRecord(e0, c->stream));
    }
    ~psba_prof_scope() {
        LAUNCH_CHECK();
        if (!on) return;
        CUDA_CHECK(cudaEventRecord(e1, c->stream));
        c->prof_pending.push_back({id, e0, e1});
    }
    explicit operator bool() const { return true; }
};
#define PROF(c, id) if (psba_prof_scope prof_scope_##id{c, id})
void psba_prof_collect(psba_ctx *c);

// opt-in dynamic shared memory is an attribute per (function, device); it is also a LIMIT (a launch that asks for more
// than the value set fails with "invalid argument"), so the largest request seen so far is what is set
#include <map>
#include <utility>
static inline void psba_set_smem(const void *fn, int bytes)
{
    static std::map<std::pair<const void *, int>, int> done;
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    int &cur = done[{fn, dev}];
    if (bytes > cur) { CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); cur = bytes; }
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
