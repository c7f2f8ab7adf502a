// psba_api.cu -- context life cycle, structure build and the operator-level C ABI
// (include/psba_b200.h).  Mirrors PSBA/cl_psba.cpp (setup_cl / fill_initBuffer2 / fill_idxBuffer /
// release_buffer) and the wrappers of PSBA/sba_func.cpp, cl_spdinv.cpp, cl_cholmod.cpp,
// cl_linearalg.cpp.  Every operator is synchronous when it returns a host value, as the
// reference's wrappers are (they all end in clFinish).
#include "psba_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <chrono>

static void die(const char *msg)
{
    fprintf(stderr, "psba_b200: %s\n", msg);
    exit(EXIT_FAILURE);
}

static void check_dims(int cnp, int pnp, int mnp)
{
    if (cnp != PSBA_CNP || pnp != PSBA_PNP || mnp != PSBA_MNP)
        die("only cnp=6, pnp=3, mnp=2 are supported (CL_files/PSBA.cl:5-7)");
}

// Device memory comes from the stream-ordered allocator; the pool keeps what a released context gives back
// (release threshold = never), so a service that opens one problem after another pays for page mapping once.
void *psba_dev_alloc(psba_ctx *c, size_t bytes, bool zero)
{
    static bool pool_ready[64] = {false};
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    if (!pool_ready[dev & 63]) {
        cudaMemPool_t pool;
        CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep = ~0ull;
        CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        pool_ready[dev & 63] = true;
    }
    void *p = nullptr;
    bytes = std::max<size_t>(bytes, 16);
    CUDA_CHECK(cudaMallocAsync(&p, bytes, c->stream));
    if (zero) CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, c->stream));
    return p;
}
void psba_dev_free(psba_ctx *c, void *p)
{
    if (p) CUDA_CHECK(cudaFreeAsync(p, c->stream));
}
template <class T> static T *dalloc(psba_ctx *c, size_t n)
{
    return (T *)psba_dev_alloc(c, std::max<size_t>(n, 1) * sizeof(T), true);
}

extern "C" void *psba_host_alloc(size_t bytes)
{
    void *p = nullptr;
    CUDA_CHECK(cudaHostAlloc(&p, std::max<size_t>(bytes, 16), cudaHostAllocDefault));
    return p;
}
extern "C" void psba_host_free(void *p) { if (p) CUDA_CHECK(cudaFreeHost(p)); }

extern "C" const char *psba_version(void) { return "psba_b200 0.1 (sm_100a, FP64)"; }

extern "C" psba_ctx *psba_setup_cl(int cnp, int pnp, int mnp, int nCams, int n3Dpts, int n2Dprojs)
{
    check_dims(cnp, pnp, mnp);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) die("no CUDA device: this library has no CPU fallback");
    psba_ctx *c = new psba_ctx();
    c->m = nCams; c->n_glob = n3Dpts; c->o_glob = n2Dprojs;
    c->N = 6 * nCams; c->T_glob = 6 * nCams + 3 * n3Dpts;
    c->n = 0; c->o = 0; c->p_off = 0; c->o_off = 0;
    c->rank = psba_comm_active() ? psba_comm_rank() : 0;
    c->nranks = psba_comm_active() ? psba_comm_size() : 1;
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    c->cur = 0; c->cache_valid[0] = c->cache_valid[1] = false;
    c->ext.kc = c->ext.wgt = nullptr; c->ext_on = false;
    c->lin_valid = false; c->S_valid = false; c->factor_valid = false; c->chol_graph_ok = false; c->bw_graph_ok = false;
    c->Sdense = c->Sdense_aux = nullptr; c->tmpA = c->tmpB = nullptr;
    c->mu_pending = 0.0; c->coeff_uvw = 1.0; c->coeff_g = 1.0;
    c->itno = 0; c->max_iter = 50; c->verbose = 0; c->lm_only = 0; c->initErr = 0.0; c->tr_fused = 1;
    c->n_cholmod_events = 0; c->cholmod_max_l_over_beta = 0.0;
    c->st_tries = c->st_exqt = c->st_lin = c->st_launches = 0;
    c->profile = false; c->timer_init = false;
    for (int k = 0; k < KID_COUNT; ++k) { c->prof_ms[k] = 0; c->prof_n[k] = 0; }
    c->comm = nullptr;
    { int dev = 0; CUDA_CHECK(cudaGetDevice(&dev)); CUDA_CHECK(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, dev)); }
    c->d_small_list = c->d_big_list = nullptr; c->n_small = c->n_big = 0;
    c->n_seg = 0; c->seg_desc = nullptr; c->sched_chunk = nullptr; c->sch_beg = c->sch_end = nullptr; c->tri_vr = nullptr;
    c->ring_wrow_ptr = nullptr; c->ring_rows = nullptr; c->ring_info = nullptr; c->ring_n_rows = 0;
    c->tile_block = nullptr; c->tile_block_bytes = 0;
    c->pchunk_pair = nullptr; c->pchunk_beg = c->pchunk_end = nullptr;
    c->chol_pdl = !(getenv("PSBA_NO_PDL") && atoi(getenv("PSBA_NO_PDL")));
    // dataflow factorisation (one flag-driven launch): measured 1.148 ms against 1.079 ms for the step kernels on the headline
    // workload and 123 against 120 us on a 7-panel chain -- kept as an option, off by default
    c->chol_flow = getenv("PSBA_CHOL_FLOW") && atoi(getenv("PSBA_CHOL_FLOW")) != 0;
    c->pcg_work = nullptr; c->camera_solver = 0; c->pcg_tol = 1e-10; c->pcg_max_iter = 1000; c->pcg_last_iters = 0;
    c->d_flow_tasks = nullptr; c->d_flow_final = c->d_flow_defseq = c->d_flow_bseq = c->d_flow_critneed = c->d_flow_ver = nullptr; c->n_flow_tasks = 0;
    c->stage_impts = c->stage_pts = nullptr;
    c->K = dalloc<double>(c, (size_t)nCams * 5);
    c->initcams = dalloc<double>(c, (size_t)nCams * 4);
    for (int s = 0; s < 2; ++s) {
        c->cams[s] = dalloc<double>(c, (size_t)nCams * 6);
        c->camcache[s] = dalloc<double>(c, (size_t)nCams * (CAMC + 16));   // full entries, then the compact records of the pipelined point pass
    }
    c->U = dalloc<double>(c, (size_t)nCams * 42);          // [U | 6m doubles: ga as it comes out of the camera pass on N > 1 GPUs (one all-reduce for both)]
    c->d_status = dalloc<int>(c, 4);
    c->d_scal = dalloc<double>(c, NSCAL);
    CUDA_CHECK(cudaMallocHost(&c->h_scal, NSCAL * sizeof(double)));
    CUDA_CHECK(cudaMallocHost(&c->h_status, 4 * sizeof(int)));
    c->h_status[0] = 0;
    c->d_mu = dalloc<double>(c, 4);
    CUDA_CHECK(cudaMallocHost(&c->h_mu_ring, 64 * 4 * sizeof(double)));
    c->h_mu_next = 0; c->st_seq_replays = 0; c->capturing = nullptr; c->seqs = new std::map<unsigned long long, psba_ctx::seq_graph>();
    c->use_graphs = getenv("PSBA_SEQ_GRAPHS") ? atoi(getenv("PSBA_SEQ_GRAPHS")) : -1;
    return c;
}

extern "C" void psba_fill_initBuffer2(psba_ctx *c, int cnp, int pnp, int mnp, int nCams, int n3Dpts, int n2Dprojs,
                                      const double *Kparas, const double *impts, const double *initcams,
                                      const double *camsEx, const double *pts3Ds)
{
    check_dims(cnp, pnp, mnp);
    if (nCams != c->m || n3Dpts != c->n_glob || n2Dprojs != c->o_glob) die("fill_initBuffer2: sizes differ from setup_cl");
    // the caller keeps its arrays: everything is copied to the device here (the local slice of the
    // per-observation / per-point arrays is cut out once fill_idxBuffer knows this rank's range)
    cudaStream_t st = c->stream;
    CUDA_CHECK(cudaMemcpyAsync(c->K, Kparas, (size_t)nCams * 5 * 8, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(c->initcams, initcams, (size_t)nCams * 4 * 8, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(c->cams[0], camsEx, (size_t)nCams * 6 * 8, cudaMemcpyHostToDevice, st));
    psba_dev_free(c, c->stage_impts); psba_dev_free(c, c->stage_pts);
    c->stage_impts = (double *)psba_dev_alloc(c, (size_t)n2Dprojs * 16, false);
    c->stage_pts = (double *)psba_dev_alloc(c, (size_t)n3Dpts * 24, false);
    if (n2Dprojs) CUDA_CHECK(cudaMemcpyAsync(c->stage_impts, impts, (size_t)n2Dprojs * 16, cudaMemcpyHostToDevice, st));
    if (n3Dpts) CUDA_CHECK(cudaMemcpyAsync(c->stage_pts, pts3Ds, (size_t)n3Dpts * 24, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
}

extern "C" void psba_local_range(int n, int o, const int *iidx, int rank, int nranks, int *p0, int *p1, int *o0, int *o1)
{
    // contiguous point ranges balanced by observation count: rank r owns the points whose first
    // observation index lies in [r*o/R, (r+1)*o/R)   (host helper; the engine derives the same range
    // from the device-built CSR in structure.cu)
    std::vector<int> ptr((size_t)n + 1, 0);
    for (int k = 0; k < o; ++k) ptr[iidx[k] + 1]++;
    for (int i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
    auto first_pt = [&](int r) -> int {
        if (r >= nranks) return n;
        long long target = (long long)o * r / nranks;
        return (int)(std::lower_bound(ptr.begin(), ptr.begin() + n, (int)target) - ptr.begin());
    };
    *p0 = first_pt(rank); *p1 = first_pt(rank + 1);
    *o0 = ptr[*p0]; *o1 = ptr[*p1];
}

extern "C" void psba_fill_idxBuffer(psba_ctx *c, int nCams, int n3Dpts, int n2Dprojs, const int *iidx, const int *jidx)
{
    if (nCams != c->m || n3Dpts != c->n_glob || n2Dprojs != c->o_glob) die("fill_idxBuffer: sizes differ from setup_cl");
    if (!c->stage_pts) die("fill_idxBuffer called before fill_initBuffer2");
    const bool timing = getenv("PSBA_SETUP_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    // ---- index structure (device-built: CSR, camera order, camera-pair triples, tile structure of S)
    psba_build_structure(c, iidx, jidx);
    const int n = c->n, o = c->o;
    // ---- this rank's slice of the per-observation / per-point inputs
    cudaStream_t st = c->stream;
    if (c->nranks == 1) { c->impts = c->stage_impts; c->pts[0] = c->stage_pts; }
    else {
        c->impts = (double *)psba_dev_alloc(c, (size_t)o * 16, false);
        c->pts[0] = (double *)psba_dev_alloc(c, (size_t)n * 24, false);
        if (o) CUDA_CHECK(cudaMemcpyAsync(c->impts, c->stage_impts + (size_t)c->o_off * 2, (size_t)o * 16, cudaMemcpyDeviceToDevice, st));
        if (n) CUDA_CHECK(cudaMemcpyAsync(c->pts[0], c->stage_pts + (size_t)c->p_off * 3, (size_t)n * 24, cudaMemcpyDeviceToDevice, st));
        psba_dev_free(c, c->stage_impts); psba_dev_free(c, c->stage_pts);
    }
    c->stage_impts = c->stage_pts = nullptr;
    c->pts[1] = dalloc<double>(c, (size_t)n * 3);
    psba_build_camera_major_copies(c);
    // ---- work buffers
    const size_t Tl = (size_t)c->N + 3 * (size_t)n;
    // W, V and Vinv are written in full (point pass, k_vinv) before anything reads them: no zero fill (0.9 GB at 5 M observations)
    c->W = (double *)psba_dev_alloc(c, std::max<size_t>((size_t)o * 18, 1) * sizeof(double), false);
    c->V = (double *)psba_dev_alloc(c, std::max<size_t>((size_t)n * 6, 1) * sizeof(double), false);
    c->Vinv = (double *)psba_dev_alloc(c, std::max<size_t>((size_t)n * 6, 1) * sizeof(double), false);
    c->g = dalloc<double>(c, Tl); c->dp = dalloc<double>(c, Tl); c->eab = dalloc<double>(c, Tl);
    c->P_U = dalloc<double>(c, Tl); c->P_B = dalloc<double>(c, Tl); c->P = dalloc<double>(c, Tl);
    c->cam_part = dalloc<double>(c, (size_t)c->n_cchunk * 27);
    c->pair_part = dalloc<double>(c, (size_t)c->n_pchunk * 42);
    c->d_part = dalloc<double>(c, ((size_t)cdiv(o, 128) + c->n_ptchunk + 512) * 8);
    c->chol_aux = dalloc<double>(c, (size_t)3 * c->N + 2 * TS);
    c->chol_diag = dalloc<double>(c, (size_t)3 * c->N + 2 * TS);
    c->chol_E = dalloc<double>(c, (size_t)c->N + TS);
    c->UVdiag_scr = dalloc<double>(c, Tl);
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (timing)
        fprintf(stderr, "psba setup: fill_idxBuffer total          %8.2f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
}

extern "C" void psba_release_buffer(psba_ctx *c)
{
    if (!c) return;
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    void *ptrs[] = {c->K, c->initcams, c->impts, c->cams[0], c->cams[1], c->pts[0], c->pts[1], c->camcache[0], c->camcache[1],
                    c->iidx, c->jidx, c->pt_ptr, c->ptchunk, c->ptdesc, c->d_small_list, c->d_big_list, c->cam_obs, c->cam_pt, c->cam_impts, c->cchunk_cam, c->cchunk_beg, c->cchunk_end,
                    c->cam_cchunk_ptr, c->tri_oa, c->tri_ob, c->tri_pt, c->pair_k, c->pair_l, c->pair_chunk_ptr, c->pchunk_pair,
                    c->pchunk_beg, c->pchunk_end, c->W, c->V, c->Vinv, c->U, c->g, c->UVdiag_scr, c->cam_part, c->pair_part,
                    c->tile_index, c->Stiles, c->Linv, c->eab, c->dp, c->d_status, c->Ldiag, c->cam2pos, c->pos2cam, c->d_crit_I, c->d_crit_K,
                    c->d_psrc_ptr, c->d_psrc, c->d_b_J, c->d_b_sptr, c->d_b_slot, c->d_def_I, c->d_def_J, c->d_def_sptr, c->d_def_src,
                    c->d_step_panels, c->d_crit_desc, c->d_def_desc, c->d_crit_src, c->d_def_srcs, c->d_bw_order, c->bw_xbuf, c->contrib, c->d_coltile_ptr, c->d_coltile_row, c->d_coltile_slot,
                    c->Sdense, c->Sdense_aux, c->chol_aux, c->chol_diag, c->chol_E, c->d_part, c->d_scal, c->P_U, c->P_B, c->P,
                    c->pcg_work, c->d_flow_tasks, c->d_flow_final, c->d_flow_defseq, c->d_flow_bseq, c->d_flow_critneed, c->d_flow_ver,
                    c->tmpA, c->tmpB, (void *)c->ext.kc, (void *)c->ext.wgt, c->seg_desc, c->sched_chunk, c->sch_beg, c->sch_end, c->tri_vr,
                    c->ring_wrow_ptr, c->ring_rows, c->ring_info};
    for (void *p : ptrs) {
        const char *q = (const char *)p, *b = (const char *)c->tile_block;   // tables inside the schedule block are not allocations
        if (b && q >= b && q < b + c->tile_block_bytes) continue;
        psba_dev_free(c, p);
    }
    psba_dev_free(c, c->tile_block);
    psba_dev_free(c, c->stage_impts); psba_dev_free(c, c->stage_pts);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (c->chol_graph_ok) cudaGraphExecDestroy(c->chol_graph);
    if (c->bw_graph_ok) cudaGraphExecDestroy(c->bw_graph);
    for (auto &kv : *c->seqs) if (kv.second.ok) cudaGraphExecDestroy(kv.second.exec);
    delete c->seqs;
    psba_dev_free(c, c->d_mu);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    cudaFreeHost(c->h_mu_ring);
    cudaFreeHost(c->h_scal); cudaFreeHost(c->h_status);
    cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join);
    cudaStreamDestroy(c->stream2);
    cudaStreamDestroy(c->stream);
    delete c;
}

// ------------------------------------------------------------------------------------------------
static void d2h(psba_ctx *c, double *host, const double *dev, size_t n)
{
    CUDA_CHECK(cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

static void need_lin(psba_ctx *c, double cu, double cg)
{
    if (!c->lin_valid || c->coeff_uvw != cu || c->coeff_g != cg) psba_launch_linearize(c, cu, cg);
}

extern "C" double psba_compute_exQT(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, int params, double *ex)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    const int set = params == PSBA_PARAMS_CUR ? c->cur : 1 - c->cur;
    double *exd = nullptr;
    if (ex) { if (!c->tmpA) c->tmpA = dalloc<double>(c, (size_t)c->o * 18); exd = c->tmpA; }
    const double cost = psba_launch_cost(c, set, exd);
    if (ex) d2h(c, ex, exd, (size_t)c->o * 2);
    return cost;
}

extern "C" void psba_compute_jacobiQT(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *jac_A, double *jac_B)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    c->lin_valid = false;            // a new linearisation point; products are rebuilt on demand
    if (jac_A || jac_B) {
        if (!c->tmpA) c->tmpA = dalloc<double>(c, (size_t)c->o * 18);
        if (!c->tmpB) c->tmpB = dalloc<double>(c, (size_t)c->o * 18);
        psba_launch_jac_materialize(c, c->tmpA, c->tmpB);
        if (jac_A) d2h(c, jac_A, c->tmpA, (size_t)c->o * 12);
        if (jac_B) d2h(c, jac_B, c->tmpB, (size_t)c->o * 6);
    }
}

extern "C" void psba_compute_U(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double coeff, double *out)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    need_lin(c, coeff, c->coeff_g);
    c->mu_pending = 0.0;
    if (out) d2h(c, out, c->U, (size_t)c->m * 36);
}

extern "C" void psba_compute_V(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double coeff, double *out)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    need_lin(c, coeff, c->coeff_g);
    if (out) {
        std::vector<double> v((size_t)c->n * 6);
        d2h(c, v.data(), c->V, v.size());
        for (int i = 0; i < c->n; ++i) {
            const double *p = &v[(size_t)i * 6]; double *q = out + (size_t)i * 9;
            q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; q[3] = p[1]; q[4] = p[3]; q[5] = p[4]; q[6] = p[2]; q[7] = p[4]; q[8] = p[5];
        }
    }
}

extern "C" void psba_compute_Wblks(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs,
                                   const int *iidx, const int *jidx, double coeff, double *Wblks)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs; (void)iidx; (void)jidx;
    need_lin(c, coeff, c->coeff_g);
    if (Wblks) d2h(c, Wblks, c->W, (size_t)c->o * 18);
}

extern "C" void psba_compute_g(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double coeff, double *g)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    need_lin(c, c->coeff_uvw, coeff);
    if (g) d2h(c, g, c->g, (size_t)c->N + 3 * (size_t)c->n);
}

extern "C" double psba_maxElmOfUV(psba_ctx *c, int totalParas, double *UVdiag)
{
    (void)totalParas;
    if (!c->lin_valid) die("maxElmOfUV before compute_U/compute_V");
    if (UVdiag) {
        std::vector<double> u((size_t)c->m * 36), v((size_t)c->n * 6);
        d2h(c, u.data(), c->U, u.size()); d2h(c, v.data(), c->V, v.size());
        for (int j = 0; j < c->m; ++j) for (int r = 0; r < 6; ++r) UVdiag[j * 6 + r] = u[(size_t)j * 36 + r * 7];
        for (int i = 0; i < c->n; ++i) { UVdiag[c->N + i * 3] = v[(size_t)i * 6]; UVdiag[c->N + i * 3 + 1] = v[(size_t)i * 6 + 3]; UVdiag[c->N + i * 3 + 2] = v[(size_t)i * 6 + 5]; }
    }
    return psba_launch_maxdiag(c);
}

extern "C" void psba_update_UV(psba_ctx *c, int cnp, int pnp, int n3Dpts, int nCams, double mu, double *U, double *V)
{
    (void)cnp; (void)pnp; (void)n3Dpts; (void)nCams; (void)U; (void)V;
    c->mu_pending += mu;             // U[rc] = U[rc] + mu (update_UV.cl:21,29)
    c->S_valid = false;
}

extern "C" void psba_restore_UVdiag(psba_ctx *c, int cnp, int pnp, int n3Dpts, int nCams)
{
    (void)cnp; (void)pnp; (void)n3Dpts; (void)nCams;
    c->mu_pending = 0.0;
    c->S_valid = false;
}

extern "C" double psba_compute_Vinv(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *V)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    if (!c->lin_valid) die("compute_Vinv before compute_V");
    psba_launch_vinv(c, c->mu_pending);
    int flag = 0;
    CUDA_CHECK(cudaMemcpyAsync(&flag, c->d_status + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (V) {   // mixed-triangle layout of compute_Vinv.cl:76-86
        std::vector<double> v((size_t)c->n * 6), vi((size_t)c->n * 6);
        d2h(c, v.data(), c->V, v.size()); d2h(c, vi.data(), c->Vinv, vi.size());
        for (int i = 0; i < c->n; ++i) {
            const double *p = &v[(size_t)i * 6], *q = &vi[(size_t)i * 6]; double *o9 = V + (size_t)i * 9;
            o9[0] = q[0]; o9[1] = p[1]; o9[2] = p[2];
            o9[3] = q[1]; o9[4] = q[3]; o9[5] = p[4];
            o9[6] = q[2]; o9[7] = q[4]; o9[8] = q[5];
        }
    }
    return flag ? 1.0 : 0.0;
}

extern "C" void psba_compute_Yblks(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs,
                                   const int *iidx, const int *jidx, double *Yblks)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs; (void)iidx; (void)jidx;
    if (Yblks) {
        if (!c->tmpB) c->tmpB = dalloc<double>(c, (size_t)c->o * 18);
        psba_launch_Y_materialize(c, c->tmpB);
        d2h(c, Yblks, c->tmpB, (size_t)c->o * 18);
    }
}

static void ensure_dense(psba_ctx *c)
{
    const size_t nn = (size_t)c->N * c->N;
    if (!c->Sdense) c->Sdense = dalloc<double>(c, nn);
    if (!c->Sdense_aux) c->Sdense_aux = dalloc<double>(c, nn);
}

extern "C" void psba_compute_S(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *S)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    if (!c->lin_valid) die("compute_S before compute_U/V/Wblks");
    psba_launch_schur(c, c->mu_pending);
    if (S) {
        ensure_dense(c);
        psba_tiles_to_dense(c, c->Sdense, true);
        d2h(c, S, c->Sdense, (size_t)c->N * c->N);
    }
}

extern "C" void psba_compute_ea(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *ea)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    // ea is produced together with S (same pass over the camera pairs)
    if (ea) d2h(c, ea, c->eab, (size_t)c->N);
}

extern "C" double psba_SPDinv(psba_ctx *c, int matSize, double *outMat)
{
    if (matSize != c->N) die("SPDinv: matSize != 6*nCams");
    if (!c->S_valid) die("SPDinv before compute_S");
    const double ret = psba_launch_factor(c);
    if (ret == 0.0 && outMat) {
        ensure_dense(c);
        double *tmp = dalloc<double>(c, (size_t)c->N * c->N);
        psba_launch_explicit_inverse(c, tmp);
        d2h(c, outMat, tmp, (size_t)c->N * c->N);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        psba_dev_free(c, tmp);
    }
    return ret;
}

// the three stages of SPDinv as the reference exports them (PSBA/cl_spdinv.h:10-18).  The engine factorises once; each entry
// hands out its product in the CALLERS' camera order: cholesky -> M with M M^T = S (M = P^T L P: the lower-triangular factor
// itself whenever the solver kept the natural camera order, i.e. for every dense system), trigMat_inv -> M^-1, trigMat_mul ->
// S^-1.  Dense N x N outputs: small systems only.
extern "C" double psba_cholesky(psba_ctx *c, int matSize, double *outMat)
{
    if (matSize != c->N) die("cholesky: matSize != 6*nCams");
    if (!c->S_valid) die("cholesky before compute_S");
    const double ret = psba_launch_factor(c);
    if (ret == 0.0 && outMat) {
        double *tmp = dalloc<double>(c, (size_t)c->N * c->N);
        psba_launch_factor_products(c, tmp, 2);
        d2h(c, outMat, tmp, (size_t)c->N * c->N);
        psba_dev_free(c, tmp);
    }
    return ret;
}
static double factor_product(psba_ctx *c, int matSize, double *outMat, int what, const char *who)
{
    if (matSize != c->N) die("trigMat_*: matSize != 6*nCams");
    if (!c->factor_valid) { fprintf(stderr, "psba_b200: %s before a successful cholesky / SPDinv\n", who); exit(EXIT_FAILURE); }
    if (outMat) {
        double *tmp = dalloc<double>(c, (size_t)c->N * c->N);
        psba_launch_factor_products(c, tmp, what);
        d2h(c, outMat, tmp, (size_t)c->N * c->N);
        psba_dev_free(c, tmp);
    }
    return 0.0;
}
extern "C" double psba_trigMat_inv(psba_ctx *c, int matSize, double *outMat) { return factor_product(c, matSize, outMat, 1, "trigMat_inv"); }
extern "C" void psba_trigMat_mul(psba_ctx *c, int matSize, double *outMat) { factor_product(c, matSize, outMat, 0, "trigMat_mul"); }
// get_delta_beta / compute_cholmod_E (PSBA/cl_cholmod.h:14-19) on the S of the last compute_S / the E of the last cholmod_blk
extern "C" void psba_get_delta_beta(psba_ctx *c, int matSize, double *delta, double *beta)
{
    if (matSize != c->N) die("get_delta_beta: matSize != 6*nCams");
    if (!c->S_valid) die("get_delta_beta before compute_S");
    psba_tile_delta_beta(c, delta, beta);
}
extern "C" void psba_compute_cholmod_E(psba_ctx *c, int matSize, double *Eout)
{
    if (matSize != c->N) die("compute_cholmod_E: matSize != 6*nCams");
    if (Eout) d2h(c, Eout, c->chol_E, (size_t)c->N);
}

extern "C" void psba_matVec_mul(psba_ctx *c, int mat_rsize, int mat_csize, double *out)
{
    (void)mat_rsize; (void)mat_csize;
    if (!c->factor_valid) die("matVec_mul before a successful SPDinv");
    psba_launch_solve(c);
    if (out) d2h(c, out, c->dp, (size_t)c->N);
}

extern "C" void psba_compute_eb(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *eab)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    psba_launch_backsub(c, c->mu_pending, false, nullptr);
    if (eab) d2h(c, eab, c->eab, (size_t)c->N + 3 * (size_t)c->n);
}

extern "C" void psba_compute_dpb(psba_ctx *c, int cnp, int pnp, int nCams, int n3Dpts, double *dp)
{
    (void)cnp; (void)pnp; (void)nCams; (void)n3Dpts;
    // dpb was produced by the fused back-substitution of compute_eb
    if (dp) d2h(c, dp, c->dp, (size_t)c->N + 3 * (size_t)c->n);
}

extern "C" void psba_compute_newp(psba_ctx *c, int nCamParas, int n3DptsParas, double *new_p)
{
    (void)nCamParas; (void)n3DptsParas;
    psba_launch_newp(c);
    if (new_p) { d2h(c, new_p, c->cams[1 - c->cur], (size_t)c->N); d2h(c, new_p + c->N, c->pts[1 - c->cur], 3 * (size_t)c->n); }
}

extern "C" void psba_update_p(psba_ctx *c, int nCamParas, int n3DptsParas, double *p)
{
    (void)nCamParas; (void)n3DptsParas;
    c->cur = 1 - c->cur;             // p <- new_p is a pointer swap
    c->lin_valid = false;
    if (p) { d2h(c, p, c->cams[c->cur], (size_t)c->N); d2h(c, p + c->N, c->pts[c->cur], 3 * (size_t)c->n); }
}

extern "C" double psba_compute_Jmultiply(psba_ctx *c, int mnp, int n3Dpts, int nCams, int n2Dprojs, int x, double *out)
{
    (void)mnp; (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    const double *xv = x == PSBA_VEC_G ? c->g : c->dp;
    double res[3];
    double *jx = nullptr;
    if (out) { if (!c->tmpA) c->tmpA = dalloc<double>(c, (size_t)c->o * 18); jx = c->tmpA; }
    psba_launch_Jdot(c, xv, xv, jx, res);
    if (out) d2h(c, out, jx, (size_t)c->o * 2);
    return res[0];
}

extern "C" void psba_upload_vec(psba_ctx *c, int vec, const double *host, int n)
{
    double *dst = vec == PSBA_VEC_G ? c->g : c->dp;
    CUDA_CHECK(cudaMemcpyAsync(dst, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

extern "C" double psba_cholmod_blk(psba_ctx *c, int matSize, double *E, double *delta, double *beta, int *n_scalar_blocks)
{
    if (matSize != c->N) die("cholmod_blk: matSize != 6*nCams");
    if (!c->S_valid) die("cholmod_blk needs the S of the last compute_S (call it again after a failed SPDinv)");
    if (psba_cholmod_use_tiles(c)) {
        double ratio = 0.0;
        const double sum = psba_launch_cholmod_tiles(c, delta, beta, n_scalar_blocks, E, &ratio);
        const char *e = getenv("PSBA_CHOLMOD_TILES");
        if (!(ratio > 1.0 && psba_cholmod_dense_possible(c) && !(e && atoi(e)))) return sum;
        psba_launch_schur(c, c->mu_pending);          // the `> beta` rescue lives in the dense kernel: rebuild S for it
    }
    ensure_dense(c);
    psba_tiles_to_dense(c, c->Sdense, true);
    const double sum = psba_launch_cholmod(c, delta, beta, n_scalar_blocks);
    if (E) d2h(c, E, c->chol_E, (size_t)c->N);
    return sum;
}

// the reference's cholmod_blk takes the matrix buffer itself (cl_cholmod.h:10-12): same here for a HOST matrix
// (symmetric, row-major, matSize a multiple of 3); mat receives the factor (lower triangle, zero above), E the E_i.
extern "C" double psba_cholmod_blk_mat(psba_ctx *c, int matSize, double *mat, double *E, double *delta, double *beta, int *n_scalar_blocks)
{
    if (matSize <= 0 || matSize % 3) die("cholmod_blk_mat: matSize must be a positive multiple of 3");
    const size_t nn = (size_t)matSize * matSize;
    double *d = (double *)psba_dev_alloc(c, nn * 8, false), *aux = (double *)psba_dev_alloc(c, ((size_t)3 * matSize + 2 * TS) * 8, true);
    double *dinv = (double *)psba_dev_alloc(c, ((size_t)3 * matSize + 2 * TS) * 8, true), *Ed = (double *)psba_dev_alloc(c, ((size_t)matSize + TS) * 8, true);
    CUDA_CHECK(cudaMemcpyAsync(d, mat, nn * 8, cudaMemcpyHostToDevice, c->stream));
    const double sum = psba_launch_cholmod_dense(c, matSize, d, aux, dinv, Ed, delta, beta, n_scalar_blocks);
    // the kernel addresses its matrix transposed (element (r, c) at [c * N + r]): hand the factor back row-major
    std::vector<double> t(nn);
    d2h(c, t.data(), d, nn);
    for (int r = 0; r < matSize; ++r) for (int q = 0; q < matSize; ++q) mat[(size_t)r * matSize + q] = q <= r ? t[(size_t)q * matSize + r] : 0.0;
    if (E) d2h(c, E, Ed, (size_t)matSize);
    psba_dev_free(c, d); psba_dev_free(c, aux); psba_dev_free(c, dinv); psba_dev_free(c, Ed);
    return sum;
}

extern "C" void psba_get_params(psba_ctx *c, int params, double *cams, double *pts)
{
    const int set = params == PSBA_PARAMS_CUR ? c->cur : 1 - c->cur;
    if (cams) d2h(c, cams, c->cams[set], (size_t)c->N);
    if (pts) d2h(c, pts, c->pts[set], 3 * (size_t)c->n);
}

extern "C" void psba_set_params(psba_ctx *c, const double *cams, const double *pts_global)
{
    // restart from given parameters: cams[m*6] (all cameras), pts_global[n3Dpts*3] (this rank takes its slice)
    if (cams) CUDA_CHECK(cudaMemcpyAsync(c->cams[c->cur], cams, (size_t)c->N * 8, cudaMemcpyHostToDevice, c->stream));
    if (pts_global && c->n) CUDA_CHECK(cudaMemcpyAsync(c->pts[c->cur], pts_global + (size_t)c->p_off * 3, (size_t)c->n * 24, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->cache_valid[0] = c->cache_valid[1] = false;
    c->lin_valid = false; c->S_valid = false; c->factor_valid = false; c->mu_pending = 0.0;
}

static void ext_changed(psba_ctx *c)
{
    const bool force = getenv("PSBA_FORCE_EXT") && atoi(getenv("PSBA_FORCE_EXT"));   // tests: the extended kernels with kc = 0
    c->ext_on = c->ext.kc || c->ext.wgt || force;
    c->lin_valid = false; c->S_valid = false; c->factor_valid = false;
}

// Fixed lens distortion per camera: kc[m*5] = (k1, k2, p1, p2, k3) of the sba "varKD" camera (data/54camsvarKD.txt
// columns 6-10; PSBA/misc.cpp:27-29 copies them through quat2vec, the reference's kernels then ignore them).
// NULL or all-zero coefficients switch the model off (the reference's undistorted projection, bit for bit).
// captured chains hold kernel arguments by value: anything that replaces a device array they name drops them
static void seq_invalidate(psba_ctx *c)
{
    for (auto &kv : *c->seqs) if (kv.second.ok) cudaGraphExecDestroy(kv.second.exec);
    c->seqs->clear();
    c->capturing = nullptr;
}

extern "C" void psba_set_distortion(psba_ctx *c, const double *kc)
{
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    seq_invalidate(c);
    psba_dev_free(c, (void *)c->ext.kc); c->ext.kc = nullptr;
    bool any = false;
    if (kc) for (int q = 0; q < c->m * 5; ++q) any |= kc[q] != 0.0;
    if (any) {
        double *d = (double *)psba_dev_alloc(c, (size_t)c->m * 5 * 8, false);
        CUDA_CHECK(cudaMemcpyAsync(d, kc, (size_t)c->m * 5 * 8, cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->ext.kc = d;
    }
    ext_changed(c);
}

// Image-point covariances (covimgpts of readInitialSBAEstimate, PSBA/readparams.cpp:272-283, 380-413: parsed, never used
// by any reference kernel): cov[n2Dprojs * covsz], covsz = 4 (full 2x2, row-major) or 3 (upper triangle s00 s01 s11), for
// ALL observations (a rank takes its slice).  The residual of observation k becomes W_k e_k with W_k^T W_k = Sigma_k^-1
// (W = L^-1, Sigma = L L^T), as Lourakis' sba weights its residuals.  NULL switches the weights off.  Call after
// fill_idxBuffer.  Returns 0, or 1 if a covariance is not positive definite.
extern "C" int psba_set_covariances(psba_ctx *c, const double *cov, int covsz)
{
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    seq_invalidate(c);
    psba_dev_free(c, (void *)c->ext.wgt); c->ext.wgt = nullptr;
    if (cov) {
        if (covsz != 3 && covsz != 4) die("set_covariances: covsz must be 3 or 4");
        if (!c->iidx) die("set_covariances before fill_idxBuffer");
        std::vector<double> w((size_t)c->o * 3);
        for (int k = 0; k < c->o; ++k) {
            const double *s = cov + (size_t)(c->o_off + k) * covsz;
            const double s00 = s[0], s01 = s[1], s11 = covsz == 4 ? s[3] : s[2];
            if (!(s00 > 0.0) || !(s00 * s11 - s01 * s01 > 0.0)) { fprintf(stderr, "psba_b200: covariance of observation %d is not positive definite\n", c->o_off + k); ext_changed(c); return 1; }
            const double l00 = std::sqrt(s00), l10 = s01 / l00, l11 = std::sqrt(s11 - l10 * l10);
            w[(size_t)k * 3] = 1.0 / l00; w[(size_t)k * 3 + 1] = -l10 / (l00 * l11); w[(size_t)k * 3 + 2] = 1.0 / l11;
        }
        double *d = (double *)psba_dev_alloc(c, std::max<size_t>(w.size(), 1) * 8, false);
        if (!w.empty()) CUDA_CHECK(cudaMemcpyAsync(d, w.data(), w.size() * 8, cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->ext.wgt = d;
    }
    ext_changed(c);
    return 0;
}

extern "C" void psba_set_option(psba_ctx *c, const char *name, double v)
{
    std::string s(name);
    if (s == "verbose") c->verbose = (int)v;
    else if (s == "max_iter") c->max_iter = (int)v;
    else if (s == "itno") c->itno = (int)v;
    else if (s == "lm_only") c->lm_only = (int)v;
    else if (s == "tr_fused") c->tr_fused = (int)v;
    else if (s == "seq_graphs") c->use_graphs = (int)v;
    else if (s == "camera_solver") { c->camera_solver = (int)v; c->factor_valid = false; }      // 0 tiled Cholesky (default), 1 block-Jacobi PCG
    else if (s == "pcg_tol") c->pcg_tol = v;
    else if (s == "pcg_max_iter") c->pcg_max_iter = (int)v;
    else if (s == "profile") { psba_prof_collect(c); c->profile = v != 0; }
    else if (s == "profile_reset") { psba_prof_collect(c); for (int k = 0; k < KID_COUNT; ++k) { c->prof_ms[k] = 0; c->prof_n[k] = 0; } }
    else if (s == "trace_reset") { c->trace.clear(); c->n_cholmod_events = 0; }
    else if (s == "stats_reset") { c->st_tries = c->st_exqt = c->st_lin = c->st_launches = c->st_seq_replays = 0; }
    else if (s == "timer_start") {
        if (!c->timer_init) { CUDA_CHECK(cudaEventCreate(&c->timer_e0)); CUDA_CHECK(cudaEventCreate(&c->timer_e1)); c->timer_init = true; }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        CUDA_CHECK(cudaEventRecord(c->timer_e0, c->stream));
    }
    else die("set_option: unknown option");
}

extern "C" double psba_get_stat(psba_ctx *c, const char *name)
{
    std::string s(name);
    if (s == "tries") return c->st_tries;
    if (s == "exqt") return c->st_exqt;
    if (s == "linearizations") return c->st_lin;
    if (s == "launches") return c->st_launches;
    if (s == "itno") return c->itno;
    if (s == "initErr") return c->initErr;
    if (s == "n_local") return c->n;
    if (s == "o_local") return c->o;
    if (s == "p_off") return c->p_off;
    if (s == "o_off") return c->o_off;
    if (s == "ntriples") return (double)c->ntri;
    if (s == "n_pairs") return c->n_pair;
    if (s == "n_tiles") return c->n_tiles;
    if (s == "n_tiles_S") return c->n_tiles_S;
    if (s == "nt") return c->nt;
    if (s == "n_steps") return c->n_steps;
    if (s == "chol_flops") return c->chol_flops;
    if (s == "n_ptchunk") return c->n_ptchunk;
    if (s == "n_cchunk") return c->n_cchunk;
    if (s == "n_pchunk") return c->n_pchunk;
    if (s == "pair_G") return c->pair_G;
    if (s == "pair_mode") return c->pair_mode;
    if (s == "n_seg") return c->n_seg;
    if (s == "seg_v") return c->seg_v;
    if (s == "seq_replays") return c->st_seq_replays;
    if (s == "seq_graphs") { int k = 0; for (auto &kv : *c->seqs) k += kv.second.ok ? 1 : 0; return k; }
    if (s == "seq_keys") return (double)c->seqs->size();
    if (s == "ring_rows") return (double)c->ring_n_rows;
    if (s == "ring_cfg") return c->pair_mode == 6 ? c->ring_cfg : -1;
    if (s == "ring_rt") return c->pair_mode == 6 ? c->ring_rt : -1;
    if (s == "cholmod_events") return c->n_cholmod_events;
    if (s == "pcg_iterations") return c->pcg_last_iters;
    if (s == "cholmod_max_l_over_beta") return c->cholmod_max_l_over_beta;
    if (s == "timer_ms") {   // device time since "timer_start" on the engine's stream
        if (!c->timer_init) die("timer_ms before timer_start");
        float ms = 0;
        CUDA_CHECK(cudaEventRecord(c->timer_e1, c->stream));
        CUDA_CHECK(cudaEventSynchronize(c->timer_e1));
        CUDA_CHECK(cudaEventElapsedTime(&ms, c->timer_e0, c->timer_e1));
        return ms;
    }
    if (s.rfind("ms.", 0) == 0 || s.rfind("n.", 0) == 0) {
        psba_prof_collect(c);
        const bool want_ms = s[0] == 'm';
        const std::string kn = s.substr(want_ms ? 3 : 2);
        for (int k = 0; k < KID_COUNT; ++k) if (kn == psba_kid_name[k]) return want_ms ? c->prof_ms[k] : c->prof_n[k];
        die("get_stat: unknown kernel name");
    }
    die("get_stat: unknown name");
    return 0;
}

extern "C" long long psba_get_index(psba_ctx *c, const char *name, void *out, long long max_count)
{
    const std::string s(name);
    const void *src = nullptr; long long cnt = 0; size_t esz = 4;
    if (s == "pt_ptr") { src = c->pt_ptr; cnt = c->n + 1; }
    else if (s == "cam_obs") { src = c->cam_obs; cnt = c->o; }
    else if (s == "pair_k") { src = c->pair_k; cnt = c->n_pair; }
    else if (s == "pair_l") { src = c->pair_l; cnt = c->n_pair; }
    else if (s == "tri_oa") { src = c->tri_oa; cnt = c->ntri; }
    else if (s == "tri_ob") { src = c->tri_ob; cnt = c->ntri; }
    else if (s == "tri_pt") { src = c->tri_pt; cnt = c->ntri; }
    else if (s == "pchunk_pair") { src = c->pchunk_pair; cnt = c->pchunk_pair ? c->n_pchunk : 0; }
    else if (s == "pchunk_beg") { src = c->pchunk_beg; cnt = c->pchunk_beg ? c->n_pchunk : 0; esz = 8; }
    else if (s == "pchunk_end") { src = c->pchunk_end; cnt = c->pchunk_end ? c->n_pchunk : 0; esz = 8; }
    else if (s == "pair_chunk_ptr") { src = c->pair_chunk_ptr; cnt = c->n_pair + 1; }
    else if (s == "chunk_beg") { src = c->sch_beg; cnt = c->sch_beg ? c->n_pchunk : 0; }
    else if (s == "chunk_end") { src = c->sch_end; cnt = c->sch_end ? c->n_pchunk : 0; }
    else if (s == "sched_chunk") { src = c->sched_chunk; cnt = c->sched_chunk ? c->n_pchunk : 0; }
    else if (s == "seg_desc") { src = c->seg_desc; cnt = (long long)c->n_seg * 6; }
    else if (s == "cam2pos") { src = c->cam2pos; cnt = c->m; }
    else die("get_index: unknown table");
    if (out && cnt > 0) {
        CUDA_CHECK(cudaMemcpyAsync(out, src, (size_t)std::min(cnt, max_count) * esz, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    return cnt;
}

void psba_prof_collect(psba_ctx *c)
{
    if (c->prof_pending.empty()) return;
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (auto &r : c->prof_pending) {
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, r.e0, r.e1));
        c->prof_ms[r.id] += ms; c->prof_n[r.id] += 1;
        c->prof_pool.push_back(r.e0); c->prof_pool.push_back(r.e1);
    }
    c->prof_pending.clear();
}

extern "C" void psba_force_lambda(psba_ctx *c, const double *lam, int n)
{
    c->force_lambda.assign(lam, lam + n);
}

extern "C" int psba_trace_count(psba_ctx *c) { return (int)c->trace.size(); }
extern "C" void psba_trace_get(psba_ctx *c, int k, psba_trace_rec *rec) { *rec = c->trace[k]; }

// fused hot-path steps -------------------------------------------------------------------------
extern "C" void psba_linearize(psba_ctx *c, double coeff_uvw, double coeff_g)
{
    psba_launch_linearize(c, coeff_uvw, coeff_g);
}


// ---- device-resident scalars and sequence graphs ------------------------------------------------------------------
void psba_set_scalars(psba_ctx *c, double mu, double a, double b)
{
    // a captured chain copies from ITS slot (the replay writes the slot before the launch); plain launches rotate through
    // the other slots: a slot is rewritten 32 calls later at the earliest, and every try ends with a host synchronisation
    const int slot = c->capturing ? c->capturing->slot : 32 + c->h_mu_next;
    if (!c->capturing) c->h_mu_next = (c->h_mu_next + 1) & 31;
    double *h = c->h_mu_ring + (size_t)slot * 4;
    h[0] = mu; h[1] = a; h[2] = b; h[3] = 0.0;
    CUDA_CHECK(cudaMemcpyAsync(c->d_mu, h, 4 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
}

static bool seq_graphs_on(const psba_ctx *c)
{
    if (c->profile || c->nranks > 1 || c->camera_solver == 1) return false;      // event scopes / NCCL / host-driven PCG stay out of graphs
    if (c->use_graphs >= 0) return c->use_graphs != 0;
    return c->o_glob <= 2000000;           // the chains of a large problem are not launch-bound (2.84 ms of kernels in 2.84 ms)
}

bool psba_seq_begin(psba_ctx *c, int kind, double mu, double a, double b)
{
    if (!seq_graphs_on(c)) return true;
    const unsigned long long key = (unsigned long long)kind | ((unsigned long long)c->cur << 8) | ((unsigned long long)c->cache_valid[0] << 9) |
                                   ((unsigned long long)c->cache_valid[1] << 10) | ((unsigned long long)(c->ext_on ? 1 : 0) << 11) |
                                   ((unsigned long long)c->pair_mode << 12) | ((unsigned long long)(c->S_valid ? 1 : 0) << 20) |
                                   ((unsigned long long)(c->lin_valid ? 1 : 0) << 21);
    psba_ctx::seq_graph &g = (*c->seqs)[key];
    if (g.seen == 0 && !g.ok) { g.exec = nullptr; g.slot = (int)((c->seqs->size() - 1) & 31); }
    if (g.ok) {
        double *h = c->h_mu_ring + (size_t)g.slot * 4;
        h[0] = mu; h[1] = a; h[2] = b; h[3] = 0.0;
        CUDA_CHECK(cudaGraphLaunch(g.exec, c->stream));
        c->st_seq_replays += 1;
        c->st_launches += g.d_launches; c->st_tries += g.d_tries; c->st_exqt += g.d_exqt; c->st_lin += g.d_lin;
        c->cache_valid[0] = g.cv0; c->cache_valid[1] = g.cv1; c->S_valid = g.S_valid; c->factor_valid = g.factor_valid; c->lin_valid = g.lin_valid;
        if (kind == SEQ_LIN_LM || kind == SEQ_LIN_TR) { c->coeff_uvw = g.cu; c->coeff_g = g.cg; }
        return false;
    }
    g.seen += 1;
    if (g.seen < 2 || c->seqs->size() > 32) { c->capturing = nullptr; return true; }   // first sighting: plain launches (lazy set-up of the kernels runs here)
    c->capturing = &g;
    c->snap_launches = c->st_launches; c->snap_tries = c->st_tries; c->snap_exqt = c->st_exqt; c->snap_lin = c->st_lin;
    CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    return true;
}

void psba_seq_end(psba_ctx *c)
{
    psba_ctx::seq_graph *g = c->capturing;
    if (!g) return;
    c->capturing = nullptr;
    cudaGraph_t graph;
    CUDA_CHECK(cudaStreamEndCapture(c->stream, &graph));
    CUDA_CHECK(cudaGraphInstantiate(&g->exec, graph, 0));
    CUDA_CHECK(cudaGraphDestroy(graph));
    g->ok = true;
    g->d_launches = c->st_launches - c->snap_launches; g->d_tries = c->st_tries - c->snap_tries;
    g->d_exqt = c->st_exqt - c->snap_exqt; g->d_lin = c->st_lin - c->snap_lin;
    g->cv0 = c->cache_valid[0]; g->cv1 = c->cache_valid[1]; g->S_valid = c->S_valid; g->factor_valid = c->factor_valid; g->lin_valid = c->lin_valid;
    g->cu = c->coeff_uvw; g->cg = c->coeff_g;
    CUDA_CHECK(cudaGraphLaunch(g->exec, c->stream));            // the captured chain has not run yet
}

extern "C" void psba_try_step(psba_ctx *c, double mu, psba_try_result *res)
{
    if (!c->lin_valid) die("try_step before linearize");
    res->cost_new = res->dp_L2 = res->dp_dot = res->p_new_L2 = NAN;
    if (psba_seq_begin(c, SEQ_TRY, mu, 0.0, 0.0)) {  // the whole try is one chain: a CUDA graph from its third use on (small problems)
        c->st_tries += 1;
        psba_launch_schur(c, mu);
        if (c->camera_solver == 1) psba_launch_pcg(c);   // optional iterative camera solve (kernels_pcg.cu)
        else {
            psba_launch_factor(c, true);                 // no host round trip between factorisation and solves
            psba_launch_solve(c);
        }
        psba_enqueue_backsub(c, mu, true);               // step scalars and status word on their way to the host
        psba_seq_end(c);
    }
    psba_finish_try(c, res);
    if (c->h_status[0] > 1) { fprintf(stderr, "psba_b200: camera solve failed with status %d (broken dataflow schedule)\n", c->h_status[0]); exit(EXIT_FAILURE); }
    res->solve_status = c->h_status[0] ? 1.0 : 0.0;
    if (res->solve_status != 0.0) { c->factor_valid = false; res->cost_new = res->dp_L2 = res->dp_dot = res->p_new_L2 = NAN; }
}
