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

template <class T> static T *dalloc(size_t n)
{
    T *p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    CUDA_CHECK(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}
template <class T> static T *dupload(const std::vector<T> &h)
{
    T *p = dalloc<T>(h.size());
    if (!h.empty()) CUDA_CHECK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}

// host staging between fill_initBuffer2 and fill_idxBuffer (the local slice is only known once
// the observation -> point map arrives)
struct psba_stage {
    std::vector<double> K, impts, initcams, cams, pts;
};
static psba_stage *g_stage_of(psba_ctx *c);
#include <map>
static std::map<psba_ctx *, psba_stage> g_stage;
static psba_stage *g_stage_of(psba_ctx *c) { return &g_stage[c]; }

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
    c->cur = 0; c->cache_valid[0] = c->cache_valid[1] = false;
    c->lin_valid = false; c->S_valid = false; c->factor_valid = false; c->chol_graph_ok = false; c->bw_graph_ok = false;
    c->Sdense = c->Sdense_aux = nullptr; c->tmpA = c->tmpB = nullptr;
    c->mu_pending = 0.0; c->coeff_uvw = 1.0; c->coeff_g = 1.0;
    c->itno = 0; c->max_iter = 50; c->verbose = 0; c->lm_only = 0; c->initErr = 0.0;
    c->n_cholmod_events = 0;
    c->st_tries = c->st_exqt = c->st_lin = c->st_launches = 0;
    c->profile = false; c->timer_init = false;
    for (int k = 0; k < KID_COUNT; ++k) { c->prof_ms[k] = 0; c->prof_n[k] = 0; }
    c->comm = nullptr;
    c->K = dalloc<double>((size_t)nCams * 5);
    c->initcams = dalloc<double>((size_t)nCams * 4);
    for (int s = 0; s < 2; ++s) {
        c->cams[s] = dalloc<double>((size_t)nCams * 6);
        c->camcache[s] = dalloc<double>((size_t)nCams * CAMC);
    }
    c->U = dalloc<double>((size_t)nCams * 36);
    c->d_status = dalloc<int>(4);
    c->d_scal = dalloc<double>(NSCAL);
    CUDA_CHECK(cudaMallocHost(&c->h_scal, NSCAL * sizeof(double)));
    return c;
}

extern "C" void psba_fill_initBuffer2(psba_ctx *c, int cnp, int pnp, int mnp, int nCams, int n3Dpts, int n2Dprojs,
                                      const double *Kparas, const double *impts, const double *initcams,
                                      const double *camsEx, const double *pts3Ds)
{
    check_dims(cnp, pnp, mnp);
    if (nCams != c->m || n3Dpts != c->n_glob || n2Dprojs != c->o_glob) die("fill_initBuffer2: sizes differ from setup_cl");
    psba_stage *st = g_stage_of(c);
    st->K.assign(Kparas, Kparas + (size_t)nCams * 5);
    st->impts.assign(impts, impts + (size_t)n2Dprojs * 2);
    st->initcams.assign(initcams, initcams + (size_t)nCams * 4);
    st->cams.assign(camsEx, camsEx + (size_t)nCams * 6);
    st->pts.assign(pts3Ds, pts3Ds + (size_t)n3Dpts * 3);
}

extern "C" void psba_local_range(int n, int o, const int *iidx, int rank, int nranks, int *p0, int *p1, int *o0, int *o1)
{
    // contiguous point ranges balanced by observation count: rank r owns the points whose first
    // observation index lies in [r*o/R, (r+1)*o/R)
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
    psba_stage *st = g_stage_of(c);
    if (st->pts.size() != (size_t)n3Dpts * 3) die("fill_idxBuffer called before fill_initBuffer2");
    const int m = c->m;
    const bool timing = getenv("PSBA_SETUP_TIMING") != nullptr;
    auto tprev = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        CUDA_CHECK(cudaDeviceSynchronize());
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "psba setup: %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tprev).count());
        tprev = now;
    };
    // observations must be point-major, cameras ascending (generate_idxs, misc.cpp:189-217)
    for (int k = 1; k < n2Dprojs; ++k) {
        if (iidx[k] < iidx[k - 1] || (iidx[k] == iidx[k - 1] && jidx[k] <= jidx[k - 1]))
            die("fill_idxBuffer: observations are not point-major with ascending cameras");
    }
    int p0, p1, o0, o1;
    psba_local_range(n3Dpts, n2Dprojs, iidx, c->rank, c->nranks, &p0, &p1, &o0, &o1);
    c->p_off = p0; c->o_off = o0; c->n = p1 - p0; c->o = o1 - o0;
    const int n = c->n, o = c->o;

    lap("validate + partition");
    // ---- parameters
    CUDA_CHECK(cudaMemcpy(c->K, st->K.data(), (size_t)m * 5 * 8, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(c->initcams, st->initcams.data(), (size_t)m * 4 * 8, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(c->cams[0], st->cams.data(), (size_t)m * 6 * 8, cudaMemcpyHostToDevice));
    c->impts = dalloc<double>((size_t)o * 2);
    if (o) CUDA_CHECK(cudaMemcpy(c->impts, st->impts.data() + (size_t)o0 * 2, (size_t)o * 16, cudaMemcpyHostToDevice));
    for (int s = 0; s < 2; ++s) c->pts[s] = dalloc<double>((size_t)n * 3);
    if (n) CUDA_CHECK(cudaMemcpy(c->pts[0], st->pts.data() + (size_t)p0 * 3, (size_t)n * 24, cudaMemcpyHostToDevice));
    g_stage.erase(c);

    lap("parameter upload");
    // ---- local CSR
    std::vector<int> li(o), lj(o), ptr((size_t)n + 1, 0);
    for (int k = 0; k < o; ++k) { li[k] = iidx[o0 + k] - p0; lj[k] = jidx[o0 + k]; ptr[li[k] + 1]++; }
    for (int i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
    c->iidx = dupload(li); c->jidx = dupload(lj); c->pt_ptr = dupload(ptr);
    // point chunks
    std::vector<int> pch(1, 0);
    {
        int cnt_o = 0, cnt_p = 0;
        for (int i = 0; i < n; ++i) {
            const int d = ptr[i + 1] - ptr[i];
            if (cnt_p > 0 && (cnt_o + d > PT_CTA || cnt_p == PT_CTA)) { pch.push_back(i); cnt_o = 0; cnt_p = 0; }
            cnt_o += d; cnt_p++;
        }
        if (n > 0) pch.push_back(n);
    }
    c->n_ptchunk = (int)pch.size() - 1;
    c->ptchunk = dupload(pch);
    // camera-major lists + chunks
    std::vector<int> cptr((size_t)m + 1, 0), cobs(o);
    for (int k = 0; k < o; ++k) cptr[lj[k] + 1]++;
    for (int j = 0; j < m; ++j) cptr[j + 1] += cptr[j];
    {
        std::vector<int> fill(m, 0);
        for (int k = 0; k < o; ++k) cobs[cptr[lj[k]] + fill[lj[k]]++] = k;
    }
    c->cam_obs = dupload(cobs);
    std::vector<int> cc_cam, cc_beg, cc_end, cc_ptr(1, 0);
    const int CCH = CAM_CTA * CAM_OPT;
    for (int j = 0; j < m; ++j) {
        for (int b = cptr[j]; b < cptr[j + 1]; b += CCH) { cc_cam.push_back(j); cc_beg.push_back(b); cc_end.push_back(std::min(b + CCH, cptr[j + 1])); }
        cc_ptr.push_back((int)cc_cam.size());
    }
    c->n_cchunk = (int)cc_cam.size();
    c->cchunk_cam = dupload(cc_cam); c->cchunk_beg = dupload(cc_beg); c->cchunk_end = dupload(cc_end);
    c->cam_cchunk_ptr = dupload(cc_ptr);

    lap("CSR + chunks");
    // ---- camera-pair structure: GLOBAL set of pairs (k >= l) so that every rank builds the same S layout
    std::vector<int> pair_id((size_t)m * m, -1);
    {
        std::vector<int> gptr((size_t)n3Dpts + 1, 0);
        for (int k = 0; k < n2Dprojs; ++k) gptr[iidx[k] + 1]++;
        for (int i = 0; i < n3Dpts; ++i) gptr[i + 1] += gptr[i];
        for (int i = 0; i < n3Dpts; ++i)
            for (int a = gptr[i]; a < gptr[i + 1]; ++a)
                for (int b = gptr[i]; b <= a; ++b) pair_id[(size_t)jidx[a] * m + jidx[b]] = 0;
    }
    for (int j = 0; j < m; ++j) pair_id[(size_t)j * m + j] = 0;   // every diagonal block exists (U_k)
    std::vector<int> pk, pl;
    std::vector<std::pair<int, int>> pairs;
    for (int k = 0; k < m; ++k)
        for (int l = 0; l <= k; ++l)
            if (pair_id[(size_t)k * m + l] == 0) { pair_id[(size_t)k * m + l] = (int)pk.size(); pk.push_back(k); pl.push_back(l); pairs.push_back({k, l}); }
    c->n_pair = (int)pk.size();
    c->pair_k = dupload(pk); c->pair_l = dupload(pl);
    lap("global pair set");
    // local triples, counting sort by pair (points ascending inside a pair)
    std::vector<long long> tptr((size_t)c->n_pair + 1, 0);
    for (int i = 0; i < n; ++i)
        for (int a = ptr[i]; a < ptr[i + 1]; ++a)
            for (int b = ptr[i]; b <= a; ++b) tptr[pair_id[(size_t)lj[a] * m + lj[b]] + 1]++;
    for (int p = 0; p < c->n_pair; ++p) tptr[p + 1] += tptr[p];
    c->ntri = tptr[c->n_pair];
    std::vector<int> toa((size_t)c->ntri), tob((size_t)c->ntri);
    {
        std::vector<long long> fill(tptr.begin(), tptr.end() - 1);
        for (int i = 0; i < n; ++i)
            for (int a = ptr[i]; a < ptr[i + 1]; ++a)
                for (int b = ptr[i]; b <= a; ++b) {
                    long long at = fill[pair_id[(size_t)lj[a] * m + lj[b]]]++;
                    toa[at] = a; tob[at] = b;
                }
    }
    c->tri_oa = dupload(toa); c->tri_ob = dupload(tob);
    std::vector<int> pc_pair, pc_ptr(1, 0);
    std::vector<long long> pc_beg, pc_end;
    // lane-group size of the pair pass: every group of G lanes owns one chunk of <= G*PAIR_TPL triples of
    // one camera pair; G follows the mean run length so that a lane streams ~PAIR_TPL triples before the
    // (shuffle) reduction -- 4 for the synthetic ring (117 triples / pair), 32 for BAL (~10^3 / pair)
    {
        long long nonempty = 0;
        for (int p = 0; p < c->n_pair; ++p) if (tptr[p + 1] > tptr[p]) ++nonempty;
        const double avg = nonempty ? (double)c->ntri / (double)nonempty : 1.0;
        int G = 1;
        while (G < 32 && avg > (double)G * PAIR_TPL) G *= 2;
        c->pair_G = G;
    }
    const long long PCH = (long long)c->pair_G * PAIR_TPL;
    for (int p = 0; p < c->n_pair; ++p) {
        for (long long b = tptr[p]; b < tptr[p + 1]; b += PCH) { pc_pair.push_back(p); pc_beg.push_back(b); pc_end.push_back(std::min(b + PCH, tptr[p + 1])); }
        pc_ptr.push_back((int)pc_pair.size());
    }
    c->n_pchunk = (int)pc_pair.size();
    c->pchunk_pair = dupload(pc_pair); c->pchunk_beg = dupload(pc_beg); c->pchunk_end = dupload(pc_end);
    c->pair_chunk_ptr = dupload(pc_ptr);

    lap("triple sort + upload");
    // ---- camera system tiles
    psba_build_tile_structure(c, pairs);

    lap("tile structure");
    // ---- work buffers
    const size_t Tl = (size_t)c->N + 3 * (size_t)n;
    c->W = dalloc<double>((size_t)o * 18);
    c->V = dalloc<double>((size_t)n * 6);
    c->Vinv = dalloc<double>((size_t)n * 6);
    c->g = dalloc<double>(Tl); c->dp = dalloc<double>(Tl); c->eab = dalloc<double>(Tl);
    c->P_U = dalloc<double>(Tl); c->P_B = dalloc<double>(Tl); c->P = dalloc<double>(Tl);
    c->cam_part = dalloc<double>((size_t)c->n_cchunk * 27);
    c->pair_part = dalloc<double>((size_t)c->n_pchunk * 42);
    c->d_part = dalloc<double>(((size_t)cdiv(o, 128) + c->n_ptchunk + 512) * 8);
    c->chol_aux = dalloc<double>((size_t)3 * c->N + 2 * TS);
    c->chol_diag = dalloc<double>((size_t)3 * c->N + 2 * TS);
    c->chol_E = dalloc<double>((size_t)c->N + TS);
    c->UVdiag_scr = dalloc<double>(Tl);
    CUDA_CHECK(cudaDeviceSynchronize());
    lap("work buffers");
}

extern "C" void psba_release_buffer(psba_ctx *c)
{
    if (!c) return;
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    void *ptrs[] = {c->K, c->initcams, c->impts, c->cams[0], c->cams[1], c->pts[0], c->pts[1], c->camcache[0], c->camcache[1],
                    c->iidx, c->jidx, c->pt_ptr, c->ptchunk, c->cam_obs, c->cchunk_cam, c->cchunk_beg, c->cchunk_end,
                    c->cam_cchunk_ptr, c->tri_oa, c->tri_ob, c->pair_k, c->pair_l, c->pair_chunk_ptr, c->pchunk_pair,
                    c->pchunk_beg, c->pchunk_end, c->W, c->V, c->Vinv, c->U, c->g, c->UVdiag_scr, c->cam_part, c->pair_part,
                    c->tile_index, c->Stiles, c->Linv, c->eab, c->dp, c->d_status, c->Ldiag, c->cam2pos, c->pos2cam, c->d_crit_I, c->d_crit_K,
                    c->d_psrc_ptr, c->d_psrc, c->d_b_J, c->d_b_sptr, c->d_b_slot, c->d_def_I, c->d_def_J, c->d_def_sptr, c->d_def_src,
                    c->d_step_panels, c->contrib, c->d_coltile_ptr, c->d_coltile_row, c->d_coltile_slot,
                    c->Sdense, c->Sdense_aux, c->chol_aux, c->chol_diag, c->chol_E, c->d_part, c->d_scal, c->P_U, c->P_B, c->P,
                    c->tmpA, c->tmpB};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (c->chol_graph_ok) cudaGraphExecDestroy(c->chol_graph);
    if (c->bw_graph_ok) cudaGraphExecDestroy(c->bw_graph);
    cudaFreeHost(c->h_scal);
    cudaStreamDestroy(c->stream);
    g_stage.erase(c);
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
    if (ex) { if (!c->tmpA) c->tmpA = dalloc<double>((size_t)c->o * 18); exd = c->tmpA; }
    const double cost = psba_launch_cost(c, set, exd);
    if (ex) d2h(c, ex, exd, (size_t)c->o * 2);
    return cost;
}

extern "C" void psba_compute_jacobiQT(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *jac_A, double *jac_B)
{
    check_dims(cnp, pnp, mnp); (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    c->lin_valid = false;            // a new linearisation point; products are rebuilt on demand
    if (jac_A || jac_B) {
        if (!c->tmpA) c->tmpA = dalloc<double>((size_t)c->o * 18);
        if (!c->tmpB) c->tmpB = dalloc<double>((size_t)c->o * 18);
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
        if (!c->tmpB) c->tmpB = dalloc<double>((size_t)c->o * 18);
        psba_launch_Y_materialize(c, c->tmpB);
        d2h(c, Yblks, c->tmpB, (size_t)c->o * 18);
    }
}

static void ensure_dense(psba_ctx *c)
{
    const size_t nn = (size_t)c->N * c->N;
    if (!c->Sdense) c->Sdense = dalloc<double>(nn);
    if (!c->Sdense_aux) c->Sdense_aux = dalloc<double>(nn);
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
        double *tmp = dalloc<double>((size_t)c->N * c->N);
        psba_launch_explicit_inverse(c, tmp);
        d2h(c, outMat, tmp, (size_t)c->N * c->N);
        cudaFree(tmp);
    }
    return ret;
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
    if (out) { if (!c->tmpA) c->tmpA = dalloc<double>((size_t)c->o * 18); jx = c->tmpA; }
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
    ensure_dense(c);
    psba_tiles_to_dense(c, c->Sdense, true);
    const double sum = psba_launch_cholmod(c, delta, beta, n_scalar_blocks);
    if (E) d2h(c, E, c->chol_E, (size_t)c->N);
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

extern "C" void psba_set_option(psba_ctx *c, const char *name, double v)
{
    std::string s(name);
    if (s == "verbose") c->verbose = (int)v;
    else if (s == "max_iter") c->max_iter = (int)v;
    else if (s == "itno") c->itno = (int)v;
    else if (s == "lm_only") c->lm_only = (int)v;
    else if (s == "profile") { psba_prof_collect(c); c->profile = v != 0; }
    else if (s == "profile_reset") { psba_prof_collect(c); for (int k = 0; k < KID_COUNT; ++k) { c->prof_ms[k] = 0; c->prof_n[k] = 0; } }
    else if (s == "stats_reset") { c->st_tries = c->st_exqt = c->st_lin = c->st_launches = 0; }
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
    if (s == "nt") return c->nt;
    if (s == "n_steps") return c->n_steps;
    if (s == "n_ptchunk") return c->n_ptchunk;
    if (s == "n_cchunk") return c->n_cchunk;
    if (s == "n_pchunk") return c->n_pchunk;
    if (s == "pair_G") return c->pair_G;
    if (s == "cholmod_events") return c->n_cholmod_events;
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

extern "C" void psba_try_step(psba_ctx *c, double mu, psba_try_result *res)
{
    if (!c->lin_valid) die("try_step before linearize");
    c->st_tries += 1;
    psba_launch_schur(c, mu);
    res->solve_status = psba_launch_factor(c);
    res->cost_new = res->dp_L2 = res->dp_dot = NAN;
    if (res->solve_status != 0.0) return;
    psba_launch_solve(c);
    psba_launch_backsub(c, mu, true, res);
}
