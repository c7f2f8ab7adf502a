// structure.cu -- index structure of a problem, built ON THE DEVICE.
//
// Replaces generate_idxs (PSBA/misc.cpp:178-218), whose dense tables blk_idx[n*m] and comm3DIdx[m*m*n]
// (O(n m^2) host loop, 14.6 TB at the synthetic size) become
//   pt_ptr   : CSR by point over the point-major observation list (misc.cpp:189-197 order),
//   cam_obs  : the same observations in camera-major order, ascending point (stable sort by camera),
//   triples  : every (obs_k, obs_l) with both observations on one point and camera k >= camera l, sorted by
//              camera pair with ascending point inside a pair -- the enumeration order of comm3DIdx
//              (misc.cpp:199-209) restricted to the lower triangle the solvers read,
//   pairs    : the camera pairs present GLOBALLY (every rank builds the same S layout) + every diagonal.
// Sorting and scanning use CUB (library radix sort / scan / run-length encode, as cuBLAS would be used for
// a plain GEMM); everything else is written here.  Nothing in this file is on the per-iteration path.
#include "psba_internal.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <chrono>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <cstdint>

typedef unsigned long long u64;

static int bits_for(u64 maxval)
{
    int b = 1;
    while (b < 64 && (maxval >> b)) ++b;
    return b;
}

// ---- small kernels -------------------------------------------------------------------------------
// observations must be point-major with ascending cameras (generate_idxs, misc.cpp:189-217)
__global__ void k_check_order(int o, int n, int m, const int *__restrict__ iidx, const int *__restrict__ jidx, int *__restrict__ flag)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= o) return;
    const int i = iidx[k], j = jidx[k];
    bool bad = i < 0 || i >= n || j < 0 || j >= m;
    if (k > 0) { const int ip = iidx[k - 1]; bad |= i < ip || (i == ip && j <= jidx[k - 1]); }
    if (bad) *flag = 1;
}

// ptr[v] = first position k with key[k] >= v, for a sorted key array (ptr has nv + 1 entries)
__global__ void k_segment_ptr(int cnt, int nv, const int *__restrict__ key, int *__restrict__ ptr)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    // keys are clamped: the order check (k_check_order) runs in the same stream and its verdict is read afterwards
    const int cur = min(max(key[k], 0), nv - 1), prev = k > 0 ? min(max(key[k - 1], 0), nv - 1) : -1;
    for (int v = prev + 1; v <= cur; ++v) ptr[v] = k;
    if (k == cnt - 1) for (int v = cur + 1; v <= nv; ++v) ptr[v] = cnt;
}

__global__ void k_local_idx(int o, int o0, int p0, const int *__restrict__ gi, const int *__restrict__ gj, int *__restrict__ li, int *__restrict__ lj, int *__restrict__ iota)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= o) return;
    li[k] = gi[o0 + k] - p0; lj[k] = gj[o0 + k]; iota[k] = k;
}

__global__ void k_local_ptr(int n1, int p0, int o0, const int *__restrict__ gptr, int *__restrict__ ptr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1) ptr[i] = gptr[p0 + i] - o0;
}

// number of lower-triangle triples of every point of [p0, p0 + np)
__global__ void k_triple_count(int np, int p0, const int *__restrict__ gptr, long long *__restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const long long d = gptr[p0 + i + 1] - gptr[p0 + i];
    cnt[i] = d * (d + 1) / 2;
}

// key = k*m + l (k >= l: cameras of the two observations), value = (obs_k << 32 | obs_l), obs ids relative to o0
__global__ void k_triple_emit(int np, int p0, int o0, int m, const int *__restrict__ gptr, const int *__restrict__ gj,
                              const long long *__restrict__ off, u64 *__restrict__ key, u64 *__restrict__ val)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const int a0 = gptr[p0 + i], a1 = gptr[p0 + i + 1];
    long long t = off[i];
    for (int a = a0; a < a1; ++a) {
        const u64 ka = (u64)gj[a] * (u64)m;
        for (int b = a0; b <= a; ++b, ++t) {
            key[t] = ka + (u64)gj[b];
            if (val) val[t] = ((u64)(unsigned)(a - o0) << 32) | (u64)(unsigned)(b - o0);
        }
    }
}

__global__ void k_split_vals(long long cnt, const u64 *__restrict__ val, const int *__restrict__ iidx_loc, int *__restrict__ oa,
                             int *__restrict__ ob, int *__restrict__ pt)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    const u64 v = val[t];
    oa[t] = (int)(v >> 32); ob[t] = (int)(v & 0xffffffffu);
    pt[t] = iidx_loc[(int)(v >> 32)];                 // local point of the triple: the pair pass never gathers iidx
}

// tptr[p] = first triple whose key is >= the key of pair p (binary search in the sorted local keys)
__global__ void k_pair_ptr(int n_pair, int m, const int *__restrict__ pk, const int *__restrict__ pl, long long ntri,
                           const u64 *__restrict__ keys, long long *__restrict__ tptr)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_pair) return;
    if (p == n_pair) { tptr[p] = ntri; return; }
    const u64 want = (u64)pk[p] * (u64)m + (u64)pl[p];
    long long lo = 0, hi = ntri;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (keys[mid] < want) lo = mid + 1; else hi = mid; }
    tptr[p] = lo;
}

__global__ void k_chunk_count(int n_pair, long long pch, const long long *__restrict__ tptr, int *__restrict__ cnt, int *__restrict__ nonempty)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pair) return;
    const long long c = tptr[p + 1] - tptr[p];
    cnt[p] = (int)((c + pch - 1) / pch);
    if (c > 0 && nonempty) atomicAdd(nonempty, 1);
}

__global__ void k_chunk_fill(int n_pair, long long pch, const long long *__restrict__ tptr, const int *__restrict__ cptr,
                             int *__restrict__ ch_pair, long long *__restrict__ ch_beg, long long *__restrict__ ch_end)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pair) return;
    int q = cptr[p];
    for (long long b = tptr[p]; b < tptr[p + 1]; b += pch, ++q) { ch_pair[q] = p; ch_beg[q] = b; ch_end[q] = min(b + pch, tptr[p + 1]); }
}

__global__ void k_camera_major_copy(int o, const int *__restrict__ cam_obs, const int *__restrict__ iidx, const double *__restrict__ impts,
                                    int *__restrict__ cam_pt, double *__restrict__ cam_impts)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= o) return;
    const int k = cam_obs[t];
    cam_pt[t] = iidx[k];
    reinterpret_cast<double2 *>(cam_impts)[t] = reinterpret_cast<const double2 *>(impts)[k];
}


// ---- segment kernel of the pair pass (kernels_schur.cu: k_schur_segs) --------------------------------------
// A "visit" is an observation in camera-major order: camera k looks at point i.  The visits of a camera are cut into
// segments of equal length (<= seg_v); a chunk is the set of triples of ONE camera pair whose visit (the observation of
// camera k) lies in ONE segment -- a contiguous piece of the pair's run in the pair-sorted triple list.
struct seg_desc_h { int row, v0, v1, diag_chunk, sched0, sched1; };

__global__ void k_inverse_perm(int o, const int *__restrict__ perm, int *__restrict__ inv)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < o) inv[perm[v]] = v;
}

// per triple: global segment id, rank of its visit in the segment, head flag of its (pair, segment) chunk
__global__ void k_tri_segment(long long ntri, const int *__restrict__ tri_oa, const int *__restrict__ tri_ob, const int *__restrict__ jidx,
                              const int *__restrict__ cam_pos, const int *__restrict__ cam_ptr, const int *__restrict__ seglen,
                              const int *__restrict__ row_seg_ptr, int *__restrict__ tri_seg, unsigned short *__restrict__ tri_vr,
                              int *__restrict__ head)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntri) return;
    auto seg_of = [&](long long u, int &k, int &l, int &rank) {
        const int a = tri_oa[u];
        k = jidx[a]; l = jidx[tri_ob[u]];
        const int rel = cam_pos[a] - cam_ptr[k], sl = seglen[k];
        rank = rel % sl;
        return row_seg_ptr[k] + rel / sl;
    };
    int k, l, rank, kp, lp, rp;
    const int sg = seg_of(t, k, l, rank);
    tri_seg[t] = sg; tri_vr[t] = (unsigned short)rank;
    int h = 1;
    if (t > 0) { const int sp = seg_of(t - 1, kp, lp, rp); h = (sp != sg || kp != k || lp != l) ? 1 : 0; }
    head[t] = h;
}

// chunk records at the head triples: first triple, segment, diagonal flag; the end of the previous chunk
__global__ void k_chunk_heads(long long ntri, const int *__restrict__ head, const int *__restrict__ cid, const int *__restrict__ tri_seg,
                              const int *__restrict__ tri_oa, const int *__restrict__ tri_ob, int n_pair, const long long *__restrict__ tptr,
                              int *__restrict__ ch_beg, int *__restrict__ ch_end, int *__restrict__ ch_seg, int *__restrict__ ch_pair,
                              int *__restrict__ ch_diag)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntri) return;
    if (t == ntri - 1) ch_end[cid[t] - 1] = (int)ntri;
    if (!head[t]) return;
    const int c = cid[t] - 1;                               // cid is the inclusive scan of the head flags
    ch_beg[c] = (int)t;
    if (c > 0) ch_end[c - 1] = (int)t;
    ch_seg[c] = tri_seg[t];
    ch_diag[c] = tri_oa[t] == tri_ob[t] ? 1 : 0;
    int lo = 0, hi = n_pair;                                // pair p with tptr[p] <= t < tptr[p + 1]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tptr[mid] <= t) lo = mid; else hi = mid; }
    ch_pair[c] = lo;
}

// schedule key: off-diagonal chunks segment by segment, largest first; diagonal chunks (handled by phase 1 of their
// segment) behind everything else
__global__ void k_chunk_keys(int n_chunk, int n_seg, const int *__restrict__ ch_beg, const int *__restrict__ ch_end, const int *__restrict__ ch_seg,
                             const int *__restrict__ ch_diag, u64 *__restrict__ key, int *__restrict__ val, int *__restrict__ seg_diag)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunk) return;
    val[c] = c;
    if (ch_diag[c]) { key[c] = ((u64)n_seg << 32); seg_diag[ch_seg[c]] = c; }
    else key[c] = ((u64)ch_seg[c] << 32) | (u64)(0x7fffffff - (ch_end[c] - ch_beg[c]));
}

// ptr[s] = first position whose key belongs to segment >= s (s = 0..n_seg; keys sorted)
__global__ void k_sched_ptr(int n_chunk, int n_seg, const u64 *__restrict__ key, int *__restrict__ ptr)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n_chunk) return;
    const int cur = j < n_chunk ? (int)(key[j] >> 32) : n_seg, prev = j > 0 ? (int)(key[j - 1] >> 32) : -1;
    for (int v = prev + 1; v <= cur; ++v) ptr[v] = j;
}

__global__ void k_fill_seg_desc(int n_seg, seg_desc_h *__restrict__ d, const int *__restrict__ seg_diag, const int *__restrict__ sptr)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    d[s].diag_chunk = seg_diag[s]; d[s].sched0 = sptr[s]; d[s].sched1 = sptr[s + 1];
}


// ---- ring kernel of the pair pass (kernels_schur.cu: k_schur_ring) -----------------------------------------
// The off-diagonal chunks of a segment (sorted largest first) become TASKS: the next 32/G chunks share a warp, G lanes
// each (G = the power of two that brings the largest chunk of the task down to ~rt rows), one ROW = 32 lane slots = one
// triple per lane.  Tasks are dealt to the nw warps of the CTA, largest first to the least loaded warp; the rows of a warp
// are contiguous so that the warp streams them through its copy ring without looking at task boundaries.
__device__ __forceinline__ int ring_lanes(int len, int rt)
{
    const int want = (len + rt - 1) / rt;
    int G = 1;
    while (G < 32 && G < want) G *= 2;
    return G;
}

// thread per segment.  fill = 0: rows of every (segment, warp) -> wrows;  fill = 1: task records at the position of the
// task's first chunk in the schedule: {first row, rows, chunks, log2 G} (rows = 0 at the other chunks of a task)
__global__ void k_ring_plan(int n_seg, const seg_desc_h *__restrict__ segs, const int *__restrict__ sched, const int *__restrict__ ch_beg,
                            const int *__restrict__ ch_end, int nw, int rt, int fill, int *__restrict__ wrows,
                            const int *__restrict__ wrow_ptr, int4 *__restrict__ task)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    int load[16];
    for (int w = 0; w < 16; ++w) load[w] = 0;
    const int j0 = segs[s].sched0, j1 = segs[s].sched1;
    for (int j = j0; j < j1;) {
        const int c = sched[j], len = ch_end[c] - ch_beg[c];
        const int G = ring_lanes(len, rt), nch = min(32 / G, j1 - j), nrows = (len + G - 1) / G;
        int w = 0;
        for (int q = 1; q < nw; ++q) if (load[q] < load[w]) w = q;
        if (fill) {
            int lg = 0;
            while ((1 << lg) < G) ++lg;
            task[j] = make_int4(wrow_ptr[(size_t)s * nw + w] + load[w], nrows, nch, lg);
            for (int q = 1; q < nch; ++q) task[j + q] = make_int4(0, 0, 0, 0);
        }
        load[w] += nrows;
        j += nch;
    }
    if (!fill) for (int w = 0; w < nw; ++w) wrows[(size_t)s * nw + w] = load[w];
}

// warp per schedule position that starts a task: the rows of the task
__global__ void k_ring_rows(int n_sched, const int4 *__restrict__ task, const int *__restrict__ sched, const int *__restrict__ ch_beg,
                            const int *__restrict__ ch_end, const unsigned short *__restrict__ tri_vr, const int *__restrict__ tri_ob,
                            int2 *__restrict__ rows)
{
    const int j = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (j >= n_sched) return;
    const int4 tk = task[j];
    if (tk.y == 0) return;
    const int G = 1 << tk.w, q = lane >> tk.w, gl = lane & (G - 1);
    int beg = 0, end = 0;
    if (q < tk.z) { const int c = sched[j + q]; beg = ch_beg[c]; end = ch_end[c]; }
    for (int it = 0; it < tk.y; ++it) {
        const int t = beg + gl + it * G;
        rows[((size_t)tk.x + it) * 34 + lane] = t < end ? make_int2(tri_ob[t], (int)tri_vr[t]) : make_int2(-1, 0);
        if (lane == 0) {       // slot 32: the task word of the row; slot 33: pad (a row record is seventeen 16-byte pieces)
            rows[((size_t)tk.x + it) * 34 + 32] = make_int2(tk.w | ((it == tk.y - 1) ? 256 : 0) | (tk.z << 16), j);
            rows[((size_t)tk.x + it) * 34 + 33] = make_int2(0, 0);
        }
    }
}

// ---- helpers ---------------------------------------------------------------------------------------
template <class T> static T *salloc(psba_ctx *c, size_t n)
{
    return (T *)psba_dev_alloc(c, std::max<size_t>(n, 1) * sizeof(T), false);
}
template <class T> static T *supload(psba_ctx *c, const std::vector<T> &h)
{
    T *p = salloc<T>(c, h.size());
    if (!h.empty()) CUDA_CHECK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return p;
}

struct lap_timer {
    bool on; psba_ctx *c;
    std::chrono::steady_clock::time_point t;
    lap_timer(psba_ctx *c_) : on(getenv("PSBA_SETUP_TIMING") != nullptr), c(c_), t(std::chrono::steady_clock::now()) {}
    void lap(const char *what)
    {
        if (!on) return;
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "psba setup: %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

// sorted (key, value) triples of the points [p0, p0+np); returns device arrays the caller frees
static void sorted_triples(psba_ctx *c, int np, int p0, int o0, const int *gptr, const int *gj, bool want_vals,
                           u64 **keys_out, u64 **vals_out, long long *ntri_out)
{
    cudaStream_t st = c->stream;
    long long *cnt = salloc<long long>(c, (size_t)np + 1), *off = salloc<long long>(c, (size_t)np + 1);
    CUDA_CHECK(cudaMemsetAsync(cnt, 0, ((size_t)np + 1) * sizeof(long long), st));
    if (np > 0) k_triple_count<<<cdiv(np, 256), 256, 0, st>>>(np, p0, gptr, cnt);
    size_t tb = 0;
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, off, np + 1, st));
    void *tmp = psba_dev_alloc(c, std::max<size_t>(tb, 16), false);
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, off, np + 1, st));
    long long ntri = 0;
    CUDA_CHECK(cudaMemcpyAsync(&ntri, off + np, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    psba_dev_free(c, tmp); psba_dev_free(c, cnt);
    if (ntri >= (1ll << 31)) { fprintf(stderr, "psba_b200: more than 2^31 camera-pair triples on one GPU\n"); exit(EXIT_FAILURE); }
    u64 *k0 = salloc<u64>(c, (size_t)ntri), *k1 = salloc<u64>(c, (size_t)ntri);
    u64 *v0 = want_vals ? salloc<u64>(c, (size_t)ntri) : nullptr, *v1 = want_vals ? salloc<u64>(c, (size_t)ntri) : nullptr;
    if (np > 0) k_triple_emit<<<cdiv(np, 128), 128, 0, st>>>(np, p0, o0, c->m, gptr, gj, off, k0, v0);
    psba_dev_free(c, off);
    const int bits = bits_for((u64)c->m * (u64)c->m);
    tb = 0;
    if (want_vals) CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, v1, (int)ntri, 0, bits, st));
    else CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tb, k0, k1, (int)ntri, 0, bits, st));
    tmp = psba_dev_alloc(c, std::max<size_t>(tb, 16), false);
    if (ntri > 0) {
        // LSD radix sort is stable: triples of one pair keep their emission order = ascending point
        if (want_vals) CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, v1, (int)ntri, 0, bits, st));
        else CUDA_CHECK(cub::DeviceRadixSort::SortKeys(tmp, tb, k0, k1, (int)ntri, 0, bits, st));
    }
    psba_dev_free(c, tmp); psba_dev_free(c, k0);
    if (v0) psba_dev_free(c, v0);
    *keys_out = k1; *vals_out = v1; *ntri_out = ntri;
}


static int ring_seg_v(const psba_ctx *c);
// tables of the segment kernel: segments of every camera row, (pair, segment) chunks, schedule
static void build_segments(psba_ctx *c, const long long *tptr, const std::vector<int> &cptr)
{
    cudaStream_t st = c->stream;
    const int m = c->m, o = c->o;
    c->n_seg = 0; c->n_pchunk = 0;
    c->seg_cfg = 2;
    c->seg_v = 1280;                                         // Y tile: 144 B per visit, one resident CTA per SM (kernels_schur.cu: launch shape)
    if (c->pair_mode == 6) c->seg_v = ring_seg_v(c);
    // small problems: enough segments to fill the machine four times over (Venice-52: 312 segments of 1 280 visits were three
    // uneven waves of one CTA per SM)
    if ((long long)o < (long long)c->seg_v * 4 * c->n_sm) c->seg_v = std::max(128, std::min(c->seg_v, cdiv(o, 4 * c->n_sm)));
    if (getenv("PSBA_SEG_V")) c->seg_v = std::max(32, std::min(c->pair_mode == 6 ? ring_seg_v(c) : SEG_V_MAX, atoi(getenv("PSBA_SEG_V"))));
    // ---- segments (host: m cameras)
    std::vector<seg_desc_h> segs;
    std::vector<int> row_seg_ptr((size_t)m + 1, 0), seglen(m, 1);
    for (int k = 0; k < m; ++k) {
        const int cnt = cptr[k + 1] - cptr[k];
        if (cnt > 0) {
            const int ns = cdiv(cnt, c->seg_v), sl = cdiv(cnt, ns);
            seglen[k] = sl;
            for (int b = 0; b < cnt; b += sl) segs.push_back({k, cptr[k] + b, cptr[k] + std::min(b + sl, cnt), -1, 0, 0});
        }
        row_seg_ptr[k + 1] = (int)segs.size();
    }
    c->n_seg = (int)segs.size();
    c->pair_chunk_ptr = salloc<int>(c, (size_t)c->n_pair + 1);
    CUDA_CHECK(cudaMemsetAsync(c->pair_chunk_ptr, 0, ((size_t)c->n_pair + 1) * 4, st));
    c->tri_vr = salloc<unsigned short>(c, (size_t)c->ntri + 8);
    c->seg_desc = supload(c, segs);
    if (c->ntri == 0 || c->n_seg == 0) {
        c->sched_chunk = salloc<int>(c, 1); c->sch_beg = salloc<int>(c, 1); c->sch_end = salloc<int>(c, 1);
        CUDA_CHECK(cudaStreamSynchronize(st));
        return;
    }
    int *cam_pos = salloc<int>(c, o), *cam_ptr = supload(c, cptr), *d_seglen = supload(c, seglen), *d_rsp = supload(c, row_seg_ptr);
    k_inverse_perm<<<cdiv(o, 256), 256, 0, st>>>(o, c->cam_obs, cam_pos);
    int *tri_seg = salloc<int>(c, (size_t)c->ntri), *head = salloc<int>(c, (size_t)c->ntri), *cid = salloc<int>(c, (size_t)c->ntri);
    k_tri_segment<<<cdiv(c->ntri, 256), 256, 0, st>>>(c->ntri, c->tri_oa, c->tri_ob, c->jidx, cam_pos, cam_ptr, d_seglen, d_rsp, tri_seg, c->tri_vr, head);
    {
        size_t tb = 0;
        CUDA_CHECK(cub::DeviceScan::InclusiveSum(nullptr, tb, head, cid, (int)c->ntri, st));
        void *tmp = psba_dev_alloc(c, std::max<size_t>(tb, 16), false);
        CUDA_CHECK(cub::DeviceScan::InclusiveSum(tmp, tb, head, cid, (int)c->ntri, st));
        psba_dev_free(c, tmp);
    }
    int n_chunk = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n_chunk, cid + (c->ntri - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));                 // also: the host vectors behind the uploads above are done with
    c->n_pchunk = n_chunk;
    c->sch_beg = salloc<int>(c, n_chunk); c->sch_end = salloc<int>(c, n_chunk);
    int *ch_seg = salloc<int>(c, n_chunk), *ch_pair = salloc<int>(c, n_chunk), *ch_diag = salloc<int>(c, n_chunk);
    k_chunk_heads<<<cdiv(c->ntri, 256), 256, 0, st>>>(c->ntri, head, cid, tri_seg, c->tri_oa, c->tri_ob, c->n_pair, tptr, c->sch_beg, c->sch_end,
                                                     ch_seg, ch_pair, ch_diag);
    k_segment_ptr<<<cdiv(n_chunk, 256), 256, 0, st>>>(n_chunk, c->n_pair, ch_pair, c->pair_chunk_ptr);
    // ---- schedule
    u64 *key0 = salloc<u64>(c, n_chunk), *key1 = salloc<u64>(c, n_chunk);
    int *val0 = salloc<int>(c, n_chunk), *seg_diag = salloc<int>(c, c->n_seg), *sptr = salloc<int>(c, (size_t)c->n_seg + 1);
    c->sched_chunk = salloc<int>(c, n_chunk);
    k_chunk_keys<<<cdiv(n_chunk, 256), 256, 0, st>>>(n_chunk, c->n_seg, c->sch_beg, c->sch_end, ch_seg, ch_diag, key0, val0, seg_diag);
    {
        size_t tb = 0;
        const int bits = 32 + bits_for((u64)c->n_seg);
        CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, key0, key1, val0, c->sched_chunk, n_chunk, 0, bits, st));
        void *tmp = psba_dev_alloc(c, std::max<size_t>(tb, 16), false);
        CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tb, key0, key1, val0, c->sched_chunk, n_chunk, 0, bits, st));
        psba_dev_free(c, tmp);
    }
    k_sched_ptr<<<cdiv(n_chunk + 1, 256), 256, 0, st>>>(n_chunk, c->n_seg, key1, sptr);
    k_fill_seg_desc<<<cdiv(c->n_seg, 256), 256, 0, st>>>(c->n_seg, (seg_desc_h *)c->seg_desc, seg_diag, sptr);
    // lanes per chunk from the mean size of an off-diagonal chunk (a lane should stream ~8 triples before the butterfly)
    {
        const double avg = (double)(c->ntri - o) / std::max(1, n_chunk - c->n_seg);
        int G = 1;
        while (G < 32 && avg > 5.0 * G) G *= 2;
        if (getenv("PSBA_SEG_G")) G = std::max(1, std::min(32, atoi(getenv("PSBA_SEG_G"))));
        while (G & (G - 1)) G &= G - 1;
        c->pair_G = G;
    }
    // the temporaries are freed in stream order (cudaFreeAsync): no host round trip here
    for (void *p : {(void *)cam_pos, (void *)cam_ptr, (void *)d_seglen, (void *)d_rsp, (void *)tri_seg, (void *)head, (void *)cid, (void *)ch_seg,
                    (void *)ch_pair, (void *)ch_diag, (void *)key0, (void *)key1, (void *)val0, (void *)seg_diag, (void *)sptr})
        psba_dev_free(c, p);
}


// launch shapes of the ring kernel: {threads, stages per warp, CTAs per SM}; the Y tile takes what the copy rings and the row
// records leave of the SM's shared memory
static const int RING_SHAPES[6][3] = {{384, 2, 1}, {256, 2, 1}, {256, 3, 1}, {128, 2, 2}, {128, 3, 2}, {192, 2, 1}};
static size_t ring_fixed_smem(int cfg)
{
    const int nw = RING_SHAPES[cfg][0] / 32, st = RING_SHAPES[cfg][1];
    return (size_t)nw * st * 4608 + (size_t)nw * (2 * st + 1) * 272;
}
size_t psba_ring_smem(int cfg, int seg_v) { return (size_t)seg_v * 144 + ring_fixed_smem(cfg); }

static void ring_config(psba_ctx *c)
{
    c->ring_cfg = 3;                                         // measured on the headline workload: 2 x 128 threads per SM 0.73 ms, 256 x 1 0.75-0.77, 384 x 1 0.87
    if (getenv("PSBA_RING_CFG")) c->ring_cfg = std::max(0, std::min(5, atoi(getenv("PSBA_RING_CFG"))));
    c->ring_nw = RING_SHAPES[c->ring_cfg][0] / 32; c->ring_stages = RING_SHAPES[c->ring_cfg][1];
    c->ring_rt = 5;
    if (getenv("PSBA_RING_RT")) c->ring_rt = std::max(1, std::min(64, atoi(getenv("PSBA_RING_RT"))));
}
// visits per segment: the Y tile beside the copy rings (227 KB per CTA, 228 KB per SM less 1 KB per resident CTA; 4 KB kept
// for the static arrays)
static int ring_seg_v(const psba_ctx *c)
{
    const int per_sm = RING_SHAPES[c->ring_cfg][2];
    const long long cta = per_sm == 1 ? 227ll * 1024 : (228ll * 1024) / per_sm - 1024;
    const long long avail = cta - 4096 - (long long)ring_fixed_smem(c->ring_cfg);
    return (int)std::min<long long>(SEG_V_MAX, avail / 144);
}

static void build_ring_tables(psba_ctx *c)
{
    cudaStream_t st = c->stream;
    const int nw = c->ring_nw, n_sched = c->n_pchunk;
    c->ring_n_rows = 0;
    c->ring_wrow_ptr = salloc<int>(c, (size_t)c->n_seg * nw + 1);
    if (c->n_seg == 0 || c->ntri == 0) {
        CUDA_CHECK(cudaMemsetAsync(c->ring_wrow_ptr, 0, ((size_t)c->n_seg * nw + 1) * 4, st));
        c->ring_rows = salloc<int2>(c, 34); c->ring_info = nullptr;
        return;
    }
    const size_t nwr = (size_t)c->n_seg * nw + 1;
    int *wrows = salloc<int>(c, nwr);
    CUDA_CHECK(cudaMemsetAsync(wrows, 0, nwr * 4, st));
    int4 *task = salloc<int4>(c, (size_t)n_sched);
    CUDA_CHECK(cudaMemsetAsync(task, 0, (size_t)n_sched * sizeof(int4), st));     // positions of the diagonal chunks start no task
    k_ring_plan<<<cdiv(c->n_seg, 128), 128, 0, st>>>(c->n_seg, (const seg_desc_h *)c->seg_desc, c->sched_chunk, c->sch_beg, c->sch_end, nw, c->ring_rt, 0,
                                                     wrows, nullptr, nullptr);
    size_t tb = 0;
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, wrows, c->ring_wrow_ptr, (int)nwr, st));
    void *tmp = psba_dev_alloc(c, std::max<size_t>(tb, 16), false);
    CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb, wrows, c->ring_wrow_ptr, (int)nwr, st));
    int n_rows = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n_rows, c->ring_wrow_ptr + (nwr - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    psba_dev_free(c, tmp);
    c->ring_n_rows = n_rows;
    c->ring_rows = salloc<int2>(c, (size_t)n_rows * 34 + 34); c->ring_info = nullptr;
    k_ring_plan<<<cdiv(c->n_seg, 128), 128, 0, st>>>(c->n_seg, (const seg_desc_h *)c->seg_desc, c->sched_chunk, c->sch_beg, c->sch_end, nw, c->ring_rt, 1,
                                                     nullptr, c->ring_wrow_ptr, task);
    k_ring_rows<<<cdiv((long long)n_sched * 32, 256), 256, 0, st>>>(n_sched, task, c->sched_chunk, c->sch_beg, c->sch_end, c->tri_vr, c->tri_ob,
                                                                   c->ring_rows);
    psba_dev_free(c, wrows); psba_dev_free(c, task);       // stream-ordered frees: no host round trip
}


// one helper thread per process for host work that can run under device work of the same set-up (created once; the jobs it
// gets issue no CUDA calls -- two threads feeding one stream serialise on each other's blocking copies)
struct host_worker {
    std::thread th; std::mutex mu; std::condition_variable cv;
    std::function<void()> job; bool busy = false, quit = false;
    static host_worker &get() { static host_worker w; return w; }
    host_worker() { th = std::thread([this] { loop(); }); }
    ~host_worker() { { std::lock_guard<std::mutex> l(mu); quit = true; } cv.notify_all(); th.join(); }
    void loop()
    {
        std::unique_lock<std::mutex> l(mu);
        for (;;) {
            cv.wait(l, [this] { return quit || (busy && job); });
            if (quit) return;
            std::function<void()> j = std::move(job); job = nullptr;
            l.unlock(); j(); l.lock();
            busy = false; cv.notify_all();
        }
    }
    void run(std::function<void()> j) { { std::lock_guard<std::mutex> l(mu); job = std::move(j); busy = true; } cv.notify_all(); }
    void wait() { std::unique_lock<std::mutex> l(mu); cv.wait(l, [this] { return !busy; }); }
};


// page-locked scratch for the point CSR that comes back from the device (4 MB at 1 M points: 0.1 ms instead of 0.5 ms into pageable
// memory); one block per process, handed to one set-up at a time (a second concurrent set-up falls back to pageable memory)
struct pinned_scratch {
    static std::mutex &mu() { static std::mutex m; return m; }
    static int *acquire(size_t n_ints)
    {
        static int *buf = nullptr; static size_t cap = 0;
        std::lock_guard<std::mutex> l(mu());
        if (in_use()) return nullptr;
        if (cap < n_ints) {
            if (buf) cudaFreeHost(buf);
            buf = nullptr; cap = 0;
            const size_t want = n_ints + n_ints / 4 + 1024;
            if (cudaHostAlloc((void **)&buf, want * sizeof(int), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); buf = nullptr; return nullptr; }
            cap = want;
        }
        in_use() = true;
        return buf;
    }
    static void release() { std::lock_guard<std::mutex> l(mu()); in_use() = false; }
    static bool &in_use() { static bool b = false; return b; }
};

// ---- the build ---------------------------------------------------------------------------------------
void psba_build_structure(psba_ctx *c, const int *iidx, const int *jidx)
{
    // one set-up at a time per process: the helper thread has ONE job slot (host_worker) and the page-locked scratch blocks are
    // shared; two contexts set up from two host threads take turns here (the solves themselves run concurrently)
    static std::mutex setup_mu;
    std::lock_guard<std::mutex> setup_lock(setup_mu);
    cudaStream_t st = c->stream;
    const int m = c->m, ng = c->n_glob, og = c->o_glob;
    lap_timer T(c);
    // ---- global index arrays, validation, CSR by point
    int *gi = salloc<int>(c, og), *gj = salloc<int>(c, og), *gptr = salloc<int>(c, (size_t)ng + 1);
    if (og) {
        CUDA_CHECK(cudaMemcpyAsync(gi, iidx, (size_t)og * 4, cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(gj, jidx, (size_t)og * 4, cudaMemcpyHostToDevice, st));
    }
    CUDA_CHECK(cudaMemsetAsync(c->d_status + 3, 0, sizeof(int), st));
    CUDA_CHECK(cudaMemsetAsync(gptr, 0, ((size_t)ng + 1) * 4, st));
    if (og) {
        k_check_order<<<cdiv(og, 256), 256, 0, st>>>(og, ng, m, gi, gj, c->d_status + 3);
        k_segment_ptr<<<cdiv(og, 256), 256, 0, st>>>(og, ng, gi, gptr);
    }
    std::vector<int> hptr_pageable;
    int *hptr = pinned_scratch::acquire((size_t)ng + 1);
    const bool hptr_pinned = hptr != nullptr;
    if (!hptr) { hptr_pageable.resize((size_t)ng + 1); hptr = hptr_pageable.data(); }
    int bad = 0;
    CUDA_CHECK(cudaMemcpyAsync(&bad, c->d_status + 3, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(hptr, gptr, ((size_t)ng + 1) * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (bad) { fprintf(stderr, "psba_b200: fill_idxBuffer: observations are not point-major with ascending cameras\n"); exit(EXIT_FAILURE); }
    T.lap("index upload + CSR");
    // ---- this rank's contiguous point range, balanced by observation count (psba_local_range)
    auto first_pt = [&](int r) -> int {
        if (r >= c->nranks) return ng;
        const long long target = (long long)og * r / c->nranks;
        return (int)(std::lower_bound(hptr, hptr + ng, (int)target) - hptr);
    };
    const int p0 = first_pt(c->rank), p1 = first_pt(c->rank + 1), o0 = hptr[p0], o1 = hptr[p1];
    c->p_off = p0; c->o_off = o0; c->n = p1 - p0; c->o = o1 - o0;
    const int n = c->n, o = c->o;
    c->iidx = salloc<int>(c, o); c->jidx = salloc<int>(c, o); c->pt_ptr = salloc<int>(c, (size_t)n + 1);
    int *iota = salloc<int>(c, o);
    if (o) k_local_idx<<<cdiv(o, 256), 256, 0, st>>>(o, o0, p0, gi, gj, c->iidx, c->jidx, iota);
    k_local_ptr<<<cdiv(n + 1, 256), 256, 0, st>>>(n + 1, p0, o0, gptr, c->pt_ptr);
    // ---- camera-major order: stable radix sort of the local observations by camera (enqueued first: the host loop over the
    // points below runs under it)
    c->cam_obs = salloc<int>(c, o);
    std::vector<int> cptr((size_t)m + 1, 0);
    int *skey = salloc<int>(c, o), *cp = salloc<int>(c, (size_t)m + 1);
    void *sort_tmp = nullptr;
    {
        CUDA_CHECK(cudaMemsetAsync(cp, 0, ((size_t)m + 1) * 4, st));
        size_t tb = 0;
        const int bits = bits_for((u64)m);
        CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, c->jidx, skey, iota, c->cam_obs, o, 0, bits, st));
        sort_tmp = psba_dev_alloc(c, std::max<size_t>(tb, 16), false);
        if (o) {
            CUDA_CHECK(cub::DeviceRadixSort::SortPairs(sort_tmp, tb, c->jidx, skey, iota, c->cam_obs, o, 0, bits, st));
            k_segment_ptr<<<cdiv(o, 256), 256, 0, st>>>(o, m, skey, cp);
        }
    }
    // point chunks: whole points, <= PT_CTA observations and <= PT_CTA points per CTA (greedy over the n points: a millisecond of
    // host work at 1 M points) -- on the helper thread, under the sorts below; uploaded where the worker is joined
    std::vector<int> pch(1, 0), small_chunks, big_chunks;
    std::vector<int4> ptdesc_h;
    host_worker::get().run([&pch, &small_chunks, &big_chunks, &ptdesc_h, hptr, n, p0, o0]() {
        int cnt_o = 0, cnt_p = 0;
        for (int i = 0; i < n; ++i) {
            const int d = hptr[p0 + i + 1] - hptr[p0 + i];
            if (cnt_p > 0 && (cnt_o + d > PT_CTA || cnt_p == PT_CTA)) { pch.push_back(i); cnt_o = 0; cnt_p = 0; }
            cnt_o += d; cnt_p++;
        }
        if (n > 0) pch.push_back(n);
        const int nch = (int)pch.size() - 1;
        ptdesc_h.resize(nch);                                  // {p0, p1, o0, o1}: one 16-byte load per CTA
        for (int q = 0; q < nch; ++q)
            ptdesc_h[q] = make_int4(pch[q], pch[q + 1], hptr[p0 + pch[q]] - o0, hptr[p0 + pch[q + 1]] - o0);
        // chunks that fit one wave go through the pipelined kernels; a chunk that is one point with more than
        // PT_CTA observations keeps the wave loop
        for (int q = 0; q < nch; ++q) (ptdesc_h[q].w - ptdesc_h[q].z <= PT_CTA ? small_chunks : big_chunks).push_back(q);
    });
    {
        CUDA_CHECK(cudaMemcpyAsync(cptr.data(), cp, ((size_t)m + 1) * 4, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        psba_dev_free(c, sort_tmp); psba_dev_free(c, skey); psba_dev_free(c, cp); psba_dev_free(c, iota);
    }
    std::vector<int> cc_cam, cc_beg, cc_end, cc_ptr(1, 0);
    const int CCH = CAM_CTA * CAM_OPT;
    for (int j = 0; j < m; ++j) {
        for (int b = cptr[j]; b < cptr[j + 1]; b += CCH) { cc_cam.push_back(j); cc_beg.push_back(b); cc_end.push_back(std::min(b + CCH, cptr[j + 1])); }
        cc_ptr.push_back((int)cc_cam.size());
    }
    c->n_cchunk = (int)cc_cam.size();
    c->cchunk_cam = supload(c, cc_cam); c->cchunk_beg = supload(c, cc_beg); c->cchunk_end = supload(c, cc_end);
    c->cam_cchunk_ptr = supload(c, cc_ptr);
    T.lap("local CSR + camera order");
    // ---- triples of the local points, sorted by camera pair
    u64 *lkeys = nullptr, *lvals = nullptr;
    sorted_triples(c, n, p0, o0, gptr, gj, true, &lkeys, &lvals, &c->ntri);
    c->tri_oa = salloc<int>(c, (size_t)c->ntri); c->tri_ob = salloc<int>(c, (size_t)c->ntri); c->tri_pt = salloc<int>(c, (size_t)c->ntri);
    if (c->ntri) k_split_vals<<<cdiv(c->ntri, 256), 256, 0, st>>>(c->ntri, lvals, c->iidx, c->tri_oa, c->tri_ob, c->tri_pt);
    psba_dev_free(c, lvals);
    T.lap("triple sort");
    // ---- global pair set (all ranks agree): unique keys of ALL points' triples + every diagonal block (U_k)
    std::vector<u64> ukeys;
    {
        u64 *gkeys = lkeys, *gv = nullptr;
        long long gtri = c->ntri;
        if (c->nranks > 1) sorted_triples(c, ng, 0, 0, gptr, gj, false, &gkeys, &gv, &gtri);
        u64 *uk = salloc<u64>(c, (size_t)gtri);
        int *ucnt = salloc<int>(c, (size_t)gtri), *nruns = salloc<int>(c, 1);
        size_t tb = 0;
        CUDA_CHECK(cub::DeviceRunLengthEncode::Encode(nullptr, tb, gkeys, uk, ucnt, nruns, (int)gtri, st));
        void *tmp = psba_dev_alloc(c, std::max<size_t>(tb, 16), false);
        int hr = 0;
        if (gtri > 0) {
            CUDA_CHECK(cub::DeviceRunLengthEncode::Encode(tmp, tb, gkeys, uk, ucnt, nruns, (int)gtri, st));
            CUDA_CHECK(cudaMemcpyAsync(&hr, nruns, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
        }
        ukeys.resize(hr);
        if (hr) CUDA_CHECK(cudaMemcpyAsync(ukeys.data(), uk, (size_t)hr * 8, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        psba_dev_free(c, tmp); psba_dev_free(c, uk); psba_dev_free(c, ucnt); psba_dev_free(c, nruns);
        if (c->nranks > 1) psba_dev_free(c, gkeys);
    }
    psba_dev_free(c, gi); psba_dev_free(c, gj); psba_dev_free(c, gptr);
    std::vector<int> pk, pl;
    std::vector<std::pair<int, int>> pairs;
    {
        pk.reserve(ukeys.size() + m); pl.reserve(ukeys.size() + m); pairs.reserve(ukeys.size() + m);
        size_t q = 0;
        auto put = [&](int k, int l) { pk.push_back(k); pl.push_back(l); pairs.push_back({k, l}); };
        for (int j = 0; j < m; ++j) {          // keys are sorted by (k, l): the diagonal is the last pair of row k
            const u64 dkey = (u64)j * (u64)m + (u64)j;
            while (q < ukeys.size() && ukeys[q] < dkey) { put((int)(ukeys[q] / (u64)m), (int)(ukeys[q] % (u64)m)); ++q; }
            if (q < ukeys.size() && ukeys[q] == dkey) ++q;
            put(j, j);
        }
    }
    c->n_pair = (int)pk.size();
    c->pair_k = supload(c, pk); c->pair_l = supload(c, pl);
    // ---- camera system tiles (symbolic factorisation, 2 ms of host work at 2 000 cameras) on a helper thread, under the device
    // work and the host round trips of the segment / ring tables below (disjoint fields of the context, same stream)
    host_worker::get().wait();                                // the point chunks are done (uploaded below, under the next job)
    if (hptr_pinned) pinned_scratch::release();                // nothing reads the point CSR on the host any more
    const auto t_job = std::chrono::steady_clock::now();
    host_worker::get().run([c, &pairs, t_job]() {          // no CUDA call inside: uploads are flushed below
        const auto t_in = std::chrono::steady_clock::now();
        psba_build_tile_structure(c, pairs);
        if (getenv("PSBA_SETUP_TIMING"))
            fprintf(stderr, "psba setup:   tiles: job picked up after %.2f ms, ran %.2f ms\n", std::chrono::duration<double, std::milli>(t_in - t_job).count(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_in).count());
    });
    c->n_ptchunk = (int)pch.size() - 1;
    c->ptchunk = supload(c, pch);
    c->ptdesc = supload(c, ptdesc_h);
    c->n_small = (int)small_chunks.size(); c->n_big = (int)big_chunks.size();
    c->d_small_list = big_chunks.empty() ? nullptr : supload(c, small_chunks);
    c->d_big_list = big_chunks.empty() ? nullptr : supload(c, big_chunks);
    // ---- triple range of every pair, chunks of the pair pass
    long long *tptr = salloc<long long>(c, (size_t)c->n_pair + 1);
    k_pair_ptr<<<cdiv(c->n_pair + 1, 256), 256, 0, st>>>(c->n_pair, m, c->pair_k, c->pair_l, c->ntri, lkeys, tptr);
    // pair pass: 6 (default) = ring kernel, 5 = segment kernel, 0 = the pair-major gather kernel of round 1
    c->pair_mode = 6;
    if (getenv("PSBA_PAIR_MODE")) { const int pm = atoi(getenv("PSBA_PAIR_MODE")); c->pair_mode = pm == 0 ? 0 : (pm == 5 ? 5 : 6); }
    if (c->pair_mode == 6) ring_config(c);
    int *ccnt = salloc<int>(c, (size_t)c->n_pair + 1), *nonempty = salloc<int>(c, 1);
    int h_nonempty = 0;
    if (c->pair_mode == 0) {                                  // mean run length of a pair (the other kernels do not need it: no round trip)
        CUDA_CHECK(cudaMemsetAsync(nonempty, 0, sizeof(int), st));
        CUDA_CHECK(cudaMemsetAsync(ccnt, 0, ((size_t)c->n_pair + 1) * 4, st));
        k_chunk_count<<<cdiv(c->n_pair, 256), 256, 0, st>>>(c->n_pair, 1, tptr, ccnt, nonempty);   // pch = 1: counts triples; only `nonempty` is used
        CUDA_CHECK(cudaMemcpyAsync(&h_nonempty, nonempty, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
    }
    c->n_seg = 0; c->seg_desc = nullptr; c->sched_chunk = nullptr; c->sch_beg = c->sch_end = nullptr; c->tri_vr = nullptr;
    c->pchunk_pair = nullptr; c->pchunk_beg = c->pchunk_end = nullptr;
    if (c->pair_mode == 5 || c->pair_mode == 6) {
        build_segments(c, tptr, cptr);
        T.lap("segments + chunks");
        if (c->pair_mode == 6) { build_ring_tables(c); T.lap("ring rows"); }
    } else {
        // every group of G lanes owns one chunk of <= G*PAIR_TPL triples of one camera pair; G follows the mean run
        // length so that a lane streams ~PAIR_TPL triples before the (shuffle) reduction
        const double avg = h_nonempty ? (double)c->ntri / (double)h_nonempty : 1.0;
        int G = 1;
        while (G < 32 && avg > (double)G * PAIR_TPL) G *= 2;
        c->pair_G = G;
        const long long PCH = (long long)G * PAIR_TPL;
        k_chunk_count<<<cdiv(c->n_pair, 256), 256, 0, st>>>(c->n_pair, PCH, tptr, ccnt, nullptr);
        c->pair_chunk_ptr = salloc<int>(c, (size_t)c->n_pair + 1);
        size_t tb = 0;
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, ccnt, c->pair_chunk_ptr, c->n_pair + 1, st));
        void *tmp = psba_dev_alloc(c, std::max<size_t>(tb, 16), false);
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb, ccnt, c->pair_chunk_ptr, c->n_pair + 1, st));
        CUDA_CHECK(cudaMemcpyAsync(&c->n_pchunk, c->pair_chunk_ptr + c->n_pair, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        psba_dev_free(c, tmp);
        c->pchunk_pair = salloc<int>(c, c->n_pchunk);
        c->pchunk_beg = salloc<long long>(c, c->n_pchunk); c->pchunk_end = salloc<long long>(c, c->n_pchunk);
        k_chunk_fill<<<cdiv(c->n_pair, 256), 256, 0, st>>>(c->n_pair, PCH, tptr, c->pair_chunk_ptr, c->pchunk_pair, c->pchunk_beg, c->pchunk_end);
        T.lap("pair set + chunks");
    }
    psba_dev_free(c, tptr); psba_dev_free(c, ccnt); psba_dev_free(c, nonempty); psba_dev_free(c, lkeys);
    host_worker::get().wait();
    T.lap("tile structure (wait)");
    psba_flush_tile_uploads(c);
    T.lap("tile structure (uploads)");
}

// camera-major copies of the per-observation constants the camera pass reads (point index, measurement)
void psba_build_camera_major_copies(psba_ctx *c)
{
    c->cam_pt = salloc<int>(c, c->o);
    c->cam_impts = salloc<double>(c, (size_t)c->o * 2);
    if (c->o) k_camera_major_copy<<<cdiv(c->o, 256), 256, 0, c->stream>>>(c->o, c->cam_obs, c->iidx, c->impts, c->cam_pt, c->cam_impts);
}
