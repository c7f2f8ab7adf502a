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


// ---- row-sweep pair pass (kernels_schur.cu: k_schur_rows) -----------------------------------------
// A "visit" is an observation in camera-major order: camera k looks at point i.  The blocks a visit needs are
// the observations of point i with camera <= k: a contiguous prefix of the point's observations (cameras
// ascend inside a point).  Visits of one camera are cut into chunks by a running cost (prefix length + 2 per
// visit): chunk = floor(cost_prefix / ROW_BUDGET), computable per visit without a sequential pass.
__global__ void k_visit_len(int o, const int *__restrict__ cam_obs, const int *__restrict__ iidx, const int *__restrict__ pt_ptr,
                            int *__restrict__ len, int *__restrict__ cam_pos, int *__restrict__ maxlen)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= o) return;
    const int a = cam_obs[v];
    const int l = a - pt_ptr[iidx[a]] + 1;
    len[v] = l; cam_pos[a] = v;
    atomicMax(maxlen, l);
}

__device__ __forceinline__ int row_chunk_of(int v, int row_first, const int *__restrict__ vis_off, int budget)
{
    return ((vis_off[v] + 2 * v) - (vis_off[row_first] + 2 * row_first)) / budget;
}

__global__ void k_row_chunks(int RB, int m, const int *__restrict__ cam_ptr, const int *__restrict__ vis_off, int *__restrict__ nch)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const int b = cam_ptr[k], e = cam_ptr[k + 1];
    nch[k] = e > b ? row_chunk_of(e - 1, b, vis_off, RB) + 1 : 0;
}

__global__ void k_chunk_first(int RB, int o, const int *__restrict__ cam_obs, const int *__restrict__ jidx, const int *__restrict__ cam_ptr,
                              const int *__restrict__ vis_off, const int *__restrict__ row_chunk_base, int n_rchunk, int *__restrict__ chunk_first)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= o) return;
    const int k = jidx[cam_obs[v]], b = cam_ptr[k];
    const int c = row_chunk_of(v, b, vis_off, RB);
    if (v == b || row_chunk_of(v - 1, b, vis_off, RB) != c) chunk_first[row_chunk_base[k] + c] = v;
    if (v == o - 1) chunk_first[n_rchunk] = o;
}

__global__ void k_vis_desc(int RB, int o, const int *__restrict__ cam_obs, const int *__restrict__ iidx, const int *__restrict__ jidx,
                           const int *__restrict__ pt_ptr, const int *__restrict__ cam_ptr, const int *__restrict__ vis_off,
                           const int *__restrict__ row_chunk_base, const int *__restrict__ chunk_first, int4 *__restrict__ desc)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= o) return;
    const int a = cam_obs[v], i = iidx[a], k = jidx[a], b = cam_ptr[k];
    const int vf = chunk_first[row_chunk_base[k] + row_chunk_of(v, b, vis_off, RB)];
    desc[v] = make_int4(pt_ptr[i], vis_off[v + 1] - vis_off[v], vis_off[v] - vis_off[vf], i);
}

// observation behind every staged block, in staging order (visit-major; the blocks of a visit are the observations
// pt_ptr[i] .. own observation of its point)
__global__ void k_block_src(int o, const int4 *__restrict__ desc, const int *__restrict__ vis_off, int *__restrict__ src)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= o) return;
    const int4 d = desc[v];
    const int b0 = vis_off[v];
    for (int j = 0; j < d.y; ++j) src[b0 + j] = d.x + j;
}

// one 16-byte record per chunk: first visit, visits, staged blocks
__global__ void k_chunk_desc(int n_rchunk, const int *__restrict__ chunk_first, const int4 *__restrict__ desc, const int *__restrict__ vis_off, int4 *__restrict__ cdesc)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_rchunk) return;
    const int v0 = chunk_first[c], v1 = chunk_first[c + 1];
    const int4 l = desc[v1 - 1];
    cdesc[c] = make_int4(v0, v1 - v0, l.z + l.y, vis_off[v0]);       // first visit, visits, staged blocks, rank of the first staged block
}

// per triple (pair-sorted): where the row kernel finds its operands
__global__ void k_tri_meta(int packing, int RB, long long ntri, int SC, const int *__restrict__ tri_oa, const int *__restrict__ tri_ob, const int *__restrict__ cam_pos,
                           const int *__restrict__ jidx, const int *__restrict__ cam_ptr, const int *__restrict__ vis_off,
                           const int *__restrict__ row_chunk_base, const int *__restrict__ chunk_first, const int4 *__restrict__ desc,
                           unsigned *__restrict__ meta)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntri) return;
    const int a = tri_oa[t], bo = tri_ob[t];
    const int v = cam_pos[a], k = jidx[a], rb = cam_ptr[k];
    const int cir = row_chunk_of(v, rb, vis_off, RB);
    const int vf = chunk_first[row_chunk_base[k] + cir];
    const int4 d = desc[v];
    if (packing == 0) meta[t] = ((unsigned)(cir % SC) << 18) | ((unsigned)(v - vf) << 10) | (unsigned)(d.z + (bo - d.x));
    else               // flow kernel: chunk in segment (6) | visit in chunk (8) | stage slot of W_il (9) | stage slot of the visit's own block (9)
        meta[t] = ((unsigned)(cir % SC) << 26) | ((unsigned)(v - vf) << 18) | ((unsigned)(d.z + (bo - d.x)) << 9) | (unsigned)(d.z + d.y - 1);
}

// triple range of every (segment, off-diagonal slot): the triples of the pair whose visit lies in the segment
__global__ void k_seg_runs(int RB, int SC, const int *__restrict__ seg_row, const int *__restrict__ seg_slot_base, const int *__restrict__ row_seg_ptr,
                           const int *__restrict__ row_pair0, const long long *__restrict__ tptr, const int *__restrict__ tri_oa,
                           const int *__restrict__ cam_pos, const int *__restrict__ cam_ptr, const int *__restrict__ vis_off,
                           int2 *__restrict__ runs)
{
    const int s = blockIdx.x;
    const int k = seg_row[s], j = s - row_seg_ptr[k], nseg = row_seg_ptr[k + 1] - row_seg_ptr[k];
    const int pair0 = row_pair0[k], noff = row_pair0[k + 1] - pair0 - 1, rb = cam_ptr[k];
    for (int slot = threadIdx.x; slot < noff; slot += blockDim.x) {
        const long long t0 = tptr[pair0 + slot], t1 = tptr[pair0 + slot + 1];
        long long be = t0, en = t1;
        if (nseg > 1) {
            auto lower = [&](int want) {         // first triple of the pair whose segment is >= want
                long long lo = t0, hi = t1;
                while (lo < hi) {
                    const long long mid = (lo + hi) >> 1;
                    if (row_chunk_of(cam_pos[tri_oa[mid]], rb, vis_off, RB) / SC < want) lo = mid + 1; else hi = mid;
                }
                return lo;
            };
            be = lower(j); en = lower(j + 1);
        }
        runs[seg_slot_base[s] + slot] = make_int2((int)be, (int)en);
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


// tables of the row-sweep pair pass; falls back (rows_ok = false) when a prefix or a row exceeds the kernel's caps
static void build_row_sweep(psba_ctx *c, const long long *tptr, const std::vector<int> &cptr, const std::vector<int> &pk)
{
    cudaStream_t st = c->stream;
    const int m = c->m, o = c->o;
    c->rows_ok = false; c->n_rchunk = c->n_rseg = c->n_rpart = 0;
    if (c->pair_mode != 2 && c->pair_mode != 4) return;
    // pairs of a row are contiguous (sorted by k, then l; the diagonal is the last one)
    std::vector<int> row_pair0((size_t)m + 1, 0);
    {
        size_t q = 0;
        for (int k = 0; k <= m; ++k) { while (q < pk.size() && pk[q] < k) ++q; row_pair0[k] = (int)q; }
    }
    int max_off = 0;
    for (int k = 0; k < m; ++k) max_off = std::max(max_off, row_pair0[k + 1] - row_pair0[k] - 1);
    if (max_off > 384 || o == 0) return;
    int RB = 304;
    if (getenv("PSBA_ROW_BUDGET") && atoi(getenv("PSBA_ROW_BUDGET")) == 640 && c->pair_mode == 2) RB = 640;
    int maxlen_cap = ROW_MAXLEN;
    if (c->pair_mode == 4) {                              // flow kernel: (budget, stages, longest prefix) = (304,2,64) (192,3,32) (144,4,32) (112,5,32)
        RB = 192;
        if (getenv("PSBA_FLOW_B")) RB = atoi(getenv("PSBA_FLOW_B"));
        if (RB != 304 && RB != 192 && RB != 144 && RB != 112) RB = 192;
        maxlen_cap = RB == 304 ? 64 : 32;
    }
    c->row_budget = RB;
    int *len = salloc<int>(c, (size_t)o + 1), *cam_pos = salloc<int>(c, o), *vis_off = salloc<int>(c, (size_t)o + 1);
    int *maxlen = salloc<int>(c, 1), *cam_ptr = supload(c, cptr);
    CUDA_CHECK(cudaMemsetAsync(maxlen, 0, sizeof(int), st));
    CUDA_CHECK(cudaMemsetAsync(len + o, 0, sizeof(int), st));
    k_visit_len<<<cdiv(o, 256), 256, 0, st>>>(o, c->cam_obs, c->iidx, c->pt_ptr, len, cam_pos, maxlen);
    {
        size_t tb = 0;
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, len, vis_off, o + 1, st));
        void *tmp = psba_dev_alloc(c, std::max<size_t>(tb, 16), false);
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb, len, vis_off, o + 1, st));
        psba_dev_free(c, tmp);
    }
    int *nch = salloc<int>(c, m);
    k_row_chunks<<<cdiv(m, 256), 256, 0, st>>>(RB, m, cam_ptr, vis_off, nch);
    std::vector<int> hnch(m);
    int hmax = 0;
    CUDA_CHECK(cudaMemcpyAsync(hnch.data(), nch, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(&hmax, maxlen, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    psba_dev_free(c, nch); psba_dev_free(c, maxlen); psba_dev_free(c, len);
    if (hmax > maxlen_cap) { psba_dev_free(c, cam_pos); psba_dev_free(c, vis_off); psba_dev_free(c, cam_ptr); return; }
    std::vector<int> chunk_base((size_t)m + 1, 0);
    for (int k = 0; k < m; ++k) chunk_base[k + 1] = chunk_base[k] + hnch[k];
    c->n_rchunk = chunk_base[m];
    // segments (one CTA each): SC consecutive chunks of one row
    int SC = std::max(1, cdiv(c->n_rchunk, 8 * c->n_sm));
    if (getenv("PSBA_ROW_SEG")) SC = std::max(1, atoi(getenv("PSBA_ROW_SEG")));
    SC = std::min(SC, c->pair_mode == 4 ? 63 : (1 << 14));      // mode 4: 6 bits, 63 is the "no triple" mark
    std::vector<int> seg_row, seg_slot_base, row_seg_ptr((size_t)m + 1, 0);
    std::vector<int2> seg_chunks;
    int nslot_total = 0;
    for (int k = 0; k < m; ++k) {
        for (int b = 0; b < hnch[k]; b += SC) {
            seg_row.push_back(k);
            seg_chunks.push_back(make_int2(chunk_base[k] + b, chunk_base[k] + std::min(b + SC, hnch[k])));
            seg_slot_base.push_back(nslot_total);
            nslot_total += row_pair0[k + 1] - row_pair0[k];
        }
        row_seg_ptr[k + 1] = (int)seg_row.size();
    }
    c->n_rseg = (int)seg_row.size(); c->n_rpart = nslot_total;
    c->rows_nt = max_off <= 168 ? 256 : 544;
    // (NT - 32) / 4 groups of four lanes, three pairs per group; the flow kernel (mode 4) exists in the small size only
    if (c->pair_mode == 4 && max_off > 144) { c->n_rseg = c->n_rpart = 0; psba_dev_free(c, cam_pos); psba_dev_free(c, vis_off); psba_dev_free(c, cam_ptr); return; }
    int *row_chunk_base = supload(c, chunk_base);
    c->rchunk_first = salloc<int>(c, (size_t)c->n_rchunk + 1);
    k_chunk_first<<<cdiv(o, 256), 256, 0, st>>>(RB, o, c->cam_obs, c->jidx, cam_ptr, vis_off, row_chunk_base, c->n_rchunk, c->rchunk_first);
    c->vis_desc = salloc<int4>(c, o);
    k_vis_desc<<<cdiv(o, 256), 256, 0, st>>>(RB, o, c->cam_obs, c->iidx, c->jidx, c->pt_ptr, cam_ptr, vis_off, row_chunk_base, c->rchunk_first, c->vis_desc);
    c->rchunk_desc = salloc<int4>(c, (size_t)c->n_rchunk + 4);
    CUDA_CHECK(cudaMemsetAsync(c->rchunk_desc, 0, ((size_t)c->n_rchunk + 4) * sizeof(int4), st));
    if (c->n_rchunk) k_chunk_desc<<<cdiv(c->n_rchunk, 256), 256, 0, st>>>(c->n_rchunk, c->rchunk_first, c->vis_desc, vis_off, c->rchunk_desc);
    c->rblk_src = nullptr;
    if (c->pair_mode == 4) {
        c->rblk_src = salloc<int>(c, (size_t)c->ntri + 512);
        CUDA_CHECK(cudaMemsetAsync(c->rblk_src, 0, ((size_t)c->ntri + 512) * sizeof(int), st));
        k_block_src<<<cdiv(o, 256), 256, 0, st>>>(o, c->vis_desc, vis_off, c->rblk_src);
    }
    c->tri_meta = salloc<unsigned>(c, (size_t)c->ntri + 4);
    if (c->ntri) k_tri_meta<<<cdiv(c->ntri, 256), 256, 0, st>>>(c->pair_mode == 4 ? 1 : 0, RB, c->ntri, SC, c->tri_oa, c->tri_ob, cam_pos, c->jidx, cam_ptr, vis_off, row_chunk_base,
                                                              c->rchunk_first, c->vis_desc, c->tri_meta);
    c->rseg_row = supload(c, seg_row); c->rseg_chunks = supload(c, seg_chunks); c->rseg_slot_base = supload(c, seg_slot_base);
    c->row_pair0 = supload(c, row_pair0); c->row_seg_ptr = supload(c, row_seg_ptr);
    c->rseg_runs = salloc<int2>(c, (size_t)nslot_total);
    CUDA_CHECK(cudaMemsetAsync(c->rseg_runs, 0, std::max<size_t>(nslot_total, 1) * sizeof(int2), st));
    if (c->n_rseg) k_seg_runs<<<c->n_rseg, 128, 0, st>>>(RB, SC, c->rseg_row, c->rseg_slot_base, c->row_seg_ptr, c->row_pair0, tptr, c->tri_oa, cam_pos,
                                                        cam_ptr, vis_off, c->rseg_runs);
    CUDA_CHECK(cudaStreamSynchronize(st));          // the host vectors above are read by the async uploads
    psba_dev_free(c, cam_pos); psba_dev_free(c, vis_off); psba_dev_free(c, cam_ptr); psba_dev_free(c, row_chunk_base);
    c->rows_ok = true;
}

// ---- the build ---------------------------------------------------------------------------------------
void psba_build_structure(psba_ctx *c, const int *iidx, const int *jidx)
{
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
    std::vector<int> hptr((size_t)ng + 1);
    int bad = 0;
    CUDA_CHECK(cudaMemcpyAsync(&bad, c->d_status + 3, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(hptr.data(), gptr, ((size_t)ng + 1) * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (bad) { fprintf(stderr, "psba_b200: fill_idxBuffer: observations are not point-major with ascending cameras\n"); exit(EXIT_FAILURE); }
    T.lap("index upload + CSR");
    // ---- this rank's contiguous point range, balanced by observation count (psba_local_range)
    auto first_pt = [&](int r) -> int {
        if (r >= c->nranks) return ng;
        const long long target = (long long)og * r / c->nranks;
        return (int)(std::lower_bound(hptr.begin(), hptr.begin() + ng, (int)target) - hptr.begin());
    };
    const int p0 = first_pt(c->rank), p1 = first_pt(c->rank + 1), o0 = hptr[p0], o1 = hptr[p1];
    c->p_off = p0; c->o_off = o0; c->n = p1 - p0; c->o = o1 - o0;
    const int n = c->n, o = c->o;
    c->iidx = salloc<int>(c, o); c->jidx = salloc<int>(c, o); c->pt_ptr = salloc<int>(c, (size_t)n + 1);
    int *iota = salloc<int>(c, o);
    if (o) k_local_idx<<<cdiv(o, 256), 256, 0, st>>>(o, o0, p0, gi, gj, c->iidx, c->jidx, iota);
    k_local_ptr<<<cdiv(n + 1, 256), 256, 0, st>>>(n + 1, p0, o0, gptr, c->pt_ptr);
    // point chunks: whole points, <= PT_CTA observations and <= PT_CTA points per CTA (greedy, host)
    std::vector<int> pch(1, 0);
    {
        int cnt_o = 0, cnt_p = 0;
        for (int i = 0; i < n; ++i) {
            const int d = hptr[p0 + i + 1] - hptr[p0 + i];
            if (cnt_p > 0 && (cnt_o + d > PT_CTA || cnt_p == PT_CTA)) { pch.push_back(i); cnt_o = 0; cnt_p = 0; }
            cnt_o += d; cnt_p++;
        }
        if (n > 0) pch.push_back(n);
    }
    c->n_ptchunk = (int)pch.size() - 1;
    c->ptchunk = supload(c, pch);
    {
        std::vector<int4> desc(c->n_ptchunk);            // {p0, p1, o0, o1}: one 16-byte load per CTA
        for (int q = 0; q < c->n_ptchunk; ++q)
            desc[q] = make_int4(pch[q], pch[q + 1], hptr[p0 + pch[q]] - o0, hptr[p0 + pch[q + 1]] - o0);
        c->ptdesc = supload(c, desc);
        // chunks that fit one wave go through the pipelined kernels; a chunk that is one point with more than
        // PT_CTA observations keeps the wave loop
        std::vector<int> small, big;
        for (int q = 0; q < c->n_ptchunk; ++q) (desc[q].w - desc[q].z <= PT_CTA ? small : big).push_back(q);
        c->n_small = (int)small.size(); c->n_big = (int)big.size();
        c->d_small_list = big.empty() ? nullptr : supload(c, small);
        c->d_big_list = big.empty() ? nullptr : supload(c, big);
    }
    // ---- camera-major order: stable radix sort of the local observations by camera
    c->cam_obs = salloc<int>(c, o);
    std::vector<int> cptr((size_t)m + 1, 0);
    {
        int *skey = salloc<int>(c, o), *cp = salloc<int>(c, (size_t)m + 1);
        CUDA_CHECK(cudaMemsetAsync(cp, 0, ((size_t)m + 1) * 4, st));
        size_t tb = 0;
        const int bits = bits_for((u64)m);
        CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, c->jidx, skey, iota, c->cam_obs, o, 0, bits, st));
        void *tmp = psba_dev_alloc(c, std::max<size_t>(tb, 16), false);
        if (o) {
            CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tb, c->jidx, skey, iota, c->cam_obs, o, 0, bits, st));
            k_segment_ptr<<<cdiv(o, 256), 256, 0, st>>>(o, m, skey, cp);
        }
        CUDA_CHECK(cudaMemcpyAsync(cptr.data(), cp, ((size_t)m + 1) * 4, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        psba_dev_free(c, tmp); psba_dev_free(c, skey); psba_dev_free(c, cp); psba_dev_free(c, iota);
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
    // ---- triple range of every pair, chunks of the pair pass
    long long *tptr = salloc<long long>(c, (size_t)c->n_pair + 1);
    k_pair_ptr<<<cdiv(c->n_pair + 1, 256), 256, 0, st>>>(c->n_pair, m, c->pair_k, c->pair_l, c->ntri, lkeys, tptr);
    int *ccnt = salloc<int>(c, (size_t)c->n_pair + 1), *nonempty = salloc<int>(c, 1);
    CUDA_CHECK(cudaMemsetAsync(nonempty, 0, sizeof(int), st));
    CUDA_CHECK(cudaMemsetAsync(ccnt, 0, ((size_t)c->n_pair + 1) * 4, st));
    k_chunk_count<<<cdiv(c->n_pair, 256), 256, 0, st>>>(c->n_pair, 1, tptr, ccnt, nonempty);   // pch = 1: counts triples; only `nonempty` is used
    int h_nonempty = 0;
    CUDA_CHECK(cudaMemcpyAsync(&h_nonempty, nonempty, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    // lane-group size of the pair pass: every group of G lanes owns one chunk of <= G*PAIR_TPL triples of
    // one camera pair; G follows the mean run length so that a lane streams ~PAIR_TPL triples before the
    // (shuffle) reduction -- 4..8 for the synthetic ring (117 triples / pair), 32 for BAL (~10^3 / pair)
    // pair_mode 0 (default): a lane per triple; 1: the lanes of a group work as quads (four lanes per triple, a 3x3
    // quadrant of the block each), a chunk is <= (G/4)*PAIR_TPQ triples; 2: row sweep; 3: lane per triple with the
    // operands fetched cooperatively through a per-warp stage.  PSBA_PAIR_MODE selects the measured-and-slower
    // variants
    c->pair_mode = 0;
    if (getenv("PSBA_PAIR_MODE")) c->pair_mode = atoi(getenv("PSBA_PAIR_MODE"));
    long long PCH;
    {
        const double avg = h_nonempty ? (double)c->ntri / (double)h_nonempty : 1.0;
        if (c->pair_mode == 1) {
            int G = 4;
            while (G < 32 && avg > (double)(G / 4) * PAIR_TPQ) G *= 2;
            c->pair_G = G;
            PCH = (long long)(G / 4) * PAIR_TPQ;
        } else {
            int G = 1;
            while (G < 32 && avg > (double)G * PAIR_TPL) G *= 2;
            c->pair_G = G;
            PCH = (long long)G * PAIR_TPL;
        }
    }
    k_chunk_count<<<cdiv(c->n_pair, 256), 256, 0, st>>>(c->n_pair, PCH, tptr, ccnt, nullptr);
    c->pair_chunk_ptr = salloc<int>(c, (size_t)c->n_pair + 1);
    {
        size_t tb = 0;
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tb, ccnt, c->pair_chunk_ptr, c->n_pair + 1, st));
        void *tmp = psba_dev_alloc(c, std::max<size_t>(tb, 16), false);
        CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb, ccnt, c->pair_chunk_ptr, c->n_pair + 1, st));
        CUDA_CHECK(cudaMemcpyAsync(&c->n_pchunk, c->pair_chunk_ptr + c->n_pair, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        psba_dev_free(c, tmp);
    }
    c->pchunk_pair = salloc<int>(c, c->n_pchunk);
    c->pchunk_beg = salloc<long long>(c, c->n_pchunk); c->pchunk_end = salloc<long long>(c, c->n_pchunk);
    k_chunk_fill<<<cdiv(c->n_pair, 256), 256, 0, st>>>(c->n_pair, PCH, tptr, c->pair_chunk_ptr, c->pchunk_pair, c->pchunk_beg, c->pchunk_end);
    T.lap("pair set + chunks");
    build_row_sweep(c, tptr, cptr, pk);
    T.lap("row-sweep tables");
    psba_dev_free(c, tptr); psba_dev_free(c, ccnt); psba_dev_free(c, nonempty); psba_dev_free(c, lkeys);
    // ---- camera system tiles (symbolic factorisation, host)
    psba_build_tile_structure(c, pairs);
    T.lap("tile structure");
}

// camera-major copies of the per-observation constants the camera pass reads (point index, measurement)
void psba_build_camera_major_copies(psba_ctx *c)
{
    c->cam_pt = salloc<int>(c, c->o);
    c->cam_impts = salloc<double>(c, (size_t)c->o * 2);
    if (c->o) k_camera_major_copy<<<cdiv(c->o, 256), 256, 0, c->stream>>>(c->o, c->cam_obs, c->iidx, c->impts, c->cam_pt, c->cam_impts);
}
