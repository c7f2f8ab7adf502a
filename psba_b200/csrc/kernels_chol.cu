// kernels_chol.cu -- camera solve: tiled FP64 Cholesky of the reduced camera system S fused with
// the forward substitution, and the backward substitution.
//
// Replaces kern_cholesky / kern_cholesky_s2 / kern_trigMat_inv / kern_trigMat_mul / kern_fill_rest
// (CL_files/SPD_inv.cl:20-411, host PSBA/cl_spdinv.cpp:18-204) and kern_matVec_mul
// (matVec_mul.cl:7-17): dpa = S^-1 ea is obtained from S = L L^T by two triangular solves instead of
// the explicit inverse (parity-safe, LM costs agree to 4e-15).
//
// S lives in a pool of 48x48 tiles (8 camera blocks per tile edge); only the tiles of the symbolic
// factor exist.  The dependent chain of a Cholesky factorisation is its panel sequence, so the cost
// at these sizes (N = 6m: 42..828 for BAL, 12000 block-banded for the synthetic ring) is launch
// latency x panels.  One kernel per panel K, all panels in one CUDA graph:
//   critical CTAs (one per tile row I >= K with a tile (I,K)):
//       apply the DEFERRED update of panel K-1 to tile (I,K) and to the diagonal tile (K,K),
//       factor the diagonal tile (every CTA redundantly: no grid-wide dependency), invert the
//       factor, L_IK = A_IK L_KK^-T, and the forward substitution y_K = L_KK^-1 b_K,
//       b_I -= L_IK y_K  (the right-hand side rides along as an extra column);
//   deferred CTAs: A_IJ -= L_I,K-1 L_J,K-1^T for the trailing tiles J > K of panel K-1.
// Failure (pivot <= 0 or not finite) sets a status word; the reference reports the same event as a
// non-finite factor entry (SPD_inv.cl:66-107).  FP64 on CUDA cores: tcgen05 has no FP64 kind and
// the panels (48 wide) are latency-, not throughput-bound.
#include "dev_math.cuh"
#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <cstring>

#define LDT (TS + 1)            // padded leading dimension in shared memory
#define TILE_SM (TS * LDT)      // doubles per shared tile
#define LDD (TS + 4)            // leading dimension of the tiles staged for the tensor-core product (8 rows x 32 B per fragment load: 2 wavefronts)
#define CHOL_SMEM (2 * TS * LDD * sizeof(double))   // two staged tiles (factor tiles of panel K-1, then D and A_IK)

// ---------------------------------------------------------------------------------------------
// Symbolic phase (host, once per problem).
//
// Ordering: nested dissection of the TILE graph (tile = 8 consecutive cameras) by BFS level
// structures (George's automatic nested dissection): a connected subgraph is rooted at a
// pseudo-peripheral tile, the level that balances the two sides is the separator, the two sides
// are ordered recursively and the separator goes last.  Subgraphs whose level structure has fewer
// than three levels (every dense BAL system) keep their natural order.  For the block-banded ring
// (250 tile columns, half band width 8 tiles) this turns a dependent chain of 250 panels into ~50
// steps of independent panels.
struct nd_ctx {
    const std::vector<std::vector<int>> *adj;
    std::vector<int> mark, dist;
    int tag;
    int min_size;
    // degree of a root candidate counted inside the subgraph (PSBA_ND_ROOT=1) instead of in the whole tile graph (default, what
    // every GPU measurement of rounds 1-2 ran with).  In a band every tile has the same degree in the whole graph, so the default
    // takes the FIRST tile of the last level, not the end of the band: the level structure then starts [1, 15, 8, ...] and a 23-tile
    // segment splits into 1 + 13..15 (separator) + 9.  Counted inside the subgraph the root is the end of the band, every separator
    // of the headline ring is 8 tiles and the schedule has 49 steps instead of 56 (same 4 506 factor tiles; CPU-checked plan,
    // tests/test_tile_plan_cpu.py) -- found after the GPU budget of round 2 was spent, hence not yet the default.
    bool subset_degree = false;
    std::vector<int> order;
    std::vector<int> front;     // front[p]: id of the leaf / separator the tile at position p belongs to (consecutive positions)
    int n_front = 0;
    void append(const std::vector<int> &nodes) { order.insert(order.end(), nodes.begin(), nodes.end()); front.insert(front.end(), nodes.size(), n_front++); }
};

static void nd_bfs(nd_ctx &x, int root, int tag, std::vector<int> &visit, int &nlev)
{
    // BFS inside the subset {v : mark[v] == tag}; dist doubles as the visited flag (-1 = unseen)
    visit.clear();
    visit.push_back(root);
    x.dist[root] = 0;
    for (size_t h = 0; h < visit.size(); ++h) {
        const int v = visit[h];
        for (int w : (*x.adj)[v])
            if (x.mark[w] == tag && x.dist[w] < 0) { x.dist[w] = x.dist[v] + 1; visit.push_back(w); }
    }
    nlev = x.dist[visit.back()] + 1;
}

static void nd_recurse(nd_ctx &x, std::vector<int> nodes)
{
    std::sort(nodes.begin(), nodes.end());
    if ((int)nodes.size() <= x.min_size) { x.append(nodes); return; }
    const int tag = ++x.tag;
    for (int v : nodes) { x.mark[v] = tag; x.dist[v] = -1; }
    std::vector<int> visit;
    int nlev = 0;
    nd_bfs(x, nodes[0], tag, visit, nlev);
    if (visit.size() < nodes.size()) {
        // disconnected: every component is an independent subtree of the elimination forest
        std::vector<std::vector<int>> comps;
        comps.push_back(visit);
        for (int v : nodes)
            if (x.dist[v] < 0) { nd_bfs(x, v, tag, visit, nlev); comps.push_back(visit); }
        for (auto &cmp : comps) nd_recurse(x, cmp);
        return;
    }
    // pseudo-peripheral root: restart from the lowest-degree tile of the last level while the depth grows
    int root = nodes[0];
    for (int it = 0; it < 4; ++it) {
        int best = -1;
        size_t bdeg = (size_t)-1;
        for (int v : visit) {
            if (x.dist[v] != nlev - 1) continue;
            size_t deg = (*x.adj)[v].size();
            if (x.subset_degree) { deg = 0; for (int w : (*x.adj)[v]) deg += x.mark[w] == tag; }
            if (deg < bdeg) { bdeg = deg; best = v; }
        }
        for (int v : nodes) x.dist[v] = -1;
        int nl2 = 0;
        nd_bfs(x, best, tag, visit, nl2);
        const bool grew = nl2 > nlev;
        root = best; nlev = nl2;
        if (!grew) break;
    }
    (void)root;
    if (nlev < 3) { x.append(nodes); return; }
    std::vector<int> cnt(nlev, 0);
    for (int v : nodes) cnt[x.dist[v]]++;
    int best_l = 1;
    long long best_score = -1;
    {
        long long below = cnt[0];
        for (int l = 1; l <= nlev - 2; ++l) {
            const long long above = (long long)nodes.size() - below - cnt[l];
            const long long score = std::max(below, above) + cnt[l];      // longest remaining chain
            if (best_score < 0 || score < best_score) { best_score = score; best_l = l; }
            below += cnt[l];
        }
    }
    std::vector<int> A, B, S;
    for (int v : nodes) (x.dist[v] < best_l ? A : x.dist[v] > best_l ? B : S).push_back(v);
    nd_recurse(x, A);
    nd_recurse(x, B);
    std::sort(S.begin(), S.end());
    x.append(S);
}

// The symbolic part below runs on a helper thread under device work of the set-up (structure.cu); it issues no CUDA call:
// every upload / zero-filled allocation is recorded here and carried out by psba_flush_tile_uploads on the calling thread
struct pending_upload { void **dst; std::vector<char> bytes; size_t zero_bytes; };
static std::map<psba_ctx *, std::vector<pending_upload>> g_pending;   // one set-up at a time per context
static std::mutex g_pending_mu;
static std::vector<pending_upload> &pending(psba_ctx *c)
{
    std::lock_guard<std::mutex> l(g_pending_mu);
    return g_pending[c];
}
template <class T> static void up_vec(psba_ctx *c, T **d, const std::vector<T> &h)
{
    pending_upload u;
    u.dst = (void **)d; u.zero_bytes = 0;
    u.bytes.assign((const char *)h.data(), (const char *)h.data() + h.size() * sizeof(T));
    pending(c).push_back(std::move(u));
}
template <class T> static void zero_alloc(psba_ctx *c, T **d, size_t bytes)
{
    pending_upload u;
    u.dst = (void **)d; u.zero_bytes = std::max<size_t>(bytes, 16);
    pending(c).push_back(std::move(u));
}
void psba_flush_tile_uploads(psba_ctx *c)
{
    // ONE device block and ONE host-to-device copy for all tables (forty separate pageable copies cost 1.3 ms at 2 000
    // cameras); the zero-filled work arrays (tile pool, ...) stay separate allocations
    std::vector<pending_upload> &q = pending(c);
    size_t total = 0;
    auto room = [](const pending_upload &u) { return std::max<size_t>(256, (u.bytes.size() + 255) & ~(size_t)255); };   // distinct addresses for empty tables
    for (pending_upload &u : q) if (!u.zero_bytes) total += room(u);
    // the block is assembled in page-locked memory kept by the process (one set-up at a time; pageable memory otherwise)
    static std::mutex pin_mu;
    static char *pin_buf = nullptr; static size_t pin_cap = 0; static bool pin_busy = false;
    const size_t need = std::max<size_t>(total, 256);
    std::vector<char> pageable;
    char *stage = nullptr;
    {
        std::lock_guard<std::mutex> l(pin_mu);
        if (!pin_busy) {
            if (pin_cap < need) {
                if (pin_buf) cudaFreeHost(pin_buf);
                pin_buf = nullptr; pin_cap = 0;
                if (cudaHostAlloc((void **)&pin_buf, need + need / 2, cudaHostAllocDefault) == cudaSuccess) pin_cap = need + need / 2;
                else { cudaGetLastError(); pin_buf = nullptr; }
            }
            if (pin_buf) { stage = pin_buf; pin_busy = true; }
        }
    }
    const bool pinned = stage != nullptr;
    if (!pinned) { pageable.resize(need); stage = pageable.data(); }
    c->tile_block = psba_dev_alloc(c, need, false);
    c->tile_block_bytes = need;
    size_t off = 0;
    for (pending_upload &u : q) {
        if (u.zero_bytes) { *u.dst = psba_dev_alloc(c, u.zero_bytes, true); continue; }
        if (!u.bytes.empty()) memcpy(stage + off, u.bytes.data(), u.bytes.size());
        *u.dst = (char *)c->tile_block + off;
        off += room(u);
    }
    CUDA_CHECK(cudaMemcpyAsync(c->tile_block, stage, std::min(need, std::max<size_t>(off, 1)), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));            // the staged bytes are free again
    if (pinned) { std::lock_guard<std::mutex> l(pin_mu); pin_busy = false; }
    { std::lock_guard<std::mutex> l(g_pending_mu); g_pending.erase(c); }   // `q` is gone from here on
}

void psba_build_tile_structure(psba_ctx *c, const std::vector<std::pair<int, int>> &pairs)
{
    const int nt = (c->m + 7) / 8;
    c->nt = nt;
    const bool lapon = getenv("PSBA_SETUP_TIMING") != nullptr;
    auto lap_t = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!lapon) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "psba setup:   tiles: %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - lap_t).count());
        lap_t = now;
    };
    // ---- tile graph in natural numbering
    std::vector<char> nat((size_t)nt * nt, 0);
    for (auto &p : pairs) { const int a = p.first / 8, b = p.second / 8; nat[(size_t)a * nt + b] = 1; nat[(size_t)b * nt + a] = 1; }
    std::vector<std::vector<int>> adj(nt);
    for (int a = 0; a < nt; ++a)
        for (int b = 0; b < nt; ++b)
            if (a != b && nat[(size_t)a * nt + b]) adj[a].push_back(b);
    // ---- ordering
    std::vector<int> tpos(nt);                       // natural tile -> position
    std::vector<int> front(nt, 0);                   // position -> front (leaf or separator of the dissection)
    {
        const char *e = getenv("PSBA_ND_MIN");
        int min_size = e ? atoi(e) : 12;
        nd_ctx x;
        x.adj = &adj; x.mark.assign(nt, 0); x.dist.assign(nt, -1); x.tag = 0; x.min_size = std::max(1, min_size);
        x.subset_degree = getenv("PSBA_ND_ROOT") && atoi(getenv("PSBA_ND_ROOT")) != 0;
        std::vector<int> all(nt);
        for (int t = 0; t < nt; ++t) all[t] = t;
        if (min_size <= 0 || min_size >= nt) x.append(all);      // natural order
        else nd_recurse(x, all);
        if ((int)x.order.size() != nt) { fprintf(stderr, "psba_b200: internal error in the tile ordering\n"); exit(EXIT_FAILURE); }
        for (int p = 0; p < nt; ++p) tpos[x.order[p]] = p;
        front = x.front;
    }
    c->h_cam2pos.resize(c->m);
    std::vector<int> pos2cam((size_t)nt * 8, -1);
    for (int j = 0; j < c->m; ++j) { c->h_cam2pos[j] = tpos[j / 8] * 8 + j % 8; pos2cam[c->h_cam2pos[j]] = j; }
    up_vec(c, &c->cam2pos, c->h_cam2pos);
    up_vec(c, &c->pos2cam, pos2cam);
    lap("tile graph + ordering");
    // ---- pattern in permuted numbering + symbolic factorisation at tile granularity
    std::vector<char> present((size_t)nt * nt, 0);
    for (int I = 0; I < nt; ++I) present[(size_t)I * nt + I] = 1;
    for (int a = 0; a < nt; ++a)
        for (int b : adj[a]) { const int I = std::max(tpos[a], tpos[b]), J = std::min(tpos[a], tpos[b]); present[(size_t)I * nt + J] = 1; }
    const std::vector<char> presentS = present;      // tiles of S itself (before fill-in)
    std::vector<std::vector<int>> rows(nt);          // rows[K]: tile rows I > K of the factor column K
    for (int K = 0; K < nt; ++K) {
        for (int I = K + 1; I < nt; ++I) if (present[(size_t)I * nt + K]) rows[K].push_back(I);
        for (size_t a = 0; a < rows[K].size(); ++a)
            for (size_t b = 0; b <= a; ++b) present[(size_t)rows[K][a] * nt + rows[K][b]] = 1;
    }
    c->h_tile_index.assign((size_t)nt * nt, -1);
    // slots: the tiles of S first (the only ones the multi-GPU all-reduce has to move), then the fill-in
    int slot = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int I = 0; I < nt; ++I)
            for (int J = 0; J <= I; ++J)
                if (present[(size_t)I * nt + J] && (presentS[(size_t)I * nt + J] != 0) == (pass == 0)) c->h_tile_index[(size_t)I * nt + J] = slot++;
        if (pass == 0) { c->n_tiles_S = slot; slot += (c->N + TS * TS - 1) / (TS * TS); }   // room for ea behind the S tiles: one all-reduce for both
    }
    c->n_tiles = slot;
    lap("symbolic factor");
    // ---- steps: a panel runs one step after the last panel it depends on
    std::vector<int> step(nt, 0);
    std::vector<std::vector<int>> cols(nt);          // cols[K]: panels P < K with a factor tile (K,P), ascending
    for (int P = 0; P < nt; ++P) for (int I : rows[P]) cols[I].push_back(P);
    int n_steps = 0;
    for (int K = 0; K < nt; ++K) {
        int s = 0;
        for (int P : cols[K]) s = std::max(s, step[P] + 1);
        step[K] = s;
        n_steps = std::max(n_steps, s + 1);
    }
    c->n_steps = n_steps;
    // algorithmic flops of the factorisation from the symbolic factor (bench.py: roofline_fp64): per panel with R tile rows
    // below the diagonal  potrf TS^3/3 + trsm R TS^3 + trailing updates R (R + 1) / 2 * 2 TS^3; forward and backward solve 4 TS^2 per tile
    {
        double fl = 0.0;
        const double t3 = (double)TS * TS * TS;
        for (int K = 0; K < nt; ++K) { const double R = (double)rows[K].size(); fl += t3 / 3.0 + R * t3 + R * (R + 1.0) * t3 + 4.0 * (R + 1.0) * TS * TS; }
        c->chol_flops = fl;
    }
    std::vector<std::vector<int>> by_step(n_steps);
    for (int K = 0; K < nt; ++K) by_step[step[K]].push_back(K);
    c->chain_schedule = true;
    for (auto &v : by_step) if (v.size() != 1) c->chain_schedule = false;
    // ---- task lists
    std::vector<int> critI, critK, psrc_ptr(1, 0), psrc;
    std::vector<int> defI, defJ, def_sptr(1, 0), def_src, step_panels;
    std::vector<int> bJ, b_sptr(1, 0), b_slot;
    std::vector<std::vector<int>> psrc_of(nt);
    for (int K = 0; K < nt; ++K) {
        for (int P : cols[K])
            if (step[P] == step[K] - 1) psrc_of[K].push_back(P);
        psrc.insert(psrc.end(), psrc_of[K].begin(), psrc_of[K].end());
        psrc_ptr.push_back((int)psrc.size());
    }
    c->step_crit_ptr.assign(1, 0); c->step_def_ptr.assign(1, 0); c->step_panel_ptr.assign(1, 0); c->step_b_ptr.assign(1, 0);
    // ---- plan of the deferred trailing updates.  Panel P updates every tile (I,J), I >= J in rows[P].  When J runs in the step
    // right after P the critical CTAs of panel J apply the update themselves; every other update is deferred to a step e with
    // step[P] < e < step[J].  Updates of one target that fall into the same step form ONE task with several sources (one
    // read-modify-write of the target, the source tiles streaming under the tensor-core products).  Measured on the headline
    // system (profiles/notes_r02_experiments.md): right after the source panel 1.083 ms; a whole front per task 1.32 ms (the thin
    // upper steps wait for their 8-source tasks); as late as possible with 3 sources per target and step 1.030 ms (default).
    // PSBA_DEF_MERGE=0 restores the step-by-step schedule, PSBA_DEF_CAP sets the sources per target and step.
    lap("steps + sources");
    struct def_item { int e; int key; int P; };
    std::vector<def_item> def_plan;                  // ordered by step, target, source
    {
        const int mode = getenv("PSBA_DEF_MERGE") ? atoi(getenv("PSBA_DEF_MERGE")) : 2;
        const int cap = std::max(1, getenv("PSBA_DEF_CAP") ? atoi(getenv("PSBA_DEF_CAP")) : 3);
        // sources of every target, grouped by target without a sort: count, prefix sum, fill (P ascending)
        std::vector<int> kcnt((size_t)nt * nt + 1, 0);
        for (int P = 0; P < nt; ++P)
            for (size_t a = 0; a < rows[P].size(); ++a)
                for (size_t b = 0; b <= a; ++b) {
                    const int I = rows[P][a], J = rows[P][b];
                    if (step[J] != step[P] + 1) kcnt[(size_t)I * nt + J + 1]++;      // else: handled by the critical CTAs of panel J
                }
        for (size_t q = 0; q < (size_t)nt * nt; ++q) kcnt[q + 1] += kcnt[q];
        std::vector<int> ksrc(kcnt.back()), kfill(kcnt.begin(), kcnt.end() - 1);
        for (int P = 0; P < nt; ++P)
            for (size_t a = 0; a < rows[P].size(); ++a)
                for (size_t b = 0; b <= a; ++b) {
                    const int I = rows[P][a], J = rows[P][b];
                    if (step[J] != step[P] + 1) ksrc[kfill[(size_t)I * nt + J]++] = P;
                }
        // step of every (target, source): mode 0 right after the source panel; mode 2 (default) as late as possible, `cap`
        // sources per target and step: the m deferred sources of a target (ascending step) run in the last ceil(m / cap)
        // steps before its panel, never before their own panel is done
        std::vector<int> kstep(ksrc.size());
        std::vector<int> per_step(n_steps + 1, 0);
        for (int key = 0; key < nt * nt; ++key) {
            const int q0 = kcnt[key], q1 = kcnt[key + 1], mm = q1 - q0;
            if (mm == 0) continue;
            std::sort(ksrc.begin() + q0, ksrc.begin() + q1, [&](int x, int y) { return step[x] != step[y] ? step[x] < step[y] : x < y; });
            const int D = step[key % nt] - 1;
            for (int q = q0; q < q1; ++q) {
                const int P = ksrc[q];
                kstep[q] = mode == 2 ? std::max(step[P] + 1, D - (mm - 1 - (q - q0)) / cap) : step[P] + 1;
                per_step[kstep[q] + 1]++;
            }
        }
        // bucket by step (stable: targets ascending, sources in the order above)
        for (int e = 0; e < n_steps; ++e) per_step[e + 1] += per_step[e];
        def_plan.resize(ksrc.size());
        for (int key = 0; key < nt * nt; ++key)
            for (int q = kcnt[key]; q < kcnt[key + 1]; ++q) def_plan[per_step[kstep[q]]++] = {kstep[q], key, ksrc[q]};
    }
    lap("deferred plan");
    size_t dq = 0;
    for (int s = 0; s < n_steps; ++s) {
        for (int K : by_step[s]) {
            critI.push_back(K); critK.push_back(K);
            for (int I : rows[K]) { critI.push_back(I); critK.push_back(K); }
            step_panels.push_back(K);
        }
        c->step_crit_ptr.push_back((int)critI.size());
        c->step_panel_ptr.push_back((int)step_panels.size());
        // deferred updates that run in step s (see def_plan above): one task per target tile, sources in ascending panel order
        while (dq < def_plan.size() && def_plan[dq].e == s) {
            const int key = def_plan[dq].key;
            defI.push_back((int)(key / nt)); defJ.push_back((int)(key % nt));
            while (dq < def_plan.size() && def_plan[dq].e == s && def_plan[dq].key == key) def_src.push_back(def_plan[dq++].P);
            def_sptr.push_back((int)def_src.size());
        }
        c->step_def_ptr.push_back((int)defI.size());
        // right-hand-side tasks: b_J -= sum_P L_JP y_P for the panels P of step s-1 and rows J of later steps
        if (s > 0) {
            std::vector<int> tg;
            std::vector<std::vector<int>> sl;
            std::vector<int> at(nt, -1);
            for (int P : by_step[s - 1])
                for (int J : rows[P]) {
                    if (step[J] == s) continue;                  // summed by the critical CTAs of panel J
                    if (at[J] < 0) { at[J] = (int)tg.size(); tg.push_back(J); sl.emplace_back(); }
                    sl[at[J]].push_back(c->h_tile_index[(size_t)J * nt + P]);
                }
            for (size_t t = 0; t < tg.size(); ++t) {
                bJ.push_back(tg[t]);
                b_slot.insert(b_slot.end(), sl[t].begin(), sl[t].end());
                b_sptr.push_back((int)b_slot.size());
            }
        }
        c->step_b_ptr.push_back((int)bJ.size());
    }
    lap("task lists");
    std::vector<int> cptr(1, 0), crow, cslot;
    for (int J = 0; J < nt; ++J) {
        for (int I : rows[J]) { crow.push_back(I); cslot.push_back(c->h_tile_index[(size_t)I * nt + J]); }
        cptr.push_back((int)crow.size());
    }
    up_vec(c, &c->tile_index, c->h_tile_index);
    up_vec(c, &c->d_crit_I, critI); up_vec(c, &c->d_crit_K, critK);
    {
        // flat task descriptors: everything a CTA needs to find its tiles in ONE dependent load
        //   critical: {I, K, slot(I,K), slot(K,K)} {src_begin, src_end}  + per source {slot(K,P), slot(I,P) or -1}
        //   deferred: {slot(I,J), src_begin, src_end, -}                + per source {slot(I,P), slot(J,P)}
        auto slot_of = [&](int A, int B) { return c->h_tile_index[(size_t)A * nt + B]; };
        std::vector<int4> cd; std::vector<int2> cs;
        for (size_t t = 0; t < critI.size(); ++t) {
            const int I = critI[t], K = critK[t];
            const int b0 = (int)cs.size();
            for (int P : psrc_of[K]) cs.push_back(make_int2(slot_of(K, P), I == K ? -1 : slot_of(I, P)));
            cd.push_back(make_int4(I, K, slot_of(I, K), slot_of(K, K)));
            cd.push_back(make_int4(b0, (int)cs.size(), 0, 0));
        }
        std::vector<int4> dd; std::vector<int2> dsrc;
        for (size_t t = 0; t < defI.size(); ++t) {
            const int I = defI[t], J = defJ[t];
            const int b0 = (int)dsrc.size();
            for (int q = def_sptr[t]; q < def_sptr[t + 1]; ++q) dsrc.push_back(make_int2(slot_of(I, def_src[q]), slot_of(J, def_src[q])));
            dd.push_back(make_int4(slot_of(I, J), b0, (int)dsrc.size(), 0));
        }
        up_vec(c, &c->d_crit_desc, cd); up_vec(c, &c->d_crit_src, cs);
        up_vec(c, &c->d_def_desc, dd); up_vec(c, &c->d_def_srcs, dsrc);
    }
    if (c->chol_flow) {
        // ---- dataflow schedule (k_panel_flow): ONE launch, tasks in step order; a task waits on per-tile write counters
        //   tver[slot] = completed writes of the tile (its deferred updates in step order, then the factor tile itself)
        //   bver[J]    = completed right-hand-side updates of panel J
        // instead of on a kernel boundary.  Everything a task waits for belongs to an earlier step = a smaller block index.
        std::vector<int> ndef(c->n_tiles, 0), nrhs(nt, 0), task_type, task_idx, b_seq(bJ.size(), 0);
        for (size_t t = 0; t < defI.size(); ++t) { const int sl = c->h_tile_index[(size_t)defI[t] * nt + defJ[t]]; ndef[sl]++; }
        std::vector<int> tfinal(c->n_tiles + nt, 0);
        for (int sl = 0; sl < c->n_tiles; ++sl) tfinal[sl] = ndef[sl] + 1;
        std::vector<int> seen_def(c->n_tiles, 0), seen_b(nt, 0), def_seq(defI.size(), 0), crit_need(critI.size() * 2, 0);
        for (int s2 = 0; s2 < n_steps; ++s2) {
            for (int t = c->step_crit_ptr[s2]; t < c->step_crit_ptr[s2 + 1]; ++t) { task_type.push_back(0); task_idx.push_back(t); }
            for (int t = c->step_def_ptr[s2]; t < c->step_def_ptr[s2 + 1]; ++t) {
                const int sl = c->h_tile_index[(size_t)defI[t] * nt + defJ[t]];
                def_seq[t] = seen_def[sl]++;
                task_type.push_back(1); task_idx.push_back(t);
            }
            for (int t = c->step_b_ptr[s2]; t < c->step_b_ptr[s2 + 1]; ++t) { b_seq[t] = seen_b[bJ[t]]++; task_type.push_back(2); task_idx.push_back(t); }
        }
        for (int J = 0; J < nt; ++J) { nrhs[J] = seen_b[J]; tfinal[c->n_tiles + J] = nrhs[J]; }
        for (size_t t = 0; t < critI.size(); ++t) {
            crit_need[2 * t] = ndef[c->h_tile_index[(size_t)critI[t] * nt + critK[t]]];
            crit_need[2 * t + 1] = ndef[c->h_tile_index[(size_t)critK[t] * nt + critK[t]]];
        }
        std::vector<int2> tasks(task_type.size());
        for (size_t q = 0; q < tasks.size(); ++q) tasks[q] = make_int2(task_type[q], task_idx[q]);
        c->n_flow_tasks = (int)tasks.size();
        up_vec(c, &c->d_flow_tasks, tasks); up_vec(c, &c->d_flow_final, tfinal); up_vec(c, &c->d_flow_defseq, def_seq);
        up_vec(c, &c->d_flow_bseq, b_seq); up_vec(c, &c->d_flow_critneed, crit_need);
        zero_alloc(c, &c->d_flow_ver, (size_t)(c->n_tiles + nt) * sizeof(int));
    }
    up_vec(c, &c->d_psrc_ptr, psrc_ptr); up_vec(c, &c->d_psrc, psrc);
    up_vec(c, &c->d_b_J, bJ); up_vec(c, &c->d_b_sptr, b_sptr); up_vec(c, &c->d_b_slot, b_slot);
    up_vec(c, &c->d_def_I, defI); up_vec(c, &c->d_def_J, defJ); up_vec(c, &c->d_def_sptr, def_sptr); up_vec(c, &c->d_def_src, def_src);
    up_vec(c, &c->d_step_panels, step_panels);
    {
        std::vector<int> order(step_panels.rbegin(), step_panels.rend());        // last step first
        up_vec(c, &c->d_bw_order, order);
        zero_alloc(c, &c->bw_xbuf, (size_t)nt * TS * sizeof(double));
    }
    up_vec(c, &c->d_coltile_ptr, cptr); up_vec(c, &c->d_coltile_row, crow); up_vec(c, &c->d_coltile_slot, cslot);
    zero_alloc(c, &c->Stiles, (size_t)c->n_tiles * TS * TS * sizeof(double));
    zero_alloc(c, &c->contrib, (size_t)c->n_tiles * TS * sizeof(double));
    zero_alloc(c, &c->Linv, (size_t)nt * TS * TS * sizeof(double));
    zero_alloc(c, &c->Ldiag, (size_t)nt * TS * TS * sizeof(double));
    lap("descriptors");
    c->chol_graph_ok = false; c->bw_graph_ok = false;
    if (getenv("PSBA_SETUP_TIMING"))
        fprintf(stderr, "psba setup: camera system %d tiles/edge, %d factor tiles, %d steps, %zu critical + %zu deferred + %zu rhs tasks\n",
                nt, c->n_tiles, n_steps, critI.size(), defI.size(), bJ.size());
}

// ---------------------------------------------------------------------------------------------
// a 48x48 tile travels global -> registers (all loads in flight at once) -> padded shared memory
// ---- host-only readback of the plan (include/psba_b200.h: psba_plan_open / _get / _close; tests/test_tile_plan_cpu.py) --------
// psba_build_tile_structure issues no CUDA call and records its uploads in g_pending: run it on a throwaway context and keep the
// tables as host arrays.
struct psba_plan { std::map<std::string, std::vector<int>> arr; };
extern "C" void *psba_plan_open(int nCams, long long npairs, const int *pair_k, const int *pair_l)
{
    if (nCams <= 0 || npairs < 0 || (npairs > 0 && (!pair_k || !pair_l))) return nullptr;
    std::vector<std::pair<int, int>> pairs;
    pairs.reserve((size_t)npairs);
    for (long long q = 0; q < npairs; ++q) {
        if (pair_k[q] < 0 || pair_k[q] >= nCams || pair_l[q] < 0 || pair_l[q] >= nCams) return nullptr;
        pairs.push_back({pair_k[q], pair_l[q]});
    }
    psba_ctx *c = new psba_ctx();                      // value-initialised: no stream, no device memory
    c->m = nCams; c->N = 6 * nCams;
    c->chol_flow = getenv("PSBA_CHOL_FLOW") && atoi(getenv("PSBA_CHOL_FLOW")) != 0;
    psba_build_tile_structure(c, pairs);
    psba_plan *pl = new psba_plan();
    pl->arr["stats"] = {c->nt, c->n_steps, c->n_tiles_S, c->n_tiles, c->chain_schedule ? 1 : 0};
    pl->arr["cam2pos"] = c->h_cam2pos;
    pl->arr["tile_index"] = c->h_tile_index;
    pl->arr["step_crit_ptr"] = c->step_crit_ptr; pl->arr["step_def_ptr"] = c->step_def_ptr;
    pl->arr["step_b_ptr"] = c->step_b_ptr; pl->arr["step_panel_ptr"] = c->step_panel_ptr;
    const std::pair<const char *, void **> named[] = {
        {"crit_I", (void **)&c->d_crit_I}, {"crit_K", (void **)&c->d_crit_K}, {"psrc_ptr", (void **)&c->d_psrc_ptr}, {"psrc", (void **)&c->d_psrc},
        {"def_I", (void **)&c->d_def_I}, {"def_J", (void **)&c->d_def_J}, {"def_sptr", (void **)&c->d_def_sptr}, {"def_src", (void **)&c->d_def_src},
        {"b_J", (void **)&c->d_b_J}, {"b_sptr", (void **)&c->d_b_sptr}, {"b_slot", (void **)&c->d_b_slot}, {"step_panels", (void **)&c->d_step_panels},
        // the flat descriptors the step kernels read (int4 / int2 records as consecutive ints)
        {"bw_order", (void **)&c->d_bw_order}, {"coltile_ptr", (void **)&c->d_coltile_ptr}, {"coltile_row", (void **)&c->d_coltile_row},
        {"coltile_slot", (void **)&c->d_coltile_slot}, {"crit_desc", (void **)&c->d_crit_desc}, {"crit_src", (void **)&c->d_crit_src}, {"def_desc", (void **)&c->d_def_desc}, {"def_srcs", (void **)&c->d_def_srcs}};
    {
        std::lock_guard<std::mutex> l(g_pending_mu);
        for (const pending_upload &u : g_pending[c])
            for (const auto &nm : named)
                if (u.dst == nm.second) {
                    const int *b = (const int *)u.bytes.data();
                    pl->arr[nm.first].assign(b, b + u.bytes.size() / sizeof(int));
                }
        g_pending.erase(c);
    }
    delete c;
    return pl;
}
extern "C" long long psba_plan_get(void *plan, const char *name, int *out, long long max_count)
{
    if (!plan || !name) return -1;
    const psba_plan *pl = (const psba_plan *)plan;
    const auto it = pl->arr.find(name);
    if (it == pl->arr.end()) return -1;
    const long long len = (long long)it->second.size();
    if (out && max_count > 0 && len > 0) memcpy(out, it->second.data(), (size_t)std::min(len, max_count) * sizeof(int));
    return len;
}
extern "C" void psba_plan_close(void *plan) { delete (psba_plan *)plan; }

template <int NT> struct TileRegs { double2 v[(TS * TS / 2 + NT - 1) / NT]; };
// CG: the tile may have been written by another CTA of the SAME launch (dataflow factorisation): read it through L2
template <int NT, bool CG = false>
__device__ __forceinline__ void tile_ldg(TileRegs<NT> &t, const double *g)
{
    constexpr int Q = (TS * TS / 2 + NT - 1) / NT;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int e = threadIdx.x + q * NT;
        t.v[q] = e < TS * TS / 2 ? (CG ? __ldcg(reinterpret_cast<const double2 *>(g) + e) : reinterpret_cast<const double2 *>(g)[e]) : make_double2(0.0, 0.0);
    }
}
template <bool CG> __device__ __forceinline__ double2 ld2c(const double *p) { return CG ? __ldcg(reinterpret_cast<const double2 *>(p)) : *reinterpret_cast<const double2 *>(p); }
template <bool CG> __device__ __forceinline__ double ld1c(const double *p) { return CG ? __ldcg(p) : *p; }

// ---- flags of the dataflow kernels: release / acquire at GPU scope
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const double *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_f64(double *p, double v)
{
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
#define X_SENTINEL 0xFFFFFFFFFFFFFFFFull      // "not there yet" in the solution buffer of the dataflow backward solve (a NaN no solve produces)
__device__ __forceinline__ void red_release_add(int *p, int v)
{
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// one thread waits until counter idx has reached `need`; gives up when the factorisation has failed elsewhere or the
// bounded spin runs out (status 3: a broken schedule becomes an error, not a hang)
__device__ __forceinline__ bool flow_wait(const int *ver, int idx, int need, int *status)
{
    int spins = 0;
    while (ld_acquire(ver + idx) < need) {
        if ((++spins & 255) == 0) {
            if (ld_relaxed(status) != 0) return false;
            if (spins > (1 << 23)) { atomicMax(status, 3); return false; }
        }
    }
    return true;
}
template <int NT>
__device__ __forceinline__ void tile_sts(double *sm, const TileRegs<NT> &t)
{
    constexpr int Q = (TS * TS / 2 + NT - 1) / NT;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int e = threadIdx.x + q * NT;
        if (e < TS * TS / 2) { const int r = (2 * e) / TS, cc = (2 * e) % TS; sm[r * LDT + cc] = t.v[q].x; sm[r * LDT + cc + 1] = t.v[q].y; }
    }
}
// lower_only: entries above the diagonal are written as zero (diagonal factor tiles)
template <int NT>
__device__ __forceinline__ void tile_stg(double *__restrict__ g, const double *sm, bool lower_only = false)
{
    constexpr int Q = (TS * TS / 2 + NT - 1) / NT;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int e = threadIdx.x + q * NT;
        if (e < TS * TS / 2) {
            const int r = (2 * e) / TS, cc = (2 * e) % TS;
            double x = sm[r * LDT + cc], y = sm[r * LDT + cc + 1];
            if (lower_only) { if (cc > r) x = 0.0; if (cc + 1 > r) y = 0.0; }
            reinterpret_cast<double2 *>(g)[e] = make_double2(x, y);
        }
    }
}

// ---- FP64 tensor-core product for the deferred updates ------------------------------------------------
// A 6x3 register-blocked product (rounds 1a-1c) reads 9 doubles from shared memory per 18 FMA and thread; ncu showed the leaf
// steps at 73 % of the L1 data pipe and 25 % of the FP64 pipe.  DMMA.8x8x4 (mma.sync m8n8k4 f64) takes one double
// of A and one of B per lane for 256 FMA: a warp that owns a 24x24 block of the tile (3x3 fragments) loads 6
// doubles per lane and k-step of 4 for 9 DMMA = 2 304 FMA.
template <int NT>
__device__ __forceinline__ void tile_sts_ldd(double *sm, const TileRegs<NT> &t)
{
    constexpr int Q = (TS * TS / 2 + NT - 1) / NT;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int e = threadIdx.x + q * NT;
        if (e < TS * TS / 2) { const int r = (2 * e) / TS, cc = (2 * e) % TS; *reinterpret_cast<double2 *>(sm + r * LDD + cc) = t.v[q]; }
    }
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// c (3x3 fragments of the warp's 24x24 block, rows rb.., columns cb..) -= A[rb..][:] * B[cb..][:]^T, tiles staged with LDD
__device__ __forceinline__ void tile_abt_dmma_sub(const double *A, const double *B, int rb, int cb, double c[3][3][2])
{
    const int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
    const double *ap = A + (rb + fr) * LDD + fk, *bp = B + (cb + fr) * LDD + fk;
#pragma unroll 4
    for (int ks = 0; ks < TS / 4; ++ks) {
        double a[3], b[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) { a[t] = -ap[t * 8 * LDD + ks * 4]; b[t] = bp[t * 8 * LDD + ks * 4]; }
#pragma unroll
        for (int ti = 0; ti < 3; ++ti)
#pragma unroll
            for (int tj = 0; tj < 3; ++tj) dmma884(c[ti][tj][0], c[ti][tj][1], a[ti], b[tj]);
    }
}

// c1 -= A1[rb..] * B[cb..]^T and (TWO) c2 -= A2[rb..] * B[cb..]^T with the B fragments shared
template <bool TWO>
__device__ __forceinline__ void tile_abt_dmma_sub2(const double *A1, const double *A2, const double *B, int rb, int cb,
                                                   double c1[3][3][2], double c2[3][3][2])
{
    const int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
    const double *a1p = A1 + (rb + fr) * LDD + fk, *a2p = A2 + (rb + fr) * LDD + fk, *bp = B + (cb + fr) * LDD + fk;
#pragma unroll 2
    for (int ks = 0; ks < TS / 4; ++ks) {
        double a1[3], a2[3], b[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            a1[t] = -a1p[t * 8 * LDD + ks * 4]; b[t] = bp[t * 8 * LDD + ks * 4];
            if (TWO) a2[t] = -a2p[t * 8 * LDD + ks * 4];
        }
#pragma unroll
        for (int ti = 0; ti < 3; ++ti)
#pragma unroll
            for (int tj = 0; tj < 3; ++tj) {
                dmma884(c1[ti][tj][0], c1[ti][tj][1], a1[ti], b[tj]);
                if (TWO) dmma884(c2[ti][tj][0], c2[ti][tj][1], a2[ti], b[tj]);
            }
    }
}

// X = L^-1 for the lower-triangular 48x48 factor in smem, by recursive doubling: the eight 6x6 diagonal
// blocks are inverted by substitution (one thread per column, <= 5 dependent steps), then blocks are
// merged pairwise (h = 6, 12, 24):  inv [[A,0],[B,C]] = [[A^-1,0],[-C^-1 B A^-1, C^-1]].  The h x h
// scratch T = B A^-1 lives in the upper-right corner of the pair, which is zero in the result.
template <int H>
__device__ __forceinline__ void tri_inverse_merge(const double *L, double *X)
{
    const int tid = threadIdx.x;
    constexpr int NP = TS / (2 * H), PER = H * H;
    for (int e = tid; e < NP * PER; e += 256) {               // T = L21 * X11
        const int o = (e / PER) * 2 * H, r = (e % PER) / H, cc = e % H;
        double s = 0.0;
        for (int k = cc; k < H; ++k) s += L[(o + H + r) * LDT + o + k] * X[(o + k) * LDT + o + cc];
        X[(o + r) * LDT + o + H + cc] = s;
    }
    __syncthreads();
    for (int e = tid; e < NP * PER; e += 256) {               // X21 = -X22 * T
        const int o = (e / PER) * 2 * H, r = (e % PER) / H, cc = e % H;
        double s = 0.0;
        for (int k = 0; k <= r; ++k) s += X[(o + H + r) * LDT + o + H + k] * X[(o + k) * LDT + o + H + cc];
        X[(o + H + r) * LDT + o + cc] = -s;
    }
    __syncthreads();
    for (int e = tid; e < NP * PER; e += 256) {
        const int o = (e / PER) * 2 * H, r = (e % PER) / H, cc = e % H;
        X[(o + r) * LDT + o + H + cc] = 0.0;
    }
    __syncthreads();
}

__device__ __forceinline__ void tile_tri_inverse(const double *L, double *X)
{
    const int tid = threadIdx.x;
    for (int e = tid; e < TS * LDT; e += 256) X[e] = 0.0;
    __syncthreads();
    if (tid < TS) {
        const int b0 = (tid / 6) * 6, cidx = tid % 6, gc = b0 + cidx;
        double x[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double v = 0.0;
            if (r == cidx) v = 1.0 / L[gc * LDT + gc];
            else if (r > cidx) {
                double sum = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) if (k >= cidx && k < r) sum += L[(b0 + r) * LDT + b0 + k] * x[k];
                v = -sum / L[(b0 + r) * LDT + b0 + r];
            }
            x[r] = v;
            X[(b0 + r) * LDT + gc] = v;
        }
    }
    __syncthreads();
    tri_inverse_merge<6>(L, X);
    tri_inverse_merge<12>(L, X);
    tri_inverse_merge<24>(L, X);
}

#define PANEL_NT 128
// optional phase stamps of the diagonal CTA of every step (PSBA_PANEL_DEBUG=1): clock64 at phase boundaries
__device__ long long *g_panel_dbg = nullptr;
#define PANEL_STAMP(i) do { if (dbg && tid == 0) dbg[i] = clock64(); } while (0)
__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PANEL_WALL(i) do { if (dbg && tid == 0) dbg[i] = global_ns(); } while (0)
// One kernel per STEP (128 threads per CTA); a step holds every panel whose dependencies are met.
//  critical CTA for tile row I of panel K (I = K: the diagonal CTA): loads D = A_KK and A_IK, applies the
//  pending updates of the source panels P of the previous step (D -= L_KP L_KP^T, A_IK -= L_IP L_KP^T), then
//  ONE column sweep factors the stacked panel [D ; A_IK ; b_K^T] (97 x 48): l_ij = d_ij * rsqrt(d_jj),
//  d_ic -= l_ij l_cj.  Rows of D become L_KK, rows of A_IK become L_IK = A_IK L_KK^-T (no explicit inverse, no
//  separate trsm) and the extra row b_K^T becomes y_K^T = (L_KK^-1 b_K)^T, the forward substitution; b_K is the
//  right-hand side minus the contributions L_KP y_P that the CTAs of the tiles (K,P) left behind (summed in
//  ascending P: fixed order, no atomics).  Every thread keeps one 6x3 block of D and one of A_IK in registers
//  for the updates and one row of the stacked panel for the sweep; per column only the 97 column entries go
//  through shared memory (double-buffered, one barrier per column); the loop body is branch-free (dead
//  entries are updated too, a non-positive pivot poisons the panel with NaN and is reported once at the end).
//  deferred CTAs: A_IJ -= sum_P L_IP L_JP^T for trailing tiles whose panel J runs in a later step.
// MOD: the MODIFIED factorisation of the trust-region fallback (cholmod_blk.cl:446-475 on the tile pool): a pivot that is not
// positive and finite is replaced by max(|pivot|, delta) instead of failing, and the 3-column blocks that needed it are counted.
// FLOW: ONE launch for the whole factorisation.  Block b runs task tasks[b] (step order); instead of a kernel boundary it
// waits -- one thread per dependency, acquire loads -- until the write counters of the tiles it reads have reached the
// values the static schedule prescribes (flow tables of psba_build_tile_structure), reads everything another CTA of
// the launch may have written through L2 (ld.cg), and bumps the counter of what it wrote with a release.  A panel's
// critical chain then runs ahead of the bulk of the trailing updates of its step, and nothing pays a launch gap.
struct flow_args { const int2 *tasks; const int *tfinal, *defseq, *bseq, *critneed; int *ver; int n_tiles; };
template <bool MOD, bool FLOW>
__global__ void __launch_bounds__(PANEL_NT) k_panel_step(int ncrit, const int4 *__restrict__ crit_desc, const int2 *__restrict__ crit_src,
                                                         int ndef, const int4 *__restrict__ def_desc, const int2 *__restrict__ def_srcs,
                                                         const int *__restrict__ bJ, const int *__restrict__ b_sptr, const int *__restrict__ b_slot,
                                                         double *Stiles, double *__restrict__ Ldiag, double *bwork,
                                                         double *__restrict__ ywork, double *contrib, int *status,
                                                         double delta, int *__restrict__ nmod, flow_args fl)
{
    extern __shared__ double smem[];
    __shared__ __align__(16) double colbuf[2][2 * TS + 4];
    // programmatic dependent launch: the CTAs of step s+1 are scheduled while step s drains; they fetch
    // their (static) task descriptors and then wait for the previous step's memory to become visible
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int4 cd0 = make_int4(0, 0, 0, 0), cd1 = cd0;
    int2 sl0 = make_int2(-1, -1);                              // first source of the task: static data, fetched before the wait too
    int role, li;                                             // 0 critical, 1 deferred, 2 right-hand side; index in the task arrays
    if (FLOW) { const int2 tk = __ldg(fl.tasks + blockIdx.x); role = tk.x; li = tk.y; }
    else { role = (int)blockIdx.x < ncrit ? 0 : ((int)blockIdx.x < ncrit + ndef ? 1 : 2); li = role == 0 ? blockIdx.x : (role == 1 ? blockIdx.x - ncrit : blockIdx.x - ncrit - ndef); }
    if (role == 0) {
        cd0 = __ldg(crit_desc + 2 * li); cd1 = __ldg(crit_desc + 2 * li + 1);
        if (cd1.x < cd1.y) sl0 = __ldg(crit_src + cd1.x);
    } else if (role == 1) {
        cd0 = __ldg(def_desc + li);
        sl0 = __ldg(def_srcs + cd0.y);
    }
    if (!FLOW) asm volatile("griddepcontrol.wait;" ::: "memory");
    // the status word of the earlier steps is loaded here and tested after the first tile loads have been issued:
    // one L2 round trip less in front of every step
    int failed = FLOW ? 0 : *reinterpret_cast<volatile int *>(status);
    const int tid = threadIdx.x;
    double *B0 = smem, *B1 = smem + TILE_SM;
    int sig = -1;                                             // FLOW: the counter this task bumps when it is done
    if (FLOW) {
        // one thread per dependency
        bool ok = true;
        if (role == 0) {
            const int nsrc = cd1.y - cd1.x, K_ = cd0.y;
            if (cd0.x != cd0.y) sig = cd0.z;
            for (int d = tid; d < 3 + 2 * nsrc; d += PANEL_NT) {
                int idx, need;
                if (d == 0) { idx = cd0.z; need = __ldg(fl.critneed + 2 * li); }
                else if (d == 1) { idx = cd0.w; need = __ldg(fl.critneed + 2 * li + 1); }
                else if (d == 2) { idx = fl.n_tiles + K_; need = __ldg(fl.tfinal + idx); }
                else { const int2 sl = __ldg(crit_src + cd1.x + (d - 3) / 2); idx = ((d - 3) & 1) ? sl.y : sl.x; need = idx >= 0 ? __ldg(fl.tfinal + idx) : 0; }
                if (idx >= 0) ok = flow_wait(fl.ver, idx, need, status) && ok;
            }
        } else if (role == 1) {
            const int nsrc = cd0.z - cd0.y;
            sig = cd0.x;
            for (int d = tid; d < 1 + 2 * nsrc; d += PANEL_NT) {
                int idx, need;
                if (d == 0) { idx = cd0.x; need = __ldg(fl.defseq + li); }
                else { const int2 sl = __ldg(def_srcs + cd0.y + (d - 1) / 2); idx = ((d - 1) & 1) ? sl.y : sl.x; need = __ldg(fl.tfinal + idx); }
                ok = flow_wait(fl.ver, idx, need, status) && ok;
            }
        } else {
            const int q0 = b_sptr[li], q1 = b_sptr[li + 1], J_ = bJ[li];
            sig = fl.n_tiles + J_;
            for (int d = tid; d < 1 + (q1 - q0); d += PANEL_NT) {
                int idx, need;
                if (d == 0) { idx = fl.n_tiles + J_; need = __ldg(fl.bseq + li); }
                else { idx = b_slot[q0 + d - 1]; need = __ldg(fl.tfinal + idx); }
                ok = flow_wait(fl.ver, idx, need, status) && ok;
            }
        }
        failed = __syncthreads_and(ok) ? 0 : 1;
    }
    // FLOW: whatever happens, the task's counter is bumped (a failed factorisation must drain, not hang)
#define FLOW_DONE() do { if (FLOW && sig >= 0) { __syncthreads(); if (tid == 0) red_release_add(fl.ver + sig, 1); } } while (0)

    if (role == 2) {
        // ---- right-hand side of a later panel:  b_J -= sum_P L_JP y_P  (ascending P)
        if (failed) { FLOW_DONE(); return; }
        const int t = li;
        if (tid < TS) {
            double sum = 0.0;
            for (int q = b_sptr[t]; q < b_sptr[t + 1]; ++q) sum += ld1c<FLOW>(contrib + (size_t)b_slot[q] * TS + tid);
            bwork[bJ[t] * TS + tid] = ld1c<FLOW>(bwork + bJ[t] * TS + tid) - sum;
        }
        FLOW_DONE();
        return;
    }
    if (role == 1) {
        // ---- deferred trailing update:  A_IJ -= sum_P L_IP L_JP^T
        const int4 dd = cd0;
        const int sb = dd.y, se = dd.z;
        double *tij = Stiles + (size_t)dd.x * TS * TS;
        TileRegs<PANEL_NT> ra, rb_;
        tile_ldg<PANEL_NT, FLOW>(ra, Stiles + (size_t)sl0.x * TS * TS);
        tile_ldg<PANEL_NT, FLOW>(rb_, Stiles + (size_t)sl0.y * TS * TS);
        if (failed) { FLOW_DONE(); return; }
        // warp w owns the 24x24 block (w / 2, w % 2) of the target: 3x3 DMMA fragments, two doubles per lane each
        const int wrp = tid >> 5, lane = tid & 31, rb = (wrp >> 1) * 24, cb = (wrp & 1) * 24;
        double cf[3][3][2];
#pragma unroll
        for (int ti = 0; ti < 3; ++ti)
#pragma unroll
            for (int tj = 0; tj < 3; ++tj) {
                const double2 v = ld2c<FLOW>(tij + (rb + ti * 8 + (lane >> 2)) * TS + cb + tj * 8 + 2 * (lane & 3));
                cf[ti][tj][0] = v.x; cf[ti][tj][1] = v.y;
            }
        double *D0 = smem, *D1 = smem + TS * LDD;
        for (int s = sb; s < se; ++s) {
            if (s > sb) __syncthreads();
            tile_sts_ldd(D0, ra); tile_sts_ldd(D1, rb_);
            __syncthreads();
            if (s + 1 < se) {                                  // next source's tiles fly during the product
                const int2 sl = __ldg(def_srcs + s + 1);
                tile_ldg<PANEL_NT, FLOW>(ra, Stiles + (size_t)sl.x * TS * TS);
                tile_ldg<PANEL_NT, FLOW>(rb_, Stiles + (size_t)sl.y * TS * TS);
            }
            tile_abt_dmma_sub(D0, D1, rb, cb, cf);
        }
#pragma unroll
        for (int ti = 0; ti < 3; ++ti)
#pragma unroll
            for (int tj = 0; tj < 3; ++tj)
                *reinterpret_cast<double2 *>(tij + (rb + ti * 8 + (lane >> 2)) * TS + cb + tj * 8 + 2 * (lane & 3)) = make_double2(cf[ti][tj][0], cf[ti][tj][1]);
        FLOW_DONE();
        return;
    }

    // ---- critical path of panel K for tile row I
    const int I = cd0.x, K = cd0.y;
    const bool diagcta = I == K;
    long long *dbg = (g_panel_dbg && (FLOW ? diagcta : blockIdx.x == 0)) ? g_panel_dbg + (size_t)K * 8 : nullptr;
    PANEL_STAMP(0); PANEL_WALL(6);
    const int sb = cd1.x, se = cd1.y;
    const int slot_ik = cd0.z;
    double *tik = Stiles + (size_t)slot_ik * TS * TS;
    const double *tkk = Stiles + (size_t)cd0.w * TS * TS;
    // all global loads are issued before anything waits on them
    TileRegs<PANEL_NT> rp, ri;
    bool upd = false;
    if (sb < se) {
        tile_ldg<PANEL_NT, FLOW>(rp, Stiles + (size_t)sl0.x * TS * TS);                           // L_KP
        upd = sl0.y >= 0;
        if (upd) tile_ldg<PANEL_NT, FLOW>(ri, Stiles + (size_t)sl0.y * TS * TS);                  // L_IP
    }
    // D and A_IK as DMMA accumulator fragments: warp w owns the 24x24 block (w / 2, w % 2), 3x3 fragments of 8x8,
    // lane l holds the entries (l / 4, 2 (l % 4) + {0, 1}) of each
    const int wrp = tid >> 5, lane = tid & 31, rb = (wrp >> 1) * 24, cb = (wrp & 1) * 24;
    double d[3][3][2], a[3][3][2];
#pragma unroll
    for (int ti = 0; ti < 3; ++ti)
#pragma unroll
        for (int tj = 0; tj < 3; ++tj) {
            const int off = (rb + ti * 8 + (lane >> 2)) * TS + cb + tj * 8 + 2 * (lane & 3);
            const double2 dv = ld2c<FLOW>(tkk + off);
            d[ti][tj][0] = dv.x; d[ti][tj][1] = dv.y;
            double2 av = make_double2(0.0, 0.0);
            if (!diagcta) av = ld2c<FLOW>(tik + off);
            a[ti][tj][0] = av.x; a[ti][tj][1] = av.y;
        }
    if (failed) { FLOW_DONE(); return; }
    // right-hand side of the panel: what the rhs tasks of earlier steps left in bwork minus the
    // contributions of the source panels (ascending P)
    PANEL_STAMP(1);
    double bk = 0.0;
    if (tid < TS) {
        double sum = 0.0;
        for (int s = sb; s < se; ++s) sum += ld1c<FLOW>(contrib + (size_t)__ldg(crit_src + s).x * TS + tid);
        bk = ld1c<FLOW>(bwork + K * TS + tid) - sum;
    }
    for (int s = sb; s < se; ++s) {
        if (s > sb) {                                          // further sources (first panel of a separator)
            __syncthreads();
            const int2 sl = __ldg(crit_src + s);
            tile_ldg<PANEL_NT, FLOW>(rp, Stiles + (size_t)sl.x * TS * TS);
            upd = sl.y >= 0;
            if (upd) tile_ldg<PANEL_NT, FLOW>(ri, Stiles + (size_t)sl.y * TS * TS);
        }
        tile_sts_ldd(B0, rp);
        if (upd) tile_sts_ldd(B0 + TS * LDD, ri);
        __syncthreads();
        if (upd) tile_abt_dmma_sub2<true>(B0, B0 + TS * LDD, B0, rb, cb, d, a);
        else tile_abt_dmma_sub2<false>(B0, B0, B0, rb, cb, d, a);
    }
    PANEL_STAMP(2);
    // ---- hand the updated blocks over to the ROW-OWNER layout of the sweep: thread t < 48 owns row t of
    // D, thread 48+r owns row r of A_IK, thread 96 owns the right-hand-side row b_K^T
    __syncthreads();                                           // the factor tiles in B0/B1 are consumed
#pragma unroll
    for (int ti = 0; ti < 3; ++ti)
#pragma unroll
        for (int tj = 0; tj < 3; ++tj) {
            const int off = (rb + ti * 8 + (lane >> 2)) * LDT + cb + tj * 8 + 2 * (lane & 3);
            B0[off] = d[ti][tj][0]; B0[off + 1] = d[ti][tj][1];
            B1[off] = a[ti][tj][0]; B1[off + 1] = a[ti][tj][1];
        }
    if (tid < TS) colbuf[0][tid] = bk;
    __syncthreads();
    const bool active = tid < TS || (tid < 2 * TS && !diagcta) || tid == 2 * TS;
    double row[TS];
    if (tid == 2 * TS) {
#pragma unroll
        for (int c = 0; c < TS; ++c) row[c] = colbuf[0][c];
    } else {
        const double *src = tid < TS ? B0 + tid * LDT : B1 + (tid - TS) * LDT;
#pragma unroll
        for (int c = 0; c < TS; ++c) row[c] = active ? src[c] : 0.0;
    }
    __syncthreads();                                           // colbuf[0] is reused below
    PANEL_STAMP(3);
    // ---- blocked sweep over the stacked panel [D ; A_IK ; b^T] (97 x 48), one row per thread in registers,
    // SB = 6 columns per block step (8 steps, two barriers each instead of one barrier per column):
    //   A  every thread factors the current 6x6 diagonal block redundantly in registers (the only truly
    //      sequential part: 6 dependent rsqrt) and runs its own row through it (l_i = row_i L_bb^-T);
    //   B  the rows of D publish their six factor entries;
    //   C1 every thread updates the six columns of the NEXT block, whose owners publish the next diagonal
    //      block immediately;
    //   C2 the remaining columns are updated at the top of the next step, in the same basic block as the
    //      latency-bound factorisation A so that the FMA stream hides under the rsqrt chain.
    // Dead entries (above the diagonal of D) are updated too: the loop is branch-free; a non-positive pivot
    // poisons the panel with NaN and is reported once at the end.
    constexpr int SB = 6, NB = TS / SB;
    double *pubD = colbuf[0];                                  // [2][SB*SB]
    double *pubL = smem;                                       // [2][TS*SB]  (the tiles in B0 are consumed)
    bool bad = false;
    unsigned modmask = 0;                                      // MOD: 3-column blocks of this panel with a replaced pivot
    if (tid < SB) {
#pragma unroll
        for (int k = 0; k < SB; ++k) pubD[tid * SB + k] = row[k];
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        __syncthreads();
        // A: 6x6 diagonal block, redundantly, fused with this thread's own row.  C2 of the previous block step
        // (columns beyond the current block) is dealt out in six chunks, one in front of every rsqrt, so
        // that its FMA stream runs under the latency of the MUFU + Newton chain (same basic block).
        {
            const double *pd = pubD + (b & 1) * SB * SB;
            const double *pl = pubL + ((b + 1) & 1) * TS * SB;         // l-values of block b-1
            double dd[SB][SB];
#pragma unroll
            for (int i = 0; i < SB; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) dd[i][j] = pd[i * SB + j];
            double *x = row + b * SB;
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                const double piv = dd[j][j];
                if (b > 0) {
                    const double *xp = row + (b - 1) * SB;
#pragma unroll
                    for (int c = (b + 1) * SB + j; c < TS; c += SB) {
                        const double2 l01 = *reinterpret_cast<const double2 *>(pl + c * SB);
                        const double2 l23 = *reinterpret_cast<const double2 *>(pl + c * SB + 2);
                        const double2 l45 = *reinterpret_cast<const double2 *>(pl + c * SB + 4);
                        double r = row[c];
                        r = fma(-xp[0], l01.x, r); r = fma(-xp[1], l01.y, r); r = fma(-xp[2], l23.x, r);
                        r = fma(-xp[3], l23.y, r); r = fma(-xp[4], l45.x, r); r = fma(-xp[5], l45.y, r);
                        row[c] = r;
                    }
                }
                double pv = piv;
                if (MOD) {
                    if (!(piv > 0.0 && piv < 1e300)) {           // cholmod_blk.cl:465-475: d = max(|d|, delta); NaN -> delta
                        pv = fmax(fabs(piv), delta);
                        if (!(pv < 1e300)) pv = delta;
                        modmask |= 1u << ((b * SB + j) / 3);
                        if (tid == b * SB + j) x[j] = pv;      // the row of D that holds this diagonal entry
                    }
                } else bad |= !(piv > 0.0 && piv < 1e300);
                const double inv = rsqrt(pv);
                x[j] *= inv;
#pragma unroll
                for (int i = j + 1; i < SB; ++i) dd[i][j] *= inv;
#pragma unroll
                for (int i = j + 1; i < SB; ++i) {
                    x[i] -= x[j] * dd[i][j];
#pragma unroll
                    for (int k = j + 1; k <= i; ++k) dd[i][k] -= dd[i][j] * dd[k][j];
                }
            }
        }
        if (b == NB - 1) break;
        // B: publish the factor entries of the rows of D
        double *plw = pubL + (b & 1) * TS * SB;
        if (tid < TS) {
            const double *x = row + b * SB;
            *reinterpret_cast<double2 *>(plw + tid * SB) = make_double2(x[0], x[1]);
            *reinterpret_cast<double2 *>(plw + tid * SB + 2) = make_double2(x[2], x[3]);
            *reinterpret_cast<double2 *>(plw + tid * SB + 4) = make_double2(x[4], x[5]);
        }
        __syncthreads();
        // C1: the columns of the next block, then its diagonal block goes out
        {
            const double *x = row + b * SB;
#pragma unroll
            for (int c = (b + 1) * SB; c < (b + 2) * SB; ++c) {
                const double2 l01 = *reinterpret_cast<const double2 *>(plw + c * SB);
                const double2 l23 = *reinterpret_cast<const double2 *>(plw + c * SB + 2);
                const double2 l45 = *reinterpret_cast<const double2 *>(plw + c * SB + 4);
                row[c] -= (x[0] * l01.x + x[1] * l01.y) + (x[2] * l23.x + x[3] * l23.y) + (x[4] * l45.x + x[5] * l45.y);
            }
            if (tid >= (b + 1) * SB && tid < (b + 2) * SB) {
                double *pdw = pubD + ((b + 1) & 1) * SB * SB + (tid - (b + 1) * SB) * SB;
#pragma unroll
                for (int k = 0; k < SB; ++k) pdw[k] = row[(b + 1) * SB + k];
            }
        }
    }
    PANEL_STAMP(4);
    if (__syncthreads_or(bad)) { if (tid == 0) atomicMax(status, 1); FLOW_DONE(); return; }
    if (MOD && diagcta && tid == 0 && modmask) atomicAdd(nmod, __popc(modmask));
    // ---- results: factor rows to global, y_K to shared, then the contribution L_IK y_K of this tile
    double *ysh = colbuf[0];
    if (tid == 2 * TS) {
#pragma unroll
        for (int c = 0; c < TS; ++c) ysh[c] = row[c];
        if (diagcta) {
#pragma unroll
            for (int c = 0; c < TS; ++c) ywork[K * TS + c] = row[c];
        }
    }
    // the rows leave through shared memory so that the stores are coalesced (a thread that writes its own 384-byte row issues
    // 24 stores that each touch 32 different lines across the warp)
    double *stg = smem + TILE_SM;                              // the sweep's publication buffers live in the first tile
    if (diagcta) {
        if (tid < TS) {
#pragma unroll
            for (int c = 0; c < TS; ++c) stg[tid * LDT + c] = row[c];
        }
        __syncthreads();
        tile_stg<PANEL_NT>(Ldiag + (size_t)K * TS * TS, stg, true);
        PANEL_STAMP(5); PANEL_WALL(7);
        return;
    }
    if (tid >= TS && tid < 2 * TS) {
#pragma unroll
        for (int c = 0; c < TS; ++c) stg[(tid - TS) * LDT + c] = row[c];
    }
    __syncthreads();
    tile_stg<PANEL_NT>(tik, stg);
    if (tid >= TS && tid < 2 * TS) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < TS; c += 2) s += row[c] * ysh[c] + row[c + 1] * ysh[c + 1];
        contrib[(size_t)slot_ik * TS + (tid - TS)] = s;
    }
    FLOW_DONE();
}
#undef FLOW_DONE
// L_KK^-1 for every diagonal tile (needed by the backward substitution only): one CTA per tile,
// all tiles in parallel, off the critical path of the factorisation
__global__ void __launch_bounds__(256) k_diag_inverse(const double *__restrict__ Ldiag, double *__restrict__ Linv, const int *__restrict__ status)
{
    extern __shared__ double smem[];
    if (*status != 0) return;
    double *B0 = smem, *B1 = smem + TILE_SM;
    TileRegs<256> r;
    tile_ldg(r, Ldiag + (size_t)blockIdx.x * TS * TS);
    tile_sts(B0, r);
    __syncthreads();
    tile_tri_inverse(B0, B1);
    tile_stg<256>(Linv + (size_t)blockIdx.x * TS * TS, B1);
}

// b0 = ea in the camera ordering of S (zero on the padding)
__global__ void k_init_rhs(int npad, const int *__restrict__ pos2cam, const double *__restrict__ ea, double *__restrict__ b, double *__restrict__ xbuf)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npad) return;
    const int cam = pos2cam[k / 6];
    b[k] = cam >= 0 ? ea[cam * 6 + k % 6] : 0.0;
    if (xbuf) reinterpret_cast<unsigned long long *>(xbuf)[k] = X_SENTINEL;   // the backward solve of this factorisation starts from "nothing known"
}

static void enqueue_factor(psba_ctx *c, std::vector<cudaEvent_t> *ev = nullptr, bool mod = false, double delta = 0.0)
{
    const int npad = c->nt * TS;
    k_init_rhs<<<cdiv(npad, 256), 256, 0, c->stream>>>(npad, c->pos2cam, c->eab, c->chol_aux, c->bw_xbuf);
    flow_args fl = {c->d_flow_tasks, c->d_flow_final, c->d_flow_defseq, c->d_flow_bseq, c->d_flow_critneed, c->d_flow_ver, c->n_tiles};
    if (c->chol_flow && !ev) {
        // dataflow: one launch for every step; the write counters start from zero
        CUDA_CHECK(cudaMemsetAsync(c->d_flow_ver, 0, (size_t)(c->n_tiles + c->nt) * sizeof(int), c->stream));
        auto kern = mod ? k_panel_step<true, true> : k_panel_step<false, true>;
        psba_set_smem((const void *)kern, (int)CHOL_SMEM);
        kern<<<c->n_flow_tasks, PANEL_NT, CHOL_SMEM, c->stream>>>(0, (const int4 *)c->d_crit_desc, (const int2 *)c->d_crit_src, 0, (const int4 *)c->d_def_desc,
                                                              (const int2 *)c->d_def_srcs, (const int *)c->d_b_J, (const int *)c->d_b_sptr, (const int *)c->d_b_slot,
                                                              c->Stiles, c->Ldiag, c->chol_aux, c->chol_diag, c->contrib, c->d_status, delta, c->d_status + 2, fl);
        k_diag_inverse<<<c->nt, 256, 2 * TILE_SM * sizeof(double), c->stream>>>(c->Ldiag, c->Linv, c->d_status);
        return;
    }
    for (int s = 0; s < c->n_steps; ++s) {
        if (ev) CUDA_CHECK(cudaEventRecord((*ev)[s], c->stream));
        const int cb = c->step_crit_ptr[s], ncrit = c->step_crit_ptr[s + 1] - cb;
        const int db = c->step_def_ptr[s], ndef = c->step_def_ptr[s + 1] - db;
        const int bb = c->step_b_ptr[s], nb = c->step_b_ptr[s + 1] - bb;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ncrit + ndef + nb); cfg.blockDim = dim3(PANEL_NT); cfg.dynamicSmemBytes = CHOL_SMEM; cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = c->chol_pdl ? 1 : 0;
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, mod ? k_panel_step<true, false> : k_panel_step<false, false>, ncrit, (const int4 *)(c->d_crit_desc + 2 * cb),
                                      (const int2 *)c->d_crit_src, ndef,
                                      (const int4 *)(c->d_def_desc + db), (const int2 *)c->d_def_srcs, (const int *)(c->d_b_J + bb),
                                      (const int *)(c->d_b_sptr + bb), (const int *)c->d_b_slot, c->Stiles, c->Ldiag, c->chol_aux,
                                      c->chol_diag, c->contrib, c->d_status, delta, c->d_status + 2, fl));
    }
    if (ev) CUDA_CHECK(cudaEventRecord((*ev)[c->n_steps], c->stream));
    k_diag_inverse<<<c->nt, 256, 2 * TILE_SM * sizeof(double), c->stream>>>(c->Ldiag, c->Linv, c->d_status);
}

// defer_status: do not wait for the outcome here (the fused try reads the status word together with the
// step scalars after the back-substitution; the kernels behind a failed factorisation run on garbage and
// their results are discarded)
double psba_launch_factor(psba_ctx *c, bool defer_status)
{
    CUDA_CHECK(cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream));
    if (!c->chol_graph_ok) {
        psba_set_smem((const void *)k_panel_step<false, false>, (int)CHOL_SMEM);
        cudaGraph_t graph;
        CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        enqueue_factor(c);
        CUDA_CHECK(cudaStreamEndCapture(c->stream, &graph));
        CUDA_CHECK(cudaGraphInstantiate(&c->chol_graph, graph, 0));
        CUDA_CHECK(cudaGraphDestroy(graph));
        c->chol_graph_ok = true;
    }
    static int dbg_runs = getenv("PSBA_PANEL_DEBUG") ? 2 : 0;
    long long *dbg_dev = nullptr;
    if (dbg_runs > 0) {
        CUDA_CHECK(cudaMalloc(&dbg_dev, (size_t)c->nt * 8 * sizeof(long long)));
        CUDA_CHECK(cudaMemset(dbg_dev, 0, (size_t)c->nt * 8 * sizeof(long long)));
        CUDA_CHECK(cudaMemcpyToSymbol(g_panel_dbg, &dbg_dev, sizeof(dbg_dev)));
    }
    static int step_timing = getenv("PSBA_STEP_TIMING") ? 3 : 0;      // third factorisation: per-step device times, no graph
    if (step_timing > 0 && --step_timing == 0) {
        std::vector<cudaEvent_t> ev(c->n_steps + 1);
        for (auto &e : ev) CUDA_CHECK(cudaEventCreate(&e));
        const bool pdl = c->chol_pdl; c->chol_pdl = false;
        enqueue_factor(c, &ev);
        c->chol_pdl = pdl;
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        for (int s = 0; s < c->n_steps; ++s) {
            float ms = 0; CUDA_CHECK(cudaEventElapsedTime(&ms, ev[s], ev[s + 1]));
            fprintf(stderr, "step %3d: %7.1f us  crit %5d  deferred %5d  rhs %4d\n", s, ms * 1e3, c->step_crit_ptr[s + 1] - c->step_crit_ptr[s],
                    c->step_def_ptr[s + 1] - c->step_def_ptr[s], c->step_b_ptr[s + 1] - c->step_b_ptr[s]);
        }
        for (auto &e : ev) cudaEventDestroy(e);
    } else if (c->capturing) enqueue_factor(c);                  // inside a captured chain (psba_seq_begin): the step kernels join that graph
    else
    PROF(c, KID_FACTOR) CUDA_CHECK(cudaGraphLaunch(c->chol_graph, c->stream));
    if (dbg_dev) {
        std::vector<long long> h((size_t)c->nt * 8);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        CUDA_CHECK(cudaMemcpy(h.data(), dbg_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        long long *nul = nullptr;
        CUDA_CHECK(cudaMemcpyToSymbol(g_panel_dbg, &nul, sizeof(nul)));
        CUDA_CHECK(cudaFree(dbg_dev));
        if (--dbg_runs == 0)
            for (int K = 0; K < c->nt; ++K)
                if (h[(size_t)K * 8])
                    fprintf(stderr, "panel %4d: loads %6lld  update %6lld  handover %6lld  sweep %6lld  store %6lld  total %6lld clk   start %lld end %lld ns\n", K,
                            h[K * 8 + 1] - h[K * 8], h[K * 8 + 2] - h[K * 8 + 1], h[K * 8 + 3] - h[K * 8 + 2], h[K * 8 + 4] - h[K * 8 + 3],
                            h[K * 8 + 5] - h[K * 8 + 4], h[K * 8 + 5] - h[K * 8], h[K * 8 + 6] - h[7], h[K * 8 + 7] - h[7]);
    }
    c->st_launches += c->n_steps + 2;
    LAUNCH_CHECK();
    c->S_valid = false;      // the factor overwrote the tile pool
    if (defer_status) { c->factor_valid = true; return 0.0; }
    int st = 0;
    CUDA_CHECK(cudaMemcpyAsync(&st, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->factor_valid = (st == 0);
    c->S_valid = false;      // the factor overwrote the tile pool
    return st ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Modified Cholesky of the trust-region fallback ON THE TILE POOL (cholmod_blk / get_delta_beta / compute_cholmod_E,
// PSBA/cl_cholmod.cpp:25-202 -> CL_files/cholmod_blk.cl:87-847) for camera systems that are too large for the dense
// single-CTA kernel of kernels_solve.cu.  Same quantities -- xi, gamma -> delta, beta; L L^T = S + diag(E);
// E_i = sum_k L_ik^2 - S_ii; the number of 3-column blocks that took the modified path -- computed by the tiled,
// multi-CTA step schedule above, with two documented differences from the reference's kernel:
//   * the elimination runs in the solver's nested-dissection ORDER (P S P^T), not in the callers' camera order.  E
//     is nonzero only where a pivot was replaced or at rounding level elsewhere, and lambda = |sum E| / N is a
//     rounding-noise quantity by construction, so the value differs like it differs between two builds
//     of the reference itself;
//   * the `L_ij > beta` rescue (cholmod_blk.cl:534-611: theta / beta) is not applied; the largest factor entry is
//     measured instead and reported (stat "cholmod_max_l_over_beta" > 1 means the reference would have rescaled).
__global__ void k_tile_maxabs(int n_tiles_S, int nt, const int *__restrict__ tile_index, const int *__restrict__ pos2cam,
                              const double *__restrict__ Stiles, double *__restrict__ diag_out, double *__restrict__ part2)
{
    // one CTA per tile slot (I, J) looked up from the index: |offdiag| and |diag| maxima, diagonal saved for E
    __shared__ double sx[256], sg[256];
    const int I = blockIdx.x / nt, J = blockIdx.x % nt;
    double xi = 0.0, ga = 0.0;
    const int slot = J <= I ? tile_index[I * nt + J] : -1;
    if (slot >= 0 && slot < n_tiles_S) {
        const double *t = Stiles + (size_t)slot * TS * TS;
        for (int e = threadIdx.x; e < TS * TS; e += 256) {
            const int r = e / TS, cc = e % TS;
            const bool real = pos2cam[(I * TS + r) / 6] >= 0 && pos2cam[(J * TS + cc) / 6] >= 0;
            if (!real || (I == J && cc > r)) continue;
            const double v = fabs(t[e]);
            if (I == J && r == cc) { ga = fmax(ga, v); diag_out[I * TS + r] = t[e]; } else xi = fmax(xi, v);
        }
    }
    sx[threadIdx.x] = xi; sg[threadIdx.x] = ga;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { sx[threadIdx.x] = fmax(sx[threadIdx.x], sx[threadIdx.x + w]); sg[threadIdx.x] = fmax(sg[threadIdx.x], sg[threadIdx.x + w]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part2[2 * blockIdx.x] = sx[0]; part2[2 * blockIdx.x + 1] = sg[0]; }
}
__global__ void k_max_pairs(int n, const double *__restrict__ part2, double *__restrict__ out2)
{
    __shared__ double sx[256], sg[256];
    double xi = 0.0, ga = 0.0;
    for (int q = threadIdx.x; q < n; q += 256) { xi = fmax(xi, part2[2 * q]); ga = fmax(ga, part2[2 * q + 1]); }
    sx[threadIdx.x] = xi; sg[threadIdx.x] = ga;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { sx[threadIdx.x] = fmax(sx[threadIdx.x], sx[threadIdx.x + w]); sg[threadIdx.x] = fmax(sg[threadIdx.x], sg[threadIdx.x + w]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = sx[0]; out2[1] = sg[0]; }
}
// E of row r (solver's ordering) = sum of the squares of row r of the factor - S_rr; also max |L_rc| off the diagonal
__global__ void k_tile_E(int nt, const int *__restrict__ tile_index, const int *__restrict__ pos2cam, const double *__restrict__ Stiles,
                         const double *__restrict__ Ldiag, const double *__restrict__ Sdiag, double *__restrict__ E_cam, double *__restrict__ rowmax)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nt * TS) return;
    const int I = r / TS, rr = r % TS, cam = pos2cam[r / 6];
    if (cam < 0) { rowmax[r] = 0.0; return; }
    double sum = 0.0, mx = 0.0;
    for (int J = 0; J < I; ++J) {
        const int slot = tile_index[I * nt + J];
        if (slot < 0) continue;
        const double *t = Stiles + (size_t)slot * TS * TS + rr * TS;
        for (int cc = 0; cc < TS; ++cc) { const double v = t[cc]; sum += v * v; mx = fmax(mx, v); }
    }
    const double *d = Ldiag + (size_t)I * TS * TS + rr * TS;
    for (int cc = 0; cc <= rr; ++cc) { const double v = d[cc]; sum += v * v; if (cc < rr) mx = fmax(mx, v); }
    E_cam[cam * 6 + r % 6] = sum - Sdiag[r];
    rowmax[r] = mx;                                            // signed maximum, as the reference's test `L_ij > beta`
}

// The dense single-CTA kernel of kernels_solve.cu implements the reference's kernel in full (3x3 block path, `> beta`
// roll-back, scalar Gill-Murray path with the theta / beta rescue) and is used up to N = 1536; the tile pool takes over
// beyond (a 1.15 GB dense copy and O(N^3) on one SM at N = 12 000).  The null-space pivots of a gauge-free camera system
// are rounding noise of either sign, and the beta / theta logic decides what replaces them: on the 7-camera set the
// per-pivot rule of the tile version flags 2 block columns where the reference flags 1.  lambda = |sum E| / N is noise in
// both.  PSBA_CHOLMOD_TILES=0/1 forces either.
bool psba_cholmod_use_tiles(psba_ctx *c)
{
    const char *e = getenv("PSBA_CHOLMOD_TILES");
    if (e) return atoi(e) != 0;
    return c->N > 1536;
}
bool psba_cholmod_dense_possible(psba_ctx *c) { return c->N <= 1536; }

// get_delta_beta (PSBA/cl_cholmod.cpp:109-167) on the S of the tile pool
void psba_tile_delta_beta(psba_ctx *c, double *delta, double *beta)
{
    const int nt = c->nt, npad = nt * TS, N = c->N;
    double *Sdiag = (double *)psba_dev_alloc(c, (size_t)npad * 8, true), *part2 = (double *)psba_dev_alloc(c, (size_t)nt * nt * 16, true);
    k_tile_maxabs<<<nt * nt, 256, 0, c->stream>>>(c->n_tiles_S, nt, c->tile_index, c->pos2cam, c->Stiles, Sdiag, part2);
    k_max_pairs<<<1, 256, 0, c->stream>>>(nt * nt, part2, c->d_scal + 8);
    LAUNCH_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(c->h_scal + 8, c->d_scal + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    psba_dev_free(c, Sdiag); psba_dev_free(c, part2);
    const double xi = c->h_scal[8], gamma = c->h_scal[9];
    *delta = 1e-15 * fmax(xi + gamma, 1.0);
    double b = fmax(gamma, 1e-15);
    *beta = sqrt(fmax(b, xi / sqrt((double)N * N - 1)));
}

double psba_launch_cholmod_tiles(psba_ctx *c, double *delta_out, double *beta_out, int *nmod_out, double *E_host, double *max_l_over_beta)
{
    const int nt = c->nt, npad = nt * TS, N = c->N;
    double *Sdiag = (double *)psba_dev_alloc(c, (size_t)npad * 8, true), *part2 = (double *)psba_dev_alloc(c, (size_t)nt * nt * 16, true);
    double *rowmax = (double *)psba_dev_alloc(c, (size_t)npad * 8, true);
    k_tile_maxabs<<<nt * nt, 256, 0, c->stream>>>(c->n_tiles_S, nt, c->tile_index, c->pos2cam, c->Stiles, Sdiag, part2);
    k_max_pairs<<<1, 256, 0, c->stream>>>(nt * nt, part2, c->d_scal + 8);
    LAUNCH_CHECK();
    CUDA_CHECK(cudaMemcpyAsync(c->h_scal + 8, c->d_scal + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    const double xi = c->h_scal[8], gamma = c->h_scal[9];
    const double delta = 1e-15 * fmax(xi + gamma, 1.0);           // cl_cholmod.cpp:161-164
    double beta = fmax(gamma, 1e-15);
    beta = sqrt(fmax(beta, xi / sqrt((double)N * N - 1)));
    CUDA_CHECK(cudaMemsetAsync(c->d_status, 0, 4 * sizeof(int), c->stream));
    psba_set_smem((const void *)k_panel_step<true, false>, (int)CHOL_SMEM);
    PROF(c, KID_CHOLMOD) enqueue_factor(c, nullptr, true, delta);  // once per trust-region phase: launched step by step, no graph
    k_tile_E<<<cdiv(npad, 128), 128, 0, c->stream>>>(nt, c->tile_index, c->pos2cam, c->Stiles, c->Ldiag, Sdiag, c->chol_E, rowmax);
    c->st_launches += c->n_steps + 5;
    LAUNCH_CHECK();
    std::vector<double> E(N), rm(npad);
    int st[4] = {0, 0, 0, 0};
    CUDA_CHECK(cudaMemcpyAsync(E.data(), c->chol_E, (size_t)N * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(rm.data(), rowmax, (size_t)npad * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(st, c->d_status, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    psba_dev_free(c, Sdiag); psba_dev_free(c, part2); psba_dev_free(c, rowmax);
    double sum = 0.0, lmax = 0.0;
    for (int i = 0; i < N; ++i) sum += E[i];                      // trust_region.cpp:358-362 (callers' camera order)
    for (int r = 0; r < npad; ++r) lmax = fmax(lmax, rm[r]);
    if (E_host) for (int i = 0; i < N; ++i) E_host[i] = E[i];
    if (delta_out) *delta_out = delta;
    if (beta_out) *beta_out = beta;
    if (nmod_out) *nmod_out = st[2];
    if (max_l_over_beta) *max_l_over_beta = lmax / beta;
    c->cholmod_max_l_over_beta = lmax / beta;
    c->S_valid = false; c->factor_valid = false;                 // the modified factor overwrote the tile pool
    return sum;
}

// ---------------------------------------------------------------------------------------------
// backward substitution x_I = L_II^-T (y_I - sum_{J>I} L_JI^T x_J).  A CTA of 960 threads = 48 columns x 20
// tile slots handles one panel: every thread streams one tile column (48 independent coalesced loads in
// flight, four independent FMA chains), partial sums are combined in a fixed order; L_II^-1 does not depend
// on x and is prefetched into shared memory while the tile column streams.  The solution leaves in the
// callers' camera order (pos2cam).
//   chain schedule (dense S): ONE persistent CTA walks the panels from last to first;
//   otherwise: one launch per step in reverse order, one CTA per panel of the step (all in one CUDA graph).
#define BW_SLOTS 20
struct bw_smem {
    double part[BW_SLOTS][TS];
    double acc[TS];
    double invs[TS * TS];
    double mv[4][TS];
};

__device__ __forceinline__ void backward_panel(bw_smem &sm, int I, int beg, int end, int myslot, int myrow, const int *__restrict__ crow,
                                               const int *__restrict__ cslot, const double *__restrict__ Stiles,
                                               const double *__restrict__ Linv, double *__restrict__ ywork,
                                               const int *__restrict__ pos2cam, double *__restrict__ sol)
{
    const int tid = threadIdx.x, col = tid % TS, slotid = tid / TS;
    for (int e = tid; e < TS * TS; e += TS * BW_SLOTS) sm.invs[e] = __ldg(Linv + (size_t)I * TS * TS + e);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int t = beg + slotid; t < end; t += BW_SLOTS) {
        const bool first = t == beg + slotid;
        const double *L = Stiles + (size_t)(first ? myslot : cslot[t]) * TS * TS + col;
        const double *x = ywork + (first ? myrow : crow[t]) * TS;
        double lv[TS];
#pragma unroll
        for (int r = 0; r < TS; ++r) lv[r] = __ldg(L + r * TS);
#pragma unroll
        for (int r = 0; r < TS; r += 4) { s0 += lv[r] * x[r]; s1 += lv[r + 1] * x[r + 1]; s2 += lv[r + 2] * x[r + 2]; s3 += lv[r + 3] * x[r + 3]; }
    }
    sm.part[slotid][col] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (tid < TS) {
        double a = 0.0;
#pragma unroll
        for (int p = 0; p < BW_SLOTS; ++p) a += sm.part[p][tid];
        sm.acc[tid] = ywork[I * TS + tid] - a;
    }
    __syncthreads();
    if (tid < 4 * TS) {          // x_I[c] = sum_{r >= c} Linv[r][c] acc[r], rows split over 4 partial sums
        const int cidx = tid % TS, q = tid / TS;
        double a = 0.0;
        for (int r = cidx + q; r < TS; r += 4) a += sm.invs[r * TS + cidx] * sm.acc[r];
        sm.mv[q][cidx] = a;
    }
    __syncthreads();
    if (tid < TS) {
        const double a = (sm.mv[0][tid] + sm.mv[1][tid]) + (sm.mv[2][tid] + sm.mv[3][tid]);
        ywork[I * TS + tid] = a;
        const int pos = I * TS + tid, cam = pos2cam[pos / 6];
        if (cam >= 0) sol[cam * 6 + pos % 6] = a;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TS * BW_SLOTS) k_backward(int nt, const int *__restrict__ cptr, const int *__restrict__ crow,
                                                           const int *__restrict__ cslot, const double *__restrict__ Stiles,
                                                           const double *__restrict__ Linv, double *__restrict__ ywork,
                                                           const int *__restrict__ pos2cam, double *__restrict__ sol)
{
    __shared__ bw_smem sm;
    const int slotid = threadIdx.x / TS;
    // tile indices of a column do not depend on x: they are fetched one step ahead so that the only
    // dependent global accesses inside a step are the tile column itself and the x vectors
    int nbeg = cptr[nt - 1], nend = cptr[nt];
    int nslot = -1, nrow = 0;
    if (nbeg + slotid < nend) { nslot = cslot[nbeg + slotid]; nrow = crow[nbeg + slotid]; }
    for (int I = nt - 1; I >= 0; --I) {
        const int beg = nbeg, end = nend, myslot = nslot, myrow = nrow;
        if (I > 0) {
            nbeg = cptr[I - 1]; nend = beg;
            nslot = -1;
            if (nbeg + slotid < nend) { nslot = cslot[nbeg + slotid]; nrow = crow[nbeg + slotid]; }
        }
        backward_panel(sm, I, beg, end, myslot, myrow, crow, cslot, Stiles, Linv, ywork, pos2cam, sol);
    }
}

// ---- dataflow variant for tree schedules: ONE launch, one CTA per panel in reverse step order.  A CTA
// prefetches everything that does not depend on x (its tile column into registers, L_II^-1 into shared
// memory), then waits for the x_J it needs on per-panel flags (release / acquire, epoch-stamped: no reset
// between solves).  Forward progress: a CTA only waits for CTAs with a smaller block index, which the
// hardware dispatches first; a bounded spin turns a broken schedule into an error instead of a hang.
#define BWD_SLOTS 10
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define BW_STAMP(i) do { if (dbg && threadIdx.x == 0) dbg[i] = gtimer(); } while (0)
#define FLAG_STRIDE 32        // ints between two flags (128 bytes)

__global__ void __launch_bounds__(TS * BWD_SLOTS, 1) k_backward_flow(const int *__restrict__ order, const int *__restrict__ cptr,
                                                                    const int *__restrict__ crow, const int *__restrict__ cslot,
                                                                    const double *__restrict__ Stiles, const double *__restrict__ Linv,
                                                                    const double *__restrict__ ywork, const int *__restrict__ pos2cam, double *__restrict__ sol,
                                                                    double *xbuf, int *__restrict__ status)
{
    __shared__ double part[BWD_SLOTS][TS];
    __shared__ double xs[BWD_SLOTS][TS];
    __shared__ double acc[TS];
    __shared__ double invs[TS * TS];
    __shared__ double mv[4][TS];
    const int tid = threadIdx.x, col = tid % TS, slotid = tid / TS;
    const int I = order[blockIdx.x];
    const int beg = cptr[I], end = cptr[I + 1];
    long long *dbg = g_panel_dbg ? g_panel_dbg + (size_t)I * 8 : nullptr;
    BW_STAMP(0);
    // independent of x: first tile column of this slot, L_II^-1, y_I
    double lv[TS];
    int myrow = -1;
    if (beg + slotid < end) {
        myrow = crow[beg + slotid];
        const double *L = Stiles + (size_t)cslot[beg + slotid] * TS * TS + col;
#pragma unroll
        for (int r = 0; r < TS; ++r) lv[r] = __ldg(L + r * TS);
    }
    for (int e = tid; e < TS * TS; e += TS * BWD_SLOTS) invs[e] = __ldg(Linv + (size_t)I * TS * TS + e);
    const double yI = tid < TS ? __ldg(ywork + I * TS + tid) : 0.0;
    BW_STAMP(1);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int t0 = beg; t0 < end; t0 += BWD_SLOTS) {
        // x_J of this slot's tile: every thread polls ITS element of the solution buffer until the producer's value has
        // replaced the sentinel -- the data is its own flag (one L2 round trip after the store instead of flag poll +
        // barrier + coherent load); a bounded spin turns a broken schedule into status 3
        const int t = t0 + slotid;
        if (t < end) {
            const int J = t0 == beg ? myrow : crow[t];
            const double *src = xbuf + J * TS + col;
            unsigned long long bits = ld_relaxed_u64(src);
            int spins = 0;
            while (bits == X_SENTINEL) {
                if (++spins > (1 << 22)) { atomicMax(status, 3); bits = 0; break; }
                bits = ld_relaxed_u64(src);
            }
            xs[slotid][col] = __longlong_as_double((long long)bits);
            if (t0 != beg) {
                const double *L = Stiles + (size_t)cslot[t] * TS * TS + col;
#pragma unroll
                for (int r = 0; r < TS; ++r) lv[r] = __ldg(L + r * TS);
            }
        }
        __syncthreads();
        if (t < end) {
            const double *x = xs[slotid];
#pragma unroll
            for (int r = 0; r < TS; r += 4) { s0 += lv[r] * x[r]; s1 += lv[r + 1] * x[r + 1]; s2 += lv[r + 2] * x[r + 2]; s3 += lv[r + 3] * x[r + 3]; }
        }
        __syncthreads();
    }
    BW_STAMP(3);
    part[slotid][col] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (tid < TS) {
        double a = 0.0;
#pragma unroll
        for (int p = 0; p < BWD_SLOTS; ++p) a += part[p][tid];
        acc[tid] = yI - a;
    }
    __syncthreads();
    if (tid < 4 * TS) {          // x_I[c] = sum_{r >= c} Linv[r][c] acc[r], rows split over 4 partial sums
        const int cidx = tid % TS, q = tid / TS;
        double a = 0.0;
        for (int r = cidx + q; r < TS; r += 4) a += invs[r * TS + cidx] * acc[r];
        mv[q][cidx] = a;
    }
    __syncthreads();
    if (tid < TS) {
        double a = (mv[0][tid] + mv[1][tid]) + (mv[2][tid] + mv[3][tid]);
        if (__double_as_longlong(a) == (long long)X_SENTINEL) a = __longlong_as_double(0x7FF8000000000000ll);   // a NaN solution must not look like "not there yet"
        st_relaxed_f64(xbuf + I * TS + tid, a);
        const int pos = I * TS + tid, cam = pos2cam[pos / 6];
        if (cam >= 0) sol[cam * 6 + pos % 6] = a;
    }
    BW_STAMP(4);
    BW_STAMP(5);
}

void psba_launch_solve(psba_ctx *c)
{
    if (c->chain_schedule) {
        PROF(c, KID_TRI_SOLVE) k_backward<<<1, TS * BW_SLOTS, 0, c->stream>>>(c->nt, c->d_coltile_ptr, c->d_coltile_row, c->d_coltile_slot,
                                                                   c->Stiles, c->Linv, c->chol_diag, c->pos2cam, c->dp);
        c->st_launches += 1;
        LAUNCH_CHECK();
        return;
    }
    static int dbg_runs = getenv("PSBA_BW_DEBUG") ? 3 : 0;
    long long *dbg_dev = nullptr;
    if (dbg_runs > 0) {
        CUDA_CHECK(cudaMalloc(&dbg_dev, (size_t)c->nt * 8 * sizeof(long long)));
        CUDA_CHECK(cudaMemset(dbg_dev, 0, (size_t)c->nt * 8 * sizeof(long long)));
        CUDA_CHECK(cudaMemcpyToSymbol(g_panel_dbg, &dbg_dev, sizeof(dbg_dev)));
    }
    PROF(c, KID_TRI_SOLVE) k_backward_flow<<<c->nt, TS * BWD_SLOTS, 0, c->stream>>>(c->d_bw_order, c->d_coltile_ptr, c->d_coltile_row, c->d_coltile_slot,
                                                                        c->Stiles, c->Linv, c->chol_diag, c->pos2cam, c->dp, c->bw_xbuf, c->d_status);
    c->st_launches += 1;
    LAUNCH_CHECK();
    if (dbg_dev) {
        std::vector<long long> h((size_t)c->nt * 8), ord(c->nt);
        std::vector<int> order(c->nt);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        CUDA_CHECK(cudaMemcpy(h.data(), dbg_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(order.data(), c->d_bw_order, c->nt * sizeof(int), cudaMemcpyDeviceToHost));
        long long *nul = nullptr;
        CUDA_CHECK(cudaMemcpyToSymbol(g_panel_dbg, &nul, sizeof(nul)));
        CUDA_CHECK(cudaFree(dbg_dev));
        if (--dbg_runs == 0) {
            long long t0 = h[(size_t)order[0] * 8];
            for (int b = 0; b < c->nt; ++b) {
                const long long *q = &h[(size_t)order[b] * 8];
                fprintf(stderr, "bw blk %3d panel %3d: start %7lld prefetched %7lld flags %7lld fenced %7lld computed %7lld released %7lld ns\n", b, order[b],
                        q[0] - t0, q[1] - t0, q[2] - t0, q[3] - t0, q[4] - t0, q[5] - t0);
            }
        }
    }
}
