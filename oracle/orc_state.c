/*
 * orc_state.c -- ORACLE (test infrastructure): problem state and index structures.
 *
 * Buffer list follows setup_cl (PSBA/cl_psba.cpp:40-84).  The index structures enumerate
 * the same entries in the same order as generate_idxs (PSBA/misc.cpp:178-218):
 *   cam_obs  : for camera j, observation ids in ascending point order (= the order in
 *              which the reference scans blk_idx[i*m+j], i = 0..n-1, compute_U.cl:22-29);
 *   pair_*   : for camera pair (k,l), the common points in ascending order (= comm3DIdx
 *              rows, misc.cpp:199-209) with the observation ids blk_idx[i*m+k], blk_idx[i*m+l].
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "psba_oracle.h"

static void *xcalloc(size_t n, size_t sz)
{
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "oracle: allocation of %zu x %zu bytes failed\n", n, sz); exit(1); }
    return p;
}

orc_state *orc_create(int m, int n, int o, const double *K, const double *impts,
                      const double *initcams, const double *camsEx, const double *pts,
                      const int *iidx, const int *jidx, int want_dense)
{
    orc_state *s = (orc_state *)xcalloc(1, sizeof(orc_state));
    int i, j, k, a, b;
    long long t;
    s->m = m; s->n = n; s->o = o; s->N = 6 * m; s->T = 6 * m + 3 * n;
    s->K = (double *)xcalloc((size_t)m * 5, 8);       memcpy(s->K, K, (size_t)m * 5 * 8);
    s->impts = (double *)xcalloc((size_t)o * 2, 8);   memcpy(s->impts, impts, (size_t)o * 2 * 8);
    s->initcams = (double *)xcalloc((size_t)m * 4, 8); memcpy(s->initcams, initcams, (size_t)m * 4 * 8);
    s->cams = (double *)xcalloc((size_t)m * 6, 8);    memcpy(s->cams, camsEx, (size_t)m * 6 * 8);
    s->pts = (double *)xcalloc((size_t)n * 3, 8);     memcpy(s->pts, pts, (size_t)n * 3 * 8);
    s->newcams = (double *)xcalloc((size_t)m * 6, 8);
    s->newpts = (double *)xcalloc((size_t)n * 3, 8);
    s->iidx = (int *)xcalloc(o, 4); memcpy(s->iidx, iidx, (size_t)o * 4);
    s->jidx = (int *)xcalloc(o, 4); memcpy(s->jidx, jidx, (size_t)o * 4);

    /* CSR by point */
    s->pt_ptr = (int *)xcalloc((size_t)n + 1, 4);
    for (k = 0; k < o; ++k) s->pt_ptr[iidx[k] + 1]++;
    for (i = 0; i < n; ++i) s->pt_ptr[i + 1] += s->pt_ptr[i];
    /* camera-major lists, ascending point (observations are already point-major) */
    s->cam_ptr = (int *)xcalloc((size_t)m + 1, 4);
    s->cam_obs = (int *)xcalloc(o, 4);
    for (k = 0; k < o; ++k) s->cam_ptr[jidx[k] + 1]++;
    for (j = 0; j < m; ++j) s->cam_ptr[j + 1] += s->cam_ptr[j];
    {
        int *fill = (int *)xcalloc(m, 4);
        for (k = 0; k < o; ++k) { j = jidx[k]; s->cam_obs[s->cam_ptr[j] + fill[j]++] = k; }
        free(fill);
    }
    /* camera-pair triples (both orders (k,l) and (l,k), as comm3DIdx is filled symmetrically) */
    s->pair_ptr = (long long *)xcalloc((size_t)m * m + 1, sizeof(long long));
    t = 0;
    for (i = 0; i < n; ++i) {
        long long d = s->pt_ptr[i + 1] - s->pt_ptr[i];
        t += d * d;
        for (a = s->pt_ptr[i]; a < s->pt_ptr[i + 1]; ++a)
            for (b = s->pt_ptr[i]; b < s->pt_ptr[i + 1]; ++b)
                s->pair_ptr[(size_t)jidx[a] * m + jidx[b] + 1]++;
    }
    s->ntriples = t;
    for (t = 0; t < (long long)m * m; ++t) s->pair_ptr[t + 1] += s->pair_ptr[t];
    s->pair_oa = (int *)xcalloc((size_t)s->ntriples, 4);
    s->pair_ob = (int *)xcalloc((size_t)s->ntriples, 4);
    {
        long long *fill = (long long *)xcalloc((size_t)m * m, sizeof(long long));
        for (i = 0; i < n; ++i)
            for (a = s->pt_ptr[i]; a < s->pt_ptr[i + 1]; ++a)
                for (b = s->pt_ptr[i]; b < s->pt_ptr[i + 1]; ++b) {
                    size_t p = (size_t)jidx[a] * m + jidx[b];
                    long long at = s->pair_ptr[p] + fill[p]++;
                    s->pair_oa[at] = a; s->pair_ob[at] = b;
                }
        free(fill);
    }
    if (want_dense) {   /* the reference's own tables, misc.cpp:178-218 (restated from the CSR lists) */
        size_t rowsize = (size_t)m * n;
        s->blk_idx = (int *)xcalloc((size_t)n * m, 4);
        s->comm3DIdx = (int *)xcalloc((size_t)n * m * m, 4);
        s->comm3DIdxCnt = (int *)xcalloc((size_t)m * m, 4);
        for (t = 0; t < (long long)n * m; ++t) s->blk_idx[t] = -1;
        for (k = 0; k < o; ++k) s->blk_idx[(size_t)iidx[k] * m + jidx[k]] = k;
        for (a = 0; a < m; ++a)
            for (b = 0; b < m; ++b) {
                size_t p = (size_t)a * m + b;
                long long q;
                s->comm3DIdxCnt[p] = (int)(s->pair_ptr[p + 1] - s->pair_ptr[p]);
                for (q = s->pair_ptr[p]; q < s->pair_ptr[p + 1]; ++q)
                    s->comm3DIdx[(size_t)a * rowsize + (size_t)b * n + (q - s->pair_ptr[p])] = iidx[s->pair_oa[q]];
            }
    }
    s->ex = (double *)xcalloc((size_t)o * 2, 8);
    s->JA = (double *)xcalloc((size_t)o * 12, 8);
    s->JB = (double *)xcalloc((size_t)o * 6, 8);
    s->U = (double *)xcalloc((size_t)m * 36, 8);
    s->V = (double *)xcalloc((size_t)n * 9, 8);
    s->UVdiag = (double *)xcalloc(s->T, 8);
    s->W = (double *)xcalloc((size_t)o * 18, 8);
    s->Y = (double *)xcalloc((size_t)o * 18, 8);
    s->S = (double *)xcalloc((size_t)s->N * s->N, 8);
    s->Saux = (double *)xcalloc((size_t)s->N * s->N, 8);
    s->diagAux = (double *)xcalloc((size_t)3 * s->N, 8);     /* 18m doubles, cl_psba.cpp:70-71 */
    s->blkBackup = (double *)xcalloc((size_t)3 * s->N, 8);
    s->E = (double *)xcalloc(s->N, 8);
    s->g = (double *)xcalloc(s->T, 8);
    s->dp = (double *)xcalloc(s->T, 8);
    s->eab = (double *)xcalloc(s->T, 8);
    s->Jx1 = (double *)xcalloc((size_t)o * 2, 8);
    s->Jx2 = (double *)xcalloc((size_t)o * 2, 8);
    s->use_explicit_inverse = 1;
    s->nthreads = 1;
    s->ops = orc_native_ops();
    return s;
}

void orc_use_ops(orc_state *s, const orc_ops *ops) { s->ops = ops; }

void orc_set_ext(orc_state *s, const double *kc, const double *wgt)
{
    free(s->kc); free(s->wgt); s->kc = s->wgt = NULL;
    if (kc) { s->kc = (double *)xcalloc((size_t)s->m * 5, 8); memcpy(s->kc, kc, (size_t)s->m * 5 * 8); }
    if (wgt) { s->wgt = (double *)xcalloc((size_t)s->o * 3, 8); memcpy(s->wgt, wgt, (size_t)s->o * 3 * 8); }
}

void orc_destroy(orc_state *s)
{
    if (!s) return;
    free(s->kc); free(s->wgt);
    free(s->K); free(s->impts); free(s->initcams); free(s->cams); free(s->newcams);
    free(s->pts); free(s->newpts); free(s->iidx); free(s->jidx); free(s->pt_ptr);
    free(s->cam_ptr); free(s->cam_obs); free(s->pair_ptr); free(s->pair_oa); free(s->pair_ob);
    free(s->blk_idx); free(s->comm3DIdx); free(s->comm3DIdxCnt);
    free(s->ex); free(s->JA); free(s->JB); free(s->U); free(s->V); free(s->UVdiag);
    free(s->W); free(s->Y); free(s->S); free(s->Saux); free(s->diagAux); free(s->blkBackup);
    free(s->E); free(s->g); free(s->dp); free(s->eab); free(s->Jx1); free(s->Jx2);
    free(s);
}
