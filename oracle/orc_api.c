/*
 * orc_api.c -- ORACLE (test infrastructure): name-based accessors so that the Python
 * harness (oracle/__init__.py, ctypes) does not have to mirror the struct layout.
 */
#include <string.h>
#include <stdio.h>
#include "psba_oracle.h"

#define D(name) if (!strcmp(n, #name)) return s->name
double *orc_ptr_d(orc_state *s, const char *n)
{
    D(K); D(impts); D(initcams); D(cams); D(newcams); D(pts); D(newpts);
    D(ex); D(JA); D(JB); D(U); D(V); D(UVdiag); D(W); D(Y); D(S); D(Saux); D(diagAux);
    D(blkBackup); D(E); D(g); D(dp); D(eab); D(Jx1); D(Jx2);
    fprintf(stderr, "orc_ptr_d: unknown buffer %s\n", n);
    return 0;
}
int *orc_ptr_i(orc_state *s, const char *n)
{
    D(iidx); D(jidx); D(pt_ptr); D(cam_ptr); D(cam_obs); D(pair_oa); D(pair_ob);
    D(blk_idx); D(comm3DIdx); D(comm3DIdxCnt);
    fprintf(stderr, "orc_ptr_i: unknown buffer %s\n", n);
    return 0;
}
long long *orc_pair_ptr(orc_state *s) { return s->pair_ptr; }
#undef D
#define G(name) if (!strcmp(n, #name)) return (double)s->name
double orc_get(orc_state *s, const char *n)
{
    G(m); G(n); G(o); G(N); G(T); G(ntriples); G(ret); G(itno); G(initErr); G(finalErr);
    G(use_explicit_inverse); G(n_cholmod_events); G(ntrace); G(nthreads);
    G(t_total); G(t_lin); G(t_schur); G(t_solve); G(t_backsub); G(t_cost);
    G(n_tries); G(n_exqt); G(n_lin);
    fprintf(stderr, "orc_get: unknown field %s\n", n);
    return -1;
}
#undef G
void orc_set(orc_state *s, const char *n, double v)
{
    if (!strcmp(n, "use_explicit_inverse")) s->use_explicit_inverse = (int)v;
    else if (!strcmp(n, "verbose")) s->verbose = (int)v;
    else if (!strcmp(n, "nthreads")) s->nthreads = (int)v;
    else if (!strcmp(n, "itno")) s->itno = (int)v;
    else fprintf(stderr, "orc_set: unknown field %s\n", n);
}
void orc_force_lambda(orc_state *s, const double *lam, int n)
{
    int i;
    if (n > 64) n = 64;
    for (i = 0; i < n; ++i) s->force_lambda[i] = lam[i];
    s->n_force_lambda = n;
}
void orc_trace_get(orc_state *s, int k, double *out8)
{
    const orc_trace_rec *r = &s->trace[k];
    out8[0] = r->phase; out8[1] = r->itno; out8[2] = r->err; out8[3] = r->rho;
    out8[4] = r->mu; out8[5] = r->delta; out8[6] = r->pnorm; out8[7] = r->accepted;
}
/* call one operator by name: the reference's wrappers of PSBA/sba_func.h */
double orc_call(orc_state *s, const char *n, double a)
{
    const orc_ops *op = s->ops;
    if (!strcmp(n, "exQT")) { op->exQT(s, s->cams, s->pts, s->ex); return orc_L2_sq(2 * s->o, s->ex); }
    if (!strcmp(n, "exQT_new")) { op->exQT(s, s->newcams, s->newpts, s->ex); return orc_L2_sq(2 * s->o, s->ex); }
    if (!strcmp(n, "jacobiQT")) { op->jacobiQT(s); return 0; }
    if (!strcmp(n, "U")) { op->U(s, a); return 0; }
    if (!strcmp(n, "V")) { op->V(s, a); return 0; }
    if (!strcmp(n, "Wblks")) { op->Wblks(s, a); return 0; }
    if (!strcmp(n, "g")) { op->g(s, a); return 0; }
    if (!strcmp(n, "update_UV")) { op->update_UV(s, a); return 0; }
    if (!strcmp(n, "restore_UVdiag")) { op->restore_UVdiag(s); return 0; }
    if (!strcmp(n, "Vinv")) return op->Vinv(s);
    if (!strcmp(n, "Yblks")) { op->Yblks(s); return 0; }
    if (!strcmp(n, "S")) { op->S(s); return 0; }
    if (!strcmp(n, "ea")) { op->ea(s); return 0; }
    if (!strcmp(n, "SPDinv")) return orc_SPDinv(s->S, s->diagAux, s->N);
    if (!strcmp(n, "potrf_solve")) return orc_potrf_solve(s->S, s->eab, s->dp, s->N);
    if (!strcmp(n, "matVec")) { op->matVec(s); return 0; }
    if (!strcmp(n, "eb")) { op->eb(s); return 0; }
    if (!strcmp(n, "dpb")) { op->dpb(s); return 0; }
    if (!strcmp(n, "newp")) { op->newp(s); return 0; }
    if (!strcmp(n, "update_p")) { op->update_p(s); return 0; }
    if (!strcmp(n, "Jmultiply_g")) { op->Jmultiply(s, s->g, s->Jx1); return orc_dot(s->Jx1, s->Jx1, 2 * s->o); }
    if (!strcmp(n, "Jmultiply_dp")) { op->Jmultiply(s, s->dp, s->Jx1); return orc_dot(s->Jx1, s->Jx1, 2 * s->o); }
    if (!strcmp(n, "cholmod")) {
        double delta, beta, sum = 0; int ns = 0, i;
        orc_get_delta_beta(s->S, s->N, &delta, &beta);
        orc_cholmod_blk(s->S, s->blkBackup, s->diagAux, s->E, s->N, beta, delta, &ns);
        orc_cholmod_E(s->S, s->E, s->N);
        for (i = 0; i < s->N; ++i) sum += s->E[i];
        s->ret = ns;
        return sum;
    }
    fprintf(stderr, "orc_call: unknown operator %s\n", n);
    return -1;
}
