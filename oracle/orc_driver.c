/*
 * orc_driver.c -- ORACLE (test infrastructure): restatement of the optimiser drivers.
 *
 *   orc_levmar        PSBA/levmar.cpp:45-256   (+ compute_rho :271-280)
 *   orc_trust_region  PSBA/trust_region.cpp:49-288, compute_PB :292-405, compute_p_2 :520-595
 *   orc_solve         PSBA/main.cpp:192-209
 *
 * Control flow, constants and branch order are the reference's; the OpenCL plumbing is
 * replaced by calls through the operator table (s->ops).  abs() on doubles
 * (trust_region.cpp:197,252,584) is read as fabs.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "psba_oracle.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

double orc_L2_sq(int n, const double *x)
{   /* misc.cpp:151-157 */
    double sum = 0; int i;
    for (i = 0; i < n; i++) sum += x[i] * x[i];
    return sum;
}

double orc_dot(const double *a, const double *b, int n)
{   /* misc.cpp:162-169 */
    double sum = 0.0; int i;
    for (i = 0; i < n; i++) sum += a[i] * b[i];
    return sum;
}

static void trace(orc_state *s, int phase, double err, double rho, double mu, double delta, double pnorm, int acc)
{
    if (s->ntrace < ORC_TRACE_MAX) {
        orc_trace_rec *r = &s->trace[s->ntrace++];
        r->phase = phase; r->itno = s->itno; r->err = err; r->rho = rho; r->mu = mu;
        r->delta = delta; r->pnorm = pnorm; r->accepted = acc;
    }
}

/* camera solve used by both drivers: dp[0..N) from S and eab[0..N). Returns 0.0 ok. */
static double camera_solve(orc_state *s, int ea_before)
{
    double ret, t0 = now_s();
    if (s->use_explicit_inverse) {
        if (ea_before) s->ops->ea(s);
        ret = orc_SPDinv(s->S, s->diagAux, s->N);                 /* levmar.cpp:134 */
        if (ret == 0.0) {
            if (!ea_before) s->ops->ea(s);                         /* trust_region.cpp:378 */
            s->ops->matVec(s);                                     /* levmar.cpp:139 */
        }
    } else {
        s->ops->ea(s);
        ret = orc_potrf_solve(s->S, s->eab, s->dp, s->N);
    }
    s->t_solve += now_s() - t0;
    return ret;
}

/* levmar.cpp:271-280 */
static double compute_rho(double ex_L2, double new_ex_L2, double mu, int nPara, const double *dp, const double *g)
{
    double sum = 0; int i;
    for (i = 0; i < nPara; i++) sum += dp[i] * (mu * dp[i] + g[i]);
    return (ex_L2 - new_ex_L2) / sum;
}

int orc_levmar(orc_state *s)
{
    const orc_ops *op = s->ops;
    int iter_flag, gooditer_cnt = 0, nu = 2, first = 1, T = s->T, k;
    double tau = ORC_INIT_MU, mu = 0, rho, p_L2 = 0, dp_L2, ex_L2, new_ex_L2, ret, t0;
    double *new_p = (double *)malloc((size_t)T * 8);

    t0 = now_s();
    op->exQT(s, s->cams, s->pts, s->ex); s->n_exqt++;              /* levmar.cpp:93 */
    ex_L2 = orc_L2_sq(s->o * 2, s->ex);
    s->t_cost += now_s() - t0;
    s->initErr = ex_L2;

    iter_flag = ORC_ITER_CONTINUE;
    for (; s->itno < ORC_MAX_ITER && iter_flag == ORC_ITER_CONTINUE; s->itno++) {
        t0 = now_s();
        op->jacobiQT(s);                                          /* levmar.cpp:103-108 */
        op->U(s, 1); op->V(s, 1); op->Wblks(s, 1); op->g(s, 1);
        s->t_lin += now_s() - t0; s->n_lin++;
        if (first) {                                              /* levmar.cpp:114-120, sba_func.cpp:422-444 */
            double mx = 0.0;
            for (k = 0; k < T; k++) if (s->UVdiag[k] > mx) mx = s->UVdiag[k];
            mu = tau * mx; first = 0; p_L2 = 1e+3; nu = 2;
        }
        while (1) {
            s->n_tries++;
            t0 = now_s();
            op->update_UV(s, mu);                                 /* levmar.cpp:126-131 */
            op->Vinv(s); op->Yblks(s); op->S(s);
            s->t_schur += now_s() - t0;
            ret = camera_solve(s, 1);
            if (ret == 0.0) {
                t0 = now_s();
                op->eb(s); op->dpb(s);                            /* levmar.cpp:150-155 */
                s->t_backsub += now_s() - t0;
                dp_L2 = orc_L2_sq(T, s->dp);
                if (dp_L2 < p_L2 * ORC_STOP_THRESH * ORC_STOP_THRESH) { iter_flag = ORC_ITER_DP_NO_CHANGE; break; }
                if (dp_L2 >= (p_L2 + ORC_STOP_THRESH) / (ORC_EPSILON * ORC_EPSILON)) { iter_flag = ORC_ITER_ERR; break; }
                op->restore_UVdiag(s);                            /* levmar.cpp:182 */
                op->newp(s);
                t0 = now_s();
                op->exQT(s, s->newcams, s->newpts, s->ex); s->n_exqt++;
                new_ex_L2 = orc_L2_sq(s->o * 2, s->ex);
                s->t_cost += now_s() - t0;
                rho = compute_rho(ex_L2, new_ex_L2, mu, T, s->dp, s->g);
                if (s->verbose) printf("itno=%d\t\tErr=%.15E\t\trho=%f\t\tmu=%f\n", s->itno, new_ex_L2, rho, mu);
                trace(s, 0, new_ex_L2, rho, mu, 0, sqrt(dp_L2), rho > 0);
                if (rho > 0) {                                    /* levmar.cpp:200-223 */
                    double tmp = 2 * rho - 1;
                    tmp = 1.0 - tmp * tmp * tmp;
                    mu = mu * ((tmp >= (1.0 / 3.0)) ? tmp : (1.0 / 3.0));
                    nu = 2;
                    op->update_p(s);
                    memcpy(new_p, s->cams, (size_t)s->N * 8);
                    memcpy(new_p + s->N, s->pts, (size_t)3 * s->n * 8);
                    p_L2 = orc_L2_sq(T, new_p);
                    ex_L2 = new_ex_L2;
                    if (fabs(rho - 1) < (1.0 / 5.0)) {
                        gooditer_cnt++;
                        if (gooditer_cnt >= 5) { iter_flag = ORC_ITER_TURN_TO_TR; break; }
                    } else gooditer_cnt = 0;
                    break;
                }
            } else {
                gooditer_cnt = 0;
                op->restore_UVdiag(s);                            /* levmar.cpp:229 */
                trace(s, 0, NAN, NAN, mu, 0, 0, 0);
            }
            mu *= nu;                                             /* levmar.cpp:237-244 */
            /* "dtype nu2 = 2*nu; if (nu2 <= nu)": 32-bit wrap-around test of the int nu */
            if (nu >= (1 << 30)) { iter_flag = ORC_ITER_ERR; break; }
            nu = 2 * nu;
        }
        if (ex_L2 <= ORC_STOP_THRESH) iter_flag = ORC_ITER_ERR_SMALL_ENOUGH;   /* levmar.cpp:247 */
    }
    s->finalErr = ex_L2;
    free(new_p);
    return iter_flag;
}

/* trust_region.cpp:292-405 */
static int compute_PB(orc_state *s, double *lambda, double *P_B)
{
    const orc_ops *op = s->ops;
    int N = s->N, T = s->T, i;
    double ret, t0 = now_s();
    s->n_tries++;
    op->update_UV(s, *lambda);
    op->Vinv(s); op->Yblks(s); op->S(s);
    s->t_schur += now_s() - t0;
    memcpy(s->Saux, s->S, (size_t)N * N * 8);                     /* trust_region.cpp:330-332 */
    ret = camera_solve(s, 0);
    if (ret != 0.0) {
        if (*lambda == 0.0) {
            double delta, beta, sum = 0.0;
            int nscalar = 0;
            memcpy(s->S, s->Saux, (size_t)N * N * 8);
            orc_get_delta_beta(s->S, N, &delta, &beta);           /* cl_cholmod.cpp:39 */
            orc_cholmod_blk(s->S, s->blkBackup, s->diagAux, s->E, N, beta, delta, &nscalar);
            orc_cholmod_E(s->S, s->E, N);
            for (i = 0; i < N; i++) sum += s->E[i];               /* trust_region.cpp:358-364 */
            *lambda = fabs(sum) / N;
            if (s->n_cholmod_events < s->n_force_lambda) *lambda = s->force_lambda[s->n_cholmod_events];
            s->n_cholmod_events++;
            trace(s, 2, (double)nscalar, sum, *lambda, 0, 0, 0);
            return 0;
        } else {
            *lambda = 2 * (*lambda);
            return 0;
        }
    }
    t0 = now_s();
    op->eb(s); op->dpb(s);                                        /* trust_region.cpp:389-392 */
    s->t_backsub += now_s() - t0;
    for (i = 0; i < T; i++) P_B[i] = -s->dp[i];
    return 1;
}

/* trust_region.cpp:520-595 */
static double compute_p_2(int nT, double pUtBpU, double pUtBpB, double pBtBpB, double delta,
                          const double *P_U, const double *P_B, double *p, const double *g)
{
    double pUg, pBg, eta1, eta2, p_norm, pU_norm, pB_norm;
    int i;
    pUg = orc_dot(P_U, g, nT);
    pBg = orc_dot(P_B, g, nT);
    eta1 = (pBg * pUtBpB) / (-pUtBpB * pUtBpB + pBtBpB * pUtBpU) - (pBtBpB * pUg) / (-pUtBpB * pUtBpB + pBtBpB * pUtBpU);
    eta2 = (pUg * pUtBpB) / (-pUtBpB * pUtBpB + pBtBpB * pUtBpU) - (pBg * pUtBpU) / (-pUtBpB * pUtBpB + pBtBpB * pUtBpU);
    p_norm = 0.0;
    for (i = 0; i < nT; i++) { p[i] = eta1 * P_U[i] + eta2 * P_B[i]; p_norm += p[i] * p[i]; }
    p_norm = sqrt(p_norm);
    if (p_norm > delta) {
        pU_norm = 0.0; pB_norm = 0.0;
        for (i = 0; i < nT; i++) { pU_norm += P_U[i] * P_U[i]; pB_norm += P_B[i] * P_B[i]; }
        pU_norm = sqrt(pU_norm); pB_norm = sqrt(pB_norm);
        if (pU_norm > delta) {
            for (i = 0; i < nT; i++) p[i] = delta * P_U[i] / pU_norm;
            return delta;
        } else if (pB_norm <= delta) {
            for (i = 0; i < nT; i++) { p[i] = P_B[i]; p_norm += p[i] * p[i]; }   /* SURVEY A.5(10) */
            return sqrt(p_norm);
        } else {
            double a = 0.0, b = 0.0, c = 0.0, Ai, Bi, b2_4ac, tau;
            for (i = 0; i < nT; i++) {
                Ai = P_B[i] - P_U[i]; Bi = 2 * P_U[i] - P_B[i];
                a += Ai * Ai; b += Ai * Bi; c += Bi * Bi;
            }
            b = 2 * b; c = c - delta * delta;
            b2_4ac = b * b - 4 * a * c; if (fabs(b2_4ac) < 1e-12) b2_4ac = 0;
            tau = (-b + sqrt(b2_4ac)) / (2 * a);
            for (i = 0; i < nT; i++) p[i] = P_U[i] + (tau - 1) * (P_B[i] - P_U[i]);
            return delta;
        }
    }
    return p_norm;
}

int orc_trust_region(orc_state *s)
{
    const orc_ops *op = s->ops;
    int iter_flag, solved, notgood_cnt = 0, good_iters = 0, nu = 2, T = s->T, o2 = s->o * 2, i;
    double ex_L2, pred_ex_L2, act_ex_L2, gtBg, gtg, dk = 1, lambda = 0, p_norm, origin_lambda = 0.0, t0;
    double *P_U = (double *)malloc((size_t)T * 8), *P_B = (double *)malloc((size_t)T * 8), *P = (double *)malloc((size_t)T * 8);

    t0 = now_s();
    op->exQT(s, s->cams, s->pts, s->ex); s->n_exqt++;             /* trust_region.cpp:106-107 */
    ex_L2 = orc_L2_sq(o2, s->ex);
    s->t_cost += now_s() - t0;
    iter_flag = ORC_ITER_CONTINUE;
    for (; s->itno < ORC_MAX_ITER; s->itno++) {
        t0 = now_s();
        op->jacobiQT(s);                                          /* trust_region.cpp:117 */
        op->g(s, -2);                                             /* :122 */
        op->Jmultiply(s, s->g, s->Jx1);                           /* :125 */
        gtBg = 2 * orc_dot(s->Jx1, s->Jx1, o2);
        gtg = orc_dot(s->g, s->g, T);
        for (i = 0; i < T; i++) P_U[i] = -(s->g[i] * gtg) / gtBg;
        op->U(s, 2); op->V(s, 2); op->Wblks(s, 2);                /* :133-137 */
        s->t_lin += now_s() - t0; s->n_lin++;
        solved = 0;
        while (!solved) {                                         /* :141-163 */
            solved = compute_PB(s, &lambda, P_B);
            if (!solved) {
                if (origin_lambda != 0.0) {
                    if (nu > 4) {
                        s->finalErr = ex_L2;
                        free(P_U); free(P_B); free(P);
                        return ORC_ITER_TURN_TO_LM;
                    } else {
                        lambda = lambda * nu; nu = nu * 2;
                        op->restore_UVdiag(s);
                    }
                } else op->restore_UVdiag(s);
            } else { nu = 2; origin_lambda = lambda; }
        }
        {
            double pUtBpU, pUtBpB, pBtBpB;
            memcpy(s->dp, P_U, (size_t)T * 8); op->Jmultiply(s, s->dp, s->Jx1);   /* :166-171 */
            memcpy(s->dp, P_B, (size_t)T * 8); op->Jmultiply(s, s->dp, s->Jx2);
            pUtBpU = 2 * orc_dot(s->Jx1, s->Jx1, o2);
            pUtBpB = 2 * orc_dot(s->Jx1, s->Jx2, o2);
            pBtBpB = 2 * orc_dot(s->Jx2, s->Jx2, o2);

            iter_flag = ORC_ITER_CONTINUE;
            while (iter_flag == ORC_ITER_CONTINUE) {              /* :180 */
                double Jx_norm, rho;
                int acc = 0;
                p_norm = compute_p_2(T, pUtBpU, pUtBpB, pBtBpB, dk, P_U, P_B, P, s->g);
                memcpy(s->dp, P, (size_t)T * 8);
                op->newp(s);
                t0 = now_s();
                op->exQT(s, s->newcams, s->newpts, s->ex); s->n_exqt++;
                act_ex_L2 = orc_L2_sq(o2, s->ex);
                s->t_cost += now_s() - t0;
                if (fabs((ex_L2 - act_ex_L2) / ex_L2) < ORC_EPSILON2) { iter_flag = ORC_ITER_DP_NO_CHANGE; break; }
                op->Jmultiply(s, s->dp, s->Jx1);                  /* :208-212 */
                Jx_norm = 2 * orc_L2_sq(o2, s->Jx1);
                pred_ex_L2 = orc_dot(s->g, P, T);
                pred_ex_L2 += ex_L2 + Jx_norm / 2;
                rho = (ex_L2 - act_ex_L2) / (ex_L2 - pred_ex_L2);
                if (rho < (1.0 / 4.0) || act_ex_L2 > ex_L2) {
                    dk = dk / 4;
                } else if (rho >= (3.0 / 4.0) && act_ex_L2 < ex_L2) {
                    iter_flag = ORC_ITER_PASS; acc = 1;
                    op->update_p(s);
                    s->finalErr = act_ex_L2;
                    dk = fmin(2 * dk, ORC_MAX_DELTA);
                } else if (rho >= (1.0 / 4.0) && rho < (3.0 / 4.0) && act_ex_L2 < ex_L2) {
                    iter_flag = ORC_ITER_PASS; acc = 1;
                    op->update_p(s);
                    s->finalErr = act_ex_L2;
                } else if (isnan(rho)) {
                    s->finalErr = ex_L2;
                    free(P_U); free(P_B); free(P);
                    return ORC_ITER_TURN_TO_LM;
                }
                if (s->verbose)
                    printf("itno=%d\tErr:%.15E\tDelta=%f\tRho=%f\tnorm_p=%f\tLambda=%E\n", s->itno, act_ex_L2, dk, rho, p_norm, lambda);
                trace(s, 1, act_ex_L2, rho, lambda, dk, p_norm, acc);
                if (fabs((act_ex_L2 - ex_L2) / ex_L2) <= ORC_EPSILON2) { iter_flag = ORC_ITER_ERR_SMALL_ENOUGH; break; }
                if (rho < 1.0 / 4) {
                    notgood_cnt++;
                    if (notgood_cnt >= 5) { iter_flag = ORC_ITER_TURN_TO_LM; break; }
                } else notgood_cnt = 0;
                if (rho > 3.0 / 4 && act_ex_L2 < ex_L2) {
                    good_iters++;
                    if (good_iters >= 10) { lambda = 0.0; origin_lambda = 0.0; good_iters = 0; }
                } else good_iters = 0;
                if (rho > (1.0 / 4) && act_ex_L2 < ex_L2) ex_L2 = act_ex_L2;
            }
        }
        if (iter_flag != ORC_ITER_PASS) break;
    }
    free(P_U); free(P_B); free(P);
    return iter_flag;
}

/* main.cpp:192-209 */
int orc_solve(orc_state *s)
{
    int flag;
    double t0 = now_s();
    s->itno = 0; s->ntrace = 0; s->n_cholmod_events = 0;
    while (1) {
        flag = orc_levmar(s);
        if (flag != ORC_ITER_TURN_TO_TR) break;
        flag = orc_trust_region(s);
        if (flag != ORC_ITER_TURN_TO_LM) break;
    }
    s->t_total = now_s() - t0;
    return flag;
}
