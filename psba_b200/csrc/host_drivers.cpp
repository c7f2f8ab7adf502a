// host_drivers.cpp -- the optimiser drivers of PSBA re-expressed over the fused device-resident
// steps.  Same constants, same branch order, same accept / reject rules as
//   levmar        PSBA/levmar.cpp:45-256 (compute_rho :271-280)
//   trust_region  PSBA/trust_region.cpp:49-288 (compute_PB :292-405, compute_p_2 :520-595)
//   main loop     PSBA/main.cpp:192-209
// What changes: nothing but a handful of scalars crosses PCIe per damping / radius try (the
// reference reads back ex (16 B/obs), dp and the dense 2mn Jx vectors every time); all vectors
// (g, dp, P_U, P_B, P) live on the device.  abs() on doubles is read as fabs.
#include "psba_internal.h"
#include <cmath>
#include <cstring>

#define PSBA_INIT_MU      1e-03      // PSBA/psba.h:6
#define PSBA_STOP_THRESH  1e-12      // PSBA/psba.h:7
#define PSBA_EPSILON      1e-12      // PSBA/psba.h:8
#define PSBA_EPSILON2     1e-12      // PSBA/psba.h:9
#define PSBA_MAX_DELTA    10000.0    // PSBA/trust_region.cpp:18

static void trace(psba_ctx *c, int phase, double err, double rho, double mu, double delta, double pnorm, int acc)
{
    psba_trace_rec r;
    r.phase = phase; r.itno = c->itno; r.err = err; r.rho = rho; r.mu = mu; r.delta = delta; r.pnorm = pnorm; r.accepted = acc;
    c->trace.push_back(r);
}


extern "C" int psba_levmar(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs, double *finalErr)
{
    (void)cnp; (void)pnp; (void)mnp; (void)n3Dpts; (void)nCams; (void)n2Dprojs;
    int iter_flag, gooditer_cnt = 0, nu = 2;
    bool first = true;
    const double tau = PSBA_INIT_MU;
    double mu = 0.0, rho, p_L2 = 0.0, ex_L2, new_ex_L2;

    ex_L2 = psba_launch_cost(c, c->cur, nullptr);                 // levmar.cpp:93-95
    c->initErr = ex_L2;
    iter_flag = PSBA_ITER_CONTINUE;
    for (; c->itno < c->max_iter && iter_flag == PSBA_ITER_CONTINUE; c->itno++) {
        if (psba_seq_begin(c, SEQ_LIN_LM, 0.0, 0.0, 0.0)) { psba_launch_linearize(c, 1.0, 1.0); psba_seq_end(c); }   // levmar.cpp:103-108
        if (first) {                                              // levmar.cpp:114-120
            mu = tau * psba_launch_maxdiag(c);
            first = false; p_L2 = 1e+3; nu = 2;
        }
        while (1) {
            psba_try_result res;
            psba_try_step(c, mu, &res);                           // levmar.cpp:126-155, 182-193
            if (res.solve_status == 0.0) {
                const double dp_L2 = res.dp_L2;
                if (dp_L2 < p_L2 * PSBA_STOP_THRESH * PSBA_STOP_THRESH) { iter_flag = PSBA_ITER_DP_NO_CHANGE; break; }
                if (dp_L2 >= (p_L2 + PSBA_STOP_THRESH) / (PSBA_EPSILON * PSBA_EPSILON)) {
                    printf("the matrix of the augmented normal equations is almost singular\n");
                    iter_flag = PSBA_ITER_ERR; break;
                }
                new_ex_L2 = res.cost_new;
                rho = (ex_L2 - new_ex_L2) / res.dp_dot;           // levmar.cpp:271-280
                if (c->verbose) printf("itno=%d\t\tErr=%.15E\t\trho=%f\t\tmu=%f\n", c->itno, new_ex_L2, rho, mu);
                trace(c, 0, new_ex_L2, rho, mu, 0, sqrt(dp_L2), rho > 0);
                if (rho > 0) {                                    // levmar.cpp:200-223
                    double tmp = 2 * rho - 1;
                    tmp = 1.0 - tmp * tmp * tmp;
                    mu = mu * ((tmp >= (1.0 / 3.0)) ? tmp : (1.0 / 3.0));
                    nu = 2;
                    c->cur = 1 - c->cur;                          // update_p: pointer swap
                    c->lin_valid = false;
                    p_L2 = res.p_new_L2;                          // fused into the back-substitution pass
                    ex_L2 = new_ex_L2;
                    if (fabs(rho - 1) < (1.0 / 5.0)) {
                        gooditer_cnt++;
                        // lm_only (bench / tests): keep iterating in LM instead of handing over to TR
                        if (gooditer_cnt >= 5 && !c->lm_only) { iter_flag = PSBA_ITER_TURN_TO_TR; break; }
                    } else gooditer_cnt = 0;
                    break;
                }
            } else {
                gooditer_cnt = 0;
                trace(c, 0, NAN, NAN, mu, 0, 0, 0);
            }
            mu *= nu;                                             // levmar.cpp:237-244
            if (nu >= (1 << 30)) { printf("too many failed attempts to increase the damping factor.\n"); iter_flag = PSBA_ITER_ERR; break; }
            nu = 2 * nu;
        }
        if (ex_L2 <= PSBA_STOP_THRESH) iter_flag = PSBA_ITER_ERR_SMALL_ENOUGH;   // levmar.cpp:247
    }
    *finalErr = ex_L2;
    return iter_flag;
}

// trust_region.cpp:292-405.  Returns true when P_B = -(B + lambda I)^-1 g is available in c->P_B.
static bool compute_PB(psba_ctx *c, double *lambda)
{
    c->st_tries += 1;
    psba_launch_schur(c, *lambda);                                // update_UV, Vinv, Yblks, S (+ ea)
    double ret;
    if (c->camera_solver == 1) {                                  // optional PCG camera solve: "not positive definite / no convergence" = failed
        psba_launch_pcg(c);
        int st = 0;
        CUDA_CHECK(cudaMemcpyAsync(&st, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        ret = st ? 1.0 : 0.0;
    } else ret = psba_launch_factor(c);
    if (ret != 0.0) {
        if (*lambda == 0.0) {
            // the failed factorisation overwrote the tile pool: rebuild S (the reference restores
            // it from Saux, trust_region.cpp:345-346), then the modified Cholesky picks lambda
            psba_launch_schur(c, 0.0);
            double delta, beta, sum; int nscalar = 0;
            bool dense = !psba_cholmod_use_tiles(c);
            if (!dense) {
                double ratio = 0.0;
                sum = psba_launch_cholmod_tiles(c, &delta, &beta, &nscalar, nullptr, &ratio);
                const char *e = getenv("PSBA_CHOLMOD_TILES");
                if (ratio > 1.0 && psba_cholmod_dense_possible(c) && !(e && atoi(e))) { dense = true; psba_launch_schur(c, 0.0); }   // the `> beta` rescue lives in the dense kernel
            }
            if (dense) {
                const size_t nn = (size_t)c->N * c->N;
                if (!c->Sdense) c->Sdense = (double *)psba_dev_alloc(c, nn * sizeof(double), true);
                psba_tiles_to_dense(c, c->Sdense, true);
                sum = psba_launch_cholmod(c, &delta, &beta, &nscalar);
            }
            *lambda = fabs(sum) / c->N;                           // trust_region.cpp:358-364
            if ((size_t)c->n_cholmod_events < c->force_lambda.size()) *lambda = c->force_lambda[c->n_cholmod_events];
            c->n_cholmod_events++;
            trace(c, 2, (double)nscalar, sum, *lambda, 0, 0, 0);
            return false;
        }
        *lambda = 2 * (*lambda);
        return false;
    }
    if (c->camera_solver != 1) psba_launch_solve(c);              // dpa
    {   // the dataflow backward solve reports a broken schedule (bounded spin) through the status word
        int st = 0;
        CUDA_CHECK(cudaMemcpyAsync(&st, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (st != 0) { fprintf(stderr, "psba_b200: camera solve failed with status %d after a successful factorisation\n", st); exit(EXIT_FAILURE); }
    }
    psba_launch_backsub(c, *lambda, false, nullptr);              // eb, dpb
    psba_launch_axpby(c, -1.0, c->dp, 0.0, c->dp, c->P_B);        // P_B = -dp
    return true;
}


// ---- trust region with ONE device round trip per Gauss-Newton step and ONE per radius try -----------------------------
// Same iteration as psba_trust_region_explicit below (trust_region.cpp:112-272, compute_p_2 :520-595); what changes is where the
// scalars come from.  P_U = -(g gtg)/gtBg is a multiple of g, P = eta1 P_U + eta2 P_B (or one of the three dog-leg forms) is a
// combination of g and P_B, and every scalar the reference forms from those vectors -- pUpU, pUg, pBpB, pBg, |P|, the
// dog-leg quadratic, g.P, |J P|^2 -- is a function of the six numbers  g.g, g.P_B, P_B.P_B, |J g|^2, Jg.JP_B, |J P_B|^2.
// They are computed ONCE per step, behind the camera solve and without a host round trip in between (the status word
// of the solve travels with them); a radius try is then  dp = a g + b P_B, candidate parameters, candidate cost.
// The reference needs 4 J-products, 5-6 triple dot products and 3-6 vector updates with a read-back each per step; here:
// 1 J-product, 1 triple dot product, 1 vector update per radius try, 2 read-backs.  Values agree with the explicit
// evaluation to rounding (the explicit path stays available: psba_set_option "tr_fused" = 0).
static int enqueue_PB(psba_ctx *c, double lambda)
{
    c->st_tries += 1;
    psba_launch_schur(c, lambda);
    psba_launch_factor(c, true);                                  // status word read later, with the scalars
    psba_launch_solve(c);
    psba_launch_backsub(c, lambda, false, nullptr);               // eb, dpb
    psba_launch_axpby(c, -1.0, c->dp, 0.0, c->dp, c->P_B);        // P_B = -dp
    return 0;
}

static int psba_trust_region_fused(psba_ctx *c, double *finalErr)
{
    int iter_flag, notgood_cnt = 0, good_iters = 0, nu = 2;
    double ex_L2, pred_ex_L2, act_ex_L2, dk = 1, lambda = 0, p_norm, origin_lambda = 0.0;

    ex_L2 = psba_launch_cost(c, c->cur, nullptr);                 // trust_region.cpp:106-107
    iter_flag = PSBA_ITER_CONTINUE;
    for (; c->itno < c->max_iter; c->itno++) {
        if (psba_seq_begin(c, SEQ_LIN_TR, 0.0, 0.0, 0.0)) { psba_launch_linearize(c, 2.0, -2.0); psba_seq_end(c); }   // :117-122, 133-137
        double JgJg = 0, JgJB = 0, JBJB = 0, gg = 0, gB = 0, BB = 0;
        bool solved = false;
        while (!solved) {                                         // :141-163 with compute_PB (:292-405) inlined
            if (psba_seq_begin(c, SEQ_TR_STEP, lambda, 0.0, 0.0)) {   // one chain; a CUDA graph from its third use on (small problems)
                enqueue_PB(c, lambda);
                psba_enqueue_Jdot(c, c->g, c->P_B, nullptr, 16);  // |J g|^2, Jg.JP_B, |J P_B|^2
                psba_enqueue_dots(c, c->g, c->P_B, c->g, 20);     // g.g, g.P_B, ., P_B.P_B (camera part, point part)
                CUDA_CHECK(cudaMemcpyAsync(c->h_scal + 16, c->d_scal + 16, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                CUDA_CHECK(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                psba_seq_end(c);
            }
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            const int st = c->h_status[0];
            if (st > 1) { fprintf(stderr, "psba_b200: camera solve failed with status %d (broken dataflow schedule)\n", st); exit(EXIT_FAILURE); }
            if (st == 0) {
                c->factor_valid = true;
                JgJg = c->h_scal[16]; JgJB = c->h_scal[17]; JBJB = c->h_scal[18];
                gg = c->h_scal[20] + c->h_scal[26]; gB = c->h_scal[21] + c->h_scal[27]; BB = c->h_scal[23] + c->h_scal[29];
                solved = true; nu = 2; origin_lambda = lambda;
                break;
            }
            c->factor_valid = false;
            if (lambda == 0.0) {
                // the failed factorisation overwrote the tile pool: rebuild S, then the modified Cholesky picks lambda
                // (trust_region.cpp:345-364)
                psba_launch_schur(c, 0.0);
                double delta, beta, sum; int nscalar = 0;
                bool dense = !psba_cholmod_use_tiles(c);
                if (!dense) {
                    double ratio = 0.0;
                    sum = psba_launch_cholmod_tiles(c, &delta, &beta, &nscalar, nullptr, &ratio);
                    const char *e = getenv("PSBA_CHOLMOD_TILES");
                    if (ratio > 1.0 && psba_cholmod_dense_possible(c) && !(e && atoi(e))) { dense = true; psba_launch_schur(c, 0.0); }
                }
                if (dense) {
                    const size_t nn = (size_t)c->N * c->N;
                    if (!c->Sdense) c->Sdense = (double *)psba_dev_alloc(c, nn * sizeof(double), true);
                    psba_tiles_to_dense(c, c->Sdense, true);
                    sum = psba_launch_cholmod(c, &delta, &beta, &nscalar);
                }
                lambda = fabs(sum) / c->N;
                if ((size_t)c->n_cholmod_events < c->force_lambda.size()) lambda = c->force_lambda[c->n_cholmod_events];
                c->n_cholmod_events++;
                {
                    psba_trace_rec r;
                    r.phase = 2; r.itno = c->itno; r.err = (double)nscalar; r.rho = sum; r.mu = lambda; r.delta = 0; r.pnorm = 0; r.accepted = 0;
                    c->trace.push_back(r);
                }
            } else lambda = 2 * lambda;
            if (c->verbose) printf("chol failed.\n");
            if (origin_lambda != 0.0) {
                if (nu > 4) { *finalErr = ex_L2; return PSBA_ITER_TURN_TO_LM; }
                lambda = lambda * nu; nu = nu * 2;
            }
        }
        // ---- the scalars of trust_region.cpp:125-130, 166-176 from the six numbers
        const double gtBg = 2 * JgJg, gtg = gg;
        const double alpha = -gtg / gtBg;                          // P_U = alpha g
        const double pUtBpU = alpha * alpha * gtBg, pUtBpB = 2 * alpha * JgJB, pBtBpB = 2 * JBJB;
        const double pUpU = alpha * alpha * gtg, pUg = alpha * gtg, pBpB = BB, pBg = gB, pUpB = alpha * gB;

        iter_flag = PSBA_ITER_CONTINUE;
        while (iter_flag == PSBA_ITER_CONTINUE) {                 // :180
            // ---- compute_p_2 (:520-595): P = cU P_U + cB P_B
            const double den = -pUtBpB * pUtBpB + pBtBpB * pUtBpU;
            const double eta1 = (pBg * pUtBpB) / den - (pBtBpB * pUg) / den;
            const double eta2 = (pUg * pUtBpB) / den - (pBg * pUtBpU) / den;
            double cU = eta1, cB = eta2;
            p_norm = sqrt(eta1 * eta1 * pUpU + 2 * eta1 * eta2 * pUpB + eta2 * eta2 * pBpB);
            if (p_norm > dk) {
                const double pU_norm = sqrt(pUpU), pB_norm = sqrt(pBpB);
                if (pU_norm > dk) { cU = dk / pU_norm; cB = 0.0; p_norm = dk; }
                else if (pB_norm <= dk) { cU = 0.0; cB = 1.0; p_norm = sqrt(p_norm + pBpB); }   // printed value only
                else {
                    // dog-leg: A = P_B - P_U, B = 2 P_U - P_B
                    const double a = pBpB - 2 * pUpB + pUpU;
                    double b = 3 * pUpB - pBpB - 2 * pUpU, cc = 4 * pUpU - 4 * pUpB + pBpB;
                    b = 2 * b; cc = cc - dk * dk;
                    double b2_4ac = b * b - 4 * a * cc; if (fabs(b2_4ac) < 1e-12) b2_4ac = 0;
                    const double tau = (-b + sqrt(b2_4ac)) / (2 * a);
                    cU = 2 - tau; cB = tau - 1;                    // P_U + (tau - 1)(P_B - P_U)
                    p_norm = dk;
                }
            }
            // ---- candidate = p + P, actual cost (:184-194)
            if (psba_seq_begin(c, SEQ_TR_RADIUS, 0.0, cU * alpha, cB)) {
                psba_launch_step_newp(c, cU * alpha, c->g, cB, c->P_B);
                psba_enqueue_cost(c, 1 - c->cur, nullptr);
                CUDA_CHECK(cudaMemcpyAsync(c->h_scal, c->d_scal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                psba_seq_end(c);
            }
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            act_ex_L2 = c->h_scal[0];
            if (fabs((ex_L2 - act_ex_L2) / ex_L2) < PSBA_EPSILON2) { iter_flag = PSBA_ITER_DP_NO_CHANGE; break; }
            // ---- predicted cost (:208-212)
            const double Jx_norm = cU * cU * pUtBpU + 2 * cU * cB * pUtBpB + cB * cB * pBtBpB;
            pred_ex_L2 = cU * pUg + cB * pBg;
            pred_ex_L2 += ex_L2 + Jx_norm / 2;
            const double rho = (ex_L2 - act_ex_L2) / (ex_L2 - pred_ex_L2);
            int acc = 0;
            if (rho < (1.0 / 4.0) || act_ex_L2 > ex_L2) {
                dk = dk / 4;
                if (c->verbose) printf("iter %d reduce region\n", c->itno);
            } else if (rho >= (3.0 / 4.0) && act_ex_L2 < ex_L2) {
                iter_flag = PSBA_ITER_PASS; acc = 1;
                c->cur = 1 - c->cur; c->lin_valid = false;        // update_p
                *finalErr = act_ex_L2;
                dk = fmin(2 * dk, PSBA_MAX_DELTA);
            } else if (rho >= (1.0 / 4.0) && rho < (3.0 / 4.0) && act_ex_L2 < ex_L2) {
                iter_flag = PSBA_ITER_PASS; acc = 1;
                c->cur = 1 - c->cur; c->lin_valid = false;
                *finalErr = act_ex_L2;
            } else if (std::isnan(rho)) {
                *finalErr = ex_L2;
                return PSBA_ITER_TURN_TO_LM;
            }
            if (c->verbose)
                printf("itno=%d\tErr:%.15E\tDelta=%f\tRho=%f\tnorm_p=%f\tLambda=%E\n", c->itno, act_ex_L2, dk, rho, p_norm, lambda);
            trace(c, 1, act_ex_L2, rho, lambda, dk, p_norm, acc);
            if (fabs((act_ex_L2 - ex_L2) / ex_L2) <= PSBA_EPSILON2) { iter_flag = PSBA_ITER_ERR_SMALL_ENOUGH; break; }
            if (rho < 1.0 / 4) {
                notgood_cnt++;
                if (notgood_cnt >= 5) { iter_flag = PSBA_ITER_TURN_TO_LM; break; }
            } else notgood_cnt = 0;
            if (rho > 3.0 / 4 && act_ex_L2 < ex_L2) {
                good_iters++;
                if (good_iters >= 10) { lambda = 0.0; origin_lambda = 0.0; good_iters = 0; }
            } else good_iters = 0;
            if (rho > (1.0 / 4) && act_ex_L2 < ex_L2) ex_L2 = act_ex_L2;
        }
        if (iter_flag != PSBA_ITER_PASS) break;
    }
    return iter_flag;
}

extern "C" int psba_trust_region(psba_ctx *c, int cnp, int pnp, int mnp, int n3Dpts, int nCams, int n2Dprojs,
                                 int *blk_idx, double *finalErr)
{
    (void)cnp; (void)pnp; (void)mnp; (void)n3Dpts; (void)nCams; (void)n2Dprojs; (void)blk_idx;
    if (c->tr_fused && c->camera_solver != 1) return psba_trust_region_fused(c, finalErr);
    int iter_flag, notgood_cnt = 0, good_iters = 0, nu = 2;
    double ex_L2, pred_ex_L2, act_ex_L2, gtBg, gtg, dk = 1, lambda = 0, p_norm, origin_lambda = 0.0;
    const size_t Tl = (size_t)c->N + 3 * (size_t)c->n;
    double r3[3], d6[6];

    ex_L2 = psba_launch_cost(c, c->cur, nullptr);                 // trust_region.cpp:106-107
    iter_flag = PSBA_ITER_CONTINUE;
    for (; c->itno < c->max_iter; c->itno++) {
        psba_launch_linearize(c, 2.0, -2.0);                      // :117-122, 133-137 (one fused pass)
        psba_launch_Jdot(c, c->g, c->g, nullptr, r3);             // :125-126
        gtBg = 2 * r3[0];
        psba_launch_dots(c, c->g, c->g, c->g, d6);
        gtg = d6[0];
        // P_U[i] = -(g[i]*gtg)/gtBg  (:128-130)
        psba_launch_axpby(c, -gtg / gtBg, c->g, 0.0, c->g, c->P_U);
        bool solved = false;
        while (!solved) {                                         // :141-163
            solved = compute_PB(c, &lambda);
            if (!solved) {
                if (c->verbose) printf("chol failed.\n");
                if (origin_lambda != 0.0) {
                    if (nu > 4) { *finalErr = ex_L2; return PSBA_ITER_TURN_TO_LM; }
                    lambda = lambda * nu; nu = nu * 2;
                }
            } else { nu = 2; origin_lambda = lambda; }
        }
        psba_launch_Jdot(c, c->P_U, c->P_B, nullptr, r3);         // :166-176
        const double pUtBpU = 2 * r3[0], pUtBpB = 2 * r3[1], pBtBpB = 2 * r3[2];
        psba_launch_dots(c, c->P_U, c->P_B, c->g, d6);
        const double pUpU = d6[0], pUg = d6[2], pBpB = d6[3], pBg = d6[4];

        iter_flag = PSBA_ITER_CONTINUE;
        while (iter_flag == PSBA_ITER_CONTINUE) {                 // :180
            // ---- compute_p_2 (:520-595)
            const double den = -pUtBpB * pUtBpB + pBtBpB * pUtBpU;
            const double eta1 = (pBg * pUtBpB) / den - (pBtBpB * pUg) / den;
            const double eta2 = (pUg * pUtBpB) / den - (pBg * pUtBpU) / den;
            psba_launch_axpby(c, eta1, c->P_U, eta2, c->P_B, c->P);
            psba_launch_dots(c, c->P, c->P, c->g, d6);
            p_norm = sqrt(d6[0]);
            if (p_norm > dk) {
                const double pU_norm = sqrt(pUpU), pB_norm = sqrt(pBpB);
                if (pU_norm > dk) {
                    psba_launch_axpby(c, dk / pU_norm, c->P_U, 0.0, c->P_U, c->P);
                    p_norm = dk;
                } else if (pB_norm <= dk) {
                    CUDA_CHECK(cudaMemcpyAsync(c->P, c->P_B, Tl * 8, cudaMemcpyDeviceToDevice, c->stream));
                    p_norm = sqrt(p_norm + pBpB);                 // printed value only
                } else {
                    // dog-leg: A = P_B - P_U, B = 2 P_U - P_B
                    double *A = c->UVdiag_scr, *B = c->P;
                    psba_launch_axpby(c, 1.0, c->P_B, -1.0, c->P_U, A);
                    psba_launch_axpby(c, 2.0, c->P_U, -1.0, c->P_B, B);
                    psba_launch_dots(c, A, B, B, d6);
                    const double a = d6[0];
                    double b = d6[1], cc = d6[3];
                    b = 2 * b; cc = cc - dk * dk;
                    double b2_4ac = b * b - 4 * a * cc; if (fabs(b2_4ac) < 1e-12) b2_4ac = 0;
                    const double tau = (-b + sqrt(b2_4ac)) / (2 * a);
                    psba_launch_axpby(c, 1.0, c->P_U, tau - 1, A, c->P);
                    p_norm = dk;
                }
            }
            // ---- candidate = p + P, actual cost (:184-194)
            CUDA_CHECK(cudaMemcpyAsync(c->dp, c->P, Tl * 8, cudaMemcpyDeviceToDevice, c->stream));
            psba_launch_newp(c);
            act_ex_L2 = psba_launch_cost(c, 1 - c->cur, nullptr);
            if (fabs((ex_L2 - act_ex_L2) / ex_L2) < PSBA_EPSILON2) { iter_flag = PSBA_ITER_DP_NO_CHANGE; break; }
            // ---- predicted cost (:208-212)
            psba_launch_Jdot(c, c->dp, c->dp, nullptr, r3);
            const double Jx_norm = 2 * r3[0];
            psba_launch_dots(c, c->g, c->P, c->P, d6);
            pred_ex_L2 = d6[1];
            pred_ex_L2 += ex_L2 + Jx_norm / 2;
            const double rho = (ex_L2 - act_ex_L2) / (ex_L2 - pred_ex_L2);
            int acc = 0;
            if (rho < (1.0 / 4.0) || act_ex_L2 > ex_L2) {
                dk = dk / 4;
                if (c->verbose) printf("iter %d reduce region\n", c->itno);
            } else if (rho >= (3.0 / 4.0) && act_ex_L2 < ex_L2) {
                iter_flag = PSBA_ITER_PASS; acc = 1;
                c->cur = 1 - c->cur; c->lin_valid = false;        // update_p
                *finalErr = act_ex_L2;
                dk = fmin(2 * dk, PSBA_MAX_DELTA);
            } else if (rho >= (1.0 / 4.0) && rho < (3.0 / 4.0) && act_ex_L2 < ex_L2) {
                iter_flag = PSBA_ITER_PASS; acc = 1;
                c->cur = 1 - c->cur; c->lin_valid = false;
                *finalErr = act_ex_L2;
            } else if (std::isnan(rho)) {
                *finalErr = ex_L2;
                return PSBA_ITER_TURN_TO_LM;
            }
            if (c->verbose)
                printf("itno=%d\tErr:%.15E\tDelta=%f\tRho=%f\tnorm_p=%f\tLambda=%E\n", c->itno, act_ex_L2, dk, rho, p_norm, lambda);
            trace(c, 1, act_ex_L2, rho, lambda, dk, p_norm, acc);
            if (fabs((act_ex_L2 - ex_L2) / ex_L2) <= PSBA_EPSILON2) { iter_flag = PSBA_ITER_ERR_SMALL_ENOUGH; break; }
            if (rho < 1.0 / 4) {
                notgood_cnt++;
                if (notgood_cnt >= 5) { iter_flag = PSBA_ITER_TURN_TO_LM; break; }
            } else notgood_cnt = 0;
            if (rho > 3.0 / 4 && act_ex_L2 < ex_L2) {
                good_iters++;
                if (good_iters >= 10) { lambda = 0.0; origin_lambda = 0.0; good_iters = 0; }
            } else good_iters = 0;
            if (rho > (1.0 / 4) && act_ex_L2 < ex_L2) ex_L2 = act_ex_L2;
        }
        if (iter_flag != PSBA_ITER_PASS) break;
    }
    return iter_flag;
}

extern "C" int psba_solve(psba_ctx *c, double *initErr, double *finalErr, int *itno)
{
    int flag;
    double fe = 0.0;
    c->itno = 0; c->trace.clear(); c->n_cholmod_events = 0;
    while (true) {                                                // main.cpp:193-208
        flag = psba_levmar(c, 6, 3, 2, c->n_glob, c->m, c->o_glob, &fe);
        if (flag != PSBA_ITER_TURN_TO_TR) break;
        flag = psba_trust_region(c, 6, 3, 2, c->n_glob, c->m, c->o_glob, nullptr, &fe);
        if (flag != PSBA_ITER_TURN_TO_LM) break;
    }
    if (initErr) *initErr = c->initErr;
    if (finalErr) *finalErr = fe;
    if (itno) *itno = c->itno;
    return flag;
}
