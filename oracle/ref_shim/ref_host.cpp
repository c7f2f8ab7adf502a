/*
 * ref_host.cpp -- ORACLE PIN (test infrastructure): C entry points onto the reference's own
 * loader and index generator (PSBA/readparams.cpp, PSBA/misc.cpp, compiled in place).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "psba_oracle.h"

/* PSBA/psba.h:45 declares it; the loaders never touch it */
FILE *debug_file = NULL;

/* PSBA/readparams.h:8-11, PSBA/misc.h:8,21-22 */
void readInitialSBAEstimate(char *camsfname, char *ptsfname, int cnp, int pnp, int mnp,
    void (*infilter)(double *pin, int nin, double *pout, int nout), int cnfp,
    int *ncams, int *n3Dpts, int *n2Dprojs,
    double **motstruct, double **initrot, double **imgpts, double **covimgpts, char **vmask);
void quat2vec(double *inp, int nin, double *outp, int nout);
void generate_idxs(int nCams, int n3Dpts, int n2Dprojs, double *impts_data, char *vmask,
    int *comm3DIdx, int *comm3DIdxCnt, int *iidx, int *jidx, int *blk_idx);

extern "C" {

/* main.cpp:102-106 : readInitialSBAEstimate(cams, pts, origin_cnp, 3, 2, quat2vec, origin_cnp+1, ...) */
int ref_read_sba(const char *cams, const char *pts, int origin_cnp, int *m, int *n, int *o,
                 double **mot, double **initrot, double **impts, char **vmask)
{
    double *cov = NULL;
    readInitialSBAEstimate((char *)cams, (char *)pts, origin_cnp, 3, 2, quat2vec, origin_cnp + 1,
                           m, n, o, mot, initrot, impts, &cov, vmask);
    if (cov) free(cov);
    return 0;
}

void ref_quat2vec(double *inp, int nin, double *outp, int nout) { quat2vec(inp, nin, outp, nout); }

/* main.cpp:183-188 */
void ref_generate_idxs(int m, int n, int o, double *impts, char *vmask,
                       int *comm3DIdx, int *comm3DIdxCnt, int *iidx, int *jidx, int *blk_idx)
{
    generate_idxs(m, n, o, impts, vmask, comm3DIdx, comm3DIdxCnt, iidx, jidx, blk_idx);
}

void ref_free(void *p) { free(p); }

}
