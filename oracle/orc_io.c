/*
 * orc_io.c -- ORACLE (test infrastructure): restatement of the SBA text loaders.
 *
 * Follows PSBA/readparams.cpp:444-518 (readInitialSBAEstimate) and its helpers
 * (findNcameras :30-52, countNDoubles :121-156, readCameraParams :169-232,
 * readNpointsAndNprojections :247-290, readPointParamsAndProjections :332-423),
 * the input filter PSBA/misc.cpp:21-49 (quat2vec) and the split in PSBA/main.cpp:131-149.
 * The dense visibility mask vmask[n*m] (readparams.cpp:415) is replaced by per-point frame
 * lists in file order; orc_generate_idxs() reproduces what generate_idxs() derives from it.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "psba_oracle.h"

#define MAXSTRLEN 2048

static void skip_line(FILE *fp)
{   /* readparams.cpp:16-20 */
    char buf[MAXSTRLEN];
    while (!feof(fp))
        if (!fgets(buf, MAXSTRLEN - 1, fp) || buf[strlen(buf) - 1] == '\n') break;
}

/* readparams.cpp:30-52 : one camera per non-comment line */
static int find_ncameras(FILE *fp)
{
    int ncams = 0, ch;
    while (!feof(fp)) {
        if ((ch = fgetc(fp)) == '#') { skip_line(fp); continue; }
        if (feof(fp)) break;
        ungetc(ch, fp);
        skip_line(fp);
        if (ferror(fp)) return -1;
        ++ncams;
    }
    return ncams;
}

/* readparams.cpp:121-156 : number of doubles on the first non-comment line; rewinds */
static int count_ndoubles(FILE *fp)
{
    int ch, np, i;
    char buf[MAXSTRLEN], *s;
    double dummy;
    while (!feof(fp)) {
        if ((ch = fgetc(fp)) == '#') { skip_line(fp); continue; }
        if (feof(fp)) return 0;
        ungetc(ch, fp);
        if (!fgets(buf, MAXSTRLEN - 1, fp)) return -1;
        for (np = i = 0, s = buf; 1; ++np, s += i) {
            ch = sscanf(s, "%lf%n", &dummy, &i);
            if (ch == 0 || ch == EOF) break;
        }
        rewind(fp);
        return np;
    }
    return 0;
}

static int read_ndoubles(FILE *fp, double *vals, int nvals)
{   /* readparams.cpp:58-77 */
    int i, n = 0, j;
    for (i = 0; i < nvals; ++i) {
        j = fscanf(fp, "%lf", vals + i);
        if (j == EOF) return EOF;
        if (j != 1 || ferror(fp)) return EOF - 1;
        n += j;
    }
    return n;
}

/* misc.cpp:21-49 : copy intrinsics(+distortion); normalise the quaternion, make its scalar
 * part non-negative, keep the vector part; copy the translation. */
void orc_quat2vec(const double *inp, int nin, double *outp, int nout)
{
    double mag, sg;
    int i;
    if (nin > 7) for (i = 0; i < nin - 7; ++i) outp[i] = inp[i];
    else i = 0;
    mag = sqrt(inp[i] * inp[i] + inp[i + 1] * inp[i + 1] + inp[i + 2] * inp[i + 2] + inp[i + 3] * inp[i + 3]);
    sg = (inp[i] >= 0.0) ? 1.0 : -1.0;
    mag = sg / mag;
    outp[i] = inp[i + 1] * mag;
    outp[i + 1] = inp[i + 2] * mag;
    outp[i + 2] = inp[i + 3] * mag;
    i += 3;
    for (; i < nout; ++i) outp[i] = inp[i + 1];
}

int orc_read_sba(const char *camsfname, const char *ptsfname, int cnp,
                 int *ncams, int *n3Dpts, int *n2Dprojs,
                 double **motstruct, double **initrot, double **imgpts,
                 int **pt_nframes, int **frames)
{
    const int pnp = 3, mnp = 2, filecnp = cnp + 1;
    FILE *fpc = fopen(camsfname, "r"), *fpp = fopen(ptsfname, "r");
    int ch, nfirst, npts = 0, nprojs = 0, covvals = 0, nframes, n, i, m;
    double tofilter[64], cov[4];
    if (!fpc || !fpp) { fprintf(stderr, "orc_read_sba: cannot open %s / %s\n", camsfname, ptsfname); return 1; }
    if (filecnp > 64) return 2;

    m = find_ncameras(fpc);                       /* readparams.cpp:463 */
    /* first pass over the points file, readparams.cpp:247-290 */
    nfirst = count_ndoubles(fpp);
    while (!feof(fpp)) {
        if ((ch = fgetc(fpp)) == '#') { skip_line(fpp); continue; }
        if (feof(fpp)) break;
        ungetc(ch, fpp);
        for (i = 0; i < pnp; ++i) { double d; if (fscanf(fpp, "%lf", &d) != 1) break; }
        if (fscanf(fpp, "%d", &nframes) != 1) { fprintf(stderr, "orc_read_sba: bad frame count\n"); return 3; }
        if (npts == 0) {
            int rest = nfirst - (pnp + 1);
            if (rest == nframes * (mnp + 1 + mnp * mnp)) covvals = mnp * mnp;              /* FULLCOV */
            else if (rest == nframes * (mnp + 1 + mnp * (mnp + 1) / 2)) covvals = mnp * (mnp + 1) / 2; /* TRICOV */
            else covvals = 0;
        }
        skip_line(fpp);
        nprojs += nframes;
        ++npts;
    }
    *ncams = m; *n3Dpts = npts; *n2Dprojs = nprojs;
    *motstruct = (double *)malloc(((size_t)m * cnp + (size_t)npts * pnp) * sizeof(double));
    *initrot = (double *)malloc((size_t)m * 4 * sizeof(double));
    *imgpts = (double *)malloc((size_t)nprojs * mnp * sizeof(double));
    *pt_nframes = (int *)malloc((size_t)npts * sizeof(int));
    *frames = (int *)malloc((size_t)nprojs * sizeof(int));
    rewind(fpc); rewind(fpp);

    /* cameras, readparams.cpp:169-232 */
    if ((n = count_ndoubles(fpc)) != filecnp) {
        fprintf(stderr, "orc_read_sba: expected %d camera parameters, first line contains %d!\n", filecnp, n);
        return 4;
    }
    {
        double *params = *motstruct, *ir = *initrot;
        while (!feof(fpc)) {
            if ((ch = fgetc(fpc)) == '#') { skip_line(fpc); continue; }
            if (feof(fpc)) break;
            ungetc(ch, fpc);
            n = read_ndoubles(fpc, tofilter, filecnp);
            if (n == EOF) break;
            if (n != filecnp) { fprintf(stderr, "orc_read_sba: short camera line\n"); return 5; }
            orc_quat2vec(tofilter, filecnp, params, cnp);
            /* readparams.cpp:222-226 : q0 recomputed from the filtered vector part */
            ir[1] = params[cnp - 6]; ir[2] = params[cnp - 5]; ir[3] = params[cnp - 4];
            ir[0] = sqrt(1.0 - ir[1] * ir[1] - ir[2] * ir[2] - ir[3] * ir[3]);
            params += cnp; ir += 4;
        }
    }
    /* points + projections, readparams.cpp:332-423 */
    {
        double *params = *motstruct + (size_t)m * cnp, *projs = *imgpts;
        int *fr = *frames, ptno = 0, frameno;
        while (!feof(fpp)) {
            if ((ch = fgetc(fpp)) == '#') { skip_line(fpp); continue; }
            if (feof(fpp)) break;
            ungetc(ch, fpp);
            n = read_ndoubles(fpp, params, pnp);
            if (n == EOF) break;
            if (n != pnp) { fprintf(stderr, "orc_read_sba: bad point line %d\n", ptno); return 6; }
            params += pnp;
            if (fscanf(fpp, "%d", &nframes) != 1) return 7;
            (*pt_nframes)[ptno] = nframes;
            for (i = 0; i < nframes; ++i) {
                if (fscanf(fpp, "%d", &frameno) != 1) return 8;
                if (frameno >= m) {
                    fprintf(stderr, "orc_read_sba: projection for frame %d but only %d cameras\n", frameno, m);
                    return 9;
                }
                if (read_ndoubles(fpp, projs, mnp) != mnp) return 10;
                projs += mnp;
                if (covvals) if (read_ndoubles(fpp, cov, covvals) != covvals) return 11;
                *fr++ = frameno;
            }
            if (fscanf(fpp, "\n") == EOF) { /* trailing newline consumed, readparams.cpp:418 */ }
            ptno++;
        }
    }
    fclose(fpc); fclose(fpp);
    return 0;
}

/* main.cpp:131-149 generalised to the three column layouts:
 * origin_cnp 11 = K(5)|qv(3)|t(3);  16 = K(5)|kc(5)|qv(3)|t(3) (kc dropped: the reference has
 * no distortion model);  6 = qv(3)|t(3) with K supplied by the caller (Kparas untouched). */
void orc_split_motion(const double *mot, int origin_cnp, int m, double *Kparas, double *camsEx)
{
    int j, k;
    for (j = 0; j < m; ++j) {
        const double *c = mot + (size_t)j * origin_cnp;
        if (origin_cnp >= 11) for (k = 0; k < 5; ++k) Kparas[j * 5 + k] = c[k];
        camsEx[j * 6 + 0] = 0.0; camsEx[j * 6 + 1] = 0.0; camsEx[j * 6 + 2] = 0.0;  /* main.cpp:131-136 */
        for (k = 0; k < 3; ++k) camsEx[j * 6 + 3 + k] = c[origin_cnp - 3 + k];
    }
}

static int cmp_int(const void *a, const void *b) { return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }

/* misc.cpp:178-218 : the reference scans vmask[i*m+j] with j ascending, so the observation
 * list is point-major with cameras ascending.  impts stays in FILE order (readparams.cpp:399),
 * identical only when every point lists its frames in ascending order. */
void orc_generate_idxs(int m, int n, int o, const int *pt_nframes, const int *frames, int *iidx, int *jidx)
{
    int i, k = 0, f, a;
    (void)m; (void)o;
    for (i = 0; i < n; ++i) {
        int nf = pt_nframes[i];
        for (f = 0; f < nf; ++f) { iidx[k + f] = i; jidx[k + f] = frames[k + f]; }
        for (a = 1; a < nf; ++a) if (jidx[k + a] < jidx[k + a - 1]) { qsort(jidx + k, nf, sizeof(int), cmp_int); break; }
        k += nf;
    }
}
