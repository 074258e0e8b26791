/*
 * oracle/ref_glue.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Flat C entry points around the reference's three mexFunctions, which are
 * compiled UNMODIFIED from /root/reference/toolbox/bundle/ by oracle/Makefile
 * (renamed at compile time with -DmexFunction=vlgref_mexN) into
 * oracle/_ref/libvlgref.so.  Each wrapper builds the positional pin[] array the
 * reference expects, calls it, copies the requested outputs and frees them.
 *
 *   vlgref_stage1 <- mex_bundle_1_XABeUVWeAeB.c:72-337  (pin/pout order :76-83,:136-175)
 *   vlgref_stage2 <- mex_bundle_2_Se_.c:15-158          (pin order :21-27)
 *   vlgref_stage3 <- mex_bundle_3_db_new.c:12-170       (pin order :18-29)
 *
 * Nothing in the product (bundleadjustmentmatlab_b200/) links or loads this.
 */
#include "mex.h"

void vlgref_mex1(int nout, mxArray *pout[], int nin, const mxArray *pin[]);
void vlgref_mex2(int nout, mxArray *pout[], int nin, const mxArray *pin[]);
void vlgref_mex3(int nout, mxArray *pout[], int nin, const mxArray *pin[]);
#ifndef VLG_GLUE_NO_PROJECTIVE
void vlgref_pmex1(int nout, mxArray *pout[], int nin, const mxArray *pin[]);
void vlgref_pmex2(int nout, mxArray *pout[], int nin, const mxArray *pin[]);
void vlgref_pmex3(int nout, mxArray *pout[], int nin, const mxArray *pin[]);
#endif

static mxArray wrap(const double *p, size_t ndim, size_t d0, size_t d1, size_t d2, size_t d3)
{
    mxArray a;
    memset(&a, 0, sizeof(a));
    a.pr = (double *)p;
    a.ndim = ndim;
    a.dims[0] = d0; a.dims[1] = d1; a.dims[2] = d2; a.dims[3] = d3;
    a.owns = 0;
    return a;
}

static void take(double *dst, mxArray *src, size_t count)
{
    if (dst) memcpy(dst, src->pr, count * sizeof(double));
    mxDestroyArray(src);
}

/* K 4xm, a num_a x m, b 3xn, X 2xnxm, visible nxm (all column-major doubles). */
void vlgref_stage1(int m, int n, int num_a,
                   const double *K, const double *a, const double *b,
                   const double *X, const double *visible,
                   double *X_hat, double *A, double *B, double *e,
                   double *U, double *V, double *W, double *eA, double *eB)
{
    mxArray in[5];
    const mxArray *pin[5];
    mxArray *pout[9];
    size_t nm = (size_t)n * (size_t)m, na = (size_t)num_a;
    int k;
    in[0] = wrap(K, 2, 4, m, 1, 1);
    in[1] = wrap(a, 2, na, m, 1, 1);
    in[2] = wrap(b, 2, 3, n, 1, 1);
    in[3] = wrap(X, 3, 2, n, m, 1);
    in[4] = wrap(visible, 2, n, m, 1, 1);
    for (k = 0; k < 5; k++) pin[k] = &in[k];
    vlgref_mex1(9, pout, 5, pin);
    take(X_hat, pout[0], 2 * nm);
    take(A,     pout[1], 2 * na * nm);
    take(B,     pout[2], 6 * nm);
    take(e,     pout[3], 2 * nm);
    take(U,     pout[4], na * na * m);
    take(V,     pout[5], 9 * (size_t)n);
    take(W,     pout[6], na * 3 * nm);
    take(eA,    pout[7], na * m);
    take(eB,    pout[8], 3 * (size_t)n);
}

/* Y, W num_a x3xnxm; U_ num_a x num_a x m; eA num_a x m; eB 3xn. */
void vlgref_stage2(int m, int n, int num_a,
                   const double *Y, const double *W, const double *U_,
                   const double *eA, const double *eB,
                   double *S, double *e_)
{
    mxArray in[5];
    const mxArray *pin[5];
    mxArray *pout[2];
    size_t na = (size_t)num_a;
    int k;
    in[0] = wrap(Y, 4, na, 3, n, m);
    in[1] = wrap(W, 4, na, 3, n, m);
    in[2] = wrap(U_, 3, na, na, m, 1);
    in[3] = wrap(eA, 2, na, m, 1, 1);
    in[4] = wrap(eB, 2, 3, n, 1, 1);
    for (k = 0; k < 5; k++) pin[k] = &in[k];
    vlgref_mex2(2, pout, 5, pin);
    take(S,  pout[0], na * m * na * m);
    take(e_, pout[1], na * m);
}

void vlgref_stage3(int m, int n, int num_a,
                   const double *W, const double *da, const double *eB,
                   const double *Vinv, const double *K, const double *a,
                   const double *b, const double *X, const double *visible,
                   double *db, double *a_new, double *b_new, double *X_hat)
{
    mxArray in[9];
    const mxArray *pin[9];
    mxArray *pout[4];
    size_t na = (size_t)num_a, nm = (size_t)n * (size_t)m;
    int k;
    in[0] = wrap(W, 4, na, 3, n, m);
    in[1] = wrap(da, 2, na * m, 1, 1, 1);
    in[2] = wrap(eB, 2, 3, n, 1, 1);
    in[3] = wrap(Vinv, 3, 3, 3, n, 1);
    in[4] = wrap(K, 2, 4, m, 1, 1);
    in[5] = wrap(a, 2, na, m, 1, 1);
    in[6] = wrap(b, 2, 3, n, 1, 1);
    in[7] = wrap(X, 3, 2, n, m, 1);
    in[8] = wrap(visible, 2, n, m, 1, 1);
    for (k = 0; k < 9; k++) pin[k] = &in[k];
    vlgref_mex3(4, pout, 9, pin);
    take(db,    pout[0], 3 * (size_t)n);
    take(a_new, pout[1], na * m);
    take(b_new, pout[2], 3 * (size_t)n);
    take(X_hat, pout[3], 2 * nm);
}

#ifndef VLG_GLUE_NO_PROJECTIVE   /* mex/Makefile's shim build has no projective drop-in wrappers to call */
/* ---- projective BA: mex_bundle_proj_1_XABeUVWeAeB.c:88-, mex_bundle_proj_2_Se_.c:15-, mex_bundle_proj_3_db_new.c:34-
 * a 12xm (vec of the 3x4 projection matrix), b 3xn, X 2xnxm, visible nxm. */
void vlgref_pstage1(int m, int n, const double *a, const double *b, const double *X, const double *visible,
                    double *X_hat, double *A, double *B, double *e,
                    double *U, double *V, double *W, double *eA, double *eB)
{
    mxArray in[4];
    const mxArray *pin[4];
    mxArray *pout[9];
    size_t nm = (size_t)n * (size_t)m, na = 12;
    int k;
    in[0] = wrap(a, 2, na, m, 1, 1);
    in[1] = wrap(b, 2, 3, n, 1, 1);
    in[2] = wrap(X, 3, 2, n, m, 1);
    in[3] = wrap(visible, 2, n, m, 1, 1);
    for (k = 0; k < 4; k++) pin[k] = &in[k];
    vlgref_pmex1(9, pout, 4, pin);
    take(X_hat, pout[0], 2 * nm);
    take(A,     pout[1], 2 * na * nm);
    take(B,     pout[2], 6 * nm);
    take(e,     pout[3], 2 * nm);
    take(U,     pout[4], na * na * m);
    take(V,     pout[5], 9 * (size_t)n);
    take(W,     pout[6], na * 3 * nm);
    take(eA,    pout[7], na * m);
    take(eB,    pout[8], 3 * (size_t)n);
}

void vlgref_pstage2(int m, int n, const double *Y, const double *W, const double *U_,
                    const double *eA, const double *eB, double *S, double *e_)
{
    mxArray in[5];
    const mxArray *pin[5];
    mxArray *pout[2];
    size_t na = 12;
    int k;
    in[0] = wrap(Y, 4, na, 3, n, m);
    in[1] = wrap(W, 4, na, 3, n, m);
    in[2] = wrap(U_, 3, na, na, m, 1);
    in[3] = wrap(eA, 2, na, m, 1, 1);
    in[4] = wrap(eB, 2, 3, n, 1, 1);
    for (k = 0; k < 5; k++) pin[k] = &in[k];
    vlgref_pmex2(2, pout, 5, pin);
    take(S,  pout[0], na * m * na * m);
    take(e_, pout[1], na * m);
}

void vlgref_pstage3(int m, int n, const double *W, const double *da, const double *eB, const double *Vinv,
                    const double *a, const double *b, const double *X, const double *visible,
                    double *db, double *a_new, double *b_new, double *X_hat)
{
    mxArray in[8];
    const mxArray *pin[8];
    mxArray *pout[4];
    size_t na = 12, nm = (size_t)n * (size_t)m;
    int k;
    in[0] = wrap(W, 4, na, 3, n, m);
    in[1] = wrap(da, 2, na * m, 1, 1, 1);
    in[2] = wrap(eB, 2, 3, n, 1, 1);
    in[3] = wrap(Vinv, 3, 3, 3, n, 1);
    in[4] = wrap(a, 2, na, m, 1, 1);
    in[5] = wrap(b, 2, 3, n, 1, 1);
    in[6] = wrap(X, 3, 2, n, m, 1);
    in[7] = wrap(visible, 2, n, m, 1, 1);
    for (k = 0; k < 8; k++) pin[k] = &in[k];
    vlgref_pmex3(4, pout, 8, pin);
    take(db,    pout[0], 3 * (size_t)n);
    take(a_new, pout[1], na * m);
    take(b_new, pout[2], 3 * (size_t)n);
    take(X_hat, pout[3], 2 * nm);
}
#endif
