/*
 * oracle/shim/mex.h -- TEST INFRASTRUCTURE ONLY (CPU oracle build).
 *
 * Minimal stand-in for MATLAB's mex.h so that the reference's three mex C
 * files (toolbox/bundle/mex_bundle_{1_XABeUVWeAeB,2_Se_,3_db_new}.c) compile
 * UNMODIFIED, from where they lie under /root/reference, into oracle/_ref/.
 * Only the API those files use is provided (reference call sites:
 * mex_bundle_1_XABeUVWeAeB.c:85-175, mex_bundle_2_Se_.c:29-66,
 * mex_bundle_3_db_new.c:29-86): mxGetPr, mxGetM, mxGetN,
 * mxCreateNumericArray, mxCreateDoubleMatrix (both zero-initialised, which
 * mex1 relies on for its "+=" accumulation, mex_bundle_1_XABeUVWeAeB.c:285).
 */
#ifndef VLG_ORACLE_SHIM_MEX_H
#define VLG_ORACLE_SHIM_MEX_H

#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>

typedef size_t mwSize;
typedef size_t mwIndex;

typedef enum { mxDOUBLE_CLASS = 6 } mxClassID;
typedef enum { mxREAL = 0 } mxComplexity;

typedef struct mxArray_tag {
    double *pr;
    mwSize  ndim;
    mwSize  dims[8];
    int     owns;
} mxArray;

static inline double *mxGetPr(const mxArray *a) { return a->pr; }
static inline size_t  mxGetM (const mxArray *a) { return a->dims[0]; }
static inline size_t  mxGetN (const mxArray *a)
{
    size_t n = 1, k;
    for (k = 1; k < a->ndim; k++) n *= a->dims[k];
    return n;
}
static inline mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims,
                                            mxClassID cls, mxComplexity cplx)
{
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    size_t total = 1, k;
    (void)cls; (void)cplx;
    a->ndim = ndim;
    for (k = 0; k < ndim; k++) { a->dims[k] = dims[k]; total *= dims[k]; }
    a->pr = (double *)calloc(total ? total : 1, sizeof(double));
    a->owns = 1;
    return a;
}
static inline mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity cplx)
{
    mwSize d[2];
    d[0] = m; d[1] = n;
    return mxCreateNumericArray(2, d, mxDOUBLE_CLASS, cplx);
}
static inline void *mxCalloc(size_t n, size_t size) { return calloc(n ? n : 1, size); }
static inline void mxFree(void *p) { free(p); }
/* last error raised through mexErrMsgIdAndTxt (tests read it; MATLAB would throw) */
static char vlg_shim_last_error[512];
static inline void mexErrMsgIdAndTxt(const char *id, const char *msg)
{
    snprintf(vlg_shim_last_error, sizeof(vlg_shim_last_error), "%s: %s", id, msg);
    fprintf(stderr, "mexErrMsgIdAndTxt: %s\n", vlg_shim_last_error);
    abort();
}
/* mexLock / mexIsLocked / mexAtExit: the shim records them; vlg_shim_run_atexit() is what "MATLAB exits" does */
static int vlg_shim_locked;
static void (*vlg_shim_atexit_fn)(void);
static inline void mexLock(void) { vlg_shim_locked++; }
static inline void mexUnlock(void) { if (vlg_shim_locked > 0) vlg_shim_locked--; }
static inline int mexIsLocked(void) { return vlg_shim_locked > 0; }
static inline int mexAtExit(void (*fn)(void)) { vlg_shim_atexit_fn = fn; return 0; }
static inline void vlg_shim_run_atexit(void) { if (vlg_shim_atexit_fn) vlg_shim_atexit_fn(); }
static inline void mxDestroyArray(mxArray *a)
{
    if (!a) return;
    if (a->owns) free(a->pr);
    free(a);
}

#endif
