/*
 * mex_bundle_projective_gpu.c -- the whole of bundle_projective.m:60-229 behind one mex call.
 *
 *   [Pp_ Xp_ error_] = mex_bundle_projective_gpu(Pp, Xp, x, visible, flags)
 *     Pp 3x4xm, Xp 4xn, x 3xnxm                          (bundle_projective.m:4-7)
 *     visible nxm or [] (derive from x, bundle_projective.m:38)
 *     flags = [fix_structure fix_motion verbose]         (parsed by bundle_projective_gpu.m from the same
 *     option strings as bundle_projective.m:44-56)
 */
#include "mex.h"
#include "vlg_ba.h"

void mexFunction(int nout, mxArray *pout[], int nin, const mxArray *pin[])
{
    vlg_ba_opts o;
    int m, n, rc, ne = 0, k;
    const double *vis = 0, *fl;
    double *err;
    mwSize dimP[3];
    (void)nout;
    if (nin != 5) mexErrMsgIdAndTxt("vlg:projective:nargin", "mex_bundle_projective_gpu(Pp, Xp, x, visible, flags)");
    n = (int)mxGetN(pin[1]);
    if (mxGetM(pin[0]) != 3 || mxGetN(pin[0]) % 4 != 0 || mxGetM(pin[1]) != 4 || mxGetM(pin[2]) != 3)
        mexErrMsgIdAndTxt("vlg:projective:shape", "expected Pp 3x4xm, Xp 4xn, x 3xnxm");
    m = (int)(mxGetN(pin[0]) / 4);                      /* mxGetN of a 3x4xm array is 4*m */
    if ((size_t)mxGetN(pin[2]) != (size_t)n * m) mexErrMsgIdAndTxt("vlg:projective:shape", "x must be 3xnxm");
    if (mxGetM(pin[3]) * mxGetN(pin[3]) == (size_t)n * m) vis = mxGetPr(pin[3]);
    if (mxGetM(pin[4]) * mxGetN(pin[4]) < 3) mexErrMsgIdAndTxt("vlg:projective:flags", "flags must have 3 entries");
    fl = mxGetPr(pin[4]);
    vlg_ba_opts_default(&o);
    o.model = VLG_BA_MODEL_PROJECTIVE;
    o.fix_structure = fl[0] != 0; o.fix_motion = fl[1] != 0; o.verbose = fl[2] != 0;
    dimP[0] = 3; dimP[1] = 4; dimP[2] = (mwSize)m;
    pout[0] = mxCreateNumericArray(3, dimP, mxDOUBLE_CLASS, mxREAL);
    pout[1] = mxCreateDoubleMatrix(4, n, mxREAL);
    err = (double *)mxCalloc((size_t)o.max_iter + 2, sizeof(double));
    rc = vlg_ba_bundle_projective(&o, m, n, mxGetPr(pin[0]), mxGetPr(pin[1]), mxGetPr(pin[2]), vis,
                                  mxGetPr(pout[0]), mxGetPr(pout[1]), err, &ne);
    if (rc != VLG_BA_OK) mexErrMsgIdAndTxt("vlg:projective:gpu", vlg_ba_last_error(0));
    pout[2] = mxCreateDoubleMatrix(1, ne, mxREAL);
    for (k = 0; k < ne; k++) mxGetPr(pout[2])[k] = err[k];
    mxFree(err);
}
