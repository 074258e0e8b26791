/*
 * mex_bundle_euclid_gpu.c -- the whole of bundle_euclid.m:81-267 behind one mex call.
 *
 *   [K_ Te_ w_ Xe_ error_] = mex_bundle_euclid_gpu(K, Te, w, Xe, x, visible, pivot, flags)
 *     K 4xm, Te 3xm, w 3xm, Xe 4xn, x 3xnxm             (bundle_euclid.m:5-9)
 *     visible nxm or [] (derive from x, bundle_euclid.m:50), pivot 1xm or []
 *     flags = [num_variableK fix_structure fix_motion verbose]  (parsed by bundle_euclid_gpu.m from the
 *     same option strings as bundle_euclid.m:54-78)
 */
#include "mex.h"
#include "vlg_ba.h"

void mexFunction(int nout, mxArray *pout[], int nin, const mxArray *pin[])
{
    vlg_ba_opts o;
    int m, n, rc, ne = 0, k;
    const double *vis = 0, *piv = 0, *fl;
    double *err;
    (void)nout;
    if (nin != 8) mexErrMsgIdAndTxt("vlg:euclid:nargin", "mex_bundle_euclid_gpu(K, Te, w, Xe, x, visible, pivot, flags)");
    m = (int)mxGetN(pin[2]);
    n = (int)mxGetN(pin[3]);
    if (mxGetM(pin[0]) != 4 || (int)mxGetN(pin[0]) != m || mxGetM(pin[1]) != 3 || (int)mxGetN(pin[1]) != m || mxGetM(pin[2]) != 3 ||
        mxGetM(pin[3]) != 4 || mxGetM(pin[4]) != 3 || (size_t)mxGetN(pin[4]) != (size_t)n * m)
        mexErrMsgIdAndTxt("vlg:euclid:shape", "expected K 4xm, Te 3xm, w 3xm, Xe 4xn, x 3xnxm");
    if (mxGetM(pin[5]) * mxGetN(pin[5]) == (size_t)n * m) vis = mxGetPr(pin[5]);
    if (mxGetM(pin[6]) * mxGetN(pin[6]) == (size_t)m) piv = mxGetPr(pin[6]);
    if (mxGetM(pin[7]) * mxGetN(pin[7]) < 4) mexErrMsgIdAndTxt("vlg:euclid:flags", "flags must have 4 entries");
    fl = mxGetPr(pin[7]);
    vlg_ba_opts_default(&o);
    o.num_variableK = (int)fl[0]; o.fix_structure = fl[1] != 0; o.fix_motion = fl[2] != 0; o.verbose = fl[3] != 0;
    pout[0] = mxCreateDoubleMatrix(4, m, mxREAL);
    pout[1] = mxCreateDoubleMatrix(3, m, mxREAL);
    pout[2] = mxCreateDoubleMatrix(3, m, mxREAL);
    pout[3] = mxCreateDoubleMatrix(4, n, mxREAL);
    err = (double *)mxCalloc((size_t)o.max_iter + 2, sizeof(double));
    rc = vlg_ba_bundle_euclid(&o, m, n, mxGetPr(pin[0]), mxGetPr(pin[1]), mxGetPr(pin[2]), mxGetPr(pin[3]), mxGetPr(pin[4]),
                              vis, piv, mxGetPr(pout[0]), mxGetPr(pout[1]), mxGetPr(pout[2]), mxGetPr(pout[3]), err, &ne);
    if (rc != VLG_BA_OK) mexErrMsgIdAndTxt("vlg:euclid:gpu", vlg_ba_last_error(0));
    pout[4] = mxCreateDoubleMatrix(1, ne, mxREAL);
    for (k = 0; k < ne; k++) mxGetPr(pout[4])[k] = err[k];
    mxFree(err);
}
