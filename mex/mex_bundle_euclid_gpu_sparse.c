/*
 * mex_bundle_euclid_gpu_sparse.c -- bundle_euclid.m:81-267 behind one mex call, on an observation LIST.
 *
 *   [K_ Te_ w_ Xe_ error_] = mex_bundle_euclid_gpu_sparse(K, Te, w, Xe, obs_xy, obs_pt, obs_cam, pivot, flags)
 *     K 4xm, Te 3xm, w 3xm, Xe 4xn                       (bundle_euclid.m:5-8)
 *     obs_xy 2 x nobs, obs_pt / obs_cam 1 x nobs (1-based point and camera of every visible cell, in the
 *     reference's traversal order: camera-major, points ascending within a camera) -- what the dense
 *     interface spells as x(1:2,i,j) with visibility(i,j) ~= 0 (bundle_euclid.m:9,18,50,81)
 *     pivot 1xm or [], flags = [num_variableK fix_structure fix_motion verbose]
 */
#include <stdint.h>
#include "mex.h"
#include "vlg_ba.h"

void mexFunction(int nout, mxArray *pout[], int nin, const mxArray *pin[])
{
    vlg_ba_opts o;
    int m, n, rc, ne = 0, k;
    size_t nobs, t;
    const double *piv = 0, *fl, *dp, *dc;
    int32_t *pt, *cam;
    double *err;
    (void)nout;
    if (nin != 9)
        mexErrMsgIdAndTxt("vlg:euclid_sparse:nargin", "mex_bundle_euclid_gpu_sparse(K, Te, w, Xe, obs_xy, obs_pt, obs_cam, pivot, flags)");
    m = (int)mxGetN(pin[2]);
    n = (int)mxGetN(pin[3]);
    nobs = mxGetN(pin[4]);
    if (mxGetM(pin[0]) != 4 || (int)mxGetN(pin[0]) != m || mxGetM(pin[1]) != 3 || (int)mxGetN(pin[1]) != m || mxGetM(pin[2]) != 3 ||
        mxGetM(pin[3]) != 4 || (nobs > 0 && mxGetM(pin[4]) != 2) || mxGetM(pin[5]) * mxGetN(pin[5]) != nobs ||
        mxGetM(pin[6]) * mxGetN(pin[6]) != nobs)
        mexErrMsgIdAndTxt("vlg:euclid_sparse:shape", "expected K 4xm, Te 3xm, w 3xm, Xe 4xn, obs_xy 2xnobs, obs_pt 1xnobs, obs_cam 1xnobs");
    if (mxGetM(pin[7]) * mxGetN(pin[7]) == (size_t)m) piv = mxGetPr(pin[7]);
    if (mxGetM(pin[8]) * mxGetN(pin[8]) < 4) mexErrMsgIdAndTxt("vlg:euclid_sparse:flags", "flags must have 4 entries");
    fl = mxGetPr(pin[8]);
    vlg_ba_opts_default(&o);
    o.num_variableK = (int)fl[0]; o.fix_structure = fl[1] != 0; o.fix_motion = fl[2] != 0; o.verbose = fl[3] != 0;
    /* MATLAB's 1-based doubles -> 0-based int32 */
    pt = (int32_t *)mxCalloc(nobs + 1, sizeof(int32_t));
    cam = (int32_t *)mxCalloc(nobs + 1, sizeof(int32_t));
    dp = mxGetPr(pin[5]); dc = mxGetPr(pin[6]);
    for (t = 0; t < nobs; t++) { pt[t] = (int32_t)dp[t] - 1; cam[t] = (int32_t)dc[t] - 1; }
    pout[0] = mxCreateDoubleMatrix(4, m, mxREAL);
    pout[1] = mxCreateDoubleMatrix(3, m, mxREAL);
    pout[2] = mxCreateDoubleMatrix(3, m, mxREAL);
    pout[3] = mxCreateDoubleMatrix(4, n, mxREAL);
    err = (double *)mxCalloc((size_t)o.max_iter + 2, sizeof(double));
    rc = vlg_ba_bundle_euclid_sparse(&o, m, n, mxGetPr(pin[0]), mxGetPr(pin[1]), mxGetPr(pin[2]), mxGetPr(pin[3]), (int64_t)nobs,
                                     mxGetPr(pin[4]), pt, cam, piv, mxGetPr(pout[0]), mxGetPr(pout[1]), mxGetPr(pout[2]),
                                     mxGetPr(pout[3]), err, &ne);
    mxFree(pt); mxFree(cam);
    if (rc != VLG_BA_OK) mexErrMsgIdAndTxt("vlg:euclid_sparse:gpu", vlg_ba_last_error(0));
    pout[4] = mxCreateDoubleMatrix(1, ne, mxREAL);
    for (k = 0; k < ne; k++) mxGetPr(pout[4])[k] = err[k];
    mxFree(err);
}
