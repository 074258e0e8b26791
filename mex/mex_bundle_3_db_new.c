/*
 * mex_bundle_3_db_new.c -- GPU drop-in for toolbox/bundle/mex_bundle_3_db_new.c:12-170.
 *
 *   [db a_new b_new X_hat] = mex_bundle_3_db_new(W, da, eB, V_inv, K, a, b, X, visible)
 *     W num_a x 3 x n x m, da (num_a*m) x 1, eB 3xn, V_inv 3x3xn, K 4xm, a num_a x m, b 3xn,
 *     X 2xnxm, visible nxm                                                (reference :18-29)
 *     db 3xn, a_new num_a x m, b_new 3xn, X_hat 2xnxm                     (reference :67-86)
 *   Like the reference, only the first six camera parameters enter the back-substitution
 *   (reference :113-120).
 */
#include "mex.h"
#include "vlg_ba.h"
#include "vlg_mex_state.h"

void mexFunction(int nout, mxArray *pout[], int nin, const mxArray *pin[])
{
    mwSize dX[3];
    int m, n, num_a, rc;
    (void)nout;
    vlg_mex_keep_state();
    if (nin != 9) mexErrMsgIdAndTxt("vlg:mex3:nargin", "mex_bundle_3_db_new(W, da, eB, V_inv, K, a, b, X, visible): 9 inputs required");
    m = (int)mxGetN(pin[5]);
    n = (int)mxGetN(pin[6]);
    num_a = (int)mxGetM(pin[0]);
    if (!(num_a == 6 || num_a == 7 || num_a == 10) || (int)mxGetM(pin[5]) != num_a) mexErrMsgIdAndTxt("vlg:mex3:num_a", "W and a must have 6, 7 or 10 rows");
    if ((size_t)mxGetN(pin[0]) != (size_t)3 * n * m) mexErrMsgIdAndTxt("vlg:mex3:W", "W must be num_a x 3 x n x m");
    if ((size_t)(mxGetM(pin[1]) * mxGetN(pin[1])) != (size_t)num_a * m) mexErrMsgIdAndTxt("vlg:mex3:da", "da must have num_a*m entries");
    if (mxGetM(pin[2]) != 3 || (int)mxGetN(pin[2]) != n || mxGetM(pin[3]) != 3 || (size_t)mxGetN(pin[3]) != (size_t)3 * n)
        mexErrMsgIdAndTxt("vlg:mex3:eBV", "eB must be 3 x n and V_inv 3 x 3 x n");
    if (mxGetM(pin[4]) != 4 || (int)mxGetN(pin[4]) != m) mexErrMsgIdAndTxt("vlg:mex3:K", "K must be 4 x m");
    if (mxGetM(pin[7]) != 2 || (size_t)mxGetN(pin[7]) != (size_t)n * m || (int)mxGetM(pin[8]) != n || (int)mxGetN(pin[8]) != m)
        mexErrMsgIdAndTxt("vlg:mex3:X", "X must be 2 x n x m and visible n x m");
    dX[0] = 2; dX[1] = n; dX[2] = m;
    pout[0] = mxCreateDoubleMatrix(3, n, mxREAL);
    pout[1] = mxCreateDoubleMatrix(num_a, m, mxREAL);
    pout[2] = mxCreateDoubleMatrix(3, n, mxREAL);
    pout[3] = mxCreateNumericArray(3, dX, mxDOUBLE_CLASS, mxREAL);
    rc = vlg_ba_mex3_dense(m, n, num_a, mxGetPr(pin[0]), mxGetPr(pin[1]), mxGetPr(pin[2]), mxGetPr(pin[3]), mxGetPr(pin[4]),
                           mxGetPr(pin[5]), mxGetPr(pin[6]), mxGetPr(pin[7]), mxGetPr(pin[8]), mxGetPr(pout[0]),
                           mxGetPr(pout[1]), mxGetPr(pout[2]), mxGetPr(pout[3]));
    if (rc != VLG_BA_OK) mexErrMsgIdAndTxt("vlg:mex3:gpu", vlg_ba_last_error(0));
}
