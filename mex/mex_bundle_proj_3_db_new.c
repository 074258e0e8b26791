/*
 * mex_bundle_proj_3_db_new.c -- GPU drop-in for toolbox/bundle/mex_bundle_proj_3_db_new.c:34-170.
 *
 *   [db a_new b_new X_hat] = mex_bundle_proj_3_db_new(W, da, eB, V_inv, a, b, X, visible)
 *     W 12x3xnxm, da (12m) x 1, eB 3xn, V_inv 3x3xn, a 12xm, b 3xn, X 2xnxm, visible nxm  (reference :42-49)
 *     db 3xn, a_new 12xm, b_new 3xn, X_hat 2xnxm                                          (reference :86-100)
 *   Like the reference, only the first six camera parameters enter the back-substitution (:117-131).
 */
#include "mex.h"
#include "vlg_ba.h"
#include "vlg_mex_state.h"

void mexFunction(int nout, mxArray *pout[], int nin, const mxArray *pin[])
{
    mwSize dX[3];
    const int num_a = 12;
    int m, n, rc;
    (void)nout;
    vlg_mex_keep_state();
    if (nin != 8) mexErrMsgIdAndTxt("vlg:pmex3:nargin", "mex_bundle_proj_3_db_new(W, da, eB, V_inv, a, b, X, visible): 8 inputs required");
    m = (int)mxGetN(pin[4]);
    n = (int)mxGetN(pin[5]);
    if ((int)mxGetM(pin[0]) != num_a || (int)mxGetM(pin[4]) != num_a) mexErrMsgIdAndTxt("vlg:pmex3:num_a", "W and a must have 12 rows");
    if ((size_t)mxGetN(pin[0]) != (size_t)3 * n * m) mexErrMsgIdAndTxt("vlg:pmex3:W", "W must be 12 x 3 x n x m");
    if ((size_t)(mxGetM(pin[1]) * mxGetN(pin[1])) != (size_t)num_a * m) mexErrMsgIdAndTxt("vlg:pmex3:da", "da must have 12*m entries");
    if (mxGetM(pin[2]) != 3 || (int)mxGetN(pin[2]) != n || mxGetM(pin[3]) != 3 || (size_t)mxGetN(pin[3]) != (size_t)3 * n)
        mexErrMsgIdAndTxt("vlg:pmex3:eBV", "eB must be 3 x n and V_inv 3 x 3 x n");
    if (mxGetM(pin[6]) != 2 || (size_t)mxGetN(pin[6]) != (size_t)n * m || (int)mxGetM(pin[7]) != n || (int)mxGetN(pin[7]) != m)
        mexErrMsgIdAndTxt("vlg:pmex3:X", "X must be 2 x n x m and visible n x m");
    dX[0] = 2; dX[1] = n; dX[2] = m;
    pout[0] = mxCreateDoubleMatrix(3, n, mxREAL);
    pout[1] = mxCreateDoubleMatrix(num_a, m, mxREAL);
    pout[2] = mxCreateDoubleMatrix(3, n, mxREAL);
    pout[3] = mxCreateNumericArray(3, dX, mxDOUBLE_CLASS, mxREAL);
    rc = vlg_ba_mex3_dense(m, n, num_a, mxGetPr(pin[0]), mxGetPr(pin[1]), mxGetPr(pin[2]), mxGetPr(pin[3]), 0,
                           mxGetPr(pin[4]), mxGetPr(pin[5]), mxGetPr(pin[6]), mxGetPr(pin[7]), mxGetPr(pout[0]),
                           mxGetPr(pout[1]), mxGetPr(pout[2]), mxGetPr(pout[3]));
    if (rc != VLG_BA_OK) mexErrMsgIdAndTxt("vlg:pmex3:gpu", vlg_ba_last_error(0));
}
