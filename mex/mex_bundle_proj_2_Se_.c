/*
 * mex_bundle_proj_2_Se_.c -- GPU drop-in for toolbox/bundle/mex_bundle_proj_2_Se_.c:15-150.
 *
 *   [S e_] = mex_bundle_proj_2_Se_(Y, W, U_, eA, eB)
 *     Y, W 12x3xnxm, U_ 12x12xm, eA 12xm, eB 3xn                      (reference :22-26)
 *     S (12m)^2, e_ (12m) x 1                                          (reference :59-65)
 */
#include "mex.h"
#include "vlg_ba.h"
#include "vlg_mex_state.h"

void mexFunction(int nout, mxArray *pout[], int nin, const mxArray *pin[])
{
    const int num_a = 12;
    int m, n, rc;
    (void)nout;
    vlg_mex_keep_state();
    if (nin != 5) mexErrMsgIdAndTxt("vlg:pmex2:nargin", "mex_bundle_proj_2_Se_(Y, W, U_, eA, eB): 5 inputs required");
    m = (int)mxGetN(pin[3]);
    n = (int)mxGetN(pin[4]);
    if ((int)mxGetM(pin[0]) != num_a || (int)mxGetM(pin[1]) != num_a || (size_t)mxGetN(pin[0]) != (size_t)3 * n * m ||
        (size_t)mxGetN(pin[1]) != (size_t)3 * n * m)
        mexErrMsgIdAndTxt("vlg:pmex2:YW", "Y and W must be 12 x 3 x n x m");
    if ((int)mxGetM(pin[2]) != num_a || (size_t)mxGetN(pin[2]) != (size_t)num_a * m) mexErrMsgIdAndTxt("vlg:pmex2:U", "U_ must be 12 x 12 x m");
    if ((int)mxGetM(pin[3]) != num_a || mxGetM(pin[4]) != 3) mexErrMsgIdAndTxt("vlg:pmex2:e", "eA must be 12 x m, eB 3 x n");
    pout[0] = mxCreateDoubleMatrix(num_a * m, num_a * m, mxREAL);
    pout[1] = mxCreateDoubleMatrix(num_a * m, 1, mxREAL);
    rc = vlg_ba_mex2_dense(m, n, num_a, mxGetPr(pin[0]), mxGetPr(pin[1]), mxGetPr(pin[2]), mxGetPr(pin[3]),
                           mxGetPr(pin[4]), mxGetPr(pout[0]), mxGetPr(pout[1]));
    if (rc != VLG_BA_OK) mexErrMsgIdAndTxt("vlg:pmex2:gpu", vlg_ba_last_error(0));
}
