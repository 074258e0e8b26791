/*
 * mex_bundle_2_Se_.c -- GPU drop-in for toolbox/bundle/mex_bundle_2_Se_.c:15-158.
 *
 *   [S e_] = mex_bundle_2_Se_(Y, W, U_, eA, eB)
 *     Y, W num_a x 3 x n x m, U_ num_a x num_a x m, eA num_a x m, eB 3xn   (reference :21-27)
 *     S (num_a*m)^2, e_ (num_a*m) x 1                                        (reference :59-66)
 *   m, n from eA, eB and num_a from Y, as the reference does (:53-57).
 */
#include "mex.h"
#include "vlg_ba.h"
#include "vlg_mex_state.h"

void mexFunction(int nout, mxArray *pout[], int nin, const mxArray *pin[])
{
    int m, n, num_a, rc;
    (void)nout;
    vlg_mex_keep_state();
    if (nin != 5) mexErrMsgIdAndTxt("vlg:mex2:nargin", "mex_bundle_2_Se_(Y, W, U_, eA, eB): 5 inputs required");
    m = (int)mxGetN(pin[3]);
    n = (int)mxGetN(pin[4]);
    num_a = (int)mxGetM(pin[0]);
    if (!(num_a == 6 || num_a == 7 || num_a == 10)) mexErrMsgIdAndTxt("vlg:mex2:num_a", "Y must have 6, 7 or 10 rows");
    if ((size_t)mxGetN(pin[0]) != (size_t)3 * n * m || (size_t)mxGetN(pin[1]) != (size_t)3 * n * m ||
        (int)mxGetM(pin[1]) != num_a)
        mexErrMsgIdAndTxt("vlg:mex2:YW", "Y and W must be num_a x 3 x n x m");
    if ((int)mxGetM(pin[2]) != num_a || (size_t)mxGetN(pin[2]) != (size_t)num_a * m) mexErrMsgIdAndTxt("vlg:mex2:U", "U_ must be num_a x num_a x m");
    if ((int)mxGetM(pin[3]) != num_a || mxGetM(pin[4]) != 3) mexErrMsgIdAndTxt("vlg:mex2:e", "eA must be num_a x m, eB 3 x n");
    pout[0] = mxCreateDoubleMatrix(num_a * m, num_a * m, mxREAL);
    pout[1] = mxCreateDoubleMatrix(num_a * m, 1, mxREAL);
    rc = vlg_ba_mex2_dense(m, n, num_a, mxGetPr(pin[0]), mxGetPr(pin[1]), mxGetPr(pin[2]), mxGetPr(pin[3]),
                           mxGetPr(pin[4]), mxGetPr(pout[0]), mxGetPr(pout[1]));
    if (rc != VLG_BA_OK) mexErrMsgIdAndTxt("vlg:mex2:gpu", vlg_ba_last_error(0));
}
