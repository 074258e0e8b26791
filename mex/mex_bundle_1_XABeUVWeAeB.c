/*
 * mex_bundle_1_XABeUVWeAeB.c -- GPU drop-in for the reference's mex file of the same name
 * (toolbox/bundle/mex_bundle_1_XABeUVWeAeB.c:72-337).  Same positional arguments:
 *
 *   [X_hat A B e U V W eA eB] = mex_bundle_1_XABeUVWeAeB(K, a, b, X, visible)
 *     K 4xm, a num_a x m, b 3xn, X 2xnxm, visible nxm (double)      (reference :76-83)
 *     X_hat 2xnxm, A 2 x num_a x n x m, B 2x3xnxm, e 2xnxm, U num_a x num_a x m, V 3x3xn,
 *     W num_a x 3 x n x m, eA num_a x m, eB 3xn                       (reference :136-175)
 *
 * All arithmetic happens in libvlgba.so (CUDA, sm_100a); this file only checks shapes (the
 * reference does not) and forwards pointers.  Build: mex -I../include mex_bundle_1_XABeUVWeAeB.c -L.. -lvlgba
 */
#include "mex.h"
#include "vlg_ba.h"
#include "vlg_mex_state.h"

void mexFunction(int nout, mxArray *pout[], int nin, const mxArray *pin[])
{
    mwSize dX[3], dA[4], dB[4], dU[3], dV[3], dW[4];
    int m, n, num_a, rc;
    (void)nout;
    vlg_mex_keep_state();
    if (nin != 5) mexErrMsgIdAndTxt("vlg:mex1:nargin", "mex_bundle_1_XABeUVWeAeB(K, a, b, X, visible): 5 inputs required");
    m = (int)mxGetN(pin[1]);
    n = (int)mxGetN(pin[2]);
    num_a = (int)mxGetM(pin[1]);
    if (!(num_a == 6 || num_a == 7 || num_a == 10)) mexErrMsgIdAndTxt("vlg:mex1:num_a", "a must have 6, 7 or 10 rows");
    if (mxGetM(pin[0]) != 4 || (int)mxGetN(pin[0]) != m) mexErrMsgIdAndTxt("vlg:mex1:K", "K must be 4 x m");
    if (mxGetM(pin[2]) != 3) mexErrMsgIdAndTxt("vlg:mex1:b", "b must be 3 x n");
    if (mxGetM(pin[3]) != 2 || (size_t)mxGetN(pin[3]) != (size_t)n * m) mexErrMsgIdAndTxt("vlg:mex1:X", "X must be 2 x n x m");
    if ((int)mxGetM(pin[4]) != n || (int)mxGetN(pin[4]) != m) mexErrMsgIdAndTxt("vlg:mex1:visible", "visible must be n x m");

    dX[0] = 2; dX[1] = n; dX[2] = m;
    dA[0] = 2; dA[1] = num_a; dA[2] = n; dA[3] = m;
    dB[0] = 2; dB[1] = 3; dB[2] = n; dB[3] = m;
    dU[0] = num_a; dU[1] = num_a; dU[2] = m;
    dV[0] = 3; dV[1] = 3; dV[2] = n;
    dW[0] = num_a; dW[1] = 3; dW[2] = n; dW[3] = m;
    pout[0] = mxCreateNumericArray(3, dX, mxDOUBLE_CLASS, mxREAL);
    pout[1] = mxCreateNumericArray(4, dA, mxDOUBLE_CLASS, mxREAL);
    pout[2] = mxCreateNumericArray(4, dB, mxDOUBLE_CLASS, mxREAL);
    pout[3] = mxCreateNumericArray(3, dX, mxDOUBLE_CLASS, mxREAL);
    pout[4] = mxCreateNumericArray(3, dU, mxDOUBLE_CLASS, mxREAL);
    pout[5] = mxCreateNumericArray(3, dV, mxDOUBLE_CLASS, mxREAL);
    pout[6] = mxCreateNumericArray(4, dW, mxDOUBLE_CLASS, mxREAL);
    pout[7] = mxCreateDoubleMatrix(num_a, m, mxREAL);
    pout[8] = mxCreateDoubleMatrix(3, n, mxREAL);
    rc = vlg_ba_mex1_dense(m, n, num_a, mxGetPr(pin[0]), mxGetPr(pin[1]), mxGetPr(pin[2]), mxGetPr(pin[3]),
                           mxGetPr(pin[4]), mxGetPr(pout[0]), mxGetPr(pout[1]), mxGetPr(pout[2]), mxGetPr(pout[3]),
                           mxGetPr(pout[4]), mxGetPr(pout[5]), mxGetPr(pout[6]), mxGetPr(pout[7]), mxGetPr(pout[8]));
    if (rc != VLG_BA_OK) mexErrMsgIdAndTxt("vlg:mex1:gpu", vlg_ba_last_error(0));
}
