/*
 * mex_bundle_proj_1_XABeUVWeAeB.c -- GPU drop-in for the reference's projective stage-1 mex file
 * (toolbox/bundle/mex_bundle_proj_1_XABeUVWeAeB.c:88-300).  Same positional arguments:
 *
 *   [X_hat A B e U V W eA eB] = mex_bundle_proj_1_XABeUVWeAeB(a, b, X, visible)
 *     a 12xm (3x4 camera matrices, column-major), b 3xn, X 2xnxm, visible nxm   (reference :95-98)
 *     X_hat 2xnxm, A 2x12xnxm, B 2x3xnxm, e 2xnxm, U 12x12xm, V 3x3xn, W 12x3xnxm, eA 12xm, eB 3xn
 *                                                                            (reference :146-187)
 * All arithmetic happens in libvlgba.so; num_a = 12 selects the projective model there.
 */
#include "mex.h"
#include "vlg_ba.h"
#include "vlg_mex_state.h"

void mexFunction(int nout, mxArray *pout[], int nin, const mxArray *pin[])
{
    mwSize dX[3], dA[4], dB[4], dU[3], dV[3], dW[4];
    const int num_a = 12;
    int m, n, rc;
    (void)nout;
    vlg_mex_keep_state();
    if (nin != 4) mexErrMsgIdAndTxt("vlg:pmex1:nargin", "mex_bundle_proj_1_XABeUVWeAeB(a, b, X, visible): 4 inputs required");
    m = (int)mxGetN(pin[0]);
    n = (int)mxGetN(pin[1]);
    if ((int)mxGetM(pin[0]) != num_a) mexErrMsgIdAndTxt("vlg:pmex1:a", "a must be 12 x m");
    if (mxGetM(pin[1]) != 3) mexErrMsgIdAndTxt("vlg:pmex1:b", "b must be 3 x n");
    if (mxGetM(pin[2]) != 2 || (size_t)mxGetN(pin[2]) != (size_t)n * m) mexErrMsgIdAndTxt("vlg:pmex1:X", "X must be 2 x n x m");
    if ((int)mxGetM(pin[3]) != n || (int)mxGetN(pin[3]) != m) mexErrMsgIdAndTxt("vlg:pmex1:visible", "visible must be n x m");

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
    rc = vlg_ba_mex1_dense(m, n, num_a, 0, mxGetPr(pin[0]), mxGetPr(pin[1]), mxGetPr(pin[2]), mxGetPr(pin[3]),
                           mxGetPr(pout[0]), mxGetPr(pout[1]), mxGetPr(pout[2]), mxGetPr(pout[3]), mxGetPr(pout[4]),
                           mxGetPr(pout[5]), mxGetPr(pout[6]), mxGetPr(pout[7]), mxGetPr(pout[8]));
    if (rc != VLG_BA_OK) mexErrMsgIdAndTxt("vlg:pmex1:gpu", vlg_ba_last_error(0));
}
