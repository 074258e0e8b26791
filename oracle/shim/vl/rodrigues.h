/*
 * oracle/shim/vl/rodrigues.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of VLFeat's vl_rodrigues (forward map only).  VLFeat is a
 * third-party dependency of the reference that is NOT vendored under
 * /root/reference (Makefile:24,37 point at ../vlfeat, version unpinned); the
 * single call site on the hot path is toolbox/bundle/reproject_point.h:44,
 * vl_rodrigues(R, 0, a) -- i.e. the derivative output is a null pointer.
 * The published algorithm (Rodrigues' formula with a theta < 1e-6 identity
 * branch) and the operand order below are the ones SURVEY.md section 8(a5) pinned
 * against the disassembly of the reference's shipped
 * toolbox/bundle/mex_bundle_1_XABeUVWeAeB.mexglx.
 *
 * R_pt is 3x3 column-major: R(i,j) = R_pt[i + 3*j].
 */
#ifndef VLG_ORACLE_SHIM_VL_RODRIGUES_H
#define VLG_ORACLE_SHIM_VL_RODRIGUES_H

#include <math.h>

static void vl_rodrigues(double *R_pt, double *dR_pt, const double *om_pt)
{
    const double small = 1e-6;
    double th = sqrt(om_pt[0]*om_pt[0] + om_pt[1]*om_pt[1] + om_pt[2]*om_pt[2]);
    (void)dR_pt;
    if (th < small) {
        R_pt[0] = 1.0; R_pt[3] = 0.0; R_pt[6] = 0.0;
        R_pt[1] = 0.0; R_pt[4] = 1.0; R_pt[7] = 0.0;
        R_pt[2] = 0.0; R_pt[5] = 0.0; R_pt[8] = 1.0;
        return;
    }
    {
        double x = om_pt[0] / th;
        double y = om_pt[1] / th;
        double z = om_pt[2] / th;
        double xx = x*x, xy = x*y, xz = x*z;
        double yy = y*y, yz = y*z, zz = z*z;
        double sth  = sin(th);
        double cth  = cos(th);
        double mcth = 1.0 - cth;
        R_pt[0] = 1.0     - mcth*(yy+zz);
        R_pt[1] =   sth*z + mcth*xy;
        R_pt[2] = - sth*y + mcth*xz;
        R_pt[3] = - sth*z + mcth*xy;
        R_pt[4] = 1.0     - mcth*(zz+xx);
        R_pt[5] =   sth*x + mcth*yz;
        R_pt[6] =   sth*y + mcth*xz;
        R_pt[7] = - sth*x + mcth*yz;
        R_pt[8] = 1.0     - mcth*(xx+yy);
    }
}

#endif
