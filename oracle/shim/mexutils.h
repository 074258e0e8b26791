/* oracle/shim/mexutils.h -- TEST INFRASTRUCTURE ONLY.
 * VLFeat's toolbox/mexutils.h is not vendored by the reference; the three
 * bundle mex files use it only to pull in mex.h
 * (mex_bundle_1_XABeUVWeAeB.c:9). */
#ifndef VLG_ORACLE_SHIM_MEXUTILS_H
#define VLG_ORACLE_SHIM_MEXUTILS_H
#include "mex.h"
#endif
