/*
 * vlg_mex_state.h -- the stage wrappers keep their GPU contexts between calls: libvlgba caches one
 * context for mex1/mex3 and one for mex2 per host thread (vlg_ba_mex{1,2,3}_dense); the mex file locks itself in memory so
 * that MATLAB's `clear mex` cannot unload the library under a live context, and frees the contexts when MATLAB exits
 * or the file is unlocked and cleared.  The reference's mex files are stateless (mex_bundle_1_XABeUVWeAeB.c:72-73);
 * results are identical either way, only the per-call set-up disappears.
 */
#ifndef VLG_MEX_STATE_H
#define VLG_MEX_STATE_H
#include "mex.h"
#include "vlg_ba.h"

static void vlg_mex_cleanup(void) { vlg_ba_dense_release(); }

static void vlg_mex_keep_state(void)
{
    if (!mexIsLocked()) {
        mexLock();
        mexAtExit(vlg_mex_cleanup);
    }
}
#endif
