// vlg_ba_dense.inl -- included inside the extern "C" block of vlg_ba_host.inl.
//
// Entry points with exactly the dense argument layout of the reference's three mexFunctions, for the
// drop-in mex wrappers in mex/ (bundle_euclid.m then runs unmodified), plus the whole of
// bundle_euclid.m behind one call.  Each compacts the dense n x m arrays to the visible cells, runs
// the same CUDA kernels as the fused path, and scatters back.
//
// The three stage entries keep their contexts between calls (bundle_euclid.m calls mex1, mex2,
// mex3 once per trial step on the same visibility pattern): one context for mex1/mex3 and one for mex2 per host
// thread, re-used while (m, n, num_a) and the list of contributing cells stay the same -- then a call only uploads
// its operands instead of creating a CUDA stream, compacting, building the CSR / block structure and cudaMalloc-ing
// everything three times per trial step.  vlg_ba_dense_release() drops them (the mex wrappers register it with
// mexAtExit and mexLock themselves); vlg_ba_dense_cache_stats() counts hits and rebuilds for the tests.

}  // extern "C" (the cache helpers have C++ linkage)

namespace {

struct DenseSlot {
    vlg_ba_ctx* ctx = nullptr;
    int m = 0, n = 0, na = 0;      // no destructor: at process exit the CUDA runtime may be gone before thread-local storage is
};
thread_local DenseSlot g_dense13, g_dense2;
thread_local int64_t g_dense_hits = 0, g_dense_builds = 0;

void dense_drop(DenseSlot& s)
{
    if (s.ctx) vlg_ba_destroy(s.ctx);
    s.ctx = nullptr;
}

// the context of a dense (K, a, b, X, visible) call: re-used when the visible cells are the same list as last time
int dense_acquire13(int m, int n, int num_a, const double* K, const double* a, const double* b, const double* X,
                    const double* visible, vlg_ba_ctx** out)
{
    *out = nullptr;
    if (m <= 0 || n < 0 || !a || (n > 0 && (!b || !X || !visible))) return fail(nullptr, VLG_BA_EINVAL, "dense entry: NULL argument");
    std::vector<double> xy;
    std::vector<int32_t> pt, cam;
    for (int j = 0; j < m; j++)
        for (int i = 0; i < n; i++) {
            const size_t c = (size_t)i + (size_t)n * j;
            if (visible[c] != 0.0) {                               // mex_bundle_1_XABeUVWeAeB.c:196
                xy.push_back(X[2 * c]); xy.push_back(X[2 * c + 1]);
                pt.push_back(i); cam.push_back(j);
            }
        }
    DenseSlot& s = g_dense13;
    vlg_ba_ctx* ctx = s.ctx;
    const bool proj = num_a == kNaProjective;
    if (ctx && s.m == m && s.n == n && s.na == num_a && ctx->h_obs_pt == pt && ctx->h_obs_cam == cam) {
        g_dense_hits++;
        CU(cudaSetDevice(ctx->device));
        ctx->h_obs_xy = xy;
        CHK(upload(ctx, (double*)ctx->obs_xy, xy.data(), xy.size()));
        if (!proj && K) { ctx->h_K.assign(K, K + 4 * (size_t)m); CHK(upload(ctx, ctx->K4, K, 4 * (size_t)m)); }
        CHK(vlg_ba_set_state(ctx, a, b, -1.0, -1.0));               // synchronises: the host vectors above may go
        *out = ctx;
        return VLG_BA_OK;
    }
    dense_drop(s);
    g_dense_builds++;
    vlg_ba_opts o;
    vlg_ba_opts_default(&o);
    // num_a = 12 (and K = NULL) is the projective model: mex_bundle_proj_1_XABeUVWeAeB(a, b, X, visible)
    if (proj) o.model = VLG_BA_MODEL_PROJECTIVE; else o.num_variableK = num_a - 6;
    o.order = VLG_BA_ORDER_REFERENCE;            // U, eA in the reference's exact accumulation order
    o.solver = VLG_BA_SOLVER_PCG;                // no block structure needed for stage 1 / stage 3
    o.pcg_cluster = 0;
    int r = vlg_ba_create(&o, &ctx);
    if (r != VLG_BA_OK) return r;
    r = build_problem(ctx, m, n, K, a, b, (int64_t)pt.size(), xy.data(), pt.data(), cam.data(), nullptr);
    if (r != VLG_BA_OK) {
        snprintf(g_create_error, sizeof(g_create_error), "%s", ctx->err);
        vlg_ba_destroy(ctx);
        return r;
    }
    s.ctx = ctx; s.m = m; s.n = n; s.na = num_a;
    *out = ctx;
    return VLG_BA_OK;
}

// a failed call leaves nothing cached (the context may be in an undefined state)
int dense_fail13(vlg_ba_ctx* ctx, int r)
{
    if (ctx) snprintf(g_create_error, sizeof(g_create_error), "%s", ctx->err);
    dense_drop(g_dense13);
    return r;
}

}  // namespace

extern "C" {

void vlg_ba_dense_release(void)
{
    dense_drop(g_dense13);
    dense_drop(g_dense2);
}

void vlg_ba_dense_cache_stats(int64_t* hits, int64_t* builds)
{
    if (hits) *hits = g_dense_hits;
    if (builds) *builds = g_dense_builds;
}

// mex_bundle_1_XABeUVWeAeB(K, a, b, X, visible) -> X_hat A B e U V W eA eB
// (mex_bundle_1_XABeUVWeAeB.c:72-337; outputs may be NULL).
int vlg_ba_mex1_dense(int m, int n, int num_a, const double* K, const double* a, const double* b, const double* X,
                      const double* visible, double* X_hat, double* A, double* B, double* e, double* U, double* V,
                      double* W, double* eA, double* eB)
{
    vlg_ba_ctx* ctx = nullptr;
    int r = dense_acquire13(m, n, num_a, K, a, b, X, visible, &ctx);
    if (r != VLG_BA_OK) return r;
    const size_t no = (size_t)ctx->nobs, na = (size_t)num_a, nm = (size_t)n * m;
    std::vector<double> sX(2 * no), sA(2 * na * no), sB(6 * no), se(2 * no), sW(3 * na * no);
    if (r == VLG_BA_OK) r = vlg_ba_get_jacobians(ctx, sX.data(), sA.data(), sB.data(), se.data());
    if (r == VLG_BA_OK) r = vlg_ba_stage1(ctx, nullptr);
    if (r == VLG_BA_OK) r = vlg_ba_get_blocks(ctx, U, V, sW.data(), eA, eB);
    if (r == VLG_BA_OK) {
        // invisible cells: X_hat = X, everything else zero (mex_bundle_1_XABeUVWeAeB.c:226-252)
        if (X_hat) memcpy(X_hat, X, sizeof(double) * 2 * nm);
        if (A) memset(A, 0, sizeof(double) * 2 * na * nm);
        if (B) memset(B, 0, sizeof(double) * 6 * nm);
        if (e) memset(e, 0, sizeof(double) * 2 * nm);
        if (W) memset(W, 0, sizeof(double) * 3 * na * nm);
        for (size_t t = 0; t < no; t++) {
            const size_t c = (size_t)ctx->h_obs_pt[t] + (size_t)n * ctx->h_obs_cam[t];
            if (X_hat) { X_hat[2 * c] = sX[2 * t]; X_hat[2 * c + 1] = sX[2 * t + 1]; }
            if (A) memcpy(A + 2 * na * c, sA.data() + 2 * na * t, sizeof(double) * 2 * na);
            if (B) memcpy(B + 6 * c, sB.data() + 6 * t, sizeof(double) * 6);
            if (e) { e[2 * c] = se[2 * t]; e[2 * c + 1] = se[2 * t + 1]; }
            if (W) memcpy(W + 3 * na * c, sW.data() + 3 * na * t, sizeof(double) * 3 * na);
        }
    } else {
        return dense_fail13(ctx, r);
    }
    return r;
}

// mex_bundle_2_Se_(Y, W, U_, eA, eB) -> S e_   (mex_bundle_2_Se_.c:15-158)
int vlg_ba_mex2_dense(int m, int n, int num_a, const double* Y, const double* W, const double* U_, const double* eA,
                      const double* eB, double* S, double* e_)
{
    const size_t na = (size_t)num_a, nw = 3 * na;
    // cells that can contribute: a non-zero Y_ij or W_ij block
    std::vector<int32_t> pt, cam;
    std::vector<double> sY, sW;
    for (int j = 0; j < m; j++)
        for (int i = 0; i < n; i++) {
            const size_t c = (size_t)i + (size_t)n * j;
            bool nz = false;
            for (size_t k = 0; k < nw && !nz; k++) nz = (Y[nw * c + k] != 0.0) || (W[nw * c + k] != 0.0);
            if (nz) {
                pt.push_back(i); cam.push_back(j);
                sY.insert(sY.end(), Y + nw * c, Y + nw * c + nw);
                sW.insert(sW.end(), W + nw * c, W + nw * c + nw);
            }
        }
    const size_t no = pt.size();
    DenseSlot& slot = g_dense2;
    vlg_ba_ctx* ctx = slot.ctx;
    int r = VLG_BA_OK;
    if (ctx && slot.m == m && slot.n == n && slot.na == num_a && ctx->h_obs_pt == pt && ctx->h_obs_cam == cam) {
        g_dense_hits++;                          // same contributing cells as last time: the block structure of S stands
    } else {
        dense_drop(slot);
        g_dense_builds++;
        vlg_ba_opts o;
        vlg_ba_opts_default(&o);
        if (num_a == kNaProjective) o.model = VLG_BA_MODEL_PROJECTIVE; else o.num_variableK = num_a - 6;
        o.solver = VLG_BA_SOLVER_CHOL;           // builds the block structure of S
        r = vlg_ba_create(&o, &ctx);
        if (r != VLG_BA_OK) return r;
        std::vector<double> K1(4 * (size_t)m, 1.0), a0(na * m, 0.0), b0(3 * (size_t)std::max(n, 1), 0.0), xy(2 * std::max<size_t>(no, 1), 0.0);
        r = vlg_ba_set_problem_sparse(ctx, m, n, K1.data(), a0.data(), b0.data(), (int64_t)no, xy.data(), pt.data(), cam.data(), nullptr);
        if (r != VLG_BA_OK) { snprintf(g_create_error, sizeof(g_create_error), "%s", ctx->err); vlg_ba_destroy(ctx); return r; }
        slot.ctx = ctx; slot.m = m; slot.n = n; slot.na = num_a;
    }
    double* dY = nullptr;
    const int N = num_a * m;
    auto body = [&]() -> int {
        CU(cudaSetDevice(ctx->device));
        CU(cudaMalloc(&dY, sizeof(double) * std::max<size_t>(nw * no, 1)));
        CHK(upload(ctx, dY, sY.data(), nw * no));
        CHK(upload(ctx, ctx->W, sW.data(), nw * no));
        CHK(upload(ctx, ctx->Ud, U_, na * na * m));
        CHK(upload(ctx, ctx->eA, eA, na * m));
        CHK(upload(ctx, ctx->eB, eB, 3 * (size_t)n));
        const int Np = ctx->Np;
        CU(cudaMemsetAsync(ctx->S, 0, sizeof(double) * (size_t)Np * Np, ctx->stream));
        const int g = cdiv(ctx->nblocks, kWarpsPerBlock), th = kWarpsPerBlock * 32, g2 = cdiv((int64_t)N, 128);
#define VLG_MEX2(NA_)                                                                                                      \
    do {                                                                                                                  \
        k_schur_blocks<NA_><<<g, th, 0, ctx->stream>>>((int)ctx->nblocks, Np, 1, ctx->blk_j, ctx->blk_k, ctx->blk_ptr,     \
                                                       ctx->pairs, ctx->obs_pt, ctx->W, ctx->Vinv, ctx->Ud, ctx->S, dY);   \
        k_ebar_from_Y<NA_><<<g2, 128, 0, ctx->stream>>>(m, ctx->cam_ptr, ctx->obs_pt, dY, ctx->eA, ctx->eB, ctx->ebar);    \
    } while (0)
        if (num_a == 6) VLG_MEX2(6); else if (num_a == 7) VLG_MEX2(7); else if (num_a == 10) VLG_MEX2(10); else VLG_MEX2(12);
#undef VLG_MEX2
        ctx->launches += 2;
        CU(cudaGetLastError());
        if (S) CU(cudaMemcpy2DAsync(S, sizeof(double) * N, ctx->S, sizeof(double) * (size_t)Np, sizeof(double) * N, N,
                                    cudaMemcpyDeviceToHost, ctx->stream));
        CHK(download(ctx, e_, ctx->ebar, (size_t)N));
        CU(cudaStreamSynchronize(ctx->stream));
        return VLG_BA_OK;
    };
    if (r == VLG_BA_OK) r = body();
    if (dY) cudaFree(dY);
    if (r != VLG_BA_OK) { snprintf(g_create_error, sizeof(g_create_error), "%s", ctx->err); dense_drop(slot); }
    return r;
}

// mex_bundle_3_db_new(W, da, eB, V_inv, K, a, b, X, visible) -> db a_new b_new X_hat
// (mex_bundle_3_db_new.c:12-170)
int vlg_ba_mex3_dense(int m, int n, int num_a, const double* W, const double* da, const double* eB, const double* Vinv,
                      const double* K, const double* a, const double* b, const double* X, const double* visible,
                      double* db, double* a_new, double* b_new, double* X_hat)
{
    // num_a = 12 (and K = NULL): mex_bundle_proj_3_db_new(W, da, eB, V_inv, a, b, X, visible)
    vlg_ba_ctx* ctx = nullptr;
    int r = dense_acquire13(m, n, num_a, K, a, b, X, visible, &ctx);
    if (r != VLG_BA_OK) return r;
    const size_t na = (size_t)num_a, nw = 3 * na, nm = (size_t)n * m;
    double* dXh = nullptr;
    auto body = [&]() -> int {
        const size_t no = (size_t)ctx->nobs;
        std::vector<double> sW(nw * no), sX(2 * no);
        for (size_t t = 0; t < no; t++) {
            const size_t c = (size_t)ctx->h_obs_pt[t] + (size_t)n * ctx->h_obs_cam[t];
            memcpy(sW.data() + nw * t, W + nw * c, sizeof(double) * nw);
        }
        CU(cudaSetDevice(ctx->device));
        CU(cudaMalloc(&dXh, sizeof(double) * std::max<size_t>(2 * no, 1)));
        CHK(upload(ctx, ctx->W, sW.data(), nw * no));
        CHK(upload(ctx, ctx->da, da, na * m));
        CHK(upload(ctx, ctx->eB, eB, 3 * (size_t)n));
        CHK(upload(ctx, ctx->Vinv, Vinv, 9 * (size_t)n));
        CU(cudaMemsetAsync(ctx->eA, 0, sizeof(double) * na * m, ctx->stream));
        ctx->xhat_out = dXh;
        ctx->s1_valid = true; ctx->s2_valid = true;
        double nc = 0, dn = 0;
        CHK(do_stage3(ctx, 0.0, &nc, &dn));
        ctx->xhat_out = nullptr;
        CHK(vlg_ba_get_update(ctx, db, a_new, b_new));
        CHK(download(ctx, sX.data(), dXh, 2 * no));
        CU(cudaStreamSynchronize(ctx->stream));
        if (X_hat) {
            memcpy(X_hat, X, sizeof(double) * 2 * nm);     // invisible: X_hat = X (mex_bundle_3_db_new.c:158-163)
            for (size_t t = 0; t < no; t++) {
                const size_t c = (size_t)ctx->h_obs_pt[t] + (size_t)n * ctx->h_obs_cam[t];
                X_hat[2 * c] = sX[2 * t]; X_hat[2 * c + 1] = sX[2 * t + 1];
            }
        }
        return VLG_BA_OK;
    };
    if (r == VLG_BA_OK) r = body();
    ctx->xhat_out = nullptr;
    if (dXh) cudaFree(dXh);
    if (r != VLG_BA_OK) return dense_fail13(ctx, r);
    ctx->s1_valid = ctx->s2_valid = ctx->s3_valid = false;      // the stage flags were forced for this call only
    return r;
}

// [K_ Te_ w_ Xe_ error_] = bundle_euclid(K, Te, w, Xe, x, ...)  (bundle_euclid.m:1-269) behind
// one call: K 4xm, Te 3xm, w 3xm, Xe 4xn, x 3xnxm, visible nxm (NULL = derive from x as
// bundle_euclid.m:50), pivot m (NULL = none); opts carries the parsed option strings.
int vlg_ba_bundle_euclid(const vlg_ba_opts* opts, int m, int n, const double* K, const double* Te, const double* w,
                         const double* Xe, const double* x, const double* visible, const double* pivot, double* K_,
                         double* Te_, double* w_, double* Xe_, double* error_, int* n_error)
{
    vlg_ba_opts o;
    if (opts) o = *opts; else vlg_ba_opts_default(&o);
    const int nk = o.num_variableK, na = 6 + nk;
    if (m <= 0 || n < 0 || !K || !Te || !w || !Xe || !x) return fail(nullptr, VLG_BA_EINVAL, "bundle_euclid: NULL argument");
    std::vector<double> a((size_t)na * m), b(3 * (size_t)n), X(2 * (size_t)n * m), vis((size_t)n * m), x4((size_t)n);
    for (int j = 0; j < m; j++) {                                   // bundle_euclid.m:89-96
        for (int k = 0; k < 3; k++) { a[(size_t)na * j + k] = w[3 * (size_t)j + k]; a[(size_t)na * j + 3 + k] = Te[3 * (size_t)j + k]; }
        if (nk == 1) a[(size_t)na * j + 6] = K[4 * (size_t)j];
        if (nk == 4) for (int k = 0; k < 4; k++) a[(size_t)na * j + 6 + k] = K[4 * (size_t)j + k];
    }
    for (int i = 0; i < n; i++) {                                   // :99
        for (int k = 0; k < 3; k++) b[3 * (size_t)i + k] = Xe[4 * (size_t)i + k];
        x4[i] = Xe[4 * (size_t)i + 3];
    }
    for (size_t c = 0; c < (size_t)n * m; c++) {                    // :50, :81, :102
        X[2 * c] = x[3 * c]; X[2 * c + 1] = x[3 * c + 1];
        vis[c] = visible ? (visible[c] != 0.0 ? 1.0 : 0.0) : ((x[3 * c] != 0.0 || x[3 * c + 1] != 0.0) ? 1.0 : 0.0);
    }
    vlg_ba_ctx* ctx = nullptr;
    int r = vlg_ba_create(&o, &ctx);
    if (r != VLG_BA_OK) return r;
    r = vlg_ba_set_problem_dense(ctx, m, n, K, a.data(), b.data(), X.data(), vis.data(), pivot);
    if (r == VLG_BA_OK) r = vlg_ba_solve(ctx, K_, Te_, w_, Xe_, x4.data(), error_, n_error);
    if (r != VLG_BA_OK) snprintf(g_create_error, sizeof(g_create_error), "%s", ctx->err);
    vlg_ba_destroy(ctx);
    return r;
}

// The same whole-loop entry on an observation LIST instead of the dense 3 x n x m / n x m arrays
// (bundle_euclid.m:9,18,50,81 make every caller materialise n x m cells; at Venice
// shape that is 43 GB of x and 14 GB of visibility for 5 M observations).  obs_* in the reference's
// traversal order, ascending i + n*j, 0-based indices.
int vlg_ba_bundle_euclid_sparse(const vlg_ba_opts* opts, int m, int n, const double* K, const double* Te, const double* w,
                                const double* Xe, int64_t nobs, const double* obs_xy, const int32_t* obs_pt,
                                const int32_t* obs_cam, const double* pivot, double* K_, double* Te_, double* w_,
                                double* Xe_, double* error_, int* n_error)
{
    vlg_ba_opts o;
    if (opts) o = *opts; else vlg_ba_opts_default(&o);
    const int nk = o.num_variableK, na = 6 + nk;
    if (m <= 0 || n < 0 || nobs < 0 || !K || !Te || !w || !Xe || (nobs > 0 && (!obs_xy || !obs_pt || !obs_cam)))
        return fail(nullptr, VLG_BA_EINVAL, "bundle_euclid_sparse: NULL argument");
    std::vector<double> a((size_t)na * m), b(3 * (size_t)n), x4((size_t)n);
    for (int j = 0; j < m; j++) {                                   // bundle_euclid.m:89-96
        for (int k = 0; k < 3; k++) { a[(size_t)na * j + k] = w[3 * (size_t)j + k]; a[(size_t)na * j + 3 + k] = Te[3 * (size_t)j + k]; }
        if (nk == 1) a[(size_t)na * j + 6] = K[4 * (size_t)j];
        if (nk == 4) for (int k = 0; k < 4; k++) a[(size_t)na * j + 6 + k] = K[4 * (size_t)j + k];
    }
    for (int i = 0; i < n; i++) {                                   // :99
        for (int k = 0; k < 3; k++) b[3 * (size_t)i + k] = Xe[4 * (size_t)i + k];
        x4[i] = Xe[4 * (size_t)i + 3];
    }
    vlg_ba_ctx* ctx = nullptr;
    int r = vlg_ba_create(&o, &ctx);
    if (r != VLG_BA_OK) return r;
    r = vlg_ba_set_problem_sparse(ctx, m, n, K, a.data(), b.data(), nobs, obs_xy, obs_pt, obs_cam, pivot);
    if (r == VLG_BA_OK) r = vlg_ba_solve(ctx, K_, Te_, w_, Xe_, x4.data(), error_, n_error);
    if (r != VLG_BA_OK) snprintf(g_create_error, sizeof(g_create_error), "%s", ctx->err);
    vlg_ba_destroy(ctx);
    return r;
}

// bundle_projective.m:1-229 in one call (dense interface, like vlg_ba_bundle_euclid)
int vlg_ba_bundle_projective(const vlg_ba_opts* opts, int m, int n, const double* Pp, const double* Xp, const double* x,
                             const double* visible, double* Pp_, double* Xp_, double* error_, int* n_error)
{
    vlg_ba_opts o;
    if (opts) o = *opts; else vlg_ba_opts_default(&o);
    o.model = VLG_BA_MODEL_PROJECTIVE;
    if (m <= 0 || n < 0 || !Pp || !Xp || !x) return fail(nullptr, VLG_BA_EINVAL, "bundle_projective: NULL argument");
    // a(1:12,j) = reshape(Pp(:,:,j),12,1) is the memory layout of Pp itself (bundle_projective.m:69-72)
    std::vector<double> b(3 * (size_t)n), X(2 * (size_t)n * m), vis((size_t)n * m);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) b[3 * (size_t)i + k] = Xp[4 * (size_t)i + k];                 // :75
    for (size_t c = 0; c < (size_t)n * m; c++) {                                                    // :38, :61, :78
        X[2 * c] = x[3 * c]; X[2 * c + 1] = x[3 * c + 1];
        vis[c] = visible ? (visible[c] != 0.0 ? 1.0 : 0.0) : ((x[3 * c] != 0.0 || x[3 * c + 1] != 0.0) ? 1.0 : 0.0);
    }
    vlg_ba_ctx* ctx = nullptr;
    int r = vlg_ba_create(&o, &ctx);
    if (r != VLG_BA_OK) return r;
    r = vlg_ba_set_problem_dense(ctx, m, n, nullptr, Pp, b.data(), X.data(), vis.data(), nullptr);
    if (r == VLG_BA_OK) {
        r = vlg_ba_lm_reset(ctx, nullptr, nullptr);
        while (r == VLG_BA_OK && vlg_ba_lm_continue(ctx)) {
            const int it = ctx->iter;
            vlg_ba_trial_info info;
            r = do_trial(ctx, &info);
            if (r == VLG_BA_OK && info.accepted && o.verbose)
                printf("iter %d: error= %.5g -> %.5g\n", it, ctx->err_hist[it - 1], ctx->err_hist[it]);
        }
    }
    if (r == VLG_BA_OK) {
        std::vector<double> bb(3 * (size_t)n);
        r = vlg_ba_get_state(ctx, Pp_, bb.data(), nullptr, nullptr, nullptr, nullptr);      // Pp_ = reshape(a, 3, 4, m): :213-217
        if (r == VLG_BA_OK && Xp_)
            for (int i = 0; i < n; i++) {                                                   // Xp_ = [b; Xp(4,:)]: :219
                for (int k = 0; k < 3; k++) Xp_[4 * (size_t)i + k] = bb[3 * (size_t)i + k];
                Xp_[4 * (size_t)i + 3] = Xp[4 * (size_t)i + 3];
            }
        const int ne = ctx->err_hist.empty() ? 0 : ctx->iter;
        if (error_) for (int k = 0; k < ne; k++) error_[k] = ctx->err_hist[k];
        if (n_error) *n_error = ne;
    }
    if (r != VLG_BA_OK) snprintf(g_create_error, sizeof(g_create_error), "%s", ctx->err);
    vlg_ba_destroy(ctx);
    return r;
}

