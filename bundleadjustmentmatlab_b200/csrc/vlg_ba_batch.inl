// vlg_ba_batch.inl -- included inside the extern "C" block of vlg_ba_host.inl.
//
// vlg_ba_solve_cameras_independent: B independent single-camera motion-only bundle adjustments in one context
// (SURVEY.md 8f N2).  The reference runs them one bundle_euclid call at a time (estimate_camera.m:247-253:
// bundle_euclid(K, T, Omega, X, x, 'fix_structure', 'fix_calibration', 'visibility', inlier') with m = 1, once per added
// camera in incr_reconstruction.m:230); each call here would cost a context, a problem set-up and ~30 launches per trial
// step for a 6 x 6 system.  With the structure fixed nothing couples the cameras: S is block diagonal (S_jj = U*_j,
// e_j = eA_j), db = 0, the cost is a sum of per-camera costs.  So the problem is set up ONCE with all B cameras
// (camera j sees only its own points) and the B Levenberg-Marquardt loops of bundle_euclid.m:111-249 advance side by side,
// one ROUND = one trial step of every camera that has not stopped yet: each camera has its own lambda, nu, iter, iter2,
// error_ history, accept decision and stop rule, exactly as if bundle_euclid had been called on it alone.

namespace {

template <int NA>
int run_independent_round(vlg_ba_ctx* ctx, const double* d_lam, const unsigned char* d_active, double* d_cost_cam /* [2][m]: old, new */,
                          double* d_denom, double* d_rtab9)
{
    const int m = ctx->m, N = NA * m;
    // stage 1 at the current a (mex1): U_j, eA_j; the per-observation cost of the current state
    CHK(run_stage1<NA>(ctx, nullptr, nullptr, nullptr, nullptr));
    for (int j = 0; j < m; j++)
        for (int q = 0; q < 9; q++) ctx->h_rtab_new[(size_t)9 * j + q] = ctx->h_rtab[(size_t)36 * j + q];
    CHK(upload(ctx, d_rtab9, ctx->h_rtab_new.data(), (size_t)9 * m));
    k_new_cost<NA><<<cdiv(ctx->nobs, 256), 256, 0, ctx->stream>>>(ctx->nobs, ctx->obs_xy, ctx->obs_pt, ctx->obs_cam, ctx->K4, ctx->a, ctx->b, d_rtab9,
                                                                 ctx->cost_obs, nullptr);
    k_cam_seg_sum<<<cdiv(m, 4), 128, 0, ctx->stream>>>(m, ctx->cam_ptr, ctx->cost_obs, d_cost_cam);
    // damping, da_j = pinv(U*_j) eA_j
    k_damp_U_vec<NA><<<cdiv((int64_t)m * NA * NA, 256), 256, 0, ctx->stream>>>(m, d_lam, ctx->U, ctx->Ud);
    k_cam_solve_diag<NA><<<cdiv(m, 64), 64, 0, ctx->stream>>>(m, ctx->Ud, ctx->eA, d_lam, d_active, ctx->da, d_denom);
    ctx->launches += 4;
    CU(cudaGetLastError());
    // a_new = a + da on the host (same IEEE add as mex_bundle_3_db_new.c:137-140) and its rotation matrices from the host libm
    CHK(download(ctx, ctx->h_da.data(), ctx->da, (size_t)N));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int t = 0; t < N; t++) ctx->h_a_new[t] = ctx->h_a[t] + ctx->h_da[t];
    CHK(upload(ctx, ctx->a_new, ctx->h_a_new.data(), (size_t)N));
    if (ctx->opt.rtable == VLG_BA_RTABLE_HOST_LIBM) {
        rtab_host(m, NA, ctx->h_a_new.data(), 1, ctx->h_rtab_new.data());
        CHK(upload(ctx, ctx->rtab_new, ctx->h_rtab_new.data(), (size_t)9 * m));
    } else {
        k_rtab<NA><<<cdiv(m, 128), 128, 0, ctx->stream>>>(m, ctx->a_new, 1, ctx->rtab_new);
        ctx->launches++;
    }
    // new cost per camera (mex3 with db = 0: b_new = b)
    k_new_cost<NA><<<cdiv(ctx->nobs, 256), 256, 0, ctx->stream>>>(ctx->nobs, ctx->obs_xy, ctx->obs_pt, ctx->obs_cam, ctx->K4, ctx->a_new, ctx->b,
                                                                 ctx->rtab_new, ctx->cost_obs, nullptr);
    k_cam_seg_sum<<<cdiv(m, 4), 128, 0, ctx->stream>>>(m, ctx->cam_ptr, ctx->cost_obs, d_cost_cam + m);
    ctx->launches += 2;
    CU(cudaGetLastError());
    return VLG_BA_OK;
}

}  // namespace

extern "C" int vlg_ba_solve_cameras_independent(vlg_ba_ctx* ctx, double* a_out, double* error_, int* n_error, int* rounds)
{
    if (!ctx || !ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    const vlg_ba_opts& o = ctx->opt;
    if (!o.fix_structure || o.model != VLG_BA_MODEL_EUCLID)
        return fail(ctx, VLG_BA_EINVAL, "solve_cameras_independent needs opts.fix_structure = 1 (Euclidean model): only then do the cameras decouple");
    if (ctx->nranks > 1) return fail(ctx, VLG_BA_EINVAL, "solve_cameras_independent is a single-GPU entry");
    CU(cudaSetDevice(ctx->device));
    const int m = ctx->m, na = ctx->na, N = na * m, maxit = o.max_iter;
    std::vector<double> lam((size_t)m, o.lambda0), nu((size_t)m, o.nu0), err((size_t)m * (maxit + 1), 0.0), nvis((size_t)m, 0.0);
    std::vector<int> it((size_t)m, 1), it2((size_t)m, 0);
    std::vector<unsigned char> active((size_t)m, 1);
    {
        std::vector<int> cnt((size_t)m, 0);
        for (int64_t t = 0; t < ctx->nobs; t++) cnt[(size_t)ctx->h_obs_cam[(size_t)t]]++;
        for (int j = 0; j < m; j++) nvis[(size_t)j] = (double)cnt[(size_t)j];     // num_vis of camera j's own call (bundle_euclid.m:82)
    }
    double *d_lam = nullptr, *d_cost = nullptr, *d_denom = nullptr, *d_rtab9 = nullptr;
    unsigned char* d_active = nullptr;
    std::vector<double> h_cost((size_t)2 * m), h_denom((size_t)m);
    int nround = 0;
    auto body = [&]() -> int {
        CU(cudaMalloc(&d_lam, sizeof(double) * m)); CU(cudaMalloc(&d_cost, sizeof(double) * 2 * m)); CU(cudaMalloc(&d_denom, sizeof(double) * m));
        CU(cudaMalloc(&d_rtab9, sizeof(double) * 9 * m)); CU(cudaMalloc(&d_active, (size_t)m));
        auto cont = [&](int j) {                                                  // bundle_euclid.m:120-123
            bool go = it[(size_t)j] < o.max_iter && it2[(size_t)j] < o.max_iter2;
            if (go && it[(size_t)j] >= 3) {
                const double* e = err.data() + (size_t)j * (maxit + 1);
                go = e[it[(size_t)j] - 1] > o.abs_tol && (e[it[(size_t)j] - 2] - e[it[(size_t)j] - 1]) > o.rel_tol * e[it[(size_t)j] - 2];
            }
            return go;
        };
        for (;;) {
            int nact = 0;
            for (int j = 0; j < m; j++) { active[(size_t)j] = cont(j) ? 1 : 0; nact += active[(size_t)j]; }
            if (nact == 0) break;
            CHK(upload(ctx, d_lam, lam.data(), (size_t)m));
            CHK(upload(ctx, d_active, active.data(), (size_t)m));
            CHK(DISPATCH_NA(ctx, run_independent_round)(ctx, d_lam, d_active, d_cost, d_denom, d_rtab9));
            CHK(download(ctx, h_cost.data(), d_cost, (size_t)2 * m));
            CHK(download(ctx, h_denom.data(), d_denom, (size_t)m));
            CU(cudaStreamSynchronize(ctx->stream));
            nround++;
            bool any_accept = false;
            for (int j = 0; j < m; j++) {
                if (!active[(size_t)j]) continue;
                const double oldc = h_cost[(size_t)j], newc = h_cost[(size_t)m + j];
                const double rho = (oldc - newc) / h_denom[(size_t)j];           // :217
                double* e = err.data() + (size_t)j * (maxit + 1);
                if (oldc - newc > 0) {                                            // :218-232
                    for (int k = 0; k < na; k++) ctx->h_a[(size_t)na * j + k] = ctx->h_a_new[(size_t)na * j + k];
                    const double f = 1 - (2 * rho - 1) * (2 * rho - 1) * (2 * rho - 1);
                    lam[(size_t)j] *= std::max(1.0 / 3.0, f);
                    nu[(size_t)j] = 2.0;
                    e[it[(size_t)j] - 1] = oldc / nvis[(size_t)j];
                    it[(size_t)j] += 1;
                    e[it[(size_t)j] - 1] = newc / nvis[(size_t)j];
                    it2[(size_t)j] = 0;
                    any_accept = true;
                } else {                                                          // :233-241
                    lam[(size_t)j] *= nu[(size_t)j];
                    nu[(size_t)j] *= 2;
                    it2[(size_t)j] += 1;
                }
            }
            if (any_accept) {
                CHK(upload(ctx, ctx->a, ctx->h_a.data(), (size_t)N));
                ctx->rtab_valid = false;
            }
            ctx->s1_valid = ctx->s2_valid = ctx->s3_valid = false;
        }
        return VLG_BA_OK;
    };
    const int r = body();
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_lam); cudaFree(d_cost); cudaFree(d_denom); cudaFree(d_rtab9); cudaFree(d_active);
    if (r != VLG_BA_OK) return r;
    if (a_out) memcpy(a_out, ctx->h_a.data(), sizeof(double) * (size_t)N);
    for (int j = 0; j < m; j++) {
        const bool any = it[(size_t)j] > 1;                                       // error_ stays empty when no step was accepted (:119)
        const int ne = any ? it[(size_t)j] : 0;
        if (n_error) n_error[j] = ne;
        if (error_) for (int k = 0; k < maxit; k++) error_[(size_t)j * maxit + k] = k < ne ? err[(size_t)j * (maxit + 1) + k] : 0.0;
    }
    if (rounds) *rounds = nround;
    return VLG_BA_OK;
}
