// vlg_ba_host.inl -- included by vlg_ba.cu: stage 2 / stage 3 launch sequences, the LM
// driver (bundle_euclid.m:111-267) and the extern "C" entry points of include/vlg_ba.h.

namespace {

int* pcg_done_ptr(vlg_ba_ctx* ctx) { return (int*)((char*)ctx->pcg_sc + offsetof(PcgScalars, done)); }

// all-reduce of the per-iteration PCG vector: peer-memory mailboxes when imported, NCCL otherwise
int allreduce_pcg_vector(vlg_ba_ctx* ctx, double* v, int n, const int* done)
{
    if (ctx->nranks <= 1) return VLG_BA_OK;
    if (!ctx->p2p_ready) return allreduce(ctx, v, (size_t)n);
    ctx->p2p_epoch++;
    k_p2p_allreduce<<<cdiv(n, kP2pThreads), kP2pThreads, 0, ctx->stream>>>(ctx->p2p, n, ctx->p2p_epoch, done, v);
    ctx->launches++;
    CU(cudaGetLastError());
    return VLG_BA_OK;
}

// S (or this rank's share of it) from the Y kept by k_cam_schur_diag: heavy blocks a warp each, light blocks a thread each
template <int NA>
int assemble_S(vlg_ba_ctx* ctx, int add_U, double* S, int ccams = 0)
{
    k_schur_diag_fill<NA><<<cdiv((int64_t)ctx->m * NA * NA, 256), 256, 0, ctx->stream>>>(ctx->m, ctx->Np, ccams, add_U, ctx->red2_local, ctx->Ud, S);
    ctx->launches++;
    if (ctx->nsblk > 0) {
        // the heavy blocks: their (block, chunk) partials were left by k_schur_chunk in this LM step's camera pass
        k_schur_fold<NA><<<cdiv(ctx->nsblk, 4), 128, 0, ctx->stream>>>(ctx->nsblk, ctx->sblk, ctx->sblk_hptr, ctx->Np, ccams, ctx->blk_j, ctx->blk_k,
                                                                      ctx->Hpart, S);
        ctx->launches++;
    }
    if (ctx->nheavy > 0) {
        k_schur_blocks_heavy<NA><<<cdiv(ctx->nheavy, kWarpsPerBlock), kWarpsPerBlock * 32, 0, ctx->stream>>>(
            ctx->nheavy, ctx->blk_heavy, ctx->Np, ccams, add_U, ctx->blk_j, ctx->blk_k, ctx->blk_ptr, ctx->pairs, ctx->Ybuf, ctx->W, ctx->Ud, S);
        ctx->launches++;
    }
    if (ctx->nlight > 0) {
        k_schur_blocks_light<NA><<<cdiv(ctx->nlight, 128), 128, 0, ctx->stream>>>(
            ctx->nlight, ctx->blk_light, ctx->Np, ccams, add_U,
            (ctx->use_explicit && ccams == 0) ? 2 * (128 / NA) + 8 : 1 << 30,      // PCG on S: upper blocks only near the diagonal
            ctx->blk_j, ctx->blk_k, ctx->blk_ptr, ctx->pairs, ctx->Ybuf, ctx->W, ctx->Ud, S);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return VLG_BA_OK;
}

// reprojection-error map of the current state (vlg_ba_reproj_errors)
template <int NA>
int run_reproj_errors(vlg_ba_ctx* ctx, double depth_max, double* d_err, double* d_depth, double* d_bad)
{
    if (!ctx->rtab_valid) CHK(run_rtab<NA>(ctx, ctx->h_a, ctx->a, 4, ctx->h_rtab, ctx->rtab));
    ctx->rtab_valid = true;
    if (ctx->nobs > 0) {
        k_reproj_errors<NA><<<cdiv(ctx->nobs, 256), 256, 0, ctx->stream>>>(ctx->nobs, ctx->obs_xy, ctx->obs_pt, ctx->obs_cam, ctx->K4, ctx->a,
                                                                          ctx->b, ctx->rtab, depth_max, d_err, d_depth, d_bad);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return VLG_BA_OK;
}

// damping + V*^-1 + camera-keyed Schur sums + (Cholesky: S, factor, solve | PCG)  ->  da
template <int NA>
int run_stage2(vlg_ba_ctx* ctx, double lambda)
{
    constexpr int NU = nu_of(NA);
    const int m = ctx->m, n = ctx->n, N = NA * m;
    {
        TimedScope ts(ctx, T_VINV);
        k_damp_U<NA><<<cdiv((int64_t)m * NA * NA, 256), 256, 0, ctx->stream>>>(m, lambda, ctx->U, ctx->Ud);
        ctx->launches++;
        if (n > 0) {
            k_vinv_damp<<<cdiv(n, 128), 128, 0, ctx->stream>>>(n, lambda, ctx->V, ctx->Vinv, ctx->eB, ctx->VE);
            ctx->launches++;
        }
    }
    {
        TimedScope ts(ctx, T_SCHUR);
        const bool chunk_kernel = ctx->nstiles > 0 && ctx->schur_chunk_ok;
        if (chunk_kernel) {
            if constexpr (NA == 6) {
                SchurChunkArgs sa;
                sa.chunk_meta = ctx->stile_meta; sa.obs_pt = ctx->obs_pt; sa.W = ctx->W; sa.VE = ctx->VE; sa.part = ctx->Spart; sa.Yout = ctx->Ybuf;
                sa.chunk_round_ptr = ctx->nsegs > 0 ? ctx->chunk_seg_ptr : nullptr; sa.rounds = ctx->segs; sa.pairs = ctx->pairs; sa.Hpart = ctx->Hpart;
                CU(cudaFuncSetAttribute(k_schur_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSchurChunkSmem));
                k_schur_chunk<<<ctx->nstiles, kSchurTile, kSchurChunkSmem, ctx->stream>>>(sa);
                ctx->launches++;
            }
        } else if (ctx->nchunks > 0) {
            k_cam_schur_diag<NA><<<cdiv(ctx->nchunks, kWarpsPerBlock), kWarpsPerBlock * 32, 0, ctx->stream>>>(
                ctx->nchunks, ctx->chunk_begin, ctx->chunk_end, ctx->obs_pt, ctx->W, ctx->Vinv, ctx->eB, ctx->Spart, ctx->Ybuf);
            ctx->launches++;
        }
        k_cam_sum_partials<<<cdiv((int64_t)m * NU, 128), 128, 0, ctx->stream>>>(m, NU, chunk_kernel ? ctx->cam_stile_ptr : ctx->cam_chunk_ptr, ctx->Spart, nullptr,
                                                                                 ctx->red2);
        ctx->launches++;
    }
    if (ctx->red2_local)      // this rank's own sums feed the diagonal blocks of its share of S
        CU(cudaMemcpyAsync(ctx->red2_local, ctx->red2, sizeof(double) * (size_t)NU * m, cudaMemcpyDeviceToDevice, ctx->stream));
    CHK(allreduce(ctx, ctx->red2, (size_t)NU * m));
    k_cam_schur_finalize<NA><<<cdiv(m, 64), 64, 0, ctx->stream>>>(m, ctx->red2, ctx->Ud, ctx->eA, ctx->Sjj, ctx->ebar,
                                                                  ctx->use_chol ? nullptr : ctx->Minv);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));

    if (ctx->use_chol) {
        const int Np = ctx->Np, nb = Np / kNB;
        {
            TimedScope ts(ctx, T_SCHUR_BLK);
            CU(cudaMemsetAsync(ctx->S, 0, sizeof(double) * (size_t)Np * Np, ctx->stream));
            CHK(assemble_S<NA>(ctx, ctx->rank == 0 ? 1 : 0, ctx->S));
        }
        CHK(allreduce(ctx, ctx->S, (size_t)Np * Np));
        {
            TimedScope ts(ctx, T_CHOL);
            CholArgs ca;
            ca.S = ctx->S; ca.ld = Np; ca.nb = nb; ca.N = N; ca.rhs = ctx->ebar; ca.R = ctx->chol_R; ca.Ld = ctx->chol_Ld; ca.x = ctx->da;
            ca.Dinv = ctx->chol_Dinv; ca.barrier = ctx->chol_bar; ca.prof = nullptr;
            CU(cudaMemsetAsync(ctx->chol_bar, 0, sizeof(unsigned int), ctx->stream));
            void* args[] = {&ca};
            CU(cudaLaunchCooperativeKernel((void*)k_chol_coop, dim3(ctx->chol_grid), dim3(kCholWarps * 32), args, 0, ctx->stream));
            ctx->launches++;
        }
        CU(cudaGetLastError());
        ctx->last_solver = VLG_BA_SOLVER_CHOL;
        ctx->last_pcg_iters = 0; ctx->last_pcg_relres = 0.0;
    } else {
        const double rtol = ctx->opt.pcg_rtol;
        int* done = pcg_done_ptr(ctx);
        constexpr int NW = 3 * NA;
        const size_t sm_pt = sizeof(double) * (kPtTile * NW + kPtTile * 3) + 16;
        const size_t sm_cam = sizeof(double) * (kCamTile * NW + kCamWarps * NA) + 16;
        const bool tiled_cam = ctx->chunk_size <= kCamTile;
        const bool need_wq = ctx->nranks > 1 || ctx->coop_grid == 0 || ctx->use_explicit;
        constexpr size_t kSymvSmem = kSymvSmemBytes;
        if (ctx->use_explicit) {
            CU(cudaFuncSetAttribute(k_symv_lower, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSymvSmem));
            // this rank's share of sum_i Y_ij W_ik' (S without U*; U* p is added by the update kernel)
            TimedScope ts(ctx, T_SCHUR_BLK);
            // the block pattern is static and nothing else writes S on this path: the structurally zero blocks only
            // have to be cleared once per problem (0.91 GB of memset per LM step otherwise)
            if (!ctx->S_zeroed) {
                CU(cudaMemsetAsync(ctx->S, 0, sizeof(double) * (size_t)ctx->Np * ctx->Np, ctx->stream));
                ctx->S_zeroed = true;
            } else if (ctx->s_split) {
                // this rank's column block holds last step's SUM over ranks, whose pattern is wider than the local one
                const size_t off = (size_t)ctx->Np * kSymvCols * ctx->s_J0, cnt = (size_t)ctx->Np * kSymvCols * (ctx->s_J1 - ctx->s_J0);
                if (cnt) CU(cudaMemsetAsync(ctx->S + off, 0, sizeof(double) * cnt, ctx->stream));
            }
            CHK(assemble_S<NA>(ctx, 0, ctx->S));
        }
        const double* McL = nullptr;
        // the update kernel keeps its cluster inverse in shared memory when every cluster gets an SM of its own
        const size_t upd_smem = sizeof(double) * (size_t)Cluster<NA>::NC * 128;
        int mcl_in_smem = ctx->McL && ctx->coop_grid <= ctx->nsm ? 1 : 0;
        if (mcl_in_smem) {
            CU(cudaFuncSetAttribute(k_pcg_update_coop<NA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
            CU(cudaFuncSetAttribute(k_pcg_update_coop<NA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
        }
        bool upd_batch = false;
        if (ctx->coop_grid > 0) {
            int per_sm = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_update_coop<NA, true>, 128, mcl_in_smem ? upd_smem : 0));
            upd_batch = ctx->coop_grid <= per_sm * ctx->nsm;
        }
        // assembled S on one GPU (or with the peer-memory exchange): the whole solve is one persistent kernel
        const bool persistent = ctx->use_explicit && ctx->persist_ok && ctx->coop_grid > 0 && ctx->coop_grid <= ctx->symv_grid &&
                                (ctx->nranks == 1 || ctx->p2p_ready) && ctx->opt.pcg_max_iter > 0;
        // ... and only that kernel applies the second (shifted) cluster partition of the overlapping preconditioner
        const bool overlap = persistent && ctx->McL && ctx->Mc2 && ctx->overlap_ok;
        const double* Mc2 = nullptr;
        if (ctx->McL) {
            // cluster-Jacobi: diagonal blocks of S over the update kernel's CTAs, summed over ranks, inverted
            TimedScope ts(ctx, T_PRECOND);
            const size_t nblk = (size_t)ctx->coop_grid * (overlap ? 2 : 1) + (overlap ? 1 : 0);
            if (ctx->use_explicit) {
                k_cluster_gather<NA><<<ctx->coop_grid, 128, 0, ctx->stream>>>(m, ctx->Np, ctx->S, ctx->Cblk, 0);
                ctx->launches++;
                if (overlap) {
                    k_cluster_gather<NA><<<ctx->coop_grid + 1, 128, 0, ctx->stream>>>(m, ctx->Np, ctx->S, ctx->Cblk + (size_t)ctx->coop_grid * 128 * 128,
                                                                                      Cluster<NA>::kShift);
                    ctx->launches++;
                }
            } else {
                // implicit path: only the within-cluster blocks are assembled (pair lists restricted to them)
                CU(cudaMemsetAsync(ctx->Cblk, 0, sizeof(double) * (size_t)ctx->coop_grid * 128 * 128, ctx->stream));
                CHK(assemble_S<NA>(ctx, 0, ctx->Cblk, Cluster<NA>::kCams));
            }
            CHK(allreduce(ctx, ctx->Cblk, nblk * 128 * 128));
            CU(cudaFuncSetAttribute(k_cluster_inverse<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cluster<NA>::kSmem));
            k_cluster_inverse<NA><<<ctx->coop_grid, 256, Cluster<NA>::kSmem, ctx->stream>>>(m, 1, 0, ctx->coop_grid, ctx->Cblk, ctx->Ud, ctx->McL);
            ctx->launches++;
            if (overlap) {
                k_cluster_inverse<NA><<<ctx->coop_grid + 1, 256, Cluster<NA>::kSmem, ctx->stream>>>(
                    m, 1, Cluster<NA>::kShift, ctx->coop_grid, ctx->Cblk + (size_t)ctx->coop_grid * 128 * 128, ctx->Ud, ctx->Mc2);
                ctx->launches++;
                Mc2 = ctx->Mc2;
            }
            CU(cudaGetLastError());
            McL = ctx->McL;
        }
        CU(cudaFuncSetAttribute(k_sweep_pt_tiled<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pt));
        CU(cudaFuncSetAttribute(k_sweep_cam_tiled<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_cam));
        // persistent ring variants: 2 stages, as many CTAs per SM as shared memory allows
        constexpr int RS = 2;
        const size_t sm_ring_cam = sizeof(double) * ((size_t)RS * kRingTile * NW + 2 * (kRingTile / 32) * NA) + sizeof(int2) * kRingMaxTiles + 8 * RS + 16;
        const size_t sm_ring_pt = sizeof(double) * ((size_t)RS * kRingTile * NW + 2 * kRingTile * 3) + sizeof(int4) * kRingMaxTiles + 8 * RS + 16;
        const int g_ring_cam = ctx->nsm * (int)std::max<size_t>(1, (size_t)(227 * 1024) / (sm_ring_cam + 1024));
        const int g_ring_pt = ctx->nsm * (int)std::max<size_t>(1, (size_t)(227 * 1024) / (sm_ring_pt + 1024));
        const bool ring_cam = (ctx->use_ring & 1) && tiled_cam && ctx->nchunks <= (int64_t)g_ring_cam * kRingMaxTiles && ctx->chunk_size <= kRingTile;
        const bool ring_pt = (ctx->use_ring & 2) && ctx->tiled_ok && ctx->pt_tile <= kRingTile && ctx->nptiles <= (int64_t)g_ring_pt * kRingMaxTiles;
        if (ring_cam) CU(cudaFuncSetAttribute(k_sweep_cam_ring<NA, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_ring_cam));
        if (ring_pt) CU(cudaFuncSetAttribute(k_sweep_pt_ring<NA, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_ring_pt));
        // the two sweeps of W V*^-1 W' v: chunk partials land in qpart (and, all-reduced, in wq)
        if (ctx->use_explicit && ctx->s_split) {
            // S <- sum over ranks, column block r on rank r (in place, one grouped ncclReduce per block); the cluster blocks
            // above were taken from the local shares and all-reduced on their own
            TimedScope ts(ctx, T_SCHUR_BLK);
            if (ctx->p2p_ready && ctx->peer_S_ready) {
                // every rank's share must be complete before anyone pulls: a 1-element all-reduce is the barrier (the solve
                // itself keeps the ranks in lockstep afterwards, so nobody rewrites its share while it is being read)
                CHK(allreduce(ctx, ctx->scal3 + 3, 1));
                const int ncols = kSymvCols * (ctx->s_J1 - ctx->s_J0);
                if (ncols > 0) {
                    k_pull_reduce_block<<<ncols, 256, 0, ctx->stream>>>(ctx->peer_S, ctx->Np, ctx->s_J0, ctx->S);
                    ctx->launches++;
                }
                CU(cudaGetLastError());
                // ... and before anyone's next assembly overwrites a share, every pull must be over: same barrier
                CHK(allreduce(ctx, ctx->scal3 + 3, 1));
            } else {
            int rc = g_nccl.GroupStart();
            for (int r = 0; r < ctx->nranks && rc == 0; r++) {
                const size_t off = (size_t)ctx->Np * kSymvCols * ctx->s_bounds[(size_t)r];
                const size_t cnt = (size_t)ctx->Np * kSymvCols * (ctx->s_bounds[(size_t)r + 1] - ctx->s_bounds[(size_t)r]);
                if (cnt) rc = g_nccl.Reduce(ctx->S + off, ctx->S + off, cnt, kNcclFloat64, kNcclSum, r, ctx->comm, ctx->stream);
            }
            if (rc == 0) rc = g_nccl.GroupEnd();
            if (rc != 0) return fail(ctx, VLG_BA_ENCCL, "ncclReduce (column blocks of S): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
            }
        }
        auto sweeps = [&](const double* v, const int* dn) -> int {
            if (ctx->use_explicit) {
                // wq = (sum Y W') v = -(S - U*) v from the lower triangle of the assembled S
                TimedScope ts(ctx, T_SYMV);
                k_symv_lower<<<ctx->symv_grid, kSymvRows + 32, kSymvSmem, ctx->stream>>>(ctx->Np, N, ctx->S, v, ctx->symv_tile_ptr,
                                                                                   ctx->symv_tiles, dn, ctx->symv_rowpart, ctx->symv_colpart);
                const SymvFold fl{ctx->symv_row_ptr, ctx->symv_row_list, ctx->symv_col_ptr, ctx->symv_col_list};
                const bool fused = ctx->nranks > 1 && ctx->p2p_ready;     // the exchange with the peers happens inside the finish kernel
                if (fused) ctx->p2p_epoch++;
                k_symv_finish<<<ctx->Np / 32, 1024, 0, ctx->stream>>>(N, -1.0, fl, ctx->symv_rowpart, ctx->symv_colpart, dn, ctx->wq,
                                                                      fused ? ctx->p2p_dev : nullptr, ctx->p2p_epoch);
                ctx->launches += 2;
                CU(cudaGetLastError());
                if (!fused) CHK(allreduce_pcg_vector(ctx, ctx->wq, N, dn));
                return VLG_BA_OK;
            }
            if (n > 0) {
                TimedScope ts(ctx, T_SWEEP_PT);
                if (ring_pt)
                    k_sweep_pt_ring<NA, RS><<<std::min(g_ring_pt, ctx->nptiles), kRingTile, sm_ring_pt, ctx->stream>>>(
                        ctx->nptiles, ctx->ptile_meta, ctx->pt_ptr, ctx->pt_cam, ctx->Wp, ctx->Vinv, v, dn, ctx->tvec);
                else if (ctx->tiled_ok)
                    k_sweep_pt_tiled<NA><<<ctx->nptiles, kPtTile, sm_pt, ctx->stream>>>(ctx->ptile_meta, ctx->pt_ptr, ctx->pt_cam,
                                                                                      ctx->Wp, ctx->Vinv, v, dn, ctx->tvec);
                else
                    k_sweep_pt<NA><<<cdiv(n, 128), 128, 0, ctx->stream>>>(n, ctx->pt_ptr, ctx->pt_obs, ctx->pt_cam, ctx->W, ctx->Vinv,
                                                                         v, dn, ctx->tvec);
                ctx->launches++;
            }
            if (ctx->nchunks > 0) {
                TimedScope ts(ctx, T_SWEEP_CAM);
                if (ring_cam)
                    k_sweep_cam_ring<NA, RS><<<std::min(g_ring_cam, ctx->nchunks), kRingTile, sm_ring_cam, ctx->stream>>>(
                        ctx->nchunks, ctx->chunk_meta, ctx->obs_pt, ctx->W, ctx->tvec, dn, ctx->qpart);
                else if (tiled_cam)
                    k_sweep_cam_tiled<NA><<<ctx->nchunks, kCamTile, sm_cam, ctx->stream>>>(ctx->chunk_meta, ctx->obs_pt,
                                                                                         ctx->W, ctx->tvec, dn, ctx->qpart);
                else
                    k_sweep_cam<NA><<<cdiv(ctx->nchunks, kWarpsPerBlock), kWarpsPerBlock * 32, 0, ctx->stream>>>(
                        ctx->nchunks, ctx->chunk_begin, ctx->chunk_end, ctx->obs_pt, ctx->W, ctx->tvec, dn, ctx->qpart);
                ctx->launches++;
            }
            if (need_wq) {
                k_cam_sum_partials<<<cdiv((int64_t)m * NA, 128), 128, 0, ctx->stream>>>(m, NA, ctx->cam_chunk_ptr, ctx->qpart, dn, ctx->wq);
                ctx->launches++;
                CHK(allreduce_pcg_vector(ctx, ctx->wq, N, dn));
            }
            CU(cudaGetLastError());
            return VLG_BA_OK;
        };
        const bool defl = ctx->opt.pcg_deflate && ctx->coop_grid > 0 && !ctx->opt.fix_structure && NA != kNaProjective;   // the gauge vectors are the Euclidean model's
        if (defl) {
            k_gauge_vectors<NA><<<cdiv(m, 128), 128, 0, ctx->stream>>>(m, ctx->a, ctx->rtab, ctx->cam_fixed, ctx->cam_chunk_ptr, ctx->Zd);
            ctx->launches++;
            for (int d = 0; d < kDefl; d++) {
                CHK(sweeps(ctx->Zd + (size_t)d * N, nullptr));
                k_apply_S_finalize<NA><<<cdiv((int64_t)N, 128), 128, 0, ctx->stream>>>(m, ctx->cam_chunk_ptr, ctx->qpart,
                                                                                    need_wq ? ctx->wq : nullptr, ctx->Ud,
                                                                                    ctx->Zd + (size_t)d * N, ctx->SZd + (size_t)d * N);
                ctx->launches++;
            }
            if (ctx->init_part && ctx->init_coop_cap < 0) {
                // can the whole update grid be co-resident with this kernel's register footprint?
                int per_sm = 0;
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_init_defl_coop<NA>, 128, 0));
                ctx->init_coop_cap = per_sm * ctx->nsm;
            }
            if (ctx->init_part && ctx->coop_grid <= ctx->init_coop_cap) {
                // cooperative set-up on the update kernel's grid
                int m_ = m;
                const double *eb = ctx->ebar, *mi = ctx->Minv, *zz = ctx->Zd, *sz = ctx->SZd, *mcl = McL, *mc2 = Mc2;
                DeflScalars* dsc = ctx->defl_sc;
                double *xx = ctx->da, *rr = ctx->pr, *pzv = ctx->pz, *ppv = ctx->pp, *ip = ctx->init_part;
                PcgScalars* sc = ctx->pcg_sc;
                unsigned int* ib = ctx->init_bar;
                CU(cudaMemsetAsync(ctx->init_bar, 0, sizeof(unsigned int), ctx->stream));
                void* args[] = {&m_, &eb, &mi, &zz, &sz, &dsc, &xx, &rr, &pzv, &ppv, &sc, &mcl, &ip, &ib, &mc2};
                CU(cudaLaunchCooperativeKernel((void*)k_pcg_init_defl_coop<NA>, dim3(ctx->coop_grid), dim3(128), args, 0, ctx->stream));
            } else {
                k_pcg_init_defl<NA><<<1, 1024, 0, ctx->stream>>>(m, ctx->ebar, ctx->Minv, ctx->Zd, ctx->SZd, ctx->defl_sc, ctx->da, ctx->pr,
                                                               ctx->pz, ctx->pp, ctx->pcg_sc, McL, Mc2);
            }
        } else {
            k_pcg_init<NA><<<1, 1024, 0, ctx->stream>>>(m, ctx->ebar, ctx->Minv, ctx->da, ctx->pr, ctx->pz, ctx->pp, ctx->pcg_sc, rtol, McL, Mc2);
        }
        ctx->launches++;
        int launched = 0;
        bool deferred = false;
        const int batch = 8;
        if (persistent) {
            TimedScope ts(ctx, T_PCG_PERSIST);
            PcgPersistArgs pa;
            pa.Np = ctx->Np; pa.ld = ctx->Np; pa.N = N; pa.m = m; pa.max_iter = ctx->opt.pcg_max_iter; pa.nclusters = ctx->coop_grid;
            pa.rtol = rtol; pa.S = ctx->S; pa.tile_ptr = ctx->symv_tile_ptr; pa.tiles = ctx->symv_tiles;
            pa.rowpart = ctx->symv_rowpart; pa.colpart = ctx->symv_colpart; pa.wq = ctx->wq;
            pa.Ud = ctx->Ud; pa.Minv = ctx->Minv; pa.McL = McL; pa.Mc2 = Mc2; pa.x = ctx->da; pa.r = ctx->pr; pa.p = ctx->pp; pa.p2 = ctx->pp2; pa.zbuf = ctx->pz;
            pa.sc = ctx->pcg_sc; pa.blkpart = ctx->blkpart; pa.Z = defl ? ctx->Zd : nullptr; pa.SZ = defl ? ctx->SZd : nullptr;
            pa.ds = ctx->defl_sc; pa.barrier = ctx->persist_bar;
            pa.mb = ctx->nranks > 1 ? ctx->p2p_dev : nullptr; pa.epoch0 = ctx->p2p_epoch + 1;
            pa.fold = SymvFold{ctx->symv_row_ptr, ctx->symv_row_list, ctx->symv_col_ptr, ctx->symv_col_list};
            pa.prof = nullptr;
            pa.stat = ctx->symv_learn_left > 0 ? ctx->symv_stat : nullptr;
            // staging the cluster inverses with an evict-first hint keeps them (22 MB with both partitions) from crowding the matvec's
            // partials out of L2: measured -2.2 us (matvec phase) + 0.9 us (fold) per iteration at Venice shape
            { const char* e = getenv("VLG_BA_MCL_EVICT"); pa.mcl_evict_first = e ? atoi(e) : (overlap ? 1 : 0); }
            long long*& d_prof = ctx->persist_prof;
            const bool want_prof = getenv("VLG_BA_PERSIST_PROF") != nullptr;
            if (want_prof) {
                if (!d_prof) CU(cudaMalloc(&d_prof, (32 + 1024) * sizeof(long long)));
                cudaMemsetAsync(d_prof, 0, (32 + 1024) * sizeof(long long), ctx->stream);
                pa.prof = d_prof;
            }
            CU(cudaMemsetAsync(ctx->persist_bar, 0, sizeof(unsigned int), ctx->stream));
            CU(cudaFuncSetAttribute(k_pcg_persistent<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSymvSmemBytes));
            void* args[] = {&pa};
            CU(cudaLaunchCooperativeKernel((void*)k_pcg_persistent<NA>, dim3(ctx->symv_grid), dim3(kSymvRows + 32), args, kSymvSmemBytes,
                                           ctx->stream));
            ctx->launches++;
            launched = ctx->opt.pcg_max_iter;
            CU(cudaMemcpyAsync(ctx->h_pcg, ctx->pcg_sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream));
            // the host only has to wait here when it needs the iteration count now (mailbox epochs of the next exchange,
            // autotuning, profile); otherwise the scalars are picked up at the step's final synchronisation
            deferred = !(ctx->nranks > 1 || pa.stat || want_prof);
            if (!deferred) CU(cudaStreamSynchronize(ctx->stream));
            if (ctx->nranks > 1) ctx->p2p_epoch += (unsigned int)ctx->h_pcg->exchanges;   // one mailbox epoch per matvec exchange, on every rank (the breakdown exit has used one more than it completed iterations)
            if (pa.stat && ctx->h_pcg->iters > 0) CHK(symv_learn(ctx));
            if (want_prof) {
                long long hp[18];
                cudaMemcpy(hp, d_prof, sizeof(hp), cudaMemcpyDeviceToHost);
                int khz = 0;
                cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
                const char* nm[18] = {"matvec", "barrier", "fold-rest", "barrier", "phase3", "barrier", "phase4-rest", "barrier", "phase5", "barrier",
                                      "p4:x,r", "p4:wait-McL", "p4:matvec", "p4:partials", "p4:sync", "fold:loads", "fold:tma-issue", "fold:list-ptrs"};
                fprintf(stderr, "k_pcg_persistent, %d iterations, us per iteration (CTA 0):", ctx->h_pcg->iters);
                for (int k = 0; k < 18; k++) fprintf(stderr, " %s %.1f", nm[k], hp[k] / (khz * 1e-3) / std::max(ctx->h_pcg->iters, 1));
                fprintf(stderr, "\n");
                std::vector<long long> hc((size_t)ctx->symv_grid);
                cudaMemcpy(hc.data(), d_prof + 32, sizeof(long long) * hc.size(), cudaMemcpyDeviceToHost);
                fprintf(stderr, "  matvec us per iteration by CTA:");
                for (size_t k = 0; k < hc.size(); k++) fprintf(stderr, " %.1f", hc[k] / (khz * 1e-3) / std::max(ctx->h_pcg->iters, 1));
                fprintf(stderr, "\n");
            }
        }
        while (launched < ctx->opt.pcg_max_iter) {
            for (int it = 0; it < batch && launched < ctx->opt.pcg_max_iter; it++, launched++) {
                CHK(sweeps(ctx->pp, done));
                {
                    TimedScope ts(ctx, T_PCG_UPDATE);
                    if (ctx->coop_grid > 0) {
                        int m_ = m;
                        const int* ccp = ctx->cam_chunk_ptr;
                        const double* qp = ctx->qpart;
                        const double* wqp = need_wq ? ctx->wq : nullptr;
                        const double* ud = ctx->Ud;
                        const double* mi = ctx->Minv;
                        double *xx = ctx->da, *rr = ctx->pr, *ppv = ctx->pp, *bp = ctx->blkpart;
                        PcgScalars* sc = ctx->pcg_sc;
                        double rt = rtol;
                        const double* zz = defl ? ctx->Zd : nullptr;
                        const double* sz = defl ? ctx->SZd : nullptr;
                        const DeflScalars* dsc = ctx->defl_sc;
                        const double* mcl = McL;
                        void* args[] = {&m_, &ccp, &qp, &wqp, &ud, &mi, &xx, &rr, &ppv, &sc, &bp, &rt, &zz, &sz, &dsc, &mcl, &mcl_in_smem};
                        CU(cudaLaunchCooperativeKernel(upd_batch ? (void*)k_pcg_update_coop<NA, true> : (void*)k_pcg_update_coop<NA, false>,
                                                       dim3(ctx->coop_grid), dim3(128), args, mcl_in_smem ? upd_smem : 0, ctx->stream));
                    } else {
                        k_pcg_update<NA><<<1, 1024, 0, ctx->stream>>>(m, ctx->Ud, ctx->Minv, ctx->wq, ctx->da, ctx->pr, ctx->pz, ctx->pp,
                                                                      ctx->pq, ctx->pcg_sc, rtol);
                    }
                    ctx->launches++;
                }
            }
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(ctx->h_pcg, ctx->pcg_sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            if (ctx->h_pcg->done) break;
        }
        if (!deferred && !ctx->h_pcg->done) {
            CU(cudaMemcpyAsync(ctx->h_pcg, ctx->pcg_sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
        if (defl) {
            k_pcg_defl_final<NA><<<1, 1024, 0, ctx->stream>>>(m, ctx->Zd, ctx->SZd, ctx->defl_sc, ctx->da);
            ctx->launches++;
            CU(cudaGetLastError());
        }
        ctx->last_solver = ctx->use_explicit ? VLG_BA_SOLVER_PCG_EXPLICIT : VLG_BA_SOLVER_PCG;
        ctx->pcg_pending = deferred;
        if (!deferred) {
            ctx->last_pcg_iters = ctx->h_pcg->iters;
            ctx->last_pcg_relres = ctx->h_pcg->r0n2 > 0.0 ? sqrt(ctx->h_pcg->rn2 / ctx->h_pcg->r0n2) : 0.0;
        }
    }
    return VLG_BA_OK;
}

// iteration count / residual of the last solve, once the host has synchronised with the stream
static int resolve_pcg_info(vlg_ba_ctx* ctx)
{
    if (!ctx->pcg_pending) return VLG_BA_OK;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->last_pcg_iters = ctx->h_pcg->iters;
    ctx->last_pcg_relres = ctx->h_pcg->r0n2 > 0.0 ? sqrt(ctx->h_pcg->rn2 / ctx->h_pcg->r0n2) : 0.0;
    ctx->pcg_pending = false;
    return VLG_BA_OK;
}

// db, a_new, b_new, new cost and dp'(lambda dp + g)
template <int NA>
int run_stage3(vlg_ba_ctx* ctx, double lambda, double* new_cost, double* denom)
{
    const int m = ctx->m, n = ctx->n, N = NA * m;
    const bool host_tab = ctx->opt.rtable == VLG_BA_RTABLE_HOST_LIBM;
    ctx->rtab_next_valid = false;
    if (host_tab) CU(cudaEventRecord(ctx->ev_da, ctx->stream));       // da is final here
    else {
        k_axpy1<<<cdiv(N, 256), 256, 0, ctx->stream>>>(N, ctx->a, ctx->da, ctx->a_new);
        ctx->launches++;
        CHK(run_rtab<NA>(ctx, ctx->h_a_new, ctx->a_new, 1, ctx->h_rtab_new, ctx->rtab_new));
    }
    {
        TimedScope ts(ctx, T_STAGE3);
        if (n > 0) {
            bool coop_done = false;
            if constexpr ((3 * NA) % 2 == 0) {
                static const bool coop_on = []() { const char* e = getenv("VLG_BA_BACKSUB_COOP"); return e ? atoi(e) != 0 : true; }();
                if (ctx->ns1tiles > 0 && coop_on) {
                    const size_t sm = sizeof(double) * (size_t)kS1Tile * 3 * NA;
                    CU(cudaFuncSetAttribute(k_backsub_coop<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                    k_backsub_coop<NA><<<ctx->ns1tiles, kS1Tile, sm, ctx->stream>>>(ctx->s1tile_meta, ctx->pt_ptr, ctx->pt_obs, ctx->pt_cam, ctx->W,
                                                                                   ctx->Vinv, ctx->eB, ctx->da, ctx->b, lambda,
                                                                                   ctx->opt.backsub_all_rows, ctx->db, ctx->b_new, ctx->denom_pt);
                    coop_done = true;
                }
            }
            if (coop_done) {
            } else if (ctx->ns1tiles > 0)
                k_backsub_tiled<NA><<<ctx->ns1tiles, kS1Tile, 0, ctx->stream>>>(ctx->s1tile_meta, ctx->pt_ptr, ctx->pt_obs, ctx->pt_cam, ctx->W,
                                                                             ctx->Vinv, ctx->eB, ctx->da, ctx->b, lambda,
                                                                             ctx->opt.backsub_all_rows, ctx->db, ctx->b_new, ctx->denom_pt);
            else
                k_backsub<NA><<<cdiv(n, 128), 128, 0, ctx->stream>>>(n, ctx->pt_ptr, ctx->pt_obs, ctx->pt_cam, ctx->W, ctx->Vinv, ctx->eB,
                                                                    ctx->da, ctx->b, lambda, ctx->opt.backsub_all_rows, ctx->db,
                                                                    ctx->b_new, ctx->denom_pt);
            ctx->launches++;
            if (ctx->d2h_bnew) {
                // the caller's (pinned) buffer gets b_new over the copy stream while the new cost is still being computed
                CU(cudaEventRecord(ctx->ev_bnew, ctx->stream));
                CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_bnew, 0));
                CU(cudaMemcpyAsync(ctx->d2h_bnew, ctx->b_new, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, ctx->copy_stream));
            }
        }
        if (host_tab) {
            // a_new = a + da on the host (same IEEE add as mex_bundle_3_db_new.c:137-140) so that the rotation matrices of the
            // candidate come from the host libm.  da is fetched on a second stream and the table is computed WHILE the
            // back-substitution runs (it only needs da): all four matrices per camera, so that the next stage 1 finds its
            // table ready when the step is accepted
            CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_da, 0));
            CU(cudaMemcpyAsync(ctx->h_da.data(), ctx->da, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, ctx->copy_stream));
            CU(cudaStreamSynchronize(ctx->copy_stream));
            for (int t = 0; t < N; t++) ctx->h_a_new[t] = ctx->h_a[t] + ctx->h_da[t];
            CHK(upload(ctx, ctx->a_new, ctx->h_a_new.data(), (size_t)N));
            if (NA != kNaProjective) {
                // only the candidate's BASE matrices stand between da and the new cost; the three perturbed ones per camera
                // (three quarters of the libm work: 0.2 ms at Venice shape, serial on purpose) are computed further down,
                // while the GPU is busy with the new cost and its reductions
                rtab_host(m, NA, ctx->h_a_new.data(), 1, ctx->h_rtab_new.data());
                CHK(upload(ctx, ctx->rtab_new, ctx->h_rtab_new.data(), (size_t)9 * m));
            }
        }
        if (ctx->nobs > 0) {
            k_new_cost<NA><<<cdiv(ctx->nobs, 256), 256, 0, ctx->stream>>>(ctx->nobs, ctx->obs_xy, ctx->obs_pt, ctx->obs_cam, ctx->K4,
                                                                         ctx->a_new, ctx->b_new, ctx->rtab_new, ctx->cost_obs, ctx->xhat_out);
            ctx->launches++;
        }
    }
    CU(cudaGetLastError());
    CHK(reduce_to(ctx, ctx->cost_obs, (size_t)ctx->nobs, ctx->scal3));
    CHK(reduce_to(ctx, ctx->denom_pt, (size_t)n, ctx->scal3 + 1));
    CHK(allreduce(ctx, ctx->scal3, 2));
    k_denom_cam<<<1, 1024, 0, ctx->stream>>>(N, ctx->da, ctx->eA, lambda, ctx->scal3 + 2);
    ctx->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ctx->h_pin, ctx->scal3, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (host_tab && NA != kNaProjective) {
        // the candidate's full table (base + the three perturbed matrices per camera), so that the next stage 1 finds it
        // ready when the step is accepted: computed now, under the kernels queued above
        for (int j = 0; j < m; j++)
            for (int q = 0; q < 9; q++) ctx->h_rtab_next[(size_t)36 * j + q] = ctx->h_rtab_new[(size_t)9 * j + q];
        rtab_host(m, NA, ctx->h_a_new.data(), 4, ctx->h_rtab_next.data(), 1);
        CHK(upload(ctx, ctx->rtab_next, ctx->h_rtab_next.data(), (size_t)36 * m));
        ctx->rtab_next_valid = true;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    *new_cost = ctx->h_pin[0];
    // dp = [da; db], g = [eA; eB] (bundle_euclid.m:215-217)
    *denom = ctx->h_pin[2] + ctx->h_pin[1];
    if (ctx->opt.rtable != VLG_BA_RTABLE_HOST_LIBM) {
        CHK(download(ctx, ctx->h_a_new.data(), ctx->a_new, (size_t)N));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return VLG_BA_OK;
}

static int resolve_cost(vlg_ba_ctx* ctx)
{
    if (!ctx->cost_pending) return VLG_BA_OK;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->old_cost = ctx->h_pin[4];
    if (!ctx->num_vis_user) ctx->num_vis = ctx->h_pin[5];
    ctx->cost_pending = false;
    return VLG_BA_OK;
}

#define DISPATCH_NA(ctx, call)                                                   \
    ((ctx)->na == 6 ? call<6> : (ctx)->na == 7 ? call<7> : (ctx)->na == 10 ? call<10> : call<12>)

// `defer`: do not wait for the cost on the host (the LM step only needs it for the accept test after stage 3)
int do_stage1(vlg_ba_ctx* ctx, bool defer = false)
{
    if (!ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    // nvis travels with the cost in the same all-reduce (num_vis, bundle_euclid.m:82)
    const double nv = (double)ctx->nobs;
    CHK(upload(ctx, ctx->scal1 + 1, &nv, 1));
    CHK(DISPATCH_NA(ctx, run_stage1)(ctx, nullptr, nullptr, nullptr, nullptr));
    CHK(allreduce(ctx, ctx->red1, (size_t)ctx->na * ctx->na * ctx->m + (size_t)ctx->na * ctx->m + 2));
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaMemcpyAsync(ctx->h_pin + 4, ctx->scal1, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->cost_pending = true;
    if (!defer) CHK(resolve_cost(ctx));
    ctx->s1_valid = true; ctx->s2_valid = false; ctx->s3_valid = false;
    return VLG_BA_OK;
}

int do_stage2(vlg_ba_ctx* ctx, double lambda)
{
    if (!ctx->s1_valid) return fail(ctx, VLG_BA_ESTATE, "stage2 needs stage1 at the current state");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->s1_valid) {}
    CHK(DISPATCH_NA(ctx, run_stage2)(ctx, lambda));
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    ctx->s2_valid = true; ctx->s2_lambda = lambda; ctx->s3_valid = false;
    return VLG_BA_OK;
}

int do_stage3(vlg_ba_ctx* ctx, double lambda, double* new_cost, double* denom)
{
    if (!ctx->s2_valid) return fail(ctx, VLG_BA_ESTATE, "stage3 needs stage2");
    CU(cudaSetDevice(ctx->device));
    CHK(DISPATCH_NA(ctx, run_stage3)(ctx, lambda, new_cost, denom));
    CU(cudaEventRecord(ctx->ev[4], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->s3_valid = true;
    return VLG_BA_OK;
}

// one trip of the while loop, bundle_euclid.m:139-241
int do_trial(vlg_ba_ctx* ctx, vlg_ba_trial_info* info)
{
    const double lambda = ctx->lambda;
    bool fresh1 = false;
    if (!ctx->s1_valid) { CHK(do_stage1(ctx, true)); fresh1 = true; }   // a rejected step left (a,b) unchanged: mex1 would give the same bits
    else CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CHK(do_stage2(ctx, lambda));
    double new_cost = 0.0, denom = 0.0;
    CHK(do_stage3(ctx, lambda, &new_cost, &denom));
    CHK(resolve_cost(ctx));
    const double old_cost = ctx->old_cost;
    const double rho = (old_cost - new_cost) / denom;                       // :217
    const bool accept = (old_cost - new_cost) > 0;                          // :218
    if (info) {
        memset(info, 0, sizeof(*info));
        info->old_cost = old_cost; info->new_cost = new_cost; info->denom = denom; info->rho = rho;
        info->lambda_used = lambda; info->accepted = accept ? 1 : 0;
        CHK(resolve_pcg_info(ctx));
        info->solver_used = ctx->last_solver; info->pcg_iters = ctx->last_pcg_iters; info->pcg_relres = ctx->last_pcg_relres;
        float ms = 0.f;
        if (fresh1 && cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) info->ms_stage1 = ms;
        if (cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]) == cudaSuccess) info->ms_schur = ms;
        if (cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]) == cudaSuccess) info->ms_solve = ms;
        if (cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]) == cudaSuccess) info->ms_stage3 = ms;
    }
    if (!isfinite(new_cost) && !isfinite(old_cost)) return fail(ctx, VLG_BA_ENUM, "non-finite cost");
    if (accept) {
        std::swap(ctx->a, ctx->a_new);
        std::swap(ctx->b, ctx->b_new);
        ctx->h_a.swap(ctx->h_a_new);
        std::swap(ctx->rtab, ctx->rtab_next);                               // the candidate's table, if stage 3 left one
        ctx->h_rtab.swap(ctx->h_rtab_next);
        ctx->rtab_valid = ctx->rtab_next_valid;
        ctx->rtab_next_valid = false;
        if (ctx->opt.model == VLG_BA_MODEL_PROJECTIVE) {
            ctx->lambda = lambda / 10;                                      // bundle_projective.m:194
        } else {
            const double f = 1 - (2 * rho - 1) * (2 * rho - 1) * (2 * rho - 1);
            ctx->lambda = lambda * std::max(1.0 / 3.0, f);                  // bundle_euclid.m:227
            ctx->nu = 2.0;
        }
        if ((int)ctx->err_hist.size() < ctx->iter + 1) ctx->err_hist.resize((size_t)ctx->iter + 1, 0.0);
        ctx->err_hist[ctx->iter - 1] = old_cost / ctx->num_vis;             // :219-220,229-231
        ctx->err_hist[ctx->iter] = new_cost / ctx->num_vis;
        ctx->iter += 1;
        ctx->iter2 = 0;
        ctx->s1_valid = false;
    } else {
        if (ctx->opt.model == VLG_BA_MODEL_PROJECTIVE) {
            ctx->lambda = lambda * 10;                                      // bundle_projective.m:204
        } else {
            ctx->lambda = lambda * ctx->nu;                                 // bundle_euclid.m:238-240
            ctx->nu = 2 * ctx->nu;
        }
        ctx->iter2 += 1;
    }
    ctx->s2_valid = false; ctx->s3_valid = false;
    if (info) { info->lambda_next = ctx->lambda; info->nu_next = ctx->nu; }
    return VLG_BA_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

void vlg_ba_opts_default(vlg_ba_opts* o)
{
    memset(o, 0, sizeof(*o));
    o->num_variableK = 4;
    o->lambda0 = 0.001; o->nu0 = 2.0; o->max_iter = 20; o->max_iter2 = 10; o->rel_tol = 1e-3; o->abs_tol = 1e-20;
    o->backsub_all_rows = 0;
    o->solver = VLG_BA_SOLVER_AUTO; o->chol_max_cams = 600; o->pcg_rtol = 1e-8; o->pcg_max_iter = 1000;
    o->rtable = VLG_BA_RTABLE_HOST_LIBM; o->order = VLG_BA_ORDER_CHUNKED; o->device = -1; o->verbose = 0;
    o->pcg_deflate = 1;
    o->pcg_cluster = 1;
    o->model = VLG_BA_MODEL_EUCLID;
    o->pcg_autotune = 0;
}

const char* vlg_ba_version(void) { return "vlgba 0.1 (sm_100a, fp64)"; }

const char* vlg_ba_last_error(const vlg_ba_ctx* ctx) { return ctx ? ctx->err : g_create_error; }   // NULL: this thread's last context-less failure

int vlg_ba_create(const vlg_ba_opts* opts, vlg_ba_ctx** out)
{
    vlg_ba_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, VLG_BA_EINVAL, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, VLG_BA_ECUDA, "no CUDA device (%s): this library has no CPU fallback", cudaGetErrorString(e));
    vlg_ba_ctx* c = new vlg_ba_ctx();
    if (opts) c->opt = *opts; else vlg_ba_opts_default(&c->opt);
    int dev = c->opt.device;
    if (dev < 0) cudaGetDevice(&dev);
    if (dev >= ndev) { delete c; return fail(nullptr, VLG_BA_EINVAL, "device %d of %d", dev, ndev); }
    c->device = dev;
    ctx = c;
    if (cudaSetDevice(dev) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_pin, 16 * sizeof(double)) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_pcg, sizeof(PcgScalars)) != cudaSuccess) {
        int r = fail(nullptr, VLG_BA_ECUDA, "context set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return r;
    }
    for (int k = 0; k < 5; k++) cudaEventCreate(&c->ev[k]);
    for (int k = 0; k < 2; k++) cudaEventCreate(&c->ev_sw[k]);
    cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c->ev_da, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_bnew, cudaEventDisableTiming);
    memset(c->h_pcg, 0, sizeof(PcgScalars));
    (void)ctx;
    *out = c;
    return VLG_BA_OK;
}

void vlg_ba_destroy(vlg_ba_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    p2p_close(ctx);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    free_problem(ctx);
    for (int t = 0; t < T_COUNT; t++)
        for (auto& pr : ctx->timers[t].pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (int k = 0; k < 5; k++) if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    for (int k = 0; k < 2; k++) if (ctx->ev_sw[k]) cudaEventDestroy(ctx->ev_sw[k]);
    if (ctx->persist_prof) cudaFree(ctx->persist_prof);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->h_pcg) cudaFreeHost(ctx->h_pcg);
    if (ctx->ev_da) cudaEventDestroy(ctx->ev_da);
    if (ctx->ev_bnew) cudaEventDestroy(ctx->ev_bnew);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int vlg_ba_nccl_unique_id(void* unique_id_128)
{
    if (!nccl_load()) return fail(nullptr, VLG_BA_ENCCL, "libnccl.so.2 not found");
    nccl_uid id;
    int r = g_nccl.GetUniqueId(&id);
    if (r != 0) return fail(nullptr, VLG_BA_ENCCL, "ncclGetUniqueId failed (%d)", r);
    memcpy(unique_id_128, &id, sizeof(id));
    return VLG_BA_OK;
}

int vlg_ba_set_comm(vlg_ba_ctx* ctx, int rank, int nranks, const void* unique_id_128)
{
    if (!ctx) return VLG_BA_EINVAL;
    if (nranks <= 1) { ctx->rank = 0; ctx->nranks = 1; return VLG_BA_OK; }
    if (!nccl_load()) return fail(ctx, VLG_BA_ENCCL, "libnccl.so.2 not found");
    CU(cudaSetDevice(ctx->device));
    nccl_uid id;
    memcpy(&id, unique_id_128, sizeof(id));
    int r = g_nccl.CommInitRank(&ctx->comm, nranks, id, rank);
    if (r != 0) return fail(ctx, VLG_BA_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    ctx->rank = rank; ctx->nranks = nranks;
    return VLG_BA_OK;
}

int vlg_ba_p2p_export(vlg_ba_ctx* ctx, void* ipc_handle_64)
{
    if (!ctx || !ipc_handle_64) return VLG_BA_EINVAL;
    if (!ctx->have_problem || ctx->nranks <= 1) return fail(ctx, VLG_BA_ESTATE, "p2p_export needs set_comm (nranks > 1) and a problem");
    if (ctx->nranks > kP2pMaxRanks) return fail(ctx, VLG_BA_EINVAL, "p2p supports up to %d ranks", kP2pMaxRanks);
    CU(cudaSetDevice(ctx->device));
    p2p_close(ctx);
    const int N = ctx->na * ctx->m;
    P2PMail& mb = ctx->p2p;
    memset(&mb, 0, sizeof(mb));
    mb.nranks = ctx->nranks; mb.rank = ctx->rank;
    mb.nslot = std::max(ctx->Np, cdiv(N, kP2pThreads) * kP2pThreads);
    ctx->p2p_bytes = sizeof(uint4) * 2 * (size_t)mb.nranks * mb.nslot;
    CU(cudaMalloc(&ctx->p2p_base, ctx->p2p_bytes));
    CU(cudaMemset(ctx->p2p_base, 0, ctx->p2p_bytes));      // flag 0 = nothing sent yet (epochs start at 1)
    CU(cudaMalloc(&ctx->p2p_dev, sizeof(P2PMail)));
    CU(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->p2p_base));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    memset(ipc_handle_64, 0, VLG_BA_P2P_HANDLE_BYTES);
    memcpy(ipc_handle_64, &h, 64);
    if (ctx->use_explicit && ctx->s_split && ctx->S) {       // second handle: this rank's share of S
        CU(cudaIpcGetMemHandle(&h, ctx->S));
        memcpy((char*)ipc_handle_64 + 64, &h, 64);
    }
    return VLG_BA_OK;
}

int vlg_ba_p2p_import(vlg_ba_ctx* ctx, const void* ipc_handles)
{
    if (!ctx || !ipc_handles) return VLG_BA_EINVAL;
    if (!ctx->p2p_base) return fail(ctx, VLG_BA_ESTATE, "p2p_import needs p2p_export first");
    CU(cudaSetDevice(ctx->device));
    P2PMail& mb = ctx->p2p;
    ctx->p2p_peer_base.assign((size_t)mb.nranks, nullptr);
    for (int r = 0; r < mb.nranks; r++) {
        void* base = ctx->p2p_base;
        if (r != mb.rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)ipc_handles + VLG_BA_P2P_HANDLE_BYTES * (size_t)r, 64);
            cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return fail(ctx, VLG_BA_ECUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
        ctx->p2p_peer_base[(size_t)r] = base;
        mb.data[r] = (uint4*)base;
    }
    // the peers' shares of S (column-block pull), when every rank exported one
    ctx->peer_S_ready = false;
    if (ctx->use_explicit && ctx->s_split && ctx->S) {
        bool all = true;
        for (int r = 0; r < mb.nranks && all; r++) {
            const char* hp = (const char*)ipc_handles + VLG_BA_P2P_HANDLE_BYTES * (size_t)r + 64;
            bool nz = false;
            for (int k = 0; k < 64; k++) nz = nz || hp[k] != 0;
            all = nz;
        }
        if (all) {
            ctx->p2p_peer_S.assign((size_t)mb.nranks, nullptr);
            ctx->peer_S.nranks = mb.nranks; ctx->peer_S.rank = mb.rank;
            for (int r = 0; r < mb.nranks; r++) {
                void* base = ctx->S;
                if (r != mb.rank) {
                    cudaIpcMemHandle_t h;
                    memcpy(&h, (const char*)ipc_handles + VLG_BA_P2P_HANDLE_BYTES * (size_t)r + 64, 64);
                    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
                    if (e != cudaSuccess) return fail(ctx, VLG_BA_ECUDA, "cudaIpcOpenMemHandle(S of rank %d): %s", r, cudaGetErrorString(e));
                    ctx->p2p_peer_S[(size_t)r] = base;
                }
                ctx->peer_S.S[r] = (const double*)base;
            }
            ctx->peer_S_ready = true;
        }
    }
    CU(cudaMemcpy(ctx->p2p_dev, &mb, sizeof(P2PMail), cudaMemcpyHostToDevice));
    ctx->p2p_epoch = 0;
    ctx->p2p_ready = true;
    return VLG_BA_OK;
}

int vlg_ba_set_problem_sparse(vlg_ba_ctx* ctx, int m, int n, const double* K, const double* a, const double* b,
                              int64_t nobs, const double* obs_xy, const int32_t* obs_pt, const int32_t* obs_cam,
                              const double* pivot)
{
    if (!ctx) return VLG_BA_EINVAL;
    if (nobs > 0 && (!obs_xy || !obs_pt || !obs_cam)) return fail(ctx, VLG_BA_EINVAL, "observation list is NULL");
    return build_problem(ctx, m, n, K, a, b, nobs, obs_xy, obs_pt, obs_cam, pivot);
}

// visibility compaction: {(i,j): visible(i,j) != 0} in ascending i + n*j, exactly the cells the
// reference treats as visible (mex_bundle_1_XABeUVWeAeB.c:196)
int vlg_ba_set_problem_dense(vlg_ba_ctx* ctx, int m, int n, const double* K, const double* a, const double* b,
                             const double* X, const double* visible, const double* pivot)
{
    if (!ctx) return VLG_BA_EINVAL;
    if (m <= 0 || n < 0 || (n > 0 && (!X || !visible))) return fail(ctx, VLG_BA_EINVAL, "X and visible are required");
    std::vector<double> xy;
    std::vector<int32_t> pt, cam;
    for (int j = 0; j < m; j++)
        for (int i = 0; i < n; i++) {
            const size_t c = (size_t)i + (size_t)n * j;
            if (visible[c] != 0.0) {
                xy.push_back(X[2 * c]); xy.push_back(X[2 * c + 1]);
                pt.push_back(i); cam.push_back(j);
            }
        }
    return build_problem(ctx, m, n, K, a, b, (int64_t)pt.size(), xy.data(), pt.data(), cam.data(), pivot);
}

int vlg_ba_set_num_vis(vlg_ba_ctx* ctx, double num_vis)
{
    if (!ctx) return VLG_BA_EINVAL;
    if (!(num_vis > 0.0)) { ctx->num_vis_user = false; return VLG_BA_OK; }     // back to the all-reduced observation count
    ctx->num_vis = num_vis;
    ctx->num_vis_user = true;
    return VLG_BA_OK;
}

int64_t vlg_ba_nobs(const vlg_ba_ctx* ctx) { return ctx ? ctx->nobs : 0; }

int vlg_ba_get_obs(vlg_ba_ctx* ctx, double* obs_xy, int32_t* obs_pt, int32_t* obs_cam)
{
    if (!ctx || !ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    if (obs_xy) memcpy(obs_xy, ctx->h_obs_xy.data(), sizeof(double) * 2 * (size_t)ctx->nobs);
    if (obs_pt) memcpy(obs_pt, ctx->h_obs_pt.data(), sizeof(int32_t) * (size_t)ctx->nobs);
    if (obs_cam) memcpy(obs_cam, ctx->h_obs_cam.data(), sizeof(int32_t) * (size_t)ctx->nobs);
    return VLG_BA_OK;
}

int vlg_ba_set_state(vlg_ba_ctx* ctx, const double* a, const double* b, double lambda, double nu)
{
    if (!ctx || !ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    CU(cudaSetDevice(ctx->device));
    const size_t N = (size_t)ctx->na * ctx->m;
    if (a) { ctx->h_a.assign(a, a + N); CHK(upload(ctx, ctx->a, a, N)); ctx->rtab_valid = false; }
    if (b) CHK(upload(ctx, ctx->b, b, (size_t)3 * ctx->n));
    CU(cudaStreamSynchronize(ctx->stream));
    if (lambda > 0) ctx->lambda = lambda;
    if (nu > 0) ctx->nu = nu;
    if (a || b) ctx->s1_valid = ctx->s2_valid = ctx->s3_valid = false;
    return VLG_BA_OK;
}

int vlg_ba_get_state(vlg_ba_ctx* ctx, double* a, double* b, double* lambda, double* nu, int* iter, int* iter2)
{
    if (!ctx || !ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    CU(cudaSetDevice(ctx->device));
    CHK(download(ctx, a, ctx->a, (size_t)ctx->na * ctx->m));
    CHK(download(ctx, b, ctx->b, (size_t)3 * ctx->n));
    CU(cudaStreamSynchronize(ctx->stream));
    if (lambda) *lambda = ctx->lambda;
    if (nu) *nu = ctx->nu;
    if (iter) *iter = ctx->iter;
    if (iter2) *iter2 = ctx->iter2;
    return VLG_BA_OK;
}

int vlg_ba_stage1(vlg_ba_ctx* ctx, double* cost)
{
    if (!ctx) return VLG_BA_EINVAL;
    CHK(do_stage1(ctx));
    if (cost) *cost = ctx->old_cost;
    return VLG_BA_OK;
}

int vlg_ba_get_blocks(vlg_ba_ctx* ctx, double* U, double* V, double* W, double* eA, double* eB)
{
    if (!ctx || !ctx->s1_valid) return fail(ctx, VLG_BA_ESTATE, "stage1 has not run at the current state");
    CU(cudaSetDevice(ctx->device));
    const size_t N = (size_t)ctx->na * ctx->m;
    CHK(download(ctx, U, ctx->U, (size_t)ctx->na * N));
    CHK(download(ctx, V, ctx->V, (size_t)9 * ctx->n));
    CHK(download(ctx, W, ctx->W, (size_t)3 * ctx->na * ctx->nobs));
    CHK(download(ctx, eA, ctx->eA, N));
    CHK(download(ctx, eB, ctx->eB, (size_t)3 * ctx->n));
    CU(cudaStreamSynchronize(ctx->stream));
    return VLG_BA_OK;
}

int vlg_ba_get_jacobians(vlg_ba_ctx* ctx, double* X_hat, double* A, double* B, double* e)
{
    if (!ctx || !ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    CU(cudaSetDevice(ctx->device));
    const size_t no = (size_t)ctx->nobs, na = (size_t)ctx->na;
    double *dX = nullptr, *dA = nullptr, *dB = nullptr, *de = nullptr;
    int r = VLG_BA_OK;
    {
        cudaError_t e = cudaMalloc(&dX, sizeof(double) * std::max<size_t>(2 * no, 1));
        if (e == cudaSuccess) e = cudaMalloc(&dA, sizeof(double) * std::max<size_t>(2 * na * no, 1));
        if (e == cudaSuccess) e = cudaMalloc(&dB, sizeof(double) * std::max<size_t>(6 * no, 1));
        if (e == cudaSuccess) e = cudaMalloc(&de, sizeof(double) * std::max<size_t>(2 * no, 1));
        if (e != cudaSuccess) r = fail(ctx, VLG_BA_ENOMEM, "get_jacobians: cudaMalloc: %s", cudaGetErrorString(e));
    }
    if (r == VLG_BA_OK) r = DISPATCH_NA(ctx, run_stage1)(ctx, dX, dA, dB, de);
    if (r == VLG_BA_OK) r = download(ctx, X_hat, dX, 2 * no);
    if (r == VLG_BA_OK) r = download(ctx, A, dA, 2 * na * no);
    if (r == VLG_BA_OK) r = download(ctx, B, dB, 6 * no);
    if (r == VLG_BA_OK) r = download(ctx, e, de, 2 * no);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(dX); cudaFree(dA); cudaFree(dB); cudaFree(de);
    ctx->s1_valid = false; ctx->s2_valid = false; ctx->s3_valid = false;   // the launch rewrote local (un-reduced) sums
    return r;
}

int vlg_ba_stage2(vlg_ba_ctx* ctx, double lambda)
{
    if (!ctx) return VLG_BA_EINVAL;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    return do_stage2(ctx, lambda);
}

int vlg_ba_get_reduced(vlg_ba_ctx* ctx, double* Vinv, double* S, double* e_, double* da)
{
    if (!ctx || !ctx->s2_valid) return fail(ctx, VLG_BA_ESTATE, "stage2 has not run");
    CU(cudaSetDevice(ctx->device));
    const size_t N = (size_t)ctx->na * ctx->m;
    CHK(download(ctx, Vinv, ctx->Vinv, (size_t)9 * ctx->n));
    CHK(download(ctx, e_, ctx->ebar, N));
    CHK(download(ctx, da, ctx->da, N));
    if (S) {
        if (!ctx->use_chol) return fail(ctx, VLG_BA_ESTATE, "S is only formed on the Cholesky path");
        // note: after the factorisation the lower triangle of the device S holds L; S itself is
        // re-assembled here from the same kernels so that tests can compare it with mex2's output
        const int Np = ctx->Np;
        double* tmp = nullptr;
        CU(cudaMalloc(&tmp, sizeof(double) * (size_t)Np * Np));
        CU(cudaMemsetAsync(tmp, 0, sizeof(double) * (size_t)Np * Np, ctx->stream));
        {
            const int r = DISPATCH_NA(ctx, assemble_S)(ctx, 1, tmp, 0);
            if (r != VLG_BA_OK) { cudaFree(tmp); return r; }
        }
        cudaError_t e = cudaMemcpy2DAsync(S, sizeof(double) * N, tmp, sizeof(double) * (size_t)Np, sizeof(double) * N, N,
                                          cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(tmp);
        if (e != cudaSuccess) return fail(ctx, VLG_BA_ECUDA, "copy of S: %s", cudaGetErrorString(e));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return VLG_BA_OK;
}

int64_t vlg_ba_symv_bytes(const vlg_ba_ctx* ctx) { return ctx && ctx->use_explicit ? ctx->symv_bytes : 0; }

int vlg_ba_symv_plan(int Np, int G, int J0, int J1, const double* speed, int32_t* tiles4, int32_t* tile_ptr, int32_t* row_ptr,
                     int32_t* row_list, int32_t* col_ptr, int32_t* col_list, int64_t* sizes)
{
    return vlg_ba_symv_plan_occ(Np, G, J0, J1, speed, nullptr, tiles4, tile_ptr, row_ptr, row_list, col_ptr, col_list, sizes);
}

int vlg_ba_symv_plan_occ(int Np, int G, int J0, int J1, const double* speed, const unsigned char* occ, int32_t* tiles4, int32_t* tile_ptr,
                         int32_t* row_ptr, int32_t* row_list, int32_t* col_ptr, int32_t* col_list, int64_t* sizes)
{
    if (Np <= 0 || Np % kSymvCols != 0 || G <= 0 || J0 < 0 || J1 > Np / kSymvCols || J0 > J1 || !sizes)
        return fail(nullptr, VLG_BA_EINVAL, "symv_plan: Np must be a positive multiple of 32, 0 <= J0 <= J1 <= Np/32, G > 0");
    std::vector<vlg_ba_ctx::SymvTileH> seq;
    std::vector<double> cum, sp((size_t)G, 1.0);
    int ncell = 0;
    symv_sequence(Np, J0, J1, seq, cum, ncell, occ);
    if (speed) sp.assign(speed, speed + G);
    SymvPlan pl;
    symv_cut(Np, G, seq, cum, sp, pl);
    sizes[0] = (int64_t)pl.tiles.size(); sizes[1] = pl.nfrag; sizes[2] = (int64_t)pl.rlist.size(); sizes[3] = (int64_t)pl.clist.size();
    sizes[4] = (int64_t)pl.rptr.size() - 1; sizes[5] = kSymvBlkRows; sizes[6] = kSymvSlab; sizes[7] = ncell;
    if (tiles4) for (size_t t = 0; t < pl.tiles.size(); t++) { tiles4[4 * t] = pl.tiles[t].x; tiles4[4 * t + 1] = pl.tiles[t].y; tiles4[4 * t + 2] = pl.tiles[t].z; tiles4[4 * t + 3] = pl.tiles[t].w; }
    if (tile_ptr) std::copy(pl.tptr.begin(), pl.tptr.end(), tile_ptr);
    if (row_ptr) std::copy(pl.rptr.begin(), pl.rptr.end(), row_ptr);
    if (row_list) std::copy(pl.rlist.begin(), pl.rlist.end(), row_list);
    if (col_ptr) std::copy(pl.cptr.begin(), pl.cptr.end(), col_ptr);
    if (col_list) std::copy(pl.clist.begin(), pl.clist.end(), col_list);
    return VLG_BA_OK;
}

int vlg_ba_reproj_errors(vlg_ba_ctx* ctx, double depth_max, double* err, double* depth, double* mean_err, double* max_sq_err,
                         int64_t* argmax, int64_t* n_bad_depth)
{
    if (!ctx || !ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    CU(cudaSetDevice(ctx->device));
    const size_t no = (size_t)ctx->nobs;
    double *d_err = nullptr, *d_depth = nullptr, *d_bad = nullptr, *d_sc = nullptr;
    long long* d_idx = nullptr;
    auto body = [&]() -> int {
        CU(cudaMalloc(&d_err, sizeof(double) * std::max<size_t>(no, 1)));
        CU(cudaMalloc(&d_depth, sizeof(double) * std::max<size_t>(no, 1)));
        CU(cudaMalloc(&d_bad, sizeof(double) * std::max<size_t>(no, 1)));
        CU(cudaMalloc(&d_sc, sizeof(double) * 4));
        CU(cudaMalloc(&d_idx, sizeof(long long)));
        CHK(DISPATCH_NA(ctx, run_reproj_errors)(ctx, depth_max, d_err, d_depth, d_bad));
        CHK(reduce_to(ctx, d_err, no, d_sc));
        CHK(reduce_to(ctx, d_bad, no, d_sc + 1));
        k_argmax_sq<<<1, 1024, 0, ctx->stream>>>((int64_t)no, d_err, d_bad, d_sc + 2, d_idx);
        ctx->launches++;
        CU(cudaGetLastError());
        double h[4];
        long long hi = -1;
        CHK(download(ctx, h, d_sc, 3));
        CU(cudaMemcpyAsync(&hi, d_idx, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CHK(download(ctx, err, d_err, no));
        CHK(download(ctx, depth, d_depth, no));
        CU(cudaStreamSynchronize(ctx->stream));
        if (mean_err) *mean_err = no ? h[0] / (double)no : 0.0;
        if (n_bad_depth) *n_bad_depth = (int64_t)(h[1] + 0.5);
        if (max_sq_err) *max_sq_err = h[2];
        if (argmax) *argmax = (int64_t)hi;
        return VLG_BA_OK;
    };
    const int r = body();
    cudaFree(d_err); cudaFree(d_depth); cudaFree(d_bad); cudaFree(d_sc); cudaFree(d_idx);
    return r;
}

int vlg_ba_set_da(vlg_ba_ctx* ctx, const double* da)
{
    if (!ctx || !ctx->s2_valid) return fail(ctx, VLG_BA_ESTATE, "stage2 has not run");
    CU(cudaSetDevice(ctx->device));
    CHK(upload(ctx, ctx->da, da, (size_t)ctx->na * ctx->m));
    CU(cudaStreamSynchronize(ctx->stream));
    return VLG_BA_OK;
}

int vlg_ba_stage3(vlg_ba_ctx* ctx, double lambda, double* new_cost, double* denom)
{
    if (!ctx) return VLG_BA_EINVAL;
    double nc = 0, dn = 0;
    CHK(do_stage3(ctx, lambda, &nc, &dn));
    if (new_cost) *new_cost = nc;
    if (denom) *denom = dn;
    return VLG_BA_OK;
}

int vlg_ba_get_update(vlg_ba_ctx* ctx, double* db, double* a_new, double* b_new)
{
    if (!ctx || !ctx->s3_valid) return fail(ctx, VLG_BA_ESTATE, "stage3 has not run");
    CU(cudaSetDevice(ctx->device));
    CHK(download(ctx, db, ctx->db, (size_t)3 * ctx->n));
    CHK(download(ctx, a_new, ctx->a_new, (size_t)ctx->na * ctx->m));
    CHK(download(ctx, b_new, ctx->b_new, (size_t)3 * ctx->n));
    CU(cudaStreamSynchronize(ctx->stream));
    return VLG_BA_OK;
}

int vlg_ba_trial_step(vlg_ba_ctx* ctx, vlg_ba_trial_info* info)
{
    if (!ctx || !ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    return do_trial(ctx, info);
}

int vlg_ba_lm_reset(vlg_ba_ctx* ctx, const double* a, const double* b)
{
    if (!ctx || !ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    if (a || b) CHK(vlg_ba_set_state(ctx, a, b, -1.0, -1.0));
    ctx->lambda = ctx->opt.lambda0; ctx->nu = ctx->opt.nu0; ctx->iter = 1; ctx->iter2 = 0;    // bundle_euclid.m:111-119
    ctx->err_hist.clear();
    return VLG_BA_OK;
}

int vlg_ba_lm_continue(const vlg_ba_ctx* ctx)
{
    if (!ctx || !ctx->have_problem) return 0;
    const vlg_ba_opts& o = ctx->opt;
    const int it = ctx->iter;
    bool go = it < o.max_iter && ctx->iter2 < o.max_iter2;                                      // :120-123
    if (go && it >= 3) {
        const std::vector<double>& err = ctx->err_hist;
        go = err[it - 1] > o.abs_tol && (err[it - 2] - err[it - 1]) > o.rel_tol * err[it - 2];
    }
    return go ? 1 : 0;
}

int vlg_ba_solve(vlg_ba_ctx* ctx, double* K_, double* Te_, double* w_, double* Xe_, const double* Xe4, double* error_,
                 int* n_error)
{
    if (!ctx || !ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    const vlg_ba_opts& o = ctx->opt;
    CHK(vlg_ba_lm_reset(ctx, nullptr, nullptr));
    while (vlg_ba_lm_continue(ctx)) {
        const int it = ctx->iter;
        vlg_ba_trial_info info;
        CHK(do_trial(ctx, &info));
        if (info.accepted && o.verbose) printf("iter %d: error= %.5g -> %.5g\n", it, ctx->err_hist[it - 1], ctx->err_hist[it]);
    }
    const bool any = !ctx->err_hist.empty();
    const std::vector<double>& err = ctx->err_hist;
    const int m = ctx->m, n = ctx->n, na = ctx->na;
    std::vector<double> a((size_t)na * m), b((size_t)3 * n);
    CHK(vlg_ba_get_state(ctx, a.data(), b.data(), nullptr, nullptr, nullptr, nullptr));
    for (int j = 0; j < m; j++) {                                                // :256-267
        if (K_) {
            for (int k = 0; k < 4; k++) K_[4 * (size_t)j + k] = ctx->h_K[4 * (size_t)j + k];
            if (o.num_variableK == 1) { K_[4 * (size_t)j] = a[(size_t)na * j + 6]; K_[4 * (size_t)j + 1] = a[(size_t)na * j + 6]; }
            if (o.num_variableK == 4) for (int k = 0; k < 4; k++) K_[4 * (size_t)j + k] = a[(size_t)na * j + 6 + k];
        }
        for (int k = 0; k < 3; k++) {
            if (w_) w_[3 * (size_t)j + k] = a[(size_t)na * j + k];
            if (Te_) Te_[3 * (size_t)j + k] = a[(size_t)na * j + 3 + k];
        }
    }
    if (Xe_)
        for (int i = 0; i < n; i++) {
            for (int k = 0; k < 3; k++) Xe_[4 * (size_t)i + k] = b[3 * (size_t)i + k];
            Xe_[4 * (size_t)i + 3] = Xe4 ? Xe4[i] : 1.0;
        }
    const int ne = any ? ctx->iter : 0;
    if (error_) for (int k = 0; k < ne; k++) error_[k] = err[k];
    if (n_error) *n_error = ne;
    return VLG_BA_OK;
}

int vlg_ba_trial_step_host(vlg_ba_ctx* ctx, const double* a, const double* b, const double* obs_xy, double lambda,
                           double* a_new, double* b_new, vlg_ba_trial_info* info)
{
    if (!ctx || !ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    CU(cudaSetDevice(ctx->device));
    if (obs_xy && ctx->nobs > 0) {
        CHK(upload(ctx, (double*)ctx->obs_xy, obs_xy, 2 * (size_t)ctx->nobs));
    }
    // new state without a host synchronisation: the caller's buffers stay untouched until this call returns, and the
    // rotation table is computed from the host copy of a while the observations are still on their way to the device
    {
        const size_t N = (size_t)ctx->na * ctx->m;
        if (a) { ctx->h_a.assign(a, a + N); CHK(upload(ctx, ctx->a, a, N)); ctx->rtab_valid = false; }
        if (b) CHK(upload(ctx, ctx->b, b, (size_t)3 * ctx->n));
        if (lambda > 0) ctx->lambda = lambda;
        ctx->s1_valid = ctx->s2_valid = ctx->s3_valid = false;
        if (!ctx->rtab_valid && ctx->opt.rtable == VLG_BA_RTABLE_HOST_LIBM) {
            CHK(DISPATCH_NA(ctx, run_rtab)(ctx, ctx->h_a, ctx->a, 4, ctx->h_rtab, ctx->rtab));
            ctx->rtab_valid = true;
        }
    }
    // the candidate always lands in (a_new, b_new) on the device; b_new (the big one) leaves over the copy stream as soon as
    // the back-substitution has produced it, when the destination is pinned memory
    const double lam = ctx->lambda;
    {
        cudaPointerAttributes at;
        const bool pinned = b_new && cudaPointerGetAttributes(&at, b_new) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        ctx->d2h_bnew = pinned ? b_new : nullptr;
    }
    CHK(do_stage1(ctx, true));
    CHK(do_stage2(ctx, lam));
    double nc = 0, dn = 0;
    const bool overlapped = ctx->d2h_bnew != nullptr;
    const int r3 = do_stage3(ctx, lam, &nc, &dn);
    ctx->d2h_bnew = nullptr;
    CHK(r3);
    CHK(resolve_cost(ctx));
    CHK(download(ctx, a_new, ctx->a_new, (size_t)ctx->na * ctx->m));
    if (!overlapped) CHK(download(ctx, b_new, ctx->b_new, (size_t)3 * ctx->n));
    CU(cudaStreamSynchronize(ctx->stream));
    if (overlapped) CU(cudaStreamSynchronize(ctx->copy_stream));
    if (info) {
        memset(info, 0, sizeof(*info));
        info->old_cost = ctx->old_cost; info->new_cost = nc; info->denom = dn; info->rho = (ctx->old_cost - nc) / dn;
        info->lambda_used = lam; info->accepted = (ctx->old_cost - nc) > 0;
        CHK(resolve_pcg_info(ctx));
        info->solver_used = ctx->last_solver; info->pcg_iters = ctx->last_pcg_iters; info->pcg_relres = ctx->last_pcg_relres;
    }
    return VLG_BA_OK;
}

int vlg_ba_get_schur_structure(vlg_ba_ctx* ctx, int64_t* n_blocks, int32_t* blk_j, int32_t* blk_k)
{
    if (!ctx || !ctx->have_problem) return fail(ctx, VLG_BA_ESTATE, "no problem set");
    if (!ctx->use_chol) return fail(ctx, VLG_BA_ESTATE, "block structure is only built on the Cholesky path");
    if (n_blocks) *n_blocks = ctx->nblocks;
    if (blk_j) memcpy(blk_j, ctx->h_blk_j.data(), sizeof(int32_t) * (size_t)ctx->nblocks);
    if (blk_k) memcpy(blk_k, ctx->h_blk_k.data(), sizeof(int32_t) * (size_t)ctx->nblocks);
    return VLG_BA_OK;
}

int vlg_ba_selftest_quotients(int device, int64_t nsamples, uint64_t seed, int64_t* mismatches)
{
    vlg_ba_ctx* ctx = nullptr;
    if (!mismatches || nsamples < 0) return fail(nullptr, VLG_BA_EINVAL, "selftest_quotients: bad arguments");
    if (device >= 0) CU(cudaSetDevice(device));
    unsigned long long* d_bad = nullptr;
    CU(cudaMalloc(&d_bad, sizeof(unsigned long long)));
    const int threads = 256, blocks = 148 * 8, per = (int)((nsamples + (int64_t)threads * blocks - 1) / ((int64_t)threads * blocks));
    cudaError_t e = cudaMemset(d_bad, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) {
        k_selftest_quotients<<<blocks, threads>>>((unsigned long long)seed, per, d_bad);
        e = cudaGetLastError();
    }
    unsigned long long h = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&h, d_bad, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    if (e != cudaSuccess) return fail(nullptr, VLG_BA_ECUDA, "selftest_quotients: %s", cudaGetErrorString(e));
    *mismatches = (int64_t)h;
    return VLG_BA_OK;
}

int64_t vlg_ba_kernel_launches(const vlg_ba_ctx* ctx) { return ctx ? ctx->launches : 0; }

int vlg_ba_reset_timers(vlg_ba_ctx* ctx, int enable)
{
    if (!ctx) return VLG_BA_EINVAL;
    resolve_timers(ctx);
    for (int t = 0; t < T_COUNT; t++) { ctx->timers[t].total_ms = 0.0; ctx->timers[t].count = 0; }
    ctx->timers_on = enable != 0;
    return VLG_BA_OK;
}

int vlg_ba_kernel_time(vlg_ba_ctx* ctx, const char* name, double* avg_ms, int64_t* count)
{
    if (!ctx || !name) return VLG_BA_EINVAL;
    resolve_timers(ctx);
    for (int t = 0; t < T_COUNT; t++)
        if (strcmp(name, kTimerNames[t]) == 0) {
            if (avg_ms) *avg_ms = ctx->timers[t].count ? ctx->timers[t].total_ms / (double)ctx->timers[t].count : 0.0;
            if (count) *count = ctx->timers[t].count;
            return VLG_BA_OK;
        }
    return fail(ctx, VLG_BA_EINVAL, "unknown kernel group '%s'", name);
}

int vlg_ba_timer_start(vlg_ba_ctx* ctx)
{
    if (!ctx) return VLG_BA_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaEventRecord(ctx->ev_sw[0], ctx->stream));
    return VLG_BA_OK;
}

int vlg_ba_timer_stop(vlg_ba_ctx* ctx, float* elapsed_ms)
{
    if (!ctx) return VLG_BA_EINVAL;
    CU(cudaEventRecord(ctx->ev_sw[1], ctx->stream));
    CU(cudaEventSynchronize(ctx->ev_sw[1]));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev_sw[0], ctx->ev_sw[1]));
    if (elapsed_ms) *elapsed_ms = ms;
    return VLG_BA_OK;
}

#include "vlg_ba_dense.inl"

}  // extern "C"

#include "vlg_ba_batch.inl"
