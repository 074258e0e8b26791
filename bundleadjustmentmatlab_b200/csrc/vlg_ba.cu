// vlg_ba.cu -- host side of libvlgba.so: context, problem set-up, the LM driver of
// bundle_euclid.m:111-267 and the C ABI declared in include/vlg_ba.h.
//
// Host code is plain C-style C++ calling the CUDA kernels of ba_kernels.cuh; no torch, no
// CPU fallback (vlg_ba_create fails without a device).  NCCL is bound lazily with dlopen so
// that the single-GPU library has no NCCL dependency.
#include "../../include/vlg_ba.h"
#include "ba_kernels.cuh"
#include "ba_chol.cuh"
#include "ba_pcg.cuh"
#include "ba_schur.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

using namespace vlgba;

// ------------------------------------------------------------------------------------------
// NCCL, bound at run time
// ------------------------------------------------------------------------------------------
namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
const int kNcclFloat64 = 8, kNcclSum = 0;

bool nccl_load()
{
    if (g_nccl.h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return false;
    g_nccl.GetUniqueId = (int (*)(nccl_uid*))dlsym(g_nccl.h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(nccl_comm*, int, nccl_uid, int))dlsym(g_nccl.h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t))dlsym(g_nccl.h, "ncclAllReduce");
    g_nccl.Reduce = (int (*)(const void*, void*, size_t, int, int, int, nccl_comm, cudaStream_t))dlsym(g_nccl.h, "ncclReduce");
    g_nccl.GroupStart = (int (*)())dlsym(g_nccl.h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(g_nccl.h, "ncclGroupEnd");
    g_nccl.CommDestroy = (int (*)(nccl_comm))dlsym(g_nccl.h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.h, "ncclGetErrorString");
    return g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy;
}

thread_local char g_create_error[512] = "";     // failures without a context (create, stateless entries): per host thread

enum TimerId { T_STAGE1_CAM = 0, T_STAGE1_PT, T_VINV, T_SCHUR, T_SCHUR_BLK, T_CHOL, T_SWEEP_PT, T_SWEEP_CAM, T_STAGE3, T_PCG_UPDATE, T_W_COPY, T_SYMV, T_PRECOND, T_PCG_PERSIST, T_COUNT };
const char* kTimerNames[T_COUNT] = {"stage1_cam", "stage1_pt", "vinv", "schur", "schur_blocks", "chol", "pcg_sweep_pt",
                                    "pcg_sweep_cam", "stage3", "pcg_update", "w_copy", "pcg_symv", "precond", "pcg_persistent"};

struct KTimer {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    double total_ms = 0.0;
    int64_t count = 0;
};

}  // namespace

struct vlg_ba_ctx {
    vlg_ba_opts opt;
    int device = 0;
    cudaStream_t stream = nullptr;
    char err[512] = "";
    // problem
    int m = 0, n = 0, na = 6;
    int64_t nobs = 0;
    double num_vis = 0.0;
    bool have_problem = false;
    int nchunks = 0, chunk_size = 256;
    bool use_chol = false;
    bool use_explicit = false;        // PCG on the assembled dense S (k_symv_lower) instead of the two W sweeps
    int4* symv_tiles = nullptr;       // tile list of k_symv_lower, grouped per CTA
    int* symv_tile_ptr = nullptr;
    int symv_grid = 0, symv_nfrag = 0;
    double *symv_rowpart = nullptr, *symv_colpart = nullptr;   // [nfrag][kSymvBlkRows], [nfrag][32 kSymvSlab]
    int *symv_row_ptr = nullptr, *symv_row_list = nullptr, *symv_col_ptr = nullptr, *symv_col_list = nullptr;   // fold lists (SymvFold)
    // the tile sequence and its cost prefix (host), the per-CTA speed weights the cut uses, and their measurement
    struct SymvTileH { int strip, r0, rows, cell, sl; };
    std::vector<SymvTileH> h_symv_seq;
    std::vector<double> h_symv_cum, symv_speed;
    std::vector<int> h_symv_tptr, symv_smid;
    int symv_ncell = 0, symv_learn_left = 0;
    int64_t symv_bytes = 0;                   // bytes of S one matvec of this rank streams (kept tiles only)
    long long* symv_stat = nullptr;           // device [2 G]: matvec clocks of the last solve, SM id, per CTA
    unsigned int* persist_bar = nullptr;   // grid barrier counter of k_pcg_persistent
    unsigned int* init_bar = nullptr;      // ... of k_pcg_init_defl_coop
    double* init_part = nullptr;           // its block partials [17 x coop_grid]
    int init_coop_cap = -1;                // CTAs of k_pcg_init_defl_coop that can be co-resident (-1: not asked yet)
    bool persist_ok = false;
    bool s_split = false;                  // multi-GPU: S is reduce-scattered by column blocks, each rank multiplies its own
    int s_J0 = 0, s_J1 = 0;                // this rank's strips [J0, J1)
    std::vector<int> s_bounds;             // strip boundaries of all ranks (nranks + 1), equal lower-triangle areas
    bool S_zeroed = false;                 // explicit-S PCG: structurally zero blocks of S cleared for this problem
    double *Cblk = nullptr, *McL = nullptr;   // cluster-Jacobi preconditioner: gathered diagonal blocks of S, their inverses [coop_grid][128 x 128]
    double* Mc2 = nullptr;                    // overlapping variant: composite rows of the shifted partition's inverses [coop_grid][128 x 128]
    bool overlap_ok = true;                   // VLG_BA_OVERLAP=0 switches the second partition off
    int Np = 0;               // padded order of S
    int64_t nblocks = 0, npairs = 0;
    // host copies
    std::vector<double> h_K, h_a, h_a_new, h_rtab, h_rtab_new, h_rtab_next, h_da;
    // host-libm rotation tables: `rtab` is valid for the current a (skip the recomputation), `rtab_next` holds the 4-matrix
    // table of the candidate a_new, computed on the host while the back-substitution runs and swapped in on accept
    bool rtab_valid = false, rtab_next_valid = false;
    double* rtab_next = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_da = nullptr, ev_bnew = nullptr;
    double* d2h_bnew = nullptr;        // pinned host destination of b_new for this step (vlg_ba_trial_step_host), fetched while the cost is computed
    std::vector<int32_t> h_obs_pt, h_obs_cam;
    std::vector<double> h_obs_xy;
    std::vector<int32_t> h_blk_j, h_blk_k;
    // device: structure
    double2* obs_xy = nullptr;
    int *obs_pt = nullptr, *obs_cam = nullptr, *cam_ptr = nullptr, *pt_ptr = nullptr, *pt_obs = nullptr, *pt_cam = nullptr;
    int *chunk_cam = nullptr, *chunk_begin = nullptr, *chunk_end = nullptr, *cam_chunk_ptr = nullptr;
    unsigned char* cam_fixed = nullptr;
    int *blk_j = nullptr, *blk_k = nullptr;
    int64_t* blk_ptr = nullptr;
    int2* pairs = nullptr;
    int *blk_heavy = nullptr, *blk_light = nullptr;   // S-assembly work lists: blocks with many / few pairs
    int nheavy = 0, nlight = 0;
    // chunk kernel of ba_schur.cuh (num_a = 6, chunked order): per-point records, segments of the heavy blocks
    bool schur_chunk_ok = false;
    double *VE = nullptr, *Hpart = nullptr;
    int *chunk_seg_ptr = nullptr, *sblk = nullptr, *sblk_hptr = nullptr;
    int2* stile_meta = nullptr;       // the chunk kernel's own tiles of the camera segments: (begin, nob), nob <= kSchurTile
    int* cam_stile_ptr = nullptr;     // [m + 1]
    int nstiles = 0;
    int4* segs = nullptr;
    int nsblk = 0, nsegs = 0;
    double* Ybuf = nullptr;                           // Y = W V*^-1 per observation (C-order), explicit-S paths
    double* red2_local = nullptr;                     // this rank's per-camera Schur sums before the all-reduce
    int4* ptile_meta = nullptr;      // point tiles of the PCG point sweep: (q0, nob, p0, npts)
    int4* s1tile_meta = nullptr;     // point tiles of the stage-1 point pass (<= kS1Tile observations)
    int ns1tiles = 0;
    int2* chunk_meta = nullptr;      // camera chunks: (begin, nob)
    int nptiles = 0;
    bool tiled_ok = false;           // every track fits one tile
    int use_ring = 0;                // persistent ring sweeps (ba_pcg.cuh): bit 0 = camera sweep, bit 1 = point sweep
    int pt_tile = 0, nsm = 148;
    double *Wp = nullptr, *blkpart = nullptr;
    double *Zd = nullptr, *SZd = nullptr;   // deflation vectors [4][N] and S*Z
    DeflScalars* defl_sc = nullptr;
    int coop_grid = 0;               // 0 = cooperative update kernel not usable
    // device: parameters
    double *K4 = nullptr, *a = nullptr, *b = nullptr, *a_new = nullptr, *b_new = nullptr, *rtab = nullptr, *rtab_new = nullptr;
    // device: stage 1 (red1 = U | eA | cost | nvis contiguous for one all-reduce)
    double *red1 = nullptr, *U = nullptr, *eA = nullptr, *scal1 = nullptr;
    double *W = nullptr, *Upart = nullptr, *V = nullptr, *eB = nullptr, *cost_pt = nullptr, *red_part = nullptr;
    double* BeP = nullptr;            // [nobs][8] P-order: B | e of every observation (stage 1: camera pass -> point pass)
    int* obs_slot = nullptr;          // C-order position -> P-order slot (inverse of pt_obs)
    // device: stage 2
    double *Ud = nullptr, *Vinv = nullptr, *Spart = nullptr, *S = nullptr;
    double *chol_R = nullptr, *chol_Ld = nullptr, *chol_Dinv = nullptr;
    unsigned int* chol_bar = nullptr;   // k_chol_coop: right-hand-side row tiles, factored diagonal tiles
    int chol_grid = 0;
    double *red2 = nullptr, *Sjj = nullptr, *ebar = nullptr;   // red2 = per-camera sums [m][NU], all-reduced
    double *Minv = nullptr, *da = nullptr;
    double *pr = nullptr, *pz = nullptr, *pp = nullptr, *pp2 = nullptr, *pq = nullptr, *wq = nullptr, *tvec = nullptr, *qpart = nullptr;
    PcgScalars* pcg_sc = nullptr;
    // device: stage 3 (scal3 = new_cost | denom_pt_sum | denom_cam)
    double *db = nullptr, *denom_pt = nullptr, *cost_obs = nullptr, *scal3 = nullptr;
    double* xhat_out = nullptr;      // optional per-observation X_hat_new (dense mex3 drop-in)
    // pinned host scalars
    double* h_pin = nullptr;
    PcgScalars* h_pcg = nullptr;
    // LM state
    double lambda = 1e-3, nu = 2.0;
    int iter = 1, iter2 = 0;
    std::vector<double> err_hist;     // error_ of bundle_euclid.m:119,229-231
    bool s1_valid = false, s2_valid = false, s3_valid = false;
    double old_cost = 0.0;
    double s2_lambda = 0.0;
    int last_solver = 0, last_pcg_iters = 0;
    bool cost_pending = false;         // stage 1's cost and visible count are still on their way to h_pin[4..5]
    bool pcg_pending = false;          // the solve's scalars are still on their way to h_pcg (no host sync after the persistent kernel)
    double last_pcg_relres = 0.0;
    // multi-GPU
    nccl_comm comm = nullptr;
    int rank = 0, nranks = 1;
    // peer-memory all-reduce of the PCG vector (vlg_ba_p2p_export / _import)
    void* p2p_base = nullptr;             // this rank's mailbox: data then flags
    size_t p2p_bytes = 0;
    std::vector<void*> p2p_peer_base;     // opened peer mailboxes (own entry = p2p_base)
    P2PMail p2p;
    P2PMail* p2p_dev = nullptr;           // device copy (argument of k_symv_finish)
    PeerS peer_S;                         // peers' S allocations (IPC), for the column-block pull
    bool peer_S_ready = false;
    std::vector<void*> p2p_peer_S;        // opened peer S mappings
    bool p2p_ready = false;
    unsigned int p2p_epoch = 0;
    long long* persist_prof = nullptr;    // VLG_BA_PERSIST_PROF: in-kernel clock64 profile of k_pcg_persistent (per context)
    bool num_vis_user = false;            // vlg_ba_set_num_vis: the caller's num_vis (bundle_euclid.m:82) overrides the all-reduced count
    // accounting
    int64_t launches = 0;
    bool timers_on = false;
    KTimer timers[T_COUNT];
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_sw[2] = {nullptr, nullptr};
    std::vector<void*> allocs;
};

namespace {

int fail(vlg_ba_ctx* c, int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : g_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, VLG_BA_ECUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

#define CHK(call)                        \
    do {                                 \
        int r_ = (call);                 \
        if (r_ != VLG_BA_OK) return r_;  \
    } while (0)

template <class T>
int dalloc(vlg_ba_ctx* ctx, T** p, size_t count)
{
    *p = nullptr;
    if (count == 0) count = 1;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) return fail(ctx, VLG_BA_ENOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
    ctx->allocs.push_back(q);
    *p = (T*)q;
    return VLG_BA_OK;
}

int allreduce(vlg_ba_ctx* ctx, double* buf, size_t count);

// closes every peer mapping (mailboxes, shares of S) and frees this rank's mailbox
void p2p_close(vlg_ba_ctx* ctx)
{
    for (void* q : ctx->p2p_peer_S)
        if (q) cudaIpcCloseMemHandle(q);
    ctx->p2p_peer_S.clear();
    ctx->peer_S_ready = false;
    for (size_t r = 0; r < ctx->p2p_peer_base.size(); r++)
        if (ctx->p2p_peer_base[r] && ctx->p2p_peer_base[r] != ctx->p2p_base) cudaIpcCloseMemHandle(ctx->p2p_peer_base[r]);
    ctx->p2p_peer_base.clear();
    if (ctx->p2p_base) cudaFree(ctx->p2p_base);
    if (ctx->p2p_dev) cudaFree(ctx->p2p_dev);
    ctx->p2p_base = nullptr; ctx->p2p_dev = nullptr; ctx->p2p_ready = false; ctx->p2p_epoch = 0;
}


void free_problem(vlg_ba_ctx* ctx)
{
    // peer-memory state is sized by (and, for S, points into) the problem: close the peers' mappings first; on a live
    // multi-rank context (set_problem_* is collective there) nobody may free an exported allocation before every importer
    // has closed it, so the ranks meet in a 1-element all-reduce in between
    const bool had_p2p = ctx->p2p_ready || ctx->peer_S_ready || ctx->p2p_base;
    p2p_close(ctx);
    if (had_p2p && ctx->nranks > 1 && ctx->comm && ctx->scal3) {
        allreduce(ctx, ctx->scal3 + 3, 1);
        cudaStreamSynchronize(ctx->stream);
    }
    ctx->scal3 = nullptr;
    for (void* p : ctx->allocs) cudaFree(p);
    ctx->allocs.clear();
    ctx->have_problem = false;
    ctx->p2p_ready = false;           // the mailbox is sized by the problem: export/import again
    ctx->s1_valid = ctx->s2_valid = ctx->s3_valid = false;
}

template <class T>
int upload(vlg_ba_ctx* ctx, T* dst, const T* src, size_t count)
{
    if (count) CU(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return VLG_BA_OK;
}

template <class T>
int download(vlg_ba_ctx* ctx, T* dst, const T* src, size_t count)
{
    if (count && dst) CU(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return VLG_BA_OK;
}

constexpr int kSchurHeavy = 16;   // blocks with more pairs than this get a warp, the others a thread

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- kernel timers ------------------------------------------------------------------------
struct TimedScope {
    vlg_ba_ctx* c;
    int id;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    TimedScope(vlg_ba_ctx* ctx, int tid) : c(ctx), id(tid)
    {
        if (!c->timers_on) return;
        auto get = [&]() {
            cudaEvent_t e;
            if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        e0 = get(); e1 = get();
        cudaEventRecord(e0, c->stream);
    }
    ~TimedScope()
    {
        if (!e0) return;
        cudaEventRecord(e1, c->stream);
        c->timers[id].pending.push_back({e0, e1});
    }
};

void resolve_timers(vlg_ba_ctx* c)
{
    cudaStreamSynchronize(c->stream);
    for (int t = 0; t < T_COUNT; t++) {
        for (auto& pr : c->timers[t].pending) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
                c->timers[t].total_ms += ms;
                c->timers[t].count += 1;
            }
            c->ev_pool.push_back(pr.first);
            c->ev_pool.push_back(pr.second);
        }
        c->timers[t].pending.clear();
    }
}

// ---- host rotation table: vl_rodrigues with the host libm (bit parity with the CPU reference) ----
// Same statement as the reference's dependency (VLFeat vl_rodrigues, call site
// reproject_point.h:44); compiled with -ffp-contract=off.
void rodrigues_host(double w0, double w1, double w2, double* R)
{
    double th = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    if (th < 1e-6) {
        R[0] = 1.0; R[3] = 0.0; R[6] = 0.0;
        R[1] = 0.0; R[4] = 1.0; R[7] = 0.0;
        R[2] = 0.0; R[5] = 0.0; R[8] = 1.0;
        return;
    }
    double x = w0 / th, y = w1 / th, z = w2 / th;
    double xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
    double sth = sin(th), cth = cos(th), mcth = 1.0 - cth;
    R[0] = 1.0 - mcth * (yy + zz);
    R[1] = sth * z + mcth * xy;
    R[2] = -sth * y + mcth * xz;
    R[3] = -sth * z + mcth * xy;
    R[4] = 1.0 - mcth * (zz + xx);
    R[5] = sth * x + mcth * yz;
    R[6] = sth * y + mcth * xz;
    R[7] = -sth * x + mcth * yz;
    R[8] = 1.0 - mcth * (xx + yy);
}

// matrices k0 .. nmat-1 of every camera (k = 0: base, k = 1..3: rotation component k-1 perturbed by h)
void rtab_host(int m, int na, const double* a, int nmat, double* out, int k0 = 0)
{
    const double h = 1e-10;
    // serial on purpose: a host thread team per rank (8 ranks x 8 spinning libgomp threads) cost 20 ms per LM step on
    // the 8-GPU box, while the table itself is ~0.1 ms at Venice shape
    for (int j = 0; j < m; j++) {
        const double* w = a + (size_t)na * j;
        for (int k = k0; k < nmat; k++) {
            double w0 = w[0], w1 = w[1], w2 = w[2];
            if (k == 1) w0 = w0 + h;
            if (k == 2) w1 = w1 + h;
            if (k == 3) w2 = w2 + h;
            rodrigues_host(w0, w1, w2, out + ((size_t)j * nmat + k) * 9);
        }
    }
}

// ---- all-reduce helper -----------------------------------------------------------------------
int allreduce(vlg_ba_ctx* ctx, double* buf, size_t count)
{
    if (ctx->nranks <= 1 || count == 0) return VLG_BA_OK;
    int r = g_nccl.AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
    if (r != 0) return fail(ctx, VLG_BA_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return VLG_BA_OK;
}

int reduce_to(vlg_ba_ctx* ctx, const double* in, size_t n, double* out)
{
    k_reduce_partial<<<kRedBlocks, kRedThreads, 0, ctx->stream>>>(in, n, ctx->red_part);
    k_reduce_final<<<1, kRedThreads, 0, ctx->stream>>>(ctx->red_part, kRedBlocks, out);
    ctx->launches += 2;
    CU(cudaGetLastError());
    return VLG_BA_OK;
}

// ---- problem set-up ----------------------------------------------------------------------------
// Host-side plan of the explicit-S matvec (pure functions: also reachable without a GPU through vlg_ba_symv_plan, which the
// CPU tests use to replay the decomposition).  symv_sequence: the tiles of the strips [J0, J1) of the lower triangle in
// cell order with their cost prefix.  symv_cut: one contiguous piece per CTA, cost proportional to the CTA's speed weight,
// and what depends on the cut -- tile flags, fragments, fold lists (ba_pcg.cuh, SymvSmem).
struct SymvPlan {
    std::vector<int4> tiles;
    std::vector<int> tptr, rptr, rlist, cptr, clist;
    int nfrag = 0;
};

// `occ` (optional, [ceil(Np / 256)][Np / 32], row-major): 1 where the 256-row slot x 32-column strip of the lower triangle
// holds a structurally non-zero block of S; tiles of empty slots are left out of the sequence -- they would stream zeros
// (a banded scene, i.e. cameras that only share points with their neighbours, has most of the triangle empty).
void symv_sequence(int Np, int J0, int J1, std::vector<vlg_ba_ctx::SymvTileH>& seq, std::vector<double>& cum, int& ncell,
                   const unsigned char* occ = nullptr)
{
    const int nstrips_all = Np / kSymvCols;
    seq.clear();
    ncell = 0;
    for (int Js = J0; Js < J1; Js += kSymvSlab) {
        const int Je = std::min(J1, Js + kSymvSlab);
        for (int b0 = (kSymvCols * Js) / kSymvBlkRows * kSymvBlkRows; b0 < Np; b0 += kSymvBlkRows, ncell++)
            for (int J = Js; J < Je; J++) {
                int lo = std::max(kSymvCols * J, b0);
                const int hi = std::min(Np, b0 + kSymvBlkRows);
                while (lo < hi) {
                    const int nx = std::min(hi, (lo / kSymvRows + 1) * kSymvRows);
                    if (!occ || occ[(size_t)(lo / kSymvRows) * nstrips_all + J]) seq.push_back({J, lo, nx - lo, ncell, J - Js});
                    lo = nx;
                }
            }
    }
    // cost of a tile in row units (a 256-row tile streams in ~1.4 us): its rows, but never less than the ring's latency
    // floor, plus a fixed part (barrier waits, descriptor, x prefetch); the first tile of a run in a strip also pays the
    // strip's x (an exposed L2 round trip) and the column reduction at the run's end
    cum.assign(seq.size() + 1, 0.0);
    for (size_t t = 0; t < seq.size(); t++) {
        const bool newrun = t == 0 || seq[t - 1].strip != seq[t].strip || seq[t - 1].cell != seq[t].cell;
        cum[t + 1] = cum[t] + std::max(seq[t].rows, 64) + 40 + (newrun ? 150 : 0);
    }
}

void symv_cut(int Np, int G, const std::vector<vlg_ba_ctx::SymvTileH>& seq, const std::vector<double>& cum,
              const std::vector<double>& speed, SymvPlan& pl)
{
    const int nstrips = Np / kSymvCols;
    std::vector<int>& tptr = pl.tptr;
    tptr.assign((size_t)G + 1, 0);
    double wsum = 0.0, wacc = 0.0;
    for (int g = 0; g < G; g++) wsum += speed[(size_t)g];
    for (int g = 1; g < G; g++) {
        wacc += speed[(size_t)g - 1];
        const double target = cum.back() * wacc / wsum;
        tptr[g] = std::max(tptr[g - 1], (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin()));
        tptr[g] = std::min(tptr[g], (int)seq.size());
    }
    tptr[G] = (int)seq.size();
    // fragments: maximal runs of one cell inside one CTA's piece; flags and fold lists
    pl.tiles.resize(seq.size());
    const int nrowblk = (Np + kSymvBlkRows - 1) / kSymvBlkRows;
    std::vector<std::vector<int>> rowl((size_t)nrowblk), coll((size_t)nstrips);
    int nfrag = 0;
    for (int g = 0; g < G; g++)
        for (int t = tptr[g]; t < tptr[g + 1]; t++) {
            const vlg_ba_ctx::SymvTileH& c = seq[(size_t)t];
            const bool ffrag = t == tptr[g] || seq[(size_t)t - 1].cell != c.cell;
            const bool lfrag = t + 1 == tptr[g + 1] || seq[(size_t)t + 1].cell != c.cell;
            const bool fstrip = ffrag || seq[(size_t)t - 1].strip != c.strip;
            const bool lstrip = lfrag || seq[(size_t)t + 1].strip != c.strip;
            if (ffrag) { nfrag++; rowl[(size_t)(c.r0 / kSymvBlkRows)].push_back(nfrag - 1); }
            if (fstrip) coll[(size_t)c.strip].push_back((nfrag - 1) * kSymvCols * kSymvSlab + kSymvCols * c.sl);
            const int flags = (fstrip ? kSymvFirstStrip : 0) | (lstrip ? kSymvLastStrip : 0) | (ffrag ? kSymvFirstFrag : 0) | (lfrag ? kSymvLastFrag : 0);
            pl.tiles[(size_t)t] = make_int4(c.strip, c.r0, c.rows | flags, (nfrag - 1) | (c.sl << 20));
        }
    pl.rptr.assign(1, 0); pl.rlist.clear(); pl.cptr.assign(1, 0); pl.clist.clear();
    for (auto& v : rowl) { pl.rlist.insert(pl.rlist.end(), v.begin(), v.end()); pl.rptr.push_back((int)pl.rlist.size()); }
    for (auto& v : coll) { pl.clist.insert(pl.clist.end(), v.begin(), v.end()); pl.cptr.push_back((int)pl.clist.size()); }
    pl.nfrag = nfrag;
}

// The context's cut: weights 1 at first; with opts.pcg_autotune re-measured over the first solves (the SMs of a B200 do not
// all stream at the same rate, +-4 % by position).  The partial results are indexed by fragment and folded in list order, so
// the product depends on the cut only through the order of summation inside a row (covered by the PCG tolerance,
// identical on every rank).
int symv_partition(vlg_ba_ctx* ctx)
{
    SymvPlan pl;
    symv_cut(ctx->Np, ctx->symv_grid, ctx->h_symv_seq, ctx->h_symv_cum, ctx->symv_speed, pl);
    ctx->h_symv_tptr = pl.tptr;
    const std::vector<int4>& tiles = pl.tiles;
    const std::vector<int>&tptr = pl.tptr, &rptr = pl.rptr, &rlist = pl.rlist, &cptr = pl.cptr, &clist = pl.clist;
    const int nfrag = pl.nfrag;
    ctx->symv_nfrag = nfrag;
    CHK(upload(ctx, ctx->symv_tiles, tiles.data(), tiles.size()));
    CHK(upload(ctx, ctx->symv_tile_ptr, tptr.data(), tptr.size()));
    CHK(upload(ctx, ctx->symv_row_ptr, rptr.data(), rptr.size())); CHK(upload(ctx, ctx->symv_row_list, rlist.data(), rlist.size()));
    CHK(upload(ctx, ctx->symv_col_ptr, cptr.data(), cptr.size())); CHK(upload(ctx, ctx->symv_col_list, clist.data(), clist.size()));
    CU(cudaStreamSynchronize(ctx->stream));      // the host vectors go out of scope
    return VLG_BA_OK;
}

// After a solve by the persistent kernel: per-CTA matvec clocks -> speed weights -> new cut (first solves of a problem only).
int symv_learn(vlg_ba_ctx* ctx)
{
    if (ctx->symv_learn_left <= 0 || !ctx->symv_stat) return VLG_BA_OK;
    const int G = ctx->symv_grid;
    std::vector<long long> st((size_t)2 * G);
    CU(cudaMemcpyAsync(st.data(), ctx->symv_stat, sizeof(long long) * st.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemsetAsync(ctx->symv_stat, 0, sizeof(long long) * st.size(), ctx->stream));
    std::vector<int> smid((size_t)G);
    for (int g = 0; g < G; g++) smid[(size_t)g] = (int)st[(size_t)2 * g + 1];
    if (!ctx->symv_smid.empty() && smid != ctx->symv_smid) {
        // the CTAs landed on other SMs than the weights were measured on: start over
        ctx->symv_speed.assign((size_t)G, 1.0);
        ctx->symv_smid = smid;
        return symv_partition(ctx);
    }
    ctx->symv_smid = smid;
    double mean = 0.0;
    std::vector<double> sp((size_t)G, 0.0);
    for (int g = 0; g < G; g++) {
        const double cost = ctx->h_symv_cum[(size_t)ctx->h_symv_tptr[(size_t)g + 1]] - ctx->h_symv_cum[(size_t)ctx->h_symv_tptr[(size_t)g]];
        const double clk = (double)st[(size_t)2 * g];
        if (!(cost > 0.0) || !(clk > 0.0)) return VLG_BA_OK;       // a CTA without work or without a measurement: keep the cut
        sp[(size_t)g] = cost / clk;
        mean += sp[(size_t)g] / G;
    }
    for (int g = 0; g < G; g++) {
        const double rel = std::min(1.25, std::max(0.8, sp[(size_t)g] / mean));
        ctx->symv_speed[(size_t)g] = rel;
    }
    ctx->symv_learn_left--;
    return symv_partition(ctx);
}

int build_problem(vlg_ba_ctx* ctx, int m, int n, const double* K, const double* a, const double* b, int64_t nobs,
                  const double* obs_xy, const int32_t* obs_pt, const int32_t* obs_cam, const double* pivot)
{
    const vlg_ba_opts& o = ctx->opt;
    if (m <= 0 || n < 0 || nobs < 0) return fail(ctx, VLG_BA_EINVAL, "bad sizes m=%d n=%d nobs=%lld", m, n, (long long)nobs);
    if ((!K && ctx->opt.model != VLG_BA_MODEL_PROJECTIVE) || !a || (n > 0 && !b)) return fail(ctx, VLG_BA_EINVAL, "K, a and b are required");
    if (nobs >= (int64_t)1 << 31) return fail(ctx, VLG_BA_EINVAL, "nobs must be < 2^31 per context");
    const bool proj = o.model == VLG_BA_MODEL_PROJECTIVE;
    const int na = proj ? kNaProjective : 6 + o.num_variableK;
    if (!proj && !(o.num_variableK == 0 || o.num_variableK == 1 || o.num_variableK == 4))
        return fail(ctx, VLG_BA_EINVAL, "num_variableK must be 0, 1 or 4");
    // validate list order (ascending i + n*j) and ranges
    for (int64_t t = 0; t < nobs; t++) {
        if (obs_pt[t] < 0 || obs_pt[t] >= n || obs_cam[t] < 0 || obs_cam[t] >= m)
            return fail(ctx, VLG_BA_EINVAL, "observation %lld out of range (pt %d, cam %d)", (long long)t, obs_pt[t], obs_cam[t]);
        if (t > 0) {
            const int64_t k0 = (int64_t)obs_pt[t - 1] + (int64_t)n * obs_cam[t - 1], k1 = (int64_t)obs_pt[t] + (int64_t)n * obs_cam[t];
            if (k1 <= k0) return fail(ctx, VLG_BA_EINVAL, "observation list must be strictly ascending in i + n*j (at %lld)", (long long)t);
        }
    }
    free_problem(ctx);
    // optional buffers: which of them exist depends on the solver path of THIS problem
    ctx->init_part = nullptr; ctx->init_bar = nullptr; ctx->persist_bar = nullptr; ctx->init_coop_cap = -1; ctx->S_zeroed = false; ctx->s_split = false;
    ctx->S = nullptr; ctx->Ybuf = nullptr; ctx->red2_local = nullptr; ctx->Cblk = nullptr; ctx->McL = nullptr; ctx->Mc2 = nullptr;
    ctx->Wp = nullptr; ctx->ptile_meta = nullptr; ctx->s1tile_meta = nullptr;
    ctx->blk_heavy = nullptr; ctx->blk_light = nullptr; ctx->nheavy = 0; ctx->nlight = 0;
    ctx->schur_chunk_ok = false; ctx->VE = nullptr; ctx->Hpart = nullptr; ctx->chunk_seg_ptr = nullptr; ctx->sblk = nullptr;
    ctx->sblk_hptr = nullptr; ctx->segs = nullptr; ctx->nsblk = 0; ctx->nsegs = 0; ctx->stile_meta = nullptr; ctx->cam_stile_ptr = nullptr; ctx->nstiles = 0;
    ctx->symv_tiles = nullptr; ctx->symv_tile_ptr = nullptr; ctx->symv_rowpart = nullptr; ctx->symv_colpart = nullptr;
    ctx->chol_R = nullptr; ctx->chol_Ld = nullptr; ctx->chol_Dinv = nullptr; ctx->chol_bar = nullptr;
    CU(cudaSetDevice(ctx->device));
    ctx->m = m; ctx->n = n; ctx->na = na; ctx->nobs = nobs;
    ctx->num_vis = (double)nobs;
    ctx->num_vis_user = false;
    const size_t N = (size_t)na * m;

    if (K) ctx->h_K.assign(K, K + 4 * (size_t)m);
    else ctx->h_K.assign(4 * (size_t)m, 0.0);
    ctx->h_a.assign(a, a + N);
    ctx->h_a_new.assign(N, 0.0);
    ctx->h_da.assign(N, 0.0);
    ctx->h_rtab.assign((size_t)36 * m, 0.0);
    ctx->h_rtab_new.assign((size_t)9 * m, 0.0);
    ctx->h_rtab_next.assign((size_t)36 * m, 0.0);
    ctx->rtab_valid = ctx->rtab_next_valid = false;
    ctx->h_obs_pt.assign(obs_pt, obs_pt + nobs);
    ctx->h_obs_cam.assign(obs_cam, obs_cam + nobs);
    ctx->h_obs_xy.assign(obs_xy, obs_xy + 2 * nobs);

    // CSR by camera and by point
    std::vector<int> cam_ptr(m + 1, 0), pt_ptr(n + 1, 0), pt_obs(nobs), pt_cam(nobs);
    for (int64_t t = 0; t < nobs; t++) { cam_ptr[obs_cam[t] + 1]++; pt_ptr[obs_pt[t] + 1]++; }
    for (int j = 0; j < m; j++) cam_ptr[j + 1] += cam_ptr[j];
    for (int i = 0; i < n; i++) pt_ptr[i + 1] += pt_ptr[i];
    {
        std::vector<int> fill(pt_ptr.begin(), pt_ptr.end() - 1);
        for (int64_t t = 0; t < nobs; t++) {
            const int q = fill[obs_pt[t]]++;
            pt_obs[q] = (int)t; pt_cam[q] = obs_cam[t];
        }
    }
    // chunks of the camera segments
    int cs;
    if (o.order == VLG_BA_ORDER_REFERENCE) cs = 1 << 30;
    else {
        int64_t want = nobs / (148 * 8);
        cs = (int)std::min<int64_t>(kCamTile, std::max<int64_t>(32, (want + 31) / 32 * 32));
    }
    ctx->chunk_size = cs;
    std::vector<int> chunk_cam, chunk_begin, chunk_end, cam_chunk_ptr(m + 1, 0);
    for (int j = 0; j < m; j++) {
        for (int64_t s = cam_ptr[j]; s < cam_ptr[j + 1]; s += cs) {
            chunk_cam.push_back(j);
            chunk_begin.push_back((int)s);
            chunk_end.push_back((int)std::min<int64_t>(s + cs, cam_ptr[j + 1]));
        }
        cam_chunk_ptr[j + 1] = (int)chunk_cam.size();
    }
    ctx->nchunks = (int)chunk_cam.size();
    {
        const char* e = getenv("VLG_BA_SCHUR_CHUNK");      // 0: the round-1 kernels (camera pass + pair lists for every block)
        ctx->schur_chunk_ok = na == 6 && o.order != VLG_BA_ORDER_REFERENCE && (e ? atoi(e) != 0 : true);
    }
    // ... and that kernel's tiles
    std::vector<int2> stile_meta;
    std::vector<int> cam_stile_ptr((size_t)m + 1, 0);
    if (ctx->schur_chunk_ok) {
        for (int j = 0; j < m; j++) {
            for (int64_t sb = cam_ptr[j]; sb < cam_ptr[j + 1]; sb += kSchurTile)
                stile_meta.push_back(make_int2((int)sb, (int)std::min<int64_t>(kSchurTile, cam_ptr[j + 1] - sb)));
            cam_stile_ptr[(size_t)j + 1] = (int)stile_meta.size();
        }
        ctx->nstiles = (int)stile_meta.size();
    }

    std::vector<unsigned char> fixed(m, 0);
    for (int j = 0; j < m; j++) fixed[j] = (o.fix_motion || (pivot && pivot[j] != 0.0)) ? 1 : 0;

    ctx->use_chol = (o.solver == VLG_BA_SOLVER_CHOL) || (o.solver == VLG_BA_SOLVER_AUTO && m <= o.chol_max_cams);
    ctx->Np = (int)((N + kNB - 1) / kNB * kNB);
    {
        // explicit-S PCG: one iteration streams the lower triangle (4 Np^2 bytes) instead of W twice
        // (2 x (24 na + 8) bytes per observation); assembling S costs about a dozen sweeps.  Every rank multiplies 1/nranks
        // of S and sweeps its own shard of W, so the rule compares the WHOLE problem's figures -- and it must: the choice of
        // solver is collective (ranks on different paths would wait for each other's exchanges forever), so it may only
        // depend on quantities that are identical on every rank
        double nobs_all = (double)nobs;
        if (ctx->nranks > 1) {
            double* d_cnt = nullptr;
            CU(cudaMalloc(&d_cnt, sizeof(double)));
            int rc = upload(ctx, d_cnt, &nobs_all, 1);
            if (rc == VLG_BA_OK) rc = allreduce(ctx, d_cnt, 1);
            if (rc == VLG_BA_OK) rc = download(ctx, &nobs_all, d_cnt, 1);
            cudaStreamSynchronize(ctx->stream);
            cudaFree(d_cnt);
            CHK(rc);
        }
        const double s_bytes = 8.0 * (double)ctx->Np * (double)ctx->Np;
        const bool pays = 0.75 * s_bytes < 2.0 * (24.0 * na + 8.0) * nobs_all && s_bytes <= 8e9 && (int64_t)m * m <= ((int64_t)1 << 28);
        ctx->use_explicit = !ctx->use_chol && (o.solver == VLG_BA_SOLVER_PCG_EXPLICIT || (o.solver == VLG_BA_SOLVER_AUTO && pays));
    }
    const bool need_S = ctx->use_chol || ctx->use_explicit;
    // cluster-Jacobi preconditioner (PCG paths): needs the cooperative update kernel, one CTA per cluster
    bool cluster_pc = false;
    {
        int coop = 0, nsm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
        cluster_pc = !ctx->use_chol && o.pcg_cluster && coop && cdiv(m, 128 / na) <= nsm * 8;
    }
    const bool cluster_blocks = cluster_pc && !ctx->use_explicit;     // implicit path: assemble the within-cluster blocks only

    // block structure of S: cameras j <= k sharing a point; pair lists in ascending point order
    std::vector<int64_t> blk_ptr;
    std::vector<int2> pairs;
    ctx->h_blk_j.clear(); ctx->h_blk_k.clear();
    if (need_S) {
        if ((int64_t)m * m > ((int64_t)1 << 28)) return fail(ctx, VLG_BA_EINVAL, "explicit Schur/Cholesky path supports m <= 16384 cameras");
        std::vector<int64_t> cnt((size_t)m * m + 1, 0);
        for (int i = 0; i < n; i++)
            for (int qa = pt_ptr[i]; qa < pt_ptr[i + 1]; qa++)
                for (int qb = qa; qb < pt_ptr[i + 1]; qb++) cnt[(size_t)pt_cam[qa] * m + pt_cam[qb] + 1]++;
        // diagonal blocks always exist (S_jj = U*_j even without observations)
        std::vector<int64_t> slot((size_t)m * m, -1);
        blk_ptr.push_back(0);
        for (int j = 0; j < m; j++)
            for (int k = j; k < m; k++) {
                const int64_t c = cnt[(size_t)j * m + k + 1];
                if (c > 0 || j == k) {
                    slot[(size_t)j * m + k] = (int64_t)ctx->h_blk_j.size();
                    ctx->h_blk_j.push_back(j); ctx->h_blk_k.push_back(k);
                    blk_ptr.push_back(blk_ptr.back() + c);
                }
            }
        pairs.resize((size_t)blk_ptr.back());
        std::vector<int64_t> fill(blk_ptr.begin(), blk_ptr.end() - 1);
        for (int i = 0; i < n; i++)
            for (int qa = pt_ptr[i]; qa < pt_ptr[i + 1]; qa++)
                for (int qb = qa; qb < pt_ptr[i + 1]; qb++) {
                    const int64_t s = slot[(size_t)pt_cam[qa] * m + pt_cam[qb]];
                    pairs[(size_t)fill[s]++] = make_int2(pt_obs[qa], pt_obs[qb]);
                }
        ctx->nblocks = (int64_t)ctx->h_blk_j.size();
        ctx->npairs = (int64_t)pairs.size();
    } else if (cluster_blocks) {
        // implicit-Schur PCG with the cluster-Jacobi preconditioner: only the blocks (j,k) whose cameras sit in
        // the same cluster of kc consecutive cameras are ever assembled; same pair-list layout, compact keys
        const int kc = 128 / na, gcl = (m + kc - 1) / kc;
        auto key = [&](int j, int k) { return ((size_t)(j / kc) * kc + (size_t)(j % kc)) * kc + (size_t)(k % kc); };
        std::vector<int64_t> cnt((size_t)gcl * kc * kc + 1, 0);
        for (int i = 0; i < n; i++)
            for (int qa = pt_ptr[i]; qa < pt_ptr[i + 1]; qa++)
                for (int qb = qa + 1; qb < pt_ptr[i + 1]; qb++)
                    if (pt_cam[qa] / kc == pt_cam[qb] / kc) cnt[key(pt_cam[qa], pt_cam[qb]) + 1]++;
        std::vector<int64_t> slot((size_t)gcl * kc * kc, -1);
        blk_ptr.push_back(0);
        for (int j = 0; j < m; j++)
            for (int k = j + 1; k < std::min(m, (j / kc + 1) * kc); k++) {
                const int64_t c = cnt[key(j, k) + 1];
                if (c > 0) {
                    slot[key(j, k)] = (int64_t)ctx->h_blk_j.size();
                    ctx->h_blk_j.push_back(j); ctx->h_blk_k.push_back(k);
                    blk_ptr.push_back(blk_ptr.back() + c);
                }
            }
        pairs.resize((size_t)blk_ptr.back());
        std::vector<int64_t> fill(blk_ptr.begin(), blk_ptr.end() - 1);
        for (int i = 0; i < n; i++)
            for (int qa = pt_ptr[i]; qa < pt_ptr[i + 1]; qa++)
                for (int qb = qa + 1; qb < pt_ptr[i + 1]; qb++)
                    if (pt_cam[qa] / kc == pt_cam[qb] / kc) {
                        const int64_t s = slot[key(pt_cam[qa], pt_cam[qb])];
                        pairs[(size_t)fill[s]++] = make_int2(pt_obs[qa], pt_obs[qb]);
                    }
        ctx->nblocks = (int64_t)ctx->h_blk_j.size();
        ctx->npairs = (int64_t)pairs.size();
    } else {
        ctx->nblocks = 0; ctx->npairs = 0;
    }

    // point tiles for the PCG point sweep: consecutive whole points, <= pt_tile observations
    std::vector<int> ptile_first;
    ctx->tiled_ok = !need_S;
    {
        const char* e = getenv("VLG_BA_RING");
        // measured on B200 (Venice shape, gpurun_out/bench_venice_r01h_*): the ring helps the camera
        // sweep (0.1485 vs 0.1526 ms) and hurts the point sweep (0.170 vs 0.148 ms: 512-observation
        // tiles do not fit a 2-stage ring at 3 CTAs/SM), so only the camera ring is on by default
        ctx->use_ring = (3 * na) % 2 == 0 ? (e ? atoi(e) : 1) : 0;
        ctx->pt_tile = (ctx->use_ring & 2) ? kRingTile : kPtTile;
        cudaDeviceGetAttribute(&ctx->nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    }
    if (!need_S) {
        int64_t acc = 0;
        int cntp = 0;
        ptile_first.push_back(0);
        for (int i = 0; i < n; i++) {
            const int t = pt_ptr[i + 1] - pt_ptr[i];
            if (t > ctx->pt_tile) { ctx->tiled_ok = false; break; }
            if (acc + t > ctx->pt_tile || cntp == ctx->pt_tile) { ptile_first.push_back(i); acc = 0; cntp = 0; }
            acc += t; cntp++;
        }
        ptile_first.push_back(n);
        if (n == 0) ctx->tiled_ok = false;
    }
    ctx->nptiles = ctx->tiled_ok ? (int)ptile_first.size() - 1 : 0;
    // tiles of the stage-1 point pass (any solver): whole points, <= kS1Tile observations; a track longer
    // than a tile sends the pass back to the thread-per-point kernel
    std::vector<int4> s1tiles;
    {
        bool ok = n > 0;
        int p_first = 0;
        int64_t acc = 0;
        for (int i = 0; i < n && ok; i++) {
            const int t = pt_ptr[i + 1] - pt_ptr[i];
            if (t > kS1Tile) { ok = false; break; }
            if (acc + t > kS1Tile || i - p_first == kS1Tile) {
                s1tiles.push_back(make_int4(pt_ptr[p_first], pt_ptr[i] - pt_ptr[p_first], p_first, i - p_first));
                p_first = i; acc = 0;
            }
            acc += t;
        }
        if (ok) s1tiles.push_back(make_int4(pt_ptr[p_first], pt_ptr[n] - pt_ptr[p_first], p_first, n - p_first));
        else s1tiles.clear();
    }
    ctx->ns1tiles = (int)s1tiles.size();

    const int NU = nu_of(na);
    // ---- device allocations
    CHK(dalloc(ctx, &ctx->obs_xy, (size_t)nobs));
    CHK(dalloc(ctx, &ctx->obs_pt, (size_t)nobs)); CHK(dalloc(ctx, &ctx->obs_cam, (size_t)nobs));
    CHK(dalloc(ctx, &ctx->cam_ptr, (size_t)m + 1)); CHK(dalloc(ctx, &ctx->pt_ptr, (size_t)n + 1));
    CHK(dalloc(ctx, &ctx->pt_obs, (size_t)nobs)); CHK(dalloc(ctx, &ctx->pt_cam, (size_t)nobs));
    CHK(dalloc(ctx, &ctx->chunk_cam, (size_t)ctx->nchunks)); CHK(dalloc(ctx, &ctx->chunk_begin, (size_t)ctx->nchunks));
    CHK(dalloc(ctx, &ctx->chunk_end, (size_t)ctx->nchunks)); CHK(dalloc(ctx, &ctx->cam_chunk_ptr, (size_t)m + 1));
    CHK(dalloc(ctx, &ctx->cam_fixed, (size_t)m));
    CHK(dalloc(ctx, &ctx->K4, (size_t)4 * m)); CHK(dalloc(ctx, &ctx->a, N)); CHK(dalloc(ctx, &ctx->a_new, N));
    CHK(dalloc(ctx, &ctx->b, (size_t)3 * n)); CHK(dalloc(ctx, &ctx->b_new, (size_t)3 * n));
    CHK(dalloc(ctx, &ctx->rtab, (size_t)36 * m)); CHK(dalloc(ctx, &ctx->rtab_new, (size_t)9 * m)); CHK(dalloc(ctx, &ctx->rtab_next, (size_t)36 * m));
    CHK(dalloc(ctx, &ctx->red1, (size_t)na * N + N + 2));
    ctx->U = ctx->red1; ctx->eA = ctx->red1 + (size_t)na * N; ctx->scal1 = ctx->eA + N;
    CHK(dalloc(ctx, &ctx->W, (size_t)3 * na * nobs)); CHK(dalloc(ctx, &ctx->Upart, (size_t)NU * ctx->nchunks));
    CHK(dalloc(ctx, &ctx->BeP, (size_t)8 * nobs)); CHK(dalloc(ctx, &ctx->obs_slot, (size_t)nobs));
    CHK(dalloc(ctx, &ctx->V, (size_t)9 * n)); CHK(dalloc(ctx, &ctx->eB, (size_t)3 * n));
    CHK(dalloc(ctx, &ctx->cost_pt, (size_t)n)); CHK(dalloc(ctx, &ctx->red_part, (size_t)kRedBlocks));
    CHK(dalloc(ctx, &ctx->Ud, (size_t)na * N)); CHK(dalloc(ctx, &ctx->Vinv, (size_t)9 * n));
    if (ctx->schur_chunk_ok) {
        CHK(dalloc(ctx, &ctx->VE, (size_t)kVE * n));
        CHK(dalloc(ctx, &ctx->stile_meta, stile_meta.size())); CHK(dalloc(ctx, &ctx->cam_stile_ptr, cam_stile_ptr.size()));
        CHK(upload(ctx, ctx->stile_meta, stile_meta.data(), stile_meta.size()));
        CHK(upload(ctx, ctx->cam_stile_ptr, cam_stile_ptr.data(), cam_stile_ptr.size()));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    CHK(dalloc(ctx, &ctx->Spart, (size_t)NU * std::max(ctx->nchunks, ctx->nstiles)));
    CHK(dalloc(ctx, &ctx->red2, (size_t)NU * m));
    CHK(dalloc(ctx, &ctx->Sjj, (size_t)na * N)); CHK(dalloc(ctx, &ctx->ebar, N));
    CHK(dalloc(ctx, &ctx->Minv, (size_t)na * N)); CHK(dalloc(ctx, &ctx->da, N));
    CHK(dalloc(ctx, &ctx->pr, N)); CHK(dalloc(ctx, &ctx->pz, N)); CHK(dalloc(ctx, &ctx->pp, N)); CHK(dalloc(ctx, &ctx->pp2, N)); CHK(dalloc(ctx, &ctx->pq, N));
    CHK(dalloc(ctx, &ctx->wq, N)); CHK(dalloc(ctx, &ctx->tvec, (size_t)4 * n)); CHK(dalloc(ctx, &ctx->qpart, (size_t)na * ctx->nchunks));
    CHK(dalloc(ctx, &ctx->pcg_sc, 1));
    CHK(dalloc(ctx, &ctx->db, (size_t)3 * n)); CHK(dalloc(ctx, &ctx->denom_pt, (size_t)n));
    CHK(dalloc(ctx, &ctx->cost_obs, (size_t)nobs)); CHK(dalloc(ctx, &ctx->scal3, 4));
    CU(cudaMemsetAsync(ctx->scal3, 0, 4 * sizeof(double), ctx->stream));      // scal3[3] doubles as the operand of barrier all-reduces
    ctx->coop_grid = 0;
    if (!ctx->use_chol) {
        int coop = 0, nsm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
        const int g = cdiv(m, 128 / na);            // k_pcg_update_coop: a CTA owns 128/na whole cameras
        if (coop && g <= nsm * 8) ctx->coop_grid = g;
        CHK(dalloc(ctx, &ctx->blkpart, (size_t)11 * std::max(g, 1)));
        if (ctx->coop_grid > 0) { CHK(dalloc(ctx, &ctx->init_part, (size_t)17 * g)); CHK(dalloc(ctx, &ctx->init_bar, 1)); }
        if (cluster_pc && ctx->coop_grid > 0) {
            // Cblk holds the g blocks of the cluster partition and, behind them, the g + 1 blocks of the shifted one
            const bool ovl = ctx->use_explicit;
            CHK(dalloc(ctx, &ctx->Cblk, (size_t)(ovl ? 2 * g + 1 : g) * 128 * 128)); CHK(dalloc(ctx, &ctx->McL, (size_t)g * 128 * 128));
            if (ovl) {
                CHK(dalloc(ctx, &ctx->Mc2, (size_t)g * 128 * 128));
                CU(cudaMemsetAsync(ctx->Mc2, 0, sizeof(double) * (size_t)g * 128 * 128, ctx->stream));     // padding rows are staged, never used
                const char* e = getenv("VLG_BA_OVERLAP");
                ctx->overlap_ok = e ? atoi(e) != 0 : true;
            }
        }
        CHK(dalloc(ctx, &ctx->Zd, (size_t)kDefl * N)); CHK(dalloc(ctx, &ctx->SZd, (size_t)kDefl * N)); CHK(dalloc(ctx, &ctx->defl_sc, 1));
        if (ctx->tiled_ok) {
            CHK(dalloc(ctx, &ctx->Wp, (size_t)3 * na * nobs));
            std::vector<int4> pm((size_t)ctx->nptiles);
            for (int t = 0; t < ctx->nptiles; t++) {
                const int p0 = ptile_first[t], p1 = ptile_first[t + 1];
                pm[t] = make_int4(pt_ptr[p0], pt_ptr[p1] - pt_ptr[p0], p0, p1 - p0);
            }
            CHK(dalloc(ctx, &ctx->ptile_meta, pm.size()));
            CHK(upload(ctx, ctx->ptile_meta, pm.data(), pm.size()));
            CU(cudaStreamSynchronize(ctx->stream));
        }
    }
    if (ctx->use_explicit) {
        // tiles of the lower triangle in cell order (cells = kSymvSlab strips x kSymvBlkRows rows, strip by strip inside a
        // cell), cut into one contiguous, equally expensive piece per persistent CTA; see ba_pcg.cuh (SymvSmem)
        const int Np = ctx->Np, nstrips = Np / kSymvCols;
        // multi-GPU (set_comm before set_problem): every rank assembles its share of S in full, rank r then holds the SUM
        // of its column block (strips [J_r, J_r+1), chosen so that every rank gets the same lower-triangle area), and the
        // matvec of rank r only walks those strips; the partial products meet in the same vector exchange as before
        ctx->s_split = ctx->nranks > 1 && g_nccl.Reduce && g_nccl.GroupStart && g_nccl.GroupEnd;
        ctx->s_bounds.assign((size_t)ctx->nranks + 1, nstrips);
        ctx->s_bounds[0] = 0;
        if (ctx->s_split) {
            // strip J of the lower triangle has nstrips - J tiles: boundaries of equal area
            for (int r = 1; r < ctx->nranks; r++)
                ctx->s_bounds[(size_t)r] = std::min(nstrips, std::max(ctx->s_bounds[(size_t)r - 1],
                                                    (int)lround(nstrips * (1.0 - sqrt(1.0 - (double)r / ctx->nranks)))));
        }
        ctx->s_J0 = ctx->s_split ? ctx->s_bounds[(size_t)ctx->rank] : 0;
        ctx->s_J1 = ctx->s_split ? ctx->s_bounds[(size_t)ctx->rank + 1] : nstrips;
        std::vector<vlg_ba_ctx::SymvTileH>& seq = ctx->h_symv_seq;
        int ncell = 0;
        // which (256-row slot, strip) pairs of the lower triangle hold a non-zero block: block (j, k), j <= k, puts the rows
        // of camera k against the columns of camera j; a strip's own diagonal tile is always kept (U* sits there)
        std::vector<unsigned char> occ((size_t)((Np + kSymvRows - 1) / kSymvRows) * nstrips, 0);
        {
            const char* e = getenv("VLG_BA_SYMV_DENSE");         // 1: stream every tile of the triangle (round-1 behaviour)
            if (e && atoi(e) != 0) std::fill(occ.begin(), occ.end(), 1);
            for (size_t bb = 0; bb < ctx->h_blk_j.size(); bb++) {
                const int j = ctx->h_blk_j[bb], k = ctx->h_blk_k[bb];
                for (int rs = (na * k) / kSymvRows; rs <= (na * k + na - 1) / kSymvRows; rs++)
                    for (int J = (na * j) / kSymvCols; J <= (na * j + na - 1) / kSymvCols; J++) occ[(size_t)rs * nstrips + J] = 1;
            }
            for (int J = 0; J < nstrips; J++) occ[(size_t)((kSymvCols * J) / kSymvRows) * nstrips + J] = 1;
        }
        if (ctx->nranks > 1) {
            // a rank multiplies its column block of the SUM of all ranks' shares: the pattern is the union over ranks
            std::vector<double> od(occ.begin(), occ.end());
            double* d_occ = nullptr;
            CU(cudaMalloc(&d_occ, sizeof(double) * od.size()));
            int rc = upload(ctx, d_occ, od.data(), od.size());
            if (rc == VLG_BA_OK) rc = allreduce(ctx, d_occ, od.size());
            if (rc == VLG_BA_OK) rc = download(ctx, od.data(), d_occ, od.size());
            cudaStreamSynchronize(ctx->stream);
            cudaFree(d_occ);
            CHK(rc);
            for (size_t t = 0; t < occ.size(); t++) occ[t] = od[t] > 0.0 ? 1 : 0;
        }
        symv_sequence(Np, ctx->s_J0, ctx->s_J1, seq, ctx->h_symv_cum, ncell, occ.data());
        ctx->symv_bytes = 0;
        for (const auto& t : seq) ctx->symv_bytes += (int64_t)t.rows * kSymvCols * 8;
        // one CTA per SM at most; never fewer CTAs than preconditioner clusters (when those fit one per SM), even if this
        // rank's column block has fewer tiles than that (surplus CTAs get an empty piece): whether the whole solve runs as the
        // persistent kernel (coop_grid <= symv_grid, run_stage2) must come out the same on every rank
        const int G = std::max(1, std::min(ctx->nsm, std::max((int)seq.size(), ctx->coop_grid)));
        ctx->symv_grid = G;
        ctx->symv_ncell = ncell;
        ctx->symv_speed.assign((size_t)G, 1.0);
        ctx->symv_smid.clear();
        {
            const char* e = getenv("VLG_BA_SYMV_LEARN");     // overrides opts.pcg_autotune
            ctx->symv_learn_left = e ? atoi(e) : o.pcg_autotune;
        }
        const size_t nfrag_max = (size_t)ncell + G, nrun_max = (size_t)ncell * kSymvSlab + G;
        CHK(dalloc(ctx, &ctx->symv_tiles, seq.size())); CHK(dalloc(ctx, &ctx->symv_tile_ptr, (size_t)G + 1));
        CHK(dalloc(ctx, &ctx->symv_row_ptr, (size_t)(Np + kSymvBlkRows - 1) / kSymvBlkRows + 1)); CHK(dalloc(ctx, &ctx->symv_row_list, nfrag_max));
        CHK(dalloc(ctx, &ctx->symv_col_ptr, (size_t)nstrips + 1)); CHK(dalloc(ctx, &ctx->symv_col_list, nrun_max));
        CHK(dalloc(ctx, &ctx->symv_rowpart, nfrag_max * kSymvBlkRows));
        CHK(dalloc(ctx, &ctx->symv_colpart, nfrag_max * kSymvCols * kSymvSlab));
        CHK(dalloc(ctx, &ctx->symv_stat, (size_t)2 * G));
        CU(cudaMemsetAsync(ctx->symv_stat, 0, sizeof(long long) * 2 * G, ctx->stream));
        CHK(symv_partition(ctx));
        CHK(dalloc(ctx, &ctx->persist_bar, 1));
        {
            const char* e = getenv("VLG_BA_PERSIST");      // 0: one launch per PCG phase instead of the persistent kernel
            ctx->persist_ok = e ? atoi(e) != 0 : true;
        }

    }
    if (ctx->use_chol) {
        const int nb = ctx->Np / kNB;
        CHK(dalloc(ctx, &ctx->chol_R, (size_t)nb * kNB * kNB)); CHK(dalloc(ctx, &ctx->chol_Ld, (size_t)nb * kNB * kNB));
        CHK(dalloc(ctx, &ctx->chol_Dinv, (size_t)ctx->Np)); CHK(dalloc(ctx, &ctx->chol_bar, 1));
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        if (!coop) return fail(ctx, VLG_BA_ECUDA, "device lacks cooperative launch (needed by k_chol_coop)");
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chol_coop, kCholWarps * 32, 0));
        if (per_sm < 1) return fail(ctx, VLG_BA_ECUDA, "k_chol_coop does not fit an SM");
        // enough CTAs for the first trailing update (one 32 x 32 tile per warp), never more than one per SM
        const int64_t tiles = (int64_t)nb * (nb + 1) / 2;
        ctx->chol_grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->nsm, (tiles + kCholWarps - 1) / kCholWarps));
    }
    if (need_S) {
        CHK(dalloc(ctx, &ctx->S, (size_t)ctx->Np * ctx->Np));
    }
    if (need_S || cluster_blocks) {
        CHK(dalloc(ctx, &ctx->blk_j, (size_t)ctx->nblocks)); CHK(dalloc(ctx, &ctx->blk_k, (size_t)ctx->nblocks));
        CHK(dalloc(ctx, &ctx->blk_ptr, (size_t)ctx->nblocks + 1)); CHK(dalloc(ctx, &ctx->pairs, (size_t)ctx->npairs));
        CHK(upload(ctx, ctx->blk_j, ctx->h_blk_j.data(), (size_t)ctx->nblocks));
        CHK(upload(ctx, ctx->blk_k, ctx->h_blk_k.data(), (size_t)ctx->nblocks));
        CHK(upload(ctx, ctx->blk_ptr, blk_ptr.data(), (size_t)ctx->nblocks + 1));
        CHK(upload(ctx, ctx->pairs, pairs.data(), (size_t)ctx->npairs));
        {
            std::vector<int> heavy, light, sblk, sblk_hptr(1, 0);
            struct SegH { int chunk, blk, q0, q1, slot; };
            std::vector<SegH> segh;
            // diagonal blocks are not on these lists: the camera pass already sums them (k_schur_diag_fill).  Blocks with very
            // many pairs go to the chunk kernel (ba_schur.cuh): cut into (block, chunk of camera j) segments -- a camera's
            // observations ascend in the point index, and so do a block's pairs, hence a segment is a contiguous pair range
            int64_t seg_heavy = kSegHeavy;
            { const char* e = getenv("VLG_BA_SEG_HEAVY"); if (e) seg_heavy = atoll(e); }       // experiments: threshold of the chunk kernel's blocks
            for (int64_t bb = 0; bb < ctx->nblocks; bb++) {
                const int j = ctx->h_blk_j[(size_t)bb];
                if (j == ctx->h_blk_k[(size_t)bb]) continue;
                const int64_t np_b = blk_ptr[bb + 1] - blk_ptr[bb];
                if (ctx->schur_chunk_ok && np_b > seg_heavy && pairs.size() < ((size_t)1 << 31)) {
                    int cur = -1;
                    for (int64_t q = blk_ptr[bb]; q < blk_ptr[bb + 1]; q++) {
                        const int ch = cam_stile_ptr[(size_t)j] + (pairs[(size_t)q].x - cam_ptr[j]) / kSchurTile;
                        if (ch != cur) {
                            if (cur >= 0) segh.back().q1 = (int)q;
                            segh.push_back({ch, (int)bb, (int)q, (int)q, (int)segh.size()});
                            cur = ch;
                        }
                    }
                    segh.back().q1 = (int)blk_ptr[bb + 1];
                    sblk.push_back((int)bb);
                    sblk_hptr.push_back((int)segh.size());
                } else {
                    (np_b > kSchurHeavy ? heavy : light).push_back((int)bb);
                }
            }
            if (ctx->schur_chunk_ok) {
                // segments grouped by chunk (counting sort; inside a chunk in block order), cut into ROUNDS of <= 32 pairs; warp
                // w of the chunk's CTA takes the rounds [nr w / 8, nr (w + 1) / 8) of the chunk, so a segment that straddles
                // two warps leaves one partial per warp ("piece").  Partial slots are numbered block-major, then chunk, then
                // piece: k_schur_fold adds a block's slots in that order.
                const int nseg = (int)segh.size();
                std::vector<int> cptr((size_t)ctx->nstiles + 1, 0), crounds((size_t)ctx->nstiles + 1, 0);
                for (const SegH& g : segh) { cptr[(size_t)g.chunk + 1]++; crounds[(size_t)g.chunk + 1] += (g.q1 - g.q0 + 31) / 32; }
                for (int c = 0; c < ctx->nstiles; c++) { cptr[(size_t)c + 1] += cptr[(size_t)c]; crounds[(size_t)c + 1] += crounds[(size_t)c]; }
                std::vector<int> fill(cptr.begin(), cptr.end() - 1), order((size_t)nseg);
                for (int g = 0; g < nseg; g++) order[(size_t)fill[(size_t)segh[(size_t)g].chunk]++] = g;
                // a chunk's segments are dealt to the 8 warps of its CTA, longest first to the least loaded warp; a warp's
                // rounds are contiguous in `rounds`, a segment never straddles two warps (one partial per segment)
                std::vector<int> slot_base((size_t)nseg + 1, 0);
                for (int g = 0; g < nseg; g++) slot_base[(size_t)g + 1] = slot_base[(size_t)g] + 1;
                for (size_t t = 0; t < sblk_hptr.size(); t++) sblk_hptr[t] = slot_base[(size_t)sblk_hptr[t]];      // segment index -> slot index
                std::vector<int4> rounds((size_t)crounds.back());
                std::vector<int> wptr((size_t)kSchurWarps * ctx->nstiles + 1, 0);
                {
                    std::vector<int> segs_c, load(kSchurWarps);
                    std::vector<std::vector<int>> mine(kSchurWarps);
                    int pos = 0;
                    for (int c = 0; c < ctx->nstiles; c++) {
                        segs_c.assign(order.begin() + cptr[(size_t)c], order.begin() + cptr[(size_t)c + 1]);
                        auto nrounds = [&](int g) { return (segh[(size_t)g].q1 - segh[(size_t)g].q0 + 31) / 32; };
                        std::stable_sort(segs_c.begin(), segs_c.end(), [&](int x, int y) { return nrounds(x) > nrounds(y); });
                        std::fill(load.begin(), load.end(), 0);
                        for (auto& v : mine) v.clear();
                        for (int g : segs_c) {
                            const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
                            load[(size_t)w] += nrounds(g) + 1;        // + the flush
                            mine[(size_t)w].push_back(g);
                        }
                        for (int w = 0; w < kSchurWarps; w++) {
                            wptr[(size_t)kSchurWarps * c + w] = pos;
                            for (int g : mine[(size_t)w]) {
                                const SegH& sg = segh[(size_t)g];
                                const int n = nrounds(g);
                                for (int t = 0; t < n; t++)
                                    rounds[(size_t)pos++] = make_int4(sg.q0 + 32 * t, std::min(32, sg.q1 - (sg.q0 + 32 * t)), slot_base[(size_t)g], t + 1 == n ? 1 : 0);
                            }
                        }
                    }
                    wptr.back() = pos;
                }
                const std::vector<int>& crounds_up = wptr;
                ctx->nsblk = (int)sblk.size(); ctx->nsegs = (int)rounds.size();
                CHK(dalloc(ctx, &ctx->chunk_seg_ptr, crounds_up.size())); CHK(dalloc(ctx, &ctx->segs, rounds.size()));
                CHK(dalloc(ctx, &ctx->sblk, sblk.size())); CHK(dalloc(ctx, &ctx->sblk_hptr, sblk_hptr.size()));
                CHK(dalloc(ctx, &ctx->Hpart, (size_t)slot_base.back() * (size_t)na * na));
                CHK(upload(ctx, ctx->chunk_seg_ptr, crounds_up.data(), crounds_up.size())); CHK(upload(ctx, ctx->segs, rounds.data(), rounds.size()));
                CHK(upload(ctx, ctx->sblk, sblk.data(), sblk.size())); CHK(upload(ctx, ctx->sblk_hptr, sblk_hptr.data(), sblk_hptr.size()));
                CU(cudaStreamSynchronize(ctx->stream));
            }
            ctx->nheavy = (int)heavy.size(); ctx->nlight = (int)light.size();
            CHK(dalloc(ctx, &ctx->blk_heavy, heavy.size())); CHK(dalloc(ctx, &ctx->blk_light, light.size()));
            CHK(upload(ctx, ctx->blk_heavy, heavy.data(), heavy.size())); CHK(upload(ctx, ctx->blk_light, light.data(), light.size()));
            CU(cudaStreamSynchronize(ctx->stream));
            CHK(dalloc(ctx, &ctx->Ybuf, (size_t)3 * na * nobs));
            CHK(dalloc(ctx, &ctx->red2_local, (size_t)nu_of(na) * m));
        }
    }
    CHK(upload(ctx, (double*)ctx->obs_xy, obs_xy, 2 * (size_t)nobs));
    CHK(upload(ctx, ctx->obs_pt, obs_pt, (size_t)nobs)); CHK(upload(ctx, ctx->obs_cam, obs_cam, (size_t)nobs));
    CHK(upload(ctx, ctx->cam_ptr, cam_ptr.data(), (size_t)m + 1)); CHK(upload(ctx, ctx->pt_ptr, pt_ptr.data(), (size_t)n + 1));
    CHK(upload(ctx, ctx->pt_obs, pt_obs.data(), (size_t)nobs)); CHK(upload(ctx, ctx->pt_cam, pt_cam.data(), (size_t)nobs));
    {
        std::vector<int> slot((size_t)nobs);
        for (int64_t q = 0; q < nobs; q++) slot[(size_t)pt_obs[(size_t)q]] = (int)q;
        CHK(upload(ctx, ctx->obs_slot, slot.data(), (size_t)nobs));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->ns1tiles > 0) {
        CHK(dalloc(ctx, &ctx->s1tile_meta, s1tiles.size()));
        CHK(upload(ctx, ctx->s1tile_meta, s1tiles.data(), s1tiles.size()));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    CHK(upload(ctx, ctx->chunk_cam, chunk_cam.data(), (size_t)ctx->nchunks));
    CHK(upload(ctx, ctx->chunk_begin, chunk_begin.data(), (size_t)ctx->nchunks));
    CHK(upload(ctx, ctx->chunk_end, chunk_end.data(), (size_t)ctx->nchunks));
    {
        std::vector<int2> cm((size_t)ctx->nchunks);
        for (int c = 0; c < ctx->nchunks; c++) cm[c] = make_int2(chunk_begin[c], chunk_end[c] - chunk_begin[c]);
        CHK(dalloc(ctx, &ctx->chunk_meta, cm.size()));
        CHK(upload(ctx, ctx->chunk_meta, cm.data(), cm.size()));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    CHK(upload(ctx, ctx->cam_chunk_ptr, cam_chunk_ptr.data(), (size_t)m + 1));
    CHK(upload(ctx, ctx->cam_fixed, fixed.data(), (size_t)m));
    CHK(upload(ctx, ctx->K4, ctx->h_K.data(), (size_t)4 * m)); CHK(upload(ctx, ctx->a, a, N)); CHK(upload(ctx, ctx->b, b, (size_t)3 * n));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_problem = true;
    ctx->lambda = o.lambda0; ctx->nu = o.nu0; ctx->iter = 1; ctx->iter2 = 0;
    ctx->err_hist.clear();
    ctx->s1_valid = ctx->s2_valid = ctx->s3_valid = false;
    return VLG_BA_OK;
}

// ---- stage launches, templated on NA -----------------------------------------------------------
template <int NA>
int run_rtab(vlg_ba_ctx* ctx, const std::vector<double>& h_a, const double* d_a, int nmat, std::vector<double>& h_tab, double* d_tab)
{
    if (NA == kNaProjective) return VLG_BA_OK;      // no rotation in the projective model
    if (ctx->opt.rtable == VLG_BA_RTABLE_HOST_LIBM) {
        rtab_host(ctx->m, NA, h_a.data(), nmat, h_tab.data());
        CHK(upload(ctx, d_tab, h_tab.data(), (size_t)9 * nmat * ctx->m));
    } else {
        k_rtab<NA><<<cdiv((int64_t)ctx->m * nmat, 128), 128, 0, ctx->stream>>>(ctx->m, d_a, nmat, d_tab);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    return VLG_BA_OK;
}

template <int NA>
int run_stage1(vlg_ba_ctx* ctx, double* diag_X, double* diag_A, double* diag_B, double* diag_e)
{
    constexpr int NU = nu_of(NA);
    const bool diag = diag_X || diag_A || diag_B || diag_e;
    if (!ctx->rtab_valid) CHK(run_rtab<NA>(ctx, ctx->h_a, ctx->a, 4, ctx->h_rtab, ctx->rtab));
    ctx->rtab_valid = true;
    Stage1Args p;
    p.m = ctx->m; p.n = ctx->n; p.nchunks = ctx->nchunks;
    p.obs_xy = ctx->obs_xy; p.obs_pt = ctx->obs_pt;
    p.chunk_cam = ctx->chunk_cam; p.chunk_begin = ctx->chunk_begin; p.chunk_end = ctx->chunk_end;
    p.K4 = ctx->K4; p.a = ctx->a; p.b = ctx->b; p.rtab = ctx->rtab; p.cam_fixed = ctx->cam_fixed;
    p.fix_structure = ctx->opt.fix_structure;
    p.ref_order = ctx->opt.order == VLG_BA_ORDER_REFERENCE ? 1 : 0;
    p.obs_slot = ctx->obs_slot; p.W = ctx->W; p.BeP = ctx->BeP; p.Upart = ctx->Upart;
    p.dX_hat = diag_X; p.dA = diag_A; p.dB = diag_B; p.de = diag_e;
    const size_t smem = (size_t)kWarpsPerBlock * (36 + NA + 4 + 32 * NU + 32 * 3 * NA) * sizeof(double);
    if (ctx->nchunks > 0) {
        TimedScope ts(ctx, T_STAGE1_CAM);
        if (diag) {
            CU(cudaFuncSetAttribute(k_stage1_cam<NA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_stage1_cam<NA, true><<<cdiv(ctx->nchunks, kWarpsPerBlock), kWarpsPerBlock * 32, smem, ctx->stream>>>(p);
        } else {
            CU(cudaFuncSetAttribute(k_stage1_cam<NA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_stage1_cam<NA, false><<<cdiv(ctx->nchunks, kWarpsPerBlock), kWarpsPerBlock * 32, smem, ctx->stream>>>(p);
        }
        ctx->launches++;
    }
    k_stage1_cam_finalize<NA><<<cdiv((int64_t)ctx->m * NU, 128), 128, 0, ctx->stream>>>(ctx->m, ctx->cam_chunk_ptr, ctx->Upart,
                                                                                         ctx->cam_fixed, ctx->U, ctx->eA);
    ctx->launches++;
    if (ctx->n > 0) {
        TimedScope ts(ctx, T_STAGE1_PT);
        if (ctx->ns1tiles > 0)
            k_stage1_pt_tiled<<<ctx->ns1tiles, kS1Tile, 0, ctx->stream>>>(ctx->s1tile_meta, ctx->pt_ptr, ctx->BeP, ctx->opt.fix_structure,
                                                                       ctx->V, ctx->eB, ctx->cost_pt);
        else
            k_stage1_pt<<<cdiv(ctx->n, 128), 128, 0, ctx->stream>>>(ctx->n, ctx->pt_ptr, ctx->BeP, ctx->opt.fix_structure, ctx->V, ctx->eB,
                                                                    ctx->cost_pt);
        ctx->launches++;
    }
    if (ctx->tiled_ok && ctx->nobs > 0 && !diag) {
        TimedScope ts(ctx, T_W_COPY);
        k_w_to_porder<NA><<<cdiv(ctx->nobs * 3 * NA, 256), 256, 0, ctx->stream>>>(ctx->nobs, ctx->pt_obs, ctx->W, ctx->Wp);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    CHK(reduce_to(ctx, ctx->cost_pt, (size_t)ctx->n, ctx->scal1));
    return VLG_BA_OK;
}

}  // namespace

// The remaining host logic (stage 2 solve paths, stage 3, LM driver, C ABI) follows in
// vlg_ba_host.inl to keep this translation unit readable.
#include "vlg_ba_host.inl"
