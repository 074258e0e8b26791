// ba_pcg.cuh -- the implicit-Schur PCG iteration, tiled for HBM streaming.
//
// One PCG iteration reads every W block twice (point-keyed sweep, camera-keyed sweep); at
// Venice shape that is 2 x 720 MB and nothing else of comparable size, so the iteration is
// HBM-bound and each sweep is built as a pure stream:
//   * W is kept in BOTH orders: C-order (written by stage 1) and a P-order copy Wp made once per
//     LM trial step by k_w_to_porder, so that each sweep reads one contiguous range per CTA;
//   * a CTA owns one tile (<= 512 observations of whole points for sweep 1, one <= 256-observation chunk of one camera
//     for sweep 2) and pulls it into shared memory with ONE TMA bulk copy
//     (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), several CTAs per SM keep
//     >100 KB per SM in flight; the small gathers (p_j, t_i) overlap the bulk copy;
//   * the 6m-vector algebra of an iteration is one cooperative kernel (two grid syncs).
// All reductions have a fixed shape (in-order per point, xor tree per warp, in-order over warps
// and chunks), so results are bit-reproducible run to run.
#pragma once
#include "ba_kernels.cuh"
#include "ba_chol.cuh"      // grid_barrier, fast_rcp_pos
#include <cooperative_groups.h>

namespace vlgba {

namespace cg = cooperative_groups;

// observations per tile == threads per sweep CTA.  Measured on B200 (Venice shape): point tiles
// of 512 (72 KB of W, 3 CTAs/SM) stream at 5.9 TB/s vs 5.3 TB/s at 256; camera chunks are better
// at 256 (4.98 vs 4.56 TB/s: a camera's last chunk is partial, and the reduction is per chunk).
constexpr int kPtTile = 512;
constexpr int kCamTile = 256;
constexpr int kCamWarps = kCamTile / 32;

// ---- TMA bulk copy + mbarrier (PTX ISA: cp.async.bulk, mbarrier) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

// global -> shared, contiguous, bytes % 16 == 0, both addresses 16-byte aligned.  The W
// stream (720 MB at Venice shape) is tagged L2 evict-first so that it does not push the small
// gathered vectors (t: 32 MB, written by sweep 1 and gathered by sweep 2) out of the 126 MB L2.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}

__device__ __forceinline__ void tma_load_1d_pol(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// Pulls a whole tile into shared memory.  TMA when the tile is 16-byte granular (issued by
// thread 0 straight after it initialised the mbarrier, i.e. before the CTA-wide barrier that
// publishes the init); otherwise (NA = 7 with an odd observation count/offset) a plain
// coalesced copy after that barrier.
template <int NW>
__device__ __forceinline__ bool tile_tma_ok(const double* src, int nob)
{
    return (((uint32_t)nob * NW * 8u) % 16u == 0u) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0u);
}

template <int NW>
__device__ __forceinline__ void load_tile_pre(double* dst, const double* src, int nob, uint64_t* bar)
{
    if (threadIdx.x == 0 && tile_tma_ok<NW>(src, nob)) {
        const uint32_t bytes = (uint32_t)nob * NW * 8u;
        mbar_expect_tx(bar, bytes);
        tma_load_1d(dst, src, bytes, bar);
    }
}

template <int NW>
__device__ __forceinline__ void load_tile_post(double* dst, const double* src, int nob)
{
    if (!tile_tma_ok<NW>(src, nob))
        for (int t = threadIdx.x; t < nob * NW; t += blockDim.x) dst[t] = src[t];
}

template <int NW>
__device__ __forceinline__ void wait_tile(const double* src, int nob, uint64_t* bar)
{
    if (tile_tma_ok<NW>(src, nob)) mbar_wait(bar, 0);
    else __syncthreads();
}

// A thread's NW-double block of the shared-memory tile -> registers.  With NW even the block
// is 16-byte aligned and read as double2: at a 144-byte stride a quarter-warp's LDS.128 covers
// all 32 banks exactly once (conflict-free), and it halves the MIO instruction count.
template <int NW>
__device__ __forceinline__ void load_block(const double* __restrict__ tile, int idx, double* __restrict__ w)
{
    if constexpr (NW % 2 == 0) {
        const double2* src = reinterpret_cast<const double2*>(tile + (size_t)idx * NW);
#pragma unroll
        for (int k = 0; k < NW / 2; k++) { const double2 v = src[k]; w[2 * k] = v.x; w[2 * k + 1] = v.y; }
    } else {
#pragma unroll
        for (int k = 0; k < NW; k++) w[k] = tile[(size_t)idx * NW + k];
    }
}

// Sum of NV values over the 32 lanes by recursive halving: at each xor step a lane keeps half
// of its values and ships the other half, so NV values cost ~NV shuffles instead of 5*NV.
// Returns, in every lane, the total of value index `warp_reduce_owner<NPAD>(lane)`; the
// combination order is fixed (bit-reproducible).
template <int NPAD>
__device__ __forceinline__ int warp_reduce_owner(int lane)
{
    if (NPAD == 8) return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

template <int NPAD>
__device__ __forceinline__ double warp_reduce_many(double* v, int lane)
{
    static_assert(NPAD == 8 || NPAD == 16, "NPAD");
    int width = NPAD;
#pragma unroll
    for (int mask = 16; width > 1; mask >>= 1, width >>= 1) {
        const bool hi = (lane & mask) != 0;
#pragma unroll
        for (int k = 0; k < NPAD / 2; k++) {
            if (k < width / 2) {
                const double send = hi ? v[k] : v[k + width / 2];
                const double keep = hi ? v[k + width / 2] : v[k];
                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
            }
        }
    }
    double s = v[0];
    if (NPAD == 8) { s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 1); }
    else { s += __shfl_xor_sync(0xffffffffu, s, 1); }
    return s;
}

// ---- P-order copy of W: Wp[q] = W[pt_obs[q]] ---------------------------------------------------
template <int NA>
__global__ void __launch_bounds__(256) k_w_to_porder(int64_t nobs, const int* __restrict__ pt_obs,
                                                     const double* __restrict__ W, double* __restrict__ Wp)
{
    constexpr int NW = 3 * NA;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nobs * NW) return;
    const int64_t q = t / NW;
    const int k = (int)(t - q * NW);
    Wp[t] = W[(size_t)pt_obs[q] * NW + k];
}

// ---- sweep 1, point-keyed: t_i = V*_i^-1 sum_j W_ij' p_j ---------------------------------------
// tile = points [ptile_first[b], ptile_first[b+1]) with <= kPtTile observations in all
template <int NA>
__global__ void __launch_bounds__(kPtTile)
k_sweep_pt_tiled(const int4* __restrict__ ptile_meta /* (q0, nob, p0, npts) */, const int* __restrict__ pt_ptr,
                 const int* __restrict__ pt_cam, const double* __restrict__ Wp, const double* __restrict__ Vinv,
                 const double* __restrict__ p, const int* __restrict__ done, double* __restrict__ t_out)
{
    constexpr int NW = 3 * NA;
    extern __shared__ __align__(128) unsigned char smraw[];
    double* wt = reinterpret_cast<double*>(smraw);            // kPtTile x NW
    double* sv = wt + kPtTile * NW;                          // kPtTile x 3
    uint64_t* bar = reinterpret_cast<uint64_t*>(sv + kPtTile * 3);
    const int tid = threadIdx.x;
    // one 16-byte descriptor per tile and the stop flag, fetched together: the bulk copy is
    // issued after a single (L2) round trip instead of three dependent ones
    const int4 meta = __ldg(ptile_meta + blockIdx.x);
    const int stop = done ? *done : 0;
    if (stop) return;
    const int q0 = meta.x, nob = meta.y, p0 = meta.z, np = meta.w;
    const double* src = Wp + (size_t)q0 * NW;
    if (tid == 0) mbar_init(bar, 1);
    load_tile_pre<NW>(wt, src, nob, bar);
    __syncthreads();
    load_tile_post<NW>(wt, src, nob);
    double pj[NA];
    if (tid < nob) {
        const double* pp = p + (size_t)NA * pt_cam[q0 + tid];
#pragma unroll
        for (int r = 0; r < NA; r++) pj[r] = __ldg(pp + r);
    }
    // everything the per-point epilogue needs is requested now, under the shadow of the bulk copy
    int o0 = 0, o1 = 0;
    double Vi[9];
    if (tid < np) {
        o0 = pt_ptr[p0 + tid] - q0; o1 = pt_ptr[p0 + tid + 1] - q0;
#pragma unroll
        for (int k = 0; k < 9; k++) Vi[k] = __ldg(Vinv + (size_t)9 * (p0 + tid) + k);
    }
    wait_tile<NW>(src, nob, bar);
    if (tid < nob) {
        double w[NW];
        load_block<NW>(wt, tid, w);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int r = 0; r < NA; r++) {
            s0 += w[r] * pj[r]; s1 += w[r + NA] * pj[r]; s2 += w[r + 2 * NA] * pj[r];
        }
        sv[tid * 3] = s0; sv[tid * 3 + 1] = s1; sv[tid * 3 + 2] = s2;
    }
    __syncthreads();
    if (tid < np) {
        const int i = p0 + tid;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int o = o0; o < o1; o++) { s0 += sv[o * 3]; s1 += sv[o * 3 + 1]; s2 += sv[o * 3 + 2]; }
        // t is stored padded to 4 doubles per point: the camera sweep gathers one aligned
        // 32-byte sector per observation
        double4 tv;
        tv.x = Vi[0] * s0 + Vi[3] * s1 + Vi[6] * s2;
        tv.y = Vi[1] * s0 + Vi[4] * s1 + Vi[7] * s2;
        tv.z = Vi[2] * s0 + Vi[5] * s1 + Vi[8] * s2;
        tv.w = 0.0;
        reinterpret_cast<double4*>(t_out)[i] = tv;
    }
}

// ---- sweep 2, camera-keyed: chunk partial of sum_i W_ij t_i ------------------------------------
// one CTA per chunk (<= 256 observations of one camera, contiguous in C-order W)
template <int NA>
__global__ void __launch_bounds__(kPtTile)
k_sweep_cam_tiled(const int2* __restrict__ chunk_meta /* (begin, nob) */, const int* __restrict__ obs_pt,
                  const double* __restrict__ W, const double* __restrict__ t_in, const int* __restrict__ done,
                  double* __restrict__ part /* [nchunks][NA] */)
{
    constexpr int NW = 3 * NA;
    extern __shared__ __align__(128) unsigned char smraw[];
    double* wt = reinterpret_cast<double*>(smraw);            // kCamTile x NW
    double* red = wt + kCamTile * NW;                         // kCamWarps x NA
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + kCamWarps * NA);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int2 meta = __ldg(chunk_meta + blockIdx.x);
    const int stop = done ? *done : 0;
    if (stop) return;
    const int beg = meta.x, nob = meta.y;
    const double* src = W + (size_t)beg * NW;
    if (tid == 0) mbar_init(bar, 1);
    load_tile_pre<NW>(wt, src, nob, bar);
    __syncthreads();
    load_tile_post<NW>(wt, src, nob);
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    if (tid < nob) {
        const int i = obs_pt[beg + tid];
        const double2 ta = __ldg(reinterpret_cast<const double2*>(t_in) + 2 * (size_t)i);
        const double2 tb = __ldg(reinterpret_cast<const double2*>(t_in) + 2 * (size_t)i + 1);
        t0 = ta.x; t1 = ta.y; t2 = tb.x;
    }
    wait_tile<NW>(src, nob, bar);
    constexpr int NPAD = NA <= 8 ? 8 : 16;
    double acc[NPAD];
#pragma unroll
    for (int r = 0; r < NPAD; r++) acc[r] = 0.0;
    if (tid < nob) {
        double w[NW];
        load_block<NW>(wt, tid, w);
#pragma unroll
        for (int r = 0; r < NA; r++) acc[r] = w[r] * t0 + w[r + NA] * t1 + w[r + 2 * NA] * t2;
    }
    {
        const double tot = warp_reduce_many<NPAD>(acc, lane);
        const int own = warp_reduce_owner<NPAD>(lane);
        if ((lane & (32 / NPAD - 1)) == 0 && own < NA) red[warp * NA + own] = tot;
    }
    __syncthreads();
    if (tid < NA) {
        double s = 0.0;
        const int nw = (nob + 31) >> 5;
        for (int w = 0; w < nw; w++) s += red[w * NA + tid];
        part[(size_t)NA * blockIdx.x + tid] = s;
    }
}

// ---- cluster-Jacobi preconditioner ----------------------------------------------------------------
// With S assembled, the preconditioner is the inverse of the diagonal blocks of S over CLUSTERS of
// consecutive cameras instead of single cameras: a cluster is the 128/NA cameras (21 for NA = 6, 126
// unknowns) that one CTA of the update kernel owns, so applying it is a dense 126 x 126 product
// against the CTA's own slice of r in shared memory (10 MB per iteration at Venice shape, 2 % of the
// matvec's traffic).  Measured on the CPU (m = 300, lambda = 1e-3..1e-5): 120/186/215 PCG iterations
// with per-camera blocks, 54/71/90 with 21-camera clusters.
// The inverse is pinv-like: non-positive pivots are eliminated (zero rows/columns), exactly like
// the per-camera blocks (sym_pinv) and the dense Cholesky.
template <int NA>
struct Cluster {
    static constexpr int kCams = 128 / NA;
    static constexpr int NC = kCams * NA;          // unknowns per cluster (<= 128)
    static constexpr int LD = NC | 1;              // odd row stride in shared memory: conflict-free rows
    static constexpr size_t kSmem = sizeof(double) * (size_t)NC * LD;
    // second, SHIFTED partition (overlapping additive Schwarz, see k_cluster_inverse): shifted block k covers the
    // cameras [k kCams - kShift, k kCams + kHalf), i.e. the last kShift cameras of cluster k-1 and the first kHalf
    // cameras of cluster k
    static constexpr int kHalf = kCams / 2;
    static constexpr int kShift = kCams - kHalf;
};

// Cblk[cl][r + 128 c] = this rank's S block of cluster cl (without U*: added after the all-reduce); `shift` > 0
// gathers the blocks of the shifted partition (gridDim.x = clusters + 1), cameras outside [0, m) give zeros
template <int NA>
__global__ void __launch_bounds__(128)
k_cluster_gather(int m, int ld, const double* __restrict__ S, double* __restrict__ Cblk, int shift)
{
    using C = Cluster<NA>;
    const int cl = blockIdx.x, t = threadIdx.x;
    const long long base = (long long)NA * ((long long)cl * C::kCams - shift), N = (long long)NA * m;
    const long long gr = base + t;
    const bool rok = t < C::NC && gr >= 0 && gr < N;
    for (int c = 0; c < 128; c++) {
        const long long gc = base + c;
        Cblk[((size_t)cl * 128 + c) * 128 + t] = (rok && c < C::NC && gc >= 0 && gc < N) ? S[gr + (long long)ld * gc] : 0.0;
    }
}

// In-place Gauss-Jordan inversion of a cluster block (SPD => no pivot search).  The block lives in REGISTERS:
// 256 threads as a 16 x 16 grid, thread (ty, tx) owns the 8 x 8 elements (ty + 16 i, tx + 16 j).  Per pivot k the
// owners of row k and of column k publish them through shared memory (double-buffered: one barrier per pivot)
// and every thread does 64 independent FMAs on its registers against 8 + 8 broadcast values -- the first
// version kept the block in shared memory and was bound by its bandwidth (3 accesses per FMA: 0.36 ms for 85
// blocks at Venice shape; this one: see DESIGN.md).  The pivot loop is unrolled over i (the register index of
// the pivot row/column must be static) and runs over the 16 owners inside.  A non-positive pivot eliminates
// its row and column (zero row/column of the inverse: pinv-like, as sym_pinv and the dense Cholesky do).
//
// shift == 0: block cl = cluster cl, inverse -> McL[cl] (column-major, ld 128, symmetric by construction).
// shift  > 0: block cl of the shifted partition.  Its rows go to the COMPOSITE Mc2: cluster c's unknown t gets
//   the row of the shifted block that contains its camera -- block c for the first kHalf cameras of the
//   cluster (block rows kShift NA + t), block c + 1 for the rest (block rows t - kHalf NA) -- as
//   Mc2[c][t + 128 col], col = the block's column.  z = (M1^-1 + M2^-1) r is then two 126-term dot products
//   per unknown, the second against the r of the cameras [c kCams - kShift, ...) or [c kCams + kHalf, ...).
template <int NA>
__global__ void __launch_bounds__(256)
k_cluster_inverse(int m, int add_U, int shift, int nclusters, const double* __restrict__ Cblk, const double* __restrict__ Ud,
                  double* __restrict__ Mout)
{
    using C = Cluster<NA>;
    extern __shared__ double A[];                  // NC x LD, row-major: only for the symmetric copy-out
    __shared__ double rowb[2][128], colb[2][128];
    const int cl = blockIdx.x, tid = threadIdx.x;
    const int ty = tid & 15, tx = tid >> 4;
    const int cam0 = cl * C::kCams - shift;
    double a[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int r = ty + 16 * i, c = tx + 16 * j;
            const int jr = cam0 + r / NA, jc = cam0 + c / NA;
            double v = 0.0;
            if (r < C::NC && c < C::NC && jr >= 0 && jr < m && jc >= 0 && jc < m) {
                v = Cblk[((size_t)cl * 128 + c) * 128 + r];
                if (add_U && jr == jc) v += Ud[(size_t)NA * NA * jr + (r - (r / NA) * NA) + NA * (c - (c / NA) * NA)];
            }
            a[i][j] = v;
        }
    int buf = 0;
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
        for (int kq = 0; kq < 16; kq++) {
            const int k = kq + 16 * kk;
            if (k >= C::NC) break;                 // uniform
            if (ty == kq) {
#pragma unroll
                for (int j = 0; j < 8; j++) rowb[buf][tx + 16 * j] = a[kk][j];
            }
            if (tx == kq) {
#pragma unroll
                for (int i = 0; i < 8; i++) colb[buf][ty + 16 * i] = a[i][kk];
            }
            __syncthreads();
            const double p = fast_rcp_pos(rowb[buf][k]);      // 0 for a non-positive pivot
            double rk[8], ci[8];
#pragma unroll
            for (int j = 0; j < 8; j++) rk[j] = rowb[buf][tx + 16 * j] * p;
#pragma unroll
            for (int i = 0; i < 8; i++) ci[i] = colb[buf][ty + 16 * i];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) a[i][j] = fma(-ci[i], rk[j], a[i][j]);
            if (ty == kq) {                        // the pivot row: scaled by 1/d
#pragma unroll
                for (int j = 0; j < 8; j++) a[kk][j] = rk[j];
            }
            if (tx == kq) {                        // the pivot column
#pragma unroll
                for (int i = 0; i < 8; i++) a[i][kk] = -ci[i] * p;
            }
            if (ty == kq && tx == kq) a[kk][kk] = p;
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int r = ty + 16 * i, c = tx + 16 * j;
            if (r < C::NC && c < C::NC) A[r * C::LD + c] = a[i][j];
        }
    __syncthreads();
    // symmetric copy out (lower triangle mirrored); rows/columns of absent cameras are zero
    const int t = tid & 127, half = tid >> 7;
    if (shift == 0) {
        double* out = Mout + (size_t)cl * 128 * 128;
        for (int c = half; c < 128; c += 2) {
            double v = 0.0;
            if (t < C::NC && c < C::NC) v = t >= c ? A[t * C::LD + c] : A[c * C::LD + t];
            out[t + 128 * c] = v;
        }
    } else if (t < C::NC) {
        const bool upper = t >= C::kShift * NA;                // first kHalf cameras of cluster cl
        const int dst = upper ? cl : cl - 1;
        const int trow = upper ? t - C::kShift * NA : t + C::kHalf * NA;
        if (dst >= 0 && dst < nclusters) {
            double* out = Mout + (size_t)dst * 128 * 128 + trow;
            for (int c = half; c < C::NC; c += 2) out[128 * c] = t >= c ? A[t * C::LD + c] : A[c * C::LD + t];
        }
    }
}

// second term of the overlapping preconditioner for unknown t from a global r: the composite row of Mc2 against
// the r of the shifted block that holds t's camera
template <int NA, bool COHERENT>
__device__ __forceinline__ double apply_cluster2_global(int m, int t, const double* __restrict__ Mc2, const double* r)
{
    using C = Cluster<NA>;
    const int cl = t / C::NC, lt = t - cl * C::NC;
    const long long N = (long long)NA * m;
    const long long base = (long long)NA * ((long long)cl * C::kCams - C::kShift) + (lt < C::kHalf * NA ? 0 : C::NC);
    const double* M = Mc2 + (size_t)cl * 128 * 128 + lt;
    double z0 = 0.0, z1 = 0.0;
    for (int c = 0; c < C::NC; c += 2) {
        const long long g0 = base + c, g1 = g0 + 1;
        if (g0 >= 0 && g0 < N) z0 += M[128 * c] * (COHERENT ? __ldcg(r + g0) : r[g0]);
        if (g1 >= 0 && g1 < N) z1 += M[128 * (c + 1)] * (COHERENT ? __ldcg(r + g1) : r[g1]);
    }
    return z0 + z1;
}

// z_t = (M^-1 r)_t for unknown t from global r: per-camera blocks, or the cluster blocks when McL is given
template <int NA>
__device__ __forceinline__ double apply_precond_global(int m, int t, const double* __restrict__ Minv,
                                                       const double* __restrict__ McL, const double* __restrict__ r,
                                                       const double* __restrict__ Mc2 = nullptr)
{
    using C = Cluster<NA>;
    double zz = 0.0;
    if (McL) {
        const int cl = t / C::NC, lt = t - cl * C::NC;
        const int nc = min(C::kCams, m - cl * C::kCams) * NA;
        const double* M = McL + (size_t)cl * 128 * 128 + lt;
        const double* rc = r + (size_t)cl * C::NC;
        double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
        int c = 0;
        for (; c + 3 < nc; c += 4) {
            z0 += M[128 * c] * rc[c]; z1 += M[128 * (c + 1)] * rc[c + 1];
            z2 += M[128 * (c + 2)] * rc[c + 2]; z3 += M[128 * (c + 3)] * rc[c + 3];
        }
        for (; c < nc; c++) z0 += M[128 * c] * rc[c];
        zz = (z0 + z1) + (z2 + z3);
        if (Mc2) zz += apply_cluster2_global<NA, false>(m, t, Mc2, r);
    } else {
        const int j = t / NA, row = t - j * NA;
#pragma unroll
        for (int c = 0; c < NA; c++) zz += Minv[(size_t)NA * NA * j + row + NA * c] * r[(size_t)NA * j + c];
    }
    return zz;
}

// r = e_, x = 0, z = M^-1 r, p = z, rz = r'z, r0n2 = r'r
template <int NA>
__global__ void __launch_bounds__(1024) k_pcg_init(int m, const double* __restrict__ ebar, const double* __restrict__ Minv,
                                                   double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                                                   double* __restrict__ p, PcgScalars* __restrict__ sc, double rtol,
                                                   const double* __restrict__ McL = nullptr, const double* __restrict__ Mc2 = nullptr)
{
    __shared__ double sh[32];
    const int N = NA * m;
    double rz = 0.0, rr = 0.0;
    for (int t = threadIdx.x; t < N; t += 1024) {
        const double zz = apply_precond_global<NA>(m, t, Minv, McL, ebar, Mc2);
        const double rv = ebar[t];
        x[t] = 0.0; r[t] = rv; z[t] = zz; p[t] = zz;
        rz += rv * zz; rr += rv * rv;
    }
    rz = block_sum_1024(rz, sh);
    rr = block_sum_1024(rr, sh);
    if (threadIdx.x == 0) {
        sc->rz = rz; sc->r0n2 = rr; sc->rn2 = rr; sc->pq = 0.0; sc->iters = 0;
        sc->done = (rr == 0.0) ? 1 : 0;
        (void)rtol;
    }
}

// ---- the vector algebra of one PCG iteration, one cooperative kernel ---------------------------
// one thread per camera (its NA unknowns stay in registers through the three phases):
//   q = U* p - W V*^-1 W' p ;  alpha = r'z / p'q ;  x += alpha p ;  r -= alpha q ;
//   z = M^-1 r ;  beta = r'z_new / r'z ;  p = z + beta p ;  stop when |r| <= rtol |r0|
__device__ __forceinline__ double block_sum_fixed(double v, double* sh)
{
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; w++) s += sh[w];
    return s;   // every thread, same order
}

// Deflation space: the 4 gauge directions the cost cannot see (world translation x3, scale) --
// exactly the 4 eigenvalues ~lambda that stall block-Jacobi PCG on S.
constexpr int kDefl = 4;
struct DeflScalars { double Einv[kDefl * kDefl]; double c0[kDefl]; };

// BATCH: the scalar partials of all blocks are loaded into registers before they are summed (faster: no
// L2 round trip per partial on the in-order issue path, but ~125 registers: at most 4 CTAs per SM); the
// host falls back to BATCH = false (48 registers) when the grid would not be co-resident (Final shape:
// 652 clusters).
template <int NA, bool BATCH>
__global__ void __launch_bounds__(128)
k_pcg_update_coop(int m, const int* __restrict__ cam_chunk_ptr, const double* __restrict__ qpart,
                  const double* __restrict__ wq, const double* __restrict__ Ud, const double* __restrict__ Minv,
                  double* __restrict__ x, double* __restrict__ r, double* __restrict__ p,
                  PcgScalars* __restrict__ sc, double* __restrict__ blkpart /* 11 * gridDim.x */, double rtol,
                  const double* __restrict__ Z /* [kDefl][N] or NULL */, const double* __restrict__ SZ,
                  const DeflScalars* __restrict__ ds, const double* __restrict__ McL /* cluster inverses or NULL */,
                  int mcl_in_smem /* launched with room for the CTA's cluster inverse in dynamic shared memory */)
{
    // one thread per reduced unknown; a CTA owns kCams whole cameras so that the NA x NA block
    // products (U* p, M^-1 r) only need the CTA's own slice of p and r (shared memory)
    constexpr int ND = 1 + 2 * kDefl;             // p'w, Z'w, SZ'p
    constexpr int kCams = 128 / NA;
    __shared__ double shd[4 * ND];
    __shared__ double pv[128], rv[128];
    __shared__ uint64_t mbar;
    extern __shared__ __align__(128) unsigned char smraw[];   // the CTA's cluster inverse (NC columns x 128), when McL is given
    double* Ms = reinterpret_cast<double*>(smraw);
    cg::grid_group grid = cg::this_grid();
    if (sc->done) return;                         // uniform: sc is only written after the last grid sync
    if (McL && mcl_in_smem) {
        // the 126 x 126 preconditioner block of this CTA is needed only after the first grid sync: pull it
        // into shared memory now with one TMA bulk copy (129 KB, contiguous), so that z = M^-1 r reads shared
        // memory instead of 126 dependent-latency L2 loads per thread
        if (threadIdx.x == 0) {
            mbar_init(&mbar, 1);
            constexpr uint32_t bytes = (uint32_t)(kCams * NA) * 128u * 8u;
            mbar_expect_tx(&mbar, bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(Ms)), "l"(McL + (size_t)blockIdx.x * 128 * 128), "r"(bytes), "r"(smem_u32(&mbar)) : "memory");
        }
    }
    const double rz = sc->rz, r0n2 = sc->r0n2;
    const int nb = gridDim.x;
    const size_t N = (size_t)NA * m;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lc = threadIdx.x / NA, row = threadIdx.x - lc * NA;       // local camera, row
    const int j = blockIdx.x * kCams + lc;
    const bool act = lc < kCams && j < m;
    const size_t t = (size_t)NA * j + row;
    double dots[ND];
#pragma unroll
    for (int k = 0; k < ND; k++) dots[k] = 0.0;
    double pt = 0.0, qt = 0.0;
    if (act) { pt = p[t]; pv[threadIdx.x] = pt; }
    __syncthreads();
    if (act) {
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < NA; c++) v += Ud[(size_t)NA * NA * j + row + NA * c] * pv[lc * NA + c];
        double w = 0.0;
        if (wq) w = wq[t];
        else for (int c = cam_chunk_ptr[j]; c < cam_chunk_ptr[j + 1]; c++) w += qpart[(size_t)NA * c + row];
        qt = v - w;
        dots[0] = pt * qt;
        if (Z) {
#pragma unroll
            for (int d = 0; d < kDefl; d++) {
                dots[1 + d] = Z[d * N + t] * qt;
                dots[1 + kDefl + d] = SZ[d * N + t] * pt;
            }
        }
    }
    const int nd = Z ? ND : 1;
    // all nd block sums with one barrier: xor tree per warp, then warps in order
    for (int k = 0; k < nd; k++) {
        const double v = warp_sum(dots[k]);
        if (lane == 0) shd[warp * ND + k] = v;
    }
    __syncthreads();
    if (threadIdx.x < nd) {
        double v = 0.0;
        for (int w = 0; w < 4; w++) v += shd[w * ND + threadIdx.x];
        blkpart[(size_t)threadIdx.x * nb + blockIdx.x] = v;
    }
    grid.sync();
    // every warp folds the per-block partials the same way (strided by lane, xor tree)
    if constexpr (!BATCH) {
        for (int k = 0; k < nd; k++) {
            double v = 0.0;
            for (int b = lane; b < nb; b += 32) v += __ldcg(blkpart + (size_t)k * nb + b);
            dots[k] = warp_sum(v);
        }
    } else {
        // every partial load is issued before the first add (an add right behind its load stalls the
        // in-order issue for an L2 round trip per partial: measured in k_pcg_persistent)
        constexpr int KB = 4;
        double ldv[ND][KB];
#pragma unroll
        for (int k = 0; k < ND; k++)
#pragma unroll
            for (int u = 0; u < KB; u++) {
                const int b = lane + 32 * u;
                ldv[k][u] = (k < nd && b < nb) ? __ldcg(blkpart + (size_t)k * nb + b) : 0.0;
            }
#pragma unroll
        for (int k = 0; k < ND; k++) {
            double v = 0.0;
#pragma unroll
            for (int u = 0; u < KB; u++) v += ldv[k][u];
            if (k < nd)
                for (int b = lane + 32 * KB; b < nb; b += 32) v += __ldcg(blkpart + (size_t)k * nb + b);
            dots[k] = k < nd ? warp_sum(v) : 0.0;
        }
    }
    // deflated operator: w <- P S p = S p - SZ Einv Z' S p ;  p'(P S p) = p'Sp - (SZ'p)' Einv (Z'Sp)
    double y[kDefl];
    double pq = dots[0];
    if (Z) {
#pragma unroll
        for (int d = 0; d < kDefl; d++) {
            double v = 0.0;
#pragma unroll
            for (int e = 0; e < kDefl; e++) v += ds->Einv[d + kDefl * e] * dots[1 + e];
            y[d] = v;
            pq -= dots[1 + kDefl + d] * v;
        }
    }
    if (!(pq > 0.0)) {                            // breakdown: every thread sees the same pq
        if (blockIdx.x == 0 && threadIdx.x == 0) { sc->done = 2; sc->pq = pq; }
        if (McL && mcl_in_smem && threadIdx.x == 0) mbar_wait(&mbar, 0);      // the staged inverse must land before the CTA exits
        __syncthreads();
        return;
    }
    const double alpha = rz / pq;
    double rzn = 0.0, rr = 0.0, zt = 0.0;
    if (act) {
        if (Z) {
#pragma unroll
            for (int d = 0; d < kDefl; d++) qt -= SZ[d * N + t] * y[d];
        }
        x[t] += alpha * pt;
        const double rt = r[t] - alpha * qt;
        r[t] = rt;
        rv[threadIdx.x] = rt;
        rr = rt * rt;
    } else {
        rv[threadIdx.x] = 0.0;
    }
    __syncthreads();
    if (act) {
        if (McL) {
            // cluster block of this CTA (column-major, ld 128: coalesced over threads), 4 independent chains
            // column-major, ld 128: consecutive threads read consecutive words; from shared memory when it fits
            // (one CTA per SM, <= 148 clusters), else from L2
            if (mcl_in_smem) mbar_wait(&mbar, 0);
            const double* M = mcl_in_smem ? Ms + threadIdx.x : McL + (size_t)blockIdx.x * 128 * 128 + threadIdx.x;
            double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
            constexpr int NC = kCams * NA;
#pragma unroll 4
            for (int c = 0; c + 3 < NC; c += 4) {
                z0 += M[128 * c] * rv[c];
                z1 += M[128 * (c + 1)] * rv[c + 1];
                z2 += M[128 * (c + 2)] * rv[c + 2];
                z3 += M[128 * (c + 3)] * rv[c + 3];
            }
#pragma unroll
            for (int c = NC - NC % 4; c < NC; c++) z0 += M[128 * c] * rv[c];
            zt = (z0 + z1) + (z2 + z3);
        } else {
#pragma unroll
            for (int c = 0; c < NA; c++) zt += Minv[(size_t)NA * NA * j + row + NA * c] * rv[lc * NA + c];
        }
        rzn = rv[threadIdx.x] * zt;
    }
    rzn = warp_sum(rzn);
    rr = warp_sum(rr);
    if (lane == 0) { shd[warp * 2] = rzn; shd[warp * 2 + 1] = rr; }   // shd's first use ended before grid.sync
    __syncthreads();
    double* bp2 = blkpart + (size_t)ND * nb;
    if (threadIdx.x < 2) {
        double v = 0.0;
        for (int w = 0; w < 4; w++) v += shd[w * 2 + threadIdx.x];
        bp2[2 * blockIdx.x + threadIdx.x] = v;
    }
    grid.sync();
    {
        double v0 = 0.0, v1 = 0.0;
        for (int b = lane; b < nb; b += 32) { v0 += __ldcg(bp2 + 2 * b); v1 += __ldcg(bp2 + 2 * b + 1); }
        rzn = warp_sum(v0); rr = warp_sum(v1);
    }
    const double beta = rzn / rz;
    if (act) p[t] = zt + beta * pt;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->rz = rzn; sc->rn2 = rr; sc->pq = pq; sc->iters += 1;
        if (rr <= rtol * rtol * r0n2) sc->done = 1;
    }
}

// gauge vectors restricted to the cameras: world translation by e_d moves T_j by -R_j e_d,
// scaling the world scales T_j; rotations and intrinsics do not move.  Z = [kDefl][N].
template <int NA>
__global__ void k_gauge_vectors(int m, const double* __restrict__ a, const double* __restrict__ rtab,
                                const unsigned char* __restrict__ cam_fixed, const int* __restrict__ cam_chunk_ptr,
                                double* __restrict__ Z)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const size_t N = (size_t)NA * m;
    const double* R = rtab + (size_t)36 * j;
    for (int d = 0; d < kDefl; d++)
        for (int k = 0; k < NA; k++) Z[d * N + (size_t)NA * j + k] = 0.0;
    // rows of S that are structurally zero (fixed or unobserved cameras) must stay out of the
    // deflation space: pinv(S) leaves them at zero (bundle_euclid.m:145-154,193)
    if (cam_fixed[j] || cam_chunk_ptr[j + 1] == cam_chunk_ptr[j]) return;
    for (int d = 0; d < 3; d++)
        for (int r = 0; r < 3; r++) Z[d * N + (size_t)NA * j + 3 + r] = -R[r + 3 * d];
    for (int r = 0; r < 3; r++) Z[3 * N + (size_t)NA * j + 3 + r] = a[(size_t)NA * j + 3 + r];
}

// out = S v for one vector from the sweep partials: out_j = U*_j v_j - (W V*^-1 W' v)_j
template <int NA>
__global__ void k_apply_S_finalize(int m, const int* __restrict__ cam_chunk_ptr, const double* __restrict__ qpart,
                                   const double* __restrict__ wq, const double* __restrict__ Ud,
                                   const double* __restrict__ v, double* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * NA) return;
    const int j = t / NA, row = t % NA;
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < NA; c++) s += Ud[(size_t)NA * NA * j + row + NA * c] * v[(size_t)NA * j + c];
    double w = 0.0;
    if (wq) w = wq[t];
    else for (int c = cam_chunk_ptr[j]; c < cam_chunk_ptr[j + 1]; c++) w += qpart[(size_t)NA * c + row];
    out[t] = s - w;
}

// E = Z'SZ, Einv, c0 = Z'b; then the deflated start: x^ = 0, r = b - SZ Einv c0, z = M^-1 r, p = z.
// The stop test stays relative to |b| (the reduced right-hand side).
template <int NA>
__global__ void __launch_bounds__(1024)
k_pcg_init_defl(int m, const double* __restrict__ ebar, const double* __restrict__ Minv, const double* __restrict__ Z,
                const double* __restrict__ SZ, DeflScalars* __restrict__ ds, double* __restrict__ x,
                double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, PcgScalars* __restrict__ sc,
                const double* __restrict__ McL, const double* __restrict__ Mc2 = nullptr)
{
    __shared__ double sh[32];
    __shared__ double E[kDefl * kDefl], Ei[kDefl * kDefl], c0[kDefl], y0[kDefl];
    const int N = NA * m;
    for (int d = 0; d < kDefl; d++)
        for (int e = d; e < kDefl; e++) {
            double v = 0.0;
            for (int t = threadIdx.x; t < N; t += 1024) v += 0.5 * (Z[(size_t)d * N + t] * SZ[(size_t)e * N + t] + Z[(size_t)e * N + t] * SZ[(size_t)d * N + t]);
            v = block_sum_1024(v, sh);
            if (threadIdx.x == 0) { E[d + kDefl * e] = v; E[e + kDefl * d] = v; }
        }
    for (int d = 0; d < kDefl; d++) {
        double v = 0.0;
        for (int t = threadIdx.x; t < N; t += 1024) v += Z[(size_t)d * N + t] * ebar[t];
        v = block_sum_1024(v, sh);
        if (threadIdx.x == 0) c0[d] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        sym_pinv<kDefl>(E, Ei);
        for (int d = 0; d < kDefl; d++) {
            double v = 0.0;
            for (int e = 0; e < kDefl; e++) v += Ei[d + kDefl * e] * c0[e];
            y0[d] = v;
            ds->c0[d] = c0[d];
        }
        for (int k = 0; k < kDefl * kDefl; k++) ds->Einv[k] = Ei[k];
    }
    __syncthreads();
    double rz = 0.0, bb = 0.0;
    for (int t = threadIdx.x; t < N; t += 1024) {
        double rv = ebar[t];
        bb += rv * rv;
        for (int d = 0; d < kDefl; d++) rv -= SZ[(size_t)d * N + t] * y0[d];
        r[t] = rv; x[t] = 0.0;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < N; t += 1024) {
        const double zz = apply_precond_global<NA>(m, t, Minv, McL, r, Mc2);
        z[t] = zz; p[t] = zz;
        rz += r[t] * zz;
    }
    rz = block_sum_1024(rz, sh);
    bb = block_sum_1024(bb, sh);
    if (threadIdx.x == 0) {
        sc->rz = rz; sc->r0n2 = bb; sc->rn2 = bb; sc->pq = 0.0; sc->iters = 0;
        sc->done = (bb == 0.0 || rz == 0.0) ? 1 : 0;
    }
}

// The same set-up on the update kernel's grid (one CTA per cluster, thread per unknown, two grid
// barriers): the single-CTA version above is a chain of 16 block reductions plus the cluster
// product by 1024 threads -- 0.39 ms per LM step at Venice shape, 3.8 % of the step.
template <int NA>
__global__ void __launch_bounds__(128)
k_pcg_init_defl_coop(int m, const double* __restrict__ ebar, const double* __restrict__ Minv, const double* __restrict__ Z,
                     const double* __restrict__ SZ, DeflScalars* __restrict__ ds, double* __restrict__ x,
                     double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, PcgScalars* __restrict__ sc,
                     const double* __restrict__ McL, double* __restrict__ part /* 17 * gridDim.x */,
                     unsigned int* __restrict__ barrier, const double* __restrict__ Mc2)
{
    constexpr int kCams = 128 / NA, NC = kCams * NA;
    constexpr int NP = kDefl * (kDefl + 1) / 2 + kDefl + 1;      // E (upper triangle), c0, b'b
    __shared__ double shd[4 * NP], bc[NP + 1], rv[128];
    __shared__ double E[kDefl * kDefl], Ei[kDefl * kDefl], y0[kDefl];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, nb = gridDim.x;
    const size_t N = (size_t)NA * m;
    const int lc = tid / NA, row = tid - lc * NA;
    const int j = cta * kCams + lc;
    const bool act = lc < kCams && j < m;
    const size_t t = (size_t)NA * j + row;
    unsigned int bar_target = 0;
    double zt[kDefl], st[kDefl], bt = 0.0;
#pragma unroll
    for (int d = 0; d < kDefl; d++) { zt[d] = act ? Z[d * N + t] : 0.0; st[d] = act ? SZ[d * N + t] : 0.0; }
    if (act) bt = ebar[t];
    double pr[NP];
    {
        int k = 0;
#pragma unroll
        for (int d = 0; d < kDefl; d++)
#pragma unroll
            for (int e = d; e < kDefl; e++) pr[k++] = 0.5 * (zt[d] * st[e] + zt[e] * st[d]);
#pragma unroll
        for (int d = 0; d < kDefl; d++) pr[k++] = zt[d] * bt;
        pr[k] = bt * bt;
    }
#pragma unroll
    for (int k = 0; k < NP; k++) {
        const double v = warp_sum(pr[k]);
        if (lane == 0) shd[warp * NP + k] = v;
    }
    __syncthreads();
    if (tid < NP) {
        double v = 0.0;
        for (int w = 0; w < 4; w++) v += shd[w * NP + tid];
        part[(size_t)tid * nb + cta] = v;
    }
    grid_barrier(barrier, bar_target, nb);
    if (warp == 0) {
        constexpr int KB = 5;
        double ldv[NP][KB];
#pragma unroll
        for (int k = 0; k < NP; k++)
#pragma unroll
            for (int u = 0; u < KB; u++) {
                const int b = lane + 32 * u;
                ldv[k][u] = b < nb ? __ldcg(part + (size_t)k * nb + b) : 0.0;
            }
#pragma unroll
        for (int k = 0; k < NP; k++) {
            double v = 0.0;
#pragma unroll
            for (int u = 0; u < KB; u++) v += ldv[k][u];
            for (int b = lane + 32 * KB; b < nb; b += 32) v += __ldcg(part + (size_t)k * nb + b);
            v = warp_sum(v);
            if (lane == 0) bc[k] = v;
        }
        if (lane == 0) {
            int k = 0;
            for (int d = 0; d < kDefl; d++)
                for (int e = d; e < kDefl; e++) { E[d + kDefl * e] = bc[k]; E[e + kDefl * d] = bc[k]; k++; }
            sym_pinv<kDefl>(E, Ei);
            for (int d = 0; d < kDefl; d++) {
                double v = 0.0;
                for (int e = 0; e < kDefl; e++) v += Ei[d + kDefl * e] * bc[kDefl * (kDefl + 1) / 2 + e];
                y0[d] = v;
            }
            if (cta == 0) {
                for (int d = 0; d < kDefl; d++) ds->c0[d] = bc[kDefl * (kDefl + 1) / 2 + d];
                for (int q = 0; q < kDefl * kDefl; q++) ds->Einv[q] = Ei[q];
            }
        }
    }
    __syncthreads();
    double rt = 0.0, zz = 0.0;
    if (act) {
        rt = bt;
#pragma unroll
        for (int d = 0; d < kDefl; d++) rt -= st[d] * y0[d];
        r[t] = rt; x[t] = 0.0;
    }
    rv[tid] = rt;
    __syncthreads();
    if (Mc2) grid_barrier(barrier, bar_target, nb);     // the overlapping term reads the neighbours' r
    if (act) {
        if (McL) {
            const double* M = McL + (size_t)cta * 128 * 128 + tid;
            double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
#pragma unroll 4
            for (int c = 0; c + 3 < NC; c += 4) {
                z0 += __ldg(M + 128 * c) * rv[c]; z1 += __ldg(M + 128 * (c + 1)) * rv[c + 1];
                z2 += __ldg(M + 128 * (c + 2)) * rv[c + 2]; z3 += __ldg(M + 128 * (c + 3)) * rv[c + 3];
            }
#pragma unroll
            for (int c = NC - NC % 4; c < NC; c++) z0 += __ldg(M + 128 * c) * rv[c];
            zz = (z0 + z1) + (z2 + z3);
            if (Mc2) zz += apply_cluster2_global<NA, true>(m, (int)t, Mc2, r);
        } else {
#pragma unroll
            for (int c = 0; c < NA; c++) zz += Minv[(size_t)NA * NA * j + row + NA * c] * rv[lc * NA + c];
        }
        z[t] = zz; p[t] = zz;
    }
    {
        const double v = warp_sum(act ? rt * zz : 0.0);
        if (lane == 0) shd[warp] = v;
    }
    __syncthreads();
    if (tid == 0) part[(size_t)NP * nb + cta] = (shd[0] + shd[1]) + (shd[2] + shd[3]);
    grid_barrier(barrier, bar_target, nb);
    if (cta == 0 && warp == 0) {
        double v = 0.0;
        for (int b = lane; b < nb; b += 32) v += __ldcg(part + (size_t)NP * nb + b);
        v = warp_sum(v);
        if (lane == 0) {
            const double bb = bc[NP - 1];
            sc->rz = v; sc->r0n2 = bb; sc->rn2 = bb; sc->pq = 0.0; sc->iters = 0;
            sc->done = (bb == 0.0 || v == 0.0) ? 1 : 0;
        }
    }
}

// x = x^ + Z Einv (Z'b - SZ'x^)
template <int NA>
__global__ void __launch_bounds__(1024)
k_pcg_defl_final(int m, const double* __restrict__ Z, const double* __restrict__ SZ, const DeflScalars* __restrict__ ds,
                 double* __restrict__ x)
{
    __shared__ double sh[32];
    __shared__ double y[kDefl];
    const int N = NA * m;
    double g[kDefl];
    for (int d = 0; d < kDefl; d++) {
        double v = 0.0;
        for (int t = threadIdx.x; t < N; t += 1024) v += SZ[(size_t)d * N + t] * x[t];
        g[d] = ds->c0[d] - block_sum_1024(v, sh);
    }
    if (threadIdx.x == 0)
        for (int d = 0; d < kDefl; d++) {
            double v = 0.0;
            for (int e = 0; e < kDefl; e++) v += ds->Einv[d + kDefl * e] * g[e];
            y[d] = v;
        }
    __syncthreads();
    for (int t = threadIdx.x; t < N; t += 1024) {
        double v = x[t];
        for (int d = 0; d < kDefl; d++) v += Z[(size_t)d * N + t] * y[d];
        x[t] = v;
    }
}

}  // namespace vlgba

// =========================================================================================
// Persistent ("ring") versions of the two sweeps.  tools/stream_probe.cu shows that a CTA-per-
// tile TMA stream reaches 7.2 TB/s when the CTA does nothing else, but every sweep CTA also has
// a prologue (descriptor fetch) and an epilogue (reduction, store) during which its shared-
// memory tile is not being refilled (measured: 5.0-5.3 TB/s).  Here a CTA stays resident, owns
// every gridDim.x-th tile and keeps kStages bulk copies in flight: the copy of tile k+kStages is
// issued the moment tile k has been consumed, the small gathers of tile k+1 (and their indices
// for k+2) are requested one tile ahead, and all tile descriptors are fetched once up front.
// =========================================================================================
namespace vlgba {

constexpr int kRingTile = 256;      // observations per ring tile == threads per CTA
constexpr int kRingMaxTiles = 384;  // tile descriptors cached in shared memory per CTA

template <int NA, int STAGES>
__global__ void __launch_bounds__(kRingTile)
k_sweep_cam_ring(int ntiles, const int2* __restrict__ chunk_meta, const int* __restrict__ obs_pt,
                 const double* __restrict__ W, const double* __restrict__ t_in, const int* __restrict__ done,
                 double* __restrict__ part /* [nchunks][NA] */)
{
    constexpr int NW = 3 * NA;
    constexpr int NPAD = NA <= 8 ? 8 : 16;
    constexpr int NWARP = kRingTile / 32;
    extern __shared__ __align__(128) unsigned char smraw[];
    double* wt = reinterpret_cast<double*>(smraw);                       // STAGES x kRingTile x NW
    double* red = wt + (size_t)STAGES * kRingTile * NW;                  // 2 x NWARP x NA
    int2* metas = reinterpret_cast<int2*>(red + 2 * NWARP * NA);         // kRingMaxTiles
    uint64_t* bar = reinterpret_cast<uint64_t*>(metas + kRingMaxTiles);  // STAGES
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (done && *done) return;
    const int first = blockIdx.x, step = gridDim.x;
    const int mine = first < ntiles ? (ntiles - first + step - 1) / step : 0;
    for (int k = tid; k < mine; k += kRingTile) metas[k] = __ldg(chunk_meta + first + (size_t)k * step);
    if (tid == 0)
        for (int s = 0; s < STAGES; s++) mbar_init(bar + s, 1);
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < STAGES && s < mine; s++) {
            const uint32_t bytes = (uint32_t)metas[s].y * NW * 8u;
            mbar_expect_tx(bar + s, bytes);
            tma_load_1d(wt + (size_t)s * kRingTile * NW, W + (size_t)metas[s].x * NW, bytes, bar + s);
        }
    // gather pipeline: index of tile k+2 and t of tile k+1 are requested while tile k is consumed
    int idx1 = 0;
    double tc0 = 0.0, tc1 = 0.0, tc2 = 0.0;
    if (mine > 0 && tid < metas[0].y) {
        const int i = obs_pt[metas[0].x + tid];
        const double2 ta = __ldg(reinterpret_cast<const double2*>(t_in) + 2 * (size_t)i);
        const double2 tb = __ldg(reinterpret_cast<const double2*>(t_in) + 2 * (size_t)i + 1);
        tc0 = ta.x; tc1 = ta.y; tc2 = tb.x;
    }
    if (mine > 1 && tid < metas[1].y) idx1 = obs_pt[metas[1].x + tid];
    for (int k = 0; k < mine; k++) {
        const int slot = k % STAGES;
        const int nob = metas[k].y;
        int idx2 = 0;
        if (k + 2 < mine && tid < metas[k + 2].y) idx2 = obs_pt[metas[k + 2].x + tid];
        double tn0 = 0.0, tn1 = 0.0, tn2 = 0.0;
        if (k + 1 < mine && tid < metas[k + 1].y) {
            const double2 ta = __ldg(reinterpret_cast<const double2*>(t_in) + 2 * (size_t)idx1);
            const double2 tb = __ldg(reinterpret_cast<const double2*>(t_in) + 2 * (size_t)idx1 + 1);
            tn0 = ta.x; tn1 = ta.y; tn2 = tb.x;
        }
        mbar_wait(bar + slot, (uint32_t)((k / STAGES) & 1));
        double acc[NPAD];
#pragma unroll
        for (int r = 0; r < NPAD; r++) acc[r] = 0.0;
        if (tid < nob) {
            double w[NW];
            load_block<NW>(wt + (size_t)slot * kRingTile * NW, tid, w);
#pragma unroll
            for (int r = 0; r < NA; r++) acc[r] = w[r] * tc0 + w[r + NA] * tc1 + w[r + 2 * NA] * tc2;
        }
        double* rd = red + (k & 1) * NWARP * NA;
        {
            const double tot = warp_reduce_many<NPAD>(acc, lane);
            const int own = warp_reduce_owner<NPAD>(lane);
            if ((lane & (32 / NPAD - 1)) == 0 && own < NA) rd[warp * NA + own] = tot;
        }
        __syncthreads();       // tile consumed by every thread, partials of all warps visible
        if (tid == 0 && k + STAGES < mine) {
            const int2 mt = metas[k + STAGES];
            const uint32_t bytes = (uint32_t)mt.y * NW * 8u;
            mbar_expect_tx(bar + slot, bytes);
            tma_load_1d(wt + (size_t)slot * kRingTile * NW, W + (size_t)mt.x * NW, bytes, bar + slot);
        }
        if (tid < NA) {
            double s = 0.0;
            const int nw = (nob + 31) >> 5;
            for (int w = 0; w < nw; w++) s += rd[w * NA + tid];
            part[(size_t)NA * (first + (size_t)k * step) + tid] = s;
        }
        tc0 = tn0; tc1 = tn1; tc2 = tn2;
        idx1 = idx2;
    }
}

template <int NA, int STAGES>
__global__ void __launch_bounds__(kRingTile)
k_sweep_pt_ring(int ntiles, const int4* __restrict__ ptile_meta /* (q0, nob, p0, npts) */, const int* __restrict__ pt_ptr,
                const int* __restrict__ pt_cam, const double* __restrict__ Wp, const double* __restrict__ Vinv,
                const double* __restrict__ p, const int* __restrict__ done, double* __restrict__ t_out)
{
    constexpr int NW = 3 * NA;
    extern __shared__ __align__(128) unsigned char smraw[];
    double* wt = reinterpret_cast<double*>(smraw);                       // STAGES x kRingTile x NW
    double* sv = wt + (size_t)STAGES * kRingTile * NW;                   // 2 x kRingTile x 3
    int4* metas = reinterpret_cast<int4*>(sv + 2 * kRingTile * 3);       // kRingMaxTiles
    uint64_t* bar = reinterpret_cast<uint64_t*>(metas + kRingMaxTiles);  // STAGES
    const int tid = threadIdx.x;
    if (done && *done) return;
    const int first = blockIdx.x, step = gridDim.x;
    const int mine = first < ntiles ? (ntiles - first + step - 1) / step : 0;
    for (int k = tid; k < mine; k += kRingTile) metas[k] = __ldg(ptile_meta + first + (size_t)k * step);
    if (tid == 0)
        for (int s = 0; s < STAGES; s++) mbar_init(bar + s, 1);
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < STAGES && s < mine; s++) {
            const uint32_t bytes = (uint32_t)metas[s].y * NW * 8u;
            mbar_expect_tx(bar + s, bytes);
            tma_load_1d(wt + (size_t)s * kRingTile * NW, Wp + (size_t)metas[s].x * NW, bytes, bar + s);
        }
    // per-tile operands, requested one tile ahead: p_j of the observation's camera (index two ahead),
    // and the point epilogue's track bounds and V*^-1
    double pc[NA], Vc[9];
    int oc0 = 0, oc1 = 0, cam1 = 0;
#pragma unroll
    for (int r = 0; r < NA; r++) pc[r] = 0.0;
    if (mine > 0) {
        const int4 m0 = metas[0];
        if (tid < m0.y) {
            const double* pp = p + (size_t)NA * pt_cam[m0.x + tid];
#pragma unroll
            for (int r = 0; r < NA; r++) pc[r] = __ldg(pp + r);
        }
        if (tid < m0.w) {
            oc0 = pt_ptr[m0.z + tid] - m0.x; oc1 = pt_ptr[m0.z + tid + 1] - m0.x;
#pragma unroll
            for (int q = 0; q < 9; q++) Vc[q] = __ldg(Vinv + (size_t)9 * (m0.z + tid) + q);
        }
    }
    if (mine > 1 && tid < metas[1].y) cam1 = pt_cam[metas[1].x + tid];
    for (int k = 0; k < mine; k++) {
        const int slot = k % STAGES;
        const int4 mk = metas[k];
        int cam2 = 0;
        if (k + 2 < mine && tid < metas[k + 2].y) cam2 = pt_cam[metas[k + 2].x + tid];
        double pn[NA], Vn[9];
        int on0 = 0, on1 = 0;
#pragma unroll
        for (int r = 0; r < NA; r++) pn[r] = 0.0;
        if (k + 1 < mine) {
            const int4 mn = metas[k + 1];
            if (tid < mn.y) {
                const double* pp = p + (size_t)NA * cam1;
#pragma unroll
                for (int r = 0; r < NA; r++) pn[r] = __ldg(pp + r);
            }
            if (tid < mn.w) {
                on0 = pt_ptr[mn.z + tid] - mn.x; on1 = pt_ptr[mn.z + tid + 1] - mn.x;
#pragma unroll
                for (int q = 0; q < 9; q++) Vn[q] = __ldg(Vinv + (size_t)9 * (mn.z + tid) + q);
            }
        }
        mbar_wait(bar + slot, (uint32_t)((k / STAGES) & 1));
        double* svk = sv + (k & 1) * kRingTile * 3;
        if (tid < mk.y) {
            double w[NW];
            load_block<NW>(wt + (size_t)slot * kRingTile * NW, tid, w);
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int r = 0; r < NA; r++) {
                s0 += w[r] * pc[r]; s1 += w[r + NA] * pc[r]; s2 += w[r + 2 * NA] * pc[r];
            }
            svk[tid * 3] = s0; svk[tid * 3 + 1] = s1; svk[tid * 3 + 2] = s2;
        }
        __syncthreads();       // tile consumed, per-observation products visible
        if (tid == 0 && k + STAGES < mine) {
            const int4 mt = metas[k + STAGES];
            const uint32_t bytes = (uint32_t)mt.y * NW * 8u;
            mbar_expect_tx(bar + slot, bytes);
            tma_load_1d(wt + (size_t)slot * kRingTile * NW, Wp + (size_t)mt.x * NW, bytes, bar + slot);
        }
        if (tid < mk.w) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            for (int o = oc0; o < oc1; o++) { s0 += svk[o * 3]; s1 += svk[o * 3 + 1]; s2 += svk[o * 3 + 2]; }
            double4 tv;
            tv.x = Vc[0] * s0 + Vc[3] * s1 + Vc[6] * s2;
            tv.y = Vc[1] * s0 + Vc[4] * s1 + Vc[7] * s2;
            tv.z = Vc[2] * s0 + Vc[5] * s1 + Vc[8] * s2;
            tv.w = 0.0;
            reinterpret_cast<double4*>(t_out)[mk.z + tid] = tv;
        }
#pragma unroll
        for (int r = 0; r < NA; r++) pc[r] = pn[r];
#pragma unroll
        for (int q = 0; q < 9; q++) Vc[q] = Vn[q];
        oc0 = on0; oc1 = on1; cam1 = cam2;
    }
}

}  // namespace vlgba

// =========================================================================================
// Explicit-S matvec for mid-sized camera counts.  When the assembled reduced system is smaller
// than the W stream (4 Np^2 bytes for the lower triangle vs 304 B/observation for the two
// implicit sweeps; Venice shape: 0.46 GB vs 1.52 GB) PCG runs on the dense block-assembled S of
// k_schur_blocks instead, and one iteration is a symmetric matrix-vector product that reads the
// LOWER TRIANGLE ONLY: every element a = S[r][c] read once feeds both y_r += a x_c and
// y_c += a x_r.
//   * a tile is 256 rows x 32 columns of S (column-major => 32 contiguous 2-KB runs), pulled
//     into shared memory by 32 TMA bulk copies on one mbarrier; a persistent CTA per SM walks its
//     contiguous piece of a host-built tile sequence with a 3-stage ring, i.e. up to 192 KB in
//     flight per SM;
//   * thread r & 255 owns row r: row sums are accumulated per fragment in shared memory, the 32
//     column sums stay in registers over a run of tiles of one strip ("Work decomposition" below);
//   * the partials are added in the order of host-built lists -> bit-reproducible, no atomics.
// =========================================================================================
namespace vlgba {

constexpr int kSymvCols = 32;
constexpr int kSymvRows = 256;    // rows per tile == threads per CTA
constexpr int kSymvStages = 3;
constexpr int kSymvSlab = 8;         // strips per cell
constexpr int kSymvBlkRows = 2048;   // rows per cell (absolute row grid); also the length of a fragment's row-sum vector

// tile descriptor: x = strip, y = first row, z = rows | flags << 16, w = fragment | strip-in-cell << 20.
// Tiles are aligned to the absolute 256-row grid (the first tile of a strip starts at its diagonal, 32 J, and is
// short): row r always belongs to thread r & 255.
constexpr int kSymvFirstStrip = 1 << 16, kSymvLastStrip = 1 << 17, kSymvFirstFrag = 1 << 18, kSymvLastFrag = 1 << 19;

// all 32 values of every lane summed over the warp; lane l ends up owning value l
__device__ __forceinline__ double warp_reduce_32(double* v, int lane)
{
    int width = 32;
#pragma unroll
    for (int mask = 16; mask >= 1; mask >>= 1, width >>= 1) {
        const bool hi = (lane & mask) != 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (k < width / 2) {
                const double send = hi ? v[k] : v[k + width / 2];
                const double keep = hi ? v[k + width / 2] : v[k];
                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
            }
        }
    }
    return v[0];
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void consumer_sync()      // named barrier 1: the 256 consumer threads only
{
    asm volatile("bar.sync 1, %0;" ::"n"(kSymvRows) : "memory");
}

// Work decomposition.  The lower triangle is cut into CELLS of kSymvSlab strips x kSymvBlkRows rows; the tiles
// of a cell are ordered strip by strip, the cells slab by slab, and every CTA gets a contiguous piece of that
// sequence (equal cost).  The piece of a cell that one CTA owns is a FRAGMENT, the unit of the partial results:
//   * row sums: thread r & 255 adds the row sum of every tile into its own slot of a kSymvBlkRows-long vector in
//     shared memory (no barrier: a slot has one owner) and the fragment leaves ONE vector of row partials;
//   * column sums: the 32 accumulators of a thread live over a run of tiles of one strip, are reduced at its end
//     (warp tree, then the 8 warps through shared memory) and leave as the fragment's partial for that strip.
// A row is then the sum of <= Np / (32 kSymvSlab) + a few fragment partials and a column of <= Np / kSymvBlkRows + a
// few (host-built lists; summed in list order => bit-reproducible, no atomics).  The first version wrote a row
// partial per tile (one per 32 columns): 14 MB per product at Venice shape and 355 terms per row to fold.
//
// warps 0..7 consume, warp 8 produces: lane c issues the bulk copy of column c, so a whole tile is requested by
// one warp instruction and the producer runs ahead of the consumers by kSymvStages tiles (full/empty mbarrier
// pairs, no CTA-wide barrier per tile).  Both halves are device functions so that the persistent PCG kernel
// below runs the same code; `kbase` is the number of tiles this CTA has already pushed through the ring (the
// mbarrier phases go on across calls), `L2X` makes x a coherent L2 read (x written by other CTAs of the kernel).
struct SymvSmem {
    double* st;        // kSymvStages x 32 x 256
    double* cred;      // 8 x 32
    double* xs;        // 32
    double* yacc;      // kSymvBlkRows
    uint64_t* full;    // kSymvStages
    uint64_t* empty;   // kSymvStages
};

__device__ __forceinline__ SymvSmem symv_smem(unsigned char* smraw)
{
    SymvSmem m;
    m.st = reinterpret_cast<double*>(smraw);
    m.cred = m.st + (size_t)kSymvStages * kSymvCols * kSymvRows;
    m.xs = m.cred + (kSymvRows / 32) * kSymvCols;
    m.yacc = m.xs + kSymvCols;
    m.full = reinterpret_cast<uint64_t*>(m.yacc + kSymvBlkRows);
    m.empty = m.full + kSymvStages;
    return m;
}

constexpr size_t kSymvSmemBytes =
    sizeof(double) * ((size_t)kSymvStages * kSymvCols * kSymvRows + (kSymvRows / 32) * kSymvCols + kSymvCols + kSymvBlkRows) + 16 * kSymvStages + 16;

// fold lists: the fragments whose row vector covers row block b (of kSymvBlkRows rows), and for strip J the offsets
// (fragment * 32 kSymvSlab + 32 strip-in-cell) of its column partials
struct SymvFold { const int* row_ptr; const int* row_list; const int* col_ptr; const int* col_list; };

__device__ __forceinline__ void symv_producer(const SymvSmem& sm, int ld, const double* __restrict__ S,
                                              const int4* __restrict__ tiles, int t0, int nt, int kbase, int lane)
{
    const uint64_t policy = l2_evict_first_policy();
    for (int k = 0; k < nt; k++) {
        const int g = kbase + k, slot = g % kSymvStages;
        const int4 d = __ldg(tiles + t0 + k);
        const uint32_t bytes = (uint32_t)(d.z & 0xffff) * 8u;
        if (g >= kSymvStages) mbar_wait(sm.empty + slot, (uint32_t)((g / kSymvStages - 1) & 1));
        if (lane == 0) mbar_expect_tx(sm.full + slot, bytes * kSymvCols);
        __syncwarp();
        tma_load_1d_pol(sm.st + ((size_t)slot * kSymvCols + lane) * kSymvRows + (d.y & (kSymvRows - 1)),
                        S + (size_t)ld * (kSymvCols * d.x + lane) + d.y, bytes, sm.full + slot, policy);
    }
}

// x_r: plain, or x_r + beta xp_r when xp is given (the persistent PCG kernel multiplies by p = z + beta p_old without
// waiting for anyone to store p)
template <bool L2X>
__device__ __forceinline__ double symv_ldx(const double* x, const double* xp, double beta, int r)
{
    if (!L2X) return __ldg(x + r);
    const double v = __ldcg(x + r);
    return xp ? fma(beta, __ldcg(xp + r), v) : v;
}

template <bool L2X>
__device__ __forceinline__ void symv_consumer(const SymvSmem& sm, int N, const double* __restrict__ x,
                                              const int4* __restrict__ tiles, int t0, int nt, int kbase,
                                              double* __restrict__ rowpart, double* __restrict__ colpart,
                                              const double* __restrict__ xp = nullptr, double beta = 0.0)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double colacc[kSymvCols];
#pragma unroll
    for (int c = 0; c < kSymvCols; c++) colacc[c] = 0.0;
    // x of this thread's row, requested one tile ahead
    int4 dn = __ldg(tiles + t0);
    double xr_next = 0.0;
    {
        const int r = (dn.y & ~(kSymvRows - 1)) + tid;
        if (r >= dn.y && r < dn.y + (dn.z & 0xffff) && r < N) xr_next = symv_ldx<L2X>(x, xp, beta, r);
    }
    for (int k = 0; k < nt; k++) {
        const int g = kbase + k, slot = g % kSymvStages;
        const int4 d = dn;
        const int rows = d.z & 0xffff, c0 = kSymvCols * d.x, r = (d.y & ~(kSymvRows - 1)) + tid;
        const bool live = r >= d.y && r < d.y + rows;
        double xr = xr_next;
        if (k + 1 < nt) {
            dn = __ldg(tiles + t0 + k + 1);
            const int rn = (dn.y & ~(kSymvRows - 1)) + tid;
            xr_next = (rn >= dn.y && rn < dn.y + (dn.z & 0xffff) && rn < N) ? symv_ldx<L2X>(x, xp, beta, rn) : 0.0;
        }
        if (d.z & kSymvFirstFrag) {                 // own slots only: no barrier
#pragma unroll
            for (int i = 0; i < kSymvBlkRows / kSymvRows; i++) sm.yacc[tid + kSymvRows * i] = 0.0;
        }
        if (d.z & kSymvFirstStrip) {                // first tile of a run in this strip: the strip's x
            if (tid < kSymvCols) sm.xs[tid] = (c0 + tid < N) ? symv_ldx<L2X>(x, xp, beta, c0 + tid) : 0.0;
            consumer_sync();
        }
        mbar_wait(sm.full + slot, (uint32_t)((g / kSymvStages) & 1));
        double ra = 0.0;
        if (live) {
            // the diagonal 32 x 32 block is stored in full: whole-row dot product there, no mirrored part
            if (r < c0 + kSymvCols) xr = 0.0;
            const double* a = sm.st + (size_t)slot * kSymvCols * kSymvRows + tid;
#pragma unroll
            for (int c = 0; c < kSymvCols; c++) {
                const double v = a[c * kSymvRows];
                ra += v * sm.xs[c];
                colacc[c] += v * xr;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.empty + slot);   // this warp is done with the slot
        if (live) sm.yacc[r & (kSymvBlkRows - 1)] += ra;
        const int frag = d.w & 0xfffff;
        if (d.z & kSymvLastStrip) {                 // last tile of the run: column sums out
            const double tot = warp_reduce_32(colacc, lane);
            sm.cred[warp * kSymvCols + lane] = tot;
            consumer_sync();
            if (tid < kSymvCols) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < kSymvRows / 32; w++) s += sm.cred[w * kSymvCols + tid];
                colpart[(size_t)frag * (kSymvCols * kSymvSlab) + kSymvCols * (d.w >> 20) + tid] = s;
            }
#pragma unroll
            for (int c = 0; c < kSymvCols; c++) colacc[c] = 0.0;
        }
        if (d.z & kSymvLastFrag) {                  // the fragment's row sums out (own slots only)
#pragma unroll
            for (int i = 0; i < kSymvBlkRows / kSymvRows; i++)
                rowpart[(size_t)frag * kSymvBlkRows + tid + kSymvRows * i] = sm.yacc[tid + kSymvRows * i];
        }
    }
}

__global__ void __launch_bounds__(kSymvRows + 32, 1)
k_symv_lower(int ld, int N, const double* __restrict__ S, const double* __restrict__ x,
             const int* __restrict__ tile_ptr /* [gridDim.x + 1] */, const int4* __restrict__ tiles,
             const int* __restrict__ done, double* __restrict__ rowpart /* [nfrag][kSymvBlkRows] */,
             double* __restrict__ colpart /* [nfrag][32 kSymvSlab] */)
{
    extern __shared__ __align__(128) unsigned char smraw[];
    const SymvSmem sm = symv_smem(smraw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (done && *done) return;
    const int t0 = tile_ptr[blockIdx.x], nt = tile_ptr[blockIdx.x + 1] - t0;
    if (nt <= 0) return;
    if (tid == 0)
        for (int s = 0; s < kSymvStages; s++) { mbar_init(sm.full + s, 1); mbar_init(sm.empty + s, kSymvRows / 32); }
    __syncthreads();
    if (warp == kSymvRows / 32) symv_producer(sm, ld, S, tiles, t0, nt, 0, lane);
    else symv_consumer<false>(sm, N, x, tiles, t0, nt, 0, rowpart, colpart);
}

// term q of the 32-row block rb (lane = row in the block): the fragments' row partials first, then the column partials
__device__ __forceinline__ const double* symv_term(const SymvFold& f, const double* rowpart, const double* colpart, int rb, int lane,
                                                   int rbeg, int nrow, int cbeg, int q)
{
    return q < nrow ? rowpart + (size_t)__ldg(f.row_list + rbeg + q) * kSymvBlkRows + ((rb & (kSymvBlkRows / 32 - 1)) << 5) + lane
                    : colpart + __ldg(f.col_list + cbeg + q - nrow) + lane;
}

struct P2PMail;
__device__ __forceinline__ double p2p_allreduce_elem(const P2PMail& mb, unsigned int epoch, int t, double x);

// With `mb` (multi-GPU) the row's value is exchanged with the peers right here: it is stored into every
// rank's mailbox and the sum over ranks comes back, see the P2PMail section at the end of this file.
__global__ void __launch_bounds__(1024)
k_symv_finish(int N, double sign, SymvFold f, const double* __restrict__ rowpart, const double* __restrict__ colpart,
              const int* __restrict__ done, double* __restrict__ out, const P2PMail* __restrict__ mb = nullptr,
              unsigned int epoch = 0)
{
    __shared__ double red[32][33];
    if (done && *done) return;
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int rb = blockIdx.x, r = rb * 32 + lane;
    const int b = rb / (kSymvBlkRows / 32);
    const int rbeg = f.row_ptr[b], nrow = f.row_ptr[b + 1] - rbeg, cbeg = f.col_ptr[rb], nterm = nrow + f.col_ptr[rb + 1] - cbeg;
    double s = 0.0;
    for (int q = g; q < nterm; q += 32) s += __ldcs(symv_term(f, rowpart, colpart, rb, lane, rbeg, nrow, cbeg, q));
    red[g][lane] = s;
    __syncthreads();
    if (g == 0 && r < N) {
        double y = 0.0;
#pragma unroll
        for (int q = 0; q < 32; q++) y += red[q][lane];
        y *= sign;
        if (mb) y = p2p_allreduce_elem(*mb, epoch, r, y);
        out[r] = y;
    }
}

}  // namespace vlgba

// =========================================================================================
// All-reduce of the per-iteration PCG vector over NVLink peer memory (multi-GPU, one process per
// GPU).  The vector is small (6m doubles: 85 kB at Venice shape) and sits on the critical path of
// every iteration, so it is latency that counts, not bandwidth.  Each rank owns a MAILBOX
// (cudaMalloc'ed, opened by every peer through CUDA IPC); an element travels as one 16-byte store
// {lo, epoch, hi, epoch} into slot [parity][my rank][t] of EVERY rank's mailbox -- the flag rides
// with the data (each 8-byte half carries its own copy, so the protocol does not depend on 16-byte
// store atomicity), hence no fence and no separate flag round trip -- and the receiving thread polls
// its own element of every rank's slot and adds them in rank order: every rank gets bit-identical
// sums.  No barrier of any kind; a thread always publishes before it polls, so there is no
// deadlock.  Slots are double-buffered by epoch parity: a peer can only send epoch e+2 after it
// consumed my epoch e+1, which I send after my epoch-e kernel has finished.
// The exchange is a device function so that it runs INSIDE the kernel that produces the vector
// (k_symv_finish: transfer of one row block overlaps the partial sums of the others).
// =========================================================================================
namespace vlgba {

constexpr int kP2pMaxRanks = 16;
constexpr int kP2pThreads = 256;

struct P2PMail {
    uint4* data[kP2pMaxRanks];           // data[r]: rank r's mailbox, [2][nranks][nslot] packets
    int nranks, rank, nslot;
};

__device__ __forceinline__ double p2p_allreduce_elem(const P2PMail& mb, unsigned int epoch, int t, double x)
{
    const int par = (int)(epoch & 1u);
    const size_t mine = ((size_t)par * mb.nranks + mb.rank) * mb.nslot + t;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    const uint4 pkt = make_uint4((unsigned int)bits, epoch, (unsigned int)(bits >> 32), epoch);
    for (int r = 0; r < mb.nranks; r++) {
        uint4* dst = mb.data[r] + mine;
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(pkt.x), "r"(pkt.y), "r"(pkt.z), "r"(pkt.w) : "memory");
    }
    double s = 0.0;
    for (int r = 0; r < mb.nranks; r++) {
        const uint4* src = mb.data[mb.rank] + ((size_t)par * mb.nranks + r) * mb.nslot + t;
        uint4 v;
        do {
            asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
        } while (v.y != epoch || v.w != epoch);
        s += __longlong_as_double((long long)(((unsigned long long)v.z << 32) | v.x));
    }
    return s;
}

// stand-alone form (implicit-Schur path: after k_cam_sum_partials)
__global__ void __launch_bounds__(kP2pThreads)
k_p2p_allreduce(P2PMail mb, int n, unsigned int epoch, const int* __restrict__ done, double* __restrict__ v /* in: partial, out: sum */)
{
    if (done && *done) return;            // identical on every rank (the stop flag comes from all-reduced values)
    const int t = blockIdx.x * kP2pThreads + threadIdx.x;
    if (t < n) v[t] = p2p_allreduce_elem(mb, epoch, t, v[t]);
}

}  // namespace vlgba

// =========================================================================================
// The whole PCG solve on the assembled S as ONE persistent cooperative kernel.
// Launched per iteration, the loop is three kernels (matvec 75 us, fold of its partials 10 us,
// vector update 22 us with two grid syncs) plus their launch gaps and a host check of the stop
// flag every 8 iterations.  Here one CTA per SM stays resident for the whole solve and an
// iteration is four phases, each closed by the light grid barrier of ba_chol.cuh:
//   1. matvec tiles through the TMA ring (symv_producer / symv_consumer, same code as k_symv_lower),
//      multiplying by p = z + beta p_old formed on the fly from the two vectors of the previous
//      iteration (the owners store p for the later phases, nobody waits for it);
//   2. fold of the fragment partials into wq (CTA c folds the 32-row blocks c, c + G, c + 2G at once,
//      term offsets staged in shared memory once per solve; the NVLink mailbox exchange of multi-GPU
//      runs happens here, per element); meanwhile the TMA engine pulls this CTA's cluster inverse M1
//      and the head of the composite M2 into the (now idle) ring memory;
//   3. q = U* p - wq and the dot products p'q, Z'q, SZ'p (block partials; halo threads fetch the
//      operands of the neighbouring cameras' residual);
//   4. alpha (scalar partials folded by one warp per scalar), x, r -- own and halo --,
//      z = (M1^-1 + M2^-1) r from shared memory, r'z and r'r (block partials); after the barrier
//      beta, the owners' p, the stop flag.
// Clusters (update CTAs) are the first `nclusters` CTAs; every CTA folds the scalar partials so
// that alpha, beta and the stop decision are uniform without a broadcast.
// =========================================================================================
namespace vlgba {

struct PcgPersistArgs {
    int Np, ld, N, m, max_iter, nclusters;
    double rtol;
    const double* S;
    const int* tile_ptr;
    const int4* tiles;
    double *rowpart, *colpart, *wq;
    const double *Ud, *Minv, *McL;
    const double* Mc2;               // composite rows of the shifted partition's inverses (overlapping preconditioner) or NULL
    double *x, *r, *p;
    double *p2, *zbuf;               // second p buffer and z: iteration k multiplies by z + beta p_old read from these (no barrier for p)
    PcgScalars* sc;
    double* blkpart;                 // 11 * nclusters
    const double *Z, *SZ;            // deflation vectors or NULL
    const DeflScalars* ds;
    unsigned int* barrier;           // zero at launch
    const P2PMail* mb;               // multi-GPU exchange or NULL
    unsigned int epoch0;             // mailbox epoch of the first matvec of this launch
    long long* prof;                 // optional: clock64 totals per phase seen by CTA 0 / thread 0 (VLG_BA_PERSIST_PROF), or NULL
    SymvFold fold;                   // fragment lists of the row / column partials
    long long* stat;                 // optional [2 G]: matvec clocks (accumulated) and SM id per CTA (opts.pcg_autotune) or NULL
    int mcl_evict_first;             // stage the cluster inverses with an L2 evict-first hint (VLG_BA_MCL_EVICT overrides)
};

#define PCG_PROF(slot)                                                           \
    do {                                                                         \
        if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) {                     \
            const long long now_ = clock64();                                    \
            a.prof[slot] += now_ - prof_t;                                       \
            prof_t = now_;                                                       \
        }                                                                        \
    } while (0)

template <int NA>
__global__ void __launch_bounds__(kSymvRows + 32, 1)
k_pcg_persistent(PcgPersistArgs a)
{
    constexpr int ND = 1 + 2 * kDefl;
    static_assert(ND <= (kSymvRows + 32) / 32, "one warp per scalar partial");
    using CL = Cluster<NA>;
    constexpr int kCams = 128 / NA, NC = kCams * NA;
    // overlapping preconditioner: the ring memory holds 192 columns of 128 doubles -- M1 (NC columns) and the first CA
    // columns of the composite M2; its last CB columns replace M1's first CB once those have been used
    constexpr int kRingCols = kSymvStages * kSymvCols * kSymvRows / 128, CA = kRingCols - NC, CB = NC - CA;
    constexpr int kOwn0 = CL::kShift * NA;        // rx: r of the cameras [cta kCams - kShift, cta kCams + kCams + kHalf), own part at kOwn0
    static_assert(CA > 0 && CB > 0 && CB <= NC && 2 * NC <= 256, "ring window of the overlapping preconditioner");
    extern __shared__ __align__(128) unsigned char smraw[];
    const SymvSmem sm = symv_smem(smraw);
    __shared__ double shd[4 * ND], pv[128], pv2[128], rx[256], z2s[128], fold[3][8][33], bcast[ND + 2];
    double* const rv = rx + kOwn0;
    __shared__ uint64_t mbarM, mbarM2;
    __shared__ int fterm[3][64], fmeta[3][4];      // fold terms of this CTA's first three row blocks (constant over the solve)
    __shared__ double einv[kDefl * kDefl];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, G = gridDim.x, nb = a.nclusters;
    const bool producer = warp == kSymvRows / 32;
    const bool cluster_cta = cta < nb;
    const int t0 = a.tile_ptr[cta], nt = a.tile_ptr[cta + 1] - t0;
    const size_t N = (size_t)a.N;
    if (tid == 0) {
        for (int s = 0; s < kSymvStages; s++) { mbar_init(sm.full + s, 1); mbar_init(sm.empty + s, kSymvRows / 32); }
        mbar_init(&mbarM, 1); mbar_init(&mbarM2, 1);
    }
    if (a.Z && tid < kDefl * kDefl) einv[tid] = a.ds->Einv[tid];     // constant over the solve
    if (a.stat && tid == 0) {
        unsigned int smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        a.stat[2 * blockIdx.x + 1] = (long long)smid;
    }
    // the fold lists of the CTA's first three 32-row blocks, resolved to offsets once: row partials as offsets into rowpart,
    // column partials as offsets into colpart with the sign bit set (an acquire at every grid barrier empties L1, so reading
    // pointer -> list -> data from global was three dependent L2 round trips per iteration)
    for (int u3 = 0; u3 < 3; u3++) {
        const int rb = u3 * (int)gridDim.x + (int)blockIdx.x;
        int rbeg = 0, nrow = 0, cbeg = 0, nterm = 0;
        if (rb < a.Np / 32) {
            const int b = rb / (kSymvBlkRows / 32);
            rbeg = a.fold.row_ptr[b]; nrow = a.fold.row_ptr[b + 1] - rbeg;
            cbeg = a.fold.col_ptr[rb]; nterm = nrow + a.fold.col_ptr[rb + 1] - cbeg;
        }
        if (tid < 64 && tid < nterm)
            fterm[u3][tid] = tid < nrow ? a.fold.row_list[rbeg + tid] * kSymvBlkRows + ((rb & (kSymvBlkRows / 32 - 1)) << 5)
                                        : (int)(0x80000000u | (unsigned int)a.fold.col_list[cbeg + tid - nrow]);
        if (tid == 0) { fmeta[u3][0] = rbeg; fmeta[u3][1] = nrow; fmeta[u3][2] = cbeg; fmeta[u3][3] = nterm; }
    }
    __syncthreads();
    unsigned int bar_target = 0;
    int kbase = 0;
    const double r0n2 = a.sc->r0n2;
    double rz = a.sc->rz;
    // p of iteration k lives in buffer k & 1 (a.p, a.p2), stored by its owner once beta is known and read by the halo
    // threads two barriers later; the matvec does not wait for it: it forms z + beta p_old itself
    double beta_prev = 0.0, pt_next = 0.0;
    const int lc = tid / NA, row = tid - lc * NA;             // update role: local camera, row (threads < 128)
    const int j = cta * kCams + lc;
    const bool act = cluster_cta && tid < 128 && lc < kCams && j < a.m;
    const size_t t = (size_t)NA * j + row;
    const int nrb = a.Np / 32;
    // halo role (threads 128..255 of a cluster CTA, overlapping preconditioner): r of the neighbouring cameras is
    // recomputed here with the owner's arithmetic instead of being fetched after one more grid barrier
    const bool ovl = cluster_cta && a.McL && a.Mc2;
    const int h = tid - 128;
    const int eh = h < kOwn0 ? h : NC + h;
    const long long gh = (long long)NA * ((long long)cta * kCams - CL::kShift) + eh;
    const bool hrole = ovl && tid >= 128 && tid < 256;
    const bool hact = hrole && h < NC && gh >= 0 && gh < (long long)N;
    const int jh = hact ? (int)(gh / NA) : 0, rowh = hact ? (int)(gh - (long long)NA * jh) : 0, lch = (h & 127) / NA;

    long long prof_t = a.prof ? clock64() : 0;
    for (int it = 0; it < a.max_iter; it++) {
        // ---- 1. matvec tiles
        const long long mv_t0 = (a.prof || a.stat) ? clock64() : 0;
        if (nt > 0) {
            if (producer) symv_producer(sm, a.ld, a.S, a.tiles, t0, nt, kbase, lane);
            else if (it == 0) symv_consumer<true>(sm, a.N, a.p, a.tiles, t0, nt, kbase, a.rowpart, a.colpart);
            else symv_consumer<true>(sm, a.N, a.zbuf, a.tiles, t0, nt, kbase, a.rowpart, a.colpart, (it & 1) ? a.p : a.p2, beta_prev);
            kbase += nt;
        }
        if (tid == 0 && (a.prof || a.stat)) {                               // per-CTA matvec time (load balance of the cut)
            const long long dt = clock64() - mv_t0;
            if (a.prof) a.prof[32 + cta] += dt;
            if (a.stat) a.stat[2 * cta] += dt;
        }
        PCG_PROF(0);
        grid_barrier(a.barrier, bar_target, G);
        PCG_PROF(1);
        // ---- 2. cluster inverse -> ring memory (asynchronously), fold of the matvec partials -> wq
        if (cluster_cta && a.McL && tid == 0) {
            constexpr uint32_t bytes = (uint32_t)NC * 128u * 8u, bytes2 = (uint32_t)CA * 128u * 8u;
            mbar_expect_tx(&mbarM, ovl ? bytes + bytes2 : bytes);
            if (a.mcl_evict_first) {
                const uint64_t pol = l2_evict_first_policy();
                tma_load_1d_pol(sm.st, a.McL + (size_t)cta * 128 * 128, bytes, &mbarM, pol);
                if (ovl) tma_load_1d_pol(sm.st + (size_t)NC * 128, a.Mc2 + (size_t)cta * 128 * 128, bytes2, &mbarM, pol);
            } else {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(sm.st)), "l"(a.McL + (size_t)cta * 128 * 128), "r"(bytes), "r"(smem_u32(&mbarM)) : "memory");
                if (ovl)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     smem_u32(sm.st + (size_t)NC * 128)), "l"(a.Mc2 + (size_t)cta * 128 * 128), "r"(bytes2), "r"(smem_u32(&mbarM)) : "memory");
            }
        }
        PCG_PROF(16);
        if (!producer) {
            const int gq = tid >> 5;                               // term group 0..7
            // 32-row blocks are dealt round-robin to the CTAs, three rounds folded TOGETHER: per block the 256 threads
            // are 32 rows x 8 groups, group gq takes terms gq, gq + 8, ... of the block's fragment lists; the loads of
            // all three blocks are issued before the first add and three different warps finish the three blocks.
            for (int qb = 0; qb * G < nrb; qb += 3) {
                int rbv[3], rbeg[3], nrowv[3], cbeg[3], ntermv[3];
                int maxterm = 0;
#pragma unroll
                for (int u3 = 0; u3 < 3; u3++) {
                    const int rb = (qb + u3) * G + cta;
                    rbv[u3] = rb;
                    ntermv[u3] = 0; rbeg[u3] = 0; nrowv[u3] = 0; cbeg[u3] = 0;
                    if (rb < nrb) {                                 // uniform over the CTA
                        if (qb == 0) {
                            rbeg[u3] = fmeta[u3][0]; nrowv[u3] = fmeta[u3][1]; cbeg[u3] = fmeta[u3][2]; ntermv[u3] = fmeta[u3][3];
                        } else {
                            const int b = rb / (kSymvBlkRows / 32);
                            rbeg[u3] = __ldg(a.fold.row_ptr + b); nrowv[u3] = __ldg(a.fold.row_ptr + b + 1) - rbeg[u3];
                            cbeg[u3] = __ldg(a.fold.col_ptr + rb); ntermv[u3] = nrowv[u3] + __ldg(a.fold.col_ptr + rb + 1) - cbeg[u3];
                        }
                    }
                    maxterm = max(maxterm, ntermv[u3]);
                }
                double acc[3] = {0.0, 0.0, 0.0};
                PCG_PROF(17);
                for (int q0 = gq; q0 < maxterm + gq; q0 += 64) {    // bound uniform over the CTA
                    double v[3][8];
                    if (qb == 0 && q0 < 64) {
                        // the common case: offsets from shared memory, one L2 round trip for all the partials
#pragma unroll
                        for (int u3 = 0; u3 < 3; u3++)
#pragma unroll
                            for (int u = 0; u < 8; u++) {
                                const int tq = q0 + 8 * u;
                                const int off = fterm[u3][tq];
                                const double* src = off >= 0 ? a.rowpart + off + lane : a.colpart + (off & 0x7fffffff) + lane;
                                v[u3][u] = tq < ntermv[u3] ? __ldcg(src) : 0.0;
                            }
                    } else {
#pragma unroll
                        for (int u3 = 0; u3 < 3; u3++)
#pragma unroll
                            for (int u = 0; u < 8; u++) {
                                const int tq = q0 + 8 * u;
                                v[u3][u] = tq < ntermv[u3]
                                               ? __ldcg(symv_term(a.fold, a.rowpart, a.colpart, rbv[u3], lane, rbeg[u3], nrowv[u3], cbeg[u3], tq))
                                               : 0.0;
                            }
                    }
#pragma unroll
                    for (int u3 = 0; u3 < 3; u3++)
#pragma unroll
                        for (int u = 0; u < 8; u++) acc[u3] += v[u3][u];
                }
                PCG_PROF(15);
                if (qb > 0) consumer_sync();                        // the previous group's sums have been read
#pragma unroll
                for (int u3 = 0; u3 < 3; u3++) fold[u3][gq][lane] = acc[u3];
                consumer_sync();
                if (gq < 3) {
                    const int rb = gq == 0 ? rbv[0] : gq == 1 ? rbv[1] : rbv[2];
                    const int r = rb * 32 + lane;
                    if (rb < nrb && r < a.N) {
                        double y = 0.0;
#pragma unroll
                        for (int g8 = 0; g8 < 8; g8++) y += fold[gq][g8][lane];
                        y = -y;
                        if (a.mb) y = p2p_allreduce_elem(*a.mb, a.epoch0 + (unsigned int)it, r, y);
                        a.wq[r] = y;
                    }
                }
            }
        }
        PCG_PROF(2);
        grid_barrier(a.barrier, bar_target, G);
        PCG_PROF(3);
        // ---- 3. q = U* p - wq, dot products
        double dots[ND];
#pragma unroll
        for (int k = 0; k < ND; k++) dots[k] = 0.0;
        double pt = 0.0, qt = 0.0;
        double xt = 0.0, rt0 = 0.0, szv[kDefl];            // phase-4 operands, requested a barrier early
#pragma unroll
        for (int d = 0; d < kDefl; d++) szv[d] = 0.0;
        const int nd = a.Z ? ND : 1;
        if (cluster_cta && tid < 128) {
            if (act) {
                pt = it == 0 ? __ldcg(a.p + t) : pt_next; pv[tid] = pt;
                xt = __ldcg(a.x + t); rt0 = __ldcg(a.r + t);
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (act) {
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < NA; c++) v += a.Ud[(size_t)NA * NA * j + row + NA * c] * pv[lc * NA + c];
                qt = v - __ldcg(a.wq + t);
                dots[0] = pt * qt;
                if (a.Z) {
#pragma unroll
                    for (int d = 0; d < kDefl; d++) {
                        szv[d] = a.SZ[d * N + t];
                        dots[1 + d] = a.Z[d * N + t] * qt;
                        dots[1 + kDefl + d] = szv[d] * pt;
                    }
                }
            }
            for (int k = 0; k < nd; k++) {
                const double v = warp_sum(dots[k]);
                if (lane == 0) shd[warp * ND + k] = v;
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (tid < nd) {
                double v = 0.0;
                for (int w = 0; w < 4; w++) v += shd[w * ND + tid];
                a.blkpart[(size_t)tid * nb + cta] = v;
            }
        }
        double qh = 0.0, rh0 = 0.0, szh[kDefl];
#pragma unroll
        for (int d = 0; d < kDefl; d++) szh[d] = 0.0;
        if (hrole) {
            if (hact) { pv2[h] = __ldcg(((it & 1) ? a.p2 : a.p) + gh); rh0 = __ldcg(a.r + gh); }
            asm volatile("bar.sync 3, 128;" ::: "memory");
            if (hact) {
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < NA; c++) v += a.Ud[(size_t)NA * NA * jh + rowh + NA * c] * pv2[lch * NA + c];
                qh = v - __ldcg(a.wq + gh);
                if (a.Z) {
#pragma unroll
                    for (int d = 0; d < kDefl; d++) szh[d] = a.SZ[d * N + gh];
                }
            }
        }
        PCG_PROF(4);
        grid_barrier(a.barrier, bar_target, G);
        PCG_PROF(5);
        // every warp of every CTA folds the block partials the same way: uniform alpha / stop decisions
        // all the partial loads are requested before the first use (a loop over k with a runtime bound
        // serialises one L2 round trip per scalar: measured 10 us of an iteration)
        // ONE warp per CTA reads the partials (148 x 9 warps hammering the same 6 KB of L2 cost 9 us per
        // iteration) and publishes the sums through shared memory
        // every warp of every CTA would read the same 6 KB of L2 (148 x 9 warps: 9 us per iteration, measured); instead
        // warp k folds scalar k (ND = 9 scalars, 9 warps) -- all its loads before the first add -- and publishes the sum
        // through shared memory
        if (warp < nd) {
            constexpr int KB = 5;                      // nb <= 160 clusters covered in one pass
            double ldv[KB];
#pragma unroll
            for (int u = 0; u < KB; u++) {
                const int b = lane + 32 * u;
                ldv[u] = b < nb ? __ldcg(a.blkpart + (size_t)warp * nb + b) : 0.0;
            }
            double v = 0.0;
#pragma unroll
            for (int u = 0; u < KB; u++) v += ldv[u];
            for (int b = lane + 32 * KB; b < nb; b += 32) v += __ldcg(a.blkpart + (size_t)warp * nb + b);
            v = warp_sum(v);
            if (lane == 0) bcast[warp] = v;
        } else if (warp < ND && lane == 0) {
            bcast[warp] = 0.0;
        }
        PCG_PROF(13);
        __syncthreads();
        PCG_PROF(14);
#pragma unroll
        for (int k = 0; k < ND; k++) dots[k] = bcast[k];
        double y[kDefl];
        double pq = dots[0];
        if (a.Z) {
#pragma unroll
            for (int d = 0; d < kDefl; d++) {
                double v = 0.0;
#pragma unroll
                for (int e = 0; e < kDefl; e++) v += einv[d + kDefl * e] * dots[1 + e];
                y[d] = v;
                pq -= dots[1 + kDefl + d] * v;
            }
        }
        if (!(pq > 0.0)) {                            // breakdown: uniform over the grid
            if (cta == 0 && tid == 0) { a.sc->done = 2; a.sc->pq = pq; a.sc->rz = rz; a.sc->iters += it; a.sc->exchanges = it + 1; }
            // the bulk copy of this iteration's cluster inverse is still on its way into this CTA's shared memory:
            // let it land before the CTA (and its shared memory) goes away
            if (cluster_cta && a.McL && tid == 0) mbar_wait(&mbarM, (uint32_t)(it & 1));
            __syncthreads();
            return;
        }
        const double alpha = rz / pq;
        // ---- 4. x, r, z = M^-1 r
        double rzn = 0.0, rr = 0.0, zt = 0.0;
        double* bp2 = a.blkpart + (size_t)ND * nb;
        if (cluster_cta && tid < 256) {
            if (tid < 128) {
                double rt = 0.0;
                if (act) {
                    if (a.Z) {
#pragma unroll
                        for (int d = 0; d < kDefl; d++) qt -= szv[d] * y[d];
                    }
                    a.x[t] = xt + alpha * pt;
                    rt = rt0 - alpha * qt;
                    a.r[t] = rt;
                    rr = rt * rt;
                }
                if (tid < NC) rv[tid] = rt;
            } else if (hrole) {
                double rh = 0.0;
                if (hact) {
                    if (a.Z) {
#pragma unroll
                        for (int d = 0; d < kDefl; d++) qh -= szh[d] * y[d];
                    }
                    rh = rh0 - alpha * qh;
                }
                if (h < NC) rx[eh] = rh;
            }
            if (ovl) consumer_sync();
            else if (tid < 128) asm volatile("bar.sync 2, 128;" ::: "memory");
            PCG_PROF(10);
            if (ovl) {
                // z = M1^-1 r + M2^-1 r: threads 0..127 take the first term, threads 128..255 the second (composite rows)
                if (tid < 128) {
                    mbar_wait(&mbarM, (uint32_t)(it & 1));
                    PCG_PROF(11);
                    const double* M = sm.st + tid;
                    double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
                    if (tid < NC) {
#pragma unroll 4
                        for (int c = 0; c + 3 < CB; c += 4) {
                            z0 += M[128 * c] * rv[c]; z1 += M[128 * (c + 1)] * rv[c + 1];
                            z2 += M[128 * (c + 2)] * rv[c + 2]; z3 += M[128 * (c + 3)] * rv[c + 3];
                        }
#pragma unroll
                        for (int c = CB - CB % 4; c < CB; c++) z0 += M[128 * c] * rv[c];
                    }
                    // M1's first CB columns are used up: the tail of M2 takes their place
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                    if (tid == 0) {
                        constexpr uint32_t bytes3 = (uint32_t)CB * 128u * 8u;
                        mbar_expect_tx(&mbarM2, bytes3);
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                         smem_u32(sm.st)), "l"(a.Mc2 + ((size_t)cta * 128 + CA) * 128), "r"(bytes3), "r"(smem_u32(&mbarM2)) : "memory");
                    }
                    if (tid < NC) {
#pragma unroll 4
                        for (int c = CB; c + 3 < NC; c += 4) {
                            z0 += M[128 * c] * rv[c]; z1 += M[128 * (c + 1)] * rv[c + 1];
                            z2 += M[128 * (c + 2)] * rv[c + 2]; z3 += M[128 * (c + 3)] * rv[c + 3];
                        }
#pragma unroll
                        for (int c = CB + (NC - CB) / 4 * 4; c < NC; c++) z0 += M[128 * c] * rv[c];
                    }
                    zt = (z0 + z1) + (z2 + z3);
                } else {
                    mbar_wait(&mbarM, (uint32_t)(it & 1));
                    const double* rb = rx + (h < CL::kHalf * NA ? 0 : NC);
                    double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
                    if (h < NC) {
                        const double* M = sm.st + (size_t)NC * 128 + h;
#pragma unroll 4
                        for (int c = 0; c + 3 < CA; c += 4) {
                            z0 += M[128 * c] * rb[c]; z1 += M[128 * (c + 1)] * rb[c + 1];
                            z2 += M[128 * (c + 2)] * rb[c + 2]; z3 += M[128 * (c + 3)] * rb[c + 3];
                        }
#pragma unroll
                        for (int c = CA - CA % 4; c < CA; c++) z0 += M[128 * c] * rb[c];
                    }
                    mbar_wait(&mbarM2, (uint32_t)(it & 1));
                    if (h < NC) {
                        const double* M = sm.st + h;
                        const double* rc = rb + CA;
#pragma unroll 4
                        for (int c = 0; c + 3 < CB; c += 4) {
                            z0 += M[128 * c] * rc[c]; z1 += M[128 * (c + 1)] * rc[c + 1];
                            z2 += M[128 * (c + 2)] * rc[c + 2]; z3 += M[128 * (c + 3)] * rc[c + 3];
                        }
#pragma unroll
                        for (int c = CB - CB % 4; c < CB; c++) z0 += M[128 * c] * rc[c];
                    }
                    z2s[h] = (z0 + z1) + (z2 + z3);
                }
                consumer_sync();
                if (act) { zt += z2s[tid]; rzn = rv[tid] * zt; a.zbuf[t] = zt; }
            } else if (act) {
                if (a.McL) {
                    mbar_wait(&mbarM, (uint32_t)(it & 1));
                    PCG_PROF(11);
                    const double* M = sm.st + tid;                  // column-major, ld 128
                    double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
#pragma unroll 4
                    for (int c = 0; c + 3 < NC; c += 4) {
                        z0 += M[128 * c] * rv[c]; z1 += M[128 * (c + 1)] * rv[c + 1];
                        z2 += M[128 * (c + 2)] * rv[c + 2]; z3 += M[128 * (c + 3)] * rv[c + 3];
                    }
#pragma unroll
                    for (int c = NC - NC % 4; c < NC; c++) z0 += M[128 * c] * rv[c];
                    zt = (z0 + z1) + (z2 + z3);
                } else {
#pragma unroll
                    for (int c = 0; c < NA; c++) zt += a.Minv[(size_t)NA * NA * j + row + NA * c] * rv[lc * NA + c];
                }
                rzn = rv[tid] * zt;
                a.zbuf[t] = zt;
            }
            PCG_PROF(12);
            if (tid < 128) {
                rzn = warp_sum(rzn);
                rr = warp_sum(rr);
                if (lane == 0) { shd[warp * 2] = rzn; shd[warp * 2 + 1] = rr; }
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (tid < 2) {
                    double v = 0.0;
                    for (int w = 0; w < 4; w++) v += shd[w * 2 + tid];
                    bp2[2 * cta + tid] = v;
                }
            }
        }
        PCG_PROF(6);
        grid_barrier(a.barrier, bar_target, G);
        PCG_PROF(7);
        if (warp == 0) {
            constexpr int KB = 5;
            double l0[KB], l1[KB];
#pragma unroll
            for (int u = 0; u < KB; u++) {
                const int b = lane + 32 * u;
                l0[u] = b < nb ? __ldcg(bp2 + 2 * b) : 0.0;
                l1[u] = b < nb ? __ldcg(bp2 + 2 * b + 1) : 0.0;
            }
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int u = 0; u < KB; u++) { v0 += l0[u]; v1 += l1[u]; }
            for (int b = lane + 32 * KB; b < nb; b += 32) { v0 += __ldcg(bp2 + 2 * b); v1 += __ldcg(bp2 + 2 * b + 1); }
            v0 = warp_sum(v0); v1 = warp_sum(v1);
            if (lane == 0) { bcast[ND] = v0; bcast[ND + 1] = v1; }
        }
        __syncthreads();
        rzn = bcast[ND]; rr = bcast[ND + 1];
        // ---- 5. p = z + beta p, stop test
        const double beta = rzn / rz;
        if (act) { pt_next = fma(beta, pt, zt); (((it + 1) & 1) ? a.p2 : a.p)[t] = pt_next; }
        rz = rzn; beta_prev = beta;
        const bool stop = rr <= a.rtol * a.rtol * r0n2 || it + 1 == a.max_iter;
        if (stop) {
            if (cta == 0 && tid == 0) {
                a.sc->rz = rzn; a.sc->rn2 = rr; a.sc->pq = pq; a.sc->iters += it + 1; a.sc->exchanges = it + 1;
                if (rr <= a.rtol * a.rtol * r0n2) a.sc->done = 1;
            }
            return;
        }
        PCG_PROF(8);
    }
}

// ---- multi-GPU: sum of the ranks' shares of S, column block by column block, over peer memory -----------
// Every rank assembles its share of S (its points) in full.  Rank r then PULLS the column block it will
// multiply (strips [J0, J1)) from every peer's S through NVLink (CUDA-IPC mapped pointers) and adds the
// shares in rank order -- only the rows the symmetric matvec reads (row >= 32 J, i.e. the lower triangle
// and the full diagonal tile).  One thread block per column, 16-byte loads along the column.  (A grouped
// ncclReduce per block did the same in 1.1 ms at 2 GPUs but serialised to ~24 ms at 8.)
struct PeerS { const double* S[kP2pMaxRanks]; int nranks, rank; };

__global__ void __launch_bounds__(256)
k_pull_reduce_block(PeerS ps, int Np, int J0, double* __restrict__ Sown)
{
    const int col = kSymvCols * J0 + blockIdx.x;            // global column
    const int r0 = kSymvCols * (col / kSymvCols);           // first row the matvec reads in this column
    const size_t off = (size_t)Np * col;
    for (int r = r0 + 2 * threadIdx.x; r < Np; r += 2 * blockDim.x) {
        double2 acc = make_double2(0.0, 0.0);
        for (int q = 0; q < ps.nranks; q++) {
            const double2 v = q == ps.rank ? *reinterpret_cast<const double2*>(Sown + off + r)
                                           : __ldcs(reinterpret_cast<const double2*>(ps.S[q] + off + r));
            acc.x += v.x; acc.y += v.y;
        }
        *reinterpret_cast<double2*>(Sown + off + r) = acc;
    }
}

}  // namespace vlgba
