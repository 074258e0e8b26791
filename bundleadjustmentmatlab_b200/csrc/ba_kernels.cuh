// ba_kernels.cuh -- CUDA kernels (sm_100a, FP64) of the LM bundle-adjustment hot path.
//
// Data layout in HBM.  Observations are kept in the reference's own
// traversal order (ascending i + n*j, camera-major: mex_bundle_1_XABeUVWeAeB.c:192-196),
// called C-order: a camera's observations are one contiguous segment, so everything that
// depends on the camera (rotation tables, a_j, K_j) is warp-uniform in the heavy kernels.
// A second, point-major view (P-order: pt_ptr/pt_obs/pt_cam/pt_xy) serves the point-keyed
// sums; pt_obs maps a P-order slot to its C-order position.
//   W      [nobs][3][NA]   C-order, column-major NA x 3 block per observation
//   U,Ud   [m][NA][NA]     eA,e_,da [m][NA]     V,Vinv [n][3][3]   eB,db,t [n][3]
//   rtab   [m][4][9]       rotation matrices: base, w0+h, w1+h, w2+h
#pragma once
#include "ba_math.cuh"
#include <stdint.h>

namespace vlgba {

constexpr int kWarpsPerBlock = 4;

__host__ __device__ constexpr int nu_of(int na) { return na * (na + 1) / 2 + na; }

struct Stage1Args {
    int m, n, nchunks;
    const double2* obs_xy;      // C-order
    const int* obs_pt;          // C-order
    const int* chunk_cam;
    const int* chunk_begin;
    const int* chunk_end;
    const double* K4;
    const double* a;
    const double* b;
    const double* rtab;
    const unsigned char* cam_fixed;
    int fix_structure;
    int ref_order;              // VLG_BA_ORDER_REFERENCE: U/eA accumulated in one chain, ascending observation order
    const int* obs_slot;        // C-order -> P-order slot of each observation
    double* W;
    double* BeP;                // [nobs][8], P-order: B (2 x 3) and e of each observation, for the point pass
    double* Upart;              // [nchunks][NU]
    // diagnostics (NULL in production)
    double* dX_hat; double* dA; double* dB; double* de;
};

// ---------------------------------------------------------------------------------------
// stage 1, camera pass ("resid_jac_normal"): one warp per chunk of one camera's segment.
// Per observation: X_hat, A, B, e ONCE (bit-exact, ba_math.cuh), W_ij = A'B written through a
// shared-memory tile so that the global stores are contiguous, B and e left as one 64-byte record (two
// whole 32-byte sectors) at the observation's P-order slot for the point pass, and the NU = NA(NA+1)/2 + NA
// distinct entries of A'A and A'e staged in shared memory and accumulated by lane r in
// ascending observation order -- the reference's order (mex_bundle_1_XABeUVWeAeB.c:281-290,
// :317-323).  One partial per chunk; k_stage1_cam_finalize adds the chunks of a camera in
// order.  With one chunk per camera (VLG_BA_ORDER_REFERENCE) U and eA are bit-identical to
// the reference.
// ---------------------------------------------------------------------------------------
// the same observation with one __ddiv_rn per quotient: taken when a depth is so small or so large that the shared
// reciprocals of the fast path could over/underflow (never on a sane scene)
template <int NA>
__device__ __noinline__ void obs_jacobian_exact(const double* __restrict__ R4, const double* __restrict__ a, double fx, double fy,
                                                double cx, double cy, double b0, double b1, double b2, double ox, double oy,
                                                double* __restrict__ X0, double* __restrict__ A, double* __restrict__ B,
                                                double* __restrict__ e)
{
    if constexpr (NA == kNaProjective) obs_jacobian_proj<false>(a, b0, b1, b2, ox, oy, X0, A, B, e);
    else obs_jacobian<NA, false>(R4, a, fx, fy, cx, cy, b0, b1, b2, ox, oy, X0, A, B, e);
}

template <int NA, bool DIAG>
#ifndef VLG_S1_MINB
#define VLG_S1_MINB 3
#endif
__global__ void __launch_bounds__(kWarpsPerBlock * 32, NA == 6 ? VLG_S1_MINB : 1)      // NA = 6: <= 168 registers, 3 CTAs (12 warps) per SM -- with the two-deep prefetch 128 registers spill (measured 0.48 vs 0.39 ms at Venice shape)
k_stage1_cam(Stage1Args p)
{
    constexpr int NU = nu_of(NA);
    constexpr int NW = 3 * NA;
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* my = smem + (size_t)warp * (36 + NA + 4 + 32 * NU + 32 * NW);
    double* camR = my;                 // 36
    double* cama = my + 36;            // NA
    double* camK = my + 36 + NA;       // 4
    double* ust = my + 36 + NA + 4;    // 32 x NU
    double* wst = ust + 32 * NU;       // 32 x NW
    const int c = blockIdx.x * kWarpsPerBlock + warp;
    if (c >= p.nchunks) return;
    const int j = p.chunk_cam[c];
    const int beg = p.chunk_begin[c], end = p.chunk_end[c];
    constexpr bool kProj = NA == kNaProjective;        // projective model: the camera is its 12 parameters, no K, no rotation table
    if (!kProj) {
        for (int t = lane; t < 36; t += 32) camR[t] = p.rtab[(size_t)36 * j + t];
        if (lane < 4) camK[lane] = p.K4[(size_t)4 * j + lane];
    }
    if (lane < NA) cama[lane] = p.a[(size_t)NA * j + lane];
    __syncwarp();
    double fx = 0.0, fy = 0.0, cx = 0.0, cy = 0.0;
    if (!kProj) effective_K<NA>(camK, cama, fx, fy, cx, cy);
    const bool zeroW = p.fix_structure || p.cam_fixed[j];

    // U/eA partial of the chunk: lane r owns entry r (and r+32, r+64 for NA = 10).  Reference order: one chain in ascending
    // observation order; chunked order (default): four interleaved chains (observation k -> chain k & 3), combined as
    // (c0 + c1) + (c2 + c3) at the end -- a fixed shape as well, and a quarter of the dependent-add latency.
    constexpr int NACC = (NU + 31) / 32;
    double acc[NACC][4];
#pragma unroll
    for (int t = 0; t < NACC; t++)
#pragma unroll
        for (int u = 0; u < 4; u++) acc[t][u] = 0.0;
    const bool ref_order = p.ref_order != 0;

    // software pipeline over the batches of 32 observations: the observation records (image point, point id, P-order
    // slot) are fetched two batches ahead and the point coordinates one batch ahead, so that a batch never waits for the
    // two dependent memory round trips id -> b (ncu, round 2: long-scoreboard stalls were 60 % of this kernel's issue slots)
    struct Rec { double2 xy; int i, slot; };
    auto load_rec = [&](int o) {
        Rec r;
        r.xy = make_double2(0.0, 0.0); r.i = 0; r.slot = 0;
        if (o < end) { r.xy = __ldg(p.obs_xy + o); r.i = __ldg(p.obs_pt + o); r.slot = __ldg(p.obs_slot + o); }
        return r;
    };
    Rec cur = load_rec(beg + lane), nxt = load_rec(beg + 32 + lane);
    double b0 = 0.0, b1 = 0.0, b2 = 0.0;
    if (beg + lane < end) { b0 = __ldg(p.b + (size_t)3 * cur.i); b1 = __ldg(p.b + (size_t)3 * cur.i + 1); b2 = __ldg(p.b + (size_t)3 * cur.i + 2); }

    for (int base = beg; base < end; base += 32) {
        const int o = base + lane;
        const int cnt = min(32, end - base);
        // next batch's point, the batch after's record
        double n0 = 0.0, n1 = 0.0, n2 = 0.0;
        if (o + 32 < end) { n0 = __ldg(p.b + (size_t)3 * nxt.i); n1 = __ldg(p.b + (size_t)3 * nxt.i + 1); n2 = __ldg(p.b + (size_t)3 * nxt.i + 2); }
        const Rec nn = load_rec(o + 64);
        if (o < end) {
            const double2 xy = cur.xy;
            double X0[2], A[2 * NA], B[6], e[2];
            bool ok;
            if constexpr (kProj) ok = obs_jacobian_proj<true>(cama, b0, b1, b2, xy.x, xy.y, X0, A, B, e);
            else ok = obs_jacobian<NA, true>(camR, cama, fx, fy, cx, cy, b0, b1, b2, xy.x, xy.y, X0, A, B, e);
            if (!ok) {
                // a depth outside the reciprocals' safe window: one __ddiv_rn per quotient (own buffers, so that the
                // arrays of the fast path never have their address taken and stay in registers)
                double Xs[2], As[2 * NA], Bs[6], es[2];
                obs_jacobian_exact<NA>(camR, cama, fx, fy, cx, cy, b0, b1, b2, xy.x, xy.y, Xs, As, Bs, es);
                X0[0] = Xs[0]; X0[1] = Xs[1]; e[0] = es[0]; e[1] = es[1];
#pragma unroll
                for (int k = 0; k < 2 * NA; k++) A[k] = As[k];
#pragma unroll
                for (int k = 0; k < 6; k++) B[k] = Bs[k];
            }
            {
                // B | e -> the observation's P-order slot: 64 bytes, 64-byte aligned
                double2* rec = reinterpret_cast<double2*>(p.BeP + (size_t)8 * cur.slot);
                rec[0] = make_double2(B[0], B[1]); rec[1] = make_double2(B[2], B[3]);
                rec[2] = make_double2(B[4], B[5]); rec[3] = make_double2(e[0], e[1]);
            }
            if (DIAG) {
                if (p.dX_hat) { p.dX_hat[(size_t)2 * o] = X0[0]; p.dX_hat[(size_t)2 * o + 1] = X0[1]; }
                if (p.dA) for (int k = 0; k < 2 * NA; k++) p.dA[(size_t)2 * NA * o + k] = A[k];
                if (p.dB) for (int k = 0; k < 6; k++) p.dB[(size_t)6 * o + k] = B[k];
                if (p.de) { p.de[(size_t)2 * o] = e[0]; p.de[(size_t)2 * o + 1] = e[1]; }
            }
            // W(:,:,i,j) = A'B (mex_bundle_1_XABeUVWeAeB.c:305-314)
#pragma unroll
            for (int col = 0; col < 3; col++)
#pragma unroll
                for (int row = 0; row < NA; row++)
                    wst[lane * NW + row + NA * col] =
                        zeroW ? 0.0 : dot2(A[2 * row], A[2 * row + 1], B[2 * col], B[2 * col + 1]);
            // upper triangle of A'A, then A'e
#pragma unroll
            for (int col = 0; col < NA; col++)
#pragma unroll
                for (int row = 0; row <= col; row++)
                    ust[lane * NU + col * (col + 1) / 2 + row] =
                        dot2(A[2 * row], A[2 * row + 1], A[2 * col], A[2 * col + 1]);
#pragma unroll
            for (int row = 0; row < NA; row++)
                ust[lane * NU + NA * (NA + 1) / 2 + row] = dot2(A[2 * row], A[2 * row + 1], e[0], e[1]);
        }
        __syncwarp();
        // contiguous store of the W tile
        {
            double* Wg = p.W + (size_t)NW * base;
            const int tot = cnt * NW;
            for (int t = lane; t < tot; t += 32) Wg[t] = wst[t];
        }
#pragma unroll
        for (int t = 0; t < NACC; t++) {
            const int r = lane + 32 * t;
            if (r < NU) {
                if (ref_order) {
                    double sv = acc[t][0];
                    for (int k = 0; k < cnt; k++) sv = VLG_P(sv, ust[k * NU + r]);
                    acc[t][0] = sv;
                } else {
                    int k = 0;
                    for (; k + 3 < cnt; k += 4) {
                        acc[t][0] = VLG_P(acc[t][0], ust[k * NU + r]);
                        acc[t][1] = VLG_P(acc[t][1], ust[(k + 1) * NU + r]);
                        acc[t][2] = VLG_P(acc[t][2], ust[(k + 2) * NU + r]);
                        acc[t][3] = VLG_P(acc[t][3], ust[(k + 3) * NU + r]);
                    }
                    for (; k < cnt; k++) acc[t][0] = VLG_P(acc[t][0], ust[k * NU + r]);     // tail of the chunk's last batch
                }
            }
        }
        __syncwarp();
        cur = nxt; nxt = nn;
        b0 = n0; b1 = n1; b2 = n2;
    }
#pragma unroll
    for (int t = 0; t < NACC; t++) {
        const int r = lane + 32 * t;
        if (r < NU) p.Upart[(size_t)NU * c + r] = ref_order ? acc[t][0] : VLG_P(VLG_P(acc[t][0], acc[t][1]), VLG_P(acc[t][2], acc[t][3]));
    }
}

// U_j, eA_j = sum of the camera's chunk partials in chunk order; fix_motion / fix_pivot
// cameras are zeroed here (bundle_euclid.m:145-154).  One thread per (camera, entry).
template <int NA>
__global__ void k_stage1_cam_finalize(int m, const int* __restrict__ cam_chunk_ptr,
                                      const double* __restrict__ Upart,
                                      const unsigned char* __restrict__ cam_fixed,
                                      double* __restrict__ U, double* __restrict__ eA)
{
    constexpr int NU = nu_of(NA);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * NU) return;
    const int j = t / NU, r = t % NU;
    double s = 0.0;
    const int c0 = cam_chunk_ptr[j], c1 = cam_chunk_ptr[j + 1];
    if (c1 > c0) {
        s = Upart[(size_t)NU * c0 + r];
        for (int c = c0 + 1; c < c1; c++) s = VLG_P(s, Upart[(size_t)NU * c + r]);
    }
    if (cam_fixed[j]) s = 0.0;
    if (r < NA * (NA + 1) / 2) {
        int col = 0;
        while ((col + 1) * (col + 2) / 2 <= r) col++;
        const int row = r - col * (col + 1) / 2;
        U[(size_t)NA * NA * j + row + NA * col] = s;
        U[(size_t)NA * NA * j + col + NA * row] = s;
    } else {
        eA[(size_t)NA * j + (r - NA * (NA + 1) / 2)] = s;
    }
}

// The 10 products a point accumulates per observation, from the record the camera pass left: B'B (upper triangle),
// B'e and e'e, each the reference's two-term sum (mex_bundle_1_XABeUVWeAeB.c:293-302, :326-332)
__device__ __forceinline__ void point_products(const double* __restrict__ rec, double* __restrict__ pr)
{
    const double2* r = reinterpret_cast<const double2*>(rec);
    const double2 B0 = r[0], B1 = r[1], B2 = r[2], e = r[3];
    pr[0] = dot2(B0.x, B0.y, B0.x, B0.y);
    pr[1] = dot2(B1.x, B1.y, B0.x, B0.y);
    pr[2] = dot2(B2.x, B2.y, B0.x, B0.y);
    pr[3] = dot2(B1.x, B1.y, B1.x, B1.y);
    pr[4] = dot2(B2.x, B2.y, B1.x, B1.y);
    pr[5] = dot2(B2.x, B2.y, B2.x, B2.y);
    pr[6] = dot2(B0.x, B0.y, e.x, e.y);
    pr[7] = dot2(B1.x, B1.y, e.x, e.y);
    pr[8] = dot2(B2.x, B2.y, e.x, e.y);
    pr[9] = e.x * e.x + e.y * e.y;
}

// ---------------------------------------------------------------------------------------
// stage 1, point pass: V_i, eB_i and the point's share of the cost e'e (bundle_euclid.m:209) from the
// B | e records of the camera pass (P-order, so a point's track is contiguous), added in ascending
// camera order -- the reference's order for V_i and eB_i.  Nothing is reprojected a second time.
// Thread-per-point form (tracks longer than a tile).
// ---------------------------------------------------------------------------------------
__global__ void k_stage1_pt(int n, const int* __restrict__ pt_ptr, const double* __restrict__ BeP, int fix_structure,
                            double* __restrict__ V, double* __restrict__ eB, double* __restrict__ cost_pt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v00 = 0, v10 = 0, v20 = 0, v11 = 0, v21 = 0, v22 = 0, g0 = 0, g1 = 0, g2 = 0, cost = 0;
    const int q0 = pt_ptr[i], q1 = pt_ptr[i + 1];
    for (int q = q0; q < q1; q++) {
        double pr[10];
        point_products(BeP + (size_t)8 * q, pr);
        v00 = VLG_P(v00, pr[0]); v10 = VLG_P(v10, pr[1]); v20 = VLG_P(v20, pr[2]);
        v11 = VLG_P(v11, pr[3]); v21 = VLG_P(v21, pr[4]); v22 = VLG_P(v22, pr[5]);
        g0 = VLG_P(g0, pr[6]); g1 = VLG_P(g1, pr[7]); g2 = VLG_P(g2, pr[8]);
        cost += pr[9];
    }
    if (fix_structure) { v00 = v10 = v20 = v11 = v21 = v22 = 0.0; g0 = g1 = g2 = 0.0; }   // bundle_euclid.m:140-144
    double* Vi = V + (size_t)9 * i;
    Vi[0] = v00; Vi[1] = v10; Vi[2] = v20;
    Vi[3] = v10; Vi[4] = v11; Vi[5] = v21;
    Vi[6] = v20; Vi[7] = v21; Vi[8] = v22;
    eB[(size_t)3 * i] = g0; eB[(size_t)3 * i + 1] = g1; eB[(size_t)3 * i + 2] = g2;
    cost_pt[i] = cost;
}

// ---------------------------------------------------------------------------------------
// The same with one LANE per observation: a CTA owns a tile of whole points (<= kS1Tile observations); every
// thread forms the 10 products of one record, then one thread per point adds its track in ascending camera
// order -- bit-identical to k_stage1_pt, without the divergence of a thread-per-point loop over tracks of 2..64.
// ---------------------------------------------------------------------------------------
constexpr int kS1Tile = 512;

__global__ void __launch_bounds__(kS1Tile)
k_stage1_pt_tiled(const int4* __restrict__ ptile_meta /* (q0, nob, p0, npts) */, const int* __restrict__ pt_ptr,
                  const double* __restrict__ BeP, int fix_structure, double* __restrict__ V,
                  double* __restrict__ eB, double* __restrict__ cost_pt)
{
    __shared__ double pr[10][kS1Tile];
    const int4 meta = __ldg(ptile_meta + blockIdx.x);
    const int q0 = meta.x, nob = meta.y, p0 = meta.z, np = meta.w, tid = threadIdx.x;
    if (tid < nob) {
        double v[10];
        point_products(BeP + (size_t)8 * (q0 + tid), v);
#pragma unroll
        for (int k = 0; k < 10; k++) pr[k][tid] = v[k];
    }
    __syncthreads();
    if (tid < np) {
        const int i = p0 + tid;
        const int o0 = pt_ptr[i] - q0, o1 = pt_ptr[i + 1] - q0;
        double v00 = 0, v10 = 0, v20 = 0, v11 = 0, v21 = 0, v22 = 0, g0 = 0, g1 = 0, g2 = 0, cost = 0;
        for (int o = o0; o < o1; o++) {
            v00 = VLG_P(v00, pr[0][o]); v10 = VLG_P(v10, pr[1][o]); v20 = VLG_P(v20, pr[2][o]);
            v11 = VLG_P(v11, pr[3][o]); v21 = VLG_P(v21, pr[4][o]); v22 = VLG_P(v22, pr[5][o]);
            g0 = VLG_P(g0, pr[6][o]); g1 = VLG_P(g1, pr[7][o]); g2 = VLG_P(g2, pr[8][o]);
            cost += pr[9][o];
        }
        if (fix_structure) { v00 = v10 = v20 = v11 = v21 = v22 = 0.0; g0 = g1 = g2 = 0.0; }   // bundle_euclid.m:140-144
        double* Vi = V + (size_t)9 * i;
        Vi[0] = v00; Vi[1] = v10; Vi[2] = v20;
        Vi[3] = v10; Vi[4] = v11; Vi[5] = v21;
        Vi[6] = v20; Vi[7] = v21; Vi[8] = v22;
        eB[(size_t)3 * i] = g0; eB[(size_t)3 * i + 1] = g1; eB[(size_t)3 * i + 2] = g2;
        cost_pt[i] = cost;
    }
}

// ---------------------------------------------------------------------------------------
// deterministic sum of a long vector: fixed-shape two-level tree (grid of kRedBlocks CTAs,
// then one CTA).  out[slot] = sum(in[0..n)).
// ---------------------------------------------------------------------------------------
constexpr int kRedBlocks = 296;
constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum_256(double v, double* sh)
{
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;   // valid in warp 0
}

__global__ void __launch_bounds__(kRedThreads) k_reduce_partial(const double* __restrict__ in, size_t n,
                                                                double* __restrict__ part)
{
    __shared__ double sh[32];
    double v = 0.0;
    for (size_t t = (size_t)blockIdx.x * kRedThreads + threadIdx.x; t < n; t += (size_t)kRedBlocks * kRedThreads)
        v += in[t];
    v = block_sum_256(v, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
}

__global__ void __launch_bounds__(kRedThreads) k_reduce_final(const double* __restrict__ part, int nparts,
                                                              double* __restrict__ out)
{
    __shared__ double sh[32];
    double v = 0.0;
    for (int t = threadIdx.x; t < nparts; t += kRedThreads) v += part[t];
    v = block_sum_256(v, sh);
    if (threadIdx.x == 0) *out = v;
}

// ---------------------------------------------------------------------------------------
// damping (bundle_euclid.m:162-173) and V*^-1 = pinv(V*) (bundle_euclid.m:178-181)
// ---------------------------------------------------------------------------------------
template <int NA>
__global__ void k_damp_U(int m, double lambda, const double* __restrict__ U, double* __restrict__ Ud)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * NA * NA) return;
    const int r = t % (NA * NA);
    const double u = U[t];
    Ud[t] = (r % NA == r / NA) ? (1 + lambda) * u : u;
}

// VE (optional): per point V*^-1 | eB as one 96-byte record (three whole sectors) for the cooperative gathers of ba_schur.cuh
__global__ void k_vinv_damp(int n, double lambda, const double* __restrict__ V, double* __restrict__ Vinv,
                            const double* __restrict__ eB = nullptr, double* __restrict__ VE = nullptr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double Vd[9], Vi[9];
#pragma unroll
    for (int k = 0; k < 9; k++) Vd[k] = V[(size_t)9 * i + k];
    Vd[0] = (1 + lambda) * Vd[0]; Vd[4] = (1 + lambda) * Vd[4]; Vd[8] = (1 + lambda) * Vd[8];
    sym_pinv<3>(Vd, Vi);
#pragma unroll
    for (int k = 0; k < 9; k++) Vinv[(size_t)9 * i + k] = Vi[k];
    if (VE) {
        double2* r = reinterpret_cast<double2*>(VE + (size_t)12 * i);
        r[0] = make_double2(Vi[0], Vi[1]); r[1] = make_double2(Vi[2], Vi[3]); r[2] = make_double2(Vi[4], Vi[5]);
        r[3] = make_double2(Vi[6], Vi[7]); r[4] = make_double2(Vi[8], eB[(size_t)3 * i]);
        r[5] = make_double2(eB[(size_t)3 * i + 1], eB[(size_t)3 * i + 2]);
    }
}

// An observation's NW-double block in global memory <-> registers; 16-byte accesses when NW is even
// (every block then starts on a 16-byte boundary): half the load/store instructions, i.e. half the
// LSU wavefronts of a gather whose lanes all hit different lines.
template <int NW>
__device__ __forceinline__ void load_block_g(const double* __restrict__ base, size_t idx, double* __restrict__ w)
{
    if constexpr (NW % 2 == 0) {
        const double2* src = reinterpret_cast<const double2*>(base + idx * NW);
#pragma unroll
        for (int k = 0; k < NW / 2; k++) { const double2 v = __ldg(src + k); w[2 * k] = v.x; w[2 * k + 1] = v.y; }
    } else {
#pragma unroll
        for (int k = 0; k < NW; k++) w[k] = __ldg(base + idx * NW + k);
    }
}

template <int NW>
__device__ __forceinline__ void store_block_g(double* __restrict__ base, size_t idx, const double* __restrict__ w)
{
    if constexpr (NW % 2 == 0) {
        double2* dst = reinterpret_cast<double2*>(base + idx * NW);
#pragma unroll
        for (int k = 0; k < NW / 2; k++) dst[k] = make_double2(w[2 * k], w[2 * k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < NW; k++) base[idx * NW + k] = w[k];
    }
}

// ---------------------------------------------------------------------------------------
// camera-keyed Schur pieces: e_j = eA_j - sum_i Y_ij eB_i (mex_bundle_2_Se_.c:132-155) and
// the diagonal blocks S_jj = U*_j - sum_i Y_ij W_ij' (mex_bundle_2_Se_.c:80-118 with k = j),
// Y_ij = W_ij V*_i^-1 formed on the fly (bundle_euclid.m:182-184; Y is never stored).
// One warp per chunk, fixed xor-tree over lanes, chunk partials added in order.
// ---------------------------------------------------------------------------------------
template <int NA>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_cam_schur_diag(int nchunks, const int* __restrict__ chunk_begin, const int* __restrict__ chunk_end,
                 const int* __restrict__ obs_pt, const double* __restrict__ W, const double* __restrict__ Vinv,
                 const double* __restrict__ eB, double* __restrict__ part /* [nchunks][NU] */,
                 double* __restrict__ Yout = nullptr /* [nobs][3 NA], C-order: kept for the assembly of S */)
{
    constexpr int NU = nu_of(NA);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * kWarpsPerBlock + warp;
    if (c >= nchunks) return;
    const int beg = chunk_begin[c], end = chunk_end[c];
    double acc[NU];
#pragma unroll
    for (int t = 0; t < NU; t++) acc[t] = 0.0;
    for (int o = beg + lane; o < end; o += 32) {
        const int i = obs_pt[o];
        double Wo[3 * NA], Vi[9], g[3], Y[3 * NA];
        load_block_g<3 * NA>(W, (size_t)o, Wo);
#pragma unroll
        for (int k = 0; k < 9; k++) Vi[k] = __ldg(Vinv + (size_t)9 * i + k);
#pragma unroll
        for (int k = 0; k < 3; k++) g[k] = __ldg(eB + (size_t)3 * i + k);
#pragma unroll
        for (int cc = 0; cc < 3; cc++)
#pragma unroll
            for (int r = 0; r < NA; r++)
                Y[r + NA * cc] = Wo[r] * Vi[3 * cc] + Wo[r + NA] * Vi[1 + 3 * cc] + Wo[r + 2 * NA] * Vi[2 + 3 * cc];
        if (Yout) store_block_g<3 * NA>(Yout, o, Y);
#pragma unroll
        for (int col = 0; col < NA; col++)
#pragma unroll
            for (int row = 0; row <= col; row++)
                acc[col * (col + 1) / 2 + row] +=
                    Y[row] * Wo[col] + Y[row + NA] * Wo[col + NA] + Y[row + 2 * NA] * Wo[col + 2 * NA];
#pragma unroll
        for (int r = 0; r < NA; r++)
            acc[NA * (NA + 1) / 2 + r] += Y[r] * g[0] + Y[r + NA] * g[1] + Y[r + 2 * NA] * g[2];
    }
#pragma unroll
    for (int t = 0; t < NU; t++) {
        double v = warp_sum(acc[t]);
        if (lane == (t & 31)) part[(size_t)NU * c + t] = v;
    }
}

// per-camera sum of chunk partials in chunk order: out[j][r] = sum_c part[c][r], r < nv
__global__ void k_cam_sum_partials(int m, int nv, const int* __restrict__ cam_chunk_ptr, const double* __restrict__ part,
                                   const int* __restrict__ done, double* __restrict__ out)
{
    if (done && *done) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * nv) return;
    const int j = t / nv, r = t % nv;
    double s = 0.0;
    for (int c = cam_chunk_ptr[j]; c < cam_chunk_ptr[j + 1]; c++) s += part[(size_t)nv * c + r];
    out[t] = s;
}

// S_jj = U*_j - sums (full NA x NA, symmetric), e_j = eA_j - sums, and optionally the
// block-Jacobi preconditioner M_j^-1 = pinv(S_jj).  sums = [m][NU] (all-reduced over ranks).
template <int NA>
__global__ void k_cam_schur_finalize(int m, const double* __restrict__ sums, const double* __restrict__ Ud,
                                     const double* __restrict__ eA, double* __restrict__ Sjj,
                                     double* __restrict__ ebar, double* __restrict__ Minv)
{
    constexpr int NU = nu_of(NA);
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double* s = sums + (size_t)NU * j;
    double M[NA * NA];
#pragma unroll
    for (int col = 0; col < NA; col++)
#pragma unroll
        for (int row = 0; row <= col; row++) {
            const double v = Ud[(size_t)NA * NA * j + row + NA * col] - s[col * (col + 1) / 2 + row];
            M[row + NA * col] = v;
            M[col + NA * row] = v;
        }
#pragma unroll
    for (int r = 0; r < NA; r++) ebar[(size_t)NA * j + r] = eA[(size_t)NA * j + r] - s[NA * (NA + 1) / 2 + r];
    if (Sjj)
#pragma unroll
        for (int k = 0; k < NA * NA; k++) Sjj[(size_t)NA * NA * j + k] = M[k];
    if (Minv) {
        double Mi[NA * NA];
        sym_pinv<NA>(M, Mi);
#pragma unroll
        for (int k = 0; k < NA * NA; k++) Minv[(size_t)NA * NA * j + k] = Mi[k];
    }
}

// ---------------------------------------------------------------------------------------
// explicit Schur complement, block-sparse: one warp per structurally non-zero block (j,k),
// j <= k, walking the precomputed list of (observation of j, observation of k) pairs that
// share a point, ascending point index (mex_bundle_2_Se_.c:103-118 restricted to the terms
// that are not exactly zero).  Writes S_jk and its mirror S_kj = S_jk' into the dense S.
// ---------------------------------------------------------------------------------------
template <int NA>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_schur_blocks(int nblocks, int N, int add_U, const int* __restrict__ blk_j, const int* __restrict__ blk_k,
               const int64_t* __restrict__ blk_ptr, const int2* __restrict__ pairs,
               const int* __restrict__ obs_pt, const double* __restrict__ W, const double* __restrict__ Vinv,
               const double* __restrict__ Ud, double* __restrict__ S, const double* __restrict__ Yext = nullptr)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bidx = blockIdx.x * kWarpsPerBlock + warp;
    if (bidx >= nblocks) return;
    const int j = blk_j[bidx], k = blk_k[bidx];
    double acc[NA * NA];
#pragma unroll
    for (int t = 0; t < NA * NA; t++) acc[t] = 0.0;
    for (int64_t q = blk_ptr[bidx] + lane; q < blk_ptr[bidx + 1]; q += 32) {
        const int2 pr = pairs[q];
        const int i = obs_pt[pr.x];
        double Wj[3 * NA], Wk[3 * NA], Vi[9], Y[3 * NA];
#pragma unroll
        for (int t = 0; t < 3 * NA; t++) Wj[t] = W[(size_t)3 * NA * pr.x + t];
#pragma unroll
        for (int t = 0; t < 3 * NA; t++) Wk[t] = W[(size_t)3 * NA * pr.y + t];
        if (Yext) {                 // mex2 drop-in: Y is an input (mex_bundle_2_Se_.c:21)
#pragma unroll
            for (int t = 0; t < 3 * NA; t++) Y[t] = Yext[(size_t)3 * NA * pr.x + t];
        } else {
#pragma unroll
            for (int t = 0; t < 9; t++) Vi[t] = __ldg(Vinv + (size_t)9 * i + t);
#pragma unroll
            for (int cc = 0; cc < 3; cc++)
#pragma unroll
                for (int r = 0; r < NA; r++)
                    Y[r + NA * cc] = Wj[r] * Vi[3 * cc] + Wj[r + NA] * Vi[1 + 3 * cc] + Wj[r + 2 * NA] * Vi[2 + 3 * cc];
        }
#pragma unroll
        for (int col = 0; col < NA; col++)
#pragma unroll
            for (int row = 0; row < NA; row++)
                acc[row + NA * col] += Y[row] * Wk[col] + Y[row + NA] * Wk[col + NA] + Y[row + 2 * NA] * Wk[col + 2 * NA];
    }
    double mine[(NA * NA + 31) / 32];
#pragma unroll
    for (int t = 0; t < NA * NA; t++) {
        const double v = warp_sum(acc[t]);
        if (lane == (t & 31)) mine[t >> 5] = v;
    }
#pragma unroll
    for (int u = 0; u < (NA * NA + 31) / 32; u++) {
        const int t = lane + 32 * u;
        if (t < NA * NA) {
            const int row = t % NA, col = t / NA;
            double v = -mine[u];
            if (j == k && add_U) v += Ud[(size_t)NA * NA * j + t];
            S[(size_t)(NA * j + row) + (size_t)N * (NA * k + col)] = v;
            if (j != k) S[(size_t)(NA * k + col) + (size_t)N * (NA * j + row)] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Assembly of S at scale (explicit-S PCG, Venice shape: 0.93 M non-zero blocks, 21 M pairs).  The
// pair counts are extremely skewed -- 2 % of the blocks (cameras that share a track window) hold
// 92 % of the pairs, the rest hold 1-3 pairs each -- and the work per pair is a gather of two
// 144-byte blocks, i.e. bound by LSU wavefronts (one per distinct line per instruction), not by
// DRAM or FP64.  Hence: Y = W V*^-1 is stored once per observation by k_cam_schur_diag (a pair
// then gathers Y_ij and W_ik with 9 + 9 16-byte loads instead of 18 + 18 + 9 8-byte ones), heavy
// blocks get a warp each (lanes over pairs, one tree reduction per block), light blocks a thread
// each (pairs in ascending point order, no reduction at all).
// ---------------------------------------------------------------------------------------
template <int NA>
__device__ __forceinline__ void schur_pair(const double* __restrict__ Y, const double* __restrict__ W, int2 pr,
                                           double* __restrict__ acc)
{
    double Yj[3 * NA], Wk[3 * NA];
    load_block_g<3 * NA>(Y, (size_t)pr.x, Yj);
    load_block_g<3 * NA>(W, (size_t)pr.y, Wk);
#pragma unroll
    for (int col = 0; col < NA; col++)
#pragma unroll
        for (int row = 0; row < NA; row++)
            acc[row + NA * col] += Yj[row] * Wk[col] + Yj[row + NA] * Wk[col + NA] + Yj[row + 2 * NA] * Wk[col + 2 * NA];
}

// Where block (j,k) lives: in the dense S (leading dimension ld), or -- cluster mode, ccams > 0 -- in the
// 128 x 128 diagonal block of the cluster of `ccams` consecutive cameras that holds both j and k
// (the cluster-Jacobi preconditioner of the implicit-Schur PCG assembles only those blocks).
struct SchurDest { double* base; size_t ld; int jr, kr; };

template <int NA>
__device__ __forceinline__ SchurDest schur_dest(double* __restrict__ S, int ld, int ccams, int j, int k)
{
    SchurDest d;
    if (ccams > 0) {
        const int cl = j / ccams;
        d.base = S + (size_t)cl * 128 * 128; d.ld = 128; d.jr = NA * (j - cl * ccams); d.kr = NA * (k - cl * ccams);
    } else {
        d.base = S; d.ld = (size_t)ld; d.jr = NA * j; d.kr = NA * k;
    }
    return d;
}

// v[row + NA col] -> S block (j,k), j <= k (the upper one; skipped when `upper` is false), and its mirror (k,j);
// 16-byte stores when NA is even
template <int NA>
__device__ __forceinline__ void schur_store_block(const SchurDest& t, bool offdiag, const double* __restrict__ v, bool upper = true)
{
    if constexpr (NA % 2 == 0) {
        if (upper || !offdiag) {
#pragma unroll
            for (int col = 0; col < NA; col++) {
                double2* d = reinterpret_cast<double2*>(t.base + (size_t)t.jr + t.ld * (t.kr + col));
#pragma unroll
                for (int h = 0; h < NA / 2; h++) d[h] = make_double2(v[2 * h + NA * col], v[2 * h + 1 + NA * col]);
            }
        }
        if (offdiag) {
#pragma unroll
            for (int row = 0; row < NA; row++) {
                double2* d = reinterpret_cast<double2*>(t.base + (size_t)t.kr + t.ld * (t.jr + row));
#pragma unroll
                for (int h = 0; h < NA / 2; h++) d[h] = make_double2(v[row + NA * (2 * h)], v[row + NA * (2 * h + 1)]);
            }
        }
    } else {
#pragma unroll
        for (int col = 0; col < NA; col++)
#pragma unroll
            for (int row = 0; row < NA; row++) {
                if (upper || !offdiag) t.base[(size_t)(t.jr + row) + t.ld * (t.kr + col)] = v[row + NA * col];
                if (offdiag) t.base[(size_t)(t.kr + col) + t.ld * (t.jr + row)] = v[row + NA * col];
            }
    }
}

// diagonal blocks of S from the per-camera sums of k_cam_schur_diag (packed upper triangle of
// sum_i Y_ij W_ij', this rank's share): S_jj = (add_U ? U*_j : 0) - sums_j, exactly symmetric
template <int NA>
__global__ void k_schur_diag_fill(int m, int ld, int ccams, int add_U, const double* __restrict__ sums /* [m][NU] */,
                                  const double* __restrict__ Ud, double* __restrict__ S)
{
    constexpr int NU = nu_of(NA);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * NA * NA) return;
    const int j = t / (NA * NA), u = t - j * NA * NA, row = u % NA, col = u / NA;
    const int r = row < col ? row : col, c = row < col ? col : row;
    double v = -sums[(size_t)NU * j + c * (c + 1) / 2 + r];
    if (add_U) v += Ud[(size_t)NA * NA * j + u];
    const SchurDest d = schur_dest<NA>(S, ld, ccams, j, j);
    d.base[(size_t)(d.jr + row) + d.ld * (d.kr + col)] = v;
}

// light blocks: one thread per block, pairs in ascending point order (the reference's order of
// summation, mex_bundle_2_Se_.c:103-118)
template <int NA>
__global__ void __launch_bounds__(128)
k_schur_blocks_light(int nlist, const int* __restrict__ list, int ld, int ccams, int add_U, int upper_band, const int* __restrict__ blk_j,
                     const int* __restrict__ blk_k, const int64_t* __restrict__ blk_ptr, const int2* __restrict__ pairs,
                     const double* __restrict__ Y, const double* __restrict__ W, const double* __restrict__ Ud,
                     double* __restrict__ S)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nlist) return;
    const int b = list[t];
    const int j = blk_j[b], k = blk_k[b];
    double acc[NA * NA];
#pragma unroll
    for (int u = 0; u < NA * NA; u++) acc[u] = 0.0;
    for (int64_t q = blk_ptr[b]; q < blk_ptr[b + 1]; q++) schur_pair<NA>(Y, W, pairs[q], acc);
#pragma unroll
    for (int u = 0; u < NA * NA; u++) {
        double v = -acc[u];
        if (j == k && add_U) v += Ud[(size_t)NA * NA * j + u];
        acc[u] = v;
    }
    // far from the diagonal only the lower block is stored when the consumer reads the lower triangle (assembled-S PCG:
    // the symmetric matvec, the peer pull; the cluster gathers stay within upper_band cameras of the diagonal)
    schur_store_block<NA>(schur_dest<NA>(S, ld, ccams, j, k), j != k, acc, k - j <= upper_band);
}

// heavy blocks: one warp per block, lanes stride over the pairs, values folded by recursive halving
template <int NA>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_schur_blocks_heavy(int nlist, const int* __restrict__ list, int ld, int ccams, int add_U, const int* __restrict__ blk_j,
                     const int* __restrict__ blk_k, const int64_t* __restrict__ blk_ptr, const int2* __restrict__ pairs,
                     const double* __restrict__ Y, const double* __restrict__ W, const double* __restrict__ Ud,
                     double* __restrict__ S)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= nlist) return;
    const int b = list[t];
    const int j = blk_j[b], k = blk_k[b];
    double acc[NA * NA];
#pragma unroll
    for (int u = 0; u < NA * NA; u++) acc[u] = 0.0;
    if constexpr (NA % 2 == 0) {
        // Y W' = sum over the 3 point coordinates c of (column c of Y)(column c of W)': three lanes share a pair, lane
        // s takes column s of both blocks (NA doubles = 3 x 16 B when NA = 6) and adds its outer product into its own
        // accumulators -- the warp reduction below sums the three anyway.  The three lanes of a pair read the same
        // one or two lines, so a load instruction costs ~1/3 of the LSU wavefronts of the lane-per-pair form.
        const int ps = lane / 3, sc = lane - 3 * ps;                 // pair slot 0..9, column; lanes 30, 31 idle
        if (ps < 10) {
            // software pipeline: the pair index is fetched two iterations ahead and the two half-blocks one
            // iteration ahead, so that an iteration costs FMAs, not two dependent memory round trips
            const int64_t q1 = blk_ptr[b + 1];
            int64_t q = blk_ptr[b] + ps;
            int2 prn = q < q1 ? pairs[q] : make_int2(0, 0);
            int2 prnn = q + 10 < q1 ? pairs[q + 10] : make_int2(0, 0);
            double2 yn[NA / 2], wn[NA / 2];
            if (q < q1) {
                const double2* ys = reinterpret_cast<const double2*>(Y + (size_t)3 * NA * prn.x + NA * sc);
                const double2* ws = reinterpret_cast<const double2*>(W + (size_t)3 * NA * prn.y + NA * sc);
#pragma unroll
                for (int h = 0; h < NA / 2; h++) { yn[h] = __ldg(ys + h); wn[h] = __ldg(ws + h); }
            }
            for (; q < q1; q += 10) {
                double yc[NA], wc[NA];
#pragma unroll
                for (int h = 0; h < NA / 2; h++) { yc[2 * h] = yn[h].x; yc[2 * h + 1] = yn[h].y; wc[2 * h] = wn[h].x; wc[2 * h + 1] = wn[h].y; }
                const int2 pnext = prnn;
                if (q + 20 < q1) prnn = pairs[q + 20];
                if (q + 10 < q1) {
                    const double2* ys = reinterpret_cast<const double2*>(Y + (size_t)3 * NA * pnext.x + NA * sc);
                    const double2* ws = reinterpret_cast<const double2*>(W + (size_t)3 * NA * pnext.y + NA * sc);
#pragma unroll
                    for (int h = 0; h < NA / 2; h++) { yn[h] = __ldg(ys + h); wn[h] = __ldg(ws + h); }
                }
#pragma unroll
                for (int col = 0; col < NA; col++)
#pragma unroll
                    for (int row = 0; row < NA; row++) acc[row + NA * col] += yc[row] * wc[col];
            }
        }
    } else {
        for (int64_t q = blk_ptr[b] + lane; q < blk_ptr[b + 1]; q += 32) schur_pair<NA>(Y, W, pairs[q], acc);
    }
    double mine[(NA * NA + 31) / 32];
#pragma unroll
    for (int u = 0; u < NA * NA; u++) {
        const double v = warp_sum(acc[u]);
        if (lane == (u & 31)) mine[u >> 5] = v;
    }
#pragma unroll
    for (int w = 0; w < (NA * NA + 31) / 32; w++) {
        const int u = lane + 32 * w;
        if (u < NA * NA) {
            const int row = u % NA, col = u / NA;
            double v = -mine[w];
            if (j == k && add_U) v += Ud[(size_t)NA * NA * j + u];
            const SchurDest d = schur_dest<NA>(S, ld, ccams, j, k);
            d.base[(size_t)(d.jr + row) + d.ld * (d.kr + col)] = v;
            if (j != k) d.base[(size_t)(d.kr + col) + d.ld * (d.jr + row)] = v;
        }
    }
}

// e_j = eA_j - sum_i Y_ij eB_i with Y given (mex2 drop-in, mex_bundle_2_Se_.c:132-155): one
// thread per (camera, row), ascending i -- the reference's order.
template <int NA>
__global__ void k_ebar_from_Y(int m, const int* __restrict__ cam_ptr, const int* __restrict__ obs_pt,
                              const double* __restrict__ Y, const double* __restrict__ eA,
                              const double* __restrict__ eB, double* __restrict__ ebar)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * NA) return;
    const int j = t / NA, row = t % NA;
    double s = 0.0;
    for (int o = cam_ptr[j]; o < cam_ptr[j + 1]; o++) {
        const double* Yo = Y + (size_t)3 * NA * o;
        const double* g = eB + (size_t)3 * obs_pt[o];
        s = VLG_P(s, VLG_P(VLG_P(VLG_M(Yo[row], g[0]), VLG_M(Yo[row + NA], g[1])), VLG_M(Yo[row + 2 * NA], g[2])));
    }
    ebar[t] = VLG_S(eA[t], s);
}

// ---------------------------------------------------------------------------------------
// implicit Schur matvec q = S p = U* p - W V*^-1 W' p (never forming S), two sweeps.
// ---------------------------------------------------------------------------------------
// sweep 1, point-keyed: t_i = V*_i^-1 sum_j W_ij' p_j  (one thread per point, ascending j)
template <int NA>
__global__ void k_sweep_pt(int n, const int* __restrict__ pt_ptr, const int* __restrict__ pt_obs,
                           const int* __restrict__ pt_cam, const double* __restrict__ W,
                           const double* __restrict__ Vinv, const double* __restrict__ p,
                           const int* __restrict__ done, double* __restrict__ t_out)
{
    if (done && *done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int q = pt_ptr[i]; q < pt_ptr[i + 1]; q++) {
        const double* Wo = W + (size_t)3 * NA * pt_obs[q];
        const double* pj = p + (size_t)NA * pt_cam[q];
#pragma unroll
        for (int r = 0; r < NA; r++) {
            const double pr = __ldg(pj + r);
            s0 += Wo[r] * pr; s1 += Wo[r + NA] * pr; s2 += Wo[r + 2 * NA] * pr;
        }
    }
    const double* Vi = Vinv + (size_t)9 * i;
    t_out[(size_t)4 * i] = Vi[0] * s0 + Vi[3] * s1 + Vi[6] * s2;      // padded to 4 doubles per point
    t_out[(size_t)4 * i + 1] = Vi[1] * s0 + Vi[4] * s1 + Vi[7] * s2;
    t_out[(size_t)4 * i + 2] = Vi[2] * s0 + Vi[5] * s1 + Vi[8] * s2;
    t_out[(size_t)4 * i + 3] = 0.0;
}

// sweep 2, camera-keyed: chunk partial of sum_i W_ij t_i (one warp per chunk)
template <int NA>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_sweep_cam(int nchunks, const int* __restrict__ chunk_begin, const int* __restrict__ chunk_end,
            const int* __restrict__ obs_pt, const double* __restrict__ W, const double* __restrict__ t_in,
            const int* __restrict__ done, double* __restrict__ part /* [nchunks][NA] */)
{
    if (done && *done) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * kWarpsPerBlock + warp;
    if (c >= nchunks) return;
    double acc[NA];
#pragma unroll
    for (int r = 0; r < NA; r++) acc[r] = 0.0;
    for (int o = chunk_begin[c] + lane; o < chunk_end[c]; o += 32) {
        const int i = obs_pt[o];
        const double t0 = __ldg(t_in + (size_t)4 * i), t1 = __ldg(t_in + (size_t)4 * i + 1), t2 = __ldg(t_in + (size_t)4 * i + 2);
        const double* Wo = W + (size_t)3 * NA * o;
#pragma unroll
        for (int r = 0; r < NA; r++) acc[r] += Wo[r] * t0 + Wo[r + NA] * t1 + Wo[r + 2 * NA] * t2;
    }
#pragma unroll
    for (int r = 0; r < NA; r++) {
        const double v = warp_sum(acc[r]);
        if (lane == r) part[(size_t)NA * c + r] = v;
    }
}

// PCG scalars kept on the device: [0] rz, [1] r0norm2, [2] rnorm2, [3] iterations, [4] pq
// `exchanges`: matvec vector exchanges (mailbox epochs) the last launch of the persistent kernel consumed -- one more than
// its completed iterations when it left through the breakdown exit
struct PcgScalars { double rz, r0n2, rn2, pq; int iters; int done; int exchanges; int pad_; };

// one-CTA vector kernels over the N = NA*m reduced unknowns (replicated on every rank)
__device__ __forceinline__ double block_sum_1024(double v, double* sh)
{
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = sh[lane];
    r = warp_sum(r);
    return r;   // every thread
}

// q = U* p - wq;  alpha = rz / p'q;  x += alpha p;  r -= alpha q;  z = M^-1 r;
// beta = r'z / rz;  p = z + beta p;  stop flag when |r| <= rtol |r0|
template <int NA>
__global__ void __launch_bounds__(1024) k_pcg_update(int m, const double* __restrict__ Ud, const double* __restrict__ Minv,
                                                     const double* __restrict__ wq, double* __restrict__ x,
                                                     double* __restrict__ r, double* __restrict__ z, double* __restrict__ p,
                                                     double* __restrict__ q, PcgScalars* __restrict__ sc, double rtol)
{
    __shared__ double sh[32];
    if (sc->done) return;
    const int N = NA * m;
    double pq = 0.0;
    for (int t = threadIdx.x; t < N; t += 1024) {
        const int j = t / NA, row = t % NA;
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < NA; c++) v += Ud[(size_t)NA * NA * j + row + NA * c] * p[(size_t)NA * j + c];
        v -= wq[t];
        q[t] = v;
        pq += p[t] * v;
    }
    pq = block_sum_1024(pq, sh);
    const double rz = sc->rz;
    if (!(pq > 0.0)) {           // breakdown (p in the null space): stop with the current x
        __syncthreads();
        if (threadIdx.x == 0) { sc->done = 2; sc->pq = pq; }
        return;
    }
    const double alpha = rz / pq;
    double rr = 0.0;
    for (int t = threadIdx.x; t < N; t += 1024) {
        x[t] += alpha * p[t];
        const double rv = r[t] - alpha * q[t];
        r[t] = rv;
        rr += rv * rv;
    }
    rr = block_sum_1024(rr, sh);
    __syncthreads();
    double rz_new = 0.0;
    for (int t = threadIdx.x; t < N; t += 1024) {
        const int j = t / NA, row = t % NA;
        double zz = 0.0;
#pragma unroll
        for (int c = 0; c < NA; c++) zz += Minv[(size_t)NA * NA * j + row + NA * c] * r[(size_t)NA * j + c];
        z[t] = zz;
        rz_new += r[t] * zz;
    }
    rz_new = block_sum_1024(rz_new, sh);
    const double beta = rz_new / rz;
    for (int t = threadIdx.x; t < N; t += 1024) p[t] = z[t] + beta * p[t];
    __syncthreads();
    if (threadIdx.x == 0) {
        sc->rz = rz_new; sc->rn2 = rr; sc->pq = pq; sc->iters += 1;
        if (rr <= rtol * rtol * sc->r0n2) sc->done = 1;
    }
}

// ---------------------------------------------------------------------------------------
// stage 3 (mex_bundle_3_db_new.c:100-146): db_i = V*_i^-1 (eB_i - sum_j W_ij' da_j) with the
// reference's association order and (by default) only 6 camera rows (:113-120); b_new = b + db;
// per-point share of db'(lambda db + eB) (bundle_euclid.m:215-217).
// ---------------------------------------------------------------------------------------
template <int NA>
__global__ void k_backsub(int n, const int* __restrict__ pt_ptr, const int* __restrict__ pt_obs,
                          const int* __restrict__ pt_cam, const double* __restrict__ W,
                          const double* __restrict__ Vinv, const double* __restrict__ eB,
                          const double* __restrict__ da, const double* __restrict__ b, double lambda, int all_rows,
                          double* __restrict__ db, double* __restrict__ b_new, double* __restrict__ denom_pt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double w0 = eB[(size_t)3 * i], w1 = eB[(size_t)3 * i + 1], w2 = eB[(size_t)3 * i + 2];
    const double g0 = w0, g1 = w1, g2 = w2;
    const int nrow = all_rows ? NA : 6;
    for (int q = pt_ptr[i]; q < pt_ptr[i + 1]; q++) {
        double Wo[3 * NA];
        load_block_g<3 * NA>(W, (size_t)pt_obs[q], Wo);     // 16-byte gathers: half the LSU wavefronts of 8-byte ones
        const double* dj = da + (size_t)NA * pt_cam[q];
        double s0 = VLG_M(Wo[0], dj[0]), s1 = VLG_M(Wo[NA], dj[0]), s2 = VLG_M(Wo[2 * NA], dj[0]);
#pragma unroll
        for (int r = 1; r < NA; r++) {
            if (r < nrow) {                 // static register indices; nrow = 6 or NA
                const double d = dj[r];
                s0 = VLG_P(s0, VLG_M(Wo[r], d));
                s1 = VLG_P(s1, VLG_M(Wo[r + NA], d));
                s2 = VLG_P(s2, VLG_M(Wo[r + 2 * NA], d));
            }
        }
        w0 = VLG_S(w0, s0); w1 = VLG_S(w1, s1); w2 = VLG_S(w2, s2);
    }
    const double* Vi = Vinv + (size_t)9 * i;
    const double d0 = VLG_P(VLG_P(VLG_M(Vi[0], w0), VLG_M(Vi[3], w1)), VLG_M(Vi[6], w2));
    const double d1 = VLG_P(VLG_P(VLG_M(Vi[1], w0), VLG_M(Vi[4], w1)), VLG_M(Vi[7], w2));
    const double d2 = VLG_P(VLG_P(VLG_M(Vi[2], w0), VLG_M(Vi[5], w1)), VLG_M(Vi[8], w2));
    db[(size_t)3 * i] = d0; db[(size_t)3 * i + 1] = d1; db[(size_t)3 * i + 2] = d2;
    b_new[(size_t)3 * i] = VLG_P(b[(size_t)3 * i], d0);
    b_new[(size_t)3 * i + 1] = VLG_P(b[(size_t)3 * i + 1], d1);
    b_new[(size_t)3 * i + 2] = VLG_P(b[(size_t)3 * i + 2], d2);
    denom_pt[i] = d0 * (lambda * d0 + g0) + d1 * (lambda * d1 + g1) + d2 * (lambda * d2 + g2);
}

// back-substitution with one LANE per observation (tiles of whole points, as k_stage1_pt_tiled): every
// thread forms W_ij' da_j for one observation, one thread per point then subtracts its track in
// ascending camera order -- the same operations in the same order as k_backsub, without the
// divergence of tracks of 2..64 observations.
template <int NA>
__global__ void __launch_bounds__(kS1Tile)
k_backsub_tiled(const int4* __restrict__ ptile_meta, const int* __restrict__ pt_ptr, const int* __restrict__ pt_obs,
                const int* __restrict__ pt_cam, const double* __restrict__ W, const double* __restrict__ Vinv,
                const double* __restrict__ eB, const double* __restrict__ da, const double* __restrict__ b, double lambda,
                int all_rows, double* __restrict__ db, double* __restrict__ b_new, double* __restrict__ denom_pt)
{
    __shared__ double pr[3][kS1Tile];
    const int4 meta = __ldg(ptile_meta + blockIdx.x);
    const int q0 = meta.x, nob = meta.y, p0 = meta.z, np = meta.w, tid = threadIdx.x;
    const int nrow = all_rows ? NA : 6;
    if (tid < nob) {
        const int q = q0 + tid;
        double Wo[3 * NA];
        load_block_g<3 * NA>(W, (size_t)pt_obs[q], Wo);
        const double* dj = da + (size_t)NA * pt_cam[q];
        double s0 = VLG_M(Wo[0], dj[0]), s1 = VLG_M(Wo[NA], dj[0]), s2 = VLG_M(Wo[2 * NA], dj[0]);
#pragma unroll
        for (int r = 1; r < NA; r++) {
            if (r < nrow) {
                const double d = dj[r];
                s0 = VLG_P(s0, VLG_M(Wo[r], d));
                s1 = VLG_P(s1, VLG_M(Wo[r + NA], d));
                s2 = VLG_P(s2, VLG_M(Wo[r + 2 * NA], d));
            }
        }
        pr[0][tid] = s0; pr[1][tid] = s1; pr[2][tid] = s2;
    }
    __syncthreads();
    if (tid < np) {
        const int i = p0 + tid;
        double w0 = eB[(size_t)3 * i], w1 = eB[(size_t)3 * i + 1], w2 = eB[(size_t)3 * i + 2];
        const double g0 = w0, g1 = w1, g2 = w2;
        const int o0 = pt_ptr[i] - q0, o1 = pt_ptr[i + 1] - q0;
        for (int o = o0; o < o1; o++) { w0 = VLG_S(w0, pr[0][o]); w1 = VLG_S(w1, pr[1][o]); w2 = VLG_S(w2, pr[2][o]); }
        const double* Vi = Vinv + (size_t)9 * i;
        const double d0 = VLG_P(VLG_P(VLG_M(Vi[0], w0), VLG_M(Vi[3], w1)), VLG_M(Vi[6], w2));
        const double d1 = VLG_P(VLG_P(VLG_M(Vi[1], w0), VLG_M(Vi[4], w1)), VLG_M(Vi[7], w2));
        const double d2 = VLG_P(VLG_P(VLG_M(Vi[2], w0), VLG_M(Vi[5], w1)), VLG_M(Vi[8], w2));
        db[(size_t)3 * i] = d0; db[(size_t)3 * i + 1] = d1; db[(size_t)3 * i + 2] = d2;
        b_new[(size_t)3 * i] = VLG_P(b[(size_t)3 * i], d0);
        b_new[(size_t)3 * i + 1] = VLG_P(b[(size_t)3 * i + 1], d1);
        b_new[(size_t)3 * i + 2] = VLG_P(b[(size_t)3 * i + 2], d2);
        denom_pt[i] = d0 * (lambda * d0 + g0) + d1 * (lambda * d1 + g1) + d2 * (lambda * d2 + g2);
    }
}

// new residual at (a_new, b_new): per-observation squared error (mex_bundle_3_db_new.c:149-166
// + bundle_euclid.m:205,210).  C-order: the camera is (nearly) warp-uniform.
template <int NA>
__global__ void k_new_cost(int64_t nobs, const double2* __restrict__ obs_xy, const int* __restrict__ obs_pt,
                           const int* __restrict__ obs_cam, const double* __restrict__ K4,
                           const double* __restrict__ a_new, const double* __restrict__ b_new,
                           const double* __restrict__ rtab_new /* [m][9] */, double* __restrict__ cost_obs,
                           double* __restrict__ xhat_out /* optional [nobs][2] */)
{
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nobs) return;
    const int i = obs_pt[o], j = obs_cam[o];
    double x, y;
    if constexpr (NA == kNaProjective) {
        double Pl[12];
#pragma unroll
        for (int k = 0; k < 12; k++) Pl[k] = __ldg(a_new + (size_t)NA * j + k);
        project_P(Pl, b_new[(size_t)3 * i], b_new[(size_t)3 * i + 1], b_new[(size_t)3 * i + 2], x, y);
    } else {
        double R[9], al[NA], Kl[4];
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = __ldg(rtab_new + (size_t)9 * j + k);
#pragma unroll
        for (int k = 3; k < NA; k++) al[k] = __ldg(a_new + (size_t)NA * j + k);
#pragma unroll
        for (int k = 0; k < 4; k++) Kl[k] = __ldg(K4 + (size_t)4 * j + k);
        double fx, fy, cx, cy;
        effective_K<NA>(Kl, al, fx, fy, cx, cy);
        project_R(R, al[3], al[4], al[5], fx, fy, cx, cy, b_new[(size_t)3 * i], b_new[(size_t)3 * i + 1],
                  b_new[(size_t)3 * i + 2], x, y);
    }
    const double2 xy = obs_xy[o];
    const double e0 = VLG_S(xy.x, x), e1 = VLG_S(xy.y, y);
    cost_obs[o] = e0 * e0 + e1 * e1;
    if (xhat_out) { xhat_out[2 * o] = x; xhat_out[2 * o + 1] = y; }
}

// rotation tables on the device (VLG_BA_RTABLE_DEVICE): 4 matrices per camera for stage 1,
// or 1 (base only) for the new cost.
template <int NA>
__global__ void k_rtab(int m, const double* __restrict__ a, int nmat, double* __restrict__ rtab)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * nmat) return;
    const int j = t / nmat, k = t % nmat;
    double w0 = a[(size_t)NA * j], w1 = a[(size_t)NA * j + 1], w2 = a[(size_t)NA * j + 2];
    if (k == 1) w0 = VLG_P(w0, kFdStep);
    if (k == 2) w1 = VLG_P(w1, kFdStep);
    if (k == 3) w2 = VLG_P(w2, kFdStep);
    double R[9];
    rodrigues_dev(w0, w1, w2, R);
#pragma unroll
    for (int q = 0; q < 9; q++) rtab[(size_t)9 * t + q] = R[q];
}

__global__ void k_axpy1(int N, const double* __restrict__ a, const double* __restrict__ da, double* __restrict__ a_new)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < N) a_new[t] = VLG_P(a[t], da[t]);
}

// da'(lambda da + eA) over the N reduced unknowns, one CTA
__global__ void __launch_bounds__(1024) k_denom_cam(int N, const double* __restrict__ da, const double* __restrict__ eA,
                                                    double lambda, double* __restrict__ out)
{
    __shared__ double sh[32];
    double v = 0.0;
    for (int t = threadIdx.x; t < N; t += 1024) v += da[t] * (lambda * da[t] + eA[t]);
    v = block_sum_1024(v, sh);
    if (threadIdx.x == 0) *out = v;
}

// ---------------------------------------------------------------------------------------
// reprojection-error map of the current state on the observation list:
// error_reproj.m:72-84 (err(i,j) = ||x(1:2,i,j) - x_reproj(1:2)||) and the statistics of
// remove_outlier (incr_reconstruction.m:363-390: depth test x_reproj(3) < 0 || > depth_max, largest
// squared error and where).  One thread per observation; the projection is project_R / project_P.
// ---------------------------------------------------------------------------------------
template <int NA>
__global__ void k_reproj_errors(int64_t nobs, const double2* __restrict__ obs_xy, const int* __restrict__ obs_pt,
                                const int* __restrict__ obs_cam, const double* __restrict__ K4, const double* __restrict__ a,
                                const double* __restrict__ b, const double* __restrict__ rtab /* [m][4][9] */,
                                double depth_max, double* __restrict__ err, double* __restrict__ depth,
                                double* __restrict__ bad /* 1.0 where the depth test fails */)
{
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nobs) return;
    const int i = obs_pt[o], j = obs_cam[o];
    const double b0 = b[(size_t)3 * i], b1 = b[(size_t)3 * i + 1], b2 = b[(size_t)3 * i + 2];
    double x, y, z;
    if constexpr (NA == kNaProjective) {
        double Pl[12];
#pragma unroll
        for (int k = 0; k < 12; k++) Pl[k] = __ldg(a + (size_t)NA * j + k);
        project_P(Pl, b0, b1, b2, x, y);
        z = VLG_P(VLG_P(VLG_P(VLG_M(Pl[2], b0), VLG_M(Pl[5], b1)), VLG_M(Pl[8], b2)), Pl[11]);
    } else {
        double R[9], al[NA], Kl[4];
#pragma unroll
        for (int k = 0; k < 9; k++) R[k] = __ldg(rtab + (size_t)36 * j + k);
#pragma unroll
        for (int k = 3; k < NA; k++) al[k] = __ldg(a + (size_t)NA * j + k);
#pragma unroll
        for (int k = 0; k < 4; k++) Kl[k] = __ldg(K4 + (size_t)4 * j + k);
        double fx, fy, cx, cy;
        effective_K<NA>(Kl, al, fx, fy, cx, cy);
        project_R(R, al[3], al[4], al[5], fx, fy, cx, cy, b0, b1, b2, x, y);
        z = VLG_P(VLG_P(VLG_P(VLG_M(R[2], b0), VLG_M(R[5], b1)), VLG_M(R[8], b2)), al[5]);
    }
    const double2 xy = obs_xy[o];
    const double dx = xy.x - x, dy = xy.y - y;
    const bool isbad = z < 0.0 || z > depth_max;
    err[o] = sqrt(dx * dx + dy * dy);
    depth[o] = z;
    bad[o] = isbad ? 1.0 : 0.0;
}

// largest squared error over the observations that pass the depth test, first occurrence in list order
// (remove_outlier uses a strict '>' while walking j outer / i inner = the list order); one CTA.
__global__ void __launch_bounds__(1024) k_argmax_sq(int64_t nobs, const double* __restrict__ err, const double* __restrict__ bad,
                                                    double* __restrict__ out_val, long long* __restrict__ out_idx)
{
    __shared__ double sv[1024];
    __shared__ long long si[1024];
    double best = -1.0;
    long long bi = -1;
    for (int64_t o = threadIdx.x; o < nobs; o += 1024) {
        if (bad[o] != 0.0) continue;
        const double v = err[o] * err[o];
        if (v > best) { best = v; bi = o; }
    }
    sv[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double v = sv[threadIdx.x + s];
            const long long k = si[threadIdx.x + s];
            if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && k >= 0 && (si[threadIdx.x] < 0 || k < si[threadIdx.x]))) {
                sv[threadIdx.x] = v; si[threadIdx.x] = k;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { *out_val = sv[0] < 0.0 ? 0.0 : sv[0]; *out_idx = si[0]; }
}

// ---------------------------------------------------------------------------------------
// Batched independent single-camera problems (estimate_camera.m:247-253 runs bundle_euclid with
// 'fix_structure' on ONE camera; incr_reconstruction.m:223-348 does that once per added camera).  With the structure
// fixed the reduced system is block diagonal -- S_jj = U*_j, e_j = eA_j (mex_bundle_2_Se_.c:80-101,132-155 with Y = 0) --
// so B such problems are one context with B cameras whose LM loops run side by side, each with its own lambda, nu,
// accept decision and stop rule (vlg_ba_solve_cameras_independent).  These kernels are the per-camera pieces.
// ---------------------------------------------------------------------------------------
// U*_j = U_j with its diagonal scaled by (1 + lambda_j)   (bundle_euclid.m:162-167, one lambda per camera)
template <int NA>
__global__ void k_damp_U_vec(int m, const double* __restrict__ lam, const double* __restrict__ U, double* __restrict__ Ud)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * NA * NA) return;
    const int j = t / (NA * NA), r = t - j * NA * NA;
    const double u = U[t];
    Ud[t] = (r % NA == r / NA) ? (1 + lam[j]) * u : u;
}

// da_j = pinv(U*_j) eA_j (bundle_euclid.m:193 on a block-diagonal S) and da_j'(lambda_j da_j + eA_j) (:217 with db = 0);
// cameras that are not active get da_j = 0
template <int NA>
__global__ void k_cam_solve_diag(int m, const double* __restrict__ Ud, const double* __restrict__ eA, const double* __restrict__ lam,
                                 const unsigned char* __restrict__ active, double* __restrict__ da, double* __restrict__ denom)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double M[NA * NA], Mi[NA * NA], g[NA];
#pragma unroll
    for (int k = 0; k < NA * NA; k++) M[k] = Ud[(size_t)NA * NA * j + k];
#pragma unroll
    for (int k = 0; k < NA; k++) g[k] = eA[(size_t)NA * j + k];
    sym_pinv<NA>(M, Mi);
    double dn = 0.0;
#pragma unroll
    for (int r = 0; r < NA; r++) {
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < NA; c++) v += Mi[r + NA * c] * g[c];
        if (!active[j]) v = 0.0;
        da[(size_t)NA * j + r] = v;
        dn += v * (lam[j] * v + g[r]);
    }
    denom[j] = dn;
}

// per-camera sum of a per-observation quantity over the camera's (contiguous, C-order) segment: one warp per camera,
// lanes stride over the segment, fixed xor tree
__global__ void k_cam_seg_sum(int m, const int* __restrict__ cam_ptr, const double* __restrict__ v, double* __restrict__ out)
{
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= m) return;
    double s = 0.0;
    for (int o = cam_ptr[j] + lane; o < cam_ptr[j + 1]; o += 32) s += v[o];
    s = warp_sum(s);
    if (lane == 0) out[j] = s;
}

// ---------------------------------------------------------------------------------------
// self-test of the shared-reciprocal quotients (ba_math.cuh: Den, fd_quot) against __ddiv_rn on random operands:
// class 0: a, d with random significands and exponents in +-200 (d inside the reciprocals' safe window);
// every class also compares the branch-free reciprocal (rcp_fast) with __drcp_rn;
// class 1: depths d in [0.01, 1e4), numerators |a| < 1e7 (what a reprojection divides);
// class 2: a = difference of two nearby image coordinates, d = h (what a forward difference divides).
// Every thread runs `per_thread` samples of a counter-based generator; mismatches are counted bitwise.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long& x)
{
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void k_selftest_quotients(unsigned long long seed, int per_thread, unsigned long long* __restrict__ mismatches)
{
    unsigned long long st = seed + 0x632BE59BD9B4E019ull * ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x + 1);
    unsigned long long bad = 0;
    for (int k = 0; k < per_thread; k++) {
        const unsigned long long u = splitmix64(st), v = splitmix64(st), w = splitmix64(st);
        const int cls = (int)(w % 3);
        double a, d;
        if (cls == 0) {
            const long long ea = 1023 - 200 + (long long)((w >> 8) % 401), ed = 1023 - 200 + (long long)((w >> 24) % 401);
            a = __longlong_as_double((long long)((u & 0x800FFFFFFFFFFFFFull) | ((unsigned long long)ea << 52)));
            d = __longlong_as_double((long long)((v & 0x800FFFFFFFFFFFFFull) | ((unsigned long long)ed << 52)));
        } else if (cls == 1) {
            d = 0.01 * exp2(19.93 * ((double)(v >> 11) * 0x1.0p-53));
            a = 2e7 * ((double)(u >> 11) * 0x1.0p-53) - 1e7;
        } else {
            const double x0 = 500.0 * ((double)(u >> 11) * 0x1.0p-53);
            const double x1 = x0 + (((double)(v >> 11) * 0x1.0p-53) - 0.5) * 1e-6 * (double)(1 + (w >> 40) % 1000);
            a = x1 - x0; d = kFdStep;
        }
        const double ref = __ddiv_rn(a, d);
        const double got = cls == 2 ? fd_quot<true>(a + 0.0, 0.0) : Den<true>(d)(a);
        bad += __double_as_longlong(ref) != __double_as_longlong(got);
        // the branch-free reciprocal against CUDA's own, wherever the callers would accept it
        const double rf = rcp_fast(d);
        if (rcp_in_window(rf)) bad += __double_as_longlong(rf) != __double_as_longlong(__drcp_rn(d));
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace vlgba
