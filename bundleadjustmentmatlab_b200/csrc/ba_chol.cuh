// ba_chol.cuh -- the dense reduced-system solve da = pinv(S) e_ (bundle_euclid.m:193) for small
// camera counts: ONE cooperative kernel that factors S = L L' (right-looking, 32-wide panels),
// carries the right-hand side through the factorisation as an extra row tile (forward solve
// for free) and finishes with the backward solve.
//
// Why one kernel: at Ladybug/Trafalgar size the factorisation is a chain of ~nb dependent panel
// steps of a few microseconds each; launched as potrf/trsm/syrk kernels it costs 3 launches per
// panel (round 1a: 0.51 ms at 294 unknowns, 3.0 ms at 1542, of which >90 % launch latency).  Here a
// panel step costs one grid barrier:
//   * look-ahead: in step kb every CTA redundantly forms and factors the NEXT diagonal tile
//     (A[kb+1][kb+1] - X X', 32 x 32, in shared memory), and the owner of row tile i brings
//     A[i][kb+1] up to date and solves it against that factor, so panel kb+1 is final when the
//     step's single grid barrier falls;
//   * the trailing update C[i][j] -= X_i X_j' runs on the FP64 tensor pipe
//     (mma.sync.aligned.m8n8k4.f64 -- tcgen05 has no f64 kind), one 32 x 32 tile per warp,
//     operands straight from L2 in fragment layout;
//   * non-positive pivots are eliminated (row/column of L set to zero, solution component 0):
//     the exactly-zero rows/columns of S (camera 0 has omega = 0, fix_* options) get pinv's answer.
// Everything is fixed-order => bit-reproducible.
#pragma once
#include "ba_math.cuh"
#include <cooperative_groups.h>

namespace vlgba {

constexpr int kNB = 32;
constexpr int kCholWarps = 8;
constexpr int kLsLd = kNB + 1;

struct CholArgs {
    double* S;       // Np x Np, column-major, lower triangle (tiles i >= j) is factored in place
    int ld;
    int nb;          // Np / 32
    int N;           // unknowns (<= Np)
    const double* rhs;   // N
    double* R;       // nb tiles of 32 x 32 (ld 32): the right-hand side lives in row 0 of tile j
    double* Ld;      // nb factored diagonal tiles, 32 x 32 (ld 32), upper part zero
    double* Dinv;    // Np reciprocals of the factor's diagonal (0 = eliminated pivot)
    unsigned int* barrier;   // grid barrier counter, zero at launch
    long long* prof;         // optional (tools/chol_probe.cu): clock64 totals per phase seen by CTA 0 thread 0, or NULL
    double* x;       // N: solution
};

__device__ __forceinline__ double* chol_tile(const CholArgs& p, int i, int j, int& ld)
{
    if (i < p.nb) { ld = p.ld; return p.S + (size_t)kNB * i + (size_t)p.ld * kNB * j; }
    ld = kNB;
    return p.R + (size_t)kNB * kNB * j;
}

// D = C - A * B for m8n8k4: A 8x4 (row), B 4x8 (col), C/D 8x8
__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// C (32 x 32) -= Xi (32 x 32) * Xj (32 x 32)', one warp, fragments straight from global/L2
__device__ __forceinline__ void tile_update_dmma(double* __restrict__ C, int ldc, const double* __restrict__ Xi, int ldi,
                                                 const double* __restrict__ Xj, int ldj, int lane)
{
    const int g = lane >> 2, q = lane & 3;
    double a[4][8], b[4][8], c[4][4][2];
#pragma unroll
    for (int ks = 0; ks < 8; ks++)
#pragma unroll
        for (int mb = 0; mb < 4; mb++) {
            a[mb][ks] = -__ldcg(Xi + (8 * mb + g) + (size_t)ldi * (4 * ks + q));
            b[mb][ks] = __ldcg(Xj + (8 * mb + g) + (size_t)ldj * (4 * ks + q));
        }
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nbk = 0; nbk < 4; nbk++)
#pragma unroll
            for (int e = 0; e < 2; e++) c[mb][nbk][e] = __ldcg(C + (8 * mb + g) + (size_t)ldc * (8 * nbk + 2 * q + e));
#pragma unroll
    for (int ks = 0; ks < 8; ks++)
#pragma unroll
        for (int mb = 0; mb < 4; mb++)
#pragma unroll
            for (int nbk = 0; nbk < 4; nbk++) dmma_884(c[mb][nbk][0], c[mb][nbk][1], a[mb][ks], b[nbk][ks]);
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nbk = 0; nbk < 4; nbk++)
#pragma unroll
            for (int e = 0; e < 2; e++) C[(8 * mb + g) + (size_t)ldc * (8 * nbk + 2 * q + e)] = c[mb][nbk][e];
}

// 1/d and 1/sqrt(d) for d > 0 from the hardware's double-precision seeds (MUFU.RCP64H / RSQ64H,
// ~20 bits) and two Newton steps: branch-free, 4-5 dependent FP64 operations.  d <= 0 gives 0
// (eliminated pivot).  FP64 operations have ~30-cycle latency on this part, so the number of
// DEPENDENT operations per pivot is what the factorisation's critical path is made of.
__device__ __forceinline__ double fast_rcp_pos(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return d > 0.0 ? r : 0.0;
}

__device__ __forceinline__ double fast_rsqrt_pos(double d)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double hd = 0.5 * d;
    r = r * fma(-hd * r, r, 1.5);
    r = r * fma(-hd * r, r, 1.5);
    return d > 0.0 ? r : 0.0;
}

// Cholesky of a 32 x 32 block by one warp: lane r holds row r (a[c], c <= r used).  Pivots <= 0
// eliminate their row and column.  On return a[c] = L[r][c] for c <= r and 0 above; the return value
// is 1 / L[r][r] (0 for an eliminated pivot).
// The chain of 32 dependent pivots is the critical path of the whole factorisation.  It runs in
// LDL' form -- per pivot: reciprocal of the diagonal entry (own scalar of lane j), one shuffle,
// t = u / d, own diagonal entry -= t * u -- with the columns left unscaled and broadcast through
// shared memory (col: 2 x 32 doubles) off the chain; the 32 reciprocal square roots that turn
// L D L' into L L' are taken at the end, all at once.
__device__ __forceinline__ double potrf_warp(double (&a)[kNB], int lane, double* __restrict__ col)
{
    double dg = 0.0, dpiv = 0.0;
#pragma unroll
    for (int c = 0; c < kNB; c++)
        if (lane == c) dg = a[c];
#pragma unroll
    for (int j = 0; j < kNB; j++) {
        double* cb = col + (j & 1) * kNB;
        cb[lane] = lane > j ? a[j] : 0.0;                       // column j, unscaled (known before 1/d_j is)
        const double rcp = __shfl_sync(0xffffffffu, fast_rcp_pos(dg), j);
        if (lane == j) dpiv = dg;
        const double t = lane > j ? a[j] * rcp : 0.0;           // u_rj / d_j
        dg = lane > j ? fma(-t, a[j], dg) : 1.0;                // own diagonal entry; finished lanes keep a benign value
        if (lane < j) a[j] = 0.0;
#pragma unroll
        for (int c = 0; c < kNB; c++)
            if (c > j) a[c] = fma(-t, cb[c], a[c]);             // lanes <= j have t = 0; entries above the diagonal are junk, never read
    }
    // L = U diag(1/sqrt(d)): lane c publishes its 1/sqrt(d_c)
    const double rs = fast_rsqrt_pos(dpiv);
    __syncwarp();
    col[lane] = rs;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < kNB; c++) {
        const double v = a[c] * col[c];
        a[c] = lane > c ? v : (lane == c ? dpiv * rs : 0.0);
    }
    return rs;
}

// x (row of 32, one per lane) <- x * L^-T with L in shared memory (ld kLsLd) and the reciprocals
// of its diagonal (0 for an eliminated pivot => solution component 0)
__device__ __forceinline__ void trsm_row(double (&x)[kNB], const double* __restrict__ Ls, const double* __restrict__ dinv)
{
#pragma unroll
    for (int c = 0; c < kNB; c++) {
        x[c] = x[c] * dinv[c];
#pragma unroll
        for (int c2 = 0; c2 < kNB; c2++)
            if (c2 > c) x[c2] -= x[c] * Ls[c2 * kLsLd + c];
    }
}

// grid-wide barrier for the co-resident CTAs of a cooperative launch: one release-add and an
// acquire spin per CTA on a monotonically increasing counter (zeroed by the host before the launch)
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, unsigned int nctas)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nctas;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// A[i][kn] - X[i][kb] X[kn][kb]' for row `lane` of row tile i, into registers
__device__ __forceinline__ void row_tile_update(const CholArgs& p, int i, int kn, int kb, const double* __restrict__ Xs,
                                                int lane, double (&x)[kNB])
{
    int ldt;
    const double* T = chol_tile(p, i, kn, ldt);
#pragma unroll
    for (int c = 0; c < kNB; c++) x[c] = __ldcg(T + lane + (size_t)ldt * c);
    if (kb >= 0) {
        int ldx;
        const double* Xi = chol_tile(p, i, kb, ldx);
        double xi[kNB];
#pragma unroll
        for (int k = 0; k < kNB; k++) xi[k] = __ldcg(Xi + lane + (size_t)ldx * k);
#pragma unroll
        for (int c = 0; c < kNB; c++) {
            double s = x[c];
#pragma unroll
            for (int k = 0; k < kNB; k++) s -= xi[k] * Xs[c * kLsLd + k];
            x[c] = s;
        }
    }
}

__device__ __forceinline__ void row_tile_store(const CholArgs& p, int i, int kn, int lane, const double (&x)[kNB])
{
    int ldt;
    double* T = chol_tile(p, i, kn, ldt);
#pragma unroll
    for (int c = 0; c < kNB; c++) T[lane + (size_t)ldt * c] = x[c];
}

#define CHOL_PROF(slot)                                                          \
    do {                                                                         \
        if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) {                     \
            const long long now_ = clock64();                                    \
            p.prof[slot] += now_ - prof_t;                                       \
            prof_t = now_;                                                       \
        }                                                                        \
    } while (0)

__global__ void __launch_bounds__(kCholWarps * 32, 1) k_chol_coop(CholArgs p)
{
    long long prof_t = p.prof ? clock64() : 0;
    __shared__ double Xs[kNB * kLsLd];     // X[kb+1][kb] (row-major with pad): Xs[c * kLsLd + k]
    __shared__ double Ls[kNB * kLsLd];     // diagonal tile being formed / its factor: Ls[r * kLsLd + c]
    __shared__ double dinv[kNB];           // reciprocal diagonal of the factor
    __shared__ double xsol[kNB];
    __shared__ __align__(16) double pcol[2 * kNB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x, nb = p.nb;
    const int gw = cta * kCholWarps + warp, nwarps = G * kCholWarps;
    unsigned int bar_target = 0;

    // right-hand side into row 0 of the extra row tile
    for (size_t t = (size_t)cta * blockDim.x + tid; t < (size_t)nb * kNB * kNB; t += (size_t)G * blockDim.x) {
        const int j = (int)(t / (kNB * kNB)), rem = (int)(t % (kNB * kNB)), row = rem % kNB, c = rem / kNB;
        const int u = kNB * j + c;
        p.R[t] = (row == 0 && u < p.N) ? p.rhs[u] : 0.0;
    }
    grid_barrier(p.barrier, bar_target, G);
    CHOL_PROF(0);

    // ---- factorisation: step kb = -1 is the prologue (no previous panel), steps 0..nb-2 carry a trailing update
    for (int kb = -1; kb < nb - 1; kb++) {
        const int kn = kb + 1;                       // panel finished by this step
        // (1) X[kn][kb] into shared memory; this warp's 4 columns of the next diagonal tile are requested
        //     at the same time (they do not depend on the panel)
        double dv[kNB / kCholWarps];
        {
            const double* A = p.S + (size_t)kNB * kn + (size_t)p.ld * kNB * kn;
#pragma unroll
            for (int cc = 0; cc < kNB / kCholWarps; cc++)
                dv[cc] = __ldcg(A + lane + (size_t)p.ld * (warp * (kNB / kCholWarps) + cc));
        }
        if (kb >= 0) {
            const double* X = p.S + (size_t)kNB * kn + (size_t)p.ld * kNB * kb;
            for (int t = tid; t < kNB * kNB; t += blockDim.x) {
                const int r = t % kNB, k = t / kNB;
                Xs[r * kLsLd + k] = __ldcg(X + r + (size_t)p.ld * k);
            }
        }
        __syncthreads();
        CHOL_PROF(1);
        // (2) the next diagonal tile, brought up to date: every CTA, 4 columns per warp
        if (kb >= 0) {
#pragma unroll 8
            for (int k = 0; k < kNB; k++) {
                const double xr = Xs[lane * kLsLd + k];
#pragma unroll
                for (int cc = 0; cc < kNB / kCholWarps; cc++) dv[cc] -= xr * Xs[(warp * (kNB / kCholWarps) + cc) * kLsLd + k];
            }
        }
#pragma unroll
        for (int cc = 0; cc < kNB / kCholWarps; cc++) Ls[lane * kLsLd + warp * (kNB / kCholWarps) + cc] = dv[cc];
        __syncthreads();
        CHOL_PROF(2);
        // (3) warp 0 factors it.  Meanwhile warps 1..7 bring their row tile of panel kn up to date (the
        //     q-th row tile of this CTA belongs to warp 1 + q % 7) and run the trailing update with
        //     panel kb: tiles (i, j), kn < j < nb, j <= i <= nb, dealt round-robin to the non-factoring warps
        int i_mine = -1;
        if (warp > 0) {
            const int i = kn + 1 + cta + G * (warp - 1);
            if (i <= nb) i_mine = i;
        }
        double x[kNB];
        if (warp == 0) {
#pragma unroll
            for (int c = 0; c < kNB; c++) x[c] = Ls[lane * kLsLd + c];
            const double inv = potrf_warp(x, lane, pcol);
#pragma unroll
            for (int c = 0; c < kNB; c++) Ls[lane * kLsLd + c] = x[c];
            dinv[lane] = inv;
            if (cta == 0) {
#pragma unroll
                for (int c = 0; c < kNB; c++) p.Ld[(size_t)kNB * kNB * kn + lane + kNB * c] = x[c];
                p.Dinv[kNB * kn + lane] = inv;
            }
        } else {
            if (i_mine >= 0) row_tile_update(p, i_mine, kn, kb, Xs, lane, x);
            if (kb >= 0) {
                int j = kn + 1, cnt = nb - j + 1;
                for (int t = cta * (kCholWarps - 1) + warp - 1; j < nb; t += G * (kCholWarps - 1)) {
                    while (j < nb && t >= cnt) { t -= cnt; j++; cnt--; }
                    if (j >= nb) break;
                    const int i = j + t;
                    int ldc, ldi, ldj;
                    double* C = chol_tile(p, i, j, ldc);
                    const double* Xi = chol_tile(p, i, kb, ldi);
                    const double* Xj = chol_tile(p, j, kb, ldj);
                    tile_update_dmma(C, ldc, Xi, ldi, Xj, ldj, lane);
                }
            }
        }
        CHOL_PROF(3);
        __syncthreads();
        CHOL_PROF(4);
        // (4) solve against the new factor; further row tiles of this warp follow one by one
        if (i_mine >= 0) {
            trsm_row(x, Ls, dinv);
            row_tile_store(p, i_mine, kn, lane, x);
            for (int i = i_mine + G * (kCholWarps - 1); i <= nb; i += G * (kCholWarps - 1)) {
                row_tile_update(p, i, kn, kb, Xs, lane, x);
                trsm_row(x, Ls, dinv);
                row_tile_store(p, i, kn, lane, x);
            }
        }
        CHOL_PROF(5);
        grid_barrier(p.barrier, bar_target, G);
        CHOL_PROF(6);
    }

    // ---- backward solve L' x = y, y in row 0 of the extra row tile.  Everything that does not depend on
    //      x_kb (the factor tiles, the y entries to be updated) is requested before the triangular solve.
    double ldpre[kNB * kNB / (kCholWarps * 32)];
    double dipre = 0.0;
    {
        const int kb = nb - 1;
#pragma unroll
        for (int u = 0; u < kNB * kNB / (kCholWarps * 32); u++) ldpre[u] = __ldcg(p.Ld + (size_t)kNB * kNB * kb + tid + u * kCholWarps * 32);
        if (tid < kNB) dipre = __ldcg(p.Dinv + kNB * kb + tid);
    }
    for (int kb = nb - 1; kb >= 0; kb--) {
#pragma unroll
        for (int u = 0; u < kNB * kNB / (kCholWarps * 32); u++) {
            const int t = tid + u * kCholWarps * 32, r = t % kNB, c = t / kNB;
            Ls[r * kLsLd + c] = ldpre[u];
        }
        if (tid < kNB) dinv[tid] = dipre;
        // this warp's first tile of row kb and its y
        const int j0 = gw;
        double2 lt[kNB / 2];
        double yj0 = 0.0;
        if (j0 < kb) {
            const double* Lt = p.S + (size_t)kNB * kb + (size_t)p.ld * (kNB * j0 + lane);
#pragma unroll
            for (int r = 0; r < kNB / 2; r++) lt[r] = __ldcg(reinterpret_cast<const double2*>(Lt) + r);
            yj0 = __ldcg(p.R + (size_t)kNB * kNB * j0 + kNB * lane);
        }
        double yk = 0.0;
        if (warp == 0) yk = __ldcg(p.R + (size_t)kNB * kNB * kb + kNB * lane);      // lane = column c
        if (kb > 0) {
#pragma unroll
            for (int u = 0; u < kNB * kNB / (kCholWarps * 32); u++) ldpre[u] = __ldcg(p.Ld + (size_t)kNB * kNB * (kb - 1) + tid + u * kCholWarps * 32);
            if (tid < kNB) dipre = __ldcg(p.Dinv + kNB * (kb - 1) + tid);
        }
        __syncthreads();
        CHOL_PROF(7);
        if (warp == 0) {
            double y = yk;
#pragma unroll
            for (int r = kNB - 1; r >= 0; r--) {
                const double xr = __shfl_sync(0xffffffffu, y, r) * dinv[r];
                if (lane == r) y = xr;
                else if (lane < r) y -= Ls[r * kLsLd + lane] * xr;
            }
            xsol[lane] = y;
            if (cta == 0 && kNB * kb + lane < p.N) p.x[kNB * kb + lane] = y;
        }
        __syncthreads();
        CHOL_PROF(8);
        // y_j -= L[kb][j]' x_kb for j < kb: lane = column, its 32 rows are contiguous
        if (j0 < kb) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < kNB / 2; r++) s += lt[r].x * xsol[2 * r] + lt[r].y * xsol[2 * r + 1];
            p.R[(size_t)kNB * kNB * j0 + kNB * lane] = yj0 - s;
        }
        for (int j = j0 + nwarps; j < kb; j += nwarps) {
            const double* Lt = p.S + (size_t)kNB * kb + (size_t)p.ld * (kNB * j + lane);
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < kNB; r += 2) {
                const double2 v = __ldcg(reinterpret_cast<const double2*>(Lt + r));
                s += v.x * xsol[r] + v.y * xsol[r + 1];
            }
            double* yj = p.R + (size_t)kNB * kNB * j + kNB * lane;
            *yj = __ldcg(yj) - s;
        }
        CHOL_PROF(9);
        if (kb > 0) grid_barrier(p.barrier, bar_target, G);
        CHOL_PROF(10);
    }
}

}  // namespace vlgba
