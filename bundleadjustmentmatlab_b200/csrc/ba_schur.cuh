// ba_schur.cuh -- the Schur complement of one camera chunk in one kernel (num_a = 6).
//
// Reference: bundle_euclid.m:178-184 (Y_ij = W_ij V*_i^-1) and mex_bundle_2_Se_.c:72-155
// (S_jk = delta_jk U*_j - sum_i Y_ij W_ik', e_j = eA_j - sum_i Y_ij eB_i).
//
// Round 1 did this in three kernels -- a camera pass that formed Y, the diagonal sums and Y eB with per-lane gathers of
// V*^-1 and eB (0.44 ms at Venice shape), and two pair-list kernels that gathered Y_ij AND W_ik from global memory for
// every pair (1.06 ms) -- and ncu (profiles/ncu_r02c_*) showed what bounds them: not DRAM (40 %) and not FP64 (13 %),
// but L1/LSU wavefronts: a lane-private 144-byte record costs nine 16-byte requests to nine wavefronts.  Here
//   * a CTA owns one tile of one camera's segment (<= kSchurTile observations): the tile's W arrives as ONE TMA
//     bulk copy, the points' (V*^-1 | eB) records (96 bytes, written by k_vinv_damp) by warp-cooperative 16-byte
//     cp.async -- adjacent lanes fetch adjacent pieces of a record, ~1.5 wavefronts per record instead of 12;
//   * lane t forms Y_t (kept in shared memory, and stored to HBM with one TMA bulk store for the light blocks), the
//     chunk's share of S_jj and of sum Y eB;
//   * the chunk's pairs with the cameras k that share MANY points with j (blocks with > kSegHeavy pairs: 2 % of the blocks,
//     87 % of the pairs at Venice shape) are processed right here: Y_ij comes from shared memory, only W_ik is fetched
//     (cooperatively, through the shared memory the W tile no longer needs), a lane accumulates whole pairs and the warp
//     folds its 36 sums by recursive halving into one partial per (block, chunk) segment;
//   * k_schur_fold adds a block's segments in chunk order (fixed shape: bit-reproducible) and stores S_jk and S_kj.
// Blocks with few pairs keep the pair-list kernels of ba_kernels.cuh (they read the Y stored here).
#pragma once
#include "ba_pcg.cuh"

namespace vlgba {

// observations per tile == threads per CTA.  The tiles are this kernel's own cut of the camera segments (not the <= 256-
// observation chunks of stage 1): 128 observations = 55 KB of shared memory and 128 x 128 registers, i.e. FOUR CTAs per
// SM in different phases instead of two (the camera phase and the pair phase of one CTA are serial: round 2, 256-
// observation tiles: 0.99 ms at Venice shape)
#ifndef VLG_SCHUR_TILE
#define VLG_SCHUR_TILE 128
#endif
constexpr int kSchurTile = VLG_SCHUR_TILE;
constexpr int kSchurWarps = kSchurTile / 32;
constexpr int kSegHeavy = 256;      // blocks with more pairs than this are processed by the chunk kernel (measured 64 / 128 / 256 / 512: Schur pieces 1.34 / 1.25 / 1.21 / 1.3 ms)
constexpr int kVE = 12;             // doubles per point record: V*^-1 (9, column-major) | eB (3)

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// 32 records of Q 16-byte units each, record r from gbase + rec_idx(lane r) * 16 Q (skipped when that index is negative),
// into dst + r * 16 Q.  Unit u of the 32 Q units goes to lane u & 31 in round u >> 5: adjacent lanes fetch adjacent
// pieces of the same record.
template <int Q>
__device__ __forceinline__ void warp_gather_records(const char* __restrict__ gbase, int my_idx, char* __restrict__ dst, int lane)
{
#pragma unroll
    for (int u = 0; u < Q; u++) {
        const int id = u * 32 + lane, rec = id / Q, part = id - rec * Q;
        const int src = __shfl_sync(0xffffffffu, my_idx, rec);
        if (src >= 0) cp_async16(dst + ((size_t)rec * Q + part) * 16, gbase + ((size_t)src * Q + part) * 16);
    }
}

// sum over the 32 lanes of NP (a power of two <= 32) values per lane by recursive halving; returns the total of value
// ((lane >> (5 - log2 NP)) in every lane; fixed combination order
template <int NP>
__device__ __forceinline__ double warp_halving_sum(double* v, int lane)
{
    int width = NP;
#pragma unroll
    for (int mask = 16; mask > 0; mask >>= 1) {
        if (width > 1) {
            const bool hi = (lane & mask) != 0;
#pragma unroll
            for (int k = 0; k < NP / 2; k++)
                if (k < width / 2) {
                    const double send = hi ? v[k] : v[k + width / 2];
                    const double keep = hi ? v[k + width / 2] : v[k];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
                }
            width >>= 1;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], mask);
        }
    }
    return v[0];
}

struct SchurChunkArgs {
    const int2* chunk_meta;        // (begin, nob) per chunk
    const int* obs_pt;             // C-order
    const double* W;               // [nobs][18], C-order
    const double* VE;              // [n][12]: V*^-1 | eB
    double* part;                  // [nchunks][27]: the chunk's share of sum Y W' (upper triangle) and sum Y eB
    double* Yout;                  // [nobs][18] or NULL
    const int* chunk_round_ptr;    // [8 nchunks + 1] or NULL (no heavy blocks): rounds of warp w of chunk c = [ptr[8c + w], ptr[8c + w + 1])
    const int4* rounds;            // (first pair, pairs <= 32, slot in Hpart, 1 = flush after this round), grouped by (chunk, warp)
    const int2* pairs;             // (observation of j, observation of k), ascending point inside a block
    double* Hpart;                 // [slots][36]
};

// W tile (later: gather buffers A) | gather buffers B (before: the points' records) | Y tile | warp sums | barrier
constexpr size_t kSchurChunkSmem = sizeof(double) * ((size_t)3 * kSchurTile * 18 + kSchurWarps * 27) + 16;

__global__ void __launch_bounds__(kSchurTile, 65536 / (kSchurTile * 128))
k_schur_chunk(SchurChunkArgs p)
{
    constexpr int NA = 6, NW = 18, NU = 27;
    extern __shared__ __align__(128) unsigned char smraw[];
    double* Wt = reinterpret_cast<double*>(smraw);             // kSchurTile x 18
    double* Bt = Wt + kSchurTile * NW;                         // kSchurTile x 18 (camera phase: kSchurTile x 12 point records)
    double* Yt = Bt + kSchurTile * NW;                         // kSchurTile x 18
    double* red = Yt + kSchurTile * NW;                        // kSchurWarps x 27
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + kSchurWarps * NU);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = blockIdx.x;
    const int2 meta = __ldg(p.chunk_meta + c);
    const int beg = meta.x, nob = meta.y;
    if (tid == 0) {
        mbar_init(bar, 1);
        const uint32_t bytes = (uint32_t)nob * NW * 8u;
        mbar_expect_tx(bar, bytes);
        tma_load_1d(Wt, p.W + (size_t)beg * NW, bytes, bar);
    }
    // the pair phase's first rounds: their descriptors and pair entries are requested now, behind the tile
    int r0 = 0, r1 = 0;
    if (p.chunk_round_ptr) {
        r0 = __ldg(p.chunk_round_ptr + (size_t)kSchurWarps * c + warp);
        r1 = __ldg(p.chunk_round_ptr + (size_t)kSchurWarps * c + warp + 1);
    }
    int4 rd_a = make_int4(0, 0, 0, 0), rd_b = make_int4(0, 0, 0, 0);
    int2 pr_a = make_int2(-1, -1), pr_b = make_int2(-1, -1);
    if (r0 < r1) { rd_a = __ldg(p.rounds + r0); if (lane < rd_a.y) pr_a = __ldg(p.pairs + rd_a.x + lane); }
    if (r0 + 1 < r1) { rd_b = __ldg(p.rounds + r0 + 1); if (lane < rd_b.y) pr_b = __ldg(p.pairs + rd_b.x + lane); }
    // the points' records, gathered by warps
    const int i = tid < nob ? __ldg(p.obs_pt + beg + tid) : -1;
    warp_gather_records<kVE / 2>(reinterpret_cast<const char*>(p.VE), i, reinterpret_cast<char*>(Bt + (size_t)warp * 32 * kVE), lane);
    cp_async_wait_all();
    __syncthreads();                                            // publishes the barrier's initialisation as well
    mbar_wait(bar, 0);

    // ---- camera phase: Y_t, the chunk's share of S_jj and of sum Y eB
    {
        double acc[NU];
#pragma unroll
        for (int t = 0; t < NU; t++) acc[t] = 0.0;
        if (tid < nob) {
            double Wo[NW], Y[NW], ve[kVE];
            load_block<NW>(Wt, tid, Wo);
            load_block<kVE>(Bt, tid, ve);
#pragma unroll
            for (int cc = 0; cc < 3; cc++)
#pragma unroll
                for (int r = 0; r < NA; r++)
                    Y[r + NA * cc] = Wo[r] * ve[3 * cc] + Wo[r + NA] * ve[1 + 3 * cc] + Wo[r + 2 * NA] * ve[2 + 3 * cc];
            double2* yd = reinterpret_cast<double2*>(Yt + (size_t)tid * NW);
#pragma unroll
            for (int k = 0; k < NW / 2; k++) yd[k] = make_double2(Y[2 * k], Y[2 * k + 1]);
#pragma unroll
            for (int col = 0; col < NA; col++)
#pragma unroll
                for (int row = 0; row <= col; row++)
                    acc[col * (col + 1) / 2 + row] = Y[row] * Wo[col] + Y[row + NA] * Wo[col + NA] + Y[row + 2 * NA] * Wo[col + 2 * NA];
#pragma unroll
            for (int r = 0; r < NA; r++) acc[NA * (NA + 1) / 2 + r] = Y[r] * ve[9] + Y[r + NA] * ve[10] + Y[r + 2 * NA] * ve[11];
        }
        // 27 sums over the warp: 16 + 8 by recursive halving (lane l ends up with entry l >> 1 resp. 16 + (l >> 2)), 3 by xor trees
        const double s0 = warp_halving_sum<16>(acc, lane);
        const double s1 = warp_halving_sum<8>(acc + 16, lane);
        const double s2 = warp_sum(acc[24]), s3 = warp_sum(acc[25]), s4 = warp_sum(acc[26]);
        if ((lane & 1) == 0) red[warp * NU + (lane >> 1)] = s0;
        if ((lane & 3) == 0) red[warp * NU + 16 + (lane >> 2)] = s1;
        if (lane == 0) { red[warp * NU + 24] = s2; red[warp * NU + 25] = s3; red[warp * NU + 26] = s4; }
    }
    __syncthreads();                                            // Y tile complete; the W tile and the records are free from here on
    if (tid < NU) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kSchurWarps; w++) s += red[w * NU + tid];
        p.part[(size_t)NU * c + tid] = s;
    }
    if (p.Yout && tid == 0) {
        // one bulk store of the chunk's Y (generic-proxy writes made visible to the async proxy first)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.Yout + (size_t)beg * NW), "r"(smem_u32(Yt)),
                     "r"((uint32_t)nob * NW * 8u)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }

    // ---- pair phase.  The chunk's pairs with the heavy cameras k, cut at set-up time into ROUNDS of <= 32 pairs of one
    // block; warp w takes the w-th eighth of the chunk's rounds.  Two-deep software pipeline: while round r is multiplied,
    // the W_ik records of round r+1 are on their way into the other gather buffer (cp.async) and the pair entries of round
    // r+2 into registers.  A lane accumulates whole pairs; after a block's last round (in this warp) the 36 sums are folded
    // over the lanes and stored as one partial.
    if (r0 < r1) {
        char* stA = reinterpret_cast<char*>(Wt + (size_t)warp * 32 * NW);
        char* stB = reinterpret_cast<char*>(Bt + (size_t)warp * 32 * NW);
        double a[NA * NA];
#pragma unroll
        for (int u = 0; u < NA * NA; u++) a[u] = 0.0;
        warp_gather_records<NW / 2>(reinterpret_cast<const char*>(p.W), pr_a.y, stA, lane);
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int r = r0; r < r1; r++) {
            char* cur = ((r - r0) & 1) ? stB : stA;
            char* nxt = ((r - r0) & 1) ? stA : stB;
            // round r+1: gather; round r+2: descriptor and pair entries
            if (r + 1 < r1) warp_gather_records<NW / 2>(reinterpret_cast<const char*>(p.W), pr_b.y, nxt, lane);
            asm volatile("cp.async.commit_group;" ::: "memory");
            int4 rd_c = make_int4(0, 0, 0, 0);
            int2 pr_c = make_int2(-1, -1);
            if (r + 2 < r1) { rd_c = __ldg(p.rounds + r + 2); if (lane < rd_c.y) pr_c = __ldg(p.pairs + rd_c.x + lane); }
            asm volatile("cp.async.wait_group 1;" ::: "memory");      // round r's records have landed
            __syncwarp();
            if (lane < rd_a.y) {
                const double* yj = Yt + (size_t)(pr_a.x - beg) * NW;
                const double* wk = reinterpret_cast<const double*>(cur) + (size_t)lane * NW;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    double y[NA], w[NA];
#pragma unroll
                    for (int h = 0; h < NA / 2; h++) {
                        const double2 yv = reinterpret_cast<const double2*>(yj + NA * d)[h], wv = reinterpret_cast<const double2*>(wk + NA * d)[h];
                        y[2 * h] = yv.x; y[2 * h + 1] = yv.y; w[2 * h] = wv.x; w[2 * h + 1] = wv.y;
                    }
#pragma unroll
                    for (int col = 0; col < NA; col++)
#pragma unroll
                        for (int row = 0; row < NA; row++) a[row + NA * col] += y[row] * w[col];
                }
            }
            if (rd_a.w) {
                // the segment's last round: fold the 36 sums over the lanes through the gather buffer that has just been read
                // out (16 rows of 36 doubles) -- two halves of the warp in turn; lane e adds entry e over the rows in row
                // order, entries 32..35 are split over lane groups and finished by three xor steps.  (The recursive-halving
                // shuffle tree cost ~400 instructions per flush, this ~130.)
                double* fb = reinterpret_cast<double*>(cur);
                double t0 = 0.0, t1 = 0.0;
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    __syncwarp();
                    if ((lane >> 4) == half) {
                        double2* row = reinterpret_cast<double2*>(fb + (size_t)(lane & 15) * (NA * NA));
#pragma unroll
                        for (int u = 0; u < NA * NA / 2; u++) row[u] = make_double2(a[2 * u], a[2 * u + 1]);
                    }
                    __syncwarp();
                    double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
                    for (int rr = 0; rr < 16; rr += 4) {
                        c0 += fb[(size_t)rr * (NA * NA) + lane]; c1 += fb[(size_t)(rr + 1) * (NA * NA) + lane];
                        c2 += fb[(size_t)(rr + 2) * (NA * NA) + lane]; c3 += fb[(size_t)(rr + 3) * (NA * NA) + lane];
                    }
                    t0 += (c0 + c1) + (c2 + c3);
                    // entry 32 + (lane & 3), rows 2 (lane >> 2) and 2 (lane >> 2) + 1
                    t1 += fb[(size_t)(2 * (lane >> 2)) * (NA * NA) + 32 + (lane & 3)] + fb[(size_t)(2 * (lane >> 2) + 1) * (NA * NA) + 32 + (lane & 3)];
                }
                t1 += __shfl_xor_sync(0xffffffffu, t1, 4); t1 += __shfl_xor_sync(0xffffffffu, t1, 8); t1 += __shfl_xor_sync(0xffffffffu, t1, 16);
                double* out = p.Hpart + (size_t)rd_a.z * (NA * NA);
                out[lane] = t0;
                if (lane < 4) out[32 + lane] = t1;
#pragma unroll
                for (int u = 0; u < NA * NA; u++) a[u] = 0.0;
            }
            __syncwarp();                                       // `cur` is read out: the round after next may overwrite it
            rd_a = rd_b; pr_a = pr_b; rd_b = rd_c; pr_b = pr_c;
        }
    }
    if (p.Yout && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // Yt must outlive the bulk store's reads
}

// ---------------------------------------------------------------------------------------
// Back-substitution (mex_bundle_3_db_new.c:100-146) with one lane per observation over tiles of whole points, as
// k_backsub_tiled -- the same operations in the same order, bit-identical -- but the W records (C-order, 144 bytes,
// scattered when walked in point order) arrive by warp-cooperative cp.async instead of nine lane-private 16-byte
// gathers: ncu of the round-1 kernel showed 0.41 ms at 3 % FP64 / 5 % issue utilisation, all of it LSU wavefronts.
// ---------------------------------------------------------------------------------------
template <int NA>
__global__ void __launch_bounds__(kS1Tile)
k_backsub_coop(const int4* __restrict__ ptile_meta, const int* __restrict__ pt_ptr, const int* __restrict__ pt_obs,
               const int* __restrict__ pt_cam, const double* __restrict__ W, const double* __restrict__ Vinv,
               const double* __restrict__ eB, const double* __restrict__ da, const double* __restrict__ b, double lambda,
               int all_rows, double* __restrict__ db, double* __restrict__ b_new, double* __restrict__ denom_pt)
{
    constexpr int NW = 3 * NA;
    static_assert(NW % 2 == 0, "16-byte units");
    extern __shared__ __align__(128) unsigned char smraw[];
    double* wt = reinterpret_cast<double*>(smraw);                 // kS1Tile x NW, one 32-record slab per warp
    __shared__ double pr[3][kS1Tile];
    const int4 meta = __ldg(ptile_meta + blockIdx.x);
    const int q0 = meta.x, nob = meta.y, p0 = meta.z, np = meta.w, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nrow = all_rows ? NA : 6;
    if (warp * 32 < nob) {
        const int o = tid < nob ? __ldg(pt_obs + q0 + tid) : -1;
        const int j = tid < nob ? __ldg(pt_cam + q0 + tid) : 0;
        double* slab = wt + (size_t)warp * 32 * NW;
        warp_gather_records<NW / 2>(reinterpret_cast<const char*>(W), o, reinterpret_cast<char*>(slab), lane);
        double dj[NA];
#pragma unroll
        for (int r = 0; r < NA; r++) dj[r] = __ldg(da + (size_t)NA * j + r);
        cp_async_wait_all();
        __syncwarp();
        if (tid < nob) {
            double Wo[NW];
            load_block<NW>(slab, lane, Wo);
            double s0 = VLG_M(Wo[0], dj[0]), s1 = VLG_M(Wo[NA], dj[0]), s2 = VLG_M(Wo[2 * NA], dj[0]);
#pragma unroll
            for (int r = 1; r < NA; r++) {
                if (r < nrow) {
                    s0 = VLG_P(s0, VLG_M(Wo[r], dj[r]));
                    s1 = VLG_P(s1, VLG_M(Wo[r + NA], dj[r]));
                    s2 = VLG_P(s2, VLG_M(Wo[r + 2 * NA], dj[r]));
                }
            }
            pr[0][tid] = s0; pr[1][tid] = s1; pr[2][tid] = s2;
        }
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

// S_jk (and S_kj) of the heavy blocks: the block's segments added in chunk order
template <int NA>
__global__ void __launch_bounds__(128)
k_schur_fold(int nsblk, const int* __restrict__ sblk, const int* __restrict__ hptr, int ld, int ccams, const int* __restrict__ blk_j,
             const int* __restrict__ blk_k, const double* __restrict__ Hpart, double* __restrict__ S)
{
    constexpr int VV = NA * NA;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= nsblk) return;
    const int b = sblk[t], j = blk_j[b], k = blk_k[b];
    const int h0 = hptr[t], h1 = hptr[t + 1];
    const SchurDest d = schur_dest<NA>(S, ld, ccams, j, k);
    for (int e = lane; e < VV; e += 32) {
        double s = 0.0;
        for (int h = h0; h < h1; h++) s += __ldg(Hpart + (size_t)h * VV + e);
        const int row = e % NA, col = e / NA;
        d.base[(size_t)(d.jr + row) + d.ld * (d.kr + col)] = -s;
        d.base[(size_t)(d.kr + col) + d.ld * (d.jr + row)] = -s;
    }
}

}  // namespace vlgba
