// tools/stream_probe.cu -- what read bandwidth can a "sum-reduce a 720 MB stream" kernel reach on
// this GPU, as a function of how the stream is pulled in?  Evidence for the tile design of
// ba_pcg.cuh.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/stream_probe tools/stream_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

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
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// (1) plain grid-stride LDG.128
__global__ void k_ldg(const double2* __restrict__ in, size_t n2, double* out)
{
    double s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 v = in[i];
        s += v.x + v.y;
    }
    if (s == 1.2345) out[0] = s;
}

// (2) one CTA per tile, `nsplit` bulk copies per tile
template <int THREADS>
__global__ void k_tma_tile(const double* __restrict__ in, int tile_doubles, int nsplit, double* out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    double* t = (double*)sm;
    uint64_t* bar = (uint64_t*)(t + tile_doubles);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, tile_doubles * 8);
        const int part = tile_doubles / nsplit;
        for (int k = 0; k < nsplit; k++) tma_load_1d(t + k * part, in + (size_t)blockIdx.x * tile_doubles + k * part, part * 8, bar);
    }
    mbar_wait(bar, 0);
    double s = 0;
    for (int i = threadIdx.x; i < tile_doubles; i += THREADS) s += t[i];
    if (s == 1.2345) out[0] = s;
}

// (3) persistent CTAs, ring of `stages` tiles
template <int THREADS>
__global__ void k_tma_persist(const double* __restrict__ in, int ntiles, int tile_doubles, int stages, double* out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    double* t = (double*)sm;
    uint64_t* bar = (uint64_t*)(t + (size_t)stages * tile_doubles);
    if (threadIdx.x == 0) for (int s = 0; s < stages; s++) mbar_init(bar + s, 1);
    __syncthreads();
    double acc = 0;
    int issued = 0, done = 0;
    const int first = blockIdx.x, step = gridDim.x;
    const int mine = first < ntiles ? (ntiles - first + step - 1) / step : 0;
    for (; issued < stages && issued < mine; issued++)
        if (threadIdx.x == 0) { mbar_expect_tx(bar + issued, tile_doubles * 8); tma_load_1d(t + (size_t)issued * tile_doubles, in + (size_t)(first + issued * step) * tile_doubles, tile_doubles * 8, bar + issued); }
    for (; done < mine; done++) {
        const int slot = done % stages;
        mbar_wait(bar + slot, (done / stages) & 1);
        const double* tt = t + (size_t)slot * tile_doubles;
        for (int i = threadIdx.x; i < tile_doubles; i += THREADS) acc += tt[i];
        __syncthreads();
        if (issued < mine) {
            if (threadIdx.x == 0) { mbar_expect_tx(bar + slot, tile_doubles * 8); tma_load_1d(t + (size_t)slot * tile_doubles, in + (size_t)(first + issued * step) * tile_doubles, tile_doubles * 8, bar + slot); }
            issued++;
        }
    }
    if (acc == 1.2345) out[0] = acc;
}

template <class F>
float timeit(F f, int reps = 20)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main()
{
    const size_t bytes = 720ull << 20;
    const size_t n = bytes / 8;
    double *in, *out;
    cudaMalloc(&in, bytes); cudaMalloc(&out, 8);
    cudaMemset(in, 0, bytes);
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    printf("stream of %zu MB, %d SMs\n", bytes >> 20, nsm);
    for (int mult : {8, 16, 32}) {
        float ms = timeit([&] { k_ldg<<<nsm * mult, 256>>>((const double2*)in, n / 2, out); });
        printf("ldg.128 grid=%dxSM              : %.3f ms  %.0f GB/s\n", mult, ms, bytes / ms / 1e6);
    }
    for (int tile_kb : {9, 18, 36, 72})
        for (int nsplit : {1, 4}) {
            const int td = tile_kb * 1024 / 8;
            const int ntiles = (int)(n / td);
            const size_t smem = (size_t)td * 8 + 16;
            cudaFuncSetAttribute(k_tma_tile<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            float ms = timeit([&] { k_tma_tile<256><<<ntiles, 256, smem>>>(in, td, nsplit, out); });
            printf("tma tile %2d KB x%d, CTA/tile     : %.3f ms  %.0f GB/s\n", tile_kb, nsplit, ms, (double)ntiles * td * 8 / ms / 1e6);
        }
    for (int tile_kb : {18, 36})
        for (int stages : {2, 3, 4})
            for (int cps : {1, 2}) {
                const int td = tile_kb * 1024 / 8;
                const int ntiles = (int)(n / td);
                const size_t smem = (size_t)stages * td * 8 + 64;
                if (smem * cps > 220 * 1024) continue;
                cudaFuncSetAttribute(k_tma_persist<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                float ms = timeit([&] { k_tma_persist<256><<<nsm * cps, 256, smem>>>(in, ntiles, td, stages, out); });
                printf("tma persistent %2d KB, %d stages, %d CTA/SM: %.3f ms  %.0f GB/s\n", tile_kb, stages, cps, ms, (double)ntiles * td * 8 / ms / 1e6);
            }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
