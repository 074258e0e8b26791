// tools/chol_probe.cu -- k_chol_coop on a random SPD system: residual check against the right-hand
// side, wall time, and where the time goes (clock64 totals per phase as seen by CTA 0 / thread 0).
// Evidence for the design notes of csrc/ba_chol.cuh.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/chol_probe tools/chol_probe.cu
// Run:   tools/chol_probe [N ...]      (defaults: 294 1542 3072)
#include "../bundleadjustmentmatlab_b200/csrc/ba_chol.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace vlgba;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int run(int N, int nsm, int clock_khz)
{
    const int Np = (N + kNB - 1) / kNB * kNB, nb = Np / kNB;
    // S = B B' + N I on the first N rows/cols, rows 3..5 zeroed (eliminated pivots), padding zero
    std::vector<double> S((size_t)Np * Np, 0.0), rhs(N), Bm((size_t)N * 8);
    srand(1);
    for (auto& v : Bm) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            double s = i == j ? 1.0 + 0.01 * i : 0.0;
            for (int k = 0; k < 8; k++) s += Bm[(size_t)i * 8 + k] * Bm[(size_t)j * 8 + k];
            if ((i >= 3 && i < 6) || (j >= 3 && j < 6)) s = 0.0;
            S[i + (size_t)Np * j] = s;
        }
    for (int i = 0; i < N; i++) rhs[i] = (i >= 3 && i < 6) ? 0.0 : sin(0.1 * i);
    double *dS, *dS0, *dR, *dLd, *dDinv, *dx, *drhs;
    unsigned int* dbar;
    long long* dprof;
    CK(cudaMalloc(&dS, sizeof(double) * Np * Np)); CK(cudaMalloc(&dS0, sizeof(double) * Np * Np));
    CK(cudaMalloc(&dR, sizeof(double) * nb * kNB * kNB)); CK(cudaMalloc(&dLd, sizeof(double) * nb * kNB * kNB));
    CK(cudaMalloc(&dDinv, sizeof(double) * Np)); CK(cudaMalloc(&dx, sizeof(double) * Np)); CK(cudaMalloc(&drhs, sizeof(double) * N));
    CK(cudaMalloc(&dbar, 4)); CK(cudaMalloc(&dprof, 16 * sizeof(long long)));
    CK(cudaMemcpy(dS0, S.data(), sizeof(double) * Np * Np, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(drhs, rhs.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
    const long long tiles = (long long)nb * (nb + 1) / 2;
    const int G = (int)std::max<long long>(1, std::min<long long>(nsm, (tiles + kCholWarps - 1) / kCholWarps));
    CholArgs ca;
    ca.S = dS; ca.ld = Np; ca.nb = nb; ca.N = N; ca.rhs = drhs; ca.R = dR; ca.Ld = dLd; ca.Dinv = dDinv; ca.x = dx; ca.barrier = dbar;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        const bool prof = rep == 5;
        ca.prof = prof ? dprof : nullptr;
        CK(cudaMemcpy(dS, dS0, sizeof(double) * Np * Np, cudaMemcpyDeviceToDevice));
        CK(cudaMemset(dbar, 0, 4)); CK(cudaMemset(dprof, 0, 16 * sizeof(long long)));
        void* args[] = {&ca};
        cudaEventRecord(e0);
        CK(cudaLaunchCooperativeKernel((void*)k_chol_coop, dim3(G), dim3(kCholWarps * 32), args, 0, 0));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (!prof && rep > 0) best = std::min(best, ms);
    }
    std::vector<double> x(N);
    long long prof[16];
    CK(cudaMemcpy(x.data(), dx, sizeof(double) * N, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(prof, dprof, sizeof(prof), cudaMemcpyDeviceToHost));
    double rmax = 0.0, bmax = 0.0;
    for (int i = 0; i < N; i++) {
        double s = 0.0;
        for (int j = 0; j < N; j++) s += S[i + (size_t)Np * j] * x[j];
        rmax = std::max(rmax, fabs(s - rhs[i])); bmax = std::max(bmax, fabs(rhs[i]));
    }
    const double flops = (double)N * N * N / 3.0;
    printf("N=%d nb=%d grid=%d: %.3f ms (%.2f TFLOP/s), residual %.2e (rel %.2e), x[3..5]=%g %g %g\n", N, nb, G, best,
           flops / (best * 1e-3) / 1e12, rmax, rmax / bmax, x[3], x[4], x[5]);
    const char* names[11] = {"init+barrier", "f1 load X", "f2 diag update", "f3 potrf (warp 0)", "f4 wait other warps", "f5 trsm+store",
                             "f6 barrier", "b1 load Ld", "b2 tri-solve", "b3 y update", "b4 barrier"};
    for (int k = 0; k < 11; k++) printf("    %-24s %9.1f us\n", names[k], prof[k] / (clock_khz * 1e-3));
    cudaFree(dS); cudaFree(dS0); cudaFree(dR); cudaFree(dLd); cudaFree(dDinv); cudaFree(dx); cudaFree(drhs); cudaFree(dbar); cudaFree(dprof);
    return 0;
}

int main(int argc, char** argv)
{
    int nsm = 0, khz = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    std::vector<int> Ns;
    for (int k = 1; k < argc; k++) Ns.push_back(atoi(argv[k]));
    if (Ns.empty()) Ns = {294, 1542, 3072};
    for (int N : Ns)
        if (run(N, nsm, khz)) return 1;
    return 0;
}
