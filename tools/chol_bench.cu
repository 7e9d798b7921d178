// Diagnostic micro-benchmark: cycles per 64x64 chol_solve by one warp (als_common.cuh), alone on an SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --extended-lambda -I../lkpy_b200/csrc chol_bench.cu -o chol_bench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "als_common.cuh"

namespace lk {
void set_error(const char *, ...) {}
int sm_count() { return 148; }
}  // namespace lk

template <int WARPS>
__global__ void bench(const float *A0, const float *y0, float *xout, long long *cycles, int reps)
{
    constexpr int KP = 64, LDA = KP + 4;
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *As = sm + warp * (KP * LDA + 2 * KP);
    float *ys = As + KP * LDA;
    float *dinv = ys + KP;
    long long total = 0;
    bool bad = false;
    for (int r = 0; r < reps; r++) {
        for (int i = lane; i < KP * KP; i += 32) As[(i / KP) * LDA + (i % KP)] = A0[i];
        for (int i = lane; i < KP; i += 32) ys[i] = y0[i];
        __syncwarp();
        const long long t0 = clock64();
        bad |= lk::chol_solve<KP, 1>(As, ys, dinv, lane);
        total += clock64() - t0;
        __syncwarp();
    }
    if (lane == 0) cycles[blockIdx.x * WARPS + warp] = total / reps;
    if (blockIdx.x == 0 && warp == 0)
        for (int i = lane; i < KP; i += 32) xout[i] = bad ? -1.f : ys[i];
}

int main()
{
    const int K = 64;
    std::vector<float> M(200 * K), A(K * K, 0.f), y(K);
    srand(1);
    for (auto &v : M) v = (rand() / (float)RAND_MAX - 0.5f) * 0.2f;
    for (int n = 0; n < 200; n++)
        for (int a = 0; a < K; a++)
            for (int b = 0; b < K; b++) A[a * K + b] += 40.f * M[n * K + a] * M[n * K + b];
    for (int a = 0; a < K; a++) A[a * K + a] += 0.5f, y[a] = rand() / (float)RAND_MAX;
    float *dA, *dy, *dx;
    long long *dc;
    cudaMalloc(&dA, K * K * 4); cudaMalloc(&dy, K * 4); cudaMalloc(&dx, K * 4); cudaMalloc(&dc, 148 * 32 * 8);
    cudaMemcpy(dA, A.data(), K * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dy, y.data(), K * 4, cudaMemcpyHostToDevice);
    auto run = [&](auto kern, int warps, int grid, const char *label) {
        const int smem = warps * (K * 68 + 2 * K) * 4;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        kern<<<grid, warps * 32, smem>>>(dA, dy, dx, dc, 20);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> c(grid * warps);
        cudaMemcpy(c.data(), dc, c.size() * 8, cudaMemcpyDeviceToHost);
        double s = 0;
        for (auto v : c) s += v;
        printf("%-40s %8.0f cycles/solve (%s)\n", label, s / c.size(), cudaGetErrorString(e));
    };
    run(bench<1>, 1, 1, "1 warp alone");
    run(bench<4>, 4, 148, "4 warps/SM (1 per scheduler)");
    run(bench<8>, 8, 148, "8 warps/SM");
    run(bench<12>, 12, 148, "12 warps/SM (3 per scheduler)");
    // residual check (of the last kernel run)
    std::vector<float> x(K);
    cudaMemcpy(x.data(), dx, K * 4, cudaMemcpyDeviceToHost);
    double rn = 0, yn = 0;
    for (int a = 0; a < K; a++) {
        double s = 0;
        for (int b = 0; b < K; b++) s += (double)A[a * K + b] * x[b];
        rn += (s - y[a]) * (s - y[a]);
        yn += (double)y[a] * y[a];
    }
    printf("relative residual %.2e\n", sqrt(rn / yn));
    return 0;
}
