// Micro-benchmark / unit test of the tensor-core blocked Cholesky (lkpy_b200/csrc/chol_tc.cuh):
// random SPD 64x64 systems are stored into TMEM in the accumulator layout, solved, and checked
// against a double-precision solve on the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../lkpy_b200/csrc -I../include chol_tc_bench.cu -o chol_tc_bench
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "chol_tc.cuh"

namespace lk {
void set_error(const char *, ...) {}
int sm_count() { return 148; }
}  // namespace lk

using namespace lk;

template <bool SWAPPED>
__global__ void __launch_bounds__(128, 3)
bench(const float *A, const float *y, float *x, int *badout, int n_groups, long long *cycles, long long *phases)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    ctc::Workspace ws = ctc::carve(base);
    uint64_t *bar = reinterpret_cast<uint64_t *>(base + ctc::WS_BYTES + 16);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar + 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(bar, 4);
        mbar_init(bar + 1, 4);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
    const int r = lane & 15, h = lane >> 4;
    uint32_t par = 0;
    long long total = 0;
    long long prof[32] = {0};
    int ngr = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        float yv[2];
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const int s = 2 * p + h;
            const int R = 16 * warp + r;
            const float *row = A + ((size_t)(4 * g + s) * 64 + R) * 64;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = row[16 * q + i];
                ctc::tmem_st16(lane_taddr + 64 * p + 16 * q, v);
            }
            yv[p] = y[(size_t)(4 * g + s) * 64 + R];
        }
        if (tid < 4) ws.bad[tid] = 0;
        tmem_fence_before();
        __syncthreads();
        tmem_fence_after();
        const long long t0 = clock64();
        ctc::solve4<false, SWAPPED>(tmem_base, yv, ws, bar, par, tid, prof);  // template flag: block Gauss-Jordan
        total += clock64() - t0;
        ngr++;
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const int s = 2 * p + h;
            x[(size_t)(4 * g + s) * 64 + 16 * warp + r] = yv[p];
        }
        if (tid < 4) badout[4 * g + tid] = ws.bad[tid];
        __syncthreads();
    }
    if (tid == 0) {
        cycles[blockIdx.x] = ngr ? total / ngr : 0;
        for (int i = 0; i < 32; i++)
            if (i < 4 || i >= 16) phases[blockIdx.x * 32 + i] = ngr ? prof[i] / ngr : 0;
    }
    if (tid == 96)
        for (int i = 4; i < 16; i++) phases[blockIdx.x * 32 + i] = ngr ? prof[i] / ngr : 0;
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
}

static void solve_ref(const float *A, const float *y, double *x)
{
    const int K = 64;
    std::vector<double> L(K * K, 0.0), z(K);
    for (int i = 0; i < K; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[i * K + j];
            for (int k = 0; k < j; k++) s -= L[i * K + k] * L[j * K + k];
            L[i * K + j] = (i == j) ? std::sqrt(s) : s / L[j * K + j];
        }
    for (int i = 0; i < K; i++) {
        double s = y[i];
        for (int k = 0; k < i; k++) s -= L[i * K + k] * z[k];
        z[i] = s / L[i * K + i];
    }
    for (int i = K - 1; i >= 0; i--) {
        double s = z[i];
        for (int k = i + 1; k < K; k++) s -= L[k * K + i] * x[k];
        x[i] = s / L[i * K + i];
    }
}

int main(int argc, char **argv)
{
    const int K = 64;
    const int n_groups = argc > 1 ? atoi(argv[1]) : 148 * 4 * 8;
    const int n_sys = 4 * n_groups;
    const int n_unique = 64;  // distinct systems, tiled over n_sys
    std::vector<float> A((size_t)n_sys * K * K), y((size_t)n_sys * K);
    srand(7);
    for (int u = 0; u < n_unique; u++) {
        const int n = 20 + (u * 37) % 400;  // rows in the Gram: rank-deficient to well-conditioned
        std::vector<float> M((size_t)n * K);
        // every fourth system is badly conditioned (large Gram, small ridge: cond ~1e4-1e5)
        const float mscale = (u % 4 == 3) ? 1.5f : 0.3f;
        const float ridge = (u % 4 == 3) ? 0.1f : 0.6f;
        for (auto &v : M) v = (rand() / (float)RAND_MAX - 0.5f) * mscale;
        float *a = &A[(size_t)u * K * K];
        for (int i = 0; i < K * K; i++) a[i] = 0.f;
        for (int t = 0; t < n; t++)
            for (int i = 0; i < K; i++)
                for (int j = 0; j < K; j++) a[i * K + j] += 40.f * M[t * K + i] * M[t * K + j];
        // OtOr-like dense SPD term + ridge
        for (int i = 0; i < K; i++)
            for (int j = 0; j < K; j++) a[i * K + j] += 0.05f + (i == j ? ridge : 0.f);
        for (int i = 0; i < K; i++) y[(size_t)u * K + i] = rand() / (float)RAND_MAX * 41.f;
    }
    for (int sidx = n_unique; sidx < n_sys; sidx++) {
        memcpy(&A[(size_t)sidx * K * K], &A[(size_t)(sidx % n_unique) * K * K], K * K * 4);
        memcpy(&y[(size_t)sidx * K], &y[(size_t)(sidx % n_unique) * K], K * 4);
    }
    std::vector<double> xref((size_t)n_unique * K);
    for (int u = 0; u < n_unique; u++) solve_ref(&A[(size_t)u * K * K], &y[(size_t)u * K], &xref[(size_t)u * K]);

    float *dA, *dy, *dx;
    int *dbad;
    long long *dc;
    cudaMalloc(&dA, A.size() * 4);
    cudaMalloc(&dy, y.size() * 4);
    cudaMalloc(&dx, y.size() * 4);
    cudaMalloc(&dbad, n_sys * 4);
    cudaMalloc(&dc, 148 * 8 * 8);
    long long *dph;
    cudaMalloc(&dph, 148 * 8 * 8 * 32);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dy, y.data(), y.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 1024 + ctc::WS_BYTES + 64;

    auto run = [&](auto kern, int occ, const char *label) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaMemset(dx, 0, y.size() * 4);
        const int grid = 148 * occ;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        kern<<<grid, 128, smem>>>(dA, dy, dx, dbad, n_groups, dc, dph);  // warm-up
        cudaEventRecord(e0);
        kern<<<grid, 128, smem>>>(dA, dy, dx, dbad, n_groups, dc, dph);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("%s: CUDA error %s\n", label, cudaGetErrorString(e));
            return;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<float> xs(y.size());
        std::vector<int> bad(n_sys);
        std::vector<long long> cyc(grid);
        cudaMemcpy(xs.data(), dx, xs.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(bad.data(), dbad, n_sys * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(cyc.data(), dc, grid * 8, cudaMemcpyDeviceToHost);
        double worst = 0, mean = 0, mean_ill = 0;
        int n_ill = 0;
        int nbad = 0, worst_s = -1;
        for (int sidx = 0; sidx < n_sys; sidx++) {
            const double *xr = &xref[(size_t)(sidx % n_unique) * K];
            double num = 0, den = 0;
            for (int i = 0; i < K; i++) {
                const double d = xs[(size_t)sidx * K + i] - xr[i];
                num += d * d, den += xr[i] * xr[i];
            }
            double rel = std::sqrt(num / den);
            if (!(rel == rel)) rel = 1e30;
            mean += rel;
            if ((sidx % n_unique) % 4 == 3) mean_ill += rel, n_ill++;
            if (rel > worst) worst = rel, worst_s = sidx;
            nbad += bad[sidx];
        }
        long long csum = 0;
        for (auto c : cyc) csum += c;
        std::vector<long long> ph(grid * 32);
        cudaMemcpy(ph.data(), dph, grid * 256, cudaMemcpyDeviceToHost);
        long long phs[32] = {0};
        for (int b = 0; b < grid; b++)
            for (int i = 0; i < 32; i++) phs[i] += ph[b * 32 + i];
        printf("   warp 3 (cycles per group): wait-diag %lld  tmem-ld %lld  trsm %lld  st+tiles %lld  fence+sync %lld  mma-wait %lld | back: products+reduce %lld  sync+wait %lld\n",
               phs[4] / grid, phs[5] / grid, phs[6] / grid, phs[7] / grid, phs[8] / grid, phs[9] / grid, phs[10] / grid, phs[11] / grid);
        printf("   per step (thread 0): diag+wait %lld %lld %lld %lld | trsm %lld %lld %lld %lld | mma issue %lld %lld %lld\n",
               phs[16] / grid, phs[17] / grid, phs[18] / grid, phs[19] / grid, phs[20] / grid, phs[21] / grid,
               phs[22] / grid, phs[23] / grid, phs[24] / grid, phs[25] / grid, phs[26] / grid);
        printf("   phases per group (cycles): diag %lld  trsm+tiles %lld  mma %lld  back-subst %lld\n", phs[0] / grid,
               phs[1] / grid, phs[2] / grid, phs[3] / grid);
        // SM cycles per system: kernel time * clock / (systems per SM)
        const double per_sys_us = ms * 1e3 / n_sys * 148.0;
        printf("%-28s occ %d: %.3f ms for %d systems = %.2f us*SM per system (%.0f cycles @1.9GHz); "
               "group latency %lld cycles; rel err worst %.3e (sys %d) mean %.3e (badly conditioned quarter %.3e); "
               "bad flags %d\n",
               label, occ, ms, n_sys, per_sys_us, per_sys_us * 1900.0, csum / grid, worst, worst_s,
               mean / n_sys, mean_ill / (n_ill ? n_ill : 1), nbad);
        if (worst > 1e-3) {
            const int sidx = worst_s < 0 ? 0 : worst_s;
            printf("   x[%d][0..7] gpu:", sidx);
            for (int i = 0; i < 8; i++) printf(" %.5f", xs[(size_t)sidx * K + i]);
            printf("\n   ref:            ");
            for (int i = 0; i < 8; i++) printf(" %.5f", xref[(size_t)(sidx % n_unique) * K + i]);
            printf("\n   x[0][56..63] gpu:");
            for (int i = 56; i < 64; i++) printf(" %.5f", xs[i]);
            printf("\n   ref:             ");
            for (int i = 56; i < 64; i++) printf(" %.5f", xref[i]);
            printf("\n");
        }
    };
    for (int occ = 1; occ <= 3; occ++) run(bench<false>, occ, "blocked Cholesky");
    for (int occ = 1; occ <= 3; occ++) run(bench<true>, occ, "block Gauss-Jordan");
        return 0;
}
