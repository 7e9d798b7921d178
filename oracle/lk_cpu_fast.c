/*
 * lk_cpu_fast.c — the TIMED CPU baseline of the ALS half-epoch (bench.py's cpu_baseline /
 * --impl reference legs only; test infrastructure like the rest of oracle/).
 *
 * Same algorithm as lk_oracle.c's lk_oracle_als_half_f32, i.e. the reference's
 *   src/accel/als/implicit.rs:87-125 / explicit.rs:80-119 (gather, Gram, right-hand side)
 *   src/accel/als/solve.rs:65-106 (sposv: Cholesky factor + two triangular solves)
 *   src/accel/als/implicit.rs:55-85 (parallel map over rows, sum of squared deltas)
 * restated so that it scales on a many-core host, which the plain oracle does not:
 *   * the reference's Gram is `matrixmultiply`'s sgemm (a register-blocked AVX/FMA micro-kernel
 *     called from every rayon worker without any shared state).  Calling SciPy's OpenBLAS sgemm
 *     from 128 OpenMP threads instead serialises on OpenBLAS's global buffer table (round 1: no
 *     scaling past 16 threads).  Here the Gram is a thread-private register-blocked rank-n update
 *     (4 x 64 accumulator block, the compiler vectorises the 64-wide rows: AVX-512 / AVX2 / SSE
 *     clones picked at load time);
 *   * the k x k Cholesky is the row-oriented (dot-product) form whose inner loops vectorise.
 * Numerics: f32 throughout like the reference; summation order differs from the oracle's (as the
 * reference's own sgemm does) — tests/test_oracle.py checks it against lk_oracle.c to 1e-4.
 *
 * Build: oracle/Makefile (gcc -O3 -fopenmp, no -march: target_clones does the dispatch so the
 * library built in the CPU container also runs on the GPU box's host).
 */

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LK_EXPORT __attribute__((visibility("default")))
#define LK_CLONES __attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))

typedef float v16 __attribute__((vector_size(64), aligned(4)));

/* A[k x k] (row-major, lda = k) += sum_j Ml[j, :]^T M[j, :]  for k % 16 == 0: a 4 x 64 block of A
 * lives in sixteen vector registers across the whole j loop (k = 64: the full row width) */
LK_CLONES static void gram_rank_update(const float *restrict Ml, const float *restrict M, int64_t n,
                                       int k, float *restrict A)
{
    for (int a0 = 0; a0 < k; a0 += 4) {
        int b0 = 0;
        for (; b0 + 64 <= k; b0 += 64) {
            v16 c00 = {0}, c01 = {0}, c02 = {0}, c03 = {0}, c10 = {0}, c11 = {0}, c12 = {0}, c13 = {0};
            v16 c20 = {0}, c21 = {0}, c22 = {0}, c23 = {0}, c30 = {0}, c31 = {0}, c32 = {0}, c33 = {0};
            for (int64_t j = 0; j < n; j++) {
                const v16 *m = (const v16 *)(M + j * k + b0);
                const v16 m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3];
                const float *l = Ml + j * k + a0;
                const float l0 = l[0], l1 = l[1], l2 = l[2], l3 = l[3];
                c00 += l0 * m0, c01 += l0 * m1, c02 += l0 * m2, c03 += l0 * m3;
                c10 += l1 * m0, c11 += l1 * m1, c12 += l1 * m2, c13 += l1 * m3;
                c20 += l2 * m0, c21 += l2 * m1, c22 += l2 * m2, c23 += l2 * m3;
                c30 += l3 * m0, c31 += l3 * m1, c32 += l3 * m2, c33 += l3 * m3;
            }
            v16 *r0 = (v16 *)(A + (a0 + 0) * k + b0), *r1 = (v16 *)(A + (a0 + 1) * k + b0);
            v16 *r2 = (v16 *)(A + (a0 + 2) * k + b0), *r3 = (v16 *)(A + (a0 + 3) * k + b0);
            r0[0] += c00, r0[1] += c01, r0[2] += c02, r0[3] += c03;
            r1[0] += c10, r1[1] += c11, r1[2] += c12, r1[3] += c13;
            r2[0] += c20, r2[1] += c21, r2[2] += c22, r2[3] += c23;
            r3[0] += c30, r3[1] += c31, r3[2] += c32, r3[3] += c33;
        }
        for (; b0 < k; b0 += 16) { /* k % 64 != 0: 16-wide remainder blocks */
            v16 c0 = {0}, c1 = {0}, c2 = {0}, c3 = {0};
            for (int64_t j = 0; j < n; j++) {
                const v16 m0 = *(const v16 *)(M + j * k + b0);
                const float *l = Ml + j * k + a0;
                c0 += l[0] * m0, c1 += l[1] * m0, c2 += l[2] * m0, c3 += l[3] * m0;
            }
            *(v16 *)(A + (a0 + 0) * k + b0) += c0, *(v16 *)(A + (a0 + 1) * k + b0) += c1;
            *(v16 *)(A + (a0 + 2) * k + b0) += c2, *(v16 *)(A + (a0 + 3) * k + b0) += c3;
        }
    }
}

/* sposv (solve.rs:65-106): A = U^T U on the upper triangle of the (symmetric, fully stored) row-major
 * matrix, right-looking so that every inner loop is an axpy over a contiguous row segment; then
 * U^T z = b and U x = z. */
LK_CLONES static int chol_solve(float *restrict A, float *restrict b, int k)
{
    for (int j = 0; j < k; j++) {
        const float d = A[j * k + j];
        if (!(d > 0.0f)) return j + 1;
        const float inv = 1.0f / sqrtf(d);
        float *restrict uj = A + j * k;
#pragma omp simd
        for (int c = j; c < k; c++) uj[c] *= inv;
        for (int i = j + 1; i < k; i++) {
            const float f = uj[i];
            float *restrict ai = A + i * k;
#pragma omp simd
            for (int c = i; c < k; c++) ai[c] -= f * uj[c];
        }
    }
    for (int i = 0; i < k; i++) { /* U^T z = b: column-oriented, axpy over row i of U */
        const float z = b[i] / A[i * k + i];
        b[i] = z;
        const float *restrict ui = A + i * k;
#pragma omp simd
        for (int c = i + 1; c < k; c++) b[c] -= ui[c] * z;
    }
    for (int i = k - 1; i >= 0; i--) { /* U x = z: dot product over row i of U */
        const float *restrict ui = A + i * k;
        float s = 0.0f;
#pragma omp simd reduction(+ : s)
        for (int c = i + 1; c < k; c++) s += ui[c] * b[c];
        b[i] = (b[i] - s) / ui[i];
    }
    return 0;
}

/* Same contract as lk_oracle_als_half_f32 (lk_oracle.c): mode 0 implicit / 1 explicit, `this_`
 * updated in place, returns 0 or (row + 1) of the first failed solve. */
LK_EXPORT int64_t lk_cpu_als_half_f32(int mode, const int64_t *indptr, const int32_t *cols,
                                      const float *vals, int64_t n_rows, int k, float *this_,
                                      const float *other, const float *otor, float reg, int nthreads,
                                      double *out_sqdelta)
{
    int64_t max_n = 1;
    for (int64_t r = 0; r < n_rows; r++) {
        int64_t n = indptr[r + 1] - indptr[r];
        if (n > max_n) max_n = n;
    }
    const int kp = (k + 15) & ~15; /* padded row length of the gathered copies */
    double total = 0.0;
    int64_t fail = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads) reduction(+ : total)
    {
        float *A = (float *)aligned_alloc(64, sizeof(float) * (size_t)kp * kp);
        float *y = (float *)aligned_alloc(64, sizeof(float) * (size_t)kp);
        /* long rows are gathered and accumulated in blocks so that the copies stay cache-resident */
        const int64_t blk = 1024;
        float *M = (float *)aligned_alloc(64, sizeof(float) * (size_t)blk * kp);
        float *Ml = (float *)aligned_alloc(64, sizeof(float) * (size_t)blk * kp);
#pragma omp for schedule(dynamic, 8)
        for (int64_t r = 0; r < n_rows; r++) {
            float *row = this_ + r * k;
            const int64_t s = indptr[r], n = indptr[r + 1] - s;
            if (n == 0) { /* implicit.rs:98-101 */
                for (int a = 0; a < k; a++) row[a] = 0.0f;
                continue;
            }
            memset(A, 0, sizeof(float) * (size_t)kp * kp);
            memset(y, 0, sizeof(float) * (size_t)kp);
            for (int64_t j0 = 0; j0 < n; j0 += blk) {
                const int64_t nb = n - j0 < blk ? n - j0 : blk;
                for (int64_t j = 0; j < nb; j++) { /* other.select + mt * vals (implicit.rs:108-111) */
                    const float *o = other + (int64_t)cols[s + j0 + j] * k;
                    const float v = vals[s + j0 + j];
                    const float w = mode == 0 ? v + 1.0f : v; /* vals += 1 (:116) */
                    float *mj = M + j * kp, *lj = Ml + j * kp;
                    for (int a = 0; a < k; a++) {
                        mj[a] = o[a];
                        lj[a] = mode == 0 ? o[a] * v : o[a];
                        y[a] += o[a] * w; /* y = mt.dot(vals) (:117) */
                    }
                    for (int a = k; a < kp; a++) mj[a] = lj[a] = 0.0f;
                }
                gram_rank_update(Ml, M, nb, kp, A); /* mtm = mtl.dot(o_picked) (:112) */
            }
            if (kp != k) { /* compact the padded system */
                for (int a = 0; a < k; a++) memmove(A + a * k, A + a * kp, sizeof(float) * k);
            }
            if (mode == 0) {
                for (int i = 0; i < k * k; i++) A[i] += otor[i]; /* a = otor + mtm (:115) */
            } else {
                const float rn = reg * (float)n; /* explicit.rs:106-108 */
                for (int a = 0; a < k; a++) A[a * k + a] += rn;
            }
            if (chol_solve(A, y, k) != 0) {
#pragma omp critical
                {
                    if (fail == 0 || r + 1 < fail) fail = r + 1;
                }
                continue;
            }
            double d2 = 0.0;
            for (int a = 0; a < k; a++) {
                const float d = y[a] - row[a];
                d2 += (double)d * (double)d;
                row[a] = y[a];
            }
            total += d2;
        }
        free(A);
        free(y);
        free(M);
        free(Ml);
    }
    if (out_sqdelta) *out_sqdelta = total;
    return fail;
}
