// als_common.cuh — pieces shared by the SIMT and tensor-core ALS kernels:
// the in-shared-memory Cholesky solve (role of LAPACK sposv, solve.rs:65-106)
// and the row write-back with the |x - x_old|^2 reduction (implicit.rs:121-124).
#pragma once

#include "common.cuh"

namespace lk {

template <int NW>
__device__ __forceinline__ void cta_sync()
{
    if constexpr (NW == 1)
        __syncwarp();
    else
        __syncthreads();
}

// Solve A x = y for a KP x KP SPD matrix held row-major in shared memory with
// row stride KP+4 (lower triangle used), rhs in ys, scratch dinv[KP].  Executed
// by NW warps together (NW == 1: one warp, only __syncwarp).  Thread t owns rows
// t + q*32*NW.  On return ys holds x; the result is true when a pivot was not
// positive (LAPACK info != 0).  Right-looking, 4-column panels: every thread
// factors the 4x4 diagonal block redundantly from broadcast reads, solves the
// panel entries of its own rows, then updates the trailing matrix in 4x4 tiles.
template <int KP, int NW>
__device__ __forceinline__ bool chol_solve(float *As, float *ys, float *dinv, const int tid)
{
    constexpr int NT = NW * 32;
    constexpr int LDA = KP + 4;
    constexpr int RPT = KP / NT;
    static_assert(RPT >= 1, "at most KP threads may share one system");
    bool bad = false;
    for (int j0 = 0; j0 < KP; j0 += 4) {
        const float a00 = As[(j0 + 0) * LDA + j0];
        const float a10 = As[(j0 + 1) * LDA + j0], a11 = As[(j0 + 1) * LDA + j0 + 1];
        const float a20 = As[(j0 + 2) * LDA + j0], a21 = As[(j0 + 2) * LDA + j0 + 1],
                    a22 = As[(j0 + 2) * LDA + j0 + 2];
        const float a30 = As[(j0 + 3) * LDA + j0], a31 = As[(j0 + 3) * LDA + j0 + 1],
                    a32 = As[(j0 + 3) * LDA + j0 + 2], a33 = As[(j0 + 3) * LDA + j0 + 3];
        const float l00 = sqrtf(a00), i0 = 1.0f / l00;
        const float l10 = a10 * i0, l20 = a20 * i0, l30 = a30 * i0;
        const float d1 = a11 - l10 * l10;
        const float l11 = sqrtf(d1), i1 = 1.0f / l11;
        const float l21 = (a21 - l20 * l10) * i1, l31 = (a31 - l30 * l10) * i1;
        const float d2 = a22 - l20 * l20 - l21 * l21;
        const float l22 = sqrtf(d2), i2 = 1.0f / l22;
        const float l32 = (a32 - l30 * l20 - l31 * l21) * i2;
        const float d3 = a33 - l30 * l30 - l31 * l31 - l32 * l32;
        const float l33 = sqrtf(d3), i3 = 1.0f / l33;
        bad |= !(a00 > 0.0f && d1 > 0.0f && d2 > 0.0f && d3 > 0.0f);
        cta_sync<NW>();  // all reads of the diagonal block precede its overwrite

        float x[RPT][4];
#pragma unroll
        for (int q = 0; q < RPT; q++) {
            const int i = tid + q * NT;
            x[q][0] = x[q][1] = x[q][2] = x[q][3] = 0.0f;
            if (i >= j0 + 4) {
                float4 ar = *reinterpret_cast<const float4 *>(As + i * LDA + j0);
                const float x0 = ar.x * i0;
                const float x1 = (ar.y - x0 * l10) * i1;
                const float x2 = (ar.z - x0 * l20 - x1 * l21) * i2;
                const float x3 = (ar.w - x0 * l30 - x1 * l31 - x2 * l32) * i3;
                x[q][0] = x0, x[q][1] = x1, x[q][2] = x2, x[q][3] = x3;
                *reinterpret_cast<float4 *>(As + i * LDA + j0) = make_float4(x0, x1, x2, x3);
            } else if (i >= j0) {
                const int r = i - j0;
                float4 lr = r == 0   ? make_float4(l00, 0.f, 0.f, 0.f)
                            : r == 1 ? make_float4(l10, l11, 0.f, 0.f)
                            : r == 2 ? make_float4(l20, l21, l22, 0.f)
                                     : make_float4(l30, l31, l32, l33);
                *reinterpret_cast<float4 *>(As + i * LDA + j0) = lr;
                dinv[i] = r == 0 ? i0 : r == 1 ? i1 : r == 2 ? i2 : i3;
            }
        }
        cta_sync<NW>();
        for (int cc = j0 + 4; cc < KP; cc += 4) {
            const float4 L0 = *reinterpret_cast<const float4 *>(As + (cc + 0) * LDA + j0);
            const float4 L1 = *reinterpret_cast<const float4 *>(As + (cc + 1) * LDA + j0);
            const float4 L2 = *reinterpret_cast<const float4 *>(As + (cc + 2) * LDA + j0);
            const float4 L3 = *reinterpret_cast<const float4 *>(As + (cc + 3) * LDA + j0);
#pragma unroll
            for (int q = 0; q < RPT; q++) {
                const int i = tid + q * NT;
                if (i >= cc) {
                    float4 av = *reinterpret_cast<float4 *>(As + i * LDA + cc);
                    av.x -= x[q][0] * L0.x + x[q][1] * L0.y + x[q][2] * L0.z + x[q][3] * L0.w;
                    av.y -= x[q][0] * L1.x + x[q][1] * L1.y + x[q][2] * L1.z + x[q][3] * L1.w;
                    av.z -= x[q][0] * L2.x + x[q][1] * L2.y + x[q][2] * L2.z + x[q][3] * L2.w;
                    av.w -= x[q][0] * L3.x + x[q][1] * L3.y + x[q][2] * L3.z + x[q][3] * L3.w;
                    *reinterpret_cast<float4 *>(As + i * LDA + cc) = av;
                }
            }
        }
        cta_sync<NW>();
    }
    // forward substitution L z = y (column oriented)
    for (int j = 0; j < KP; j++) {
        const float zj = ys[j] * dinv[j];
        cta_sync<NW>();
#pragma unroll
        for (int q = 0; q < RPT; q++) {
            const int i = tid + q * NT;
            if (i > j)
                ys[i] -= As[i * LDA + j] * zj;
            else if (i == j)
                ys[i] = zj;
        }
        cta_sync<NW>();
    }
    // back substitution L^T x = z
    for (int j = KP - 1; j >= 0; j--) {
        const float xj = ys[j] * dinv[j];
        cta_sync<NW>();
#pragma unroll
        for (int q = 0; q < RPT; q++) {
            const int i = tid + q * NT;
            if (i < j)
                ys[i] -= As[j * LDA + i] * xj;
            else if (i == j)
                ys[i] = xj;
        }
        cta_sync<NW>();
    }
    return bad;
}

// Write x (in ys) to the row of `this` (and its replicas on peer GPUs) and add
// |x - x_old|^2 to the delta accumulator; a failed solve leaves the row
// untouched and records the row in the status word.
template <int KP, int NW>
__device__ __forceinline__ void write_row(const lk_als_args &a, const int row, float *thisrow,
                                          const float *ys, const int tid, bool bad)
{
    constexpr int NT = NW * 32;
    constexpr int RPT = KP / NT;
    const int lane = tid & 31;
    const int k = a.k;
    if constexpr (NW == 1)
        bad = __any_sync(FULL, bad);
    else
        bad = __syncthreads_or(bad);
    if (bad) {
        if (tid == 0) atomicCAS(a.d_status, 0, row + 1);
        return;
    }
    float d2 = 0.0f;
#pragma unroll
    for (int q = 0; q < RPT; q++) {
        const int i = tid + q * NT;
        if (i < k) {
            const float xn = ys[i];
            const float d = xn - thisrow[i];
            d2 = fmaf(d, d, d2);
            thisrow[i] = xn;
            for (int r = 0; r < a.n_replicas; r++)
                a.d_replicas[r][(size_t)(a.replica_row0 + row) * k + i] = xn;
        }
    }
    d2 = warp_sum(d2);
    if (lane == 0 && d2 != 0.0f) atomicAdd(a.d_sqdelta, (double)d2);
}

}  // namespace lk
