// chol_tc.cuh — blocked Cholesky solve of four 64x64 SPD systems that live in TMEM
// (role of LAPACK sposv, reference src/accel/als/solve.rs:65-106), with the rank-16
// trailing updates on the tensor cores.
//
// A CTA of 4 warps owns 4 systems as two "pairs": the M=64 accumulator layout puts
// rows 16w..16w+15 of a system into TMEM lanes 32w..32w+15 (system 2p) or
// 32w+16..32w+31 (system 2p+1) at columns 64p..64p+63, so warp w can touch block-row
// w of every system and lane l = 16h + r owns row 16w + r of system 2p + h.
//
// Right-looking, block size 16.  Step j (0..3):
//   1. warp j factors the two 16x16 diagonal blocks per pair in registers (row r in
//      lane r; pivots and multipliers travel by 16-lane shuffles), carries the
//      right-hand side along (fused forward substitution), and publishes L_jj^T,
//      the inverse pivots and z_j in shared memory;
//   2. warps w > j solve X L_jj^T = A_wj for their rows (one row per lane, L_jj^T
//      broadcast from shared memory), update their right-hand sides, store X back
//      into the dead panel columns of TMEM (the back substitution reads it there)
//      and write X as two tf32 K-major operand tiles: hi = X with the low 13
//      mantissa bits cleared, lo = X - hi (exact in f32);
//   3. one elected lane per warp issues, for "its" system, D[:, 16(j+1):] -= X X^T as
//      hi.hi + hi.lo + lo.hi — six tcgen05.mma.kind::tf32 (K = 8 each, A negated
//      through the instruction descriptor), fp32 accumulation in place in TMEM;
//      only the warp that factors the next diagonal block polls their mbarrier.
// Back substitution L^T x = z runs block-wise from the bottom: warp j takes the
// contributions of the finished blocks off its right-hand side, solves its
// transposed 16x16 block (every lane redundantly, from broadcast reads of L_jj),
// multiplies its rows of block column j-1 by x_j and sums them over the 16 lanes
// (shuffle reduce-scatter) before the barrier — the one contribution warp j-1 is
// still missing — and the older block columns after it.
// GJ = true (solve4's template flag) is the block Gauss-Jordan variant described there.
#pragma once

#include "tc_common.cuh"

namespace lk {
namespace ctc {

constexpr int LDT = 20;                       // row stride (floats) of a transposed diagonal block
constexpr int TILE_BYTES = 64 * 16 * 4;       // one operand tile: 64 rows x 16 tf32, K-major, no swizzle
constexpr int TILES_BYTES = 4 * 2 * TILE_BYTES;          // [system][hi, lo]
constexpr int LT_FLOATS = 4 * 16 * LDT;                  // [system][16][LDT]: L_jj^T of the current step
constexpr int LD_FLOATS = 4 * 4 * 16 * LDT;              // [system][step][16][LDT]: L_jj, kept for the back substitution
constexpr int INVD_FLOATS = 4 * 64;                      // [system][64]
constexpr int ZB_FLOATS = 4 * 16;                        // [system][16]
constexpr int TSUM_FLOATS = 4 * 4 * 4 * 16;              // [block row][block column][system][16]
constexpr int WS_BYTES =
    TILES_BYTES + (LT_FLOATS + LD_FLOATS + INVD_FLOATS + ZB_FLOATS + TSUM_FLOATS) * 4 + 16;

// instruction descriptor (mma_sm100_desc.hpp): D f32, A = B = tf32, both K-major, A negated, M = 64; N added at run time
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | ((64u >> 4) << 24);
// operand tile address(row, k) = (row / 8) * 512 + (k / 4) * 128 + (row % 8) * 16 + (k % 4) * 4:
// core matrix = 8 rows x 16 B; LBO (next core matrix along K) = 128 B, SBO (next 8-row group) = 512 B
constexpr uint64_t DESC_KMAJOR = (uint64_t(128 >> 4) << 16) | (uint64_t(512 >> 4) << 32) | (1ull << 46);
constexpr uint64_t DESC_KMAJOR_SWAPPED = (uint64_t(512 >> 4) << 16) | (uint64_t(128 >> 4) << 32) | (1ull << 46);

struct Workspace {
    unsigned char *tiles;  // 128-byte aligned
    float *lt, *ld, *invd, *zb, *tsum;
    int *bad;              // [4] per-system "pivot not positive" flags
};

__device__ __forceinline__ Workspace carve(unsigned char *p)
{
    Workspace w;
    w.tiles = p;
    w.lt = reinterpret_cast<float *>(p + TILES_BYTES);
    w.ld = w.lt + LT_FLOATS;
    w.invd = w.ld + LD_FLOATS;
    w.zb = w.invd + INVD_FLOATS;
    w.tsum = w.zb + ZB_FLOATS;
    w.bad = reinterpret_cast<int *>(w.tsum + TSUM_FLOATS);
    return w;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16])
{
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = __uint_as_float(u[i]);
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])),
          "r"(__float_as_uint(r[3])), "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])),
          "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7])), "r"(__float_as_uint(r[8])),
          "r"(__float_as_uint(r[9])), "r"(__float_as_uint(r[10])), "r"(__float_as_uint(r[11])),
          "r"(__float_as_uint(r[12])), "r"(__float_as_uint(r[13])), "r"(__float_as_uint(r[14])),
          "r"(__float_as_uint(r[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 32 lanes x 64 consecutive 32-bit columns, registers -> TMEM
__device__ __forceinline__ void tmem_st64(uint32_t taddr, const uint32_t (&r)[64])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, "
        "%33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, "
        "%49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
          "r"(r[32]), "r"(r[33]), "r"(r[34]), "r"(r[35]), "r"(r[36]), "r"(r[37]), "r"(r[38]), "r"(r[39]),
          "r"(r[40]), "r"(r[41]), "r"(r[42]), "r"(r[43]), "r"(r[44]), "r"(r[45]), "r"(r[46]), "r"(r[47]),
          "r"(r[48]), "r"(r[49]), "r"(r[50]), "r"(r[51]), "r"(r[52]), "r"(r[53]), "r"(r[54]), "r"(r[55]),
          "r"(r[56]), "r"(r[57]), "r"(r[58]), "r"(r[59]), "r"(r[60]), "r"(r[61]), "r"(r[62]), "r"(r[63])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
}

// MUFU.RCP alone (max relative error 2^-23, the size of an f32 rounding): it sits on the pivot
// chain of the diagonal blocks, where a Newton step would cost two more dependent operations
__device__ __forceinline__ float rcp_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// one lane of a converged warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// non-blocking probe of an mbarrier phase (the waiting warp polls: try_wait parks the thread and
// was measured to wake it ~1 k cycles late)
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// Solve the four systems.  On entry the lower triangles of A are in TMEM (tmem_base = column 0 of
// pair 0, lane field 0) and yv[p] holds the right-hand-side entry of this lane's row of system
// 2p + (lane >> 4); on exit yv[p] holds the solution entry.  `bar` points at two mbarriers (count 4 each: one commit per warp
// and update pass) used only here, `par` is their running phase parity.  ws.bad[s] is set when a pivot of system s was not
// positive (caller zeroes it).  All 128 threads must call.
//
// GJ = true turns the block elimination into a block Gauss-Jordan: at step j *every* other block
// row w (above the pivot block as well as below) gets X_w = A_wj L_jj^-T, the tensor cores update
// all 64 rows (they do anyway: M = 64), the right-hand sides of all rows lose X_w z_j, and the pivot
// block row itself is left alone (its operand-tile rows are zeroed).  After four steps the system
// is block diagonal with the pivot blocks L_jj L_jj^T, so every warp finishes its own 16 unknowns
// at the same time — no serial block back substitution.  Measured error vs Cholesky: see DESIGN.md.
template <bool SWAPPED_DESC = false, bool GJ = false>
__device__ __forceinline__ void solve4(const uint32_t tmem_base, float (&yv)[2], const Workspace &ws, uint64_t *bar,
                                       uint32_t &par, const int tid, long long *prof = nullptr)
{
    long long t_prev = prof ? clock64() : 0;
    auto mark = [&](int i, int step = -1) {  // thread 0's view; per-step copies at prof[16 + 4 * i + step]
        if (prof != nullptr && tid == 0) {
            const long long t = clock64();
            prof[i] += t - t_prev;
            if (step >= 0) prof[16 + 4 * i + step] += t - t_prev;
            t_prev = t;
        }
    };
    long long t_prev3 = t_prev;
    auto mark3 = [&](int i) {  // warp 3's view (it runs the triangular solve of every step)
        if (prof != nullptr && tid == 96) {
            const long long t = clock64();
            prof[i] += t - t_prev3;
            t_prev3 = t;
        }
    };
    const int lane = tid & 31, warp = tid >> 5;
    const int r = lane & 15, h = lane >> 4;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
    constexpr uint64_t DESC = SWAPPED_DESC ? DESC_KMAJOR_SWAPPED : DESC_KMAJOR;
    // before step j rewrites the operand tiles: every trailing-update instruction of step j - 1 has completed
    // (Cholesky variant: bar[1] collects the commits issued after the second pass, its phase parity follows
    // bar[0]'s; the Gauss-Jordan variant commits everything to bar[0], which the next step has waited for)
    auto wait_tiles_free = [&](const int j) {
        if (!GJ && j > 0) {
            while (!mbar_test(bar + 1, par ^ 1u)) {
            }
        }
    };

#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        if (warp == j) {
            // ---- diagonal blocks of both pairs, interleaved for ILP ----
            float a[2][16];
            tmem_ld16(lane_taddr + 16 * j, a[0]);
            tmem_ld16(lane_taddr + 64 + 16 * j, a[1]);
            float inv_mine[2] = {0.0f, 0.0f};
            bool bad[2] = {false, false};
            float zt[2] = {yv[0], yv[1]};  // becomes z_j = L_jj^-1 y_j
#pragma unroll
            for (int k = 0; k < 16; k++) {
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    // critical path: pivot broadcast -> reciprocal -> scaled column -> rank-1 update;
                    // the square-root scaling of L and the right-hand side ride beside it
                    const float akk = __shfl_sync(FULL, a[p][k], k, 16);
                    bad[p] |= !(akk > 0.0f);
                    const float ak = (r >= k) ? a[p][k] : 0.0f;
                    const float sk = ak * rcp_fast(akk);
#pragma unroll
                    for (int c = k + 1; c < 16; c++) {
                        const float ac = __shfl_sync(FULL, ak, c, 16);
                        a[p][c] = fmaf(-sk, ac, a[p][c]);
                    }
                    const float inv = rsqrt_nr(akk);
                    const float lk = ak * inv;
                    a[p][k] = lk;
                    if (r == k) inv_mine[p] = inv;
                    const float zk = __shfl_sync(FULL, zt[p], k, 16) * inv;
                    zt[p] = (r == k) ? zk : fmaf(-lk, zk, zt[p]);
                }
            }
            if constexpr (!GJ) {  // Cholesky: the right-hand side continues as z; Gauss-Jordan keeps y
                yv[0] = zt[0];
                yv[1] = zt[1];
            }
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const int s = 2 * p + h;
                float *lt = ws.lt + s * 16 * LDT;
#pragma unroll
                for (int c = 0; c < 16; c++) lt[c * LDT + r] = a[p][c];  // Lt[c][r] = L[r][c]
                float *ldr = ws.ld + ((s * 4 + j) * 16 + r) * LDT;       // row r of L_jj
#pragma unroll
                for (int q = 0; q < 4; q++)
                    *reinterpret_cast<float4 *>(ldr + 4 * q) =
                        make_float4(a[p][4 * q], a[p][4 * q + 1], a[p][4 * q + 2], a[p][4 * q + 3]);
                ws.invd[s * 64 + 16 * j + r] = inv_mine[p];
                ws.zb[s * 16 + r] = zt[p];
                if (bad[p] && r == 0) ws.bad[s] = 1;
                if constexpr (GJ) {
                    if (j < 3) {  // the pivot block row takes no part in this step's update
                        if (p == 0) wait_tiles_free(j);
                        const int R = 16 * warp + r;
                        unsigned char *thi = ws.tiles + (s * 2) * TILE_BYTES + (R >> 3) * 512 + (R & 7) * 16;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            *reinterpret_cast<float4 *>(thi + q * 128) = make_float4(0.f, 0.f, 0.f, 0.f);
                            *reinterpret_cast<float4 *>(thi + TILE_BYTES + q * 128) = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
            }
        }
        if (!GJ && j == 3) break;
        tmem_fence_before();
        __syncthreads();
        tmem_fence_after();
        mark(0, j);
        mark3(4);
        if (GJ ? (warp != j) : (warp > j)) {
            // both pairs in one pass (two independent dependency chains per lane)
            float a[2][16];
            tmem_ld16(lane_taddr + 16 * j, a[0]);
            tmem_ld16(lane_taddr + 64 + 16 * j, a[1]);
            mark3(5);
            float iv[2][16];
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const int s = 2 * p + h;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 t = *reinterpret_cast<const float4 *>(ws.invd + s * 64 + 16 * j + 4 * q);
                    iv[p][4 * q] = t.x, iv[p][4 * q + 1] = t.y, iv[p][4 * q + 2] = t.z, iv[p][4 * q + 3] = t.w;
                }
            }
            // row k of Lt holds L[m][k] for m > k in 16-byte groups m = 4q .. 4q+3; the rows are
            // fetched one step ahead of their use so that no shared-memory latency sits on the chain
            float4 row[2][2][4];
#pragma unroll
            for (int p = 0; p < 2; p++)
#pragma unroll
                for (int q = 0; q < 4; q++)
                    row[0][p][q] = *reinterpret_cast<const float4 *>(ws.lt + (2 * p + h) * 16 * LDT + 4 * q);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k + 1 < 16) {
#pragma unroll
                    for (int p = 0; p < 2; p++)
#pragma unroll
                        for (int q = (k + 2) / 4; q < 4; q++)
                            row[(k + 1) & 1][p][q] = *reinterpret_cast<const float4 *>(
                                ws.lt + (2 * p + h) * 16 * LDT + (k + 1) * LDT + 4 * q);
                }
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const float xk = a[p][k] * iv[p][k];
                    a[p][k] = xk;
#pragma unroll
                    for (int q = (k + 1) / 4; q < 4; q++) {
                        const float4 t = row[k & 1][p][q];
                        if (4 * q + 0 > k) a[p][4 * q + 0] = fmaf(-xk, t.x, a[p][4 * q + 0]);
                        if (4 * q + 1 > k) a[p][4 * q + 1] = fmaf(-xk, t.y, a[p][4 * q + 1]);
                        if (4 * q + 2 > k) a[p][4 * q + 2] = fmaf(-xk, t.z, a[p][4 * q + 2]);
                        if (4 * q + 3 > k) a[p][4 * q + 3] = fmaf(-xk, t.w, a[p][4 * q + 3]);
                    }
                }
            }
            mark3(6);
            wait_tiles_free(j);
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const int s = 2 * p + h;
                // right-hand side: y_R -= X[R][:] . z_j (two partial sums)
                float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 u = *reinterpret_cast<const float4 *>(ws.zb + s * 16 + 4 * q);
                    acc0 = fmaf(a[p][4 * q + 0], u.x, acc0);
                    acc1 = fmaf(a[p][4 * q + 1], u.y, acc1);
                    acc0 = fmaf(a[p][4 * q + 2], u.z, acc0);
                    acc1 = fmaf(a[p][4 * q + 3], u.w, acc1);
                }
                yv[p] -= acc0 + acc1;
                if constexpr (GJ) {
                    if (j == 3) continue;  // last step: no trailing columns left, only the right-hand sides
                } else {
                    tmem_st16(lane_taddr + 64 * p + 16 * j, a[p]);  // the back substitution reads it there
                }
                // tf32 operand tiles: hi keeps the top 19 bits, lo = x - hi exactly
                const int R = 16 * warp + r;
                unsigned char *thi = ws.tiles + (s * 2) * TILE_BYTES + (R >> 3) * 512 + (R & 7) * 16;
                unsigned char *tlo = thi + TILE_BYTES;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(a[p][4 * q + 0]) & 0xffffe000u);
                    hi.y = __uint_as_float(__float_as_uint(a[p][4 * q + 1]) & 0xffffe000u);
                    hi.z = __uint_as_float(__float_as_uint(a[p][4 * q + 2]) & 0xffffe000u);
                    hi.w = __uint_as_float(__float_as_uint(a[p][4 * q + 3]) & 0xffffe000u);
                    lo.x = a[p][4 * q + 0] - hi.x, lo.y = a[p][4 * q + 1] - hi.y;
                    lo.z = a[p][4 * q + 2] - hi.z, lo.w = a[p][4 * q + 3] - hi.w;
                    *reinterpret_cast<float4 *>(thi + q * 128) = hi;
                    *reinterpret_cast<float4 *>(tlo + q * 128) = lo;
                }
            }
        }
        mark3(7);
        if (GJ && j == 3) break;
        fence_proxy_async();
        tmem_fence_before();
        __syncthreads();
        mark(1, j);
        mark3(8);
        // lane 0 of warp s issues the six MMAs of system s and commits them (the barrier counts 4
        // arrivals); its sibling lanes wait at the __syncwarp so their spinning cannot starve it
        // (the warp index is re-derived through a shuffle so that the compiler knows it is uniform
        // and keeps the descriptors in uniform registers instead of broadcasting them per MMA)
        const int warp_u = __shfl_sync(FULL, warp, 0);
        if (elect_one()) {
            tmem_fence_after();
            const int c0 = 16 * (j + 1);
            const int s = warp_u;
            const uint32_t d = tmem_base + ((uint32_t)((s & 1) * 16) << 16) + (uint32_t)((s >> 1) * 64 + c0);
            const uint32_t hi = smem_u32(ws.tiles + (s * 2) * TILE_BYTES);
            const uint32_t lo = hi + TILE_BYTES;
            // D[:, c .. c + n) -= X X[c .. c + n)^T as hi.hi + hi.lo + lo.hi over the two K = 8 slices
            auto update = [&](const int c, const int n) {
                const uint32_t idesc = IDESC_TF32 | ((uint32_t)(n >> 3) << 17);
                const uint32_t boff = (uint32_t)(c >> 3) * 512;  // B starts at row c
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    const uint64_t a_hi = DESC | (uint64_t)(((hi + kk * 256) >> 4) & 0x3fffu);
                    const uint64_t a_lo = DESC | (uint64_t)(((lo + kk * 256) >> 4) & 0x3fffu);
                    const uint64_t b_hi = DESC | (uint64_t)(((hi + boff + kk * 256) >> 4) & 0x3fffu);
                    const uint64_t b_lo = DESC | (uint64_t)(((lo + boff + kk * 256) >> 4) & 0x3fffu);
                    umma_tf32(d + (uint32_t)(c - c0), a_hi, b_hi, idesc);
                    umma_tf32(d + (uint32_t)(c - c0), a_hi, b_lo, idesc);
                    umma_tf32(d + (uint32_t)(c - c0), a_lo, b_hi, idesc);
                }
            };
            if constexpr (GJ) {
                // (look-ahead as below measured slower here: 3.64 vs 3.54 us*SM per system, tools/chol_tc_bench.cu)
                update(c0, 64 - c0);
                umma_commit(bar);
            } else {
                // look-ahead: the 16 columns the next step works on (its diagonal block and triangular solves)
                // first, and the barrier after them; the rest of the trailing update runs in their shadow and is
                // committed to the second barrier, which the next step's writers of the operand tiles wait for
                // (wait_tiles_free) — the tiles are read by these instructions until they complete
                update(c0, 16);
                umma_commit(bar);
                if (c0 + 16 < 64) update(c0 + 16, 64 - c0 - 16);
                umma_commit(bar + 1);
            }
        }
        __syncwarp();
        // only the warp that factors the next diagonal block needs the updated accumulators now;
        // the others meet it at the barrier after that factorisation
        if (warp == j + 1) {
            while (!mbar_test(bar, par)) {
            }
            tmem_fence_after();
        }
        par ^= 1u;
        mark(2, j);
        mark3(9);
    }

    if constexpr (GJ) {
        // block-diagonal system left: L_ww L_ww^T x_w = y_w — every warp solves its own two pairs
        float rowl[2][16], coll[2][16], inv[2];
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const int s = 2 * p + h;
            const float *ldw = ws.ld + (s * 4 + warp) * 16 * LDT;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 t = *reinterpret_cast<const float4 *>(ldw + r * LDT + 4 * q);  // L[r][k], k < r
                rowl[p][4 * q] = t.x, rowl[p][4 * q + 1] = t.y, rowl[p][4 * q + 2] = t.z, rowl[p][4 * q + 3] = t.w;
            }
#pragma unroll
            for (int k = 0; k < 16; k++) coll[p][k] = ldw[k * LDT + r];  // L[k][r], k > r
            inv[p] = ws.invd[s * 64 + 16 * warp + r];
        }
#pragma unroll
        for (int k = 0; k < 16; k++) {
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const float zk = __shfl_sync(FULL, yv[p] * inv[p], k, 16);
                yv[p] = (r == k) ? zk : ((r > k) ? fmaf(-rowl[p][k], zk, yv[p]) : yv[p]);
            }
        }
#pragma unroll
        for (int k = 15; k >= 0; k--) {
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const float xk = __shfl_sync(FULL, yv[p] * inv[p], k, 16);
                yv[p] = (r == k) ? xk : ((r < k) ? fmaf(-coll[p][k], xk, yv[p]) : yv[p]);
            }
        }
        if (prof != nullptr) {
            __syncthreads();
            mark(3);
        }
        return;
    }

    // ---- back substitution L^T x = z; yv holds z ----
    // ts[(w*4 + jj)*4 + s][c]: what block row w (its x final) takes off the right-hand side of block jj.
    // Step j: warp j finishes x_j, multiplies its rows of block column j-1 (the only contribution the
    // next step is still missing) before the barrier and the older block columns after it, in the
    // shadow of warp j-1's solve.
    auto contribute = [&](const int jj) {
        float v[2][16];
        tmem_ld16(lane_taddr + 16 * jj, v[0]);
        tmem_ld16(lane_taddr + 64 + 16 * jj, v[1]);
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int c = 0; c < 16; c++) v[p][c] *= yv[p];
        // reduce-scatter over the 16 lanes of each half: lane r ends with column r
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1) {
            const bool up = (r & w) != 0;
#pragma unroll
            for (int i = 0; i < w; i++) {
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const float send = up ? v[p][i] : v[p][i + w];
                    const float keep = up ? v[p][i + w] : v[p][i];
                    v[p][i] = keep + __shfl_xor_sync(FULL, send, w, 16);
                }
            }
        }
        ws.tsum[((warp * 4 + jj) * 4 + h) * 16 + r] = v[0][0];
        ws.tsum[((warp * 4 + jj) * 4 + 2 + h) * 16 + r] = v[1][0];
    };
#pragma unroll 1
    for (int j = 3; j >= 0; j--) {
        if (warp == j) {
            // every lane solves the whole 16x16 transposed system of its half redundantly: the
            // right-hand side is broadcast once, then no shuffle sits on the dependency chain
            float x[2][16];
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const int s = 2 * p + h;
                float rhs = yv[p];
                for (int w = j + 1; w < 4; w++) rhs -= ws.tsum[((w * 4 + j) * 4 + s) * 16 + r];
#pragma unroll
                for (int k = 0; k < 16; k++) x[p][k] = __shfl_sync(FULL, rhs, k, 16);
            }
#pragma unroll
            for (int k = 15; k >= 0; k--) {
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const int s = 2 * p + h;
                    const float *ldk = ws.ld + ((s * 4 + j) * 16 + k) * LDT;  // row k of L_jj: L[k][m], m < k
                    const float xk = x[p][k] * ws.invd[s * 64 + 16 * j + k];
                    if (r == k) yv[p] = xk;
#pragma unroll
                    for (int q = 0; 4 * q < k; q++) {
                        const float4 t = *reinterpret_cast<const float4 *>(ldk + 4 * q);
                        if (4 * q + 0 < k) x[p][4 * q + 0] = fmaf(-t.x, xk, x[p][4 * q + 0]);
                        if (4 * q + 1 < k) x[p][4 * q + 1] = fmaf(-t.y, xk, x[p][4 * q + 1]);
                        if (4 * q + 2 < k) x[p][4 * q + 2] = fmaf(-t.z, xk, x[p][4 * q + 2]);
                        if (4 * q + 3 < k) x[p][4 * q + 3] = fmaf(-t.w, xk, x[p][4 * q + 3]);
                    }
                }
            }
            if (j > 0) contribute(j - 1);
        }
        if (j == 0) break;
        __syncthreads();
        if (warp == j)
            for (int jj = j - 2; jj >= 0; jj--) contribute(jj);
    }
    if (prof != nullptr) {
        __syncthreads();
        mark(3);
    }
}

}  // namespace ctc
}  // namespace lk
