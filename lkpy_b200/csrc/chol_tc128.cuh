// chol_tc128.cuh — blocked Cholesky solve of two 128x128 SPD systems that live in TMEM
// (role of LAPACK sposv at k = 128, reference src/accel/als/solve.rs:65-106), with the rank-16
// trailing updates on the tensor cores.  The 64x64 solver of chol_tc.cuh, one level up.
//
// Layout.  A 128x128 lower triangle is three 64x64 blocks, each an M=64 tcgen05 accumulator:
//      G11 (rows 0..63,   cols 0..63)      slot P0, the lanes 0..15  of every 32-lane quarter
//      G21 (rows 64..127, cols 0..63)      slot P0, the lanes 16..31 of every quarter
//      G22 (rows 64..127, cols 64..127)    slot P1, the lanes 16..31 (the lanes 0..15 of P1 are unused)
// with slot P0 = TMEM columns 128 p .. 128 p + 63 and P1 = 128 p + 64 .. 128 p + 127 of system p = 0, 1.
// Warp w can touch TMEM lanes 32 w .. 32 w + 31 only, so lane l = 16 h + r of warp w owns row
//      R = 64 h + 16 w + r
// of BOTH systems: its half h = 0 is a row of the top block row [G11], its half h = 1 a row of the
// bottom block row [G21 G22] — and one 32x32b load of 16 columns of P0 returns, in the two halves
// of the warp, the top and the bottom rows of the same block column of the left panel.
//
// Right-looking, block size 16, eight steps j = 4 half + jb:
//   1. the 16 lanes (half) of warp jb that hold the diagonal block factor it for both systems (row per
//      lane, pivots and multipliers by 16-lane shuffles, forward substitution fused) and publish
//      L_jj^T, L_jj, the inverse pivots and z_j in shared memory;
//   2. every lane whose row lies below the block solves X L_jj^T = A_Rj for its row, updates its
//      right-hand side, stores X back into the dead panel columns of TMEM and writes X as tf32
//      hi / lo K-major operand tiles (128 rows x 16);
//   3. the elected lane of warp p issues, for system p, the trailing updates as hi.hi + hi.lo + lo.hi
//      tcgen05.mma.kind::tf32 (M = 64, A negated):
//          half 0:  G11[:, c0:] -= Xt Xt[c0:]^T    G21[:, c0:] -= Xb Xt[c0:]^T    G22 -= Xb Xb^T
//          half 1:  G22[:, c0:] -= Xb Xb[c0:]^T                    (c0 = 16 (jb + 1); Xt / Xb = rows 0..63 / 64..127)
// then the block back substitution L^T x = z from the bottom, as in chol_tc.cuh.
#pragma once

#include "chol_tc.cuh"

namespace lk {
namespace ctc128 {

using ctc::elect_one;
using ctc::mbar_test;
using ctc::rcp_fast;
using ctc::tmem_ld16;
using ctc::tmem_st16;
using ctc::umma_tf32;

constexpr int NSYS = 2;
constexpr int LDT = 20;                                   // row stride (floats) of a 16x16 block
constexpr int TILE_BYTES = 128 * 16 * 4;                  // one operand tile: 128 rows x 16 tf32, K-major, no swizzle
constexpr int TILES_BYTES = NSYS * 2 * TILE_BYTES;        // [system][hi, lo] = 32 KB
constexpr int LT_FLOATS = NSYS * 16 * LDT;                // L_jj^T of the current step
constexpr int LD_FLOATS = NSYS * 8 * 16 * LDT;            // L_jj of every step (back substitution)
constexpr int INVD_FLOATS = NSYS * 128;
constexpr int ZB_FLOATS = NSYS * 16;
constexpr int TSUM_FLOATS = 8 * 8 * NSYS * 16;            // [row block][column block][system][16]
constexpr int WS_BYTES = TILES_BYTES + (LT_FLOATS + LD_FLOATS + INVD_FLOATS + ZB_FLOATS + TSUM_FLOATS) * 4 + 16;

struct Workspace {
    unsigned char *tiles;  // 128-byte aligned
    float *lt, *ld, *invd, *zb, *tsum;
    int *bad;  // [NSYS]
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

// TMEM column of block column jb of the half's slot, system p (relative to the CTA's base column)
__device__ __forceinline__ uint32_t blk_col(int p, int half, int jb) { return (uint32_t)(128 * p + 64 * half + 16 * jb); }

// Solve the two systems.  On entry the lower triangles are in TMEM (layout above, tmem_base = column 0,
// lane field 0) and yv[p] holds the right-hand-side entry of this lane's row R of system p; on exit
// yv[p] holds the solution entry.  `bar`: two mbarriers with count NSYS (one commit per issuing warp and update
// pass), `par` their running phase parity.  ws.bad[p] is set when a pivot of system p was not positive (caller
// zeroes it).  All 128 threads must call.
__device__ __forceinline__ void solve2(const uint32_t tmem_base, float (&yv)[NSYS], const Workspace &ws, uint64_t *bar,
                                       uint32_t &par, const int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    const int r = lane & 15, h = lane >> 4;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
    const int R = 64 * h + 16 * warp + r;

#pragma unroll 1
    for (int j = 0; j < 8; j++) {
        const int half = j >> 2, jb = j & 3;
        if (warp == jb) {
            // ---- diagonal blocks of both systems (the half that owns them; the other half computes on
            //      whatever it loaded and publishes nothing) ----
            float a[NSYS][16];
#pragma unroll
            for (int p = 0; p < NSYS; p++) tmem_ld16(lane_taddr + blk_col(p, half, jb), a[p]);
            const bool own = h == half;
            float inv_mine[NSYS] = {0.0f, 0.0f};
            bool bad[NSYS] = {false, false};
            float zt[NSYS] = {yv[0], yv[1]};
#pragma unroll
            for (int k = 0; k < 16; k++) {
#pragma unroll
                for (int p = 0; p < NSYS; p++) {
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
            if (own) {
#pragma unroll
                for (int p = 0; p < NSYS; p++) {
                    yv[p] = zt[p];
                    float *lt = ws.lt + p * 16 * LDT;
#pragma unroll
                    for (int c = 0; c < 16; c++) lt[c * LDT + r] = a[p][c];  // Lt[c][r] = L[r][c]
                    float *ldr = ws.ld + ((p * 8 + j) * 16 + r) * LDT;       // row r of L_jj
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        *reinterpret_cast<float4 *>(ldr + 4 * q) =
                            make_float4(a[p][4 * q], a[p][4 * q + 1], a[p][4 * q + 2], a[p][4 * q + 3]);
                    ws.invd[p * 128 + 16 * j + r] = inv_mine[p];
                    ws.zb[p * 16 + r] = zt[p];
                    if (bad[p] && r == 0) ws.bad[p] = 1;
                }
            }
        }
        if (j == 7) break;
        tmem_fence_before();
        __syncthreads();
        tmem_fence_after();
        // ---- rows below the diagonal block: X L_jj^T = A_Rj ----
        // tcgen05.ld / .st are warp-wide (.sync.aligned): a warp takes part as a whole when any of its rows
        // lies below the block; the lanes whose rows do not (`below` false) run the same arithmetic on
        // entries nobody reads again (the diagonal block itself, the never-used upper triangle, the unused
        // half of slot P1), keep their right-hand side and leave their operand-tile rows alone.
        const bool below = half == 0 ? (h == 1 || warp > jb) : (h == 1 && warp > jb);
        if (half == 0 || warp > jb) {
            float a[NSYS][16];
#pragma unroll
            for (int p = 0; p < NSYS; p++) tmem_ld16(lane_taddr + blk_col(p, half, jb), a[p]);
            float iv[NSYS][16];
#pragma unroll
            for (int p = 0; p < NSYS; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 t = *reinterpret_cast<const float4 *>(ws.invd + p * 128 + 16 * j + 4 * q);
                    iv[p][4 * q] = t.x, iv[p][4 * q + 1] = t.y, iv[p][4 * q + 2] = t.z, iv[p][4 * q + 3] = t.w;
                }
            float4 row[2][NSYS][4];
#pragma unroll
            for (int p = 0; p < NSYS; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) row[0][p][q] = *reinterpret_cast<const float4 *>(ws.lt + p * 16 * LDT + 4 * q);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k + 1 < 16) {
#pragma unroll
                    for (int p = 0; p < NSYS; p++)
#pragma unroll
                        for (int q = (k + 2) / 4; q < 4; q++)
                            row[(k + 1) & 1][p][q] =
                                *reinterpret_cast<const float4 *>(ws.lt + p * 16 * LDT + (k + 1) * LDT + 4 * q);
                }
#pragma unroll
                for (int p = 0; p < NSYS; p++) {
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
            // the operand tiles are read by the previous step's trailing-update instructions until they complete:
            // bar[1] collects the commits issued after their second pass (its phase parity follows bar[0]'s)
            if (j > 0) {
                while (!mbar_test(bar + 1, par ^ 1u)) {
                }
            }
#pragma unroll
            for (int p = 0; p < NSYS; p++) {
                float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 u = *reinterpret_cast<const float4 *>(ws.zb + p * 16 + 4 * q);
                    acc0 = fmaf(a[p][4 * q + 0], u.x, acc0);
                    acc1 = fmaf(a[p][4 * q + 1], u.y, acc1);
                    acc0 = fmaf(a[p][4 * q + 2], u.z, acc0);
                    acc1 = fmaf(a[p][4 * q + 3], u.w, acc1);
                }
                if (below) yv[p] -= acc0 + acc1;
                tmem_st16(lane_taddr + blk_col(p, half, jb), a[p]);  // the back substitution reads it there
                unsigned char *thi = ws.tiles + (p * 2) * TILE_BYTES + (R >> 3) * 512 + (R & 7) * 16;
                unsigned char *tlo = thi + TILE_BYTES;
#pragma unroll
                for (int q = 0; below && q < 4; q++) {
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
        fence_proxy_async();
        tmem_fence_before();
        __syncthreads();
        // ---- trailing updates: warp p issues for system p ----
        const int warp_u = __shfl_sync(FULL, warp, 0);
        if (warp_u < NSYS && elect_one()) {
            tmem_fence_after();
            const int p = warp_u;
            const int c0 = 16 * (jb + 1);
            const uint32_t hi = smem_u32(ws.tiles + (p * 2) * TILE_BYTES);
            const uint32_t lo = hi + TILE_BYTES;
            const uint32_t bot = 8 * 512;  // rows 64..127 of a tile
            auto update = [&](uint32_t d, uint32_t a_off, uint32_t b_off, int n) {
                if (n <= 0) return;
                const uint32_t idesc = ctc::IDESC_TF32 | ((uint32_t)(n >> 3) << 17);
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    const uint64_t a_hi = ctc::DESC_KMAJOR | (uint64_t)(((hi + a_off + kk * 256) >> 4) & 0x3fffu);
                    const uint64_t a_lo = ctc::DESC_KMAJOR | (uint64_t)(((lo + a_off + kk * 256) >> 4) & 0x3fffu);
                    const uint64_t b_hi = ctc::DESC_KMAJOR | (uint64_t)(((hi + b_off + kk * 256) >> 4) & 0x3fffu);
                    const uint64_t b_lo = ctc::DESC_KMAJOR | (uint64_t)(((lo + b_off + kk * 256) >> 4) & 0x3fffu);
                    umma_tf32(d, a_hi, b_hi, idesc);
                    umma_tf32(d, a_hi, b_lo, idesc);
                    umma_tf32(d, a_lo, b_hi, idesc);
                }
            };
            const uint32_t top_d = tmem_base + (uint32_t)(128 * p);                    // lanes 0..15 of every quarter
            const uint32_t bot_d = tmem_base + ((uint32_t)16 << 16) + (uint32_t)(128 * p);  // lanes 16..31
            const uint32_t boff_top = (uint32_t)(c0 >> 3) * 512;        // B = X rows c0.. of the top half
            const uint32_t boff_bot = bot + (uint32_t)(c0 >> 3) * 512;  // B = X rows 64 + c0..
            // look-ahead (as in chol_tc.cuh): the 16 columns the next step works on first, the barrier after
            // them, the rest of the trailing update in their shadow, committed to the second barrier
            auto update16 = [&](uint32_t d, uint32_t a_off, uint32_t b_off, int n, bool first) {
                if (first)
                    update(d, a_off, b_off, n > 16 ? 16 : n);
                else if (n > 16)
                    update(d + 16, a_off, b_off + 2 * 512, n - 16);
            };
            for (int pass = 0; pass < 2; pass++) {
                const bool first = pass == 0;
                if (half == 0) {
                    if (jb < 3) {
                        // next step: block column jb + 1 of the left panel (top and bottom rows)
                        update16(top_d + c0, 0, boff_top, 64 - c0, first);    // G11[:, c0:] -= Xt Xt[c0:]^T
                        update16(bot_d + c0, bot, boff_top, 64 - c0, first);  // G21[:, c0:] -= Xb Xt[c0:]^T
                        if (!first) update(bot_d + 64, bot, bot, 64);         // G22        -= Xb Xb^T
                    } else {
                        // jb == 3: the next step starts on G22 — its first 16 columns are the urgent part
                        update16(bot_d + 64, bot, bot, 64, first);
                    }
                } else {
                    update16(bot_d + 64 + c0, bot, boff_bot, 64 - c0, first);  // G22[:, c0:] -= Xb Xb[c0:]^T
                }
                if (first) umma_commit(bar);
            }
            umma_commit(bar + 1);  // second pass done: the tiles may be rewritten once this phase completes
        }
        __syncwarp();
        // only the warp that factors the next diagonal block needs the updated accumulators now
        if (warp == ((j + 1) & 3)) {
            while (!mbar_test(bar, par)) {
            }
            tmem_fence_after();
        }
        par ^= 1u;
    }

    // ---- back substitution L^T x = z; yv holds z.  Row block b = 4 h + w; ts[(b * 8 + jj) * NSYS + p][c] is what
    //      row block b (its x final) takes off the right-hand side of block jj < b. ----
    auto contribute = [&](const int jj, const bool active) {
        float v[NSYS][16];
#pragma unroll
        for (int p = 0; p < NSYS; p++) tmem_ld16(lane_taddr + blk_col(p, jj >> 2, jj & 3), v[p]);
#pragma unroll
        for (int p = 0; p < NSYS; p++)
#pragma unroll
            for (int c = 0; c < 16; c++) v[p][c] *= yv[p];
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1) {
            const bool up = (r & w) != 0;
#pragma unroll
            for (int i = 0; i < w; i++) {
#pragma unroll
                for (int p = 0; p < NSYS; p++) {
                    const float send = up ? v[p][i] : v[p][i + w];
                    const float keep = up ? v[p][i + w] : v[p][i];
                    v[p][i] = keep + __shfl_xor_sync(FULL, send, w, 16);
                }
            }
        }
        if (active) {
            const int b = 4 * h + warp;
#pragma unroll
            for (int p = 0; p < NSYS; p++) ws.tsum[((b * 8 + jj) * NSYS + p) * 16 + r] = v[p][0];
        }
    };
#pragma unroll 1
    for (int j = 7; j >= 0; j--) {
        const int half = j >> 2, jb = j & 3;
        const bool own = warp == jb && h == half;
        if (warp == jb) {
            // every lane of the owning half solves the whole 16x16 transposed system redundantly
            float x[NSYS][16];
#pragma unroll
            for (int p = 0; p < NSYS; p++) {
                float rhs = yv[p];
                if (own)
                    for (int b = j + 1; b < 8; b++) rhs -= ws.tsum[((b * 8 + j) * NSYS + p) * 16 + r];
#pragma unroll
                for (int k = 0; k < 16; k++) x[p][k] = __shfl_sync(FULL, rhs, k, 16);
            }
#pragma unroll
            for (int k = 15; k >= 0; k--) {
#pragma unroll
                for (int p = 0; p < NSYS; p++) {
                    const float *ldk = ws.ld + ((p * 8 + j) * 16 + k) * LDT;  // row k of L_jj: L[k][m], m < k
                    const float xk = x[p][k] * ws.invd[p * 128 + 16 * j + k];
                    if (own && r == k) yv[p] = xk;
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
            if (j > 0) contribute(j - 1, own);
        }
        if (j == 0) break;
        __syncthreads();
        if (warp == jb)
            for (int jj = j - 2; jj >= 0; jj--) contribute(jj, own);
    }
}

}  // namespace ctc128
}  // namespace lk
