// als_tcr.cu — tensor-core ALS half-epoch, second generation: the Gram comes out
// of TMEM straight into a *register-resident* Cholesky.
//
// Same contract and Gram pipeline as als_tc.cu (k = 64, bf16 gathered rows,
// unweighted / uniformly weighted Gram on tcgen05.mma; see that file's header).
// What changes is the solve, which dominated als_tc.cu (58 % of the user
// half-epoch, ~10 k issue slots per 64x64 system when one warp factors a matrix
// held in shared memory):
//
//   * tcgen05.ld hands every lane one full row of a system (for M = 64 the rows
//     16w..16w+15 of an accumulator sit in the TMEM lanes of warp w; with two
//     accumulators interleaved in lanes 0-15 / 16-31 a warp holds 16 rows of two
//     systems).  The four warps of the CTA therefore own, row by row, a pair of
//     64x64 systems entirely in registers (64 floats per lane).
//   * Right-looking Cholesky with 4-column panels on that layout: the 4x4
//     diagonal block and the four right-hand-side pivots travel through a 160-byte
//     shared buffer, every lane factors the block redundantly, solves its own
//     row's panel entries, publishes them (16 B per row) and then applies the
//     rank-4 update to its own row out of registers: 1 broadcast LDS.128 + 4 FFMA
//     per updated element instead of LDS + 16 FFMA + STS per 4x4 tile.  Columns
//     right of a warp's last row are skipped warp-uniformly (upper triangle).
//     The forward substitution rides along in the panel loop.
//   * The factor is then written once, lower triangle only, in a ragged packed
//     layout (row i holds 4*(i/4+1) floats; 8.7 KB per system, aliased on the
//     dead gather ring) and each warp back-substitutes one system.
//
// No 17 KB-per-warp system buffers any more: the CTA needs 55 KB of shared memory
// and 128 registers per thread, so four CTAs (16 warps, 512 TMEM columns) fit an SM.
// Rows split over several chunks only deposit their partial Gram here; the
// reduction + solve of those (few hundred) rows is done by als_split_fixup_kernel.

#include "tc_common.cuh"

namespace lk {

namespace tcr {
constexpr int KP = 64;
constexpr int WARPS = 4;
constexpr int NT = WARPS * 32;
constexpr int STAGE_ROWS = 32;
constexpr int NSTAGE = 3;
constexpr int ROW_BYTES = KP * 2;
constexpr int STAGE_BYTES = STAGE_ROWS * ROW_BYTES;  // 4096
constexpr int RING_BYTES = NSTAGE * STAGE_BYTES;     // 12288 per warp
constexpr int PK = 8 * (KP / 4) * (KP / 4 + 1);      // 2176 floats: packed lower triangle
static_assert(PK * 4 <= RING_BYTES, "a packed factor must fit the warp's dead gather ring");
constexpr int TMEM_COLS = 128;
constexpr int SLOTF = KP * KP + KP;
// shared memory map (after 1 KB alignment):
//   [0, 4*RING)            gather rings, later the packed factors Lp[4]
//   ys[4][64] zs[4][64] dinv[4][64]   right-hand sides, z / x, 1/diag
//   px[2][64] float4       panel entries of the current pair
//   dgb[2][16] dyb[2][4]   diagonal block + pivots' rhs of the current pair
//   mbarriers, tmem ptr, misc
constexpr int OFF_YS = WARPS * RING_BYTES;
constexpr int OFF_ZS = OFF_YS + WARPS * KP * 4;
constexpr int OFF_DINV = OFF_ZS + WARPS * KP * 4;
constexpr int OFF_PX = OFF_DINV + WARPS * KP * 4;
constexpr int OFF_DG = OFF_PX + 2 * KP * 16;
constexpr int OFF_DY = OFF_DG + 2 * 16 * 4;
constexpr int OFF_BARS = OFF_DY + 2 * 4 * 4;
constexpr int OFF_TMEM = OFF_BARS + (WARPS * NSTAGE + WARPS) * 8;
constexpr int OFF_MISC = OFF_TMEM + 16;
constexpr int SMEM_BYTES = 1024 + OFF_MISC + 64 * 4;
using tcd::DESC_HI;
using tcd::DESC_LBO;

__host__ __device__ constexpr int rowoff(int i) { return 8 * (i >> 2) * ((i >> 2) + 1) + (i & 3) * 4 * ((i >> 2) + 1); }
}  // namespace tcr

// Back substitution L^T x = z for one 64x64 system by one warp; L in the ragged
// packed layout, z in zs (overwritten with x), 1/diag in dinv.
__device__ __forceinline__ void back_subst_packed64(const float *Lp, float *zs, const float *dinv, const int lane)
{
    float yv[2] = {zs[lane], zs[lane + 32]};
    __syncwarp();
    for (int j0 = 60; j0 >= 0; j0 -= 4) {
        const int b = j0 >> 2;
        const float *blk = Lp + 8 * b * (b + 1);  // rows j0..j0+3, each 4*(b+1) long
        const int s = 4 * (b + 1);
        const float l10 = blk[1 * s + j0];
        const float l20 = blk[2 * s + j0], l21 = blk[2 * s + j0 + 1];
        const float l30 = blk[3 * s + j0], l31 = blk[3 * s + j0 + 1], l32 = blk[3 * s + j0 + 2];
        const float x3 = zs[j0 + 3] * dinv[j0 + 3];
        const float x2 = (zs[j0 + 2] - l32 * x3) * dinv[j0 + 2];
        const float x1 = (zs[j0 + 1] - l21 * x2 - l31 * x3) * dinv[j0 + 1];
        const float x0 = (zs[j0] - l10 * x1 - l20 * x2 - l30 * x3) * dinv[j0];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int i = lane + 32 * q;
            if (i < j0) {
                yv[q] -= blk[i] * x0 + blk[s + i] * x1 + blk[2 * s + i] * x2 + blk[3 * s + i] * x3;
                if (i >= j0 - 4) zs[i] = yv[q];
            } else if (i < j0 + 4) {
                const int r = i - j0;
                yv[q] = r == 0 ? x0 : r == 1 ? x1 : r == 2 ? x2 : x3;
            }
        }
        __syncwarp();
    }
    zs[lane] = yv[0];
    zs[lane + 32] = yv[1];
    __syncwarp();
}

template <int MODE>
__global__ void __launch_bounds__(tcr::NT, 4) als_tcr_kernel(lk_als_args a)
{
    using namespace tcr;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    unsigned char *wreg = base + warp * RING_BYTES;  // this warp's gather ring
    float *ys_all = reinterpret_cast<float *>(base + OFF_YS);
    float *zs_all = reinterpret_cast<float *>(base + OFF_ZS);
    float *dinv_all = reinterpret_cast<float *>(base + OFF_DINV);
    float4 *px_all = reinterpret_cast<float4 *>(base + OFF_PX);
    float *dg_all = reinterpret_cast<float *>(base + OFF_DG);
    float *dy_all = reinterpret_cast<float *>(base + OFF_DY);
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + OFF_BARS);
    uint64_t *stage_free = bars + warp * NSTAGE;
    uint64_t *acc_full = bars + WARPS * NSTAGE;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(base + OFF_TMEM);
    int *s_misc = reinterpret_cast<int *>(base + OFF_MISC);  // [0] group; [8+8c..] chunk metadata; [48+c] bad flag

    const __nv_bfloat16 *__restrict__ other = reinterpret_cast<const __nv_bfloat16 *>(a.d_other);
    constexpr int k = KP;

    if (tid == 0) {
        for (int i = 0; i < WARPS * NSTAGE + WARPS; i++) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem_base = *s_tmem;
    // accumulator of warp w: columns 64*(w/2), lanes 16*(w%2) of every 32-lane quarter
    const uint32_t my_acc = tmem_base + ((uint32_t)((warp & 1) * 16) << 16) + (uint32_t)((warp >> 1) * 64);

    long long t_prev = clock64();
    auto prof = [&](int idx) {
        if (a.d_prof != nullptr && tid == 0) {
            const long long t = clock64();
            atomicAdd(a.d_prof + idx, (unsigned long long)(t - t_prev));
            t_prev = t;
        }
    };

    uint32_t free_par = 0;  // bit s: parity of the number of commits issued on stage_free[s]
    uint32_t full_par = 0;  // bit c: parity of the number of commits seen on acc_full[c]

    for (;;) {
        if (tid == 0) s_misc[0] = atomicAdd(a.d_work_counter, 1);
        __syncthreads();
        const int64_t g = s_misc[0];
        __syncthreads();
        if (g * WARPS >= a.n_chunks) break;
        prof(0);
        const int64_t ci = g * WARPS + warp;
        const bool active = ci < a.n_chunks;
        int row = -1, begin = 0, len = 0, nparts = 1, slot0 = 0, part = 0;
        if (active) {
            const int4 c0 = __ldg(reinterpret_cast<const int4 *>(a.d_chunks) + 2 * ci);
            const int4 c1 = __ldg(reinterpret_cast<const int4 *>(a.d_chunks) + 2 * ci + 1);
            row = c0.x, begin = c0.y, len = c0.z, nparts = c0.w;
            slot0 = c1.x, part = c1.y;
        }
        const bool has_gram = active && len > 0;
        int n_row = 0;
        if (active) n_row = __ldg(a.d_indptr + row + 1) - __ldg(a.d_indptr + row);
        if (lane == 0) {
            int *m = s_misc + 8 + 8 * warp;
            m[0] = has_gram ? 1 : 0;
            m[1] = nparts;
            m[2] = slot0 + part;
            m[3] = n_row;
            m[4] = row;
            s_misc[48 + warp] = 0;
        }

        // ------------------------------------------------------------------
        // phase 1: gather -> tcgen05.mma, y on the side (as in als_tc.cu)
        // ------------------------------------------------------------------
        float y0 = 0.0f, y1 = 0.0f;  // features 2*lane, 2*lane+1
        if (has_gram) {
            const int n_it = (len + STAGE_ROWS - 1) / STAGE_ROWS;
            const int32_t *cols = a.d_cols + begin;
            const float *vals = a.d_vals + begin;
            float vst[NSTAGE];
#pragma unroll
            for (int s = 0; s < NSTAGE; s++) vst[s] = 0.0f;

            auto fetch = [&](int it, int &c, float &v) {
                const int idx = it * STAGE_ROWS + lane;
                if (it < n_it && idx < len) {
                    c = __ldg(cols + idx);
                    v = __ldg(vals + idx);
                } else {
                    c = 0;
                    v = 0.0f;
                }
            };
            auto issue = [&](int it, int s, int c) {
                if (it < n_it) {
                    const int nrows = min(STAGE_ROWS, len - it * STAGE_ROWS);
                    const int npad = (nrows + 15) & ~15;
                    const uint32_t sbase = smem_u32(wreg + s * STAGE_BYTES);
                    const int chunk = lane & 7;
#pragma unroll
                    for (int t = 0; t < STAGE_ROWS / 4; t++) {
                        const int r = 4 * t + (lane >> 3);
                        const int cr = __shfl_sync(FULL, c, r);
                        const uint32_t dst = sbase + r * ROW_BYTES + ((chunk ^ (r & 7)) << 4);
                        if (r < nrows) {
                            cp_async16(dst, other + (size_t)cr * k + chunk * 8);
                        } else if (r < npad) {
                            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0) : "memory");
                        }
                    }
                }
                cp_async_commit();
            };

            int cpre[NSTAGE - 1];
            float vpre[NSTAGE - 1];
#pragma unroll
            for (int j = 0; j < NSTAGE - 1; j++) fetch(j, cpre[j], vpre[j]);
#pragma unroll
            for (int j = 0; j < NSTAGE - 1; j++) {
                issue(j, j, cpre[j]);
                vst[j] = vpre[j];
            }
            int cnext;
            float vnext;
            fetch(NSTAGE - 1, cnext, vnext);

            for (int it0 = 0; it0 < n_it; it0 += NSTAGE) {
#pragma unroll
                for (int s = 0; s < NSTAGE; s++) {
                    const int it = it0 + s;
                    if (it < n_it) {
                        const int sj = (s + NSTAGE - 1) % NSTAGE;  // constant after unrolling
                        const int j = it + NSTAGE - 1;
                        if (j < n_it && j >= NSTAGE) mbar_wait(&stage_free[sj], ((free_par >> sj) & 1u) ^ 1u);
                        issue(j, sj, cnext);
                        if (j < n_it) vst[sj] = vnext;
                        fetch(j + 1, cnext, vnext);
                        cp_async_wait<NSTAGE - 1>();
                        fence_proxy_async();
                        __syncwarp();
                        const int nrows = min(STAGE_ROWS, len - it * STAGE_ROWS);
                        if (lane == 0) {
                            tmem_fence_after();
                            const uint32_t sbase = smem_u32(wreg + s * STAGE_BYTES);
                            const int nk = (nrows + 15) >> 4;
                            for (int kk = 0; kk < nk; kk++) {
                                const uint64_t desc =
                                    DESC_HI | DESC_LBO | (uint64_t)(((sbase + kk * 2048) >> 4) & 0x3fffu);
                                umma_bf16_64x64x16(my_acc, desc, (it > 0 || kk > 0) ? 1u : 0u);
                            }
                            umma_commit(&stage_free[s]);
                            if (it == n_it - 1) umma_commit(&acc_full[warp]);
                        }
                        free_par ^= (1u << s);
                        {
                            const unsigned char *st = wreg + s * STAGE_BYTES;
                            const int chunk = lane >> 2, within = (lane & 3) * 4;
                            for (int r = 0; r < nrows; r++) {
                                float w = __shfl_sync(FULL, vst[s], r);
                                if constexpr (MODE == LK_ALS_IMPLICIT) w += 1.0f;
                                const uint32_t two = *reinterpret_cast<const uint32_t *>(
                                    st + r * ROW_BYTES + ((chunk ^ (r & 7)) << 4) + within);
                                y0 = fmaf(__uint_as_float(two << 16), w, y0);
                                y1 = fmaf(__uint_as_float(two & 0xffff0000u), w, y1);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            cp_async_wait<0>();
        }
        // right-hand side of this warp's chunk: to shared memory, or to the partial slot
        if (has_gram) {
            if (nparts == 1) {
                *reinterpret_cast<float2 *>(ys_all + warp * KP + 2 * lane) = make_float2(y0, y1);
            } else {
                float *slot = a.d_partials + (size_t)(slot0 + part) * SLOTF + KP * KP;
                __stcg(reinterpret_cast<float2 *>(slot) + lane, make_float2(y0, y1));
            }
        }
        __syncthreads();  // every warp has issued its MMAs; metadata and rhs are visible
        prof(1);

        int gram[WARPS], parts[WARPS], slotc[WARPS], nrowc[WARPS];
#pragma unroll
        for (int c = 0; c < WARPS; c++) {
            const int *m = s_misc + 8 + 8 * c;
            gram[c] = m[0], parts[c] = m[1], slotc[c] = m[2], nrowc[c] = m[3];
        }
#pragma unroll
        for (int c = 0; c < WARPS; c++) {
            if (gram[c]) {
                full_par ^= (1u << c);
                mbar_wait(&acc_full[c], ((full_par >> c) & 1u) ^ 1u);
            }
        }
        tmem_fence_after();
        prof(2);

        // ------------------------------------------------------------------
        // phase 2: pairs of systems, one row per lane, factored in registers
        // ------------------------------------------------------------------
        const int sys = lane >> 4;                // which system of the pair this lane serves
        const int gi = 16 * warp + (lane & 15);   // the row of that system held by this lane
#pragma unroll 1
        for (int pp = 0; pp < 2; pp++) {
            const int c = 2 * pp + sys;
            if (!(gram[2 * pp] || gram[2 * pp + 1])) continue;  // uniform
            float av[KP];
            {
                uint32_t r[64];
                tmem_ld_32x32b_x64(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(pp * 64), r);
                if (gram[c] && parts[c] > 1) {
                    // chunk of a split row: deposit the partial Gram, als_split_fixup_kernel finishes it
                    float *slot = a.d_partials + (size_t)slotc[c] * SLOTF + gi * KP;
#pragma unroll
                    for (int q = 0; q < 16; q++)
                        __stcg(reinterpret_cast<float4 *>(slot) + q,
                               make_float4(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1]),
                                           __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
                }
                if constexpr (MODE == LK_ALS_IMPLICIT) {
                    const float v = a.uniform_val;
                    const float4 *ot = reinterpret_cast<const float4 *>(a.d_otor + gi * k);
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        const float4 o = __ldg(ot + q);  // a = otor + v * mtm (implicit.rs:115)
                        av[4 * q + 0] = fmaf(v, __uint_as_float(r[4 * q + 0]), o.x);
                        av[4 * q + 1] = fmaf(v, __uint_as_float(r[4 * q + 1]), o.y);
                        av[4 * q + 2] = fmaf(v, __uint_as_float(r[4 * q + 2]), o.z);
                        av[4 * q + 3] = fmaf(v, __uint_as_float(r[4 * q + 3]), o.w);
                    }
                } else {
                    const float regn = a.reg * (float)nrowc[c];  // explicit.rs:106-108
#pragma unroll
                    for (int j = 0; j < KP; j++) av[j] = __uint_as_float(r[j]) + (j == gi ? regn : 0.0f);
                }
            }
            const bool solvable = gram[c] && parts[c] == 1;
            if (!solvable) {
                // keep the arithmetic finite for a lane pair-mate that has nothing to solve
#pragma unroll
                for (int j = 0; j < KP; j++) av[j] = (j == gi) ? 1.0f : 0.0f;
            }
            float yv = solvable ? ys_all[c * KP + gi] : 0.0f;
            float mydinv = 1.0f;
            bool bad = false;
            float *dg = dg_all + sys * 16;
            float *dy = dy_all + sys * 4;
            float4 *pxs = px_all + sys * KP;
            if (gi < 4) {
                *reinterpret_cast<float4 *>(dg + gi * 4) = make_float4(av[0], av[1], av[2], av[3]);
                dy[gi] = yv;
            }
            __syncthreads();
#pragma unroll
            for (int j0 = 0; j0 < KP; j0 += 4) {
                const float4 g0 = *reinterpret_cast<const float4 *>(dg + 0);
                const float4 g1 = *reinterpret_cast<const float4 *>(dg + 4);
                const float4 g2 = *reinterpret_cast<const float4 *>(dg + 8);
                const float4 g3 = *reinterpret_cast<const float4 *>(dg + 12);
                const float4 yp = *reinterpret_cast<const float4 *>(dy);
                const float a00 = g0.x, a10 = g1.x, a11 = g1.y, a20 = g2.x, a21 = g2.y, a22 = g2.z;
                const float a30 = g3.x, a31 = g3.y, a32 = g3.z, a33 = g3.w;
                const float i0 = rsqrt_nr(a00), l00 = a00 * i0;
                const float l10 = a10 * i0, l20 = a20 * i0, l30 = a30 * i0;
                const float d1 = a11 - l10 * l10;
                const float i1 = rsqrt_nr(d1), l11 = d1 * i1;
                const float l21 = (a21 - l20 * l10) * i1, l31 = (a31 - l30 * l10) * i1;
                const float d2 = a22 - l20 * l20 - l21 * l21;
                const float i2 = rsqrt_nr(d2), l22 = d2 * i2;
                const float l32 = (a32 - l30 * l20 - l31 * l21) * i2;
                const float d3 = a33 - l30 * l30 - l31 * l31 - l32 * l32;
                const float i3 = rsqrt_nr(d3), l33 = d3 * i3;
                bad |= !(a00 > 0.0f && d1 > 0.0f && d2 > 0.0f && d3 > 0.0f);
                const float z0 = yp.x * i0;
                const float z1 = (yp.y - l10 * z0) * i1;
                const float z2 = (yp.z - l20 * z0 - l21 * z1) * i2;
                const float z3 = (yp.w - l30 * z0 - l31 * z1 - l32 * z2) * i3;

                float x0 = 0.0f, x1 = 0.0f, x2 = 0.0f, x3 = 0.0f;
                if (gi >= j0 + 4) {
                    x0 = av[j0] * i0;
                    x1 = (av[j0 + 1] - x0 * l10) * i1;
                    x2 = (av[j0 + 2] - x0 * l20 - x1 * l21) * i2;
                    x3 = (av[j0 + 3] - x0 * l30 - x1 * l31 - x2 * l32) * i3;
                    yv -= x0 * z0 + x1 * z1 + x2 * z2 + x3 * z3;
                } else if (gi >= j0) {
                    const int r = gi - j0;
                    x0 = r == 0 ? l00 : r == 1 ? l10 : r == 2 ? l20 : l30;
                    x1 = r == 0 ? 0.f : r == 1 ? l11 : r == 2 ? l21 : l31;
                    x2 = r <= 1 ? 0.f : r == 2 ? l22 : l32;
                    x3 = r <= 2 ? 0.f : l33;
                    yv = r == 0 ? z0 : r == 1 ? z1 : r == 2 ? z2 : z3;
                    mydinv = r == 0 ? i0 : r == 1 ? i1 : r == 2 ? i2 : i3;
                }
                av[j0] = x0, av[j0 + 1] = x1, av[j0 + 2] = x2, av[j0 + 3] = x3;  // row of L
                pxs[gi] = make_float4(x0, x1, x2, x3);
                __syncthreads();  // panel entries published; the diagonal buffer has been read by all
                // rank-4 update of this lane's row; a warp never needs columns right of its last row
#pragma unroll
                for (int blk = (j0 + 4) >> 4; blk < 4; blk++) {
                    if (warp >= blk) {
#pragma unroll
                        for (int cc = (blk * 16 > j0 + 4 ? blk * 16 : j0 + 4); cc < blk * 16 + 16; cc++) {
                            const float4 Lc = pxs[cc];
                            av[cc] -= x0 * Lc.x + x1 * Lc.y + x2 * Lc.z + x3 * Lc.w;
                        }
                    }
                }
                if (j0 + 4 < KP && (gi >> 2) == ((j0 + 4) >> 2)) {
                    // rows of the next diagonal block are final: publish block and pivots' rhs
                    *reinterpret_cast<float4 *>(dg + (gi & 3) * 4) =
                        make_float4(av[j0 + 4], av[j0 + 5], av[j0 + 6], av[j0 + 7]);
                    dy[gi & 3] = yv;
                }
                __syncthreads();
            }
            // the factor row (lower part), z and 1/diag go to shared memory for the back substitution
            if (solvable) {
                float *Lrow = reinterpret_cast<float *>(base + c * RING_BYTES) + rowoff(gi);
#pragma unroll
                for (int q = 0; q < 16; q++)
                    if (q <= (gi >> 2))
                        *reinterpret_cast<float4 *>(Lrow + 4 * q) =
                            make_float4(av[4 * q], av[4 * q + 1], av[4 * q + 2], av[4 * q + 3]);
                zs_all[c * KP + gi] = yv;
                dinv_all[c * KP + gi] = mydinv;
                if (bad && gi == 0) s_misc[48 + c] = 1;
            }
        }
        tmem_fence_before();
        __threadfence();  // partial slots (cheap when nothing was written)
        __syncthreads();
        prof(3);

        // ------------------------------------------------------------------
        // phase 3: one warp back-substitutes one system and writes the row
        // ------------------------------------------------------------------
        if (active) {
            float *thisrow = a.d_this + (size_t)row * k;
            if (has_gram && nparts == 1) {
                float *zs = zs_all + warp * KP;
                back_subst_packed64(reinterpret_cast<const float *>(base + warp * RING_BYTES), zs,
                                    dinv_all + warp * KP, lane);
                write_row<KP, 1>(a, row, thisrow, zs, lane, s_misc[48 + warp] != 0);
            } else if (len == 0 && nparts == 1) {
                for (int i = lane; i < k; i += 32) {  // empty row: x = 0, no delta (implicit.rs:98-101)
                    thisrow[i] = 0.0f;
                    for (int rr = 0; rr < a.n_replicas; rr++)
                        a.d_replicas[rr][(size_t)(a.replica_row0 + row) * k + i] = 0.0f;
                }
            }
        }
        prof(5);
        __syncthreads();  // the factors alias the gather rings of the next group
        prof(6);
    }

    tmem_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)tcr::TMEM_COLS)
                     : "memory");
    }
}

// Rows split over several chunks: sum the partial Grams in slot order
// (deterministic), add OtOr / reg*n*I, solve, write.  One warp per split row.
template <int MODE>
__global__ void __launch_bounds__(128) als_split_fixup_kernel(lk_als_args a)
{
    constexpr int KP = 64, LDA = KP + 4, SLOTF = KP * KP + KP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *As = reinterpret_cast<float *>(smem_raw) + warp * (KP * LDA + 2 * KP);
    float *ys = As + KP * LDA;
    float *dinv = ys + KP;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int64_t ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ci < a.n_chunks; ci += warps_total) {
        const int4 c0 = __ldg(reinterpret_cast<const int4 *>(a.d_chunks) + 2 * ci);
        const int4 c1 = __ldg(reinterpret_cast<const int4 *>(a.d_chunks) + 2 * ci + 1);
        if (c0.w <= 1 || c1.y != 0) continue;  // only the first part of a split row
        const int row = c0.x, nparts = c0.w, slot0 = c1.x;
        const int n_row = __ldg(a.d_indptr + row + 1) - __ldg(a.d_indptr + row);
        const float regn = a.reg * (float)n_row;
        for (int idx = lane; idx < KP * KP; idx += 32) {
            float s = 0.0f;
            for (int p = 0; p < nparts; p++) s += __ldcg(a.d_partials + (size_t)(slot0 + p) * SLOTF + idx);
            const int gi = idx >> 6, gc = idx & 63;
            if constexpr (MODE == LK_ALS_IMPLICIT)
                s = fmaf(a.uniform_val, s, __ldg(a.d_otor + idx));
            else if (gi == gc)
                s += regn;
            As[gi * LDA + gc] = s;
        }
        for (int f = lane; f < KP; f += 32) {
            float s = 0.0f;
            for (int p = 0; p < nparts; p++) s += __ldcg(a.d_partials + (size_t)(slot0 + p) * SLOTF + KP * KP + f);
            ys[f] = s;
        }
        __syncwarp();
        const bool bad = chol_solve<KP, 1>(As, ys, dinv, lane);
        write_row<KP, 1>(a, row, a.d_this + (size_t)row * a.k, ys, lane, bad);
        __syncwarp();
    }
}

template <int MODE>
static int launch_tcr(const lk_als_args &a, cudaStream_t st)
{
    auto kern = als_tcr_kernel<MODE>;
    const int smem = tcr::SMEM_BYTES;
    LK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t groups = (a.n_chunks + tcr::WARPS - 1) / tcr::WARPS;
    const int occ = 512 / tcr::TMEM_COLS;  // 4 CTAs per SM: 55 KB smem, <= 128 registers, 128 TMEM columns each
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)sm_count() * occ, groups));
    kern<<<(unsigned)grid, tcr::NT, smem, st>>>(a);
    LK_CUDA_TRY(cudaGetLastError());
    if (a.n_split_rows > 0) {
        auto fix = als_split_fixup_kernel<MODE>;
        const int fsmem = 4 * (64 * 68 + 128) * 4;
        LK_CUDA_TRY(cudaFuncSetAttribute(fix, cudaFuncAttributeMaxDynamicSharedMemorySize, fsmem));
        const int fgrid = (int)std::max<int64_t>(1, std::min<int64_t>(sm_count() * 2, (a.n_split_rows + 3) / 4 * 4));
        fix<<<fgrid, 128, fsmem, st>>>(a);
        LK_CUDA_TRY(cudaGetLastError());
    }
    return LK_OK;
}

// returns LK_OK when this kernel took the launch, 1 when the caller should use another one
int launch_als_tcr(const lk_als_args &a, cudaStream_t st)
{
    if (a.k != tcr::KP || a.other_dtype != LK_DTYPE_BF16) return 1;
    if (a.mode == LK_ALS_IMPLICIT && !a.vals_uniform) return 1;
    if (reinterpret_cast<uintptr_t>(a.d_other) % 16 != 0) return 1;
    return a.mode == LK_ALS_IMPLICIT ? launch_tcr<LK_ALS_IMPLICIT>(a, st) : launch_tcr<LK_ALS_EXPLICIT>(a, st);
}

}  // namespace lk
