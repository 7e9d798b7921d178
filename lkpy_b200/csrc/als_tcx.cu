// als_tcx.cu — ALS half-epoch on the tensor cores for everything the bf16 fast path
// (als_tc.cu) does not take at k = 64: fp32 gathered rows (the reference's own arithmetic,
// src/accel/als/implicit.rs:87-125 / explicit.rs:80-119) and non-uniform confidence weights
// (use_ratings=True), for fp32 or bf16 rows.
//
//      z_j = s_j * other[c_j, :]      s_j = 1 (uniform weights / explicit) or sqrt(v_j)
//      G   = sum_j z_j z_j^T          (fp32-accurate on the tensor cores, see below)
//      implicit: A = OtOr + c * G     c = v (uniform) or 1;   y = sum_j (v_j + 1) o_j
//      explicit: A = G + reg * n * I                           y = sum_j v_j o_j
//
// fp32-accurate Gram from tf32 MMAs.  Every z is split exactly into hi = z with the low 13
// mantissa bits cleared (a tf32 number) and lo = z - hi (exact in f32; the tensor core keeps its
// top 11 bits), and G is accumulated as hi.hi + hi.lo + lo.hi — three tcgen05.mma.kind::tf32 per
// 8 gathered rows, fp32 accumulation in TMEM.  The dropped lo.lo term and the truncation of lo
// are both ~2^-22 relative per product: the size of the f32 rounding the reference's own sgemm
// makes on every product.  (The same split carries the trailing updates of the in-TMEM Cholesky,
// chol_tc.cuh.)
//
// The rows go through registers: a warp reads 8 rows per stage (4 x LDG.128 per lane: lane =
// (row quad, feature quad), a 4 x 4 micro-block), scales / splits them, accumulates the right-hand
// side on the way (4 FMAs per 4 loaded values — no tensor-core pass, no extra TMEM columns), and
// stores the micro-block transposed, i.e. as 16-byte K-vectors of the K-major no-swizzle operand
// layout (core matrix = 8 features x 4 rows; the 8-feature groups are 272 bytes apart so that the
// transposed 16-byte stores of a quarter-warp hit 32 different banks).  A = Z^T and B = Z are the
// same tile.  Everything after the Gram — OtOr preload, split-row reduction, the blocked Cholesky
// in TMEM with tensor-core trailing updates, write-back into `this` and its peer replicas — is
// shared with als_tc.cu's design (chol_tc.cuh); a CTA is 4 warps working on groups of 4 chunks.

#include <algorithm>

#include "chol_tc.cuh"

namespace lk {

namespace tcx {
constexpr int KP = 64;
constexpr int WARPS = 4;
constexpr int NT = WARPS * 32;
constexpr int STAGE_ROWS = 8;
constexpr int NSTAGE = 3;
constexpr int GROUP_STRIDE = 272;                   // bytes between 8-feature groups (256 + 16 pad)
constexpr int TILE_BYTES = 8 * GROUP_STRIDE;        // 64 features x 8 rows of tf32, K-major
constexpr int STAGE_BYTES = 2 * TILE_BYTES;         // hi, lo
constexpr int RING_BYTES = NSTAGE * STAGE_BYTES;    // per warp
constexpr int TMEM_COLS = 128;
// the block Gauss-Jordan variant of the in-TMEM solve (chol_tc.cuh): ~4 % faster than Cholesky + block back
// substitution at the same measured accuracy (profiles/r02_experiments.md); als_tc.cu switches by option
constexpr bool TCX_GJ = true;
constexpr int SLOTF = KP * KP + KP;
constexpr int WS_ALIGNED = (ctc::WS_BYTES + 127) & ~127;
constexpr int UNION_BYTES = WS_ALIGNED > WARPS * RING_BYTES ? WS_ALIGNED : WARPS * RING_BYTES;
// union (rings | solve workspace), right-hand sides, barriers, misc
constexpr int SMEM_BYTES = 128 /*alignment slack*/ + UNION_BYTES + WARPS * KP * 4 + (WARPS * NSTAGE + WARPS + 2) * 8 +
                           16 + 64 * 4;
// instruction descriptor: D f32, A = B = tf32, both K-major, M = 64, N = 64
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
// K-major, no swizzle: LBO (next core matrix along K: rows 4..7) = 128 B, SBO (next 8-feature group) = 272 B
constexpr uint64_t DESC = (uint64_t(128 >> 4) << 16) | (uint64_t(GROUP_STRIDE >> 4) << 32) | (1ull << 46);
}  // namespace tcx

// 4 consecutive features of a gathered row as f32
__device__ __forceinline__ float4 load_quad(const float *other, int row, int fq)
{
    return __ldg(reinterpret_cast<const float4 *>(other + (size_t)row * tcx::KP) + fq);
}
__device__ __forceinline__ float4 load_quad(const __nv_bfloat16 *other, int row, int fq)
{
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(other + (size_t)row * tcx::KP) + fq);
    return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u), __uint_as_float(t.y << 16),
                       __uint_as_float(t.y & 0xffff0000u));
}

// MODE: LK_ALS_IMPLICIT / LK_ALS_EXPLICIT.  ET: float or __nv_bfloat16 rows of `other`.
// WEIGHTED: implicit mode with per-nonzero confidences (z = sqrt(v) o); otherwise the Gram is
// unweighted and scaled by the uniform confidence afterwards (implicit) or not at all (explicit).
template <int MODE, typename ET, bool WEIGHTED>
__global__ void __launch_bounds__(tcx::NT, 3) als_tcx_kernel(lk_als_args a)
{
    using namespace tcx;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(FULL, tid >> 5, 0);

    unsigned char *ring = base + warp * RING_BYTES;
    float *ys_all = reinterpret_cast<float *>(base + UNION_BYTES);  // [WARPS][KP]
    uint64_t *bars = reinterpret_cast<uint64_t *>(ys_all + WARPS * KP);
    uint64_t *stage_free = bars + warp * NSTAGE;
    uint64_t *acc_full = bars + WARPS * NSTAGE;
    uint64_t *solve_bar = acc_full + WARPS;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(solve_bar + 2);  // solve_bar[1]: second pass of the updates
    int *s_misc = reinterpret_cast<int *>(s_tmem + 4);

    const ET *__restrict__ other = reinterpret_cast<const ET *>(a.d_other);
    constexpr int k = KP;
    constexpr bool IMPLICIT = MODE == LK_ALS_IMPLICIT;
    // A = OtOr + scale * G; the in-TMEM system is A / scale, solved against y / scale
    const float scale = (IMPLICIT && !WEIGHTED) ? a.uniform_val : 1.0f;
    const float rscale = 1.0f / scale;

    if (tid == 0) {
        for (int i = 0; i < WARPS * NSTAGE + WARPS; i++) mbar_init(&bars[i], 1);
        mbar_init(solve_bar, 4);
        mbar_init(solve_bar + 1, 4);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem_base = *s_tmem;
    // accumulator of warp w: columns 64*(w/2), lanes 16*(w%2) of every 32-lane quarter
    const uint32_t my_acc = tmem_base + ((uint32_t)((warp & 1) * 16) << 16) + (uint32_t)((warp >> 1) * 64);

    uint32_t free_par = 0, full_par = 0, solve_par = 0;

    // implicit mode: accumulators start every group holding OtOr / scale (warp w: rows 16w..16w+15 of all systems)
    auto preload_otor = [&]() {
        const float4 *ot = reinterpret_cast<const float4 *>(a.d_otor + (16 * warp + (lane & 15)) * k);
        uint32_t r[64];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const float4 o = __ldg(ot + q);
            r[4 * q + 0] = __float_as_uint(o.x * rscale), r[4 * q + 1] = __float_as_uint(o.y * rscale);
            r[4 * q + 2] = __float_as_uint(o.z * rscale), r[4 * q + 3] = __float_as_uint(o.w * rscale);
        }
        const uint32_t t = tmem_base + ((uint32_t)(32 * warp) << 16);
        ctc::tmem_st64(t, r);
        ctc::tmem_st64(t + 64u, r);
        tmem_fence_before();
    };
    if constexpr (IMPLICIT) preload_otor();

    // lane = (row quad, feature quad) of the 8 x 64 stage
    const int rq = lane >> 4, fq = lane & 15;
    // byte offset of this lane's K-vector of feature 4*fq + j inside a tile: group (4fq+j)/8, row 4*rq.., core row (4fq+j)%8
    const uint32_t st_off = (uint32_t)((fq >> 1) * GROUP_STRIDE + rq * 128 + (fq & 1) * 64);

    if (tid == 0) s_misc[0] = fetch_work(a.d_work_counter, a.d_cancel);
    __syncthreads();
    for (;;) {
        const int64_t g = __shfl_sync(FULL, s_misc[0], 0);
        __syncthreads();
        if (g * WARPS >= a.n_chunks) break;
        const int64_t ci = g * WARPS + warp;
        const bool active = ci < a.n_chunks;
        int row = -1, begin = 0, len = 0, nparts = 1, slot0 = 0, part = 0, split_idx = 0;
        if (active) {
            const int4 c0 = __ldg(reinterpret_cast<const int4 *>(a.d_chunks) + 2 * ci);
            const int4 c1 = __ldg(reinterpret_cast<const int4 *>(a.d_chunks) + 2 * ci + 1);
            row = c0.x, begin = c0.y, len = c0.z, nparts = c0.w;
            slot0 = c1.x, part = c1.y, split_idx = c1.z;
        }
        row = __shfl_sync(FULL, row, 0), begin = __shfl_sync(FULL, begin, 0), len = __shfl_sync(FULL, len, 0);
        nparts = __shfl_sync(FULL, nparts, 0), slot0 = __shfl_sync(FULL, slot0, 0);
        part = __shfl_sync(FULL, part, 0), split_idx = __shfl_sync(FULL, split_idx, 0);
        const bool has_gram = active && len > 0;
        int n_row = 0;
        if (active) n_row = __ldg(a.d_indptr + row + 1) - __ldg(a.d_indptr + row);
        if (lane == 0) {
            int *m = s_misc + 8 + 8 * warp;
            m[0] = has_gram ? 1 : 0;
            m[1] = nparts;
            m[2] = slot0 + part;
            m[3] = n_row;
            m[4] = slot0;
            m[5] = row;
        }

        // ------------------------------------------------------------------
        // phase 1: rows -> registers -> (y, hi / lo K-major tiles) -> tcgen05.mma
        // ------------------------------------------------------------------
        if (has_gram) {
            const int n_it = (len + STAGE_ROWS - 1) / STAGE_ROWS;
            const int32_t *cols = a.d_cols + begin;
            const float *vals = a.d_vals + begin;
            float4 ysum = make_float4(0.f, 0.f, 0.f, 0.f);  // features 4*fq..4*fq+3 over this lane's rows

            // Two-level software pipeline.  Stage `it` = rows it*8 + 4*rq + i (i = 0..3) for this lane.
            //   column indices / values: a ring of IDX_AHEAD stages, loaded that far ahead, so that
            //   the row loads: issued two stages ahead, never wait for the DRAM latency of their own index
            //   (with both in one step every stage paid that latency in line — the phase was bound by it).
            constexpr int IDX_AHEAD = 4;
            int ci[IDX_AHEAD][4];
            float vi[IDX_AHEAD][4];
            auto load_idx = [&](int it, int (&c)[4], float (&v)[4]) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int idx = it * STAGE_ROWS + 4 * rq + i;
                    const bool in = it < n_it && idx < len;
                    c[i] = in ? __ldg(cols + idx) : -1;
                    v[i] = in ? __ldg(vals + idx) : 0.0f;
                }
            };
            auto load_rows = [&](const int (&c)[4], float4 (&x)[4]) {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    x[i] = c[i] >= 0 ? load_quad(other, c[i], fq) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            float4 xr[2][4];
#pragma unroll
            for (int u = 0; u < IDX_AHEAD; u++) load_idx(u, ci[u], vi[u]);
            load_rows(ci[0], xr[0]);
            load_rows(ci[1], xr[1]);

            auto consume = [&](int it, float4 (&x)[4], float (&v)[4]) {
                const int s = it % NSTAGE;
                if (it >= NSTAGE) {  // the MMAs of stage it - NSTAGE were the last commit on this buffer
                    mbar_wait(&stage_free[s], ((free_par >> s) & 1u) ^ 1u);
                }
                unsigned char *thi = ring + s * STAGE_BYTES + st_off;
                unsigned char *tlo = thi + TILE_BYTES;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    // right-hand side: implicit y += (v + 1) o, explicit y += v o (rows past the end carry x = 0)
                    const float w = IMPLICIT ? v[i] + 1.0f : v[i];
                    ysum.x = fmaf(w, x[i].x, ysum.x), ysum.y = fmaf(w, x[i].y, ysum.y);
                    ysum.z = fmaf(w, x[i].z, ysum.z), ysum.w = fmaf(w, x[i].w, ysum.w);
                    if constexpr (WEIGHTED) {
                        const float sv = sqrtf(v[i]);
                        x[i].x *= sv, x[i].y *= sv, x[i].z *= sv, x[i].w *= sv;
                    }
                }
                // transposed stores: feature 4*fq + j, rows 4*rq .. 4*rq+3 -> one 16-byte K-vector
                auto put = [&](int j, float e0, float e1, float e2, float e3) {
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(e0) & 0xffffe000u);
                    hi.y = __uint_as_float(__float_as_uint(e1) & 0xffffe000u);
                    hi.z = __uint_as_float(__float_as_uint(e2) & 0xffffe000u);
                    hi.w = __uint_as_float(__float_as_uint(e3) & 0xffffe000u);
                    lo.x = e0 - hi.x, lo.y = e1 - hi.y, lo.z = e2 - hi.z, lo.w = e3 - hi.w;
                    *reinterpret_cast<float4 *>(thi + j * 16) = hi;
                    *reinterpret_cast<float4 *>(tlo + j * 16) = lo;
                };
                put(0, x[0].x, x[1].x, x[2].x, x[3].x);
                put(1, x[0].y, x[1].y, x[2].y, x[3].y);
                put(2, x[0].z, x[1].z, x[2].z, x[3].z);
                put(3, x[0].w, x[1].w, x[2].w, x[3].w);
                fence_proxy_async();
                __syncwarp();
                if (ctc::elect_one()) {
                    tmem_fence_after();
                    const uint32_t hi_a = smem_u32(ring + s * STAGE_BYTES);
                    const uint64_t dh = DESC | (uint64_t)((hi_a >> 4) & 0x3fffu);
                    const uint64_t dl = DESC | (uint64_t)(((hi_a + TILE_BYTES) >> 4) & 0x3fffu);
                    umma_tf32_acc(my_acc, dh, dh, IDESC, (IMPLICIT || it > 0) ? 1u : 0u);
                    umma_tf32_acc(my_acc, dh, dl, IDESC, 1u);
                    umma_tf32_acc(my_acc, dl, dh, IDESC, 1u);
                    umma_commit(&stage_free[s]);
                    if (it == n_it - 1) umma_commit(&acc_full[warp]);
                }
                free_par ^= (1u << s);
                __syncwarp();
            };
            for (int it0 = 0; it0 < n_it; it0 += IDX_AHEAD) {
#pragma unroll
                for (int u = 0; u < IDX_AHEAD; u++) {
                    const int it = it0 + u;
                    if (it < n_it) {
                        consume(it, xr[u & 1], vi[u]);
                        load_rows(ci[(u + 2) % IDX_AHEAD], xr[u & 1]);  // rows of stage it + 2
                        load_idx(it + IDX_AHEAD, ci[u], vi[u]);         // indices of stage it + IDX_AHEAD
                    }
                }
            }
            // right-hand side of the chunk: fold the two row quads, features 4*fq..4*fq+3 -> lanes 0..15
            ysum.x += __shfl_xor_sync(FULL, ysum.x, 16), ysum.y += __shfl_xor_sync(FULL, ysum.y, 16);
            ysum.z += __shfl_xor_sync(FULL, ysum.z, 16), ysum.w += __shfl_xor_sync(FULL, ysum.w, 16);
            if (lane < 16) {
                if (nparts == 1)
                    *reinterpret_cast<float4 *>(ys_all + warp * KP + 4 * fq) = ysum;
                else
                    __stcg(reinterpret_cast<float4 *>(a.d_partials + (size_t)(slot0 + part) * SLOTF + KP * KP) + fq, ysum);
            }
        }
        __syncthreads();  // every warp has issued its MMAs; chunk metadata and right-hand sides are visible

        // ------------------------------------------------------------------
        // phase 2: finish the systems in place in TMEM
        // ------------------------------------------------------------------
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
        const ctc::Workspace ws = ctc::carve(base);  // aliases the rings: all their MMAs have completed
        const int r16 = lane & 15, hh = lane >> 4;
        const int gi = 16 * warp + r16;  // Gram row / feature held by this lane (of system 2p + hh)
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
        const bool anysplit = parts[0] > 1 || parts[1] > 1 || parts[2] > 1 || parts[3] > 1;
        float *dsum = reinterpret_cast<float *>(s_misc + 44);  // [warp][system] partial |delta|^2
        float yv[2] = {0.0f, 0.0f};
        // right-hand sides of the unsplit systems (read before the workspace is written: ys_all is outside the union)
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const int c = 2 * p + hh;
            if (gram[c] && parts[c] == 1) yv[p] = ys_all[c * KP + gi] * rscale;
        }
        if (tid < 4) ws.bad[tid] = 0;
#pragma unroll
        for (int p = 0; p < 2; p++) {
            if (!(gram[2 * p] || gram[2 * p + 1])) continue;
            if constexpr (IMPLICIT) {
                // preloaded with OtOr / scale: unsplit rows are complete as they stand
                if (!((gram[2 * p] && parts[2 * p] > 1) || (gram[2 * p + 1] && parts[2 * p + 1] > 1))) continue;
            }
            const int c = 2 * p + hh;
            uint32_t r[64];
            tmem_ld_32x32b_x64(lane_taddr + (uint32_t)(p * 64), r);
            if (gram[c]) {
                if (parts[c] == 1) {
                    if constexpr (!IMPLICIT) {
                        const float regn = a.reg * (float)nrowc[c];
#pragma unroll
                        for (int i = 0; i < 64; i++)
                            if (i == gi) r[i] = __float_as_uint(__uint_as_float(r[i]) + regn);
                    }
                } else {
                    float *slot = a.d_partials + (size_t)slotc[c] * SLOTF + gi * KP;
#pragma unroll
                    for (int q = 0; q < 16; q++)
                        __stcg(reinterpret_cast<float4 *>(slot) + q,
                               make_float4(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1]),
                                           __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
                }
            }
            if (!IMPLICIT && ((gram[2 * p] && parts[2 * p] == 1) || (gram[2 * p + 1] && parts[2 * p + 1] == 1)))
                ctc::tmem_st64(lane_taddr + (uint32_t)(p * 64), r);
        }
        tmem_fence_before();
        if (anysplit) __threadfence();  // partial slots only
        __syncthreads();
        tmem_fence_after();
        uint32_t solve_mask = 0;
#pragma unroll
        for (int c = 0; c < WARPS; c++)
            if (gram[c] && parts[c] == 1) solve_mask |= 1u << c;
        // split rows: the last part to arrive sums the slots in order, straight into TMEM
        if (anysplit) {
            if (lane == 0)
                s_misc[40 + warp] =
                    (active && nparts > 1 && atomicAdd(a.d_split_counters + split_idx, 1) == nparts - 1) ? 1 : 0;
            __syncthreads();
#pragma unroll
            for (int c = 0; c < WARPS; c++) {
                if (!(parts[c] > 1 && s_misc[40 + c])) continue;
                __threadfence();
                const int p = c >> 1;
                const int slot0c = s_misc[8 + 8 * c + 4];
                uint32_t r[64];
                tmem_ld_32x32b_x64(lane_taddr + (uint32_t)(p * 64), r);
                if (hh == (c & 1)) {
                    const float regn = a.reg * (float)nrowc[c];
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int pp = 0; pp < parts[c]; pp++) {
                            const float4 t = __ldcg(
                                reinterpret_cast<const float4 *>(a.d_partials + (size_t)(slot0c + pp) * SLOTF + gi * KP) + q);
                            sacc.x += t.x, sacc.y += t.y, sacc.z += t.z, sacc.w += t.w;
                        }
                        if constexpr (IMPLICIT) {
                            // every part carries one copy of the preloaded OtOr / scale: keep exactly one
                            const float4 o = __ldg(reinterpret_cast<const float4 *>(a.d_otor + gi * k) + q);
                            const float extra = (float)(parts[c] - 1) * rscale;
                            sacc.x = fmaf(-extra, o.x, sacc.x), sacc.y = fmaf(-extra, o.y, sacc.y);
                            sacc.z = fmaf(-extra, o.z, sacc.z), sacc.w = fmaf(-extra, o.w, sacc.w);
                        } else {
                            if (4 * q + 0 == gi) sacc.x += regn;
                            if (4 * q + 1 == gi) sacc.y += regn;
                            if (4 * q + 2 == gi) sacc.z += regn;
                            if (4 * q + 3 == gi) sacc.w += regn;
                        }
                        r[4 * q + 0] = __float_as_uint(sacc.x), r[4 * q + 1] = __float_as_uint(sacc.y);
                        r[4 * q + 2] = __float_as_uint(sacc.z), r[4 * q + 3] = __float_as_uint(sacc.w);
                    }
                    float sy = 0.0f;
                    for (int pp = 0; pp < parts[c]; pp++)
                        sy += __ldcg(a.d_partials + (size_t)(slot0c + pp) * SLOTF + KP * KP + gi);
                    yv[p] = sy * rscale;
                }
                ctc::tmem_st64(lane_taddr + (uint32_t)(p * 64), r);
                solve_mask |= 1u << c;
            }
        }

        // ------------------------------------------------------------------
        // phase 3: blocked Cholesky on the tensor cores (chol_tc.cuh), write-back
        // ------------------------------------------------------------------
        if (tid == 0) s_misc[0] = fetch_work(a.d_work_counter, a.d_cancel);  // next group, read after the closing barrier
        if (solve_mask) {
            // old values of the rows about to be written: fetched before the solve so that the
            // write-back does not wait for them
            float xold[2] = {0.0f, 0.0f};
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const int c = 2 * p + hh;
                if ((solve_mask >> c) & 1u) xold[p] = a.d_this[(size_t)s_misc[8 + 8 * c + 5] * k + gi];
            }
            ctc::solve4<false, tcx::TCX_GJ>(tmem_base, yv, ws, solve_bar, solve_par, tid);
            __syncthreads();  // pivot flags
            float dpart[2] = {0.0f, 0.0f};
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const int c = 2 * p + hh;
                if ((solve_mask >> c) & 1u) {
                    const int rowc = s_misc[8 + 8 * c + 5];
                    if (ws.bad[c]) {
                        if (warp == 0 && r16 == 0) atomicCAS(a.d_status, 0, rowc + 1);
                    } else {
                        const float xn = yv[p];
                        const float d = xn - xold[p];
                        dpart[p] = d * d;
                        a.d_this[(size_t)rowc * k + gi] = xn;
                        for (int rr = 0; rr < a.n_replicas; rr++)
                            a.d_replicas[rr][(size_t)(a.replica_row0 + rowc) * k + gi] = xn;
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < 2; p++) {
                float v = dpart[p];
#pragma unroll
                for (int w = 8; w >= 1; w >>= 1) v += __shfl_xor_sync(FULL, v, w, 16);
                if (r16 == 0) dsum[warp * 4 + 2 * p + hh] = v;
            }
        }
        if (active && len == 0 && nparts == 1) {
            // empty row: x = 0, no delta (implicit.rs:98-101)
            float *thisrow = a.d_this + (size_t)row * k;
            for (int i = lane; i < k; i += 32) {
                thisrow[i] = 0.0f;
                for (int rr = 0; rr < a.n_replicas; rr++)
                    a.d_replicas[rr][(size_t)(a.replica_row0 + row) * k + i] = 0.0f;
            }
        }
        if constexpr (IMPLICIT) preload_otor();  // accumulators of the next group (this warp's rows)
        __syncthreads();  // the workspace aliases the rings of the next group
        if (tid < 4 && ((solve_mask >> tid) & 1u)) {
            const float t = ((dsum[tid] + dsum[4 + tid]) + dsum[8 + tid]) + dsum[12 + tid];
            if (t != 0.0f) atomicAdd(a.d_sqdelta, (double)t);
        }
    }

    tmem_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)tcx::TMEM_COLS)
                     : "memory");
    }
}

// returns LK_OK when this kernel took the launch, 1 when the caller should fall back to the SIMT kernel
int launch_als_tcx(const lk_als_args &a, cudaStream_t st)
{
    const Options &opt = options();
    if (a.k != tcx::KP || opt.als_tf32 == 0 || opt.als_tcs == 0 || opt.als_tc_interleave == 0) return 1;
    const bool f32 = a.other_dtype == LK_DTYPE_F32;
    if (reinterpret_cast<uintptr_t>(a.d_other) % 16 != 0) return 1;
    const bool implicit = a.mode == LK_ALS_IMPLICIT;
    const bool weighted = implicit && !a.vals_uniform;
    // the weighted Gram uses z = sqrt(v) o: confidences must not be negative (the plan passes the minimum)
    if (weighted && !(a.uniform_val >= 0.0f)) return 1;
    // uniform confidence 0: (A / v) is undefined
    if (implicit && !weighted && !(fabsf(a.uniform_val) > 1e-20f)) return 1;
    const int smem = tcx::SMEM_BYTES;
    int occ = 3;
    if (opt.als_tc_occ > 0) occ = std::max(1, std::min(3, opt.als_tc_occ));
    const int64_t groups = (a.n_chunks + tcx::WARPS - 1) / tcx::WARPS;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)sm_count() * occ, groups));
    auto launch = [&](auto kern) -> int {
        LK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<(unsigned)grid, tcx::NT, smem, st>>>(a);
        return LK_OK;
    };
    int rc;
    if (implicit) {
        if (weighted)
            rc = f32 ? launch(als_tcx_kernel<LK_ALS_IMPLICIT, float, true>)
                     : launch(als_tcx_kernel<LK_ALS_IMPLICIT, __nv_bfloat16, true>);
        else
            rc = f32 ? launch(als_tcx_kernel<LK_ALS_IMPLICIT, float, false>)
                     : launch(als_tcx_kernel<LK_ALS_IMPLICIT, __nv_bfloat16, false>);
    } else {
        rc = f32 ? launch(als_tcx_kernel<LK_ALS_EXPLICIT, float, false>)
                 : launch(als_tcx_kernel<LK_ALS_EXPLICIT, __nv_bfloat16, false>);
    }
    if (rc != LK_OK) return rc;
    LK_CUDA_TRY(cudaGetLastError());
    return LK_OK;
}

}  // namespace lk
