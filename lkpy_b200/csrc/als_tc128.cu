// als_tc128.cu — ALS half-epoch at k = 128 on the tensor cores (BASELINE configs[3]: 100 M interactions,
// features = 128), the 128x128 systems solved in TMEM (chol_tc128.cuh).  Same contract as
// als_half_kernel<128, ...> (reference src/accel/als/implicit.rs:87-125, explicit.rs:80-119,
// solve.rs:65-106), which keeps what this kernel does not take (negative or zero confidences).
// Two gather paths, one per operand type:
//   * bf16 rows, unweighted (explicit) or uniformly weighted (implicit without ratings): exact bf16 products,
//     tcgen05.mma.kind::f16 — described below;
//   * fp32 rows (the reference's own arithmetic) and/or per-nonzero confidences (use_ratings=True): the
//     tf32 hi/lo split of als_tcx.cu at twice the width — see `gather_tf32`.
//
// A 128x128 lower triangle is three M=64 accumulators (chol_tc128.cuh): with the gathered row split into
// its two 64-feature halves m = [m0 | m1] (each exactly one 128-byte row of an MN-major SWIZZLE_128B tile),
//      G11 += M0^T M0      G21 += M1^T M0      G22 += M1^T M1
// are three tcgen05.mma.kind::f16 (M = N = 64, K = 16) per 16 gathered rows, fp32 accumulation in TMEM,
// exact products (the operands are bf16).  256 TMEM columns hold two systems; two CTAs share an SM.
//
// CTA = 4 warps working on groups of 2 row chunks.  Warps 0 and 1 gather one chunk each: a lane reads
// 4 features (8 bytes) of a row — 32 lanes = one 256-byte row per load instruction, 16 rows per stage —
// keeps the running column sums of its features in registers (the right-hand side is (v + 1) x those
// sums, or sum v_j m_j in explicit mode: no tensor-core pass, no extra TMEM columns) and stores the 8
// bytes into the swizzled tiles; column indices run a ring of stages ahead of the row loads.  Then all
// four warps finish the systems in place (OtOr / v preloaded, reg * n on the diagonal, split rows reduced
// in slot order) and solve them together.

#include <algorithm>

#include "chol_tc128.cuh"

namespace lk {

namespace t128 {
constexpr int KP = 128;
constexpr int WARPS = 4;
constexpr int NT = WARPS * 32;
constexpr int NSYS = 2;
constexpr int STAGE_ROWS = 16;
constexpr int NSTAGE = 3;
constexpr int ATOM_BYTES = STAGE_ROWS * 128;    // 16 rows x 64 bf16
constexpr int STAGE_BYTES = 2 * ATOM_BYTES;     // features 0..63 | 64..127
constexpr int RING_BYTES = NSTAGE * STAGE_BYTES;  // per gathering warp: 12 KB
// tf32 path: 8 rows per stage; K-major no-swizzle tiles of 64 features x 8 rows (als_tcx.cu's layout: the
// 8-feature groups 272 bytes apart), four per stage: features 0..63 hi, lo, features 64..127 hi, lo
constexpr int X_STAGE_ROWS = 8;
constexpr int X_GROUP_STRIDE = 272;
constexpr int X_TILE_BYTES = 8 * X_GROUP_STRIDE;
constexpr int X_STAGE_BYTES = 4 * X_TILE_BYTES;
constexpr int X_RING_BYTES = NSTAGE * X_STAGE_BYTES;  // per gathering warp: 25.5 KB
constexpr uint32_t X_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
constexpr uint64_t X_DESC = (uint64_t(128 >> 4) << 16) | (uint64_t(X_GROUP_STRIDE >> 4) << 32) | (1ull << 46);
constexpr int TMEM_COLS = 256;
constexpr int SLOTF = KP * KP + KP;
constexpr int WS_ALIGNED = (ctc128::WS_BYTES + 1023) & ~1023;
static_assert(X_RING_BYTES >= RING_BYTES && X_RING_BYTES % 16 == 0, "the union is sized for the larger ring");
constexpr int UNION_BYTES = WS_ALIGNED > NSYS * X_RING_BYTES ? WS_ALIGNED : NSYS * X_RING_BYTES;
constexpr int SMEM_BYTES = 1024 /*alignment slack*/ + UNION_BYTES + NSYS * KP * 4 + (NSYS * NSTAGE + NSYS + 2) * 8 + 16 + 64 * 4;
}  // namespace t128

// 4 consecutive features (quad q of 32) of a gathered 128-feature row as f32
__device__ __forceinline__ float4 load_quad128(const float *other, int row, int q)
{
    return __ldg(reinterpret_cast<const float4 *>(other + (size_t)row * t128::KP) + q);
}
__device__ __forceinline__ float4 load_quad128(const __nv_bfloat16 *other, int row, int q)
{
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(other + (size_t)row * t128::KP) + q);
    return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u), __uint_as_float(t.y << 16),
                       __uint_as_float(t.y & 0xffff0000u));
}

// MODE: LK_ALS_IMPLICIT / LK_ALS_EXPLICIT.  ET: float or __nv_bfloat16 rows of `other`.  WEIGHTED: implicit
// mode with per-nonzero confidences (z = sqrt(v) o, A = OtOr + Z^T Z); otherwise the Gram is unweighted and
// scaled by the uniform confidence (implicit) or not at all (explicit).
template <int MODE, typename ET, bool WEIGHTED>
__global__ void __launch_bounds__(t128::NT, 2) als_tc128_kernel(lk_als_args a)
{
    using namespace t128;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(FULL, tid >> 5, 0);

    float *ys_all = reinterpret_cast<float *>(base + UNION_BYTES);  // [NSYS][KP]
    uint64_t *bars = reinterpret_cast<uint64_t *>(ys_all + NSYS * KP);
    uint64_t *acc_full = bars + NSYS * NSTAGE;
    uint64_t *solve_bar = acc_full + NSYS;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(solve_bar + 2);  // solve_bar[1]: second pass of the updates
    int *s_misc = reinterpret_cast<int *>(s_tmem + 4);

    const ET *__restrict__ other = reinterpret_cast<const ET *>(a.d_other);
    constexpr int k = KP;
    constexpr bool IMPLICIT = MODE == LK_ALS_IMPLICIT;
    constexpr bool TF32 = WEIGHTED || sizeof(ET) == 4;  // bf16 rows with uniform weights take the kind::f16 path
    // A = OtOr + scale * G; the in-TMEM system is A / scale, solved against y / scale
    const float scale = (IMPLICIT && !WEIGHTED) ? a.uniform_val : 1.0f;
    const float rscale = 1.0f / scale;

    if (tid == 0) {
        for (int i = 0; i < NSYS * NSTAGE + NSYS; i++) mbar_init(&bars[i], 1);
        mbar_init(solve_bar, ctc128::NSYS);
        mbar_init(solve_bar + 1, ctc128::NSYS);
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
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
    const int r16 = lane & 15, hh = lane >> 4;
    const int R = 64 * hh + 16 * warp + r16;  // the row of both systems this lane owns (chol_tc128.cuh)

    uint32_t free_par = 0, full_par = 0, solve_par = 0;

    // implicit mode: every group starts with OtOr / scale in the accumulators (row R: columns 0..63 into slot P0,
    // and for the bottom rows columns 64..127 into slot P1; the top half of P1 is never read)
    auto preload_otor = [&]() {
        uint32_t r[64];
#pragma unroll
        for (int part = 0; part < 2; part++) {
            const float4 *ot = reinterpret_cast<const float4 *>(a.d_otor + (size_t)R * k + 64 * part);
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const float4 o = __ldg(ot + q);
                r[4 * q + 0] = __float_as_uint(o.x * rscale), r[4 * q + 1] = __float_as_uint(o.y * rscale);
                r[4 * q + 2] = __float_as_uint(o.z * rscale), r[4 * q + 3] = __float_as_uint(o.w * rscale);
            }
#pragma unroll
            for (int p = 0; p < NSYS; p++) ctc::tmem_st64(lane_taddr + (uint32_t)(128 * p + 64 * part), r);
        }
        tmem_fence_before();
    };
    if constexpr (IMPLICIT) preload_otor();

    if (tid == 0) s_misc[0] = fetch_work(a.d_work_counter, a.d_cancel);
    __syncthreads();
    for (;;) {
        const int64_t g = __shfl_sync(FULL, s_misc[0], 0);
        __syncthreads();
        if (g * NSYS >= a.n_chunks) break;
        // chunk of system s = warp (warps 0, 1 gather; every warp reads the metadata of both afterwards)
        const int64_t ci = g * NSYS + warp;
        const bool active = warp < NSYS && ci < a.n_chunks;
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
        if (lane == 0 && warp < NSYS) {
            int *m = s_misc + 8 + 8 * warp;
            m[0] = has_gram ? 1 : 0;
            m[1] = nparts;
            m[2] = slot0 + part;
            m[3] = n_row;
            m[4] = slot0;
            m[5] = active ? row : -1;
            m[6] = (active && len == 0 && nparts == 1) ? 1 : 0;  // empty row
        }

        // ------------------------------------------------------------------
        // phase 1 (warps 0, 1): rows -> registers -> (column sums, swizzled bf16 tiles) -> tcgen05.mma
        // ------------------------------------------------------------------
        if constexpr (TF32) {
          if (has_gram) {
            // gather_tf32: a lane is (row quad rq, feature quad fq) of the 8 x 128 stage and holds a 4 x 4
            // micro-block of each 64-feature half: 8 LDG.128 per stage (4 x LDG.64 for bf16 rows).  The rows are
            // scaled (sqrt(v) when WEIGHTED), accumulated into the right-hand side on the way, split exactly into
            // hi (top 19 bits, a tf32 number) and lo = z - hi, and stored transposed as 16-byte K-vectors.  Per
            // 8 rows: G11 += Z0^T Z0, G21 += Z1^T Z0, G22 += Z1^T Z1, each as hi.hi + hi.lo + lo.hi — nine
            // tcgen05.mma.kind::tf32 (M = N = 64, K = 8).
            unsigned char *ring = base + warp * X_RING_BYTES;
            uint64_t *stage_free = bars + warp * NSTAGE;
            const int n_it = (len + X_STAGE_ROWS - 1) / X_STAGE_ROWS;
            const int32_t *cols = a.d_cols + begin;
            const float *vals = a.d_vals + begin;
            const uint32_t top_d = tmem_base + (uint32_t)(128 * warp);
            const uint32_t bot_d = tmem_base + ((uint32_t)16 << 16) + (uint32_t)(128 * warp);
            const int rq = lane >> 4, fq = lane & 15;
            const uint32_t st_off = (uint32_t)((fq >> 1) * X_GROUP_STRIDE + rq * 128 + (fq & 1) * 64);
            float4 ysum[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};

            constexpr int IDX_AHEAD = 4;
            int ci_[IDX_AHEAD][4];
            float vi_[IDX_AHEAD][4];
            auto load_idx = [&](int it, int (&c)[4], float (&v)[4]) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int idx = it * X_STAGE_ROWS + 4 * rq + i;
                    const bool in = it < n_it && idx < len;
                    c[i] = in ? __ldg(cols + idx) : -1;
                    v[i] = in ? __ldg(vals + idx) : 0.0f;
                }
            };
            auto load_rows = [&](const int (&c)[4], float4 (&x)[2][4]) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
#pragma unroll
                    for (int h2 = 0; h2 < 2; h2++)
                        x[h2][i] = c[i] >= 0 ? load_quad128(other, c[i], fq + 16 * h2) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            float4 xr[2][2][4];
#pragma unroll
            for (int u = 0; u < IDX_AHEAD; u++) load_idx(u, ci_[u], vi_[u]);
            load_rows(ci_[0], xr[0]);
            load_rows(ci_[1], xr[1]);

            auto consume = [&](int it, float4 (&x)[2][4], float (&v)[4]) {
                const int s = it % NSTAGE;
                if (it >= NSTAGE) mbar_wait(&stage_free[s], ((free_par >> s) & 1u) ^ 1u);
                unsigned char *st = ring + s * X_STAGE_BYTES + st_off;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    // right-hand side: implicit y += (v + 1) o, explicit y += v o (rows past the end carry x = 0)
                    const float w = IMPLICIT ? v[i] + 1.0f : v[i];
                    float sv = 1.0f;
                    if constexpr (WEIGHTED) sv = sqrtf(v[i]);
#pragma unroll
                    for (int h2 = 0; h2 < 2; h2++) {
                        ysum[h2].x = fmaf(w, x[h2][i].x, ysum[h2].x), ysum[h2].y = fmaf(w, x[h2][i].y, ysum[h2].y);
                        ysum[h2].z = fmaf(w, x[h2][i].z, ysum[h2].z), ysum[h2].w = fmaf(w, x[h2][i].w, ysum[h2].w);
                        if constexpr (WEIGHTED)
                            x[h2][i].x *= sv, x[h2][i].y *= sv, x[h2][i].z *= sv, x[h2][i].w *= sv;
                    }
                }
#pragma unroll
                for (int h2 = 0; h2 < 2; h2++) {
                    unsigned char *thi = st + (2 * h2) * X_TILE_BYTES;
                    unsigned char *tlo = thi + X_TILE_BYTES;
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
                    put(0, x[h2][0].x, x[h2][1].x, x[h2][2].x, x[h2][3].x);
                    put(1, x[h2][0].y, x[h2][1].y, x[h2][2].y, x[h2][3].y);
                    put(2, x[h2][0].z, x[h2][1].z, x[h2][2].z, x[h2][3].z);
                    put(3, x[h2][0].w, x[h2][1].w, x[h2][2].w, x[h2][3].w);
                }
                fence_proxy_async();
                __syncwarp();
                if (ctc::elect_one()) {
                    tmem_fence_after();
                    const uint32_t t0 = smem_u32(ring + s * X_STAGE_BYTES);
                    auto dsc = [&](int tile) { return X_DESC | (uint64_t)(((t0 + tile * X_TILE_BYTES) >> 4) & 0x3fffu); };
                    const uint64_t h0 = dsc(0), l0 = dsc(1), h1 = dsc(2), l1 = dsc(3);
                    const uint32_t acc = (IMPLICIT || it > 0) ? 1u : 0u;
                    umma_tf32_acc(top_d, h0, h0, X_IDESC, acc);  // G11 += Z0^T Z0
                    umma_tf32_acc(top_d, h0, l0, X_IDESC, 1u);
                    umma_tf32_acc(top_d, l0, h0, X_IDESC, 1u);
                    umma_tf32_acc(bot_d, h1, h0, X_IDESC, acc);  // G21 += Z1^T Z0
                    umma_tf32_acc(bot_d, h1, l0, X_IDESC, 1u);
                    umma_tf32_acc(bot_d, l1, h0, X_IDESC, 1u);
                    umma_tf32_acc(bot_d + 64, h1, h1, X_IDESC, acc);  // G22 += Z1^T Z1
                    umma_tf32_acc(bot_d + 64, h1, l1, X_IDESC, 1u);
                    umma_tf32_acc(bot_d + 64, l1, h1, X_IDESC, 1u);
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
                        consume(it, xr[u & 1], vi_[u]);
                        load_rows(ci_[(u + 2) % IDX_AHEAD], xr[u & 1]);  // rows of stage it + 2
                        load_idx(it + IDX_AHEAD, ci_[u], vi_[u]);        // indices of stage it + IDX_AHEAD
                    }
                }
            }
            // right-hand side of the chunk: fold the two row quads; lanes 0..15 hold features 4*fq.. of each half
#pragma unroll
            for (int h2 = 0; h2 < 2; h2++) {
                ysum[h2].x += __shfl_xor_sync(FULL, ysum[h2].x, 16), ysum[h2].y += __shfl_xor_sync(FULL, ysum[h2].y, 16);
                ysum[h2].z += __shfl_xor_sync(FULL, ysum[h2].z, 16), ysum[h2].w += __shfl_xor_sync(FULL, ysum[h2].w, 16);
                if (lane < 16) {
                    if (nparts == 1)
                        *reinterpret_cast<float4 *>(ys_all + warp * KP + 64 * h2 + 4 * fq) = ysum[h2];
                    else
                        __stcg(reinterpret_cast<float4 *>(a.d_partials + (size_t)(slot0 + part) * SLOTF + KP * KP) +
                                   16 * h2 + fq,
                               ysum[h2]);
                }
            }
          }
        } else if (has_gram) {
            const __nv_bfloat16 *other16 = reinterpret_cast<const __nv_bfloat16 *>(a.d_other);
            unsigned char *ring = base + warp * RING_BYTES;
            uint64_t *stage_free = bars + warp * NSTAGE;
            const int n_it = (len + STAGE_ROWS - 1) / STAGE_ROWS;
            const int32_t *cols = a.d_cols + begin;
            const float *vals = a.d_vals + begin;
            const uint32_t top_d = tmem_base + (uint32_t)(128 * warp);
            const uint32_t bot_d = tmem_base + ((uint32_t)16 << 16) + (uint32_t)(128 * warp);
            float4 ysum = make_float4(0.f, 0.f, 0.f, 0.f);  // features 4*lane .. 4*lane+3
            // lane's 8 bytes inside a row: atom lane/16, 16-byte chunk (lane%16)/2 (XOR row&7), half lane&1
            const uint32_t st_atom = (uint32_t)(lane >> 4) * ATOM_BYTES;
            const int st_chunk = (lane & 15) >> 1;
            const uint32_t st_half = (uint32_t)(lane & 1) * 8;

            // column index / value of row it*16 + lane%16 (both halves of the warp hold a copy)
            constexpr int IDX_AHEAD = 4;
            int cq[IDX_AHEAD];
            float vq[IDX_AHEAD];
            auto load_idx = [&](int it, int &c, float &v) {
                const int idx = it * STAGE_ROWS + (lane & 15);
                const bool in = it < n_it && idx < len;
                c = in ? __ldg(cols + idx) : -1;
                v = in ? __ldg(vals + idx) : 0.0f;
            };
            // the 16 rows of a stage: 8 bytes per lane and row
            auto load_rows = [&](int c, uint2 (&x)[STAGE_ROWS]) {
#pragma unroll
                for (int i = 0; i < STAGE_ROWS; i++) {
                    const int ci_ = __shfl_sync(FULL, c, i);
                    x[i] = ci_ >= 0 ? __ldg(reinterpret_cast<const uint2 *>(other16 + (size_t)ci_ * k) + lane) : make_uint2(0u, 0u);
                }
            };
            uint2 xr[2][STAGE_ROWS];
#pragma unroll
            for (int u = 0; u < IDX_AHEAD; u++) load_idx(u, cq[u], vq[u]);
            load_rows(cq[0], xr[0]);
            load_rows(cq[1], xr[1]);

            auto consume = [&](int it, uint2 (&x)[STAGE_ROWS], float vrow) {
                const int s = it % NSTAGE;
                if (it >= NSTAGE) mbar_wait(&stage_free[s], ((free_par >> s) & 1u) ^ 1u);
                unsigned char *tile = ring + s * STAGE_BYTES + st_atom;
#pragma unroll
                for (int i = 0; i < STAGE_ROWS; i++) {
                    // right-hand side: implicit — plain column sums (x (v + 1) at the end); explicit — weighted by v_j
                    float w = 1.0f;
                    if constexpr (!IMPLICIT) w = __shfl_sync(FULL, vrow, i);
                    ysum.x = fmaf(w, __uint_as_float(x[i].x << 16), ysum.x);
                    ysum.y = fmaf(w, __uint_as_float(x[i].x & 0xffff0000u), ysum.y);
                    ysum.z = fmaf(w, __uint_as_float(x[i].y << 16), ysum.z);
                    ysum.w = fmaf(w, __uint_as_float(x[i].y & 0xffff0000u), ysum.w);
                    *reinterpret_cast<uint2 *>(tile + i * 128 + ((st_chunk ^ (i & 7)) << 4) + st_half) = x[i];
                }
                fence_proxy_async();
                __syncwarp();
                if (ctc::elect_one()) {
                    tmem_fence_after();
                    const uint32_t t0 = smem_u32(ring + s * STAGE_BYTES);
                    const uint64_t d0 = tcd::DESC_HI | tcd::DESC_LBO | (uint64_t)((t0 >> 4) & 0x3fffu);
                    const uint64_t d1 = tcd::DESC_HI | tcd::DESC_LBO | (uint64_t)(((t0 + ATOM_BYTES) >> 4) & 0x3fffu);
                    const uint32_t acc = (IMPLICIT || it > 0) ? 1u : 0u;
                    umma_bf16_ab(top_d, d0, d0, tcd::IDESC, acc);       // G11 += M0^T M0
                    umma_bf16_ab(bot_d, d1, d0, tcd::IDESC, acc);       // G21 += M1^T M0
                    umma_bf16_ab(bot_d + 64, d1, d1, tcd::IDESC, acc);  // G22 += M1^T M1
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
                        consume(it, xr[u & 1], vq[u]);
                        load_rows(cq[(u + 2) % IDX_AHEAD], xr[u & 1]);  // rows of stage it + 2
                        load_idx(it + IDX_AHEAD, cq[u], vq[u]);         // indices of stage it + IDX_AHEAD
                    }
                }
            }
            if constexpr (IMPLICIT) {
                const float w1 = a.uniform_val + 1.0f;
                ysum.x *= w1, ysum.y *= w1, ysum.z *= w1, ysum.w *= w1;
            }
            if (nparts == 1)
                *reinterpret_cast<float4 *>(ys_all + warp * KP + 4 * lane) = ysum;
            else
                __stcg(reinterpret_cast<float4 *>(a.d_partials + (size_t)(slot0 + part) * SLOTF + KP * KP) + lane, ysum);
        }
        __syncthreads();  // MMAs issued, chunk metadata and right-hand sides visible

        // ------------------------------------------------------------------
        // phase 2: finish the systems in place in TMEM
        // ------------------------------------------------------------------
        int gram[NSYS], parts[NSYS], slotc[NSYS], nrowc[NSYS];
#pragma unroll
        for (int c = 0; c < NSYS; c++) {
            const int *m = s_misc + 8 + 8 * c;
            gram[c] = m[0], parts[c] = m[1], slotc[c] = m[2], nrowc[c] = m[3];
        }
#pragma unroll
        for (int c = 0; c < NSYS; c++) {
            if (gram[c]) {
                full_par ^= (1u << c);
                mbar_wait(&acc_full[c], ((full_par >> c) & 1u) ^ 1u);
            }
        }
        tmem_fence_after();
        const ctc128::Workspace ws = ctc128::carve(base);  // aliases the rings: all their MMAs have completed
        const bool anysplit = parts[0] > 1 || parts[1] > 1;
        float *dsum = reinterpret_cast<float *>(s_misc + 44);  // [warp][system] partial |delta|^2
        float yv[NSYS] = {0.0f, 0.0f};
#pragma unroll
        for (int p = 0; p < NSYS; p++)
            if (gram[p] && parts[p] == 1) yv[p] = ys_all[p * KP + R] * rscale;
        if (tid < NSYS) ws.bad[tid] = 0;
#pragma unroll
        for (int p = 0; p < NSYS; p++) {
            if (!gram[p]) continue;
            if (IMPLICIT && parts[p] == 1) continue;  // preloaded with OtOr / scale: complete as it stands
#pragma unroll
            for (int part2 = 0; part2 < 2; part2++) {  // slot P0 (columns 0..63 of row R), slot P1 (columns 64..127, bottom rows)
                uint32_t r[64];
                tmem_ld_32x32b_x64(lane_taddr + (uint32_t)(128 * p + 64 * part2), r);
                if (parts[p] == 1) {
                    // explicit: A += reg * n * I — row R's diagonal entry sits in P0 for a top row, in P1 for a bottom row
                    const float regn = a.reg * (float)nrowc[p];
                    if (part2 == hh) {
#pragma unroll
                        for (int i = 0; i < 64; i++)
                            if (i == 16 * warp + r16) r[i] = __float_as_uint(__uint_as_float(r[i]) + regn);
                    }
                    ctc::tmem_st64(lane_taddr + (uint32_t)(128 * p + 64 * part2), r);
                } else if (part2 == 0 || hh == 1) {
                    float *slot = a.d_partials + (size_t)slotc[p] * SLOTF + (size_t)R * KP + 64 * part2;
#pragma unroll
                    for (int q = 0; q < 16; q++)
                        __stcg(reinterpret_cast<float4 *>(slot) + q,
                               make_float4(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1]),
                                           __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
                }
            }
        }
        tmem_fence_before();
        if (anysplit) __threadfence();
        __syncthreads();
        tmem_fence_after();
        uint32_t solve_mask = 0;
#pragma unroll
        for (int c = 0; c < NSYS; c++)
            if (gram[c] && parts[c] == 1) solve_mask |= 1u << c;
        // split rows: the last part to arrive sums the slots in order, straight into TMEM
        if (anysplit) {
            if (lane == 0 && warp < NSYS)
                s_misc[40 + warp] =
                    (active && nparts > 1 && atomicAdd(a.d_split_counters + split_idx, 1) == nparts - 1) ? 1 : 0;
            __syncthreads();
#pragma unroll
            for (int p = 0; p < NSYS; p++) {
                if (!(parts[p] > 1 && s_misc[40 + p])) continue;
                __threadfence();
                const int slot0c = s_misc[8 + 8 * p + 4];
                const float regn = a.reg * (float)nrowc[p];
#pragma unroll
                for (int part2 = 0; part2 < 2; part2++) {
                    uint32_t r[64];
                    if (part2 == 0 || hh == 1) {
#pragma unroll
                        for (int q = 0; q < 16; q++) {
                            float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
                            for (int pp = 0; pp < parts[p]; pp++) {
                                const float4 t = __ldcg(reinterpret_cast<const float4 *>(
                                                            a.d_partials + (size_t)(slot0c + pp) * SLOTF + (size_t)R * KP + 64 * part2) + q);
                                sacc.x += t.x, sacc.y += t.y, sacc.z += t.z, sacc.w += t.w;
                            }
                            if constexpr (IMPLICIT) {
                                // every part carries one copy of the preloaded OtOr / scale: keep exactly one
                                const float4 o = __ldg(reinterpret_cast<const float4 *>(a.d_otor + (size_t)R * k + 64 * part2) + q);
                                const float extra = (float)(parts[p] - 1) * rscale;
                                sacc.x = fmaf(-extra, o.x, sacc.x), sacc.y = fmaf(-extra, o.y, sacc.y);
                                sacc.z = fmaf(-extra, o.z, sacc.z), sacc.w = fmaf(-extra, o.w, sacc.w);
                            } else if (part2 == hh) {
                                const int dcol = 16 * warp + r16;
                                if (4 * q + 0 == dcol) sacc.x += regn;
                                if (4 * q + 1 == dcol) sacc.y += regn;
                                if (4 * q + 2 == dcol) sacc.z += regn;
                                if (4 * q + 3 == dcol) sacc.w += regn;
                            }
                            r[4 * q + 0] = __float_as_uint(sacc.x), r[4 * q + 1] = __float_as_uint(sacc.y);
                            r[4 * q + 2] = __float_as_uint(sacc.z), r[4 * q + 3] = __float_as_uint(sacc.w);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 64; i++) r[i] = 0u;  // the unused top half of slot P1
                    }
                    ctc::tmem_st64(lane_taddr + (uint32_t)(128 * p + 64 * part2), r);
                }
                float sy = 0.0f;
                for (int pp = 0; pp < parts[p]; pp++) sy += __ldcg(a.d_partials + (size_t)(slot0c + pp) * SLOTF + KP * KP + R);
                yv[p] = sy * rscale;
                solve_mask |= 1u << p;
            }
        }

        // ------------------------------------------------------------------
        // phase 3: blocked Cholesky in TMEM (chol_tc128.cuh), write-back
        // ------------------------------------------------------------------
        if (tid == 0) s_misc[0] = fetch_work(a.d_work_counter, a.d_cancel);  // next group, read after the closing barrier
        if (solve_mask) {
            float xold[NSYS] = {0.0f, 0.0f};
#pragma unroll
            for (int p = 0; p < NSYS; p++)
                if ((solve_mask >> p) & 1u) xold[p] = a.d_this[(size_t)s_misc[8 + 8 * p + 5] * k + R];
            ctc128::solve2(tmem_base, yv, ws, solve_bar, solve_par, tid);
            __syncthreads();  // pivot flags
            float dpart[NSYS] = {0.0f, 0.0f};
#pragma unroll
            for (int p = 0; p < NSYS; p++) {
                if ((solve_mask >> p) & 1u) {
                    const int rowc = s_misc[8 + 8 * p + 5];
                    if (ws.bad[p]) {
                        if (tid == 0) atomicCAS(a.d_status, 0, rowc + 1);
                    } else {
                        const float xn = yv[p];
                        const float d = xn - xold[p];
                        dpart[p] = d * d;
                        a.d_this[(size_t)rowc * k + R] = xn;
                        for (int rr = 0; rr < a.n_replicas; rr++)
                            a.d_replicas[rr][(size_t)(a.replica_row0 + rowc) * k + R] = xn;
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < NSYS; p++) {
                const float v = warp_sum(dpart[p]);
                if (lane == 0) dsum[warp * NSYS + p] = v;
            }
        }
#pragma unroll
        for (int p = 0; p < NSYS; p++) {
            if (s_misc[8 + 8 * p + 6]) {  // empty row: x = 0, no delta (implicit.rs:98-101)
                const int rowc = s_misc[8 + 8 * p + 5];
                a.d_this[(size_t)rowc * k + R] = 0.0f;
                for (int rr = 0; rr < a.n_replicas; rr++)
                    a.d_replicas[rr][(size_t)(a.replica_row0 + rowc) * k + R] = 0.0f;
            }
        }
        if constexpr (IMPLICIT) preload_otor();  // accumulators of the next group
        __syncthreads();  // the workspace aliases the rings of the next group
        if (tid < NSYS && ((solve_mask >> tid) & 1u)) {
            const float t = ((dsum[tid] + dsum[NSYS + tid]) + dsum[2 * NSYS + tid]) + dsum[3 * NSYS + tid];
            if (t != 0.0f) atomicAdd(a.d_sqdelta, (double)t);
        }
    }

    tmem_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)t128::TMEM_COLS)
                     : "memory");
    }
}

// returns LK_OK when this kernel took the launch, 1 when the caller should fall back to the SIMT kernel
int launch_als_tc128(const lk_als_args &a, cudaStream_t st)
{
    const Options &opt = options();
    if (a.k != t128::KP || opt.als_tcs == 0) return 1;
    if (reinterpret_cast<uintptr_t>(a.d_other) % 16 != 0) return 1;
    const bool f32 = a.other_dtype == LK_DTYPE_F32;
    const bool implicit = a.mode == LK_ALS_IMPLICIT;
    const bool weighted = implicit && !a.vals_uniform;
    if ((f32 || weighted) && opt.als_tf32 == 0) return 1;  // diagnostics: keep these on the SIMT kernel
    // the weighted Gram uses z = sqrt(v) o: confidences must not be negative (the plan passes the minimum)
    if (weighted && !(a.uniform_val >= 0.0f)) return 1;
    // uniform confidence 0: (A / v) is undefined
    if (implicit && !weighted && !(fabsf(a.uniform_val) > 1e-20f)) return 1;
    const int smem = t128::SMEM_BYTES;
    const int64_t groups = (a.n_chunks + t128::NSYS - 1) / t128::NSYS;
    int occ = 2;  // 256 TMEM columns per CTA
    if (opt.als_tc_occ > 0) occ = std::max(1, std::min(2, opt.als_tc_occ));
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)sm_count() * occ, groups));
    auto launch = [&](auto kern) -> int {
        LK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<(unsigned)grid, t128::NT, smem, st>>>(a);
        return LK_OK;
    };
    int rc;
    if (implicit) {
        if (weighted)
            rc = f32 ? launch(als_tc128_kernel<LK_ALS_IMPLICIT, float, true>)
                     : launch(als_tc128_kernel<LK_ALS_IMPLICIT, __nv_bfloat16, true>);
        else
            rc = f32 ? launch(als_tc128_kernel<LK_ALS_IMPLICIT, float, false>)
                     : launch(als_tc128_kernel<LK_ALS_IMPLICIT, __nv_bfloat16, false>);
    } else {
        rc = f32 ? launch(als_tc128_kernel<LK_ALS_EXPLICIT, float, false>)
                 : launch(als_tc128_kernel<LK_ALS_EXPLICIT, __nv_bfloat16, false>);
    }
    if (rc != LK_OK) return rc;
    LK_CUDA_TRY(cudaGetLastError());
    return LK_OK;
}

}  // namespace lk
