// prep.cu — item-kNN input preparation on the device.
//
// Replaces the host stage of ItemKNNScorer.train, src/lenskit/knn/item.py:202-228
// (_center_ratings, _normalize_rows: SciPy / NumPy), which on ML-25M-shaped data takes 45x
// longer than the similarity build it feeds.  The f32 values it produces are the *inputs* of a
// bit-exact kernel, so this kernel reproduces the host arithmetic to the bit:
//
//   sums[i]   = np.add.reduceat(data_f32)           NumPy's float32 pairwise summation
//   means[i]  = float32(float64(sums[i]) / count)   np.divide(f32, int64, out=f32)
//   c         = data - means[i]                     float32
//   norms[i]  = sqrt(np.add.reduceat(float64(c)**2))  float64 pairwise summation
//   out       = float32(float64(c) * (1.0 / max(norms[i], FLT_MIN)))
//
// NumPy's reduction of a segment is first + pairwise(rest), where pairwise(a, n) adds
// sequentially for n < 8, keeps eight interleaved partial sums for n <= 128 (combined as
// ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)), then the tail sequentially), and otherwise splits at
// n/2 rounded down to a multiple of 8 (numpy/_core/src/umath/loops_utils.h.src).  Every add is
// an explicit round-to-nearest intrinsic: no FMA contraction, no reassociation.
// tests/test_prep_gpu.py compares the result bit for bit with lkpy_b200.data.knn_item_matrices
// (the SciPy calls of the reference).
//
// One thread per item column: the columns are independent, a column's pairwise tree is walked
// depth-first (eight independent accumulators per leaf give the memory-level parallelism).

#include <algorithm>

#include "common.cuh"

namespace lk {

struct F32Ops {
    using T = float;
    __device__ static T add(T a, T b) { return __fadd_rn(a, b); }
};
struct F64Ops {
    using T = double;
    __device__ static T add(T a, T b) { return __dadd_rn(a, b); }
};

// leaf of the pairwise tree (n <= 128)
template <typename Ops, typename Get>
__device__ __forceinline__ typename Ops::T pairwise_leaf(const Get &get, int64_t lo, int64_t n)
{
    using T = typename Ops::T;
    if (n < 8) {
        T res = (T)0;
        for (int64_t i = 0; i < n; i++) res = Ops::add(res, get(lo + i));
        return res;
    }
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = get(lo + j);
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = Ops::add(r[j], get(lo + i + j));
    }
    T res = Ops::add(Ops::add(Ops::add(r[0], r[1]), Ops::add(r[2], r[3])),
                     Ops::add(Ops::add(r[4], r[5]), Ops::add(r[6], r[7])));
    for (; i < n; i++) res = Ops::add(res, get(lo + i));
    return res;
}

// pairwise(a[lo .. lo+n)) with element i produced by get(i): the recursion of NumPy's pairwise sum
// (split at n/2 rounded down to a multiple of 8 while n > 128) walked depth-first with an explicit
// stack — depth <= log2(n / 64) < 32
template <typename Ops, typename Get>
__device__ typename Ops::T pairwise_sum(const Get &get, int64_t lo, int64_t n)
{
    using T = typename Ops::T;
    struct Frame {
        int64_t lo, n;
        T left;
        int stage;  // 1: left child running, 2: right child running
    };
    Frame stk[32];
    int sp = 0;
    int64_t clo = lo, cn = n;
    for (;;) {
        // descend along left children to a leaf
        while (cn > 128) {
            int64_t n2 = cn / 2;
            n2 -= n2 % 8;
            stk[sp].lo = clo, stk[sp].n = cn, stk[sp].stage = 1;
            sp++;
            cn = n2;
        }
        T res = pairwise_leaf<Ops>(get, clo, cn);
        // climb: a finished left child starts the right one, a finished right child closes the node
        for (;;) {
            if (sp == 0) return res;
            Frame &f = stk[sp - 1];
            if (f.stage == 1) {
                int64_t n2 = f.n / 2;
                n2 -= n2 % 8;
                f.left = res;
                f.stage = 2;
                clo = f.lo + n2, cn = f.n - n2;
                break;
            }
            res = Ops::add(f.left, res);
            sp--;
        }
    }
}

// np.add.reduceat over one segment: first element, then += pairwise(rest)
template <typename Ops, typename Get>
__device__ typename Ops::T segment_reduce(const Get &get, int64_t lo, int64_t n)
{
    if (n == 1) return get(lo);
    return Ops::add(get(lo), pairwise_sum<Ops>(get, lo + 1, n - 1));
}

__global__ void __launch_bounds__(128) knn_prep_columns_kernel(const int32_t *__restrict__ indptr,
                                                             const float *__restrict__ vals, int n_cols,
                                                             int centre, float *__restrict__ means,
                                                             float *__restrict__ out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cols; i += gridDim.x * blockDim.x) {
        const int64_t lo = indptr[i];
        const int64_t n = (int64_t)indptr[i + 1] - lo;
        float mean = 0.0f;
        if (centre && n > 0) {
            const float s = segment_reduce<F32Ops>([&](int64_t p) { return __ldg(vals + p); }, lo, n);
            mean = __double2float_rn(__ddiv_rn((double)s, (double)n));
        }
        if (means != nullptr) means[i] = mean;
        if (n == 0) continue;
        auto centred = [&](int64_t p) { return centre ? __fsub_rn(__ldg(vals + p), mean) : __ldg(vals + p); };
        const double s2 = segment_reduce<F64Ops>(
            [&](int64_t p) {
                const double c = (double)centred(p);
                return __dmul_rn(c, c);
            },
            lo, n);
        const double norm = sqrt(s2);  // IEEE: correctly rounded
        const double recip = __ddiv_rn(1.0, fmax(norm, (double)1.17549435e-38f));
        for (int64_t p = lo; p < lo + n; p++) out[p] = __double2float_rn(__dmul_rn((double)centred(p), recip));
    }
}

}  // namespace lk

using namespace lk;

extern "C" {

int lk_knn_prep_columns(const int32_t *d_indptr, const float *d_vals, int32_t n_cols, int32_t centre,
                        float *d_means, float *d_out, void *stream)
{
    LK_REQUIRE(d_indptr && d_vals && d_out && n_cols >= 0, LK_ERR_INVALID, "lk_knn_prep_columns: bad arguments");
    if (n_cols == 0) return LK_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = std::min((n_cols + 127) / 128, sm_count() * 16);
    knn_prep_columns_kernel<<<blocks, 128, 0, st>>>(d_indptr, d_vals, n_cols, centre, d_means, d_out);
    LK_CUDA_TRY(cudaGetLastError());
    return LK_OK;
}

}  // extern "C"
