// Batch-affine pre-reduction of the bucket runs of the variable-base MSM (instantiated for G1 in msm_ba_g1.cu).
//
// msm_accumulate adds the points of a bucket one by one into an XYZZ accumulator: 8M + 2S = 10 products per point (9.5 with the
// lazy reduction), and it runs at the integer-pipe roof, so only fewer products per point make the MSM faster.  The sum of two
// AFFINE points into an affine point costs 2M + 1S plus ONE field inversion, and an inversion shared by a batch (Montgomery's
// trick) costs 3M per element: 6 products per addition.  A bucket is a sum, so its run of points may be added in any grouping:
//   round 0 : every bucket run (its start aligned to 2^R entries of the sorted index array, padding = sentinel) is cut into
//             adjacent pairs (2g, 2g + 1); all pairs of all windows are independent additions, P1[g] = P[2g] + P[2g+1];
//   round r : P_{r+1}[g] = P_r[2g] + P_r[2g+1], sequential reads;
//   then    : msm_accumulate_pre walks the 2^R-times shorter runs of P_R with the XYZZ mixed addition as before.
// One thread handles M pair slots of a round with ONE inversion (a ~330-product Fermat ladder, so M is in the hundreds): forward
// pass over the slots multiplying the denominators into a running product (kept per slot in local memory), inversion, backward
// pass that peels the individual inverses off and finishes the additions.  Lanes of a warp take adjacent slots, so index and
// round >= 1 point loads are coalesced.  All special cases of BNG1.add are kept (BNG1.java:42-81): a missing operand (padding, a
// base at infinity) copies the other one, P + P doubles (denominator 2y, numerator 3x^2: the reference profiler's input is N copies
// of one base), P + (-P) gives infinity, encoded (0, 0) like every affine infinity in this library.
#pragma once
#include "msm_impl.cuh"

namespace ozk {

static constexpr int kBaMaxM = 512;                 // most pair slots per thread (local array of running products: 16 KB)
static constexpr uint32_t kBaSentinel = 0xffffffffu;   // index-array padding: no point

// The denominator and the case of one slot.  kind: 0 = result is `a` (b missing; also both missing: a = infinity), 1 = result is b,
// 2 = generic addition (den = xb - xa), 3 = doubling (den = 2 ya), 4 = infinity (a = -b)
template <class F>
__device__ __forceinline__ int ba_classify(const Affine<F>& a, bool has_a, const Affine<F>& b, bool has_b, F& den) {
    den = F::one();
    if (!has_b) return 0;
    if (!has_a) return 1;
    if (a.x != b.x) {
        den = F::sub(b.x, a.x);
        return 2;
    }
    if (a.y == b.y) {
        den = F::dbl(a.y);
        return 3;
    }
    return 4;
}

// finishes the addition of one slot given 1 / den
template <class F>
__device__ __forceinline__ Affine<F> ba_finish(int kind, const Affine<F>& a, const Affine<F>& b, const F& inv_den) {
    if (kind == 0) return a;
    if (kind == 1) return b;
    if (kind == 4) return Affine<F>::inf();
    F num;
    if (kind == 2) {
        num = F::sub(b.y, a.y);
    } else {
        const F xx = F::sqr(a.x);
        num = F::add(F::dbl(xx), xx);
    }
    const F lam = F::mul(num, inv_den);
    Affine<F> r;
    r.x = F::sub(F::sub(F::sqr(lam), a.x), b.x);          // doubling: b.x == a.x
    r.y = F::sub(F::mul(lam, F::sub(a.x, r.x)), a.y);
    return r;
}

// ROUND0: operands come from `bases` through the sorted index array; otherwise from the previous round's array `in`.
// Both passes are software-pipelined: the index pair of slot k+2 and the operands of slot k+1 are requested before slot k is
// computed, so the two dependent DRAM latencies of a slot (index, then a random 128-byte line per operand) overlap the
// arithmetic of the slots before it -- without this the kernel is latency-bound (ncu: 6.3 long-scoreboard stalls per issue,
// integer pipe 57 % busy).
template <class F, bool ROUND0>
struct BaSlot {
    // where the operands of slot g come from
    __device__ __forceinline__ static uint2 entry(const uint32_t* __restrict__ sorted, size_t g, size_t total) {
        if (!ROUND0) return make_uint2(0u, 0u);
        return g < total ? reinterpret_cast<const uint2*>(sorted)[g] : make_uint2(kBaSentinel, kBaSentinel);
    }
    __device__ __forceinline__ static const uint4* addr(const uint4* __restrict__ bases, const uint4* __restrict__ in, uint32_t e, size_t g, int which) {
        constexpr int U = FieldIO<F>::kU4;
        if (ROUND0) return bases + (size_t)(e & 0x7fffffffu) * (2 * U);
        return in + (2 * g + which) * (2 * U);
    }
    __device__ __forceinline__ static bool present(uint32_t e, size_t g, size_t total) { return ROUND0 ? e != kBaSentinel : g < total; }
};

template <class F, bool ROUND0>
__global__ void __launch_bounds__(128, sizeof(F) == 32 ? 4 : 1) msm_ba_round(const uint4* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                                         const uint4* __restrict__ in, size_t total, uint4* __restrict__ out, int M) {
    using S = BaSlot<F, ROUND0>;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const size_t base = warp * 32 * (size_t)M;
    if (base >= total) return;
    F pre[kBaMaxM];
    F run = F::one();
    auto slot_of = [&](int k) { return base + (size_t)k * 32 + lane; };
    auto load_full = [&](uint2 e, size_t g, Affine<F>& a, bool& ha, Affine<F>& b, bool& hb) {
        ha = S::present(e.x, g, total);
        hb = S::present(e.y, g, total);
        a = Affine<F>::inf();
        b = Affine<F>::inf();
        if (ha) {
            a = load_affine<F>(S::addr(bases, in, e.x, g, 0), 0);
            ha = !a.is_inf();
            if (ROUND0 && ha && (e.x >> 31)) a.y = F::neg(a.y);
        }
        if (hb) {
            b = load_affine<F>(S::addr(bases, in, e.y, g, 1), 0);
            hb = !b.is_inf();
            if (ROUND0 && hb && (e.y >> 31)) b.y = F::neg(b.y);
        }
    };
    // ---- forward: running product of the denominators.  Only the x coordinates are read here; equal or zero x -- a doubling, a
    // cancellation, an operand at infinity, rare except for the profiler's input -- takes the full loads.
    {
        uint2 e0 = S::entry(sorted, slot_of(0), total);                 // entry of the current slot
        uint2 e2 = M > 1 ? S::entry(sorted, slot_of(1), total) : make_uint2(kBaSentinel, kBaSentinel);   // of the next one
        F xa = F::zero(), xb = F::zero();
        {
            const size_t g = slot_of(0);
            if (S::present(e0.x, g, total)) xa = FieldIO<F>::load(S::addr(bases, in, e0.x, g, 0));
            if (S::present(e0.y, g, total)) xb = FieldIO<F>::load(S::addr(bases, in, e0.y, g, 1));
        }
#pragma unroll 1
        for (int k = 0; k < M; k++) {
            const size_t g = slot_of(k);
            // requests for the following slots
            const uint2 e3 = k + 2 < M ? S::entry(sorted, slot_of(k + 2), total) : make_uint2(kBaSentinel, kBaSentinel);
            F xa_n = F::zero(), xb_n = F::zero();
            if (k + 1 < M) {
                const size_t g1 = slot_of(k + 1);
                if (S::present(e2.x, g1, total)) xa_n = FieldIO<F>::load(S::addr(bases, in, e2.x, g1, 0));
                if (S::present(e2.y, g1, total)) xb_n = FieldIO<F>::load(S::addr(bases, in, e2.y, g1, 1));
            }
            F den = F::one();
            const bool ha = S::present(e0.x, g, total), hb = S::present(e0.y, g, total);
            if (ha && hb && xa != xb && !xa.is_zero() && !xb.is_zero()) {
                den = F::sub(xb, xa);
            } else if (ha || hb) {
                Affine<F> a, b;
                bool fa, fb;
                load_full(e0, g, a, fa, b, fb);
                ba_classify(a, fa, b, fb, den);
            }
            run = F::mul(run, den);
            pre[k] = run;
            e0 = e2;
            e2 = e3;
            xa = xa_n;
            xb = xb_n;
        }
    }
    F inv = field_inv_ni(run);
    // ---- backward: inverse of each denominator, then the addition
    {
        uint2 e0 = S::entry(sorted, slot_of(M - 1), total);
        uint2 e2 = M > 1 ? S::entry(sorted, slot_of(M - 2), total) : make_uint2(kBaSentinel, kBaSentinel);
        Affine<F> a, b;
        bool ha, hb;
        load_full(e0, slot_of(M - 1), a, ha, b, hb);
#pragma unroll 1
        for (int k = M - 1; k >= 0; k--) {
            const size_t g = slot_of(k);
            const uint2 e3 = k >= 2 ? S::entry(sorted, slot_of(k - 2), total) : make_uint2(kBaSentinel, kBaSentinel);
            Affine<F> a_n = Affine<F>::inf(), b_n = Affine<F>::inf();
            bool ha_n = false, hb_n = false;
            if (k >= 1) load_full(e2, slot_of(k - 1), a_n, ha_n, b_n, hb_n);
            F den;
            const int kind = ba_classify(a, ha, b, hb, den);
            const F inv_den = k ? F::mul(inv, pre[k - 1]) : inv;
            inv = F::mul(inv, den);
            if (g < total) store_affine<F>(out, g, ba_finish(kind, a, b, inv_den));
            e2 = e3;
            a = a_n;
            b = b_n;
            ha = ha_n;
            hb = hb_n;
        }
    }
}

// msm_accumulate on the pre-reduced runs: task t < nbuckets_total takes bucket order[t]; its run holds ceil(count / 2^rounds) affine
// points at P_R[(w * wstride + start) >> rounds ...]; overflow tasks take segment seg of seg_len ORIGINAL entries (seg_len is a
// multiple of 2^rounds).  Same outputs as msm_accumulate.
template <class F>
__global__ void __launch_bounds__(128, sizeof(F) == 32 ? 4 : 1) msm_accumulate_pre(const uint4* __restrict__ pts, const uint32_t* __restrict__ start,
                                                                               const uint32_t* __restrict__ count,
                                                                               const OvfTask* __restrict__ ovf_tasks, const uint32_t* __restrict__ ovf_count,
                                                                               const uint32_t* __restrict__ order, uint32_t nbuckets_total, uint32_t log_nb,
                                                                               size_t wstride, uint32_t seg_len, uint32_t rounds, uint32_t resume,
                                                                               uint4* __restrict__ buckets, uint4* __restrict__ ovf_partial) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bucket, seg;
    if (t < nbuckets_total) {
        bucket = order[t];
        seg = 0;
    } else {
        const uint32_t k = t - nbuckets_total;
        if (k >= *ovf_count) return;
        bucket = ovf_tasks[k].bucket;
        seg = ovf_tasks[k].seg;
    }
    const uint32_t cnt = count[bucket];
    const uint32_t w = bucket >> log_nb;
    const uint32_t lo = (seg * seg_len) >> rounds;
    const uint32_t hi = (min(cnt, seg * seg_len + seg_len) + (1u << rounds) - 1) >> rounds;
    const size_t first = ((size_t)w * wstride + start[bucket]) >> rounds;
    XYZZ<F> acc = (resume && t < nbuckets_total) ? load_xyzz<F>(buckets, bucket) : XYZZ<F>::inf();
    if (lo < hi) {
        Affine<F> p = load_affine<F>(pts, first + lo);
        for (uint32_t j = lo; j < hi; j++) {
            Affine<F> p_next = p;
            if (j + 1 < hi) p_next = load_affine<F>(pts, first + j + 1);
            xyzz_madd_hot(acc, p);
            p = p_next;
        }
    }
    if (t < nbuckets_total) store_xyzz<F>(buckets, bucket, acc);
    else store_xyzz<F>(ovf_partial, t - nbuckets_total, acc);
}

// ---- host-side launch wrappers (defined in msm_ba_g1.cu) ------------------------------------------------------------------------
struct MsmBaLaunch {
    // out[g] = P[sorted[2g]] + P[sorted[2g+1]], g < total
    int (*round0)(cudaStream_t, const void* bases, const uint32_t* sorted, size_t total, void* out, int M);
    // out[g] = in[2g] + in[2g+1], g < total
    int (*round)(cudaStream_t, const void* in, size_t total, void* out, int M);
    int (*accumulate_pre)(cudaStream_t, const void* pts, const uint32_t* start, const uint32_t* count, const OvfTask* tasks, const uint32_t* ovf_count,
                          const uint32_t* order, uint32_t nbuckets_total, uint32_t log_nb, size_t wstride, uint32_t seg_len, uint32_t rounds,
                          uint32_t resume, uint32_t ovf_cap, void* buckets, void* ovf_partial);
    size_t affine_bytes;
};
extern const MsmBaLaunch kMsmBaG1;

#define OZK_DEFINE_MSM_BA_LAUNCH(F, NAME)                                                                                                  \
    static unsigned NAME##_grid(size_t total, int M) {                                                                                     \
        const size_t warps = (total + (size_t)32 * M - 1) / ((size_t)32 * M);                                                              \
        return (unsigned)((warps * 32 + 127) / 128);                                                                                       \
    }                                                                                                                                      \
    static int NAME##_round0(cudaStream_t s, const void* bases, const uint32_t* sorted, size_t total, void* out, int M) {                 \
        if (total == 0) return 0;                                                                                                          \
        msm_ba_round<F, true><<<NAME##_grid(total, M), 128, 0, s>>>((const uint4*)bases, sorted, nullptr, total, (uint4*)out, M);          \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                                 \
    }                                                                                                                                      \
    static int NAME##_round(cudaStream_t s, const void* in, size_t total, void* out, int M) {                                              \
        if (total == 0) return 0;                                                                                                          \
        msm_ba_round<F, false><<<NAME##_grid(total, M), 128, 0, s>>>(nullptr, nullptr, (const uint4*)in, total, (uint4*)out, M);           \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                                 \
    }                                                                                                                                      \
    static int NAME##_accumulate_pre(cudaStream_t s, const void* pts, const uint32_t* start, const uint32_t* count, const OvfTask* tasks, \
                                     const uint32_t* ovf_count, const uint32_t* order, uint32_t nbt, uint32_t log_nb, size_t wstride,      \
                                     uint32_t seg_len, uint32_t rounds, uint32_t resume, uint32_t ovf_cap, void* buckets, void* ovf_partial) { \
        const size_t total = (size_t)nbt + ovf_cap;                                                                                        \
        msm_accumulate_pre<F><<<(unsigned)((total + 127) / 128), 128, 0, s>>>((const uint4*)pts, start, count, tasks, ovf_count, order, nbt, \
                                                                              log_nb, wstride, seg_len, rounds, resume, (uint4*)buckets,   \
                                                                              (uint4*)ovf_partial);                                        \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                                 \
    }                                                                                                                                      \
    const MsmBaLaunch NAME = {NAME##_round0, NAME##_round, NAME##_accumulate_pre, sizeof(Affine<F>)};

}  // namespace ozk
