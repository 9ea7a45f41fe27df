// Curve-generic kernels of the variable-base MSM (instantiated for G1 in msm_g1.cu and G2 in msm_g2.cu).
//
// Pipeline (msm.cu drives it; replaces pippengerMSMG1/G2, algebra_msm_VariableBaseMSM.cu:1246-1604, which loops over
// windows on the host with ~6c+6 synchronous launches per window and one warp per big integer):
//   convert  : Jacobian canonical bases -> affine Montgomery (batched inversion, Z == 1 fast path)
//   digits   : signed c-bit digits of every scalar, histogram of (window, bucket)            [msm.cu]
//   scan     : per-window exclusive scan -> bucket start offsets, overflow task list         [msm.cu]
//   scatter  : counting-sort the point indices of every window by bucket                     [msm.cu]
//   accumulate : one thread per (window, bucket) adds its run of points into an XYZZ accumulator (mixed add 8M+2S);
//                runs longer than SEG are split into extra tasks and merged by a warp-cooperative reduction
//   reduce   : sum_b (b+1) * B_b per window by a hierarchy of running sums (groups of 8 while a level fills the machine,
//              groups of 2 -- one addition deep -- once it is latency-bound; all windows in parallel)
//   tail     : per-window recombination of the levels and Horner over the windows, every doubling / addition spread over
//              the lanes of a warp (level schedules of curve.cuh); conversion to the canonical Jacobian wire format
#pragma once
#include <cstdlib>

#include "curve.cuh"

namespace ozk {

static constexpr int kSegMax = 1024;     // longest run of points a single accumulate task handles (runtime value <= this)
static constexpr int kWsumLogSBig = 3;   // bucket reduction: groups of 8 (a serial chain of 15 additions per thread) on levels with enough
static constexpr int kWsumLogSSmall = 1; // inputs to fill the machine, groups of 2 (ONE addition per thread and level) below that
static constexpr uint32_t kWsumBigMin = 1u << 18;   // inputs (all windows together) from which a level counts as throughput-bound
static constexpr int kMaxReduceLevels = 24;
static constexpr int kConvBatchMax = 64; // most points per thread in the batched normalisations (one inversion per thread)
// points per thread: large batches amortise the ~380-product inversion, small inputs keep enough threads in flight
static inline int conv_batch_for(size_t n) {
    size_t b = n / 32768;
    return b < 8 ? 8 : (b > (size_t)kConvBatchMax ? kConvBatchMax : (int)b);
}

struct OvfTask {
    uint32_t bucket;   // w * nb + b
    uint32_t seg;      // segment index >= 1 inside the bucket's run
};
struct OvfBucket {
    uint32_t bucket;
    uint32_t first_task;
    uint32_t ntasks;
};

// ---- canonical <-> field element I/O --------------------------------------------------------------------------
__device__ __forceinline__ Fq ld_fq(const uint4* p) {
    uint4 a = p[0], b = p[1];
    Fq r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fq(uint4* p, const Fq& r) {
    p[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    p[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
template <class F> struct FieldIO;
template <> struct FieldIO<Fq> {
    static constexpr int kU4 = 2;      // uint4 per element
    __device__ __forceinline__ static Fq load(const uint4* p) { return ld_fq(p); }
    __device__ __forceinline__ static void store(uint4* p, const Fq& v) { st_fq(p, v); }
};
template <> struct FieldIO<Fq2> {
    static constexpr int kU4 = 4;
    __device__ __forceinline__ static Fq2 load(const uint4* p) { return {ld_fq(p), ld_fq(p + 2)}; }
    __device__ __forceinline__ static void store(uint4* p, const Fq2& v) { st_fq(p, v.c0); st_fq(p + 2, v.c1); }
};

template <class F>
__device__ __forceinline__ Affine<F> load_affine(const uint4* base, size_t idx) {
    const uint4* p = base + idx * (2 * FieldIO<F>::kU4);
    return {FieldIO<F>::load(p), FieldIO<F>::load(p + FieldIO<F>::kU4)};
}
template <class F>
__device__ __forceinline__ void store_affine(uint4* base, size_t idx, const Affine<F>& a) {
    uint4* p = base + idx * (2 * FieldIO<F>::kU4);
    FieldIO<F>::store(p, a.x);
    FieldIO<F>::store(p + FieldIO<F>::kU4, a.y);
}
template <class F>
__device__ __forceinline__ XYZZ<F> load_xyzz(const uint4* base, size_t idx) {
    const uint4* p = base + idx * (4 * FieldIO<F>::kU4);
    return {FieldIO<F>::load(p), FieldIO<F>::load(p + FieldIO<F>::kU4), FieldIO<F>::load(p + 2 * FieldIO<F>::kU4),
            FieldIO<F>::load(p + 3 * FieldIO<F>::kU4)};
}
template <class F>
__device__ __forceinline__ void store_xyzz(uint4* base, size_t idx, const XYZZ<F>& a) {
    uint4* p = base + idx * (4 * FieldIO<F>::kU4);
    FieldIO<F>::store(p, a.x);
    FieldIO<F>::store(p + FieldIO<F>::kU4, a.y);
    FieldIO<F>::store(p + 2 * FieldIO<F>::kU4, a.zz);
    FieldIO<F>::store(p + 3 * FieldIO<F>::kU4, a.zzz);
}

// Out-of-line copies of the rarely taken / cold-path group operations: one compiled body per field instead of one per
// call site (the carry-chain code is large; see DESIGN.md "build time").
template <class F> __device__ __noinline__ void xyzz_add_ni(XYZZ<F>& acc, const XYZZ<F>& q) { xyzz_add(acc, q); }
template <class F> __device__ __noinline__ void xyzz_dbl_ni(XYZZ<F>& p) { p = xyzz_dbl(p); }
template <class F> __device__ __noinline__ void xyzz_dbl_affine_ni(XYZZ<F>& acc, const Affine<F>& q) { acc = xyzz_dbl_affine(q); }
template <class F> __device__ __noinline__ F field_inv_ni(const F& a) { return F::inv(a); }

// Products of the hot loop: inlined for Fq; for Fq2 (CALLS) one called copy of the product / square each (operands and result
// travel in registers).  The inlined G2 loop body is ~128 KB of code against a 32 KB L1.5 instruction cache, and with only two
// warps per scheduler (242 registers) 18 % of the warp samples of msm_accumulate<Fq2> were "no instruction" stalls at 66 % pipe
// utilisation (profiles/r2_ncu_full_msm_accumulate_g2.txt).  An Fq2 product is ~700 instructions with three independent
// carry-chain streams inside, so nothing is lost by not interleaving it with its neighbours -- unlike the 183-instruction Fr
// product of the NTT butterflies, where the same change cost 14 % (ntt.cu).  (The inlined G2 kernel and a 168-register, three-CTA
// build of the called one were measured and removed again: they cost 4 minutes of build time; numbers in msm.cu.)
template <class F, bool CALLS>
struct HotOps {
    __device__ __forceinline__ static F mul(const F& a, const F& b) { return F::mul(a, b); }
    __device__ __forceinline__ static F sqr(const F& a) { return F::sqr(a); }
    __device__ __forceinline__ static F mul_sub(const F& a, const F& b, const F& c, const F& d) { return F::mul_sub(a, b, c, d); }
};
static __device__ __noinline__ Fq2 fq2_mul_ni(Fq2 a, Fq2 b) { return Fq2::mul(a, b); }
static __device__ __noinline__ Fq2 fq2_sqr_ni(Fq2 a) { return Fq2::sqr(a); }
template <>
struct HotOps<Fq2, true> {
    __device__ __forceinline__ static Fq2 mul(const Fq2& a, const Fq2& b) { return fq2_mul_ni(a, b); }
    __device__ __forceinline__ static Fq2 sqr(const Fq2& a) { return fq2_sqr_ni(a); }
    __device__ __forceinline__ static Fq2 mul_sub(const Fq2& a, const Fq2& b, const Fq2& c, const Fq2& d) {
        return Fq2::sub(fq2_mul_ni(a, b), fq2_mul_ni(c, d));
    }
};

// default: inlined products for Fq, called ones for Fq2 (G2 accumulate 2^24: 134.0 -> 120.3 ms, profiles/r2_sweep_g2_called_products.jsonl)
template <class F> struct HotCallsDefault { static constexpr bool value = sizeof(F) != 32; };

// mixed add for the hot loop: the common path is inline, the doubling special case is a call
template <class F, bool CALLS = HotCallsDefault<F>::value>
__device__ __forceinline__ void xyzz_madd_hot(XYZZ<F>& acc, const Affine<F>& q) {
    using H = HotOps<F, CALLS>;
    if (q.is_inf()) return;
    if (acc.is_inf()) {
        acc = {q.x, q.y, F::one(), F::one()};
        return;
    }
    F U2 = H::mul(q.x, acc.zz);
    F S2 = H::mul(q.y, acc.zzz);
    F Pp = F::sub(U2, acc.x);
    F Rr = F::sub(S2, acc.y);
    if (Pp.is_zero()) {
        if (Rr.is_zero()) xyzz_dbl_affine_ni(acc, q);
        else acc = XYZZ<F>::inf();
        return;
    }
    F PP = H::sqr(Pp);
    F PPP = H::mul(Pp, PP);
    F Q = H::mul(acc.x, PP);
    F X3 = F::sub(F::sub(H::sqr(Rr), PPP), F::dbl(Q));
    acc.y = H::mul_sub(Rr, F::sub(Q, X3), acc.y, PPP);       // one reduction for both products (lazy)
    acc.x = X3;
    acc.zz = H::mul(acc.zz, PP);
    acc.zzz = H::mul(acc.zzz, PPP);
}

// full addition for the throughput-bound reduction levels, products through HotOps (Fq2: called)
template <class F, bool CALLS = HotCallsDefault<F>::value>
__device__ __forceinline__ void xyzz_add_hot(XYZZ<F>& acc, const XYZZ<F>& q) {
    using H = HotOps<F, CALLS>;
    if (q.is_inf()) return;
    if (acc.is_inf()) {
        acc = q;
        return;
    }
    F U1 = H::mul(acc.x, q.zz);
    F U2 = H::mul(q.x, acc.zz);
    F S1 = H::mul(acc.y, q.zzz);
    F S2 = H::mul(q.y, acc.zzz);
    F Pp = F::sub(U2, U1);
    F Rr = F::sub(S2, S1);
    if (Pp.is_zero()) {
        if (Rr.is_zero()) xyzz_dbl_ni(acc);
        else acc = XYZZ<F>::inf();
        return;
    }
    F PP = H::sqr(Pp);
    F PPP = H::mul(Pp, PP);
    F Q = H::mul(U1, PP);
    F X3 = F::sub(F::sub(H::sqr(Rr), PPP), F::dbl(Q));
    acc.y = H::mul_sub(Rr, F::sub(Q, X3), S1, PPP);
    acc.x = X3;
    acc.zz = H::mul(H::mul(acc.zz, q.zz), PP);
    acc.zzz = H::mul(H::mul(acc.zzz, q.zzz), PPP);
}

__device__ __forceinline__ Fq canon_one(Fq*) {
    Fq r = Fq::zero();
    r.v[0] = 1;
    return r;
}
__device__ __forceinline__ Fq2 canon_one(Fq2*) { return {canon_one((Fq*)nullptr), Fq::zero()}; }

// ---- convert: Jacobian canonical -> affine Montgomery ------------------------------------------------------------
// in : n x [X|Y|Z] canonical little-endian (the reference wire format, VariableBaseMSM.java:224-227 / :279-285)
// out: n x [x|y] Montgomery, (0,0) for infinity.  flag |= 1 when a coordinate is not reduced.
template <class F>
__global__ void __launch_bounds__(128) msm_convert_bases(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, uint32_t* flag, int batch) {
    constexpr int U = FieldIO<F>::kU4;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    F prefix[kConvBatchMax];
    F zm[kConvBatchMax];
    bool all_unit = true;
    bool bad = false;
    const F one_canon = canon_one((F*)nullptr);
    // forward: running product of the Z's (Z == 0 counts as 1)
#pragma unroll 1
    for (int k = 0; k < batch; k++) {
        size_t i = tid + (size_t)k * nthreads;
        F z = F::one();
        if (i < n) {
            F zc = FieldIO<F>::load(in + i * (3 * U) + 2 * U);
            bad |= !zc.is_canonical();
            if (!zc.is_zero()) {
                if (zc != one_canon) {
                    all_unit = false;
                    z = F::to_mont(zc);
                }
            }
        }
        zm[k] = z;
        prefix[k] = k ? F::mul(prefix[k - 1], z) : z;
    }
    F inv = F::one();
    if (!all_unit) inv = field_inv_ni(prefix[batch - 1]);
#pragma unroll 1
    for (int k = batch - 1; k >= 0; k--) {
        size_t i = tid + (size_t)k * nthreads;
        F zi = F::one();
        if (!all_unit) {
            zi = k ? F::mul(inv, prefix[k - 1]) : inv;
            inv = F::mul(inv, zm[k]);
        }
        if (i >= n) continue;
        const uint4* p = in + i * (3 * U);
        F xc = FieldIO<F>::load(p), yc = FieldIO<F>::load(p + U), zc = FieldIO<F>::load(p + 2 * U);
        bad |= !xc.is_canonical() || !yc.is_canonical();
        Affine<F> a;
        if (zc.is_zero()) {
            a = Affine<F>::inf();
        } else {
            a.x = F::to_mont(xc);
            a.y = F::to_mont(yc);
            if (!all_unit) {
                F zi2 = F::sqr(zi);
                a.x = F::mul(a.x, zi2);
                a.y = F::mul(a.y, F::mul(zi2, zi));
            }
        }
        store_affine<F>(out, i, a);
    }
    if (bad) atomicOr(flag, 1u);
}

// ---- accumulate ---------------------------------------------------------------------------------------------------
// task t < nbuckets_total : bucket order[t], entries [0, min(cnt, seg_len)) of its run  -> buckets[bucket]
// task t >= nbuckets_total: overflow task (bucket, seg), entries [seg*seg_len, ...)     -> ovf_partial[t - nbuckets_total]
// sorted[w * n + pos] = point index | sign << 31 ; start/count are per (window, bucket), start is window-local.
// G1: four 128-thread CTAs per SM (<= 128 registers per thread); G2 needs ~240 registers and runs two.
template <class F, bool CALLS = HotCallsDefault<F>::value>
__global__ void __launch_bounds__(128, sizeof(F) == 32 ? 4 : 1) msm_accumulate(const uint4* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                      const uint32_t* __restrict__ start, const uint32_t* __restrict__ count,
                                                      const OvfTask* __restrict__ ovf_tasks, const uint32_t* __restrict__ ovf_count,
                                                      const uint32_t* __restrict__ order,
                                                      uint32_t nbuckets_total, uint32_t log_nb, size_t n, uint32_t seg_len,
                                                      uint32_t resume, uint4* __restrict__ buckets, uint4* __restrict__ ovf_partial) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bucket, seg;
    if (t < nbuckets_total) {
        // buckets are visited in order of decreasing run length, so the 32 lanes of a warp have (nearly) equal work
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
    uint32_t lo = seg * seg_len;
    uint32_t hi = min(cnt, lo + seg_len);
    const uint32_t* run = sorted + (size_t)w * n + start[bucket];
    // resume != 0: the buckets already hold the sums of earlier slices of the same MSM (host-pointer entry, msm.cu)
    XYZZ<F> acc = (resume && t < nbuckets_total) ? load_xyzz<F>(buckets, bucket) : XYZZ<F>::inf();
    if (lo < hi) {
        uint32_t e = run[lo];
        Affine<F> p = load_affine<F>(bases, e & 0x7fffffffu);
        for (uint32_t j = lo; j < hi; j++) {
            // prefetch the next point while this one is being added
            uint32_t e_next = 0;
            Affine<F> p_next = p;
            if (j + 1 < hi) {
                e_next = run[j + 1];
                p_next = load_affine<F>(bases, e_next & 0x7fffffffu);
            }
            if (e >> 31) p.y = F::neg(p.y);
            xyzz_madd_hot<F, CALLS>(acc, p);
            e = e_next;
            p = p_next;
        }
    }
    if (t < nbuckets_total) store_xyzz<F>(buckets, bucket, acc);
    else store_xyzz<F>(ovf_partial, t - nbuckets_total, acc);
}

// ---- accumulate, shared-memory-resident variant -----------------------------------------------------------------------
// The same tasks as msm_accumulate, but the accumulator and the prefetched next point live in shared memory (one column per
// thread of a [row][thread] array of 16-byte words: conflict-free), so a thread needs registers only for the values in flight
// inside one addition.  That buys a fifth resident CTA per SM for G1 (96 instead of 128 registers) and a third for G2 (168
// instead of 242, where the register-resident kernel runs two warps per scheduler).  The next point arrives by cp.async
// straight into the thread's column, so the prefetch holds no registers either.  Selected by OZK_MSM_SMEM (msm.cu).
template <class F> struct SmIO;
template <> struct SmIO<Fq> {
    __device__ __forceinline__ static Fq ld(const uint4* row) {
        const uint4 a = row[0], b = row[128];
        Fq r;
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
        r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        return r;
    }
    __device__ __forceinline__ static void st(uint4* row, const Fq& r) {
        row[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
        row[128] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
    }
};
template <> struct SmIO<Fq2> {
    __device__ __forceinline__ static Fq2 ld(const uint4* row) { return {SmIO<Fq>::ld(row), SmIO<Fq>::ld(row + 2 * 128)}; }
    __device__ __forceinline__ static void st(uint4* row, const Fq2& r) {
        SmIO<Fq>::st(row, r.c0);
        SmIO<Fq>::st(row + 2 * 128, r.c1);
    }
};
// element `slot` of a thread's column (slot s occupies rows s*U .. s*U + U - 1)
template <class F> __device__ __forceinline__ F sm_ld(const uint4* col, int slot) { return SmIO<F>::ld(col + slot * FieldIO<F>::kU4 * 128); }
template <class F> __device__ __forceinline__ void sm_st(uint4* col, int slot, const F& v) { SmIO<F>::st(col + slot * FieldIO<F>::kU4 * 128, v); }

__device__ __forceinline__ void cp_async16_g2s(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}

// acc (slots 0..3 = x, y, zz, zzz of the column; infinity iff zz == 0) += p, same formulas and special cases as xyzz_madd_hot
template <class F>
__device__ __forceinline__ void xyzz_madd_sm(uint4* acc, const Affine<F>& p) {
    using H = HotOps<F, HotCallsDefault<F>::value>;
    if (p.is_inf()) return;
    const F zz = sm_ld<F>(acc, 2);
    if (zz.is_zero()) {
        sm_st<F>(acc, 0, p.x);
        sm_st<F>(acc, 1, p.y);
        sm_st<F>(acc, 2, F::one());
        sm_st<F>(acc, 3, F::one());
        return;
    }
    const F Pp = F::sub(H::mul(p.x, zz), sm_ld<F>(acc, 0));
    const F Rr = F::sub(H::mul(p.y, sm_ld<F>(acc, 3)), sm_ld<F>(acc, 1));
    if (Pp.is_zero()) {
        if (Rr.is_zero()) {
            XYZZ<F> d;
            xyzz_dbl_affine_ni(d, p);
            sm_st<F>(acc, 0, d.x);
            sm_st<F>(acc, 1, d.y);
            sm_st<F>(acc, 2, d.zz);
            sm_st<F>(acc, 3, d.zzz);
        } else {
            sm_st<F>(acc, 2, F::zero());
        }
        return;
    }
    const F PP = H::sqr(Pp);
    sm_st<F>(acc, 2, H::mul(sm_ld<F>(acc, 2), PP));
    const F PPP = H::mul(Pp, PP);
    const F Q = H::mul(sm_ld<F>(acc, 0), PP);
    sm_st<F>(acc, 3, H::mul(sm_ld<F>(acc, 3), PPP));
    const F X3 = F::sub(F::sub(H::sqr(Rr), PPP), F::dbl(Q));
    sm_st<F>(acc, 0, X3);
    sm_st<F>(acc, 1, H::mul_sub(Rr, F::sub(Q, X3), sm_ld<F>(acc, 1), PPP));
}

template <class F>
__global__ void __launch_bounds__(128, sizeof(F) == 32 ? 5 : 3) msm_accumulate_sm(const uint4* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                      const uint32_t* __restrict__ start, const uint32_t* __restrict__ count,
                                                      const OvfTask* __restrict__ ovf_tasks, const uint32_t* __restrict__ ovf_count,
                                                      const uint32_t* __restrict__ order,
                                                      uint32_t nbuckets_total, uint32_t log_nb, size_t n, uint32_t seg_len,
                                                      uint32_t resume, uint4* __restrict__ buckets, uint4* __restrict__ ovf_partial) {
    extern __shared__ uint4 sm_acc[];
    constexpr int U = FieldIO<F>::kU4;
    uint4* acc = sm_acc + threadIdx.x;                       // rows 0 .. 4U-1
    uint4* pbuf = sm_acc + 4 * U * 128 + threadIdx.x;        // two point buffers of 2U rows each
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
    const uint32_t lo = seg * seg_len;
    const uint32_t hi = min(cnt, lo + seg_len);
    const uint32_t* run = sorted + (size_t)w * n + start[bucket];
    if (resume && t < nbuckets_total) {
        const XYZZ<F> b = load_xyzz<F>(buckets, bucket);
        sm_st<F>(acc, 0, b.x);
        sm_st<F>(acc, 1, b.y);
        sm_st<F>(acc, 2, b.zz);
        sm_st<F>(acc, 3, b.zzz);
    } else {
        sm_st<F>(acc, 2, F::zero());
    }
    if (lo < hi) {
        uint32_t e = run[lo];
        {
            const uint4* src = bases + (size_t)(e & 0x7fffffffu) * (2 * U);
#pragma unroll
            for (int k = 0; k < 2 * U; k++) cp_async16_g2s(pbuf + k * 128, src + k);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        for (uint32_t j = lo; j < hi; j++) {
            uint4* cur = pbuf + ((j - lo) & 1) * (2 * U * 128);
            uint32_t e_next = 0;
            if (j + 1 < hi) {
                e_next = run[j + 1];
                uint4* nxt = pbuf + ((j + 1 - lo) & 1) * (2 * U * 128);
                const uint4* src = bases + (size_t)(e_next & 0x7fffffffu) * (2 * U);
#pragma unroll
                for (int k = 0; k < 2 * U; k++) cp_async16_g2s(nxt + k * 128, src + k);
                asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;\n" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            }
            Affine<F> p = {sm_ld<F>(cur, 0), sm_ld<F>(cur, 1)};
            if (e >> 31) p.y = F::neg(p.y);
            xyzz_madd_sm<F>(acc, p);
            e = e_next;
        }
    }
    XYZZ<F> r = XYZZ<F>::inf();
    {
        const F zz = sm_ld<F>(acc, 2);
        if (!zz.is_zero()) r = {sm_ld<F>(acc, 0), sm_ld<F>(acc, 1), zz, sm_ld<F>(acc, 3)};
    }
    if (t < nbuckets_total) store_xyzz<F>(buckets, bucket, r);
    else store_xyzz<F>(ovf_partial, t - nbuckets_total, r);
}

// buckets[b] += sum of the overflow partials of bucket b.  Two launches over the overflow-bucket list:
//   DENSE == false: one thread per overflow bucket with at most 32 partials (serial sum; the common case for short task lengths)
//   DENSE == true : one warp per overflow bucket with more than 32 partials (lane-strided sums + shuffle butterfly)
template <class F, bool DENSE>
__global__ void __launch_bounds__(128) msm_merge_overflow(const OvfBucket* __restrict__ ovf_buckets, const uint32_t* __restrict__ ovf_bucket_count,
                                                          const uint4* __restrict__ ovf_partial, uint4* __restrict__ buckets) {
    const uint32_t total = *ovf_bucket_count;
    if (!DENSE) {
        const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
        if (idx >= total) return;
        const OvfBucket ob = ovf_buckets[idx];
        if (ob.ntasks > 32) return;
        XYZZ<F> acc = load_xyzz<F>(buckets, ob.bucket);
        for (uint32_t k = 0; k < ob.ntasks; k++) {
            XYZZ<F> q = load_xyzz<F>(ovf_partial, ob.first_task + k);
            xyzz_add_ni(acc, q);
        }
        store_xyzz<F>(buckets, ob.bucket, acc);
    } else {
        const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const uint32_t lane = threadIdx.x & 31;
        if (warp >= total) return;
        const OvfBucket ob = ovf_buckets[warp];
        if (ob.ntasks <= 32) return;
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t k = lane; k < ob.ntasks; k += 32) {
            XYZZ<F> q = load_xyzz<F>(ovf_partial, ob.first_task + k);
            xyzz_add_ni(acc, q);
        }
        constexpr int W = sizeof(XYZZ<F>) / 4;
#pragma unroll 1
        for (int off = 16; off >= 1; off >>= 1) {
            XYZZ<F> other;
            uint32_t* sp = reinterpret_cast<uint32_t*>(&acc);
            uint32_t* dp = reinterpret_cast<uint32_t*>(&other);
#pragma unroll
            for (int i = 0; i < W; i++) dp[i] = __shfl_xor_sync(0xffffffffu, sp[i], off);
            xyzz_add_ni(acc, other);
        }
        if (lane == 0) {
            XYZZ<F> bsum = load_xyzz<F>(buckets, ob.bucket);
            xyzz_add_ni(bsum, acc);
            store_xyzz<F>(buckets, ob.bucket, bsum);
        }
    }
}

// ---- hierarchical bucket reduction ---------------------------------------------------------------------------------
// One launch per level handles several arrays at once (blockIdx.y = job): job 0 reduces the level's input with index
// weights -- for every window and every group g of S consecutive inputs run[g] = sum_j in[gS+j], acc[g] = sum_j j * in[gS+j] --
// and the other jobs carry the acc arrays of the earlier levels one step further towards their plain sums (same shapes), so
// the whole reduction is `levels` launches with a serial chain of 2S-1 additions each.  Inputs are prefetched one ahead.
struct ReduceJob {
    const uint4* in;
    uint4* out_run;      // plain sums of the groups
    uint4* out_acc;      // weighted sums of the groups, or null for a plain-sum job
};
struct ReduceArgs {
    ReduceJob job[kMaxReduceLevels];
    uint32_t njobs;
    uint32_t m_in;       // elements per window in every input of this level
    uint32_t nwin;
    uint32_t s;          // group size of this level (8 or 2)
};

template <class F>
__global__ void __launch_bounds__(128) msm_reduce_level(ReduceArgs a) {
    const ReduceJob jb = a.job[blockIdx.y];
    const uint32_t S = a.s;
    const uint32_t groups = (a.m_in + S - 1) / S;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= groups * a.nwin) return;
    const uint32_t w = t / groups, g = t % groups;
    const size_t base = (size_t)w * a.m_in + (size_t)g * S;
    const uint32_t len = min(S, a.m_in - g * S);
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    XYZZ<F> q = load_xyzz<F>(jb.in, base + len - 1);
    for (int j = (int)len - 1; j >= 1; j--) {
        XYZZ<F> qn = load_xyzz<F>(jb.in, base + j - 1);
        xyzz_add_hot<F>(run, q);
        if (jb.out_acc) {
            if (j == (int)len - 1) acc = run;          // inf + run: a copy, not an addition (groups of 2 then cost ONE addition)
            else xyzz_add_hot<F>(acc, run);
        }
        q = qn;
    }
    xyzz_add_hot<F>(run, q);
    store_xyzz<F>(jb.out_run, (size_t)w * groups + g, run);
    if (jb.out_acc) store_xyzz<F>(jb.out_acc, (size_t)w * groups + g, acc);
}

// ---- tail --------------------------------------------------------------------------------------------------------------
// sum_acc[l][w] (l < nlevels): sum over groups of acc at level l; total[w]: plain sum of all buckets of window w.
// Window value W_w = total[w] + sum_l (prod_{k<l} S_k) * sum_acc[l][w]; result = sum_w 2^(c w) W_w.  Written as canonical
// Jacobian (X|Y|Z, 32 bytes each per base-field element), (0,1,0) for infinity (what the reference emits,
// algebra_msm_VariableBaseMSM.cu:1274-1276).
struct FinalArgs {
    const uint4* sum_acc[kMaxReduceLevels];
    const uint4* total;
    uint32_t nlevels;
    uint32_t nwin;
    uint32_t c;
    uint8_t log_s[kMaxReduceLevels];
};

// mul4 of the level schedules (curve.cuh) over the lanes of one warp.  Every lane holds the same A, B; lane i (mod 4) computes
// product i and the four results are broadcast back with shuffles, so one level costs ONE product time whatever the number of
// products.  Fq2: four lanes share a product, one schoolbook term each (below) -- an Fq2 product in the time of one Fq product.
// All 32 lanes run the same instructions
// (the upper lanes repeat the work of the lower ones), so the whole warp stays converged and holds identical state.
template <class F> struct CoopMul4;
template <> struct CoopMul4<Fq> {
    __device__ __forceinline__ void operator()(const Fq (&A)[4], const Fq (&B)[4], Fq (&R)[4]) const {
        const int i = threadIdx.x & 3;
        Fq a = A[0], b = B[0];
#pragma unroll
        for (int k = 1; k < 4; k++) {
            if (i == k) { a = A[k]; b = B[k]; }
        }
        const Fq r = Fq::mul(a, b);
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int l = 0; l < 8; l++) R[k].v[l] = __shfl_sync(0xffffffffu, r.v[l], k);
        }
    }
};
template <> struct CoopMul4<Fq2> {
    // Four lanes per product, schoolbook: lane 4 i + t computes term t of product i (a0 b0, a1 b1, a0 b1, a1 b0); neighbouring
    // lanes combine their terms with one shuffle exchange (c0 = a0 b0 - a1 b1 in lane t = 0, c1 = a0 b1 + a1 b0 in lane t = 2) BEFORE
    // the broadcast, so a level costs one Fq product, one Fq addition and 72 shuffles (the three-lane Karatsuba form needed three
    // subtractions per product after the broadcast: 12 per level, a third of the level's instructions).
    __device__ __forceinline__ void operator()(const Fq2 (&A)[4], const Fq2 (&B)[4], Fq2 (&R)[4]) const {
        const int lane = threadIdx.x & 15;
        const int i = lane >> 2, t = lane & 3;
        Fq2 a = A[0], b = B[0];
#pragma unroll
        for (int k = 1; k < 4; k++) {
            if (i == k) { a = A[k]; b = B[k]; }
        }
        const Fq x = (t == 0 || t == 2) ? a.c0 : a.c1;
        const Fq y = (t == 0 || t == 3) ? b.c0 : b.c1;
        const Fq r = Fq::mul(x, y);
        Fq o;
#pragma unroll
        for (int l = 0; l < 8; l++) o.v[l] = __shfl_xor_sync(0xffffffffu, r.v[l], 1);
        if (t == 0) o = Fq::neg(o);                       // lanes t = 1, 3 compute values nobody reads
        const Fq c = Fq::add(r, o);
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int l = 0; l < 8; l++) {
                R[k].c0.v[l] = __shfl_sync(0xffffffffu, c.v[l], 4 * k);
                R[k].c1.v[l] = __shfl_sync(0xffffffffu, c.v[l], 4 * k + 2);
            }
        }
    }
};
template <class F> __device__ __noinline__ XYZZ<F> coop_dbl(XYZZ<F> p) { return xyzz_dbl_levels(p, CoopMul4<F>()); }
template <class F> __device__ __noinline__ XYZZ<F> coop_add(XYZZ<F> p, XYZZ<F> q) { return xyzz_add_levels(p, q, CoopMul4<F>()); }

// One warp per window recombines the levels of that window (nlevels doublings-and-additions: the other latency chain of the
// tail); the warp that finishes last (a counter in global memory) then runs the Horner chain over the windows -- one launch,
// no second kernel boundary.  `done` must be zero on entry and is left zero.
template <class F>
__global__ void __launch_bounds__(32) msm_tail(FinalArgs a, uint4* __restrict__ window_vals, uint4* __restrict__ out, uint32_t* __restrict__ done) {
    constexpr int U = FieldIO<F>::kU4;
    const uint32_t w = blockIdx.x;
    {
        XYZZ<F> T = XYZZ<F>::inf();
        for (int l = (int)a.nlevels - 1; l >= 0; l--) {
            for (int d = 0; d < (int)a.log_s[l]; d++) T = coop_dbl<F>(T);
            T = coop_add<F>(T, load_xyzz<F>(a.sum_acc[l], w));
        }
        T = coop_add<F>(T, load_xyzz<F>(a.total, w));
        if (threadIdx.x == 0) store_xyzz<F>(window_vals, w, T);
    }
    __threadfence();
    uint32_t last = 0;
    if (threadIdx.x == 0) last = (atomicAdd(done, 1u) == a.nwin - 1) ? 1u : 0u;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int ww = (int)a.nwin - 1; ww >= 0; ww--) {
        if (ww != (int)a.nwin - 1) {
            for (uint32_t d = 0; d < a.c; d++) acc = coop_dbl<F>(acc);
        }
        // window values were written by other blocks during this launch: read them around L1
        XYZZ<F> q;
        {
            uint4 raw[4 * U];
            const uint4* src = window_vals + (size_t)ww * (4 * U);
#pragma unroll
            for (int k = 0; k < 4 * U; k++) raw[k] = __ldcg(src + k);
            q = load_xyzz<F>(raw, 0);
        }
        acc = coop_add<F>(acc, q);
    }
    if (threadIdx.x == 0) {
        Jacobian<F> j = xyzz_to_jacobian(acc);
        FieldIO<F>::store(out, F::from_mont(j.x));
        FieldIO<F>::store(out + U, F::from_mont(j.y));
        FieldIO<F>::store(out + 2 * U, F::from_mont(j.z));
        *done = 0;
    }
}

// out = sum of k points given in the canonical Jacobian wire format (partial results of a chunked MSM), same format out
template <class F>
__global__ void msm_sum_wire_points(const uint4* __restrict__ in, uint32_t k, uint4* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    constexpr int U = FieldIO<F>::kU4;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = 0; i < k; i++) {
        const uint4* p = in + (size_t)i * (3 * U);
        F zc = FieldIO<F>::load(p + 2 * U);
        if (zc.is_zero()) continue;
        F z = F::to_mont(zc);
        XYZZ<F> q;
        q.x = F::to_mont(FieldIO<F>::load(p));
        q.y = F::to_mont(FieldIO<F>::load(p + U));
        q.zz = F::sqr(z);
        q.zzz = F::mul(q.zz, z);
        xyzz_add_ni(acc, q);
    }
    Jacobian<F> j = xyzz_to_jacobian(acc);
    FieldIO<F>::store(out, F::from_mont(j.x));
    FieldIO<F>::store(out + U, F::from_mont(j.y));
    FieldIO<F>::store(out + 2 * U, F::from_mont(j.z));
}

// ---- host-side launch wrappers (declared here, defined per field in msm_g1.cu / msm_g2.cu) --------------------------
struct MsmLaunch {
    int (*convert)(cudaStream_t, const void* in, void* out, size_t n, uint32_t* flag, int sm_count);
    int (*accumulate)(cudaStream_t, const void* bases, const uint32_t* sorted, const uint32_t* start, const uint32_t* count,
                      const OvfTask* tasks, const uint32_t* ovf_count, const uint32_t* order, uint32_t nbuckets_total, uint32_t log_nb, size_t n,
                      uint32_t seg_len, uint32_t resume,
                      uint32_t ovf_cap, void* buckets, void* ovf_partial);
    int (*merge)(cudaStream_t, const OvfBucket* ob, const uint32_t* ob_count, uint32_t ob_cap, const void* ovf_partial, void* buckets);
    int (*reduce_level)(cudaStream_t, const ReduceArgs& a);
    int (*final)(cudaStream_t, const FinalArgs& a, void* window_vals, void* out, uint32_t* done);
    int (*sum_wire)(cudaStream_t, const void* in, uint32_t k, void* out);
    size_t affine_bytes;     // per point
    size_t jac_bytes;        // per point, wire format
    size_t xyzz_bytes;
};

extern const MsmLaunch kMsmG1;
extern const MsmLaunch kMsmG2;
// which accumulate kernel runs (msm.cu: OZK_MSM_SMEM bit 0 = G1, bit 1 = G2)
bool msm_use_smem_accumulate(bool g2);

#define OZK_DEFINE_MSM_LAUNCH(F, NAME)                                                                                         \
    static int NAME##_convert(cudaStream_t s, const void* in, void* out, size_t n, uint32_t* flag, int sm_count) {             \
        const int batch = conv_batch_for(n);                                                                                   \
        size_t threads_needed = (n + batch - 1) / batch;                                                                       \
        unsigned grid = (unsigned)((threads_needed + 127) / 128);                                                              \
        if (grid == 0) grid = 1;                                                                                               \
        (void)sm_count;                                                                                                        \
        msm_convert_bases<F><<<grid, 128, 0, s>>>((const uint4*)in, (uint4*)out, n, flag, batch);                              \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                     \
    }                                                                                                                          \
    static int NAME##_accumulate(cudaStream_t s, const void* bases, const uint32_t* sorted, const uint32_t* start,             \
                                 const uint32_t* count, const OvfTask* tasks, const uint32_t* ovf_count, const uint32_t* order, \
                                 uint32_t nbt, uint32_t log_nb, size_t n, uint32_t seg_len, uint32_t resume, uint32_t ovf_cap,   \
                                 void* buckets, void* ovf_partial) {                                                           \
        size_t total = (size_t)nbt + ovf_cap;                                                                                  \
        unsigned grid = (unsigned)((total + 127) / 128);                                                                       \
        if (msm_use_smem_accumulate(sizeof(F) != 32)) {                                                                        \
            const size_t smem = (size_t)8 * FieldIO<F>::kU4 * 128 * 16;                                                        \
            static bool attr_done = false;                                                                                     \
            if (!attr_done) {                                                                                                  \
                if (cudaFuncSetAttribute(msm_accumulate_sm<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1; \
                attr_done = true;                                                                                              \
            }                                                                                                                  \
            msm_accumulate_sm<F><<<grid, 128, smem, s>>>((const uint4*)bases, sorted, start, count, tasks, ovf_count, order, nbt, log_nb, n, \
                                                         seg_len, resume, (uint4*)buckets, (uint4*)ovf_partial);                       \
            return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                 \
        }                                                                                                                      \
        msm_accumulate<F><<<grid, 128, 0, s>>>((const uint4*)bases, sorted, start, count, tasks, ovf_count, order, nbt, log_nb, n, seg_len, \
                                               resume, (uint4*)buckets, (uint4*)ovf_partial);                                          \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                     \
    }                                                                                                                          \
    static int NAME##_merge(cudaStream_t s, const OvfBucket* ob, const uint32_t* ob_count, uint32_t ob_cap,                    \
                            const void* ovf_partial, void* buckets) {                                                          \
        if (ob_cap == 0) return 0;                                                                                             \
        unsigned grid1 = (unsigned)(((size_t)ob_cap + 127) / 128);          /* thread per overflow bucket */                  \
        unsigned grid32 = (unsigned)(((size_t)ob_cap * 32 + 127) / 128);    /* warp per overflow bucket */                    \
        msm_merge_overflow<F, false><<<grid1, 128, 0, s>>>(ob, ob_count, (const uint4*)ovf_partial, (uint4*)buckets);          \
        msm_merge_overflow<F, true><<<grid32, 128, 0, s>>>(ob, ob_count, (const uint4*)ovf_partial, (uint4*)buckets);          \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                     \
    }                                                                                                                          \
    static int NAME##_reduce_level(cudaStream_t s, const ReduceArgs& a) {                                                      \
        uint32_t groups = (a.m_in + a.s - 1) / a.s;                                                                            \
        dim3 grid((groups * a.nwin + 127) / 128, a.njobs);                                                                     \
        msm_reduce_level<F><<<grid, 128, 0, s>>>(a);                                                                           \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                     \
    }                                                                                                                          \
    static int NAME##_final(cudaStream_t s, const FinalArgs& a, void* window_vals, void* out, uint32_t* done) {                \
        msm_tail<F><<<a.nwin, 32, 0, s>>>(a, (uint4*)window_vals, (uint4*)out, done);                                          \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                     \
    }                                                                                                                          \
    static int NAME##_sum_wire(cudaStream_t s, const void* in, uint32_t k, void* out) {                                        \
        msm_sum_wire_points<F><<<1, 32, 0, s>>>((const uint4*)in, k, (uint4*)out);                                             \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                     \
    }                                                                                                                          \
    const MsmLaunch NAME = {NAME##_convert, NAME##_accumulate, NAME##_merge, NAME##_reduce_level, NAME##_final, NAME##_sum_wire, \
                            sizeof(Affine<F>), sizeof(Jacobian<F>), sizeof(XYZZ<F>)};

}  // namespace ozk
