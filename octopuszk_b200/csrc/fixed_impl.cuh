// Curve-generic kernels of the fixed-base batch MSM (instantiated in msm_g1.cu / msm_g2.cu).
//
// Replaces fixed_batch_MSMG1/G2 (algebra_msm_FixedBaseMSM.cu:997-1137): the reference rebuilds a Jacobian window table
// of numWindows x 2^w entries (192 B each) on every call (calculateBaseOuterG1Helper :928, getWindowTableG1 :851) and
// walks it with one warp per scalar (fixedbase_MSM_unit_processing_G1 :750).  Here the result is defined only by
// (s mod 2^(outerc*w)) * B, so the device table has its own shape: signed t-bit digits, affine Montgomery entries
// j * 2^(t k) * B for j = 1..2^(t-1) (64 B each for G1; 16 x 2^15 x 64 B = 32 MiB at t = 16, L2-resident), cached per
// base on the context; one thread per scalar does ceil(255/t) mixed adds into an XYZZ accumulator; a second kernel
// normalises the results to affine (Z = 1) with a batched inversion, so later variable-base MSMs over these points take
// their Z == 1 fast path.
#pragma once
#include "msm_impl.cuh"

namespace ozk {

static constexpr int kFixedPowers = 256;     // 2^i * B for i < 256

// single thread: pow_aff[i] = 2^i * B (affine Montgomery), from the canonical Jacobian base; flag |= 1 if not reduced
template <class F>
__global__ void fixed_powers(const uint4* __restrict__ base_canon, uint4* __restrict__ pow_xyzz, uint32_t* flag) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    constexpr int U = FieldIO<F>::kU4;
    F xc = FieldIO<F>::load(base_canon), yc = FieldIO<F>::load(base_canon + U), zc = FieldIO<F>::load(base_canon + 2 * U);
    if (!xc.is_canonical() || !yc.is_canonical() || !zc.is_canonical()) atomicOr(flag, 1u);
    XYZZ<F> p;
    if (zc.is_zero()) {
        p = XYZZ<F>::inf();
    } else {
        // Jacobian (X, Y, Z) -> XYZZ (X, Y, Z^2, Z^3)
        F z = F::to_mont(zc);
        p.x = F::to_mont(xc);
        p.y = F::to_mont(yc);
        p.zz = F::sqr(z);
        p.zzz = F::mul(p.zz, z);
    }
    for (int i = 0; i < kFixedPowers; i++) {
        store_xyzz<F>(pow_xyzz, i, p);
        xyzz_dbl_ni(p);
    }
}

// in-place style normalisation of an XYZZ array: out_aff[i] = affine(in[i]) (Montgomery, (0,0) for infinity).
// One thread per `batch` strided elements, one inversion per thread (same scheme as msm_convert_bases).
template <class F>
__global__ void __launch_bounds__(128) xyzz_to_affine_batch(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, int batch) {
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    F prefix[kConvBatchMax];
#pragma unroll 1
    for (int k = 0; k < batch; k++) {
        size_t i = tid + (size_t)k * nthreads;
        F d = F::one();
        if (i < n) {
            XYZZ<F> p = load_xyzz<F>(in, i);
            if (!p.is_inf()) d = F::mul(p.zz, p.zzz);
        }
        prefix[k] = k ? F::mul(prefix[k - 1], d) : d;
    }
    F inv = field_inv_ni(prefix[batch - 1]);
#pragma unroll 1
    for (int k = batch - 1; k >= 0; k--) {
        size_t i = tid + (size_t)k * nthreads;
        XYZZ<F> p = XYZZ<F>::inf();
        F d = F::one();
        if (i < n) {
            p = load_xyzz<F>(in, i);
            if (!p.is_inf()) d = F::mul(p.zz, p.zzz);
        }
        F di = k ? F::mul(inv, prefix[k - 1]) : inv;      // 1 / (zz zzz)
        inv = F::mul(inv, d);
        if (i >= n) continue;
        Affine<F> a = Affine<F>::inf();
        if (!p.is_inf()) {
            a.x = F::mul(p.x, F::mul(di, p.zzz));          // X / ZZ
            a.y = F::mul(p.y, F::mul(di, p.zz));           // Y / ZZZ
        }
        store_affine<F>(out, i, a);
    }
}

// same, but writes the canonical Jacobian wire format (x, y, 1) / (0, 1, 0)
template <class F>
__global__ void __launch_bounds__(128) xyzz_to_wire_batch(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, int batch) {
    constexpr int U = FieldIO<F>::kU4;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    F prefix[kConvBatchMax];
#pragma unroll 1
    for (int k = 0; k < batch; k++) {
        size_t i = tid + (size_t)k * nthreads;
        F d = F::one();
        if (i < n) {
            XYZZ<F> p = load_xyzz<F>(in, i);
            if (!p.is_inf()) d = F::mul(p.zz, p.zzz);
        }
        prefix[k] = k ? F::mul(prefix[k - 1], d) : d;
    }
    F inv = field_inv_ni(prefix[batch - 1]);
    const F one_c = canon_one((F*)nullptr);
#pragma unroll 1
    for (int k = batch - 1; k >= 0; k--) {
        size_t i = tid + (size_t)k * nthreads;
        XYZZ<F> p = XYZZ<F>::inf();
        F d = F::one();
        if (i < n) {
            p = load_xyzz<F>(in, i);
            if (!p.is_inf()) d = F::mul(p.zz, p.zzz);
        }
        F di = k ? F::mul(inv, prefix[k - 1]) : inv;
        inv = F::mul(inv, d);
        if (i >= n) continue;
        uint4* o = out + i * (3 * U);
        if (p.is_inf()) {
            FieldIO<F>::store(o, F::zero());
            FieldIO<F>::store(o + U, one_c);
            FieldIO<F>::store(o + 2 * U, F::zero());
        } else {
            // from_mont(a * b) = mont_mul(mont_mul(a, b), 1): fold the conversion into the last product instead:
            // mont_mul(x_mont, from_mont(t)) = x * t canonical.
            F tx = F::from_mont(F::mul(di, p.zzz));
            F ty = F::from_mont(F::mul(di, p.zz));
            FieldIO<F>::store(o, F::mul(p.x, tx));
            FieldIO<F>::store(o + U, F::mul(p.y, ty));
            FieldIO<F>::store(o + 2 * U, one_c);
        }
    }
}

// Jacobian wire format WITHOUT normalisation: (X ZZ ZZZ^2, Y ZZ^3 ZZZ^2, ZZ ZZZ), the accumulator's own denominator, one thread
// per element and no inversion.  This is what the reference's fixed-base outputs look like on the wire -- Jacobian triples with
// arbitrary Z straight out of its window walk (algebra_msm_FixedBaseMSM.cu:750-850) -- and what a later variable-base MSM over
// them has to normalise (OZK_FIXED_KEEP_Z in include/octozk.h).
template <class F>
__global__ void __launch_bounds__(128) xyzz_to_wire_raw(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    constexpr int U = FieldIO<F>::kU4;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const XYZZ<F> p = load_xyzz<F>(in, i);
    uint4* o = out + i * (3 * U);
    if (p.is_inf()) {
        FieldIO<F>::store(o, F::zero());
        FieldIO<F>::store(o + U, canon_one((F*)nullptr));
        FieldIO<F>::store(o + 2 * U, F::zero());
        return;
    }
    const Jacobian<F> j = xyzz_to_jacobian(p);
    FieldIO<F>::store(o, F::from_mont(j.x));
    FieldIO<F>::store(o + U, F::from_mont(j.y));
    FieldIO<F>::store(o + 2 * U, F::from_mont(j.z));
}

// table_xyzz[k * half + (j-1)] = j * 2^(t k) * B for j in 1..half (half = 2^(t-1)); the reference builds its table by binary
// decomposition of every j over the powers of the base (algebra_msm_FixedBaseMSM.cu:851-884), ~t/2 additions per entry.
// Two-level construction (one addition per entry): j = jl + jh 2^h with h = ceil((t-1)/2).
//   sub[k][jl]           = jl * 2^(t k) B,        jl < 2^h                 (binary decomposition, a few thousand entries)
//   sub[k][2^h + jh]     = jh * 2^(h + t k) B,    jh <= 2^(t-1-h)
// normalised to affine, then table[k][j-1] = sub[k][jl] + sub[k][2^h + jh].
__host__ __device__ __forceinline__ uint32_t fixed_sub_h(uint32_t t) { return t / 2; }                       // ceil((t-1)/2)
__host__ __device__ __forceinline__ uint32_t fixed_sub_count(uint32_t t) {
    const uint32_t h = fixed_sub_h(t);
    return (1u << h) + (1u << (t - 1 - h)) + 1u;
}
template <class F>
__global__ void __launch_bounds__(128) fixed_sub_build(const uint4* __restrict__ pow_aff, uint4* __restrict__ sub_xyzz, uint32_t t, uint32_t nwin) {
    const uint32_t S = fixed_sub_count(t), h = fixed_sub_h(t);
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S * nwin) return;
    const uint32_t k = e / S, se = e % S;
    const bool hi = se >= (1u << h);
    const uint32_t j = hi ? se - (1u << h) : se;
    const uint32_t first = k * t + (hi ? h : 0);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t b = 0; b < t; b++) {
        if ((j >> b) & 1) {
            const uint32_t pw = first + b;
            if (pw < (uint32_t)kFixedPowers) {
                Affine<F> q = load_affine<F>(pow_aff, pw);
                xyzz_madd_hot(acc, q);
            }
        }
    }
    store_xyzz<F>(sub_xyzz, e, acc);
}
template <class F>
__global__ void __launch_bounds__(128) fixed_table_combine(const uint4* __restrict__ sub_aff, uint4* __restrict__ table_xyzz, uint32_t t, uint32_t nwin) {
    const uint32_t half = 1u << (t - 1), S = fixed_sub_count(t), h = fixed_sub_h(t);
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)half * nwin) return;
    const uint32_t k = (uint32_t)(e >> (t - 1));
    const uint32_t j = (uint32_t)(e & (half - 1)) + 1;
    const Affine<F> lo = load_affine<F>(sub_aff, (size_t)k * S + (j & ((1u << h) - 1)));
    const Affine<F> hi = load_affine<F>(sub_aff, (size_t)k * S + (1u << h) + (j >> h));
    XYZZ<F> acc = XYZZ<F>::inf();
    xyzz_madd_hot(acc, lo);
    xyzz_madd_hot(acc, hi);
    store_xyzz<F>(table_xyzz, e, acc);
}

// One thread per scalar: acc = sum_k sign_k * table[k][|d_k| - 1] over the signed t-bit digits of
// (s mod 2^bits).  bits = min(outerc * w, 256) is the caller's truncation (the reference walks exactly outerc windows
// of w bits, algebra_msm_FixedBaseMSM.cu:765-779).
template <class F>
__global__ void __launch_bounds__(128) fixed_walk(const uint4* __restrict__ scalars, size_t n, const uint4* __restrict__ table_aff,
                                                  uint32_t t, uint32_t nwin, uint32_t bits, uint4* __restrict__ out_xyzz, uint32_t* flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 a = scalars[2 * i], b = scalars[2 * i + 1];
    uint32_t s[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    {
        bool lt = false, decided = false;
        for (int q = 7; q >= 0 && !decided; q--) {
            const uint32_t m = FrParams::mod(q);
            if (s[q] < m) { lt = true; decided = true; }
            else if (s[q] > m) { decided = true; }
        }
        if (!lt) atomicOr(flag, 2u);
    }
    // truncate to `bits`
    for (int q = 0; q < 8; q++) {
        const uint32_t lo = q * 32;
        if (bits <= lo) s[q] = 0;
        else if (bits < lo + 32) s[q] &= (1u << (bits - lo)) - 1;
    }
    const uint32_t half = 1u << (t - 1);
    XYZZ<F> acc = XYZZ<F>::inf();
    uint32_t carry = 0;
    for (uint32_t k = 0; k < nwin; k++) {
        const uint32_t pos = k * t;
        uint32_t d = carry;
        if (pos < 256) {
            const uint32_t limb = pos >> 5, off = pos & 31;
            uint64_t v = s[limb];
            if (limb + 1 < 8) v |= (uint64_t)s[limb + 1] << 32;
            d += (uint32_t)(v >> off) & ((1u << t) - 1);
        }
        carry = 0;
        bool neg = false;
        if (d > half) {
            d = (1u << t) - d;
            neg = true;
            carry = 1;
        }
        if (d == 0) continue;
        Affine<F> q = load_affine<F>(table_aff, (size_t)k * half + (d - 1));
        if (neg) q.y = F::neg(q.y);
        xyzz_madd_hot(acc, q);
    }
    store_xyzz<F>(out_xyzz, i, acc);
}

struct FixedLaunch {
    int (*powers)(cudaStream_t, const void* base_canon, void* pow_xyzz, uint32_t* flag);
    int (*to_affine)(cudaStream_t, const void* in_xyzz, void* out_aff, size_t n);
    int (*to_wire)(cudaStream_t, const void* in_xyzz, void* out_wire, size_t n);
    int (*to_wire_raw)(cudaStream_t, const void* in_xyzz, void* out_wire, size_t n);
    // sub: scratch of nwin * fixed_sub_count(t) * (xyzz_bytes + affine_bytes) bytes
    int (*table)(cudaStream_t, const void* pow_aff, void* sub, void* table_xyzz, uint32_t t, uint32_t nwin);
    int (*walk)(cudaStream_t, const void* scalars, size_t n, const void* table_aff, uint32_t t, uint32_t nwin, uint32_t bits,
                void* out_xyzz, uint32_t* flag);
    size_t affine_bytes, jac_bytes, xyzz_bytes;
};
extern const FixedLaunch kFixedG1;
extern const FixedLaunch kFixedG2;

#define OZK_DEFINE_FIXED_LAUNCH(F, NAME)                                                                                    \
    static int NAME##_powers(cudaStream_t s, const void* b, void* p, uint32_t* flag) {                                       \
        fixed_powers<F><<<1, 32, 0, s>>>((const uint4*)b, (uint4*)p, flag);                                                  \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                   \
    }                                                                                                                        \
    static unsigned NAME##_batch_grid(size_t n) {                                                                            \
        const int batch = conv_batch_for(n);                                                                                 \
        size_t threads = (n + batch - 1) / batch;                                                                            \
        unsigned g = (unsigned)((threads + 127) / 128);                                                                      \
        return g ? g : 1;                                                                                                    \
    }                                                                                                                        \
    static int NAME##_to_affine(cudaStream_t s, const void* in, void* out, size_t n) {                                       \
        xyzz_to_affine_batch<F><<<NAME##_batch_grid(n), 128, 0, s>>>((const uint4*)in, (uint4*)out, n, conv_batch_for(n));   \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                   \
    }                                                                                                                        \
    static int NAME##_to_wire(cudaStream_t s, const void* in, void* out, size_t n) {                                         \
        xyzz_to_wire_batch<F><<<NAME##_batch_grid(n), 128, 0, s>>>((const uint4*)in, (uint4*)out, n, conv_batch_for(n));     \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                   \
    }                                                                                                                        \
    static int NAME##_to_wire_raw(cudaStream_t s, const void* in, void* out, size_t n) {                                     \
        xyzz_to_wire_raw<F><<<(unsigned)((n + 127) / 128), 128, 0, s>>>((const uint4*)in, (uint4*)out, n);                   \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                   \
    }                                                                                                                        \
    static int NAME##_table(cudaStream_t s, const void* pw, void* sub, void* tb, uint32_t t, uint32_t nwin) {                \
        const size_t nsub = (size_t)fixed_sub_count(t) * nwin;                                                               \
        void* sub_aff = (char*)sub + nsub * sizeof(XYZZ<F>);                                                                 \
        fixed_sub_build<F><<<(unsigned)((nsub + 127) / 128), 128, 0, s>>>((const uint4*)pw, (uint4*)sub, t, nwin);           \
        if (NAME##_to_affine(s, sub, sub_aff, nsub)) return -1;                                                              \
        size_t total = ((size_t)1 << (t - 1)) * nwin;                                                                        \
        fixed_table_combine<F><<<(unsigned)((total + 127) / 128), 128, 0, s>>>((const uint4*)sub_aff, (uint4*)tb, t, nwin);  \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                   \
    }                                                                                                                        \
    static int NAME##_walk(cudaStream_t s, const void* sc, size_t n, const void* tb, uint32_t t, uint32_t nwin, uint32_t bits, \
                           void* out, uint32_t* flag) {                                                                      \
        fixed_walk<F><<<(unsigned)((n + 127) / 128), 128, 0, s>>>((const uint4*)sc, n, (const uint4*)tb, t, nwin, bits,      \
                                                                  (uint4*)out, flag);                                        \
        return cudaGetLastError() == cudaSuccess ? 0 : -1;                                                                   \
    }                                                                                                                        \
    const FixedLaunch NAME = {NAME##_powers, NAME##_to_affine, NAME##_to_wire, NAME##_to_wire_raw, NAME##_table, NAME##_walk,                    \
                              sizeof(Affine<F>), sizeof(Jacobian<F>), sizeof(XYZZ<F>)};

}  // namespace ozk
