// 256-bit prime-field arithmetic in Montgomery form (8 x 32-bit limbs, R = 2^256) for BN254a Fq and Fr.
//
// Replaces the reference's CGBN warp-cooperative `cgbn_mul` + `cgbn_rem` (division-based reduction, one warp
// per 254-bit integer: algebra_msm_VariableBaseMSM.cu:149-264, algebra_msm_FixedBaseMSM.cu:1241-1266) with one
// thread per element and explicit mad.lo.cc/madc.hi.cc chains.  Values are kept fully reduced in [0, m) so that
// "canonical out"  only needs a from-Montgomery multiply.
#pragma once
#include "consts.cuh"
#include "ptx_arith.cuh"

namespace ozk {

template <class P>
struct Fp {
    uint32_t v[8];

    // ---- constants ---------------------------------------------------------------------------------------
    OZK_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }
    OZK_HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
        return r;
    }
    OZK_HD static Fp rr() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::rr(i);
        return r;
    }

    // ---- predicates --------------------------------------------------------------------------------------
    OZK_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i];
        return o == 0;
    }
    OZK_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    OZK_HD bool operator!=(const Fp& b) const { return !(*this == b); }

    // ---- add / sub (inputs and outputs in [0, m)) ---------------------------------------------------------
    OZK_HD static Fp add(const Fp& a, const Fp& b) {
        using namespace ptx;
        uint32_t s[8], d[8];
        s[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 7; i++) s[i] = addc_cc(a.v[i], b.v[i]);
        s[7] = addc(a.v[7], b.v[7]);                 // a + b < 2m < 2^255: no carry out
        d[0] = sub_cc(s[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < 8; i++) d[i] = subc_cc(s[i], P::mod(i));
        uint32_t borrow = subc(0u, 0u);              // 0xffffffff when s < m
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = borrow ? s[i] : d[i];
        return r;
    }
    OZK_HD static Fp sub(const Fp& a, const Fp& b) {
        using namespace ptx;
        uint32_t d[8];
        d[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) d[i] = subc_cc(a.v[i], b.v[i]);
        uint32_t mask = subc(0u, 0u);                // all ones when a < b
        Fp r;
        r.v[0] = add_cc(d[0], P::mod(0) & mask);
#pragma unroll
        for (int i = 1; i < 7; i++) r.v[i] = addc_cc(d[i], P::mod(i) & mask);
        r.v[7] = addc(d[7], P::mod(7) & mask);
        return r;
    }
    OZK_HD static Fp dbl(const Fp& a) { return add(a, a); }
    OZK_HD static Fp neg(const Fp& a) {
        if (a.is_zero()) return a;
        using namespace ptx;
        Fp r;
        r.v[0] = sub_cc(P::mod(0), a.v[0]);
#pragma unroll
        for (int i = 1; i < 7; i++) r.v[i] = subc_cc(P::mod(i), a.v[i]);
        r.v[7] = subc(P::mod(7), a.v[7]);
        return r;
    }

    // ---- Montgomery multiplication -------------------------------------------------------------------------
    // Operand scanning with the reduction interleaved.  The running sum T is held as two 8-limb accumulators,
    // `e` aligned at limb 0 and `o` aligned at limb 1 (T = e + o * 2^32): products a[j]*b_i with even j land in
    // `e`, with odd j in `o`, so every 64-bit product is added by one lo/hi pair on a single carry chain
    // (= one IMAD.WIDE.X each).  After a round T is divisible by 2^32; dropping the zero limb makes the old `o`
    // the new limb-0 accumulator, so the two arrays swap roles every round.  T < 2m throughout, which is why the
    // `o` chains can never carry out and the `e` chains carry into o[7].
    // Cost: 8 rounds x (16 wide multiply-adds + 1 low multiply) = 136 integer-pipe multiplies.
    struct Acc {
        uint32_t e[8], o[8];
    };
    template <bool FIRST>
    OZK_HD static void round(uint32_t* e, uint32_t* o, const uint32_t* a, uint32_t bi) {
        using namespace ptx;
        // On entry (FIRST == false): `o` is last round's limb-0 accumulator (its limb 0 is zero, limb 1 still has to
        // be folded into e[0]) and `e` is last round's limb-1 accumulator, i.e. already this round's limb-0 one.
        if (FIRST) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                e[j] = mul_lo(a[j], bi);
                e[j + 1] = mul_hi(a[j], bi);
                o[j] = mul_lo(a[j + 1], bi);
                o[j + 1] = mul_hi(a[j + 1], bi);
            }
        } else {
            e[0] = add_cc(e[0], o[1]);
            // new limb-1 accumulator = (old limb-0 accumulator >> 64) + odd products, carry from the fold above
#pragma unroll
            for (int j = 0; j < 6; j += 2) {
                o[j] = madc_lo_cc(a[j + 1], bi, o[j + 2]);
                o[j + 1] = madc_hi_cc(a[j + 1], bi, o[j + 3]);
            }
            o[6] = madc_lo_cc(a[7], bi, 0u);
            o[7] = madc_hi(a[7], bi, 0u);
            e[0] = mad_lo_cc(a[0], bi, e[0]);
            e[1] = madc_hi_cc(a[0], bi, e[1]);
#pragma unroll
            for (int j = 2; j < 8; j += 2) {
                e[j] = madc_lo_cc(a[j], bi, e[j]);
                e[j + 1] = madc_hi_cc(a[j], bi, e[j + 1]);
            }
            o[7] = addc(o[7], 0u);
        }
        uint32_t m = mul_lo(e[0], P::NP0);
        o[0] = mad_lo_cc(P::mod(1), m, o[0]);
        o[1] = madc_hi_cc(P::mod(1), m, o[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            o[j] = madc_lo_cc(P::mod(j + 1), m, o[j]);
            o[j + 1] = madc_hi_cc(P::mod(j + 1), m, o[j + 1]);
        }
        e[0] = mad_lo_cc(P::mod(0), m, e[0]);
        e[1] = madc_hi_cc(P::mod(0), m, e[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            e[j] = madc_lo_cc(P::mod(j), m, e[j]);
            e[j + 1] = madc_hi_cc(P::mod(j), m, e[j + 1]);
        }
        o[7] = addc(o[7], 0u);
    }

    OZK_HD static Fp mul(const Fp& a, const Fp& b) {
        using namespace ptx;
        uint32_t x[8], y[8];
        round<true>(x, y, a.v, b.v[0]);
        round<false>(y, x, a.v, b.v[1]);
        round<false>(x, y, a.v, b.v[2]);
        round<false>(y, x, a.v, b.v[3]);
        round<false>(x, y, a.v, b.v[4]);
        round<false>(y, x, a.v, b.v[5]);
        round<false>(x, y, a.v, b.v[6]);
        round<false>(y, x, a.v, b.v[7]);
        // after the last round: y was the limb-0 accumulator (y[0] == 0), x the limb-1 one.  T / 2^32 = x + (y >> 32).
        uint32_t s[8], d[8];
        s[0] = add_cc(x[0], y[1]);
#pragma unroll
        for (int i = 1; i < 7; i++) s[i] = addc_cc(x[i], y[i + 1]);
        s[7] = addc(x[7], 0u);
        d[0] = sub_cc(s[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < 8; i++) d[i] = subc_cc(s[i], P::mod(i));
        uint32_t borrow = subc(0u, 0u);
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = borrow ? s[i] : d[i];
        return r;
    }
    OZK_HD static Fp sqr(const Fp& a) { return mul(a, a); }

    // ---- conversions ---------------------------------------------------------------------------------------
    OZK_HD static Fp to_mont(const Fp& a) { return mul(a, rr()); }     // a * R
    OZK_HD static Fp from_mont(const Fp& a) {                           // a / R
        Fp o = zero();
        o.v[0] = 1;
        return mul(a, o);
    }
    // true when the raw 256-bit value is < modulus
    OZK_HD bool is_canonical() const {
        using namespace ptx;
        (void)sub_cc(v[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < 8; i++) (void)subc_cc(v[i], P::mod(i));
        return subc(0u, 0u) != 0;
    }

    // ---- exponentiation / inversion (Fermat; Fp.inverse is BigInteger.modInverse, algebra/fields/Fp.java:88-90) --
    OZK_HD static Fp pow_u32(const Fp& a, uint32_t e) {
        Fp r = one(), b = a;
        while (e) {
            if (e & 1) r = mul(r, b);
            b = sqr(b);
            e >>= 1;
        }
        return r;
    }
    // a^(m-2); inverse of zero is zero.
    OZK_HD static Fp inv(const Fp& a) {
        // 4-bit fixed window over the exponent m-2, most significant nibble first.
        Fp tbl[16];
        tbl[0] = one();
        tbl[1] = a;
        for (int i = 2; i < 16; i++) tbl[i] = mul(tbl[i - 1], a);
        uint32_t ex[8];
        {
            // m - 2 (m is odd and its low limb is > 1 for both fields)
#pragma unroll
            for (int i = 0; i < 8; i++) ex[i] = P::mod(i);
            ex[0] -= 2;
        }
        Fp r = one();
        bool started = false;
        for (int i = 7; i >= 0; i--) {
            for (int s = 28; s >= 0; s -= 4) {
                uint32_t nib = (ex[i] >> s) & 15;
                if (started) {
                    r = sqr(r); r = sqr(r); r = sqr(r); r = sqr(r);
                }
                if (nib) {
                    r = started ? mul(r, tbl[nib]) : tbl[nib];
                    started = true;
                }
            }
        }
        return r;
    }
};

using Fq = Fp<FqParams>;
using Fr = Fp<FrParams>;

}  // namespace ozk
