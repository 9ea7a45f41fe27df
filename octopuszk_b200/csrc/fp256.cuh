// 256-bit prime-field arithmetic in Montgomery form (8 x 32-bit limbs, R = 2^256) for BN254a Fq and Fr.
//
// Replaces the reference's CGBN warp-cooperative `cgbn_mul` + `cgbn_rem` (division-based reduction, one warp
// per 254-bit integer: algebra_msm_VariableBaseMSM.cu:149-264, algebra_msm_FixedBaseMSM.cu:1241-1266) with one
// thread per element and explicit mad.lo.cc/madc.hi.cc chains.  Values are kept fully reduced in [0, m) so that
// "canonical out"  only needs a from-Montgomery multiply.
#pragma once
#include "chains.cuh"
#include "consts.cuh"
#include "ptx_arith.cuh"

namespace ozk {

template <class P>
struct alignas(16) Fp {
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
    OZK_HD static void load_mod(uint32_t (&m)[8]) {
#pragma unroll
        for (int i = 0; i < 8; i++) m[i] = P::mod(i);
    }
    // r = s >= m ? s - m : s
    OZK_HD static Fp reduce_once(const uint32_t (&s)[8]) {
        uint32_t m[8], d[8], bw;
        load_mod(m);
        chain::sub8(d, bw, s, m);
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = bw ? s[i] : d[i];
        return r;
    }
    OZK_HD static Fp add(const Fp& a, const Fp& b) {
        uint32_t s[8];
        chain::add8(s, a.v, b.v);                    // a + b < 2m < 2^255: no carry out
        return reduce_once(s);
    }
    OZK_HD static Fp sub(const Fp& a, const Fp& b) {
        uint32_t d[8], mm[8], bw;
        chain::sub8(d, bw, a.v, b.v);
#pragma unroll
        for (int i = 0; i < 8; i++) mm[i] = P::mod(i) & bw;   // add the modulus back when a < b
        Fp r;
        chain::add8(r.v, d, mm);
        return r;
    }
    OZK_HD static Fp dbl(const Fp& a) { return add(a, a); }
    OZK_HD static Fp neg(const Fp& a) {
        if (a.is_zero()) return a;
        uint32_t m[8], bw;
        load_mod(m);
        Fp r;
        chain::sub8(r.v, bw, m, a.v);
        return r;
    }

    // ---- Montgomery multiplication -------------------------------------------------------------------------
    // Operand scanning with the reduction interleaved.  The running sum T is held as two 8-limb accumulators,
    // `e` aligned at limb 0 and `o` aligned at limb 1 (T = e + o * 2^32): products a[j]*b_i with even j land in
    // `e`, with odd j in `o`, so every 64-bit product is added by one lo/hi pair on a single carry chain
    // (= one IMAD.WIDE.U32.X each).  After a round T is divisible by 2^32; dropping the zero limb makes the old `o`
    // the new limb-0 accumulator, so the two arrays swap roles every round.  T < 2m throughout, which is why the
    // `o` chains can never carry out and the `e` chains carry into o[7].  (chains.cuh holds the two blocks.)
    // Cost: 8 rounds x (16 wide multiply-adds + 1 low multiply) = 136 integer-pipe multiplies.
    template <bool REDUCE>
    OZK_HD static Fp mul_t(const Fp& a, const Fp& b) {
        uint32_t x[8], y[8], m[8];
        load_mod(m);
        // round 0: plain products, nothing to accumulate yet (x: limb-0 accumulator, y: limb-1 accumulator)
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            uint64_t pe = (uint64_t)a.v[j] * b.v[0];
            uint64_t po = (uint64_t)a.v[j + 1] * b.v[0];
            x[j] = (uint32_t)pe;
            x[j + 1] = (uint32_t)(pe >> 32);
            y[j] = (uint32_t)po;
            y[j + 1] = (uint32_t)(po >> 32);
        }
        chain::mont_round_mp(x, y, m, P::NP0);
        chain::mont_round_ab(y, x, a.v, b.v[1]);
        chain::mont_round_mp(y, x, m, P::NP0);
        chain::mont_round_ab(x, y, a.v, b.v[2]);
        chain::mont_round_mp(x, y, m, P::NP0);
        chain::mont_round_ab(y, x, a.v, b.v[3]);
        chain::mont_round_mp(y, x, m, P::NP0);
        chain::mont_round_ab(x, y, a.v, b.v[4]);
        chain::mont_round_mp(x, y, m, P::NP0);
        chain::mont_round_ab(y, x, a.v, b.v[5]);
        chain::mont_round_mp(y, x, m, P::NP0);
        chain::mont_round_ab(x, y, a.v, b.v[6]);
        chain::mont_round_mp(x, y, m, P::NP0);
        chain::mont_round_ab(y, x, a.v, b.v[7]);
        chain::mont_round_mp(y, x, m, P::NP0);
        // y was the limb-0 accumulator of the last round (y[0] == 0), x the limb-1 one: T / 2^32 = x + (y >> 32) < 2m
        uint32_t sh[8], s[8];
#pragma unroll
        for (int i = 0; i < 7; i++) sh[i] = y[i + 1];
        sh[7] = 0;
        chain::add8(s, x, sh);
        if (!REDUCE) {
            Fp r;
#pragma unroll
            for (int i = 0; i < 8; i++) r.v[i] = s[i];
            return r;
        }
        return reduce_once(s);
    }
    OZK_HD static Fp mul(const Fp& a, const Fp& b) { return mul_t<true>(a, b); }
    OZK_HD static Fp sqr(const Fp& a) { return mul(a, a); }

    // ---- "lazy" domain [0, 2m): the NTT butterflies keep their values only partially reduced ---------------------------------
    // 4m < 2^256 for both fields, so sums and offset differences of values below 2m fit eight limbs without a reduction, and the
    // Montgomery product absorbs the slack: for a < 4m, b < m the running sum of mul_t stays below a + m < 2^256 and the result
    // (a b + q m) / 2^256 < a m / 2^256 + m < 2m (m / 2^256 < 0.19), so the final conditional subtraction is simply left out.
    // Per butterfly that is 40 + 21 instead of 48 + 37 additions / selects around the same 136 multiply-adds.
    //   a * b / 2^256 mod m for a < 4m, b < m; result in [0, 2m)
    OZK_HD static Fp mul_lazy(const Fp& a, const Fp& b) { return mul_t<false>(a, b); }
    //   a + b mod m for a, b in [0, 2m); result in [0, 2m)
    OZK_HD static Fp add_lazy(const Fp& a, const Fp& b) {
        uint32_t s[8], m2[8], d[8], bw;
        chain::add8(s, a.v, b.v);                    // < 4m < 2^256
#pragma unroll
        for (int i = 0; i < 8; i++) m2[i] = P::mod2(i);
        chain::sub8(d, bw, s, m2);
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = bw ? s[i] : d[i];
        return r;
    }
    //   a - b + 2m for a, b in [0, 2m): a representative of a - b in (0, 4m), fit to be the first operand of mul_lazy
    OZK_HD static Fp sub_lazy_wide(const Fp& a, const Fp& b) {
        uint32_t t[8], m2[8], bw;
#pragma unroll
        for (int i = 0; i < 8; i++) m2[i] = P::mod2(i);
        chain::add8(t, a.v, m2);
        Fp r;
        chain::sub8(r.v, bw, t, b.v);
        return r;
    }
    //   a - b mod m for a, b in [0, 2m); result in [0, 2m)
    OZK_HD static Fp sub_lazy(const Fp& a, const Fp& b) {
        uint32_t d[8], mm[8], bw;
        chain::sub8(d, bw, a.v, b.v);
#pragma unroll
        for (int i = 0; i < 8; i++) mm[i] = P::mod2(i) & bw;   // add 2m back when a < b
        Fp r;
        chain::add8(r.v, d, mm);
        return r;
    }
    //   [0, 2m) -> [0, m)
    OZK_HD static Fp reduce_lazy(const Fp& a) { return reduce_once(a.v); }

    // ---- lazy reduction: full 512-bit products, 512-bit add / sub, one Montgomery reduction for a sum of products --------
    // mul_wide is the multiplication above without the m*p half of every round (64 multiply-adds); redc is the other half
    // on its own (72): together the same 136, but a*b - c*d costs 2 x 64 + 72 instead of 2 x 136.
    struct Wide {
        uint32_t v[16];
    };
    // a, b: any 256-bit values
    OZK_HD static Wide mul_wide(const Fp& a, const Fp& b) {
        uint32_t x[8], y[8];
        Wide t;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            uint64_t pe = (uint64_t)a.v[j] * b.v[0];
            uint64_t po = (uint64_t)a.v[j + 1] * b.v[0];
            x[j] = (uint32_t)pe;
            x[j + 1] = (uint32_t)(pe >> 32);
            y[j] = (uint32_t)po;
            y[j + 1] = (uint32_t)(po >> 32);
        }
        // invariant before round i: the limb-0 accumulator's low limb is limb i-1 of the product, final
        t.v[0] = x[0];
        chain::mont_round_ab(y, x, a.v, b.v[1]);
        t.v[1] = y[0];
        chain::mont_round_ab(x, y, a.v, b.v[2]);
        t.v[2] = x[0];
        chain::mont_round_ab(y, x, a.v, b.v[3]);
        t.v[3] = y[0];
        chain::mont_round_ab(x, y, a.v, b.v[4]);
        t.v[4] = x[0];
        chain::mont_round_ab(y, x, a.v, b.v[5]);
        t.v[5] = y[0];
        chain::mont_round_ab(x, y, a.v, b.v[6]);
        t.v[6] = x[0];
        chain::mont_round_ab(y, x, a.v, b.v[7]);
        // y: limb-0 accumulator at limb 7, x: limb-1 accumulator at limb 8
        t.v[7] = y[0];
        uint32_t sh[8];
#pragma unroll
        for (int i = 0; i < 7; i++) sh[i] = y[i + 1];
        sh[7] = 0;
        chain::add8(t.v + 8, x, sh);
        return t;
    }
    // t / 2^256 mod m for t < m * 2^256, result in [0, m)
    OZK_HD static Fp redc(const Wide& t) {
        uint32_t x[8], y[8], m[8];
        load_mod(m);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x[i] = t.v[i];
            y[i] = 0;
        }
        chain::mont_round_mp(x, y, m, P::NP0);
        chain::redc_shift(y, x, t.v[8]);
        chain::mont_round_mp(y, x, m, P::NP0);
        chain::redc_shift(x, y, t.v[9]);
        chain::mont_round_mp(x, y, m, P::NP0);
        chain::redc_shift(y, x, t.v[10]);
        chain::mont_round_mp(y, x, m, P::NP0);
        chain::redc_shift(x, y, t.v[11]);
        chain::mont_round_mp(x, y, m, P::NP0);
        chain::redc_shift(y, x, t.v[12]);
        chain::mont_round_mp(y, x, m, P::NP0);
        chain::redc_shift(x, y, t.v[13]);
        chain::mont_round_mp(x, y, m, P::NP0);
        chain::redc_shift(y, x, t.v[14]);
        chain::mont_round_mp(y, x, m, P::NP0);
        // y: limb-0 accumulator (y[0] == 0), x: limb-1 accumulator; the top input limb joins at limb 7 of the result
        uint32_t sh[8], s[8];
#pragma unroll
        for (int i = 0; i < 7; i++) sh[i] = y[i + 1];
        sh[7] = t.v[15];
        chain::add8(s, x, sh);
        return reduce_once(s);
    }
    OZK_HD static Wide wide_add(const Wide& a, const Wide& b) {
        Wide r;
        uint32_t c, x;
        chain::add8c(r.v, c, x, a.v, b.v, 0u, 0xffffffffu);
        chain::add8c(r.v + 8, c, x, a.v + 8, b.v + 8, c, 0xffffffffu);
        return r;
    }
    // a - b, requires a >= b
    OZK_HD static Wide wide_sub(const Wide& a, const Wide& b) {
        Wide r;
        uint32_t bw, x;
        chain::sub8b(r.v, bw, x, a.v, b.v, 0u);
        chain::sub8b(r.v + 8, bw, x, a.v + 8, b.v + 8, bw & 1u);
        return r;
    }
    // a + b without the final reduction (a, b < m < 2^254, so the sum fits): an operand for mul_wide
    OZK_HD static Fp add_raw(const Fp& a, const Fp& b) {
        Fp r;
        chain::add8(r.v, a.v, b.v);
        return r;
    }
    // a*b - c*d with a single Montgomery reduction
    OZK_HD static Fp mul_sub(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
        return redc(wide_sub_lazy(mul_wide(a, b), mul_wide(c, d)));
    }
    // a + m^2 - b: the non-negative representative of a - b for a, b < m^2 (result < 2 m^2 < m 2^256, fit for redc)
    OZK_HD static Wide wide_sub_lazy(const Wide& a, const Wide& b) {
        Wide k;
#pragma unroll
        for (int i = 0; i < 16; i++) k.v[i] = P::msq(i);
        return wide_sub(wide_add(a, k), b);
    }

    // ---- conversions ---------------------------------------------------------------------------------------
    OZK_HD static Fp to_mont(const Fp& a) { return mul(a, rr()); }     // a * R
    OZK_HD static Fp from_mont(const Fp& a) {                           // a / R
        Fp o = zero();
        o.v[0] = 1;
        return mul(a, o);
    }
    // true when the raw 256-bit value is < modulus
    OZK_HD bool is_canonical() const {
        uint32_t m[8], d[8], bw;
        load_mod(m);
        chain::sub8(d, bw, v, m);
        return bw != 0;
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
