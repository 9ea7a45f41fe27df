// BN254a group arithmetic for G1 (over Fq) and G2 (over Fq2 = Fq[u]/(u^2+1)), y^2 = x^3 + b, a = 0.
//
// The reference works in Jacobian coordinates with add-2007-bl / dbl-2009-l (BNG1.java:38-97,133-161,
// BNG2.java:43-126; CUDA copies at algebra_msm_VariableBaseMSM.cu:278-623).  Parity is equality of group
// elements (BNG1.equals compares projectively, BNG1.java:191-224), so the device code is free to use cheaper
// representations: affine inputs (after a batched inversion), XYZZ accumulators (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2; mixed add 8M+2S instead of 11M+5S), and Jacobian only on the wire.
#pragma once
#include "fp256.cuh"

namespace ozk {

// ---- Fq2 (algebra/fields/Fp2.java:44-118 with non-residue p-1, BN254aFq2Parameters.java:38) ---------------
struct Fq2 {
    Fq c0, c1;
    OZK_HD static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    OZK_HD static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    OZK_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    OZK_HD bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
    OZK_HD bool operator!=(const Fq2& b) const { return !(*this == b); }
    OZK_HD static Fq2 add(const Fq2& a, const Fq2& b) { return {Fq::add(a.c0, b.c0), Fq::add(a.c1, b.c1)}; }
    OZK_HD static Fq2 sub(const Fq2& a, const Fq2& b) { return {Fq::sub(a.c0, b.c0), Fq::sub(a.c1, b.c1)}; }
    OZK_HD static Fq2 dbl(const Fq2& a) { return {Fq::dbl(a.c0), Fq::dbl(a.c1)}; }
    OZK_HD static Fq2 neg(const Fq2& a) { return {Fq::neg(a.c0), Fq::neg(a.c1)}; }
    // Karatsuba, u^2 = -1: (a0 b0 - a1 b1) + ((a0+a1)(b0+b1) - a0 b0 - a1 b1) u, with lazy reduction: three 512-bit
    // products, the combinations formed in 512 bits, two Montgomery reductions instead of three (336 instead of 408
    // multiply-adds).  Bounds: t0, t1 < m^2, t2 < 4 m^2 < 2^512; t0 - t1 + m^2 and t2 - t0 - t1 = a0 b1 + a1 b0 lie in
    // [0, 2 m^2) subset [0, m 2^256), which is what redc needs.
    OZK_HD static Fq2 mul(const Fq2& a, const Fq2& b) {
        Fq::Wide t0 = Fq::mul_wide(a.c0, b.c0);
        Fq::Wide t1 = Fq::mul_wide(a.c1, b.c1);
        Fq::Wide t2 = Fq::mul_wide(Fq::add_raw(a.c0, a.c1), Fq::add_raw(b.c0, b.c1));
        Fq2 r;
        r.c1 = Fq::redc(Fq::wide_sub(Fq::wide_sub(t2, t0), t1));
        r.c0 = Fq::redc(Fq::wide_sub_lazy(t0, t1));
        return r;
    }
    // a*b - c*d
    OZK_HD static Fq2 mul_sub(const Fq2& a, const Fq2& b, const Fq2& c, const Fq2& d) { return sub(mul(a, b), mul(c, d)); }
    // complex squaring: (a0+a1)(a0-a1) + 2 a0 a1 u
    OZK_HD static Fq2 sqr(const Fq2& a) {
        Fq t = Fq::mul(a.c0, a.c1);
        Fq s = Fq::mul(Fq::add(a.c0, a.c1), Fq::sub(a.c0, a.c1));
        return {s, Fq::dbl(t)};
    }
    // Fp2.inverse (Fp2.java:109-118): conj(a) / (a0^2 + a1^2)
    OZK_HD static Fq2 inv(const Fq2& a) {
        Fq n = Fq::add(Fq::sqr(a.c0), Fq::sqr(a.c1));
        Fq ni = Fq::inv(n);
        return {Fq::mul(a.c0, ni), Fq::neg(Fq::mul(a.c1, ni))};
    }
    OZK_HD static Fq2 to_mont(const Fq2& a) { return {Fq::to_mont(a.c0), Fq::to_mont(a.c1)}; }
    OZK_HD static Fq2 from_mont(const Fq2& a) { return {Fq::from_mont(a.c0), Fq::from_mont(a.c1)}; }
    OZK_HD bool is_canonical() const { return c0.is_canonical() && c1.is_canonical(); }
};

// ---- point representations --------------------------------------------------------------------------------
// Affine: infinity is encoded as (0, 0), which is on neither curve (b != 0).
template <class F>
struct Affine {
    F x, y;
    OZK_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    OZK_HD static Affine inf() { return {F::zero(), F::zero()}; }
};

// Jacobian (the reference's wire representation): x = X/Z^2, y = Y/Z^3; infinity iff Z == 0 (BNG1.java:103-105).
template <class F>
struct Jacobian {
    F x, y, z;
    OZK_HD bool is_inf() const { return z.is_zero(); }
    OZK_HD static Jacobian inf() { return {F::zero(), F::one(), F::zero()}; }   // (0,1,0), BN254aG1Parameters.java:52-55
};

// XYZZ accumulator: infinity iff ZZ == 0.
template <class F>
struct XYZZ {
    F x, y, zz, zzz;
    OZK_HD bool is_inf() const { return zz.is_zero(); }
    OZK_HD static XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    OZK_HD static XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        return {p.x, p.y, F::one(), F::one()};
    }
};

// 2 * (affine) -> XYZZ  (EFD mdbl-2008-s-1, a = 0).  p must not be infinity; y != 0 on a prime-order curve.
template <class F>
OZK_HD XYZZ<F> xyzz_dbl_affine(const Affine<F>& p) {
    F U = F::dbl(p.y);
    F V = F::sqr(U);
    F W = F::mul(U, V);
    F S = F::mul(p.x, V);
    F xx = F::sqr(p.x);
    F M = F::add(F::dbl(xx), xx);
    XYZZ<F> r;
    r.x = F::sub(F::sqr(M), F::dbl(S));
    r.y = F::mul_sub(M, F::sub(S, r.x), W, p.y);
    r.zz = V;
    r.zzz = W;
    return r;
}

// 2 * XYZZ (EFD dbl-2008-s-1, a = 0)
template <class F>
OZK_HD XYZZ<F> xyzz_dbl(const XYZZ<F>& p) {
    if (p.is_inf()) return p;
    F U = F::dbl(p.y);
    F V = F::sqr(U);
    F W = F::mul(U, V);
    F S = F::mul(p.x, V);
    F xx = F::sqr(p.x);
    F M = F::add(F::dbl(xx), xx);
    XYZZ<F> r;
    r.x = F::sub(F::sqr(M), F::dbl(S));
    r.y = F::mul_sub(M, F::sub(S, r.x), W, p.y);
    r.zz = F::mul(V, p.zz);
    r.zzz = F::mul(W, p.zzz);
    return r;
}

// acc += (affine) q   (EFD madd-2008-s, 8M + 2S) with the special cases the reference handles explicitly in
// BNG1.add (BNG1.java:42-81): O + q, acc + O, acc == q (double), acc == -q (infinity).
template <class F>
OZK_HD void xyzz_madd(XYZZ<F>& acc, const Affine<F>& q) {
    if (q.is_inf()) return;
    if (acc.is_inf()) {
        acc = {q.x, q.y, F::one(), F::one()};
        return;
    }
    F U2 = F::mul(q.x, acc.zz);
    F S2 = F::mul(q.y, acc.zzz);
    F Pp = F::sub(U2, acc.x);
    F Rr = F::sub(S2, acc.y);
    if (Pp.is_zero()) {
        if (Rr.is_zero()) acc = xyzz_dbl_affine(q);
        else acc = XYZZ<F>::inf();
        return;
    }
    F PP = F::sqr(Pp);
    F PPP = F::mul(Pp, PP);
    F Q = F::mul(acc.x, PP);
    F X3 = F::sub(F::sub(F::sqr(Rr), PPP), F::dbl(Q));
    F Y3 = F::mul_sub(Rr, F::sub(Q, X3), acc.y, PPP);
    acc.x = X3;
    acc.y = Y3;
    acc.zz = F::mul(acc.zz, PP);
    acc.zzz = F::mul(acc.zzz, PPP);
}

// acc += q (both XYZZ)  (EFD add-2008-s, 12M + 2S) with the same special cases.
template <class F>
OZK_HD void xyzz_add(XYZZ<F>& acc, const XYZZ<F>& q) {
    if (q.is_inf()) return;
    if (acc.is_inf()) {
        acc = q;
        return;
    }
    F U1 = F::mul(acc.x, q.zz);
    F U2 = F::mul(q.x, acc.zz);
    F S1 = F::mul(acc.y, q.zzz);
    F S2 = F::mul(q.y, acc.zzz);
    F Pp = F::sub(U2, U1);
    F Rr = F::sub(S2, S1);
    if (Pp.is_zero()) {
        if (Rr.is_zero()) acc = xyzz_dbl(acc);
        else acc = XYZZ<F>::inf();
        return;
    }
    F PP = F::sqr(Pp);
    F PPP = F::mul(Pp, PP);
    F Q = F::mul(U1, PP);
    F X3 = F::sub(F::sub(F::sqr(Rr), PPP), F::dbl(Q));
    F Y3 = F::mul_sub(Rr, F::sub(Q, X3), S1, PPP);
    acc.x = X3;
    acc.y = Y3;
    acc.zz = F::mul(F::mul(acc.zz, q.zz), PP);
    acc.zzz = F::mul(F::mul(acc.zzz, q.zzz), PPP);
}

// ---- XYZZ doubling / addition as LEVELS of up to four independent products ----------------------------------------------
// The tail of an MSM (recombining the windows: ~255 dependent doublings on ONE value, and the last levels of the bucket reduction)
// is pure latency: a lone warp is bound by the issue rate of the multiplier (one IMAD.WIDE per four cycles), so a thread doing
// the 9 products of a doubling one after the other takes 9 product times.  The formulas have only 3 (doubling) and 4 (addition)
// DEPENDENT levels of products, so they are written here once as level schedules over a functor
//     mul4(A, B, R):  R[i] = A[i] * B[i],  i < 4
// that the MSM tail kernel implements with the lanes of a warp (msm_impl.cuh, CoopMul4: every lane holds the whole state,
// lane i computes product i -- for Fq2 four lanes share a product, one schoolbook term each -- and the results are
// broadcast with shuffles), and that tests/host_arith_check.cc runs with a plain loop against the oracle.
// Infinity is all zeros and stays all zeros through the doubling; the addition handles its special cases explicitly, like
// xyzz_add (all lanes hold the same values, so those branches are uniform).
template <class F, class Mul4>
OZK_HD XYZZ<F> xyzz_dbl_levels(const XYZZ<F>& p, Mul4&& mul4) {
    F A[4], B[4], R[4];
    const F U = F::dbl(p.y);
    A[0] = U;   B[0] = U;                // V = U^2
    A[1] = p.x; B[1] = p.x;              // XX
    A[2] = F::zero(); B[2] = F::zero();
    A[3] = F::zero(); B[3] = F::zero();
    mul4(A, B, R);
    const F V = R[0];
    const F M = F::add(F::dbl(R[1]), R[1]);
    A[0] = U;   B[0] = V;                // W = U V
    A[1] = M;   B[1] = M;                // M^2
    A[2] = p.x; B[2] = V;                // S = X V
    A[3] = V;   B[3] = p.zz;             // ZZ3
    mul4(A, B, R);
    const F W = R[0], S = R[2];
    XYZZ<F> r;
    r.x = F::sub(R[1], F::dbl(S));
    r.zz = R[3];
    A[0] = M;   B[0] = F::sub(S, r.x);   // M (S - X3)
    A[1] = W;   B[1] = p.y;              // W Y
    A[2] = W;   B[2] = p.zzz;            // ZZZ3
    A[3] = F::zero(); B[3] = F::zero();
    mul4(A, B, R);
    r.y = F::sub(R[0], R[1]);
    r.zzz = R[2];
    return r;
}

template <class F, class Mul4>
OZK_HD XYZZ<F> xyzz_add_levels(const XYZZ<F>& p, const XYZZ<F>& q, Mul4&& mul4) {
    if (q.is_inf()) return p;
    if (p.is_inf()) return q;
    F A[4], B[4], R[4];
    A[0] = p.x; B[0] = q.zz;             // U1
    A[1] = q.x; B[1] = p.zz;             // U2
    A[2] = p.y; B[2] = q.zzz;            // S1
    A[3] = q.y; B[3] = p.zzz;            // S2
    mul4(A, B, R);
    const F U1 = R[0], S1 = R[2];
    const F Pp = F::sub(R[1], U1);
    const F Rr = F::sub(R[3], S1);
    if (Pp.is_zero()) {
        if (Rr.is_zero()) return xyzz_dbl_levels(p, mul4);
        return XYZZ<F>::inf();
    }
    A[0] = Pp;   B[0] = Pp;              // PP
    A[1] = Rr;   B[1] = Rr;              // R^2
    A[2] = p.zz; B[2] = q.zz;
    A[3] = p.zzz; B[3] = q.zzz;
    mul4(A, B, R);
    const F PP = R[0], RR = R[1], ZZ12 = R[2], ZZZ12 = R[3];
    A[0] = Pp;   B[0] = PP;              // PPP
    A[1] = U1;   B[1] = PP;              // Q
    A[2] = ZZ12; B[2] = PP;              // ZZ3
    A[3] = F::zero(); B[3] = F::zero();
    mul4(A, B, R);
    const F PPP = R[0], Q = R[1];
    XYZZ<F> r;
    r.zz = R[2];
    r.x = F::sub(F::sub(RR, PPP), F::dbl(Q));
    A[0] = Rr;    B[0] = F::sub(Q, r.x); // R (Q - X3)
    A[1] = S1;    B[1] = PPP;            // S1 PPP
    A[2] = ZZZ12; B[2] = PPP;            // ZZZ3
    A[3] = F::zero(); B[3] = F::zero();
    mul4(A, B, R);
    r.y = F::sub(R[0], R[1]);
    r.zzz = R[2];
    return r;
}
// the functor of the CPU-side test (and of any caller without lanes to spare)
template <class F>
struct SerialMul4 {
    OZK_HD void operator()(const F (&A)[4], const F (&B)[4], F (&R)[4]) const {
        for (int i = 0; i < 4; i++) R[i] = F::mul(A[i], B[i]);
    }
};

template <class F>
OZK_HD Affine<F> affine_neg(const Affine<F>& p) {
    return {p.x, F::neg(p.y)};
}

// XYZZ -> Jacobian without an inversion: choose Z = ZZ*ZZZ (= z^5), then X_j = x Z^2 = X ZZ ZZZ^2 and
// Y_j = y Z^3 = Y ZZ^3 ZZZ^2.  Any representative is acceptable on the wire (projective equality).
template <class F>
OZK_HD Jacobian<F> xyzz_to_jacobian(const XYZZ<F>& p) {
    if (p.is_inf()) return Jacobian<F>::inf();
    F zzz2 = F::sqr(p.zzz);
    F t = F::mul(p.zz, zzz2);          // ZZ ZZZ^2
    Jacobian<F> r;
    r.x = F::mul(p.x, t);
    r.y = F::mul(F::mul(p.y, t), F::sqr(p.zz));
    r.z = F::mul(p.zz, p.zzz);
    return r;
}

// XYZZ -> affine with one field inversion.
template <class F>
OZK_HD Affine<F> xyzz_to_affine(const XYZZ<F>& p) {
    if (p.is_inf()) return Affine<F>::inf();
    F i = F::inv(F::mul(p.zz, p.zzz));      // 1/(ZZ ZZZ)
    F izz = F::mul(i, p.zzz);
    F izzz = F::mul(i, p.zz);
    return {F::mul(p.x, izz), F::mul(p.y, izzz)};
}

using G1Affine = Affine<Fq>;
using G2Affine = Affine<Fq2>;
using G1XYZZ = XYZZ<Fq>;
using G2XYZZ = XYZZ<Fq2>;
using G1Jac = Jacobian<Fq>;
using G2Jac = Jacobian<Fq2>;

}  // namespace ozk
