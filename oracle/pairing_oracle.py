"""CPU oracle for the reference's Groth16 verifier: the ate pairing over BN254a and `Verifier.verify` restated with
Python integers.

TEST INFRASTRUCTURE (see oracle/dizk_oracle.py): only tests/ import this.  The reference's only end-to-end assertion is
`Verifier.verify == true` (SerialzkSNARKTest.java:63-78); with this module the tests can make the same assertion on proofs
whose MSMs and transforms ran on the GPU, independently of the "equation in the exponent" check of groth16_oracle.py.

Restated (reference paths under src/main/java/):
  doubling_step / mixed_addition_step   algebra/curves/barreto_naehrig/BNPairing.java:84-151
  mul_by_q                               BNPairing.java:153-158
  precompute_g2                          BNPairing.java:283-322
  miller_loop                            BNPairing.java:238-274
  final_exponentiation                   BNPairing.java:168-236 (as one exponent, see FINAL_EXPONENT below)
  reduced_pairing                        BNPairing.java:324-336, bn254a/BN254aPairing.java
  verify                                 zk_proof_systems/zkSNARK/Verifier.java:25-59
  constants                              bn254a/BN254aPublicParameters.java:24-43
Fq12 is Fq2[w]/(w^6 - xi) with xi = 9 + u; the reference's tower element c0 + c1 w with c0 = (z0, z1, z2), c1 = (z3, z4, z5)
over v = w^2 (Fp12_2Over3Over2.java:17-31, Fp6_3Over2.java) is z0 + z3 w + z1 w^2 + z4 w^3 + z2 w^4 + z5 w^5 here, so
`mulBy024(ell0, ellVW, ellVV)` (Fp12_2Over3Over2.java:200-262: x0 at z0, ellVV at z2, ellVW at z4) multiplies by
ell0 + ellVW w^3 + ellVV w^4.  Multiplication is schoolbook on the six Fq2 coefficients: this is a checker, not a fast
pairing.

The final exponentiation of the reference is elt -> elt^((q^6-1)(q^2+1)) followed by the Fuentes-Castaneda chain whose
comments state the exponent q^3 (12z^3+6z^2+4z-1) + q^2 (12z^3+6z^2+6z) + q (12z^3+6z^2+4z) + (12z^3+12z^2+6z+1)
(BNPairing.java:189-224).  The same power is taken here by square-and-multiply, so GT values equal the reference's.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from . import dizk_oracle as O

P = O.P
F2 = O.Fq2Field

XI = (9, 1)                                               # twist, BN254aPublicParameters.java:25
TWIST_COEFF_B = F2.mul(F2.inv(XI), (3, 0))                # twist^-1 * coefficientB, :26
Q_X_MUL_TWIST = (21575463638280843010398324269430826099269044274347216827212613867836435027261,
                 10307601595873709700152284273816112264069230130616436755625194854815875713954)      # :29-31
Q_Y_MUL_TWIST = (2821565182194536844548159561693502659359617185244120367078079554186484126554,
                 3505843767911556378687030309984248845540243509899259641013678093033130930403)       # :32-34
ATE_LOOP_COUNT = 29793968203157093288                     # :37, not negative (:38)
FINAL_EXPONENT_Z = 4965661367192848881                    # :40, not negative (:41)
_z = FINAL_EXPONENT_Z
FINAL_EXPONENT = (P ** 6 - 1) * (P ** 2 + 1) * (
    P ** 3 * (12 * _z ** 3 + 6 * _z ** 2 + 4 * _z - 1) + P ** 2 * (12 * _z ** 3 + 6 * _z ** 2 + 6 * _z)
    + P * (12 * _z ** 3 + 6 * _z ** 2 + 4 * _z) + (12 * _z ** 3 + 12 * _z ** 2 + 6 * _z + 1))

Fq12 = List[Tuple[int, int]]                              # six Fq2 coefficients of 1, w, ..., w^5
FQ12_ONE: Fq12 = [F2.one] + [F2.zero] * 5


def fq12_mul(a: Sequence, b: Sequence) -> Fq12:
    t = [F2.zero] * 11
    for i in range(6):
        if F2.is_zero(a[i]):
            continue
        for j in range(6):
            if F2.is_zero(b[j]):
                continue
            t[i + j] = F2.add(t[i + j], F2.mul(a[i], b[j]))
    return [F2.add(t[k], F2.mul(XI, t[k + 6])) if k < 5 else t[k] for k in range(6)]      # w^6 = xi


def fq12_pow(a: Sequence, e: int) -> Fq12:
    r = FQ12_ONE
    for bit in bin(e)[2:]:
        r = fq12_mul(r, r)
        if bit == "1":
            r = fq12_mul(r, a)
    return r


def mul_by_024(f: Sequence, ell0, ell_vw, ell_vv) -> Fq12:
    return fq12_mul(f, [ell0, F2.zero, F2.zero, ell_vw, ell_vv, F2.zero])


def _fq2_frobenius(a):                                     # Fp2.FrobeniusMap(1): conjugation (u^q = -u, q = 3 mod 4)
    return (a[0], (-a[1]) % P)


def doubling_step(two_inv: int, cur: list):
    """BNPairing.doublingStepForFlippedMillerLoop (BNPairing.java:84-117); cur = [X, Y, Z] is updated in place."""
    X, Y, Z = cur
    half = (two_inv, 0)
    A = F2.mul(F2.mul(X, Y), half)
    B = F2.sqr(Y)
    C = F2.sqr(Z)
    D = F2.add(F2.add(C, C), C)
    E = F2.mul(TWIST_COEFF_B, D)
    F = F2.add(F2.add(E, E), E)
    G = F2.mul(F2.add(B, F), half)
    H = F2.sub(F2.sqr(F2.add(Y, Z)), F2.add(B, C))
    I = F2.sub(E, B)
    J = F2.sqr(X)
    E2 = F2.sqr(E)
    cur[0] = F2.mul(A, F2.sub(B, F))
    cur[1] = F2.sub(F2.sqr(G), F2.add(F2.add(E2, E2), E2))
    cur[2] = F2.mul(B, H)
    return (F2.mul(XI, I), F2.neg(H), F2.add(F2.add(J, J), J))          # (ell0, ellVW, ellVV)


def mixed_addition_step(base, cur: list):
    """BNPairing.mixedAdditionStepForFlippedMillerLoop (BNPairing.java:119-151)."""
    X1, Y1, Z1 = cur
    x2, y2 = base[0], base[1]
    D = F2.sub(X1, F2.mul(x2, Z1))
    E = F2.sub(Y1, F2.mul(y2, Z1))
    F = F2.sqr(D)
    G = F2.sqr(E)
    H = F2.mul(D, F)
    I = F2.mul(X1, F)
    J = F2.sub(F2.add(H, F2.mul(Z1, G)), F2.add(I, I))
    cur[0] = F2.mul(D, J)
    cur[1] = F2.sub(F2.mul(E, F2.sub(I, J)), F2.mul(H, Y1))
    cur[2] = F2.mul(Z1, H)
    return (F2.mul(XI, F2.sub(F2.mul(E, x2), F2.mul(D, y2))), D, F2.neg(E))   # (ell0, ellVW = D, ellVV = -E)


def mul_by_q(pt):
    """BNPairing.mulByQ (BNPairing.java:153-158)."""
    return (F2.mul(Q_X_MUL_TWIST, _fq2_frobenius(pt[0])), F2.mul(Q_Y_MUL_TWIST, _fq2_frobenius(pt[1])), _fq2_frobenius(pt[2]))


def _loop_bits():
    """The bits of ateLoopCount below its most significant one, MSB first (BNPairing.java:247-254,296-303)."""
    return [(ATE_LOOP_COUNT >> i) & 1 for i in range(ATE_LOOP_COUNT.bit_length() - 2, -1, -1)]


def precompute_g2(Q):
    """BNPairing.precomputeG2 (BNPairing.java:283-322): the line coefficients of the flipped Miller loop."""
    qa = O.G2.to_affine(Q)
    two_inv = pow(2, -1, P)
    R = [qa[0], qa[1], F2.one]
    coeffs = []
    for bit in _loop_bits():
        coeffs.append(doubling_step(two_inv, R))
        if bit:
            coeffs.append(mixed_addition_step(qa, R))
    q1 = mul_by_q((qa[0], qa[1], F2.one))
    q2 = mul_by_q(q1)
    assert q1[2] == F2.one and q2[2] == F2.one
    q2 = (q2[0], F2.neg(q2[1]), q2[2])
    coeffs.append(mixed_addition_step(q1, R))
    coeffs.append(mixed_addition_step(q2, R))
    return coeffs


def miller_loop(Pt, coeffs) -> Fq12:
    """BNPairing.millerLoop (BNPairing.java:238-274) on the affine G1 point and the precomputed G2 coefficients."""
    pa = O.G1.to_affine(Pt)
    px, py = pa[0], pa[1]

    def line(f, c):
        ell0, ell_vw, ell_vv = c
        return mul_by_024(f, ell0, (ell_vw[0] * py % P, ell_vw[1] * py % P), (ell_vv[0] * px % P, ell_vv[1] * px % P))

    f = FQ12_ONE
    idx = 0
    for bit in _loop_bits():
        f = fq12_mul(f, f)
        f = line(f, coeffs[idx]); idx += 1
        if bit:
            f = line(f, coeffs[idx]); idx += 1
    f = line(f, coeffs[idx]); idx += 1
    f = line(f, coeffs[idx])
    return f


def reduced_pairing(Pt, Q) -> Fq12:
    """e(P, Q) = finalExponentiation(atePairing(P, Q)) (BNPairing.java:324-336).  Infinity on either side gives one
    (the Java would divide by zero in toAffineCoordinates; the Groth16 flow never pairs infinity)."""
    if O.G1.is_zero(Pt) or O.G2.is_zero(Q):
        return list(FQ12_ONE)
    return fq12_pow(miller_loop(Pt, precompute_g2(Q)), FINAL_EXPONENT)


def verify(alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc: Sequence, primary_input: Sequence[int], proof) -> bool:
    """Verifier.verify (Verifier.java:25-59): e(A, B) == e(alpha, beta) * e(sum_i x_i gammaABC_i, gamma) * e(C, delta).
    `proof` = (gA, gB, gC); primary_input[0] must be one (:33-34).  GT "add" in the Java is multiplication in Fq12
    (BNGT.java:37-39).  alphaG1betaG2 is recomputed here from alpha*G1 and beta*G2 (SerialSetup.java:166)."""
    assert primary_input[0] == 1
    gA, gB, gC = proof
    ab = reduced_pairing(gA, gB)
    alpha_beta = reduced_pairing(alpha_g1, beta_g2)
    c_delta = reduced_pairing(gC, delta_g2)
    evaluation_abc = O.pippenger_msm(O.G1, list(primary_input), list(gamma_abc))
    rhs = fq12_mul(fq12_mul(alpha_beta, reduced_pairing(evaluation_abc, gamma_g2)), c_delta)
    return ab == rhs
