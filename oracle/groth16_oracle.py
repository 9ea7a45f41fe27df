"""CPU oracle for the end-to-end flow around the hot path: a restatement of the reference's synthetic R1CS generator,
R1CS-to-QAP reduction, Groth16 serial setup and serial prover (BASELINE.json configs[0]).

TEST INFRASTRUCTURE (see oracle/dizk_oracle.py): only tests/ import this.

Restated (reference paths under src/main/java/):
  serial_construct       profiler/generation/R1CSConstruction.java:31-110
  r1cs_to_qap_relation   reductions/r1cs_to_qap/R1CStoQAP.java:38-97
  r1cs_to_qap_witness    reductions/r1cs_to_qap/R1CStoQAP.java:126-238
  setup                  zk_proof_systems/zkSNARK/SerialSetup.java:32-192   (the pairing value alphaG1betaG2 is not computed)
  prove                  zk_proof_systems/zkSNARK/SerialProver.java:26-119
Every random() in the reference's serial flow is Fp.random(seed = 10) (configuration/Configuration.java:52), so the whole
flow is deterministic: t = alpha = beta = gamma = delta = r = s = a = b = SEED10.

The reference's only end-to-end assertion is Verifier.verify == true (pairings).  Here the toxic waste is known, so the
same equation is checked in the exponent: with A = a*G1, B = b*G2, C = c*G1 the Groth16 equation
e(A,B) = e(alpha,beta) * e(sum_i x_i gammaABC_i, gamma) * e(C, delta) holds iff a*b = alpha*beta + (sum_i x_i gammaABC_i)*gamma + c*delta
in Fr; `proof_exponents` returns (a, b, c) computed directly from the QAP, and the tests require the proof POINTS to equal
a*G1gen, b*G2gen, c*G1gen."""
from __future__ import annotations

from typing import List, Tuple

from . import dizk_oracle as O

R = O.R
SEED = 10


def seed10() -> int:
    return O.fp_random(SEED, R)


# ---- R1CS ------------------------------------------------------------------------------------------------------
# a constraint is (A, B, C) with each a list of (index, coefficient) terms
def serial_construct(num_constraints: int, num_inputs: int):
    """R1CSConstruction.serialConstruct (R1CSConstruction.java:31-110): alternating a+b / a*b chain, last constraint
    (sum)^2.  Returns (constraints, num_inputs, num_auxiliary, primary, auxiliary)."""
    assert num_inputs <= num_constraints + 1
    num_auxiliary = 3 + num_constraints - num_inputs
    num_variables = num_inputs + num_auxiliary
    a = seed10()
    b = seed10()
    full = [1, a, b]
    constraints = []
    for i in range(num_constraints - 1):
        if i % 2 != 0:
            A, B, C = [(i + 1, 1)], [(i + 2, 1)], [(i + 3, 1)]
            tmp = a * b % R
        else:
            A, B, C = [(i + 1, 1), (i + 2, 1)], [(0, 1)], [(i + 3, 1)]
            tmp = (a + b) % R
        a, b = b, tmp
        full.append(tmp)
        constraints.append((A, B, C))
    A = [(i, 1) for i in range(1, num_variables - 1)]
    B = [(i, 1) for i in range(1, num_variables - 1)]
    res = 0
    for i in range(1, num_variables - 1):
        res = (res + full[i]) % R
    C = [(num_variables - 1, 1)]
    full.append(res * res % R)
    constraints.append((A, B, C))
    assert len(full) == num_variables and len(constraints) == num_constraints
    return constraints, num_inputs, num_auxiliary, full[:num_inputs], full[num_inputs:]


def evaluate(lc, assignment) -> int:
    """LinearCombination.evaluate."""
    acc = 0
    for idx, val in lc:
        acc += val * assignment[idx]
    return acc % R


def is_satisfied(constraints, full) -> bool:
    return all(evaluate(A, full) * evaluate(B, full) % R == evaluate(C, full) for A, B, C in constraints)


# ---- QAP -------------------------------------------------------------------------------------------------------
def r1cs_to_qap_relation(constraints, num_inputs, num_variables, t: int):
    """R1CStoQAP.R1CStoQAPRelation (R1CStoQAP.java:38-97)."""
    num_constraints = len(constraints)
    dom = O.SerialFFT(num_constraints + num_inputs)
    At = [0] * num_variables
    Bt = [0] * num_variables
    Ct = [0] * num_variables
    lag = dom.lagrange_coefficients(t)
    for i in range(num_inputs):
        At[i] = lag[num_constraints + i]
    for i, (A, B, C) in enumerate(constraints):
        li = lag[i]
        for idx, val in A:
            At[idx] = (At[idx] + li * val) % R
        for idx, val in B:
            Bt[idx] = (Bt[idx] + li * val) % R
        for idx, val in C:
            Ct[idx] = (Ct[idx] + li * val) % R
    Ht = []
    ti = 1
    for _ in range(dom.domain_size + 1):
        Ht.append(ti)
        ti = ti * t % R
    return {"At": At, "Bt": Bt, "Ct": Ct, "Ht": Ht, "Zt": dom.compute_z(t), "degree": dom.domain_size}


def r1cs_to_qap_witness(constraints, num_inputs, primary, auxiliary, fft=None):
    """R1CStoQAP.R1CStoQAPWitness (R1CStoQAP.java:126-238): coefficients of H = (A*B - C)/Z via 3 inverse FFTs, 3 coset FFTs,
    divideByZOnCoset and one coset inverse FFT.  `fft` may replace the Python transforms with faster ones of the same
    definition (the C oracle)."""
    num_constraints = len(constraints)
    g = O.FR_MULT_GEN
    dom = O.SerialFFT(num_constraints + num_inputs)
    n = dom.domain_size
    full = list(primary) + list(auxiliary)
    f = fft or _PyFFT(dom)
    A = [0] * n
    B = [0] * n
    for i in range(num_inputs):
        A[i + num_constraints] = full[i]
    for i, (a, b, _) in enumerate(constraints):
        A[i] = (evaluate(a, full) + A[i]) % R
        B[i] = evaluate(b, full)
    f.inverse(A)
    f.inverse(B)
    f.coset(A, g)
    f.coset(B, g)
    H = [x * y % R for x, y in zip(A, B)]
    C = [0] * n
    for i, (_, _, c) in enumerate(constraints):
        C[i] = evaluate(c, full)
    f.inverse(C)
    f.coset(C, g)
    H = [(h - c) % R for h, c in zip(H, C)]
    inv = pow(dom.compute_z(g), -1, R)
    H = [h * inv % R for h in H]
    f.coset_inverse(H, g)
    H.append(0)
    return H


class _PyFFT:
    def __init__(self, dom):
        self.dom = dom

    def inverse(self, a):
        self.dom.radix2_inverse_fft(a)

    def coset(self, a, g):
        self.dom.radix2_coset_fft(a, g)

    def coset_inverse(self, a, g):
        self.dom.radix2_coset_inverse_fft(a, g)


class CFFT:
    """Same wrappers over the C oracle's serialRadix2FFT (oracle/dizk_oracle.c), for sizes Python cannot reach."""

    def __init__(self, n):
        from . import c_oracle as C
        self.C = C
        self.n = n
        self.omega = O.root_of_unity(n)

    def _fft(self, a, omega):
        out = self.C.fft_fr(O.pack_scalars(a), O.le32(omega))
        for i in range(self.n):
            a[i] = O.from_le(out[32 * i:32 * i + 32])

    def inverse(self, a):
        self._fft(a, pow(self.omega, -1, R))
        c = pow(self.n, -1, R)
        for i in range(self.n):
            a[i] = a[i] * c % R

    def coset(self, a, g):
        O.multiply_by_coset(a, g)
        self._fft(a, self.omega)

    def coset_inverse(self, a, g):
        self.inverse(a)
        O.multiply_by_coset(a, pow(g, -1, R))


# ---- Groth16 -----------------------------------------------------------------------------------------------------
def setup_scalars(constraints, num_inputs, num_variables):
    """The field part of SerialSetup.generate (SerialSetup.java:43-112,146-150,160): every scalar vector that is then
    encoded by a fixed-base batch MSM, plus the window parameters the Java derives."""
    t = alpha = beta = gamma = delta = seed10()
    inv_gamma = pow(gamma, -1, R)
    inv_delta = pow(delta, -1, R)
    qap = r1cs_to_qap_relation(constraints, num_inputs, num_variables, t)
    abc = [(beta * qap["At"][i] + alpha * qap["Bt"][i] + qap["Ct"][i]) % R for i in range(num_variables)]
    gammaABC = [abc[i] * inv_gamma % R for i in range(num_inputs)]
    deltaABC = [abc[i] * inv_delta % R for i in range(num_inputs, num_variables)]
    non_zero_at = sum(1 for v in qap["At"] if v)
    non_zero_bt = sum(1 for v in qap["Bt"] if v)
    g1 = O.G1.random(SEED)
    g2 = O.G2.random(SEED)
    count_g1 = non_zero_at + non_zero_bt + num_variables
    inverse_delta_zt = qap["Zt"] * inv_delta % R
    Ht = [h * inverse_delta_zt % R for h in qap["Ht"]]
    return {
        "t": t, "alpha": alpha, "beta": beta, "gamma": gamma, "delta": delta, "qap": qap,
        "gammaABC": gammaABC, "deltaABC": deltaABC, "queryH_scalars": Ht,
        "g1": g1, "g2": g2,
        "scalarSizeG1": O.G1.bit_size(g1), "windowSizeG1": O.get_window_size(count_g1, O.G1),
        "scalarSizeG2": O.G2.bit_size(g2), "windowSizeG2": O.get_window_size(non_zero_bt, O.G2),
    }


def setup_literal(constraints, num_inputs, num_variables):
    """SerialSetup.generate restated literally (small sizes only): the proving key as lists of group elements."""
    s = setup_scalars(constraints, num_inputs, num_variables)
    G1, G2 = O.G1, O.G2
    g1, g2 = s["g1"], s["g2"]
    ss1, w1, ss2, w2 = s["scalarSizeG1"], s["windowSizeG1"], s["scalarSizeG2"], s["windowSizeG2"]
    pk = {
        "alphaG1": G1.mul(g1, s["alpha"]), "betaG1": G1.mul(g1, s["beta"]), "betaG2": G2.mul(g2, s["beta"]),
        "deltaG1": G1.mul(g1, s["delta"]), "deltaG2": G2.mul(g2, s["delta"]),
        "deltaABCG1": O.fixed_batch_msm(G1, ss1, w1, g1, s["deltaABC"]),
        "queryA": O.fixed_batch_msm(G1, ss1, w1, g1, s["qap"]["At"]),
        "queryB": list(zip(O.fixed_batch_msm(G1, ss1, w1, g1, s["qap"]["Bt"]), O.fixed_batch_msm(G2, ss2, w2, g2, s["qap"]["Bt"]))),
        "queryH": O.fixed_batch_msm(G1, ss1, w1, g1, s["queryH_scalars"]),
    }
    vk = {"gammaG2": G2.mul(g2, s["gamma"]), "deltaG2": pk["deltaG2"],
          "gammaABCG1": O.fixed_batch_msm(G1, ss1, w1, g1, s["gammaABC"])}
    return s, pk, vk


def prove_literal(pk, constraints, num_inputs, primary, auxiliary, fft=None):
    """SerialProver.prove restated literally (SerialProver.java:26-119)."""
    G1, G2 = O.G1, O.G2
    H = r1cs_to_qap_witness(constraints, num_inputs, primary, auxiliary, fft)
    r = s = seed10()
    num_variables = len(primary) + len(auxiliary)
    rs_delta = G1.mul(pk["deltaG1"], r * s % R)
    qa, qb = pk["queryA"], pk["queryB"]
    ev_at = G1.add(O.serial_msm(G1, primary, qa[:num_inputs]), O.serial_msm(G1, auxiliary, qa[num_inputs:num_variables]))
    bp1, bp2 = O.double_msm(primary, [q[0] for q in qb[:num_inputs]], [q[1] for q in qb[:num_inputs]])
    bw1, bw2 = O.double_msm(auxiliary, [q[0] for q in qb[num_inputs:num_variables]], [q[1] for q in qb[num_inputs:num_variables]])
    ev_b1, ev_b2 = G1.add(bp1, bw1), G2.add(bp2, bw2)
    ev_h = O.serial_msm(G1, H, pk["queryH"])
    num_witness = num_variables - num_inputs
    ev_abc = G1.add(O.serial_msm(G1, auxiliary[:num_witness], pk["deltaABCG1"][:num_witness]), ev_h)
    A = G1.add(G1.add(pk["alphaG1"], ev_at), G1.mul(pk["deltaG1"], r))
    B1 = G1.add(G1.add(pk["betaG1"], ev_b1), G1.mul(pk["deltaG1"], s))
    B2 = G2.add(G2.add(pk["betaG2"], ev_b2), G2.mul(pk["deltaG2"], s))
    C = G1.sub(G1.add(G1.add(ev_abc, G1.mul(A, s)), G1.mul(B1, r)), rs_delta)
    return (A, B2, C), H


def proof_exponents(setup, primary, auxiliary, H) -> Tuple[int, int, int]:
    """Discrete logs (to the bases g1 / g2 of the setup) of the proof elements A, B, C that SerialProver.prove must output,
    computed straight from the QAP values -- the toxic waste is known in this deterministic flow."""
    full = list(primary) + list(auxiliary)
    qap = setup["qap"]
    alpha, beta, delta = setup["alpha"], setup["beta"], setup["delta"]
    r = s = seed10()
    num_inputs = len(primary)
    ev_a = sum(x * y for x, y in zip(full, qap["At"])) % R
    ev_b = sum(x * y for x, y in zip(full, qap["Bt"])) % R
    ev_h = sum(x * y for x, y in zip(H, setup["queryH_scalars"])) % R
    ev_abc = (sum(x * y for x, y in zip(auxiliary, setup["deltaABC"])) + ev_h) % R
    a = (alpha + ev_a + r * delta) % R
    b = (beta + ev_b + s * delta) % R
    c = (ev_abc + a * s + b * r - r * s * delta) % R
    # the Groth16 verification equation in the exponent (Verifier.java:36-51)
    x_gamma = sum(x * y for x, y in zip(primary, setup["gammaABC"])) % R
    assert a * b % R == (alpha * beta + x_gamma * setup["gamma"] + c * delta) % R, "oracle proof does not verify"
    assert num_inputs == len(setup["gammaABC"])
    return a, b, c


def proof_exponents_synthetic(num_constraints: int, num_inputs: int, threads: int = 0) -> Tuple[int, int, int]:
    """The same (a, b, c) as `proof_exponents` for the reference's synthetic circuit, at ANY size, without building the QAP:
    with L_j(t) the Lagrange coefficients of the domain at the setup point t (FFTAuxiliary.java:249-302),
        sum_i z_i At_i = sum_j L_j(t) a_j        (a_j = <A_j, z>, inputs appended: R1CStoQAP.java:54-80 against :143-160)
    and likewise for B and C, and H(t) Z(t) = A(t) B(t) - C(t) because the circuit is satisfied (QAPRelation.isSatisfied), so
        a = alpha + A(t) + r delta,   b = beta + B(t) + s delta,
        c = [beta A_aux(t) + alpha B_aux(t) + C_aux(t) + H(t) Z(t)] / delta + s a + r b - r s delta
    where X_aux(t) = X(t) - (the same sum over the primary variables only).  Everything is O(domain) field work in the C oracle
    (oracle_r1cs_chain, oracle_r1cs_synth_eval, oracle_fr_lagrange_eval); no group operation and no GPU result enters."""
    from . import c_oracle as C
    t = alpha = beta = gamma = delta = r = s = seed10()
    nc, ni = num_constraints, num_inputs
    n = O.SerialFFT(nc + ni).domain_size
    omega = O.root_of_unity(n)
    z = C.r1cs_chain(nc, seed10(), seed10())
    nv = nc + 3
    full = [C.fr_lagrange_eval(v, n, omega, t, threads) for v in C.r1cs_synth_eval(nc, ni, z, nv, n)]
    prim = [C.fr_lagrange_eval(v, n, omega, t, threads) for v in C.r1cs_synth_eval(nc, ni, z, ni, n)]
    at, bt, ct = full
    hz = (at * bt - ct) % R
    abc_aux = (beta * (at - prim[0]) + alpha * (bt - prim[1]) + (ct - prim[2])) % R
    a = (alpha + at + r * delta) % R
    b = (beta + bt + s * delta) % R
    c = ((abc_aux + hz) * pow(delta, -1, R) + s * a + r * b - r * s * delta) % R
    # the verification equation in the exponent (Verifier.java:36-51): the primary part of abc is what gammaABC carries
    x_gamma = (beta * prim[0] + alpha * prim[1] + prim[2]) % R          # (sum_i x_i gammaABC_i) * gamma
    assert a * b % R == (alpha * beta + x_gamma + c * delta) % R, "oracle proof does not verify"
    return a, b, c


def expected_proof_synthetic(num_constraints: int, num_inputs: int, threads: int = 0):
    """The proof points SerialProver.prove must output for the synthetic circuit, as affine-normalised Jacobian triples:
    (a g1, b g2, c g1) with g1 = BNG1.random(seed 10) = seed10 * G1one, g2 = seed10 * G2one (BNG1.java:125-127)."""
    a, b, c = proof_exponents_synthetic(num_constraints, num_inputs, threads)
    rho = seed10()
    A = O.G1.to_affine(O.G1.mul(O.G1.generator, rho * a % R))
    B = O.G2.to_affine(O.G2.mul(O.G2.generator, rho * b % R))
    Cp = O.G1.to_affine(O.G1.mul(O.G1.generator, rho * c % R))
    return A, B, Cp
