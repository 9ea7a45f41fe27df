"""The end-to-end oracle (oracle/groth16_oracle.py): the literal restatement of SerialSetup + SerialProver at a small
size must produce a proof that (a) satisfies the Groth16 verification equation (checked in the exponent, the toxic waste
being the known seed-10 value) and (b) equals the exponent shortcut used at large sizes."""
from oracle import dizk_oracle as O
from oracle import groth16_oracle as G


def test_r1cs_is_satisfied_and_h_divides():
    cons, ni, na, prim, aux = G.serial_construct(16, 4)
    assert len(prim) == 4 and len(aux) == 3 + 16 - 4
    assert G.is_satisfied(cons, prim + aux)
    H = G.r1cs_to_qap_witness(cons, ni, prim, aux)
    n = O.SerialFFT(16 + 4).domain_size
    assert len(H) == n + 1 and H[n] == 0 and H[n - 1] == 0 and H[n - 2] != 0       # SerialProver.java:43-49
    # QAP relation holds at the setup point t: A(t) B(t) - C(t) = H(t) Z(t)   (QAPRelation.isSatisfied)
    t = G.seed10()
    qap = G.r1cs_to_qap_relation(cons, ni, ni + na, t)
    full = prim + aux
    a = sum(x * y for x, y in zip(full, qap["At"])) % O.R
    b = sum(x * y for x, y in zip(full, qap["Bt"])) % O.R
    c = sum(x * y for x, y in zip(full, qap["Ct"])) % O.R
    h = sum(x * y for x, y in zip(H, qap["Ht"])) % O.R
    assert (a * b - c) % O.R == h * qap["Zt"] % O.R
    # the C transforms give the same H
    assert G.r1cs_to_qap_witness(cons, ni, prim, aux, fft=G.CFFT(n)) == H


def test_literal_setup_and_prover_verify_in_the_exponent():
    cons, ni, na, prim, aux = G.serial_construct(8, 3)
    setup, pk, vk = G.setup_literal(cons, ni, ni + na)
    (A, B, C), H = G.prove_literal(pk, cons, ni, prim, aux)
    a, b, c = G.proof_exponents(setup, prim, aux, H)              # asserts the verification equation
    assert O.G1.equals(A, O.G1.mul(setup["g1"], a))
    assert O.G2.equals(B, O.G2.mul(setup["g2"], b))
    assert O.G1.equals(C, O.G1.mul(setup["g1"], c))
    assert setup["scalarSizeG1"] == 253 and setup["scalarSizeG2"] == 254


def test_exponent_shortcut_at_scale_matches_the_literal_flow():
    """proof_exponents_synthetic (C oracle, O(domain), used to check proofs at 2^20..2^24 constraints) gives the same exponents as
    the literal restatement of setup + witness map, and the same proof points as SerialSetup + SerialProver restated."""
    for nc, ni in ((8, 3), (64, 7), (300, 20)):
        cons, _, na, prim, aux = G.serial_construct(nc, ni)
        setup = G.setup_scalars(cons, ni, ni + na)
        H = G.r1cs_to_qap_witness(cons, ni, prim, aux)
        assert G.proof_exponents_synthetic(nc, ni, 2) == G.proof_exponents(setup, prim, aux, H)
    cons, ni, na, prim, aux = G.serial_construct(8, 3)
    _, pk, _ = G.setup_literal(cons, ni, ni + na)
    (A, B, C), _ = G.prove_literal(pk, cons, ni, prim, aux)
    eA, eB, eC = G.expected_proof_synthetic(8, 3, 2)
    assert O.G1.to_affine(A) == eA and O.G2.to_affine(B) == eB and O.G1.to_affine(C) == eC
