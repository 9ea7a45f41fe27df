"""The pairing verifier of the oracle (oracle/pairing_oracle.py, Verifier.java:25-59 + BNPairing.java restated): bilinearity
and non-degeneracy of the reduced ate pairing, and `verify == true` -- the reference's own end-to-end assertion
(SerialzkSNARKTest.java:63-78) -- on a proof made by the literal restatement of setup + prover."""
import random

from oracle import dizk_oracle as O
from oracle import groth16_oracle as G
from oracle import pairing_oracle as PO


def test_bilinear_and_non_degenerate():
    rng = random.Random(5)
    a, b = rng.randrange(1, O.R), rng.randrange(1, O.R)
    g1, g2 = O.G1.generator, O.G2.generator
    e = PO.reduced_pairing(g1, g2)
    assert e != PO.FQ12_ONE
    assert PO.fq12_pow(e, O.R) == PO.FQ12_ONE                                  # GT has order r
    lhs = PO.reduced_pairing(O.G1.mul(g1, a), O.G2.mul(g2, b))
    assert lhs == PO.fq12_pow(e, a * b % O.R)
    assert lhs == PO.reduced_pairing(O.G1.mul(g1, a * b % O.R), g2)
    # Jacobian inputs with Z != 1 are normalised first (BNPairing.java:277-286)
    p = O.G1.mul(g1, a)
    assert p[2] != 1 and PO.reduced_pairing(p, g2) == PO.fq12_pow(e, a)


def test_verifier_accepts_oracle_proof_and_rejects_a_wrong_one():
    nc, ni = 16, 4
    cons, ni, na, primary, aux = G.serial_construct(nc, ni)
    s, pk, vk = G.setup_literal(cons, ni, ni + na)
    (A, B, C), _H = G.prove_literal(pk, cons, ni, primary, aux)
    assert PO.verify(pk["alphaG1"], pk["betaG2"], vk["gammaG2"], vk["deltaG2"], vk["gammaABCG1"], primary, (A, B, C))
    assert not PO.verify(pk["alphaG1"], pk["betaG2"], vk["gammaG2"], vk["deltaG2"], vk["gammaABCG1"], primary,
                         (A, B, O.G1.add(C, O.G1.generator)))
    bad_input = list(primary)
    bad_input[1] = (bad_input[1] + 1) % O.R
    assert not PO.verify(pk["alphaG1"], pk["betaG2"], vk["gammaG2"], vk["deltaG2"], vk["gammaABCG1"], bad_input, (A, B, C))
