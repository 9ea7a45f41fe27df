"""BASELINE.json configs[0]: serial Groth16 setup + prove on BN254a for the reference's synthetic R1CS
(SerialzkSNARKTest.java:69-93 with R1CSConstruction.serialConstruct), every MSM and FFT on the GPU through
octopuszk_b200/groth16.py, compared with the oracle: the witness polynomial H bit for bit, sampled proving-key elements
and the three proof points as group elements; the proof also satisfies the Groth16 equation (checked in the exponent
by the oracle, the toxic waste being the known seed-10 value) and is accepted by the oracle's pairing verifier
(oracle/pairing_oracle.py, Verifier.java:25-59 restated)."""
import random

import pytest

from oracle import dizk_oracle as O
from oracle import groth16_oracle as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from octopuszk_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("num_constraints,num_inputs", [(64, 7), (1 << 15, 1023)])
def test_serial_groth16_matches_oracle(ctx, num_constraints, num_inputs):
    from octopuszk_b200.groth16 import Groth16
    gz = Groth16(ctx)
    cons, ni, na, prim, aux = Groth16.serial_construct(num_constraints, num_inputs)
    o_cons, _, _, o_prim, o_aux = G.serial_construct(num_constraints, num_inputs)
    assert (cons, prim, aux) == (o_cons, o_prim, o_aux)
    nv = ni + na
    pk, vk, info = gz.setup(cons, ni, nv)
    setup = G.setup_scalars(o_cons, ni, nv)
    assert info["windowSizeG1"] == setup["windowSizeG1"] and info["windowSizeG2"] == setup["windowSizeG2"]
    g1, g2 = setup["g1"], setup["g2"]
    assert O.G1.equals(info["g1"], g1) and O.G2.equals(info["g2"], g2)
    # sampled proving / verification key elements (fixed-base outputs) against scalar * generator
    rng = random.Random(1)
    for i in [0, 1, ni - 1, ni, nv - 1] + [rng.randrange(nv) for _ in range(6)]:
        assert O.G1.equals(pk["queryA"][i], O.G1.mul(g1, setup["qap"]["At"][i]))
        assert O.G1.equals(pk["queryB"][i][0], O.G1.mul(g1, setup["qap"]["Bt"][i]))
    for i in [0, 1, ni, nv - 1]:
        assert O.G2.equals(pk["queryB"][i][1], O.G2.mul(g2, setup["qap"]["Bt"][i]))
    for i in [0, len(setup["deltaABC"]) - 1]:
        assert O.G1.equals(pk["deltaABCG1"][i], O.G1.mul(g1, setup["deltaABC"][i]))
    for i in [0, 1, len(setup["queryH_scalars"]) - 1]:
        assert O.G1.equals(pk["queryH"][i], O.G1.mul(g1, setup["queryH_scalars"][i]))
    assert O.G1.equals(vk["gammaABCG1"][ni - 1], O.G1.mul(g1, setup["gammaABC"][ni - 1]))
    assert O.G2.equals(vk["gammaG2"], O.G2.mul(g2, setup["gamma"]))
    # prove
    (A, B, C), H = gz.prove(pk, cons, ni, prim, aux)
    n = O.SerialFFT(num_constraints + ni).domain_size
    fft = G.CFFT(n) if n > 4096 else None
    H_exp = G.r1cs_to_qap_witness(o_cons, ni, o_prim, o_aux, fft)
    assert H == H_exp                                             # 7 transforms + pointwise ops, bit-exact
    a, b, c = G.proof_exponents(setup, o_prim, o_aux, H_exp)      # also asserts the verification equation
    assert O.G1.equals(A, O.G1.mul(g1, a))
    assert O.G2.equals(B, O.G2.mul(g2, b))
    assert O.G1.equals(C, O.G1.mul(g1, c))
    # the reference's own end-to-end assertion, Verifier.verify == true (SerialzkSNARKTest.java:63-78), by the oracle's
    # restatement of the pairing verifier, on the proof and the keys the GPU path produced
    from oracle import pairing_oracle as PO
    assert PO.verify(pk["alphaG1"], pk["betaG2"], vk["gammaG2"], vk["deltaG2"], vk["gammaABCG1"], prim, (A, B, C))
    if num_constraints <= 64:
        assert not PO.verify(pk["alphaG1"], pk["betaG2"], vk["gammaG2"], vk["deltaG2"], vk["gammaABCG1"], prim,
                             (A, B, O.G1.add(C, g1)))
        # literal restatement of SerialSetup + SerialProver: identical proof
        _, o_pk, _ = G.setup_literal(o_cons, ni, nv)
        (oA, oB, oC), _ = G.prove_literal(o_pk, o_cons, ni, o_prim, o_aux)
        assert O.G1.equals(A, oA) and O.G2.equals(B, oB) and O.G1.equals(C, oC)
        assert O.G1.to_affine(A) == O.G1.to_affine(oA) and O.G2.to_affine(B) == O.G2.to_affine(oB)   # identical affine points
