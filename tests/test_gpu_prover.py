"""Device-resident Groth16 setup + prover (octopuszk_b200/prover.py) against the oracle: BASELINE.json configs[0] (2^15 constraints)
and the first sizes of configs[4], on one GPU.  The proof must be the points (a g1, b g2, c g1) whose exponents the C oracle
derives from the QAP at the setup point (oracle/groth16_oracle.expected_proof_synthetic, itself pinned on the literal restatement
of SerialSetup + SerialProver in tests/test_oracle_groth16.py), i.e. identical affine proof points; at the small sizes the proving
key, the H polynomial and the pairing verifier are checked too."""
import random

import numpy as np
import pytest
import torch

from oracle import dizk_oracle as O
from oracle import groth16_oracle as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from octopuszk_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _z(prim, aux):
    return torch.frombuffer(bytearray(O.pack_scalars(list(prim) + list(aux))), dtype=torch.uint8).view(-1, 32).cuda()


def _rows(csr):
    """CSR on the device -> list of sorted column lists"""
    ptr = csr.row_ptr.cpu().tolist()
    col = csr.col.cpu().tolist()
    return [sorted(col[ptr[i]:ptr[i + 1]]) for i in range(csr.rows)]


@pytest.mark.parametrize("nc,ni", [(8, 3), (9, 4), (64, 7)])
def test_synthetic_rows_match_the_reference_generator(ctx, nc, ni):
    """synthetic_r1cs == R1CSConstruction.serialConstruct restated (oracle), whole and as cyclic shards of 2 and 4 ranks."""
    from octopuszk_b200.prover import synthetic_r1cs
    cons, _, _, _, _ = G.serial_construct(nc, ni)
    for world in (1, 2, 4):
        for rank in range(world):
            r = synthetic_r1cs(nc, ni, torch.device("cuda"), world, rank)
            want = cons[rank::world]
            for k, M in enumerate((r.A, r.B, r.C)):
                assert M.coeff is None and M.cols == nc + 3
                assert _rows(M) == [sorted(i for i, _ in c[k]) for c in want]


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("nc,ni", [(64, 7), (1 << 15, 1023)])
def test_device_prover_matches_oracle(ctx, nc, ni):
    from octopuszk_b200.prover import DeviceGroth16, Csr, R1cs, synthetic_r1cs
    gz = DeviceGroth16(ctx)
    cons, _, na, prim, aux = G.serial_construct(nc, ni)
    nv = ni + na
    full = synthetic_r1cs(nc, ni, torch.device("cuda"))
    pk, vk = gz.setup(full)
    setup = G.setup_scalars(cons, ni, nv)
    assert pk.info["windowSizeG1"] == setup["windowSizeG1"] and pk.info["windowSizeG2"] == setup["windowSizeG2"]
    g1, g2 = setup["g1"], setup["g2"]
    # proving-key entries through unit-vector MSMs on the resident keys
    rng = random.Random(3)

    def entry(key, i, n, g2_=False):
        e = np.zeros((n, 32), dtype=np.uint8)
        e[i, 0] = 1
        out = ctx.msm_keyed(torch.from_numpy(e).cuda(), key, n, device=True)
        return O.unpack_g2(out)[0] if g2_ else O.unpack_g1(out)[0]

    for i in [0, 1, ni - 1, ni, nv - 1] + [rng.randrange(nv) for _ in range(3)]:
        assert O.G1.equals(entry(pk.queryA, i, nv), O.G1.mul(g1, setup["qap"]["At"][i]))
        assert O.G1.equals(entry(pk.queryB1, i, nv), O.G1.mul(g1, setup["qap"]["Bt"][i]))
    for i in [0, nv - 1]:
        assert O.G2.equals(entry(pk.queryB2, i, nv, True), O.G2.mul(g2, setup["qap"]["Bt"][i]))
    n = pk.info["domain"]
    for i in [0, 1, n - 1]:
        assert O.G1.equals(entry(pk.queryH, i, n), O.G1.mul(g1, setup["queryH_scalars"][i]))
    for i in [0, nv - ni - 1]:
        assert O.G1.equals(entry(pk.deltaABC, i, nv - ni), O.G1.mul(g1, setup["deltaABC"][i]))
    assert O.G1.equals(vk["gammaABCG1"][ni - 1], O.G1.mul(g1, setup["gammaABC"][ni - 1]))
    # the witness polynomial, bit for bit
    d_z = _z(prim, aux)
    H = gz.witness_map(full, d_z).cpu().numpy().tobytes()
    H_exp = G.r1cs_to_qap_witness(cons, ni, prim, aux, G.CFFT(n) if n > 4096 else None)
    assert [O.from_le(H[32 * i:32 * i + 32]) for i in range(n)] == H_exp[:n] and H_exp[n] == 0
    # the proof
    A, B, C = gz.prove(pk, full, d_z)
    eA, eB, eC = G.expected_proof_synthetic(nc, ni)
    assert (O.G1.to_affine(A), O.G2.to_affine(B), O.G1.to_affine(C)) == (eA, eB, eC)
    a, b, c = G.proof_exponents(setup, prim, aux, H_exp)
    assert O.G1.equals(A, O.G1.mul(g1, a)) and O.G2.equals(B, O.G2.mul(g2, b)) and O.G1.equals(C, O.G1.mul(g1, c))
    # the reference's own end-to-end assertion: Verifier.verify == true (pairing oracle)
    from oracle import pairing_oracle as PO
    assert PO.verify(pk.alphaG1, pk.betaG2, vk["gammaG2"], vk["deltaG2"], vk["gammaABCG1"], prim, (A, B, C))
    if nc <= 64:
        assert not PO.verify(pk.alphaG1, pk.betaG2, vk["gammaG2"], vk["deltaG2"], vk["gammaABCG1"], prim, (A, B, O.G1.add(C, g1)))
        # a general (non-unit) circuit goes through Csr.from_rows: scale every row of A by 3 and C by 3 -> still satisfied
        cons3 = [([(i, 3 * v % O.R) for i, v in a_], b_, [(i, 3 * v % O.R) for i, v in c_]) for a_, b_, c_ in cons]
        dev = torch.device("cuda")
        gen = R1cs(Csr.from_rows([c_[0] for c_ in cons3], nv, dev), Csr.from_rows([c_[1] for c_ in cons3], nv, dev),
                   Csr.from_rows([c_[2] for c_ in cons3], nv, dev), nc, ni, nv)
        H3 = gz.witness_map(gen, d_z).cpu().numpy().tobytes()
        H3_exp = G.r1cs_to_qap_witness(cons3, ni, prim, aux)
        assert [O.from_le(H3[32 * i:32 * i + 32]) for i in range(n)] == H3_exp[:n]
    pk.free()


@pytest.mark.timeout(1800)
def test_device_prover_2_20_constraints(ctx):
    """2^20 constraints (domain 2^21): sizes only the C oracle reaches; identical affine proof points."""
    from octopuszk_b200.groth16 import fr_random
    from octopuszk_b200.prover import DeviceGroth16, synthetic_r1cs
    from oracle import c_oracle as C
    nc, ni = 1 << 20, 1023
    gz = DeviceGroth16(ctx)
    full = synthetic_r1cs(nc, ni, torch.device("cuda"))
    pk, _ = gz.setup(full, keep_vk=False)
    d_z = torch.from_numpy(C.r1cs_chain(nc, fr_random(), fr_random())).cuda()
    A, B, Cp = gz.prove(pk, full, d_z)
    assert (O.G1.to_affine(A), O.G2.to_affine(B), O.G1.to_affine(Cp)) == G.expected_proof_synthetic(nc, ni)
    pk.free()
