"""Golden vectors recorded from the reference itself: tests/golden/reference_cuda_vectors.json holds seeded inputs and the raw
output bytes of the reference's own CUDA sources (compiled unmodified, called through their JNI entry points on a B200 by
tests/golden/make_reference_cuda_vectors.py).  Without a GPU the oracle must reproduce every vector (this pins the oracle on
outputs of the reference itself); with one, liboctozk must."""
import json
import os

import pytest

from oracle import dizk_oracle as O

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cuda_vectors.json")


def _cases():
    with open(PATH) as f:
        return json.load(f)["cases"]


def _scalars(h):
    b = bytes.fromhex(h)
    return [O.from_le(b[i:i + 32]) for i in range(0, len(b), 32)]


def test_fixture_present_and_complete():
    kinds = [c["kind"] for c in _cases()]
    assert kinds.count("var_msm_g1") == 3 and {"var_msm_g2", "var_double_msm", "fixed_batch_g1", "fixed_batch_g2", "fixed_double_batch",
                                               "field_batch"} <= set(kinds)


def test_oracle_reproduces_reference_outputs():
    for c in _cases():
        out = bytes.fromhex(c["out"])
        sc = _scalars(c["scalars"])
        if c["kind"] == "var_msm_g1":
            bases = O.unpack_g1(bytes.fromhex(c["bases"]))
            assert O.G1.equals(O.unpack_g1(out, stride=64)[0], O.pippenger_msm(O.G1, sc, bases))
            if c["n"] == 4:                                # SerialVariableBaseMSMTest.java:31-77 carried to G1: 75 G
                assert O.G1.equals(O.unpack_g1(out, stride=64)[0], O.G1.mul(O.G1.generator, 75))
        elif c["kind"] == "var_msm_g2":
            bases = O.unpack_g2(bytes.fromhex(c["bases"]))
            assert O.G2.equals(O.unpack_g2(out, stride=64)[0], O.pippenger_msm(O.G2, sc, bases))
        elif c["kind"] == "var_double_msm":
            e1, e2 = O.double_msm(sc, O.unpack_g1(bytes.fromhex(c["bases1"])), O.unpack_g2(bytes.fromhex(c["bases2"])))
            assert O.G1.equals(O.unpack_g1(out[:192], stride=64)[0], e1)
            assert O.G2.equals(O.unpack_g2(out[192:], stride=64)[0], e2)
        elif c["kind"] == "fixed_batch_g1":
            base = O.unpack_g1(bytes.fromhex(c["base"]))[0]
            exp = O.fixed_batch_msm(O.G1, c["scalar_size"], c["window"], base, sc)
            got = O.unpack_g1(out, stride=64, big_endian=True)
            assert len(got) == c["n"] and all(O.G1.equals(a, b) for a, b in zip(got, exp))
        elif c["kind"] == "fixed_batch_g2":
            base = O.unpack_g2(bytes.fromhex(c["base"]))[0]
            exp = O.fixed_batch_msm(O.G2, c["scalar_size"], c["window"], base, sc)
            got = O.unpack_g2(out, stride=64, big_endian=True)
            assert len(got) == c["n"] and all(O.G2.equals(a, b) for a, b in zip(got, exp))
        elif c["kind"] == "fixed_double_batch":
            b1, b2 = O.unpack_g1(bytes.fromhex(c["base1"]))[0], O.unpack_g2(bytes.fromhex(c["base2"]))[0]
            e1 = O.fixed_batch_msm(O.G1, c["scalar_size1"], c["window1"], b1, sc)
            e2 = O.fixed_batch_msm(O.G2, c["scalar_size2"], c["window2"], b2, sc)
            for i in range(c["n"]):                     # element i = G1 (192 B) || G2 (384 B), SURVEY.md Appendix A.2
                blk = out[576 * i:576 * (i + 1)]
                assert O.G1.equals(O.unpack_g1(blk[:192], stride=64, big_endian=True)[0], e1[i])
                assert O.G2.equals(O.unpack_g2(blk[192:], stride=64, big_endian=True)[0], e2[i])
        elif c["kind"] == "field_batch":
            b = O.from_le(bytes.fromhex(c["b"]))
            assert [int.from_bytes(out[64 * i:64 * i + 64], "big") for i in range(c["n"])] == O.field_batch_msm(sc, b)
        else:
            raise AssertionError(c["kind"])


@pytest.mark.gpu
def test_liboctozk_reproduces_reference_outputs():
    from octopuszk_b200 import Context
    ctx = Context(0)
    for c in _cases():
        out = bytes.fromhex(c["out"])
        sb = bytes.fromhex(c["scalars"])
        n = c["n"]
        if c["kind"] == "var_msm_g1":
            assert O.G1.equals(O.unpack_g1(ctx.msm_g1(sb, bytes.fromhex(c["bases"]), n))[0], O.unpack_g1(out, stride=64)[0])
        elif c["kind"] == "var_msm_g2":
            assert O.G2.equals(O.unpack_g2(ctx.msm_g2(sb, bytes.fromhex(c["bases"]), n))[0], O.unpack_g2(out, stride=64)[0])
        elif c["kind"] == "var_double_msm":
            got = ctx.msm_g1g2(sb, bytes.fromhex(c["bases1"]), bytes.fromhex(c["bases2"]), n)
            assert O.G1.equals(O.unpack_g1(got[:96])[0], O.unpack_g1(out[:192], stride=64)[0])
            assert O.G2.equals(O.unpack_g2(got[96:])[0], O.unpack_g2(out[192:], stride=64)[0])
        elif c["kind"] == "fixed_batch_g1":
            got = O.unpack_g1(ctx.fixed_g1(bytes.fromhex(c["base"]), sb, n, c["outerc"], c["window"]))
            ref = O.unpack_g1(out, stride=64, big_endian=True)
            assert all(O.G1.equals(a, b) for a, b in zip(got, ref))
        elif c["kind"] == "fixed_batch_g2":
            got = O.unpack_g2(ctx.fixed_g2(bytes.fromhex(c["base"]), sb, n, c["outerc"], c["window"]))
            ref = O.unpack_g2(out, stride=64, big_endian=True)
            assert len(got) == n and all(O.G2.equals(a, b) for a, b in zip(got, ref))
        elif c["kind"] == "fixed_double_batch":
            g1 = O.unpack_g1(ctx.fixed_g1(bytes.fromhex(c["base1"]), sb, n, c["outerc1"], c["window1"]))
            g2 = O.unpack_g2(ctx.fixed_g2(bytes.fromhex(c["base2"]), sb, n, c["outerc2"], c["window2"]))
            for i in range(n):
                blk = out[576 * i:576 * (i + 1)]
                assert O.G1.equals(g1[i], O.unpack_g1(blk[:192], stride=64, big_endian=True)[0])
                assert O.G2.equals(g2[i], O.unpack_g2(blk[192:], stride=64, big_endian=True)[0])
        elif c["kind"] == "field_batch":
            got = ctx.fr_scale(sb, bytes.fromhex(c["b"]))
            assert [O.from_le(got[32 * i:32 * i + 32]) for i in range(n)] == [int.from_bytes(out[64 * i:64 * i + 64], "big") for i in range(n)]
    ctx.close()
