"""GPU parity: fixed-base batch MSM (G1, G2) through the C ABI against the oracle's restatement of
FixedBaseMSM.getWindowTable + serialMSM (what batchMSM must return as group elements)."""
import random

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import dizk_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from octopuszk_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _check(G, ctx, base, scalars, scalar_size, w):
    outerc = (scalar_size + w - 1) // w
    n = len(scalars)
    if G is O.G1:
        out = O.unpack_g1(ctx.fixed_g1(O.pack_g1([base]), O.pack_scalars(scalars), n, outerc, w))
    else:
        out = O.unpack_g2(ctx.fixed_g2(O.pack_g2([base]), O.pack_scalars(scalars), n, outerc, w))
    exp = O.fixed_batch_msm(G, scalar_size, w, base, scalars)
    assert len(out) == n
    for got, e in zip(out, exp):
        assert G.equals(got, e)
        # normalised output: Z == 1 or the reference's infinity (0, 1, 0)
        if G.is_zero(e):
            assert got == (G.F.zero, G.F.one, G.F.zero)
        else:
            assert got[2] == G.F.one


@pytest.mark.parametrize("scalar_size,w", [(253, 13), (253, 17), (253, 20), (253, 11), (253, 1), (254, 16), (10, 3)])
def test_g1_matches_oracle(ctx, scalar_size, w):
    rng = random.Random(scalar_size * 100 + w)
    base = O.G1.random(10)                       # seed-10 generator, Jacobian with Z != 1
    scalars = [0, 1, 2, O.R - 1, O.R - 2, (1 << 253) + 5, (1 << 253) - 1] + [rng.randrange(O.R) for _ in range(60)]
    _check(O.G1, ctx, base, scalars, scalar_size, w)


def test_g1_infinity_base_and_zero_windows(ctx):
    scalars = [5, 7, O.R - 1]
    out = O.unpack_g1(ctx.fixed_g1(O.pack_g1([O.G1.zero()]), O.pack_scalars(scalars), 3, 20, 13))
    assert all(O.G1.is_zero(p) for p in out)
    out = O.unpack_g1(ctx.fixed_g1(O.pack_g1([O.G1.generator]), O.pack_scalars(scalars), 3, 0, 13))
    assert all(O.G1.is_zero(p) for p in out)


@pytest.mark.parametrize("scalar_size,w", [(254, 5), (254, 16), (254, 19)])
def test_g2_matches_oracle(ctx, scalar_size, w):
    rng = random.Random(scalar_size * 100 + w)
    base = O.G2.random(10)
    scalars = [0, 1, O.R - 1] + [rng.randrange(O.R) for _ in range(25)]
    _check(O.G2, ctx, base, scalars, scalar_size, w)


def test_g1_large_vs_c_oracle(ctx):
    """2^14 scalars against the C restatement of getWindowTable + serialMSM, every point compared after affine
    normalisation; plus the checksum sum_i out_i == (sum_i s_i) * B over 2^18 scalars."""
    n = 1 << 14
    base = O.pack_g1([O.G1.random(10)])
    raw = util.rand_scalars_bytes(n, seed=21)
    w, ss = 13, 253
    outerc = (ss + w - 1) // w
    got = ctx.fixed_g1(base, raw.tobytes(), n, outerc, w)
    exp = C.fixed_g1(base, raw.tobytes(), n, ss, w, C.max_threads())
    for i in range(0, n, 1):
        a, b = got[96 * i:96 * i + 96], exp[96 * i:96 * i + 96]
        if a != C.g1_to_affine(b):
            assert C.g1_equal(a, b), i
    n = 1 << 18
    raw = util.rand_scalars_bytes(n, seed=22)
    got = ctx.fixed_g1(base, raw.tobytes(), n, (253 + 16) // 17, 17)
    ones = np.zeros((n, 32), dtype=np.uint8)
    ones[:, 0] = 1
    total = ctx.msm_g1(ones.tobytes(), got, n)
    s = util.column_sums(raw, 1)[0] % O.R
    assert O.G1.equals(O.unpack_g1(total)[0], O.G1.mul(O.G1.random(10), s))


def test_rejects_bad_arguments(ctx):
    from octopuszk_b200 import OzkError
    base = O.pack_g1([O.G1.generator])
    with pytest.raises(OzkError):
        ctx.fixed_g1(base, O.le32(O.R), 1, 20, 13)
    with pytest.raises(OzkError):
        ctx.fixed_g1(base, O.le32(1), 1, 20, 0)
