"""The Python mirror of the reference's Java operator interface (octopuszk_b200/algebra.py), written the way the
reference's own tests are: SerialVariableBaseMSMTest.java:31-77, SerialFixedBaseMSMTest.java:55-72,
SerialFFTTest.java:168-190, DistributedFFTTest.java (inverse / coset round trips)."""
import random

import pytest

from oracle import dizk_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from octopuszk_b200 import Context
    c = Context(0)
    yield c
    c.close()


def test_serial_variable_base_msm(ctx):
    from octopuszk_b200.algebra import VariableBaseMSM
    msm = VariableBaseMSM(ctx)
    bases = [O.G1.mul(O.G1.generator, k) for k in (5, 2, 7, 3)]
    result = msm.serialMSM([3, 11, 2, 8], bases)
    assert O.G1.equals(result, O.G1.mul(O.G1.generator, 75))
    g2 = [O.G2.mul(O.G2.generator, k) for k in (5, 2, 7, 3)]
    r1, r2 = msm.doubleMSM([3, 11, 2, 8], list(zip(bases, g2)))
    assert O.G1.equals(r1, O.G1.mul(O.G1.generator, 75)) and O.G2.equals(r2, O.G2.mul(O.G2.generator, 75))
    with pytest.raises(ValueError):
        msm.serialMSM([], [])


def test_fixed_base_msm(ctx):
    from octopuszk_b200.algebra import FixedBaseMSM
    fb = FixedBaseMSM(ctx)
    base = O.G1.random(10)
    assert FixedBaseMSM.getWindowSize(1 << 20, base) == 17 and FixedBaseMSM.getWindowSize(1 << 20, O.G2.generator) == 17
    scalar_size = O.G1.bit_size(base)
    window = 2
    rng = random.Random(1)
    scalars = [4294967296, 200] + [rng.randrange(O.R) for _ in range(6)]
    out = fb.batchMSM(scalar_size, window, base, scalars)
    table = O.get_window_table(O.G1, base, scalar_size, window)
    for s, got in zip(scalars, out):
        assert O.G1.equals(got, O.fixed_serial_msm(O.G1, scalar_size, window, table, s))
    assert fb.batchFieldMSM([1, 2, O.R - 1], 7) == [7, 14, O.R - 7]
    pairs = fb.doubleBatchMSM(scalar_size, 5, 254, 5, base, O.G2.generator, scalars[:3])
    for s, (a, b) in zip(scalars, pairs):
        assert O.G1.equals(a, O.G1.mul(base, s)) and O.G2.equals(b, O.G2.mul(O.G2.generator, s))


def test_serial_fft_domain(ctx):
    from octopuszk_b200.algebra import FFTAuxiliary, SerialFFT
    dom = SerialFFT(ctx, 4)
    a = [2, 5, 3, 8]
    dom.radix2FFT(a)
    assert a == [O.naive_evaluate([2, 5, 3, 8], pow(dom.omega, i, O.R)) for i in range(4)]
    rng = random.Random(2)
    m = 64
    dom = SerialFFT(ctx, m)
    ref = O.SerialFFT(m)
    x = [rng.randrange(O.R) for _ in range(m)]
    g = O.FR_MULT_GEN
    for ours, theirs in ((lambda v: dom.radix2InverseFFT(v), lambda v: ref.radix2_inverse_fft(v)),
                         (lambda v: dom.radix2CosetFFT(v, g), lambda v: ref.radix2_coset_fft(v, g)),
                         (lambda v: dom.radix2CosetInverseFFT(v, g), lambda v: ref.radix2_coset_inverse_fft(v, g)),
                         (lambda v: dom.divideByZOnCoset(g, v), lambda v: ref.divide_by_z_on_coset(g, v))):
        a, b = list(x), list(x)
        ours(a)
        theirs(b)
        assert a == b
    a = list(x)
    FFTAuxiliary(ctx).multiplyByCoset(a, g)
    b = list(x)
    O.multiply_by_coset(b, g)
    assert a == b
    a = list(x)
    FFTAuxiliary(ctx).serialRadix2FFT(a, dom.omega)
    b = list(x)
    O.serial_radix2_fft(b, ref.omega)
    assert a == b
