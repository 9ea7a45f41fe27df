"""Pins oracle/dizk_oracle.py against every known-answer test the reference's own unit tests hold for the hot
path and against the constants of SURVEY.md Appendix B."""
import random

from oracle import dizk_oracle as O


def test_java_random_seed10():
    # configuration/Configuration.java:52 seeds with 10; Fp.random = new Random(seed).nextLong() (Fp.java:72-80)
    assert O.JavaRandom(10).next_long() == -4972683369271453960
    assert O.fp_random(10, O.R) == 21888242871839275222246405745257275088548364400416034343693231503206537041657


def test_msm_integer_group_kat():
    # src/test/java/algebra/msm/SerialVariableBaseMSMTest.java:31-77 -> 75
    assert O.naive_msm(O.ZGROUP, [3, 11, 2, 8], [5, 2, 7, 3]) == 75
    assert O.pippenger_msm(O.ZGROUP, [3, 11, 2, 8], [5, 2, 7, 3]) == 75
    # src/test/java/algebra/msm/DistributedVariableBaseMSMTest.java:92-124 -> 60 (duplicates)
    assert O.pippenger_msm(O.ZGROUP, [3] * 4, [5] * 4) == 60


def test_msm_g1_kat_transfer():
    # the same vector carried to BN254 G1: [5G,2G,7G,3G] -> 75 G
    bases = [O.G1.mul(O.G1.generator, k) for k in (5, 2, 7, 3)]
    r = O.G1.to_affine(O.pippenger_msm(O.G1, [3, 11, 2, 8], bases))
    assert r[:2] == (14670023805213312856584033961079180710026959676164645964476657106778352781859,
                     211633134735504671946091929992244044834074118928621612299531666035417451988)
    assert O.G1.equals(O.naive_msm(O.G1, [3, 11, 2, 8], bases), O.G1.mul(O.G1.generator, 75))


def test_fft_kat_large_fp():
    # src/test/java/algebra/fft/SerialFFTTest.java:168-190: FFT([2,5,3,8]) == naive evaluation at omega^i
    dom = O.SerialFFT(4, O.LARGE_FP_MODULUS, O.LARGE_FP_ROOT)
    a = [2, 5, 3, 8]
    dom.radix2_fft(a)
    assert a == [O.naive_evaluate([2, 5, 3, 8], pow(dom.omega, i, O.LARGE_FP_MODULUS), O.LARGE_FP_MODULUS)
                 for i in range(4)]
    assert a == [18, 863244948084361541713705916379447741823013005020402232,
                 1532495540865888858358347027150309183618765510462668793,
                 669250592781527316644641110770861441795752505442266567]


def test_fft_kat_fr():
    a = [2, 5, 3, 8]
    O.SerialFFT(4).radix2_fft(a)
    assert a == [18, 21888242871839275209022642834368543560924422484752198131886912786320552141471,
                 21888242871839275222246405745257275088548364400416034343698204186575808495609,
                 13223762910888731527623941915663836211811291400255256354144]


def test_root_of_unity():
    # src/test/java/algebra/curves/BNFieldsTest.java:74
    w = O.root_of_unity(8)
    assert pow(w, 8, O.R) == 1 and pow(w, 4, O.R) != 1
    assert pow(O.FR_ROOT, 1 << 28, O.R) == 1 and pow(O.FR_ROOT, 1 << 27, O.R) == O.R - 1


def test_group_laws():
    # src/test/java/algebra/curves/CurvesTest.java:62-81
    rng = random.Random(3)
    for G in (O.G1, O.G2):
        a = G.mul(G.generator, rng.randrange(O.R))
        r1, r2 = rng.randrange(O.R), rng.randrange(O.R)
        assert G.equals(G.twice(a), G.add(a, a))
        assert G.is_zero(G.sub(a, a))
        assert G.equals(G.add(G.mul(a, r1), G.mul(a, r2)), G.mul(a, (r1 + r2) % O.R))
        assert G.is_zero(G.mul(G.generator, O.R))


def test_seed10_generators():
    g = O.G1.random(10)
    assert g == (4917613592469748221377412155284963277687977530704438532678615470445113848758,
                 10213824969900307996792374339763350851485120556585809505860694351865747169886,
                 7394481779713641977695769229485301335956112496977197210873819694116977828251)
    assert O.G1.bit_size(g) == 253
    assert O.G2.bit_size(O.G2.random(10)) == 254


def test_window_rules():
    assert [O.pippenger_window(n) for n in (1023, 1 << 16, 1 << 20, 1 << 23, 1 << 24, 1 << 26)] == [6, 11, 14, 16, 16, 18]
    assert O.get_window_size(1 << 20, O.G1) == 17 and O.get_window_size(1 << 24, O.G1) == 20
    assert O.get_window_size(57818, O.G1) == 13 and O.get_window_size(34552892, O.G1) == 22


def test_fixed_base_table_walk():
    # FixedBaseMSM.getWindowTable + serialMSM (the pure-Java oracle) == scalar multiplication
    rng = random.Random(5)
    for G, ss in ((O.G1, 253), (O.G2, 254)):
        w = 3
        base = G.random(10)
        table = O.get_window_table(G, base, ss, w)
        s = rng.randrange(O.R)
        outerc = (ss + w - 1) // w
        got = O.fixed_serial_msm(G, ss, w, table, s)
        assert G.equals(got, G.mul(base, s & ((1 << (outerc * w)) - 1)))
        assert G.equals(got, O.fixed_batch_msm(G, ss, w, base, [s])[0])


def test_distributed_fft_equals_serial():
    # src/test/java/algebra/fft/DistributedFFTTest.java:41-67 and the inverse/coset variants
    rng = random.Random(1)
    x = [rng.randrange(O.R) for _ in range(64)]
    y = list(x)
    O.SerialFFT(64).radix2_fft(y)
    for rows, cols in ((8, 8), (4, 16), (16, 4), (1, 64), (64, 1)):
        assert O.distributed_radix2_fft(x, rows, cols, False) == y
    y = list(x)
    O.SerialFFT(64).radix2_inverse_fft(y)
    assert O.distributed_radix2_fft(x, 16, 4, True) == y
    z = list(y)
    O.SerialFFT(64).radix2_fft(z)
    assert z == x


def test_coset_roundtrip_and_divide_by_z():
    rng = random.Random(2)
    dom = O.SerialFFT(32)
    x = [rng.randrange(O.R) for _ in range(32)]
    y = list(x)
    dom.radix2_coset_fft(y, O.FR_MULT_GEN)
    assert y == [O.naive_evaluate(x, O.FR_MULT_GEN * pow(dom.omega, i, O.R) % O.R) for i in range(32)]
    dom.radix2_coset_inverse_fft(y, O.FR_MULT_GEN)
    assert y == x


def test_wire_format_roundtrip():
    pts = [O.G1.random(10), O.G1.zero()]
    assert O.unpack_g1(O.pack_g1(pts)) == pts
    q = [O.G2.random(10)]
    assert O.unpack_g2(O.pack_g2(q)) == q
    assert O.le32(1) == b"\x01" + b"\x00" * 31
