"""The C restatement (oracle/dizk_oracle.c) must agree with the Python oracle, which is what the reference's own
known-answer tests pin (tests/test_oracle.py)."""
import random

from oracle import c_oracle as C
from oracle import dizk_oracle as O
from tests import util


def test_msm_kat_and_random():
    bases = [O.G1.mul(O.G1.generator, k) for k in (5, 2, 7, 3)]
    out = C.msm_g1(O.pack_scalars([3, 11, 2, 8]), O.pack_g1(bases), 4)
    aff = O.unpack_g1(C.g1_to_affine(out))[0]
    assert aff == (14670023805213312856584033961079180710026959676164645964476657106778352781859,
                   211633134735504671946091929992244044834074118928621612299531666035417451988, 1)
    rng = random.Random(2)
    ks, pool = util.known_dlog_points(O.G1, 10, seed=2)
    for n in (1, 2, 33, 200):
        pts = [pool[rng.randrange(10)] for _ in range(n)]
        pts[0] = O.G1.zero()
        sc = [rng.randrange(O.R) for _ in range(n)]
        sc[-1] = 0
        exp = O.pippenger_msm(O.G1, sc, pts)
        for threads in (1, 3):
            got = O.unpack_g1(C.msm_g1(O.pack_scalars(sc), O.pack_g1(pts), n, threads))[0]
            assert O.G1.equals(got, exp)
        if n == 33:
            # single thread follows VariableBaseMSM.pippengerMSM operation for operation: identical Jacobian triple
            assert O.unpack_g1(C.msm_g1(O.pack_scalars(sc), O.pack_g1(pts), n, 1))[0] == exp


def test_msm_large_known_dlog():
    n = 1 << 14
    ks, pool = util.known_dlog_points(O.G1, 64, seed=14)
    raw = util.rand_scalars_bytes(n, seed=14)
    bases = util.tiled_bases_bytes(O.G1, pool, n)
    out = C.msm_g1(raw.tobytes(), bases.tobytes(), n, C.max_threads())
    exp = util.expected_from_dlogs(O.G1, ks, util.column_sums(raw, 64))
    assert C.g1_equal(out, O.pack_g1([exp]))


def test_fft_matches_python():
    rng = random.Random(3)
    for log_n in (1, 2, 5, 10):
        n = 1 << log_n
        x = [rng.randrange(O.R) for _ in range(n)]
        omega = O.root_of_unity(n)
        exp = list(x)
        O.serial_radix2_fft(exp, omega)
        got = C.fft_fr(O.pack_scalars(x), O.le32(omega))
        assert [O.from_le(got[i:i + 32]) for i in range(0, len(got), 32)] == exp


def test_fixed_base_matches_python():
    rng = random.Random(4)
    base = O.G1.random(10)
    sc = [0, 1, O.R - 1] + [rng.randrange(O.R) for _ in range(20)]
    for ss, w in ((253, 5), (253, 11), (254, 4)):
        table = O.get_window_table(O.G1, base, ss, w)
        out = O.unpack_g1(C.fixed_g1(O.pack_g1([base]), O.pack_scalars(sc), len(sc), ss, w, 2))
        for s, got in zip(sc, out):
            assert got == O.fixed_serial_msm(O.G1, ss, w, table, s)      # same operations, same triple


def test_fr_scale_matches_python():
    rng = random.Random(5)
    a = [0, 1, O.R - 1] + [rng.randrange(O.R) for _ in range(50)]
    b = rng.randrange(O.R)
    got = C.fr_scale(O.pack_scalars(a), O.le32(b))
    assert [O.from_le(got[i:i + 32]) for i in range(0, len(got), 32)] == O.field_batch_msm(a, b)
