"""GPU parity through the JNI boundary: the Java_* symbols of the three shim libraries are called with the byte
layouts the reference's Java produces via a fake JNIEnv, and their returned arrays are parsed
exactly as the Java callers parse them (VariableBaseMSM.java:239-258,293-327,541-589; FixedBaseMSM.java:233-246,
560-589,771-777; FFTAuxiliary.java:83-93)."""
import ctypes
import random

import pytest

from oracle import dizk_oracle as O
from tests import util
from tests.jni_harness import FakeJvm

pytestmark = pytest.mark.gpu
vp, i32 = ctypes.c_void_p, ctypes.c_int32


@pytest.fixture(scope="module")
def jvm():
    return FakeJvm()


def test_variable_base_serial_and_double(jvm):
    rng = random.Random(31)
    n = 200
    k1, p1 = util.known_dlog_points(O.G1, 8, seed=1)
    k2, p2 = util.known_dlog_points(O.G2, 8, seed=2)
    b1 = [p1[i % 8] for i in range(n)]
    b2 = [p2[i % 8] for i in range(n)]
    sc = [rng.randrange(O.R) for _ in range(n)]
    f = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_variableBaseSerialMSMNativeHelper", [vp, vp, i32, i32, i32])
    r = f(jvm.env, None, jvm.bytes_(O.pack_g1(b1)), jvm.bytes_(O.pack_scalars(sc)), n, 1, 0)
    assert r and jvm.exception() is None
    out = jvm.read(r)
    assert len(out) == 192                                   # X|Y|Z, 64 bytes little-endian each
    got = O.unpack_g1(out, stride=64)[0]
    assert all(v < O.P for v in got)
    e1, e2 = O.double_msm(sc, b1, b2)
    assert O.G1.equals(got, e1)
    r = f(jvm.env, None, jvm.bytes_(O.pack_g2(b2)), jvm.bytes_(O.pack_scalars(sc)), n, 2, 5)     # taskID % nGPU
    out = jvm.read(r)
    assert len(out) == 384
    assert O.G2.equals(O.unpack_g2(out, stride=64)[0], e2)
    g = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_variableBaseDoubleMSMNativeHelper", [vp, vp, vp, i32, i32])
    r = g(jvm.env, None, jvm.bytes_(O.pack_g1(b1)), jvm.bytes_(O.pack_g2(b2)), jvm.bytes_(O.pack_scalars(sc)), n, 0)
    out = jvm.read(r)
    assert len(out) == 576
    assert O.G1.equals(O.unpack_g1(out[:192], stride=64)[0], e1)
    assert O.G2.equals(O.unpack_g2(out[192:], stride=64)[0], e2)
    assert jvm.live_pins() == 0                              # every pinned array was released
    # direct-buffer variant: 32-byte elements in and out
    d = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_variableBaseMSMDirect", [vp, vp, vp, i32, i32, i32, vp], i32)
    ob = jvm.direct(size=288)
    rc = d(jvm.env, None, jvm.direct(O.pack_g1(b1)), jvm.direct(O.pack_g2(b2)), jvm.direct(O.pack_scalars(sc)), n, 3, 0, ob)
    assert rc == 0
    out = jvm.read(ob)
    assert O.G1.equals(O.unpack_g1(out[:96])[0], e1) and O.G2.equals(O.unpack_g2(out[96:])[0], e2)
    # persistent proving-key handles: upload once, then only the scalars are passed
    i64 = ctypes.c_int64
    up = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_uploadBasesDirect", [vp, i32, i32, i32], i64)
    keyed = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_variableBaseMSMKeyedDirect",
                   [i64, i64, vp, i32, i32, i32, i32, vp], i32)
    free = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_freeBases", [i64, i32], None)
    k1 = up(jvm.env, None, jvm.direct(O.pack_g1(b1)), n, 1, 0)
    k2 = up(jvm.env, None, jvm.direct(O.pack_g2(b2)), n, 2, 0)
    assert k1 and k2
    ob = jvm.direct(size=288)
    assert keyed(jvm.env, None, k1, k2, jvm.direct(O.pack_scalars(sc)), 0, n, 3, 0, ob) == 0
    out = jvm.read(ob)
    assert O.G1.equals(O.unpack_g1(out[:96])[0], e1) and O.G2.equals(O.unpack_g2(out[96:])[0], e2)
    ob = jvm.direct(size=96)
    assert keyed(jvm.env, None, k1, 0, jvm.direct(O.pack_scalars(sc[:5])), 3, 5, 1, 0, ob) == 0
    assert O.G1.equals(O.unpack_g1(jvm.read(ob))[0], O.pippenger_msm(O.G1, sc[:5], b1[3:8]))
    assert keyed(jvm.env, None, k1, 0, jvm.direct(O.pack_scalars(sc)), 1, n, 1, 0, ob) == -1       # range exceeds the key
    assert "RuntimeException" in jvm.exception()
    jvm.clear()
    free(jvm.env, None, k1, 0)
    free(jvm.env, None, k2, 0)


def test_errors_become_runtime_exceptions(jvm):
    f = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_variableBaseSerialMSMNativeHelper", [vp, vp, i32, i32, i32])
    r = f(jvm.env, None, jvm.bytes_(b"\0" * 96), jvm.bytes_(O.le32(1)), 2, 1, 0)       # arrays shorter than batch_size
    assert not r and "RuntimeException" in jvm.exception()
    jvm.clear()
    r = f(jvm.env, None, jvm.bytes_(O.pack_g1([O.G1.generator])), jvm.bytes_(O.le32(O.R)), 1, 1, 0)    # scalar not reduced
    assert not r and "not reduced" in jvm.exception()
    jvm.clear()
    assert jvm.live_pins() == 0


def test_fixed_base_batch_double_and_field(jvm):
    rng = random.Random(32)
    n = 40
    sc = [0, 1, O.R - 1] + [rng.randrange(O.R) for _ in range(n - 3)]
    g1, g2 = O.G1.random(10), O.G2.random(10)
    ss1, ss2 = O.G1.bit_size(g1), O.G2.bit_size(g2)            # 253 / 254, SURVEY.md Appendix C.2
    w1, w2 = 13, 12
    oc1, oc2 = (ss1 + w1 - 1) // w1, (ss2 + w2 - 1) // w2
    f = jvm.fn("libAlgebraMSMFixedBaseMSM.so", "Java_algebra_msm_FixedBaseMSM_batchMSMNativeHelper",
               [i32, i32, i32, i32, i32, i32, vp, vp, i32, i32])
    r = f(jvm.env, None, oc1, w1, oc1, 1 << w1, n, ss1, jvm.bytes_(O.pack_g1([g1])), jvm.bytes_(O.pack_scalars(sc)), 1, 0)
    out = jvm.read(r)
    assert len(out) == n * 192                               # 64-byte BIG-endian coordinates
    e1 = O.fixed_batch_msm(O.G1, ss1, w1, g1, sc)
    for got, e in zip(O.unpack_g1(out, stride=64, big_endian=True), e1):
        assert O.G1.equals(got, e) and all(v < O.P for v in got)
    r = f(jvm.env, None, oc2, w2, oc2, 1 << w2, n, ss2, jvm.bytes_(O.pack_g2([g2])), jvm.bytes_(O.pack_scalars(sc)), 2, 0)
    out = jvm.read(r)
    assert len(out) == n * 384
    e2 = O.fixed_batch_msm(O.G2, ss2, w2, g2, sc)
    for got, e in zip(O.unpack_g2(out, stride=64, big_endian=True), e2):
        assert O.G2.equals(got, e)
    g = jvm.fn("libAlgebraMSMFixedBaseMSM.so", "Java_algebra_msm_FixedBaseMSM_doubleBatchMSMNativeHelper",
               [i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32])
    r = g(jvm.env, None, oc1, w1, oc2, w2, oc1, 1 << w1, oc2, 1 << w2, n, jvm.bytes_(O.pack_g1([g1])), jvm.bytes_(O.pack_g2([g2])),
          jvm.bytes_(O.pack_scalars(sc)), 0)
    out = jvm.read(r)
    assert len(out) == n * 576
    for i in range(n):
        rec = out[576 * i:576 * (i + 1)]
        assert O.G1.equals(O.unpack_g1(rec[:192], stride=64, big_endian=True)[0], e1[i])
        assert O.G2.equals(O.unpack_g2(rec[192:], stride=64, big_endian=True)[0], e2[i])
    # field batch: N scalars followed by the multiplier (FixedBaseMSM.java:761-767), N x 64 bytes big-endian back
    h = jvm.fn("libAlgebraMSMFixedBaseMSM.so", "Java_algebra_msm_FixedBaseMSM_fieldBatchMSMNativeHelper", [vp, i32, i32])
    b = rng.randrange(O.R)
    r = h(jvm.env, None, jvm.bytes_(O.pack_scalars(sc) + O.le32(b)), n, 0)
    out = jvm.read(r)
    assert [int.from_bytes(out[64 * i:64 * i + 64], "big") for i in range(n)] == O.field_batch_msm(sc, b)
    assert jvm.live_pins() == 0


def test_fft_list_of_byte_arrays_and_direct(jvm):
    rng = random.Random(33)
    n = 256
    x = [rng.randrange(O.R) for _ in range(n)]
    x[0], x[1] = 5, 0                                        # short arrays: 4-byte multiples (FFTAuxiliary.java:41-51)

    def cgbn_bytes(v):                                       # bigIntegerToByteArrayHelperCGBN of FFTAuxiliary.java
        raw = v.to_bytes((v.bit_length() + 8) // 8, "big")   # BigInteger.toByteArray: minimal two's complement
        res = bytearray((len(raw) + 3) // 4 * 4)
        for i, byte in enumerate(raw):
            res[len(raw) - i - 1] = byte
        return bytes(res)

    omega = O.root_of_unity(n)
    f = jvm.fn("libAlgebraFFTAuxiliary.so", "Java_algebra_fft_FFTAuxiliary_serialRadix2FFTNativeHelper", [vp, vp, i32])
    r = f(jvm.env, None, jvm.list_([cgbn_bytes(v) for v in x]), jvm.bytes_(cgbn_bytes(omega)), 0)
    assert r and jvm.exception() is None
    out = jvm.read(r)
    assert len(out) == n * 64
    exp = list(x)
    O.serial_radix2_fft(exp, omega)
    assert [int.from_bytes(out[64 * i:64 * i + 64], "little") for i in range(n)] == exp
    d = jvm.fn("libAlgebraFFTAuxiliary.so", "Java_algebra_fft_FFTAuxiliary_serialRadix2FFTDirect", [vp, i32, vp, i32], i32)
    buf = jvm.direct(O.pack_scalars(x))
    assert d(jvm.env, None, buf, n, jvm.bytes_(O.le32(omega)), 0) == 0
    out = jvm.read(buf)
    assert [O.from_le(out[32 * i:32 * i + 32]) for i in range(n)] == exp


def test_concurrent_executor_threads(jvm):
    """Spark "Executor task launch worker" threads call the natives concurrently in one JVM (hs_err_pid98479.log:260-272;
    SURVEY.md section 8b): every thread gets its own context (stream, scratch) inside the shim, results must not mix."""
    import threading
    f = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_variableBaseSerialMSMNativeHelper", [vp, vp, i32, i32, i32])
    jobs = []
    for t in range(6):
        n = 3000 + 500 * t
        ks, pool = util.known_dlog_points(O.G1, 64, seed=500 + t, random_z=True)
        raw = util.rand_scalars_bytes(n, seed=500 + t)
        bases = util.tiled_bases_bytes(O.G1, pool, n)
        exp = util.expected_from_dlogs(O.G1, ks, util.column_sums(raw, 64))
        jobs.append((n, jvm.bytes_(bases.tobytes()), jvm.bytes_(raw.tobytes()), exp))
    results = [None] * len(jobs)

    def worker(i):
        n, jb, js, _ = jobs[i]
        outs = []
        for _ in range(4):                                   # several calls per thread, interleaving with the others
            outs.append(jvm.read(f(jvm.env, None, jb, js, n, 1, i)))
        results[i] = outs

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert jvm.exception() is None
    for (n, _, _, exp), outs in zip(jobs, results):
        for out in outs:
            assert len(out) == 192 and O.G1.equals(O.unpack_g1(out, stride=64)[0], exp)
    assert jvm.live_pins() == 0


def test_critical_regions_hold_no_jni_calls_and_are_per_slice(jvm):
    """JNI rule: no JNI function between Get/ReleasePrimitiveArrayCritical; and the big MSM arrays are pinned one slice at a
    time (ozk_msm_feed), not across the GPU call.  A pin that fails must surface as a RuntimeException raised OUTSIDE any
    critical region, with every other pin released and the context usable afterwards."""
    f = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_variableBaseSerialMSMNativeHelper", [vp, vp, i32, i32, i32])
    n = 1 << 18                                              # two slices (ozk_msm_plan_slices)
    ks, pool = util.known_dlog_points(O.G1, 64, seed=91, random_z=True)
    raw = util.rand_scalars_bytes(n, seed=91)
    bases = util.tiled_bases_bytes(O.G1, pool, n)
    exp = util.expected_from_dlogs(O.G1, ks, util.column_sums(raw, 64))
    jb, js = jvm.bytes_(bases.tobytes()), jvm.bytes_(raw.tobytes())
    p0 = jvm.pin_calls()
    r = f(jvm.env, None, jb, js, n, 1, 0)
    assert r and jvm.exception() is None
    assert O.G1.equals(O.unpack_g1(jvm.read(r), stride=64)[0], exp)
    assert jvm.pin_calls() - p0 == 4                         # 2 slices x (scalars, bases)
    assert jvm.live_pins() == 0 and jvm.critical_violations() == 0
    # a pin fails: exception, no pin left, nothing thrown from inside a critical region
    jvm.fail_next_pins(1)                                    # the very first pin of the call fails
    r = f(jvm.env, None, jb, js, n, 1, 0)
    assert not r and "could not pin" in jvm.exception()
    jvm.clear()
    jvm.fail_next_pins(0)
    assert jvm.live_pins() == 0 and jvm.critical_violations() == 0
    r = f(jvm.env, None, jb, js, n, 1, 0)                    # the context is usable again
    assert r and O.G1.equals(O.unpack_g1(jvm.read(r), stride=64)[0], exp)
    # fixed-base and field natives never pin
    p0 = jvm.pin_calls()
    h = jvm.fn("libAlgebraMSMFixedBaseMSM.so", "Java_algebra_msm_FixedBaseMSM_fieldBatchMSMNativeHelper", [vp, i32, i32])
    r = h(jvm.env, None, jvm.bytes_(O.pack_scalars([1, 2, 3]) + O.le32(5)), 3, 0)
    out = jvm.read(r)
    assert [int.from_bytes(out[64 * i:64 * i + 64], "big") for i in range(3)] == [5, 10, 15]
    assert jvm.pin_calls() == p0 and jvm.critical_violations() == 0


def test_stale_key_handles_throw(jvm):
    """A jlong that is not a live handle of uploadBasesDirect (freed, or garbage) is a RuntimeException, not a wild pointer."""
    i64 = ctypes.c_int64
    up = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_uploadBasesDirect", [vp, i32, i32, i32], i64)
    keyed = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_variableBaseMSMKeyedDirect",
                   [i64, i64, vp, i32, i32, i32, i32, vp], i32)
    free = jvm.fn("libAlgebraMSMVariableBaseMSM.so", "Java_algebra_msm_VariableBaseMSM_freeBases", [i64, i32], None)
    pts = [O.G1.mul(O.G1.generator, k) for k in (3, 4)]
    k1 = up(jvm.env, None, jvm.direct(O.pack_g1(pts)), 2, 1, 0)
    assert k1
    ob = jvm.direct(size=96)
    sc = jvm.direct(O.pack_scalars([5, 6]))
    assert keyed(jvm.env, None, k1, 0, sc, 0, 2, 1, 0, ob) == 0
    assert O.G1.equals(O.unpack_g1(jvm.read(ob))[0], O.G1.mul(O.G1.generator, 39))
    free(jvm.env, None, k1, 0)
    assert jvm.exception() is None
    assert keyed(jvm.env, None, k1, 0, sc, 0, 2, 1, 0, ob) == -1           # freed handle
    assert "live handle" in jvm.exception()
    jvm.clear()
    assert keyed(jvm.env, None, 0x1234, 0, sc, 0, 2, 1, 0, ob) == -1       # garbage
    assert "live handle" in jvm.exception()
    jvm.clear()
    free(jvm.env, None, k1, 0)                                             # double free
    assert "live handle" in jvm.exception()
    jvm.clear()
