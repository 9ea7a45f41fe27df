"""End-to-end timing through the reference-facing JNI symbols (SURVEY.md section 8d "and through the fake-JNI byte[] path"):
the legacy byte[] native the unmodified dizk.jar calls, the direct-ByteBuffer native and the keyed native (bases resident),
on the same G1 inputs.  Wall clock of the whole native call as the JVM would see it (the fake JNIEnv hands out malloc'ed,
pageable arrays, like a JVM heap).  Sizes follow the Java's own chunking: at most 2^23 G1 points per call
(VariableBaseMSM.java:211).

    python tools/jni_bench.py [log_n ...]"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import dizk_oracle as O  # noqa: E402
from tests import util  # noqa: E402
from tests.jni_harness import FakeJvm  # noqa: E402

vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
jvm = FakeJvm()
LIB = "libAlgebraMSMVariableBaseMSM.so"
legacy = jvm.fn(LIB, "Java_algebra_msm_VariableBaseMSM_variableBaseSerialMSMNativeHelper", [vp, vp, i32, i32, i32])
direct = jvm.fn(LIB, "Java_algebra_msm_VariableBaseMSM_variableBaseMSMDirect", [vp, vp, vp, i32, i32, i32, vp], i32)
upload = jvm.fn(LIB, "Java_algebra_msm_VariableBaseMSM_uploadBasesDirect", [vp, i32, i32, i32], i64)
keyed = jvm.fn(LIB, "Java_algebra_msm_VariableBaseMSM_variableBaseMSMKeyedDirect", [i64, i64, vp, i32, i32, i32, i32, vp], i32)
free = jvm.fn(LIB, "Java_algebra_msm_VariableBaseMSM_freeBases", [i64, i32], None)


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, r


for log_n in [int(a) for a in sys.argv[1:]] or [20, 22, 23]:
    n = 1 << log_n
    ks, pool = util.known_dlog_points(O.G1, 64, seed=log_n, random_z=True)
    raw = util.rand_scalars_bytes(n, seed=log_n)
    bases = np.ascontiguousarray(util.tiled_bases_bytes(O.G1, pool, n))
    exp = util.expected_from_dlogs(O.G1, ks, util.column_sums(raw, 64))
    jb, js = jvm.bytes_(bases.tobytes()), jvm.bytes_(raw.tobytes())
    legacy(jvm.env, None, jb, js, n, 1, 0)                                                   # warm-up (context, buffers)
    ms_legacy, r = best(lambda: legacy(jvm.env, None, jb, js, n, 1, 0))
    ok = O.G1.equals(O.unpack_g1(jvm.read(r), stride=64)[0], exp)
    db, ds, do = jvm.direct(bases.tobytes()), jvm.direct(raw.tobytes()), jvm.direct(size=96)
    ms_direct, rc = best(lambda: direct(jvm.env, None, db, None, ds, n, 1, 0, do))
    ok = ok and rc == 0 and O.G1.equals(O.unpack_g1(jvm.read(do))[0], exp)
    key = upload(jvm.env, None, db, n, 1, 0)
    ms_keyed, rc = best(lambda: keyed(jvm.env, None, key, 0, ds, 0, n, 1, 0, do))
    ok = ok and rc == 0 and O.G1.equals(O.unpack_g1(jvm.read(do))[0], exp)
    free(jvm.env, None, key, 0)
    print(json.dumps({"op": "jni_var_msm_g1", "log_n": log_n, "ok": bool(ok), "legacy_byte_array_ms": ms_legacy, "direct_buffer_ms": ms_direct,
                      "keyed_direct_ms": ms_keyed, "Mpairs_per_s_legacy": n / ms_legacy / 1e3,
                      "note": "pageable host memory (malloc), wall clock of the whole native call"}), flush=True)
