"""ctypes loader for oracle/_ref/ -- the reference's own CUDA sources compiled unmodified for sm_100a by
oracle/Makefile.ref (test infrastructure; needs a GPU to run).  Byte layouts are the reference's JNI formats
(SURVEY.md Appendix A): 32-byte LE inputs, 64-byte LE (variable-base, FFT) or 64-byte BE (fixed-base, field) outputs."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int


def available(name: str) -> bool:
    return os.path.exists(os.path.join(_REF, f"libref_{name}.so"))


def _load(name):
    return ctypes.CDLL(os.path.join(_REF, f"libref_{name}.so"))


def var_msm(bases: bytes, scalars: bytes, n: int, type_: int) -> bytes:
    lib = _load("varmsm")
    lib.ref_var_msm.restype = ctypes.c_long
    lib.ref_var_msm.argtypes = [vp, sz, vp, sz, i32, i32, vp, sz]
    out = ctypes.create_string_buffer(384)
    r = lib.ref_var_msm(bases, len(bases), scalars, len(scalars), n, type_, out, 384)
    assert r in (192, 384), r
    return out.raw[:r]


def var_double_msm(bases1: bytes, bases2: bytes, scalars: bytes, n: int) -> bytes:
    lib = _load("varmsm")
    lib.ref_var_double_msm.restype = ctypes.c_long
    lib.ref_var_double_msm.argtypes = [vp, sz, vp, sz, vp, sz, i32, vp, sz]
    out = ctypes.create_string_buffer(576)
    r = lib.ref_var_double_msm(bases1, len(bases1), bases2, len(bases2), scalars, len(scalars), n, out, 576)
    assert r == 576, r
    return out.raw


def fixed_batch(outerc, window, out_len, inner_len, n, scalar_size, base: bytes, scalars: bytes, bn_type: int) -> bytes:
    lib = _load("fixedmsm")
    lib.ref_fixed_batch.restype = ctypes.c_long
    lib.ref_fixed_batch.argtypes = [i32, i32, i32, i32, i32, i32, vp, sz, vp, sz, i32, vp, sz]
    cap = n * (192 if bn_type == 1 else 384)
    out = ctypes.create_string_buffer(cap)
    r = lib.ref_fixed_batch(outerc, window, out_len, inner_len, n, scalar_size, base, len(base), scalars, len(scalars), bn_type, out, cap)
    assert r == cap, r
    return out.raw


def fixed_double_batch(outerc1, window1, outerc2, window2, out_len1, inner_len1, out_len2, inner_len2, n, base1: bytes, base2: bytes,
                       scalars: bytes) -> bytes:
    lib = _load("fixedmsm")
    lib.ref_fixed_double_batch.restype = ctypes.c_long
    lib.ref_fixed_double_batch.argtypes = [i32] * 9 + [vp, sz, vp, sz, vp, sz, vp, sz]
    cap = n * 576
    out = ctypes.create_string_buffer(cap)
    r = lib.ref_fixed_double_batch(outerc1, window1, outerc2, window2, out_len1, inner_len1, out_len2, inner_len2, n, base1, len(base1),
                                   base2, len(base2), scalars, len(scalars), out, cap)
    assert r == cap, r
    return out.raw


def field_batch(scalars_plus_base: bytes, n: int) -> bytes:
    lib = _load("fixedmsm")
    lib.ref_field_batch.restype = ctypes.c_long
    lib.ref_field_batch.argtypes = [vp, sz, i32, vp, sz]
    out = ctypes.create_string_buffer(n * 64)
    r = lib.ref_field_batch(scalars_plus_base, len(scalars_plus_base), n, out, n * 64)
    assert r == n * 64, r
    return out.raw


def fft(data: bytes, omega: bytes) -> bytes:
    lib = _load("fft")
    lib.ref_fft.restype = ctypes.c_long
    lib.ref_fft.argtypes = [vp, i32, i32, vp, i32, vp, sz]
    n = len(data) // 32
    out = ctypes.create_string_buffer(n * 64)
    r = lib.ref_fft(data, n, 32, omega, len(omega), out, n * 64)
    assert r == n * 64, r
    return out.raw


# ---- isolation -------------------------------------------------------------------------------------------------------
# The reference's code leaks device memory, synchronises the whole device and exit(-1)s on CUDA errors
# (algebra_msm_VariableBaseMSM.cu:1417-1422), so tests call it in a child process with a time limit: a crash or a hang of
# the reference can then neither take the test process down nor stall the suite.
def _child(name, args, q):
    try:
        q.put(("ok", globals()[name](*args)))
    except BaseException as e:      # noqa: BLE001
        q.put(("err", repr(e)))


def isolated(name: str, *args, timeout: float = 180.0):
    """Run ref_cuda.<name>(*args) in a spawned child; returns its result, or raises TimeoutError / RuntimeError."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_child, args=(name, args, q), daemon=True)
    p.start()
    try:
        status, val = q.get(timeout=timeout)
    except Exception:
        p.kill()
        p.join(10)
        raise TimeoutError(f"reference CUDA call {name} did not answer within {timeout} s")
    p.join(30)
    if p.is_alive():
        p.kill()
    if status != "ok":
        raise RuntimeError(f"reference CUDA call {name} failed: {val}")
    return val
