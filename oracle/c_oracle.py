"""ctypes loader for oracle/dizk_oracle.c (test infrastructure: see that file's header).  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdizk_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_HERE, "dizk_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    lib = ctypes.CDLL(_SO)
    vp, sz, i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.oracle_msm_g1.argtypes = [vp, vp, sz, i, vp]
    lib.oracle_g1_equal.argtypes = [vp, vp]
    lib.oracle_g1_to_affine.argtypes = [vp, vp]
    lib.oracle_fft_fr.argtypes = [vp, sz, vp]
    lib.oracle_fft_fr_batch.argtypes = [vp, sz, sz, vp, i]
    lib.oracle_fixed_g1.argtypes = [vp, vp, sz, i, i, i, vp]
    lib.oracle_fr_scale.argtypes = [vp, sz, vp, vp]
    lib.oracle_fr_dot.argtypes = [vp, vp, sz, i, vp]
    lib.oracle_fr_coset_scale.argtypes = [vp, sz, vp, vp]
    lib.oracle_fr_horner.argtypes = [vp, sz, vp, sz, i, vp]
    lib.oracle_fr_lagrange_eval.argtypes = [vp, sz, vp, vp, i, vp]
    lib.oracle_r1cs_chain.argtypes = [sz, vp, vp, vp]
    lib.oracle_r1cs_synth_eval.argtypes = [sz, sz, vp, sz, sz, vp, vp, vp]
    _lib = lib
    return lib


def max_threads() -> int:
    return load().oracle_max_threads()


def _addr(x):
    if isinstance(x, (bytes, ctypes.Array)):
        return x
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(type(x))


def msm_g1(scalars, bases, n: int, threads: int = 1) -> bytes:
    out = ctypes.create_string_buffer(96)
    assert load().oracle_msm_g1(_addr(scalars), _addr(bases), n, threads, out) == 0
    return out.raw


def g1_equal(a: bytes, b: bytes) -> bool:
    return bool(load().oracle_g1_equal(a, b))


def g1_to_affine(a: bytes) -> bytes:
    out = ctypes.create_string_buffer(96)
    load().oracle_g1_to_affine(a, out)
    return out.raw


def fft_fr(data: bytes, omega: bytes) -> bytes:
    n = len(data) // 32
    buf = ctypes.create_string_buffer(bytes(data), len(data))
    assert load().oracle_fft_fr(buf, n, omega) == 0
    return buf.raw


def fft_fr_batch_inplace(buf, n: int, batch: int, omega: bytes, threads: int):
    assert load().oracle_fft_fr_batch(_addr(buf), n, batch, omega, threads) == 0


def fixed_g1(base: bytes, scalars, n: int, scalar_size: int, window: int, threads: int = 1) -> bytes:
    out = ctypes.create_string_buffer(96 * n)
    assert load().oracle_fixed_g1(base, _addr(scalars), n, scalar_size, window, threads, out) == 0
    return out.raw


def fr_scale(a: bytes, b: bytes) -> bytes:
    n = len(a) // 32
    out = ctypes.create_string_buffer(32 * n)
    load().oracle_fr_scale(a, n, b, out)
    return out.raw


def default_threads() -> int:
    import os
    return os.cpu_count() or max_threads()


def fr_dot(a, b, n: int, threads: int = 0) -> int:
    """sum_i a[i] * b[i] mod r of two arrays of n canonical 32-byte elements."""
    out = ctypes.create_string_buffer(32)
    assert load().oracle_fr_dot(_addr(a), _addr(b), n, threads or default_threads(), out) == 0
    return int.from_bytes(out.raw, "little")


def fr_horner(coeffs, n: int, xs, threads: int = 0):
    """[sum_j coeffs[j] * x^j mod r for x in xs] (xs: list of ints)."""
    xb = b"".join(int(x).to_bytes(32, "little") for x in xs)
    out = ctypes.create_string_buffer(32 * len(xs))
    assert load().oracle_fr_horner(_addr(coeffs), n, xb, len(xs), threads or default_threads(), out) == 0
    return [int.from_bytes(out.raw[32 * q:32 * q + 32], "little") for q in range(len(xs))]


def fr_lagrange_eval(vals, n: int, omega: int, t: int, threads: int = 0) -> int:
    """sum_j vals[j] * L_j(t) over the domain {omega^j}: the interpolant of vals evaluated at t."""
    out = ctypes.create_string_buffer(32)
    rc = load().oracle_fr_lagrange_eval(_addr(vals), n, int(omega).to_bytes(32, "little"), int(t).to_bytes(32, "little"),
                                        threads or default_threads(), out)
    assert rc == 0, rc
    return int.from_bytes(out.raw, "little")


def r1cs_chain(num_constraints: int, a0: int, b0: int):
    """Full assignment (numpy (num_constraints + 3, 32) uint8) of R1CSConstruction.serialConstruct."""
    import numpy as np
    out = np.empty((num_constraints + 3, 32), dtype=np.uint8)
    assert load().oracle_r1cs_chain(num_constraints, int(a0).to_bytes(32, "little"), int(b0).to_bytes(32, "little"), out.ctypes.data) == 0
    return out


def fr_coset_scale(a, g: int) -> bytes:
    """[a_i * g^i mod r] (FFTAuxiliary.multiplyByCoset)."""
    n = len(a) // 32
    out = ctypes.create_string_buffer(32 * n)
    load().oracle_fr_coset_scale(_addr(a), n, int(g).to_bytes(32, "little"), out)
    return out.raw


def r1cs_synth_eval(nc: int, ni: int, z, col_limit: int, n_dom: int):
    """(a, b, c) evaluation vectors ((n_dom, 32) uint8 each) of the synthetic circuit at assignment z, columns < col_limit only."""
    import numpy as np
    a, b, c = (np.empty((n_dom, 32), dtype=np.uint8) for _ in range(3))
    rc = load().oracle_r1cs_synth_eval(nc, ni, _addr(z), col_limit, n_dom, a.ctypes.data, b.ctypes.data, c.ctypes.data)
    assert rc == 0, rc
    return a, b, c
