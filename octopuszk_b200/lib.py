"""ctypes binding of the C ABI in include/octozk.h.

The library is built in-tree by octopuszk_b200/build.py.  There is no fallback of any kind: if liboctozk.so is
missing, or no CUDA device is usable, loading / context creation raises."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[ctypes.CDLL] = None

OZK_OK = 0


class OzkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"octozk error {code}: {msg}")
        self.code = code


def library_path() -> str:
    return os.path.join(_HERE, "lib", "liboctozk.so")


_c_u8p = ctypes.c_char_p
_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_int = ctypes.c_int

# name -> (restype, argtypes); every symbol include/octozk.h declares
SIGNATURES = {
    "ozk_device_count": (_int, []),
    "ozk_ctx_create": (_int, [_int, ctypes.POINTER(_vp)]),
    "ozk_ctx_destroy": (None, [_vp]),
    "ozk_ctx_set_stream": (_int, [_vp, _vp]),
    "ozk_ctx_sync": (_int, [_vp]),
    "ozk_ctx_launches": (ctypes.c_ulonglong, [_vp]),
    "ozk_last_error": (ctypes.c_char_p, []),
    "ozk_version": (ctypes.c_char_p, []),
    "ozk_fr_scale": (_int, [_vp, _vp, _sz, _c_u8p, _vp]),
    "ozk_fr_scale_dev": (_int, [_vp, _vp, _vp, _sz, _c_u8p]),
    "ozk_fr_scale_powers_dev": (_int, [_vp, _vp, _vp, _sz, _c_u8p, _c_u8p, ctypes.c_uint64]),
    "ozk_fr_mul_sub_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _sz]),
    "ozk_fr_spmv_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ozk_fr_spmv_ex_dev": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _sz, _vp]),
    "ozk_fr_lincomb_dev": (_int, [_vp, _vp, _sz, _vp, _c_u8p, _vp, _c_u8p, _vp, _c_u8p]),
    "ozk_fr_lagrange_dev": (_int, [_vp, _vp, _sz, _c_u8p, _c_u8p]),
    "ozk_ntt_fr_scatter_dev": (_int, [_vp, _vp, ctypes.POINTER(_vp), _sz, _sz, _sz, _c_u8p, _c_u8p]),
    "ozk_fr_dft_small_scatter_dev": (_int, [_vp, _vp, ctypes.POINTER(_vp), _sz, _sz, _sz, _c_u8p, _c_u8p]),
    "ozk_peer_alloc": (_int, [_vp, _sz, ctypes.POINTER(_vp), ctypes.c_char_p]),
    "ozk_peer_open": (_int, [_vp, ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "ozk_peer_close": (_int, [_vp, _vp]),
    "ozk_peer_free": (_int, [_vp, _vp]),
    "ozk_fr_dft_small_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _c_u8p]),
    "ozk_ntt_fr": (_int, [_vp, _vp, _sz, _c_u8p]),
    "ozk_ntt_fr_dev": (_int, [_vp, _vp, _vp, _sz, _c_u8p]),
    "ozk_ntt_fr_ex_dev": (_int, [_vp, _vp, _vp, _sz, _c_u8p, _c_u8p, _c_u8p, _c_u8p]),
    "ozk_imad_peak": (_int, [_vp, ctypes.POINTER(ctypes.c_double)]),
    "ozk_modmul_peak": (_int, [_vp, ctypes.POINTER(ctypes.c_double)]),
    "ozk_pipe_probe": (_int, [_vp, _int, ctypes.POINTER(ctypes.c_double)]),
    # MSM_BEGIN
    "ozk_msm_g1": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "ozk_msm_g1_dev": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "ozk_msm_g2": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "ozk_msm_g2_dev": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "ozk_msm_g1g2": (_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "ozk_msm_g1g2_dev": (_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "ozk_msm_begin": (_int, [_vp, _int, _sz, _sz, _vp, _vp, _sz]),
    "ozk_msm_feed": (_int, [_vp, _vp, _vp, _vp, _sz]),
    "ozk_msm_end": (_int, [_vp, _vp]),
    "ozk_msm_plan_slices": (_int, [_sz, ctypes.POINTER(_sz), _int]),
    "ozk_fixed_g1": (_int, [_vp, _c_u8p, _vp, _sz, _int, _int, _vp]),
    "ozk_fixed_g1_dev": (_int, [_vp, _c_u8p, _vp, _sz, _int, _int, _vp]),
    "ozk_fixed_g2": (_int, [_vp, _c_u8p, _vp, _sz, _int, _int, _vp]),
    "ozk_fixed_g2_dev": (_int, [_vp, _c_u8p, _vp, _sz, _int, _int, _vp]),
    "ozk_fixed_g1_ex_dev": (_int, [_vp, _c_u8p, _vp, _sz, _int, _int, ctypes.c_uint, _vp]),
    "ozk_fixed_g2_ex_dev": (_int, [_vp, _c_u8p, _vp, _sz, _int, _int, ctypes.c_uint, _vp]),
    "ozk_sum_g1_dev": (_int, [_vp, _vp, _sz, _vp]),
    "ozk_sum_g2_dev": (_int, [_vp, _vp, _sz, _vp]),
    "ozk_msm_last_stats": (_int, [_vp, ctypes.POINTER(ctypes.c_double), _int]),
    "ozk_bases_upload_g1": (_int, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "ozk_bases_upload_g1_dev": (_int, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "ozk_bases_upload_g2": (_int, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "ozk_bases_upload_g2_dev": (_int, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "ozk_bases_len": (_sz, [_vp]),
    "ozk_bases_free": (None, [_vp, _vp]),
    "ozk_msm_g1_keyed": (_int, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "ozk_msm_g1_keyed_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "ozk_msm_g2_keyed": (_int, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "ozk_msm_g2_keyed_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "ozk_msm_g1g2_keyed": (_int, [_vp, _vp, _vp, _vp, _sz, _sz, _vp]),
    "ozk_msm_g1g2_keyed_dev": (_int, [_vp, _vp, _vp, _vp, _sz, _sz, _vp]),
    # MSM_END
}


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or library_path()
    if not os.path.exists(p):
        raise OzkError(-100, f"{p} not found: build it with `python -m octopuszk_b200.build` (no CPU fallback exists)")
    lib = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _LIB = lib
    return lib


def _ptr(x):
    """Address of a bytes-like / numpy / torch buffer, or a raw int device pointer."""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    if isinstance(x, bytes):
        return x                      # ctypes passes the buffer address and keeps the object alive for the call
    if isinstance(x, bytearray):
        return (ctypes.c_char * len(x)).from_buffer(x)
    if isinstance(x, ctypes.Array):
        return x
    raise TypeError(type(x))


class Bases:
    """Persistent device-resident bases (a proving-key query vector), see ozk_bases_upload_* in include/octozk.h."""

    def __init__(self, ctx: "Context", handle, group: int):
        self.ctx, self._h, self.group = ctx, handle, group

    def __len__(self) -> int:
        return int(self.ctx.lib.ozk_bases_len(self._h)) if self._h is not None else 0

    def free(self):
        if self._h is not None and self.ctx._h is not None:
            self.ctx.lib.ozk_bases_free(self.ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One liboctozk context (own stream + scratch) on one device."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = load_library()
        h = ctypes.c_void_p()
        self._h = None
        self._check(self.lib.ozk_ctx_create(device, ctypes.byref(h)))
        self._h = h
        self.device = device
        self._stream = None                  # None: the context's own stream
        if stream is not None:
            self.set_stream(stream)

    def _check(self, rc: int):
        if rc != OZK_OK:
            raise OzkError(rc, self.lib.ozk_last_error().decode())

    def close(self):
        if self._h is not None:
            self.lib.ozk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.ozk_ctx_set_stream(self._h, ctypes.c_void_p(cuda_stream)))
        self._stream = cuda_stream

    def use_torch_stream(self):
        """Run all later work of this context on torch's CURRENT stream of the context's device.  The "_dev" entry points only
        enqueue on the context's stream (stream contract in include/octozk.h), so work that mixes torch operations, NCCL
        collectives and liboctozk kernels on the same buffers is ordered only if they share a stream.  Every wrapper below that
        is handed a CUDA torch tensor calls this first, so the contract holds for any caller (also inside
        `with torch.cuda.stream(s):`); switching streams is ordered on the device and costs nothing when the stream is unchanged."""
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s != self._stream:
            self.set_stream(s)

    def _dp(self, x):
        """Device-pointer argument: address of a CUDA torch tensor (binding the context to torch's current stream) or a raw int."""
        if x is not None and getattr(x, "is_cuda", False):
            self.use_torch_stream()
        return _ptr(x)

    def sync(self):
        self._check(self.lib.ozk_ctx_sync(self._h))

    def launches(self) -> int:
        return int(self.lib.ozk_ctx_launches(self._h))

    # ---- diagnostics
    def imad_peak(self) -> float:
        v = ctypes.c_double()
        self._check(self.lib.ozk_imad_peak(self._h, ctypes.byref(v)))
        return v.value

    def pipe_probe(self, which: int) -> float:
        v = ctypes.c_double()
        self._check(self.lib.ozk_pipe_probe(self._h, which, ctypes.byref(v)))
        return v.value

    def modmul_peak(self) -> float:
        v = ctypes.c_double()
        self._check(self.lib.ozk_modmul_peak(self._h, ctypes.byref(v)))
        return v.value

    # ---- Fr
    def fr_scale(self, a: bytes, b: bytes) -> bytes:
        n = len(a) // 32
        out = ctypes.create_string_buffer(n * 32)
        self._check(self.lib.ozk_fr_scale(self._h, a, n, b, out))
        return out.raw

    def fr_scale_dev(self, d_a, d_out, n: int, b: bytes):
        self._check(self.lib.ozk_fr_scale_dev(self._h, self._dp(d_a), self._dp(d_out), n, b))

    def fr_scale_powers_dev(self, d_a, d_out, n: int, scale=None, coset=None, first_index: int = 0):
        self._check(self.lib.ozk_fr_scale_powers_dev(self._h, self._dp(d_a), self._dp(d_out), n, scale, coset, first_index))

    def fr_mul_sub_dev(self, d_a, d_b, d_c, d_out, n: int):
        self._check(self.lib.ozk_fr_mul_sub_dev(self._h, self._dp(d_a), self._dp(d_b), self._dp(d_c) if d_c is not None else None, self._dp(d_out), n))

    def fr_spmv_dev(self, d_row_ptr, d_col, d_coeff, d_z, rows: int, d_out):
        self._check(self.lib.ozk_fr_spmv_dev(self._h, self._dp(d_row_ptr), self._dp(d_col), self._dp(d_coeff), self._dp(d_z), rows, self._dp(d_out)))

    def fr_spmv_ex_dev(self, d_row_ptr, d_col, d_coeff, d_z, z_len: int, rows: int, d_out):
        """d_coeff None: unit coefficients; column indices are checked against z_len."""
        self._check(self.lib.ozk_fr_spmv_ex_dev(self._h, self._dp(d_row_ptr), self._dp(d_col), self._dp(d_coeff) if d_coeff is not None else None,
                                                self._dp(d_z), z_len, rows, self._dp(d_out)))

    def fr_lincomb_dev(self, d_out, n: int, d_a, ca: bytes, d_b=None, cb: bytes = None, d_c=None, cc: bytes = None):
        self._check(self.lib.ozk_fr_lincomb_dev(self._h, self._dp(d_out), n, self._dp(d_a), ca, self._dp(d_b) if d_b is not None else None, cb,
                                                self._dp(d_c) if d_c is not None else None, cc))

    def fr_lagrange_dev(self, d_out, m: int, t: bytes, omega: bytes):
        self._check(self.lib.ozk_fr_lagrange_dev(self._h, self._dp(d_out), m, t, omega))

    def ntt_scatter_dev(self, d_in, peer_ptrs, rank: int, n_local: int, omega_local: bytes, twiddle_base: bytes):
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        self._check(self.lib.ozk_ntt_fr_scatter_dev(self._h, self._dp(d_in), arr, len(peer_ptrs), rank, n_local, omega_local, twiddle_base))

    def dft_small_scatter_dev(self, d_in, peer_ptrs, rank: int, length: int, omega_g: bytes, omega_n: bytes):
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        self._check(self.lib.ozk_fr_dft_small_scatter_dev(self._h, self._dp(d_in), arr, len(peer_ptrs), rank, length, omega_g, omega_n))

    def peer_alloc(self, nbytes: int):
        """(device pointer, 64-byte IPC handle) of a fresh cudaMalloc block other processes can map."""
        p = ctypes.c_void_p()
        h = ctypes.create_string_buffer(64)
        self._check(self.lib.ozk_peer_alloc(self._h, nbytes, ctypes.byref(p), h))
        return int(p.value), h.raw

    def peer_open(self, handle: bytes) -> int:
        p = ctypes.c_void_p()
        self._check(self.lib.ozk_peer_open(self._h, handle, ctypes.byref(p)))
        return int(p.value)

    def peer_close(self, ptr: int):
        self._check(self.lib.ozk_peer_close(self._h, ctypes.c_void_p(ptr)))

    def peer_free(self, ptr: int):
        self._check(self.lib.ozk_peer_free(self._h, ctypes.c_void_p(ptr)))

    def fr_dft_small_dev(self, d_in, d_out, groups: int, length: int, omega_g: bytes):
        self._check(self.lib.ozk_fr_dft_small_dev(self._h, self._dp(d_in), self._dp(d_out), groups, length, omega_g))

    # ---- NTT
    def ntt(self, data: bytes, omega: bytes) -> bytes:
        n = len(data) // 32
        buf = ctypes.create_string_buffer(bytes(data), n * 32)
        self._check(self.lib.ozk_ntt_fr(self._h, buf, n, omega))
        return buf.raw

    def ntt_host_inplace(self, buf, n: int, omega: bytes):
        self._check(self.lib.ozk_ntt_fr(self._h, _ptr(buf), n, omega))

    def ntt_dev(self, d_in, d_out, n: int, omega: bytes):
        self._check(self.lib.ozk_ntt_fr_dev(self._h, self._dp(d_in), self._dp(d_out), n, omega))

    def ntt_ex_dev(self, d_in, d_out, n: int, omega: bytes, pre_coset=None, post_scale=None, post_coset=None):
        self._check(self.lib.ozk_ntt_fr_ex_dev(self._h, self._dp(d_in), self._dp(d_out), n, omega, pre_coset, post_scale, post_coset))

    # ---- variable-base MSM
    def msm_g1(self, scalars, bases, n: int) -> bytes:
        out = ctypes.create_string_buffer(96)
        self._check(self.lib.ozk_msm_g1(self._h, _ptr(scalars), _ptr(bases), n, out))
        return out.raw

    def msm_g1_dev(self, d_scalars, d_bases, n: int) -> bytes:
        out = ctypes.create_string_buffer(96)
        self._check(self.lib.ozk_msm_g1_dev(self._h, self._dp(d_scalars), self._dp(d_bases), n, out))
        return out.raw

    def msm_g2(self, scalars, bases, n: int) -> bytes:
        out = ctypes.create_string_buffer(192)
        self._check(self.lib.ozk_msm_g2(self._h, _ptr(scalars), _ptr(bases), n, out))
        return out.raw

    def msm_g2_dev(self, d_scalars, d_bases, n: int) -> bytes:
        out = ctypes.create_string_buffer(192)
        self._check(self.lib.ozk_msm_g2_dev(self._h, self._dp(d_scalars), self._dp(d_bases), n, out))
        return out.raw

    def msm_g1g2(self, scalars, bases1, bases2, n: int) -> bytes:
        out = ctypes.create_string_buffer(288)
        self._check(self.lib.ozk_msm_g1g2(self._h, _ptr(scalars), _ptr(bases1), _ptr(bases2), n, out))
        return out.raw

    def msm_g1g2_dev(self, d_scalars, d_bases1, d_bases2, n: int) -> bytes:
        out = ctypes.create_string_buffer(288)
        self._check(self.lib.ozk_msm_g1g2_dev(self._h, self._dp(d_scalars), self._dp(d_bases1), self._dp(d_bases2), n, out))
        return out.raw

    # ---- streaming form: announce the total, feed slices of host arrays, collect the result
    def msm_begin(self, groups: int, n_total: int, max_slice: int = 0, key1=None, key2=None, first: int = 0):
        self._check(self.lib.ozk_msm_begin(self._h, groups, n_total, max_slice, key1._h if key1 else None, key2._h if key2 else None, first))

    def msm_feed(self, scalars, bases1, bases2, n: int):
        self._check(self.lib.ozk_msm_feed(self._h, _ptr(scalars), _ptr(bases1), _ptr(bases2), n))

    def msm_end(self, groups: int) -> bytes:
        out = ctypes.create_string_buffer({1: 96, 2: 192, 3: 288}[groups])
        self._check(self.lib.ozk_msm_end(self._h, out))
        return out.raw

    # ---- persistent bases + keyed MSM (scalars only per call)
    def upload_bases(self, group: int, bases, n: int, device: bool = False) -> Bases:
        fn = getattr(self.lib, f"ozk_bases_upload_g{group}" + ("_dev" if device else ""))
        h = ctypes.c_void_p()
        self._check(fn(self._h, self._dp(bases) if device else _ptr(bases), n, ctypes.byref(h)))
        return Bases(self, h, group)

    def msm_keyed(self, scalars, key: Bases, n: int, first: int = 0, device: bool = False) -> bytes:
        out = ctypes.create_string_buffer(96 if key.group == 1 else 192)
        fn = getattr(self.lib, f"ozk_msm_g{key.group}_keyed" + ("_dev" if device else ""))
        self._check(fn(self._h, self._dp(scalars) if device else _ptr(scalars), key._h, first, n, out))
        return out.raw

    def msm_g1g2_keyed(self, scalars, key1: Bases, key2: Bases, n: int, first: int = 0, device: bool = False) -> bytes:
        out = ctypes.create_string_buffer(288)
        fn = getattr(self.lib, "ozk_msm_g1g2_keyed" + ("_dev" if device else ""))
        self._check(fn(self._h, self._dp(scalars) if device else _ptr(scalars), key1._h, key2._h, first, n, out))
        return out.raw

    def sum_points_dev(self, group: int, d_points, k: int) -> bytes:
        out = ctypes.create_string_buffer(96 if group == 1 else 192)
        self._check(getattr(self.lib, f"ozk_sum_g{group}_dev")(self._h, self._dp(d_points), k, out))
        return out.raw

    def msm_last_stats(self):
        arr = (ctypes.c_double * 16)()
        k = self.lib.ozk_msm_last_stats(self._h, arr, 16)
        return list(arr)[:max(k, 0)]

    # ---- fixed-base batch MSM
    def fixed_g1(self, base: bytes, scalars, n: int, outerc: int, window: int) -> bytes:
        out = ctypes.create_string_buffer(n * 96)
        self._check(self.lib.ozk_fixed_g1(self._h, base, _ptr(scalars), n, outerc, window, out))
        return out.raw

    def fixed_g1_dev(self, base: bytes, d_scalars, n: int, outerc: int, window: int, d_out, keep_z: bool = False):
        self._check(self.lib.ozk_fixed_g1_ex_dev(self._h, base, self._dp(d_scalars), n, outerc, window, 1 if keep_z else 0, self._dp(d_out)))

    def fixed_g2(self, base: bytes, scalars, n: int, outerc: int, window: int) -> bytes:
        out = ctypes.create_string_buffer(n * 192)
        self._check(self.lib.ozk_fixed_g2(self._h, base, _ptr(scalars), n, outerc, window, out))
        return out.raw

    def fixed_g2_dev(self, base: bytes, d_scalars, n: int, outerc: int, window: int, d_out, keep_z: bool = False):
        self._check(self.lib.ozk_fixed_g2_ex_dev(self._h, base, self._dp(d_scalars), n, outerc, window, 1 if keep_z else 0, self._dp(d_out)))
