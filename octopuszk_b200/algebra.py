"""Host-side mirror of the reference's Java operator interface for the hot path, on top of the C ABI.

Same names, argument meaning and error behaviour as the Java (snake_case aliases aside), so that the parity tests read
like the reference's own tests:

    algebra.msm.VariableBaseMSM   serialMSM / doubleMSM / distributedMSM      src/main/java/algebra/msm/VariableBaseMSM.java:199,480,772
    algebra.msm.FixedBaseMSM      getWindowSize / batchMSM / doubleBatchMSM   src/main/java/algebra/msm/FixedBaseMSM.java:49,185,487
                                  batchFieldMSMPartition (field batch)        :752-781
    algebra.fft.SerialFFT         radix2FFT / radix2InverseFFT / radix2CosetFFT / radix2CosetInverseFFT /
                                  divideByZOnCoset                            src/main/java/algebra/fft/SerialFFT.java:24-162
    algebra.fft.FFTAuxiliary      serialRadix2FFT                             src/main/java/algebra/fft/FFTAuxiliary.java:60

Field elements are Python ints in [0, modulus); G1 points are (X, Y, Z) int triples, G2 points ((X0,X1),(Y0,Y1),(Z0,Z1))
(Jacobian, infinity iff Z == 0), i.e. what BN254aG1.BN254G1ToBigInteger / BN254aG2.BN254G2ToBigInteger hand to the JNI
layer.  Everything is computed on the GPU through liboctozk; there is no CPU path here."""
from __future__ import annotations

from typing import List, Sequence, Tuple

from .lib import Context

# BN254aFrParameters.java:33-35, BN254aFqParameters.java:33
FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ_MODULUS = 21888242871839275222246405745257275088696311157297823662689037894645226208583
FR_ROOT = 19103219067921713944291392827692070036145651957329286315305642004821462161904
FR_MULTIPLICATIVE_GENERATOR = 5

# BN254aG1Parameters.java:25-50, BN254aG2Parameters.java:33-58
_G1_WINDOW_TABLE = [1, 5, 11, 32, 55, 162, 360, 815, 2373, 6978, 7122, 0, 57818, 0, 169679, 439759, 936073, 0, 4666555,
                    7580404, 0, 34552892]
_G2_WINDOW_TABLE = [1, 5, 10, 25, 59, 154, 334, 743, 2034, 4988, 8888, 26271, 39768, 106276, 141703, 462423, 926872, 0,
                    4873049, 5706708, 0, 31673815]


def _le32(x: int) -> bytes:
    return x.to_bytes(32, "little")          # raises OverflowError for values >= 2^256 or negative


def _is_g2(p) -> bool:
    return isinstance(p[0], (tuple, list))


def _pack_points(points) -> Tuple[bytes, bool]:
    g2 = _is_g2(points[0])
    if g2:
        return b"".join(_le32(c) for p in points for f in p for c in f), True
    return b"".join(_le32(c) for p in points for c in p), False


def _unpack_g1(b: bytes):
    return tuple(int.from_bytes(b[k * 32:(k + 1) * 32], "little") for k in range(3))


def _unpack_g2(b: bytes):
    v = [int.from_bytes(b[k * 32:(k + 1) * 32], "little") for k in range(6)]
    return ((v[0], v[1]), (v[2], v[3]), (v[4], v[5]))


class VariableBaseMSM:
    """algebra.msm.VariableBaseMSM on one GPU context."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def serialMSM(self, scalars: Sequence[int], bases: Sequence):
        """VariableBaseMSM.serialMSM (VariableBaseMSM.java:199-338): sum_i scalars[i] * bases[i]; G1 or G2 is chosen by
        the type of the bases, as the Java does by class name (:209,:266).  No 2^23 / 2^22 chunking is needed."""
        assert len(bases) == len(scalars)
        if len(scalars) == 0:
            raise ValueError("serialMSM: empty input (the Java dereferences bases.get(0))")
        pts, g2 = _pack_points(bases)
        sc = b"".join(_le32(s) for s in scalars)
        if g2:
            return _unpack_g2(self.ctx.msm_g2(sc, pts, len(scalars)))
        return _unpack_g1(self.ctx.msm_g1(sc, pts, len(scalars)))

    def doubleMSM(self, scalars: Sequence[int], bases: Sequence[Tuple]):
        """VariableBaseMSM.doubleMSM (VariableBaseMSM.java:480-606): bases is a list of (G1, G2) pairs."""
        assert len(bases) == len(scalars)
        if len(scalars) == 0:
            raise ValueError("doubleMSM: empty input")
        b1, _ = _pack_points([b[0] for b in bases])
        b2, _ = _pack_points([b[1] for b in bases])
        out = self.ctx.msm_g1g2(b"".join(_le32(s) for s in scalars), b1, b2, len(scalars))
        return _unpack_g1(out[:96]), _unpack_g2(out[96:])

    serial_msm = serialMSM
    double_msm = doubleMSM


class FixedBaseMSM:
    """algebra.msm.FixedBaseMSM on one GPU context."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    @staticmethod
    def getWindowSize(numScalars: int, groupFactory) -> int:
        """FixedBaseMSM.getWindowSize (FixedBaseMSM.java:49-66) with the per-curve tables."""
        table = _G2_WINDOW_TABLE if _is_g2(groupFactory) else _G1_WINDOW_TABLE
        window = 1
        for i in range(len(table) - 1, -1, -1):
            if table[i] != 0 and numScalars >= table[i]:
                window = i + 1
                break
        return window

    def batchMSM(self, scalarSize: int, windowSize: int, baseElement, scalars: Sequence[int]) -> List:
        """FixedBaseMSM.batchMSM (FixedBaseMSM.java:185-315): [s * base for s in scalars], each s truncated to
        outerc * windowSize bits with outerc = ceil(scalarSize / windowSize) (:214).  The Java's out_size / in_size
        arguments describe its own table and are dropped."""
        outerc = (scalarSize + windowSize - 1) // windowSize
        n = len(scalars)
        sc = b"".join(_le32(s) for s in scalars)
        base, g2 = _pack_points([baseElement])
        if g2:
            out = self.ctx.fixed_g2(base, sc, n, outerc, windowSize)
            return [_unpack_g2(out[192 * i:192 * (i + 1)]) for i in range(n)]
        out = self.ctx.fixed_g1(base, sc, n, outerc, windowSize)
        return [_unpack_g1(out[96 * i:96 * (i + 1)]) for i in range(n)]

    def doubleBatchMSM(self, scalarSize1: int, windowSize1: int, scalarSize2: int, windowSize2: int, base1, base2,
                       scalars: Sequence[int]) -> List[Tuple]:
        """FixedBaseMSM.doubleBatchMSM (FixedBaseMSM.java:487-602): [(s * base1, s * base2)]."""
        a = self.batchMSM(scalarSize1, windowSize1, base1, scalars)
        b = self.batchMSM(scalarSize2, windowSize2, base2, scalars)
        return list(zip(a, b))

    def batchFieldMSM(self, scalars: Sequence[int], b: int) -> List[int]:
        """Field batch of batchFieldMSMPartition (FixedBaseMSM.java:752-781): [a * b mod r]."""
        out = self.ctx.fr_scale(b"".join(_le32(s) for s in scalars), _le32(b))
        return [int.from_bytes(out[32 * i:32 * i + 32], "little") for i in range(len(scalars))]

    get_window_size = getWindowSize
    batch_msm = batchMSM
    double_batch_msm = doubleBatchMSM


class FFTAuxiliary:
    def __init__(self, ctx: Context):
        self.ctx = ctx

    def serialRadix2FFT(self, input: List[int], omega: int) -> None:
        """FFTAuxiliary.serialRadix2FFT (FFTAuxiliary.java:60-124): in place, natural order in and out."""
        n = len(input)
        if n == 1:
            return
        out = self.ctx.ntt(b"".join(_le32(v) for v in input), _le32(omega))
        for i in range(n):
            input[i] = int.from_bytes(out[32 * i:32 * i + 32], "little")

    def multiplyByCoset(self, input: List[int], g: int) -> None:
        """FFTAuxiliary.multiplyByCoset (FFTAuxiliary.java:224-232): input[i] *= g^i."""
        import torch
        n = len(input)
        d = torch.frombuffer(bytearray(b"".join(_le32(v) for v in input)), dtype=torch.uint8).to(torch.device("cuda", self.ctx.device))
        self.ctx.fr_scale_powers_dev(d, d, n, None, _le32(g))
        self.ctx.sync()
        out = d.cpu().numpy().tobytes()
        for i in range(n):
            input[i] = int.from_bytes(out[32 * i:32 * i + 32], "little")

    serial_radix2_fft = serialRadix2FFT


class SerialFFT:
    """algebra.fft.SerialFFT (SerialFFT.java:24-162) for BN254a Fr."""

    def __init__(self, ctx: Context, domainSize: int):
        assert domainSize > 1
        self.ctx = ctx
        d = 1
        while d < domainSize:
            d <<= 1
        self.domainSize = d                                                 # MathUtils.lowestPowerOfTwo, :26
        self.omega = pow(FR_ROOT, FR_MODULUS // d, FR_MODULUS)              # Fp.rootOfUnity, Fp.java:98-102

    def _run(self, input: List[int], **kw) -> None:
        import torch
        assert len(input) == self.domainSize
        n = self.domainSize
        dev = torch.device("cuda", self.ctx.device)
        d = torch.frombuffer(bytearray(b"".join(_le32(v) for v in input)), dtype=torch.uint8).to(dev)
        self.ctx.ntt_ex_dev(d, d, n, **kw)
        self.ctx.sync()
        out = d.cpu().numpy().tobytes()
        for i in range(n):
            input[i] = int.from_bytes(out[32 * i:32 * i + 32], "little")

    def radix2FFT(self, input: List[int]) -> None:                         # :75-80
        self._run(input, omega=_le32(self.omega))

    def radix2InverseFFT(self, input: List[int]) -> None:                  # :86-95
        self._run(input, omega=_le32(pow(self.omega, -1, FR_MODULUS)), post_scale=_le32(pow(self.domainSize, -1, FR_MODULUS)))

    def radix2CosetFFT(self, input: List[int], g: int) -> None:            # :100-105
        self._run(input, omega=_le32(self.omega), pre_coset=_le32(g))

    def radix2CosetInverseFFT(self, input: List[int], g: int) -> None:     # :111-115
        self._run(input, omega=_le32(pow(self.omega, -1, FR_MODULUS)), post_scale=_le32(pow(self.domainSize, -1, FR_MODULUS)),
                  post_coset=_le32(pow(g, -1, FR_MODULUS)))

    def lagrangeCoefficients(self, t: int) -> List[int]:                   # :126-128 -> FFTAuxiliary.java:249-302
        """All Lagrange polynomials of the domain at t, on the GPU (one shared inversion per 32 coefficients in place of
        the Java's inversion per coefficient)."""
        import torch
        n = self.domainSize
        d = torch.empty(n * 32, dtype=torch.uint8, device=torch.device("cuda", self.ctx.device))
        self.ctx.fr_lagrange_dev(d, n, _le32(t % FR_MODULUS), _le32(self.omega))
        out = d.cpu().numpy().tobytes()
        return [int.from_bytes(out[32 * i:32 * i + 32], "little") for i in range(n)]

    def computeZ(self, t: int) -> int:                                     # :139-141
        return (pow(t, self.domainSize, FR_MODULUS) - 1) % FR_MODULUS

    def divideByZOnCoset(self, coset: int, input: List[int]) -> None:      # :157-162
        inv = pow(self.computeZ(coset), -1, FR_MODULUS)
        out = self.ctx.fr_scale(b"".join(_le32(v) for v in input), _le32(inv))
        for i in range(self.domainSize):
            input[i] = int.from_bytes(out[32 * i:32 * i + 32], "little")

    radix2_fft = radix2FFT
    radix2_inverse_fft = radix2InverseFFT
    radix2_coset_fft = radix2CosetFFT
    radix2_coset_inverse_fft = radix2CosetInverseFFT
    divide_by_z_on_coset = divideByZOnCoset
    lagrange_coefficients = lagrangeCoefficients
