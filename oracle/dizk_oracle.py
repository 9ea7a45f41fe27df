"""CPU oracle: a Python-int restatement of the reference's (brucechin/OctopusZK, a DIZK fork)
Java algorithms for the Groth16 arithmetic hot path.

THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.  The product path (octopuszk_b200/) never does.

Parity status: the reference stores no BN254 golden vectors.  This oracle is pinned by
the known-answer tests the reference's own unit tests hold, restated in tests/test_oracle.py:
  * MSM over the additive integer group: [3,11,2,8] x [5,2,7,3] = 75, and 4 x (3*5) = 60
    (src/test/java/algebra/msm/SerialVariableBaseMSMTest.java:31-77,
     src/test/java/algebra/msm/DistributedVariableBaseMSMTest.java:92-124)
  * FFT([2,5,3,8]) over LargeFpParameters equals naive evaluation at omega^i
    (src/test/java/algebra/fft/SerialFFTTest.java:168-190)
  * rootOfUnity(8)^8 == 1 (src/test/java/algebra/curves/BNFieldsTest.java:74)
  * group laws (src/test/java/algebra/curves/CurvesTest.java:62-81)
  * distributed FFT == serial FFT (src/test/java/algebra/fft/DistributedFFTTest.java:41-67)
and by the constants of SURVEY.md Appendix B.  The Java itself cannot run here (no JVM in the image).

Every function cites the reference file:line it restates (paths relative to the reference root,
Java sources under src/main/java/).
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple

# --------------------------------------------------------------------------------------------
# Parameters
# --------------------------------------------------------------------------------------------
# algebra/curves/barreto_naehrig/bn254a/bn254a_parameters/BN254aFqParameters.java:33
P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
# .../BN254aFrParameters.java:33-35,39
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FR_ROOT = 19103219067921713944291392827692070036145651957329286315305642004821462161904
FR_MULT_GEN = 5
FR_TWO_ADICITY = 28
# .../BN254aFq2Parameters.java:38 : Fq2 = Fq[u]/(u^2 - nonresidue), nonresidue = p - 1
FQ2_NONRESIDUE = P - 1
# algebra/curves/barreto_naehrig/bn254a/BN254aPublicParameters.java:24-26
COEFF_B = 3
# algebra/fields/fieldparameters/LargeFpParameters.java:33,41 (the field SerialFFTTest uses)
LARGE_FP_MODULUS = 1532495540865888858358347027150309183618765510462668801
LARGE_FP_ROOT = 6


# --------------------------------------------------------------------------------------------
# java.util.Random (JDK LCG) -- needed because Configuration seeds every random() with 10
# (configuration/Configuration.java:52) and Fp.random(seed) = new Random(seed).nextLong()
# (algebra/fields/Fp.java:72-80).
# --------------------------------------------------------------------------------------------
class JavaRandom:
    _MULT = 0x5DEECE66D
    _MASK = (1 << 48) - 1

    def __init__(self, seed: int):
        self.seed = (seed ^ self._MULT) & self._MASK

    def _next(self, bits: int) -> int:
        self.seed = (self.seed * self._MULT + 0xB) & self._MASK
        v = self.seed >> (48 - bits)
        if v >= 1 << (bits - 1):  # to signed int32
            v -= 1 << bits
        return v

    def next_long(self) -> int:
        hi = self._next(32)
        lo = self._next(32)
        v = (hi << 32) + lo
        v &= (1 << 64) - 1
        if v >= 1 << 63:
            v -= 1 << 64
        return v


def fp_random(seed: int, modulus: int) -> int:
    """Fp.random(seed, null): algebra/fields/Fp.java:72-80 with Fp(long) reducing mod p (Fp.java:21-24)."""
    return JavaRandom(seed).next_long() % modulus


# --------------------------------------------------------------------------------------------
# common/MathUtils.java
# --------------------------------------------------------------------------------------------
def java_log2(x: int) -> int:
    """MathUtils.log2: (int)(Math.log(x)/Math.log(2)) -- floating point on purpose (MathUtils.java:12-14)."""
    return int(math.log(x) / math.log(2))


def lowest_power_of_two(n: int) -> int:
    """MathUtils.lowestPowerOfTwo (MathUtils.java:24-33)."""
    if n < 1:
        return 1
    r = 1
    while r < n:
        r <<= 1
    return r


def bitreverse(n: int, bits: int) -> int:
    """MathUtils.bitreverse (MathUtils.java:47-56)."""
    count = bits - 1
    reverse = n
    n >>= 1
    while n > 0:
        reverse = (reverse << 1) | (n & 1)
        n >>= 1
        count -= 1
    return (reverse << count) & ((1 << bits) - 1)


# --------------------------------------------------------------------------------------------
# Fields.  Fp ops reduce after every operation (algebra/fields/Fp.java:21-24,38-49).
# --------------------------------------------------------------------------------------------
class FqField:
    """BN254a Fq as plain ints (algebra/fields/Fp.java)."""
    zero = 0
    one = 1

    @staticmethod
    def add(a, b): return (a + b) % P
    @staticmethod
    def sub(a, b): return (a - b) % P
    @staticmethod
    def mul(a, b): return (a * b) % P
    @staticmethod
    def sqr(a): return (a * a) % P
    @staticmethod
    def neg(a): return (-a) % P
    @staticmethod
    def inv(a): return pow(a, -1, P)          # Fp.inverse = BigInteger.modInverse (Fp.java:88-90)
    @staticmethod
    def is_zero(a): return a == 0
    @staticmethod
    def eq(a, b): return a == b
    @staticmethod
    def bit_size(a): return a.bit_length()     # Fp.bitSize (Fp.java:104-106)


class Fq2Field:
    """BN254a Fq2 = Fq[u]/(u^2+1) as (c0, c1) tuples (algebra/fields/Fp2.java)."""
    zero = (0, 0)
    one = (1, 0)

    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)            # Fp2.java:44-46
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)            # Fp2.java:48-50

    @staticmethod
    def mul(a, b):                                                           # Fp2.java:59-72 (Karatsuba)
        c0c0 = (a[0] * b[0]) % P
        c1c1 = (a[1] * b[1]) % P
        return ((c0c0 + FQ2_NONRESIDUE * c1c1) % P,
                ((a[0] + a[1]) * (b[0] + b[1]) - c0c0 - c1c1) % P)

    @staticmethod
    def sqr(a):                                                              # Fp2.java:98-107 (complex squaring)
        c0c1 = (a[0] * a[1]) % P
        factor = ((a[0] + a[1]) * (a[0] + FQ2_NONRESIDUE * a[1])) % P
        return ((factor - c0c1 - FQ2_NONRESIDUE * c0c1) % P, (2 * c0c1) % P)

    @staticmethod
    def neg(a): return ((-a[0]) % P, (-a[1]) % P)

    @staticmethod
    def inv(a):                                                              # Fp2.java:109-118
        t2 = (a[0] * a[0] - FQ2_NONRESIDUE * a[1] * a[1]) % P
        t3 = pow(t2, -1, P)
        return ((a[0] * t3) % P, (-(a[1] * t3)) % P)

    @staticmethod
    def is_zero(a): return a[0] == 0 and a[1] == 0
    @staticmethod
    def eq(a, b): return a[0] == b[0] and a[1] == b[1]
    @staticmethod
    def bit_size(a): return max(a[0].bit_length(), a[1].bit_length())


# --------------------------------------------------------------------------------------------
# Groups
# --------------------------------------------------------------------------------------------
class JacobianGroup:
    """Jacobian short-Weierstrass group y^2 = x^3 + b over field F, formula-for-formula from
    algebra/curves/barreto_naehrig/BNG1.java:38-97,133-172,191-224 (BNG2.java:43-126 is identical over Fq2).
    Points are (X, Y, Z) tuples of field elements; infinity is any triple with Z == 0."""

    def __init__(self, F, generator, zero, name):
        self.F = F
        self.generator = generator
        self._zero = zero
        self.name = name

    def zero(self):
        return self._zero

    def is_zero(self, p):                       # BNG1.java:103-105
        return self.F.is_zero(p[2])

    def twice(self, p):                         # BNG1.java:133-161 (dbl-2009-l)
        F = self.F
        if self.is_zero(p):
            return p
        X1, Y1, Z1 = p
        A = F.sqr(X1)
        B = F.sqr(Y1)
        C = F.sqr(B)
        D = F.sub(F.sub(F.sqr(F.add(X1, B)), A), C)
        D = F.add(D, D)
        E = F.add(F.add(A, A), A)
        Fv = F.sqr(E)
        X3 = F.sub(Fv, F.add(D, D))
        eightC = F.add(C, C)
        eightC = F.add(eightC, eightC)
        eightC = F.add(eightC, eightC)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), eightC)
        Y1Z1 = F.mul(Y1, Z1)
        Z3 = F.add(Y1Z1, Y1Z1)
        return (X3, Y3, Z3)

    def add(self, p, q):                        # BNG1.java:38-97 (add-2007-bl with explicit O and P==Q checks)
        F = self.F
        if self.is_zero(p):
            return q
        if self.is_zero(q):
            return p
        X1, Y1, Z1 = p
        X2, Y2, Z2 = q
        Z1Z1 = F.sqr(Z1)
        Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        Z1c = F.mul(Z1, Z1Z1)
        Z2c = F.mul(Z2, Z2Z2)
        S1 = F.mul(Y1, Z2c)
        S2 = F.mul(Y2, Z1c)
        if F.eq(U1, U2) and F.eq(S1, S2):
            return self.twice(p)
        H = F.sub(U2, U1)
        S2mS1 = F.sub(S2, S1)
        I = F.sqr(F.add(H, H))
        J = F.mul(H, I)
        r = F.add(S2mS1, S2mS1)
        V = F.mul(U1, I)
        X3 = F.sub(F.sub(F.sqr(r), J), F.add(V, V))
        S1J = F.mul(S1, J)
        Y3 = F.sub(F.mul(r, F.sub(V, X3)), F.add(S1J, S1J))
        Z3 = F.mul(F.sub(F.sub(F.sqr(F.add(Z1, Z2)), Z1Z1), Z2Z2), H)
        return (X3, Y3, Z3)

    def negate(self, p):                        # BNG1.java:129-131
        return (p[0], self.F.neg(p[1]), p[2])

    def sub(self, p, q):                        # BNG1.java:99-101
        return self.add(p, self.negate(q))

    def mul(self, p, scalar: int):              # algebra/groups/AbstractGroup.java:29-51 (MSB-first double-and-add)
        if scalar == 1:
            return p
        result = self.zero()
        found = False
        for i in range(scalar.bit_length() - 1, -1, -1):
            if found:
                result = self.twice(result)
            if (scalar >> i) & 1:
                found = True
                result = self.add(result, p)
        return result

    def to_affine(self, p):                     # BNG1.java:163-172
        F = self.F
        if self.is_zero(p):
            return (F.zero, F.one, F.zero)
        zi = F.inv(p[2])
        z2 = F.sqr(zi)
        z3 = F.mul(z2, zi)
        return (F.mul(p[0], z2), F.mul(p[1], z3), F.one)

    def equals(self, p, q):                     # BNG1.java:191-224 (projective equality)
        F = self.F
        if self.is_zero(p):
            return self.is_zero(q)
        if self.is_zero(q):
            return False
        z1s = F.sqr(p[2])
        z2s = F.sqr(q[2])
        if not F.eq(F.mul(p[0], z2s), F.mul(q[0], z1s)):
            return False
        z1c = F.mul(p[2], z1s)
        z2c = F.mul(q[2], z2s)
        return F.eq(F.mul(p[1], z2c), F.mul(q[1], z1c))

    def bit_size(self, p):                      # BNG1.java:174-176
        F = self.F
        return max(F.bit_size(p[0]), F.bit_size(p[1]), F.bit_size(p[2]))

    def random(self, seed: int):                # BNG1.java:125-127 : one().mul(Fr.random(seed))
        return self.mul(self.generator, fp_random(seed, R))


# BN254aG1Parameters.java:23-24,52-55 : ZERO = (0,1,0), ONE = (1,2,1)
G1 = JacobianGroup(FqField, (1, 2, 1), (0, 1, 0), "G1")
# BN254aG2Parameters.java:25-32 (ONE), :60-68 (ZERO() = (0,0,0) in this fork)
G2 = JacobianGroup(
    Fq2Field,
    ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
      11559732032986387107991004021392285783925812861821192530917403151452391805634),
     (8495653923123431417604973247489272438418190587263600148770280649306958101930,
      4082367875863433681332203403145435568316851327593401208105741076214120093531),
     (1, 0)),
    ((0, 0), (0, 0), (0, 0)), "G2")


class AdditiveIntegerGroup:
    """algebra/groups/AdditiveIntegerGroup.java -- the group the reference's MSM unit tests run on."""
    name = "Z"

    def zero(self): return 0
    def is_zero(self, p): return p == 0
    def add(self, p, q): return p + q
    def twice(self, p): return p + p
    def negate(self, p): return -p
    def sub(self, p, q): return p - q
    def mul(self, p, s): return p * s
    def equals(self, p, q): return p == q


ZGROUP = AdditiveIntegerGroup()


# --------------------------------------------------------------------------------------------
# Variable-base MSM
# --------------------------------------------------------------------------------------------
def naive_msm(group, scalars: Sequence[int], bases: Sequence) -> object:
    """algebra/msm/NaiveMSM.java:21-46 : sum of base.mul(scalar)."""
    acc = group.zero()
    for s, b in zip(scalars, bases):
        acc = group.add(acc, group.mul(b, s))
    return acc


def pippenger_window(n: int) -> int:
    """c = L - floor(L/3), L = max(1, floor(log2 n)) (algebra/msm/VariableBaseMSM.java:137-139;
    same rule on the GPU side, algebra_msm_VariableBaseMSM.cu:1269-1272)."""
    L = max(1, java_log2(n))
    return L - (L // 3)


def pippenger_msm(group, scalars: Sequence[int], bases: Sequence, num_bits: int = 254):
    """algebra/msm/VariableBaseMSM.java:134-188 (pippengerMSM): unsigned c-bit windows from the top,
    bucket 0 skipped, running-sum bucket reduce (:171-177), c doublings between windows (:180-184).
    num_bits defaults to the 254 the native path hard-codes (algebra_msm_VariableBaseMSM.cu:1267)."""
    length = len(scalars)
    c = pippenger_window(length)
    num_buckets = 1 << c
    num_groups = (num_bits + c - 1) // c
    zero = group.zero()
    result = zero
    for k in range(num_groups - 1, -1, -1):
        buckets = [zero] * num_buckets
        for i in range(length):
            idx = (scalars[i] >> (k * c)) & (num_buckets - 1)
            if idx == 0:
                continue
            buckets[idx] = group.add(buckets[idx], bases[i])
        running = zero
        for i in range(num_buckets - 1, 0, -1):
            running = group.add(running, buckets[i])
            result = group.add(result, running)
        if k > 0:
            for _ in range(c):
                result = group.twice(result)
    return result


def serial_msm(group, scalars: Sequence[int], bases: Sequence, chunk: int | None = None):
    """VariableBaseMSM.serialMSM (VariableBaseMSM.java:199-338): split in chunks of 2^23 (G1) / 2^22 (G2),
    run the (native) Pippenger on each, add the chunk results (:261-265)."""
    if chunk is None:
        chunk = 1 << 23 if group is G1 else 1 << 22
    acc = group.zero()
    for lo in range(0, len(scalars), chunk):
        acc = group.add(acc, pippenger_msm(group, scalars[lo:lo + chunk], bases[lo:lo + chunk]))
    return acc


def double_msm(scalars: Sequence[int], bases1: Sequence, bases2: Sequence):
    """VariableBaseMSM.doubleMSM (VariableBaseMSM.java:480-606): same scalars on G1 and G2 bases."""
    return serial_msm(G1, scalars, bases1, 1 << 21), serial_msm(G2, scalars, bases2, 1 << 21)


# --------------------------------------------------------------------------------------------
# Fixed-base MSM
# --------------------------------------------------------------------------------------------
# BN254aG1Parameters.java:25-50 and BN254aG2Parameters.java:33-58
G1_FIXED_BASE_WINDOW_TABLE = [1, 5, 11, 32, 55, 162, 360, 815, 2373, 6978, 7122, 0, 57818, 0, 169679, 439759,
                              936073, 0, 4666555, 7580404, 0, 34552892]
G2_FIXED_BASE_WINDOW_TABLE = [1, 5, 10, 25, 59, 154, 334, 743, 2034, 4988, 8888, 26271, 39768, 106276, 141703,
                              462423, 926872, 0, 4873049, 5706708, 0, 31673815]


def get_window_size(num_scalars: int, group) -> int:
    """FixedBaseMSM.getWindowSize (algebra/msm/FixedBaseMSM.java:49-66)."""
    table = G1_FIXED_BASE_WINDOW_TABLE if group is G1 else G2_FIXED_BASE_WINDOW_TABLE
    window = 1
    for i in range(len(table) - 1, -1, -1):
        if table[i] != 0 and num_scalars >= table[i]:
            window = i + 1
            break
    return window


def get_window_table(group, base, scalar_size: int, window_size: int):
    """FixedBaseMSM.getWindowTable (FixedBaseMSM.java:71-99): table[outer][inner] = inner * 2^(w*outer) * base."""
    num_windows = scalar_size // window_size if scalar_size % window_size == 0 else scalar_size // window_size + 1
    inner_limit = 1 << window_size
    if num_windows == 0:
        return [[group.zero()]]
    table = []
    base_outer = base
    for _ in range(num_windows):
        row = []
        base_inner = group.zero()
        for _ in range(inner_limit):
            row.append(base_inner)
            base_inner = group.add(base_inner, base_outer)
        table.append(row)
        for _ in range(window_size):
            base_outer = group.twice(base_outer)
    return table


def fixed_serial_msm(group, scalar_size: int, window_size: int, table, scalar: int):
    """FixedBaseMSM.serialMSM (FixedBaseMSM.java:141-167)."""
    outerc = (scalar_size + window_size - 1) // window_size
    res = table[0][0]
    for outer in range(outerc):
        inner = (scalar >> (outer * window_size)) & ((1 << window_size) - 1)
        res = group.add(res, table[outer][inner])
    return res


def fixed_batch_msm(group, scalar_size: int, window_size: int, base, scalars: Sequence[int]):
    """What FixedBaseMSM.batchMSM (FixedBaseMSM.java:185-315) returns as group elements:
    out[i] = (s_i mod 2^(outerc*w)) * base with outerc = ceil(scalarSize/w) (:214; the native side walks
    exactly outerc windows, algebra_msm_FixedBaseMSM.cu:750-790).  Computed by double-and-add so that large
    windows do not need the 2^w-entry table in Python."""
    outerc = (scalar_size + window_size - 1) // window_size
    mask = (1 << (outerc * window_size)) - 1
    return [group.mul(base, s & mask) for s in scalars]


def field_batch_msm(scalars: Sequence[int], b: int) -> List[int]:
    """fieldBatchMSMNativeHelper / field_MSM (algebra_msm_FixedBaseMSM.cu:1241-1266): a_i * b mod r."""
    return [(a * b) % R for a in scalars]


# --------------------------------------------------------------------------------------------
# FFT
# --------------------------------------------------------------------------------------------
def root_of_unity(order: int, modulus: int = R, root: int = FR_ROOT) -> int:
    """Fp.rootOfUnity (algebra/fields/Fp.java:98-102): root^(modulus / order) with integer division."""
    return pow(root, modulus // order, modulus)


def serial_radix2_fft(a: List[int], omega: int, modulus: int = R) -> None:
    """FFTAuxiliary.serialRadix2FFT, the active pure-Java loop (algebra/fft/FFTAuxiliary.java:100-123):
    bit-reverse swap, then log n DIT stages.  In place; natural order in and out."""
    n = len(a)
    if n == 1:
        return
    logn = java_log2(n)
    assert n == 1 << logn
    for k in range(n):
        rk = bitreverse(k, logn)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    m = 1
    for _ in range(1, logn + 1):
        w_m = pow(omega, n // (2 * m), modulus)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = (w * a[k + j + m]) % modulus
                a[k + j + m] = (a[k + j] - t) % modulus
                a[k + j] = (a[k + j] + t) % modulus
                w = (w * w_m) % modulus
        m *= 2


def naive_evaluate(a: Sequence[int], x: int, modulus: int = R) -> int:
    """common/NaiveEvaluation.evaluatePolynomial (Horner) -- what SerialFFTTest compares against."""
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % modulus
    return acc


def multiply_by_coset(a: List[int], g: int, modulus: int = R) -> None:
    """FFTAuxiliary.multiplyByCoset (FFTAuxiliary.java:224-232): a[i] *= g^i."""
    coset = g
    for i in range(1, len(a)):
        a[i] = (a[i] * coset) % modulus
        coset = (coset * g) % modulus


class SerialFFT:
    """algebra/fft/SerialFFT.java."""

    def __init__(self, domain_size: int, modulus: int = R, root: int = FR_ROOT):
        self.modulus = modulus
        self.domain_size = lowest_power_of_two(domain_size)           # SerialFFT.java:24-28
        self.omega = root_of_unity(self.domain_size, modulus, root)

    def radix2_fft(self, a: List[int]) -> None:                       # :75-80
        assert len(a) == self.domain_size
        serial_radix2_fft(a, self.omega, self.modulus)

    def radix2_inverse_fft(self, a: List[int]) -> None:               # :86-95
        assert len(a) == self.domain_size
        serial_radix2_fft(a, pow(self.omega, -1, self.modulus), self.modulus)
        c = pow(self.domain_size, -1, self.modulus)
        for i in range(self.domain_size):
            a[i] = (a[i] * c) % self.modulus

    def radix2_coset_fft(self, a: List[int], g: int) -> None:         # :100-105
        multiply_by_coset(a, g, self.modulus)
        self.radix2_fft(a)

    def radix2_coset_inverse_fft(self, a: List[int], g: int) -> None:  # :111-115
        self.radix2_inverse_fft(a)
        multiply_by_coset(a, pow(g, -1, self.modulus), self.modulus)

    def compute_z(self, t: int) -> int:                               # :139-141
        return (pow(t, self.domain_size, self.modulus) - 1) % self.modulus

    def divide_by_z_on_coset(self, coset: int, a: List[int]) -> None:  # :157-162
        inv = pow(self.compute_z(coset), -1, self.modulus)
        for i in range(self.domain_size):
            a[i] = (a[i] * inv) % self.modulus

    def lagrange_coefficients(self, t: int) -> List[int]:            # :127-129 -> FFTAuxiliary.java:249-302
        return serial_radix2_lagrange_coefficients(t, self.domain_size, self.modulus,
                                                   root_of_unity(self.domain_size, self.modulus,
                                                                 FR_ROOT if self.modulus == R else LARGE_FP_ROOT))


def serial_radix2_lagrange_coefficients(t: int, m: int, modulus: int, omega: int) -> List[int]:
    """FFTAuxiliary.serialRadix2LagrangeCoefficients (FFTAuxiliary.java:249-302)."""
    if m == 1:
        return [1]
    coeffs = [0] * m
    if pow(t, m, modulus) == 1:
        omega_i = 1
        for i in range(m):
            if omega_i == t:
                coeffs[i] = 1
                return coeffs
            omega_i = (omega_i * omega) % modulus
    Z = (pow(t, m, modulus) - 1) % modulus
    l = (Z * pow(m, -1, modulus)) % modulus
    r = 1
    for i in range(m):
        coeffs[i] = (l * pow((t - r) % modulus, -1, modulus)) % modulus
        l = (l * omega) % modulus
        r = (r * omega) % modulus
    return coeffs


def distributed_radix2_fft(a: Sequence[int], rows: int, columns: int, inverse: bool,
                           modulus: int = R, root: int = FR_ROOT) -> List[int]:
    """FFTAuxiliary.distributedRadix2FFT (FFTAuxiliary.java:129-219), the Spark four-step, restated on a list.
    Input index e -> group e % rows, position e / rows; step 1: length-`columns` (inverse) FFT per group;
    twiddle omega_size^(group*i) (inverse: its inverse, :177-178) and re-key to i*rows+group; step 2:
    length-`rows` (inverse) FFT per i; output index j*columns + i."""
    size = rows * columns
    omega_shift = root_of_unity(size, modulus, root)
    row_dom = SerialFFT(rows, modulus, root)
    col_dom = SerialFFT(columns, modulus, root)
    stage = [0] * size
    for g in range(rows):
        grp = [a[k * rows + g] for k in range(columns)]
        if columns > 1:
            (col_dom.radix2_inverse_fft if inverse else col_dom.radix2_fft)(grp)
        for i in range(columns):
            nth = pow(omega_shift, g * i, modulus)
            if inverse:
                nth = pow(nth, -1, modulus)
            stage[i * rows + g] = (nth * grp[i]) % modulus
    out = [0] * size
    for i in range(columns):
        grp = [stage[i * rows + g] for g in range(rows)]
        if rows > 1:
            (row_dom.radix2_inverse_fft if inverse else row_dom.radix2_fft)(grp)
        for j in range(rows):
            out[j * columns + i] = grp[j]
    return out


# --------------------------------------------------------------------------------------------
# Wire formats
# --------------------------------------------------------------------------------------------
def le32(x: int) -> bytes:
    """bigIntegerToByteArrayHelperCGBN (VariableBaseMSM.java:121-131): 32-byte little-endian, zero padded."""
    return x.to_bytes(32, "little")


def from_le(b: bytes) -> int:
    return int.from_bytes(b, "little")


def pack_scalars(scalars: Sequence[int]) -> bytes:
    return b"".join(le32(s) for s in scalars)


def pack_g1(points: Sequence[Tuple[int, int, int]]) -> bytes:
    """N x [X|Y|Z], 96 B per point (VariableBaseMSM.java:224-227)."""
    return b"".join(le32(p[0]) + le32(p[1]) + le32(p[2]) for p in points)


def pack_g2(points) -> bytes:
    """N x [X.c0|X.c1|Y.c0|Y.c1|Z.c0|Z.c1], 192 B per point (VariableBaseMSM.java:279-285)."""
    return b"".join(le32(p[0][0]) + le32(p[0][1]) + le32(p[1][0]) + le32(p[1][1]) + le32(p[2][0]) + le32(p[2][1])
                    for p in points)


def unpack_g1(b: bytes, stride: int = 32, big_endian: bool = False) -> List[Tuple[int, int, int]]:
    """Inverse of pack_g1 for `stride`-byte coordinates (32 = C ABI, 64 = legacy JNI return format)."""
    order = "big" if big_endian else "little"
    out = []
    for o in range(0, len(b), 3 * stride):
        out.append(tuple(int.from_bytes(b[o + k * stride:o + (k + 1) * stride], order) for k in range(3)))
    return out


def unpack_g2(b: bytes, stride: int = 32, big_endian: bool = False):
    order = "big" if big_endian else "little"
    out = []
    for o in range(0, len(b), 6 * stride):
        v = [int.from_bytes(b[o + k * stride:o + (k + 1) * stride], order) for k in range(6)]
        out.append(((v[0], v[1]), (v[2], v[3]), (v[4], v[5])))
    return out
