"""GPU parity: radix-2 NTT over Fr and the Fr vector scale, through the C ABI, against the oracle
(FFTAuxiliary.serialRadix2FFT restated in oracle/dizk_oracle.py)."""
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


def _pack(v):
    return b"".join(O.le32(x) for x in v)


def _unpack(b):
    return [O.from_le(b[i:i + 32]) for i in range(0, len(b), 32)]


def test_fr_scale_matches_oracle(ctx):
    rng = random.Random(11)
    a = [0, 1, O.R - 1, O.R - 2, (1 << 253)] + [rng.randrange(O.R) for _ in range(1000)]
    for b in (0, 1, O.R - 1, rng.randrange(O.R)):
        got = _unpack(ctx.fr_scale(_pack(a), O.le32(b)))
        assert got == O.field_batch_msm(a, b)


def test_fr_scale_rejects_unreduced(ctx):
    from octopuszk_b200 import OzkError
    with pytest.raises(OzkError):
        ctx.fr_scale(_pack([1, 2]), O.le32(O.R))


def test_ntt_reference_kat(ctx):
    # SURVEY.md Appendix B / SerialFFTTest.java:168-190 carried to Fr
    got = _unpack(ctx.ntt(_pack([2, 5, 3, 8]), O.le32(O.root_of_unity(4))))
    assert got == [18, 21888242871839275209022642834368543560924422484752198131886912786320552141471,
                   21888242871839275222246405745257275088548364400416034343698204186575808495609,
                   13223762910888731527623941915663836211811291400255256354144]


@pytest.mark.parametrize("log_n", list(range(0, 14)))
def test_ntt_matches_oracle_all_small_sizes(ctx, log_n):
    rng = random.Random(100 + log_n)
    n = 1 << log_n
    x = [rng.randrange(O.R) for _ in range(n)]
    if n >= 4:
        x[0], x[1], x[2] = 0, O.R - 1, 1
    omega = O.root_of_unity(n)
    exp = list(x)
    O.serial_radix2_fft(exp, omega)
    assert _unpack(ctx.ntt(_pack(x), O.le32(omega))) == exp
    # inverse direction: omega^-1 (SerialFFT.radix2InverseFFT without the n^-1 scaling)
    inv = pow(omega, -1, O.R)
    exp2 = list(x)
    O.serial_radix2_fft(exp2, inv)
    assert _unpack(ctx.ntt(_pack(x), O.le32(inv))) == exp2


@pytest.mark.parametrize("log_n", [16, 18, 19, 20, 22])
def test_ntt_large_spot_check_and_roundtrip(ctx, log_n):
    """Sizes the Python oracle cannot transform in seconds: compare sampled outputs with Horner evaluation of the
    input at omega^k (the definition FFTAuxiliary.serialRadix2FFT satisfies), then invert and compare everything."""
    import numpy as np
    rng = random.Random(200 + log_n)
    n = 1 << log_n
    from tests import util
    x_bytes = util.rand_scalars_full_range(n, seed=200 + log_n).tobytes()      # uniform in [0, r)
    omega = O.root_of_unity(n)
    y_bytes = ctx.ntt(x_bytes, O.le32(omega))
    ks = [0, 1, n - 1, n // 2] + [rng.randrange(n) for _ in range(2)]
    if log_n <= 18:
        x = _unpack(x_bytes)
        for k in ks:
            assert O.from_le(y_bytes[32 * k:32 * k + 32]) == O.naive_evaluate(x, pow(omega, k, O.R))
    z_bytes = ctx.ntt(y_bytes, O.le32(pow(omega, -1, O.R)))
    ninv = pow(n, -1, O.R)
    back = ctx.fr_scale(z_bytes, O.le32(ninv))
    assert back == x_bytes
    # linearity / delta response, size independent: NTT(e_j)[k] = omega^(jk)
    j = rng.randrange(n)
    d = bytearray(n * 32)
    d[32 * j:32 * j + 32] = O.le32(1)
    yd = ctx.ntt(bytes(d), O.le32(omega))
    for k in ks:
        assert O.from_le(yd[32 * k:32 * k + 32]) == pow(omega, j * k, O.R)


@pytest.mark.parametrize("log_n", [19, 20, 22])
def test_ntt_three_pass_full_compare_c_oracle(ctx, log_n):
    """The three-pass sizes (2^19 and up; the headline 2^26 runs the same kernels) compared ELEMENT FOR ELEMENT with the C
    restatement of FFTAuxiliary.serialRadix2FFT (oracle/dizk_oracle.c, itself checked against the Python oracle in
    tests/test_oracle_c.py): forward, inverse (SerialFFT.radix2InverseFFT) and coset forward (radix2CosetFFT) on inputs
    uniform in [0, r) with the edge values 0, 1, r - 1 forced in (full compare up to 2^22)."""
    import numpy as np
    import torch
    from oracle import c_oracle as C
    from tests import util
    n = 1 << log_n
    raw = util.rand_scalars_full_range(n, seed=300 + log_n)
    for pos, v in enumerate([0, 1, O.R - 1, O.R - 2, 1 << 253]):
        raw[pos] = np.frombuffer(O.le32(v), dtype=np.uint8)
    x_bytes = raw.tobytes()
    omega = O.root_of_unity(n)
    w, winv = O.le32(omega), O.le32(pow(omega, -1, O.R))
    g = O.FR_MULT_GEN
    d_in = torch.from_numpy(raw).cuda()
    d_out = torch.empty_like(d_in)

    def gpu(**kw):
        ctx.ntt_ex_dev(d_in, d_out, n, **kw)
        ctx.sync()
        return d_out.cpu().numpy().tobytes()

    assert gpu(omega=w) == C.fft_fr(x_bytes, w), "forward"
    exp = C.fr_scale(C.fft_fr(x_bytes, winv), O.le32(pow(n, -1, O.R)))
    assert gpu(omega=winv, post_scale=O.le32(pow(n, -1, O.R))) == exp, "inverse"
    assert gpu(omega=w, pre_coset=O.le32(g)) == C.fft_fr(C.fr_coset_scale(x_bytes, g), w), "coset forward"
    # the host-pointer entry point (what the JNI shim calls) on the same input
    assert ctx.ntt(x_bytes, w) == C.fft_fr(x_bytes, w)


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("log_n", [24, 26])
def test_ntt_headline_sizes_horner_spots(ctx, log_n):
    """2^24 and the headline 2^26 on inputs uniform in [0, r): K = 64 outputs (fixed corner indices and random ones) against
    the definition out[k] = sum_j in[j] omega^(jk), evaluated by the C oracle's Horner loop over the whole input
    (SerialFFTTest.java:168-190 compares with naive evaluation)."""
    import torch
    from oracle import c_oracle as C
    from tests import util
    n = 1 << log_n
    raw = util.rand_scalars_full_range(n, seed=400 + log_n)
    omega = O.root_of_unity(n)
    d_in = torch.from_numpy(raw).cuda()
    d_out = torch.empty_like(d_in)
    ctx.ntt_dev(d_in, d_out, n, O.le32(omega))
    ctx.sync()
    rng = random.Random(log_n)
    ks = [0, 1, 2, n - 1, n // 2, n // 2 + 1, (1 << 9) - 1, 1 << 9, 1 << 18, (1 << 18) + 1] + [rng.randrange(n) for _ in range(54)]
    got = [O.from_le(d_out[k].cpu().numpy().tobytes()) for k in ks]
    exp = C.fr_horner(raw, n, [pow(omega, k, O.R) for k in ks])
    assert got == exp


def test_ntt_rejects_bad_omega(ctx):
    from octopuszk_b200 import OzkError
    with pytest.raises(OzkError):
        ctx.ntt(_pack([1, 2, 3, 4]), O.le32(O.root_of_unity(8)))     # primitive 8th root, not 4th
    with pytest.raises(OzkError):
        ctx.ntt(_pack([1, 2, 3]), O.le32(1))                          # not a power of two


def test_ntt_ex_wrappers_match_serialfft(ctx):
    """radix2InverseFFT / radix2CosetFFT / radix2CosetInverseFFT (SerialFFT.java:86-115) via the fused entry point."""
    import torch
    rng = random.Random(5)
    n = 1 << 10
    dom = O.SerialFFT(n)
    x = [rng.randrange(O.R) for _ in range(n)]
    g = O.FR_MULT_GEN
    dev = torch.device("cuda:0")
    d_in = torch.frombuffer(bytearray(_pack(x)), dtype=torch.uint8).to(dev)
    d_out = torch.empty_like(d_in)
    ninv = pow(n, -1, O.R)

    def run(**kw):
        ctx.ntt_ex_dev(d_in, d_out, n, **kw)
        ctx.sync()
        return _unpack(d_out.cpu().numpy().tobytes())

    e = list(x); dom.radix2_inverse_fft(e)
    assert run(omega=O.le32(pow(dom.omega, -1, O.R)), post_scale=O.le32(ninv)) == e
    e = list(x); dom.radix2_coset_fft(e, g)
    assert run(omega=O.le32(dom.omega), pre_coset=O.le32(g)) == e
    e = list(x); dom.radix2_coset_inverse_fft(e, g)
    assert run(omega=O.le32(pow(dom.omega, -1, O.R)), post_scale=O.le32(ninv), post_coset=O.le32(pow(g, -1, O.R))) == e


@pytest.mark.parametrize("m", [1, 2, 8, 32, 64, 1000, 1 << 12])
def test_lagrange_coefficients_match_oracle(ctx, m):
    """ozk_fr_lagrange_dev against FFTAuxiliary.serialRadix2LagrangeCoefficients restated (FFTAuxiliary.java:249-302;
    the reference's own test is DistributedFFTTest/SerialFFTTest "Lagrange" against naive interpolation): ordinary t, the
    seed-10 value the setup uses, t inside the domain (unit vector), and ragged sizes of the last thread's batch."""
    import torch
    from octopuszk_b200 import OzkError
    if m & (m - 1):
        with pytest.raises(OzkError):
            ctx.fr_lagrange_dev(torch.empty(32 * m, dtype=torch.uint8, device="cuda"), m, O.le32(5), O.le32(1))
        return
    omega = O.root_of_unity(m) if m > 1 else 1
    rng = random.Random(m)
    ts = [rng.randrange(O.R), O.fp_random(10, O.R), 0, 1]
    if m > 2:
        ts += [pow(omega, m - 1, O.R), pow(omega, m // 2 + 1, O.R)]
    for t in ts:
        d = torch.empty(32 * m, dtype=torch.uint8, device="cuda")
        ctx.fr_lagrange_dev(d, m, O.le32(t), O.le32(omega))
        got = d.cpu().numpy().tobytes()
        got = [O.from_le(got[32 * i:32 * i + 32]) for i in range(m)]
        assert got == O.serial_radix2_lagrange_coefficients(t, m, O.R, omega), (m, t)
    if m > 2:
        with pytest.raises(OzkError):
            ctx.fr_lagrange_dev(torch.empty(32 * m, dtype=torch.uint8, device="cuda"), m, O.le32(5), O.le32(pow(omega, 2, O.R)))


def test_lagrange_large_interpolates(ctx):
    """2^20 coefficients: sum_i L_i(t) = 1 and sum_i L_i(t) omega^(i k) = t^k (interpolation of x^k), exact in Fr."""
    import numpy as np
    import torch
    m = 1 << 20
    omega = O.root_of_unity(m)
    t = O.fp_random(10, O.R)
    d = torch.empty(32 * m, dtype=torch.uint8, device="cuda")
    ctx.fr_lagrange_dev(d, m, O.le32(t), O.le32(omega))
    raw = d.cpu().numpy().reshape(m, 32)
    from tests import util
    assert util.column_sums(raw, 1)[0] % O.R == 1
    # k = m / 4: omega^(i k) cycles through the four 4th roots of unity, so the sum splits into four column sums
    w4 = pow(omega, m // 4, O.R)
    cols = util.column_sums(raw, 4)
    assert sum(c * pow(w4, j, O.R) for j, c in enumerate(cols)) % O.R == pow(t, m // 4, O.R)


@pytest.mark.parametrize("log_n", [24, 26])
def test_ntt_headline_sizes_device_resident(ctx, log_n):
    """BASELINE.json's NTT sizes (2^24, and the headline 2^26) through size-independent exact properties, all on the device:
    forward then inverse transform times n^-1 gives the input back bit for bit; the transform of x plus a unit vector e_j
    differs from the transform of x by exactly omega^(jk) at sampled k (linearity + delta response)."""
    import torch
    n = 1 << log_n
    g = torch.Generator(device="cuda")
    g.manual_seed(log_n)
    x = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    x[:, 31] &= 0x0F                                  # < 2^252: x + e_j stays canonical without a reduction
    x[:, 0] &= 0xFE                                   # low bit clear: adding 1 to element j cannot carry
    omega = O.root_of_unity(n)
    y = torch.empty_like(x)
    ctx.ntt_dev(x, y, n, O.le32(omega))
    back = torch.empty_like(x)
    ctx.ntt_ex_dev(y, back, n, O.le32(pow(omega, -1, O.R)), None, O.le32(pow(n, -1, O.R)), None)
    ctx.sync()
    assert torch.equal(back, x)
    rng = random.Random(log_n)
    j = rng.randrange(n)
    x2 = x.clone()
    x2[j, 0] |= 1                                     # x2 = x + e_j
    y2 = torch.empty_like(x)
    ctx.ntt_dev(x2, y2, n, O.le32(omega))
    ctx.sync()
    for k in [0, 1, n - 1, n // 2, j] + [rng.randrange(n) for _ in range(8)]:
        a = O.from_le(y[k].cpu().numpy().tobytes())
        b = O.from_le(y2[k].cpu().numpy().tobytes())
        assert (b - a) % O.R == pow(omega, j * k, O.R), (log_n, k)


@pytest.mark.parametrize("log_n", [27, 28])
def test_ntt_beyond_headline_direct_tables(ctx, log_n):
    """2^27 and 2^28 (the 2-adicity of Fr): on a device with memory to spare the first pass reads a direct inter-pass twiddle table
    of 2^27 / 2^28 entries (ntt.cu, direct_log_limit).  Same exact properties as above, all on the device -- forward, inverse and
    n^-1 give the input back bit for bit; the response to a unit vector e_j is omega^(jk) at sampled k -- and at 2^27 three outputs
    against Horner evaluation of the whole input by the C oracle."""
    import torch
    n = 1 << log_n
    torch.cuda.empty_cache()                          # blocks cached by earlier tests count as used otherwise
    free, _total = torch.cuda.mem_get_info()
    if free < 9 * n * 32:
        pytest.skip("not enough free device memory for five 2^%d-element vectors, the scratch and the tables" % log_n)
    g = torch.Generator(device="cuda")
    g.manual_seed(log_n)
    x = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    x[:, 31] &= 0x0F
    x[:, 0] &= 0xFE
    omega = O.root_of_unity(n)
    y = torch.empty_like(x)
    ctx.ntt_dev(x, y, n, O.le32(omega))
    back = torch.empty_like(x)
    ctx.ntt_ex_dev(y, back, n, O.le32(pow(omega, -1, O.R)), None, O.le32(pow(n, -1, O.R)), None)
    ctx.sync()
    assert torch.equal(back, x)
    del back
    rng = random.Random(log_n)
    if log_n == 27:
        from oracle import c_oracle as C
        ks = [1, n // 2 + 3, rng.randrange(n)]
        exp = C.fr_horner(x.cpu().numpy(), n, [pow(omega, k, O.R) for k in ks])
        assert [O.from_le(y[k].cpu().numpy().tobytes()) for k in ks] == exp
    j = rng.randrange(n)
    x[j, 0] |= 1                                      # x + e_j, in place
    y2 = torch.empty_like(x)
    ctx.ntt_dev(x, y2, n, O.le32(omega))
    ctx.sync()
    for k in [0, 1, n - 1, n // 2, j] + [rng.randrange(n) for _ in range(8)]:
        a = O.from_le(y[k].cpu().numpy().tobytes())
        b = O.from_le(y2[k].cpu().numpy().tobytes())
        assert (b - a) % O.R == pow(omega, j * k, O.R), (log_n, k)
    del x, y, y2
    torch.cuda.empty_cache()


def test_sparse_rows_times_assignment(ctx):
    """ozk_fr_spmv_dev against LinearCombination.evaluate restated: empty rows, single terms, rows at and above the
    one-thread limit (64 terms), one row over all variables (the last constraint of R1CSConstruction.serialConstruct)."""
    import numpy as np
    import torch
    rng = random.Random(77)
    nvars = 5000
    z = [rng.randrange(O.R) for _ in range(nvars)]
    z[0] = 1
    lengths = [0, 1, 2, 64, 65, 300, nvars, 3, 0, 1000] + [rng.randrange(0, 5) for _ in range(2000)]
    rows = []
    for L in lengths:
        idx = rng.sample(range(nvars), L) if L < nvars else list(range(nvars))
        rows.append([(i, rng.choice([1, 1, O.R - 1, rng.randrange(O.R)])) for i in idx])
    row_ptr = np.zeros(len(rows) + 1, dtype=np.uint32)
    cols, coeffs = [], []
    for i, lc in enumerate(rows):
        for j, v in lc:
            cols.append(j)
            coeffs.append(v)
        row_ptr[i + 1] = len(cols)
    d_out = torch.empty(len(rows) * 32, dtype=torch.uint8, device="cuda")
    ctx.fr_spmv_dev(torch.from_numpy(row_ptr).cuda(), torch.from_numpy(np.asarray(cols, dtype=np.uint32)).cuda(),
                    torch.frombuffer(bytearray(O.pack_scalars(coeffs)), dtype=torch.uint8).cuda(),
                    torch.frombuffer(bytearray(O.pack_scalars(z)), dtype=torch.uint8).cuda(), len(rows), d_out)
    got = d_out.cpu().numpy().tobytes()
    exp = [sum(v * z[j] for j, v in lc) % O.R for lc in rows]
    assert [O.from_le(got[32 * i:32 * i + 32]) for i in range(len(rows))] == exp
