"""GPU parity: variable-base MSM (G1, G2, paired) through the C ABI against the oracle's restatement of
VariableBaseMSM.pippengerMSM / serialMSM / doubleMSM.  Equality is equality of group elements (BNG1.equals)."""
import random

import numpy as np
import pytest

from oracle import dizk_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from octopuszk_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _run(ctx, group, scalars, bases):
    n = len(scalars)
    sb = O.pack_scalars(scalars)
    if group is O.G1:
        out = ctx.msm_g1(sb, O.pack_g1(bases), n)
        assert len(out) == 96
    else:
        out = ctx.msm_g2(sb, O.pack_g2(bases), n)
        assert len(out) == 192
    pt = util.unpack_point(group, out)
    # returned coordinates are fully reduced (Java stores them without mod, BN254aG1.java:55-59)
    flat = pt if group is O.G1 else [c for f in pt for c in f]
    assert all(0 <= v < O.P for v in flat)
    return pt


def test_reference_kat_75G(ctx):
    # SerialVariableBaseMSMTest.java:31-77 carried to G1
    bases = [O.G1.mul(O.G1.generator, k) for k in (5, 2, 7, 3)]
    got = O.G1.to_affine(_run(ctx, O.G1, [3, 11, 2, 8], bases))
    assert got[:2] == (14670023805213312856584033961079180710026959676164645964476657106778352781859,
                       211633134735504671946091929992244044834074118928621612299531666035417451988)
    # duplicates: 4 x (3 * 5G) = 60 G (DistributedVariableBaseMSMTest.java:92-124)
    got = _run(ctx, O.G1, [3] * 4, [bases[0]] * 4)
    assert O.G1.equals(got, O.G1.mul(O.G1.generator, 60))


@pytest.mark.parametrize("gname", ["G1", "G2"])
@pytest.mark.parametrize("n", [1, 2, 3, 17, 100, 1023])
def test_small_random_vs_oracle(ctx, gname, n):
    G = O.G1 if gname == "G1" else O.G2
    if G is O.G2 and n > 100:
        n = 257
    rng = random.Random(1000 + n)
    ks, pool = util.known_dlog_points(G, min(n, 12), seed=n, random_z=True)
    bases = [pool[rng.randrange(len(pool))] for _ in range(n)]
    scalars = [rng.randrange(O.R) for _ in range(n)]
    # edge cases the reference meets in Groth16 inputs: infinity bases, zero / one / r-1 scalars,
    # a point and its negation, the same point many times
    if n >= 17:
        bases[0] = G.zero()
        bases[1] = (pool[0][0], pool[0][1], G.F.zero)            # Z == 0 with junk X, Y
        scalars[2] = 0
        scalars[3] = 1
        scalars[4] = O.R - 1
        bases[5] = pool[1]
        bases[6] = G.negate(pool[1])
        scalars[5] = scalars[6] = rng.randrange(O.R)
        bases[7] = bases[8] = bases[9] = pool[2]
        scalars[7] = scalars[8] = scalars[9] = 12345
        bases[10] = G.to_affine(pool[3])                          # Z == 1
    exp = O.pippenger_msm(G, scalars, bases)
    assert G.equals(_run(ctx, G, scalars, bases), exp)
    assert G.equals(exp, O.naive_msm(G, scalars, bases))


def test_all_zero_scalars_and_empty(ctx):
    ks, pool = util.known_dlog_points(O.G1, 4, seed=3)
    assert O.G1.is_zero(_run(ctx, O.G1, [0, 0, 0, 0], pool))
    out = ctx.msm_g1(b"", b"", 0)
    assert O.G1.is_zero(O.unpack_g1(out)[0])
    out = ctx.msm_g2(b"", b"", 0)
    assert O.G2.is_zero(O.unpack_g2(out)[0])


def test_rejects_unreduced_inputs(ctx):
    from octopuszk_b200 import OzkError
    ks, pool = util.known_dlog_points(O.G1, 2, seed=4)
    with pytest.raises(OzkError):
        ctx.msm_g1(O.le32(O.R) + O.le32(1), O.pack_g1(pool), 2)
    bad = [(O.P, pool[0][1], pool[0][2]), pool[1]]
    with pytest.raises(OzkError):
        ctx.msm_g1(O.le32(1) + O.le32(1), b"".join(O.le32(v) for p in bad for v in p), 2)


def test_paired_matches_separate(ctx):
    # VariableBaseMSM.doubleMSM: one scalar vector on G1 and G2 bases (VariableBaseMSM.java:480-606)
    rng = random.Random(77)
    n = 300
    k1, p1 = util.known_dlog_points(O.G1, 8, seed=5)
    k2, p2 = util.known_dlog_points(O.G2, 8, seed=6)
    b1 = [p1[i % 8] for i in range(n)]
    b2 = [p2[i % 8] for i in range(n)]
    scalars = [rng.randrange(O.R) for _ in range(n)]
    out = ctx.msm_g1g2(O.pack_scalars(scalars), O.pack_g1(b1), O.pack_g2(b2), n)
    assert len(out) == 288
    e1, e2 = O.double_msm(scalars, b1, b2)
    assert O.G1.equals(O.unpack_g1(out[:96])[0], e1)
    assert O.G2.equals(O.unpack_g2(out[96:])[0], e2)


@pytest.mark.parametrize("gname,log_n", [("G1", 14), ("G1", 16), ("G1", 18), ("G1", 20), ("G2", 14), ("G2", 16)])
def test_large_known_dlog(ctx, gname, log_n):
    """Sizes the oracle cannot add point by point: n DISTINCT bases P_i = k_i G made on the GPU (fixed-base path, every
    point with its own Jacobian Z; sampled entries re-checked by the oracle) and scalars uniform in [0, r) with the edge
    values forced in, so the exact answer is (sum_i s_i k_i mod r) G.  An error in any bit of a point
    index -- a truncated index, a mis-strided slice -- changes the result."""
    G = O.G1 if gname == "G1" else O.G2
    n = 1 << log_n
    d_b, ks = util.gpu_distinct_bases(ctx, G, n, seed=log_n, keep_z=True)
    raw = util.force_edge_scalars(util.rand_scalars_full_range(n, seed=log_n))
    fn = ctx.msm_g1 if G is O.G1 else ctx.msm_g2
    out = fn(raw, d_b.cpu().numpy(), n)                       # host-pointer entry point (pageable numpy buffers)
    assert G.equals(util.unpack_point(G, out), util.expected_from_dot(G, raw, ks))


def test_large_tiled_random_z_pool(ctx):
    """The previous form of the large-size check, kept for the normalisation path with adversarial Z values: bases tiled from
    64 oracle-made points with random Z."""
    n = 1 << 16
    ks, pool = util.known_dlog_points(O.G1, 64, seed=16, random_z=True)
    raw = util.rand_scalars_full_range(n, seed=16)
    bases = util.tiled_bases_bytes(O.G1, pool, n)
    out = ctx.msm_g1(raw.tobytes(), bases.tobytes(), n)
    assert O.G1.equals(O.unpack_g1(out)[0], util.expected_from_dlogs(O.G1, ks, util.column_sums(raw, 64)))


def test_profiler_distribution_identical_bases(ctx):
    """The reference profiler's input (VariableBaseMSMProfiling.java:20-31): N copies of the seed-10 generator and
    scalars Fr(random long): half of them are r - |x|, so every high window has one bucket holding N/2 points, and
    every bucket sees the same point again and again (doubling case)."""
    n = 1 << 16
    rng = O.JavaRandom(10)
    g = O.G1.random(10)
    scalars = [rng.next_long() % O.R for _ in range(n)]
    out = ctx.msm_g1(O.pack_scalars(scalars), O.pack_g1([g]) * n, n)
    exp = O.G1.mul(g, sum(scalars) % O.R)
    assert O.G1.equals(O.unpack_g1(out)[0], exp)
    stats = ctx.msm_last_stats()
    assert stats[3] > 0 and stats[4] > 0          # the dense-bucket path really ran


def test_device_resident_entry_point(ctx):
    import torch
    n = 1 << 12
    ks, pool = util.known_dlog_points(O.G1, 16, seed=9)
    raw = util.rand_scalars_bytes(n, seed=9)
    bases = util.tiled_bases_bytes(O.G1, pool, n)
    d_s = torch.from_numpy(raw.copy()).cuda()
    d_b = torch.from_numpy(np.ascontiguousarray(bases)).cuda()
    out = ctx.msm_g1_dev(d_s, d_b, n)
    exp = util.expected_from_dlogs(O.G1, ks, util.column_sums(raw, 16))
    assert O.G1.equals(O.unpack_g1(out)[0], exp)


# ---- sliced host-pointer path: all slices add into one set of buckets (msm_run_host, csrc/msm.cu) -----------------------
@pytest.mark.parametrize("slices", [2, 3, 8])
def test_host_slices_share_buckets(ctx, slices, monkeypatch):
    monkeypatch.setenv("OZK_HOST_SLICES", str(slices))
    n = 100003                                   # ragged last slice
    ks, pool = util.known_dlog_points(O.G1, 64, seed=31, random_z=True)
    raw = util.rand_scalars_bytes(n, seed=31)
    bases = util.tiled_bases_bytes(O.G1, pool, n)
    out = ctx.msm_g1(raw.tobytes(), bases.tobytes(), n)
    exp = util.expected_from_dlogs(O.G1, ks, util.column_sums(raw, 64))
    assert O.G1.equals(O.unpack_g1(out)[0], exp)


def test_host_slices_paired_and_dense_buckets(ctx, monkeypatch):
    monkeypatch.setenv("OZK_HOST_SLICES", "4")
    # paired G1 + G2 (each group keeps its own buckets across the slices)
    n = 1 << 15
    k1, p1 = util.known_dlog_points(O.G1, 64, seed=32)
    k2, p2 = util.known_dlog_points(O.G2, 64, seed=33)
    raw = util.rand_scalars_bytes(n, seed=34)
    out = ctx.msm_g1g2(raw.tobytes(), util.tiled_bases_bytes(O.G1, p1, n).tobytes(), util.tiled_bases_bytes(O.G2, p2, n).tobytes(), n)
    sums = util.column_sums(raw, 64)
    assert O.G1.equals(O.unpack_g1(out[:96])[0], util.expected_from_dlogs(O.G1, k1, sums))
    assert O.G2.equals(O.unpack_g2(out[96:])[0], util.expected_from_dlogs(O.G2, k2, sums))
    # the profiler's distribution through four slices: overflow tasks are merged into resumed buckets
    n = 1 << 16
    rng = O.JavaRandom(10)
    g = O.G1.random(10)
    scalars = [rng.next_long() % O.R for _ in range(n)]
    out = ctx.msm_g1(O.pack_scalars(scalars), O.pack_g1([g]) * n, n)
    assert O.G1.equals(O.unpack_g1(out)[0], O.G1.mul(g, sum(scalars) % O.R))


# ---- persistent bases (ProvingKey.java query vectors stay on the device) --------------------
def test_keyed_msm_matches_plain_and_oracle(ctx):
    import torch
    rng = random.Random(41)
    n = 3000
    ks, pool = util.known_dlog_points(O.G1, 16, seed=41, random_z=True)
    bases = [pool[rng.randrange(16)] for _ in range(n)]
    bases[0] = O.G1.zero()
    bases[7] = (pool[0][0], pool[0][1], 0)
    bases[9] = O.G1.to_affine(pool[3])
    scalars = [rng.randrange(O.R) for _ in range(n)]
    packed = O.pack_g1(bases)
    key = ctx.upload_bases(1, packed, n)
    assert len(key) == n
    sb = O.pack_scalars(scalars)
    plain = O.unpack_g1(ctx.msm_g1(sb, packed, n))[0]
    got = O.unpack_g1(ctx.msm_keyed(sb, key, n))[0]
    assert O.G1.equals(got, plain)
    # a sub-range, as the prover's subList calls and the Java chunk loop need (VariableBaseMSM.java:211-265)
    first, m = 1000, 777
    got = O.unpack_g1(ctx.msm_keyed(O.pack_scalars(scalars[:m]), key, m, first=first))[0]
    assert O.G1.equals(got, O.pippenger_msm(O.G1, scalars[:m], bases[first:first + m]))
    # device-resident scalars, and a key built from device-resident wire-format points
    d_s = torch.frombuffer(bytearray(sb), dtype=torch.uint8).cuda()
    d_b = torch.frombuffer(bytearray(packed), dtype=torch.uint8).cuda()
    key_d = ctx.upload_bases(1, d_b, n, device=True)
    got = O.unpack_g1(ctx.msm_keyed(d_s, key_d, n, device=True))[0]
    assert O.G1.equals(got, plain)
    assert O.G1.is_zero(O.unpack_g1(ctx.msm_keyed(b"", key, 0))[0])
    key.free()
    key_d.free()


def test_keyed_paired_and_errors(ctx):
    from octopuszk_b200 import OzkError
    rng = random.Random(43)
    n = 500
    k1, p1 = util.known_dlog_points(O.G1, 8, seed=43)
    k2, p2 = util.known_dlog_points(O.G2, 8, seed=44)
    b1 = [p1[i % 8] for i in range(n)]
    b2 = [p2[i % 8] for i in range(n)]
    scalars = [rng.randrange(O.R) for _ in range(n)]
    key1 = ctx.upload_bases(1, O.pack_g1(b1), n)
    key2 = ctx.upload_bases(2, O.pack_g2(b2), n)
    sb = O.pack_scalars(scalars)
    out = ctx.msm_g1g2_keyed(sb, key1, key2, n)
    e1, e2 = O.double_msm(scalars, b1, b2)
    assert O.G1.equals(O.unpack_g1(out[:96])[0], e1)
    assert O.G2.equals(O.unpack_g2(out[96:])[0], e2)
    assert O.G2.equals(O.unpack_g2(ctx.msm_keyed(sb, key2, n))[0], e2)
    with pytest.raises(OzkError):
        ctx.msm_keyed(sb, key1, n, first=1)                     # range exceeds the key
    with pytest.raises(OzkError):
        ctx.msm_g1g2_keyed(sb, key2, key1, n)                   # groups swapped
    bad = [(O.P, p1[0][1], p1[0][2])]
    with pytest.raises(OzkError):
        ctx.upload_bases(1, b"".join(O.le32(v) for p in bad for v in p), 1)
    key1.free()
    key2.free()


def test_sum_of_wire_points(ctx):
    """ozk_sum_g1_dev / _g2_dev: the reduce(add) of per-shard partial sums (VariableBaseMSM.java:777-783), with infinity,
    a point twice (doubling) and a point with its negation among the summands."""
    import torch
    for G, grp, pack in ((O.G1, 1, O.pack_g1), (O.G2, 2, O.pack_g2)):
        ks, pool = util.known_dlog_points(G, 5, seed=61, random_z=True)
        pts = [pool[0], G.zero(), pool[1], pool[1], pool[2], G.negate(pool[2]), pool[3], G.to_affine(pool[4])]
        d = torch.frombuffer(bytearray(pack(pts)), dtype=torch.uint8).cuda()
        got = util.unpack_point(G, ctx.sum_points_dev(grp, d, len(pts)))
        exp = G.zero()
        for p in pts:
            exp = G.add(exp, p)
        assert G.equals(got, exp)
        assert G.is_zero(util.unpack_point(G, ctx.sum_points_dev(grp, d, 0)))
        assert G.equals(util.unpack_point(G, ctx.sum_points_dev(grp, d, 1)), pool[0])


@pytest.mark.timeout(900)
def test_headline_size_2_24_known_dlog(ctx):
    """BASELINE.json configs[1] at its headline size: 2^24 (scalar, point) pairs, 2^24 DISTINCT bases k_i G generated on the
    GPU (unnormalised Jacobian, every point its own Z), scalars uniform in [0, r) with forced edge values; exact answer
    (sum_i s_i k_i mod r) G.  Device-resident entry point, then the same through a persistent key with half of the pairs at
    an offset, then the paired G1+G2 call at 2^20 on distinct G2 bases."""
    import torch
    n = 1 << 24
    d_b, ks = util.gpu_distinct_bases(ctx, O.G1, n, seed=24, keep_z=True)
    raw = util.force_edge_scalars(util.rand_scalars_full_range(n, seed=24))
    d_s = torch.from_numpy(raw).cuda()
    out = ctx.msm_g1_dev(d_s, d_b, n)
    assert O.G1.equals(O.unpack_g1(out)[0], util.expected_from_dot(O.G1, raw, ks))
    key = ctx.upload_bases(1, d_b, n, device=True)
    half = n // 2
    out = ctx.msm_keyed(d_s[half:].contiguous(), key, half, first=half, device=True)
    assert O.G1.equals(O.unpack_g1(out)[0], util.expected_from_dot(O.G1, raw[half:], ks[half:]))
    key.free()
    del d_b, key
    torch.cuda.empty_cache()
    m = 1 << 20
    d_b1, k1 = util.gpu_distinct_bases(ctx, O.G1, m, seed=25, keep_z=False)
    d_b2, k2 = util.gpu_distinct_bases(ctx, O.G2, m, seed=26, keep_z=True)
    out = ctx.msm_g1g2_dev(d_s[:m].contiguous(), d_b1, d_b2, m)
    assert O.G1.equals(O.unpack_g1(out[:96])[0], util.expected_from_dot(O.G1, raw[:m], k1))
    assert O.G2.equals(O.unpack_g2(out[96:])[0], util.expected_from_dot(O.G2, raw[:m], k2))


def test_streaming_begin_feed_end(ctx):
    """ozk_msm_begin / _feed / _end: ragged feeds (also a single pair, and empty ones) into one MSM equal the whole-array call;
    paired groups; a persistent key for one group; misuse is an error, not a crash."""
    from octopuszk_b200 import OzkError
    rng = random.Random(71)
    n = 70001
    k1, p1 = util.known_dlog_points(O.G1, 64, seed=71, random_z=True)
    k2, p2 = util.known_dlog_points(O.G2, 64, seed=72, random_z=True)
    raw = util.rand_scalars_bytes(n, seed=71)
    b1 = np.ascontiguousarray(util.tiled_bases_bytes(O.G1, p1, n))
    b2 = np.ascontiguousarray(util.tiled_bases_bytes(O.G2, p2, n))
    sums = util.column_sums(raw, 64)
    e1, e2 = util.expected_from_dlogs(O.G1, k1, sums), util.expected_from_dlogs(O.G2, k2, sums)
    cuts = [0, 1, 1, 4097, 30000, 30000, 69999, n]
    ctx.msm_begin(3, n, max_slice=40000)
    for lo, hi in zip(cuts, cuts[1:]):
        ctx.msm_feed(raw[lo:hi], b1[lo:hi], b2[lo:hi], hi - lo)
    out = ctx.msm_end(3)
    assert O.G1.equals(O.unpack_g1(out[:96])[0], e1) and O.G2.equals(O.unpack_g2(out[96:])[0], e2)
    # G1 from a persistent key at an offset, scalars streamed
    key = ctx.upload_bases(1, b1, n)
    first = 64 * 100
    m = n - first
    ctx.msm_begin(1, m, key1=key, first=first)
    ctx.msm_feed(raw[:m // 2], None, None, m // 2)
    ctx.msm_feed(raw[m // 2:m], None, None, m - m // 2)
    out = ctx.msm_end(1)
    exp = util.expected_from_dlogs(O.G1, k1, util.column_sums(raw[:m], 64))     # key[first + i] = pool[i mod 64] since 64 | first
    assert O.G1.equals(O.unpack_g1(out)[0], exp)
    key.free()
    with pytest.raises(OzkError):
        ctx.msm_feed(raw[:1], b1[:1], None, 1)               # nothing in progress
    ctx.msm_begin(1, 10)
    ctx.msm_feed(raw[:4], b1[:4], None, 4)
    with pytest.raises(OzkError):
        ctx.msm_end(1)                                       # fewer pairs than announced
    ctx.msm_begin(1, 4)
    with pytest.raises(OzkError):
        ctx.msm_feed(raw[:5], b1[:5], None, 5)               # more than announced
    out = ctx.msm_g1(raw[:100], b1[:100], 100)               # the context still works
    assert O.G1.equals(O.unpack_g1(out)[0], util.expected_from_dlogs(O.G1, k1, util.column_sums(raw[:100], 64)))


# ---- batch-affine pre-reduction of the bucket runs (csrc/msm_ba_impl.cuh), forced on through OZK_MSM_BA ---------------------
@pytest.mark.parametrize("rounds", [1, 2, 3])
def test_batch_affine_rounds_match_oracle(ctx, rounds, monkeypatch):
    """Same results with 1, 2 or 3 rounds of pairwise affine additions before the XYZZ walk: small random inputs with every
    special case (infinity bases, P with -P, repeated points = doublings inside a pair, zero / one / r-1 scalars), ragged and tiny
    batch sizes, the profiler's all-identical-bases input (every pair is a doubling, dense buckets), distinct bases at 2^18,
    slices sharing buckets, and the paired call (G2 keeps the classic path on the padded index layout)."""
    monkeypatch.setenv("OZK_MSM_BA", str(rounds))
    rng = random.Random(900 + rounds)
    for n, m in ((1, None), (2, None), (17, "8"), (100, None), (1023, "16"), (5000, "512")):
        if m:
            monkeypatch.setenv("OZK_MSM_BA_M", m)
        else:
            monkeypatch.delenv("OZK_MSM_BA_M", raising=False)
        ks, pool = util.known_dlog_points(O.G1, min(n, 12), seed=n, random_z=True)
        bases = [pool[rng.randrange(len(pool))] for _ in range(n)]
        scalars = [rng.randrange(O.R) for _ in range(n)]
        if n >= 17:
            bases[0] = O.G1.zero()
            bases[1] = (pool[0][0], pool[0][1], 0)
            scalars[2], scalars[3], scalars[4] = 0, 1, O.R - 1
            bases[5], bases[6] = pool[1], O.G1.negate(pool[1])
            scalars[5] = scalars[6] = rng.randrange(O.R)
            bases[7] = bases[8] = bases[9] = pool[2]
            scalars[7] = scalars[8] = scalars[9] = 12345
        assert O.G1.equals(_run(ctx, O.G1, scalars, bases), O.pippenger_msm(O.G1, scalars, bases)), (rounds, n)
    monkeypatch.delenv("OZK_MSM_BA_M", raising=False)
    # profiler distribution: N copies of one base
    n = 1 << 16
    jr = O.JavaRandom(10)
    g = O.G1.random(10)
    scalars = [jr.next_long() % O.R for _ in range(n)]
    out = ctx.msm_g1(O.pack_scalars(scalars), O.pack_g1([g]) * n, n)
    assert O.G1.equals(O.unpack_g1(out)[0], O.G1.mul(g, sum(scalars) % O.R))
    # distinct bases, full-range scalars, host path in slices
    monkeypatch.setenv("OZK_HOST_SLICES", "3")
    n = 1 << 18
    d_b, ks = util.gpu_distinct_bases(ctx, O.G1, n, seed=77, keep_z=True)
    raw = util.force_edge_scalars(util.rand_scalars_full_range(n, seed=77))
    out = ctx.msm_g1(raw, d_b.cpu().numpy(), n)
    assert O.G1.equals(O.unpack_g1(out)[0], util.expected_from_dot(O.G1, raw, ks))
    # paired call
    m = 3000
    k1, p1 = util.known_dlog_points(O.G1, 64, seed=5)
    k2, p2 = util.known_dlog_points(O.G2, 64, seed=6)
    rw = util.rand_scalars_bytes(m, seed=7)
    out = ctx.msm_g1g2(rw.tobytes(), util.tiled_bases_bytes(O.G1, p1, m).tobytes(), util.tiled_bases_bytes(O.G2, p2, m).tobytes(), m)
    sums = util.column_sums(rw, 64)
    assert O.G1.equals(O.unpack_g1(out[:96])[0], util.expected_from_dlogs(O.G1, k1, sums))
    assert O.G2.equals(O.unpack_g2(out[96:])[0], util.expected_from_dlogs(O.G2, k2, sums))


# ---- shared-memory-resident accumulate kernel (csrc/msm_impl.cuh, msm_accumulate_sm), forced on through OZK_MSM_SMEM --------
@pytest.mark.parametrize("gname,mask", [("G1", "3"), ("G2", "3"), ("G2", "0")])
def test_smem_accumulate_matches_oracle(ctx, gname, mask, monkeypatch):
    """Both accumulate kernels of both groups, whatever the default is (G1: registers, G2: shared memory; mask "3" forces the
    shared-memory kernel, "0" the register-resident one).
    The accumulate variant that keeps the bucket accumulator and the prefetched point in shared memory gives the same sums:
    small inputs with every special case of the mixed addition (infinity bases, P with -P, the same point repeatedly = the
    doubling branch, zero / one / r-1 scalars), the profiler's identical-bases input (dense buckets, overflow tasks), distinct
    bases at 2^16, and host slices that resume the shared buckets."""
    import torch
    monkeypatch.setenv("OZK_MSM_SMEM", mask)
    G = O.G1 if gname == "G1" else O.G2
    rng = random.Random(77)
    for n in (1, 2, 17, 100, 1023 if G is O.G1 else 257):
        ks, pool = util.known_dlog_points(G, min(n, 12), seed=n, random_z=True)
        bases = [pool[rng.randrange(len(pool))] for _ in range(n)]
        scalars = [rng.randrange(O.R) for _ in range(n)]
        if n >= 17:
            bases[0] = G.zero()
            bases[1] = (pool[0][0], pool[0][1], G.F.zero)
            scalars[2], scalars[3], scalars[4] = 0, 1, O.R - 1
            bases[5], bases[6] = pool[1], G.negate(pool[1])
            scalars[5] = scalars[6] = rng.randrange(O.R)
            bases[7] = bases[8] = bases[9] = pool[2]
            scalars[7] = scalars[8] = scalars[9] = 12345
        assert G.equals(_run(ctx, G, scalars, bases), O.pippenger_msm(G, scalars, bases)), (gname, n)
    # N copies of one base, scalars Fr(random long): one bucket per high window holds N/2 points, every addition is a doubling
    n = 1 << 14
    jr = O.JavaRandom(10)
    g = G.random(10)
    scalars = [jr.next_long() % O.R for _ in range(n)]
    assert G.equals(_run(ctx, G, scalars, [g] * n), G.mul(g, sum(scalars) % O.R))
    # distinct bases k_i G, scalars uniform in [0, r), device-resident and through host slices (resume of the shared buckets)
    log_n = 16 if G is O.G1 else 14
    n = 1 << log_n
    d_b, ksv = util.gpu_distinct_bases(ctx, G, n, seed=5, keep_z=True, verify=2)
    raw = util.force_edge_scalars(util.rand_scalars_full_range(n, seed=6))
    exp = util.expected_from_dot(G, raw, ksv)
    fn = ctx.msm_g1_dev if G is O.G1 else ctx.msm_g2_dev
    assert G.equals(util.unpack_point(G, fn(torch.from_numpy(raw).cuda(), d_b, n)), exp)
    monkeypatch.setenv("OZK_HOST_SLICES", "3")
    hb = d_b.cpu().numpy()
    out = ctx.msm_g1(raw, hb, n) if G is O.G1 else ctx.msm_g2(raw, hb, n)
    assert G.equals(util.unpack_point(G, out), exp)


def test_host_schedule_replan_on_slow_link(ctx, monkeypatch):
    """The whole-array host entry point measures the rate of its first slice's upload and, on a slow host link, re-plans the
    remaining slices (small last slice).  Forced here by a threshold no link reaches: same result as the known answer, at a size
    with eight slices, with a Z per base; the measured rate is reported in the stats."""
    import torch
    n = 1 << 22
    d_b, ks = util.gpu_distinct_bases(ctx, O.G1, n, seed=31, keep_z=True, verify=2)
    raw = util.force_edge_scalars(util.rand_scalars_full_range(n, seed=32))
    exp = util.expected_from_dot(O.G1, raw, ks)
    h_s = torch.from_numpy(raw).pin_memory()
    h_b = d_b.cpu().pin_memory()
    del d_b
    out_plain = ctx.msm_g1(h_s, h_b, n)
    assert O.G1.equals(O.unpack_g1(out_plain)[0], exp)
    assert ctx.msm_last_stats()[10] > 1.0                      # GB/s of the first slice: measured
    monkeypatch.setenv("OZK_HOST_ADAPT_GBPS", "100000")
    out_replanned = ctx.msm_g1(h_s, h_b, n)
    assert O.G1.equals(O.unpack_g1(out_replanned)[0], exp)
    monkeypatch.setenv("OZK_HOST_NO_ADAPT", "1")
    assert O.G1.equals(O.unpack_g1(ctx.msm_g1(h_s, h_b, n))[0], exp)
