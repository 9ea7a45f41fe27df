"""Shared helpers for the GPU parity tests: fast (numpy) generation of large MSM instances whose exact answer is
known from discrete logs, so that sizes far beyond what the Python oracle can add up are still checked exactly
("scale-parity technique")."""
import random

import numpy as np

from oracle import dizk_oracle as O


def rand_scalars_bytes(n, seed, bits=None):
    """n uniformly random scalars as an (n, 32) uint8 array: uniform in [0, r) by default (rejection sampling, see
    rand_scalars_full_range), or below 2^bits (< r for bits <= 253) when `bits` is given."""
    if bits is None:
        return rand_scalars_full_range(n, seed)
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    full, rem = divmod(bits, 8)
    raw[:, full + (1 if rem else 0):] = 0
    if rem:
        raw[:, full] &= (1 << rem) - 1
    return raw


def scalars_from_bytes(raw):
    return [int.from_bytes(raw[i].tobytes(), "little") for i in range(raw.shape[0])]


def column_sums(raw, mod_classes):
    """sums[j] = sum of the scalars whose index is congruent to j mod mod_classes (exact, via 32-bit limb sums)."""
    n = raw.shape[0]
    limbs = raw.view(np.uint32).reshape(n, 8).astype(np.uint64)
    out = []
    for j in range(mod_classes):
        col = limbs[j::mod_classes].sum(axis=0, dtype=np.uint64)
        out.append(sum(int(col[k]) << (32 * k) for k in range(8)))
    return out


def known_dlog_points(group, count, seed, random_z=True):
    """count points k_j * G with known k_j, as Jacobian triples with random Z (fixed-base outputs in the reference
    arrive with arbitrary Z)."""
    rng = random.Random(seed)
    ks, pts = [], []
    F = group.F
    for _ in range(count):
        k = rng.randrange(1, O.R)
        p = group.to_affine(group.mul(group.generator, k))
        if random_z:
            if group is O.G1:
                lam = rng.randrange(1, O.P)
            else:
                lam = (rng.randrange(1, O.P), rng.randrange(O.P))
            l2 = F.sqr(lam)
            p = (F.mul(p[0], l2), F.mul(p[1], F.mul(l2, lam)), lam)
        ks.append(k)
        pts.append(p)
    return ks, pts


def tiled_bases_bytes(group, pts, n):
    """(n, 96|192) uint8 array with bases[i] = pts[i % len(pts)] in the wire format."""
    packed = O.pack_g1(pts) if group is O.G1 else O.pack_g2(pts)
    stride = 96 if group is O.G1 else 192
    table = np.frombuffer(packed, dtype=np.uint8).reshape(len(pts), stride)
    idx = np.arange(n) % len(pts)
    return table[idx]


def expected_from_dlogs(group, ks, sums):
    k = sum(a * b for a, b in zip(ks, sums)) % O.R
    return group.mul(group.generator, k)


def unpack_point(group, b):
    return O.unpack_g1(b)[0] if group is O.G1 else O.unpack_g2(b)[0]


# ---- full-range scalars and distinct GPU-generated bases --------------------
_R_LIMBS = [(O.R >> (64 * k)) & ((1 << 64) - 1) for k in range(4)]


def rand_scalars_full_range(n, seed):
    """n scalars uniform in [0, r) as an (n, 32) uint8 array: 254-bit candidates from a seeded stream, rejection of
    values >= r, so the range [2^253, r) -- where the top signed digit and the last carry of the
    recoding live -- is hit at every size."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, 32), dtype=np.uint8)
    have = 0
    while have < n:
        m = int((n - have) * 1.4) + 64
        raw = rng.integers(0, 256, size=(m, 32), dtype=np.uint8)
        raw[:, 31] &= 0x3F
        limbs = raw.view("<u8").reshape(m, 4)
        lt = np.zeros(m, dtype=bool)
        eq = np.ones(m, dtype=bool)
        for k in (3, 2, 1, 0):
            lt |= eq & (limbs[:, k] < np.uint64(_R_LIMBS[k]))
            eq &= limbs[:, k] == np.uint64(_R_LIMBS[k])
        good = raw[lt]
        take = min(n - have, good.shape[0])
        out[have:have + take] = good[:take]
        have += take
    return out


def force_edge_scalars(raw, c_bits=(16, 17, 20)):
    """Overwrite the first entries with the edge values SURVEY.md section 8d prescribes: 0, 1, r - 1 and window boundaries."""
    vals = [0, 1, O.R - 1, O.R - 2, (1 << 253), (1 << 253) - 1, (O.R - 1) >> 1]
    for c in c_bits:
        vals += [1 << c, (1 << c) + 1, (1 << c) - 1, 1 << (c - 1), (1 << (c - 1)) + 1]
    for pos, v in enumerate(vals):
        if pos < raw.shape[0]:
            raw[pos] = np.frombuffer(O.le32(v % O.R), dtype=np.uint8)
    return raw


def gpu_distinct_bases(ctx, group, n, seed, keep_z=False, verify=8):
    """n DISTINCT bases P_i = k_i * G made on the GPU by the fixed-base path (ozk_fixed_g1/g2_ex_dev) from seeded uniform
    k_i in [0, r); returns (device uint8 tensor (n, 96|192), k as (n, 32) uint8 numpy).  keep_z=True leaves every point
    with its own Jacobian Z (what the reference's fixed-base outputs look like), otherwise Z = 1.  `verify` sampled
    entries (first, last and random ones) are recomputed with the Python oracle, so a later MSM check over these bases
    does not rest on the fixed-base kernels alone."""
    import torch
    dev = torch.device("cuda", ctx.device)
    if n > (1 << 22):
        d_k = gpu_rand_scalars(n, seed ^ 0x5EED, dev)
        d_k[0] = torch.from_numpy(np.frombuffer(O.le32(1), dtype=np.uint8).copy()).to(dev)
        ks = d_k.cpu().numpy()
    else:
        ks = rand_scalars_full_range(n, seed ^ 0x5EED)
        ks[0] = np.frombuffer(O.le32(1), dtype=np.uint8)          # P_0 = G itself
        d_k = torch.from_numpy(ks).to(dev)
    stride = 96 if group is O.G1 else 192
    d_b = torch.empty((n, stride), dtype=torch.uint8, device=dev)
    torch.cuda.current_stream().synchronize()
    if group is O.G1:
        ctx.fixed_g1_dev(O.pack_g1([O.G1.generator]), d_k, n, 16, 16, d_b, keep_z=keep_z)
    else:
        ctx.fixed_g2_dev(O.pack_g2([O.G2.generator]), d_k, n, 16, 16, d_b, keep_z=keep_z)
    ctx.sync()
    rng = random.Random(seed)
    for i in sorted({0, n - 1} | {rng.randrange(n) for _ in range(max(0, verify - 2))}):
        got = unpack_point(group, d_b[i].cpu().numpy().tobytes())
        k = int.from_bytes(ks[i].tobytes(), "little")
        assert group.equals(got, group.mul(group.generator, k)), f"generated base {i} is not k_i * G"
    return d_b, ks


def expected_from_dot(group, scalars_raw, ks_raw, threads=0):
    """(sum_i s_i k_i mod r) * G with the dot product from the C oracle (oracle_fr_dot)."""
    from oracle import c_oracle as C
    n = scalars_raw.shape[0]
    k = C.fr_dot(np.ascontiguousarray(scalars_raw), np.ascontiguousarray(ks_raw), n, threads)
    return group.mul(group.generator, k)


def gpu_rand_scalars(n, seed, device):
    """Uniform scalars in [0, r) generated on the GPU (torch; same rejection rule as rand_scalars_full_range, a different
    stream): (n, 32) uint8 CUDA tensor.  For inputs too large to draw on the host in reasonable time (2^26 and up)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, 32), dtype=torch.uint8, device=device)
    sign = -(1 << 63)
    rl = [((v + (1 << 63)) % (1 << 64)) - (1 << 63) for v in _R_LIMBS]            # limbs of r as int64 bit patterns
    have = 0
    while have < n:
        m = int((n - have) * 1.4) + 64
        raw = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device=device, generator=g)
        raw[:, 31] &= 0x3F
        limbs = raw.view(torch.int64).view(m, 4)
        lt = torch.zeros(m, dtype=torch.bool, device=device)
        eq = torch.ones(m, dtype=torch.bool, device=device)
        for k in (3, 2, 1, 0):
            a = limbs[:, k] ^ sign                                               # unsigned order through signed compare
            b = rl[k] ^ sign
            lt |= eq & (a < b)
            eq &= limbs[:, k] == rl[k]
        good = raw[lt]
        take = min(n - have, good.shape[0])
        out[have:have + take] = good[:take]
        have += take
        del raw, limbs, good
    return out
