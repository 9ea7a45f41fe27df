"""Shared helpers for the GPU parity tests: fast (numpy) generation of large MSM instances whose exact answer is
known from discrete logs, so that sizes far beyond what the Python oracle can add up are still checked exactly
("scale-parity technique")."""
import random

import numpy as np

from oracle import dizk_oracle as O


def rand_scalars_bytes(n, seed, bits=253):
    """n uniformly random scalars below 2^bits (< r for bits <= 253) as an (n, 32) uint8 array."""
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
