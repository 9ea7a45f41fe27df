"""Runs the device arithmetic headers (fp256.cuh, curve.cuh) on the CPU through the carry-flag emulation in
ptx_arith.cuh and compares every result with the Python oracle.  This is the no-GPU check that the limb-level
algorithms the kernels use are right; the -m gpu tests then check the kernels themselves."""
import os
import random
import subprocess

import pytest

from oracle import dizk_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    exe = os.path.join(ROOT, "build", "host_arith_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host_arith_check.cc")])

    def run(cmds):
        out = subprocess.run([exe], input="\n".join(cmds) + "\n", capture_output=True, text=True, check=True).stdout
        return out.split("\n")[:len(cmds)]
    return run


def h(x):
    return "%x" % x


def test_field_ops(checker):
    rng = random.Random(7)
    cmds, exp = [], []
    for name, m in (("fq", O.P), ("fr", O.R)):
        edge = [0, 1, 2, m - 1, m - 2, (1 << 256) % m, (1 << 255) % m, m >> 1]
        vals = edge + [rng.randrange(m) for _ in range(200)]
        for _ in range(500):
            a, b = rng.choice(vals), rng.choice(vals)
            cmds += [f"{name} mul {h(a)} {h(b)}", f"{name} add {h(a)} {h(b)}", f"{name} sub {h(a)} {h(b)}"]
            exp += ["%064x" % (a * b % m), "%064x" % ((a + b) % m), "%064x" % ((a - b) % m)]
            # lazy reduction: wide products, 512-bit add/sub, a single Montgomery reduction
            cmds += [f"{name} lazymul {h(a)} {h(b)}", f"{name} lazydiff {h(a)} {h(b)}", f"{name} lazysum {h(a)} {h(b)}"]
            exp += ["%064x" % (a * b % m), "%064x" % ((a * b - b * b) % m), "%064x" % (a * a % m)]
        for a in vals[:40]:
            cmds += [f"{name} neg {h(a)}", f"{name} inv {h(a)}", f"{name} sqr {h(a)}"]
            exp += ["%064x" % ((-a) % m), "%064x" % (pow(a, -1, m) if a else 0), "%064x" % (a * a % m)]
        cmds += [f"{name} canon {h(m)}", f"{name} canon {h(m - 1)}", f"{name} canon {h((1 << 256) - 1)}"]
        exp += ["0", "1", "0"]
    assert checker(cmds) == exp


def test_lazy_domain_butterfly(checker):
    """The NTT butterflies keep values in [0, 2m): sums reduced against 2m, differences offset by 2m, and a Montgomery product
    without its final subtraction (fp256.cuh, "lazy" domain).  Inputs cover the whole range including 2m - 1."""
    rng = random.Random(11)
    for name, m in (("fr", O.R), ("fq", O.P)):
        edge = [0, 1, m - 1, m, m + 1, 2 * m - 1, 2 * m - 2]
        vals = edge + [rng.randrange(2 * m) for _ in range(100)]
        ws = [0, 1, m - 1] + [rng.randrange(m) for _ in range(20)]
        cases = [(rng.choice(vals), rng.choice(vals), rng.choice(ws)) for _ in range(600)]
        cases += [(2 * m - 1, 0, m - 1), (0, 2 * m - 1, m - 1), (2 * m - 1, 2 * m - 1, m - 1)]
        out = checker([f"{name} lz {h(a)} {h(b)} {h(w)}" for a, b, w in cases])
        for (a, b, w), line in zip(cases, out):
            s, d, p, sr = (int(x, 16) for x in line.split())
            assert s < 2 * m and s % m == (a + b) % m
            assert d < 2 * m and d % m == (a - b) % m
            assert p < 2 * m and p % m == (a - b) * w % m
            assert sr == s % m or (sr < m and sr % m == s % m)
            assert sr < m or s >= 2 * m


def test_fq2_ops(checker):
    rng = random.Random(8)
    F = O.Fq2Field
    cmds, exp = [], []
    vals = [(0, 0), (1, 0), (0, 1), (O.P - 1, O.P - 1)] + [(rng.randrange(O.P), rng.randrange(O.P)) for _ in range(50)]
    f2 = lambda a: "%064x %064x" % a
    for _ in range(200):
        a, b = rng.choice(vals), rng.choice(vals)
        cmds += [f"fq2 mul {h(a[0])} {h(a[1])} {h(b[0])} {h(b[1])}", f"fq2 sqr {h(a[0])} {h(a[1])}",
                 f"fq2 sub {h(a[0])} {h(a[1])} {h(b[0])} {h(b[1])}"]
        exp += [f2(F.mul(a, b)), f2(F.sqr(a)), f2(F.sub(a, b))]
    for a in vals[1:20]:
        cmds.append(f"fq2 inv {h(a[0])} {h(a[1])}")
        exp.append(f2(F.inv(a)))
    assert checker(cmds) == exp


def _aff(G, p):
    a = G.to_affine(p)
    if G.is_zero(p):
        return (0, 0) if G is O.G1 else ((0, 0), (0, 0))
    return a[:2]


def _fmt(G, a):
    if G is O.G1:
        return "%x %x" % a
    return "%x %x %x %x" % (a[0][0], a[0][1], a[1][0], a[1][1])


def _out(G, a):
    if G is O.G1:
        return "%064x %064x" % a
    return "%064x %064x %064x %064x" % (a[0][0], a[0][1], a[1][0], a[1][1])


@pytest.mark.parametrize("gname", ["g1", "g2"])
def test_curve_ops(checker, gname):
    G = O.G1 if gname == "g1" else O.G2
    rng = random.Random(9)
    pts = [G.mul(G.generator, rng.randrange(1, O.R)) for _ in range(12)]
    inf = G.zero()
    cmds, exp = [], []
    # chains of mixed adds including: repeated point (doubling), P + (-P) (infinity), infinity inputs
    chains = [
        [pts[0], pts[1], pts[2]],
        [pts[0], pts[0]],                      # double via madd
        [pts[0], G.negate(pts[0])],            # -> infinity
        [pts[0], G.negate(pts[0]), pts[3]],    # infinity then restart
        [inf, pts[4], inf, pts[5]],
        [pts[1], pts[1], pts[1], pts[1]],
        [pts[2], pts[3], G.negate(G.add(pts[2], pts[3]))],  # sum hits infinity through a non-affine accumulator
        [pts[2], pts[3], G.add(pts[2], pts[3])],            # acc == q with acc not affine -> double
        pts,
    ]
    for ch in chains:
        cmds.append(f"{gname} chain {len(ch)} " + " ".join(_fmt(G, _aff(G, p)) for p in ch))
        acc = G.zero()
        for p in ch:
            acc = G.add(acc, p)
        exp.append(_out(G, _aff(G, acc)))
    quads = [
        (pts[0], pts[1], pts[2], pts[3]),
        (pts[0], pts[1], pts[0], pts[1]),                      # equal XYZZ operands -> double
        (pts[0], pts[1], G.negate(pts[0]), G.negate(pts[1])),  # opposite -> infinity
        (inf, inf, pts[2], pts[3]),
        (pts[2], pts[3], inf, inf),
    ]
    for q in quads:
        for op in ("add", "coopadd"):        # coopadd: the four-level schedule of the MSM tail (xyzz_add_levels), same special cases
            cmds.append(f"{gname} {op} " + " ".join(_fmt(G, _aff(G, p)) for p in q))
            exp.append(_out(G, _aff(G, G.add(G.add(q[0], q[1]), G.add(q[2], q[3])))))
    for a, b in ((pts[5], pts[6]), (pts[7], inf), (inf, inf)):
        cmds.append(f"{gname} dbl " + " ".join(_fmt(G, _aff(G, p)) for p in (a, b)))
        exp.append(_out(G, _aff(G, G.twice(G.add(a, b)))))
    # the three-level doubling schedule of the MSM tail (xyzz_dbl_levels)
    for k, (a, b) in ((1, (pts[5], pts[6])), (5, (pts[7], pts[8])), (3, (inf, inf)), (2, (pts[9], inf))):
        cmds.append(f"{gname} coopdbl {k} " + " ".join(_fmt(G, _aff(G, p)) for p in (a, b)))
        e = G.add(a, b)
        for _ in range(k):
            e = G.twice(e)
        exp.append(_out(G, _aff(G, e)))
    assert checker(cmds) == exp
