#!/usr/bin/env python3
"""Generates tests/golden/reference_cuda_vectors.json: inputs and the RAW OUTPUT BYTES of the reference's own code for this
path -- its three CUDA sources compiled unmodified from /root/reference into oracle/_ref/ (oracle/Makefile.ref) and called
through their own Java_* JNI entry points (oracle/ref_driver.cc) -- on seeded inputs.  The reference holds no BN254 golden
bytes of its own, and its Java cannot run in this image; these vectors are the outputs of the reference
itself, recorded on a B200, and they travel with the repository.

Needs a GPU and oracle/_ref (built by __graft_entry__.build() where /root/reference is mounted):
    gpurun -- python tests/golden/make_reference_cuda_vectors.py gpurun_out/reference_cuda_vectors.json
Consumers: tests/test_golden_vectors.py (-m "not gpu": the oracle reproduces every vector; -m gpu: liboctozk does).
The reference's dormant FFT kernel is not recorded: it reads uninitialised memory (algebra_fft_FFTAuxiliary.cu:117-119,129-131).
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dizk_oracle as O  # noqa: E402
from oracle import ref_cuda as R  # noqa: E402
from tests import util  # noqa: E402


def main(out_path):
    vec = {"generator": "tests/golden/make_reference_cuda_vectors.py", "formats": "inputs: 32-byte LE field elements (SURVEY.md "
           "Appendix A); outputs: the reference's raw return arrays (64-byte slots, LE for variable-base, BE for fixed-base / field)",
           "cases": []}

    def add(kind, **kw):
        vec["cases"].append({"kind": kind, **{k: (v.hex() if isinstance(v, (bytes, bytearray)) else v) for k, v in kw.items()}})

    # variable-base G1: the reference's ZZ-group KAT carried to G1 ([5G,2G,7G,3G] . [3,11,2,8] = 75 G), then random cases with
    # random-Z bases, infinity, repeated points, a point and its negation, zero / one / r-1 scalars
    bases = [O.G1.mul(O.G1.generator, k) for k in (5, 2, 7, 3)]
    sb, bb = O.pack_scalars([3, 11, 2, 8]), O.pack_g1(bases)
    add("var_msm_g1", n=4, scalars=sb, bases=bb, out=R.isolated("var_msm", bb, sb, 4, 1))
    for n, seed in ((33, 1), (100, 2)):
        rng = random.Random(seed)
        ks, pool = util.known_dlog_points(O.G1, 8, seed=seed, random_z=True)
        bs = [pool[rng.randrange(8)] for _ in range(n)]
        sc = [rng.randrange(O.R) for _ in range(n)]
        bs[0] = O.G1.zero()
        bs[1], bs[2] = pool[1], O.G1.negate(pool[1])
        sc[1] = sc[2]
        bs[3] = bs[4] = pool[2]
        sc[5], sc[6], sc[7] = 0, 1, O.R - 1
        sb, bb = O.pack_scalars(sc), O.pack_g1(bs)
        add("var_msm_g1", n=n, scalars=sb, bases=bb, out=R.isolated("var_msm", bb, sb, n, 1))
    # variable-base G2 and the paired call
    rng = random.Random(3)
    n = 24
    k1, p1 = util.known_dlog_points(O.G1, 6, seed=11, random_z=True)
    k2, p2 = util.known_dlog_points(O.G2, 6, seed=12, random_z=True)
    b1 = [p1[i % 6] for i in range(n)]
    b2 = [p2[i % 6] for i in range(n)]
    sc = [rng.randrange(O.R) for _ in range(n)]
    sb = O.pack_scalars(sc)
    add("var_msm_g2", n=n, scalars=sb, bases=O.pack_g2(b2), out=R.isolated("var_msm", O.pack_g2(b2), sb, n, 2))
    add("var_double_msm", n=n, scalars=sb, bases1=O.pack_g1(b1), bases2=O.pack_g2(b2),
        out=R.isolated("var_double_msm", O.pack_g1(b1), O.pack_g2(b2), sb, n))
    # fixed-base batch (G1) with the seed-10 generator and the field batch
    rng = random.Random(4)
    n = 16
    base = O.G1.random(10)
    ss, w = 253, 5
    outerc = (ss + w - 1) // w
    sc = [0, 1, O.R - 1] + [rng.randrange(O.R) for _ in range(n - 3)]
    sb = O.pack_scalars(sc)
    add("fixed_batch_g1", n=n, scalar_size=ss, window=w, outerc=outerc, base=O.pack_g1([base]), scalars=sb,
        out=R.isolated("fixed_batch", outerc, w, outerc, 1 << w, n, ss, O.pack_g1([base]), sb, 1))
    # fixed-base batch on G2 (BNType != 1) and the paired G1 + G2 call (doubleBatchMSMNativeHelper,
    # algebra_msm_FixedBaseMSM.cu:1395-1491), seed-10 generators, scalarSize 253 / 254 as the Java passes
    base2 = O.G2.random(10)
    ss2, w2 = 254, 4
    outerc2 = (ss2 + w2 - 1) // w2
    m = 12
    sb2 = O.pack_scalars(sc[:m])
    add("fixed_batch_g2", n=m, scalar_size=ss2, window=w2, outerc=outerc2, base=O.pack_g2([base2]), scalars=sb2,
        out=R.isolated("fixed_batch", outerc2, w2, outerc2, 1 << w2, m, ss2, O.pack_g2([base2]), sb2, 2))
    add("fixed_double_batch", n=m, scalar_size1=ss, window1=w, outerc1=outerc, scalar_size2=ss2, window2=w2, outerc2=outerc2,
        base1=O.pack_g1([base]), base2=O.pack_g2([base2]), scalars=sb2,
        out=R.isolated("fixed_double_batch", outerc, w, outerc2, w2, outerc, 1 << w, outerc2, 1 << w2, m, O.pack_g1([base]),
                       O.pack_g2([base2]), sb2))
    b = rng.randrange(O.R)
    add("field_batch", n=n, scalars=sb, b=O.le32(b), out=R.isolated("field_batch", sb + O.le32(b), n))
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(vec, f, indent=0)
    print("wrote", out_path, len(vec["cases"]), "cases")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "reference_cuda_vectors.json"))
