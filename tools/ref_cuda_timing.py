"""Baseline B2 of BASELINE.md: the reference's own CUDA (oracle/_ref, compiled unmodified for sm_100a) timed on the same
B200 through its JNI entry point, next to liboctozk's host-pointer entry point on the same inputs.  Wall clock of the
whole call in both cases (the reference allocates, copies and synchronises inside the call)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from octopuszk_b200 import Context  # noqa: E402
from oracle import dizk_oracle as O  # noqa: E402
from oracle import ref_cuda as R  # noqa: E402
from tests import util  # noqa: E402

ctx = Context(0)
ks, pool = util.known_dlog_points(O.G1, 64, seed=1, random_z=True)
for log_n in [int(a) for a in sys.argv[1:]] or [14, 16, 18, 20]:
    n = 1 << log_n
    raw = util.rand_scalars_bytes(n, seed=log_n)
    bases = np.ascontiguousarray(util.tiled_bases_bytes(O.G1, pool, n))
    sb, bb = raw.tobytes(), bases.tobytes()
    exp = util.expected_from_dlogs(O.G1, ks, util.column_sums(raw, 64))
    ctx.msm_g1(sb, bb, n)
    t0 = time.perf_counter()
    ours = ctx.msm_g1(sb, bb, n)
    t_ours = time.perf_counter() - t0
    R.var_msm(bb[:96 * 64], sb[:32 * 64], 64, 1)          # context / module warm-up
    t0 = time.perf_counter()
    ref = R.var_msm(bb, sb, n, 1)
    t_ref = time.perf_counter() - t0
    ok_ours = O.G1.equals(O.unpack_g1(ours)[0], exp)
    ok_ref = O.G1.equals(O.unpack_g1(ref, stride=64)[0], exp)
    print(json.dumps({"g1_msm_log_n": log_n, "ours_ms": t_ours * 1e3, "reference_cuda_ms": t_ref * 1e3,
                      "speedup": t_ref / t_ours, "ours_ok": ok_ours, "reference_ok": ok_ref}), flush=True)
