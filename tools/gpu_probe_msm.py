"""On-GPU probe: device-resident G1 MSM timings (CUDA events)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from octopuszk_b200 import Context  # noqa: E402
from oracle import dizk_oracle as O  # noqa: E402
from tests import util  # noqa: E402

ctx = Context(0, stream=torch.cuda.current_stream().cuda_stream)
ks, pool = util.known_dlog_points(O.G1, 64, seed=1, random_z=("--randz" in sys.argv))
for log_n in [int(a) for a in sys.argv[1:] if a.isdigit()] or [16, 20, 22, 24]:
    n = 1 << log_n
    raw = util.rand_scalars_bytes(n, seed=log_n)
    bases = util.tiled_bases_bytes(O.G1, pool, n)
    d_s = torch.from_numpy(raw).cuda()
    d_b = torch.from_numpy(np.ascontiguousarray(bases)).cuda()
    out = ctx.msm_g1_dev(d_s, d_b, n)
    ok = O.G1.equals(O.unpack_g1(out)[0], util.expected_from_dlogs(O.G1, ks, util.column_sums(raw, 64)))
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.msm_g1_dev(d_s, d_b, n)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(json.dumps({"msm_g1_log_n": log_n, "ok": ok, "ms": ms, "Mpts_per_s": n / ms / 1e3, "stats": ctx.msm_last_stats()}), flush=True)
