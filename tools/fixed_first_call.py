"""Fixed-base batch MSM: first call on a fresh context (window table built) against the cached steady state."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from octopuszk_b200 import Context  # noqa: E402
from oracle import dizk_oracle as O  # noqa: E402
from tests import util  # noqa: E402

base = O.pack_g1([O.G1.random(10)])
for log_n in [int(a) for a in sys.argv[1:]] or [20, 22, 24]:
    n = 1 << log_n
    d_s = torch.from_numpy(util.rand_scalars_bytes(n, seed=log_n)).cuda()
    d_o = torch.empty((n, 96), dtype=torch.uint8, device="cuda")
    ctx = Context(0)
    ctx.fr_scale_dev(d_s, d_s.clone(), 4, O.le32(1))        # CUDA context / module warm-up without touching the fixed-base path
    ctx.sync()
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.fixed_g1_dev(base, d_s, n, 13, 20, d_o)
        ctx.sync()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"op": "fixed_g1_first_call", "log_n": log_n, "first_ms": ts[0], "cached_ms": min(ts[1:])}), flush=True)
    ctx.close()
