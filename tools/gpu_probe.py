"""Quick on-GPU probe: integer-pipe microbenchmarks and NTT timings (device-resident, CUDA events)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from octopuszk_b200 import Context  # noqa: E402
from oracle import dizk_oracle as O  # noqa: E402

ctx = Context(0, stream=torch.cuda.current_stream().cuda_stream)
res = {"gpu": torch.cuda.get_device_name(0)}
res["imad_gops"] = ctx.imad_peak()
res["modmul_gops"] = ctx.modmul_peak()
res["modmul_as_imad_gops"] = res["modmul_gops"] * 136
print(json.dumps(res), flush=True)

for log_n in [int(a) for a in sys.argv[1:]] or [16, 20, 22, 24, 26]:
    n = 1 << log_n
    d = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda")
    d[:, 31] &= 0x1F
    out = torch.empty_like(d)
    omega = O.le32(O.root_of_unity(n))
    for _ in range(3):
        ctx.ntt_dev(d, out, n, omega)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.ntt_dev(d, out, n, omega)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(json.dumps({"ntt_log_n": log_n, "ms": ms, "alg_GBps": n * 128 / ms / 1e6,
                      "modmul_per_s_alg": n * log_n / 2 / ms / 1e6}), flush=True)
