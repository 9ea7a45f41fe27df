"""End-to-end G1 MSM from pinned host memory under different slice schedules (OZK_HOST_PLAN): prints one JSON line per plan.

    python tools/e2e_plan_sweep.py [log_n] [plan ...]        plan = comma-separated fractions, or "g1.3x8" = 8 slices growing x1.3"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from octopuszk_b200 import Context  # noqa: E402
from oracle import c_oracle as C  # noqa: E402
from oracle import dizk_oracle as O  # noqa: E402
from tests import util  # noqa: E402

HAMMER = "--hammer" in sys.argv          # a second process streams pinned host memory to the same GPU: a contended host link
args = [a for a in sys.argv[1:] if a != "--hammer"]
log_n = int(args[0]) if args and args[0].isdigit() else 24
plans = [a for a in args if not a.isdigit()] or ["default", "g1.3x8", "g1.2x8", "g1.0x8", "g1.3x6", "g1.3x10", "g1.5x6"]
n = 1 << log_n
dev = torch.device("cuda")
ctx = Context(0, stream=torch.cuda.current_stream().cuda_stream)
d_k = util.gpu_rand_scalars(n, 1000, dev)
d_s = util.gpu_rand_scalars(n, 2000, dev)
d_bz = torch.empty((n, 96), dtype=torch.uint8, device=dev)
ctx.fixed_g1_dev(O.pack_g1([O.G1.generator]), d_k, n, 16, 16, d_bz, keep_z=True)
torch.cuda.synchronize()
h_k, h_sraw = d_k.cpu().numpy(), d_s.cpu().numpy()
expected = O.G1.mul(O.G1.generator, C.fr_dot(h_sraw, h_k, n, os.cpu_count() or 1))
h_s = torch.from_numpy(h_sraw).pin_memory()
h_b = d_bz.cpu().pin_memory()
del d_k, d_s, d_bz
hammer = None
if HAMMER:
    import subprocess
    code = ("import torch,time\n"
            "h=torch.empty(1<<30,dtype=torch.uint8).pin_memory(); d=torch.empty(1<<30,dtype=torch.uint8,device='cuda')\n"
            "s=torch.cuda.Stream()\n"
            "with torch.cuda.stream(s):\n"
            "    while True:\n"
            "        d.copy_(h,non_blocking=True); s.synchronize()\n")
    hammer = subprocess.Popen([sys.executable, "-c", code])
    time.sleep(8)
for plan in plans:
    for k in ("OZK_HOST_PLAN", "OZK_HOST_SLICES", "OZK_HOST_SLICE_GROWTH", "OZK_HOST_NO_ADAPT", "OZK_HOST_ADAPT_GBPS"):
        os.environ.pop(k, None)
    if plan == "noadapt":
        os.environ["OZK_HOST_NO_ADAPT"] = "1"
    elif plan == "forced_replan":
        os.environ["OZK_HOST_ADAPT_GBPS"] = "100000"
    if plan.startswith("g"):
        g, k = plan[1:].split("x")
        os.environ["OZK_HOST_SLICES"] = k
        os.environ["OZK_HOST_SLICE_GROWTH"] = g
    elif plan not in ("default", "noadapt", "forced_replan"):
        os.environ["OZK_HOST_PLAN"] = plan
    out = None
    for _ in range(2):
        out = ctx.msm_g1(h_s, h_b, n)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = ctx.msm_g1(h_s, h_b, n)
        ts.append((time.perf_counter() - t0) * 1e3)
    ok = O.G1.equals(O.unpack_g1(out)[0], expected)
    st = ctx.msm_last_stats()
    print(json.dumps({"op": "e2e_g1_pinned", "log_n": log_n, "plan": plan, "contended_link": HAMMER, "ok": ok, "ms_median": sorted(ts)[2],
                      "ms_min": min(ts), "first_slice_GBps": st[10] if len(st) > 10 else None}), flush=True)
if hammer:
    hammer.kill()
