"""Multi-GPU NTT (BASELINE.json configs[2], "four-step all-to-all beyond one GPU"): one length-n transform sharded over
the ranks of a torchrun job with octopuszk_b200.distributed.ntt_distributed, in both forms: "nccl" (local M-point transforms,
twiddle pass, ONE all_to_all, G-point cross-shard transform) and "fused" (the last pass of the local transform applies the
twiddle and stores into the peers' receive buffers over NVLink; no collective on the data path).  Device-resident, CUDA
events, max over ranks.

    torchrun --nproc-per-node N tools/ntt_multi_bench.py [log_n ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from octopuszk_b200 import Context  # noqa: E402
from octopuszk_b200 import distributed as D  # noqa: E402
from oracle import dizk_oracle as O  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
ctx = Context(local, stream=torch.cuda.current_stream().cuda_stream)
ops = D.GpuOps(ctx)

for log_n in [int(a) for a in sys.argv[1:]] or [26, 28]:
    n = 1 << log_n
    m = n // world
    omega = O.root_of_unity(n)
    x = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device=dev)
    x[:, 31] &= 0x1F
    x = x.view(-1)
    # correctness on a small instance is covered by tests/test_gpu_distributed.py; here: a delta-response spot check
    if log_n <= 26:
        d = torch.zeros(m * 32, dtype=torch.uint8, device=dev)
        j = 12345 * world + 1          # global index j lives on rank j % world at local position j // world
        if rank == j % world:
            d[32 * (j // world)] = 1
        out = D.ntt_distributed(ops, d, n, omega)
        torch.cuda.synchronize()
        c = m // world
        # rank holds X[k1*M + rank*c + t]; check k = rank*c + 5 (k1 = 0)
        k = rank * c + 5
        got = O.from_le(out[32 * 5:32 * 6].cpu().numpy().tobytes())
        assert got == pow(omega, j * k, O.R), "distributed NTT delta response mismatch"
    ex = D.PeerExchange(ctx, m * 32, stream_ordered=True) if world > 1 else None
    if ex is not None and log_n <= 26:
        a1 = D.ntt_distributed(ops, x, n, omega, exchange=ex)
        a2 = D.ntt_distributed(ops, x.clone(), n, omega)
        torch.cuda.synchronize()
        assert torch.equal(a1, a2), "fused and NCCL forms disagree"
        del a1, a2
    for form in (["nccl", "fused"] if world > 1 else ["single"]):
        ts = []
        for it in range(6):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = D.ntt_distributed(ops, x, n, omega, exchange=ex if form == "fused" else None)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it >= 2:
                ts.append(float(t.item()))
        if rank == 0:
            ms = sorted(ts)[len(ts) // 2]
            print(json.dumps({"op": "ntt_fr_distributed", "form": form, "log_n": log_n, "n_gpus": world, "ms": ms,
                              "exchange_bytes_per_gpu": (m * 32) * (world - 1) // world}), flush=True)
    if ex is not None:
        ex.close()
    del x, out
if world > 1:
    dist.destroy_process_group()
