"""Groth16 prover, arithmetic hot path only (BASELINE.json metric "Groth16 prove s"): the sequence of liboctozk calls
SerialProver.prove makes for a circuit with 2^log_m constraints -- R1CStoQAPWitness's 3 inverse + 3 coset transforms,
pointwise A*B-C, divide-by-Z + coset inverse transform (R1CStoQAP.java:165-227), then the MSMs on queries A, B (G1 and G2),
H and deltaABC (SerialProver.java:76-101) -- on synthetic device-resident inputs of the right shapes: random witness values,
proving-key points produced by the fixed-base path (Z = 1, as this library's setup emits them).  The host-side field loops
of the Java (linear-combination evaluation) are outside the hot path and not included.

    python tools/prove_bench.py [log_m ...]            # one GPU
    torchrun --nproc-per-node N tools/prove_bench.py   # everything sharded over N GPUs: the seven transforms as four-step
                                                       # transforms with the exchange fused into the kernels (peer stores,
                                                       # distributed.witness_map_distributed), the MSMs over point shards
    OZK_PROVE_REPLICATED_NTT=1 keeps the earlier form (every rank runs the whole transforms) for comparison."""
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
R = O.R


def rand_fr(n, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    t = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    t[:, 31] &= 0x1F
    return t


for log_m in [int(a) for a in sys.argv[1:]] or [20]:
    m = 1 << log_m                       # constraints ~ variables
    n = 2 * m                            # FFT domain: lowestPowerOfTwo(numConstraints + numInputs)
    shard = (m + world - 1) // world     # MSM points per rank
    g1, g2 = O.G1.random(10), O.G2.random(10)
    # proving key shards (fixed-base outputs, Z = 1)
    k = rand_fr(shard, 1 + rank)
    qa = torch.empty((shard, 96), dtype=torch.uint8, device=dev)
    qb2 = torch.empty((shard, 192), dtype=torch.uint8, device=dev)
    ctx.fixed_g1_dev(O.pack_g1([g1]), k, shard, 15, 17, qa)
    ctx.fixed_g2_dev(O.pack_g2([g2]), k, shard, 15, 17, qb2)
    hshard = (n + world - 1) // world
    kh = rand_fr(hshard, 100 + rank)
    qh = torch.empty((hshard, 96), dtype=torch.uint8, device=dev)
    ctx.fixed_g1_dev(O.pack_g1([g1]), kh, hshard, 15, 17, qh)
    del k, kh
    w = rand_fr(shard, 7 + rank)         # this rank's slice of the witness
    sharded = world > 1 and not os.environ.get("OZK_PROVE_REPLICATED_NTT")
    nloc = n // world if sharded else n  # elements of A, B, C this rank holds (its cyclic shard, or everything)
    A0, B0, C0 = rand_fr(nloc, 11 + 100 * rank), rand_fr(nloc, 12 + 100 * rank), rand_fr(nloc, 13 + 100 * rank)
    A, B, C = A0.clone(), B0.clone(), C0.clone()
    ex = D.PeerExchange(ctx, nloc * 32, stream_ordered=True) if sharded else None
    omega = O.root_of_unity(n)
    wf, wi = O.le32(omega), O.le32(pow(omega, -1, R))
    ninv, g = O.le32(pow(n, -1, R)), O.FR_MULT_GEN
    zscale = O.le32(pow(n, -1, R) * pow((pow(g, n, R) - 1) % R, -1, R) % R)

    def witness_map():
        if sharded:
            # the evaluations arrive from the host each proof; here: fresh copies of the synthetic shard
            A.copy_(A0); B.copy_(B0); C.copy_(C0)
            return D.witness_map_distributed(ops, A.view(-1), B.view(-1), C.view(-1), n, exchange=ex)
        for d in (A, B, C):
            ctx.ntt_ex_dev(d, d, n, wi, None, ninv, None)
            ctx.ntt_ex_dev(d, d, n, wf, O.le32(g), None, None)
        ctx.fr_mul_sub_dev(A, B, C, A, n)
        ctx.ntt_ex_dev(A, A, n, wi, None, zscale, O.le32(pow(g, -1, R)))
        return A

    hbuf = [None]

    def msms():
        if sharded:
            h = hbuf[0].view(nloc, 32)             # this rank's coefficients of H (blocked layout; queryH is stored to match)
        else:
            h = A.view(n, 32)[rank * hshard:(rank + 1) * hshard]
        out = [D.msm_distributed(ops, w, qa, shard),                       # query A
               D.msm_distributed(ops, w, qa, shard),                       # query B, G1 half
               D.msm_distributed(ops, w, qb2, shard, g2=True),             # query B, G2 half
               D.msm_distributed(ops, h.contiguous(), qh, h.shape[0]),     # query H
               D.msm_distributed(ops, w, qa, shard)]                       # deltaABC
        return out

    def step():
        hbuf[0] = witness_map()
        return msms()

    step()
    torch.cuda.synchronize()
    res = {}
    for name, fn in (("witness_map_7_ntt", witness_map), ("msms_A_B1_B2_H_delta", msms), ("prove_hot_path", step)):
        ts = []
        for _ in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        res[name + "_ms"] = sorted(ts)[1]
    if rank == 0:
        print(json.dumps({"op": "groth16_prove_hot_path", "log_constraints": log_m, "fft_domain_log": log_m + 1, "n_gpus": world,
                          "transforms": "sharded, fused exchange" if sharded else ("replicated" if world > 1 else "single GPU"), **res}), flush=True)
    if ex is not None:
        ex.close()
    del A, B, C, A0, B0, C0, qa, qb2, qh, w
if world > 1:
    dist.destroy_process_group()
