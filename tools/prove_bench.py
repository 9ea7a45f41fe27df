"""Groth16 prove on the reference's synthetic circuit (BASELINE.json configs[4]: 2^24 constraints, sharded over the GPUs of one
node; configs[0] is the same flow at 2^15), through octopuszk_b200/prover.py: setup (proving key resident on the device, sharded)
and prover (sparse rows x assignment, seven transforms, pointwise product, four keyed MSMs, final assembly), with the proof
checked against the expected discrete logarithms computed by the C oracle (oracle/groth16_oracle.expected_proof_synthetic: the
Groth16 verification equation in the exponent, the toxic waste being the reference's fixed seed).

    python tools/prove_bench.py [log2(constraints) ...]                      # one GPU
    python -m torch.distributed.run --nproc-per-node N tools/prove_bench.py 20 24   # sharded over N GPUs

bench.py calls `bench_leg` and puts the result into its JSON line as "groth16"."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NUM_INPUTS = 1023            # SerialzkSNARKTest / the profiler's default (BASELINE.json configs[0])


def cpu_prover_sample(log_nc: int, threads: int):
    """The CPU prover's arithmetic (C restatement of the reference's serial algorithms) on a BOUNDED sample circuit of 2^log_nc
    constraints: seven serial radix-2 transforms of the domain (three at a time across threads, as many as are independent) and
    the four G1 MSMs (pippengerMSM per thread + reduce(add)).  The G2 MSM of queryB is NOT included (the C oracle has no Fq2
    port), so this is a lower bound of the CPU time; the reference's Java cannot run here."""
    import numpy as np

    from oracle import c_oracle as C
    from oracle import dizk_oracle as O
    from tests import util
    nc = 1 << log_nc
    n = 2 * nc
    omega = O.le32(O.root_of_unity(n))
    data = util.rand_scalars_full_range(3 * n, 5)
    ks = util.rand_scalars_full_range(n, 6)
    bases = np.frombuffer(C.fixed_g1(O.pack_g1([O.G1.generator]), ks, n, 254, 16, threads), dtype=np.uint8).reshape(n, 96)
    sc = util.rand_scalars_full_range(n, 7)
    t0 = time.perf_counter()
    C.fft_fr_batch_inplace(data, n, 3, omega, threads)        # inverse transforms of A, B, C
    C.fft_fr_batch_inplace(data, n, 3, omega, threads)        # coset transforms
    C.fft_fr_batch_inplace(data, n, 1, omega, threads)        # coset inverse of H
    t_fft = time.perf_counter() - t0
    t0 = time.perf_counter()
    for cnt in (nc, nc, nc, n):                               # queryA, queryB (G1 half), deltaABC, queryH
        C.msm_g1(sc[:cnt], bases[:cnt], cnt, threads)
    t_msm = time.perf_counter() - t0
    return {"constraints_log2": log_nc, "seconds": t_fft + t_msm, "fft_s": t_fft, "msm_g1_s": t_msm, "cores": threads, "kind": "port",
            "note": "C restatement of the reference's serial FFT and pippengerMSM on a bounded sample circuit; the G2 MSM is not ported "
                    "(lower bound); cost grows ~linearly with the constraint count; the reference's Java cannot run here (no JVM)"}


def bench_leg(ctx, dev, world, rank, log_nc, host_threads, steps=3, cpu_sample_log=18, check=True):
    import numpy as np
    import torch
    import torch.distributed as dist

    from octopuszk_b200 import distributed as D
    from octopuszk_b200.groth16 import fr_random
    from octopuszk_b200.prover import DeviceGroth16, domain_size, synthetic_r1cs
    from oracle import c_oracle as C
    from oracle import dizk_oracle as O
    from oracle import groth16_oracle as GO

    nc, ni = 1 << log_nc, NUM_INPUTS
    n = domain_size(nc, ni)
    if world > 1 and n % (world * world):
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # the assignment of R1CSConstruction.serialConstruct (input generation; every rank holds it, as DistributedProver's
    # broadcast of the full assignment would leave it)
    h_z = torch.from_numpy(C.r1cs_chain(nc, fr_random(), fr_random())).pin_memory()
    d_z = h_z.to(dev)
    ex = D.PeerExchange(ctx, (n // world) * 32, stream_ordered=True) if world > 1 else None
    gz = DeviceGroth16(ctx, exchange=ex)
    barrier()
    t0 = time.perf_counter()
    full = synthetic_r1cs(nc, ni, dev)
    pk, _ = gz.setup(full, keep_vk=False)
    del full
    barrier()
    setup_s = time.perf_counter() - t0
    local = synthetic_r1cs(nc, ni, dev, world, rank)
    torch.cuda.empty_cache()

    def run(from_host):
        if from_host:
            d_z.copy_(h_z, non_blocking=True)
        return gz.prove(pk, local, d_z)

    proof = run(False)                                        # warm-up (plans, tables, scratch)
    out = {}
    for name, from_host in (("prove", False), ("prove_e2e", True)):
        run(from_host)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches()
        e0.record()
        for _ in range(steps):
            proof = run(from_host)
        e1.record()
        barrier()
        out[name + "_s"] = max_over_ranks(e0.elapsed_time(e1)) / steps / 1e3
        out[name + "_launches"] = (ctx.launches() - l0) // steps
    # phases of one more proof (witness map vs the MSMs), device events on this rank
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    gz.witness_map(local, d_z)
    ev[1].record()
    torch.cuda.synchronize()
    out["witness_map_ms"] = max_over_ranks(ev[0].elapsed_time(ev[1]))
    checked = None
    if check:
        A, B, Cp = proof
        mine = O.G1.to_affine(A), O.G2.to_affine(B), O.G1.to_affine(Cp)
        if world > 1:                                         # every rank must hold the same proof
            flat = b"".join(O.le32(v) for v in (mine[0][0], mine[0][1], mine[1][0][0], mine[1][0][1], mine[1][1][0], mine[1][1][1], mine[2][0], mine[2][1]))
            t = torch.frombuffer(bytearray(flat), dtype=torch.uint8).to(dev).to(torch.int32)
            mx, mn = t.clone(), t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            assert torch.equal(mx, mn), "groth16: ranks disagree on the proof"
        if rank == 0:
            exp = GO.expected_proof_synthetic(nc, ni, os.cpu_count() or host_threads)
            assert mine == exp, "groth16: the proof differs from the expected points (verification equation in the exponent)"
            checked = ("proof (A, B, C) == (a g1, b g2, c g1) with a, b, c from the QAP evaluated by the C oracle at the setup point; "
                       "a b = alpha beta + x gamma + c delta asserted; all ranks hold the identical affine proof")
    pk.free()
    if ex is not None:
        ex.close()
    del gz, local, d_z
    torch.cuda.empty_cache()
    res = {"constraints_log2": log_nc, "num_inputs": ni, "domain_log2": n.bit_length() - 1, "gpus": world, "prove_s": out["prove_s"],
           "prove_e2e_s": out["prove_e2e_s"], "e2e_h2d_bytes": int(h_z.numel()), "setup_s": setup_s, "witness_map_ms": out["witness_map_ms"],
           "gpu_launches_per_proof": out["prove_launches"], "checked": checked,
           "what": "SerialProver/DistributedProver.prove on R1CSConstruction.serialConstruct: rows x assignment, 7 NTTs, A*B-C, MSMs on queryA, "
                   "queryB (G1+G2), deltaABC, queryH against a device-resident sharded proving key, final assembly; assignment resident (prove_s) "
                   "or copied from pinned host memory inside the timed region (prove_e2e_s)"}
    if rank == 0 and cpu_sample_log:
        res["cpu_prover"] = cpu_prover_sample(cpu_sample_log, os.cpu_count() or host_threads)
    return res


def main():
    import torch
    import torch.distributed as dist

    from octopuszk_b200 import Context
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = Context(local, stream=torch.cuda.current_stream().cuda_stream)
    threads = max(1, (os.cpu_count() or 1) // world)
    logs = [int(a) for a in sys.argv[1:] if a.isdigit()] or [20]
    for k, log_nc in enumerate(logs):
        r = bench_leg(ctx, dev, world, rank, log_nc, threads, cpu_sample_log=18 if (k == 0 and "--cpu" in sys.argv) else 0)
        if rank == 0:
            print(json.dumps(r), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
