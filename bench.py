#!/usr/bin/env python3
"""Headline benchmark: BN254 G1 variable-base MSM, 2^24 (scalar, point) pairs per GPU (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-n 24]

One "step" = one pass of the hot path (ozk_msm_g1*) over one batch of synthetic input.
  value  : pairs/s with inputs already resident in HBM (device-pointer entry point), max over ranks, whole job.
  e2e    : pairs/s through the host-buffer entry point the JNI shim calls (pinned host buffers; H2D of scalars and
           bases and D2H of the result inside the timed region).
  roofline: the dominant kernel (msm_accumulate) against the integer pipe.  MSM is bounded by 32x32+64-bit multiply-adds
           (SURVEY.md section 8d), not by HBM or tensor cores, and MEASURED_PEAKS.json carries no integer peak, so the
           denominator is measured live by ozk_imad_peak (independent IMAD.WIDE chains on all SMs).  The HBM view of the
           second headline kernel (NTT 2^26) is reported in "ntt".
  cpu_baseline: the C restatement of the reference's CPU algorithm (oracle/dizk_oracle.c, "port": the Java itself
           cannot run, no JVM in the image) on all host cores, on a bounded sample.
  msm_strong: the same MSM with a FIXED total (2^24 and 2^26 pairs) split over the N GPUs (strong scaling).
  fixed_multi: fixed-base batch MSM (G1 and G2, BASELINE.json configs[3]) over a fixed total of 2^24 scalars split over the N GPUs.
  ntt_multi (N > 1): the four-step transform over all GPUs at 2^26 and 2^28, exchange fused into the kernels (peer stores over
           NVLink) and as an NCCL all_to_all, checked on the box against the single-GPU transform (2^22) and against Horner
           evaluation of the gathered input (full size).
  groth16: Groth16 prove seconds on the reference's synthetic circuit (BASELINE.json configs[4]) sharded over the N GPUs,
           proof checked against the expected discrete logarithms (the verification equation in the exponent).
--impl reference times the C restatement as the reference arm on the SAME 2^24 workload.

Inputs: scalars uniform in [0, r) by rejection, with the edge values forced in at fixed positions;
2^24 DISTINCT bases P_i = k_i G per GPU made by the fixed-base kernels from uniform k_i, so the exact answer at any size
and on any number of GPUs is (sum over all ranks of sum_i s_i k_i mod r) G -- asserted for every measured path.
Multi-GPU (torchrun): every rank owns 2^log_n pairs (weak scaling, no data-path collective); the only exchange is an
all_gather of the 96-byte partial sums, added on every rank by one tiny launch (ozk_sum_g1_dev).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMAD_PER_PAIR = 21760          # 16 windows x (8M+2S) x 136 multiply-adds
MODMUL_PER_PAIR = 160
IMAD_PER_NTT_ELEMENT_PER_LOG = 68


def _clock_sampler(stop, samples, gpu_index):
    """SM clock and throttle reasons while the timed region runs: NVML every 10 ms when pynvml is importable (a timed
    region is a few hundred ms), else nvidia-smi every 200 ms.  Rows: [sm_mhz, sm_max_mhz, _, hw_slowdown, hw_thermal,
    sw_thermal, sw_power_cap] with "Active"/"Not Active" strings as nvidia-smi prints them."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = [getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4))]
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not stop.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = get_reasons(h)
            samples.append([str(sm), str(mx), ""] + ["Active" if r & b else "Not Active" for b in bits])
            stop.wait(0.01)
        return
    except Exception:
        pass
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            if out:
                samples.append([x.strip() for x in out.split(",")])
        except Exception:
            pass
        stop.wait(0.2)


def _clock_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    mx = max((int(s[1]) for s in samples if s[1].isdigit()), default=None)
    reasons = set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for s in samples:
        for k, nm in enumerate(names):
            if len(s) > 3 + k and s[3 + k].lower().startswith("active"):
                reasons.add(nm)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(samples)}


WORKLOAD = ("BN254 G1 variable-base MSM, 2^{ln} pairs per GPU (BASELINE.json configs[1]); scalars uniform in [0, r) with forced "
            "edge values, 2^{ln} distinct bases k_i G (SURVEY.md 8d)")


def _edge_values():
    from oracle import dizk_oracle as O
    vals = [0, 1, O.R - 1, O.R - 2, 1 << 253, (1 << 253) - 1]
    for c in (16, 17, 20):
        vals += [1 << c, (1 << c) + 1, (1 << c) - 1]
    return vals


def _cpu_inputs(n, seed, threads):
    """The workload for the CPU arms, generated without the GPU library: the same distributions as the GPU arm (uniform scalars in
    [0, r) with the forced edge values; n distinct bases k_i G from the C oracle's fixed-base walk).  Returns scalar and base
    byte arrays and the expected point (sum_i s_i k_i) G."""
    import numpy as np
    from oracle import c_oracle as C
    from oracle import dizk_oracle as O
    from tests import util
    raw = util.rand_scalars_full_range(n, seed)
    for pos, v in enumerate(_edge_values()):
        if pos < n:
            raw[pos] = np.frombuffer(O.le32(v), dtype=np.uint8)
    ks = util.rand_scalars_full_range(n, seed ^ 0x5EED)
    bases = np.frombuffer(C.fixed_g1(O.pack_g1([O.G1.generator]), ks, n, 254, 16, threads), dtype=np.uint8).reshape(n, 96)
    expected = O.G1.mul(O.G1.generator, C.fr_dot(raw, ks, n, threads))
    return raw, bases, expected


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (C restatement, all host threads) on the bench workload itself.  One step is
    a whole 2^log_n MSM (the same configuration as the GPU arm); on a box too small for that within the time budget the sample
    shrinks and the line says so.  Steps beyond the time budget are not run (a CPU step takes seconds)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle as C
    from oracle import dizk_oracle as O
    threads = os.cpu_count() or C.max_threads()          # explicit: torchrun exports OMP_NUM_THREADS=1
    # probe the speed on 2^17 pairs, then take the largest power of two <= 2^log_n that fits ~40 s per step
    raw, bases, expected = _cpu_inputs(1 << 17, 7, threads)
    t0 = time.perf_counter()
    out = C.msm_g1(raw, bases, 1 << 17, threads)
    probe = time.perf_counter() - t0
    assert O.G1.equals(O.unpack_g1(out)[0], expected)
    log_s = args.ref_log_n if args.ref_log_n else args.log_n
    while log_s > 17 and probe * (1 << (log_s - 17)) * 0.8 > 40.0:      # larger inputs run ~20 % faster per pair (wider windows)
        log_s -= 1
    n = 1 << log_s
    raw, bases, expected = _cpu_inputs(n, 100, threads)
    budget = 150.0
    times = []
    t_start = time.perf_counter()
    for _ in range(max(1, args.steps)):
        t0 = time.perf_counter()
        out = C.msm_g1(raw, bases, n, threads)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start + times[-1] > budget:
            break
    assert O.G1.equals(O.unpack_g1(out)[0], expected), "reference arm: result differs from the known answer"
    sec = sum(times) / len(times)
    v = n / sec
    full = log_s == args.log_n
    sample = (f"{len(times)} step(s) of {'the whole' if full else 'a 2^%d-pair sample of the' % log_s} 2^{args.log_n}-pair workload; "
              "C restatement of VariableBaseMSM.pippengerMSM per thread + reduce(add); the reference's Java cannot run here (no JVM)")
    print(json.dumps({
        "impl": "reference", "metric": "VarMSM points/sec (BN254 G1)", "value": v, "unit": "points/s", "n_gpus": args.gpus,
        "steps": len(times), "steps_requested": args.steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64x4 (256-bit integers mod p)", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(ln=args.log_n), "sample": sample},
        "cpu_baseline": {"value": v, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from octopuszk_b200 import Context
    from octopuszk_b200 import distributed as D
    from oracle import c_oracle as C
    from oracle import dizk_oracle as O
    from tests import util

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = Context(local_rank, stream=torch.cuda.current_stream().cuda_stream)
    host_threads = max(1, (os.cpu_count() or 1) // world)         # for the oracle's answer checks (OMP_NUM_THREADS is 1 under torchrun)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out, ctx.launches() - l0

    def global_point(local_dot):
        """(sum over ranks of the local dot products) G: the exact answer of the whole job."""
        tot = local_dot
        if world > 1:
            mine = torch.frombuffer(bytearray(O.le32(local_dot)), dtype=torch.uint8).to(dev)
            allv = torch.empty(world * 32, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allv, mine)
            b = allv.cpu().numpy().tobytes()
            tot = sum(O.from_le(b[32 * r:32 * r + 32]) for r in range(world))
        return O.G1.mul(O.G1.generator, tot % O.R)

    # ---- inputs: made on the GPU (fixed-base kernels for the bases), copied back once for the answer check ----------------------
    n = 1 << args.log_n
    strong_logs = [] if args.no_strong else [t for t in (24, 26) if (1 << t) >= world]
    n_strong = {t: (1 << t) // world for t in strong_logs}
    n_gen = max([n] + list(n_strong.values()))
    d_k = util.gpu_rand_scalars(n_gen, 1000 + rank, dev)
    d_s = util.gpu_rand_scalars(n_gen, 2000 + rank, dev)
    for pos, v in enumerate(_edge_values()):
        d_s[pos] = torch.from_numpy(np.frombuffer(O.le32(v), dtype=np.uint8).copy()).to(dev)
    gen_packed = O.pack_g1([O.G1.generator])
    d_b = torch.empty((n_gen, 96), dtype=torch.uint8, device=dev)
    ctx.fixed_g1_dev(gen_packed, d_k, n_gen, 16, 16, d_b)                       # P_i = k_i G, normalised (Z = 1)
    d_bz = torch.empty((n, 96), dtype=torch.uint8, device=dev)
    ctx.fixed_g1_dev(gen_packed, d_k, n, 16, 16, d_bz, keep_z=True)             # the same points, every one with its own Z
    torch.cuda.synchronize()
    h_k, h_sraw = d_k.cpu().numpy(), d_s.cpu().numpy()
    for i in (0, n - 1, n // 3):                                              # sampled bases against the Python oracle
        ki = O.from_le(h_k[i].tobytes())
        for arr in (d_b, d_bz):
            assert O.G1.equals(O.unpack_g1(arr[i].cpu().numpy().tobytes())[0], O.G1.mul(O.G1.generator, ki)), "bench: generated base is not k_i G"
    dot_n = C.fr_dot(h_sraw[:n], h_k[:n], n, host_threads)
    expected = global_point(dot_n)
    expected_local = O.G1.mul(O.G1.generator, dot_n)
    h_s = torch.from_numpy(h_sraw[:n]).pin_memory()
    h_b = d_bz.cpu().pin_memory()
    del d_k

    def step_device(cnt=n):
        out = ctx.msm_g1_dev(d_s[:cnt], d_b[:cnt], cnt)
        if world > 1:
            part = torch.frombuffer(bytearray(out), dtype=torch.uint8).to(dev)
            gathered = torch.empty((world, 96), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(gathered, part)
            out = ctx.sum_points_dev(1, gathered, world)
        return out

    def step_e2e():
        return ctx.msm_g1(h_s, h_b, n)

    # integer-pipe peak, measured live on this GPU (rank 0's value is reported)
    imad_peak = ctx.imad_peak()
    modmul_peak = ctx.modmul_peak()

    for _ in range(args.warmup):
        out = step_device()
    # parity of the measured path on every rank: the GLOBAL sum must be the known answer
    assert O.G1.equals(O.unpack_g1(out)[0], expected), "bench: MSM result differs from the known answer"
    stop, samples = threading.Event(), []
    th = threading.Thread(target=_clock_sampler, args=(stop, samples, local_rank), daemon=True)
    if rank == 0:
        th.start()
    ms_total, out, launches = timed(step_device, args.steps)
    assert O.G1.equals(O.unpack_g1(out)[0], expected), "bench: MSM result of the timed region differs from the known answer"
    # dominant-kernel time from the context's own events (same stream), averaged over a few more steps
    acc_ms, phase = [], None
    for _ in range(min(args.steps, 3)):
        ctx.msm_g1_dev(d_s[:n], d_b[:n], n)
        st = ctx.msm_last_stats()
        acc_ms.append(st[7])
        phase = st
    stop.set()
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)

    # the same device-resident call on the unnormalised Jacobian bases (adds the batched normalisation)
    for _ in range(2):
        ctx.msm_g1_dev(d_s[:n], d_bz, n)
    rz_steps = max(1, min(args.steps, 3))
    ms_rz, out_rz, _ = timed(lambda: ctx.msm_g1_dev(d_s[:n], d_bz, n), rz_steps)
    ms_rz /= rz_steps
    assert O.G1.equals(O.unpack_g1(out_rz)[0], expected_local), "bench: random-Z MSM result differs from the known answer"
    del d_bz

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    e2e_steps = max(1, min(args.steps, 5))
    ms_e2e, out_e2e, _ = timed(step_e2e, e2e_steps)
    ms_e2e /= e2e_steps
    assert O.G1.equals(O.unpack_g1(out_e2e)[0], expected_local), "bench: e2e MSM result differs from the known answer"
    e2e_value = world * n / (ms_e2e * 1e-3)
    _st = ctx.msm_last_stats()
    e2e_first_slice_gbps = _st[10] if len(_st) > 10 else None      # host-link rate the library measured on the first slice (rank 0)

    # the same call with the bases uploaded once as a persistent proving-key vector (ozk_bases_upload_g1, untimed, like
    # the reference's proving key they are the same for every proof): only the scalars cross PCIe in the timed region
    key = ctx.upload_bases(1, h_b, n)
    for _ in range(2):
        ctx.msm_keyed(h_s, key, n)
    ms_key, out_key, _ = timed(lambda: ctx.msm_keyed(h_s, key, n), e2e_steps)
    ms_key /= e2e_steps
    assert O.G1.equals(O.unpack_g1(out_key)[0], expected_local), "bench: keyed MSM result differs from the known answer"
    key.free()
    del h_b
    probes = None
    if rank == 0:
        probes = {"dfma_G_per_s": ctx.pipe_probe(1), "dfma_with_imad_wide_interleaved_G_per_s": ctx.pipe_probe(2),
                  "iadd3_G_per_s": ctx.pipe_probe(3), "imad32_G_per_s": ctx.pipe_probe(4)}

    # ---- strong scaling: a FIXED total split over the GPUs ---------------------------------------------------------------------
    msm_strong = []
    for t in strong_logs:
        cnt = n_strong[t]
        if world == 1 and cnt == n:
            msm_strong.append({"total_log_n": t, "pairs_per_gpu": cnt, "ms_per_step": ms_step, "value": value, "unit": "points/s",
                               "imad_frac": cnt * IMAD_PER_PAIR / (ms_step * 1e-3) / 1e9 / imad_peak, "checked": True})
            continue
        exp_t = global_point(C.fr_dot(h_sraw[:cnt], h_k[:cnt], cnt, host_threads))
        for _ in range(2):
            o = step_device(cnt)
        assert O.G1.equals(O.unpack_g1(o)[0], exp_t), f"bench: strong-scaling MSM (2^{t} total) differs from the known answer"
        k = max(1, min(args.steps, 5 if t <= 24 else 3))
        ms_t, o, _ = timed(lambda: step_device(cnt), k)
        ms_t /= k
        assert O.G1.equals(O.unpack_g1(o)[0], exp_t)
        msm_strong.append({"total_log_n": t, "pairs_per_gpu": cnt, "ms_per_step": ms_t, "value": (1 << t) / (ms_t * 1e-3), "unit": "points/s",
                           "imad_frac": cnt * IMAD_PER_PAIR / (ms_t * 1e-3) / 1e9 / imad_peak, "checked": True})
    # ---- fixed-base batch, a FIXED total of 2^24 scalars split over the GPUs (FixedBaseMSM.distributedBatchMSM; outputs stay sharded) --
    fixed_multi = []
    if not args.no_strong:
        cnt = (1 << 24) // world
        for grp, pt_bytes, w, oc, gen in ((1, 96, 20, 13, gen_packed), (2, 192, 20, 13, O.pack_g2([O.G2.generator]))):
            d_o = torch.empty((cnt, pt_bytes), dtype=torch.uint8, device=dev)
            fn = ctx.fixed_g1_dev if grp == 1 else ctx.fixed_g2_dev
            for _ in range(2):
                fn(gen, d_s[:cnt], cnt, oc, w, d_o)
            ms_f, _, _ = timed(lambda: fn(gen, d_s[:cnt], cnt, oc, w, d_o), 3)
            ms_f /= 3
            # exact, size-independent: the sum of the first 4096 outputs is (sum of the first 4096 scalars) * B
            G = O.G1 if grp == 1 else O.G2
            tot = ctx.sum_points_dev(grp, d_o[:4096].contiguous(), 4096)
            want = G.mul(G.generator, util.column_sums(h_sraw[:4096], 1)[0] % O.R)
            got = O.unpack_g1(tot)[0] if grp == 1 else O.unpack_g2(tot)[0]
            assert G.equals(got, want), "bench: fixed-base batch outputs do not sum to (sum s_i) B"
            fixed_multi.append({"group": "G1" if grp == 1 else "G2", "total_log_n": 24, "scalars_per_gpu": cnt, "window": w, "outerc": oc, "ms": ms_f,
                                "value": (1 << 24) / (ms_f * 1e-3), "unit": "scalars/s",
                                "imad_frac": cnt * oc * 1360 * (1 if grp == 1 else 3) / (ms_f * 1e-3) / 1e9 / imad_peak, "checked": True})
            del d_o
        torch.cuda.empty_cache()
    del h_k, h_sraw
    if n_gen > n:
        d_s, d_b = d_s[:n].clone(), d_b[:n].clone()
        torch.cuda.empty_cache()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)

    def ntt_views(total_elems, ln, ms, gpus):
        per_gpu = total_elems / gpus
        return {"hbm": {"bound": "hbm", "achieved": per_gpu * 128 / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s per GPU",
                        "frac": per_gpu * 128 / (ms * 1e-3) / 1e9 / hbm, "peak_source": "measured" if peaks else "fallback"},
                "imad": {"bound": "imad", "achieved": per_gpu * IMAD_PER_NTT_ELEMENT_PER_LOG * ln / (ms * 1e-3) / 1e9, "peak": imad_peak,
                         "unit": "GIMAD/s per GPU", "frac": per_gpu * IMAD_PER_NTT_ELEMENT_PER_LOG * ln / (ms * 1e-3) / 1e9 / imad_peak}}

    # ---- second headline kernel: NTT 2^26 on one GPU (device-resident), reported against HBM and the integer pipe ------------------
    ntt = None
    if rank == 0 and not args.no_ntt:
        ln = args.ntt_log_n
        nn = 1 << ln
        d = util.gpu_rand_scalars(nn, 77, dev)
        o = torch.empty_like(d)
        omega_i = O.root_of_unity(nn)
        omega = O.le32(omega_i)
        for _ in range(3):
            ctx.ntt_dev(d, o, nn, omega)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            ctx.ntt_dev(d, o, nn, omega)
            a1.record()
            torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
        nms = sorted(ts)[len(ts) // 2]
        # parity of the measured transform: K outputs against Horner evaluation of the whole input (C oracle)
        ks = [0, 1, nn - 1, nn // 2 + 1, 12345 % nn, (nn // 3) | 1]
        got = [O.from_le(o[k].cpu().numpy().tobytes()) for k in ks]
        exp = C.fr_horner(d.cpu().numpy(), nn, [pow(omega_i, k, O.R) for k in ks], os.cpu_count() or 1)
        assert got == exp, "bench: NTT outputs differ from Horner evaluation of the input"
        ntt = {"log_n": ln, "ms": nms, "checked": f"{len(ks)} outputs == Horner evaluation of the input (C oracle)"}
        ntt.update(ntt_views(nn, ln, nms, 1))
        ntt["note"] = "integer-pipe bound: 13 modmul/element at the measured modmul peak is the floor, HBM is <15% busy"
        del d, o
    barrier()

    # ---- multi-GPU four-step transform ------------------------------------------------------------------------------------------
    ntt_multi = None
    if world > 1 and not args.no_ntt:
        ntt_multi = _bench_ntt_multi(args, ctx, dev, world, rank, timed, ntt_views)

    # ---- Groth16 prove (BASELINE.json configs[4]) -------------------------------------------------------------------------------
    groth16 = None
    if not args.no_groth16:
        from tools import prove_bench
        if hasattr(prove_bench, "bench_leg"):
            groth16 = prove_bench.bench_leg(ctx, dev, world, rank, args.groth16_log_n, host_threads)

    cpu = None
    if rank == 0 and not args.no_cpu:
        threads = os.cpu_count() or C.max_threads()      # explicit: torchrun exports OMP_NUM_THREADS=1
        raw1, bases1, exp1 = _cpu_inputs(1 << 17, 7, threads)
        t1 = time.perf_counter()
        o1 = C.msm_g1(raw1, bases1, 1 << 17, 1)
        sec1 = time.perf_counter() - t1
        assert O.G1.equals(O.unpack_g1(o1)[0], exp1)
        t1 = time.perf_counter()
        C.msm_g1(raw1, bases1, 1 << 17, threads)
        probe = time.perf_counter() - t1
        ls = args.log_n
        while ls > 17 and probe * (1 << (ls - 17)) * 0.8 > 25.0:
            ls -= 1
        m = 1 << ls
        rawm, basesm, expm = _cpu_inputs(m, 100, threads)
        t0 = time.perf_counter()
        om = C.msm_g1(rawm, basesm, m, threads)
        sec = time.perf_counter() - t0
        assert O.G1.equals(O.unpack_g1(om)[0], expm), "bench: CPU baseline result differs from the known answer"
        cpu = {"value": m / sec, "unit": "points/s", "cores": threads, "kind": "port",
               "single_thread": {"value": (1 << 17) / sec1, "unit": "points/s", "sample": "2^17 pairs, one thread"},
               "sample": f"{'the whole' if ls == args.log_n else '2^%d pairs of the' % ls} 2^{args.log_n}-pair workload (same generator), C restatement of "
                         "VariableBaseMSM.pippengerMSM per thread + reduce(add); the reference's Java cannot run here (no JVM)"}

    if rank == 0:
        th.join(timeout=2)
        acc = sum(acc_ms) / len(acc_ms)
        achieved = n * IMAD_PER_PAIR / (acc * 1e-3) / 1e9
        traffic, traffic_src = None, None
        for name in ("r2_msm_accumulate_traffic.json", "r1_msm_accumulate_traffic.json"):
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", name))).get("dram_bytes_per_launch")
                traffic_src = f"profiles/{name} (one ncu --set full capture of this command; not re-measured in this run)"
                break
            except Exception:
                pass
        line = {
            "metric": "VarMSM points/sec (BN254 G1)", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (256-bit integers mod p, Montgomery form)", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(ln=args.log_n) + "; bases Z = 1 for value, unnormalised Jacobian (own Z per point) for e2e",
                       "l2": "inputs (2 GiB) and the sorted index (1 GiB) exceed the 126 MB L2; no flush needed",
                       "partition": f"{world} x 2^{args.log_n} pairs, all_gather of 96-byte partial sums",
                       "answer_check": "global sum == (sum over ranks of sum_i s_i k_i mod r) G on every rank, for value, e2e, keyed and strong legs"},
            "clocks": _clock_summary(samples),
            "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": n * 128,
                    "d2h_bytes_per_step": 96, "first_slice_h2d_GBps": e2e_first_slice_gbps,
                    "slice_schedule": ("small last slice (host link below 38 GB/s)" if e2e_first_slice_gbps and e2e_first_slice_gbps < 38.0
                                       else "x1.3 growth")},
            "e2e_resident_key": {"value": world * n / (ms_key * 1e-3), "unit": "points/s", "ms_per_step": ms_key,
                                 "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 96,
                                 "note": "bases uploaded once with ozk_bases_upload_g1 (persistent proving-key vector); scalars from pinned host memory per step"},
            "gpu_launches": launches,
            "pipe_probes": probes,
            "value_random_z_bases": {"value": world * n / (ms_rz * 1e-3), "unit": "points/s", "ms_per_step": ms_rz},
            "roofline": {"bound": "imad", "kernel": "msm_accumulate", "achieved": achieved, "peak": imad_peak, "unit": "GIMAD/s",
                         "frac": achieved / imad_peak, "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": acc,
                         "peak_source": "ozk_imad_peak, measured live (MEASURED_PEAKS.json has no integer-pipe figure)",
                         "modmul_peak_G_per_s": modmul_peak,
                         "algorithmic_work": f"{IMAD_PER_PAIR} IMAD per pair = {MODMUL_PER_PAIR} modmul x 136"},
            "phases_ms": {"sort": phase[5], "convert": phase[6], "accumulate": phase[7], "merge": phase[8], "reduce_final": phase[9]},
            "msm_shape": {"window_bits": phase[0], "windows": phase[1], "buckets_per_window": phase[2]},
            "msm_strong": msm_strong,
            "fixed_multi": fixed_multi,
            "ntt": ntt,
            "ntt_multi": ntt_multi,
            "groth16": groth16,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        barrier()
        dist.destroy_process_group()


def _bench_ntt_multi(args, ctx, dev, world, rank, timed, ntt_views):
    """The four-step transform over all ranks (octopuszk_b200/distributed.py) at 2^26 and 2^28 points, exchange fused into the
    kernels and as an NCCL all_to_all; checked on the box at 2^22 against the single-GPU transform of the gathered input and at
    full size against Horner evaluation (C oracle) of the gathered input at K outputs."""
    import torch
    import torch.distributed as dist

    from octopuszk_b200 import distributed as D
    from oracle import c_oracle as C
    from oracle import dizk_oracle as O
    from tests import util
    ops = D.GpuOps(ctx)
    results = []
    max_log = max(args.ntt_multi_logs)
    ex = D.PeerExchange(ctx, ((1 << max_log) // world) * 32, stream_ordered=True)

    def gather_natural(x_local, m):
        """natural-order vector from the cyclic shards (x[d + G i2] = shard_d[i2]) on every rank"""
        allx = torch.empty((world, m, 32), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allx.view(-1), x_local.view(-1))
        return allx.permute(1, 0, 2).contiguous().view(world * m, 32)

    def run_forms(ln):
        nn = 1 << ln
        m = nn // world
        c = m // world
        omega = O.root_of_unity(nn)
        x = util.gpu_rand_scalars(m, 5000 + 16 * ln + rank, dev)
        out_f = D.ntt_distributed(ops, x.view(-1), nn, omega, exchange=ex).view(m, 32)
        out_n = D.ntt_distributed(ops, x.clone().view(-1), nn, omega).view(m, 32)
        torch.cuda.synchronize()
        assert torch.equal(out_f, out_n), "bench: fused and NCCL forms of the multi-GPU transform differ"
        return nn, m, c, omega, x, out_f

    # (a) 2^22: the whole output against the single-GPU transform of the gathered input
    nn, m, c, omega, x, out = run_forms(22)
    xn = gather_natural(x, m)
    ref = torch.empty_like(xn)
    ctx.ntt_dev(xn, ref, nn, O.le32(omega))
    torch.cuda.synchronize()
    assert torch.equal(out.view(world, c, 32), ref.view(world, world, c, 32)[:, rank]), "bench: multi-GPU transform != single-GPU transform at 2^22"
    del xn, ref
    for ln in args.ntt_multi_logs:
        nn, m, c, omega, x, out = run_forms(ln)
        # (b) K outputs against Horner evaluation of the gathered input on rank 0's host cores
        rng_ks = [0, 1, nn - 1, nn // 2 + 1, m + c + 1, 3 * (nn // 7), (nn // world) * (world - 1) + 5, 977 * 977 % nn]
        mine = torch.zeros((len(rng_ks), 32), dtype=torch.int32, device=dev)
        for q, k in enumerate(rng_ks):                 # X[k1 M + d c + t] lives on rank d at [k1][t]
            k1, rem = divmod(k, m)
            d, t = divmod(rem, c)
            if d == rank:
                mine[q] = out[k1 * c + t].to(torch.int32)
        dist.all_reduce(mine)
        xn = gather_natural(x, m)
        checked = None
        if rank == 0:
            got = [O.from_le(bytes(mine[q].to(torch.uint8).cpu().numpy().tobytes())) for q in range(len(rng_ks))]
            exp = C.fr_horner(xn.cpu().numpy(), nn, [pow(omega, k, O.R) for k in rng_ks], os.cpu_count() or 1)
            assert got == exp, f"bench: multi-GPU transform at 2^{ln} differs from Horner evaluation of the input"
            checked = f"{len(rng_ks)} outputs == Horner evaluation of the gathered input (C oracle); fused == NCCL form bit for bit"
        del xn
        for form in ("fused", "nccl"):
            def step():
                if form == "fused":
                    return D.ntt_distributed(ops, x.view(-1), nn, omega, exchange=ex)
                return D.ntt_distributed(ops, x.view(-1), nn, omega)      # overwrites x: still a valid input for timing
            for _ in range(2):
                step()
            k = 5
            ms, _, _ = timed(step, k)
            ms /= k
            r = {"log_n": ln, "form": form, "ms": ms, "gpus": world, "checked": checked}
            r.update(ntt_views(nn, ln, ms, world))
            results.append(r)
        del x, out
        torch.cuda.empty_cache()
    ex.close()
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--ntt-log-n", type=int, default=26)
    ap.add_argument("--ntt-multi-logs", type=int, nargs="*", default=[26, 28])
    ap.add_argument("--groth16-log-n", type=int, default=24)
    ap.add_argument("--ref-log-n", type=int, default=0, help="sample size of the CPU reference arm (default: the bench size, shrunk only if a step would exceed ~40 s)")
    ap.add_argument("--no-ntt", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-groth16", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
