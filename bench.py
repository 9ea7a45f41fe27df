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
--impl reference times that same CPU restatement as the reference arm.
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


def _make_inputs(n, seed, random_z=True):
    """Synthetic workload: uniform scalars below 2^253 (< r) and bases tiled from 64 points with
    known discrete logs, so the exact answer is known at any n.  Bases are supplied affine (Z = 1) for the device-resident
    number and as random-Z Jacobian triples (what the reference's own fixed-base outputs look like on the wire) for the
    JNI-facing end-to-end number, as section 8d prescribes."""
    import numpy as np
    from oracle import dizk_oracle as O
    from tests import util
    ks, pool = util.known_dlog_points(O.G1, 64, seed=seed, random_z=random_z)
    raw = util.rand_scalars_bytes(n, seed=seed)
    # forced edge entries at fixed positions: 0, 1, r - 1, and the digit boundaries of the 16/17-bit windows
    for pos, val in enumerate([0, 1, O.R - 1, 1 << 16, (1 << 16) + 1, (1 << 17) - 1, 1 << 17, (1 << 253) - 1]):
        if pos < n:
            raw[pos] = np.frombuffer(O.le32(val), dtype=np.uint8)
    bases = np.ascontiguousarray(util.tiled_bases_bytes(O.G1, pool, n))
    expected = util.expected_from_dlogs(O.G1, ks, util.column_sums(raw, 64))
    return raw, bases, expected


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (C restatement, all host threads) on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle as C
    threads = os.cpu_count() or C.max_threads()          # explicit: torchrun exports OMP_NUM_THREADS=1
    log_s = args.ref_log_n
    n = 1 << log_s
    raw, bases, expected = _make_inputs(n, seed=7)
    sb, bb = raw.tobytes(), bases.tobytes()
    for _ in range(args.warmup if args.warmup < 2 else 1):
        C.msm_g1(sb, bb, n, threads)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = C.msm_g1(sb, bb, n, threads)
        times.append(time.perf_counter() - t0)
    from oracle import dizk_oracle as O
    assert O.G1.equals(O.unpack_g1(out)[0], expected)
    sec = sum(times) / len(times)
    v = n / sec
    sample = f"2^{log_s} pairs per step of the 2^{args.log_n} workload (same generator), pippengerMSM per thread + reduce(add)"
    print(json.dumps({
        "impl": "reference", "metric": "VarMSM points/sec (BN254 G1)", "value": v, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32x8 (256-bit integers mod p)", "data": "synthetic",
        "config": {"workload": f"BN254 G1 variable-base MSM, 2^{args.log_n} pairs per GPU (BASELINE.json configs[1])",
                   "sample": sample},
        "cpu_baseline": {"value": v, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from octopuszk_b200 import Context
    from oracle import dizk_oracle as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = Context(local_rank, stream=torch.cuda.current_stream().cuda_stream)

    n = 1 << args.log_n
    raw, bases, expected = _make_inputs(n, seed=100 + rank, random_z=True)     # wire form for the e2e path
    h_s = torch.from_numpy(raw).pin_memory()
    h_b = torch.from_numpy(bases).pin_memory()
    d_s = h_s.to(dev)
    d_bz = h_b.to(dev)
    from tests import util as _util
    _ks, _pool = _util.known_dlog_points(O.G1, 64, seed=100 + rank, random_z=True)
    _pool = [O.G1.to_affine(p) for p in _pool]                                          # the same points with Z = 1
    d_b = torch.from_numpy(np.ascontiguousarray(_util.tiled_bases_bytes(O.G1, _pool, n))).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        out = ctx.msm_g1_dev(d_s, d_b, n)
        if world > 1:
            part = torch.frombuffer(bytearray(out), dtype=torch.uint8).to(dev)
            gathered = torch.empty((world, 96), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(gathered, part)
            out = ctx.sum_points_dev(1, gathered, world)
        return out

    def step_e2e():
        return ctx.msm_g1(h_s, h_b, n)

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

    # integer-pipe peak, measured live on this GPU (rank 0's value is reported)
    imad_peak = ctx.imad_peak()
    modmul_peak = ctx.modmul_peak()

    for _ in range(args.warmup):
        out = step_device()
    # parity of the measured path: local result (N == 1) or the global sum (N > 1, same generator per rank seed)
    if world == 1:
        assert O.G1.equals(O.unpack_g1(out)[0], expected), "bench: MSM result differs from the known answer"
    stop, samples = threading.Event(), []
    th = threading.Thread(target=_clock_sampler, args=(stop, samples, local_rank), daemon=True)
    if rank == 0:
        th.start()
    ms_total, out, launches = timed(step_device, args.steps)
    # dominant-kernel time from the context's own events (same stream), averaged over a few more steps
    acc_ms, phase = [], None
    for _ in range(min(args.steps, 3)):
        ctx.msm_g1_dev(d_s, d_b, n)
        st = ctx.msm_last_stats()
        acc_ms.append(st[7])
        phase = st
    stop.set()
    if world > 1:
        # every rank must hold the same global sum
        chk = torch.frombuffer(bytearray(out), dtype=torch.uint8).to(dev).to(torch.int32)
        mx, mn = chk.clone(), chk.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        assert torch.equal(mx, mn), "bench: ranks disagree on the global MSM result"
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)
    # the same device-resident call on random-Z Jacobian bases (adds the batched normalisation)
    for _ in range(2):
        ctx.msm_g1_dev(d_s, d_bz, n)
    ms_rz, out_rz, _ = timed(lambda: ctx.msm_g1_dev(d_s, d_bz, n), max(1, min(args.steps, 3)))
    ms_rz /= max(1, min(args.steps, 3))
    if world == 1:
        assert O.G1.equals(O.unpack_g1(out_rz)[0], expected), "bench: random-Z MSM result differs from the known answer"
    del d_bz

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    e2e_steps = max(1, min(args.steps, 5))
    ms_e2e, out_e2e, _ = timed(step_e2e, e2e_steps)
    ms_e2e /= e2e_steps
    if world == 1:
        assert O.G1.equals(O.unpack_g1(out_e2e)[0], expected), "bench: e2e MSM result differs from the known answer"
    e2e_value = world * n / (ms_e2e * 1e-3)

    # the same call with the bases uploaded once as a persistent proving-key vector (ozk_bases_upload_g1, untimed, like
    # the reference's proving key they are the same for every proof): only the scalars cross PCIe in the timed region
    key = ctx.upload_bases(1, h_b, n)
    for _ in range(2):
        ctx.msm_keyed(h_s, key, n)
    ms_key, out_key, _ = timed(lambda: ctx.msm_keyed(h_s, key, n), e2e_steps)
    ms_key /= e2e_steps
    if world == 1:
        assert O.G1.equals(O.unpack_g1(out_key)[0], expected), "bench: keyed MSM result differs from the known answer"
    key.free()
    probes = None
    if rank == 0:
        probes = {"dfma_G_per_s": ctx.pipe_probe(1), "dfma_with_imad_wide_interleaved_G_per_s": ctx.pipe_probe(2),
                  "iadd3_G_per_s": ctx.pipe_probe(3), "imad32_G_per_s": ctx.pipe_probe(4)}

    # second headline kernel: NTT 2^26 (device-resident), reported against HBM and the integer pipe
    ntt = None
    if rank == 0 and not args.no_ntt:
        ln = args.ntt_log_n
        nn = 1 << ln
        d = torch.randint(0, 256, (nn, 32), dtype=torch.uint8, device=dev)
        d[:, 31] &= 0x1F
        o = torch.empty_like(d)
        omega = O.le32(O.root_of_unity(nn))
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
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = peaks.get("hbm_gbs", 6650.0)
        ntt = {"log_n": ln, "ms": nms,
               "hbm": {"bound": "hbm", "achieved": nn * 128 / (nms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                       "frac": nn * 128 / (nms * 1e-3) / 1e9 / hbm, "peak_source": "measured" if peaks else "fallback"},
               "imad": {"bound": "imad", "achieved": nn * 68 * ln / (nms * 1e-3) / 1e9, "peak": imad_peak, "unit": "GIMAD/s",
                        "frac": nn * 68 * ln / (nms * 1e-3) / 1e9 / imad_peak},
               "note": "integer-pipe bound: 13 modmul/element at the measured modmul peak is the floor, HBM is <15% busy"}
        del d, o

    cpu = None
    if rank == 0 and not args.no_cpu:
        from oracle import c_oracle as C
        threads = os.cpu_count() or C.max_threads()      # explicit: torchrun exports OMP_NUM_THREADS=1
        ls = args.ref_log_n
        m = 1 << ls
        sb, bb = raw[:m].tobytes(), bases[:m].tobytes()
        t0 = time.perf_counter()
        C.msm_g1(sb, bb, m, threads)
        sec = time.perf_counter() - t0
        m1 = min(m, 1 << 17)
        t1 = time.perf_counter()
        C.msm_g1(raw[:m1].tobytes(), bases[:m1].tobytes(), m1, 1)
        sec1 = time.perf_counter() - t1
        cpu = {"value": m / sec, "unit": "points/s", "cores": threads, "kind": "port",
               "single_thread": {"value": m1 / sec1, "unit": "points/s", "sample": f"first 2^{m1.bit_length() - 1} pairs, one thread"},
               "sample": f"first 2^{ls} pairs of the workload, C restatement of VariableBaseMSM.pippengerMSM per thread + reduce(add); "
                         "the reference's Java cannot run here (no JVM)"}

    if rank == 0:
        th.join(timeout=2)
        acc = sum(acc_ms) / len(acc_ms)
        achieved = n * IMAD_PER_PAIR / (acc * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_msm_accumulate_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": "VarMSM points/sec (BN254 G1)", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (256-bit integers mod p, Montgomery form)", "data": "synthetic",
            "config": {"workload": f"BN254 G1 variable-base MSM, 2^{args.log_n} pairs per GPU (BASELINE.json configs[1]); "
                                   "uniform scalars < 2^253; bases affine (Z = 1) for value, random-Z Jacobian for e2e (SURVEY.md 8d)",
                       "l2": "inputs (2 GiB) and the sorted index (1 GiB) exceed the 126 MB L2; no flush needed",
                       "partition": f"{world} x 2^{args.log_n} pairs, all_gather of 96-byte partial sums"},
            "clocks": _clock_summary(samples),
            "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": n * 128,
                    "d2h_bytes_per_step": 96},
            "e2e_resident_key": {"value": world * n / (ms_key * 1e-3), "unit": "points/s", "ms_per_step": ms_key,
                                 "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 96,
                                 "note": "bases uploaded once with ozk_bases_upload_g1 (persistent proving-key vector); scalars from pinned host memory per step"},
            "gpu_launches": launches,
            "pipe_probes": probes,
            "value_random_z_bases": {"value": world * n / (ms_rz * 1e-3), "unit": "points/s", "ms_per_step": ms_rz},
            "roofline": {"bound": "imad", "kernel": "msm_accumulate", "achieved": achieved, "peak": imad_peak, "unit": "GIMAD/s",
                         "frac": achieved / imad_peak, "traffic": traffic, "kernel_ms": acc,
                         "peak_source": "ozk_imad_peak, measured live (MEASURED_PEAKS.json has no integer-pipe figure)",
                         "modmul_peak_G_per_s": modmul_peak,
                         "algorithmic_work": f"{IMAD_PER_PAIR} IMAD per pair = {MODMUL_PER_PAIR} modmul x 136"},
            "phases_ms": {"sort": phase[5], "convert": phase[6], "accumulate": phase[7], "merge": phase[8], "reduce_final": phase[9]},
            "msm_shape": {"window_bits": phase[0], "windows": phase[1], "buckets_per_window": phase[2]},
            "ntt": ntt,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--ntt-log-n", type=int, default=26)
    ap.add_argument("--ref-log-n", type=int, default=20, help="sample size of the CPU arm (bounded: ~10-30 s of CPU work)")
    ap.add_argument("--no-ntt", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
