"""Sweep of BASELINE.json configs[1..3] on one GPU, device-resident, CUDA-event timed (median of 5 after 2 warm-ups):
G1 / G2 variable-base MSM, NTT, fixed-base batch MSM.  Prints one JSON line per point; used for profiles/rNN_sweep.jsonl.

    python tools/sweep.py [msm|cfg1|msm2|skew|ntt|fixed|fixed2 ...]   (OZK_SWEEP_LOGS=24,26 restricts the G1 MSM sizes)"""
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
IMAD_PEAK = ctx.imad_peak()          # measured integer-pipe peak (G multiply-adds / s) for the roofline fractions
what = [a for a in sys.argv[1:] if not a.startswith("-")] or ["msm", "msm2", "ntt", "fixed", "fixed2"]


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def msm_sweep(G, sizes, name):
    """sizes: pair counts.  n DISTINCT bases k_i G made on the GPU, scalars uniform in [0, r): exact answer (sum s_i k_i) G."""
    for n in sizes:
        d_b, ks = util.gpu_distinct_bases(ctx, G, n, seed=n & 0xFFFF, keep_z=False, verify=3)
        raw = util.force_edge_scalars(util.rand_scalars_full_range(n, seed=n & 0xFFFF)) if n <= (1 << 22) else None
        if raw is None:
            d_s = util.gpu_rand_scalars(n, n & 0xFFFF, torch.device("cuda"))
            raw = d_s.cpu().numpy()
        else:
            d_s = torch.from_numpy(raw).cuda()
        fn = ctx.msm_g1_dev if G is O.G1 else ctx.msm_g2_dev
        out = fn(d_s, d_b, n)
        ok = G.equals(util.unpack_point(G, out), util.expected_from_dot(G, raw, ks))
        ms = timeit(lambda: fn(d_s, d_b, n))
        st = ctx.msm_last_stats()
        imad = n * 21760 * (1 if G is O.G1 else 3) / (ms * 1e-3) / 1e9
        print(json.dumps({"op": name, "n": n, "log_n": round(np.log2(n), 3), "ok": ok, "ms": ms, "Mpairs_per_s": n / ms / 1e3, "window_bits": st[0],
                          "imad_frac": imad / IMAD_PEAK,
                          "phases_ms": {"sort": st[5], "convert": st[6], "accumulate": st[7], "merge": st[8], "reduce_final": st[9]}}), flush=True)
        del d_s, d_b
        torch.cuda.empty_cache()


def profiler_scalars(n, seed):
    """The reference profiler's scalar distribution (VariableBaseMSMProfiling.java:27-31: Fr(random long)): a uniform
    63-bit magnitude x, taken as x or as r - x with equal probability, as an (n, 32) uint8 array."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
    negative = rng.integers(0, 2, size=n, dtype=np.uint8).astype(bool)
    r = [(O.R >> (64 * k)) & ((1 << 64) - 1) for k in range(4)]
    limbs = np.zeros((n, 4), dtype=np.uint64)
    limbs[:, 0] = x
    lo = (np.uint64(r[0]) - x)                                   # mod 2^64
    borrow = (x > np.uint64(r[0])).astype(np.uint64)
    limbs[negative, 0] = lo[negative]
    limbs[negative, 1] = (np.uint64(r[1]) - borrow)[negative]    # r[1] != 0: the borrow stops here
    limbs[negative, 2] = np.uint64(r[2])
    limbs[negative, 3] = np.uint64(r[3])
    return limbs.view(np.uint8).reshape(n, 32)


if "skew" in what:
    # robustness: N copies of one base, the profiler's scalars -- every high window has one bucket
    # with N/2 points; must be correct and at least half as fast as the uniform case
    g = O.G1.random(10)
    for log_n in [20, 22, 24]:
        n = 1 << log_n
        raw = np.ascontiguousarray(profiler_scalars(n, seed=log_n))
        d_s = torch.from_numpy(raw).cuda()
        d_b = torch.from_numpy(np.frombuffer(O.pack_g1([O.G1.to_affine(g)]), dtype=np.uint8).copy()).cuda().repeat(n, 1).contiguous()
        out = ctx.msm_g1_dev(d_s, d_b, n)
        total = util.column_sums(raw, 1)[0] % O.R
        ok = O.G1.equals(O.unpack_g1(out)[0], O.G1.mul(g, total))
        ms = timeit(lambda: ctx.msm_g1_dev(d_s, d_b, n))
        st = ctx.msm_last_stats()
        print(json.dumps({"op": "varmsm_g1_profiler_distribution", "log_n": log_n, "ok": ok, "ms": ms, "Mpairs_per_s": n / ms / 1e3,
                          "window_bits": st[0], "overflow_tasks": st[3],
                          "phases_ms": {"sort": st[5], "convert": st[6], "accumulate": st[7], "merge": st[8], "reduce_final": st[9]}}), flush=True)
        del d_s, d_b

if "msm" in what:
    msm_sweep(O.G1, [1 << int(x) for x in os.environ.get("OZK_SWEEP_LOGS", "16,18,20,22,24,26").split(",")], "varmsm_g1")
if "cfg1" in what:
    # the MSM sizes of BASELINE.json configs[0] (2^15 constraints, 1023 inputs: SURVEY.md section 8 a1)
    msm_sweep(O.G1, [1023, 31748, 65537], "varmsm_g1_cfg1")
if "msm2" in what:
    msm_sweep(O.G2, [1 << int(x) for x in os.environ.get("OZK_SWEEP_LOGS2", "16,18,20,22,24").split(",")], "varmsm_g2")
if "ntt" in what or "ntt26" in what:
    ntt_logs = [int(x) for x in os.environ.get("OZK_SWEEP_NTT_LOGS", "16,18,20,22,24,26,28").split(",")]
    for log_n in ([26] if "ntt26" in what else ntt_logs):
        n = 1 << log_n
        d = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda")
        d[:, 31] &= 0x1F
        o = torch.empty_like(d)
        omega = O.le32(O.root_of_unity(n))
        ms = timeit(lambda: ctx.ntt_dev(d, o, n, omega))
        print(json.dumps({"op": "ntt_fr", "log_n": log_n, "ms": ms, "alg_GBps": n * 128 / ms / 1e6, "Gmodmul_per_s_alg": n * log_n / 2 / ms / 1e6,
                          "imad_frac": n * 68 * log_n / (ms * 1e-3) / 1e9 / IMAD_PEAK}), flush=True)
        del d, o
for key, G, logs in (("fixed", O.G1, [20, 22, 24]), ("fixed2", O.G2, [20, 22, 24])):
    if key not in what:
        continue
    base_pt = G.random(10)
    base = O.pack_g1([base_pt]) if G is O.G1 else O.pack_g2([base_pt])
    ss = G.bit_size(base_pt)
    for log_n in logs:
        n = 1 << log_n
        w = O.get_window_size(n, G)
        outerc = (ss + w - 1) // w
        raw = util.rand_scalars_bytes(n, seed=log_n)
        d_s = torch.from_numpy(raw).cuda()
        d_o = torch.empty((n, 96 if G is O.G1 else 192), dtype=torch.uint8, device="cuda")
        fn = ctx.fixed_g1_dev if G is O.G1 else ctx.fixed_g2_dev
        ms = timeit(lambda: fn(base, d_s, n, outerc, w, d_o))
        # spot check three outputs
        outb = d_o[:3].cpu().numpy().tobytes()
        pts = O.unpack_g1(outb) if G is O.G1 else O.unpack_g2(outb)
        sc = util.scalars_from_bytes(raw[:3])
        ok = all(G.equals(p, e) for p, e in zip(pts, O.fixed_batch_msm(G, ss, w, base_pt, sc)))
        imad = n * outerc * 1360 * (1 if G is O.G1 else 3) / (ms * 1e-3) / 1e9      # outerc x 10 x 136 IMAD per scalar, x3 for G2
        # size-independent exact check: sum_i out_i == (sum_i s_i) * B
        tot = ctx.sum_points_dev(1 if G is O.G1 else 2, d_o[:4096].contiguous(), 4096)
        ok = ok and G.equals(util.unpack_point(G, tot), G.mul(base_pt, util.column_sums(raw[:4096], 1)[0] % O.R))
        print(json.dumps({"op": "fixed_" + ("g1" if G is O.G1 else "g2"), "log_n": log_n, "window": w, "outerc": outerc, "ok": ok, "ms": ms,
                          "Mscalars_per_s": n / ms / 1e3, "imad_frac": imad / IMAD_PEAK}), flush=True)
        del d_s, d_o
