"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, the all-gather of MSM partial sums and the
one-all-to-all four-step transform.  The local compute is the oracle here (tests only) -- what is under test is the
index arithmetic and the collectives of octopuszk_b200/distributed.py, checked against the serial oracle results the
reference's DistributedFFTTest / DistributedVariableBaseMSMTest compare with."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle as C
from oracle import dizk_oracle as O
from tests import util


class OracleOps:
    """CPU stand-in for GpuOps (same method names), backed by the oracle."""

    def empty_like(self, t):
        return torch.empty_like(t)

    def to_device(self, b):
        return torch.frombuffer(bytearray(b), dtype=torch.uint8)

    def ntt(self, x, n, omega):
        out = C.fft_fr(x.numpy().tobytes(), O.le32(omega))
        x.copy_(torch.frombuffer(bytearray(out), dtype=torch.uint8))

    def scale_powers(self, x, n, coset, scale=None, first_index=0):
        v = [O.from_le(bytes(x[32 * i:32 * i + 32].tolist())) for i in range(n)]
        s = 1 if scale is None else scale
        g = 1 if coset is None else coset
        v = [a * s * pow(g, first_index + i, O.R) % O.R for i, a in enumerate(v)]
        x.copy_(torch.frombuffer(bytearray(O.pack_scalars(v)), dtype=torch.uint8))

    def mul_sub(self, a, b, c, out, n):
        f = lambda t: [O.from_le(bytes(t[32 * i:32 * i + 32].tolist())) for i in range(n)]
        va, vb, vc = f(a), f(b), f(c)
        out.copy_(torch.frombuffer(bytearray(O.pack_scalars([(x * y - z) % O.R for x, y, z in zip(va, vb, vc)])), dtype=torch.uint8))

    def dft_small(self, x, out, groups, length, omega_g):
        raw = x.numpy().tobytes()
        v = [O.from_le(raw[32 * i:32 * i + 32]) for i in range(groups * length)]
        res = [0] * (groups * length)
        for j in range(length):
            col = [v[i1 * length + j] for i1 in range(groups)]
            for k1 in range(groups):
                res[k1 * length + j] = sum(col[i1] * pow(omega_g, i1 * k1, O.R) for i1 in range(groups)) % O.R
        out.copy_(torch.frombuffer(bytearray(O.pack_scalars(res)), dtype=torch.uint8))

    def msm_g1(self, scalars, bases, n):
        return C.msm_g1(scalars.numpy().tobytes(), bases.numpy().tobytes(), n, 1)

    def fixed(self, base, scalars, n, outerc, window, g2=False):
        assert not g2
        sc = [O.from_le(bytes(scalars.reshape(-1)[32 * i:32 * i + 32].tolist())) for i in range(n)]
        pts = [O.G1.mul(O.unpack_g1(base)[0], s % (1 << (outerc * window))) for s in sc]
        return torch.frombuffer(bytearray(O.pack_g1(pts)), dtype=torch.uint8).view(n, 96)

    def fr_scale(self, a, n, b):
        out = C.fr_scale(a.numpy().tobytes(), O.le32(b))
        return torch.frombuffer(bytearray(out), dtype=torch.uint8).view(a.shape)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from octopuszk_b200 import distributed as D
    ops = OracleOps()
    # ---- NTT: cyclic shards in, [k1][t] chunks out, one all-to-all
    n = 64
    rng = random.Random(5)
    x = [rng.randrange(O.R) for _ in range(n)]
    omega = O.root_of_unity(n)
    shard = torch.frombuffer(bytearray(D.ntt_scatter_cyclic(O.pack_scalars(x), world, rank)), dtype=torch.uint8)
    out = D.ntt_distributed(ops, shard, n, omega)
    # ---- MSM: contiguous shards, all-gather of partial sums
    ks, pool = util.known_dlog_points(O.G1, 8, seed=3)
    total = 40
    bases = [pool[i % 8] for i in range(total)]
    scalars = [rng.randrange(O.R) for _ in range(total)]
    lo, hi = rank * total // world, (rank + 1) * total // world
    res = D.msm_distributed(ops, torch.frombuffer(bytearray(O.pack_scalars(scalars[lo:hi])), dtype=torch.uint8),
                            torch.frombuffer(bytearray(O.pack_g1(bases[lo:hi])), dtype=torch.uint8), hi - lo)
    # ---- mirrored transform: blocked layout in, cyclic out (chains with the one above without a re-layout)
    back = D.ntt_distributed_blocked_in(ops, out.clone(), n, pow(omega, -1, O.R))     # inverse of the transform above, times n
    # ---- sharded R1CStoQAPWitness transform chain
    ev = [[rng.randrange(O.R) for _ in range(n)] for _ in range(3)]
    sh = [torch.frombuffer(bytearray(D.ntt_scatter_cyclic(O.pack_scalars(v), world, rank)), dtype=torch.uint8) for v in ev]
    h = D.witness_map_distributed(ops, sh[0], sh[1], sh[2], n)
    # ---- fixed-base batch and field batch over scalar slices (FixedBaseMSM.distributedBatchMSM / distributedFieldBatchMSM)
    fsc = [rng.randrange(O.R) for _ in range(12)]
    flo, fhi = rank * 12 // world, (rank + 1) * 12 // world
    fs = torch.frombuffer(bytearray(O.pack_scalars(fsc[flo:fhi])), dtype=torch.uint8).view(-1, 32)
    gbase = O.pack_g1([O.G1.random(10)])
    f_sharded = D.fixed_batch_distributed(ops, gbase, fs, fhi - flo, 23, 11)
    f_all = D.fixed_batch_distributed(ops, gbase, fs, fhi - flo, 23, 11, gather=True)
    fld = D.field_batch_distributed(ops, fs, fhi - flo, 12345, gather=True)
    assert f_sharded.shape == (fhi - flo, 96) and f_all.shape == (12, 96)
    got = O.unpack_g1(f_all.numpy().tobytes())
    assert all(O.G1.equals(a, b) for a, b in zip(got, O.fixed_batch_msm(O.G1, 253, 11, O.G1.random(10), fsc)))
    fb = fld.numpy().tobytes()
    assert [O.from_le(fb[32 * i:32 * i + 32]) for i in range(12)] == O.field_batch_msm(fsc, 12345)
    q.put((rank, out.numpy().tobytes(), res, x, scalars, bases, back.numpy().tobytes(), ev, h.numpy().tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from octopuszk_b200 import distributed as D
    x = results[0][3]
    n = len(x)
    exp = list(x)
    O.serial_radix2_fft(exp, O.root_of_unity(n))                       # DistributedFFTTest.java:41-67: distributed == serial
    got = D.ntt_gather_natural([r[1] for r in results], n)
    assert [O.from_le(got[32 * i:32 * i + 32]) for i in range(n)] == exp
    # blocked-in transform with omega^-1 undoes it (up to the factor n), and returns the cyclic layout
    for r in results:
        cyc = [O.from_le(r[6][32 * i:32 * i + 32]) for i in range(n // world)]
        assert cyc == [x[r[0] + world * i] * n % O.R for i in range(n // world)]
    # witness map: H = cosetIFFT((cosetFFT(IFFT(A)) * cosetFFT(IFFT(B)) - cosetFFT(IFFT(C))) / Z), R1CStoQAP.java:165-227
    ev = results[0][7]
    dom = O.SerialFFT(n)
    g = O.FR_MULT_GEN
    cos = []
    for v in ev:
        v = list(v)
        dom.radix2_inverse_fft(v)
        dom.radix2_coset_fft(v, g)
        cos.append(v)
    hh = [(a * b - c) % O.R for a, b, c in zip(*cos)]
    dom.divide_by_z_on_coset(g, hh)
    dom.radix2_coset_inverse_fft(hh, g)
    got_h = D.ntt_gather_natural([r[8] for r in results], n)
    assert [O.from_le(got_h[32 * i:32 * i + 32]) for i in range(n)] == hh
    scalars, bases = results[0][4], results[0][5]
    e = O.pippenger_msm(O.G1, scalars, bases)
    for r in results:                                                  # every rank holds the same global sum
        assert O.G1.equals(O.unpack_g1(r[2])[0], e)


def test_scatter_gather_layouts_single_process():
    from octopuszk_b200 import distributed as D
    n, world = 32, 4
    x = O.pack_scalars(list(range(1, n + 1)))
    shards = [D.ntt_scatter_cyclic(x, world, r) for r in range(world)]
    assert O.from_le(shards[1][:32]) == 2 and O.from_le(shards[1][32:64]) == 6
    # gather is the inverse of "rank d holds X[k1*M + d*M/G + t] at [k1][t]"
    m, c = n // world, n // world // world
    outs = []
    for d in range(world):
        outs.append(b"".join(x[(k1 * m + d * c + t) * 32:(k1 * m + d * c + t) * 32 + 32] for k1 in range(world) for t in range(c)))
    assert D.ntt_gather_natural(outs, n) == x


def test_prover_shards_partition_everything():
    """ProverShards (octopuszk_b200/prover.py): over all ranks the variable slices, the auxiliary slices and the H blocks cover every
    index exactly once, the H blocks are the blocked layout witness_map_distributed leaves, and the rows that carry the primary
    inputs are found on the right rank at the right local row."""
    from octopuszk_b200.prover import ProverShards, domain_size
    for nc, ni, world in ((64, 7, 1), (64, 7, 2), (61, 6, 4), (1 << 10, 50, 8), (1 << 15, 1023, 8)):
        nv = nc + 3
        n = domain_size(nc, ni)
        seen_v, seen_a, seen_h, seen_in = [], [], [], {}
        for rank in range(world):
            sh = ProverShards(world, rank, n, ni, nv)
            seen_v += list(range(*sh.var_range))
            seen_a += list(range(*sh.aux_range))
            for first, cnt in sh.h_blocks():
                seen_h += list(range(first, first + cnt))
            if world > 1:
                m, c = n // world, n // world // world
                assert sh.h_blocks() == [(k1 * m + rank * c, c) for k1 in range(world)]
            ir = sh.input_rows(nc)
            if ir is not None:
                row0, var0, count, step = ir
                for k in range(count):
                    g_row = rank + world * (row0 + k)          # global domain index of local row row0 + k
                    seen_in[g_row] = var0 + k * step
        assert seen_v == list(range(nv)) and seen_a == list(range(nv - ni)) and sorted(seen_h) == list(range(n))
        assert seen_in == {nc + i: i for i in range(ni)}
