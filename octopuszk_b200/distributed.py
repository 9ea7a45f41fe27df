"""Multi-GPU forms of the hot path: one process per GPU, torch.distributed for the plumbing.

Replaces the reference's Spark distribution:
  * VariableBaseMSM.distributedMSM (src/main/java/algebra/msm/VariableBaseMSM.java:772-786: mapPartitions + reduce(add))
    -> every rank runs the full single-GPU MSM on its contiguous shard; the only exchange is an all_gather of the
       96/192-byte partial sums, added on every rank (ozk_sum_g1_dev / _g2_dev).  No data-path collective.
  * FFTAuxiliary.distributedRadix2FFT (src/main/java/algebra/fft/FFTAuxiliary.java:129-219: two shuffles)
    -> four-step with n = G * M over G ranks and ONE all-to-all:
         input  : rank d holds x[d + G * i2], i2 < M                      (cyclic shard)
         step 1 : local M-point transform over i2 with omega^G             (the single-GPU kernel)
         twiddle: Y[k2] *= omega^(d * k2)
         exchange: all_to_all of M/G-element chunks (chunk c = k2 in [c M/G, (c+1) M/G) goes to rank c)
         step 2 : G-point transform across the received chunks with omega^M
         output : rank d holds X[k1 * M + d * M/G + t] at [k1][t]           (G chunks of M/G, k1-major)
       `ntt_gather_natural` turns the distributed output into the natural-order vector for checks.
       With a `PeerExchange` (receive buffers of all ranks mapped into every process over CUDA IPC) step 1, the twiddle
       and the exchange are ONE kernel sequence: the last pass of the local transform multiplies by omega^(d * k2) and
       stores every element straight into the owning rank's receive buffer over NVLink (ozk_ntt_fr_scatter_dev); the
       NCCL all_to_all and the separate twiddle pass disappear, only two tiny stream-ordered barriers remain.

The local compute is behind a small `ops` object so the index logic and the collectives can be exercised on CPU with
the gloo backend (tests/test_distributed_cpu.py plugs the oracle in there); `GpuOps` is the product path."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
from .algebra import FR_MULTIPLICATIVE_GENERATOR, FR_ROOT  # noqa: E402  (BN254aFrParameters.java:29-34)


def _le32(x: int) -> bytes:
    return x.to_bytes(32, "little")


class GpuOps:
    """Product path: liboctozk kernels on this rank's GPU; tensors are uint8 CUDA tensors of 32-byte elements."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)

    def empty_like(self, t):
        return torch.empty_like(t)

    def ntt(self, x, n, omega):
        self.ctx.ntt_dev(x, x, n, _le32(omega))

    def scale_powers(self, x, n, coset, scale=None, first_index=0):
        self.ctx.fr_scale_powers_dev(x, x, n, None if scale is None else _le32(scale), None if coset is None else _le32(coset), first_index)

    def ntt_from(self, src, out, n, omega):
        self.ctx.ntt_dev(src, out, n, _le32(omega))

    def dft_small_scatter(self, x, peer_ptrs, rank, length, omega_g, omega_n):
        self.ctx.dft_small_scatter_dev(x, peer_ptrs, rank, length, _le32(omega_g), _le32(omega_n))

    def mul_sub(self, a, b, c, out, n):
        self.ctx.fr_mul_sub_dev(a, b, c, out, n)

    def ntt_scatter(self, x, peer_ptrs, rank, n_local, omega_local, twiddle_base):
        self.ctx.ntt_scatter_dev(x, peer_ptrs, rank, n_local, _le32(omega_local), _le32(twiddle_base))

    def dft_small(self, x, out, groups, length, omega_g):
        self.ctx.fr_dft_small_dev(x, out, groups, length, _le32(omega_g))

    def msm_g1(self, scalars, bases, n):
        return self.ctx.msm_g1_dev(scalars, bases, n)

    def msm_g2(self, scalars, bases, n):
        return self.ctx.msm_g2_dev(scalars, bases, n)

    def sum_points(self, points, k, g2=False):
        return self.ctx.sum_points_dev(2 if g2 else 1, points, k)

    def to_device(self, b: bytes):
        return torch.frombuffer(bytearray(b), dtype=torch.uint8).to(self.device)

    def fixed(self, base: bytes, scalars, n, outerc, window, g2=False):
        out = torch.empty((n, 192 if g2 else 96), dtype=torch.uint8, device=self.device)
        (self.ctx.fixed_g2_dev if g2 else self.ctx.fixed_g1_dev)(base, scalars, n, outerc, window, out)
        return out

    def fr_scale(self, a, n, b: int):
        out = torch.empty_like(a)
        self.ctx.fr_scale_dev(a, out, n, _le32(b))
        return out


def _world(group):
    return dist.get_world_size(group) if dist.is_initialized() else 1


def _rank(group):
    return dist.get_rank(group) if dist.is_initialized() else 0


def msm_distributed(ops, scalars_local, bases_local, n_local: int, g2: bool = False, group=None) -> bytes:
    """sum over all ranks' shards of scalars[i] * bases[i]; every rank returns the same point (wire format)."""
    world = _world(group)
    fn = ops.msm_g2 if g2 else ops.msm_g1
    part = fn(scalars_local, bases_local, n_local)
    if world == 1:
        return part
    size = 192 if g2 else 96
    mine = ops.to_device(part)
    gathered = torch.empty((world * size,), dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    if hasattr(ops, "sum_points"):
        return ops.sum_points(gathered, world, g2)            # one tiny launch: reduce(add) of the partial sums
    ones = bytearray(32 * world)
    for r in range(world):
        ones[32 * r] = 1
    return fn(ops.to_device(bytes(ones)), gathered, world)


def _gather_equal(local, group):
    """all ranks' equally sized shards, concatenated in rank order, on every rank"""
    world = _world(group)
    if world == 1:
        return local
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out.view((world * local.shape[0],) + tuple(local.shape[1:]))


def fixed_batch_distributed(ops, base: bytes, scalars_local, n_local: int, outerc: int, window: int, g2: bool = False, gather: bool = False,
                            group=None):
    """FixedBaseMSM.distributedBatchMSM (src/main/java/algebra/msm/FixedBaseMSM.java:446-472: mapPartitions over the scalar
    partitions against the broadcast window table), used by DistributedSetup.generate for queryA / queryH / deltaABC
    (zk_proof_systems/zkSNARK/DistributedSetup.java:69-99,137-164): out[i] = scalars[i] * base for this rank's slice.  Embarrassingly
    parallel: every rank walks its own slice against its own cached window table (built once per rank; nothing is broadcast), and
    the outputs STAY SHARDED, in input order, where the sharded prover consumes them.  gather=True returns the whole vector on
    every rank instead (equal slices required)."""
    out = ops.fixed(base, scalars_local, n_local, outerc, window, g2)
    return _gather_equal(out, group) if gather else out


def fixed_double_batch_distributed(ops, base1: bytes, base2: bytes, scalars_local, n_local: int, outerc1: int, window1: int, outerc2: int,
                                   window2: int, gather: bool = False, group=None):
    """FixedBaseMSM.distributedDoubleBatchMSM (FixedBaseMSM.java:712-741; queryB of DistributedSetup, :110-135): the pairs
    (s_i * base1 in G1, s_i * base2 in G2) of this rank's slice, as two sharded vectors."""
    return (fixed_batch_distributed(ops, base1, scalars_local, n_local, outerc1, window1, False, gather, group),
            fixed_batch_distributed(ops, base2, scalars_local, n_local, outerc2, window2, True, gather, group))


def field_batch_distributed(ops, scalars_local, n_local: int, b: int, gather: bool = False, group=None):
    """FixedBaseMSM.distributedFieldBatchMSM (FixedBaseMSM.java:880-900; DistributedSetup.java:69-76,169): a_i * b mod r on this
    rank's slice."""
    out = ops.fr_scale(scalars_local, n_local, b)
    return _gather_equal(out, group) if gather else out


class PeerExchange:
    """The receive buffers of all ranks of one node, mapped into this process (cudaMalloc + CUDA IPC through
    ozk_peer_alloc / ozk_peer_open), for the fused exchange of `ntt_distributed`.  One instance serves transforms of up
    to `nbytes` bytes per rank.  The barriers below must order the KERNELS of all ranks: with stream_ordered=True the caller
    guarantees that the context runs on the torch current stream (Context(stream=torch.cuda.current_stream().cuda_stream)),
    so the NCCL all_reduce enqueued on that stream is a device-side barrier and nothing synchronises with the host; otherwise
    (default, any context) each barrier also drains the context's stream and the torch stream on the host."""

    def __init__(self, ctx, nbytes: int, group=None, stream_ordered: bool = False):
        self.ctx, self.group, self.nbytes, self.stream_ordered = ctx, group, nbytes, stream_ordered
        self.world, self.rank = _world(group), _rank(group)
        dev = torch.device("cuda", ctx.device)
        self.ptr, handle = ctx.peer_alloc(nbytes)
        mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
        allh = torch.empty(self.world * 64, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        hs = allh.cpu().numpy().tobytes()
        self.peers = [self.ptr if r == self.rank else ctx.peer_open(hs[64 * r:64 * r + 64]) for r in range(self.world)]
        self._token = torch.zeros(1, dtype=torch.int32, device=dev)

    def barrier(self):
        """Barrier over the ranks: every kernel enqueued before it, on every rank, completes before any kernel enqueued after
        it starts (a 4-byte all_reduce; see the class docstring for the two modes)."""
        if not self.stream_ordered:
            self.ctx.sync()
        dist.all_reduce(self._token, group=self.group)
        if not self.stream_ordered:
            torch.cuda.current_stream().synchronize()

    def close(self):
        self.barrier()
        torch.cuda.synchronize()
        for r, p in enumerate(self.peers):
            if r != self.rank:
                self.ctx.peer_close(p)
        self.ctx.peer_free(self.ptr)
        self.peers = []


def ntt_distributed(ops, x_local, n: int, omega: int, group=None, exchange: Optional[PeerExchange] = None):
    """Forward transform of a length-n vector sharded cyclically over the ranks (see the module docstring for layouts).
    x_local: uint8 tensor of M * 32 bytes (overwritten by the NCCL form, preserved by the fused form).  Returns a uint8
    tensor of the same size laid out [k1][t]."""
    world = _world(group)
    rank = _rank(group)
    assert n % world == 0 and world in (1, 2, 4, 8)
    m = n // world
    assert x_local.numel() == m * 32
    if world == 1:
        ops.ntt(x_local, n, omega)
        return x_local
    assert m % world == 0
    if exchange is not None:
        assert exchange.nbytes >= m * 32 and exchange.world == world
        exchange.barrier()                                                   # every rank's receive buffer is free again
        ops.ntt_scatter(x_local, exchange.peers, rank, m, pow(omega, world, FR_MODULUS), pow(omega, rank, FR_MODULUS))
        exchange.barrier()                                                   # all remote stores have landed
        out = ops.empty_like(x_local)
        ops.dft_small(exchange.ptr, out, world, m // world, pow(omega, m, FR_MODULUS))   # step 2
        return out
    ops.ntt(x_local, m, pow(omega, world, FR_MODULUS))                       # step 1
    ops.scale_powers(x_local, m, pow(omega, rank, FR_MODULUS))               # twiddle omega^(d * k2)
    recv = ops.empty_like(x_local)
    dist.all_to_all_single(recv, x_local, group=group)                       # chunk c -> rank c
    out = ops.empty_like(x_local)
    ops.dft_small(recv, out, world, m // world, pow(omega, m, FR_MODULUS))   # step 2
    return out


def ntt_distributed_blocked_in(ops, x_local, n: int, omega: int, group=None, exchange: Optional[PeerExchange] = None):
    """The mirrored transform: x_local is in the OUTPUT layout of `ntt_distributed` (rank d holds x[a * M + d * c + t] at
    [a][t], c = M / G); returns the transform in the cyclic layout (rank d holds X[d + G * k2] at k2).
         step 1 : G-point transform over a (local)           y[k1][t] = sum_a x[a][t] omega_G^(a k1)
         twiddle: y[k1][t] *= omega^((d c + t) k1)
         exchange: block y[k1] goes to rank k1, placed at d c  (fused: stored there by the step-1 kernel)
         step 2 : local M-point transform with omega^G
    Alternating the two forms chains transforms without ever re-laying data out (sharded R1CStoQAPWitness below)."""
    world = _world(group)
    rank = _rank(group)
    assert n % world == 0 and world in (1, 2, 4, 8)
    m = n // world
    assert x_local.numel() == m * 32
    if world == 1:
        ops.ntt(x_local, n, omega)
        return x_local
    assert m % world == 0
    c = m // world
    omega_g = pow(omega, m, FR_MODULUS)
    if exchange is not None:
        assert exchange.nbytes >= m * 32 and exchange.world == world
        exchange.barrier()
        ops.dft_small_scatter(x_local, exchange.peers, rank, c, omega_g, omega)
        exchange.barrier()
        out = ops.empty_like(x_local)
        ops.ntt_from(exchange.ptr, out, m, pow(omega, world, FR_MODULUS))
        return out
    y = ops.empty_like(x_local)
    ops.dft_small(x_local, y, world, c, omega_g)
    for k1 in range(1, world):
        blk = y[k1 * c * 32:(k1 + 1) * c * 32]
        ops.scale_powers(blk, c, pow(omega, k1, FR_MODULUS), None, rank * c)           # omega^(k1 (d c + t))
    recv = ops.empty_like(y)
    dist.all_to_all_single(recv, y, group=group)                                     # block k1 -> rank k1, ordered by source d
    ops.ntt(recv, m, pow(omega, world, FR_MODULUS))
    return recv


def scale_blocked(ops, x_local, n: int, world: int, rank: int, scale: Optional[int], coset: Optional[int]):
    """x[i] *= scale * coset^i for data in the blocked layout (global index i = a M + rank c + t at [a][t])."""
    m = n // world
    c = m // world
    for a in range(world):
        blk = x_local[a * c * 32:(a + 1) * c * 32]
        ops.scale_powers(blk, c, coset, scale, a * m + rank * c)


def witness_map_distributed(ops, A, B, C, n: int, group=None, exchange: Optional[PeerExchange] = None):
    """R1CStoQAP.R1CStoQAPWitness's transform chain (src/main/java/reductions/r1cs_to_qap/R1CStoQAP.java:165-227) on a
    domain of n points sharded over the ranks: A, B, C hold the evaluations a_i, b_i, c_i of this rank's cyclic shard
    (index rank + G * i2, overwritten).  Inverse transform (cyclic -> blocked), coset scaling in the blocked layout, coset
    forward transform (blocked -> cyclic), H' = A B - C pointwise, inverse coset transform (cyclic -> blocked) with the
    divide-by-Z factor folded in.  Returns H in the blocked layout: rank d holds the coefficients h[k1 M + d c + t] at [k1][t]
    -- the order in which this rank's slice of queryH must be stored for the H MSM."""
    world = _world(group)
    rank = _rank(group)
    R = FR_MODULUS
    g = FR_MULTIPLICATIVE_GENERATOR
    omega = pow(FR_ROOT, R // n, R)                                                  # Fp.rootOfUnity, Fp.java:98-102
    omega_inv = pow(omega, -1, R)
    n_inv = pow(n, -1, R)
    m = n // world
    on_coset = []
    for X in (A, B, C):
        if world == 1:
            ops.ntt(X, n, omega_inv)
            ops.scale_powers(X, n, g, n_inv, 0)
            ops.ntt(X, n, omega)
            on_coset.append(X)
            continue
        co = ntt_distributed(ops, X, n, omega_inv, group, exchange)                   # coefficients * n, blocked
        scale_blocked(ops, co, n, world, rank, n_inv, g)                              # radix2InverseFFT's 1/n and multiplyByCoset
        on_coset.append(ntt_distributed_blocked_in(ops, co, n, omega, group, exchange))   # evaluations on the coset, cyclic
    a, b, c = on_coset
    ops.mul_sub(a, b, c, a, m)                                                        # R1CStoQAP.java:211-214
    z_inv = pow((pow(g, n, R) - 1) % R, -1, R)                                        # divideByZOnCoset, SerialFFT.java:157-162
    if world == 1:
        ops.ntt(a, n, omega_inv)
        ops.scale_powers(a, n, pow(g, -1, R), n_inv * z_inv % R, 0)
        return a
    h = ntt_distributed(ops, a, n, omega_inv, group, exchange)
    scale_blocked(ops, h, n, world, rank, n_inv * z_inv % R, pow(g, -1, R))           # radix2CosetInverseFFT
    return h


def ntt_scatter_cyclic(x: bytes, world: int, rank: int) -> bytes:
    """The cyclic shard of a natural-order vector: elements rank, rank + world, ..."""
    n = len(x) // 32
    return b"".join(x[32 * i:32 * i + 32] for i in range(rank, n, world))


def ntt_gather_natural(shards, n: int) -> bytes:
    """Natural-order vector from the per-rank outputs of ntt_distributed (shards[d] = bytes of rank d's output)."""
    world = len(shards)
    m = n // world
    c = m // world
    out = bytearray(n * 32)
    for d, sh in enumerate(shards):
        for k1 in range(world):
            src = sh[(k1 * c) * 32:(k1 * c + c) * 32]
            dst = (k1 * m + d * c) * 32
            out[dst:dst + c * 32] = src
    return bytes(out)
