"""Device-resident Groth16 setup and prover, on one GPU or sharded over the GPUs of one node (BASELINE.json configs[0] and [4]).

The callers either side of the hot path, restated so that NOTHING of size O(constraints) ever exists as
Python objects or returns to the host between the assignment and the proof:

  R1CS rows              CSR on the device (`Csr`); the reference's synthetic circuit (profiler/generation/R1CSConstruction.java:
                         48-104) is generated directly in that form (`synthetic_r1cs`), any other circuit through `Csr.from_rows`
  SerialSetup.generate   zk_proof_systems/zkSNARK/SerialSetup.java:32-192      -> `setup`: Lagrange coefficients (ozk_fr_lagrange_dev),
  DistributedSetup       zk_proof_systems/zkSNARK/DistributedSetup.java:36-204    At/Bt/Ct as TRANSPOSED sparse products (ozk_fr_spmv_ex_dev),
                                                                                 the vector combinations (ozk_fr_lincomb_dev), the fixed-base
                                                                                 batches (ozk_fixed_g1/g2_dev), query vectors kept as
                                                                                 persistent device-resident bases (ozk_bases_upload_*_dev)
  R1CStoQAPWitness       reductions/r1cs_to_qap/R1CStoQAP.java:126-238         -> `witness_map`: sparse rows x assignment, 7 transforms,
                                                                                 pointwise product, all on the device
  SerialProver.prove     zk_proof_systems/zkSNARK/SerialProver.java:26-119     -> `prove`: keyed MSMs on device-resident scalars; H goes from
  DistributedProver      zk_proof_systems/zkSNARK/DistributedProver.java:28-167   the last transform straight into the H MSM

Sharding (one process per GPU, torch.distributed; replaces the Spark RDDs): rank d of G
  * evaluates the constraint rows d, d + G, ... (cyclic shard of the domain) and runs the seven transforms sharded
    (distributed.witness_map_distributed: four-step transforms, exchange fused into the kernels when a PeerExchange is given);
  * owns a contiguous slice of the variables for queryA / queryB / deltaABC and, for queryH, exactly the coefficients the
    sharded witness map leaves on it (blocked layout) -- `ProverShards`; the setup builds only those slices on each rank
    (FixedBaseMSM.distributedBatchMSM / distributedDoubleBatchMSM, FixedBaseMSM.java:446-472,712-741: slice the scalars,
    window table per rank, outputs stay where the prover consumes them);
  * the five partial sums of a proof (96 + 288 + 96 + 96 bytes) are gathered once and added on every rank.
All "random" field elements are Fp.random(seed 10) as in the reference's Configuration (configuration/Configuration.java:52)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from . import distributed as D
from .algebra import FR_MODULUS as R
from .algebra import FR_MULTIPLICATIVE_GENERATOR, FR_ROOT, FixedBaseMSM, VariableBaseMSM, _le32
from .groth16 import G1_ONE, G2_ONE, fr_random
from .lib import Bases, Context


def domain_size(num_constraints: int, num_inputs: int) -> int:
    """SerialFFT's domain for the QAP (R1CStoQAP.java:139-141, SerialFFT.java:26): the next power of two."""
    d = 1
    while d < num_constraints + num_inputs:
        d <<= 1
    return d


# ---- sparse matrices ---------------------------------------------------------------------------------------------------------
class Csr:
    """rows x cols sparse matrix over Fr in CSR on the device: row_ptr int32[rows + 1], col int32[nnz], coeff (nnz, 32) uint8
    canonical elements or None when every coefficient is 1."""

    def __init__(self, row_ptr, col, coeff, rows: int, cols: int):
        self.row_ptr, self.col, self.coeff, self.rows, self.cols = row_ptr, col, coeff, rows, cols

    @staticmethod
    def from_rows(rows, cols: int, device) -> "Csr":
        """From a list of linear combinations [(index, value), ...] (the LinearCombination objects of the Java)."""
        ptr, col, coeff = [0], [], []
        for lc in rows:
            for idx, val in lc:
                col.append(idx)
                coeff.append(val % R)
            ptr.append(len(col))
        unit = all(v == 1 for v in coeff)
        cf = None
        if not unit:
            cf = torch.frombuffer(bytearray(b"".join(_le32(v) for v in coeff)), dtype=torch.uint8).view(-1, 32).to(device)
        return Csr(torch.tensor(ptr, dtype=torch.int32, device=device), torch.tensor(col if col else [0], dtype=torch.int32, device=device),
                   cf, len(rows), cols)

    def transpose(self) -> "Csr":
        """cols x rows matrix (for the column sums of R1CStoQAPRelation: At[j] = sum_i L_i A[i][j])."""
        dev = self.col.device
        nnz = int(self.row_ptr[-1].item())
        lens = (self.row_ptr[1:] - self.row_ptr[:-1]).long()
        r = torch.repeat_interleave(torch.arange(self.rows, device=dev), lens)
        c = self.col[:nnz].long()
        order = torch.argsort(c, stable=True)
        counts = torch.bincount(c, minlength=self.cols)
        ptr = torch.zeros(self.cols + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(counts, 0)
        new_col = r[order].to(torch.int32)
        if new_col.numel() == 0:
            new_col = torch.zeros(1, dtype=torch.int32, device=dev)
        coeff = self.coeff[order].contiguous() if self.coeff is not None else None
        return Csr(ptr.to(torch.int32), new_col, coeff, self.cols, self.rows)

    def times(self, ctx: Context, d_z, z_len: int, d_out):
        """d_out[i] = <row i, z> for every row (ozk_fr_spmv_ex_dev)."""
        ctx.fr_spmv_ex_dev(self.row_ptr, self.col, self.coeff, d_z, z_len, self.rows, d_out)


@dataclass
class R1cs:
    """Constraint rows A, B, C (all constraints, or the rows `row_index` of a shard) of a system with num_constraints constraints."""
    A: Csr
    B: Csr
    C: Csr
    num_constraints: int
    num_inputs: int
    num_variables: int
    row_index: Optional[torch.Tensor] = None        # global constraint index of every local row (None: all rows, in order)


def synthetic_r1cs(num_constraints: int, num_inputs: int, device, world: int = 1, rank: int = 0) -> R1cs:
    """The constraint rows rank, rank + world, ... of R1CSConstruction.serialConstruct(numConstraints, numInputs)
    (R1CSConstruction.java:48-104), built directly as CSR on the device with unit coefficients:
        odd i  : x_{i+1} * x_{i+2} = x_{i+3}            even i : (x_{i+1} + x_{i+2}) * x_0 = x_{i+3}
        last   : (sum_{1 <= j <= nv-2} x_j)^2 = x_{nv-1}."""
    nc = num_constraints
    nv = nc + 3
    rows = torch.arange(rank, nc, world, device=device, dtype=torch.int64)
    k = rows.numel()
    last = rows == nc - 1
    odd = (rows % 2 == 1) & ~last
    even = ~odd & ~last
    has_last = bool(last.any().item()) if k else False
    dense = torch.arange(1, nv - 1, device=device, dtype=torch.int64)

    def build(lens, first, second=None):
        ptr = torch.zeros(k + 1, dtype=torch.int64, device=device)
        ptr[1:] = torch.cumsum(lens, 0)
        nnz = int(ptr[-1].item()) if k else 0
        col = torch.zeros(max(nnz, 1), dtype=torch.int64, device=device)
        pos = ptr[:-1]
        normal = ~last
        col[pos[normal]] = first[normal]
        if second is not None:
            col[pos[even] + 1] = second[even]
        if has_last:
            p = int(pos[last].item())
            if lens[last].item() == 1:
                col[p] = first[last]
            else:
                col[p:p + nv - 2] = dense
        return Csr(ptr.to(torch.int32), col.to(torch.int32), None, k, nv)

    one = torch.ones_like(rows)
    lens_a = torch.where(last, (nv - 2) * one, torch.where(odd, one, 2 * one))
    lens_b = torch.where(last, (nv - 2) * one, one)
    A = build(lens_a, rows + 1, rows + 2)
    B = build(lens_b, torch.where(odd, rows + 2, torch.zeros_like(rows)))
    C = build(one, rows + 3)
    return R1cs(A, B, C, nc, num_inputs, nv, rows if world > 1 else None)


# ---- who owns what ------------------------------------------------------------------------------------------------------------
@dataclass
class ProverShards:
    """Index ranges of one rank (see the module docstring).  world == 1: everything."""
    world: int
    rank: int
    n: int                 # domain size
    num_inputs: int
    num_variables: int

    @property
    def m(self) -> int:    # domain points per rank
        return self.n // self.world

    @property
    def var_range(self):   # variables whose queryA / queryB entries and assignment values this rank multiplies
        return self.rank * self.num_variables // self.world, (self.rank + 1) * self.num_variables // self.world

    @property
    def aux_range(self):   # auxiliary variables (offsets into deltaABC = variables num_inputs + offset)
        na = self.num_variables - self.num_inputs
        return self.rank * na // self.world, (self.rank + 1) * na // self.world

    def h_blocks(self):
        """[(first global coefficient index, length)] of the H coefficients the sharded witness map leaves on this rank, in local
        order: block k1 holds h[k1 m + rank c + t], t < c = m / world (distributed.witness_map_distributed)."""
        if self.world == 1:
            return [(0, self.n)]
        c = self.m // self.world
        return [(k1 * self.m + self.rank * c, c) for k1 in range(self.world)]

    def input_rows(self, num_constraints: int):
        """(local row index, variable index) of the rows num_constraints + i that carry primary input i in the evaluation vector
        of A (R1CStoQAP.java:151-153), restricted to this rank's cyclic shard: two aligned ranges (start, stop, step)."""
        g, d = self.world, self.rank
        first = num_constraints + ((d - num_constraints) % g)        # smallest global row >= num_constraints congruent to d
        stop = num_constraints + self.num_inputs
        if first >= stop:
            return None
        count = (stop - first + g - 1) // g
        return (first - d) // g, first - num_constraints, count, g


# ---- keys -----------------------------------------------------------------------------------------------------------------------
@dataclass
class ProvingKeyDev:
    """SerialSetup's ProvingKey (zk_proof_systems/zkSNARK/objects/ProvingKey.java:16-47) with the query vectors resident on the
    device as persistent bases (this rank's slices when sharded) and the five single elements as host integers."""
    alphaG1: tuple
    betaG1: tuple
    betaG2: tuple
    deltaG1: tuple
    deltaG2: tuple
    queryA: Bases
    queryB1: Bases
    queryB2: Bases
    deltaABC: Bases
    queryH: Bases
    shards: ProverShards
    info: dict

    def free(self):
        for k in (self.queryA, self.queryB1, self.queryB2, self.deltaABC, self.queryH):
            k.free()


class DeviceGroth16:
    def __init__(self, ctx: Context, group=None, exchange: Optional[D.PeerExchange] = None):
        self.ctx, self.group, self.exchange = ctx, group, exchange
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device("cuda", ctx.device)
        self.msm = VariableBaseMSM(ctx)
        self.ops = D.GpuOps(ctx)
        self.timings = {}

    # ---- small helpers
    def _mul(self, p, s: int):
        return self.msm.serialMSM([s % R], [p])

    def _vec(self, n: int):
        return torch.empty((n, 32), dtype=torch.uint8, device=self.device)

    def _nonzero(self, v) -> int:
        return int((v != 0).any(dim=1).sum().item())

    def shards(self, r1cs: R1cs) -> ProverShards:
        return ProverShards(self.world, self.rank, domain_size(r1cs.num_constraints, r1cs.num_inputs), r1cs.num_inputs, r1cs.num_variables)

    # ---- setup
    def qap_relation(self, full: R1cs, t: int):
        """R1CStoQAP.R1CStoQAPRelation (R1CStoQAP.java:38-97) on the device: At, Bt, Ct (num_variables elements each) and
        Ht = [t^k, k <= n] as (.., 32) uint8 tensors, and Z(t)."""
        assert full.row_index is None, "the QAP relation needs all constraint rows"
        ctx = self.ctx
        nc, ni, nv = full.num_constraints, full.num_inputs, full.num_variables
        n = domain_size(nc, ni)
        omega = pow(FR_ROOT, R // n, R)                                          # Fp.rootOfUnity, Fp.java:98-102
        lag = self._vec(n)
        ctx.fr_lagrange_dev(lag, n, _le32(t % R), _le32(omega))                  # SerialFFT.lagrangeCoefficients, :54-56
        out = []
        for k, M in enumerate((full.A, full.B, full.C)):
            v = self._vec(nv)
            M.transpose().times(ctx, lag, nc, v)                                 # v[j] = sum_i L_i M[i][j]   (:62-80)
            if k == 0:
                one = _le32(1)
                ctx.fr_lincomb_dev(v[:ni], ni, v[:ni], one, lag[nc:nc + ni], one)    # At[i] += L_{nc+i}, i < numInputs (:57-60)
            out.append(v)
        ones = torch.zeros((n + 1, 32), dtype=torch.uint8, device=self.device)
        ones[:, 0] = 1
        Ht = self._vec(n + 1)
        ctx.fr_scale_powers_dev(ones, Ht, n + 1, None, _le32(t % R))             # Ht[k] = t^k (:82-88)
        del ones
        return out[0], out[1], out[2], Ht, (pow(t, n, R) - 1) % R, n

    def setup(self, full: R1cs, scalar_size_g1: int = 253, scalar_size_g2: int = 254, keep_vk: bool = True):
        """SerialSetup.generate / DistributedSetup.generate: returns (ProvingKeyDev, vk dict).  Every rank computes the (cheap) field
        part and encodes only the slices of the query vectors it will multiply (ProverShards).  scalarSize is generator.bitSize()
        of the Java representative (253 / 254 for the seed-10 generators)."""
        ctx = self.ctx
        nc, ni, nv = full.num_constraints, full.num_inputs, full.num_variables
        t = alpha = beta = gamma = delta = fr_random()
        inv_gamma, inv_delta = pow(gamma, -1, R), pow(delta, -1, R)
        At, Bt, Ct, Ht, Zt, n = self.qap_relation(full, t)
        sh = ProverShards(self.world, self.rank, n, ni, nv)
        non_zero_at, non_zero_bt = self._nonzero(At), self._nonzero(Bt)
        abc = self._vec(nv)
        ctx.fr_lincomb_dev(abc, nv, At, _le32(beta), Bt, _le32(alpha), Ct, _le32(1))      # SerialSetup.java:66-70
        del Ct
        delta_abc = self._vec(nv - ni)
        ctx.fr_scale_dev(abc[ni:], delta_abc, nv - ni, _le32(inv_delta))                  # :79-85
        gamma_abc = self._vec(ni)
        ctx.fr_scale_dev(abc[:ni], gamma_abc, ni, _le32(inv_gamma))                       # :72-77
        del abc
        ctx.fr_scale_dev(Ht, Ht, n + 1, _le32(Zt * inv_delta % R))                        # queryH scalars, :146-150
        g1, g2 = self._mul(G1_ONE, fr_random()), self._mul(G2_ONE, fr_random())           # BNG1.random / BNG2.random
        w1 = FixedBaseMSM.getWindowSize(non_zero_at + non_zero_bt + nv, g1)               # :92-94
        w2 = FixedBaseMSM.getWindowSize(non_zero_bt, g2)                                  # :96-98
        oc1, oc2 = (scalar_size_g1 + w1 - 1) // w1, (scalar_size_g2 + w2 - 1) // w2
        g1b = b"".join(_le32(v) for v in g1)
        g2b = b"".join(_le32(v) for f in g2 for v in f)

        def encode(scalars, group: int) -> Bases:
            """fixed-base batch of a slice of scalars -> persistent bases"""
            cnt = scalars.shape[0]
            if cnt == 0:                                                                  # a rank may own nothing of a tiny circuit
                scalars = torch.zeros((1, 32), dtype=torch.uint8, device=self.device)
                cnt = 1
            scalars = scalars.contiguous()
            wire = torch.empty((cnt, 96 if group == 1 else 192), dtype=torch.uint8, device=self.device)
            if group == 1:
                ctx.fixed_g1_dev(g1b, scalars, cnt, oc1, w1, wire)
            else:
                ctx.fixed_g2_dev(g2b, scalars, cnt, oc2, w2, wire)
            key = ctx.upload_bases(group, wire, cnt, device=True)
            del wire
            return key

        lo, hi = sh.var_range
        alo, ahi = sh.aux_range
        h_sel = torch.cat([Ht[f:f + c] for f, c in sh.h_blocks()]) if self.world > 1 else Ht[:n]
        pk = ProvingKeyDev(
            alphaG1=self._mul(g1, alpha), betaG1=self._mul(g1, beta), betaG2=self._mul(g2, beta),
            deltaG1=self._mul(g1, delta), deltaG2=self._mul(g2, delta),
            queryA=encode(At[lo:hi], 1), queryB1=encode(Bt[lo:hi], 1), queryB2=encode(Bt[lo:hi], 2),
            deltaABC=encode(delta_abc[alo:ahi], 1), queryH=encode(h_sel, 1), shards=sh,
            info={"g1": g1, "g2": g2, "windowSizeG1": w1, "windowSizeG2": w2, "nonZeroAt": non_zero_at, "nonZeroBt": non_zero_bt,
                  "domain": n})
        vk = None
        if keep_vk:
            wire = torch.empty((ni, 96), dtype=torch.uint8, device=self.device)
            ctx.fixed_g1_dev(g1b, gamma_abc, ni, oc1, w1, wire)
            ctx.sync()
            wb = wire.cpu().numpy().tobytes()
            vk = {"gammaG2": self._mul(g2, gamma), "deltaG2": pk.deltaG2,
                  "gammaABCG1": [tuple(int.from_bytes(wb[96 * i + 32 * k:96 * i + 32 * k + 32], "little") for k in range(3)) for i in range(ni)]}
        return pk, vk

    # ---- prover
    def witness_map(self, local: R1cs, d_z):
        """R1CStoQAPWitness: this rank's shard of the coefficients of H (n // world elements; natural order on one GPU, blocked layout
        when sharded), device resident.  `local` holds the constraint rows rank, rank + world, ...; d_z the full assignment."""
        ctx = self.ctx
        nc, ni, nv = local.num_constraints, local.num_inputs, local.num_variables
        sh = self.shards(local)
        m = sh.m
        vecs = []
        for k, M in enumerate((local.A, local.B, local.C)):
            v = torch.zeros((m, 32), dtype=torch.uint8, device=self.device)
            M.times(ctx, d_z, nv, v)                                                      # a_i = <A_i, z> (R1CStoQAP.java:143-160)
            if k == 0:
                ir = sh.input_rows(nc)
                if ir is not None:                                                        # inputs appended to A (:151-153)
                    row0, var0, count, step = ir
                    v[row0:row0 + count] = d_z[var0:var0 + count * step:step]
            vecs.append(v.view(-1))
        return D.witness_map_distributed(self.ops, vecs[0], vecs[1], vecs[2], sh.n, group=self.group, exchange=self.exchange).view(m, 32)

    def prove(self, pk: ProvingKeyDev, local: R1cs, d_z):
        """SerialProver.prove / DistributedProver.prove: the proof (A in G1, B in G2, C in G1) as integer Jacobian triples, the same
        on every rank.  d_z: (num_variables, 32) uint8 device tensor, the full assignment (primary then auxiliary)."""
        ctx, sh = self.ctx, pk.shards
        ni, nv = local.num_inputs, local.num_variables
        h = self.witness_map(local, d_z)
        lo, hi = sh.var_range
        alo, ahi = sh.aux_range
        parts = bytearray(576)
        parts[0:96] = ctx.msm_keyed(d_z[lo:hi], pk.queryA, hi - lo, device=True)                               # evaluationAt
        parts[96:384] = ctx.msm_g1g2_keyed(d_z[lo:hi], pk.queryB1, pk.queryB2, hi - lo, device=True)           # evaluationBt (G1, G2)
        parts[384:480] = ctx.msm_keyed(d_z[ni + alo:ni + ahi], pk.deltaABC, ahi - alo, device=True)            # evaluationABC, witness part
        parts[480:576] = ctx.msm_keyed(h, pk.queryH, h.shape[0], device=True)                                  # evaluationHt (H[n] = 0 is skipped)
        if self.world > 1:
            mine = torch.frombuffer(parts, dtype=torch.uint8).to(self.device)
            allp = torch.empty((self.world, 576), dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allp.view(-1), mine, group=self.group)
            parts = bytearray(576)
            for off, size, grp in ((0, 96, 1), (96, 96, 1), (192, 192, 2), (384, 96, 1), (480, 96, 1)):
                parts[off:off + size] = ctx.sum_points_dev(grp, allp[:, off:off + size].contiguous(), self.world)

        def g1(off):
            return tuple(int.from_bytes(parts[off + 32 * k:off + 32 * k + 32], "little") for k in range(3))

        ev_at, ev_b1, ev_abc, ev_h = g1(0), g1(96), g1(384), g1(480)
        v = [int.from_bytes(parts[192 + 32 * k:192 + 32 * k + 32], "little") for k in range(6)]
        ev_b2 = ((v[0], v[1]), (v[2], v[3]), (v[4], v[5]))
        r = s = fr_random()
        msm = self.msm.serialMSM
        # SerialProver.java:106-114: A = alpha + evAt + r delta, B = beta + evBt + s delta (G1 and G2), C = evABC + s A + r B1 - r s delta.
        # B1 only enters C, so C is expanded over the independent points (the delta terms add up to + r s delta): three small MSMs.
        A = msm([1, 1, r], [pk.alphaG1, ev_at, pk.deltaG1])
        B2 = msm([1, 1, s], [pk.betaG2, ev_b2, pk.deltaG2])
        C = msm([1, 1, s, s, r, r, r * s % R], [ev_abc, ev_h, pk.alphaG1, ev_at, pk.betaG1, ev_b1, pk.deltaG1])
        return A, B2, C
