"""Host driver restating the reference's serial Groth16 setup and prover on top of the GPU operator mirror
(SURVEY.md section 8f-1): the callers either side of the hot path, so that a whole proof can be produced with every MSM
and every FFT on the GPU and compared point for point with the reference's result.

Mirrors (reference paths under src/main/java/):
  SerialSetup.generate     zk_proof_systems/zkSNARK/SerialSetup.java:32-192   (pairing value alphaG1betaG2 omitted)
  SerialProver.prove       zk_proof_systems/zkSNARK/SerialProver.java:26-119
  R1CStoQAPRelation        reductions/r1cs_to_qap/R1CStoQAP.java:38-97       (host-side field loops, as in the Java)
  R1CStoQAPWitness         reductions/r1cs_to_qap/R1CStoQAP.java:126-238     (7 transforms on the GPU, device resident)
  R1CSConstruction.serialConstruct  profiler/generation/R1CSConstruction.java:31-110 (synthetic circuit)
  DistributedProver.prove  zk_proof_systems/zkSNARK/DistributedProver.java:28-167  (prove_distributed: one rank per GPU)
Every random() is Fp.random(seed 10) as in the reference's Configuration (configuration/Configuration.java:52).

The O(n) field loops of the Java also run on the GPU where they sit on the path: the constraint rows times the assignment
(ozk_fr_spmv_dev) in the prover and the Lagrange coefficients (ozk_fr_lagrange_dev) in the setup; all group
arithmetic and all transforms go through liboctozk.  Single scalar multiplications and additions of the Java
(AbstractGroup.mul / add, SerialProver.java:67,106-114) are issued as tiny MSMs."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from .algebra import FR_MODULUS as R
from .algebra import FR_MULTIPLICATIVE_GENERATOR, FR_ROOT, FixedBaseMSM, SerialFFT, VariableBaseMSM, _le32
from .lib import Context

G1_ONE = (1, 2, 1)                      # BN254aG1Parameters.java:24
G2_ONE = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531),
          (1, 0))                        # BN254aG2Parameters.java:25-32


def java_random_next_long(seed: int) -> int:
    """new java.util.Random(seed).nextLong()."""
    mult, mask = 0x5DEECE66D, (1 << 48) - 1
    s = (seed ^ mult) & mask
    out = []
    for _ in range(2):
        s = (s * mult + 0xB) & mask
        v = s >> 16
        out.append(v - (1 << 32) if v >= 1 << 31 else v)
    v = ((out[0] << 32) + out[1]) & ((1 << 64) - 1)
    return v - (1 << 64) if v >= 1 << 63 else v


def fr_random(seed: int = 10) -> int:
    """Fp.random(seed) (algebra/fields/Fp.java:72-80)."""
    return java_random_next_long(seed) % R


def bit_size(p) -> int:
    """BNG1.bitSize / BNG2.bitSize (BNG1.java:174-176): the largest coordinate bit length."""
    flat = [c for f in p for c in (f if isinstance(f, (tuple, list)) else (f,))]
    return max(v.bit_length() for v in flat)


class Groth16:
    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.msm = VariableBaseMSM(ctx)
        self.fixed = FixedBaseMSM(ctx)

    # ---- small group helpers (AbstractGroup.mul / add on the GPU)
    def mul(self, p, s: int):
        return self.msm.serialMSM([s % R], [p])

    def add(self, *pts):
        return self.msm.serialMSM([1] * len(pts), list(pts))

    def random_g1(self, seed: int = 10):
        """BNG1.random (BNG1.java:125-127) = one().mul(Fr.random(seed)).  NOTE: the Jacobian representative differs from
        the Java's double-and-add result, so bitSize() must not be taken from it (see scalar sizes below)."""
        return self.mul(G1_ONE, fr_random(seed))

    def random_g2(self, seed: int = 10):
        return self.mul(G2_ONE, fr_random(seed))

    # ---- synthetic circuit
    @staticmethod
    def serial_construct(num_constraints: int, num_inputs: int):
        num_auxiliary = 3 + num_constraints - num_inputs
        num_variables = num_inputs + num_auxiliary
        a = fr_random()
        b = fr_random()
        full = [1, a, b]
        cons = []
        for i in range(num_constraints - 1):
            if i % 2 != 0:
                A, B, C = [(i + 1, 1)], [(i + 2, 1)], [(i + 3, 1)]
                tmp = a * b % R
            else:
                A, B, C = [(i + 1, 1), (i + 2, 1)], [(0, 1)], [(i + 3, 1)]
                tmp = (a + b) % R
            a, b = b, tmp
            full.append(tmp)
            cons.append((A, B, C))
        lc = [(i, 1) for i in range(1, num_variables - 1)]
        res = sum(full[1:num_variables - 1]) % R
        full.append(res * res % R)
        cons.append((lc, list(lc), [(num_variables - 1, 1)]))
        return cons, num_inputs, num_auxiliary, full[:num_inputs], full[num_inputs:]

    # ---- R1CS -> QAP
    @staticmethod
    def _evaluate(lc, assignment) -> int:
        acc = 0
        for idx, val in lc:
            acc += val * assignment[idx]
        return acc % R

    @staticmethod
    def _csr(rows):
        """CSR arrays (row_ptr, col, coefficient list) of a list of linear combinations [(index, value), ...]."""
        import numpy as np
        row_ptr = np.zeros(len(rows) + 1, dtype=np.uint32)
        cols, coeffs = [], []
        for i, lc in enumerate(rows):
            for idx, val in lc:
                cols.append(idx)
                coeffs.append(val % R)
            row_ptr[i + 1] = len(cols)
        return row_ptr, np.asarray(cols if cols else [0], dtype=np.uint32), coeffs

    def r1cs_to_qap_relation(self, cons, num_inputs, num_variables, t):
        num_constraints = len(cons)
        dom = SerialFFT(self.ctx, num_constraints + num_inputs)
        At, Bt, Ct = [0] * num_variables, [0] * num_variables, [0] * num_variables
        lag = dom.lagrangeCoefficients(t)                      # GPU (SerialFFT.lagrangeCoefficients, SerialFFT.java:126-128)
        for i in range(num_inputs):
            At[i] = lag[num_constraints + i]
        for i, (A, B, C) in enumerate(cons):
            li = lag[i]
            for idx, val in A:
                At[idx] = (At[idx] + li * val) % R
            for idx, val in B:
                Bt[idx] = (Bt[idx] + li * val) % R
            for idx, val in C:
                Ct[idx] = (Ct[idx] + li * val) % R
        Ht, ti = [], 1
        for _ in range(dom.domainSize + 1):
            Ht.append(ti)
            ti = ti * t % R
        return {"At": At, "Bt": Bt, "Ct": Ct, "Ht": Ht, "Zt": dom.computeZ(t), "degree": dom.domainSize}

    def r1cs_to_qap_witness(self, cons, num_inputs, primary, auxiliary) -> List[int]:
        """R1CStoQAPWitness with the seven transforms on the GPU; A, B, C and H never leave the device between them."""
        ctx = self.ctx
        num_constraints = len(cons)
        g = FR_MULTIPLICATIVE_GENERATOR
        dom = SerialFFT(ctx, num_constraints + num_inputs)
        n = dom.domainSize
        full = list(primary) + list(auxiliary)
        dev = torch.device("cuda", ctx.device)

        def up(v):
            return torch.frombuffer(bytearray(b"".join(_le32(x) for x in v)), dtype=torch.uint8).to(dev)

        # a_i, b_i, c_i = <row i, assignment> on the GPU (ozk_fr_spmv_dev; R1CStoQAP.java:143-160), inputs appended to A (:151-153)
        d_z = up(full)
        dA, dB, dC = (torch.zeros(n * 32, dtype=torch.uint8, device=dev) for _ in range(3))
        for which, d in ((0, dA), (1, dB), (2, dC)):
            rp, col, cf = self._csr([c[which] for c in cons])
            ctx.fr_spmv_dev(torch.from_numpy(rp).to(dev), torch.from_numpy(col).to(dev), up(cf) if cf else torch.zeros(32, dtype=torch.uint8, device=dev),
                            d_z, num_constraints, d)
        dA.view(n, 32)[num_constraints:num_constraints + num_inputs] = d_z.view(-1, 32)[:num_inputs]

        w, winv = _le32(dom.omega), _le32(pow(dom.omega, -1, R))
        ninv = _le32(pow(n, -1, R))
        for d in (dA, dB, dC):
            ctx.ntt_ex_dev(d, d, n, winv, None, ninv, None)                 # radix2InverseFFT
            ctx.ntt_ex_dev(d, d, n, w, _le32(g), None, None)                # radix2CosetFFT
        # pointwise A*B - C on the device (R1CStoQAP.java:180-214); nothing returns to the host between the transforms
        dH = dA
        ctx.fr_mul_sub_dev(dA, dB, dC, dH, n)
        inv_z = pow(dom.computeZ(g), -1, R)                                 # divideByZOnCoset folded into the post-scale
        scale = pow(n, -1, R) * inv_z % R
        ctx.ntt_ex_dev(dH, dH, n, winv, None, _le32(scale), _le32(pow(g, -1, R)))   # (divide by Z) + radix2CosetInverseFFT
        ctx.sync()
        hb = dH.cpu().numpy().tobytes()
        H = [int.from_bytes(hb[32 * i:32 * i + 32], "little") for i in range(n)]
        H.append(0)
        return H

    # ---- setup
    def setup(self, cons, num_inputs, num_variables, scalar_size_g1: int = 253, scalar_size_g2: int = 254):
        """SerialSetup.generate.  scalarSize is generator.bitSize() of the *Java* representative (253 / 254 for the seed-10
        generators); it only matters when outerc * window < 254."""
        t = alpha = beta = gamma = delta = fr_random()
        inv_gamma, inv_delta = pow(gamma, -1, R), pow(delta, -1, R)
        qap = self.r1cs_to_qap_relation(cons, num_inputs, num_variables, t)
        abc = [(beta * qap["At"][i] + alpha * qap["Bt"][i] + qap["Ct"][i]) % R for i in range(num_variables)]
        gammaABC = [abc[i] * inv_gamma % R for i in range(num_inputs)]
        deltaABC = [abc[i] * inv_delta % R for i in range(num_inputs, num_variables)]
        non_zero_at = sum(1 for v in qap["At"] if v)
        non_zero_bt = sum(1 for v in qap["Bt"] if v)
        g1, g2 = self.random_g1(), self.random_g2()
        w1 = FixedBaseMSM.getWindowSize(non_zero_at + non_zero_bt + num_variables, g1)
        w2 = FixedBaseMSM.getWindowSize(non_zero_bt, g2)
        fb = self.fixed
        inverse_delta_zt = qap["Zt"] * inv_delta % R
        Ht = [h * inverse_delta_zt % R for h in qap["Ht"]]
        pk = {
            "alphaG1": self.mul(g1, alpha), "betaG1": self.mul(g1, beta), "betaG2": self.mul(g2, beta),
            "deltaG1": self.mul(g1, delta), "deltaG2": self.mul(g2, delta),
            "deltaABCG1": fb.batchMSM(scalar_size_g1, w1, g1, deltaABC),
            "queryA": fb.batchMSM(scalar_size_g1, w1, g1, qap["At"]),
            "queryB": fb.doubleBatchMSM(scalar_size_g1, w1, scalar_size_g2, w2, g1, g2, qap["Bt"]),
            "queryH": fb.batchMSM(scalar_size_g1, w1, g1, Ht),
        }
        vk = {"gammaG2": self.mul(g2, gamma), "deltaG2": pk["deltaG2"], "gammaABCG1": fb.batchMSM(scalar_size_g1, w1, g1, gammaABC)}
        return pk, vk, {"g1": g1, "g2": g2, "windowSizeG1": w1, "windowSizeG2": w2}

    # ---- prover
    def prove(self, pk, cons, num_inputs, primary: Sequence[int], auxiliary: Sequence[int]):
        """SerialProver.prove: returns ((A, B, C), coefficientsH)."""
        msm = self.msm
        H = self.r1cs_to_qap_witness(cons, num_inputs, primary, auxiliary)
        r = s = fr_random()
        num_variables = len(primary) + len(auxiliary)
        rs_delta = self.mul(pk["deltaG1"], r * s % R)
        qa, qb = pk["queryA"], pk["queryB"]
        ev_at = self.add(msm.serialMSM(primary, qa[:num_inputs]), msm.serialMSM(auxiliary, qa[num_inputs:num_variables]))
        bp1, bp2 = msm.doubleMSM(primary, qb[:num_inputs])
        bw1, bw2 = msm.doubleMSM(auxiliary, qb[num_inputs:num_variables])
        ev_b1, ev_b2 = self.add(bp1, bw1), self.add(bp2, bw2)
        ev_h = msm.serialMSM(H, pk["queryH"])
        num_witness = num_variables - num_inputs
        ev_abc = self.add(msm.serialMSM(auxiliary[:num_witness], pk["deltaABCG1"][:num_witness]), ev_h)
        A = self.add(pk["alphaG1"], ev_at, self.mul(pk["deltaG1"], r))
        B1 = self.add(pk["betaG1"], ev_b1, self.mul(pk["deltaG1"], s))
        B2 = self.add(pk["betaG2"], ev_b2, self.mul(pk["deltaG2"], s))
        neg_rs_delta = (rs_delta[0], (-rs_delta[1]) % FQ, rs_delta[2])
        C = self.add(ev_abc, self.mul(A, s), self.mul(B1, r), neg_rs_delta)
        return (A, B2, C), H


    # ---- distributed prover
    def prove_distributed(self, pk, cons, num_inputs, primary: Sequence[int], auxiliary: Sequence[int], group=None, exchange=None):
        """DistributedProver.prove (zk_proof_systems/zkSNARK/DistributedProver.java:28-167) on the ranks of a torch.distributed
        group, one GPU each, replacing the Spark RDD plumbing: every rank evaluates the constraints of its cyclic shard of the
        domain, the seven transforms run sharded (distributed.witness_map_distributed: four-step transforms whose exchange is
        fused into the kernels), every MSM runs on a contiguous shard of (scalar, base) pairs followed by a gather of the
        partial sums.  The H coefficients come out in the blocked layout, so each rank pairs them with the matching entries of
        queryH.  Returns the same ((A, B, C)) on every rank -- the same proof as `prove` -- and this rank's H shard."""
        import torch.distributed as dist

        from . import distributed as D
        ctx = self.ctx
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        ops = D.GpuOps(ctx)
        dev = torch.device("cuda", ctx.device)
        num_constraints = len(cons)
        n = SerialFFT(ctx, num_constraints + num_inputs).domainSize
        assert n % (world * world) == 0
        m, c = n // world, n // world // world
        full = list(primary) + list(auxiliary)
        num_variables = len(full)

        def up(v):
            return torch.frombuffer(bytearray(b"".join(_le32(x) for x in v)), dtype=torch.uint8).to(dev)

        # this rank's rows of the QAP evaluation vectors: domain index i = rank + world * i2 (R1CStoQAP.java:143-160)
        A, B, C = [0] * m, [0] * m, [0] * m
        for i2 in range(m):
            i = rank + world * i2
            if i < num_constraints:
                a, b, cc = cons[i]
                A[i2], B[i2], C[i2] = self._evaluate(a, full), self._evaluate(b, full), self._evaluate(cc, full)
            elif i < num_constraints + num_inputs:
                A[i2] = full[i - num_constraints]
        h_dev = D.witness_map_distributed(ops, up(A), up(B), up(C), n, group=group, exchange=exchange)
        ctx.sync()
        # global coefficient index of local element [k1][t]: k1 * m + rank * c + t (H[n] = 0 contributes nothing)
        h_index = [k1 * m + rank * c + t for k1 in range(world) for t in range(c)]
        qh = [pk["queryH"][i] for i in h_index]

        def shard(seq):
            lo, hi = rank * len(seq) // world, (rank + 1) * len(seq) // world
            return seq[lo:hi]

        def dmsm(scalars, bases, g2=False):
            """sum over all ranks of this rank's (scalars, bases): local MSM, gather of the partial sums, one small sum"""
            k = len(scalars)
            if g2:
                local = self.msm.serialMSM(list(scalars), list(bases)) if k else ((0, 0), (1, 0), (0, 0))
                packed = b"".join(_le32(v) for f in local for v in f)
            else:
                local = self.msm.serialMSM(list(scalars), list(bases)) if k else (0, 1, 0)
                packed = b"".join(_le32(v) for v in local)
            if world == 1:
                return local
            mine = torch.frombuffer(bytearray(packed), dtype=torch.uint8).to(dev)
            gathered = torch.empty(world * len(packed), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(gathered, mine, group=group)
            out = ctx.sum_points_dev(2 if g2 else 1, gathered, world)
            vals = [int.from_bytes(out[32 * i:32 * i + 32], "little") for i in range(len(out) // 32)]
            return ((vals[0], vals[1]), (vals[2], vals[3]), (vals[4], vals[5])) if g2 else tuple(vals)

        r = s = fr_random()
        qa, qb = pk["queryA"], pk["queryB"]
        ev_at = dmsm(shard(full), shard(qa[:num_variables]))
        ev_b1 = dmsm(shard(full), shard([q[0] for q in qb[:num_variables]]))
        ev_b2 = dmsm(shard(full), shard([q[1] for q in qb[:num_variables]]), g2=True)
        hb = h_dev.cpu().numpy().tobytes()
        h_local = [int.from_bytes(hb[32 * i:32 * i + 32], "little") for i in range(m)]
        ev_h = dmsm(h_local, qh)
        num_witness = num_variables - num_inputs
        ev_w = dmsm(shard(list(auxiliary[:num_witness])), shard(pk["deltaABCG1"][:num_witness]))
        ev_abc = self.add(ev_w, ev_h)
        rs_delta = self.mul(pk["deltaG1"], r * s % R)
        A_pt = self.add(pk["alphaG1"], ev_at, self.mul(pk["deltaG1"], r))
        B1 = self.add(pk["betaG1"], ev_b1, self.mul(pk["deltaG1"], s))
        B2 = self.add(pk["betaG2"], ev_b2, self.mul(pk["deltaG2"], s))
        neg_rs_delta = (rs_delta[0], (-rs_delta[1]) % FQ, rs_delta[2])
        C_pt = self.add(ev_abc, self.mul(A_pt, s), self.mul(B1, r), neg_rs_delta)
        return (A_pt, B2, C_pt), h_local


FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
