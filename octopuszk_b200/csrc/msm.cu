// Variable-base MSM driver: scalar decomposition, counting sort of (window, bucket) pairs, and the C ABI.
//
// Replaces Java_algebra_msm_VariableBaseMSM_variableBaseSerialMSMNativeHelper / ...DoubleMSMNativeHelper and the host
// loops pippengerMSMG1 / pippengerMSMG2 (algebra_msm_VariableBaseMSM.cu:1246-1788).  Differences that do not change
// the result as a group element: signed digits (half the buckets), all windows processed at once, a counting sort of
// 4-byte point indices instead of scattering whole 192-byte points per window (:758), no host round trips.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "msm_ba_impl.cuh"
#include "msm_impl.cuh"

namespace ozk {

// ---- scalar digits ---------------------------------------------------------------------------------------------
// Signed c-bit recoding, least significant window first: d = bits + carry; if d > 2^(c-1): d -= 2^c, carry = 1.
// So |d| <= 2^(c-1): bucket index |d| - 1 in [0, 2^(c-1)), sign separate.  Scalars are < r < 2^254, and
// nwin = ceil(255 / c) leaves room for the last carry.
__device__ __forceinline__ uint32_t scalar_bits(const uint32_t (&s)[8], uint32_t pos, uint32_t c) {
    if (pos >= 256) return 0;
    const uint32_t limb = pos >> 5, off = pos & 31;
    uint64_t v = s[limb];
    if (limb + 1 < 8) v |= (uint64_t)s[limb + 1] << 32;
    return (uint32_t)(v >> off) & ((1u << c) - 1);
}

__device__ __forceinline__ bool scalar_lt_r(const uint32_t (&s)[8]) {
    for (int i = 7; i >= 0; i--) {
        const uint32_t m = FrParams::mod(i);
        if (s[i] < m) return true;
        if (s[i] > m) return false;
    }
    return false;
}

// MODE 0: histogram (count[w * nb + b] += 1).  MODE 1: scatter (sorted[w * n + cursor++] = i | sign << 31).
// Only windows [win_lo, win_hi) are emitted by one launch (the carries of the lower windows are still computed): the
// scatter is launched per group of windows small enough that the live 32-byte store sectors (one per bucket cursor) stay
// in L2, so the 4-byte scattered stores combine there instead of going to DRAM one sector at a time.
// The atomic of the warp's first emitting lane is aggregated: every lane that hits the same (window, bucket) as that lane
// is served by one atomicAdd of the group's size and takes consecutive positions (two ballots and a shuffle; a full
// match.any cost 1 ms per 2^24 pairs).  With uniform scalars the group is the one lane; with skewed scalars -- the reference
// profiler's input (VariableBaseMSMProfiling.java:27-31: half of all scalars share every high digit), 0/1-heavy witnesses,
// a top window with one or two significant bits -- the dominant bucket of a window is in nearly every warp's first lane
// group, which turns up to 32 same-address atomics into one and the group's stores into one coalesced run.
template <int MODE>
__global__ void __launch_bounds__(256) msm_digits(const uint4* __restrict__ scalars, size_t n, size_t wstride, uint32_t c, uint32_t nwin,
                                                  uint32_t win_lo, uint32_t win_hi,
                                                  uint32_t* __restrict__ count_or_cursor, uint32_t* __restrict__ sorted,
                                                  uint32_t* flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;                      // no early return: every lane takes part in the warp votes below
    const uint32_t lane = threadIdx.x & 31;
    uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (valid) {
        uint4 a = scalars[2 * i], b = scalars[2 * i + 1];
        s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
        s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
        if (MODE == 0 && win_lo == 0 && !scalar_lt_r(s)) atomicOr(flag, 2u);
    }
    const uint32_t half = 1u << (c - 1);
    const uint32_t log_nb = c - 1;
    uint32_t carry = 0;
    (void)nwin;
    for (uint32_t w = 0; w < win_hi; w++) {
        uint32_t d = scalar_bits(s, w * c, c) + carry;
        uint32_t neg = 0;
        carry = 0;
        if (d > half) {
            d = (1u << c) - d;
            neg = 1;
            carry = 1;
        }
        if (w < win_lo) continue;                  // uniform across the warp
        const bool emit = valid && d != 0;
        const unsigned act = __ballot_sync(0xffffffffu, emit);
        if (act == 0) continue;
        const uint32_t slot = (w << log_nb) + (d - 1);
        // the first emitting lane's slot is the candidate hot bucket of this warp
        const uint32_t first = __ffs(act) - 1;
        const uint32_t slot0 = __shfl_sync(0xffffffffu, slot, first);
        const bool in_group = emit && slot == slot0;
        const unsigned group = __ballot_sync(0xffffffffu, in_group);
        uint32_t base = 0;
        if (lane == first) base = atomicAdd(&count_or_cursor[slot0], (uint32_t)__popc(group));
        if (MODE == 1) base = __shfl_sync(0xffffffffu, base, first);
        if (!emit) continue;
        if (MODE == 0) {
            if (!in_group) atomicAdd(&count_or_cursor[slot], 1u);
        } else {
            const uint32_t pos = in_group ? base + __popc(group & ((1u << lane) - 1u)) : atomicAdd(&count_or_cursor[slot], 1u);
            sorted[(size_t)w * wstride + pos] = (uint32_t)i | (neg << 31);
        }
    }
}

// One CTA per window: exclusive scan of the bucket counts -> window-local start offsets (also copied to `cursor`),
// and the overflow work list for buckets with more than seg_len entries.  Each thread scans kScanPer consecutive buckets
// per step, so a window of 2^19 buckets takes 64 block-wide steps.
static constexpr int kScanPer = 8;
__global__ void __launch_bounds__(1024) msm_scan(const uint32_t* __restrict__ count, uint32_t nb, uint32_t* __restrict__ start,
                                                 uint32_t* __restrict__ cursor, OvfTask* __restrict__ ovf_tasks,
                                                 uint32_t* __restrict__ ovf_task_count, OvfBucket* __restrict__ ovf_buckets,
                                                 uint32_t* __restrict__ ovf_bucket_count, uint32_t ovf_task_cap, uint32_t ovf_bucket_cap, uint32_t seg_len,
                                                 uint32_t align_mask) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const uint32_t w = blockIdx.x;
    const uint32_t* cnt = count + (size_t)w * nb;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024 * kScanPer) {
        const uint32_t b0 = base + threadIdx.x * kScanPer;
        uint32_t v[kScanPer];
        uint32_t tsum = 0;
#pragma unroll
        for (int k = 0; k < kScanPer; k++) {
            v[k] = (b0 + k) < nb ? cnt[b0 + k] : 0;
            tsum += (v[k] + align_mask) & ~align_mask;          // bucket runs start at multiples of 2^rounds (batch-affine pre-reduction)
        }
        // block-wide exclusive scan of the per-thread sums
        uint32_t x = tsum;
        const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane];
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, ws, off);
                if (lane >= off) ws += y;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        uint32_t excl = carry_s + (wid ? warp_sums[wid - 1] : 0) + x - tsum;
#pragma unroll
        for (int k = 0; k < kScanPer; k++) {
            const uint32_t b = b0 + k;
            if (b < nb) {
                start[(size_t)w * nb + b] = excl;
                cursor[(size_t)w * nb + b] = excl;
                if (v[k] > seg_len) {
                    const uint32_t extra = (v[k] - 1) / seg_len;
                    const uint32_t t0 = atomicAdd(ovf_task_count, extra);
                    const uint32_t kk = atomicAdd(ovf_bucket_count, 1u);
                    if (kk < ovf_bucket_cap && t0 + extra <= ovf_task_cap) {
                        ovf_buckets[kk] = {w * nb + b, t0, extra};
                        for (uint32_t t = 0; t < extra; t++) ovf_tasks[t0 + t] = {w * nb + b, t + 1};
                    }
                }
            }
            excl += (v[k] + align_mask) & ~align_mask;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl;
        __syncthreads();
    }
}

// ---- bucket order -----------------------------------------------------------------------------------------------
// Counting sort of the bucket ids by run length (clamped to seg_len), longest first.  msm_accumulate walks the buckets in
// this order so that the lanes of a warp do the same number of additions (run lengths are ~Poisson, otherwise every
// warp waits for its longest lane), and long runs start first.
static constexpr int kOrderKeys = kSegMax + 1;

__global__ void __launch_bounds__(256) msm_order_hist(const uint32_t* __restrict__ count, uint32_t nbt, uint32_t* __restrict__ ohist, uint32_t seg_len) {
    __shared__ uint32_t h[kOrderKeys];
    for (int k = threadIdx.x; k < kOrderKeys; k += blockDim.x) h[k] = 0;
    __syncthreads();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbt) atomicAdd(&h[seg_len - min(count[b], seg_len)], 1u);
    __syncthreads();
    for (int k = threadIdx.x; k < kOrderKeys; k += blockDim.x)
        if (h[k]) atomicAdd(&ohist[k], h[k]);
}

__global__ void __launch_bounds__(1024) msm_order_scan(const uint32_t* __restrict__ ohist, uint32_t* __restrict__ ocursor) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < (uint32_t)kOrderKeys; base += 1024) {
        const uint32_t k = base + threadIdx.x;
        const uint32_t v = k < (uint32_t)kOrderKeys ? ohist[k] : 0;
        uint32_t x = v;
        const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane];
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, ws, off);
                if (lane >= off) ws += y;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const uint32_t excl = carry_s + (wid ? warp_sums[wid - 1] : 0) + x - v;
        if (k < (uint32_t)kOrderKeys) ocursor[k] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) msm_order_scatter(const uint32_t* __restrict__ count, uint32_t nbt, uint32_t* __restrict__ ocursor,
                                                         uint32_t* __restrict__ order, uint32_t seg_len) {
    __shared__ uint32_t h[kOrderKeys];
    for (int k = threadIdx.x; k < kOrderKeys; k += blockDim.x) h[k] = 0;
    __syncthreads();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t key = 0, rank = 0;
    if (b < nbt) {
        key = seg_len - min(count[b], seg_len);
        rank = atomicAdd(&h[key], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kOrderKeys; k += blockDim.x)
        if (h[k]) h[k] = atomicAdd(&ocursor[k], h[k]);      // h[k] becomes this block's base for key k
    __syncthreads();
    if (b < nbt) order[h[key] + rank] = b;
}

// ---- window choice ---------------------------------------------------------------------------------------------
// Cost model in units of one mixed addition per (point, window), calibrated on B200 at n = 2^24 and 2^26
// (profiles/r1_window_sweep.txt):
//   accumulate : n * (1 + lane imbalance) + ~11.5 per bucket (task set-up, bucket store, two full additions in the reduce)
//   sort       : 0.13 n (measured 5.8-8.5 ms per 2^24 x 13-15 windows for c = 16..20)
//   tail       : a fixed per-window cost
//   hot top    : when the top window holds only t < c - 1 significant bits of the 254-bit scalars, its 2^t buckets each
//                receive n / 2^t points (same-address atomics in the sort, overflow tasks in accumulate); penalised
//                when that is more than 2^15 points per bucket (c = 18, 19 at 2^24: +5 .. +7 ms measured)
// 2^23 .. 2^25 -> c = 17 (15 windows); >= 2^26 -> c = 20 (13 windows: 170 ms against 191 ms at 2^26).
static constexpr uint32_t kMaxWindowBits = 20;
static uint32_t choose_window(size_t n) {
    if (const char* e = getenv("OZK_MSM_WINDOW")) {
        int c = atoi(e);
        if (c >= 2 && c <= (int)kMaxWindowBits) return (uint32_t)c;
    }
    uint32_t best = 2;
    double best_cost = 1e300;
    for (uint32_t c = 2; c <= kMaxWindowBits; c++) {
        const uint32_t nwin_i = (255 + c - 1) / c;
        const double nwin = nwin_i;
        const double nb = (double)(1u << (c - 1));
        const double per_bucket = (double)n / nb;
        const double imbalance = per_bucket >= 1 ? 0.5 / std::sqrt(per_bucket) : 2.0;
        double cost = nwin * ((double)n * (1.0 + imbalance + 0.13) + 11.5 * nb + 2000.0);
        const int top_bits = 254 - (int)((nwin_i - 1) * c);
        if (top_bits > 0 && top_bits < (int)c - 1 && ((double)n / (double)(1u << top_bits)) > 32768.0) cost += 0.35 * (double)n;
        if (cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
}

struct MsmShape {
    uint32_t c, nwin, nb, log_nb;
    uint32_t seg;                      // longest run one accumulate task handles
    uint32_t ovf_task_cap, ovf_bucket_cap;
    uint32_t ba;                       // batch-affine pre-reduction rounds (msm_ba_impl.cuh); 0: off
    size_t wstride;                    // entries per window in the sorted index array (len, plus the alignment padding when ba > 0)
};

// Batch-affine pre-reduction (G1): rounds by slice length.  Each round halves the runs at ~6.6 instead of ~9.5 products per
// addition but costs a pass over HBM, so it pays only while the runs are long.  OZK_MSM_BA = 0..3 overrides.
static uint32_t ba_rounds_for(size_t len, uint32_t c) {
    if (const char* e = getenv("OZK_MSM_BA")) {
        int r = atoi(e);
        if (r >= 0 && r <= 3) return (uint32_t)r;
    }
    (void)len; (void)c;
    return 0;
}

// n_total fixes the window shape (slices of one MSM share their buckets); len is the number of pairs sorted and
// accumulated at a time and fixes the task length and the overflow capacities.
static MsmShape msm_shape(size_t n_total, size_t len) {
    MsmShape s;
    s.c = choose_window(n_total);
    s.nwin = (255 + s.c - 1) / s.c;
    s.log_nb = s.c - 1;
    s.nb = 1u << s.log_nb;
    // task length: long enough to amortise a task, short enough that the serial chain of one task (about 2.3 us per
    // addition) does not dominate small inputs: aim for ~2^17 tasks
    size_t want = ((size_t)s.nwin * len) >> 17;
    s.seg = 32;
    while (s.seg < (uint32_t)kSegMax && s.seg * 2 <= want) s.seg *= 2;
    size_t cap = ((size_t)s.nwin * len) / s.seg + 1;
    s.ovf_task_cap = (uint32_t)std::min<size_t>(cap, 0x7fffffffu);
    s.ovf_bucket_cap = s.ovf_task_cap;
    s.ba = ba_rounds_for(len, s.c);
    if (len >= ((size_t)1 << 31) - 1) s.ba = 0;           // the sentinel must not be a valid entry
    const size_t a = (size_t)1 << s.ba;
    s.wstride = s.ba ? ((len + (size_t)s.nb * (a - 1) + 63) & ~(size_t)63) : len;
    return s;
}
static MsmShape msm_shape(size_t n) { return msm_shape(n, n); }

// buffers in ctx->msm[]
enum { B_AFF1 = 0, B_AFF2, B_COUNT, B_START, B_CURSOR, B_SORTED, B_BUCKETS, B_OVFTASK, B_OVFBUCKET, B_OVFPART, B_SCRATCH, B_MISC, B_ORDER, B_BUCKETS2, B_PAIR0, B_PAIR1 };
// Which accumulate kernel runs (msm_impl.cuh): bit 0 = G1, bit 1 = G2 use the shared-memory-resident one.  Default: G2 only.
// Measured at 2^24 (profiles/r2_sweep_g2_called_products.jsonl, r2_sweep_smem_accumulate_experiment.jsonl), accumulate ms:
//   G2  register-resident, products inlined, 2 CTAs per SM (round 1)   134.0
//       products called                                                120.3
//       products called, 168-register cap for 3 CTAs (516 B of spills) 116.8
//       products called + accumulator in shared memory, 3 CTAs         113.1   <- default
//   G1  register-resident, 4 CTAs per SM                                37.5   <- default
//       accumulator in shared memory, 5 CTAs                            37.3   (within noise; not worth a second code path)
bool msm_use_smem_accumulate(bool g2) {
    const char* e = getenv("OZK_MSM_SMEM");       // read on every call: the parity test switches it inside one process
    const int mask = e ? atoi(e) : 2;
    return (mask >> (g2 ? 1 : 0)) & 1;
}
static constexpr int kMiscOhist = 64, kMiscOcursor = 2048, kMiscWords = 4096;   // word offsets inside B_MISC

// sort phase shared by G1 / G2 / paired calls: fills start/count/sorted and the overflow lists
static int msm_sort(ozk_ctx* ctx, const void* d_scalars, size_t n, const MsmShape& sh) {
    cudaStream_t st = ctx->stream;
    const size_t nbt = (size_t)sh.nwin * sh.nb;
    OZK_TRY(ctx->msm[B_COUNT].reserve(nbt * 4, st));
    OZK_TRY(ctx->msm[B_START].reserve(nbt * 4, st));
    OZK_TRY(ctx->msm[B_CURSOR].reserve(nbt * 4, st));
    OZK_TRY(ctx->msm[B_SORTED].reserve((size_t)sh.nwin * sh.wstride * 4, st));
    OZK_TRY(ctx->msm[B_OVFTASK].reserve((size_t)sh.ovf_task_cap * sizeof(OvfTask), st));
    OZK_TRY(ctx->msm[B_OVFBUCKET].reserve((size_t)sh.ovf_bucket_cap * sizeof(OvfBucket), st));
    OZK_TRY(ctx->msm[B_MISC].reserve(kMiscWords * 4, st));
    OZK_TRY(ctx->msm[B_ORDER].reserve(nbt * 4, st));
    if (sh.ba) OZK_CUDA(cudaMemsetAsync(ctx->msm[B_SORTED].p, 0xff, (size_t)sh.nwin * sh.wstride * 4, st));     // padding = sentinel
    uint32_t* misc = (uint32_t*)ctx->msm[B_MISC].p;      // [0] flag (cleared by the caller), [1] ovf task count, [2] ovf bucket count, run-length histogram
    OZK_CUDA(cudaMemsetAsync(misc + 1, 0, (kMiscWords - 1) * 4, st));
    OZK_CUDA(cudaMemsetAsync(ctx->msm[B_COUNT].p, 0, nbt * 4, st));
    const unsigned grid = (unsigned)((n + 255) / 256);
    // windows per launch: keep (buckets x 32-byte sectors) of one launch within ~24 MB of L2
    uint32_t wpl = (uint32_t)std::max<size_t>(1, ((size_t)24 << 20) / ((size_t)sh.nb * 32));
    if (const char* e = getenv("OZK_MSM_WPL")) wpl = (uint32_t)std::max(1, atoi(e));
    uint32_t nlaunch = 0;
    for (uint32_t w0 = 0; w0 < sh.nwin; w0 += wpl, nlaunch++)
        msm_digits<0><<<grid, 256, 0, st>>>((const uint4*)d_scalars, n, sh.wstride, sh.c, sh.nwin, w0, std::min(sh.nwin, w0 + wpl),
                                            (uint32_t*)ctx->msm[B_COUNT].p, nullptr, misc);
    msm_scan<<<sh.nwin, 1024, 0, st>>>((const uint32_t*)ctx->msm[B_COUNT].p, sh.nb, (uint32_t*)ctx->msm[B_START].p,
                                       (uint32_t*)ctx->msm[B_CURSOR].p, (OvfTask*)ctx->msm[B_OVFTASK].p, misc + 1,
                                       (OvfBucket*)ctx->msm[B_OVFBUCKET].p, misc + 2, sh.ovf_task_cap, sh.ovf_bucket_cap, sh.seg,
                                       (1u << sh.ba) - 1u);
    for (uint32_t w0 = 0; w0 < sh.nwin; w0 += wpl, nlaunch++)
        msm_digits<1><<<grid, 256, 0, st>>>((const uint4*)d_scalars, n, sh.wstride, sh.c, sh.nwin, w0, std::min(sh.nwin, w0 + wpl),
                                            (uint32_t*)ctx->msm[B_CURSOR].p, (uint32_t*)ctx->msm[B_SORTED].p, misc);
    {
        const unsigned og = (unsigned)((nbt + 255) / 256);
        msm_order_hist<<<og, 256, 0, st>>>((const uint32_t*)ctx->msm[B_COUNT].p, (uint32_t)nbt, misc + kMiscOhist, sh.seg);
        msm_order_scan<<<1, 1024, 0, st>>>(misc + kMiscOhist, misc + kMiscOcursor);
        msm_order_scatter<<<og, 256, 0, st>>>((const uint32_t*)ctx->msm[B_COUNT].p, (uint32_t)nbt, misc + kMiscOcursor,
                                              (uint32_t*)ctx->msm[B_ORDER].p, sh.seg);
    }
    ctx->launches += 4 + nlaunch;
    OZK_CUDA(cudaGetLastError());
    return OZK_OK;
}

// Where the bases of one group come from: the wire format (device pointer; converted to affine Montgomery form into
// ctx->msm[aff_slot] first) or an already converted persistent array (ozk_bases_*, below).
struct BaseSrc {
    const void* wire = nullptr;
    const void* affine = nullptr;
    bool any() const { return wire || affine; }
};

// The normalisation of a slice's bases (wire Jacobian -> affine Montgomery: integer-pipe bound, 1.0 ms per 2^24 points with
// Z = 1 and 3.7 ms with a Z per point) needs only the bases, the sort of the same slice (L2-atomic bound, 5.9 ms) only the scalars:
// the conversion is forked onto a high-priority side stream BEFORE the sort is enqueued and joined before the accumulate.
// The fork event sits behind everything already on the main stream, so the conversion cannot overwrite the affine buffer while
// the previous slice's accumulate still reads it.  OZK_MSM_CONVERT_FORK=0 keeps everything on one stream.
static int msm_convert_fork(ozk_ctx* ctx, const MsmLaunch& L, const BaseSrc& src, size_t n, int aff_slot, int group_bit) {
    if (!src.wire) return OZK_OK;
    if (const char* e = getenv("OZK_MSM_CONVERT_FORK")) {
        if (atoi(e) == 0) return OZK_OK;
    }
    cudaStream_t st = ctx->stream, side = ctx->side_stream;
    if (ctx->msm[aff_slot].bytes < n * L.affine_bytes) OZK_CUDA(cudaStreamSynchronize(side));   // a regrow frees the buffer: nothing may be in flight on it
    OZK_TRY(ctx->msm[aff_slot].reserve(n * L.affine_bytes, st));
    cudaEvent_t fork = ctx->copy_ev[kCopyChunks], join = ctx->copy_ev[kCopyChunks + 1 + group_bit];
    if (!(ctx->conv_forked)) {
        OZK_CUDA(cudaEventRecord(fork, st));
        OZK_CUDA(cudaStreamWaitEvent(side, fork, 0));
    }
    if (L.convert(side, src.wire, ctx->msm[aff_slot].p, n, (uint32_t*)ctx->msm[B_MISC].p, ctx->sm_count)) { set_error("msm: convert launch failed"); return OZK_ERR_CUDA; }
    OZK_CUDA(cudaEventRecord(join, side));
    ctx->launches += 1;
    ctx->conv_forked |= 1 << group_bit;
    return OZK_OK;
}

// bucket phase, part 1: convert `n` bases and add them into the buckets of their digits.  resume == false starts from empty
// buckets; resume == true adds to what earlier slices of the same MSM left there (same shape c / nwin / nb).
static int msm_accumulate_phase(ozk_ctx* ctx, const MsmLaunch& L, const BaseSrc& src, size_t n, const MsmShape& sh, int aff_slot, int bkt_slot,
                                bool resume) {
    cudaStream_t st = ctx->stream;
    const size_t nbt = (size_t)sh.nwin * sh.nb;
    uint32_t* misc = (uint32_t*)ctx->msm[B_MISC].p;
    OZK_TRY(ctx->msm[bkt_slot].reserve(nbt * L.xyzz_bytes, st));
    OZK_TRY(ctx->msm[B_OVFPART].reserve((size_t)sh.ovf_task_cap * L.xyzz_bytes, st));
    OZK_CUDA(cudaEventRecord(ctx->evs[1], st));
    const void* aff = src.affine;
    if (!aff) {
        const int group_bit = aff_slot == B_AFF2 ? 1 : 0;
        if (ctx->conv_forked & (1 << group_bit)) {
            // converted on the side stream while the sort ran (msm_convert_fork): join
            OZK_CUDA(cudaStreamWaitEvent(st, ctx->copy_ev[kCopyChunks + 1 + group_bit], 0));
            ctx->conv_forked &= ~(1 << group_bit);
        } else {
            OZK_TRY(ctx->msm[aff_slot].reserve(n * L.affine_bytes, st));
            if (L.convert(st, src.wire, ctx->msm[aff_slot].p, n, misc, ctx->sm_count)) { set_error("msm: convert launch failed"); return OZK_ERR_CUDA; }
            ctx->launches += 1;
        }
        aff = ctx->msm[aff_slot].p;
    }
    OZK_CUDA(cudaEventRecord(ctx->evs[2], st));
    if (sh.ba && &L == &kMsmG1) {
        // batch-affine pre-reduction: rounds of pairwise affine additions over all bucket runs, then the XYZZ walk over the short runs
        const MsmBaLaunch& B = kMsmBaG1;
        const size_t pairs0 = (size_t)sh.nwin * sh.wstride / 2;
        OZK_TRY(ctx->msm[B_PAIR0].reserve(pairs0 * B.affine_bytes, st));
        if (sh.ba > 1) OZK_TRY(ctx->msm[B_PAIR1].reserve(pairs0 / 2 * B.affine_bytes, st));
        auto batch_for = [&](size_t total) {
            int M = 32;
            if (const char* e = getenv("OZK_MSM_BA_M")) M = std::max(8, std::min(kBaMaxM, atoi(e)));
            else while (M < kBaMaxM && total / ((size_t)M * 2) >= (size_t)ctx->sm_count * 512 * 2) M *= 2;     // keep >= 2 waves of threads
            return M;
        };
        if (B.round0(st, aff, (const uint32_t*)ctx->msm[B_SORTED].p, pairs0, ctx->msm[B_PAIR0].p, batch_for(pairs0))) { set_error("msm: pair round launch failed"); return OZK_ERR_CUDA; }
        const void* cur = ctx->msm[B_PAIR0].p;
        size_t total = pairs0;
        for (uint32_t r = 1; r < sh.ba; r++) {
            total /= 2;
            void* dst = (r & 1) ? ctx->msm[B_PAIR1].p : ctx->msm[B_PAIR0].p;
            if (B.round(st, cur, total, dst, batch_for(total))) { set_error("msm: pair round launch failed"); return OZK_ERR_CUDA; }
            cur = dst;
        }
        ctx->launches += sh.ba;
        if (B.accumulate_pre(st, cur, (const uint32_t*)ctx->msm[B_START].p, (const uint32_t*)ctx->msm[B_COUNT].p, (const OvfTask*)ctx->msm[B_OVFTASK].p,
                             misc + 1, (const uint32_t*)ctx->msm[B_ORDER].p, (uint32_t)nbt, sh.log_nb, sh.wstride, sh.seg, sh.ba, resume ? 1u : 0u,
                             sh.ovf_task_cap, ctx->msm[bkt_slot].p, ctx->msm[B_OVFPART].p)) { set_error("msm: accumulate launch failed"); return OZK_ERR_CUDA; }
    } else if (L.accumulate(st, aff, (const uint32_t*)ctx->msm[B_SORTED].p, (const uint32_t*)ctx->msm[B_START].p,
                     (const uint32_t*)ctx->msm[B_COUNT].p, (const OvfTask*)ctx->msm[B_OVFTASK].p, misc + 1, (const uint32_t*)ctx->msm[B_ORDER].p,
                     (uint32_t)nbt, sh.log_nb, sh.wstride, sh.seg, resume ? 1u : 0u,
                     sh.ovf_task_cap, ctx->msm[bkt_slot].p, ctx->msm[B_OVFPART].p)) { set_error("msm: accumulate launch failed"); return OZK_ERR_CUDA; }
    OZK_CUDA(cudaEventRecord(ctx->evs[3], st));
    if (L.merge(st, (const OvfBucket*)ctx->msm[B_OVFBUCKET].p, misc + 2, std::min<uint32_t>(sh.ovf_bucket_cap, (uint32_t)nbt),
                ctx->msm[B_OVFPART].p, ctx->msm[bkt_slot].p)) { set_error("msm: merge launch failed"); return OZK_ERR_CUDA; }
    OZK_CUDA(cudaEventRecord(ctx->evs[4], st));
    ctx->launches += 3;
    OZK_CUDA(cudaGetLastError());
    return OZK_OK;
}

// bucket phase, part 2: buckets -> window sums -> result in d_out (jac_bytes, canonical wire format)
static int msm_reduce_phase(ozk_ctx* ctx, const MsmLaunch& L, const MsmShape& sh, int bkt_slot, void* d_out) {
    cudaStream_t st = ctx->stream;

    // hierarchical reduction: one launch per level (msm_reduce_level).  Level l turns A_l (m[l] per window) into run = A_{l+1}
    // and acc_l (m[l+1] per window each) and carries acc_0 .. acc_{l-1} one summation step further.  Levels with enough inputs
    // to fill the machine use groups of 8 (a serial chain of 15 additions per thread, fewest bytes moved); below that the chain
    // is pure latency and groups of 2 -- one addition per thread and level, and fewer additions in total -- finish sooner.
    uint32_t m[kMaxReduceLevels + 1], logs[kMaxReduceLevels], nlev = 0;
    m[0] = sh.nb;
    while (m[nlev] > 1) {
        if (nlev >= (uint32_t)kMaxReduceLevels) { set_error("msm: too many reduction levels"); return OZK_ERR_ARG; }
        logs[nlev] = ((size_t)m[nlev] * sh.nwin >= kWsumBigMin) ? kWsumLogSBig : kWsumLogSSmall;
        const uint32_t S = 1u << logs[nlev];
        m[nlev + 1] = (m[nlev] + S - 1) / S;
        nlev++;
    }
    size_t per_win = 4;
    for (uint32_t l = 0; l < nlev; l++) per_win += (size_t)(l + 2) * m[l + 1];
    OZK_TRY(ctx->msm[B_SCRATCH].reserve((per_win * sh.nwin + 512) * L.xyzz_bytes, st));
    char* sp = (char*)ctx->msm[B_SCRATCH].p;
    auto take = [&](size_t elems_per_win) {
        void* p = sp;
        sp += elems_per_win * sh.nwin * L.xyzz_bytes;
        return p;
    };
    FinalArgs fa;
    memset(&fa, 0, sizeof fa);
    const void* level_in = ctx->msm[bkt_slot].p;
    const void* partial[kMaxReduceLevels] = {};          // partial[k]: current (partially summed) acc array of level k
    for (uint32_t l = 0; l < nlev; l++) {
        ReduceArgs ra;
        memset(&ra, 0, sizeof ra);
        ra.m_in = m[l];
        ra.nwin = sh.nwin;
        ra.s = 1u << logs[l];
        void* run = take(m[l + 1]);
        void* acc = take(m[l + 1]);
        ra.job[0] = {(const uint4*)level_in, (uint4*)run, (uint4*)acc};
        ra.njobs = 1;
        for (uint32_t k = 0; k < l; k++) {
            void* nxt = take(m[l + 1]);
            ra.job[ra.njobs++] = {(const uint4*)partial[k], (uint4*)nxt, nullptr};
            partial[k] = nxt;
        }
        partial[l] = acc;
        if (L.reduce_level(st, ra)) { set_error("msm: reduce launch failed"); return OZK_ERR_CUDA; }
        ctx->launches += 1;
        level_in = run;
    }
    for (uint32_t l = 0; l < nlev; l++) {
        fa.sum_acc[l] = (const uint4*)partial[l];
        fa.log_s[l] = (uint8_t)logs[l];
    }
    fa.total = (const uint4*)level_in;     // one element per window: the plain sum of all buckets (nb == 1: the bucket itself)
    fa.nlevels = nlev;
    fa.nwin = sh.nwin;
    fa.c = sh.c;
    void* window_vals = take(1);
    // completion counter of msm_tail: the last word of B_MISC, zeroed when the buffer is first sized and left zero by the kernel
    if (L.final(st, fa, window_vals, d_out, (uint32_t*)ctx->msm[B_MISC].p + (kMiscWords - 1))) { set_error("msm: final launch failed"); return OZK_ERR_CUDA; }
    ctx->launches += 1;
    OZK_CUDA(cudaEventRecord(ctx->evs[5], st));
    OZK_CUDA(cudaGetLastError());
    return OZK_OK;
}

static int msm_finish(ozk_ctx* ctx, const void* d_res, size_t bytes, uint8_t* out) {
    // result + flag come back in one small pinned block
    uint8_t* pin = (uint8_t*)ctx->pinned;
    OZK_CUDA(cudaMemcpyAsync(pin, d_res, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    OZK_CUDA(cudaMemcpyAsync(pin + 1024, ctx->msm[B_MISC].p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    OZK_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint32_t* misc = (const uint32_t*)(pin + 1024);
    ctx->msm_stats[3] = misc[1];
    ctx->msm_stats[4] = misc[2];
    for (int k = 0; k < 5; k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->evs[k], ctx->evs[k + 1]) != cudaSuccess) { ms = -1.f; cudaGetLastError(); }
        ctx->msm_stats[5 + k] = ms;
    }
    if (misc[0] & 1u) { set_error("msm: a base coordinate is not reduced mod p"); return OZK_ERR_DOMAIN; }
    if (misc[0] & 2u) { set_error("msm: a scalar is not reduced mod r"); return OZK_ERR_DOMAIN; }
    memcpy(out, pin, bytes);
    return OZK_OK;
}

// n == 0: the empty sum
static void write_inf(uint8_t* out, size_t coord_bytes) {
    memset(out, 0, 3 * coord_bytes);
    out[coord_bytes] = 1;      // (0, 1, 0)
}

// Device-resident scalars.  Large inputs are still cut into slices of kSliceLen pairs that share one set of buckets (and
// one reduction): the sort scratch (nwin x len x 4 B) stays bounded.
static constexpr size_t kDevSliceLen = (size_t)1 << 26;

static int msm_run(ozk_ctx* ctx, const void* d_scalars, BaseSrc s1, BaseSrc s2, size_t n, uint8_t* out) {
    OZK_ARG(n > 0 && n < ((size_t)1 << 31), "msm: between 1 and 2^31 - 1 points per call");
    OZK_TRY(ctx->msm[B_MISC].reserve(kMiscWords * 4, ctx->stream));
    OZK_CUDA(cudaMemsetAsync(ctx->msm[B_MISC].p, 0, 4, ctx->stream));
    OZK_CUDA(cudaEventRecord(ctx->evs[0], ctx->stream));
    OZK_TRY(ctx->io_out.reserve(512, ctx->stream));
    char* d_res = (char*)ctx->io_out.p;
    MsmShape sh = msm_shape(n, std::min(n, kDevSliceLen));
    ctx->msm_stats[0] = sh.c;
    ctx->msm_stats[1] = sh.nwin;
    ctx->msm_stats[2] = sh.nb;
    for (size_t lo = 0; lo < n; lo += kDevSliceLen) {
        const size_t len = std::min(kDevSliceLen, n - lo);
        sh = msm_shape(n, len);
        ctx->conv_forked = 0;
        if (s1.wire) OZK_TRY(msm_convert_fork(ctx, kMsmG1, BaseSrc{(const char*)s1.wire + lo * kMsmG1.jac_bytes, nullptr}, len, B_AFF1, 0));
        if (s2.wire) OZK_TRY(msm_convert_fork(ctx, kMsmG2, BaseSrc{(const char*)s2.wire + lo * kMsmG2.jac_bytes, nullptr}, len, B_AFF2, 1));
        OZK_TRY(msm_sort(ctx, (const char*)d_scalars + lo * 32, len, sh));
        if (s1.any()) {
            BaseSrc b = {s1.wire ? (const char*)s1.wire + lo * kMsmG1.jac_bytes : nullptr, s1.affine ? (const char*)s1.affine + lo * kMsmG1.affine_bytes : nullptr};
            OZK_TRY(msm_accumulate_phase(ctx, kMsmG1, b, len, sh, B_AFF1, B_BUCKETS, lo != 0));
        }
        if (s2.any()) {
            BaseSrc b = {s2.wire ? (const char*)s2.wire + lo * kMsmG2.jac_bytes : nullptr, s2.affine ? (const char*)s2.affine + lo * kMsmG2.affine_bytes : nullptr};
            OZK_TRY(msm_accumulate_phase(ctx, kMsmG2, b, len, sh, B_AFF2, B_BUCKETS2, lo != 0));
        }
    }
    size_t bytes = 0;
    if (s1.any()) {
        OZK_TRY(msm_reduce_phase(ctx, kMsmG1, sh, B_BUCKETS, d_res));
        bytes += 96;
    }
    if (s2.any()) {
        OZK_TRY(msm_reduce_phase(ctx, kMsmG2, sh, B_BUCKETS2, d_res + bytes));
        bytes += 192;
    }
    return msm_finish(ctx, d_res, bytes, out);
}

// Host-pointer entry, as a stream of slices.  begin fixes the window shape from the TOTAL number of pairs; every feed uploads one
// slice on the copy stream (or through the bounce-buffer stager when the memory is pageable, which is what the JNI shims pass) and
// enqueues its sort / normalise / accumulate behind the copy, so slice k+1 crosses PCIe while slice k is computed; all slices add
// into the same buckets (an MSM is a sum over points, so a bucket may receive its points in any grouping), so the bucket
// reduction and the Horner tail run once, in end.  This is the partition the Java does with its 2^23-element chunks
// (VariableBaseMSM.java:211-265), except that here it only exists to overlap the copies -- and to let a JNI caller hold its
// array pins for one slice's host copy at a time instead of the whole call (csrc/jni/jni_shim.cc).
// Host base arrays may be null when the group's bases are a persistent key (affine device memory).
struct MsmStream {
    bool active = false;
    size_t n_total = 0, fed = 0;
    BaseSrc s1, s2;                 // device sources of the two groups (wire staging or persistent affine)
    bool host1 = false, host2 = false;
    bool first = true;
    bool time_first = false;        // bracket the uploads of the first slice with the timing events ev0 / ev1 (msm_run_host)
    int slices = 0;
};
static MsmStream& stream_of(ozk_ctx* ctx) {
    static_assert(sizeof(MsmStream) <= sizeof(ctx->msm_stream), "msm_stream storage in common.h is too small");
    return *reinterpret_cast<MsmStream*>(ctx->msm_stream);
}

// all per-slice scratch for slices of up to `len` pairs, reserved once: a regrow in the middle of the pipeline is a stream
// synchronisation plus cudaFree / cudaMalloc exactly where the copies should overlap the compute
static int msm_reserve_slices(ozk_ctx* ctx, size_t n_total, size_t len, bool g1, bool g2, bool conv1, bool conv2) {
    cudaStream_t st = ctx->stream;
    const MsmShape sh = msm_shape(n_total, len);
    OZK_TRY(ctx->msm[B_SORTED].reserve((size_t)sh.nwin * sh.wstride * 4, st));
    OZK_TRY(ctx->msm[B_OVFTASK].reserve((size_t)sh.ovf_task_cap * sizeof(OvfTask), st));
    OZK_TRY(ctx->msm[B_OVFBUCKET].reserve((size_t)sh.ovf_bucket_cap * sizeof(OvfBucket), st));
    const size_t part = (size_t)sh.ovf_task_cap * (g2 ? kMsmG2.xyzz_bytes : kMsmG1.xyzz_bytes);
    OZK_TRY(ctx->msm[B_OVFPART].reserve(part, st));
    if (g1 && conv1) OZK_TRY(ctx->msm[B_AFF1].reserve(len * kMsmG1.affine_bytes, st));
    if (g2 && conv2) OZK_TRY(ctx->msm[B_AFF2].reserve(len * kMsmG2.affine_bytes, st));
    return OZK_OK;
}

static int msm_stream_begin(ozk_ctx* ctx, bool g1, bool g2, size_t n_total, size_t max_slice, const void* aff1, const void* aff2) {
    OZK_ARG(n_total > 0 && n_total < ((size_t)1 << 31), "msm: between 1 and 2^31 - 1 points per call");
    OZK_ARG(g1 || g2, "msm: no group selected");
    MsmStream& ms = stream_of(ctx);
    ms = MsmStream();
    cudaStream_t st = ctx->stream;
    ms.host1 = g1 && !aff1;
    ms.host2 = g2 && !aff2;
    OZK_TRY(ctx->io_a.reserve(n_total * 32, st));
    if (ms.host1) OZK_TRY(ctx->io_b.reserve(n_total * 96, st));
    if (ms.host2) OZK_TRY(ctx->io_c.reserve(n_total * 192, st));
    OZK_TRY(ctx->msm[B_MISC].reserve(kMiscWords * 4, st));
    OZK_TRY(ctx->io_out.reserve(512, st));
    if (max_slice) OZK_TRY(msm_reserve_slices(ctx, n_total, std::min(max_slice, n_total), g1, g2, ms.host1, ms.host2));
    OZK_CUDA(cudaMemsetAsync(ctx->msm[B_MISC].p, 0, 4, st));
    OZK_CUDA(cudaEventRecord(ctx->evs[0], st));
    // the copy stream must not overtake earlier work on the main stream that still uses the staging buffers
    OZK_CUDA(cudaEventRecord(ctx->copy_ev[2 * kCopyChunks], st));
    OZK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[2 * kCopyChunks], 0));
    if (g1) ms.s1 = ms.host1 ? BaseSrc{ctx->io_b.p, nullptr} : BaseSrc{nullptr, aff1};
    if (g2) ms.s2 = ms.host2 ? BaseSrc{ctx->io_c.p, nullptr} : BaseSrc{nullptr, aff2};
    ms.n_total = n_total;
    ms.active = true;
    const MsmShape sh = msm_shape(n_total, n_total);
    ctx->msm_stats[0] = sh.c;
    ctx->msm_stats[1] = sh.nwin;
    ctx->msm_stats[2] = sh.nb;
    return OZK_OK;
}

// release_host: return only when the host arrays are no longer read (always true for pageable memory, which goes through
// the stager; for pinned memory it costs a wait for the slice's DMA)
static int msm_stream_feed(ozk_ctx* ctx, const uint8_t* scalars, const uint8_t* b1, const uint8_t* b2, size_t len, bool release_host) {
    MsmStream& ms = stream_of(ctx);
    OZK_ARG(ms.active, "ozk_msm_feed: no MSM in progress on this context (ozk_msm_begin first)");
    OZK_ARG(len <= ms.n_total - ms.fed, "ozk_msm_feed: more pairs than announced to ozk_msm_begin");
    OZK_ARG(len == 0 || (scalars && (!ms.host1 || b1) && (!ms.host2 || b2)), "ozk_msm_feed: null pointer");
    if (len == 0) return OZK_OK;
    cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
    const size_t lo = ms.fed;
    cudaEvent_t after = ctx->copy_ev[2 * kCopyChunks];
    bool any_async = false;
    auto upload = [&](void* dst, const uint8_t* src, size_t nbytes) -> int {
        if (nbytes >= ((size_t)8 << 20) && host_pointer_is_pageable(src)) return staged_h2d(ctx, dst, src, nbytes, after, st);
        OZK_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, cs));
        any_async = any_async || !host_pointer_is_pageable(src);
        return OZK_OK;
    };
    const bool timed = ms.time_first && ms.slices == 0;
    if (timed) OZK_CUDA(cudaEventRecord(ctx->ev0, cs));
    OZK_TRY(upload((char*)ctx->io_a.p + lo * 32, scalars, len * 32));
    if (ms.host1) OZK_TRY(upload((char*)ctx->io_b.p + lo * 96, b1, len * 96));
    if (ms.host2) OZK_TRY(upload((char*)ctx->io_c.p + lo * 192, b2, len * 192));
    if (timed) OZK_CUDA(cudaEventRecord(ctx->ev1, cs));      // right behind the copies: host time spent below must not count
    cudaEvent_t ev = ctx->copy_ev[ms.slices % kCopyChunks];
    OZK_CUDA(cudaEventRecord(ev, cs));
    OZK_CUDA(cudaStreamWaitEvent(st, ev, 0));
    if (release_host && any_async) OZK_CUDA(cudaEventSynchronize(ev));
    const MsmShape sh = msm_shape(ms.n_total, len);
    ctx->conv_forked = 0;
    if (ms.s1.wire) OZK_TRY(msm_convert_fork(ctx, kMsmG1, BaseSrc{(const char*)ms.s1.wire + lo * kMsmG1.jac_bytes, nullptr}, len, B_AFF1, 0));
    if (ms.s2.wire) OZK_TRY(msm_convert_fork(ctx, kMsmG2, BaseSrc{(const char*)ms.s2.wire + lo * kMsmG2.jac_bytes, nullptr}, len, B_AFF2, 1));
    OZK_TRY(msm_sort(ctx, (char*)ctx->io_a.p + lo * 32, len, sh));
    if (ms.s1.any()) {
        BaseSrc b = {ms.s1.wire ? (const char*)ms.s1.wire + lo * kMsmG1.jac_bytes : nullptr,
                     ms.s1.affine ? (const char*)ms.s1.affine + lo * kMsmG1.affine_bytes : nullptr};
        OZK_TRY(msm_accumulate_phase(ctx, kMsmG1, b, len, sh, B_AFF1, B_BUCKETS, !ms.first));
    }
    if (ms.s2.any()) {
        BaseSrc b = {ms.s2.wire ? (const char*)ms.s2.wire + lo * kMsmG2.jac_bytes : nullptr,
                     ms.s2.affine ? (const char*)ms.s2.affine + lo * kMsmG2.affine_bytes : nullptr};
        OZK_TRY(msm_accumulate_phase(ctx, kMsmG2, b, len, sh, B_AFF2, B_BUCKETS2, !ms.first));
    }
    ms.first = false;
    ms.fed += len;
    ms.slices++;
    return OZK_OK;
}

static int msm_stream_end(ozk_ctx* ctx, uint8_t* out) {
    MsmStream& ms = stream_of(ctx);
    OZK_ARG(ms.active, "ozk_msm_end: no MSM in progress on this context");
    ms.active = false;
    OZK_ARG(ms.fed == ms.n_total, "ozk_msm_end: fewer pairs were fed than announced to ozk_msm_begin");
    const MsmShape sh = msm_shape(ms.n_total, ms.n_total);
    char* d_res = (char*)ctx->io_out.p;
    size_t bytes = 0;
    if (ms.s1.any()) {
        OZK_TRY(msm_reduce_phase(ctx, kMsmG1, sh, B_BUCKETS, d_res));
        bytes += 96;
    }
    if (ms.s2.any()) {
        OZK_TRY(msm_reduce_phase(ctx, kMsmG2, sh, B_BUCKETS2, d_res + bytes));
        bytes += 192;
    }
    return msm_finish(ctx, d_res, bytes, out);
}

// Slice schedule of a whole-array call.  The compute of a slice starts when its copy has landed and the copies run back to
// back, so a step ends at max(first copy + all compute + per-slice overheads, all copies + compute of the LAST slice).  Every
// slice costs ~0.5 ms of fixed work (each accumulate launch reloads and stores all buckets and starts every run cold).
//  * pinned memory on one GPU: copy 2.3 ns per 128-byte pair, compute 2.8 ns -- both ends bind at once.  Eight slices growing
//    x1.3 measured best (profiles/r2_e2e_slice_plans.jsonl: 53.7 ms at 2^24; equal slices 56.3, x1.2 54.6, x1.5 57.8, six or
//    twelve slices 54.4, a schedule with small first AND last slices 55.0 -- it is then bound by the per-slice overheads).
//  * pageable memory (what the JNI shims pass; it crosses through the bounce-buffer stager at ~25 GB/s) is bound by the copies,
//    and the step ends one last-slice compute after the last byte: there the schedule with a small last slice wins
//    (2^23 pairs through the legacy JNI symbol: 37.7 ms against 43-50 ms).
// OZK_HOST_SLICES / OZK_HOST_SLICE_GROWTH / OZK_HOST_PLAN override (tools/e2e_plan_sweep.py).
static constexpr int kMaxSlices = 16;
static int msm_plan_slices(size_t n, size_t* bounds, int cap, bool copy_bound) {
    static const double plan8[8] = {0.04, 0.07, 0.11, 0.15, 0.19, 0.20, 0.16, 0.08};
    static const double plan4[4] = {0.10, 0.27, 0.38, 0.25};
    static const double plan2[2] = {0.35, 0.65};
    int nslices = 1;
    if (n >= ((size_t)1 << 22)) nslices = 8;
    else if (n >= ((size_t)1 << 20)) nslices = 4;
    else if (n >= ((size_t)1 << 18)) nslices = 2;
    double growth = 0;
    if (const char* e = getenv("OZK_HOST_SLICES")) {
        nslices = std::max(1, std::min(kMaxSlices, atoi(e)));
        growth = 1.3;
    }
    if (const char* e = getenv("OZK_HOST_SLICE_GROWTH")) growth = std::max(1.0, atof(e));
    nslices = std::min(nslices, cap);
    double frac[kMaxSlices];
    if (const char* e = getenv("OZK_HOST_PLAN")) {
        // explicit fractions "0.05,0.1,..." (normalised to 1): the tuning knob tools/e2e_plan_sweep.py turns
        int k = 0;
        double total = 0;
        for (const char* q = e; *q && k < std::min(cap, kMaxSlices);) {
            char* end = nullptr;
            const double v = strtod(q, &end);
            if (end == q) break;
            if (v > 0) { frac[k++] = v; total += v; }
            q = (*end == ',') ? end + 1 : end;
        }
        if (k > 0 && n >= ((size_t)1 << 18)) {
            double acc = 0;
            bounds[0] = 0;
            for (int i = 0; i < k; i++) {
                acc += frac[i] / total;
                size_t b = ((size_t)((double)n * acc) + 255) & ~(size_t)255;
                bounds[i + 1] = (i + 1 == k) ? n : std::min(b, n);
                if (bounds[i + 1] < bounds[i]) bounds[i + 1] = bounds[i];
            }
            return k;
        }
    }
    const double* tab = !copy_bound ? nullptr : nslices == 8 ? plan8 : nslices == 4 ? plan4 : nslices == 2 ? plan2 : nullptr;
    if (growth == 0 && tab) {
        for (int k = 0; k < nslices; k++) frac[k] = tab[k];
    } else {
        if (growth == 0) growth = 1.3;
        double total = 0, w = 1;
        for (int k = 0; k < nslices; k++, w *= growth) total += w;
        w = 1;
        for (int k = 0; k < nslices; k++, w *= growth) frac[k] = w / total;
    }
    double acc = 0;
    bounds[0] = 0;
    for (int k = 0; k < nslices; k++) {
        acc += frac[k];
        size_t b = (size_t)((double)n * acc);
        b = (b + 255) & ~(size_t)255;
        bounds[k + 1] = (k + 1 == nslices) ? n : std::min(b, n);
        if (bounds[k + 1] < bounds[k]) bounds[k + 1] = bounds[k];
    }
    return nslices;
}

static int msm_run_host(ozk_ctx* ctx, const uint8_t* scalars, const uint8_t* b1, const uint8_t* b2, BaseSrc s1, BaseSrc s2, size_t n, uint8_t* out) {
    OZK_ARG(n > 0 && n < ((size_t)1 << 31), "msm: between 1 and 2^31 - 1 points per call");
    size_t bounds[kMaxSlices + 1];
    const bool pageable = host_pointer_is_pageable(b1 ? b1 : b2 ? b2 : scalars);
    int nslices = msm_plan_slices(n, bounds, kMaxSlices, pageable);
    size_t slice = 0;                                    // the longest slice sizes the per-slice scratch
    for (int k = 0; k < nslices; k++) slice = std::max(slice, bounds[k + 1] - bounds[k]);
    OZK_TRY(msm_stream_begin(ctx, b1 || s1.any(), b2 || s2.any(), n, slice, s1.affine, s2.affine));
    // Pinned memory: the schedule assumes the full PCIe rate.  When several GPUs share a host link (4 or 8 ranks on one box each
    // moving 2 GiB per step: ~21 GB/s per GPU instead of ~54) the copies bind and the step ends one last-slice compute after the
    // last byte, so the rate of the FIRST slice's copy is measured (two events on the copy stream, one host wait of a millisecond
    // or two that the GPU does not notice) and a slow link switches the remaining slices to the small-last-slice schedule.
    const bool adaptive = !pageable && nslices == 8 && b1 && !b2 && !getenv("OZK_HOST_SLICES") && !getenv("OZK_HOST_SLICE_GROWTH") &&
                          !getenv("OZK_HOST_PLAN") && !getenv("OZK_HOST_NO_ADAPT");
    for (int k = 0; k < nslices; k++) {
        const size_t lo = bounds[k], len = bounds[k + 1] - bounds[k];
        if (adaptive && k == 0) stream_of(ctx).time_first = true;
        int rc = msm_stream_feed(ctx, scalars + lo * 32, b1 ? b1 + lo * 96 : nullptr, b2 ? b2 + lo * 192 : nullptr, len, false);
        if (rc != OZK_OK) {
            stream_of(ctx).active = false;
            return rc;
        }
        if (adaptive && k == 0) {
            float ms = 0;
            if (cudaEventSynchronize(ctx->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess && ms > 0) {
                const double gbps = (double)len * 128 / (ms * 1e6);
                ctx->msm_stats[10] = gbps;
                double slow = 38.0;                                   // GB/s; OZK_HOST_ADAPT_GBPS moves the threshold (tests force the re-plan with it)
                if (const char* e = getenv("OZK_HOST_ADAPT_GBPS")) slow = atof(e);
                if (gbps < slow) {
                    // re-plan what is left: fractions 7, 11, 15, 19, 20, 16, 8 (% of the whole) behind the first slice
                    static const double rest[7] = {0.07, 0.11, 0.15, 0.19, 0.20, 0.16, 0.08};
                    double total = 0, acc = 0;
                    for (double f : rest) total += f;
                    const size_t left = n - bounds[1];
                    for (int q = 0; q < 7; q++) {
                        acc += rest[q] / total;
                        size_t b = bounds[1] + (((size_t)((double)left * acc) + 255) & ~(size_t)255);
                        bounds[q + 2] = (q == 6) ? n : std::min(b, n);
                        if (bounds[q + 2] - bounds[q + 1] > slice) bounds[q + 2] = bounds[q + 1] + slice;   // never beyond the reserved scratch
                    }
                    bounds[8] = n;
                    if (bounds[8] - bounds[7] > slice) {      // (cannot happen with these fractions; keep the invariant anyway)
                        stream_of(ctx).active = false;
                        set_error("msm: slice re-plan exceeded the reserved scratch");
                        return OZK_ERR_ARG;
                    }
                }
            } else {
                cudaGetLastError();
            }
        }
    }
    return msm_stream_end(ctx, out);
}

// ---- persistent bases -------------------------------------------------------------------------------------------
// A proving key's query vectors are the same for every proof (ProvingKey.java:16-47; SerialProver.prove reads
// queryA/queryB/deltaABCG1/queryH, SerialProver.java:70-106), so they can be uploaded and normalised to affine
// Montgomery form once; later MSMs then move only the scalars.
static int bases_upload(ozk_ctx* ctx, const MsmLaunch& L, int group, const uint8_t* host, const void* dev, size_t n, ozk_bases** out) {
    OZK_ARG(out != nullptr && n > 0 && n < ((size_t)1 << 31) && (host || dev), "ozk_bases_upload: bad argument");
    cudaStream_t st = ctx->stream;
    ozk_bases* kb = new ozk_bases();
    kb->group = group;
    kb->n = n;
    kb->device = ctx->device;
    if (cudaMalloc(&kb->d_affine, n * L.affine_bytes) != cudaSuccess) {
        cudaGetLastError();
        delete kb;
        set_error("ozk_bases_upload: out of device memory for %zu points", n);
        return OZK_ERR_CUDA;
    }
    auto fail = [&](int rc) {
        cudaFree(kb->d_affine);
        delete kb;
        return rc;
    };
    if (ctx->msm[B_MISC].reserve(kMiscWords * 4, st) != OZK_OK) return fail(OZK_ERR_CUDA);
    uint32_t* misc = (uint32_t*)ctx->msm[B_MISC].p;
    if (cudaMemsetAsync(misc, 0, 4, st) != cudaSuccess) return fail(OZK_ERR_CUDA);
    const size_t chunk = (size_t)1 << 22;
    if (host && ctx->io_b.reserve(std::min(n, chunk) * L.jac_bytes, st) != OZK_OK) return fail(OZK_ERR_CUDA);
    for (size_t lo = 0; lo < n; lo += chunk) {
        const size_t len = std::min(chunk, n - lo);
        const void* src = dev ? (const void*)((const char*)dev + lo * L.jac_bytes) : ctx->io_b.p;
        if (host && cudaMemcpyAsync(ctx->io_b.p, host + lo * L.jac_bytes, len * L.jac_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(OZK_ERR_CUDA);
        if (L.convert(st, src, (char*)kb->d_affine + lo * L.affine_bytes, len, misc, ctx->sm_count)) return fail(OZK_ERR_CUDA);
        ctx->launches += 1;
    }
    uint32_t* pin = (uint32_t*)ctx->pinned;
    if (cudaMemcpyAsync(pin, misc, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("ozk_bases_upload: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(OZK_ERR_CUDA);
    }
    if (pin[0] & 1u) {
        set_error("ozk_bases_upload: a base coordinate is not reduced mod p");
        return fail(OZK_ERR_DOMAIN);
    }
    *out = kb;
    return OZK_OK;
}

static int keyed_check(ozk_ctx* ctx, const ozk_bases* kb, int group, size_t first, size_t n) {
    OZK_ARG(kb != nullptr, "keyed msm: null bases handle");
    OZK_ARG(kb->group == group, "keyed msm: bases handle belongs to the other group");
    OZK_ARG(kb->device == ctx->device, "keyed msm: bases handle lives on another device");
    OZK_ARG(first <= kb->n && n <= kb->n - first, "keyed msm: [first, first + n) exceeds the uploaded bases");
    return OZK_OK;
}

}  // namespace ozk

using namespace ozk;

extern "C" {

int ozk_msm_g1_dev(ozk_ctx* ctx, const void* d_scalars, const void* d_bases, size_t n, uint8_t out[96]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || (d_scalars && d_bases)), "ozk_msm_g1_dev: null pointer");
    if (n == 0) { write_inf(out, 32); return OZK_OK; }
    return msm_run(ctx, d_scalars, BaseSrc{d_bases, nullptr}, BaseSrc{}, n, out);
}
int ozk_msm_g2_dev(ozk_ctx* ctx, const void* d_scalars, const void* d_bases, size_t n, uint8_t out[192]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || (d_scalars && d_bases)), "ozk_msm_g2_dev: null pointer");
    if (n == 0) { write_inf(out, 64); return OZK_OK; }
    return msm_run(ctx, d_scalars, BaseSrc{}, BaseSrc{d_bases, nullptr}, n, out);
}
int ozk_msm_g1g2_dev(ozk_ctx* ctx, const void* d_scalars, const void* d_bases1, const void* d_bases2, size_t n, uint8_t out[288]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || (d_scalars && d_bases1 && d_bases2)), "ozk_msm_g1g2_dev: null pointer");
    if (n == 0) {
        write_inf(out, 32);
        write_inf(out + 96, 64);
        return OZK_OK;
    }
    return msm_run(ctx, d_scalars, BaseSrc{d_bases1, nullptr}, BaseSrc{d_bases2, nullptr}, n, out);
}
int ozk_msm_g1(ozk_ctx* ctx, const uint8_t* scalars, const uint8_t* bases, size_t n, uint8_t out[96]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || (scalars && bases)), "ozk_msm_g1: null pointer");
    if (n == 0) { write_inf(out, 32); return OZK_OK; }
    return msm_run_host(ctx, scalars, bases, nullptr, BaseSrc{}, BaseSrc{}, n, out);
}
int ozk_msm_g2(ozk_ctx* ctx, const uint8_t* scalars, const uint8_t* bases, size_t n, uint8_t out[192]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || (scalars && bases)), "ozk_msm_g2: null pointer");
    if (n == 0) { write_inf(out, 64); return OZK_OK; }
    return msm_run_host(ctx, scalars, nullptr, bases, BaseSrc{}, BaseSrc{}, n, out);
}
int ozk_msm_g1g2(ozk_ctx* ctx, const uint8_t* scalars, const uint8_t* bases1, const uint8_t* bases2, size_t n, uint8_t out[288]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || (scalars && bases1 && bases2)), "ozk_msm_g1g2: null pointer");
    if (n == 0) {
        write_inf(out, 32);
        write_inf(out + 96, 64);
        return OZK_OK;
    }
    return msm_run_host(ctx, scalars, bases1, bases2, BaseSrc{}, BaseSrc{}, n, out);
}

int ozk_msm_begin(ozk_ctx* ctx, int groups, size_t n_total, size_t max_slice, const ozk_bases* key1, const ozk_bases* key2, size_t first) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(groups >= 1 && groups <= 3, "ozk_msm_begin: groups must be 1 (G1), 2 (G2) or 3 (both)");
    if (key1) OZK_TRY(keyed_check(ctx, key1, 1, first, n_total));
    if (key2) OZK_TRY(keyed_check(ctx, key2, 2, first, n_total));
    OZK_ARG((!key1 || (groups & 1)) && (!key2 || (groups & 2)), "ozk_msm_begin: a key was given for a group that is not selected");
    return msm_stream_begin(ctx, (groups & 1) != 0, (groups & 2) != 0, n_total, max_slice,
                            key1 ? (const char*)key1->d_affine + first * kMsmG1.affine_bytes : nullptr,
                            key2 ? (const char*)key2->d_affine + first * kMsmG2.affine_bytes : nullptr);
}
int ozk_msm_feed(ozk_ctx* ctx, const uint8_t* scalars, const uint8_t* bases1, const uint8_t* bases2, size_t len) {
    OZK_TRY(ctx_enter(ctx));
    int rc = msm_stream_feed(ctx, scalars, bases1, bases2, len, true);
    if (rc != OZK_OK) stream_of(ctx).active = false;
    return rc;
}
int ozk_msm_end(ozk_ctx* ctx, uint8_t* out) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out != nullptr, "ozk_msm_end: null pointer");
    return msm_stream_end(ctx, out);
}
int ozk_msm_plan_slices(size_t n, size_t* bounds, int cap) {
    if (!bounds || cap < 1) return 0;
    if (n == 0) { bounds[0] = 0; return 0; }
    return msm_plan_slices(n, bounds, std::min(cap, kCopyChunks), true);     // callers that feed slices themselves hold pageable (JVM heap) arrays
}

int ozk_bases_upload_g1(ozk_ctx* ctx, const uint8_t* bases, size_t n, ozk_bases** out) {
    OZK_TRY(ctx_enter(ctx));
    return bases_upload(ctx, kMsmG1, 1, bases, nullptr, n, out);
}
int ozk_bases_upload_g2(ozk_ctx* ctx, const uint8_t* bases, size_t n, ozk_bases** out) {
    OZK_TRY(ctx_enter(ctx));
    return bases_upload(ctx, kMsmG2, 2, bases, nullptr, n, out);
}
int ozk_bases_upload_g1_dev(ozk_ctx* ctx, const void* d_bases, size_t n, ozk_bases** out) {
    OZK_TRY(ctx_enter(ctx));
    return bases_upload(ctx, kMsmG1, 1, nullptr, d_bases, n, out);
}
int ozk_bases_upload_g2_dev(ozk_ctx* ctx, const void* d_bases, size_t n, ozk_bases** out) {
    OZK_TRY(ctx_enter(ctx));
    return bases_upload(ctx, kMsmG2, 2, nullptr, d_bases, n, out);
}
size_t ozk_bases_len(const ozk_bases* kb) { return kb ? kb->n : 0; }
void ozk_bases_free(ozk_ctx* ctx, ozk_bases* kb) {
    if (!kb) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(kb->d_affine);
    delete kb;
}

int ozk_msm_g1_keyed(ozk_ctx* ctx, const uint8_t* scalars, const ozk_bases* key, size_t first, size_t n, uint8_t out[96]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || scalars), "ozk_msm_g1_keyed: null pointer");
    OZK_TRY(keyed_check(ctx, key, 1, first, n));
    if (n == 0) { write_inf(out, 32); return OZK_OK; }
    return msm_run_host(ctx, scalars, nullptr, nullptr, BaseSrc{nullptr, (const char*)key->d_affine + first * kMsmG1.affine_bytes}, BaseSrc{}, n, out);
}
int ozk_msm_g1_keyed_dev(ozk_ctx* ctx, const void* d_scalars, const ozk_bases* key, size_t first, size_t n, uint8_t out[96]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || d_scalars), "ozk_msm_g1_keyed_dev: null pointer");
    OZK_TRY(keyed_check(ctx, key, 1, first, n));
    if (n == 0) { write_inf(out, 32); return OZK_OK; }
    return msm_run(ctx, d_scalars, BaseSrc{nullptr, (const char*)key->d_affine + first * kMsmG1.affine_bytes}, BaseSrc{}, n, out);
}
int ozk_msm_g2_keyed(ozk_ctx* ctx, const uint8_t* scalars, const ozk_bases* key, size_t first, size_t n, uint8_t out[192]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || scalars), "ozk_msm_g2_keyed: null pointer");
    OZK_TRY(keyed_check(ctx, key, 2, first, n));
    if (n == 0) { write_inf(out, 64); return OZK_OK; }
    return msm_run_host(ctx, scalars, nullptr, nullptr, BaseSrc{}, BaseSrc{nullptr, (const char*)key->d_affine + first * kMsmG2.affine_bytes}, n, out);
}
int ozk_msm_g2_keyed_dev(ozk_ctx* ctx, const void* d_scalars, const ozk_bases* key, size_t first, size_t n, uint8_t out[192]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || d_scalars), "ozk_msm_g2_keyed_dev: null pointer");
    OZK_TRY(keyed_check(ctx, key, 2, first, n));
    if (n == 0) { write_inf(out, 64); return OZK_OK; }
    return msm_run(ctx, d_scalars, BaseSrc{}, BaseSrc{nullptr, (const char*)key->d_affine + first * kMsmG2.affine_bytes}, n, out);
}
int ozk_msm_g1g2_keyed(ozk_ctx* ctx, const uint8_t* scalars, const ozk_bases* key1, const ozk_bases* key2, size_t first, size_t n, uint8_t out[288]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || scalars), "ozk_msm_g1g2_keyed: null pointer");
    OZK_TRY(keyed_check(ctx, key1, 1, first, n));
    OZK_TRY(keyed_check(ctx, key2, 2, first, n));
    if (n == 0) {
        write_inf(out, 32);
        write_inf(out + 96, 64);
        return OZK_OK;
    }
    return msm_run_host(ctx, scalars, nullptr, nullptr, BaseSrc{nullptr, (const char*)key1->d_affine + first * kMsmG1.affine_bytes},
                        BaseSrc{nullptr, (const char*)key2->d_affine + first * kMsmG2.affine_bytes}, n, out);
}
int ozk_msm_g1g2_keyed_dev(ozk_ctx* ctx, const void* d_scalars, const ozk_bases* key1, const ozk_bases* key2, size_t first, size_t n, uint8_t out[288]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(out && (n == 0 || d_scalars), "ozk_msm_g1g2_keyed_dev: null pointer");
    OZK_TRY(keyed_check(ctx, key1, 1, first, n));
    OZK_TRY(keyed_check(ctx, key2, 2, first, n));
    if (n == 0) {
        write_inf(out, 32);
        write_inf(out + 96, 64);
        return OZK_OK;
    }
    return msm_run(ctx, d_scalars, BaseSrc{nullptr, (const char*)key1->d_affine + first * kMsmG1.affine_bytes},
                   BaseSrc{nullptr, (const char*)key2->d_affine + first * kMsmG2.affine_bytes}, n, out);
}

// out = sum of k points in the wire format (device memory): the reduce(add) of the partial sums of a sharded MSM
// (VariableBaseMSM.distributedMSM, VariableBaseMSM.java:777-783) as one tiny launch instead of a k-point MSM.
static int sum_points(ozk_ctx* ctx, const MsmLaunch& L, const void* d_points, size_t k, uint8_t* out, size_t coord_bytes) {
    OZK_ARG(out && (k == 0 || d_points) && k <= 4096, "ozk_sum: null pointer or more than 4096 points");
    if (k == 0) { write_inf(out, coord_bytes); return OZK_OK; }
    OZK_TRY(ctx->io_out.reserve(512, ctx->stream));
    if (L.sum_wire(ctx->stream, d_points, (uint32_t)k, ctx->io_out.p)) { set_error("sum: launch failed"); return OZK_ERR_CUDA; }
    ctx->launches += 1;
    uint8_t* pin = (uint8_t*)ctx->pinned;
    OZK_CUDA(cudaMemcpyAsync(pin, ctx->io_out.p, L.jac_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    OZK_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(out, pin, L.jac_bytes);
    return OZK_OK;
}
int ozk_sum_g1_dev(ozk_ctx* ctx, const void* d_points, size_t k, uint8_t out[96]) {
    OZK_TRY(ctx_enter(ctx));
    return sum_points(ctx, kMsmG1, d_points, k, out, 32);
}
int ozk_sum_g2_dev(ozk_ctx* ctx, const void* d_points, size_t k, uint8_t out[192]) {
    OZK_TRY(ctx_enter(ctx));
    return sum_points(ctx, kMsmG2, d_points, k, out, 64);
}

int ozk_msm_last_stats(ozk_ctx* ctx, double* out, int cap) {
    if (!ctx || !out) return 0;
    int k = cap < 11 ? cap : 11;       // [10]: GB/s of the first slice's upload in the last host-pointer call that measured it
    for (int i = 0; i < k; i++) out[i] = ctx->msm_stats[i];
    return k;
}

}  // extern "C"
