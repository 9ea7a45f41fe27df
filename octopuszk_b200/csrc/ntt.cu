// Radix-2 NTT over BN254a Fr as a multi-pass ("four-step" generalised to k passes) transform.
//
// Replaces FFTAuxiliary.serialRadix2FFT (src/main/java/algebra/fft/FFTAuxiliary.java:100-123: bit-reverse swap, then
// log n stages each touching the whole array) and the reference's dormant CUDA version (algebra_fft_FFTAuxiliary.cu:
// one launch per stage, every thread recomputing its twiddle with two modular exponentiations, :96-140).
//
// Factor n = n_1 n_2 ... n_k with n_j <= 512.  Writing the input index as i = (i_1, ..., i_k) with i_1 slowest and the
// output index as k = k_1 + n_1 k_2 + n_1 n_2 k_3 + ..., pass j replaces digit i_j by k_j with an n_j-point transform
// held entirely in shared memory, and (for j < k) multiplies by the inter-pass twiddle
// omega_n^{(n / (n_j inner)) * (remaining index) * k_j}.  A CTA owns a tile of 1024 elements (kTileLogDefault) and reads and
// writes runs of 64 to 128 contiguous bytes:
//   - passes 1..k-1 walk the transform dimension with stride `inner` and take 2+ adjacent columns per CTA;
//   - the last pass reads contiguous rows and writes the transposed (natural-order) result, taking 2+ rows that are
//     adjacent in the OUTPUT per CTA.
// So there is no bit-reversal pass and the data crosses HBM 2k times (k = 3 for 2^19..2^27).  HBM is < 10 % busy in these
// passes (they are bound by the integer pipe), which is why short runs and the direct twiddle tables below cost nothing.
//
// Data stays in canonical (non-Montgomery) form end to end: every multiplication is data x twiddle, and the twiddles
// are stored as w*2^256 mod r, so mont_mul(x, w~) = x*w needs no conversions at the boundary.
//
// Inside a CTA the n_j-point transform is decimation-in-frequency, radix-8 per thread in registers (three butterfly
// stages per shared-memory round trip); outputs are picked up in bit-reversed position by the store phase.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "fp256.cuh"

namespace ozk {

static constexpr int kMaxLogT = 9;          // largest in-CTA transform by default: 512 points
// 1024-point in-CTA transforms (one column per tile, 16 KB of twiddles, four CTAs per SM) are supported and would keep 2^28 at
// three passes (10+9+9) instead of four (7+7+7+7), but measured slower on B200: 72.7 ms against 67.6 ms -- the first pass then
// needs the two-level inter-pass twiddle (one more product per element) and runs one column per tile.  OZK_NTT_MAX_LOGT=10 selects it.
static constexpr int kMaxLogTBig = 10;
static inline int max_log_t_for(int) { return kMaxLogT; }
// Tile of 1024 elements (32 KB) per 128-thread CTA, five CTAs per SM (__launch_bounds__(128, 5): 96 registers, ~100 bytes of
// spills) whose load / transform / store phases interleave (measured 2^26: 16.3 ms with 2048-element tiles and two CTAs per SM,
// 15.9 ms with 1024-element tiles and four, 14.7 ms with five together with the direct twiddle tables below).
static constexpr int kTileLogDefault = 10;
// Largest inter-pass twiddle range that gets a direct table: 2^26 entries = 2 GiB per (n, omega) plan.  HBM is < 10 % busy in
// these passes, so one more 32-byte read per element is free, and it replaces the product of two table entries (one of the
// ~13.4 multiplications per element at 2^26: 15.9 -> 15.0 ms).  Larger ranges fall back to the two-level table.
static constexpr int kDirectLogDefault = 26;
static int env_int(const char* name, int dflt, int lo, int hi) {
    if (const char* e = getenv(name)) {
        int v = atoi(e);
        if (v >= lo && v <= hi) return v;
    }
    return dflt;
}
// ... and up to 2^28 entries (8 GiB per plan) when the device has the memory to spare: at 2^27 the direct table of the first pass
// takes the transform from 31.1 to 28.8 ms.  Decided from the free device memory when the first plan is made (the plan-cache
// budget below grows with it: a forward / inverse pair at 2^28 holds 17 GiB of tables); OZK_NTT_DIRECT_LOG overrides.
static int direct_log_limit() {
    static const int lim = [] {
        if (getenv("OZK_NTT_DIRECT_LOG")) return env_int("OZK_NTT_DIRECT_LOG", kDirectLogDefault, 0, 28);
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
            cudaGetLastError();
            return kDirectLogDefault;
        }
        if (free_b >= ((size_t)96 << 30)) return 28;
        if (free_b >= ((size_t)48 << 30)) return 27;
        return kDirectLogDefault;
    }();
    return lim;
}
static constexpr int kMaxPasses = 4;

struct NttPlan {
    int log_n = 0;
    int npass = 0;
    int logt[kMaxPasses] = {0, 0, 0, 0};
    int H = 0;                    // two-level split of the inter-pass twiddle exponent
    Fr* wsub[kMaxPasses] = {};    // per pass: omega_{n_j}^t, t < n_j/2 (Montgomery form)
    Fr* tlo = nullptr;            // omega_n^e, e < 2^H
    Fr* thi = nullptr;            // omega_n^(e 2^H), e < 2^(log_n - H)
    Fr* tdir[kMaxPasses] = {};    // per pass: direct inter-pass twiddles omega_n^(outer * col * k) at [k][col], when n_j * inner is small enough
    void* block = nullptr;
    size_t bytes = 0;             // size of `block`
    unsigned long long last_use = 0;
};

// The plan cache is bounded: a plan with direct tables is up to 2 GiB per (n, omega), the prover uses four of them per domain,
// and a JNI host keeps one context per executor thread for the thread's lifetime.  Least-recently-used plans are dropped when
// a new one would push the context over the budget (OZK_NTT_CACHE_MB, default 16 GiB); a dropped plan is simply rebuilt on its
// next use (a few ms; ~15 ms for the 8 GiB table of a 2^28-point transform).
static size_t plan_budget_bytes() {
    static const size_t b = (size_t)env_int("OZK_NTT_CACHE_MB", direct_log_limit() >= 28 ? 40960 : direct_log_limit() == 27 ? 24576 : 16384, 64, 1 << 20) << 20;
    return b;
}
static unsigned long long g_plan_clock = 0;
static int plan_cache_make_room(ozk_ctx* ctx, size_t incoming) {
    size_t total = incoming;
    for (auto& kv : ctx->ntt_plans) total += kv.second->bytes;
    bool synced = false;
    while (total > plan_budget_bytes() && !ctx->ntt_plans.empty()) {
        auto victim = ctx->ntt_plans.begin();
        for (auto it = ctx->ntt_plans.begin(); it != ctx->ntt_plans.end(); ++it)
            if (it->second->last_use < victim->second->last_use) victim = it;
        if (!synced) {
            OZK_CUDA(cudaStreamSynchronize(ctx->stream));      // kernels reading the victim's tables may still be in flight
            synced = true;
        }
        total -= victim->second->bytes;
        if (victim->second->block) cudaFree(victim->second->block);
        delete victim->second;
        ctx->ntt_plans.erase(victim);
    }
    return OZK_OK;
}

// ---- small helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ Fr load_fr(const uint4* p) {
    uint4 a = p[0], b = p[1];
    Fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void store_fr(uint4* p, const Fr& r) {
    p[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    p[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

__device__ Fr fr_pow(Fr base, uint32_t e) {
    Fr r = Fr::one();
    while (e) {
        if (e & 1) r = Fr::mul(r, base);
        base = Fr::sqr(base);
        e >>= 1;
    }
    return r;
}

// ---- twiddle tables ------------------------------------------------------------------------------------------
// table[i] = omega^(i * step) in Montgomery form; `omega_canon` is the canonical 32-byte value from the caller.
// A thread fills kGenRun consecutive entries: one exponentiation, then one multiplication per entry (the direct inter-pass
// tables have up to 2^26 entries).
static constexpr uint32_t kGenRun = 16;
__global__ void ntt_gen_table(Fr* table, uint32_t count, uint32_t step, Fr omega_canon) {
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * kGenRun;
    if (i0 >= count) return;
    const Fr w = Fr::to_mont(omega_canon);
    const Fr ws = fr_pow(w, step);
    Fr v = fr_pow(ws, i0);
    const uint32_t end = min(count, i0 + kGenRun);
    for (uint32_t i = i0; i < end; i++) {
        table[i] = v;
        v = Fr::mul(v, ws);
    }
}

// Direct inter-pass twiddle table of one pass, laid out like the data the pass stores: table[(k << log_inner) + col] =
// omega^(step * col * k), k < 2^log_t, col < 2^log_inner.  (A table indexed by the product col * k has the same size but is
// read as 32-byte pieces scattered over 128-byte lines: 9.8 GB instead of 4.3 GB of DRAM reads in the first pass at 2^26.)
__global__ void ntt_gen_table2d(Fr* table, uint32_t log_t, uint32_t log_inner, uint32_t step, Fr omega_canon) {
    const size_t total = (size_t)1 << (log_t + log_inner);
    const size_t e0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * kGenRun;
    if (e0 >= total) return;
    const uint32_t run = log_inner >= 4 ? kGenRun : (1u << log_inner);     // a run never crosses a row when 2^log_inner >= kGenRun
    const Fr ws = fr_pow(Fr::to_mont(omega_canon), step);
    for (size_t e = e0; e < e0 + kGenRun && e < total; e += run) {
        const uint32_t k = (uint32_t)(e >> log_inner), col = (uint32_t)(e & (((size_t)1 << log_inner) - 1));
        const Fr wk = fr_pow(ws, k);
        Fr v = fr_pow(wk, col);
        for (uint32_t q = 0; q < run && e + q < total; q++) {
            table[e + q] = v;
            v = Fr::mul(v, wk);
        }
    }
}

// flag[0] = 1 when omega is canonical and omega^(n/2) == -1 (i.e. omega is a primitive n-th root of unity)
__global__ void ntt_check_omega(uint32_t* flag, Fr omega_canon, uint32_t half_n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    bool ok = omega_canon.is_canonical();
    if (ok) {
        Fr w = Fr::to_mont(omega_canon);
        Fr h = half_n ? fr_pow(w, half_n) : w;           // n == 1: omega must be 1
        Fr target = half_n ? Fr::neg(Fr::one()) : Fr::one();
        ok = (h == target);
    }
    flag[0] = ok ? 1u : 0u;
}

// ---- the pass kernel -----------------------------------------------------------------------------------------
struct PassArgs {
    const uint4* in;
    uint4* out;
    const Fr* wsub;
    const Fr* tlo;
    const Fr* thi;
    const Fr* tdir;        // direct table of this pass's inter-pass twiddles, or null
    uint32_t H;
    uint32_t log_n;
    uint32_t log_inner;    // non-last passes: stride of the transform dimension
    uint32_t log_outer;    // number of independent transforms per column = n / (T * inner)
    uint32_t log_c;        // columns (non-last) or rows (last) per tile
    uint32_t log_n1;       // last pass: radix of pass 1 (rows adjacent in the output differ in k_1)
    uint32_t nmid;         // last pass: number of middle passes (2..k-1) and their radices
    uint32_t logmid0, logmid1;
    uint32_t log_t;        // this pass's transform size
    // last pass of a shard-local transform inside the multi-GPU four-step (ozk_ntt_fr_scatter_dev): element i of the output
    // is multiplied by g^i (g = omega_n^rank, two-level table ptlo / pthi) and stored into the buffer of the rank that owns
    // chunk i >> log_chunk, at [my_rank][i mod chunk] -- the inter-shard twiddle and the all-to-all fused into the epilogue.
    uint32_t npeers;       // 0: plain store to `out`
    uint32_t my_rank;
    uint32_t log_chunk;
    uint32_t pH;
    const Fr* ptlo;
    const Fr* pthi;
    uint4* peer[8];
};

template <bool LAST>
struct Layout {
    // shared-memory slot (in 16-byte units, per plane) of element t of column c
    __device__ __forceinline__ static uint32_t slot(uint32_t t, uint32_t c, uint32_t log_t, uint32_t log_c) {
        if (LAST) {
            uint32_t pitch = (1u << log_t) + (1u << (log_t > 3 ? log_t - 3 : 0)) + 1;   // +1 slot per 8, odd pitch
            return c * pitch + t + (t >> 3);
        } else {
            return (t << log_c) + c;
        }
    }
    __device__ __forceinline__ static uint32_t plane_slots(uint32_t log_t, uint32_t log_c) {
        if (LAST) {
            uint32_t pitch = (1u << log_t) + (1u << (log_t > 3 ? log_t - 3 : 0)) + 1;
            return (pitch << log_c) + 4;
        } else {
            return (1u << (log_t + log_c)) + 4;    // +64 B so the two planes start in different bank groups
        }
    }
};

static size_t pass_smem_bytes(bool last, int log_t, int log_c) {
    uint32_t slots;
    if (last) {
        uint32_t pitch = (1u << log_t) + (1u << (log_t > 3 ? log_t - 3 : 0)) + 1;
        slots = (pitch << log_c) + 4;
    } else {
        slots = (1u << (log_t + log_c)) + 4;
    }
    size_t data = (size_t)slots * 16 * 2;
    size_t tw = (size_t)(log_t ? (1u << (log_t - 1)) : 1) * 32;
    return data + tw;
}

__device__ __forceinline__ Fr lds_fr(const uint4* lo, const uint4* hi, uint32_t s) {
    uint4 a = lo[s], b = hi[s];
    Fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sts_fr(uint4* lo, uint4* hi, uint32_t s, const Fr& r) {
    lo[s] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    hi[s] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// Asynchronous 16-byte global -> shared copies (LDGSTS): the tile load issues all of a thread's copies back to back and waits once,
// instead of one load-to-register / store-to-shared round trip per 16 bytes (16 dependent DRAM latencies per thread and tile).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// Q decimation-in-frequency stages on 2^Q registers.  Element u sits at index base + u*m of the transform; stage a
// pairs u with u + 2^(Q-1-a).  The twiddle of a pair depends only on (u mod half) and j = index mod m.
// M_IS_ONE (the final step, m == 1): exponents with u_local == 0 are zero, those multiplications are skipped.
// LAZY: values stay in [0, 2r) (inputs may be anywhere in that range): sums are reduced against 2r, the difference goes into the
// product as p - q + 2r < 4r and the product leaves out its final subtraction; the caller reduces once at the very end
// (fp256.cuh, "lazy" domain).  2^26: 14.57 -> 13.87 ms.  The products stay INLINED: the kernel is 111 KB of straight-line code
// that every warp runs through once per tile (15 % of the warp samples are "no instruction" stalls), but calling one shared
// copy of the product instead (operands and result in registers, 38 KB of code) was slower, 16.6 ms, and two products per call
// 16.2 ms: a called product cannot be interleaved with the additions of the neighbouring butterflies.
template <int Q, bool M_IS_ONE, bool LAZY>
__device__ __forceinline__ void dif_step(Fr (&x)[1 << Q], const Fr* wtab, uint32_t j, uint32_t log_m, uint32_t log_t) {
#pragma unroll
    for (int a = 0; a < Q; a++) {
        const int half = 1 << (Q - 1 - a);           // pair distance in units of m
        const uint32_t sh = log_t - 1 - (log_m + (Q - 1 - a));   // exponent scale: T / (2 * half * m)
#pragma unroll
        for (int ul = 0; ul < half; ul++) {
            const bool trivial = M_IS_ONE && ul == 0;
            Fr w;
            if (!trivial) w = wtab[(((uint32_t)ul << log_m) + j) << sh];
#pragma unroll
            for (int blk = 0; blk < (1 << Q); blk += 2 * half) {
                Fr& p = x[blk + ul];
                Fr& q = x[blk + ul + half];
                if (LAZY) {
                    const Fr s = Fr::add_lazy(p, q);
                    q = trivial ? Fr::sub_lazy(p, q) : Fr::mul_lazy(Fr::sub_lazy_wide(p, q), w);
                    p = s;
                } else {
                    Fr s = Fr::add(p, q);
                    Fr d = Fr::sub(p, q);
                    p = s;
                    q = trivial ? d : Fr::mul(d, w);
                }
            }
        }
    }
}

template <int Q, bool M_IS_ONE, bool LAST>
__device__ __forceinline__ void run_step(uint4* lo, uint4* hi, const Fr* wtab, uint32_t log_m, uint32_t log_t, uint32_t log_c) {
    const uint32_t groups_per_col = 1u << (log_t - Q);
    const uint32_t ngroups = groups_per_col << log_c;
    for (uint32_t gi = threadIdx.x; gi < ngroups; gi += blockDim.x) {
        uint32_t g, c;
        if (LAST) {
            g = gi & (groups_per_col - 1);
            c = gi >> (log_t - Q);
        } else {
            c = gi & ((1u << log_c) - 1);
            g = gi >> log_c;
        }
        const uint32_t j = g & ((1u << log_m) - 1);
        const uint32_t base = ((g >> log_m) << (log_m + Q)) + j;
        Fr x[1 << Q];
#pragma unroll
        for (int u = 0; u < (1 << Q); u++) x[u] = lds_fr(lo, hi, Layout<LAST>::slot(base + ((uint32_t)u << log_m), c, log_t, log_c));
        dif_step<Q, M_IS_ONE, true>(x, wtab, j, log_m, log_t);
#pragma unroll
        for (int u = 0; u < (1 << Q); u++) sts_fr(lo, hi, Layout<LAST>::slot(base + ((uint32_t)u << log_m), c, log_t, log_c), x[u]);
    }
}

template <bool LAST>
__global__ void __launch_bounds__(128, 5) ntt_pass_kernel(PassArgs a) {
    extern __shared__ uint4 smem[];
    const uint32_t LOGT = a.log_t;
    const uint32_t T = 1u << LOGT;
    const uint32_t log_c = a.log_c;
    const uint32_t C = 1u << log_c;
    uint4* lo = smem;
    uint4* hi = smem + Layout<LAST>::plane_slots(LOGT, log_c);
    Fr* wtab = reinterpret_cast<Fr*>(hi + Layout<LAST>::plane_slots(LOGT, log_c));

    // ---- tile decode
    size_t in_base, out_base;
    uint32_t col0 = 0;     // first column index (non-last): the "remaining index" of the inter-pass twiddle
    if (!LAST) {
        const uint32_t tiles_per_o = 1u << (a.log_inner - log_c);
        const uint32_t o = blockIdx.x >> (a.log_inner - log_c);
        col0 = (blockIdx.x & (tiles_per_o - 1)) << log_c;
        in_base = (((size_t)o << LOGT) << a.log_inner) + col0;
        out_base = in_base;
    } else {
        // rows o = (k_1, rho) with k_1 slowest; a tile takes C consecutive k_1 for one rho
        const uint32_t log_rho = a.log_outer - a.log_n1;
        const uint32_t rho = blockIdx.x & ((1u << log_rho) - 1);
        const uint32_t k10 = (blockIdx.x >> log_rho) << log_c;
        in_base = (((size_t)k10 << log_rho) + rho) << LOGT;
        // output index = k_1 + n_1 * (k_2 + n_2 * (k_3 ...)) + (n/T) * t ; rho = (k_2, k_3, ...) with k_2 slowest.
        // Two middle passes at most (kMaxPasses == 4): rho = k_2 * n_3 + k_3  ->  k_2 + n_2 * k_3.
        uint32_t rev = rho;
        if (a.nmid == 2) {
            const uint32_t k3 = rho & ((1u << a.logmid1) - 1);
            const uint32_t k2 = rho >> a.logmid1;
            rev = k2 + (k3 << a.logmid0);
        }
        out_base = (size_t)k10 + ((size_t)rev << a.log_n1);
    }

    // ---- twiddles of the in-CTA transform
    for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) {     // T/2 entries of two uint4 each
        cp_async16(reinterpret_cast<uint4*>(wtab) + i, reinterpret_cast<const uint4*>(a.wsub) + i);
    }

    // ---- load the tile (16 bytes per thread per step, fastest index = the contiguous one in global memory)
    if (!LAST) {
        const uint32_t per_row = C * 2;                     // uint4 per row of the tile
        const uint32_t total = T * per_row;
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
            uint32_t t = i >> (log_c + 1), r = i & (per_row - 1);
            uint32_t c = r >> 1, h = r & 1;
            cp_async16(&(h ? hi : lo)[Layout<LAST>::slot(t, c, LOGT, log_c)], &a.in[(in_base + ((size_t)t << a.log_inner) + c) * 2 + h]);
        }
    } else {
        const uint32_t log_rho = a.log_outer - a.log_n1;
        const uint32_t per_row = T * 2;
        const uint32_t total = C * per_row;
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
            uint32_t c = i >> (LOGT + 1), r = i & (per_row - 1);
            uint32_t t = r >> 1, h = r & 1;
            cp_async16(&(h ? hi : lo)[Layout<LAST>::slot(t, c, LOGT, log_c)], &a.in[(in_base + ((((size_t)c << log_rho)) << LOGT) + t) * 2 + h]);
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- DIF stages from half-size T/2 down to 1: (LOGT mod 3) single stages first, then radix-8 steps.  (Merging two leftover
    // stages into one radix-4 step was measured slower: 15.7 against 15.0 ms at 2^26.)
    {
        uint32_t log_m = LOGT;
        uint32_t r0 = LOGT % 3;
        if (LOGT < 3) r0 = LOGT;
        while (r0 > 0) {
            log_m -= 1;
            r0 -= 1;
            if (log_m == 0) run_step<1, true, LAST>(lo, hi, wtab, 0, LOGT, log_c);
            else run_step<1, false, LAST>(lo, hi, wtab, log_m, LOGT, log_c);
            __syncthreads();
        }
        while (log_m > 0) {
            log_m -= 3;
            if (log_m == 0) run_step<3, true, LAST>(lo, hi, wtab, 0, LOGT, log_c);
            else run_step<3, false, LAST>(lo, hi, wtab, log_m, LOGT, log_c);
            __syncthreads();
        }
    }

    // ---- store: element k of the transform sits at bit-reversed position; lanes run along the contiguous output index.
    // kStoreBatch elements per thread are handled together: their inter-pass twiddles (one 32-byte global read each, no reuse) are
    // requested first, so the DRAM latencies of a batch overlap instead of adding up.
    const uint32_t total = T << log_c;
    constexpr int kStoreBatch = 4;
    for (uint32_t i0 = threadIdx.x; i0 < total; i0 += blockDim.x * kStoreBatch) {
        Fr w[kStoreBatch];
        if (!LAST) {
#pragma unroll
            for (int u = 0; u < kStoreBatch; u++) {
                const uint32_t i = i0 + u * blockDim.x;
                const uint32_t c = i & (C - 1), k = i >> log_c;
                const uint32_t ek = (col0 + c) * k;
                if (i < total && ek != 0) {
                    if (a.tdir) {
                        w[u] = a.tdir[((size_t)k << a.log_inner) + col0 + c];
                    } else {
                        const uint32_t e = ek << a.log_outer;
                        w[u] = Fr::mul_lazy(a.tlo[e & ((1u << a.H) - 1)], a.thi[e >> a.H]);     // < 2r: fine as a factor (v w < 4 r^2)
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kStoreBatch; u++) {
            const uint32_t i = i0 + u * blockDim.x;
            if (i >= total) break;
            const uint32_t c = i & (C - 1);
            const uint32_t k = i >> log_c;
            const uint32_t pos = LOGT ? (__brev(k) >> (32 - LOGT)) : 0;
            Fr v = lds_fr(lo, hi, Layout<LAST>::slot(pos, c, LOGT, log_c));      // in [0, 2r)
            if (!LAST) {
                // inter-pass twiddle omega_n^(outer * column * k); the value stays in [0, 2r) between the passes
                if ((col0 + c) * k != 0) v = Fr::mul_lazy(v, w[u]);
                store_fr(a.out + (out_base + ((size_t)k << a.log_inner) + c) * 2, v);
            } else {
                const size_t idx = out_base + ((size_t)k << a.log_outer) + c;
                if (a.npeers) {
                    if (idx != 0 && a.my_rank != 0) {
                        const Fr wp = Fr::mul_lazy(a.ptlo[idx & ((1u << a.pH) - 1)], a.pthi[idx >> a.pH]);
                        v = Fr::mul_lazy(v, wp);
                    }
                    v = Fr::reduce_lazy(v);                      // canonical on the wire
                    const uint32_t dest = (uint32_t)(idx >> a.log_chunk);
                    const size_t local = ((size_t)a.my_rank << a.log_chunk) + (idx & (((size_t)1 << a.log_chunk) - 1));
                    store_fr(a.peer[dest] + local * 2, v);       // peer memory over NVLink (or this GPU's own buffer)
                } else {
                    store_fr(a.out + idx * 2, Fr::reduce_lazy(v));   // the last pass hands out canonical values
                }
            }
        }
    }
}

// ---- elementwise helpers ---------------------------------------------------------------------------------------
// out[i] = in[i] * scale * coset^i   (scale / coset given in Montgomery form; `has_coset` selects the power term)
__global__ void __launch_bounds__(256) fr_scale_powers_kernel(const uint4* in, uint4* out, size_t n, Fr scale_canon, Fr coset_canon,
                                                              int has_scale, int has_coset, unsigned long long first_index) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (i >= n) return;
    Fr f = has_scale ? Fr::to_mont(scale_canon) : Fr::one(), step = Fr::one();
    if (has_coset) {
        // f = scale * coset^i, step = coset^stride
        const Fr coset = Fr::to_mont(coset_canon);
        Fr b = coset;
        unsigned long long e = first_index + i;
        while (e) {
            if (e & 1) f = Fr::mul(f, b);
            b = Fr::sqr(b);
            e >>= 1;
        }
        b = coset;
        e = stride;
        while (e) {
            if (e & 1) step = Fr::mul(step, b);
            b = Fr::sqr(b);
            e >>= 1;
        }
    }
    for (; i < n; i += stride) {
        Fr v = load_fr(in + i * 2);
        v = Fr::mul(v, f);
        store_fr(out + i * 2, v);
        if (has_coset) f = Fr::mul(f, step);
    }
}

// out[i] = a[i] * b[i] - c[i] (c may be null): the pointwise step between the transforms of the QAP witness map
// (R1CStoQAP.java:180-182,211-214).  All values canonical: to_mont(a) * b = a*b in canonical form.
__global__ void __launch_bounds__(256) fr_mul_sub_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ c,
                                                         uint4* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr x = Fr::to_mont(load_fr(a + i * 2));
        Fr v = Fr::mul(x, load_fr(b + i * 2));
        if (c) v = Fr::sub(v, load_fr(c + i * 2));
        store_fr(out + i * 2, v);
    }
}

// ---- Lagrange coefficients ----------------------------------------------------------------------------------------------
// out[i] = L_i(t) over S = {omega^0 .. omega^(m-1)} = (t^m - 1) / m * omega^i / (t - omega^i), or the unit vector when t is
// omega^i itself: FFTAuxiliary.serialRadix2LagrangeCoefficients (src/main/java/algebra/fft/FFTAuxiliary.java:249-302),
// which inverts (t - omega^i) one element at a time.  Here every thread takes kLagBatch consecutive indices and shares one
// inversion among them (Montgomery's trick); hit[0] = 1, hit[1] = i when some t - omega^i is zero.
static constexpr int kLagBatch = 32;

// consts[0] = (t^m - 1) / m in Montgomery form
__global__ void fr_lagrange_setup(Fr* consts, Fr t_canon, uint32_t log_m) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Fr t = Fr::to_mont(t_canon);
    Fr tm = t;
    for (uint32_t k = 0; k < log_m; k++) tm = Fr::sqr(tm);
    Fr mm = Fr::zero();
    mm.v[0] = 1u << (log_m & 31);                 // m = 2^log_m <= 2^28, canonical
    consts[0] = Fr::mul(Fr::sub(tm, Fr::one()), Fr::inv(Fr::to_mont(mm)));
}

__global__ void __launch_bounds__(128) fr_lagrange_kernel(uint4* __restrict__ out, size_t m, Fr t_canon, Fr omega_canon,
                                                         const Fr* __restrict__ consts, uint32_t* __restrict__ hit) {
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * kLagBatch;
    if (i0 >= m) return;
    const int cnt = (int)((m - i0) < (size_t)kLagBatch ? (m - i0) : (size_t)kLagBatch);
    const Fr t = Fr::to_mont(t_canon), omega = Fr::to_mont(omega_canon);
    Fr r = fr_pow(omega, (uint32_t)i0);             // omega^i0
    Fr num[kLagBatch], prefix[kLagBatch], den[kLagBatch];
#pragma unroll 1
    for (int k = 0; k < cnt; k++) {
        num[k] = r;
        Fr d = Fr::sub(t, r);
        if (d.is_zero()) {
            hit[0] = 1u;
            hit[1] = (uint32_t)(i0 + k);
            d = Fr::one();
        }
        den[k] = d;
        prefix[k] = k ? Fr::mul(prefix[k - 1], d) : d;
        r = Fr::mul(r, omega);
    }
    Fr inv = Fr::inv(prefix[cnt - 1]);
    const Fr l0 = consts[0];
#pragma unroll 1
    for (int k = cnt - 1; k >= 0; k--) {
        const Fr di = k ? Fr::mul(inv, prefix[k - 1]) : inv;       // 1 / den[k]
        inv = Fr::mul(inv, den[k]);
        const Fr v = Fr::mul(Fr::mul(l0, num[k]), di);
        store_fr(out + (i0 + k) * 2, Fr::from_mont(v));
    }
}

// t is a point of S: the coefficients are the unit vector e_{hit[1]} (FFTAuxiliary.java:268-279)
__global__ void __launch_bounds__(256) fr_lagrange_unit(uint4* __restrict__ out, size_t m, const uint32_t* __restrict__ hit) {
    if (hit[0] == 0) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        Fr v = Fr::zero();
        if (i == hit[1]) v.v[0] = 1;
        store_fr(out + i * 2, v);
    }
}

// ---- sparse R1CS rows times the assignment ---------------------------------------------------------------------------------
// out[i] = sum_{k in [row_ptr[i], row_ptr[i+1])} coeff[k] * z[col[k]] mod r: the evaluation of the linear combinations a_i, b_i,
// c_i of every constraint at the full assignment, the host loop of R1CStoQAP.R1CStoQAPWitness
// (src/main/java/reductions/r1cs_to_qap/R1CStoQAP.java:143-160 with LinearCombination.evaluate, relations/objects/
// LinearCombination.java:47-55).  Rows of up to kSpmvShort terms take one thread; longer rows (the reference's synthetic
// circuit ends with one constraint over all variables, R1CSConstruction.java:91-104) are queued and take one block each.
static constexpr uint32_t kSpmvShort = 64;
static constexpr uint32_t kSpmvLongCap = 4096;

// coeff == nullptr: every coefficient is 1 (the reference's synthetic circuit, R1CSConstruction.java:48-104, and most gates of
// real circuits).  A column index >= z_len sets flag bit 0 and contributes nothing, so a malformed matrix cannot read out of bounds.
__device__ __forceinline__ Fr spmv_term(const uint4* coeff, const uint4* z, const uint32_t* col, size_t k, size_t z_len, uint32_t* flag) {
    const size_t c = col[k];
    if (c >= z_len) {
        atomicOr(flag, 1u);
        return Fr::zero();
    }
    const Fr zv = load_fr(z + c * 2);
    if (!coeff) return zv;
    return Fr::mul(Fr::to_mont(load_fr(coeff + k * 2)), zv);     // canonical coeff * z
}

__global__ void __launch_bounds__(256) fr_spmv_short(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, const uint4* __restrict__ coeff,
                                                     const uint4* __restrict__ z, size_t z_len, size_t rows, uint4* __restrict__ out,
                                                     uint32_t* __restrict__ long_rows, uint32_t* __restrict__ long_count, uint32_t* __restrict__ flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const uint32_t lo = row_ptr[i], hi = row_ptr[i + 1];
    if (hi - lo > kSpmvShort) {
        const uint32_t q = atomicAdd(long_count, 1u);
        if (q < kSpmvLongCap) long_rows[q] = (uint32_t)i;
        return;
    }
    Fr acc = Fr::zero();
    for (uint32_t k = lo; k < hi; k++) acc = Fr::add(acc, spmv_term(coeff, z, col, k, z_len, flag));
    store_fr(out + i * 2, acc);
}

// Long rows are cut into segments of kSpmvSeg terms; blocks walk the (row, segment) pairs, each leaving one partial sum, and
// fr_spmv_long_finish adds the partials of every row (a row over all 2^24 variables -- the last constraint of the synthetic
// circuit -- is 1024 segments spread over all SMs instead of one block's 65 536 serial steps).
static constexpr uint32_t kSpmvSeg = 1u << 14;
static constexpr uint32_t kSpmvMaxSegs = 1u << 17;        // partial sums available (4 MB of scratch)
__global__ void __launch_bounds__(256) fr_spmv_long(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, const uint4* __restrict__ coeff,
                                                    const uint4* __restrict__ z, size_t z_len, uint4* __restrict__ partial,
                                                    const uint32_t* __restrict__ long_rows, const uint32_t* __restrict__ long_count,
                                                    uint32_t* __restrict__ flag) {
    __shared__ Fr part[256];
    const uint32_t total = min(*long_count, kSpmvLongCap);
    uint32_t seg_base = 0;                                  // index of the first segment of row q among all segments
    for (uint32_t q = 0; q < total; q++) {
        const uint32_t i = long_rows[q];
        const uint32_t lo = row_ptr[i], hi = row_ptr[i + 1];
        const uint32_t nseg = (hi - lo + kSpmvSeg - 1) / kSpmvSeg;
        // segments of this row taken by this block: those whose global segment index is congruent to blockIdx.x
        uint32_t first = (blockIdx.x + gridDim.x - seg_base % gridDim.x) % gridDim.x;
        for (uint32_t sg = first; sg < nseg; sg += gridDim.x) {
            if (seg_base + sg >= kSpmvMaxSegs) {
                if (threadIdx.x == 0) atomicOr(flag, 2u);
                break;
            }
            const uint32_t a = lo + sg * kSpmvSeg, b = min(hi, a + kSpmvSeg);
            Fr acc = Fr::zero();
            for (uint32_t k = a + threadIdx.x; k < b; k += blockDim.x) acc = Fr::add(acc, spmv_term(coeff, z, col, k, z_len, flag));
            part[threadIdx.x] = acc;
            __syncthreads();
            for (uint32_t s = 128; s >= 1; s >>= 1) {
                if (threadIdx.x < s) part[threadIdx.x] = Fr::add(part[threadIdx.x], part[threadIdx.x + s]);
                __syncthreads();
            }
            if (threadIdx.x == 0) store_fr(partial + (size_t)(seg_base + sg) * 2, part[0]);
            __syncthreads();
        }
        seg_base += nseg;
    }
}
__global__ void __launch_bounds__(128) fr_spmv_long_finish(const uint32_t* __restrict__ row_ptr, const uint4* __restrict__ partial, uint4* __restrict__ out,
                                                           const uint32_t* __restrict__ long_rows, const uint32_t* __restrict__ long_count) {
    const uint32_t total = min(*long_count, kSpmvLongCap);
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total) return;
    uint32_t seg_base = 0;
    for (uint32_t p = 0; p < q; p++) {
        const uint32_t ip = long_rows[p];
        seg_base += (row_ptr[ip + 1] - row_ptr[ip] + kSpmvSeg - 1) / kSpmvSeg;
    }
    const uint32_t i = long_rows[q];
    const uint32_t nseg = (row_ptr[i + 1] - row_ptr[i] + kSpmvSeg - 1) / kSpmvSeg;
    Fr acc = Fr::zero();
    for (uint32_t sg = 0; sg < nseg && seg_base + sg < kSpmvMaxSegs; sg++) acc = Fr::add(acc, load_fr(partial + (size_t)(seg_base + sg) * 2));
    store_fr(out + (size_t)i * 2, acc);
}

// out[i] = ca * a[i] + cb * b[i] + cc * c[i] (b, c optional): the linear combinations of the setup, beta*At + alpha*Bt + Ct and
// their division by gamma / delta (SerialSetup.java:66-85), on vectors that stay on the device.  Constants canonical.
__global__ void __launch_bounds__(256) fr_lincomb_kernel(uint4* __restrict__ out, size_t n, const uint4* __restrict__ a, Fr ca_canon,
                                                         const uint4* __restrict__ b, Fr cb_canon, const uint4* __restrict__ c, Fr cc_canon) {
    const Fr ca = Fr::to_mont(ca_canon), cb = Fr::to_mont(cb_canon), cc = Fr::to_mont(cc_canon);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr v = Fr::mul(ca, load_fr(a + i * 2));                      // Montgomery constant * canonical value = canonical product
        if (b) v = Fr::add(v, Fr::mul(cb, load_fr(b + i * 2)));
        if (c) v = Fr::add(v, Fr::mul(cc, load_fr(c + i * 2)));
        store_fr(out + i * 2, v);
    }
}

// flag |= 4 when some element of in[0..n) is not reduced mod r (the host-pointer entry points reject such input; Fr::add / sub
// rely on operands below the modulus)
__global__ void __launch_bounds__(256) fr_check_canonical_kernel(const uint4* __restrict__ in, size_t n, uint32_t* __restrict__ flag) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) bad |= !load_fr(in + i * 2).is_canonical();
    if (bad) atomicOr(flag, 4u);
}

// ---- small cross-shard DFT (the second step of the multi-GPU transform) ---------------------------------------------------
// out[k1 * len + j] = sum_{i1 < G} in[i1 * len + j] * omega_G^(i1 k1), G = 2^Q <= 8: one thread per j, all G values in registers.
template <int Q>
__global__ void __launch_bounds__(256) fr_dft_small_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t len,
                                                           const Fr* __restrict__ wtab) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    Fr x[1 << Q];
#pragma unroll
    for (int u = 0; u < (1 << Q); u++) x[u] = load_fr(in + ((size_t)u * len + j) * 2);
    dif_step<Q, true, false>(x, wtab, 0, 0, Q);
#pragma unroll
    for (int u = 0; u < (1 << Q); u++) {
        const uint32_t k1 = __brev((uint32_t)u) >> (32 - Q);
        store_fr(out + ((size_t)k1 * len + j) * 2, x[u]);
    }
}

// The mirrored first step of the multi-GPU transform (ozk_fr_dft_small_scatter_dev): this rank holds x[a * M + i2] for all
// a < G and its own block of i2 (len values starting at rank * len).  One thread per i2: the G-point transform over a, the
// twiddle omega_n^(i2 * k1) (two-level table), and the store of result k1 straight into rank k1's buffer at position i2,
// so every rank ends up with a full natural-order M-vector: the exchange is the epilogue of this kernel.
struct PeerPtrs {
    uint4* p[8];
};
template <int Q>
__global__ void __launch_bounds__(256) fr_dft_small_scatter_kernel(const uint4* __restrict__ in, size_t len, const Fr* __restrict__ wtab,
                                                                   const Fr* __restrict__ tlo, const Fr* __restrict__ thi, uint32_t H,
                                                                   uint32_t rank, PeerPtrs peers) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    Fr x[1 << Q];
#pragma unroll
    for (int u = 0; u < (1 << Q); u++) x[u] = load_fr(in + ((size_t)u * len + j) * 2);
    dif_step<Q, true, false>(x, wtab, 0, 0, Q);
    const size_t i2 = (size_t)rank * len + j;
#pragma unroll
    for (int u = 0; u < (1 << Q); u++) {
        const uint32_t k1 = __brev((uint32_t)u) >> (32 - Q);
        Fr v = x[u];
        const uint32_t e = (uint32_t)i2 * k1;            // < n <= 2^28
        if (e != 0) v = Fr::mul(v, Fr::mul(tlo[e & ((1u << H) - 1)], thi[e >> H]));
        store_fr(peers.p[k1] + i2 * 2, v);
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
static void fr_from_bytes(Fr& f, const uint8_t* b) { memcpy(f.v, b, 32); }

void ntt_free_plans(ozk_ctx* ctx) {
    for (auto& kv : ctx->ntt_plans) {
        if (kv.second->block) cudaFree(kv.second->block);
        delete kv.second;
    }
    ctx->ntt_plans.clear();
}

static int ntt_get_plan(ozk_ctx* ctx, int log_n, const uint8_t omega[32], NttPlan** out) {
    std::string key((const char*)omega, 32);
    key.push_back((char)log_n);
    auto it = ctx->ntt_plans.find(key);
    if (it != ctx->ntt_plans.end()) {
        it->second->last_use = ++g_plan_clock;
        *out = it->second;
        return OZK_OK;
    }
    Fr om;
    fr_from_bytes(om, omega);
    // validate omega on the device
    uint32_t* flag = (uint32_t*)ctx->pinned;
    OZK_TRY(ctx->io_out.reserve(256, ctx->stream));
    ntt_check_omega<<<1, 32, 0, ctx->stream>>>((uint32_t*)ctx->io_out.p, om, log_n ? (1u << (log_n - 1)) : 0u);
    OZK_CUDA(cudaMemcpyAsync(flag, ctx->io_out.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    OZK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag[0] != 1u) {
        set_error("ozk_ntt_fr: omega is not a canonical primitive 2^%d-th root of unity in Fr", log_n);
        return OZK_ERR_DOMAIN;
    }
    NttPlan* p = new NttPlan();
    p->log_n = log_n;
    const int max_t = env_int("OZK_NTT_MAX_LOGT", max_log_t_for(log_n), 7, kMaxLogTBig);
    p->npass = log_n <= max_t ? 1 : (log_n + max_t - 1) / max_t;
    {
        int rem = log_n;
        for (int j = 0; j < p->npass; j++) {
            int left = p->npass - j;
            p->logt[j] = (rem + left - 1) / left;      // balanced, larger radices first
            rem -= p->logt[j];
        }
    }
    p->H = (log_n + 1) / 2;
    size_t count = 0;
    size_t off_w[kMaxPasses];
    for (int j = 0; j < p->npass; j++) {
        off_w[j] = count;
        count += p->logt[j] ? (size_t)1 << (p->logt[j] - 1) : 1;
    }
    size_t off_lo = count;
    count += (size_t)1 << p->H;
    size_t off_hi = count;
    count += (size_t)1 << (log_n - p->H);
    // direct tables: pass j (not the last) needs exponents (column * k) < n_j * inner_j = 2^(log_n - log_outer_j)
    size_t off_dir[kMaxPasses] = {};
    int dir_log[kMaxPasses] = {};
    {
        int log_outer = 0;
        for (int j = 0; j + 1 < p->npass; j++) {
            const int range_log = log_n - log_outer;
            if (range_log <= direct_log_limit()) {
                off_dir[j] = count;
                dir_log[j] = range_log;
                count += (size_t)1 << range_log;
            }
            log_outer += p->logt[j];
        }
    }
    p->bytes = count * sizeof(Fr);
    p->last_use = ++g_plan_clock;
    if (plan_cache_make_room(ctx, p->bytes) != OZK_OK || cudaMalloc(&p->block, p->bytes) != cudaSuccess) {
        cudaGetLastError();
        delete p;
        set_error("ozk_ntt_fr: out of device memory for the twiddle tables of a 2^%d-point transform", log_n);
        return OZK_ERR_CUDA;
    }
    Fr* base = (Fr*)p->block;
    for (int j = 0; j < p->npass; j++) {
        p->wsub[j] = base + off_w[j];
        uint32_t cnt = p->logt[j] ? 1u << (p->logt[j] - 1) : 1u;
        ntt_gen_table<<<(cnt / kGenRun + 128) / 128, 128, 0, ctx->stream>>>(p->wsub[j], cnt, 1u << (log_n - p->logt[j]), om);
    }
    p->tlo = base + off_lo;
    p->thi = base + off_hi;
    {
        uint32_t cnt = 1u << p->H;
        ntt_gen_table<<<(cnt / kGenRun + 128) / 128, 128, 0, ctx->stream>>>(p->tlo, cnt, 1u, om);
        cnt = 1u << (log_n - p->H);
        ntt_gen_table<<<(cnt / kGenRun + 128) / 128, 128, 0, ctx->stream>>>(p->thi, cnt, 1u << p->H, om);
    }
    {
        int log_outer = 0;
        for (int j = 0; j + 1 < p->npass; j++) {
            if (dir_log[j]) {
                p->tdir[j] = base + off_dir[j];
                uint32_t cnt = 1u << dir_log[j];
                // range = n_j * inner_j: rows k < n_j, columns col < inner_j = 2^(range_log - logt[j])
                ntt_gen_table2d<<<(cnt / kGenRun + 128) / 128, 128, 0, ctx->stream>>>(p->tdir[j], (uint32_t)p->logt[j], (uint32_t)(dir_log[j] - p->logt[j]),
                                                                                     1u << log_outer, om);
            }
            log_outer += p->logt[j];
        }
    }
    OZK_CUDA(cudaGetLastError());
    ctx->ntt_plans[key] = p;
    *out = p;
    return OZK_OK;
}

template <bool LAST>
static int launch_pass(ozk_ctx* ctx, int log_t, const PassArgs& a, uint32_t grid) {
    size_t smem = pass_smem_bytes(LAST, log_t, a.log_c);
    uint32_t threads = 1u << (log_t + a.log_c >= 3 ? log_t + a.log_c - 3 : 0);
    if (threads < 32) threads = 32;
    if (threads > 256) threads = 256;
    static bool attr_done[64] = {};
    if (!attr_done[ctx->device & 63]) {
        OZK_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<LAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_done[ctx->device & 63] = true;
    }
    ntt_pass_kernel<LAST><<<grid, threads, smem, ctx->stream>>>(a);
    ctx->launches += 1;
    OZK_CUDA(cudaGetLastError());
    return OZK_OK;
}

struct ScatterDesc {
    uint32_t npeers = 0, my_rank = 0, log_chunk = 0, pH = 0;
    const Fr* ptlo = nullptr;
    const Fr* pthi = nullptr;
    uint4* peer[8] = {};
};

static int ntt_run(ozk_ctx* ctx, const void* d_in, void* d_out, int log_n, const uint8_t omega[32], const ScatterDesc* sc = nullptr) {
    if (log_n == 0 && !sc) {
        // n == 1: FFTAuxiliary.serialRadix2FFT returns at once whatever omega is (FFTAuxiliary.java:64-66)
        if (d_in != d_out) OZK_CUDA(cudaMemcpyAsync(d_out, d_in, 32, cudaMemcpyDeviceToDevice, ctx->stream));
        return OZK_OK;
    }
    NttPlan* p;
    OZK_TRY(ntt_get_plan(ctx, log_n, omega, &p));
    const size_t bytes = ((size_t)32) << log_n;
    void* work = nullptr;
    if (p->npass > 1) {
        OZK_TRY(ctx->work.reserve(bytes, ctx->stream));
        work = ctx->work.p;
    }
    int log_inner = log_n;
    for (int j = 0; j < p->npass; j++) {
        const int lt = p->logt[j];
        log_inner -= lt;
        const bool last = (j == p->npass - 1);
        PassArgs a;
        memset(&a, 0, sizeof a);
        a.in = (const uint4*)(j == 0 ? d_in : work);
        a.out = (uint4*)(last ? d_out : work);
        a.wsub = p->wsub[j];
        a.tlo = p->tlo;
        a.thi = p->thi;
        a.tdir = last ? nullptr : p->tdir[j];
        a.H = p->H;
        a.log_n = log_n;
        a.log_inner = log_inner;
        a.log_outer = log_n - lt - log_inner;
        a.log_t = lt;
        if (!last) {
            int lc = std::max(0, env_int("OZK_NTT_TILE_LOG", kTileLogDefault, 9, 10) - lt);
            if (lc > log_inner) lc = log_inner;
            a.log_c = lc;
            uint32_t grid = 1u << (log_n - lt - lc);
            OZK_TRY(launch_pass<false>(ctx, lt, a, grid));
        } else {
            a.log_n1 = p->npass > 1 ? p->logt[0] : 0;
            a.nmid = p->npass > 2 ? p->npass - 2 : 0;
            a.logmid0 = a.nmid > 0 ? p->logt[1] : 0;
            a.logmid1 = a.nmid > 1 ? p->logt[2] : 0;
            int lc = std::max(0, env_int("OZK_NTT_TILE_LOG", kTileLogDefault, 9, 10) - lt);
            if (lc > (int)a.log_n1) lc = a.log_n1;
            a.log_c = lc;
            if (sc) {
                a.npeers = sc->npeers;
                a.my_rank = sc->my_rank;
                a.log_chunk = sc->log_chunk;
                a.pH = sc->pH;
                a.ptlo = sc->ptlo;
                a.pthi = sc->pthi;
                for (int r = 0; r < 8; r++) a.peer[r] = sc->peer[r];
            }
            uint32_t grid = 1u << (log_n - lt - lc);
            OZK_TRY(launch_pass<true>(ctx, lt, a, grid));
        }
    }
    return OZK_OK;
}

static int ilog2_exact(size_t n) {
    if (n == 0 || (n & (n - 1))) return -1;
    int l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}

// byte-wise "value < r" on a 32-byte little-endian number (plain comparison; no field arithmetic on the host)
static bool fr_bytes_canonical(const uint8_t* b) {
    static const uint32_t mod[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    uint32_t v[8];
    memcpy(v, b, 32);
    for (int i = 7; i >= 0; i--) {
        if (v[i] < mod[i]) return true;
        if (v[i] > mod[i]) return false;
    }
    return false;
}

// two-level table of g^e, e < 2^log_range (tlo[e] = g^e, e < 2^H; thi[e] = g^(e 2^H)), cached on the context like a plan
static int get_pow_table(ozk_ctx* ctx, const uint8_t base[32], int log_range, NttPlan** out) {
    if (!fr_bytes_canonical(base)) {
        set_error("twiddle base is not reduced mod r");
        return OZK_ERR_DOMAIN;
    }
    std::string key("t");
    key.append((const char*)base, 32);
    key.push_back((char)log_range);
    auto it = ctx->ntt_plans.find(key);
    if (it != ctx->ntt_plans.end()) {
        it->second->last_use = ++g_plan_clock;
        *out = it->second;
        return OZK_OK;
    }
    NttPlan* tw = new NttPlan();
    tw->log_n = log_range;
    tw->H = (log_range + 1) / 2;
    const size_t nlo = (size_t)1 << tw->H, nhi = (size_t)1 << (log_range - tw->H);
    tw->bytes = (nlo + nhi) * sizeof(Fr);
    tw->last_use = ++g_plan_clock;
    if (plan_cache_make_room(ctx, tw->bytes) != OZK_OK || cudaMalloc(&tw->block, tw->bytes) != cudaSuccess) {
        delete tw;
        set_error("twiddle table: out of device memory");
        cudaGetLastError();
        return OZK_ERR_CUDA;
    }
    tw->tlo = (Fr*)tw->block;
    tw->thi = tw->tlo + nlo;
    Fr g;
    fr_from_bytes(g, base);
    ntt_gen_table<<<(unsigned)((nlo / kGenRun + 128) / 128), 128, 0, ctx->stream>>>(tw->tlo, (uint32_t)nlo, 1u, g);
    ntt_gen_table<<<(unsigned)((nhi / kGenRun + 128) / 128), 128, 0, ctx->stream>>>(tw->thi, (uint32_t)nhi, 1u << tw->H, g);
    ctx->launches += 2;
    ctx->ntt_plans[key] = tw;
    *out = tw;
    return OZK_OK;
}

static int scale_powers(ozk_ctx* ctx, const void* d_in, void* d_out, size_t n, const uint8_t* scale, const uint8_t* coset,
                        unsigned long long first_index = 0) {
    Fr s, g;
    memset(&s, 0, sizeof s);
    memset(&g, 0, sizeof g);
    // the constants go to the kernel in canonical form; it converts them to Montgomery form itself
    if (scale) {
        if (!fr_bytes_canonical(scale)) {
            set_error("fr: scale factor is not reduced mod r");
            return OZK_ERR_DOMAIN;
        }
        fr_from_bytes(s, scale);
    }
    if (coset) {
        if (!fr_bytes_canonical(coset)) {
            set_error("fr: coset generator is not reduced mod r");
            return OZK_ERR_DOMAIN;
        }
        fr_from_bytes(g, coset);
    }
    size_t blocks = (n + 255) / 256;
    size_t cap = (size_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    ctx->launches += 1;
    fr_scale_powers_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const uint4*)d_in, (uint4*)d_out, n, s, g, scale ? 1 : 0, coset ? 1 : 0, first_index);
    OZK_CUDA(cudaGetLastError());
    return OZK_OK;
}

// Range check of the host-pointer entry points (they synchronise for the download anyway): one HBM-bound pass over the uploaded
// input that raises a flag, read back after the download.  The "_dev" entry points only enqueue and state the precondition.
static int check_canonical_begin(ozk_ctx* ctx, const void* d_in, size_t n) {
    OZK_TRY(ctx->io_out.reserve(512 + (size_t)kSpmvLongCap * 4, ctx->stream));
    uint32_t* flag = (uint32_t*)((char*)ctx->io_out.p + 400);
    OZK_CUDA(cudaMemsetAsync(flag, 0, 4, ctx->stream));
    size_t blocks = std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 16);
    fr_check_canonical_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const uint4*)d_in, n, flag);
    ctx->launches += 1;
    OZK_CUDA(cudaMemcpyAsync((char*)ctx->pinned + 2048, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    return OZK_OK;
}
static int check_canonical_end(ozk_ctx* ctx, const char* who) {
    OZK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*(const uint32_t*)((const char*)ctx->pinned + 2048) & 4u) {
        set_error("%s: an input element is not reduced mod r", who);
        return OZK_ERR_DOMAIN;
    }
    return OZK_OK;
}

}  // namespace ozk

using namespace ozk;

extern "C" {

int ozk_ntt_fr_dev(ozk_ctx* ctx, const void* d_in, void* d_out, size_t n, const uint8_t omega[32]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(d_in && d_out && omega, "ozk_ntt_fr_dev: null pointer");
    int log_n = ilog2_exact(n);
    OZK_ARG(log_n >= 0 && log_n <= 28, "ozk_ntt_fr_dev: n must be a power of two <= 2^28 (Fr has 2-adicity 28)");
    return ntt_run(ctx, d_in, d_out, log_n, omega);
}

int ozk_ntt_fr_ex_dev(ozk_ctx* ctx, const void* d_in, void* d_out, size_t n, const uint8_t omega[32],
                      const uint8_t* pre_coset, const uint8_t* post_scale, const uint8_t* post_coset) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(d_in && d_out && omega, "ozk_ntt_fr_ex_dev: null pointer");
    int log_n = ilog2_exact(n);
    OZK_ARG(log_n >= 0 && log_n <= 28, "ozk_ntt_fr_ex_dev: n must be a power of two <= 2^28");
    const void* src = d_in;
    if (pre_coset) {
        OZK_TRY(scale_powers(ctx, d_in, d_out, n, nullptr, pre_coset));
        src = d_out;
    }
    OZK_TRY(ntt_run(ctx, src, d_out, log_n, omega));
    if (post_scale || post_coset) OZK_TRY(scale_powers(ctx, d_out, d_out, n, post_scale, post_coset));
    return OZK_OK;
}

int ozk_ntt_fr(ozk_ctx* ctx, uint8_t* data, size_t n, const uint8_t omega[32]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(data && omega, "ozk_ntt_fr: null pointer");
    int log_n = ilog2_exact(n);
    OZK_ARG(log_n >= 0 && log_n <= 28, "ozk_ntt_fr: n must be a power of two <= 2^28 (Fr has 2-adicity 28)");
    const size_t bytes = n * 32;
    OZK_TRY(ctx->io_a.reserve(bytes, ctx->stream));
    OZK_TRY(upload_any(ctx, ctx->io_a.p, data, bytes, ctx->stream));
    OZK_TRY(check_canonical_begin(ctx, ctx->io_a.p, n));
    OZK_TRY(ntt_run(ctx, ctx->io_a.p, ctx->io_a.p, log_n, omega));
    OZK_TRY(download_any(ctx, data, ctx->io_a.p, bytes, ctx->stream));
    return check_canonical_end(ctx, "ozk_ntt_fr");
}

int ozk_fr_scale_dev(ozk_ctx* ctx, const void* d_a, void* d_out, size_t n, const uint8_t b[32]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(d_a && d_out && b, "ozk_fr_scale_dev: null pointer");
    if (n == 0) return OZK_OK;
    return scale_powers(ctx, d_a, d_out, n, b, nullptr);
}

int ozk_fr_mul_sub_dev(ozk_ctx* ctx, const void* d_a, const void* d_b, const void* d_c, void* d_out, size_t n) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(n == 0 || (d_a && d_b && d_out), "ozk_fr_mul_sub_dev: null pointer");
    if (n == 0) return OZK_OK;
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    fr_mul_sub_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((const uint4*)d_a, (const uint4*)d_b, (const uint4*)d_c, (uint4*)d_out, n);
    ctx->launches += 1;
    OZK_CUDA(cudaGetLastError());
    return OZK_OK;
}

int ozk_fr_dft_small_dev(ozk_ctx* ctx, const void* d_in, void* d_out, size_t groups, size_t len, const uint8_t omega_g[32]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(d_in && d_out && omega_g && d_in != d_out, "ozk_fr_dft_small_dev: null or aliased pointers");
    int q = ilog2_exact(groups);
    OZK_ARG(q >= 0 && q <= 3, "ozk_fr_dft_small_dev: groups must be 1, 2, 4 or 8");
    if (len == 0) return OZK_OK;
    if (q == 0) {
        OZK_CUDA(cudaMemcpyAsync(d_out, d_in, len * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        return OZK_OK;
    }
    NttPlan* p;
    OZK_TRY(ntt_get_plan(ctx, q, omega_g, &p));       // validates omega_g and holds omega_g^t, t < groups / 2
    const unsigned grid = (unsigned)((len + 255) / 256);
    if (q == 1) fr_dft_small_kernel<1><<<grid, 256, 0, ctx->stream>>>((const uint4*)d_in, (uint4*)d_out, len, p->wsub[0]);
    else if (q == 2) fr_dft_small_kernel<2><<<grid, 256, 0, ctx->stream>>>((const uint4*)d_in, (uint4*)d_out, len, p->wsub[0]);
    else fr_dft_small_kernel<3><<<grid, 256, 0, ctx->stream>>>((const uint4*)d_in, (uint4*)d_out, len, p->wsub[0]);
    ctx->launches += 1;
    OZK_CUDA(cudaGetLastError());
    return OZK_OK;
}

int ozk_fr_dft_small_scatter_dev(ozk_ctx* ctx, const void* d_in, void* const* peer_out, size_t groups, size_t rank, size_t len,
                                 const uint8_t omega_g[32], const uint8_t omega_n[32]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(d_in && peer_out && omega_g && omega_n, "ozk_fr_dft_small_scatter_dev: null pointer");
    const int q = ilog2_exact(groups);
    OZK_ARG(q >= 1 && q <= 3, "ozk_fr_dft_small_scatter_dev: groups must be 2, 4 or 8");
    OZK_ARG(rank < groups, "ozk_fr_dft_small_scatter_dev: rank out of range");
    const int log_len = ilog2_exact(len);
    OZK_ARG(log_len >= 0 && log_len + 2 * q <= 28, "ozk_fr_dft_small_scatter_dev: len must be a power of two, groups^2 * len <= 2^28");
    for (size_t r = 0; r < groups; r++) OZK_ARG(peer_out[r] != nullptr, "ozk_fr_dft_small_scatter_dev: null peer buffer");
    NttPlan* p;
    OZK_TRY(ntt_get_plan(ctx, q, omega_g, &p));       // validates omega_g and holds omega_g^t, t < groups / 2
    NttPlan* tw = nullptr;
    OZK_TRY(get_pow_table(ctx, omega_n, log_len + 2 * q, &tw));
    PeerPtrs pp;
    memset(&pp, 0, sizeof pp);
    for (size_t r = 0; r < groups; r++) pp.p[r] = (uint4*)peer_out[r];
    const unsigned grid = (unsigned)((len + 255) / 256);
    if (q == 1) fr_dft_small_scatter_kernel<1><<<grid, 256, 0, ctx->stream>>>((const uint4*)d_in, len, p->wsub[0], tw->tlo, tw->thi, (uint32_t)tw->H, (uint32_t)rank, pp);
    else if (q == 2) fr_dft_small_scatter_kernel<2><<<grid, 256, 0, ctx->stream>>>((const uint4*)d_in, len, p->wsub[0], tw->tlo, tw->thi, (uint32_t)tw->H, (uint32_t)rank, pp);
    else fr_dft_small_scatter_kernel<3><<<grid, 256, 0, ctx->stream>>>((const uint4*)d_in, len, p->wsub[0], tw->tlo, tw->thi, (uint32_t)tw->H, (uint32_t)rank, pp);
    ctx->launches += 1;
    OZK_CUDA(cudaGetLastError());
    return OZK_OK;
}

int ozk_fr_scale_powers_dev(ozk_ctx* ctx, const void* d_a, void* d_out, size_t n, const uint8_t* scale, const uint8_t* coset,
                            uint64_t first_index) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(n == 0 || (d_a && d_out), "ozk_fr_scale_powers_dev: null pointer");
    if (n == 0) return OZK_OK;
    return scale_powers(ctx, d_a, d_out, n, scale, coset, first_index);
}

/* ---- multi-GPU four-step, fused exchange ---------------------------------------------------------------------------------- */
int ozk_peer_alloc(ozk_ctx* ctx, size_t bytes, void** d_ptr, uint8_t handle[64]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(d_ptr && handle && bytes > 0, "ozk_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    OZK_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("ozk_peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        cudaGetLastError();
        return OZK_ERR_CUDA;
    }
    memcpy(handle, &h, 64);
    *d_ptr = p;
    return OZK_OK;
}
int ozk_peer_open(ozk_ctx* ctx, const uint8_t handle[64], void** d_ptr) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(d_ptr && handle, "ozk_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    OZK_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return OZK_OK;
}
int ozk_peer_close(ozk_ctx* ctx, void* d_ptr) {
    OZK_TRY(ctx_enter(ctx));
    OZK_CUDA(cudaStreamSynchronize(ctx->stream));
    OZK_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return OZK_OK;
}
int ozk_peer_free(ozk_ctx* ctx, void* d_ptr) {
    OZK_TRY(ctx_enter(ctx));
    OZK_CUDA(cudaStreamSynchronize(ctx->stream));
    OZK_CUDA(cudaFree(d_ptr));
    return OZK_OK;
}

int ozk_ntt_fr_scatter_dev(ozk_ctx* ctx, const void* d_in, void* const* peer_out, size_t groups, size_t rank, size_t n_local,
                           const uint8_t omega_local[32], const uint8_t twiddle_base[32]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(d_in && peer_out && omega_local && twiddle_base, "ozk_ntt_fr_scatter_dev: null pointer");
    OZK_ARG(groups == 2 || groups == 4 || groups == 8, "ozk_ntt_fr_scatter_dev: groups must be 2, 4 or 8");
    OZK_ARG(rank < groups, "ozk_ntt_fr_scatter_dev: rank out of range");
    const int log_m = ilog2_exact(n_local);
    OZK_ARG(log_m >= 3 && log_m <= 28, "ozk_ntt_fr_scatter_dev: n_local must be a power of two in [8, 2^28]");
    for (size_t r = 0; r < groups; r++) OZK_ARG(peer_out[r] != nullptr, "ozk_ntt_fr_scatter_dev: null peer buffer");
    int log_g = 0;
    while (((size_t)1 << log_g) < groups) log_g++;
    NttPlan* tw = nullptr;
    OZK_TRY(get_pow_table(ctx, twiddle_base, log_m, &tw));
    ScatterDesc sc;
    sc.npeers = (uint32_t)groups;
    sc.my_rank = (uint32_t)rank;
    sc.log_chunk = (uint32_t)(log_m - log_g);
    sc.pH = (uint32_t)tw->H;
    sc.ptlo = tw->tlo;
    sc.pthi = tw->thi;
    for (size_t r = 0; r < groups; r++) sc.peer[r] = (uint4*)peer_out[r];
    return ntt_run(ctx, d_in, nullptr, log_m, omega_local, &sc);
}

static int spmv_run(ozk_ctx* ctx, const void* d_row_ptr, const void* d_col, const void* d_coeff, const void* d_z, size_t z_len, size_t rows,
                    void* d_out, const char* who) {
    if (rows == 0) return OZK_OK;
    cudaStream_t st = ctx->stream;
    OZK_TRY(ctx->io_out.reserve(512 + (size_t)kSpmvLongCap * 4, st));
    uint32_t* long_count = (uint32_t*)((char*)ctx->io_out.p + 384);
    uint32_t* flag = long_count + 1;
    uint32_t* long_rows = (uint32_t*)((char*)ctx->io_out.p + 512);
    OZK_CUDA(cudaMemsetAsync(long_count, 0, 8, st));
    fr_spmv_short<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>((const uint32_t*)d_row_ptr, (const uint32_t*)d_col, (const uint4*)d_coeff,
                                                                  (const uint4*)d_z, z_len, rows, (uint4*)d_out, long_rows, long_count, flag);
    OZK_TRY(ctx->work.reserve((size_t)kSpmvMaxSegs * 32, st));            // partial sums of the long rows (the NTT scratch, same stream)
    fr_spmv_long<<<ctx->sm_count * 4, 256, 0, st>>>((const uint32_t*)d_row_ptr, (const uint32_t*)d_col, (const uint4*)d_coeff, (const uint4*)d_z, z_len,
                                                    (uint4*)ctx->work.p, long_rows, long_count, flag);
    fr_spmv_long_finish<<<(kSpmvLongCap + 127) / 128, 128, 0, st>>>((const uint32_t*)d_row_ptr, (const uint4*)ctx->work.p, (uint4*)d_out, long_rows,
                                                                   long_count);
    ctx->launches += 3;
    OZK_CUDA(cudaGetLastError());
    uint32_t* pin = (uint32_t*)ctx->pinned;
    OZK_CUDA(cudaMemcpyAsync(pin, long_count, 8, cudaMemcpyDeviceToHost, st));
    OZK_CUDA(cudaStreamSynchronize(st));
    if (pin[0] > kSpmvLongCap) {
        set_error("%s: more than %u rows with over %u terms", who, kSpmvLongCap, kSpmvShort);
        return OZK_ERR_ARG;
    }
    if (pin[1] & 1u) {
        set_error("%s: a column index is outside the assignment (>= %zu)", who, z_len);
        return OZK_ERR_ARG;
    }
    if (pin[1] & 2u) {
        set_error("%s: the rows with over %u terms hold more than %u x %u terms in total", who, kSpmvShort, kSpmvMaxSegs, kSpmvSeg);
        return OZK_ERR_ARG;
    }
    return OZK_OK;
}

int ozk_fr_spmv_dev(ozk_ctx* ctx, const void* d_row_ptr, const void* d_col, const void* d_coeff, const void* d_z, size_t rows, void* d_out) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(rows == 0 || (d_row_ptr && d_col && d_coeff && d_z && d_out), "ozk_fr_spmv_dev: null pointer");
    OZK_ARG(rows < ((size_t)1 << 31), "ozk_fr_spmv_dev: too many rows");
    return spmv_run(ctx, d_row_ptr, d_col, d_coeff, d_z, ~(size_t)0, rows, d_out, "ozk_fr_spmv_dev");
}

int ozk_fr_spmv_ex_dev(ozk_ctx* ctx, const void* d_row_ptr, const void* d_col, const void* d_coeff, const void* d_z, size_t z_len, size_t rows,
                       void* d_out) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(rows == 0 || (d_row_ptr && d_col && d_z && d_out), "ozk_fr_spmv_ex_dev: null pointer");
    OZK_ARG(rows < ((size_t)1 << 31), "ozk_fr_spmv_ex_dev: too many rows");
    return spmv_run(ctx, d_row_ptr, d_col, d_coeff, d_z, z_len, rows, d_out, "ozk_fr_spmv_ex_dev");
}

int ozk_fr_lincomb_dev(ozk_ctx* ctx, void* d_out, size_t n, const void* d_a, const uint8_t ca[32], const void* d_b, const uint8_t cb[32],
                       const void* d_c, const uint8_t cc[32]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(n == 0 || (d_out && d_a && ca && (!d_b || cb) && (!d_c || cc)), "ozk_fr_lincomb_dev: null pointer");
    if (n == 0) return OZK_OK;
    Fr fa, fb, fc;
    memset(&fb, 0, sizeof fb);
    memset(&fc, 0, sizeof fc);
    if (!fr_bytes_canonical(ca) || (d_b && !fr_bytes_canonical(cb)) || (d_c && !fr_bytes_canonical(cc))) {
        set_error("ozk_fr_lincomb_dev: a constant is not reduced mod r");
        return OZK_ERR_DOMAIN;
    }
    fr_from_bytes(fa, ca);
    if (d_b) fr_from_bytes(fb, cb);
    if (d_c) fr_from_bytes(fc, cc);
    size_t blocks = std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 16);
    fr_lincomb_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((uint4*)d_out, n, (const uint4*)d_a, fa, (const uint4*)d_b, fb, (const uint4*)d_c, fc);
    ctx->launches += 1;
    OZK_CUDA(cudaGetLastError());
    return OZK_OK;
}

int ozk_fr_lagrange_dev(ozk_ctx* ctx, void* d_out, size_t m, const uint8_t t[32], const uint8_t omega[32]) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(d_out && t && omega, "ozk_fr_lagrange_dev: null pointer");
    OZK_ARG(m >= 1 && m <= ((size_t)1 << 28) && (m & (m - 1)) == 0, "ozk_fr_lagrange_dev: m must be a power of two <= 2^28");
    if (!fr_bytes_canonical(t) || !fr_bytes_canonical(omega)) {
        set_error("ozk_fr_lagrange_dev: t or omega is not reduced mod r");
        return OZK_ERR_DOMAIN;
    }
    Fr tt, ww;
    fr_from_bytes(tt, t);
    fr_from_bytes(ww, omega);
    uint32_t log_m = 0;
    while (((size_t)1 << log_m) < m) log_m++;
    cudaStream_t st = ctx->stream;
    OZK_TRY(ctx->io_out.reserve(512, st));
    Fr* consts = (Fr*)ctx->io_out.p;
    uint32_t* hit = (uint32_t*)((char*)ctx->io_out.p + 256);
    OZK_CUDA(cudaMemsetAsync(hit, 0, 8, st));
    // omega must generate S: omega^(m/2) == -1 (m == 1: omega == 1), checked on the device like the NTT's
    ntt_check_omega<<<1, 32, 0, st>>>(hit + 2, ww, (uint32_t)(m / 2));
    fr_lagrange_setup<<<1, 32, 0, st>>>(consts, tt, log_m);
    const size_t threads = (m + kLagBatch - 1) / kLagBatch;
    fr_lagrange_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>((uint4*)d_out, m, tt, ww, consts, hit);
    size_t blocks = std::min<size_t>((m + 255) / 256, (size_t)ctx->sm_count * 8);
    fr_lagrange_unit<<<(unsigned)blocks, 256, 0, st>>>((uint4*)d_out, m, hit);
    ctx->launches += 4;
    OZK_CUDA(cudaGetLastError());
    uint32_t* pin = (uint32_t*)ctx->pinned;
    OZK_CUDA(cudaMemcpyAsync(pin, hit, 12, cudaMemcpyDeviceToHost, st));
    OZK_CUDA(cudaStreamSynchronize(st));
    if (pin[2] != 1u) {
        set_error("ozk_fr_lagrange_dev: omega is not a primitive m-th root of unity");
        return OZK_ERR_DOMAIN;
    }
    return OZK_OK;
}

int ozk_fr_scale(ozk_ctx* ctx, const uint8_t* a, size_t n, const uint8_t b[32], uint8_t* out) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(b && (n == 0 || (a && out)), "ozk_fr_scale: null pointer");
    if (n == 0) return OZK_OK;
    const size_t bytes = n * 32;
    OZK_TRY(ctx->io_a.reserve(bytes, ctx->stream));
    OZK_TRY(upload_any(ctx, ctx->io_a.p, a, bytes, ctx->stream));
    OZK_TRY(check_canonical_begin(ctx, ctx->io_a.p, n));
    OZK_TRY(scale_powers(ctx, ctx->io_a.p, ctx->io_a.p, n, b, nullptr));
    OZK_TRY(download_any(ctx, out, ctx->io_a.p, bytes, ctx->stream));
    return check_canonical_end(ctx, "ozk_fr_scale");
}

}  // extern "C"
