// Fixed-base batch MSM driver and C ABI: out[i] = (s_i mod 2^(outerc*w)) * B.
//
// Replaces Java_algebra_msm_FixedBaseMSM_batchMSMNativeHelper (algebra_msm_FixedBaseMSM.cu:1276-1384) and, through two
// calls, ...doubleBatchMSMNativeHelper (:1395-1491).  See fixed_impl.cuh for the kernels.  The device window table is
// cached per (group, base, table window) on the context, so the four batchMSM calls of SerialSetup on the same
// generator (SerialSetup.java:123-164) build it once; the reference rebuilds it on every call (:1030-1041).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "fixed_impl.cuh"

namespace ozk {

struct FixedTable {
    void* table_aff = nullptr;
    uint32_t t = 0, nwin = 0;
};

void fixed_free_tables(ozk_ctx* ctx) {
    for (auto& kv : ctx->fixed_tables) {
        if (kv.second->table_aff) cudaFree(kv.second->table_aff);
        delete kv.second;
    }
    ctx->fixed_tables.clear();
}

// Table window t (signed digits, 2^(t-1) affine entries per window, ceil(255 / t) windows), in units of one mixed addition:
//   walk  : windows * n             (measured 2^24 scalars: 41.5 ms at t = 16, 34.5 at 20, 32.2 at 22: the walk is bound by
//                                    the additions, not by the 64-byte gathers, also when the table is 1.6 GB of HBM)
//   build : 2.7 per entry (one mixed addition and a share of a batched inversion), divided by an assumed reuse of 2 --
//           the table is cached per (group, base, t) and SerialSetup.generate walks the same G1 table five times
//           (SerialSetup.java:123-164), every later setup or proof key too
// capped at t = 22 (12 windows x 2^21 entries = 1.6 GB for G1) and at 4 GB of table.
static uint32_t fixed_choose_t(size_t n, size_t affine_bytes) {
    if (const char* e = getenv("OZK_FIXED_T")) {
        int t = atoi(e);
        if (t >= 4 && t <= 22) return (uint32_t)t;
    }
    uint32_t best = 4;
    double best_cost = 1e300;
    for (uint32_t t = 4; t <= 22; t++) {
        const double nwin = (255 + t - 1) / t;
        const double entries = (double)(1u << (t - 1));
        if (nwin * entries * (double)affine_bytes > 4.0e9) break;
        const double cost = nwin * ((double)n + entries * (2.7 / 2.0));
        if (cost < best_cost) {
            best_cost = cost;
            best = t;
        }
    }
    return best;
}

enum { FB_SMALL = 0, FB_TABLE_TMP, FB_OUT_XYZZ, FB_SCALARS };

static int fixed_get_table(ozk_ctx* ctx, const FixedLaunch& L, int tag, const uint8_t* base, uint32_t t, FixedTable** out) {
    std::string key;
    key.push_back((char)tag);
    key.push_back((char)t);
    key.append((const char*)base, L.jac_bytes);
    auto it = ctx->fixed_tables.find(key);
    if (it != ctx->fixed_tables.end()) {
        *out = it->second;
        return OZK_OK;
    }
    cudaStream_t st = ctx->stream;
    const uint32_t nwin = (255 + t - 1) / t;
    const size_t entries = ((size_t)1 << (t - 1)) * nwin;
    // small block: [flag 256 B][base][pow_xyzz 256][pow_aff 256]
    const size_t nsub = (size_t)fixed_sub_count(t) * nwin;
    const size_t small = 256 + 512 + kFixedPowers * (L.xyzz_bytes + L.affine_bytes) + nsub * (L.xyzz_bytes + L.affine_bytes);
    OZK_TRY(ctx->fb[FB_SMALL].reserve(small, st));
    char* sp = (char*)ctx->fb[FB_SMALL].p;
    uint32_t* flag = (uint32_t*)sp;
    void* d_base = sp + 256;
    void* pow_xyzz = sp + 256 + 512;
    void* pow_aff = (char*)pow_xyzz + kFixedPowers * L.xyzz_bytes;
    void* sub = (char*)pow_aff + kFixedPowers * L.affine_bytes;
    OZK_CUDA(cudaMemsetAsync(flag, 0, 16, st));
    OZK_CUDA(cudaMemcpyAsync(d_base, base, L.jac_bytes, cudaMemcpyHostToDevice, st));
    OZK_TRY(ctx->fb[FB_TABLE_TMP].reserve(entries * L.xyzz_bytes, st));
    FixedTable* ft = new FixedTable();
    ft->t = t;
    ft->nwin = nwin;
    if (cudaMalloc(&ft->table_aff, entries * L.affine_bytes) != cudaSuccess) {
        delete ft;
        set_error("fixed-base: cudaMalloc of the window table failed");
        cudaGetLastError();
        return OZK_ERR_CUDA;
    }
    int rc = 0;
    rc |= L.powers(st, d_base, pow_xyzz, flag);
    rc |= L.to_affine(st, pow_xyzz, pow_aff, kFixedPowers);
    rc |= L.table(st, pow_aff, sub, ctx->fb[FB_TABLE_TMP].p, t, nwin);
    rc |= L.to_affine(st, ctx->fb[FB_TABLE_TMP].p, ft->table_aff, entries);
    ctx->launches += 6;
    uint32_t* hflag = (uint32_t*)ctx->pinned;
    cudaError_t e = cudaMemcpyAsync(hflag, flag, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (rc || e != cudaSuccess) {
        cudaFree(ft->table_aff);
        delete ft;
        set_error("fixed-base: table build failed: %s", cudaGetErrorString(e));
        return OZK_ERR_CUDA;
    }
    if (hflag[0] & 1u) {
        cudaFree(ft->table_aff);
        delete ft;
        set_error("fixed-base: a base coordinate is not reduced mod p");
        return OZK_ERR_DOMAIN;
    }
    // keep the cache small: tables are tens of MiB
    if (ctx->fixed_tables.size() >= 8) fixed_free_tables(ctx);
    ctx->fixed_tables[key] = ft;
    *out = ft;
    return OZK_OK;
}

static int fixed_run(ozk_ctx* ctx, const FixedLaunch& L, int tag, const uint8_t* base, const void* d_scalars, size_t n, int outerc,
                     int window, void* d_out, unsigned flags = 0) {
    OZK_ARG(outerc >= 0 && window >= 1 && window <= 256, "fixed-base: outerc must be >= 0 and 1 <= windowSize <= 256");
    if (n == 0) return OZK_OK;
    long long bits_ll = (long long)outerc * window;
    const uint32_t bits = (uint32_t)std::min<long long>(bits_ll, 256);
    const uint32_t t = fixed_choose_t(n, L.affine_bytes);
    FixedTable* ft;
    OZK_TRY(fixed_get_table(ctx, L, tag, base, t, &ft));
    cudaStream_t st = ctx->stream;
    OZK_TRY(ctx->fb[FB_OUT_XYZZ].reserve(n * L.xyzz_bytes, st));
    uint32_t* flag = (uint32_t*)ctx->fb[FB_SMALL].p;
    OZK_CUDA(cudaMemsetAsync(flag, 0, 16, st));
    if (L.walk(st, d_scalars, n, ft->table_aff, ft->t, ft->nwin, bits, ctx->fb[FB_OUT_XYZZ].p, flag) ||
        ((flags & OZK_FIXED_KEEP_Z) ? L.to_wire_raw : L.to_wire)(st, ctx->fb[FB_OUT_XYZZ].p, d_out, n)) {
        set_error("fixed-base: kernel launch failed");
        return OZK_ERR_CUDA;
    }
    ctx->launches += 2;
    uint32_t* hflag = (uint32_t*)ctx->pinned;
    OZK_CUDA(cudaMemcpyAsync(hflag, flag, 4, cudaMemcpyDeviceToHost, st));
    OZK_CUDA(cudaStreamSynchronize(st));
    if (hflag[0] & 2u) {
        set_error("fixed-base: a scalar is not reduced mod r");
        return OZK_ERR_DOMAIN;
    }
    return OZK_OK;
}

static int fixed_run_host(ozk_ctx* ctx, const FixedLaunch& L, int tag, const uint8_t* base, const uint8_t* scalars, size_t n, int outerc,
                          int window, uint8_t* out) {
    if (n == 0) return OZK_OK;
    OZK_TRY(ctx->fb[FB_SCALARS].reserve(n * 32, ctx->stream));
    OZK_TRY(ctx->io_b.reserve(n * L.jac_bytes, ctx->stream));
    // pageable buffers (JNI byte[]) go through the bounce-buffer stager (stage.cu)
    OZK_TRY(upload_any(ctx, ctx->fb[FB_SCALARS].p, scalars, n * 32, ctx->stream));
    OZK_TRY(fixed_run(ctx, L, tag, base, ctx->fb[FB_SCALARS].p, n, outerc, window, ctx->io_b.p));
    return download_any(ctx, out, ctx->io_b.p, n * L.jac_bytes, ctx->stream);
}

}  // namespace ozk

using namespace ozk;

extern "C" {

int ozk_fixed_g1_dev(ozk_ctx* ctx, const uint8_t base[96], const void* d_scalars, size_t n, int outerc, int window, void* d_out) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(base && (n == 0 || (d_scalars && d_out)), "ozk_fixed_g1_dev: null pointer");
    return fixed_run(ctx, kFixedG1, 1, base, d_scalars, n, outerc, window, d_out);
}
int ozk_fixed_g2_dev(ozk_ctx* ctx, const uint8_t base[192], const void* d_scalars, size_t n, int outerc, int window, void* d_out) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(base && (n == 0 || (d_scalars && d_out)), "ozk_fixed_g2_dev: null pointer");
    return fixed_run(ctx, kFixedG2, 2, base, d_scalars, n, outerc, window, d_out);
}
int ozk_fixed_g1_ex_dev(ozk_ctx* ctx, const uint8_t base[96], const void* d_scalars, size_t n, int outerc, int window, unsigned flags, void* d_out) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(base && (n == 0 || (d_scalars && d_out)), "ozk_fixed_g1_ex_dev: null pointer");
    OZK_ARG((flags & ~(unsigned)OZK_FIXED_KEEP_Z) == 0, "ozk_fixed_g1_ex_dev: unknown flag");
    return fixed_run(ctx, kFixedG1, 1, base, d_scalars, n, outerc, window, d_out, flags);
}
int ozk_fixed_g2_ex_dev(ozk_ctx* ctx, const uint8_t base[192], const void* d_scalars, size_t n, int outerc, int window, unsigned flags, void* d_out) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(base && (n == 0 || (d_scalars && d_out)), "ozk_fixed_g2_ex_dev: null pointer");
    OZK_ARG((flags & ~(unsigned)OZK_FIXED_KEEP_Z) == 0, "ozk_fixed_g2_ex_dev: unknown flag");
    return fixed_run(ctx, kFixedG2, 2, base, d_scalars, n, outerc, window, d_out, flags);
}
int ozk_fixed_g1(ozk_ctx* ctx, const uint8_t base[96], const uint8_t* scalars, size_t n, int outerc, int window, uint8_t* out) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(base && (n == 0 || (scalars && out)), "ozk_fixed_g1: null pointer");
    return fixed_run_host(ctx, kFixedG1, 1, base, scalars, n, outerc, window, out);
}
int ozk_fixed_g2(ozk_ctx* ctx, const uint8_t base[192], const uint8_t* scalars, size_t n, int outerc, int window, uint8_t* out) {
    OZK_TRY(ctx_enter(ctx));
    OZK_ARG(base && (n == 0 || (scalars && out)), "ozk_fixed_g2: null pointer");
    return fixed_run_host(ctx, kFixedG2, 2, base, scalars, n, outerc, window, out);
}

}  // extern "C"
