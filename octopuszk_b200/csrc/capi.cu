// Context management, error reporting and the integer-pipe microbenchmarks of liboctozk.so.
#include <cstdlib>

#include "common.h"
#include "fp256.cuh"

namespace ozk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int ctx_enter(ozk_ctx* ctx) {
    if (!ctx) {
        set_error("null context");
        return OZK_ERR_ARG;
    }
    OZK_CUDA(cudaSetDevice(ctx->device));
    return OZK_OK;
}

void ntt_free_plans(ozk_ctx* ctx);
void fixed_free_tables(ozk_ctx* ctx);

// ---- microbenchmarks ---------------------------------------------------------------------------------------
// 8 independent 64-bit accumulators per thread, each fed by mad.lo.cc/madc.hi pairs (what ptxas turns into
// IMAD.WIDE.U32): the same instruction the Montgomery rounds issue.
__global__ void __launch_bounds__(256) imad_peak_kernel(uint32_t* out, uint32_t seed, int iters) {
    using namespace ptx;
    uint32_t b = seed * 3 + blockIdx.x + 0x9e3779b9u;
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        lo[k] = threadIdx.x * 8 + k + 1;
        hi[k] = seed + k;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                // (hi:lo)[k] += lo[k] * b, carry into the next pair: the same lo/hi chain the Montgomery rounds use.
                // The multiplier is the accumulator's own low word: a true dependency, nothing to hoist or share.
                uint32_t m0 = lo[k], m1 = lo[k + 1];
                lo[k] = mad_lo_cc(m0, b, lo[k]);
                hi[k] = madc_hi_cc(m0, b, hi[k]);
                lo[k + 1] = madc_lo_cc(m1, b, lo[k + 1]);
                hi[k + 1] = madc_hi(m1, b, hi[k + 1]);
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= lo[k] ^ hi[k];
    if (s == 0x12345678u) out[0] = s;
}

__global__ void __launch_bounds__(256) modmul_peak_kernel(uint32_t* out, uint32_t seed, int iters) {
    Fr x = Fr::one(), y = Fr::rr();
    x.v[0] += threadIdx.x + seed;
    y.v[1] ^= blockIdx.x;
    Fr z = x;
    for (int it = 0; it < iters; it++) {
        x = Fr::mul(x, y);
        z = Fr::mul(z, x);
    }
    if (z.v[0] == 0x12345678u && x.v[3] == 7) out[0] = z.v[1];
}

// Issue-rate probes for the other pipes a 256-bit multiplier could use (planning data for mixed-pipe arithmetic):
//   which 1: DFMA chains (fp64 pipe)   2: DFMA and IMAD.WIDE chains interleaved in the same thread
//   which 3: IADD3 carry chains (alu)  4: 32-bit IMAD (low half only) chains
__global__ void __launch_bounds__(256) pipe_probe_kernel(uint32_t* out, uint32_t seed, int iters, int which) {
    using namespace ptx;
    double d[8];
    uint32_t lo[8], hi[8];
    const double m = 1.0 + 1e-9 * (seed & 7);
    uint32_t b = seed * 3 + blockIdx.x + 0x9e3779b9u;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        d[k] = 1.0 + threadIdx.x * 1e-3 + k;
        lo[k] = threadIdx.x * 8 + k + 1;
        hi[k] = seed + k;
    }
    if (which == 1) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int k = 0; k < 8; k++) d[k] = fma(d[k], m, 0.5);
        }
    } else if (which == 2) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 8; r++) {
#pragma unroll
                for (int k = 0; k < 8; k++) d[k] = fma(d[k], m, 0.5);
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    uint32_t m0 = lo[k], m1 = lo[k + 1];
                    lo[k] = mad_lo_cc(m0, b, lo[k]);
                    hi[k] = madc_hi_cc(m0, b, hi[k]);
                    lo[k + 1] = madc_lo_cc(m1, b, lo[k + 1]);
                    hi[k + 1] = madc_hi(m1, b, hi[k + 1]);
                }
            }
        }
    } else if (which == 3) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 8; r++) {
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    lo[k] = add_cc(lo[k], hi[k + 1]);
                    hi[k] = addc_cc(hi[k], lo[k + 1]);
                    lo[k + 1] = addc_cc(lo[k + 1], hi[k]);
                    hi[k + 1] = addc(hi[k + 1], lo[k]);
                }
            }
        }
    } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int k = 0; k < 8; k++) lo[k] = lo[k] * b + hi[k];
        }
    }
    uint32_t s = 0;
    double ds = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        s ^= lo[k] ^ hi[k];
        ds += d[k];
    }
    if (s == 0x12345678u && ds == 3.25) out[0] = s;
}

}  // namespace ozk

using namespace ozk;

extern "C" {

const char* ozk_last_error(void) { return g_err; }
const char* ozk_version(void) { return "octozk-b200 0.1 (sm_100a)"; }

int ozk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        set_error("cudaGetDeviceCount failed: no usable CUDA device (there is no CPU fallback)");
        cudaGetLastError();
        return 0;
    }
    return n;
}

void ozk_ctx_destroy(ozk_ctx* c);

int ozk_ctx_create(int device, ozk_ctx** out) {
    OZK_ARG(out != nullptr, "ozk_ctx_create: out is null");
    int n = 0;
    OZK_CUDA(cudaGetDeviceCount(&n));
    OZK_ARG(device >= 0 && device < n, "ozk_ctx_create: device index out of range");
    OZK_CUDA(cudaSetDevice(device));
    if (const char* e = getenv("OZK_L2_FETCH")) {
        // experiment: L2 fill granularity for the 64-byte random gathers of the MSM (a device-wide hint: 32, 64 or 128 bytes)
        const int g = atoi(e);
        if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g);
        cudaGetLastError();
    }
    ozk_ctx* c = new ozk_ctx();
    c->device = device;
    // anything that fails below must not leak the half-built context
    auto build = [&]() -> int {
        cudaDeviceProp prop;
        OZK_CUDA(cudaGetDeviceProperties(&prop, device));
        c->sm_count = prop.multiProcessorCount;
        OZK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        OZK_CUDA(cudaEventCreate(&c->ev0));
        OZK_CUDA(cudaEventCreate(&c->ev1));
        for (auto& e : c->evs) OZK_CUDA(cudaEventCreate(&e));
        OZK_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        {
            int lo_prio = 0, hi_prio = 0;
            OZK_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
            OZK_CUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, hi_prio));
        }
        for (auto& e : c->copy_ev) OZK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        OZK_CUDA(cudaMallocHost(&c->pinned, 4096));
        return OZK_OK;
    };
    const int rc = build();
    if (rc != OZK_OK) {
        ozk_ctx_destroy(c);
        return rc;
    }
    *out = c;
    return OZK_OK;
}

void ozk_ctx_destroy(ozk_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    ntt_free_plans(c);
    fixed_free_tables(c);
    stager_free(c);
    c->io_a.release(); c->io_b.release(); c->io_c.release(); c->io_out.release();
    c->work.release();
    for (auto& b : c->msm) b.release();
    for (auto& b : c->fb) b.release();
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (auto& e : c->evs) if (e) cudaEventDestroy(e);
    for (auto& e : c->copy_ev) if (e) cudaEventDestroy(e);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->side_stream) { cudaStreamSynchronize(c->side_stream); cudaStreamDestroy(c->side_stream); }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int ozk_ctx_set_stream(ozk_ctx* c, void* s) {
    OZK_TRY(ctx_enter(c));
    if ((cudaStream_t)s == c->stream) return OZK_OK;
    // Work already enqueued on the old stream stays ordered before anything enqueued on the new one (device-side, no host
    // synchronisation): the new stream waits for an event recorded at the tail of the old one.
    cudaEvent_t ev = c->copy_ev[19];
    OZK_CUDA(cudaEventRecord(ev, c->stream));
    OZK_CUDA(cudaStreamWaitEvent((cudaStream_t)s, ev, 0));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);     // returns at once; released when its work has drained
    c->stream = (cudaStream_t)s;
    c->own_stream = false;
    return OZK_OK;
}

unsigned long long ozk_ctx_launches(ozk_ctx* c) { return c ? c->launches : 0; }

int ozk_ctx_sync(ozk_ctx* c) {
    OZK_TRY(ctx_enter(c));
    OZK_CUDA(cudaStreamSynchronize(c->stream));
    return OZK_OK;
}

static int time_kernel(ozk_ctx* c, void (*k)(uint32_t*, uint32_t, int), int iters, float* ms) {
    OZK_TRY(c->io_out.reserve(256, c->stream));
    int grid = c->sm_count * 8;
    k<<<grid, 256, 0, c->stream>>>((uint32_t*)c->io_out.p, 1u, 16);   // warm-up
    c->launches += 4;
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        OZK_CUDA(cudaEventRecord(c->ev0, c->stream));
        k<<<grid, 256, 0, c->stream>>>((uint32_t*)c->io_out.p, 7u + rep, iters);
        OZK_CUDA(cudaEventRecord(c->ev1, c->stream));
        OZK_CUDA(cudaEventSynchronize(c->ev1));
        float t;
        OZK_CUDA(cudaEventElapsedTime(&t, c->ev0, c->ev1));
        if (t < best) best = t;
    }
    OZK_CUDA(cudaGetLastError());
    *ms = best;
    return OZK_OK;
}

int ozk_imad_peak(ozk_ctx* c, double* out) {
    OZK_TRY(ctx_enter(c));
    OZK_ARG(out != nullptr, "ozk_imad_peak: out is null");
    const int iters = 4096;
    float ms;
    OZK_TRY(time_kernel(c, imad_peak_kernel, iters, &ms));
    double ops = (double)c->sm_count * 8 * 256 * (double)iters * 64.0;
    *out = ops / (ms * 1e-3) / 1e9;
    return OZK_OK;
}

int ozk_pipe_probe(ozk_ctx* c, int which, double* out) {
    OZK_TRY(ctx_enter(c));
    OZK_ARG(out != nullptr && which >= 1 && which <= 4, "ozk_pipe_probe: which must be 1..4");
    OZK_TRY(c->io_out.reserve(256, c->stream));
    const int iters = 2048, grid = c->sm_count * 8;
    pipe_probe_kernel<<<grid, 256, 0, c->stream>>>((uint32_t*)c->io_out.p, 1u, 16, which);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        OZK_CUDA(cudaEventRecord(c->ev0, c->stream));
        pipe_probe_kernel<<<grid, 256, 0, c->stream>>>((uint32_t*)c->io_out.p, 7u + rep, iters, which);
        OZK_CUDA(cudaEventRecord(c->ev1, c->stream));
        OZK_CUDA(cudaEventSynchronize(c->ev1));
        float t;
        OZK_CUDA(cudaEventElapsedTime(&t, c->ev0, c->ev1));
        if (t < best) best = t;
    }
    c->launches += 4;
    OZK_CUDA(cudaGetLastError());
    // operations of the probed kind per iteration and thread: 64 (which 2 counts the 64 DFMA; the 32 IMAD.WIDE pairs ride along)
    *out = (double)c->sm_count * 8 * 256 * (double)iters * 64.0 / (best * 1e-3) / 1e9;
    return OZK_OK;
}

int ozk_modmul_peak(ozk_ctx* c, double* out) {
    OZK_TRY(ctx_enter(c));
    OZK_ARG(out != nullptr, "ozk_modmul_peak: out is null");
    const int iters = 2048;
    float ms;
    OZK_TRY(time_kernel(c, modmul_peak_kernel, iters, &ms));
    double ops = (double)c->sm_count * 8 * 256 * (double)iters * 2.0;
    *out = ops / (ms * 1e-3) / 1e9;
    return OZK_OK;
}

}  // extern "C"
