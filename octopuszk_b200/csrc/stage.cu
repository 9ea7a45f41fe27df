// Host -> device upload of PAGEABLE memory at close to PCIe speed.
//
// The JNI shims hand liboctozk pointers into JVM heap arrays (GetPrimitiveArrayCritical), i.e. pageable memory.
// cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread at ~8 GB/s (measured: 2^23 G1 pairs,
// 1.07 GB, 110 ms), five times slower than the link.  Here kStageThreads host threads copy interleaved chunks into their
// own pinned bounce buffers and enqueue the DMA of each chunk on their own stream, so the page-touching memcpy runs in
// parallel with itself and with the DMA of the previous chunks.  The reference has the same problem one level up (it
// marshals BigIntegers element by element, VariableBaseMSM.java:217-237); the byte[] it finally passes is what arrives here.
#include <algorithm>
#include <cstring>
#include <thread>

#include "common.h"

namespace ozk {

struct Stager {
    static constexpr int kThreads = 8;          // upper bound; `active` of them are used (OZK_STAGE_THREADS, default 4)
    static constexpr int kBufs = 2;
    int active = 4;
    static constexpr size_t kChunk = (size_t)4 << 20;
    void* pinned[kThreads][kBufs] = {};
    cudaStream_t st[kThreads] = {};
    cudaEvent_t ev[kThreads][kBufs] = {};
    cudaEvent_t done[kThreads] = {};
    cudaEvent_t start = nullptr;
    bool ok = false;
};

void stager_free(ozk_ctx* ctx) {
    Stager* s = ctx->stager;
    if (!s) return;
    for (int t = 0; t < Stager::kThreads; t++) {
        if (s->st[t]) { cudaStreamSynchronize(s->st[t]); cudaStreamDestroy(s->st[t]); }
        for (int b = 0; b < Stager::kBufs; b++) {
            if (s->pinned[t][b]) cudaFreeHost(s->pinned[t][b]);
            if (s->ev[t][b]) cudaEventDestroy(s->ev[t][b]);
        }
        if (s->done[t]) cudaEventDestroy(s->done[t]);
    }
    if (s->start) cudaEventDestroy(s->start);
    delete s;
    ctx->stager = nullptr;
}

static int stager_get(ozk_ctx* ctx, Stager** out) {
    if (ctx->stager && ctx->stager->ok) {
        *out = ctx->stager;
        return OZK_OK;
    }
    if (ctx->stager) stager_free(ctx);              // a half-built one left behind by an earlier failure
    Stager* s = new Stager();
    ctx->stager = s;
    if (const char* e = getenv("OZK_STAGE_THREADS")) s->active = std::max(1, std::min((int)Stager::kThreads, atoi(e)));
    for (int t = 0; t < s->active; t++) {
        OZK_CUDA(cudaStreamCreateWithFlags(&s->st[t], cudaStreamNonBlocking));
        OZK_CUDA(cudaEventCreateWithFlags(&s->done[t], cudaEventDisableTiming));
        for (int b = 0; b < Stager::kBufs; b++) {
            OZK_CUDA(cudaMallocHost(&s->pinned[t][b], Stager::kChunk));
            OZK_CUDA(cudaEventCreateWithFlags(&s->ev[t][b], cudaEventDisableTiming));
        }
    }
    OZK_CUDA(cudaEventCreateWithFlags(&s->start, cudaEventDisableTiming));
    s->ok = true;
    *out = s;
    return OZK_OK;
}

// work(0) on the calling thread, work(1..n-1) on helper threads; a helper that cannot be started (thread limit reached)
// has its share run inline instead -- no exception may cross the C ABI.
template <class W>
static void run_on_threads(W& work, int nthreads) {
    std::thread th[Stager::kThreads];
    bool started[Stager::kThreads] = {};
    for (int t = 1; t < nthreads; t++) {
        try {
            th[t] = std::thread([&work, t] { work(t); });
            started[t] = true;
        } catch (...) {
            started[t] = false;
        }
    }
    work(0);
    for (int t = 1; t < nthreads; t++) {
        if (started[t]) th[t].join();
        else work(t);
    }
}

bool host_pointer_is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// dst (device) <- src (host, pageable), `bytes` bytes.  The copies start after `after` (an event recorded where the staging
// destination was last used; may be null) and `consumer` is made to wait for all of them.  Returns when the host side is
// done with src (all chunks are in bounce buffers or on the device).
int staged_h2d(ozk_ctx* ctx, void* dst, const void* src, size_t bytes, cudaEvent_t after, cudaStream_t consumer) {
    if (bytes == 0) return OZK_OK;
    Stager* s;
    OZK_TRY(stager_get(ctx, &s));
    const size_t nchunks = (bytes + Stager::kChunk - 1) / Stager::kChunk;
    const int nthreads = (int)std::min<size_t>((size_t)s->active, nchunks);
    cudaError_t err[Stager::kThreads];
    auto work = [&](int t) {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e == cudaSuccess && after) e = cudaStreamWaitEvent(s->st[t], after, 0);
        int i = 0;
        for (size_t c = (size_t)t; c < nchunks && e == cudaSuccess; c += (size_t)nthreads, i++) {
            const int b = i % Stager::kBufs;
            const size_t off = c * Stager::kChunk;
            const size_t len = std::min(Stager::kChunk, bytes - off);
            e = cudaEventSynchronize(s->ev[t][b]);                 // the DMA that last read this bounce buffer has finished
            if (e != cudaSuccess) break;
            memcpy(s->pinned[t][b], (const char*)src + off, len);
            e = cudaMemcpyAsync((char*)dst + off, s->pinned[t][b], len, cudaMemcpyHostToDevice, s->st[t]);
            if (e == cudaSuccess) e = cudaEventRecord(s->ev[t][b], s->st[t]);
        }
        if (e == cudaSuccess) e = cudaEventRecord(s->done[t], s->st[t]);
        err[t] = e;
    };
    run_on_threads(work, nthreads);
    for (int t = 0; t < nthreads; t++) {
        if (err[t] != cudaSuccess) {
            set_error("staged upload: %s", cudaGetErrorString(err[t]));
            return OZK_ERR_CUDA;
        }
        OZK_CUDA(cudaStreamWaitEvent(consumer, s->done[t], 0));
    }
    return OZK_OK;
}

// dst (host, pageable) <- src (device): the mirror image.  The copies start after `after` (an event recorded after the
// producer of src; may be null: then the caller has synchronised already).  Returns when dst is complete.
int staged_d2h(ozk_ctx* ctx, void* dst, const void* src, size_t bytes, cudaEvent_t after) {
    if (bytes == 0) return OZK_OK;
    Stager* s;
    OZK_TRY(stager_get(ctx, &s));
    const size_t nchunks = (bytes + Stager::kChunk - 1) / Stager::kChunk;
    const int nthreads = (int)std::min<size_t>((size_t)s->active, nchunks);
    cudaError_t err[Stager::kThreads];
    auto work = [&](int t) {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e == cudaSuccess && after) e = cudaStreamWaitEvent(s->st[t], after, 0);
        // software pipeline over this thread's chunks: the DMA of chunk i+1 runs while chunk i is copied out of its bounce buffer
        size_t pend_off[Stager::kBufs] = {}, pend_len[Stager::kBufs] = {};
        bool pending[Stager::kBufs] = {};
        int i = 0;
        auto drain = [&](int b) {
            if (!pending[b] || e != cudaSuccess) return;
            e = cudaEventSynchronize(s->ev[t][b]);
            if (e == cudaSuccess) memcpy((char*)dst + pend_off[b], s->pinned[t][b], pend_len[b]);
            pending[b] = false;
        };
        for (size_t c = (size_t)t; c < nchunks && e == cudaSuccess; c += (size_t)nthreads, i++) {
            const int b = i % Stager::kBufs;
            drain(b);                                              // the bounce buffer is free again
            if (e != cudaSuccess) break;
            const size_t off = c * Stager::kChunk;
            const size_t len = std::min(Stager::kChunk, bytes - off);
            e = cudaMemcpyAsync(s->pinned[t][b], (const char*)src + off, len, cudaMemcpyDeviceToHost, s->st[t]);
            if (e == cudaSuccess) e = cudaEventRecord(s->ev[t][b], s->st[t]);
            pend_off[b] = off;
            pend_len[b] = len;
            pending[b] = (e == cudaSuccess);
        }
        for (int k = 0; k < Stager::kBufs; k++) drain((i + k) % Stager::kBufs);
        err[t] = e;
    };
    run_on_threads(work, nthreads);
    for (int t = 0; t < nthreads; t++) {
        if (err[t] != cudaSuccess) {
            set_error("staged download: %s", cudaGetErrorString(err[t]));
            return OZK_ERR_CUDA;
        }
    }
    return OZK_OK;
}

// convenience for the host-pointer entry points: pick the stager for large pageable buffers, plain async copies otherwise.
// upload: ordered after everything already on `stream`, and `stream` waits for it.  download: waits for `stream` first and
// returns with dst complete.
int upload_any(ozk_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    if (bytes >= ((size_t)8 << 20) && host_pointer_is_pageable(src)) {
        Stager* s;
        OZK_TRY(stager_get(ctx, &s));
        OZK_CUDA(cudaEventRecord(s->start, stream));
        return staged_h2d(ctx, dst, src, bytes, s->start, stream);
    }
    OZK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return OZK_OK;
}
int download_any(ozk_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    if (bytes >= ((size_t)8 << 20) && host_pointer_is_pageable(dst)) {
        Stager* s;
        OZK_TRY(stager_get(ctx, &s));
        OZK_CUDA(cudaEventRecord(s->start, stream));
        return staged_d2h(ctx, dst, src, bytes, s->start);
    }
    OZK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    OZK_CUDA(cudaStreamSynchronize(stream));
    return OZK_OK;
}

}  // namespace ozk
