// Internal (non-ABI) definitions shared by the .cu translation units of liboctozk.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/octozk.h"

namespace ozk {

void set_error(const char* fmt, ...);

#define OZK_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            ::ozk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return OZK_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define OZK_TRY(call)              \
    do {                           \
        int rc__ = (call);         \
        if (rc__ != OZK_OK) return rc__; \
    } while (0)

#define OZK_ARG(cond, msg)                 \
    do {                                   \
        if (!(cond)) {                     \
            ::ozk::set_error("%s", msg);   \
            return OZK_ERR_ARG;            \
        }                                  \
    } while (0)

// A grow-only device buffer.  Growing synchronises the stream first (the old block may still be in use).
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int reserve(size_t need, cudaStream_t s) {
        if (need <= bytes) return OZK_OK;
        if (p) {
            OZK_CUDA(cudaStreamSynchronize(s));
            OZK_CUDA(cudaFree(p));
            p = nullptr;
            bytes = 0;
        }
        size_t want = need + need / 8;
        OZK_CUDA(cudaMalloc(&p, want));
        bytes = want;
        return OZK_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

struct NttPlan;
struct FixedTable;
struct Stager;

}  // namespace ozk

// The context: one per (thread, device) user.  Re-entrancy contract: a context must not be
// used from two threads at once; separate contexts are independent (own stream, own scratch, own caches).
struct ozk_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t copy_stream = nullptr;        // second stream for chunked host->device uploads (host-pointer MSM entry)
    cudaEvent_t copy_ev[20] = {};
    cudaStream_t side_stream = nullptr;        // high-priority side stream: the normalisation of a slice's bases runs here while the
    int conv_forked = 0;                       // same slice is sorted on the main stream (msm.cu, msm_convert_fork; bit g = group g forked)
    cudaEvent_t evs[8] = {};                   // phase marks of the last MSM (see ozk_msm_last_stats)
    unsigned long long launches = 0;           // kernels launched through this context
    // scratch
    ozk::DevBuf io_a, io_b, io_c, io_out;     // staging for the host-pointer entry points
    ozk::DevBuf work;                          // NTT ping-pong buffer
    ozk::DevBuf msm[16];                       // MSM pipeline buffers (see msm.cu)
    ozk::DevBuf fb[4];                         // fixed-base buffers
    alignas(8) unsigned char msm_stream[96] = {};  // state of the host-pointer MSM in progress (MsmStream, msm.cu)
    double msm_stats[16] = {};                  // last MSM: window c, windows, buckets/window, overflow tasks, overflow buckets
    void* pinned = nullptr;                    // small pinned host block for flags / results
    std::map<std::string, ozk::NttPlan*> ntt_plans;
    std::map<std::string, ozk::FixedTable*> fixed_tables;
    ozk::Stager* stager = nullptr;              // bounce buffers for uploads from pageable host memory (stage.cu)
};

// Persistent device-resident bases (ozk_bases_upload_*): affine Montgomery coordinates, (0,0) for infinity.
struct ozk_bases {
    int group = 0;            // 1 = G1, 2 = G2
    int device = 0;
    size_t n = 0;
    void* d_affine = nullptr;
};

namespace ozk {
static constexpr int kCopyChunks = 8;      // chunks per uploaded base array (<= 2 arrays x 8 + 1 events)
// activates ctx->device for the calling thread
int ctx_enter(ozk_ctx* ctx);
// stage.cu: uploads from pageable host memory through pinned bounce buffers on helper threads
bool host_pointer_is_pageable(const void* p);
int staged_h2d(ozk_ctx* ctx, void* dst, const void* src, size_t bytes, cudaEvent_t after, cudaStream_t consumer);
int staged_d2h(ozk_ctx* ctx, void* dst, const void* src, size_t bytes, cudaEvent_t after);
int upload_any(ozk_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t stream);      // ordered on `stream`
int download_any(ozk_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t stream);    // returns with dst complete
void stager_free(ozk_ctx* ctx);
}  // namespace ozk
