/* liboctozk -- B200 (sm_100a) implementation of the Groth16 arithmetic hot path of brucechin/OctopusZK (GPU-DIZK):
 * variable-base MSM, fixed-base batch MSM and the radix-2 NTT over BN254a, behind a plain C ABI.
 *
 * This is the drop-in boundary.  The three JNI libraries the reference's Java loads
 * (libAlgebraMSMVariableBaseMSM.so, libAlgebraMSMFixedBaseMSM.so, libAlgebraFFTAuxiliary.so; see jni/ and
 * INTEGRATION.md) are thin shims over these entry points.  Paths below are relative to the reference tree.
 *
 * Data formats:
 *   field element  : 32 bytes little-endian, canonical residue in [0, modulus), never Montgomery form
 *                    (what bigIntegerToByteArrayHelperCGBN produces, src/main/java/algebra/msm/VariableBaseMSM.java:121-131)
 *   G1 point       : X|Y|Z Jacobian, 3 x 32 B (VariableBaseMSM.java:224-227); infinity iff Z == 0
 *   G2 point       : X.c0|X.c1|Y.c0|Y.c1|Z.c0|Z.c1, 6 x 32 B (VariableBaseMSM.java:279-285)
 * Outputs use the same 32-byte layouts, fully reduced; returned points are equal to the reference's as group
 * elements (the reference compares projectively, src/main/java/algebra/curves/barreto_naehrig/BNG1.java:191-224).
 * The legacy 64-byte-per-coordinate return formats exist only in the JNI shims.
 *
 * "_dev" entry points take device pointers (cudaMalloc memory on the context's device) and enqueue on the context's
 * stream; the others take host pointers and copy in and out themselves.  Every function returns OZK_OK or a
 * negative error code and never terminates the process (the reference exit(-1)s, algebra_msm_VariableBaseMSM.cu:1417-1422);
 * ozk_last_error() gives the message for the calling thread.  There is no CPU fallback: without a usable CUDA device
 * every compute entry point fails with OZK_ERR_CUDA.
 */
#ifndef OCTOZK_H
#define OCTOZK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OZK_OK 0
#define OZK_ERR_ARG (-1)
#define OZK_ERR_CUDA (-2)
#define OZK_ERR_DOMAIN (-3) /* omega is not a primitive n-th root of unity, value not reduced, ... */

typedef struct ozk_ctx ozk_ctx;

/* ---- context -------------------------------------------------------------------------------------------
 * One context per calling thread and device: own stream, scratch arena and caches, so concurrent JVM executor
 * threads do not share state (the reference shares the default stream and device-synchronises after every launch,
 * algebra_msm_VariableBaseMSM.cu:1286-1413).  The reference's placement rule is device = taskID % deviceCount
 * (algebra_msm_VariableBaseMSM.cu:1248-1257). */
int ozk_device_count(void);
int ozk_ctx_create(int device, ozk_ctx** out);
void ozk_ctx_destroy(ozk_ctx* ctx);
/* use an externally owned cudaStream_t (e.g. the caller's current stream) for all later work */
int ozk_ctx_set_stream(ozk_ctx* ctx, void* cuda_stream);
int ozk_ctx_sync(ozk_ctx* ctx);
const char* ozk_last_error(void);
const char* ozk_version(void);

/* ---- Fr vector x constant ------------------------------------------------------------------------------
 * out[i] = a[i] * b mod r.  Replaces field_MSM / Java_algebra_msm_FixedBaseMSM_fieldBatchMSMNativeHelper
 * (algebra_msm_FixedBaseMSM.cu:1241-1266, :1500-1558). */
int ozk_fr_scale(ozk_ctx* ctx, const uint8_t* a, size_t n, const uint8_t b[32], uint8_t* out);
int ozk_fr_scale_dev(ozk_ctx* ctx, const void* d_a, void* d_out, size_t n, const uint8_t b[32]);

/* ---- radix-2 NTT over Fr ---------------------------------------------------------------------------------
 * out[k] = sum_j in[j] * omega^(j k), natural order in and out, n a power of two <= 2^28, omega a primitive n-th
 * root of unity.  Replaces FFTAuxiliary.serialRadix2FFT (src/main/java/algebra/fft/FFTAuxiliary.java:60-124) and the
 * dormant Java_algebra_fft_FFTAuxiliary_serialRadix2FFTNativeHelper / best_fft (algebra_fft_FFTAuxiliary.cu:167-260).
 * d_in == d_out is allowed. */
int ozk_ntt_fr(ozk_ctx* ctx, uint8_t* data, size_t n, const uint8_t omega[32]);
int ozk_ntt_fr_dev(ozk_ctx* ctx, const void* d_in, void* d_out, size_t n, const uint8_t omega[32]);
/* Fused wrappers of src/main/java/algebra/fft/SerialFFT.java:75-115,157-162:
 *   in[i] *= pre_coset^i (if non-NULL), transform with omega, out[i] *= post_scale * post_coset^i (either may be NULL).
 *   radix2InverseFFT      : omega^-1, post_scale = n^-1
 *   radix2CosetFFT        : pre_coset = g
 *   radix2CosetInverseFFT : omega^-1, post_scale = n^-1, post_coset = g^-1
 *   divideByZOnCoset      : folds into post_scale */
int ozk_ntt_fr_ex_dev(ozk_ctx* ctx, const void* d_in, void* d_out, size_t n, const uint8_t omega[32],
                      const uint8_t* pre_coset, const uint8_t* post_scale, const uint8_t* post_coset);

/* ---- diagnostics ----------------------------------------------------------------------------------------- */
/* Integer-pipe microbenchmark: independent 32x32+64 multiply-add chains on every SM; reports billions of
 * multiply-adds per second.  bench.py uses it as the measured integer roofline (MEASURED_PEAKS.json has none). */
int ozk_imad_peak(ozk_ctx* ctx, double* gimad_per_s);
/* Fr Montgomery multiplications per second with all operands in registers (upper bound for the field kernels). */
int ozk_modmul_peak(ozk_ctx* ctx, double* gmodmul_per_s);

#ifdef __cplusplus
}
#endif
#endif /* OCTOZK_H */
