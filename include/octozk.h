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

#if defined(__GNUC__)
#define OZK_API __attribute__((visibility("default")))
#else
#define OZK_API
#endif

#define OZK_OK 0
#define OZK_ERR_ARG (-1)
#define OZK_ERR_CUDA (-2)
#define OZK_ERR_DOMAIN (-3) /* omega is not a primitive n-th root of unity, value not reduced, ... */

typedef struct ozk_ctx ozk_ctx;
typedef struct ozk_bases ozk_bases; /* persistent device-resident bases, see below */

/* ---- context -------------------------------------------------------------------------------------------
 * One context per calling thread and device: own stream, scratch arena and caches, so concurrent JVM executor
 * threads do not share state (the reference shares the default stream and device-synchronises after every launch,
 * algebra_msm_VariableBaseMSM.cu:1286-1413).  The reference's placement rule is device = taskID % deviceCount
 * (algebra_msm_VariableBaseMSM.cu:1248-1257). */
OZK_API int ozk_device_count(void);
OZK_API int ozk_ctx_create(int device, ozk_ctx** out);
OZK_API void ozk_ctx_destroy(ozk_ctx* ctx);
/* Use an externally owned cudaStream_t (e.g. the caller's current stream) for all later work.  Work already enqueued by this
 * context is ordered before it (the new stream waits for the old one on the device).
 * STREAM CONTRACT of every "_dev" entry point: it only ENQUEUES on the context's stream.  Buffers written by the caller on
 * another stream must be ordered before the call, and results must be ordered before the caller reads them, by the caller:
 * either make the context run on the caller's stream with this function, or synchronise explicitly (ozk_ctx_sync).  The
 * Python binding does the former automatically whenever it is handed a CUDA torch tensor (octopuszk_b200/lib.py). */
OZK_API int ozk_ctx_set_stream(ozk_ctx* ctx, void* cuda_stream);
OZK_API int ozk_ctx_sync(ozk_ctx* ctx);
/* number of kernels this context has launched so far (bench.py reports the per-step delta as gpu_launches) */
OZK_API unsigned long long ozk_ctx_launches(ozk_ctx* ctx);
OZK_API const char* ozk_last_error(void);
OZK_API const char* ozk_version(void);

/* ---- Fr vector x constant ------------------------------------------------------------------------------
 * out[i] = a[i] * b mod r.  Replaces field_MSM / Java_algebra_msm_FixedBaseMSM_fieldBatchMSMNativeHelper
 * (algebra_msm_FixedBaseMSM.cu:1241-1266, :1500-1558). */
OZK_API int ozk_fr_scale(ozk_ctx* ctx, const uint8_t* a, size_t n, const uint8_t b[32], uint8_t* out);
OZK_API int ozk_fr_scale_dev(ozk_ctx* ctx, const void* d_a, void* d_out, size_t n, const uint8_t b[32]);

/* out[i] = a[i] * scale * coset^i (scale and/or coset may be NULL = 1).  FFTAuxiliary.multiplyByCoset
 * (src/main/java/algebra/fft/FFTAuxiliary.java:224-232) and distributedMultiplyByCoset (:237-243) with `first_index` as the
 * global index of a[0], so a shard of a larger vector can be scaled in place. */
OZK_API int ozk_fr_scale_powers_dev(ozk_ctx* ctx, const void* d_a, void* d_out, size_t n, const uint8_t* scale, const uint8_t* coset,
                            uint64_t first_index);

/* out[i] = a[i] * b[i] - c[i] mod r (d_c may be NULL: plain product; d_out may alias an input): the pointwise step between
 * the transforms of R1CStoQAP.R1CStoQAPWitness (src/main/java/reductions/r1cs_to_qap/R1CStoQAP.java:180-182,211-214), so
 * that A, B, C and H stay on the device from the first inverse FFT to the H MSM. */
OZK_API int ozk_fr_mul_sub_dev(ozk_ctx* ctx, const void* d_a, const void* d_b, const void* d_c, void* d_out, size_t n);

/* out[i] = L_i(t), the m Lagrange coefficients of the domain S = {omega^0 .. omega^(m-1)} at t (the unit vector e_i when
 * t == omega^i): FFTAuxiliary.serialRadix2LagrangeCoefficients (src/main/java/algebra/fft/FFTAuxiliary.java:249-302), the
 * m field inversions of R1CStoQAP.R1CStoQAPRelation (src/main/java/reductions/r1cs_to_qap/R1CStoQAP.java:54-56) in the
 * setup.  m a power of two <= 2^28, omega a primitive m-th root of unity.  SURVEY.md section 8f row 4. */
OZK_API int ozk_fr_lagrange_dev(ozk_ctx* ctx, void* d_out, size_t m, const uint8_t t[32], const uint8_t omega[32]);

/* out[i] = sum_k coeff[k] * z[col[k]] over k in [row_ptr[i], row_ptr[i+1]) mod r: the linear combinations a_i, b_i, c_i of all
 * constraints evaluated at the full assignment z (CSR: row_ptr has rows + 1 uint32 entries, col uint32 indices into z, coeff and
 * z 32-byte canonical elements) -- the host loop of R1CStoQAP.R1CStoQAPWitness (src/main/java/reductions/r1cs_to_qap/
 * R1CStoQAP.java:143-160, LinearCombination.evaluate).  At most 4096 rows may have more than 64 terms.  SURVEY.md 8f row 4. */
OZK_API int ozk_fr_spmv_dev(ozk_ctx* ctx, const void* d_row_ptr, const void* d_col, const void* d_coeff, const void* d_z, size_t rows,
                            void* d_out);
/* Same with the length of z given: a column index >= z_len is reported as OZK_ERR_ARG instead of being dereferenced (the form
 * above trusts the matrix).  d_coeff == NULL: every coefficient is 1 (the reference's synthetic circuit,
 * src/main/java/profiler/generation/R1CSConstruction.java:48-104, has no other), which also serves the transposed products of
 * R1CStoQAP.R1CStoQAPRelation (At[j] = sum_i L_i(t) A[i][j], R1CStoQAP.java:62-80) with z = the Lagrange coefficients. */
OZK_API int ozk_fr_spmv_ex_dev(ozk_ctx* ctx, const void* d_row_ptr, const void* d_col, const void* d_coeff, const void* d_z, size_t z_len,
                               size_t rows, void* d_out);
/* out[i] = ca * a[i] + cb * b[i] + cc * c[i] mod r (d_b / d_c may be NULL; out may alias an input): the vector combinations of
 * SerialSetup.generate (beta At + alpha Bt + Ct and the divisions by gamma / delta,
 * src/main/java/zk_proof_systems/zkSNARK/SerialSetup.java:66-85) on device-resident vectors. */
OZK_API int ozk_fr_lincomb_dev(ozk_ctx* ctx, void* d_out, size_t n, const void* d_a, const uint8_t ca[32], const void* d_b,
                               const uint8_t cb[32], const void* d_c, const uint8_t cc[32]);

/* ---- radix-2 NTT over Fr ---------------------------------------------------------------------------------
 * out[k] = sum_j in[j] * omega^(j k), natural order in and out, n a power of two <= 2^28, omega a primitive n-th
 * root of unity.  Replaces FFTAuxiliary.serialRadix2FFT (src/main/java/algebra/fft/FFTAuxiliary.java:60-124) and the
 * dormant Java_algebra_fft_FFTAuxiliary_serialRadix2FFTNativeHelper / best_fft (algebra_fft_FFTAuxiliary.cu:167-260).
 * d_in == d_out is allowed.  Elements must be reduced mod r: the host-pointer entry points (ozk_ntt_fr, ozk_fr_scale) check it
 * and return OZK_ERR_DOMAIN; the "_dev" entry points only enqueue work and do not (unreduced elements give unspecified
 * values there, never a memory fault). */
OZK_API int ozk_ntt_fr(ozk_ctx* ctx, uint8_t* data, size_t n, const uint8_t omega[32]);
OZK_API int ozk_ntt_fr_dev(ozk_ctx* ctx, const void* d_in, void* d_out, size_t n, const uint8_t omega[32]);
/* Fused wrappers of src/main/java/algebra/fft/SerialFFT.java:75-115,157-162:
 *   in[i] *= pre_coset^i (if non-NULL), transform with omega, out[i] *= post_scale * post_coset^i (either may be NULL).
 *   radix2InverseFFT      : omega^-1, post_scale = n^-1
 *   radix2CosetFFT        : pre_coset = g
 *   radix2CosetInverseFFT : omega^-1, post_scale = n^-1, post_coset = g^-1
 *   divideByZOnCoset      : folds into post_scale */
OZK_API int ozk_ntt_fr_ex_dev(ozk_ctx* ctx, const void* d_in, void* d_out, size_t n, const uint8_t omega[32],
                      const uint8_t* pre_coset, const uint8_t* post_scale, const uint8_t* post_coset);

/* Cross-shard step of the multi-GPU transform (four-step with n = G * M, see octopuszk_b200/distributed.py, replacing the
 * two Spark shuffles of FFTAuxiliary.distributedRadix2FFT, src/main/java/algebra/fft/FFTAuxiliary.java:129-219):
 * out[k1 * len + j] = sum_{i1 < groups} in[i1 * len + j] * omega_g^(i1 k1), groups in {1,2,4,8}, omega_g a primitive
 * groups-th root of unity.  d_in != d_out. */
OZK_API int ozk_fr_dft_small_dev(ozk_ctx* ctx, const void* d_in, void* d_out, size_t groups, size_t len, const uint8_t omega_g[32]);

/* Fused form of step 1 + twiddle + exchange of the same four-step transform: the shard-local n_local-point transform (with
 * omega_local = omega_n^groups) whose last pass multiplies output element i by twiddle_base^i (twiddle_base = omega_n^rank)
 * and stores it straight into the buffer of the rank that owns chunk i / (n_local / groups), at element
 * rank * (n_local / groups) + i mod (n_local / groups).  peer_out[r] is rank r's receive buffer (n_local x 32 B) mapped into
 * this process (ozk_peer_open; peer_out[rank] is this rank's own), so the all-to-all travels over NVLink as coalesced 128-byte
 * stores from the epilogue of the transform instead of a separate NCCL collective and two more passes over HBM.  The caller
 * orders the ranks (a barrier before: receive buffers free; a barrier after: all stores landed) and then runs
 * ozk_fr_dft_small_dev on its receive buffer.  d_in is not modified. */
OZK_API int ozk_ntt_fr_scatter_dev(ozk_ctx* ctx, const void* d_in, void* const* peer_out, size_t groups, size_t rank, size_t n_local,
                                   const uint8_t omega_local[32], const uint8_t twiddle_base[32]);
/* The mirrored transform, for data that is already in the output layout of the one above (rank d holds x[a * M + d * len + t],
 * a < groups, t < len = M / groups, as [a][t]): groups-point transforms over a (omega_g = omega_n^M), twiddle
 * omega_n^((d len + t) k1), and result k1 stored straight into rank k1's buffer at element d * len + t.  After a barrier every
 * rank runs the plain M-point transform (omega_n^groups) on its receive buffer and holds X[rank + groups * k2]: the cyclic
 * layout again.  The two forms alternate through the prover's transform chain (inverse, coset forward, ..., R1CStoQAP.java:
 * 165-227) so no re-layout is ever needed.  The omega_n table is validated only through omega_g = omega_n^M by the caller. */
OZK_API int ozk_fr_dft_small_scatter_dev(ozk_ctx* ctx, const void* d_in, void* const* peer_out, size_t groups, size_t rank, size_t len,
                                         const uint8_t omega_g[32], const uint8_t omega_n[32]);
/* Receive buffers that other processes on the same node can map: cudaMalloc + CUDA IPC handle (64 bytes, to be sent to the
 * peers by any means, e.g. torch.distributed.all_gather); ozk_peer_open maps a peer's buffer with peer access enabled. */
OZK_API int ozk_peer_alloc(ozk_ctx* ctx, size_t bytes, void** d_ptr, uint8_t handle[64]);
OZK_API int ozk_peer_open(ozk_ctx* ctx, const uint8_t handle[64], void** d_ptr);
OZK_API int ozk_peer_close(ozk_ctx* ctx, void* d_ptr);
OZK_API int ozk_peer_free(ozk_ctx* ctx, void* d_ptr);

/* ---- variable-base MSM --------------------------------------------------------------------------------------
 * out = sum_i scalars[i] * bases[i].  Replaces VariableBaseMSM.serialMSM's native leg
 * Java_algebra_msm_VariableBaseMSM_variableBaseSerialMSMNativeHelper (algebra_msm_VariableBaseMSM.cu:1614-1695,
 * pippengerMSMG1 :1246-1428, pippengerMSMG2 :1433-1604) and, for the paired form, ...variableBaseDoubleMSMNativeHelper
 * (:1712-1788), which runs G1 then G2 on the same scalars; here both share one scalar sort.
 * scalars: n x 32 B (< r).  G1 bases: n x 96 B, G2 bases: n x 192 B (Jacobian, any Z; Z == 0 is infinity).
 * out: one point in the same layout (G1 96 B, G2 192 B, paired: G1 || G2 = 288 B); infinity is (0,1,0).
 * n == 0 gives infinity.  Unlike the Java caller's 2^23 / 2^22 / 2^21 chunking (VariableBaseMSM.java:211,268,494) any
 * n < 2^31 that fits in device memory is accepted in one call. */
OZK_API int ozk_msm_g1(ozk_ctx* ctx, const uint8_t* scalars, const uint8_t* bases, size_t n, uint8_t out[96]);
OZK_API int ozk_msm_g1_dev(ozk_ctx* ctx, const void* d_scalars, const void* d_bases, size_t n, uint8_t out[96]);
OZK_API int ozk_msm_g2(ozk_ctx* ctx, const uint8_t* scalars, const uint8_t* bases, size_t n, uint8_t out[192]);
OZK_API int ozk_msm_g2_dev(ozk_ctx* ctx, const void* d_scalars, const void* d_bases, size_t n, uint8_t out[192]);
OZK_API int ozk_msm_g1g2(ozk_ctx* ctx, const uint8_t* scalars, const uint8_t* bases1, const uint8_t* bases2, size_t n, uint8_t out[288]);
OZK_API int ozk_msm_g1g2_dev(ozk_ctx* ctx, const void* d_scalars, const void* d_bases1, const void* d_bases2, size_t n, uint8_t out[288]);
/* Streaming form of the host-pointer MSM (what ozk_msm_g1 / _g2 / _g1g2 and the keyed forms do internally): announce the total,
 * feed the (scalar, point) pairs in slices, collect the result.  groups: 1 = G1, 2 = G2, 3 = both on the same scalars.
 *   ozk_msm_begin(ctx, groups, n_total, max_slice, key1, key2, first)   key1/key2: persistent bases for that group (then feed
 *                                      takes NULL for its base array) or NULL; max_slice: longest slice to come (0 = unknown),
 *                                      lets all scratch be reserved up front
 *   ozk_msm_feed(ctx, scalars, bases1, bases2, len)   next `len` pairs; returns as soon as the host arrays are no longer read
 *                                      (copied to pinned staging or to the device) while the GPU works on them -- a JNI
 *                                      caller pins its byte[] only around this call, one slice at a time, instead of across the
 *                                      whole MSM (the Java's own 2^23-element chunk loop, VariableBaseMSM.java:211-265, maps
 *                                      onto feeds of one MSM too)
 *   ozk_msm_end(ctx, out)              bucket reduction + window recombination, result as for ozk_msm_*
 *   ozk_msm_plan_slices(n, bounds, cap) the slice schedule the whole-array entry points use for pageable host memory (a small
 *                                      first and last slice: the call is bound by the copies): returns k <= cap and fills
 *                                      bounds[0..k] (bounds[0] = 0, bounds[k] = n)
 * Unreduced inputs are reported by ozk_msm_end.  One MSM in progress per context. */
OZK_API int ozk_msm_begin(ozk_ctx* ctx, int groups, size_t n_total, size_t max_slice, const ozk_bases* key1, const ozk_bases* key2,
                          size_t first);
OZK_API int ozk_msm_feed(ozk_ctx* ctx, const uint8_t* scalars, const uint8_t* bases1, const uint8_t* bases2, size_t len);
OZK_API int ozk_msm_end(ozk_ctx* ctx, uint8_t* out);
OZK_API int ozk_msm_plan_slices(size_t n, size_t* bounds, int cap);

/* ---- persistent bases (device-resident proving key) ---------------------------------------------------------
 * The query vectors of a Groth16 proving key (src/main/java/zk_proof_systems/zkSNARK/objects/ProvingKey.java:16-47:
 * queryA, queryB, deltaABCG1, queryH) are the same for every proof, but the reference re-marshals and re-uploads them on
 * every MSM (VariableBaseMSM.java:217-237; SerialProver.java:70-106 passes them to serialMSM / doubleMSM each time).
 * ozk_bases_upload_* copies n wire-format points (host memory, or device memory for _dev) once and keeps them on the
 * context's device in the affine Montgomery form the bucket kernels read; the _keyed MSMs then take only scalars:
 *   out = sum_{i < n} scalars[i] * key[first + i]
 * (`first` covers both the Java chunk loop and the subList calls of the prover.)  Results are identical to the plain
 * entry points on the same points.  A handle belongs to the device of the context that made it; free it with
 * ozk_bases_free before destroying that context.  SURVEY.md section 8f, row 3. */
OZK_API int ozk_bases_upload_g1(ozk_ctx* ctx, const uint8_t* bases, size_t n, ozk_bases** out);
OZK_API int ozk_bases_upload_g1_dev(ozk_ctx* ctx, const void* d_bases, size_t n, ozk_bases** out);
OZK_API int ozk_bases_upload_g2(ozk_ctx* ctx, const uint8_t* bases, size_t n, ozk_bases** out);
OZK_API int ozk_bases_upload_g2_dev(ozk_ctx* ctx, const void* d_bases, size_t n, ozk_bases** out);
OZK_API size_t ozk_bases_len(const ozk_bases* key);
OZK_API void ozk_bases_free(ozk_ctx* ctx, ozk_bases* key);
OZK_API int ozk_msm_g1_keyed(ozk_ctx* ctx, const uint8_t* scalars, const ozk_bases* key, size_t first, size_t n, uint8_t out[96]);
OZK_API int ozk_msm_g1_keyed_dev(ozk_ctx* ctx, const void* d_scalars, const ozk_bases* key, size_t first, size_t n, uint8_t out[96]);
OZK_API int ozk_msm_g2_keyed(ozk_ctx* ctx, const uint8_t* scalars, const ozk_bases* key, size_t first, size_t n, uint8_t out[192]);
OZK_API int ozk_msm_g2_keyed_dev(ozk_ctx* ctx, const void* d_scalars, const ozk_bases* key, size_t first, size_t n, uint8_t out[192]);
OZK_API int ozk_msm_g1g2_keyed(ozk_ctx* ctx, const uint8_t* scalars, const ozk_bases* key1, const ozk_bases* key2, size_t first, size_t n,
                               uint8_t out[288]);
OZK_API int ozk_msm_g1g2_keyed_dev(ozk_ctx* ctx, const void* d_scalars, const ozk_bases* key1, const ozk_bases* key2, size_t first, size_t n,
                                   uint8_t out[288]);

/* out = sum of k <= 4096 points given in the wire format in device memory: the reduce(add) over the partial sums of a
 * sharded MSM (VariableBaseMSM.distributedMSM, src/main/java/algebra/msm/VariableBaseMSM.java:777-783). */
OZK_API int ozk_sum_g1_dev(ozk_ctx* ctx, const void* d_points, size_t k, uint8_t out[96]);
OZK_API int ozk_sum_g2_dev(ozk_ctx* ctx, const void* d_points, size_t k, uint8_t out[192]);

/* last MSM on this context: {window bits, windows, buckets per window, overflow tasks, overflow buckets,
 * ms sort, ms convert, ms accumulate, ms merge, ms reduce+final, GB/s of the first slice's upload in the last whole-array
 * host-pointer call from pinned memory (what its slice schedule was chosen by)} (device times from events on the context's
 * stream; for the paired call the per-group phases are those of the G2 half) */
OZK_API int ozk_msm_last_stats(ozk_ctx* ctx, double* out, int cap);

/* ---- fixed-base batch MSM -----------------------------------------------------------------------------------
 * out[i] = (scalars[i] mod 2^(outerc * windowSize)) * base, i.e. exactly the outerc windows of windowSize bits the
 * reference walks (fixedbase_MSM_unit_processing_G1/G2, algebra_msm_FixedBaseMSM.cu:750-850; Java oracle
 * FixedBaseMSM.serialMSM, src/main/java/algebra/msm/FixedBaseMSM.java:141-167).  Replaces
 * Java_algebra_msm_FixedBaseMSM_batchMSMNativeHelper (algebra_msm_FixedBaseMSM.cu:1276-1384); the paired
 * ...doubleBatchMSMNativeHelper (:1395-1491) is one G1 and one G2 call on the same scalars.
 * base: one point (G1 96 B / G2 192 B), host memory in both variants.  scalars: n x 32 B (< r).
 * out: n points in the wire layout, normalised to Z = 1 (infinity as (0,1,0)).  The window table is cached on the context. */
OZK_API int ozk_fixed_g1(ozk_ctx* ctx, const uint8_t base[96], const uint8_t* scalars, size_t n, int outerc, int windowSize, uint8_t* out);
OZK_API int ozk_fixed_g1_dev(ozk_ctx* ctx, const uint8_t base[96], const void* d_scalars, size_t n, int outerc, int windowSize, void* d_out);
OZK_API int ozk_fixed_g2(ozk_ctx* ctx, const uint8_t base[192], const uint8_t* scalars, size_t n, int outerc, int windowSize, uint8_t* out);
OZK_API int ozk_fixed_g2_dev(ozk_ctx* ctx, const uint8_t base[192], const void* d_scalars, size_t n, int outerc, int windowSize, void* d_out);
/* Same with flags.  OZK_FIXED_KEEP_Z: skip the normalisation and return every point as the Jacobian triple of the accumulator's
 * own denominator (arbitrary Z, one launch less) -- the form the reference's own outputs have on the wire
 * (fixedbase_MSM_unit_processing_G1, algebra_msm_FixedBaseMSM.cu:750-850, returns unnormalised Jacobian sums).  The points are
 * the same group elements; only the representative differs. */
#define OZK_FIXED_KEEP_Z 1u
OZK_API int ozk_fixed_g1_ex_dev(ozk_ctx* ctx, const uint8_t base[96], const void* d_scalars, size_t n, int outerc, int windowSize,
                                unsigned flags, void* d_out);
OZK_API int ozk_fixed_g2_ex_dev(ozk_ctx* ctx, const uint8_t base[192], const void* d_scalars, size_t n, int outerc, int windowSize,
                                unsigned flags, void* d_out);

/* ---- diagnostics ----------------------------------------------------------------------------------------- */
/* Integer-pipe microbenchmark: independent 32x32+64 multiply-add chains on every SM; reports billions of
 * multiply-adds per second.  bench.py uses it as the measured integer roofline (MEASURED_PEAKS.json has none). */
OZK_API int ozk_imad_peak(ozk_ctx* ctx, double* gimad_per_s);
/* Issue-rate probes of the other pipes (billions of operations per second over all SMs): which = 1 DFMA chains,
 * 2 DFMA chains with IMAD.WIDE chains interleaved (reports the DFMA rate; 32 wide multiply-adds ride along per 64 DFMA),
 * 3 IADD3 carry chains, 4 32-bit IMAD chains.  Planning data for DESIGN.md section 4. */
OZK_API int ozk_pipe_probe(ozk_ctx* ctx, int which, double* gops_per_s);
/* Fr Montgomery multiplications per second with all operands in registers (upper bound for the field kernels). */
OZK_API int ozk_modmul_peak(ozk_ctx* ctx, double* gmodmul_per_s);

#ifdef __cplusplus
}
#endif
#endif /* OCTOZK_H */
