// JNI shims: the six native methods the reference's Java declares, bound to the C ABI of include/octozk.h.
//
// Built three times (build.py) into the three libraries the Java static initialisers load:
//   -DOZK_SHIM_VARMSM   -> libAlgebraMSMVariableBaseMSM.so  (System.loadLibrary, src/main/java/algebra/msm/VariableBaseMSM.java:31-34)
//   -DOZK_SHIM_FIXEDMSM -> libAlgebraMSMFixedBaseMSM.so     (src/main/java/algebra/msm/FixedBaseMSM.java:44-47)
//   -DOZK_SHIM_FFT      -> libAlgebraFFTAuxiliary.so        (src/main/java/algebra/fft/SerialFFT.java:20-23)
// Symbol names, argument order and byte layouts are those of the javah headers (algebra_msm_VariableBaseMSM.h:10-24,
// algebra_msm_FixedBaseMSM.h:10-32, algebra_fft_FFTAuxiliary.h:10-16) and of SURVEY.md Appendix A.  What changes:
//   * the reference calls GetByteArrayElements and never releases (algebra_msm_VariableBaseMSM.cu:1624,1634).  Here the large
//     MSM inputs are pinned with Get/ReleasePrimitiveArrayCritical(JNI_ABORT) for ONE SLICE's host copy at a time (ozk_msm_feed
//     returns as soon as the slice sits in pinned staging; the GPU work, the final reduction and every JNI call -- FindClass,
//     ThrowNew, NewByteArray -- happen outside any critical region, so the collector is never locked out for a whole MSM and
//     the JNI rule "no JNI calls inside a critical region" holds); the smaller fixed-base / field inputs are copied out with
//     GetByteArrayRegion and never pinned;
//   * errors become java.lang.RuntimeException via ThrowNew and a NULL return -- the reference prints and continues or
//     exit(-1)s (algebra_msm_VariableBaseMSM.cu:23-28,1417-1422);
//   * one liboctozk context per (thread, device): Spark executor threads enter concurrently (hs_err_pid98479.log:260-272);
//   * device = taskID % deviceCount as in the reference (algebra_msm_VariableBaseMSM.cu:1248-1257).
// The "...Direct" natives take java.nio direct ByteBuffers of 32-byte elements in and out (no per-element BigInteger
// marshalling, no 64-byte padding): the Java edits that call them are in INTEGRATION.md.
// No JVM exists in this image: the shims are exercised through a fake JNIEnv (tests/fake_jni.cc).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#include "jni_min.h"
#include "octozk.h"

namespace {

struct ThreadCtx {
    std::map<int, ozk_ctx*> by_device;
    ~ThreadCtx() {
        for (auto& kv : by_device) ozk_ctx_destroy(kv.second);
    }
};

ozk_ctx* context_for_task(jint task_id) {
    static thread_local ThreadCtx tc;
    int ndev = ozk_device_count();
    if (ndev <= 0) return nullptr;
    int dev = (int)(((long long)task_id % ndev + ndev) % ndev);
    auto it = tc.by_device.find(dev);
    if (it != tc.by_device.end()) return it->second;
    ozk_ctx* c = nullptr;
    if (ozk_ctx_create(dev, &c) != OZK_OK) return nullptr;
    tc.by_device[dev] = c;
    return c;
}

jbyteArray fail(JNIEnv* env, const char* where) {
    char msg[640];
    snprintf(msg, sizeof msg, "%s: %s", where, ozk_last_error());
    jclass cls = env->FindClass("java/lang/RuntimeException");
    if (cls) env->ThrowNew(cls, msg);
    return nullptr;
}
jbyteArray fail_msg(JNIEnv* env, const char* msg) {
    jclass cls = env->FindClass("java/lang/RuntimeException");
    if (cls) env->ThrowNew(cls, msg);
    return nullptr;
}

// RAII pin of a Java byte[]
struct Pinned {
    JNIEnv* env;
    jbyteArray arr;
    void* p;
    Pinned(JNIEnv* e, jbyteArray a) : env(e), arr(a), p(a ? e->GetPrimitiveArrayCritical(a, nullptr) : nullptr) {}
    ~Pinned() {
        if (p) env->ReleasePrimitiveArrayCritical(arr, p, JNI_ABORT);
    }
    const uint8_t* bytes() const { return (const uint8_t*)p; }
};

// 32-byte little-endian coordinates -> 64-byte slots, little-endian (variable-base, FFT) or big-endian (fixed-base,
// field batch): SURVEY.md Appendix A.1 / A.2.
void widen(const uint8_t* in, size_t coords, uint8_t* out, bool big_endian) {
    for (size_t k = 0; k < coords; k++) {
        uint8_t* o = out + 64 * k;
        const uint8_t* s = in + 32 * k;
        memset(o, 0, 64);
        if (big_endian) {
            for (int j = 0; j < 32; j++) o[63 - j] = s[j];
        } else {
            memcpy(o, s, 32);
        }
    }
}

jbyteArray to_java(JNIEnv* env, const std::vector<uint8_t>& v) {
    if (v.size() > 0x7fffffffu) return fail_msg(env, "result exceeds the 2 GiB limit of a Java byte[]");
    jbyteArray out = env->NewByteArray((jsize)v.size());
    if (!out) return nullptr;
    env->SetByteArrayRegion(out, 0, (jsize)v.size(), (const jbyte*)v.data());
    return out;
}

bool long_enough(JNIEnv* env, jbyteArray a, size_t need) { return a && (size_t)env->GetArrayLength(a) >= need; }

// first `bytes` bytes of a Java byte[] into a native buffer, in pieces (no pin; a piece is a plain memcpy inside the JVM)
void copy_out(JNIEnv* env, jbyteArray a, size_t bytes, uint8_t* dst) {
    const size_t piece = (size_t)4 << 20;
    for (size_t off = 0; off < bytes; off += piece) {
        const size_t len = bytes - off < piece ? bytes - off : piece;
        env->GetByteArrayRegion(a, (jsize)off, (jsize)len, (jbyte*)(dst + off));
    }
}

// One variable-base MSM over Java arrays, slice by slice: the arrays are pinned only while ozk_msm_feed copies a slice out of
// them.  groups: 1 = G1 (b1), 2 = G2 (b2), 3 = both.  Returns OZK_OK, an ozk error (message in ozk_last_error), or 1 when an
// array could not be pinned.  No JNI call is made while a pin is live.
int msm_over_arrays(JNIEnv* env, ozk_ctx* ctx, int groups, jbyteArray scalars, jbyteArray b1, jbyteArray b2, size_t n, uint8_t* res) {
    if (n == 0) return groups == 1 ? ozk_msm_g1(ctx, nullptr, nullptr, 0, res) : groups == 2 ? ozk_msm_g2(ctx, nullptr, nullptr, 0, res)
                                                                                             : ozk_msm_g1g2(ctx, nullptr, nullptr, nullptr, 0, res);
    size_t bounds[9];
    const int k = ozk_msm_plan_slices(n, bounds, 8);
    size_t longest = 0;
    for (int i = 0; i < k; i++) longest = bounds[i + 1] - bounds[i] > longest ? bounds[i + 1] - bounds[i] : longest;
    int rc = ozk_msm_begin(ctx, groups, n, longest, nullptr, nullptr, 0);
    if (rc != OZK_OK) return rc;
    bool pin_failed = false;
    for (int i = 0; i < k && rc == OZK_OK && !pin_failed; i++) {
        const size_t lo = bounds[i], len = bounds[i + 1] - bounds[i];
        if (len == 0) continue;
        Pinned s(env, scalars), p1(env, (groups & 1) ? b1 : nullptr), p2(env, (groups & 2) ? b2 : nullptr);
        if (!s.p || ((groups & 1) && !p1.p) || ((groups & 2) && !p2.p)) {
            pin_failed = true;
        } else {
            rc = ozk_msm_feed(ctx, s.bytes() + lo * 32, (groups & 1) ? p1.bytes() + lo * 96 : nullptr,
                              (groups & 2) ? p2.bytes() + lo * 192 : nullptr, len);
        }
    }                                                    // pins are released here, before anything else happens
    if (pin_failed) {
        uint8_t scratch[288];
        ozk_msm_end(ctx, scratch);                       // closes the unfinished stream (reports "fewer pairs fed")
        return 1;
    }
    if (rc != OZK_OK) return rc;
    return ozk_msm_end(ctx, res);
}

// handles of persistent bases handed to Java: a stale or foreign jlong must become a RuntimeException, not a wild pointer
std::mutex g_keys_mu;
std::set<const ozk_bases*> g_keys;
bool key_known(jlong h) {
    std::lock_guard<std::mutex> g(g_keys_mu);
    return g_keys.count((const ozk_bases*)(intptr_t)h) != 0;
}

}  // namespace

extern "C" {

#ifdef OZK_SHIM_VARMSM
// algebra_msm_VariableBaseMSM.h:10-16, reference implementation algebra_msm_VariableBaseMSM.cu:1614-1695
JNIEXPORT jbyteArray JNICALL Java_algebra_msm_VariableBaseMSM_variableBaseSerialMSMNativeHelper(
    JNIEnv* env, jclass, jbyteArray basesXYZ, jbyteArray scalars, jint batch_size, jint type, jint taskID) {
    if (batch_size < 0) return fail_msg(env, "variableBaseSerialMSMNativeHelper: negative batch_size");
    const size_t n = (size_t)batch_size;
    const bool g1 = (type == 1);                     // anything else is G2, like the reference (:1631,:1661)
    if (!long_enough(env, scalars, n * 32) || !long_enough(env, basesXYZ, n * (g1 ? 96 : 192)))
        return fail_msg(env, "variableBaseSerialMSMNativeHelper: input arrays shorter than batch_size elements");
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) return fail(env, "variableBaseSerialMSMNativeHelper");
    uint8_t res[192];
    const int rc = msm_over_arrays(env, ctx, g1 ? 1 : 2, scalars, g1 ? basesXYZ : nullptr, g1 ? nullptr : basesXYZ, n, res);
    if (rc == 1) return fail_msg(env, "variableBaseSerialMSMNativeHelper: could not pin the input arrays");
    if (rc != OZK_OK) return fail(env, "variableBaseSerialMSMNativeHelper");
    std::vector<uint8_t> out(g1 ? 192 : 384);
    widen(res, g1 ? 3 : 6, out.data(), false);
    return to_java(env, out);
}

// algebra_msm_VariableBaseMSM.h:18-24, reference implementation algebra_msm_VariableBaseMSM.cu:1712-1788
JNIEXPORT jbyteArray JNICALL Java_algebra_msm_VariableBaseMSM_variableBaseDoubleMSMNativeHelper(
    JNIEnv* env, jclass, jbyteArray bases1, jbyteArray bases2, jbyteArray scalars, jint batch_size, jint taskID) {
    if (batch_size < 0) return fail_msg(env, "variableBaseDoubleMSMNativeHelper: negative batch_size");
    const size_t n = (size_t)batch_size;
    if (!long_enough(env, scalars, n * 32) || !long_enough(env, bases1, n * 96) || !long_enough(env, bases2, n * 192))
        return fail_msg(env, "variableBaseDoubleMSMNativeHelper: input arrays shorter than batch_size elements");
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) return fail(env, "variableBaseDoubleMSMNativeHelper");
    uint8_t res[288];
    const int rc = msm_over_arrays(env, ctx, 3, scalars, bases1, bases2, n, res);
    if (rc == 1) return fail_msg(env, "variableBaseDoubleMSMNativeHelper: could not pin the input arrays");
    if (rc != OZK_OK) return fail(env, "variableBaseDoubleMSMNativeHelper");
    std::vector<uint8_t> out(576);
    widen(res, 9, out.data(), false);                // G1 (X,Y,Z) || G2 (Xa,Xb,Ya,Yb,Za,Zb), :1781-1784
    return to_java(env, out);
}

// North-star variant: direct ByteBuffers, 32-byte elements in and out.  type 1 = G1, 2 = G2, 3 = paired (bases2 used).
// `out` receives 96 / 192 / 288 bytes.  Returns 0, or throws RuntimeException and returns -1.
JNIEXPORT jint JNICALL Java_algebra_msm_VariableBaseMSM_variableBaseMSMDirect(
    JNIEnv* env, jclass, jobject bases1, jobject bases2, jobject scalars, jint batch_size, jint type, jint taskID, jobject out) {
    const size_t n = batch_size < 0 ? 0 : (size_t)batch_size;
    const uint8_t* s = scalars ? (const uint8_t*)env->GetDirectBufferAddress(scalars) : nullptr;
    const uint8_t* b1 = bases1 ? (const uint8_t*)env->GetDirectBufferAddress(bases1) : nullptr;
    const uint8_t* b2 = bases2 ? (const uint8_t*)env->GetDirectBufferAddress(bases2) : nullptr;
    uint8_t* o = out ? (uint8_t*)env->GetDirectBufferAddress(out) : nullptr;
    const size_t out_need = type == 1 ? 96 : type == 2 ? 192 : 288;
    if (batch_size < 0 || !o || (size_t)env->GetDirectBufferCapacity(out) < out_need ||
        (n && (!s || (size_t)env->GetDirectBufferCapacity(scalars) < n * 32)) ||
        (n && type != 2 && (!b1 || (size_t)env->GetDirectBufferCapacity(bases1) < n * 96)) ||
        (n && type != 1 && (!b2 || (size_t)env->GetDirectBufferCapacity(bases2) < n * 192))) {
        fail_msg(env, "variableBaseMSMDirect: buffers must be direct and hold batch_size elements");
        return -1;
    }
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) { fail(env, "variableBaseMSMDirect"); return -1; }
    int rc = type == 1 ? ozk_msm_g1(ctx, s, b1, n, o) : type == 2 ? ozk_msm_g2(ctx, s, b2, n, o) : ozk_msm_g1g2(ctx, s, b1, b2, n, o);
    if (rc != OZK_OK) { fail(env, "variableBaseMSMDirect"); return -1; }
    return 0;
}
// Persistent proving-key vectors: upload a query vector once, then pass its handle.
// type 1 = G1 (96-byte points), 2 = G2 (192-byte points).  Returns the handle (never 0), or throws and returns 0.
// The handle lives on device taskID % deviceCount; later calls must use a taskID that maps to the same device.
JNIEXPORT jlong JNICALL Java_algebra_msm_VariableBaseMSM_uploadBasesDirect(
    JNIEnv* env, jclass, jobject bases, jint count, jint type, jint taskID) {
    const size_t n = count < 0 ? 0 : (size_t)count;
    const uint8_t* b = bases ? (const uint8_t*)env->GetDirectBufferAddress(bases) : nullptr;
    const size_t pt = type == 1 ? 96 : 192;
    if (count <= 0 || (type != 1 && type != 2) || !b || (size_t)env->GetDirectBufferCapacity(bases) < n * pt) {
        fail_msg(env, "uploadBasesDirect: bases must be a direct buffer of count > 0 points, type 1 (G1) or 2 (G2)");
        return 0;
    }
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) { fail(env, "uploadBasesDirect"); return 0; }
    ozk_bases* key = nullptr;
    int rc = type == 1 ? ozk_bases_upload_g1(ctx, b, n, &key) : ozk_bases_upload_g2(ctx, b, n, &key);
    if (rc != OZK_OK) { fail(env, "uploadBasesDirect"); return 0; }
    {
        std::lock_guard<std::mutex> g(g_keys_mu);
        g_keys.insert(key);
    }
    return (jlong)(intptr_t)key;
}

JNIEXPORT void JNICALL Java_algebra_msm_VariableBaseMSM_freeBases(JNIEnv* env, jclass, jlong key, jint taskID) {
    if (!key) return;
    {
        std::lock_guard<std::mutex> g(g_keys_mu);
        if (g_keys.erase((const ozk_bases*)(intptr_t)key) == 0) {
            fail_msg(env, "freeBases: not a live handle of uploadBasesDirect (stale or already freed)");
            return;
        }
    }
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) { fail(env, "freeBases"); return; }
    ozk_bases_free(ctx, (ozk_bases*)(intptr_t)key);
}

// out = sum_{i < batch_size} scalars[i] * key[first + i]; type 1: key1 is a G1 handle, 2: key2 is a G2 handle, 3: both
// (paired, out = G1 || G2).  Only the scalars cross PCIe.  Returns 0, or throws RuntimeException and returns -1.
JNIEXPORT jint JNICALL Java_algebra_msm_VariableBaseMSM_variableBaseMSMKeyedDirect(
    JNIEnv* env, jclass, jlong key1, jlong key2, jobject scalars, jint first, jint batch_size, jint type, jint taskID, jobject out) {
    const size_t n = batch_size < 0 ? 0 : (size_t)batch_size;
    const uint8_t* s = scalars ? (const uint8_t*)env->GetDirectBufferAddress(scalars) : nullptr;
    uint8_t* o = out ? (uint8_t*)env->GetDirectBufferAddress(out) : nullptr;
    const size_t out_need = type == 1 ? 96 : type == 2 ? 192 : 288;
    if (batch_size < 0 || first < 0 || type < 1 || type > 3 || !o || (size_t)env->GetDirectBufferCapacity(out) < out_need ||
        (n && (!s || (size_t)env->GetDirectBufferCapacity(scalars) < n * 32))) {
        fail_msg(env, "variableBaseMSMKeyedDirect: buffers must be direct and hold batch_size elements");
        return -1;
    }
    if ((type != 2 && !key_known(key1)) || (type != 1 && !key_known(key2))) {
        fail_msg(env, "variableBaseMSMKeyedDirect: not a live handle of uploadBasesDirect (stale or already freed)");
        return -1;
    }
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) { fail(env, "variableBaseMSMKeyedDirect"); return -1; }
    const ozk_bases* k1 = (const ozk_bases*)(intptr_t)key1;
    const ozk_bases* k2 = (const ozk_bases*)(intptr_t)key2;
    int rc = type == 1 ? ozk_msm_g1_keyed(ctx, s, k1, (size_t)first, n, o)
           : type == 2 ? ozk_msm_g2_keyed(ctx, s, k2, (size_t)first, n, o)
                       : ozk_msm_g1g2_keyed(ctx, s, k1, k2, (size_t)first, n, o);
    if (rc != OZK_OK) { fail(env, "variableBaseMSMKeyedDirect"); return -1; }
    return 0;
}
#endif  // OZK_SHIM_VARMSM

#ifdef OZK_SHIM_FIXEDMSM
// algebra_msm_FixedBaseMSM.h:10-16, reference implementation algebra_msm_FixedBaseMSM.cu:1276-1384.
// out_len / inner_len / scalarSize describe the reference's own table and are not needed: the result is defined by
// outerc windows of windowSize bits (algebra_msm_FixedBaseMSM.cu:765-779).
JNIEXPORT jbyteArray JNICALL Java_algebra_msm_FixedBaseMSM_batchMSMNativeHelper(
    JNIEnv* env, jclass, jint outerc, jint windowSize, jint out_len, jint inner_len, jint batch_size, jint scalarSize,
    jbyteArray base, jbyteArray scalars, jint BNType, jint taskID) {
    (void)out_len; (void)inner_len; (void)scalarSize;
    if (batch_size < 0) return fail_msg(env, "batchMSMNativeHelper: negative batch_size");
    const size_t n = (size_t)batch_size;
    const bool g1 = (BNType == 1);
    const size_t pt = g1 ? 96 : 192;
    if (!long_enough(env, scalars, n * 32) || !long_enough(env, base, pt))
        return fail_msg(env, "batchMSMNativeHelper: input arrays too short");
    if (n * pt * 2 > 0x7fffffffu) return fail_msg(env, "batchMSMNativeHelper: result exceeds the 2 GiB limit of a Java byte[]");
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) return fail(env, "batchMSMNativeHelper");
    std::vector<uint8_t> res(n * pt), sc(n * 32);
    uint8_t b[192];
    copy_out(env, base, pt, b);
    copy_out(env, scalars, n * 32, sc.data());
    const int rc = g1 ? ozk_fixed_g1(ctx, b, sc.data(), n, outerc, windowSize, res.data())
                      : ozk_fixed_g2(ctx, b, sc.data(), n, outerc, windowSize, res.data());
    if (rc != OZK_OK) return fail(env, "batchMSMNativeHelper");
    std::vector<uint8_t> out(n * pt * 2);
    widen(res.data(), n * (g1 ? 3 : 6), out.data(), true);       // 64-byte big-endian coordinates, :783-787
    return to_java(env, out);
}

// algebra_msm_FixedBaseMSM.h:18-24, reference implementation algebra_msm_FixedBaseMSM.cu:1395-1491:
// element i of the result = G1 (X,Y,Z) || G2 (Xa,Xb,Ya,Yb,Za,Zb), 9 x 64 bytes big-endian.
JNIEXPORT jbyteArray JNICALL Java_algebra_msm_FixedBaseMSM_doubleBatchMSMNativeHelper(
    JNIEnv* env, jclass, jint outerc1, jint windowSize1, jint outerc2, jint windowSize2, jint out_len1, jint inner_len1,
    jint out_len2, jint inner_len2, jint batch_size, jbyteArray baseG1, jbyteArray baseG2, jbyteArray scalars, jint taskID) {
    (void)out_len1; (void)inner_len1; (void)out_len2; (void)inner_len2;
    if (batch_size < 0) return fail_msg(env, "doubleBatchMSMNativeHelper: negative batch_size");
    const size_t n = (size_t)batch_size;
    if (!long_enough(env, scalars, n * 32) || !long_enough(env, baseG1, 96) || !long_enough(env, baseG2, 192))
        return fail_msg(env, "doubleBatchMSMNativeHelper: input arrays too short");
    if (n * 576 > 0x7fffffffu) return fail_msg(env, "doubleBatchMSMNativeHelper: result exceeds the 2 GiB limit of a Java byte[]");
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) return fail(env, "doubleBatchMSMNativeHelper");
    std::vector<uint8_t> r1(n * 96), r2(n * 192), sc(n * 32);
    uint8_t b1[96], b2[192];
    copy_out(env, baseG1, 96, b1);
    copy_out(env, baseG2, 192, b2);
    copy_out(env, scalars, n * 32, sc.data());
    int rc = ozk_fixed_g1(ctx, b1, sc.data(), n, outerc1, windowSize1, r1.data());
    if (rc == OZK_OK) rc = ozk_fixed_g2(ctx, b2, sc.data(), n, outerc2, windowSize2, r2.data());
    if (rc != OZK_OK) return fail(env, "doubleBatchMSMNativeHelper");
    std::vector<uint8_t> out(n * 576);
    for (size_t i = 0; i < n; i++) {
        widen(r1.data() + 96 * i, 3, out.data() + 576 * i, true);
        widen(r2.data() + 192 * i, 6, out.data() + 576 * i + 192, true);
    }
    return to_java(env, out);
}

// algebra_msm_FixedBaseMSM.h:26-32, reference implementation algebra_msm_FixedBaseMSM.cu:1500-1558: the input holds
// batch_size scalars followed by the multiplier (:1520-1522); the result is batch_size x 64 bytes big-endian.
JNIEXPORT jbyteArray JNICALL Java_algebra_msm_FixedBaseMSM_fieldBatchMSMNativeHelper(
    JNIEnv* env, jclass, jbyteArray scalarsPlusBase, jint batch_size, jint taskID) {
    if (batch_size < 0) return fail_msg(env, "fieldBatchMSMNativeHelper: negative batch_size");
    const size_t n = (size_t)batch_size;
    if (!long_enough(env, scalarsPlusBase, (n + 1) * 32)) return fail_msg(env, "fieldBatchMSMNativeHelper: input array too short");
    if (n * 64 > 0x7fffffffu) return fail_msg(env, "fieldBatchMSMNativeHelper: result exceeds the 2 GiB limit of a Java byte[]");
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) return fail(env, "fieldBatchMSMNativeHelper");
    std::vector<uint8_t> res(n * 32), in((n + 1) * 32);
    copy_out(env, scalarsPlusBase, (n + 1) * 32, in.data());
    const int rc = ozk_fr_scale(ctx, in.data(), n, in.data() + n * 32, res.data());
    if (rc != OZK_OK) return fail(env, "fieldBatchMSMNativeHelper");
    std::vector<uint8_t> out(n * 64);
    widen(res.data(), n, out.data(), true);
    return to_java(env, out);
}

// North-star variant of batchMSM: direct buffers, 32-byte coordinates, BNType 1 = G1, 2 = G2; out holds n points.
JNIEXPORT jint JNICALL Java_algebra_msm_FixedBaseMSM_batchMSMDirect(
    JNIEnv* env, jclass, jint outerc, jint windowSize, jint batch_size, jobject base, jobject scalars, jint BNType, jint taskID, jobject out) {
    const size_t n = batch_size < 0 ? 0 : (size_t)batch_size;
    const size_t pt = BNType == 1 ? 96 : 192;
    const uint8_t* b = base ? (const uint8_t*)env->GetDirectBufferAddress(base) : nullptr;
    const uint8_t* s = scalars ? (const uint8_t*)env->GetDirectBufferAddress(scalars) : nullptr;
    uint8_t* o = out ? (uint8_t*)env->GetDirectBufferAddress(out) : nullptr;
    if (batch_size < 0 || !b || (size_t)env->GetDirectBufferCapacity(base) < pt ||
        (n && (!s || !o || (size_t)env->GetDirectBufferCapacity(scalars) < n * 32 || (size_t)env->GetDirectBufferCapacity(out) < n * pt))) {
        fail_msg(env, "batchMSMDirect: buffers must be direct and hold batch_size elements");
        return -1;
    }
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) { fail(env, "batchMSMDirect"); return -1; }
    int rc = BNType == 1 ? ozk_fixed_g1(ctx, b, s, n, outerc, windowSize, o) : ozk_fixed_g2(ctx, b, s, n, outerc, windowSize, o);
    if (rc != OZK_OK) { fail(env, "batchMSMDirect"); return -1; }
    return 0;
}
#endif  // OZK_SHIM_FIXEDMSM

#ifdef OZK_SHIM_FFT
// algebra_fft_FFTAuxiliary.h:10-16, reference implementation algebra_fft_FFTAuxiliary.cu:219-260.  `input` is a
// java.util.List<byte[]>; every element is a little-endian array whose length is a multiple of 4 and at most 32
// (FFTAuxiliary.java:41-51); the result is n x 64 bytes little-endian (:252-255).  Dormant in the reference's Java
// (the call is commented out, FFTAuxiliary.java:72-97) but part of the library's interface.
JNIEXPORT jbyteArray JNICALL Java_algebra_fft_FFTAuxiliary_serialRadix2FFTNativeHelper(
    JNIEnv* env, jclass, jobject input, jbyteArray omega, jint taskID) {
    if (!input || !omega) return fail_msg(env, "serialRadix2FFTNativeHelper: null argument");
    jclass list_cls = env->FindClass("java/util/List");
    if (!list_cls) return nullptr;
    jmethodID m_size = env->GetMethodID(list_cls, "size", "()I");
    jmethodID m_get = env->GetMethodID(list_cls, "get", "(I)Ljava/lang/Object;");
    if (!m_size || !m_get) return nullptr;
    const jint n = env->CallIntMethod(input, m_size);
    if (n < 0) return fail_msg(env, "serialRadix2FFTNativeHelper: negative list size");
    if ((size_t)n * 64 > 0x7fffffffu) return fail_msg(env, "serialRadix2FFTNativeHelper: result exceeds the 2 GiB limit of a Java byte[]");
    std::vector<uint8_t> data((size_t)n * 32, 0);
    for (jint i = 0; i < n; i++) {
        jbyteArray e = (jbyteArray)env->CallObjectMethod(input, m_get, i);
        if (!e) return fail_msg(env, "serialRadix2FFTNativeHelper: null list element");
        jsize len = env->GetArrayLength(e);
        if (len > 32) {
            // BigInteger.toByteArray may add a sign byte: the padded array is then 36 bytes with zeros above byte 31
            len = 32;
        }
        env->GetByteArrayRegion(e, 0, len, (jbyte*)(data.data() + 32 * (size_t)i));
        env->DeleteLocalRef(e);
    }
    uint8_t w[32] = {0};
    jsize wl = env->GetArrayLength(omega);
    env->GetByteArrayRegion(omega, 0, wl > 32 ? 32 : wl, (jbyte*)w);
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) return fail(env, "serialRadix2FFTNativeHelper");
    if (n > 0 && ozk_ntt_fr(ctx, data.data(), (size_t)n, w) != OZK_OK) return fail(env, "serialRadix2FFTNativeHelper");
    std::vector<uint8_t> out((size_t)n * 64);
    widen(data.data(), (size_t)n, out.data(), false);
    return to_java(env, out);
}

// Bulk variant the reference's own TODOs ask for (algebra_fft_FFTAuxiliary.cu:226,251): one direct ByteBuffer of
// n x 32 bytes, transformed in place.
JNIEXPORT jint JNICALL Java_algebra_fft_FFTAuxiliary_serialRadix2FFTDirect(
    JNIEnv* env, jclass, jobject data, jint n, jbyteArray omega, jint taskID) {
    uint8_t* d = data ? (uint8_t*)env->GetDirectBufferAddress(data) : nullptr;
    if (n < 0 || !omega || (n && (!d || (size_t)env->GetDirectBufferCapacity(data) < (size_t)n * 32))) {
        fail_msg(env, "serialRadix2FFTDirect: data must be a direct buffer of n 32-byte elements");
        return -1;
    }
    uint8_t w[32] = {0};
    jsize wl = env->GetArrayLength(omega);
    env->GetByteArrayRegion(omega, 0, wl > 32 ? 32 : wl, (jbyte*)w);
    ozk_ctx* ctx = context_for_task(taskID);
    if (!ctx) { fail(env, "serialRadix2FFTDirect"); return -1; }
    if (n > 0 && ozk_ntt_fr(ctx, d, (size_t)n, w) != OZK_OK) { fail(env, "serialRadix2FFTDirect"); return -1; }
    return 0;
}
#endif  // OZK_SHIM_FFT

}  // extern "C"
