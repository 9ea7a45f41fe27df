// C entry points around the reference's own JNI functions, compiled unmodified from /root/reference into oracle/_ref/
// (see oracle/Makefile.ref).  Test infrastructure: lets the GPU tests compare liboctozk's results with the reference's
// own CUDA implementation running on the same B200, and bench tools time it.  Never linked into the product.
#include "jni.h"

extern "C" {
jbyteArray Java_algebra_msm_VariableBaseMSM_variableBaseSerialMSMNativeHelper(JNIEnv*, jclass, jbyteArray, jbyteArray, jint, jint, jint);
jbyteArray Java_algebra_msm_VariableBaseMSM_variableBaseDoubleMSMNativeHelper(JNIEnv*, jclass, jbyteArray, jbyteArray, jbyteArray, jint, jint);
jbyteArray Java_algebra_msm_FixedBaseMSM_batchMSMNativeHelper(JNIEnv*, jclass, jint, jint, jint, jint, jint, jint, jbyteArray, jbyteArray, jint, jint);
jbyteArray Java_algebra_msm_FixedBaseMSM_doubleBatchMSMNativeHelper(JNIEnv*, jclass, jint, jint, jint, jint, jint, jint, jint, jint, jint, jbyteArray,
                                                                   jbyteArray, jbyteArray, jint);
jbyteArray Java_algebra_msm_FixedBaseMSM_fieldBatchMSMNativeHelper(JNIEnv*, jclass, jbyteArray, jint, jint);
jbyteArray Java_algebra_fft_FFTAuxiliary_serialRadix2FFTNativeHelper(JNIEnv*, jclass, jobject, jbyteArray, jint);
}

static jbyteArray mk(const void* p, size_t n) {
    _jobject* o = new _jobject();
    o->bytes.assign((const jbyte*)p, (const jbyte*)p + n);
    return o;
}
static long take(jbyteArray r, void* out, size_t cap) {
    if (!r) return -1;
    size_t n = r->bytes.size();
    if (n > cap) n = cap;
    std::memcpy(out, r->bytes.data(), n);
    long full = (long)r->bytes.size();
    delete r;
    return full;
}

extern "C" {
#define EXPORT __attribute__((visibility("default")))

#ifdef REF_VARMSM
EXPORT long ref_var_msm(const void* bases, size_t bases_len, const void* scalars, size_t scalars_len, int n, int type, void* out, size_t cap) {
    JNIEnv env;
    jbyteArray b = mk(bases, bases_len), s = mk(scalars, scalars_len);
    long r = take(Java_algebra_msm_VariableBaseMSM_variableBaseSerialMSMNativeHelper(&env, nullptr, b, s, n, type, 0), out, cap);
    delete b; delete s;
    return r;
}
EXPORT long ref_var_double_msm(const void* bases1, size_t l1, const void* bases2, size_t l2, const void* scalars, size_t ls, int n, void* out, size_t cap) {
    JNIEnv env;
    jbyteArray b1 = mk(bases1, l1), b2 = mk(bases2, l2), s = mk(scalars, ls);
    long r = take(Java_algebra_msm_VariableBaseMSM_variableBaseDoubleMSMNativeHelper(&env, nullptr, b1, b2, s, n, 0), out, cap);
    delete b1; delete b2; delete s;
    return r;
}
#endif
#ifdef REF_FIXEDMSM
EXPORT long ref_fixed_batch(int outerc, int window, int out_len, int inner_len, int n, int scalar_size, const void* base, size_t base_len,
                            const void* scalars, size_t ls, int bn_type, void* out, size_t cap) {
    JNIEnv env;
    jbyteArray b = mk(base, base_len), s = mk(scalars, ls);
    long r = take(Java_algebra_msm_FixedBaseMSM_batchMSMNativeHelper(&env, nullptr, outerc, window, out_len, inner_len, n, scalar_size, b, s, bn_type, 0), out, cap);
    delete b; delete s;
    return r;
}
EXPORT long ref_fixed_double_batch(int outerc1, int window1, int outerc2, int window2, int out_len1, int inner_len1, int out_len2, int inner_len2,
                                   int n, const void* base1, size_t l1, const void* base2, size_t l2, const void* scalars, size_t ls, void* out,
                                   size_t cap) {
    JNIEnv env;
    jbyteArray b1 = mk(base1, l1), b2 = mk(base2, l2), s = mk(scalars, ls);
    long r = take(Java_algebra_msm_FixedBaseMSM_doubleBatchMSMNativeHelper(&env, nullptr, outerc1, window1, outerc2, window2, out_len1, inner_len1,
                                                                         out_len2, inner_len2, n, b1, b2, s, 0), out, cap);
    delete b1; delete b2; delete s;
    return r;
}
EXPORT long ref_field_batch(const void* scalars_plus_base, size_t len, int n, void* out, size_t cap) {
    JNIEnv env;
    jbyteArray a = mk(scalars_plus_base, len);
    long r = take(Java_algebra_msm_FixedBaseMSM_fieldBatchMSMNativeHelper(&env, nullptr, a, n, 0), out, cap);
    delete a;
    return r;
}
#endif
#ifdef REF_FFT
EXPORT long ref_fft(const void* data, int n, int elem_bytes, const void* omega, int omega_bytes, void* out, size_t cap) {
    JNIEnv env;
    _jobject* list = new _jobject();
    for (int i = 0; i < n; i++) list->items.push_back(mk((const char*)data + (size_t)i * elem_bytes, (size_t)elem_bytes));
    jbyteArray w = mk(omega, (size_t)omega_bytes);
    long r = take(Java_algebra_fft_FFTAuxiliary_serialRadix2FFTNativeHelper(&env, nullptr, list, w, 0), out, cap);
    for (auto* it : list->items) delete it;
    delete list; delete w;
    return r;
}
#endif
}
