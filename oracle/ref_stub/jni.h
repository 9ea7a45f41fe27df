// Stub <jni.h> for compiling the reference's three .cu files where they lie under /root/reference, without a JDK, into
// oracle/_ref/ (test infrastructure; see oracle/Makefile.ref).  It declares exactly the JNI names those files use
// (algebra_msm_VariableBaseMSM.cu:1614-1790, algebra_msm_FixedBaseMSM.cu:1276-1560, algebra_fft_FFTAuxiliary.cu:219-260)
// over a trivial in-memory object model, so oracle/ref_driver.cc can call the reference's own Java_* entry points.
#ifndef OZK_REF_STUB_JNI_H
#define OZK_REF_STUB_JNI_H
#include <cstdint>
#include <cstring>
#include <vector>

#define JNIEXPORT
#define JNICALL

typedef int32_t jint;
typedef int64_t jlong;
typedef signed char jbyte;
typedef unsigned char jboolean;
typedef jint jsize;

struct _jobject {
    std::vector<jbyte> bytes;          // byte[]
    std::vector<_jobject*> items;      // java.util.List<byte[]>
};
typedef _jobject* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jobject jbyteArray;
typedef jobject jobjectArray;
struct _jmethodID { int id; };
typedef _jmethodID* jmethodID;

struct JNIEnv_ {
    jobject NewGlobalRef(jobject o) { return o; }
    jclass FindClass(const char*) { static _jobject cls; return &cls; }
    jmethodID GetMethodID(jclass, const char* name, const char*) {
        static _jmethodID size_id{1}, get_id{2};
        return std::strcmp(name, "size") == 0 ? &size_id : &get_id;
    }
    jint CallIntMethod(jobject o, jmethodID, ...) { return (jint)o->items.size(); }
    jobject CallObjectMethod(jobject o, jmethodID, jint i) { return o->items[(size_t)i]; }
    jbyte* GetByteArrayElements(jbyteArray a, jboolean*) { return a->bytes.data(); }
    jsize GetArrayLength(jarray a) { return (jsize)a->bytes.size(); }
    jbyteArray NewByteArray(jsize n) {
        _jobject* o = new _jobject();
        o->bytes.resize((size_t)n);
        return o;
    }
    void SetByteArrayRegion(jbyteArray a, jsize start, jsize len, const jbyte* buf) { std::memcpy(a->bytes.data() + start, buf, (size_t)len); }
};
typedef JNIEnv_ JNIEnv;
#endif
