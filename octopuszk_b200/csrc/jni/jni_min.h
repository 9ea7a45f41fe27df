// Minimal, self-written subset of the Java Native Interface for building the shim libraries without a JDK (this
// image has no jni.h).  Layout follows the JNI specification: JNIEnv is a pointer to a pointer to the function table;
// the table starts with four reserved slots and lists the functions in specification order.  Only the slots the shims
// use are typed; their indices are the specification's and are asserted in tests/fake_jni.cc.
// With a real JDK this header can be replaced by <jni.h> without touching jni_shim.cc.
#ifndef OZK_JNI_MIN_H
#define OZK_JNI_MIN_H

#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_OK 0

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef jint jsize;

struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jthrowable;
typedef jobject jarray;
typedef jarray jbyteArray;
struct _jmethodID;
typedef struct _jmethodID* jmethodID;

struct JNIEnv_;
typedef JNIEnv_ JNIEnv;

enum {
    OZK_JNI_FindClass = 6,
    OZK_JNI_ThrowNew = 14,
    OZK_JNI_DeleteLocalRef = 23,
    OZK_JNI_GetMethodID = 33,
    OZK_JNI_CallObjectMethod = 34,
    OZK_JNI_CallIntMethod = 49,
    OZK_JNI_GetArrayLength = 171,
    OZK_JNI_NewByteArray = 176,
    OZK_JNI_GetByteArrayRegion = 200,
    OZK_JNI_SetByteArrayRegion = 208,
    OZK_JNI_GetPrimitiveArrayCritical = 222,
    OZK_JNI_ReleasePrimitiveArrayCritical = 223,
    OZK_JNI_ExceptionCheck = 228,
    OZK_JNI_GetDirectBufferAddress = 230,
    OZK_JNI_GetDirectBufferCapacity = 231,
    OZK_JNI_TABLE_SLOTS = 235
};

struct JNINativeInterface_ {
    void* slot[OZK_JNI_TABLE_SLOTS];
};

struct JNIEnv_ {
    const JNINativeInterface_* functions;

    template <class Fn>
    Fn fn(int idx) const { return reinterpret_cast<Fn>(functions->slot[idx]); }

    jclass FindClass(const char* name) { return fn<jclass (*)(JNIEnv*, const char*)>(OZK_JNI_FindClass)(this, name); }
    jint ThrowNew(jclass c, const char* msg) { return fn<jint (*)(JNIEnv*, jclass, const char*)>(OZK_JNI_ThrowNew)(this, c, msg); }
    void DeleteLocalRef(jobject o) { fn<void (*)(JNIEnv*, jobject)>(OZK_JNI_DeleteLocalRef)(this, o); }
    jmethodID GetMethodID(jclass c, const char* name, const char* sig) {
        return fn<jmethodID (*)(JNIEnv*, jclass, const char*, const char*)>(OZK_JNI_GetMethodID)(this, c, name, sig);
    }
    jobject CallObjectMethod(jobject o, jmethodID m, jint arg) {
        return fn<jobject (*)(JNIEnv*, jobject, jmethodID, ...)>(OZK_JNI_CallObjectMethod)(this, o, m, arg);
    }
    jint CallIntMethod(jobject o, jmethodID m) { return fn<jint (*)(JNIEnv*, jobject, jmethodID, ...)>(OZK_JNI_CallIntMethod)(this, o, m); }
    jsize GetArrayLength(jarray a) { return fn<jsize (*)(JNIEnv*, jarray)>(OZK_JNI_GetArrayLength)(this, a); }
    jbyteArray NewByteArray(jsize n) { return fn<jbyteArray (*)(JNIEnv*, jsize)>(OZK_JNI_NewByteArray)(this, n); }
    void GetByteArrayRegion(jbyteArray a, jsize start, jsize len, jbyte* buf) {
        fn<void (*)(JNIEnv*, jbyteArray, jsize, jsize, jbyte*)>(OZK_JNI_GetByteArrayRegion)(this, a, start, len, buf);
    }
    void SetByteArrayRegion(jbyteArray a, jsize start, jsize len, const jbyte* buf) {
        fn<void (*)(JNIEnv*, jbyteArray, jsize, jsize, const jbyte*)>(OZK_JNI_SetByteArrayRegion)(this, a, start, len, buf);
    }
    void* GetPrimitiveArrayCritical(jarray a, jboolean* is_copy) {
        return fn<void* (*)(JNIEnv*, jarray, jboolean*)>(OZK_JNI_GetPrimitiveArrayCritical)(this, a, is_copy);
    }
    void ReleasePrimitiveArrayCritical(jarray a, void* p, jint mode) {
        fn<void (*)(JNIEnv*, jarray, void*, jint)>(OZK_JNI_ReleasePrimitiveArrayCritical)(this, a, p, mode);
    }
    jboolean ExceptionCheck() { return fn<jboolean (*)(JNIEnv*)>(OZK_JNI_ExceptionCheck)(this); }
    void* GetDirectBufferAddress(jobject b) { return fn<void* (*)(JNIEnv*, jobject)>(OZK_JNI_GetDirectBufferAddress)(this, b); }
    jlong GetDirectBufferCapacity(jobject b) { return fn<jlong (*)(JNIEnv*, jobject)>(OZK_JNI_GetDirectBufferCapacity)(this, b); }
};

#endif
