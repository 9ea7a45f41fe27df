// A fake JNIEnv for driving the JNI shims without a JVM (none exists in this image).
// Implements the handful of function-table slots jni_shim.cc uses over plain C++ objects and exposes helpers to
// Python (ctypes) for building byte[] / List<byte[]> / direct ByteBuffer arguments and reading results back.
// Test scaffolding only.
#include <cstdarg>
#include <cstdio>
#include <atomic>
#include <cstring>
#include <string>
#include <vector>

#include "../octopuszk_b200/csrc/jni/jni_min.h"

namespace {

enum Kind { K_BYTES = 1, K_LIST = 2, K_CLASS = 3, K_DIRECT = 4 };

struct Obj {
    int kind;
    std::vector<uint8_t> bytes;       // K_BYTES / K_DIRECT
    std::vector<Obj*> items;          // K_LIST
    std::string name;                 // K_CLASS
    int pins = 0;
};

std::string g_exception;
std::atomic<int> g_live_pins{0};   // natives are entered from several threads at once (test_concurrent_executor_threads)
// JNI rule: between GetPrimitiveArrayCritical and ReleasePrimitiveArrayCritical a thread must not call any other JNI function.
// Every fake function below except those two calls not_in_critical(); violations are counted for the tests.
thread_local int t_pins = 0;
std::atomic<int> g_critical_violations{0};
std::atomic<int> g_pin_calls{0};
std::atomic<int> g_fail_next_pin{0};   // test hook: the next GetPrimitiveArrayCritical calls return NULL
void not_in_critical() {
    if (t_pins > 0) g_critical_violations++;
}
int g_method_size = 1, g_method_get = 2;

Obj* O(jobject o) { return reinterpret_cast<Obj*>(o); }
jobject J(Obj* o) { return reinterpret_cast<jobject>(o); }

jclass FindClass(JNIEnv*, const char* name) {
    not_in_critical();
    Obj* c = new Obj{K_CLASS};
    c->name = name;
    return J(c);
}
jint ThrowNew(JNIEnv*, jclass c, const char* msg) {
    not_in_critical();
    g_exception = O(c)->name + ": " + msg;
    return 0;
}
void DeleteLocalRef(JNIEnv*, jobject) { not_in_critical(); }
jmethodID GetMethodID(JNIEnv*, jclass, const char* name, const char*) {
    not_in_critical();
    if (!strcmp(name, "size")) return reinterpret_cast<jmethodID>(&g_method_size);
    if (!strcmp(name, "get")) return reinterpret_cast<jmethodID>(&g_method_get);
    return nullptr;
}
jobject CallObjectMethod(JNIEnv*, jobject o, jmethodID m, ...) {
    not_in_critical();
    va_list ap;
    va_start(ap, m);
    jint idx = va_arg(ap, jint);
    va_end(ap);
    if (m != reinterpret_cast<jmethodID>(&g_method_get) || O(o)->kind != K_LIST) return nullptr;
    if (idx < 0 || (size_t)idx >= O(o)->items.size()) return nullptr;
    return J(O(o)->items[idx]);
}
jint CallIntMethod(JNIEnv*, jobject o, jmethodID m, ...) {
    not_in_critical();
    if (m != reinterpret_cast<jmethodID>(&g_method_size) || O(o)->kind != K_LIST) return -1;
    return (jint)O(o)->items.size();
}
jsize GetArrayLength(JNIEnv*, jarray a) {
    not_in_critical();
    return (jsize)O(a)->bytes.size();
}
jbyteArray NewByteArray(JNIEnv*, jsize n) {
    not_in_critical();
    Obj* o = new Obj{K_BYTES};
    o->bytes.resize((size_t)n);
    return J(o);
}
void GetByteArrayRegion(JNIEnv*, jbyteArray a, jsize start, jsize len, jbyte* buf) {
    not_in_critical();
    memcpy(buf, O(a)->bytes.data() + start, (size_t)len);
}
void SetByteArrayRegion(JNIEnv*, jbyteArray a, jsize start, jsize len, const jbyte* buf) {
    not_in_critical();
    memcpy(O(a)->bytes.data() + start, buf, (size_t)len);
}
void* GetPrimitiveArrayCritical(JNIEnv*, jarray a, jboolean* is_copy) {
    if (is_copy) *is_copy = 0;
    g_pin_calls++;
    if (g_fail_next_pin > 0) {
        g_fail_next_pin--;
        return nullptr;
    }
    O(a)->pins++;
    g_live_pins++;
    t_pins++;
    return O(a)->bytes.data();
}
void ReleasePrimitiveArrayCritical(JNIEnv*, jarray a, void*, jint) {
    O(a)->pins--;
    g_live_pins--;
    t_pins--;
}
jboolean ExceptionCheck(JNIEnv*) {
    not_in_critical();
    return g_exception.empty() ? 0 : 1;
}
void* GetDirectBufferAddress(JNIEnv*, jobject b) {
    not_in_critical();
    return O(b)->kind == K_DIRECT ? O(b)->bytes.data() : nullptr;
}
jlong GetDirectBufferCapacity(JNIEnv*, jobject b) {
    not_in_critical();
    return O(b)->kind == K_DIRECT ? (jlong)O(b)->bytes.size() : -1;
}

JNINativeInterface_ g_table;
JNIEnv_ g_env;

}  // namespace

extern "C" {

#define EXPORT __attribute__((visibility("default")))

EXPORT void* fj_env() {
    static bool init = false;
    if (!init) {
        memset(&g_table, 0, sizeof g_table);
        // slot indices are the JNI specification's
        static_assert(OZK_JNI_FindClass == 6 && OZK_JNI_ThrowNew == 14 && OZK_JNI_GetMethodID == 33 && OZK_JNI_CallObjectMethod == 34 &&
                          OZK_JNI_CallIntMethod == 49 && OZK_JNI_GetArrayLength == 171 && OZK_JNI_NewByteArray == 176 &&
                          OZK_JNI_GetByteArrayRegion == 200 && OZK_JNI_SetByteArrayRegion == 208 &&
                          OZK_JNI_GetPrimitiveArrayCritical == 222 && OZK_JNI_ReleasePrimitiveArrayCritical == 223 &&
                          OZK_JNI_ExceptionCheck == 228 && OZK_JNI_GetDirectBufferAddress == 230 && OZK_JNI_GetDirectBufferCapacity == 231,
                      "JNI function-table indices");
        g_table.slot[OZK_JNI_FindClass] = (void*)FindClass;
        g_table.slot[OZK_JNI_ThrowNew] = (void*)ThrowNew;
        g_table.slot[OZK_JNI_DeleteLocalRef] = (void*)DeleteLocalRef;
        g_table.slot[OZK_JNI_GetMethodID] = (void*)GetMethodID;
        g_table.slot[OZK_JNI_CallObjectMethod] = (void*)CallObjectMethod;
        g_table.slot[OZK_JNI_CallIntMethod] = (void*)CallIntMethod;
        g_table.slot[OZK_JNI_GetArrayLength] = (void*)GetArrayLength;
        g_table.slot[OZK_JNI_NewByteArray] = (void*)NewByteArray;
        g_table.slot[OZK_JNI_GetByteArrayRegion] = (void*)GetByteArrayRegion;
        g_table.slot[OZK_JNI_SetByteArrayRegion] = (void*)SetByteArrayRegion;
        g_table.slot[OZK_JNI_GetPrimitiveArrayCritical] = (void*)GetPrimitiveArrayCritical;
        g_table.slot[OZK_JNI_ReleasePrimitiveArrayCritical] = (void*)ReleasePrimitiveArrayCritical;
        g_table.slot[OZK_JNI_ExceptionCheck] = (void*)ExceptionCheck;
        g_table.slot[OZK_JNI_GetDirectBufferAddress] = (void*)GetDirectBufferAddress;
        g_table.slot[OZK_JNI_GetDirectBufferCapacity] = (void*)GetDirectBufferCapacity;
        g_env.functions = &g_table;
        init = true;
    }
    return &g_env;
}
EXPORT void* fj_new_bytes(const void* data, size_t len) {
    Obj* o = new Obj{K_BYTES};
    o->bytes.assign((const uint8_t*)data, (const uint8_t*)data + len);
    return o;
}
EXPORT void* fj_new_direct(const void* data, size_t len) {
    Obj* o = new Obj{K_DIRECT};
    o->bytes.resize(len);
    if (data) memcpy(o->bytes.data(), data, len);
    return o;
}
EXPORT void* fj_new_list() { return new Obj{K_LIST}; }
EXPORT void fj_list_add(void* list, void* item) { O((jobject)list)->items.push_back((Obj*)item); }
EXPORT size_t fj_len(void* o) { return ((Obj*)o)->bytes.size(); }
EXPORT const void* fj_data(void* o) { return ((Obj*)o)->bytes.data(); }
EXPORT void fj_free(void* o) {
    Obj* p = (Obj*)o;
    if (!p) return;
    if (p->kind == K_LIST)
        for (Obj* i : p->items) delete i;
    delete p;
}
EXPORT const char* fj_exception() { return g_exception.empty() ? nullptr : g_exception.c_str(); }
EXPORT void fj_clear_exception() { g_exception.clear(); }
EXPORT int fj_live_pins() { return g_live_pins; }
EXPORT int fj_critical_violations() { return g_critical_violations; }
EXPORT int fj_pin_calls() { return g_pin_calls; }
EXPORT void fj_fail_next_pins(int k) { g_fail_next_pin = k; }
}
