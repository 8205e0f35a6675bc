// fake_jni.cpp -- a JNIEnv with just the three array functions, to drive libnativedsp.so's
// exported JNI symbols from a test process that has no JVM (test support).
#include <cstddef>
#include <cstring>

#include "../../rfanalyzer_b200/csrc/jni_min/jni.h"

namespace {
struct FakeArray {
    jsize length;
    jfloat *data;
};
jsize get_len(JNIEnv *, jarray a) { return reinterpret_cast<FakeArray *>(a)->length; }
void get_region(JNIEnv *, jfloatArray a, jsize s, jsize l, jfloat *b) {
    memcpy(b, reinterpret_cast<FakeArray *>(a)->data + s, sizeof(jfloat) * (size_t)l);
}
void set_region(JNIEnv *, jfloatArray a, jsize s, jsize l, const jfloat *b) {
    memcpy(reinterpret_cast<FakeArray *>(a)->data + s, b, sizeof(jfloat) * (size_t)l);
}
}  // namespace

typedef void (*jni_fn)(JNIEnv *, jobject, jfloatArray, jfloatArray);

extern "C" void fake_jni_call(void *fn, float *in, int in_len, float *out, int out_len) {
    static JNINativeInterface_ table;
    table.GetArrayLength = get_len;
    table.GetFloatArrayRegion = get_region;
    table.SetFloatArrayRegion = set_region;
    JNIEnv env;
    env.functions = &table;
    FakeArray a{in_len, in}, b{out_len, out};
    reinterpret_cast<jni_fn>(fn)(&env, nullptr, reinterpret_cast<jfloatArray>(&a), reinterpret_cast<jfloatArray>(&b));
}

extern "C" int fake_jni_slot_offsets(int *get_len_slot, int *get_region_slot, int *set_region_slot) {
    *get_len_slot = (int)(offsetof(JNINativeInterface_, GetArrayLength) / sizeof(void *));
    *get_region_slot = (int)(offsetof(JNINativeInterface_, GetFloatArrayRegion) / sizeof(void *));
    *set_region_slot = (int)(offsetof(JNINativeInterface_, SetFloatArrayRegion) / sizeof(void *));
    return (int)(sizeof(JNINativeInterface_) / sizeof(void *));
}
