/*
 * Minimal stand-in for <jni.h> -- TEST INFRASTRUCTURE ONLY.
 * Just enough of the JNI C++ surface for the reference's
 * nativedsp/src/main/cpp/nativedsp.cpp (:19-81) to compile unmodified, in place,
 * on a box with no JDK/NDK: jfloatArray is a {length, data} pair and the three
 * JNIEnv members the file calls are plain memcpy's.
 */
#ifndef RFA_ORACLE_STUB_JNI_H
#define RFA_ORACLE_STUB_JNI_H
#include <string.h>

typedef int jint;
typedef int jsize;
typedef float jfloat;
struct _jobject {};
typedef _jobject *jobject;
struct _jfloatArray {
    jsize length;
    jfloat *data;
};
typedef _jfloatArray *jfloatArray;

struct JNIEnv {
    jsize GetArrayLength(jfloatArray a) { return a->length; }
    void GetFloatArrayRegion(jfloatArray a, jsize start, jsize len, jfloat *buf) {
        memcpy(buf, a->data + start, sizeof(jfloat) * (size_t)len);
    }
    void SetFloatArrayRegion(jfloatArray a, jsize start, jsize len, const jfloat *buf) {
        memcpy(a->data + start, buf, sizeof(jfloat) * (size_t)len);
    }
};

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#endif
