/*
 * jni_min/jni.h -- the subset of the Java Native Interface that the nativedsp shim needs,
 * laid out per the JNI specification's function table (slot numbers are normative: 171
 * GetArrayLength, 205 GetFloatArrayRegion, 213 SetFloatArrayRegion).  Used only when no JDK/NDK
 * <jni.h> is on the include path (this build box has neither); with a real <jni.h> the shim
 * compiles against that instead.  Binary compatible with both HotSpot and ART.
 */
#ifndef RFA_JNI_MIN_H
#define RFA_JNI_MIN_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t jint;
typedef jint jsize;
typedef float jfloat;
struct _jobject;
typedef struct _jobject *jobject;
typedef jobject jarray;
typedef jarray jfloatArray;

struct JNINativeInterface_;
#ifdef __cplusplus
struct JNIEnv_;
typedef JNIEnv_ JNIEnv;
#else
typedef const struct JNINativeInterface_ *JNIEnv;
#endif

struct JNINativeInterface_ {
    void *slots_0_170[171];
#ifdef __cplusplus
    jsize (*GetArrayLength)(JNIEnv *env, jarray array);                                                /* 171 */
    void *slots_172_204[33];
    void (*GetFloatArrayRegion)(JNIEnv *env, jfloatArray array, jsize start, jsize len, jfloat *buf);   /* 205 */
    void *slots_206_212[7];
    void (*SetFloatArrayRegion)(JNIEnv *env, jfloatArray array, jsize start, jsize len, const jfloat *buf); /* 213 */
#else
    jsize (*GetArrayLength)(JNIEnv *env, jarray array);
    void *slots_172_204[33];
    void (*GetFloatArrayRegion)(JNIEnv *env, jfloatArray array, jsize start, jsize len, jfloat *buf);
    void *slots_206_212[7];
    void (*SetFloatArrayRegion)(JNIEnv *env, jfloatArray array, jsize start, jsize len, const jfloat *buf);
#endif
    void *slots_214_234[21];
};

#ifdef __cplusplus
struct JNIEnv_ {
    const struct JNINativeInterface_ *functions;
    jsize GetArrayLength(jarray a) { return functions->GetArrayLength(this, a); }
    void GetFloatArrayRegion(jfloatArray a, jsize s, jsize l, jfloat *b) { functions->GetFloatArrayRegion(this, a, s, l, b); }
    void SetFloatArrayRegion(jfloatArray a, jsize s, jsize l, const jfloat *b) { functions->SetFloatArrayRegion(this, a, s, l, b); }
};
#endif

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL

#ifdef __cplusplus
}
#endif
#endif
