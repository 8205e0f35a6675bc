// jni_shim.cpp -> libnativedsp.so: the reference's two JNI entry points on top of librfa_b200.
//
// Drop-in for nativedsp/src/main/cpp/nativedsp.cpp: same mangled symbols, same array contracts
//   Java_com_mantz_1it_nativedsp_NativeDsp_performFFT           (nativedsp.cpp:19-42)
//   Java_com_mantz_1it_nativedsp_NativeDsp_performFFTAndLogMag  (nativedsp.cpp:44-81)
// so `System.loadLibrary("nativedsp")` (NativeDsp.kt:30-35) binds them without touching a line
// of Kotlin.  Like the reference it keeps process-global state and is not thread safe
// (NativeDsp.kt:23-26); unlike the reference it does not leak its buffers on a size change.
// The transform size is GetArrayLength(input)/2; sizes other than a power of two in 16..65536
// are reported on stderr and leave the output untouched (the reference would crash on a NULL
// pffft setup).
#if __has_include(<jni.h>)
#include <jni.h>
#else
#include "jni_min/jni.h"
#endif

#include <cstdio>

#include "../../include/rfa_b200.h"

namespace {
rfa_ctx *g_ctx = nullptr;
float *g_in = nullptr, *g_out = nullptr;  // pinned staging, grow-only
jsize g_cap = 0;

bool ready(jsize length) {
    if (!g_ctx && rfa_ctx_create(0, nullptr, &g_ctx) != RFA_OK) {
        fprintf(stderr, "nativedsp(b200): %s\n", rfa_last_error());
        return false;
    }
    if (length > g_cap) {
        rfa_host_free(g_in);
        rfa_host_free(g_out);
        g_in = g_out = nullptr;
        g_cap = 0;
        if (rfa_host_alloc(sizeof(float) * (size_t)length, (void **)&g_in) != RFA_OK ||
            rfa_host_alloc(sizeof(float) * (size_t)length, (void **)&g_out) != RFA_OK) {
            fprintf(stderr, "nativedsp(b200): %s\n", rfa_last_error());
            return false;
        }
        g_cap = length;
    }
    return true;
}
}  // namespace

extern "C" JNIEXPORT void JNICALL Java_com_mantz_1it_nativedsp_NativeDsp_performFFT(JNIEnv *env, jobject,
                                                                                   jfloatArray inputArray,
                                                                                   jfloatArray outputArray) {
    const jsize length = env->GetArrayLength(inputArray);
    if (!ready(length)) return;
    env->GetFloatArrayRegion(inputArray, 0, length, g_in);
    if (rfa_fft_c2c(g_ctx, g_in, g_out, length / 2, 1, RFA_MEM_HOST) != RFA_OK) {
        fprintf(stderr, "nativedsp(b200): performFFT: %s\n", rfa_last_error());
        return;
    }
    env->SetFloatArrayRegion(outputArray, 0, length, g_out);
}

extern "C" JNIEXPORT void JNICALL Java_com_mantz_1it_nativedsp_NativeDsp_performFFTAndLogMag(JNIEnv *env, jobject,
                                                                                            jfloatArray inputArray,
                                                                                            jfloatArray outputArray) {
    const jsize length = env->GetArrayLength(inputArray);
    const jsize outputLength = length / 2;
    if (!ready(length)) return;
    env->GetFloatArrayRegion(inputArray, 0, length, g_in);
    if (rfa_fft_logmag(g_ctx, g_in, g_out, outputLength, 1, RFA_MEM_HOST) != RFA_OK) {
        fprintf(stderr, "nativedsp(b200): performFFTAndLogMag: %s\n", rfa_last_error());
        return;
    }
    env->SetFloatArrayRegion(outputArray, 0, outputLength, g_out);
}
