/* Stand-in for <android/log.h> -- TEST INFRASTRUCTURE ONLY (nativedsp.cpp only
 * defines logging macros from it and never calls them). */
#ifndef RFA_ORACLE_STUB_ANDROID_LOG_H
#define RFA_ORACLE_STUB_ANDROID_LOG_H
enum { ANDROID_LOG_INFO = 4, ANDROID_LOG_ERROR = 6 };
static inline int __android_log_print(int, const char *, const char *, ...) { return 0; }
#endif
