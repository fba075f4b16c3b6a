/* tests/stubs/jni.h — NOT the JDK's header.  A minimal stand-in declaring only the JNI types and the JNIEnv entries that
 * java/jni/hq_jni.c uses, so that tests/test_abi.py can at least type-check the shim against include/hq_b200.h in an image
 * without a JDK (gcc -fsyntax-only).  Signatures follow the JNI specification (jni.h of any JDK). */
#ifndef HQ_TEST_STUB_JNI_H
#define HQ_TEST_STUB_JNI_H
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef float jfloat;
typedef uint8_t jboolean;
struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jbyteArray;
typedef jarray jfloatArray;
typedef jarray jlongArray;
typedef jobject jthrowable;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv* env, const char* name);
    jint (*ThrowNew)(JNIEnv* env, jclass clazz, const char* msg);
    void* (*GetPrimitiveArrayCritical)(JNIEnv* env, jarray array, jboolean* isCopy);
    void (*ReleasePrimitiveArrayCritical)(JNIEnv* env, jarray array, void* carray, jint mode);
};
#endif
