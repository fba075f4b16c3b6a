/* tests/stubs/jni.h — NOT the JDK's header.  A minimal stand-in declaring only the JNI types and the JNIEnv entries that
 * java/jni/hq_jni.c uses, so that the shim can be type-checked (tests/test_abi.py, gcc -fsyntax-only) and EXECUTED through a
 * fake JNIEnv (tests/cpp/jni_harness.c) in an image without a JDK.  Signatures follow the JNI specification (jni.h of any
 * JDK); the member ORDER of the function table is irrelevant here because the shim only uses the members by name. */
#ifndef HQ_TEST_STUB_JNI_H
#define HQ_TEST_STUB_JNI_H
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef float jfloat;
typedef double jdouble;
typedef uint8_t jboolean;
typedef jint jsize;
struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jbyteArray;
typedef jarray jintArray;
typedef jarray jfloatArray;
typedef jarray jlongArray;
typedef jarray jdoubleArray;
typedef jobject jthrowable;
struct _jmethodID;
typedef struct _jmethodID* jmethodID;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv* env, const char* name);
    jint (*ThrowNew)(JNIEnv* env, jclass clazz, const char* msg);
    jboolean (*ExceptionCheck)(JNIEnv* env);
    jsize (*GetArrayLength)(JNIEnv* env, jarray array);
    jclass (*GetObjectClass)(JNIEnv* env, jobject obj);
    jmethodID (*GetMethodID)(JNIEnv* env, jclass clazz, const char* name, const char* sig);
    void (*CallVoidMethod)(JNIEnv* env, jobject obj, jmethodID methodID, ...);
    jbyte* (*GetByteArrayElements)(JNIEnv* env, jbyteArray array, jboolean* isCopy);
    jint* (*GetIntArrayElements)(JNIEnv* env, jintArray array, jboolean* isCopy);
    jlong* (*GetLongArrayElements)(JNIEnv* env, jlongArray array, jboolean* isCopy);
    jfloat* (*GetFloatArrayElements)(JNIEnv* env, jfloatArray array, jboolean* isCopy);
    jdouble* (*GetDoubleArrayElements)(JNIEnv* env, jdoubleArray array, jboolean* isCopy);
    void (*ReleaseByteArrayElements)(JNIEnv* env, jbyteArray array, jbyte* elems, jint mode);
    void (*ReleaseIntArrayElements)(JNIEnv* env, jintArray array, jint* elems, jint mode);
    void (*ReleaseLongArrayElements)(JNIEnv* env, jlongArray array, jlong* elems, jint mode);
    void (*ReleaseFloatArrayElements)(JNIEnv* env, jfloatArray array, jfloat* elems, jint mode);
    void (*ReleaseDoubleArrayElements)(JNIEnv* env, jdoubleArray array, jdouble* elems, jint mode);
};
#endif
