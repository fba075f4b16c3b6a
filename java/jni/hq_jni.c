/* hq_jni.c — thin JNI shim between plugins.dbrasseur.hybridquantization.CudaImageManipulation and
 * the C ABI (include/hq_b200.h).  Compile-guarded: this image has no JDK (no jni.h), so the file
 * is NOT built or run here; tests/test_abi.py type-checks it against a stub of jni.h (tests/stubs/jni.h).  With a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       -o libhq_jni.so hq_jni.c -L../../hybridquantization_b200 -lhq_b200
 */
#if defined(__has_include)
#if __has_include(<jni.h>)
#define HQ_HAVE_JNI 1
#endif
#endif
#ifdef HQ_HAVE_JNI
#include <jni.h>
#include <stdint.h>
#include <stdio.h>
#include "hq_b200.h"

#define CLS(name) Java_plugins_dbrasseur_hybridquantization_CudaImageManipulation_##name

static void throw_hq(JNIEnv* env, hq_ctx* ctx, const char* what) {
    char msg[600];
    snprintf(msg, sizeof msg, "%s: %s", what, hq_last_error(ctx));
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), msg);
}

JNIEXPORT jlong JNICALL CLS(nCreate)(JNIEnv* env, jclass c, jint device) {
    hq_ctx* ctx = NULL;
    if (hq_create(device, &ctx) != HQ_OK) { throw_hq(env, NULL, "hq_create"); return 0; }
    return (jlong)(intptr_t)ctx;
}
JNIEXPORT void JNICALL CLS(nDestroy)(JNIEnv* env, jclass c, jlong h) { hq_destroy((hq_ctx*)(intptr_t)h); }
JNIEXPORT jlong JNICALL CLS(nPixels)(JNIEnv* env, jclass c, jlong h) { return (jlong)hq_image_pixels((hq_ctx*)(intptr_t)h); }

JNIEXPORT void JNICALL CLS(nSetImage)(JNIEnv* env, jclass c, jlong h, jbyteArray rgb, jint w, jint rows, jint wp) {
    hq_ctx* ctx = (hq_ctx*)(intptr_t)h;
    jbyte* p = (*env)->GetPrimitiveArrayCritical(env, rgb, NULL);
    const int rc = hq_set_image_u8(ctx, (const uint8_t*)p, w, rows, wp);
    (*env)->ReleasePrimitiveArrayCritical(env, rgb, p, JNI_ABORT);
    if (rc != HQ_OK) throw_hq(env, ctx, "hq_set_image_u8");
}

/* float[] planes exactly as im.getDataXYCAsFloat() holds them (HybridQuantization.java:98) */
JNIEXPORT void JNICALL CLS(nSetImageFloat)(JNIEnv* env, jclass c, jlong h, jfloatArray r, jfloatArray g, jfloatArray b, jint w, jint rows, jint wp) {
    hq_ctx* ctx = (hq_ctx*)(intptr_t)h;
    jfloat* pr = (*env)->GetPrimitiveArrayCritical(env, r, NULL);
    jfloat* pg = (*env)->GetPrimitiveArrayCritical(env, g, NULL);
    jfloat* pb = (*env)->GetPrimitiveArrayCritical(env, b, NULL);
    const int rc = hq_set_image_f32_planar(ctx, pr, pg, pb, w, rows, wp);
    (*env)->ReleasePrimitiveArrayCritical(env, b, pb, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, g, pg, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, r, pr, JNI_ABORT);
    if (rc != HQ_OK) throw_hq(env, ctx, "hq_set_image_f32_planar");
}

JNIEXPORT void JNICALL CLS(nEvalPalettes)(JNIEnv* env, jclass c, jlong h, jfloatArray pal, jint b, jint k, jint space,
                                          jlongArray errFx, jlongArray counts) {
    hq_ctx* ctx = (hq_ctx*)(intptr_t)h;
    jfloat* pp = (*env)->GetPrimitiveArrayCritical(env, pal, NULL);
    jlong* pe = (*env)->GetPrimitiveArrayCritical(env, errFx, NULL);
    jlong* pc = (*env)->GetPrimitiveArrayCritical(env, counts, NULL);
    /* a Java host that keeps its own annealing loop gets the exact pruned scoring where it pays (same integers) */
    const int flags = hq_search_eval_flags(ctx, k, space, HQ_COST_LAB);
    const int rc = hq_eval_palettes(ctx, pp, b, k, space, flags, (int64_t*)pe, (uint64_t*)pc, NULL);
    (*env)->ReleasePrimitiveArrayCritical(env, counts, pc, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, errFx, pe, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, pal, pp, JNI_ABORT);
    if (rc != HQ_OK) throw_hq(env, ctx, "hq_eval_palettes");
}

JNIEXPORT void JNICALL CLS(nQuantize)(JNIEnv* env, jclass c, jlong h, jfloatArray pal, jint k, jint space, jbyteArray out) {
    hq_ctx* ctx = (hq_ctx*)(intptr_t)h;
    jfloat* pp = (*env)->GetPrimitiveArrayCritical(env, pal, NULL);
    jbyte* po = (*env)->GetPrimitiveArrayCritical(env, out, NULL);
    const int rc = hq_quantize(ctx, pp, k, space, (uint8_t*)po, NULL, NULL);
    (*env)->ReleasePrimitiveArrayCritical(env, out, po, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, pal, pp, JNI_ABORT);
    if (rc != HQ_OK) throw_hq(env, ctx, "hq_quantize");
}
#else
/* no JDK: nothing to compile */
typedef int hq_jni_unavailable;
#endif
