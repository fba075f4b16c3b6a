/* hq_jni.c — JNI shim between plugins.dbrasseur.hybridquantization.CudaImageManipulation and the C ABI
 * (include/hq_b200.h).  One native per C entry the Java class needs; no logic of its own beyond marshaling.
 *
 * Array access: Get<Type>ArrayElements / Release<Type>ArrayElements (a copy or a pin, at the JVM's choice).  Critical
 * sections (GetPrimitiveArrayCritical) are NOT used: every entry below runs blocking CUDA work (allocation, copies, stream
 * waits), which the JNI specification forbids inside a critical region and which would stall the collector of the whole JVM.
 * Every array is checked for NULL and for the length the C call will touch BEFORE the call (IllegalArgumentException /
 * NullPointerException), so a short Java array can never become an out-of-bounds access of the JVM heap; a failed
 * Get...Elements (OutOfMemoryError pending) returns without calling into the library.
 *
 * This image has no JDK (no jni.h): the file is compiled against tests/stubs/jni.h and EXECUTED on the GPU through a fake
 * JNIEnv (tests/cpp/jni_harness.c, tests/test_gpu_cpp_api.py) — every native below runs there.  With a JDK:
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
#include <stdlib.h>
#include "hq_b200.h"

#define CLS(name) Java_plugins_dbrasseur_hybridquantization_CudaImageManipulation_##name
#define CTX(h) ((hq_ctx*)(intptr_t)(h))

static void throw_cls(JNIEnv* env, const char* cls, const char* msg) {
    jclass c = (*env)->FindClass(env, cls);
    if (c) (*env)->ThrowNew(env, c, msg);
}
static void throw_hq(JNIEnv* env, hq_ctx* ctx, const char* what) {
    char msg[600];
    snprintf(msg, sizeof msg, "%s: %s", what, hq_last_error(ctx));
    throw_cls(env, "java/lang/RuntimeException", msg);
}
/* 1 when `a` is non-NULL and holds at least `need` elements; otherwise throws and returns 0 */
static int need(JNIEnv* env, jarray a, jlong n, const char* name) {
    char msg[200];
    if (!a) { snprintf(msg, sizeof msg, "%s is null", name); throw_cls(env, "java/lang/NullPointerException", msg); return 0; }
    if (n < 0 || (jlong)(*env)->GetArrayLength(env, a) < n) {
        snprintf(msg, sizeof msg, "%s has %d elements, the call needs %lld", name, (int)(*env)->GetArrayLength(env, a), (long long)n);
        throw_cls(env, "java/lang/IllegalArgumentException", msg);
        return 0;
    }
    return 1;
}
static int bad_arg(JNIEnv* env, const char* msg) { throw_cls(env, "java/lang/IllegalArgumentException", msg); return 0; }

/* ---- lifetime: the constructor / close() (ImageManipulation.java:52-93, :265-269).  devices: one entry = hq_create,
 * several = ONE context over all of them (hq_create_multi: rows split and all-reduced inside the library) */
JNIEXPORT jlong JNICALL CLS(nCreate)(JNIEnv* env, jclass c, jintArray devices) {
    hq_ctx* ctx = NULL;
    if (!need(env, devices, 1, "devices")) return 0;
    const jint nd = (*env)->GetArrayLength(env, devices);
    jint* d = (*env)->GetIntArrayElements(env, devices, NULL);
    if (!d) return 0;
    int* dv = (int*)malloc(sizeof(int) * (size_t)nd);
    for (jint i = 0; dv && i < nd; ++i) dv[i] = (int)d[i];
    const int rc = dv ? hq_create_multi(dv, (int)nd, &ctx) : HQ_ERR_INVALID;
    free(dv);
    (*env)->ReleaseIntArrayElements(env, devices, d, JNI_ABORT);
    if (rc != HQ_OK) { throw_hq(env, NULL, "hq_create_multi"); return 0; }
    return (jlong)(intptr_t)ctx;
}
JNIEXPORT void JNICALL CLS(nDestroy)(JNIEnv* env, jclass c, jlong h) { hq_destroy(CTX(h)); }
JNIEXPORT jlong JNICALL CLS(nPixels)(JNIEnv* env, jclass c, jlong h) { return (jlong)hq_image_pixels(CTX(h)); }
JNIEXPORT jint JNICALL CLS(nDeviceCount)(JNIEnv* env, jclass c, jlong h) { return (jint)hq_multi_device_count(CTX(h)); }
JNIEXPORT void JNICALL CLS(nSetPruning)(JNIEnv* env, jclass c, jlong h, jint mode) {
    if (hq_set_pruning(CTX(h), mode) != HQ_OK) throw_hq(env, CTX(h), "hq_set_pruning");
}
JNIEXPORT void JNICALL CLS(nSetDeltaE)(JNIEnv* env, jclass c, jlong h, jint type) {   /* deltaETypes ordinal (:20): 0 CIE76, 1 CIE94, 2 CIEDE2000 (refused) */
    if (hq_set_delta_e(CTX(h), type) != HQ_OK) throw_hq(env, CTX(h), "hq_set_delta_e");
}
JNIEXPORT void JNICALL CLS(nRequestStop)(JNIEnv* env, jclass c, jlong h) { hq_request_stop(CTX(h)); }

/* ---- image */
JNIEXPORT void JNICALL CLS(nSetImage)(JNIEnv* env, jclass c, jlong h, jbyteArray rgb, jint w, jint rows, jint wp) {
    if (w < 0 || rows < 0) { bad_arg(env, "negative image size"); return; }
    if (!need(env, rgb, (jlong)w * rows * 3, "rgb")) return;
    jbyte* p = (*env)->GetByteArrayElements(env, rgb, NULL);
    if (!p) return;
    const int rc = hq_set_image_u8(CTX(h), (const uint8_t*)p, w, rows, wp);
    (*env)->ReleaseByteArrayElements(env, rgb, p, JNI_ABORT);
    if (rc != HQ_OK) throw_hq(env, CTX(h), "hq_set_image_u8");
}
/* float[] planes exactly as im.getDataXYCAsFloat() holds them (HybridQuantization.java:98) */
JNIEXPORT void JNICALL CLS(nSetImageFloat)(JNIEnv* env, jclass c, jlong h, jfloatArray r, jfloatArray g, jfloatArray b, jint w, jint rows, jint wp) {
    if (w < 0 || rows < 0) { bad_arg(env, "negative image size"); return; }
    const jlong n = (jlong)w * rows;
    if (!need(env, r, n, "R") || !need(env, g, n, "G") || !need(env, b, n, "B")) return;
    jfloat* pr = (*env)->GetFloatArrayElements(env, r, NULL);
    jfloat* pg = pr ? (*env)->GetFloatArrayElements(env, g, NULL) : NULL;
    jfloat* pb = pg ? (*env)->GetFloatArrayElements(env, b, NULL) : NULL;
    int rc = HQ_OK;
    if (pb) rc = hq_set_image_f32_planar(CTX(h), pr, pg, pb, w, rows, wp);
    if (pb) (*env)->ReleaseFloatArrayElements(env, b, pb, JNI_ABORT);
    if (pg) (*env)->ReleaseFloatArrayElements(env, g, pg, JNI_ABORT);
    if (pr) (*env)->ReleaseFloatArrayElements(env, r, pr, JNI_ABORT);
    if (pb && rc != HQ_OK) throw_hq(env, CTX(h), "hq_set_image_f32_planar");
}

/* ---- candidate evaluation.  scielab == 0: identity-filter Lab cost (hq_eval_palettes, exact pruning where the library's
 * policy says it pays); 1: the reference's full chain (hq_eval_palettes_scielab) */
JNIEXPORT void JNICALL CLS(nEvalPalettes)(JNIEnv* env, jclass c, jlong h, jfloatArray pal, jint b, jint k, jint space, jint scielab,
                                          jlongArray errFx, jlongArray counts) {
    if (b < 1 || k < 1) { bad_arg(env, "population and palette size must be >= 1"); return; }
    if (!need(env, pal, (jlong)b * k * 4, "palettes") || !need(env, errFx, b, "errFx") || !need(env, counts, (jlong)b * k, "counts")) return;
    jfloat* pp = (*env)->GetFloatArrayElements(env, pal, NULL);
    jlong* pe = pp ? (*env)->GetLongArrayElements(env, errFx, NULL) : NULL;
    jlong* pc = pe ? (*env)->GetLongArrayElements(env, counts, NULL) : NULL;
    int rc = HQ_OK;
    if (pc) {
        if (scielab) rc = hq_eval_palettes_scielab(CTX(h), pp, b, k, space, (int64_t*)pe, (uint64_t*)pc);
        else rc = hq_eval_palettes(CTX(h), pp, b, k, space, hq_search_eval_flags(CTX(h), k, space, HQ_COST_LAB), (int64_t*)pe, (uint64_t*)pc, NULL);
    }
    if (pc) (*env)->ReleaseLongArrayElements(env, counts, pc, 0);
    if (pe) (*env)->ReleaseLongArrayElements(env, errFx, pe, 0);
    if (pp) (*env)->ReleaseFloatArrayElements(env, pal, pp, JNI_ABORT);
    if (pc && rc != HQ_OK) throw_hq(env, CTX(h), scielab ? "hq_eval_palettes_scielab" : "hq_eval_palettes");
}

/* ---- S-CIELAB stage */
JNIEXPORT void JNICALL CLS(nScielabConfigure)(JNIEnv* env, jclass c, jlong h, jint dpi, jfloat distance) {
    if (hq_scielab_configure(CTX(h), dpi, distance) != HQ_OK) throw_hq(env, CTX(h), "hq_scielab_configure");
}
/* filters7: [7][taps] = O1g1,O1g2,O1g3,O2g1,O2g2,O3g1,O3g2 flattened; abs3: |O1g3| (updateOpenCLFilters, :800-841) */
JNIEXPORT void JNICALL CLS(nScielabSetFilters)(JNIEnv* env, jclass c, jlong h, jfloatArray filters7, jfloatArray abs3, jint taps) {
    if (taps < 1) { bad_arg(env, "taps must be >= 1"); return; }
    if (!need(env, filters7, (jlong)7 * taps, "filters") || !need(env, abs3, taps, "absfilters")) return;
    jfloat* pf = (*env)->GetFloatArrayElements(env, filters7, NULL);
    jfloat* pa = pf ? (*env)->GetFloatArrayElements(env, abs3, NULL) : NULL;
    int rc = HQ_OK;
    if (pa) rc = hq_scielab_set_filters(CTX(h), pf, pa, taps);
    if (pa) (*env)->ReleaseFloatArrayElements(env, abs3, pa, JNI_ABORT);
    if (pf) (*env)->ReleaseFloatArrayElements(env, filters7, pf, JNI_ABORT);
    if (pa && rc != HQ_OK) throw_hq(env, CTX(h), "hq_scielab_set_filters");
}
/* the caller's S-CIELAB image of the original, [n][4] (findBestQuantization's inlineScielabOriginal, :383) */
JNIEXPORT void JNICALL CLS(nScielabSetImage)(JNIEnv* env, jclass c, jlong h, jfloatArray lab4) {
    if (!need(env, lab4, (jlong)hq_image_pixels(CTX(h)) * 4, "inlineScielabOriginal")) return;
    jfloat* p = (*env)->GetFloatArrayElements(env, lab4, NULL);
    if (!p) return;
    const int rc = hq_scielab_set_image(CTX(h), p);
    (*env)->ReleaseFloatArrayElements(env, lab4, p, JNI_ABORT);
    if (rc != HQ_OK) throw_hq(env, CTX(h), "hq_scielab_set_image");
}
/* RGBtoXYZ (:100-152) */
JNIEXPORT void JNICALL CLS(nRgbToXyz)(JNIEnv* env, jclass c, jlong h, jfloatArray r, jfloatArray g, jfloatArray b, jfloatArray xyz4) {
    if (!need(env, r, 0, "R")) return;
    const jlong n = (*env)->GetArrayLength(env, r);
    if (!need(env, g, n, "G") || !need(env, b, n, "B") || !need(env, xyz4, 4 * n, "output")) return;
    jfloat* pr = (*env)->GetFloatArrayElements(env, r, NULL);
    jfloat* pg = pr ? (*env)->GetFloatArrayElements(env, g, NULL) : NULL;
    jfloat* pb = pg ? (*env)->GetFloatArrayElements(env, b, NULL) : NULL;
    jfloat* po = pb ? (*env)->GetFloatArrayElements(env, xyz4, NULL) : NULL;
    int rc = HQ_OK;
    if (po) rc = hq_rgb_to_xyz(CTX(h), pr, pg, pb, (size_t)n, po);
    if (po) (*env)->ReleaseFloatArrayElements(env, xyz4, po, 0);
    if (pb) (*env)->ReleaseFloatArrayElements(env, b, pb, JNI_ABORT);
    if (pg) (*env)->ReleaseFloatArrayElements(env, g, pg, JNI_ABORT);
    if (pr) (*env)->ReleaseFloatArrayElements(env, r, pr, JNI_ABORT);
    if (po && rc != HQ_OK) throw_hq(env, CTX(h), "hq_rgb_to_xyz");
}
/* XYZtoScielab (:285-370) with the context's filter bank */
JNIEXPORT void JNICALL CLS(nXyzToScielab)(JNIEnv* env, jclass c, jlong h, jfloatArray xyz4, jint w, jfloatArray illuminant, jfloatArray lab4) {
    if (w < 1) { bad_arg(env, "width must be >= 1"); return; }
    if (!need(env, xyz4, 0, "XYZ") || !need(env, illuminant, 3, "illuminant")) return;
    const jlong n4 = (*env)->GetArrayLength(env, xyz4);
    if (n4 % (4 * (jlong)w) != 0) { bad_arg(env, "XYZ.length is not a multiple of 4 * width"); return; }
    if (!need(env, lab4, n4, "output")) return;
    jfloat* px = (*env)->GetFloatArrayElements(env, xyz4, NULL);
    jfloat* pi = px ? (*env)->GetFloatArrayElements(env, illuminant, NULL) : NULL;
    jfloat* po = pi ? (*env)->GetFloatArrayElements(env, lab4, NULL) : NULL;
    int rc = HQ_OK;
    if (po) rc = hq_xyz_to_scielab(CTX(h), px, w, (int)(n4 / 4 / w), pi, po);
    if (po) (*env)->ReleaseFloatArrayElements(env, lab4, po, 0);
    if (pi) (*env)->ReleaseFloatArrayElements(env, illuminant, pi, JNI_ABORT);
    if (px) (*env)->ReleaseFloatArrayElements(env, xyz4, px, JNI_ABORT);
    if (po && rc != HQ_OK) throw_hq(env, CTX(h), "hq_xyz_to_scielab");
}

/* ---- output image: quantize (:770-798).  Either output may be null: outRgb packed u8 [n][3], outF32 [n][4] (what the
 * reference returns) */
JNIEXPORT void JNICALL CLS(nQuantize)(JNIEnv* env, jclass c, jlong h, jfloatArray pal, jint k, jint space, jbyteArray outRgb, jfloatArray outF32) {
    const jlong n = (jlong)hq_image_pixels(CTX(h));
    if (k < 1) { bad_arg(env, "palette size must be >= 1"); return; }
    if (!need(env, pal, (jlong)k * 4, "colors")) return;
    if (outRgb && !need(env, outRgb, 3 * n, "outRgb")) return;
    if (outF32 && !need(env, outF32, 4 * n, "outF32")) return;
    jfloat* pp = (*env)->GetFloatArrayElements(env, pal, NULL);
    if (!pp) return;
    jbyte* po = outRgb ? (*env)->GetByteArrayElements(env, outRgb, NULL) : NULL;
    jfloat* pf = outF32 && (po || !outRgb) ? (*env)->GetFloatArrayElements(env, outF32, NULL) : NULL;
    const int ok = (!outRgb || po) && (!outF32 || pf);
    int rc = HQ_OK;
    if (ok) rc = hq_quantize(CTX(h), pp, k, space, (uint8_t*)po, pf, NULL);
    if (pf) (*env)->ReleaseFloatArrayElements(env, outF32, pf, 0);
    if (po) (*env)->ReleaseByteArrayElements(env, outRgb, po, 0);
    (*env)->ReleaseFloatArrayElements(env, pal, pp, JNI_ABORT);
    if (ok && rc != HQ_OK) throw_hq(env, CTX(h), "hq_quantize");
}

/* ---- error-image mode */
/* computeError as the reference declares it (:858-894): two Lab images, error image in lanes 0..2; returns the mean dE */
JNIEXPORT jdouble JNICALL CLS(nDeltaEImages)(JNIEnv* env, jclass c, jlong h, jfloatArray a, jfloatArray b, jfloatArray errorImage) {
    if (!need(env, a, 0, "original")) return 0;
    const jlong n4 = (*env)->GetArrayLength(env, a);
    if (!need(env, b, n4, "quantized")) return 0;
    if (errorImage && !need(env, errorImage, n4, "errorImage")) return 0;
    jfloat* pa = (*env)->GetFloatArrayElements(env, a, NULL);
    jfloat* pb = pa ? (*env)->GetFloatArrayElements(env, b, NULL) : NULL;
    jfloat* pe = pb && errorImage ? (*env)->GetFloatArrayElements(env, errorImage, NULL) : NULL;
    const int ok = pb && (!errorImage || pe);
    double mean = 0;
    int rc = HQ_OK;
    if (ok) rc = hq_delta_e_images(CTX(h), pa, pb, (size_t)(n4 / 4), pe, &mean);
    if (pe) (*env)->ReleaseFloatArrayElements(env, errorImage, pe, 0);
    if (pb) (*env)->ReleaseFloatArrayElements(env, b, pb, JNI_ABORT);
    if (pa) (*env)->ReleaseFloatArrayElements(env, a, pa, JNI_ABORT);
    if (ok && rc != HQ_OK) throw_hq(env, CTX(h), "hq_delta_e_images");
    return mean;
}
/* HybridQuantization.errorImage (:139-182) in one call: second image as packed u8 (rgb8) or float planes; map [n] may be null */
JNIEXPORT jdouble JNICALL CLS(nErrorImage)(JNIEnv* env, jclass c, jlong h, jbyteArray rgb8, jfloatArray map) {
    const jlong n = (jlong)hq_image_pixels(CTX(h));
    if (!need(env, rgb8, 3 * n, "quantized")) return 0;
    if (map && !need(env, map, n, "errorMap")) return 0;
    jbyte* pq = (*env)->GetByteArrayElements(env, rgb8, NULL);
    jfloat* pm = pq && map ? (*env)->GetFloatArrayElements(env, map, NULL) : NULL;
    const int ok = pq && (!map || pm);
    double mean = 0;
    int rc = HQ_OK;
    if (ok) rc = hq_error_image(CTX(h), (const uint8_t*)pq, pm, NULL, &mean);
    if (pm) (*env)->ReleaseFloatArrayElements(env, map, pm, 0);
    if (pq) (*env)->ReleaseByteArrayElements(env, rgb8, pq, JNI_ABORT);
    if (ok && rc != HQ_OK) throw_hq(env, CTX(h), "hq_error_image");
    return mean;
}
JNIEXPORT jdouble JNICALL CLS(nErrorImageFloat)(JNIEnv* env, jclass c, jlong h, jfloatArray r, jfloatArray g, jfloatArray b, jfloatArray map) {
    const jlong n = (jlong)hq_image_pixels(CTX(h));
    if (!need(env, r, n, "R") || !need(env, g, n, "G") || !need(env, b, n, "B")) return 0;
    if (map && !need(env, map, n, "errorMap")) return 0;
    jfloat* pr = (*env)->GetFloatArrayElements(env, r, NULL);
    jfloat* pg = pr ? (*env)->GetFloatArrayElements(env, g, NULL) : NULL;
    jfloat* pb = pg ? (*env)->GetFloatArrayElements(env, b, NULL) : NULL;
    jfloat* pm = pb && map ? (*env)->GetFloatArrayElements(env, map, NULL) : NULL;
    const int ok = pb && (!map || pm);
    double mean = 0;
    int rc = HQ_OK;
    if (ok) rc = hq_error_image_f32_planar(CTX(h), pr, pg, pb, pm, NULL, &mean);
    if (pm) (*env)->ReleaseFloatArrayElements(env, map, pm, 0);
    if (pb) (*env)->ReleaseFloatArrayElements(env, b, pb, JNI_ABORT);
    if (pg) (*env)->ReleaseFloatArrayElements(env, g, pg, JNI_ABORT);
    if (pr) (*env)->ReleaseFloatArrayElements(env, r, pr, JNI_ABORT);
    if (ok && rc != HQ_OK) throw_hq(env, CTX(h), "hq_error_image_f32_planar");
    return mean;
}

/* ---- the whole annealing search inside the library (hq_find_best_quantization: same SWASA schedule, java.util.Random
 * with an explicit seed).  ip = {population, imax, iTc, convergence, space, costModel}; fp = {delta, convDelay, convSpread,
 * T0, alpha, s0, beta}; listener (may be null) receives progress(int iteration, int maxIterations, double bestError) every
 * 10 iterations on THIS thread (where the reference updates the progress bar, :546-551) and may call requestStop(). */
typedef struct { JNIEnv* env; jobject listener; jmethodID mid; } hq_jni_progress;
static void jni_progress(void* user, int it, int max_it, double best) {
    hq_jni_progress* p = (hq_jni_progress*)user;
    if ((*p->env)->ExceptionCheck(p->env)) return;   /* a listener threw: make no further calls into the JVM */
    (*p->env)->CallVoidMethod(p->env, p->listener, p->mid, (jint)it, (jint)max_it, (jdouble)best);
}
JNIEXPORT jint JNICALL CLS(nFindBestQuantization)(JNIEnv* env, jclass c, jlong h, jint k, jintArray ip, jfloatArray fp, jlong seed,
                                                  jfloatArray best, jdoubleArray bestError, jdoubleArray trace, jobject listener) {
    if (k < 1) { bad_arg(env, "palette size must be >= 1"); return 0; }
    if (!need(env, ip, 6, "intParams") || !need(env, fp, 7, "floatParams") || !need(env, best, (jlong)4 * k, "bestColors") || !need(env, bestError, 1, "bestError")) return 0;
    jint iv[6];
    jfloat fv[7];
    {
        jint* pi = (*env)->GetIntArrayElements(env, ip, NULL);
        if (!pi) return 0;
        for (int i = 0; i < 6; ++i) iv[i] = pi[i];
        (*env)->ReleaseIntArrayElements(env, ip, pi, JNI_ABORT);
        jfloat* pf = (*env)->GetFloatArrayElements(env, fp, NULL);
        if (!pf) return 0;
        for (int i = 0; i < 7; ++i) fv[i] = pf[i];
        (*env)->ReleaseFloatArrayElements(env, fp, pf, JNI_ABORT);
    }
    if (iv[0] < 1 || iv[1] < 1) { bad_arg(env, "population and imax must be >= 1"); return 0; }
    if (trace && !need(env, trace, ((jlong)iv[1] + 1) * iv[0], "trace")) return 0;
    hq_swasa_params p;
    hq_swasa_default_params(&p);
    p.population = iv[0]; p.imax = iv[1]; p.iTc = iv[2]; p.convergence = iv[3]; p.space = iv[4]; p.cost_model = iv[5];
    p.delta = fv[0]; p.conv_delay = fv[1]; p.conv_spread = fv[2]; p.t0 = fv[3]; p.alpha = fv[4]; p.s0 = fv[5]; p.beta = fv[6];
    p.seed = (int64_t)seed;
    hq_jni_progress prog = {env, listener, NULL};
    if (listener) {
        jclass lc = (*env)->GetObjectClass(env, listener);
        prog.mid = lc ? (*env)->GetMethodID(env, lc, "progress", "(IID)V") : NULL;
        if (!prog.mid) return 0;   /* NoSuchMethodError pending */
        hq_set_progress(CTX(h), jni_progress, &prog);
    }
    jfloat* pb = (*env)->GetFloatArrayElements(env, best, NULL);
    jdouble* pe = pb ? (*env)->GetDoubleArrayElements(env, bestError, NULL) : NULL;
    jdouble* pt = pe && trace ? (*env)->GetDoubleArrayElements(env, trace, NULL) : NULL;
    const int ok = pe && (!trace || pt);
    int its = 0, rc = HQ_OK;
    if (ok) rc = hq_find_best_quantization(CTX(h), k, &p, 0, pb, pe, pt, &its);
    if (listener) hq_set_progress(CTX(h), NULL, NULL);
    if (pt) (*env)->ReleaseDoubleArrayElements(env, trace, pt, 0);
    if (pe) (*env)->ReleaseDoubleArrayElements(env, bestError, pe, 0);
    if (pb) (*env)->ReleaseFloatArrayElements(env, best, pb, 0);
    if (ok && rc != HQ_OK && !(*env)->ExceptionCheck(env)) throw_hq(env, CTX(h), "hq_find_best_quantization");
    return (jint)its;
}
#else
/* no JDK: nothing to compile */
typedef int hq_jni_unavailable;
#endif
