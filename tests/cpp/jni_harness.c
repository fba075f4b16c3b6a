/* Drives java/jni/hq_jni.c WITHOUT a JVM: the shim is compiled against tests/stubs/jni.h and called through a fake JNIEnv
 * whose arrays are plain C buffers with a length (a jarray is a pointer to a {data, len} record), whose exceptions are
 * recorded, and whose CallVoidMethod delivers the progress callback to a C "listener".  This executes EVERY native method
 * of plugins.dbrasseur.hybridquantization.CudaImageManipulation on the GPU exactly as the JVM would invoke them — array
 * length checks, element access, the C-ABI calls, the error paths — and prints one JSON object for
 * tests/test_gpu_cpp_api.py to compare with the oracle. */
#include <jni.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct _jobject { void* data; int len; };
struct _jmethodID { int dummy; };
static int g_thrown = 0;
static char g_msg[700], g_cls[100];
static jclass fake_FindClass(JNIEnv* env, const char* name) {
    (void)env;
    static struct _jobject cls[8];
    static char names[8][100];
    static int n = 0;
    const int i = n++ & 7;
    snprintf(names[i], sizeof names[i], "%s", name);
    cls[i].data = names[i]; cls[i].len = 0;
    return &cls[i];
}
static jint fake_ThrowNew(JNIEnv* env, jclass c, const char* msg) {
    (void)env;
    g_thrown = 1;
    snprintf(g_cls, sizeof g_cls, "%s", c && c->data ? (const char*)c->data : "?");
    snprintf(g_msg, sizeof g_msg, "%s", msg);
    return 0;
}
static jboolean fake_ExceptionCheck(JNIEnv* env) { (void)env; return (jboolean)g_thrown; }
static jsize fake_GetArrayLength(JNIEnv* env, jarray a) { (void)env; return a->len; }
static jclass fake_GetObjectClass(JNIEnv* env, jobject o) { (void)env; return o; }
static struct _jmethodID g_mid;
static jmethodID fake_GetMethodID(JNIEnv* env, jclass c, const char* name, const char* sig) {
    (void)env; (void)c;
    return (strcmp(name, "progress") == 0 && strcmp(sig, "(IID)V") == 0) ? &g_mid : NULL;
}
/* the "Java" progress listener: records the calls; optionally asks for a stop at its first call */
static int g_prog_calls = 0, g_prog_first = 0, g_prog_last = 0, g_prog_max = 0, g_stop_at_first = 0;
static double g_prog_best = 0;
static jlong g_ctx_for_stop = 0;
#define CLS(name) Java_plugins_dbrasseur_hybridquantization_CudaImageManipulation_##name
void CLS(nRequestStop)(JNIEnv*, jclass, jlong);
static void fake_CallVoidMethod(JNIEnv* env, jobject obj, jmethodID mid, ...) {
    (void)obj; (void)mid;
    va_list ap;
    va_start(ap, mid);
    const int it = va_arg(ap, int), mx = va_arg(ap, int);
    const double best = va_arg(ap, double);
    va_end(ap);
    if (g_prog_calls++ == 0) g_prog_first = it;
    g_prog_last = it; g_prog_max = mx; g_prog_best = best;
    if (g_stop_at_first) CLS(nRequestStop)(env, NULL, g_ctx_for_stop);
}
static void* fake_get(jarray a) { return a->data; }
static jbyte* fake_GetB(JNIEnv* e, jbyteArray a, jboolean* c) { (void)e; if (c) *c = 0; return fake_get(a); }
static jint* fake_GetI(JNIEnv* e, jintArray a, jboolean* c) { (void)e; if (c) *c = 0; return fake_get(a); }
static jlong* fake_GetL(JNIEnv* e, jlongArray a, jboolean* c) { (void)e; if (c) *c = 0; return fake_get(a); }
static jfloat* fake_GetF(JNIEnv* e, jfloatArray a, jboolean* c) { (void)e; if (c) *c = 0; return fake_get(a); }
static jdouble* fake_GetD(JNIEnv* e, jdoubleArray a, jboolean* c) { (void)e; if (c) *c = 0; return fake_get(a); }
static void fake_RelB(JNIEnv* e, jbyteArray a, jbyte* p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static void fake_RelI(JNIEnv* e, jintArray a, jint* p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static void fake_RelL(JNIEnv* e, jlongArray a, jlong* p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static void fake_RelF(JNIEnv* e, jfloatArray a, jfloat* p, jint m) { (void)e; (void)a; (void)p; (void)m; }
static void fake_RelD(JNIEnv* e, jdoubleArray a, jdouble* p, jint m) { (void)e; (void)a; (void)p; (void)m; }

jlong CLS(nCreate)(JNIEnv*, jclass, jintArray);
void CLS(nDestroy)(JNIEnv*, jclass, jlong);
jlong CLS(nPixels)(JNIEnv*, jclass, jlong);
jint CLS(nDeviceCount)(JNIEnv*, jclass, jlong);
void CLS(nSetPruning)(JNIEnv*, jclass, jlong, jint);
void CLS(nSetDeltaE)(JNIEnv*, jclass, jlong, jint);
void CLS(nSetImage)(JNIEnv*, jclass, jlong, jbyteArray, jint, jint, jint);
void CLS(nSetImageFloat)(JNIEnv*, jclass, jlong, jfloatArray, jfloatArray, jfloatArray, jint, jint, jint);
void CLS(nEvalPalettes)(JNIEnv*, jclass, jlong, jfloatArray, jint, jint, jint, jint, jlongArray, jlongArray);
void CLS(nScielabConfigure)(JNIEnv*, jclass, jlong, jint, jfloat);
void CLS(nScielabSetFilters)(JNIEnv*, jclass, jlong, jfloatArray, jfloatArray, jint);
void CLS(nScielabSetImage)(JNIEnv*, jclass, jlong, jfloatArray);
void CLS(nRgbToXyz)(JNIEnv*, jclass, jlong, jfloatArray, jfloatArray, jfloatArray, jfloatArray);
void CLS(nXyzToScielab)(JNIEnv*, jclass, jlong, jfloatArray, jint, jfloatArray, jfloatArray);
void CLS(nQuantize)(JNIEnv*, jclass, jlong, jfloatArray, jint, jint, jbyteArray, jfloatArray);
jdouble CLS(nDeltaEImages)(JNIEnv*, jclass, jlong, jfloatArray, jfloatArray, jfloatArray);
jdouble CLS(nErrorImage)(JNIEnv*, jclass, jlong, jbyteArray, jfloatArray);
jdouble CLS(nErrorImageFloat)(JNIEnv*, jclass, jlong, jfloatArray, jfloatArray, jfloatArray, jfloatArray);
jint CLS(nFindBestQuantization)(JNIEnv*, jclass, jlong, jint, jintArray, jfloatArray, jlong, jfloatArray, jdoubleArray, jdoubleArray, jobject);
int hq_scielab_build_filters(int dpi, float viewing_distance_cm, float* filters7, float* abs3, int* taps);

static unsigned long long splitmix64(unsigned long long* s) {
    unsigned long long z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static unsigned long long fnv(const void* p, size_t bytes) {
    unsigned long long h = 0;
    const unsigned char* b = p;
    for (size_t j = 0; j < bytes; ++j) h = h * 1099511628211ULL + b[j];
    return h;
}
#define MUST(what) do { if (g_thrown) { fprintf(stderr, "%s threw %s: %s\n", what, g_cls, g_msg); return 2; } } while (0)
static void print_ll(const char* key, const jlong* v, int n) {
    printf("\"%s\": [", key);
    for (int i = 0; i < n; ++i) printf("%s%lld", i ? ", " : "", (long long)v[i]);
    printf("], ");
}

int main(int argc, char** argv) {
    const int w = argc > 1 ? atoi(argv[1]) : 320, h = argc > 2 ? atoi(argv[2]) : 240, K = argc > 3 ? atoi(argv[3]) : 48, B = 3;
    const struct JNINativeInterface_ table = {fake_FindClass, fake_ThrowNew, fake_ExceptionCheck, fake_GetArrayLength, fake_GetObjectClass,
                                              fake_GetMethodID, fake_CallVoidMethod, fake_GetB, fake_GetI, fake_GetL, fake_GetF, fake_GetD,
                                              fake_RelB, fake_RelI, fake_RelL, fake_RelF, fake_RelD};
    JNIEnv envp = &table;
    JNIEnv* env = &envp;
    /* same bytes as hybridquantization_b200.synth.synth_image(w, h, 5) */
    const size_t n = (size_t)w * h;
    unsigned char* rgb = malloc(n * 3);
    unsigned long long s = 5, cur = 0;
    splitmix64(&s); splitmix64(&s);
    for (size_t j = 0; j < n * 3; ++j) { if ((j & 7) == 0) cur = splitmix64(&s); rgb[j] = (unsigned char)(cur >> (8 * (j & 7))); }
    /* palettes: a simple deterministic lattice the Python side rebuilds */
    float* pal = calloc((size_t)B * K * 4, sizeof(float));
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < 3; ++c) pal[((size_t)b * K + k) * 4 + c] = (float)((b * 7919 + k * 104729 + c * 1299709) % 1000) / 999.0f;
    jlong *err = calloc(B, sizeof(jlong)), *errf = calloc(B, sizeof(jlong)), *errs = calloc(B, sizeof(jlong)), *errs2 = calloc(B, sizeof(jlong)), *errs3 = calloc(B, sizeof(jlong));
    jlong *cnt = calloc((size_t)B * K, sizeof(jlong)), *cnts = calloc((size_t)B * K, sizeof(jlong)), *cnt2 = calloc((size_t)B * K, sizeof(jlong));
    unsigned char* out = malloc(n * 3);
    float* outf = malloc(n * 16);
    float* planes = malloc(n * 12);
    for (size_t j = 0; j < n; ++j)
        for (int c = 0; c < 3; ++c) planes[(size_t)c * n + j] = (float)(rgb[3 * j + c] / 255.0);   /* the plugin's float planes (c/255) */
    jint dev0[1] = {0};
    struct _jobject a_dev = {dev0, 1}, a_rgb = {rgb, (int)(n * 3)}, a_pal = {pal, B * K * 4}, a_err = {err, B}, a_errf = {errf, B}, a_errs = {errs, B},
                    a_errs2 = {errs2, B}, a_errs3 = {errs3, B}, a_cnt = {cnt, B * K}, a_cnts = {cnts, B * K}, a_cnt2 = {cnt2, B * K}, a_out = {out, (int)(n * 3)},
                    a_outf = {outf, (int)(n * 4)}, a_r = {planes, (int)n}, a_g = {planes + n, (int)n}, a_b = {planes + 2 * n, (int)n};

    const jlong ctx = CLS(nCreate)(env, NULL, &a_dev); MUST("nCreate");
    g_ctx_for_stop = ctx;
    const int ndev = CLS(nDeviceCount)(env, NULL, ctx);
    CLS(nSetImage)(env, NULL, ctx, &a_rgb, w, h, 0); MUST("nSetImage");
    const jlong px = CLS(nPixels)(env, NULL, ctx);
    CLS(nSetPruning)(env, NULL, ctx, 1); MUST("nSetPruning");
    CLS(nEvalPalettes)(env, NULL, ctx, &a_pal, B, K, 0, 0, &a_err, &a_cnt); MUST("nEvalPalettes");
    CLS(nQuantize)(env, NULL, ctx, &a_pal, K, 0, &a_out, &a_outf); MUST("nQuantize");
    const unsigned long long img_hash = fnv(out, n * 3), f32_hash = fnv(outf, n * 16);

    /* ---- the reference-faithful chain: filters from dpi / distance, then the same bank installed explicitly */
    CLS(nScielabConfigure)(env, NULL, ctx, 72, 45.0f); MUST("nScielabConfigure");
    CLS(nEvalPalettes)(env, NULL, ctx, &a_pal, B, K, 1, 1, &a_errs, &a_cnts); MUST("nEvalPalettes (S-CIELAB)");
    int taps = 0;
    hq_scielab_build_filters(72, 45.0f, NULL, NULL, &taps);
    float *f7 = malloc(sizeof(float) * 7 * taps), *ab = malloc(sizeof(float) * taps);
    int cap = taps;
    hq_scielab_build_filters(72, 45.0f, f7, ab, &cap);
    struct _jobject a_f7 = {f7, 7 * taps}, a_ab = {ab, taps};
    CLS(nScielabSetFilters)(env, NULL, ctx, &a_f7, &a_ab, taps); MUST("nScielabSetFilters");
    CLS(nEvalPalettes)(env, NULL, ctx, &a_pal, B, K, 1, 1, &a_errs2, &a_cnt2); MUST("nEvalPalettes (S-CIELAB, installed bank)");

    /* ---- the reference's own route to S-CIELAB(original): RGBtoXYZ -> XYZtoScielab, installed as the comparison target */
    float *xyz = malloc(n * 16), *lab = malloc(n * 16), *xyzq = malloc(n * 16), *labq = malloc(n * 16), *eimg = calloc(n * 4, sizeof(float));
    float ill[3] = {0.95047f, 1.0f, 1.0883f};
    struct _jobject a_xyz = {xyz, (int)(n * 4)}, a_lab = {lab, (int)(n * 4)}, a_xyzq = {xyzq, (int)(n * 4)}, a_labq = {labq, (int)(n * 4)},
                    a_eimg = {eimg, (int)(n * 4)}, a_ill = {ill, 3};
    CLS(nRgbToXyz)(env, NULL, ctx, &a_r, &a_g, &a_b, &a_xyz); MUST("nRgbToXyz");
    CLS(nXyzToScielab)(env, NULL, ctx, &a_xyz, w, &a_ill, &a_lab); MUST("nXyzToScielab");
    CLS(nScielabSetImage)(env, NULL, ctx, &a_lab); MUST("nScielabSetImage");
    CLS(nEvalPalettes)(env, NULL, ctx, &a_pal, B, K, 1, 1, &a_errs3, &a_cnt2); MUST("nEvalPalettes (S-CIELAB, installed original)");

    /* ---- error image three ways: computeError on two Lab images, and the one-call forms (u8 and float second image) */
    float* qplanes = malloc(n * 12);
    for (size_t j = 0; j < n; ++j)
        for (int c = 0; c < 3; ++c) qplanes[(size_t)c * n + j] = (float)(out[3 * j + c] / 255.0);
    struct _jobject a_qr = {qplanes, (int)n}, a_qg = {qplanes + n, (int)n}, a_qb = {qplanes + 2 * n, (int)n};
    CLS(nRgbToXyz)(env, NULL, ctx, &a_qr, &a_qg, &a_qb, &a_xyzq); MUST("nRgbToXyz (quantised)");
    CLS(nXyzToScielab)(env, NULL, ctx, &a_xyzq, w, &a_ill, &a_labq); MUST("nXyzToScielab (quantised)");
    const double de_lab = CLS(nDeltaEImages)(env, NULL, ctx, &a_lab, &a_labq, &a_eimg); MUST("nDeltaEImages");
    float* map = malloc(n * 4);
    struct _jobject a_map = {map, (int)n};
    const double de_u8 = CLS(nErrorImage)(env, NULL, ctx, &a_out, &a_map); MUST("nErrorImage");
    int map_equal = 1;
    for (size_t j = 0; j < n; ++j) if (memcmp(&map[j], &eimg[4 * j], 4) || memcmp(&map[j], &eimg[4 * j + 2], 4) || eimg[4 * j + 3] != 0.0f) { map_equal = 0; break; }
    const unsigned long long map_hash = fnv(map, n * 4);
    const double de_f32 = CLS(nErrorImageFloat)(env, NULL, ctx, &a_qr, &a_qg, &a_qb, &a_map); MUST("nErrorImageFloat");
    const int map_equal_f32 = fnv(map, n * 4) == map_hash;

    /* ---- the same image as the plugin's float planes: same integers expected */
    CLS(nSetImageFloat)(env, NULL, ctx, &a_r, &a_g, &a_b, w, h, 0); MUST("nSetImageFloat");
    CLS(nEvalPalettes)(env, NULL, ctx, &a_pal, B, K, 0, 0, &a_errf, &a_cnt2); MUST("nEvalPalettes (float image)");

    /* ---- the library-side search with a progress listener; then one the listener stops at its first call */
    const int imax = 45, P = 4;
    jint ipv[6] = {P, imax, 20, 1, 0, 0};
    jfloat fpv[7] = {2.0f, 0.75f, 0.15f, 20.0f, 0.9f, 100.0f, 5.3f};
    float* best = malloc(sizeof(float) * 4 * K);
    double berr[1] = {0}, *trace = malloc(sizeof(double) * (imax + 1) * P);
    struct _jobject a_ip = {ipv, 6}, a_fp = {fpv, 7}, a_best = {best, 4 * K}, a_berr = {berr, 1}, a_trace = {trace, (imax + 1) * P}, listener = {NULL, 0};
    const int its = CLS(nFindBestQuantization)(env, NULL, ctx, K, &a_ip, &a_fp, 4242, &a_best, &a_berr, &a_trace, &listener); MUST("nFindBestQuantization");
    const int prog_calls = g_prog_calls, prog_first = g_prog_first, prog_last = g_prog_last, prog_max = g_prog_max;
    const double prog_best = g_prog_best, best_error = berr[0];
    const unsigned long long trace_hash = fnv(trace, sizeof(double) * (imax + 1) * P), best_hash = fnv(best, sizeof(float) * 4 * K);
    g_prog_calls = 0; g_stop_at_first = 1;
    ipv[1] = 100000;
    const int its_stopped = CLS(nFindBestQuantization)(env, NULL, ctx, K, &a_ip, &a_fp, 4242, &a_best, &a_berr, NULL, &listener); MUST("nFindBestQuantization (stopped)");
    g_stop_at_first = 0;

    /* ---- error paths: nothing may reach the library, every one must surface as a Java exception */
    int threw_short = 0, threw_null = 0, threw_bad_k = 0, threw_short_out = 0;
    char cls_short[100] = "", cls_null[100] = "";
    struct _jobject a_short = {rgb, (int)(n * 3) - 1};
    CLS(nSetImage)(env, NULL, ctx, &a_short, w, h, 0);
    threw_short = g_thrown; snprintf(cls_short, sizeof cls_short, "%s", g_cls); g_thrown = 0;
    CLS(nSetImage)(env, NULL, ctx, NULL, w, h, 0);
    threw_null = g_thrown; snprintf(cls_null, sizeof cls_null, "%s", g_cls); g_thrown = 0;
    CLS(nEvalPalettes)(env, NULL, ctx, &a_pal, 1, 100000, 0, 0, &a_err, &a_cnt);      /* palettes / counts far too short for K */
    threw_bad_k = g_thrown; g_thrown = 0;
    struct _jobject a_small_out = {out, 16};
    CLS(nQuantize)(env, NULL, ctx, &a_pal, K, 0, &a_small_out, NULL);
    threw_short_out = g_thrown; g_thrown = 0;
    /* deltaETypes: CIEDE2000 is refused (an empty stub in the reference), CIE94 and CIE76 are accepted */
    CLS(nSetDeltaE)(env, NULL, ctx, 2);
    const int threw_ciede2000 = g_thrown; g_thrown = 0;
    CLS(nSetDeltaE)(env, NULL, ctx, 1); MUST("nSetDeltaE(CIE94)");
    const double de94 = CLS(nDeltaEImages)(env, NULL, ctx, &a_lab, &a_labq, NULL); MUST("nDeltaEImages (CIE94)");
    CLS(nSetDeltaE)(env, NULL, ctx, 0); MUST("nSetDeltaE(CIE76)");
    /* the image must still be intact after the refused calls */
    jlong* err_again = calloc(B, sizeof(jlong));
    struct _jobject a_err_again = {err_again, B};
    CLS(nEvalPalettes)(env, NULL, ctx, &a_pal, B, K, 0, 0, &a_err_again, &a_cnt2); MUST("nEvalPalettes (after refused calls)");
    CLS(nDestroy)(env, NULL, ctx);

    printf("{\"pixels\": %lld, \"devices\": %d, \"image_hash\": %llu, \"f32_hash\": %llu, ", (long long)px, ndev, img_hash, f32_hash);
    print_ll("err_fx", err, B); print_ll("counts", cnt, B * K); print_ll("err_fx_float_image", errf, B); print_ll("err_fx_again", err_again, B);
    print_ll("sc_err_fx", errs, B); print_ll("sc_counts", cnts, B * K); print_ll("sc_err_fx_installed_bank", errs2, B);
    print_ll("sc_err_fx_installed_original", errs3, B);
    printf("\"de_lab\": \"%a\", \"de_u8\": \"%a\", \"de_f32\": \"%a\", \"map_hash\": %llu, \"map_equal\": %d, \"map_equal_f32\": %d, ", de_lab, de_u8, de_f32,
           map_hash, map_equal, map_equal_f32);
    printf("\"iterations\": %d, \"best_error\": \"%a\", \"trace_hash\": %llu, \"best_hash\": %llu, \"progress\": [%d, %d, %d, %d], \"progress_best\": \"%a\", ",
           its, best_error, trace_hash, best_hash, prog_calls, prog_first, prog_last, prog_max, prog_best);
    printf("\"iterations_stopped\": %d, \"threw_short\": %d, \"cls_short\": \"%s\", \"threw_null\": %d, \"cls_null\": \"%s\", \"threw_bad_k\": %d, \"threw_short_out\": %d, "
           "\"threw_ciede2000\": %d, \"de94_is_nan_or_positive\": %d}\n",
           its_stopped, threw_short, cls_short, threw_null, cls_null, threw_bad_k, threw_short_out, threw_ciede2000, (de94 != de94 || de94 > 0) ? 1 : 0);
    return 0;
}
