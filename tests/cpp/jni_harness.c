/* Drives java/jni/hq_jni.c WITHOUT a JVM: the shim is compiled against tests/stubs/jni.h and called through a fake JNIEnv
 * whose "arrays" are plain C buffers (a jarray is a pointer to a {data} record).  This executes every native method of
 * plugins.dbrasseur.hybridquantization.CudaImageManipulation on the GPU exactly as the JVM would invoke them — array pinning,
 * the C-ABI calls, the error path — and prints the integers for tests/test_gpu_cpp_api.py to compare with the oracle. */
#include <jni.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct _jobject { void* data; };
static int g_thrown = 0;
static char g_msg[700];
static struct _jobject g_class;
static jclass fake_FindClass(JNIEnv* env, const char* name) { (void)env; (void)name; return &g_class; }
static jint fake_ThrowNew(JNIEnv* env, jclass c, const char* msg) { (void)env; (void)c; g_thrown = 1; snprintf(g_msg, sizeof g_msg, "%s", msg); return 0; }
static void* fake_Get(JNIEnv* env, jarray a, jboolean* isCopy) { (void)env; if (isCopy) *isCopy = 0; return a->data; }
static void fake_Release(JNIEnv* env, jarray a, void* p, jint mode) { (void)env; (void)a; (void)p; (void)mode; }

#define CLS(name) Java_plugins_dbrasseur_hybridquantization_CudaImageManipulation_##name
jlong CLS(nCreate)(JNIEnv*, jclass, jint);
void CLS(nDestroy)(JNIEnv*, jclass, jlong);
jlong CLS(nPixels)(JNIEnv*, jclass, jlong);
void CLS(nSetImage)(JNIEnv*, jclass, jlong, jbyteArray, jint, jint, jint);
void CLS(nSetImageFloat)(JNIEnv*, jclass, jlong, jfloatArray, jfloatArray, jfloatArray, jint, jint, jint);
void CLS(nEvalPalettes)(JNIEnv*, jclass, jlong, jfloatArray, jint, jint, jint, jlongArray, jlongArray);
void CLS(nQuantize)(JNIEnv*, jclass, jlong, jfloatArray, jint, jint, jbyteArray);

static unsigned long long splitmix64(unsigned long long* s) {
    unsigned long long z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

int main(int argc, char** argv) {
    const int w = argc > 1 ? atoi(argv[1]) : 320, h = argc > 2 ? atoi(argv[2]) : 240, K = argc > 3 ? atoi(argv[3]) : 48, B = 3;
    const struct JNINativeInterface_ table = {fake_FindClass, fake_ThrowNew, fake_Get, fake_Release};
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
    jlong* err = calloc((size_t)B, sizeof(jlong));
    jlong* cnt = calloc((size_t)B * K, sizeof(jlong));
    unsigned char* out = malloc(n * 3);
    struct _jobject a_rgb = {rgb}, a_pal = {pal}, a_err = {err}, a_cnt = {cnt}, a_out = {out};

    const jlong ctx = CLS(nCreate)(env, &g_class, 0);
    if (g_thrown) { fprintf(stderr, "nCreate threw: %s\n", g_msg); return 2; }
    CLS(nSetImage)(env, &g_class, ctx, &a_rgb, w, h, 0);
    if (g_thrown) { fprintf(stderr, "nSetImage threw: %s\n", g_msg); return 2; }
    const jlong px = CLS(nPixels)(env, &g_class, ctx);
    CLS(nEvalPalettes)(env, &g_class, ctx, &a_pal, B, K, 0, &a_err, &a_cnt);
    if (g_thrown) { fprintf(stderr, "nEvalPalettes threw: %s\n", g_msg); return 2; }
    CLS(nQuantize)(env, &g_class, ctx, &a_pal, K, 0, &a_out);
    if (g_thrown) { fprintf(stderr, "nQuantize threw: %s\n", g_msg); return 2; }
    unsigned long long hash = 0;
    for (size_t j = 0; j < n * 3; ++j) hash = hash * 1099511628211ULL + out[j];
    /* the same image as the plugin's float planes (c/255, HybridQuantization.java:95-98): same integers expected */
    float* planes = malloc(n * 3 * sizeof(float));
    for (size_t j = 0; j < n; ++j)
        for (int c = 0; c < 3; ++c) planes[(size_t)c * n + j] = (float)(rgb[3 * j + c] / 255.0);
    jlong* errf = calloc((size_t)B, sizeof(jlong));
    struct _jobject a_r = {planes}, a_g = {planes + n}, a_b = {planes + 2 * n}, a_errf = {errf};
    CLS(nSetImageFloat)(env, &g_class, ctx, &a_r, &a_g, &a_b, w, h, 0);
    if (g_thrown) { fprintf(stderr, "nSetImageFloat threw: %s\n", g_msg); return 2; }
    CLS(nEvalPalettes)(env, &g_class, ctx, &a_pal, B, K, 0, &a_errf, &a_cnt);
    if (g_thrown) { fprintf(stderr, "nEvalPalettes (float image) threw: %s\n", g_msg); return 2; }
    /* error path: K beyond the limit must surface as a Java exception, not as silence */
    CLS(nEvalPalettes)(env, &g_class, ctx, &a_pal, 1, 100000, 0, &a_err, &a_cnt);
    const int threw_on_bad_k = g_thrown;
    CLS(nDestroy)(env, &g_class, ctx);

    printf("{\"pixels\": %lld, \"image_hash\": %llu, \"threw_on_bad_k\": %d, \"err_fx\": [", (long long)px, hash, threw_on_bad_k);
    for (int b = 0; b < B; ++b) printf("%s%lld", b ? ", " : "", (long long)err[b]);
    printf("], \"err_fx_float_image\": [");
    for (int b = 0; b < B; ++b) printf("%s%lld", b ? ", " : "", (long long)errf[b]);
    printf("], \"counts\": [");
    for (int i = 0; i < B * K; ++i) printf("%s%lld", i ? ", " : "", (long long)cnt[i]);
    printf("]}\n");
    return 0;
}
