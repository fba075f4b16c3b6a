/* multi_c_bench.c — what a single-process host (a JVM behind the JNI shim, any C program) gets from a multi-device context: no
 * Python, no torch, no CUDA headers, only include/hq_b200.h.  hq_create_multi replaces JavaCL.createBestContext() + one queue
 * (ImageManipulation.java:58-59).  Prints one JSON line per device count:
 *   - the headline step (K colours, B candidates per call, 3840 x 2160 rows PER DEVICE = weak scaling) through hq_eval_palettes
 *     with host buffers, and the same through the exact pruned kernel;
 *   - a default-parameter search (512 x 512, 8 colours, population 4) sharded over the devices: us per iteration, i.e. one fused
 *     launch per device with the exchange over NVLink peer memory inside it (HQ_PEER_EXCHANGE=0: ncclAllReduce instead).
 *   gcc -O2 -Iinclude -o tools/multi_c_bench tools/multi_c_bench.c -Lhybridquantization_b200 -lhq_b200 -Wl,-rpath,$PWD/hybridquantization_b200
 *   tools/multi_c_bench <max devices> [K B steps]                                                                                  */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "hq_b200.h"

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }
static uint64_t lcg(uint64_t* s) { *s = *s * 6364136223846793005ull + 1442695040888963407ull; return *s >> 33; }
#define CHECK(call) do { int rc__ = (call); if (rc__ != HQ_OK) { fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #call, rc__, hq_last_error(ctx)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int maxdev = argc > 1 ? atoi(argv[1]) : 2;
    const int K = argc > 2 ? atoi(argv[2]) : 256, B = argc > 3 ? atoi(argv[3]) : 64, steps = argc > 4 ? atoi(argv[4]) : 10;
    const int w = 3840, rows_per_dev = 2160;
    for (int ndev = 1; ndev <= maxdev; ndev *= 2) {
        hq_ctx* ctx = NULL;
        int devs[64];
        for (int i = 0; i < ndev; ++i) devs[i] = i;
        CHECK(hq_create_multi(devs, ndev, &ctx));
        const int h = rows_per_dev * ndev;
        const size_t n = (size_t)w * h;
        uint8_t* img = malloc(n * 3);
        uint64_t s = 0x48510003ull;
        for (size_t i = 0; i < n * 3; ++i) img[i] = (uint8_t)(lcg(&s) & 255);
        CHECK(hq_set_image_u8(ctx, img, w, h, HQ_WHITEPOINT_D65));
        hq_java_random rng;
        hq_java_random_seed(&rng, 77760);
        float* pal = malloc(sizeof(float) * 4 * (size_t)B * K);
        for (int b = 0; b < B; ++b) hq_swasa_generate_random_colors(&rng, K, pal + (size_t)b * K * 4);
        int64_t* err = malloc(sizeof(int64_t) * B);
        uint64_t* cnt = malloc(sizeof(uint64_t) * (size_t)B * K);
        double t_ex = 0, t_pr = 0;
        for (int mode = 0; mode < 2; ++mode) {
            const int flags = mode ? HQ_EVAL_PRUNE : 0;
            for (int i = 0; i < 3; ++i) CHECK(hq_eval_palettes(ctx, pal, B, K, HQ_SPACE_LAB, flags, err, cnt, NULL));
            const double t0 = now_s();
            for (int i = 0; i < steps; ++i) CHECK(hq_eval_palettes(ctx, pal, B, K, HQ_SPACE_LAB, flags, err, cnt, NULL));
            *(mode ? &t_pr : &t_ex) = (now_s() - t0) / steps;
            uint64_t tot = 0;
            for (int k = 0; k < K; ++k) tot += cnt[k];
            if (tot != n) { fprintf(stderr, "counts sum %llu != %zu\n", (unsigned long long)tot, n); return 2; }
        }
        /* the plugin's defaults, sharded */
        const int sw = 512, sh = 512, sk = 8;
        uint8_t* simg = malloc((size_t)sw * sh * 3);
        s = 0x48510077ull;   /* the same small image for every device count: best_error must not depend on it */
        for (size_t i = 0; i < (size_t)sw * sh * 3; ++i) simg[i] = (uint8_t)(lcg(&s) & 255);
        CHECK(hq_set_image_u8(ctx, simg, sw, sh, HQ_WHITEPOINT_D65));
        hq_swasa_params p;
        hq_swasa_default_params(&p);
        p.imax = 3000; p.seed = 77760;
        float best[4 * 8];
        double best_err = 0, t_search = 1e30;
        int its = 0;
        for (int rep = 0; rep < 3; ++rep) {
            const double t0 = now_s();
            CHECK(hq_find_best_quantization(ctx, sk, &p, 0, best, &best_err, NULL, &its));
            const double dt = now_s() - t0;
            if (dt < t_search) t_search = dt;
        }
        int nccl = 0, size = 0;
        CHECK(hq_comm_info(ctx, NULL, &size, &nccl));
        printf("{\"host\": \"plain C, one process\", \"devices\": %d, \"peer_exchange\": %d, \"nccl\": %d, \"workload\": \"%dx%d (2160 rows per device), K=%d, %d candidates per hq_eval_palettes call, host buffers\", "
               "\"exhaustive_gpixel_per_s\": %.2f, \"exhaustive_ms_per_call\": %.3f, \"pruned_gpixel_per_s\": %.1f, \"pruned_ms_per_call\": %.3f, "
               "\"default_search_512x512_k8_p4\": {\"iterations\": %d, \"us_per_iteration\": %.2f, \"best_error\": %.9f}}\n",
               ndev, hq_comm_peers_open(ctx), nccl, w, h, K, B, (double)n * B / t_ex / 1e9, 1e3 * t_ex, (double)n * B / t_pr / 1e9, 1e3 * t_pr, its,
               1e6 * t_search / (its + 1), best_err);
        fflush(stdout);
        hq_destroy(ctx);
        free(img); free(simg); free(pal); free(err); free(cnt);
    }
    return 0;
}
