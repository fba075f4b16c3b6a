/* multi_c_test.c — a plain C host (no Python, no torch, no CUDA headers) drives a sharded search on N GPUs of one process
 * through the C ABI alone: hq_create_multi replaces JavaCL.createBestContext() + one queue (ImageManipulation.java:58-59).
 * Everything the N-device context returns must equal, integer for integer and bit for bit, what a 1-device context
 * returns for the same image: evaluation totals (exhaustive and pruned), the S-CIELAB chain, Lab / S-CIELAB planes,
 * the annealing trajectories of both cost models, the output image and the error image.
 *
 *   multi_c_test <ndev> [width rows K B]        exit 0 and "MULTI_C_TEST OK ..." on success
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hq_b200.h"

#define CHECK(call)                                                                                  \
    do {                                                                                             \
        int rc__ = (call);                                                                           \
        if (rc__ != HQ_OK) {                                                                         \
            fprintf(stderr, "%s:%d %s -> %d: %s | %s\n", __FILE__, __LINE__, #call, rc__, hq_last_error(one), multi ? hq_last_error(multi) : hq_last_error(NULL)); \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)
#define SAME(a, b, bytes, what)                                                  \
    do {                                                                         \
        if (memcmp((a), (b), (bytes)) != 0) { fprintf(stderr, "MISMATCH: %s\n", what); return 2; } \
        ++checks;                                                                \
    } while (0)

static uint64_t lcg(uint64_t* s) { *s = *s * 6364136223846793005ull + 1442695040888963407ull; return *s >> 33; }

int main(int argc, char** argv) {
    const int ndev = argc > 1 ? atoi(argv[1]) : 2;
    const int w = argc > 2 ? atoi(argv[2]) : 1031, h = argc > 3 ? atoi(argv[3]) : 517;
    const int K = argc > 4 ? atoi(argv[4]) : 64, B = argc > 5 ? atoi(argv[5]) : 5;
    const size_t n = (size_t)w * h;
    int checks = 0;
    hq_ctx *one = NULL, *multi = NULL;
    int devs[64];
    for (int i = 0; i < ndev && i < 64; ++i) devs[i] = i;

    /* a smooth gradient + noise image: near-ties and crowded colours */
    uint8_t* img = malloc(n * 3);
    uint64_t s = 0x48510000ull + 77;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c) {
                int v = (c == 0 ? 255 * x / w : c == 1 ? 255 * y / h : 255 * (x + y) / (w + h)) + (int)(lcg(&s) % 17) - 8;
                img[((size_t)y * w + x) * 3 + c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }

    CHECK(hq_create(0, &one));
    CHECK(hq_create_multi(devs, ndev, &multi));
    if (hq_multi_device_count(multi) != ndev) { fprintf(stderr, "device count %d\n", hq_multi_device_count(multi)); return 3; }
    int nccl_version = 0, size = 0;
    CHECK(hq_comm_info(multi, NULL, &size, &nccl_version));
    CHECK(hq_set_image_u8(one, img, w, h, HQ_WHITEPOINT_D65));
    CHECK(hq_set_image_u8(multi, img, w, h, HQ_WHITEPOINT_D65));
    if (hq_image_pixels(multi) != n || hq_image_pixels(one) != n) { fprintf(stderr, "pixel count\n"); return 3; }

    /* entries that cannot work on a multi-device context refuse loudly */
    if (ndev > 1) {
        if (hq_set_image_u8_sharded(multi, img, w, h, 0, 0, 0, h, 0) != HQ_ERR_UNSUPPORTED) { fprintf(stderr, "sharded upload not refused\n"); return 3; }
        if (hq_eval_palettes_device(multi, img, 1, K, 0, 0, img, NULL) != HQ_ERR_UNSUPPORTED) { fprintf(stderr, "device eval not refused\n"); return 3; }
    }

    hq_java_random rng;
    hq_java_random_seed(&rng, 77760);
    float* pal = malloc(sizeof(float) * 4 * (size_t)B * K);
    for (int b = 0; b < B; ++b) hq_swasa_generate_random_colors(&rng, K, pal + (size_t)b * K * 4);

    /* Lab planes */
    float *lab1 = malloc(n * 12), *labN = malloc(n * 12);
    CHECK(hq_get_lab(one, lab1));
    CHECK(hq_get_lab(multi, labN));
    SAME(lab1, labN, n * 12, "Lab planes");

    /* evaluation totals: exhaustive, pruned, sRGB space; small (direct host I/O) and large (DMA) transfers */
    int64_t *e1 = malloc(8 * (size_t)B), *eN = malloc(8 * (size_t)B);
    uint64_t *c1 = malloc(8 * (size_t)B * K), *cN = malloc(8 * (size_t)B * K);
    int64_t *s1 = malloc(8 * 3 * (size_t)B * K), *sN = malloc(8 * 3 * (size_t)B * K);
    const int modes[4][2] = {{HQ_SPACE_LAB, HQ_EVAL_SUMS}, {HQ_SPACE_LAB, HQ_EVAL_SUMS | HQ_EVAL_PRUNE}, {HQ_SPACE_SRGB, HQ_EVAL_SUMS}, {HQ_SPACE_LAB, HQ_EVAL_SUMS | HQ_EVAL_FORCE_PREFILTER}};
    for (int m = 0; m < 4; ++m) {
        CHECK(hq_eval_palettes(one, pal, B, K, modes[m][0], modes[m][1], e1, c1, s1));
        CHECK(hq_eval_palettes(multi, pal, B, K, modes[m][0], modes[m][1], eN, cN, sN));
        SAME(e1, eN, 8 * (size_t)B, "err_fx"); SAME(c1, cN, 8 * (size_t)B * K, "counts"); SAME(s1, sN, 24 * (size_t)B * K, "sums");
        uint64_t tot = 0;
        for (int k = 0; k < K; ++k) tot += cN[k];
        if (tot != n) { fprintf(stderr, "counts sum %llu != %zu\n", (unsigned long long)tot, n); return 2; }
    }
    {   /* a population beyond the direct-I/O thresholds (64 KB of palettes) */
        const int B2 = 24, K2 = 300;
        float* pal2 = malloc(sizeof(float) * 4 * (size_t)B2 * K2);
        for (int b = 0; b < B2; ++b) hq_swasa_generate_random_colors(&rng, K2, pal2 + (size_t)b * K2 * 4);
        int64_t *f1 = malloc(8 * B2), *fN = malloc(8 * B2);
        uint64_t *g1 = malloc(8 * (size_t)B2 * K2), *gN = malloc(8 * (size_t)B2 * K2);
        CHECK(hq_eval_palettes(one, pal2, B2, K2, HQ_SPACE_LAB, 0, f1, g1, NULL));
        CHECK(hq_eval_palettes(multi, pal2, B2, K2, HQ_SPACE_LAB, 0, fN, gN, NULL));
        SAME(f1, fN, 8 * B2, "err_fx (large population)"); SAME(g1, gN, 8 * (size_t)B2 * K2, "counts (large population)");
        free(pal2); free(f1); free(fN); free(g1); free(gN);
    }

    /* the reference-faithful chain (S-CIELAB spatial filters): halo rows are uploaded by the library itself */
    CHECK(hq_scielab_configure(one, 72, 45.0f));
    CHECK(hq_scielab_configure(multi, 72, 45.0f));
    float *sc1 = malloc(n * 12), *scN = malloc(n * 12);
    CHECK(hq_scielab_get_image(one, sc1));
    CHECK(hq_scielab_get_image(multi, scN));
    SAME(sc1, scN, n * 12, "S-CIELAB planes of the original");
    CHECK(hq_eval_palettes_scielab(one, pal, B, K, HQ_SPACE_SRGB, e1, c1));
    CHECK(hq_eval_palettes_scielab(multi, pal, B, K, HQ_SPACE_SRGB, eN, cN));
    SAME(e1, eN, 8 * (size_t)B, "S-CIELAB err_fx"); SAME(c1, cN, 8 * (size_t)B * K, "S-CIELAB counts");

    /* the annealing search, both cost models: trajectory, best palette, best error */
    for (int model = 0; model < 2; ++model) {
        hq_swasa_params p;
        hq_swasa_default_params(&p);
        p.population = 4; p.imax = model ? 12 : 60; p.seed = 2024 + model;
        p.cost_model = model ? HQ_COST_SCIELAB : HQ_COST_LAB; p.space = model ? HQ_SPACE_SRGB : HQ_SPACE_LAB;
        const size_t nt = (size_t)(p.imax + 1) * p.population;
        double *t1 = malloc(8 * nt), *tN = malloc(8 * nt), b1 = 0, bN = 0;
        float *best1 = malloc(16 * (size_t)K), *bestN = malloc(16 * (size_t)K);
        int it1 = 0, itN = 0;
        CHECK(hq_find_best_quantization(one, K, &p, 0, best1, &b1, t1, &it1));
        CHECK(hq_find_best_quantization(multi, K, &p, 0, bestN, &bN, tN, &itN));
        SAME(t1, tN, 8 * nt, model ? "S-CIELAB search trajectory" : "LAB search trajectory");
        SAME(best1, bestN, 16 * (size_t)K, "best palette"); SAME(&b1, &bN, 8, "best error");
        if (it1 != p.imax || itN != p.imax) { fprintf(stderr, "iterations %d %d\n", it1, itN); return 2; }
        if (model == 0) {   /* output image of the found palette + error image of that output */
            uint8_t *q1 = malloc(n * 3), *qN = malloc(n * 3);
            uint16_t *i1 = malloc(n * 2), *iN = malloc(n * 2);
            float *f1 = malloc(n * 16), *fN = malloc(n * 16);
            CHECK(hq_quantize(one, best1, K, HQ_SPACE_LAB, q1, f1, i1));
            CHECK(hq_quantize(multi, bestN, K, HQ_SPACE_LAB, qN, fN, iN));
            SAME(q1, qN, n * 3, "quantised image"); SAME(i1, iN, n * 2, "index image"); SAME(f1, fN, n * 16, "float image");
            float *m1 = malloc(n * 4), *mN = malloc(n * 4);
            uint8_t *u1 = malloc(n), *uN = malloc(n);
            double d1 = 0, dN = 0;
            CHECK(hq_error_image(one, q1, m1, u1, &d1));
            CHECK(hq_error_image(multi, qN, mN, uN, &dN));
            SAME(m1, mN, n * 4, "error map"); SAME(u1, uN, n, "error map u8"); SAME(&d1, &dN, 8, "mean dE");
            free(q1); free(qN); free(i1); free(iN); free(f1); free(fN); free(m1); free(mN); free(u1); free(uN);
        }
        free(t1); free(tN); free(best1); free(bestN);
    }
    printf("MULTI_C_TEST OK ndev=%d comm_size=%d nccl=%d peers=%d image=%dx%d K=%d B=%d checks=%d\n", ndev, size, nccl_version, hq_comm_peers_open(multi), w, h, K, B, checks);
    hq_destroy(multi);
    hq_destroy(one);
    free(img); free(pal); free(lab1); free(labN); free(e1); free(eN); free(c1); free(cN); free(s1); free(sN); free(sc1); free(scN);
    return 0;
}
