/* hq_oracle.c — CPU ORACLE (test infrastructure only; see hq_oracle.h for the rules).
 *
 * Plain C restatement of the reference's algorithm for the hot path, pinned bit for bit against the
 * reference's own sources compiled for the CPU (oracle/_ref, tests/test_ref_pinning.py; see hq_oracle.h).
 * Compile with -ffp-contract=off: every fp32 expression below is meant to round after
 * each operation, as Java float arithmetic does; fused multiply-adds appear only where
 * they are written as fmaf().  Transcendentals are glibc's pow/exp/tanh in double, the
 * counterpart of Java's Math.pow/exp/tanh.
 *
 * Citations: File:line under /root/reference/src/plugins/dbrasseur/hybridquantization/.
 */
#define _GNU_SOURCE
#include "hq_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ colour arithmetic */

/* ScielabProcessor.java:20-21 */
static const float WHITE[2][3] = {{0.95047f, 1.0f, 1.0883f}, {0.966797f, 1.0f, 0.825188f}};

/* ScielabProcessor.java:59-61 — evaluated at run time in fp32, in Java's order:
 *   LABDELTA = 6f/29f; LABDELTA2 = LABDELTA*LABDELTA; LABDELTA3 = LABDELTA2*LABDELTA;
 * :301 uses 3*LABDELTA2 (int*float -> float) and 4.0f/29.0f. */
static float C_LABDELTA3, C_3LABDELTA2, C_4_29;
__attribute__((constructor)) static void init_constants(void) {
    volatile float six = 6.0f, tn = 29.0f, four = 4.0f, three = 3.0f;
    volatile float d = six / tn;
    volatile float d2 = d * d;
    volatile float d3 = d2 * d;
    volatile float t3 = three * d2;
    volatile float f429 = four / tn;
    C_LABDELTA3 = d3; C_3LABDELTA2 = t3; C_4_29 = f429;
}

float hqo_lab_constants(int which) {
    if (which == 0) return C_LABDELTA3;
    if (which == 1) return C_3LABDELTA2;
    return C_4_29;
}

/* HybridQuantization.java:95 (Icy convertToType(FLOAT, rescale) — third-party, unpinned):
 * the build defines u8 -> float as (float)(c / 255.0). */
float hqo_u8_to_unit(unsigned c) { return (float)((double)c / 255.0); }

/* (float)Math.pow(b, 2.4f): the float literal widens to 2.4000000953674316 */
float hqo_pow_2p4(float b) { return (float)pow((double)b, (double)2.4f); }
/* (float)Math.pow(t, 1.0 / 3.0) */
float hqo_cbrt_pow(float t) { return (float)pow((double)t, 1.0 / 3.0); }

/* ScielabProcessor.java:282 */
float hqo_srgb_decode(float c) {
    if (c <= 0.04045f) return c / 12.92f;
    return hqo_pow_2p4((c + 0.055f) / 1.055f);
}

/* ScielabProcessor.java:279-291 */
void hqo_srgb_to_opp(const float rgb[3], float opp[3]) {
    const float R = hqo_srgb_decode(rgb[0]);
    const float G = hqo_srgb_decode(rgb[1]);
    const float B = hqo_srgb_decode(rgb[2]);
    opp[0] = 0.26641335000823f * R + 0.60316740257478f * G + 0.0011333302293f * B;
    opp[1] = -0.12197400229389f * R + 0.05598088396616f * G + 0.01326365114329f * B;
    opp[2] = -0.08033445917708f * R + -0.33146741170125f * G + 0.44913244757774f * B;
}

static float lab_f(float t) { /* ScielabProcessor.java:301 */
    if (t > C_LABDELTA3) return hqo_cbrt_pow(t);
    return (t / C_3LABDELTA2) + C_4_29;
}

/* ScielabProcessor.java:293-311 */
void hqo_opp_to_lab(const float opp[3], int whitepoint, float lab[3]) {
    const float* ill = WHITE[whitepoint == HQO_WHITE_D50 ? 1 : 0];
    const float X = 0.97959616044562807864f * opp[0] + -1.5347157012664408981f * opp[1] +
                    0.44459764330437399288f * opp[2];
    const float Y = 1.188977906742323787f * opp[0] + 0.7643549575179937615f * opp[1] +
                    0.13512574791125839373f * opp[2];
    const float Z = 1.2318333139247290457f * opp[0] + 1.1631592597636512884f * opp[1] +
                    2.0784075888008567862f * opp[2];
    const float fx = lab_f(X / ill[0]);
    const float fy = lab_f(Y / ill[1]);
    const float fz = lab_f(Z / ill[2]);
    lab[0] = 116.0f * fy - 16.0f;
    lab[1] = 500.0f * (fx - fy);
    lab[2] = 200.0f * (fy - fz);
}

/* ScielabProcessor.java:432 : OpptoLab(sRGBtoOpp(px)) */
void hqo_srgb_to_lab(const float rgb[3], int whitepoint, float lab[3]) {
    float opp[3];
    hqo_srgb_to_opp(rgb, opp);
    hqo_opp_to_lab(opp, whitepoint, lab);
}

/* ------------------------------------------------------------------ threading helper */
typedef void (*range_fn)(void* ctx, size_t lo, size_t hi, int tid);
typedef struct { range_fn fn; void* ctx; size_t lo, hi; int tid; } range_job;
static void* range_tramp(void* p) {
    range_job* j = (range_job*)p;
    j->fn(j->ctx, j->lo, j->hi, j->tid);
    return NULL;
}
static void parallel_ranges(size_t n, int threads, range_fn fn, void* ctx) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    if (threads == 1) { fn(ctx, 0, n, 0); return; }
    pthread_t th[256];
    range_job jobs[256];
    for (int t = 0; t < threads; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].tid = t;
        jobs[t].lo = n * (size_t)t / (size_t)threads;
        jobs[t].hi = n * (size_t)(t + 1) / (size_t)threads;
        pthread_create(&th[t], NULL, range_tramp, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------ image planes */
typedef struct {
    const uint8_t* rgb; int wp;
    float *ur, *ug, *ub, *ll, *la, *lb;
    float unit[256], lin[256];
} planes_ctx;

static void planes_range(void* p, size_t lo, size_t hi, int tid) {
    (void)tid;
    planes_ctx* c = (planes_ctx*)p;
    for (size_t i = lo; i < hi; ++i) {
        const unsigned r = c->rgb[3 * i], g = c->rgb[3 * i + 1], b = c->rgb[3 * i + 2];
        if (c->ur) { c->ur[i] = c->unit[r]; c->ug[i] = c->unit[g]; c->ub[i] = c->unit[b]; }
        if (c->ll) {
            /* identical to hqo_srgb_to_lab on (unit[r],unit[g],unit[b]); the decode of a
             * u8 channel only has 256 possible values, so it is tabulated */
            const float R = c->lin[r], G = c->lin[g], B = c->lin[b];
            float opp[3], lab[3];
            opp[0] = 0.26641335000823f * R + 0.60316740257478f * G + 0.0011333302293f * B;
            opp[1] = -0.12197400229389f * R + 0.05598088396616f * G + 0.01326365114329f * B;
            opp[2] = -0.08033445917708f * R + -0.33146741170125f * G + 0.44913244757774f * B;
            hqo_opp_to_lab(opp, c->wp, lab);
            c->ll[i] = lab[0]; c->la[i] = lab[1]; c->lb[i] = lab[2];
        }
    }
}

void hqo_image_planes(const uint8_t* rgb, size_t n, int whitepoint, float* unit_r, float* unit_g,
                      float* unit_b, float* lab_l, float* lab_a, float* lab_b, int threads) {
    planes_ctx c;
    c.rgb = rgb; c.wp = whitepoint;
    c.ur = unit_r; c.ug = unit_g; c.ub = unit_b; c.ll = lab_l; c.la = lab_a; c.lb = lab_b;
    for (unsigned v = 0; v < 256; ++v) {
        c.unit[v] = hqo_u8_to_unit(v);
        c.lin[v] = hqo_srgb_decode(c.unit[v]);
    }
    parallel_ranges(n, threads, planes_range, &c);
}

/* float image planes (HybridQuantization.java:95,98: im.getDataXYCAsFloat(), any Icy type rescaled to [0,1]) */
typedef struct { const float* unit3; size_t n; int wp; float* lab3; } planes_f32_ctx;
static void planes_f32_range(void* p, size_t lo, size_t hi, int tid) {
    (void)tid;
    planes_f32_ctx* c = (planes_f32_ctx*)p;
    for (size_t i = lo; i < hi; ++i) {
        const float rgb[3] = {c->unit3[i], c->unit3[c->n + i], c->unit3[2 * c->n + i]};
        float lab[3];
        hqo_srgb_to_lab(rgb, c->wp, lab);
        c->lab3[i] = lab[0]; c->lab3[c->n + i] = lab[1]; c->lab3[2 * c->n + i] = lab[2];
    }
}
void hqo_image_planes_f32(const float* unit3, size_t n, int whitepoint, float* lab3, int threads) {
    planes_f32_ctx c = {unit3, n, whitepoint, lab3};
    parallel_ranges(n, threads, planes_f32_range, &c);
}

/* ------------------------------------------------------------------ assign + reduce */
static int64_t to_fx(float v) { return (int64_t)llrintf(v * 16777216.0f); }

/* squared distance as the build pins OpenCL distance() (OptimizedConvolution.cl:180,185):
 * fp32 differences, dx*dx, then fma(dy,dy,.), fma(dz,dz,.) */
static inline float dist2(float x0, float x1, float x2, float p0, float p1, float p2) {
    const float d0 = x0 - p0, d1 = x1 - p1, d2 = x2 - p2;
    return fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
}

typedef struct {
    const float *f0, *f1, *f2; /* assignment features [n] */
    const float *l0, *l1, *l2; /* Lab planes [n] */
    const float* pal_feat;     /* [B][K][3] assignment features of the palettes */
    const float* pal_lab;      /* [B][K][3] */
    int B, K, space;
    size_t n;
    uint16_t* idx;             /* [B][n] or NULL */
    int threads;
    int64_t* t_err;            /* [threads][B] */
    uint64_t* t_cnt;           /* [threads][B][K] */
    int64_t* t_sum;            /* [threads][B][K][3] */
} assign_ctx;

#define BLK 256
static void assign_range(void* p, size_t lo, size_t hi, int tid) {
    assign_ctx* c = (assign_ctx*)p;
    const int K = c->K;
    float best[BLK];
    int bidx[BLK];
    for (int b = 0; b < c->B; ++b) {
        const float* pf = c->pal_feat + (size_t)b * K * 3;
        const float* pl = c->pal_lab + (size_t)b * K * 3;
        int64_t err = 0;
        uint64_t* cnt = c->t_cnt + ((size_t)tid * c->B + b) * K;
        int64_t* sum = c->t_sum + ((size_t)tid * c->B + b) * K * 3;
        for (size_t base = lo; base < hi; base += BLK) {
            const int m = (int)((hi - base) < BLK ? (hi - base) : BLK);
            const float* x0 = c->f0 + base; const float* x1 = c->f1 + base; const float* x2 = c->f2 + base;
            /* OptimizedConvolution.cl:179-192: start from colour 0, replace on strict < */
            for (int i = 0; i < m; ++i) { best[i] = dist2(x0[i], x1[i], x2[i], pf[0], pf[1], pf[2]); bidx[i] = 0; }
            for (int k = 1; k < K; ++k) {
                const float p0 = pf[3 * k], p1 = pf[3 * k + 1], p2 = pf[3 * k + 2];
                for (int i = 0; i < m; ++i) {
                    const float d = dist2(x0[i], x1[i], x2[i], p0, p1, p2);
                    const int lt = d < best[i];
                    best[i] = lt ? d : best[i];
                    bidx[i] = lt ? k : bidx[i];
                }
            }
            for (int i = 0; i < m; ++i) {
                const size_t px = base + i;
                const int k = bidx[i];
                float d2v = best[i];
                if (c->space == HQO_SPACE_SRGB) /* score in Lab (OptimizedConvolution.cl:209) */
                    d2v = dist2(c->l0[px], c->l1[px], c->l2[px], pl[3 * k], pl[3 * k + 1], pl[3 * k + 2]);
                err += to_fx(sqrtf(d2v));
                cnt[k] += 1;
                sum[3 * k] += to_fx(c->l0[px]);
                sum[3 * k + 1] += to_fx(c->l1[px]);
                sum[3 * k + 2] += to_fx(c->l2[px]);
                if (c->idx) c->idx[(size_t)b * c->n + px] = (uint16_t)k;
            }
        }
        c->t_err[(size_t)tid * c->B + b] = err;
    }
}

void hqo_assign_reduce_planes(const float* unit_rgb3, const float* lab3, size_t n, int whitepoint,
                              const float* palettes, int B, int K, int space, int64_t* err_fx,
                              uint64_t* counts, int64_t* sums_fx, uint16_t* idx, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    assign_ctx c;
    memset(&c, 0, sizeof c);
    float* pal_lab = (float*)malloc(sizeof(float) * (size_t)B * K * 3);
    float* pal_rgb = (float*)malloc(sizeof(float) * (size_t)B * K * 3);
    for (size_t j = 0; j < (size_t)B * K; ++j) {
        const float* s = palettes + 4 * j;
        pal_rgb[3 * j] = s[0]; pal_rgb[3 * j + 1] = s[1]; pal_rgb[3 * j + 2] = s[2];
        hqo_srgb_to_lab(s, whitepoint, pal_lab + 3 * j);
    }
    c.l0 = lab3; c.l1 = lab3 + n; c.l2 = lab3 + 2 * n;
    if (space == HQO_SPACE_SRGB) { c.f0 = unit_rgb3; c.f1 = unit_rgb3 + n; c.f2 = unit_rgb3 + 2 * n; c.pal_feat = pal_rgb; }
    else { c.f0 = c.l0; c.f1 = c.l1; c.f2 = c.l2; c.pal_feat = pal_lab; }
    c.pal_lab = pal_lab;
    c.B = B; c.K = K; c.space = space; c.n = n; c.idx = idx; c.threads = threads;
    c.t_err = (int64_t*)calloc((size_t)threads * B, sizeof(int64_t));
    c.t_cnt = (uint64_t*)calloc((size_t)threads * B * K, sizeof(uint64_t));
    c.t_sum = (int64_t*)calloc((size_t)threads * B * K * 3, sizeof(int64_t));
    parallel_ranges(n, threads, assign_range, &c);
    for (int b = 0; b < B; ++b) {
        int64_t e = 0;
        for (int t = 0; t < threads; ++t) e += c.t_err[(size_t)t * B + b];
        if (err_fx) err_fx[b] = e;
        for (int k = 0; k < K; ++k) {
            uint64_t cn = 0; int64_t s0 = 0, s1 = 0, s2 = 0;
            for (int t = 0; t < threads; ++t) {
                const size_t o = ((size_t)t * B + b) * K + k;
                cn += c.t_cnt[o]; s0 += c.t_sum[3 * o]; s1 += c.t_sum[3 * o + 1]; s2 += c.t_sum[3 * o + 2];
            }
            if (counts) counts[(size_t)b * K + k] = cn;
            if (sums_fx) { int64_t* d = sums_fx + ((size_t)b * K + k) * 3; d[0] = s0; d[1] = s1; d[2] = s2; }
        }
    }
    free(c.t_err); free(c.t_cnt); free(c.t_sum); free(pal_lab); free(pal_rgb);
}

void hqo_assign_reduce(const uint8_t* rgb, size_t n, int whitepoint, const float* palettes, int B,
                       int K, int space, int64_t* err_fx, uint64_t* counts, int64_t* sums_fx,
                       uint16_t* idx, int threads) {
    float* lab = (float*)malloc(sizeof(float) * 3 * (n ? n : 1));
    float* unit = (float*)malloc(sizeof(float) * 3 * (n ? n : 1));
    hqo_image_planes(rgb, n, whitepoint, unit, unit + n, unit + 2 * n, lab, lab + n, lab + 2 * n, threads);
    hqo_assign_reduce_planes(unit, lab, n, whitepoint, palettes, B, K, space, err_fx, counts, sums_fx, idx, threads);
    free(lab); free(unit);
}

/* ImageManipulation.java:712 : averageArray(err) + computePenalty(used)
 *   averageArray: double sum / length (:736-752);  computePenalty: += delta per unused (SWASA.java:74-82)
 * The build's sum is the 2^-24 fixed-point integer (exact, order independent). */
double hqo_cost(int64_t err_fx, const uint64_t* counts, int K, uint64_t n_total, float delta) {
    double penalty = 0;
    for (int k = 0; k < K; ++k)
        if (counts[k] == 0) penalty += delta;
    const double sum = (double)err_fx * (1.0 / 16777216.0);
    return sum / (double)n_total + penalty;
}

/* OptimizedConvolution.cl:147-170 + HybridQuantization.java:111-122 */
void hqo_quantize(const uint8_t* rgb, size_t n, int whitepoint, const float* palette, int K,
                  int space, uint8_t* out_rgb, float* out_f32, uint16_t* idx, int threads) {
    uint16_t* own = idx ? idx : (uint16_t*)malloc(sizeof(uint16_t) * (n ? n : 1));
    hqo_assign_reduce(rgb, n, whitepoint, palette, 1, K, space, NULL, NULL, NULL, own, threads);
    for (size_t i = 0; i < n; ++i) {
        const float* c = palette + 4 * (size_t)own[i];
        if (out_f32) { out_f32[4 * i] = c[0]; out_f32[4 * i + 1] = c[1]; out_f32[4 * i + 2] = c[2]; out_f32[4 * i + 3] = c[3]; }
        if (out_rgb)
            for (int ch = 0; ch < 3; ++ch) out_rgb[3 * i + ch] = (uint8_t)(int)(c[ch] * 255.0f + 0.5f);
    }
    if (!idx) free(own);
}

/* ------------------------------------------------------------------ java.util.Random */
void hqo_rng_seed(hqo_rng* r, int64_t seed) {
    r->state = ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
}
int32_t hqo_rng_next(hqo_rng* r, int bits) {
    r->state = (r->state * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int32_t)(int64_t)(r->state >> (48 - bits));
}
float hqo_rng_next_float(hqo_rng* r) { return (float)hqo_rng_next(r, 24) / (float)(1 << 24); }
double hqo_rng_next_double(hqo_rng* r) {
    const int64_t hi = (int64_t)hqo_rng_next(r, 26) << 27;
    const int64_t lo = hqo_rng_next(r, 27);
    return (double)(hi + lo) * 0x1.0p-53;
}

/* ------------------------------------------------------------------ SWASA */
void hqo_swasa_defaults(hqo_swasa_params* p) { /* HybridQuantization.java:192-233 */
    p->population = 4; p->imax = 5000; p->iTc = 20; p->delta = 2.0f; p->convergence = 1;
    p->conv_delay = 0.75f; p->conv_spread = 0.15f; p->t0 = 20.0f; p->alpha = 0.9f;
    p->s0 = 100.0f; p->beta = 5.3f; p->whitepoint = HQO_WHITE_D65; p->space = HQO_SPACE_LAB;
    p->seed = 77760; p->cost_model = 0; p->dpi = 72; p->viewing_distance = 45.0f;
}

/* SWASA.java:40-52 */
void hqo_generate_random_colors(hqo_rng* r, int K, float* colors) {
    for (int i = 0; i < K; ++i) {
        colors[4 * i] = hqo_rng_next_float(r);
        colors[4 * i + 1] = hqo_rng_next_float(r);
        colors[4 * i + 2] = hqo_rng_next_float(r);
        colors[4 * i + 3] = 0.0f;
    }
}

/* SWASA.java:69-72 : (float)(2*s0/(1+Math.exp(beta*i/imax))) */
float hqo_max_step_width(const hqo_swasa_params* p, int ite) {
    const float arg = p->beta * (float)ite / (float)p->imax;
    const float two_s0 = 2.0f * p->s0;
    return (float)((double)two_s0 / (1.0 + exp((double)arg)));
}

static float clampf(float v, float mn, float mx) { /* SWASA.java:103-106 */
    return v > mn ? (v > mx ? mx : v) : mn;
}

/* SWASA.java:91-101 */
void hqo_generate_neighboring_colors(const hqo_swasa_params* p, hqo_rng* r, const float* colors,
                                     float* next, int K, int ite) {
    const float w = hqo_max_step_width(p, ite) / 256.0f;
    for (int i = 0; i < K; ++i) {
        for (int ch = 0; ch < 3; ++ch) {
            const float u = hqo_rng_next_float(r) * 2.0f - 1.0f;
            next[4 * i + ch] = clampf(colors[4 * i + ch] + u * w, 0.0f, 1.0f);
        }
        next[4 * i + 3] = 0.0f;
    }
}

/* ImageManipulation.java:383-591 with the scoring of SURVEY D2 (identity spatial filter) */
double hqo_find_best_quantization(const uint8_t* rgb, int w, int h, int K,
                                  const hqo_swasa_params* p, float* best_colors,
                                  double* trace_costs, int threads) {
    const size_t n = (size_t)w * h;
    const int P = p->population;
    const size_t pal = (size_t)K * 4;
    float* lab = (float*)malloc(sizeof(float) * 3 * n);
    float* unit = (float*)malloc(sizeof(float) * 3 * n);
    hqo_image_planes(rgb, n, p->whitepoint, unit, unit + n, unit + 2 * n, lab, lab + n, lab + 2 * n, threads);
    float* colors = (float*)malloc(sizeof(float) * P * pal);
    float* current = (float*)malloc(sizeof(float) * P * pal);
    double* cur_err = (double*)calloc((size_t)(P > 0 ? P : 1), sizeof(double));
    double* errs = (double*)malloc(sizeof(double) * P);
    int64_t* err_fx = (int64_t*)malloc(sizeof(int64_t) * P);
    uint64_t* counts = (uint64_t*)malloc(sizeof(uint64_t) * P * K);
    /* scoring: identity-filter Lab cost, or the reference's full S-CIELAB chain (ImageManipulation.java:635-699) */
    float *sc_filters = NULL, *sc_abs3 = NULL, *sc_orig = NULL; int sc_taps = 0;
    if (p->cost_model == 1) {
        sc_filters = (float*)malloc(sizeof(float) * 7 * 4096); sc_abs3 = (float*)malloc(sizeof(float) * 4096);
        sc_taps = hqo_scielab_filters(p->dpi, (double)p->viewing_distance, sc_filters, sc_abs3, 4096);
        sc_orig = (float*)malloc(sizeof(float) * 3 * n);
        hqo_scielab_image(rgb, w, h, p->whitepoint, sc_filters, sc_abs3, sc_taps, sc_orig, threads);
    }
#define EVAL_POPULATION(pal)                                                                                             \
    do {                                                                                                                 \
        if (p->cost_model == 1)                                                                                          \
            hqo_scielab_eval(rgb, w, h, p->whitepoint, sc_filters, sc_abs3, sc_taps, sc_orig, (pal), P, K, p->space, err_fx, counts, threads); \
        else                                                                                                             \
            hqo_assign_reduce_planes(unit, lab, n, p->whitepoint, (pal), P, K, p->space, err_fx, counts, NULL, NULL, threads); \
    } while (0)
    hqo_rng rng;
    hqo_rng_seed(&rng, p->seed);
    float temperature = p->t0; /* SWASA.reset(), :30-34 */

    for (int i = 0; i < P; ++i) hqo_generate_random_colors(&rng, K, colors + i * pal); /* :413-417 */
    EVAL_POPULATION(colors);
    for (int i = 0; i < P; ++i) {
        cur_err[i] = hqo_cost(err_fx[i], counts + (size_t)i * K, K, n, p->delta);
        if (trace_costs) trace_costs[i] = cur_err[i];
    }
    int mn = 0; /* argmin, :843-856 (strict >) */
    for (int i = 1; i < P; ++i) if (cur_err[mn] > cur_err[i]) mn = i;
    double best_err = cur_err[mn];
    memcpy(best_colors, colors + mn * pal, sizeof(float) * pal);

    for (int ite = 1; ite <= p->imax; ++ite) { /* :497 */
        if (ite % p->iTc == 0) temperature *= p->alpha; /* SWASA.java:84-89 */
        for (int j = 0; j < P; ++j)                     /* :508-511 */
            hqo_generate_neighboring_colors(p, &rng, colors + j * pal, current + j * pal, K, ite);
        EVAL_POPULATION(current);
        for (int i = 0; i < P; ++i) {
            errs[i] = hqo_cost(err_fx[i], counts + (size_t)i * K, K, n, p->delta);
            if (trace_costs) trace_costs[(size_t)ite * P + i] = errs[i];
        }
        double minerror = DBL_MAX; int minidx = 0; /* :516-517 */
        for (int i = 0; i < P; ++i) {
            if (P > 1 && errs[i] < minerror) { minerror = errs[i]; minidx = i; } /* :520-524 */
            const double dE = errs[i] - cur_err[i];
            /* SWASA.java:54-57,64-67 : the RNG is drawn only when dE > 0 */
            const int accepted = dE <= 0 || exp(-dE / (double)temperature) > hqo_rng_next_double(&rng);
            if (accepted) {
                cur_err[i] = errs[i];
                memcpy(colors + i * pal, current + i * pal, sizeof(float) * pal);
                if (cur_err[i] < best_err) {
                    best_err = cur_err[i];
                    memcpy(best_colors, current + i * pal, sizeof(float) * pal);
                }
            }
        }
        for (int i = 0; p->convergence && P > 1 && i < P; ++i) { /* :538-545 */
            /* SWASA.java:59-62 : float argument, double tanh */
            const float num = (float)ite - p->conv_delay * (float)p->imax;
            const float den = p->conv_spread * (float)p->imax;
            const double keep = -(tanh((double)(num / den))) / 2 + 0.5;
            if (!(keep > hqo_rng_next_double(&rng))) {
                cur_err[i] = minerror;
                memcpy(colors + i * pal, current + minidx * pal, sizeof(float) * pal);
            }
        }
    }
    free(lab); free(unit); free(colors); free(current); free(cur_err); free(errs); free(err_fx); free(counts);
    free(sc_filters); free(sc_abs3); free(sc_orig);
#undef EVAL_POPULATION
    return best_err;
}

/* ------------------------------------------------------------------ S-CIELAB filter bank
 * ScielabProcessor.java:66-181 (constructor), :185-254 (conv1D, resize1D, extractWithIndices, gauss).
 * Java float/double promotion is followed expression by expression. */
static const float SC_WEIGHTS[3][3] = {{1.00327f, 0.114416f, -0.117686f}, {0.616725f, 0.383275f, 0.f}, {0.567885f, 0.432115f, 0.f}}; /* :44-48 */
static const float SC_HALFWIDTHS[3][3] = {{0.05f, 0.225f, 7.0f}, {0.0685f, 0.826f, 0.f}, {0.0920f, 0.6451f, 0.f}};                  /* :49-53 */
static const int SC_COUNT[3] = {3, 2, 2};

static void sc_gauss(float halfwidth, int width, float* result) { /* :238-254 */
    const float alpha = 2 * (float)sqrt(log(2)) / (halfwidth - 1);
    const int offset = width / 2;
    double sum = 0;
    for (int i = 0; i < width; ++i) {
        const float arg = -alpha * alpha * (float)(i - offset) * (float)(i - offset);
        result[i] = (float)exp((double)arg);
        sum += result[i];
    }
    for (int i = 0; i < width; ++i) result[i] = (float)((double)result[i] / sum);
}

static void sc_conv1d(const float* data, int n, const float* filter, int fl, float* result) { /* :185-201 */
    const int offset = fl / 2;
    for (int i = 0; i < n; ++i) {
        float acc = 0.0f;
        for (int j = -offset; j <= offset; ++j)
            if (!(i + j < 0 || i + j >= n)) acc += filter[j + offset] * data[i + j];
        result[i] = acc;
    }
}

/* filters[7][taps]: O1 (three Gaussians), O2 (two), O3 (two), in the order
 * [0]=O1g1 [1]=O1g2 [2]=O1g3 [3]=O2g1 [4]=O2g2 [5]=O3g1 [6]=O3g2; abs3[taps] = |O1g3|.
 * Returns the number of taps (<= max_taps) or -1. */
int hqo_scielab_filters(int dpi, double viewing_distance, float* filters, float* abs3, int max_taps) {
    int samp = (int)llround((double)dpi / ((180 / M_PI) * atan(2.54 / viewing_distance))); /* :80 */
    int uprate = 1;
    if (samp < 224) { uprate = (int)ceil(224 * 1.0 / samp); samp *= uprate; }                /* :81-88 */
    const int width = (int)(ceil(samp / 2.0)) * 2 - 1;                                       /* :102 */
    float* g[7]; int gi = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < SC_COUNT[i]; ++j) {
            float* f = (float*)malloc(sizeof(float) * width);
            const float spread = SC_HALFWIDTHS[i][j] * (float)samp;                          /* :97 */
            sc_gauss(spread, width, f);                                                      /* :112 */
            const float w = SC_WEIGHTS[i][j];
            const float factor = (float)sqrt((double)fabsf(w)) * (w > 0 ? 1.0f : (w < 0 ? -1.0f : 0.0f)); /* :113 */
            for (int k = 0; k < width; ++k) f[k] *= factor;
            g[gi++] = f;
        }
    int taps = width;
    if (uprate > 1) {                                                                         /* :122-173 */
        const int ul = uprate * 2 - 1;
        const int fl = ul + width - 1;
        float* up = (float*)calloc(fl, sizeof(float));
        const int pad = (fl - ul) / 2;
        for (int i = 0; i < ul; ++i) up[pad + i] = (float)(uprate - abs(uprate - i - 1)) * 1.0f / (float)uprate; /* :129, resize :203-220 */
        const int mid = width / 2;
        taps = 2 * (mid / uprate) + 1;                                                        /* :148 */
        if (taps > max_taps) { for (int q = 0; q < 7; ++q) free(g[q]); free(up); return -1; }
        int* downs = (int*)malloc(sizeof(int) * taps);
        const int nlow = mid / uprate + 1;                                                    /* :149-163 */
        for (int i = 0; i < nlow; ++i) downs[i] = mid - (nlow - 1 - i) * uprate;
        for (int i = nlow, j = mid + uprate; i < taps; ++i, j += uprate) downs[i] = j;
        float* tmp = (float*)malloc(sizeof(float) * width);
        for (int q = 0; q < 7; ++q) {
            sc_conv1d(g[q], width, up, fl, tmp);                                              /* :136-144 */
            for (int i = 0; i < taps; ++i) filters[(size_t)q * taps + i] = tmp[downs[i]];      /* :166-172 */
        }
        free(tmp); free(downs); free(up);
    } else {
        if (taps > max_taps) { for (int q = 0; q < 7; ++q) free(g[q]); return -1; }
        for (int q = 0; q < 7; ++q) memcpy(filters + (size_t)q * taps, g[q], sizeof(float) * taps);
    }
    for (int i = 0; i < taps; ++i) { const float v = filters[2 * (size_t)taps + i]; abs3[i] = v * (v < 0 ? -1 : 1); } /* :174-178 */
    for (int q = 0; q < 7; ++q) free(g[q]);
    return taps;
}

/* ------------------------------------------------------------------ S-CIELAB pipeline (next row 1)
 * The plugin's real cost: S-CIELAB(original) vs S-CIELAB(quantised), CIE76, mean + penalty.
 * Follows the OpenCL kernels (the only implementation the reference has of this stage):
 *   RGB2XYZ cl:79-90, XYZ2Opp cl:111-116, convolve4Channels cl:2-40, convolve1Channel cl:42-74,
 *   quantizeAndConvertToOpp cl:172-199, computeScielabKernelsTemp cl:234-272,
 *   computeScielabKernelsEnd cl:274-306, Opp2LAB cl:124-145, CIEDE cl:201-209,
 *   host sequencing ImageManipulation.java:285-370 (original) and :620-727 (candidates).
 * Device-defined OpenCL builtins are pinned: pow(x,2.4f) and cbrt correctly rounded to fp32, dot()
 * left to right without contraction, fma() exact, distance() = sqrtf(fma chain). */
static const float CL_RGB2XYZ[3][3] = {{0.4124564f, 0.3575761f, 0.1804375f}, {0.2126729f, 0.7151522f, 0.0721750f}, {0.0193339f, 0.1191920f, 0.9503041f}}; /* cl:77 */
static const float CL_XYZ2OPP[3][3] = {{0.2787336f, 0.7218031f, -0.1065520f}, {-0.4487736f, 0.2898056f, -0.0771569f}, {0.0859513f, -0.5899859f, 0.5011089f}}; /* cl:110 */
static const float CL_OPP2XYZ[3][3] = {{0.624045f, -1.87044f, -0.155304f}, {1.36606f, 0.931563f, 0.433903f}, {1.5013f, 1.41761f, 2.53307f}};                  /* cl:118 */
static const float CL_RGB2OPP[3][3] = {{0.266413f, 0.603167f, 0.00113333f}, {-0.124957f, 0.0375879f, -0.133381f}, {-0.0803345f, -0.331467f, 0.449132f}};    /* cl:171 */

static inline float cl_dot3(const float m[3], float x, float y, float z) { return (x * m[0] + y * m[1]) + z * m[2]; }
static inline int reflect(int off, int n) { /* cl:20-27 */
    if (off < 0) off = -off - 1;
    else if (off >= n) off = (n << 1) - off - 1;
    return off;
}

static void cl_opp_to_lab(const float opp[3], const float ill[3], float lab[3]) { /* cl:124-145 */
    volatile float n216 = 216.0f, d24389 = 24389.0f, n27 = 27.0f;
    const float LABDELTA3 = n216 / d24389, kappa = d24389 / n27; /* cl:122-123 */
    float f[3];
    for (int c = 0; c < 3; ++c) {
        const float v = cl_dot3(CL_OPP2XYZ[c], opp[0], opp[1], opp[2]);
        const float t = v / ill[c];
        f[c] = (t > LABDELTA3) ? hqo_cbrt_pow(t) : fmaf(kappa, t, 16.0f) / 116.0f;
    }
    lab[0] = 116.0f * f[1] - 16.0f;
    lab[1] = 500.0f * (f[0] - f[1]);
    lab[2] = 200.0f * (f[1] - f[2]);
}

/* one separable pass of convolve4Channels / convolve1Channel on `nch` interleaved channels: in is
 * [rows][cols][nch], out is TRANSPOSED [cols][rows][nch]; k[t][c]; accumulates (+=) if update */
typedef struct { const float* in; float* out; const float* k; int rows, cols, nch, taps, update; } conv_ctx;
static void conv_rows(void* p, size_t lo, size_t hi, int tid) {
    (void)tid;
    conv_ctx* c = (conv_ctx*)p;
    const int half = c->taps / 2;
    for (size_t i = lo; i < hi; ++i)
        for (int j = 0; j < c->cols; ++j)
            for (int ch = 0; ch < c->nch; ++ch) {
                float acc = 0.0f;
                for (int t = 0; t < c->taps; ++t) {
                    const int off = reflect(j + t - half, c->cols);
                    acc = fmaf(c->in[(i * c->cols + off) * c->nch + ch], c->k[t * c->nch + ch], acc);
                }
                float* o = &c->out[((size_t)j * c->rows + i) * c->nch + ch];
                *o = c->update ? *o + acc : acc;
            }
}
static void conv_pass(const float* in, float* out, const float* k, int rows, int cols, int nch, int taps, int update, int threads) {
    conv_ctx c = {in, out, k, rows, cols, nch, taps, update};
    parallel_ranges((size_t)rows, threads, conv_rows, &c);
}

/* filters [7][taps] as hqo_scielab_filters returns them -> k1[t][3], k2[t][3] (updateOpenCLFilters, ImageManipulation.java:800-841) */
static void pack_filters(const float* filters, int taps, float* k1, float* k2) {
    for (int t = 0; t < taps; ++t) {
        k1[3 * t] = filters[0 * taps + t]; k1[3 * t + 1] = filters[3 * taps + t]; k1[3 * t + 2] = filters[5 * taps + t];
        k2[3 * t] = filters[1 * taps + t]; k2[3 * t + 1] = filters[4 * taps + t]; k2[3 * t + 2] = filters[6 * taps + t];
    }
}

/* S-CIELAB representation of the ORIGINAL image: sRGBToScielab (ScielabProcessor.java:374-381) */
/* lin: linear-light RGB [n][3] */
static void scielab_image_lin(const float* lin, int w, int h, int whitepoint, const float* filters, const float* abs3,
                              int taps, float* lab /*[3][n]*/, int threads) {
    const size_t n = (size_t)w * h;
    const float* ill = WHITE[whitepoint == HQO_WHITE_D50 ? 1 : 0];
    float* opp = (float*)malloc(sizeof(float) * 3 * n);
    float* tmp = (float*)malloc(sizeof(float) * 3 * n);
    float* conv = (float*)malloc(sizeof(float) * 3 * n);
    float* o1 = (float*)malloc(sizeof(float) * n);
    float* t1 = (float*)malloc(sizeof(float) * n);
    float* c1 = (float*)malloc(sizeof(float) * n);
    float* k1 = (float*)malloc(sizeof(float) * 3 * taps);
    float* k2 = (float*)malloc(sizeof(float) * 3 * taps);
    pack_filters(filters, taps, k1, k2);
    for (size_t i = 0; i < n; ++i) { /* RGB2XYZ cl:79-90 then XYZ2Opp cl:111-116 */
        const float R = lin[3 * i], G = lin[3 * i + 1], B = lin[3 * i + 2];
        const float X = cl_dot3(CL_RGB2XYZ[0], R, G, B), Y = cl_dot3(CL_RGB2XYZ[1], R, G, B), Z = cl_dot3(CL_RGB2XYZ[2], R, G, B);
        for (int c = 0; c < 3; ++c) opp[3 * i + c] = cl_dot3(CL_XYZ2OPP[c], X, Y, Z);
        o1[i] = opp[3 * i];
    }
    /* ImageManipulation.java:319-346: (H then V) with k1, += (H then V) with k2, .x += (H k3, V |k3|) */
    conv_pass(opp, tmp, k1, h, w, 3, taps, 0, threads);
    conv_pass(tmp, conv, k1, w, h, 3, taps, 0, threads);
    conv_pass(opp, tmp, k2, h, w, 3, taps, 0, threads);
    conv_pass(tmp, conv, k2, w, h, 3, taps, 1, threads);
    conv_pass(o1, t1, filters + 2 * (size_t)taps, h, w, 1, taps, 0, threads);
    for (size_t i = 0; i < n; ++i) c1[i] = conv[3 * i];
    conv_pass(t1, c1, abs3, w, h, 1, taps, 1, threads);
    for (size_t i = 0; i < n; ++i) {
        float o[3] = {c1[i], conv[3 * i + 1], conv[3 * i + 2]}, l[3];
        cl_opp_to_lab(o, ill, l);
        lab[i] = l[0]; lab[n + i] = l[1]; lab[2 * n + i] = l[2];
    }
    free(opp); free(tmp); free(conv); free(o1); free(t1); free(c1); free(k1); free(k2);
}

void hqo_scielab_image(const uint8_t* rgb, int w, int h, int whitepoint, const float* filters, const float* abs3,
                       int taps, float* lab /*[3][n]*/, int threads) {
    const size_t n = (size_t)w * h;
    float tab[256];
    for (unsigned v = 0; v < 256; ++v) tab[v] = hqo_srgb_decode(hqo_u8_to_unit(v));
    float* lin = (float*)malloc(sizeof(float) * 3 * (n ? n : 1));
    for (size_t i = 0; i < 3 * n; ++i) lin[i] = tab[rgb[i]];
    scielab_image_lin(lin, w, h, whitepoint, filters, abs3, taps, lab, threads);
    free(lin);
}

void hqo_scielab_image_f32(const float* unit3 /*[3][n]*/, int w, int h, int whitepoint, const float* filters, const float* abs3,
                           int taps, float* lab /*[3][n]*/, int threads) {
    const size_t n = (size_t)w * h;
    float* lin = (float*)malloc(sizeof(float) * 3 * (n ? n : 1));
    for (size_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) lin[3 * i + c] = hqo_srgb_decode(unit3[(size_t)c * n + i]);
    scielab_image_lin(lin, w, h, whitepoint, filters, abs3, taps, lab, threads);
    free(lin);
}

typedef struct {
    const uint16_t* idx; const float* opp_tab; /* [K][3] */
    const float *k1, *k2, *k3, *abs3; int w, h, taps;
    float *t1, *t2, *t3; /* transposed [w][h][3], [w][h][3], [w][h] */
    const float* orig; const float* ill; size_t n;
    int64_t* t_err; /* per thread */
} sc_ctx;
static void sc_temp_rows(void* p, size_t lo, size_t hi, int tid) { /* computeScielabKernelsTemp cl:234-272 */
    (void)tid;
    sc_ctx* c = (sc_ctx*)p;
    const int half = c->taps / 2, w = c->w, h = c->h;
    for (size_t i = lo; i < hi; ++i)
        for (int j = 0; j < w; ++j) {
            float a1[3] = {0, 0, 0}, a2[3] = {0, 0, 0}, a3 = 0;
            for (int t = 0; t < c->taps; ++t) {
                const float* in = c->opp_tab + 3 * (size_t)c->idx[i * w + reflect(j + t - half, w)];
                for (int ch = 0; ch < 3; ++ch) { a1[ch] = fmaf(in[ch], c->k1[3 * t + ch], a1[ch]); a2[ch] = fmaf(in[ch], c->k2[3 * t + ch], a2[ch]); }
                a3 = fmaf(in[0], c->k3[t], a3);
            }
            const size_t o = (size_t)j * h + i;
            for (int ch = 0; ch < 3; ++ch) { c->t1[3 * o + ch] = a1[ch]; c->t2[3 * o + ch] = a2[ch]; }
            c->t3[o] = a3;
        }
}
static void sc_end_rows(void* p, size_t lo, size_t hi, int tid) { /* computeScielabKernelsEnd cl:274-306 + Opp2LAB + CIEDE */
    sc_ctx* c = (sc_ctx*)p;
    const int half = c->taps / 2, w = c->w, h = c->h;
    int64_t err = 0;
    for (size_t j = lo; j < hi; ++j)      /* rows of the transposed images = image columns */
        for (int i = 0; i < h; ++i) {
            float out[3] = {0, 0, 0};
            for (int t = 0; t < c->taps; ++t) {
                const size_t o = j * h + reflect(i + t - half, h);
                for (int ch = 0; ch < 3; ++ch) out[ch] = fmaf(c->t1[3 * o + ch], c->k1[3 * t + ch], fmaf(c->t2[3 * o + ch], c->k2[3 * t + ch], out[ch]));
                out[0] = fmaf(c->t3[o], c->abs3[t], out[0]);
            }
            float lab[3];
            cl_opp_to_lab(out, c->ill, lab);
            const size_t px = (size_t)i * w + j;
            err += to_fx(sqrtf(dist2(c->orig[px], c->orig[c->n + px], c->orig[2 * c->n + px], lab[0], lab[1], lab[2]))); /* cl:209 */
        }
    c->t_err[tid] += err;
}

/* candidate costs with the S-CIELAB pipeline.  scielab_orig = hqo_scielab_image(...) [3][n] */
void hqo_scielab_eval(const uint8_t* rgb, int w, int h, int whitepoint, const float* filters, const float* abs3, int taps,
                      const float* scielab_orig, const float* palettes, int B, int K, int space, int64_t* err_fx,
                      uint64_t* counts, int threads) {
    const size_t n = (size_t)w * h;
    float* lab = (float*)malloc(sizeof(float) * 3 * (n ? n : 1));
    float* unit = (float*)malloc(sizeof(float) * 3 * (n ? n : 1));
    hqo_image_planes(rgb, n, whitepoint, unit, unit + n, unit + 2 * n, lab, lab + n, lab + 2 * n, threads);
    hqo_scielab_eval_planes(unit, lab, w, h, whitepoint, filters, abs3, taps, scielab_orig, palettes, B, K, space, err_fx, counts, threads);
    free(lab); free(unit);
}

void hqo_scielab_eval_planes(const float* unit3, const float* lab3, int w, int h, int whitepoint, const float* filters, const float* abs3,
                             int taps, const float* scielab_orig, const float* palettes, int B, int K, int space, int64_t* err_fx,
                             uint64_t* counts, int threads) {
    const size_t n = (size_t)w * h;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    uint16_t* idx = (uint16_t*)malloc(sizeof(uint16_t) * n * B);
    hqo_assign_reduce_planes(unit3, lab3, n, whitepoint, palettes, B, K, space, NULL, counts, NULL, idx, threads);
    sc_ctx c; memset(&c, 0, sizeof c);
    float* k1 = (float*)malloc(sizeof(float) * 3 * taps); float* k2 = (float*)malloc(sizeof(float) * 3 * taps);
    pack_filters(filters, taps, k1, k2);
    float* tab = (float*)malloc(sizeof(float) * 3 * K);
    c.k1 = k1; c.k2 = k2; c.k3 = filters + 2 * (size_t)taps; c.abs3 = abs3; c.w = w; c.h = h; c.taps = taps; c.n = n;
    c.t1 = (float*)malloc(sizeof(float) * 3 * n); c.t2 = (float*)malloc(sizeof(float) * 3 * n); c.t3 = (float*)malloc(sizeof(float) * n);
    c.orig = scielab_orig; c.ill = WHITE[whitepoint == HQO_WHITE_D50 ? 1 : 0]; c.opp_tab = tab;
    c.t_err = (int64_t*)malloc(sizeof(int64_t) * 256);
    for (int b = 0; b < B; ++b) {
        for (int k = 0; k < K; ++k) { /* cl:194-198: decode the chosen palette colour, RGB -> Opp */
            const float* pc = palettes + ((size_t)b * K + k) * 4;
            const float R = hqo_srgb_decode(pc[0]), G = hqo_srgb_decode(pc[1]), Bl = hqo_srgb_decode(pc[2]);
            for (int ch = 0; ch < 3; ++ch) tab[3 * k + ch] = cl_dot3(CL_RGB2OPP[ch], R, G, Bl);
        }
        c.idx = idx + (size_t)b * n;
        memset(c.t_err, 0, sizeof(int64_t) * 256);
        parallel_ranges((size_t)h, threads, sc_temp_rows, &c);
        parallel_ranges((size_t)w, threads, sc_end_rows, &c);
        int64_t e = 0;
        for (int t = 0; t < 256; ++t) e += c.t_err[t];
        err_fx[b] = e;
    }
    free(idx); free(k1); free(k2); free(tab); free(c.t1); free(c.t2); free(c.t3); free(c.t_err);
}

/* error-image mode: ImageManipulation.computeError (:858-894) on two S-CIELAB images */
static double error_image_of_labs(const float* la, const float* lb, size_t n, float* error_map, uint8_t* error_map_u8) {
    int64_t sum = 0;
    for (size_t i = 0; i < n; ++i) {
        const float e = sqrtf(dist2(la[i], la[n + i], la[2 * n + i], lb[i], lb[n + i], lb[2 * n + i])); /* cl:209 */
        const float d = 255.0f - e;
        const float v = (d * d) / 65025.0f; /* :890 ((255-e)*(255-e))/(255*255) */
        if (error_map) error_map[i] = v;
        if (error_map_u8) { float q = v * 255.0f + 0.5f; q = q < 0 ? 0 : (q > 255 ? 255 : q); error_map_u8[i] = (uint8_t)(int)q; }
        sum += to_fx(e);
    }
    return n ? ((double)sum * (1.0 / 16777216.0)) / (double)n : 0.0;
}

double hqo_error_image(const uint8_t* rgb_a, const uint8_t* rgb_b, int w, int h, int whitepoint, const float* filters, const float* abs3,
                       int taps, float* error_map, uint8_t* error_map_u8, int threads) {
    const size_t n = (size_t)w * h;
    float* la = (float*)malloc(sizeof(float) * 3 * n); float* lb = (float*)malloc(sizeof(float) * 3 * n);
    hqo_scielab_image(rgb_a, w, h, whitepoint, filters, abs3, taps, la, threads);
    hqo_scielab_image(rgb_b, w, h, whitepoint, filters, abs3, taps, lb, threads);
    const double mean = error_image_of_labs(la, lb, n, error_map, error_map_u8);
    free(la); free(lb);
    return mean;
}

/* both images as float planes (errorImage converts both sequences to FLOAT, HybridQuantization.java:142-143) */
double hqo_error_image_f32(const float* unit3_a, const float* unit3_b, int w, int h, int whitepoint, const float* filters, const float* abs3,
                           int taps, float* error_map, uint8_t* error_map_u8, int threads) {
    const size_t n = (size_t)w * h;
    float* la = (float*)malloc(sizeof(float) * 3 * n); float* lb = (float*)malloc(sizeof(float) * 3 * n);
    hqo_scielab_image_f32(unit3_a, w, h, whitepoint, filters, abs3, taps, la, threads);
    hqo_scielab_image_f32(unit3_b, w, h, whitepoint, filters, abs3, taps, lb, threads);
    const double mean = error_image_of_labs(la, lb, n, error_map, error_map_u8);
    free(la); free(lb);
    return mean;
}

/* ------------------------------------------------------------------ CIE94 (scope row f4: restated and pinned, not built on the GPU)
 * The CIE94 branch of the reference's CIEDE kernel (OptimizedConvolution.cl:217-226, selected by -DCIE94,
 * ImageManipulation.java:63; the plugin itself always passes CIE76, HybridQuantization.java:96,145).  fp32 arithmetic with exact
 * fma(); the weights 1 + 0.045*C1 and 1 + 0.015*C1 are evaluated in double (the literals are doubles in OpenCL C) and narrowed
 * on assignment; deltaH takes the square root of da^2 + db^2 - dC^2, which rounding makes slightly negative for collinear chroma
 * vectors: the result is then NaN, exactly as the reference kernel's. */
float hqo_delta_e94(const float lab1[3], const float lab2[3]) {
    const float dL = lab1[0] - lab2[0];
    const float c1 = sqrtf(fmaf(lab1[1], lab1[1], lab1[2] * lab1[2]));
    const float dC = c1 - sqrtf(fmaf(lab2[1], lab2[1], lab2[2] * lab2[2]));
    const float da = lab1[1] - lab2[1], db = lab1[2] - lab2[2];
    const float dH = sqrtf(fmaf(da, da, db * db) - dC * dC);
    const float sc = (float)(1 + 0.045 * (double)c1), sh = (float)(1 + 0.015 * (double)c1);
    const float qc = dC / sc, qh = dH / sh;
    return sqrtf(fmaf(dL, dL, fmaf(qc, qc, qh * qh)));
}
void hqo_delta_e94_array(const float* lab1 /*[n][3]*/, const float* lab2, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) out[i] = hqo_delta_e94(lab1 + 3 * i, lab2 + 3 * i);
}

/* ------------------------------------------------------------------ range evaluation (tests) */
typedef struct { int which; uint32_t first; float* out; } mrange_ctx;
static void mrange_fn(void* p, size_t lo, size_t hi, int tid) {
    (void)tid;
    mrange_ctx* c = (mrange_ctx*)p;
    for (size_t i = lo; i < hi; ++i) {
        const uint32_t u = c->first + (uint32_t)i;
        float v; memcpy(&v, &u, 4);
        float r;
        switch (c->which) {
            case 0: r = hqo_cbrt_pow(v); break;
            case 1: r = hqo_pow_2p4(v); break;
            case 2: r = hqo_srgb_decode(v); break;
            case 3: r = v / WHITE[0][0]; break;   /* X / illuminant[0], D65 (ScielabProcessor.java:300) */
            case 4: r = v / WHITE[0][2]; break;   /* Z / illuminant[2], D65 (:304) */
            case 5: r = v / C_3LABDELTA2; break;  /* t / (3*LABDELTA2) (:301) */
            case 6: r = v / WHITE[1][0]; break;
            default: r = v / WHITE[1][2]; break;
        }
        c->out[i] = r;
    }
}
/* which: 0 (float)pow(t,1.0/3.0), 1 (float)pow(b,2.4f), 2 sRGB decode, 3-7 the divisions by constants; over `count`
 * consecutive float bit patterns starting at first_bits */
void hqo_math_range(int which, uint32_t first_bits, uint32_t count, float* out, int threads) {
    mrange_ctx c = {which, first_bits, out};
    parallel_ranges(count, threads, mrange_fn, &c);
}

/* ------------------------------------------------------------------ synthetic images */
static uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* uniform: byte j of the image is byte (j%8) (little-endian) of SplitMix64 output j/8 + 2.
 * smooth: per channel a bilinear blend of four corner bytes (outputs 0 and 1) plus
 * ((uniform byte) % 17) - 8, clamped to 0..255. */
void hqo_synth_image(uint8_t* rgb, int w, int h, uint64_t seed, int smooth) {
    uint64_t s = seed;
    const uint64_t o0 = splitmix64(&s), o1 = splitmix64(&s);
    const size_t nbytes = (size_t)w * h * 3;
    uint64_t cur = 0;
    for (size_t j = 0; j < nbytes; ++j) {
        if ((j & 7) == 0) cur = splitmix64(&s);
        rgb[j] = (uint8_t)(cur >> (8 * (j & 7)));
    }
    if (!smooth) return;
    int corner[3][4];
    for (int c = 0; c < 3; ++c)
        for (int q = 0; q < 4; ++q) {
            const int b = c * 4 + q;
            corner[c][q] = (int)(((b < 8 ? o0 : o1) >> (8 * (b & 7))) & 0xff);
        }
    const int64_t wd = w > 1 ? w - 1 : 1, hd = h > 1 ? h - 1 : 1;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c) {
                const size_t j = ((size_t)y * w + x) * 3 + c;
                const int64_t top = corner[c][0] * (wd - x) + (int64_t)corner[c][1] * x;
                const int64_t bot = corner[c][2] * (wd - x) + (int64_t)corner[c][3] * x;
                const int64_t base = (top * (hd - y) + bot * y) / (wd * hd);
                int64_t v = base + (int)(rgb[j] % 17) - 8;
                rgb[j] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
}
