/* hq_oracle.h — CPU ORACLE of the HybridQuantization hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  Nothing under hybridquantization_b200/ links or calls it.
 *
 * PARITY UNPINNED: the reference (Helios77760/HybridQuantization) ships no tests, golden
 * vectors or fixtures, has no CPU implementation of this path (ImageManipulation.java:397,
 * :773 return zeros without OpenCL) and cannot be built here (no JDK, three un-vendored
 * jars, no OpenCL).  This file is a plain-C restatement of the reference's algorithm; it is
 * pinned only by analytic anchors, java.util.Random known answers (App. B)
 * and the committed fixtures under tests/golden/ that this oracle itself generated.
 *
 * Citations are File:line under /root/reference/src/plugins/dbrasseur/hybridquantization/.
 */
#ifndef HQ_ORACLE_H
#define HQ_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { HQO_WHITE_D65 = 0, HQO_WHITE_D50 = 1 };
/* LAB: assign + score in CIELAB (north_star).  SRGB: assign by sRGB distance as
 * OptimizedConvolution.cl:178-193 does, score by CIELAB distance (cl:209). */
enum { HQO_SPACE_LAB = 0, HQO_SPACE_SRGB = 1 };

/* ---- colour arithmetic (ScielabProcessor.java:279-311, :20-21, :59-61) ---- */
float hqo_u8_to_unit(unsigned c);
float hqo_srgb_decode(float c);                               /* :282 */
void hqo_srgb_to_opp(const float rgb[3], float opp[3]);       /* :279-291 */
void hqo_opp_to_lab(const float opp[3], int whitepoint, float lab[3]); /* :293-311 */
void hqo_srgb_to_lab(const float rgb[3], int whitepoint, float lab[3]); /* :432 */
float hqo_lab_constants(int which); /* 0 LABDELTA3, 1 3*LABDELTA2, 2 4f/29f */
float hqo_pow_2p4(float b);         /* (float)Math.pow(b, 2.4f) */
float hqo_cbrt_pow(float t);        /* (float)Math.pow(t, 1.0/3.0) */

/* packed u8 RGB (3 B/px) -> planar fp32: unit sRGB planes and Lab planes (either may be NULL) */
void hqo_image_planes(const uint8_t* rgb, size_t n, int whitepoint, float* unit_r, float* unit_g,
                      float* unit_b, float* lab_l, float* lab_a, float* lab_b, int threads);

/* ---- assign + reduce (OptimizedConvolution.cl:178-193, :209) ----
 * palettes: [B][K][4] sRGB floats (R,G,B,0) as SWASA.java:42-50 lays them out.
 * Outputs per candidate (any may be NULL): err_fx[B] = sum of llrintf(dE * 2^24);
 * counts[B][K]; sums_fx[B][K][3] = sum of llrintf(Lab * 2^24) of the assigned pixels;
 * idx[B][n] nearest-colour index. */
void hqo_assign_reduce(const uint8_t* rgb, size_t n, int whitepoint, const float* palettes, int B,
                       int K, int space, int64_t* err_fx, uint64_t* counts, int64_t* sums_fx,
                       uint16_t* idx, int threads);
/* same, from precomputed planes (so a timed loop does not redo RGB->Lab) */
void hqo_assign_reduce_planes(const float* unit_rgb3 /*[3][n] or NULL*/,
                              const float* lab3 /*[3][n]*/, size_t n, int whitepoint,
                              const float* palettes, int B, int K, int space, int64_t* err_fx,
                              uint64_t* counts, int64_t* sums_fx, uint16_t* idx, int threads);

/* cost = (sum dE)/N + delta * #unused  (ImageManipulation.java:712,736-752; SWASA.java:74-82) */
double hqo_cost(int64_t err_fx, const uint64_t* counts, int K, uint64_t n_total, float delta);

/* final image (OptimizedConvolution.cl:147-170): out_rgb packed u8 = round(255*c) of the chosen
 * palette colour; out_f32 [n][4] floats exactly as the reference's quantize() returns */
void hqo_quantize(const uint8_t* rgb, size_t n, int whitepoint, const float* palette, int K,
                  int space, uint8_t* out_rgb, float* out_f32, uint16_t* idx, int threads);

/* ---- java.util.Random (JDK spec LCG) ---- */
typedef struct { uint64_t state; } hqo_rng;
void hqo_rng_seed(hqo_rng* r, int64_t seed);
int32_t hqo_rng_next(hqo_rng* r, int bits);
float hqo_rng_next_float(hqo_rng* r);
double hqo_rng_next_double(hqo_rng* r);

/* ---- SWASA (SWASA.java:30-106) + findBestQuantization (ImageManipulation.java:383-591) ---- */
typedef struct {
    int population, imax, iTc;        /* HybridQuantization.java:197,199,214 */
    float delta;                      /* :201 */
    int convergence;                  /* :204 */
    float conv_delay, conv_spread;    /* :206,208 */
    float t0, alpha, s0, beta;        /* :212,216,223,224 */
    int whitepoint, space;
    int64_t seed;
    int cost_model;                   /* 0 Lab / identity filter, 1 full S-CIELAB chain */
    int dpi; float viewing_distance;  /* HybridQuantization.java:229-231 */
} hqo_swasa_params;
void hqo_swasa_defaults(hqo_swasa_params* p);

void hqo_generate_random_colors(hqo_rng* r, int K, float* colors /*[K][4]*/); /* SWASA.java:40 */
float hqo_max_step_width(const hqo_swasa_params* p, int ite);                   /* :69 */
void hqo_generate_neighboring_colors(const hqo_swasa_params* p, hqo_rng* r, const float* colors,
                                     float* next, int K, int ite);            /* :91 */

/* Full search.  best_colors [K][4]; trace_costs (optional) receives (imax+1)*P candidate
 * costs in evaluation order: the initial population, then each iteration's candidates.
 * Returns best error. */
double hqo_find_best_quantization(const uint8_t* rgb, int w, int h, int K,
                                  const hqo_swasa_params* p, float* best_colors,
                                  double* trace_costs, int threads);

/* ---- S-CIELAB (next row 1): filter bank of ScielabProcessor.java:66-181.  filters [7][taps]
 * (O1g1,O1g2,O1g3,O2g1,O2g2,O3g1,O3g2), abs3 [taps]; returns taps or -1 */
int hqo_scielab_filters(int dpi, double viewing_distance, float* filters, float* abs3, int max_taps);
/* S-CIELAB representation of the original image, lab [3][n] (ScielabProcessor.sRGBToScielab :374-381) */
void hqo_scielab_image(const uint8_t* rgb, int w, int h, int whitepoint, const float* filters, const float* abs3,
                       int taps, float* lab, int threads);
/* candidate costs through quantise -> Opp -> separable filters -> Lab -> CIE76 vs scielab_orig
 * (ImageManipulation.java:620-727): err_fx[B], counts[B][K] */
void hqo_scielab_eval(const uint8_t* rgb, int w, int h, int whitepoint, const float* filters, const float* abs3, int taps,
                      const float* scielab_orig, const float* palettes, int B, int K, int space, int64_t* err_fx,
                      uint64_t* counts, int threads);

/* error-image mode (ImageManipulation.computeError :858-894): returns mean dE; maps optional */
double hqo_error_image(const uint8_t* rgb_a, const uint8_t* rgb_b, int w, int h, int whitepoint, const float* filters, const float* abs3,
                       int taps, float* error_map, uint8_t* error_map_u8, int threads);

/* tests: evaluate which (0 cube-root pow, 1 pow 2.4f, 2 sRGB decode) over consecutive float bit patterns */
void hqo_math_range(int which, uint32_t first_bits, uint32_t count, float* out, int threads);

/* synthetic inputs: SplitMix64-seeded uniform / smooth images */
void hqo_synth_image(uint8_t* rgb, int w, int h, uint64_t seed, int smooth);

#ifdef __cplusplus
}
#endif
#endif
