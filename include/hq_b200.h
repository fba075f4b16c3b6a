/* hq_b200.h — C ABI of the B200-native HybridQuantization hot path (libhq_b200.so).
 *
 * This is the drop-in boundary: it replaces the reference's JavaCL/OpenCL backend class
 * `ImageManipulation` (ImageManipulation.java) for the path
 *     RGB -> CIELAB  ->  nearest-palette assignment  ->  per-colour counts / Lab sums /
 *     total error  ->  SWASA candidate scoring,
 * and nothing else.  Plain pointers and sizes only; a JNI / cgo / ctypes stub can bind
 * every entry directly (INTEGRATION.md shows the JNI stub for the reference's plugin).
 *
 * Citations are File:line under
 * /root/reference/src/plugins/dbrasseur/hybridquantization/.
 *
 * Error behaviour: every entry returns HQ_OK (0) or an error code; the text is available
 * from hq_last_error().  There is NO CPU / OpenCL fallback: where the reference silently
 * returns zero arrays without a device (ImageManipulation.java:397,773), this library
 * fails loudly.
 *
 * Threading: a context is single-caller, like the reference's backend object which is only
 * used from EzPlug's execute() thread.  hq_request_stop() alone may be called from another
 * thread (the reference's stopExecution(), HybridQuantization.java:311-315).
 */
#ifndef HQ_B200_H
#define HQ_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
    HQ_OK = 0,
    HQ_ERR_INVALID = 1,     /* bad argument */
    HQ_ERR_CUDA = 2,        /* CUDA runtime / driver error, no device */
    HQ_ERR_NO_IMAGE = 3,    /* hq_set_image_* has not been called */
    HQ_ERR_UNSUPPORTED = 4, /* e.g. K > HQ_MAX_COLORS_ANY, a 16-bit index image for K > 65,535 */
    HQ_ERR_CALLBACK = 5     /* the all-reduce hook reported failure */
};

enum { HQ_WHITEPOINT_D65 = 0, HQ_WHITEPOINT_D50 = 1 }; /* ScielabProcessor.java:19-21 */

/* HQ_SPACE_LAB : assign and score in CIELAB (the accelerated path's default).
 * HQ_SPACE_SRGB: assign by sRGB distance exactly as OptimizedConvolution.cl:178-193 does,
 *                score by CIELAB distance (CIEDE/CIE76, cl:209). */
enum { HQ_SPACE_LAB = 0, HQ_SPACE_SRGB = 1 };

/* what a candidate is scored with: HQ_COST_LAB = sum ||Lab(px) - Lab(P[idx])|| (identity spatial filter,
 * the accelerated path); HQ_COST_SCIELAB = the reference's full S-CIELAB chain (hq_eval_palettes_scielab) */
enum { HQ_COST_LAB = 0, HQ_COST_SCIELAB = 1 };

/* ImageManipulation.deltaETypes (ImageManipulation.java:20; the constructor's first argument, :52,:63).  The plugin itself only
 * ever passes CIE76 (HybridQuantization.java:96,145).  HQ_DELTAE_CIE94 is the reference kernel's CIE94 branch
 * (OptimizedConvolution.cl:217-226) bit for bit, INCLUDING its latent NaN: deltaH takes the square root of a difference that
 * rounding makes slightly negative for about one generic pixel pair in 4,000, so the mean of any real image is NaN.  An integer
 * sum cannot hold a NaN: such pixels are counted, and every entry that returns err_fx then reports HQ_ERR_FX_NAN for that
 * candidate, which hq_cost (and the built-in search) turn into a NaN cost — exactly what the reference's host mean would be.
 * HQ_DELTAE_CIEDE2000: the reference's branch is an empty stub (cl:227-229); hq_set_delta_e refuses it (HQ_ERR_UNSUPPORTED).
 * Affects hq_eval_palettes (scored from index images in a second pass), hq_eval_palettes_scielab (two-kernel path),
 * hq_error_image*, hq_delta_e_images and the search. */
enum { HQ_DELTAE_CIE76 = 0, HQ_DELTAE_CIE94 = 1, HQ_DELTAE_CIEDE2000 = 2 };
#define HQ_ERR_FX_NAN INT64_MIN

enum {
    HQ_EVAL_SUMS = 1,            /* also reduce per-colour Lab sums */
    HQ_EVAL_FORCE_DIRECT = 2,    /* kernel variant selection, for tests / profiling */
    HQ_EVAL_FORCE_CHUNKED = 4,
    HQ_EVAL_FORCE_PREFILTER = 8, /* expanded-form prefilter + exact re-check (default for K > 32) */
    HQ_EVAL_PRUNE = 16,          /* exact assignment with geometric pruning (LAB space; same integers as the exhaustive
                                  * kernel, ~10x less arithmetic at K=256): see hq_set_pruning */
    HQ_EVAL_ALLREDUCE = 32       /* hq_eval_palettes_device only: d_results holds the totals over the context's communicator
                                  * when the call's work completes — folded into the scoring kernel's last CTA over peer
                                  * memory where that path is open (hq_comm_open_peers), else hq_comm_allreduce behind it */
};

#define HQ_MAX_COLORS 1024          /* palette sizes the exhaustive kernel stages in shared memory */
#define HQ_MAX_COLORS_PRUNED 4096   /* palette sizes of the pruned kernel: 1024 < K <= 4096 works wherever that kernel applies
                                     * (LAB-space scoring, the S-CIELAB chain, hq_quantize) and is selected automatically */
#define HQ_MAX_COLORS_ANY (1 << 24) /* the plugin's own range ("Number of colors" in [1, 2^24], HybridQuantization.java:192): beyond the
                                     * staged kernels the palette is swept in chunks against a per-pixel running best (same integers,
                                     * untuned: ~10 instructions per pair).  hq_eval_palettes (both spaces, sums) and hq_quantize's
                                     * image outputs take any K up to here; 16-bit index images (hq_quantize's out_idx, the S-CIELAB
                                     * chain) stop at 65,535 */

typedef struct hq_ctx hq_ctx;

/* ---- lifetime: replaces the constructor / close() (ImageManipulation.java:52-93, :265-269) */
int hq_create(int device, hq_ctx** out);
void hq_destroy(hq_ctx* ctx);
/* ctx may be NULL: then the message of the last failed hq_create() on this thread */
const char* hq_last_error(const hq_ctx* ctx);
int hq_device_info(const hq_ctx* ctx, int* sm_count, int* sm_clock_khz, char* name, int name_len);
int hq_set_delta_e(hq_ctx* ctx, int type); /* HQ_DELTAE_*; default CIE76 */

/* ---- image upload + RGB->CIELAB: replaces the uploads at ImageManipulation.java:451,471-472
 * and RGBtoXYZ/XYZtoScielab (:100, :285) with the identity spatial filter.
 * rgb: packed u8, 3 bytes per pixel, row-major, `rows` rows of `width` pixels: the caller's
 * shard of the image (all of it on one GPU).  Host pointer. */
int hq_set_image_u8(hq_ctx* ctx, const uint8_t* rgb, int width, int rows, int whitepoint);
/* A row shard WITH halo rows (needed by the S-CIELAB stage, whose vertical filter reaches taps/2 rows
 * into the neighbours): rgb holds halo_top + own_rows + halo_bottom rows; the first OWN row is row
 * global_row0 of an image of global_rows rows.  Every reduction (counts, sums, errors), hq_image_pixels,
 * hq_get_lab, hq_quantize and hq_scielab_get_image cover the own rows only; halo pixels are assigned so
 * that the quantised halo feeds the filter.  hq_set_image_u8 == no halo, the whole image. */
int hq_set_image_u8_sharded(hq_ctx* ctx, const uint8_t* rgb, int width, int own_rows, int halo_top, int halo_bottom,
                            int global_row0, int global_rows, int whitepoint);
/* same with the packed RGB already resident in device memory; runs on `stream` (a
 * cudaStream_t, NULL = the context's stream) and does not synchronise */
int hq_set_image_u8_device(hq_ctx* ctx, const void* d_rgb, int width, int rows, int whitepoint,
                           void* stream);
/* (the conversion is ordered before every later entry: the context's own stream, and any other stream passed to
 * hq_eval_palettes_device, waits on an event recorded behind it) */
/* The image as the plugin itself holds it: planar floats in [0,1], one array per channel (`im.getDataXYCAsFloat()` after
 * `IcyBufferedImageUtil.convertToType(..., DataType.FLOAT, true)`, HybridQuantization.java:95-98) — so 16-bit and float
 * Icy images need no detour through u8.  Each value is decoded on the device exactly as ScielabProcessor.java:282-284
 * does (double pow, rounded once to float); a u8-derived float image gives the same bits as hq_set_image_u8.
 * Values outside [0,1] or NaN: HQ_ERR_INVALID.  Host pointers; r, g, b each hold rows*width floats. */
int hq_set_image_f32_planar(hq_ctx* ctx, const float* r, const float* g, const float* b, int width, int rows, int whitepoint);
/* row shard with halo rows, arguments as hq_set_image_u8_sharded */
int hq_set_image_f32_planar_sharded(hq_ctx* ctx, const float* r, const float* g, const float* b, int width, int own_rows,
                                    int halo_top, int halo_bottom, int global_row0, int global_rows, int whitepoint);
/* debug / parity: the Lab planes [3][n] (L plane, a plane, b plane) */
int hq_get_lab(hq_ctx* ctx, float* planes);
uint64_t hq_image_pixels(const hq_ctx* ctx);

/* ---- candidate evaluation: replaces computeQuantizationErrorPopulation
 * (ImageManipulation.java:620-727): kernels quantizeAndConvertToOpp + CIEDE, the
 * used-colour flags and the host-side averageArray.
 * palettes: [B][K][4] floats, sRGB in [0,1] laid out R,G,B,0 as SWASA.java:42-50; a NaN or a value outside [0,1]
 * (nothing SWASA.java:93-106 can produce) is refused with HQ_ERR_INVALID by every host-pointer entry.
 * Outputs (each may be NULL), all exact integers so that shards add up bit-identically:
 *   err_fx [B]        sum over pixels of round(dE * 2^24)
 *   counts [B][K]     pixels assigned to each colour (colour used <=> count > 0)
 *   sums_fx[B][K][3]  sum of round(L|a|b * 2^24) of the pixels of each colour
 *                     (requires HQ_EVAL_SUMS in flags)
 * If an all-reduce hook is installed the outputs are the totals over all ranks. */
int hq_eval_palettes(hq_ctx* ctx, const float* palettes, int B, int K, int space, int flags,
                     int64_t* err_fx, uint64_t* counts, int64_t* sums_fx);

/* number of 8-byte words per candidate in the device result buffer:
 * [0] err_fx, [1..K] counts, then with HQ_EVAL_SUMS [1+K .. 1+4K) sums */
int hq_result_words(int K, int flags);
/* fully asynchronous variant on device memory: d_palettes [B][K][4] floats, d_results
 * [B][hq_result_words] words (zeroed by the call).  Runs on `stream` (cudaStream_t; NULL =
 * the context's stream), no host synchronisation, no all-reduce unless HQ_EVAL_ALLREDUCE asks for it.  (With HQ_EVAL_PRUNE the first call after an image
 * change builds the cell-sorted copy of the image and synchronises `stream` once to read the chunk count.) */
int hq_eval_palettes_device(hq_ctx* ctx, const void* d_palettes, int B, int K, int space,
                            int flags, void* d_results, void* stream);

/* ---- exact pruning (added; no reference counterpart: the reference sweeps all K colours for every pixel, cl:178-193).
 * Every output of the scoring step is a sum over pixels, so the pixel order is free: once per image the own pixels are
 * sorted by a coarse CIELAB cell into chunks with exact bounding boxes, and per (chunk, candidate) only the colours that
 * can be the nearest one of some pixel of the chunk are compared.  Outputs are bit-identical to the exhaustive kernel.
 * HQ_EVAL_PRUNE selects it per call (LAB space only; ignored otherwise).  hq_set_pruning chooses what the annealing search
 * (hq_find_best_quantization) does: HQ_PRUNE_OFF never, HQ_PRUNE_AUTO (default) when it pays — K >= 32 on a shard of >= 65,536
 * pixels, K >= 12 on one of >= 786,432 — and the search runs in LAB space with the LAB cost model, HQ_PRUNE_ON whenever the space and cost model allow.
 * The same mode governs the index-producing pruned kernel inside hq_eval_palettes_scielab (AUTO: K >= 32 and >= 65,536 pixels)
 * and hq_quantize (AUTO: only for K > HQ_MAX_COLORS); palettes with HQ_MAX_COLORS < K <= HQ_MAX_COLORS_PRUNED always use it.
 * hq_pruning_stats: chunks of the resident image and, when profiling is enabled, the mean number of colours that
 * survived per (chunk, candidate) since the image was set. */
/* CUDA-graph replay of hq_eval_palettes (added): from the third call with identical arguments on, the launch set of one
 * evaluation (H2D palettes, palette kernel, scoring kernel, D2H totals) is replayed as one graph launch.  Measured on a
 * 512x512, K=16, population-4 search: 29.7 -> 26.0 us per iteration.  OFF by default: the first graph instantiation of a
 * process costs the driver 40-60 ms, which only a host that runs several long searches on small images earns back.
 * Ignored while an all-reduce hook or profiling is active.  The environment variable HQ_CUDA_GRAPHS=1 turns it on for
 * every new context. */
int hq_set_graphs(hq_ctx* ctx, int enabled);

enum { HQ_PRUNE_OFF = 0, HQ_PRUNE_AUTO = 1, HQ_PRUNE_ON = 2 };
int hq_set_pruning(hq_ctx* ctx, int mode);
/* the HQ_EVAL_* flags a search loop over hq_eval_palettes should pass for the resident image under the context's pruning
 * mode (what hq_find_best_quantization and the C++ / Java hosts use): HQ_EVAL_PRUNE or 0 */
int hq_search_eval_flags(const hq_ctx* ctx, int K, int space, int cost_model);
int hq_pruning_stats(hq_ctx* ctx, uint32_t* chunks, double* mean_survivors);

/* cost = (sum dE)/N + delta * #{unused colours}: averageArray + computePenalty
 * (ImageManipulation.java:712,736-752; SWASA.java:74-82).  Pure host arithmetic. */
double hq_cost(int64_t err_fx, const uint64_t* counts, int K, uint64_t n_total, float delta);

/* ---- final image: replaces quantize() (ImageManipulation.java:770-798, kernel cl:147-170).
 * Any output may be NULL.  out_rgb packed u8 [n][3]; out_f32 [n][4] floats (what the
 * reference returns); out_idx [n] palette indices. */
int hq_quantize(hq_ctx* ctx, const float* palette, int K, int space, uint8_t* out_rgb,
                float* out_f32, uint16_t* out_idx);

/* ---- S-CIELAB spatial-filter stage (scope row "next 1"): the plugin's real cost is the mean
 * CIE76 between S-CIELAB(original) and S-CIELAB(quantised image).
 * hq_scielab_configure builds the separable filter bank of ScielabProcessor.java:66-181 from the
 * plugin's "Dpi" / "Viewing distance" parameters (HybridQuantization.java:229-231; defaults 72, 45 cm
 * are applied if it is never called); hq_scielab_set_filters installs a caller-built bank instead
 * (filters7 = [7][taps]: O1g1,O1g2,O1g3,O2g1,O2g2,O3g1,O3g2; abs3 = |O1g3|; taps odd).
 * hq_scielab_get_image returns sRGBToScielab(original) (ScielabProcessor.java:374-381) as planes
 * [3][n].  hq_eval_palettes_scielab replaces computeQuantizationErrorPopulation with its full kernel
 * chain (ImageManipulation.java:635-699): err_fx[B] = sum of round(dE * 2^24), counts[B][K].
 * Row shards need taps/2 halo rows (hq_set_image_u8_sharded); the all-reduce hook then sums errors and counts. */
int hq_scielab_configure(hq_ctx* ctx, int dpi, float viewing_distance_cm);
int hq_scielab_set_filters(hq_ctx* ctx, const float* filters7, const float* abs3, int taps);
/* *taps: in = capacity of the arrays (entries per filter), out = actual taps */
int hq_scielab_get_filters(const hq_ctx* ctx, float* filters7, float* abs3, int* taps);
int hq_scielab_get_image(hq_ctx* ctx, float* planes);
/* error-image mode (scope row "next 2"): HybridQuantization.errorImage (HybridQuantization.java:139-182)
 * + ImageManipulation.computeError (ImageManipulation.java:858-894).  quantized_rgb: packed u8 image of
 * the resident image's size.  error_map [n] = ((255 - dE)^2)/(255*255) (:890), error_map_u8 [n] its
 * 8-bit rendering, *mean_de = mean dE between the two S-CIELAB images.  Outputs may be NULL. */
int hq_error_image(hq_ctx* ctx, const uint8_t* quantized_rgb, float* error_map, uint8_t* error_map_u8, double* mean_de);
/* the second image as float planes in [0,1] (errorImage converts both sequences to FLOAT, HybridQuantization.java:142-143) */
int hq_error_image_f32_planar(hq_ctx* ctx, const float* r, const float* g, const float* b, float* error_map, uint8_t* error_map_u8,
                              double* mean_de);
/* ---- the reference class's remaining one-shot entries, on its own interleaved layouts ([n][4] floats, 4th lane 0), so that a
 * Java drop-in can keep the reference's method signatures (java/plugins/.../CudaImageManipulation.java):
 *   hq_rgb_to_xyz         RGBtoXYZ (ImageManipulation.java:100-152, kernel RGB2XYZ cl:79-90): planar sRGB in [0,1] -> XYZ
 *   hq_xyz_to_scielab     XYZtoScielab (:285-370): XYZ2Opp, the three separable filter pairs, Opp2LAB with the caller's
 *                         illuminant[3]; uses the context's filter bank (hq_scielab_configure / _set_filters; defaults 72 dpi, 45 cm)
 *   hq_scielab_set_image  installs the caller's S-CIELAB image of the resident image (findBestQuantization's
 *                         inlineScielabOriginal argument, :383) as the target candidates are compared with, instead of the
 *                         one hq_eval_palettes_scielab would compute itself (whole image; also on multi-device contexts)
 *   hq_delta_e_images     computeError (:858-894): CIEDE/CIE76 between two Lab images, the error image value
 *                         ((255 - e)^2) / (255 * 255) in lanes 0..2 of error_rgba4 (may be NULL), *mean_de = sum in double / n */
int hq_rgb_to_xyz(hq_ctx* ctx, const float* r, const float* g, const float* b, size_t n, float* xyz4);
int hq_xyz_to_scielab(hq_ctx* ctx, const float* xyz4, int width, int rows, const float* illuminant3, float* lab4);
int hq_scielab_set_image(hq_ctx* ctx, const float* lab4);
int hq_delta_e_images(hq_ctx* ctx, const float* lab4_a, const float* lab4_b, size_t n, float* error_rgba4, double* mean_de);
/* test hook: 1 = always run the generic any-tap-count kernels instead of the 21-tap specialisations; 2 = the 21-tap
 * candidate stage as two kernels per candidate with an intermediate in HBM (round 1) instead of the fused kernel; 0 = default */
int hq_scielab_force_generic(hq_ctx* ctx, int enabled);
/* host only (no GPU needed): the filter bank for (dpi, viewing distance); *taps as above */
int hq_scielab_build_filters(int dpi, float viewing_distance_cm, float* filters7, float* abs3, int* taps);
int hq_eval_palettes_scielab(hq_ctx* ctx, const float* palettes, int B, int K, int space,
                             int64_t* err_fx, uint64_t* counts);

/* ---- multi-GPU: pixel rows are sharded across ranks (one context per GPU); the only
 * exchange is a sum of the integer result words.  The hook is called on the context's
 * stream order with the DEVICE buffer; it must leave the element-wise int64 sum over all
 * ranks in place (e.g. ncclAllReduce(ncclInt64, ncclSum) / torch.distributed.all_reduce)
 * and return 0. */
typedef int (*hq_allreduce_fn)(void* user, void* d_words, size_t n_words, void* stream);
int hq_set_allreduce(hq_ctx* ctx, hq_allreduce_fn fn, void* user);

/* ---- the exchange step inside the library: native NCCL (added in round 2).
 * (a) One process per GPU (torchrun, MPI, ...): rank 0 calls hq_comm_get_unique_id and ships the HQ_COMM_ID_BYTES to the
 *     other ranks by any means; every rank then calls hq_comm_init_rank on its context (collective).  From then on
 *     hq_eval_palettes, hq_eval_palettes_scielab, hq_error_image* and the search return totals over all ranks through
 *     ncclAllReduce(ncclInt64, ncclSum) on the context's stream; hq_comm_allreduce does the same for the buffers of the
 *     device-pointer API.  A hook installed with hq_set_allreduce takes precedence.
 * (b) One process, several GPUs — what replaces JavaCL.createBestContext() + one queue (ImageManipulation.java:58-59) for
 *     a JVM host: hq_create_multi returns ONE context over the listed devices.  hq_set_image_u8 / hq_set_image_f32_planar
 *     take the WHOLE image and split its rows over the devices (rows i*H/G .. (i+1)*H/G on device i, plus the halo rows
 *     the S-CIELAB stage needs: 10, or taps/2 of the filter bank configured at that time); every evaluation runs on all
 *     devices and one grouped ncclAllReduce follows; hq_quantize, hq_get_lab, hq_scielab_get_image and hq_error_image*
 *     gather whole-image outputs.  Entries that take device pointers or explicit shards return HQ_ERR_UNSUPPORTED on it.
 * The sums are exact integers: totals and the annealing trajectory are identical for any number of GPUs.
 * NCCL is loaded at run time (libnccl.so.2 of the process, else of the system; HQ_NCCL_LIB overrides); without it these
 * entries fail with HQ_ERR_UNSUPPORTED. */
#define HQ_COMM_ID_BYTES 128
int hq_comm_get_unique_id(void* id128);
int hq_comm_init_rank(hq_ctx* ctx, const void* id128, int nranks, int rank);
int hq_comm_allreduce(hq_ctx* ctx, void* d_words, size_t n_words, void* stream);
/* rank / size of the context's communicator (0 / 1 without one; size = devices of a multi-device context) and the NCCL
 * version in use (0 = not loaded); each pointer may be NULL */
int hq_comm_info(const hq_ctx* ctx, int* rank, int* size, int* nccl_version);
int hq_create_multi(const int* devices, int ndev, hq_ctx** out);
int hq_multi_device_count(const hq_ctx* ctx);
/* ---- the same exchange over NVLink / NVSwitch PEER MEMORY (no reference counterpart: the reference owns one device and one
 * queue, ImageManipulation.java:58-59) for small payloads (<= 4,096 result words: every search with
 * B x (K + 1) <= 4,096): the CTA that finishes an evaluation stores its words into a mailbox in every rank's HBM, signals,
 * waits for the other ranks' signals and adds the slots up — inside the scoring kernel itself for K <= 32, else in a one-CTA
 * launch behind it; no collective launch, identical integers.  Larger payloads stay on ncclAllReduce.
 * A multi-device context (hq_create_multi) sets it up itself (cudaDeviceEnablePeerAccess).  One process per GPU: after
 * hq_comm_init_rank every rank calls hq_comm_peer_handle, the host ships the HQ_PEER_HANDLE_BYTES of every rank to every
 * rank (rank order), and every rank calls hq_comm_open_peers (cudaIpcOpenMemHandle).  COLLECTIVE: if it fails on any rank
 * (HQ_ERR_UNSUPPORTED: no IPC between the processes, HQ_PEER_EXCHANGE=0), call hq_comm_close_peers on every rank and the
 * exchange stays on NCCL.  A rank that never arrives turns into HQ_ERR_CUDA after HQ_PEER_TIMEOUT_MS (60 s), not a hang. */
#define HQ_PEER_HANDLE_BYTES 64
int hq_comm_peer_handle(hq_ctx* ctx, void* handle64);
int hq_comm_open_peers(hq_ctx* ctx, const void* handles /* [nranks][HQ_PEER_HANDLE_BYTES] */, int nranks, int rank);
void hq_comm_close_peers(hq_ctx* ctx);
int hq_comm_peers_open(const hq_ctx* ctx);   /* 1: small exchanges go over peer memory */

/* ---- the annealing search: replaces findBestQuantization (ImageManipulation.java:383-591)
 * with SWASA.java's schedule.  Accept/reject logic and RNG stay on the host; only the
 * population scoring runs on the GPU (one launch per iteration, batch = population). */
typedef struct {
    int population;     /* "Population size"            HybridQuantization.java:197 (4) */
    int imax;           /* "Max iterations"             :199 (5000) */
    int iTc;            /* "Iterations per temperature" :214 (20) */
    float delta;        /* "Penalty Constant"           :201 (2) */
    int convergence;    /* "Pop Convergence"            :204 (true) */
    float conv_delay;   /* "Convergence delay"          :206 (0.75) */
    float conv_spread;  /* "Convergence spread"         :208 (0.15) */
    float t0;           /* "Initial temperature"        :212 (20) */
    float alpha;        /* "Cooling coefficient"        :216 (0.9) */
    float s0;           /* "Initial Step size"          :223 (100) */
    float beta;         /* "Adaptation constant"        :224 (5.3) */
    int space;          /* HQ_SPACE_* (added; default LAB) */
    int64_t seed;       /* java.util.Random seed (added: the reference's RNG is unseeded) */
    int cost_model;     /* HQ_COST_* (added; default HQ_COST_LAB) */
} hq_swasa_params;
void hq_swasa_default_params(hq_swasa_params* p);

/* n_total: pixels of the WHOLE image (all ranks); 0 = this context's pixel count.
 * best_colors [K][4]; trace_costs (optional) receives (imax+1)*population costs in
 * evaluation order; iterations_done (optional) < imax if hq_request_stop() intervened. */
int hq_find_best_quantization(hq_ctx* ctx, int K, const hq_swasa_params* p, uint64_t n_total,
                              float* best_colors, double* best_error, double* trace_costs,
                              int* iterations_done);
void hq_request_stop(hq_ctx* ctx); /* EzStoppable.stopExecution, HybridQuantization.java:311 */
/* progress hook: called from hq_find_best_quantization every 10 iterations with (iteration, max
 * iterations, best error so far), where the reference updates Icy's progress bar
 * (ImageManipulation.java:546-551, HybridQuantization.updateProgressBar :265-270).  NULL removes it. */
typedef void (*hq_progress_fn)(void* user, int iteration, int max_iterations, double best_error);
int hq_set_progress(hq_ctx* ctx, hq_progress_fn fn, void* user);

/* ---- host-side pieces of the plugin that callers outside C++ may want ---- */
/* java.util.Random-compatible generator */
typedef struct { uint64_t state; } hq_java_random;
void hq_java_random_seed(hq_java_random* r, int64_t seed);
int32_t hq_java_random_next(hq_java_random* r, int bits);
float hq_java_random_next_float(hq_java_random* r);
double hq_java_random_next_double(hq_java_random* r);
/* SWASA.java:40-52, :91-101, :69-72 */
void hq_swasa_generate_random_colors(hq_java_random* r, int K, float* colors);
void hq_swasa_generate_neighboring_colors(const hq_swasa_params* p, hq_java_random* r,
                                          const float* colors, float* next_colors, int K,
                                          int iteration);
float hq_swasa_max_step_width(const hq_swasa_params* p, int iteration);

/* ---- measurement hooks (bench.py): CUDA events recorded on the launching stream right
 * around the assign+reduce kernel of the most recent evaluation; and an FFMA-saturating
 * microbenchmark giving the FP32 CUDA-core ceiling of THIS device at its current clocks
 * (scalar FFMA and packed FFMA2), the denominator of the large-K roofline. */
int hq_set_profiling(hq_ctx* ctx, int enabled);
int hq_last_assign_ms(hq_ctx* ctx, float* ms);
int hq_last_rgb_to_lab_ms(hq_ctx* ctx, float* ms);
/* the S-CIELAB filter-stage launch(es) of the last sub-batch of the most recent hq_eval_palettes_scielab, and how many
 * candidates that sub-batch held (row f1: cl:234-306 + Opp2LAB + CIEDE, without the assignment) */
int hq_last_scielab_stage_ms(hq_ctx* ctx, float* ms, int* candidates);
int hq_measure_fp32_peak(hq_ctx* ctx, double* tflops_ffma, double* tflops_ffma2);

/* ---- test hooks: the single-source arithmetic of csrc/hq_math.h evaluated on the host
 * (which: 0 cube root, 1 pow 2.4f, 2 sRGB decode) over `count` consecutive float bit
 * patterns starting at first_bits, and the same on the device. */
int hq_host_math_range(int which, uint32_t first_bits, uint32_t count, float* out, int threads);
int hq_device_math_range(hq_ctx* ctx, int which, uint32_t first_bits, uint32_t count, float* out);
void hq_host_srgb_to_lab(const float rgb[3], int whitepoint, float lab[3]);

#ifdef __cplusplus
}
#endif
#endif /* HQ_B200_H */
