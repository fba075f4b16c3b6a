// hq_api.cu — the C ABI (include/hq_b200.h): context, device memory, launch orchestration.
// No CPU fallback: every compute entry fails with HQ_ERR_CUDA if the device is unusable.
#include "hq_ctx.h"

namespace hqi { thread_local std::string g_create_error; }
using namespace hqi;

namespace {

// sRGB-assign mode needs the unit planes; they are produced on first use from the resident RGB
int ensure_unit(hq_ctx* c, cudaStream_t st) {
    if (c->have_unit) return HQ_OK;
    HQ_CUDA(c, c->d_unit.reserve(3 * c->stride));
    HQ_CUDA(c, hq::launch_rgb_to_lab(c->d_rgb.p, c->n, c->stride, c->whitepoint, c->d_table.p, c->d_lab.p, c->d_unit.p, c->sm_count, st));
    c->have_unit = true;
    return HQ_OK;
}

int convert_image(hq_ctx* c, int width, int own_rows, int halo_top, int halo_bottom, int g_row0, int g_rows, int whitepoint, cudaStream_t st) {
    c->width = width; c->rows = halo_top + own_rows + halo_bottom; c->whitepoint = whitepoint;
    c->halo_top = halo_top; c->halo_bottom = halo_bottom; c->own_rows = own_rows; c->g_row0 = g_row0; c->g_rows = g_rows;
    c->own_lo = (size_t)halo_top * width; c->own_hi = (size_t)(halo_top + own_rows) * width;
    c->have_unit = c->image_f32;  // a float image IS its unit planes
    c->sc_image_ready = false;
    c->pr_own.ready = false;
    c->pr_all.ready = false;
    ++c->image_gen;
    HQ_CUDA(c, c->d_lab.reserve(3 * c->stride > 0 ? 3 * c->stride : 1));
    if (c->profiling) HQ_CUDA(c, cudaEventRecord(c->ev2, st));
    if (c->image_f32) HQ_CUDA(c, hq::launch_unit_to_lab(c->d_unit.p, c->n, c->stride, whitepoint, c->d_lab.p, c->d_flag.p, c->sm_count, st));
    else HQ_CUDA(c, hq::launch_rgb_to_lab(c->d_rgb.p, c->n, c->stride, whitepoint, c->d_table.p, c->d_lab.p, nullptr, c->sm_count, st));
    if (c->profiling) { HQ_CUDA(c, cudaEventRecord(c->ev3, st)); c->ev_rl_valid = true; }
    c->have_image = true;
    return HQ_OK;
}

int check_eval_args(hq_ctx* c, int B, int K, int space) {
    if (!c) return HQ_ERR_INVALID;
    if (!c->have_image) return fail(c, HQ_ERR_NO_IMAGE, "no image: call hq_set_image_u8 or hq_set_image_f32_planar first");
    if (B < 1 || K < 1) return fail(c, HQ_ERR_INVALID, "B and K must be >= 1 (got B=%d K=%d)", B, K);
    if (K > HQ_MAX_COLORS_ANY) return fail(c, HQ_ERR_UNSUPPORTED, "K=%d exceeds the plugin's own range [1, 2^24] (HybridQuantization.java:192)", K);
    if (space != HQ_SPACE_LAB && space != HQ_SPACE_SRGB) return fail(c, HQ_ERR_INVALID, "unknown space %d", space);
    return HQ_OK;
}

// cell-sorted copy of pixels [lo, hi) of the feature planes of `space` + chunk table + boxes for the pruned kernel
// (one host synchronisation, once per image and set)
int ensure_pruned(hq_ctx* c, hq_ctx::PrunedSet& ps, int space, size_t lo, size_t hi, bool want_perm, cudaStream_t st) {
    if (ps.ready && ps.space == space) return HQ_OK;
    ps.ready = false;
    const size_t n = hi - lo;
    if (n >= 0xffffffffull)  // 32-bit pixel positions in the sort permutation and the chunk table
        return fail(c, HQ_ERR_UNSUPPORTED, "the pruned kernel handles fewer than 2^32 pixels per context (got %zu): use the exhaustive kernel or shard the image", n);
    ps.sstride = hq::plane_stride(n);
    const size_t words = hq::pruned_scratch_words();
    HQ_CUDA(c, ps.sorted.reserve(3 * ps.sstride > 0 ? 3 * ps.sstride : 1));
    if (want_perm) HQ_CUDA(c, ps.perm.reserve(n ? n : 1));
    HQ_CUDA(c, c->d_pr_scratch.reserve(words));
    if (!c->d_pr_stats.p) { HQ_CUDA(c, c->d_pr_stats.reserve(2)); }
    HQ_CUDA(c, cudaMemsetAsync(c->d_pr_stats.p, 0, 16, st));
    const float* feat = space == HQ_SPACE_SRGB ? c->d_unit.p : c->d_lab.p;
    const int cell_bits = hq::pruned_cell_bits(n, space);
    HQ_CUDA(c, hq::launch_pruned_build_cells(feat, c->stride, space, cell_bits, lo, hi, c->d_pr_scratch.p, ps.sorted.p, ps.sstride,
                                             want_perm ? ps.perm.p : nullptr, c->sm_count, st));
    unsigned totals[2] = {0, 0};
    HQ_CUDA(c, cudaMemcpyAsync(totals, c->d_pr_scratch.p + words - 2, sizeof totals, cudaMemcpyDeviceToHost, st));
    HQ_CUDA(c, cudaStreamSynchronize(st));
    if (totals[0] != n) return fail(c, HQ_ERR_CUDA, "pruning: cell sort covered %u of %zu pixels", totals[0], n);
    ps.nchunks = totals[1];
    HQ_CUDA(c, ps.chunk_start.reserve(ps.nchunks ? ps.nchunks : 1));
    HQ_CUDA(c, ps.chunk_len.reserve(ps.nchunks ? ps.nchunks : 1));
    HQ_CUDA(c, ps.box.reserve(ps.nchunks ? 6 * (size_t)ps.nchunks : 1));
    HQ_CUDA(c, hq::launch_pruned_build_chunks(c->d_pr_scratch.p, cell_bits, ps.sorted.p, ps.sstride, ps.nchunks, ps.chunk_start.p, ps.chunk_len.p, ps.box.p, st));
    ps.ready = true; ps.space = space;
    return HQ_OK;
}

// which evaluations go through the pruned kernel: asked for (HQ_EVAL_PRUNE) or forced by K > HQ_MAX_COLORS, and possible:
// scoring (no index image) needs the features to BE CIELAB (the error is the CIELAB distance); index-producing
// evaluations work in either space
// palettes no tuned kernel stages: beyond the pruned kernel's 4,096 colours, or beyond the exhaustive kernel's 1,024 where the
// pruned one cannot score (sRGB-space assignment without an index image) -> the chunked sweep of hq_bigk.cu
bool use_bigk(int K, int space, bool want_idx) {
    return K > HQ_MAX_COLORS_PRUNED || (K > HQ_MAX_COLORS && !(want_idx || space == HQ_SPACE_LAB));
}
bool use_pruned(int K, int space, int flags, bool want_idx) {
    if (use_bigk(K, space, want_idx)) return false;
    const bool wanted = (flags & HQ_EVAL_PRUNE) != 0 || K > HQ_MAX_COLORS;
    return wanted && (want_idx || space == HQ_SPACE_LAB);
}
int prepare_pruned(hq_ctx* c, int K, int space, int flags, bool want_idx, cudaStream_t st) {
    if (!use_pruned(K, space, flags, want_idx)) return HQ_OK;
    if (space == HQ_SPACE_SRGB) { int rc = ensure_unit(c, st); if (rc) return rc; }
    return want_idx ? ensure_pruned(c, c->pr_all, space, 0, c->n, true, st) : ensure_pruned(c, c->pr_own, HQ_SPACE_LAB, c->own_lo, c->own_hi, false, st);
}

// tail (optional): have the scoring kernel itself export the results to pinned host memory; *tail_used reports whether a
// kernel carrying it was actually launched (nothing is launched for an empty image)
int eval_device(hq_ctx* c, const float* d_palettes, int B, int K, int space, int flags, unsigned long long* d_results,
                void* d_idx, cudaStream_t st, const hq::ExportTail* tail = nullptr, bool* tail_used = nullptr) {
    const int K8 = hq::padded_colors(K);
    const bool sums = (flags & HQ_EVAL_SUMS) != 0;
    const int words = hq::result_words(K, sums);
    HQ_CUDA(c, c->d_pal_lab.reserve((size_t)B * K8));
    HQ_CUDA(c, c->d_pal_rgb.reserve((size_t)B * K8));
    if (space == HQ_SPACE_SRGB) { int rc = ensure_unit(c, st); if (rc) return rc; }
    HQ_CUDA(c, hq::launch_palette_features(d_palettes, B, K, c->whitepoint, c->d_pal_lab.p, c->d_pal_rgb.p, st, d_results, (size_t)B * words));
    hq::AssignArgs a;
    a.lab = c->d_lab.p; a.unit = c->d_unit.p; a.n = c->n; a.stride = c->stride;
    a.pal_lab = c->d_pal_lab.p; a.pal_rgb = c->d_pal_rgb.p;
    a.B = B; a.K = K; a.space = space; a.want_sums = sums;
    a.results = d_results; a.idx_out = d_idx; a.sm_count = c->sm_count;
    a.own_lo = c->own_lo; a.own_hi = c->own_hi;
    a.variant = (flags & HQ_EVAL_FORCE_DIRECT) ? 1 : ((flags & HQ_EVAL_FORCE_CHUNKED) ? 2 : ((flags & HQ_EVAL_FORCE_PREFILTER) ? 3 : 0));
    const bool prune = use_pruned(K, space, flags, d_idx != nullptr);
    { int rc = prepare_pruned(c, K, space, flags, d_idx != nullptr, st); if (rc) return rc; }
    if (c->profiling) HQ_CUDA(c, cudaEventRecord(c->ev0, st));
    if (use_bigk(K, space, d_idx != nullptr)) {
        if (d_idx && K > 65535) return fail(c, HQ_ERR_UNSUPPORTED, "index images hold 16 bits: K=%d > 65,535 is available for scoring and for the output image only", K);
        HQ_CUDA(c, c->d_big_d2.reserve(c->n ? c->n : 1));
        HQ_CUDA(c, c->d_big_idx.reserve(c->n ? c->n : 1));
        const float* feat = space == HQ_SPACE_SRGB ? c->d_unit.p : c->d_lab.p;
        for (int b = 0; b < B; ++b)
            HQ_CUDA(c, hq::launch_bigk_candidate(feat, c->d_lab.p, c->n, c->stride, c->own_lo, c->own_hi,
                                                 (space == HQ_SPACE_SRGB ? c->d_pal_rgb.p : c->d_pal_lab.p) + (size_t)b * K8, c->d_pal_lab.p + (size_t)b * K8, K,
                                                 space == HQ_SPACE_SRGB, sums, c->d_big_d2.p, c->d_big_idx.p, d_results + (size_t)b * words,
                                                 d_idx ? static_cast<uint16_t*>(d_idx) + (size_t)b * c->stride : nullptr, c->sm_count, st));
    } else if (prune) {
        const hq_ctx::PrunedSet& ps = d_idx ? c->pr_all : c->pr_own;
        hq::PrunedArgs pa;
        pa.sorted = ps.sorted.p; pa.sstride = ps.sstride; pa.chunk_start = ps.chunk_start.p; pa.chunk_len = ps.chunk_len.p;
        pa.box = ps.box.p; pa.nchunks = ps.nchunks; pa.pal = (d_idx && space == HQ_SPACE_SRGB) ? c->d_pal_rgb.p : c->d_pal_lab.p;
        pa.B = B; pa.K = K; pa.want_sums = sums && !d_idx;
        pa.results = d_results; pa.stats = c->profiling ? c->d_pr_stats.p : nullptr; pa.sm_count = c->sm_count;
        if (d_idx) { pa.perm = ps.perm.p; pa.idx_out = d_idx; pa.istride = c->stride; pa.own_lo = c->own_lo; pa.own_hi = c->own_hi; }
        HQ_CUDA(c, hq::launch_pruned_assign(pa, st));
    } else {
        const bool direct_variant = a.variant == 1 || (a.variant == 0 && K <= hq::kDirectMaxColors);  // the kernel that carries the tail
        if (tail && c->n > 0 && direct_variant) { a.tail = *tail; if (tail_used) *tail_used = true; }
        HQ_CUDA(c, hq::launch_assign_reduce(a, st));
    }
    if (c->profiling) { HQ_CUDA(c, cudaEventRecord(c->ev1, st)); c->ev_valid = true; }
    return HQ_OK;
}

}  // namespace

extern "C" {

int hq_create(int device, hq_ctx** out) {
    if (!out) return fail(nullptr, HQ_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, HQ_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return fail(nullptr, HQ_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    hq_ctx* c = new hq_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete c;
        return fail(nullptr, HQ_ERR_CUDA, "device %d init failed: %s", device, cudaGetErrorString(e));
    }
    if (prop.major < 10) {
        cudaStreamDestroy(c->stream);
        delete c;
        return fail(nullptr, HQ_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    }
    if ((e = c->h_flag.reserve(1)) == cudaSuccess) c->h_flag.p[0] = 0ull;
    if (e == cudaSuccess && (e = c->d_export_counter.reserve(1)) == cudaSuccess) e = cudaMemsetAsync(c->d_export_counter.p, 0, sizeof(unsigned), c->stream);
    if (e != cudaSuccess || (e = c->d_table.reserve(512)) != cudaSuccess || (e = hq::launch_decode_table(c->d_table.p, c->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(c->stream)) != cudaSuccess) {
        cudaStreamDestroy(c->stream);
        delete c;
        return fail(nullptr, HQ_ERR_CUDA, "device %d: decode table kernel failed: %s (is the library built for this GPU?)", device, cudaGetErrorString(e));
    }
    {
        const char* d = std::getenv("HQ_DIRECT_IO");
        c->direct_io = !(d && d[0] == '0');
        const char* sm = std::getenv("HQ_SMALL_EVAL");   // 0: keep the two-launch latency path (A/B measurements)
        c->small_eval = !(sm && sm[0] == '0');
        const char* pe = std::getenv("HQ_PERSIST");   // 1: persistent evaluator inside hq_find_best_quantization (measured: -9 %; off by default)
        c->persist_enabled = pe && pe[0] == '1';
    }
    {   // HQ_SC_UNFUSED=1: the S-CIELAB candidate stage as two kernels per candidate with a 7-plane intermediate (round 1; A/B runs)
        const char* u = std::getenv("HQ_SC_UNFUSED");
        c->sc_unfused = u && u[0] == '1';
    }
    {   // HQ_CUDA_GRAPHS=1 turns hq_set_graphs on for every new context
        const char* g = std::getenv("HQ_CUDA_GRAPHS");
        c->use_graphs = g && g[0] == '1';
    }
    c->sm_count = prop.multiProcessorCount;
    cudaDeviceGetAttribute(&c->clock_khz, cudaDevAttrClockRate, device);
    snprintf(c->name, sizeof c->name, "%s", prop.name);
    *out = c;
    return HQ_OK;
}

void hq_destroy(hq_ctx* c) {
    if (!c) return;
    if (c->is_multi()) {   // the leader owns its members
        std::vector<hq_ctx*> ms(c->members.begin() + 1, c->members.end());
        c->members.clear();
        for (hq_ctx* m : ms) { m->leader = nullptr; hq_destroy(m); }
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    comm_release(c);
    if (c->ev_image) cudaEventDestroy(c->ev_image);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->ev3) cudaEventDestroy(c->ev3);
    if (c->ev4) cudaEventDestroy(c->ev4);
    if (c->ev5) cudaEventDestroy(c->ev5);
    c->d_rgb.release(); c->d_flag.release(); c->d_lab.release(); c->d_unit.release(); c->d_table.release(); c->d_pal.release();
    c->d_pal_lab.release(); c->d_pal_rgb.release(); c->d_results.release(); c->d_idx.release();
    c->d_out_rgb.release(); c->d_out_f32.release(); c->h_pal.release(); c->h_results.release(); c->h_flag.release(); c->d_export_counter.release(); c->d_results_small.release(); c->h_persist.release(); c->d_persist_cmd.release();
    c->d_sc_filters.release(); c->d_sc_opp.release(); c->d_sc_tmp.release(); c->d_sc_lab.release(); c->d_sc_tab.release(); c->d_sc_err.release();
    c->d_sc_lab2.release(); c->d_sc_map.release(); c->d_sc_rgb2.release(); c->d_sc_map8.release();
    c->pr_own.release(); c->pr_all.release(); c->d_pr_scratch.release(); c->d_big_d2.release(); c->d_big_idx.release();
    c->d_pr_stats.release(); c->h_pr_small.release();
    delete c;
}

const char* hq_last_error(const hq_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int hq_device_info(const hq_ctx* c, int* sm_count, int* sm_clock_khz, char* name, int name_len) {
    if (!c) return HQ_ERR_INVALID;
    if (sm_count) *sm_count = c->sm_count;
    if (sm_clock_khz) *sm_clock_khz = c->clock_khz;
    if (name && name_len > 0) snprintf(name, (size_t)name_len, "%s", c->name);
    return HQ_OK;
}

uint64_t hq_image_pixels(const hq_ctx* c) {
    if (!c) return 0;
    if (c->is_multi()) return (uint64_t)c->m_width * (uint64_t)c->m_rows;  // the whole image the members share
    return (uint64_t)(c->own_hi - c->own_lo);
}

namespace {
// the image of a context has changed hands: forget a foreign-stream conversion (hq_set_image_u8_device)
void image_on_own_stream(hq_ctx* c) { c->image_foreign = false; }

// row block of member i of G (rows i*H/G .. (i+1)*H/G) plus the neighbour rows the S-CIELAB stage's vertical
// filter reaches; an empty block carries no halo (there is nothing to filter)
struct Shard { int r0, own, ht, hb; };
Shard shard_of(int H, int G, int i, int halo) {
    const int r0 = (int)((long long)H * i / G), r1 = (int)((long long)H * (i + 1) / G);
    if (r1 == r0) return Shard{r0, 0, 0, 0};
    return Shard{r0, r1 - r0, halo < r0 ? halo : r0, halo < H - r1 ? halo : H - r1};
}
// halo rows of a multi-device context: the filter bank configured when the image arrives (21 taps by default -> 10)
int multi_halo(const hq_ctx* c) { const int h = c->sc_taps / 2; return h > 10 ? h : 10; }
int member_rc(hq_ctx* leader, hq_ctx* m, int rc) {
    if (rc != HQ_OK && m != leader) leader->err = "device " + std::to_string(m->device) + ": " + m->err;
    return rc;
}
int sync_members(hq_ctx* c) {
    for (hq_ctx* m : c->members) {
        HQ_CUDA(c, cudaSetDevice(m->device));
        HQ_CUDA(c, cudaStreamSynchronize(m->stream));
    }
    return bind_device(c);
}

int set_image_u8_shard(hq_ctx* c, const uint8_t* rgb, int width, int own_rows, int halo_top, int halo_bottom,
                       int global_row0, int global_rows, int whitepoint, bool sync) {
    const long long rows = (long long)halo_top + own_rows + halo_bottom;
    if (width < 0 || own_rows < 0 || halo_top < 0 || halo_bottom < 0 || (!rgb && (size_t)width * rows > 0)) return fail(c, HQ_ERR_INVALID, "bad image arguments");
    if (global_row0 < halo_top || global_row0 + own_rows + halo_bottom > global_rows)
        return fail(c, HQ_ERR_INVALID, "shard rows [%d,%d) with halos %d/%d do not fit a %d-row image", global_row0, global_row0 + own_rows, halo_top, halo_bottom, global_rows);
    if (whitepoint != HQ_WHITEPOINT_D65 && whitepoint != HQ_WHITEPOINT_D50) return fail(c, HQ_ERR_INVALID, "unknown white point %d", whitepoint);
    int rc = bind_device(c); if (rc) return rc;
    c->have_image = false;
    c->image_f32 = false;
    c->n = (size_t)width * rows;
    c->stride = hq::plane_stride(c->n);
    HQ_CUDA(c, c->d_rgb.reserve(c->n * 3 > 0 ? c->n * 3 : 1));
    if (c->n) HQ_CUDA(c, cudaMemcpyAsync(c->d_rgb.p, rgb, c->n * 3, cudaMemcpyHostToDevice, c->stream));
    image_on_own_stream(c);
    rc = convert_image(c, width, own_rows, halo_top, halo_bottom, global_row0, global_rows, whitepoint, c->stream); if (rc) return rc;
    if (sync) HQ_CUDA(c, cudaStreamSynchronize(c->stream));
    return HQ_OK;
}
const char* kMultiShards = "a multi-device context shards the image itself: pass the whole image to hq_set_image_u8 / hq_set_image_f32_planar";
}  // namespace

int hq_set_image_u8_sharded(hq_ctx* c, const uint8_t* rgb, int width, int own_rows, int halo_top, int halo_bottom,
                            int global_row0, int global_rows, int whitepoint) try {
    if (!c) return HQ_ERR_INVALID;
    if (c->is_multi()) return fail(c, HQ_ERR_UNSUPPORTED, "%s", kMultiShards);
    return set_image_u8_shard(c, rgb, width, own_rows, halo_top, halo_bottom, global_row0, global_rows, whitepoint, true);
} catch (const std::exception& ex) { return api_exception(c, ex); }

namespace {
// waits for a float image's conversion and judges the range flag its kernel raised
int f32_image_verdict(hq_ctx* c) {
    int rc = bind_device(c); if (rc) return rc;
    unsigned int bad = 0;
    HQ_CUDA(c, cudaMemcpyAsync(&bad, c->d_flag.p, sizeof bad, cudaMemcpyDeviceToHost, c->stream));
    HQ_CUDA(c, cudaStreamSynchronize(c->stream));
    if (bad) {
        c->have_image = false;
        return fail(c, HQ_ERR_INVALID, "float image values must lie in [0,1] (Icy's rescaled convertToType, HybridQuantization.java:95)");
    }
    return HQ_OK;
}
// sync = false: upload and conversion are only enqueued; the caller runs f32_image_verdict later
int set_image_f32_shard(hq_ctx* c, const float* r, const float* g, const float* b, int width, int own_rows, int halo_top,
                        int halo_bottom, int global_row0, int global_rows, int whitepoint, bool sync) {
    const long long rows = (long long)halo_top + own_rows + halo_bottom;
    if (width < 0 || own_rows < 0 || halo_top < 0 || halo_bottom < 0 || ((!r || !g || !b) && (size_t)width * rows > 0))
        return fail(c, HQ_ERR_INVALID, "bad image arguments");
    if (global_row0 < halo_top || global_row0 + own_rows + halo_bottom > global_rows)
        return fail(c, HQ_ERR_INVALID, "shard rows [%d,%d) with halos %d/%d do not fit a %d-row image", global_row0, global_row0 + own_rows, halo_top, halo_bottom, global_rows);
    if (whitepoint != HQ_WHITEPOINT_D65 && whitepoint != HQ_WHITEPOINT_D50) return fail(c, HQ_ERR_INVALID, "unknown white point %d", whitepoint);
    int rc = bind_device(c); if (rc) return rc;
    c->have_image = false;
    c->image_f32 = true;
    c->n = (size_t)width * rows;
    c->stride = hq::plane_stride(c->n);
    HQ_CUDA(c, c->d_unit.reserve(3 * c->stride > 0 ? 3 * c->stride : 1));
    HQ_CUDA(c, c->d_flag.reserve(1));
    HQ_CUDA(c, cudaMemsetAsync(c->d_flag.p, 0, sizeof(unsigned int), c->stream));
    const float* planes[3] = {r, g, b};
    for (int pl = 0; pl < 3 && c->n; ++pl)
        HQ_CUDA(c, cudaMemcpyAsync(c->d_unit.p + (size_t)pl * c->stride, planes[pl], c->n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    image_on_own_stream(c);
    rc = convert_image(c, width, own_rows, halo_top, halo_bottom, global_row0, global_rows, whitepoint, c->stream); if (rc) return rc;
    return sync ? f32_image_verdict(c) : HQ_OK;
}
}  // namespace

int hq_set_image_f32_planar_sharded(hq_ctx* c, const float* r, const float* g, const float* b, int width, int own_rows, int halo_top,
                                    int halo_bottom, int global_row0, int global_rows, int whitepoint) try {
    if (!c) return HQ_ERR_INVALID;
    if (c->is_multi()) return fail(c, HQ_ERR_UNSUPPORTED, "%s", kMultiShards);
    return set_image_f32_shard(c, r, g, b, width, own_rows, halo_top, halo_bottom, global_row0, global_rows, whitepoint, true);
} catch (const std::exception& ex) { return api_exception(c, ex); }

int hq_set_image_f32_planar(hq_ctx* c, const float* r, const float* g, const float* b, int width, int rows, int whitepoint) try {
    if (!c) return HQ_ERR_INVALID;
    if (!c->is_multi()) return set_image_f32_shard(c, r, g, b, width, rows, 0, 0, 0, rows, whitepoint, true);
    if (width < 0 || rows < 0 || ((!r || !g || !b) && (size_t)width * rows > 0)) return fail(c, HQ_ERR_INVALID, "bad image arguments");
    const int G = (int)c->members.size(), halo = multi_halo(c);
    for (hq_ctx* m : c->members) m->have_image = false;
    for (int i = 0; i < G; ++i) {   // every member's upload and conversion are enqueued before any is waited for
        const Shard sh = shard_of(rows, G, i, halo);
        const size_t off = (size_t)(sh.r0 - sh.ht) * width;
        hq_ctx* m = c->members[i];
        const int rc = member_rc(c, m, set_image_f32_shard(m, r + off, g + off, b + off, width, sh.own, sh.ht, sh.hb, sh.r0, rows, whitepoint, false));
        if (rc) return rc;
    }
    int verdict = HQ_OK;
    for (hq_ctx* m : c->members) { const int rc = member_rc(c, m, f32_image_verdict(m)); if (rc && !verdict) verdict = rc; }
    if (verdict) { for (hq_ctx* m : c->members) m->have_image = false; return verdict; }
    c->m_width = width; c->m_rows = rows;
    return bind_device(c);
} catch (const std::exception& ex) { return api_exception(c, ex); }

int hq_set_image_u8(hq_ctx* c, const uint8_t* rgb, int width, int rows, int whitepoint) try {
    if (!c) return HQ_ERR_INVALID;
    if (!c->is_multi()) return set_image_u8_shard(c, rgb, width, rows, 0, 0, 0, rows, whitepoint, true);
    if (width < 0 || rows < 0 || (!rgb && (size_t)width * rows > 0)) return fail(c, HQ_ERR_INVALID, "bad image arguments");
    const int G = (int)c->members.size(), halo = multi_halo(c);
    for (hq_ctx* m : c->members) m->have_image = false;
    for (int i = 0; i < G; ++i) {
        const Shard sh = shard_of(rows, G, i, halo);
        hq_ctx* m = c->members[i];
        const int rc = member_rc(c, m, set_image_u8_shard(m, rgb + (size_t)(sh.r0 - sh.ht) * width * 3, width, sh.own, sh.ht, sh.hb, sh.r0, rows, whitepoint, false));
        if (rc) return rc;
    }
    c->m_width = width; c->m_rows = rows;
    return sync_members(c);
} catch (const std::exception& ex) { return api_exception(c, ex); }

int hq_set_image_u8_device(hq_ctx* c, const void* d_rgb, int width, int rows, int whitepoint, void* stream) {
    if (!c) return HQ_ERR_INVALID;
    if (c->is_multi()) return fail(c, HQ_ERR_UNSUPPORTED, "a device buffer belongs to one device: a multi-device context takes host images");
    if (width < 0 || rows < 0 || (!d_rgb && (size_t)width * rows > 0)) return fail(c, HQ_ERR_INVALID, "bad image arguments");
    if (whitepoint != HQ_WHITEPOINT_D65 && whitepoint != HQ_WHITEPOINT_D50) return fail(c, HQ_ERR_INVALID, "unknown white point %d", whitepoint);
    int rc = bind_device(c); if (rc) return rc;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    c->have_image = false;
    c->image_f32 = false;
    c->n = (size_t)width * rows;
    c->stride = hq::plane_stride(c->n);
    HQ_CUDA(c, c->d_rgb.reserve(c->n * 3 > 0 ? c->n * 3 : 1));
    if (c->n) HQ_CUDA(c, cudaMemcpyAsync(c->d_rgb.p, d_rgb, c->n * 3, cudaMemcpyDeviceToDevice, st));
    rc = convert_image(c, width, rows, 0, 0, 0, rows, whitepoint, st); if (rc) return rc;
    c->image_foreign = st != c->stream;
    if (c->image_foreign) {
        // the conversion runs on the caller's stream; the context's own (non-blocking) stream — every host-buffer entry —
        // and any other stream handed to hq_eval_palettes_device wait for it through this event
        if (!c->ev_image) HQ_CUDA(c, cudaEventCreateWithFlags(&c->ev_image, cudaEventDisableTiming));
        HQ_CUDA(c, cudaEventRecord(c->ev_image, st));
        HQ_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_image, 0));
    }
    return HQ_OK;
}

namespace {
// own pixels of three device planes [3][stride] -> planes[pl * plane_len + at ...] (a member of a multi-device context
// writes its row block into the whole image's planes)
int own_planes_to_host(hq_ctx* c, const float* d_planes, float* planes, size_t plane_len, size_t at) {
    int rc = bind_device(c); if (rc) return rc;
    const size_t no = c->own_hi - c->own_lo;
    for (int pl = 0; pl < 3 && no; ++pl)
        HQ_CUDA(c, cudaMemcpyAsync(planes + (size_t)pl * plane_len + at, d_planes + (size_t)pl * c->stride + c->own_lo, no * sizeof(float),
                                   cudaMemcpyDeviceToHost, c->stream));
    HQ_CUDA(c, cudaStreamSynchronize(c->stream));
    return HQ_OK;
}
}  // namespace

int hq_get_lab(hq_ctx* c, float* planes) {
    if (!c || !planes) return HQ_ERR_INVALID;
    if (!c->have_image) return fail(c, HQ_ERR_NO_IMAGE, "no image");
    if (!c->is_multi()) return own_planes_to_host(c, c->d_lab.p, planes, c->own_hi - c->own_lo, 0);
    const size_t n_all = (size_t)c->m_width * c->m_rows;
    for (hq_ctx* m : c->members) {
        const int rc = member_rc(c, m, own_planes_to_host(m, m->d_lab.p, planes, n_all, (size_t)m->g_row0 * m->width));
        if (rc) return rc;
    }
    return bind_device(c);
}

int hq_result_words(int K, int flags) { return hq::result_words(K, (flags & HQ_EVAL_SUMS) != 0); }

int hq_eval_palettes_device(hq_ctx* c, const void* d_palettes, int B, int K, int space, int flags, void* d_results, void* stream) {
    int rc = check_eval_args(c, B, K, space); if (rc) return rc;
    if (!d_palettes || !d_results) return fail(c, HQ_ERR_INVALID, "NULL device buffer");
    rc = bind_device(c); if (rc) return rc;
    if (c->is_multi()) return fail(c, HQ_ERR_UNSUPPORTED, "a device buffer belongs to one device: use hq_eval_palettes on a multi-device context");
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    if (c->image_foreign && st != c->stream) HQ_CUDA(c, cudaStreamWaitEvent(st, c->ev_image, 0));  // image converted on another stream
    unsigned long long* words = static_cast<unsigned long long*>(d_results);
    if ((flags & HQ_EVAL_ALLREDUCE) && reduces(c)) { rc = peer_check(c); if (rc) return rc; }   // (an EARLIER exchange that timed out: this entry never waits)
    if (!(flags & HQ_EVAL_ALLREDUCE) || !reduces(c))
        return eval_device(c, static_cast<const float*>(d_palettes), B, K, space, flags & ~HQ_EVAL_ALLREDUCE, words, nullptr, st);
    // scoring + exchange: over peer memory the last CTA of the scoring kernel all-reduces (K <= 32) or a one-CTA launch behind it does
    const size_t nwords = (size_t)B * hq::result_words(K, (flags & HQ_EVAL_SUMS) != 0);
    if (!peer_ready(c, nwords)) {
        rc = eval_device(c, static_cast<const float*>(d_palettes), B, K, space, flags & ~HQ_EVAL_ALLREDUCE, words, nullptr, st); if (rc) return rc;
        return reduce_words(c, words, nwords, st);
    }
    HQ_CUDA(c, c->d_export_counter.reserve(1));
    hq::ExportTail tail;
    tail.peer_only = true; tail.counter = c->d_export_counter.p; tail.src = words; tail.nwords = (unsigned)nwords;
    tail.peer = peer_next(c);
    bool used = false;
    rc = eval_device(c, static_cast<const float*>(d_palettes), B, K, space, flags & ~HQ_EVAL_ALLREDUCE, words, nullptr, st, &tail, &used); if (rc) return rc;
    if (!used) HQ_CUDA(c, hq::launch_peer_allreduce(tail.peer, words, nwords, nullptr, nullptr, 0, st));
    return HQ_OK;
}

namespace {
// palette colours must be sRGB in [0,1] (what SWASA.java:93-106 produces): NaN or out-of-range values would leave the domain
// hq_srgb_decode is verified on.  The copy into the pinned staging buffer is the one pass over them anyway.
bool copy_palettes_checked(float* dst, const float* src, size_t npal) {
    bool ok = true;
    for (size_t i = 0; i < npal; i += 4) {
        const float r = src[i], g = src[i + 1], b = src[i + 2];
        dst[i] = r; dst[i + 1] = g; dst[i + 2] = b; dst[i + 3] = src[i + 3];
        ok &= (r >= 0.f) & (r <= 1.f) & (g >= 0.f) & (g <= 1.f) & (b >= 0.f) & (b <= 1.f);  // false for NaN
    }
    return ok;
}
const char* kBadPalette = "palette colours must be finite sRGB values in [0,1] (SWASA.java:93-106 clamps them)";

// ---- persistent evaluator of a small search: started by hq_find_best_quantization, fed by hq_eval_palettes
void persist_end(hq_ctx* c) {
    if (!c->persist_on) return;
    c->persist_on = false;
    cudaSetDevice(c->device);
    *static_cast<volatile unsigned long long*>(c->h_persist.p) = hq::kPersistQuitCmd;
    cudaStreamSynchronize(c->stream);
}
// true: the kernel is running and every evaluation with this signature goes through its mailbox
bool persist_begin(hq_ctx* c, int B, int K, int space, bool sums) {
    if (!c->persist_enabled || !c->small_eval || !c->direct_io || c->is_multi() || reduces(c) || c->profiling || c->use_graphs || c->n == 0 ||
        c->delta_e != HQ_DELTAE_CIE76 || K > hq::kDirectMaxColors || (long long)B * K > hq::kSmallPalColors)
        return false;
    if (cudaSetDevice(c->device) != cudaSuccess) return false;
    const size_t npal = (size_t)B * K * 4, nwords = (size_t)B * hq::result_words(K, sums);
    if (c->h_pal.reserve(npal) != cudaSuccess || c->h_results.reserve(nwords) != cudaSuccess || c->h_persist.reserve(16) != cudaSuccess ||
        c->d_persist_cmd.reserve(1) != cudaSuccess || c->d_pal.reserve(npal) != cudaSuccess)
        return false;
    if (space == HQ_SPACE_SRGB && ensure_unit(c, c->stream) != HQ_OK) return false;
    if (nwords > c->d_results_small.cap) {
        if (c->d_results_small.reserve(nwords > 4096 ? nwords : 4096) != cudaSuccess) return false;
        if (cudaMemsetAsync(c->d_results_small.p, 0, c->d_results_small.cap * 8, c->stream) != cudaSuccess) return false;
    }
    const unsigned long long first = c->export_seq;
    c->h_persist.p[0] = first; c->h_persist.p[8] = 0ull;
    if (cudaMemcpyAsync(c->d_persist_cmd.p, c->h_persist.p, 8, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return false;
    hq::AssignArgs a;
    a.lab = c->d_lab.p; a.unit = c->d_unit.p; a.n = c->n; a.stride = c->stride;
    a.pal_lab = nullptr; a.pal_rgb = nullptr;
    a.B = B; a.K = K; a.space = space; a.want_sums = sums;
    a.results = c->d_results_small.p; a.idx_out = nullptr; a.sm_count = c->sm_count;
    a.own_lo = c->own_lo; a.own_hi = c->own_hi;
    a.variant = 1;
    a.tail.host_dst = c->h_results.p; a.tail.host_flag = c->h_flag.p; a.tail.seq = 0; a.tail.counter = c->d_export_counter.p;
    a.tail.src = c->d_results_small.p; a.tail.nwords = (unsigned)nwords;
    unsigned long long idle_ns = 200ull * 1000000ull;   // a search posts an evaluation every ~15 us; HQ_PERSIST_IDLE_MS
    if (const char* t = std::getenv("HQ_PERSIST_IDLE_MS")) { const long long ms = std::atoll(t); if (ms > 0) idle_ns = (unsigned long long)ms * 1000000ull; }
    const cudaError_t e = hq::launch_assign_persist(a, c->h_pal.p, c->d_pal.p, c->whitepoint, c->h_persist.p, c->h_persist.p + 8, c->d_persist_cmd.p, first, idle_ns, c->stream);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }   // (not resident in one wave, no cooperative launch: one launch per evaluation)
    c->persist_on = true;
    c->persist_B = B; c->persist_K = K; c->persist_space = space; c->persist_sums = sums; c->persist_image_gen = c->image_gen;
    return true;
}

struct EvalPlan {
    int B, K, space, flags, words;
    size_t npal, nwords;
    bool direct;   // palettes read from / results exported to pinned host memory by kernels (small transfers)
};
// everything an evaluation may allocate or build lazily on context m (outside any graph capture)
int eval_prepare(hq_ctx* m, const EvalPlan& e) {
    int rc = bind_device(m); if (rc) return rc;
    const int K8 = hq::padded_colors(e.K);
    HQ_CUDA(m, m->d_pal.reserve(e.npal));
    HQ_CUDA(m, m->d_results.reserve(e.nwords));
    HQ_CUDA(m, m->d_pal_lab.reserve((size_t)e.B * K8));
    HQ_CUDA(m, m->d_pal_rgb.reserve((size_t)e.B * K8));
    if (e.space == HQ_SPACE_SRGB) { rc = ensure_unit(m, m->stream); if (rc) return rc; }
    return prepare_pruned(m, e.K, e.space, e.flags, false, m->stream);
}
// palettes (pinned, portable host memory) -> result words of this context's pixels in m->d_results, on m->stream
int eval_enqueue(hq_ctx* m, const float* h_pal, const EvalPlan& e, const hq::ExportTail* tail, bool* tail_used) {
    int rc = bind_device(m); if (rc) return rc;
    if (e.direct) return eval_device(m, h_pal, e.B, e.K, e.space, e.flags, m->d_results.p, nullptr, m->stream, tail, tail_used);
    HQ_CUDA(m, cudaMemcpyAsync(m->d_pal.p, h_pal, e.npal * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    return eval_device(m, m->d_pal.p, e.B, e.K, e.space, e.flags, m->d_results.p, nullptr, m->stream);
}
void unpack_results(const unsigned long long* h, int B, int K, int words, int64_t* err_fx, uint64_t* counts, int64_t* sums_fx) {
    for (int b = 0; b < B; ++b) {
        const unsigned long long* w = h + (size_t)b * words;
        if (err_fx) err_fx[b] = (int64_t)w[0];
        if (counts) std::memcpy(counts + (size_t)b * K, w + 1, sizeof(uint64_t) * K);
        if (sums_fx) std::memcpy(sums_fx + (size_t)b * K * 3, w + 1 + K, sizeof(int64_t) * 3 * K);
    }
}
// the exchange step: sum of the result words over every shard, left in place on every device
int eval_reduce(hq_ctx* c, size_t nwords) {
    if (c->is_multi()) {
        std::vector<unsigned long long*> bufs;
        for (hq_ctx* m : c->members) bufs.push_back(m->d_results.p);
        return group_reduce(c, bufs, nwords);
    }
    return reduce_words(c, c->d_results.p, nwords, c->stream);
}
}  // namespace

int hq_eval_palettes(hq_ctx* c, const float* palettes, int B, int K, int space, int flags, int64_t* err_fx, uint64_t* counts, int64_t* sums_fx) try {
    int rc = check_eval_args(c, B, K, space); if (rc) return rc;
    if (!palettes) return fail(c, HQ_ERR_INVALID, "palettes is NULL");
    const bool sums = (flags & HQ_EVAL_SUMS) != 0;
    if (sums_fx && !sums) return fail(c, HQ_ERR_INVALID, "sums_fx requires HQ_EVAL_SUMS");
    rc = bind_device(c); if (rc) return rc;
    EvalPlan e{};
    e.B = B; e.K = K; e.space = space; e.flags = flags;
    e.npal = (size_t)B * K * 4;
    e.words = hq::result_words(K, sums);
    e.nwords = (size_t)B * e.words;
    const size_t npal = e.npal, nwords = e.nwords;
    const int words = e.words;
    HQ_CUDA(c, c->h_pal.reserve(npal));
    HQ_CUDA(c, c->h_results.reserve(nwords));
    if (!copy_palettes_checked(c->h_pal.p, palettes, npal)) return fail(c, HQ_ERR_INVALID, "%s", kBadPalette);
    const bool multi = c->is_multi(), reduce = reduces(c);
    std::vector<hq_ctx*> self(1, c);
    const std::vector<hq_ctx*>& targets = multi ? c->members : self;
    if (c->delta_e != HQ_DELTAE_CIE76) {
        // scope row f4: the assignment kernels score with the squared distance they minimise; another dE is scored from the index
        // images in a second pass, with B words of NaN-pixel counts behind the result words (the CIE94 branch's latent NaN)
        const bool idx16 = K > 256;
        for (size_t i = targets.size(); i-- > 0;) {
            hq_ctx* m = targets[i];
            m->delta_e = c->delta_e;
            rc = member_rc(c, m, bind_device(m)); if (rc) return rc;
            HQ_CUDA(m, m->d_pal.reserve(npal));
            HQ_CUDA(m, m->d_results.reserve(nwords + B));
            HQ_CUDA(m, m->d_sc_err.reserve(B));
            HQ_CUDA(m, m->d_idx.reserve((size_t)B * (m->stride ? m->stride : 1) * (idx16 ? 2 : 1)));
            HQ_CUDA(m, cudaMemcpyAsync(m->d_pal.p, c->h_pal.p, npal * sizeof(float), cudaMemcpyHostToDevice, m->stream));
            rc = member_rc(c, m, eval_device(m, m->d_pal.p, B, K, space, flags & ~HQ_EVAL_PRUNE, m->d_results.p, m->d_idx.p, m->stream)); if (rc) return rc;
            HQ_CUDA(m, cudaMemsetAsync(m->d_results.p + nwords, 0, (size_t)B * 8, m->stream));
            HQ_CUDA(m, cudaMemsetAsync(m->d_sc_err.p, 0, (size_t)B * 8, m->stream));
            HQ_CUDA(m, hq::launch_sc_score_indices(m->d_idx.p, idx16, m->d_lab.p, m->stride, m->own_lo, m->own_hi, m->d_pal_lab.p, hq::padded_colors(K), B,
                                                   m->delta_e, m->d_sc_err.p, m->d_results.p + nwords, m->sm_count, m->stream));
            HQ_CUDA(m, cudaMemcpy2DAsync(m->d_results.p, (size_t)words * 8, m->d_sc_err.p, 8, 8, (size_t)B, cudaMemcpyDeviceToDevice, m->stream));
        }
        if (reduce) { rc = eval_reduce(c, nwords + B); if (rc) return rc; }
        rc = bind_device(c); if (rc) return rc;
        HQ_CUDA(c, c->h_results.reserve(nwords + B));
        HQ_CUDA(c, cudaMemcpyAsync(c->h_results.p, c->d_results.p, (nwords + B) * 8, cudaMemcpyDeviceToHost, c->stream));
        HQ_CUDA(c, wait_stream(c->stream));
        for (int b = 0; b < B; ++b)
            if (c->h_results.p[nwords + b]) c->h_results.p[(size_t)b * words] = (unsigned long long)HQ_ERR_FX_NAN;
        unpack_results(c->h_results.p, B, K, words, err_fx, counts, sums_fx);
        return HQ_OK;
    }
    for (hq_ctx* m : targets) { rc = member_rc(c, m, eval_prepare(m, e)); if (rc) return rc; }
    rc = bind_device(c); if (rc) return rc;
    hq_ctx::EvalKey key;
    key.B = B; key.K = K; key.space = space; key.flags = flags; key.image_gen = c->image_gen; key.d_pal = c->d_pal.p; key.d_results = c->d_results.p;
    key.h_pal = c->h_pal.p; key.h_results = c->h_results.p; key.d_pal_lab = c->d_pal_lab.p; key.d_pal_rgb = c->d_pal_rgb.p;
    const bool graphable = c->use_graphs && !reduce && !c->profiling;
    if (graphable && c->graph_exec && key == c->graph_key) {
        HQ_CUDA(c, cudaGraphLaunch(c->graph_exec, c->stream));  // the third and later identical calls: one launch for the whole step
    } else {
        // first call with this signature: plain launches (also configures the kernels); second: captured into a graph
        const bool capture = graphable && key == c->seen_key;
        c->seen_key = key;
        // (small transfers only: one CTA pushing 131 KB of results over PCIe took 0.16 ms at 64 candidates x 256 colours, a DMA copy 5 us)
        e.direct = !graphable && c->direct_io && npal * sizeof(float) <= 65536 && nwords * 8 <= 32768;
        // ONE launch (round 2): small searches (B*K <= 192 colours, K <= 32 — the plugin's defaults are 8 colours x 4 candidates) pass the
        // palettes as a kernel parameter; every CTA converts its candidate's palette itself and the last CTA exports and re-zeroes the
        // result words (d_results_small is only ever touched by these launches: zero between them).
        // a sharded image: the exchange rides in the exporting CTA over peer memory when that path is open (hq_kernels.cuh)
        const bool peer = reduce && e.direct && peer_ready(c, nwords);
        bool all_nonempty = true;
        for (hq_ctx* m : targets) all_nonempty = all_nonempty && m->n > 0;
        if (c->persist_on) {
            // a search is running its persistent evaluator: this evaluation is one mailbox write (the palettes are already in h_pal)
            if (B == c->persist_B && K == c->persist_K && space == c->persist_space && sums == c->persist_sums && flags == (sums ? HQ_EVAL_SUMS : 0) &&
                c->image_gen == c->persist_image_gen && *static_cast<volatile unsigned long long*>(c->h_persist.p + 8) == 0ull) {
                const unsigned long long seq = ++c->export_seq;
                std::atomic_thread_fence(std::memory_order_release);
                *static_cast<volatile unsigned long long*>(c->h_persist.p) = seq;
                const cudaError_t we = wait_flag(c->h_flag.p, seq, c->stream);
                if (we == cudaSuccess) goto unpack;
                cudaGetLastError();
                if (we != cudaErrorUnknown) return fail(c, HQ_ERR_CUDA, "persistent evaluation failed: %s", cudaGetErrorString(we));
                // the kernel left (idle time-out) before it saw the command: one launch per evaluation from here on
            }
            persist_end(c);
        }
        if (e.direct && c->small_eval && (!reduce || peer) && all_nonempty && !c->profiling && K <= hq::kDirectMaxColors && (long long)B * K <= hq::kSmallPalColors &&
            !(flags & (HQ_EVAL_PRUNE | HQ_EVAL_FORCE_DIRECT | HQ_EVAL_FORCE_CHUNKED | HQ_EVAL_FORCE_PREFILTER))) {
            const unsigned long long seq = ++c->export_seq;
            cudaError_t le = cudaSuccess;
            for (size_t i = targets.size(); i-- > 0 && le == cudaSuccess;) {   // the leader last: its launch exports
                hq_ctx* m = targets[i];
                rc = bind_device(m); if (rc) return rc;
                if (nwords > m->d_results_small.cap) {
                    HQ_CUDA(c, m->d_results_small.reserve(nwords > 4096 ? nwords : 4096));
                    HQ_CUDA(c, cudaMemsetAsync(m->d_results_small.p, 0, m->d_results_small.cap * 8, m->stream));
                }
                hq::AssignArgs a;
                a.lab = m->d_lab.p; a.unit = m->d_unit.p; a.n = m->n; a.stride = m->stride;
                a.pal_lab = nullptr; a.pal_rgb = nullptr;
                a.B = B; a.K = K; a.space = space; a.want_sums = sums;
                a.results = m->d_results_small.p; a.idx_out = nullptr; a.sm_count = m->sm_count;
                a.own_lo = m->own_lo; a.own_hi = m->own_hi;
                a.variant = 1;
                a.tail.host_dst = m == c ? c->h_results.p : nullptr; a.tail.host_flag = m == c ? c->h_flag.p : nullptr; a.tail.seq = seq; a.tail.counter = m->d_export_counter.p;
                a.tail.src = m->d_results_small.p; a.tail.nwords = (unsigned)nwords;
                a.tail.peer_only = m != c;
                if (peer) a.tail.peer = peer_next(m);
                le = hq::launch_assign_small(a, c->h_pal.p, m->whitepoint, m->stream);
            }
            rc = bind_device(c); if (rc) return rc;
            cudaError_t we = le;
            if (le == cudaSuccess) we = wait_flag(c->h_flag.p, seq, c->stream);
            if (we != cudaSuccess) {
                for (hq_ctx* m : targets) { cudaSetDevice(m->device); m->d_results_small.release(); }   // whatever the failed launch left behind is not reused
                cudaSetDevice(c->device);
                return fail(c, HQ_ERR_CUDA, "one-launch evaluation failed: %s", cudaGetErrorString(we));
            }
            goto unpack;
        }
        if (e.direct) {
            // Latency path (a search iteration is four dependent stream operations; this makes it two): the palette kernel reads
            // the pinned host copy directly (UVA: pinned host memory is device-accessible) and a one-CTA kernel writes the
            // result words plus a sequence number back into pinned host memory, which the host spins on.
            const unsigned long long seq = ++c->export_seq;
            bool tail_used = false;
            for (size_t i = targets.size(); i-- > 0;) {   // the leader (targets[0]) last: its stream carries the export
                hq_ctx* m = targets[i];
                // single GPU, small palettes: the scoring kernel's last CTA exports; sharded with the peer path open: every rank's
                // last CTA exchanges, the leader's also exports; otherwise one-CTA kernels after it
                hq::ExportTail tail;
                tail.host_dst = m == c ? c->h_results.p : nullptr; tail.host_flag = m == c ? c->h_flag.p : nullptr; tail.seq = seq; tail.counter = m->d_export_counter.p;
                tail.src = m->d_results.p; tail.nwords = (unsigned)nwords; tail.peer_only = m != c;
                if (peer) tail.peer = peer_next(m);
                bool used = false;
                rc = member_rc(c, m, eval_enqueue(m, c->h_pal.p, e, (peer || (!reduce && m == c)) ? &tail : nullptr, &used)); if (rc) return rc;
                if (peer && !used)
                    HQ_CUDA(c, hq::launch_peer_allreduce(tail.peer, m->d_results.p, nwords, m == c ? c->h_results.p : nullptr, m == c ? c->h_flag.p : nullptr, seq, m->stream));
                if (m == c) tail_used = used || peer;
            }
            if (reduce && !peer) { rc = eval_reduce(c, nwords); if (rc) return rc; }
            rc = bind_device(c); if (rc) return rc;
            if (!tail_used) HQ_CUDA(c, hq::launch_export_results(c->d_results.p, c->h_results.p, nwords, c->h_flag.p, seq, c->stream));
            HQ_CUDA(c, wait_flag(c->h_flag.p, seq, c->stream));
            goto unpack;
        }
        if (capture) HQ_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t ce = cudaSuccess;
        rc = HQ_OK;
        for (size_t i = targets.size(); i-- > 0 && rc == HQ_OK;) rc = member_rc(c, targets[i], eval_enqueue(targets[i], c->h_pal.p, e, nullptr, nullptr));
        if (rc == HQ_OK && reduce) rc = eval_reduce(c, nwords);
        if (rc == HQ_OK) rc = bind_device(c);
        if (rc == HQ_OK) ce = cudaMemcpyAsync(c->h_results.p, c->d_results.p, nwords * 8, cudaMemcpyDeviceToHost, c->stream);
        if (capture) {
            cudaGraph_t g = nullptr;
            const cudaError_t ec = cudaStreamEndCapture(c->stream, &g);
            if (rc == HQ_OK && ce == cudaSuccess && ec == cudaSuccess && g) {
                if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
                const cudaError_t ei = cudaGraphInstantiate(&c->graph_exec, g, 0);
                cudaGraphDestroy(g);
                if (ei != cudaSuccess) { c->graph_exec = nullptr; return fail(c, HQ_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei)); }
                c->graph_key = key;
                HQ_CUDA(c, cudaGraphLaunch(c->graph_exec, c->stream));
            } else {
                if (g) cudaGraphDestroy(g);
                if (rc != HQ_OK) return rc;
                return fail(c, HQ_ERR_CUDA, "graph capture of the evaluation failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ec));
            }
        } else {
            if (rc != HQ_OK) return rc;
            if (ce != cudaSuccess) return fail(c, HQ_ERR_CUDA, "evaluation launch failed: %s", cudaGetErrorString(ce));
        }
    }
    HQ_CUDA(c, wait_stream(c->stream));
unpack:
    if (reduce) { rc = peer_check(c); if (rc) return rc; }
    unpack_results(c->h_results.p, B, K, words, err_fx, counts, sums_fx);
    return HQ_OK;
} catch (const std::exception& ex) { return api_exception(c, ex); }
double hq_cost(int64_t err_fx, const uint64_t* counts, int K, uint64_t n_total, float delta) {
    double penalty = 0;
    for (int k = 0; k < K; ++k)
        if (counts[k] == 0) penalty += delta;
    if (err_fx == HQ_ERR_FX_NAN) return std::nan("");   // a NaN pixel of the CIE94 branch: the reference's mean, hence its cost, is NaN
    const double sum = (double)err_fx * (1.0 / 16777216.0);
    return sum / (double)n_total + penalty;
}

namespace {
int quantize_one(hq_ctx* c, const float* palette, int K, int space, uint8_t* out_rgb, float* out_f32, uint16_t* out_idx) {
    int rc = check_eval_args(c, 1, K, space); if (rc) return rc;
    rc = bind_device(c); if (rc) return rc;
    const size_t n = c->n;
    const bool idx16 = K > 256;
    const size_t npal = (size_t)K * 4;
    const int words = hq::result_words(K, false);
    HQ_CUDA(c, c->h_pal.reserve(npal));
    HQ_CUDA(c, c->d_pal.reserve(npal));
    HQ_CUDA(c, c->d_results.reserve(words));
    HQ_CUDA(c, c->d_idx.reserve((c->stride ? c->stride : 1) * (idx16 ? 2 : 1)));
    if (!copy_palettes_checked(c->h_pal.p, palette, npal)) return fail(c, HQ_ERR_INVALID, "%s", kBadPalette);
    HQ_CUDA(c, cudaMemcpyAsync(c->d_pal.p, c->h_pal.p, npal * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    const bool big = K > 65535;   // beyond 16-bit index images: the chunked sweep's own 32-bit assignment feeds the output image
    if (big && out_idx) return fail(c, HQ_ERR_UNSUPPORTED, "out_idx holds 16 bits: pass NULL for K=%d > 65,535", K);
    rc = eval_device(c, c->d_pal.p, 1, K, space, (K > HQ_MAX_COLORS || c->prune_mode == HQ_PRUNE_ON) ? HQ_EVAL_PRUNE : 0, c->d_results.p,
                     big ? nullptr : c->d_idx.p, c->stream); if (rc) return rc;
    if (out_rgb) HQ_CUDA(c, c->d_out_rgb.reserve(n * 3 > 0 ? n * 3 : 1));
    if (out_f32) HQ_CUDA(c, c->d_out_f32.reserve(n * 4 > 0 ? n * 4 : 1));
    if ((out_rgb || out_f32) && big)
        HQ_CUDA(c, hq::launch_apply_palette_u32(c->d_big_idx.p, n, c->d_pal.p, out_rgb ? c->d_out_rgb.p : nullptr, out_f32 ? c->d_out_f32.p : nullptr, c->stream));
    else if (out_rgb || out_f32)
        HQ_CUDA(c, hq::launch_apply_palette(c->d_idx.p, idx16, n, c->d_pal.p, K, out_rgb ? c->d_out_rgb.p : nullptr,
                                            out_f32 ? c->d_out_f32.p : nullptr, c->stream));
    const size_t lo = c->own_lo, no = c->own_hi - c->own_lo;  // a shard returns its own rows only
    if (no) {
        if (out_rgb) HQ_CUDA(c, cudaMemcpyAsync(out_rgb, c->d_out_rgb.p + lo * 3, no * 3, cudaMemcpyDeviceToHost, c->stream));
        if (out_f32) HQ_CUDA(c, cudaMemcpyAsync(out_f32, c->d_out_f32.p + lo * 4, no * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    std::vector<uint8_t> idx8;
    if (out_idx && no) {
        if (idx16) {
            HQ_CUDA(c, cudaMemcpyAsync(out_idx, c->d_idx.p + lo * 2, no * 2, cudaMemcpyDeviceToHost, c->stream));
        } else {
            idx8.resize(no);
            HQ_CUDA(c, cudaMemcpyAsync(idx8.data(), c->d_idx.p + lo, no, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    HQ_CUDA(c, cudaStreamSynchronize(c->stream));
    if (out_idx && no && !idx16)
        for (size_t i = 0; i < no; ++i) out_idx[i] = idx8[i];
    return HQ_OK;
}
}  // namespace

int hq_quantize(hq_ctx* c, const float* palette, int K, int space, uint8_t* out_rgb, float* out_f32, uint16_t* out_idx) try {
    if (!c) return HQ_ERR_INVALID;
    if (!palette) return fail(c, HQ_ERR_INVALID, "palette is NULL");
    if (!c->is_multi()) return quantize_one(c, palette, K, space, out_rgb, out_f32, out_idx);
    for (hq_ctx* m : c->members) {   // every member writes its own row block of the outputs
        const size_t at = (size_t)m->g_row0 * m->width;
        const int rc = member_rc(c, m, quantize_one(m, palette, K, space, out_rgb ? out_rgb + at * 3 : nullptr, out_f32 ? out_f32 + at * 4 : nullptr,
                                                    out_idx ? out_idx + at : nullptr));
        if (rc) return rc;
    }
    return bind_device(c);
} catch (const std::exception& ex) { return api_exception(c, ex); }

// ------------------------------------------------------------------ S-CIELAB stage
static int sc_upload_filters(hq_ctx* c) {
    const int T = c->sc_taps;
    std::vector<float> blk((size_t)8 * T);
    const float* f = c->sc_filters7.data();
    for (int t = 0; t < T; ++t) {  // updateOpenCLFilters, ImageManipulation.java:800-841
        blk[3 * t] = f[0 * T + t]; blk[3 * t + 1] = f[3 * T + t]; blk[3 * t + 2] = f[5 * T + t];                    // k1 = first Gaussian of O1,O2,O3
        blk[3 * T + 3 * t] = f[1 * T + t]; blk[3 * T + 3 * t + 1] = f[4 * T + t]; blk[3 * T + 3 * t + 2] = f[6 * T + t];  // k2 = second
        blk[6 * T + t] = f[2 * T + t];                                                                               // k3 = third Gaussian of O1
        blk[7 * T + t] = c->sc_abs3[t];
    }
    c->sc_block = blk;
    HQ_CUDA(c, c->d_sc_filters.reserve(blk.size()));
    HQ_CUDA(c, cudaMemcpyAsync(c->d_sc_filters.p, blk.data(), blk.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    HQ_CUDA(c, cudaStreamSynchronize(c->stream));
    c->sc_image_ready = false;
    return HQ_OK;
}

int hq_scielab_set_filters(hq_ctx* c, const float* filters7, const float* abs3, int taps) try {
    if (!c || !filters7 || !abs3) return c ? fail(c, HQ_ERR_INVALID, "NULL filter arrays") : HQ_ERR_INVALID;
    if (taps < 1 || taps > hq::kMaxScielabTaps || (taps & 1) == 0) return fail(c, HQ_ERR_UNSUPPORTED, "taps must be odd and in [1,%d] (got %d)", hq::kMaxScielabTaps, taps);
    std::vector<hq_ctx*> self(1, c);
    for (hq_ctx* m : (c->is_multi() ? c->members : self)) {   // every member filters its own rows with the same bank
        int rc = bind_device(m); if (rc) return member_rc(c, m, rc);
        m->sc_filters7.assign(filters7, filters7 + (size_t)7 * taps);
        m->sc_abs3.assign(abs3, abs3 + taps);
        m->sc_taps = taps;
        rc = member_rc(c, m, sc_upload_filters(m)); if (rc) return rc;
    }
    return bind_device(c);
} catch (const std::exception& ex) { return api_exception(c, ex); }

int hq_scielab_configure(hq_ctx* c, int dpi, float viewing_distance_cm) try {
    if (!c) return HQ_ERR_INVALID;
    if (dpi < 1 || !(viewing_distance_cm >= 1.0f)) return fail(c, HQ_ERR_INVALID, "dpi >= 1 and viewing distance >= 1 cm required (HybridQuantization.java:229-231)");
    const hq::ScielabProcessor::FilterBank bank = hq::ScielabProcessor::buildFilters(dpi, (double)viewing_distance_cm);
    const std::vector<float> flat = bank.flat();
    return hq_scielab_set_filters(c, flat.data(), bank.absOfilters.data(), bank.taps());
} catch (const std::exception& ex) { return api_exception(c, ex); }

int hq_scielab_force_generic(hq_ctx* c, int enabled) {
    if (!c) return HQ_ERR_INVALID;
    std::vector<hq_ctx*> self(1, c);
    for (hq_ctx* m : (c->is_multi() ? c->members : self)) { m->sc_generic = enabled == 1; m->sc_unfused = enabled == 2; m->sc_image_ready = false; }
    return HQ_OK;
}

int hq_scielab_build_filters(int dpi, float viewing_distance_cm, float* filters7, float* abs3, int* taps) {
    if (!taps || dpi < 1 || !(viewing_distance_cm >= 1.0f)) return HQ_ERR_INVALID;
    const hq::ScielabProcessor::FilterBank bank = hq::ScielabProcessor::buildFilters(dpi, (double)viewing_distance_cm);
    const int T = bank.taps();
    if (filters7 && abs3 && *taps >= T) {
        const std::vector<float> flat = bank.flat();
        std::memcpy(filters7, flat.data(), sizeof(float) * 7 * T);
        std::memcpy(abs3, bank.absOfilters.data(), sizeof(float) * T);
    }
    *taps = T;
    return HQ_OK;
}

int hq_scielab_get_filters(const hq_ctx* c, float* filters7, float* abs3, int* taps) {
    if (!c || !taps) return HQ_ERR_INVALID;
    if (c->sc_taps == 0) return HQ_ERR_INVALID;
    if (filters7 && abs3 && *taps >= c->sc_taps) {
        std::memcpy(filters7, c->sc_filters7.data(), sizeof(float) * 7 * c->sc_taps);
        std::memcpy(abs3, c->sc_abs3.data(), sizeof(float) * c->sc_taps);
    }
    *taps = c->sc_taps;
    return HQ_OK;
}

static hq::ScRows sc_rows(const hq_ctx* c) { return hq::ScRows{c->halo_top, c->own_rows, c->g_row0 - c->halo_top, c->g_rows}; }

// S-CIELAB representation of the resident image (sRGBToScielab, ScielabProcessor.java:374-381)
static int sc_ensure_image(hq_ctx* c) {
    if (c->sc_image_ready) return HQ_OK;
    if (!c->have_image) return fail(c, HQ_ERR_NO_IMAGE, "no image: call hq_set_image_u8 or hq_set_image_f32_planar first");
    if (c->sc_taps == 0) { int rc = hq_scielab_configure(c, 72, 45.0f); if (rc) return rc; }  // plugin defaults :229-231
    const int half = c->sc_taps / 2;
    if (c->width < half || c->g_rows < half)
        return fail(c, HQ_ERR_UNSUPPORTED, "image %dx%d is smaller than the filter half-width %d (the reference's single reflection, "
                    "OptimizedConvolution.cl:20-27, would read out of bounds)", c->width, c->g_rows, half);
    if (c->rows > 65535)  // the filter kernels put one image row per blockIdx.y
        return fail(c, HQ_ERR_UNSUPPORTED, "the S-CIELAB stage handles at most 65,535 local rows per context (got %d): shard the image by rows", c->rows);
    {   // a row shard must carry the neighbours' rows the vertical filter reaches (reflection is at the GLOBAL borders)
        const int need_top = c->g_row0 < half ? c->g_row0 : half;
        const int below = c->g_rows - (c->g_row0 + c->own_rows);
        const int need_bottom = below < half ? below : half;
        if (c->halo_top < need_top || c->halo_bottom < need_bottom || (c->own_rows > 0 && c->rows < half))
            return fail(c, HQ_ERR_UNSUPPORTED, "S-CIELAB on a row shard needs %d halo rows above and %d below (got %d / %d): use hq_set_image_u8_sharded / hq_set_image_f32_planar_sharded",
                        need_top, need_bottom, c->halo_top, c->halo_bottom);
    }
    HQ_CUDA(c, c->d_sc_opp.reserve(3 * c->stride));
    HQ_CUDA(c, c->d_sc_tmp.reserve(7 * c->stride));
    HQ_CUDA(c, c->d_sc_lab.reserve(3 * c->stride));
    if (c->image_f32) HQ_CUDA(c, hq::launch_sc_unit_to_opp(c->d_unit.p, c->n, c->stride, c->d_sc_opp.p, nullptr, c->stream));  // range checked at upload
    else HQ_CUDA(c, hq::launch_sc_rgb_to_opp(c->d_rgb.p, c->n, c->stride, c->d_table.p, c->d_sc_opp.p, c->stream));
    HQ_CUDA(c, hq::launch_sc_original(c->d_sc_opp.p, c->width, c->rows, c->stride, c->d_sc_filters.p, c->sc_generic ? nullptr : c->sc_block.data(), c->sc_taps,
                                      c->whitepoint, sc_rows(c), c->d_sc_tmp.p, c->d_sc_lab.p, c->stream));
    c->sc_image_ready = true;
    return HQ_OK;
}

int hq_scielab_get_image(hq_ctx* c, float* planes) {
    if (!c || !planes) return HQ_ERR_INVALID;
    std::vector<hq_ctx*> self(1, c);
    const size_t n_all = c->is_multi() ? (size_t)c->m_width * c->m_rows : c->own_hi - c->own_lo;
    for (hq_ctx* m : (c->is_multi() ? c->members : self)) {
        if (c->is_multi() && m->own_rows == 0) continue;   // an empty row block
        int rc = bind_device(m); if (rc) return member_rc(c, m, rc);
        rc = member_rc(c, m, sc_ensure_image(m)); if (rc) return rc;
        rc = member_rc(c, m, own_planes_to_host(m, m->d_sc_lab.p, planes, n_all, c->is_multi() ? (size_t)m->g_row0 * m->width : 0)); if (rc) return rc;
    }
    return bind_device(c);
}

// ---- the reference class's one-shot entries on its own layouts (what a Java drop-in with the reference's signatures calls)
int hq_rgb_to_xyz(hq_ctx* c, const float* r, const float* g, const float* b, size_t n, float* xyz4) try {
    if (!c || !xyz4 || ((!r || !g || !b) && n)) return c ? fail(c, HQ_ERR_INVALID, "NULL array") : HQ_ERR_INVALID;
    int rc = bind_device(c); if (rc) return rc;
    DevBuf<float> in, out;
    HQ_CUDA(c, in.reserve(n ? 3 * n : 1));
    HQ_CUDA(c, out.reserve(n ? 4 * n : 1));
    HQ_CUDA(c, c->d_flag.reserve(1));
    cudaError_t e = cudaMemsetAsync(c->d_flag.p, 0, sizeof(unsigned int), c->stream);
    const float* src[3] = {r, g, b};
    for (int pl = 0; pl < 3 && n && e == cudaSuccess; ++pl) e = cudaMemcpyAsync(in.p + (size_t)pl * n, src[pl], n * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = hq::launch_sc_unit_to_xyz4(in.p, in.p + n, in.p + 2 * n, n, out.p, c->d_flag.p, c->stream);
    unsigned int bad = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, c->d_flag.p, sizeof bad, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(xyz4, out.p, 4 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    in.release(); out.release();
    if (e != cudaSuccess) return fail(c, HQ_ERR_CUDA, "hq_rgb_to_xyz failed: %s", cudaGetErrorString(e));
    if (bad) return fail(c, HQ_ERR_INVALID, "float image values must lie in [0,1] (Icy's rescaled convertToType, HybridQuantization.java:95)");
    return HQ_OK;
} catch (const std::exception& ex) { return api_exception(c, ex); }

int hq_xyz_to_scielab(hq_ctx* c, const float* xyz4, int width, int rows, const float* illuminant3, float* lab4) try {
    if (!c || !xyz4 || !lab4 || !illuminant3) return c ? fail(c, HQ_ERR_INVALID, "NULL array") : HQ_ERR_INVALID;
    if (width < 1 || rows < 1) return fail(c, HQ_ERR_INVALID, "bad image size %d x %d", width, rows);
    int rc = bind_device(c); if (rc) return rc;
    if (c->sc_taps == 0) { rc = hq_scielab_configure(c, 72, 45.0f); if (rc) return rc; }  // plugin defaults :229-231
    const int half = c->sc_taps / 2;
    if (width < half || rows < half)
        return fail(c, HQ_ERR_UNSUPPORTED, "image %dx%d is smaller than the filter half-width %d (the reference's single reflection, "
                    "OptimizedConvolution.cl:20-27, would read out of bounds)", width, rows, half);
    if (rows > 65535) return fail(c, HQ_ERR_UNSUPPORTED, "at most 65,535 rows (got %d)", rows);
    const size_t n = (size_t)width * rows, stride = hq::plane_stride(n);
    DevBuf<float> io, opp, tmp, lab;
    cudaError_t e = io.reserve(4 * n);
    if (e == cudaSuccess) e = opp.reserve(3 * stride);
    if (e == cudaSuccess) e = tmp.reserve(7 * stride);
    if (e == cudaSuccess) e = lab.reserve(3 * stride);
    if (e == cudaSuccess) e = cudaMemcpyAsync(io.p, xyz4, 4 * n * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = hq::launch_sc_xyz4_to_opp(io.p, n, stride, opp.p, c->stream);
    if (e == cudaSuccess) e = hq::launch_sc_original(opp.p, width, rows, stride, c->d_sc_filters.p, c->sc_generic ? nullptr : c->sc_block.data(), c->sc_taps,
                                                     HQ_WHITEPOINT_D65, hq::ScRows{0, rows, 0, rows}, tmp.p, lab.p, c->stream, illuminant3);
    if (e == cudaSuccess) e = hq::launch_sc_planes_to_f4(lab.p, n, stride, io.p, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(lab4, io.p, 4 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    io.release(); opp.release(); tmp.release(); lab.release();
    if (e != cudaSuccess) return fail(c, HQ_ERR_CUDA, "hq_xyz_to_scielab failed: %s", cudaGetErrorString(e));
    return HQ_OK;
} catch (const std::exception& ex) { return api_exception(c, ex); }

namespace {
// lab4: the WHOLE image's [n][4] S-CIELAB values; context m takes the rows it holds (halo rows included)
int scielab_set_image_one(hq_ctx* m, const float* lab4_whole) {
    int rc = bind_device(m); if (rc) return rc;
    if (!m->have_image) return fail(m, HQ_ERR_NO_IMAGE, "no image: call hq_set_image_u8 or hq_set_image_f32_planar first");
    if (m->sc_taps == 0) { rc = hq_scielab_configure(m, 72, 45.0f); if (rc) return rc; }
    if (m->n == 0) { m->sc_image_ready = true; return HQ_OK; }
    const size_t first = (size_t)(m->g_row0 - m->halo_top) * m->width;
    DevBuf<float> in;
    HQ_CUDA(m, in.reserve(4 * m->n));
    HQ_CUDA(m, m->d_sc_opp.reserve(3 * m->stride));
    HQ_CUDA(m, m->d_sc_tmp.reserve(7 * m->stride));
    HQ_CUDA(m, m->d_sc_lab.reserve(3 * m->stride));
    cudaError_t e = cudaMemcpyAsync(in.p, lab4_whole + 4 * first, 4 * m->n * sizeof(float), cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess) e = hq::launch_sc_f4_to_planes(in.p, m->n, m->stride, m->d_sc_lab.p, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    in.release();
    if (e != cudaSuccess) return fail(m, HQ_ERR_CUDA, "hq_scielab_set_image failed: %s", cudaGetErrorString(e));
    m->sc_image_ready = true;
    return HQ_OK;
}
}  // namespace

int hq_scielab_set_image(hq_ctx* c, const float* lab4) try {
    if (!c || !lab4) return c ? fail(c, HQ_ERR_INVALID, "lab4 is NULL") : HQ_ERR_INVALID;
    if (!c->is_multi() && (c->halo_top || c->halo_bottom || c->g_rows != c->own_rows))
        return fail(c, HQ_ERR_UNSUPPORTED, "hq_scielab_set_image takes the whole image: not available on an explicit row shard");
    std::vector<hq_ctx*> self(1, c);
    for (hq_ctx* m : (c->is_multi() ? c->members : self)) { const int rc = member_rc(c, m, scielab_set_image_one(m, lab4)); if (rc) return rc; }
    return bind_device(c);
} catch (const std::exception& ex) { return api_exception(c, ex); }

int hq_delta_e_images(hq_ctx* c, const float* lab4_a, const float* lab4_b, size_t n, float* error_rgba4, double* mean_de) try {
    if (!c || ((!lab4_a || !lab4_b) && n)) return c ? fail(c, HQ_ERR_INVALID, "NULL array") : HQ_ERR_INVALID;
    int rc = bind_device(c); if (rc) return rc;
    DevBuf<float> a, b, e, img;
    std::vector<float> he(n);
    cudaError_t ce = a.reserve(n ? 4 * n : 1);
    if (ce == cudaSuccess) ce = b.reserve(n ? 4 * n : 1);
    if (ce == cudaSuccess) ce = e.reserve(n ? n : 1);
    if (ce == cudaSuccess && error_rgba4) ce = img.reserve(n ? 4 * n : 1);
    if (ce == cudaSuccess && n) ce = cudaMemcpyAsync(a.p, lab4_a, 4 * n * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (ce == cudaSuccess && n) ce = cudaMemcpyAsync(b.p, lab4_b, 4 * n * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (ce == cudaSuccess && n && error_rgba4) ce = cudaMemcpyAsync(img.p, error_rgba4, 4 * n * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (ce == cudaSuccess) ce = hq::launch_sc_delta_e4(a.p, b.p, n, e.p, error_rgba4 ? img.p : nullptr, c->delta_e, c->stream);
    if (ce == cudaSuccess && n) ce = cudaMemcpyAsync(he.data(), e.p, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (ce == cudaSuccess && n && error_rgba4) ce = cudaMemcpyAsync(error_rgba4, img.p, 4 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
    a.release(); b.release(); e.release(); img.release();
    if (ce != cudaSuccess) return fail(c, HQ_ERR_CUDA, "hq_delta_e_images failed: %s", cudaGetErrorString(ce));
    double sum = 0.0;   // ImageManipulation.java:886-893: the floats summed in a double, in pixel order, then / n
    for (size_t i = 0; i < n; ++i) sum += (double)he[i];
    if (mean_de) *mean_de = n ? sum / (double)n : 0.0;
    return HQ_OK;
} catch (const std::exception& ex) { return api_exception(c, ex); }

// error-image mode: HybridQuantization.errorImage (:139-182) + ImageManipulation.computeError (:858-894)
namespace {
// the second image of error-image mode, as packed u8 (rgb8) or as float planes (f32[3])
// raw_sum (optional): a member of a multi-device context — no exchange, no mean; the fixed-point sum of its own rows and the
// number of NaN pixels (CIE94 only) go to raw_sum[0], raw_sum[1]
int error_image_common(hq_ctx* c, const uint8_t* rgb8, const float* const* f32, float* error_map, uint8_t* error_map_u8, double* mean_de,
                       unsigned long long* raw_sum = nullptr) {
    int rc = bind_device(c); if (rc) return rc;
    rc = sc_ensure_image(c); if (rc) return rc;
    const size_t n = c->n;
    HQ_CUDA(c, c->d_sc_lab2.reserve(3 * c->stride));
    HQ_CUDA(c, c->d_sc_err.reserve(2));
    if (error_map) HQ_CUDA(c, c->d_sc_map.reserve(n ? n : 1));
    if (error_map_u8) HQ_CUDA(c, c->d_sc_map8.reserve(n ? n : 1));
    // S-CIELAB of the second image through the same route as the original (sRGBToScielab, ScielabProcessor.java:374-381)
    if (rgb8) {
        HQ_CUDA(c, c->d_sc_rgb2.reserve(n * 3 > 0 ? n * 3 : 1));
        HQ_CUDA(c, cudaMemcpyAsync(c->d_sc_rgb2.p, rgb8, n * 3, cudaMemcpyHostToDevice, c->stream));
        HQ_CUDA(c, hq::launch_sc_rgb_to_opp(c->d_sc_rgb2.p, n, c->stride, c->d_table.p, c->d_sc_opp.p, c->stream));
    } else {  // the planes land in the opponent buffer and are converted in place (element-wise kernel)
        HQ_CUDA(c, c->d_flag.reserve(1));
        HQ_CUDA(c, cudaMemsetAsync(c->d_flag.p, 0, sizeof(unsigned int), c->stream));
        for (int pl = 0; pl < 3 && n; ++pl)
            HQ_CUDA(c, cudaMemcpyAsync(c->d_sc_opp.p + (size_t)pl * c->stride, f32[pl], n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        HQ_CUDA(c, hq::launch_sc_unit_to_opp(c->d_sc_opp.p, n, c->stride, c->d_sc_opp.p, c->d_flag.p, c->stream));
        unsigned int bad = 0;
        HQ_CUDA(c, cudaMemcpyAsync(&bad, c->d_flag.p, sizeof bad, cudaMemcpyDeviceToHost, c->stream));
        HQ_CUDA(c, cudaStreamSynchronize(c->stream));
        if (bad) return fail(c, HQ_ERR_INVALID, "float image values must lie in [0,1] (Icy's rescaled convertToType, HybridQuantization.java:142-143)");
    }
    HQ_CUDA(c, hq::launch_sc_original(c->d_sc_opp.p, c->width, c->rows, c->stride, c->d_sc_filters.p, c->sc_generic ? nullptr : c->sc_block.data(),
                                      c->sc_taps, c->whitepoint, sc_rows(c), c->d_sc_tmp.p, c->d_sc_lab2.p, c->stream));
    HQ_CUDA(c, cudaMemsetAsync(c->d_sc_err.p, 0, 16, c->stream));
    const size_t lo = c->own_lo, no = c->own_hi - c->own_lo;  // maps and the sum cover the own rows
    HQ_CUDA(c, hq::launch_sc_error_image(c->d_sc_lab.p + lo, c->d_sc_lab2.p + lo, no, c->stride, error_map ? c->d_sc_map.p : nullptr,
                                         error_map_u8 ? c->d_sc_map8.p : nullptr, c->d_sc_err.p, c->delta_e, c->stream));
    const bool reduce = !raw_sum && reduces(c);
    if (reduce) { rc = reduce_words(c, c->d_sc_err.p, 2, c->stream); if (rc) return rc; }
    unsigned long long sums[2] = {0, 0};
    HQ_CUDA(c, cudaMemcpyAsync(sums, c->d_sc_err.p, 16, cudaMemcpyDeviceToHost, c->stream));
    const unsigned long long& sum = sums[0];
    if (error_map && no) HQ_CUDA(c, cudaMemcpyAsync(error_map, c->d_sc_map.p, no * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (error_map_u8 && no) HQ_CUDA(c, cudaMemcpyAsync(error_map_u8, c->d_sc_map8.p, no, cudaMemcpyDeviceToHost, c->stream));
    HQ_CUDA(c, cudaStreamSynchronize(c->stream));
    const double n_all = (double)c->width * (double)c->g_rows;  // with the hook the sum is over the whole image
    const double n_div = reduce ? n_all : (double)no;
    if (raw_sum) { raw_sum[0] = sums[0]; raw_sum[1] = sums[1]; }
    if (mean_de && sums[1]) *mean_de = std::nan("");   // a NaN pixel of the CIE94 branch makes the reference's mean NaN (:886-893)
    else if (mean_de) *mean_de = n_div > 0 ? ((double)(int64_t)sum * (1.0 / 16777216.0)) / n_div : 0.0;  // :893 error/errorArray.length
    return HQ_OK;
}
}  // namespace

int hq_error_image(hq_ctx* c, const uint8_t* quantized_rgb, float* error_map, uint8_t* error_map_u8, double* mean_de) try {
    if (!c || !quantized_rgb) return c ? fail(c, HQ_ERR_INVALID, "quantized_rgb is NULL") : HQ_ERR_INVALID;
    if (!c->is_multi()) return error_image_common(c, quantized_rgb, nullptr, error_map, error_map_u8, mean_de);
    long long total = 0;
    unsigned long long nans = 0;
    for (hq_ctx* m : c->members) {   // each member: its rows (+ halo) of the second image in, its rows of the maps out
        if (m->own_rows == 0) continue;
        const size_t in_at = (size_t)(m->g_row0 - m->halo_top) * m->width, out_at = (size_t)m->g_row0 * m->width;
        unsigned long long part[2] = {0, 0};
        const int rc = member_rc(c, m, error_image_common(m, quantized_rgb + in_at * 3, nullptr, error_map ? error_map + out_at : nullptr,
                                                          error_map_u8 ? error_map_u8 + out_at : nullptr, nullptr, part));
        if (rc) return rc;
        total += (long long)part[0]; nans += part[1];
    }
    const double n_all = (double)c->m_width * (double)c->m_rows;
    if (mean_de) *mean_de = nans ? std::nan("") : (n_all > 0 ? ((double)total * (1.0 / 16777216.0)) / n_all : 0.0);
    return bind_device(c);
} catch (const std::exception& ex) { return api_exception(c, ex); }

int hq_error_image_f32_planar(hq_ctx* c, const float* r, const float* g, const float* b, float* error_map, uint8_t* error_map_u8, double* mean_de) try {
    if (!c || !r || !g || !b) return c ? fail(c, HQ_ERR_INVALID, "an image plane is NULL") : HQ_ERR_INVALID;
    const float* planes[3] = {r, g, b};
    if (!c->is_multi()) return error_image_common(c, nullptr, planes, error_map, error_map_u8, mean_de);
    long long total = 0;
    unsigned long long nans = 0;
    for (hq_ctx* m : c->members) {
        if (m->own_rows == 0) continue;
        const size_t in_at = (size_t)(m->g_row0 - m->halo_top) * m->width, out_at = (size_t)m->g_row0 * m->width;
        const float* mp[3] = {r + in_at, g + in_at, b + in_at};
        unsigned long long part[2] = {0, 0};
        const int rc = member_rc(c, m, error_image_common(m, nullptr, mp, error_map ? error_map + out_at : nullptr,
                                                          error_map_u8 ? error_map_u8 + out_at : nullptr, nullptr, part));
        if (rc) return rc;
        total += (long long)part[0]; nans += part[1];
    }
    const double n_all = (double)c->m_width * (double)c->m_rows;
    if (mean_de) *mean_de = nans ? std::nan("") : (n_all > 0 ? ((double)total * (1.0 / 16777216.0)) / n_all : 0.0);
    return bind_device(c);
} catch (const std::exception& ex) { return api_exception(c, ex); }

namespace {
// one context's share of a reference-faithful evaluation, enqueued on m->stream: result words (error in word 0, counts) of its
// own rows in m->d_results.  h_pal: pinned, portable.
// direct: palettes read from the pinned host copy by the kernels and no result shuffle at the end (the caller exports with sc_export_kernel)
int sc_eval_enqueue(hq_ctx* m, const float* h_pal, int B, int K, int space, bool direct = false) {
    int rc = bind_device(m); if (rc) return rc;
    const size_t npal = (size_t)B * K * 4;
    const int words = hq::result_words(K, false);
    const size_t nwords = (size_t)B * words;
    const bool idx16 = K > 256;
    const bool de94 = m->delta_e != HQ_DELTAE_CIE76;   // then B more words follow the result words: NaN pixels per candidate
    HQ_CUDA(m, m->d_results.reserve(nwords + B));
    if (de94) HQ_CUDA(m, cudaMemsetAsync(m->d_results.p + nwords, 0, (size_t)B * 8, m->stream));
    if (m->own_rows == 0) {   // an empty row block of a multi-device context contributes zeros
        HQ_CUDA(m, cudaMemsetAsync(m->d_results.p, 0, nwords * 8, m->stream));
        return HQ_OK;
    }
    rc = sc_ensure_image(m); if (rc) return rc;
    HQ_CUDA(m, m->d_pal.reserve(npal));
    HQ_CUDA(m, m->d_sc_err.reserve(B));
    HQ_CUDA(m, m->d_sc_tab.reserve((size_t)B * K));
    // Candidates go through assignment and filter stage in SUB-BATCHES whose index images stay in L2 (126 MB): with all 16
    // candidates of a 4K population in flight (133 MB of indices) the pruned kernel's scattered 1-byte index stores turned into
    // partial-sector DRAM writes — 0.338 ms per candidate against 0.284 ms at four (profiles/r02).
    const size_t idx_bytes = (m->stride ? m->stride : 1) * (idx16 ? 2 : 1);
    const bool prune_idx = K > HQ_MAX_COLORS || m->prune_mode == HQ_PRUNE_ON || (m->prune_mode == HQ_PRUNE_AUTO && K >= 32 && m->n >= 65536);
    int bs = prune_idx ? (int)(((size_t)40 << 20) / idx_bytes) : B;   // (the exhaustive kernel stores its indices coalesced and likes large batches)
    bs = bs < 1 ? 1 : (bs > B ? B : bs);
    HQ_CUDA(m, m->d_idx.reserve((size_t)bs * idx_bytes));
    const float* pal_src = direct ? h_pal : m->d_pal.p;   // pinned host memory is device-accessible (UVA): no copy node for a few KB
    if (!direct) HQ_CUDA(m, cudaMemcpyAsync(m->d_pal.p, h_pal, npal * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    // the K opponent colours each quantised image is made of (cl:194-198); the same launch clears the error sums
    HQ_CUDA(m, hq::launch_sc_palette_opp(pal_src, B * K, m->d_sc_tab.p, m->stream, m->d_sc_err.p, B));
    bool tmp_reserved = false;
    for (int b0 = 0; b0 < B; b0 += bs) {
        const int nb = B - b0 < bs ? B - b0 : bs;
        // 1. assignment (quantizeAndConvertToOpp's argmin, cl:178-193): indices + counts for every candidate of the sub-batch
        //    (exact pruned kernel where it pays: same indices and counts)
        rc = eval_device(m, pal_src + (size_t)b0 * K * 4, nb, K, space, prune_idx ? HQ_EVAL_PRUNE : 0, m->d_results.p + (size_t)b0 * words, m->d_idx.p, m->stream);
        if (rc) return rc;
        // 2. per candidate: separable filters, Opp2LAB, CIE76 against the original, fixed-point sum
        if (m->profiling) HQ_CUDA(m, cudaEventRecord(m->ev4, m->stream));
        cudaError_t fe = (m->sc_generic || m->sc_unfused || de94) ? cudaErrorNotSupported   // (the fused kernel is CIE76 only)
                         : hq::launch_sc_candidates_fused(m->d_idx.p, idx16, m->d_sc_tab.p + (size_t)b0 * K, K, nb, m->width, m->rows, m->stride, m->sc_block.data(),
                                                          m->sc_taps, m->whitepoint, sc_rows(m), m->d_sc_lab.p, m->d_sc_err.p + b0, m->sm_count, m->stream);
        if (fe != cudaSuccess && fe != cudaErrorNotSupported) return fail(m, HQ_ERR_CUDA, "fused S-CIELAB kernel launch failed: %s", cudaGetErrorString(fe));
        if (fe == cudaErrorNotSupported && !tmp_reserved) { HQ_CUDA(m, m->d_sc_tmp.reserve(7 * m->stride)); tmp_reserved = true; }
        for (int b = 0; b < nb && fe == cudaErrorNotSupported; ++b) {   // other tap counts, K > 1024: two kernels per candidate
            const uint8_t* idx_b = m->d_idx.p + (size_t)b * idx_bytes;
            HQ_CUDA(m, hq::launch_sc_candidate(idx_b, idx16, m->d_sc_tab.p + (size_t)(b0 + b) * K, m->width, m->rows, m->stride, m->d_sc_filters.p,
                                               m->sc_generic ? nullptr : m->sc_block.data(), m->sc_taps, m->whitepoint, sc_rows(m), m->d_sc_tmp.p,
                                               m->d_sc_lab.p, m->d_sc_err.p + b0 + b, m->stream, m->delta_e, m->d_results.p + nwords + b0 + b));
        }
        if (m->profiling) { HQ_CUDA(m, cudaEventRecord(m->ev5, m->stream)); m->ev_sc_valid = true; m->ev_sc_candidates = nb; }
    }
    // word 0 of every candidate <- the S-CIELAB error sum, so that one all-reduce covers error and counts
    if (!direct) HQ_CUDA(m, cudaMemcpy2DAsync(m->d_results.p, (size_t)words * 8, m->d_sc_err.p, 8, 8, (size_t)B, cudaMemcpyDeviceToDevice, m->stream));
    return HQ_OK;
}
}  // namespace

int hq_eval_palettes_scielab(hq_ctx* c, const float* palettes, int B, int K, int space, int64_t* err_fx, uint64_t* counts) try {
    int rc = check_eval_args(c, B, K, space); if (rc) return rc;
    if (!palettes) return fail(c, HQ_ERR_INVALID, "palettes is NULL");
    rc = bind_device(c); if (rc) return rc;
    const size_t npal = (size_t)B * K * 4;
    const int words = hq::result_words(K, false);
    const size_t nwords = (size_t)B * words;
    HQ_CUDA(c, c->h_pal.reserve(npal));
    HQ_CUDA(c, c->h_results.reserve(nwords + B));
    if (!copy_palettes_checked(c->h_pal.p, palettes, npal)) return fail(c, HQ_ERR_INVALID, "%s", kBadPalette);
    std::vector<hq_ctx*> self(1, c);
    const std::vector<hq_ctx*>& targets = c->is_multi() ? c->members : self;
    const size_t tail = c->delta_e != HQ_DELTAE_CIE76 ? (size_t)B : 0;   // NaN pixel counts of the CIE94 branch, reduced with the rest
    for (hq_ctx* m : targets) m->delta_e = c->delta_e;
    // one device, a few KB each way (the plugin's default search): no copy nodes — the kernels read the pinned palettes, a one-CTA
    // kernel merges the error sums into the result words, writes them to pinned host memory and raises the flag the host spins on
    const bool direct = c->direct_io && !c->is_multi() && !reduces(c) && tail == 0 && !c->profiling && c->own_rows > 0 &&
                        npal * sizeof(float) <= 65536 && nwords * 8 <= 32768;
    for (size_t i = targets.size(); i-- > 0;) { rc = member_rc(c, targets[i], sc_eval_enqueue(targets[i], c->h_pal.p, B, K, space, direct)); if (rc) return rc; }
    if (reduces(c)) { rc = eval_reduce(c, nwords + tail); if (rc) return rc; }
    rc = bind_device(c); if (rc) return rc;
    if (direct) {
        const unsigned long long seq = ++c->export_seq;
        HQ_CUDA(c, hq::launch_sc_export(c->d_results.p, c->d_sc_err.p, words, nwords, c->h_results.p, c->h_flag.p, seq, c->stream));
        HQ_CUDA(c, wait_flag(c->h_flag.p, seq, c->stream));
    } else {
        HQ_CUDA(c, cudaMemcpyAsync(c->h_results.p, c->d_results.p, (nwords + tail) * 8, cudaMemcpyDeviceToHost, c->stream));
        HQ_CUDA(c, wait_stream(c->stream));
    }
    if (reduces(c)) { rc = peer_check(c); if (rc) return rc; }
    for (int b = 0; b < B; ++b) {
        if (err_fx) err_fx[b] = (tail && c->h_results.p[nwords + b]) ? HQ_ERR_FX_NAN : (int64_t)c->h_results.p[(size_t)b * words];
        if (counts) std::memcpy(counts + (size_t)b * K, c->h_results.p + (size_t)b * words + 1, sizeof(uint64_t) * K);
    }
    return HQ_OK;
} catch (const std::exception& ex) { return api_exception(c, ex); }

int hq_set_allreduce(hq_ctx* c, hq_allreduce_fn fn, void* user) {
    if (!c) return HQ_ERR_INVALID;
    if (c->is_multi() && fn) return fail(c, HQ_ERR_UNSUPPORTED, "a multi-device context reduces over its own NCCL communicator");
    c->allreduce = fn;
    c->allreduce_user = user;
    return HQ_OK;
}

void hq_swasa_default_params(hq_swasa_params* p) {
    if (!p) return;
    const hq::HybridQuantization d;
    p->population = d.populationSize; p->imax = d.imax; p->iTc = d.iTc; p->delta = d.delta;
    p->convergence = d.convEnable ? 1 : 0; p->conv_delay = d.convDelay; p->conv_spread = d.convSpread;
    p->t0 = d.T0; p->alpha = d.alpha; p->s0 = d.s0; p->beta = d.beta; p->space = d.space; p->seed = d.seed;
    p->cost_model = d.costModel;
}

int hq_find_best_quantization(hq_ctx* c, int K, const hq_swasa_params* p, uint64_t n_total, float* best_colors,
                              double* best_error, double* trace_costs, int* iterations_done) {
    if (!c || !p || !best_colors) return c ? fail(c, HQ_ERR_INVALID, "NULL argument") : HQ_ERR_INVALID;
    int rc = check_eval_args(c, p->population, K, p->space); if (rc) return rc;
    if (p->imax < 1 || p->iTc < 1) return fail(c, HQ_ERR_INVALID, "imax and iTc must be >= 1");
    c->stop.store(false);
    c->stop_flag_view = false;
    try {
        hq::ImageManipulation backend(c, false, p->convergence != 0);
        backend.setStopFlag(&c->stop_flag_view);
        backend.setCostModel(p->cost_model);
        backend.setProgress(c->progress, c->progress_user);
        const int eval_flags = hq_search_eval_flags(c, K, p->space, p->cost_model);
        backend.setEvalFlags(eval_flags);  // exact pruning where it pays: same costs, same trajectory
        hq::JavaRandom random(p->seed);
        hq::SWASA swasa(p->population, p->imax, p->iTc, p->delta, p->conv_delay, p->conv_spread, p->t0, p->alpha, p->s0, p->beta, &random);
        double err = 0;
        // a small search with the LAB cost (the plugin's defaults: 8 colours x 4 candidates) is thousands of evaluations whose cost is
        // the kernel launch: with HQ_PERSIST=1 one persistent kernel serves them all through a mailbox (same integers; hq_eval_palettes
        // falls back to one launch per evaluation whenever the kernel is not there).  Off by default: measured 18.3 against 20.0 us
        // per iteration — the mailbox hops (PCIe poll, palette fetch, republish) cost almost what the launch did.
        if (p->cost_model == HQ_COST_LAB && eval_flags == 0 && p->imax >= 16) persist_begin(c, p->population, K, p->space, false);
        std::vector<float> best;
        try {
            best = backend.findBestQuantization(K, swasa, n_total, p->space, &err, trace_costs, iterations_done);
        } catch (...) {
            persist_end(c);
            throw;
        }
        persist_end(c);
        std::memcpy(best_colors, best.data(), sizeof(float) * best.size());
        if (best_error) *best_error = err;
    } catch (const std::exception& ex) {
        if (c->err.empty()) c->err = ex.what();
        return HQ_ERR_CUDA;
    }
    return bind_device(c);
}

int hq_search_eval_flags(const hq_ctx* c, int K, int space, int cost_model) {
    if (!c || !c->have_image) return 0;
    if (space != HQ_SPACE_LAB || cost_model != HQ_COST_LAB) return 0;
    // measured crossover (4 candidates, profiles/r01/sweep_v10.json + DESIGN.md section 6): from K = 32 on every image of
    // >= 256 x 256, from K = 12 on images of >= 1024 x 768
    const size_t own = c->own_hi - c->own_lo;
    const bool pays = ((K >= 32 && own >= 65536) || (K >= 12 && own >= 786432)) && own < 0xffffffffull;
    return (K > HQ_MAX_COLORS || c->prune_mode == HQ_PRUNE_ON || (c->prune_mode == HQ_PRUNE_AUTO && pays)) ? HQ_EVAL_PRUNE : 0;
}

int hq_set_graphs(hq_ctx* c, int enabled) {
    if (!c) return HQ_ERR_INVALID;
    c->use_graphs = enabled != 0;
    if (!c->use_graphs && c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; c->graph_key = hq_ctx::EvalKey(); }
    return HQ_OK;
}

int hq_set_delta_e(hq_ctx* c, int type) {
    if (!c) return HQ_ERR_INVALID;
    if (type == HQ_DELTAE_CIEDE2000)
        return fail(c, HQ_ERR_UNSUPPORTED, "CIEDE2000 has no definition to be faithful to: the reference's branch is an empty stub (OptimizedConvolution.cl:227-229)");
    if (type != HQ_DELTAE_CIE76 && type != HQ_DELTAE_CIE94) return fail(c, HQ_ERR_INVALID, "unknown dE type %d", type);
    c->delta_e = type;
    for (hq_ctx* m : c->members) m->delta_e = type;
    return HQ_OK;
}

int hq_set_pruning(hq_ctx* c, int mode) {
    if (!c) return HQ_ERR_INVALID;
    if (mode != HQ_PRUNE_OFF && mode != HQ_PRUNE_AUTO && mode != HQ_PRUNE_ON) return fail(c, HQ_ERR_INVALID, "unknown pruning mode %d", mode);
    c->prune_mode = mode;
    for (hq_ctx* m : c->members) m->prune_mode = mode;
    return HQ_OK;
}

int hq_pruning_stats(hq_ctx* c, uint32_t* chunks, double* mean_survivors) {
    if (!c) return HQ_ERR_INVALID;
    if (!c->have_image) return fail(c, HQ_ERR_NO_IMAGE, "no image");
    int rc = bind_device(c); if (rc) return rc;
    rc = ensure_pruned(c, c->pr_own, HQ_SPACE_LAB, c->own_lo, c->own_hi, false, c->stream); if (rc) return rc;
    if (chunks) *chunks = c->pr_own.nchunks;
    if (mean_survivors) {
        unsigned long long st[2] = {0, 0};
        HQ_CUDA(c, cudaMemcpyAsync(st, c->d_pr_stats.p, sizeof st, cudaMemcpyDeviceToHost, c->stream));
        HQ_CUDA(c, cudaStreamSynchronize(c->stream));
        *mean_survivors = st[1] ? (double)st[0] / (double)st[1] : 0.0;
    }
    return HQ_OK;
}

int hq_set_progress(hq_ctx* c, hq_progress_fn fn, void* user) {
    if (!c) return HQ_ERR_INVALID;
    c->progress = fn;
    c->progress_user = user;
    return HQ_OK;
}

void hq_request_stop(hq_ctx* c) {
    if (!c) return;
    c->stop.store(true);
    c->stop_flag_view = true;
}

void hq_java_random_seed(hq_java_random* r, int64_t seed) { hq::JavaRandom j(seed); r->state = j.state(); }
int32_t hq_java_random_next(hq_java_random* r, int bits) { hq::JavaRandom j; j.setState(r->state); const int32_t v = j.next(bits); r->state = j.state(); return v; }
float hq_java_random_next_float(hq_java_random* r) { hq::JavaRandom j; j.setState(r->state); const float v = j.nextFloat(); r->state = j.state(); return v; }
double hq_java_random_next_double(hq_java_random* r) { hq::JavaRandom j; j.setState(r->state); const double v = j.nextDouble(); r->state = j.state(); return v; }

static hq::SWASA make_swasa(const hq_swasa_params* p, hq::JavaRandom* j) {
    return hq::SWASA(p->population, p->imax, p->iTc, p->delta, p->conv_delay, p->conv_spread, p->t0, p->alpha, p->s0, p->beta, j);
}
void hq_swasa_generate_random_colors(hq_java_random* r, int K, float* colors) {
    hq_swasa_params d; hq_swasa_default_params(&d);
    hq::JavaRandom j; j.setState(r->state);
    make_swasa(&d, &j).generateRandomColors(K, colors);
    r->state = j.state();
}
void hq_swasa_generate_neighboring_colors(const hq_swasa_params* p, hq_java_random* r, const float* colors, float* next_colors, int K, int iteration) {
    hq::JavaRandom j; j.setState(r->state);
    make_swasa(p, &j).generateNeighboringColors(colors, next_colors, K, iteration);
    r->state = j.state();
}
float hq_swasa_max_step_width(const hq_swasa_params* p, int iteration) {
    hq::JavaRandom j;
    return make_swasa(p, &j).maxStepWidth(iteration);
}

int hq_set_profiling(hq_ctx* c, int enabled) {
    if (!c) return HQ_ERR_INVALID;
    int rc = bind_device(c); if (rc) return rc;
    if (enabled && !c->ev0) {
        HQ_CUDA(c, cudaEventCreate(&c->ev0)); HQ_CUDA(c, cudaEventCreate(&c->ev1));
        HQ_CUDA(c, cudaEventCreate(&c->ev2)); HQ_CUDA(c, cudaEventCreate(&c->ev3));
        HQ_CUDA(c, cudaEventCreate(&c->ev4)); HQ_CUDA(c, cudaEventCreate(&c->ev5));
    }
    c->profiling = enabled != 0;
    c->ev_valid = false;
    c->ev_rl_valid = false;
    c->ev_sc_valid = false;
    return HQ_OK;
}

int hq_last_assign_ms(hq_ctx* c, float* ms) {
    if (!c || !ms) return HQ_ERR_INVALID;
    if (!c->ev_valid) return fail(c, HQ_ERR_INVALID, "no profiled evaluation yet (hq_set_profiling)");
    int rc = bind_device(c); if (rc) return rc;
    HQ_CUDA(c, cudaEventSynchronize(c->ev1));
    HQ_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return HQ_OK;
}

int hq_last_rgb_to_lab_ms(hq_ctx* c, float* ms) {
    if (!c || !ms) return HQ_ERR_INVALID;
    if (!c->ev_rl_valid) return fail(c, HQ_ERR_INVALID, "no profiled image conversion yet (hq_set_profiling)");
    int rc = bind_device(c); if (rc) return rc;
    HQ_CUDA(c, cudaEventSynchronize(c->ev3));
    HQ_CUDA(c, cudaEventElapsedTime(ms, c->ev2, c->ev3));
    return HQ_OK;
}

int hq_last_scielab_stage_ms(hq_ctx* c, float* ms, int* candidates) {
    if (!c || !ms) return HQ_ERR_INVALID;
    if (!c->ev_sc_valid) return fail(c, HQ_ERR_INVALID, "no profiled S-CIELAB evaluation yet (hq_set_profiling)");
    int rc = bind_device(c); if (rc) return rc;
    HQ_CUDA(c, cudaEventSynchronize(c->ev5));
    HQ_CUDA(c, cudaEventElapsedTime(ms, c->ev4, c->ev5));
    if (candidates) *candidates = c->ev_sc_candidates;
    return HQ_OK;
}

int hq_measure_fp32_peak(hq_ctx* c, double* tflops_ffma, double* tflops_ffma2) {
    if (!c) return HQ_ERR_INVALID;
    int rc = bind_device(c); if (rc) return rc;
    DevBuf<float> d;
    HQ_CUDA(c, d.reserve(1));
    cudaEvent_t e0, e1;
    HQ_CUDA(c, cudaEventCreate(&e0));
    HQ_CUDA(c, cudaEventCreate(&e1));
    const int iters = 40000;  // 40000 * 32 FMA * 1024 threads/SM: ~20 ms per launch
    double best[2] = {0, 0};
    for (int packed = 0; packed < 2; ++packed)
        for (int rep = 0; rep < 4; ++rep) {
            HQ_CUDA(c, cudaEventRecord(e0, c->stream));
            HQ_CUDA(c, hq::launch_fp32_peak(packed != 0, iters, c->sm_count, d.p, c->stream));
            HQ_CUDA(c, cudaEventRecord(e1, c->stream));
            HQ_CUDA(c, cudaEventSynchronize(e1));
            float ms = 0;
            HQ_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
            const double flop = (double)iters * 32.0 * (packed ? 4.0 : 2.0) * 256.0 * 4.0 * c->sm_count;
            const double tf = flop / (ms * 1e-3) / 1e12;
            if (rep > 0 && tf > best[packed]) best[packed] = tf;
        }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    d.release();
    if (tflops_ffma) *tflops_ffma = best[0];
    if (tflops_ffma2) *tflops_ffma2 = best[1];
    return HQ_OK;
}

int hq_host_math_range(int which, uint32_t first_bits, uint32_t count, float* out, int threads) {
    if (!out || which < 0 || which > 7) return HQ_ERR_INVALID;
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([=] {
            const uint64_t lo = (uint64_t)count * t / threads, hi = (uint64_t)count * (t + 1) / threads;
            for (uint64_t i = lo; i < hi; ++i) {
                const float v = HQ_U2F(first_bits + (uint32_t)i);
                out[i] = hq_math_probe(which, v);
            }
        });
    for (auto& th : pool) th.join();
    return HQ_OK;
}

int hq_device_math_range(hq_ctx* c, int which, uint32_t first_bits, uint32_t count, float* out) {
    if (!c || !out || which < 0 || which > 7) return HQ_ERR_INVALID;
    int rc = bind_device(c); if (rc) return rc;
    DevBuf<float> d;
    HQ_CUDA(c, d.reserve(count ? count : 1));
    cudaError_t e = hq::launch_math_probe(which, first_bits, count, d.p, c->stream);
    if (e == cudaSuccess && count) e = cudaMemcpyAsync(out, d.p, (size_t)count * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    d.release();
    if (e != cudaSuccess) return fail(c, HQ_ERR_CUDA, "math probe failed: %s", cudaGetErrorString(e));
    return HQ_OK;
}

void hq_host_srgb_to_lab(const float rgb[3], int whitepoint, float lab[3]) {
    const hq_float3 v = hq_srgb_to_lab(rgb[0], rgb[1], rgb[2], hq_make_white(whitepoint));
    lab[0] = v.x; lab[1] = v.y; lab[2] = v.z;
}

}  // extern "C"
