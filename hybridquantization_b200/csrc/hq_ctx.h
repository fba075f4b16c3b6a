// hq_ctx.h — the context object behind the C ABI and the small helpers every translation unit of the library shares
// (hq_api.cu: single-device entries; hq_multi.cu: native NCCL communicator + single-process multi-device contexts).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hq_b200.h"
#include "../../include/hq_plugin.hpp"
#include "hq_kernels.cuh"
#include "hq_math.h"


namespace hqi {


extern thread_local std::string g_create_error;  // hq_api.cu

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t count) {
        if (count <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
        if (e == cudaSuccess) cap = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t count) {
        if (count <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&p), count * sizeof(T), cudaHostAllocPortable);  // every device of a multi-device context reads it
        if (e == cudaSuccess) cap = count;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace hqi
using hqi::DevBuf;
using hqi::PinBuf;

struct hq_ctx {
    int device = 0;
    int sm_count = 0;
    int clock_khz = 0;
    char name[128] = {0};
    cudaStream_t stream = nullptr;
    std::string err;

    // image state (this rank's shard)
    size_t n = 0, stride = 0;
    int width = 0, rows = 0, whitepoint = 0;   // rows = local rows INCLUDING halo rows
    int halo_top = 0, halo_bottom = 0, own_rows = 0;  // rows [halo_top, halo_top + own_rows) are this rank's own
    int g_row0 = 0, g_rows = 0;                // global index of the first own row, global image height
    size_t own_lo = 0, own_hi = 0;             // own pixel range inside the local arrays
    bool have_image = false, have_unit = false;
    bool image_f32 = false;                    // the resident image is the planar float one (d_unit); d_rgb is not used then
    DevBuf<unsigned int> d_flag;               // out-of-range report of the float conversion
    DevBuf<uint8_t> d_rgb;
    DevBuf<float> d_lab, d_unit, d_table;

    // evaluation scratch
    DevBuf<float> d_pal;
    DevBuf<float4> d_pal_lab, d_pal_rgb;
    DevBuf<unsigned long long> d_results;
    DevBuf<uint8_t> d_idx;
    DevBuf<uint8_t> d_out_rgb;
    DevBuf<float> d_out_f32;
    PinBuf<float> h_pal;
    PinBuf<unsigned long long> h_results;
    PinBuf<unsigned long long> h_flag;      // sequence number written by export_results_kernel after the result words
    unsigned long long export_seq = 0;
    DevBuf<unsigned> d_export_counter;      // ticket counter of the scoring kernels' export tail (zero between launches)
    DevBuf<unsigned long long> d_results_small;   // result words of the one-launch evaluations: zero between launches (the last CTA re-zeroes them)
    bool small_eval = true;                 // HQ_SMALL_EVAL=0: never take the one-launch path
    // persistent evaluator of a small search (hq_kernels.cu, assign_persist_kernel): alive between persist_begin and persist_end
    // inside hq_find_best_quantization
    bool persist_enabled = false;           // HQ_PERSIST=1 (off by default: -9 % measured, a kernel that owns the GPU for the length of a search)
    bool persist_on = false;
    int persist_B = 0, persist_K = 0, persist_space = 0;
    bool persist_sums = false;
    unsigned long long persist_image_gen = 0;
    PinBuf<unsigned long long> h_persist;   // [0] command word the kernel polls, [8] exit word it writes (separate cache lines)
    DevBuf<unsigned long long> d_persist_cmd;
    bool direct_io = true;                  // HQ_DIRECT_IO=0: the H2D copy / D2H copy / stream wait path instead (A/B measurements)

    // exact pruning (hq_pruned.cu): cell-sorted copy of the own pixels, chunk table, boxes; built on first use per image
    int prune_mode = HQ_PRUNE_AUTO;
    struct PrunedSet {   // one cell-sorted copy of (a range of) the resident image
        bool ready = false;
        int space = -1;
        size_t sstride = 0;
        unsigned nchunks = 0;
        DevBuf<float> sorted, box;
        DevBuf<unsigned> perm, chunk_start, chunk_len;
        void release() { sorted.release(); box.release(); perm.release(); chunk_start.release(); chunk_len.release(); ready = false; }
    };
    PrunedSet pr_own;   // CIELAB features of the OWN pixels: the LAB cost model (error, counts, sums; no indices)
    PrunedSet pr_all;   // features of EVERY local pixel (own + halo) in the space asked for, with their image positions:
                        // index-producing evaluations (the S-CIELAB chain, hq_quantize)
    DevBuf<float> d_big_d2;                 // palettes beyond the staged kernels (hq_bigk.cu): per-pixel running best distance ...
    DevBuf<unsigned> d_big_idx;             // ... and index of the candidate being swept (the final assignment afterwards)
    DevBuf<unsigned> d_pr_scratch;
    DevBuf<unsigned long long> d_pr_stats;
    PinBuf<unsigned long long> h_pr_small;

    int delta_e = HQ_DELTAE_CIE76;            // ImageManipulation.deltaETypes (:20): what a pixel pair is scored with (hq_set_delta_e)

    // S-CIELAB stage (next row 1)
    std::vector<float> sc_filters7, sc_abs3;  // [7][taps], [taps] as ScielabProcessor builds them
    std::vector<float> sc_block;              // [8][taps] device layout, host copy
    bool sc_generic = false;                  // test hook: force the generic (any-taps) kernels
    bool sc_unfused = false;                  // HQ_SC_UNFUSED=1: the two-kernel 21-tap candidate path of round 1 (A/B measurements, tests)
    int sc_taps = 0;
    bool sc_image_ready = false;
    DevBuf<float> d_sc_filters, d_sc_opp, d_sc_tmp, d_sc_lab, d_sc_lab2, d_sc_map;
    DevBuf<uint8_t> d_sc_rgb2, d_sc_map8;
    DevBuf<float4> d_sc_tab;
    DevBuf<unsigned long long> d_sc_err;

    // CUDA graph of one host-buffer evaluation (H2D palettes, palette kernel, scoring kernel, D2H results): a search repeats
    // the same launch set thousands of times; for small images the per-launch driver cost dominated an iteration
    struct EvalKey {
        int B = 0, K = 0, space = 0, flags = 0; unsigned long long image_gen = 0;
        const void *d_pal = nullptr, *d_results = nullptr, *h_pal = nullptr, *h_results = nullptr, *d_pal_lab = nullptr, *d_pal_rgb = nullptr;
        bool operator==(const EvalKey& o) const {
            return B == o.B && K == o.K && space == o.space && flags == o.flags && image_gen == o.image_gen && d_pal == o.d_pal &&
                   d_results == o.d_results && h_pal == o.h_pal && h_results == o.h_results && d_pal_lab == o.d_pal_lab && d_pal_rgb == o.d_pal_rgb;
        }
    };
    EvalKey graph_key, seen_key;
    cudaGraphExec_t graph_exec = nullptr;
    unsigned long long image_gen = 0;
    bool use_graphs = false;  // off by default: see hq_set_graphs

    // an image converted on a caller's stream (hq_set_image_u8_device): every other stream waits on this event before it
    // touches the image buffers
    cudaEvent_t ev_image = nullptr;
    bool image_foreign = false;             // the resident image was converted on a stream other than c->stream

    // ---- native NCCL (hq_multi.cu).  `comm` is this context's communicator: rank `comm_rank` of `comm_size`, either one
    // process per GPU (hq_comm_init_rank) or one member of a single-process multi-device context (hq_create_multi)
    void* comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    // single-process multi-device: the LEADER (the context the caller holds) lists every member, itself first
    std::vector<hq_ctx*> members;
    hq_ctx* leader = nullptr;               // set on the members the leader owns
    int m_width = 0, m_rows = 0;            // whole image of a multi-device context
    bool is_multi() const { return members.size() > 1; }
    // ---- the same exchange over NVLink peer memory for small payloads (hq_kernels.cuh, PeerExchange): this context's mailbox
    // and every other rank's, mapped (peer access inside one process, CUDA IPC between processes)
    DevBuf<unsigned long long> d_peer_box;
    unsigned long long* peer_box[hq::kPeerMaxRanks] = {};
    bool peer_ipc[hq::kPeerMaxRanks] = {};    // opened with cudaIpcOpenMemHandle: closed in comm_release
    bool peer_open = false;
    unsigned long long peer_seq = 0;          // exchanges done; in lockstep on every rank
    unsigned long long peer_timeout_ns = 60ull * 1000000000ull;   // HQ_PEER_TIMEOUT_MS
    PinBuf<unsigned long long> h_peer_status; // one word: sequence number of an exchange that timed out

    hq_progress_fn progress = nullptr;
    void* progress_user = nullptr;
    hq_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    bool profiling = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev4 = nullptr, ev5 = nullptr;
    bool ev_valid = false, ev_rl_valid = false, ev_sc_valid = false;
    int ev_sc_candidates = 0;               // candidates of the filter-stage launch between ev4 and ev5
    std::atomic<bool> stop{false};
    volatile bool stop_flag_view = false;
};

namespace hqi {

inline int fail(hq_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define HQ_CUDA(c, call)                                                                   \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return fail((c), HQ_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

// Completion wait of the per-iteration calls (hq_eval_palettes*): poll the stream for a few milliseconds before falling
// back to cudaStreamSynchronize.  The default synchronisation may yield the CPU, and in a process with other busy threads
// (a Python host with a thread pool, a JVM) the wake-up then costs as much as a whole pruned scoring step — measured: a
// 1080p search of 1,000 iterations took 0.18 s instead of 0.09 s, a 4K / 64-candidate one 0.67 s instead of 0.37 s.
inline cudaError_t wait_stream(cudaStream_t st) {
    using clock = std::chrono::steady_clock;
    const auto t0 = clock::now();
    for (;;) {
        for (int i = 0; i < 64; ++i) {
            const cudaError_t e = cudaStreamQuery(st);
            if (e != cudaErrorNotReady) return e;
        }
        if (clock::now() - t0 > std::chrono::milliseconds(8)) return cudaStreamSynchronize(st);
    }
}

// Spin on the sequence number export_results_kernel writes into pinned host memory after the result words; the stream is
// queried now and then so that a failed launch surfaces as an error instead of a hang.
inline cudaError_t wait_flag(const unsigned long long* flag, unsigned long long seq, cudaStream_t st) {
    using clock = std::chrono::steady_clock;
    const volatile unsigned long long* f = flag;
    const auto t0 = clock::now();
    for (;;) {
        for (int i = 0; i < 2048; ++i)
            if (*f == seq) { std::atomic_thread_fence(std::memory_order_acquire); return cudaSuccess; }  // result words are read after this
        cudaError_t e = cudaStreamQuery(st);
        if (e == cudaErrorNotReady && clock::now() - t0 > std::chrono::milliseconds(8)) e = cudaStreamSynchronize(st);  // long kernel: stop burning a core
        if (e == cudaSuccess) return *f == seq ? cudaSuccess : cudaErrorUnknown;
        if (e != cudaErrorNotReady) return e;
    }
}

// no C++ exception (std::bad_alloc from a host-side vector, std::runtime_error from the host classes) may cross the C ABI
inline int api_exception(hq_ctx* c, const std::exception& ex) {
    if (c) c->err = std::string("internal error: ") + ex.what();
    return HQ_ERR_CUDA;
}

inline int bind_device(hq_ctx* c) {
    HQ_CUDA(c, cudaSetDevice(c->device));
    return HQ_OK;
}


// ---- hq_multi.cu
bool reduces(const hq_ctx* c);   // an all-reduce follows every evaluation (hook, communicator or multi-device context)
int reduce_words(hq_ctx* c, unsigned long long* d_words, size_t n_words, cudaStream_t st);          // one rank per process
int group_reduce(hq_ctx* leader, const std::vector<unsigned long long*>& bufs, size_t n_words);     // one process, all members
void comm_release(hq_ctx* c);
bool peer_ready(const hq_ctx* c, size_t n_words);   // this exchange goes over peer memory (the same answer on every rank)
hq::PeerExchange peer_next(hq_ctx* c);              // parameters of the context's next exchange (advances its sequence)
int peer_check(hq_ctx* c);                          // after a completed call: a timed-out exchange is an error, not a wrong total
}  // namespace hqi
