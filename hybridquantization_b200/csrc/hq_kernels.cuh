// hq_kernels.cuh — launch interface of the sm_100a kernels (implemented in hq_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace hq {

constexpr int kMaxColors = 1024;  // palette sizes the fused kernel stages in shared memory
constexpr int kPadColors = 8;     // palettes are padded to a multiple of this with far colours

inline int padded_colors(int K) { return (K + kPadColors - 1) / kPadColors * kPadColors; }

// result words (8 bytes each) per candidate: [0] sum of error (2^-24 fixed point, int64),
// [1 .. K] per-colour pixel counts (u64), then if sums are requested
// [1+K .. 1+4K) per-colour (sum L, sum a, sum b) in 2^-24 fixed point (int64).
inline int result_words(int K, bool want_sums) { return 1 + K + (want_sums ? 3 * K : 0); }

// plane stride (floats / index elements) for n pixels: rows of every SoA plane start 128-B aligned
inline size_t plane_stride(size_t n) { return (n + 31) / 32 * 32; }

// packed u8 RGB -> fp32 planes.  lab = [3][stride] (L, a, b); unit = [3][stride] (r, g, b)/255 or null.
// d_table: 512 floats built once by launch_decode_table (u8 -> unit, u8 -> linear light)
cudaError_t launch_decode_table(float* d_table, cudaStream_t stream);
// d_src[nwords] -> pinned host memory (device-accessible under UVA), then *h_flag = seq (system-scope ordered after the words)
cudaError_t launch_export_results(const unsigned long long* d_src, unsigned long long* h_dst_mapped, size_t nwords, unsigned long long* h_flag_mapped,
                                  unsigned long long seq, cudaStream_t stream);
// planar float sRGB in [0,1] ([3][stride]) -> Lab planes; *d_bad |= 1 when a value is outside [0,1] or NaN
cudaError_t launch_unit_to_lab(const float* d_unit, size_t n, size_t stride, int whitepoint, float* d_lab, unsigned int* d_bad,
                               int sm_count, cudaStream_t stream);
cudaError_t launch_rgb_to_lab(const uint8_t* d_rgb, size_t n, size_t stride, int whitepoint, const float* d_table,
                              float* d_lab, float* d_unit, int sm_count, cudaStream_t stream);

// palettes [B][K][4] sRGB floats -> padded feature tables [B][K8] float4:
// pal_lab = (L, a, b, 0), pal_rgb = (r, g, b, 0); pad entries are 1e18.  Optionally clears d_zero[0 .. zero_words) (the
// result words of the evaluation launched next on the same stream).
cudaError_t launch_palette_features(const float* d_palettes, int B, int K, int whitepoint,
                                    float4* d_pal_lab, float4* d_pal_rgb, cudaStream_t stream, unsigned long long* d_zero = nullptr,
                                    size_t zero_words = 0);

// ---- the exchange step over NVLink / NVSwitch PEER MEMORY (round 2): an all-reduce of the integer result words done by the
// exporting CTA itself — no separate collective launch.  Every rank owns a MAILBOX in its own HBM (one arrival flag per sender,
// two parity sets of one slot per sender); the other ranks hold it mapped (cudaDeviceEnablePeerAccess inside one process,
// cudaIpcOpenMemHandle between processes).  Per exchange `seq` the CTA
//   1. stores its words into slot[seq & 1][my rank] of EVERY rank's mailbox (plain 8-byte stores travelling over NVLink),
//   2. fences at system scope and releases `seq` into its flag in every mailbox,
//   3. acquires the flags of all senders in its OWN mailbox (local polling; a globaltimer bound turns a dead peer into an
//      error word instead of a hung GPU),
//   4. adds the slots up locally — integer sums, so the order is free and every rank gets the same bits.
// Two parity sets suffice: a rank cannot start exchange seq + 2 before it has completed seq + 1, which needs every
// peer's flag for seq + 1, which that peer releases only after it has read its slots of seq.
// Payloads above kPeerCapWords (64 candidates x 256 colours = 16 k words) stay on ncclAllReduce: there the collective's
// launch latency no longer matters and one CTA's stores would.
constexpr int kPeerMaxRanks = 16;
constexpr unsigned kPeerCapWords = 4096;                 // 32 KB per sender and parity
constexpr unsigned kPeerFlagStride = 16;                 // words: one 128-byte line per sender's flag
constexpr size_t kPeerSlotBase = (size_t)kPeerMaxRanks * kPeerFlagStride;
constexpr size_t kPeerBoxWords = kPeerSlotBase + (size_t)2 * kPeerMaxRanks * kPeerCapWords;   // 1 MB + 2 KB
struct PeerExchange {
    unsigned long long* box[kPeerMaxRanks] = {};   // every rank's mailbox in THIS device's address space (box[rank] is local)
    int nranks = 0, rank = 0;                      // nranks == 0: no exchange
    unsigned long long seq = 0;                    // 1, 2, 3, ... in lockstep on every rank
    unsigned long long timeout_ns = 0;
    unsigned long long* status = nullptr;          // pinned host word: seq of an exchange that timed out (0 = fine)
};
#ifdef __CUDACC__
// called by ALL threads of ONE CTA (any block size >= nranks); words[0 .. nwords) in: this rank's partial sums (complete and
// visible: the caller has fenced), out: the totals (or zero, zero_after); host_dst (optional): the totals as well
__device__ __noinline__ static void peer_allreduce_cta(const PeerExchange& px, unsigned long long* words, unsigned nwords,
                                                       unsigned long long* host_dst, bool zero_after) {
    const int nr = px.nranks, me = px.rank;
    const unsigned tid = threadIdx.x;
    const size_t slot_mine = kPeerSlotBase + ((size_t)(px.seq & 1ull) * kPeerMaxRanks + me) * kPeerCapWords;
    for (unsigned i = tid; i < nwords; i += blockDim.x) {
        const unsigned long long v = __ldcg(words + i);
        for (int k = 0; k < nr; ++k) {
            int r = me + k; r -= r >= nr ? nr : 0;   // own mailbox first, then round the ring: the ranks do not all hit rank 0 at once
            *reinterpret_cast<volatile unsigned long long*>(px.box[r] + slot_mine + i) = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_dead;
    if (tid == 0) s_dead = 0;
    if ((int)tid < nr)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(px.box[tid] + (size_t)me * kPeerFlagStride), "l"(px.seq) : "memory");
    __syncthreads();
    if ((int)tid < nr) {
        const unsigned long long* f = px.box[me] + (size_t)tid * kPeerFlagStride;
        unsigned long long t0 = 0, v;
        for (unsigned spin = 0;; ++spin) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= px.seq) break;
            if ((spin & 255u) == 255u) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t0 == 0) t0 = t;
                else if (t - t0 > px.timeout_ns) { s_dead = 1; break; }
            }
        }
    }
    __syncthreads();
    if (s_dead && tid == 0 && px.status) *reinterpret_cast<volatile unsigned long long*>(px.status) = px.seq;
    const unsigned long long* base = px.box[me] + kPeerSlotBase + (size_t)(px.seq & 1ull) * kPeerMaxRanks * kPeerCapWords;
    for (unsigned i = tid; i < nwords; i += blockDim.x) {
        unsigned long long s = 0;
        for (int r = 0; r < nr; ++r) s += *reinterpret_cast<const volatile unsigned long long*>(base + (size_t)r * kPeerCapWords + i);
        if (host_dst) host_dst[i] = s;
        words[i] = zero_after ? 0ull : s;
    }
}
#endif
cudaError_t launch_peer_allreduce(const PeerExchange& px, unsigned long long* d_words, size_t nwords, unsigned long long* h_dst_mapped,
                                  unsigned long long* h_flag_mapped, unsigned long long host_seq, cudaStream_t stream);

// Optional tail of the small-palette scoring kernel (assign_reduce_kernel variant 1, K <= 32 — where a search iteration is
// latency-bound): the LAST CTA of the grid to finish copies every result word to pinned host memory and then writes a
// sequence number there (what export_results_kernel does as a separate launch) — one dependent stream operation fewer per
// search iteration (C1: 28.1 -> 26.1 us).  Not compiled into variant 3 or the pruned kernel: there it bought nothing and its
// mere presence cost the pruned kernel 1.5 %.  counter: one device word, zero on entry; the last CTA resets it.
struct ExportTail {
    unsigned long long* host_dst = nullptr;   // null = no tail
    unsigned long long* host_flag = nullptr;
    unsigned long long seq = 0;
    unsigned* counter = nullptr;
    const unsigned long long* src = nullptr;  // [nwords] all result words of the launch
    unsigned nwords = 0;
    bool peer_only = false;                   // a member of a multi-device context: exchange, no host export (host_dst unused)
    PeerExchange peer;                        // nranks > 0: the last CTA all-reduces the words over peer memory before it exports
    bool active() const { return host_dst != nullptr || peer_only; }
};
#ifdef __CUDACC__
// every thread of every CTA calls this after its last result atomic
// zero_after: the result words are left zero for the next launch (the one-launch evaluation has no clearing pass)
__device__ __forceinline__ void export_tail(const ExportTail& x, unsigned total_ctas, bool zero_after = false, unsigned long long seq_override = 0ull) {
    if (x.host_dst == nullptr && !x.peer_only) return;
    __shared__ unsigned s_ticket;
    __threadfence();     // this thread's result atomics are performed device-wide before the ticket is taken
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(x.counter, 1u);
    __syncthreads();
    if (s_ticket != total_ctas - 1) return;
    __threadfence();
    if (x.peer.nranks > 0) {   // sharded image: totals over all ranks first, over NVLink peer memory
        peer_allreduce_cta(x.peer, const_cast<unsigned long long*>(x.src), x.nwords, x.peer_only ? nullptr : x.host_dst, zero_after);
    } else {
        for (unsigned i = threadIdx.x; i < x.nwords; i += blockDim.x) {
            x.host_dst[i] = __ldcg(x.src + i);
            if (zero_after) const_cast<unsigned long long*>(x.src)[i] = 0ull;
        }
    }
    __threadfence_system();   // the words are visible to the host before the sequence number is
    __syncthreads();
    if (threadIdx.x == 0) {
        *x.counter = 0u;
        if (!x.peer_only) *reinterpret_cast<volatile unsigned long long*>(x.host_flag) = seq_override ? seq_override : x.seq;
    }
}
#endif

constexpr size_t kAllPixels = ~(size_t)0;
struct AssignArgs {
    const float* lab;      // [3][stride]
    const float* unit;     // [3][stride], required when space == 1
    size_t n, stride;
    const float4* pal_lab; // [B][K8]
    const float4* pal_rgb; // [B][K8]
    int B, K;
    int space;             // 0 LAB, 1 SRGB-assign / LAB-score
    bool want_sums;
    unsigned long long* results;  // [B][result_words], must be zeroed by the caller
    void* idx_out;         // [B][stride] u8 (K <= 256) or u16, or null
    int sm_count;
    size_t own_lo = 0, own_hi = kAllPixels;  // reduce only pixels [own_lo, own_hi) (default: all): halo rows of a shard are assigned but not counted
    int variant;           // 0 auto, 1 direct index tracking, 2 chunked min + recompute, 3 prefilter + exact
    ExportTail tail;       // optional (see above); only with variant 1 (explicit, or auto with K <= kDirectMaxColors)
};
constexpr int kDirectMaxColors = 32;  // auto picks variant 1 up to here (measured crossover, hq_kernels.cu)
cudaError_t launch_assign_reduce(const AssignArgs& a, cudaStream_t stream);
// ONE launch per evaluation for the searches that are latency-bound (the plugin's defaults: 8 colours, population 4): B * K <=
// kSmallPalColors sRGB colours travel as a kernel parameter, every CTA converts its candidate's palette itself, the last CTA
// exports the result words (a.tail, required) and zeroes them again.  h_palettes: [B][K][4] host floats (validated by the
// caller); a.results must be zero on entry and is zero again when the launch completes; K <= kDirectMaxColors, no index image.
constexpr int kSmallPalColors = 192;
cudaError_t launch_assign_small(const AssignArgs& a, const float* h_palettes, int whitepoint, cudaStream_t stream);
// The same evaluation as a PERSISTENT kernel for the length of a search (hq_kernels.cu, assign_persist_kernel): launched once, fed
// through a pinned mailbox word (h_cmd: sequence number of the evaluation wanted, ~0 = leave), palettes read from the pinned buffer
// h_palettes_mapped each iteration, totals + the sequence number exported as by launch_assign_small; h_exit becomes non-zero when
// the kernel has left (asked to, or idle for idle_ns).
constexpr unsigned long long kPersistQuitCmd = ~0ull;
cudaError_t launch_assign_persist(const AssignArgs& a, const float* h_palettes_mapped, float* d_palettes, int whitepoint, const unsigned long long* h_cmd,
                                  unsigned long long* h_exit, unsigned long long* d_cmd, unsigned long long first_seq, unsigned long long idle_ns, cudaStream_t stream);

// ---- exact assignment with geometric pruning (hq_pruned.cu): the own pixels are counting-sorted once per image by a
// coarse CIELAB cell into chunks of <= kPrunedChunkPx pixels with exact bounding boxes; per (chunk, candidate) only
// the colours that can be nearest to some pixel of the chunk are swept.  Results are bit-identical to
// launch_assign_reduce (LAB space, no index image).
constexpr int kPrunedChunkPx = 2048;
constexpr int kMaxColorsPruned = 4096;  // the pruned kernel's shared memory holds ~26 B (50 B with sums) per colour
size_t pruned_scratch_words();  // unsigned words of scratch launch_pruned_build_cells needs (totals at [.. - 2]: pixels, chunks)
// d_feat: [3][stride] feature planes (Lab for space 0, unit sRGB for space 1); pixels [lo, hi) are sorted into d_sorted
// [3][sstride]; d_perm (optional) receives the image position of every sorted pixel
// cell_bits (3..5, from pruned_cell_bits) = bits per axis of the coarse cells chunks never straddle
int pruned_cell_bits(size_t n, int space);
cudaError_t launch_pruned_build_cells(const float* d_feat, size_t stride, int space, int cell_bits, size_t lo, size_t hi, unsigned* d_scratch,
                                      float* d_sorted, size_t sstride, unsigned* d_perm, int sm_count, cudaStream_t st);
cudaError_t launch_pruned_build_chunks(const unsigned* d_scratch, int cell_bits, const float* d_sorted, size_t sstride, unsigned nchunks,
                                       unsigned* d_chunk_start, unsigned* d_chunk_len, float* d_box, cudaStream_t st);
struct PrunedArgs {
    const float* sorted; size_t sstride;          // [3][sstride] features of the pixels in cell order
    const unsigned* chunk_start; const unsigned* chunk_len; const float* box; unsigned nchunks;
    const float4* pal;                            // [B][K8] palette in the same feature space
    int B, K;
    bool want_sums;
    unsigned long long* results;                  // [B][result_words], zeroed by the caller
    unsigned long long* stats;                    // optional [2]: survivors summed over (chunk, candidate), number of (chunk, candidate)
    int sm_count;
    // index-producing mode: idx_out [B][istride] (u8 for K <= 256, else u16) written through perm; error word stays 0,
    // counts cover the image positions [own_lo, own_hi) only
    const unsigned* perm = nullptr; void* idx_out = nullptr; size_t istride = 0, own_lo = 0, own_hi = 0;
};
cudaError_t launch_pruned_assign(const PrunedArgs& a, cudaStream_t st);

// ---- palettes of ANY size (hq_bigk.cu): chunked sweep against a per-pixel running best in HBM, then one reduction pass.
// One candidate per call: d_pal_feat / d_pal_lab are that candidate's [K8] tables, d_out its result words (zeroed by the
// caller), d_best_d2 / d_best_idx [n] scratch that holds the final assignment afterwards, d_idx16 (optional, K <= 65535) the
// index image of every local pixel.
constexpr int kBigChunk = 2048;
cudaError_t launch_bigk_candidate(const float* d_feat, const float* d_lab, size_t n, size_t stride, size_t own_lo, size_t own_hi, const float4* d_pal_feat,
                                  const float4* d_pal_lab, int K, bool srgb, bool want_sums, float* d_best_d2, unsigned* d_best_idx,
                                  unsigned long long* d_out, uint16_t* d_idx16, int sm_count, cudaStream_t st);
cudaError_t launch_apply_palette_u32(const unsigned* d_idx, size_t n, const float* d_palette, uint8_t* d_out_rgb, float* d_out_f32, cudaStream_t st);

// indices -> output image of the chosen palette colours (u8 RGB packed and/or float RGBA)
cudaError_t launch_apply_palette(const void* d_idx, bool idx16, size_t n, const float* d_palette /*[K][4]*/,
                                 int K, uint8_t* d_out_rgb, float* d_out_f32, cudaStream_t stream);

// test hooks: run the single-source routines of hq_math.h on the device over a float range
cudaError_t launch_math_probe(int which, uint32_t first_bits, uint32_t count, float* d_out,
                              cudaStream_t stream);

// rows of a shard for the vertical filter: outputs are local rows [y_begin, y_begin + y_count); local row 0 is
// global row g0 of an image of gh rows (reflection happens at the GLOBAL borders)
struct ScRows {
    int y_begin, y_count, g0, gh;
};

// ---- S-CIELAB stage (hq_scielab.cu).  d_filters: [8][taps] = k1[t][3], k2[t][3], k3[t], |k3|[t]
constexpr int kMaxScielabTaps = 255;
cudaError_t launch_sc_rgb_to_opp(const uint8_t* d_rgb, size_t n, size_t stride, const float* d_table, float* d_opp, cudaStream_t st);
// float image planes -> opponent planes (in place allowed); d_bad (optional) |= 1 for values outside [0,1]
cudaError_t launch_sc_unit_to_opp(const float* d_unit, size_t n, size_t stride, float* d_opp, unsigned int* d_bad, cudaStream_t st);
cudaError_t launch_sc_palette_opp(const float* d_palettes, int total, float4* d_tab, cudaStream_t st, unsigned long long* d_zero = nullptr, int zero_words = 0);
cudaError_t launch_sc_export(const unsigned long long* d_results, const unsigned long long* d_sc_err, int words, size_t nwords, unsigned long long* h_dst_mapped,
                             unsigned long long* h_flag_mapped, unsigned long long seq, cudaStream_t st);
// original image: opp planes -> S-CIELAB Lab planes (d_tmp: 7 planes of scratch)
// h_filters: the same block on the host (nullptr = always use the generic kernels); with taps == 21
// (plugin defaults) the specialised kernels take it as a kernel parameter
// illuminant3 (optional): an explicit white point instead of the enum's (XYZtoScielab's float[] illuminant, ImageManipulation.java:285)
cudaError_t launch_sc_original(const float* d_opp, int w, int h, size_t stride, const float* d_filters, const float* h_filters, int taps,
                               int whitepoint, ScRows rows, float* d_tmp, float* d_lab_out, cudaStream_t st, const float* illuminant3 = nullptr);
// the reference class's one-shot entries on its interleaved float4 layouts (hq_rgb_to_xyz, hq_xyz_to_scielab, hq_scielab_set_image,
// hq_delta_e_images)
cudaError_t launch_sc_unit_to_xyz4(const float* d_r, const float* d_g, const float* d_b, size_t n, float* d_xyz4, unsigned int* d_bad, cudaStream_t st);
cudaError_t launch_sc_xyz4_to_opp(const float* d_xyz4, size_t n, size_t stride, float* d_opp, cudaStream_t st);
cudaError_t launch_sc_planes_to_f4(const float* d_planes, size_t n, size_t stride, float* d_out4, cudaStream_t st);
cudaError_t launch_sc_f4_to_planes(const float* d_in4, size_t n, size_t stride, float* d_planes, cudaStream_t st);
cudaError_t launch_sc_delta_e4(const float* d_a4, const float* d_b4, size_t n, float* d_e, float* d_err_img4, int de_type, cudaStream_t st);
// one candidate: index image + opponent table -> fixed-point sum of dE against d_lab_orig, added to *d_err
cudaError_t launch_sc_candidate(const void* d_idx, bool idx16, const float4* d_tab, int w, int h, size_t stride, const float* d_filters,
                                const float* h_filters, int taps, int whitepoint, ScRows rows, float* d_tmp, const float* d_lab_orig,
                                unsigned long long* d_err, cudaStream_t st, int de_type = 0, unsigned long long* d_nan = nullptr);
// de_type: 0 CIE76 (cl:209), 1 the CIE94 branch (cl:217-226); d_nan counts the pixels whose CIE94 value is NaN

// the whole population in one launch, no intermediate in HBM (taps == 21 and K <= kMaxColors; otherwise cudaErrorNotSupported:
// call launch_sc_candidate per candidate).  d_idx [B][stride], d_tab [B][K], d_err [B]
cudaError_t launch_sc_candidates_fused(const void* d_idx, bool idx16, const float4* d_tab, int K, int B, int w, int h, size_t stride,
                                       const float* h_filters, int taps, int whitepoint, ScRows rows, const float* d_lab_orig,
                                       unsigned long long* d_err, int sm_count, cudaStream_t st);

// error-image mode: dE map between two S-CIELAB images + fixed-point sum
// d_err: [2] words: fixed-point sum, NaN count (CIE94 only)
cudaError_t launch_sc_error_image(const float* d_lab_a, const float* d_lab_b, size_t n, size_t stride, float* d_map, uint8_t* d_map_u8,
                                  unsigned long long* d_err, int de_type, cudaStream_t st);
// identity-filter cost from index images under a dE type: d_err[B] += sum of dE(Lab(pixel), Lab(P[idx])) over [own_lo, own_hi), d_nan[B]
cudaError_t launch_sc_score_indices(const void* d_idx, bool idx16, const float* d_lab, size_t stride, size_t own_lo, size_t own_hi,
                                    const float4* d_pal_lab, int K8, int B, int de_type, unsigned long long* d_err, unsigned long long* d_nan,
                                    int sm_count, cudaStream_t st);

// FFMA-saturating probe: `iters` x 32 dependent-chain FMAs per thread (8 chains), scalar or packed
cudaError_t launch_fp32_peak(bool packed, int iters, int sm_count, float* d_out, cudaStream_t stream);

}  // namespace hq
