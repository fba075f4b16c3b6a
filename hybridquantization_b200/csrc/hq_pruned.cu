// hq_pruned.cu — EXACT nearest-palette assignment with geometric pruning (opt-in fast path of the
// population scoring, HQ_EVAL_PRUNE).
//
// The exhaustive kernel (assign_reduce_kernel, hq_kernels.cu) evaluates all N*K pixel-colour pairs per
// candidate, as quantizeAndConvertToOpp does (OptimizedConvolution.cl:178-193).  Every output of the
// scoring step is a SUM over pixels (error, per-colour counts, Lab sums), so the pixels may be visited
// in any order.  Once per image the pixels are therefore counting-sorted by a coarse cell of the feature space
// (3 to 5 bits per axis, chosen from the image size), and inside a cell by the Morton code of the finer bits, and
// cut into chunks of <= 2048 pixels that never straddle a cell; each chunk keeps its exact bounding box.  Per
// (chunk, candidate) a CTA then
//   1. bounds every colour against the box:  dmin_k <= |x - p_k| <= dmax_k  for every pixel x of the chunk,
//   2. takes U = min_k dmax_k (some colour is within U of every pixel) and keeps the colours with
//      dmin_k^2 <= U^2 * (1 + 2^-18): a discarded colour is strictly farther from every pixel of the chunk
//      than the colour that realises U, by a margin ~250x the rounding error of the fp32 distances, so it
//      can neither win nor tie,
//   3. runs the exact direct-form sweep (hq_dist2, strict '<', ascending colour index = the reference's
//      first-wins rule, cl:186) over the survivors only — typically 4-10 of 256.
// Winner, distance, counts, sums and error are bit-identical to the exhaustive kernel
// (tests/test_gpu_pruned.py); what changes is the amount of arithmetic: ~N*(S + c) instead of N*K.
//
// The index-producing mode (S-CIELAB chain, hq_quantize) sorts every local pixel in the assignment space, keeps
// each pixel's image position and scatters the winner's index there; the scoring mode needs no index image.
#include <cstdlib>

#include <mutex>

#include "hq_kernels.cuh"
#include "hq_math.h"

namespace hq {
namespace {

constexpr int kThreads = 256;
constexpr int kPxPerThread = kPrunedChunkPx / kThreads;  // 8
constexpr int kAxisBits = 7;                       // every feature axis is quantised to 7 bits for the ORDER of the sort
constexpr int kMaxCellBits = 5;                    // of which the top cb (3..5) name the coarse cell: chunks never straddle one
constexpr int kMaxCells = 1 << (3 * kMaxCellBits); // 32768
constexpr int kBins = 1 << (3 * kAxisBits);        // 2,097,152 counting-sort bins = cells x Morton sub-grid of the cell

// Bin of a pixel: coarse cell of the feature space (top cb bits per axis after the affine map q = (f + off) * scale) in
// the high bits, the Morton code of its position inside the cell (the remaining 7 - cb bits per axis) in the low bits.
//   CIELAB: L in [0,100] -> scale 1.28; a, b in [-128,128) -> off 128, scale 0.5   (cb = 5: cells of 3.1 x 8 x 8)
//   sRGB:   r, g, b in [0,1] -> scale 128                                         (cb = 5: cells of 1/32 per axis)
// Only an ORDER: any deterministic function works, exactness is irrelevant (the chunk boxes are exact).  Ordering the
// inside of a cell makes the chunks of a crowded cell compact (and their composition deterministic) instead of random
// subsets of the cell.  cb is chosen per image so that a cell holds about a chunk's worth of pixels (pruned_cell_bits).
struct BinMap { float off[3], scale[3]; int cb; };
__device__ __forceinline__ unsigned bin_of(float f0, float f1, float f2, const BinMap& m) {
    const unsigned l = (unsigned)min(max(__float2int_rd((f0 + m.off[0]) * m.scale[0]), 0), 127);
    const unsigned u = (unsigned)min(max(__float2int_rd((f1 + m.off[1]) * m.scale[1]), 0), 127);
    const unsigned v = (unsigned)min(max(__float2int_rd((f2 + m.off[2]) * m.scale[2]), 0), 127);
    const int sb = kAxisBits - m.cb;
    const unsigned cell = (((l >> sb) << m.cb | (u >> sb)) << m.cb) | (v >> sb);
    unsigned sub = 0;
    for (int i = sb - 1; i >= 0; --i) sub = (sub << 3) | (((l >> i) & 1u) << 2) | (((u >> i) & 1u) << 1) | ((v >> i) & 1u);
    return (cell << (3 * sb)) | sub;
}

__global__ void cell_hist_kernel(const float* __restrict__ feat, size_t stride, size_t lo, size_t hi, BinMap m, unsigned* __restrict__ hist) {
    for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[bin_of(feat[i], feat[stride + i], feat[2 * stride + i], m)], 1u);
}

// population of every cell = sum of its nsub bins (one thread per cell)
__global__ void cell_count_kernel(const unsigned* __restrict__ hist, int ncells, int nsub, unsigned* __restrict__ cell_cnt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    unsigned h = 0;
    for (int j = 0; j < nsub; ++j) h += hist[(size_t)c * nsub + j];
    cell_cnt[c] = h;
}

// one CTA of 1024 threads, `per` consecutive cells per thread: exclusive scans of the cell populations (first sorted position
// of every cell) and of the chunks per cell
__global__ void __launch_bounds__(1024) cell_scan_kernel(const unsigned* __restrict__ cell_cnt, int ncells, unsigned* __restrict__ cell_off,
                                                         unsigned* __restrict__ chunk_base, unsigned* __restrict__ totals) {
    __shared__ unsigned s_px[1024], s_ch[1024];
    const int t = threadIdx.x;
    const int per = (ncells + 1023) / 1024;
    unsigned px = 0, ch = 0;
    for (int i = 0; i < per; ++i) {
        const int c = t * per + i;
        const unsigned h = c < ncells ? cell_cnt[c] : 0u;
        px += h; ch += (h + kPrunedChunkPx - 1) / kPrunedChunkPx;
    }
    s_px[t] = px; s_ch[t] = ch;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
        const unsigned a = t >= off ? s_px[t - off] : 0u, b = t >= off ? s_ch[t - off] : 0u;
        __syncthreads();
        s_px[t] += a; s_ch[t] += b;
        __syncthreads();
    }
    unsigned opx = s_px[t] - px, och = s_ch[t] - ch;
    for (int i = 0; i < per; ++i) {
        const int c = t * per + i;
        if (c >= ncells) break;
        const unsigned h = cell_cnt[c];
        cell_off[c] = opx; chunk_base[c] = och;
        opx += h; och += (h + kPrunedChunkPx - 1) / kPrunedChunkPx;
    }
    if (t == 1023) { totals[0] = s_px[t]; totals[1] = s_ch[t]; }
}

// first sorted position of every bin: the cell's offset plus the populations of the cell's earlier bins (one thread per cell)
__global__ void bin_offset_kernel(const unsigned* __restrict__ hist, const unsigned* __restrict__ cell_off, int ncells, int nsub,
                                  unsigned* __restrict__ bin_off) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    unsigned o = cell_off[c];
    for (int j = 0; j < nsub; ++j) {
        bin_off[(size_t)c * nsub + j] = o;
        o += hist[(size_t)c * nsub + j];
    }
}

__global__ void cell_scatter_kernel(const float* __restrict__ feat, size_t stride, size_t lo, size_t hi, BinMap m, const unsigned* __restrict__ bin_off,
                                    unsigned* __restrict__ cursor, float* __restrict__ sorted, size_t sstride, unsigned* __restrict__ perm) {
    for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
        const float f0 = feat[i], f1 = feat[stride + i], f2 = feat[2 * stride + i];
        const unsigned c = bin_of(f0, f1, f2, m);
        const size_t pos = (size_t)bin_off[c] + atomicAdd(&cursor[c], 1u);
        sorted[pos] = f0; sorted[sstride + pos] = f1; sorted[2 * sstride + pos] = f2;
        if (perm) perm[pos] = (unsigned)i;  // where the pixel sits in the image (index-producing evaluations)
    }
}

__global__ void chunk_table_kernel(const unsigned* __restrict__ cell_cnt, const unsigned* __restrict__ cell_off, const unsigned* __restrict__ chunk_base,
                                   int ncells, unsigned* __restrict__ chunk_start, unsigned* __restrict__ chunk_len) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const unsigned h = cell_cnt[c], first = cell_off[c];
    for (unsigned j = 0, left = h; left > 0; ++j) {
        const unsigned len = left < (unsigned)kPrunedChunkPx ? left : (unsigned)kPrunedChunkPx;
        chunk_start[chunk_base[c] + j] = first + j * kPrunedChunkPx;
        chunk_len[chunk_base[c] + j] = len;
        left -= len;
    }
}

// exact bounding box of every chunk: one warp per chunk; box = (lo0, lo1, lo2, hi0, hi1, hi2)
__global__ void chunk_box_kernel(const float* __restrict__ sorted, size_t sstride, const unsigned* __restrict__ chunk_start,
                                 const unsigned* __restrict__ chunk_len, unsigned nchunks, float* __restrict__ box) {
    const unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nchunks) return;
    const size_t s = chunk_start[w];
    const unsigned len = chunk_len[w];
    const float INF = __int_as_float(0x7f800000);
    float lo[3] = {INF, INF, INF}, hi[3] = {-INF, -INF, -INF};
    for (unsigned i = lane; i < len; i += 32)
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            const float v = sorted[pl * sstride + s + i];
            lo[pl] = fminf(lo[pl], v); hi[pl] = fmaxf(hi[pl], v);
        }
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lo[pl] = fminf(lo[pl], __shfl_xor_sync(0xffffffffu, lo[pl], off));
            hi[pl] = fmaxf(hi[pl], __shfl_xor_sync(0xffffffffu, hi[pl], off));
        }
    if (lane == 0) {
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) { box[6 * (size_t)w + pl] = lo[pl]; box[6 * (size_t)w + 3 + pl] = hi[pl]; }
    }
}

struct PrunedParams {
    const float* sorted; size_t sstride;
    const unsigned* chunk_start; const unsigned* chunk_len; const float* box; unsigned nchunks;
    const float4* pal;      // [B][K8] palette in the feature space of the sorted copy
    int B, K, K8, words, b_per_cta;
    unsigned long long* results;
    unsigned long long* stats;  // optional: [0] += survivors summed over (chunk, candidate), [1] += (chunk, candidate) pairs
    // index-producing mode (IDXW != 0): the winner of sorted pixel i goes to idx_out[b * istride + perm[i]]; only pixels
    // whose image position lies in [own_lo, own_hi) are counted (halo rows of a shard are assigned, not counted)
    const unsigned* perm; void* idx_out; size_t istride, own_lo, own_hi;
};

// shared-memory carve-up: survivors live in per-warp SEGMENTS (segment g = colours [32g, 32g+32), filled from slot 32g
// upwards in ascending colour order), so the compaction needs no prefix sum across warps and the sweep still visits
// the survivors in ascending colour index.
extern __shared__ __align__(16) unsigned char pruned_smem_raw[];
template <bool SUMS>
struct PrunedSmem {  // typed views of the dynamic shared memory, recomputed where they are used so that they stay LDS/STS/ATOMS
    float4* surv;               // [K32] feature of the survivor in each slot
    unsigned long long* sum;    // [3*K32] per-slot Lab sums (SUMS)
    unsigned* cnt;              // [K32] per-slot pixel counts
    float* dmin;                // [K32] lower bound dmin^2 of every colour against the box
    unsigned* segcnt;           // [K32/32] survivors per segment
    unsigned short* list;       // [K32] colour index of each slot
    int K32;                    // K rounded up to whole segments
    __device__ __forceinline__ explicit PrunedSmem(int K) {
        K32 = (K + 31) & ~31;
        surv = reinterpret_cast<float4*>(pruned_smem_raw);
        sum = reinterpret_cast<unsigned long long*>(pruned_smem_raw + (size_t)K32 * 16);
        cnt = reinterpret_cast<unsigned*>(pruned_smem_raw + (size_t)K32 * 16 + (SUMS ? (size_t)K32 * 24 : 0));
        dmin = reinterpret_cast<float*>(cnt + K32);
        segcnt = reinterpret_cast<unsigned*>(dmin + K32);
        list = reinterpret_cast<unsigned short*>(segcnt + K32 / 32);
    }
};
inline size_t pruned_smem_bytes(int K, bool sums) {
    const size_t K32 = ((size_t)K + 31) & ~(size_t)31;
    return K32 * (16 + 4 + 4 + 2) + K32 / 32 * 4 + (sums ? K32 * 24 : 0);
}

// One candidate on one chunk whose pixels sit in registers (NS slots of 256 pixels).  3 CTA barriers.
// IDXW: 0 = scoring (error, counts, optional sums), 1 / 2 = write u8 / u16 indices through perm and count own pixels only.
template <int NS, bool SUMS, int IDXW>
__device__ __forceinline__ void score_candidate(const PrunedParams& p, unsigned* s_U, long long* s_err, int b, int parity,
                                                const float (&x0)[NS], const float (&x1)[NS], const float (&x2)[NS], size_t start, unsigned len,
                                                const float (&lo)[3], const float (&hi)[3]) {
    const int K = p.K, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PrunedSmem<SUMS> sm(K);
    const float INF = __int_as_float(0x7f800000);
    const int nseg = (K + 31) >> 5;
    const float4* pal = p.pal + (size_t)b * p.K8;
    // ---- 1. bounds of every colour against the box; U = min over colours of dmax^2 (non-negative floats order like their bits).
    // Colour k = tid stays in registers (all there is for K <= 256); colours tid + 256, tid + 512, ... go through sm.dmin.
    float umin = INF;
    auto bounds = [&](const float4& c, float& mn) {
        const float pc[3] = {c.x, c.y, c.z};
        float mx = 0.f;
        mn = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float near = fmaxf(fmaxf(lo[a] - pc[a], pc[a] - hi[a]), 0.f);  // > 0 when the colour lies outside the slab
            const float far = fmaxf(pc[a] - lo[a], hi[a] - pc[a]);
            mn = fmaf(near, near, mn); mx = fmaf(far, far, mx);
        }
        umin = fminf(umin, mx);
    };
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f);
    float mn0 = INF;
    if (tid < K) { c0 = __ldg(pal + tid); bounds(c0, mn0); }
    for (int k = tid + kThreads; k < sm.K32; k += kThreads) {
        float mn = INF;
        if (k < K) bounds(__ldg(pal + k), mn);
        sm.dmin[k] = mn;  // read back by the same thread after barrier (1); nothing the previous candidate's flush still reads
    }
    const unsigned wmin = __reduce_min_sync(0xffffffffu, __float_as_uint(umin));
    if (lane == 0) atomicMin(&s_U[parity], wmin);
    __syncthreads();  // (1)
    // survivors: dmin^2 <= U * (1 + 2^-18) (+ an absolute floor for U == 0); every other colour is strictly farther
    const float U = __uint_as_float(s_U[parity]);
    const float thr = fmaf(U, 0x1p-18f, U) + 1e-30f;
    if (tid == 0) s_U[parity ^ 1] = 0x7f800000u;  // for the next candidate (its atomicMin comes after barriers 2 and 3)
    for (int k = tid; k < sm.K32; k += kThreads) {  // k >> 5 is warp-uniform: each warp compacts its own 32-colour segment
        const bool keep = (k == tid ? mn0 : sm.dmin[k]) <= thr;  // false for the padding (INF)
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int slot = (k & ~31) + __popc(bal & ((1u << lane) - 1u));
        if (keep) { sm.surv[slot] = (k == tid) ? c0 : __ldg(pal + k); sm.list[slot] = (unsigned short)k; }
        if (lane == 0) sm.segcnt[k >> 5] = __popc(bal);
        sm.cnt[k] = 0u;
        if (SUMS) { sm.sum[3 * k] = 0ull; sm.sum[3 * k + 1] = 0ull; sm.sum[3 * k + 2] = 0ull; }
    }
    __syncthreads();  // (2)
    // ---- 2. exact sweep over the survivors: strict '<', ascending colour index = lowest index wins
    float best[NS];
    int bi[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) { best[j] = INF; bi[j] = 0; }
    unsigned S = 0;
    for (int g = 0; g < nseg; ++g) {
        const unsigned c = sm.segcnt[g];
        S += c;
        for (unsigned i = 0; i < c; ++i) {
            const float4 q = sm.surv[g * 32 + i];
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const float d = hq_dist2(x0[j], x1[j], x2[j], q.x, q.y, q.z);
                if (d < best[j]) { best[j] = d; bi[j] = g * 32 + (int)i; }
            }
        }
    }
    // ---- 3. per-slot counts, and either error / sums (scoring) or the index image
    long long err_acc = 0;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const unsigned i = j * kThreads + tid;
        if (i < len) {
            if (IDXW == 0) {
                err_acc += hq_to_fx(HQ_FSQRT(best[j]));
                atomicAdd(&sm.cnt[bi[j]], 1u);
                if (SUMS) {
                    atomicAdd(&sm.sum[3 * bi[j]], (unsigned long long)hq_to_fx(x0[j]));
                    atomicAdd(&sm.sum[3 * bi[j] + 1], (unsigned long long)hq_to_fx(x1[j]));
                    atomicAdd(&sm.sum[3 * bi[j] + 2], (unsigned long long)hq_to_fx(x2[j]));
                }
            } else {
                const size_t px = __ldg(p.perm + start + i);
                const unsigned k = sm.list[bi[j]];
                if (IDXW == 1) reinterpret_cast<uint8_t*>(p.idx_out)[(size_t)b * p.istride + px] = (uint8_t)k;
                else reinterpret_cast<uint16_t*>(p.idx_out)[(size_t)b * p.istride + px] = (uint16_t)k;
                if (px >= p.own_lo && px < p.own_hi) atomicAdd(&sm.cnt[bi[j]], 1u);
            }
        }
    }
    if (IDXW == 0) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) err_acc += __shfl_down_sync(0xffffffffu, err_acc, off);
        if (lane == 0) s_err[warp] = err_acc;
    }
    __syncthreads();  // (3) also orders the shared atomics before the flush; the next candidate's barrier (1) orders the flush
                      // before its writes to these arrays
    unsigned long long* out = p.results + (size_t)b * p.words;
    if (tid == 0) {
        if (IDXW == 0) {
            long long e = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) e += s_err[w];
            if (e != 0) atomicAdd(out, (unsigned long long)e);
        }
        if (p.stats) { atomicAdd(p.stats, (unsigned long long)S); atomicAdd(p.stats + 1, 1ull); }
    }
    for (int slot = tid; slot < sm.K32; slot += kThreads) {
        const unsigned c = ((unsigned)(slot & 31) < sm.segcnt[slot >> 5]) ? sm.cnt[slot] : 0u;
        if (c) {
            const int k = sm.list[slot];
            atomicAdd(out + 1 + k, (unsigned long long)c);
            if (SUMS) {
                atomicAdd(out + 1 + K + 3 * k, sm.sum[3 * slot]);
                atomicAdd(out + 1 + K + 3 * k + 1, sm.sum[3 * slot + 1]);
                atomicAdd(out + 1 + K + 3 * k + 2, sm.sum[3 * slot + 2]);
            }
        }
    }
}

// all candidates of this CTA on one chunk of <= NS*256 pixels
template <int NS, bool SUMS, int IDXW>
__device__ __forceinline__ void score_chunk(const PrunedParams& p, unsigned* s_U, long long* s_err, size_t start, unsigned len,
                                            const float (&lo)[3], const float (&hi)[3]) {
    const int tid = threadIdx.x;
    float x0[NS], x1[NS], x2[NS];  // the chunk's pixels stay in registers for every candidate
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const unsigned i = j * kThreads + tid;
        const bool ok = i < len;
        x0[j] = ok ? __ldg(p.sorted + start + i) : 0.f;
        x1[j] = ok ? __ldg(p.sorted + p.sstride + start + i) : 0.f;
        x2[j] = ok ? __ldg(p.sorted + 2 * p.sstride + start + i) : 0.f;
    }
    const int b_begin = blockIdx.y * p.b_per_cta, b_end = min(p.B, b_begin + p.b_per_cta);
    for (int b = b_begin; b < b_end; ++b) score_candidate<NS, SUMS, IDXW>(p, s_U, s_err, b, (b - b_begin) & 1, x0, x1, x2, start, len, lo, hi);
}

template <bool SUMS, int IDXW>
__global__ void __launch_bounds__(kThreads, SUMS ? 2 : 4) pruned_assign_kernel(const PrunedParams p) {
    __shared__ unsigned s_U[2];
    __shared__ long long s_err[kThreads / 32];
    if (threadIdx.x < 2) s_U[threadIdx.x] = 0x7f800000u;
    __syncthreads();
    const unsigned chunk = blockIdx.x;
    const size_t start = p.chunk_start[chunk];
    const unsigned len = p.chunk_len[chunk];
    float lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { lo[a] = __ldg(p.box + 6 * (size_t)chunk + a); hi[a] = __ldg(p.box + 6 * (size_t)chunk + 3 + a); }
    // the sweep and epilogue are specialised on the number of 256-pixel slots the chunk occupies: a half-empty chunk costs half
    if (len <= 2 * kThreads) score_chunk<2, SUMS, IDXW>(p, s_U, s_err, start, len, lo, hi);
    else if (len <= 4 * kThreads) score_chunk<4, SUMS, IDXW>(p, s_U, s_err, start, len, lo, hi);
    else if (len <= 6 * kThreads) score_chunk<6, SUMS, IDXW>(p, s_U, s_err, start, len, lo, hi);
    else score_chunk<kPxPerThread, SUMS, IDXW>(p, s_U, s_err, start, len, lo, hi);
}

template <bool SUMS, int IDXW>
cudaError_t launch_pruned_t(const PrunedParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    // process-wide and monotonic: the attribute belongs to the function on the device, not to the calling thread
    static std::mutex mu;
    static size_t configured[64];  // per device: largest dynamic shared memory set so far
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lock(mu);
        size_t& have = configured[dev & 63];
        if (smem > have) {
            e = cudaFuncSetAttribute(pruned_assign_kernel<SUMS, IDXW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            have = smem;
        }
    }
    pruned_assign_kernel<SUMS, IDXW><<<grid, kThreads, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace

// scratch layout (unsigned words): hist[kBins], bin_off[kBins], cursor[kBins], cell_cnt[kMaxCells], cell_off[kMaxCells],
// chunk_base[kMaxCells], totals[2]
size_t pruned_scratch_words() { return (size_t)kBins * 3 + (size_t)kMaxCells * 3 + 2; }

// Cell size for an image of n pixels: the finest of 5 / 4 / 3 bits per axis at which a non-empty cell still holds about a
// chunk's worth of pixels.  `occupancy` = expected fraction of non-empty cells for a maximally spread image: the sRGB gamut
// fills ~1/4 of the CIELAB bounding box, the sRGB cube is full.  (Too fine a grid leaves most chunks nearly empty — 63 pixels
// per chunk for a 1080p noise image in sRGB space at 5 bits — and the per-chunk work then dominates.)
int pruned_cell_bits(size_t n, int space) {
    if (const char* e = std::getenv("HQ_PRUNE_CELL_BITS")) { const int v = std::atoi(e); if (v >= 3 && v <= kMaxCellBits) return v; }
    const double occupancy = space == 1 ? 1.0 : 0.25;
    int cb = kMaxCellBits;
    while (cb > 3 && (double)n < 0.45 * kPrunedChunkPx * occupancy * (double)(1 << (3 * cb))) --cb;
    return cb;
}

static BinMap bin_map(int space, int cb) {
    BinMap m;
    if (space == 1) { for (int a = 0; a < 3; ++a) { m.off[a] = 0.f; m.scale[a] = 128.f; } }
    else { m.off[0] = 0.f; m.scale[0] = 1.28f; m.off[1] = m.off[2] = 128.f; m.scale[1] = m.scale[2] = 0.5f; }
    m.cb = cb;
    return m;
}

cudaError_t launch_pruned_build_cells(const float* d_feat, size_t stride, int space, int cell_bits, size_t lo, size_t hi, unsigned* d_scratch,
                                      float* d_sorted, size_t sstride, unsigned* d_perm, int sm_count, cudaStream_t st) {
    if (cell_bits < 3 || cell_bits > kMaxCellBits) return cudaErrorInvalidValue;
    unsigned *hist = d_scratch, *bin_off = d_scratch + kBins, *cursor = d_scratch + 2 * (size_t)kBins, *cell_cnt = d_scratch + 3 * (size_t)kBins,
             *cell_off = cell_cnt + kMaxCells, *chunk_base = cell_off + kMaxCells, *totals = chunk_base + kMaxCells;
    if (hi - lo >= 0xffffffffull || hi >= 0xffffffffull) return cudaErrorInvalidValue;  // 32-bit sorted positions / image positions
    cudaError_t e = cudaMemsetAsync(d_scratch, 0, pruned_scratch_words() * sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    const BinMap m = bin_map(space, cell_bits);
    const int ncells = 1 << (3 * cell_bits), nsub = kBins / ncells;
    if (hi > lo) cell_hist_kernel<<<sm_count * 8, 256, 0, st>>>(d_feat, stride, lo, hi, m, hist);
    cell_count_kernel<<<(ncells + 255) / 256, 256, 0, st>>>(hist, ncells, nsub, cell_cnt);
    cell_scan_kernel<<<1, 1024, 0, st>>>(cell_cnt, ncells, cell_off, chunk_base, totals);
    bin_offset_kernel<<<(ncells + 255) / 256, 256, 0, st>>>(hist, cell_off, ncells, nsub, bin_off);
    if (hi > lo) cell_scatter_kernel<<<sm_count * 8, 256, 0, st>>>(d_feat, stride, lo, hi, m, bin_off, cursor, d_sorted, sstride, d_perm);
    return cudaGetLastError();
}

cudaError_t launch_pruned_build_chunks(const unsigned* d_scratch, int cell_bits, const float* d_sorted, size_t sstride, unsigned nchunks,
                                       unsigned* d_chunk_start, unsigned* d_chunk_len, float* d_box, cudaStream_t st) {
    const unsigned *cell_cnt = d_scratch + 3 * (size_t)kBins, *cell_off = cell_cnt + kMaxCells, *chunk_base = cell_off + kMaxCells;
    const int ncells = 1 << (3 * cell_bits);
    chunk_table_kernel<<<(ncells + 255) / 256, 256, 0, st>>>(cell_cnt, cell_off, chunk_base, ncells, d_chunk_start, d_chunk_len);
    if (nchunks) chunk_box_kernel<<<(nchunks * 32 + 255) / 256, 256, 0, st>>>(d_sorted, sstride, d_chunk_start, d_chunk_len, nchunks, d_box);
    return cudaGetLastError();
}

cudaError_t launch_pruned_assign(const PrunedArgs& a, cudaStream_t st) {
    if (a.B <= 0 || a.K <= 0 || a.K > kMaxColorsPruned) return cudaErrorInvalidValue;
    if (a.nchunks == 0) return cudaSuccess;
    if (a.idx_out && (a.want_sums || !a.perm)) return cudaErrorInvalidValue;
    PrunedParams p;
    p.sorted = a.sorted; p.sstride = a.sstride; p.chunk_start = a.chunk_start; p.chunk_len = a.chunk_len; p.box = a.box; p.nchunks = a.nchunks;
    p.pal = a.pal; p.B = a.B; p.K = a.K; p.K8 = padded_colors(a.K); p.words = result_words(a.K, a.want_sums);
    p.results = a.results; p.stats = a.stats;
    p.perm = a.perm; p.idx_out = a.idx_out; p.istride = a.istride; p.own_lo = a.own_lo; p.own_hi = a.own_hi;
    // every CTA keeps its chunk in registers and loops over candidates; split the candidates over gridDim.y only as far
    // as needed to fill the machine (>= 8 CTAs per SM)
    int groups = 1;
    while ((long long)a.nchunks * groups < (long long)a.sm_count * 8 && groups < a.B) groups *= 2;
    if (groups > a.B) groups = a.B;
    p.b_per_cta = (a.B + groups - 1) / groups;
    groups = (a.B + p.b_per_cta - 1) / p.b_per_cta;
    const size_t smem = pruned_smem_bytes(a.K, a.want_sums);
    const dim3 grid(a.nchunks, (unsigned)groups);
    if (a.idx_out) return a.K <= 256 ? launch_pruned_t<false, 1>(p, grid, smem, st) : launch_pruned_t<false, 2>(p, grid, smem, st);
    return a.want_sums ? launch_pruned_t<true, 0>(p, grid, smem, st) : launch_pruned_t<false, 0>(p, grid, smem, st);
}

}  // namespace hq
