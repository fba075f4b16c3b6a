// hq_bigk.cu — palettes beyond what one CTA can stage (the plugin's "Number of colors" goes up to 2^24,
// HybridQuantization.java:192): the palette is swept in chunks of kBigChunk colours against a per-pixel RUNNING best
// (squared distance + index) kept in HBM, then one pass reduces error / counts / Lab sums from the final assignment.
// Same arithmetic and the same rule as everywhere (hq_dist2, strict '<', chunks in ascending order = first wins,
// OptimizedConvolution.cl:178-193), so the integers equal the exhaustive kernel's wherever both apply.  This is the
// coverage path, not a tuned one: ~10 instructions per (pixel, colour) pair.
#include "hq_kernels.cuh"
#include "hq_math.h"

namespace hq {
namespace {

constexpr int kBigThreads = 256;

__global__ void __launch_bounds__(kBigThreads) bigk_sweep_kernel(const float* __restrict__ feat, size_t n, size_t stride, const float4* __restrict__ pal, int k0,
                                                                 int kc, float* __restrict__ best_d2, unsigned* __restrict__ best_idx, int first) {
    __shared__ float4 s_p[kBigChunk];
    for (int k = threadIdx.x; k < kc; k += kBigThreads) s_p[k] = pal[k0 + k];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * kBigThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kBigThreads) {
        const float x0 = feat[i], x1 = feat[stride + i], x2 = feat[2 * stride + i];
        float bd = first ? __int_as_float(0x7f800000) : best_d2[i];
        unsigned bi = first ? 0u : best_idx[i];
#pragma unroll 4
        for (int k = 0; k < kc; ++k) {
            const float4 p = s_p[k];
            const float d = hq_dist2(x0, x1, x2, p.x, p.y, p.z);
            if (d < bd) { bd = d; bi = (unsigned)(k0 + k); }   // strict '<', ascending index: first wins
        }
        best_d2[i] = bd; best_idx[i] = bi;
    }
}

template <bool SUMS>
__global__ void __launch_bounds__(kBigThreads) bigk_finalize_kernel(const float* __restrict__ lab, size_t n, size_t stride, size_t own_lo, size_t own_hi,
                                                                    const float* __restrict__ best_d2, const unsigned* __restrict__ best_idx,
                                                                    const float4* __restrict__ pal_lab, int srgb, int K, unsigned long long* __restrict__ out,
                                                                    uint16_t* __restrict__ idx16) {
    long long fx = 0;
    for (size_t i = (size_t)blockIdx.x * kBigThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kBigThreads) {
        const unsigned k = best_idx[i];
        if (idx16) idx16[i] = (uint16_t)k;
        if (i < own_lo || i >= own_hi) continue;   // halo pixel of a row shard: assigned, not counted
        const float q0 = lab[i], q1 = lab[stride + i], q2 = lab[2 * stride + i];
        float d2 = best_d2[i];
        if (srgb) { const float4 pl = __ldg(pal_lab + k); d2 = hq_dist2(q0, q1, q2, pl.x, pl.y, pl.z); }   // assign in sRGB, score in CIELAB (cl:209)
        fx += hq_to_fx(HQ_FSQRT(d2));
        atomicAdd(out + 1 + k, 1ull);
        if (SUMS) {
            atomicAdd(out + 1 + K + 3 * (size_t)k, (unsigned long long)hq_to_fx(q0));
            atomicAdd(out + 1 + K + 3 * (size_t)k + 1, (unsigned long long)hq_to_fx(q1));
            atomicAdd(out + 1 + K + 3 * (size_t)k + 2, (unsigned long long)hq_to_fx(q2));
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) fx += __shfl_down_sync(0xffffffffu, fx, off);
    if ((threadIdx.x & 31) == 0 && fx) atomicAdd(out, (unsigned long long)fx);
}

__global__ void apply_palette_u32_kernel(const unsigned* __restrict__ idx, size_t n, const float4* __restrict__ pal, uint8_t* __restrict__ out_rgb,
                                         float4* __restrict__ out_f32) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = __ldg(pal + idx[i]);
    if (out_f32) out_f32[i] = c;
    if (out_rgb) {   // as apply_palette_kernel: float -> u8 is (int)(c * 255 + 0.5)
        out_rgb[3 * i] = (uint8_t)__float2int_rz(HQ_FADD(HQ_FMUL(c.x, 255.0f), 0.5f));
        out_rgb[3 * i + 1] = (uint8_t)__float2int_rz(HQ_FADD(HQ_FMUL(c.y, 255.0f), 0.5f));
        out_rgb[3 * i + 2] = (uint8_t)__float2int_rz(HQ_FADD(HQ_FMUL(c.z, 255.0f), 0.5f));
    }
}

}  // namespace

cudaError_t launch_bigk_candidate(const float* d_feat, const float* d_lab, size_t n, size_t stride, size_t own_lo, size_t own_hi, const float4* d_pal_feat,
                                  const float4* d_pal_lab, int K, bool srgb, bool want_sums, float* d_best_d2, unsigned* d_best_idx,
                                  unsigned long long* d_out, uint16_t* d_idx16, int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    size_t grid = (n + kBigThreads - 1) / kBigThreads;
    const size_t cap = (size_t)(sm_count > 0 ? sm_count : 148) * 8;
    if (grid > cap) grid = cap;
    for (int k0 = 0; k0 < K; k0 += kBigChunk)
        bigk_sweep_kernel<<<(unsigned)grid, kBigThreads, 0, st>>>(d_feat, n, stride, d_pal_feat, k0, K - k0 < kBigChunk ? K - k0 : kBigChunk, d_best_d2, d_best_idx, k0 == 0);
    if (want_sums) bigk_finalize_kernel<true><<<(unsigned)grid, kBigThreads, 0, st>>>(d_lab, n, stride, own_lo, own_hi, d_best_d2, d_best_idx, d_pal_lab, srgb ? 1 : 0, K, d_out, d_idx16);
    else bigk_finalize_kernel<false><<<(unsigned)grid, kBigThreads, 0, st>>>(d_lab, n, stride, own_lo, own_hi, d_best_d2, d_best_idx, d_pal_lab, srgb ? 1 : 0, K, d_out, d_idx16);
    return cudaGetLastError();
}

cudaError_t launch_apply_palette_u32(const unsigned* d_idx, size_t n, const float* d_palette, uint8_t* d_out_rgb, float* d_out_f32, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    apply_palette_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_idx, n, reinterpret_cast<const float4*>(d_palette), d_out_rgb, reinterpret_cast<float4*>(d_out_f32));
    return cudaGetLastError();
}

}  // namespace hq
