// hq_kernels.cu — hand-written sm_100a kernels of the HybridQuantization hot path.
//
//   rgb_to_lab_kernel        RGB->CIELAB of every pixel   (replaces RGB2XYZ + Opp2LAB,
//                            OptimizedConvolution.cl:79-90,124-145, with the Java CPU
//                            arithmetic of ScielabProcessor.java:279-311)
//   palette_features_kernel  candidate palettes -> Lab    (same arithmetic, B*K colours)
//   assign_reduce_kernel     nearest-palette argmin fused with the per-colour counts /
//                            Lab sums / total-error reduction, many candidates per launch
//                            (replaces quantizeAndConvertToOpp :172-199, CIEDE :201-209,
//                            the used-colour flags and the host-side averageArray of
//                            ImageManipulation.java:620-727,736-768)
//   apply_palette_kernel     final image                  (quantize, :147-170)
//
// HBM layout: the image lives as SoA fp32 planes [3][stride] (stride = n rounded up to 32
// pixels) so that every warp load is one fully coalesced 128-bit access; palettes live as
// padded float4 tables [B][K8].  All reductions are integers (counts, 2^-24 fixed point),
// hence independent of block order, grid size and GPU count.
//
// Tensor cores are deliberately not used: the inner dimension of the distance is 3.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "hq_kernels.cuh"
#include "hq_math.h"

namespace hq {
namespace {

constexpr int kThreads = 256;
constexpr int kPxPerThread = 4;
constexpr int kTilePx = kThreads * kPxPerThread;  // 1024 pixels per CTA iteration
constexpr int kChunk = 8;                         // colours per running-min chunk
constexpr float kFar = 1e18f;                     // padding colour coordinate

// ---- packed fp32x2 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2), IEEE round-to-nearest,
// ---- no flush-to-zero: bit-identical per lane to __fsub_rn / __fmul_rn / __fmaf_rn.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t dup2(float v) { return pack2(v, v); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float min3(float a, float b, float c) {  // FMNMX3
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// squared distances of one pixel to a PAIR of colours; same rounding sequence as hq_dist2
__device__ __forceinline__ uint64_t dist2_pair(uint64_t X, uint64_t Y, uint64_t Z, uint64_t P0,
                                               uint64_t P1, uint64_t P2) {
    const uint64_t d0 = sub2(X, P0), d1 = sub2(Y, P1), d2 = sub2(Z, P2);
    return fma2(d2, d2, fma2(d1, d1, mul2(d0, d0)));
}

// ====================================================================== RGB -> Lab

// 256-entry u8 -> unit / linear-light tables ([0..255] unit, [256..511] linear), built once per context
__global__ void decode_table_kernel(float* __restrict__ table) {
    const int v = threadIdx.x;
    const float u = hq_u8_to_unit((uint32_t)v);
    table[v] = u;
    table[256 + v] = hq_srgb_decode(u);
}

// ---- Lab of a PAIR of pixels from their linear-light RGB, packed f32x2 (sm_100a FMUL2 / FADD2 / FFMA2: one
// instruction per two pixels).  Every packed lane performs the operation sequence of the single-source scalar
// routines of hq_math.h — hq_dot3 (products and sums rounded, no contraction), the Markstein constant division
// of hq_div_const, the fp32 cube root of hq_cbrtf — so each lane's result is bit-identical to theirs.  What is
// dropped is hq_div_const's operand-range check: the kernel's inputs are u8 triples, and ALL 2^24 of them are
// compared with the oracle for both white points on the device (test_rgb_to_lab_all_16m_colours).
// Returns a 2-bit mask of the lanes whose cube roots need the fp64 routine (then the pixel is redone by
// lab_of_pixel_f64).
__device__ __noinline__ hq_float3 lab_of_pixel_f64(float R, float G, float B, hq_white white) {
    return hq_linrgb_to_lab_f64(R, G, B, white);
}
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even though both carry an explicit rounding
// modifier (observed on CUDA 12.9: L = 116*fy - 16 came out unrounded), which Java float arithmetic must not do.
// Wherever a packed PRODUCT feeds a packed SUM the sum is therefore written as fma(p, ONE, q) with ONE = (1.0f, 1.0f)
// taken from a kernel parameter: p*1 + q rounds exactly like p + q, and an fma cannot be fused any further.
__device__ __forceinline__ uint64_t add2_of_product(uint64_t p, uint64_t q, uint64_t ONE) { return fma2(p, ONE, q); }
__device__ __forceinline__ uint64_t dot3_pair(float a, uint64_t x, float b, uint64_t y, float c, uint64_t z, uint64_t ONE) {
    return add2_of_product(mul2(z, dup2(c)), add2_of_product(mul2(y, dup2(b)), mul2(x, dup2(a)), ONE), ONE);
}
__device__ __forceinline__ uint64_t div_const_pair(uint64_t x, float c, float rc) {  // RN(x / c), Markstein
    const uint64_t q0 = mul2(x, dup2(rc));
    const uint64_t rem = fma2(q0, dup2(-c), x);
    return fma2(rem, dup2(rc), q0);
}
// cube roots of two values, lane for lane the operation sequence of hq_cbrtf_fast (hq_math.h); amb is set when
// either lane's two roundings disagree (ordered compare: a lane holding 0 produces NaN here, is never flagged and
// is replaced by the linear segment in lab_of_pixel_pair)
__device__ __forceinline__ uint64_t cbrt_pair(uint64_t t2, bool& amb) {
    float t0, t1;
    unpack2(t2, t0, t1);
    float lg0, lg1, e0, e1, y0, y1, rq0, rq1;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg0) : "f"(t0));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg1) : "f"(t1));
    unpack2(mul2(pack2(lg0, lg1), dup2(0x1.555556p-2f)), e0, e1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(e0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(e1));
    const uint64_t y = pack2(y0, y1), yn = pack2(-y0, -y1);
    const uint64_t q = mul2(y, y);
    const uint64_t nql = fma2(yn, y, q);      // q - y*y, exact
    const uint64_t p = mul2(q, y);
    const uint64_t npl = fma2(q, yn, p);      // p - q*y, exact
    const uint64_t r = fma2(nql, y, add2(sub2(t2, p), npl));  // t - y^3
    float q0, q1;
    unpack2(q, q0, q1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rq0) : "f"(q0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rq1) : "f"(q1));
    const uint64_t d = mul2(mul2(r, dup2(0x1.555556p-2f)), pack2(rq0, rq1));
    const uint64_t s_lo = add2(y, fma2(y, dup2(-HQ_CBRT_ETA), d));
    const uint64_t s_hi = add2(y, fma2(y, dup2(HQ_CBRT_ETA), d));
    float a0, a1, b0, b1;
    unpack2(s_lo, a0, a1); unpack2(s_hi, b0, b1);
    amb = amb || (a0 < b0) || (a1 < b1);
    return s_lo;
}
// the linear segment of ScielabProcessor.java:301, t / (3*LABDELTA2) + 4f/29f, for the lanes with t <= LABDELTA3
__device__ __forceinline__ uint64_t lab_f_blend_pair(uint64_t t2, uint64_t cb) {
    float t0, t1, c0, c1, l0, l1;
    unpack2(t2, t0, t1); unpack2(cb, c0, c1);
    unpack2(add2(div_const_pair(t2, HQ_3LABDELTA2, HQ_RCP_3LABDELTA2), dup2(HQ_4_OVER_29)), l0, l1);
    return pack2(t0 > HQ_LABDELTA3 ? c0 : l0, t1 > HQ_LABDELTA3 ? c1 : l1);
}
__device__ __forceinline__ bool lab_of_pixel_pair(uint64_t R, uint64_t G, uint64_t B, const hq_white& white, uint64_t ONE,
                                                  uint64_t& L, uint64_t& A, uint64_t& Bv) {
    // ScielabProcessor.java:286-290, :295-298 (constants as in hq_linrgb_to_opp / hq_opp_to_lab)
    const uint64_t ox = dot3_pair(0.26641335000823f, R, 0.60316740257478f, G, 0.0011333302293f, B, ONE);
    const uint64_t oy = dot3_pair(-0.12197400229389f, R, 0.05598088396616f, G, 0.01326365114329f, B, ONE);
    const uint64_t oz = dot3_pair(-0.08033445917708f, R, -0.33146741170125f, G, 0.44913244757774f, B, ONE);
    const uint64_t X = dot3_pair(0.97959616044562807864f, ox, -1.5347157012664408981f, oy, 0.44459764330437399288f, oz, ONE);
    const uint64_t ty = dot3_pair(1.188977906742323787f, ox, 0.7643549575179937615f, oy, 0.13512574791125839373f, oz, ONE);
    const uint64_t Z = dot3_pair(1.2318333139247290457f, ox, 1.1631592597636512884f, oy, 2.0784075888008567862f, oz, ONE);
    const uint64_t tx = div_const_pair(X, white.x, white.rx);  // illuminant[1] == 1.0f for both white points (:20-21): ty = Y
    const uint64_t tz = div_const_pair(Z, white.z, white.rz);
    bool amb = false;
    uint64_t fx = cbrt_pair(tx, amb), fy = cbrt_pair(ty, amb), fz = cbrt_pair(tz, amb);
    float x0, x1, y0, y1, z0, z1;
    unpack2(tx, x0, x1); unpack2(ty, y0, y1); unpack2(tz, z0, z1);
    if (!(fminf(min3(x0, x1, y0), min3(y1, z0, z1)) > HQ_LABDELTA3)) {  // a dark pixel: some value on the linear segment
        fx = lab_f_blend_pair(tx, fx); fy = lab_f_blend_pair(ty, fy); fz = lab_f_blend_pair(tz, fz);
    }
    L = add2_of_product(mul2(dup2(116.0f), fy), dup2(-16.0f), ONE);
    A = mul2(dup2(500.0f), sub2(fx, fy));
    Bv = mul2(dup2(200.0f), sub2(fy, fz));
    return amb;
}

constexpr int kRlWarpPx = 128;  // pixels per warp iteration: 384 B of packed RGB, 4 pixels per lane
template <bool UNIT>
__global__ void __launch_bounds__(kThreads, 4)
rgb_to_lab_kernel(const uint8_t* __restrict__ rgb, size_t n, size_t stride, int whitepoint, float one,
                  const float* __restrict__ table, float* __restrict__ lab, float* __restrict__ unit) {
    const uint64_t ONE = dup2(one);  // 1.0f that ptxas cannot see (add2_of_product)
    __shared__ __align__(16) uint32_t s_stage[kThreads / 32][kRlWarpPx * 3 / 4];  // 96 words per warp
    __shared__ float s_lin[256];
    __shared__ float s_unit[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_unit[tid] = table[tid];
    s_lin[tid] = table[256 + tid];
    __syncthreads();  // the only CTA-wide barrier: every warp stages and converts its own tiles
    const hq_white white = hq_make_white(whitepoint);
    uint32_t* stage = s_stage[warp];
    const uint32_t ntiles = (uint32_t)((n + kRlWarpPx - 1) / kRlWarpPx);
    // tiles [0, nfull) are complete and 16-byte aligned: staged with one 128-bit load per lane
    const uint32_t nfull = (reinterpret_cast<uintptr_t>(rgb) & 15) == 0 ? (uint32_t)(n / kRlWarpPx) : 0u;
    const uint32_t tstride = gridDim.x * (kThreads / 32);
    const bool loader = lane < kRlWarpPx * 3 / 16;  // 24 lanes x 16 B = 384 B, fully coalesced
    const uint4* src = reinterpret_cast<const uint4*>(rgb) + lane;
    uint32_t tile = blockIdx.x * (kThreads / 32) + warp;  // a CTA's 8 warps cover 1024 consecutive pixels
    // software pipeline: the next tile's 16 bytes per lane are in flight while this tile is converted
    uint4 pre = make_uint4(0, 0, 0, 0);
    if (tile < nfull && loader) pre = __ldg(src + (size_t)tile * (kRlWarpPx * 3 / 16));
    for (; tile < ntiles; tile += tstride) {
        const size_t px0 = (size_t)tile * kRlWarpPx;
        __syncwarp();  // the previous tile's words have been read by every lane
        if (tile < nfull) {
            if (loader) reinterpret_cast<uint4*>(stage)[lane] = pre;
        } else {
            const size_t byte0 = px0 * 3, nbytes = (n - px0 < (size_t)kRlWarpPx ? n - px0 : (size_t)kRlWarpPx) * 3;
            uint8_t* dst = reinterpret_cast<uint8_t*>(stage);
            for (size_t i = lane; i < (size_t)kRlWarpPx * 3; i += 32) dst[i] = (i < nbytes) ? rgb[byte0 + i] : (uint8_t)0;
        }
        __syncwarp();
        if (tile + tstride < nfull && loader) pre = __ldg(src + (size_t)(tile + tstride) * (kRlWarpPx * 3 / 16));
        // 4 pixels = 3 words per lane, bank-conflict free (word stride 3 is coprime with 32)
        const uint32_t w0 = stage[3 * lane], w1 = stage[3 * lane + 1], w2 = stage[3 * lane + 2];
        auto byte_of = [&](int i) { const uint32_t w = i < 4 ? w0 : (i < 8 ? w1 : w2); return (w >> (8 * (i & 3))) & 255u; };
        float lin[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) lin[i] = s_lin[byte_of(i)];
        float L[4], A[4], Bv[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // two pixel PAIRS, packed f32x2 arithmetic
            uint64_t l2, a2, b2;
            const bool amb = lab_of_pixel_pair(pack2(lin[6 * h], lin[6 * h + 3]), pack2(lin[6 * h + 1], lin[6 * h + 4]),
                                               pack2(lin[6 * h + 2], lin[6 * h + 5]), white, ONE, l2, a2, b2);
            unpack2(l2, L[2 * h], L[2 * h + 1]); unpack2(a2, A[2 * h], A[2 * h + 1]); unpack2(b2, Bv[2 * h], Bv[2 * h + 1]);
            if (amb) {  // a cube root within 2^-15 ulp of a rounding boundary (about one pair in 3,000): both pixels through fp64
#pragma unroll
                for (int j = 2 * h; j < 2 * h + 2; ++j) {
                    const hq_float3 v = lab_of_pixel_f64(lin[3 * j], lin[3 * j + 1], lin[3 * j + 2], white);
                    L[j] = v.x; A[j] = v.y; Bv[j] = v.z;
                }
            }
        }
        const size_t p = px0 + 4 * (size_t)lane;
        if (p + 4 <= n) {  // 128-bit coalesced plane stores
            *reinterpret_cast<float4*>(lab + p) = make_float4(L[0], L[1], L[2], L[3]);
            *reinterpret_cast<float4*>(lab + stride + p) = make_float4(A[0], A[1], A[2], A[3]);
            *reinterpret_cast<float4*>(lab + 2 * stride + p) = make_float4(Bv[0], Bv[1], Bv[2], Bv[3]);
            if (UNIT) {
                *reinterpret_cast<float4*>(unit + p) = make_float4(s_unit[byte_of(0)], s_unit[byte_of(3)], s_unit[byte_of(6)], s_unit[byte_of(9)]);
                *reinterpret_cast<float4*>(unit + stride + p) = make_float4(s_unit[byte_of(1)], s_unit[byte_of(4)], s_unit[byte_of(7)], s_unit[byte_of(10)]);
                *reinterpret_cast<float4*>(unit + 2 * stride + p) = make_float4(s_unit[byte_of(2)], s_unit[byte_of(5)], s_unit[byte_of(8)], s_unit[byte_of(11)]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (p + j < n) {
                    lab[p + j] = L[j]; lab[stride + p + j] = A[j]; lab[2 * stride + p + j] = Bv[j];
                    if (UNIT) {
                        unit[p + j] = s_unit[byte_of(3 * j)];
                        unit[stride + p + j] = s_unit[byte_of(3 * j + 1)];
                        unit[2 * stride + p + j] = s_unit[byte_of(3 * j + 2)];
                    }
                }
        }
    }
}

// ---- planar float image (HybridQuantization.java:95-98: any Icy data type rescaled to floats in [0,1]) -> Lab.
// No table applies to arbitrary floats: every channel goes through the single-source hq_srgb_decode (exact double pow
// to float, hq_math.h) like a palette colour does.  Values outside [0,1] (or NaN) lie outside the domain that
// routine is verified on; they are reported through `bad` and the image is refused.
__global__ void __launch_bounds__(256) unit_to_lab_kernel(const float* __restrict__ unit, size_t n, size_t stride, int whitepoint,
                                                          float* __restrict__ lab, unsigned int* __restrict__ bad) {
    const hq_white white = hq_make_white(whitepoint);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float r = unit[i], g = unit[stride + i], b = unit[2 * stride + i];
        hq_float3 v; v.x = v.y = v.z = 0.f;
        if (r >= 0.f && r <= 1.f && g >= 0.f && g <= 1.f && b >= 0.f && b <= 1.f) v = hq_srgb_to_lab(r, g, b, white);
        else atomicOr(bad, 1u);
        lab[i] = v.x; lab[stride + i] = v.y; lab[2 * stride + i] = v.z;
    }
}

// ====================================================================== palettes -> features
__global__ void palette_features_kernel(const float* __restrict__ pal, int B, int K, int K8,
                                        int whitepoint, float4* __restrict__ pal_lab,
                                        float4* __restrict__ pal_rgb, unsigned long long* __restrict__ zero, size_t zero_words) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // the result words of the evaluation that follows on the same stream are cleared here (saves a memset node per iteration)
    for (size_t z = (size_t)i; z < zero_words; z += (size_t)gridDim.x * blockDim.x) zero[z] = 0ull;
    if (i >= B * K8) return;
    const int b = i / K8, k = i - b * K8;
    float4 lab = make_float4(kFar, kFar, kFar, 0.f), rgbv = lab;
    if (k < K) {
        const float4 c = reinterpret_cast<const float4*>(pal)[(size_t)b * K + k];
        const hq_float3 v = hq_srgb_to_lab(c.x, c.y, c.z, hq_make_white(whitepoint));
        lab = make_float4(v.x, v.y, v.z, 0.f);
        rgbv = make_float4(c.x, c.y, c.z, 0.f);
    }
    pal_lab[i] = lab;
    pal_rgb[i] = rgbv;
}

// ---- TMA bulk copies (cp.async.bulk, global -> shared, completion counted in bytes on an mbarrier) and the mbarrier
// primitives of the pixel pipeline of the small-palette kernel
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "HQ_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra HQ_MBAR_DONE_%=;\n"
        "bra HQ_MBAR_WAIT_%=;\n"
        "HQ_MBAR_DONE_%=:\n"
        "}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// ====================================================================== assign + reduce
struct AssignParams {
    const float* feat;   // [3][stride] planes the argmin runs on (Lab, or unit sRGB)
    const float* lab;    // [3][stride] Lab planes (scoring / sums)
    size_t n, stride;
    const float4* pal_feat;  // [B][K8]
    const float4* pal_lab;   // [B][K8]
    int K, K8, words;
    float xmax0, xmax1, xmax2;  // bounds of |feature| over the image (prefilter error bound)
    size_t own_lo, own_hi;      // pixels [own_lo, own_hi) are reduced; the rest (halo rows of a shard) only get an index
    unsigned long long* results;
    void* idx_out;
    ExportTail tail;
    int use_tma;   // variant 1: pixel tiles staged by TMA bulk copies through a 3-stage shared-memory ring
    int whitepoint;   // assign_small_kernel converts its palettes itself
    float one_f = 1.0f, zero_f = 0.0f;   // a 1 and a 0 the compiler cannot see through (track_min)
};
constexpr int kPxStages = 3;   // tiles in flight per CTA (variant 1 with TMA): 3 x 12 KB

// Ambiguous pixels deferred per CTA (variant 3) and the most pixels a CTA of that variant is given: ~0.2 % of the pixels are
// ambiguous at K = 256 (0.7 % at K = 1,024, where the chunk id takes 7 mantissa bits), and an entry that finds the list full is
// resolved on the spot by ONE lane while 31 wait.  Round 1 sized the grid by waves alone: a CTA of a 64 MP image held 900 k pixels,
// overflowed its 1,024 entries and the 64 MP runs sat 1.5 points below the 4K ones; so did K = 1,024 at 64 candidates per launch.
// (Per-warp lists drained inside the tile loop were measured too: no overflow at any size, but 2.5 % slower at 4K.)
constexpr int kWorklistCap = 2048;
constexpr long long kV3MaxPxPerCta = 128 * 1024;
#ifndef HQ_PREFILTER_SCALAR
#define HQ_PREFILTER_SCALAR 0  // 1: scalar FFMA sweep instead of packed FFMA2 (experiment)
#endif

// shared-memory carve-up, identical on host (size) and device (pointers)
template <int VARIANT, bool SRGB, bool SUMS>
struct AssignSmem {
    size_t off_pairs01, off_lab, off_sum, off_pairs2, off_coef, off_skew, off_cnt, off_wl, off_px, total;
    int skew_len;
    __host__ __device__ explicit AssignSmem(int K8) {
        skew_len = (K8 / kChunk) * (kChunk + 1);
        size_t o = 0;
        off_pairs01 = o; if (VARIANT != 3) o += (size_t)K8 / 2 * 16;       // (f0,f0',f1,f1') per colour pair
        off_coef = o;    if (VARIANT == 3) o += (size_t)((K8 + 15) / 16 * 16) * 16;   // (-2p0,-2p1,-2p2,|p|^2) per colour, an even number of chunks
        off_lab = o;     if (SRGB) o += (size_t)K8 * 16;                   // Lab of the palette for scoring
        off_sum = o;     if (SUMS) o += (size_t)K8 * 3 * 8;                // per-colour Lab sums
        off_pairs2 = o;  if (VARIANT != 3) o += (size_t)K8 / 2 * 8;        // (f2,f2') per colour pair
        off_skew = o;    if (VARIANT != 1) o += (size_t)3 * skew_len * 4;  // skewed SoA copy of the features
        off_cnt = o;     o += (size_t)K8 * 4;                              // per-colour counts
        off_wl = o;      if (VARIANT == 3) o += (size_t)kWorklistCap * 4;
        o = (o + 127) / 128 * 128;
        off_px = o;      if (VARIANT == 1) o += (size_t)kPxStages * 3 * kTilePx * 4;   // TMA ring of pixel tiles: [stage][plane][1024]
        total = o;
    }
};

// exact squared distances of one pixel to the 8 colours of chunk c: lowest index of the minimum
__device__ __forceinline__ void exact_chunk(const float* __restrict__ s_sk, int skew_len, int c, float x0, float x1,
                                            float x2, float& bd, int& bi) {
    const int s0 = c * (kChunk + 1);
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
        const float d = hq_dist2(x0, x1, x2, s_sk[s0 + i], s_sk[skew_len + s0 + i], s_sk[2 * skew_len + s0 + i]);
        if (d < bd) { bd = d; bi = c * kChunk + i; }  // strict '<', ascending index: first wins
    }
}

// exact sweep over every colour (the ambiguous pixels of variant 3: rare, so kept out of line): (squared distance bits, index)
__device__ __noinline__ unsigned long long exact_all_colours(const float* __restrict__ s_sk, int skew_len, int nchunks, float x0, float x1, float x2) {
    float bd = __int_as_float(0x7f800000);
    int bi = 0;
    for (int c = 0; c < nchunks; ++c) exact_chunk(s_sk, skew_len, c, x0, x1, x2, bd, bi);
    return ((unsigned long long)(unsigned)bi << 32) | (unsigned long long)__float_as_uint(bd);
}

// VARIANT 1: running (min, index) per colour, direct (x-p)^2 form — small palettes.
// VARIANT 2: direct form, running min per chunk of 8 colours with FMNMX3; the winning chunk is
//            re-evaluated once per pixel to recover the index (first colour whose distance equals
//            the minimum = the reference's strict '<' / lowest-index rule, OptimizedConvolution.cl:186).
// VARIANT 3: PREFILTER.  s_k = x.(-2 p_k) + |p_k|^2 = d^2 - |x|^2 costs 3 FMA per (pixel, colour)
//            instead of 6 FMA-pipe operations, but its fp32 cancellation error is too large to
//            decide with.  It is only used to find the chunk that can contain the winner:
//              |s_k - (D_k - |x|^2)| <= E      (E from the magnitudes actually present, see below)
//              exact winner k* satisfies  s(k*) <= s_min + T,  T = 2E + 2.1*gamma*D_min
//            (gamma = 6*2^-24 bounds the relative error of the exact fp32 distance).  If the
//            second-best CHUNK minimum exceeds s_min + T the winner lies in the best chunk, which
//            is then evaluated EXACTLY (direct form, first-wins).  Otherwise the pixel is ambiguous
//            (≈0.1 % at K=256) and goes through the exact sweep over all colours — deferred to a
//            per-CTA worklist so that warps do not diverge.  Results are bit-identical to variants 1/2.
// 2 CTAs/SM (121 registers, no stack frame) measured 68.8 % of FP32 peak vs 65.5 % at 3 CTAs/SM (80 registers)
#ifndef HQ_V3_MIN_CTAS
#define HQ_V3_MIN_CTAS 2
#endif
#ifndef HQ_V1_MIN_CTAS
#define HQ_V1_MIN_CTAS 3   // variant 1 (K <= 32): 80 registers; 4 (64 registers) was measured too
#endif
#ifndef HQ_V3_UNROLL
#define HQ_V3_UNROLL 2
#endif
[[maybe_unused]] constexpr int kV3Unroll = HQ_V3_UNROLL;  // chunks per unrolled iteration of the prefilter sweep (HQ_V3_BOOK == 0)
// Round 2: (a) 8 pixels per thread in variant 3 (two 1024-pixel sub-tiles per iteration, four packed pixel pairs): every
// LDS.128 of a colour's coefficients now feeds 12 FFMA2 instead of 6; (b) chunk bookkeeping through the MANTISSA: the chunk
// id replaces the low bits of the chunk minimum (one LOP3), so "best and second-best chunk" of a PAIR of chunks is
// min, max, max, min3, min — 3.5 ALU-pipe instructions per pixel and chunk instead of 5 (FMNMX x2, FSETP, FSEL, SEL), and the
// winner's chunk id falls out of the best value itself.  The perturbation (< 2^bits ulp) is added to the ambiguity threshold.
#ifndef HQ_V3_PX
#define HQ_V3_PX 8
#endif
#ifndef HQ_V3_BOOK
#define HQ_V3_BOOK 1
#endif
static_assert(HQ_V3_PX == 4 || HQ_V3_PX == 8, "variant 3 handles one or two 1024-pixel sub-tiles per iteration");
template <int VARIANT> struct AssignGeom {
    static constexpr int kPx = VARIANT == 3 ? HQ_V3_PX : kPxPerThread;   // pixels per thread and iteration
    static constexpr int kSub = kPx / kPxPerThread;                      // 1024-pixel sub-tiles per iteration
    static constexpr int kTile = kTilePx * kSub;
};
// Palettes as a KERNEL PARAMETER (the one-launch evaluation of small searches, launch_assign_small): B*K <= kSmallPalColors
// sRGB colours travel with the launch command itself — no H2D copy, no separate palette kernel, no PCIe read from the kernel.
// (sized in steps: the plugin's defaults, 8 colours x 4 candidates, launch with 512 bytes of parameters, not 3 KB)
template <int NCOL> struct SmallPalettes { float v[NCOL * 4]; };   // [B][K][4] as the plugin lays them out (SWASA.java:42-50)

// sRGB palette of candidate b (parameter space) -> the staged tables of variant 1, with the arithmetic of
// palette_features_kernel (hq_srgb_to_lab = hq_linrgb_to_lab of the three decoded channels): 3K threads decode one channel
// each (the exact pow(., 2.4) is the long pole: in parallel it costs one evaluation instead of three), K threads finish.
// Out of line: its fp64 code must not shape the register allocation of the sweep.
__device__ __noinline__ void stage_small_palette(const float* __restrict__ cpal, int K, int K8, int whitepoint, bool srgb, float* __restrict__ pla_f,
                                                 float* __restrict__ pb_f, float4* __restrict__ s_lab, float* __restrict__ s_lin, bool fresh = false) {
    // fresh: the palettes live in pinned host memory that the host rewrites between the iterations of a PERSISTENT kernel (below):
    // every read goes to the source (ld.cv), never to a line this SM fetched an iteration ago
    auto rd = [&](int i) { return fresh ? __ldcv(cpal + i) : cpal[i]; };
    const int tid = threadIdx.x;
    if (tid < 3 * K) s_lin[tid] = hq_srgb_decode(rd((tid / 3) * 4 + tid % 3));
    __syncthreads();
    if (tid < K8) {
        const int k = tid;
        float f0 = kFar, f1 = kFar, f2 = kFar;
        float4 lab = make_float4(kFar, kFar, kFar, 0.f);
        if (k < K) {
            const hq_float3 v = hq_linrgb_to_lab(s_lin[3 * k], s_lin[3 * k + 1], s_lin[3 * k + 2], hq_make_white(whitepoint));
            lab = make_float4(v.x, v.y, v.z, 0.f);
            if (srgb) { f0 = rd(4 * k); f1 = rd(4 * k + 1); f2 = rd(4 * k + 2); }
            else { f0 = v.x; f1 = v.y; f2 = v.z; }
        }
        pla_f[(k >> 1) * 4 + (k & 1)] = f0;
        pla_f[(k >> 1) * 4 + 2 + (k & 1)] = f1;
        pb_f[k] = f2;
        if (srgb) s_lab[k] = lab;
    }
}

// Running (minimum, index) of variant 1.  `if (d < best) { best = d; idx = k; }` compiles to FSETP + FSEL + SEL, three ALU-pipe
// instructions per (pixel, colour) pair beside the six FMA-pipe operations of the distance.  EXPERIMENT (HQ_V1_PRED_TRACK=1, off):
// keep the FSETP but make the two updates predicated FFMAs (best = d * 1 + (-0), index as a float = d * 0 + k, the 1 and the 0 from
// kernel parameters so that ptxas cannot fold them back into selects; a predicated integer move IS folded back into a SEL).  ptxas
// emits exactly that (64 @P FFMA per 32 pairs, no FSEL / SEL left) and the integers are unchanged, but it measured SLOWER — K=32 x 64
// candidates 4.48 -> 5.12 ms, K=8 1.68 -> 1.72 ms (profiles/r02/smallk_pred_track_ab.txt): the packed distance arithmetic already
// keeps the FMA pipe busier than the ALU pipe, and the selects ride beside it for free.
#ifndef HQ_V1_PRED_TRACK
#define HQ_V1_PRED_TRACK 0
#endif
__device__ __forceinline__ void track_min(float d, float kf, float& best, float& idxf, float one_f, float zero_f) {
#if HQ_V1_PRED_TRACK
    // (the index update multiplies the finite, non-negative distance by an opaque zero: d * 0 + kf == kf, and nothing of it can be hoisted)
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %2, %0;\n\t@p fma.rn.f32 %1, %2, %5, %3;\n\t@p fma.rn.f32 %0, %2, %4, 0f80000000;\n\t}"
        : "+f"(best), "+f"(idxf) : "f"(d), "f"(kf), "f"(one_f), "f"(zero_f));
#else
    if (d < best) { best = d; idxf = kf; }   // (not compiled into the default build: the select form keeps its integer index)
#endif
}

template <int VARIANT, bool SRGB, bool SUMS, int IDXW, bool FUSED>
__device__ __forceinline__ void assign_body(const AssignParams& p, const float* __restrict__ cpal, unsigned long long seq_override = 0ull, bool fresh_pal = false) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K, K8 = p.K8;
    const int tid = threadIdx.x;
    const int b = blockIdx.x;  // candidate index varies fastest: co-resident CTAs share pixels in L2
    const AssignSmem<VARIANT, SRGB, SUMS> L(K8);
    const int skew_len = L.skew_len;
    float4* s_pla = reinterpret_cast<float4*>(smem_raw + L.off_pairs01);
    float4* s_coef = reinterpret_cast<float4*>(smem_raw + L.off_coef);
    float4* s_lab = reinterpret_cast<float4*>(smem_raw + L.off_lab);
    unsigned long long* s_sum = reinterpret_cast<unsigned long long*>(smem_raw + L.off_sum);
    float2* s_pb = reinterpret_cast<float2*>(smem_raw + L.off_pairs2);
    float* s_sk = reinterpret_cast<float*>(smem_raw + L.off_skew);
    unsigned int* s_cnt = reinterpret_cast<unsigned int*>(smem_raw + L.off_cnt);
    unsigned int* s_wl = reinterpret_cast<unsigned int*>(smem_raw + L.off_wl);
    __shared__ unsigned int s_wl_n;
    __shared__ unsigned int s_bound[4];  // max |p0|, |p1|, |p2|, |p|^2 over the real colours (float bits)
    if (tid < 4) s_bound[tid] = 0u;
    if (tid == 0) s_wl_n = 0u;
    __syncthreads();
    // one-launch evaluation: the first tile's pixels are requested before the palette is converted (the exact pow of the sRGB decode
    // is a ~1 us dependent chain; the loads fly meanwhile)
    float4 pf0 = make_float4(0.f, 0.f, 0.f, 0.f), pf1 = pf0, pf2 = pf0;
    bool pf_ok = false;
    if (FUSED) {
        const size_t bs = (size_t)blockIdx.y * kTilePx + (size_t)kPxPerThread * tid;
        if (bs + 4 <= p.n) {
            pf0 = __ldg(reinterpret_cast<const float4*>(p.feat + bs));
            pf1 = __ldg(reinterpret_cast<const float4*>(p.feat + p.stride + bs));
            pf2 = __ldg(reinterpret_cast<const float4*>(p.feat + 2 * p.stride + bs));
            pf_ok = true;
        }
    }
    if (FUSED) {
        __shared__ float s_lin[3 * kDirectMaxColors];
        stage_small_palette(cpal + (size_t)b * K * 4, K, K8, p.whitepoint, SRGB, reinterpret_cast<float*>(s_pla), reinterpret_cast<float*>(s_pb), s_lab, s_lin, fresh_pal);
        for (int k = tid; k < K8; k += kThreads) {
            s_cnt[k] = 0u;
            if (SUMS) { s_sum[3 * k] = 0ull; s_sum[3 * k + 1] = 0ull; s_sum[3 * k + 2] = 0ull; }
        }
    } else {
        const float4* gf = p.pal_feat + (size_t)b * K8;
        const float4* gl = p.pal_lab + (size_t)b * K8;
        float* pla_f = reinterpret_cast<float*>(s_pla);
        float* pb_f = reinterpret_cast<float*>(s_pb);
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, me = 0.f;
        if (VARIANT == 3)   // the sweep takes chunks in pairs: a palette with an odd chunk count gets one more chunk that never wins
            for (int k = K8 + tid; k < (K8 + 15) / 16 * 16; k += kThreads) s_coef[k] = make_float4(0.f, 0.f, 0.f, 1e30f);
        for (int k = tid; k < K8; k += kThreads) {
            const float4 f = gf[k];
            if (VARIANT != 3) {
                pla_f[(k >> 1) * 4 + (k & 1)] = f.x;
                pla_f[(k >> 1) * 4 + 2 + (k & 1)] = f.y;
                pb_f[k] = f.z;
            } else {
                if (k < K) {
                    // |p|^2 rounded once from fp64; -2p is exact
                    const float e = (float)__fma_rn((double)f.z, (double)f.z, __fma_rn((double)f.y, (double)f.y, __dmul_rn((double)f.x, (double)f.x)));
                    s_coef[k] = make_float4(-2.f * f.x, -2.f * f.y, -2.f * f.z, e);
                    m0 = fmaxf(m0, fabsf(f.x)); m1 = fmaxf(m1, fabsf(f.y)); m2 = fmaxf(m2, fabsf(f.z)); me = fmaxf(me, e);
                } else {
                    s_coef[k] = make_float4(0.f, 0.f, 0.f, 1e30f);  // padding never wins
                }
            }
            if (VARIANT != 1) {
                const int s = (k / kChunk) * (kChunk + 1) + (k % kChunk);
                s_sk[s] = f.x; s_sk[skew_len + s] = f.y; s_sk[2 * skew_len + s] = f.z;
            }
            if (SRGB) s_lab[k] = gl[k];
            s_cnt[k] = 0u;
            if (SUMS) { s_sum[3 * k] = 0ull; s_sum[3 * k + 1] = 0ull; s_sum[3 * k + 2] = 0ull; }
        }
        if (VARIANT == 3) {  // non-negative floats order like their bit patterns
            atomicMax(&s_bound[0], __float_as_uint(m0)); atomicMax(&s_bound[1], __float_as_uint(m1));
            atomicMax(&s_bound[2], __float_as_uint(m2)); atomicMax(&s_bound[3], __float_as_uint(me));
        }
    }
    __syncthreads();
    // prefilter error bound: |p|^2 carries one rounding, each of the 3 fma's one more, all of
    // magnitude <= R = max|p|^2 + 2*sum_i max|x_i|*max|p_i|;  E = 4u*R, inflated by 25 %.
    float E = 0.f;
    if (VARIANT == 3) {
        const float R = __uint_as_float(s_bound[3]) + 2.f * (p.xmax0 * __uint_as_float(s_bound[0]) + p.xmax1 * __uint_as_float(s_bound[1]) +
                                                             p.xmax2 * __uint_as_float(s_bound[2]));
        E = 5.0f * 5.9604645e-8f * R + 1e-30f;
    }

    const float* f0 = p.feat; const float* f1 = p.feat + p.stride; const float* f2 = p.feat + 2 * p.stride;
    const float* l0 = p.lab;  const float* l1 = p.lab + p.stride;  const float* l2 = p.lab + 2 * p.stride;
    const size_t n = p.n;
    constexpr int PX = AssignGeom<VARIANT>::kPx, SUB = AssignGeom<VARIANT>::kSub, TILE = AssignGeom<VARIANT>::kTile;
    const size_t ntiles = (n + TILE - 1) / TILE;
    const float INF = __int_as_float(0x7f800000);
    const int nchunks = K8 / kChunk;
    long long err_acc = 0;
    unsigned long long* out = p.results + (size_t)b * p.words;

    // error / counts / sums / index of ONE resolved pixel (q = its Lab, d2v = exact squared distance in
    // the assignment space)
    auto account = [&](int k, float d2v, float q0, float q1, float q2) {
        if (SRGB) {  // assign in sRGB, score in CIELAB (OptimizedConvolution.cl:209)
            const float4 pl = s_lab[k];
            d2v = hq_dist2(q0, q1, q2, pl.x, pl.y, pl.z);
        }
        err_acc += hq_to_fx(HQ_FSQRT(d2v));
        atomicAdd(&s_cnt[k], 1u);
        if (SUMS) {
            atomicAdd(&s_sum[3 * k], (unsigned long long)hq_to_fx(q0));
            atomicAdd(&s_sum[3 * k + 1], (unsigned long long)hq_to_fx(q1));
            atomicAdd(&s_sum[3 * k + 2], (unsigned long long)hq_to_fx(q2));
        }
    };
    auto resolve = [&](size_t px, int k, float d2v, float q0, float q1, float q2, bool write_idx) {
        if (write_idx) {
            if (IDXW == 1) reinterpret_cast<uint8_t*>(p.idx_out)[(size_t)b * p.stride + px] = (uint8_t)k;
            if (IDXW == 2) reinterpret_cast<uint16_t*>(p.idx_out)[(size_t)b * p.stride + px] = (uint16_t)k;
        }
        if (px < p.own_lo || px >= p.own_hi) return;  // halo pixel of a row shard: assigned, not counted
        account(k, d2v, q0, q1, q2);
    };
    // exact sweep over every colour (ambiguous pixels of variant 3)
    auto exact_all = [&](float x0, float x1, float x2, float& bd, int& bi) {
        const unsigned long long r = exact_all_colours(s_sk, skew_len, nchunks, x0, x1, x2);
        bd = __uint_as_float((unsigned)r); bi = (int)(r >> 32);
    };

    // ---- variant 1: the feature planes of this CTA's tiles arrive through a ring of kPxStages tiles in shared memory, filled by
    // TMA bulk copies (3 x 4 KB per tile, one elected thread, completion counted on an mbarrier) kPxStages - 1 tiles ahead of
    // the arithmetic: no registers held for data in flight, no load instructions in the warps' instruction stream
    float* s_px = reinterpret_cast<float*>(smem_raw + L.off_px);
    __shared__ __align__(8) unsigned long long s_full[kPxStages], s_empty[kPxStages];
    const bool tma = VARIANT == 1 && p.use_tma != 0;
    const size_t my_tiles = blockIdx.y < ntiles ? (ntiles - blockIdx.y + gridDim.y - 1) / gridDim.y : 0;
    auto tma_issue = [&](size_t i) {   // tile i of this CTA -> stage i % kPxStages (thread 0 only)
        const size_t tb = (blockIdx.y + i * gridDim.y) * kTilePx;
        const unsigned cnt = (unsigned)(n - tb < (size_t)kTilePx ? n - tb : (size_t)kTilePx);
        const unsigned bytes = (cnt * 4u + 15u) & ~15u;   // the planes are padded to 32 pixels: the rounded copy stays inside
        const int st = (int)(i % kPxStages);
        mbar_expect_tx(&s_full[st], 3u * bytes);
        bulk_g2s(s_px + (size_t)(st * 3 + 0) * kTilePx, f0 + tb, bytes, &s_full[st]);
        bulk_g2s(s_px + (size_t)(st * 3 + 1) * kTilePx, f1 + tb, bytes, &s_full[st]);
        bulk_g2s(s_px + (size_t)(st * 3 + 2) * kTilePx, f2 + tb, bytes, &s_full[st]);
    };
    if (tma) {
        if (tid == 0) {
            for (int st = 0; st < kPxStages; ++st) { mbar_init(&s_full[st], 1u); mbar_init(&s_empty[st], (unsigned)kThreads); }
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0)
            for (size_t i = 0; i < my_tiles && i < (size_t)kPxStages - 1; ++i) tma_issue(i);
    }

    // variant 3, chunk id in the mantissa: ids 0 .. 2*ceil(nchunks/2)-1 need `bits` bits; a value moves by < 2^bits ulp
    const int nchunks2 = (nchunks + 1) & ~1;
    const int id_bits = nchunks2 > 1 ? 32 - __clz(nchunks2 - 1) : 1;
    const unsigned id_mask = (1u << id_bits) - 1u;
    const float rho25 = 2.5f * __uint_as_float((unsigned)(127 + id_bits - 23) << 23);   // 2.5 * 2^(bits-23)
    (void)id_mask; (void)rho25;

    size_t it = 0;
    for (size_t tile = blockIdx.y; tile < ntiles; tile += gridDim.y, ++it) {
        // pixel j of this thread: base + (j >> 2) * kTilePx + (j & 3) — every sub-tile is read with one 128-bit load per plane
        const size_t base = tile * TILE + (size_t)kPxPerThread * tid;
        auto pix = [&](int j) { return base + (size_t)(j >> 2) * kTilePx + (size_t)(j & 3); };
        float x0[PX], x1[PX], x2[PX];
        int nv[SUB];   // valid pixels of each sub-tile's quad
        if (tma) {
            // refill first: tile it + kPxStages - 1 goes into the stage tile it - 1 used, once every thread has read that one
            if (tid == 0 && it + kPxStages - 1 < my_tiles) {
                if (it >= 1) mbar_wait(&s_empty[(it - 1) % kPxStages], (unsigned)(((it - 1) / kPxStages) & 1));
                tma_issue(it + kPxStages - 1);
            }
            const int st = (int)(it % kPxStages);
            mbar_wait(&s_full[st], (unsigned)((it / kPxStages) & 1));
            const float* sp = s_px + (size_t)st * 3 * kTilePx + kPxPerThread * tid;
            const float4 a = *reinterpret_cast<const float4*>(sp);
            const float4 c = *reinterpret_cast<const float4*>(sp + kTilePx);
            const float4 d = *reinterpret_cast<const float4*>(sp + 2 * kTilePx);
            mbar_arrive(&s_empty[st]);
            x0[0] = a.x; x0[1] = a.y; x0[2] = a.z; x0[3] = a.w;
            x1[0] = c.x; x1[1] = c.y; x1[2] = c.z; x1[3] = c.w;
            x2[0] = d.x; x2[1] = d.y; x2[2] = d.z; x2[3] = d.w;
            nv[0] = base + 4 <= n ? 4 : (base < n ? (int)(n - base) : 0);
            if (nv[0] < 4) {   // the tail of the last tile: what lies beyond the image is not pixels
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j >= nv[0]) { x0[j] = 0.f; x1[j] = 0.f; x2[j] = 0.f; }
            }
        } else {
#pragma unroll
            for (int s = 0; s < SUB; ++s) {
                const size_t bs = base + (size_t)s * kTilePx;
                if (bs + 4 <= n) {
                    const bool pre = FUSED && it == 0 && pf_ok;
                    const float4 a = pre ? pf0 : __ldg(reinterpret_cast<const float4*>(f0 + bs));
                    const float4 c = pre ? pf1 : __ldg(reinterpret_cast<const float4*>(f1 + bs));
                    const float4 d = pre ? pf2 : __ldg(reinterpret_cast<const float4*>(f2 + bs));
                    x0[4 * s] = a.x; x0[4 * s + 1] = a.y; x0[4 * s + 2] = a.z; x0[4 * s + 3] = a.w;
                    x1[4 * s] = c.x; x1[4 * s + 1] = c.y; x1[4 * s + 2] = c.z; x1[4 * s + 3] = c.w;
                    x2[4 * s] = d.x; x2[4 * s + 1] = d.y; x2[4 * s + 2] = d.z; x2[4 * s + 3] = d.w;
                    nv[s] = 4;
                } else {
                    nv[s] = bs < n ? (int)(n - bs) : 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool ok = j < nv[s];
                        x0[4 * s + j] = ok ? f0[bs + j] : 0.f; x1[4 * s + j] = ok ? f1[bs + j] : 0.f; x2[4 * s + j] = ok ? f2[bs + j] : 0.f;
                    }
                }
            }
        }
        auto valid = [&](int j) { return (j & 3) < nv[j >> 2]; };

        float best[PX];  // exact squared distance of the winner
        int idx[PX];
        bool deferred[PX];
        bool any_def = false;
#pragma unroll
        for (int j = 0; j < PX; ++j) { best[j] = INF; idx[j] = 0; deferred[j] = false; }

        if (VARIANT == 1) {
            uint64_t X[PX], Y[PX], Z[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) { X[j] = pack2(x0[j], x0[j]); Y[j] = pack2(x1[j], x1[j]); Z[j] = pack2(x2[j], x2[j]); }
#if HQ_V1_PRED_TRACK
            float idxf[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) idxf[j] = 0.f;
            float kf = 0.f;
#pragma unroll 4
            for (int q = 0; q < K8 / 2; ++q) {
                const float4 la = s_pla[q];
                const float2 bb = s_pb[q];
                const uint64_t P0 = pack2(la.x, la.y), P1 = pack2(la.z, la.w), P2 = pack2(bb.x, bb.y);
                const float kf1 = kf + 1.0f;
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    float dlo, dhi;
                    unpack2(dist2_pair(X[j], Y[j], Z[j], P0, P1, P2), dlo, dhi);
                    track_min(dlo, kf, best[j], idxf[j], p.one_f, p.zero_f);
                    track_min(dhi, kf1, best[j], idxf[j], p.one_f, p.zero_f);
                }
                kf += 2.0f;
            }
#pragma unroll
            for (int j = 0; j < PX; ++j) idx[j] = __float2int_rn(idxf[j]);
#else
#pragma unroll 4
            for (int q = 0; q < K8 / 2; ++q) {
                const float4 la = s_pla[q];
                const float2 bb = s_pb[q];
                const uint64_t P0 = pack2(la.x, la.y), P1 = pack2(la.z, la.w), P2 = pack2(bb.x, bb.y);
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    float dlo, dhi;
                    unpack2(dist2_pair(X[j], Y[j], Z[j], P0, P1, P2), dlo, dhi);
                    if (dlo < best[j]) { best[j] = dlo; idx[j] = 2 * q; }
                    if (dhi < best[j]) { best[j] = dhi; idx[j] = 2 * q + 1; }
                }
            }
#endif
        } else if (VARIANT == 2) {
            uint64_t X[PX], Y[PX], Z[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) { X[j] = pack2(x0[j], x0[j]); Y[j] = pack2(x1[j], x1[j]); Z[j] = pack2(x2[j], x2[j]); }
            int cidx[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) cidx[j] = 0;
            for (int c = 0; c < nchunks; ++c) {
                float m[PX];
#pragma unroll
                for (int j = 0; j < PX; ++j) m[j] = INF;
#pragma unroll
                for (int q = 0; q < kChunk / 2; ++q) {
                    const float4 la = s_pla[c * (kChunk / 2) + q];
                    const float2 bb = s_pb[c * (kChunk / 2) + q];
                    const uint64_t P0 = pack2(la.x, la.y), P1 = pack2(la.z, la.w), P2 = pack2(bb.x, bb.y);
#pragma unroll
                    for (int j = 0; j < PX; ++j) {
                        float dlo, dhi;
                        unpack2(dist2_pair(X[j], Y[j], Z[j], P0, P1, P2), dlo, dhi);
                        m[j] = min3(m[j], dlo, dhi);
                    }
                }
#pragma unroll
                for (int j = 0; j < PX; ++j)
                    if (m[j] < best[j]) { best[j] = m[j]; cidx[j] = c; }
            }
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const int s0 = cidx[j] * (kChunk + 1);
                int found = 0;
#pragma unroll
                for (int i = kChunk - 1; i >= 0; --i) {
                    const float d = hq_dist2(x0[j], x1[j], x2[j], s_sk[s0 + i], s_sk[skew_len + s0 + i], s_sk[2 * skew_len + s0 + i]);
                    if (d == best[j]) found = i;
                }
                idx[j] = cidx[j] * kChunk + found;
            }
        } else {
            // ---- prefilter sweep: PX / 2 pixel PAIRS per thread, palette coefficients broadcast
            constexpr int NP = PX / 2;
            uint64_t X[NP], Y[NP], Z[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) { X[q] = pack2(x0[2 * q], x0[2 * q + 1]); Y[q] = pack2(x1[2 * q], x1[2 * q + 1]); Z[q] = pack2(x2[2 * q], x2[2 * q + 1]); }
            float sbest[PX], second[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) { sbest[j] = INF; second[j] = INF; }
            // minimum of the 8 prefilter values of chunk c, per pixel
            auto chunk_min = [&](int c, float (&m)[PX]) {
#pragma unroll
                for (int j = 0; j < PX; ++j) m[j] = INF;
#pragma unroll
                for (int q = 0; q < kChunk; q += 2) {
                    const float4 u = s_coef[c * kChunk + q], v = s_coef[c * kChunk + q + 1];
#if HQ_PREFILTER_SCALAR
#pragma unroll
                    for (int j = 0; j < PX; ++j) {
                        const float su = __fmaf_rn(x0[j], u.x, __fmaf_rn(x1[j], u.y, __fmaf_rn(x2[j], u.z, u.w)));
                        const float sv = __fmaf_rn(x0[j], v.x, __fmaf_rn(x1[j], v.y, __fmaf_rn(x2[j], v.z, v.w)));
                        m[j] = min3(m[j], su, sv);
                    }
#else
                    const uint64_t ua = pack2(u.x, u.x), ub = pack2(u.y, u.y), uc = pack2(u.z, u.z), ue = pack2(u.w, u.w);
                    const uint64_t va = pack2(v.x, v.x), vb = pack2(v.y, v.y), vc = pack2(v.z, v.z), ve = pack2(v.w, v.w);
#pragma unroll
                    for (int r = 0; r < NP; ++r) {
                        float a0, a1, b0, b1;
                        unpack2(fma2(X[r], ua, fma2(Y[r], ub, fma2(Z[r], uc, ue))), a0, a1);
                        unpack2(fma2(X[r], va, fma2(Y[r], vb, fma2(Z[r], vc, ve))), b0, b1);
                        m[2 * r] = min3(m[2 * r], a0, b0); m[2 * r + 1] = min3(m[2 * r + 1], a1, b1);
                    }
#endif
                }
            };
#if HQ_V3_BOOK
            // two chunks per iteration; each chunk minimum carries its chunk id in its low mantissa bits, so the two smallest
            // chunk minima (and the winner's id) are kept with min / max alone
            const unsigned nmask = ~id_mask;
#pragma unroll 1
            for (int c = 0; c < nchunks2; c += 2) {
                float ma[PX], mb[PX];
                chunk_min(c, ma);
                chunk_min(c + 1, mb);
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const float a = __uint_as_float((__float_as_uint(ma[j]) & nmask) | (unsigned)c);
                    const float d = __uint_as_float((__float_as_uint(mb[j]) & nmask) | (unsigned)(c + 1));
                    const float lo = fminf(a, d), hi = fmaxf(a, d);
                    second[j] = min3(second[j], hi, fmaxf(lo, sbest[j]));
                    sbest[j] = fminf(sbest[j], lo);
                }
            }
#else
            int cidx[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) cidx[j] = 0;
#pragma unroll (kV3Unroll)
            for (int c = 0; c < nchunks; ++c) {
                float m[PX];
                chunk_min(c, m);
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    second[j] = fminf(second[j], fmaxf(m[j], sbest[j]));  // 2nd smallest chunk minimum
                    if (m[j] < sbest[j]) { sbest[j] = m[j]; cidx[j] = c; }
                }
            }
#endif
            // ---- decide: unique chunk -> exact evaluation of that chunk; else defer / exact sweep.
            // Straight-line: the winning chunk of ALL the thread's pixels is evaluated exactly in one basic block (independent
            // chains the scheduler interleaves; as a branch per pixel the 8 evaluations ran one after the other, each waiting for
            // its own shared-memory loads); the rare ambiguous pixels are dealt with afterwards and overwrite what was computed.
            bool amb[PX];
            bool any_amb = false;
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const float nx = fmaf(x2[j], x2[j], fmaf(x1[j], x1[j], x0[j] * x0[j]));
                float T = 2.01f * E + 1.0e-6f * (fabsf(sbest[j] + nx) + E);
#if HQ_V3_BOOK
                // every carried value is within 2^bits ulp of the chunk minimum it stands for: if the winner lay outside the
                // best chunk, second <= sbest + T + rho (|m_best| + |m_winner's chunk|) would hold, both magnitudes <= (|sbest| + T)(1 + 4 rho)
                T += rho25 * (fabsf(sbest[j]) + T);
                const int cw = min((int)(__float_as_uint(sbest[j]) & id_mask), nchunks - 1);   // (a padding chunk never wins)
#else
                const int cw = cidx[j];
#endif
                amb[j] = valid(j) && !(second[j] > sbest[j] + T);
                any_amb = any_amb || amb[j];
                exact_chunk(s_sk, skew_len, cw, x0[j], x1[j], x2[j], best[j], idx[j]);
            }
            if (any_amb) {
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    if (amb[j]) {
                        const unsigned slot = atomicAdd(&s_wl_n, 1u);
                        if (slot < (unsigned)kWorklistCap && pix(j) < 0xffffffffull) {
                            s_wl[slot] = (unsigned)pix(j);
                            deferred[j] = true;
                            any_def = true;
                        } else {
                            exact_all(x0[j], x1[j], x2[j], best[j], idx[j]);  // worklist full: resolve here
                        }
                    }
                }
            }
        }

        // ---- per-pixel epilogue: error, counts, sums, indices
        float q0[PX], q1[PX], q2[PX];  // Lab of the pixel
        if (SRGB) {
#pragma unroll
            for (int s = 0; s < SUB; ++s) {
                const size_t bs = base + (size_t)s * kTilePx;
                if (nv[s] == 4) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(l0 + bs));
                    const float4 c = __ldg(reinterpret_cast<const float4*>(l1 + bs));
                    const float4 d = __ldg(reinterpret_cast<const float4*>(l2 + bs));
                    q0[4 * s] = a.x; q0[4 * s + 1] = a.y; q0[4 * s + 2] = a.z; q0[4 * s + 3] = a.w;
                    q1[4 * s] = c.x; q1[4 * s + 1] = c.y; q1[4 * s + 2] = c.z; q1[4 * s + 3] = c.w;
                    q2[4 * s] = d.x; q2[4 * s + 1] = d.y; q2[4 * s + 2] = d.z; q2[4 * s + 3] = d.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool ok = j < nv[s];
                        q0[4 * s + j] = ok ? l0[bs + j] : 0.f; q1[4 * s + j] = ok ? l1[bs + j] : 0.f; q2[4 * s + j] = ok ? l2[bs + j] : 0.f;
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < PX; ++j) { q0[j] = x0[j]; q1[j] = x1[j]; q2[j] = x2[j]; }
        }
        // small palettes are instruction-bound in this epilogue: a thread whose four pixels are all own pixels (every thread but
        // those at the edges of a shard) skips the per-pixel range tests
        bool all_own = !any_def;   // (variant 3: and none of them waits on the worklist)
#pragma unroll
        for (int s = 0; s < SUB; ++s) all_own = all_own && nv[s] == 4 && base + (size_t)s * kTilePx >= p.own_lo && base + (size_t)s * kTilePx + 4 <= p.own_hi;
        if (VARIANT != 2 && all_own) {
#pragma unroll
            for (int j = 0; j < PX; ++j) account(idx[j], best[j], q0[j], q1[j], q2[j]);
        } else {
#pragma unroll
            for (int j = 0; j < PX; ++j)
                if (valid(j) && !deferred[j]) resolve(pix(j), idx[j], best[j], q0[j], q1[j], q2[j], false);
        }
#pragma unroll
        for (int s = 0; s < SUB; ++s) {
            const size_t bs = base + (size_t)s * kTilePx;
            if (IDXW == 1) {
                uint8_t* o = reinterpret_cast<uint8_t*>(p.idx_out) + (size_t)b * p.stride + bs;
                if (nv[s] == 4) {
                    *reinterpret_cast<uint32_t*>(o) = (uint32_t)idx[4 * s] | ((uint32_t)idx[4 * s + 1] << 8) | ((uint32_t)idx[4 * s + 2] << 16) | ((uint32_t)idx[4 * s + 3] << 24);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < nv[s]) o[j] = (uint8_t)idx[4 * s + j];
                }
            } else if (IDXW == 2) {
                uint16_t* o = reinterpret_cast<uint16_t*>(p.idx_out) + (size_t)b * p.stride + bs;
                if (nv[s] == 4) {
                    *reinterpret_cast<uint2*>(o) = make_uint2((uint32_t)idx[4 * s] | ((uint32_t)idx[4 * s + 1] << 16), (uint32_t)idx[4 * s + 2] | ((uint32_t)idx[4 * s + 3] << 16));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < nv[s]) o[j] = (uint16_t)idx[4 * s + j];
                }
            }
        }
    }

    if (VARIANT == 3) {
        // ---- deferred (ambiguous) pixels: exact sweep over all colours, one pixel per thread.
        // The barrier also orders the packed index stores above before the per-pixel fix-ups.
        __syncthreads();
        const unsigned nwl = min(s_wl_n, (unsigned)kWorklistCap);
        for (unsigned i = tid; i < nwl; i += kThreads) {
            const size_t px = s_wl[i];
            float bd; int bi;
            exact_all(f0[px], f1[px], f2[px], bd, bi);
            resolve(px, bi, bd, l0[px], l1[px], l2[px], true);
        }
    }

    // ---- CTA reduction: warp shuffles -> shared -> one global atomic per (CTA, quantity)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) err_acc += __shfl_down_sync(0xffffffffu, err_acc, off);
    __shared__ long long s_err[kThreads / 32];
    if ((tid & 31) == 0) s_err[tid >> 5] = err_acc;
    __syncthreads();  // also orders every shared atomic before the flush below
    if (tid == 0) {
        long long e = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) e += s_err[w];
        if (e != 0) atomicAdd(out, (unsigned long long)e);
    }
    for (int k = tid; k < K; k += kThreads) {
        const unsigned int c = s_cnt[k];
        if (c) {
            atomicAdd(out + 1 + k, (unsigned long long)c);
            if (SUMS) {
                atomicAdd(out + 1 + K + 3 * k, s_sum[3 * k]);
                atomicAdd(out + 1 + K + 3 * k + 1, s_sum[3 * k + 1]);
                atomicAdd(out + 1 + K + 3 * k + 2, s_sum[3 * k + 2]);
            }
        }
    }
    if (VARIANT == 1) export_tail(p.tail, gridDim.x * gridDim.y, FUSED, seq_override);
}

template <int VARIANT, bool SRGB, bool SUMS, int IDXW>
__global__ void __launch_bounds__(kThreads, VARIANT == 3 ? HQ_V3_MIN_CTAS : (VARIANT == 1 ? HQ_V1_MIN_CTAS : 2)) assign_reduce_kernel(const __grid_constant__ AssignParams p) {
    assign_body<VARIANT, SRGB, SUMS, IDXW, false>(p, nullptr);
}
// one launch per evaluation: variant 1 with the palettes in parameter space; the last CTA exports the result words to pinned
// host memory and leaves them ZERO for the next launch (no clearing pass)
template <bool SRGB, bool SUMS, int NCOL>
__global__ void __launch_bounds__(kThreads, HQ_V1_MIN_CTAS) assign_small_kernel(const __grid_constant__ AssignParams p, const __grid_constant__ SmallPalettes<NCOL> pal) {
    assign_body<1, SRGB, SUMS, 0, true>(p, pal.v);
}

// PERSISTENT evaluator of a small search (hq_find_best_quantization with the LAB cost: thousands of identical evaluations whose
// cost is the launch itself — an empty kernel takes 8.8 us from launch to a flag in host memory on this box, the evaluation's own
// work ~6 us, profiles/r02/microbench5_launch_floor.json).  Launched ONCE per search with every CTA resident (cooperative launch);
// per iteration the host writes the palettes into pinned memory and then a sequence number into its mailbox word; CTA (0,0) polls that
// word over PCIe and republishes it in a device word the other CTAs poll in L2; every CTA then runs the one-launch evaluation body
// unchanged (palettes read fresh from the pinned buffer, last CTA exports the totals + the same sequence number and re-zeroes the
// result words).  kQuit, or no command for idle_ns (a host that went away, a progress callback that blocks), ends the kernel — it
// reports that in h_exit and the library falls back to one launch per evaluation.
struct PersistCtl {
    const unsigned long long* h_cmd;   // pinned host word: sequence number of the evaluation wanted (kPersistQuit: leave)
    unsigned long long* h_exit;        // pinned host word: non-zero once the kernel has left
    unsigned long long* d_cmd;         // device word: the command as seen by the dispatcher CTA
    const float4* h_pal;               // pinned host palettes [B * K] (r, g, b, .)
    float4* d_pal;                     // their device copy, refreshed by the dispatcher CTA before it republishes the command
    int ncolors;
    unsigned long long first_seq;      // the sequence number BEFORE the first evaluation
    unsigned long long idle_ns;
};
constexpr unsigned long long kPersistQuit = ~0ull;
template <bool SRGB, bool SUMS>
__global__ void __launch_bounds__(kThreads, HQ_V1_MIN_CTAS) assign_persist_kernel(const __grid_constant__ AssignParams p, const __grid_constant__ PersistCtl ctl) {
    __shared__ unsigned long long s_cmd;
    unsigned long long last = ctl.first_seq;
    const bool dispatcher = blockIdx.x == 0 && blockIdx.y == 0;
    for (;;) {
        if (threadIdx.x == 0) {
            unsigned long long v = last, t0 = 0;
            for (unsigned spin = 0;; ++spin) {
                if (dispatcher) v = *reinterpret_cast<const volatile unsigned long long*>(ctl.h_cmd);
                else asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctl.d_cmd) : "memory");
                if (v != last) break;
                if ((spin & 63u) == 63u) {
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    if (t0 == 0) t0 = t;
                    else if (t - t0 > (dispatcher ? ctl.idle_ns : 4 * ctl.idle_ns)) { v = kPersistQuit; break; }   // (the others outlast the dispatcher: they leave on ITS word)
                }
            }
            s_cmd = v;
        }
        __syncthreads();
        const unsigned long long cmd = s_cmd;
        if (dispatcher) {
            // the palettes cross PCIe ONCE per iteration, in a few coalesced 16-byte loads (every CTA reading its own colours from host
            // memory was ~10,000 four-byte PCIe reads per iteration: 138 us), then the command is republished for the other CTAs
            if (cmd != kPersistQuit) {
                __threadfence_system();   // the palettes the host wrote before the command are read after it
                for (int i = threadIdx.x; i < ctl.ncolors; i += blockDim.x) ctl.d_pal[i] = __ldcv(ctl.h_pal + i);
                __threadfence();
            }
            __syncthreads();
            if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(ctl.d_cmd), "l"(cmd) : "memory");
        }
        __syncthreads();
        if (cmd == kPersistQuit) break;
        last = cmd;
        assign_body<1, SRGB, SUMS, 0, true>(p, reinterpret_cast<const float*>(ctl.d_pal), cmd, true);
    }
    if (threadIdx.x == 0 && dispatcher) {
        *reinterpret_cast<volatile unsigned long long*>(ctl.h_exit) = last + 1;
        __threadfence_system();
    }
}

// per (kernel instantiation, device), PROCESS-wide: cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the function on
// the device, not to the calling thread, so it is only ever RAISED (two host threads with different palette sizes must not
// lower it under each other); the occupancy of the last launch shape is remembered beside it so that the per-iteration
// launches of a search pay neither cudaFuncSetAttribute nor an occupancy query
struct LaunchCache { std::mutex mu; size_t attr = 0; size_t smem = ~(size_t)0; int occ = 0; };
template <int VARIANT, bool SRGB, bool SUMS, int IDXW>
cudaError_t launch_assign_t(const AssignParams& p, int B, int sm_count, cudaStream_t stream) {
    auto kern = assign_reduce_kernel<VARIANT, SRGB, SUMS, IDXW>;
    const size_t smem = AssignSmem<VARIANT, SRGB, SUMS>(p.K8).total;
    static LaunchCache cache[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int occ;
    {
        LaunchCache& lc = cache[dev & 63];
        std::lock_guard<std::mutex> lock(lc.mu);
        if (smem > lc.attr) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            lc.attr = smem;
        }
        if (lc.smem != smem) {
            int o = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, smem);
            if (e != cudaSuccess) return e;
            lc.smem = smem; lc.occ = o < 1 ? 1 : o;
        }
        occ = lc.occ;
    }
    const long long slots = (long long)sm_count * occ;
    const long long ntiles = (long long)((p.n + AssignGeom<VARIANT>::kTile - 1) / AssignGeom<VARIANT>::kTile);
    // CTAs per candidate (G): B*G must be a whole number of waves of the `slots` resident CTAs,
    // otherwise the last wave leaves SMs idle (first measurement: 256 CTAs on 296 slots ->
    // 20 SMs half empty).  G = slots / gcd(B, slots) is the smallest such count; it is then
    // doubled while every CTA still gets >= 32 x 1024 pixels, which shortens the ragged tail.
    auto gcd = [](long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; };
    long long G = slots / gcd((long long)B, slots);
    while (G * 2 * (32 / AssignGeom<VARIANT>::kSub) <= ntiles && G * 2 * B <= slots * 16) G *= 2;   // (32 x 1024 pixels per CTA)
    if (VARIANT == 3) {   // bound the pixels per CTA (worklist capacity, above), in whole waves
        const long long unit = slots / gcd((long long)B, slots);
        const long long need = ((long long)p.n + kV3MaxPxPerCta - 1) / kV3MaxPxPerCta;
        if (G < need) G = (need + unit - 1) / unit * unit;
    }
    if (G > ntiles) G = ntiles;
    if (G < 1) G = 1;
    if (G > 65535) G = 65535;
    const dim3 grid((unsigned)B, (unsigned)G);
    kern<<<grid, kThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <bool SRGB, bool SUMS, int NCOL>
cudaError_t launch_small_t(const AssignParams& p, const float* h_palettes, int B, int sm_count, cudaStream_t stream) {
    auto kern = assign_small_kernel<SRGB, SUMS, NCOL>;
    SmallPalettes<NCOL> pal;
    std::memcpy(pal.v, h_palettes, (size_t)B * p.K * 4 * sizeof(float));
    const size_t smem = AssignSmem<1, SRGB, SUMS>(p.K8).total;
    static LaunchCache cache[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int occ;
    {
        LaunchCache& lc = cache[dev & 63];
        std::lock_guard<std::mutex> lock(lc.mu);
        if (smem > lc.attr) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            lc.attr = smem;
        }
        if (lc.smem != smem) {
            int o = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, smem);
            if (e != cudaSuccess) return e;
            lc.smem = smem; lc.occ = o < 1 ? 1 : o;
        }
        occ = lc.occ;
    }
    // a latency-bound launch: ONE wave at most, every CTA with at least two tiles (fewer CTAs = fewer result atomics and a
    // shorter ticket queue in front of the exporting CTA)
    const long long slots = (long long)sm_count * occ;
    const long long ntiles = (long long)((p.n + kTilePx - 1) / kTilePx);
    long long G = slots / B;
    if (G > (ntiles + 1) / 2) G = (ntiles + 1) / 2;
    if (G < 1) G = 1;
    if (G > 65535) G = 65535;
    const dim3 grid((unsigned)B, (unsigned)G);
    kern<<<grid, kThreads, smem, stream>>>(p, pal);
    return cudaGetLastError();
}

template <int VARIANT, bool SRGB, bool SUMS>
cudaError_t launch_assign_i(const AssignParams& p, int B, int idxw, int sm, cudaStream_t st) {
    if (idxw == 0) return launch_assign_t<VARIANT, SRGB, SUMS, 0>(p, B, sm, st);
    if (idxw == 1) return launch_assign_t<VARIANT, SRGB, SUMS, 1>(p, B, sm, st);
    return launch_assign_t<VARIANT, SRGB, SUMS, 2>(p, B, sm, st);
}
template <int VARIANT>
cudaError_t launch_assign_v(const AssignParams& p, int B, bool srgb, bool sums, int idxw, int sm, cudaStream_t st) {
    if (srgb) return sums ? launch_assign_i<VARIANT, true, true>(p, B, idxw, sm, st) : launch_assign_i<VARIANT, true, false>(p, B, idxw, sm, st);
    return sums ? launch_assign_i<VARIANT, false, true>(p, B, idxw, sm, st) : launch_assign_i<VARIANT, false, false>(p, B, idxw, sm, st);
}

// ====================================================================== final image
template <typename IdxT>
__global__ void apply_palette_kernel(const IdxT* __restrict__ idx, size_t n, const float4* __restrict__ pal,
                                     int K, uint8_t* __restrict__ out_rgb, float4* __restrict__ out_f32) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = (int)idx[i];
    if (k >= K) k = K - 1;
    const float4 c = __ldg(pal + k);
    if (out_f32) out_f32[i] = c;  // quantize() returns the palette colour itself (cl:168)
    if (out_rgb) {
        // HybridQuantization.java:122 (Icy convertToType(UBYTE, rescale), third-party):
        // the build defines float -> u8 as (int)(c * 255 + 0.5)
        out_rgb[3 * i] = (uint8_t)__float2int_rz(HQ_FADD(HQ_FMUL(c.x, 255.0f), 0.5f));
        out_rgb[3 * i + 1] = (uint8_t)__float2int_rz(HQ_FADD(HQ_FMUL(c.y, 255.0f), 0.5f));
        out_rgb[3 * i + 2] = (uint8_t)__float2int_rz(HQ_FADD(HQ_FMUL(c.z, 255.0f), 0.5f));
    }
}

__global__ void math_probe_kernel(int which, uint32_t first_bits, uint32_t count, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float v = __uint_as_float(first_bits + i);
    out[i] = hq_math_probe(which, v);
}

// FP32 ceiling probe: 8 independent FMA chains per thread, 8 warps per CTA, 4 CTAs per SM
template <bool PACKED>
__global__ void __launch_bounds__(kThreads) fp32_peak_kernel(int iters, float* __restrict__ out) {
    float s = 1.0f + 1e-7f * (float)threadIdx.x, t = 1e-9f * (float)blockIdx.x;
    float a[8];
    uint64_t A[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = (float)i; A[i] = pack2((float)i, (float)-i); }
    const uint64_t S = pack2(s, s), T = pack2(t, t);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (PACKED) A[i] = fma2(A[i], S, T);
                else a[i] = __fmaf_rn(a[i], s, t);
            }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; unpack2(A[i], lo, hi); r += a[i] + lo + hi; }
    if (r == 12345.678f) out[0] = r;
}

}  // namespace

cudaError_t launch_fp32_peak(bool packed, int iters, int sm_count, float* d_out, cudaStream_t stream) {
    if (packed) fp32_peak_kernel<true><<<sm_count * 4, kThreads, 0, stream>>>(iters, d_out);
    else fp32_peak_kernel<false><<<sm_count * 4, kThreads, 0, stream>>>(iters, d_out);
    return cudaGetLastError();
}

// ====================================================================== launchers
cudaError_t launch_decode_table(float* d_table, cudaStream_t stream) {
    decode_table_kernel<<<1, 256, 0, stream>>>(d_table);
    return cudaGetLastError();
}

// Result words -> the caller's pinned (mapped) host buffer, then a sequence number the host spins on: replaces the D2H copy
// node and the stream wait of a host-buffer evaluation (two of the four dependent stream operations of a search iteration).
__global__ void __launch_bounds__(256) export_results_kernel(const unsigned long long* __restrict__ src, unsigned long long* dst, size_t nwords,
                                                             volatile unsigned long long* flag, unsigned long long seq) {
    for (size_t i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();   // every thread's stores are visible to the host before ...
    __syncthreads();
    if (threadIdx.x == 0) *flag = seq;  // ... the flag is
}

cudaError_t launch_export_results(const unsigned long long* d_src, unsigned long long* h_dst_mapped, size_t nwords, unsigned long long* h_flag_mapped,
                                  unsigned long long seq, cudaStream_t stream) {
    export_results_kernel<<<1, 256, 0, stream>>>(d_src, h_dst_mapped, nwords, h_flag_mapped, seq);
    return cudaGetLastError();
}

// The exchange step as its own one-CTA launch, for the scoring kernels that carry no export tail (prefilter, pruned, the
// S-CIELAB chain, caller-owned buffers): all-reduce of the words over peer memory (hq_kernels.cuh), in place, and — when a
// host buffer is given — the export of the totals plus the sequence number, i.e. what ncclAllReduce + export_results_kernel did.
__global__ void __launch_bounds__(1024) peer_allreduce_kernel(const __grid_constant__ PeerExchange px, unsigned long long* words, unsigned nwords,
                                                              unsigned long long* host_dst, volatile unsigned long long* host_flag, unsigned long long host_seq) {
    peer_allreduce_cta(px, words, nwords, host_dst, false);
    if (host_dst) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *host_flag = host_seq;
    }
}

cudaError_t launch_peer_allreduce(const PeerExchange& px, unsigned long long* d_words, size_t nwords, unsigned long long* h_dst_mapped,
                                  unsigned long long* h_flag_mapped, unsigned long long host_seq, cudaStream_t stream) {
    if (px.nranks < 1 || px.nranks > kPeerMaxRanks || nwords > kPeerCapWords) return cudaErrorInvalidValue;
    const unsigned threads = nwords > 512 ? 1024u : (nwords > 128 ? 512u : 128u);
    peer_allreduce_kernel<<<1, threads, 0, stream>>>(px, d_words, (unsigned)nwords, h_dst_mapped, h_flag_mapped, host_seq);
    return cudaGetLastError();
}

cudaError_t launch_unit_to_lab(const float* d_unit, size_t n, size_t stride, int whitepoint, float* d_lab, unsigned int* d_bad,
                               int sm_count, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    size_t grid = (n + 255) / 256;
    const size_t cap = (size_t)(sm_count > 0 ? sm_count : 148) * 8;
    if (grid > cap) grid = cap;
    unit_to_lab_kernel<<<(unsigned)grid, 256, 0, stream>>>(d_unit, n, stride, whitepoint, d_lab, d_bad);
    return cudaGetLastError();
}

cudaError_t launch_rgb_to_lab(const uint8_t* d_rgb, size_t n, size_t stride, int whitepoint, const float* d_table,
                              float* d_lab, float* d_unit, int sm_count, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (n / kRlWarpPx >= 0xffffffffull) return cudaErrorInvalidValue;  // 32-bit tile counters
    const size_t ntiles = (n + kRlWarpPx - 1) / kRlWarpPx;
    size_t grid = (size_t)sm_count * 4;  // one wave of the 4 resident CTAs per SM (__launch_bounds__), warps grid-stride over 128-px tiles
    const size_t need = (ntiles + kThreads / 32 - 1) / (kThreads / 32);
    if (grid > need) grid = need;
    if (d_unit) rgb_to_lab_kernel<true><<<(unsigned)grid, kThreads, 0, stream>>>(d_rgb, n, stride, whitepoint, 1.0f, d_table, d_lab, d_unit);
    else rgb_to_lab_kernel<false><<<(unsigned)grid, kThreads, 0, stream>>>(d_rgb, n, stride, whitepoint, 1.0f, d_table, d_lab, d_unit);
    return cudaGetLastError();
}

cudaError_t launch_palette_features(const float* d_palettes, int B, int K, int whitepoint,
                                    float4* d_pal_lab, float4* d_pal_rgb, cudaStream_t stream, unsigned long long* d_zero, size_t zero_words) {
    const int K8 = padded_colors(K);
    const int total = B * K8;
    if (total == 0) return cudaSuccess;
    palette_features_kernel<<<(total + 127) / 128, 128, 0, stream>>>(d_palettes, B, K, K8, whitepoint, d_pal_lab, d_pal_rgb, d_zero, zero_words);
    return cudaGetLastError();
}

cudaError_t launch_assign_reduce(const AssignArgs& a, cudaStream_t stream) {
    if (a.B <= 0 || a.K <= 0 || a.K > kMaxColors) return cudaErrorInvalidValue;
    if (a.n == 0) return cudaSuccess;
    if (a.space == 1 && a.unit == nullptr) return cudaErrorInvalidValue;
    AssignParams p;
    p.feat = a.space == 1 ? a.unit : a.lab;
    p.lab = a.lab;
    p.n = a.n; p.stride = a.stride;
    p.pal_feat = a.space == 1 ? a.pal_rgb : a.pal_lab;
    p.pal_lab = a.pal_lab;
    p.K = a.K; p.K8 = padded_colors(a.K);
    p.words = result_words(a.K, a.want_sums);
    p.results = a.results;
    p.idx_out = a.idx_out;
    p.tail = a.tail;
    p.whitepoint = 0;
    {   // HQ_V1_TMA=1: the small-palette kernel stages its pixel tiles with TMA bulk copies + mbarriers instead of LDG.128.
        // OFF by default — measured on the B200 (tools/smallk_bench.py, profiles/r02/smallk_tma_ab.txt): the kernel is bound by
        // instruction issue, not by exposed load latency, and the ring costs instructions: K=8, 4K image, one candidate 41.7 -> 43.0 us,
        // 64 candidates 1.68 -> 2.01 ms.  Kept selectable so that the measurement can be repeated.
        static const int tma_env = [] { const char* e = std::getenv("HQ_V1_TMA"); return (e && e[0] == '1') ? 1 : 0; }();
        p.use_tma = tma_env;
    }
    // (an EMPTY own range — a shard that only carries halo rows — is honoured as empty: indices for every pixel, no reduction)
    p.own_lo = a.own_lo; p.own_hi = a.own_hi == kAllPixels ? a.n : a.own_hi;
    const int idxw = a.idx_out ? (a.K <= 256 ? 1 : 2) : 0;
    // |feature| bounds of the image for the prefilter's error bound: CIELAB of in-gamut sRGB has
    // L in [0,100], |a|,|b| < 128 (extremes 98.3 / 107.9); unit sRGB is in [0,1]
    if (a.space == 1) { p.xmax0 = p.xmax1 = p.xmax2 = 1.0f; }
    else { p.xmax0 = 100.5f; p.xmax1 = 128.0f; p.xmax2 = 128.0f; }
    int variant = a.variant;
    if (variant == 0) variant = (a.K <= kDirectMaxColors) ? 1 : 3;  // measured crossover (4K, 64 candidates): K=32 direct 42 % vs prefilter 36 %, K=64 46 % vs 50 %
    if (a.tail.active() && variant != 1) return cudaErrorInvalidValue;  // the export tail only exists in variant 1
    if (variant == 1) return launch_assign_v<1>(p, a.B, a.space == 1, a.want_sums, idxw, a.sm_count, stream);
    if (variant == 2) return launch_assign_v<2>(p, a.B, a.space == 1, a.want_sums, idxw, a.sm_count, stream);
    return launch_assign_v<3>(p, a.B, a.space == 1, a.want_sums, idxw, a.sm_count, stream);
}

cudaError_t launch_assign_small(const AssignArgs& a, const float* h_palettes, int whitepoint, cudaStream_t stream) {
    if (a.B <= 0 || a.K <= 0 || a.K > kDirectMaxColors || (long long)a.B * a.K > kSmallPalColors || a.idx_out || !a.tail.active() || !h_palettes) return cudaErrorInvalidValue;
    if (a.space == 1 && a.unit == nullptr) return cudaErrorInvalidValue;
    if (a.n == 0) return cudaErrorInvalidValue;   // (nothing would export: the caller handles an empty image)
    AssignParams p;
    p.feat = a.space == 1 ? a.unit : a.lab;
    p.lab = a.lab;
    p.n = a.n; p.stride = a.stride;
    p.pal_feat = nullptr; p.pal_lab = nullptr;
    p.K = a.K; p.K8 = padded_colors(a.K);
    p.words = result_words(a.K, a.want_sums);
    p.results = a.results;
    p.idx_out = nullptr;
    p.tail = a.tail;
    p.use_tma = 0;
    p.whitepoint = whitepoint;
    p.xmax0 = p.xmax1 = p.xmax2 = 0.f;
    p.own_lo = a.own_lo; p.own_hi = a.own_hi == kAllPixels ? a.n : a.own_hi;
    const int ncol = a.B * a.K;
    auto go = [&](auto srgb, auto sums) -> cudaError_t {
        constexpr bool S = decltype(srgb)::value, U = decltype(sums)::value;
        if (ncol <= 32) return launch_small_t<S, U, 32>(p, h_palettes, a.B, a.sm_count, stream);
        if (ncol <= 64) return launch_small_t<S, U, 64>(p, h_palettes, a.B, a.sm_count, stream);
        if (ncol <= 128) return launch_small_t<S, U, 128>(p, h_palettes, a.B, a.sm_count, stream);
        return launch_small_t<S, U, kSmallPalColors>(p, h_palettes, a.B, a.sm_count, stream);
    };
    if (a.space == 1) return a.want_sums ? go(std::true_type{}, std::true_type{}) : go(std::true_type{}, std::false_type{});
    return a.want_sums ? go(std::false_type{}, std::true_type{}) : go(std::false_type{}, std::false_type{});
}

cudaError_t launch_assign_persist(const AssignArgs& a, const float* h_palettes_mapped, float* d_palettes, int whitepoint, const unsigned long long* h_cmd,
                                  unsigned long long* h_exit, unsigned long long* d_cmd, unsigned long long first_seq, unsigned long long idle_ns, cudaStream_t stream) {
    if (a.B <= 0 || a.K <= 0 || a.K > kDirectMaxColors || (long long)a.B * a.K > kSmallPalColors || a.idx_out || !a.tail.active() || !h_palettes_mapped || !d_palettes || a.n == 0)
        return cudaErrorInvalidValue;
    if (a.space == 1 && a.unit == nullptr) return cudaErrorInvalidValue;
    AssignParams p;
    p.feat = a.space == 1 ? a.unit : a.lab;
    p.lab = a.lab;
    p.n = a.n; p.stride = a.stride;
    p.pal_feat = nullptr; p.pal_lab = nullptr;
    p.K = a.K; p.K8 = padded_colors(a.K);
    p.words = result_words(a.K, a.want_sums);
    p.results = a.results;
    p.idx_out = nullptr;
    p.tail = a.tail;
    p.use_tma = 0;
    p.whitepoint = whitepoint;
    p.xmax0 = p.xmax1 = p.xmax2 = 0.f;
    p.own_lo = a.own_lo; p.own_hi = a.own_hi == kAllPixels ? a.n : a.own_hi;
    PersistCtl ctl;
    ctl.h_cmd = h_cmd; ctl.h_exit = h_exit; ctl.d_cmd = d_cmd; ctl.first_seq = first_seq; ctl.idle_ns = idle_ns;
    ctl.h_pal = reinterpret_cast<const float4*>(h_palettes_mapped); ctl.d_pal = reinterpret_cast<float4*>(d_palettes); ctl.ncolors = a.B * a.K;
    auto go = [&](auto srgb, auto sums) -> cudaError_t {
        constexpr bool S = decltype(srgb)::value, U = decltype(sums)::value;
        auto kern = assign_persist_kernel<S, U>;
        const size_t smem = AssignSmem<1, S, U>(p.K8).total;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        // the same grid as the one-launch evaluation: one wave at most (every CTA must be resident: they wait for each other's tickets)
        const long long slots = (long long)a.sm_count * occ;
        const long long ntiles = (long long)((p.n + kTilePx - 1) / kTilePx);
        long long G = slots / a.B;
        if (G > (ntiles + 1) / 2) G = (ntiles + 1) / 2;
        if (G < 1) G = 1;
        if (G > 65535) G = 65535;
        if ((long long)a.B * G > slots) return cudaErrorLaunchOutOfResources;
        void* args[] = {(void*)&p, (void*)&ctl};
        return cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)a.B, (unsigned)G), dim3(kThreads), args, smem, stream);
    };
    if (a.space == 1) return a.want_sums ? go(std::true_type{}, std::true_type{}) : go(std::true_type{}, std::false_type{});
    return a.want_sums ? go(std::false_type{}, std::true_type{}) : go(std::false_type{}, std::false_type{});
}

cudaError_t launch_apply_palette(const void* d_idx, bool idx16, size_t n, const float* d_palette, int K,
                                 uint8_t* d_out_rgb, float* d_out_f32, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (idx16)
        apply_palette_kernel<uint16_t><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(d_idx), n, reinterpret_cast<const float4*>(d_palette), K, d_out_rgb, reinterpret_cast<float4*>(d_out_f32));
    else
        apply_palette_kernel<uint8_t><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint8_t*>(d_idx), n, reinterpret_cast<const float4*>(d_palette), K, d_out_rgb, reinterpret_cast<float4*>(d_out_f32));
    return cudaGetLastError();
}

cudaError_t launch_math_probe(int which, uint32_t first_bits, uint32_t count, float* d_out, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    math_probe_kernel<<<(count + 255) / 256, 256, 0, stream>>>(which, first_bits, count, d_out);
    return cudaGetLastError();
}

}  // namespace hq
