// hq_scielab.cu — S-CIELAB spatial-filter stage (scope table row "next 1"): the plugin's real
// cost is mean CIE76 between S-CIELAB(original) and S-CIELAB(quantised).
//   original : RGB2XYZ + XYZ2Opp (cl:79-90,111-116) -> 3 separable filter pairs (convolve4Channels /
//              convolve1Channel, cl:2-74, sequenced at ImageManipulation.java:316-346) -> Opp2LAB
//   candidate: quantizeAndConvertToOpp (cl:172-199) -> computeScielabKernelsTemp (cl:234-272) ->
//              computeScielabKernelsEnd (cl:274-306) -> Opp2LAB (cl:124-145) -> CIEDE (cl:201-209)
// First version: one thread per pixel and per pass, taps read in ascending order so that every fma
// chain has the reference's order (bit-exact against the oracle).  The reference transposes between
// the passes; here the vertical pass reads column neighbours directly (coalesced across a row), which
// is the same arithmetic.
#include <cmath>
#include <mutex>

#include "hq_kernels.cuh"
#include "hq_math.h"

namespace hq {
namespace {

constexpr int kScThreads = 256;

// filter block in global memory: [0,3T) k1[t][3], [3T,6T) k2[t][3], [6T,7T) k3[t], [7T,8T) |k3|[t]
struct ScFilters {
    const float* data;
    int taps;
};

// dE of one pixel pair in 2^-24 fixed point: CIE76 (cl:209) or the CIE94 branch (cl:217-226, ImageManipulation.deltaETypes :20).
// A NaN of the CIE94 branch cannot enter an integer sum: it is COUNTED (*nan_out) and the caller reports the whole sum as NaN.
__device__ __forceinline__ long long sc_delta_e_fx(int de_type, float L1, float a1, float b1, float L2, float a2, float b2,
                                                   unsigned long long* nan_out) {
    if (de_type == 0) return hq_to_fx(HQ_FSQRT(hq_dist2(L1, a1, b1, L2, a2, b2)));
    const float e = hq_cl_delta_e94(L1, a1, b1, L2, a2, b2);
    if (e != e) { atomicAdd(nan_out, 1ull); return 0; }
    return hq_to_fx(e);
}

__global__ void sc_rgb_to_opp_kernel(const uint8_t* __restrict__ rgb, size_t n, size_t stride,
                                     const float* __restrict__ table, float* __restrict__ opp) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float R = __ldg(table + 256 + rgb[3 * i]), G = __ldg(table + 256 + rgb[3 * i + 1]), B = __ldg(table + 256 + rgb[3 * i + 2]);
    const hq_float3 o = hq_cl_linrgb_to_opp_via_xyz(R, G, B);
    opp[i] = o.x; opp[stride + i] = o.y; opp[2 * stride + i] = o.z;
}

// the same for a planar float image: decoded per value, as the reference's RGB2XYZ kernel does (cl:79-90)
// (unit and opp may be the same buffer: element-wise; bad, when given, reports values outside [0,1])
__global__ void sc_unit_to_opp_kernel(const float* unit, size_t n, size_t stride, float* opp, unsigned int* __restrict__ bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float r = unit[i], g = unit[stride + i], b = unit[2 * stride + i];
    hq_float3 o; o.x = o.y = o.z = 0.f;
    if (!bad || (r >= 0.f && r <= 1.f && g >= 0.f && g <= 1.f && b >= 0.f && b <= 1.f)) o = hq_cl_linrgb_to_opp_via_xyz(hq_srgb_decode(r), hq_srgb_decode(g), hq_srgb_decode(b));
    else atomicOr(bad, 1u);
    opp[i] = o.x; opp[stride + i] = o.y; opp[2 * stride + i] = o.z;
}

// palette colours -> opponent table (the K values the quantised image can take)
// (optionally clears zero[0 .. zero_words): the error sums of the evaluation that follows on the same stream)
__global__ void sc_palette_opp_kernel(const float* __restrict__ pal, int total, float4* __restrict__ tab, unsigned long long* __restrict__ zero, int zero_words) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < zero_words) zero[i] = 0ull;
    if (i >= total) return;
    const float4 c = reinterpret_cast<const float4*>(pal)[i];
    const hq_float3 o = hq_cl_linrgb_to_opp(hq_srgb_decode(c.x), hq_srgb_decode(c.y), hq_srgb_decode(c.z));
    tab[i] = make_float4(o.x, o.y, o.z, 0.f);
}

// horizontal pass.  MODE 0: input = opponent planes; MODE 1: input = palette index image + table.
// Output: 7 planes [7][stride]: (k1 * O1, k1 * O2, k1 * O3, k2 * O1, k2 * O2, k2 * O3, k3 * O1).
template <int MODE, typename IdxT>
__global__ void __launch_bounds__(kScThreads)
sc_hpass_kernel(const float* __restrict__ opp, const IdxT* __restrict__ idx, const float4* __restrict__ tab,
                int w, int h, size_t stride, ScFilters f, float* __restrict__ tmp) {
    extern __shared__ float s_f[];  // 7 * taps
    for (int i = threadIdx.x; i < 7 * f.taps; i += kScThreads) s_f[i] = f.data[i];
    __syncthreads();
    const int x = blockIdx.x * kScThreads + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int half = f.taps / 2;
    const size_t row = (size_t)y * w;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, c0 = 0.f;
    for (int t = 0; t < f.taps; ++t) {
        const size_t p = row + hq_reflect(x + t - half, w);
        float i0, i1, i2;
        if (MODE == 0) { i0 = __ldg(opp + p); i1 = __ldg(opp + stride + p); i2 = __ldg(opp + 2 * stride + p); }
        else { const float4 v = __ldg(tab + idx[p]); i0 = v.x; i1 = v.y; i2 = v.z; }
        a0 = HQ_FFMA(i0, s_f[3 * t], a0); a1 = HQ_FFMA(i1, s_f[3 * t + 1], a1); a2 = HQ_FFMA(i2, s_f[3 * t + 2], a2);
        b0 = HQ_FFMA(i0, s_f[3 * f.taps + 3 * t], b0); b1 = HQ_FFMA(i1, s_f[3 * f.taps + 3 * t + 1], b1); b2 = HQ_FFMA(i2, s_f[3 * f.taps + 3 * t + 2], b2);
        c0 = HQ_FFMA(i0, s_f[6 * f.taps + t], c0);
    }
    const size_t o = row + x;
    tmp[o] = a0; tmp[stride + o] = a1; tmp[2 * stride + o] = a2;
    tmp[3 * stride + o] = b0; tmp[4 * stride + o] = b1; tmp[5 * stride + o] = b2; tmp[6 * stride + o] = c0;
}

// vertical pass + combination + Opp2LAB.
// MODE 0 (original, ImageManipulation.java:319-346): conv = V1(H1) ; conv += V2(H2) ; conv.x += V|3|(H3),
//         each V a full tap sum started from 0; writes the S-CIELAB planes.
// MODE 1 (candidate, cl:292-304): out = fma(t1,k1, fma(t2,k2,out)); out.x = fma(t3,|k3|,out.x) per tap;
//         dE vs the original's S-CIELAB, fixed-point sum -> one atomic per CTA.
template <int MODE>
__global__ void __launch_bounds__(kScThreads)
sc_vpass_kernel(const float* __restrict__ tmp, int w, int h, size_t stride, ScFilters f, hq_float3 ill, ScRows rows,
                float* __restrict__ lab_out, const float* __restrict__ lab_orig, unsigned long long* __restrict__ err_out,
                int de_type, unsigned long long* __restrict__ nan_out) {
    extern __shared__ float s_f[];  // 8 * taps
    for (int i = threadIdx.x; i < 8 * f.taps; i += kScThreads) s_f[i] = f.data[i];
    __syncthreads();
    const int x = blockIdx.x * kScThreads + threadIdx.x, y = rows.y_begin + blockIdx.y;
    long long fx = 0;
    if (x < w) {
        const int half = f.taps / 2, T = f.taps;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, c0 = 0.f;
        for (int t = 0; t < T; ++t) {
            const size_t p = (size_t)(hq_reflect(rows.g0 + y + t - half, rows.gh) - rows.g0) * w + x;
            const float t10 = __ldg(tmp + p), t11 = __ldg(tmp + stride + p), t12 = __ldg(tmp + 2 * stride + p);
            const float t20 = __ldg(tmp + 3 * stride + p), t21 = __ldg(tmp + 4 * stride + p), t22 = __ldg(tmp + 5 * stride + p);
            const float t3 = __ldg(tmp + 6 * stride + p);
            if (MODE == 0) {
                a0 = HQ_FFMA(t10, s_f[3 * t], a0); a1 = HQ_FFMA(t11, s_f[3 * t + 1], a1); a2 = HQ_FFMA(t12, s_f[3 * t + 2], a2);
                b0 = HQ_FFMA(t20, s_f[3 * T + 3 * t], b0); b1 = HQ_FFMA(t21, s_f[3 * T + 3 * t + 1], b1); b2 = HQ_FFMA(t22, s_f[3 * T + 3 * t + 2], b2);
                c0 = HQ_FFMA(t3, s_f[7 * T + t], c0);
            } else {
                a0 = HQ_FFMA(t10, s_f[3 * t], HQ_FFMA(t20, s_f[3 * T + 3 * t], a0));
                a1 = HQ_FFMA(t11, s_f[3 * t + 1], HQ_FFMA(t21, s_f[3 * T + 3 * t + 1], a1));
                a2 = HQ_FFMA(t12, s_f[3 * t + 2], HQ_FFMA(t22, s_f[3 * T + 3 * t + 2], a2));
                a0 = HQ_FFMA(t3, s_f[7 * T + t], a0);
            }
        }
        float o0, o1, o2;
        if (MODE == 0) { o0 = HQ_FADD(HQ_FADD(a0, b0), c0); o1 = HQ_FADD(a1, b1); o2 = HQ_FADD(a2, b2); }
        else { o0 = a0; o1 = a1; o2 = a2; }
        const hq_float3 lab = hq_cl_opp_to_lab(o0, o1, o2, ill);
        const size_t p = (size_t)y * w + x;
        if (MODE == 0) {
            lab_out[p] = lab.x; lab_out[stride + p] = lab.y; lab_out[2 * stride + p] = lab.z;
        } else {
            fx = sc_delta_e_fx(de_type, __ldg(lab_orig + p), __ldg(lab_orig + stride + p), __ldg(lab_orig + 2 * stride + p), lab.x, lab.y, lab.z, nan_out);
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) fx += __shfl_down_sync(0xffffffffu, fx, off);
        __shared__ long long s_err[kScThreads / 32];
        if ((threadIdx.x & 31) == 0) s_err[threadIdx.x >> 5] = fx;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long e = 0;
#pragma unroll
            for (int i = 0; i < kScThreads / 32; ++i) e += s_err[i];
            if (e) atomicAdd(err_out, (unsigned long long)e);
        }
    }
}

// ---------------------------------------------------------------- specialised kernels, taps = 21
// (the plugin's default geometry: dpi 72, 45 cm).  Same arithmetic and tap order as the generic
// kernels above; the filter bank travels as a kernel parameter so every coefficient is a constant-bank
// operand of its FFMA, inputs are staged once in shared memory (horizontal) or streamed through a
// register sliding window (vertical), so each value is loaded ~1.3x / 3.5x instead of 21x.
constexpr int kT = 21, kHalf = 10;
struct Filt21 { float v[8 * kT]; };  // k1[t][3], k2[t][3], k3[t], |k3|[t]

constexpr int kHOut = 4;                          // outputs per thread (horizontal)
constexpr int kHSeg = kScThreads * kHOut;         // 1024 pixels of one row per CTA

template <int MODE, typename IdxT>
__global__ void __launch_bounds__(kScThreads)
sc_hpass21_kernel(const float* __restrict__ opp, const IdxT* __restrict__ idx, const float4* __restrict__ tab,
                  int w, int h, size_t stride, const __grid_constant__ Filt21 f, float* __restrict__ tmp) {
    __shared__ float s_in[3][kHSeg + 2 * kHalf];
    const int x0 = blockIdx.x * kHSeg, y = blockIdx.y;
    const size_t row = (size_t)y * w;
    for (int i = threadIdx.x; i < kHSeg + 2 * kHalf; i += kScThreads) {
        const int xx = x0 - kHalf + i;
        float i0 = 0.f, i1 = 0.f, i2 = 0.f;
        if (xx < w + kHalf) {  // positions beyond the last output's right halo are never used
            const size_t p = row + hq_reflect(xx, w);
            if (MODE == 0) { i0 = __ldg(opp + p); i1 = __ldg(opp + stride + p); i2 = __ldg(opp + 2 * stride + p); }
            else { const float4 v = __ldg(tab + idx[p]); i0 = v.x; i1 = v.y; i2 = v.z; }
        }
        s_in[0][i] = i0; s_in[1][i] = i1; s_in[2][i] = i2;
    }
    __syncthreads();
    const int lx = threadIdx.x * kHOut;
    if (x0 + lx >= w) return;
    float acc[kHOut][7];
#pragma unroll
    for (int o = 0; o < kHOut; ++o)
#pragma unroll
        for (int c = 0; c < 7; ++c) acc[o][c] = 0.f;
    float win[3][kHOut + kT - 1];
#pragma unroll
    for (int i = 0; i < kHOut + kT - 1; ++i) { win[0][i] = s_in[0][lx + i]; win[1][i] = s_in[1][lx + i]; win[2][i] = s_in[2][lx + i]; }
#pragma unroll
    for (int t = 0; t < kT; ++t)
#pragma unroll
        for (int o = 0; o < kHOut; ++o) {
            const float i0 = win[0][o + t], i1 = win[1][o + t], i2 = win[2][o + t];
            acc[o][0] = HQ_FFMA(i0, f.v[3 * t], acc[o][0]); acc[o][1] = HQ_FFMA(i1, f.v[3 * t + 1], acc[o][1]); acc[o][2] = HQ_FFMA(i2, f.v[3 * t + 2], acc[o][2]);
            acc[o][3] = HQ_FFMA(i0, f.v[3 * kT + 3 * t], acc[o][3]); acc[o][4] = HQ_FFMA(i1, f.v[3 * kT + 3 * t + 1], acc[o][4]); acc[o][5] = HQ_FFMA(i2, f.v[3 * kT + 3 * t + 2], acc[o][5]);
            acc[o][6] = HQ_FFMA(i0, f.v[6 * kT + t], acc[o][6]);
        }
    const size_t o0 = row + x0 + lx;
    if (x0 + lx + kHOut <= w && ((row + x0 + lx) & 3) == 0) {
#pragma unroll
        for (int c = 0; c < 7; ++c)
            *reinterpret_cast<float4*>(tmp + c * stride + o0) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
    } else {
#pragma unroll
        for (int o = 0; o < kHOut; ++o)
            if (x0 + lx + o < w)
#pragma unroll
                for (int c = 0; c < 7; ++c) tmp[c * stride + o0 + o] = acc[o][c];
    }
}

constexpr int kVRows = 8;  // output rows per thread (vertical): 28 input rows stream through registers

template <int MODE>
__global__ void __launch_bounds__(kScThreads)
sc_vpass21_kernel(const float* __restrict__ tmp, int w, int h, size_t stride, const __grid_constant__ Filt21 f, hq_float3 ill, ScRows rows,
                  float* __restrict__ lab_out, const float* __restrict__ lab_orig, unsigned long long* __restrict__ err_out,
                  int de_type, unsigned long long* __restrict__ nan_out) {
    const int x = blockIdx.x * kScThreads + threadIdx.x, y0 = rows.y_begin + blockIdx.y * kVRows;
    long long fx = 0;
    if (x < w) {
        float a[kVRows][3], b[kVRows][3], c3[kVRows];
#pragma unroll
        for (int o = 0; o < kVRows; ++o) { a[o][0] = a[o][1] = a[o][2] = 0.f; b[o][0] = b[o][1] = b[o][2] = 0.f; c3[o] = 0.f; }
#pragma unroll
        for (int r = 0; r < kVRows + kT - 1; ++r) {
            // reflection at the GLOBAL image borders; rows only needed by discarded outputs are clamped into the local array
            int lr = hq_reflect(rows.g0 + y0 - kHalf + r, rows.gh) - rows.g0;
            lr = lr < 0 ? 0 : (lr >= h ? h - 1 : lr);
            const size_t p = (size_t)lr * w + x;
            const float t10 = __ldg(tmp + p), t11 = __ldg(tmp + stride + p), t12 = __ldg(tmp + 2 * stride + p);
            const float t20 = __ldg(tmp + 3 * stride + p), t21 = __ldg(tmp + 4 * stride + p), t22 = __ldg(tmp + 5 * stride + p);
            const float t3 = __ldg(tmp + 6 * stride + p);
#pragma unroll
            for (int o = 0; o < kVRows; ++o) {
                const int t = r - o;  // tap index of input row r for output row o: ascending in r, i.e. the reference's order
                if (t >= 0 && t < kT) {
                    if (MODE == 0) {
                        a[o][0] = HQ_FFMA(t10, f.v[3 * t], a[o][0]); a[o][1] = HQ_FFMA(t11, f.v[3 * t + 1], a[o][1]); a[o][2] = HQ_FFMA(t12, f.v[3 * t + 2], a[o][2]);
                        b[o][0] = HQ_FFMA(t20, f.v[3 * kT + 3 * t], b[o][0]); b[o][1] = HQ_FFMA(t21, f.v[3 * kT + 3 * t + 1], b[o][1]); b[o][2] = HQ_FFMA(t22, f.v[3 * kT + 3 * t + 2], b[o][2]);
                        c3[o] = HQ_FFMA(t3, f.v[7 * kT + t], c3[o]);
                    } else {
                        a[o][0] = HQ_FFMA(t10, f.v[3 * t], HQ_FFMA(t20, f.v[3 * kT + 3 * t], a[o][0]));
                        a[o][1] = HQ_FFMA(t11, f.v[3 * t + 1], HQ_FFMA(t21, f.v[3 * kT + 3 * t + 1], a[o][1]));
                        a[o][2] = HQ_FFMA(t12, f.v[3 * t + 2], HQ_FFMA(t22, f.v[3 * kT + 3 * t + 2], a[o][2]));
                        a[o][0] = HQ_FFMA(t3, f.v[7 * kT + t], a[o][0]);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < kVRows; ++o) {
            const int y = y0 + o;
            if (y < rows.y_begin + rows.y_count) {
                float o0, o1, o2;
                if (MODE == 0) { o0 = HQ_FADD(HQ_FADD(a[o][0], b[o][0]), c3[o]); o1 = HQ_FADD(a[o][1], b[o][1]); o2 = HQ_FADD(a[o][2], b[o][2]); }
                else { o0 = a[o][0]; o1 = a[o][1]; o2 = a[o][2]; }
                const hq_float3 lab = hq_cl_opp_to_lab(o0, o1, o2, ill);
                const size_t p = (size_t)y * w + x;
                if (MODE == 0) {
                    lab_out[p] = lab.x; lab_out[stride + p] = lab.y; lab_out[2 * stride + p] = lab.z;
                } else {
                    fx += sc_delta_e_fx(de_type, __ldg(lab_orig + p), __ldg(lab_orig + stride + p), __ldg(lab_orig + 2 * stride + p), lab.x, lab.y, lab.z, nan_out);
                }
            }
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) fx += __shfl_down_sync(0xffffffffu, fx, off);
        __shared__ long long s_err[kScThreads / 32];
        if ((threadIdx.x & 31) == 0) s_err[threadIdx.x >> 5] = fx;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long e = 0;
#pragma unroll
            for (int i = 0; i < kScThreads / 32; ++i) e += s_err[i];
            if (e) atomicAdd(err_out, (unsigned long long)e);
        }
    }
}

// ---------------------------------------------------------------- fused candidate stage, taps = 21 (round 2)
// One kernel per population instead of two per candidate with a 7-plane fp32 intermediate in HBM (28 B/px written and
// read back ~3.5x: ~510 MB per 4K candidate where the stage needs the 1 B/px of indices and the 12 B/px of the original's
// S-CIELAB).  A CTA owns a vertical STRIP of kSW columns of one candidate and walks down it in bands of kSBand rows; the
// horizontally filtered rows live in a ring of kSRing = kSBand + 20 rows in shared memory (7 planes), so every image row
// is filtered horizontally ONCE per strip (plus 20 rows at each segment start):
//   H phase (cl:234-272): task = (row, 8 outputs): 28 palette indices from HBM/L2 -> opponent colours by look-up in the
//       candidate's K-entry table (cl:194-198, shared memory) -> 7 x 21 fma chains per output in ascending tap order
//       -> ring;
//   V phase (cl:274-306): task = (column, 8 output rows): 28 ring rows stream through registers, per tap
//       fma(t1,k1,fma(t2,k2,out)); out.x = fma(t3,|k3|,out.x) -> Opp2LAB (cl:124-145) -> CIE76 against the resident
//       S-CIELAB of the original (cl:209, its 24 values prefetched before the fma work) -> 2^-24 fixed point.
// One atomic per CTA.  Arithmetic, operand order and fma nesting are those of sc_hpass21_kernel / sc_vpass21_kernel<1>:
// same bits (tests/test_gpu_scielab.py holds all three paths and the oracle together).
// The first version of the fusion (a 32 x 128 tile with all 148 filtered rows in 145 KB of shared memory, 1 CTA/SM) ran
// 1.8x SLOWER than the two kernels it replaced: ncu showed 28 % issue utilisation at 2 warps per scheduler, 23 % of the
// stall samples on the index staging loop and 17 % on the S-CIELAB loads of the epilogue (profiles/r02).
// Second version (this one): strips + a TRANSPOSED vertical filter.  A thread owns one column and streams down its rows;
// input row v updates the 21 pending outputs o = v+10-t with tap t (ascending t per output = the reference's order) and
// completes output v-10, so only the CURRENT band of horizontally filtered rows has to sit in shared memory (57 KB for
// 128 columns -> 3 CTAs/SM) and each ring value is read from shared memory once instead of 3.5 times.  Code size matters:
// a fully unrolled 8-row register window per phase (38 KB of FFMAs, twice for the prologue) stalled 1.7 cycles per issue
// on instruction fetch (L1.5 I-cache: 32 KB); here the H task is 4 outputs + a window shift, looped, and the V loop is
// unrolled by 4 rows with one accumulator shift per 4 rows.
// The two phases have different register needs (the V accumulators persist for the whole strip, the H window is 84 floats),
// so they run in different WARPS: warps 0-3 filter vertically (one column each thread), warps 4-7 filter horizontally one
// band of 8 rows ahead, through a double-buffered ring and four named barriers (full / empty per buffer).
#ifndef HQ_SC_VUNROLL
#define HQ_SC_VUNROLL 4   // rows per unrolled V group (4 or 8): one accumulator shift per group vs code size
#endif
#ifndef HQ_SC_HHALVES
#define HQ_SC_HHALVES 2   // 2: the H task is 4 outputs + a window shift, looped twice; 1: 8 outputs fully unrolled
#endif
#ifndef HQ_SC_SW
#define HQ_SC_SW 128      // columns per strip (128: 2 CTAs of 256 threads per SM; 256: 1 CTA of 512 threads, half the horizontal halo overhead)
#endif
#ifndef HQ_SC_EPI_H
#define HQ_SC_EPI_H 0     // who runs Opp2LAB + dE: 0 the V warps (every row; default), 1 the H warps (every row, one band later, through
#endif                    // shared memory), 2 both: of every group of four rows the V warps score two in place and hand two to the H
                          // warps.  EXPERIMENT: per pixel the V warps execute ~342 instructions and the H warps ~200, so 1 and 2 were
                          // built to even the halves out (2: 289 / 275) — identical integers, and SLOWER: 0.285 -> 0.318 (1) / 0.336 (2)
                          // ms per 4K candidate (profiles/r02/sc_stage_ab_variants.txt).  The H warps' slack is not what limits the kernel.
constexpr int kSW = HQ_SC_SW, kSBand = 8, kSThreads = 2 * kSW, kSCtasPerSm = 256 / kSW;
constexpr int kEpiV = HQ_SC_EPI_H == 0 ? 4 : (HQ_SC_EPI_H == 1 ? 0 : 2);   // rows of each group of four that the V warps score themselves
constexpr int kEpiHRows = kSBand * (4 - kEpiV) / 4;                       // rows per band handed to the H warps
constexpr size_t kSOutFloats = (size_t)2 * kEpiHRows * 3 * kSW;           // their vertically filtered opponent values: [2][kEpiHRows][3][kSW]
constexpr int kSHOut = 8;                                              // outputs per H task
constexpr int kSHPer = kSHOut / HQ_SC_HHALVES;                         // ... per unrolled pass
constexpr int kSVUnroll = HQ_SC_VUNROLL;
static_assert(kSBand % kSVUnroll == 0 && kSBand * (kSW / kSHOut) == kSThreads / 2 && kSW == kSThreads / 2, "task geometry");
static_assert(HQ_SC_EPI_H == 0 || kSVUnroll == 4, "the epilogue split is per group of four rows");

// named barriers with IMMEDIATE ids (a register id makes ptxas reserve all 16 barriers for the CTA)
template <int ID> __device__ __forceinline__ void named_bar_sync_i(int count) { asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(count) : "memory"); }
template <int ID> __device__ __forceinline__ void named_bar_arrive_i(int count) { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(count) : "memory"); }
template <int ID0> __device__ __forceinline__ void named_bar_sync(int which, int count) { if (which) named_bar_sync_i<ID0 + 1>(count); else named_bar_sync_i<ID0>(count); }
template <int ID0> __device__ __forceinline__ void named_bar_arrive(int which, int count) { if (which) named_bar_arrive_i<ID0 + 1>(count); else named_bar_arrive_i<ID0>(count); }

// Opp2LAB (cl:124-145) of one pixel in STRAIGHT-LINE form.  hq_cl_opp_to_lab_white branches nine times per pixel (the range
// check of each constant division, the dark-value segment and the fp64 decision of each cube root), and every one of those
// branches cost the V warps a BSSY / BRA / BSYNC triple and a basic-block boundary the scheduler cannot move loads across
// (ncu, round 2: 25 control instructions per pixel, 0.86 stall cycles per issue waiting for instructions).  Here the common
// case of every branch is computed unconditionally — the same operations in the same order, so the same bits — and ONE flag
// collects "some value needs another branch" (a value outside the constant division's verified range, a value on the linear
// segment, a cube root the fp32 path cannot decide): such a pixel is redone by the reference routine, out of line.
#ifndef HQ_SC_STRAIGHT
#define HQ_SC_STRAIGHT 1
#endif
__device__ __noinline__ hq_float3 sc_opp_to_lab_general(float o0, float o1, float o2, hq_white w) { return hq_cl_opp_to_lab_white(o0, o1, o2, w); }
__device__ __forceinline__ float sc_cbrt_common(float t, bool& other) {   // hq_cbrtf_fast up to its decision
    float lg, y0, rq;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(t));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(__fmul_rn(lg, 0x1.555556p-2f)));
    const float q = __fmul_rn(y0, y0);
    const float nql = __fmaf_rn(-y0, y0, q);
    const float p = __fmul_rn(q, y0);
    const float npl = __fmaf_rn(q, -y0, p);
    const float r = __fmaf_rn(nql, y0, __fadd_rn(__fsub_rn(t, p), npl));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rq) : "f"(q));
    const float d = __fmul_rn(__fmul_rn(r, 0x1.555556p-2f), rq);
    const float s_lo = __fadd_rn(y0, __fmaf_rn(y0, -HQ_CBRT_ETA, d));
    const float s_hi = __fadd_rn(y0, __fmaf_rn(y0, HQ_CBRT_ETA, d));
    other = other || !(s_lo == s_hi) || !(t > HQ_LABDELTA3);   // (a NaN from a negative t lands here too)
    return s_lo;
}
__device__ __forceinline__ float sc_div_const_common(float x, float c, float rc, bool& other) {   // hq_div_const inside its verified range
    const uint32_t ax = __float_as_uint(x) & 0x7fffffffu;
    other = other || !(ax - 0x21800000u <= 0x40800000u - 0x21800000u);
    const float q0 = __fmul_rn(x, rc);
    const float rem = __fmaf_rn(-q0, c, x);
    return __fmaf_rn(rem, rc, q0);
}
__device__ __forceinline__ hq_float3 sc_opp_to_lab_straight(float o0, float o1, float o2, const hq_white& w) {
#if HQ_SC_STRAIGHT
    const float X = hq_cl_dot3(0.624045f, -1.87044f, -0.155304f, o0, o1, o2);
    const float Y = hq_cl_dot3(1.36606f, 0.931563f, 0.433903f, o0, o1, o2);
    const float Z = hq_cl_dot3(1.5013f, 1.41761f, 2.53307f, o0, o1, o2);
    bool other = false;
    const float fx = sc_cbrt_common(sc_div_const_common(X, w.x, w.rx, other), other);
    const float fy = sc_cbrt_common(Y, other);
    const float fz = sc_cbrt_common(sc_div_const_common(Z, w.z, w.rz, other), other);
    hq_float3 lab;
    lab.x = HQ_FSUB(HQ_FMUL(116.0f, fy), 16.0f);
    lab.y = HQ_FMUL(500.0f, HQ_FSUB(fx, fy));
    lab.z = HQ_FMUL(200.0f, HQ_FSUB(fy, fz));
    if (other) lab = sc_opp_to_lab_general(o0, o1, o2, w);
    return lab;
#else
    return hq_cl_opp_to_lab_white(o0, o1, o2, w);
#endif
}

// Never called (seg_rows is positive).  Without a call site inside the V loop ptxas hoists the loop-invariant filter coefficients
// out of it into ~70 ordinary registers and then spills them (372 bytes); with one it keeps re-loading them into uniform
// registers inside the loop, as it did while the Opp2LAB epilogue (which has a rare out-of-line call) lived there.
__device__ __noinline__ void sc_never_called() { asm volatile("trap;"); }

template <typename IdxT>
__global__ void __launch_bounds__(kSThreads, kSCtasPerSm)
sc_candidate_strip21_kernel(const IdxT* __restrict__ idx, const float4* __restrict__ tab, int K, int w, int h, size_t stride, int seg_rows,
                            const __grid_constant__ Filt21 f, hq_white white, ScRows rows, const float* __restrict__ lab_orig,
                            unsigned long long* __restrict__ err_out) {
    extern __shared__ __align__(16) unsigned char strip_smem[];
    float* s_ring = reinterpret_cast<float*>(strip_smem);                                              // [2][kSBand][7][kSW]
    float* s_out = reinterpret_cast<float*>(strip_smem + (size_t)2 * kSBand * 7 * kSW * sizeof(float));    // [2][kEpiHRows][3][kSW]
    float4* s_lut = reinterpret_cast<float4*>(strip_smem + ((size_t)2 * kSBand * 7 * kSW + kSOutFloats) * sizeof(float));  // [K]
    (void)s_out;
    __shared__ long long s_err[kSW / 32], s_err_v[kSW / 32];   // partial sums of the H warps / of the V warps
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kSW;
    const int ya = rows.y_begin + blockIdx.y * seg_rows;                 // output rows [ya, yb) of this segment (local rows)
    const int yb = min(ya + seg_rows, rows.y_begin + rows.y_count);
    // input (horizontally filtered) rows v = ya-10 .. yb+9 in bands of 8; row v completes output row v-10
    const int nbands = (yb - ya + 2 * kHalf + kSBand - 1) / kSBand;
    enum { kBarFull = 1, kBarEmpty = 3, kBarOutFull = 7, kBarOutEmpty = 9 };   // named barriers 1,2 (full[buf]), 3,4 (empty[buf]), 7,8 / 9,10 (the same for s_out); 5, 6: one half only; 0 is __syncthreads

    if (tid >= kSW) {
        // =============================== H warps (cl:234-272): one task (row, 8 outputs) per thread and band
        const int q = tid - kSW;
        const IdxT* idx_b = idx + (size_t)b * stride;
        for (int k = q; k < K; k += kSW) s_lut[k] = __ldg(tab + (size_t)b * K + k);
        named_bar_sync_i<5>(kSW);                      // the table is staged (H warps only)
        const int hrow = q / (kSW / kSHOut), hg = q - hrow * (kSW / kSHOut);
        const int xs = x0 + hg * kSHOut - kHalf;
        const bool interior = xs >= 0 && xs + kSHOut + kT - 1 <= w;
#if HQ_SC_EPI_H
        // Opp2LAB (cl:124-145) + CIE76 (cl:209) + fixed-point sum for the rows the V warps hand over, one band later (they run one band
        // behind this half).  With the whole epilogue on their side the V warps carried 342 of the stage's 591 instructions per pixel
        // against 200 here: the H warps waited at the barriers 40 % of the time and a scheduler had two busy warps.  Task = kEpiHRows
        // pixels of one handed-over row.
        long long hfx = 0;
        constexpr int kEpiPerRow = kSW / kEpiHRows;                   // threads per handed-over row
        const int er = q / kEpiPerRow, ec0 = (q - er * kEpiPerRow) * kEpiHRows;
        const int erow = (er / (4 - kEpiV)) * 4 + kEpiV + er % (4 - kEpiV);   // its row inside the band
        auto h_epilogue = [&](int eb) {
            const int ebuf = eb & 1;
            const int y = ya - 2 * kHalf + eb * kSBand + erow;        // band eb's input row v = ya-10+8*eb+erow completed output v-10
            const int yc = y < ya ? ya : (y >= yb ? yb - 1 : y);
            const bool rowok = y >= ya && y < yb;
            float lo[3][kEpiHRows];
#pragma unroll
            for (int i = 0; i < kEpiHRows; ++i) {
                int x = x0 + ec0 + i;
                x = x < w ? x : w - 1;
                const size_t pp = (size_t)yc * w + x;
                lo[0][i] = __ldg(lab_orig + pp); lo[1][i] = __ldg(lab_orig + stride + pp); lo[2][i] = __ldg(lab_orig + 2 * stride + pp);
            }
            named_bar_sync<kBarOutFull>(ebuf, kSThreads);
            const float* src = s_out + ((size_t)ebuf * kEpiHRows + er) * 3 * kSW + ec0;
#pragma unroll 1
            for (int i4 = 0; i4 < kEpiHRows; i4 += 4) {   // four pixels at a time (registers, code size)
                const float4 o0 = *reinterpret_cast<const float4*>(src + i4);
                const float4 o1 = *reinterpret_cast<const float4*>(src + kSW + i4);
                const float4 o2 = *reinterpret_cast<const float4*>(src + 2 * kSW + i4);
                const float a0[4] = {o0.x, o0.y, o0.z, o0.w}, a1[4] = {o1.x, o1.y, o1.z, o1.w}, a2[4] = {o2.x, o2.y, o2.z, o2.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const hq_float3 lab = sc_opp_to_lab_straight(a0[i], a1[i], a2[i], white);
                    const float l0 = (kEpiHRows > 4 && i4) ? lo[0][(4 + i) % kEpiHRows] : lo[0][i], l1 = (kEpiHRows > 4 && i4) ? lo[1][(4 + i) % kEpiHRows] : lo[1][i],
                                l2 = (kEpiHRows > 4 && i4) ? lo[2][(4 + i) % kEpiHRows] : lo[2][i];
                    const long long e = hq_to_fx(HQ_FSQRT(hq_dist2(l0, l1, l2, lab.x, lab.y, lab.z)));
                    hfx += (rowok && x0 + ec0 + i4 + i < w) ? e : 0ll;
                }
            }
            if (eb + 2 < nbands) named_bar_arrive<kBarOutEmpty>(ebuf, kSThreads);   // the buffer is read: a V warp may refill it
        };
#endif
        for (int band = 0; band < nbands; ++band) {
            const int buf = band & 1;
            // virtual row -> array row: reflection at the GLOBAL borders; rows only discarded outputs need are clamped
            int lr = hq_reflect(rows.g0 + ya - kHalf + band * kSBand + hrow, rows.gh) - rows.g0;
            lr = lr < 0 ? 0 : (lr >= h ? h - 1 : lr);
            const IdxT* irow = idx_b + (size_t)lr * w;
            float win[3][kSHOut + kT - 1];
            if (interior) {
                IdxT id[kSHOut + kT - 1];
#pragma unroll
                for (int i = 0; i < kSHOut + kT - 1; ++i) id[i] = __ldg(irow + xs + i);
#pragma unroll
                for (int i = 0; i < kSHOut + kT - 1; ++i) { const float4 c = s_lut[id[i]]; win[0][i] = c.x; win[1][i] = c.y; win[2][i] = c.z; }
            } else {
#pragma unroll
                for (int i = 0; i < kSHOut + kT - 1; ++i) {
                    int xx = xs + i;
                    xx = xx < w + kHalf ? hq_reflect(xx, w) : 0;   // beyond the last output's right halo: never used
                    xx = xx < 0 ? 0 : (xx >= w ? w - 1 : xx);
                    const float4 c = s_lut[__ldg(irow + xx)];
                    win[0][i] = c.x; win[1][i] = c.y; win[2][i] = c.z;
                }
            }
            if (band >= 2) named_bar_sync<kBarEmpty>(buf, kSThreads);   // the V warps have read this buffer's previous band
            float* dst = s_ring + ((size_t)buf * kSBand + hrow) * 7 * kSW + hg * kSHOut;
#pragma unroll 1
            for (int half = 0; half < HQ_SC_HHALVES; ++half) {   // kSHPer outputs from the head of the window, then the window moves down
                float acc[kSHPer][7];
#pragma unroll
                for (int o = 0; o < kSHPer; ++o)
#pragma unroll
                    for (int c = 0; c < 7; ++c) acc[o][c] = 0.f;
#pragma unroll
                for (int t = 0; t < kT; ++t)
#pragma unroll
                    for (int o = 0; o < kSHPer; ++o) {
                        const float i0 = win[0][o + t], i1 = win[1][o + t], i2 = win[2][o + t];
                        acc[o][0] = HQ_FFMA(i0, f.v[3 * t], acc[o][0]); acc[o][1] = HQ_FFMA(i1, f.v[3 * t + 1], acc[o][1]); acc[o][2] = HQ_FFMA(i2, f.v[3 * t + 2], acc[o][2]);
                        acc[o][3] = HQ_FFMA(i0, f.v[3 * kT + 3 * t], acc[o][3]); acc[o][4] = HQ_FFMA(i1, f.v[3 * kT + 3 * t + 1], acc[o][4]); acc[o][5] = HQ_FFMA(i2, f.v[3 * kT + 3 * t + 2], acc[o][5]);
                        acc[o][6] = HQ_FFMA(i0, f.v[6 * kT + t], acc[o][6]);
                    }
#pragma unroll
                for (int c = 0; c < 7; ++c)
#pragma unroll
                    for (int o4 = 0; o4 < kSHPer; o4 += 4)
                        *reinterpret_cast<float4*>(dst + c * kSW + kSHPer * half + o4) = make_float4(acc[o4][c], acc[o4 + 1][c], acc[o4 + 2][c], acc[o4 + 3][c]);
                if (HQ_SC_HHALVES > 1) {
#pragma unroll
                    for (int i = 0; i < kSHOut + kT - 1 - kSHPer; ++i) { win[0][i] = win[0][i + kSHPer]; win[1][i] = win[1][i + kSHPer]; win[2][i] = win[2][i + kSHPer]; }
                }
            }
            named_bar_arrive<kBarFull>(buf, kSThreads);
#if HQ_SC_EPI_H
            if (band >= 1) h_epilogue(band - 1);
#endif
        }
#if HQ_SC_EPI_H
        h_epilogue(nbands - 1);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) hfx += __shfl_down_sync(0xffffffffu, hfx, off);
        if ((q & 31) == 0) s_err[q >> 5] = hfx;
        named_bar_sync_i<5>(kSW);                      // H warps only
        if (q == 0) {
            long long e = 0;
#pragma unroll
            for (int i = 0; i < kSW / 32; ++i) e += s_err[i];
            if (e) atomicAdd(err_out + b, (unsigned long long)e);
        }
#endif
        return;
    }

    // =================================== V warps (cl:274-306), transposed: one column per thread
    const int x = x0 + tid;
    const bool vact = x < w;
    float A[kT + kSVUnroll - 1][3];                   // pending outputs: at phase p of a row group, tap t lives in slot t + kSVUnroll - 1 - p
#pragma unroll
    for (int j = 0; j < kT + kSVUnroll - 1; ++j) A[j][0] = A[j][1] = A[j][2] = 0.f;
    long long fx = 0;
    for (int band = 0; band < nbands; ++band) {
        const int buf = band & 1;
        const int vb = ya - kHalf + band * kSBand;    // first input row of the band
        named_bar_sync<kBarFull>(buf, kSThreads);
#if HQ_SC_EPI_H
        if (band >= 2) named_bar_sync<kBarOutEmpty>(buf, kSThreads);   // the H warps hold this buffer's previous outputs in registers
#endif
        if (vact) {
#pragma unroll 1
            for (int r0 = 0; r0 < kSBand; r0 += kSVUnroll) {
                // the outputs this group completes and the V warps score themselves: rows vb + r0 + p - 10, p < kEpiV; their S-CIELAB
                // originals fly during the fma work
                float lo[kEpiV > 0 ? kEpiV : 1][3];
#pragma unroll
                for (int p = 0; p < kEpiV; ++p) {
                    int y = vb + r0 + p - kHalf;
                    y = y < ya ? ya : (y >= yb ? yb - 1 : y);
                    const size_t pp = (size_t)y * w + x;
                    lo[p][0] = __ldg(lab_orig + pp); lo[p][1] = __ldg(lab_orig + stride + pp); lo[p][2] = __ldg(lab_orig + 2 * stride + pp);
                }
#pragma unroll
                for (int p = 0; p < kSVUnroll; ++p) {
                    const float* src = s_ring + ((size_t)buf * kSBand + r0 + p) * 7 * kSW + tid;
                    const float t10 = src[0], t11 = src[kSW], t12 = src[2 * kSW], t20 = src[3 * kSW], t21 = src[4 * kSW], t22 = src[5 * kSW], t3 = src[6 * kSW];
#pragma unroll
                    for (int t = 0; t < kT; ++t) {
                        float* a = A[t + kSVUnroll - 1 - p];
                        a[0] = HQ_FFMA(t10, f.v[3 * t], HQ_FFMA(t20, f.v[3 * kT + 3 * t], a[0]));
                        a[1] = HQ_FFMA(t11, f.v[3 * t + 1], HQ_FFMA(t21, f.v[3 * kT + 3 * t + 1], a[1]));
                        a[2] = HQ_FFMA(t12, f.v[3 * t + 2], HQ_FFMA(t22, f.v[3 * kT + 3 * t + 2], a[2]));
                        a[0] = HQ_FFMA(t3, f.v[7 * kT + t], a[0]);
                    }
                    const float* a = A[kT + kSVUnroll - 2 - p];            // completed: tap 20 was the last
                    if (p >= kEpiV) {   // to the H warps, which score it one band later
                        if (HQ_SC_EPI_H == 1 && seg_rows < 0) sc_never_called();
                        float* dst = s_out + ((size_t)buf * kEpiHRows + (r0 / 4) * (4 - kEpiV) + (p - kEpiV)) * 3 * kSW + tid;
                        dst[0] = a[0]; dst[kSW] = a[1]; dst[2 * kSW] = a[2];
                    } else {
                        const int y = vb + r0 + p - kHalf;
#if HQ_SC_STRAIGHT
                        // rows outside the segment are computed like the others and dropped by a select (no divergence, no branch)
                        const hq_float3 lab = sc_opp_to_lab_straight(a[0], a[1], a[2], white);
                        const long long e = hq_to_fx(HQ_FSQRT(hq_dist2(lo[p < kEpiV ? p : 0][0], lo[p < kEpiV ? p : 0][1], lo[p < kEpiV ? p : 0][2], lab.x, lab.y, lab.z)));
                        fx += (y >= ya && y < yb) ? e : 0ll;
#else
                        if (y >= ya && y < yb) {
                            const hq_float3 lab = hq_cl_opp_to_lab_white(a[0], a[1], a[2], white);
                            fx += hq_to_fx(HQ_FSQRT(hq_dist2(lo[p < kEpiV ? p : 0][0], lo[p < kEpiV ? p : 0][1], lo[p < kEpiV ? p : 0][2], lab.x, lab.y, lab.z)));
                        }
#endif
                    }
                }
                // kSVUnroll rows further every pending output's tap has grown by as much: slot j -> j + kSVUnroll, fresh outputs enter below
#pragma unroll
                for (int j = kT + kSVUnroll - 2; j >= kSVUnroll; --j) { A[j][0] = A[j - kSVUnroll][0]; A[j][1] = A[j - kSVUnroll][1]; A[j][2] = A[j - kSVUnroll][2]; }
#pragma unroll
                for (int j = 0; j < kSVUnroll; ++j) A[j][0] = A[j][1] = A[j][2] = 0.f;
            }
        }
#if HQ_SC_EPI_H
        named_bar_arrive<kBarOutFull>(buf, kSThreads);
#endif
        if (band + 2 < nbands) named_bar_arrive<kBarEmpty>(buf, kSThreads);   // an H warp will wait for it
    }
    if (kEpiV > 0) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) fx += __shfl_down_sync(0xffffffffu, fx, off);
        if ((tid & 31) == 0) s_err_v[tid >> 5] = fx;
        named_bar_sync_i<6>(kSW);                          // V warps only
        if (tid == 0) {
            long long e = 0;
#pragma unroll
            for (int i = 0; i < kSW / 32; ++i) e += s_err_v[i];
            if (e) atomicAdd(err_out + b, (unsigned long long)e);
        }
    }
}

// error-image mode (ImageManipulation.computeError :858-894): dE between two S-CIELAB images, the
// map value ((255 - dE)^2) / (255*255) (:890) and the fixed-point sum of dE
__global__ void sc_error_image_kernel(const float* __restrict__ lab_a, const float* __restrict__ lab_b, size_t n, size_t stride,
                                      float* __restrict__ map_out, uint8_t* __restrict__ map_u8, unsigned long long* __restrict__ err_out,
                                      int de_type) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long fx = 0;
    if (i < n) {
        const float e = de_type == 0 ? HQ_FSQRT(hq_dist2(lab_a[i], lab_a[stride + i], lab_a[2 * stride + i], lab_b[i], lab_b[stride + i], lab_b[2 * stride + i]))
                                     : hq_cl_delta_e94(lab_a[i], lab_a[stride + i], lab_a[2 * stride + i], lab_b[i], lab_b[stride + i], lab_b[2 * stride + i]);
        const float d = HQ_FSUB(255.0f, e);
        const float v = HQ_FDIV(HQ_FMUL(d, d), 65025.0f);
        if (map_out) map_out[i] = v;
        if (map_u8) {
            const float q = HQ_FADD(HQ_FMUL(v, 255.0f), 0.5f);
            map_u8[i] = (uint8_t)__float2int_rz(fminf(fmaxf(q, 0.0f), 255.0f));   // (a NaN of the CIE94 branch renders as 0)
        }
        if (e != e) atomicAdd(err_out + 1, 1ull);   // word 1: NaN count (CIE94 only)
        else fx = hq_to_fx(e);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) fx += __shfl_down_sync(0xffffffffu, fx, off);
    __shared__ long long s_err[8];
    if ((threadIdx.x & 31) == 0) s_err[threadIdx.x >> 5] = fx;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long e = 0;
        for (int k = 0; k < 8; ++k) e += s_err[k];
        if (e) atomicAdd(err_out, (unsigned long long)e);
    }
}

// ---- the reference class's one-shot entries on its own interleaved layouts (float4 per pixel, 4th lane 0)
// RGBtoXYZ (ImageManipulation.java:100-152, kernel RGB2XYZ cl:79-90): planar sRGB floats -> XYZ float4
__global__ void sc_unit_to_xyz4_kernel(const float* __restrict__ r, const float* __restrict__ g, const float* __restrict__ b, size_t n,
                                       float4* __restrict__ xyz, unsigned int* __restrict__ bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float cr = r[i], cg = g[i], cb = b[i];
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cr >= 0.f && cr <= 1.f && cg >= 0.f && cg <= 1.f && cb >= 0.f && cb <= 1.f) {
        const float R = hq_srgb_decode(cr), G = hq_srgb_decode(cg), B = hq_srgb_decode(cb);
        o.x = hq_cl_dot3(0.4124564f, 0.3575761f, 0.1804375f, R, G, B);
        o.y = hq_cl_dot3(0.2126729f, 0.7151522f, 0.0721750f, R, G, B);
        o.z = hq_cl_dot3(0.0193339f, 0.1191920f, 0.9503041f, R, G, B);
    } else {
        atomicOr(bad, 1u);
    }
    xyz[i] = o;
}
// XYZ2Opp (cl:111-116) into the planes the filter kernels read
__global__ void sc_xyz4_to_opp_kernel(const float4* __restrict__ xyz, size_t n, size_t stride, float* __restrict__ opp) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = xyz[i];
    opp[i] = hq_cl_dot3(0.2787336f, 0.7218031f, -0.1065520f, v.x, v.y, v.z);
    opp[stride + i] = hq_cl_dot3(-0.4487736f, 0.2898056f, -0.0771569f, v.x, v.y, v.z);
    opp[2 * stride + i] = hq_cl_dot3(0.0859513f, -0.5899859f, 0.5011089f, v.x, v.y, v.z);
}
__global__ void sc_planes_to_f4_kernel(const float* __restrict__ planes, size_t n, size_t stride, float4* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(planes[i], planes[stride + i], planes[2 * stride + i], 0.f);
}
__global__ void sc_f4_to_planes_kernel(const float4* __restrict__ in, size_t n, size_t stride, float* __restrict__ planes) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = in[i];
    planes[i] = v.x; planes[stride + i] = v.y; planes[2 * stride + i] = v.z;
}
// computeError (ImageManipulation.java:858-894): CIEDE kernel (cl:201-209) on two interleaved Lab images + the error image
// value ((255 - e) * (255 - e)) / (255 * 255) of :890, written to the first three lanes as the reference does
__global__ void sc_delta_e4_kernel(const float4* __restrict__ a, const float4* __restrict__ b, size_t n, float* __restrict__ e_out,
                                   float4* __restrict__ err_img, int de_type) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = a[i], q = b[i];
    const float e = de_type == 0 ? HQ_FSQRT(hq_dist2(p.x, p.y, p.z, q.x, q.y, q.z)) : hq_cl_delta_e94(p.x, p.y, p.z, q.x, q.y, q.z);
    e_out[i] = e;
    if (err_img) {
        const float d = HQ_FSUB(255.0f, e);
        const float v = HQ_FDIV(HQ_FMUL(d, d), 65025.0f);
        float4 o = err_img[i];   // the 4th lane is left as the caller has it (:890 writes off .. off+2)
        o.x = v; o.y = v; o.z = v;
        err_img[i] = o;
    }
}

// identity-filter cost under a dE other than CIE76: the assignment kernels score with the squared distance they minimise, so
// for CIE94 the population is assigned first (index images) and scored here: dE94(Lab(pixel), Lab(P[idx])) over the own pixels
template <typename IdxT>
__global__ void sc_score_indices_kernel(const IdxT* __restrict__ idx, const float* __restrict__ lab, size_t stride, size_t own_lo, size_t own_hi,
                                        const float4* __restrict__ pal_lab, int K8, int de_type, unsigned long long* __restrict__ err_out,
                                        unsigned long long* __restrict__ nan_out) {
    const int b = blockIdx.y;
    long long fx = 0;
    for (size_t i = own_lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < own_hi; i += (size_t)gridDim.x * blockDim.x) {
        const float4 c = __ldg(pal_lab + (size_t)b * K8 + idx[(size_t)b * stride + i]);
        fx += sc_delta_e_fx(de_type, lab[i], lab[stride + i], lab[2 * stride + i], c.x, c.y, c.z, nan_out + b);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) fx += __shfl_down_sync(0xffffffffu, fx, off);
    if ((threadIdx.x & 31) == 0 && fx) atomicAdd(err_out + b, (unsigned long long)fx);
}

}  // namespace

cudaError_t launch_sc_score_indices(const void* d_idx, bool idx16, const float* d_lab, size_t stride, size_t own_lo, size_t own_hi,
                                    const float4* d_pal_lab, int K8, int B, int de_type, unsigned long long* d_err, unsigned long long* d_nan,
                                    int sm_count, cudaStream_t st) {
    if (own_hi <= own_lo || B == 0) return cudaSuccess;
    const dim3 grid((unsigned)((sm_count > 0 ? sm_count : 148) * 4), (unsigned)B);
    if (idx16) sc_score_indices_kernel<uint16_t><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(d_idx), d_lab, stride, own_lo, own_hi, d_pal_lab, K8, de_type, d_err, d_nan);
    else sc_score_indices_kernel<uint8_t><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(d_idx), d_lab, stride, own_lo, own_hi, d_pal_lab, K8, de_type, d_err, d_nan);
    return cudaGetLastError();
}

cudaError_t launch_sc_unit_to_xyz4(const float* d_r, const float* d_g, const float* d_b, size_t n, float* d_xyz4, unsigned int* d_bad, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    sc_unit_to_xyz4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_r, d_g, d_b, n, reinterpret_cast<float4*>(d_xyz4), d_bad);
    return cudaGetLastError();
}
cudaError_t launch_sc_xyz4_to_opp(const float* d_xyz4, size_t n, size_t stride, float* d_opp, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    sc_xyz4_to_opp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(d_xyz4), n, stride, d_opp);
    return cudaGetLastError();
}
cudaError_t launch_sc_planes_to_f4(const float* d_planes, size_t n, size_t stride, float* d_out4, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    sc_planes_to_f4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_planes, n, stride, reinterpret_cast<float4*>(d_out4));
    return cudaGetLastError();
}
cudaError_t launch_sc_f4_to_planes(const float* d_in4, size_t n, size_t stride, float* d_planes, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    sc_f4_to_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(d_in4), n, stride, d_planes);
    return cudaGetLastError();
}
cudaError_t launch_sc_delta_e4(const float* d_a4, const float* d_b4, size_t n, float* d_e, float* d_err_img4, int de_type, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    sc_delta_e4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(d_a4), reinterpret_cast<const float4*>(d_b4), n, d_e,
                                                                     reinterpret_cast<float4*>(d_err_img4), de_type);
    return cudaGetLastError();
}

cudaError_t launch_sc_error_image(const float* d_lab_a, const float* d_lab_b, size_t n, size_t stride, float* d_map, uint8_t* d_map_u8,
                                  unsigned long long* d_err, int de_type, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    sc_error_image_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_lab_a, d_lab_b, n, stride, d_map, d_map_u8, d_err, de_type);
    return cudaGetLastError();
}

cudaError_t launch_sc_rgb_to_opp(const uint8_t* d_rgb, size_t n, size_t stride, const float* d_table, float* d_opp, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    sc_rgb_to_opp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_rgb, n, stride, d_table, d_opp);
    return cudaGetLastError();
}

cudaError_t launch_sc_palette_opp(const float* d_palettes, int total, float4* d_tab, cudaStream_t st, unsigned long long* d_zero, int zero_words) {
    const int nthreads = total > zero_words ? total : zero_words;
    if (nthreads == 0) return cudaSuccess;
    sc_palette_opp_kernel<<<(nthreads + 127) / 128, 128, 0, st>>>(d_palettes, total, d_tab, d_zero, d_zero ? zero_words : 0);
    return cudaGetLastError();
}

// Result words of a reference-faithful evaluation -> the caller's pinned host buffer, word 0 of every candidate taken from the
// S-CIELAB error sums, then the sequence number the host spins on: one launch instead of a strided D2D copy, a D2H copy and a
// stream wait (small result sets on one device: the plugin's default search is a chain of dependent launches).
__global__ void __launch_bounds__(256) sc_export_kernel(const unsigned long long* __restrict__ results, const unsigned long long* __restrict__ sc_err, int words,
                                                        unsigned nwords, unsigned long long* dst, volatile unsigned long long* flag, unsigned long long seq) {
    for (unsigned i = threadIdx.x; i < nwords; i += blockDim.x) {
        const unsigned b = i / (unsigned)words;
        dst[i] = (i - b * (unsigned)words == 0u) ? sc_err[b] : results[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *flag = seq;
}
cudaError_t launch_sc_export(const unsigned long long* d_results, const unsigned long long* d_sc_err, int words, size_t nwords, unsigned long long* h_dst_mapped,
                             unsigned long long* h_flag_mapped, unsigned long long seq, cudaStream_t st) {
    sc_export_kernel<<<1, 256, 0, st>>>(d_results, d_sc_err, words, (unsigned)nwords, h_dst_mapped, h_flag_mapped, seq);
    return cudaGetLastError();
}

cudaError_t launch_sc_unit_to_opp(const float* d_unit, size_t n, size_t stride, float* d_opp, unsigned int* d_bad, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    sc_unit_to_opp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_unit, n, stride, d_opp, d_bad);
    return cudaGetLastError();
}

cudaError_t launch_sc_original(const float* d_opp, int w, int h, size_t stride, const float* d_filters, const float* h_filters, int taps,
                               int whitepoint, ScRows rows, float* d_tmp, float* d_lab_out, cudaStream_t st, const float* illuminant3) {
    if (w == 0 || h == 0 || rows.y_count == 0) return cudaSuccess;
    hq_float3 ill = hq_whitepoint(whitepoint);
    if (illuminant3) { ill.x = illuminant3[0]; ill.y = illuminant3[1]; ill.z = illuminant3[2]; }   // XYZtoScielab's float[] illuminant (:285)
    if (taps == kT && h_filters) {
        Filt21 f21;
        for (int i = 0; i < 8 * kT; ++i) f21.v[i] = h_filters[i];
        const dim3 gh((unsigned)((w + kHSeg - 1) / kHSeg), (unsigned)h), gv((unsigned)((w + kScThreads - 1) / kScThreads), (unsigned)((rows.y_count + kVRows - 1) / kVRows));
        sc_hpass21_kernel<0, uint8_t><<<gh, kScThreads, 0, st>>>(d_opp, nullptr, nullptr, w, h, stride, f21, d_tmp);
        sc_vpass21_kernel<0><<<gv, kScThreads, 0, st>>>(d_tmp, w, h, stride, f21, ill, rows, d_lab_out, nullptr, nullptr, 0, nullptr);
        return cudaGetLastError();
    }
    const ScFilters f{d_filters, taps};
    const dim3 grid((unsigned)((w + kScThreads - 1) / kScThreads), (unsigned)h), gridv((unsigned)((w + kScThreads - 1) / kScThreads), (unsigned)rows.y_count);
    sc_hpass_kernel<0, uint8_t><<<grid, kScThreads, 7 * taps * sizeof(float), st>>>(d_opp, nullptr, nullptr, w, h, stride, f, d_tmp);
    sc_vpass_kernel<0><<<gridv, kScThreads, 8 * taps * sizeof(float), st>>>(d_tmp, w, h, stride, f, ill, rows, d_lab_out, nullptr, nullptr, 0, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_sc_candidate(const void* d_idx, bool idx16, const float4* d_tab, int w, int h, size_t stride, const float* d_filters,
                                const float* h_filters, int taps, int whitepoint, ScRows rows, float* d_tmp, const float* d_lab_orig,
                                unsigned long long* d_err, cudaStream_t st, int de_type, unsigned long long* d_nan) {
    if (w == 0 || h == 0 || rows.y_count == 0) return cudaSuccess;
    if (taps == kT && h_filters) {
        Filt21 f21;
        for (int i = 0; i < 8 * kT; ++i) f21.v[i] = h_filters[i];
        const dim3 gh((unsigned)((w + kHSeg - 1) / kHSeg), (unsigned)h), gv((unsigned)((w + kScThreads - 1) / kScThreads), (unsigned)((rows.y_count + kVRows - 1) / kVRows));
        if (idx16) sc_hpass21_kernel<1, uint16_t><<<gh, kScThreads, 0, st>>>(nullptr, static_cast<const uint16_t*>(d_idx), d_tab, w, h, stride, f21, d_tmp);
        else sc_hpass21_kernel<1, uint8_t><<<gh, kScThreads, 0, st>>>(nullptr, static_cast<const uint8_t*>(d_idx), d_tab, w, h, stride, f21, d_tmp);
        sc_vpass21_kernel<1><<<gv, kScThreads, 0, st>>>(d_tmp, w, h, stride, f21, hq_whitepoint(whitepoint), rows, nullptr, d_lab_orig, d_err, de_type, d_nan);
        return cudaGetLastError();
    }
    const ScFilters f{d_filters, taps};
    const dim3 grid((unsigned)((w + kScThreads - 1) / kScThreads), (unsigned)h), gridv((unsigned)((w + kScThreads - 1) / kScThreads), (unsigned)rows.y_count);
    if (idx16) sc_hpass_kernel<1, uint16_t><<<grid, kScThreads, 7 * taps * sizeof(float), st>>>(nullptr, static_cast<const uint16_t*>(d_idx), d_tab, w, h, stride, f, d_tmp);
    else sc_hpass_kernel<1, uint8_t><<<grid, kScThreads, 7 * taps * sizeof(float), st>>>(nullptr, static_cast<const uint8_t*>(d_idx), d_tab, w, h, stride, f, d_tmp);
    sc_vpass_kernel<1><<<gridv, kScThreads, 8 * taps * sizeof(float), st>>>(d_tmp, w, h, stride, f, hq_whitepoint(whitepoint), rows, nullptr, d_lab_orig, d_err, de_type, d_nan);
    return cudaGetLastError();
}

// all B candidates of a population in ONE launch (taps = 21, K <= kMaxColors); d_idx: [B][stride] indices, d_tab: [B][K] opponent
// colours, d_err: [B] fixed-point sums (added to).  Returns cudaErrorNotSupported when the fused kernel does not apply
// (then the caller runs launch_sc_candidate per candidate).
cudaError_t launch_sc_candidates_fused(const void* d_idx, bool idx16, const float4* d_tab, int K, int B, int w, int h, size_t stride,
                                       const float* h_filters, int taps, int whitepoint, ScRows rows, const float* d_lab_orig,
                                       unsigned long long* d_err, int sm_count, cudaStream_t st) {
    if (taps != kT || !h_filters || K > kMaxColors || B > 65535) return cudaErrorNotSupported;
    if (w == 0 || h == 0 || rows.y_count == 0 || B == 0) return cudaSuccess;
    Filt21 f21;
    for (int i = 0; i < 8 * kT; ++i) f21.v[i] = h_filters[i];
    const size_t smem = ((size_t)2 * kSBand * 7 * kSW + kSOutFloats) * sizeof(float) + (size_t)K * sizeof(float4);
    const int strips = (w + kSW - 1) / kSW;
    // Rows per segment.  Every segment start costs 20 extra horizontally filtered rows, a CTA walks its segment band by band (8 rows),
    // and `slots` CTAs run at a time: the segment length that minimises bands-per-CTA x rounds.  Round 2 first used "as long as the
    // grid still holds >= 2.5 waves", which is what this picks for a 4K population too (7 segments, 840 CTAs), but it left a 512 x 512
    // population at 128 CTAs of 11 bands on 296 slots (58 us; the plugin's default search is launch- and tail-bound there).
    const long long slots = (long long)(sm_count > 0 ? sm_count : 148) * kSCtasPerSm;
    int seg_rows = kSBand;
    {
        double best_cost = 1e300;
        for (int cand = 2 * kSBand; cand <= 1024 + kSBand; cand += kSBand) {
            const int nseg = (rows.y_count + cand - 1) / cand;
            const int seg = ((rows.y_count + nseg - 1) / nseg + kSBand - 1) / kSBand * kSBand;   // equal segments: the last one must not be a stub
            const int nseg_eq = (rows.y_count + seg - 1) / seg;
            const double waves = (double)strips * B * nseg_eq / (double)slots;
            const double rounds = waves < 3.0 ? std::ceil(waves) : waves + 0.5;                  // few waves run in lockstep, many average out
            const double cost = rounds * (double)((seg + 2 * kHalf + kSBand - 1) / kSBand);
            if (cost <= best_cost) { best_cost = cost; seg_rows = seg; }                           // ties: the longer segment
            if (nseg == 1) break;
        }
    }
    const dim3 grid((unsigned)strips, (unsigned)((rows.y_count + seg_rows - 1) / seg_rows), (unsigned)B);
    if (grid.y > 65535) return cudaErrorNotSupported;
    static std::mutex mu;
    static size_t configured[2][64];   // per index width and device: largest dynamic shared memory set so far (process-wide, monotonic)
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lock(mu);
        size_t& have = configured[idx16 ? 1 : 0][dev & 63];
        if (smem > have) {
            e = idx16 ? cudaFuncSetAttribute(sc_candidate_strip21_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : cudaFuncSetAttribute(sc_candidate_strip21_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            have = smem;
        }
    }
    const hq_white white = hq_make_white(whitepoint);
    if (idx16) sc_candidate_strip21_kernel<uint16_t><<<grid, kSThreads, smem, st>>>(static_cast<const uint16_t*>(d_idx), d_tab, K, w, h, stride, seg_rows, f21, white, rows, d_lab_orig, d_err);
    else sc_candidate_strip21_kernel<uint8_t><<<grid, kSThreads, smem, st>>>(static_cast<const uint8_t*>(d_idx), d_tab, K, w, h, stride, seg_rows, f21, white, rows, d_lab_orig, d_err);
    return cudaGetLastError();
}

}  // namespace hq
