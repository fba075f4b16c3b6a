// microbench3.cu — candidate inner loops for the expanded-form prefilter
//   s(k) = x0*a_k + x1*b_k + x2*c_k + e_k   (a=-2p0, b=-2p1, c=-2p2, e=|p|^2), 3 FMA per (pixel, colour)
// with chunked running-min bookkeeping (best, second-best chunk min, chunk index), palette in
// shared memory, R pixels per thread.  Reports cycles per (pixel, colour) pair per SMSP-lane.
//   V1 scalar FFMA, LDS.128 per colour
//   V2 FFMA2 packed over colour pairs, pixel feature broadcast (.F32)
//   V3 FFMA2 packed over pixel pairs, palette scalar broadcast (.F32)
//   V0 the current direct form (sub,mul,fma,fma packed over colour pairs) for reference
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
// integer 3-input min on the bit patterns (valid float order when every value is >= 0): VIMNMX3
__device__ __forceinline__ float imin3(float a, float b, float c) { return __int_as_float(min(min(__float_as_int(a), __float_as_int(b)), __float_as_int(c))); }

constexpr int K = 256, CH = 8;
static int g_only_v = -1, g_only_r = -1;

template <int V, int R>
__global__ void __launch_bounds__(256) sweep(const float* __restrict__ in, float* __restrict__ out, int iters, unsigned long long* cyc) {
    __shared__ __align__(16) float4 s_aos[K];       // (a,b,c,e) per colour
    __shared__ __align__(16) float4 s_p01[K / 2];   // (a_k,a_k+1,b_k,b_k+1)
    __shared__ __align__(16) float4 s_p23[K / 2];   // (c_k,c_k+1,e_k,e_k+1)
    for (int k = threadIdx.x; k < K; k += 256) {
        const float a = in[k & 63] * (k + 1), b = in[(k + 7) & 63] - k, c = in[(k + 13) & 63] + 0.5f * k, e = a * a + b * b + c * c;
        s_aos[k] = make_float4(a, b, c, e);
        float* p01 = reinterpret_cast<float*>(s_p01); float* p23 = reinterpret_cast<float*>(s_p23);
        p01[(k >> 1) * 4 + (k & 1)] = a; p01[(k >> 1) * 4 + 2 + (k & 1)] = b;
        p23[(k >> 1) * 4 + (k & 1)] = c; p23[(k >> 1) * 4 + 2 + (k & 1)] = e;
    }
    __syncthreads();
    float x0[R], x1[R], x2[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { x0[j] = in[(threadIdx.x + j) & 63]; x1[j] = in[(threadIdx.x + 3 * j + 1) & 63] * 2.f; x2[j] = in[(threadIdx.x + 5 * j + 2) & 63] - 1.f; }
    float acc = 0.f;
    const float INF = __int_as_float(0x7f800000);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float best[R], second[R]; int cidx[R];
#pragma unroll
        for (int j = 0; j < R; ++j) { best[j] = INF; second[j] = INF; cidx[j] = 0; }
        for (int c = 0; c < K / CH; ++c) {
            float m[R];
#pragma unroll
            for (int j = 0; j < R; ++j) m[j] = INF;
            if (V == 1) {
#pragma unroll
                for (int q = 0; q < CH; q += 2) {
                    const float4 p = s_aos[c * CH + q], r = s_aos[c * CH + q + 1];
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const float s0 = __fmaf_rn(x0[j], p.x, __fmaf_rn(x1[j], p.y, __fmaf_rn(x2[j], p.z, p.w)));
                        const float s1 = __fmaf_rn(x0[j], r.x, __fmaf_rn(x1[j], r.y, __fmaf_rn(x2[j], r.z, r.w)));
                        m[j] = min3(m[j], s0, s1);
                    }
                }
            } else if (V == 2) {
#pragma unroll
                for (int q = 0; q < CH / 2; ++q) {
                    const float4 u = s_p01[c * (CH / 2) + q], w = s_p23[c * (CH / 2) + q];
                    const uint64_t A = pk(u.x, u.y), B = pk(u.z, u.w), C = pk(w.x, w.y), E = pk(w.z, w.w);
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        float lo, hi;
                        upk(fma2(pk(x0[j], x0[j]), A, fma2(pk(x1[j], x1[j]), B, fma2(pk(x2[j], x2[j]), C, E))), lo, hi);
                        m[j] = min3(m[j], lo, hi);
                    }
                }
            } else if (V == 5) {
#pragma unroll
                for (int q = 0; q < CH / 2; ++q) {
                    const float4 u = s_p01[c * (CH / 2) + q], w = s_p23[c * (CH / 2) + q];
                    const uint64_t A = pk(u.x, u.y), B = pk(u.z, u.w), C = pk(w.x, w.y), E = pk(w.z, w.w);
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        float lo, hi;
                        upk(fma2(pk(x0[j], x0[j]), A, fma2(pk(x1[j], x1[j]), B, fma2(pk(x2[j], x2[j]), C, E))), lo, hi);
                        m[j] = imin3(m[j], lo, hi);
                    }
                }
            } else if (V == 4) {
#pragma unroll
                for (int q = 0; q < CH; q += 2) {
                    const float4 p = s_aos[c * CH + q], r = s_aos[c * CH + q + 1];
#pragma unroll
                    for (int j = 0; j < R; j += 2) {
                        const uint64_t X = pk(x0[j], x0[j + 1]), Y = pk(x1[j], x1[j + 1]), Z = pk(x2[j], x2[j + 1]);
                        float a0, a1, b0, b1;
                        upk(fma2(X, pk(p.x, p.x), fma2(Y, pk(p.y, p.y), fma2(Z, pk(p.z, p.z), pk(p.w, p.w)))), a0, a1);
                        upk(fma2(X, pk(r.x, r.x), fma2(Y, pk(r.y, r.y), fma2(Z, pk(r.z, r.z), pk(r.w, r.w)))), b0, b1);
                        m[j] = imin3(m[j], a0, b0);
                        m[j + 1] = imin3(m[j + 1], a1, b1);
                    }
                }
            } else if (V == 6) {  // scalar FFMA + integer min3
#pragma unroll
                for (int q = 0; q < CH; q += 2) {
                    const float4 p = s_aos[c * CH + q], r = s_aos[c * CH + q + 1];
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const float s0 = __fmaf_rn(x0[j], p.x, __fmaf_rn(x1[j], p.y, __fmaf_rn(x2[j], p.z, p.w)));
                        const float s1 = __fmaf_rn(x0[j], r.x, __fmaf_rn(x1[j], r.y, __fmaf_rn(x2[j], r.z, r.w)));
                        m[j] = imin3(m[j], s0, s1);
                    }
                }
            } else if (V == 3) {
#pragma unroll
                for (int q = 0; q < CH; q += 2) {
                    const float4 p = s_aos[c * CH + q], r = s_aos[c * CH + q + 1];
#pragma unroll
                    for (int j = 0; j < R; j += 2) {
                        const uint64_t X = pk(x0[j], x0[j + 1]), Y = pk(x1[j], x1[j + 1]), Z = pk(x2[j], x2[j + 1]);
                        float a0, a1, b0, b1;
                        upk(fma2(X, pk(p.x, p.x), fma2(Y, pk(p.y, p.y), fma2(Z, pk(p.z, p.z), pk(p.w, p.w)))), a0, a1);
                        upk(fma2(X, pk(r.x, r.x), fma2(Y, pk(r.y, r.y), fma2(Z, pk(r.z, r.z), pk(r.w, r.w)))), b0, b1);
                        m[j] = min3(m[j], a0, b0);
                        m[j + 1] = min3(m[j + 1], a1, b1);
                    }
                }
            } else {  // V0: direct form, colour pairs
#pragma unroll
                for (int q = 0; q < CH / 2; ++q) {
                    const float4 u = s_p01[c * (CH / 2) + q], w = s_p23[c * (CH / 2) + q];
                    const uint64_t A = pk(u.x, u.y), B = pk(u.z, u.w), C = pk(w.x, w.y);
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const uint64_t d0 = sub2(pk(x0[j], x0[j]), A), d1 = sub2(pk(x1[j], x1[j]), B), d2 = sub2(pk(x2[j], x2[j]), C);
                        float lo, hi;
                        upk(fma2(d2, d2, fma2(d1, d1, mul2(d0, d0))), lo, hi);
                        m[j] = min3(m[j], lo, hi);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < R; ++j) {
                second[j] = fminf(second[j], fmaxf(m[j], best[j]));
                if (m[j] < best[j]) { best[j] = m[j]; cidx[j] = c; }
            }
        }
#pragma unroll
        for (int j = 0; j < R; ++j) { acc += best[j] + second[j] + (float)cidx[j]; x0[j] += 1e-3f; }
    }
    const long long t1 = clock64();
    if (acc == 123.456f) out[threadIdx.x] = acc;
    if (threadIdx.x == 0) atomicMax(cyc, (unsigned long long)(t1 - t0));
}

template <int V, int R>
int run(const char* name, int sms, const float* din, float* dout, unsigned long long* dc) {
    if (g_only_v >= 0 && (g_only_v != V || g_only_r != R)) return 0;
    const int iters = g_only_v >= 0 ? 50 : 200;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep<V, R>, 256, 0));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, sweep<V, R>));
    for (int o = (g_only_v >= 0 ? (occ < 4 ? occ : 4) : 1); o <= occ && o <= 4; ++o) {
        unsigned long long best = ~0ull;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemset(dc, 0, 8));
            sweep<V, R><<<sms * o, 256>>>(din, dout, iters, dc);
            CK(cudaDeviceSynchronize());
            unsigned long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
            if (rep && c < best) best = c;
        }
        const double warps_per_smsp = 8.0 * o / 4.0;
        const double cyc_per_pair = (double)best / iters / (double)(K * R) / warps_per_smsp;
        printf("  {\"name\": \"%s\", \"R\": %d, \"regs\": %d, \"ctas_per_sm\": %d, \"cycles_per_pair\": %.3f, \"flop8_per_clk_lane\": %.3f, \"pct_of_2flop_peak\": %.1f},\n",
               name, R, fa.numRegs, o, cyc_per_pair, 8.0 / cyc_per_pair, 100.0 * 8.0 / cyc_per_pair / 2.0);
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 2) { g_only_v = atoi(argv[1]); g_only_r = atoi(argv[2]); }
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.5f + 0.013f * i;
    float *din, *dout; unsigned long long* dc;
    CK(cudaMalloc(&din, sizeof h)); CK(cudaMalloc(&dout, 4096)); CK(cudaMalloc(&dc, 8));
    CK(cudaMemcpy(din, h, sizeof h, cudaMemcpyHostToDevice));
    printf("{\"device\": \"%s\", \"results\": [\n", prop.name);
    run<0, 4>("V0 direct packed colour-pairs", sms, din, dout, dc);
    run<1, 4>("V1 expanded scalar", sms, din, dout, dc);
    run<1, 8>("V1 expanded scalar", sms, din, dout, dc);
    run<2, 4>("V2 expanded packed colour-pairs", sms, din, dout, dc);
    run<2, 8>("V2 expanded packed colour-pairs", sms, din, dout, dc);
    run<3, 4>("V3 expanded packed pixel-pairs", sms, din, dout, dc);
    run<3, 8>("V3 expanded packed pixel-pairs", sms, din, dout, dc);
    run<4, 4>("V4 = V3 + VIMNMX3", sms, din, dout, dc);
    run<4, 8>("V4 = V3 + VIMNMX3", sms, din, dout, dc);
    run<5, 4>("V5 = V2 + VIMNMX3", sms, din, dout, dc);
    run<6, 4>("V6 = V1 scalar + VIMNMX3", sms, din, dout, dc);
    run<6, 8>("V6 = V1 scalar + VIMNMX3", sms, din, dout, dc);
    printf("  {}]}\n");
    return 0;
}
