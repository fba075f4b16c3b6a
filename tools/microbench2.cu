// microbench2.cu — issue-model probes for sm_100a: does an ALU-pipe op (FMNMX / FMNMX3 / IADD3 /
// FSETP+SEL) co-issue in the shadow of FMA-pipe ops, scalar (FFMA) or packed (FFMA2)?
// Every op is asm volatile on its own dependent chain, 8 warps/CTA, 4 CTAs/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

#define FFMA(d, a, b) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(d) : "f"(a), "f"(b))
#define FFMA2(d, a, b) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d) : "l"(a), "l"(b))
// pair * broadcast-scalar + pair accumulator (the operand form of the pixel-pair packed sweep)
#define FFMA2P(d, x, sc) asm volatile("{ .reg .b64 t; mov.b64 t, {%2,%2}; fma.rn.f32x2 %0, %1, t, %0; }" : "+l"(d) : "l"(x), "f"(sc))
#define FMIN(d, a) asm volatile("min.f32 %0, %0, %1;" : "+f"(d) : "f"(a))
#define FMIN3(d, a, b) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(d) : "f"(a), "f"(b))
#define IADD(d, a) asm volatile("add.s32 %0, %0, %1;" : "+r"(d) : "r"(a))
#define IMNMX(d, a) asm volatile("min.s32 %0, %0, %1;" : "+r"(d) : "r"(a))
#define SETPSEL(d, a, b) asm volatile("{ .reg .pred p; setp.lt.f32 p, %1, %2; selp.f32 %0, %1, %0, p; }" : "+f"(d) : "f"(a), "f"(b))
#define LOP(d, a) asm volatile("xor.b32 %0, %0, %1;" : "+r"(d) : "r"(a))
// 3-input integer min (VIMNMX3): candidate replacement for FMNMX3 on non-negative floats reinterpreted as ints
#define VIMIN3(d, a, b) d = __vimin3_s32(d, a, b)

enum { M_FFMA, M_FFMA2, M_FMIN, M_FMIN3, M_IADD, M_IMNMX, M_SETPSEL, M_LOP,
       M_FFMA24_FMIN3_4, M_FFMA2_12_FMIN3_4, M_FFMA24_FMIN_8, M_FFMA24_IADD_8, M_FFMA2_12_IADD_8,
       M_FFMA24_SETPSEL_4, M_FFMA2_12_FMIN_8, M_FFMA24_LOP_8, M_FFMA16_FMIN3_8, M_FFMA2P_12, M_FFMA2P_12_FMIN3_4, M_FFMA2P_12_FMIN3_8, M_FFMA2P_12_FMIN_8, M_VIMIN3, M_FFMA2P_12_VIMIN3_4, M_FFMA2P_12_VIMIN3_8, M_COUNT };
static const char* names[M_COUNT] = {"ffma", "ffma2", "fmin", "fmin3", "iadd", "imnmx", "setpsel", "lop",
  "ffma24+fmin3x4", "ffma2x12+fmin3x4", "ffma24+fminx8", "ffma24+iaddx8", "ffma2x12+iaddx8",
  "ffma24+setpselx4", "ffma2x12+fminx8", "ffma24+lopx8", "ffma16+fmin3x8", "ffma2p x12", "ffma2p x12+fmin3x4", "ffma2p x12+fmin3x8", "ffma2p x12+fminx8", "vimin3", "ffma2p x12+vimin3x4", "ffma2p x12+vimin3x8"};
// warp-level instructions per loop iteration (setp+sel counted as 2)
static const int instrs[M_COUNT] = {24, 12, 8, 8, 8, 8, 16, 8, 28, 16, 32, 32, 20, 32, 20, 32, 24, 12, 16, 20, 20, 8, 16, 20};
static const int fma_equiv[M_COUNT] = {24, 24, 0, 0, 0, 0, 0, 0, 24, 24, 24, 24, 24, 24, 24, 24, 16, 24, 24, 24, 24, 0, 24, 24};

template <int M>
__global__ void __launch_bounds__(256) k(const float* __restrict__ in, float* __restrict__ out, int iters, unsigned long long* cyc) {
    float f[24]; uint64_t F[12]; float g[8]; int n[8];
    const float s = in[threadIdx.x & 31], t = in[32 + (threadIdx.x & 31)];
    uint64_t S, T; asm("mov.b64 %0, {%1,%2};" : "=l"(S) : "f"(s), "f"(s)); asm("mov.b64 %0, {%1,%2};" : "=l"(T) : "f"(t), "f"(t));
#pragma unroll
    for (int i = 0; i < 24; ++i) f[i] = s + i;
#pragma unroll
    for (int i = 0; i < 12; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(F[i]) : "f"(s + i), "f"(t - i));
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = t * (i + 1); n[i] = threadIdx.x + i; }
    const int ni = (int)(s * 3.f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (M == M_FFMA) {
#pragma unroll
            for (int i = 0; i < 24; ++i) FFMA(f[i], s, t);
        } else if (M == M_FFMA2) {
#pragma unroll
            for (int i = 0; i < 12; ++i) FFMA2(F[i], S, T);
        } else if (M == M_FMIN) {
#pragma unroll
            for (int i = 0; i < 8; ++i) FMIN(g[i], f[i]);
        } else if (M == M_FMIN3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) FMIN3(g[i], f[i], f[i + 8]);
        } else if (M == M_IADD) {
#pragma unroll
            for (int i = 0; i < 8; ++i) IADD(n[i], ni);
        } else if (M == M_IMNMX) {
#pragma unroll
            for (int i = 0; i < 8; ++i) IMNMX(n[i], ni);
        } else if (M == M_SETPSEL) {
#pragma unroll
            for (int i = 0; i < 8; ++i) SETPSEL(g[i], f[i], f[i + 8]);
        } else if (M == M_LOP) {
#pragma unroll
            for (int i = 0; i < 8; ++i) LOP(n[i], ni);
        } else if (M == M_FFMA24_FMIN3_4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { FFMA(f[6*i], s, t); FFMA(f[6*i+1], s, t); FFMA(f[6*i+2], s, t); FMIN3(g[i], s, t); FFMA(f[6*i+3], s, t); FFMA(f[6*i+4], s, t); FFMA(f[6*i+5], s, t); }
        } else if (M == M_FFMA2_12_FMIN3_4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { FFMA2(F[3*i], S, T); FFMA2(F[3*i+1], S, T); FMIN3(g[i], s, t); FFMA2(F[3*i+2], S, T); }
        } else if (M == M_FFMA24_FMIN_8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { FFMA(f[3*i], s, t); FFMA(f[3*i+1], s, t); FMIN(g[i], s); FFMA(f[3*i+2], s, t); }
        } else if (M == M_FFMA24_IADD_8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { FFMA(f[3*i], s, t); FFMA(f[3*i+1], s, t); IADD(n[i], ni); FFMA(f[3*i+2], s, t); }
        } else if (M == M_FFMA2_12_IADD_8) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { FFMA2(F[3*i], S, T); IADD(n[2*i], ni); FFMA2(F[3*i+1], S, T); IADD(n[2*i+1], ni); FFMA2(F[3*i+2], S, T); }
        } else if (M == M_FFMA24_SETPSEL_4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { FFMA(f[6*i], s, t); FFMA(f[6*i+1], s, t); FFMA(f[6*i+2], s, t); SETPSEL(g[i], s, t); FFMA(f[6*i+3], s, t); FFMA(f[6*i+4], s, t); FFMA(f[6*i+5], s, t); }
        } else if (M == M_FFMA2_12_FMIN_8) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { FFMA2(F[3*i], S, T); FMIN(g[2*i], s); FFMA2(F[3*i+1], S, T); FMIN(g[2*i+1], s); FFMA2(F[3*i+2], S, T); }
        } else if (M == M_FFMA24_LOP_8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { FFMA(f[3*i], s, t); FFMA(f[3*i+1], s, t); LOP(n[i], ni); FFMA(f[3*i+2], s, t); }
        } else if (M == M_FFMA2P_12) {
#pragma unroll
            for (int i = 0; i < 12; ++i) FFMA2P(F[i], F[(i + 5) % 12], f[i]);
        } else if (M == M_FFMA2P_12_FMIN3_4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { FFMA2P(F[3*i], F[(3*i + 5) % 12], f[i]); FFMA2P(F[3*i+1], F[(3*i + 6) % 12], f[i+4]); FMIN3(g[i], s, t); FFMA2P(F[3*i+2], F[(3*i + 7) % 12], f[i+8]); }
        } else if (M == M_FFMA2P_12_FMIN3_8) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { FFMA2P(F[3*i], F[(3*i + 5) % 12], f[i]); FMIN3(g[2*i], s, t); FFMA2P(F[3*i+1], F[(3*i + 6) % 12], f[i+4]); FMIN3(g[2*i+1], s, t); FFMA2P(F[3*i+2], F[(3*i + 7) % 12], f[i+8]); }
        } else if (M == M_FFMA2P_12_FMIN_8) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { FFMA2P(F[3*i], F[(3*i + 5) % 12], f[i]); FMIN(g[2*i], s); FFMA2P(F[3*i+1], F[(3*i + 6) % 12], f[i+4]); FMIN(g[2*i+1], s); FFMA2P(F[3*i+2], F[(3*i + 7) % 12], f[i+8]); }
        } else if (M == M_VIMIN3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { VIMIN3(n[i], __float_as_int(f[i]) + it, __float_as_int(f[i + 8])); }
        } else if (M == M_FFMA2P_12_VIMIN3_4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { FFMA2P(F[3*i], F[(3*i + 5) % 12], f[i]); FFMA2P(F[3*i+1], F[(3*i + 6) % 12], f[i+4]); VIMIN3(n[i], ni + it, __float_as_int(t)); FFMA2P(F[3*i+2], F[(3*i + 7) % 12], f[i+8]); }
        } else if (M == M_FFMA2P_12_VIMIN3_8) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { FFMA2P(F[3*i], F[(3*i + 5) % 12], f[i]); VIMIN3(n[2*i], ni + it, __float_as_int(t)); FFMA2P(F[3*i+1], F[(3*i + 6) % 12], f[i+4]); VIMIN3(n[2*i+1], ni - it, __float_as_int(s)); FFMA2P(F[3*i+2], F[(3*i + 7) % 12], f[i+8]); }
        } else if (M == M_FFMA16_FMIN3_8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { FFMA(f[2*i], s, t); FMIN3(g[i], s, t); FFMA(f[2*i+1], s, t); }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) r += f[i];
#pragma unroll
    for (int i = 0; i < 12; ++i) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(F[i])); r += a + b; }
#pragma unroll
    for (int i = 0; i < 8; ++i) r += g[i] + (float)n[i];
    if (r == 123.456f) out[threadIdx.x] = r;
    if (threadIdx.x == 0) atomicMax(cyc, (unsigned long long)(t1 - t0));
}

template <int M> void launch(int grid, const float* in, float* out, int iters, unsigned long long* cyc) { k<M><<<grid, 256>>>(in, out, iters, cyc); }
typedef void (*launch_fn)(int, const float*, float*, int, unsigned long long*);
template <int... Ms> struct Table { static constexpr launch_fn fns[sizeof...(Ms)] = {launch<Ms>...}; };
using T = Table<0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23>;

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.5f + 0.001f * i;
    float *din, *dout; unsigned long long* dc;
    CK(cudaMalloc(&din, sizeof h)); CK(cudaMalloc(&dout, 4096)); CK(cudaMalloc(&dc, 8));
    CK(cudaMemcpy(din, h, sizeof h, cudaMemcpyHostToDevice));
    const int iters = 20000, occ = 4;
    printf("{\"device\": \"%s\", \"sms\": %d, \"probes\": {", prop.name, sms);
    for (int m = 0; m < M_COUNT; ++m) {
        unsigned long long best = ~0ull;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemset(dc, 0, 8));
            T::fns[m](sms * occ, din, dout, iters, dc);
            CK(cudaDeviceSynchronize());
            unsigned long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
            if (rep && c < best) best = c;
        }
        // cycles per loop iteration per SMSP: 8 warps/CTA * occ CTAs / 4 SMSPs = 8 warps per SMSP
        const double cyc_per_iter_warp = (double)best / iters / (8.0 * occ / 4.0);
        printf("%s\"%s\": {\"cycles_per_iter_per_warp_smsp\": %.3f, \"instrs\": %d, \"fma_scalar_equiv\": %d, \"issue_per_clk_smsp\": %.3f}",
               m ? ", " : "", names[m], cyc_per_iter_warp, instrs[m], fma_equiv[m], instrs[m] / cyc_per_iter_warp);
    }
    printf("}}\n");
    return 0;
}
