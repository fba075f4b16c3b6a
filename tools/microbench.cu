// microbench.cu — measures the FP32 CUDA-core ceilings the assign kernel is judged against
// ("P_fp32 is not in MEASURED_PEAKS.json: measure an FFMA-saturating
// microbenchmark in the same gpurun job").  Prints one JSON object.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench tools/microbench.cu
//
// Patterns (per thread, 8 independent chains, 256 threads/CTA, `occ` CTAs per SM):
//   ffma      scalar FFMA                     -> flop/clk/SM, TFLOP/s
//   ffma2     packed fma.rn.f32x2             -> same units, counts 4 flop / instr
//   fadd2     packed add.rn.f32x2
//   fmul2     packed mul.rn.f32x2
//   dist2     the assign inner body: 3 FADD2 + FMUL2 + 2 FFMA2 per colour pair (8 flop*2)
//   dist2min  dist2 + FMNMX3 per colour pair
//   fmnmx     scalar FMNMX (alu pipe)
//   fmnmx3    FMNMX3
//   setpsel   FSETP + 2 SEL (direct index tracking)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(uint64_t v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(uint64_t v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fmin1(float a, float b) { float d; asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

enum { P_FFMA, P_FFMA2, P_FADD2, P_FMUL2, P_DIST2, P_DIST2MIN, P_FMNMX, P_FMNMX3, P_SETPSEL, P_COUNT };

template <int PAT>
__global__ void __launch_bounds__(256) bench(const float* __restrict__ in, float* __restrict__ out, int iters, unsigned long long* cyc) {
    const float s0 = in[threadIdx.x & 31], s1 = in[32 + (threadIdx.x & 31)], s2 = in[64 + (threadIdx.x & 31)];
    float a[8];
    uint64_t A[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = s0 + i; A[i] = pk(s0 + i, s1 - i); }
    const uint64_t B = pk(s1, s2), C = pk(s2, s0);
    float m0 = 3e38f, m1 = 3e38f, m2 = 3e38f, m3 = 3e38f;
    int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
    float z[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i] = s2 * (float)(i + 1);
    __shared__ float smem[512 + 256];
    for (int i = threadIdx.x; i < 768; i += 256) smem[i] = in[i & 63] * (float)(1 + (i >> 6));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (PAT == P_FFMA) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = ffma(a[i], s1, s2);
        } else if (PAT == P_FFMA2) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) A[i] = fma2(A[i], B, C);
        } else if (PAT == P_FADD2) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) A[i] = add2(A[i], B);
        } else if (PAT == P_FMUL2) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) A[i] = mul2(A[i], B);
        } else if (PAT == P_DIST2 || PAT == P_DIST2MIN) {
            // the assign inner body: 4 pixels (a[0..3], a[4..7], s0..) x 8 colours (4 pairs)
            // read from shared memory exactly as the real kernel does
            const float4* pla = reinterpret_cast<const float4*>(smem) + (it & 31) * 4;
            const float2* pb = reinterpret_cast<const float2*>(smem + 512) + (it & 31) * 4;
            float c0 = 3e38f, c1 = 3e38f, c2 = 3e38f, c3 = 3e38f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 la = pla[q];
                const float2 bb = pb[q];
                const uint64_t P0 = pk(la.x, la.y), P1 = pk(la.z, la.w), P2 = pk(bb.x, bb.y);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t X = pk(a[j], a[j]), Y = pk(a[4 + j], a[4 + j]), Z = pk(z[j], z[j]);
                    const uint64_t d0 = sub2(X, P0), d1 = sub2(Y, P1), d2 = sub2(Z, P2);
                    uint64_t acc = mul2(d0, d0);
                    acc = fma2(d1, d1, acc);
                    acc = fma2(d2, d2, acc);
                    if (PAT == P_DIST2MIN) {
                        if (j == 0) c0 = min3(c0, lo(acc), hi(acc));
                        if (j == 1) c1 = min3(c1, lo(acc), hi(acc));
                        if (j == 2) c2 = min3(c2, lo(acc), hi(acc));
                        if (j == 3) c3 = min3(c3, lo(acc), hi(acc));
                    } else {
                        if (j == 0) c0 += lo(acc) + hi(acc);  // 2 FADD stand in for the reduction
                        if (j == 1) c1 += lo(acc) + hi(acc);
                        if (j == 2) c2 += lo(acc) + hi(acc);
                        if (j == 3) c3 += lo(acc) + hi(acc);
                    }
                }
            }
            if (PAT == P_DIST2MIN) {
                if (c0 < m0) { m0 = c0; i0 = it; }
                if (c1 < m1) { m1 = c1; i1 = it; }
                if (c2 < m2) { m2 = c2; i2 = it; }
                if (c3 < m3) { m3 = c3; i3 = it; }
            } else { m0 += c0; m1 += c1; m2 += c2; m3 += c3; }
        } else if (PAT == P_FMNMX) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmin1(a[i], s1 + (float)u);
        } else if (PAT == P_FMNMX3) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = min3(a[i], s1, s2);
        } else if (PAT == P_SETPSEL) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float cand = a[(i + 1) & 7];
                    int lt;
                    asm volatile("{ .reg .pred p; setp.lt.f32 p, %1, %2; selp.s32 %0, 1, 0, p; }" : "=r"(lt) : "f"(cand), "f"(a[i]));
                    if (lt) { m0 = cand; i0 = it + i; }
                }
        }
    }
    const long long t1 = clock64();
    float r = m0 + m1 + m2 + m3 + (float)i0 + (float)i1 + (float)i2 + (float)i3;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i] + lo(A[i]) + hi(A[i]);
    if (r == 123.456f) out[threadIdx.x] = r;
    if (threadIdx.x == 0) atomicMax(cyc, (unsigned long long)(t1 - t0));
}

struct Pat { const char* name; double instr_per_iter; double flop_per_iter; };

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float h_in[96];
    for (int i = 0; i < 96; ++i) h_in[i] = 0.5f + 0.001f * i;
    float *d_in, *d_out;
    unsigned long long* d_cyc;
    CK(cudaMalloc(&d_in, sizeof h_in)); CK(cudaMalloc(&d_out, 4096)); CK(cudaMalloc(&d_cyc, 8));
    CK(cudaMemcpy(d_in, h_in, sizeof h_in, cudaMemcpyHostToDevice));
    // per-thread, per-iteration instruction / flop counts
    const Pat pats[P_COUNT] = {
        {"ffma", 32, 64}, {"ffma2", 32, 128}, {"fadd2", 32, 64}, {"fmul2", 32, 64},
        {"dist2", 96 + 48, 8 * 8 * 4}, {"dist2min", 96 + 16 + 12, 8 * 8 * 4}, {"fmnmx", 32, 0}, {"fmnmx3", 32, 0}, {"setpsel", 96, 0}};
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_rate_khz\": %d, \"patterns\": {", prop.name, sms, clk_khz);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
    for (int p = 0; p < P_COUNT; ++p) {
        for (int occ = 1; occ <= 4; occ *= 2) {  // CTAs (of 8 warps) per SM
            const int grid = sms * occ;
            float best_ms = 1e30f; unsigned long long cyc = 0;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaMemset(d_cyc, 0, 8));
                CK(cudaEventRecord(e0));
                switch (p) {
                    case P_FFMA: bench<P_FFMA><<<grid, 256>>>(d_in, d_out, iters, d_cyc); break;
                    case P_FFMA2: bench<P_FFMA2><<<grid, 256>>>(d_in, d_out, iters, d_cyc); break;
                    case P_FADD2: bench<P_FADD2><<<grid, 256>>>(d_in, d_out, iters, d_cyc); break;
                    case P_FMUL2: bench<P_FMUL2><<<grid, 256>>>(d_in, d_out, iters, d_cyc); break;
                    case P_DIST2: bench<P_DIST2><<<grid, 256>>>(d_in, d_out, iters, d_cyc); break;
                    case P_DIST2MIN: bench<P_DIST2MIN><<<grid, 256>>>(d_in, d_out, iters, d_cyc); break;
                    case P_FMNMX: bench<P_FMNMX><<<grid, 256>>>(d_in, d_out, iters, d_cyc); break;
                    case P_FMNMX3: bench<P_FMNMX3><<<grid, 256>>>(d_in, d_out, iters, d_cyc); break;
                    case P_SETPSEL: bench<P_SETPSEL><<<grid, 256>>>(d_in, d_out, iters, d_cyc); break;
                }
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep > 0 && ms < best_ms) { best_ms = ms; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); }
            }
            const double thr_instr = pats[p].instr_per_iter * iters * 256.0 * occ;  // thread-instr per SM
            const double warp_instr_per_clk = thr_instr / 32.0 / (double)cyc;      // per SM
            const double tflops = pats[p].flop_per_iter * iters * 256.0 * grid / (best_ms * 1e-3) / 1e12;
            printf("%s\"%s_occ%d\": {\"ms\": %.4f, \"cycles\": %llu, \"warp_instr_per_clk_per_sm\": %.3f, \"flop_per_clk_per_sm\": %.1f, \"tflops\": %.2f, \"eff_mhz\": %.0f}",
                   (p == 0 && occ == 1) ? "" : ", ", pats[p].name, occ, best_ms, cyc, warp_instr_per_clk,
                   pats[p].flop_per_iter * iters * 256.0 * occ / (double)cyc, tflops, (double)cyc / (best_ms * 1e-3) / 1e6);
        }
    }
    printf("}}\n");
    return 0;
}
