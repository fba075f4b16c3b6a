// microbench5.cu — the floor under a one-launch search iteration: how long does "host launches a kernel, the
// kernel's last CTA writes a sequence number into pinned host memory, the host sees it" take on this box, for an empty grid, for the
// grid shape of the small-palette kernel (4 x 111 CTAs of 256 threads, ticket counter, 36 result words exported), and with a
// kernel-parameter block of the one-launch evaluation's size?  Prints one JSON object.  nvcc -O3 -arch=sm_100a.
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

struct Params512 { float v[128]; };

__global__ void flag_only(volatile unsigned long long* flag, unsigned long long seq) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) { __threadfence_system(); *flag = seq; }
}
__global__ void ticket_export(const __grid_constant__ Params512 pal, unsigned* counter, const unsigned long long* src, unsigned long long* dst,
                              volatile unsigned long long* flag, unsigned long long seq, float* sink) {
    __shared__ unsigned s_ticket;
    if (pal.v[threadIdx.x & 127] == 123456.f) sink[0] = 1.f;   // the parameters are read
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(counter, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x * gridDim.y - 1) return;
    __threadfence();
    for (unsigned i = threadIdx.x; i < 36; i += blockDim.x) dst[i] = __ldcg(src + i);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { *counter = 0u; *flag = seq; }
}

template <typename F>
static double run(F launch, volatile unsigned long long* h_flag, int iters) {
    unsigned long long seq = 0;
    for (int i = 0; i < 200; ++i) { launch(++seq); while (*h_flag != seq) {} }
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) { launch(++seq); while (*h_flag != seq) {} }
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
}

int main() {
    unsigned long long *h_flag, *h_dst, *d_src;
    unsigned* d_counter;
    float* d_sink;
    cudaHostAlloc(&h_flag, 64, cudaHostAllocPortable);
    cudaHostAlloc(&h_dst, 4096, cudaHostAllocPortable);
    cudaMalloc(&d_src, 4096); cudaMemset(d_src, 0, 4096);
    cudaMalloc(&d_counter, 4); cudaMemset(d_counter, 0, 4);
    cudaMalloc(&d_sink, 4);
    *h_flag = 0;
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    Params512 pal = {};
    const int iters = 20000;
    const double a = run([&](unsigned long long s) { flag_only<<<1, 32, 0, st>>>(h_flag, s); }, h_flag, iters);
    const double b = run([&](unsigned long long s) { flag_only<<<dim3(4, 111), 256, 0, st>>>(h_flag, s); }, h_flag, iters);
    const double c = run([&](unsigned long long s) { ticket_export<<<dim3(4, 111), 256, 0, st>>>(pal, d_counter, d_src, h_dst, h_flag, s, d_sink); }, h_flag, iters);
    const double d = run([&](unsigned long long s) { ticket_export<<<dim3(4, 111), 256, 0, st>>>(pal, d_counter, d_src, h_dst, h_flag, s, d_sink); cudaStreamQuery(st); }, h_flag, iters);
    printf("{\"what\": \"us per launch -> flag seen by the host (spin on pinned memory), %d iterations each\", \"one_warp_flag_only\": %.2f, "
           "\"grid_4x111x256_flag_only\": %.2f, \"grid_4x111x256_ticket_and_36_word_export_512B_params\": %.2f, \"same_plus_cudaStreamQuery\": %.2f}\n",
           iters, a, b, c, d);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
