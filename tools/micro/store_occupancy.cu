// Microbenchmark: achieved write bandwidth of the warp_span pattern (each warp streams its own
// 24 KB span) as a function of resident warps per SM, store width (128 / 256 bit) and a dependent
// ALU chain of WORK instructions between consecutive stores (emulates the renderer's bookkeeping).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_occupancy store_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void st256(void* p, uint32_t v) {
    asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "r"(v) : "memory");
}

template <int WIDTH, int UNROLL, int WORK>
__global__ void __launch_bounds__(256) k(uint8_t* out, unsigned seed) {
    extern __shared__ uint32_t pad_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t item = (size_t)blockIdx.x * 8 + warp;
    uint8_t* p = out + item * 24576 + lane * (WIDTH / 8);
    const int iters = 24576 / (32 * (WIDTH / 8));
    uint32_t v = seed + lane;
#pragma unroll UNROLL
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int w = 0; w < WORK; ++w) v = v * 1664525u + 1013904223u;   // dependent chain
        if (WIDTH == 128) __stcs(reinterpret_cast<uint4*>(p), make_uint4(v, v, v, v));
        else st256(p, v);
        p += 32 * (WIDTH / 8);
    }
    if (v == 0x12345678u) pad_smem[0] = v;
}

template <int WIDTH, int UNROLL, int WORK>
void run(uint8_t* d, size_t bytes, int ctas_per_sm) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = (int)(bytes / 24576 / 8);
    // limit residency with dynamic shared memory: 227 KB / ctas_per_sm
    const int smem = ctas_per_sm >= 8 ? 0 : (227 * 1024 / ctas_per_sm) - 1024;
    cudaFuncSetAttribute(k<WIDTH, UNROLL, WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float best = 1e9;
    for (int r = 0; r < 8; ++r) {
        cudaEventRecord(a);
        k<WIDTH, UNROLL, WORK><<<grid, 256, smem>>>(d, r);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r >= 2 && ms < best) best = ms;
    }
    printf("width=%d unroll=%d work=%2d ctas/sm=%d (warps/sm=%2d)  %.4f ms  %5.0f GB/s %s\n", WIDTH, UNROLL, WORK,
           ctas_per_sm, ctas_per_sm * 8, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t bytes = (size_t)4096 * 393216;
    uint8_t* d; cudaMalloc(&d, bytes);
    for (int c : {1, 2, 3, 4, 6, 8}) run<128, 1, 0>(d, bytes, c);
    for (int c : {1, 2, 3, 4, 6, 8}) run<128, 4, 0>(d, bytes, c);
    for (int c : {1, 2, 3, 4, 6, 8}) run<256, 1, 0>(d, bytes, c);
    for (int c : {1, 2, 3, 4, 6, 8}) run<256, 4, 0>(d, bytes, c);
    for (int c : {2, 4, 8}) run<128, 4, 16>(d, bytes, c);
    for (int c : {2, 4, 8}) run<128, 4, 32>(d, bytes, c);
    for (int c : {2, 4, 8}) run<256, 4, 32>(d, bytes, c);
    for (int c : {2, 4, 8}) run<256, 4, 64>(d, bytes, c);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
