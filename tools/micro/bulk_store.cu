// Microbenchmark: write-only stream made of small TMA bulk stores (cp.async.bulk shared -> global)
// issued per lane, as the column renderer would: each column of 768 B is written as `PIECES`
// bulk copies out of a constant pattern buffer in shared memory.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_store bulk_store.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"((uint32_t)__cvta_generic_to_shared(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// MODE 0: every lane writes its column (768 B) as PIECES bulk stores
// MODE 1: lane 0 writes the warp's 24 KB span as 32 bulk stores of 768 B (fewer issuing threads)
// MODE 2: lane 0 writes the warp's 24 KB span as ONE bulk store
template <int MODE, int PIECES>
__global__ void __launch_bounds__(256) k(uint8_t* out, int items_per_warp, unsigned seed) {
    __shared__ __align__(128) uint8_t pat[24576 + 64];
    for (int i = threadIdx.x; i < (24576 + 64) / 4; i += 256) ((uint32_t*)pat)[i] = seed + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int it = 0; it < items_per_warp; ++it) {
        const size_t item = ((size_t)blockIdx.x * items_per_warp + it) * 8 + warp;
        uint8_t* span = out + item * 24576;
        if (MODE == 0) {
            uint8_t* col = span + lane * 768;
            constexpr int sz = 768 / PIECES;
#pragma unroll
            for (int p = 0; p < PIECES; ++p) bulk_s2g(col + p * sz, pat + ((p * sz) % 48), sz);
        } else if (MODE == 1) {
            if (lane == 0)
                for (int c = 0; c < 32; ++c) bulk_s2g(span + c * 768, pat, 768);
        } else {
            if (lane == 0) bulk_s2g(span, pat, 24576);
        }
        bulk_commit();
    }
    bulk_wait_read0();
}

template <int MODE, int PIECES>
void run(const char* name, uint8_t* d, size_t bytes, int items_per_warp) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = (int)(bytes / 24576 / 8 / items_per_warp);
    float best = 1e9;
    for (int r = 0; r < 12; ++r) {
        cudaEventRecord(a);
        k<MODE, PIECES><<<grid, 256>>>(d, items_per_warp, r);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r >= 2 && ms < best) best = ms;
    }
    printf("%-22s pieces=%d items/warp=%d grid=%6d  %.4f ms  %.0f GB/s  (%s)\n", name, PIECES, items_per_warp, grid, best,
           bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t bytes = (size_t)4096 * 393216;
    uint8_t* d; cudaMalloc(&d, bytes);
    run<0, 1>("lane_per_column", d, bytes, 1);
    run<0, 3>("lane_per_column", d, bytes, 1);
    run<0, 6>("lane_per_column", d, bytes, 1);
    run<0, 12>("lane_per_column", d, bytes, 1);
    run<0, 3>("lane_per_column", d, bytes, 4);
    run<0, 3>("lane_per_column", d, bytes, 16);
    run<1, 1>("lane0_32x768", d, bytes, 1);
    run<2, 1>("lane0_1x24576", d, bytes, 1);
    run<2, 1>("lane0_1x24576", d, bytes, 16);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
