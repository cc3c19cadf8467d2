// Same-address global atomic throughput (design input for ticket / arrival-counter schemes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/atomic_rate tools/micro/atomic_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_red(unsigned long long* p, int iters, int stride) {
    // one lane per warp issues; `stride` words apart per CTA (0 = all CTAs on one address)
    unsigned long long* q = p + (size_t)blockIdx.x * stride;
    if ((threadIdx.x & 31) == 0)
        for (int i = 0; i < iters; ++i) atomicAdd(q, 1ull);
}
__global__ void k_atom(unsigned int* p, int iters, unsigned int* sink) {
    unsigned int acc = 0;
    if ((threadIdx.x & 31) == 0)
        for (int i = 0; i < iters; ++i) acc += atomicAdd(p, 1u);  // returning form: each waits for the previous
    if (acc == 0xFFFFFFFFu) *sink = acc;
}
int main() {
    unsigned long long* d; unsigned int* d32; unsigned int* sink;
    cudaMalloc(&d, 1 << 24); cudaMalloc(&d32, 1024); cudaMalloc(&sink, 4);
    cudaMemset(d, 0, 1 << 24); cudaMemset(d32, 0, 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int stride : {0, 1, 16, 4096}) {
        for (int grid : {148, 148 * 8}) {
            const int iters = 256, warps = 4;
            k_red<<<grid, 32 * warps>>>(d, 8, stride); cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_red<<<grid, 32 * warps>>>(d, iters, stride);
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double n = (double)grid * warps * iters;
            printf("RED  stride %5d words grid %5d: %.1f us, %.2f ns per atomic (%.0f atomics)\n", stride, grid, ms * 1e3, ms * 1e6 / n, n);
        }
    }
    for (int grid : {148, 148 * 8}) {
        const int iters = 64, warps = 1;
        cudaEventRecord(e0);
        k_atom<<<grid, 32 * warps>>>(d32, iters, sink);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double n = (double)grid * warps * iters;
        printf("ATOM (returning, 1 address) grid %5d: %.1f us, %.2f ns per atomic, %.2f us per dependent round trip\n", grid, ms * 1e3, ms * 1e6 / n, ms * 1e3 / iters);
    }
    return 0;
}
