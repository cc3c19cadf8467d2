// cub_bar.cu -- the "library bar" (VERDICT r1 item 3): cub::DeviceRadixSort::SortPairs and
// cub::DeviceScan::ExclusiveSum behind a tiny C ABI, so tests can assert byte equality with the hand-written sorts and
// tools / bench.py can print the library's time beside ours. Measurement infrastructure only: nothing in the product
// links or loads this file (built into tools/bin/libcub_bar.so by __graft_entry__.build()).
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

extern "C" {

// Stable ascending sort of (key, payload) pairs on bits [beginBit, endBit). Scratch is sized and allocated per call.
// Returns 0 or a cudaError_t. If ms != nullptr the sort is run `reps` more times between two events (temp storage
// reused, inputs re-copied outside the timed region is not possible with in-place semantics, so the timed runs sort the
// ORIGINAL unsorted input each time from a pristine copy made on the device before the events).
int cub_sort_pairs(const void* keysIn, const uint32_t* valsIn, void* keysOut, uint32_t* valsOut, uint32_t n, int keyBits,
                   int beginBit, int endBit, void* stream, int reps, float* ms) {
    cudaStream_t s = (cudaStream_t)stream;
    size_t tempBytes = 0;
    cudaError_t e;
    if (keyBits == 16)
        e = cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, (const uint16_t*)keysIn, (uint16_t*)keysOut, valsIn, valsOut, (int)n, beginBit, endBit, s);
    else
        e = cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, (const uint32_t*)keysIn, (uint32_t*)keysOut, valsIn, valsOut, (int)n, beginBit, endBit, s);
    if (e != cudaSuccess) return (int)e;
    void* temp = nullptr;
    if ((e = cudaMalloc(&temp, tempBytes ? tempBytes : 1)) != cudaSuccess) return (int)e;
    auto run = [&]() {
        return keyBits == 16
            ? cub::DeviceRadixSort::SortPairs(temp, tempBytes, (const uint16_t*)keysIn, (uint16_t*)keysOut, valsIn, valsOut, (int)n, beginBit, endBit, s)
            : cub::DeviceRadixSort::SortPairs(temp, tempBytes, (const uint32_t*)keysIn, (uint32_t*)keysOut, valsIn, valsOut, (int)n, beginBit, endBit, s);
    };
    e = run();
    if (e == cudaSuccess && ms && reps > 0) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, s);
        for (int i = 0; i < reps && e == cudaSuccess; ++i) e = run();
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float t = 0;
        cudaEventElapsedTime(&t, a, b);
        *ms = t / reps;
        cudaEventDestroy(a); cudaEventDestroy(b);
    }
    cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(temp);
    return (int)(e != cudaSuccess ? e : e2);
}

int cub_exclusive_sum(const uint32_t* in, uint32_t* out, uint32_t n, void* stream, int reps, float* ms) {
    cudaStream_t s = (cudaStream_t)stream;
    size_t tempBytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tempBytes, in, out, (int)n, s);
    if (e != cudaSuccess) return (int)e;
    void* temp = nullptr;
    if ((e = cudaMalloc(&temp, tempBytes ? tempBytes : 1)) != cudaSuccess) return (int)e;
    e = cub::DeviceScan::ExclusiveSum(temp, tempBytes, in, out, (int)n, s);
    if (e == cudaSuccess && ms && reps > 0) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, s);
        for (int i = 0; i < reps && e == cudaSuccess; ++i) e = cub::DeviceScan::ExclusiveSum(temp, tempBytes, in, out, (int)n, s);
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float t = 0;
        cudaEventElapsedTime(&t, a, b);
        *ms = t / reps;
        cudaEventDestroy(a); cudaEventDestroy(b);
    }
    cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(temp);
    return (int)(e != cudaSuccess ? e : e2);
}

}  // extern "C"
