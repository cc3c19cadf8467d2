// Per-tile timeline of one onesweep digit pass (diagnostic; not part of the product library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -DGSM_SORT_TRACE -I gsm_renderer_b200/csrc -I include
//        -o tools/bin/sort_trace tools/sort_trace.cu
//   tools/bin/sort_trace N keyBits > trace.csv     (columns: pass, tile, then ns since the pass's first CTA entry)
#include "../gsm_renderer_b200/csrc/sort.cu"
namespace gsm { bool pdlEnabled() { const char* e = getenv("GSM_PDL"); return !(e && e[0] == '0'); } }
#include <cstdio>
#include <cstdlib>
#include <cstdlib>
#include <vector>
#include <random>
using namespace gsm;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atol(argv[1]) : 709202u;
    const int keyBits = argc > 2 ? atoi(argv[2]) : 32;
    const int passes = keyBits / 8;
    const bool large = keyBits == 32 && n >= 3000000u;
    const uint32_t tile = sortTileSize(keyBits, large), tiles = (n + tile - 1) / tile;
    int dev = 0, numSMs = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev));
    std::mt19937 rng(1);
    std::vector<uint32_t> hk(n);
    for (auto& k : hk) k = keyBits == 32 ? rng() : rng() % 8160u;
    std::vector<uint16_t> hk16(hk.begin(), hk.end());
    const size_t keyBytes = (size_t)n * (keyBits / 8);
    void *k0, *k1; uint32_t *v0, *v1, *cnt, *hist, *status, *gstatus, *tickets; unsigned long long* trace;
    CK(cudaMalloc(&k0, keyBytes)); CK(cudaMalloc(&k1, keyBytes));
    CK(cudaMalloc(&v0, (size_t)n * 4)); CK(cudaMalloc(&v1, (size_t)n * 4));
    CK(cudaMalloc(&cnt, 4)); CK(cudaMalloc(&hist, 4 * 256 * 4)); CK(cudaMalloc(&tickets, 16));
    const size_t stBytes = (size_t)passes * tiles * 256 * 4, gstBytes = (size_t)passes * sortGroupRows(tiles) * 256 * 4;
    CK(cudaMalloc(&status, stBytes)); CK(cudaMalloc(&gstatus, gstBytes));
    const size_t trBytes = (size_t)passes * tiles * 16 * 8;
    CK(cudaMalloc(&trace, trBytes));
    CK(cudaMemcpy(cnt, &n, 4, cudaMemcpyHostToDevice));
    void* flush; CK(cudaMalloc(&flush, 256u << 20));
    std::vector<unsigned long long> ht((size_t)passes * tiles * 16);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemcpy(k0, keyBits == 32 ? (void*)hk.data() : (void*)hk16.data(), keyBytes, cudaMemcpyHostToDevice));
        CK(cudaMemset(hist, 0, 4 * 256 * 4)); CK(cudaMemset(tickets, 0, 16));
        CK(cudaMemset(status, 0, stBytes)); CK(cudaMemset(gstatus, 0, gstBytes)); CK(cudaMemset(trace, 0, trBytes));
        if (rep & 1) CK(cudaMemset(flush, rep, 256u << 20));  // odd reps start with a cold L2
        SortPlan p;
        p.k0 = k0; p.k1 = k1; p.v0 = v0; p.v1 = v1; p.countPtr = cnt; p.countCap = n; p.hist = hist; p.status = status;
        p.gstatus = gstatus; p.tickets = tickets; p.tilesCap = tiles; p.keyBits = keyBits; p.numPasses = passes; p.numSMs = numSMs;
        p.largeTiles = large; p.histogramReady = false;
        // one launch per pass so that every pass gets its own trace slice
        unsigned long long* tr = trace;
        CK(cudaMemcpyToSymbol(g_sortTrace, &tr, sizeof(tr)));
        CK(cudaEventRecord(e0));
        CK(launchSort(0, p));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(ht.data(), trace, trBytes, cudaMemcpyDeviceToHost));
        if (rep >= 2) {
            for (int ps = 0; ps < passes; ++ps) {
                unsigned long long t0 = ~0ull;
                for (uint32_t t = 0; t < tiles; ++t) { unsigned long long v = ht[((size_t)ps * tiles + t) * 16]; if (v && v < t0) t0 = v; }
                for (uint32_t t = 0; t < tiles; ++t) {
                    printf("%d,%d,%d", rep, ps, t);
                    for (int k = 0; k < 8; ++k) printf(",%lld", (long long)(ht[((size_t)ps * tiles + t) * 16 + k] - t0));
                    printf("\n");
                }
            }
        }
        fprintf(stderr, "rep %d (%s L2): n=%u keyBits=%d tiles=%u sort %.1f us (histogram + %d passes)\n", rep, (rep & 1) ? "cold" : "warm", n, keyBits, tiles, ms * 1e3f, passes);
    }
    return 0;
}
