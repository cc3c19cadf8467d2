// blend.cu -- stage 8: front-to-back alpha blending of 16x16 tiles, 8x8 threads x 2x2 pixels, all pixel
// arithmetic in binary16 exactly as depthFirstRender (DFS.metal:1703-1811) and depthFirstStereoRender
// (DFS.metal:1825-1982) do it, with the clear (DFS.metal:2020-2034, :1813-1823) and the stereo copy
// (DFS.metal:1984-2018) fused in.
//
// B200 mapping: one CTA per tile; the tile's splat list is staged through shared memory in chunks (one
// 32-byte pre-expanded record per splat, gathered once per tile instead of once per thread); the four pixels
// of a thread are two half2 rows so every half op is a packed HADD2/HMUL2; exp() is the canonical
// polynomial of gsm_dmath.cuh; a CTA-wide vote stops fetching once every thread has closed its quad.
// Per-thread semantics are untouched: a thread's early exit depends only on its own four transmittances
// (quirk Q8), and unfused mul/add keep one rounding per reference operation.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "gsm_common.cuh"
#include "gsm_dmath.cuh"
#include "gsm_kernels.h"

// A/B knobs of the mono blend (tools/build_variant.sh builds variants; the defaults are the measured best)
#ifndef GSM_BLEND_GROUPS
#define GSM_BLEND_GROUPS 5
#endif
#ifndef GSM_BLEND_CTAS
#define GSM_BLEND_CTAS 2       // resident CTAs per SM the launch bounds ask for
#endif
#ifndef GSM_BLEND_PREFETCH
#define GSM_BLEND_PREFETCH 1       // 1: records one chunk and indices two chunks ahead (C2 blend 124 -> 119 us at 4 groups, 114 us at 5); 2: indices only
#endif
#ifndef GSM_BLEND_TABLE
#define GSM_BLEND_TABLE 0      // 1: exact exp table in shared memory instead of the polynomial. Measured SLOWER (C2 blend 162.8 us
                               // against 146.8 us at the same shape, profiles/r2_blend_ab*.txt): four bank-conflicting 2-byte
                               // loads per (thread, splat) cost more LSU time than the 28 FMA-pipe operations they replace
#endif

namespace gsm {

constexpr int kBlendThreads = 64;
constexpr int kBlendChunk = 64;

__device__ __forceinline__ __half2 h2(float v) { return __float2half2_rn(v); }
__device__ __forceinline__ uint32_t h2bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }

// exp(-0.5h * p) for two packed pairs. EXPM 1: the guard-free XU-pipe form of gsm_dmath.cuh (MUFU.EX2 on a tuned argument, one
// exceptional input sent to the polynomial) -- C2 mono blend 147 -> 123 us, the FMA pipe loses 28 of its 47 operations per (warp,
// splat); EXPM 0: the canonical polynomial itself. Bit-identical on all inputs (blendExpSelfTest below decides per device).
template <int EXPM>
__device__ __forceinline__ void expNegHalfPairs(__half2 p0, __half2 p1, __half2& e0, __half2& e1) {
    if (EXPM == 1) {
        dhexp2_neghalf_tuned(p0, p1, e0, e1);
    } else {
        e0 = dhexp2_neghalf_packed(p0);
        e1 = dhexp2_neghalf_packed(p1);
    }
}

struct QuadState {
    __half2 T0, T1;             // transmittance rows (x, x+1)
    __half2 r0, g0, b0, d0;     // row 0 accumulators
    __half2 r1, g1, b1, d1;     // row 1
};

// one splat against one eye's quad; returns false if all four alphas are zero (nothing to do)
__device__ __forceinline__ void accumulate(QuadState& q, __half2 a0, __half2 a1, __half2 cr, __half2 cg, __half2 cb, __half2 cd,
                                           bool withDepth) {
    const __half2 one = h2(1.0f);
    __half2 w0 = __hmul2_rn(a0, q.T0), w1 = __hmul2_rn(a1, q.T1);
    // color += gColor * (a*T): contracted to one fused half FMA (canonical semantics)
    q.r0 = __hfma2(cr, w0, q.r0); q.r1 = __hfma2(cr, w1, q.r1);
    q.g0 = __hfma2(cg, w0, q.g0); q.g1 = __hfma2(cg, w1, q.g1);
    q.b0 = __hfma2(cb, w0, q.b0); q.b1 = __hfma2(cb, w1, q.b1);
    if (withDepth) {
        q.d0 = __hfma2(cd, w0, q.d0); q.d1 = __hfma2(cd, w1, q.d1);
    }
    q.T0 = __hmul2_rn(q.T0, __hsub2_rn(one, a0));
    q.T1 = __hmul2_rn(q.T1, __hsub2_rn(one, a1));
}

// p = dx*dx*cxx + dy*dy*cyy + dx*dy*cxy2 for one row (two pixels) (DFS.metal:1770), contracted as
// fma(dx*dy, cxy2, fma(dy*dy, cyy, (dx*dx)*cxx))
__device__ __forceinline__ __half2 power(__half2 dx, __half2 dy, __half2 cxx, __half2 cyy, __half2 cxy2) {
    __half2 t0 = __hmul2_rn(__hmul2_rn(dx, dx), cxx);
    __half2 in = __hfma2(__hmul2_rn(dy, dy), cyy, t0);
    return __hfma2(__hmul2_rn(dx, dy), cxy2, in);
}

__device__ __forceinline__ bool quadClosed(const __half2& T0, const __half2& T1, __half thr) {
    __half2 m = __hmax2(T0, T1);
    __half mm = __hmax(__low2half(m), __high2half(m));
    return __hlt(mm, thr);  // DFS.metal:1746-1747
}

__device__ __forceinline__ void storePixelRow(__half* color, __half* depth, uint32_t width, uint32_t height, uint32_t x, uint32_t y,
                                              __half2 r, __half2 g, __half2 b, __half2 a, __half2 d, uint32_t pitch = 0u) {
    if (y >= height || x >= width) return;   // `width` x `height` bound the stores; rows are `pitch` pixels apart (0 = width)
    const size_t o = (size_t)y * (pitch ? pitch : width) + x;
    // the thread's two pixels of a row are 16 contiguous bytes of colour and 4 of depth: one store each when the pair is whole and
    // aligned (x is even, so an even row pitch is enough) -- half the store instructions, and whole 128-byte lines per 8 lanes when the
    // target is a peer's memory (strip-sharded frames write the image owner's buffer over NVLink); else 8- and 2-byte stores
    const uint32_t rowPitch = pitch ? pitch : width;
    if (x + 1 < width && (rowPitch & 1u) == 0u && ((reinterpret_cast<uintptr_t>(color) & 15u) == 0u)) {
        uint4 v;
        v.x = h2bits(__halves2half2(__low2half(r), __low2half(g)));
        v.y = h2bits(__halves2half2(__low2half(b), __low2half(a)));
        v.z = h2bits(__halves2half2(__high2half(r), __high2half(g)));
        v.w = h2bits(__halves2half2(__high2half(b), __high2half(a)));
        *reinterpret_cast<uint4*>(color + 4 * o) = v;
        if (depth) {
            if ((reinterpret_cast<uintptr_t>(depth) & 3u) == 0u) *reinterpret_cast<__half2*>(depth + o) = d;
            else { depth[o] = __low2half(d); depth[o + 1] = __high2half(d); }
        }
        return;
    }
    uint2 v;  // 8-byte pixel stores: always aligned
    v.x = h2bits(__halves2half2(__low2half(r), __low2half(g)));
    v.y = h2bits(__halves2half2(__low2half(b), __low2half(a)));
    *reinterpret_cast<uint2*>(color + 4 * o) = v;
    if (depth) depth[o] = __low2half(d);
    if (x + 1 < width) {
        v.x = h2bits(__halves2half2(__high2half(r), __high2half(g)));
        v.y = h2bits(__halves2half2(__high2half(b), __high2half(a)));
        *reinterpret_cast<uint2*>(color + 4 * o + 4) = v;
        if (depth) depth[o + 1] = __high2half(d);
    }
}

// DFS.metal:1304-1312: the tile's {offset,count} header and, for a non-empty tile, one atomic append to the active list
__device__ __forceinline__ void publishTile(const TileOut& tout, uint32_t tile, uint32_t start, uint32_t count) {
    GSMGaussianHeader h;
    h.offset = start;
    h.count = count;
    tout.tileHeaders[tile] = h;
    if (count > 0) tout.activeTiles[atomicAdd(tout.activeTileCount, 1u)] = tile;
}

// tools/blend_stats.py builds this file with GSM_BLEND_STATS to count, on the bench workload, how many (warp, splat)
// evaluations of the inner loop find every lane's alphas zero (diagnostic; not in the product build).
#ifdef GSM_BLEND_STATS
__device__ unsigned long long g_blendStats[4];  // warp-evaluations, of those no lane uses the splat, lane-evaluations, of those used
__device__ __forceinline__ void blendStat(bool use) {
    const unsigned act = __activemask();
    const unsigned useMask = __ballot_sync(act, use);
    if ((threadIdx.x & 31u) == (unsigned)(__ffs(act) - 1)) {
        atomicAdd(&g_blendStats[0], 1ull);
        if (useMask == 0u) atomicAdd(&g_blendStats[1], 1ull);
        atomicAdd(&g_blendStats[2], (unsigned long long)__popc(act));
        atomicAdd(&g_blendStats[3], (unsigned long long)__popc(useMask));
    }
}
#define GSM_BLEND_STAT(use) blendStat(use)
#else
#define GSM_BLEND_STAT(use) do { } while (0)
#endif

// Staged form of one splat for one tile. p = fma(dx*dy, cxy2, fma(dy*dy, cyy, (dx*dx)*cxx)) (DFS.metal:1770) has a
// per-COLUMN part T0 = (dx*dx)*cxx and a per-ROW part dy*dy; T0, dy*dy, dx, dy are the same half operations on the
// same operands whichever thread evaluates them, so they are computed once per (splat, column pair) and (splat,
// row pair) by the staging thread instead of once per pixel (ncu r1_v2: the kernel was FMA-pipe bound at 73 %,
// 19 packed half ops per pixel quad for p; this form needs 6).
struct StagedSplat {
    uint2 col[8];  // per x pair k: {T0(x=2k), T0(2k+1)} , {dx(2k), dx(2k+1)}
    uint2 row[8];  // per y pair k: {dy*dy(y=2k), dy*dy(2k+1)} , {dy(2k), dy(2k+1)}
    uint4 m0;      // cxy2|cxy2, op|op, r|r, g|g
    uint4 m1;      // b|b, depth|depth, unused, cyy|cyy
};

__device__ __forceinline__ void zeroStaged(StagedSplat& sp) {
    uint4* w = reinterpret_cast<uint4*>(&sp);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(StagedSplat) / 16); ++k) w[k] = make_uint4(0u, 0u, 0u, 0u);
}

// exp(-0.5h * p) of both halves of p from the exact table (tab[bits(p)] == dhexp2_neghalf_packed(p), built by
// blend_exp_table_kernel from that very function, so bit-identical by construction and checked on all inputs by
// gsm_probe_math op 13). The table covers the non-negative halfs, NaNs and +inf included; p is a sum of products of
// non-negative terms plus a cross term, so a set sign bit is rare: that lane pair takes the polynomial.
// Why a table was tried: 28 of the 47 FMA-pipe operations per (warp, splat) are this polynomial (ncu r1_v14: FMA pipe 66 % busy).
// It is exact but slower (see GSM_BLEND_TABLE above), so the default build calls the polynomial; the table path stays as an
// A/B build and as a probe (gsm_probe_math op 13) that proves its bit-equality on all 65 536 inputs.
__device__ __forceinline__ __half2 expNegHalfTab(const unsigned short* __restrict__ tab, __half2 p) {
    const uint32_t b = h2bits(p);
    if (!GSM_BLEND_TABLE || (b & 0x80008000u)) return dhexp2_neghalf_packed(p);
    const uint32_t lo = tab[b & 0xFFFFu], hi = tab[b >> 16];
    const uint32_t r = lo | (hi << 16);
    return *reinterpret_cast<const __half2*>(&r);
}

// Alphas of one staged splat on this thread's quad (DFS.metal:1770-1781), branch-free. Returns false when the splat does
// nothing here: an invalid instance (DFS.metal:1750) or all four alphas zero (:1781) -- which covers p > 35, where
// -0.5h * p < -17.5 makes the canonical exp exactly +0.
template <int EXPM>
__device__ __forceinline__ bool evalAlphas(const StagedSplat& sp, const unsigned short* __restrict__ tab, unsigned lx, unsigned ly,
                                           __half2& a0, __half2& a1) {
    const uint2 c = sp.col[lx], r = sp.row[ly];
    const uint4 m0 = sp.m0;
    const uint32_t m1y = sp.m1.w;  // cyy|cyy
    const __half2 t0 = *reinterpret_cast<const __half2*>(&c.x), dx = *reinterpret_cast<const __half2*>(&c.y);
    const __half2 dy2p = *reinterpret_cast<const __half2*>(&r.x), dyp = *reinterpret_cast<const __half2*>(&r.y);
    const __half2 cxy2 = *reinterpret_cast<const __half2*>(&m0.x), cyy = *reinterpret_cast<const __half2*>(&m1y);
    const __half2 op = *reinterpret_cast<const __half2*>(&m0.y);
    const __half2 h099 = h2(0.99f);
    // fma(dx*dy, cxy2, fma(dy*dy, cyy, t0)) for the two rows of the quad
    const __half2 p0 = __hfma2(__hmul2_rn(dx, __low2half2(dyp)), cxy2, __hfma2(__low2half2(dy2p), cyy, t0));
    const __half2 p1 = __hfma2(__hmul2_rn(dx, __high2half2(dyp)), cxy2, __hfma2(__high2half2(dy2p), cyy, t0));
    __half2 e0, e1;
#if GSM_BLEND_TABLE
    e0 = expNegHalfTab(tab, p0); e1 = expNegHalfTab(tab, p1);
#else
    expNegHalfPairs<EXPM>(p0, p1, e0, e1);
#endif
    a0 = __hmin2(__hmul2_rn(op, e0), h099);
    a1 = __hmin2(__hmul2_rn(op, e1), h099);
    // an invalid instance (DFS.metal:1750) is staged as all zeros: p = 0, opacity 0, alphas +0 -- it leaves here like any other
    // splat without effect, no flag to test
    return ((h2bits(a0) | h2bits(a1)) & 0x7FFF7FFFu) != 0u;
}

// ---- persistent form: kBlendGroups tiles in flight per CTA (one 64-thread group each) sharing the CTA's exp table
constexpr int kBlendGroups = GSM_BLEND_GROUPS;
constexpr uint32_t kExpTableEntries = 32768u;                       // the non-negative halfs
constexpr uint32_t kExpTableBytes = kExpTableEntries * 2u;          // 64 KB
constexpr uint32_t kBlendTableSmem = GSM_BLEND_TABLE ? kExpTableBytes : 0u;
constexpr size_t kBlendSmemBytes = kBlendTableSmem + (size_t)kBlendGroups * (kBlendChunk + 1) * sizeof(StagedSplat);

__device__ __forceinline__ void groupBarrier(unsigned group) { asm volatile("bar.sync %0, 64;" ::"r"(group + 1u) : "memory"); }
__device__ __forceinline__ bool groupAll(unsigned group, bool pred) {  // barrier + AND-reduction over the group's 64 threads
    uint32_t r;
    asm volatile("{ .reg .pred p, q; setp.ne.u32 q, %1, 0; barrier.cta.red.and.pred p, %2, 64, q; selp.u32 %0, 1, 0, p; }"
                 : "=r"(r) : "r"(pred ? 1u : 0u), "r"(group + 1u) : "memory");
    return r != 0u;
}

__global__ void __launch_bounds__(256) blend_exp_table_kernel(unsigned short* __restrict__ tab) {
    const uint32_t i = (blockIdx.x * 256u + threadIdx.x) * 2u;  // two inputs per thread: the packed function is evaluated as it is in the blend
    if (i >= kExpTableEntries) return;
    const uint32_t in = i | ((i + 1u) << 16);
    const __half2 r = dhexp2_neghalf_packed(*reinterpret_cast<const __half2*>(&in));
    const uint32_t out = h2bits(r);
    tab[i] = (unsigned short)(out & 0xFFFFu);
    tab[i + 1u] = (unsigned short)(out >> 16);
}

// probe (gsm_probe_math op 13): the blend's table path on arbitrary half2 inputs, table staged in shared memory as in the blend
__global__ void __launch_bounds__(256) blend_exp_probe_kernel(const unsigned short* __restrict__ tab, const unsigned short* __restrict__ in,
                                                              unsigned short* __restrict__ out, uint32_t n) {
    extern __shared__ uint4 s_raw[];
    for (uint32_t i = threadIdx.x; i < kExpTableBytes / 16u; i += 256u) s_raw[i] = __ldg(reinterpret_cast<const uint4*>(tab) + i);
    __syncthreads();
    const unsigned short* s_tab = reinterpret_cast<const unsigned short*>(s_raw);
    for (uint32_t i = (blockIdx.x * 256u + threadIdx.x) * 2u; i < n; i += gridDim.x * 512u) {
        const uint32_t a = in[i], b = (i + 1u < n) ? in[i + 1u] : 0u;
        const uint32_t pk = a | (b << 16);
        const uint32_t r = h2bits(expNegHalfTab(s_tab, *reinterpret_cast<const __half2*>(&pk)));
        out[i] = (unsigned short)(r & 0xFFFFu);
        if (i + 1u < n) out[i + 1u] = (unsigned short)(r >> 16);
    }
}

template <int EXPM>
__global__ void __launch_bounds__(kBlendThreads * kBlendGroups, GSM_BLEND_CTAS) blend_mono_kernel(const uint32_t* __restrict__ lowerBounds,
                                                                   const BlendSplat* __restrict__ splats,
                                                                   const int32_t* __restrict__ instanceIdx, uint32_t width,
                                                                   uint32_t height, uint32_t tilesX, uint32_t tileRowFirst,
                                                                   uint32_t numTiles, const unsigned short* __restrict__ expTable,
                                                                   uint32_t* ticket,
                                                                   __half* __restrict__ color, __half* __restrict__ depth, TileOut tout) {
    extern __shared__ uint4 s_raw[];
    __shared__ uint32_t s_tileOf[kBlendGroups];
    const unsigned short* s_tab = reinterpret_cast<const unsigned short*>(s_raw);
    const unsigned group = threadIdx.x >> 6, tid = threadIdx.x & 63u;
    StagedSplat* s_sp = reinterpret_cast<StagedSplat*>(reinterpret_cast<char*>(s_raw) + kBlendTableSmem) + group * (kBlendChunk + 1);  // + the invalid sentinel that ends an odd chunk
    const unsigned lx = tid & 7u, ly = tid >> 3;
    pdlLaunchDependents();
    // the table is immutable after gsm_renderer_create: staging it does not wait for the predecessor kernel
    for (uint32_t i = threadIdx.x; i < kBlendTableSmem / 16u; i += kBlendThreads * kBlendGroups)
        s_raw[i] = __ldg(reinterpret_cast<const uint4*>(expTable) + i);
    __syncthreads();
    pdlWait();
    const __half thr = __float2half_rn(1.0f / 255.0f);  // half(1.0h/255.0h): both roundings agree (0x1C04)
    const __half2 zero = h2(0.0f), one = h2(1.0f);

    while (true) {
        if (tid == 0) s_tileOf[group] = atomicAdd(ticket, 1u);
        groupBarrier(group);
        const uint32_t t = s_tileOf[group];
        if (t >= numTiles) break;
        const uint32_t tileX = t % tilesX, tileY = tileRowFirst + t / tilesX;
        const uint32_t tile = tileY * tilesX + tileX;
        const uint32_t start = lowerBounds[tile];
        const uint32_t end = lowerBounds[tile + 1];
        const uint32_t count = end > start ? end - start : 0u;

        const uint32_t baseX = tileX * 16u + lx * 2u, baseY = tileY * 16u + ly * 2u;
        // pixel coordinates of the tile as half (quirk Q7): pair k = (16*tile + 2k, +1)
        __half2 pxs[8], pys[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            pxs[k] = __halves2half2(__uint2half_rn(tileX * 16u + 2u * k), __uint2half_rn(tileX * 16u + 2u * k + 1u));
            pys[k] = __halves2half2(__uint2half_rn(tileY * 16u + 2u * k), __uint2half_rn(tileY * 16u + 2u * k + 1u));
        }

        QuadState q;
        q.T0 = one; q.T1 = one;
        q.r0 = q.g0 = q.b0 = q.d0 = q.r1 = q.g1 = q.b1 = q.d1 = zero;
        bool done = false;

#if GSM_BLEND_PREFETCH == 2
        int32_t giCur = tid < count ? __ldg(instanceIdx + start + tid) : -1;   // index one chunk ahead only (one register)
#elif GSM_BLEND_PREFETCH
        // The staging of a chunk is two dependent global loads (instance index -> 32-byte record). They are issued one chunk
        // (records) and two chunks (indices) ahead, so their latency runs under the previous chunk's blending instead of in
        // front of this chunk's barrier.
        int32_t giCur = tid < count ? __ldg(instanceIdx + start + tid) : -1;
        int32_t giNext = kBlendChunk + tid < count ? __ldg(instanceIdx + start + kBlendChunk + tid) : -1;
        uint4 ra = make_uint4(0u, 0u, 0u, 0u), rb = ra;
        if (giCur >= 0) {
            const uint4* src = reinterpret_cast<const uint4*>(splats + giCur);
            ra = __ldg(src); rb = __ldg(src + 1);
        }
#endif
        for (uint32_t base = 0; base < count; base += kBlendChunk) {
            const uint32_t n = min((uint32_t)kBlendChunk, count - base);
            if (tid < n) {
#if GSM_BLEND_PREFETCH
                const int32_t gi = giCur;
#else
                const int32_t gi = __ldg(instanceIdx + start + base + tid);
#endif
                StagedSplat& sp = s_sp[tid];
                if (gi >= 0) {
#if GSM_BLEND_PREFETCH != 1
                    const uint4* src = reinterpret_cast<const uint4*>(splats + gi);
                    const uint4 ra = __ldg(src), rb = __ldg(src + 1);
#endif
                    const __half2 mean = *reinterpret_cast<const __half2*>(&ra.x);
                    const __half2 cxx_cyy = *reinterpret_cast<const __half2*>(&ra.y);
                    const __half2 cxy2_op = *reinterpret_cast<const __half2*>(&ra.z);
                    const __half2 rg = *reinterpret_cast<const __half2*>(&ra.w);
                    const __half2 b_d = *reinterpret_cast<const __half2*>(&rb.x);
                    const __half2 mx = __low2half2(mean), my = __high2half2(mean);
                    const __half2 cxx = __low2half2(cxx_cyy), cyy = __high2half2(cxx_cyy);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const __half2 dx = __hsub2_rn(pxs[k], mx);
                        const __half2 t0 = __hmul2_rn(__hmul2_rn(dx, dx), cxx);
                        sp.col[k] = make_uint2(h2bits(t0), h2bits(dx));
                        const __half2 dy = __hsub2_rn(pys[k], my);
                        sp.row[k] = make_uint2(h2bits(__hmul2_rn(dy, dy)), h2bits(dy));
                    }
                    sp.m0 = make_uint4(h2bits(__low2half2(cxy2_op)), h2bits(__high2half2(cxy2_op)), h2bits(__low2half2(rg)),
                                       h2bits(__high2half2(rg)));
                    sp.m1 = make_uint4(h2bits(__low2half2(b_d)), h2bits(__high2half2(b_d)), 0u, h2bits(cyy));
                } else {
                    zeroStaged(sp);  // "continue" (DFS.metal:1750): opacity 0 makes every alpha +0
                }
            }
            if (tid == 0) zeroStaged(s_sp[n]);  // the loop below reads slot j + 1 unconditionally
            groupBarrier(group);
#if GSM_BLEND_PREFETCH == 2
            giCur = base + kBlendChunk + tid < count ? __ldg(instanceIdx + start + base + kBlendChunk + tid) : -1;
#elif GSM_BLEND_PREFETCH
            giCur = giNext;
            if (giCur >= 0) {
                const uint4* src = reinterpret_cast<const uint4*>(splats + giCur);
                ra = __ldg(src); rb = __ldg(src + 1);
            }
            giNext = base + 2u * kBlendChunk + tid < count ? __ldg(instanceIdx + start + base + 2u * kBlendChunk + tid) : -1;
#endif
            if (!done) {
                // Two splats per trip: the alphas of splat j+1 do not depend on splat j (only the accumulation does), and
                // evaluating both before either is accumulated gives each warp four independent chains instead of two
                // (ncu r1_v11: "wait" -- fixed-latency dependency -- was the largest stall, 2.7 warps per issue slot). The
                // evaluation has no side effect, so doing it for a splat the quad then closes before is invisible. Three and
                // four per trip were measured too and are slower (106 / 118 registers): profiles/README.md, v12.
                for (uint32_t j = 0; j < n; j += 2) {
                    if (quadClosed(q.T0, q.T1, thr)) { done = true; break; }
                    __half2 aA0, aA1, aB0, aB1;
                    const bool useA = evalAlphas<EXPM>(s_sp[j], s_tab, lx, ly, aA0, aA1);
                    const bool useB = evalAlphas<EXPM>(s_sp[j + 1u], s_tab, lx, ly, aB0, aB1);  // slot n is the invalid sentinel
                    GSM_BLEND_STAT(useA);
                    if (useA) {
                        const StagedSplat& sp = s_sp[j];
                        const uint4 m0 = sp.m0, m1 = sp.m1;
                        accumulate(q, aA0, aA1, *reinterpret_cast<const __half2*>(&m0.z), *reinterpret_cast<const __half2*>(&m0.w),
                                   *reinterpret_cast<const __half2*>(&m1.x), *reinterpret_cast<const __half2*>(&m1.y), true);
                    }
                    GSM_BLEND_STAT(useB);
                    if (useB) {
                        // DFS.metal:1746-1747 runs before every splat; T only moved if splat j was used. (When splat j + 1 is
                        // not used the test is simply the one at the top of the next trip.)
                        if (useA && quadClosed(q.T0, q.T1, thr)) { done = true; break; }
                        const StagedSplat& sp = s_sp[j + 1u];
                        const uint4 m0 = sp.m0, m1 = sp.m1;
                        accumulate(q, aB0, aB1, *reinterpret_cast<const __half2*>(&m0.z), *reinterpret_cast<const __half2*>(&m0.w),
                                   *reinterpret_cast<const __half2*>(&m1.x), *reinterpret_cast<const __half2*>(&m1.y), true);
                    }
                }
            }
            if (groupAll(group, done)) break;
        }

        // active tiles: alpha = 1 - T; inactive tiles keep the clear value alpha = 1 (quirk Q6)
        __half2 al0, al1;
        if (count > 0) { al0 = __hsub2_rn(one, q.T0); al1 = __hsub2_rn(one, q.T1); }
        else { al0 = one; al1 = one; }
        storePixelRow(color, depth, width, height, baseX, baseY, q.r0, q.g0, q.b0, al0, q.d0);
        storePixelRow(color, depth, width, height, baseX, baseY + 1u, q.r1, q.g1, q.b1, al1, q.d1);
        if (tid == 0) publishTile(tout, tile, start, count);
        groupBarrier(group);  // the staging buffer and s_tileOf are reused by the next tile
    }
}

// one eye of depthFirstStereoRender for one splat (DFS.metal:1881-1920)
template <int EXPM>
__device__ __forceinline__ void stereoEye(QuadState& q, bool eyeOpen, __half2 mean, __half2 cxx_cyy, __half cxy2h, __half2 op,
                                          __half2 cr, __half2 cg, __half2 cb, __half2 px, __half2 py0, __half2 py1) {
    if (!eyeOpen) return;
    if (!__hge(__low2half(mean), __float2half_rn(-60000.0f))) return;  // invisible eye: mean = -inf
    const __half2 h099 = h2(0.99f), r2Max = h2(9.0f), zero = h2(0.0f);
    const __half2 mx = __low2half2(mean), my = __high2half2(mean);
    const __half2 cxx = __low2half2(cxx_cyy), cyy = __high2half2(cxx_cyy), cxy2 = __half2half2(cxy2h);
    const __half2 dx = __hsub2_rn(px, mx);
    const __half2 p0 = power(dx, __hsub2_rn(py0, my), cxx, cyy, cxy2), p1 = power(dx, __hsub2_rn(py1, my), cxx, cyy, cxy2);
    const __half2 out0 = __hgt2(p0, r2Max), out1 = __hgt2(p1, r2Max);  // 1.0 where p > r2Max (false for NaN)
    const uint32_t o0 = h2bits(out0), o1 = h2bits(out1);
    if (o0 == 0x3C003C00u && o1 == 0x3C003C00u) return;  // all four beyond the cutoff: alphas stay 0
    __half2 e0, e1;
    expNegHalfPairs<EXPM>(p0, p1, e0, e1);
    __half2 a0 = __hmin2(__hmul2_rn(op, e0), h099);
    __half2 a1 = __hmin2(__hmul2_rn(op, e1), h099);
    // per-pixel cutoff: alpha = 0 where p > r2Max
    uint32_t m0 = ((o0 & 0xFFFFu) ? 0u : 0xFFFFu) | ((o0 >> 16) ? 0u : 0xFFFF0000u);
    uint32_t m1 = ((o1 & 0xFFFFu) ? 0u : 0xFFFFu) | ((o1 >> 16) ? 0u : 0xFFFF0000u);
    uint32_t b0 = h2bits(a0) & m0, b1 = h2bits(a1) & m1;
    a0 = *reinterpret_cast<__half2*>(&b0);
    a1 = *reinterpret_cast<__half2*>(&b1);
    if (((b0 | b1) & 0x7FFF7FFFu) == 0u) return;
    accumulate(q, a0, a1, cr, cg, cb, zero, false);
}

// Exact cutoff test for one eye of one splat over a rectangle of the tile (16 columns x `rows` rows), run once by the staging thread. The stereo instance lists hold
// every tile of the union box (DFS.metal:816-825), so for most (splat, tile) pairs every pixel fails p > r2Max and each of
// the 64 threads would find that out on its own. With cxx, cyy >= 0 the rounded evaluation
//   p = RN(m * cxy2 + inner),  m = RN(dx * dy),  inner = RN(RN(dy * dy) * cyy + RN(RN(dx * dx) * cxx))
// is monotone: inner is smallest at the tile's smallest |dx|, |dy| (0 if the tile straddles the mean), m lies between the
// extremes of its four corner products, and RN is monotone, so every pixel's p is >= min(RN(mMin * cxy2 + innerMin),
// RN(mMax * cxy2 + innerMin)). If that bound exceeds r2Max, every alpha is exactly 0 and the eye is skipped. Any NaN in the
// bound (0 * inf, inf - inf) makes the comparison false: no skip, the per-pixel path decides.
__device__ __forceinline__ bool eyeRectBeyondCutoff(__half2 mean, __half2 cxx_cyy, __half cxy2, uint32_t X0, uint32_t Y0, uint32_t rows) {
    const __half zero = __float2half_rn(0.0f), r2Max = __float2half_rn(9.0f);
    const __half mx = __low2half(mean), my = __high2half(mean), cxx = __low2half(cxx_cyy), cyy = __high2half(cxx_cyy);
    if (!(__hge(cxx, zero) && __hge(cyy, zero))) return false;
    const __half xa = __hsub_rn(__uint2half_rn(X0), mx), xb = __hsub_rn(__uint2half_rn(X0 + 15u), mx);
    const __half ya = __hsub_rn(__uint2half_rn(Y0), my), yb = __hsub_rn(__uint2half_rn(Y0 + rows - 1u), my);
    const __half dxm = __hgt(xa, zero) ? xa : (__hlt(xb, zero) ? xb : zero);
    const __half dym = __hgt(ya, zero) ? ya : (__hlt(yb, zero) ? yb : zero);
    const __half inner = __hfma(__hmul_rn(dym, dym), cyy, __hmul_rn(__hmul_rn(dxm, dxm), cxx));
    const __half m1 = __hmul_rn(xa, ya), m2 = __hmul_rn(xa, yb), m3 = __hmul_rn(xb, ya), m4 = __hmul_rn(xb, yb);
    if (__hisnan(m1) || __hisnan(m2) || __hisnan(m3) || __hisnan(m4)) return false;
    const __half mMin = __hmin(__hmin(m1, m2), __hmin(m3, m4)), mMax = __hmax(__hmax(m1, m2), __hmax(m3, m4));
    return __hgt(__hfma(mMin, cxy2, inner), r2Max) && __hgt(__hfma(mMax, cxy2, inner), r2Max);
}

template <int EXPM>
__global__ void __launch_bounds__(kBlendThreads) blend_stereo_kernel(const uint32_t* __restrict__ lowerBounds,
                                                                     const GSMStereoTiledRenderData* __restrict__ splats,
                                                                     const int32_t* __restrict__ instanceIdx, uint32_t width,
                                                                     uint32_t height, uint32_t tilesX,
                                                                     __half* __restrict__ dstSideBySide, int flipY, int eyeMask, TileOut tout) {
    __shared__ uint4 s_rec[kBlendChunk][2];
    __shared__ uint4 s_col[kBlendChunk];   // op|op, r|r, g|g, b|b as half2: the four u8 / 255 divisions, once per splat
    __shared__ uint32_t s_valid[kBlendChunk];
    const bool doL = (eyeMask & 1) != 0, doR = (eyeMask & 2) != 0;  // one-eye-per-GPU split
    const unsigned tid = threadIdx.x;
    const uint32_t tileX = blockIdx.x % tilesX, tileY = blockIdx.x / tilesX;
    const uint32_t tile = tileY * tilesX + tileX;
    pdlLaunchDependents();
    pdlWait();
    const uint32_t start = lowerBounds[tile];
    const uint32_t end = lowerBounds[tile + 1];
    const uint32_t count = end > start ? end - start : 0u;
    const uint32_t baseX = tileX * 16u + (tid & 7u) * 2u, baseY = tileY * 16u + (tid >> 3) * 2u;
    const __half2 px = __halves2half2(__uint2half_rn(baseX), __uint2half_rn(baseX + 1u));
    const __half2 py0 = __half2half2(__uint2half_rn(baseY)), py1 = __half2half2(__uint2half_rn(baseY + 1u));
    const __half thr = __float2half_rn(1.0f / 255.0f);
    const __half2 zero = h2(0.0f), one = h2(1.0f);
    QuadState qL, qR;
    qL.T0 = qL.T1 = qR.T0 = qR.T1 = one;
    qL.r0 = qL.g0 = qL.b0 = qL.d0 = qL.r1 = qL.g1 = qL.b1 = qL.d1 = zero;
    qR = qL;
    bool done = false;
    for (uint32_t base = 0; base < count; base += kBlendChunk) {
        const uint32_t n = min((uint32_t)kBlendChunk, count - base);
        if (tid < n) {
            const int32_t gi = __ldg(instanceIdx + start + base + tid);
            uint32_t flags = gi >= 0 ? 1u : 0u;  // bit 0 valid; bits 1-4: an eye is provably beyond the cutoff on a warp's half of the tile
            if (gi >= 0) {
                const uint4* src = reinterpret_cast<const uint4*>(splats + gi);
                const uint4 ra = __ldg(src);
                s_rec[tid][0] = ra;
                const uint4 rb = __ldg(src + 1);
                s_rec[tid][1] = rb;
                // halfs: ra = {LmeanX,LmeanY | Lcxx,Lcyy | Lcxy2,Ldepth | RmeanX,RmeanY}; rb = {Rcxx,Rcyy | Rcxy2,Rdepth | ...}
                // per eye and per warp (warp w blends rows 8w .. 8w+7 of the tile): bits 1,2 = warp 0 left/right, bits 3,4 = warp 1
                const __half2 mL = *reinterpret_cast<const __half2*>(&ra.x), cL = *reinterpret_cast<const __half2*>(&ra.y);
                const __half xL = __low2half(*reinterpret_cast<const __half2*>(&ra.z));
                const __half2 mR = *reinterpret_cast<const __half2*>(&ra.w), cR = *reinterpret_cast<const __half2*>(&rb.x);
                const __half xR = __low2half(*reinterpret_cast<const __half2*>(&rb.y));
#pragma unroll
                for (uint32_t w = 0; w < 2u; ++w) {
                    if (eyeRectBeyondCutoff(mL, cL, xL, tileX * 16u, tileY * 16u + 8u * w, 8u)) flags |= 2u << (2u * w);
                    if (eyeRectBeyondCutoff(mR, cR, xR, tileX * 16u, tileY * 16u + 8u * w, 8u)) flags |= 4u << (2u * w);
                }
                // every thread of the tile would otherwise redo these IEEE divisions for every splat (703 -> 506 us at C4).
                // Staging the per-column / per-row terms of p as the mono kernel does was tried and is slower here (577 us):
                // the stereo lists hold every AABB tile, most splats leave at the cutoff test, and the staging is not repaid.
                s_col[tid] = make_uint4(h2bits(__half2half2(__float2half_rn((float)(rb.z >> 24) / 255.0f))),
                                        h2bits(__half2half2(__float2half_rn((float)(rb.z & 0xFFu) / 255.0f))),
                                        h2bits(__half2half2(__float2half_rn((float)((rb.z >> 8) & 0xFFu) / 255.0f))),
                                        h2bits(__half2half2(__float2half_rn((float)((rb.z >> 16) & 0xFFu) / 255.0f))));
            }
            s_valid[tid] = flags;
        }
        __syncthreads();
        if (!done) {
            for (uint32_t j = 0; j < n; ++j) {
                // Most entries of a stereo list do nothing on this half of the tile (the lists hold every tile of the union box):
                // their flags are looked at FIRST. The exit test of DFS.metal:1868-1871 runs before every entry in the reference;
                // running it only before the entries that can change a pixel ends the thread before the same accumulations.
                const uint32_t flags = s_valid[j];
                const uint32_t wf = flags >> (2u * (tid >> 5));  // this warp's pair of bits
                if (!(flags & 1u) || (wf & 6u) == 6u) continue;
                const bool closedL = !doL || quadClosed(qL.T0, qL.T1, thr), closedR = !doR || quadClosed(qR.T0, qR.T1, thr);
                if (closedL && closedR) { done = true; break; }
                const bool skipL = closedL || (wf & 2u), skipR = closedR || (wf & 4u);
                if (skipL && skipR) continue;  // nothing this splat can change on this tile
                const uint4 ra = s_rec[j][0], rb = s_rec[j][1];
                // halfs: ra = {LmeanX,LmeanY | Lcxx,Lcyy | Lcxy2,Ldepth | RmeanX,RmeanY}; rb = {Rcxx,Rcyy | Rcxy2,Rdepth | r,g,b,op | cDepth,pad}
                const uint4 sc = s_col[j];
                const __half2 op = *reinterpret_cast<const __half2*>(&sc.x), cr = *reinterpret_cast<const __half2*>(&sc.y);
                const __half2 cg = *reinterpret_cast<const __half2*>(&sc.z), cb = *reinterpret_cast<const __half2*>(&sc.w);
                stereoEye<EXPM>(qL, !skipL, *reinterpret_cast<const __half2*>(&ra.x), *reinterpret_cast<const __half2*>(&ra.y),
                          __low2half(*reinterpret_cast<const __half2*>(&ra.z)), op, cr, cg, cb, px, py0, py1);
                stereoEye<EXPM>(qR, !skipR, *reinterpret_cast<const __half2*>(&ra.w), *reinterpret_cast<const __half2*>(&rb.x),
                          __low2half(*reinterpret_cast<const __half2*>(&rb.y)), op, cr, cg, cb, px, py0, py1);
            }
            // the exit test once more at the end of the chunk, so that the vote below sees quads closed by its last entries
            if (!done && (!doL || quadClosed(qL.T0, qL.T1, thr)) && (!doR || quadClosed(qR.T0, qR.T1, thr))) done = true;
        }
        if (__syncthreads_and(done ? 1 : 0)) break;
    }
    // write both eyes straight into the side-by-side target (left at x in [0,W), right at [W,2W)); the literal
    // stereoCopy maps NDC(-1,-1) to uv(0,0), i.e. dst row y = src row H-1-y (quirk Q9).
    const uint32_t sbsWidth = 2u * width;
#pragma unroll
    for (int row = 0; row < 2; ++row) {
        const uint32_t y = baseY + row;
        if (y >= height) continue;
        const uint32_t dy = flipY ? (height - 1u - y) : y;
        __half2 aL, aR;
        if (count > 0) {
            aL = __hsub2_rn(one, row ? qL.T1 : qL.T0);
            aR = __hsub2_rn(one, row ? qR.T1 : qR.T0);
        } else { aL = one; aR = one; }
        // bounds are per-eye: x < width
        if (baseX < width) {
            const bool two = baseX + 1u < width;
            __half* l = dstSideBySide + 4 * ((size_t)dy * sbsWidth + baseX);
            __half* r = dstSideBySide + 4 * ((size_t)dy * sbsWidth + width + baseX);
            const __half2 Lr = row ? qL.r1 : qL.r0, Lg = row ? qL.g1 : qL.g0, Lb = row ? qL.b1 : qL.b0;
            const __half2 Rr = row ? qR.r1 : qR.r0, Rg = row ? qR.g1 : qR.g0, Rb = row ? qR.b1 : qR.b0;
            // both pixels of the pair as one 16-byte store when the pair is whole and aligned (baseX is even: an even eye width and a
            // 16-byte aligned target are enough) -- what a peer's memory wants (one eye per GPU writes rank 0's target over NVLink)
            const bool wide = two && (width & 1u) == 0u && (reinterpret_cast<uintptr_t>(dstSideBySide) & 15u) == 0u;
            if (wide) {
                uint4 w;
                if (doL) {
                    w.x = h2bits(__halves2half2(__low2half(Lr), __low2half(Lg))); w.y = h2bits(__halves2half2(__low2half(Lb), __low2half(aL)));
                    w.z = h2bits(__halves2half2(__high2half(Lr), __high2half(Lg))); w.w = h2bits(__halves2half2(__high2half(Lb), __high2half(aL)));
                    *reinterpret_cast<uint4*>(l) = w;
                }
                if (doR) {
                    w.x = h2bits(__halves2half2(__low2half(Rr), __low2half(Rg))); w.y = h2bits(__halves2half2(__low2half(Rb), __low2half(aR)));
                    w.z = h2bits(__halves2half2(__high2half(Rr), __high2half(Rg))); w.w = h2bits(__halves2half2(__high2half(Rb), __high2half(aR)));
                    *reinterpret_cast<uint4*>(r) = w;
                }
                continue;
            }
            uint2 v;
            if (doL) {
                v.x = h2bits(__halves2half2(__low2half(Lr), __low2half(Lg))); v.y = h2bits(__halves2half2(__low2half(Lb), __low2half(aL)));
                *reinterpret_cast<uint2*>(l) = v;
            }
            if (doR) {
                v.x = h2bits(__halves2half2(__low2half(Rr), __low2half(Rg))); v.y = h2bits(__halves2half2(__low2half(Rb), __low2half(aR)));
                *reinterpret_cast<uint2*>(r) = v;
            }
            if (two && doL) {
                v.x = h2bits(__halves2half2(__high2half(Lr), __high2half(Lg))); v.y = h2bits(__halves2half2(__high2half(Lb), __high2half(aL)));
                *reinterpret_cast<uint2*>(l + 4) = v;
            }
            if (two && doR) {
                v.x = h2bits(__halves2half2(__high2half(Rr), __high2half(Rg))); v.y = h2bits(__halves2half2(__high2half(Rb), __high2half(aR)));
                *reinterpret_cast<uint2*>(r + 4) = v;
            }
        }
    }
    if (tid == 0) publishTile(tout, tile, start, count);
}

// ---------------------------------------------------------------- GlobalRenderer: globalRender (GlobalShaders.metal:1036-1187)
// One CTA per 32 x 16 tile of the LIMITS, 8 x 8 threads of 4 x 2 pixels; the clear (GlobalShaders.metal:140-154: colour
// (0,0,0,1), depth 0) is fused: tiles without a list write it themselves. Same half arithmetic as the DepthFirst blend
// (power / dhexp2_neghalf_packed / contracted accumulate); the early exit looks at the thread's eight pixels.
struct GlobalStaged {
    __half2 mean;            // meanX, meanY
    __half2 cxx, cyy, cxy2;  // each broadcast to both halves
    __half2 op, r, g, b, d;
    uint32_t valid;
};

template <int EXPM>
__global__ void __launch_bounds__(64) global_render_kernel(GlobalFrame f, uint32_t width, uint32_t height, uint32_t maxWidth,
                                                           uint32_t maxHeight, __half* __restrict__ color, __half* __restrict__ depth) {
    __shared__ GlobalStaged s_sp[64];
    const uint32_t tile = blockIdx.x, tid = threadIdx.x;
    const GSMGaussianHeader hdr = f.tileHeaders[tile];
    const uint32_t tileX = tile % f.tilesX, tileY = tile / f.tilesX;
    const uint32_t baseX = tileX * 32u + (tid & 7u) * 4u, baseY = tileY * 16u + (tid >> 3) * 2u;
    const __half2 one = h2(1.0f), zero = h2(0.0f), h099 = h2(0.99f);
    const __half thr = __float2half_rn(1.0f / 255.0f);
    const __half2 pxA = __halves2half2(__uint2half_rn(baseX), __uint2half_rn(baseX + 1u));
    const __half2 pxB = __halves2half2(__uint2half_rn(baseX + 2u), __uint2half_rn(baseX + 3u));
    const __half2 py0 = __half2half2(__uint2half_rn(baseY)), py1 = __half2half2(__uint2half_rn(baseY + 1u));
    // pair k: 0 = row 0 (x0,x1), 1 = row 0 (x2,x3), 2 = row 1 (x0,x1), 3 = row 1 (x2,x3)
    __half2 T[4] = {one, one, one, one}, R[4] = {zero, zero, zero, zero}, G[4] = {zero, zero, zero, zero},
            B[4] = {zero, zero, zero, zero}, D[4] = {zero, zero, zero, zero};
    bool done = false;
    for (uint32_t c0 = 0; c0 < hdr.count; c0 += 64u) {
        __syncthreads();
        {
            GlobalStaged sp;
            sp.valid = 0u;
            if (c0 + tid < hdr.count) {
                const int32_t gi = f.sortedIndices[hdr.offset + c0 + tid];
                if (gi >= 0) {
                    const uint4 rd = f.renderData[gi];
                    const float theta = (float)(rd.y & 0xFFFFu) * GSM_THETA_UNPACK;
                    float sn, cs;
                    dsincos(theta, sn, cs);
                    const float sig1 = dmax(__half2float(__ushort_as_half((unsigned short)(rd.y >> 16))), 1e-4f);
                    const float sig2 = dmax(__half2float(__ushort_as_half((unsigned short)(rd.z & 0xFFFFu))), 1e-4f);
                    const float invVar1 = 1.0f / (sig1 * sig1), invVar2 = 1.0f / (sig2 * sig2);
                    const float cc = cs * cs, ss = sn * sn, csn = cs * sn;
                    const float A = cc * invVar1 + ss * invVar2, Bq = csn * (invVar1 - invVar2), C = ss * invVar1 + cc * invVar2;
                    sp.mean = *reinterpret_cast<const __half2*>(&rd.x);
                    sp.cxx = __half2half2(__float2half_rn(A));
                    sp.cyy = __half2half2(__float2half_rn(C));
                    sp.cxy2 = __half2half2(__float2half_rn(2.0f * Bq));
                    sp.op = __half2half2(__float2half_rn((float)(rd.w >> 24) / 255.0f));
                    sp.r = __half2half2(__float2half_rn((float)(rd.w & 0xFFu) / 255.0f));
                    sp.g = __half2half2(__float2half_rn((float)((rd.w >> 8) & 0xFFu) / 255.0f));
                    sp.b = __half2half2(__float2half_rn((float)((rd.w >> 16) & 0xFFu) / 255.0f));
                    sp.d = __half2half2(__ushort_as_half((unsigned short)(rd.z >> 16)));
                    sp.valid = 1u;
                }
            }
            s_sp[tid] = sp;
        }
        __syncthreads();
        if (!done) {
            const uint32_t m = min(64u, hdr.count - c0);
            for (uint32_t j = 0; j < m; ++j) {
                {   // GlobalShaders.metal:1081-1084, before every list entry
                    const __half2 mm = __hmax2(__hmax2(T[0], T[1]), __hmax2(T[2], T[3]));
                    if (__hlt(__hmax(__low2half(mm), __high2half(mm)), thr)) { done = true; break; }
                }
                const GlobalStaged& sp = s_sp[j];
                if (!sp.valid) continue;
                const __half2 mx = __low2half2(sp.mean), my = __high2half2(sp.mean);
                const __half2 dxA = __hsub2_rn(pxA, mx), dxB = __hsub2_rn(pxB, mx);
                const __half2 dy0 = __hsub2_rn(py0, my), dy1 = __hsub2_rn(py1, my);
                __half2 a[4];
                __half2 e[4];
                expNegHalfPairs<EXPM>(power(dxA, dy0, sp.cxx, sp.cyy, sp.cxy2), power(dxB, dy0, sp.cxx, sp.cyy, sp.cxy2), e[0], e[1]);
                expNegHalfPairs<EXPM>(power(dxA, dy1, sp.cxx, sp.cyy, sp.cxy2), power(dxB, dy1, sp.cxx, sp.cyy, sp.cxy2), e[2], e[3]);
#pragma unroll
                for (int k = 0; k < 4; ++k) a[k] = __hmin2(__hmul2_rn(sp.op, e[k]), h099);
                if (((h2bits(a[0]) | h2bits(a[1]) | h2bits(a[2]) | h2bits(a[3])) & 0x7FFF7FFFu) == 0u) continue;   // all eight alphas are (+-)0
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const __half2 w = __hmul2_rn(a[k], T[k]);
                    R[k] = __hfma2(sp.r, w, R[k]);
                    G[k] = __hfma2(sp.g, w, G[k]);
                    B[k] = __hfma2(sp.b, w, B[k]);
                    D[k] = __hfma2(sp.d, w, D[k]);
                    T[k] = __hmul2_rn(T[k], __hsub2_rn(one, a[k]));
                }
            }
        }
        if (__syncthreads_and(done ? 1 : 0)) break;
    }
    const uint32_t W = min(width, maxWidth), H = min(height, maxHeight);   // texture writes outside the target are dropped
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = baseX + (uint32_t)(k & 1) * 2u, y = baseY + (uint32_t)(k >> 1);
        if (y >= H) continue;
        const __half2 al = hdr.count > 0u ? __hsub2_rn(one, T[k]) : one;
        const size_t o = (size_t)y * width + x;
        if (x < W) {
            uint2 v;
            v.x = h2bits(__halves2half2(__low2half(R[k]), __low2half(G[k])));
            v.y = h2bits(__halves2half2(__low2half(B[k]), __low2half(al)));
            *reinterpret_cast<uint2*>(color + 4 * o) = v;
            if (depth) depth[o] = __low2half(D[k]);
        }
        if (x + 1u < W) {
            uint2 v;
            v.x = h2bits(__halves2half2(__high2half(R[k]), __high2half(G[k])));
            v.y = h2bits(__halves2half2(__high2half(B[k]), __high2half(al)));
            *reinterpret_cast<uint2*>(color + 4 * o + 4) = v;
            if (depth) depth[o + 1] = __high2half(D[k]);
        }
    }
}

// ---- globalRender in the mono blend's form: persistent CTAs (tile groups of 64 threads on a ticket), the 32-byte record written
// once per Gaussian by the tile count, per-column / per-row terms of p staged once per (tile, splat), staging loads issued one and
// two chunks ahead. A thread owns 4 x 2 pixels = two 2 x 2 quads side by side; its exit test looks at all eight transmittances and
// a splat is skipped when all eight alphas are zero, as in the reference (accumulating a quad whose four alphas are +0 changes
// nothing: w = +0, fma(c, +0, acc) = acc for the non-negative accumulators, T * (1 - 0) = T). C2 cloud: 295 -> see profiles.
struct StagedSplatG {
    uint2 col[16];  // per x pair k of the 32 columns: {T0(2k), T0(2k+1)}, {dx(2k), dx(2k+1)}
    uint2 row[8];   // per y pair k of the 16 rows: {dy*dy(2k), dy*dy(2k+1)}, {dy(2k), dy(2k+1)}
    uint4 m0;       // cxy2|cxy2, op|op, r|r, g|g
    uint4 m1;       // b|b, depth|depth, unused, cyy|cyy
};
constexpr int kGlobalGroups = 4;
constexpr size_t kGlobalSmemBytes = (size_t)kGlobalGroups * (kBlendChunk + 1) * sizeof(StagedSplatG);

template <int EXPM>
__global__ void __launch_bounds__(kBlendThreads * kGlobalGroups, 2) global_render_staged_kernel(GlobalFrame f, uint32_t width, uint32_t height,
                                                                                                 uint32_t maxWidth, uint32_t maxHeight,
                                                                                                 __half* __restrict__ color,
                                                                                                 __half* __restrict__ depth) {
    extern __shared__ uint4 s_raw[];
    __shared__ uint32_t s_tileOf[kGlobalGroups];
    const unsigned group = threadIdx.x >> 6, tid = threadIdx.x & 63u;
    StagedSplatG* s_sp = reinterpret_cast<StagedSplatG*>(s_raw) + group * (kBlendChunk + 1);
    const unsigned lx = tid & 7u, ly = tid >> 3;
    const uint32_t numTiles = f.tilesX * f.tilesY;
    const __half thr = __float2half_rn(1.0f / 255.0f);
    const __half2 zero = h2(0.0f), one = h2(1.0f), h099 = h2(0.99f);
    const uint32_t W = min(width, maxWidth), H = min(height, maxHeight);   // texture writes outside the target are dropped

    while (true) {
        if (tid == 0) s_tileOf[group] = atomicAdd(f.renderTicket, 1u);
        groupBarrier(group);
        const uint32_t tile = s_tileOf[group];
        if (tile >= numTiles) break;
        const uint32_t tileX = tile % f.tilesX, tileY = tile / f.tilesX;
        const GSMGaussianHeader hdr = f.tileHeaders[tile];
        const uint32_t start = hdr.offset, count = hdr.count;
        const uint32_t baseX = tileX * 32u + lx * 4u, baseY = tileY * 16u + ly * 2u;

        QuadState qa, qb;   // columns baseX, baseX + 1 and baseX + 2, baseX + 3
        qa.T0 = qa.T1 = one;
        qa.r0 = qa.g0 = qa.b0 = qa.d0 = qa.r1 = qa.g1 = qa.b1 = qa.d1 = zero;
        qb = qa;
        bool done = false;

        int32_t giCur = tid < count ? __ldg(f.sortedIndices + start + tid) : -1;
        int32_t giNext = kBlendChunk + tid < count ? __ldg(f.sortedIndices + start + kBlendChunk + tid) : -1;
        uint4 ra = make_uint4(0u, 0u, 0u, 0u), rb = ra;
        if (giCur >= 0) {
            const uint4* src = reinterpret_cast<const uint4*>(f.blendSplats + giCur);
            ra = __ldg(src); rb = __ldg(src + 1);
        }
        for (uint32_t base = 0; base < count; base += kBlendChunk) {
            const uint32_t n = min((uint32_t)kBlendChunk, count - base);
            if (tid < n) {
                StagedSplatG& sp = s_sp[tid];
                if (giCur >= 0) {
                    const __half2 mean = *reinterpret_cast<const __half2*>(&ra.x);
                    const __half2 cxx_cyy = *reinterpret_cast<const __half2*>(&ra.y);
                    const __half2 cxy2_op = *reinterpret_cast<const __half2*>(&ra.z);
                    const __half2 rg = *reinterpret_cast<const __half2*>(&ra.w);
                    const __half2 b_d = *reinterpret_cast<const __half2*>(&rb.x);
                    const __half2 mx = __low2half2(mean), my = __high2half2(mean);
                    const __half2 cxx = __low2half2(cxx_cyy), cyy = __high2half2(cxx_cyy);
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const __half2 px = __halves2half2(__uint2half_rn(tileX * 32u + 2u * k), __uint2half_rn(tileX * 32u + 2u * k + 1u));
                        const __half2 dx = __hsub2_rn(px, mx);
                        sp.col[k] = make_uint2(h2bits(__hmul2_rn(__hmul2_rn(dx, dx), cxx)), h2bits(dx));
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const __half2 py = __halves2half2(__uint2half_rn(tileY * 16u + 2u * k), __uint2half_rn(tileY * 16u + 2u * k + 1u));
                        const __half2 dy = __hsub2_rn(py, my);
                        sp.row[k] = make_uint2(h2bits(__hmul2_rn(dy, dy)), h2bits(dy));
                    }
                    sp.m0 = make_uint4(h2bits(__low2half2(cxy2_op)), h2bits(__high2half2(cxy2_op)), h2bits(__low2half2(rg)),
                                       h2bits(__high2half2(rg)));
                    sp.m1 = make_uint4(h2bits(__low2half2(b_d)), h2bits(__high2half2(b_d)), 0u, h2bits(cyy));
                } else {
                    uint4* w = reinterpret_cast<uint4*>(&sp);   // invalid instance: opacity 0 makes every alpha +0
#pragma unroll
                    for (int k = 0; k < (int)(sizeof(StagedSplatG) / 16); ++k) w[k] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            groupBarrier(group);
            giCur = giNext;
            if (giCur >= 0) {
                const uint4* src = reinterpret_cast<const uint4*>(f.blendSplats + giCur);
                ra = __ldg(src); rb = __ldg(src + 1);
            }
            giNext = base + 2u * kBlendChunk + tid < count ? __ldg(f.sortedIndices + start + base + 2u * kBlendChunk + tid) : -1;
            if (!done) {
                for (uint32_t j = 0; j < n; ++j) {
                    // GlobalShaders.metal:1081-1084, before every list entry: all eight pixels of the thread
                    if (quadClosed(qa.T0, qa.T1, thr) && quadClosed(qb.T0, qb.T1, thr)) { done = true; break; }
                    const StagedSplatG& sp = s_sp[j];
                    const uint4 c = *reinterpret_cast<const uint4*>(&sp.col[2u * lx]);   // columns of quad a (x, y), of quad b (z, w)
                    const uint2 r = sp.row[ly];
                    const uint4 m0 = sp.m0, m1 = sp.m1;
                    const __half2 t0a = *reinterpret_cast<const __half2*>(&c.x), dxa = *reinterpret_cast<const __half2*>(&c.y);
                    const __half2 t0b = *reinterpret_cast<const __half2*>(&c.z), dxb = *reinterpret_cast<const __half2*>(&c.w);
                    const __half2 dy2p = *reinterpret_cast<const __half2*>(&r.x), dyp = *reinterpret_cast<const __half2*>(&r.y);
                    const __half2 cxy2 = *reinterpret_cast<const __half2*>(&m0.x), op = *reinterpret_cast<const __half2*>(&m0.y);
                    const __half2 cyy = *reinterpret_cast<const __half2*>(&m1.w);
                    const __half2 dy0 = __low2half2(dyp), dy1 = __high2half2(dyp);
                    const __half2 in0a = __hfma2(__low2half2(dy2p), cyy, t0a), in1a = __hfma2(__high2half2(dy2p), cyy, t0a);
                    const __half2 in0b = __hfma2(__low2half2(dy2p), cyy, t0b), in1b = __hfma2(__high2half2(dy2p), cyy, t0b);
                    const __half2 p0a = __hfma2(__hmul2_rn(dxa, dy0), cxy2, in0a), p1a = __hfma2(__hmul2_rn(dxa, dy1), cxy2, in1a);
                    const __half2 p0b = __hfma2(__hmul2_rn(dxb, dy0), cxy2, in0b), p1b = __hfma2(__hmul2_rn(dxb, dy1), cxy2, in1b);
                    __half2 e0a, e1a, e0b, e1b;
                    expNegHalfPairs<EXPM>(p0a, p1a, e0a, e1a);
                    expNegHalfPairs<EXPM>(p0b, p1b, e0b, e1b);
                    const __half2 a0a = __hmin2(__hmul2_rn(op, e0a), h099), a1a = __hmin2(__hmul2_rn(op, e1a), h099);
                    const __half2 a0b = __hmin2(__hmul2_rn(op, e0b), h099), a1b = __hmin2(__hmul2_rn(op, e1b), h099);
                    const bool useA = ((h2bits(a0a) | h2bits(a1a)) & 0x7FFF7FFFu) != 0u, useB = ((h2bits(a0b) | h2bits(a1b)) & 0x7FFF7FFFu) != 0u;
                    const __half2 cr = *reinterpret_cast<const __half2*>(&m0.z), cg = *reinterpret_cast<const __half2*>(&m0.w);
                    const __half2 cb = *reinterpret_cast<const __half2*>(&m1.x), cd = *reinterpret_cast<const __half2*>(&m1.y);
                    if (useA) accumulate(qa, a0a, a1a, cr, cg, cb, cd, true);
                    if (useB) accumulate(qb, a0b, a1b, cr, cg, cb, cd, true);
                }
            }
            if (groupAll(group, done)) break;
        }

        __half2 al0a, al1a, al0b, al1b;
        if (count > 0) { al0a = __hsub2_rn(one, qa.T0); al1a = __hsub2_rn(one, qa.T1); al0b = __hsub2_rn(one, qb.T0); al1b = __hsub2_rn(one, qb.T1); }
        else { al0a = al1a = al0b = al1b = one; }
        storePixelRow(color, depth, W, H, baseX, baseY, qa.r0, qa.g0, qa.b0, al0a, qa.d0, width);
        storePixelRow(color, depth, W, H, baseX, baseY + 1u, qa.r1, qa.g1, qa.b1, al1a, qa.d1, width);
        storePixelRow(color, depth, W, H, baseX + 2u, baseY, qb.r0, qb.g0, qb.b0, al0b, qb.d0, width);
        storePixelRow(color, depth, W, H, baseX + 2u, baseY + 1u, qb.r1, qb.g1, qb.b1, al1b, qb.d1, width);
        groupBarrier(group);  // the staging buffer and s_tileOf are reused by the next tile
    }
}

// ---- which form of exp(-0.5h * p) the blend kernels of a device run (see expNegHalfPairs): decided once per device by comparing
// the tuned XU-pipe form with the canonical polynomial on all 65 536 half inputs, in both lanes and in both pair slots. A device
// on which any input differs (another MUFU implementation) runs the polynomial; GSM_BLEND_EXP=poly forces it (A/B, tests).
__global__ void __launch_bounds__(256) blend_exp_selftest_kernel(uint32_t* __restrict__ mismatches) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;  // 65 536 threads: input i in the low lane, a different input in the high lane
    const uint32_t a = i | (((i * 40503u + 977u) & 0xFFFFu) << 16), b = (a >> 16) | (a << 16);
    const __half2 pa = *reinterpret_cast<const __half2*>(&a), pb = *reinterpret_cast<const __half2*>(&b);
    __half2 t0, t1;
    dhexp2_neghalf_tuned(pa, pb, t0, t1);
    const __half2 c0 = dhexp2_neghalf_packed(pa), c1 = dhexp2_neghalf_packed(pb);
    // NaN inputs: the tuned form gives NaN, the polynomial +inf; both become 0.99h in min(opacity * e, 0.99h) -- not compared
    const bool nanLo = (i & 0x7FFFu) > 0x7C00u, nanHi = ((a >> 16) & 0x7FFFu) > 0x7C00u;
    uint32_t bad = 0;
    if (!nanLo) bad += (h2bits(t0) & 0xFFFFu) != (h2bits(c0) & 0xFFFFu);
    if (!nanHi) bad += (h2bits(t0) >> 16) != (h2bits(c0) >> 16);
    if (!nanHi) bad += (h2bits(t1) & 0xFFFFu) != (h2bits(c1) & 0xFFFFu);
    if (!nanLo) bad += (h2bits(t1) >> 16) != (h2bits(c1) >> 16);
    if (bad) atomicAdd(mismatches, bad);
}

constexpr int kMaxDevices = 64;
static int g_blendExpMode[kMaxDevices];  // 0 not tested yet (the polynomial runs), 1 polynomial, 2 tuned XU-pipe form
static std::mutex g_blendExpMutex;

cudaError_t blendExpSelfTest(int device) {  // called by gsm_renderer_create with the device current; synchronises (never in a frame)
    if (device < 0 || device >= kMaxDevices) return cudaSuccess;
    std::lock_guard<std::mutex> lock(g_blendExpMutex);
    if (g_blendExpMode[device] != 0) return cudaSuccess;
    const char* env = getenv("GSM_BLEND_EXP");
    if (env && strcmp(env, "poly") == 0) { g_blendExpMode[device] = 1; return cudaSuccess; }
    uint32_t* d = nullptr;
    uint32_t h = 1;
    cudaError_t e = cudaMalloc((void**)&d, sizeof(uint32_t));
    if (e != cudaSuccess) return e;
    e = cudaMemset(d, 0, sizeof(uint32_t));
    if (e == cudaSuccess) { blend_exp_selftest_kernel<<<256, 256>>>(d); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(&h, d, sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    g_blendExpMode[device] = (h == 0u) ? 2 : 1;
    if (h != 0u) fprintf(stderr, "[gsm] device %d: MUFU.EX2 form of the blend's exp differs from the canonical polynomial on %u checks; using the polynomial\n", device, h);
    return cudaSuccess;
}
int blendExpMode(int device) { return (device >= 0 && device < kMaxDevices) ? g_blendExpMode[device] : 0; }
static bool tunedExpOnCurrentDevice() {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess && blendExpMode(dev) == 2;
}

cudaError_t launchGlobalRender(cudaStream_t s, const GlobalFrame& f, uint32_t width, uint32_t height, uint32_t maxWidth, uint32_t maxHeight,
                               __half* color, __half* depth, int numSMs) {
    // GSM_GLOBAL_RENDER=simple keeps the first form (one CTA per tile, conic per (tile, splat)): A/B measurement
    static const bool simple = [] { const char* e = getenv("GSM_GLOBAL_RENDER"); return e && strcmp(e, "simple") == 0; }();
    const bool tuned = tunedExpOnCurrentDevice();
    if (simple) {
        if (tuned) global_render_kernel<1><<<f.tilesX * f.tilesY, 64, 0, s>>>(f, width, height, maxWidth, maxHeight, color, depth);
        else global_render_kernel<0><<<f.tilesX * f.tilesY, 64, 0, s>>>(f, width, height, maxWidth, maxHeight, color, depth);
        return cudaGetLastError();
    }
    static bool attrSet = false;
    if (!attrSet) {
        cudaError_t e = cudaFuncSetAttribute(global_render_staged_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGlobalSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(global_render_staged_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGlobalSmemBytes);
        if (e != cudaSuccess) return e;
        attrSet = true;
    }
    const uint32_t numTiles = f.tilesX * f.tilesY;
    uint32_t grid = (numTiles + kGlobalGroups - 1) / kGlobalGroups;
    if (grid > (uint32_t)numSMs * 2u) grid = (uint32_t)numSMs * 2u;
    if (tuned) global_render_staged_kernel<1><<<grid, kBlendThreads * kGlobalGroups, kGlobalSmemBytes, s>>>(f, width, height, maxWidth, maxHeight, color, depth);
    else global_render_staged_kernel<0><<<grid, kBlendThreads * kGlobalGroups, kGlobalSmemBytes, s>>>(f, width, height, maxWidth, maxHeight, color, depth);
    return cudaGetLastError();
}

cudaError_t launchBlendMono(cudaStream_t s, const uint32_t* lowerBounds, const BlendSplat* splats, const int32_t* instanceIdx,
                            uint32_t width, uint32_t height, uint32_t tilesX, uint32_t tilesY, uint32_t tileRowFirst,
                            uint32_t tileRowCount, __half* color, __half* depth, TileOut tout, const unsigned short* expTable,
                            uint32_t* ticket, int numSMs) {
    (void)tilesY;
    if (tileRowCount == 0) return cudaSuccess;
    static bool attrSet = false;
    if (!attrSet) {
        cudaError_t e = cudaFuncSetAttribute(blend_mono_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlendSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(blend_mono_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlendSmemBytes);
        if (e != cudaSuccess) return e;
        attrSet = true;
    }
    const uint32_t numTiles = tilesX * tileRowCount;
    uint32_t grid = (numTiles + kBlendGroups - 1) / kBlendGroups;
    if (grid > (uint32_t)numSMs * GSM_BLEND_CTAS) grid = (uint32_t)numSMs * GSM_BLEND_CTAS;  // persistent: GSM_BLEND_CTAS CTAs of kBlendGroups tile groups per SM
    return launchChainedSmem(tunedExpOnCurrentDevice() ? blend_mono_kernel<1> : blend_mono_kernel<0>, grid, kBlendThreads * kBlendGroups, s,
                             kBlendSmemBytes, lowerBounds, splats, instanceIdx, width, height, tilesX, tileRowFirst, numTiles, expTable,
                             ticket, color, depth, tout);
}

cudaError_t buildBlendExpTable(cudaStream_t s, unsigned short* table) {
    blend_exp_table_kernel<<<(kExpTableEntries / 2u + 255u) / 256u, 256, 0, s>>>(table);
    return cudaGetLastError();
}
size_t blendExpTableBytes() { return kExpTableBytes; }
bool blendUsesExpTable() { return GSM_BLEND_TABLE != 0; }

cudaError_t launchBlendExpProbe(cudaStream_t s, const unsigned short* table, const unsigned short* in, unsigned short* out, uint32_t n) {
    cudaError_t e = cudaFuncSetAttribute(blend_exp_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExpTableBytes);
    if (e != cudaSuccess) return e;
    blend_exp_probe_kernel<<<64, 256, kExpTableBytes, s>>>(table, in, out, n);
    return cudaGetLastError();
}

cudaError_t launchBlendStereo(cudaStream_t s, const uint32_t* lowerBounds, const GSMStereoTiledRenderData* splats,
                              const int32_t* instanceIdx, uint32_t width, uint32_t height, uint32_t tilesX, uint32_t tilesY,
                              __half* dstSideBySide, int eyeMask, int flipY, TileOut tout) {
    return launchChained(tunedExpOnCurrentDevice() ? blend_stereo_kernel<1> : blend_stereo_kernel<0>, tilesX * tilesY, kBlendThreads, s, lowerBounds, splats, instanceIdx, width, height,
                         tilesX, dstSideBySide, flipY, eyeMask, tout);
}

}  // namespace gsm

#ifdef GSM_BLEND_STATS
extern "C" int gsm_blend_stats_read(unsigned long long* out, int reset) {
    unsigned long long zero[4] = {0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out, gsm::g_blendStats, sizeof zero) != cudaSuccess) return 1;
    if (reset && cudaMemcpyToSymbol(gsm::g_blendStats, zero, sizeof zero) != cudaSuccess) return 1;
    return 0;
}
#endif
