// scene.cu -- scene ingest on the device (include/gsm/gsm_scene.h): the reference's PLYLoader (standard 3DGS layout,
// PLYLoader.swift:517-741, and the PlayCanvas / splat-transform compressed layout, :285-513), SH re-layout to planar
// (:692-719), recentering (:496-503, :722-730), GaussianSceneBuilder.bounds (Scene.swift:159-190), the packing into
// PackedWorldGaussian(+Half) (PLYBenchmarkTests.swift:139-149) and the Morton pre-sort (Scene.swift:47-138).
//
// The header is parsed on the host (text, a few hundred bytes); the body is copied to the device once and decoded there.
// Standard layout: pass A reads the seven properties that decide the placeholder flags and the position bounds; pass B
// decodes, recenters, packs and writes the renderer's buffers directly (records and SH rows staged through shared memory
// so that both the file and the outputs move in coalesced 16-byte accesses) and reduces the scene radius. The compressed
// layout (16 B per vertex) decodes into float records first. Placeholder vertices (PLYLoader.swift:658-660) are rare, so
// their order-preserving compaction index is a kernel that only runs when pass A found any. Arithmetic is the canonical set of gsm_dmath.cuh
// (IEEE + - * / sqrt, Cephes exp), the same definitions the CPU oracle uses, so the outputs are compared bit for bit.
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gsm/gsm_scene.h"
#include "gsm_common.cuh"
#include "gsm_dmath.cuh"
#include "gsm_kernels.h"

namespace gsm {

// ---------------------------------------------------------------- host: header
enum PlyType : int { T_I8, T_U8, T_I16, T_U16, T_I32, T_U32, T_F32, T_F64, T_LIST };
struct PlyProp { std::string name; int type; };
struct PlyElem { std::string name; uint32_t count = 0; std::vector<PlyProp> props; };
struct PlyHeader { int format = -1; std::vector<PlyElem> elems; size_t bodyStart = 0; };

static int plyTypeWidth(int t) {
    switch (t) {
        case T_I8: case T_U8: return 1;
        case T_I16: case T_U16: return 2;
        case T_I32: case T_U32: case T_F32: return 4;
        case T_F64: return 8;
        default: return 0;  // list (PLYLoader.swift:40-45)
    }
}
static int plyTypeFromString(const std::string& s) {  // PLYLoader.swift:191-205
    if (s == "int8" || s == "char") return T_I8;
    if (s == "uint8" || s == "uchar") return T_U8;
    if (s == "int16" || s == "short") return T_I16;
    if (s == "uint16" || s == "ushort") return T_U16;
    if (s == "int32" || s == "int") return T_I32;
    if (s == "uint32" || s == "uint") return T_U32;
    if (s == "float32" || s == "float") return T_F32;
    if (s == "float64" || s == "double") return T_F64;
    return -1;
}
static std::string lowered(std::string s) {
    for (auto& c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}
static bool startsWith(const std::string& s, const char* p) { return s.compare(0, std::strlen(p), p) == 0; }
static bool isShName(const std::string& lname) {
    return startsWith(lname, "f_dc_") || startsWith(lname, "f_rest_") || startsWith(lname, "sh_") ||
           startsWith(lname, "spherical_harmonics_");
}

static size_t findSeq(const unsigned char* d, size_t n, const char* needle) {
    const size_t m = std::strlen(needle);
    for (size_t i = 0; i + m <= n; ++i)
        if (!std::memcmp(d + i, needle, m)) return i;
    return (size_t)-1;
}

// PLYHeader.decodeASCII (PLYLoader.swift:112-187); end_header search as PLYLoader.swift:250-255 (LF first, then CRLF)
static gsm_status parsePlyHeader(const unsigned char* d, size_t n, PlyHeader& h) {
    size_t e = findSeq(d, n, "end_header\n"), len = 11;
    if (e == (size_t)-1) { e = findSeq(d, n, "end_header\r\n"); len = 12; }
    if (e == (size_t)-1) return reportFailure((gsm_status)GSM_ERR_PLY_INVALID_HEADER, "PLY: end_header not found");
    h.bodyStart = e + len;
    size_t pos = 0;
    while (pos < h.bodyStart) {
        size_t eol = pos;
        while (eol < h.bodyStart && d[eol] != '\n' && d[eol] != '\r') eol++;
        std::string line((const char*)d + pos, eol - pos);
        pos = eol;
        while (pos < h.bodyStart && (d[pos] == '\n' || d[pos] == '\r')) pos++;
        std::vector<std::string> tok;
        size_t i = 0;
        while (i < line.size()) {
            while (i < line.size() && (line[i] == ' ' || line[i] == '\t')) i++;
            size_t j = i;
            while (j < line.size() && line[j] != ' ' && line[j] != '\t') j++;
            if (j > i) tok.push_back(line.substr(i, j - i));
            i = j;
        }
        if (tok.empty()) continue;
        const std::string& kw = tok[0];
        if (kw == "ply" || kw == "comment" || kw == "obj_info") continue;
        if (kw == "end_header") break;
        auto bad = [&](const char* what) { return reportFailure((gsm_status)GSM_ERR_PLY_INVALID_HEADER, (std::string("PLY header: ") + what + ": \"" + line + "\"").c_str()); };
        if (kw == "format") {
            if (h.format >= 0) return bad("unexpected keyword");
            if (tok.size() < 3) return bad("invalid line");
            if (tok[1] == "ascii") h.format = 0;
            else if (tok[1] == "binary_little_endian") h.format = 1;
            else if (tok[1] == "binary_big_endian") h.format = 2;
            else return bad("invalid format type");
        } else if (kw == "element") {
            if (h.format < 0) return bad("unexpected keyword");
            if (tok.size() < 3) return bad("invalid line");
            PlyElem el;
            el.name = tok[1];
            el.count = (uint32_t)std::strtoul(tok[2].c_str(), nullptr, 10);
            h.elems.push_back(el);
        } else if (kw == "property") {
            if (h.format < 0 || h.elems.empty()) return bad("unexpected keyword");
            PlyProp p;
            if (tok.size() >= 5 && tok[1] == "list") {
                if (plyTypeFromString(tok[2]) < 0 || plyTypeFromString(tok[3]) < 0) return bad("unknown property type");
                p.type = T_LIST;
                p.name = tok[4];
            } else if (tok.size() >= 3) {
                p.type = plyTypeFromString(tok[1]);
                if (p.type < 0) return bad("unknown property type");
                p.name = tok[2];
            } else {
                return bad("invalid line");
            }
            h.elems.back().props.push_back(p);
        } else {
            return bad("unknown keyword");
        }
    }
    if (h.format < 0) return reportFailure((gsm_status)GSM_ERR_PLY_INVALID_HEADER, "PLY header: format missing");
    return GSM_OK;
}
static const PlyElem* findElem(const PlyHeader& h, const char* name) {
    for (auto& e : h.elems)
        if (e.name == name) return &e;
    return nullptr;
}
static bool hasProp(const PlyElem& e, const char* name) {
    for (auto& p : e.props)
        if (p.name == name) return true;
    return false;
}
static bool isCompressed(const PlyHeader& h, const PlyElem& vx) {  // PLYLoader.swift:269-274
    return findElem(h, "chunk") && hasProp(vx, "packed_position") && hasProp(vx, "packed_rotation") &&
           hasProp(vx, "packed_scale") && hasProp(vx, "packed_color");
}

// ---------------------------------------------------------------- device: decode
constexpr int kMaxShProps = 64;
struct StandardLayout {
    uint32_t stride, count;
    int32_t off[11];   // x y z s0 s1 s2 r0 r1 r2 r3 opacity: byte offset in the vertex record, -1 = absent
    int32_t type[11];
    int32_t shOff[kMaxShProps], shType[kMaxShProps];
    uint32_t shProps, shComponents;
    uint32_t scaleIsLogSpace, opacityIsLogit;
};
struct SceneScratch {      // device, zero-initialised (bounds keys start at their identity, see initScratch)
    uint32_t minKey[3], maxKey[3];  // order-preserving integer images of the position bounds
    uint32_t radiusKey;             // max over records of |p - c| + max scale (non-negative floats order as integers)
    uint32_t placeholders;
    uint32_t kept;
    uint32_t ticket;
};

__device__ __forceinline__ uint32_t floatOrderKey(float f) {  // monotonic: -0 below +0, like the min / max of gsm_dmath.cuh
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float floatFromOrderKey(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// getFloat (PLYLoader.swift:598-618) on a possibly unaligned record
__device__ __forceinline__ float readPlyProp(const unsigned char* p, int type) {
    switch (type) {
        case T_F32: {
            uint32_t u;
            if (((uintptr_t)p & 3u) == 0u) u = *reinterpret_cast<const uint32_t*>(p);
            else u = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
            return __uint_as_float(u);
        }
        case T_F64: {
            unsigned long long u = 0;
            for (int i = 0; i < 8; ++i) u |= (unsigned long long)p[i] << (8 * i);
            return (float)__longlong_as_double((long long)u);
        }
        case T_U8: return (float)p[0] / 255.0f;
        case T_I8: return (float)(signed char)p[0];
        case T_I16: return (float)(short)((uint32_t)p[0] | ((uint32_t)p[1] << 8));
        case T_U16: return (float)((uint32_t)p[0] | ((uint32_t)p[1] << 8));
        case T_I32: return (float)(int)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
        case T_U32: return (float)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
        default: return 0.0f;
    }
}

__device__ __forceinline__ void reduceBounds(SceneScratch* sc, bool valid, float x, float y, float z) {
    // warp min / max of the order keys, one atomic per warp and axis; NaN positions lose (fmin / fmax semantics)
    const float v[3] = {x, y, z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const bool ok = valid && v[k] == v[k];
        uint32_t lo = ok ? floatOrderKey(v[k]) : 0xFFFFFFFFu, hi = ok ? floatOrderKey(v[k]) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
        }
        if ((threadIdx.x & 31u) == 0u) {
            // same-address atomics cost ~0.7 ns each (tools/micro/atomic_rate.cu): 6 per warp would dominate the kernel, and
            // after the first few warps almost none of them changes the bounds -- look before updating
            if (lo < *(volatile uint32_t*)&sc->minKey[k]) atomicMin(&sc->minKey[k], lo);
            if (hi > *(volatile uint32_t*)&sc->maxKey[k]) atomicMax(&sc->maxKey[k], hi);
        }
    }
}

// Standard layout, pass A: keep flags (placeholders, PLYLoader.swift:658-660) and the position bounds. Reads only the
// seven properties it needs -- a fraction of each record's sectors.
__global__ void __launch_bounds__(256) ply_scan_standard_kernel(const unsigned char* __restrict__ body, const __grid_constant__ StandardLayout L,
                                                                uint8_t* __restrict__ keep, SceneScratch* sc) {
    const uint32_t v = blockIdx.x * 256u + threadIdx.x;
    bool kept = false;
    float px = 0, py = 0, pz = 0;
    if (v < L.count) {
        const unsigned char* rec = body + (size_t)v * L.stride;
        auto get = [&](int i) { return L.off[i] >= 0 ? readPlyProp(rec + L.off[i], L.type[i]) : 0.0f; };
        const float s0 = get(3), s1 = get(4), s2 = get(5), opRaw = get(10);
        kept = !(s0 == 2.0f && s1 == 2.0f && s2 == 2.0f && fabsf(opRaw - 4.8402f) < 0.001f);
        keep[v] = kept ? 1 : 0;
        if (kept) { px = get(0); py = get(1); pz = get(2); }
    }
    const unsigned lost = __ballot_sync(0xFFFFFFFFu, v < L.count && !kept);
    if ((threadIdx.x & 31u) == 0u && lost) atomicAdd(&sc->placeholders, (uint32_t)__popc(lost));
    reduceBounds(sc, kept, px, py, pz);
}

// the recentering shift and the center of the recentered bounds, from the reduced keys; every thread (and the host)
// derives them with the same float arithmetic
struct SceneCenters { float c[3], c2[3]; bool shift; };
__host__ __device__ inline float sceneKeyToFloat(uint32_t k) {
    const uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    float f;
    memcpy(&f, &b, 4);
    return f;
}
__host__ __device__ inline SceneCenters sceneCenters(const uint32_t minKey[3], const uint32_t maxKey[3]) {
    SceneCenters r;
    float mn[3], mx[3];
    const bool any = minKey[0] != 0xFFFFFFFFu;
    for (int k = 0; k < 3; ++k) {
        mn[k] = any ? sceneKeyToFloat(minKey[k]) : 0.0f;
        mx[k] = any ? sceneKeyToFloat(maxKey[k]) : 0.0f;
        r.c[k] = (mn[k] + mx[k]) * 0.5f;
    }
    r.shift = sqrtf((r.c[0] * r.c[0] + r.c[1] * r.c[1]) + r.c[2] * r.c[2]) > 1e-6f;  // PLYLoader.swift:725
    for (int k = 0; k < 3; ++k) {
        if (!r.shift) r.c[k] = 0.0f;
        // bounds of the recentered records: subtracting the same c is monotonic, so min' = min - c and max' = max - c exactly
        const float mn2 = r.shift ? mn[k] - r.c[k] : mn[k], mx2 = r.shift ? mx[k] - r.c[k] : mx[k];
        r.c2[k] = (mn2 + mx2) * 0.5f;
    }
    return r;
}

// Standard layout, pass B: decode, recenter, pack and write -- no intermediate arrays. A CTA stages its 128 consecutive
// records in shared memory with coalesced 16-byte loads (a thread-per-record read of 248-byte records touches 32 lines
// per instruction and re-fetched the file twice from DRAM, ncu r1_scene) and stages its planar SH rows the same way on the
// way out. Also reduces the scene radius (Scene.swift:179-188).
constexpr uint32_t kPlyBlock = 128;
template <bool HALF>
__global__ void __launch_bounds__(kPlyBlock) ply_decode_pack_standard_kernel(const unsigned char* __restrict__ body, const __grid_constant__ StandardLayout L,
                                                                             const uint8_t* __restrict__ keep, const uint32_t* __restrict__ dstIndex,
                                                                             void* __restrict__ gaussiansOut, void* __restrict__ harmonicsOut,
                                                                             SceneScratch* sc, uint32_t stageRecords) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    const uint32_t first = blockIdx.x * kPlyBlock, v = first + threadIdx.x;
    const uint32_t nHere = min(kPlyBlock, L.count - first);
    const unsigned char* rec = body + (size_t)v * L.stride;
    if (stageRecords) {  // the body buffer is 16-byte aligned and a block starts at 128 * stride bytes: always a multiple of 16
        const size_t startByte = (size_t)first * L.stride, bytes = (size_t)nHere * L.stride;
        const size_t vecs = bytes / 16u;
        const uint4* src = reinterpret_cast<const uint4*>(body + startByte);
        for (size_t i = threadIdx.x; i < vecs; i += kPlyBlock) reinterpret_cast<uint4*>(s_dyn)[i] = __ldcs(src + i);
        for (size_t i = vecs * 16u + threadIdx.x; i < bytes; i += kPlyBlock) s_dyn[i] = body[startByte + i];
        __syncthreads();
        rec = s_dyn + (size_t)threadIdx.x * L.stride;
    }
    const SceneCenters cen = sceneCenters(sc->minKey, sc->maxKey);
    float r = 0.0f;
    const bool kept = v < L.count && keep[v];
    const uint32_t shStride = L.shComponents > 0 ? L.shProps : 0u;
    float* shStage = reinterpret_cast<float*>(s_dyn + (((size_t)(stageRecords ? kPlyBlock : 0) * L.stride + 15u) & ~(size_t)15u));
    const bool stageSh = dstIndex == nullptr && shStride > 0;  // contiguous destination rows: stage and write coalesced
    if (kept) {
        auto get = [&](int i) { return L.off[i] >= 0 ? readPlyProp(rec + L.off[i], L.type[i]) : 0.0f; };
        const uint32_t d = dstIndex ? dstIndex[v] : v;
        float p[3] = {get(0), get(1), get(2)};
#pragma unroll
        for (int k = 0; k < 3; ++k) p[k] = cen.shift ? p[k] - cen.c[k] : p[k];
        const float s0 = get(3), s1 = get(4), s2 = get(5), opRaw = get(10);
        const float sx = L.scaleIsLogSpace ? dexp(s0) : s0, sy = L.scaleIsLogSpace ? dexp(s1) : s1, sz = L.scaleIsLogSpace ? dexp(s2) : s2;
        float qx = get(7), qy = get(8), qz = get(9), qw = get(6);  // simd_quatf(ix: rot_1, iy: rot_2, iz: rot_3, r: rot_0)
        const float inv = 1.0f / sqrtf(((qx * qx + qy * qy) + qz * qz) + qw * qw);
        qx *= inv; qy *= inv; qz *= inv; qw *= inv;
        const float op = L.opacityIsLogit ? 1.0f / (1.0f + dexp(-opRaw)) : opRaw;
        if (HALF) {
            GSMPackedWorldGaussianHalf g;
            g.px = p[0]; g.py = p[1]; g.pz = p[2];
            g.opacity = __half_as_ushort(__float2half_rn(op));
            g.sx = __half_as_ushort(__float2half_rn(sx)); g.sy = __half_as_ushort(__float2half_rn(sy)); g.sz = __half_as_ushort(__float2half_rn(sz));
            g.rx = __half_as_ushort(__float2half_rn(qx)); g.ry = __half_as_ushort(__float2half_rn(qy));
            g.rz = __half_as_ushort(__float2half_rn(qz)); g.rw = __half_as_ushort(__float2half_rn(qw));
            g._pad0 = 0; g._pad1 = 0;
            uint4* o = reinterpret_cast<uint4*>(gaussiansOut) + 2 * (size_t)d;
            o[0] = reinterpret_cast<const uint4*>(&g)[0];
            o[1] = reinterpret_cast<const uint4*>(&g)[1];
        } else {
            float4* o = reinterpret_cast<float4*>(gaussiansOut) + 3 * (size_t)d;
            o[0] = make_float4(p[0], p[1], p[2], op);
            o[1] = make_float4(sx, sy, sz, 0.0f);
            o[2] = make_float4(qx, qy, qz, qw);
        }
        if (shStride > 0) {  // PLY [DC_R, DC_G, DC_B, R1.., G1.., B1..] -> shader [R0.., G0.., B0..] (PLYLoader.swift:700-719)
            const uint32_t K = L.shComponents, hoc = K - 1u;
            for (uint32_t k = 0; k < shStride; ++k) {
                // destination slot k <- source property: channel ch = k / K, coefficient c = k % K
                float val = 0.0f;
                if (k < 3u * K) {
                    const uint32_t ch = k / K, c = k - ch * K;
                    const uint32_t src = c == 0u ? ch : 3u + ch * hoc + (c - 1u);
                    val = readPlyProp(rec + L.shOff[src], L.shType[src]);
                }
                if (stageSh) shStage[(size_t)threadIdx.x * shStride + k] = val;
                else if (HALF) reinterpret_cast<__half*>(harmonicsOut)[(size_t)d * shStride + k] = __float2half_rn(val);
                else reinterpret_cast<float*>(harmonicsOut)[(size_t)d * shStride + k] = val;
            }
        }
        const float ox = p[0] - cen.c2[0], oy = p[1] - cen.c2[1], oz = p[2] - cen.c2[2];
        r = sqrtf((ox * ox + oy * oy) + oz * oz) + dmax(sx, dmax(sy, sz));
        if (!(r == r)) r = 0.0f;  // a NaN candidate loses (max semantics of gsm_dmath.cuh)
    }
    if (stageSh) {
        __syncthreads();
        const size_t total = (size_t)nHere * shStride, base = (size_t)first * shStride;
        for (size_t i = threadIdx.x; i < total; i += kPlyBlock) {
            if (HALF) reinterpret_cast<__half*>(harmonicsOut)[base + i] = __float2half_rn(shStage[i]);
            else reinterpret_cast<float*>(harmonicsOut)[base + i] = shStage[i];
        }
    }
    uint32_t rk = __float_as_uint(dmax(r, 0.0f));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rk = max(rk, __shfl_xor_sync(0xFFFFFFFFu, rk, o));
    if ((threadIdx.x & 31u) == 0u && rk > *(volatile uint32_t*)&sc->radiusKey) atomicMax(&sc->radiusKey, rk);
}

struct CompressedLayout {
    uint32_t chunkStride, vertexStride, count;
    uint64_t vertexStart;  // byte offset of the vertex element from the chunk element's start
    int32_t chunkOff[18];  // min_x min_y min_z max_x max_y max_z min_scale_x.. max_scale_z min_r min_g min_b max_r max_g max_b
    int32_t vertexOff[4];  // packed_position packed_rotation packed_scale packed_color
};
__device__ __forceinline__ float unpackUnorm(uint32_t v, int bits) {  // PLYLoader.swift:354-357
    const uint32_t mask = (1u << bits) - 1u;
    return (float)(v & mask) / (float)mask;
}
__device__ __forceinline__ float lerpPly(float a, float b, float t) { return a * (1.0f - t) + b * t; }  // PLYLoader.swift:401-403

__global__ void __launch_bounds__(256) ply_decode_compressed_kernel(const unsigned char* __restrict__ body, const __grid_constant__ CompressedLayout L,
                                                                    float* __restrict__ pos, float* __restrict__ scale,
                                                                    float* __restrict__ rot, float* __restrict__ opacity,
                                                                    float* __restrict__ sh, uint8_t* __restrict__ keep, SceneScratch* sc) {
    const uint32_t v = blockIdx.x * 256u + threadIdx.x;
    float px = 0, py = 0, pz = 0;
    const bool valid = v < L.count;
    if (valid) {
        const unsigned char* ch = body + (size_t)(v / 256u) * L.chunkStride;
        const unsigned char* vx = body + L.vertexStart + (size_t)v * L.vertexStride;
        auto cf = [&](int i) { return L.chunkOff[i] >= 0 ? readPlyProp(ch + L.chunkOff[i], T_F32) : 0.0f; };
        auto vu = [&](int i) { return L.vertexOff[i] >= 0 ? __float_as_uint(readPlyProp(vx + L.vertexOff[i], T_F32)) : 0u; };
        const uint32_t pp = vu(0), pr = vu(1), ps = vu(2), pc = vu(3);
        px = lerpPly(cf(0), cf(3), unpackUnorm(pp >> 21, 11));
        py = lerpPly(cf(1), cf(4), unpackUnorm(pp >> 11, 10));
        pz = lerpPly(cf(2), cf(5), unpackUnorm(pp, 11));
        pos[3 * (size_t)v] = px; pos[3 * (size_t)v + 1] = py; pos[3 * (size_t)v + 2] = pz;
        const float norm = 1.0f / (sqrtf(2.0f) * 0.5f);  // PLYLoader.swift:375
        const float a = (unpackUnorm(pr >> 20, 10) - 0.5f) * norm, b = (unpackUnorm(pr >> 10, 10) - 0.5f) * norm,
                    c = (unpackUnorm(pr, 10) - 0.5f) * norm;
        const float m = sqrtf(dmax(0.0f, 1.0f - ((a * a + b * b) + c * c)));
        float qx, qy, qz, qw;  // PLYLoader.swift:392-398
        switch (pr >> 30) {
            case 0: qx = a; qy = b; qz = c; qw = m; break;
            case 1: qx = m; qy = b; qz = c; qw = a; break;
            case 2: qx = b; qy = m; qz = c; qw = a; break;
            default: qx = b; qy = c; qz = m; qw = a; break;
        }
        rot[4 * (size_t)v] = qx; rot[4 * (size_t)v + 1] = qy; rot[4 * (size_t)v + 2] = qz; rot[4 * (size_t)v + 3] = qw;
        scale[3 * (size_t)v] = dexp(lerpPly(cf(6), cf(9), unpackUnorm(ps >> 21, 11)));
        scale[3 * (size_t)v + 1] = dexp(lerpPly(cf(7), cf(10), unpackUnorm(ps >> 11, 10)));
        scale[3 * (size_t)v + 2] = dexp(lerpPly(cf(8), cf(11), unpackUnorm(ps, 11)));
        const float cr = unpackUnorm(pc >> 24, 8), cg = unpackUnorm(pc >> 16, 8), cb = unpackUnorm(pc >> 8, 8);
        opacity[v] = unpackUnorm(pc, 8);
        const float SH_C0 = 0.28209479177387814f;
        sh[3 * (size_t)v] = (lerpPly(cf(12), cf(15), cr) - 0.5f) / SH_C0;
        sh[3 * (size_t)v + 1] = (lerpPly(cf(13), cf(16), cg) - 0.5f) / SH_C0;
        sh[3 * (size_t)v + 2] = (lerpPly(cf(14), cf(17), cb) - 0.5f) / SH_C0;
        keep[v] = 1;
    }
    reduceBounds(sc, valid, px, py, pz);
}

// order-preserving compaction index of the kept vertices (only launched when placeholders exist): 256 vertices per tile,
// eager prefix over tiles (gsm_common.cuh), persistent CTAs on a ticket
__global__ void __launch_bounds__(256) ply_keep_index_kernel(const uint8_t* __restrict__ keep, uint32_t n, uint32_t* __restrict__ dstIndex,
                                                             unsigned long long* tileWords, unsigned long long* groupWords, SceneScratch* sc) {
    __shared__ uint32_t s_scan[9];
    __shared__ uint32_t s_tile, s_base;
    const uint32_t numTiles = (n + 255u) / 256u;
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&sc->ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= numTiles) break;
        const uint32_t v = tile * 256u + threadIdx.x;
        const uint32_t k = (v < n && keep[v]) ? 1u : 0u;
        uint32_t total;
        const uint32_t excl = block_exclusive_scan_256(k, s_scan, total);
        if (threadIdx.x == 0) prefixPublish(tileWords, groupWords, tile, total);
        if (threadIdx.x < 32) {
            const uint32_t b = prefixResolve(tileWords, groupWords, tile);
            if (threadIdx.x == 0) { s_base = b; if (tile == numTiles - 1u) sc->kept = b + total; }
        }
        __syncthreads();
        if (v < n) dstIndex[v] = k ? s_base + excl : 0xFFFFFFFFu;
        __syncthreads();
    }
}

// recenter (PLYLoader.swift:722-730), pack (PLYBenchmarkTests.swift:139-149) and reduce the scene radius (Scene.swift:179-188)
template <bool HALF>
__global__ void __launch_bounds__(256) scene_pack_kernel(const float* __restrict__ pos, const float* __restrict__ scale,
                                                         const float* __restrict__ rot, const float* __restrict__ opacity,
                                                         const float* __restrict__ sh, const uint8_t* __restrict__ keep,
                                                         const uint32_t* __restrict__ dstIndex, uint32_t n, uint32_t shStride,
                                                         void* __restrict__ gaussiansOut, void* __restrict__ harmonicsOut, SceneScratch* sc) {
    const uint32_t v = blockIdx.x * 256u + threadIdx.x;
    // the bounds are complete (previous kernel); every thread derives the same center with the same arithmetic
    const SceneCenters cen = sceneCenters(sc->minKey, sc->maxKey);
    const bool shift = cen.shift;
    const float* c = cen.c;
    const float* c2 = cen.c2;
    float r = 0.0f;
    if (v < n && keep[v]) {
        const uint32_t d = dstIndex ? dstIndex[v] : v;
        float p[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) p[k] = shift ? pos[3 * (size_t)v + k] - c[k] : pos[3 * (size_t)v + k];
        const float sx = scale[3 * (size_t)v], sy = scale[3 * (size_t)v + 1], sz = scale[3 * (size_t)v + 2];
        const float qx = rot[4 * (size_t)v], qy = rot[4 * (size_t)v + 1], qz = rot[4 * (size_t)v + 2], qw = rot[4 * (size_t)v + 3];
        const float op = opacity[v];
        if (HALF) {
            GSMPackedWorldGaussianHalf g;
            g.px = p[0]; g.py = p[1]; g.pz = p[2];
            g.opacity = __half_as_ushort(__float2half_rn(op));
            g.sx = __half_as_ushort(__float2half_rn(sx)); g.sy = __half_as_ushort(__float2half_rn(sy)); g.sz = __half_as_ushort(__float2half_rn(sz));
            g.rx = __half_as_ushort(__float2half_rn(qx)); g.ry = __half_as_ushort(__float2half_rn(qy));
            g.rz = __half_as_ushort(__float2half_rn(qz)); g.rw = __half_as_ushort(__float2half_rn(qw));
            g._pad0 = 0; g._pad1 = 0;
            uint4* o = reinterpret_cast<uint4*>(gaussiansOut) + 2 * (size_t)d;
            o[0] = reinterpret_cast<const uint4*>(&g)[0];
            o[1] = reinterpret_cast<const uint4*>(&g)[1];
            __half* ho = reinterpret_cast<__half*>(harmonicsOut) + (size_t)d * shStride;
            for (uint32_t k = 0; k < shStride; ++k) ho[k] = __float2half_rn(sh[(size_t)v * shStride + k]);
        } else {
            float4* o = reinterpret_cast<float4*>(gaussiansOut) + 3 * (size_t)d;
            o[0] = make_float4(p[0], p[1], p[2], op);
            o[1] = make_float4(sx, sy, sz, 0.0f);
            o[2] = make_float4(qx, qy, qz, qw);
            float* ho = reinterpret_cast<float*>(harmonicsOut) + (size_t)d * shStride;
            for (uint32_t k = 0; k < shStride; ++k) ho[k] = sh[(size_t)v * shStride + k];
        }
        const float ox = p[0] - c2[0], oy = p[1] - c2[1], oz = p[2] - c2[2];
        r = sqrtf((ox * ox + oy * oy) + oz * oz) + dmax(sx, dmax(sy, sz));
        if (!(r == r)) r = 0.0f;  // a NaN candidate loses (max semantics of gsm_dmath.cuh)
    }
    uint32_t rk = __float_as_uint(dmax(r, 0.0f));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rk = max(rk, __shfl_xor_sync(0xFFFFFFFFu, rk, o));
    if ((threadIdx.x & 31u) == 0u && rk > *(volatile uint32_t*)&sc->radiusKey) atomicMax(&sc->radiusKey, rk);
}

// ---------------------------------------------------------------- device: Morton pre-sort
__global__ void __launch_bounds__(256) scene_position_bounds_kernel(const unsigned char* __restrict__ gaussians, uint32_t strideBytes, uint32_t n,
                                                                    SceneScratch* sc) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    float x = 0, y = 0, z = 0;
    if (i < n) {
        const float* p = reinterpret_cast<const float*>(gaussians + (size_t)i * strideBytes);
        x = p[0]; y = p[1]; z = p[2];
    }
    reduceBounds(sc, i < n, x, y, z);
}
__device__ __forceinline__ unsigned long long expandBits21(unsigned long long v) {  // Scene.swift:50-58
    unsigned long long x = v & 0x1FFFFFull;
    x = (x | (x << 32)) & 0x1F00000000FFFFull;
    x = (x | (x << 16)) & 0x1F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
__global__ void __launch_bounds__(256) scene_morton_codes_kernel(const unsigned char* __restrict__ gaussians, uint32_t strideBytes, uint32_t n,
                                                                 const SceneScratch* sc, uint32_t* __restrict__ lo, uint32_t* __restrict__ hi,
                                                                 uint32_t* __restrict__ index) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(gaussians + (size_t)i * strideBytes);
    unsigned long long code = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float mn = floatFromOrderKey(sc->minKey[k]), mx = floatFromOrderKey(sc->maxKey[k]);
        const float e = mx - mn, inv = e > 1e-6f ? 1.0f / e : 0.0f;  // Scene.swift:97-102
        const float s = 2097151.0f;
        const float q = dmax(0.0f, dmin(s, ((p[k] - mn) * inv) * s));   // Scene.swift:64-67; UInt64(Float) truncates
        code |= expandBits21((unsigned long long)q) << k;
    }
    lo[i] = (uint32_t)code;
    hi[i] = (uint32_t)(code >> 32);
    index[i] = i;
}
__global__ void __launch_bounds__(256) gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ index, uint32_t n,
                                                         uint32_t* __restrict__ dst) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i < n) dst[i] = src[index[i]];
}
// out[i] = in[order[i]] for records of `words` 32-bit words
__global__ void __launch_bounds__(256) permute_records_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ order, uint32_t n,
                                                              uint32_t words, uint32_t* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * 256u + threadIdx.x;
    const size_t total = (size_t)n * words;
    if (t >= total) return;
    const uint32_t i = (uint32_t)(t / words), w = (uint32_t)(t - (size_t)i * words);
    out[t] = in[(size_t)order[i] * words + w];
}
__global__ void __launch_bounds__(256) permute_halfs_kernel(const uint16_t* __restrict__ in, const uint32_t* __restrict__ order, uint32_t n,
                                                            uint32_t elems, uint16_t* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * 256u + threadIdx.x;
    const size_t total = (size_t)n * elems;
    if (t >= total) return;
    const uint32_t i = (uint32_t)(t / elems), w = (uint32_t)(t - (size_t)i * elems);
    out[t] = in[(size_t)order[i] * elems + w];
}

static void initScratch(SceneScratch& h) {
    std::memset(&h, 0, sizeof(h));
    for (int k = 0; k < 3; ++k) h.minKey[k] = 0xFFFFFFFFu;
}

}  // namespace gsm

using namespace gsm;

#define SCENE_CUDA(call, what)                                                          \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) { st = reportFailure(GSM_ERR_RENDER_FAILED, what, e__); goto done; } \
    } while (0)

extern "C" {

gsm_status gsm_ply_probe(const void* fileBytes, size_t fileSize, gsm_ply_info* info) {
    if (!fileBytes || !info) return reportFailure(GSM_ERR_INVALID_ARGUMENT, "null argument");
    std::memset(info, 0, sizeof(*info));
    PlyHeader h;
    gsm_status st = parsePlyHeader((const unsigned char*)fileBytes, fileSize, h);
    if (st != GSM_OK) return st;
    info->format = (uint32_t)h.format;
    info->bodyOffset = h.bodyStart;
    const PlyElem* vx = findElem(h, "vertex");
    if (!vx) return reportFailure((gsm_status)GSM_ERR_PLY_MISSING_VERTEX, "PLY: no 'vertex' element");
    info->vertexCount = vx->count;
    info->compressed = isCompressed(h, *vx) ? 1u : 0u;
    uint32_t nsh = 0;
    for (auto& p : vx->props)
        if (isShName(lowered(p.name))) nsh++;
    info->shProperties = info->compressed ? 3u : nsh;
    return GSM_OK;
}

gsm_status gsm_ply_load(int device, void* stream, const void* fileBytes, size_t fileSize, int precision, void* gaussiansOut,
                        void* harmonicsOut, uint32_t gaussianCapacity, size_t harmonicsCapacity, gsm_scene_info* info) {
    if (!fileBytes || !gaussiansOut || !info) return reportFailure(GSM_ERR_INVALID_ARGUMENT, "null argument");
    std::memset(info, 0, sizeof(*info));
    const unsigned char* data = (const unsigned char*)fileBytes;
    PlyHeader h;
    gsm_status st = parsePlyHeader(data, fileSize, h);
    if (st != GSM_OK) return st;
    if (h.format != 1) return reportFailure((gsm_status)GSM_ERR_PLY_UNSUPPORTED_FORMAT, "PLY: only binary_little_endian is supported");
    const PlyElem* vx = findElem(h, "vertex");
    if (!vx) return reportFailure((gsm_status)GSM_ERR_PLY_MISSING_VERTEX, "PLY: no 'vertex' element");
    const bool compressed = isCompressed(h, *vx);
    const uint32_t n = vx->count;
    const bool half = precision == GSM_PRECISION_FLOAT16;

    StandardLayout SL;
    CompressedLayout CL;
    std::memset(&SL, 0, sizeof(SL));
    std::memset(&CL, 0, sizeof(CL));
    size_t bodyBytes = 0;
    uint32_t shStride = 0, shComponents = 0;
    if (!compressed) {
        for (auto& p : vx->props)
            if (p.type == T_LIST) return reportFailure((gsm_status)GSM_ERR_PLY_LIST_PROPERTY, "PLY: list properties in the vertex element are not supported");
        std::vector<int> offsets;
        int stride = 0;
        for (auto& p : vx->props) { offsets.push_back(stride); stride += plyTypeWidth(p.type); }
        if (fileSize - h.bodyStart < (size_t)stride * n) return reportFailure((gsm_status)GSM_ERR_PLY_INSUFFICIENT_DATA, "PLY: insufficient data for the declared vertex count");
        for (int i = 0; i < 11; ++i) SL.off[i] = -1;
        struct ShEntry { long key; int idx; };
        std::vector<ShEntry> shs;
        static const char* const alias[11][4] = {  // PLYLoader.swift:547-561
            {"x", "px", "pos_x", "position_x"}, {"y", "py", "pos_y", "position_y"}, {"z", "pz", "pos_z", "position_z"},
            {"scale_0", "scale0", "sx", "scale_x"}, {"scale_1", "scale1", "sy", "scale_y"}, {"scale_2", "scale2", "sz", "scale_z"},
            {"rot_0", "rot0", "qw", "rotation_w"}, {"rot_1", "rot1", "qx", "rotation_x"}, {"rot_2", "rot2", "qy", "rotation_y"},
            {"rot_3", "rot3", "qz", "rotation_z"}, {"opacity", "alpha", "opacity", "alpha"}};
        for (size_t i = 0; i < vx->props.size(); ++i) {
            const std::string nm = lowered(vx->props[i].name);
            bool matched = false;
            for (int a = 0; a < 11 && !matched; ++a)
                for (int b = 0; b < 4; ++b)
                    if (nm == alias[a][b]) { SL.off[a] = offsets[i]; SL.type[a] = vx->props[i].type; matched = true; break; }
            if (!matched && isShName(nm)) {
                long key = 0x7FFFFFFFL;  // PLYLoader.swift:575-580
                if (startsWith(nm, "f_dc_")) key = std::atol(nm.c_str() + 5);
                else if (startsWith(nm, "f_rest_")) key = 3 + std::atol(nm.c_str() + 7);
                else if (startsWith(nm, "sh_")) key = std::atol(nm.c_str() + 3);
                shs.push_back({key, (int)i});
            }
        }
        if (SL.off[0] < 0 || SL.off[1] < 0 || SL.off[2] < 0) return reportFailure((gsm_status)GSM_ERR_PLY_MISSING_PROPERTIES, "PLY: missing required properties x, y, z");
        if (shs.size() > (size_t)kMaxShProps) return reportFailure((gsm_status)GSM_ERR_PLY_INVALID_HEADER, "PLY: more than 64 SH properties");
        for (size_t a = 1; a < shs.size(); ++a) {  // stable insertion sort by key
            ShEntry v = shs[a];
            size_t b = a;
            while (b > 0 && shs[b - 1].key > v.key) { shs[b] = shs[b - 1]; b--; }
            shs[b] = v;
        }
        for (size_t k = 0; k < shs.size(); ++k) { SL.shOff[k] = offsets[shs[k].idx]; SL.shType[k] = vx->props[shs[k].idx].type; }
        SL.shProps = (uint32_t)shs.size();
        SL.shComponents = SL.shProps / 3u;  // PLYLoader.swift:693
        SL.stride = (uint32_t)stride;
        SL.count = n;
        // format detection on the first 100 vertices (host, on the mapped file), PLYLoader.swift:620-650
        auto hostProp = [&](uint32_t v, int i) -> float {
            if (SL.off[i] < 0) return 0.0f;
            const unsigned char* p = data + h.bodyStart + (size_t)v * stride + SL.off[i];
            switch (SL.type[i]) {
                case T_F32: { float f; std::memcpy(&f, p, 4); return f; }
                case T_F64: { double f; std::memcpy(&f, p, 8); return (float)f; }
                case T_U8: return (float)p[0] / 255.0f;
                case T_I8: return (float)(signed char)p[0];
                case T_I16: { int16_t f; std::memcpy(&f, p, 2); return (float)f; }
                case T_U16: { uint16_t f; std::memcpy(&f, p, 2); return (float)f; }
                case T_I32: { int32_t f; std::memcpy(&f, p, 4); return (float)f; }
                case T_U32: { uint32_t f; std::memcpy(&f, p, 4); return (float)f; }
                default: return 0.0f;
            }
        };
        SL.scaleIsLogSpace = 1; SL.opacityIsLogit = 1;
        const uint32_t sampleCount = n < 100u ? n : 100u;
        if (SL.off[3] >= 0 && sampleCount > 0) {
            bool hasNeg = false, hasLarge = false;
            float sum = 0.0f;
            for (uint32_t v = 0; v < sampleCount; ++v) { const float s = hostProp(v, 3); hasNeg |= s < 0.0f; hasLarge |= s > 1.0f; sum += s; }
            const float avg = sum / (float)sampleCount;
            if (hasNeg) SL.scaleIsLogSpace = 1;
            else if (!hasLarge && avg > 0.0f && avg < 0.5f) SL.scaleIsLogSpace = 0;
        }
        if (SL.off[10] >= 0 && sampleCount > 0) {
            float mn = hostProp(0, 10), mx = mn;
            for (uint32_t v = 0; v < sampleCount; ++v) { const float o = hostProp(v, 10); mn = o < mn ? o : mn; mx = o > mx ? o : mx; }
            SL.opacityIsLogit = (mn < 0.0f || mx > 1.0f) ? 1u : 0u;
        }
        bodyBytes = (size_t)stride * n;
        shComponents = SL.shComponents;
        shStride = shComponents > 0 ? SL.shProps : 0u;
        info->scaleIsLogSpace = SL.scaleIsLogSpace;
        info->opacityIsLogit = SL.opacityIsLogit;
    } else {
        const PlyElem* ch = findElem(h, "chunk");
        if (!ch) return reportFailure((gsm_status)GSM_ERR_PLY_MISSING_CHUNK, "PLY: compressed layout without a 'chunk' element");
        int chunkStride = 0, vertexStride = 0, shElemStride = 0;
        for (auto& p : ch->props) chunkStride += plyTypeWidth(p.type);
        for (auto& p : vx->props) vertexStride += plyTypeWidth(p.type);
        if (const PlyElem* sh = findElem(h, "sh")) for (auto& p : sh->props) shElemStride += plyTypeWidth(p.type);
        const size_t vertexStart = (size_t)chunkStride * ch->count, shStart = vertexStart + (size_t)vertexStride * n;
        if (fileSize < h.bodyStart + shStart + (size_t)shElemStride * n) return reportFailure((gsm_status)GSM_ERR_PLY_INSUFFICIENT_DATA, "PLY: insufficient data for the declared vertex count");
        static const char* const cnames[18] = {"min_x", "min_y", "min_z", "max_x", "max_y", "max_z", "min_scale_x", "min_scale_y", "min_scale_z",
                                               "max_scale_x", "max_scale_y", "max_scale_z", "min_r", "min_g", "min_b", "max_r", "max_g", "max_b"};
        static const char* const vnames[4] = {"packed_position", "packed_rotation", "packed_scale", "packed_color"};
        for (int k = 0; k < 18; ++k) { CL.chunkOff[k] = -1; int o = 0; for (auto& p : ch->props) { if (p.name == cnames[k]) CL.chunkOff[k] = o; o += plyTypeWidth(p.type); } }
        for (int k = 0; k < 4; ++k) { CL.vertexOff[k] = -1; int o = 0; for (auto& p : vx->props) { if (p.name == vnames[k]) CL.vertexOff[k] = o; o += plyTypeWidth(p.type); } }
        CL.chunkStride = (uint32_t)chunkStride; CL.vertexStride = (uint32_t)vertexStride; CL.count = n; CL.vertexStart = vertexStart;
        bodyBytes = shStart;
        shComponents = 1; shStride = 3;
        info->scaleIsLogSpace = 1; info->opacityIsLogit = 0;
    }
    info->compressed = compressed ? 1u : 0u;
    info->shComponents = shComponents;
    info->harmonicsStride = shStride;
    if (n > gaussianCapacity || (size_t)n * shStride > harmonicsCapacity || (shStride > 0 && !harmonicsOut))
        return reportFailure((gsm_status)GSM_ERR_PLY_INSUFFICIENT_DATA, "PLY: the caller's device buffers are too small for this file");

    if (device < 0) cudaGetDevice(&device);
    int prevDevice = -1;
    cudaGetDevice(&prevDevice);
    if (prevDevice != device) cudaSetDevice(device);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char* dBody = nullptr;
    float *dPos = nullptr, *dScale = nullptr, *dRot = nullptr, *dOp = nullptr, *dSh = nullptr;
    uint8_t* dKeep = nullptr;
    uint32_t* dIndex = nullptr;
    unsigned long long* dWords = nullptr;
    SceneScratch* dScratch = nullptr;
    SceneScratch hs;
    const uint32_t blocks = (n + 255u) / 256u;
    const size_t nn = n > 0 ? n : 1;
    st = GSM_OK;
    {
        // the body lands 16-byte aligned on the device whatever the header length was
        SCENE_CUDA(cudaMalloc((void**)&dBody, bodyBytes + 16), "scene: body buffer");
        if (compressed) {  // the compressed layout decodes into float records first (16 B in, 60 B out per vertex)
            SCENE_CUDA(cudaMalloc((void**)&dPos, nn * 12), "scene: records");
            SCENE_CUDA(cudaMalloc((void**)&dScale, nn * 12), "scene: records");
            SCENE_CUDA(cudaMalloc((void**)&dRot, nn * 16), "scene: records");
            SCENE_CUDA(cudaMalloc((void**)&dOp, nn * 4), "scene: records");
            SCENE_CUDA(cudaMalloc((void**)&dSh, nn * (shStride ? shStride : 1) * 4), "scene: harmonics");
        }
        SCENE_CUDA(cudaMalloc((void**)&dKeep, nn), "scene: flags");
        SCENE_CUDA(cudaMalloc((void**)&dScratch, sizeof(SceneScratch)), "scene: scratch");
        initScratch(hs);
        SCENE_CUDA(cudaMemcpyAsync(dScratch, &hs, sizeof(hs), cudaMemcpyHostToDevice, s), "scene: scratch init");
        if (bodyBytes) SCENE_CUDA(cudaMemcpyAsync(dBody, data + h.bodyStart, bodyBytes, cudaMemcpyHostToDevice, s), "scene: body upload");
        if (n > 0) {
            if (compressed) ply_decode_compressed_kernel<<<blocks, 256, 0, s>>>(dBody, CL, dPos, dScale, dRot, dOp, dSh, dKeep, dScratch);
            else ply_scan_standard_kernel<<<blocks, 256, 0, s>>>(dBody, SL, dKeep, dScratch);
            SCENE_CUDA(cudaGetLastError(), "scene: decode kernel");
        }
        SCENE_CUDA(cudaMemcpyAsync(&hs, dScratch, sizeof(hs), cudaMemcpyDeviceToHost, s), "scene: scratch readback");
        SCENE_CUDA(cudaStreamSynchronize(s), "scene: decode sync");
        uint32_t kept = n;
        if (hs.placeholders > 0 && n > 0) {
            const size_t tiles = blocks, words = tiles + (tiles + 31) / 32 + 8;
            SCENE_CUDA(cudaMalloc((void**)&dIndex, nn * 4), "scene: compaction index");
            SCENE_CUDA(cudaMalloc((void**)&dWords, words * 8), "scene: prefix words");
            SCENE_CUDA(cudaMemsetAsync(dWords, 0, words * 8, s), "scene: prefix words");
            int numSMs = 1;
            cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, device);
            const uint32_t grid = blocks < (uint32_t)numSMs * 8u ? blocks : (uint32_t)numSMs * 8u;
            ply_keep_index_kernel<<<grid, 256, 0, s>>>(dKeep, n, dIndex, dWords, dWords + tiles, dScratch);
            SCENE_CUDA(cudaGetLastError(), "scene: compaction kernel");
            kept = n - hs.placeholders;
        }
        if (n > 0 && compressed) {
            if (half) scene_pack_kernel<true><<<blocks, 256, 0, s>>>(dPos, dScale, dRot, dOp, dSh, dKeep, dIndex, n, shStride, gaussiansOut, harmonicsOut, dScratch);
            else scene_pack_kernel<false><<<blocks, 256, 0, s>>>(dPos, dScale, dRot, dOp, dSh, dKeep, dIndex, n, shStride, gaussiansOut, harmonicsOut, dScratch);
            SCENE_CUDA(cudaGetLastError(), "scene: pack kernel");
        } else if (n > 0) {
            // shared memory: the block's records (if they fit) + its planar SH rows (when the destination rows are contiguous)
            const size_t recBytes = ((size_t)kPlyBlock * SL.stride + 15u) & ~(size_t)15u;
            const size_t shBytes = (dIndex == nullptr) ? (size_t)kPlyBlock * shStride * 4u : 0u;
            const uint32_t stage = (recBytes + shBytes <= 160u * 1024u) ? 1u : 0u;
            const size_t smem = (stage ? recBytes : 0u) + shBytes;
            const uint32_t pblocks = (n + kPlyBlock - 1u) / kPlyBlock;
            if (half) {
                SCENE_CUDA(cudaFuncSetAttribute(ply_decode_pack_standard_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "scene: shared memory opt-in");
                ply_decode_pack_standard_kernel<true><<<pblocks, kPlyBlock, smem, s>>>(dBody, SL, dKeep, dIndex, gaussiansOut, harmonicsOut, dScratch, stage);
            } else {
                SCENE_CUDA(cudaFuncSetAttribute(ply_decode_pack_standard_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "scene: shared memory opt-in");
                ply_decode_pack_standard_kernel<false><<<pblocks, kPlyBlock, smem, s>>>(dBody, SL, dKeep, dIndex, gaussiansOut, harmonicsOut, dScratch, stage);
            }
            SCENE_CUDA(cudaGetLastError(), "scene: decode+pack kernel");
        }
        SCENE_CUDA(cudaMemcpyAsync(&hs, dScratch, sizeof(hs), cudaMemcpyDeviceToHost, s), "scene: scratch readback");
        SCENE_CUDA(cudaStreamSynchronize(s), "scene: pack sync");
        info->count = kept;
        // center / bounds on the host from the reduced keys, with the same float arithmetic as the kernels
        if (kept > 0 && hs.minKey[0] != 0xFFFFFFFFu) {
            const SceneCenters cen = sceneCenters(hs.minKey, hs.maxKey);
            float d[3];
            for (int k = 0; k < 3; ++k) {
                info->center[k] = cen.c[k];
                info->boundsCenter[k] = cen.c2[k];
                const float mx = sceneKeyToFloat(hs.maxKey[k]);
                d[k] = (cen.shift ? mx - cen.c[k] : mx) - cen.c2[k];
            }
            const float far2 = sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);  // Scene.swift:187
            float r;
            std::memcpy(&r, &hs.radiusKey, 4);
            r = r > far2 ? r : far2;
            info->boundsRadius = r > 0.5f ? r : 0.5f;
        } else {
            info->boundsRadius = 1.0f;  // Scene.swift:160
        }
    }
done:
    cudaFree(dBody); cudaFree(dPos); cudaFree(dScale); cudaFree(dRot); cudaFree(dOp); cudaFree(dSh); cudaFree(dKeep);
    cudaFree(dIndex); cudaFree(dWords); cudaFree(dScratch);
    if (prevDevice != device && prevDevice >= 0) cudaSetDevice(prevDevice);
    return st;
}

gsm_status gsm_scene_morton_sort(int device, void* stream, void* gaussians, void* harmonics, uint32_t count, uint32_t harmonicsStride,
                                 int precision) {
    if (!gaussians || (harmonicsStride > 0 && !harmonics)) return reportFailure(GSM_ERR_INVALID_ARGUMENT, "null argument");
    if (count <= 1) return GSM_OK;  // Scene.swift:78
    if (device < 0) cudaGetDevice(&device);
    int prevDevice = -1;
    cudaGetDevice(&prevDevice);
    if (prevDevice != device) cudaSetDevice(device);
    cudaStream_t s = (cudaStream_t)stream;
    const bool half = precision == GSM_PRECISION_FLOAT16;
    const uint32_t recBytes = half ? 32u : 48u, elemBytes = half ? 2u : 4u;
    const uint32_t blocks = (count + 255u) / 256u;
    uint32_t *dLo, *dHi, *dIdx, *dKey;
    unsigned char *dRec, *dSh, *arena = nullptr;
    void* dSortScratch;
    SceneScratch* dScratch;
    SceneScratch hs;
    gsm_status st = GSM_OK;
    int numSMs = 1;
    cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, device);
    {
        // one allocation for every temporary (allocation and free calls would otherwise cost more than the kernels)
        const size_t shBytes = (size_t)count * harmonicsStride * elemBytes;
        auto up = [](size_t v) { return (v + 255) / 256 * 256; };
        const size_t keyB = up((size_t)count * 4), oHi = keyB, oKey = 2 * keyB, oIdx = 3 * keyB, oRec = 4 * keyB;
        const size_t oSh = oRec + up((size_t)count * recBytes), oScr = oSh + up(shBytes), oSort = oScr + 256;
        const size_t total = oSort + sortScratchBytes(count, 32, 4);
        SCENE_CUDA(cudaMalloc((void**)&arena, total), "morton: temporaries");
        dLo = (uint32_t*)arena; dHi = (uint32_t*)(arena + oHi); dKey = (uint32_t*)(arena + oKey); dIdx = (uint32_t*)(arena + oIdx);
        dRec = arena + oRec; dSh = arena + oSh; dScratch = (SceneScratch*)(arena + oScr); dSortScratch = arena + oSort;
        initScratch(hs);
        SCENE_CUDA(cudaMemcpyAsync(dScratch, &hs, sizeof(hs), cudaMemcpyHostToDevice, s), "morton: scratch init");
        scene_position_bounds_kernel<<<blocks, 256, 0, s>>>((const unsigned char*)gaussians, recBytes, count, dScratch);
        scene_morton_codes_kernel<<<blocks, 256, 0, s>>>((const unsigned char*)gaussians, recBytes, count, dScratch, dLo, dHi, dIdx);
        SCENE_CUDA(cudaGetLastError(), "morton: code kernels");
        // stable LSD over the 64-bit code: sort (lo, index), then (hi[index], index)
        st = sortPairsStandalone(s, numSMs, dLo, dIdx, count, 32, 4, dSortScratch);
        if (st != GSM_OK) goto done;
        gather_u32_kernel<<<blocks, 256, 0, s>>>(dHi, dIdx, count, dKey);
        SCENE_CUDA(cudaGetLastError(), "morton: gather");
        st = sortPairsStandalone(s, numSMs, dKey, dIdx, count, 32, 4, dSortScratch);
        if (st != GSM_OK) goto done;
        {
            const uint32_t words = recBytes / 4u;
            const size_t totalWords = (size_t)count * words;
            permute_records_kernel<<<(unsigned)((totalWords + 255) / 256), 256, 0, s>>>((const uint32_t*)gaussians, dIdx, count, words, (uint32_t*)dRec);
            SCENE_CUDA(cudaMemcpyAsync(gaussians, dRec, (size_t)count * recBytes, cudaMemcpyDeviceToDevice, s), "morton: records back");
        }
        if (shBytes) {
            const size_t totalElems = (size_t)count * harmonicsStride;
            const uint32_t rowBytes = harmonicsStride * elemBytes;
            if (rowBytes % 4u == 0u) {  // whole 32-bit words per row (always for float, and for half rows of even length)
                const uint32_t words = rowBytes / 4u;
                const size_t totalWords = (size_t)count * words;
                permute_records_kernel<<<(unsigned)((totalWords + 255) / 256), 256, 0, s>>>((const uint32_t*)harmonics, dIdx, count, words, (uint32_t*)dSh);
            } else {
                permute_halfs_kernel<<<(unsigned)((totalElems + 255) / 256), 256, 0, s>>>((const uint16_t*)harmonics, dIdx, count, harmonicsStride, (uint16_t*)dSh);
            }
            SCENE_CUDA(cudaMemcpyAsync(harmonics, dSh, shBytes, cudaMemcpyDeviceToDevice, s), "morton: harmonics back");
        }
        SCENE_CUDA(cudaGetLastError(), "morton: permute kernels");
        SCENE_CUDA(cudaStreamSynchronize(s), "morton: sync");
    }
done:
    if (st != GSM_OK) cudaStreamSynchronize(s);
    cudaFree(arena);
    if (prevDevice != device && prevDevice >= 0) cudaSetDevice(prevDevice);
    return st;
}

}  // extern "C"
