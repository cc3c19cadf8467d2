/*
 * gsm_oracle_ply.c -- CPU restatement of the reference's scene ingest (SURVEY.md 8(f) rank 1 and the Morton pre-sort of
 * rank 2): PLY header parsing, the standard 3DGS vertex layout, the PlayCanvas / splat-transform compressed layout, SH
 * re-layout to planar, recentering, scene bounds, Morton sort and the packing into PackedWorldGaussian(+Half).
 *
 * TEST INFRASTRUCTURE ONLY (see gsm_oracle.h): nothing here is linked, imported or executed by the product path.
 *
 * Follows: PLYLoader.swift:88-205 (header), :246-281 (format dispatch), :285-513 (compressed), :517-741 (standard),
 * Scene.swift:47-138 (Morton), :159-190 (bounds), PLYBenchmarkTests.swift:139-149 (packing).
 *
 * PARITY UNPINNED for transcendental bits: the reference calls the platform's expf / sqrt / simd_normalize on the CPU; this
 * file uses the canonical gsmo_exp (gsmo_math.h), IEEE sqrt and q * (1 / sqrt(q.q)). Everything structural -- property
 * mapping and aliases, type conversion, format detection on the first 100 vertices, placeholder skipping, SH ordering and
 * planar layout, recentering threshold, bounds, Morton code and order -- is exact and is what the tests pin.
 */
#include <ctype.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gsmo_math.h"

#define GSMO_PLY_MAX_PROPS 256
#define GSMO_PLY_MAX_ELEMS 8

enum { GSMO_T_I8, GSMO_T_U8, GSMO_T_I16, GSMO_T_U16, GSMO_T_I32, GSMO_T_U32, GSMO_T_F32, GSMO_T_F64, GSMO_T_LIST };

typedef struct { char name[64]; int type; } gsmo_ply_prop;
typedef struct { char name[32]; uint32_t count; int nprops; gsmo_ply_prop props[GSMO_PLY_MAX_PROPS]; } gsmo_ply_elem;
typedef struct {
    int format; /* 0 ascii, 1 binary_little_endian, 2 binary_big_endian */
    int nelems;
    gsmo_ply_elem elems[GSMO_PLY_MAX_ELEMS];
    size_t bodyStart;
} gsmo_ply_header;

/* error codes = PLYLoaderError / DecodeError cases (PLYLoader.swift:88-110, :209-242) */
enum {
    GSMO_PLY_OK = 0, GSMO_PLY_INVALID_HEADER = 1, GSMO_PLY_UNSUPPORTED_FORMAT = 2, GSMO_PLY_MISSING_VERTEX = 3,
    GSMO_PLY_MISSING_REQUIRED = 4, GSMO_PLY_LIST_NOT_SUPPORTED = 5, GSMO_PLY_INSUFFICIENT_DATA = 6,
    GSMO_PLY_MISSING_CHUNK = 7, GSMO_PLY_HEADER_SYNTAX = 8
};

static int typeWidth(int t) {
    switch (t) {
        case GSMO_T_I8: case GSMO_T_U8: return 1;
        case GSMO_T_I16: case GSMO_T_U16: return 2;
        case GSMO_T_I32: case GSMO_T_U32: case GSMO_T_F32: return 4;
        case GSMO_T_F64: return 8;
        default: return 0; /* list: PLYLoader.swift:40-45 */
    }
}
static int typeFromString(const char* s) { /* PLYLoader.swift:191-205 */
    if (!strcmp(s, "int8") || !strcmp(s, "char")) return GSMO_T_I8;
    if (!strcmp(s, "uint8") || !strcmp(s, "uchar")) return GSMO_T_U8;
    if (!strcmp(s, "int16") || !strcmp(s, "short")) return GSMO_T_I16;
    if (!strcmp(s, "uint16") || !strcmp(s, "ushort")) return GSMO_T_U16;
    if (!strcmp(s, "int32") || !strcmp(s, "int")) return GSMO_T_I32;
    if (!strcmp(s, "uint32") || !strcmp(s, "uint")) return GSMO_T_U32;
    if (!strcmp(s, "float32") || !strcmp(s, "float")) return GSMO_T_F32;
    if (!strcmp(s, "float64") || !strcmp(s, "double")) return GSMO_T_F64;
    return -1;
}

static const unsigned char* findBytes(const unsigned char* hay, size_t n, const char* needle) {
    const size_t m = strlen(needle);
    if (n < m) return NULL;
    for (size_t i = 0; i + m <= n; ++i)
        if (!memcmp(hay + i, needle, m)) return hay + i;
    return NULL;
}

/* PLYLoader.swift:112-187 (line-wise keyword parse) + :250-258 (end_header with LF first, then CRLF) */
static int parseHeader(const unsigned char* data, size_t size, gsmo_ply_header* h) {
    memset(h, 0, sizeof(*h));
    h->format = -1;
    const unsigned char* end = findBytes(data, size, "end_header\n");
    size_t endLen = 11;
    if (!end) { end = findBytes(data, size, "end_header\r\n"); endLen = 12; }
    if (!end) return GSMO_PLY_INVALID_HEADER;
    h->bodyStart = (size_t)(end - data) + endLen;
    size_t pos = 0;
    while (pos < h->bodyStart) {
        size_t eol = pos;
        while (eol < h->bodyStart && data[eol] != '\n' && data[eol] != '\r') eol++;
        char line[512];
        size_t len = eol - pos < sizeof(line) - 1 ? eol - pos : sizeof(line) - 1;
        memcpy(line, data + pos, len);
        line[len] = 0;
        pos = eol;
        while (pos < h->bodyStart && (data[pos] == '\n' || data[pos] == '\r')) pos++;
        char* tok[8];
        int nt = 0;
        for (char* p = strtok(line, " \t"); p && nt < 8; p = strtok(NULL, " \t")) tok[nt++] = p;
        if (nt == 0) continue;
        if (!strcmp(tok[0], "ply") || !strcmp(tok[0], "comment") || !strcmp(tok[0], "obj_info")) continue;
        if (!strcmp(tok[0], "end_header")) break;
        if (!strcmp(tok[0], "format")) {
            if (h->format >= 0 || nt < 3) return GSMO_PLY_HEADER_SYNTAX;
            if (!strcmp(tok[1], "ascii")) h->format = 0;
            else if (!strcmp(tok[1], "binary_little_endian")) h->format = 1;
            else if (!strcmp(tok[1], "binary_big_endian")) h->format = 2;
            else return GSMO_PLY_HEADER_SYNTAX;
        } else if (!strcmp(tok[0], "element")) {
            if (h->format < 0 || nt < 3 || h->nelems >= GSMO_PLY_MAX_ELEMS) return GSMO_PLY_HEADER_SYNTAX;
            gsmo_ply_elem* e = &h->elems[h->nelems++];
            strncpy(e->name, tok[1], sizeof(e->name) - 1);
            e->count = (uint32_t)strtoul(tok[2], NULL, 10);
        } else if (!strcmp(tok[0], "property")) {
            if (h->format < 0 || h->nelems == 0) return GSMO_PLY_HEADER_SYNTAX;
            gsmo_ply_elem* e = &h->elems[h->nelems - 1];
            if (e->nprops >= GSMO_PLY_MAX_PROPS) return GSMO_PLY_HEADER_SYNTAX;
            gsmo_ply_prop* p = &e->props[e->nprops++];
            if (nt >= 5 && !strcmp(tok[1], "list")) {
                if (typeFromString(tok[2]) < 0 || typeFromString(tok[3]) < 0) return GSMO_PLY_HEADER_SYNTAX;
                p->type = GSMO_T_LIST;
                strncpy(p->name, tok[4], sizeof(p->name) - 1);
            } else if (nt >= 3) {
                p->type = typeFromString(tok[1]);
                if (p->type < 0) return GSMO_PLY_HEADER_SYNTAX;
                strncpy(p->name, tok[2], sizeof(p->name) - 1);
            } else {
                return GSMO_PLY_HEADER_SYNTAX;
            }
        } else {
            return GSMO_PLY_HEADER_SYNTAX; /* headerUnknownKeyword */
        }
    }
    if (h->format < 0) return GSMO_PLY_HEADER_SYNTAX; /* headerFormatMissing */
    return GSMO_PLY_OK;
}

static const gsmo_ply_elem* findElem(const gsmo_ply_header* h, const char* name) {
    for (int i = 0; i < h->nelems; ++i)
        if (!strcmp(h->elems[i].name, name)) return &h->elems[i];
    return NULL;
}
static int hasProp(const gsmo_ply_elem* e, const char* name) {
    for (int i = 0; i < e->nprops; ++i)
        if (!strcmp(e->props[i].name, name)) return 1;
    return 0;
}

/* getFloat, PLYLoader.swift:598-618 */
static float readProp(const unsigned char* p, int type) {
    switch (type) {
        case GSMO_T_F32: { float v; memcpy(&v, p, 4); return v; }
        case GSMO_T_F64: { double v; memcpy(&v, p, 8); return (float)v; }
        case GSMO_T_U8: return (float)p[0] / 255.0f;
        case GSMO_T_I8: return (float)(int8_t)p[0];
        case GSMO_T_I16: { int16_t v; memcpy(&v, p, 2); return (float)v; }
        case GSMO_T_U16: { uint16_t v; memcpy(&v, p, 2); return (float)v; }
        case GSMO_T_I32: { int32_t v; memcpy(&v, p, 4); return (float)v; }
        case GSMO_T_U32: { uint32_t v; memcpy(&v, p, 4); return (float)v; }
        default: return 0.0f;
    }
}

typedef struct {
    uint32_t count;           /* records kept (placeholders skipped) */
    uint32_t shComponents;    /* GaussianDataset.shComponents */
    uint32_t harmonicsStride; /* floats per Gaussian in harmonics[] (3 * shComponents, or the raw SH property count) */
    uint32_t compressed, scaleIsLogSpace, opacityIsLogit;
    float center[3];          /* the center that was subtracted (zero if |center| <= 1e-6) */
    float boundsCenter[3];    /* GaussianSceneBuilder.bounds of the FINAL records */
    float boundsRadius;
} gsmo_ply_result;

/* GaussianSceneBuilder.bounds, Scene.swift:159-190. pos/scale: 3 floats per record. */
void gsmo_scene_bounds(const float* pos, const float* scale, uint32_t n, float center[3], float* radius) {
    if (n == 0) { center[0] = center[1] = center[2] = 0.0f; *radius = 1.0f; return; }
    float mn[3] = {pos[0], pos[1], pos[2]}, mx[3] = {pos[0], pos[1], pos[2]};
    for (uint32_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            mn[k] = gsmo_fmin(mn[k], pos[3 * i + k]);
            mx[k] = gsmo_fmax(mx[k], pos[3 * i + k]);
        }
    for (int k = 0; k < 3; ++k) center[k] = (mn[k] + mx[k]) * 0.5f;
    float r = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        float ox = pos[3 * i] - center[0], oy = pos[3 * i + 1] - center[1], oz = pos[3 * i + 2] - center[2];
        float sm = gsmo_fmax(scale[3 * i], gsmo_fmax(scale[3 * i + 1], scale[3 * i + 2]));
        r = gsmo_fmax(r, sqrtf((ox * ox + oy * oy) + oz * oz) + sm);
    }
    float dx = mx[0] - center[0], dy = mx[1] - center[1], dz = mx[2] - center[2];
    r = gsmo_fmax(r, sqrtf((dx * dx + dy * dy) + dz * dz));
    *radius = gsmo_fmax(r, 0.5f);
}

static void recenter(float* pos, uint32_t n, gsmo_ply_result* res) { /* PLYLoader.swift:496-503, :722-730 */
    float mn[3], mx[3];
    res->center[0] = res->center[1] = res->center[2] = 0.0f;
    if (n == 0) return;
    for (int k = 0; k < 3; ++k) mn[k] = mx[k] = pos[k];
    for (uint32_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            mn[k] = gsmo_fmin(mn[k], pos[3 * i + k]);
            mx[k] = gsmo_fmax(mx[k], pos[3 * i + k]);
        }
    float c[3];
    for (int k = 0; k < 3; ++k) c[k] = (mn[k] + mx[k]) * 0.5f;
    if (sqrtf((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]) > 1e-6f) {
        for (uint32_t i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) pos[3 * i + k] -= c[k];
        for (int k = 0; k < 3; ++k) res->center[k] = c[k];
    }
}

static int lowerEq(const char* a, const char* b) {
    for (; *a && *b; ++a, ++b)
        if (tolower((unsigned char)*a) != *b) return 0;
    return *a == 0 && *b == 0;
}
static int lowerStarts(const char* a, const char* prefix) {
    for (; *prefix; ++a, ++prefix)
        if (tolower((unsigned char)*a) != *prefix) return 0;
    return 1;
}
static int anyOf(const char* name, const char* a, const char* b, const char* c, const char* d) {
    return lowerEq(name, a) || lowerEq(name, b) || lowerEq(name, c) || lowerEq(name, d);
}
static long shSortKey(const char* name) { /* PLYLoader.swift:575-580 (on the lowercased name) */
    if (lowerStarts(name, "f_dc_")) return atol(name + 5);
    if (lowerStarts(name, "f_rest_")) return 3 + atol(name + 7);
    if (lowerStarts(name, "sh_")) return atol(name + 3);
    return 0x7FFFFFFFL;
}

/* Outputs are caller-allocated for vertexCount records: pos[3n], scale[3n], rot[4n] as (x, y, z, w), opacity[n],
 * harmonics[n * shProps]. Returns a GSMO_PLY_* code. */
int gsmo_ply_load(const unsigned char* data, size_t size, float* pos, float* scale, float* rot, float* opacity,
                  float* harmonics, size_t harmonicsCapacity, gsmo_ply_result* res);

static int loadStandard(const unsigned char* data, size_t size, const gsmo_ply_header* h, const gsmo_ply_elem* vx,
                        float* pos, float* scale, float* rot, float* opacity, float* harmonics, size_t harmonicsCapacity,
                        gsmo_ply_result* res) {
    for (int i = 0; i < vx->nprops; ++i)
        if (vx->props[i].type == GSMO_T_LIST) return GSMO_PLY_LIST_NOT_SUPPORTED;
    int offsets[GSMO_PLY_MAX_PROPS];
    int stride = 0;
    for (int i = 0; i < vx->nprops; ++i) { offsets[i] = stride; stride += typeWidth(vx->props[i].type); }
    const uint32_t n = vx->count;
    if (size - h->bodyStart < (size_t)stride * n) return GSMO_PLY_INSUFFICIENT_DATA;
    int ix = -1, iy = -1, iz = -1, is0 = -1, is1 = -1, is2 = -1, ir0 = -1, ir1 = -1, ir2 = -1, ir3 = -1, iop = -1;
    int shIdx[GSMO_PLY_MAX_PROPS];
    long shKey[GSMO_PLY_MAX_PROPS];
    int nsh = 0;
    for (int i = 0; i < vx->nprops; ++i) { /* PLYLoader.swift:547-569 */
        const char* nm = vx->props[i].name;
        if (anyOf(nm, "x", "px", "pos_x", "position_x")) ix = i;
        else if (anyOf(nm, "y", "py", "pos_y", "position_y")) iy = i;
        else if (anyOf(nm, "z", "pz", "pos_z", "position_z")) iz = i;
        else if (anyOf(nm, "scale_0", "scale0", "sx", "scale_x")) is0 = i;
        else if (anyOf(nm, "scale_1", "scale1", "sy", "scale_y")) is1 = i;
        else if (anyOf(nm, "scale_2", "scale2", "sz", "scale_z")) is2 = i;
        else if (anyOf(nm, "rot_0", "rot0", "qw", "rotation_w")) ir0 = i;
        else if (anyOf(nm, "rot_1", "rot1", "qx", "rotation_x")) ir1 = i;
        else if (anyOf(nm, "rot_2", "rot2", "qy", "rotation_y")) ir2 = i;
        else if (anyOf(nm, "rot_3", "rot3", "qz", "rotation_z")) ir3 = i;
        else if (lowerEq(nm, "opacity") || lowerEq(nm, "alpha")) iop = i;
        else if (lowerStarts(nm, "f_dc_") || lowerStarts(nm, "f_rest_") || lowerStarts(nm, "sh_") ||
                 lowerStarts(nm, "spherical_harmonics_")) {
            shIdx[nsh] = i; shKey[nsh] = shSortKey(nm); nsh++;
        }
    }
    if (ix < 0 || iy < 0 || iz < 0) return GSMO_PLY_MISSING_REQUIRED;
    for (int a = 1; a < nsh; ++a) { /* insertion sort: stable (Swift's sort is not; ties do not occur in real files) */
        int vi = shIdx[a]; long vk = shKey[a]; int b = a - 1;
        while (b >= 0 && shKey[b] > vk) { shIdx[b + 1] = shIdx[b]; shKey[b + 1] = shKey[b]; b--; }
        shIdx[b + 1] = vi; shKey[b + 1] = vk;
    }
    const unsigned char* body = data + h->bodyStart;
#define GET(v, idx) ((idx) >= 0 ? readProp(body + (size_t)(v) * stride + offsets[idx], vx->props[idx].type) : 0.0f)
    /* format detection on the first 100 vertices, PLYLoader.swift:620-650 */
    int scaleLog = 1, opLogit = 1;
    const uint32_t sampleCount = n < 100 ? n : 100;
    if (is0 >= 0 && sampleCount > 0) {
        int hasNeg = 0, hasLarge = 0;
        float sum = 0.0f;
        for (uint32_t v = 0; v < sampleCount; ++v) {
            float s = GET(v, is0);
            if (s < 0.0f) hasNeg = 1;
            if (s > 1.0f) hasLarge = 1;
            sum += s;
        }
        float avg = sum / (float)sampleCount;
        if (hasNeg) scaleLog = 1;
        else if (!hasLarge && avg > 0.0f && avg < 0.5f) scaleLog = 0;
    }
    if (iop >= 0 && sampleCount > 0) {
        float mn = GET(0, iop), mx = mn;
        for (uint32_t v = 0; v < sampleCount; ++v) { float o = GET(v, iop); if (o < mn) mn = o; if (o > mx) mx = o; }
        opLogit = (mn < 0.0f || mx > 1.0f);
    }
    res->scaleIsLogSpace = (uint32_t)scaleLog;
    res->opacityIsLogit = (uint32_t)opLogit;
    /* vertices, PLYLoader.swift:652-689 */
    uint32_t m = 0;
    float* shRaw = (float*)malloc(sizeof(float) * (size_t)(nsh > 0 ? nsh : 1) * (n > 0 ? n : 1));
    for (uint32_t v = 0; v < n; ++v) {
        float s0 = GET(v, is0), s1 = GET(v, is1), s2 = GET(v, is2), op = GET(v, iop);
        if (s0 == 2.0f && s1 == 2.0f && s2 == 2.0f && fabsf(op - 4.8402f) < 0.001f) continue; /* placeholder */
        pos[3 * m] = GET(v, ix); pos[3 * m + 1] = GET(v, iy); pos[3 * m + 2] = GET(v, iz);
        if (scaleLog) { scale[3 * m] = gsmo_exp(s0); scale[3 * m + 1] = gsmo_exp(s1); scale[3 * m + 2] = gsmo_exp(s2); }
        else { scale[3 * m] = s0; scale[3 * m + 1] = s1; scale[3 * m + 2] = s2; }
        float qx = GET(v, ir1), qy = GET(v, ir2), qz = GET(v, ir3), qw = GET(v, ir0);
        float inv = 1.0f / sqrtf(((qx * qx + qy * qy) + qz * qz) + qw * qw);
        rot[4 * m] = qx * inv; rot[4 * m + 1] = qy * inv; rot[4 * m + 2] = qz * inv; rot[4 * m + 3] = qw * inv;
        opacity[m] = opLogit ? 1.0f / (1.0f + gsmo_exp(-op)) : op;
        for (int k = 0; k < nsh; ++k) shRaw[(size_t)m * nsh + k] = GET(v, shIdx[k]);
        m++;
    }
#undef GET
    /* SH re-layout, PLYLoader.swift:692-719 */
    const uint32_t shComponents = nsh == 0 ? 0 : (uint32_t)nsh / 3;
    res->shComponents = shComponents;
    res->harmonicsStride = shComponents > 0 ? (uint32_t)nsh : 0;
    if (shComponents > 0) {
        if ((size_t)m * nsh > harmonicsCapacity) { free(shRaw); return GSMO_PLY_INSUFFICIENT_DATA; }
        const uint32_t hoc = shComponents - 1;
        memset(harmonics, 0, sizeof(float) * (size_t)m * nsh);
        for (uint32_t i = 0; i < m; ++i) {
            const float* src = shRaw + (size_t)i * nsh;
            float* dst = harmonics + (size_t)i * nsh;
            dst[0] = src[0];
            for (uint32_t c = 0; c < hoc; ++c) dst[1 + c] = src[3 + c];
            dst[shComponents] = src[1];
            for (uint32_t c = 0; c < hoc; ++c) dst[shComponents + 1 + c] = src[3 + hoc + c];
            dst[2 * shComponents] = src[2];
            for (uint32_t c = 0; c < hoc; ++c) dst[2 * shComponents + 1 + c] = src[3 + 2 * hoc + c];
        }
    }
    free(shRaw);
    res->count = m;
    recenter(pos, m, res);
    return GSMO_PLY_OK;
}

static float unpackUnorm(uint32_t v, int bits) { /* PLYLoader.swift:354-357 */
    uint32_t mask = (1u << bits) - 1u;
    return (float)(v & mask) / (float)mask;
}
static float lerpf(float a, float b, float t) { return a * (1.0f - t) + b * t; } /* PLYLoader.swift:401-403 */

static int loadCompressed(const unsigned char* data, size_t size, const gsmo_ply_header* h, float* pos, float* scale,
                          float* rot, float* opacity, float* harmonics, size_t harmonicsCapacity, gsmo_ply_result* res) {
    const gsmo_ply_elem* ch = findElem(h, "chunk");
    const gsmo_ply_elem* vx = findElem(h, "vertex");
    if (!ch || !vx) return GSMO_PLY_MISSING_CHUNK;
    int chunkStride = 0, vertexStride = 0, shStride = 0;
    for (int i = 0; i < ch->nprops; ++i) chunkStride += typeWidth(ch->props[i].type);
    for (int i = 0; i < vx->nprops; ++i) vertexStride += typeWidth(vx->props[i].type);
    const gsmo_ply_elem* sh = findElem(h, "sh");
    if (sh) for (int i = 0; i < sh->nprops; ++i) shStride += typeWidth(sh->props[i].type);
    const size_t chunkStart = h->bodyStart, vertexStart = chunkStart + (size_t)chunkStride * ch->count;
    const size_t shStart = vertexStart + (size_t)vertexStride * vx->count;
    if (size < shStart + (size_t)shStride * vx->count) return GSMO_PLY_INSUFFICIENT_DATA;
    const uint32_t n = vx->count;
    if ((size_t)n * 3 > harmonicsCapacity) return GSMO_PLY_INSUFFICIENT_DATA;
#define CHOFF(pname, var) int var = -1; { int o_ = 0; for (int i = 0; i < ch->nprops; ++i) { if (!strcmp(ch->props[i].name, pname)) var = o_; o_ += typeWidth(ch->props[i].type); } }
#define VXOFF(pname, var) int var = -1; { int o_ = 0; for (int i = 0; i < vx->nprops; ++i) { if (!strcmp(vx->props[i].name, pname)) var = o_; o_ += typeWidth(vx->props[i].type); } }
    CHOFF("min_x", oMinX) CHOFF("min_y", oMinY) CHOFF("min_z", oMinZ) CHOFF("max_x", oMaxX) CHOFF("max_y", oMaxY) CHOFF("max_z", oMaxZ)
    CHOFF("min_scale_x", oMinSX) CHOFF("min_scale_y", oMinSY) CHOFF("min_scale_z", oMinSZ)
    CHOFF("max_scale_x", oMaxSX) CHOFF("max_scale_y", oMaxSY) CHOFF("max_scale_z", oMaxSZ)
    CHOFF("min_r", oMinR) CHOFF("min_g", oMinG) CHOFF("min_b", oMinB) CHOFF("max_r", oMaxR) CHOFF("max_g", oMaxG) CHOFF("max_b", oMaxB)
    VXOFF("packed_position", oPos) VXOFF("packed_rotation", oRot) VXOFF("packed_scale", oScale) VXOFF("packed_color", oColor)
#undef CHOFF
#undef VXOFF
#define CF(c, off) ((off) >= 0 ? readProp(data + chunkStart + (size_t)(c) * chunkStride + (off), GSMO_T_F32) : 0.0f)
#define VU(v, off, out) do { out = 0; if ((off) >= 0) memcpy(&out, data + vertexStart + (size_t)(v) * vertexStride + (off), 4); } while (0)
    const float norm = 1.0f / (sqrtf(2.0f) * 0.5f); /* PLYLoader.swift:375 */
    const float SH_C0 = 0.28209479177387814f;
    for (uint32_t v = 0; v < n; ++v) {
        const uint32_t c = v / 256u;
        uint32_t pp, pr, ps, pc;
        VU(v, oPos, pp); VU(v, oRot, pr); VU(v, oScale, ps); VU(v, oColor, pc);
        float px = unpackUnorm(pp >> 21, 11), py = unpackUnorm(pp >> 11, 10), pz = unpackUnorm(pp, 11);
        pos[3 * v] = lerpf(CF(c, oMinX), CF(c, oMaxX), px);
        pos[3 * v + 1] = lerpf(CF(c, oMinY), CF(c, oMaxY), py);
        pos[3 * v + 2] = lerpf(CF(c, oMinZ), CF(c, oMaxZ), pz);
        float a = (unpackUnorm(pr >> 20, 10) - 0.5f) * norm, b = (unpackUnorm(pr >> 10, 10) - 0.5f) * norm,
              cc = (unpackUnorm(pr, 10) - 0.5f) * norm;
        float m = sqrtf(gsmo_fmax(0.0f, 1.0f - ((a * a + b * b) + cc * cc)));
        float qx, qy, qz, qw; /* PLYLoader.swift:392-398 */
        switch (pr >> 30) {
            case 0: qx = a; qy = b; qz = cc; qw = m; break;
            case 1: qx = m; qy = b; qz = cc; qw = a; break;
            case 2: qx = b; qy = m; qz = cc; qw = a; break;
            default: qx = b; qy = cc; qz = m; qw = a; break;
        }
        rot[4 * v] = qx; rot[4 * v + 1] = qy; rot[4 * v + 2] = qz; rot[4 * v + 3] = qw;
        float sx = unpackUnorm(ps >> 21, 11), sy = unpackUnorm(ps >> 11, 10), sz = unpackUnorm(ps, 11);
        scale[3 * v] = gsmo_exp(lerpf(CF(c, oMinSX), CF(c, oMaxSX), sx));
        scale[3 * v + 1] = gsmo_exp(lerpf(CF(c, oMinSY), CF(c, oMaxSY), sy));
        scale[3 * v + 2] = gsmo_exp(lerpf(CF(c, oMinSZ), CF(c, oMaxSZ), sz));
        float cr = unpackUnorm(pc >> 24, 8), cg = unpackUnorm(pc >> 16, 8), cb = unpackUnorm(pc >> 8, 8), ca = unpackUnorm(pc, 8);
        opacity[v] = ca;
        harmonics[3 * v] = (lerpf(CF(c, oMinR), CF(c, oMaxR), cr) - 0.5f) / SH_C0;
        harmonics[3 * v + 1] = (lerpf(CF(c, oMinG), CF(c, oMaxG), cg) - 0.5f) / SH_C0;
        harmonics[3 * v + 2] = (lerpf(CF(c, oMinB), CF(c, oMaxB), cb) - 0.5f) / SH_C0;
    }
#undef CF
#undef VU
    res->count = n;
    res->shComponents = 1;
    res->harmonicsStride = 3;
    res->scaleIsLogSpace = 1;
    res->opacityIsLogit = 0;
    recenter(pos, n, res);
    return GSMO_PLY_OK;
}

/* PLYLoader.load, PLYLoader.swift:246-281 */
int gsmo_ply_load(const unsigned char* data, size_t size, float* pos, float* scale, float* rot, float* opacity,
                  float* harmonics, size_t harmonicsCapacity, gsmo_ply_result* res) {
    gsmo_ply_header* h = (gsmo_ply_header*)malloc(sizeof(gsmo_ply_header));
    memset(res, 0, sizeof(*res));
    int rc = parseHeader(data, size, h);
    if (rc != GSMO_PLY_OK) { free(h); return rc; }
    if (h->format != 1) { free(h); return GSMO_PLY_UNSUPPORTED_FORMAT; }
    const gsmo_ply_elem* vx = findElem(h, "vertex");
    if (!vx) { free(h); return GSMO_PLY_MISSING_VERTEX; }
    const int compressed = findElem(h, "chunk") && hasProp(vx, "packed_position") && hasProp(vx, "packed_rotation") &&
                           hasProp(vx, "packed_scale") && hasProp(vx, "packed_color");
    res->compressed = (uint32_t)compressed;
    rc = compressed ? loadCompressed(data, size, h, pos, scale, rot, opacity, harmonics, harmonicsCapacity, res)
                    : loadStandard(data, size, h, vx, pos, scale, rot, opacity, harmonics, harmonicsCapacity, res);
    if (rc == GSMO_PLY_OK) gsmo_scene_bounds(pos, scale, res->count, res->boundsCenter, &res->boundsRadius);
    free(h);
    return rc;
}

/* Header-only query: declared vertex count and the number of SH-like properties (sizes the caller's buffers). */
int gsmo_ply_probe(const unsigned char* data, size_t size, uint32_t* vertexCount, uint32_t* shProps) {
    gsmo_ply_header* h = (gsmo_ply_header*)malloc(sizeof(gsmo_ply_header));
    int rc = parseHeader(data, size, h);
    *vertexCount = 0; *shProps = 0;
    if (rc == GSMO_PLY_OK) {
        const gsmo_ply_elem* vx = findElem(h, "vertex");
        if (!vx) rc = GSMO_PLY_MISSING_VERTEX;
        else {
            *vertexCount = vx->count;
            uint32_t k = 0;
            for (int i = 0; i < vx->nprops; ++i) {
                const char* nm = vx->props[i].name;
                if (lowerStarts(nm, "f_dc_") || lowerStarts(nm, "f_rest_") || lowerStarts(nm, "sh_") || lowerStarts(nm, "spherical_harmonics_")) k++;
            }
            *shProps = k < 3 ? 3 : k;
        }
    }
    free(h);
    return rc;
}

/* ---- Morton pre-sort, Scene.swift:47-138 */
static uint64_t expandBits(uint64_t v) { /* Scene.swift:50-58 */
    uint64_t x = v & 0x1FFFFFull;
    x = (x | (x << 32)) & 0x1F00000000FFFFull;
    x = (x | (x << 16)) & 0x1F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
uint64_t gsmo_morton_code(float x, float y, float z) { /* Scene.swift:62-69: UInt64(max(0, min(scale, v * scale))) truncates */
    const float s = 2097151.0f;
    uint64_t xi = (uint64_t)gsmo_fmax(0.0f, gsmo_fmin(s, x * s));
    uint64_t yi = (uint64_t)gsmo_fmax(0.0f, gsmo_fmin(s, y * s));
    uint64_t zi = (uint64_t)gsmo_fmax(0.0f, gsmo_fmin(s, z * s));
    return expandBits(xi) | (expandBits(yi) << 1) | (expandBits(zi) << 2);
}

/* codes[n] and order[n] out. Equal codes keep their input order (the reference's indices.sort is not guaranteed stable;
 * the stable order is the deterministic choice and what the device produces). */
void gsmo_morton_order(const float* pos, uint32_t n, uint64_t* codes, uint32_t* order) {
    if (n == 0) return;
    float mn[3] = {pos[0], pos[1], pos[2]}, mx[3] = {pos[0], pos[1], pos[2]};
    for (uint32_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            mn[k] = gsmo_fmin(mn[k], pos[3 * i + k]);
            mx[k] = gsmo_fmax(mx[k], pos[3 * i + k]);
        }
    float inv[3];
    for (int k = 0; k < 3; ++k) { float e = mx[k] - mn[k]; inv[k] = e > 1e-6f ? 1.0f / e : 0.0f; }
    for (uint32_t i = 0; i < n; ++i) {
        codes[i] = gsmo_morton_code((pos[3 * i] - mn[0]) * inv[0], (pos[3 * i + 1] - mn[1]) * inv[1], (pos[3 * i + 2] - mn[2]) * inv[2]);
        order[i] = i;
    }
    /* stable LSD radix sort of (code, index), 8 x 8-bit digits */
    uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * n);
    uint32_t* a = order;
    uint32_t* b = tmp;
    for (int p = 0; p < 8; ++p) {
        size_t hist[257] = {0};
        for (uint32_t i = 0; i < n; ++i) hist[((codes[a[i]] >> (8 * p)) & 0xFF) + 1]++;
        for (int d = 0; d < 256; ++d) hist[d + 1] += hist[d];
        for (uint32_t i = 0; i < n; ++i) b[hist[(codes[a[i]] >> (8 * p)) & 0xFF]++] = a[i];
        uint32_t* t = a; a = b; b = t;
    }
    /* 8 passes: the result is back in `order` */
    free(tmp);
}

/* ---- packing, PLYBenchmarkTests.swift:139-149 / TestUtils.swift:245: records -> PackedWorldGaussian(+Half); harmonics
 * -> float or Float16 (round to nearest even). `order` (optional) applies a permutation: out[i] = in[order[i]]. */
void gsmo_pack_gaussians(const float* pos, const float* scale, const float* rot, const float* opacity, uint32_t n,
                         const uint32_t* order, int half, void* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t s = order ? order[i] : i;
        if (half) {
            unsigned char* o = (unsigned char*)out + (size_t)i * 32;
            memcpy(o, pos + 3 * s, 12);
            gsmo_half hv[10] = {gsmo_f2h(opacity[s]), gsmo_f2h(scale[3 * s]), gsmo_f2h(scale[3 * s + 1]), gsmo_f2h(scale[3 * s + 2]),
                                gsmo_f2h(rot[4 * s]), gsmo_f2h(rot[4 * s + 1]), gsmo_f2h(rot[4 * s + 2]), gsmo_f2h(rot[4 * s + 3]), 0, 0};
            memcpy(o + 12, hv, 20);
        } else {
            float* o = (float*)((unsigned char*)out + (size_t)i * 48);
            /* BridgingTypes.h:58-64: position, opacity, scale, pad, rotation (x, y, z, w) */
            o[0] = pos[3 * s]; o[1] = pos[3 * s + 1]; o[2] = pos[3 * s + 2];
            o[3] = opacity[s];
            o[4] = scale[3 * s]; o[5] = scale[3 * s + 1]; o[6] = scale[3 * s + 2];
            o[7] = 0.0f;
            o[8] = rot[4 * s]; o[9] = rot[4 * s + 1]; o[10] = rot[4 * s + 2]; o[11] = rot[4 * s + 3];
        }
    }
}
void gsmo_pack_harmonics(const float* harmonics, uint32_t n, uint32_t stride, const uint32_t* order, int half, void* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t s = order ? order[i] : i;
        for (uint32_t c = 0; c < stride; ++c) {
            if (half) ((gsmo_half*)out)[(size_t)i * stride + c] = gsmo_f2h(harmonics[(size_t)s * stride + c]);
            else ((float*)out)[(size_t)i * stride + c] = harmonics[(size_t)s * stride + c];
        }
    }
}
