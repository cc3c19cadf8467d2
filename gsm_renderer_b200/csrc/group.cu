// group.cu -- the exchange step of the strip-sharded frame, fused with its producer and its consumer
// over NVLink peer memory instead of a library collective. No reference counterpart (the reference is single-device).
//
// Every rank owns an exchange WINDOW in its HBM, mapped into every peer (cudaIpc across processes, plain pointers for
// emulated ranks in one process): [mailbox | one receive region per source rank | optional frame image]. A frame is
//   route_records_kernel  (source rank)  walks the rank's compacted, projected splats in gid order, finds the strips each
//                         one touches (tile rows of its AABB, refined by stage 1's hit mask) and STORES the 48-byte record
//                         straight into that destination's region for this source, at its order-preserving position (a
//                         single-pass multi-destination compaction: packed per-destination counts, block scan, chained
//                         look-back over tiles). The kernel's last CTA release-stores (count, frame sequence) into every
//                         destination's mailbox. The transfer IS the kernel's stores: it overlaps the routing math tile by tile.
//   ingest_routed_kernel  (destination rank) acquires the mailbox words of all sources, then reads only the records that
//                         touch ITS strip -- sources in rank order == ascending global gid, so the stable depth sort breaks
//                         ties exactly as the single-GPU frame does -- and its last CTA acks the sources (flow control for the
//                         single-buffered regions: a source starts routing frame f+1 only when every destination has consumed f).
// Nothing returns to the host: counts stay on the device, there is no stream synchronisation and no staging copy.
// (Round 1 did this with a padded NCCL all-gather of every record to every rank plus host-side count exchange:
// 205 MB gathered and ingested per rank at C3; here a rank receives ~1/world of the records.)
#include "gsm_common.cuh"
#include "gsm_kernels.h"
#include "gsm_tiletest.cuh"

namespace gsm {

constexpr int kRouteItems = 2;
constexpr uint32_t kRouteTile = 256u * kRouteItems;
constexpr uint32_t kRouteValueMask = 0x3FFFFFFFu, kRouteAggregate = 0x40000000u, kRouteInclusive = 0x80000000u;

uint32_t routeStatusWords(uint32_t maxRecords) { return ((maxRecords + kRouteTile - 1u) / kRouteTile + 1u) * kGroupMaxRanks; }

__device__ __forceinline__ uint32_t field16(unsigned long long lo, unsigned long long hi, uint32_t d) {
    return (uint32_t)(((d < 4u ? lo : hi) >> (16u * (d & 3u))) & 0xFFFFu);
}

// Destinations of one record: bit d set when strip d (tile rows [rowStart[d], rowStart[d+1])) holds at least one of the
// splat's tiles. For AABBs of at most kMaskTiles tiles stage 1's hit mask decides exactly (rows are contiguous bit ranges of
// the row-major mask); larger AABBs go to every strip their rows meet and the destination re-runs the exact walk.
__device__ __forceinline__ uint32_t recordDestinations(const int4 b, uint32_t maskLo, uint32_t maskHi, const RouteParams& P) {
    const int minTX = b.x, maxTX = b.y, minTY = b.z, maxTY = b.w;
    if (minTX > maxTX || minTY > maxTY) return 0u;
    const uint32_t w = (uint32_t)(maxTX - minTX + 1);
    const uint32_t fullTiles = w * (uint32_t)(maxTY - minTY + 1);
    const unsigned long long mask = ((unsigned long long)maskHi << 32) | maskLo;
    uint32_t dests = 0u;
    for (uint32_t d = 0; d < P.world; ++d) {
        const int lo = max(minTY, (int)P.rowStart[d]), hi = min(maxTY, (int)P.rowStart[d + 1] - 1);
        if (lo > hi) continue;
        if (fullTiles <= kMaskTiles) {
            const uint32_t shift = (uint32_t)(lo - minTY) * w, bits = (uint32_t)(hi - lo + 1) * w;
            unsigned long long m = mask >> shift;
            if (bits < 64u) m &= (1ull << bits) - 1ull;
            if (m == 0ull) continue;
        }
        dests |= 1u << d;
    }
    return dests;
}

__global__ void __launch_bounds__(256) route_records_kernel(FrameState* fs, const uint32_t* __restrict__ keys,
                                                            const int32_t* __restrict__ gids, const void* __restrict__ renderData,
                                                            const int32_t* __restrict__ bounds, const uint2* __restrict__ hitMask,
                                                            uint32_t cap, uint32_t* status, const __grid_constant__ RouteParams P) {
    __shared__ uint32_t s_tile, s_last;
    __shared__ unsigned long long s_warp[8][2];
    __shared__ uint32_t s_base[kGroupMaxRanks], s_cnt[kGroupMaxRanks];
    // The tile's records are staged once in shared memory, with one index list per destination: the copy-out then writes each
    // destination's run as consecutive 16-byte chunks by consecutive threads (512 contiguous bytes per warp store). Storing the
    // 48-byte records straight from registers (3 x 16 B at a 48-byte stride per warp instruction) made every NVLink write a
    // partial sector: 2-GPU C3 spent ~0.4 ms in this kernel against ~0.1 ms with local stores.
    __shared__ SplatRecord s_rec[kRouteTile];
    __shared__ uint16_t s_list[kGroupMaxRanks][kRouteTile];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    pdlLaunchDependents();
    pdlWait();
    // flow control: every destination has consumed this rank's records of the previous frame (regions are single-buffered)
    if (tid < P.world) {
        while ((int32_t)(ld_acquire_sys(&P.mine->ack[tid]) - (P.seq - 1u)) < 0) __nanosleep(200);
    }
    __syncthreads();
    const uint32_t count = min(fs->visibleCountRaw, cap);
    const uint32_t numTiles = (count + kRouteTile - 1u) / kRouteTile;
    while (true) {
        if (tid == 0) s_tile = atomicAdd(&fs->ticketRoute, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= numTiles) break;
        const uint32_t first = tile * kRouteTile + tid * kRouteItems;  // blocked: index order == gid order
        uint32_t dests[kRouteItems];
        unsigned long long c0 = 0ull, c1 = 0ull;  // per-destination counts of this thread, 16-bit fields (<= kRouteItems each)
#pragma unroll
        for (int i = 0; i < kRouteItems; ++i) {
            const uint32_t j = first + i;
            dests[i] = 0u;
            if (j < count) {
                const uint32_t gid = (uint32_t)gids[j];
                SplatRecord rec;
                rec.renderData = __ldg(reinterpret_cast<const uint4*>(renderData) + gid);
                rec.bounds = __ldg(reinterpret_cast<const int4*>(bounds) + gid);
                const uint2 m = __ldg(hitMask + gid);
                rec.key = keys[j];
                rec.maskLo = m.x; rec.maskHi = m.y;
                rec.gid = gid;
                dests[i] = recordDestinations(rec.bounds, m.x, m.y, P);
                for (uint32_t d = 0; d < P.world; ++d)
                    if (dests[i] >> d & 1u) { if (d < 4u) c0 += 1ull << (16u * d); else c1 += 1ull << (16u * (d - 4u)); }
                uint4* sd = reinterpret_cast<uint4*>(&s_rec[tid * kRouteItems + i]);
                const uint4* sr = reinterpret_cast<const uint4*>(&rec);
                sd[0] = sr[0]; sd[1] = sr[1]; sd[2] = sr[2];
            }
        }
        // block exclusive scan of the packed counts (a tile holds 512 records: no field overflows)
        unsigned long long i0 = c0, i1 = c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t0 = __shfl_up_sync(0xFFFFFFFFu, i0, o), t1 = __shfl_up_sync(0xFFFFFFFFu, i1, o);
            if (lane >= (unsigned)o) { i0 += t0; i1 += t1; }
        }
        if (lane == 31u) { s_warp[warp][0] = i0; s_warp[warp][1] = i1; }
        __syncthreads();
        unsigned long long w0 = 0ull, w1 = 0ull, tot0 = 0ull, tot1 = 0ull;
#pragma unroll
        for (unsigned w = 0; w < 8u; ++w) {
            const unsigned long long a = s_warp[w][0], b = s_warp[w][1];
            if (w < warp) { w0 += a; w1 += b; }
            tot0 += a; tot1 += b;
        }
        unsigned long long e0 = w0 + i0 - c0, e1 = w1 + i1 - c1;  // exclusive prefix of this thread inside the tile
        // per-destination index lists, in record order
#pragma unroll
        for (int i = 0; i < kRouteItems; ++i) {
            uint32_t dm = dests[i];
            while (dm) {
                const uint32_t d = (uint32_t)__ffs(dm) - 1u;
                dm &= dm - 1u;
                s_list[d][field16(e0, e1, d)] = (uint16_t)(tid * kRouteItems + i);
                if (d < 4u) e0 += 1ull << (16u * d); else e1 += 1ull << (16u * (d - 4u));
            }
        }
        // chained look-back over tiles, one thread per destination
        if (tid < P.world) {
            const uint32_t mineCount = field16(tot0, tot1, tid);
            uint32_t* st = status + (size_t)tile * kGroupMaxRanks + tid;
            uint32_t exclusive = 0u;
            if (tile == 0u) {
                st_status32(st, kRouteInclusive | mineCount);
            } else {
                st_status32(st, kRouteAggregate | mineCount);
                int look = (int)tile - 1;
                while (look >= 0) {
                    const uint32_t sw = ld_status32(status + (size_t)look * kGroupMaxRanks + tid);
                    if (sw & kRouteInclusive) { exclusive += sw & kRouteValueMask; break; }
                    if (sw & kRouteAggregate) { exclusive += sw & kRouteValueMask; --look; }
                }
                st_status32(st, kRouteInclusive | (exclusive + mineCount));
            }
            s_base[tid] = exclusive;
            s_cnt[tid] = mineCount;
            if (tile == numTiles - 1u) fs->routeTotals[tid] = exclusive + mineCount;
        }
        __syncthreads();
        // copy-out: destination d's run is region[d][s_base[d] .. + s_cnt[d]) -- peer memory for d != rank (NVLink stores)
        for (uint32_t d = 0; d < P.world; ++d) {
            const uint32_t n = s_cnt[d], base = s_base[d];
            if (n == 0u) continue;
            const uint32_t room = base < P.regionCap ? min(n, P.regionCap - base) : 0u;  // a source never routes more records than its shard holds
            uint4* dst = reinterpret_cast<uint4*>(P.region[d] + base);
            for (uint32_t q = tid; q < room * 3u; q += 256u) {
                const uint32_t e = q / 3u, c = q - e * 3u;
                dst[q] = reinterpret_cast<const uint4*>(&s_rec[s_list[d][e]])[c];
            }
        }
        __syncthreads();
    }
    // all of this CTA's record stores are performed system-wide before it arrives; the last CTA to arrive publishes
    __threadfence_system();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&fs->routeDone, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (s_last && tid < P.world) {
        __threadfence();
        const uint32_t total = ld_status32(&fs->routeTotals[tid]);
        GroupMailbox* mb = P.mailbox[tid];
        st_relaxed_sys(&mb->recordCount[P.rank], total);
        __threadfence_system();
        st_release_sys(&mb->recordSeq[P.rank], P.seq);
    }
}

// The destination side: the body is ingest_records_kernel's (strip.cu) over the virtual concatenation, in source order, of
// the regions of this rank's window; the record count exists only on the device, so the grid is persistent.
__global__ void __launch_bounds__(256) ingest_routed_kernel(const __grid_constant__ IngestParams P, ProjectOut o,
                                                            uint32_t* __restrict__ recTouched, uint32_t* __restrict__ recKey,
                                                            uint32_t* __restrict__ recGid) {
    __shared__ uint32_t s_prefix[kGroupMaxRanks + 1];
    __shared__ uint32_t s_last;
    __shared__ WarpTileWork s_work[8];
    const unsigned tid = threadIdx.x;
    pdlLaunchDependents();
    pdlWait();
    if (tid == 0) {
        uint32_t acc = 0u;
        for (uint32_t s = 0; s < P.world; ++s) {
            while ((int32_t)(ld_acquire_sys(&P.mine->recordSeq[s]) - P.seq) < 0) __nanosleep(200);
            s_prefix[s] = acc;
            acc += min(ld_acquire_sys(&P.mine->recordCount[s]), P.regionCap);
        }
        s_prefix[P.world] = acc;
        if (blockIdx.x == 0) o.fs->recordTotal = min(acc, o.maxOut);
    }
    __syncthreads();
    const uint32_t total = min(s_prefix[P.world], o.maxOut);
    const int rowFirst = P.rowFirst, rowLast = P.rowLast;
    for (uint32_t jb = blockIdx.x * 256u; jb < total; jb += gridDim.x * 256u) {
        const uint32_t j = jb + tid;
        if ((j & ~31u) >= total) continue;  // whole warp past the end
        const bool inRange = j < total;
        uint32_t gid = 0, nTiles = 0, keyIn = 0xFFFFFFFFu, cnt = 0;
        uint4 rd = make_uint4(0, 0, 0, 0);
        int minTX = 0, maxTX = -1, minTY = 0, maxTY = -1;
        uint2 mask = make_uint2(0u, 0u);
        QuantSplat q = {};
        if (inRange) {
            uint32_t src = 0;
            for (uint32_t s = 1; s < P.world; ++s) src += (j >= s_prefix[s]) ? 1u : 0u;
            const uint4* rp = reinterpret_cast<const uint4*>(P.region[src] + (j - s_prefix[src]));
            rd = __ldcg(rp);  // L2 loads: the region is rewritten by peers every frame (never the read-only / L1 path)
            const uint4 bw = __ldcg(rp + 1);
            const uint4 kw = __ldcg(rp + 2);
            gid = kw.w;
            keyIn = kw.x;
            minTX = (int)bw.x; maxTX = (int)bw.y; minTY = max((int)bw.z, rowFirst); maxTY = min((int)bw.w, rowLast);
            if (minTX <= maxTX && minTY <= maxTY) {
                const uint32_t w = (uint32_t)(maxTX - minTX + 1);
                const uint32_t fullTiles = w * (uint32_t)((int)bw.w - (int)bw.z + 1);
                if (fullTiles <= kMaskTiles) {
                    const uint32_t shift = (uint32_t)(minTY - (int)bw.z) * w, bits = (uint32_t)(maxTY - minTY + 1) * w;
                    unsigned long long m = (((unsigned long long)kw.z << 32) | kw.y) >> shift;
                    if (bits < 64u) m &= (1ull << bits) - 1ull;
                    mask = make_uint2((uint32_t)m, (uint32_t)(m >> 32));
                    cnt = (uint32_t)__popcll(m);
                } else {
                    nTiles = w * (uint32_t)(maxTY - minTY + 1);
                }
            }
            if (cnt > 0 || nTiles > 0)
                q = makeQuantSplat(__ushort_as_half((unsigned short)(rd.x & 0xFFFFu)), __ushort_as_half((unsigned short)(rd.x >> 16)),
                                   (uint16_t)(rd.y & 0xFFFFu), __ushort_as_half((unsigned short)(rd.y >> 16)),
                                   __ushort_as_half((unsigned short)(rd.z & 0xFFFFu)), (uint8_t)(rd.w >> 24));
            if (nTiles > 0 && !(q.d2Cutoff >= 0.0f)) nTiles = 0;
        }
        uint2 walkMask;
        const uint32_t walkCnt = warpCountTiles(s_work[tid >> 5], nTiles, q, minTX, minTY, maxTX - minTX + 1, walkMask);
        if (nTiles > 0) { cnt = walkCnt; mask = walkMask; }
        if (inRange) {
            if (cnt > 0) {
                reinterpret_cast<uint4*>(o.renderData)[gid] = rd;
                reinterpret_cast<int4*>(o.bounds)[gid] = make_int4(minTX, maxTX, minTY, maxTY);
                o.nTouched[gid] = cnt;
                o.hitMask[gid] = mask;
                storeBlendSplat(o.blendSplats + gid, q, __ushort_as_half((unsigned short)(rd.x & 0xFFFFu)),
                                __ushort_as_half((unsigned short)(rd.x >> 16)), (uint8_t)rd.w, (uint8_t)(rd.w >> 8), (uint8_t)(rd.w >> 16),
                                (uint8_t)(rd.w >> 24), __ushort_as_half((unsigned short)(rd.z >> 16)));
            }
            recTouched[j] = cnt;
            recKey[j] = keyIn;
            recGid[j] = gid;
        }
    }
    // ack the sources once every CTA has read its records: their next frame may overwrite the regions
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&o.fs->ingestDone, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (s_last && tid < P.world) {
        __threadfence();
        st_release_sys(&P.mailbox[tid]->ack[P.rank], P.seq);
    }
}

// "My part of frame `seq` is in rank `to`'s image": runs after the blend in stream order (plain launch: the blend's peer
// stores have completed), fences system-wide and release-stores the flag.
__global__ void group_signal_kernel(GroupMailbox* to, uint32_t rank, uint32_t seq) {
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(&to->frameDone[rank], seq);
    }
}

// The owner of the image waits (on the device) for the parts of the ranks in `mask`.
__global__ void group_wait_kernel(const GroupMailbox* mine, uint32_t mask, uint32_t seq) {
    const uint32_t s = threadIdx.x;
    if (s < kGroupMaxRanks && (mask >> s & 1u)) {
        while ((int32_t)(ld_acquire_sys(&mine->frameDone[s]) - seq) < 0) __nanosleep(200);
    }
}

cudaError_t launchRouteRecords(cudaStream_t s, FrameState* fs, const uint32_t* keys, const int32_t* gids, const void* renderData,
                               const int32_t* bounds, const uint2* hitMask, uint32_t cap, uint32_t* status, const RouteParams& P,
                               int numSMs) {
    uint32_t tiles = (cap + kRouteTile - 1u) / kRouteTile;
    uint32_t grid = tiles < (uint32_t)numSMs * 4u ? tiles : (uint32_t)numSMs * 4u;  // co-resident: the look-back chain needs it
    if (grid == 0) grid = 1;
    launchChained(route_records_kernel, grid, 256, s, fs, keys, gids, renderData, bounds, hitMask, cap, status, P);
    return cudaGetLastError();
}

cudaError_t launchIngestRouted(cudaStream_t s, const IngestParams& P, const ProjectOut& o, uint32_t* recTouched, uint32_t* recKey,
                               uint32_t* recGid, int numSMs) {
    launchChained(ingest_routed_kernel, (uint32_t)numSMs * 6u, 256, s, P, o, recTouched, recKey, recGid);
    return cudaGetLastError();
}

cudaError_t launchGroupSignal(cudaStream_t s, GroupMailbox* to, uint32_t rank, uint32_t seq) {
    group_signal_kernel<<<1, 32, 0, s>>>(to, rank, seq);
    return cudaGetLastError();
}

cudaError_t launchGroupWait(cudaStream_t s, const GroupMailbox* mine, uint32_t mask, uint32_t seq) {
    group_wait_kernel<<<1, 32, 0, s>>>(mine, mask, seq);
    return cudaGetLastError();
}

}  // namespace gsm
