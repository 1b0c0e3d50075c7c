// fovgs_sort.cu — level 2 of the binning sort: every tile's segment of (depth_bits << 32 | id) keys is sorted by a
// block-local LSD radix sort over the 32 depth bits (digits that are uniform inside the segment are skipped);
// runs of equal depth bits are then ordered by id.  The result equals a stable sort on (tile, depth_bits) of the
// id-ascending emission order, i.e. exactly the reference's point_list
// (cub::DeviceRadixSort::SortPairs on 32+13 bits, FOV/cuda_rasterizer/rasterizer_impl.cu:843-854),
// with ~16 B/instance of HBM traffic instead of six 24 B/instance global passes.
//
// Segments are classed by size so that keys live in shared memory whenever they fit:
//   n <= 256          rank sort (one pass, full 64-bit compare)
//   n <= CAP          radix in shared memory, CAP in {2048, 6144, 12288}  (41 / 106 / 204 KB per CTA)
//   n >  12288        radix on the global ping-pong buffers
#include "fovgs_internal.cuh"

namespace fovgs {

template <int CAP>
__device__ __forceinline__ void sort_tile_smem(const Workspace& ws, const int tile, uint32_t* __restrict__ out_ranges,
                                               uint32_t* __restrict__ out_point_list, uint32_t nmin) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);          // [2][CAP]
    uint32_t* whist = reinterpret_cast<uint32_t*>(keys + 2 * CAP);   // [8][256]
    __shared__ uint32_t totals[256];
    __shared__ uint32_t wsum[8];
    __shared__ unsigned long long vary_s;
    const uint32_t cap = ws.hdr->cap;
    uint32_t sbeg = ws.tile_offset[tile], send = ws.tile_offset[tile + 1];
    if (out_ranges && threadIdx.x == 0) {
        // reference semantics: untouched tiles keep the memset value (0,0)  (rasterizer_impl.cu:860)
        out_ranges[2 * tile] = (send > sbeg) ? sbeg : 0u;
        out_ranges[2 * tile + 1] = (send > sbeg) ? send : 0u;
    }
    sbeg = min(sbeg, cap);
    send = min(send, cap);
    const uint32_t n = send - sbeg;
    if (n <= nmin || n > (uint32_t)CAP) return;
    const uint64_t* __restrict__ gsrc = ws.keysA + sbeg;
    uint32_t* out = ws.point_list + sbeg;
    uint32_t* out2 = out_point_list ? out_point_list + sbeg : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n == 1) {
        if (tid == 0) {
            const uint32_t id = (uint32_t)gsrc[0];
            out[0] = id;
            if (out2) out2[0] = id;
        }
        return;
    }
    if (tid == 0) vary_s = 0ull;
    __syncthreads();
    {
        const uint64_t k0 = gsrc[0];
        uint64_t v = 0;
        for (uint32_t i = tid; i < n; i += 256) { const uint64_t k = gsrc[i]; keys[i] = k; v |= (k ^ k0); }
#pragma unroll
        for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicOr(&vary_s, (unsigned long long)v);
    }
    __syncthreads();
    if (n <= 256) {
        // rank sort: position = number of keys smaller than mine (keys are unique: the id is part of the key)
        if ((uint32_t)tid < n) {
            const uint64_t k = keys[tid];
            uint32_t r = 0;
            for (uint32_t j = 0; j < n; j++) r += keys[j] < k;
            const uint32_t id = (uint32_t)k;
            out[r] = id;
            if (out2) out2[r] = id;
        }
        return;
    }
    const uint64_t vary = vary_s;
    const uint32_t chunk = ((n + 7) / 8 + 31) & ~31u;
    const uint32_t wbeg = min(n, warp * chunk), wend = min(n, wbeg + chunk);
    int cur = 0;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 32 + 8 * pass;
        if (((vary >> shift) & 0xffull) == 0) continue;
        uint64_t* src = keys + cur * CAP;
        uint64_t* dst = keys + (cur ^ 1) * CAP;
        for (int i = tid; i < 8 * 256; i += 256) whist[i] = 0;
        __syncthreads();
        for (uint32_t i = wbeg + lane; i < wend; i += 32) atomicAdd(&whist[warp * 256 + ((src[i] >> shift) & 0xff)], 1u);
        __syncthreads();
        {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) { const uint32_t c = whist[w * 256 + tid]; whist[w * 256 + tid] = t; t += c; }
            uint32_t x = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) wsum[warp] = x;
            __syncthreads();
            uint32_t base = x - t;
            for (int w = 0; w < warp; w++) base += wsum[w];
            totals[tid] = base;
        }
        __syncthreads();
        for (int i = tid; i < 8 * 256; i += 256) whist[i] += totals[i & 255];
        __syncthreads();
        for (uint32_t i0 = wbeg; i0 < wend; i0 += 32) {
            const uint32_t i = i0 + lane;
            const bool valid = i < wend;
            const uint64_t key = valid ? src[i] : 0ull;
            const uint32_t d = valid ? (uint32_t)((key >> shift) & 0xff) : (256u + lane);
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t base = 0;
            if (valid) base = whist[warp * 256 + d];
            __syncwarp();
            if (valid) {
                dst[base + rank] = key;
                if (rank == 0) whist[warp * 256 + d] = base + __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        cur ^= 1;
    }
    uint64_t* src = keys + cur * CAP;
    for (uint32_t i = tid; i < n; i += 256) {
        const uint32_t dk = (uint32_t)(src[i] >> 32);
        const bool head = (i == 0) || ((uint32_t)(src[i - 1] >> 32) != dk);
        if (head && i + 1 < n && (uint32_t)(src[i + 1] >> 32) == dk) {
            uint32_t j = i + 1;
            while (j < n && (uint32_t)(src[j] >> 32) == dk) j++;
            for (uint32_t a = i + 1; a < j; a++) {
                const uint64_t k = src[a];
                uint32_t b = a;
                while (b > i && src[b - 1] > k) { src[b] = src[b - 1]; b--; }
                src[b] = k;
            }
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += 256) {
        const uint32_t id = (uint32_t)src[i];
        out[i] = id;
        if (out2) out2[i] = id;
    }
}

// size classes are powers of two and tile_order is descending by class, so every tile with more than `nmin` instances
// sits in the first cum_class[class(nmin+1)] entries: CTAs loop over that prefix only
__device__ __forceinline__ uint32_t heavy_prefix(const Workspace& ws, uint32_t nmin) {
    return nmin == 0 ? (uint32_t)ws.hdr->tiles : ws.hdr->cum_class[32 - __clz(nmin + 1)];
}

template <int CAP>
__global__ void __launch_bounds__(256) k_tile_sort_smem(Workspace ws, uint32_t* __restrict__ out_ranges,
                                                        uint32_t* __restrict__ out_point_list, uint32_t nmin) {
    const uint32_t limit = heavy_prefix(ws, nmin);
    for (uint32_t bi = blockIdx.x; bi < limit; bi += gridDim.x) {
        sort_tile_smem<CAP>(ws, (int)ws.tile_order[bi], out_ranges, out_point_list, nmin);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Segments that do not fit in shared memory: same algorithm on the global ping-pong buffers (L2-resident), with
// 32 warps per CTA and four keys in flight per lane so that the L2 round trips overlap instead of serialising.
// ------------------------------------------------------------------------------------------------------------------
constexpr int GW = 32;   // warps per CTA
constexpr int GU = 4;    // keys in flight per lane

__device__ __forceinline__ void sort_tile_global(const Workspace& ws, const int tile, uint32_t* __restrict__ out_point_list, uint32_t nmin) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* whist = reinterpret_cast<uint32_t*>(smem_raw);   // [GW][256]
    __shared__ uint32_t totals[256];
    __shared__ uint32_t wsum[GW];
    __shared__ unsigned long long vary_s;
    const uint32_t cap = ws.hdr->cap;
    uint32_t sbeg = ws.tile_offset[tile], send = ws.tile_offset[tile + 1];
    sbeg = min(sbeg, cap);
    send = min(send, cap);
    const uint32_t n = send - sbeg;
    if (n <= nmin) return;   // smaller segments are sorted in shared memory (k_tile_sort_smem)
    uint64_t* src = ws.keysA + sbeg;
    uint64_t* dst = ws.keysB + sbeg;
    uint32_t* out = ws.point_list + sbeg;
    uint32_t* out2 = out_point_list ? out_point_list + sbeg : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NT = GW * 32;
    if (tid == 0) vary_s = 0ull;
    __syncthreads();
    {
        const uint64_t k0 = src[0];
        uint64_t v = 0;
        for (uint32_t i = tid; i < n; i += NT) v |= (src[i] ^ k0);
#pragma unroll
        for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicOr(&vary_s, (unsigned long long)v);
    }
    __syncthreads();
    const uint64_t vary = vary_s;
    const uint32_t chunk = ((n + GW - 1) / GW + 31) & ~31u;
    const uint32_t wbeg = min(n, warp * chunk), wend = min(n, wbeg + chunk);
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 32 + 8 * pass;
        if (((vary >> shift) & 0xffull) == 0) continue;
        for (int i = tid; i < GW * 256; i += NT) whist[i] = 0;
        __syncthreads();
        for (uint32_t i0 = wbeg; i0 < wend; i0 += 32 * GU) {
            uint64_t k[GU];
#pragma unroll
            for (int u = 0; u < GU; u++) { const uint32_t i = i0 + u * 32 + lane; k[u] = (i < wend) ? src[i] : 0ull; }
#pragma unroll
            for (int u = 0; u < GU; u++) { const uint32_t i = i0 + u * 32 + lane; if (i < wend) atomicAdd(&whist[warp * 256 + ((k[u] >> shift) & 0xff)], 1u); }
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t t = 0;
            for (int w = 0; w < GW; w++) { const uint32_t c = whist[w * 256 + tid]; whist[w * 256 + tid] = t; t += c; }
            uint32_t x = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) wsum[warp] = x;
            totals[tid] = x - t;   // exclusive within the warp, fixed up below
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t base = totals[tid];
            for (int w = 0; w < warp; w++) base += wsum[w];
            totals[tid] = base;
        }
        __syncthreads();
        for (int i = tid; i < GW * 256; i += NT) whist[i] += totals[i & 255];
        __syncthreads();
        for (uint32_t i0 = wbeg; i0 < wend; i0 += 32 * GU) {
            uint64_t k[GU];
#pragma unroll
            for (int u = 0; u < GU; u++) { const uint32_t i = i0 + u * 32 + lane; k[u] = (i < wend) ? src[i] : 0ull; }
#pragma unroll
            for (int u = 0; u < GU; u++) {
                const uint32_t i = i0 + u * 32 + lane;
                const bool valid = i < wend;
                const uint32_t d = valid ? (uint32_t)((k[u] >> shift) & 0xff) : (256u + lane);
                const unsigned peers = __match_any_sync(0xffffffffu, d);
                const unsigned rank = __popc(peers & ((1u << lane) - 1u));
                uint32_t base = 0;
                if (valid) base = whist[warp * 256 + d];
                __syncwarp();
                if (valid) {
                    dst[base + rank] = k[u];
                    if (rank == 0) whist[warp * 256 + d] = base + __popc(peers);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        uint64_t* t = src; src = dst; dst = t;
    }
    for (uint32_t i = tid; i < n; i += NT) {
        const uint32_t dk = (uint32_t)(src[i] >> 32);
        const bool head = (i == 0) || ((uint32_t)(src[i - 1] >> 32) != dk);
        if (head && i + 1 < n && (uint32_t)(src[i + 1] >> 32) == dk) {
            uint32_t j = i + 1;
            while (j < n && (uint32_t)(src[j] >> 32) == dk) j++;
            for (uint32_t a = i + 1; a < j; a++) {
                const uint64_t k = src[a];
                uint32_t b = a;
                while (b > i && src[b - 1] > k) { src[b] = src[b - 1]; b--; }
                src[b] = k;
            }
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += NT) {
        const uint32_t id = (uint32_t)src[i];
        out[i] = id;
        if (out2) out2[i] = id;
    }
}

__global__ void __launch_bounds__(GW * 32) k_tile_sort_global(Workspace ws, uint32_t* __restrict__ out_point_list, uint32_t nmin) {
    const uint32_t limit = heavy_prefix(ws, nmin);
    for (uint32_t bi = blockIdx.x; bi < limit; bi += gridDim.x) {
        sort_tile_global(ws, (int)ws.tile_order[bi], out_point_list, nmin);
        __syncthreads();
    }
}

template <int CAP>
static cudaError_t launch_class(const Workspace& ws, int grid, uint32_t* out_ranges, uint32_t* out_point_list, uint32_t nmin,
                                cudaStream_t st) {
    const size_t smem = (size_t)2 * CAP * 8 + 8 * 256 * 4;
    static PerDeviceOnce once;       // per device (and per CAP instantiation)
    bool* configured = once.slot();
    if (!*configured) {
        cudaError_t e = cudaFuncSetAttribute(k_tile_sort_smem<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        *configured = true;
    }
    k_tile_sort_smem<CAP><<<grid, 256, smem, st>>>(ws, out_ranges, out_point_list, nmin);
    return cudaGetLastError();
}

cudaError_t launch_tile_sort(const Workspace& ws, int T, uint32_t* out_ranges, uint32_t* out_point_list, cudaStream_t st) {
    cudaError_t e;
    // the first class also publishes `ranges` for every tile
    // heaviest classes first: few long-running CTAs start while the bulk class fills the rest of the machine
    k_tile_sort_global<<<min(T, 148), GW * 32, GW * 256 * 4, st>>>(ws, out_point_list, 12288u);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = launch_class<12288>(ws, min(T, 148), nullptr, out_point_list, 6144u, st)) != cudaSuccess) return e;
    if ((e = launch_class<6144>(ws, min(T, 296), nullptr, out_point_list, 2048u, st)) != cudaSuccess) return e;
    // the bulk class also publishes `ranges` for every tile
    return launch_class<2048>(ws, T, out_ranges, out_point_list, 0u, st);
}

}  // namespace fovgs
