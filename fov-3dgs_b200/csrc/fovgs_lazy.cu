// fovgs_lazy.cu — fused "sort only what you composite" kernel: all variants (FOV, SMFR, MMFR, OBB, training family).
//
// Measurement that motivates it (6 M Gaussians, 1080p, gaze at the dense centre): the blend stage consumes 1.7 M of the
// 16.9 M binned instances (10 %) before every pixel of its tile has saturated (T < 1e-4) — the reference, and our
// full-sort path, depth-sort all of them.  Front-to-back compositing only ever needs a PREFIX of each tile's depth
// order, so this kernel produces that order lazily, per tile:
//
//   1. n <= 2048: load the tile's keys into shared memory, sort (rank sort / LSD radix on the varying depth bits, ties
//      by id), composite.
//   2. n  > 2048: one MSD pass partitions the tile's keys (global, L2-resident) into <= 256 depth buckets by their
//      highest varying depth byte; buckets are then grouped front to back into chunks (<= 512 keys first, doubling up to
//      2048), each chunk is sorted in shared memory and composited; the loop stops as soon as all 256 pixels are done — the buckets behind
//      are never sorted, never gathered.
//
// Per-pixel arithmetic and traversal order are exactly those of k_blend (fovgs_blend.cu), so images stay bit-identical
// to the reference; only WHERE batch boundaries fall differs, which no pixel can observe (a pixel's `done` is its own).
// The training family (SUM / MAX / LWMC) runs on the same producer with the reference's batch-exact semantics: the sorted
// prefix is appended to `point_list` (what its backward walks) and composited in 256-entry batches at absolute list
// positions (lazy_tile_sum).  Only `out_point_list` requests / FOVGS_OPT_FULL_SORT use the full-sort path (fovgs_sort.cu
// + fovgs_blend.cu).
#include <type_traits>
#include "fovgs_internal.cuh"
#include "fovgs_tma.cuh"

namespace fovgs {

#ifdef FOVGS_TILE_TIMING
__device__ uint32_t g_tile_times[4 * 65536];
#endif

#ifndef LAZY_LCAP
#define LAZY_LCAP 1024          // keys of one sorted group held in shared memory (two buffers of LCAP u64); A/B: 2048 is 3 % slower
#endif
constexpr int LCAP = LAZY_LCAP;
// How a 256-batch of blend records (3-4 float4 per sorted instance, gathered by Gaussian id) reaches shared memory:
//   0  register-staged: LDG into registers, STS after the barrier; LAZY_PREFETCH=1 loads the next batch one batch ahead
//      (12-16 live registers across the compositing loop), LAZY_PREFETCH=0 loads it when it is needed (latency exposed)
//      [default: LAZY_STAGE 0 with LAZY_PREFETCH 0 — measured fastest, see profiles/README.md "staging A/B"]
//   1  cp.async (LDGSTS), double buffered: the next batch flies global -> shared while the current one is composited,
//      no registers, no STS; costs a second 16 KB staging buffer per CTA (less L1 left for the gathers)
//   2  cp.async.bulk (TMA, UBLKCP) + one mbarrier per buffer, double buffered: same data flow through the bulk-copy engine
#ifndef LAZY_STAGE
#define LAZY_STAGE 0
#endif
#ifndef LAZY_PREFETCH
#define LAZY_PREFETCH 0
#endif
constexpr int NSTAGE = LAZY_STAGE ? 2 : 1;
#ifndef LAZY_RANK_MAX
#define LAZY_RANK_MAX 256u
#endif
#ifndef LAZY_FIRST_GROUP
#define LAZY_FIRST_GROUP 512u   // first sorted group of a partitioned tile; doubles up to LCAP
#endif
#ifndef LAZY_CTAS
#define LAZY_CTAS 4            // CTAs per SM the kernel is compiled for (register cap 64)
#endif
#ifndef LAZY_MU
#define LAZY_MU 8
#endif
#ifndef LAZY_DIRECT_MAX
#define LAZY_DIRECT_MAX ((uint32_t)LAZY_LCAP)    // tiles up to this many instances are sorted whole in shared memory (<= LCAP)
#endif
constexpr float kStartBlendL = 0.5f;

struct SortScratch { uint32_t whist[8][256]; uint32_t totals[256]; };
struct BlendStage { float4 sA[256], sB[256], sC[256], sD[256]; };
struct LazySmem {
    uint64_t keys[2][LCAP];
    union {                 // a tile alternates sort and composite phases (block barriers in between): one footprint
        SortScratch srt;
        BlendStage bl[NSTAGE];
    };
    alignas(8) uint64_t bar[2];   // LAZY_STAGE 2: completion barrier of each staging buffer
    uint32_t stage_parity;        // LAZY_STAGE 2: phase bits of the two barriers (they live as long as the kernel, across groups and tiles)
    uint32_t bucket_off[257];
    uint32_t bucket_cur[256];
    uint32_t wsum[8];
    unsigned long long vary;
    uint32_t ticket;        // index of the tile this CTA is working on (persistent kernel: drawn from FrameHeader::lazy_ticket)
    uint32_t consumed;      // frame statistics of the current tile, kept here instead of in registers that live across the
    uint32_t kept[8];       //   whole tile: instances staged (thread 0) / (warp, splat) pairs that passed block_may_touch
    uint8_t widx[8][256];   // per warp: batch slots whose footprint can reach the warp's 8x4 pixel block
};

// ---- conservative footprint test of one splat against a warp's pixel block -----------------------------------------
// A splat changes a pixel only if its falloff exponent is >= -4.5 and opacity*exp(exponent) >= 1/255
// (FOV/forward.cu:556-566, OBB/forward.cu:330-342).  q = -exponent is a convex quadratic in (dx, dy); its minimum over the
// block [X0, X0+7] x [Y0, Y0+3] is 0 when the centre lies inside and otherwise sits on one of the four edges, where it
// is a clamped 1-D parabola.  The splat is dropped for this warp only when that minimum exceeds the threshold by more
// than a bound on every fp32 rounding involved (relative 8e-6 of the largest term magnitudes + 2e-3 absolute), so a
// dropped splat is one the exact per-pixel code below would have skipped for all 32 pixels: images are unchanged.
__device__ __forceinline__ bool block_may_touch(const float4 a, const float conz, const float op, const float X0, const float Y0,
                                                const float tau_cap = 4.5f) {
    const float dx0 = a.x - (X0 + 7.0f), dx1 = a.x - X0, dy0 = a.y - (Y0 + 3.0f), dy1 = a.y - Y0;
    if (dx0 <= 0.0f && dx1 >= 0.0f && dy0 <= 0.0f && dy1 >= 0.0f) return true;
    const float A = a.z, B = a.w, C = conz;
    const float tau = fminf(tau_cap, __logf(255.0f * op));   // tau_cap = +inf: the vanilla mode has no falloff cut
    // MUFU.RCP (1 ulp) instead of the IEEE reciprocal (8 instructions each): the clamped parabola minimum moves by second order
    // in the error of t, far inside the slack of the final comparison
    float iA, iC;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iA) : "f"(A));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iC) : "f"(C));
    auto edge_x = [&](float ex) {   // dx fixed
        const float t = fminf(fmaxf(-B * ex * iC, dy0), dy1);
        return 0.5f * A * ex * ex + B * ex * t + 0.5f * C * t * t;
    };
    auto edge_y = [&](float ey) {   // dy fixed
        const float t = fminf(fmaxf(-B * ey * iA, dx0), dx1);
        return 0.5f * C * ey * ey + B * ey * t + 0.5f * A * t * t;
    };
    const float qmin = fminf(fminf(edge_x(dx0), edge_x(dx1)), fminf(edge_y(dy0), edge_y(dy1)));
    const float mx = fmaxf(fabsf(dx0), fabsf(dx1)), my = fmaxf(fabsf(dy0), fabsf(dy1));
    const float mag = 0.5f * A * mx * mx + 0.5f * C * my * my + fabsf(B) * mx * my;
    return !(qmin - (8e-6f * mag + 2e-3f) > tau);   // NaN anywhere -> keep
}

// ---- per-pixel compositing state (identical arithmetic to k_blend) -------------------------------------------------
struct PixPS1 {
    BlendExpConsts ek;      // OBB/forward.cu:251-384
    float T, C0, C1, C2;
    bool done;
    __device__ __forceinline__ void init(bool inside, const float* ec) { T = 1.0f; C0 = C1 = C2 = 0.f; done = !inside; ek.load(ec); }
    __device__ __forceinline__ float power(const BlendStage& bl, int j, float pixx, float pixy) const {
        const float4 a = bl.sA[j];
        const float dx = FS(a.x, pixx), dy = FS(a.y, pixy);
        return gauss_power(a.z, a.w, bl.sB[j].x, dx, dy);
    }
    // One divergent region per splat (the falloff cut, which most lanes fail); the two rarer outcomes inside it — alpha below
    // 1/255, pixel saturated — are selects, not branches: same values, no reconvergence barriers in the inner loop.
    __device__ __forceinline__ void apply(const BlendStage& bl, int j, float power) {
        if (done || power > 0.0f || power < -4.5f) return;
        const float4 b = bl.sB[j];
        const float4 c = bl.sC[j];
        const float alpha = fminf(0.99f, FM(b.y, blend_exp(power, ek)));
        const float test_T = FM(T, FS(1.0f, alpha));
        const bool vis = !(alpha < 1.0f / 255.0f);
        const bool fin = vis && test_T < 0.0001f;
        const bool acc = vis && !fin;
        const float w = FM(alpha, T);
        C0 = acc ? FF(c.x, w, C0) : C0; C1 = acc ? FF(c.y, w, C1) : C1; C2 = acc ? FF(c.z, w, C2) : C2;
        T = acc ? test_T : T;
        done = fin;
    }
};
struct PixFov {
    BlendExpConsts ek;      // FOV/forward.cu:490-609
    float T, C0, C1, C2;
    bool done;
    __device__ __forceinline__ void init(bool inside, const float* ec) { T = 1.0f; C0 = C1 = C2 = 0.f; done = !inside; ek.load(ec); }
    __device__ __forceinline__ float power(const BlendStage& bl, int j, float pixx, float pixy) const {
        const float4 a = bl.sA[j];
        const float dx = FS(a.x, pixx), dy = FS(a.y, pixy);
        return gauss_power(a.z, a.w, bl.sB[j].x, dx, dy);
    }
    __device__ __forceinline__ void apply(const BlendStage& bl, int j, float power) {   // see PixPS1::apply
        if (done || power > 0.0f || power < -4.5f) return;
        const float4 c = bl.sC[j];
        const float alpha = fminf(0.99f, FM(c.x, blend_exp(power, ek)));
        const float test_T = FM(T, FS(1.0f, alpha));
        const bool vis = !(alpha < 1.0f / 255.0f);
        const bool fin = vis && test_T < 0.0001f;
        const bool acc = vis && !fin;
        const float w = FM(alpha, T);
        C0 = acc ? FF(c.y, w, C0) : C0; C1 = acc ? FF(c.z, w, C1) : C1; C2 = acc ? FF(c.w, w, C2) : C2;
        T = acc ? test_T : T;
        done = fin;
    }
};
struct PixFovBlend {
    BlendExpConsts ek;  // FOV/forward.cu:262-476
    float T1, T2, A0, A1, A2, B0, B1, B2, L2_f;
    bool L1_done, L2_done, done;
    __device__ __forceinline__ void init(bool inside, float est, int L2, float l2f, const float* ec) {
        T1 = T2 = 1.0f; A0 = A1 = A2 = B0 = B1 = B2 = 0.f; L2_f = l2f;
        L1_done = est > (float)L2; L2_done = false; done = !inside;
        ek.load(ec);
    }
    __device__ __forceinline__ float power(const BlendStage& bl, int j, float pixx, float pixy) const {
        const float4 a = bl.sA[j];
        const float dx = FS(a.x, pixx), dy = FS(a.y, pixy);
        return gauss_power(a.z, a.w, bl.sB[j].x, dx, dy);
    }
    __device__ __forceinline__ void apply(const BlendStage& bl, int j, float power) {   // selects, not branches: see PixPS1::apply
        if (done || power > 0.0f || power < -4.5f) return;
        const float4 b = bl.sB[j];
        const float4 c1 = bl.sC[j];
        const float e = blend_exp(power, ek);
        {
            const float alpha1 = fminf(0.99f, FM(c1.x, e));
            const float test_T1 = FM(T1, FS(1.0f, alpha1));
            const bool vis = !L1_done && !(alpha1 < 1.0f / 255.0f);
            const bool fin = vis && test_T1 < 0.0001f;
            const bool acc = vis && !fin;
            const float w = FM(alpha1, T1);
            A0 = acc ? FF(c1.y, w, A0) : A0; A1 = acc ? FF(c1.z, w, A1) : A1; A2 = acc ? FF(c1.w, w, A2) : A2;
            T1 = acc ? test_T1 : T1;
            L1_done = L1_done || fin;
        }
        // the splat belongs to level L2 only when highest_level + 1 >= L2 (FOV/forward.cu:425): a property of the SPLAT, the same
        // for every lane still here, so this branch never diverges — and three in five Gaussians carry highest level 0
        if (!(FA(b.y, 1.0f) < L2_f)) {
            const float4 c2 = bl.sD[j];
            const float alpha2 = fminf(0.99f, FM(c2.x, e));
            const float test_T2 = FM(T2, FS(1.0f, alpha2));
            const bool vis = !L2_done && !(alpha2 < 1.0f / 255.0f);
            const bool fin = vis && test_T2 < 0.0001f;
            const bool acc = vis && !fin;
            const float w = FM(alpha2, T2);
            B0 = acc ? FF(c2.y, w, B0) : B0; B1 = acc ? FF(c2.z, w, B1) : B1; B2 = acc ? FF(c2.w, w, B2) : B2;
            T2 = acc ? test_T2 : T2;
            L2_done = L2_done || fin;
        }
        if (L1_done && L2_done) done = true;
    }
};

struct PixSmfrBlend {
    BlendExpConsts ek;  // naive_pcheck_obb/cuda_rasterizer/forward.cu:262-440: one alpha for both levels
    float T1, T2, A0, A1, A2, B0, B1, B2, L2_f;
    bool L1_done, L2_done, done;
    __device__ __forceinline__ void init(bool inside, float est, int L2, float l2f, const float* ec) {
        T1 = T2 = 1.0f; A0 = A1 = A2 = B0 = B1 = B2 = 0.f; L2_f = l2f;
        L1_done = est > (float)L2; L2_done = false; done = !inside;
        ek.load(ec);
    }
    __device__ __forceinline__ float power(const BlendStage& bl, int j, float pixx, float pixy) const {
        const float4 a = bl.sA[j];
        const float dx = FS(a.x, pixx), dy = FS(a.y, pixy);
        return gauss_power(a.z, a.w, bl.sB[j].x, dx, dy);
    }
    __device__ __forceinline__ void apply(const BlendStage& bl, int j, float power) {
        if (power > 0.0f || power < -4.5f) return;
        const float4 b = bl.sB[j];
        const float4 c = bl.sC[j];
        const float alpha1 = fminf(0.99f, FM(c.x, blend_exp(power, ek)));
        if (!L1_done) {
            // a live L1 drops the entry for BOTH levels (:403-405); once L1 is done the alpha test no longer applies
            if (alpha1 < 1.0f / 255.0f) return;
            const float test_T1 = FM(T1, FS(1.0f, alpha1));
            L1_done = test_T1 < 0.0001f;
            if (!L1_done) {
                const float w = FM(alpha1, T1);
                A0 = FF(c.y, w, A0); A1 = FF(c.z, w, A1); A2 = FF(c.w, w, A2);
                T1 = test_T1;
            }
        }
        if (!L2_done) {
            if (!(FA(b.y, 1.0f) < L2_f)) {
                const float test_T2 = FM(T2, FS(1.0f, alpha1));
                L2_done = test_T2 < 0.0001f;
                if (!L2_done) {
                    const float w = FM(alpha1, T2);
                    B0 = FF(c.y, w, B0); B1 = FF(c.z, w, B1); B2 = FF(c.w, w, B2);
                    T2 = test_T2;
                }
            }
        }
        if (L1_done && L2_done) done = true;
    }
};

// ---- shared-memory sort of the m keys in keys[0][0..m); returns the buffer index that holds the sorted keys ----------
__device__ __forceinline__ int lazy_sort_group(LazySmem& sm, const uint32_t m) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (m <= LAZY_RANK_MAX) {   // rank sort: m comparisons per key; above this the radix passes below are cheaper
        uint64_t k = 0;
        uint32_t r = 0;
        if ((uint32_t)tid < m) {
            k = sm.keys[0][tid];
            for (uint32_t j = 0; j < m; j++) r += sm.keys[0][j] < k;
        }
        if ((uint32_t)tid < m) sm.keys[1][r] = k;
        __syncthreads();
        return 1;
    }
    if (tid == 0) sm.vary = 0ull;
    __syncthreads();
    {
        const uint64_t k0 = sm.keys[0][0];
        uint64_t v = 0;
        for (uint32_t i = tid; i < m; i += 256) v |= (sm.keys[0][i] ^ k0);
#pragma unroll
        for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicOr(&sm.vary, (unsigned long long)v);
    }
    __syncthreads();
    const uint64_t vary = sm.vary;
    const uint32_t chunk = ((m + 7) / 8 + 31) & ~31u;
    const uint32_t wbeg = min(m, warp * chunk), wend = min(m, wbeg + chunk);
    int cur = 0;
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 32 + 8 * pass;
        if (((vary >> shift) & 0xffull) == 0) continue;
        uint64_t* src = sm.keys[cur];
        uint64_t* dst = sm.keys[cur ^ 1];
        for (int i = tid; i < 8 * 256; i += 256) (&sm.srt.whist[0][0])[i] = 0;
        __syncthreads();
        for (uint32_t i = wbeg + lane; i < wend; i += 32) atomicAdd(&sm.srt.whist[warp][(src[i] >> shift) & 0xff], 1u);
        __syncthreads();
        {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) { const uint32_t c = sm.srt.whist[w][tid]; sm.srt.whist[w][tid] = t; t += c; }
            uint32_t x = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) sm.wsum[warp] = x;
            __syncthreads();
            uint32_t base = x - t;
            for (int w = 0; w < warp; w++) base += sm.wsum[w];
            sm.srt.totals[tid] = base;
        }
        __syncthreads();
        for (int i = tid; i < 8 * 256; i += 256) (&sm.srt.whist[0][0])[i] += sm.srt.totals[i & 255];
        __syncthreads();
        for (uint32_t i0 = wbeg; i0 < wend; i0 += 32) {
            const uint32_t i = i0 + lane;
            const bool valid = i < wend;
            const uint64_t key = valid ? src[i] : 0ull;
            const uint32_t d = valid ? (uint32_t)((key >> shift) & 0xff) : (256u + lane);
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t base = 0;
            if (valid) base = sm.srt.whist[warp][d];
            __syncwarp();
            if (valid) {
                dst[base + rank] = key;
                if (rank == 0) sm.srt.whist[warp][d] = base + __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        cur ^= 1;
    }
    uint64_t* src = sm.keys[cur];
    for (uint32_t i = tid; i < m; i += 256) {
        const uint32_t dk = (uint32_t)(src[i] >> 32);
        const bool head = (i == 0) || ((uint32_t)(src[i - 1] >> 32) != dk);
        if (head && i + 1 < m && (uint32_t)(src[i + 1] >> 32) == dk) {
            uint32_t j = i + 1;
            while (j < m && (uint32_t)(src[j] >> 32) == dk) j++;
            for (uint32_t a = i + 1; a < j; a++) {
                const uint64_t k = src[a];
                uint32_t b = a;
                while (b > i && src[b - 1] > k) { src[b] = src[b - 1]; b--; }
                src[b] = k;
            }
        }
    }
    __syncthreads();
    return cur;
}

// ---- full ordering of m keys in global memory (src/dst ping-pong); returns the array that holds the result ----------
__device__ __forceinline__ const uint64_t* lazy_global_sort(LazySmem& sm, uint64_t* src, uint64_t* dst, const uint32_t m) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sm.vary = 0ull;
    __syncthreads();
    {
        const uint64_t k0 = src[0];
        uint64_t v = 0;
        for (uint32_t i = tid; i < m; i += 256) v |= (src[i] ^ k0);
#pragma unroll
        for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicOr(&sm.vary, (unsigned long long)v);
    }
    __syncthreads();
    const uint64_t vary = sm.vary;
    const uint32_t chunk = ((m + 7) / 8 + 31) & ~31u;
    const uint32_t wbeg = min(m, warp * chunk), wend = min(m, wbeg + chunk);
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 32 + 8 * pass;
        if (((vary >> shift) & 0xffull) == 0) continue;
        for (int i = tid; i < 8 * 256; i += 256) (&sm.srt.whist[0][0])[i] = 0;
        __syncthreads();
        for (uint32_t i = wbeg + lane; i < wend; i += 32) atomicAdd(&sm.srt.whist[warp][(src[i] >> shift) & 0xff], 1u);
        __syncthreads();
        {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) { const uint32_t c = sm.srt.whist[w][tid]; sm.srt.whist[w][tid] = t; t += c; }
            uint32_t x = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) sm.wsum[warp] = x;
            __syncthreads();
            uint32_t base = x - t;
            for (int w = 0; w < warp; w++) base += sm.wsum[w];
            sm.srt.totals[tid] = base;
        }
        __syncthreads();
        for (int i = tid; i < 8 * 256; i += 256) (&sm.srt.whist[0][0])[i] += sm.srt.totals[i & 255];
        __syncthreads();
        for (uint32_t i0 = wbeg; i0 < wend; i0 += 32) {
            const uint32_t i = i0 + lane;
            const bool valid = i < wend;
            const uint64_t key = valid ? src[i] : 0ull;
            const uint32_t d = valid ? (uint32_t)((key >> shift) & 0xff) : (256u + lane);
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t base = 0;
            if (valid) base = sm.srt.whist[warp][d];
            __syncwarp();
            if (valid) {
                dst[base + rank] = key;
                if (rank == 0) sm.srt.whist[warp][d] = base + __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        uint64_t* t = src; src = dst; dst = t;
    }
    for (uint32_t i = tid; i < m; i += 256) {
        const uint32_t dk = (uint32_t)(src[i] >> 32);
        const bool head = (i == 0) || ((uint32_t)(src[i - 1] >> 32) != dk);
        if (head && i + 1 < m && (uint32_t)(src[i + 1] >> 32) == dk) {
            uint32_t j = i + 1;
            while (j < m && (uint32_t)(src[j] >> 32) == dk) j++;
            for (uint32_t a = i + 1; a < j; a++) {
                const uint64_t k = src[a];
                uint32_t b = a;
                while (b > i && src[b - 1] > k) { src[b] = src[b - 1]; b--; }
                src[b] = k;
            }
        }
    }
    __syncthreads();
    return src;
}

// ---- composite the m sorted keys `sk` (shared memory) in 256-entry batches; returns true when every pixel is done ----
// Thread layout: warp w owns the 8x4 pixel block at (8*(w&1), 4*(w>>1)) of the tile.  Per batch each warp first builds
// the list of slots that can reach its block (block_may_touch, one slot per lane), then all lanes walk that list.
// KIND: 0 PS=1, 1 foveated plain tile, 2 FOV blending tile (two level records), 3 SMFR blending tile (one record).
// R = float4 per record; S1 / S2 = record slots of the level-L1 / level-L2 (opacity, r, g, b).
template <int KIND, int R, class PIX>
__device__ __forceinline__ bool lazy_blend_group(LazySmem& sm, const Workspace& ws, const uint64_t* sk, const uint32_t m,
                                                 PIX& px, const float pixx, const float pixy, const float blkx,
                                                 const float blky, const int S1, const int S2) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* __restrict__ wl = sm.widx[warp];
    const uint32_t* __restrict__ wl4 = reinterpret_cast<const uint32_t*>(sm.widx[warp]);
#if LAZY_STAGE == 0
    float4 r0, r1, r2, r3;
    bool valid;
    auto fetch = [&](uint32_t pos) {
        valid = pos < m;
        if (valid) {
            const uint32_t id = (uint32_t)sk[pos];
            const float4* __restrict__ rec = ws.rec + (size_t)R * id;
            r0 = rec[0]; r1 = rec[1];
            r2 = (KIND == 0) ? rec[2] : rec[S1];
            if (KIND == 2) r3 = rec[S2];
        }
    };
    if (LAZY_PREFETCH) fetch(tid);
#else
    // asynchronous staging: thread t copies the records of instance b0 + t straight into slot t of a staging buffer
    uint32_t parity = sm.stage_parity;      // LAZY_STAGE 2: phase bits of the two mbarriers (uniform over the CTA)
    auto issue = [&](uint32_t b0, int buf) {
        const uint32_t pos = b0 + tid;
        BlendStage& st = sm.bl[buf];
        const bool valid = pos < m;
        const float4* __restrict__ rec = nullptr;
        if (valid) rec = ws.rec + (size_t)R * (uint32_t)sk[pos];
        const float4* __restrict__ c1 = (KIND == 0) ? rec + 2 : rec + S1;
#if LAZY_STAGE == 1
        if (valid) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&st.sA[tid])), "l"(rec) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&st.sB[tid])), "l"(rec + 1) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&st.sC[tid])), "l"(c1) : "memory");
            if (KIND == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&st.sD[tid])), "l"(rec + S2) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
#else
        mbar_arrive_expect_tx(&sm.bar[buf], valid ? (KIND == 2 ? 64u : 48u) : 0u);
        if (valid) {
            bulk_g2s(&st.sA[tid], rec, 16, &sm.bar[buf]);
            bulk_g2s(&st.sB[tid], rec + 1, 16, &sm.bar[buf]);
            bulk_g2s(&st.sC[tid], c1, 16, &sm.bar[buf]);
            if (KIND == 2) bulk_g2s(&st.sD[tid], rec + S2, 16, &sm.bar[buf]);
        }
#endif
    };
    auto wait_for = [&](int buf) {
#if LAZY_STAGE == 1
        (void)buf;
        asm volatile("cp.async.wait_all;" ::: "memory");
#else
        mbar_wait(&sm.bar[buf], (parity >> buf) & 1u);
        parity ^= 1u << buf;
#endif
    };
#if LAZY_STAGE == 2
    fence_proxy_async_smem();     // the sort scratch (generic-proxy stores) shares this memory with the staging buffers
#endif
    issue(0, 0);
#endif
    int buf = 0;
    for (uint32_t b0 = 0; b0 < m; b0 += 256) {
#if LAZY_STAGE != 0
        wait_for(buf);                                              // this thread's copies of the current batch have landed
        if (__syncthreads_count(px.done) == 256) {                  // ... everybody's: nothing is in flight at this point
            if (tid == 0) sm.stage_parity = parity;                 // (read again after the caller's block barrier)
            return true;
        }
        if (b0 + 256 < m) issue(b0 + 256, buf ^ 1);                 // next batch -> the buffer the previous batch has left
#else
        if (__syncthreads_count(px.done) == 256) return true;
        if (!LAZY_PREFETCH) fetch(b0 + tid);
        if (valid) {
            sm.bl[0].sA[tid] = r0; sm.bl[0].sB[tid] = r1; sm.bl[0].sC[tid] = r2;
            if (KIND == 2) sm.bl[0].sD[tid] = r3;
        }
        __syncthreads();
#endif
        const BlendStage& bl = sm.bl[buf];
        const int lim = (int)min(256u, m - b0);
        if (tid == 0) sm.consumed += (uint32_t)lim;
#if LAZY_STAGE == 0
        if (LAZY_PREFETCH && b0 + 256 < m) fetch(b0 + 256 + tid);
#else
        buf ^= 1;
#endif
        if (__all_sync(0xffffffffu, px.done)) continue;
        uint32_t cnt = 0;
        for (int jb = 0; jb < lim; jb += 32) {
            const int j = jb + lane;
            bool keep = false;
            if (j < lim) {
                // SMFR blending tiles composite level L2 whatever the alpha once L1 is done: only the -4.5 bound holds
                const float op = (KIND == 0) ? bl.sB[j].y : (KIND == 1) ? bl.sC[j].x :
                                 (KIND == 2) ? fmaxf(bl.sC[j].x, bl.sD[j].x) : 1.0f;
                keep = block_may_touch(bl.sA[j], bl.sB[j].x, op, blkx, blky);
            }
            const unsigned mk = __ballot_sync(0xffffffffu, keep);
            if (keep) wl[cnt + __popc(mk & ((1u << lane) - 1u))] = (uint8_t)j;
            cnt += __popc(mk);
        }
        __syncwarp();
        if (lane == 0) sm.kept[warp] += cnt;
        // four falloff exponents are evaluated together (independent shared loads + FMAs in flight), then applied in
        // list order: the per-pixel sequence of operations is unchanged
        uint32_t k = 0;
        for (; !px.done && k + 3 < cnt; k += 4) {
            const uint32_t q = wl4[k >> 2];
            const int j0 = q & 0xff, j1 = (q >> 8) & 0xff, j2 = (q >> 16) & 0xff, j3 = q >> 24;
            const float p0 = px.power(bl, j0, pixx, pixy), p1 = px.power(bl, j1, pixx, pixy);
            const float p2 = px.power(bl, j2, pixx, pixy), p3 = px.power(bl, j3, pixx, pixy);
            px.apply(bl, j0, p0);
            if (!px.done) px.apply(bl, j1, p1);
            if (!px.done) px.apply(bl, j2, p2);
            if (!px.done) px.apply(bl, j3, p3);
        }
        for (; !px.done && k < cnt; k++) { const int j = wl[k]; px.apply(bl, j, px.power(bl, j, pixx, pixy)); }
    }
#if LAZY_STAGE != 0
    if (tid == 0) sm.stage_parity = parity;
#endif
    return false;
}

// ---- front-to-back producer of sorted key groups for one tile --------------------------------------------------------
// `consume(sk, m, last)` is called by the whole CTA with m sorted keys in shared memory, groups in depth order; `last` says
// no key follows.  It returns true (uniformly) when the tile needs nothing more; the keys behind are then never sorted.
// One loop, ONE sort site and ONE consume site: the three sources of a group (the whole small tile, the next run of MSD
// buckets, the next slice of an over-long bucket that was ordered in global memory) only differ in how the keys reach
// shared memory — the compositing code, which is most of the kernel, is instantiated once (it used to be inlined at three
// call sites: 170 KB of SASS per instantiation).
template <class CONSUME>
__device__ __forceinline__ void lazy_for_each_group(LazySmem& sm, const Workspace& ws, const int tile, CONSUME&& consume) {
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t cap = ws.hdr->cap;
    const uint32_t sbeg = min(ws.tile_offset[tile], cap), send = min(ws.tile_offset[tile + 1], cap);
    const uint32_t n = send - sbeg;
    if (n == 0) return;
    const uint64_t* __restrict__ gA = ws.keysA + sbeg;
    uint64_t* gB = ws.keysB + sbeg;
    const bool direct = n <= LAZY_DIRECT_MAX;
    if (!direct) {
        // ---- MSD partition of the tile's keys by their highest varying depth byte ----
        constexpr int MU = LAZY_MU;   // keys in flight per thread in the partition passes (L2 round trips overlap)
        // digit = 8 bits ending at the highest varying depth bit (all higher bits are equal inside the tile).  The position is
        // guessed from the first 2048 keys (the scatter left them in arbitrary order) and verified for free by the histogram
        // pass, which ORs the differences of ALL keys; a wrong guess (never seen in practice) repeats the histogram.
        const uint64_t k0 = gA[0];
        auto digit_shift = [](uint32_t vhi) { return vhi ? max(32, 32 + (31 - __clz(vhi)) - 7) : 32; };
        if (tid == 0) sm.vary = 0ull;
        __syncthreads();
        {
            uint64_t v = 0;
            uint64_t k[MU];
#pragma unroll
            for (int u = 0; u < MU; u++) { const uint32_t i = tid + u * 256; k[u] = (i < n) ? gA[i] : k0; }
#pragma unroll
            for (int u = 0; u < MU; u++) v |= (k[u] ^ k0);
#pragma unroll
            for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && v) atomicOr(&sm.vary, (unsigned long long)v);
        }
        __syncthreads();
        int shift = digit_shift((uint32_t)(sm.vary >> 32));
        for (;;) {
            sm.bucket_cur[tid] = 0;
            __syncthreads();
            uint64_t v = 0;
            for (uint32_t i0 = tid; i0 < n; i0 += 256 * MU) {
                uint64_t k[MU];
#pragma unroll
                for (int u = 0; u < MU; u++) { const uint32_t i = i0 + u * 256; k[u] = (i < n) ? gA[i] : k0; }
#pragma unroll
                for (int u = 0; u < MU; u++) {
                    v |= (k[u] ^ k0);
                    if (i0 + u * 256 < n) atomicAdd(&sm.bucket_cur[(k[u] >> shift) & 0xff], 1u);
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && v) atomicOr(&sm.vary, (unsigned long long)v);
            __syncthreads();
            const int exact = digit_shift((uint32_t)(sm.vary >> 32));
            if (exact == shift) break;
            shift = exact;          // uniform over the CTA: sm.vary is complete after the barrier
            __syncthreads();
        }
        {   // exclusive scan of the 256 bucket sizes
            const uint32_t c = sm.bucket_cur[tid];
            uint32_t x = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) sm.wsum[tid >> 5] = x;
            __syncthreads();
            uint32_t base = x - c;
            for (int w = 0; w < (tid >> 5); w++) base += sm.wsum[w];
            sm.bucket_off[tid] = base;
            if (tid == 255) sm.bucket_off[256] = base + c;
            sm.bucket_cur[tid] = base;
        }
        __syncthreads();
        for (uint32_t i0 = tid; i0 < n; i0 += 256 * MU) {
            uint64_t k[MU];
#pragma unroll
            for (int u = 0; u < MU; u++) { const uint32_t i = i0 + u * 256; k[u] = (i < n) ? gA[i] : 0ull; }
#pragma unroll
            for (int u = 0; u < MU; u++) {
                if (i0 + u * 256 < n) {
                    const uint32_t pos = atomicAdd(&sm.bucket_cur[(k[u] >> shift) & 0xff], 1u);
                    gB[pos] = k[u];
                }
            }
        }
        __syncthreads();   // gB is read below by this CTA only
    }
    // ---- front to back: one group per iteration ----
    int b = 0;                               // next MSD bucket
    // group size grows geometrically: a tile that saturates within its nearest few hundred splats sorts only those, a tile that
    // needs everything pays at most a few extra group boundaries
    uint32_t gcap = LAZY_FIRST_GROUP;
    const uint64_t* big = nullptr;           // an over-long bucket, completely ordered in global memory, handed over in slices
    uint32_t big_n = 0, big_at = 0, big_g0 = 0;
    bool first = true;
    for (;;) {
        const uint64_t* src;                 // where the group's keys come from
        uint32_t m, end_pos;                 // group size, position of its end inside the tile's list
        bool presorted = false;
        if (direct) {
            if (!first) break;
            src = gA; m = n; end_pos = n;
        } else if (big != nullptr) {
            m = min((uint32_t)LCAP, big_n - big_at);
            src = big + big_at; end_pos = big_g0 + big_at + m; presorted = true;
            big_at += m;
            if (big_at >= big_n) { big = nullptr; b++; }
        } else {
            if (b >= 256) break;
            const uint32_t g0 = sm.bucket_off[b];
            // buckets [b, e) fit the group together: bucket_off is monotone, so the fitting ones form a prefix — count them
            int e = b + __syncthreads_count(tid >= b && sm.bucket_off[tid + 1] - g0 <= gcap);
            if (e == b && gcap < (uint32_t)LCAP)   // the next bucket alone exceeds the small cap: take what the buffer holds
                e = b + __syncthreads_count(tid >= b && sm.bucket_off[tid + 1] - g0 <= (uint32_t)LCAP);
            gcap = min(gcap * 2u, (uint32_t)LCAP);
            if (e == b) {
                // one bucket larger than the shared buffer (>= 2048 keys agreeing in all depth bits above `shift`):
                // order it completely with the block radix sort on the global ping-pong ranges (this tile's slice of
                // keysA is free scratch after the partition), then hand it over in LCAP-sized slices.
                big_n = sm.bucket_off[b + 1] - g0;
                big = lazy_global_sort(sm, gB + g0, const_cast<uint64_t*>(gA) + g0, big_n);
                big_at = 0; big_g0 = g0;
                continue;
            }
            m = sm.bucket_off[e] - g0;
            b = e;
            if (m == 0) continue;
            src = gB + g0; end_pos = g0 + m;
        }
        first = false;
#pragma unroll 4
        for (uint32_t i = tid; i < m; i += 256) sm.keys[0][i] = src[i];
        __syncthreads();
        const int cur = presorted ? 0 : lazy_sort_group(sm, m);
        const bool all_done = consume(sm.keys[cur], m, end_pos == n);
        __syncthreads();
        if (all_done) break;
    }
}

template <int KIND, int R, class PIX>
__device__ __forceinline__ void lazy_tile(LazySmem& sm, const Workspace& ws, const int tile, PIX& px, const float pixx,
                                          const float pixy, const float blkx, const float blky, const int S1, const int S2) {
    if (threadIdx.x < 8) sm.kept[threadIdx.x] = 0;
    if (threadIdx.x == 8) sm.consumed = 0;
    // (the producer's first block barrier orders these stores before any use)
    lazy_for_each_group(sm, ws, tile, [&](const uint64_t* sk, uint32_t m, bool) {
        return lazy_blend_group<KIND, R>(sm, ws, sk, m, px, pixx, pixy, blkx, blky, S1, S2);
    });
    __syncthreads();
    if (threadIdx.x == 0 && sm.consumed) atomicAdd(&ws.hdr->stats.reserved[0], sm.consumed);
    if (threadIdx.x < 8 && sm.kept[threadIdx.x]) atomicAdd(&ws.hdr->stats.reserved[1], sm.kept[threadIdx.x]);   // (warp, splat) pairs that passed block_may_touch
}

// ---- training variant (SUM/forward.cu:298-430) on the lazy producer ---------------------------------------------------
// The reference defines three things by its 256-entry batches at ABSOLUTE list positions: a batch is staged iff some pixel
// is still live when it starts; every entry of a staged batch gets `gaussians_count += 1`; `n_contrib` is the 1-based
// position of a pixel's last contributing entry.  So the sorted ids are appended to `point_list` (the backward walks that
// prefix) as groups arrive, and batches are composited from there once complete (or once the list has ended).
// `contributions[id] += alpha*T` per hit (one fp32 atomic per pixel hit in the reference) is summed over the warp, then
// over the tile in shared memory, and leaves as one atomic per staged entry.
//
// The pruning-metric variants keep different statistics on the same traversal (STAT):
//   STAT_MAX  (pcheck_obb_max/forward.cu:381,400): `gaussians_count += 1` per (live pixel, entry) that passes the falloff
//             cut — BEFORE the alpha test, so the block footprint test may only use the -4.5 bound — and
//             `contributions = max(contributions, alpha*T)`.  alpha*T > 0, so the float maximum is the maximum of the bit
//             patterns: warp `redux.max`, shared atomicMax, one global atomicMax per staged entry; counts likewise
//             (ballot+popc, shared add, one global add).  Both are exact, hence bit-identical to the reference.
//   STAT_LWMC (pcheck_obb_loss_weighted_max_count/forward.cu:347-348,403-410,435): counts as SUM; every pixel remembers the
//             entry with its largest alpha*T (strict >, so the first maximum wins) and finally adds loss_map[pixel] to it —
//             to Gaussian 0 when nothing contributed, as the reference does.
struct SumSmemExtra {
    float acc[256];
    int ids[256];
    int cnt[256];
};

template <int STAT>
__device__ __forceinline__ void lazy_tile_sum(LazySmem& sm, SumSmemExtra& sx, const Workspace& ws, const FrameInputs& in,
                                              const int tile, const bool inside, const float pixx, const float pixy,
                                              const float blkx, const float blky, float& T, float& C0, float& C1, float& C2,
                                              uint32_t& last_contributor, int& max_idx) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t cap = ws.hdr->cap;
    const uint32_t rbeg = min(ws.tile_offset[tile], cap), rend = min(ws.tile_offset[tile + 1], cap);
    const uint32_t n = rend - rbeg;
    uint32_t* __restrict__ plist = ws.point_list + rbeg;
    uint32_t sorted_count = 0, next = 0, kept = 0;   // next = index of the next 256-batch to composite
    bool done = !inside;
    float max_contrib = 0.0f;   // STAT_LWMC
    const float cut = ws.hdr->cam.falloff_cut;   // -4.5 (SUM/forward.cu:378); -inf for FOVGS_PS1_VANILLA
    uint8_t* __restrict__ wl = sm.widx[warp];
    lazy_for_each_group(sm, ws, tile, [&](const uint64_t* sk, uint32_t m, bool last) {
        for (uint32_t i = tid; i < m; i += 256) plist[sorted_count + i] = (uint32_t)sk[i];
        sorted_count += m;
        __syncthreads();   // ids visible to the CTA; sk / sort scratch no longer needed
        while ((next + 1) * 256u <= sorted_count || (last && next * 256u < n)) {
            if (__syncthreads_count(done) == 256) return true;
            const uint32_t b0 = next * 256u;
            const int lim = (int)min(256u, n - b0);
            if (tid < lim) {
                const uint32_t id = plist[b0 + tid];
                const float4* __restrict__ rec = ws.rec + (size_t)REC_PS1 * id;
                sm.bl[0].sA[tid] = rec[0]; sm.bl[0].sB[tid] = rec[1]; sm.bl[0].sC[tid] = rec[2];
                sx.ids[tid] = (int)id;
                if (STAT != STAT_MAX) atomicAdd(&in.gaussians_count[id], 1);   // counted when the batch is staged, as in the reference
            }
            sx.acc[tid] = 0.0f;
            if (STAT == STAT_MAX) sx.cnt[tid] = 0;
            __syncthreads();
            next++;
            if (!__all_sync(0xffffffffu, done)) {
                uint32_t cnt = 0;
                for (int jb = 0; jb < lim; jb += 32) {
                    const int j = jb + lane;
                    bool keep = false;
                    // MAX counts entries that pass the falloff cut whatever their alpha: opacity 1 leaves only the -4.5 bound
                    if (j < lim) keep = block_may_touch(sm.bl[0].sA[j], sm.bl[0].sB[j].x, STAT == STAT_MAX ? 1.0f : sm.bl[0].sB[j].y, blkx, blky, -cut);
                    const unsigned mk = __ballot_sync(0xffffffffu, keep);
                    if (keep) wl[cnt + __popc(mk & ((1u << lane) - 1u))] = (uint8_t)j;
                    cnt += __popc(mk);
                }
                __syncwarp();
                kept += cnt;
                // one splat of the warp's list against this lane's pixel; returns alpha*T of the hit (0 if none)
                auto composite = [&](const int j, const float power, bool& in_cut) -> float {
                    float w = 0.0f;
                    in_cut = false;
                    if (!done && !(power > 0.0f || power < cut)) {
                        in_cut = true;
                        // the two rarer outcomes (alpha below 1/255, pixel saturated) are selects, not branches
                        const float4 c = sm.bl[0].sC[j];
                        const float alpha = fminf(0.99f, FM(sm.bl[0].sB[j].y, BLEND_EXP(power)));
                        const float test_T = FM(T, FS(1.0f, alpha));
                        const bool vis = !(alpha < 1.0f / 255.0f);
                        const bool fin = vis && test_T < 0.0001f;
                        const bool acc = vis && !fin;
                        // SUM accumulates (f*alpha)*T and records alpha*T per Gaussian (SUM/forward.cu:400-404)
                        w = acc ? FM(alpha, T) : 0.0f;
                        C0 = acc ? FF(T, FM(alpha, c.x), C0) : C0;
                        C1 = acc ? FF(T, FM(alpha, c.y), C1) : C1;
                        C2 = acc ? FF(T, FM(alpha, c.z), C2) : C2;
                        T = acc ? test_T : T;
                        last_contributor = acc ? b0 + (uint32_t)j + 1u : last_contributor;
                        done = fin;
                    }
                    return w;
                };
                auto record = [&](const int j, const float w, const bool in_cut) {
                    if (STAT == STAT_SUM) {
                        // sum of alpha*T over the warp's 32 pixels with two REDUX instead of a 5-step shuffle tree per splat:
                        // block floating point — the largest term fixes a power-of-two scale that puts it just below 2^26,
                        // the 32 scaled terms are rounded to integers and summed exactly (< 2^31).  Error <= 2^-26 of the
                        // largest term per addend, i.e. fp32-summation accuracy (the reference's atomics are no better).
                        const unsigned wb = __reduce_max_sync(0xffffffffu, __float_as_uint(w));
                        if (wb) {
                            const int e = (int)(wb >> 23) - 127;                       // largest term in [2^e, 2^(e+1))
                            const float up = __uint_as_float((unsigned)(127 + 25 - e) << 23);
                            const unsigned si = __reduce_add_sync(0xffffffffu, __float2uint_rn(w * up));
                            if (lane == 0) atomicAdd(&sx.acc[j], (float)si * __uint_as_float((unsigned)(127 - 25 + e) << 23));
                        }
                    } else if (STAT == STAT_MAX) {
                        const unsigned hits = __ballot_sync(0xffffffffu, in_cut);
                        const unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(w));   // w >= 0: bit order = value order
                        if (lane == 0) {
                            if (hits) atomicAdd(&sx.cnt[j], __popc(hits));
                            if (wmax) atomicMax(reinterpret_cast<unsigned*>(&sx.acc[j]), wmax);
                        }
                    } else {
                        if (w > max_contrib) { max_contrib = w; max_idx = sx.ids[j]; }
                    }
                };
                auto power_of = [&](const int j) {
                    const float4 a = sm.bl[0].sA[j];
                    return gauss_power(a.z, a.w, sm.bl[0].sB[j].x, FS(a.x, pixx), FS(a.y, pixy));
                };
                // four falloff exponents in flight (independent of T), then applied in list order
                const uint32_t* __restrict__ wl4 = reinterpret_cast<const uint32_t*>(wl);
                uint32_t k = 0;
                for (; k + 3 < cnt; k += 4) {
                    if (__all_sync(0xffffffffu, done)) break;
                    const uint32_t q = wl4[k >> 2];
                    const int j0 = q & 0xff, j1 = (q >> 8) & 0xff, j2 = (q >> 16) & 0xff, j3 = q >> 24;
                    const float p0 = power_of(j0), p1 = power_of(j1), p2 = power_of(j2), p3 = power_of(j3);
                    bool c0, c1, c2, c3;
                    const float w0 = composite(j0, p0, c0); record(j0, w0, c0);
                    const float w1 = composite(j1, p1, c1); record(j1, w1, c1);
                    const float w2 = composite(j2, p2, c2); record(j2, w2, c2);
                    const float w3 = composite(j3, p3, c3); record(j3, w3, c3);
                }
                for (; k < cnt; k++) {
                    if (__all_sync(0xffffffffu, done)) break;
                    const int j = wl[k];
                    bool c;
                    const float w = composite(j, power_of(j), c);
                    record(j, w, c);
                }
            }
            __syncthreads();
            if (STAT == STAT_SUM) {
                if (tid < lim && sx.acc[tid] != 0.0f) atomicAdd(&in.contributions[sx.ids[tid]], sx.acc[tid]);
            } else if (STAT == STAT_MAX) {
                if (tid < lim) {
                    if (sx.cnt[tid]) atomicAdd(&in.gaussians_count[sx.ids[tid]], sx.cnt[tid]);
                    const unsigned m = __float_as_uint(sx.acc[tid]);
                    if (m) atomicMax(reinterpret_cast<unsigned*>(&in.contributions[sx.ids[tid]]), m);
                }
            }
        }
        return false;
    });
    if (tid == 0 && next) atomicAdd(&ws.hdr->stats.reserved[0], min(n, next * 256u));
    if (lane == 0 && kept) atomicAdd(&ws.hdr->stats.reserved[1], kept);
}

// ---- one tile: per-pixel state, the lazy producer, the epilogue -------------------------------------------------------
// BK (foveated modes): 0 = the tile is a plain tile, 1 = a blending tile.  It is a COMPILE-TIME parameter: plain tiles
// (83 % of a frame) are compiled without the two-level compositing state (PixFovBlend: 12 words against PixFov's 5), the
// reference makes the same split (renderCUDA at 32 registers, renderCUDA_blending at 40: FOV/cuda_rasterizer/forward.cu:490-609
// vs :262-476).
template <int MODE, int STAT, int BK>
__device__ __forceinline__ void lazy_one_tile(LazySmem& sm, unsigned char* smem_raw, const Workspace& ws, const FrameInputs& in,
                                              const int tile) {
    const FrameHeader* __restrict__ hdr = ws.hdr;
    const int W = hdr->cam.W, H = hdr->cam.H, gx = hdr->cam.grid_x;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x;
#ifdef FOVGS_TILE_TIMING
    // measurement build only: (start, end) of every tile's CTA in ns + SM id, read back with fovgs_debug_tile_times()
    struct TT { unsigned long long t0; int tile; uint32_t n; unsigned sm;
                __device__ ~TT() { if (threadIdx.x == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    g_tile_times[4 * tile + 0] = (uint32_t)t0; g_tile_times[4 * tile + 1] = (uint32_t)t1;
                    g_tile_times[4 * tile + 2] = sm; g_tile_times[4 * tile + 3] = n; } } };
    unsigned long long tt0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt0));
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    TT tt_guard{tt0, tile, min(ws.tile_offset[tile + 1], hdr->cap) - min(ws.tile_offset[tile], hdr->cap), smid};
#endif
    const int bx = ((tid >> 5) & 1) * 8, by = (tid >> 6) * 4;             // this warp's 8x4 pixel block inside the tile
    const int lxi = bx + (tid & 7), lyi = by + ((tid >> 3) & 3);
    const int pxi = tx * TILE + lxi, pyi = ty * TILE + lyi;
    const float blkx = (float)(tx * TILE + bx), blky = (float)(ty * TILE + by);
    const bool inside = pxi < W && pyi < H;
    const uint32_t pix_id = (uint32_t)W * pyi + pxi;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const size_t HW = (size_t)H * W;
    if (is_foveated(MODE)) {
        if (MODE == MODE_MMFR && ws.tile_skip[tile]) {
            // not a tile of this level's call: the reference leaves its zero-initialised image untouched here
            if (inside) { in.out_color[pix_id] = 0.f; in.out_color[HW + pix_id] = 0.f; in.out_color[2 * HW + pix_id] = 0.f; }
            return;
        }
        const float tile_level_f = ws.tile_min[tile];
        const int L1 = (int)tile_level_f;
        constexpr int R = rec_size(MODE);
        const int S1 = (MODE == MODE_FOV) ? 2 + L1 : 2;   // SMFR / MMFR: the one shared (opacity, r, g, b) record
        if (MODE == MODE_MMFR && BK) {
            // mmfr_pcheck_obb/cuda_rasterizer/forward.cu:255-418: one composite, weighted by this level's share of the
            // smoothstep; pixels whose level estimate belongs to the other level of the pair do nothing at all
            const float cur_level = hdr->cur_level;
            const float est = FF(FF((float)lxi, ws.tile_gx[tile], FM((float)lyi, ws.tile_gy[tile])), 0.0625f, tile_level_f);
            const int L1i = (int)est;
            const float xr = FM(FS(est, FA((float)L1i, kStartBlendL)), 2.0f);   // (est - (L1 + start_blend)) / blend_width
            PixFov px;
            px.init(inside, hdr->exp_consts);
            if (xr < 0.0f && (float)L1i != cur_level) px.done = true;
            lazy_tile<1, R>(sm, ws, tile, px, pixx, pixy, blkx, blky, S1, 0);
            if (inside) {
                const float bg0 = hdr->bg[0], bg1 = hdr->bg[1], bg2 = hdr->bg[2];
                const float x = fmaxf(0.0f, fminf(1.0f, xr));
                const float m3 = FM(x, FM(x, -3.0f));
                const float nb = FF(x, FM(x, FA(x, x)), m3);
                const float w1 = FA(nb, 1.0f);
                const float used = ((float)L1i == cur_level) ? w1 : FS(1.0f, w1);
                in.out_color[pix_id] = FM(FF(bg0, px.T, px.C0), used);
                in.out_color[HW + pix_id] = FM(FF(bg1, px.T, px.C1), used);
                in.out_color[2 * HW + pix_id] = FM(FF(bg2, px.T, px.C2), used);
            }
        } else if (!BK) {
            PixFov px;
            px.init(inside, hdr->exp_consts);
            lazy_tile<1, R>(sm, ws, tile, px, pixx, pixy, blkx, blky, S1, 0);
            if (inside) {
                const float bg0 = hdr->bg[0], bg1 = hdr->bg[1], bg2 = hdr->bg[2];
                store_rgb(in, pix_id, HW, FF(bg0, px.T, px.C0), FF(bg1, px.T, px.C1), FF(bg2, px.T, px.C2));
            }
        } else {
            const int L2 = L1 + 1;
            const float dxl = (float)lxi, dyl = (float)lyi;
            const float est = FF(FF(dxl, ws.tile_gx[tile], FM(dyl, ws.tile_gy[tile])), 0.0625f, tile_level_f);
            typename std::conditional<MODE == MODE_SMFR, PixSmfrBlend, PixFovBlend>::type px;
            px.init(inside, est, L2, FA(tile_level_f, 1.0f), hdr->exp_consts);
            lazy_tile<(MODE == MODE_SMFR) ? 3 : 2, R>(sm, ws, tile, px, pixx, pixy, blkx, blky, S1, (MODE == MODE_FOV) ? S1 + 1 : S1);
            if (inside) {
                const float bg0 = hdr->bg[0], bg1 = hdr->bg[1], bg2 = hdr->bg[2];
                const float A0 = FF(bg0, px.T1, px.A0), A1 = FF(bg1, px.T1, px.A1), A2 = FF(bg2, px.T1, px.A2);
                const float B0 = FF(bg0, px.T2, px.B0), B1 = FF(bg1, px.T2, px.B1), B2 = FF(bg2, px.T2, px.B2);
                const float v = FS(est, FA((float)L1, kStartBlendL));
                const float x = __saturatef(FA(fabsf(v), fabsf(v)));
                const float m3 = FM(x, FM(x, -3.0f));
                const float nb = FF(x, FM(x, FA(x, x)), m3);
                const float w1 = FA(nb, 1.0f);
                const float w2 = FS(1.0f, w1);
                store_rgb(in, pix_id, HW, FF(A0, w1, FM(B0, w2)), FF(A1, w1, FM(B1, w2)), FF(A2, w1, FM(B2, w2)));
            }
        }
    } else if (MODE == MODE_SUM) {
        // the training family stages through bl[0] only: with two staging buffers its per-batch scratch lives in the second one
        SumSmemExtra& sx = (NSTAGE == 2) ? *reinterpret_cast<SumSmemExtra*>(&sm.bl[NSTAGE - 1])
                                         : *reinterpret_cast<SumSmemExtra*>(smem_raw + sizeof(LazySmem));
        float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
        uint32_t last_contributor = 0;
        int max_idx = 0;
        lazy_tile_sum<STAT>(sm, sx, ws, in, tile, inside, pixx, pixy, blkx, blky, T, C0, C1, C2, last_contributor, max_idx);
        if (inside) {
            const float bg0 = hdr->bg[0], bg1 = hdr->bg[1], bg2 = hdr->bg[2];
            if (STAT == STAT_LWMC) atomicAdd(&in.contributions[max_idx], in.loss_map[pix_id]);
            ws.final_T[pix_id] = T;
            ws.n_contrib[pix_id] = last_contributor;
            in.out_color[pix_id] = FF(bg0, T, C0);
            in.out_color[HW + pix_id] = FF(bg1, T, C1);
            in.out_color[2 * HW + pix_id] = FF(bg2, T, C2);
        }
    } else {
        PixPS1 px;
        px.init(inside, hdr->exp_consts);
        lazy_tile<0, REC_PS1>(sm, ws, tile, px, pixx, pixy, blkx, blky, 0, 0);
        if (inside) {
            const float bg0 = hdr->bg[0], bg1 = hdr->bg[1], bg2 = hdr->bg[2];
            in.out_color[pix_id] = FF(bg0, px.T, px.C0);
            in.out_color[HW + pix_id] = FF(bg1, px.T, px.C1);
            in.out_color[2 * HW + pix_id] = FF(bg2, px.T, px.C2);
        }
    }
}

// ---- the kernel: persistent CTAs draw tiles of ONE kind, heaviest first, from a ticket counter ---------------------------
// A foveated frame launches it twice: BK = 1 (blending tiles) then BK = 0 (plain tiles).  The two launches are independent
// (disjoint tiles, disjoint pixels), so they form a programmatic-dependent-launch pair (fovgs_internal.cuh): the first
// triggers on entry, and the plain-tile CTAs move onto the SMs as the blending-tile CTAs run out of tickets and exit — no
// idle tail between them.  The second grid's CTAs wait before THEY exit, so the stream (next stage, next frame) only
// proceeds once both grids are complete and flushed.  `pdl`: 0 single launch, 1 first of a pair, 2 second of a pair.
#ifndef LAZY_CTAS_BLEND
#define LAZY_CTAS_BLEND 4      // blending-tile instantiations (A/B: 3 CTAs/SM at 80 registers is 3 % slower)
#endif
template <int MODE, int STAT = STAT_SUM, int BK = 0>
__global__ void __launch_bounds__(256, BK ? LAZY_CTAS_BLEND : LAZY_CTAS) k_lazy_blend(Workspace ws, FrameInputs in, int pdl) {
    extern __shared__ __align__(16) unsigned char lazy_smem_raw[];
    LazySmem& sm = *reinterpret_cast<LazySmem*>(lazy_smem_raw);
    if (pdl == 1) pdl_trigger();
#if LAZY_STAGE == 2
    if (threadIdx.x == 0) {
        mbar_init(&sm.bar[0], 256);
        mbar_init(&sm.bar[1], 256);
        sm.stage_parity = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#endif
    FrameHeader* hdr = ws.hdr;
    // tile_order2 = [blending tiles, heaviest first][plain tiles, heaviest first]
    const uint32_t count = hdr->lazy_count[BK];
    const uint32_t base = BK ? 0u : hdr->lazy_count[1];
    // thread 0 always holds the NEXT ticket: it is requested when a tile starts and first looked at when the tile is done, so
    // the L2 round trip of the atomic never sits between two tiles (a CTA over-draws one ticket at the end; harmless)
    uint32_t next = 0;
    if (threadIdx.x == 0) next = atomicAdd(&hdr->lazy_ticket[BK], 1u);
    for (;;) {
        if (threadIdx.x == 0) sm.ticket = next;
        __syncthreads();
        const uint32_t t = sm.ticket;
        if (t >= count) break;
        if (threadIdx.x == 0) next = atomicAdd(&hdr->lazy_ticket[BK], 1u);
        lazy_one_tile<MODE, STAT, BK>(sm, lazy_smem_raw, ws, in, (int)ws.tile_order2[base + t]);
        __syncthreads();      // the tile's shared memory (and sm.ticket) is free again
    }
    if (pdl == 2) pdl_wait();
}

template <class K>
static cudaError_t launch_lazy_kernel(K kernel, int grid, size_t smem, cudaStream_t st, bool dependent, const Workspace& ws,
                                      const FrameInputs& in, int pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (dependent) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kernel, ws, in, pdl);
}

bool g_no_pdl = false;   // fovgs_set_option(FOVGS_OPT_NO_PDL, 1): every kernel of a frame is an ordinary launch (no programmatic dependent launches)

cudaError_t launch_lazy_blend(const Workspace& ws, const FrameInputs& in, int T, Mode mode, cudaStream_t st) {
    static_assert(sizeof(SumSmemExtra) <= sizeof(BlendStage), "SumSmemExtra must fit the second staging buffer");
    const size_t smem = sizeof(LazySmem), smem_sum = sizeof(LazySmem) + (NSTAGE == 2 ? 0 : sizeof(SumSmemExtra));
    static PerDeviceOnce once;
    bool* configured = once.slot();
    if (!*configured) {
        cudaError_t e;
#define LAZY_SET(K, BYTES)                                                                          \
        e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES));     \
        if (e != cudaSuccess) return e;
        LAZY_SET((k_lazy_blend<MODE_FOV, STAT_SUM, 0>), smem) LAZY_SET((k_lazy_blend<MODE_FOV, STAT_SUM, 1>), smem)
        LAZY_SET((k_lazy_blend<MODE_SMFR, STAT_SUM, 0>), smem) LAZY_SET((k_lazy_blend<MODE_SMFR, STAT_SUM, 1>), smem)
        LAZY_SET((k_lazy_blend<MODE_MMFR, STAT_SUM, 0>), smem) LAZY_SET((k_lazy_blend<MODE_MMFR, STAT_SUM, 1>), smem)
        LAZY_SET((k_lazy_blend<MODE_OBB, STAT_SUM, 0>), smem)
        LAZY_SET((k_lazy_blend<MODE_SUM, STAT_SUM, 0>), smem_sum) LAZY_SET((k_lazy_blend<MODE_SUM, STAT_MAX, 0>), smem_sum)
        LAZY_SET((k_lazy_blend<MODE_SUM, STAT_LWMC, 0>), smem_sum)
#undef LAZY_SET
        *configured = true;
    }
    const int sms = device_sm_count();
    const int grid0 = max(1, min(T, sms * LAZY_CTAS)), grid1 = max(1, min(T, sms * LAZY_CTAS_BLEND));
    if (is_foveated(mode)) {
        // blending tiles first (each costs two composites), then the plain tiles as a programmatic dependent launch
        const bool pdl = !g_no_pdl;
        cudaError_t e;
        if (mode == MODE_FOV) e = launch_lazy_kernel(k_lazy_blend<MODE_FOV, STAT_SUM, 1>, grid1, smem, st, false, ws, in, pdl ? 1 : 0);
        else if (mode == MODE_SMFR) e = launch_lazy_kernel(k_lazy_blend<MODE_SMFR, STAT_SUM, 1>, grid1, smem, st, false, ws, in, pdl ? 1 : 0);
        else e = launch_lazy_kernel(k_lazy_blend<MODE_MMFR, STAT_SUM, 1>, grid1, smem, st, false, ws, in, pdl ? 1 : 0);
        if (e != cudaSuccess) return e;
        if (mode == MODE_FOV) e = launch_lazy_kernel(k_lazy_blend<MODE_FOV, STAT_SUM, 0>, grid0, smem, st, pdl, ws, in, pdl ? 2 : 0);
        else if (mode == MODE_SMFR) e = launch_lazy_kernel(k_lazy_blend<MODE_SMFR, STAT_SUM, 0>, grid0, smem, st, pdl, ws, in, pdl ? 2 : 0);
        else e = launch_lazy_kernel(k_lazy_blend<MODE_MMFR, STAT_SUM, 0>, grid0, smem, st, pdl, ws, in, pdl ? 2 : 0);
        return e != cudaSuccess ? e : cudaGetLastError();
    }
    cudaError_t e;
    if (mode == MODE_SUM && in.stat == STAT_MAX) e = launch_lazy_kernel(k_lazy_blend<MODE_SUM, STAT_MAX, 0>, grid0, smem_sum, st, false, ws, in, 0);
    else if (mode == MODE_SUM && in.stat == STAT_LWMC) e = launch_lazy_kernel(k_lazy_blend<MODE_SUM, STAT_LWMC, 0>, grid0, smem_sum, st, false, ws, in, 0);
    else if (mode == MODE_SUM) e = launch_lazy_kernel(k_lazy_blend<MODE_SUM, STAT_SUM, 0>, grid0, smem_sum, st, false, ws, in, 0);
    else e = launch_lazy_kernel(k_lazy_blend<MODE_OBB, STAT_SUM, 0>, grid0, smem, st, false, ws, in, 0);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace fovgs

// ---- test support: blend_exp against libdevice expf, bit for bit, over a range of float bit patterns ---------------------
namespace fovgs {
__global__ void k_expf_check(uint32_t lo, uint32_t hi, unsigned long long* mismatches, const float* consts) {
    BlendExpConsts ek;
    ek.load(consts);
    unsigned long long bad = 0;
    const uint64_t n = (uint64_t)hi - lo + 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __uint_as_float(lo + (uint32_t)i);
        bad += __float_as_uint(expf(x)) != __float_as_uint(blend_exp(x, ek));
    }
    if (bad) atomicAdd(mismatches, bad);
}
}  // namespace fovgs
extern "C" int fovgs_debug_expf_mismatches(uint32_t lo_bits, uint32_t hi_bits, unsigned long long* mismatches_dev, void* stream) {
    if (!mismatches_dev || hi_bits < lo_bits) return FOVGS_ERR_INVALID_ARG;
    // the two constants travel through device memory (one word after the counter's 8 bytes is not available: use a static buffer)
    static float* consts = nullptr;
    if (!consts) {
        const uint32_t h[2] = {0x3bbb989du, 0x437c0000u};
        if (cudaMalloc(&consts, 8) != cudaSuccess || cudaMemcpy(consts, h, 8, cudaMemcpyHostToDevice) != cudaSuccess) return FOVGS_ERR_CUDA;
    }
    fovgs::k_expf_check<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(lo_bits, hi_bits, mismatches_dev, consts);
    return cudaGetLastError() == cudaSuccess ? 0 : FOVGS_ERR_CUDA;
}

#ifdef FOVGS_TILE_TIMING
extern "C" int fovgs_debug_tile_times(uint32_t* host_out, int tiles) {
    return (int)cudaMemcpyFromSymbol(host_out, fovgs::g_tile_times, sizeof(uint32_t) * 4 * (size_t)tiles);
}
#endif
