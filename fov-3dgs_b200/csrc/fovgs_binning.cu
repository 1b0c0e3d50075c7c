// fovgs_binning.cu — per-Gaussian stage of the forward path, written for load balance.
//
//   k_pre      persistent, warp-autonomous (one block barrier at start, one at the end).  Warps draw 64-Gaussian work
//              tickets from a global counter.  Phase 0 (lane = Gaussian): a conservative screen cull drops the Gaussians
//              whose tile rectangle must be empty; survivors are compacted through a per-warp ring.  Phase A (lane =
//              surviving Gaussian): exact projection with pinned arithmetic (fovgs_math.cuh); the candidate rectangle is
//              clipped to the level's tile bounding box and to the OBB's tile-axis band.  Phase B (lane = candidate tile):
//              WARP-LEVEL LOAD-BALANCED EXPANSION — the 32 rectangles form one candidate index space (shuffle scan), 32
//              candidates per round find their owner through a `redux.or` head mask, then run the exact level and OBB
//              tests; a splat that covers 2000 tiles costs 63 warp rounds instead of stalling one lane for 2000 serial
//              iterations (the reference's `filter`/`OBB_test` and `duplicateWithKeys`, FOV/cuda_rasterizer/
//              rasterizer_impl.cu:264-383, 423-486, are thread-per-Gaussian loops: ~58 % of its frame at 6 M Gaussians,
//              profiles/r1_launches_fov_6M_reference.csv).  Surviving (tile, depth|id) instances are counted per tile (RED
//              on 256-byte-strided counters) and staged densely in per-warp 512-slot chunks.  Phase C: radii, geometry
//              records, visible list.  Replaces preprocessCUDA + InclusiveSum x2 + filter/OBB_test + duplicateWithKeys.
//   k_tile_scan  exclusive scan of the tile histogram (= the reference's `ranges`, no identifyTileRanges) + tile orders, one
//              CTA; the colour kernel is its programmatic dependent launch and runs beside it (nothing there needs the scan).
//   k_color_tma  SH colours of the visible Gaussians (TMA bulk gathers; packed 256-byte model rows for foveated models);
//              k_color is the register-staged fallback.  Replaces compute_fov_colors / computeColorFromSH.
//   k_scatter  staged instances -> their tiles' segments (cursor atomics; trivially balanced: one thread per instance).
#include "fovgs_internal.cuh"
#include "fovgs_tma.cuh"

namespace fovgs {

__device__ constexpr float SH_C0 = 0.28209479177387814f;
__device__ constexpr float SH_C1 = 0.4886025119029199f;
__device__ constexpr float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                       -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                       0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                       -0.5900435899266435f};

// SH colour (OBB/rasterizer_impl.cu:32-82, SUM/forward.cu:20-71, FOV/rasterizer_impl.cu:37-84).
// `first` = index of the first degree-1 triplet (1 for [P,16,3] tensors, 0 for the FOV "rest" tensor).
__device__ __forceinline__ float3 sh_accumulate(const float* __restrict__ sh, int first, int deg, float x, float y, float z,
                                                float3 init) {
    float3 r = init;
    auto C = [&](int k, int ch) { return sh[3 * (first + k) + ch]; };
    if (deg > 0) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            float v = (&r.x)[ch];
            v = v - SH_C1 * y * C(0, ch) + SH_C1 * z * C(1, ch) - SH_C1 * x * C(2, ch);
            (&r.x)[ch] = v;
        }
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z;
            const float xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                float v = (&r.x)[ch];
                v = v + SH_C2[0] * xy * C(3, ch) + SH_C2[1] * yz * C(4, ch) + SH_C2[2] * (2.0f * zz - xx - yy) * C(5, ch) +
                    SH_C2[3] * xz * C(6, ch) + SH_C2[4] * (xx - yy) * C(7, ch);
                (&r.x)[ch] = v;
            }
            if (deg > 2) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    float v = (&r.x)[ch];
                    v = v + SH_C3[0] * y * (3.0f * xx - yy) * C(8, ch) + SH_C3[1] * xy * z * C(9, ch) +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * C(10, ch) +
                        SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * C(11, ch) +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * C(12, ch) + SH_C3[5] * z * (xx - yy) * C(13, ch) +
                        SH_C3[6] * x * (xx - 3.0f * yy) * C(14, ch);
                    (&r.x)[ch] = v;
                }
            }
        }
    }
    return r;
}

// precomputed-covariance variant of project_splat (FOV/forward.cu:155-158)
__device__ __forceinline__ bool project_splat_cov(const CamParams& cam, float mx, float my, float mz, const float* c3, Splat& s) {
    const float tz = xform_row(cam.view, 2, mx, my, mz);
    if (tz <= 0.2f) return false;
    const float hx = xform_row(cam.proj, 0, mx, my, mz);
    const float hy = xform_row(cam.proj, 1, mx, my, mz);
    const float hw = xform_row(cam.proj, 3, mx, my, mz);
    const float pw = __frcp_rn(FA(hw, 0.0000001f));
    const float tx = xform_row(cam.view, 0, mx, my, mz);
    const float ty = xform_row(cam.view, 1, mx, my, mz);
    cov2d_from_cov3d(cam, tx, ty, tz, c3, s.cxx, s.cxy, s.cyy);
    const float bb = FM(s.cxy, s.cxy);
    const float det = FF(s.cxx, s.cyy, -bb);
    if (det == 0.0f) return false;
    const float det_inv = __frcp_rn(det);
    s.conx = FM(s.cyy, det_inv);
    s.cony = FM(s.cxy, -det_inv);
    s.conz = FM(s.cxx, det_inv);
    const float mid = FM(FA(s.cxx, s.cyy), 0.5f);
    const float sq = __fsqrt_rn(fmaxf(FF(mid, mid, -det), 0.1f));
    const float l1 = FA(mid, sq), l2 = FS(mid, sq);
    s.radius = __float2int_ru(FM(__fsqrt_rn(fmaxf(l1, l2)), 3.0f));
    s.px = ndc2pix(FM(hx, pw), cam.W);
    s.py = ndc2pix(FM(hy, pw), cam.H);
    get_rect(s.px, s.py, s.radius, cam.grid_x, cam.grid_y, s.x0, s.y0, s.x1, s.y1);
    const unsigned tnum = (unsigned)(s.y1 - s.y0) * (unsigned)(s.x1 - s.x0);
    if (tnum == 0) return false;
    s.depth = tz;
    s.e1x = s.e1y = s.e2x = s.e2y = s.len1 = s.len2 = 0.0f;
    if (tnum > 1) {
        const float a1 = FS(s.cxx, l1), a2 = FS(s.cxx, l2);
        const float q1 = rsqrtf(FF(a1, a1, bb)), q2 = rsqrtf(FF(a2, a2, bb));
        s.e1x = FM(s.cxy, -q1); s.e1y = FM(a1, q1);
        s.e2x = FM(s.cxy, -q2); s.e2y = FM(a2, q2);
        s.len1 = FM(__fsqrt_rn(l1), 3.0f);
        s.len2 = FM(__fsqrt_rn(l2), 3.0f);
    }
    return true;
}

#ifndef PRE_PB
#define PRE_PB 256
#endif
#ifndef PRE_CTAS
#define PRE_CTAS 4
#endif
#ifndef PRE_WCHUNK
#define PRE_WCHUNK 512
#endif
#ifndef PRE_TICKET
#define PRE_TICKET 64
#endif
#ifndef PRE_UNROLL
#define PRE_UNROLL 2
#endif
constexpr int PRE_UNROLL_N = PRE_UNROLL;   // unroll factor of phase B's round loop
#ifndef SCAT_U
#define SCAT_U 8
#endif
#ifndef SCAT_CTAS
#define SCAT_CTAS 8
#endif
constexpr int PB = PRE_PB;         // threads per block
constexpr int WPB = PB / 32;       // warps per block; every warp is an autonomous worker (no block barriers)
constexpr uint32_t WCHUNK = PRE_WCHUNK;   // staging slots a warp reserves per global atomic
constexpr uint32_t VCHUNK = 128;   // visible-list slots a warp reserves per global atomic
constexpr int SHW = 65;            // k_color: floats per staged Gaussian (64 + 1 pad: conflict-free lane-strided reads)

// Per-owner data of phase B, packed so that a candidate lane fetches its owner's parameters with five 16-byte shared loads
// instead of ~20 scalar ones (lanes of one owner broadcast; phase B is instruction-bound).
struct WarpSmem {
    float4 oc[32];               // px, py, e1x, e1y
    float4 oe[32];               // e2x, e2y, g1, g2:  g_k = len_k + 8 (|e_kx| + |e_ky|)  (reach of tile + OBB along eigen-axis k)
    float4 ox[32];               // extents of the OBB's corners (tile-axis separations of the OBB test): mnx, mxx, mny, mxy
    int4 orc[32];                // candidate rectangle: first candidate index (exclusive scan), width, x0, y0
    uint4 ok[32];                // 1/width (float bits), flags (hcode | single << 8), depth bits, Gaussian id
    float l1[32], l2[32], hl1[32];   // only the exact OBB test / the non-integral level test read these
    uint32_t queue[64];          // ids that survived the conservative screen cull, waiting for a full warp of work
    uint32_t nzlist[32];         // lane of the k-th Gaussian with a non-empty candidate rectangle
    uint32_t cnt[32];            // != 0: at least one candidate tile survived (the Gaussian is visible)
};

constexpr int LC_MAX = 16384;      // tiles whose level code fits the shared table (1080p: 8160)
constexpr int PRE_CHUNK = PRE_TICKET;   // Gaussians per work ticket of k_pre
static_assert(PRE_CHUNK % 32 == 0 && PRE_CHUNK <= 256, "ticket = whole batches; its inputs are prefetched by one warp");

struct PreSmem {
    CamParams cam;
    float cull_k;                // bound on |T row| * tz for the conservative screen cull (see k_pre phase 0)
    int bbox[FOV_LEVELS][4];
    // FOV: per tile the smallest h in {1,2,3,4} with tile_min < h (5: none).  A Gaussian with integral highest_level passes
    // `tile_min < highest_level + 1` (rasterizer_impl.cu:307,344) iff highest_level + 1 >= code: the per-candidate level
    // test reads one shared byte instead of gathering a float through L1.
    alignas(16) uint8_t lvl_code[LC_MAX];
    WarpSmem w[WPB];
};

template <int MODE>
__global__ void __launch_bounds__(PB, PRE_CTAS) k_pre(Workspace ws, FrameInputs in) {
    extern __shared__ __align__(16) unsigned char pre_smem_raw[];      // sizeof(PreSmem) > 48 KB: opt-in dynamic shared memory
    PreSmem& sm = *reinterpret_cast<PreSmem*>(pre_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const int n = (int)(sizeof(CamParams) / 4);
        const uint32_t* src = (const uint32_t*)&ws.hdr->cam;
        uint32_t* dst = (uint32_t*)&sm.cam;
        for (int i = tid; i < n; i += PB) dst[i] = src[i];
        if (is_foveated(MODE) && tid < FOV_LEVELS * 4) (&sm.bbox[0][0])[tid] = (&ws.hdr->lvl_bbox[0][0])[tid];
        if (tid == 32) {
            // |row i of T| <= Wn * f * (1 + 1.3 tanfov) / tz  (cov2d_from_cov3d: T_i = J_ii * view_axis_i + J_i2 * view_axis_z,
            // |J_i2| <= f * 1.3 tanfov / tz); Wn = largest norm of the three view axes (1 for a rigid camera, not assumed)
            const float* v = ws.hdr->cam.view;
            const float n0 = sqrtf(v[0] * v[0] + v[4] * v[4] + v[8] * v[8]);
            const float n1 = sqrtf(v[1] * v[1] + v[5] * v[5] + v[9] * v[9]);
            const float n2 = sqrtf(v[2] * v[2] + v[6] * v[6] + v[10] * v[10]);
            const float fxk = ws.hdr->cam.focal_x * (1.0f + 1.3f * ws.hdr->cam.tanfovx);
            const float fyk = ws.hdr->cam.focal_y * (1.0f + 1.3f * ws.hdr->cam.tanfovy);
            sm.cull_k = fmaxf(n0, fmaxf(n1, n2)) * fmaxf(fxk, fyk) * 1.001f;
        }
        if (is_foveated(MODE)) {
            // per-tile level codes (written by k_tile_infos): 16 bytes per load, two independent loads per thread at 1080p
            const int T = ws.hdr->tiles;
            if (T <= LC_MAX)
                for (int i = tid * 16; i < T; i += PB * 16)
                    *reinterpret_cast<uint4*>(&sm.lvl_code[i]) = *reinterpret_cast<const uint4*>(&ws.tile_code[i]);
        }
    }
    __syncthreads();   // the only block barrier: from here on warps never wait for each other
    const CamParams& cam = sm.cam;
    WarpSmem& wm = sm.w[warp];
    const int gx = cam.grid_x;
    const uint32_t stage_cap = ws.stage_cap;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned le_mask = lt_mask | (1u << lane);
    // staging chunk of this warp (identical in all lanes)
    uint32_t chunk_base = 0, chunk_used = WCHUNK;
    bool has_chunk = false;
    uint32_t vbase = 0, vused = VCHUNK;   // this warp's chunk of the visible list
    bool has_vchunk = false;
    unsigned visible_total = 0, cand_total = 0;

    uint32_t qn = 0;   // entries waiting in this warp's queue (warp-uniform)
    const float cull_k = sm.cull_k;
    const float scr_w = (float)(16 * cam.grid_x), scr_h = (float)(16 * cam.grid_y);
    const bool use_codes = is_foveated(MODE) && ws.hdr->tiles <= LC_MAX;
    // Dynamic work distribution: warps draw 64-Gaussian chunks (PRE_TICKET) from a global ticket counter (splats that cover hundreds
    // of tiles are rare and random, a static partition leaves a 12 % tail).  The next ticket is requested one chunk ahead;
    // its value is only looked at when the current chunk is used up.
    const uint32_t nchunks = ((uint32_t)in.P + PRE_CHUNK - 1) / PRE_CHUNK;
    auto request_ticket = [&]() -> uint32_t { return lane == 0 ? atomicAdd(&ws.hdr->pre_chunk, 1u) : 0u; };
    uint32_t cur = __shfl_sync(0xffffffffu, request_ticket(), 0);
    uint32_t nxt_raw = request_ticket();
    int batch = 0;
    for (;;) {
        const bool have_input = cur < nchunks;
        const int base0 = have_input ? (int)(cur * PRE_CHUNK + batch * 32) : in.P;
        // ---------------- phase 0: conservative screen cull + compaction (lane = Gaussian) ----------------
        // Four in five Gaussians of a frame end with an empty tile rectangle.  The reference finds that out at the end of
        // the full projection; here a bound on the radius decides it first:  lambda_max(Sigma2D) <= 2 (k/tz)^2
        // lambda_max(Sigma3D) + 0.62  (|T row| <= k/tz; lambda_1 = mid + sqrt(max(0.1, ((cxx-cyy)/2)^2 + cxy^2)) <=
        // max(cxx,cyy) + |cxy| + 0.317), lambda_max(Sigma3D) <= (mod * s_max * (|1-|q|^2| + |q|^2))^2  (R(q) = (1-|q|^2) I +
        // |q|^2 R(q/|q|)).  A Gaussian whose mean lies further outside the tile grid than that radius (plus slack for every
        // rounding involved) has an empty rectangle in the exact code as well — it is dropped here, radii = 0, exactly
        // the reference's outcome; everything else (and anything non-finite) goes on to the exact projection.  Survivors
        // are compacted through a per-warp queue, so phases A-C always run on a full warp of live Gaussians.
        if (base0 < in.P) {
            const int idx0 = base0 + lane;
            bool keep = false;
            if (idx0 < in.P) {
                // all inputs of the bound are requested before the first one is used: one memory round trip, not two
                const float mx = in.means3D[3 * (size_t)idx0], my = in.means3D[3 * (size_t)idx0 + 1], mz = in.means3D[3 * (size_t)idx0 + 2];
                float sx = 0.f, sy = 0.f, sz = 0.f;
                float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
                if (in.cov3D_precomp == nullptr) {
                    sx = in.scales[3 * (size_t)idx0]; sy = in.scales[3 * (size_t)idx0 + 1]; sz = in.scales[3 * (size_t)idx0 + 2];
                    q = *reinterpret_cast<const float4*>(in.rotations + 4 * (size_t)idx0);
                }
                // Foveated modes: a Gaussian only lands in tiles of its level's region {tile_min < highest_level + 1}, so the
                // "tile grid" it is culled against is that region's tile bounding box, not the screen: three in five Gaussians
                // carry highest level 0, whose region is the foveal disc — outside it they used to survive this cull, pay the
                // exact projection (~450 instructions) and only then lose their whole rectangle to the same box in phase A.
                float lo_x = 0.0f, hi_x = scr_w, lo_y = 0.0f, hi_y = scr_h;
                if (is_foveated(MODE)) {
                    const float hl0 = has_level_mask(MODE) ? in.highest_levels[idx0] : 0.0f;   // MMFR: box 0 = this level's tiles
                    const int li = (int)hl0;
                    if (hl0 >= 0.0f && hl0 <= (float)(FOV_LEVELS - 1) && (float)li == hl0) {   // same condition as phase A's clip
                        lo_x = 16.0f * (float)sm.bbox[li][0]; lo_y = 16.0f * (float)sm.bbox[li][1];
                        hi_x = 16.0f * (float)sm.bbox[li][2]; hi_y = 16.0f * (float)sm.bbox[li][3];
                    }
                }
                const float tz = xform_row(cam.view, 2, mx, my, mz);
                if (!(tz <= 0.2f)) {                  // same expression, same bits as the exact near-plane test (NaN passes)
                    float lam3;                       // bound on the largest eigenvalue of Sigma3D
                    if (in.cov3D_precomp != nullptr) {
                        const float* c = in.cov3D_precomp + 6 * (size_t)idx0;
                        lam3 = sqrtf(c[0] * c[0] + c[3] * c[3] + c[5] * c[5] + 2.0f * (c[1] * c[1] + c[2] * c[2] + c[4] * c[4]));
                    } else {
                        const float qn2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
                        const float rn = fabsf(1.0f - qn2) + qn2;
                        const float sm_ = fmaxf(fabsf(sx), fmaxf(fabsf(sy), fabsf(sz))) * fabsf(cam.scale_modifier) * rn;
                        lam3 = sm_ * sm_;
                    }
                    // MUFU-approximate reciprocal / square root (1-2 ulp): this is the CONSERVATIVE bound, its 1 % and 1e-5 slack
                    // factors dwarf them; the exact projection below keeps its IEEE operations
                    float rtz, pw, sq2;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rtz) : "f"(tz));
                    const float kt = cull_k * rtz;
                    const float lam2 = 2.0f * kt * kt * lam3 * 1.01f + 0.62f;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq2) : "f"(lam2));
                    const float rb = 3.01f * sq2 + 2.0f;
                    const float hx = xform_row(cam.proj, 0, mx, my, mz);
                    const float hy = xform_row(cam.proj, 1, mx, my, mz);
                    const float hw = xform_row(cam.proj, 3, mx, my, mz);
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pw) : "f"(FA(hw, 0.0000001f)));
                    const float px = ((FM(hx, pw) + 1.0f) * (float)cam.W - 1.0f) * 0.5f;
                    const float py = ((FM(hy, pw) + 1.0f) * (float)cam.H - 1.0f) * 0.5f;
                    const float ex = rb + 1.0f + 1e-5f * fabsf(px), ey = rb + 1.0f + 1e-5f * fabsf(py);
                    // exact rect: x1 = clamp(trunc((px + r + 15)/16)), x0 = clamp(trunc((px - r)/16)); empty iff x1 <= x0 or y1 <= y0
                    // (against tile columns [bx0, bx1): x1 <= bx0 or x0 >= bx1; the screen is the box [0, grid))
                    const bool out = (px + ex + 16.0f < lo_x) || (px - ex > hi_x + 1.0f) || (py + ey + 16.0f < lo_y) || (py - ey > hi_y + 1.0f);
                    keep = !out;                      // NaN anywhere: comparisons are false -> keep
                }
                if (!keep) in.radii[idx0] = 0;
                // the reference prints and __trap()s (fatal for the context) when `prefiltered` is set and a Gaussian fails the
                // near-plane test (auxiliary.h:286-293); here the frame completes and the count travels in the statistics
                if (cam.prefiltered && tz <= 0.2f) atomicAdd(&ws.hdr->stats.reserved[3], 1u);
            }
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            if (keep) wm.queue[qn + __popc(km & lt_mask)] = (uint32_t)idx0;
            qn += __popc(km);
            __syncwarp();
        }
        if (have_input && ++batch == PRE_CHUNK / 32) {
            batch = 0;
            cur = __shfl_sync(0xffffffffu, nxt_raw, 0);
            nxt_raw = request_ticket();
            if (cur + 1 < nchunks) {   // full chunks only: the prefetched lines must lie inside the tensors
                // the new chunk's inputs start their way up from HBM now (means 12 lines, scales 12, rotations 16)
                const size_t g0 = (size_t)cur * PRE_CHUNK;
                const char* pm = (const char*)(in.means3D + 3 * g0) + 128 * lane;
                if (lane < PRE_CHUNK * 12 / 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pm));
                if (has_level_mask(MODE)) {
                    const char* pl = (const char*)(in.highest_levels + g0) + 128 * lane;
                    if (lane < PRE_CHUNK * 4 / 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pl));
                }
                if (in.cov3D_precomp == nullptr) {
                    const char* ps = (const char*)(in.scales + 3 * g0) + 128 * lane;
                    const char* pr = (const char*)(in.rotations + 4 * g0) + 128 * lane;
                    if (lane < PRE_CHUNK * 12 / 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(ps));
                    if (lane < PRE_CHUNK * 16 / 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr));
                }
            }
        }
        const bool input_done = !(cur < nchunks);
        if (qn < 32 && !(input_done && qn > 0)) {
            if (input_done) break;
            continue;
        }
        const bool valid_lane = (uint32_t)lane < qn;
        const int idx = valid_lane ? (int)wm.queue[lane] : in.P;
        __syncwarp();
        if (qn > 32) { const uint32_t v = (32 + lane < (int)qn) ? wm.queue[32 + lane] : 0u; __syncwarp(); wm.queue[lane] = v; }
        qn = qn > 32 ? qn - 32 : 0;
        // ---------------- phase A: projection (lane = Gaussian) ----------------
        Splat s;
        float c3[6];
        bool ok = false;
        float hl = 0.0f;
        if (idx < in.P) {
            const float mx = in.means3D[3 * (size_t)idx], my = in.means3D[3 * (size_t)idx + 1], mz = in.means3D[3 * (size_t)idx + 2];
            if (has_level_mask(MODE)) hl = in.highest_levels[idx];
            if (in.cov3D_precomp != nullptr) {
#pragma unroll
                for (int k = 0; k < 6; k++) c3[k] = in.cov3D_precomp[6 * (size_t)idx + k];
                ok = project_splat_cov(cam, mx, my, mz, c3, s);
            } else {
                const float sx = in.scales[3 * (size_t)idx], sy = in.scales[3 * (size_t)idx + 1], sz = in.scales[3 * (size_t)idx + 2];
                const float4 q = *reinterpret_cast<const float4*>(in.rotations + 4 * (size_t)idx);
                ok = project_splat(cam, mx, my, mz, sx, sy, sz, q.x, q.y, q.z, q.w, s, c3);
            }
        }
        uint32_t tnum = 0;
        uint32_t okx = 0, oky = 0, okz = 0;
        int rcw = 0, rcx = 0, rcy = 0;
        if (ok) {
            // "single" = no OBB test for this splat: one-tile rectangles (the reference's potential_tnum == 1 shortcut) and, in
            // the vanilla mode of the training family, every splat (its duplicateWithKeys walks the whole rectangle)
            const bool single0 = ((uint32_t)(s.y1 - s.y0) * (uint32_t)(s.x1 - s.x0)) == 1u || (MODE == MODE_SUM && cam.no_obb);
            int cx0 = s.x0, cy0 = s.y0, cx1 = s.x1, cy1 = s.y1;
            uint32_t hcode = 0;
            if (is_foveated(MODE)) {
                // tiles outside the level's bounding box fail `tile_min < hl + 1` anyway: do not even enumerate them
                const int li = (int)hl;   // MMFR: hl stays 0 -> box 0 = the tiles of this level's call, code 1
                if (hl >= 0.0f && hl <= (float)(FOV_LEVELS - 1) && (float)li == hl) {
                    cx0 = max(cx0, sm.bbox[li][0]); cy0 = max(cy0, sm.bbox[li][1]);
                    cx1 = min(cx1, sm.bbox[li][2]); cy1 = min(cy1, sm.bbox[li][3]);
                    if (use_codes) hcode = (uint32_t)li + 1u;
                }
            }
            float e_mnx = 0.f, e_mxx = 0.f, e_mny = 0.f, e_mxy = 0.f;
            if (!single0 && true) {
                // The OBB test starts with the two tile-axis separations (auxiliary.h:95-118; obb_hits_tile): a tile
                // column tx can only pass when max(vx) - (16 tx + 8) >= -8 and min(vx) - (16 tx + 8) <= 8, same for
                // rows.  Clip the candidate rectangle to that band (widened by far more than one fp32 rounding of the
                // subtraction) so the tiles it removes are exactly tiles the test would reject.
                ObbCorners oc;
                obb_corners(s.px, s.py, s.e1x, s.e1y, s.e2x, s.e2y, s.len1, s.len2, oc);
                const float mnx = fminf(fminf(oc.vx[0], oc.vx[1]), fminf(oc.vx[2], oc.vx[3]));
                const float mxx = fmaxf(fmaxf(oc.vx[0], oc.vx[1]), fmaxf(oc.vx[2], oc.vx[3]));
                const float mny = fminf(fminf(oc.vy[0], oc.vy[1]), fminf(oc.vy[2], oc.vy[3]));
                const float mxy = fmaxf(fmaxf(oc.vy[0], oc.vy[1]), fmaxf(oc.vy[2], oc.vy[3]));
                e_mnx = mnx; e_mxx = mxx; e_mny = mny; e_mxy = mxy;
                if (fabsf(mnx) < 1e8f && fabsf(mxx) < 1e8f) {
                    const float e = 0.05f + 1e-6f * (fabsf(mnx) + fabsf(mxx));
                    cx0 = max(cx0, (int)fmaxf(ceilf((mnx - 16.0f - e) * 0.0625f), 0.0f));
                    cx1 = min(cx1, (int)fminf(floorf((mxx + e) * 0.0625f) + 1.0f, 1e6f));
                }
                if (fabsf(mny) < 1e8f && fabsf(mxy) < 1e8f) {
                    const float e = 0.05f + 1e-6f * (fabsf(mny) + fabsf(mxy));
                    cy0 = max(cy0, (int)fmaxf(ceilf((mny - 16.0f - e) * 0.0625f), 0.0f));
                    cy1 = min(cy1, (int)fminf(floorf((mxy + e) * 0.0625f) + 1.0f, 1e6f));
                }
            }
            const int cw = max(cx1 - cx0, 0), ch = max(cy1 - cy0, 0);
            tnum = (uint32_t)cw * (uint32_t)ch;
            wm.oc[lane] = make_float4(s.px, s.py, s.e1x, s.e1y);
            wm.oe[lane] = make_float4(s.e2x, s.e2y, s.len1 + 8.0f * (fabsf(s.e1x) + fabsf(s.e1y)),
                                      s.len2 + 8.0f * (fabsf(s.e2x) + fabsf(s.e2y)));
            wm.ox[lane] = make_float4(e_mnx, e_mxx, e_mny, e_mxy);
            wm.l1[lane] = s.len1; wm.l2[lane] = s.len2;
            if (is_foveated(MODE)) wm.hl1[lane] = FA(hl, 1.0f);
            if (MODE == MODE_SUM) {
                // kept for the backward pass; only ever read for Gaussians that end up visible
#pragma unroll
                for (int k = 0; k < 6; k++) ws.cov3D[6 * (size_t)idx + k] = c3[k];
            }
            okx = __float_as_uint(1.0f / (float)max(cw, 1));
            // The two tile-axis separations of the OBB test (below, per candidate) are monotone in the tile index — fl(a - t) is
            // non-increasing in t — so when they hold at the candidate rectangle's two extreme columns and rows, evaluated with
            // the very same operations, they hold for every tile in it: flag 512 lets phase B skip them (and their 16-byte
            // shared load) for this splat.  The rectangle was clipped to a slightly widened band, so nearly every splat qualifies.
            bool band_ok = false;
            if (!single0 && cw > 0 && ch > 0) {
                const float tcx_lo = FF((float)cx0, 16.0f, 8.0f), tcx_hi = FF((float)(cx1 - 1), 16.0f, 8.0f);
                const float tcy_lo = FF((float)cy0, 16.0f, 8.0f), tcy_hi = FF((float)(cy1 - 1), 16.0f, 8.0f);
                band_ok = !(FS(e_mxx, tcx_hi) < -8.0f) && !(FS(e_mnx, tcx_lo) > 8.0f) && !(FS(e_mxy, tcy_hi) < -8.0f) && !(FS(e_mny, tcy_lo) > 8.0f);
            }
            oky = hcode | (single0 ? 256u : 0u) | (band_ok ? 512u : 0u);
            okz = __float_as_uint(s.depth);
            rcw = cw; rcx = cx0; rcy = cy0;
        }
        wm.cnt[lane] = 0;
        // exclusive scan of the candidate counts over the warp; compact list of non-empty owners
        uint32_t incl = tnum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        const uint32_t my_start = incl - tnum;
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned nz = __ballot_sync(0xffffffffu, tnum > 0);
        cand_total += total;
        wm.orc[lane] = make_int4((int)my_start, rcw, rcx, rcy);
        wm.ok[lane] = make_uint4(okx, oky, okz, (uint32_t)idx);
        if (tnum > 0) wm.nzlist[__popc(nz & lt_mask)] = (uint32_t)lane;
        __syncwarp();

        // ---------------- phase B: candidate tiles (lane = candidate), 32 per round ----------------
        uint32_t owners_before = 0;   // non-empty owners whose range starts before the current window
#pragma unroll PRE_UNROLL_N   // two rounds share no registers: the staging stores of round r need not drain before round r+1 starts
        for (uint32_t r = 0; r < total; r += 32) {
            const uint32_t c = r + lane;
            const bool valid = c < total;
            // heads: window positions at which a non-empty owner's range starts
            const unsigned hbit = (tnum > 0 && my_start >= r && my_start < r + 32) ? (1u << (my_start - r)) : 0u;
            const unsigned H = __reduce_or_sync(0xffffffffu, hbit);
            bool pass = false;
            uint32_t tile = 0, owner = 0;
            uint4 ko = make_uint4(0u, 0u, 0u, 0u);
            if (valid) {
                owner = wm.nzlist[owners_before + __popc(H & le_mask) - 1];
                const int4 rc = wm.orc[owner];
                ko = wm.ok[owner];
                const int t = (int)c - rc.x;
                const int w = rc.y;
                int q = (int)(((float)t + 0.5f) * __uint_as_float(ko.x));
                int rem = t - q * w;
                if (rem < 0) { q--; rem += w; } else if (rem >= w) { q++; rem -= w; }
                const int ty = rc.w + q;
                const int tx = rc.z + rem;
                tile = (uint32_t)ty * gx + tx;
                pass = true;
                if (is_foveated(MODE)) {
                    const uint32_t hc = ko.y & 0xffu;
                    if (MODE == MODE_MMFR) pass = hc ? (sm.lvl_code[tile] == 1) : (ws.tile_skip[tile] == 0);
                    else pass = hc ? (hc >= (uint32_t)sm.lvl_code[tile]) : (ws.tile_min[tile] < wm.hl1[owner]);
                }
                if (pass && !(ko.y & 256u)) {
                    // OBB separating-axis test (auxiliary.h:80-168).  The two tile-axis separations are evaluated exactly as the
                    // reference does (corner extents per splat, see obb_hits_tile_ext).  The two eigen-axis separations compare
                    // the min / max over the tile's four corners of dot(corner - centre, e_k) with +-len_k; those are affine in
                    // the tile centre: min/max = dot(tile centre - centre, e_k) -+ 8 (|e_kx| + |e_ky|).  So
                    //     t_k = |dot(d, e_k)| - g_k,  g_k = len_k + 8 (|e_kx| + |e_ky|)
                    // decides axis k except within the fp32 rounding of either formulation (a few ulp of |d| terms, bounded by
                    // eps below); only candidates inside that band — about one in 10^4 — run the reference's corner-by-corner
                    // arithmetic, so every decision is the reference's decision, bit for bit.
                    const float4 pc = wm.oc[owner];
                    const float4 pe = wm.oe[owner];
                    const float tcx = FF((float)tx, 16.0f, 8.0f), tcy = FF((float)ty, 16.0f, 8.0f);
                    bool band_fail = false;
                    if (!(ko.y & 512u)) {   // (rare: the splat's rectangle reaches past the exact band, see phase A)
                        const float4 b4 = wm.ox[owner];
                        band_fail = FS(b4.y, tcx) < -8.0f || FS(b4.x, tcx) > 8.0f || FS(b4.w, tcy) < -8.0f || FS(b4.z, tcy) > 8.0f;
                    }
                    if (band_fail) {
                        pass = false;
                    } else {
                        const float dx = tcx - pc.x, dy = tcy - pc.y;
                        const float t1 = fabsf(fmaf(pc.z, dx, pc.w * dy)) - pe.z;
                        const float t2 = fabsf(fmaf(pe.x, dx, pe.y * dy)) - pe.w;
                        const float eps = 4e-6f * (fabsf(dx) + fabsf(dy) + 16.0f);
                        if (t1 > eps || t2 > eps) pass = false;                        // separated along an eigen-axis, surely
                        else if (!(t1 < -eps && t2 < -eps)) {                          // inside the rounding band (or NaN): exact
                            const float4 px4 = wm.ox[owner];
                            pass = obb_hits_tile_ext(px4.x, px4.y, px4.z, px4.w, pc.x, pc.y, pc.z, pc.w, pe.x, pe.y, wm.l1[owner],
                                                     wm.l2[owner], tcx, tcy);
                        }
                    }
                }
                // RED (no return value).  Measured alternative: taking the returned rank here so that the scatter needs no
                // atomics costs k_pre +0.05 ms (even with the dependent store deferred by a round) and saves the scatter
                // 0.02 ms — the scatter is bound by its 8-byte scattered stores, not by its cursor atomics.
                if (pass) atomicAdd(&ws.tile_count[(size_t)tile * CSTRIDE], 1u);
            }
            // per-owner bookkeeping: a Gaussian is visible iff any of its candidate tiles survives (benign same-value race).
            // The reference also tracks the range of levels a Gaussian lands in (rasterizer_impl.cu:375-380) to colour only
            // those; k_color simply colours all four levels — the unused ones are never composited.
            if (pass) wm.cnt[owner] = 1u;
            const unsigned passmask = __ballot_sync(0xffffffffu, pass);
            // stage the surviving instances densely in this warp's chunk
            const uint32_t np = __popc(passmask);
            if (np) {
                if (chunk_used + np > WCHUNK) {
                    if (has_chunk) {
                        const uint32_t p = chunk_base + chunk_used + lane;    // < 32 slots are left
                        if (p < chunk_base + WCHUNK && p < stage_cap) ws.stage_tile[p] = TILE_INVALID;
                    }
                    uint32_t nb = 0;
                    if (lane == 0) nb = atomicAdd(&ws.hdr->stage_cursor, WCHUNK);
                    chunk_base = __shfl_sync(0xffffffffu, nb, 0);
                    chunk_used = 0;
                    has_chunk = true;
                }
                if (pass) {
                    const uint32_t pos = chunk_base + chunk_used + __popc(passmask & lt_mask);
                    if (pos < stage_cap) {
                        ws.stage_tile[pos] = tile;
                        ws.stage_key[pos] = ((uint64_t)ko.z << 32) | ko.w;
                    }
                }
                chunk_used += np;
            }
            owners_before += __popc(H);
            __syncwarp();
        }
        __syncwarp();

        // ---------------- phase C: per-Gaussian outputs; colour work is queued for k_color ----------------
        const uint32_t count = ok ? wm.cnt[lane] : 0u;
        const bool visible = (idx < in.P) && count > 0;
        if (idx < in.P) in.radii[idx] = count ? s.radius : 0;
        const unsigned vismask = __ballot_sync(0xffffffffu, visible);
        const uint32_t nv = __popc(vismask);
        visible_total += nv;
        uint32_t lv = 0;
        if (visible) {
            constexpr int R = rec_size(MODE);
            float4* rec = ws.rec + (size_t)R * idx;
            rec[0] = make_float4(s.px, s.py, s.conx, s.cony);
            if (is_foveated(MODE)) {
                rec[1] = make_float4(s.conz, hl, s.depth, 0.0f);
                lv = (uint32_t)(FOV_LEVELS - 1) << 8;   // all levels
            } else {
                rec[1] = make_float4(s.conz, in.opacities[idx], s.depth, 0.0f);
            }
        }
        if (nv) {
            if (vused + nv > VCHUNK) {
                if (has_vchunk) {
                    const uint32_t p = vbase + vused + lane;              // < 32 slots are left
                    if (p < vbase + VCHUNK && p < ws.vis_cap) ws.vis_list[p] = TILE_INVALID;
                }
                uint32_t nb = 0;
                if (lane == 0) nb = atomicAdd(&ws.hdr->vis_cursor, VCHUNK);
                vbase = __shfl_sync(0xffffffffu, nb, 0);
                vused = 0;
                has_vchunk = true;
            }
            if (visible) {
                const uint32_t p = vbase + vused + __popc(vismask & lt_mask);
                if (p < ws.vis_cap) { ws.vis_list[p] = (uint32_t)idx; ws.vis_lv[p] = lv; }
            }
            vused += nv;
        }
    }
    // retire the partially filled staging / visible-list chunks
    if (has_chunk) {
        for (uint32_t p = chunk_base + chunk_used + lane; p < chunk_base + WCHUNK; p += 32)
            if (p < stage_cap) ws.stage_tile[p] = TILE_INVALID;
    }
    if (has_vchunk) {
        for (uint32_t p = vbase + vused + lane; p < vbase + VCHUNK; p += 32)
            if (p < ws.vis_cap) ws.vis_list[p] = TILE_INVALID;
    }
    if (lane == 0 && visible_total) atomicAdd(&ws.hdr->stats.num_visible, visible_total);
    if (lane == 0 && cand_total) atomicAdd(&ws.hdr->stats.reserved[2], cand_total);   // candidate tiles enumerated
    // (the tile histogram is scanned by k_tile_scan, which runs beside the colour kernel that follows)
}

// ------------------------------------------------------------------------------------------------------------------
// k_color: SH -> RGB for the visible Gaussians, fully packed.  k_pre queued the visible ids densely; here one warp takes
// 32 of them, copies their SH / dc / opacity inputs with coalesced warp-wide loads into shared memory (8 Gaussians =
// 16 loads in flight per lane), then every lane evaluates its own Gaussian from its slot with all 32 lanes busy.
// Replaces compute_fov_colors (FOV/rasterizer_impl.cu:490-530), OBB's compute_fov_colors (:118-138) and the SH part of
// SUM's preprocessCUDA (SUM/forward.cu:20-71, 282-288).  A thread-per-Gaussian SH read costs 61 L1 wavefronts per
// Gaussian (one 4-byte element of 32 different lines per load); the cooperative copy needs 2-4.
// ------------------------------------------------------------------------------------------------------------------
constexpr int CW = 4;   // warps per CTA in k_color
bool g_no_tma = false;  // fovgs_set_option(FOVGS_OPT_NO_TMA, 1): force the register-staged colour kernel

template <int MODE>
__global__ void __launch_bounds__(CW * 32) k_color(Workspace ws, FrameInputs in) {
    __shared__ float shbuf[CW][32][SHW];
    __shared__ float campos_s[3];
    __shared__ int deg_s, M_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 3) campos_s[tid] = ws.hdr->cam.campos[tid];
    if (tid == 0) { deg_s = ws.hdr->cam.sh_degree; M_s = ws.hdr->cam.M; }
    __syncthreads();
    const int deg = deg_s, M = M_s;
    const int nsh = (in.shs != nullptr) ? 3 * M : 0;
    const uint32_t nslots = min(ws.hdr->vis_cursor, ws.vis_cap);
    const unsigned lt_mask = (1u << lane) - 1u;
    float (*buf)[SHW] = shbuf[warp];
    const uint32_t gw = blockIdx.x * CW + warp, nw = gridDim.x * CW;
    for (uint32_t s0 = gw * 32; s0 < nslots; s0 += nw * 32) {
        const uint32_t slot = s0 + lane;
        uint32_t id = TILE_INVALID, lv = 0;
        if (slot < nslots) { id = ws.vis_list[slot]; lv = ws.vis_lv[slot]; }
        const bool valid = id != TILE_INVALID;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (vmask == 0) continue;
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (valid) {
            const float mx = in.means3D[3 * (size_t)id], my = in.means3D[3 * (size_t)id + 1], mz = in.means3D[3 * (size_t)id + 2];
            dx = mx - campos_s[0]; dy = my - campos_s[1]; dz = mz - campos_s[2];
            const float len = sqrtf(dx * dx + dy * dy + dz * dz);
            dx = dx / len; dy = dy / len; dz = dz / len;
        }
        // cooperative copy, 8 Gaussians per round; the id of Gaussian g is broadcast from lane g
#pragma unroll 1
        for (int g0 = 0; g0 < 32; g0 += 8) {
            if (((vmask >> g0) & 0xffu) == 0) continue;
            float a0[8], a1[8];
#pragma unroll
            for (int g = 0; g < 8; g++) {
                a0[g] = 0.f; a1[g] = 0.f;
                const uint32_t gid32 = __shfl_sync(0xffffffffu, id, g0 + g);
                if (gid32 != TILE_INVALID) {
                    const size_t gid = gid32;
                    if (nsh) {
                        const float* src = in.shs + gid * (size_t)nsh;
                        if (lane < nsh) a0[g] = src[lane];
                        if (32 + lane < nsh) a1[g] = src[32 + lane];
                    }
                    if (MODE == MODE_FOV) {
                        // second half of the slot: [32,45) SH rest, [48,60) the 4 dc triplets, [60,64) the 4 opacities
                        if (lane >= 16 && lane < 28) a1[g] = in.shs_dcs[gid * 12 + (lane - 16)];
                        else if (lane >= 28) a1[g] = in.opacities[gid * 4 + (lane - 28)];
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < 8; g++) { buf[g0 + g][lane] = a0[g]; buf[g0 + g][32 + lane] = a1[g]; }
        }
        __syncwarp();
        if (valid) {
            const float* b = buf[lane];
            if (MODE == MODE_FOV) {
                float4* rec = ws.rec + (size_t)REC_FOV * id;
                float3 rs = make_float3(0.f, 0.f, 0.f);
                if (nsh) rs = sh_accumulate(b, 0, deg, dx, dy, dz, rs);
                rs.x += 0.5f; rs.y += 0.5f; rs.z += 0.5f;
                const int l0 = (int)(lv & 0xff), l1 = (int)((lv >> 8) & 0xff);
                // levels outside [l0,l1] are never composited (the reference leaves them uninitialised, Q4); they are
                // zeroed so that the blending tiles' unconditional level-L2 load stays finite.
#pragma unroll
                for (int l = 0; l < FOV_LEVELS; l++) {
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (l >= l0 && l <= l1) {
                        o.x = b[60 + l];
                        o.y = fmaxf(SH_C0 * b[48 + 3 * l + 0] + rs.x, 0.0f);
                        o.z = fmaxf(SH_C0 * b[48 + 3 * l + 1] + rs.y, 0.0f);
                        o.w = fmaxf(SH_C0 * b[48 + 3 * l + 2] + rs.z, 0.0f);
                    }
                    rec[2 + l] = o;
                }
            } else {
                float4* rec = ws.rec + (size_t)rec_size(MODE) * id;
                float3 c;
                bool cl0 = false, cl1 = false, cl2 = false;
                if (in.colors_precomp != nullptr) {
                    c = make_float3(in.colors_precomp[3 * (size_t)id], in.colors_precomp[3 * (size_t)id + 1], in.colors_precomp[3 * (size_t)id + 2]);
                } else {
                    c = sh_accumulate(b, 1, deg, dx, dy, dz, make_float3(SH_C0 * b[0], SH_C0 * b[1], SH_C0 * b[2]));
                    c.x += 0.5f; c.y += 0.5f; c.z += 0.5f;
                    cl0 = c.x < 0; cl1 = c.y < 0; cl2 = c.z < 0;
                    c.x = fmaxf(c.x, 0.0f); c.y = fmaxf(c.y, 0.0f); c.z = fmaxf(c.z, 0.0f);
                }
                if (shared_model(MODE)) rec[2] = make_float4(in.opacities[id], c.x, c.y, c.z);   // REC_SMFR
                else rec[2] = make_float4(c.x, c.y, c.z, 0.f);
                if (MODE == MODE_SUM) reinterpret_cast<uchar4*>(ws.clamped)[id] = make_uchar4(cl0, cl1, cl2, 0);
            }
        }
        __syncwarp();
    }
    (void)lt_mask;
}

// ------------------------------------------------------------------------------------------------------------------
// k_color_tma: same job as k_color, but the gather of each Gaussian's SH / dc / opacity blocks is done by the TMA engine:
// every lane issues 1-D bulk copies (cp.async.bulk, global -> shared, completion on a per-warp mbarrier) for "its"
// Gaussian, so 32 Gaussians x ~250 B are in flight per warp without holding a single register, and the warp wakes up
// once when the last byte has landed.  (The register-staged variant above keeps 16 x 128 B in flight per warp and is
// latency-bound on these random 180-byte blocks: 0.65 ms for 1.9 M Gaussians vs ~0.1 ms of HBM time.)
// The SH block of Gaussian g starts at byte 180*g — only 4-byte aligned — so the copy fetches the enclosing 16-byte
// aligned window and the consumer reads at the residual offset.  Requires 16-byte aligned base pointers (checked on the
// host; otherwise k_color is used).
// ------------------------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------------------------
// Scatter: staged instance -> its tile's segment of the key array (cursor allocation inside the segment).  12 B in, one
// L2 fetch-and-add and one scattered 8-byte store per instance; the next batch's loads are in flight while this batch is
// placed.  Bound by the scattered stores (61 G/s measured with or without the atomics).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void scatter_role(const Workspace& ws, const uint32_t t, const uint32_t nthreads) {
    const uint32_t n = min(ws.hdr->stage_cursor, ws.stage_cap);
    const uint32_t cap = ws.hdr->cap;
    constexpr int U = SCAT_U;
    uint32_t tl[U], ntl[U];
    unsigned long long k[U], nk[U];
    auto load = [&](uint32_t i0, uint32_t* T, unsigned long long* K) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t i = i0 + u * nthreads;
            const bool in = i < n;
            T[u] = in ? __ldcs(&ws.stage_tile[i]) : TILE_INVALID;
            K[u] = in ? __ldcs(reinterpret_cast<const unsigned long long*>(&ws.stage_key[i])) : 0ull;
        }
    };
    if (t < n) load(t, tl, k);
    for (uint32_t i0 = t; i0 < n; i0 += nthreads * U) {
        const uint32_t nx = i0 + nthreads * U;
        if (nx < n) load(nx, ntl, nk);
        // cursors start at their tile's offset (tile scan), so the atomic returns the slot: per instance one divergent
        // ATOMG and one divergent 8-byte store; the kernel is bound by those LSU wavefronts, not by HBM
        uint32_t slot[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (tl[u] != TILE_INVALID) slot[u] = atomicAdd(&ws.tile_cursor[(size_t)tl[u] * CSTRIDE], 1u);
#pragma unroll
        for (int u = 0; u < U; u++)
            if (tl[u] != TILE_INVALID && slot[u] < cap) ws.keysA[slot[u]] = k[u];
#pragma unroll
        for (int u = 0; u < U; u++) { tl[u] = ntl[u]; k[u] = nk[u]; }
    }
}

// Measured and dropped (profiles/README.md, r2): launching this kernel as the programmatic dependent of k_tile_scan and the
// colour kernel as ITS dependent, so that scatter (L2 atomics, LSU wavefronts) and colour gathers (HBM rows) run side by
// side — 1.041 ms per frame against 1.021 ms with the colour stage first and the scatter alone after it.
__global__ void __launch_bounds__(256) k_scatter(Workspace ws) {
    scatter_role(ws, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

constexpr int TSLOT_FOV = 72;   // floats per slot: [0,56) SH-rest window, [56,68) 4 dc triplets, [68,72) 4 opacities
constexpr int TSLOT_PS1 = 56;   // floats per slot: [0,56) SH window (192 B when aligned)

// Colour role of one warp: `slots` = this warp's [32][SLOT] floats of shared memory, `bar` its mbarrier (count 32),
// gw / nw = index of the warp among / number of all colour warps of the grid.
template <int MODE>
__device__ __forceinline__ void color_tma_role(const Workspace& ws, const FrameInputs& in, const size_t shs_floats,
                                               float* slots, uint64_t* bar, const uint32_t gw, const uint32_t nw,
                                               const float* campos_s, const int deg, const int M) {
    constexpr int SLOT = (MODE == MODE_FOV) ? TSLOT_FOV : TSLOT_PS1;
    const int lane = threadIdx.x & 31;
    const int nsh = (in.shs != nullptr) ? 3 * M : 0;
    const uint32_t nslots = min(ws.hdr->vis_cursor, ws.vis_cap);
    uint32_t parity = 0;
    const uintptr_t shs_beg = (uintptr_t)in.shs, shs_end = shs_beg + shs_floats * 4;
    for (uint32_t s0 = gw * 32; s0 < nslots; s0 += nw * 32) {
        const uint32_t slot = s0 + lane;
        uint32_t id = TILE_INVALID, lv = 0;
        if (slot < nslots) { id = ws.vis_list[slot]; lv = ws.vis_lv[slot]; }
        const bool valid = id != TILE_INVALID;
        if (__ballot_sync(0xffffffffu, valid) == 0) continue;
        float* b = slots + lane * SLOT;
        int sh_off = 0;          // float offset of this Gaussian's SH block inside its window
        bool sh_direct = false;  // window would leave the tensor: plain loads instead
        uint32_t tx = 0;
        // packed model rows (FOV): ONE aligned 256-byte bulk copy per Gaussian = 8 full sectors in 2 DRAM lines, instead of
        // a 4-byte-aligned 180-byte row + 48 + 16 + 12 bytes from four tensors (about 11.5 sectors in 6-7 lines)
        const bool packed = MODE == MODE_FOV && in.packed_rows != nullptr;
        if (packed) {
            mbar_arrive_expect_tx(bar, valid ? 256u : 0u);
            if (valid) bulk_g2s(b, in.packed_rows + (size_t)id * 64, 256, bar);
            mbar_wait(bar, parity);
            parity ^= 1u;
            if (valid) {
                float dx = b[61] - campos_s[0], dy = b[62] - campos_s[1], dz = b[63] - campos_s[2];
                const float len = sqrtf(dx * dx + dy * dy + dz * dz);
                dx = dx / len; dy = dy / len; dz = dz / len;
                float4* rec = ws.rec + (size_t)REC_FOV * id;
                float3 rs = make_float3(0.f, 0.f, 0.f);
                if (nsh) rs = sh_accumulate(b, 0, deg, dx, dy, dz, rs);
                rs.x += 0.5f; rs.y += 0.5f; rs.z += 0.5f;
#pragma unroll
                for (int l = 0; l < FOV_LEVELS; l++) {
                    float4 o;
                    o.x = b[57 + l];
                    o.y = fmaxf(SH_C0 * b[45 + 3 * l + 0] + rs.x, 0.0f);
                    o.z = fmaxf(SH_C0 * b[45 + 3 * l + 1] + rs.y, 0.0f);
                    o.w = fmaxf(SH_C0 * b[45 + 3 * l + 2] + rs.z, 0.0f);
                    rec[2 + l] = o;
                }
            }
            __syncwarp();
            continue;
        }
        if (valid) {
            if (nsh) {
                const uintptr_t beg = shs_beg + (size_t)id * (size_t)nsh * 4, end = beg + (size_t)nsh * 4;
                const uintptr_t wbeg = beg & ~(uintptr_t)15, wend = (end + 15) & ~(uintptr_t)15;
                sh_off = (int)((beg - wbeg) >> 2);
                if (wbeg < shs_beg || wend > shs_end || (wend - wbeg) > 56 * 4) sh_direct = true;
                else tx += (uint32_t)(wend - wbeg);
            }
            if (MODE == MODE_FOV) tx += 48 + 16;
        }
        mbar_arrive_expect_tx(bar, tx);      // every lane arrives (idle lanes with 0 bytes): 32 arrivals complete a phase
        if (valid) {
            if (nsh && !sh_direct) {
                const uintptr_t beg = shs_beg + (size_t)id * (size_t)nsh * 4, end = beg + (size_t)nsh * 4;
                const uintptr_t wbeg = beg & ~(uintptr_t)15, wend = (end + 15) & ~(uintptr_t)15;
                bulk_g2s(b, (const void*)wbeg, (uint32_t)(wend - wbeg), bar);
            }
            if (MODE == MODE_FOV) {
                bulk_g2s(b + 56, in.shs_dcs + (size_t)id * 12, 48, bar);
                bulk_g2s(b + 68, in.opacities + (size_t)id * 4, 16, bar);
            }
        }
        // overlap: view direction while the copies fly
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (valid) {
            const float mx = in.means3D[3 * (size_t)id], my = in.means3D[3 * (size_t)id + 1], mz = in.means3D[3 * (size_t)id + 2];
            dx = mx - campos_s[0]; dy = my - campos_s[1]; dz = mz - campos_s[2];
            const float len = sqrtf(dx * dx + dy * dy + dz * dz);
            dx = dx / len; dy = dy / len; dz = dz / len;
            if (sh_direct) {
                const float* src = in.shs + (size_t)id * (size_t)nsh;
                for (int k = 0; k < nsh; k++) b[k] = src[k];
                sh_off = 0;
            }
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        if (valid) {
            const float* sh = b + sh_off;
            if (MODE == MODE_FOV) {
                float4* rec = ws.rec + (size_t)REC_FOV * id;
                float3 rs = make_float3(0.f, 0.f, 0.f);
                if (nsh) rs = sh_accumulate(sh, 0, deg, dx, dy, dz, rs);
                rs.x += 0.5f; rs.y += 0.5f; rs.z += 0.5f;
                const int l0 = (int)(lv & 0xff), l1 = (int)((lv >> 8) & 0xff);
                // levels outside [l0,l1] are never composited (the reference leaves them uninitialised, Q4); zeroed so
                // that the blending tiles' unconditional level-L2 load stays finite.
#pragma unroll
                for (int l = 0; l < FOV_LEVELS; l++) {
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (l >= l0 && l <= l1) {
                        o.x = b[68 + l];
                        o.y = fmaxf(SH_C0 * b[56 + 3 * l + 0] + rs.x, 0.0f);
                        o.z = fmaxf(SH_C0 * b[56 + 3 * l + 1] + rs.y, 0.0f);
                        o.w = fmaxf(SH_C0 * b[56 + 3 * l + 2] + rs.z, 0.0f);
                    }
                    rec[2 + l] = o;
                }
            } else {
                float4* rec = ws.rec + (size_t)rec_size(MODE) * id;
                float3 c;
                bool cl0 = false, cl1 = false, cl2 = false;
                if (in.colors_precomp != nullptr) {
                    c = make_float3(in.colors_precomp[3 * (size_t)id], in.colors_precomp[3 * (size_t)id + 1], in.colors_precomp[3 * (size_t)id + 2]);
                } else {
                    c = sh_accumulate(sh, 1, deg, dx, dy, dz, make_float3(SH_C0 * sh[0], SH_C0 * sh[1], SH_C0 * sh[2]));
                    c.x += 0.5f; c.y += 0.5f; c.z += 0.5f;
                    cl0 = c.x < 0; cl1 = c.y < 0; cl2 = c.z < 0;
                    c.x = fmaxf(c.x, 0.0f); c.y = fmaxf(c.y, 0.0f); c.z = fmaxf(c.z, 0.0f);
                }
                if (shared_model(MODE)) rec[2] = make_float4(in.opacities[id], c.x, c.y, c.z);   // REC_SMFR
                else rec[2] = make_float4(c.x, c.y, c.z, 0.f);
                if (MODE == MODE_SUM) reinterpret_cast<uchar4*>(ws.clamped)[id] = make_uchar4(cl0, cl1, cl2, 0);
            }
        }
        __syncwarp();   // all generic-proxy reads of the slots are done before the next round's bulk copies overwrite them
    }
    pdl_wait();   // second of the (k_tile_scan, colour) pair, see k_color_tma
}

// Measured and dropped: running the colour gathers and the scatter as two warp roles of ONE launch.  Alone they take
// 0.159 ms and 0.155 ms, fused 0.328 ms — both are bound by the same thing (random 32-byte sectors through L2/HBM), so
// there is nothing to overlap.  They stay two back-to-back launches, each at its own best occupancy.
template <int MODE>
__global__ void __launch_bounds__(CW * 32) k_color_tma(Workspace ws, FrameInputs in, size_t shs_floats) {
    constexpr int SLOT = (MODE == MODE_FOV) ? TSLOT_FOV : TSLOT_PS1;
    __shared__ __align__(16) float tbuf[CW][32][SLOT];
    __shared__ __align__(8) uint64_t bars[CW];
    __shared__ float campos_s[3];
    __shared__ int deg_s, M_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 3) campos_s[tid] = ws.hdr->cam.campos[tid];
    if (tid == 0) { deg_s = ws.hdr->cam.sh_degree; M_s = ws.hdr->cam.M; }
    if (lane == 0) {
        mbar_init(&bars[warp], 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    color_tma_role<MODE>(ws, in, shs_floats, &tbuf[warp][0][0], &bars[warp], blockIdx.x * CW + warp, gridDim.x * CW, campos_s,
                         deg_s, M_s);
    pdl_wait();   // second of the (k_tile_scan, colour) pair: the stream goes on only when the scan is complete as well
}

// ------------------------------------------------------------------------------------------------------------------
// k_tile_scan: one CTA of 1024 threads turns the per-tile histogram k_pre left into
//   tile_offset (exclusive scan = the reference's `ranges`; no identifyTileRanges), tile_cursor (the scatter allocates
//   slots straight from it), num_rendered / overflow / max_tile_instances, cum_class, and the heavy-first tile orders
//   (tile_order2: blending tiles first, each kind by descending power-of-two size class; tile_order: by class only, written
//   when the full-sort path will run).
// Nothing in the colour stage needs any of this, so the kernel is the FIRST of a programmatic-dependent-launch pair with the
// colour kernel: it triggers the dependent launch on entry, the colour CTAs fill the other 147 SMs at once and the scan
// hides behind the colour gathers (as the last CTA of k_pre it held the whole GPU idle for ~65 us per frame: 256 threads,
// 32 dependent round trips to L2-resident counters that sit 256 bytes apart, three match/atomic sequences per tile).
// Organised around round trips: a thread owns 8 consecutive tiles and requests its 8 counters before it uses one (1080p:
// 8160 tiles = one sweep).  The counting sorts use no atomics: per sweep every warp ranks its items per class with
// `match.any` into its own row of a [warp][class] table, one block scan over the table (class-major, warp-minor) turns the
// rows into bases.  Counters are read with ld.cg: they were only ever touched by L2 atomics.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SCAN_NT = 512, SCAN_NW = SCAN_NT / 32, SCAN_IT = 8;   // 512 threads x ~54 registers: leaves room for other kernels' CTAs on its SM
constexpr int SCAN_BINS = 68;      // (kind, class): class = 32 - clz(n) in [0, 32] (0: empty), bin = class + 34 * kind
struct ScanSmem {
    uint32_t warp_sums[SCAN_NW];
    uint32_t carry, maxv;
    uint32_t table[SCAN_BINS][SCAN_NW];     // [bin][warp]: counts, then bases of tile_order2
    uint32_t table1[34][SCAN_NW];           // the same for tile_order (both kinds merged)
    uint32_t bin_total[SCAN_BINS];
};

// exclusive block scan of one value per thread (SCAN_NT threads); returns the exclusive prefix, *total = the block's sum
__device__ __forceinline__ uint32_t scan_block_excl(uint32_t x, uint32_t* warp_sums, uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    __syncthreads();                         // warp_sums may still be read by the previous call
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    uint32_t w = lane < SCAN_NW ? warp_sums[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
    *total = __shfl_sync(0xffffffffu, wi, 31);
    const uint32_t wbase = __shfl_sync(0xffffffffu, wi - w, wid);
    return wbase + incl - x;
}

__global__ void __launch_bounds__(SCAN_NT, 1) k_tile_scan(FrameHeader* hdr, const uint32_t* __restrict__ tile_count,
                                                          uint32_t* __restrict__ tile_offset, uint32_t* __restrict__ tile_cursor,
                                                          uint32_t* __restrict__ tile_order, uint32_t* __restrict__ tile_order2,
                                                          const uint8_t* __restrict__ tile_blend, uint32_t stage_cap,
                                                          int want_order1, uint32_t* stats_host) {
    static_assert(SCAN_NW <= 32, "scan_block_excl scans the warp sums in one warp");
    __shared__ ScanSmem s;
    pdl_trigger();                           // the colour kernel may start now: it shares nothing with this one
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int T = hdr->tiles;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int i = tid; i < SCAN_BINS * SCAN_NW; i += SCAN_NT) (&s.table[0][0])[i] = 0;
    for (int i = tid; i < 34 * SCAN_NW; i += SCAN_NT) (&s.table1[0][0])[i] = 0;
    if (tid == 0) { s.carry = 0; s.maxv = 0; }
    __syncthreads();
    auto load = [&](int i0, uint32_t* v, int* bin) {
#pragma unroll
        for (int u = 0; u < SCAN_IT; u++) v[u] = (i0 + u < T) ? __ldcg(&tile_count[(size_t)(i0 + u) * CSTRIDE]) : 0u;
        unsigned long long kinds = 0ull;     // 8 tile_blend bytes (sub-buffers are 256-byte aligned, i0 is a multiple of 8)
        if (tile_blend != nullptr) {
            if (i0 + SCAN_IT <= T) kinds = *reinterpret_cast<const unsigned long long*>(tile_blend + i0);
            else
                for (int u = 0; u < SCAN_IT; u++)
                    if (i0 + u < T) kinds |= (unsigned long long)tile_blend[i0 + u] << (8 * u);
        }
#pragma unroll
        for (int u = 0; u < SCAN_IT; u++) {
            const int cls = v[u] ? 32 - __clz(v[u]) : 0;
            bin[u] = (i0 + u < T) ? cls + (((kinds >> (8 * u)) & 0xffull) ? 34 : 0) : -1;
        }
    };
    // ---- pass A: offsets, cursors, per-(warp, bin) counts ----
    uint32_t local_max = 0;
    for (int base = 0; base < T; base += SCAN_NT * SCAN_IT) {
        const int i0 = base + tid * SCAN_IT;
        uint32_t v[SCAN_IT];
        int bin[SCAN_IT];
        load(i0, v, bin);
        uint32_t sum = 0;
#pragma unroll
        for (int u = 0; u < SCAN_IT; u++) { sum += v[u]; local_max = max(local_max, v[u]); }
        uint32_t total;
        uint32_t off = scan_block_excl(sum, s.warp_sums, &total) + s.carry;
#pragma unroll
        for (int u = 0; u < SCAN_IT; u++) {
            if (i0 + u < T) { tile_offset[i0 + u] = off; tile_cursor[(size_t)(i0 + u) * CSTRIDE] = off; }
            off += v[u];
        }
#pragma unroll
        for (int u = 0; u < SCAN_IT; u++) {
            const int b = bin[u] >= 0 ? bin[u] : SCAN_BINS + lane;      // idle lanes: singleton groups
            const unsigned peers = __match_any_sync(0xffffffffu, b);
            if (bin[u] >= 0 && (peers & lt_mask) == 0) s.table[b][wid] += (uint32_t)__popc(peers);   // the group's first lane owns the row entry
            if (want_order1) {
                const int b1 = bin[u] >= 0 ? bin[u] % 34 : 34 + lane;
                const unsigned peers1 = __match_any_sync(0xffffffffu, b1);
                if (bin[u] >= 0 && (peers1 & lt_mask) == 0) s.table1[b1][wid] += (uint32_t)__popc(peers1);
            }
            __syncwarp();
        }
        __syncthreads();
        if (tid == 0) s.carry += total;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if (lane == 0 && local_max) atomicMax(&s.maxv, local_max);
    __syncthreads();
    // ---- bases: exclusive scan over the tables in output order (bins descending — blending kind first —, warps ascending) ----
    {
        // tile_order2: flat entry f = r * 32 + w, r = 0..67 <-> bin 67 - r  (bins 67..34: blending tiles by descending class)
        constexpr int E2 = (SCAN_BINS * SCAN_NW + SCAN_NT - 1) / SCAN_NT;
        uint32_t c[E2], excl_total;
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < E2; k++) {
            const int f = tid * E2 + k;
            c[k] = (f < SCAN_BINS * SCAN_NW) ? s.table[SCAN_BINS - 1 - f / SCAN_NW][f % SCAN_NW] : 0u;
            mine += c[k];
        }
        uint32_t run = scan_block_excl(mine, s.warp_sums, &excl_total);
#pragma unroll
        for (int k = 0; k < E2; k++) {
            const int f = tid * E2 + k;
            if (f < SCAN_BINS * SCAN_NW) s.table[SCAN_BINS - 1 - f / SCAN_NW][f % SCAN_NW] = run;
            run += c[k];
        }
    }
    if (want_order1) {
        constexpr int E1 = (34 * SCAN_NW + SCAN_NT - 1) / SCAN_NT;
        uint32_t c[E1], excl_total;
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < E1; k++) {
            const int f = tid * E1 + k;
            c[k] = (f < 34 * SCAN_NW) ? s.table1[33 - f / SCAN_NW][f % SCAN_NW] : 0u;
            mine += c[k];
        }
        uint32_t run = scan_block_excl(mine, s.warp_sums, &excl_total);
#pragma unroll
        for (int k = 0; k < E1; k++) {
            const int f = tid * E1 + k;
            if (f < 34 * SCAN_NW) s.table1[33 - f / SCAN_NW][f % SCAN_NW] = run;
            run += c[k];
        }
    }
    __syncthreads();
    if (tid < 34) {
        // tiles of class >= b (both kinds) = position at which class b - 1 starts in tile_order... computed from the tile_order2
        // table: a bin's size = next bin's base - its base
        const int b = tid;
        auto bin_size = [&](int bn) {   // tile_order2 bases are in order 67, 66, ..., 0
            const uint32_t beg = s.table[bn][0];
            const uint32_t end = bn > 0 ? s.table[bn - 1][0] : (uint32_t)T;
            return end - beg;
        };
        s.bin_total[b] = bin_size(b);
        s.bin_total[b + 34] = bin_size(b + 34);
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t total = s.carry;
        tile_offset[T] = total;
        hdr->stats.num_rendered = total;
        hdr->stats.overflow = (total > hdr->cap || hdr->stage_cursor > stage_cap) ? 1u : 0u;
        hdr->stats.max_tile_instances = s.maxv;
        uint32_t run = 0;
#pragma unroll 1
        for (int b = 32; b >= 0; b--) { run += s.bin_total[b] + s.bin_total[b + 34]; hdr->cum_class[b] = run; }
        hdr->cum_class[33] = 0;
        uint32_t k1 = 0;
#pragma unroll 1
        for (int b = 0; b < 34; b++) k1 += s.bin_total[b + 34];
        hdr->lazy_count[1] = k1;
        hdr->lazy_count[0] = (uint32_t)T - k1;
    }
    if (stats_host != nullptr) {
        // the frame statistics the host waits for (instance count, overflow flag, visible count, ...) go straight into its pinned,
        // device-mapped buffer: no 64-byte device->host copy sits in the stream between two kernels of the frame
        __syncthreads();
        if (tid < (int)(sizeof(fovgs_frame_stats) / 4)) stats_host[tid] = reinterpret_cast<const volatile uint32_t*>(&hdr->stats)[tid];
        __threadfence_system();
    }
    // ---- pass B: positions = base of (bin, warp) + rank inside the warp; the row entry is the warp's running cursor ----
    for (int base = 0; base < T; base += SCAN_NT * SCAN_IT) {
        const int i0 = base + tid * SCAN_IT;
        uint32_t v[SCAN_IT];
        int bin[SCAN_IT];
        load(i0, v, bin);
#pragma unroll
        for (int u = 0; u < SCAN_IT; u++) {
            const int b = bin[u] >= 0 ? bin[u] : SCAN_BINS + lane;
            {
                const unsigned peers = __match_any_sync(0xffffffffu, b);
                const int leader = __ffs(peers) - 1;
                uint32_t old = 0;
                if (bin[u] >= 0 && lane == leader) { old = s.table[b][wid]; s.table[b][wid] = old + (uint32_t)__popc(peers); }
                old = __shfl_sync(0xffffffffu, old, leader);
                if (bin[u] >= 0) tile_order2[old + __popc(peers & lt_mask)] = (uint32_t)(i0 + u);
            }
            if (want_order1) {
                const int b1 = bin[u] >= 0 ? bin[u] % 34 : 34 + lane;
                const unsigned peers = __match_any_sync(0xffffffffu, b1);
                const int leader = __ffs(peers) - 1;
                uint32_t old = 0;
                if (bin[u] >= 0 && lane == leader) { old = s.table1[b1][wid]; s.table1[b1][wid] = old + (uint32_t)__popc(peers); }
                old = __shfl_sync(0xffffffffu, old, leader);
                if (bin[u] >= 0) tile_order[old + __popc(peers & lt_mask)] = (uint32_t)(i0 + u);
            }
            __syncwarp();
        }
    }
}

cudaError_t launch_tile_scan(const Workspace& ws, bool want_order1, uint32_t* stats_host, cudaStream_t st) {
    // an ordinary launch (it needs all of k_pre); the colour kernel that follows is its programmatic dependent
    return launch_chained(false, k_tile_scan, dim3(1), dim3(SCAN_NT), 0, st, ws.hdr, (const uint32_t*)ws.tile_count, ws.tile_offset,
                          ws.tile_cursor, ws.tile_order, ws.tile_order2, (const uint8_t*)ws.tile_blend, ws.stage_cap,
                          want_order1 ? 1 : 0, stats_host);
}

cudaError_t launch_pre(const Workspace& ws, const FrameInputs& in, Mode mode, int num_sms, cudaStream_t st) {
    const int need = (in.P + PB - 1) / PB;
    const int grid = max(1, min(need, min(num_sms * PRE_CTAS, (int)STAGE_MAX_BLOCKS)));
    const size_t smem = sizeof(PreSmem);
    static PerDeviceOnce once;
    bool* configured = once.slot();
    if (!*configured) {
        cudaError_t e;
#define PRE_SET(K)                                                                              \
        e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        if (e != cudaSuccess) return e;
        PRE_SET(k_pre<MODE_OBB>) PRE_SET(k_pre<MODE_SUM>) PRE_SET(k_pre<MODE_FOV>) PRE_SET(k_pre<MODE_SMFR>) PRE_SET(k_pre<MODE_MMFR>)
#undef PRE_SET
        *configured = true;
    }
    switch (mode) {
        case MODE_OBB: k_pre<MODE_OBB><<<grid, PB, smem, st>>>(ws, in); break;
        case MODE_SUM: k_pre<MODE_SUM><<<grid, PB, smem, st>>>(ws, in); break;
        case MODE_SMFR: k_pre<MODE_SMFR><<<grid, PB, smem, st>>>(ws, in); break;
        case MODE_MMFR: k_pre<MODE_MMFR><<<grid, PB, smem, st>>>(ws, in); break;
        default: k_pre<MODE_FOV><<<grid, PB, smem, st>>>(ws, in); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_color(const Workspace& ws, const FrameInputs& in, Mode mode, int num_sms, cudaStream_t st) {
    // 6 CTAs per SM on all SMs but one: k_tile_scan's 1024-thread CTA runs beside this kernel (PDL pair) and fills most of
    // an SM; with a grid that needs every slot, the CTAs that find theirs taken would run as a second wave after the scan
    const int grid = max(1, num_sms - 1) * 6;
    // TMA bulk gathers need 16-byte aligned bases (the SH window logic handles the 4-byte aligned per-Gaussian offsets)
    const bool packed = mode == MODE_FOV && in.packed_rows != nullptr && (((uintptr_t)in.packed_rows) & 15) == 0;
    const bool aligned = packed || (((uintptr_t)in.shs | (uintptr_t)in.shs_dcs | (uintptr_t)in.opacities) & 15) == 0;
    const int nsh = in.shs ? 3 * in.M : 0;
    if (aligned && nsh <= (packed ? 45 : 48) && !g_no_tma) {
        const size_t shs_floats = (size_t)in.P * (size_t)nsh;
        switch (mode) {
            case MODE_OBB: launch_chained(true, k_color_tma<MODE_OBB>, dim3(grid), dim3(CW * 32), 0, st, ws, in, shs_floats); break;
            case MODE_SUM: launch_chained(true, k_color_tma<MODE_SUM>, dim3(grid), dim3(CW * 32), 0, st, ws, in, shs_floats); break;
            case MODE_SMFR: launch_chained(true, k_color_tma<MODE_SMFR>, dim3(grid), dim3(CW * 32), 0, st, ws, in, shs_floats); break;
            case MODE_MMFR: launch_chained(true, k_color_tma<MODE_MMFR>, dim3(grid), dim3(CW * 32), 0, st, ws, in, shs_floats); break;
            default: launch_chained(true, k_color_tma<MODE_FOV>, dim3(grid), dim3(CW * 32), 0, st, ws, in, shs_floats); break;
        }
        return cudaGetLastError();
    }
    switch (mode) {
        case MODE_OBB: launch_chained(true, k_color<MODE_OBB>, dim3(grid), dim3(CW * 32), 0, st, ws, in); break;
        case MODE_SUM: launch_chained(true, k_color<MODE_SUM>, dim3(grid), dim3(CW * 32), 0, st, ws, in); break;
        case MODE_SMFR: launch_chained(true, k_color<MODE_SMFR>, dim3(grid), dim3(CW * 32), 0, st, ws, in); break;
        case MODE_MMFR: launch_chained(true, k_color<MODE_MMFR>, dim3(grid), dim3(CW * 32), 0, st, ws, in); break;
        default: launch_chained(true, k_color<MODE_FOV>, dim3(grid), dim3(CW * 32), 0, st, ws, in); break;
    }
    return cudaGetLastError();
}

// one-time re-layout of the static model tensors the colour stage gathers (see FrameInputs::packed_rows)
__global__ void k_pack_color_rows(int P, int M_rest, const float* __restrict__ means3D, const float* __restrict__ shs_rest,
                                  const float* __restrict__ shs_dcs, const float* __restrict__ opacities, float* __restrict__ rows) {
    const size_t n = (size_t)P * 64;
    const int nsh = 3 * M_rest;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t g = i >> 6;
        const int k = (int)(i & 63);
        float v = 0.0f;
        if (k < 45) v = (k < nsh) ? shs_rest[g * (size_t)nsh + k] : 0.0f;
        else if (k < 57) v = shs_dcs[g * 12 + (k - 45)];
        else if (k < 61) v = opacities[g * 4 + (k - 57)];
        else v = means3D[g * 3 + (k - 61)];
        rows[i] = v;
    }
}

cudaError_t launch_pack_color_rows(int P, int M_rest, const float* means3D, const float* shs_rest, const float* shs_dcs,
                                   const float* opacities, float* rows, cudaStream_t st) {
    k_pack_color_rows<<<148 * 16, 256, 0, st>>>(P, M_rest, means3D, shs_rest, shs_dcs, opacities, rows);
    return cudaGetLastError();
}

cudaError_t launch_scatter(const Workspace& ws, int num_sms, cudaStream_t st) {
    k_scatter<<<num_sms * SCAT_CTAS, 256, 0, st>>>(ws);
    return cudaGetLastError();
}

}  // namespace fovgs
