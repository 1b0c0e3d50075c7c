// fovgs_kernels.cu — sm_100a kernels of the forward path:
//   setup / tile levels  ->  preprocess+filter (count)  ->  tile scan  ->  emit+colour  ->  per-tile sort  ->  blend
//
// Design: binning is a two-level sort.  Level 1 is a counting sort by tile (histogram in the
// preprocess pass, one-block scan, cursor scatter in the emit pass) which yields the reference's `ranges` for
// free; level 2 sorts each tile's segment by depth bits with a block-local LSD radix sort (ties broken by
// Gaussian id), which reproduces exactly the order of the reference's stable 45-bit global radix sort
// (FOV/cuda_rasterizer/rasterizer_impl.cu:843-854).  No host synchronisation anywhere.
#include "fovgs_internal.cuh"

namespace fovgs {

// ------------------------------------------------------------------------------------------------------------------
// SH constants (reference auxiliary.h:35-52)
// ------------------------------------------------------------------------------------------------------------------
__device__ constexpr float SH_C0 = 0.28209479177387814f;
__device__ constexpr float SH_C1 = 0.4886025119029199f;
__device__ constexpr float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                       -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                       0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                       -0.5900435899266435f};

// ------------------------------------------------------------------------------------------------------------------
// FOV tile tables.  Same expression order and literal types as the reference kernels so that nvcc/ptxas emit
// the same arithmetic (checked by tools/compare_sass.py against oracle/_ref): compute_tile_levels_cuda
// (FOV/rasterizer_impl.cu:120-177), ps2level (auxiliary.h:55-66), compute_tile_level_infos_cuda (:182-260).
// ------------------------------------------------------------------------------------------------------------------
constexpr float kRealImageWidth = 2.0f;
constexpr float kRealViewingDistance = 1.0f;
constexpr float kSqrtMaxPs = 3.4641016151377544f;
constexpr float kStartBlend = 0.5f;
constexpr float kBlendWidth = 0.5f;
#ifndef FOVGS_PI
#define FOVGS_PI 3.14159265358979323846
#endif

__forceinline__ __device__ float tl_distance(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }
__forceinline__ __device__ void tl_norm(float3& v) {
    float d = tl_distance(v.x, v.y, v.z);
    v.x /= d;
    v.y /= d;
    v.z /= d;
}
__forceinline__ __device__ float3 tl_ncd2dir(const float2 ncd, const float real_width, const float real_height) {
    float3 v;
    v.x = (ncd.x - 0.5f) * real_width;
    v.y = (ncd.y - 0.5f) * real_height;
    v.z = kRealViewingDistance;
    tl_norm(v);
    return v;
}
__forceinline__ __device__ float tl_dot(const float3 a, const float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__forceinline__ __device__ void tl_ps2level(const float pooling_size, float& level) {
    const float step = (kSqrtMaxPs - 1.) / float(FOV_LEVELS - 1);
    if (pooling_size <= 1) {
        level = 0;
    } else {
        level = (sqrtf(pooling_size) - 1) / step;
    }
}

__global__ void k_tile_levels(int T, float* __restrict__ tile_levels, const float* __restrict__ gaze_ptr, const int W,
                              const int H, const int tile_width_num, const float alpha) {
    auto idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)T) return;
    const float2 gaze = make_float2(gaze_ptr[0], gaze_ptr[1]);
    int tile_y = idx / tile_width_num;
    int tile_x = idx % tile_width_num;
    float p_x = tile_x * TILE + TILE / 2;
    float p_y = tile_y * TILE + TILE / 2;
    float2 p = make_float2(p_x, p_y);
    float real_image_height = float(H) / float(W) * kRealImageWidth;

    float2 tile_ncd;
    tile_ncd.x = (p.x / W);
    tile_ncd.y = (p.y / H);
    float3 tile_dir = tl_ncd2dir(tile_ncd, kRealImageWidth, real_image_height);
    float3 gaze_dir = tl_ncd2dir(gaze, kRealImageWidth, real_image_height);
    const float2 center_ncd = make_float2(0.5, 0.5);
    float3 center_dir = tl_ncd2dir(center_ncd, kRealImageWidth, real_image_height);

    float ecc = acosf(tl_dot(gaze_dir, tile_dir));
    float ecc_center = acosf(tl_dot(tile_dir, center_dir));

    float pooling_rad = alpha * ecc * ecc;
    float angle_min = ecc_center - pooling_rad * 0.5;
    float angle_max = ecc_center + pooling_rad * 0.5;

    float distance_to_pixel =
        tl_distance((tile_ncd.x - 0.5) * kRealImageWidth, (tile_ncd.y - 0.5) * real_image_height, kRealViewingDistance);
    float major_axis = (tanf(angle_max) - tanf(angle_min)) * kRealViewingDistance;
    float minor_axis = 2.0f * distance_to_pixel * tanf(pooling_rad * 0.5f);

    float area = FOVGS_PI * major_axis * minor_axis * 0.25f;
    float real2pix_factor = W / kRealImageWidth;
    float pooling_size = sqrtf(area) * real2pix_factor;

    float level;
    tl_ps2level(pooling_size, level);
    if (level > (float(FOV_LEVELS) - 0.1)) {
        level = (float(FOV_LEVELS) - 0.1);
    }
    tile_levels[idx] = level;
}

__global__ void k_tile_infos(int T, const float* __restrict__ tile_levels, const int tile_width_num,
                             const int tile_height_num, float* __restrict__ grad_y, float* __restrict__ grad_x,
                             float* __restrict__ tile_level_min, uint8_t* __restrict__ tile_blendings,
                             FrameHeader* __restrict__ hdr) {
    auto idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool blend = false;
    if (idx < (uint64_t)T) {
        int tile_y = idx / tile_width_num;
        int tile_x = idx % tile_width_num;
        float tile_level_f = tile_levels[idx];
        float right_level = -1, left_level = -1, up_level = -1, down_level = -1;
        if (tile_x + 1 < tile_width_num) right_level = tile_levels[(tile_x + 1) + tile_width_num * tile_y];
        if (tile_x - 1 >= 0) left_level = tile_levels[(tile_x - 1) + tile_width_num * tile_y];
        if (tile_y + 1 < tile_height_num) up_level = tile_levels[tile_x + tile_width_num * (tile_y + 1)];
        if (tile_y - 1 >= 0) down_level = tile_levels[tile_x + tile_width_num * (tile_y - 1)];

        float gx = 0, gy = 0;
        if (right_level != -1 && left_level != -1) gx = (right_level - left_level) / 2.0f;
        else if (right_level != -1) gx = right_level - tile_level_f;
        else if (left_level != -1) gx = tile_level_f - left_level;
        if (up_level != -1 && down_level != -1) gy = (up_level - down_level) / 2.0f;
        else if (up_level != -1) gy = up_level - tile_level_f;
        else if (down_level != -1) gy = tile_level_f - down_level;

        float max_delta = 0.5 * (fabsf(gx) + fabsf(gy));
        float tile_min = tile_level_f - max_delta;
        tile_level_min[idx] = tile_min;
        float tile_min_i = float(int(tile_min));
        blend = ((tile_min - tile_min_i) > kStartBlend && (tile_min_i < (FOV_LEVELS - 1)));
        tile_blendings[idx] = blend ? 1 : 0;
        grad_y[idx] = gy;
        grad_x[idx] = gx;
    }
    const unsigned m = __ballot_sync(0xffffffffu, blend);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&hdr->stats.num_blend_tiles, __popc(m));
}

// ------------------------------------------------------------------------------------------------------------------
// setup: frame header (camera block read from device pointers — no host copy), zero tile counters
// ------------------------------------------------------------------------------------------------------------------
struct SetupArgs {
    const float* view;
    const float* proj;
    const float* campos;
    const float* bg;
    const float* gaze;
    float tanfovx, tanfovy, focal_x, focal_y, scale_modifier, alpha;
    int W, H, gx, gy, sh_degree, M, P, tiles;
    uint32_t cap;
};

__global__ void k_setup(Workspace ws, SetupArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.tiles) {
        ws.tile_count[t] = 0;
        ws.tile_cursor[t] = 0;
    }
    if (blockIdx.x == 0) {
        FrameHeader* h = ws.hdr;
        const int i = threadIdx.x;
        if (i < 16) {
            h->cam.view[i] = a.view[i];
            h->cam.proj[i] = a.proj[i];
            ((uint32_t*)&h->stats)[i] = 0;
        }
        if (i < 3) {
            h->cam.campos[i] = a.campos[i];
            h->bg[i] = a.bg[i];
        }
        if (i < 2) h->gaze[i] = a.gaze ? a.gaze[i] : 0.5f;
        if (i == 0) {
            h->cam.tanfovx = a.tanfovx;
            h->cam.tanfovy = a.tanfovy;
            h->cam.focal_x = a.focal_x;
            h->cam.focal_y = a.focal_y;
            h->cam.scale_modifier = a.scale_modifier;
            h->cam.W = a.W;
            h->cam.H = a.H;
            h->cam.grid_x = a.gx;
            h->cam.grid_y = a.gy;
            h->cam.sh_degree = a.sh_degree;
            h->cam.M = a.M;
            h->alpha = a.alpha;
            h->P = a.P;
            h->tiles = a.tiles;
            h->cap = a.cap;
        }
    }
}

__device__ __forceinline__ void load_cam(CamParams& dst_smem, const FrameHeader* __restrict__ hdr) {
    const int n = (int)(sizeof(CamParams) / 4);
    const uint32_t* src = (const uint32_t*)&hdr->cam;
    uint32_t* dst = (uint32_t*)&dst_smem;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------
// Pass 1: per-Gaussian preprocess + tile filter (count).  Replaces preprocessCUDA + InclusiveSum + filter/OBB_test
// (FOV/forward.cu:104-238, FOV/rasterizer_impl.cu:264-383; SUM/rasterizer_impl.cu:70-146).
// ------------------------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ unsigned filter_rect(const Workspace& ws, const Splat& s, int gx, float hl1, bool emit,
                                                uint64_t key, uint32_t cap, float& lo_level, float& hi_level,
                                                bool& any_blend) {
    const unsigned tnum = (unsigned)(s.y1 - s.y0) * (unsigned)(s.x1 - s.x0);
    unsigned count = 0;
    if (tnum == 1) {
        const uint32_t tile = (uint32_t)s.y0 * gx + s.x0;
        bool pass = true;
        if (MODE == MODE_FOV) {
            const float level = ws.tile_min[tile];
            pass = level < hl1;
            if (pass) {
                lo_level = level;
                hi_level = level;
                any_blend = any_blend || ws.tile_blend[tile];
            }
        }
        if (pass) {
            count = 1;
            if (emit) {
                const uint32_t slot = ws.tile_offset[tile] + atomicAdd(&ws.tile_cursor[tile], 1u);
                if (slot < cap) ws.keysA[slot] = key;
            } else {
                atomicAdd(&ws.tile_count[tile], 1u);
            }
        }
        return count;
    }
    ObbCorners oc;
    obb_corners(s.px, s.py, s.e1x, s.e1y, s.e2x, s.e2y, s.len1, s.len2, oc);
    for (int y = s.y0; y < s.y1; y++) {
        const float tcy = FF((float)y, 16.0f, 8.0f);
        for (int x = s.x0; x < s.x1; x++) {
            const uint32_t tile = (uint32_t)y * gx + x;
            float level = 0.0f;
            if (MODE == MODE_FOV) {
                level = ws.tile_min[tile];
                if (!(level < hl1)) continue;
            }
            const float tcx = FF((float)x, 16.0f, 8.0f);
            if (!obb_hits_tile(oc, s.px, s.py, s.e1x, s.e1y, s.e2x, s.e2y, s.len1, s.len2, tcx, tcy)) continue;
            count++;
            if (MODE == MODE_FOV) {
                lo_level = fminf(lo_level, level);
                hi_level = fmaxf(hi_level, level);
                any_blend = any_blend || ws.tile_blend[tile];
            }
            if (emit) {
                const uint32_t slot = ws.tile_offset[tile] + atomicAdd(&ws.tile_cursor[tile], 1u);
                if (slot < cap) ws.keysA[slot] = key;
            } else {
                atomicAdd(&ws.tile_count[tile], 1u);
            }
        }
    }
    return count;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_preprocess(Workspace ws, FrameInputs in) {
    __shared__ CamParams cam;
    load_cam(cam, ws.hdr);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    bool visible = false;
    if (idx < in.P) {
        const float mx = in.means3D[3 * idx], my = in.means3D[3 * idx + 1], mz = in.means3D[3 * idx + 2];
        Splat s;
        float c3[6];
        bool ok;
        if (in.cov3D_precomp != nullptr) {
            // precomputed Σ3D path (FOV/forward.cu:155-158): only the projection differs
            const float tz = xform_row(cam.view, 2, mx, my, mz);
            ok = tz > 0.2f;
            if (ok) {
                // reuse project_splat with an identity-free shortcut: compute everything but Σ3D
                for (int k = 0; k < 6; k++) c3[k] = in.cov3D_precomp[6 * idx + k];
                const float hx = xform_row(cam.proj, 0, mx, my, mz);
                const float hy = xform_row(cam.proj, 1, mx, my, mz);
                const float hw = xform_row(cam.proj, 3, mx, my, mz);
                const float pw = __frcp_rn(FA(hw, 0.0000001f));
                const float tx = xform_row(cam.view, 0, mx, my, mz);
                const float ty = xform_row(cam.view, 1, mx, my, mz);
                cov2d_from_cov3d(cam, tx, ty, tz, c3, s.cxx, s.cxy, s.cyy);
                const float bb = FM(s.cxy, s.cxy);
                const float det = FF(s.cxx, s.cyy, -bb);
                ok = det != 0.0f;
                if (ok) {
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
                    ok = tnum != 0;
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
                }
            }
        } else {
            const float sx = in.scales[3 * idx], sy = in.scales[3 * idx + 1], sz = in.scales[3 * idx + 2];
            const float4 q = *reinterpret_cast<const float4*>(in.rotations + 4 * idx);
            ok = project_splat(cam, mx, my, mz, sx, sy, sz, q.x, q.y, q.z, q.w, s, c3);
        }
        unsigned count = 0;
        float hl = 0.0f;
        if (ok) {
            float lo, hi;
            bool ab = false;
            float hl1 = 0.0f;
            if (MODE == MODE_FOV) {
                hl = in.highest_levels[idx];
                hl1 = FA(hl, 1.0f);
                lo = hl;
                hi = 0.0f;
            }
            count = filter_rect<MODE>(ws, s, cam.grid_x, hl1, false, 0ull, 0u, lo, hi, ab);
        }
        in.radii[idx] = count ? s.radius : 0;
        if (count) {
            visible = true;
            ws.geomA[idx] = make_float4(s.depth, __int_as_float(s.radius), s.len1, s.len2);
            ws.geomB[idx] = make_float4(s.e1x, s.e1y, s.e2x, s.e2y);
            const int R = (MODE == MODE_FOV) ? REC_FOV : REC_PS1;
            float4* rec = ws.rec + (size_t)R * idx;
            rec[0] = make_float4(s.px, s.py, s.conx, s.cony);
            if (MODE == MODE_FOV) {
                rec[1] = make_float4(s.conz, hl, 0.0f, 0.0f);
            } else {
                rec[1] = make_float4(s.conz, in.opacities[idx], 0.0f, 0.0f);  // colour filled by the emit pass
            }
            if (MODE == MODE_SUM) {
#pragma unroll
                for (int k = 0; k < 6; k++) ws.cov3D[6 * (size_t)idx + k] = c3[k];
            }
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, visible);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&ws.hdr->stats.num_visible, __popc(m));
}

// ------------------------------------------------------------------------------------------------------------------
// Tile scan: exclusive prefix sum of the per-tile histogram (one block); publishes N and the overflow flag.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_tile_scan(Workspace ws, int T) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    __shared__ uint32_t max_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_s = 0; max_s = 0; }
    __syncthreads();
    uint32_t local_max = 0;
    for (int base = 0; base < T; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = (i < T) ? ws.tile_count[i] : 0u;
        local_max = max(local_max, v);
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t incl = x + (wid ? warp_sums[wid - 1] : 0u) + carry;
        if (i < T) ws.tile_offset[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if (lane == 0) atomicMax(&max_s, local_max);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = carry_s;
        ws.tile_offset[T] = total;
        ws.hdr->stats.num_rendered = total;
        ws.hdr->stats.overflow = total > ws.hdr->cap ? 1u : 0u;
        ws.hdr->stats.max_tile_instances = max_s;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// SH colour (OBB/rasterizer_impl.cu:32-82, SUM/forward.cu:20-71, FOV/rasterizer_impl.cu:37-84)
// ------------------------------------------------------------------------------------------------------------------
// `sh` points at this Gaussian's coefficient triplets; `first` is the index of the degree-1 block:
// 1 for the PS1 layout (DC at 0), 0 for the FOV "rest" layout.  Returns Σ_{deg>=1}.
__device__ __forceinline__ float3 sh_rest_sum(const float* __restrict__ sh, int first, int deg, float x, float y, float z,
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

// ------------------------------------------------------------------------------------------------------------------
// Pass 2: emit (tile, depth, id) instances into the tile-binned key array + per-Gaussian colour.
// Replaces duplicateWithKeys (FOV/rasterizer_impl.cu:423-486) and compute_fov_colors (:490-530).
// ------------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) k_emit(Workspace ws, FrameInputs in) {
    __shared__ CamParams cam;
    load_cam(cam, ws.hdr);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= in.P) return;
    const int radius = in.radii[idx];
    if (radius <= 0) return;
    const int R = (MODE == MODE_FOV) ? REC_FOV : REC_PS1;
    float4* rec = ws.rec + (size_t)R * idx;
    const float4 r0 = rec[0];
    const float4 gA = ws.geomA[idx];
    const float4 gB = ws.geomB[idx];
    Splat s;
    s.px = r0.x; s.py = r0.y; s.depth = gA.x; s.radius = radius; s.len1 = gA.z; s.len2 = gA.w;
    s.e1x = gB.x; s.e1y = gB.y; s.e2x = gB.z; s.e2y = gB.w;
    get_rect(s.px, s.py, s.radius, cam.grid_x, cam.grid_y, s.x0, s.y0, s.x1, s.y1);
    float lo = 0.f, hi = 0.f, hl1 = 0.f;
    bool ab = false;
    if (MODE == MODE_FOV) {
        const float hl = rec[1].y;
        hl1 = FA(hl, 1.0f);
        lo = hl;
        hi = 0.0f;
    }
    const uint64_t key = ((uint64_t)__float_as_uint(s.depth) << 32) | (uint32_t)idx;
    filter_rect<MODE>(ws, s, cam.grid_x, hl1, true, key, ws.hdr->cap, lo, hi, ab);

    // ---- colour ----
    const float mx = in.means3D[3 * idx], my = in.means3D[3 * idx + 1], mz = in.means3D[3 * idx + 2];
    float dx = mx - cam.campos[0], dy = my - cam.campos[1], dz = mz - cam.campos[2];
    const float len = sqrtf(dx * dx + dy * dy + dz * dz);
    dx = dx / len; dy = dy / len; dz = dz / len;
    if (MODE == MODE_FOV) {
        float3 rest = make_float3(0.f, 0.f, 0.f);
        if (in.shs != nullptr && cam.M > 0) rest = sh_rest_sum(in.shs + (size_t)3 * cam.M * idx, 0, cam.sh_degree, dx, dy, dz, rest);
        rest.x += 0.5f; rest.y += 0.5f; rest.z += 0.5f;
        const int l0 = (int)lo;
        int l1 = (int)hi;
        if (ab) l1 = min(l1 + 1, FOV_LEVELS - 1);
        for (int l = l0; l <= l1; l++) {
            const float* dc = in.shs_dcs + (size_t)idx * 3 * FOV_LEVELS + l * 3;
            float4 o;
            o.x = in.opacities[(size_t)idx * FOV_LEVELS + l];
            o.y = fmaxf(SH_C0 * dc[0] + rest.x, 0.0f);
            o.z = fmaxf(SH_C0 * dc[1] + rest.y, 0.0f);
            o.w = fmaxf(SH_C0 * dc[2] + rest.z, 0.0f);
            rec[2 + l] = o;
        }
        // levels outside [l0,l1] are never composited (reference leaves them uninitialised, Q4); the blending
        // kernel may still *load* level l+1 of a skipped Gaussian, so keep those slots finite.
        for (int l = 0; l < FOV_LEVELS; l++)
            if (l < l0 || l > l1) rec[2 + l] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        float3 c;
        bool cl0 = false, cl1 = false, cl2 = false;
        if (in.colors_precomp != nullptr) {
            c = make_float3(in.colors_precomp[3 * idx], in.colors_precomp[3 * idx + 1], in.colors_precomp[3 * idx + 2]);
        } else {
            const float* sh = in.shs + (size_t)3 * cam.M * idx;
            // result = SH_C0*sh[0], higher bands accumulated onto it in the reference's order, then + 0.5
            c = sh_rest_sum(sh, 1, cam.sh_degree, dx, dy, dz, make_float3(SH_C0 * sh[0], SH_C0 * sh[1], SH_C0 * sh[2]));
            c.x += 0.5f; c.y += 0.5f; c.z += 0.5f;
            cl0 = c.x < 0; cl1 = c.y < 0; cl2 = c.z < 0;
            c.x = fmaxf(c.x, 0.0f); c.y = fmaxf(c.y, 0.0f); c.z = fmaxf(c.z, 0.0f);
        }
        float4 r1 = rec[1];
        r1.z = c.x; r1.w = c.y;
        rec[1] = r1;
        rec[2] = make_float4(c.z, 0.f, 0.f, 0.f);
        if (MODE == MODE_SUM) {
            uchar4 cl = make_uchar4(cl0, cl1, cl2, 0);
            reinterpret_cast<uchar4*>(ws.clamped)[idx] = cl;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Per-tile sort: block-local LSD radix sort on the 32 depth bits (uniform digits skipped), ties (equal depth
// bits) ordered by Gaussian id.  Output order == stable sort on (tile, depth_bits) of the id-ascending emission
// order == the reference's point_list.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tile_sort(Workspace ws, uint32_t* __restrict__ out_ranges,
                                                   uint32_t* __restrict__ out_point_list) {
    __shared__ uint32_t whist[8][256];
    __shared__ uint32_t totals[256];
    __shared__ uint32_t wsum[8];
    __shared__ unsigned long long vary_s;
    const int tile = blockIdx.x;
    const uint32_t cap = ws.hdr->cap;
    uint32_t sbeg = ws.tile_offset[tile], send = ws.tile_offset[tile + 1];
    if (out_ranges && threadIdx.x == 0) {
        // reference semantics: untouched tiles keep the memset value (0,0)
        out_ranges[2 * tile] = (send > sbeg) ? sbeg : 0u;
        out_ranges[2 * tile + 1] = (send > sbeg) ? send : 0u;
    }
    sbeg = min(sbeg, cap);
    send = min(send, cap);
    const uint32_t n = send - sbeg;
    if (n == 0) return;
    uint64_t* src = ws.keysA + sbeg;
    uint64_t* dst = ws.keysB + sbeg;
    uint32_t* out = ws.point_list + sbeg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* out2 = out_point_list ? out_point_list + sbeg : nullptr;
    if (n == 1) {
        if (tid == 0) {
            out[0] = (uint32_t)src[0];
            if (out2) out2[0] = (uint32_t)src[0];
        }
        return;
    }
    // which depth digits vary inside this segment?
    if (tid == 0) vary_s = 0ull;
    __syncthreads();
    {
        const uint64_t k0 = src[0];
        uint64_t v = 0;
        for (uint32_t i = tid; i < n; i += 256) v |= (src[i] ^ k0);
#pragma unroll
        for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicOr(&vary_s, (unsigned long long)v);
    }
    __syncthreads();
    const uint64_t vary = vary_s;
    // contiguous chunk per warp (multiple of 32 so that rounds stay warp-aligned)
    const uint32_t chunk = ((n + 7) / 8 + 31) & ~31u;
    const uint32_t wbeg = min(n, warp * chunk), wend = min(n, wbeg + chunk);
    for (int pass = 0; pass < 4; pass++) {
        const int shift = 32 + 8 * pass;
        if (((vary >> shift) & 0xffull) == 0) continue;
        for (int i = tid; i < 8 * 256; i += 256) (&whist[0][0])[i] = 0;
        __syncthreads();
        for (uint32_t i = wbeg + lane; i < wend; i += 32) atomicAdd(&whist[warp][(src[i] >> shift) & 0xff], 1u);
        __syncthreads();
        {   // thread d: exclusive scan over warps for digit d, then block scan over digits
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) { const uint32_t c = whist[w][tid]; whist[w][tid] = t; t += c; }
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
        for (int i = tid; i < 8 * 256; i += 256) (&whist[0][0])[i] += totals[i & 255];
        __syncthreads();
        for (uint32_t i0 = wbeg; i0 < wend; i0 += 32) {
            const uint32_t i = i0 + lane;
            const bool valid = i < wend;
            const uint64_t key = valid ? src[i] : 0ull;
            const uint32_t d = valid ? (uint32_t)((key >> shift) & 0xff) : (256u + lane);
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            uint32_t base = 0;
            if (valid) base = whist[warp][d];
            __syncwarp();
            if (valid) {
                dst[base + rank] = key;
                if (rank == 0) whist[warp][d] = base + __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        uint64_t* t = src; src = dst; dst = t;
    }
    // tie fix: runs of equal depth bits are ordered by id.  Heads are found first, then each head thread
    // insertion-sorts its (typically 2-element) run.
    {
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
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += 256) {
        const uint32_t id = (uint32_t)src[i];
        out[i] = id;
        if (out2) out2[i] = id;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Blend (v1: one 16x16 tile per CTA, one pixel per thread; arithmetic pinned to the reference binary).
// FOV plain tiles  : FOV/forward.cu:490-609   FOV blending tiles: FOV/forward.cu:262-476
// OBB              : OBB/forward.cu:251-384   SUM: SUM/forward.cu:298-430
// ------------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) k_blend(Workspace ws, FrameInputs in) {
    __shared__ float4 sA[256];   // px, py, conx, cony
    __shared__ float4 sB[256];   // conz, opacity(L1), (PS1: r, g)
    __shared__ float4 sC[256];   // PS1: b | FOV: colour L1 (x=op1,r,g,b)
    __shared__ float4 sD[256];   // FOV blending: level L2 (op2, r, g, b)
    __shared__ int sId[256];
    const FrameHeader* __restrict__ hdr = ws.hdr;
    const int W = hdr->cam.W, H = hdr->cam.H, gx = hdr->cam.grid_x;
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x;
    const int pxi = tx * TILE + (tid & 15), pyi = ty * TILE + (tid >> 4);
    const bool inside = pxi < W && pyi < H;
    const uint32_t pix_id = (uint32_t)W * pyi + pxi;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const uint32_t cap = hdr->cap;
    const uint32_t rbeg = min(ws.tile_offset[tile], cap), rend = min(ws.tile_offset[tile + 1], cap);
    const int total = (int)(rend - rbeg);
    const int rounds = (total + 255) / 256;
    int toDo = total;
    bool done = !inside;
    const float bg0 = hdr->bg[0], bg1 = hdr->bg[1], bg2 = hdr->bg[2];
    const size_t HW = (size_t)H * W;

    if (MODE == MODE_FOV) {
        const bool blending = ws.tile_blend[tile] != 0;
        const float tile_level_f = ws.tile_min[tile];   // Q2: kernels receive tile_level_min
        const int L1 = (int)tile_level_f;
        if (!blending) {
            float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
            for (int i = 0; i < rounds; i++, toDo -= 256) {
                if (__syncthreads_count(done) == 256) break;
                const int progress = i * 256 + tid;
                if (progress < total) {
                    const uint32_t id = ws.point_list[rbeg + progress];
                    const float4* rec = ws.rec + (size_t)REC_FOV * id;
                    sA[tid] = rec[0];
                    sB[tid] = rec[1];
                    sC[tid] = rec[2 + L1];
                }
                __syncthreads();
                const int lim = min(256, toDo);
                for (int j = 0; !done && j < lim; j++) {
                    const float4 a = sA[j];
                    const float dx = FS(a.x, pixx), dy = FS(a.y, pixy);
                    const float power = gauss_power(a.z, a.w, sB[j].x, dx, dy);
                    if (power > 0.0f || power < -4.5f) continue;
                    const float4 c = sC[j];
                    const float alpha = fminf(0.99f, FM(c.x, expf(power)));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = FM(T, FS(1.0f, alpha));
                    if (test_T < 0.0001f) { done = true; continue; }
                    const float w = FM(alpha, T);
                    C0 = FF(c.y, w, C0); C1 = FF(c.z, w, C1); C2 = FF(c.w, w, C2);
                    T = test_T;
                }
            }
            if (inside) {
                in.out_color[pix_id] = FF(bg0, T, C0);
                in.out_color[HW + pix_id] = FF(bg1, T, C1);
                in.out_color[2 * HW + pix_id] = FF(bg2, T, C2);
            }
        } else {
            const int L2 = L1 + 1;
            const float L2_f = FA(tile_level_f, 1.0f);
            const float dxl = (float)(tid & 15), dyl = (float)(tid >> 4);
            const float est = FF(FF(dxl, ws.tile_gx[tile], FM(dyl, ws.tile_gy[tile])), 0.0625f, tile_level_f);
            bool L1_done = est > (float)L2;
            bool L2_done = false;
            float T1 = 1.0f, T2 = 1.0f, A0 = 0.f, A1 = 0.f, A2 = 0.f, B0 = 0.f, B1 = 0.f, B2 = 0.f;
            for (int i = 0; i < rounds; i++, toDo -= 256) {
                if (__syncthreads_count(done) == 256) break;
                const int progress = i * 256 + tid;
                if (progress < total) {
                    const uint32_t id = ws.point_list[rbeg + progress];
                    const float4* rec = ws.rec + (size_t)REC_FOV * id;
                    sA[tid] = rec[0];
                    sB[tid] = rec[1];
                    sC[tid] = rec[2 + L1];
                    sD[tid] = rec[2 + L2];
                }
                __syncthreads();
                const int lim = min(256, toDo);
                for (int j = 0; !done && j < lim; j++) {
                    const float4 a = sA[j];
                    const float4 b = sB[j];
                    const float dx = FS(a.x, pixx), dy = FS(a.y, pixy);
                    const float power = gauss_power(a.z, a.w, b.x, dx, dy);
                    if (power > 0.0f || power < -4.5f) continue;
                    const float e = expf(power);
                    if (!L1_done) {
                        const float4 c = sC[j];
                        const float alpha1 = fminf(0.99f, FM(c.x, e));
                        if (!(alpha1 < 1.0f / 255.0f)) {
                            const float test_T1 = FM(T1, FS(1.0f, alpha1));
                            L1_done = test_T1 < 0.0001f;
                            if (!L1_done) {
                                const float w = FM(alpha1, T1);
                                A0 = FF(c.y, w, A0); A1 = FF(c.z, w, A1); A2 = FF(c.w, w, A2);
                                T1 = test_T1;
                            }
                        }
                    }
                    if (!L2_done) {
                        const float4 c = sD[j];
                        const float alpha2 = fminf(0.99f, FM(c.x, e));
                        const bool skip2 = (alpha2 < 1.0f / 255.0f) || (FA(b.y, 1.0f) < L2_f);
                        if (!skip2) {
                            const float test_T2 = FM(T2, FS(1.0f, alpha2));
                            L2_done = test_T2 < 0.0001f;
                            if (!L2_done) {
                                const float w = FM(alpha2, T2);
                                B0 = FF(c.y, w, B0); B1 = FF(c.z, w, B1); B2 = FF(c.w, w, B2);
                                T2 = test_T2;
                            }
                        }
                    }
                    if (L1_done && L2_done) { done = true; continue; }
                }
            }
            if (inside) {
                A0 = FF(bg0, T1, A0); A1 = FF(bg1, T1, A1); A2 = FF(bg2, T1, A2);
                B0 = FF(bg0, T2, B0); B1 = FF(bg1, T2, B1); B2 = FF(bg2, T2, B2);
                const float v = FS(est, FA((float)L1, kStartBlend));
                const float x = __saturatef(FA(fabsf(v), fabsf(v)));   // |v| / blend_width(0.5), clamped to [0,1]
                const float m3 = FM(x, FM(x, -3.0f));
                const float nb = FF(x, FM(x, FA(x, x)), m3);            // -(3x^2 - 2x^3)
                const float w1 = FA(nb, 1.0f);
                const float w2 = FS(1.0f, w1);
                in.out_color[pix_id] = FF(A0, w1, FM(B0, w2));
                in.out_color[HW + pix_id] = FF(A1, w1, FM(B1, w2));
                in.out_color[2 * HW + pix_id] = FF(A2, w1, FM(B2, w2));
            }
        }
        return;
    }

    // ---- PS=1 (OBB / SUM) ----
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t contributor = 0, last_contributor = 0;
    for (int i = 0; i < rounds; i++, toDo -= 256) {
        if (__syncthreads_count(done) == 256) break;
        const int progress = i * 256 + tid;
        if (progress < total) {
            const uint32_t id = ws.point_list[rbeg + progress];
            const float4* rec = ws.rec + (size_t)REC_PS1 * id;
            sA[tid] = rec[0];
            sB[tid] = rec[1];
            sC[tid] = rec[2];
            if (MODE == MODE_SUM) {
                sId[tid] = (int)id;
                atomicAdd(&in.gaussians_count[id], 1);
            }
        }
        __syncthreads();
        const int lim = min(256, toDo);
        for (int j = 0; !done && j < lim; j++) {
            contributor++;
            const float4 a = sA[j];
            const float4 b = sB[j];
            const float dx = FS(a.x, pixx), dy = FS(a.y, pixy);
            const float power = gauss_power(a.z, a.w, b.x, dx, dy);
            if (power > 0.0f || power < -4.5f) continue;
            const float alpha = fminf(0.99f, FM(b.y, expf(power)));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = FM(T, FS(1.0f, alpha));
            if (test_T < 0.0001f) { done = true; continue; }
            if (MODE == MODE_SUM) {
                // SUM accumulates (f*alpha)*T and records alpha*T per Gaussian (SUM/forward.cu:400-404)
                atomicAdd(&in.contributions[sId[j]], FM(alpha, T));
                C0 = FF(T, FM(alpha, b.z), C0);
                C1 = FF(T, FM(alpha, b.w), C1);
                C2 = FF(T, FM(alpha, sC[j].x), C2);
            } else {
                const float w = FM(alpha, T);
                C0 = FF(b.z, w, C0); C1 = FF(b.w, w, C1); C2 = FF(sC[j].x, w, C2);
            }
            T = test_T;
            last_contributor = contributor;
        }
    }
    if (inside) {
        if (MODE == MODE_SUM) {
            ws.final_T[pix_id] = T;
            ws.n_contrib[pix_id] = last_contributor;
        }
        in.out_color[pix_id] = FF(bg0, T, C0);
        in.out_color[HW + pix_id] = FF(bg1, T, C1);
        in.out_color[2 * HW + pix_id] = FF(bg2, T, C2);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// markVisible (FOV/rasterizer_impl.cu:407-419)
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_mark_visible(int P, const float* __restrict__ pts, const float* __restrict__ view, uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float z = xform_row(view, 2, pts[3 * idx], pts[3 * idx + 1], pts[3 * idx + 2]);
    present[idx] = z > 0.2f ? 1 : 0;
}

// export helpers for parity tests
__global__ void k_export_geometry(Workspace ws, int P, int mode, const int* __restrict__ radii_unused, float* means2D, float* depths,
                                  float* conic, float* cov3D, float* rgb, float* level_colors) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int R = (mode == MODE_FOV) ? REC_FOV : REC_PS1;
    const float4* rec = ws.rec + (size_t)R * idx;
    const float4 r0 = rec[0], r1 = rec[1];
    if (means2D) { means2D[2 * idx] = r0.x; means2D[2 * idx + 1] = r0.y; }
    if (depths) depths[idx] = ws.geomA[idx].x;
    if (conic) { conic[3 * idx] = r0.z; conic[3 * idx + 1] = r0.w; conic[3 * idx + 2] = r1.x; }
    if (cov3D && mode == MODE_SUM) for (int k = 0; k < 6; k++) cov3D[6 * idx + k] = ws.cov3D[6 * (size_t)idx + k];
    if (rgb && mode != MODE_FOV) { rgb[3 * idx] = r1.z; rgb[3 * idx + 1] = r1.w; rgb[3 * idx + 2] = rec[2].x; }
    if (level_colors && mode == MODE_FOV)
        for (int l = 0; l < FOV_LEVELS; l++) {
            const float4 c = rec[2 + l];
            level_colors[(idx * 4 + l) * 3 + 0] = c.y;
            level_colors[(idx * 4 + l) * 3 + 1] = c.z;
            level_colors[(idx * 4 + l) * 3 + 2] = c.w;
        }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

Workspace carve_workspace(void* base, int P, int W, int H, int64_t cap, Mode mode) {
    Workspace ws{};
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const size_t T = (size_t)gx * gy;
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t bytes) -> char* {
        char* p = b ? b + off : nullptr;
        off = align_up(off + bytes, 256);
        return p;
    };
    ws.hdr = (FrameHeader*)take(sizeof(FrameHeader));
    ws.tile_count = (uint32_t*)take(T * 4);
    ws.tile_offset = (uint32_t*)take((T + 1) * 4);
    ws.tile_cursor = (uint32_t*)take(T * 4);
    if (mode == MODE_FOV) {
        ws.tile_level = (float*)take(T * 4);
        ws.tile_min = (float*)take(T * 4);
        ws.tile_gx = (float*)take(T * 4);
        ws.tile_gy = (float*)take(T * 4);
        ws.tile_blend = (uint8_t*)take(T);
    }
    ws.geomA = (float4*)take((size_t)P * 16);
    ws.geomB = (float4*)take((size_t)P * 16);
    ws.rec = (float4*)take((size_t)P * 16 * (mode == MODE_FOV ? REC_FOV : REC_PS1));
    if (mode == MODE_SUM) {
        ws.cov3D = (float*)take((size_t)P * 24);
        ws.clamped = (uint8_t*)take((size_t)P * 4);
        ws.final_T = (float*)take((size_t)W * H * 4);
        ws.n_contrib = (uint32_t*)take((size_t)W * H * 4);
    }
    ws.keysA = (uint64_t*)take((size_t)cap * 8);
    ws.keysB = (uint64_t*)take((size_t)cap * 8);
    ws.point_list = (uint32_t*)take((size_t)cap * 4);
    ws.total_bytes = off;
    return ws;
}

// ---- optional stage timing (bench.py roofline): CUDA events between the stages of a frame ----
StageProfile g_prof;
static inline void prof_mark(int i, cudaStream_t st) {
    if (!g_prof.enabled) return;
    if (!g_prof.created) {
        for (int k = 0; k < StageProfile::N; k++) cudaEventCreate(&g_prof.ev[k]);
        g_prof.created = true;
    }
    cudaEventRecord(g_prof.ev[i], st);
    g_prof.valid = i + 1;
}

cudaError_t launch_setup(const Workspace& ws, const fovgs_camera& cam, int P, int M, Mode mode, const float* gaze,
                         float alpha, uint32_t cap, cudaStream_t st) {
    SetupArgs a;
    a.view = cam.viewmatrix; a.proj = cam.projmatrix; a.campos = cam.campos; a.bg = cam.bg; a.gaze = gaze;
    a.tanfovx = cam.tanfovx; a.tanfovy = cam.tanfovy;
    // reference: focal = dim / (2.0f * tanfov)   (FOV/rasterizer_impl.cu:656-657)
    a.focal_y = cam.image_height / (2.0f * cam.tanfovy);
    a.focal_x = cam.image_width / (2.0f * cam.tanfovx);
    a.scale_modifier = cam.scale_modifier; a.alpha = alpha;
    a.W = cam.image_width; a.H = cam.image_height;
    a.gx = (a.W + TILE - 1) / TILE; a.gy = (a.H + TILE - 1) / TILE;
    a.sh_degree = cam.sh_degree; a.M = M; a.P = P; a.tiles = a.gx * a.gy; a.cap = cap;
    const int T = a.tiles;
    prof_mark(0, st);
    k_setup<<<(T + 255) / 256, 256, 0, st>>>(ws, a);
    if (mode == MODE_FOV) {
        k_tile_levels<<<(T + 255) / 256, 256, 0, st>>>(T, ws.tile_level, gaze, a.W, a.H, a.gx, alpha);
        k_tile_infos<<<(T + 255) / 256, 256, 0, st>>>(T, ws.tile_level, a.gx, a.gy, ws.tile_gy, ws.tile_gx, ws.tile_min,
                                                     ws.tile_blend, ws.hdr);
    }
    return cudaGetLastError();
}

#define STAGE_CHECK()                                          \
    do {                                                       \
        cudaError_t e_ = cudaGetLastError();                   \
        if (e_ != cudaSuccess) return e_;                      \
        if (debug) {                                           \
            e_ = cudaStreamSynchronize(st);                    \
            if (e_ != cudaSuccess) return e_;                  \
        }                                                      \
    } while (0)

template <int MODE>
static cudaError_t forward_impl(const Workspace& ws, const FrameInputs& in, int W, int H, bool debug, cudaStream_t st) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const int T = gx * gy;
    const int pb = (in.P + 255) / 256;
    prof_mark(1, st);
    k_preprocess<MODE><<<pb, 256, 0, st>>>(ws, in);
    STAGE_CHECK();
    prof_mark(2, st);
    k_tile_scan<<<1, 1024, 0, st>>>(ws, T);
    STAGE_CHECK();
    prof_mark(3, st);
    k_emit<MODE><<<pb, 256, 0, st>>>(ws, in);
    STAGE_CHECK();
    prof_mark(4, st);
    k_tile_sort<<<T, 256, 0, st>>>(ws, in.out_ranges, in.out_point_list);
    STAGE_CHECK();
    prof_mark(5, st);
    k_blend<MODE><<<T, 256, 0, st>>>(ws, in);
    STAGE_CHECK();
    prof_mark(6, st);
    return cudaSuccess;
}

cudaError_t launch_forward(const Workspace& ws, const FrameInputs& in, int W, int H, Mode mode, bool debug, cudaStream_t st) {
    switch (mode) {
        case MODE_OBB: return forward_impl<MODE_OBB>(ws, in, W, H, debug, st);
        case MODE_SUM: return forward_impl<MODE_SUM>(ws, in, W, H, debug, st);
        default: return forward_impl<MODE_FOV>(ws, in, W, H, debug, st);
    }
}

cudaError_t launch_mark_visible(int P, const float* means3D, const float* view, const float* proj, uint8_t* present,
                                cudaStream_t st) {
    (void)proj;
    k_mark_visible<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, view, present);
    return cudaGetLastError();
}

cudaError_t launch_export_geometry(const Workspace& ws, int P, Mode mode, float* means2D, float* depths, float* conic,
                                   float* cov3D, float* rgb, float* level_colors, cudaStream_t st) {
    k_export_geometry<<<(P + 255) / 256, 256, 0, st>>>(ws, P, (int)mode, nullptr, means2D, depths, conic, cov3D, rgb, level_colors);
    return cudaGetLastError();
}

}  // namespace fovgs
