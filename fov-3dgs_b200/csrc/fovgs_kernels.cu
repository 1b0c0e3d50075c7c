// fovgs_kernels.cu — frame setup (header, tile-level tables), workspace carving and the forward launch sequence:
//   k_setup || k_tile_levels, k_tile_infos -> k_pre -> k_tile_scan || k_color_tma -> k_scatter -> k_lazy_blend (x2, overlapped)
//   (with out_point_list / FOVGS_OPT_FULL_SORT: ... -> k_tile_sort_* -> k_blend);  || = programmatic-dependent-launch pair
//
// Design: binning is a two-level sort.  Level 1 is a counting sort by tile (histogram in k_pre, scan by
// the one-CTA k_tile_scan beside the colour kernel, cursor scatter) which yields the reference's `ranges` for free; level 2 orders each tile's segment by
// depth bits, ties by Gaussian id — lazily inside the blend kernel (fovgs_lazy.cu) or completely (fovgs_sort.cu) — which
// reproduces exactly the order of the reference's stable 45-bit global radix sort
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
    // second of a programmatic-dependent-launch pair with k_setup (which touches nothing this kernel reads or writes): the
    // wait below, before exit, keeps the stream's order for k_tile_infos
    struct ExitWait { __device__ ~ExitWait() { pdl_wait(); } } exit_wait;
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

// `mmfr` != 0: the multi-model baseline's variant — tile_min clamped at 0 before it is stored and classified
// (mmfr_pcheck_obb/cuda_rasterizer/rasterizer_impl.cu:249-251) and the per-level tile selection of
// compute_tile_skips_cuda (:277-304): skip unless cur_level - blend_width < tile_min < cur_level + 1.
__global__ void k_tile_infos(int T, const float* __restrict__ tile_levels, const int tile_width_num,
                             const int tile_height_num, float* __restrict__ grad_y, float* __restrict__ grad_x,
                             float* __restrict__ tile_level_min, uint8_t* __restrict__ tile_blendings,
                             FrameHeader* __restrict__ hdr, const int mmfr, const float cur_level,
                             uint8_t* __restrict__ tile_skips, uint8_t* __restrict__ tile_code) {
    auto idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool blend = false, skip = false;
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
        if (mmfr) {
            if (tile_min < 0) tile_min = 0;
            float lb, hb;
            lb = cur_level - kBlendWidth;
            hb = cur_level + 1;
            bool min_in = tile_min > lb && tile_min < hb;
            skip = !min_in;
            tile_skips[idx] = skip ? 1 : 0;
        }
        tile_level_min[idx] = tile_min;
        float tile_min_i = float(int(tile_min));
        blend = ((tile_min - tile_min_i) > kStartBlend && (tile_min_i < (FOV_LEVELS - 1)));
        tile_blendings[idx] = blend ? 1 : 0;
        // k_pre's per-candidate level test as one byte: the smallest h in {1,2,3,4} with tile_min < h (5: none); a Gaussian
        // whose highest level is an integer l passes `tile_min < l + 1` iff l + 1 >= code.  MMFR: 1 = tile of this call, 5 = skipped.
        tile_code[idx] = mmfr ? (uint8_t)(skip ? 5 : 1)
                              : (uint8_t)((tile_min < 1.0f) ? 1 : (tile_min < 2.0f) ? 2 : (tile_min < 3.0f) ? 3 : (tile_min < 4.0f) ? 4 : 5);
        grad_y[idx] = gy;
        grad_x[idx] = gx;
    }
    const unsigned m = __ballot_sync(0xffffffffu, blend);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&hdr->stats.num_blend_tiles, __popc(m));
    // bounding box of the tiles a Gaussian whose highest level is l can still land in: {tile_min < l + 1}
    {
        const bool in_range = idx < (uint64_t)T;
        const float tm = in_range ? tile_level_min[idx] : 0.0f;
        const int tx = in_range ? (int)(idx % tile_width_num) : 0, ty = in_range ? (int)(idx / tile_width_num) : 0;
#pragma unroll
        for (int l = 0; l < FOV_LEVELS; l++) {
            // MMFR: one box (slot 0) around the tiles this level's call renders
            const bool inl = in_range && (mmfr ? (l == 0 && !skip) : (tm < (float)(l + 1)));
            const int x0 = __reduce_min_sync(0xffffffffu, inl ? tx : 0x7fffffff);
            const int y0 = __reduce_min_sync(0xffffffffu, inl ? ty : 0x7fffffff);
            const int x1 = __reduce_max_sync(0xffffffffu, inl ? tx + 1 : -1);
            const int y1 = __reduce_max_sync(0xffffffffu, inl ? ty + 1 : -1);
            if ((threadIdx.x & 31) == 0 && x1 > 0) {
                atomicMin(&hdr->lvl_bbox[l][0], x0);
                atomicMin(&hdr->lvl_bbox[l][1], y0);
                atomicMax(&hdr->lvl_bbox[l][2], x1);
                atomicMax(&hdr->lvl_bbox[l][3], y1);
            }
        }
    }
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
    float tanfovx, tanfovy, focal_x, focal_y, scale_modifier, alpha, cur_level;
    int W, H, gx, gy, sh_degree, M, P, tiles, prefiltered, vanilla;
    uint32_t cap;
};

__global__ void k_setup(Workspace ws, SetupArgs a) {
    pdl_trigger();   // k_tile_levels (foveated frames) may run beside this kernel
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.tiles) ws.tile_count[(size_t)t * CSTRIDE] = 0;   // tile_cursor is initialised by the tile scan
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
            h->cam.prefiltered = a.prefiltered;
            h->cam.no_obb = a.vanilla;
            h->cam.falloff_cut = a.vanilla ? -INFINITY : -4.5f;
            h->alpha = a.alpha;
            h->cur_level = a.cur_level;
            h->P = a.P;
            h->tiles = a.tiles;
            h->cap = a.cap;
            h->stage_cursor = 0;
            h->pre_chunk = 0;
            h->vis_cursor = 0;
            h->exp_consts[0] = __uint_as_float(0x3bbb989du);
            h->exp_consts[1] = 252.0f;
            h->lazy_ticket[0] = h->lazy_ticket[1] = 0;
            h->lazy_count[0] = h->lazy_count[1] = 0;
        }
        if (i < FOV_LEVELS * 4) (&h->lvl_bbox[0][0])[i] = ((i & 3) < 2) ? 0x7fffffff : -1;
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
// markVisible (FOV/rasterizer_impl.cu:407-419)
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_mark_visible(int P, const float* __restrict__ pts, const float* __restrict__ view, uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float z = xform_row(view, 2, pts[3 * (size_t)idx], pts[3 * (size_t)idx + 1], pts[3 * (size_t)idx + 2]);
    present[idx] = z > 0.2f ? 1 : 0;
}

// export helpers for parity tests
__global__ void k_export_geometry(Workspace ws, int P, int mode, const int* __restrict__ radii_unused, float* means2D, float* depths,
                                  float* conic, float* cov3D, float* rgb, float* level_colors) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int R = rec_size(mode);
    const float4* rec = ws.rec + (size_t)R * idx;
    const float4 r0 = rec[0], r1 = rec[1];
    if (means2D) { means2D[2 * (size_t)idx] = r0.x; means2D[2 * (size_t)idx + 1] = r0.y; }
    if (depths) depths[idx] = r1.z;
    if (conic) { conic[3 * (size_t)idx] = r0.z; conic[3 * (size_t)idx + 1] = r0.w; conic[3 * (size_t)idx + 2] = r1.x; }
    if (cov3D && mode == MODE_SUM) for (int k = 0; k < 6; k++) cov3D[6 * (size_t)idx + k] = ws.cov3D[6 * (size_t)idx + k];
    if (rgb && !is_foveated(mode)) { const float4 c = rec[2]; rgb[3 * (size_t)idx] = c.x; rgb[3 * (size_t)idx + 1] = c.y; rgb[3 * (size_t)idx + 2] = c.z; }
    if (level_colors && mode == MODE_FOV)
        for (int l = 0; l < FOV_LEVELS; l++) {
            const float4 c = rec[2 + l];
            level_colors[((size_t)idx * 4 + l) * 3 + 0] = c.y;
            level_colors[((size_t)idx * 4 + l) * 3 + 1] = c.z;
            level_colors[((size_t)idx * 4 + l) * 3 + 2] = c.w;
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
    ws.tile_count = (uint32_t*)take(T * 4 * CSTRIDE);
    ws.tile_offset = (uint32_t*)take((T + 1) * 4);
    ws.tile_cursor = (uint32_t*)take(T * 4 * CSTRIDE);
    ws.tile_order = (uint32_t*)take(T * 4);
    ws.tile_order2 = (uint32_t*)take(T * 4);
    if (is_foveated(mode)) {
        ws.tile_level = (float*)take(T * 4);
        ws.tile_min = (float*)take(T * 4);
        ws.tile_gx = (float*)take(T * 4);
        ws.tile_gy = (float*)take(T * 4);
        ws.tile_blend = (uint8_t*)take(T);
        ws.tile_code = (uint8_t*)take((T + 15) / 16 * 16);
        if (mode == MODE_MMFR) ws.tile_skip = (uint8_t*)take(T);   // after the tables every foveated layout shares
    }
    ws.rec = (float4*)take((size_t)P * 16 * rec_size(mode));
    ws.vis_cap = (uint32_t)((size_t)P + (size_t)P / 4 + (size_t)STAGE_MAX_BLOCKS * 8 * 128);
    ws.vis_list = (uint32_t*)take((size_t)ws.vis_cap * 4);
    ws.vis_lv = (uint32_t*)take((size_t)ws.vis_cap * 4);
    if (mode == MODE_SUM) {
        ws.cov3D = (float*)take((size_t)P * 24);
        ws.clamped = (uint8_t*)take((size_t)P * 4);
        ws.final_T = (float*)take((size_t)W * H * 4);
        ws.n_contrib = (uint32_t*)take((size_t)W * H * 4);
    }
    // staging holds up to cap instances plus holes: < 1/16 per retired chunk, one open chunk per block
    const size_t stage_cap = cap ? (size_t)cap + (size_t)cap / 8 + (size_t)STAGE_MAX_BLOCKS * STAGE_CHUNK : 0;
    ws.stage_cap = (uint32_t)(stage_cap > 0xfffffff0ull ? 0xfffffff0ull : stage_cap);
    ws.keysA = (uint64_t*)take((size_t)cap * 8);
    ws.keysB = (uint64_t*)take(stage_cap * 8);
    ws.stage_key = ws.keysB;
    ws.stage_tile = (uint32_t*)take(stage_cap * 4);
    ws.point_list = (uint32_t*)take((size_t)cap * 4);
    ws.total_bytes = off;
    return ws;
}

// ---- optional stage timing (bench.py roofline): CUDA events between the stages of a frame ----
StageProfile g_prof;
bool g_no_direct_stats = false;   // fovgs_set_option(FOVGS_OPT_NO_DIRECT_STATS, 1): early statistics by cudaMemcpyAsync (A/B)
bool g_force_full_sort = false;   // fovgs_set_option(FOVGS_OPT_FULL_SORT, 1): always run the complete per-tile sort
static inline void prof_mark(int i, cudaStream_t st) {
    if (!g_prof.enabled) return;
    if (!g_prof.created) {
        for (int s = 0; s < StageProfile::SLOTS; s++)
            for (int k = 0; k < StageProfile::N; k++) cudaEventCreate(&g_prof.ev[s][k]);
        g_prof.created = true;
    }
    if (i == 0) g_prof.frames++;
    if (g_prof.blend_only && i < StageProfile::N - 2) { g_prof.valid = i + 1; return; }
    cudaEventRecord(g_prof.ev[(g_prof.frames - 1) % StageProfile::SLOTS][i], st);
    g_prof.valid = i + 1;
}

cudaError_t launch_setup(const Workspace& ws, const fovgs_camera& cam, int P, int M, Mode mode, const float* gaze,
                         float alpha, float cur_level, uint32_t cap, cudaStream_t st, bool vanilla) {
    SetupArgs a;
    a.view = cam.viewmatrix; a.proj = cam.projmatrix; a.campos = cam.campos; a.bg = cam.bg; a.gaze = gaze;
    a.tanfovx = cam.tanfovx; a.tanfovy = cam.tanfovy;
    // reference: focal = dim / (2.0f * tanfov)   (FOV/rasterizer_impl.cu:656-657)
    a.focal_y = cam.image_height / (2.0f * cam.tanfovy);
    a.focal_x = cam.image_width / (2.0f * cam.tanfovx);
    a.scale_modifier = cam.scale_modifier; a.alpha = alpha; a.cur_level = cur_level;
    a.W = cam.image_width; a.H = cam.image_height;
    a.gx = (a.W + TILE - 1) / TILE; a.gy = (a.H + TILE - 1) / TILE;
    a.sh_degree = cam.sh_degree; a.prefiltered = cam.prefiltered; a.vanilla = vanilla ? 1 : 0; a.M = M; a.P = P; a.tiles = a.gx * a.gy; a.cap = cap;
    const int T = a.tiles;
    prof_mark(0, st);
    k_setup<<<(T + 255) / 256, 256, 0, st>>>(ws, a);
    if (is_foveated(mode)) {
        launch_chained(true, k_tile_levels, dim3((T + 255) / 256), dim3(256), 0, st, T, ws.tile_level, gaze, a.W, a.H, a.gx, alpha);
        k_tile_infos<<<(T + 255) / 256, 256, 0, st>>>(T, ws.tile_level, a.gx, a.gy, ws.tile_gy, ws.tile_gx, ws.tile_min,
                                                     ws.tile_blend, ws.hdr, mode == MODE_MMFR ? 1 : 0, cur_level, ws.tile_skip, ws.tile_code);
    }
    return cudaGetLastError();
}

int device_sm_count() {
    static int sms[MAX_DEVICES] = {};
    const int dev = current_device_ordinal();
    if (sms[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = n > 0 ? n : 148;
    }
    return sms[dev];
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
    const int num_sms = device_sm_count();
    prof_mark(1, st);
    launch_pre(ws, in, (Mode)MODE, num_sms, st);
    STAGE_CHECK();
    prof_mark(2, st);
    // All variants sort lazily inside the blend kernel (fovgs_lazy.cu) unless the caller asked for the complete sorted
    // lists.  The training variant appends the sorted prefix it composites to point_list: exactly what its backward walks.
    const bool lazy = in.out_point_list == nullptr && in.out_ranges == nullptr && !g_force_full_sort;
    // early statistics: written by the scan kernel straight into the caller's pinned buffer when that is device-accessible
    // (cudaHostAlloc / torch pin_memory under unified addressing), else copied out behind the colour stage
    uint32_t* stats_host_dev = nullptr;
    if (in.early_stats_host != nullptr) {
        void* d = nullptr;
        if (g_no_direct_stats) d = nullptr;
        else if (cudaHostGetDevicePointer(&d, in.early_stats_host, 0) == cudaSuccess) stats_host_dev = (uint32_t*)d;
        else (void)cudaGetLastError();
    }
    launch_tile_scan(ws, !lazy, stats_host_dev, st);  // one CTA, an ordinary launch; triggers its dependent on entry
    launch_color(ws, in, (Mode)MODE, num_sms, st);    // its programmatic dependent: runs beside the scan
    STAGE_CHECK();
    if (in.early_stats_host != nullptr) {
        // instance count, overflow flag, visible count are final here: the host can have them well before the frame ends
        if (stats_host_dev == nullptr) {
            cudaError_t e_ = cudaMemcpyAsync(in.early_stats_host, ws.hdr, sizeof(fovgs_frame_stats), cudaMemcpyDeviceToHost, st);
            if (e_ != cudaSuccess) return e_;
        }
        if (in.early_stats_event != nullptr) {
            cudaError_t e_ = cudaEventRecord((cudaEvent_t)in.early_stats_event, st);
            if (e_ != cudaSuccess) return e_;
        }
    }
    prof_mark(3, st);
    launch_scatter(ws, num_sms, st);
    STAGE_CHECK();
    prof_mark(4, st);
    if (!lazy) {
        launch_tile_sort(ws, T, in.out_ranges, in.out_point_list, st);
        STAGE_CHECK();
    }
    prof_mark(5, st);
    if (lazy) launch_lazy_blend(ws, in, T, (Mode)MODE, st);
    else launch_blend(ws, in, T, (Mode)MODE, st);
    STAGE_CHECK();
    prof_mark(6, st);
    return cudaSuccess;
}

cudaError_t launch_forward(const Workspace& ws, const FrameInputs& in, int W, int H, Mode mode, bool debug, cudaStream_t st) {
    switch (mode) {
        case MODE_OBB: return forward_impl<MODE_OBB>(ws, in, W, H, debug, st);
        case MODE_SUM: return forward_impl<MODE_SUM>(ws, in, W, H, debug, st);
        case MODE_SMFR: return forward_impl<MODE_SMFR>(ws, in, W, H, debug, st);
        case MODE_MMFR: return forward_impl<MODE_MMFR>(ws, in, W, H, debug, st);
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
