// fovgs_api.cu — the extern "C" surface declared in include/fovgs.h (argument checking, workspace carving,
// launch orchestration).  No torch types, no allocation, no hidden state.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include "fovgs_internal.cuh"

using namespace fovgs;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
static int fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "CUDA error in %s: %s", where, cudaGetErrorString(e));
    return FOVGS_ERR_CUDA;
}

namespace fovgs {
// message setter for the entry points that live in other translation units (fovgs_step.cu)
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace fovgs

// the two-word header of every args struct (FOVGS_ARGS_HEADER): read before any other field
template <class A>
static int check_header(const A* a, const char* what) {
    if (a->struct_size != (uint32_t)sizeof(A) || a->abi_version != (uint32_t)FOVGS_VERSION) {
        snprintf(g_err, sizeof(g_err), "%s: args header mismatch (struct_size %u, abi_version %u; this library expects %u, %d): "
                 "the caller was built against another include/fovgs.h", what, a->struct_size, a->abi_version, (unsigned)sizeof(A), FOVGS_VERSION);
        return FOVGS_ERR_INVALID_ARG;
    }
    return 0;
}

extern "C" {

const char* fovgs_last_error(void) { return g_err; }
int fovgs_version(void) { return FOVGS_VERSION; }

size_t fovgs_struct_size(int32_t id) {
    switch (id) {
        case FOVGS_STRUCT_CAMERA: return sizeof(fovgs_camera);
        case FOVGS_STRUCT_FRAME_STATS: return sizeof(fovgs_frame_stats);
        case FOVGS_STRUCT_FOV_FWD_ARGS: return sizeof(fovgs_fov_fwd_args);
        case FOVGS_STRUCT_SMFR_FWD_ARGS: return sizeof(fovgs_smfr_fwd_args);
        case FOVGS_STRUCT_MMFR_FWD_ARGS: return sizeof(fovgs_mmfr_fwd_args);
        case FOVGS_STRUCT_PS1_FWD_ARGS: return sizeof(fovgs_ps1_fwd_args);
        case FOVGS_STRUCT_PS1_BWD_ARGS: return sizeof(fovgs_ps1_bwd_args);
        case FOVGS_STRUCT_ADAM_GROUP: return sizeof(fovgs_adam_group);
        default: return 0;
    }
}

size_t fovgs_workspace_bytes(int32_t P, int32_t W, int32_t H, int64_t max_instances, int32_t foveated, int32_t ps1_mode) {
    if (P < 0 || W <= 0 || H <= 0 || max_instances < 0) return 0;
    const Mode mode = foveated == 3 ? MODE_MMFR : foveated == 2 ? MODE_SMFR : foveated ? MODE_FOV : (ps1_mode != FOVGS_PS1_OBB ? MODE_SUM : MODE_OBB);
    return carve_workspace(nullptr, P, W, H, max_instances, mode).total_bytes;
}

static const int32_t kMaxP = 1 << 30;   // Gaussian ids travel in 32 bits next to the depth bits; loop counters are int

static int check_cam(const fovgs_camera& c) {
    if (c.image_width <= 0 || c.image_height <= 0) return fail(FOVGS_ERR_INVALID_ARG, "image size must be positive%s");
    if (!c.bg || !c.viewmatrix || !c.projmatrix || !c.campos)
        return fail(FOVGS_ERR_INVALID_ARG, "camera pointers (bg, viewmatrix, projmatrix, campos) must be non-null%s");
    return 0;
}

int fovgs_forward_fov(const fovgs_fov_fwd_args* a, void* stream) {
    if (!a) return fail(FOVGS_ERR_INVALID_ARG, "null args%s");
    if (int r = check_header(a, "fovgs_forward_fov")) return r;
    if (int r = check_cam(a->cam)) return r;
    cudaStream_t st = (cudaStream_t)stream;
    const int W = a->cam.image_width, H = a->cam.image_height;
    if (a->P < 0 || a->P > kMaxP) return fail(FOVGS_ERR_INVALID_ARG, "means3D must have dimensions (num_points, 3), num_points <= 2^30%s");
    if ((!a->out_color && !a->out_color_u8) || (a->P > 0 && !a->radii)) return fail(FOVGS_ERR_INVALID_ARG, "null output pointer%s");
    if (a->P == 0) return 0;  // reference: outputs stay at their initial value (rasterize_points.cu:103)
    if (!a->means3D || !a->opacities || !a->scales || !a->rotations || !a->shs_dcs || !a->highest_levels || !a->gaze)
        return fail(FOVGS_ERR_INVALID_ARG, "null input pointer (means3D/opacities/scales/rotations/shs_dcs/highest_levels/gaze)%s");
    if (a->M_rest > 0 && !a->shs_rest) return fail(FOVGS_ERR_INVALID_ARG, "shs_rest is null but M_rest > 0%s");
    if (a->max_instances <= 0 || a->max_instances > 0xffffffffll) return fail(FOVGS_ERR_INVALID_ARG, "max_instances out of range%s");
    Workspace ws = carve_workspace(a->workspace, a->P, W, H, a->max_instances, MODE_FOV);
    if (!a->workspace || a->workspace_bytes < ws.total_bytes) return fail(FOVGS_ERR_WORKSPACE, "workspace too small%s");
    cudaError_t e = launch_setup(ws, a->cam, a->P, a->M_rest, MODE_FOV, a->gaze, a->alpha, 0.0f, (uint32_t)a->max_instances, st);
    if (e != cudaSuccess) return fail_cuda(e, "setup");
    FrameInputs in{};
    in.P = a->P; in.M = a->M_rest;
    in.means3D = a->means3D; in.opacities = a->opacities; in.scales = a->scales; in.rotations = a->rotations;
    in.shs = a->M_rest > 0 ? a->shs_rest : nullptr; in.shs_dcs = a->shs_dcs; in.highest_levels = a->highest_levels;
    in.radii = a->radii; in.out_color = a->out_color; in.out_color_u8 = a->out_color_u8;
    in.out_ranges = a->out_ranges; in.out_point_list = a->out_point_list;
    in.early_stats_host = a->early_stats_host; in.early_stats_event = a->early_stats_event;
    in.packed_rows = (a->M_rest <= 15) ? a->packed_color_rows : nullptr;
    e = launch_forward(ws, in, W, H, MODE_FOV, a->cam.debug != 0, st);
    if (e != cudaSuccess) return fail_cuda(e, "forward_fov");
    return 0;
}

int fovgs_pack_color_rows(int32_t P, int32_t M_rest, const float* means3D, const float* shs_rest, const float* shs_dcs,
                          const float* opacities, float* rows, void* stream) {
    if (P < 0 || M_rest < 0 || M_rest > 15) return fail(FOVGS_ERR_INVALID_ARG, "pack_color_rows: P >= 0 and 0 <= M_rest <= 15%s");
    if (P == 0) return 0;
    if (!means3D || !shs_dcs || !opacities || !rows || (M_rest > 0 && !shs_rest))
        return fail(FOVGS_ERR_INVALID_ARG, "pack_color_rows: null pointer%s");
    if (((uintptr_t)rows) & 255) return fail(FOVGS_ERR_INVALID_ARG, "pack_color_rows: rows must be 256-byte aligned%s");
    cudaError_t e = launch_pack_color_rows(P, M_rest, means3D, shs_rest, shs_dcs, opacities, rows, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail_cuda(e, "pack_color_rows");
    return 0;
}

int fovgs_forward_smfr(const fovgs_smfr_fwd_args* a, void* stream) {
    if (!a) return fail(FOVGS_ERR_INVALID_ARG, "null args%s");
    if (int r = check_header(a, "fovgs_forward_smfr")) return r;
    if (int r = check_cam(a->cam)) return r;
    cudaStream_t st = (cudaStream_t)stream;
    const int W = a->cam.image_width, H = a->cam.image_height;
    if (a->P < 0 || a->P > kMaxP) return fail(FOVGS_ERR_INVALID_ARG, "means3D must have dimensions (num_points, 3), num_points <= 2^30%s");
    if (!a->out_color || (a->P > 0 && !a->radii)) return fail(FOVGS_ERR_INVALID_ARG, "null output pointer%s");
    if (a->P == 0) return 0;
    if (!a->means3D || !a->opacities || !a->scales || !a->rotations || !a->shs || !a->highest_levels || !a->gaze)
        return fail(FOVGS_ERR_INVALID_ARG, "null input pointer (means3D/opacities/scales/rotations/shs/highest_levels/gaze)%s");
    if (a->M <= 0) return fail(FOVGS_ERR_INVALID_ARG, "shs must hold at least the DC coefficient (M >= 1)%s");
    if (a->max_instances <= 0 || a->max_instances > 0xffffffffll) return fail(FOVGS_ERR_INVALID_ARG, "max_instances out of range%s");
    Workspace ws = carve_workspace(a->workspace, a->P, W, H, a->max_instances, MODE_SMFR);
    if (!a->workspace || a->workspace_bytes < ws.total_bytes) return fail(FOVGS_ERR_WORKSPACE, "workspace too small%s");
    cudaError_t e = launch_setup(ws, a->cam, a->P, a->M, MODE_SMFR, a->gaze, a->alpha, 0.0f, (uint32_t)a->max_instances, st);
    if (e != cudaSuccess) return fail_cuda(e, "setup");
    FrameInputs in{};
    in.P = a->P; in.M = a->M;
    in.means3D = a->means3D; in.opacities = a->opacities; in.scales = a->scales; in.rotations = a->rotations;
    in.shs = a->shs; in.highest_levels = a->highest_levels;
    in.radii = a->radii; in.out_color = a->out_color;
    in.out_ranges = a->out_ranges; in.out_point_list = a->out_point_list;
    in.early_stats_host = a->early_stats_host; in.early_stats_event = a->early_stats_event;
    e = launch_forward(ws, in, W, H, MODE_SMFR, a->cam.debug != 0, st);
    if (e != cudaSuccess) return fail_cuda(e, "forward_smfr");
    return 0;
}

int fovgs_forward_mmfr(const fovgs_mmfr_fwd_args* a, void* stream) {
    if (!a) return fail(FOVGS_ERR_INVALID_ARG, "null args%s");
    if (int r = check_header(a, "fovgs_forward_mmfr")) return r;
    if (int r = check_cam(a->cam)) return r;
    cudaStream_t st = (cudaStream_t)stream;
    const int W = a->cam.image_width, H = a->cam.image_height;
    if (a->P < 0 || a->P > kMaxP) return fail(FOVGS_ERR_INVALID_ARG, "means3D must have dimensions (num_points, 3), num_points <= 2^30%s");
    if (!a->out_color || (a->P > 0 && !a->radii)) return fail(FOVGS_ERR_INVALID_ARG, "null output pointer%s");
    if (a->P == 0) return 0;
    if (!a->means3D || !a->opacities || !a->scales || !a->rotations || !a->shs || !a->gaze)
        return fail(FOVGS_ERR_INVALID_ARG, "null input pointer (means3D/opacities/scales/rotations/shs/gaze)%s");
    if (a->M <= 0) return fail(FOVGS_ERR_INVALID_ARG, "shs must hold at least the DC coefficient (M >= 1)%s");
    if (a->max_instances <= 0 || a->max_instances > 0xffffffffll) return fail(FOVGS_ERR_INVALID_ARG, "max_instances out of range%s");
    Workspace ws = carve_workspace(a->workspace, a->P, W, H, a->max_instances, MODE_MMFR);
    if (!a->workspace || a->workspace_bytes < ws.total_bytes) return fail(FOVGS_ERR_WORKSPACE, "workspace too small%s");
    cudaError_t e = launch_setup(ws, a->cam, a->P, a->M, MODE_MMFR, a->gaze, a->alpha, a->cur_level, (uint32_t)a->max_instances, st);
    if (e != cudaSuccess) return fail_cuda(e, "setup");
    FrameInputs in{};
    in.P = a->P; in.M = a->M;
    in.means3D = a->means3D; in.opacities = a->opacities; in.scales = a->scales; in.rotations = a->rotations;
    in.shs = a->shs;
    in.radii = a->radii; in.out_color = a->out_color;
    in.out_ranges = a->out_ranges; in.out_point_list = a->out_point_list;
    in.early_stats_host = a->early_stats_host; in.early_stats_event = a->early_stats_event;
    e = launch_forward(ws, in, W, H, MODE_MMFR, a->cam.debug != 0, st);
    if (e != cudaSuccess) return fail_cuda(e, "forward_mmfr");
    return 0;
}

int fovgs_forward_ps1(const fovgs_ps1_fwd_args* a, void* stream) {
    if (!a) return fail(FOVGS_ERR_INVALID_ARG, "null args%s");
    if (int r = check_header(a, "fovgs_forward_ps1")) return r;
    if (int r = check_cam(a->cam)) return r;
    cudaStream_t st = (cudaStream_t)stream;
    const int W = a->cam.image_width, H = a->cam.image_height;
    if (a->mode < FOVGS_PS1_OBB || a->mode > FOVGS_PS1_VANILLA) return fail(FOVGS_ERR_INVALID_ARG, "bad ps1 mode%s");
    const Mode mode = a->mode != FOVGS_PS1_OBB ? MODE_SUM : MODE_OBB;
    if (a->mode == FOVGS_PS1_LWMC && !a->loss_map && a->P > 0)
        return fail(FOVGS_ERR_INVALID_ARG, "the loss-weighted variant needs loss_map [H,W]%s");
    if (a->P < 0 || a->P > kMaxP) return fail(FOVGS_ERR_INVALID_ARG, "means3D must have dimensions (num_points, 3), num_points <= 2^30%s");
    if (!a->out_color || (a->P > 0 && !a->radii)) return fail(FOVGS_ERR_INVALID_ARG, "null output pointer%s");
    if (a->P == 0) return 0;
    if (!a->means3D || !a->opacities) return fail(FOVGS_ERR_INVALID_ARG, "null input pointer (means3D/opacities)%s");
    if (!a->cov3D_precomp && (!a->scales || !a->rotations))
        return fail(FOVGS_ERR_INVALID_ARG, "provide either scales+rotations or cov3D_precomp%s");
    if (!a->colors_precomp && !(a->shs && a->M > 0))
        return fail(FOVGS_ERR_INVALID_ARG, "provide either shs (M>0) or colors_precomp%s");
    if (mode == MODE_SUM && (!a->gaussians_count || !a->contributions))
        return fail(FOVGS_ERR_INVALID_ARG, "the training variants need gaussians_count and contributions%s");
    if (a->max_instances <= 0 || a->max_instances > 0xffffffffll) return fail(FOVGS_ERR_INVALID_ARG, "max_instances out of range%s");
    Workspace ws = carve_workspace(a->workspace, a->P, W, H, a->max_instances, mode);
    if (!a->workspace || a->workspace_bytes < ws.total_bytes) return fail(FOVGS_ERR_WORKSPACE, "workspace too small%s");
    cudaError_t e = launch_setup(ws, a->cam, a->P, a->colors_precomp ? 0 : a->M, mode, nullptr, 0.0f, 0.0f, (uint32_t)a->max_instances, st,
                                 a->mode == FOVGS_PS1_VANILLA);
    if (e != cudaSuccess) return fail_cuda(e, "setup");
    FrameInputs in{};
    in.P = a->P; in.M = a->M;
    in.means3D = a->means3D; in.opacities = a->opacities; in.scales = a->scales; in.rotations = a->rotations;
    in.cov3D_precomp = a->cov3D_precomp; in.shs = a->colors_precomp ? nullptr : a->shs; in.colors_precomp = a->colors_precomp;
    in.radii = a->radii; in.gaussians_count = a->gaussians_count; in.contributions = a->contributions;
    in.loss_map = a->loss_map;
    in.stat = a->mode == FOVGS_PS1_MAX ? STAT_MAX : (a->mode == FOVGS_PS1_LWMC ? STAT_LWMC : STAT_SUM);
    in.out_color = a->out_color; in.out_ranges = a->out_ranges; in.out_point_list = a->out_point_list;
    in.early_stats_host = a->early_stats_host; in.early_stats_event = a->early_stats_event;
    e = launch_forward(ws, in, W, H, mode, a->cam.debug != 0, st);
    if (e != cudaSuccess) return fail_cuda(e, "forward_ps1");
    return 0;
}

int fovgs_backward_ps1(const fovgs_ps1_bwd_args* a, void* stream) {
    if (!a) return fail(FOVGS_ERR_INVALID_ARG, "null args%s");
    if (int r = check_header(a, "fovgs_backward_ps1")) return r;
    if (int r = check_cam(a->cam)) return r;
    cudaStream_t st = (cudaStream_t)stream;
    if (a->P == 0) return 0;
    if (!a->means3D || !a->radii || !a->dL_dout_color || !a->workspace)
        return fail(FOVGS_ERR_INVALID_ARG, "null input pointer (means3D/radii/dL_dout_color/workspace)%s");
    if (!a->dL_dmeans2D || !a->dL_dconic || !a->dL_dopacity || !a->dL_dcolors || !a->dL_dmeans3D || !a->dL_dcov3D)
        return fail(FOVGS_ERR_INVALID_ARG, "null gradient output pointer%s");
    if (a->shs && !a->dL_dsh) return fail(FOVGS_ERR_INVALID_ARG, "dL_dsh is null%s");
    if (a->scales && (!a->rotations || !a->dL_dscales || !a->dL_drotations))
        return fail(FOVGS_ERR_INVALID_ARG, "scale/rotation gradient outputs are null%s");
    const int W = a->cam.image_width, H = a->cam.image_height;
    Workspace ws = carve_workspace(const_cast<void*>(a->workspace), a->P, W, H, a->max_instances, MODE_SUM);
    if (a->workspace_bytes < ws.total_bytes) return fail(FOVGS_ERR_WORKSPACE, "workspace too small%s");
    fovgs_ps1_bwd_args b = *a;
    if (b.colors_precomp) b.shs = nullptr;
    cudaError_t e = launch_backward(ws, b, st);
    if (e != cudaSuccess) return fail_cuda(e, "backward_ps1");
    return 0;
}

int fovgs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                       void* stream) {
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !projmatrix || !present)))
        return fail(FOVGS_ERR_INVALID_ARG, "mark_visible: null pointer%s");
    if (P == 0) return 0;
    cudaError_t e = launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail_cuda(e, "mark_visible");
    return 0;
}

int fovgs_read_stats_async(const void* workspace, fovgs_frame_stats* stats_host, void* stream) {
    if (!workspace || !stats_host) return fail(FOVGS_ERR_INVALID_ARG, "read_stats: null pointer%s");
    cudaError_t e = cudaMemcpyAsync(stats_host, workspace, sizeof(fovgs_frame_stats), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail_cuda(e, "read_stats");
    return 0;
}

int fovgs_fov_tile_tables(const void* workspace, int32_t W, int32_t H, float* tile_level, float* tile_min, float* grad_x,
                          float* grad_y, uint8_t* blending, void* stream) {
    if (!workspace) return fail(FOVGS_ERR_INVALID_ARG, "tile_tables: null workspace%s");
    Workspace ws = carve_workspace(const_cast<void*>(workspace), 0, W, H, 0, MODE_FOV);
    const size_t T = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (tile_level && e == cudaSuccess) e = cudaMemcpyAsync(tile_level, ws.tile_level, T * 4, cudaMemcpyDeviceToDevice, st);
    if (tile_min && e == cudaSuccess) e = cudaMemcpyAsync(tile_min, ws.tile_min, T * 4, cudaMemcpyDeviceToDevice, st);
    if (grad_x && e == cudaSuccess) e = cudaMemcpyAsync(grad_x, ws.tile_gx, T * 4, cudaMemcpyDeviceToDevice, st);
    if (grad_y && e == cudaSuccess) e = cudaMemcpyAsync(grad_y, ws.tile_gy, T * 4, cudaMemcpyDeviceToDevice, st);
    if (blending && e == cudaSuccess) e = cudaMemcpyAsync(blending, ws.tile_blend, T, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail_cuda(e, "tile_tables");
    return 0;
}

int fovgs_ps1_geometry(const void* workspace, int32_t P, int32_t W, int32_t H, int32_t ps1_mode, float* means2D, float* depths,
                       float* conic, float* cov3D, float* rgb, void* stream) {
    if (!workspace) return fail(FOVGS_ERR_INVALID_ARG, "geometry: null workspace%s");
    const Mode mode = ps1_mode != FOVGS_PS1_OBB ? MODE_SUM : MODE_OBB;
    // tile tables and per-Gaussian arrays do not depend on the instance capacity
    Workspace ws = carve_workspace(const_cast<void*>(workspace), P, W, H, 0, mode);
    cudaError_t e = launch_export_geometry(ws, P, mode, means2D, depths, conic, cov3D, rgb, nullptr, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail_cuda(e, "geometry");
    return 0;
}

int fovgs_fov_geometry(const void* workspace, int32_t P, int32_t W, int32_t H, float* means2D, float* depths, float* conic,
                       float* level_colors, void* stream) {
    if (!workspace) return fail(FOVGS_ERR_INVALID_ARG, "geometry: null workspace%s");
    Workspace ws = carve_workspace(const_cast<void*>(workspace), P, W, H, 0, MODE_FOV);
    cudaError_t e = launch_export_geometry(ws, P, MODE_FOV, means2D, depths, conic, nullptr, nullptr, level_colors, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail_cuda(e, "geometry");
    return 0;
}

int fovgs_set_option(int32_t option, int32_t value) {
    switch (option) {
        case FOVGS_OPT_FULL_SORT: g_force_full_sort = value != 0; return 0;
        case FOVGS_OPT_NO_TMA: g_no_tma = value != 0; return 0;
        case FOVGS_OPT_NO_PDL: g_no_pdl = value != 0; return 0;
        case FOVGS_OPT_NO_DIRECT_STATS: g_no_direct_stats = value != 0; return 0;
        default: return fail(FOVGS_ERR_INVALID_ARG, "unknown option%s");
    }
}

int fovgs_profile_enable(int32_t on) {
    g_prof.enabled = on != 0;
    g_prof.blend_only = on == 2;
    g_prof.valid = 0;
    g_prof.frames = 0;
    return 0;
}

int fovgs_profile_count(void) { return g_prof.frames < StageProfile::SLOTS ? g_prof.frames : StageProfile::SLOTS; }

static int profile_read_slot(int slot, float* ms_out_host, int32_t n) {
    if (!ms_out_host || n < StageProfile::N - 1) return fail(FOVGS_ERR_INVALID_ARG, "profile_read: need room for 6 floats%s");
    if (!g_prof.created || g_prof.frames == 0 || g_prof.valid < StageProfile::N)
        return fail(FOVGS_ERR_INVALID_ARG, "profile_read: no profiled frame%s");
    cudaError_t e = cudaEventSynchronize(g_prof.ev[slot][StageProfile::N - 1]);
    if (e != cudaSuccess) return fail_cuda(e, "profile_read");
    for (int i = 0; i + 1 < StageProfile::N; i++) {
        if (g_prof.blend_only && i != StageProfile::N - 2) { ms_out_host[i] = 0.0f; continue; }   // only the blend was bracketed
        e = cudaEventElapsedTime(&ms_out_host[i], g_prof.ev[slot][i], g_prof.ev[slot][i + 1]);
        if (e != cudaSuccess) return fail_cuda(e, "profile_read");
    }
    return 0;
}

int fovgs_profile_read(float* ms_out_host, int32_t n) {
    return profile_read_slot((g_prof.frames > 0 ? g_prof.frames - 1 : 0) % StageProfile::SLOTS, ms_out_host, n);
}

int fovgs_profile_read_frame(int32_t k, float* ms_out_host, int32_t n) {
    if (k < 0 || k >= fovgs_profile_count()) return fail(FOVGS_ERR_INVALID_ARG, "profile_read_frame: no such frame%s");
    // frames older than SLOTS have been overwritten: index k counts from the oldest frame still held
    const int oldest = g_prof.frames > StageProfile::SLOTS ? g_prof.frames - StageProfile::SLOTS : 0;
    return profile_read_slot((oldest + k) % StageProfile::SLOTS, ms_out_host, n);
}

}  // extern "C"
