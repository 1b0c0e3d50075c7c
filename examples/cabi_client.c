/* cabi_client.c — a torch-free client of libfovgs.so: plain C99 + the CUDA runtime API.
 *
 * Shows (and tests: tests/test_cabi_client.py) that the drop-in boundary of include/fovgs.h needs nothing but device pointers,
 * sizes and a stream: the program reads a foveated scene + camera from a flat binary file, uploads it with cudaMemcpy, sizes the
 * workspace with fovgs_workspace_bytes, renders one frame with fovgs_forward_fov (growing the workspace once if the frame
 * statistics report an overflow), and writes the image and the radii back to a flat binary file.
 *
 * input file  (little endian): int32 {magic 0x46564753, P, M_rest, W, H, sh_degree}, float {tanfovx, tanfovy, alpha, gaze_x, gaze_y},
 *                              float bg[3], view[16], proj[16], campos[3],
 *                              float means3D[P*3], opacities[P*4], scales[P*3], rotations[P*4], shs_rest[P*M_rest*3],
 *                              shs_dcs[P*12], highest_levels[P]
 * output file: uint32 {num_rendered, num_visible}, float image[3*H*W], int32 radii[P]
 *
 *   cabi_client scene.bin out.bin [initial_instance_capacity]
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/fovgs.h"

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "cabi_client: %s failed: %s\n", #call, cudaGetErrorString(e_));        \
            return 3;                                                                              \
        }                                                                                          \
    } while (0)

static float* upload(FILE* f, size_t n, int* err) {
    float* d = NULL;
    if (n == 0) return NULL;
    float* h = (float*)malloc(n * sizeof(float));
    if (!h || fread(h, sizeof(float), n, f) != n) { *err = 1; free(h); return NULL; }
    if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) *err = 2;
    free(h);
    return d;
}

int main(int argc, char** argv) {
    if (argc != 3 && argc != 4) {
        fprintf(stderr, "usage: %s scene.bin out.bin [initial_instance_capacity]   (libfovgs ABI version %d)\n", argv[0], fovgs_version());
        return 2;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t hi[6];
    float hf[5];
    if (fread(hi, 4, 6, f) != 6 || fread(hf, 4, 5, f) != 5 || hi[0] != 0x46564753) {
        fprintf(stderr, "cabi_client: %s is not a scene file\n", argv[1]);
        return 2;
    }
    const int32_t P = hi[1], M_rest = hi[2], W = hi[3], H = hi[4];
    int err = 0;
    float* bg = upload(f, 3, &err);
    float* view = upload(f, 16, &err);
    float* proj = upload(f, 16, &err);
    float* campos = upload(f, 3, &err);
    float* means3D = upload(f, (size_t)P * 3, &err);
    float* opacities = upload(f, (size_t)P * 4, &err);
    float* scales = upload(f, (size_t)P * 3, &err);
    float* rotations = upload(f, (size_t)P * 4, &err);
    float* shs_rest = upload(f, (size_t)P * M_rest * 3, &err);
    float* shs_dcs = upload(f, (size_t)P * 12, &err);
    float* levels = upload(f, (size_t)P, &err);
    fclose(f);
    if (err) { fprintf(stderr, "cabi_client: short read or CUDA allocation failure (%d)\n", err); return 3; }
    float* gaze = NULL;
    CK(cudaMalloc((void**)&gaze, 8));
    CK(cudaMemcpy(gaze, hf + 3, 8, cudaMemcpyHostToDevice));

    float* color = NULL;
    int32_t* radii = NULL;
    CK(cudaMalloc((void**)&color, (size_t)3 * W * H * 4));
    CK(cudaMalloc((void**)&radii, (size_t)P * 4));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));

    fovgs_frame_stats stats;
    /* default capacity as the Python layer's: max(2^20, 8 P); a smaller one (argv[3]) exercises the overflow protocol */
    int64_t cap = argc == 4 ? atoll(argv[3]) : ((int64_t)8 * P > (1 << 20) ? (int64_t)8 * P : (1 << 20));
    if (cap <= 0) { fprintf(stderr, "cabi_client: bad capacity\n"); return 2; }
    for (int attempt = 0; attempt < 3; ++attempt) {
        const size_t bytes = fovgs_workspace_bytes(P, W, H, cap, /*foveated=*/1, /*ps1_mode=*/0);
        if (bytes == 0) { fprintf(stderr, "cabi_client: fovgs_workspace_bytes rejected the configuration\n"); return 3; }
        void* ws = NULL;
        CK(cudaMalloc(&ws, bytes));
        fovgs_fov_fwd_args a;
        memset(&a, 0, sizeof(a));
        a.struct_size = (uint32_t)sizeof(a); a.abi_version = FOVGS_VERSION;   /* FOVGS_ARGS_HEADER */
        a.cam.image_height = H; a.cam.image_width = W;
        a.cam.tanfovx = hf[0]; a.cam.tanfovy = hf[1];
        a.cam.scale_modifier = 1.0f; a.cam.sh_degree = hi[5];
        a.cam.bg = bg; a.cam.viewmatrix = view; a.cam.projmatrix = proj; a.cam.campos = campos;
        a.P = P; a.M_rest = M_rest;
        a.means3D = means3D; a.opacities = opacities; a.scales = scales; a.rotations = rotations;
        a.shs_rest = shs_rest; a.shs_dcs = shs_dcs; a.highest_levels = levels; a.gaze = gaze;
        a.alpha = hf[2]; a.blending = 1;
        a.out_color = color; a.radii = radii;
        a.workspace = ws; a.workspace_bytes = bytes; a.max_instances = cap;
        if (fovgs_forward_fov(&a, st) != 0 || fovgs_read_stats_async(ws, &stats, st) != 0) {
            fprintf(stderr, "cabi_client: %s\n", fovgs_last_error());
            return 3;
        }
        CK(cudaStreamSynchronize(st));
        CK(cudaFree(ws));
        if (!stats.overflow) break;
        fprintf(stderr, "cabi_client: %u instances exceed the capacity %lld, growing the workspace\n", stats.num_rendered, (long long)cap);
        cap = (int64_t)stats.num_rendered + stats.num_rendered / 4 + 1024;
    }
    if (stats.overflow) { fprintf(stderr, "cabi_client: still overflowing\n"); return 3; }

    float* h_color = (float*)malloc((size_t)3 * W * H * 4);
    int32_t* h_radii = (int32_t*)malloc((size_t)P * 4 + 4);
    CK(cudaMemcpy(h_color, color, (size_t)3 * W * H * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_radii, radii, (size_t)P * 4, cudaMemcpyDeviceToHost));
    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 2; }
    const uint32_t head[2] = {stats.num_rendered, stats.num_visible};
    fwrite(head, 4, 2, o);
    fwrite(h_color, 4, (size_t)3 * W * H, o);
    fwrite(h_radii, 4, (size_t)P, o);
    fclose(o);
    printf("cabi_client: P=%d %dx%d num_rendered=%u num_visible=%u\n", P, W, H, stats.num_rendered, stats.num_visible);
    return 0;
}
