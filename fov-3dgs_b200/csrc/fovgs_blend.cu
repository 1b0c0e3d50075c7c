// fovgs_blend.cu — per-tile front-to-back alpha blend.
//
// One 16x16 tile per CTA, one pixel per thread, 256-entry batches staged through shared memory — the same
// traversal as the reference so that per-pixel results (and SUM's per-Gaussian `gaussians_count`, which counts
// fetched batches) are reproduced exactly; arithmetic is pinned to what the reference binary computes
// (fovgs_math.cuh: gauss_power, explicit FMUL/FFMA).  What is different:
//   * records: each Gaussian's blend inputs were packed by k_pre into 16-byte aligned float4 records with the
//     per-level opacity+colour already selected, so a batch entry is 3 (4 on blending tiles) 128-bit gathers
//     instead of 5-8 scalar gathers from four arrays (FOV/forward.cu:353-377);
//   * software pipelining: the gathers of batch i+1 are issued into registers before batch i is composited, so the
//     L2 gather latency (two dependent loads) is overlapped with compute instead of being exposed once per batch;
//   * scheduling: CTAs take tiles in descending instance-count class (tile_order from k_tile_scan), so the foveal
//     tiles — 10-20x the mean list length — start first instead of forming the tail of the kernel.
// Reference kernels: FOV plain tiles FOV/forward.cu:490-609, blending tiles :262-476, OBB OBB/forward.cu:251-384,
// SUM SUM/forward.cu:298-430.
#include "fovgs_internal.cuh"

namespace fovgs {

constexpr float kStartBlend = 0.5f;

template <int NREC>
struct Prefetch {
    float4 r[NREC];
    uint32_t id;
    bool valid;
};

template <int MODE>
__global__ void __launch_bounds__(256) k_blend(Workspace ws, FrameInputs in) {
    __shared__ float4 sA[256];   // px, py, conx, cony
    __shared__ float4 sB[256];   // conz, opacity | highest_level, depth
    __shared__ float4 sC[256];   // PS1: (r, g, b, -) | FOV: level L1 (opacity, r, g, b)
    __shared__ float4 sD[(MODE == MODE_FOV) ? 256 : 1];   // SMFR blending tiles reuse sC for both levels   // FOV blending tiles: level L2 (opacity, r, g, b)
    __shared__ int sId[(MODE == MODE_SUM) ? 256 : 1];
    const FrameHeader* __restrict__ hdr = ws.hdr;
    const int W = hdr->cam.W, H = hdr->cam.H, gx = hdr->cam.grid_x;
    const int tile = (int)ws.tile_order[blockIdx.x];
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
    const uint32_t* __restrict__ plist = ws.point_list + rbeg;
    int batches_done = 0;   // statistics: how much of the sorted list the tile actually consumed

    if (is_foveated(MODE)) {
        const bool blending = ws.tile_blend[tile] != 0;
        const float tile_level_f = ws.tile_min[tile];   // Q2: the render kernels receive tile_level_min
        const int L1 = (int)tile_level_f;
        constexpr int R = rec_size(MODE);
        constexpr bool SMFR = MODE == MODE_SMFR;          // one shared (opacity, r, g, b) record in slot 2
        const int S1 = shared_model(MODE) ? 2 : 2 + L1;
        if (MODE == MODE_MMFR && ws.tile_skip[tile]) {
            if (inside) { in.out_color[pix_id] = 0.f; in.out_color[HW + pix_id] = 0.f; in.out_color[2 * HW + pix_id] = 0.f; }
            return;
        }
        // MMFR blending tiles (mmfr_pcheck_obb/cuda_rasterizer/forward.cu:255-418) run the plain loop with two changes:
        // pixels that belong to the other level of the pair start out done, and the result is weighted at the end
        float mm_x = 0.0f;
        int mm_L1 = 0;
        const bool mm_blend = MODE == MODE_MMFR && blending;
        if (mm_blend) {
            const float est = FF(FF((float)(tid & 15), ws.tile_gx[tile], FM((float)(tid >> 4), ws.tile_gy[tile])), 0.0625f, tile_level_f);
            mm_L1 = (int)est;
            mm_x = FM(FS(est, FA((float)mm_L1, kStartBlend)), 2.0f);
            if (mm_x < 0.0f && (float)mm_L1 != hdr->cur_level) done = true;
        }
        if (!blending || mm_blend) {
            Prefetch<3> pf;
            auto fetch = [&](int progress) {
                pf.valid = progress < total;
                if (pf.valid) {
                    const uint32_t id = plist[progress];
                    const float4* __restrict__ rec = ws.rec + (size_t)R * id;
                    pf.r[0] = rec[0]; pf.r[1] = rec[1]; pf.r[2] = rec[S1];
                }
            };
            fetch(tid);
            float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
            for (int i = 0; i < rounds; i++, toDo -= 256) {
                if (__syncthreads_count(done) == 256) break;
                if (pf.valid) { sA[tid] = pf.r[0]; sB[tid] = pf.r[1]; sC[tid] = pf.r[2]; }
                batches_done = i + 1;
                __syncthreads();
                if (i + 1 < rounds) fetch((i + 1) * 256 + tid);
                const int lim = min(256, toDo);
                for (int j = 0; !done && j < lim; j++) {
                    const float4 a = sA[j];
                    const float dx = FS(a.x, pixx), dy = FS(a.y, pixy);
                    const float power = gauss_power(a.z, a.w, sB[j].x, dx, dy);
                    if (power > 0.0f || power < -4.5f) continue;
                    const float4 c = sC[j];
                    const float alpha = fminf(0.99f, FM(c.x, BLEND_EXP(power)));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = FM(T, FS(1.0f, alpha));
                    if (test_T < 0.0001f) { done = true; continue; }
                    const float w = FM(alpha, T);
                    C0 = FF(c.y, w, C0); C1 = FF(c.z, w, C1); C2 = FF(c.w, w, C2);
                    T = test_T;
                }
            }
            if (inside) {
                float used = 1.0f;
                if (mm_blend) {
                    const float x = fmaxf(0.0f, fminf(1.0f, mm_x));
                    const float m3 = FM(x, FM(x, -3.0f));
                    const float nb = FF(x, FM(x, FA(x, x)), m3);
                    const float w1 = FA(nb, 1.0f);
                    used = ((float)mm_L1 == hdr->cur_level) ? w1 : FS(1.0f, w1);
                }
                const float o0 = FF(bg0, T, C0), o1 = FF(bg1, T, C1), o2 = FF(bg2, T, C2);
                store_rgb(in, pix_id, HW, mm_blend ? FM(o0, used) : o0, mm_blend ? FM(o1, used) : o1, mm_blend ? FM(o2, used) : o2);
            }
            if (tid == 0 && batches_done) atomicAdd(&ws.hdr->stats.reserved[0], (uint32_t)min(total, batches_done * 256));
        } else {
            const int L2 = L1 + 1;
            const float L2_f = FA(tile_level_f, 1.0f);
            const float dxl = (float)(tid & 15), dyl = (float)(tid >> 4);
            const float est = FF(FF(dxl, ws.tile_gx[tile], FM(dyl, ws.tile_gy[tile])), 0.0625f, tile_level_f);
            bool L1_done = est > (float)L2;
            bool L2_done = false;
            Prefetch<4> pf;
            auto fetch = [&](int progress) {
                pf.valid = progress < total;
                if (pf.valid) {
                    const uint32_t id = plist[progress];
                    const float4* __restrict__ rec = ws.rec + (size_t)R * id;
                    pf.r[0] = rec[0]; pf.r[1] = rec[1]; pf.r[2] = rec[S1];
                    if (MODE == MODE_FOV) pf.r[3] = rec[S1 + 1];
                }
            };
            fetch(tid);
            float T1 = 1.0f, T2 = 1.0f, A0 = 0.f, A1 = 0.f, A2 = 0.f, B0 = 0.f, B1 = 0.f, B2 = 0.f;
            for (int i = 0; i < rounds; i++, toDo -= 256) {
                if (__syncthreads_count(done) == 256) break;
                if (pf.valid) { sA[tid] = pf.r[0]; sB[tid] = pf.r[1]; sC[tid] = pf.r[2]; if (MODE == MODE_FOV) sD[tid] = pf.r[3]; }
                batches_done = i + 1;
                __syncthreads();
                if (i + 1 < rounds) fetch((i + 1) * 256 + tid);
                const int lim = min(256, toDo);
                for (int j = 0; !done && j < lim; j++) {
                    const float4 a = sA[j];
                    const float4 b = sB[j];
                    const float dx = FS(a.x, pixx), dy = FS(a.y, pixy);
                    const float power = gauss_power(a.z, a.w, b.x, dx, dy);
                    if (power > 0.0f || power < -4.5f) continue;
                    const float e = BLEND_EXP(power);
                    if (SMFR) {
                        // naive_pcheck_obb/cuda_rasterizer/forward.cu:383-430: one alpha for both levels; a live L1 drops
                        // the entry for both when alpha < 1/255, a finished L1 lets it through to L2 untested
                        const float4 c = sC[j];
                        const float alpha1 = fminf(0.99f, FM(c.x, e));
                        if (!L1_done) {
                            if (alpha1 < 1.0f / 255.0f) continue;
                            const float test_T1 = FM(T1, FS(1.0f, alpha1));
                            L1_done = test_T1 < 0.0001f;
                            if (!L1_done) {
                                const float w = FM(alpha1, T1);
                                A0 = FF(c.y, w, A0); A1 = FF(c.z, w, A1); A2 = FF(c.w, w, A2);
                                T1 = test_T1;
                            }
                        }
                        if (!L2_done && !(FA(b.y, 1.0f) < L2_f)) {
                            const float test_T2 = FM(T2, FS(1.0f, alpha1));
                            L2_done = test_T2 < 0.0001f;
                            if (!L2_done) {
                                const float w = FM(alpha1, T2);
                                B0 = FF(c.y, w, B0); B1 = FF(c.z, w, B1); B2 = FF(c.w, w, B2);
                                T2 = test_T2;
                            }
                        }
                        if (L1_done && L2_done) done = true;
                        continue;
                    }
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
                store_rgb(in, pix_id, HW, FF(A0, w1, FM(B0, w2)), FF(A1, w1, FM(B1, w2)), FF(A2, w1, FM(B2, w2)));
            }
            if (tid == 0 && batches_done) atomicAdd(&ws.hdr->stats.reserved[0], (uint32_t)min(total, batches_done * 256));
        }
        return;
    }

    // ---- PS=1 (OBB / SUM) ----
    Prefetch<3> pf;
    auto fetch = [&](int progress) {
        pf.valid = progress < total;
        if (pf.valid) {
            const uint32_t id = plist[progress];
            const float4* __restrict__ rec = ws.rec + (size_t)REC_PS1 * id;
            pf.id = id;
            pf.r[0] = rec[0]; pf.r[1] = rec[1]; pf.r[2] = rec[2];
        }
    };
    fetch(tid);
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t contributor = 0, last_contributor = 0;
    // training family: which statistics (fovgs_lazy.cu explains the three reference variants); this full-sort path keeps
    // the reference's per-hit atomics — it serves parity runs (`out_point_list`), the lazy kernel is the fast one
    const int stat = (MODE == MODE_SUM) ? in.stat : STAT_SUM;
    const float cut = (MODE == MODE_SUM) ? ws.hdr->cam.falloff_cut : -4.5f;   // -inf: FOVGS_PS1_VANILLA
    int max_idx = 0;
    float max_contrib = 0.0f;
    for (int i = 0; i < rounds; i++, toDo -= 256) {
        if (__syncthreads_count(done) == 256) break;
        if (pf.valid) {
            sA[tid] = pf.r[0]; sB[tid] = pf.r[1]; sC[tid] = pf.r[2];
            if (MODE == MODE_SUM) {
                sId[tid] = (int)pf.id;
                if (stat != STAT_MAX) atomicAdd(&in.gaussians_count[pf.id], 1);   // counted when the batch is staged, as in the reference
            }
        }
        batches_done = i + 1;
        __syncthreads();
        if (i + 1 < rounds) fetch((i + 1) * 256 + tid);
        const int lim = min(256, toDo);
        for (int j = 0; !done && j < lim; j++) {
            contributor++;
            const float4 a = sA[j];
            const float4 b = sB[j];
            const float dx = FS(a.x, pixx), dy = FS(a.y, pixy);
            const float power = gauss_power(a.z, a.w, b.x, dx, dy);
            if (power > 0.0f || power < cut) continue;
            if (MODE == MODE_SUM && stat == STAT_MAX) atomicAdd(&in.gaussians_count[sId[j]], 1);
            const float alpha = fminf(0.99f, FM(b.y, BLEND_EXP(power)));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = FM(T, FS(1.0f, alpha));
            if (test_T < 0.0001f) { done = true; continue; }
            if (MODE == MODE_SUM) {
                // SUM accumulates (f*alpha)*T and records alpha*T per Gaussian (SUM/forward.cu:400-404)
                const float4 c = sC[j];
                const float contrib = FM(alpha, T);
                if (stat == STAT_SUM) atomicAdd(&in.contributions[sId[j]], contrib);
                else if (stat == STAT_MAX) atomicMax(reinterpret_cast<unsigned*>(&in.contributions[sId[j]]), __float_as_uint(contrib));
                else if (contrib > max_contrib) { max_contrib = contrib; max_idx = sId[j]; }
                C0 = FF(T, FM(alpha, c.x), C0);
                C1 = FF(T, FM(alpha, c.y), C1);
                C2 = FF(T, FM(alpha, c.z), C2);
            } else {
                const float4 c = sC[j];
                const float w = FM(alpha, T);
                C0 = FF(c.x, w, C0); C1 = FF(c.y, w, C1); C2 = FF(c.z, w, C2);
            }
            T = test_T;
            last_contributor = contributor;
        }
    }
    if (tid == 0 && batches_done) atomicAdd(&ws.hdr->stats.reserved[0], (uint32_t)min(total, batches_done * 256));
    if (inside) {
        if (MODE == MODE_SUM) {
            if (stat == STAT_LWMC) atomicAdd(&in.contributions[max_idx], in.loss_map[pix_id]);
            ws.final_T[pix_id] = T;
            ws.n_contrib[pix_id] = last_contributor;
        }
        in.out_color[pix_id] = FF(bg0, T, C0);
        in.out_color[HW + pix_id] = FF(bg1, T, C1);
        in.out_color[2 * HW + pix_id] = FF(bg2, T, C2);
    }
}

cudaError_t launch_blend(const Workspace& ws, const FrameInputs& in, int T, Mode mode, cudaStream_t st) {
    switch (mode) {
        case MODE_OBB: k_blend<MODE_OBB><<<T, 256, 0, st>>>(ws, in); break;
        case MODE_SUM: k_blend<MODE_SUM><<<T, 256, 0, st>>>(ws, in); break;
        case MODE_SMFR: k_blend<MODE_SMFR><<<T, 256, 0, st>>>(ws, in); break;
        case MODE_MMFR: k_blend<MODE_MMFR><<<T, 256, 0, st>>>(ws, in); break;
        default: k_blend<MODE_FOV><<<T, 256, 0, st>>>(ws, in); break;
    }
    return cudaGetLastError();
}

}  // namespace fovgs
