// fovgs_step.cu — the elementwise work either side of the rasterizer in a training step:
//   * the model's activations (fov3dgs/scene/gaussian_model.py:40-60,200-237: scaling = exp, rotation = normalize,
//     opacity = sigmoid) as ONE pass over the Gaussians, forward and backward;
//   * the optimizer update (scene/gaussian_model.py:279-289: torch.optim.Adam(l, lr=0.0, eps=1e-15), six parameter groups)
//     as ONE multi-tensor launch: per element 4 reads (param, grad, exp_avg, exp_avg_sq) and 3 writes = 28 B, where the
//     library optimizer's foreach path makes ~18 array passes.
// HBM-bound streaming kernels: 16-byte accesses, two independent float4 per array per thread in flight.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/fovgs.h"

namespace fovgs {

void set_error(const char* fmt, ...);

// ---------------------------------------------------------------- activations
// one thread per Gaussian; 8 floats in, 8 floats out.
__global__ void __launch_bounds__(256) k_activate_fwd(int P, const float* __restrict__ raw_scale, const float* __restrict__ raw_rot,
                                                      const float* __restrict__ raw_opacity, float* __restrict__ scale,
                                                      float* __restrict__ rot, float* __restrict__ opacity) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)P; i += (size_t)gridDim.x * blockDim.x) {
        if (scale) {
            scale[3 * i + 0] = expf(raw_scale[3 * i + 0]);
            scale[3 * i + 1] = expf(raw_scale[3 * i + 1]);
            scale[3 * i + 2] = expf(raw_scale[3 * i + 2]);
        }
        if (rot) {
            const float4 q = reinterpret_cast<const float4*>(raw_rot)[i];
            // torch.nn.functional.normalize: x / max(||x||_2, 1e-12)
            const float n = sqrtf(__fmaf_rn(q.w, q.w, __fmaf_rn(q.z, q.z, __fmaf_rn(q.y, q.y, __fmul_rn(q.x, q.x)))));
            const float d = fmaxf(n, 1e-12f);
            reinterpret_cast<float4*>(rot)[i] = make_float4(__fdiv_rn(q.x, d), __fdiv_rn(q.y, d), __fdiv_rn(q.z, d), __fdiv_rn(q.w, d));
        }
        if (opacity) opacity[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-raw_opacity[i])));
    }
}

// gradients w.r.t. the raw parameters from the gradients w.r.t. the activated ones.
//   exp      : d_raw = d * y
//   sigmoid  : d_raw = d * (1 - y) * y
//   normalize: y = x / max(n, eps);  d_raw = d / dn - x * (sum(d*x) / (dn*dn*n))   (second term only where n >= eps, n > 0)
__global__ void __launch_bounds__(256) k_activate_bwd(int P, const float* __restrict__ raw_rot, const float* __restrict__ scale,
                                                      const float* __restrict__ opacity, const float* __restrict__ d_scale,
                                                      const float* __restrict__ d_rot, const float* __restrict__ d_opacity,
                                                      float* __restrict__ d_raw_scale, float* __restrict__ d_raw_rot,
                                                      float* __restrict__ d_raw_opacity) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)P; i += (size_t)gridDim.x * blockDim.x) {
        if (d_raw_scale) {
            d_raw_scale[3 * i + 0] = __fmul_rn(d_scale[3 * i + 0], scale[3 * i + 0]);
            d_raw_scale[3 * i + 1] = __fmul_rn(d_scale[3 * i + 1], scale[3 * i + 1]);
            d_raw_scale[3 * i + 2] = __fmul_rn(d_scale[3 * i + 2], scale[3 * i + 2]);
        }
        if (d_raw_rot) {
            const float4 q = reinterpret_cast<const float4*>(raw_rot)[i];
            const float4 g = reinterpret_cast<const float4*>(d_rot)[i];
            const float n = sqrtf(__fmaf_rn(q.w, q.w, __fmaf_rn(q.z, q.z, __fmaf_rn(q.y, q.y, __fmul_rn(q.x, q.x)))));
            const float dn = fmaxf(n, 1e-12f);
            float4 o = make_float4(__fdiv_rn(g.x, dn), __fdiv_rn(g.y, dn), __fdiv_rn(g.z, dn), __fdiv_rn(g.w, dn));
            if (n >= 1e-12f && n > 0.0f) {
                const float dot = __fmaf_rn(g.w, q.w, __fmaf_rn(g.z, q.z, __fmaf_rn(g.y, q.y, __fmul_rn(g.x, q.x))));
                const float k = __fdiv_rn(dot, __fmul_rn(__fmul_rn(dn, dn), n));
                o.x = __fmaf_rn(-q.x, k, o.x);
                o.y = __fmaf_rn(-q.y, k, o.y);
                o.z = __fmaf_rn(-q.z, k, o.z);
                o.w = __fmaf_rn(-q.w, k, o.w);
            }
            reinterpret_cast<float4*>(d_raw_rot)[i] = o;
        }
        if (d_raw_opacity) {
            const float y = opacity[i];
            d_raw_opacity[i] = __fmul_rn(__fmul_rn(d_opacity[i], __fsub_rn(1.0f, y)), y);
        }
    }
}

// ---------------------------------------------------------------- Adam, all parameter groups in one launch
#define ADAM_THREADS 256
#define ADAM_CHUNK (ADAM_THREADS * 8)   // floats per CTA: two float4 per thread per array

struct AdamGroupDev {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    long long chunk_begin;  // first CTA of this group
    float w1, beta2, w2, bc2_sqrt, eps, neg_step;
    int vec_ok;             // all four pointers 16-byte aligned
    int pad;
};
struct AdamLaunch {
    AdamGroupDev g[FOVGS_ADAM_MAX_GROUPS];
    int n_groups;
};

// torch.optim.Adam (weight_decay = 0, amsgrad = False, maximize = False), torch/optim/adam.py _single_tensor_adam/_multi_tensor_adam:
//   exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
//   denom = exp_avg_sq.sqrt() / sqrt(1 - beta2^t) + eps;  param.addcdiv_(exp_avg, denom, value = -lr / (1 - beta1^t))
__device__ __forceinline__ void adam_one(const AdamGroupDev& G, float& p, float g, float& m, float& v) {
    const float diff = __fsub_rn(g, m);
    m = (G.w1 < 0.5f) ? __fmaf_rn(G.w1, diff, m) : __fsub_rn(g, __fmul_rn(diff, __fsub_rn(1.0f, G.w1)));
    v = __fmaf_rn(G.w2, __fmul_rn(g, g), __fmul_rn(v, G.beta2));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), G.bc2_sqrt), G.eps);
    p = __fmaf_rn(G.neg_step, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(ADAM_THREADS) k_adam(const __grid_constant__ AdamLaunch L) {
    int gi = 0;
#pragma unroll
    for (int k = 1; k < FOVGS_ADAM_MAX_GROUPS; ++k)
        if (k < L.n_groups && (long long)blockIdx.x >= L.g[k].chunk_begin) gi = k;
    const AdamGroupDev& G = L.g[gi];
    const long long base = ((long long)blockIdx.x - G.chunk_begin) * ADAM_CHUNK;
    const long long left = G.n - base;
    if (G.vec_ok && left >= ADAM_CHUNK) {
        const long long i0 = base / 4 + threadIdx.x, i1 = i0 + ADAM_THREADS;
        float4 p0 = reinterpret_cast<const float4*>(G.p)[i0], p1 = reinterpret_cast<const float4*>(G.p)[i1];
        const float4 g0 = reinterpret_cast<const float4*>(G.g)[i0], g1 = reinterpret_cast<const float4*>(G.g)[i1];
        float4 m0 = reinterpret_cast<const float4*>(G.m)[i0], m1 = reinterpret_cast<const float4*>(G.m)[i1];
        float4 v0 = reinterpret_cast<const float4*>(G.v)[i0], v1 = reinterpret_cast<const float4*>(G.v)[i1];
        adam_one(G, p0.x, g0.x, m0.x, v0.x); adam_one(G, p0.y, g0.y, m0.y, v0.y);
        adam_one(G, p0.z, g0.z, m0.z, v0.z); adam_one(G, p0.w, g0.w, m0.w, v0.w);
        adam_one(G, p1.x, g1.x, m1.x, v1.x); adam_one(G, p1.y, g1.y, m1.y, v1.y);
        adam_one(G, p1.z, g1.z, m1.z, v1.z); adam_one(G, p1.w, g1.w, m1.w, v1.w);
        reinterpret_cast<float4*>(G.p)[i0] = p0; reinterpret_cast<float4*>(G.p)[i1] = p1;
        reinterpret_cast<float4*>(G.m)[i0] = m0; reinterpret_cast<float4*>(G.m)[i1] = m1;
        reinterpret_cast<float4*>(G.v)[i0] = v0; reinterpret_cast<float4*>(G.v)[i1] = v1;
    } else {
        const long long end = left < ADAM_CHUNK ? G.n : base + ADAM_CHUNK;
        for (long long i = base + threadIdx.x; i < end; i += ADAM_THREADS) {
            float p = G.p[i], m = G.m[i], v = G.v[i];
            adam_one(G, p, G.g[i], m, v);
            G.p[i] = p; G.m[i] = m; G.v[i] = v;
        }
    }
}

}  // namespace fovgs

using namespace fovgs;

extern "C" int fovgs_activate_forward(int32_t P, const float* raw_scale, const float* raw_rot, const float* raw_opacity, float* scale,
                                      float* rot, float* opacity, void* stream) {
    if (P < 0 || (scale && !raw_scale) || (rot && !raw_rot) || (opacity && !raw_opacity)) {
        set_error("fovgs_activate_forward: an output is requested without its raw input");
        return FOVGS_ERR_INVALID_ARG;
    }
    if (P == 0 || (!scale && !rot && !opacity)) return 0;
    if (rot && ((((uintptr_t)raw_rot) | ((uintptr_t)rot)) & 15)) {
        set_error("fovgs_activate_forward: rotation arrays must be 16-byte aligned");
        return FOVGS_ERR_INVALID_ARG;
    }
    const int blocks = (P + 255) / 256 < 148 * 16 ? (P + 255) / 256 : 148 * 16;
    k_activate_fwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(P, raw_scale, raw_rot, raw_opacity, scale, rot, opacity);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("fovgs_activate_forward: %s", cudaGetErrorString(e)); return FOVGS_ERR_CUDA; }
    return 0;
}

extern "C" int fovgs_activate_backward(int32_t P, const float* raw_rot, const float* scale, const float* opacity, const float* d_scale,
                                       const float* d_rot, const float* d_opacity, float* d_raw_scale, float* d_raw_rot,
                                       float* d_raw_opacity, void* stream) {
    if (P < 0 || (d_raw_scale && (!scale || !d_scale)) || (d_raw_rot && (!raw_rot || !d_rot)) ||
        (d_raw_opacity && (!opacity || !d_opacity))) {
        set_error("fovgs_activate_backward: a gradient is requested without its inputs");
        return FOVGS_ERR_INVALID_ARG;
    }
    if (P == 0 || (!d_raw_scale && !d_raw_rot && !d_raw_opacity)) return 0;
    if (d_raw_rot && ((((uintptr_t)raw_rot) | ((uintptr_t)d_rot) | ((uintptr_t)d_raw_rot)) & 15)) {
        set_error("fovgs_activate_backward: rotation arrays must be 16-byte aligned");
        return FOVGS_ERR_INVALID_ARG;
    }
    const int blocks = (P + 255) / 256 < 148 * 16 ? (P + 255) / 256 : 148 * 16;
    k_activate_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(P, raw_rot, scale, opacity, d_scale, d_rot, d_opacity, d_raw_scale,
                                                             d_raw_rot, d_raw_opacity);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("fovgs_activate_backward: %s", cudaGetErrorString(e)); return FOVGS_ERR_CUDA; }
    return 0;
}

extern "C" int fovgs_adam_step(const fovgs_adam_group* groups, int32_t n_groups, void* stream) {
    if (n_groups < 0 || n_groups > FOVGS_ADAM_MAX_GROUPS || (n_groups > 0 && !groups)) {
        set_error("fovgs_adam_step: n_groups must be in [0, %d]", FOVGS_ADAM_MAX_GROUPS);
        return FOVGS_ERR_INVALID_ARG;
    }
    AdamLaunch L;
    L.n_groups = 0;
    long long chunks = 0;
    for (int k = 0; k < n_groups; ++k) {
        const fovgs_adam_group& a = groups[k];
        if (a.n < 0 || a.step < 1 || (a.n > 0 && (!a.param || !a.grad || !a.exp_avg || !a.exp_avg_sq))) {
            set_error("fovgs_adam_step: group %d has a null pointer, n < 0 or step < 1", k);
            return FOVGS_ERR_INVALID_ARG;
        }
        if (!(a.beta1 >= 0.0 && a.beta1 < 1.0 && a.beta2 >= 0.0 && a.beta2 < 1.0 && a.eps >= 0.0)) {
            set_error("fovgs_adam_step: group %d has betas outside [0, 1) or eps < 0", k);
            return FOVGS_ERR_INVALID_ARG;
        }
        if (a.n == 0) continue;
        AdamGroupDev& d = L.g[L.n_groups++];
        d.p = a.param; d.g = a.grad; d.m = a.exp_avg; d.v = a.exp_avg_sq; d.n = a.n; d.chunk_begin = chunks;
        // the scalars are formed in double exactly as torch/optim/adam.py forms them in Python floats, then narrowed once
        const double bc1 = 1.0 - pow(a.beta1, (double)a.step), bc2 = 1.0 - pow(a.beta2, (double)a.step);
        d.w1 = (float)(1.0 - a.beta1);
        d.beta2 = (float)a.beta2;
        d.w2 = (float)(1.0 - a.beta2);
        d.bc2_sqrt = (float)sqrt(bc2);
        d.eps = (float)a.eps;
        d.neg_step = (float)(-(a.lr / bc1));
        d.vec_ok = ((((uintptr_t)a.param) | ((uintptr_t)a.grad) | ((uintptr_t)a.exp_avg) | ((uintptr_t)a.exp_avg_sq)) & 15) == 0;
        d.pad = 0;
        chunks += (a.n + ADAM_CHUNK - 1) / ADAM_CHUNK;
    }
    if (chunks == 0) return 0;
    if (chunks > 0x7fffffffLL) { set_error("fovgs_adam_step: too many elements for one launch"); return FOVGS_ERR_INVALID_ARG; }
    k_adam<<<(unsigned)chunks, ADAM_THREADS, 0, (cudaStream_t)stream>>>(L);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("fovgs_adam_step: %s", cudaGetErrorString(e)); return FOVGS_ERR_CUDA; }
    return 0;
}
