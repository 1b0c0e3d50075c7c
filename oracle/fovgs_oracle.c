/* fovgs_oracle.c — CPU restatement of the reference rasterizer's algorithm.   *** TEST INFRASTRUCTURE ONLY ***
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or call
 * this file.  The product path (fov-3dgs_b200/) never does; it fails loudly when its CUDA library is missing.
 *
 * Parity pin: this oracle is checked against golden vectors produced by the UNMODIFIED reference CUDA extensions
 * (oracle/build_ref.py -> oracle/_ref/*.so, run on a B200 by tools/parity_gpu.py --golden; fixtures in
 * tests/golden/).  The reference ships no tests or golden vectors of its own.
 *
 * It follows the reference's *structure* (not this repo's CUDA design):
 *   preprocess            FOV/cuda_rasterizer/forward.cu:104-238 (+22-98), auxiliary.h:173-209,271-296
 *   tile levels / infos   FOV/cuda_rasterizer/rasterizer_impl.cu:120-177, 182-260 ; auxiliary.h:55-66
 *   filter / OBB_test     FOV/cuda_rasterizer/rasterizer_impl.cu:264-383 ; SUM/...:70-146 ; auxiliary.h:80-168
 *   duplicateWithKeys     FOV/cuda_rasterizer/rasterizer_impl.cu:423-486  (emission order = ascending Gaussian id)
 *   stable sort on (tile<<32 | depth bits), identifyTileRanges   :535-557, 843-871
 *   colours               FOV/...rasterizer_impl.cu:37-84,490-530 ; OBB/...:32-82 ; SUM/forward.cu:20-71
 *   blend                 FOV/forward.cu:262-476 (blending tiles), 490-609 ; OBB/forward.cu:251-384 ; SUM/forward.cu:298-430
 *   backward              SUM/cuda_rasterizer/backward.cu:20-557
 *
 * Floating point: compiled with -ffp-contract=off; fused multiply-adds are written explicitly (fmaf) where the
 * reference *binary* (nvcc 12.9 / ptxas, sm_100) fuses them, so that depth / means2D / radius / OBB decisions are
 * reproduced bit-for-bit on the index-critical chain.  expf/acosf/tanf are libm's here and libdevice's on the GPU
 * (<= 2 ulp apart); rsqrt is 1/sqrtf here and MUFU.RSQ there.
 *
 * Known, bounded disagreement with the reference BINARY (measured at 6 M Gaussians / 1080p, tools/oracle_diff.py): the OBB
 * eigenvectors are normalised with `rsqrtf` = MUFU.RSQ on the GPU (max relative error 2^-22.4, table-driven, not restatable
 * from public documentation) and with the correctly rounded 1/sqrtf here.  About one (Gaussian, tile) SAT decision in ten
 * million sits so close to its threshold that this last-place difference flips it (frame 0 of the bench workload: tile 4857,
 * Gaussian 1848335; frame 1: tile 7247, Gaussian 351137 — both kept by the reference binary and by libfovgs, dropped here).
 * With orc_set_ambiguity(1) the binning pass re-evaluates every near-threshold OBB decision with both normalisation factors
 * scaled by (1 +- 2^-21) and records the (tile, Gaussian) pairs whose outcome depends on it (orc_ambiguous()).  Full-size
 * checks then demand: instance sets equal outside that list.  tests/test_oracle_golden.py pins the two cases above.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define BLOCK_X 16
#define BLOCK_Y 16
#define BLOCK_SIZE 256
#define FOV_NUM 4

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                               0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

typedef struct orc_camera {
    int32_t W, H;
    float tanfovx, tanfovy;
    float scale_modifier;
    int32_t sh_degree;
    float bg[3];
    float view[16];
    float proj[16];
    float campos[3];
} orc_camera;

/* CUDA cvt.rzi.s32.f32: truncate, saturate, NaN -> 0 */
static inline int f2i_rz(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)(-2147483647 - 1);
    return (int)f;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* row r of the transposed 4x4 applied to (x,y,z,1): m[r]*x + m[4+r]*y + m[8+r]*z + m[12+r]   (auxiliary.h:190-209) */
static inline float xform_row(const float* m, int r, float x, float y, float z) {
    float t = y * m[4 + r];
    t = fmaf(x, m[r], t);
    t = fmaf(z, m[8 + r], t);
    return m[12 + r] + t;
}
static inline float dot3p(float ax, float ay, float az, float bx, float by, float bz) {
    return fmaf(az, bz, fmaf(ax, bx, ay * by));
}
/* auxiliary.h:173-176 (double math, one fused multiply-add in the reference binary) */
static inline float ndc2pix(float v, int S) { return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5); }

/* FOV/forward.cu:22-56 */
static void compute_cov3d(const float* s3, float mod, const float* q, float* c) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float s0 = mod * s3[0], s1 = mod * s3[1], s2 = mod * s3[2];
    const float xz = x * z, rx = r * x, rz = r * z, yy = y * y, zz = z * z;
    const float h02 = fmaf(r, y, xz), h20 = fmaf(-r, y, xz);
    const float h12 = fmaf(y, z, -rx), h21 = fmaf(y, z, rx);
    const float h01 = fmaf(x, y, -rz), h10 = fmaf(x, y, rz);
    const float a = yy + zz, b = fmaf(x, x, zz), d = fmaf(x, x, yy);
    const float R00 = 1.0f - (a + a), R11 = 1.0f - (b + b), R22 = 1.0f - (d + d);
    const float R01 = h01 + h01, R02 = h02 + h02, R10 = h10 + h10, R12 = h12 + h12, R20 = h20 + h20, R21 = h21 + h21;
    const float m00 = s0 * R00, m01 = s1 * R01, m02 = s2 * R02;
    const float m10 = s0 * R10, m11 = s1 * R11, m12 = s2 * R12;
    const float m20 = s0 * R20, m21 = s1 * R21, m22 = s2 * R22;
    c[0] = dot3p(m00, m01, m02, m00, m01, m02);
    c[1] = dot3p(m10, m11, m12, m00, m01, m02);
    c[2] = dot3p(m20, m21, m22, m00, m01, m02);
    c[3] = dot3p(m10, m11, m12, m10, m11, m12);
    c[4] = dot3p(m20, m21, m22, m10, m11, m12);
    c[5] = dot3p(m20, m21, m22, m20, m21, m22);
}

/* FOV/forward.cu:59-98 */
static void compute_cov2d(const orc_camera* cam, float fx, float fy, float tx, float ty, float tz, const float* c,
                          float* cxx, float* cxy, float* cyy) {
    const float* v = cam->view;
    const float limx = cam->tanfovx * 1.3f, limy = cam->tanfovy * 1.3f;
    const float txtz = tx / tz, tytz = ty / tz;
    const float clx = fminf(limx, fmaxf(-limx, txtz)), cly = fminf(limy, fmaxf(-limy, tytz));
    const float tz2 = tz * tz;
    const float J00 = fx / tz, J02 = (fx * (clx * -tz)) / tz2;
    const float J11 = fy / tz, J12 = (fy * (cly * -tz)) / tz2;
    const float T00 = fmaf(v[2], J02, v[0] * J00), T01 = fmaf(v[6], J02, v[4] * J00), T02 = fmaf(v[10], J02, v[8] * J00);
    const float T10 = fmaf(v[2], J12, J11 * v[1]), T11 = fmaf(v[6], J12, J11 * v[5]), T12 = fmaf(v[10], J12, J11 * v[9]);
    const float A00 = fmaf(T02, c[2], fmaf(T00, c[0], T01 * c[1]));
    const float A01 = fmaf(T12, c[2], fmaf(T10, c[0], T11 * c[1]));
    const float A10 = fmaf(T02, c[4], fmaf(T00, c[1], T01 * c[3]));
    const float A11 = fmaf(T12, c[4], fmaf(T10, c[1], T11 * c[3]));
    const float A20 = fmaf(T02, c[5], fmaf(T00, c[2], T01 * c[4]));
    const float A21 = fmaf(T12, c[5], fmaf(T10, c[2], T11 * c[4]));
    *cxx = fmaf(T02, A20, fmaf(T00, A00, T01 * A10)) + 0.3f;
    *cxy = fmaf(T02, A21, fmaf(T00, A01, T01 * A11));
    *cyy = fmaf(T12, A21, fmaf(T10, A01, T11 * A11)) + 0.3f;
}

static void get_rect(float px, float py, int radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
    const float rf = (float)radius;
    *x0 = imin(gx, imax(0, f2i_rz((px - rf) / BLOCK_X)));
    *y0 = imin(gy, imax(0, f2i_rz((py - rf) / BLOCK_Y)));
    *x1 = imin(gx, imax(0, f2i_rz((((px + rf) + 16.0f) - 1.0f) / BLOCK_X)));
    *y1 = imin(gy, imax(0, f2i_rz((((py + rf) + 16.0f) - 1.0f) / BLOCK_Y)));
}

typedef struct {
    float depth, px, py;
    int radius;
    float conx, cony, conz;
    float e1x, e1y, e2x, e2y, len1, len2;
    float cxy, a1, a2, n1, n2;   /* eigenvector inputs: e_k = (cxy * -q_k, a_k * q_k), q_k = rsqrt(n_k) */
    uint32_t tiles;   /* rect size, later exact count */
} splat_t;

/* preprocessCUDA (FOV/forward.cu:104-238).  Returns 0 when culled. */
static int preprocess_one(const orc_camera* cam, float fx, float fy, int gx, int gy, const float* mean, const float* scale,
                          const float* rot, float* cov3d, splat_t* s) {
    const float mx = mean[0], my = mean[1], mz = mean[2];
    const float tz = xform_row(cam->view, 2, mx, my, mz);
    if (tz <= 0.2f) return 0;
    const float hx = xform_row(cam->proj, 0, mx, my, mz), hy = xform_row(cam->proj, 1, mx, my, mz);
    const float hw = xform_row(cam->proj, 3, mx, my, mz);
    const float pw = 1.0f / (hw + 0.0000001f);
    const float ndcx = hx * pw, ndcy = hy * pw;
    compute_cov3d(scale, cam->scale_modifier, rot, cov3d);
    const float tx = xform_row(cam->view, 0, mx, my, mz), ty = xform_row(cam->view, 1, mx, my, mz);
    float cxx, cxy, cyy;
    compute_cov2d(cam, fx, fy, tx, ty, tz, cov3d, &cxx, &cxy, &cyy);
    const float bb = cxy * cxy;
    const float det = fmaf(cxx, cyy, -bb);
    if (det == 0.0f) return 0;
    const float det_inv = 1.0f / det;
    s->conx = cyy * det_inv; s->cony = cxy * -det_inv; s->conz = cxx * det_inv;
    const float mid = (cxx + cyy) * 0.5f;
    const float sq = sqrtf(fmaxf(fmaf(mid, mid, -det), 0.1f));
    const float l1 = mid + sq, l2 = mid - sq;
    const float rf = ceilf(sqrtf(fmaxf(l1, l2)) * 3.0f);
    s->radius = f2i_rz(rf);
    s->px = ndc2pix(ndcx, cam->W); s->py = ndc2pix(ndcy, cam->H);
    int x0, y0, x1, y1;
    get_rect(s->px, s->py, s->radius, gx, gy, &x0, &y0, &x1, &y1);
    const uint32_t tnum = (uint32_t)(y1 - y0) * (uint32_t)(x1 - x0);
    if (tnum == 0) return 0;
    s->e1x = s->e1y = s->e2x = s->e2y = s->len1 = s->len2 = 0.0f;
    if (tnum > 1) {
        const float a1 = cxx - l1, a2 = cxx - l2;
        const float q1 = 1.0f / sqrtf(fmaf(a1, a1, bb)), q2 = 1.0f / sqrtf(fmaf(a2, a2, bb));  /* GPU: rsqrt.approx */
        s->e1x = cxy * -q1; s->e1y = a1 * q1; s->e2x = cxy * -q2; s->e2y = a2 * q2;
        s->len1 = sqrtf(l1) * 3.0f; s->len2 = sqrtf(l2) * 3.0f;
        s->cxy = cxy; s->a1 = a1; s->a2 = a2; s->n1 = fmaf(a1, a1, bb); s->n2 = fmaf(a2, a2, bb);
    }
    s->depth = tz;
    s->tiles = tnum;
    return 1;
}

/* OBB_check (auxiliary.h:80-168) with the corner sums fused as in the reference binary */
typedef struct { float vx[4], vy[4]; } corners_t;
static void obb_corners(const splat_t* s, corners_t* o) {
    const float ax = fmaf(s->e1x, s->len1, s->px), bx = fmaf(-s->e1x, s->len1, s->px);
    const float ay = fmaf(s->e1y, s->len1, s->py), by = fmaf(-s->e1y, s->len1, s->py);
    o->vx[0] = fmaf(s->e2x, s->len2, ax); o->vy[0] = fmaf(s->e2y, s->len2, ay);
    o->vx[1] = fmaf(s->e2x, s->len2, bx); o->vy[1] = fmaf(s->e2y, s->len2, by);
    o->vx[2] = fmaf(-s->e2x, s->len2, bx); o->vy[2] = fmaf(-s->e2y, s->len2, by);
    o->vx[3] = fmaf(-s->e2x, s->len2, ax); o->vy[3] = fmaf(-s->e2y, s->len2, ay);
}
static int obb_check(const splat_t* s, const corners_t* o, float tcx, float tcy) {
    float mn = o->vx[0] - tcx, mx = mn;
    for (int i = 1; i < 4; i++) { float v = o->vx[i] - tcx; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    if (mx < -8.0f || mn > 8.0f) return 0;
    mn = o->vy[0] - tcy; mx = mn;
    for (int i = 1; i < 4; i++) { float v = o->vy[i] - tcy; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    if (mx < -8.0f || mn > 8.0f) return 0;
    const float rxp = (tcx + 8.0f) - s->px, rxm = (tcx + -8.0f) - s->px;
    const float ryp = (tcy + 8.0f) - s->py, rym = (tcy + -8.0f) - s->py;
    {
        const float yp = s->e1y * ryp, ym = s->e1y * rym;
        const float d0 = fmaf(s->e1x, rxp, yp), d1 = fmaf(s->e1x, rxm, yp), d2 = fmaf(s->e1x, rxm, ym), d3 = fmaf(s->e1x, rxp, ym);
        const float lo = fminf(fminf(d0, d1), fminf(d2, d3)), hi = fmaxf(fmaxf(d0, d1), fmaxf(d2, d3));
        if (s->len1 < lo || -s->len1 > hi) return 0;
    }
    {
        const float yp = s->e2y * ryp, ym = s->e2y * rym;
        const float d0 = fmaf(s->e2x, rxp, yp), d1 = fmaf(s->e2x, rxm, yp), d2 = fmaf(s->e2x, rxm, ym), d3 = fmaf(s->e2x, rxp, ym);
        const float lo = fminf(fminf(d0, d1), fminf(d2, d3)), hi = fmaxf(fmaxf(d0, d1), fmaxf(d2, d3));
        if (s->len2 < lo || -s->len2 > hi) return 0;
    }
    return 1;
}

/* ---- rsqrt sensitivity of an OBB decision (see the header comment) ------------------------------------------- */
static int g_ambiguity = 0;
static uint64_t* g_amb = NULL;       /* (tile << 32) | gaussian id of decisions that depend on the last place of rsqrt */
static int64_t g_amb_n = 0, g_amb_cap = 0;
void orc_set_ambiguity(int on) { g_ambiguity = on; }
int64_t orc_ambiguous(uint64_t* out, int64_t cap) {
    for (int64_t i = 0; i < g_amb_n && i < cap; i++) out[i] = g_amb[i];
    return g_amb_n;
}
static void amb_push(uint32_t tile, uint32_t id) {
    if (g_amb_n == g_amb_cap) { g_amb_cap = g_amb_cap ? 2 * g_amb_cap : 1024; g_amb = (uint64_t*)realloc(g_amb, sizeof(uint64_t) * g_amb_cap); }
    g_amb[g_amb_n++] = ((uint64_t)tile << 32) | id;
}
/* smallest distance of any SAT comparison of obb_check from its threshold, in pixels (exact arithmetic is irrelevant here:
 * this only selects the candidates worth re-evaluating) */
static float obb_margin(const splat_t* s, const corners_t* o, float tcx, float tcy) {
    float m = 1e30f;
    float mn = o->vx[0] - tcx, mx = mn;
    for (int i = 1; i < 4; i++) { float v = o->vx[i] - tcx; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    m = fminf(m, fminf(fabsf(mx + 8.0f), fabsf(mn - 8.0f)));
    mn = o->vy[0] - tcy; mx = mn;
    for (int i = 1; i < 4; i++) { float v = o->vy[i] - tcy; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    m = fminf(m, fminf(fabsf(mx + 8.0f), fabsf(mn - 8.0f)));
    const float rx[2] = {(tcx + 8.0f) - s->px, (tcx - 8.0f) - s->px}, ry[2] = {(tcy + 8.0f) - s->py, (tcy - 8.0f) - s->py};
    const float ex[2] = {s->e1x, s->e2x}, ey[2] = {s->e1y, s->e2y}, ln[2] = {s->len1, s->len2};
    for (int k = 0; k < 2; k++) {
        float lo = 1e30f, hi = -1e30f;
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) { const float d = ex[k] * rx[a] + ey[k] * ry[b]; lo = fminf(lo, d); hi = fmaxf(hi, d); }
        m = fminf(m, fminf(fabsf(ln[k] - lo), fabsf(-ln[k] - hi)));
    }
    return m;
}
/* 1 when the decision of obb_check changes under q_k -> q_k * (1 +- 2^-21) for any sign combination */
static int obb_rsqrt_sensitive(const splat_t* s, float tcx, float tcy, int nominal) {
    const float d = 4.76837158203125e-07f;   /* 2^-21 > MUFU.RSQ's 2^-22.4 + the half ulp of the rounded 1/sqrtf */
    const float q1 = 1.0f / sqrtf(s->n1), q2 = 1.0f / sqrtf(s->n2);
    for (int i = -1; i <= 1; i++)
        for (int j = -1; j <= 1; j++) {
            if (!i && !j) continue;
            splat_t t = *s;
            const float p1 = q1 * (1.0f + (float)i * d), p2 = q2 * (1.0f + (float)j * d);
            t.e1x = s->cxy * -p1; t.e1y = s->a1 * p1; t.e2x = s->cxy * -p2; t.e2y = s->a2 * p2;
            corners_t oc; obb_corners(&t, &oc);
            if (obb_check(&t, &oc, tcx, tcy) != nominal) return 1;
        }
    return 0;
}

/* ---- tile levels (FOV/rasterizer_impl.cu:86-177, auxiliary.h:26-66) ---------------------------------------- */
static const float real_image_width = 2.0f, real_viewing_distance = 1.0f, sqrt_max_ps = 3.4641016151377544f;
static const float start_blend = 0.5f;
static const float blend_width = 0.5f;   /* auxiliary.h:32 */

static void ncd2dir(float nx, float ny, float rw, float rh, float* d) {
    float x = (nx - 0.5f) * rw, y = (ny - 0.5f) * rh, z = real_viewing_distance;
    float n = sqrtf(fmaf(y, y, x * x) + z * z);
    d[0] = x / n; d[1] = y / n; d[2] = z / n;
}
static void tile_tables_impl(int W, int H, const float* gaze, float alpha, float* tile_level, float* tile_min, float* gxs,
                     float* gys, uint8_t* blending, int clamp0) {
    const int tw = (W + 15) / 16, th = (H + 15) / 16, T = tw * th;
    const float step = (float)((sqrt_max_ps - 1.) / (float)(FOV_NUM - 1));
    for (int idx = 0; idx < T; idx++) {
        const int ty = idx / tw, tx = idx % tw;
        const float px = (float)(tx * BLOCK_X + BLOCK_X / 2), py = (float)(ty * BLOCK_Y + BLOCK_Y / 2);
        const float rih = (float)H / (float)W * real_image_width;
        const float ncx = px / W, ncy = py / H;
        float td[3], gd[3], cd[3];
        ncd2dir(ncx, ncy, real_image_width, rih, td);
        ncd2dir(gaze[0], gaze[1], real_image_width, rih, gd);
        ncd2dir(0.5f, 0.5f, real_image_width, rih, cd);
        const float ecc = acosf(fmaf(gd[2], td[2], fmaf(gd[1], td[1], gd[0] * td[0])));
        const float eccc = acosf(fmaf(td[2], cd[2], fmaf(td[1], cd[1], td[0] * cd[0])));
        const float pr = alpha * ecc * ecc;
        const float amin = (float)(eccc - pr * 0.5), amax = (float)(eccc + pr * 0.5);
        const float ddx = (float)((ncx - 0.5) * real_image_width), ddy = (float)((ncy - 0.5) * rih);
        const float dist = sqrtf(fmaf(ddy, ddy, ddx * ddx) + 1.0f);
        const float major = (tanf(amax) - tanf(amin)) * real_viewing_distance;
        const float minor = 2.0f * dist * tanf(pr * 0.5f);
        const float area = (float)(3.14159265358979323846 * major * minor * 0.25f);
        const float r2p = W / real_image_width;
        const float ps = sqrtf(area) * r2p;
        float level;
        if (ps <= 1) level = 0; else level = (sqrtf(ps) - 1) / step;
        if (level > ((float)FOV_NUM - 0.1)) level = (float)((float)FOV_NUM - 0.1);
        tile_level[idx] = level;
    }
    for (int idx = 0; idx < T; idx++) {
        const int ty = idx / tw, tx = idx % tw;
        const float lv = tile_level[idx];
        float r = -1, l = -1, u = -1, d = -1;
        if (tx + 1 < tw) r = tile_level[(tx + 1) + tw * ty];
        if (tx - 1 >= 0) l = tile_level[(tx - 1) + tw * ty];
        if (ty + 1 < th) u = tile_level[tx + tw * (ty + 1)];
        if (ty - 1 >= 0) d = tile_level[tx + tw * (ty - 1)];
        float gx = 0, gy = 0;
        if (r != -1 && l != -1) gx = (r - l) / 2.0f; else if (r != -1) gx = r - lv; else if (l != -1) gx = lv - l;
        if (u != -1 && d != -1) gy = (u - d) / 2.0f; else if (u != -1) gy = u - lv; else if (d != -1) gy = lv - d;
        const float md = (float)(0.5 * (fabsf(gx) + fabsf(gy)));
        float tm = lv - md;
        if (clamp0 && tm < 0) tm = 0;   /* MMFR only: mmfr_pcheck_obb/cuda_rasterizer/rasterizer_impl.cu:249-251 */
        tile_min[idx] = tm;
        const float tmi = (float)f2i_rz(tm);
        blending[idx] = ((tm - tmi) > start_blend && (tmi < (FOV_NUM - 1))) ? 1 : 0;
        gys[idx] = gy; gxs[idx] = gx;
    }
}

void orc_tile_tables(int W, int H, const float* gaze, float alpha, float* tile_level, float* tile_min, float* gxs,
                     float* gys, uint8_t* blending) {
    tile_tables_impl(W, H, gaze, alpha, tile_level, tile_min, gxs, gys, blending, 0);
}

/* ---- SH colour ----------------------------------------------------------------------------------------------- */
/* `first`: index of the first degree-1 coefficient (1 for [P,16,3] tensors, 0 for the FOV rest tensor). */
static void sh_accumulate(const float* sh, int first, int deg, float x, float y, float z, float* res) {
#define C(k, ch) sh[3 * (first + (k)) + (ch)]
    if (deg > 0) {
        for (int ch = 0; ch < 3; ch++)
            res[ch] = res[ch] - SH_C1 * y * C(0, ch) + SH_C1 * z * C(1, ch) - SH_C1 * x * C(2, ch);
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            for (int ch = 0; ch < 3; ch++)
                res[ch] = res[ch] + SH_C2[0] * xy * C(3, ch) + SH_C2[1] * yz * C(4, ch) +
                          SH_C2[2] * (2.0f * zz - xx - yy) * C(5, ch) + SH_C2[3] * xz * C(6, ch) + SH_C2[4] * (xx - yy) * C(7, ch);
            if (deg > 2)
                for (int ch = 0; ch < 3; ch++)
                    res[ch] = res[ch] + SH_C3[0] * y * (3.0f * xx - yy) * C(8, ch) + SH_C3[1] * xy * z * C(9, ch) +
                              SH_C3[2] * y * (4.0f * zz - xx - yy) * C(10, ch) +
                              SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * C(11, ch) +
                              SH_C3[4] * x * (4.0f * zz - xx - yy) * C(12, ch) + SH_C3[5] * z * (xx - yy) * C(13, ch) +
                              SH_C3[6] * x * (xx - 3.0f * yy) * C(14, ch);
        }
    }
#undef C
}
static void view_dir(const orc_camera* cam, const float* mean, float* d) {
    float x = mean[0] - cam->campos[0], y = mean[1] - cam->campos[1], z = mean[2] - cam->campos[2];
    float n = sqrtf(x * x + y * y + z * z);
    d[0] = x / n; d[1] = y / n; d[2] = z / n;
}

/* ---- binning: keys, stable sort, ranges ------------------------------------------------------------------------- */
typedef struct { uint64_t key; uint32_t id; uint32_t seq_hi; uint64_t seq; } inst_t;
static int inst_cmp(const void* a, const void* b) {
    const inst_t* x = (const inst_t*)a; const inst_t* y = (const inst_t*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->seq != y->seq) return x->seq < y->seq ? -1 : 1;   /* stability: emission order */
    return 0;
}

typedef struct {
    int mode;  /* 0 obb, 1 sum, 2 fov, 3 mmfr (tile_skip instead of the level test) */
    int no_obb; /* vanilla diff-gaussian-rasterization: every tile of the rectangle gets an instance
                   (diff-gaussian-rasterization/cuda_rasterizer/rasterizer_impl.cu:70-110: duplicateWithKeys walks the whole rect) */
    const uint8_t* tile_skip;
    const orc_camera* cam;
    int P, M;
    const float *means3D, *opacity, *scales, *rot, *shs;
    const float *shs_dcs, *highest_levels;     /* fov */
    const float *tile_min; const uint8_t* tile_blend;  /* fov */
} bin_in_t;

typedef struct {
    splat_t* sp;          /* [P] */
    uint8_t* vis;         /* [P] passes preprocess */
    int* radii;           /* [P] */
    float* cov3d;         /* [P*6] */
    int32_t* lvl_lo; int32_t* lvl_hi;  /* fov level_ranges */
    inst_t* inst; int64_t n_inst;
} bin_out_t;

static int64_t run_binning(const bin_in_t* in, bin_out_t* o, float fx, float fy, int gx, int gy) {
    const int P = in->P;
    /* pass 1: preprocess */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        o->radii[i] = 0;
        o->vis[i] = (uint8_t)preprocess_one(in->cam, fx, fy, gx, gy, in->means3D + 3 * i, in->scales + 3 * i, in->rot + 4 * i,
                                            o->cov3d + 6 * (size_t)i, &o->sp[i]);
        if (o->vis[i]) o->radii[i] = o->sp[i].radius;
    }
    /* pass 2: filter / OBB_test -> exact tile lists; emission in ascending id, row-major tile order */
    g_amb_n = 0;
    int64_t cap = 1 << 20, n = 0;
    inst_t* inst = (inst_t*)malloc(sizeof(inst_t) * cap);
    for (int i = 0; i < P; i++) {
        if (!o->vis[i]) continue;
        splat_t* s = &o->sp[i];
        int x0, y0, x1, y1;
        get_rect(s->px, s->py, s->radius, gx, gy, &x0, &y0, &x1, &y1);
        const uint32_t tnum = (uint32_t)(y1 - y0) * (uint32_t)(x1 - x0);
        uint32_t count = 0;
        float hl = 0, lo = 0, hi = 0; int be_blend = 0;
        if (in->mode == 2) { hl = in->highest_levels[i]; lo = hl; hi = 0; }
        const uint64_t dbits = fbits(s->depth);
        if (n + tnum + 1 > cap) { while (n + tnum + 1 > cap) cap *= 2; inst = (inst_t*)realloc(inst, sizeof(inst_t) * cap); }
        if (tnum == 1) {
            const uint32_t tile = (uint32_t)y0 * gx + x0;
            int pass = 1;
            if (in->mode == 2) {
                const float level = in->tile_min[tile];
                pass = level < (hl + 1);
                if (pass) { lo = level; hi = level; be_blend = in->tile_blend[tile] || be_blend; }
            }
            if (in->mode == 3) pass = !in->tile_skip[tile];
            if (pass) { count = 1; inst[n].key = ((uint64_t)tile << 32) | dbits; inst[n].id = i; inst[n].seq = n; n++; }
        } else {
            corners_t oc; obb_corners(s, &oc);
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) {
                    const uint32_t tile = (uint32_t)y * gx + x;
                    float level = 0;
                    if (in->mode == 2) { level = in->tile_min[tile]; if (!(level < (hl + 1))) continue; }
                    if (in->mode == 3 && in->tile_skip[tile]) continue;
                    const float tcx = (float)x * (float)BLOCK_X + (float)BLOCK_X / 2.0f;
                    const float tcy = (float)y * (float)BLOCK_Y + (float)BLOCK_Y / 2.0f;
                    const int hit = in->no_obb ? 1 : obb_check(s, &oc, tcx, tcy);
                    if (g_ambiguity && !in->no_obb) {
                        /* a perturbation of 2^-21 moves corners / projections by at most (len + |rel|) * 2^-21 pixels */
                        const float reach = (s->len1 + s->len2 + fabsf(tcx - s->px) + fabsf(tcy - s->py) + 16.0f) * 2e-6f + 1e-4f;
                        if (obb_margin(s, &oc, tcx, tcy) <= reach && obb_rsqrt_sensitive(s, tcx, tcy, hit)) amb_push(tile, (uint32_t)i);
                    }
                    if (!hit) continue;
                    count++;
                    if (in->mode == 2) { lo = fminf(lo, level); hi = fmaxf(hi, level); be_blend = in->tile_blend[tile] || be_blend; }
                    inst[n].key = ((uint64_t)tile << 32) | dbits; inst[n].id = i; inst[n].seq = n; n++;
                }
        }
        s->tiles = count;
        if (count == 0) o->radii[i] = 0;
        else if (in->mode == 2) {
            o->lvl_lo[i] = f2i_rz(lo);
            int h = f2i_rz(hi);
            if (be_blend) h = imin(h + 1, FOV_NUM - 1);
            o->lvl_hi[i] = h;
        }
    }
    qsort(inst, (size_t)n, sizeof(inst_t), inst_cmp);
    o->inst = inst; o->n_inst = n;
    return n;
}

static void fill_lists(const bin_out_t* o, int T, uint32_t* point_list, int64_t cap, uint32_t* ranges) {
    const int64_t n = o->n_inst;
    if (ranges) memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)T);
    for (int64_t i = 0; i < n; i++) {
        if (point_list && i < cap) point_list[i] = o->inst[i].id;
        if (ranges) {
            const uint32_t t = (uint32_t)(o->inst[i].key >> 32);
            if (i == 0) ranges[2 * t] = 0;
            else {
                const uint32_t pt = (uint32_t)(o->inst[i - 1].key >> 32);
                if (t != pt) { ranges[2 * pt + 1] = (uint32_t)i; ranges[2 * t] = (uint32_t)i; }
            }
            if (i == n - 1) ranges[2 * t + 1] = (uint32_t)n;
        }
    }
}

/* ---- tile-parallel driver (pthreads; dynamic scheduling over tiles).  ORC_THREADS overrides the core count. ---- */
typedef void (*tile_fn)(void* ctx, int tile);
typedef struct { tile_fn fn; void* ctx; int T; int next; } tile_job_t;
static void* tile_worker(void* arg) {
    tile_job_t* j = (tile_job_t*)arg;
    for (;;) {
        const int t = __atomic_fetch_add(&j->next, 4, __ATOMIC_RELAXED);
        if (t >= j->T) break;
        for (int k = t; k < t + 4 && k < j->T; k++) j->fn(j->ctx, k);
    }
    return NULL;
}
int orc_num_threads(void) {
    const char* e = getenv("ORC_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}
static void run_tiles(tile_fn fn, void* ctx, int T) {
    tile_job_t job = {fn, ctx, T, 0};
    const int nt = orc_num_threads();
    if (nt <= 1) { tile_worker(&job); return; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nt - 1; i++) if (pthread_create(&th[started], NULL, tile_worker, &job) == 0) started++;
    tile_worker(&job);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
}
static inline void atomic_add_f32(float* p, float v) {
    uint32_t old = __atomic_load_n((uint32_t*)p, __ATOMIC_RELAXED), neu;
    do { float f; memcpy(&f, &old, 4); f += v; memcpy(&neu, &f, 4); }
    while (!__atomic_compare_exchange_n((uint32_t*)p, &old, neu, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

/* atomicMaxFloat of the MAX variant (pcheck_obb_max/cuda_rasterizer/auxiliary.h:41-50): CAS loop around fmaxf */
static inline void atomic_max_f32(float* p, float v) {
    uint32_t old = __atomic_load_n((uint32_t*)p, __ATOMIC_RELAXED), neu;
    do { float f; memcpy(&f, &old, 4); f = fmaxf(v, f); memcpy(&neu, &f, 4); }
    while (!__atomic_compare_exchange_n((uint32_t*)p, &old, neu, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

static inline float gauss_power(float conx, float cony, float conz, float dx, float dy) {
    const float a = dy * (dy * conz);
    const float s = fmaf(dx, dx * conx, a);
    const float c = dy * (dx * cony);
    return fmaf(s, -0.5f, -c);
}

typedef struct {
    const orc_camera* cam; int mode, W, H, gx; const uint32_t* rng; const inst_t* inst; const splat_t* sp;
    const float* opacity; const float* rgb; int* gaussians_count; float* contributions; float* final_T; uint32_t* n_contrib;
    float* out_color; const float* loss_map;
} ps1_blend_ctx;
static void blend_tile_ps1(void* vctx, int tile) {
    const ps1_blend_ctx* c = (const ps1_blend_ctx*)vctx;
    const orc_camera* cam = c->cam; const int mode = c->mode, W = c->W, H = c->H, gx = c->gx;
    const uint32_t* rng = c->rng; const float* opacity = c->opacity; const float* rgb = c->rgb;
    int* gaussians_count = c->gaussians_count; float* contributions = c->contributions;
    float* final_T = c->final_T; uint32_t* n_contrib = c->n_contrib; float* out_color = c->out_color;
    struct { const inst_t* inst; const splat_t* sp; } o = {c->inst, c->sp};

        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = rng[2 * tile], r1 = rng[2 * tile + 1];
        const int total = (int)(r1 - r0);
        float Tt[256], C[256][3]; uint8_t done[256]; uint32_t contributor[256], last[256];
        /* LWMC: per pixel the Gaussian with the largest alpha*T so far (strict >, first wins), id 0 if none
           (pcheck_obb_loss_weighted_max_count/cuda_rasterizer/forward.cu:347-348,403-408) */
        int max_idx[256]; float max_contrib[256];
        const float* loss_map = c->loss_map;
        /* mode 4 = vanilla: no `power < -4.5f` cut (diff-gaussian-rasterization/cuda_rasterizer/forward.cu:342), no statistics */
        const float cut = mode == 4 ? -INFINITY : -4.5f;
        int ndone = 0;
        for (int t = 0; t < 256; t++) {
            const int px = tx * 16 + (t & 15), py = ty * 16 + (t >> 4);
            max_idx[t] = 0; max_contrib[t] = 0.0f;
            Tt[t] = 1.0f; C[t][0] = C[t][1] = C[t][2] = 0; contributor[t] = last[t] = 0;
            done[t] = !(px < W && py < H); ndone += done[t];
        }
        const int rounds = (total + 255) / 256;
        int toDo = total;
        for (int b = 0; b < rounds; b++, toDo -= 256) {
            if (ndone == 256) break;
            const int lim = toDo < 256 ? toDo : 256;
            if (mode == 1 || mode == 3)   /* SUM, LWMC: counted when the batch is staged (SUM/forward.cu:361) */
                for (int j = 0; j < lim; j++) {
                    const uint32_t id = o.inst[r0 + (size_t)b * 256 + j].id;
                    __atomic_fetch_add(&gaussians_count[id], 1, __ATOMIC_RELAXED);
                }
            for (int t = 0; t < 256; t++) {
                if (done[t]) continue;
                const float pxf = (float)(tx * 16 + (t & 15)), pyf = (float)(ty * 16 + (t >> 4));
                for (int j = 0; j < lim && !done[t]; j++) {
                    const uint32_t id = o.inst[r0 + (size_t)b * 256 + j].id;
                    const splat_t* s = &o.sp[id];
                    contributor[t]++;
                    const float dx = s->px - pxf, dy = s->py - pyf;
                    const float power = gauss_power(s->conx, s->cony, s->conz, dx, dy);
                    if (power > 0.0f || power < cut) continue;
                    /* MAX: one count per (pixel, Gaussian) that passes the falloff cut (pcheck_obb_max/forward.cu:381) */
                    if (mode == 2) __atomic_fetch_add(&gaussians_count[id], 1, __ATOMIC_RELAXED);
                    const float alpha = fminf(0.99f, opacity[id] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = Tt[t] * (1 - alpha);
                    if (test_T < 0.0001f) { done[t] = 1; ndone++; continue; }
                    if (mode >= 1) {
                        const float contrib = alpha * Tt[t];
                        if (mode == 1) atomic_add_f32(&contributions[id], contrib);
                        else if (mode == 2) atomic_max_f32(&contributions[id], contrib);   /* pcheck_obb_max/forward.cu:400 */
                        else if (mode == 3 && contrib > max_contrib[t]) { max_contrib[t] = contrib; max_idx[t] = (int)id; }
                        for (int ch = 0; ch < 3; ch++) C[t][ch] = fmaf(Tt[t], alpha * rgb[3 * id + ch], C[t][ch]);
                    } else {
                        const float w = alpha * Tt[t];
                        for (int ch = 0; ch < 3; ch++) C[t][ch] = fmaf(rgb[3 * id + ch], w, C[t][ch]);
                    }
                    Tt[t] = test_T;
                    last[t] = contributor[t];
                }
            }
        }
        for (int t = 0; t < 256; t++) {
            const int px = tx * 16 + (t & 15), py = ty * 16 + (t >> 4);
            if (!(px < W && py < H)) continue;
            const size_t pid = (size_t)W * py + px;
            /* LWMC: the pixel's loss goes to its max contributor — Gaussian 0 when nothing contributed (forward.cu:435) */
            if (mode == 3 && loss_map) atomic_add_f32(&contributions[max_idx[t]], loss_map[pid]);
            if (final_T) final_T[pid] = Tt[t];
            if (n_contrib) n_contrib[pid] = last[t];
            for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * H * W + pid] = fmaf(cam->bg[ch], Tt[t], C[t][ch]);
        }
}

/* =================================================================================================================
 * PS=1 forward (mode 0 = pcheck_obb, 1 = pcheck_obb_sum, 2 = pcheck_obb_max, 3 = pcheck_obb_loss_weighted_max_count;
 * modes 1-3 share everything but the per-Gaussian statistics of the blend; mode 4 = the vanilla diff-gaussian-rasterization
 * of fov3dgs/submodules/diff-gaussian-rasterization: SUM's arithmetic without the OBB test, the -4.5 falloff cut and the
 * statistics — gaussians_count / contributions are not touched).  All pointers are host memory.  Optional
 * outputs may be NULL; `loss_map` [H*W] is read by mode 3 only.
 * ================================================================================================================= */
int64_t orc_forward_ps1(const orc_camera* cam, int mode, int P, int M, const float* means3D, const float* opacity,
                        const float* scales, const float* rot, const float* shs, float* out_color, int* radii,
                        int* gaussians_count, float* contributions, float* means2D, float* depths, float* conic,
                        float* cov3D, float* rgb_out, uint8_t* clamped_out, uint32_t* point_list, int64_t list_cap,
                        uint32_t* ranges, float* final_T, uint32_t* n_contrib, const float* loss_map) {
    const int W = cam->W, H = cam->H, gx = (W + 15) / 16, gy = (H + 15) / 16, T = gx * gy;
    const float fy = H / (2.0f * cam->tanfovy), fx = W / (2.0f * cam->tanfovx);
    bin_in_t in; memset(&in, 0, sizeof(in));
    in.mode = mode >= 1 ? 1 : 0; in.no_obb = mode == 4; in.cam = cam; in.P = P; in.M = M; in.means3D = means3D; in.opacity = opacity; in.scales = scales; in.rot = rot; in.shs = shs;
    bin_out_t o; memset(&o, 0, sizeof(o));
    o.sp = (splat_t*)calloc((size_t)P + 1, sizeof(splat_t)); o.vis = (uint8_t*)calloc((size_t)P + 1, 1);
    o.radii = radii; o.cov3d = (float*)calloc((size_t)P * 6 + 6, sizeof(float));
    float* rgb = (float*)calloc((size_t)P * 3 + 3, sizeof(float));
    uint8_t* clamped = (uint8_t*)calloc((size_t)P * 3 + 3, 1);
    const int64_t n = run_binning(&in, &o, fx, fy, gx, gy);
    uint32_t* rng = (uint32_t*)calloc((size_t)T * 2, sizeof(uint32_t));
    fill_lists(&o, T, point_list, list_cap, rng);
    if (ranges) memcpy(ranges, rng, sizeof(uint32_t) * 2 * (size_t)T);
    /* colours: OBB evaluates survivors after culling, SUM every preprocessed Gaussian — same values where used */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (!o.vis[i]) continue;
        float d[3]; view_dir(cam, means3D + 3 * i, d);
        const float* sh = shs + (size_t)3 * M * i;
        float res[3] = {SH_C0 * sh[0], SH_C0 * sh[1], SH_C0 * sh[2]};
        sh_accumulate(sh, 1, cam->sh_degree, d[0], d[1], d[2], res);
        for (int ch = 0; ch < 3; ch++) {
            res[ch] += 0.5f;
            clamped[3 * i + ch] = res[ch] < 0;
            rgb[3 * i + ch] = fmaxf(res[ch], 0.0f);
        }
        if (means2D) { means2D[2 * i] = o.sp[i].px; means2D[2 * i + 1] = o.sp[i].py; }
        if (depths) depths[i] = o.sp[i].depth;
        if (conic) { conic[3 * i] = o.sp[i].conx; conic[3 * i + 1] = o.sp[i].cony; conic[3 * i + 2] = o.sp[i].conz; }
    }
    if (cov3D) memcpy(cov3D, o.cov3d, sizeof(float) * 6 * (size_t)P);
    if (rgb_out) memcpy(rgb_out, rgb, sizeof(float) * 3 * (size_t)P);
    if (clamped_out) memcpy(clamped_out, clamped, (size_t)3 * P);
    /* blend: one 16x16 tile per work item, 256-entry batches with the block-wide "all done" vote
       (OBB/forward.cu:251-384, SUM/forward.cu:298-430) */
    {
        ps1_blend_ctx bc = {cam, mode, W, H, gx, rng, o.inst, o.sp, opacity, rgb, gaussians_count, contributions, final_T, n_contrib, out_color, loss_map};
        run_tiles(blend_tile_ps1, &bc, T);
    }
    free(o.sp); free(o.vis); free(o.cov3d); free(o.inst); free(rgb); free(clamped); free(rng);
    return n;
}

typedef struct {
    const orc_camera* cam; int W, H, gx; const uint32_t* rng; const inst_t* inst; const splat_t* sp;
    const float* tm; const float* tgx; const float* tgy; const uint8_t* tb; const float* opacities4; const float* fc;
    const float* highest_levels; float* out_color;
    int naive;   /* SMFR baseline (naive_pcheck_obb/cuda_rasterizer/forward.cu:383-430): one opacity/colour for all levels */
} fov_blend_ctx;
static void blend_tile_fov(void* vctx, int tile) {
    const fov_blend_ctx* c = (const fov_blend_ctx*)vctx;
    const orc_camera* cam = c->cam; const int W = c->W, H = c->H, gx = c->gx;
    const uint32_t* rng = c->rng; const float* tm = c->tm; const float* tgx = c->tgx; const float* tgy = c->tgy; const uint8_t* tb = c->tb;
    const float* opacities4 = c->opacities4; const float* fc = c->fc; const float* highest_levels = c->highest_levels; float* out_color = c->out_color;
    struct { const inst_t* inst; const splat_t* sp; } o = {c->inst, c->sp};

        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = rng[2 * tile], r1 = rng[2 * tile + 1];
        const int total = (int)(r1 - r0), rounds = (total + 255) / 256;
        const float tlf = tm[tile];              /* Q2 */
        const int L1 = f2i_rz(tlf), L2 = L1 + 1;
        const float L2f = tlf + 1.0f;
        const int blending = tb[tile];
        float T1[256], T2[256], C1[256][3], C2[256][3], est[256]; uint8_t done[256], d1[256], d2[256];
        int ndone = 0;
        for (int t = 0; t < 256; t++) {
            const int px = tx * 16 + (t & 15), py = ty * 16 + (t >> 4);
            T1[t] = T2[t] = 1.0f; for (int ch = 0; ch < 3; ch++) C1[t][ch] = C2[t][ch] = 0;
            done[t] = !(px < W && py < H); ndone += done[t];
            est[t] = fmaf(fmaf((float)(t & 15), tgx[tile], (float)(t >> 4) * tgy[tile]), 0.0625f, tlf);
            d1[t] = blending ? (est[t] > (float)L2) : 0; d2[t] = 0;
        }
        int toDo = total;
        for (int b = 0; b < rounds; b++, toDo -= 256) {
            if (ndone == 256) break;
            const int lim = toDo < 256 ? toDo : 256;
            for (int t = 0; t < 256; t++) {
                if (done[t]) continue;
                const float pxf = (float)(tx * 16 + (t & 15)), pyf = (float)(ty * 16 + (t >> 4));
                for (int j = 0; j < lim && !done[t]; j++) {
                    const uint32_t id = o.inst[r0 + (size_t)b * 256 + j].id;
                    const splat_t* s = &o.sp[id];
                    const float dx = s->px - pxf, dy = s->py - pyf;
                    const float power = gauss_power(s->conx, s->cony, s->conz, dx, dy);
                    if (power > 0.0f || power < -4.5f) continue;
                    const float e = expf(power);
                    if (!blending) {
                        const float a = fminf(0.99f, opacities4[(size_t)id * 4 + L1] * e);
                        if (a < 1.0f / 255.0f) continue;
                        const float tt = T1[t] * (1 - a);
                        if (tt < 0.0001f) { done[t] = 1; ndone++; continue; }
                        const float w = a * T1[t];
                        for (int ch = 0; ch < 3; ch++) C1[t][ch] = fmaf(fc[(size_t)id * 12 + L1 * 3 + ch], w, C1[t][ch]);
                        T1[t] = tt;
                    } else if (c->naive) {
                        /* the shared-model baseline tests alpha once: a live L1 drops the entry for both levels when
                           alpha < 1/255, but once L1 is done the entry still reaches L2 whatever its alpha */
                        const float a = fminf(0.99f, opacities4[(size_t)id * 4 + L1] * e);
                        if (!d1[t]) {
                            if (a < 1.0f / 255.0f) continue;
                            const float tt = T1[t] * (1 - a);
                            d1[t] = tt < 0.0001f;
                            if (!d1[t]) {
                                const float w = a * T1[t];
                                for (int ch = 0; ch < 3; ch++) C1[t][ch] = fmaf(fc[(size_t)id * 12 + L1 * 3 + ch], w, C1[t][ch]);
                                T1[t] = tt;
                            }
                        }
                        if (!d2[t]) {
                            if (!((highest_levels[id] + 1) < L2f)) {
                                const float tt = T2[t] * (1 - a);
                                d2[t] = tt < 0.0001f;
                                if (!d2[t]) {
                                    const float w = a * T2[t];
                                    for (int ch = 0; ch < 3; ch++) C2[t][ch] = fmaf(fc[(size_t)id * 12 + L1 * 3 + ch], w, C2[t][ch]);
                                    T2[t] = tt;
                                }
                            }
                        }
                        if (d1[t] && d2[t]) { done[t] = 1; ndone++; }
                    } else {
                        if (!d1[t]) {
                            const float a = fminf(0.99f, opacities4[(size_t)id * 4 + L1] * e);
                            if (!(a < 1.0f / 255.0f)) {
                                const float tt = T1[t] * (1 - a);
                                d1[t] = tt < 0.0001f;
                                if (!d1[t]) {
                                    const float w = a * T1[t];
                                    for (int ch = 0; ch < 3; ch++) C1[t][ch] = fmaf(fc[(size_t)id * 12 + L1 * 3 + ch], w, C1[t][ch]);
                                    T1[t] = tt;
                                }
                            }
                        }
                        if (!d2[t]) {
                            const float a = fminf(0.99f, opacities4[(size_t)id * 4 + L2] * e);
                            const int skip = (a < 1.0f / 255.0f) || ((highest_levels[id] + 1) < L2f);
                            if (!skip) {
                                const float tt = T2[t] * (1 - a);
                                d2[t] = tt < 0.0001f;
                                if (!d2[t]) {
                                    const float w = a * T2[t];
                                    for (int ch = 0; ch < 3; ch++) C2[t][ch] = fmaf(fc[(size_t)id * 12 + L2 * 3 + ch], w, C2[t][ch]);
                                    T2[t] = tt;
                                }
                            }
                        }
                        if (d1[t] && d2[t]) { done[t] = 1; ndone++; }
                    }
                }
            }
        }
        for (int t = 0; t < 256; t++) {
            const int px = tx * 16 + (t & 15), py = ty * 16 + (t >> 4);
            if (!(px < W && py < H)) continue;
            const size_t pid = (size_t)W * py + px;
            if (!blending) {
                for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * H * W + pid] = fmaf(cam->bg[ch], T1[t], C1[t][ch]);
            } else {
                const float v = est[t] - ((float)L1 + start_blend);
                float x = fabsf(v) + fabsf(v);
                x = fmaxf(0.0f, fminf(1.0f, x));
                const float nb = fmaf(x, x * (x + x), x * (x * -3.0f));
                const float w1 = nb + 1.0f, w2 = 1.0f - w1;
                for (int ch = 0; ch < 3; ch++) {
                    const float a = fmaf(cam->bg[ch], T1[t], C1[t][ch]), bb = fmaf(cam->bg[ch], T2[t], C2[t][ch]);
                    out_color[(size_t)ch * H * W + pid] = fmaf(a, w1, bb * w2);
                }
            }
        }
}

/* =================================================================================================================
 * Foveated forward (diff_gaussian_rasterization_fov_pcheck_obb).
 * ================================================================================================================= */
static int64_t forward_fov_impl(int naive, const orc_camera* cam, int P, int M_rest, const float* means3D, const float* opacities4_in,
                        const float* scales, const float* rot, const float* shs_rest, const float* shs_dcs,
                        const float* highest_levels, const float* gaze, float alpha_pool, float* out_color, int* radii,
                        float* means2D, float* depths, float* conic, uint32_t* point_list, int64_t list_cap,
                        uint32_t* ranges, float* tile_level_out, float* tile_min_out, uint8_t* tile_blend_out,
                        int32_t* level_ranges_out) {
    /* naive (SMFR): `shs_rest` is the full [P,M,3] SH tensor (DC first), `opacities4_in` is [P]; shs_dcs unused */
    const float* opacities4 = opacities4_in;
    float* op_rep = NULL;
    if (naive) {
        op_rep = (float*)malloc(sizeof(float) * 4 * ((size_t)P + 1));
        for (size_t i = 0; i < (size_t)P; i++) for (int l = 0; l < 4; l++) op_rep[4 * i + l] = opacities4_in[i];
        opacities4 = op_rep;
    }
    const int W = cam->W, H = cam->H, gx = (W + 15) / 16, gy = (H + 15) / 16, T = gx * gy;
    const float fy = H / (2.0f * cam->tanfovy), fx = W / (2.0f * cam->tanfovx);
    float* tl = (float*)malloc(sizeof(float) * T); float* tm = (float*)malloc(sizeof(float) * T);
    float* tgx = (float*)malloc(sizeof(float) * T); float* tgy = (float*)malloc(sizeof(float) * T);
    uint8_t* tb = (uint8_t*)malloc(T);
    orc_tile_tables(W, H, gaze, alpha_pool, tl, tm, tgx, tgy, tb);
    if (tile_level_out) memcpy(tile_level_out, tl, sizeof(float) * T);
    if (tile_min_out) memcpy(tile_min_out, tm, sizeof(float) * T);
    if (tile_blend_out) memcpy(tile_blend_out, tb, T);
    bin_in_t in; memset(&in, 0, sizeof(in));
    in.mode = 2; in.cam = cam; in.P = P; in.M = M_rest; in.means3D = means3D; in.scales = scales; in.rot = rot;
    in.highest_levels = highest_levels; in.tile_min = tm; in.tile_blend = tb;   /* Q2: filter receives tile_level_min */
    bin_out_t o; memset(&o, 0, sizeof(o));
    o.sp = (splat_t*)calloc((size_t)P + 1, sizeof(splat_t)); o.vis = (uint8_t*)calloc((size_t)P + 1, 1);
    o.radii = radii; o.cov3d = (float*)calloc((size_t)P * 6 + 6, sizeof(float));
    o.lvl_lo = (int32_t*)calloc((size_t)P + 1, 4); o.lvl_hi = (int32_t*)calloc((size_t)P + 1, 4);
    const int64_t n = run_binning(&in, &o, fx, fy, gx, gy);
    uint32_t* rng = (uint32_t*)calloc((size_t)T * 2, sizeof(uint32_t));
    fill_lists(&o, T, point_list, list_cap, rng);
    if (ranges) memcpy(ranges, rng, sizeof(uint32_t) * 2 * (size_t)T);
    /* compute_fov_colors (rasterizer_impl.cu:490-530): only levels in level_ranges are defined */
    float* fc = (float*)calloc((size_t)P * 12 + 12, sizeof(float));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        float d[3]; view_dir(cam, means3D + 3 * i, d);
        float res[3] = {0, 0, 0};
        if (naive) {
            /* computeColorFromSH of the shared model (naive_pcheck_obb/cuda_rasterizer/rasterizer_impl.cu:34-82) */
            const float* sh = shs_rest + (size_t)3 * M_rest * i;
            for (int ch = 0; ch < 3; ch++) res[ch] = SH_C0 * sh[ch];
            sh_accumulate(sh, 1, cam->sh_degree, d[0], d[1], d[2], res);
            for (int l = 0; l < 4; l++)
                for (int ch = 0; ch < 3; ch++) fc[(size_t)i * 12 + l * 3 + ch] = fmaxf(res[ch] + 0.5f, 0.0f);
        } else {
        if (M_rest > 0) sh_accumulate(shs_rest + (size_t)3 * M_rest * i, 0, cam->sh_degree, d[0], d[1], d[2], res);
        for (int ch = 0; ch < 3; ch++) res[ch] += 0.5f;
        for (int l = o.lvl_lo[i]; l <= o.lvl_hi[i]; l++)
            for (int ch = 0; ch < 3; ch++)
                fc[(size_t)i * 12 + l * 3 + ch] = fmaxf(SH_C0 * shs_dcs[(size_t)i * 12 + l * 3 + ch] + res[ch], 0.0f);
        }
        if (means2D) { means2D[2 * i] = o.sp[i].px; means2D[2 * i + 1] = o.sp[i].py; }
        if (depths) depths[i] = o.sp[i].depth;
        if (conic) { conic[3 * i] = o.sp[i].conx; conic[3 * i + 1] = o.sp[i].cony; conic[3 * i + 2] = o.sp[i].conz; }
        if (level_ranges_out) { level_ranges_out[2 * i] = o.lvl_lo[i]; level_ranges_out[2 * i + 1] = o.lvl_hi[i]; }
    }
    {
        fov_blend_ctx bc = {cam, W, H, gx, rng, o.inst, o.sp, tm, tgx, tgy, tb, opacities4, fc, highest_levels, out_color, naive};
        run_tiles(blend_tile_fov, &bc, T);
    }
    free(tl); free(tm); free(tgx); free(tgy); free(tb); free(o.sp); free(o.vis); free(o.cov3d); free(o.lvl_lo); free(o.lvl_hi);
    free(o.inst); free(rng); free(fc); free(op_rep);
    return n;
}

int64_t orc_forward_fov(const orc_camera* cam, int P, int M_rest, const float* means3D, const float* opacities4,
                        const float* scales, const float* rot, const float* shs_rest, const float* shs_dcs,
                        const float* highest_levels, const float* gaze, float alpha_pool, float* out_color, int* radii,
                        float* means2D, float* depths, float* conic, uint32_t* point_list, int64_t list_cap,
                        uint32_t* ranges, float* tile_level_out, float* tile_min_out, uint8_t* tile_blend_out,
                        int32_t* level_ranges_out) {
    return forward_fov_impl(0, cam, P, M_rest, means3D, opacities4, scales, rot, shs_rest, shs_dcs, highest_levels, gaze, alpha_pool,
                            out_color, radii, means2D, depths, conic, point_list, list_cap, ranges, tile_level_out, tile_min_out,
                            tile_blend_out, level_ranges_out);
}

/* =================================================================================================================
 * SMFR baseline (diff_gaussian_rasterization_naive_pcheck_obb): the foveated pipeline with ONE shared model — `shs` is
 * the full [P,M,3] tensor, `opacity` [P]; levels only subset the Gaussians (highest_levels) and pick the blending path.
 * ================================================================================================================= */
int64_t orc_forward_smfr(const orc_camera* cam, int P, int M, const float* means3D, const float* opacity,
                         const float* scales, const float* rot, const float* shs, const float* highest_levels,
                         const float* gaze, float alpha_pool, float* out_color, int* radii, uint32_t* point_list,
                         int64_t list_cap, uint32_t* ranges) {
    return forward_fov_impl(1, cam, P, M, means3D, opacity, scales, rot, shs, NULL, highest_levels, gaze, alpha_pool, out_color,
                            radii, NULL, NULL, NULL, point_list, list_cap, ranges, NULL, NULL, NULL, NULL);
}

/* =================================================================================================================
 * MMFR baseline (diff_gaussian_rasterization_mmfr_pcheck_obb): one call per level model.  `cur_level` selects the tiles
 * (compute_tile_skips_cuda, rasterizer_impl.cu:277-304); plain tiles composite normally, blending tiles composite the one
 * model and weight it (forward.cu:255-418), skipped tiles stay 0; the caller sums the four level images.
 * The reference refreshes its (process-static) tile tables only when cur_level == 0; this restatement computes them
 * on every call, which is the same thing for the four calls of one frame (same gaze, alpha, size).
 * ================================================================================================================= */
typedef struct {
    const orc_camera* cam; int W, H, gx; const uint32_t* rng; const inst_t* inst; const splat_t* sp;
    const float* tm; const float* tgx; const float* tgy; const uint8_t* tb; const uint8_t* tskip; const float* opacity;
    const float* rgb; float cur_level; float* out_color;
} mmfr_blend_ctx;
static void blend_tile_mmfr(void* vctx, int tile) {
    const mmfr_blend_ctx* c = (const mmfr_blend_ctx*)vctx;
    const int W = c->W, H = c->H, gx = c->gx;
    const int tx = tile % gx, ty = tile / gx;
    if (c->tskip[tile]) return;                       /* out_color keeps its initial 0 */
    const int blending = c->tb[tile];
    const uint32_t r0 = c->rng[2 * tile], r1 = c->rng[2 * tile + 1];
    const int total = (int)(r1 - r0), rounds = (total + 255) / 256;
    const float tlf = c->tm[tile];
    float T1[256], C1[256][3], xs[256]; int L1s[256]; uint8_t done[256];
    int ndone = 0;
    for (int t = 0; t < 256; t++) {
        const int px = tx * 16 + (t & 15), py = ty * 16 + (t >> 4);
        T1[t] = 1.0f; C1[t][0] = C1[t][1] = C1[t][2] = 0;
        done[t] = !(px < W && py < H);
        if (blending) {
            const float est = fmaf(fmaf((float)(t & 15), c->tgx[tile], (float)(t >> 4) * c->tgy[tile]), 0.0625f, tlf);
            L1s[t] = f2i_rz(est);
            xs[t] = (est - ((float)L1s[t] + start_blend)) / blend_width;
            if (xs[t] < 0 && (float)L1s[t] != c->cur_level) done[t] = 1;
        }
        ndone += done[t];
    }
    int toDo = total;
    for (int b = 0; b < rounds; b++, toDo -= 256) {
        if (ndone == 256) break;
        const int lim = toDo < 256 ? toDo : 256;
        for (int t = 0; t < 256; t++) {
            if (done[t]) continue;
            const float pxf = (float)(tx * 16 + (t & 15)), pyf = (float)(ty * 16 + (t >> 4));
            for (int j = 0; j < lim && !done[t]; j++) {
                const uint32_t id = c->inst[r0 + (size_t)b * 256 + j].id;
                const splat_t* s = &c->sp[id];
                const float dx = s->px - pxf, dy = s->py - pyf;
                const float power = gauss_power(s->conx, s->cony, s->conz, dx, dy);
                if (power > 0.0f || power < -4.5f) continue;
                const float a = fminf(0.99f, c->opacity[id] * expf(power));
                if (a < 1.0f / 255.0f) continue;
                const float tt = T1[t] * (1 - a);
                if (tt < 0.0001f) { done[t] = 1; ndone++; continue; }
                const float w = a * T1[t];
                for (int ch = 0; ch < 3; ch++) C1[t][ch] = fmaf(c->rgb[3 * id + ch], w, C1[t][ch]);
                T1[t] = tt;
            }
        }
    }
    for (int t = 0; t < 256; t++) {
        const int px = tx * 16 + (t & 15), py = ty * 16 + (t >> 4);
        if (!(px < W && py < H)) continue;
        const size_t pid = (size_t)W * py + px;
        for (int ch = 0; ch < 3; ch++) {
            const float v = fmaf(c->cam->bg[ch], T1[t], C1[t][ch]);
            if (!blending) { c->out_color[(size_t)ch * H * W + pid] = v; continue; }
            const float x = fmaxf(0.0f, fminf(1.0f, xs[t]));
            const float nb = fmaf(x, x * (x + x), x * (x * -3.0f));     /* -(3x^2 - 2x^3), as the FOV kernel's binary */
            const float w1 = nb + 1.0f;
            const float used = ((float)L1s[t] == c->cur_level) ? w1 : 1.0f - w1;
            c->out_color[(size_t)ch * H * W + pid] = v * used;
        }
    }
}

int64_t orc_forward_mmfr(const orc_camera* cam, int P, int M, const float* means3D, const float* opacity,
                         const float* scales, const float* rot, const float* shs, float cur_level, const float* gaze,
                         float alpha_pool, float* out_color, int* radii, uint32_t* point_list, int64_t list_cap,
                         uint32_t* ranges, uint8_t* tile_skip_out) {
    const int W = cam->W, H = cam->H, gx = (W + 15) / 16, gy = (H + 15) / 16, T = gx * gy;
    const float fy = H / (2.0f * cam->tanfovy), fx = W / (2.0f * cam->tanfovx);
    float* tl = (float*)malloc(sizeof(float) * T); float* tm = (float*)malloc(sizeof(float) * T);
    float* tgx = (float*)malloc(sizeof(float) * T); float* tgy = (float*)malloc(sizeof(float) * T);
    uint8_t* tb = (uint8_t*)malloc(T); uint8_t* tskip = (uint8_t*)malloc(T);
    tile_tables_impl(W, H, gaze, alpha_pool, tl, tm, tgx, tgy, tb, 1);
    {
        const float lb = cur_level - blend_width, hb = cur_level + 1;
        for (int i = 0; i < T; i++) tskip[i] = !(tm[i] > lb && tm[i] < hb);
    }
    if (tile_skip_out) memcpy(tile_skip_out, tskip, T);
    bin_in_t in; memset(&in, 0, sizeof(in));
    in.mode = 3; in.cam = cam; in.P = P; in.M = M; in.means3D = means3D; in.scales = scales; in.rot = rot; in.tile_skip = tskip;
    bin_out_t o; memset(&o, 0, sizeof(o));
    o.sp = (splat_t*)calloc((size_t)P + 1, sizeof(splat_t)); o.vis = (uint8_t*)calloc((size_t)P + 1, 1);
    o.radii = radii; o.cov3d = (float*)calloc((size_t)P * 6 + 6, sizeof(float));
    const int64_t n = run_binning(&in, &o, fx, fy, gx, gy);
    uint32_t* rng = (uint32_t*)calloc((size_t)T * 2, sizeof(uint32_t));
    fill_lists(&o, T, point_list, list_cap, rng);
    if (ranges) memcpy(ranges, rng, sizeof(uint32_t) * 2 * (size_t)T);
    float* rgb = (float*)calloc((size_t)P * 3 + 3, sizeof(float));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        float d[3]; view_dir(cam, means3D + 3 * i, d);
        const float* sh = shs + (size_t)3 * M * i;
        float res[3] = {SH_C0 * sh[0], SH_C0 * sh[1], SH_C0 * sh[2]};
        sh_accumulate(sh, 1, cam->sh_degree, d[0], d[1], d[2], res);
        for (int ch = 0; ch < 3; ch++) rgb[3 * i + ch] = fmaxf(res[ch] + 0.5f, 0.0f);
    }
    memset(out_color, 0, sizeof(float) * 3 * (size_t)W * H);
    {
        mmfr_blend_ctx bc = {cam, W, H, gx, rng, o.inst, o.sp, tm, tgx, tgy, tb, tskip, opacity, rgb, cur_level, out_color};
        run_tiles(blend_tile_mmfr, &bc, T);
    }
    free(tl); free(tm); free(tgx); free(tgy); free(tb); free(tskip); free(o.sp); free(o.vis); free(o.cov3d); free(o.inst);
    free(rng); free(rgb);
    return n;
}

/* =================================================================================================================
 * Backward of the SUM variant (SUM/cuda_rasterizer/backward.cu).  Inputs: forward intermediates.
 * Gradients are accumulated in double per Gaussian (the reference accumulates with fp32 atomics in arbitrary order).
 * ================================================================================================================= */
static void dnormvdv3(const float* v, const float* dv, float* out) {
    const float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
    out[0] = ((+sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * inv;
    out[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * inv;
    out[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * inv;
}

static int backward_ps1_impl(float cut, const orc_camera* cam, int P, int M, const float* means3D, const float* scales, const float* rot,
                     const float* shs, const float* opacity, const int* radii, const float* means2D, const float* conic,
                     const float* rgb, const uint8_t* clamped, const float* cov3D, const uint32_t* point_list,
                     const uint32_t* ranges, const float* final_T, const uint32_t* n_contrib, const float* dL_dpix,
                     float* dL_dmeans2D, float* dL_dconic4, float* dL_dopacity, float* dL_dcolors, float* dL_dmeans3D,
                     float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drot) {
    const int W = cam->W, H = cam->H, gx = (W + 15) / 16, gy = (H + 15) / 16, T = gx * gy;
    const float fy = H / (2.0f * cam->tanfovy), fx = W / (2.0f * cam->tanfovx);
    double* acc = (double*)calloc((size_t)P * 9 + 9, sizeof(double));
    const float ddelx_dx = (float)(0.5 * W), ddely_dy = (float)(0.5 * H);
    /* renderCUDA backward (backward.cu:399-557) */
    for (int tile = 0; tile < T; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        const int total = (int)(r1 - r0);
        for (int t = 0; t < 256; t++) {
            const int px = tx * 16 + (t & 15), py = ty * 16 + (t >> 4);
            if (!(px < W && py < H)) continue;
            const size_t pid = (size_t)W * py + px;
            const float pxf = (float)px, pyf = (float)py;
            const float Tf = final_T[pid];
            float Tc = Tf;
            uint32_t contributor = (uint32_t)total;
            const int last_contributor = (int)n_contrib[pid];
            float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0;
            float dpix[3]; for (int ch = 0; ch < 3; ch++) dpix[ch] = dL_dpix[(size_t)ch * H * W + pid];
            for (int k = 0; k < total; k++) {
                const uint32_t id = point_list[r1 - k - 1];
                contributor--;
                if ((int)contributor >= last_contributor) continue;
                const float dx = means2D[2 * id] - pxf, dy = means2D[2 * id + 1] - pyf;
                const float cx = conic[3 * id], cy = conic[3 * id + 1], cz = conic[3 * id + 2], op = opacity[id];
                const float power = gauss_power(cx, cy, cz, dx, dy);
                if (power > 0.0f || power < cut) continue;
                const float G = expf(power);
                const float alpha = fminf(0.99f, op * G);
                if (alpha < 1.0f / 255.0f) continue;
                Tc = Tc / (1.f - alpha);
                const float dchannel_dcolor = alpha * Tc;
                float dL_dalpha = 0.0f;
                for (int ch = 0; ch < 3; ch++) {
                    const float c = rgb[3 * id + ch];
                    accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                    last_color[ch] = c;
                    dL_dalpha += (c - accum_rec[ch]) * dpix[ch];
                    acc[(size_t)id * 9 + ch] += dchannel_dcolor * dpix[ch];
                }
                dL_dalpha *= Tc;
                last_alpha = alpha;
                float bg_dot = 0; for (int ch = 0; ch < 3; ch++) bg_dot += cam->bg[ch] * dpix[ch];
                dL_dalpha += (-Tf / (1.f - alpha)) * bg_dot;
                const float dL_dG = op * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * cx - gdy * cy, dG_ddely = -gdy * cz - gdx * cy;
                acc[(size_t)id * 9 + 3] += dL_dG * dG_ddelx * ddelx_dx;
                acc[(size_t)id * 9 + 4] += dL_dG * dG_ddely * ddely_dy;
                acc[(size_t)id * 9 + 5] += -0.5f * gdx * dx * dL_dG;
                acc[(size_t)id * 9 + 6] += -0.5f * gdx * dy * dL_dG;
                acc[(size_t)id * 9 + 7] += -0.5f * gdy * dy * dL_dG;
                acc[(size_t)id * 9 + 8] += G * dL_dalpha;
            }
        }
    }
    for (int i = 0; i < P; i++) {
        for (int ch = 0; ch < 3; ch++) dL_dcolors[3 * i + ch] = (float)acc[(size_t)i * 9 + ch];
        dL_dmeans2D[3 * i] = (float)acc[(size_t)i * 9 + 3]; dL_dmeans2D[3 * i + 1] = (float)acc[(size_t)i * 9 + 4]; dL_dmeans2D[3 * i + 2] = 0;
        dL_dconic4[4 * i] = (float)acc[(size_t)i * 9 + 5]; dL_dconic4[4 * i + 1] = (float)acc[(size_t)i * 9 + 6];
        dL_dconic4[4 * i + 2] = 0; dL_dconic4[4 * i + 3] = (float)acc[(size_t)i * 9 + 7];
        dL_dopacity[i] = (float)acc[(size_t)i * 9 + 8];
    }
    free(acc);
    const float* v = cam->view; const float* proj = cam->proj;
    /* computeCov2DCUDA (backward.cu:144-274) + preprocessCUDA (backward.cu:346-396) */
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        if (!(radii[idx] > 0)) continue;
        const float* mean = means3D + 3 * idx;
        const float* c3 = cov3D + 6 * (size_t)idx;
        const float dcx = dL_dconic4[4 * idx], dcy = dL_dconic4[4 * idx + 1], dcz = dL_dconic4[4 * idx + 3];
        float t[3];
        t[0] = v[0] * mean[0] + v[4] * mean[1] + v[8] * mean[2] + v[12];
        t[1] = v[1] * mean[0] + v[5] * mean[1] + v[9] * mean[2] + v[13];
        t[2] = v[2] * mean[0] + v[6] * mean[1] + v[10] * mean[2] + v[14];
        const float limx = 1.3f * cam->tanfovx, limy = 1.3f * cam->tanfovy;
        const float txtz = t[0] / t[2], tytz = t[1] / t[2];
        t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
        t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
        const float xgm = (txtz < -limx || txtz > limx) ? 0.f : 1.f, ygm = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float J00 = fx / t[2], J02 = -(fx * t[0]) / (t[2] * t[2]), J11 = fy / t[2], J12 = -(fy * t[1]) / (t[2] * t[2]);
        const float T00 = v[0] * J00 + v[2] * J02, T01 = v[4] * J00 + v[6] * J02, T02 = v[8] * J00 + v[10] * J02;
        const float T10 = v[1] * J11 + v[2] * J12, T11 = v[5] * J11 + v[6] * J12, T12 = v[9] * J11 + v[10] * J12;
        const float V00 = c3[0], V01 = c3[1], V02 = c3[2], V11 = c3[3], V12 = c3[4], V22 = c3[5];
        const float A0 = T00 * V00 + T01 * V01 + T02 * V02, A1 = T00 * V01 + T01 * V11 + T02 * V12, A2 = T00 * V02 + T01 * V12 + T02 * V22;
        const float B0 = T10 * V00 + T11 * V01 + T12 * V02, B1 = T10 * V01 + T11 * V11 + T12 * V12, B2 = T10 * V02 + T11 * V12 + T12 * V22;
        const float a = (A0 * T00 + A1 * T01 + A2 * T02) + 0.3f, b = (A0 * T10 + A1 * T11 + A2 * T12), c = (B0 * T10 + B1 * T11 + B2 * T12) + 0.3f;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float dcov[6] = {0, 0, 0, 0, 0, 0};
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            dcov[0] = (T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc);
            dcov[3] = (T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc);
            dcov[5] = (T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc);
            dcov[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
            dcov[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
            dcov[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
        }
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * (size_t)idx + k] = dcov[k];
        const float dT00 = 2 * A0 * dL_da + B0 * dL_db, dT01 = 2 * A1 * dL_da + B1 * dL_db, dT02 = 2 * A2 * dL_da + B2 * dL_db;
        const float dT10 = 2 * B0 * dL_dc + A0 * dL_db, dT11 = 2 * B1 * dL_dc + A1 * dL_db, dT12 = 2 * B2 * dL_dc + A2 * dL_db;
        const float dJ00 = v[0] * dT00 + v[4] * dT01 + v[8] * dT02, dJ02 = v[2] * dT00 + v[6] * dT01 + v[10] * dT02;
        const float dJ11 = v[1] * dT10 + v[5] * dT11 + v[9] * dT12, dJ12 = v[2] * dT10 + v[6] * dT11 + v[10] * dT12;
        const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = xgm * -fx * tz2 * dJ02, dty = ygm * -fy * tz2 * dJ12;
        const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[0]) * tz3 * dJ02 + (2 * fy * t[1]) * tz3 * dJ12;
        float dmean[3] = {v[0] * dtx + v[1] * dty + v[2] * dtz, v[4] * dtx + v[5] * dty + v[6] * dtz, v[8] * dtx + v[9] * dty + v[10] * dtz};
        {
            const float m_hw = proj[3] * mean[0] + proj[7] * mean[1] + proj[11] * mean[2] + proj[15];
            const float m_w = 1.0f / (m_hw + 0.0000001f);
            const float mul1 = (proj[0] * mean[0] + proj[4] * mean[1] + proj[8] * mean[2] + proj[12]) * m_w * m_w;
            const float mul2 = (proj[1] * mean[0] + proj[5] * mean[1] + proj[9] * mean[2] + proj[13]) * m_w * m_w;
            const float g2x = dL_dmeans2D[3 * idx], g2y = dL_dmeans2D[3 * idx + 1];
            dmean[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
            dmean[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
            dmean[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
        }
        if (shs) {   /* computeColorFromSH backward (backward.cu:20-139) */
            const int deg = cam->sh_degree;
            float dor[3] = {mean[0] - cam->campos[0], mean[1] - cam->campos[1], mean[2] - cam->campos[2]};
            const float len = sqrtf(dor[0] * dor[0] + dor[1] * dor[1] + dor[2] * dor[2]);
            const float x = dor[0] / len, y = dor[1] / len, z = dor[2] / len;
            const float* sh = shs + (size_t)3 * M * idx; float* dsh = dL_dsh + (size_t)3 * M * idx;
            float dRGB[3]; for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolors[3 * idx + ch] * (clamped[3 * idx + ch] ? 0.f : 1.f);
            float dx_[3] = {0, 0, 0}, dy_[3] = {0, 0, 0}, dz_[3] = {0, 0, 0};
#define S(k, ch) sh[3 * (k) + (ch)]
#define W3(k, w) do { const float w_ = (w); for (int ch = 0; ch < 3; ch++) dsh[3 * (k) + ch] = w_ * dRGB[ch]; } while (0)
            W3(0, SH_C0);
            if (deg > 0) {
                W3(1, -SH_C1 * y); W3(2, SH_C1 * z); W3(3, -SH_C1 * x);
                for (int ch = 0; ch < 3; ch++) { dx_[ch] = -SH_C1 * S(3, ch); dy_[ch] = -SH_C1 * S(1, ch); dz_[ch] = SH_C1 * S(2, ch); }
                if (deg > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    W3(4, SH_C2[0] * xy); W3(5, SH_C2[1] * yz); W3(6, SH_C2[2] * (2.f * zz - xx - yy)); W3(7, SH_C2[3] * xz); W3(8, SH_C2[4] * (xx - yy));
                    for (int ch = 0; ch < 3; ch++) {
                        dx_[ch] += SH_C2[0] * y * S(4, ch) + SH_C2[2] * 2.f * -x * S(6, ch) + SH_C2[3] * z * S(7, ch) + SH_C2[4] * 2.f * x * S(8, ch);
                        dy_[ch] += SH_C2[0] * x * S(4, ch) + SH_C2[1] * z * S(5, ch) + SH_C2[2] * 2.f * -y * S(6, ch) + SH_C2[4] * 2.f * -y * S(8, ch);
                        dz_[ch] += SH_C2[1] * y * S(5, ch) + SH_C2[2] * 2.f * 2.f * z * S(6, ch) + SH_C2[3] * x * S(7, ch);
                    }
                    if (deg > 2) {
                        W3(9, SH_C3[0] * y * (3.f * xx - yy)); W3(10, SH_C3[1] * xy * z); W3(11, SH_C3[2] * y * (4.f * zz - xx - yy));
                        W3(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)); W3(13, SH_C3[4] * x * (4.f * zz - xx - yy));
                        W3(14, SH_C3[5] * z * (xx - yy)); W3(15, SH_C3[6] * x * (xx - 3.f * yy));
                        for (int ch = 0; ch < 3; ch++) {
                            dx_[ch] += (SH_C3[0] * S(9, ch) * 3.f * 2.f * xy + SH_C3[1] * S(10, ch) * yz + SH_C3[2] * S(11, ch) * -2.f * xy +
                                        SH_C3[3] * S(12, ch) * -3.f * 2.f * xz + SH_C3[4] * S(13, ch) * (-3.f * xx + 4.f * zz - yy) +
                                        SH_C3[5] * S(14, ch) * 2.f * xz + SH_C3[6] * S(15, ch) * 3.f * (xx - yy));
                            dy_[ch] += (SH_C3[0] * S(9, ch) * 3.f * (xx - yy) + SH_C3[1] * S(10, ch) * xz + SH_C3[2] * S(11, ch) * (-3.f * yy + 4.f * zz - xx) +
                                        SH_C3[3] * S(12, ch) * -3.f * 2.f * yz + SH_C3[4] * S(13, ch) * -2.f * xy + SH_C3[5] * S(14, ch) * -2.f * yz +
                                        SH_C3[6] * S(15, ch) * -3.f * 2.f * xy);
                            dz_[ch] += (SH_C3[1] * S(10, ch) * xy + SH_C3[2] * S(11, ch) * 4.f * 2.f * yz + SH_C3[3] * S(12, ch) * 3.f * (2.f * zz - xx - yy) +
                                        SH_C3[4] * S(13, ch) * 4.f * 2.f * xz + SH_C3[5] * S(14, ch) * (xx - yy));
                        }
                    }
                }
            }
#undef S
#undef W3
            const float ddir[3] = {dx_[0] * dRGB[0] + dx_[1] * dRGB[1] + dx_[2] * dRGB[2], dy_[0] * dRGB[0] + dy_[1] * dRGB[1] + dy_[2] * dRGB[2],
                                   dz_[0] * dRGB[0] + dz_[1] * dRGB[1] + dz_[2] * dRGB[2]};
            float dm[3]; dnormvdv3(dor, ddir, dm);
            dmean[0] += dm[0]; dmean[1] += dm[1]; dmean[2] += dm[2];
        }
        for (int k = 0; k < 3; k++) dL_dmeans3D[3 * idx + k] = dmean[k];
        if (scales) {   /* computeCov3D backward (backward.cu:278-343) */
            const float mod = cam->scale_modifier;
            const float r = rot[4 * idx], x = rot[4 * idx + 1], y = rot[4 * idx + 2], z = rot[4 * idx + 3];
            float Rm[3][3];
            Rm[0][0] = 1.f - 2.f * (y * y + z * z); Rm[1][0] = 2.f * (x * y - r * z); Rm[2][0] = 2.f * (x * z + r * y);
            Rm[0][1] = 2.f * (x * y + r * z); Rm[1][1] = 1.f - 2.f * (x * x + z * z); Rm[2][1] = 2.f * (y * z - r * x);
            Rm[0][2] = 2.f * (x * z - r * y); Rm[1][2] = 2.f * (y * z + r * x); Rm[2][2] = 1.f - 2.f * (x * x + y * y);
            const float s[3] = {mod * scales[3 * idx], mod * scales[3 * idx + 1], mod * scales[3 * idx + 2]};
            float Mm[3][3], D[3][3], dM[3][3], E[3][3];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Mm[i][j] = s[i] * Rm[i][j];
            D[0][0] = dcov[0]; D[0][1] = 0.5f * dcov[1]; D[0][2] = 0.5f * dcov[2];
            D[1][0] = 0.5f * dcov[1]; D[1][1] = dcov[3]; D[1][2] = 0.5f * dcov[4];
            D[2][0] = 0.5f * dcov[2]; D[2][1] = 0.5f * dcov[4]; D[2][2] = dcov[5];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
                dM[i][j] = 2.0f * Mm[i][0] * D[0][j] + 2.0f * Mm[i][1] * D[1][j] + 2.0f * Mm[i][2] * D[2][j];
            for (int i = 0; i < 3; i++) dL_dscales[3 * idx + i] = Rm[i][0] * dM[i][0] + Rm[i][1] * dM[i][1] + Rm[i][2] * dM[i][2];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) E[i][j] = dM[i][j] * s[i];
            dL_drot[4 * idx + 0] = 2 * z * (E[0][1] - E[1][0]) + 2 * y * (E[2][0] - E[0][2]) + 2 * x * (E[1][2] - E[2][1]);
            dL_drot[4 * idx + 1] = 2 * y * (E[1][0] + E[0][1]) + 2 * z * (E[2][0] + E[0][2]) + 2 * r * (E[1][2] - E[2][1]) - 4 * x * (E[2][2] + E[1][1]);
            dL_drot[4 * idx + 2] = 2 * x * (E[1][0] + E[0][1]) + 2 * r * (E[2][0] - E[0][2]) + 2 * z * (E[1][2] + E[2][1]) - 4 * y * (E[2][2] + E[0][0]);
            dL_drot[4 * idx + 3] = 2 * r * (E[0][1] - E[1][0]) + 2 * x * (E[2][0] + E[0][2]) + 2 * y * (E[1][2] + E[2][1]) - 4 * z * (E[1][1] + E[0][0]);
        }
    }
    return 0;
}

/* SUM/cuda_rasterizer/backward.cu (the -4.5 falloff cut of the pcheck variants, :495) */
int orc_backward_ps1(const orc_camera* cam, int P, int M, const float* means3D, const float* scales, const float* rot,
                     const float* shs, const float* opacity, const int* radii, const float* means2D, const float* conic,
                     const float* rgb, const uint8_t* clamped, const float* cov3D, const uint32_t* point_list,
                     const uint32_t* ranges, const float* final_T, const uint32_t* n_contrib, const float* dL_dpix,
                     float* dL_dmeans2D, float* dL_dconic4, float* dL_dopacity, float* dL_dcolors, float* dL_dmeans3D,
                     float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drot) {
    return backward_ps1_impl(-4.5f, cam, P, M, means3D, scales, rot, shs, opacity, radii, means2D, conic, rgb, clamped, cov3D,
                             point_list, ranges, final_T, n_contrib, dL_dpix, dL_dmeans2D, dL_dconic4, dL_dopacity, dL_dcolors,
                             dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drot);
}
/* vanilla diff-gaussian-rasterization/cuda_rasterizer/backward.cu:495: only `power > 0` is skipped */
int orc_backward_vanilla(const orc_camera* cam, int P, int M, const float* means3D, const float* scales, const float* rot,
                         const float* shs, const float* opacity, const int* radii, const float* means2D, const float* conic,
                         const float* rgb, const uint8_t* clamped, const float* cov3D, const uint32_t* point_list,
                         const uint32_t* ranges, const float* final_T, const uint32_t* n_contrib, const float* dL_dpix,
                         float* dL_dmeans2D, float* dL_dconic4, float* dL_dopacity, float* dL_dcolors, float* dL_dmeans3D,
                         float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drot) {
    return backward_ps1_impl(-INFINITY, cam, P, M, means3D, scales, rot, shs, opacity, radii, means2D, conic, rgb, clamped, cov3D,
                             point_list, ranges, final_T, n_contrib, dL_dpix, dL_dmeans2D, dL_dconic4, dL_dopacity, dL_dcolors,
                             dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drot);
}

void orc_mark_visible(const orc_camera* cam, int P, const float* means3D, uint8_t* present) {
    for (int i = 0; i < P; i++)
        present[i] = xform_row(cam->view, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]) > 0.2f;
}

int orc_version(void) { return 1; }
