/*
 * raster_oracle.c -- CPU restatement of the reference rasterizer.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may load this
 * (through oracle/cpu.py).  The product (frosting_b200/) never does.
 *
 * Each function restates one piece of the reference (DGR = gaussian_splatting/submodules/
 * diff-gaussian-rasterization, CR = DGR/cuda_rasterizer):
 *   oracle_preprocess   CR/forward.cu:155-256 (preprocessCUDA) with in_frustum CR/auxiliary.h:139-164,
 *                       computeCov3D CR/forward.cu:118-152, computeCov2D CR/forward.cu:74-113,
 *                       computeColorFromSH CR/forward.cu:20-71, ndc2Pix/getRect CR/auxiliary.h:41-56
 *   oracle_bin          CR/rasterizer_impl.cu:70-111 (duplicateWithKeys), :300-308 (stable radix sort on
 *                       32+bit key bits), :116-138 (identifyTileRanges), getHigherMsb :35-50
 *   oracle_render_fwd   CR/forward.cu:261-374 (renderCUDA)
 *   oracle_render_bwd   CR/backward.cu:399-557 (renderCUDA backward)
 *   oracle_geom_bwd     CR/backward.cu:144-274 (computeCov2DCUDA), :346-396 (preprocessCUDA backward),
 *                       :20-139 (SH backward), :278-341 (cov3D backward), dnormvdv CR/auxiliary.h:107-117
 *   oracle_mark_visible CR/rasterizer_impl.cu:54-66
 *   oracle_mesh_raster  the clip-space triangle rasterisation Frosting gets from nvdiffrast
 *                       (frosting_utils/nvdiffrast.py:42-54): parity UNPINNED (nvdiffrast is not in the
 *                       reference tree and needs OpenGL), rule-level restatement only.
 *
 * Pinning: the forward/binning/blend/backward functions are pinned against the UNMODIFIED reference
 * compiled from /root/reference (oracle/_ref) on the GPU box -- tests/test_oracle_vs_ref_gpu.py -- and
 * against the golden vectors under tests/golden/ that the same reference produced
 * (tests/golden/make_golden.py).  The reference ships no tests or fixtures of its own.
 *
 * Arithmetic: fp32 with explicit fmaf() in the operation order of the reference's compiled sm_100a
 * code, so every integer-determining quantity (depth bits, radius, tile rect,
 * means2D, conic) is bit-identical to the GPU; compile with -ffp-contract=off.  expf() here is the
 * host libm's, the GPU's is libdevice's ex2.approx sequence: blend weights agree to ~2 ulp, not bitwise.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16

static inline float dot3x(float a0, float b0, float a1, float b1, float a2, float b2) {
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}
static inline float affine_row(const float* m, int r, float x, float y, float z) {
    return dot3x(x, m[r], y, m[4 + r], z, m[8 + r]) + m[12 + r];
}
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* CUDA float->int conversions saturate and map NaN to 0 */
static inline int f2i_rz(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (int)(-2147483647 - 1);
    return (int)v;
}
static inline int f2i_ceil(float v) { return f2i_rz(ceilf(v)); }
static inline float fminf_cuda(float a, float b) { return (a != a) ? b : (b != b) ? a : (a < b ? a : b); }
static inline float fmaxf_cuda(float a, float b) { return (a != a) ? b : (b != b) ? a : (a > b ? a : b); }

static const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
static const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                            -1.0925484305920792f, 0.5462742152960396f};
static const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                            -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

static void cov3d_from_scale_rot(const float* s_in, float mod, const float* q, float* cov) {
    float sx = s_in[0] * mod, sy = s_in[1] * mod, sz = s_in[2] * mod;
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float xz = x * z, rx = r * x, rz = r * z, yy = y * y, zz = z * z;
    float e00 = yy + zz, e11 = fmaf(x, x, zz), e22 = fmaf(x, x, yy);
    float R00 = -(e00 + e00) + 1.0f, R11 = -(e11 + e11) + 1.0f, R22 = -(e22 + e22) + 1.0f;
    float u;
    u = fmaf(x, y, -rz); float R01 = u + u;
    u = fmaf(r, y, xz);  float R02 = u + u;
    u = fmaf(x, y, rz);  float R10 = u + u;
    u = fmaf(y, z, -rx); float R12 = u + u;
    u = fmaf(-r, y, xz); float R20 = u + u;
    u = fmaf(y, z, rx);  float R21 = u + u;
    float M00 = sx * R00, M01 = sy * R01, M02 = sz * R02;
    float M10 = sx * R10, M11 = sy * R11, M12 = sz * R12;
    float M20 = sx * R20, M21 = sy * R21, M22 = sz * R22;
    cov[0] = dot3x(M00, M00, M01, M01, M02, M02);
    cov[1] = dot3x(M00, M10, M01, M11, M02, M12);
    cov[2] = dot3x(M00, M20, M01, M21, M02, M22);
    cov[3] = dot3x(M10, M10, M11, M11, M12, M12);
    cov[4] = dot3x(M10, M20, M11, M21, M12, M22);
    cov[5] = dot3x(M20, M20, M21, M21, M22, M22);
}

static void cov2d(float px, float py, float pz, float fx, float fy, float tanx, float tany, const float* S,
                  const float* v, float* out3) {
    float tx = affine_row(v, 0, px, py, pz), ty = affine_row(v, 1, px, py, pz), tz = affine_row(v, 2, px, py, pz);
    float limx = tanx * 1.3f, limy = tany * 1.3f;
    float txtz = tx / tz, tytz = ty / tz;
    float cx = fminf_cuda(fmaxf_cuda(txtz, -limx), limx);
    float cy = fminf_cuda(fmaxf_cuda(tytz, -limy), limy);
    float tz2 = tz * tz;
    float J00 = fx / tz, J02 = ((tz * -cx) * fx) / tz2, J11 = fy / tz, J12 = ((tz * -cy) * fy) / tz2;
    float T00 = fmaf(v[2], J02, v[0] * J00), T01 = fmaf(v[6], J02, v[4] * J00), T02 = fmaf(v[10], J02, v[8] * J00);
    float T10 = fmaf(v[2], J12, v[1] * J11), T11 = fmaf(v[6], J12, v[5] * J11), T12 = fmaf(v[10], J12, v[9] * J11);
    float P00 = dot3x(T00, S[0], T01, S[1], T02, S[2]);
    float P10 = dot3x(T00, S[1], T01, S[3], T02, S[4]);
    float P20 = dot3x(T00, S[2], T01, S[4], T02, S[5]);
    float P01 = dot3x(T10, S[0], T11, S[1], T12, S[2]);
    float P11 = dot3x(T10, S[1], T11, S[3], T12, S[4]);
    float P21 = dot3x(T10, S[2], T11, S[4], T12, S[5]);
    out3[0] = dot3x(P00, T00, P10, T01, P20, T02) + 0.3f;
    out3[1] = dot3x(P01, T00, P11, T01, P21, T02);
    out3[2] = dot3x(P01, T10, P11, T11, P21, T12) + 0.3f;
}

static inline float ndc2pix(float v, int S) { return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5); }
static inline uint32_t rect_coord(float v, uint32_t g) {
    int i = f2i_rz(v * 0.0625f);
    if (i < 0) i = 0;
    return (uint32_t)i < g ? (uint32_t)i : g;
}

static void eval_sh(int deg, const float* c, float x, float y, float z, float* out) {
    float r[3];
    for (int k = 0; k < 3; ++k) r[k] = c[k] * C0;
    if (deg > 0) {
        float k1 = y * C1, k2 = z * C1, k3 = x * C1;
        for (int k = 0; k < 3; ++k) { r[k] = fmaf(-k1, c[3 + k], r[k]); r[k] = fmaf(k2, c[6 + k], r[k]); r[k] = fmaf(-k3, c[9 + k], r[k]); }
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = y * x, yz = z * y, xz = z * x;
            float zz2 = zz + zz, xx_yy = xx + -yy;
            float b4 = xy * C2[0], b5 = yz * C2[1], b6 = (-yy + (-xx + zz2)) * C2[2], b7 = xz * C2[3], b8 = xx_yy * C2[4];
            for (int k = 0; k < 3; ++k) {
                r[k] = fmaf(b4, c[12 + k], r[k]); r[k] = fmaf(b5, c[15 + k], r[k]); r[k] = fmaf(b6, c[18 + k], r[k]);
                r[k] = fmaf(b7, c[21 + k], r[k]); r[k] = fmaf(b8, c[24 + k], r[k]);
            }
            if (deg > 2) {
                float f4 = -yy + fmaf(zz, 4.0f, -xx);
                float b9 = (y * C3[0]) * fmaf(xx, 3.0f, -yy);
                float b10 = (xy * C3[1]) * z;
                float b11 = (y * C3[2]) * f4;
                float b12 = (z * C3[3]) * fmaf(yy, -3.0f, fmaf(xx, -3.0f, zz2));
                float b13 = f4 * (x * C3[4]);
                float b14 = xx_yy * (z * C3[5]);
                float b15 = (x * C3[6]) * fmaf(yy, -3.0f, xx);
                for (int k = 0; k < 3; ++k) {
                    r[k] = fmaf(b9, c[27 + k], r[k]); r[k] = fmaf(b10, c[30 + k], r[k]); r[k] = fmaf(b11, c[33 + k], r[k]);
                    r[k] = fmaf(b12, c[36 + k], r[k]); r[k] = fmaf(b13, c[39 + k], r[k]); r[k] = fmaf(b14, c[42 + k], r[k]);
                    r[k] = fmaf(b15, c[45 + k], r[k]);
                }
            }
        }
    }
    for (int k = 0; k < 3; ++k) out[k] = r[k];
}

/* Outputs are per-Gaussian arrays in the reference's GeometryState layout (rasterizer_impl.h:30-45);
 * rect is (min.x, min.y, max.x, max.y). Entries of culled Gaussians: radii = 0, tiles_touched = 0, rest untouched. */
void oracle_preprocess(int P, int D, int M, const float* means, const float* scales, float mod, const float* rots,
                       const float* opac, const float* shs, const float* cov3D_precomp, const float* colors_precomp,
                       const float* view, const float* proj, const float* campos, int W, int H, float tanx, float tany,
                       const uint8_t* visibility,
                       int32_t* radii, float* xy, float* depths, float* cov3D, float* conic_opacity, float* rgb,
                       uint8_t* clamped, uint32_t* tiles_touched, uint32_t* rect) {
    const float focal_y = H / (2.0f * tany), focal_x = W / (2.0f * tanx);
    const uint32_t gx = (uint32_t)((W + TILE - 1) / TILE), gy = (uint32_t)((H + TILE - 1) / TILE);
    for (int i = 0; i < P; ++i) {
        radii[i] = 0; tiles_touched[i] = 0;
        rect[4 * i] = rect[4 * i + 1] = rect[4 * i + 2] = rect[4 * i + 3] = 0;
        float px = means[3 * i], py = means[3 * i + 1], pz = means[3 * i + 2];
        float depth = affine_row(view, 2, px, py, pz);
        if (depth <= 0.2f) continue;
        if (visibility && !visibility[i]) continue;
        float hx = affine_row(proj, 0, px, py, pz), hy = affine_row(proj, 1, px, py, pz), hw = affine_row(proj, 3, px, py, pz);
        float p_w = 1.0f / (hw + 0.0000001f);
        float projx = hx * p_w, projy = hy * p_w;
        float S[6];
        if (cov3D_precomp) memcpy(S, cov3D_precomp + 6 * (size_t)i, 24);
        else { cov3d_from_scale_rot(scales + 3 * (size_t)i, mod, rots + 4 * (size_t)i, S); memcpy(cov3D + 6 * (size_t)i, S, 24); }
        float cov[3];
        cov2d(px, py, pz, focal_x, focal_y, tanx, tany, S, view, cov);
        float det = fmaf(cov[0], cov[2], -(cov[1] * cov[1]));
        if (det == 0.0f) continue;
        float det_inv = 1.0f / det;
        float conic[3] = {cov[2] * det_inv, cov[1] * -det_inv, cov[0] * det_inv};
        float mid = (cov[0] + cov[2]) * 0.5f;
        float sq = sqrtf(fmaxf_cuda(fmaf(mid, mid, -det), 0.1f));
        float l1 = mid + sq, l2 = mid + -sq;
        int my_radius = f2i_ceil(sqrtf(fmaxf_cuda(l1, l2)) * 3.0f);
        float Rf = (float)my_radius;
        float pix_x = ndc2pix(projx, W), pix_y = ndc2pix(projy, H);
        uint32_t minx = rect_coord(pix_x + -Rf, gx), miny = rect_coord(pix_y + -Rf, gy);
        uint32_t maxx = rect_coord(((pix_x + Rf) + 16.0f) + -1.0f, gx), maxy = rect_coord(((pix_y + Rf) + 16.0f) + -1.0f, gy);
        uint32_t touched = (maxx - minx) * (maxy - miny);
        if (touched == 0) continue;
        if (!colors_precomp) {
            float dx = -campos[0] + px, dy = -campos[1] + py, dz = -campos[2] + pz;
            float len = sqrtf(dot3x(dx, dx, dy, dy, dz, dz));
            float res[3];
            eval_sh(D, shs + (size_t)i * M * 3, dx / len, dy / len, dz / len, res);
            for (int k = 0; k < 3; ++k) {
                float s = res[k] + 0.5f;
                clamped[3 * i + k] = s < 0.f;
                rgb[3 * i + k] = (s < 0.f) ? 0.f : s;
            }
        }
        depths[i] = depth;
        radii[i] = my_radius;
        xy[2 * i] = pix_x; xy[2 * i + 1] = pix_y;
        conic_opacity[4 * i] = conic[0]; conic_opacity[4 * i + 1] = conic[1]; conic_opacity[4 * i + 2] = conic[2];
        conic_opacity[4 * i + 3] = opac[i];
        tiles_touched[i] = touched;
        rect[4 * i] = minx; rect[4 * i + 1] = miny; rect[4 * i + 2] = maxx; rect[4 * i + 3] = maxy;
    }
}

void oracle_mark_visible(int P, const float* means, const float* view, uint8_t* present) {
    for (int i = 0; i < P; ++i)
        present[i] = !(affine_row(view, 2, means[3 * i], means[3 * i + 1], means[3 * i + 2]) <= 0.2f);
}

/* getHigherMsb, rasterizer_impl.cu:35-50 */
static uint32_t higher_msb(uint32_t n) {
    uint32_t msb = 16, step = 16;
    while (step > 1) { step /= 2; if (n >> msb) msb += step; else msb -= step; }
    if (n >> msb) msb++;
    return msb;
}

typedef struct { uint64_t key; uint32_t val; } kv_t;

/* Emits keys (tile<<32 | depth bits) in Gaussian order, sorts them with a STABLE LSD radix sort on bits
 * [0, 32+bit) like cub::DeviceRadixSort, fills ranges.  Returns R; call with keys==NULL to only count. */
int64_t oracle_bin(int P, int W, int H, const int32_t* radii, const uint32_t* rect, const float* depths,
                   uint64_t* keys_sorted, uint32_t* point_list, uint32_t* ranges /* 2T */) {
    const uint32_t gx = (uint32_t)((W + TILE - 1) / TILE), gy = (uint32_t)((H + TILE - 1) / TILE);
    int64_t R = 0;
    for (int i = 0; i < P; ++i)
        if (radii[i] > 0) R += (int64_t)(rect[4 * i + 2] - rect[4 * i]) * (rect[4 * i + 3] - rect[4 * i + 1]);
    if (!keys_sorted) return R;
    uint64_t* ka = keys_sorted; uint32_t* va = point_list;
    uint64_t* kb = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(R ? R : 1));
    uint32_t* vb = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(R ? R : 1));
    int64_t off = 0;
    for (int i = 0; i < P; ++i) {
        if (radii[i] <= 0) continue;
        for (uint32_t y = rect[4 * i + 1]; y < rect[4 * i + 3]; ++y)
            for (uint32_t x = rect[4 * i]; x < rect[4 * i + 2]; ++x) {
                ka[off] = ((uint64_t)(y * gx + x) << 32) | f2u(depths[i]);
                va[off] = (uint32_t)i;
                ++off;
            }
    }
    const int end_bit = 32 + (int)higher_msb(gx * gy);
    for (int shift = 0; shift < end_bit; shift += 8) {
        const int bits = (end_bit - shift) < 8 ? (end_bit - shift) : 8;
        const uint32_t mask = (1u << bits) - 1u;
        size_t hist[257]; memset(hist, 0, sizeof(hist));
        for (int64_t j = 0; j < R; ++j) hist[((ka[j] >> shift) & mask) + 1]++;
        for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
        for (int64_t j = 0; j < R; ++j) { size_t d = hist[(ka[j] >> shift) & mask]++; kb[d] = ka[j]; vb[d] = va[j]; }
        uint64_t* tk = ka; ka = kb; kb = tk; uint32_t* tv = va; va = vb; vb = tv;
    }
    if (ka != keys_sorted) { memcpy(keys_sorted, ka, sizeof(uint64_t) * (size_t)R); memcpy(point_list, va, sizeof(uint32_t) * (size_t)R); free(ka); free(va); }
    else { free(kb); free(vb); }
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
    for (int64_t j = 0; j < R; ++j) {
        uint32_t cur = (uint32_t)(keys_sorted[j] >> 32);
        if (j == 0) ranges[2 * cur] = 0;
        else { uint32_t prev = (uint32_t)(keys_sorted[j - 1] >> 32); if (cur != prev) { ranges[2 * prev + 1] = (uint32_t)j; ranges[2 * cur] = (uint32_t)j; } }
        if (j == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
    }
    return R;
}

static inline float blend_power(const float* xy, const float* co, float pxf, float pyf, float* dx, float* dy) {
    *dx = xy[0] + -pxf; *dy = xy[1] + -pyf;
    float q = fmaf(*dx, *dx * co[0], *dy * (*dy * co[2]));
    float u = *dy * (*dx * co[1]);
    return fmaf(q, -0.5f, -u);
}

void oracle_render_fwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* xy,
                       const float* features, const float* conic_opacity, const float* bg,
                       float* final_T, uint32_t* n_contrib, float* out_color) {
    const int gx = (W + TILE - 1) / TILE;
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            const int tile = (py / TILE) * gx + (px / TILE);
            const uint32_t lo = ranges[2 * tile], hi = ranges[2 * tile + 1];
            float T = 1.0f, C[3] = {0, 0, 0};
            uint32_t contributor = 0, last = 0;
            for (uint32_t j = lo; j < hi; ++j) {
                contributor++;
                const uint32_t id = point_list[j];
                float dx, dy;
                float power = blend_power(xy + 2 * (size_t)id, conic_opacity + 4 * (size_t)id, (float)px, (float)py, &dx, &dy);
                if (power > 0.0f) continue;
                float alpha = fminf_cuda(0.99f, conic_opacity[4 * (size_t)id + 3] * expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                float test_T = T * (1.0f + -alpha);
                if (test_T < 0.0001f) break;
                for (int ch = 0; ch < 3; ++ch) C[ch] = fmaf(T, alpha * features[3 * (size_t)id + ch], C[ch]);
                T = test_T;
                last = contributor;
            }
            const size_t pid = (size_t)py * W + px;
            final_T[pid] = T; n_contrib[pid] = last;
            for (int ch = 0; ch < 3; ++ch) out_color[(size_t)ch * H * W + pid] = C[ch] + T * bg[ch];
        }
}

/* Gradient accumulators are double (the reference uses unordered float atomics); per-pair terms are float. */
void oracle_render_bwd(int P, int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* bg,
                       const float* xy, const float* conic_opacity, const float* colors, const float* final_T,
                       const uint32_t* n_contrib, const float* dL_dpix,
                       double* dL_dmean2D /* [P,2] */, double* dL_dconic /* [P,3] x,y,w */, double* dL_dopacity, double* dL_dcolors) {
    const int gx = (W + TILE - 1) / TILE;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    memset(dL_dmean2D, 0, sizeof(double) * 2 * (size_t)P); memset(dL_dconic, 0, sizeof(double) * 3 * (size_t)P);
    memset(dL_dopacity, 0, sizeof(double) * (size_t)P); memset(dL_dcolors, 0, sizeof(double) * 3 * (size_t)P);
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            const int tile = (py / TILE) * gx + (px / TILE);
            const uint32_t lo = ranges[2 * tile];
            const size_t pid = (size_t)py * W + px;
            const float T_final = final_T[pid];
            float T = T_final;
            const uint32_t last = n_contrib[pid];
            float dp[3] = {dL_dpix[pid], dL_dpix[(size_t)H * W + pid], dL_dpix[2 * (size_t)H * W + pid]};
            float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0;
            float bg_dot = bg[0] * dp[0] + bg[1] * dp[1] + bg[2] * dp[2];
            for (int64_t pos = (int64_t)last - 1; pos >= 0; --pos) {
                const uint32_t id = point_list[lo + pos];
                const float* co = conic_opacity + 4 * (size_t)id;
                float dx, dy;
                float power = blend_power(xy + 2 * (size_t)id, co, (float)px, (float)py, &dx, &dy);
                if (power > 0.0f) continue;
                float G = expf(power);
                float alpha = fminf_cuda(0.99f, co[3] * G);
                if (alpha < 1.0f / 255.0f) continue;
                T = T / (1.f - alpha);
                float dchannel_dcolor = alpha * T;
                float dL_dalpha = 0.0f;
                for (int ch = 0; ch < 3; ++ch) {
                    float c = colors[3 * (size_t)id + ch];
                    accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                    last_color[ch] = c;
                    dL_dalpha += (c - accum_rec[ch]) * dp[ch];
                    dL_dcolors[3 * (size_t)id + ch] += dchannel_dcolor * dp[ch];
                }
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                float dL_dG = co[3] * dL_dalpha;
                float gdx = G * dx, gdy = G * dy;
                float dG_ddelx = -gdx * co[0] - gdy * co[1];
                float dG_ddely = -gdy * co[2] - gdx * co[1];
                dL_dmean2D[2 * (size_t)id] += dL_dG * dG_ddelx * ddelx_dx;
                dL_dmean2D[2 * (size_t)id + 1] += dL_dG * dG_ddely * ddely_dy;
                dL_dconic[3 * (size_t)id] += -0.5f * gdx * dx * dL_dG;
                dL_dconic[3 * (size_t)id + 1] += -0.5f * gdx * dy * dL_dG;
                dL_dconic[3 * (size_t)id + 2] += -0.5f * gdy * dy * dL_dG;
                dL_dopacity[id] += G * dL_dalpha;
            }
        }
}

/* computeCov2DCUDA + preprocessCUDA(backward).  Inputs dL_dmean2D[P,2], dL_dconic[P,3] (x,y,w), dL_dcolor[P,3]
 * as float; outputs as the reference's tensors. */
void oracle_geom_bwd(int P, int D, int M, const float* means, const int32_t* radii, const float* shs,
                     const uint8_t* clamped, const float* scales, const float* rots, float mod, const float* cov3Ds,
                     const float* view, const float* proj, const float* campos, int W, int H, float tanx, float tany,
                     const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor,
                     float* dL_dmeans, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot) {
    const float h_y = H / (2.0f * tany), h_x = W / (2.0f * tanx);
    memset(dL_dmeans, 0, sizeof(float) * 3 * (size_t)P); memset(dL_dcov3D, 0, sizeof(float) * 6 * (size_t)P);
    if (dL_dsh) memset(dL_dsh, 0, sizeof(float) * 3 * (size_t)M * P);
    memset(dL_dscale, 0, sizeof(float) * 3 * (size_t)P); memset(dL_drot, 0, sizeof(float) * 4 * (size_t)P);
    for (int i = 0; i < P; ++i) {
        if (!(radii[i] > 0)) continue;
        const float* c3 = cov3Ds + 6 * (size_t)i;
        const float m[3] = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
        const float dLc[3] = {dL_dconic[3 * i], dL_dconic[3 * i + 1], dL_dconic[3 * i + 2]};
        float t[3];
        for (int r = 0; r < 3; ++r) t[r] = view[r] * m[0] + view[4 + r] * m[1] + view[8 + r] * m[2] + view[12 + r];
        const float limx = 1.3f * tanx, limy = 1.3f * tany;
        const float txtz = t[0] / t[2], tytz = t[1] / t[2];
        t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
        t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
        const float xg = (txtz < -limx || txtz > limx) ? 0.f : 1.f, yg = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        /* T rows (2x3) = J * W, W[i][j] = view[4j+i] */
        const float J00 = h_x / t[2], J02 = -(h_x * t[0]) / (t[2] * t[2]), J11 = h_y / t[2], J12 = -(h_y * t[1]) / (t[2] * t[2]);
        float T0[3], T1[3];
        for (int j = 0; j < 3; ++j) { T0[j] = J00 * view[4 * j] + J02 * view[4 * j + 2]; T1[j] = J11 * view[4 * j + 1] + J12 * view[4 * j + 2]; }
        const float V[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
        float VT0[3], VT1[3];
        for (int r = 0; r < 3; ++r) { VT0[r] = V[r][0] * T0[0] + V[r][1] * T0[1] + V[r][2] * T0[2]; VT1[r] = V[r][0] * T1[0] + V[r][1] * T1[1] + V[r][2] * T1[2]; }
        const float a = T0[0] * VT0[0] + T0[1] * VT0[1] + T0[2] * VT0[2] + 0.3f;
        const float b = T0[0] * VT1[0] + T0[1] * VT1[1] + T0[2] * VT1[2];
        const float c = T1[0] * VT1[0] + T1[1] * VT1[1] + T1[2] * VT1[2] + 0.3f;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float* dcov = dL_dcov3D + 6 * (size_t)i;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dLc[0] + 2 * b * c * dLc[1] + (denom - a * c) * dLc[2]);
            dL_dc = denom2inv * (-a * a * dLc[2] + 2 * a * b * dLc[1] + (denom - a * c) * dLc[0]);
            dL_db = denom2inv * 2 * (b * c * dLc[0] - (denom + 2 * b * b) * dLc[1] + a * b * dLc[2]);
            dcov[0] = T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc;
            dcov[3] = T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc;
            dcov[5] = T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc;
            dcov[1] = 2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2 * T1[0] * T1[1] * dL_dc;
            dcov[2] = 2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2 * T1[0] * T1[2] * dL_dc;
            dcov[4] = 2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2 * T1[1] * T1[2] * dL_dc;
        }
        float dT0[3], dT1[3];
        for (int r = 0; r < 3; ++r) { dT0[r] = 2 * VT0[r] * dL_da + VT1[r] * dL_db; dT1[r] = 2 * VT1[r] * dL_dc + VT0[r] * dL_db; }
        const float dJ00 = view[0] * dT0[0] + view[4] * dT0[1] + view[8] * dT0[2];
        const float dJ02 = view[2] * dT0[0] + view[6] * dT0[1] + view[10] * dT0[2];
        const float dJ11 = view[1] * dT1[0] + view[5] * dT1[1] + view[9] * dT1[2];
        const float dJ12 = view[2] * dT1[0] + view[6] * dT1[1] + view[10] * dT1[2];
        const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = xg * -h_x * tz2 * dJ02, dty = yg * -h_y * tz2 * dJ12;
        const float dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * t[0]) * tz3 * dJ02 + (2 * h_y * t[1]) * tz3 * dJ12;
        float gm[3] = {view[0] * dtx + view[1] * dty + view[2] * dtz, view[4] * dtx + view[5] * dty + view[6] * dtz,
                       view[8] * dtx + view[9] * dty + view[10] * dtz};
        /* projection term, backward.cu:362-387 */
        {
            float hx = proj[0] * m[0] + proj[4] * m[1] + proj[8] * m[2] + proj[12];
            float hy = proj[1] * m[0] + proj[5] * m[1] + proj[9] * m[2] + proj[13];
            float hw = proj[3] * m[0] + proj[7] * m[1] + proj[11] * m[2] + proj[15];
            float m_w = 1.0f / (hw + 0.0000001f);
            float mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
            float gx2 = dL_dmean2D[2 * i], gy2 = dL_dmean2D[2 * i + 1];
            for (int j = 0; j < 3; ++j)
                gm[j] += (proj[4 * j] * m_w - proj[4 * j + 3] * mul1) * gx2 + (proj[4 * j + 1] * m_w - proj[4 * j + 3] * mul2) * gy2;
        }
        if (shs) {
            float dRGB[3];
            for (int k = 0; k < 3; ++k) dRGB[k] = clamped[3 * i + k] ? 0.f : dL_dcolor[3 * i + k];
            float dor[3] = {m[0] - campos[0], m[1] - campos[1], m[2] - campos[2]};
            float len = sqrtf(dor[0] * dor[0] + dor[1] * dor[1] + dor[2] * dor[2]);
            float x = dor[0] / len, y = dor[1] / len, z = dor[2] / len;
            const float* sh = shs + (size_t)i * M * 3;
            float* dsh = dL_dsh + (size_t)i * M * 3;
            float bv[16] = {0}, bx[16] = {0}, by[16] = {0}, bz[16] = {0};
            bv[0] = C0;
            if (D > 0) {
                bv[1] = -C1 * y; by[1] = -C1; bv[2] = C1 * z; bz[2] = C1; bv[3] = -C1 * x; bx[3] = -C1;
                if (D > 1) {
                    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    bv[4] = C2[0] * xy; bx[4] = C2[0] * y; by[4] = C2[0] * x;
                    bv[5] = C2[1] * yz; by[5] = C2[1] * z; bz[5] = C2[1] * y;
                    bv[6] = C2[2] * (2.f * zz - xx - yy); bx[6] = C2[2] * 2.f * -x; by[6] = C2[2] * 2.f * -y; bz[6] = C2[2] * 2.f * 2.f * z;
                    bv[7] = C2[3] * xz; bx[7] = C2[3] * z; bz[7] = C2[3] * x;
                    bv[8] = C2[4] * (xx - yy); bx[8] = C2[4] * 2.f * x; by[8] = C2[4] * 2.f * -y;
                    if (D > 2) {
                        bv[9] = C3[0] * y * (3.f * xx - yy); bx[9] = C3[0] * 3.f * 2.f * xy; by[9] = C3[0] * 3.f * (xx - yy);
                        bv[10] = C3[1] * xy * z; bx[10] = C3[1] * yz; by[10] = C3[1] * xz; bz[10] = C3[1] * xy;
                        bv[11] = C3[2] * y * (4.f * zz - xx - yy); bx[11] = C3[2] * -2.f * xy; by[11] = C3[2] * (-3.f * yy + 4.f * zz - xx); bz[11] = C3[2] * 4.f * 2.f * yz;
                        bv[12] = C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); bx[12] = C3[3] * -3.f * 2.f * xz; by[12] = C3[3] * -3.f * 2.f * yz; bz[12] = C3[3] * 3.f * (2.f * zz - xx - yy);
                        bv[13] = C3[4] * x * (4.f * zz - xx - yy); bx[13] = C3[4] * (-3.f * xx + 4.f * zz - yy); by[13] = C3[4] * -2.f * xy; bz[13] = C3[4] * 4.f * 2.f * xz;
                        bv[14] = C3[5] * z * (xx - yy); bx[14] = C3[5] * 2.f * xz; by[14] = C3[5] * -2.f * yz; bz[14] = C3[5] * (xx - yy);
                        bv[15] = C3[6] * x * (xx - 3.f * yy); bx[15] = C3[6] * 3.f * (xx - yy); by[15] = C3[6] * -3.f * 2.f * xy;
                    }
                }
            }
            int ncoef = (D + 1) * (D + 1);
            float ddir[3] = {0, 0, 0};
            for (int k = 0; k < ncoef && k < 16; ++k)
                for (int ch = 0; ch < 3; ++ch) {
                    dsh[3 * k + ch] = bv[k] * dRGB[ch];
                    ddir[0] += bx[k] * sh[3 * k + ch] * dRGB[ch];
                    ddir[1] += by[k] * sh[3 * k + ch] * dRGB[ch];
                    ddir[2] += bz[k] * sh[3 * k + ch] * dRGB[ch];
                }
            float sum2 = dor[0] * dor[0] + dor[1] * dor[1] + dor[2] * dor[2];
            float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            float vd = dor[0] * ddir[0] + dor[1] * ddir[1] + dor[2] * ddir[2];
            for (int j = 0; j < 3; ++j) gm[j] += (ddir[j] * sum2 - dor[j] * vd) * inv32;
        }
        for (int j = 0; j < 3; ++j) dL_dmeans[3 * i + j] = gm[j];
        if (scales) {
            const float* q = rots + 4 * (size_t)i;
            const float r = q[0], x = q[1], y = q[2], z = q[3];
            const float s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
            /* Rm[row][col]: Sigma = Rm^T diag(s)^2 Rm */
            const float Rm[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y + r * z), 2.f * (x * z - r * y)},
                                    {2.f * (x * y - r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + r * x)},
                                    {2.f * (x * z + r * y), 2.f * (y * z - r * x), 1.f - 2.f * (x * x + y * y)}};
            const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]}, {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            float G[3][3];
            for (int a2 = 0; a2 < 3; ++a2) {
                float acc = 0;
                for (int b2 = 0; b2 < 3; ++b2) {
                    float dM = 2.f * s[a2] * (Rm[a2][0] * dS[0][b2] + Rm[a2][1] * dS[1][b2] + Rm[a2][2] * dS[2][b2]);
                    acc += Rm[a2][b2] * dM;
                    G[a2][b2] = s[a2] * dM;
                }
                dL_dscale[3 * i + a2] = acc;
            }
            float* dq = dL_drot + 4 * (size_t)i;
            dq[0] = 2 * z * (G[0][1] - G[1][0]) + 2 * y * (G[2][0] - G[0][2]) + 2 * x * (G[1][2] - G[2][1]);
            dq[1] = 2 * y * (G[1][0] + G[0][1]) + 2 * z * (G[2][0] + G[0][2]) + 2 * r * (G[1][2] - G[2][1]) - 4 * x * (G[2][2] + G[1][1]);
            dq[2] = 2 * x * (G[1][0] + G[0][1]) + 2 * r * (G[2][0] - G[0][2]) + 2 * z * (G[1][2] + G[2][1]) - 4 * y * (G[2][2] + G[0][0]);
            dq[3] = 2 * r * (G[0][1] - G[1][0]) + 2 * x * (G[2][0] + G[0][2]) + 2 * y * (G[1][2] + G[2][1]) - 4 * z * (G[1][1] + G[0][0]);
        }
    }
}

/* ---- mesh rasterisation (occlusion prepass) -- parity UNPINNED, see header comment ------------------ */
typedef struct { float x, y, z; } sv_t;

static float edge_fn(sv_t a, sv_t b, float px, float py) {
    int swap = (a.x > b.x) || (a.x == b.x && a.y > b.y);
    sv_t p = swap ? b : a, q = swap ? a : b;
    float e = (q.x - p.x) * (py - p.y) - (q.y - p.y) * (px - p.x);
    return swap ? -e : e;
}
static int top_left(sv_t a, sv_t b) { float dx = b.x - a.x, dy = b.y - a.y; return (dy == 0.f && dx > 0.f) || (dy < 0.f); }

static void raster_tri(sv_t v0, sv_t v1, sv_t v2, int W, int H, int face, uint64_t* zbuf) {
    float area = edge_fn(v0, v1, v2.x, v2.y);
    if (area == 0.f || area != area) return;
    if (area < 0.f) { sv_t t = v1; v1 = v2; v2 = t; area = -area; }
    float minx = fminf(v0.x, fminf(v1.x, v2.x)), maxx = fmaxf(v0.x, fmaxf(v1.x, v2.x));
    float miny = fminf(v0.y, fminf(v1.y, v2.y)), maxy = fmaxf(v0.y, fmaxf(v1.y, v2.y));
    int x0 = (int)ceilf(minx - 0.5f), x1 = (int)floorf(maxx - 0.5f), y0 = (int)ceilf(miny - 0.5f), y1 = (int)floorf(maxy - 0.5f);
    if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0; if (x1 > W - 1) x1 = W - 1; if (y1 > H - 1) y1 = H - 1;
    int tl0 = top_left(v1, v2), tl1 = top_left(v2, v0), tl2 = top_left(v0, v1);
    float inv_area = 1.0f / area;
    for (int py = y0; py <= y1; ++py)
        for (int px = x0; px <= x1; ++px) {
            float cx = px + 0.5f, cy = py + 0.5f;
            float w0 = edge_fn(v1, v2, cx, cy), w1 = edge_fn(v2, v0, cx, cy), w2 = edge_fn(v0, v1, cx, cy);
            if (!((w0 > 0.f || (w0 == 0.f && tl0)) && (w1 > 0.f || (w1 == 0.f && tl1)) && (w2 > 0.f || (w2 == 0.f && tl2)))) continue;
            float z = (((w0 * v0.z) + (w1 * v1.z)) + (w2 * v2.z)) * inv_area;
            if (!(z >= 0.f && z <= 1.f)) continue;
            uint64_t key = ((uint64_t)f2u(z) << 32) | (uint32_t)face;
            uint64_t* p = zbuf + (size_t)py * W + px;
            if (key < *p) *p = key;
        }
}

typedef struct { float x, y, z, w; } cv_t;
static sv_t to_screen(cv_t c, int W, int H) {
    float iw = 1.0f / c.w; sv_t s;
    s.x = (((c.x * iw) * 0.5f) + 0.5f) * (float)W; s.y = (((c.y * iw) * 0.5f) + 0.5f) * (float)H; s.z = ((c.z * iw) * 0.5f) + 0.5f;
    return s;
}

void oracle_mesh_raster(int V, int F, const float* verts, const int32_t* faces, const float* m, int W, int H,
                        int32_t* pix_to_face, uint8_t* face_visible, int mark_last_on_bg) {
    (void)V;
    size_t N = (size_t)W * H;
    uint64_t* zbuf = (uint64_t*)malloc(sizeof(uint64_t) * N);
    memset(zbuf, 0xff, sizeof(uint64_t) * N);
    for (int f = 0; f < F; ++f) {
        cv_t c[3];
        for (int k = 0; k < 3; ++k) {
            const float* p = verts + 3 * (size_t)faces[3 * f + k];
            c[k].x = affine_row(m, 0, p[0], p[1], p[2]);
            c[k].y = affine_row(m, 1, p[0], p[1], p[2]);
            c[k].z = affine_row(m, 2, p[0], p[1], p[2]);
            c[k].w = affine_row(m, 3, p[0], p[1], p[2]);
        }
        float d[3]; int in[3], nin = 0;
        for (int k = 0; k < 3; ++k) { d[k] = c[k].z + c[k].w; in[k] = d[k] >= 0.f && c[k].w > 1e-12f; nin += in[k]; }
        if (nin == 0) continue;
        if (nin == 3) { raster_tri(to_screen(c[0], W, H), to_screen(c[1], W, H), to_screen(c[2], W, H), W, H, f, zbuf); continue; }
        cv_t poly[4]; int np = 0;
        for (int e = 0; e < 3; ++e) {
            int a = e, b = (e + 1) % 3;
            if (in[a]) poly[np++] = c[a];
            if (in[a] != in[b]) {
                float t = d[a] / (d[a] - d[b]);
                cv_t p;
                p.x = fmaf(c[b].x - c[a].x, t, c[a].x); p.y = fmaf(c[b].y - c[a].y, t, c[a].y);
                p.z = fmaf(c[b].z - c[a].z, t, c[a].z); p.w = fmaf(c[b].w - c[a].w, t, c[a].w);
                if (!(p.w > 1e-12f)) p.w = 1e-12f;
                poly[np++] = p;
            }
        }
        if (np < 3) continue;
        sv_t s0 = to_screen(poly[0], W, H), s1 = to_screen(poly[1], W, H), s2 = to_screen(poly[2], W, H);
        raster_tri(s0, s1, s2, W, H, f, zbuf);
        if (np == 4) raster_tri(s0, s2, to_screen(poly[3], W, H), W, H, f, zbuf);
    }
    if (face_visible) memset(face_visible, 0, (size_t)F);
    for (size_t i = 0; i < N; ++i) {
        int face = (zbuf[i] == ~0ull) ? -1 : (int)(uint32_t)zbuf[i];
        pix_to_face[i] = face;
        if (face_visible) { if (face >= 0) face_visible[face] = 1; else if (mark_last_on_bg && F > 0) face_visible[F - 1] = 1; }
    }
    free(zbuf);
}
