// render_bwd.cu -- backward of the alpha compositing, one CTA per 16x16 tile, warp-autonomous.
//
// Replaces renderCUDA<3> backward (DGR/cuda_rasterizer/backward.cu:399-557): back-to-front
// traversal from each pixel's last contributor, T recovered by division, gradients w.r.t. the
// per-Gaussian 2D mean, conic, opacity and colour.  Same formulas; B200-first structure:
//   * the reference issues 9 global float atomics per contributing PIXEL-Gaussian pair (its dominant
//     cost).  Here a warp (8x4 pixels) reduces its 32 lanes' contributions with a transposing
//     butterfly (14 shuffles for 9 values) and issues ONE red.global.add instruction per
//     warp-Gaussian pair (9 active lanes, 9 addresses in one 48-byte accumulator record);
//   * every warp walks the list on its own, 32 instances per step, starting at ITS deepest last
//     contributor; no CTA barrier (v1 staged 256-instance batches cooperatively and was dominated by
//     barrier stalls, profiles/r01_render_c3_v1_summary.json).  Records re-read by the 8 warps of a
//     tile are L1/L2 hits;
//   * the same conservative alpha>=1/255 extents as the forward kernel let a warp skip instances
//     that cannot touch its sub-tile.
#include "common.cuh"

namespace fb200 {

namespace {

struct __align__(16) WarpSlabB {
    float4 q0[32];
    float4 q1[32];
    float cb[32];
    uint32_t id[32];
};

__device__ __forceinline__ bool overlaps(float lo, float hi, float c, float ext) {
    return !(c + ext < lo) && !(c - ext > hi);
}

// Sum each of v[0..7] and v8 over the 32 lanes with 12 shuffles: a transposing butterfly halves the number
// of live values at every step (8 -> 4 -> 2 -> 1) while the ninth value rides along and is folded into the
// last transposing step.  On return lane 4k+{0,1} ... : lanes with (lane & 3) < 2 hold
//   value index  (lane >> 2)            if (lane & 2) == 0   [k = 0..7]
// and lanes with (lane & 2) != 0 hold value 8.  The caller uses lanes with (lane & 3) == 0 (values 0..7)
// and lane 2 (value 8).
__device__ __forceinline__ float warp_reduce9(const float (&v)[8], float v8, bool b4, bool b3, bool b2, bool b1) {
    const unsigned full = 0xffffffffu;
    float w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b4 ? v[i] : v[i + 4];
        const float keep = b4 ? v[i + 4] : v[i];
        w[i] = keep + __shfl_xor_sync(full, send, 16);
    }
    v8 += __shfl_xor_sync(full, v8, 16);
    float x[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b3 ? w[i] : w[i + 2];
        const float keep = b3 ? w[i + 2] : w[i];
        x[i] = keep + __shfl_xor_sync(full, send, 8);
    }
    v8 += __shfl_xor_sync(full, v8, 8);
    float y;
    {
        const float send = b2 ? x[0] : x[1];
        const float keep = b2 ? x[1] : x[0];
        y = keep + __shfl_xor_sync(full, send, 4);
    }
    v8 += __shfl_xor_sync(full, v8, 4);
    // y: value index (lane>>2)&7 summed over lane bits 4,3,2; v8 summed over the same bits.
    // fold the pair (y, v8) over lane bit 1: lanes with bit1 = 0 keep y, lanes with bit1 = 1 keep v8
    float z;
    {
        const float send = b1 ? y : v8;
        const float keep = b1 ? v8 : y;
        z = keep + __shfl_xor_sync(full, send, 2);
    }
    z += __shfl_xor_sync(full, z, 1);
    return z;
}

// Twelve values (row f4: colour + three extra feature channels): v[0..7] as above, e[0..3] = {v8, extra 0..2} ride a
// second, shallower transposing butterfly (4 -> 2 -> 1 over lane bits 4, 3, a plain step over bit 2) and are folded
// with the first in the bit-1 step: 13 shuffles.  On return lanes with (lane & 3) == 0 hold v[lane >> 2] and lanes
// with (lane & 7) == 2 hold e[lane >> 3].
__device__ __forceinline__ float warp_reduce12(const float (&v)[8], const float (&e)[4], bool b4, bool b3, bool b2,
                                               bool b1) {
    const unsigned full = 0xffffffffu;
    float w[4], f[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b4 ? v[i] : v[i + 4];
        const float keep = b4 ? v[i + 4] : v[i];
        w[i] = keep + __shfl_xor_sync(full, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b4 ? e[i] : e[i + 2];
        const float keep = b4 ? e[i + 2] : e[i];
        f[i] = keep + __shfl_xor_sync(full, send, 16);
    }
    float x[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b3 ? w[i] : w[i + 2];
        const float keep = b3 ? w[i + 2] : w[i];
        x[i] = keep + __shfl_xor_sync(full, send, 8);
    }
    float h;
    {
        const float send = b3 ? f[0] : f[1];
        const float keep = b3 ? f[1] : f[0];
        h = keep + __shfl_xor_sync(full, send, 8);
    }
    float y;
    {
        const float send = b2 ? x[0] : x[1];
        const float keep = b2 ? x[1] : x[0];
        y = keep + __shfl_xor_sync(full, send, 4);
    }
    h += __shfl_xor_sync(full, h, 4);
    float z;
    {
        const float send = b1 ? y : h;
        const float keep = b1 ? h : y;
        z = keep + __shfl_xor_sync(full, send, 2);
    }
    z += __shfl_xor_sync(full, z, 1);
    return z;
}

struct __align__(16) WarpSlabBX {
    float e0[32], e1[32], e2[32];
};

template <bool kExtra>
__global__ void __launch_bounds__(256, kExtra ? 3 : 4)
render_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                  const SplatRec* __restrict__ rec, int W, int H, int tiles_x,
                  const float* __restrict__ bg, const float* __restrict__ final_T,
                  const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                  float* __restrict__ acc, const int32_t* __restrict__ status, const ExtraArgs ex) {
    __shared__ WarpSlabB slabs[kWarpsPerTile];
    __shared__ WarpSlabBX slabs_x[kExtra ? kWarpsPerTile : 1];
    if (status[FB200_ST_OVERFLOW]) return;

    const unsigned full = 0xffffffffu;
    const int tile = blockIdx.x;
    const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
    const int warp = threadIdx.x >> 5;
    int lane;
    // volatile: keeps the lane id in a register (ptxas otherwise re-reads SR_TID.X inside the hot loop)
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    const bool lb4 = lane & 16, lb3 = lane & 8, lb2 = lane & 4, lb1 = lane & 2;
    // kExtra: lanes 2, 10, 18, 26 carry slots 8..11 (colour b, extras); otherwise lane 2 carries slot 8
    const bool red_lane = ((lane & 3) == 0) || (kExtra ? (lane & 7) == 2 : lane == 2);
    const int red_slot = (lane & 2) ? 8 + (lane >> 3) : (lane >> 2);
    WarpSlabB& slab = slabs[warp];
    WarpSlabBX& slabx = slabs_x[kExtra ? warp : 0];
    const int sub_x0 = tile_x * kTile + (warp & 1) * kSubW;
    const int sub_y0 = tile_y * kTile + (warp >> 1) * kSubH;
    const int pix_x = sub_x0 + (lane & 7), pix_y = sub_y0 + (lane >> 3);
    const bool inside = pix_x < W && pix_y < H;
    const float pxf = (float)pix_x, pyf = (float)pix_y;
    const float lox = (float)sub_x0, hix = (float)(sub_x0 + kSubW - 1);
    const float loy = (float)sub_y0, hiy = (float)(sub_y0 + kSubH - 1);
    const size_t pix_id = (size_t)pix_y * W + pix_x;
    const size_t HW = (size_t)H * W;

    const uint2 range = ranges[tile];

    const float T_final = inside ? final_T[pix_id] : 0.f;
    const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0u;
    float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
    if (inside) {
        dp0 = dL_dpix[pix_id];
        dp1 = dL_dpix[HW + pix_id];
        dp2 = dL_dpix[2 * HW + pix_id];
    }
    float bg_dot_dpixel = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;
    float de0 = 0.f, de1 = 0.f, de2 = 0.f;      // dL/d(extra output) at this pixel
    if (kExtra) {
        if (inside) {
            de0 = ex.dL_dout[pix_id];
            if (ex.ch > 1) de1 = ex.dL_dout[HW + pix_id];
            if (ex.ch > 2) de2 = ex.dL_dout[2 * HW + pix_id];
        }
        bg_dot_dpixel += ex.bg[0] * de0;
        if (ex.ch > 1) bg_dot_dpixel += ex.bg[1] * de1;
        if (ex.ch > 2) bg_dot_dpixel += ex.bg[2] * de2;
    }
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    // positions [0, n) of the tile list matter to this warp
    const int n = (int)__reduce_max_sync(full, last_contributor);
    if (n == 0) return;

    float T = T_final;
    float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f;      // accum_rec
    float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;      // last_color
    float last_alpha = 0.f;
    float ae0 = 0.f, ae1 = 0.f, ae2 = 0.f;      // accum_rec / last value of the extra channels
    float le0 = 0.f, le1 = 0.f, le2 = 0.f;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;         // extra features of the record in r0..r2
    auto load_extra = [&](uint32_t id) {
        const float* f = ex.feat + (size_t)id * ex.ch;
        x0 = __ldg(f);
        x1 = ex.ch > 1 ? __ldg(f + 1) : 0.f;
        x2 = ex.ch > 2 ? __ldg(f + 2) : 0.f;
    };

    // software pipeline: step s handles positions p = n-1-(s*32+lane), descending
    uint32_t id_cur = 0, id_next = 0;
    float4 r0, r1, r2;
    r0 = r1 = r2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < n) {
        id_cur = point_list[range.x + (n - 1 - lane)];
        const float4* p = reinterpret_cast<const float4*>(rec + id_cur);
        r0 = __ldg(p); r1 = __ldg(p + 1); r2 = __ldg(p + 2);
        if (kExtra) load_extra(id_cur);
    }
    if (32 + lane < n) id_next = point_list[range.x + (n - 1 - 32 - lane)];

    for (int base = 0; base < n; base += 32) {
        const bool hit = (base + lane < n) && overlaps(lox, hix, r0.x, r2.y) && overlaps(loy, hiy, r0.y, r2.z);
        uint32_t bits = __ballot_sync(full, hit);
        if (hit) {
            slab.q0[lane] = r0;
            slab.q1[lane] = r1;
            slab.cb[lane] = r2.x;
            slab.id[lane] = id_cur;
            if (kExtra) { slabx.e0[lane] = x0; slabx.e1[lane] = x1; slabx.e2[lane] = x2; }
        }
        id_cur = id_next;
        if (base + 32 + lane < n) {
            const float4* p = reinterpret_cast<const float4*>(rec + id_cur);
            r0 = __ldg(p); r1 = __ldg(p + 1); r2 = __ldg(p + 2);
            if (kExtra) load_extra(id_cur);
        }
        if (base + 64 + lane < n) id_next = point_list[range.x + (n - 1 - base - 64 - lane)];
        __syncwarp();

        while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1;
            const uint32_t pos = (uint32_t)(n - 1 - base - j);
            const float4 q0 = slab.q0[j];
            const float4 q1 = slab.q1[j];
            const float dx = fadd(q0.x, -pxf), dy = fadd(q0.y, -pyf);
            const float q = ffma(dx, fmul(dx, q0.z), fmul(dy, fmul(dy, q1.x)));
            const float u = fmul(dy, fmul(dx, q0.w));
            const float power = ffma(q, -0.5f, -u);
            const float G = expf(power);
            const float alpha = fminf(0.99f, fmul(q1.y, G));
            const bool active = (pos < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            if (!__any_sync(full, active)) continue;

            float v[8];
            float v8 = 0.f;
            float ve[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
            if (active) {
                const float cb = slab.cb[j];
                const float inv = __frcp_rn(1.f - alpha);    // T/(1-a) and T_final/(1-a) share one reciprocal
                T = T * inv;
                const float dchannel_dcolor = alpha * T;
                ar0 = last_alpha * lc0 + (1.f - last_alpha) * ar0;
                ar1 = last_alpha * lc1 + (1.f - last_alpha) * ar1;
                ar2 = last_alpha * lc2 + (1.f - last_alpha) * ar2;
                lc0 = q1.z; lc1 = q1.w; lc2 = cb;
                float dL_dalpha = (q1.z - ar0) * dp0 + (q1.w - ar1) * dp1 + (cb - ar2) * dp2;
                if (kExtra) {
                    const float f0 = slabx.e0[j], f1 = slabx.e1[j], f2 = slabx.e2[j];
                    ae0 = last_alpha * le0 + (1.f - last_alpha) * ae0;
                    ae1 = last_alpha * le1 + (1.f - last_alpha) * ae1;
                    ae2 = last_alpha * le2 + (1.f - last_alpha) * ae2;
                    le0 = f0; le1 = f1; le2 = f2;
                    dL_dalpha += (f0 - ae0) * de0 + (f1 - ae1) * de1 + (f2 - ae2) * de2;
                    ve[0] = dchannel_dcolor * de0;
                    ve[1] = dchannel_dcolor * de1;
                    ve[2] = dchannel_dcolor * de2;
                }
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final * inv) * bg_dot_dpixel;
                const float dL_dG = q1.y * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * q0.z - gdy * q0.w;
                const float dG_ddely = -gdy * q1.x - gdx * q0.w;
                v[0] = dL_dG * dG_ddelx * ddelx_dx;       // dL/dmean2D.x
                v[1] = dL_dG * dG_ddely * ddely_dy;       // dL/dmean2D.y
                v[2] = -0.5f * gdx * dx * dL_dG;          // dL/dconic.x
                v[3] = -0.5f * gdx * dy * dL_dG;          // dL/dconic.y
                v[4] = -0.5f * gdy * dy * dL_dG;          // dL/dconic.w
                v[5] = G * dL_dalpha;                     // dL/dopacity
                v[6] = dchannel_dcolor * dp0;             // dL/dcolour
                v[7] = dchannel_dcolor * dp1;
                v8 = dchannel_dcolor * dp2;
            }
            float red;
            if (kExtra) {
                const float e4[4] = {v8, ve[0], ve[1], ve[2]};
                red = warp_reduce12(v, e4, lb4, lb3, lb2, lb1);
            } else {
                red = warp_reduce9(v, v8, lb4, lb3, lb2, lb1);
            }
            // lane 4k holds value k (k = 0..7), lane 2 holds value 8 (kExtra: lanes 2, 10, 18, 26 hold 8..11)
            if (red_lane) atomicAdd(acc + (size_t)slab.id[j] * 12 + red_slot, red);
        }
        __syncwarp();   // slab is rewritten by the next step
    }
}

}  // namespace

cudaError_t launch_render_bwd_clear(const BwdArgs& a, cudaStream_t s) {
    return cudaMemsetAsync(a.acc, 0, sizeof(float) * 12 * (size_t)a.prm.P, s);
}

cudaError_t launch_render_bwd(const BwdArgs& a, cudaStream_t s) {
    const int T = a.tiles_x * a.tiles_y;
    count_launch();
    if (a.ex.ch > 0)
        render_bwd_kernel<true><<<T, 256, 0, s>>>(a.ranges, a.point_list, a.rec, a.prm.image_width,
                                                  a.prm.image_height, a.tiles_x, a.in.d_background, a.final_T,
                                                  a.n_contrib, a.dL_dpix, a.acc, a.status, a.ex);
    else
        render_bwd_kernel<false><<<T, 256, 0, s>>>(a.ranges, a.point_list, a.rec, a.prm.image_width,
                                                   a.prm.image_height, a.tiles_x, a.in.d_background, a.final_T,
                                                   a.n_contrib, a.dL_dpix, a.acc, a.status, a.ex);
    return cudaGetLastError();
}

namespace {
// dL/d(extra features): slots 9..11 of the accumulator record, zero for Gaussians that were not rendered
__global__ void __launch_bounds__(256)
extra_grad_kernel(int P, int ch, const int32_t* __restrict__ radii, const float* __restrict__ acc,
                  float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const bool vis = radii[i] > 0;
    for (int c = 0; c < ch; ++c) out[(size_t)i * ch + c] = vis ? acc[(size_t)i * 12 + 9 + c] : 0.f;
}
}  // namespace

cudaError_t launch_extra_grad(const BwdArgs& a, cudaStream_t s) {
    if (a.ex.ch <= 0 || a.prm.P <= 0) return cudaSuccess;
    extra_grad_kernel<<<(a.prm.P + 255) / 256, 256, 0, s>>>(a.prm.P, a.ex.ch, a.radii, a.acc, a.ex.dL_dfeat);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fb200
