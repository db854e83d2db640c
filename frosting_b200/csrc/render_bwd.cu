// render_bwd.cu -- backward of the alpha compositing, warp-autonomous, two Gaussians per lane-iteration on packed fp32.
//
// Replaces renderCUDA<3> backward (DGR/cuda_rasterizer/backward.cu:399-557): back-to-front traversal from each
// pixel's last contributor, T recovered by division, gradients w.r.t. the per-Gaussian 2D mean, conic, opacity and
// colour.  Same formulas; B200-first structure:
//   * the reference issues 9 global float atomics per contributing PIXEL-Gaussian pair (its dominant cost).  Here a
//     warp (8x4 pixels) sums its 32 lanes' contributions with a transposing butterfly and issues ONE red.global.add
//     instruction per PAIR of Gaussians (18 active lanes, addresses inside two 48-byte accumulator records);
//   * every warp walks the list on its own, 32 instances per step, starting at ITS deepest last contributor; no CTA
//     barrier (v1 staged 256-instance batches cooperatively and was dominated by barrier stalls,
//     profiles/r01_render_c3_v1_summary.json).  Records re-read by the warps of a tile are L1/L2 hits;
//   * the same conservative alpha>=1/255 extents as the forward kernel let a warp skip instances that cannot touch
//     its sub-tile;
//   * instruction-issue bound, hence packed fp32 on pairs of hits (kernel comment below): v2 (scalar, one hit at a
//     time) executed 469 M warp-instructions per C3 launch, this one 372 M (profiles/r01_render_c3_v3_summary.json).
#include "common.cuh"

namespace fb200 {

namespace {

__device__ __forceinline__ bool overlaps(float lo, float hi, float c, float ext) {
    return !(c + ext < lo) && !(c - ext > hi);
}

// ---- two Gaussians per lane-iteration with packed fp32 ---------------------------------------------------------------
// The per-warp hit records of a step are COMPACTED into a structure-of-arrays slab (slot = rank of the lane among the
// hits), so hits 2k and 2k+1 sit in adjacent words and one 64-bit shared load yields an aligned register pair; the
// alpha evaluation (libdevice-exact exp included) and the gradient algebra then run as FFMA2/FMUL2/FADD2 on
// (Gaussian 2k, Gaussian 2k+1).  Only the transmittance / accum_rec recurrences, which chain through the two
// Gaussians at a pixel, stay scalar.  The 18 partial gradients (2 x 9) of the pair are summed over the warp by ONE
// transposing butterfly -- 16 values fold 16->8->4->2->1 over lane bits 4..1, the two blue-channel values ride along --
// in 20 shuffles, and ONE red.global.add instruction (18 lanes) retires both Gaussians.
struct __align__(16) PairSlab {
    float x[34], y[34], A[34], B[34], C[34], op[34], r[34], g[34], b[34];
    uint32_t id[34], pos[34];
};
struct __align__(16) PairSlabX {
    float e0[34], e1[34], e2[34];
};

__device__ __forceinline__ P2 ldp(const float* a, int k) { return *reinterpret_cast<const float2*>(a + k); }

// 80 registers / 24 resident warps per SM: measured against 72 / 28 and 64 / 32 (both 5 % slower: more instructions, and
// the kernel is issue-bound, not latency-bound)
template <bool kExtra, int kWarps>
__global__ void __launch_bounds__(32 * kWarps, (kExtra ? 16 : 24) / kWarps)
render_bwd_pair_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                       const SplatRec* __restrict__ rec, int W, int H, int tiles_x,
                       const float* __restrict__ bg, const float* __restrict__ final_T,
                       const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                       float* __restrict__ acc, const int32_t* __restrict__ status, const ExtraArgs ex,
                       const uint32_t* __restrict__ tile_order) {
    __shared__ PairSlab slabs[kWarps];
    __shared__ PairSlabX slabs_x[kExtra ? kWarps : 1];
    if (status[FB200_ST_OVERFLOW]) return;

    const unsigned full = 0xffffffffu;
    constexpr int kSplit = kWarpsPerTile / kWarps;
    const int tile = (int)tile_order[blockIdx.x / kSplit];       // longest lists first (binning.cu)
    const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
    const int wslot = threadIdx.x >> 5;
    const int warp = (blockIdx.x % kSplit) * kWarps + wslot;      // sub-tile index inside the tile
    int lane;
    unsigned lt_mask;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    const bool lb4 = lane & 16, lb3 = lane & 8, lb2 = lane & 4, lb1 = lane & 2, lb0 = lane & 1;
    // after the butterfly: even lanes hold value (lane >> 1) & 7 of Gaussian (lane >> 4); odd lanes hold the ride-along
    // values: slot 8 (kExtra: 8 + ((lane >> 2) & 3)) of Gaussian (lane >> 4)
    const bool red_lane = !lb0 || (kExtra ? (lane & 3) == 1 : (lane & 15) == 1);
    const int red_slot = lb0 ? (kExtra ? 8 + ((lane >> 2) & 3) : 8) : ((lane >> 1) & 7);
    PairSlab& slab = slabs[wslot];
    PairSlabX& slabx = slabs_x[kExtra ? wslot : 0];
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
    float de0 = 0.f, de1 = 0.f, de2 = 0.f;
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

    const int n = (int)__reduce_max_sync(full, last_contributor);
    if (n == 0) return;

    float T = T_final;
    float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f;      // accum_rec (advanced form, see below)
    float ae0 = 0.f, ae1 = 0.f, ae2 = 0.f;      // same for the extra channels
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    auto load_extra = [&](uint32_t id) {
        const float* f = ex.feat + (size_t)id * ex.ch;
        x0 = __ldg(f);
        x1 = ex.ch > 1 ? __ldg(f + 1) : 0.f;
        x2 = ex.ch > 2 ? __ldg(f + 2) : 0.f;
    };

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
        const bool hit = (base + lane < n) && subtile_hit(lox, hix, loy, hiy, r0.x, r0.y, r0.z, r0.w, r1.x, r2.y, r2.z, r2.w);
        const uint32_t bits = __ballot_sync(full, hit);
        const int nhit = __popc(bits);
        if (hit) {
            const int slot = __popc(bits & lt_mask);
            slab.x[slot] = r0.x; slab.y[slot] = r0.y; slab.A[slot] = r0.z; slab.B[slot] = r0.w;
            slab.C[slot] = r1.x; slab.op[slot] = r1.y; slab.r[slot] = r1.z; slab.g[slot] = r1.w; slab.b[slot] = r2.x;
            slab.id[slot] = id_cur;
            slab.pos[slot] = (uint32_t)(n - 1 - base - lane);
            if (kExtra) { slabx.e0[slot] = x0; slabx.e1[slot] = x1; slabx.e2[slot] = x2; }
        }
        if ((nhit & 1) && lane == 0) {
            // odd count: pad with a record that can never contribute (opacity 0, position beyond every pixel's last)
            slab.x[nhit] = 0.f; slab.y[nhit] = 0.f; slab.A[nhit] = 0.f; slab.B[nhit] = 0.f; slab.C[nhit] = 0.f;
            slab.op[nhit] = 0.f; slab.r[nhit] = 0.f; slab.g[nhit] = 0.f; slab.b[nhit] = 0.f;
            slab.id[nhit] = 0u; slab.pos[nhit] = 0xffffffffu;
            if (kExtra) { slabx.e0[nhit] = 0.f; slabx.e1[nhit] = 0.f; slabx.e2[nhit] = 0.f; }
        }
        id_cur = id_next;
        if (base + 32 + lane < n) {
            const float4* p = reinterpret_cast<const float4*>(rec + id_cur);
            r0 = __ldg(p); r1 = __ldg(p + 1); r2 = __ldg(p + 2);
            if (kExtra) load_extra(id_cur);
        }
        if (base + 64 + lane < n) id_next = point_list[range.x + (n - 1 - base - 64 - lane)];
        __syncwarp();

        for (int k = 0; k < nhit; k += 2) {
            const P2 X = ldp(slab.x, k), Y = ldp(slab.y, k), A = ldp(slab.A, k), B = ldp(slab.B, k), Cc = ldp(slab.C, k);
            const P2 OP = ldp(slab.op, k);
            const uint2 POS = *reinterpret_cast<const uint2*>(slab.pos + k);
            // power = -0.5*(A dx^2 + C dy^2) - B dx dy in the reference's op order (same as the forward kernel)
            const P2 dx = add2(X, bc(-pxf)), dy = add2(Y, bc(-pyf));
            const P2 q = fma2(dx, mul2(dx, A), mul2(dy, mul2(dy, Cc)));
            const P2 u = mul2(dy, mul2(dx, B));
            const P2 power = fma2(q, bc(-0.5f), neg2(u));
            const P2 G = exp_pair(power);
            const P2 og = mul2(OP, G);
            const float al0 = fminf(0.99f, og.x), al1 = fminf(0.99f, og.y);
            const bool act0 = (POS.x < last_contributor) && !(power.x > 0.0f) && !(al0 < 1.0f / 255.0f);
            const bool act1 = (POS.y < last_contributor) && !(power.y > 0.0f) && !(al1 < 1.0f / 255.0f);
            if (!__any_sync(full, act0 || act1)) continue;

            const P2 Rc = ldp(slab.r, k), Gc = ldp(slab.g, k), Bc = ldp(slab.b, k);
            // 1/(1 - alpha) for both: alpha <= 0.99 keeps the argument in [0.01, 1], so MUFU.RCP + one packed Newton
            // step (|error| < 1 ulp) needs none of __frcp_rn's range checks
            const P2 om = add2(bc(1.0f), neg2(p2(al0, al1)));                  // 1 - alpha
            P2 INV;
            {
                float i0, i1;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i0) : "f"(om.x));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i1) : "f"(om.y));
                const P2 r = p2(i0, i1);
                INV = fma2(r, fma2(neg2(om), r, bc(1.0f)), r);
            }
            // ---- the recurrences chain through the two Gaussians: scalar, first 2k then 2k+1 ----
            // accum_rec is kept in its "already advanced" form: the reference advances it with the PREVIOUS
            // contributor's (alpha, colour) just before use (backward.cu:509-512); advancing it with the CURRENT
            // one right after use is the same arithmetic one step earlier and needs no last_alpha / last_color
            P2 AR0, AR1, AR2, TP, AE0, AE1, AE2;
            AR0.x = ar0; AR1.x = ar1; AR2.x = ar2;
            if (kExtra) { AE0.x = ae0; AE1.x = ae1; AE2.x = ae2; }
            if (act0) {
                T = T * INV.x;
                ar0 = al0 * Rc.x + om.x * ar0;
                ar1 = al0 * Gc.x + om.x * ar1;
                ar2 = al0 * Bc.x + om.x * ar2;
                if (kExtra) {
                    ae0 = al0 * slabx.e0[k] + om.x * ae0;
                    ae1 = al0 * slabx.e1[k] + om.x * ae1;
                    ae2 = al0 * slabx.e2[k] + om.x * ae2;
                }
            }
            TP.x = T;
            AR0.y = ar0; AR1.y = ar1; AR2.y = ar2;
            if (kExtra) { AE0.y = ae0; AE1.y = ae1; AE2.y = ae2; }
            if (act1) {
                T = T * INV.y;
                ar0 = al1 * Rc.y + om.y * ar0;
                ar1 = al1 * Gc.y + om.y * ar1;
                ar2 = al1 * Bc.y + om.y * ar2;
                if (kExtra) {
                    ae0 = al1 * slabx.e0[k + 1] + om.y * ae0;
                    ae1 = al1 * slabx.e1[k + 1] + om.y * ae1;
                    ae2 = al1 * slabx.e2[k + 1] + om.y * ae2;
                }
            }
            TP.y = T;

            // ---- gradient algebra on the pair; a non-contributing (lane, Gaussian) is zeroed through alpha and G ----
            const P2 al = p2(act0 ? al0 : 0.f, act1 ? al1 : 0.f);
            const P2 Gm = p2(act0 ? G.x : 0.f, act1 ? G.y : 0.f);
            const P2 dch = mul2(al, TP);                                       // d(channel)/d(colour)
            P2 dLa = fma2(add2(Rc, neg2(AR0)), bc(dp0),
                          fma2(add2(Gc, neg2(AR1)), bc(dp1), mul2(add2(Bc, neg2(AR2)), bc(dp2))));
            P2 V9, V10, V11;
            if (kExtra) {
                const P2 F0 = ldp(slabx.e0, k), F1 = ldp(slabx.e1, k), F2 = ldp(slabx.e2, k);
                dLa = fma2(add2(F0, neg2(AE0)), bc(de0), dLa);
                dLa = fma2(add2(F1, neg2(AE1)), bc(de1), dLa);
                dLa = fma2(add2(F2, neg2(AE2)), bc(de2), dLa);
                V9 = mul2(dch, bc(de0)); V10 = mul2(dch, bc(de1)); V11 = mul2(dch, bc(de2));
            }
            dLa = mul2(dLa, TP);
            dLa = fma2(mul2(bc(-T_final), INV), bc(bg_dot_dpixel), dLa);
            const P2 dL_dG = mul2(OP, dLa);
            const P2 gdx = mul2(Gm, dx), gdy = mul2(Gm, dy);
            const P2 dG_ddelx = fma2(neg2(gdx), A, neg2(mul2(gdy, B)));
            const P2 dG_ddely = fma2(neg2(gdy), Cc, neg2(mul2(gdx, B)));
            const P2 hh = mul2(dL_dG, bc(-0.5f));
            const P2 hx = mul2(hh, gdx), hy = mul2(hh, gdy);
            P2 V[8];
            V[0] = mul2(mul2(dL_dG, dG_ddelx), bc(ddelx_dx));     // dL/dmean2D.x
            V[1] = mul2(mul2(dL_dG, dG_ddely), bc(ddely_dy));     // dL/dmean2D.y
            V[2] = mul2(hx, dx);                                  // dL/dconic.x
            V[3] = mul2(hx, dy);                                  // dL/dconic.y
            V[4] = mul2(hy, dy);                                  // dL/dconic.w
            V[5] = mul2(Gm, dLa);                                 // dL/dopacity
            V[6] = mul2(dch, bc(dp0));                            // dL/dcolour
            V[7] = mul2(dch, bc(dp1));
            const P2 V8 = mul2(dch, bc(dp2));

            // ---- one butterfly for both Gaussians ----
            // bit 4: lanes 0-15 keep Gaussian 2k (.x), lanes 16-31 keep Gaussian 2k+1 (.y)
            float w[8];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const P2 send = p2(lb4 ? V[i].x : V[i].y, lb4 ? V[i + 1].x : V[i + 1].y);
                const P2 keep = p2(lb4 ? V[i].y : V[i].x, lb4 ? V[i + 1].y : V[i + 1].x);
                const P2 sum = add2(keep, p2(__shfl_xor_sync(full, send.x, 16), __shfl_xor_sync(full, send.y, 16)));
                w[i] = sum.x; w[i + 1] = sum.y;
            }
            float xx[4];
#pragma unroll
            for (int i = 0; i < 4; i += 2) {
                const P2 send = p2(lb3 ? w[i] : w[i + 4], lb3 ? w[i + 1] : w[i + 5]);
                const P2 keep = p2(lb3 ? w[i + 4] : w[i], lb3 ? w[i + 5] : w[i + 1]);
                const P2 sum = add2(keep, p2(__shfl_xor_sync(full, send.x, 8), __shfl_xor_sync(full, send.y, 8)));
                xx[i] = sum.x; xx[i + 1] = sum.y;
            }
            float yy[2];
            {
                const P2 send = p2(lb2 ? xx[0] : xx[2], lb2 ? xx[1] : xx[3]);
                const P2 keep = p2(lb2 ? xx[2] : xx[0], lb2 ? xx[3] : xx[1]);
                const P2 sum = add2(keep, p2(__shfl_xor_sync(full, send.x, 4), __shfl_xor_sync(full, send.y, 4)));
                yy[0] = sum.x; yy[1] = sum.y;
            }
            float z;
            {
                const float send = lb1 ? yy[0] : yy[1];
                const float keep = lb1 ? yy[1] : yy[0];
                z = keep + __shfl_xor_sync(full, send, 2);       // value 4*b3 + 2*b2 + b1 of Gaussian b4
            }
            // ride-along values
            float h;
            if (kExtra) {
                const P2 L[4] = {V8, V9, V10, V11};
                float f[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float send = lb4 ? L[i].x : L[i].y;
                    const float keep = lb4 ? L[i].y : L[i].x;
                    f[i] = keep + __shfl_xor_sync(full, send, 16);
                }
                float gq[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float send = lb3 ? f[i] : f[i + 2];
                    const float keep = lb3 ? f[i + 2] : f[i];
                    gq[i] = keep + __shfl_xor_sync(full, send, 8);
                }
                {
                    const float send = lb2 ? gq[0] : gq[1];
                    const float keep = lb2 ? gq[1] : gq[0];
                    h = keep + __shfl_xor_sync(full, send, 4);   // ride-along 2*b3 + b2 of Gaussian b4
                }
                h += __shfl_xor_sync(full, h, 2);
            } else {
                const float send = lb4 ? V8.x : V8.y;
                const float keep = lb4 ? V8.y : V8.x;
                h = keep + __shfl_xor_sync(full, send, 16);
                h += __shfl_xor_sync(full, h, 8);
                h += __shfl_xor_sync(full, h, 4);
                h += __shfl_xor_sync(full, h, 2);
            }
            float red;
            {
                const float send = lb0 ? z : h;
                const float keep = lb0 ? h : z;
                red = keep + __shfl_xor_sync(full, send, 1);
            }
            if (red_lane && !(lb4 && k + 1 >= nhit)) {
                const uint32_t gid = slab.id[k + (lb4 ? 1 : 0)];
                atomicAdd(acc + (size_t)gid * 12 + red_slot, red);
            }
        }
        __syncwarp();   // slab is rewritten by the next step
    }
}


// ---- v4: transposed blocks -- lane = Gaussian for the per-pair math, lane = pixel for the recurrences --------------------
// The pair kernel above keeps lane = pixel throughout, so every Gaussian's 9 partial gradients must be summed over the
// warp: 36 % of its pair loop is the transposing butterfly (20 SHFL + 36 FSEL per two Gaussians,
// profiles/r01_pair_loop_sass_budget.json).  Here the hits of a warp's 8x4 sub-tile are processed in blocks of 16
// Gaussians and the work is split by what each part needs:
//   phase 1  lane = (Gaussian g, pixel half h): the lane evaluates ITS Gaussian at its 16 pixels, two pixels at a time on
//            packed fp32 -- same op order as the forward (power / libdevice exp / alpha tests, bit-identical decisions) --
//            and leaves three values per (Gaussian, pixel) in a padded shared-memory matrix: the masked opacity*G, the
//            1/(1-alpha) of the pair and alpha * (colour . dL/dpixel);
//   phase 2  lane = pixel: the only sequential part, the back-to-front recurrences, walks the 16 Gaussians of the block
//            for the lane's own pixel -- T <- T/(1-alpha), E <- E + alpha T (c . dL/dpix) -- two loads, four flops and two
//            stores per pair; the per-pixel state (T, E) lives in registers across blocks.  It leaves T_g and
//            U = (K - E_g)/(1-alpha) in place of the inputs, so that dL/dalpha = T_g (c . dL/dpix) + U
//            (backward.cu:509-534 with accum_rec kept un-normalised: T_g * accum_rec = E_g / (1 - alpha_g));
//   phase 3  lane = (g, h) again: gradient algebra for its Gaussian, accumulated over its 16 pixels IN REGISTERS as five
//            image moments of S = dL/dG * G (sum S dx, S dy, S dx^2, S dx dy, S dy^2), sum S / opacity-side and the three
//            colour sums; no cross-lane reduction except one xor-16 add per block.  9 red.global.add per Gaussian, as before.
// No shuffle and no select in the inner loops; the matrix rows are padded to 34 floats so that the 64-bit row accesses of
// phases 1/3 (lane stride = one row) and the 32-bit column accesses of phase 2 (lane stride = one word) are both
// bank-conflict free.
constexpr int kBlk = 16;            // Gaussians per block
constexpr int kRow = 34;            // padded row length (floats) of the [kBlk][32 pixels] matrices

struct __align__(16) BwdWarpSmem {
    float m0[kBlk * kRow];          // masked opacity * G                       (phase 1 -> phase 3)
    float m1[kBlk * kRow];          // 1 / (1 - alpha)  -> T_g                  (phase 1 -> 2 -> 3)
    float m2[kBlk * kRow];          // alpha * (c . dL/dpix)  -> U              (phase 1 -> 2 -> 3)
    float dp0[32], dp1[32], dp2[32], K[32];     // per pixel: dL/dpixel, K = -T_final * (bg . dL/dpixel)
    uint32_t last[32];                          // per pixel: entries of the hit list up to its last contributor
};

// The list a warp walks is its sub-tile's HIT LIST, written by the forward blend (render_fwd.cu): the Gaussian ids
// that passed the sub-tile test, in tile-list order, and per pixel the number of entries up to its last contributor.
// The backward therefore touches no instance the forward did not blend into this sub-tile: no re-test, no walk over
// the other 7/8 of the tile list, and the per-pixel "before my last contributor" test is a compare of entry numbers.
template <int kWarps, int kResidentWarps>
__global__ void __launch_bounds__(32 * kWarps, kResidentWarps / kWarps)
render_bwd_t16_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ sub_hits,
                      const SplatRec* __restrict__ rec, int W, int H, int tiles_x,
                      const float* __restrict__ bg, const float* __restrict__ final_T,
                      const uint32_t* __restrict__ last_entry, const float* __restrict__ dL_dpix,
                      float* __restrict__ acc, const int32_t* __restrict__ status,
                      const uint32_t* __restrict__ tile_order) {
    __shared__ BwdWarpSmem smem[kWarps];
    if (status[FB200_ST_OVERFLOW]) return;

    const unsigned full = 0xffffffffu;
    constexpr int kSplit = kWarpsPerTile / kWarps;
    const int tile = (int)tile_order[blockIdx.x / kSplit];       // longest lists first (binning.cu)
    const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
    const int wslot = threadIdx.x >> 5;
    const int warp = (blockIdx.x % kSplit) * kWarps + wslot;      // sub-tile index inside the tile
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    BwdWarpSmem& sm = smem[wslot];
    const int sub_x0 = tile_x * kTile + (warp & 1) * kSubW;
    const int sub_y0 = tile_y * kTile + (warp >> 1) * kSubH;
    const uint2 range = ranges[tile];
    const uint32_t* const sub = sub_hits + (size_t)kWarpsPerTile * range.x + (size_t)warp * (range.y - range.x);

    // ---- lane = pixel: the lane's own pixel, its recurrence state and its row of the per-pixel constants ----
    float T, E = 0.f, Kp;
    int top;                        // entries [0, top) of the hit list are still to be processed
    {
        const int pix_x = sub_x0 + (lane & 7), pix_y = sub_y0 + (lane >> 3);
        const bool inside = pix_x < W && pix_y < H;
        const size_t pix_id = (size_t)pix_y * W + pix_x;
        const size_t HW = (size_t)H * W;
        const float T_final = inside ? final_T[pix_id] : 0.f;
        const uint32_t le = inside ? last_entry[pix_id] : 0u;
        float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
        if (inside) {
            dp0 = dL_dpix[pix_id];
            dp1 = dL_dpix[HW + pix_id];
            dp2 = dL_dpix[2 * HW + pix_id];
        }
        Kp = -T_final * (bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2);
        T = T_final;
        sm.dp0[lane] = dp0; sm.dp1[lane] = dp1; sm.dp2[lane] = dp2; sm.K[lane] = Kp; sm.last[lane] = le;
        top = (int)__reduce_max_sync(full, le);
    }
    if (top == 0) return;
    __syncwarp();
    // Pixel-pair slot k (0..7) of phases 1 / 3 covers pixels {2k, 2k+1} of the upper half and {16+2k, 16+2k+1} of the
    // lower half at once.  pair_last[k] = the largest "last entry" among those four pixels: a block whose entries all lie
    // at or beyond it has nothing to do in slot k.  Deep in a list only the few pixels that never saturated are still
    // live, so most slots are skipped there (C2 / C5: most of the walk).
    uint32_t pair_last;
    {
        const int k = lane & 7;
        pair_last = max(max(sm.last[2 * k], sm.last[2 * k + 1]), max(sm.last[16 + 2 * k], sm.last[17 + 2 * k]));
    }

    // ---- lane = (Gaussian g, pixel half h): rows 2h, 2h+1 of the 8x4 sub-tile ----
    const int g = lane & (kBlk - 1), h = lane >> 4;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    const float px0 = (float)sub_x0;
    const float py0 = (float)(sub_y0 + 2 * h);
    float* const row0 = sm.m0 + g * kRow + 16 * h;
    float* const row1 = sm.m1 + g * kRow + 16 * h;
    float* const row2 = sm.m2 + g * kRow + 16 * h;
    const float* const pdp0 = sm.dp0 + 16 * h;
    const float* const pdp1 = sm.dp1 + 16 * h;
    const float* const pdp2 = sm.dp2 + 16 * h;
    const uint32_t* const plast = sm.last + 16 * h;

    // Register prefetch, two stages deep: the Gaussian ids of block k+2 and the records of block k+1 are requested while
    // block k is processed, so neither hop of the dependent chain id -> record is waited for (round-2 v6 profile at C2:
    // 27 % of the stall samples sat on the id load at the top of every block).  A/B (profiles/r02_tma_ab.md): staging
    // the records with one 48-byte cp.async.bulk each into a 4-deep shared-memory ring, 3 blocks ahead, was SLOWER
    // (C3: 0.408 vs 0.337 ms) -- 16 small bulk copies per block per warp exceed what the SM's TMA unit issues and the
    // warps spin on the mbarriers.
    auto load_id = [&](int t) -> uint32_t {       // entry t-1-g of the list, 0 if beyond its start
        const int e = t - 1 - g;
        return e >= 0 ? __ldg(sub + e) : 0u;
    };
    uint32_t nid = load_id(top), nnid = load_id(top - kBlk);
    float4 n0, n1, n2;
    n0 = n1 = n2 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto fetch_rec = [&](uint32_t id, int t) {    // records of the block whose top is t
        if (t - 1 - g >= 0) {
            const float4* p = reinterpret_cast<const float4*>(rec + id);
            n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2);
        }
    };
    fetch_rec(nid, top);

    while (top > 0) {
        const int cnt = min(kBlk, top);
        const bool valid = g < cnt;
        const uint32_t entry = (uint32_t)(top - 1 - g);          // this Gaussian's entry number (valid lanes)
        const float X = n0.x, Y = n0.y, A = n0.z, B = n0.w, Cc = n1.x;
        const float op = valid ? n1.y : 0.f;
        const float cr = valid ? n1.z : 0.f, cg = valid ? n1.w : 0.f, cb = valid ? n2.x : 0.f;
        const uint32_t gid = nid;
        top -= cnt;
        nid = nnid;
        fetch_rec(nid, top);                  // next block's records (its ids arrived a block ago)
        nnid = load_id(top - kBlk);           // ids of the block after that
        // ---- phase 1: alpha of (Gaussian g) x (up to 16 pixels), the forward's op order ----
        // slots that can hold a contributing pixel for some entry of this block (entries [top, top + cnt) after the update)
        const unsigned slots = __ballot_sync(full, pair_last > (uint32_t)top) & 0xffu;
        bool any_act = false;
        unsigned my_slots = 0;          // slots where THIS lane found a contributing pixel
        auto phase1 = [&](int r, int c) {
            const float dy = fadd(Y, -(py0 + (float)r));
            const float dyC = fmul(dy, fmul(dy, Cc));
            const int k = 8 * r + 2 * c;                                   // pixel pair (k, k+1) of this half
            const P2 dx = add2(bc(X), p2(-(px0 + (float)(2 * c)), -(px0 + (float)(2 * c + 1))));
            const P2 q = fma2(dx, mul2(dx, bc(A)), bc(dyC));
            const P2 u = mul2(bc(dy), mul2(dx, bc(B)));
            const P2 power = fma2(q, bc(-0.5f), neg2(u));
            const P2 og = mul2(bc(op), exp_pair(power));
            const uint2 lastp = *reinterpret_cast<const uint2*>(plast + k);
            // contributes iff the forward blended it: not beyond the pixel's last contributor, power <= 0,
            // alpha >= 1/255 (min(0.99, og) < 1/255  <=>  og < 1/255)
            const bool a0 = valid && (entry < lastp.x) && !(power.x > 0.0f) && !(og.x < 1.0f / 255.0f);
            const bool a1 = valid && (entry < lastp.y) && !(power.y > 0.0f) && !(og.y < 1.0f / 255.0f);
            any_act |= a0 | a1;
            my_slots |= (a0 | a1) ? (1u << (4 * r + c)) : 0u;
            const P2 ogm = p2(a0 ? og.x : 0.f, a1 ? og.y : 0.f);
            const P2 am = p2(fminf(0.99f, ogm.x), fminf(0.99f, ogm.y));
            const P2 om = add2(bc(1.0f), neg2(am));                        // 1 - alpha in [0.01, 1]
            P2 inv;
            {
                float i0, i1;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i0) : "f"(om.x));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i1) : "f"(om.y));
                const P2 rr = p2(i0, i1);
                inv = fma2(rr, fma2(neg2(om), rr, bc(1.0f)), rr);          // one Newton step
            }
            const P2 cd = fma2(ldp(pdp0, k), bc(cr), fma2(ldp(pdp1, k), bc(cg), mul2(ldp(pdp2, k), bc(cb))));
            *reinterpret_cast<float2*>(row0 + k) = ogm;
            *reinterpret_cast<float2*>(row1 + k) = inv;
            *reinterpret_cast<float2*>(row2 + k) = mul2(am, cd);
        };
        if (slots == 0xffu) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) phase1(r, c);
        } else {
            // skipped slots keep whatever an earlier block left in the matrices: make them identities for phase 2
            for (unsigned todo = ~slots & 0xffu; todo; todo &= todo - 1) {
                const int sl = __ffs(todo) - 1, k = 8 * (sl >> 2) + 2 * (sl & 3);
                *reinterpret_cast<float2*>(row1 + k) = make_float2(1.f, 1.f);
                *reinterpret_cast<float2*>(row2 + k) = make_float2(0.f, 0.f);
            }
            for (unsigned todo = slots; todo; todo &= todo - 1) {
                const int sl = __ffs(todo) - 1;
                phase1(sl >> 2, sl & 3);
            }
        }
        // Gaussians of the block with at least one contributing pixel (either half); the others are identities for the
        // recurrences and add nothing to any sum
        const unsigned act_bits = __ballot_sync(full, any_act);
        const unsigned act16 = (act_bits | (act_bits >> 16)) & 0xffffu;
        if (act16 == 0) continue;
        __syncwarp();                  // phase 1's stores before phase 2's loads
        // ---- phase 2: lane = pixel, back-to-front recurrences over the block's Gaussians ----
        {
            float* c1 = sm.m1 + lane;
            float* c2 = sm.m2 + lane;
            if (__popc(act16) + 2 >= cnt) {
                // dense block: walk every row (inactive rows are identities: inv = 1, bb = 0), loads four deep
#pragma unroll 4
                for (int i = 0; i < cnt; ++i) {
                    const float inv = c1[i * kRow], bb = c2[i * kRow];
                    const float Tg = T * inv;                  // transmittance in front of Gaussian i at this pixel
                    c1[i * kRow] = Tg;
                    c2[i * kRow] = (Kp - E) * inv;             // U: everything behind it (and the background) seen through it
                    E = fmaf(bb, Tg, E);
                    T = Tg;
                }
            } else {
                for (unsigned todo = act16; todo; todo &= todo - 1) {
                    const int i = __ffs(todo) - 1;
                    const float inv = c1[i * kRow], bb = c2[i * kRow];
                    const float Tg = T * inv;
                    c1[i * kRow] = Tg;
                    c2[i * kRow] = (Kp - E) * inv;
                    E = fmaf(bb, Tg, E);
                    T = Tg;
                }
            }
        }
        __syncwarp();
        // ---- phase 3: gradient sums of Gaussian g over its 16 pixels ----
        const float rop = op > 0.f ? __frcp_rn(op) : 0.f;
        P2 sR = bc(0.f), sG = bc(0.f), sB = bc(0.f), sO = bc(0.f);
        P2 M10 = bc(0.f), M01 = bc(0.f), M20 = bc(0.f), M11 = bc(0.f), M02 = bc(0.f);
        auto phase3 = [&](int r, int c) {
            const float dy = fadd(Y, -(py0 + (float)r));
            const int k = 8 * r + 2 * c;
            const P2 dx = add2(bc(X), p2(-(px0 + (float)(2 * c)), -(px0 + (float)(2 * c + 1))));
            const P2 ogm = ldp(row0, k), Tg = ldp(row1, k), U = ldp(row2, k);
            const P2 d0 = ldp(pdp0, k), d1 = ldp(pdp1, k), d2 = ldp(pdp2, k);
            const P2 am = p2(fminf(0.99f, ogm.x), fminf(0.99f, ogm.y));
            const P2 Gm = mul2(ogm, bc(rop));                               // G (0 where the pair does not contribute)
            const P2 cd = fma2(d0, bc(cr), fma2(d1, bc(cg), mul2(d2, bc(cb))));
            const P2 dch = mul2(am, Tg);                                   // d(pixel channel)/d(colour)
            sR = fma2(dch, d0, sR); sG = fma2(dch, d1, sG); sB = fma2(dch, d2, sB);
            const P2 dLa = fma2(Tg, cd, U);                                // dL/dalpha
            const P2 S = mul2(dLa, Gm);                                    // dL/dopacity contribution
            sO = add2(sO, S);
            const P2 SG = mul2(S, bc(op));                                 // dL/dG * G
            const P2 sx = mul2(SG, dx), sy = mul2(SG, bc(dy));
            M10 = add2(M10, sx); M01 = add2(M01, sy);
            M20 = fma2(sx, dx, M20); M11 = fma2(sx, bc(dy), M11); M02 = fma2(sy, bc(dy), M02);
        };
        // only the slots where some lane found a contributing pixel carry anything
        const unsigned live_slots = __reduce_or_sync(full, my_slots);
        if (live_slots == 0xffu) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) phase3(r, c);
        } else {
            for (unsigned todo = live_slots; todo; todo &= todo - 1) {
                const int sl = __ffs(todo) - 1;
                phase3(sl >> 2, sl & 3);
            }
        }
        float v[9] = {M10.x + M10.y, M01.x + M01.y, M20.x + M20.y, M11.x + M11.y, M02.x + M02.y,
                      sO.x + sO.y, sR.x + sR.y, sG.x + sG.y, sB.x + sB.y};
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] += __shfl_xor_sync(full, v[i], 16);
        if (h == 0 && ((act16 >> g) & 1u)) {
            float* a = acc + (size_t)gid * 12;
            // dG/ddelx = -G (dx A + dy B), dG/ddely = -G (dy C + dx B)   (backward.cu:536-546)
            // two 128-bit vector reductions + one scalar (red.global.add.v4.f32, sm_90+): a third of the L2 atomic
            // transactions of nine scalar ones -- neighbouring sub-tiles hit the same 48-byte accumulator records
            atomicAdd(reinterpret_cast<float4*>(a),
                      make_float4(-ddelx_dx * fmaf(A, v[0], B * v[1]), -ddely_dy * fmaf(Cc, v[1], B * v[0]),
                                  -0.5f * v[2], -0.5f * v[3]));
            atomicAdd(reinterpret_cast<float4*>(a) + 1, make_float4(-0.5f * v[4], v[5], v[6], v[7]));
            atomicAdd(a + 8, v[8]);
        }
        __syncwarp();      // the matrices are rewritten by the next block
    }
}

}  // namespace

cudaError_t launch_render_bwd_clear(const BwdArgs& a, cudaStream_t s) {
    return cudaMemsetAsync(a.acc, 0, sizeof(float) * 12 * (size_t)a.prm.P, s);
}

cudaError_t launch_render_bwd(const BwdArgs& a, cudaStream_t s) {
    const int T = a.tiles_x * a.tiles_y;
    constexpr int kWarps = 4;   // measured on C3: 4-warp CTAs (half tiles) beat 8 (-4 %) and 2
    const int grid = T * (kWarpsPerTile / kWarps);
    count_launch();
    if (a.ex.ch > 0)
        render_bwd_pair_kernel<true, kWarps><<<grid, 32 * kWarps, 0, s>>>(
            a.ranges, a.point_list, a.rec, a.prm.image_width, a.prm.image_height, a.tiles_x, a.in.d_background,
            a.final_T, a.n_contrib, a.dL_dpix, a.acc, a.status, a.ex, a.tile_order);
    else if (a.prm.debug & 4)       // A/B switch: the round-1 pair kernel (lane = pixel + butterfly)
        render_bwd_pair_kernel<false, kWarps><<<grid, 32 * kWarps, 0, s>>>(
            a.ranges, a.point_list, a.rec, a.prm.image_width, a.prm.image_height, a.tiles_x, a.in.d_background,
            a.final_T, a.n_contrib, a.dL_dpix, a.acc, a.status, a.ex, a.tile_order);
    else if (a.prm.debug & 16)      // A/B switch: 24 resident warps (80 registers, 13 spilled words) -- measured slower
        render_bwd_t16_kernel<kWarps, 24><<<grid, 32 * kWarps, 0, s>>>(
            a.ranges, a.sub_hits, a.rec, a.prm.image_width, a.prm.image_height, a.tiles_x, a.in.d_background,
            a.final_T, a.last_entry, a.dL_dpix, a.acc, a.status, a.tile_order);
    else                            // 20 resident warps per SM, 96 registers, no spills (C3: 0.336 ms vs 0.362)
        render_bwd_t16_kernel<kWarps, 20><<<grid, 32 * kWarps, 0, s>>>(
            a.ranges, a.sub_hits, a.rec, a.prm.image_width, a.prm.image_height, a.tiles_x, a.in.d_background,
            a.final_T, a.last_entry, a.dL_dpix, a.acc, a.status, a.tile_order);
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
