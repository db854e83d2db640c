// optim.cu -- data-parallel gradient reduction fused with the Adam update.
//
// Replaces, for camera-sharded training, what a torch user would bolt onto the reference's optimizer
// wrapper (frosting_scene/frosting_optimizer.py:101 `torch.optim.Adam(l, lr=0.0, eps=1e-15)`, stepped at
// :116-118): an NCCL all-reduce of every parameter gradient followed by a multi-tensor Adam over the 6-12
// parameter groups.  B200-first structure -- ONE kernel does the collective and the update over NVLink peer
// memory:
//   * every rank keeps the full parameter slab and a full gradient slab in peer-mapped memory (cudaIpc);
//     the Adam moments exist only for the rank's own 1/world shard of the flat parameter space;
//   * rank r's kernel walks its shard: 128-bit loads of the same gradient vector from all `world` gradient
//     slabs (the local one from HBM, the others across NVSwitch), summed in rank order so the result does
//     not depend on which rank owns the shard, scaled (1/world for the mean), Adam update of the local
//     moments, and the new parameter vector stored to all `world` parameter slabs (peer stores);
//   * so the wire carries (world-1)/world of the gradients once (reduce-scatter) and (world-1)/world of the
//     parameters once (all-gather) -- the volume of a ring all-reduce -- but with no staging buffers, no
//     second pass over HBM, and 1/world of Adam's 28 B/element HBM traffic per GPU.
// The two stream-ordered rendezvous this needs (all gradients complete before any shard is read; all
// parameter stores landed before the next forward) are the scalar loss all-reduce the training loop does
// anyway and one more 4-byte all-reduce (frosting_b200/optim.py).  With world = 1 the same kernel is a plain
// fused multi-group Adam.
//
// Arithmetic follows torch.optim.Adam's single-tensor formulas (torch/optim/adam.py `_single_tensor_adam`,
// the path the reference's optimizer takes): m += (g - m)(1 - b1); v = b2 v + (1 - b2) g g;
// p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps), bias corrections computed on the host in double.
#include "common.cuh"

namespace fb200 {

namespace {

__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcg(p); }   // L2 only: peer data is never L1-cached

struct Coef {
    float one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps;
};

__device__ __forceinline__ float adam1(float g, float& m, float& v, float p, float lr_over_bc1, const Coef& c) {
    m = m + (g - m) * c.one_minus_b1;
    v = v * c.b2 + c.one_minus_b2 * g * g;
    const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
    return p - lr_over_bc1 * (m / denom);
}

__device__ __forceinline__ void mc_st(float4* p, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// kPeers: compile-time bound on `world` (1, 2, 4, 8) so that small boxes do not pay registers for eight gradient
// vectors -- more resident warps = more remote loads in flight, which is what hides the NVLink latency.
// kMasked: sparse gradient rows (fb200_adam_args.peer_row_radii): a peer's vector is fetched only if one of its (at
// most four) rows was rendered on that peer, and the elements of its other rows -- unwritten memory -- are dropped.
template <int kPeers, bool kMasked>
__global__ void __launch_bounds__(256, kMasked ? (kPeers <= 2 ? 4 : 2) : (kPeers <= 4 ? 4 : 3))
adam_shard_kernel(const fb200_adam_args a) {
    const int64_t v_lo = a.shard_lo >> 2, v_hi = a.shard_hi >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float4* __restrict__ mom1 = reinterpret_cast<float4*>(a.d_exp_avg);
    float4* __restrict__ mom2 = reinterpret_cast<float4*>(a.d_exp_avg_sq);
    float4* mc_p = reinterpret_cast<float4*>(a.mc_params);
    // torch forms 1 - beta in double and rounds once (python floats); (float)1 - (float)beta would be off by 1e-5 relative
    const Coef c{(float)(1.0 - a.beta1), (float)a.beta2, (float)(1.0 - a.beta2), a.bias_correction2_sqrt, a.eps};
    for (int64_t i = v_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v_hi; i += stride) {
        // group of this vector (group starts are multiples of 4 elements, so a vector never straddles two)
        const int64_t e = i << 2;
        int k = 0;
#pragma unroll
        for (int j = 1; j < FB200_ADAM_MAX_GROUPS; ++j)
            if (j < a.n_groups && e >= a.group_start[j]) k = j;
        const float lr = a.lr[k];
        if (lr < 0.f) continue;        // group without a gradient this step: untouched, like a torch parameter with .grad None
        const float step = lr / a.bias_correction1;

        // the rank's own state first: these loads do not depend on anything below and overlap the (dependent) radii ->
        // gradient chain of the masked path
        const int64_t li = i - v_lo;
        float4 m = mom1[li], v = mom2[li];
        float4 p = reinterpret_cast<const float4*>(a.peer_params[a.rank])[i];

        // gradient: one 128-bit load per peer, all issued before the first use
        float4 g[kPeers];
        const uint32_t w = kMasked ? (uint32_t)a.row_width[k] : 0u;
        if (kMasked && w > 0) {
            const uint32_t off = (uint32_t)(e - a.group_start[k]);
            const uint32_t last = (uint32_t)a.row_count - 1u;      // padding elements behind the last row map onto it
            uint32_t r = off / w, rem = off - r * w;
            uint32_t rows[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                rows[j] = min(r, last);
                if (++rem == w) { rem = 0; ++r; }
            }
            bool live[kPeers][4];
#pragma unroll
            for (int p = 0; p < kPeers; ++p) {
                if (p < a.world) {
                    const int32_t* rr = a.peer_row_radii[p];
                    live[p][0] = __ldcg(rr + rows[0]) > 0;
                    live[p][3] = rows[3] == rows[0] ? live[p][0] : __ldcg(rr + rows[3]) > 0;
                    // four consecutive elements span at most two rows unless rows are shorter than 3 elements
                    live[p][1] = rows[1] == rows[0] ? live[p][0] : (rows[1] == rows[3] ? live[p][3] : __ldcg(rr + rows[1]) > 0);
                    live[p][2] = rows[2] == rows[0] ? live[p][0] : (rows[2] == rows[3] ? live[p][3] : __ldcg(rr + rows[2]) > 0);
                }
            }
#pragma unroll
            for (int p = 0; p < kPeers; ++p) {
                g[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p < a.world && (live[p][0] || live[p][1] || live[p][2] || live[p][3])) {
                    const float4 t = ld_stream(reinterpret_cast<const float4*>(a.peer_grads[p]) + i);
                    g[p] = make_float4(live[p][0] ? t.x : 0.f, live[p][1] ? t.y : 0.f, live[p][2] ? t.z : 0.f, live[p][3] ? t.w : 0.f);
                }
            }
        } else {
#pragma unroll
            for (int p = 0; p < kPeers; ++p)
                if (p < a.world) g[p] = ld_stream(reinterpret_cast<const float4*>(a.peer_grads[p]) + i);
        }
        float4 s = g[0];
#pragma unroll
        for (int p = 1; p < kPeers; ++p)
            if (p < a.world) { s.x += g[p].x; s.y += g[p].y; s.z += g[p].z; s.w += g[p].w; }
        s.x *= a.grad_scale; s.y *= a.grad_scale; s.z *= a.grad_scale; s.w *= a.grad_scale;

        p.x = adam1(s.x, m.x, v.x, p.x, step, c);
        p.y = adam1(s.y, m.y, v.y, p.y, step, c);
        p.z = adam1(s.z, m.z, v.z, p.z, step, c);
        p.w = adam1(s.w, m.w, v.w, p.w, step, c);
        mom1[li] = m;
        mom2[li] = v;
        if (kMasked && mc_p != nullptr) {
            mc_st(mc_p + i, p);           // the switch writes every replica
        } else {
#pragma unroll
            for (int q = 0; q < kPeers; ++q)
                if (q < a.world) __stcg(reinterpret_cast<float4*>(a.peer_params[q]) + i, p);
        }
    }
}

// NVLS variant: the same shard walk, but the gradient vector arrives already summed over all ranks -- one
// multimem.ld_reduce on the multicast mapping of the gradient slabs makes the NVSwitch read every replica and add them --
// and one multimem.st on the multicast mapping of the parameter slabs makes the switch write every replica.  Per GPU the
// wire carries shard bytes in and shard bytes out instead of (world - 1) times that.
__device__ __forceinline__ float4 mc_ld_reduce(const float4* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256, 4)
adam_shard_mc_kernel(const fb200_adam_args a) {
    const int64_t v_lo = a.shard_lo >> 2, v_hi = a.shard_hi >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float4* __restrict__ mom1 = reinterpret_cast<float4*>(a.d_exp_avg);
    float4* __restrict__ mom2 = reinterpret_cast<float4*>(a.d_exp_avg_sq);
    const float4* mc_g = reinterpret_cast<const float4*>(a.mc_grads);
    float4* mc_p = reinterpret_cast<float4*>(a.mc_params);
    const float4* my_p = reinterpret_cast<const float4*>(a.peer_params[a.rank]);
    const Coef c{(float)(1.0 - a.beta1), (float)a.beta2, (float)(1.0 - a.beta2), a.bias_correction2_sqrt, a.eps};
    for (int64_t i = v_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v_hi; i += stride) {
        float4 s = mc_ld_reduce(mc_g + i);
        s.x *= a.grad_scale; s.y *= a.grad_scale; s.z *= a.grad_scale; s.w *= a.grad_scale;
        const int64_t e = i << 2;
        float lr = a.lr[0];
#pragma unroll
        for (int k = 1; k < FB200_ADAM_MAX_GROUPS; ++k)
            if (k < a.n_groups && e >= a.group_start[k]) lr = a.lr[k];
        if (lr < 0.f) continue;
        const float step = lr / a.bias_correction1;
        const int64_t li = i - v_lo;
        float4 m = mom1[li], v = mom2[li];
        float4 p = my_p[i];
        p.x = adam1(s.x, m.x, v.x, p.x, step, c);
        p.y = adam1(s.y, m.y, v.y, p.y, step, c);
        p.z = adam1(s.z, m.z, v.z, p.z, step, c);
        p.w = adam1(s.w, m.w, v.w, p.w, step, c);
        mom1[li] = m;
        mom2[li] = v;
        mc_st(mc_p + i, p);
    }
}

}  // namespace

cudaError_t launch_adam_shard(const fb200_adam_args& a, cudaStream_t s) {
    const int64_t vecs = (a.shard_hi - a.shard_lo) >> 2;
    if (vecs <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (vecs + 255) / 256;
    const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);   // 8 resident CTAs of 256 threads per SM
    if (a.peer_row_radii[0] != nullptr) {
        // sparse gradient rows: masked peer loads (the in-switch sum would read the unwritten rows), multicast store if mapped
        if (a.world == 1) adam_shard_kernel<1, true><<<grid, 256, 0, s>>>(a);
        else if (a.world == 2) adam_shard_kernel<2, true><<<grid, 256, 0, s>>>(a);
        else if (a.world <= 4) adam_shard_kernel<4, true><<<grid, 256, 0, s>>>(a);
        else adam_shard_kernel<8, true><<<grid, 256, 0, s>>>(a);
    }
    else if (a.world > 1 && a.mc_grads && a.mc_params) adam_shard_mc_kernel<<<grid, 256, 0, s>>>(a);
    else if (a.world == 1) adam_shard_kernel<1, false><<<grid, 256, 0, s>>>(a);
    else if (a.world == 2) adam_shard_kernel<2, false><<<grid, 256, 0, s>>>(a);
    else if (a.world <= 4) adam_shard_kernel<4, false><<<grid, 256, 0, s>>>(a);
    else adam_shard_kernel<8, false><<<grid, 256, 0, s>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fb200
