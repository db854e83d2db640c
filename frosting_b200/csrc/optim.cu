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

__device__ __forceinline__ float adam1(float g, float& m, float& v, float p, float lr_over_bc1, const fb200_adam_args& a) {
    m = m + (g - m) * (1.0f - a.beta1);
    v = v * a.beta2 + (1.0f - a.beta2) * g * g;
    const float denom = sqrtf(v) / a.bias_correction2_sqrt + a.eps;
    return p - lr_over_bc1 * (m / denom);
}

__global__ void __launch_bounds__(256)
adam_shard_kernel(const fb200_adam_args a) {
    const int64_t v_lo = a.shard_lo >> 2, v_hi = a.shard_hi >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float4* __restrict__ mom1 = reinterpret_cast<float4*>(a.d_exp_avg);
    float4* __restrict__ mom2 = reinterpret_cast<float4*>(a.d_exp_avg_sq);
    for (int64_t i = v_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v_hi; i += stride) {
        // gradient: one 128-bit load per peer, all issued before the first use
        float4 g[FB200_MAX_PEERS];
#pragma unroll
        for (int p = 0; p < FB200_MAX_PEERS; ++p)
            if (p < a.world) g[p] = ld_stream(reinterpret_cast<const float4*>(a.peer_grads[p]) + i);
        float4 s = g[0];
#pragma unroll
        for (int p = 1; p < FB200_MAX_PEERS; ++p)
            if (p < a.world) { s.x += g[p].x; s.y += g[p].y; s.z += g[p].z; s.w += g[p].w; }
        s.x *= a.grad_scale; s.y *= a.grad_scale; s.z *= a.grad_scale; s.w *= a.grad_scale;

        // group of this vector (group starts are multiples of 4 elements, so a vector never straddles two)
        const int64_t e = i << 2;
        float lr = a.lr[0];
#pragma unroll
        for (int k = 1; k < FB200_ADAM_MAX_GROUPS; ++k)
            if (k < a.n_groups && e >= a.group_start[k]) lr = a.lr[k];
        const float step = lr / a.bias_correction1;

        const int64_t li = i - v_lo;
        float4 m = mom1[li], v = mom2[li];
        float4 p = reinterpret_cast<const float4*>(a.peer_params[a.rank])[i];
        p.x = adam1(s.x, m.x, v.x, p.x, step, a);
        p.y = adam1(s.y, m.y, v.y, p.y, step, a);
        p.z = adam1(s.z, m.z, v.z, p.z, step, a);
        p.w = adam1(s.w, m.w, v.w, p.w, step, a);
        mom1[li] = m;
        mom2[li] = v;
#pragma unroll
        for (int q = 0; q < FB200_MAX_PEERS; ++q)
            if (q < a.world) __stcg(reinterpret_cast<float4*>(a.peer_params[q]) + i, p);
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
    adam_shard_kernel<<<grid, 256, 0, s>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fb200
