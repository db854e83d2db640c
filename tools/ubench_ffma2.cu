// Microbenchmark: issue cost of packed fp32 (FFMA2) vs scalar FFMA on sm_100a, alone and mixed with integer ALU work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o ubench_ffma2 tools/ubench_ffma2.cu && ./ubench_ffma2
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int* iout, float seed) {
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
    const float m = 1.0001f, c = 0.5f;
    const float2 m2 = make_float2(m, m), c2 = make_float2(c, c);
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {          // 8 scalar FFMA (8 independent chains)
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        } else if (MODE == 1) {   // 4 FFMA2 = same flops
            p0 = __ffma2_rn(p0, m2, c2); p1 = __ffma2_rn(p1, m2, c2); p2 = __ffma2_rn(p2, m2, c2); p3 = __ffma2_rn(p3, m2, c2);
        } else if (MODE == 2) {   // 8 scalar FFMA + 8 integer ops
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
            i0 = (i0 ^ 0x55) + 3; i1 = (i1 ^ 0x33) + 5; i2 = (i2 ^ 0x0f) + 7; i3 = (i3 ^ 0x71) + 9;
        } else if (MODE == 3) {   // 4 FFMA2 + 8 integer ops
            p0 = __ffma2_rn(p0, m2, c2); p1 = __ffma2_rn(p1, m2, c2); p2 = __ffma2_rn(p2, m2, c2); p3 = __ffma2_rn(p3, m2, c2);
            i0 = (i0 ^ 0x55) + 3; i1 = (i1 ^ 0x33) + 5; i2 = (i2 ^ 0x0f) + 7; i3 = (i3 ^ 0x71) + 9;
        } else if (MODE == 4) {   // 8 FFMA2 (twice the flops of mode 0)
            p0 = __ffma2_rn(p0, m2, c2); p1 = __ffma2_rn(p1, m2, c2); p2 = __ffma2_rn(p2, m2, c2); p3 = __ffma2_rn(p3, m2, c2);
            p0 = __ffma2_rn(p0, c2, m2); p1 = __ffma2_rn(p1, c2, m2); p2 = __ffma2_rn(p2, c2, m2); p3 = __ffma2_rn(p3, c2, m2);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = i0 + i1 + i2 + i3;
}

template <int MODE>
float run(float* out, int* iout, int blocks) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, 256>>>(out, iout, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<blocks, 256>>>(out, iout, 1.0f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8;     // 8 CTAs x 8 warps = 64 warps / SM
    float* out; int* iout; cudaMalloc(&out, blocks * 256 * 4); cudaMalloc(&iout, blocks * 256 * 4);
    const char* names[] = {"8 FFMA", "4 FFMA2 (same flops)", "8 FFMA + 8 int", "4 FFMA2 + 8 int", "8 FFMA2 (2x flops)"};
    float ms[5] = {run<0>(out, iout, blocks), run<1>(out, iout, blocks), run<2>(out, iout, blocks), run<3>(out, iout, blocks),
                   run<4>(out, iout, blocks)};
    for (int i = 0; i < 5; ++i) {
        // warp-iterations per SMSP: 64 warps / 4 SMSP = 16 warps, ITERS iterations each
        const double cycles = ms[i] * 1e-3 * clk * 1e3;
        printf("%-24s %.3f ms  -> %.2f cycles per warp-iteration per SMSP (clock %d MHz nominal)\n", names[i], ms[i],
               cycles / (16.0 * ITERS), clk / 1000);
    }
    return 0;
}
