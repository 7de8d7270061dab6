// Microbenchmark (GPU box): MUFU.EX2 throughput, f32 vs packed f16x2, per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_rate tools/mufu_rate.cu && /tmp/mufu_rate
#include <cstdio>
#include <cuda_fp16.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters) {
    float a[8];
    unsigned h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0x80018002u + threadIdx.x + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            else asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 1024 * 4);
    const int it = 4096;
    for (int mode = 0; mode < 2; ++mode) {
        if (mode == 0) { k<0><<<148, 1024>>>(d, it); k<0><<<148, 1024>>>(d, it); }
        else { k<1><<<148, 1024>>>(d, it); k<1><<<148, 1024>>>(d, it); }
        cudaDeviceSynchronize();
        float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
        // 32 warps/SM, 8 instr per iteration per warp
        double instr_per_sm = 32.0 * 8 * it;
        printf("%s: %.2f cycles per warp-instruction per SM -> %.1f lanes/clk/SM (%s results/clk/SM %.1f)\n",
               mode ? "ex2.f16x2" : "ex2.f32  ", cyc / instr_per_sm, 32.0 * instr_per_sm / cyc, mode ? "x2:" : "",
               (mode ? 2 : 1) * 32.0 * instr_per_sm / cyc);
    }
    return 0;
}
