// Microbenchmark (GPU box): what does the softmax inner loop of attention2_tcgen05.cu cost per 128-score block when
// nothing else is in the way? One warp (or 2 / 4 warps on the SAME SM sub-partition) runs the instruction mix on
// registers only; cycles per block via clock64. Variants isolate each ingredient.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I aigv_assessor_b200/csrc -o /tmp/sm_ub tools/softmax_ubench.cu
#include <cstdio>
#include "ptx.cuh"
using namespace aigv;

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ void exp2_poly2(f32x2 x, float& p0, float& p1) {
    float x0, x1; unpack2(x, x0, x1);
    x0 = fmaxf(x0, -126.0f); x1 = fmaxf(x1, -126.0f);
    const f32x2 xc = pack2(x0, x1);
    const f32x2 t = add2(xc, pack2(12582912.0f, 12582912.0f));
    const f32x2 n = add2(t, pack2(-12582912.0f, -12582912.0f));
    const f32x2 f = fma2(n, pack2(-1.0f, -1.0f), xc);
    f32x2 p = fma2(pack2(0.0551715381f, 0.0551715381f), f, pack2(0.242610976f, 0.242610976f));
    p = fma2(p, f, pack2(0.693261027f, 0.693261027f));
    p = fma2(p, f, pack2(0.999928117f, 0.999928117f));
    float t0, t1, q0, q1; unpack2(t, t0, t1); unpack2(p, q0, q1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

// MODE bits: 1 = scale FFMA2, 2 = MUFU, 4 = row-sum FADD2, 8 = bf16 pack, 16 = row max (FMNMX3 pass first)
template <int MODE, int POLY>
__global__ void __launch_bounds__(256, 1) k(float* out, const float* in, int iters, int active_mask) {
    const int warp = threadIdx.x >> 5;
    if (!((active_mask >> warp) & 1)) return;
    float v[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) v[i] = in[(threadIdx.x * 131 + i * 7) & 1023];
    float sc = in[1], m = in[2];
    f32x2 sum01 = pack2(0.f, 0.f), sum23 = pack2(0.f, 0.f);
    uint32_t acc = 0;
    float mxall = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE & 16) {
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < 128; c += 8) {
                mx0 = fmaxf(mx0, fmaxf(v[c], v[c + 1])); mx1 = fmaxf(mx1, fmaxf(v[c + 2], v[c + 3]));
                mx2 = fmaxf(mx2, fmaxf(v[c + 4], v[c + 5])); mx3 = fmaxf(mx3, fmaxf(v[c + 6], v[c + 7]));
            }
            mxall += fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        }
        const f32x2 sc2 = pack2(sc, sc), nm2 = pack2(-m, -m);
#pragma unroll
        for (int c = 0; c < 128; c += 32) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
                float p[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    f32x2 x = pack2(v[c + e + 2 * q], v[c + e + 2 * q + 1]);
                    if (MODE & 1) x = fma2(x, sc2, nm2);
                    if (q < POLY) exp2_poly2(x, p[2 * q], p[2 * q + 1]);
                    else {
                        float x0, x1; unpack2(x, x0, x1);
                        if (MODE & 2) { p[2 * q] = ex2f(x0); p[2 * q + 1] = ex2f(x1); } else { p[2 * q] = x0; p[2 * q + 1] = x1; }
                    }
                }
                if (MODE & 4) {
                    sum01 = add2(sum01, add2(pack2(p[0], p[1]), pack2(p[4], p[5])));
                    sum23 = add2(sum23, add2(pack2(p[2], p[3]), pack2(p[6], p[7])));
                }
                if (MODE & 8) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) pk[(e >> 1) + q] = pack_bf16x2(p[2 * q], p[2 * q + 1]);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) pk[(e >> 1) + q] = __float_as_uint(p[2 * q]) ^ __float_as_uint(p[2 * q + 1]);
                }
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) acc ^= pk[q];
        }
        // make the next iteration depend on this one a little, so nothing is hoisted
        v[0] += __uint_as_float(acc & 1u) * 1e-30f;
        m += 1e-9f;
    }
    long long t1 = clock64();
    float s0, s1, s2, s3; unpack2(sum01, s0, s1); unpack2(sum23, s2, s3);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + __uint_as_float(acc) + mxall;
    if ((threadIdx.x & 31) == 0) out[4096 + warp] = (float)(t1 - t0) / iters;
}

template <int MODE, int POLY>
void run(const char* name, float* d_out, float* d_in) {
    const int iters = 2000;
    const int masks[4] = {0x1, 0x11, 0x3, 0xff};  // 1 warp | 2 warps on ONE sub-partition | 2 warps on two | 8 warps (2 per sub-partition)
    printf("%-44s", name);
    for (int mi = 0; mi < 4; ++mi) {
        k<MODE, POLY><<<1, 256>>>(d_out, d_in, iters, masks[mi]);
        k<MODE, POLY><<<1, 256>>>(d_out, d_in, iters, masks[mi]);
        cudaDeviceSynchronize();
        float cyc; cudaMemcpy(&cyc, d_out + 4096, 4, cudaMemcpyDeviceToHost);
        printf("  %s: %6.0f", mi == 0 ? "1w" : mi == 1 ? "2w/1sp" : mi == 2 ? "2w/2sp" : "8w", cyc);
    }
    printf("   cycles per 128-score block per warp\n");
}

int main() {
    float *d_out, *d_in; cudaMalloc(&d_out, 8192 * 4); cudaMalloc(&d_in, 1024 * 4);
    float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = -0.01f * (i % 97) - 0.5f; h[1] = 0.18f; h[2] = 0.3f;
    cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<2, 0>("MUFU only (128)", d_out, d_in);
    run<3, 0>("FFMA2 + MUFU", d_out, d_in);
    run<7, 0>("FFMA2 + MUFU + FADD2(sum)", d_out, d_in);
    run<11, 0>("FFMA2 + MUFU + F2FP", d_out, d_in);
    run<15, 0>("FFMA2 + MUFU + FADD2 + F2FP (kernel mix)", d_out, d_in);
    run<31, 0>("+ row max (FMNMX3)", d_out, d_in);
    run<13, 0>("FFMA2 + FADD2 + F2FP, no MUFU", d_out, d_in);
    run<15, 1>("kernel mix, 1/4 poly", d_out, d_in);
    run<15, 2>("kernel mix, 2/4 poly", d_out, d_in);
    run<15, 4>("kernel mix, all poly", d_out, d_in);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
