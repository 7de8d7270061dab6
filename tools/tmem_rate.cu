// Microbenchmark (GPU box): tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps, and the MUFU rate
// next to a concurrent tcgen05.ld stream (does the S read of the attention kernel compete with its exponentials?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I aigv_assessor_b200/csrc -o /tmp/tmem_rate tools/tmem_rate.cu && /tmp/tmem_rate
#include <cstdio>

#include "ptx.cuh"

using namespace aigv;

// MODE 0: every warp loads 128 columns (4 x 32x32b.x32) per iteration; 1: stores 64 columns (4 x 32x32b.x16);
// MODE 2: warps 0..nw/2-1 load, the others run MUFU.EX2 (128 per iteration); 3: all warps MUFU only (baseline of 2)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int iters, int nwarps) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        tmem_alloc<512>(&slot);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float acc = 0.f;
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
        const bool loader = MODE == 0 || (MODE == 2 && warp < nwarps / 2);
        const bool storer = MODE == 1;
        for (int it = 0; it < iters; ++it) {
            if (loader) {
                uint32_t v[4][32];   // like the attention kernel: the block's 128 scores in flight, one wait
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32(base + ((warp >> 2) & 1) * 128 + c * 32, v[c]);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c) acc += __uint_as_float(v[c][0]) + __uint_as_float(v[c][31]);
            } else if (storer) {
                uint32_t p[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) p[i] = it + i;
#pragma unroll
                for (int c = 0; c < 64; c += 16) tmem_st_32x16(base + ((warp >> 2) & 1) * 128 + c, p);
                tmem_st_wait();
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r)
#pragma unroll
                    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            }
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    // the slowest warp's time
    __shared__ long long tmax;
    if (threadIdx.x == 0) tmax = 0;
    __syncthreads();
    if (lane == 0 && warp < nwarps) atomicMax(reinterpret_cast<unsigned long long*>(&tmax), static_cast<unsigned long long>(t1 - t0));
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = tmax;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<512>(slot);
    }
}

template <int MODE>
double run(float* d, long long* c, int iters, int nw) {
    k<MODE><<<148, 512>>>(d, c, iters, nw);
    k<MODE><<<148, 512>>>(d, c, iters, nw);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < 148; ++i) s += h[i];
    return s / 148;
}

int main() {
    float* d;
    long long* c;
    cudaMalloc(&d, 148 * 512 * 4);
    cudaMalloc(&c, 148 * 8);
    const int it = 2000;
    for (int nw : {1, 2, 4, 8, 16}) {
        const double cy = run<0>(d, c, it, nw);
        printf("tcgen05.ld 32x32b.x32: %2d warps  %.0f cycles per 16 KB per warp  -> %.1f B/clk/SM\n", nw, cy / it, nw * 16384.0 * it / cy);
    }
    for (int nw : {1, 4, 8, 16}) {
        const double cy = run<1>(d, c, it, nw);
        printf("tcgen05.st 32x32b.x16: %2d warps  %.0f cycles per 8 KB per warp   -> %.1f B/clk/SM\n", nw, cy / it, nw * 8192.0 * it / cy);
    }
    for (int nw : {8, 16}) {
        const double c3 = run<3>(d, c, it, nw / 2);   // nw/2 MUFU warps alone
        const double c2 = run<2>(d, c, it, nw);       // nw/2 loaders + nw/2 MUFU warps
        const double c0 = run<0>(d, c, it, nw / 2);   // nw/2 loaders alone
        printf("%2d warps: %d MUFU warps alone %.0f cycles/iter (128 ex2 per thread), %d loaders alone %.0f, together (max) %.0f\n", nw, nw / 2,
               c3 / it, nw / 2, c0 / it, c2 / it);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
