// Microbenchmark (GPU box): tcgen05.mma issue/latency behaviour for small tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I aigv_assessor_b200/csrc -o /tmp/umma_latency tools/umma_latency.cu
// Reports cycles per MMA for chains of dependent MMAs (same accumulator) vs round-robin over NACC independent
// accumulators, for N in {64,128,256} (M=128, K=16, bf16), and the commit -> mbarrier-observed round trip.
#include <cstdio>
#include "ptx.cuh"
using namespace aigv;

template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) chain_kernel(long long* out, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) {
        if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
        __syncwarp();
        tmem_alloc<512>(&slot);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0) {
        const uint32_t a = smem_u32(smem), b = a + 16384;
        const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
        const uint64_t da = umma_desc_sw128(a, 16, 1024), db = umma_desc_sw128(b, 16, 1024);
        long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int k = 0; k < NACC; ++k) umma_bf16_ss(tm + k * N, da, db, idesc, 1u);
            }
            umma_commit(&bar);
        }
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

// round trip: issue ONE group of `n` MMAs + commit, wait, repeat (like one attention block step)
template <int N>
__global__ void __launch_bounds__(128, 1) roundtrip_kernel(long long* out, int iters, int n) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) {
        if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
        __syncwarp();
        tmem_alloc<512>(&slot);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0) {
        const uint32_t a = smem_u32(smem), b = a + 16384;
        const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
        const uint64_t da = umma_desc_sw128(a, 16, 1024), db = umma_desc_sw128(b, 16, 1024);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (elect_one()) {
                for (int k = 0; k < n; ++k) umma_bf16_ss(tm, da, db, idesc, k ? 1u : 0u);
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, i & 1);
            tc_fence_after();
        }
        long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int N, int NACC>
void run_chain(long long* d, int iters) {
    auto k = chain_kernel<N, NACC>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<<<148, 128, 65536>>>(d, iters);
    k<<<148, 128, 65536>>>(d, iters);
    cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d acc=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)  err=%s\n", N, NACC,
           (double)h[0] / (iters * NACC), (double)h[1] / (iters * NACC), N / 2, cudaGetErrorString(cudaGetLastError()));
}
template <int N>
void run_rt(long long* d, int iters, int n) {
    auto k = roundtrip_kernel<N>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<<<148, 128, 65536>>>(d, iters, n);
    k<<<148, 128, 65536>>>(d, iters, n);
    cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("round trip N=%3d, %d dependent MMAs + commit + wait: %.1f cycles per group (MMA floor %d)  err=%s\n", N, n,
           (double)h[0] / iters, n * N / 2, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    const int it = 2000;
    run_chain<64, 1>(d, it); run_chain<64, 2>(d, it); run_chain<64, 4>(d, it);
    run_chain<128, 1>(d, it); run_chain<128, 2>(d, it);
    run_chain<256, 1>(d, it); run_chain<256, 2>(d, it);
    run_rt<64>(d, 500, 1); run_rt<64>(d, 500, 4); run_rt<64>(d, 500, 8);
    run_rt<128>(d, 500, 4); run_rt<256>(d, 500, 4);
    return 0;
}
