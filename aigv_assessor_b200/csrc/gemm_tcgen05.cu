// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogue.
//
//   C[M,N] (op)= colscale * act( A[M,K] (bf16, K contiguous) x W[N,K]^T (bf16, nn.Linear weight layout) + bias )
//   op is a TMA store (bf16 or fp32) or a TMA reduce-add into the fp32 residual stream (x += ...); fp32 math.
//
// Replaces the cuBLASLt GEMM + separate bias / GELU / layer-scale / residual kernels behind
//   attn.qkv  (modeling_intern_vit.py:129,163)   attn.proj + *ls1 + residual (:141,175,224)
//   mlp.fc1 + GELU (:189,193-194)                mlp.fc2 + *ls2 + residual (:190,195,226)
//   patch_embedding conv as im2col GEMM (:97)    mlp1[1]+GELU, mlp1[3] (modeling_internvl_chat.py eval2:240-242)
//
// Tile 128x256x64, 4-stage TMA->smem ring (128B swizzle), UMMA M=128 N=256 K=16 issued by one thread,
// two 256-column fp32 accumulators in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// Warps 0-7: epilogue (warp%4 = TMEM lane quadrant, warp/4 = column half; results are staged through a 128B-swizzled
// smem tile per warp and leave by TMA store / reduce-add); warp 8: TMA producer; warp 9: MMA issuer.
#include "kernels.h"
#include "ptx.cuh"

namespace aigv {

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_BYTES = BN * BK * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int STG_BYTES = 4096;  // per epilogue warp: 32 rows x 128 B (64 bf16 or 32 fp32 columns), 128B-swizzled
constexpr int OFF_STG = STAGES * STAGE_BYTES;
constexpr int OFF_BAR = OFF_STG + EPI_WARPS * STG_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr uint32_t TMEM_COLS = 512;

// acc + bias -> act -> * colscale, on 32 consecutive columns starting at col0 (fp32 registers).
template <bool kGelu>
__device__ __forceinline__ void epilogue_math(float (&x)[32], int col0, const GemmEpilogue& ep) {
    if (ep.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
            x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
        }
    }
    if (kGelu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
    }
    if (ep.colscale) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 s = __ldg(reinterpret_cast<const float4*>(ep.colscale + col0 + j));
            x[j] *= s.x; x[j + 1] *= s.y; x[j + 2] *= s.z; x[j + 3] *= s.w;
        }
    }
}

template <bool kGelu, bool kOutF32>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, int M, int N, int K, GemmEpilogue ep) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = (M + BM - 1) / BM;
    const int num_n = (N + BN - 1) / BN;
    const int num_k = (K + BK - 1) / BK;
    const int num_tiles = num_m * num_n;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("aigv gemm: dynamic smem base not 1024B aligned\n");
        __trap();
    }
    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 9) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tfull_bar[s], 1);
                mbar_init(&tempty_bar[s], EPI_WARPS);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<TMEM_COLS>(tmem_slot);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ------------------------------------------------------------ TMA producer
        // (whole warp converged, one elected lane issues: keeps addresses/coordinates in uniform registers)
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / num_n, n_blk = tile % num_n;
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    tma_load_2d(sa + A_BYTES, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------ MMA issuer
        // The warp runs the loop converged and one elected lane issues the tcgen05 instructions; inside a divergent
        // `if (lane == 0)` every UTCHMMA gets wrapped in an ELECT/BRA.U.ANY loop with R2UR transfers.
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_u + acc * BN;
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
                const uint32_t b_addr = a_addr + A_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = umma_desc_sw128(a_addr + k * UMMA_K * 2, 16, 1024);
                        const uint64_t db = umma_desc_sw128(b_addr + k * UMMA_K * 2, 16, 1024);
                        umma_bf16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
                    if (kb == num_k - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else {
        // ------------------------------------------------------------ epilogue warps
        // TMEM -> registers (tcgen05.ld, lane == row) -> fp32 math -> 128B-swizzled smem staging tile -> TMA store
        // (or TMA reduce-add into the fp32 residual stream): global traffic is fully coalesced and asynchronous.
        const int quad = warp & 3;
        const int half = warp >> 2;
        uint8_t* stg = smem + OFF_STG + warp * STG_BYTES;
        uint8_t* my_row = stg + lane * 128;
        const int sw = lane & 7;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / num_n, n_blk = tile % num_n;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int row0 = m_blk * BM + quad * 32;
            const uint32_t tm = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + half * 128;
            constexpr int CHUNK = kOutF32 ? 32 : 64;  // columns per staged tile (128 B per row)
#pragma unroll 1
            for (int c = 0; c < 128; c += CHUNK) {
                const int col0 = n_blk * BN + half * 128 + c;
                uint32_t v[32];
                float x[32];
                uint4 packed[8];
                tmem_ld_32x32(tm + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
                if (col0 < N) epilogue_math<kGelu>(x, col0, ep);
                if (kOutF32) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        packed[j] = make_uint4(__float_as_uint(x[4 * j]), __float_as_uint(x[4 * j + 1]),
                                               __float_as_uint(x[4 * j + 2]), __float_as_uint(x[4 * j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        packed[j] = make_uint4(pack_bf16x2(x[8 * j], x[8 * j + 1]), pack_bf16x2(x[8 * j + 2], x[8 * j + 3]),
                                               pack_bf16x2(x[8 * j + 4], x[8 * j + 5]), pack_bf16x2(x[8 * j + 6], x[8 * j + 7]));
                    tmem_ld_32x32(tm + c + 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
                    if (col0 + 32 < N) epilogue_math<kGelu>(x, col0 + 32, ep);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        packed[4 + j] = make_uint4(pack_bf16x2(x[8 * j], x[8 * j + 1]), pack_bf16x2(x[8 * j + 2], x[8 * j + 3]),
                                                   pack_bf16x2(x[8 * j + 4], x[8 * j + 5]), pack_bf16x2(x[8 * j + 6], x[8 * j + 7]));
                }
                // the previous TMA store of this warp must have finished reading the staging tile
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) = packed[j];
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && row0 < M && col0 < N) {
                    if (ep.accumulate) tma_reduce_add_2d(&tmC, stg, col0, row0);
                    else tma_store_2d(&tmC, stg, col0, row0);
                    tma_store_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (lane == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

template <bool kGelu, bool kOutF32>
cudaError_t launch_variant(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, int M, int N, int K,
                           const GemmEpilogue& ep, int num_sms, cudaStream_t stream) {
    auto kern = gemm_bf16_tn_kernel<kGelu, kOutF32>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int grid = tiles < num_sms ? tiles : num_sms;
    kern<<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, tmC, M, N, K, ep);
    return cudaGetLastError();
}

}  // namespace

int gemm_box_rows_a() { return BM; }
int gemm_box_rows_b() { return BN; }
int gemm_box_cols() { return BK; }

int gemm_box_rows_c() { return 32; }
int gemm_box_cols_c(bool out_f32) { return out_f32 ? 32 : 64; }

cudaError_t launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, int M, int N, int K,
                        const GemmEpilogue& ep, int num_sms, cudaStream_t stream) {
    if (M <= 0 || N <= 0 || K <= 0 || (N % 32) != 0 || (K % 8) != 0) return cudaErrorInvalidValue;
    if (ep.accumulate && !ep.out_f32) return cudaErrorInvalidValue;
    if (ep.gelu) {
        return ep.out_f32 ? launch_variant<true, true>(tmA, tmB, tmC, M, N, K, ep, num_sms, stream)
                          : launch_variant<true, false>(tmA, tmB, tmC, M, N, K, ep, num_sms, stream);
    }
    return ep.out_f32 ? launch_variant<false, true>(tmA, tmB, tmC, M, N, K, ep, num_sms, stream)
                      : launch_variant<false, false>(tmA, tmB, tmC, M, N, K, ep, num_sms, stream);
}

}  // namespace aigv
