// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogue, CTA-pair (cta_group::2) form.
//
//   C[M,N] (op)= colscale * act( A[M,K] (bf16, K contiguous) x W[N,K]^T (bf16, nn.Linear weight layout) + bias )
//   op is a TMA store (bf16 or fp32) or a TMA reduce-add into the fp32 residual stream (x += ...); fp32 math.
//
// Replaces the cuBLASLt GEMM + separate bias / GELU / layer-scale / residual kernels behind
//   attn.qkv  (modeling_intern_vit.py:129,163)   attn.proj + *ls1 + residual (:141,175,224)
//   mlp.fc1 + GELU (:189,193-194)                mlp.fc2 + *ls2 + residual (:190,195,226)
//   patch_embedding conv as im2col GEMM (:97)    mlp1[1]+GELU, mlp1[3] (modeling_internvl_chat.py eval2:240-242)
//
// A CTA pair (cluster of 2, one TPC) owns a 256x256 output tile; UMMA M=256 N=256 K=16 is issued by the leader CTA and
// each CTA stages only its own 128 rows of A and its own 128 rows of W (half of N) per k-block: 32 KiB instead of 48 KiB
// of TMA fill + operand reads per 128x256x64 of MACs. A single-CTA 128x256 tile is shared-memory-bandwidth bound (~96
// B/clk of operand reads + ~94 B/clk of TMA fills against ~128 B/clk; measured 54-73% tensor-pipe active); the pair
// needs ~125 B/clk at full MMA rate and measures 90% tensor-pipe active on the K >= 1024 shapes.
// Results are staged through a 128B-swizzled smem tile per epilogue warp and leave by TMA store / reduce-add.
//   warp 8: TMA producer (each CTA loads its own operands; transaction bytes are credited to the LEADER's full barrier)
//   warp 9: leader only: MMA issuer; commits are multicast to both CTAs' empty / accumulator-full barriers
//   warps 0-15: epilogue of this CTA's 128 rows (warp%4 = TMEM lane quadrant, warp/4 = 64-column slice; 4 warps per
//   SMSP hide the GELU / TMEM-load latencies); they release the accumulator stage on the leader's barrier (remote arrive)
//   warp 16: TMA producer, warp 17: MMA issuer
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace aigv {

namespace {

constexpr int BM = 128;          // rows per CTA (256 per pair)
constexpr int BN = 256;          // columns per pair tile
constexpr int BNH = 128;         // W rows staged per CTA
constexpr int BK = 64, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;   // 16 KiB
constexpr int B_BYTES = BNH * BK * 2;  // 16 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STG_BYTES = 4096;
constexpr int STATS_COLS = 64;  // columns per row-statistics slice of the fused-residual epilogues (one bf16 staging tile)
constexpr uint32_t TMEM_COLS = 512;

// Shapes of the same kernel: the GELU epilogue is instruction-issue heavy and wants 16 epilogue warps (4 per SMSP) at the
// price of one pipeline stage; the plain epilogues keep 8 warps and 6 stages; the fused-residual epilogue comes in two
// forms: kResid == 2 (long K, fc2: the mainloop hides a serial epilogue) re-uses ONE staging tile per warp for the x_old
// fetch, the fp32 x_new store and the bf16 copy and keeps 6 stages (measured: 4 instead of 6 stages costs 8% on fc2);
// kResid == 1 (short K, proj: the epilogue is the critical path) needs
// three staging tiles per warp (two fp32 tiles for the x_old prefetch / x_new store, one bf16 tile) and runs 4 stages.
// kResid == 3 ("delta" producer, the 300M tower's proj): proj is HBM-bound on the fp32 residual stream (269 MB read + 269 MB
// written per launch on top of 134 + 134 MB of bf16), so it does not touch the fp32 master at all: it reads the bf16 copy
// xb_old, writes xb_new = bf16(xb_old + d) in place and the update d = colscale * (acc + bias) as bf16; the NEXT producer
// (fc2, kResid == 2 with ep.delta_in) folds d into the master: x_new = x_old + d + colscale * (acc + bias). 536 instead of
// 806 MB per proj launch; fc2 (compute-bound, long mainloop) reads 134 MB more.
template <int kEpiWarps, int kResid>
struct Cfg {
    static constexpr int EPI_WARPS = kEpiWarps;          // multiple of 4: warp%4 = TMEM lane quadrant
    static constexpr int SLICES = kEpiWarps / 4;         // column slices of the 256-column tile
    static constexpr int SLICE_COLS = BN / SLICES;       // 64 or 128 columns per warp
    static constexpr int STAGES = (kResid == 1 || kResid == 3) ? 4 : (kEpiWarps == 16 ? 5 : 6);
    static constexpr int THREADS = (kEpiWarps + 2) * 32;
    static constexpr int WARP_STG = (kResid == 1 || kResid == 3) ? 3 * STG_BYTES : STG_BYTES;
    static constexpr int OFF_STG = STAGES * STAGE_BYTES;
    static constexpr int OFF_BAR = OFF_STG + kEpiWarps * WARP_STG;
    static constexpr int SMEM_BYTES = OFF_BAR + 512;     // 2*STAGES + 4 pipeline barriers, 2 per epilogue warp, TMEM slot
};

// acc -> [LayerNorm applied algebraically] -> + bias -> act -> * colscale, on 32 consecutive columns from col0.
// LN fusion: the A operand was the un-normalised activation x (bf16) and W carries gamma, so
//   LN(x) W^T = rstd * (x W'^T - mean * wsum) + (b + W beta),   wsum[n] = sum_k W'[n,k]
// with the row's mean / rstd taken from the (sum, sum of squares) the producing GEMM accumulated.
template <int kAct>   // 0: none, 1: erf GELU, 2: SwiGLU on (gate, up) column pairs (applied by the caller: halves the width), 3: rotary (by the caller)
__device__ __forceinline__ void epilogue_math(float (&x)[32], int col0, const GemmEpilogue& ep, float mean, float rstd) {
    if (ep.ln_stats) {
        // x = rstd * (acc - mean * wsum) + bias as two FMAs per element: fma(rstd, acc, fma(-mean*rstd, wsum, bias))
        const float nmr = -mean * rstd;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            float4 b = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (ep.ln_wsum) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(ep.ln_wsum + col0 + j));
                b.x = fmaf(nmr, w.x, b.x); b.y = fmaf(nmr, w.y, b.y); b.z = fmaf(nmr, w.z, b.z); b.w = fmaf(nmr, w.w, b.w);
            }
            x[j] = fmaf(rstd, x[j], b.x); x[j + 1] = fmaf(rstd, x[j + 1], b.y);
            x[j + 2] = fmaf(rstd, x[j + 2], b.z); x[j + 3] = fmaf(rstd, x[j + 3], b.w);
        }
    } else if (ep.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
            x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
        }
    }
    if (kAct == 0 && ep.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
    }
    if (kAct == 1) {
#pragma unroll
#ifdef AIGV_GELU_AS7126  // the older Abramowitz-Stegun 7.1.26 form (18 issue slots per pair; fc1 2.4% slower)
        for (int j = 0; j < 32; j += 2) gelu_erf2(x[j], x[j + 1]);
#else
        for (int j = 0; j < 32; j += 2) gelu_sig2(x[j], x[j + 1]);
#endif
    }
    if (ep.colscale) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 s = __ldg(reinterpret_cast<const float4*>(ep.colscale + col0 + j));
            x[j] *= s.x; x[j + 1] *= s.y; x[j + 2] *= s.z; x[j + 3] *= s.w;
        }
    }
}

template <int kAct, bool kOutF32, int kEpiWarps, int kResid>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg<kEpiWarps, kResid>::THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmX, int M, int N, int K,
                    GemmEpilogue ep) {
    using C = Cfg<kEpiWarps, kResid>;
    constexpr int EPI_WARPS = C::EPI_WARPS, STAGES = C::STAGES, OFF_STG = C::OFF_STG, OFF_BAR = C::OFF_BAR;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* xload_bar = tempty_bar + 2;  // [EPI_WARPS][2] (kResid): x_old staging tile landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xload_bar + 2 * EPI_WARPS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();   // 0 = leader
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int num_m = (M + 2 * BM - 1) / (2 * BM);
    const int num_n = (N + BN - 1) / BN;
    const int num_k = (K + BK - 1) / BK;
    const int num_tiles = num_m * num_n;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("aigv gemm: dynamic smem base not 1024B aligned\n");
        __trap();
    }
    if (warp == EPI_WARPS && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        if (kResid) tma_prefetch_desc(&tmX);
    }
    if (warp == EPI_WARPS + 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full_bar[s], 1);    // leader: its own producer's arrive.expect_tx (bytes of BOTH CTAs)
                mbar_init(&empty_bar[s], 1);   // multicast commit from the leader's MMA thread
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tfull_bar[s], 1);                // multicast commit
                mbar_init(&tempty_bar[s], 2 * EPI_WARPS);   // leader: epilogue warps of both CTAs
            }
            if (kResid)
                for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(&xload_bar[s], 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_pair<TMEM_COLS>(tmem_slot);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / TMA credit / multicast commit
    __syncthreads();      // (the cluster barrier already orders the TMEM-slot write; compute-sanitizer racecheck only
                          // recognises the CTA barrier and otherwise reports the slot read below as a hazard)
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above overlapped the previous kernel's tail (launch_pdl); its outputs -- this kernel's A, the residual
    // stream, the row statistics -- are complete from here on
    pdl_wait();
    pdl_launch_dependents();

    if (warp == EPI_WARPS) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            int m_blk, n_blk;
            tile_coords(tile, num_m, num_n, ep.group_m, m_blk, n_blk);
            const int row_a = m_blk * 2 * BM + rank * BM;
            const int row_b = n_blk * BN + rank * BNH;
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    tma_load_2d_pair(sa, &tmA, &full_bar[stage], kb * BK, row_a);
                    tma_load_2d_pair(sa + A_BYTES, &tmB, &full_bar[stage], kb * BK, row_b);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == EPI_WARPS + 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, 0);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
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
                            umma_bf16_ss_pair(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit_pair(&empty_bar[stage], 3);                           // free the slot in both CTAs
                        if (kb == num_k - 1) umma_commit_pair(&tfull_bar[acc], 3);        // accumulator -> both epilogues
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps (this CTA's 128 rows)
        const int quad = warp & 3;
        const int slice = warp >> 2;  // column slice of the 256-column tile
        constexpr int SLICE_COLS = C::SLICE_COLS;
        uint8_t* stg = smem + OFF_STG + warp * C::WARP_STG;
        const int sw = lane & 7;
        const uint32_t tempty_leader0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
        const uint32_t tempty_leader1 = mapa_shared(smem_u32(&tempty_bar[1]), 0);
        uint32_t xphase = 0;  // kResid: parity bits of this warp's two x_old staging barriers
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            int m_blk, n_blk;
            tile_coords(tile, num_m, num_n, ep.group_m, m_blk, n_blk);
            const int row0 = m_blk * 2 * BM + rank * BM + quad * 32;
            const int colb = n_blk * BN + slice * SLICE_COLS;
            // kResid: x_old does not depend on the MMAs: start fetching this warp's first staging tile before waiting
            auto issue_xload = [&](int c) {
                if (lane == 0 && row0 < M && colb + c < N) {
                    // kResid 2: a single staging tile; 1: fp32 tiles alternate per 32-column chunk; 3: bf16 tiles per 64 columns
                    const int buf = kResid == 2 ? 0 : (kResid == 3 ? (c >> 6) & 1 : (c >> 5) & 1);
                    uint64_t* bar = &xload_bar[2 * warp + buf];
                    mbar_arrive_expect_tx(bar, STG_BYTES);
                    tma_load_2d(stg + buf * STG_BYTES, kResid == 3 ? &tmX : &tmC, bar, colb + c, row0);
                }
            };
            if (kResid) {
                if (lane == 0) tma_store_wait_read<0>();  // staging tiles of the previous output tile have been read out
                __syncwarp();
                issue_xload(0);
                if (kResid == 3 && SLICE_COLS > 64) issue_xload(64);   // both 64-column groups of the slice, a mainloop ahead
                // x_old is streamed from HBM (the residual stream is 4 B x M x N, far beyond L2). Optional L2 prefetches of this
                // warp's rows (both measured and off by default, see the two macros below).
                auto prefetch_rows = [&](int tl) {
                    int mb, nb;
                    tile_coords(tl, num_m, num_n, ep.group_m, mb, nb);
                    const int r0 = mb * 2 * BM + static_cast<int>(rank) * BM + quad * 32 + lane;
                    const int c0 = nb * BN + slice * SLICE_COLS;
                    if (r0 < M && c0 < N)
                        prefetch_l2_bulk(ep.resid_in + static_cast<size_t>(r0) * N + c0,
                                         4u * static_cast<uint32_t>(N - c0 < SLICE_COLS ? N - c0 : SLICE_COLS));
                };
#ifndef AIGV_GEMM_PREFETCH_NEXT
#define AIGV_GEMM_PREFETCH_NEXT 0   // measured (profiles/r2c_ncu_gemm_summary.txt): +110 MB (proj) / +216 MB (fc2) of DRAM re-reads per launch and 6 % slower proj -- the lines are evicted again before the TMA fetch wants them
#endif
                if (AIGV_GEMM_PREFETCH_NEXT) {
                    if (tile == pair) prefetch_rows(tile);
                    if (tile + num_pairs < num_tiles) prefetch_rows(tile + num_pairs);
                } else if (kResid == 2) {
#ifdef AIGV_GEMM_XPREFETCH   // round 1 pulled this tile's x_old rows into L2 here; with the delta scheme it costs 1 % on fc2,
                             // the 6B tower and the LLM's wo / w2 (same-box A/B, profiles/r2r_xprefetch_ab.txt): off
                    prefetch_rows(tile);
#endif
                }
            }
            // LayerNorm-fused consumer: this thread's row statistics (fetched ahead of the accumulator wait)
            float mean = 0.f, rstd = 1.f;
            if (ep.ln_stats) {
                const int row = row0 + lane;
                float sx = 0.f, sq = 0.f;
                if (row < M)
                    for (int p = 0; p < ep.ln_nparts; ++p) {  // fixed order: results do not depend on scheduling
                        const float2 st = __ldg(ep.ln_stats + static_cast<size_t>(p) * M + row);
                        sx += st.x;
                        sq += st.y;
                    }
                mean = ep.ln_wsum ? sx * ep.ln_inv_d : 0.f;
                rstd = rsqrtf(fmaxf(sq * ep.ln_inv_d - mean * mean, 0.f) + ep.ln_eps);
            }
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t tm = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + slice * SLICE_COLS;
            // 32 accumulator columns -> (LN) / bias / GELU / scale in fp32 (in place). The accumulator stage is handed
            // back to the MMA warp as soon as this warp's LAST tcgen05.ld of the tile has completed.
            auto load32 = [&](int c, float (&x)[32]) {
                uint32_t v[32];
                tmem_ld_32x32(tm + c, v);
                tmem_ld_wait();
                if (c + 32 == SLICE_COLS) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
                if (colb + c < N) epilogue_math<kAct>(x, colb + c, ep, mean, rstd);
            };
            auto flush = [&](const uint4 (&packed)[8], int col0) {
                uint8_t* my_row = stg + lane * 128;
                if (lane == 0) tma_store_wait_read<0>();  // previous TMA store has finished reading the staging tile
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
            };
            if (kResid == 3) {
                // "Delta" producer: xb_new = bf16(xb_old + d) in place, d = colscale * (acc + bias) out as bf16, row statistics
                // of the new value; the fp32 master is brought up to date by the next producer (ep.delta_in there).
                static_assert(kResid != 3 || SLICE_COLS == 128, "two 64-column groups per warp");
                uint8_t* dtile = stg + 2 * STG_BYTES;
#pragma unroll 1
                for (int g = 0; g < SLICE_COLS / 64; ++g) {
                    float xa[32], xc[32];
                    load32(64 * g, xa);
                    load32(64 * g + 32, xc);
                    if (!(row0 < M && colb + 64 * g < N)) continue;  // warp-uniform (N is a multiple of 64)
                    mbar_wait(&xload_bar[2 * warp + g], (xphase >> g) & 1);
                    xphase ^= 1u << g;
                    if (g) {   // the d tile of group 0 must have been read out by its store
                        if (lane == 0) tma_store_wait_read<0>();
                        __syncwarp();
                    }
                    uint8_t* lrow = stg + g * STG_BYTES + lane * 128;
                    uint8_t* drow = dtile + lane * 128;
                    float rs = 0.f, rq = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float* xx = j < 4 ? &xa[8 * j] : &xc[8 * (j - 4)];
                        uint4* slot = reinterpret_cast<uint4*>(lrow + ((j ^ sw) << 4));
                        const uint4 o = *slot;
                        const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
                        uint32_t yw[4], dw[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float y0 = __uint_as_float(ow[e] << 16) + xx[2 * e];
                            const float y1 = __uint_as_float(ow[e] & 0xffff0000u) + xx[2 * e + 1];
                            rs += y0 + y1;
                            rq += y0 * y0 + y1 * y1;
                            yw[e] = pack_bf16x2(y0, y1);
                            dw[e] = pack_bf16x2(xx[2 * e], xx[2 * e + 1]);
                        }
                        *slot = make_uint4(yw[0], yw[1], yw[2], yw[3]);
                        *reinterpret_cast<uint4*>(drow + ((j ^ sw) << 4)) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmX, stg + g * STG_BYTES, colb + 64 * g, row0);   // xb_new (in place over xb_old)
                        tma_store_2d(&tmC, dtile, colb + 64 * g, row0);                  // d, bf16
                        tma_store_commit();
                    }
                    if (ep.stats_out && row0 + lane < M)
                        ep.stats_out[static_cast<size_t>((colb + 64 * g) / STATS_COLS) * M + row0 + lane] = make_float2(rs, rq);
                }
            } else if (kResid == 2) {
                // Fused residual producer, serial form: one staging tile per warp. Per 32-column chunk: x_old (TMA) ->
                // in-place x_new -> fp32 store -> (every second chunk) the 64-column bf16 copy through the same tile ->
                // fetch of the next chunk. Every step waits for the previous one; the long mainloop hides it.
                float rs = 0.f, rq = 0.f;
                uint32_t pbr[32];  // bf16 pairs of the current 64-column pair of chunks
                uint8_t* frow = stg + lane * 128;
#pragma unroll
                for (int c = 0; c < SLICE_COLS; c += 32) {
                    const int b = (c >> 5) & 1;
                    float x[32];
                    load32(c, x);
                    if (!(row0 < M && colb + c < N)) continue;  // warp-uniform: nothing of this chunk is inside C
                    if (ep.delta_in && row0 + lane < M) {
                        // the previous ("delta") producer's update, not yet in the fp32 master: this row's 32 bf16 values
                        const uint4* dp = reinterpret_cast<const uint4*>(ep.delta_in + static_cast<size_t>(row0 + lane) * N + colb + c);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 dv = __ldg(dp + q);
                            const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                x[8 * q + 2 * e] += __uint_as_float(dw[e] << 16);
                                x[8 * q + 2 * e + 1] += __uint_as_float(dw[e] & 0xffff0000u);
                            }
                        }
                    }
                    mbar_wait(&xload_bar[2 * warp], xphase & 1);
                    xphase ^= 1u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint4* slot = reinterpret_cast<uint4*>(frow + ((j ^ sw) << 4));
                        const uint4 o = *slot;
                        const float y0 = __uint_as_float(o.x) + x[4 * j], y1 = __uint_as_float(o.y) + x[4 * j + 1];
                        const float y2 = __uint_as_float(o.z) + x[4 * j + 2], y3 = __uint_as_float(o.w) + x[4 * j + 3];
                        rs += (y0 + y1) + (y2 + y3);
                        rq += (y0 * y0 + y1 * y1) + (y2 * y2 + y3 * y3);
                        *slot = make_uint4(__float_as_uint(y0), __float_as_uint(y1), __float_as_uint(y2), __float_as_uint(y3));
                        pbr[16 * b + 2 * j] = pack_bf16x2(y0, y1);
                        pbr[16 * b + 2 * j + 1] = pack_bf16x2(y2, y3);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmC, stg, colb + c, row0);  // x_new, fp32
                        tma_store_commit();
                        tma_store_wait_read<0>();
                    }
                    __syncwarp();
                    if (b) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<uint4*>(frow + ((q ^ sw) << 4)) =
                                make_uint4(pbr[4 * q], pbr[4 * q + 1], pbr[4 * q + 2], pbr[4 * q + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmX, stg, colb + c - 32, row0);  // x_new, bf16 (64 columns)
                            tma_store_commit();
                            tma_store_wait_read<0>();
                        }
                        __syncwarp();
                    }
                    if (c + 32 < SLICE_COLS) issue_xload(c + 32);
                    if (b) {  // per-row partial (sum, sum of squares) of this 64-column pair of chunks
                        if (ep.stats_out && row0 + lane < M)
                            ep.stats_out[static_cast<size_t>((colb + c) / STATS_COLS) * M + row0 + lane] = make_float2(rs, rq);
                        rs = 0.f;
                        rq = 0.f;
                    }
                }
            } else if (kResid) {
                // Fused residual producer: x_new = x_old + colscale * (acc + bias) computed here (x_old tiles arrive by
                // TMA, double buffered), written back as fp32 (the residual stream) AND as bf16 (the next GEMM's A
                // operand), with the row's (sum, sum of squares) accumulated for the LayerNorm the next GEMM applies.
                float rs = 0.f, rq = 0.f;
                uint8_t* btile = stg + 2 * STG_BYTES;  // bf16 staging: 32 rows x 64 columns
#pragma unroll 1
                for (int c = 0; c < SLICE_COLS; c += 32) {
                    const int b = (c >> 5) & 1;
                    uint8_t* frow = stg + b * STG_BYTES + lane * 128;
                    if (c + 32 < SLICE_COLS) {
                        // the other fp32 tile (and, on even chunks, the bf16 tile) must have been read out by their stores
                        if (lane == 0) tma_store_wait_read<0>();
                        __syncwarp();
                        issue_xload(c + 32);
                    }
                    float x[32];
                    load32(c, x);
                    if (!(row0 < M && colb + c < N)) continue;  // warp-uniform: nothing of this chunk is inside C
                    mbar_wait(&xload_bar[2 * warp + b], (xphase >> b) & 1);
                    xphase ^= 1u << b;
                    uint4 pb[4];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint4* slot = reinterpret_cast<uint4*>(frow + ((j ^ sw) << 4));
                        const uint4 o = *slot;
                        const float y0 = __uint_as_float(o.x) + x[4 * j], y1 = __uint_as_float(o.y) + x[4 * j + 1];
                        const float y2 = __uint_as_float(o.z) + x[4 * j + 2], y3 = __uint_as_float(o.w) + x[4 * j + 3];
                        rs += (y0 + y1) + (y2 + y3);
                        rq += (y0 * y0 + y1 * y1) + (y2 * y2 + y3 * y3);
                        *slot = make_uint4(__float_as_uint(y0), __float_as_uint(y1), __float_as_uint(y2), __float_as_uint(y3));
                        if (j & 1) { pb[j >> 1].z = pack_bf16x2(y0, y1); pb[j >> 1].w = pack_bf16x2(y2, y3); }
                        else { pb[j >> 1].x = pack_bf16x2(y0, y1); pb[j >> 1].y = pack_bf16x2(y2, y3); }
                    }
                    uint8_t* brow = btile + lane * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(brow + (((4 * b + j) ^ sw) << 4)) = pb[j];
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmC, stg + b * STG_BYTES, colb + c, row0);          // x_new, fp32
                        if (b) tma_store_2d(&tmX, btile, colb + c - 32, row0);            // x_new, bf16 (64 columns)
                        tma_store_commit();
                    }
                    if (b) {  // per-row partial (sum, sum of squares) of this 64-column pair; the consumer adds them in order
                        if (ep.stats_out && row0 + lane < M)
                            ep.stats_out[static_cast<size_t>((colb + c) / STATS_COLS) * M + row0 + lane] = make_float2(rs, rq);
                        rs = 0.f;
                        rq = 0.f;
                    }
                }
            } else if (kOutF32) {
#pragma unroll 1
                for (int c = 0; c < SLICE_COLS; c += 32) {  // 32-column fp32 staging tiles (128 B rows)
                    float x[32];
                    uint4 packed[8];
                    load32(c, x);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        packed[j] = make_uint4(__float_as_uint(x[4 * j]), __float_as_uint(x[4 * j + 1]),
                                               __float_as_uint(x[4 * j + 2]), __float_as_uint(x[4 * j + 3]));
                    flush(packed, colb + c);
                }
            } else if (kAct == 2) {
                // SwiGLU (modeling_internlm2.py:275-278): W interleaves w1 / w3 rows, so accumulator columns (2i, 2i+1) are
                // (gate, up) of output column i: out = silu(gate) * up, half as wide as the accumulator. 128 accumulator
                // columns fill one 64-column bf16 staging tile of the [M, N/2] output.
                static_assert(kAct != 2 || SLICE_COLS % 128 == 0, "SwiGLU epilogue needs 128-column slices");
#pragma unroll 1
                for (int c = 0; c < SLICE_COLS; c += 128) {
                    uint4 packed[8];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        float x[32];
                        load32(c + 32 * h, x);
                        float y[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float g = x[2 * j];
                            y[j] = __fdividef(g, 1.0f + exp2f(-1.4426950408889634f * g)) * x[2 * j + 1];
                        }
                        packed[2 * h] = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]),
                                                   pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
                        packed[2 * h + 1] = make_uint4(pack_bf16x2(y[8], y[9]), pack_bf16x2(y[10], y[11]),
                                                       pack_bf16x2(y[12], y[13]), pack_bf16x2(y[14], y[15]));
                    }
                    flush(packed, (colb + c) >> 1);
                }
            } else if (kAct == 3) {
                // Rotary embedding (InternLM2 wqkv, 128-wide heads): this warp's 128-column slice is one head slot of a kv
                // group. Chunk pairs (0, 64) and (32, 96) hold (x[j], x[j + 64]): rotate in fp32, round once.
                static_assert(kAct != 3 || SLICE_COLS == 128, "rotary epilogue: one 128-wide head per slice");
                const bool rot = (colb % ep.rope_group_cols) < ep.rope_group_cols - 128;   // not the group's value head
                const int row = row0 + lane;
                const float4* cs = reinterpret_cast<const float4*>(ep.rope_tab + static_cast<size_t>(row < M ? row % ep.rope_N : 0) * 64);
                uint4 lo[8], hi[8];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float xa[32], xc[32];
                    load32(32 * h, xa);
                    load32(32 * h + 64, xc);
                    if (rot) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float4 t = __ldg(cs + ((32 * h + j) >> 1));   // (cos, sin) of j and j + 1
                            const float a0 = xa[j], a1 = xa[j + 1], b0 = xc[j], b1 = xc[j + 1];
                            xa[j] = a0 * t.x - b0 * t.y;
                            xc[j] = b0 * t.x + a0 * t.y;
                            xa[j + 1] = a1 * t.z - b1 * t.w;
                            xc[j + 1] = b1 * t.z + a1 * t.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        lo[4 * h + j] = make_uint4(pack_bf16x2(xa[8 * j], xa[8 * j + 1]), pack_bf16x2(xa[8 * j + 2], xa[8 * j + 3]),
                                                   pack_bf16x2(xa[8 * j + 4], xa[8 * j + 5]), pack_bf16x2(xa[8 * j + 6], xa[8 * j + 7]));
                        hi[4 * h + j] = make_uint4(pack_bf16x2(xc[8 * j], xc[8 * j + 1]), pack_bf16x2(xc[8 * j + 2], xc[8 * j + 3]),
                                                   pack_bf16x2(xc[8 * j + 4], xc[8 * j + 5]), pack_bf16x2(xc[8 * j + 6], xc[8 * j + 7]));
                    }
                }
                flush(lo, colb);
                flush(hi, colb + 64);
            } else {
#pragma unroll 1
                for (int c = 0; c < SLICE_COLS; c += 64) {  // 64-column bf16 staging tiles (128 B rows)
                    uint4 packed[8];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float x[32];
                        load32(c + 32 * h, x);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            packed[4 * h + j] = make_uint4(
                                pack_bf16x2(x[8 * j], x[8 * j + 1]), pack_bf16x2(x[8 * j + 2], x[8 * j + 3]),
                                pack_bf16x2(x[8 * j + 4], x[8 * j + 5]), pack_bf16x2(x[8 * j + 6], x[8 * j + 7]));
                    }
                    flush(packed, colb + c);
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (lane == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    cluster_sync_all();   // no CTA may exit (or free TMEM) while its peer can still touch its smem / barriers / TMEM
    if (warp == EPI_WARPS + 1) {
        tc_fence_after();
        tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    }
}

template <int kAct, bool kOutF32, int kEpiWarps, int kResid>
cudaError_t launch_variant(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmX,
                           int M, int N, int K, const GemmEpilogue& ep, int num_sms, cudaStream_t stream) {
    using C = Cfg<kEpiWarps, kResid>;
    auto kern = gemm_bf16_tn_kernel<kAct, kOutF32, kEpiWarps, kResid>;
    static unsigned long long optin_mask = 0;  // per device (kernels.h)
    cudaError_t e = ensure_smem_optin(reinterpret_cast<const void*>(kern), C::SMEM_BYTES, optin_mask);
    if (e != cudaSuccess) return e;
    const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
    int pairs = num_sms / 2;
    if (tiles < pairs) pairs = tiles;
    GemmEpilogue epk = ep;
    // W beyond what L2 keeps next to the activations, A smaller than W: walk M fastest so every W tile leaves DRAM once
    // (measured on the 7B w1|w3 GEMM, M = 2108: 2.12 GB of DRAM reads per launch with N fastest against 0.25 GB algorithmic).
    // With many row tiles (several prompts per prefill: A no longer sits in L2 next to the W stream) M goes fastest inside
    // groups of 16 row tiles (7B w1|w3 at four prompts per prefill: 26.6 against 28.0 ms per prompt; 8: 27.3, 4: 27.8).
    epk.group_m = gemm_group_m(M, N, K);
    return launch_pdl(kern, dim3(2 * pairs), dim3(C::THREADS), C::SMEM_BYTES, stream, tmA, tmB, tmC, tmX, M, N, K, epk);
}

}  // namespace

int gemm_group_m(int M, int N, int K) {
    const double w_bytes = 2.0 * N * K, a_bytes = 2.0 * M * K;
    const int num_m = (M + 2 * BM - 1) / (2 * BM);
    static const int group_big = [] { const char* e = getenv("AIGV_GEMM_GROUP_M"); return e ? atoi(e) : 16; }();
    if (w_bytes > 48e6 && a_bytes < w_bytes) return (num_m <= 20 || group_big <= 0) ? num_m : group_big;
    return 0;
}

int gemm_box_rows_a() { return BM; }
int gemm_box_rows_b() { return BNH; }
int gemm_box_cols() { return BK; }
int gemm_box_rows_c() { return 32; }
int gemm_box_cols_c(bool out_f32) { return out_f32 ? 32 : 64; }
int gemm_stats_slice_cols() { return STATS_COLS; }

cudaError_t launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, int M, int N, int K,
                        const GemmEpilogue& ep, int num_sms, cudaStream_t stream, const CUtensorMap* tmX) {
    if (M <= 0 || N <= 0 || K <= 0 || (N % 32) != 0 || (K % 8) != 0) return cudaErrorInvalidValue;
    if (ep.accumulate && !ep.out_f32) return cudaErrorInvalidValue;
    if (ep.fused_residual) {
        // x_new = x_old + colscale*(acc+bias) in the SM; needs the bf16 copy map and 64-column-aligned N
        if (ep.gelu || ep.accumulate || !tmX || (N % 64) != 0) return cudaErrorInvalidValue;
        if (ep.delta_out) {   // tmC = the bf16 update matrix d, tmX = the bf16 copy xb (read and rewritten in place)
            if (ep.out_f32 || ep.delta_in) return cudaErrorInvalidValue;
            return launch_variant<0, false, 8, 3>(tmA, tmB, tmC, *tmX, M, N, K, ep, num_sms, stream);
        }
        if (!ep.out_f32) return cudaErrorInvalidValue;
        if (ep.delta_in && K < 2048) return cudaErrorInvalidValue;   // only the serial (long-K) form folds a pending update
        // long K: the mainloop hides a serial single-buffer epilogue and keeps all 6 stages; short K: pipelined, 4 stages
        if (K >= 2048) return launch_variant<0, true, 8, 2>(tmA, tmB, tmC, *tmX, M, N, K, ep, num_sms, stream);
        // short K (proj): the epilogue is the critical path. The serial form is too slow for it even on 16 warps with
        // two chunks each (4.6 vs 4.2 ms/step, same-box A/B), so x_old is double-buffered at the price of 4 stages.
        return launch_variant<0, true, 8, 1>(tmA, tmB, tmC, *tmX, M, N, K, ep, num_sms, stream);
    }
    // The GELU epilogue is instruction-issue heavy: 16 epilogue warps and 5 stages; everything else 8 warps and 6 stages
    // (measured: the fp32 reduce-add epilogue of proj does not gain from more warps: 948 vs 1016 TFLOP/s).
    if (ep.swiglu) {  // tmC describes the [M, N/2] bf16 output
        if (ep.out_f32 || ep.gelu || ep.accumulate || ep.colscale || (N % 128) != 0) return cudaErrorInvalidValue;
        return launch_variant<2, false, 8, 0>(tmA, tmB, tmC, tmC, M, N, K, ep, num_sms, stream);
    }
    if (ep.gelu) {
#ifndef AIGV_GELU_WARPS
#define AIGV_GELU_WARPS 16
#endif
        return ep.out_f32 ? launch_variant<1, true, AIGV_GELU_WARPS, 0>(tmA, tmB, tmC, tmC, M, N, K, ep, num_sms, stream)
                          : launch_variant<1, false, AIGV_GELU_WARPS, 0>(tmA, tmB, tmC, tmC, M, N, K, ep, num_sms, stream);
    }
    if (ep.out_f32) return launch_variant<0, true, 8, 0>(tmA, tmB, tmC, tmC, M, N, K, ep, num_sms, stream);
    if (ep.rope_tab) {
        if (ep.rope_N <= 0 || ep.rope_group_cols < 256 || (ep.rope_group_cols % 128) != 0 || (N % ep.rope_group_cols) != 0 || ep.colscale ||
            ep.relu)
            return cudaErrorInvalidValue;
        return launch_variant<3, false, 8, 0>(tmA, tmB, tmC, tmC, M, N, K, ep, num_sms, stream);
    }
    return launch_variant<0, false, 8, 0>(tmA, tmB, tmC, tmC, M, N, K, ep, num_sms, stream);
}

}  // namespace aigv
