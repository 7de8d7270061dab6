// Fused bidirectional flash-style attention on tcgen05/TMEM for sm_100a (head_dim 64).
//
// Replaces FlashAttention.forward -> flash_attn_varlen_qkvpacked_func (flash_attention.py:52-55, non-causal,
// dropout 0, softmax_scale = d^-0.5) reached from InternAttention._flash_attn (modeling_intern_vit.py:162-177), and is
// numerically the softmax(q*scale @ k^T) @ v of _naive_attn (modeling_intern_vit.py:143-157).
//
// One CTA = one (frame, head, 128-query block). Q/K/V tiles are read straight out of the [M, 3D] QKV GEMM output by
// TMA (no split / permute copies); channel order (three, head, d) as in modeling_intern_vit.py:164.
//   warp 4: TMA producer (Q once; K and V double buffered, 64 keys per block)     warp 5: MMA issuer (one thread)
//   warps 0-3: softmax, one query row per thread (TMEM lane == row, so row max / sum need no shuffles)
// Per 64-key block j:  S = Q K_j^T (UMMA 128x64x64 -> TMEM cols [0,64))  ->  one tcgen05.ld pass pulls the 64-key row
//   into registers (row max, p = exp2(s*scale*log2e - m_ref)) -> P_j (bf16) written to smem, 128B-swizzled K-major ->
//   O += P_j V_j (UMMA 128x64x64 accumulating in TMEM cols [64,128); V is the MN-major B operand exactly as TMA lands it).
// O lives in TMEM for the whole KV loop. m_ref is a *stale* row maximum: it is only raised (and O, l rescaled through
// tcgen05.ld/st) when the true running maximum exceeds it by more than 2^8, so p <= 256 always and the rescale is rare;
// the final O / l is exact regardless of the reference point.
// The kernel is MUFU(ex2)-bound, so residency matters more than tile size: 64 KiB smem, 128 TMEM columns and <=112
// registers give 3 CTAs (12 softmax warps) per SM, which lets other CTAs' exp work hide each CTA's barrier round trips.
// S = 1025 = 16*64 + 1: the ragged tail is one short KV block (32 columns, 1 valid, rest masked to -inf).
#include "kernels.h"
#include "ptx.cuh"

namespace aigv {

namespace {

constexpr int BQ = 128;   // query rows per CTA
constexpr int BKV = 64;   // keys per full block
constexpr int HD = 64;    // head dim
constexpr int Q_BYTES = BQ * HD * 2;    // 16 KiB
constexpr int KV_BYTES = BKV * HD * 2;  // 8 KiB
constexpr int P_BYTES = BQ * BKV * 2;   // 16 KiB (one 128B-swizzle atom column)
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;        // 2 stages
constexpr int OFF_V = OFF_K + 2 * KV_BYTES;   // 2 stages
constexpr int OFF_P = OFF_V + 2 * KV_BYTES;
constexpr int OFF_BAR = OFF_P + P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;     // 11 mbarriers + the TMEM base slot
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = 128;
constexpr uint32_t TM_S = 0;    // S block: columns [0,64)
constexpr uint32_t TM_O = 64;   // O accumulator: columns [64,128)
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#ifndef AIGV_ATTN_MIN_CTAS
#define AIGV_ATTN_MIN_CTAS 2
#endif
__global__ void __launch_bounds__(THREADS, AIGV_ATTN_MIN_CTAS)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     const __grid_constant__ CUtensorMap tm_tail, __nv_bfloat16* __restrict__ out, int S, int D,
                     int tail_rows, float scale_log2) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;    // S_j in TMEM (also: every earlier MMA, incl. P_{j-1} V_{j-1}, has completed)
    uint64_t* p_full = bars + 10;   // S_j consumed, P_j written to smem, O rescaled if needed
    uint64_t* o_done = bars + 11;   // last P V accumulated
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ;
    const int head = blockIdx.y;
    const int frame = blockIdx.z;
    const int row_base = frame * S;          // first token row of this frame in the [M, 3D] matrix
    const int n_full = S / BKV;              // full KV blocks
    const int tail_valid = S - n_full * BKV; // valid keys in the tail block (0 = no tail)
    const int n_blocks = n_full + (tail_valid > 0 ? 1 : 0);

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("aigv attention: dynamic smem base not 1024B aligned\n");
        __trap();
    }
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        tma_prefetch_desc(&tm_tail);
    }
    if (warp == 5) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            for (int s = 0; s < 2; ++s) {
                mbar_init(&k_full[s], 1);
                mbar_init(&k_empty[s], 1);
                mbar_init(&v_full[s], 1);
                mbar_init(&v_empty[s], 1);
            }
            mbar_init(s_full, 1);
            mbar_init(p_full, BQ);
            mbar_init(o_done, 1);
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

    if (warp == 4) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, Q_BYTES);
            tma_load_2d(smem + OFF_Q, &tm_q, q_full, head * HD, row_base + q0);
            for (int j = 0; j < n_blocks; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                const bool tail = j >= n_full;
                const uint32_t bytes = tail ? tail_rows * HD * 2 : KV_BYTES;
                const CUtensorMap* tm = tail ? &tm_tail : &tm_kv;
                const int r = row_base + j * BKV;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], bytes);
                tma_load_2d(smem + OFF_K + st * KV_BYTES, tm, &k_full[st], D + head * HD, r);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[st], bytes);
                tma_load_2d(smem + OFF_V + st * KV_BYTES, tm, &v_full[st], 2 * D + head * HD, r);
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t q_addr = smem_u32(smem + OFF_Q);
            const uint32_t p_addr = smem_u32(smem + OFF_P);
            const uint32_t tm_s = tmem_base + TM_S;
            const uint32_t tm_o = tmem_base + TM_O;
            auto issue_qk = [&](int j) {
                const int st = j & 1;
                const int kv = j >= n_full ? tail_rows : BKV;
                mbar_wait(&k_full[st], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(smem + OFF_K + st * KV_BYTES);
                const uint32_t idesc = umma_idesc_bf16(BQ, kv, 0, 0);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    umma_bf16_ss(tm_s, umma_desc_sw128(q_addr + k * 32, 16, 1024),
                                 umma_desc_sw128(k_addr + k * 32, 16, 1024), idesc, k != 0 ? 1u : 0u);
                }
                umma_commit(s_full);
                umma_commit(&k_empty[st]);
            };
            mbar_wait(q_full, 0);
            issue_qk(0);
            for (int j = 0; j < n_blocks; ++j) {
                const int st = j & 1;
                const int kv = j >= n_full ? tail_rows : BKV;
                mbar_wait(p_full, j & 1);  // S_j consumed (buffer free), P_j in smem, O rescaled if it had to be
                mbar_wait(&v_full[st], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(smem + OFF_V + st * KV_BYTES);
                constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, 0, 1);  // B (= V) is MN-major
                for (int k = 0; k < kv / 16; ++k) {
                    // A = P_j: K-major 128B-swizzled rows, 32 B per 16-key step; B = V rows [16k, 16k+16): 2 KiB apart.
                    umma_bf16_ss(tm_o, umma_desc_sw128(p_addr + k * 32, 16, 1024),
                                 umma_desc_sw128(v_addr + k * 2048, 16, 1024), idesc_pv, (j | k) != 0 ? 1u : 0u);
                }
                umma_commit(&v_empty[st]);
                if (j + 1 < n_blocks) issue_qk(j + 1);  // its s_full commit also covers P_j V_j
                else umma_commit(o_done);
            }
        }
    } else {
        // ------------------------------------------------------------ softmax warps: one query row per thread
        const int r = threadIdx.x;  // 0..127 == TMEM lane
        const uint32_t tm_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        const uint32_t tm_s = tm_lane + TM_S;
        const uint32_t tm_o = tm_lane + TM_O;
        uint8_t* p_row = smem + OFF_P + r * 128;
        const int sw = r & 7;
        float m_ref = -INFINITY;  // reference maximum, already multiplied by scale*log2e
        float l_run = 0.f;
        // Warps whose 32 query rows are all beyond the sequence (only in the ragged last query block) skip the softmax
        // math; they still take part in the barrier protocol so the MMA warp's schedule is unchanged.
        const bool warp_live = q0 + warp * 32 < S;

        for (int j = 0; j < n_blocks; ++j) {
            const bool tail = j >= n_full;
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            if (warp_live) {
                // Single pass over S: the whole 64-key row is pulled into registers once (TMEM reads, ~64 B/clk/SM, are
                // as scarce as MUFU throughput here), then max -> (rare) rescale -> exp2 -> bf16 P tile.
                const int ncol = tail ? tail_rows : BKV;
                uint32_t va[32], vb[32];  // columns [0,32) and [32,64) of this thread's row
                tmem_ld_32x32(tm_s, va);
                if (ncol > 32) tmem_ld_32x32(tm_s + 32, vb);
                tmem_ld_wait();
                if (tail) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (i >= tail_valid) va[i] = 0xff800000u;  // -inf: exp2 gives exactly 0
                        if (i + 32 >= tail_valid) vb[i] = 0xff800000u;
                    }
                }
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    mx0 = fmaxf(mx0, fmaxf(__uint_as_float(va[i]), __uint_as_float(vb[i])));
                    mx1 = fmaxf(mx1, fmaxf(__uint_as_float(va[i + 1]), __uint_as_float(vb[i + 1])));
                    mx2 = fmaxf(mx2, fmaxf(__uint_as_float(va[i + 2]), __uint_as_float(vb[i + 2])));
                    mx3 = fmaxf(mx3, fmaxf(__uint_as_float(va[i + 3]), __uint_as_float(vb[i + 3])));
                }
                const float mxs = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
                const bool need = mxs > m_ref + RESCALE_THRESHOLD;  // always true for the first block (m_ref = -inf)
                if (__any_sync(0xffffffffu, need)) {                 // warp-uniform: TMEM ops are warp-collective
                    const float new_ref = need ? mxs : m_ref;
                    if (j > 0) {
                        // s_full(j) implies P_{j-1} V_{j-1} has completed, so O is quiescent: rescale it in place
                        const float alpha = ex2(m_ref - new_ref);  // 1 for rows that keep their reference
#pragma unroll 1
                        for (int c = 0; c < HD; c += 32) {
                            uint32_t o[32];
                            tmem_ld_32x32(tm_o + c, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st_32x32(tm_o + c, o);
                        }
                        tmem_st_wait();
                        l_run *= alpha;
                    }
                    m_ref = new_ref;
                }
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                auto emit8 = [&](const uint32_t (&v)[32], int i0, int chunk) {
                    float p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) p[i] = ex2(fmaf(__uint_as_float(v[i0 + i]), scale_log2, -m_ref));
                    s0 += p[0] + p[4]; s1 += p[1] + p[5]; s2 += p[2] + p[6]; s3 += p[3] + p[7];
                    uint4 w;
                    w.x = pack_bf16x2(p[0], p[1]);
                    w.y = pack_bf16x2(p[2], p[3]);
                    w.z = pack_bf16x2(p[4], p[5]);
                    w.w = pack_bf16x2(p[6], p[7]);
                    *reinterpret_cast<uint4*>(p_row + ((chunk ^ sw) << 4)) = w;
                };
#pragma unroll
                for (int c = 0; c < 4; ++c) emit8(va, 8 * c, c);
                if (ncol > 32) {  // the tail block only carries tail_rows columns
#pragma unroll
                    for (int c = 0; c < 4; ++c) emit8(vb, 8 * c, 4 + c);
                }
                l_run += (s0 + s1) + (s2 + s3);
                fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            }
            tc_fence_before();
            mbar_arrive(p_full);
        }

        mbar_wait(o_done, 0);
        tc_fence_after();
        const int q = q0 + r;
        if (warp_live) {
            const float inv = 1.0f / l_run;
            __nv_bfloat16* o = out + static_cast<size_t>(row_base + q) * D + head * HD;
#pragma unroll
            for (int c = 0; c < HD; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tm_o + c, v);
                tmem_ld_wait();
                if (q < S) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 w;
                        w.x = pack_bf16x2(__uint_as_float(v[i + 0]) * inv, __uint_as_float(v[i + 1]) * inv);
                        w.y = pack_bf16x2(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
                        w.z = pack_bf16x2(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
                        w.w = pack_bf16x2(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
                        *reinterpret_cast<uint4*>(o + c + i) = w;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

}  // namespace

int attention_q_rows() { return BQ; }
int attention_kv_rows() { return BKV; }
int attention_tail_rows(int S) {
    const int t = S % BKV;
    return t == 0 ? 32 : ((t + 31) / 32) * 32;
}

cudaError_t launch_attention(const CUtensorMap& tm_q, const CUtensorMap& tm_kv, const CUtensorMap& tm_tail,
                             __nv_bfloat16* out, int frames, int S, int heads, int head_dim, float softmax_scale,
                             cudaStream_t stream) {
    if (head_dim != HD || frames <= 0 || S <= 0 || heads <= 0) return cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) {
        cudaError_t e =
            cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((S + BQ - 1) / BQ, heads, frames);
    const float scale_log2 = softmax_scale * 1.4426950408889634f;
    attention_fwd_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tm_q, tm_kv, tm_tail, out, S, heads * head_dim,
                                                                attention_tail_rows(S), scale_log2);
    return cudaGetLastError();
}

}  // namespace aigv
