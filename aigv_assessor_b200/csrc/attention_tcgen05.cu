// Fused bidirectional flash-style attention on tcgen05/TMEM for sm_100a (head_dim 64).
//
// Replaces FlashAttention.forward -> flash_attn_varlen_qkvpacked_func (flash_attention.py:52-55, non-causal,
// dropout 0, softmax_scale = d^-0.5) reached from InternAttention._flash_attn (modeling_intern_vit.py:162-177), and is
// numerically the softmax(q*scale @ k^T) @ v of _naive_attn (modeling_intern_vit.py:143-157).
//
// One CTA = one (frame, head, 128-query block). Q/K/V tiles are read straight out of the [M, 3D] QKV GEMM output by
// TMA (no split / permute copies); channel order (three, head, d) as in modeling_intern_vit.py:164.
//   warp 4: TMA producer (Q once; K and V double buffered)     warp 5: MMA issuer (one thread)
//   warps 0-3: softmax, one query row per thread (TMEM lane == row, so row max / sum need no shuffles)
// Per KV block j:  S = Q K_j^T (UMMA 128 x kv x 64 -> TMEM)  ->  online softmax in registers (exp2, scale*log2e folded)
//   -> P_j (bf16) written to smem in the 128B-swizzled K-major layout  ->  O_j = P_j V_j (UMMA 128 x 64 x kv, V is the
//   MN-major B operand exactly as TMA lands it)  ->  threads fold O_j into their fp32 register accumulator with the
//   running rescale. S is double-buffered in TMEM (O_j reuses the dead S_j buffer), QK^T runs two blocks ahead and the
//   O fold one block behind, so the softmax warps do not wait on the tensor core in steady state.
//   S = 1025 = 8*128 + 1: the ragged tail is one short KV block (32 columns, 1 valid, rest masked).
// Two CTAs are resident per SM (256 TMEM columns and ~112 KiB smem each) so one CTA's softmax overlaps the other's MMAs.
#include "kernels.h"
#include "ptx.cuh"

namespace aigv {

namespace {

constexpr int BQ = 128;        // query rows per CTA
constexpr int BKV = 128;       // keys per full block
constexpr int HD = 64;         // head dim
constexpr int KV_STAGES = 2;
constexpr int TILE_BYTES = BKV * HD * 2;  // 16 KiB (also the Q tile and one P atom)
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
constexpr int OFF_P = OFF_V + KV_STAGES * TILE_BYTES;
constexpr int OFF_BAR = OFF_P + 2 * TILE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;  // 16 mbarriers + the TMEM base slot
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = 256;
// TMEM: two 128-column buffers; buffer j%2 holds S_j and, once consumed, O_j in its first 64 columns.

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_tail,
                     __nv_bfloat16* __restrict__ out, int S, int D, int tail_rows, float scale_log2) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;    // [2]  S_j landed in TMEM buffer j%2
    uint64_t* o_full = bars + 11;   // [2]  O_j = P_j V_j landed in TMEM buffer j%2 (over the dead S_j)
    uint64_t* o_empty = bars + 13;  // [2]  O_j folded into registers -> buffer j%2 may take S_{j+2}
    uint64_t* p_full = bars + 15;   //      S_j consumed and P_j written to smem
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

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
        tma_prefetch_desc(&tm_full);
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
                mbar_init(&s_full[s], 1);
                mbar_init(&o_full[s], 1);
                mbar_init(&o_empty[s], BQ);
            }
            mbar_init(p_full, BQ);
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
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
            tma_load_2d(smem + OFF_Q, &tm_full, q_full, head * HD, row_base + q0);
            for (int j = 0; j < n_blocks; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                const bool tail = j >= n_full;
                const uint32_t bytes = tail ? tail_rows * HD * 2 : TILE_BYTES;
                const CUtensorMap* tm = tail ? &tm_tail : &tm_full;
                const int r = row_base + j * BKV;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], bytes);
                tma_load_2d(smem + OFF_K + st * TILE_BYTES, tm, &k_full[st], D + head * HD, r);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[st], bytes);
                tma_load_2d(smem + OFF_V + st * TILE_BYTES, tm, &v_full[st], 2 * D + head * HD, r);
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------ MMA issuer
        // S_j goes to TMEM buffer j%2; once the softmax warps have consumed it, O_j = P_j V_j is written over its first
        // 64 columns; once O_j is folded, the buffer takes S_{j+2}. QK^T therefore runs two blocks ahead of the softmax
        // and P V one block behind it, so the softmax warps never wait on the tensor core in steady state.
        if (lane == 0) {
            const uint32_t q_addr = smem_u32(smem + OFF_Q);
            const uint32_t p_addr = smem_u32(smem + OFF_P);
            auto issue_qk = [&](int j) {
                const int st = j & 1;
                const int kv = j >= n_full ? tail_rows : BKV;
                mbar_wait(&k_full[st], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(smem + OFF_K + st * TILE_BYTES);
                const uint32_t idesc = umma_idesc_bf16(BQ, kv, 0, 0);
                const uint32_t tm_s = tmem_base + st * BKV;
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    umma_bf16_ss(tm_s, umma_desc_sw128(q_addr + k * 32, 16, 1024),
                                 umma_desc_sw128(k_addr + k * 32, 16, 1024), idesc, k != 0 ? 1u : 0u);
                }
                umma_commit(&s_full[st]);
                umma_commit(&k_empty[st]);
            };
            mbar_wait(q_full, 0);
            issue_qk(0);
            if (n_blocks > 1) issue_qk(1);
            for (int j = 0; j < n_blocks; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                const int kv = j >= n_full ? tail_rows : BKV;
                mbar_wait(p_full, j & 1);          // S_j consumed, P_j in smem
                mbar_wait(&v_full[st], ph);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(smem + OFF_V + st * TILE_BYTES);
                const uint32_t tm_o = tmem_base + st * BKV;
                constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, 0, 1);  // B (= V) is MN-major
                for (int k = 0; k < kv / 16; ++k) {
                    // A = P: K-major, two 64-column swizzle atoms of 16 KiB; B = V rows [16k, 16k+16): 2 KiB apart.
                    const uint32_t pa = p_addr + (k >> 2) * TILE_BYTES + (k & 3) * 32;
                    umma_bf16_ss(tm_o, umma_desc_sw128(pa, 16, 1024), umma_desc_sw128(v_addr + k * 2048, 16, 1024),
                                 idesc_pv, k != 0 ? 1u : 0u);
                }
                umma_commit(&o_full[st]);
                umma_commit(&v_empty[st]);
                if (j + 2 < n_blocks) {
                    mbar_wait(&o_empty[st], ph);   // O_j folded -> buffer free for S_{j+2}
                    issue_qk(j + 2);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ softmax warps: one query row per thread
        const int r = threadIdx.x;  // 0..127 == TMEM lane
        const uint32_t tm_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        uint8_t* p_row = smem + OFF_P + r * 128;
        const int sw = r & 7;
        float m_run = -INFINITY, l_run = 0.f;
        float alpha_prev = 1.f;  // rescale that goes with the not-yet-folded O_{j-1}
        float o_acc[HD];
#pragma unroll
        for (int i = 0; i < HD; ++i) o_acc[i] = 0.f;
        // Warps whose 32 query rows are all beyond the sequence (only in the ragged last query block) skip the softmax
        // math; they still take part in the barrier protocol so the MMA warp's schedule is unchanged.
        const bool warp_live = q0 + warp * 32 < S;

        auto fold = [&](int jj, float a) {  // o_acc = o_acc * a + O_jj
            const int st = jj & 1;
            mbar_wait(&o_full[st], (jj >> 1) & 1);
            tc_fence_after();
            if (warp_live) {
#pragma unroll
                for (int c = 0; c < HD; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(tm_lane + st * BKV + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o_acc[c + i] = fmaf(o_acc[c + i], a, __uint_as_float(v[i]));
                }
            }
            tc_fence_before();
            mbar_arrive(&o_empty[st]);
        };

        for (int j = 0; j < n_blocks; ++j) {
            const bool tail = j >= n_full;
            const int kv = tail ? tail_rows : BKV;
            const uint32_t tm_s = tm_lane + (j & 1) * BKV;
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            float alpha = 1.f, m_scaled = 0.f;
            if (warp_live) {
                // pass 1: row maximum (four independent chains)
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
                if (!tail) {
#pragma unroll 1
                    for (int c = 0; c < BKV; c += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(tm_s + c, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            mx0 = fmaxf(mx0, __uint_as_float(v[i]));
                            mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
                            mx2 = fmaxf(mx2, __uint_as_float(v[i + 2]));
                            mx3 = fmaxf(mx3, __uint_as_float(v[i + 3]));
                        }
                    }
                } else {
                    for (int c = 0; c < kv; c += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(tm_s + c, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            mx0 = fmaxf(mx0, (c + i < tail_valid) ? __uint_as_float(v[i]) : -INFINITY);
                    }
                }
                const float m_new = fmaxf(fmaxf(m_run, fmaxf(mx0, mx1)), fmaxf(mx2, mx3));
                alpha = ex2((m_run - m_new) * scale_log2);
                m_scaled = m_new * scale_log2;
                m_run = m_new;
            }
            // fold the previous block's O (its P V ran while pass 1 was executing); frees P smem and the TMEM buffer
            if (j > 0) fold(j - 1, alpha_prev);
            alpha_prev = alpha;
            if (warp_live) {
                // pass 2: probabilities -> bf16 P tile in smem (swizzled K-major), row sum (four independent chains)
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                auto emit = [&](const float (&p)[32], int c) {
                    uint8_t* atom = p_row + (c >> 6) * TILE_BYTES;
                    const int chunk0 = (c & 63) >> 3;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 w;
                        w.x = pack_bf16x2(p[8 * q + 0], p[8 * q + 1]);
                        w.y = pack_bf16x2(p[8 * q + 2], p[8 * q + 3]);
                        w.z = pack_bf16x2(p[8 * q + 4], p[8 * q + 5]);
                        w.w = pack_bf16x2(p[8 * q + 6], p[8 * q + 7]);
                        *reinterpret_cast<uint4*>(atom + (((chunk0 + q) ^ sw) << 4)) = w;
                    }
                };
                if (!tail) {
#pragma unroll 1
                    for (int c = 0; c < BKV; c += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(tm_s + c, v);
                        tmem_ld_wait();
                        float p[32];
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            p[i] = ex2(fmaf(__uint_as_float(v[i]), scale_log2, -m_scaled));
                            p[i + 1] = ex2(fmaf(__uint_as_float(v[i + 1]), scale_log2, -m_scaled));
                            p[i + 2] = ex2(fmaf(__uint_as_float(v[i + 2]), scale_log2, -m_scaled));
                            p[i + 3] = ex2(fmaf(__uint_as_float(v[i + 3]), scale_log2, -m_scaled));
                            s0 += p[i]; s1 += p[i + 1]; s2 += p[i + 2]; s3 += p[i + 3];
                        }
                        emit(p, c);
                    }
                } else {
                    for (int c = 0; c < kv; c += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(tm_s + c, v);
                        tmem_ld_wait();
                        float p[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float e = ex2(fmaf(__uint_as_float(v[i]), scale_log2, -m_scaled));
                            p[i] = (c + i < tail_valid) ? e : 0.f;
                            s0 += p[i];
                        }
                        emit(p, c);
                    }
                }
                l_run = l_run * alpha + ((s0 + s1) + (s2 + s3));
                fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            }
            tc_fence_before();
            mbar_arrive(p_full);
        }
        fold(n_blocks - 1, alpha_prev);

        const int q = q0 + r;
        if (q < S) {
            const float inv = 1.0f / l_run;
            __nv_bfloat16* o = out + static_cast<size_t>(row_base + q) * D + head * HD;
#pragma unroll
            for (int c = 0; c < HD; c += 8) {
                uint4 w;
                w.x = pack_bf16x2(o_acc[c + 0] * inv, o_acc[c + 1] * inv);
                w.y = pack_bf16x2(o_acc[c + 2] * inv, o_acc[c + 3] * inv);
                w.z = pack_bf16x2(o_acc[c + 4] * inv, o_acc[c + 5] * inv);
                w.w = pack_bf16x2(o_acc[c + 6] * inv, o_acc[c + 7] * inv);
                *reinterpret_cast<uint4*>(o + c) = w;
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

int attention_tail_rows(int S) {
    const int t = S % BKV;
    return t == 0 ? 32 : ((t + 31) / 32) * 32;
}

cudaError_t launch_attention(const CUtensorMap& tm_full, const CUtensorMap& tm_tail, __nv_bfloat16* out, int frames,
                             int S, int heads, int head_dim, float softmax_scale, cudaStream_t stream) {
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
    attention_fwd_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tm_full, tm_tail, out, S, heads * head_dim,
                                                                attention_tail_rows(S), scale_log2);
    return cudaGetLastError();
}

}  // namespace aigv
