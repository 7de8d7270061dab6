// Fused bidirectional flash-style attention on tcgen05/TMEM for sm_100a (head_dim 64).
//
// Replaces FlashAttention.forward -> flash_attn_varlen_qkvpacked_func (flash_attention.py:52-55, non-causal,
// dropout 0, softmax_scale = d^-0.5) reached from InternAttention._flash_attn (modeling_intern_vit.py:162-177), and is
// numerically the softmax(q*scale @ k^T) @ v of _naive_attn (modeling_intern_vit.py:143-157).
//
// One CTA = one (frame, head) x TWO 128-query tiles. Q/K/V tiles are read straight out of the [M, 3D] QKV GEMM output
// by TMA (no split / permute copies); channel order (three, head, d) as in modeling_intern_vit.py:164. Both query tiles
// consume every K/V stage, which halves the K/V tile traffic from L2 (measured to be the first bound of a
// one-tile-per-CTA kernel: ~5.9 TB/s of tile loads with the softmax math removed).
//   warp 8: TMA producer (Q tiles once; K and V double buffered, 64 keys per block)   warp 9: MMA issuer (one thread)
//   warps 0-3 / 4-7: softmax warpgroup of query tile 0 / 1, one query row per thread (TMEM lane == row, so row max /
//   sum need no shuffles). The two warpgroups ping-pong against the single MMA thread: while one does its exp work the
//   tensor core runs the other's P V and next Q K^T.
// Per tile and 64-key block j:  S = Q K_j^T (UMMA 128x64x64 -> TMEM)  ->  row max, p = exp2(s*scale*log2e - m_ref)
//   -> P_j (bf16 pairs) written back over S in TENSOR memory (tcgen05.st)  ->  O += P_j V_j (UMMA 128x64x64 with the A
//   operand read from TMEM, accumulating in TMEM; V is the MN-major B operand exactly as TMA lands it). Keeping P out of
//   shared memory removes the per-block st.shared + proxy fence and a third of the tensor pipe's smem reads (a 128x64x16
//   MMA with both operands in smem is smem-bandwidth bound: measured 48 instead of 32 cycles).
// O lives in TMEM for the whole KV loop. m_ref is a *stale* row maximum: it is only raised (and O, l rescaled through
// tcgen05.ld/st) when the true running maximum exceeds it by more than 2^8, so p <= 256 always and the rescale is rare;
// the final O / l is exact regardless of the reference point.
// Resources: 64 KiB smem, 256 TMEM columns, <=102 registers -> 2 CTAs (16 softmax warps) per SM.
// S = 1025 = 16*64 + 1: the ragged tail is one short KV block (32 columns, 1 valid, rest masked to -inf); on the query
// side the last CTA of a (frame, head) holds the single leftover row and skips its empty second tile entirely.
#include "kernels.h"
#include "ptx.cuh"

namespace aigv {

namespace {

constexpr int BQ = 128;    // query rows per tile
constexpr int TILES = 2;   // query tiles per CTA
constexpr int BKV = 64;    // keys per full block
constexpr int SOFTMAX_WARPS = 4 * TILES;
constexpr int THREADS = (SOFTMAX_WARPS + 2) * 32;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

// Geometry per head dimension. Every operand tile is stored as 64-column (128-byte) swizzle atoms: HD/64 atoms side by
// side for Q, K (K-major: the d axis is the MMA K dimension) and V (MN-major: the d axis is the MMA N dimension).
template <int HD>
struct Geo {
    static constexpr int ATOMS = HD / 64;
    static constexpr int Q_ATOM = BQ * 128;                 // bytes of one 64-column atom of a Q tile (16 KiB)
    static constexpr int KV_ATOM = BKV * 128;               // bytes of one 64-column atom of a K / V tile (8 KiB)
    static constexpr int Q_BYTES = ATOMS * Q_ATOM;
    static constexpr int KV_BYTES = ATOMS * KV_ATOM;
    static constexpr int STAGES = HD == 64 ? 4 : 2;         // K/V ring depth (TMA latency ~ two block times)
    static constexpr int OFF_Q = 0;                         // TILES tiles
    static constexpr int OFF_K = OFF_Q + TILES * Q_BYTES;   // STAGES stages
    static constexpr int OFF_V = OFF_K + STAGES * KV_BYTES; // STAGES stages
    static constexpr int OFF_BAR = OFF_V + STAGES * KV_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;        // 1 + 4*STAGES + 3*TILES mbarriers + the TMEM base slot
    // TMEM columns per query tile: S block at +0 (64 fp32 cols) -- overwritten in place by P (bf16 pairs, 32 cols) once
    // the softmax has pulled S into registers -- and the O accumulator (HD fp32 cols) at +64.
    static constexpr uint32_t TM_O = 64;
    static constexpr uint32_t TM_TILE = HD == 64 ? 128 : 256;
    static constexpr uint32_t TMEM_COLS = TILES * TM_TILE;  // 256 (two CTAs per SM) or 512 (one)
    static constexpr int MIN_CTAS = HD == 64 ? 2 : 1;
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD>
__global__ void __launch_bounds__(THREADS, Geo<HD>::MIN_CTAS)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     const __grid_constant__ CUtensorMap tm_tail, __nv_bfloat16* __restrict__ out, int S, int D,
                     int tail_rows, float scale_log2) {
    using G = Geo<HD>;
    constexpr int ATOMS = G::ATOMS, Q_ATOM = G::Q_ATOM, KV_ATOM = G::KV_ATOM, Q_BYTES = G::Q_BYTES,
                  KV_BYTES = G::KV_BYTES, OFF_Q = G::OFF_Q, OFF_K = G::OFF_K, OFF_V = G::OFF_V, OFF_BAR = G::OFF_BAR;
    constexpr uint32_t TM_O = G::TM_O, TM_TILE = G::TM_TILE, TMEM_COLS = G::TMEM_COLS;
    constexpr int STAGES = G::STAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = q_full + 1;           // [STAGES]
    uint64_t* k_empty = k_full + STAGES;     // [STAGES]
    uint64_t* v_full = k_empty + STAGES;     // [STAGES]
    uint64_t* v_empty = v_full + STAGES;     // [STAGES]
    uint64_t* s_full = v_empty + STAGES;     // [TILES] S_j in TMEM (also: every earlier MMA, incl. this tile's P V, done)
    uint64_t* p_full = s_full + TILES;       // [TILES] S_j consumed, P_j written to TMEM, O rescaled if needed (1 arrival/warp)
    uint64_t* o_done = p_full + TILES;       // [TILES] last P V accumulated
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + TILES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (BQ * TILES);
    const int head = blockIdx.y;
    const int frame = blockIdx.z;
    const int row_base = frame * S;          // first token row of this frame in the [M, 3D] matrix
    const int n_full = S / BKV;              // full KV blocks
    const int tail_valid = S - n_full * BKV; // valid keys in the tail block (0 = no tail)
    const int n_blocks = n_full + (tail_valid > 0 ? 1 : 0);
    const int n_tiles = (q0 + BQ < S) ? TILES : 1;  // tiles of this CTA that hold at least one query row

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("aigv attention: dynamic smem base not 1024B aligned\n");
        __trap();
    }
    if (warp == SOFTMAX_WARPS && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
        tma_prefetch_desc(&tm_tail);
    }
    if (warp == SOFTMAX_WARPS + 1) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&k_full[s], 1);
                mbar_init(&k_empty[s], 1);
                mbar_init(&v_full[s], 1);
                mbar_init(&v_empty[s], 1);
            }
            for (int t = 0; t < TILES; ++t) {
                mbar_init(&s_full[t], 1);
                mbar_init(&p_full[t], 4);
                mbar_init(&o_done[t], 1);
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

    if (warp == SOFTMAX_WARPS) {
        // ------------------------------------------------------------ TMA producer
        {
            if (elect_one()) {
                mbar_arrive_expect_tx(q_full, n_tiles * Q_BYTES);
                for (int t = 0; t < n_tiles; ++t)
                    for (int a = 0; a < ATOMS; ++a)
                        tma_load_2d(smem + OFF_Q + t * Q_BYTES + a * Q_ATOM, &tm_q, q_full, head * HD + a * 64,
                                    row_base + q0 + t * BQ);
            }
            __syncwarp();
            for (int j = 0; j < n_blocks; ++j) {
                const int st = j % STAGES;
                const uint32_t ph = (j / STAGES) & 1;
                const bool tail = j >= n_full;
                const uint32_t bytes = tail ? tail_rows * HD * 2 : KV_BYTES;
                const int atom_stride = tail ? tail_rows * 128 : KV_ATOM;  // the tail box is shorter, atoms pack tighter
                const CUtensorMap* tm = tail ? &tm_tail : &tm_kv;
                const int r = row_base + j * BKV;
                mbar_wait(&k_empty[st], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&k_full[st], bytes);
                    for (int a = 0; a < ATOMS; ++a)
                        tma_load_2d(smem + OFF_K + st * KV_BYTES + a * atom_stride, tm, &k_full[st],
                                    D + head * HD + a * 64, r);
                }
                __syncwarp();
                mbar_wait(&v_empty[st], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&v_full[st], bytes);
                    for (int a = 0; a < ATOMS; ++a)
                        tma_load_2d(smem + OFF_V + st * KV_BYTES + a * atom_stride, tm, &v_full[st],
                                    2 * D + head * HD + a * 64, r);
                }
                __syncwarp();
            }
        }
    } else if (warp == SOFTMAX_WARPS + 1) {
        // ------------------------------------------------------------ MMA issuer
        // Per block j and tile t (in order): wait P_j(t) -> O(t) += P_j V_j -> S(t) = Q(t) K_{j+1}^T. A K/V stage is
        // released by the commit that follows its last reader (the last live tile).
        {
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            // Base descriptors are built once; every MMA step derives its operands with one add (the descriptor
            // bit-twiddling per UTCHMMA was most of this warp's instruction stream, and this warp is the critical path).
            const uint64_t q_desc0 = umma_desc_sw128(smem_u32(smem + OFF_Q), 16, 1024);
            const uint64_t k_desc0 = umma_desc_sw128(smem_u32(smem + OFF_K), 16, 1024);
            const uint64_t v_desc0 = umma_desc_sw128(smem_u32(smem + OFF_V), KV_ATOM, 1024);
            const uint64_t v_desc0_tail = umma_desc_sw128(smem_u32(smem + OFF_V), tail_rows * 128, 1024);
            constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, 0, 0);
            const uint32_t idesc_qk_tail = umma_idesc_bf16(BQ, tail_rows, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, 0, 1);  // B (= V) is MN-major, N = HD
            // Q K^T of block j for tile t (operands must already be resident); called inside an elected region.
            auto qk_mmas = [&](int t, int st, bool tail) {
                const uint64_t qd = umma_desc_advance(q_desc0, t * Q_BYTES);
                const uint64_t kd = umma_desc_advance(k_desc0, st * KV_BYTES);
                const uint32_t k_atom = tail ? tail_rows * 128 : KV_ATOM;
                const uint32_t idesc = tail ? idesc_qk_tail : idesc_qk;
                const uint32_t tm_s = tmem_u + t * TM_TILE;
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {  // d axis: 4 steps of 32 B inside each 64-column atom
                    umma_bf16_ss(tm_s, umma_desc_advance(qd, (k >> 2) * Q_ATOM + (k & 3) * 32),
                                 umma_desc_advance(kd, (k >> 2) * k_atom + (k & 3) * 32), idesc, k != 0 ? 1u : 0u);
                }
                umma_commit(&s_full[t]);
                if (t == n_tiles - 1) umma_commit(&k_empty[st]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            if (elect_one()) {
                for (int t = 0; t < n_tiles; ++t) qk_mmas(t, 0, n_full == 0);
            }
            __syncwarp();
            for (int j = 0; j < n_blocks; ++j) {
                const int st = j % STAGES;
                const int st1 = (j + 1) % STAGES;
                const bool tail = j >= n_full;
                const bool more = j + 1 < n_blocks;
                const bool tail1 = j + 1 >= n_full;
                const int ksteps = (tail ? tail_rows : BKV) / 16;
                const uint64_t vd = umma_desc_advance(tail ? v_desc0_tail : v_desc0, st * KV_BYTES);
                for (int t = 0; t < n_tiles; ++t) {
                    mbar_wait(&p_full[t], j & 1);  // S_j(t) consumed, P_j(t) in TMEM, O(t) rescaled if it had to be
                    if (t == 0) {
                        mbar_wait(&v_full[st], (j / STAGES) & 1);
                        if (more) mbar_wait(&k_full[st1], ((j + 1) / STAGES) & 1);
                    }
                    tc_fence_after();
                    const uint32_t tm_p = tmem_u + t * TM_TILE;  // P_j: bf16 pairs in the first 32 columns of the S block
                    const uint32_t tm_o = tm_p + TM_O;
                    // one elected region per (tile, block): O += P_j V_j, its commits, then S = Q K_{j+1}^T whose
                    // s_full commit also covers the P V just issued
                    if (elect_one()) {
                        if (!tail) {
#pragma unroll
                            for (int k = 0; k < BKV / 16; ++k)
                                // A = P_j from TMEM: 8 columns (16 keys) per step; B = V rows [16k,16k+16): 2 KiB apart
                                umma_bf16_ts(tm_o, tm_p + k * 8, umma_desc_advance(vd, k * 2048), idesc_pv,
                                             (j | k) != 0 ? 1u : 0u);
                        } else {
                            for (int k = 0; k < ksteps; ++k)
                                umma_bf16_ts(tm_o, tm_p + k * 8, umma_desc_advance(vd, k * 2048), idesc_pv,
                                             (j | k) != 0 ? 1u : 0u);
                        }
                        if (t == n_tiles - 1) umma_commit(&v_empty[st]);
                        if (more) qk_mmas(t, st1, tail1);
                        else umma_commit(&o_done[t]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ------------------------------------------------------------ softmax warpgroups: one query row per thread
        const int t = warp >> 2;                      // query tile of this warpgroup
        const int wq = warp & 3;                      // TMEM lane quadrant
        const int r = wq * 32 + lane;                 // row within the tile == TMEM lane
        const int tile_q0 = q0 + t * BQ;
        if (tile_q0 < S) {
            const uint32_t tm_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + t * TM_TILE;
            const uint32_t tm_s = tm_lane;
            const uint32_t tm_o = tm_lane + TM_O;
            float m_ref = -INFINITY;  // reference maximum, already multiplied by scale*log2e
            float l_run = 0.f;
            // Warps whose 32 query rows are all beyond the sequence (ragged last tile) skip the softmax math; they
            // still take part in the barrier protocol so the MMA warp's schedule is unchanged.
            const bool warp_live = tile_q0 + wq * 32 < S;

            for (int j = 0; j < n_blocks; ++j) {
                const bool tail = j >= n_full;
                mbar_wait(&s_full[t], j & 1);
                tc_fence_after();
                if (warp_live) {
                    // pass 1: row maximum (independent chains)
                    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
                    if (!tail) {
#pragma unroll
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
                        for (int c = 0; c < tail_rows; c += 32) {
                            uint32_t v[32];
                            tmem_ld_32x32(tm_s + c, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                mx0 = fmaxf(mx0, (c + i < tail_valid) ? __uint_as_float(v[i]) : -INFINITY);
                        }
                    }
                    const float mxs = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
                    const bool need = mxs > m_ref + RESCALE_THRESHOLD;  // always true for block 0 (m_ref = -inf)
                    if (__any_sync(0xffffffffu, need)) {                 // warp-uniform: TMEM ops are warp-collective
                        const float new_ref = need ? mxs : m_ref;
                        if (j > 0) {
                            // s_full(j) implies P_{j-1} V_{j-1} has completed, so O is quiescent: rescale it in place
                            const float alpha = ex2(m_ref - new_ref);  // 1 for rows that keep their reference
#pragma unroll 1
                            for (int c = 0; c < HD; c += 32) {
                                uint32_t v[32];
                                tmem_ld_32x32(tm_o + c, v);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                                tmem_st_32x32(tm_o + c, v);
                            }
                            tmem_st_wait();
                            l_run *= alpha;
                        }
                        m_ref = new_ref;
                    }
                    // pass 2: probabilities -> bf16 pairs, row sum (independent chains); P then overwrites S in TMEM
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    if (!tail) {
                        // packed fp32x2 arithmetic (FFMA2 / FADD2): one issue slot per two keys for the scale and the sum
                        const f32x2 sc2 = pack2(scale_log2, scale_log2), nm2 = pack2(-m_ref, -m_ref);
                        f32x2 sum01 = pack2(0.f, 0.f), sum23 = pack2(0.f, 0.f);
#pragma unroll
                        for (int c = 0; c < BKV; c += 32) {
                            uint32_t v[32];
                            tmem_ld_32x32(tm_s + c, v);
                            tmem_ld_wait();
                            uint32_t pk[16];  // 32 probabilities, two per register (even key in the low half)
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                float x0, x1, x2, x3;
                                unpack2(fma2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2), x0, x1);
                                unpack2(fma2(pack2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), sc2, nm2), x2, x3);
                                const float p0 = ex2(x0), p1 = ex2(x1), p2 = ex2(x2), p3 = ex2(x3);
                                sum01 = add2(sum01, pack2(p0, p1));
                                sum23 = add2(sum23, pack2(p2, p3));
                                pk[i >> 1] = pack_bf16x2(p0, p1);
                                pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
                            }
                            // P columns [c/2, c/2+16) overwrite S columns that this thread has already consumed
                            tmem_st_32x16(tm_s + (c >> 1), pk);
                        }
                        unpack2(sum01, s0, s1);
                        unpack2(sum23, s2, s3);
                    } else {
                        uint32_t v[32];  // tail_rows == 32 columns (attention_tail_rows caps the ragged block)
                        tmem_ld_32x32(tm_s, v);
                        tmem_ld_wait();
                        uint32_t pt[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float e0 = ex2(fmaf(__uint_as_float(v[i]), scale_log2, -m_ref));
                            const float e1 = ex2(fmaf(__uint_as_float(v[i + 1]), scale_log2, -m_ref));
                            const float p0 = (i < tail_valid) ? e0 : 0.f;
                            const float p1 = (i + 1 < tail_valid) ? e1 : 0.f;
                            s0 += p0 + p1;
                            pt[i >> 1] = pack_bf16x2(p0, p1);
                        }
                        tmem_st_32x16(tm_s, pt);
                    }
                    tmem_st_wait();
                    l_run += (s0 + s1) + (s2 + s3);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[t]);  // one arrival per softmax warp
            }

            mbar_wait(&o_done[t], 0);
            tc_fence_after();
            const int q = tile_q0 + r;
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
    }

    tc_fence_before();
    __syncthreads();
    if (warp == SOFTMAX_WARPS + 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

template <int HD>
cudaError_t launch_attention_hd(const CUtensorMap& tm_q, const CUtensorMap& tm_kv, const CUtensorMap& tm_tail,
                                __nv_bfloat16* out, int frames, int S, int heads, float softmax_scale,
                                cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Geo<HD>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((S + BQ * TILES - 1) / (BQ * TILES), heads, frames);
    const float scale_log2 = softmax_scale * 1.4426950408889634f;
    attention_fwd_kernel<HD><<<grid, THREADS, Geo<HD>::SMEM_BYTES, stream>>>(tm_q, tm_kv, tm_tail, out, S, heads * HD,
                                                                             attention_tail_rows(S), scale_log2);
    return cudaGetLastError();
}

}  // namespace

int attention_q_rows() { return BQ; }
int attention_kv_rows() { return BKV; }
int attention_tail_rows(int S) {
    (void)S;
    return 32;  // the ragged block is processed 32 keys at a time (see launch_attention: S % 64 must be <= 32)
}

cudaError_t launch_attention(const CUtensorMap& tm_q, const CUtensorMap& tm_kv, const CUtensorMap& tm_tail,
                             __nv_bfloat16* out, int frames, int S, int heads, int head_dim, float softmax_scale,
                             cudaStream_t stream) {
    if (frames <= 0 || S <= 0 || heads <= 0) return cudaErrorInvalidValue;
    if (S % BKV > 32) return cudaErrorInvalidValue;  // ragged tail of at most 32 keys (1025 % 64 == 1)
    if (head_dim == 64) return launch_attention_hd<64>(tm_q, tm_kv, tm_tail, out, frames, S, heads, softmax_scale, stream);
    if (head_dim == 128) return launch_attention_hd<128>(tm_q, tm_kv, tm_tail, out, frames, S, heads, softmax_scale, stream);
    return cudaErrorInvalidValue;
}

}  // namespace aigv
