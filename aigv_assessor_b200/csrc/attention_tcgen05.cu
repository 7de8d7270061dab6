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
// S = 1025 = 16*64 + 1. On the KEY side a ragged tail of 2..32 keys is one short KV block (32 columns, invalid ones masked
// to -inf); a tail of exactly ONE key (every ViT sequence: the class token makes S = 64k + 1) never enters the tensor
// pipeline: the producer warp leaves that key's k and v rows in shared memory as fp32 and every query row folds it in
// during its epilogue (one 64-term dot product against its own Q row, an exact max/alpha update, O += p v), which saves
// one of the 17 block round trips. On the query
// side a ViT sequence is (multiple of 256) + 1 (the class token): the one leftover query row would cost a whole extra
// CTA per (frame, head) -- 20% of the grid for 0.1% of the rows. Instead ("side row" mode) the TMA producer warp of the
// first CTA of each (frame, head) computes that row on the CUDA cores from the very K/V stages it has just loaded
// (fp32 online softmax, one warp, ~700 instructions per 64-key block, hidden in the shadow of the tensor pipeline), and
// the grid has exactly S/256 CTAs per (frame, head). Other sequence lengths keep the ceil(S/256) grid, where the last
// CTA skips its empty second tile.
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
    static constexpr int OFF_SIDE = OFF_BAR + 256;          // 1 + 4*STAGES + 3*TILES mbarriers + the TMEM base slot
    // side region (fp32): q of the side row (pre-scaled) [HD], one block of its p [BKV], the lone tail key's k [HD], v [HD]
    static constexpr int SMEM_BYTES = OFF_SIDE + (3 * HD + BKV) * 4;
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
__device__ __forceinline__ float bf_lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf_hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }

template <int HD>
__global__ void __launch_bounds__(THREADS, Geo<HD>::MIN_CTAS)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     const __grid_constant__ CUtensorMap tm_tail, const __nv_bfloat16* __restrict__ qkv,
                     __nv_bfloat16* __restrict__ out, int S, int D, int tail_rows, float scale_log2, int side_row) {
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
    uint64_t* kt_full = o_done + TILES;      // the lone tail key's k / v rows are in shared memory (fp32)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kt_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (BQ * TILES);
    const int head = blockIdx.y;
    const int frame = blockIdx.z;
    const int row_base = frame * S;          // first token row of this frame in the [M, 3D] matrix
    const int n_full = S / BKV;              // full KV blocks
    const int tail_valid = S - n_full * BKV; // valid keys in the tail block (0 = no tail)
    const bool tail_epi = tail_valid == 1 && n_full >= 1;  // lone tail key: folded in by the epilogues, no KV block
    const int n_blocks = n_full + ((tail_valid > 0 && !tail_epi) ? 1 : 0);
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
            mbar_init(kt_full, 1);
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
        // ------------------------------------------------------------ TMA producer (+ the side row, see header)
        {
            constexpr int DPL = HD / 32;  // output dims per lane of the side row
            const bool side = side_row != 0 && blockIdx.x == 0;
            float* q_side = reinterpret_cast<float*>(smem + G::OFF_SIDE);  // [HD] q * scale * log2e
            float* p_side = q_side + HD;                                   // [BKV] probabilities of the current block
            float* kt_side = p_side + BKV;                                 // [HD] k of the lone tail key
            float* vt_side = kt_side + HD;                                 // [HD] v of the lone tail key
            if (tail_epi) {
                const __nv_bfloat16* kg = qkv + static_cast<size_t>(row_base + S - 1) * (3 * D) + D + head * HD;
#pragma unroll
                for (int i = 0; i < DPL; ++i) {
                    kt_side[lane * DPL + i] = __bfloat162float(kg[lane * DPL + i]);
                    vt_side[lane * DPL + i] = __bfloat162float(kg[D + lane * DPL + i]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(kt_full);
            }
            float m_run = -INFINITY, l_lane = 0.f, o_acc[DPL];
#pragma unroll
            for (int i = 0; i < DPL; ++i) o_acc[i] = 0.f;
            if (side) {
                const __nv_bfloat16* qg = qkv + static_cast<size_t>(row_base + S - 1) * (3 * D) + head * HD;
#pragma unroll
                for (int i = 0; i < DPL; ++i) q_side[lane * DPL + i] = __bfloat162float(qg[lane * DPL + i]) * scale_log2;
                __syncwarp();
            }
            // One 64-key block of the side row: scores on lanes (keys lane, lane+32), online softmax across the warp,
            // then every lane accumulates its DPL output dims over the block's keys.
            auto side_block = [&](int jj) {
                const int st = jj % STAGES;
                const uint32_t ph = (jj / STAGES) & 1;
                const bool tail = jj >= n_full;
                const int rows = tail ? tail_rows : BKV;
                const int atom_stride = tail ? tail_rows * 128 : KV_ATOM;
                mbar_wait(&k_full[st], ph);
                const uint8_t* kb = smem + OFF_K + st * KV_BYTES;
                float sc[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int key = lane + 32 * h;
                    float a0 = 0.f, a1 = 0.f;
                    if (key < rows) {
#pragma unroll
                        for (int a = 0; a < ATOMS; ++a)
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const uint4 kk = *reinterpret_cast<const uint4*>(kb + a * atom_stride + key * 128 + ((c ^ (key & 7)) << 4));
                                const float4 qa = *reinterpret_cast<const float4*>(q_side + a * 64 + c * 8);
                                const float4 qb = *reinterpret_cast<const float4*>(q_side + a * 64 + c * 8 + 4);
                                a0 = fmaf(bf_lo(kk.x), qa.x, a0); a1 = fmaf(bf_hi(kk.x), qa.y, a1);
                                a0 = fmaf(bf_lo(kk.y), qa.z, a0); a1 = fmaf(bf_hi(kk.y), qa.w, a1);
                                a0 = fmaf(bf_lo(kk.z), qb.x, a0); a1 = fmaf(bf_hi(kk.z), qb.y, a1);
                                a0 = fmaf(bf_lo(kk.w), qb.z, a0); a1 = fmaf(bf_hi(kk.w), qb.w, a1);
                            }
                    }
                    sc[h] = (key < rows && jj * BKV + key < S) ? a0 + a1 : -INFINITY;
                }
                float mb = fmaxf(sc[0], sc[1]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
                const float m_new = fmaxf(m_run, mb);  // finite: every block holds at least one valid key
                const float corr = ex2(m_run - m_new);
                const float p0 = ex2(sc[0] - m_new), p1 = ex2(sc[1] - m_new);
                l_lane = fmaf(l_lane, corr, p0 + p1);
                m_run = m_new;
                p_side[lane] = p0;
                p_side[lane + 32] = p1;
                __syncwarp();
                mbar_wait(&v_full[st], ph);
                const uint8_t* vb = smem + OFF_V + st * KV_BYTES + ((lane * DPL) >> 6) * atom_stride;
                const int boff = ((lane * DPL) & 63) * 2;  // byte offset of this lane's dims inside the 128-byte row
#pragma unroll
                for (int i = 0; i < DPL; ++i) o_acc[i] *= corr;
#pragma unroll 4
                for (int key = 0; key < rows; ++key) {
                    const float pk = p_side[key];
                    const uint8_t* vr = vb + key * 128 + ((((boff >> 4) ^ (key & 7)) << 4) | (boff & 15));
#pragma unroll
                    for (int i = 0; i < DPL; i += 2) {
                        const uint32_t vv = *reinterpret_cast<const uint32_t*>(vr + 2 * i);
                        o_acc[i] = fmaf(pk, bf_lo(vv), o_acc[i]);
                        o_acc[i + 1] = fmaf(pk, bf_hi(vv), o_acc[i + 1]);
                    }
                }
                __syncwarp();  // p_side is rewritten by the next block
            };

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
                // the side row trails the loads by one block; its stage is only refilled STAGES-1 iterations from now
                if (side && j > 0) side_block(j - 1);
            }
            if (side) {
                side_block(n_blocks - 1);
                float l = l_lane;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
                if (tail_epi) {  // the lone tail key: one more online-softmax step, dims spread over the lanes
                    float st = 0.f;
#pragma unroll
                    for (int i = 0; i < DPL; ++i) st = fmaf(q_side[lane * DPL + i], kt_side[lane * DPL + i], st);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) st += __shfl_xor_sync(0xffffffffu, st, o);
                    const float m_new = fmaxf(m_run, st);
                    const float corr = ex2(m_run - m_new), pt = ex2(st - m_new);
                    l = fmaf(l, corr, pt);
#pragma unroll
                    for (int i = 0; i < DPL; ++i) o_acc[i] = fmaf(o_acc[i], corr, pt * vt_side[lane * DPL + i]);
                }
                const float inv = 1.0f / l;
                __nv_bfloat16* o = out + static_cast<size_t>(row_base + S - 1) * D + head * HD + lane * DPL;
#pragma unroll
                for (int i = 0; i < DPL; i += 2)
                    *reinterpret_cast<uint32_t*>(o + i) = pack_bf16x2(o_acc[i] * inv, o_acc[i + 1] * inv);
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

            // The lone tail key (see header): s_t = q_row . k_t from this row's Q tile line in shared memory, then an exact
            // reference update. Computed before the o_done wait: it only needs Q and the side region.
            float alpha = 1.f, p_t = 0.f;
            const float* vt_side = reinterpret_cast<const float*>(smem + G::OFF_SIDE) + 2 * HD + BKV;
            if (tail_epi && warp_live) {
                const float* kt_side = vt_side - HD;
                mbar_wait(kt_full, 0);
                const uint8_t* qrow = smem + OFF_Q + t * Q_BYTES + r * 128;
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int a = 0; a < ATOMS; ++a)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 qq = *reinterpret_cast<const uint4*>(qrow + a * Q_ATOM + ((c ^ (r & 7)) << 4));
                        const float4 ka = *reinterpret_cast<const float4*>(kt_side + a * 64 + c * 8);
                        const float4 kb = *reinterpret_cast<const float4*>(kt_side + a * 64 + c * 8 + 4);
                        a0 = fmaf(bf_lo(qq.x), ka.x, a0); a1 = fmaf(bf_hi(qq.x), ka.y, a1);
                        a0 = fmaf(bf_lo(qq.y), ka.z, a0); a1 = fmaf(bf_hi(qq.y), ka.w, a1);
                        a0 = fmaf(bf_lo(qq.z), kb.x, a0); a1 = fmaf(bf_hi(qq.z), kb.y, a1);
                        a0 = fmaf(bf_lo(qq.w), kb.z, a0); a1 = fmaf(bf_hi(qq.w), kb.w, a1);
                    }
                const float s_t = (a0 + a1) * scale_log2;
                const float new_ref = fmaxf(m_ref, s_t);
                alpha = ex2(m_ref - new_ref);
                p_t = ex2(s_t - new_ref);
                l_run = fmaf(l_run, alpha, p_t);
            }
            mbar_wait(&o_done[t], 0);
            tc_fence_after();
            const int q = tile_q0 + r;
            if (warp_live) {
                const float inv = 1.0f / l_run;
                const float ai = alpha * inv, pi = p_t * inv;
                __nv_bfloat16* o = out + static_cast<size_t>(row_base + q) * D + head * HD;
#pragma unroll
                for (int c = 0; c < HD; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(tm_o + c, v);
                    tmem_ld_wait();
                    if (q < S) {
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            float y[8];
                            if (tail_epi) {
                                const float4 va = *reinterpret_cast<const float4*>(vt_side + c + i);
                                const float4 vb = *reinterpret_cast<const float4*>(vt_side + c + i + 4);
                                y[0] = fmaf(__uint_as_float(v[i + 0]), ai, pi * va.x);
                                y[1] = fmaf(__uint_as_float(v[i + 1]), ai, pi * va.y);
                                y[2] = fmaf(__uint_as_float(v[i + 2]), ai, pi * va.z);
                                y[3] = fmaf(__uint_as_float(v[i + 3]), ai, pi * va.w);
                                y[4] = fmaf(__uint_as_float(v[i + 4]), ai, pi * vb.x);
                                y[5] = fmaf(__uint_as_float(v[i + 5]), ai, pi * vb.y);
                                y[6] = fmaf(__uint_as_float(v[i + 6]), ai, pi * vb.z);
                                y[7] = fmaf(__uint_as_float(v[i + 7]), ai, pi * vb.w);
                            } else {
#pragma unroll
                                for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(v[i + e]) * inv;
                            }
                            uint4 w;
                            w.x = pack_bf16x2(y[0], y[1]);
                            w.y = pack_bf16x2(y[2], y[3]);
                            w.z = pack_bf16x2(y[4], y[5]);
                            w.w = pack_bf16x2(y[6], y[7]);
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
                                const __nv_bfloat16* qkv, __nv_bfloat16* out, int frames, int S, int heads,
                                float softmax_scale, cudaStream_t stream) {
    static unsigned long long optin_mask = 0;  // per device (kernels.h)
    cudaError_t e = ensure_smem_optin(reinterpret_cast<const void*>(&attention_fwd_kernel<HD>), Geo<HD>::SMEM_BYTES,
                                      optin_mask);
    if (e != cudaSuccess) return e;
    // side-row mode: S = k * 256 + 1 (k >= 1) -- the leftover query row rides on the producer warp of CTA 0
    const int side_row = (S > BQ * TILES && S % (BQ * TILES) == 1 && qkv != nullptr) ? 1 : 0;
    dim3 grid(side_row ? S / (BQ * TILES) : (S + BQ * TILES - 1) / (BQ * TILES), heads, frames);
    const float scale_log2 = softmax_scale * 1.4426950408889634f;
    attention_fwd_kernel<HD><<<grid, THREADS, Geo<HD>::SMEM_BYTES, stream>>>(
        tm_q, tm_kv, tm_tail, qkv, out, S, heads * HD, attention_tail_rows(S), scale_log2, side_row);
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
                             const __nv_bfloat16* qkv, __nv_bfloat16* out, int frames, int S, int heads, int head_dim,
                             float softmax_scale, cudaStream_t stream) {
    if (frames <= 0 || S <= 0 || heads <= 0) return cudaErrorInvalidValue;
    if (S % BKV > 32) return cudaErrorInvalidValue;  // ragged tail of at most 32 keys (1025 % 64 == 1)
    if (head_dim == 64) return launch_attention_hd<64>(tm_q, tm_kv, tm_tail, qkv, out, frames, S, heads, softmax_scale, stream);
    if (head_dim == 128) return launch_attention_hd<128>(tm_q, tm_kv, tm_tail, qkv, out, frames, S, heads, softmax_scale, stream);
    return cudaErrorInvalidValue;
}

}  // namespace aigv
