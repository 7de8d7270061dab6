// EXPERIMENT, NOT BUILT INTO THE LIBRARY (kept as the record of a measured dead end; DESIGN.md section 5, attention).
// Result on B200 (same-box A/B, profiles/r1f_ab_experiments.txt): numerically correct (all attention and end-to-end parity
// tests pass), but 18.7 ms/step against 15.4 ms for attention_tcgen05.cu. With S double-buffered the softmax warps no
// longer wait for the tensor pipe, yet each warp's per-block chain (tcgen05.ld -> max -> pair barrier -> exp -> tcgen05.st
// -> wait::st -> arrive, ~2700 cycles for 256 cycles of MUFU work) is serial inside the warp and 4 warps per SMSP cannot
// cover it; halving the work per thread halves the useful share of that chain. To build it: add this file to
// _lib.SOURCES, declare launch_attention_pp / attention_pp_covers in kernels.h and dispatch from launch_attention.
//
// Ping-pong variant of the fused attention for head_dim 64 and ViT sequence lengths (S = 256k or 256k + 1).
//
// Same contract as attention_tcgen05.cu (FlashAttention.forward, flash_attention.py:52-55 / _naive_attn,
// modeling_intern_vit.py:143-157) and the same tile geometry (one CTA = one (frame, head) x two 128-query tiles sharing
// every K/V stage, P written over S in TMEM, O in TMEM, stale-maximum rescale, class-token row on the producer warp,
// lone tail key folded into the epilogues). What changes is the dependency chain. There, S_j is single-buffered with P_j
// aliased onto it, so per tile  softmax(j) -> P.V(j) -> Q.K^T(j+1) -> softmax(j+1)  is strictly serial and the softmax
// warps spend ~36% of their time waiting for the tensor pipe's round trip (ncu, profiles/r1f_ncu_attn_summary.txt).
// Here every tile owns TWO S buffers (192 TMEM columns per tile: S0, S1, O; one CTA per SM), the MMA warp keeps
// Q.K^T two blocks ahead ( ... P.V(j), Q.K^T(j+2) ... ), and softmax(j+1) starts the moment softmax(j) ends.
// To keep 16 softmax warps per SM with one CTA, each query row is shared by TWO threads (column halves of the block):
// one tcgen05.ld of 32 columns per thread and block -- S is read from TMEM once, both passes run from registers -- and the
// two partners (same TMEM lane quadrant, warps w and w+4) exchange their row maximum through shared memory under a
// 64-thread named barrier. That barrier also orders the partner's P store (bf16 columns [16,32) alias S columns [16,32)
// of the first half) after the first half's load.
#include "kernels.h"
#include "ptx.cuh"

namespace aigv {

namespace {

constexpr int BQ = 128;     // query rows per tile
constexpr int TILES = 2;    // query tiles per CTA
constexpr int BKV = 64;     // keys per block
constexpr int HD = 64;      // head dim
constexpr int SM_WARPS = 16;  // softmax warps: tile = w >> 3, column half = (w >> 2) & 1, TMEM lane quadrant = w & 3
constexpr int THREADS = (SM_WARPS + 2) * 32;
constexpr int STAGES = 6;   // K / V ring depth (K runs two blocks ahead of V)
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

constexpr int Q_BYTES = BQ * 128;    // one 64-column (128-byte) swizzle atom per tile
constexpr int KV_BYTES = BKV * 128;
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILES * Q_BYTES;
constexpr int OFF_V = OFF_K + STAGES * KV_BYTES;
constexpr int OFF_BAR = OFF_V + STAGES * KV_BYTES;
constexpr int OFF_XCHG = OFF_BAR + 512;                       // float [2 bufs][TILES][2 halves][BQ] row maxima, then
constexpr int OFF_LSUM = OFF_XCHG + 2 * TILES * 2 * BQ * 4;   // float [TILES][2 halves][BQ] row-sum halves
constexpr int OFF_SIDE = OFF_LSUM + TILES * 2 * BQ * 4;       // side region (fp32): q_side[HD], p_side[BKV], k_t[HD], v_t[HD]
constexpr int SMEM_BYTES = OFF_SIDE + (3 * HD + BKV) * 4;
constexpr uint32_t TM_S = 64;        // columns per S buffer
constexpr uint32_t TM_O = 128;       // O accumulator at +128
constexpr uint32_t TM_TILE = 192;
constexpr uint32_t TMEM_COLS = 512;  // 2 x 192 used

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float bf_lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf_hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }

__global__ void __launch_bounds__(THREADS, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int S, int D,
                    float scale_log2) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = q_full + 1;           // [STAGES]
    uint64_t* k_empty = k_full + STAGES;     // [STAGES]
    uint64_t* v_full = k_empty + STAGES;     // [STAGES]
    uint64_t* v_empty = v_full + STAGES;     // [STAGES]
    uint64_t* s_full = v_empty + STAGES;     // [TILES][2] S_j in buffer j&1
    // [TILES][2] P_j written into buffer j&1, O rescaled if needed (1 arrival per softmax warp of the tile). One barrier
    // per buffer: the softmax may finish blocks j and j+1 before the MMA warp has looked at block j, and a single
    // barrier two phases ahead of its waiter aliases (parity) and deadlocks.
    uint64_t* p_full = s_full + 2 * TILES;
    uint64_t* pv_done = p_full + 2 * TILES;  // [TILES] P.V(j) complete (only the rare rescale waits on it)
    uint64_t* o_done = pv_done + TILES;      // [TILES] last P.V accumulated
    uint64_t* kt_full = o_done + TILES;      // the lone tail key's k / v rows are in shared memory (fp32)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kt_full + 1);
    float* xchg = reinterpret_cast<float*>(smem + OFF_XCHG);
    float* lsum = reinterpret_cast<float*>(smem + OFF_LSUM);
    float* q_side = reinterpret_cast<float*>(smem + OFF_SIDE);
    float* p_side = q_side + HD;
    float* kt_side = p_side + BKV;
    float* vt_side = kt_side + HD;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (BQ * TILES);
    const int head = blockIdx.y;
    const int frame = blockIdx.z;
    const int row_base = frame * S;
    const int n_blocks = S / BKV;                 // full blocks only (launcher: S % 64 is 0 or 1, S >= 256)
    const bool tail_key = (S % BKV) == 1;         // folded in by the epilogues
    const bool side = tail_key && blockIdx.x == 0;  // S % 256 == 1: the leftover query row rides on this CTA's producer warp

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("aigv attention_pp: dynamic smem base not 1024B aligned\n");
        __trap();
    }
    if (warp == SM_WARPS && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
    }
    if (warp == SM_WARPS + 1) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&k_full[s], 1);
                mbar_init(&k_empty[s], 1);
                mbar_init(&v_full[s], 1);
                mbar_init(&v_empty[s], 1);
            }
            for (int t = 0; t < TILES; ++t) {
                mbar_init(&s_full[2 * t], 1);
                mbar_init(&s_full[2 * t + 1], 1);
                mbar_init(&p_full[2 * t], SM_WARPS / TILES);
                mbar_init(&p_full[2 * t + 1], SM_WARPS / TILES);
                mbar_init(&pv_done[t], 1);
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

    if (warp == SM_WARPS) {
        // ------------------------------------------------------------ TMA producer (+ the side row)
        constexpr int DPL = HD / 32;
        float m_run = -INFINITY, l_lane = 0.f, o_acc[DPL];
#pragma unroll
        for (int i = 0; i < DPL; ++i) o_acc[i] = 0.f;
        if (tail_key) {
            const __nv_bfloat16* kg = qkv + static_cast<size_t>(row_base + S - 1) * (3 * D) + D + head * HD;
#pragma unroll
            for (int i = 0; i < DPL; ++i) {
                kt_side[lane * DPL + i] = __bfloat162float(kg[lane * DPL + i]);
                vt_side[lane * DPL + i] = __bfloat162float(kg[D + lane * DPL + i]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(kt_full);
        }
        if (side) {
            const __nv_bfloat16* qg = qkv + static_cast<size_t>(row_base + S - 1) * (3 * D) + head * HD;
#pragma unroll
            for (int i = 0; i < DPL; ++i) q_side[lane * DPL + i] = __bfloat162float(qg[lane * DPL + i]) * scale_log2;
            __syncwarp();
        }
        auto side_block = [&](int jj) {
            const int st = jj % STAGES;
            const uint32_t ph = (jj / STAGES) & 1;
            mbar_wait(&k_full[st], ph);
            const uint8_t* kb = smem + OFF_K + st * KV_BYTES;
            float sc[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int key = lane + 32 * hh;
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 kk = *reinterpret_cast<const uint4*>(kb + key * 128 + ((c ^ (key & 7)) << 4));
                    const float4 qa = *reinterpret_cast<const float4*>(q_side + c * 8);
                    const float4 qb = *reinterpret_cast<const float4*>(q_side + c * 8 + 4);
                    a0 = fmaf(bf_lo(kk.x), qa.x, a0); a1 = fmaf(bf_hi(kk.x), qa.y, a1);
                    a0 = fmaf(bf_lo(kk.y), qa.z, a0); a1 = fmaf(bf_hi(kk.y), qa.w, a1);
                    a0 = fmaf(bf_lo(kk.z), qb.x, a0); a1 = fmaf(bf_hi(kk.z), qb.y, a1);
                    a0 = fmaf(bf_lo(kk.w), qb.z, a0); a1 = fmaf(bf_hi(kk.w), qb.w, a1);
                }
                sc[hh] = a0 + a1;
            }
            float mb = fmaxf(sc[0], sc[1]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
            const float m_new = fmaxf(m_run, mb);
            const float corr = ex2(m_run - m_new);
            const float p0 = ex2(sc[0] - m_new), p1 = ex2(sc[1] - m_new);
            l_lane = fmaf(l_lane, corr, p0 + p1);
            m_run = m_new;
            p_side[lane] = p0;
            p_side[lane + 32] = p1;
            __syncwarp();
            mbar_wait(&v_full[st], ph);
            const uint8_t* vb = smem + OFF_V + st * KV_BYTES;
            const int boff = lane * DPL * 2;
#pragma unroll
            for (int i = 0; i < DPL; ++i) o_acc[i] *= corr;
#pragma unroll 4
            for (int key = 0; key < BKV; ++key) {
                const float pk = p_side[key];
                const uint32_t vv = *reinterpret_cast<const uint32_t*>(
                    vb + key * 128 + ((((boff >> 4) ^ (key & 7)) << 4) | (boff & 15)));
                o_acc[0] = fmaf(pk, bf_lo(vv), o_acc[0]);
                o_acc[1] = fmaf(pk, bf_hi(vv), o_acc[1]);
            }
            __syncwarp();
        };

        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, TILES * Q_BYTES);
            for (int t = 0; t < TILES; ++t)
                tma_load_2d(smem + OFF_Q + t * Q_BYTES, &tm_q, q_full, head * HD, row_base + q0 + t * BQ);
        }
        __syncwarp();
        for (int j = 0; j < n_blocks; ++j) {
            const int st = j % STAGES;
            const uint32_t ph = (j / STAGES) & 1;
            const int r = row_base + j * BKV;
            mbar_wait(&k_empty[st], ph ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&k_full[st], KV_BYTES);
                tma_load_2d(smem + OFF_K + st * KV_BYTES, &tm_kv, &k_full[st], D + head * HD, r);
            }
            __syncwarp();
            mbar_wait(&v_empty[st], ph ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&v_full[st], KV_BYTES);
                tma_load_2d(smem + OFF_V + st * KV_BYTES, &tm_kv, &v_full[st], 2 * D + head * HD, r);
            }
            __syncwarp();
            if (side && j > 0) side_block(j - 1);  // trails the loads by one block; its stage is refilled STAGES-1 later
        }
        if (side) {
            side_block(n_blocks - 1);
            float l = l_lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
            {   // the lone tail key
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
            *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(o_acc[0] * inv, o_acc[1] * inv);
        }
    } else if (warp == SM_WARPS + 1) {
        // ------------------------------------------------------------ MMA issuer
        // prologue: S(t)[0] = Q(t) K_0^T, S(t)[1] = Q(t) K_1^T;  then per block j and tile t:
        //   wait P_j(t) -> O(t) += P_j V_j (A = P from TMEM buffer j&1) -> S(t)[j&1] = Q(t) K_{j+2}^T
        // The tensor pipe executes one thread's MMAs in issue order, so the Q.K^T that overwrites buffer j&1 runs after
        // the P.V that reads it.
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t q_desc0 = umma_desc_sw128(smem_u32(smem + OFF_Q), 16, 1024);
        const uint64_t k_desc0 = umma_desc_sw128(smem_u32(smem + OFF_K), 16, 1024);
        const uint64_t v_desc0 = umma_desc_sw128(smem_u32(smem + OFF_V), KV_BYTES, 1024);
        constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, 0, 0);
        constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, 0, 1);  // B (= V) is MN-major, N = HD
        auto qk_mmas = [&](int t, int jb) {  // S(t)[jb & 1] = Q(t) K_jb^T ; inside an elected region
            const int st = jb % STAGES;
            const uint64_t qd = umma_desc_advance(q_desc0, t * Q_BYTES);
            const uint64_t kd = umma_desc_advance(k_desc0, st * KV_BYTES);
            const uint32_t tm_s = tmem_u + t * TM_TILE + (jb & 1) * TM_S;
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
                umma_bf16_ss(tm_s, umma_desc_advance(qd, k * 32), umma_desc_advance(kd, k * 32), idesc_qk, k != 0 ? 1u : 0u);
            umma_commit(&s_full[2 * t + (jb & 1)]);
            if (t == TILES - 1) umma_commit(&k_empty[st]);
        };
        mbar_wait(q_full, 0);
        for (int jb = 0; jb < 2 && jb < n_blocks; ++jb) {
            mbar_wait(&k_full[jb % STAGES], (jb / STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
                for (int t = 0; t < TILES; ++t) qk_mmas(t, jb);
            }
            __syncwarp();
        }
        for (int j = 0; j < n_blocks; ++j) {
            const int st = j % STAGES;
            const bool more = j + 2 < n_blocks;
            const uint64_t vd = umma_desc_advance(v_desc0, st * KV_BYTES);
            for (int t = 0; t < TILES; ++t) {
                mbar_wait(&p_full[2 * t + (j & 1)], (j >> 1) & 1);
                if (t == 0) {
                    mbar_wait(&v_full[st], (j / STAGES) & 1);
                    if (more) mbar_wait(&k_full[(j + 2) % STAGES], ((j + 2) / STAGES) & 1);
                }
                tc_fence_after();
                const uint32_t tm_p = tmem_u + t * TM_TILE + (j & 1) * TM_S;  // P_j: bf16 pairs, first 32 columns of S[j&1]
                const uint32_t tm_o = tmem_u + t * TM_TILE + TM_O;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)
                        umma_bf16_ts(tm_o, tm_p + k * 8, umma_desc_advance(vd, k * 2048), idesc_pv, (j | k) != 0 ? 1u : 0u);
                    umma_commit(&pv_done[t]);
                    if (t == TILES - 1) umma_commit(&v_empty[st]);
                    if (more) qk_mmas(t, j + 2);
                    if (j == n_blocks - 1) umma_commit(&o_done[t]);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------ softmax: two threads per query row
        const int t = warp >> 3;            // query tile
        const int h = (warp >> 2) & 1;      // column half of every S block (keys [32h, 32h+32)) and of O
        const int wq = warp & 3;            // TMEM lane quadrant
        const int r = wq * 32 + lane;       // row within the tile == TMEM lane
        const uint32_t tm_tile = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + t * TM_TILE;
        const uint32_t tm_o = tm_tile + TM_O + 32 * h;
        const uint32_t pair_bar = 1 + t * 4 + wq;  // named barrier of this warp and its partner (64 threads)
        float* mx_mine = xchg + (t * 2 + h) * BQ + r;        // + buf * (TILES * 2 * BQ)
        const float* mx_other = xchg + (t * 2 + (h ^ 1)) * BQ + r;
        float m_ref = -INFINITY;  // reference maximum (scaled), identical in both partners
        float l_run = 0.f;        // this thread's half of the row sum

        for (int j = 0; j < n_blocks; ++j) {
            const int buf = j & 1;
            mbar_wait(&s_full[2 * t + buf], (j >> 1) & 1);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld_32x32(tm_tile + buf * TM_S + 32 * h, v);
            tmem_ld_wait();
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(v[i]));
                mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(v[i + 2]));
                mx3 = fmaxf(mx3, __uint_as_float(v[i + 3]));
            }
            float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            mx_mine[buf * (TILES * 2 * BQ)] = mx;
            // Partners meet: exchanges the half-row maxima AND orders the other half's P store (bf16 columns [16,32) alias
            // S columns [16,32), which belong to half 0) after half 0's load above.
            tc_fence_before();
            named_bar_sync(pair_bar, 64);
            tc_fence_after();
            mx = fmaxf(mx, mx_other[buf * (TILES * 2 * BQ)]);
            const float mxs = mx * scale_log2;
            const bool need = mxs > m_ref + RESCALE_THRESHOLD;  // always true for block 0; same decision in both partners
            if (__any_sync(0xffffffffu, need)) {
                const float new_ref = need ? mxs : m_ref;
                if (j > 0) {
                    // O must be quiescent: P.V(j-1) may still be running (s_full(j) only implies P.V(j-2))
                    mbar_wait(&pv_done[t], (j - 1) & 1);
                    tc_fence_after();
                    const float alpha = ex2(m_ref - new_ref);
                    uint32_t o[32];
                    tmem_ld_32x32(tm_o, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st_32x32(tm_o, o);
                    tmem_st_wait();
                    l_run *= alpha;
                }
                m_ref = new_ref;
            }
            // probabilities from registers -> bf16 pairs over the S buffer, row-sum half
            const f32x2 sc2 = pack2(scale_log2, scale_log2), nm2 = pack2(-m_ref, -m_ref);
            f32x2 sum01 = pack2(0.f, 0.f), sum23 = pack2(0.f, 0.f);
            uint32_t pk[16];
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
            tmem_st_32x16(tm_tile + buf * TM_S + 16 * h, pk);
            tmem_st_wait();
            float s0, s1, s2, s3;
            unpack2(sum01, s0, s1);
            unpack2(sum23, s2, s3);
            l_run += (s0 + s1) + (s2 + s3);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[2 * t + buf]);
        }

        // epilogue: the lone tail key (both partners compute the same s_t), the row sum across the two halves, O / l
        float alpha = 1.f, p_t = 0.f;
        if (tail_key) {
            mbar_wait(kt_full, 0);
            const uint8_t* qrow = smem + OFF_Q + t * Q_BYTES + r * 128;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 qq = *reinterpret_cast<const uint4*>(qrow + ((c ^ (r & 7)) << 4));
                const float4 ka = *reinterpret_cast<const float4*>(kt_side + c * 8);
                const float4 kb = *reinterpret_cast<const float4*>(kt_side + c * 8 + 4);
                a0 = fmaf(bf_lo(qq.x), ka.x, a0); a1 = fmaf(bf_hi(qq.x), ka.y, a1);
                a0 = fmaf(bf_lo(qq.y), ka.z, a0); a1 = fmaf(bf_hi(qq.y), ka.w, a1);
                a0 = fmaf(bf_lo(qq.z), kb.x, a0); a1 = fmaf(bf_hi(qq.z), kb.y, a1);
                a0 = fmaf(bf_lo(qq.w), kb.z, a0); a1 = fmaf(bf_hi(qq.w), kb.w, a1);
            }
            const float s_t = (a0 + a1) * scale_log2;
            const float new_ref = fmaxf(m_ref, s_t);
            alpha = ex2(m_ref - new_ref);
            p_t = ex2(s_t - new_ref);
        }
        lsum[(t * 2 + h) * BQ + r] = l_run;
        named_bar_sync(pair_bar, 64);
        const float l = fmaf(l_run + lsum[(t * 2 + (h ^ 1)) * BQ + r], alpha, p_t);
        mbar_wait(&o_done[t], 0);
        tc_fence_after();
        const float inv = 1.0f / l;
        const float ai = alpha * inv, pi = p_t * inv;
        __nv_bfloat16* o = out + static_cast<size_t>(row_base + q0 + t * BQ + r) * D + head * HD + 32 * h;
        uint32_t v[32];
        tmem_ld_32x32(tm_o, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(v[i + e]) * ai;
            if (tail_key) {
                const float4 va = *reinterpret_cast<const float4*>(vt_side + 32 * h + i);
                const float4 vb = *reinterpret_cast<const float4*>(vt_side + 32 * h + i + 4);
                y[0] = fmaf(pi, va.x, y[0]); y[1] = fmaf(pi, va.y, y[1]); y[2] = fmaf(pi, va.z, y[2]); y[3] = fmaf(pi, va.w, y[3]);
                y[4] = fmaf(pi, vb.x, y[4]); y[5] = fmaf(pi, vb.y, y[5]); y[6] = fmaf(pi, vb.z, y[6]); y[7] = fmaf(pi, vb.w, y[7]);
            }
            uint4 w;
            w.x = pack_bf16x2(y[0], y[1]);
            w.y = pack_bf16x2(y[2], y[3]);
            w.z = pack_bf16x2(y[4], y[5]);
            w.w = pack_bf16x2(y[6], y[7]);
            *reinterpret_cast<uint4*>(o + i) = w;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == SM_WARPS + 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

}  // namespace

bool attention_pp_covers(int S, int head_dim) {
    return head_dim == HD && S >= BQ * TILES && (S % (BQ * TILES) == 0 || S % (BQ * TILES) == 1);
}

cudaError_t launch_attention_pp(const CUtensorMap& tm_q, const CUtensorMap& tm_kv, const __nv_bfloat16* qkv,
                                __nv_bfloat16* out, int frames, int S, int heads, float softmax_scale,
                                cudaStream_t stream) {
    if (!attention_pp_covers(S, HD) || !qkv) return cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attention_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid(S / (BQ * TILES), heads, frames);
    const float scale_log2 = softmax_scale * 1.4426950408889634f;
    attention_pp_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tm_q, tm_kv, qkv, out, S, heads * HD, scale_log2);
    return cudaGetLastError();
}

}  // namespace aigv
