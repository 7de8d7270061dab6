// Persistent fused bidirectional flash-style attention on tcgen05/TMEM for sm_100a (head_dim 64 and 128), round-2 form.
//
// Replaces FlashAttention.forward -> flash_attn_varlen_qkvpacked_func (flash_attention.py:52-55, non-causal,
// dropout 0, softmax_scale = d^-0.5) reached from InternAttention._flash_attn (modeling_intern_vit.py:162-177), and is
// numerically the softmax(q*scale @ k^T) @ v of _naive_attn (modeling_intern_vit.py:143-157).
//
// What changed against the round-1 kernel (attention_tcgen05.cu) and why (profiles/r1f_ncu_attn_summary.txt: XU 54%,
// tensor 27%, 39% of the softmax warps' samples waiting for S): the per-tile chain softmax(j) -> P V(j) -> Q K^T(j+1)
// was serial because P aliased the only S buffer of a tile, every exponential went through MUFU, and a 64-key block
// paid one commit -> mbarrier -> tcgen05.ld round trip per 64 exps per thread.
//   * ONE persistent CTA per SM walks a list of work items (frame, head, 256 query rows = two 128-row tiles).
//     Barriers, TMEM and descriptors are set up once; the loads / Q K^T of item i+1 overlap the epilogue of item i.
//   * 128-key blocks: one query row per thread holds the 128 scores of a block in registers (one tcgen05.ld pass, half
//     as many round trips); the softmax warpgroups get 208 registers through setmaxnreg, the four service warps 88.
//   * head_dim 64: THREE S buffers rotate over the two tiles (512 TMEM columns = 3 x 128 S/P + 2 x 64 O). The MMA warp
//     keeps Q K^T three (tile, block) steps ahead of P V, so when a warpgroup has written P(j) its S(j+1) is already
//     in TMEM: the tensor-pipe latency leaves the per-tile chain. head_dim 128 has room for two S buffers
//     (2 x 128 + 2 x 128 columns) and keeps the alternating schedule.
//   * AIGV_ATTN_POLY of every 4 pairs of exponentials are evaluated on the FMA pipe (Cody-Waite split + cubic minimax
//     polynomial of 2^f on [-0.5, 0.5], |rel err| <= 7.5e-5, 26x below the bf16 rounding of P) with packed FFMA2.
//   warp 8: TMA producer   warp 9: MMA issuer   warps 10/11: the class-token query row (see below)
//   warps 0-3 / 4-7: softmax warpgroup of query tile 0 / 1, one query row per thread (TMEM lane == row).
// O lives in TMEM for the whole KV loop of an item. m_ref is a *stale* row maximum: it is only raised (and O, l
// rescaled through tcgen05.ld/st, after waiting for the tile's previous P V) when the true running maximum exceeds it
// by more than 2^8, so p <= 256 always and the rescale is rare; the final O / l is exact regardless of the reference.
// Ragged ends. Keys: S % 128 == 1 (every ViT sequence: 1024 patches + the class token): the lone tail key never
// enters the tensor pipeline -- its k / v rows sit in shared memory as fp32 and every query row folds it in during its
// epilogue (a head_dim-term dot product with its own Q row, an exact max/alpha update, O*alpha + p*v). Any other tail
// is one more (masked) block. Queries: S % 256 == 1: the class-token row rides on warps 10/11 of the CTA that owns the
// first 256 rows of that (frame, head) -- fp32 online softmax on the CUDA cores from the very K/V stages the tensor
// pipe consumes, 64 keys of every block per warp -- so the grid has exactly S/256 items per (frame, head).
// A masked tail block multiplies P == 0 with V rows of the NEXT frame (TMA zero-fills past the end of the matrix);
// with finite activations that contributes exactly nothing.
#include "kernels.h"
#include "ptx.cuh"

#ifndef AIGV_ATTN_POLY
#define AIGV_ATTN_POLY 0   // pairs out of every 4 pairs of exponentials evaluated on the FMA pipe (0..4)
#endif
#ifndef AIGV_ATTN_ABL
#define AIGV_ATTN_ABL 0   // timing ablations (WRONG results): 1 = no MUFU, 2 = no tcgen05.ld of S, 4 = no tcgen05.st of P, 8 = no row max
#endif
#ifndef AIGV_ATTN_PINGPONG
#define AIGV_ATTN_PINGPONG 1   // the two softmax warps of an SM sub-partition take turns in their exponential phase
#endif

namespace aigv {

namespace {

constexpr int BQ = 128;     // query rows per tile
constexpr int TILES = 2;    // query tiles per work item
constexpr int CHUNK = BQ * TILES;
constexpr int BKV = 128;    // keys per block
constexpr int SOFTMAX_WARPS = 4 * TILES;
constexpr int WARP_PRODUCER = SOFTMAX_WARPS, WARP_MMA = SOFTMAX_WARPS + 1, WARP_SIDE = SOFTMAX_WARPS + 2;
constexpr int THREADS = (SOFTMAX_WARPS + 4) * 32;   // three warpgroups (setmaxnreg works per warpgroup)
constexpr int REGS_SOFTMAX = 208, REGS_SERVICE = 88;  // 2*128*208 + 128*88 = 64512 <= 384 * 168
constexpr float RESCALE_THRESHOLD = 8.0f;           // log2 units

template <int HD>
struct Geo2 {
    static constexpr int ATOMS = HD / 64;                 // 64-column (128-byte) swizzle atoms per operand row
    static constexpr int Q_ATOM = BQ * 128;               // 16 KiB
    static constexpr int KV_ATOM = BKV * 128;             // 16 KiB
    static constexpr int Q_BYTES = ATOMS * Q_ATOM;
    static constexpr int KV_BYTES = ATOMS * KV_ATOM;
    static constexpr int NSB = HD == 64 ? 3 : 2;          // S/P buffers in TMEM
    static constexpr int STAGES = HD == 64 ? 4 : 2;       // K / V ring depth
    static constexpr int QBUFS = HD == 64 ? 2 : 1;        // Q tiles of the next item land while this one computes
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = OFF_Q + QBUFS * TILES * Q_BYTES;
    static constexpr int OFF_V = OFF_K + STAGES * KV_BYTES;
    static constexpr int OFF_BAR = OFF_V + STAGES * KV_BYTES;
    static constexpr int OFF_SIDE = OFF_BAR + 512;
    // fp32 side region: [2][2*HD] lone tail key k, v (ring of two items) | [2][HD] q of the class-token row (one copy
    // per side warp) | [2][64] its probabilities | [HD + 32] merge buffer of the second side warp
    static constexpr int SIDE_FLOATS = 2 * 2 * HD + 2 * HD + 2 * 64 + HD + 32;
    static constexpr int SMEM_BYTES = OFF_SIDE + SIDE_FLOATS * 4;
    static constexpr uint32_t TM_O = NSB * BKV;           // O accumulators behind the S buffers
    static constexpr uint32_t TMEM_COLS = 512;
    static_assert(NSB * BKV + TILES * HD <= 512, "TMEM budget");
};

struct Item {
    int frame, head, q0, row_base, n_t;
    bool side;
};

// item index -> (frame, head, chunk). Items of one (frame, head) are adjacent (their CTAs share K/V in L2); the chunk is
// rotated by the pass number so the CTA that gets the class-token row changes from pass to pass.
__device__ __forceinline__ Item decode_item(int i, int n_chunks, int rot_div, int H, int S, bool side_mode) {
    Item m;
    const int fh = i / n_chunks;
    const int raw = i - fh * n_chunks;
    const int chunk = (raw + fh / rot_div) % n_chunks;
    m.frame = fh / H;
    m.head = fh - m.frame * H;
    m.q0 = chunk * CHUNK;
    m.row_base = m.frame * S;
    m.n_t = (m.q0 + BQ < S) ? TILES : 1;
    m.side = side_mode && chunk == 0;
    return m;
}

template <uint32_t N>
__device__ __forceinline__ void reg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N));
}
template <uint32_t N>
__device__ __forceinline__ void reg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N));
}

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float bf_lo2(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf_hi2(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }

// 2^x for a pair on the FMA / ALU pipes (no MUFU): x = n + f, n = rint(x) taken from the low mantissa bits of
// x + 1.5*2^23, f in [-0.5, 0.5]; 2^f by a cubic (minimax in relative error, 7.5e-5); the exponent is added as an integer.
// x is clamped at -126 so the exponent field cannot wrap (the result is then ~1e-38 instead of 0).
__device__ __forceinline__ void exp2_poly2(f32x2 x, float& p0, float& p1) {
    float x0, x1;
    unpack2(x, x0, x1);
    x0 = fmaxf(x0, -126.0f);
    x1 = fmaxf(x1, -126.0f);
    const f32x2 xc = pack2(x0, x1);
    const f32x2 t = add2(xc, pack2(12582912.0f, 12582912.0f));
    const f32x2 n = add2(t, pack2(-12582912.0f, -12582912.0f));
    const f32x2 f = fma2(n, pack2(-1.0f, -1.0f), xc);
    f32x2 p = fma2(pack2(0.0551715381f, 0.0551715381f), f, pack2(0.242610976f, 0.242610976f));
    p = fma2(p, f, pack2(0.693261027f, 0.693261027f));
    p = fma2(p, f, pack2(0.999928117f, 0.999928117f));
    float t0, t1, q0, q1;
    unpack2(t, t0, t1);
    unpack2(p, q0, q1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

template <int HD>
__global__ void __launch_bounds__(THREADS, 1)
attention_fwd2_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __nv_bfloat16* __restrict__ qkv,
                      __nv_bfloat16* __restrict__ out, int S, int H, int n_items, int n_chunks, int rot_div,
                      int side_mode, float scale_log2) {
    using G = Geo2<HD>;
    constexpr int ATOMS = G::ATOMS, Q_ATOM = G::Q_ATOM, KV_ATOM = G::KV_ATOM, Q_BYTES = G::Q_BYTES,
                  KV_BYTES = G::KV_BYTES, NSB = G::NSB, STAGES = G::STAGES, QBUFS = G::QBUFS, OFF_Q = G::OFF_Q,
                  OFF_K = G::OFF_K, OFF_V = G::OFF_V;
    constexpr uint32_t TM_O = G::TM_O;
    constexpr int DPL = HD / 32;  // dims per lane wherever one warp spreads a head_dim vector over its lanes
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::OFF_BAR);
    uint64_t* q_full = bars;                 // [QBUFS] Q tiles landed (TMA bytes) + the lone tail key's rows written
    uint64_t* q_empty = q_full + QBUFS;      // [QBUFS] last Q K^T of the item done + every softmax warp done with Q
    uint64_t* k_full = q_empty + QBUFS;      // [STAGES]
    uint64_t* k_empty = k_full + STAGES;     // [STAGES] MMA commit + the two side warps
    uint64_t* v_full = k_empty + STAGES;     // [STAGES]
    uint64_t* v_empty = v_full + STAGES;     // [STAGES]
    uint64_t* s_full = v_empty + STAGES;     // [NSB] S of one (tile, block) in TMEM
    uint64_t* p_full = s_full + NSB;         // [NSB] S consumed, P written over it (one arrival per softmax warp)
    uint64_t* pv_done = p_full + NSB;        // [TILES] a P V of this tile has completed (O quiescent until the next P)
    uint64_t* o_full = pv_done + TILES;      // [TILES] last P V of the item accumulated
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + TILES);
    float* side_f = reinterpret_cast<float*>(smem + G::OFF_SIDE);
    float* kt_ring = side_f;                  // [2][2*HD]: k then v of the lone tail key, item parity
    float* q_side = kt_ring + 2 * 2 * HD;     // [2][HD]
    float* p_side = q_side + 2 * HD;          // [2][64]
    float* merge = p_side + 2 * 64;           // [HD + 32]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_full = S / BKV;
    const int tail_valid = S - n_full * BKV;
    const bool tail_epi = tail_valid == 1 && n_full >= 1;
    const int n_blocks = n_full + ((tail_valid > 0 && !tail_epi) ? 1 : 0);
    const bool smode = side_mode != 0;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("aigv attention: dynamic smem base not 1024B aligned\n");
        __trap();
    }
    if (warp == WARP_PRODUCER && lane == 0) tma_prefetch_desc(&tm_qkv);
    if (warp == WARP_MMA) {
        if (lane == 0) {
            for (int b = 0; b < QBUFS; ++b) {
                mbar_init(&q_full[b], 2);
                mbar_init(&q_empty[b], 1 + SOFTMAX_WARPS);
            }
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&k_full[s], 1);
                mbar_init(&k_empty[s], 3);
                mbar_init(&v_full[s], 1);
                mbar_init(&v_empty[s], 3);
            }
            for (int b = 0; b < NSB; ++b) {
                mbar_init(&s_full[b], 1);
                mbar_init(&p_full[b], 4);
            }
            for (int t = 0; t < TILES; ++t) {
                mbar_init(&pv_done[t], 1);
                mbar_init(&o_full[t], 1);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<G::TMEM_COLS>(tmem_slot);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= SOFTMAX_WARPS) {
        reg_dec<REGS_SERVICE>();
        if (warp == WARP_PRODUCER) {
            // ------------------------------------------------------------ TMA producer
            int it = 0;
            for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++it) {
                const Item m = decode_item(i, n_chunks, rot_div, H, S, smode);
                const int qb = it % QBUFS;
                auto load_kv = [&](int j) {
                    const int kv = it * n_blocks + j;
                    const int st = kv % STAGES;
                    const uint32_t ph = (kv / STAGES) & 1;
                    const int r = m.row_base + j * BKV;
                    mbar_wait(&k_empty[st], ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&k_full[st], KV_BYTES);
                        for (int a = 0; a < ATOMS; ++a)
                            tma_load_2d(smem + OFF_K + st * KV_BYTES + a * KV_ATOM, &tm_qkv, &k_full[st],
                                        H * HD + m.head * HD + a * 64, r);
                    }
                    __syncwarp();
                    mbar_wait(&v_empty[st], ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&v_full[st], KV_BYTES);
                        for (int a = 0; a < ATOMS; ++a)
                            tma_load_2d(smem + OFF_V + st * KV_BYTES + a * KV_ATOM, &tm_qkv, &v_full[st],
                                        2 * H * HD + m.head * HD + a * 64, r);
                    }
                    __syncwarp();
                };
                // With a single Q buffer the Q tiles can only land after the previous item's last Q K^T: have the
                // first K/V block in flight before waiting for that.
                int j0 = 0;
                if (QBUFS == 1 && STAGES >= 2) { load_kv(0); j0 = 1; }
                mbar_wait(&q_empty[qb], ((it / QBUFS) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&q_full[qb], m.n_t * Q_BYTES);
                    for (int t = 0; t < m.n_t; ++t)
                        for (int a = 0; a < ATOMS; ++a)
                            tma_load_2d(smem + OFF_Q + (qb * TILES + t) * Q_BYTES + a * Q_ATOM, &tm_qkv, &q_full[qb],
                                        m.head * HD + a * 64, m.row_base + m.q0 + t * BQ);
                }
                __syncwarp();
                if (tail_epi) {
                    float* kt = kt_ring + (it & 1) * 2 * HD;
                    const __nv_bfloat16* kg =
                        qkv + static_cast<size_t>(m.row_base + S - 1) * (3 * H * HD) + H * HD + m.head * HD;
#pragma unroll
                    for (int e = 0; e < DPL; ++e) {
                        kt[lane * DPL + e] = __bfloat162float(kg[lane * DPL + e]);
                        kt[HD + lane * DPL + e] = __bfloat162float(kg[H * HD + lane * DPL + e]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&q_full[qb]);
                for (int j = j0; j < n_blocks; ++j) load_kv(j);
            }
        } else if (warp == WARP_MMA) {
            // ------------------------------------------------------------ MMA issuer
            // Two cursors walk the same (item, block, tile) sequence: Q K^T runs NSB steps ahead of P V. Step g uses
            // S buffer g % NSB; the P V of step g reads P from the buffer the Q K^T of step g + NSB overwrites, and the
            // tensor pipe executes this thread's MMAs in issue order.
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint64_t q_desc0 = umma_desc_sw128(smem_u32(smem + OFF_Q), 16, 1024);
            const uint64_t k_desc0 = umma_desc_sw128(smem_u32(smem + OFF_K), 16, 1024);
            const uint64_t v_desc0 = umma_desc_sw128(smem_u32(smem + OFF_V), KV_ATOM, 1024);
            constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, HD, 0, 1);  // B (= V) is MN-major, N = HD
            struct Cur { int it, j, t, n_t; };
            auto item_nt = [&](int it) -> int {
                const long long i = static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x;
                if (i >= n_items) return 0;
                return decode_item(static_cast<int>(i), n_chunks, rot_div, H, S, smode).n_t;
            };
            auto advance = [&](Cur& c) {
                if (++c.t == c.n_t) {
                    c.t = 0;
                    if (++c.j == n_blocks) {
                        c.j = 0;
                        ++c.it;
                        c.n_t = item_nt(c.it);
                    }
                }
            };
            Cur cq{0, 0, 0, item_nt(0)};
            Cur cp = cq;
            uint32_t gq = 0, gp = 0;
            auto issue_qk = [&]() {
                const int qb = cq.it % QBUFS;
                const int kv = cq.it * n_blocks + cq.j;
                const int st = kv % STAGES;
                if (cq.j == 0 && cq.t == 0) mbar_wait(&q_full[qb], (cq.it / QBUFS) & 1);
                if (cq.t == 0) mbar_wait(&k_full[st], (kv / STAGES) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t qd = umma_desc_advance(q_desc0, (qb * TILES + cq.t) * Q_BYTES);
                    const uint64_t kd = umma_desc_advance(k_desc0, st * KV_BYTES);
                    const uint32_t b = gq % NSB;
                    const uint32_t tm_s = tmem_u + b * BKV;
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)  // d axis: 4 steps of 32 B inside each 64-column atom
                        umma_bf16_ss(tm_s, umma_desc_advance(qd, (k >> 2) * Q_ATOM + (k & 3) * 32),
                                     umma_desc_advance(kd, (k >> 2) * KV_ATOM + (k & 3) * 32), idesc_qk, k != 0 ? 1u : 0u);
                    umma_commit(&s_full[b]);
                    if (cq.t == cq.n_t - 1) {
                        umma_commit(&k_empty[st]);
                        if (cq.j == n_blocks - 1) umma_commit(&q_empty[qb]);
                    }
                }
                __syncwarp();
                ++gq;
                advance(cq);
            };
            auto issue_pv = [&]() {
                const uint32_t b = gp % NSB;
                const int kv = cp.it * n_blocks + cp.j;
                const int st = kv % STAGES;
                mbar_wait(&p_full[b], (gp / NSB) & 1);  // S consumed, P in TMEM, O rescaled if it had to be
                if (cp.t == 0) mbar_wait(&v_full[st], (kv / STAGES) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t vd = umma_desc_advance(v_desc0, st * KV_BYTES);
                    const uint32_t tm_p = tmem_u + b * BKV;  // bf16 pairs in the first 64 columns of the S buffer
                    const uint32_t tm_o = tmem_u + TM_O + cp.t * HD;
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)
                        // A = P from TMEM: 8 columns (16 keys) per step; B = V rows [16k, 16k+16): 2 KiB apart
                        umma_bf16_ts(tm_o, tm_p + k * 8, umma_desc_advance(vd, k * 2048), idesc_pv,
                                     (cp.j | k) != 0 ? 1u : 0u);
                    umma_commit(&pv_done[cp.t]);
                    if (cp.t == cp.n_t - 1) umma_commit(&v_empty[st]);
                    if (cp.j == n_blocks - 1) umma_commit(&o_full[cp.t]);
                }
                __syncwarp();
                ++gp;
                advance(cp);
            };
            while (gq < static_cast<uint32_t>(NSB) && cq.n_t) issue_qk();
            while (cp.n_t) {
                issue_pv();
                if (cq.n_t) issue_qk();
            }
        } else {
            // ------------------------------------------------------------ the class-token query row (side warps)
            // Both warps consume every K/V stage (the ring's empty barriers expect them); on items that own the
            // class-token row each warp runs an fp32 online softmax over its 64 keys of every block.
            const int w = warp - WARP_SIDE;
            float* qs = q_side + w * HD;
            float* ps = p_side + w * 64;
            int it = 0;
            for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++it) {
                const Item m = decode_item(i, n_chunks, rot_div, H, S, smode);
                float m_run = -INFINITY, l_lane = 0.f, o_acc[DPL];
#pragma unroll
                for (int e = 0; e < DPL; ++e) o_acc[e] = 0.f;
                const __nv_bfloat16* qg = qkv + static_cast<size_t>(m.row_base + S - 1) * (3 * H * HD) + m.head * HD;
                if (m.side) {
#pragma unroll
                    for (int e = 0; e < DPL; ++e) qs[lane * DPL + e] = __bfloat162float(qg[lane * DPL + e]) * scale_log2;
                    __syncwarp();
                }
                for (int j = 0; j < n_blocks; ++j) {
                    const int kv = it * n_blocks + j;
                    const int st = kv % STAGES;
                    const uint32_t ph = (kv / STAGES) & 1;
                    mbar_wait(&k_full[st], ph);
                    float corr = 1.f;
                    if (m.side) {  // side mode: every block is full (S = 256 k + 1, the tail key is folded at the end)
                        const uint8_t* kb = smem + OFF_K + st * KV_BYTES;
                        float sc[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int key = w * 64 + lane + 32 * h;
                            float a0 = 0.f, a1 = 0.f;
#pragma unroll
                            for (int a = 0; a < ATOMS; ++a)
#pragma unroll
                                for (int c = 0; c < 8; ++c) {
                                    const uint4 kk = *reinterpret_cast<const uint4*>(kb + a * KV_ATOM + key * 128 + ((c ^ (key & 7)) << 4));
                                    const float4 qa = *reinterpret_cast<const float4*>(qs + a * 64 + c * 8);
                                    const float4 qc = *reinterpret_cast<const float4*>(qs + a * 64 + c * 8 + 4);
                                    a0 = fmaf(bf_lo2(kk.x), qa.x, a0); a1 = fmaf(bf_hi2(kk.x), qa.y, a1);
                                    a0 = fmaf(bf_lo2(kk.y), qa.z, a0); a1 = fmaf(bf_hi2(kk.y), qa.w, a1);
                                    a0 = fmaf(bf_lo2(kk.z), qc.x, a0); a1 = fmaf(bf_hi2(kk.z), qc.y, a1);
                                    a0 = fmaf(bf_lo2(kk.w), qc.z, a0); a1 = fmaf(bf_hi2(kk.w), qc.w, a1);
                                }
                            sc[h] = a0 + a1;
                        }
                        float mb = fmaxf(sc[0], sc[1]);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
                        const float m_new = fmaxf(m_run, mb);
                        corr = ex2f(m_run - m_new);
                        const float p0 = ex2f(sc[0] - m_new), p1 = ex2f(sc[1] - m_new);
                        l_lane = fmaf(l_lane, corr, p0 + p1);
                        m_run = m_new;
                        ps[lane] = p0;
                        ps[lane + 32] = p1;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&k_empty[st]);
                    mbar_wait(&v_full[st], ph);
                    if (m.side) {
                        const uint8_t* vb = smem + OFF_V + st * KV_BYTES + ((lane * DPL) >> 6) * KV_ATOM;
                        const int boff = ((lane * DPL) & 63) * 2;  // byte offset of this lane's dims inside the 128-byte row
#pragma unroll
                        for (int e = 0; e < DPL; ++e) o_acc[e] *= corr;
#pragma unroll 4
                        for (int kk = 0; kk < 64; ++kk) {
                            const int key = w * 64 + kk;
                            const float pk = ps[kk];
                            const uint8_t* vr = vb + key * 128 + ((((boff >> 4) ^ (key & 7)) << 4) | (boff & 15));
#pragma unroll
                            for (int e = 0; e < DPL; e += 2) {
                                const uint32_t vv = *reinterpret_cast<const uint32_t*>(vr + 2 * e);
                                o_acc[e] = fmaf(pk, bf_lo2(vv), o_acc[e]);
                                o_acc[e + 1] = fmaf(pk, bf_hi2(vv), o_acc[e + 1]);
                            }
                        }
                    }
                    __syncwarp();  // ps is rewritten by the next block; all lanes are done with the V stage
                    if (lane == 0) mbar_arrive(&v_empty[st]);
                }
                if (m.side) {
                    float l = l_lane;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
                    if (w == 1) {
#pragma unroll
                        for (int e = 0; e < DPL; ++e) merge[lane * DPL + e] = o_acc[e];
                        if (lane == 0) { merge[HD] = m_run; merge[HD + 1] = l; }
                    }
                    named_bar_sync(1, 64);
                    if (w == 0) {
                        const float m1 = merge[HD], l1 = merge[HD + 1];
                        float mm = fmaxf(m_run, m1);
                        float a0 = ex2f(m_run - mm), a1 = ex2f(m1 - mm);
                        l = l * a0 + l1 * a1;
#pragma unroll
                        for (int e = 0; e < DPL; ++e) o_acc[e] = o_acc[e] * a0 + merge[lane * DPL + e] * a1;
                        if (tail_epi) {  // the lone tail key: one more online-softmax step, dims spread over the lanes
                            const __nv_bfloat16* kg = qg + H * HD;
                            float stl = 0.f, vt[DPL];
#pragma unroll
                            for (int e = 0; e < DPL; ++e) {
                                stl = fmaf(qs[lane * DPL + e], __bfloat162float(kg[lane * DPL + e]), stl);
                                vt[e] = __bfloat162float(kg[H * HD + lane * DPL + e]);
                            }
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) stl += __shfl_xor_sync(0xffffffffu, stl, o);
                            const float m_new = fmaxf(mm, stl);
                            const float corr = ex2f(mm - m_new), pt = ex2f(stl - m_new);
                            l = fmaf(l, corr, pt);
#pragma unroll
                            for (int e = 0; e < DPL; ++e) o_acc[e] = fmaf(o_acc[e], corr, pt * vt[e]);
                        }
                        const float inv = 1.0f / l;
                        __nv_bfloat16* o = out + static_cast<size_t>(m.row_base + S - 1) * (H * HD) + m.head * HD + lane * DPL;
#pragma unroll
                        for (int e = 0; e < DPL; e += 2)
                            *reinterpret_cast<uint32_t*>(o + e) = pack_bf16x2(o_acc[e] * inv, o_acc[e + 1] * inv);
                    }
                    named_bar_sync(1, 64);  // the merge buffer may be rewritten
                }
            }
        }
    } else {
        // ------------------------------------------------------------ softmax warpgroups: one query row per thread
        reg_inc<REGS_SOFTMAX>();
        const int t = warp >> 2;                      // query tile of this warpgroup
        const int wq = warp & 3;                      // TMEM lane quadrant
        const int r = wq * 32 + lane;                 // row within the tile == TMEM lane
        // Ping-pong (measured: without it the two softmax warps that share an SM sub-partition drift into lock step,
        // both in their MUFU-heavy phase at once and both out of it at once: XU 66% busy). Warps w (tile 0) and w + 4
        // (tile 1) sit on the same scheduler; named barrier 2 + 2*wq + t is "tile t's turn" of that pair: a warp waits
        // for its turn before its exponentials and hands the turn over right after them, so one warp's tcgen05.ld /
        // row max / tcgen05.st / barrier traffic runs under the other's exponentials. Tile 0 goes first.
        const uint32_t pp_mine = 2 + 2 * wq + t, pp_other = 2 + 2 * wq + (t ^ 1);
        if (AIGV_ATTN_PINGPONG && t == 1) named_bar_arrive(pp_other, 64);
        const uint32_t tm_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
        const uint32_t tm_o = tm_lane + TM_O + t * HD;
        uint32_t g_base = 0;     // (tile, block) steps of all earlier items
        uint32_t n_pv = 0;       // P V steps of this tile so far (phase of pv_done)
        uint32_t n_items_t = 0;  // items this tile took part in (phase of o_full)
        int it = 0;
        for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++it) {
            const Item m = decode_item(i, n_chunks, rot_div, H, S, smode);
            const int qb = it % QBUFS;
            const uint32_t qph = (it / QBUFS) & 1;
            if (t >= m.n_t) {  // this warpgroup's tile holds no query row: only release Q (and keep passing the turn)
                mbar_wait(&q_full[qb], qph);
                if (lane == 0) mbar_arrive(&q_empty[qb]);
                g_base += n_blocks * m.n_t;
                if (AIGV_ATTN_PINGPONG)
                    for (int j = 0; j < n_blocks; ++j) {
                        named_bar_sync(pp_mine, 64);
                        named_bar_arrive(pp_other, 64);
                    }
                continue;
            }
            const int tile_q0 = m.q0 + t * BQ;
            // Warps whose 32 query rows are all beyond the sequence skip the softmax math; they still take part in
            // the barrier protocol so the MMA warp's schedule is unchanged.
            const bool warp_live = tile_q0 + wq * 32 < S;
            float m_ref = -INFINITY;  // reference maximum, already multiplied by scale*log2e
            float l_run = 0.f;
            float s_t = 0.f;          // the lone tail key's score (log2 units)

            for (int j = 0; j < n_blocks; ++j) {
                const uint32_t g = g_base + j * m.n_t + t;
                const uint32_t b = g % NSB;
                const uint32_t tm_s = tm_lane + b * BKV;
                mbar_wait(&s_full[b], (g / NSB) & 1);
                tc_fence_after();
                if (warp_live) {
                    uint32_t v[BKV];
                    if (AIGV_ATTN_ABL & 2) {
#pragma unroll
                        for (int c = 0; c < BKV; ++c) v[c] = __float_as_uint(-0.01f * (c + j) - l_run * 1e-30f);
                    } else {
#pragma unroll
                        for (int c = 0; c < BKV; c += 32) tmem_ld_32x32(tm_s + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
                        tmem_ld_wait();
                    }
                    if (j >= n_full) {  // masked tail block (never taken for ViT sequences)
#pragma unroll
                        for (int c = 0; c < BKV; ++c)
                            if (c >= tail_valid) v[c] = 0xff800000u;
                    }
                    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
                    if (AIGV_ATTN_ABL & 8) mx0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[77]));
#pragma unroll
                    for (int c = 0; c < ((AIGV_ATTN_ABL & 8) ? 0 : BKV); c += 8) {
                        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v[c]), __uint_as_float(v[c + 1])));
                        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])));
                        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(v[c + 4]), __uint_as_float(v[c + 5])));
                        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(v[c + 6]), __uint_as_float(v[c + 7])));
                    }
                    const float mxs = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
                    const bool need = mxs > m_ref + RESCALE_THRESHOLD;  // always true for block 0 (m_ref = -inf)
                    if (__any_sync(0xffffffffu, need)) {                 // warp-uniform: TMEM ops are warp-collective
                        const float new_ref = need ? mxs : m_ref;
                        if (j > 0) {
                            // O must be quiescent: wait for this tile's previous P V (with two S buffers s_full already
                            // implies it), then rescale O in place
                            if (NSB > 2) {
                                mbar_wait(&pv_done[t], (n_pv - 1) & 1);
                                tc_fence_after();
                            }
                            const float alpha = ex2f(m_ref - new_ref);  // 1 for rows that keep their reference
#pragma unroll 1
                            for (int c = 0; c < HD; c += 32) {
                                uint32_t o[32];
                                tmem_ld_32x32(tm_o + c, o);
                                tmem_ld_wait();
#pragma unroll
                                for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                                tmem_st_32x32(tm_o + c, o);
                            }
                            l_run *= alpha;
                        }
                        m_ref = new_ref;
                    }
                    if (AIGV_ATTN_PINGPONG) named_bar_sync(pp_mine, 64);
                    // probabilities -> bf16 pairs written over the S buffer, row sum; packed fp32x2 arithmetic
                    const f32x2 sc2 = pack2(scale_log2, scale_log2), nm2 = pack2(-m_ref, -m_ref);
                    f32x2 sum01 = pack2(0.f, 0.f), sum23 = pack2(0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < BKV; c += 32) {
                        uint32_t pk[16];  // 32 probabilities, two per register (even key in the low half)
#pragma unroll
                        for (int e = 0; e < 32; e += 8) {
                            float p[8];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const f32x2 x = fma2(pack2(__uint_as_float(v[c + e + 2 * q]), __uint_as_float(v[c + e + 2 * q + 1])), sc2, nm2);
                                if (q < AIGV_ATTN_POLY) {
                                    exp2_poly2(x, p[2 * q], p[2 * q + 1]);
                                } else {
                                    float x0, x1;
                                    unpack2(x, x0, x1);
                                    if (AIGV_ATTN_ABL & 1) {
                                        p[2 * q] = x0;
                                        p[2 * q + 1] = x1;
                                    } else {
                                        p[2 * q] = ex2f(x0);
                                        p[2 * q + 1] = ex2f(x1);
                                    }
                                }
                            }
                            sum01 = add2(sum01, add2(pack2(p[0], p[1]), pack2(p[4], p[5])));
                            sum23 = add2(sum23, add2(pack2(p[2], p[3]), pack2(p[6], p[7])));
#pragma unroll
                            for (int q = 0; q < 4; ++q) pk[(e >> 1) + q] = pack_bf16x2(p[2 * q], p[2 * q + 1]);
                        }
                        if (AIGV_ATTN_ABL & 4) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) l_run += __uint_as_float(pk[e]) * 1e-30f;
                        } else {
                            tmem_st_32x16(tm_s + (c >> 1), pk);
                        }
                    }
                    if (AIGV_ATTN_PINGPONG) named_bar_arrive(pp_other, 64);
                    tmem_st_wait();
                    float s0, s1, s2, s3;
                    unpack2(sum01, s0, s1);
                    unpack2(sum23, s2, s3);
                    l_run += (s0 + s1) + (s2 + s3);
                } else if (AIGV_ATTN_PINGPONG) {
                    named_bar_sync(pp_mine, 64);
                    named_bar_arrive(pp_other, 64);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[b]);  // one arrival per softmax warp
                ++n_pv;
                if (j == 0) {
                    // The lone tail key's score from this row's Q line in shared memory (the producer's fp32 copy of
                    // the key is covered by q_full), then this warp is done with the Q buffer.
                    if (tail_epi && warp_live) {
                        mbar_wait(&q_full[qb], qph);
                        const float* kt = kt_ring + (it & 1) * 2 * HD;
                        const uint8_t* qrow = smem + OFF_Q + (qb * TILES + t) * Q_BYTES + r * 128;
                        float a0 = 0.f, a1 = 0.f;
#pragma unroll
                        for (int a = 0; a < ATOMS; ++a)
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const uint4 qq = *reinterpret_cast<const uint4*>(qrow + a * Q_ATOM + ((c ^ (r & 7)) << 4));
                                const float4 ka = *reinterpret_cast<const float4*>(kt + a * 64 + c * 8);
                                const float4 kc = *reinterpret_cast<const float4*>(kt + a * 64 + c * 8 + 4);
                                a0 = fmaf(bf_lo2(qq.x), ka.x, a0); a1 = fmaf(bf_hi2(qq.x), ka.y, a1);
                                a0 = fmaf(bf_lo2(qq.y), ka.z, a0); a1 = fmaf(bf_hi2(qq.y), ka.w, a1);
                                a0 = fmaf(bf_lo2(qq.z), kc.x, a0); a1 = fmaf(bf_hi2(qq.z), kc.y, a1);
                                a0 = fmaf(bf_lo2(qq.w), kc.z, a0); a1 = fmaf(bf_hi2(qq.w), kc.w, a1);
                            }
                        s_t = (a0 + a1) * scale_log2;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&q_empty[qb]);
                }
            }
            g_base += n_blocks * m.n_t;

            // epilogue: fold the lone tail key in (exact reference update), normalise, store
            float alpha = 1.f, p_t = 0.f;
            const float* vt = kt_ring + (it & 1) * 2 * HD + HD;
            if (tail_epi && warp_live) {
                const float new_ref = fmaxf(m_ref, s_t);
                alpha = ex2f(m_ref - new_ref);
                p_t = ex2f(s_t - new_ref);
                l_run = fmaf(l_run, alpha, p_t);
            }
            mbar_wait(&o_full[t], n_items_t & 1);
            ++n_items_t;
            tc_fence_after();
            const int q = tile_q0 + r;
            if (warp_live) {
                const float inv = 1.0f / l_run;
                const float ai = alpha * inv, pi = p_t * inv;
                __nv_bfloat16* o = out + static_cast<size_t>(m.row_base + q) * (H * HD) + m.head * HD;
#pragma unroll
                for (int c = 0; c < HD; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(tm_o + c, v);
                    tmem_ld_wait();
                    if (q < S) {
#pragma unroll
                        for (int e = 0; e < 32; e += 8) {
                            float y[8];
                            if (tail_epi) {
                                const float4 va = *reinterpret_cast<const float4*>(vt + c + e);
                                const float4 vb = *reinterpret_cast<const float4*>(vt + c + e + 4);
                                y[0] = fmaf(__uint_as_float(v[e + 0]), ai, pi * va.x);
                                y[1] = fmaf(__uint_as_float(v[e + 1]), ai, pi * va.y);
                                y[2] = fmaf(__uint_as_float(v[e + 2]), ai, pi * va.z);
                                y[3] = fmaf(__uint_as_float(v[e + 3]), ai, pi * va.w);
                                y[4] = fmaf(__uint_as_float(v[e + 4]), ai, pi * vb.x);
                                y[5] = fmaf(__uint_as_float(v[e + 5]), ai, pi * vb.y);
                                y[6] = fmaf(__uint_as_float(v[e + 6]), ai, pi * vb.z);
                                y[7] = fmaf(__uint_as_float(v[e + 7]), ai, pi * vb.w);
                            } else {
#pragma unroll
                                for (int u = 0; u < 8; ++u) y[u] = __uint_as_float(v[e + u]) * inv;
                            }
                            uint4 wv;
                            wv.x = pack_bf16x2(y[0], y[1]);
                            wv.y = pack_bf16x2(y[2], y[3]);
                            wv.z = pack_bf16x2(y[4], y[5]);
                            wv.w = pack_bf16x2(y[6], y[7]);
                            *reinterpret_cast<uint4*>(o + c + e) = wv;
                        }
                    }
                }
            }
            // the next item's first P V overwrites O only after this warpgroup's next p_full arrival, which follows these
            // tcgen05.ld in program order
            tc_fence_before();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        tmem_dealloc<G::TMEM_COLS>(tmem_base);
    }
}

template <int HD>
cudaError_t launch_attention2_hd(const CUtensorMap& tm_qkv, const __nv_bfloat16* qkv, __nv_bfloat16* out, int frames,
                                 int S, int heads, float softmax_scale, int num_sms, cudaStream_t stream) {
    static unsigned long long optin_mask = 0;
    cudaError_t e = ensure_smem_optin(reinterpret_cast<const void*>(&attention_fwd2_kernel<HD>), Geo2<HD>::SMEM_BYTES,
                                      optin_mask);
    if (e != cudaSuccess) return e;
    // side mode: S = k * 256 + 1 (k >= 1) -- the leftover query row rides on the side warps of chunk 0
    const int side_mode = (S > CHUNK && S % CHUNK == 1 && qkv != nullptr) ? 1 : 0;
    const int n_chunks = side_mode ? S / CHUNK : (S + CHUNK - 1) / CHUNK;
    const long long n_items = static_cast<long long>(frames) * heads * n_chunks;
    if (n_items > 0x3fffffff) return cudaErrorInvalidValue;
    int grid = num_sms > 0 ? num_sms : 148;
    if (n_items < grid) grid = static_cast<int>(n_items);
    int rot_div = grid / n_chunks;
    if (rot_div < 1) rot_div = 1;
    const float scale_log2 = softmax_scale * 1.4426950408889634f;
    attention_fwd2_kernel<HD><<<grid, THREADS, Geo2<HD>::SMEM_BYTES, stream>>>(
        tm_qkv, qkv, out, S, heads, static_cast<int>(n_items), n_chunks, rot_div, side_mode, scale_log2);
    return cudaGetLastError();
}

__global__ void exp2_poly_test_kernel(const float* __restrict__ x, float* __restrict__ y, int n) {
    const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i + 1 < n) {
        exp2_poly2(pack2(x[i], x[i + 1]), y[i], y[i + 1]);
    } else if (i < n) {
        float dummy;
        exp2_poly2(pack2(x[i], 0.f), y[i], dummy);
    }
}

}  // namespace

cudaError_t launch_exp2_poly_test(const float* x, float* y, int n, cudaStream_t stream) {
    const int pairs = (n + 1) / 2;
    exp2_poly_test_kernel<<<(pairs + 255) / 256, 256, 0, stream>>>(x, y, n);
    return cudaGetLastError();
}

int attention2_box_rows() { return BQ; }  // Q, K and V tiles are all [128 rows x 64 columns] boxes of the qkv matrix

cudaError_t launch_attention2(const CUtensorMap& tm_qkv, const __nv_bfloat16* qkv, __nv_bfloat16* out, int frames, int S,
                              int heads, int head_dim, float softmax_scale, int num_sms, cudaStream_t stream) {
    if (frames <= 0 || S <= 0 || heads <= 0) return cudaErrorInvalidValue;
    if (head_dim == 64) return launch_attention2_hd<64>(tm_qkv, qkv, out, frames, S, heads, softmax_scale, num_sms, stream);
    if (head_dim == 128) return launch_attention2_hd<128>(tm_qkv, qkv, out, frames, S, heads, softmax_scale, num_sms, stream);
    return cudaErrorInvalidValue;
}

}  // namespace aigv
