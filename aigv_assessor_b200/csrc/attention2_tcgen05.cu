// Persistent fused bidirectional flash-style attention on tcgen05/TMEM for sm_100a (head_dim 64 and 128), round-2 form.
//
// Replaces FlashAttention.forward -> flash_attn_varlen_qkvpacked_func (flash_attention.py:52-55, non-causal,
// dropout 0, softmax_scale = d^-0.5) reached from InternAttention._flash_attn (modeling_intern_vit.py:162-177), and is
// numerically the softmax(q*scale @ k^T) @ v of _naive_attn (modeling_intern_vit.py:143-157).
//
// What changed against the round-1 kernel (attention_tcgen05.cu) and why (profiles/r1f_ncu_attn_summary.txt: XU 54%,
// tensor 27%, 39% of the softmax warps' samples waiting for S): the per-tile chain softmax(j) -> P V(j) -> Q K^T(j+1)
// was serial because P aliased the only S buffer of a tile, so every block paid a commit -> mbarrier -> tcgen05.ld
// round trip that four warps per scheduler could not cover.
//   * ONE persistent CTA per SM walks a list of work items (frame, head, TILES x 128 query rows). Barriers, TMEM and
//     descriptors are set up once; the loads / Q K^T of item i+1 overlap the epilogue of item i.
//   * More S buffers than tiles, rotating: the MMA warp keeps Q K^T NSB (tile, block) steps ahead of P V, so when a
//     warpgroup has written P(j) its S(j+1) is already in TMEM -- the tensor-pipe latency leaves the per-tile chain.
//     head_dim 64:  3 tiles x 64-key blocks, 5 S/P buffers  (5 x 64 + 3 x 64 O columns = 512 TMEM columns); three
//                   softmax warps per scheduler, one query row per thread, the block's 64 scores in registers
//                   (one tcgen05.ld pass instead of round 1's two);
//     head_dim 128: 2 tiles x 128-key blocks, 2 S/P buffers (2 x 128 + 2 x 128), the alternating schedule.
//     (Measured on the way, profiles/r2_attention_notes.txt: with two tiles x 128-key blocks x 3 buffers at d = 64 the
//     period of a scheduler's two softmax warps was N + E -- the non-MUFU part N of a block, ~1700 cycles, never hid
//     under the partner's exponentials E -- 390 us against 451 us for round 1; a third warp per scheduler hides N.)
//   * The softmax warps of a scheduler take turns in their exponential phase (named barriers): one warp's tcgen05.ld /
//     row max / tcgen05.st / mbarrier traffic runs under another's MUFU stream instead of all competing for it.
//   * The MMA issuer's instruction stream sits between a warpgroup's P and its later S: ring indices / parities are
//     running counters, descriptors one add from constants, P V(g) and Q K^T(g + NSB) leave from one elected region.
//   * AIGV_ATTN_POLY (default 0) of every 4 pairs of exponentials can be evaluated on the FMA pipe (Cody-Waite split +
//     cubic minimax polynomial of 2^f on [-0.5, 0.5], |rel err| <= 7.5e-5). Measured: it LOSES here -- the packed FMA
//     work costs more issue slots than the MUFU slots it frees -- so it is off; aigv_exp2_poly keeps it tested.
//   warp 4T: TMA producer   warp 4T+1: MMA issuer   warps 4T+2/3: the class-token query row (see below)
//   warps 4t..4t+3: softmax warpgroup of query tile t, one query row per thread (TMEM lane == row).
// O lives in TMEM for the whole KV loop of an item. m_ref is a *stale* row maximum: it is only raised (and O, l
// rescaled through tcgen05.ld/st, after waiting for the tile's previous P V) when the true running maximum exceeds it
// by more than 2^8, so p <= 256 always and the rescale is rare; the final O / l is exact regardless of the reference.
// Ragged ends. Keys: S % BKV == 1 (every ViT sequence: 1024 patches + the class token): the lone tail key never
// enters the tensor pipeline -- its k / v rows sit in shared memory as fp32 and every query row folds it in during its
// epilogue (a head_dim-term dot product with its own Q row, an exact max/alpha update, O*alpha + p*v). Any other tail
// is one more (masked) block. Queries: S % 128 == 1: the class-token row rides on the two side warps of the CTA that
// owns the first rows of that (frame, head) -- fp32 online softmax on the CUDA cores from the very K/V stages the
// tensor pipe consumes, half of every block's keys per warp -- so no tile is spent on one row.
// A masked tail block multiplies P == 0 with V rows of the NEXT frame (TMA zero-fills past the end of the matrix);
// with finite activations that contributes exactly nothing.
#include <type_traits>

#include "kernels.h"
#include "ptx.cuh"

#ifndef AIGV_ATTN_POLY
#define AIGV_ATTN_POLY 0   // pairs out of every 4 pairs of exponentials evaluated on the FMA pipe (0..4)
#endif
#ifndef AIGV_ATTN_TILES64
#define AIGV_ATTN_TILES64 2   // head_dim 64 geometry: 2 = two tiles x 128-key blocks x 3 S buffers; 3 = three tiles x 64 x 5
#endif
#ifndef AIGV_ATTN_MMA1T
#define AIGV_ATTN_MMA1T 0   // 1: one elected thread runs the whole MMA-issuer loop (no per-step elect / reconvergence)
#endif
#ifndef AIGV_ATTN_PINGPONG
#define AIGV_ATTN_PINGPONG 1   // the softmax warps of an SM sub-partition take turns in their exponential phase
#endif

namespace aigv {

namespace {

constexpr int BQ = 128;                    // query rows per tile
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

template <int HD>
struct Geo2 {
    static constexpr bool T3 = HD == 64 && AIGV_ATTN_TILES64 == 3;
    static constexpr int TILES = T3 ? 3 : 2;              // query tiles per work item
    static constexpr int BKV = T3 ? 64 : 128;             // keys per block
    static constexpr int NSB = T3 ? 5 : (HD == 64 ? 3 : 2);  // S/P buffers in TMEM (>= TILES)
    static constexpr int PVR = NSB <= 2 ? 2 : (NSB <= 4 ? 4 : 8);  // ring of "P V done" barriers per tile (>= NSB, power of 2)
    static constexpr int CHUNK = BQ * TILES;
    static constexpr int SOFTMAX_WARPS = 4 * TILES;
    static constexpr int WARP_PRODUCER = SOFTMAX_WARPS, WARP_MMA = SOFTMAX_WARPS + 1, WARP_SIDE = SOFTMAX_WARPS + 2;
    static constexpr int THREADS = (SOFTMAX_WARPS + 4) * 32;  // whole warpgroups (setmaxnreg works per warpgroup)
    // registers after setmaxnreg: 12*32*144 + 4*32*80 = 65536 (launch: 512 x 128); 8*32*208 + 4*32*88 = 64512 (384 x 168)
    static constexpr int REGS_SOFTMAX = T3 ? 144 : 208;
    static constexpr int REGS_SERVICE = T3 ? 80 : 88;
    static constexpr int ATOMS = HD / 64;                 // 64-column (128-byte) swizzle atoms per operand row
    static constexpr int Q_ATOM = BQ * 128;               // 16 KiB
    static constexpr int KV_ATOM = BKV * 128;             // 8 / 16 KiB
    static constexpr int Q_BYTES = ATOMS * Q_ATOM;
    static constexpr int KV_BYTES = ATOMS * KV_ATOM;
    static constexpr int STAGES = T3 ? 6 : (HD == 64 ? 4 : 2);  // K / V ring depth
    static constexpr int QBUFS = HD == 64 ? 2 : 1;        // Q tiles of the next item land while this one computes
    static constexpr int KPW = BKV / 2;                   // keys of a block per side warp
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = OFF_Q + QBUFS * TILES * Q_BYTES;
    static constexpr int OFF_V = OFF_K + STAGES * KV_BYTES;
    static constexpr int OFF_BAR = OFF_V + STAGES * KV_BYTES;
    static constexpr int OFF_SIDE = OFF_BAR + 512;
    // fp32 side region: [2][2*HD] lone tail key k, v (ring of two items) | [2][HD] q of the class-token row (one copy
    // per side warp) | [2][KPW] its probabilities | [HD + 32] merge buffer of the second side warp
    static constexpr int SIDE_FLOATS = 2 * 2 * HD + 2 * HD + 2 * KPW + HD + 32;
    static constexpr int SMEM_BYTES = OFF_SIDE + SIDE_FLOATS * 4;
    static constexpr uint32_t TM_O = NSB * BKV;           // O accumulators behind the S buffers
    static constexpr uint32_t TMEM_COLS = 512;
    static_assert(NSB * BKV + TILES * HD <= 512, "TMEM budget");
    static_assert(NSB >= TILES, "every tile needs an S buffer of its own");
    static_assert((2 * QBUFS + 4 * STAGES + 2 * NSB + TILES * PVR + TILES) * 8 + 8 <= 512, "barrier region");
    static_assert(2 + 4 * TILES <= 16, "named barriers of the exponential turn");
};

struct Item {
    int frame, head, q0, row_base, n_t, n_b;
    bool side;
};

// Where q / k / v of head h live in a row of the fused projection output (see AttnLayout in kernels.h).
__device__ __forceinline__ int col_q(const AttnLayout& L, int h, int HD) {
    return (h / L.q_group) * L.grp_stride + (h % L.q_group) * HD;
}
__device__ __forceinline__ int col_k(const AttnLayout& L, int h) {
    return (h / L.q_group) * L.grp_stride + L.k_off + (h % L.q_group) * L.k_mul;
}

// item index -> (frame, head, chunk). Items of one (frame, head) are adjacent (their CTAs share K/V in L2); the chunk is
// rotated by the pass number so the CTA that gets the class-token row changes from pass to pass. s_rows = rows that
// tiles cover (the sequence without the class-token side row).
template <int CHUNK, int TILES, int BKV>
__device__ __forceinline__ Item decode_item(int i, int n_chunks, int rot_div, int H, int S, int s_rows, bool side_mode,
                                            int n_blocks, bool causal, int n_items_total) {
    Item m;
    int fh, chunk;
    if (causal) {   // rank i: all (frame, head) of the last chunk first, then the chunk before it, ...
        const int FH = n_items_total / n_chunks;
        const int c = i / FH;
        fh = i - c * FH;
        chunk = n_chunks - 1 - c;
    } else {
        fh = i / n_chunks;
        const int raw = i - fh * n_chunks;
        chunk = (raw + fh / rot_div) % n_chunks;
    }
    m.frame = fh / H;
    m.head = fh - m.frame * H;
    m.q0 = chunk * CHUNK;
    m.row_base = m.frame * S;
    const int left = s_rows - m.q0;
    m.n_t = left >= CHUNK ? TILES : (left + BQ - 1) / BQ;
    m.side = side_mode && chunk == 0;
    // causal: keys up to the item's last query row only (blocks past a tile's own diagonal are masked, not skipped)
    m.n_b = causal ? min(n_blocks, (min(m.q0 + m.n_t * BQ, S) + BKV - 1) / BKV) : n_blocks;
    return m;
}

// Work list. Bidirectional: CTA b takes items b, b + grid, ... (items of one (frame, head) adjacent: their CTAs share K/V
// in L2). Causal: items are ranked heaviest first (a chunk's cost grows with its position in the sequence) and dealt out
// in a snake -- pass 0 to CTAs 0..grid-1, pass 1 to grid-1..0, ... -- so every CTA's total is close to the average instead
// of up to twice it. Returns the item id of this CTA's it-th item, or -1 past its last one.
template <bool kCausal>
__device__ __forceinline__ int item_at(int it, int n_items) {
    const int grid = static_cast<int>(gridDim.x), b = static_cast<int>(blockIdx.x);
    const int i = it * grid + ((kCausal && (it & 1)) ? grid - 1 - b : b);
    return i < n_items ? i : -1;
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

template <int HD, bool kCausal>
__global__ void __launch_bounds__(Geo2<HD>::THREADS, 1)
attention_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                      const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int S, int H, int n_items,
                      int n_chunks, int rot_div, int side_mode, float scale_log2, const AttnLayout lay) {
    using G = Geo2<HD>;
    constexpr int TILES = G::TILES, BKV = G::BKV, CHUNK = G::CHUNK, SOFTMAX_WARPS = G::SOFTMAX_WARPS,
                  WARP_PRODUCER = G::WARP_PRODUCER, WARP_MMA = G::WARP_MMA, WARP_SIDE = G::WARP_SIDE, KPW = G::KPW;
    constexpr int ATOMS = G::ATOMS, Q_ATOM = G::Q_ATOM, KV_ATOM = G::KV_ATOM, Q_BYTES = G::Q_BYTES,
                  KV_BYTES = G::KV_BYTES, NSB = G::NSB, STAGES = G::STAGES, QBUFS = G::QBUFS, OFF_Q = G::OFF_Q,
                  OFF_K = G::OFF_K, OFF_V = G::OFF_V, PVR = G::PVR;
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
    // pv_done[t][k % PVR]: the k-th P V of tile t has completed (O quiescent until the tile's next P). A RING of barriers
    // per tile, because a parity wait only tells "this phase or the one before": Q K^T runs up to NSB steps ahead, so
    // when a warpgroup starts block k it knows P V(k - PVR) is complete and nothing about the ones between. With one
    // barrier a rescale at block k could take "P V(k-3) done" for "P V(k-1) done" (found by the parity tests of the
    // three-tile geometry: a wrong tile in items with fewer live tiles than S buffers).
    uint64_t* pv_done = p_full + NSB;        // [TILES][PVR]
    uint64_t* o_full = pv_done + TILES * PVR;  // [TILES] last P V of the item accumulated
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + TILES);
    float* side_f = reinterpret_cast<float*>(smem + G::OFF_SIDE);
    float* kt_ring = side_f;                  // [2][2*HD]: k then v of the lone tail key, item parity
    float* q_side = kt_ring + 2 * 2 * HD;     // [2][HD]
    float* p_side = q_side + 2 * HD;          // [2][KPW]
    float* merge = p_side + 2 * KPW;          // [HD + 32]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_full = S / BKV;
    const int tail_valid = S - n_full * BKV;
    constexpr bool causal = kCausal;  // a compile-time variant: the bidirectional (ViT) instantiation carries none of it
    const bool tail_epi = tail_valid == 1 && n_full >= 1 && !causal;
    const int n_blocks = n_full + ((tail_valid > 0 && !tail_epi) ? 1 : 0);
    const bool smode = side_mode != 0;
    const int s_rows = S - (smode ? 1 : 0);  // query rows that tiles cover

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // dynamic smem base must be 1024-byte aligned
    if (warp == WARP_PRODUCER && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_kv);
    }
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
                for (int k = 0; k < PVR; ++k) mbar_init(&pv_done[t * PVR + k], 1);
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
    pdl_wait();               // qkv (the previous kernel's output) is complete from here on; the set-up above overlapped its tail
    pdl_launch_dependents();

    if (warp >= SOFTMAX_WARPS) {
        reg_dec<G::REGS_SERVICE>();
        if (warp == WARP_PRODUCER) {
            // ------------------------------------------------------------ TMA producer
            uint32_t kst = 0, kph = 0;  // K/V ring position of the next block to load
            for (int it = 0, i; (i = item_at<kCausal>(it, n_items)) >= 0; ++it) {
                const Item m = decode_item<CHUNK, TILES, BKV>(i, n_chunks, rot_div, H, S, s_rows, smode, n_blocks, causal, n_items);
                const int qb = it % QBUFS;
                auto load_k = [&](int j) {
                    const int r = m.row_base + j * BKV;
                    mbar_wait(&k_empty[kst], kph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&k_full[kst], KV_BYTES);
                        for (int a = 0; a < ATOMS; ++a)
                            tma_load_2d(smem + OFF_K + kst * KV_BYTES + a * KV_ATOM, &tm_kv, &k_full[kst],
                                        col_k(lay, m.head) + a * 64, r);
                    }
                    __syncwarp();
                };
                auto load_v = [&](int j) {   // also advances the ring
                    const int r = m.row_base + j * BKV;
                    mbar_wait(&v_empty[kst], kph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&v_full[kst], KV_BYTES);
                        for (int a = 0; a < ATOMS; ++a)
                            tma_load_2d(smem + OFF_V + kst * KV_BYTES + a * KV_ATOM, &tm_kv, &v_full[kst],
                                        col_k(lay, m.head) + lay.v_off + a * 64, r);
                    }
                    __syncwarp();
                    if (++kst == static_cast<uint32_t>(STAGES)) { kst = 0; kph ^= 1; }
                };
                // With a single Q buffer the Q tiles can only land after the previous item's last Q K^T: have the first K
                // block in flight before waiting for that. Only K: the V stage it would take is released by a P V of the
                // previous item that the MMA warp issues together with THIS item's first Q K^T when that item had a single
                // live tile (Q K^T then runs two blocks ahead), i.e. only after these Q tiles have landed -- waiting for
                // that V stage before loading Q deadlocks (found by the first multi-item run with ragged 128-wide heads).
                const bool early_k = QBUFS == 1 && STAGES >= 2;
                if (early_k) load_k(0);
                mbar_wait(&q_empty[qb], ((it / QBUFS) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&q_full[qb], m.n_t * Q_BYTES);
                    for (int t = 0; t < m.n_t; ++t)
                        for (int a = 0; a < ATOMS; ++a)
                            tma_load_2d(smem + OFF_Q + (qb * TILES + t) * Q_BYTES + a * Q_ATOM, &tm_q, &q_full[qb],
                                        col_q(lay, m.head, HD) + a * 64, m.row_base + m.q0 + t * BQ);
                }
                __syncwarp();
                if (tail_epi) {
                    float* kt = kt_ring + (it & 1) * 2 * HD;
                    const __nv_bfloat16* kg =
                        qkv + static_cast<size_t>(m.row_base + S - 1) * lay.row_pitch + col_k(lay, m.head);
#pragma unroll
                    for (int e = 0; e < DPL; ++e) {
                        kt[lane * DPL + e] = __bfloat162float(kg[lane * DPL + e]);
                        kt[HD + lane * DPL + e] = __bfloat162float(kg[lay.v_off + lane * DPL + e]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&q_full[qb]);
                for (int j = 0; j < m.n_b; ++j) {
                    if (j > 0 || !early_k) load_k(j);
                    load_v(j);
                }
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
            int my_items = 0;
            while (item_at<kCausal>(my_items, n_items) >= 0) ++my_items;
            const bool uniform = (s_rows % CHUNK) == 0 && !causal;  // every item: all tiles live, all key blocks
            // (live tiles, key blocks) of this CTA's it-th item
            auto geom_of = [&](int it, int& nt, int& nb) {
                if (it >= my_items) { nt = 0; nb = n_blocks; return; }
                if (uniform) { nt = TILES; nb = n_blocks; return; }
                const Item m = decode_item<CHUNK, TILES, BKV>(item_at<kCausal>(it, n_items), n_chunks, rot_div, H, S, s_rows,
                                                              smode, n_blocks, causal, n_items);
                nt = m.n_t;
                nb = m.n_b;
            };
            // Q K^T cursor
            int q_it = 0, q_j = 0, q_t = 0, q_nt, q_nb;
            geom_of(0, q_nt, q_nb);
            uint32_t q_b = 0, q_kst = 0, q_kph = 0, q_qb = 0, q_qph = 0;
            // P V cursor
            int p_it = 0, p_j = 0, p_t = 0, p_nt = q_nt, p_nb = q_nb;
            uint32_t p_b = 0, p_bph = 0, p_vst = 0, p_vph = 0;
            uint32_t pv_cnt[TILES];  // P V steps issued per tile: slot of the tile's pv_done ring = count % PVR
#pragma unroll
            for (int t = 0; t < TILES; ++t) pv_cnt[t] = 0;
            auto step = [&](bool do_pv, bool do_qk) {
                if (do_pv) {
                    mbar_wait(&p_full[p_b], p_bph);  // S consumed, P in TMEM, O rescaled if it had to be
                    if (p_t == 0) mbar_wait(&v_full[p_vst], p_vph);
                }
                if (do_qk && q_t == 0) {
                    if (q_j == 0) mbar_wait(&q_full[q_qb], q_qph);
                    mbar_wait(&k_full[q_kst], q_kph);
                }
                tc_fence_after();
                const bool p_last_t = p_t == p_nt - 1, p_last_j = p_j == p_nb - 1;
                const bool q_last_t = q_t == q_nt - 1, q_last_j = q_j == q_nb - 1;
                const uint64_t vd = v_desc0 + ((p_vst * static_cast<uint32_t>(KV_BYTES)) >> 4);
                const uint32_t tm_p = tmem_u + p_b * BKV;  // bf16 pairs in the first BKV/2 columns of the S buffer
                const uint32_t tm_o = tmem_u + TM_O + p_t * HD;
                const uint64_t qd = q_desc0 + (((q_qb * TILES + q_t) * static_cast<uint32_t>(Q_BYTES)) >> 4);
                const uint64_t kd = k_desc0 + ((q_kst * static_cast<uint32_t>(KV_BYTES)) >> 4);
                const uint32_t tm_s = tmem_u + q_b * BKV;
                if (AIGV_ATTN_MMA1T || elect_one()) {
                    if (do_pv) {
#pragma unroll
                        for (int k = 0; k < BKV / 16; ++k)
                            // A = P from TMEM: 8 columns (16 keys) per step; B = V rows [16k, 16k+16): 2 KiB apart
                            umma_bf16_ts(tm_o, tm_p + k * 8, vd + ((k * 2048) >> 4), idesc_pv, (p_j | k) != 0 ? 1u : 0u);
                        uint32_t slot = 0;
#pragma unroll
                        for (int t = 0; t < TILES; ++t)
                            if (t == p_t) slot = pv_cnt[t] & (PVR - 1);
                        umma_commit(&pv_done[p_t * PVR + slot]);
                        if (p_last_t) umma_commit(&v_empty[p_vst]);
                        if (p_last_j) umma_commit(&o_full[p_t]);
                    }
                    if (do_qk) {
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k)  // d axis: 4 steps of 32 B inside each 64-column atom
                            umma_bf16_ss(tm_s, qd + (((k >> 2) * Q_ATOM + (k & 3) * 32) >> 4),
                                         kd + (((k >> 2) * KV_ATOM + (k & 3) * 32) >> 4), idesc_qk, k != 0 ? 1u : 0u);
                        umma_commit(&s_full[q_b]);
                        if (q_last_t) {
                            umma_commit(&k_empty[q_kst]);
                            if (q_last_j) umma_commit(&q_empty[q_qb]);
                        }
                    }
                }
                if (!AIGV_ATTN_MMA1T) __syncwarp();
                if (do_pv) {
#pragma unroll
                    for (int t = 0; t < TILES; ++t)
                        if (t == p_t) ++pv_cnt[t];
                    if (++p_b == static_cast<uint32_t>(NSB)) { p_b = 0; p_bph ^= 1; }
                    if (p_last_t) {
                        p_t = 0;
                        if (++p_vst == static_cast<uint32_t>(STAGES)) { p_vst = 0; p_vph ^= 1; }
                        if (p_last_j) {
                            p_j = 0;
                            ++p_it;
                            geom_of(p_it, p_nt, p_nb);
                        } else {
                            ++p_j;
                        }
                    } else {
                        ++p_t;
                    }
                }
                if (do_qk) {
                    if (++q_b == static_cast<uint32_t>(NSB)) q_b = 0;
                    if (q_last_t) {
                        q_t = 0;
                        if (++q_kst == static_cast<uint32_t>(STAGES)) { q_kst = 0; q_kph ^= 1; }
                        if (q_last_j) {
                            q_j = 0;
                            ++q_it;
                            if (++q_qb == static_cast<uint32_t>(QBUFS)) { q_qb = 0; q_qph ^= 1; }
                            geom_of(q_it, q_nt, q_nb);
                        } else {
                            ++q_j;
                        }
                    } else {
                        ++q_t;
                    }
                }
            };
            if (!AIGV_ATTN_MMA1T || elect_one()) {
                for (int pre = 0; pre < NSB && q_nt; ++pre) step(false, true);
                while (p_nt) step(true, q_nt != 0);
            }
            __syncwarp();
        } else {
            // ------------------------------------------------------------ the class-token query row (side warps)
            // Both warps consume every K/V stage (the ring's empty barriers expect them); on items that own the
            // class-token row each warp runs an fp32 online softmax over its KPW keys of every block.
            constexpr int KPL = KPW / 32;  // keys per lane
            const int w = warp - WARP_SIDE;
            float* qs = q_side + w * HD;
            float* ps = p_side + w * KPW;
            uint32_t kst = 0, kph = 0;
            for (int it = 0, i; (i = item_at<kCausal>(it, n_items)) >= 0; ++it) {
                const Item m = decode_item<CHUNK, TILES, BKV>(i, n_chunks, rot_div, H, S, s_rows, smode, n_blocks, causal, n_items);
                float m_run = -INFINITY, l_lane = 0.f, o_acc[DPL];
#pragma unroll
                for (int e = 0; e < DPL; ++e) o_acc[e] = 0.f;
                const __nv_bfloat16* qg = qkv + static_cast<size_t>(m.row_base + S - 1) * lay.row_pitch + col_q(lay, m.head, HD);
                if (m.side) {
#pragma unroll
                    for (int e = 0; e < DPL; ++e) qs[lane * DPL + e] = __bfloat162float(qg[lane * DPL + e]) * scale_log2;
                    __syncwarp();
                }
                for (int j = 0; j < m.n_b; ++j) {
                    mbar_wait(&k_full[kst], kph);
                    float corr = 1.f;
                    if (m.side) {  // side mode: every block is full (S = 128 k + 1, the tail key is folded at the end)
                        const uint8_t* kb = smem + OFF_K + kst * KV_BYTES;
                        float sc[KPL];
#pragma unroll
                        for (int h = 0; h < KPL; ++h) {
                            const int key = w * KPW + lane + 32 * h;
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
                        float mb = sc[0];
#pragma unroll
                        for (int h = 1; h < KPL; ++h) mb = fmaxf(mb, sc[h]);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
                        const float m_new = fmaxf(m_run, mb);
                        corr = ex2f(m_run - m_new);
                        float psum = 0.f;
#pragma unroll
                        for (int h = 0; h < KPL; ++h) {
                            const float pv = ex2f(sc[h] - m_new);
                            ps[lane + 32 * h] = pv;
                            psum += pv;
                        }
                        l_lane = fmaf(l_lane, corr, psum);
                        m_run = m_new;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&k_empty[kst]);
                    mbar_wait(&v_full[kst], kph);
                    if (m.side) {
                        const uint8_t* vb = smem + OFF_V + kst * KV_BYTES + ((lane * DPL) >> 6) * KV_ATOM;
                        const int boff = ((lane * DPL) & 63) * 2;  // byte offset of this lane's dims inside the 128-byte row
#pragma unroll
                        for (int e = 0; e < DPL; ++e) o_acc[e] *= corr;
#pragma unroll 4
                        for (int kk = 0; kk < KPW; ++kk) {
                            const int key = w * KPW + kk;
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
                    if (lane == 0) mbar_arrive(&v_empty[kst]);
                    if (++kst == static_cast<uint32_t>(STAGES)) { kst = 0; kph ^= 1; }
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
                            const __nv_bfloat16* kg =
                                qkv + static_cast<size_t>(m.row_base + S - 1) * lay.row_pitch + col_k(lay, m.head);
                            float stl = 0.f, vt[DPL];
#pragma unroll
                            for (int e = 0; e < DPL; ++e) {
                                stl = fmaf(qs[lane * DPL + e], __bfloat162float(kg[lane * DPL + e]), stl);
                                vt[e] = __bfloat162float(kg[lay.v_off + lane * DPL + e]);
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
        reg_inc<G::REGS_SOFTMAX>();
        const int t = warp >> 2;                      // query tile of this warpgroup
        const int wq = warp & 3;                      // TMEM lane quadrant
        const int r = wq * 32 + lane;                 // row within the tile == TMEM lane
        // Exponential turns: warps wq, wq + 4, ... (tiles 0, 1, ...) sit on the same scheduler; named barrier
        // 2 + TILES*wq + t is "tile t's turn" of that group. A warp waits for its turn before its exponentials and hands
        // the turn to the next tile right after them. Tile 0 goes first.
        const uint32_t pp_mine = 2 + TILES * wq + t, pp_next = 2 + TILES * wq + (t + 1 == TILES ? 0 : t + 1);
        if (AIGV_ATTN_PINGPONG && t == TILES - 1) named_bar_arrive(pp_next, 64);
        const uint32_t tm_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
        const uint32_t tm_o = tm_lane + TM_O + t * HD;
        uint32_t sb = 0, sph = 0;  // S buffer / parity of this tile's next step; advanced by n_t per block
        uint32_t n_pv = 0;         // P V steps of this tile so far (phase of pv_done)
        uint32_t n_items_t = 0;    // items this tile took part in (phase of o_full)
        int it = 0;
        auto advance_sb = [&](uint32_t steps) {
            sb += steps;
            while (sb >= static_cast<uint32_t>(NSB)) { sb -= NSB; sph ^= 1; }
        };
        advance_sb(t);  // tile t's first step is step t of the sequence (when all tiles are live; see below)
        for (int i; (i = item_at<kCausal>(it, n_items)) >= 0; ++it) {
            const Item m = decode_item<CHUNK, TILES, BKV>(i, n_chunks, rot_div, H, S, s_rows, smode, n_blocks, causal, n_items);
            const int qb = it % QBUFS;
            const uint32_t qph = (it / QBUFS) & 1;
            // Step index of (block j, tile t) inside this item is j * n_t + t: sb/sph were positioned for tile t with all
            // TILES live (offset t from the item's first step); with fewer live tiles the offset is the same t (t < n_t).
            if (t >= m.n_t) {  // this warpgroup's tile holds no query row: only release Q (and keep passing the turn)
                mbar_wait(&q_full[qb], qph);
                if (lane == 0) mbar_arrive(&q_empty[qb]);
                advance_sb(static_cast<uint32_t>(m.n_b * m.n_t));
                if (AIGV_ATTN_PINGPONG)
                    for (int j = 0; j < m.n_b; ++j) {
                        named_bar_sync(pp_mine, 64);
                        named_bar_arrive(pp_next, 64);
                    }
                continue;
            }
            const int tile_q0 = m.q0 + t * BQ;
            // Warps whose 32 query rows are all beyond the sequence skip the softmax math; they still take part in
            // the barrier protocol so the MMA warp's schedule is unchanged.
            const bool warp_live = tile_q0 + wq * 32 < s_rows;
            float m_ref = -INFINITY;  // reference maximum, already multiplied by scale*log2e
            float l_run = 0.f;
            float s_t = 0.f;          // the lone tail key's score (log2 units)

            // One (tile, block) step. kMasked selects the copy that masks a ragged tail block: kept out of the hot loop (its
            // 128 selects sat between the tcgen05.ld and the row max of EVERY block, jumped over).
            auto softmax_block = [&](const int j, auto masked_tag) {
                constexpr bool kMasked = decltype(masked_tag)::value;
                const uint32_t b = sb;
                const uint32_t tm_s = tm_lane + b * BKV;
                mbar_wait(&s_full[b], sph);
                tc_fence_after();
                if (warp_live) {
                    uint32_t v[BKV];
#pragma unroll
                    for (int c = 0; c < BKV; c += 32) tmem_ld_32x32(tm_s + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
                    tmem_ld_wait();
                    if (kMasked) {  // ragged tail block / causal diagonal and beyond (never taken for ViT sequences)
                        // keys of this block the row may attend to: all that exist, or those up to its own position
                        const int valid = causal ? tile_q0 + r - j * BKV + 1 : tail_valid;
#pragma unroll
                        for (int c = 0; c < BKV; ++c)
                            if (c >= valid) v[c] = 0xff800000u;
                    }
                    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                    for (int c = 0; c < BKV; c += 8) {
                        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v[c]), __uint_as_float(v[c + 1])));
                        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])));
                        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(v[c + 4]), __uint_as_float(v[c + 5])));
                        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(v[c + 6]), __uint_as_float(v[c + 7])));
                    }
                    const float mxs = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
                    const bool need = mxs > m_ref + RESCALE_THRESHOLD;  // always true for block 0 (m_ref = -inf)
                    if (__builtin_expect(__any_sync(0xffffffffu, need), 0)) {  // warp-uniform: TMEM ops are warp-collective
                        const float new_ref = need ? mxs : m_ref;
                        if (j > 0) {
                            // O must be quiescent: wait for this tile's previous P V (ring slot (n_pv-1) % PVR, use
                            // (n_pv-1) / PVR of that slot), then rescale O in place
                            mbar_wait(&pv_done[t * PVR + ((n_pv - 1) & (PVR - 1))], ((n_pv - 1) / PVR) & 1);
                            tc_fence_after();
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
                                    p[2 * q] = ex2f(x0);
                                    p[2 * q + 1] = ex2f(x1);
                                }
                            }
                            sum01 = add2(sum01, add2(pack2(p[0], p[1]), pack2(p[4], p[5])));
                            sum23 = add2(sum23, add2(pack2(p[2], p[3]), pack2(p[6], p[7])));
#pragma unroll
                            for (int q = 0; q < 4; ++q) pk[(e >> 1) + q] = pack_bf16x2(p[2 * q], p[2 * q + 1]);
                        }
                        tmem_st_32x16(tm_s + (c >> 1), pk);  // P over S columns this thread has already consumed
                    }
                    if (AIGV_ATTN_PINGPONG) named_bar_arrive(pp_next, 64);
                    tmem_st_wait();
                    float s0, s1, s2, s3;
                    unpack2(sum01, s0, s1);
                    unpack2(sum23, s2, s3);
                    l_run += (s0 + s1) + (s2 + s3);
                } else if (AIGV_ATTN_PINGPONG) {
                    named_bar_sync(pp_mine, 64);
                    named_bar_arrive(pp_next, 64);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[b]);  // one arrival per softmax warp
                ++n_pv;
                advance_sb(static_cast<uint32_t>(m.n_t));
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
            };
            {
                // blocks [0, jm) need no mask for any row of this warp: whole blocks of existing keys, and (causal) ending
                // at or before the warp's first query row
                const int jm = causal ? min(m.n_b, (tile_q0 + wq * 32 + 1) / BKV) : min(n_full, m.n_b);
                int j = 0;
                for (; j < jm; ++j) softmax_block(j, std::false_type{});
                for (; j < m.n_b; ++j) softmax_block(j, std::true_type{});
            }

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
                    if (q < s_rows) {
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
cudaError_t launch_attention2_hd(const CUtensorMap& tm_q, const CUtensorMap& tm_kv, const __nv_bfloat16* qkv,
                                 __nv_bfloat16* out, int frames, int S, int heads, float softmax_scale, int num_sms,
                                 cudaStream_t stream, const AttnLayout& lay) {
    using G = Geo2<HD>;
    static unsigned long long optin_mask[2] = {0, 0};
    auto kernel = lay.causal ? &attention_fwd2_kernel<HD, true> : &attention_fwd2_kernel<HD, false>;
    cudaError_t e = ensure_smem_optin(reinterpret_cast<const void*>(kernel), G::SMEM_BYTES, optin_mask[lay.causal ? 1 : 0]);
    if (e != cudaSuccess) return e;
    // side mode: S = k * 128 + 1 (k >= 1) -- the leftover query row rides on the side warps of the first chunk
    const int side_mode = (S > BQ && S % BQ == 1 && qkv != nullptr && !lay.causal) ? 1 : 0;
    const int s_rows = S - side_mode;
    const int n_chunks = (s_rows + G::CHUNK - 1) / G::CHUNK;
    const long long n_items = static_cast<long long>(frames) * heads * n_chunks;
    if (n_items > 0x3fffffff) return cudaErrorInvalidValue;
    int grid = num_sms > 0 ? num_sms : 148;
    if (n_items < grid) grid = static_cast<int>(n_items);
    int rot_div = grid / n_chunks;
    if (rot_div < 1) rot_div = 1;
    const float scale_log2 = softmax_scale * 1.4426950408889634f;
    return launch_pdl(kernel, dim3(grid), dim3(G::THREADS), G::SMEM_BYTES, stream, tm_q, tm_kv, qkv, out, S, heads,
                      static_cast<int>(n_items), n_chunks, rot_div, side_mode, scale_log2, lay);
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

int attention2_q_box_rows() { return BQ; }
int attention2_kv_box_rows(int head_dim) { return head_dim == 64 ? Geo2<64>::BKV : Geo2<128>::BKV; }

cudaError_t launch_attention2(const CUtensorMap& tm_q, const CUtensorMap& tm_kv, const __nv_bfloat16* qkv,
                              __nv_bfloat16* out, int frames, int S, int heads, int head_dim, float softmax_scale,
                              int num_sms, cudaStream_t stream, const AttnLayout* layout) {
    if (frames <= 0 || S <= 0 || heads <= 0) return cudaErrorInvalidValue;
    // default: the ViT's fused qkv output, channel order (three, head, d), bidirectional
    AttnLayout lay{3 * heads * head_dim, heads, 0, heads * head_dim, head_dim, heads * head_dim, 0};
    if (layout) lay = *layout;
    if (lay.q_group <= 0 || heads % lay.q_group != 0) return cudaErrorInvalidValue;
    if (head_dim == 64) return launch_attention2_hd<64>(tm_q, tm_kv, qkv, out, frames, S, heads, softmax_scale, num_sms, stream, lay);
    if (head_dim == 128) return launch_attention2_hd<128>(tm_q, tm_kv, qkv, out, frames, S, heads, softmax_scale, num_sms, stream, lay);
    return cudaErrorInvalidValue;
}

}  // namespace aigv
