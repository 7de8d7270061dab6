// InternLM2 prefill above the seam: the decoder stack the reference runs on the spliced
// inputs_embeds (/root/reference/internvl/model/internlm2/modeling_internlm2.py, "LLM" below), as a schedule of this
// library's tcgen05 kernels behind aigv_llm_* (include/aigv_b200.h).
//
//   per layer (LLM:635-695): x += wo( attention( rope( wqkv( rmsnorm(x) ) ) ) );  x += w2( silu(w1 h) * w3 h ), h = rmsnorm(x)
//     RMSNorm (LLM:129-143) is never materialised: gamma is folded into wqkv / w1 / w3 at weight load and the consumer
//     GEMM applies rstd from the row's sum of squares, which the producer GEMM (wo / w2 fused-residual epilogue, or the
//     seeding kernel for layer 0) left behind -- the same producer / consumer scheme as the ViT's LayerNorm fusion.
//     wqkv output keeps the reference's packing (kv group, G q heads + k + v, d) (LLM:373-389); rotary embedding
//     (LLM:161-196, 247-260; rotate-half form, positions 0..N-1) is applied in place to its q and k columns; attention
//     is the persistent tcgen05 kernel in its causal / grouped-query variant (K and V tiles are read per kv group,
//     LLM:404-405 repeat_kv never materialises); w1 and w3 are interleaved row by row into one GEMM whose output pairs
//     (gate, up) are adjacent for the SwiGLU (LLM:264-278).
//   after the stack: the final RMSNorm (LLM:984) and the 92 553-wide output head (LLM:1095) run only on the rows the caller
//     asks for (hidden_states[-1][:, -4] feeds mlpscore, eval2 :469-475; the answer span feeds the argmax, stage1_eval.py:
//     928-929) instead of on all B*N positions.
// Not covered (refused or documented): padding masks / position_ids other than 0..N-1 per sample (the eval runs B = 1
// unpadded prompts), the KV cache / decode steps, dynamic-NTK rescaling beyond max_position_embeddings (LLM:218-244 leaves
// the base unchanged below it).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/aigv_b200.h"
#include "kernels.h"
#include "ptx.cuh"

namespace aigv {
namespace {

// ------------------------------------------------------------------------------------------------ kernels
// Layer-0 producer: the spliced embeddings become the fp32 residual stream, its bf16 copy (the first wqkv's A operand)
// and the row's (sum, sum of squares) in statistics slice 0 (other slices zero). One warp per token.
__global__ void __launch_bounds__(256)
llm_seed_residual_kernel(const __nv_bfloat16* __restrict__ emb, float* __restrict__ x, __nv_bfloat16* __restrict__ xb,
                         float2* __restrict__ stats, int M, int H, int nparts) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const uint4* src = reinterpret_cast<const uint4*>(emb + static_cast<size_t>(row) * H);
    uint4* dstb = reinterpret_cast<uint4*>(xb + static_cast<size_t>(row) * H);
    float4* dst = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * H);
    float s = 0.f, q = 0.f;
    for (int i = lane; i < H / 8; i += 32) {
        const uint4 v = __ldg(src + i);
        dstb[i] = v;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            f[2 * e] = __uint_as_float(w[e] << 16);
            f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { s += f[e]; q = fmaf(f[e], f[e], q); }
        dst[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
        dst[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    for (int p = lane; p < nparts; p += 32) stats[static_cast<size_t>(p) * M + row] = p == 0 ? make_float2(s, q) : make_float2(0.f, 0.f);
}

// cos / sin of pos * theta^(-2j/d) (LLM:170-184: inv_freq and the outer product in fp32)
__global__ void llm_rope_table_kernel(float2* __restrict__ tab, int N, int half, float theta, float dim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * half) return;
    const int pos = i / half, j = i - pos * half;
    const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * j) / dim);
    const float a = static_cast<float>(pos) * inv_freq;
    float sn, cs;
    sincosf(a, &sn, &cs);
    tab[i] = make_float2(cs, sn);
}

// In place on the wqkv output [M, pitch]: every q head and the k head of each kv group (slots 0..G of the G + 2 per
// group) get out[j] = x[j] cos_j - x[j + d/2] sin_j, out[j + d/2] = x[j + d/2] cos_j + x[j] sin_j  (LLM:247-260), computed
// in fp32 from the bf16 values and rounded once. One thread per (token, rotary head, 8 consecutive j): 16-byte accesses.
__global__ void __launch_bounds__(256)
llm_rope_kernel(__nv_bfloat16* __restrict__ qkv, const float2* __restrict__ tab, int M, int N, int pitch, int n_kv, int G,
                int hd) {
    const int half = hd >> 1, octs = half >> 3;
    const int per_row = n_kv * (G + 1) * octs;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(M) * per_row) return;
    const int row = static_cast<int>(idx / per_row);
    int rem = static_cast<int>(idx - static_cast<long long>(row) * per_row);
    const int jo = rem % octs;
    rem /= octs;
    const int slot = rem % (G + 1), g = rem / (G + 1);
    const int pos = row % N;
    __nv_bfloat16* base = qkv + static_cast<size_t>(row) * pitch + (g * (G + 2) + slot) * hd;
    uint4* lo = reinterpret_cast<uint4*>(base + 8 * jo);
    uint4* hi = reinterpret_cast<uint4*>(base + half + 8 * jo);
    const uint4 a = *lo, b = *hi;
    const float4* cs = reinterpret_cast<const float4*>(tab + static_cast<size_t>(pos) * half + 8 * jo);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t ow[4], pw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float4 c = __ldg(cs + e);   // (cos, sin) of j = 8 jo + 2 e and of the next j
        const float a0 = __uint_as_float(aw[e] << 16), a1 = __uint_as_float(aw[e] & 0xffff0000u);
        const float b0 = __uint_as_float(bw[e] << 16), b1 = __uint_as_float(bw[e] & 0xffff0000u);
        ow[e] = pack_bf16x2(a0 * c.x - b0 * c.y, a1 * c.z - b1 * c.w);
        pw[e] = pack_bf16x2(b0 * c.x + a0 * c.y, b1 * c.z + a1 * c.w);
    }
    *lo = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    *hi = make_uint4(pw[0], pw[1], pw[2], pw[3]);
}

// SwiGLU on the interleaved w1 / w3 output: out[m, i] = silu(h[m, 2i]) * h[m, 2i + 1]  (LLM:275-278)
__global__ void __launch_bounds__(256)
llm_silu_mul_kernel(const __nv_bfloat16* __restrict__ h13, __nv_bfloat16* __restrict__ out, size_t n_out8) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_out8;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(h13) + 2 * i), b = __ldg(reinterpret_cast<const uint4*>(h13) + 2 * i + 1);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float g = __uint_as_float(w[e] << 16), u = __uint_as_float(w[e] & 0xffff0000u);
            y[e] = g / (1.0f + __expf(-g)) * u;
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]),
                                                      pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
    }
}

// Final RMSNorm (LLM:984) on selected rows of the residual stream: y = w * x * rsqrt(mean(x^2) + eps) -> fp32 and bf16.
__global__ void __launch_bounds__(256)
llm_norm_rows_kernel(const float* __restrict__ x, const int* __restrict__ rows, const float* __restrict__ w,
                     float* __restrict__ y, __nv_bfloat16* __restrict__ yb, int n, int M, int H, float eps) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const int row = rows[r];
    if (row < 0 || row >= M) return;
    const float* xr = x + static_cast<size_t>(row) * H;
    float q = 0.f;
    for (int i = lane; i < H; i += 32) q = fmaf(xr[i], xr[i], q);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / static_cast<float>(H) + eps);
    for (int i = lane; i < H; i += 32) {
        const float v = w[i] * (xr[i] * rstd);
        if (y) y[static_cast<size_t>(r) * H + i] = v;
        if (yb) yb[static_cast<size_t>(r) * H + i] = __float2bfloat16_rn(v);
    }
}

// Weight re-packing. dst[n, k] = bf16(src[n, k] * gamma[k]) (RMSNorm weight folded into the following Linear);
// rows of dst are written at n * row_mul + row_add (interleaves w1 / w3).
template <typename T>
__global__ void llm_pack_weight_kernel(const T* __restrict__ src, const float* __restrict__ gamma,
                                       __nv_bfloat16* __restrict__ dst, size_t rows, size_t K, int row_mul, int row_add) {
    const size_t n = rows * K;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = i / K, k = i - r * K;
        float v;
        if (sizeof(T) == 4) v = reinterpret_cast<const float*>(src)[i];
        else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
        if (gamma) v *= gamma[k];
        dst[(r * row_mul + row_add) * K + k] = __float2bfloat16_rn(v);
    }
}
template <typename T>
__global__ void llm_to_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        if (sizeof(T) == 4) dst[i] = reinterpret_cast<const float*>(src)[i];
        else dst[i] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
    }
}

struct LlmLayer {
    __nv_bfloat16 *wqkv = nullptr, *wo = nullptr, *w13 = nullptr, *w2 = nullptr;   // packed (gamma folded / interleaved)
    float *attn_norm = nullptr, *ffn_norm = nullptr;
    void *raw_wqkv = nullptr, *raw_w1 = nullptr, *raw_w3 = nullptr;               // as given, until finalize
    int raw_dtype[3] = {0, 0, 0};
    bool have[7] = {false, false, false, false, false, false, false};             // wqkv wo w1 w2 w3 attention_norm ffn_norm
    CUtensorMap tm_wqkv, tm_wo, tm_w13, tm_w2;
};

}  // namespace
}  // namespace aigv

using namespace aigv;

struct aigv_llm_handle {
    aigv_llm_cfg cfg{};
    int H = 0, nh = 0, nkv = 0, G = 0, hd = 0, I = 0, L = 0, V = 0, QW = 0, Vpad = 0, Mx = 0, parts = 0;
    int num_sms = 148;
    bool finalized = false;
    std::string err;
    std::vector<void*> allocs;
    std::vector<LlmLayer> layers;
    float* norm_w = nullptr;
    __nv_bfloat16* out_w = nullptr;
    bool have_norm = false, have_out = false;
    CUtensorMap tm_out_w;
    // workspace
    float* x = nullptr;
    __nv_bfloat16 *xb = nullptr, *qkv = nullptr, *attn = nullptr, *h13 = nullptr, *hm = nullptr, *sel_b = nullptr;
    float2* stats = nullptr;
    float2* rope = nullptr;
    int rope_n = 0;
    int* rows_dev = nullptr;
    float *sel_f = nullptr, *logits = nullptr;
    int cur_B = 0, cur_N = 0;
    int64_t launches = 0;
    // activation maps, rebuilt when the token count changes
    int maps_m = -1;
    CUtensorMap a_xb, a_attn, a_hm, q_map, kv_map, c_x, c_xb, c_qkv, c_h13, c_hm;
    bool fused_rope = true;     // AIGV_LLM_UNFUSED_ROPE=1: wqkv GEMM, then the standalone rotary kernel (A/B runs; always for heads != 128)
    bool fused_swiglu = true;   // AIGV_LLM_UNFUSED_SWIGLU=1: w1|w3 GEMM -> h13, then the standalone SwiGLU kernel (A/B runs)
};

namespace {

constexpr int kMaxRows = 512;   // rows per aigv_llm_hidden_rows / aigv_llm_logits_rows call

int lfail(aigv_llm_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}
int lcuda(aigv_llm_handle* h, cudaError_t e, const char* what) {
    return lfail(h, AIGV_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
template <typename T>
bool lalloc(aigv_llm_handle* h, T** p, size_t count) {
    void* q = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    if (cudaMalloc(&q, bytes) != cudaSuccess) return false;
    h->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return true;
}
void lfree(aigv_llm_handle* h, void* q) {
    if (!q) return;
    cudaFree(q);
    for (auto& a : h->allocs)
        if (a == q) { a = nullptr; break; }
}

int build_maps(aigv_llm_handle* h, int M) {
    if (h->maps_m == M) return 0;
    const int ba = gemm_box_rows_a(), bc = gemm_box_cols();
    const int cr = gemm_box_rows_c(), cb = gemm_box_cols_c(false), cf = gemm_box_cols_c(true);
    int r = 0;
    r |= make_tmap_bf16_2d(&h->a_xb, h->xb, M, h->H, h->H, ba, bc);
    r |= make_tmap_bf16_2d(&h->a_attn, h->attn, M, h->H, h->H, ba, bc);
    r |= make_tmap_bf16_2d(&h->a_hm, h->hm, M, h->I, h->I, ba, bc);
    r |= make_tmap_bf16_2d(&h->q_map, h->qkv, M, h->QW, h->QW, attention2_q_box_rows(), 64);
    r |= make_tmap_bf16_2d(&h->kv_map, h->qkv, M, h->QW, h->QW, attention2_kv_box_rows(h->hd), 64);
    r |= make_tmap_2d(&h->c_x, h->x, 1, M, h->H, h->H, cr, cf);
    r |= make_tmap_2d(&h->c_xb, h->xb, 0, M, h->H, h->H, cr, cb);
    r |= make_tmap_2d(&h->c_qkv, h->qkv, 0, M, h->QW, h->QW, cr, cb);
    r |= make_tmap_2d(&h->c_h13, h->h13, 0, M, 2 * h->I, 2 * h->I, cr, cb);
    r |= make_tmap_2d(&h->c_hm, h->hm, 0, M, h->I, h->I, cr, cb);
    if (r) return -1;
    h->maps_m = M;
    return 0;
}

// "language_model.model.layers.3.attention.wqkv.weight" -> layer 3, field "attention.wqkv.weight"
bool parse_layer(const std::string& key, int* layer, std::string* field) {
    const std::string tag = "model.layers.";
    const size_t p = key.find(tag);
    if (p == std::string::npos) return false;
    size_t q = p + tag.size(), e = q;
    while (e < key.size() && key[e] >= '0' && key[e] <= '9') ++e;
    if (e == q || e >= key.size() || key[e] != '.') return false;
    *layer = atoi(key.substr(q, e - q).c_str());
    *field = key.substr(e + 1);
    return true;
}

}  // namespace

extern "C" {

int aigv_llm_create(const aigv_llm_cfg* cfg, aigv_llm_handle** out) {
    if (!cfg || !out) return AIGV_ERR_INVALID;
    *out = nullptr;
    auto* h = new (std::nothrow) aigv_llm_handle();
    if (!h) return AIGV_ERR_CUDA;
    h->cfg = *cfg;
    h->H = cfg->hidden_size; h->nh = cfg->num_heads; h->nkv = cfg->num_kv_heads; h->I = cfg->intermediate_size;
    h->L = cfg->num_layers; h->V = cfg->vocab_size; h->Mx = cfg->max_tokens;
    auto bad = [&](int code, const char* msg) { h->err = msg; *out = h; return code; };
    if (h->H <= 0 || h->nh <= 0 || h->nkv <= 0 || h->I <= 0 || h->L <= 0 || h->V <= 0 || h->Mx <= 0)
        return bad(AIGV_ERR_INVALID, "aigv_llm_create: non-positive dimension");
    if (h->H % h->nh || h->nh % h->nkv) return bad(AIGV_ERR_INVALID, "hidden_size / num_heads / num_kv_heads do not divide");
    h->hd = h->H / h->nh;
    h->G = h->nh / h->nkv;
    if (h->hd != 64 && h->hd != 128) return bad(AIGV_ERR_UNSUPPORTED, "attention kernel covers head_dim 64 and 128");
    if (h->H % 64 || h->I % 32) return bad(AIGV_ERR_UNSUPPORTED, "hidden_size must be a multiple of 64, intermediate_size of 32");
    h->QW = (h->nh + 2 * h->nkv) * h->hd;
    h->Vpad = (h->V + 31) & ~31;
    h->parts = h->H / gemm_stats_slice_cols();
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    h->num_sms = sms > 0 ? sms : 148;
    h->layers.resize(h->L);
    {
        const char* e = getenv("AIGV_LLM_UNFUSED_SWIGLU");
        h->fused_swiglu = !(e && e[0] == '1') && (h->I % 64) == 0;
        const char* er = getenv("AIGV_LLM_UNFUSED_ROPE");
        h->fused_rope = !(er && er[0] == '1') && h->hd == 128;
    }
    const size_t M = h->Mx;
    bool ok = lalloc(h, &h->x, M * h->H) && lalloc(h, &h->xb, M * h->H) && lalloc(h, &h->qkv, M * h->QW) &&
              lalloc(h, &h->attn, M * h->H) && lalloc(h, &h->h13, M * 2 * h->I) && lalloc(h, &h->hm, M * h->I) &&
              lalloc(h, &h->stats, M * h->parts) && lalloc(h, &h->rows_dev, (size_t)kMaxRows) &&
              lalloc(h, &h->sel_f, (size_t)kMaxRows * h->H) && lalloc(h, &h->sel_b, (size_t)kMaxRows * h->H) &&
              lalloc(h, &h->logits, (size_t)kMaxRows * h->Vpad) && lalloc(h, &h->norm_w, (size_t)h->H) &&
              lalloc(h, &h->out_w, (size_t)h->V * h->H);
    for (auto& w : h->layers)
        ok = ok && lalloc(h, &w.wqkv, (size_t)h->QW * h->H) && lalloc(h, &w.wo, (size_t)h->H * h->H) &&
             lalloc(h, &w.w13, (size_t)2 * h->I * h->H) && lalloc(h, &w.w2, (size_t)h->H * h->I) &&
             lalloc(h, &w.attn_norm, (size_t)h->H) && lalloc(h, &w.ffn_norm, (size_t)h->H);
    *out = h;
    if (!ok) return lfail(h, AIGV_ERR_CUDA, "aigv_llm_create: device allocation failed");
    return AIGV_OK;
}

void aigv_llm_destroy(aigv_llm_handle* h) {
    if (!h) return;
    cudaDeviceSynchronize();
    for (void* q : h->allocs)
        if (q) cudaFree(q);
    for (auto& w : h->layers)   // raw copies of weights that were set but never finalized
        for (void* q : {w.raw_wqkv, w.raw_w1, w.raw_w3})
            if (q) cudaFree(q);
    delete h;
}

const char* aigv_llm_last_error(const aigv_llm_handle* h) { return h ? h->err.c_str() : ""; }
int64_t aigv_llm_launch_count(const aigv_llm_handle* h) { return h ? h->launches : 0; }

int aigv_llm_set_weight(aigv_llm_handle* h, const char* name, const void* ptr, int dtype, const int64_t* shape, int rank) {
    if (!h || !name || !ptr || !shape || rank <= 0) return lfail(h, AIGV_ERR_INVALID, "aigv_llm_set_weight: null argument");
    if (dtype != AIGV_F32 && dtype != AIGV_BF16) return lfail(h, AIGV_ERR_INVALID, "aigv_llm_set_weight: dtype");
    const std::string key(name);
    size_t count = 1;
    for (int i = 0; i < rank; ++i) count *= (size_t)shape[i];
    const size_t esz = dtype == AIGV_F32 ? 4 : 2;
    size_t expect = 0;
    int layer = -1;
    std::string field;
    enum { T_NONE, T_RAW, T_BF16, T_F32 } kind = T_NONE;
    void* dst = nullptr;
    void** raw_slot = nullptr;
    int* raw_dt = nullptr;
    bool* flag = nullptr;
    if (parse_layer(key, &layer, &field)) {
        if (layer < 0 || layer >= h->L) return lfail(h, AIGV_ERR_INVALID, "layer index out of range: " + key);
        LlmLayer& w = h->layers[layer];
        const size_t HH = (size_t)h->H;
        if (field == "attention.wqkv.weight") { kind = T_RAW; raw_slot = &w.raw_wqkv; raw_dt = &w.raw_dtype[0]; expect = (size_t)h->QW * HH; flag = &w.have[0]; }
        else if (field == "attention.wo.weight") { kind = T_BF16; dst = w.wo; expect = HH * HH; flag = &w.have[1]; }
        else if (field == "feed_forward.w1.weight") { kind = T_RAW; raw_slot = &w.raw_w1; raw_dt = &w.raw_dtype[1]; expect = (size_t)h->I * HH; flag = &w.have[2]; }
        else if (field == "feed_forward.w2.weight") { kind = T_BF16; dst = w.w2; expect = HH * h->I; flag = &w.have[3]; }
        else if (field == "feed_forward.w3.weight") { kind = T_RAW; raw_slot = &w.raw_w3; raw_dt = &w.raw_dtype[2]; expect = (size_t)h->I * HH; flag = &w.have[4]; }
        else if (field == "attention_norm.weight") { kind = T_F32; dst = w.attn_norm; expect = HH; flag = &w.have[5]; }
        else if (field == "ffn_norm.weight") { kind = T_F32; dst = w.ffn_norm; expect = HH; flag = &w.have[6]; }
    } else if (key.size() >= 17 && key.compare(key.size() - 17, 17, "model.norm.weight") == 0) {
        kind = T_F32; dst = h->norm_w; expect = (size_t)h->H; flag = &h->have_norm;
    } else if (key == "output.weight" || (key.size() >= 14 && key.compare(key.size() - 14, 14, ".output.weight") == 0)) {
        kind = T_BF16; dst = h->out_w; expect = (size_t)h->V * h->H; flag = &h->have_out;
    }
    if (kind == T_NONE) return lfail(h, AIGV_ERR_INVALID, "unknown weight name: " + key);
    if (count != expect)
        return lfail(h, AIGV_ERR_INVALID, "shape mismatch for " + key + ": got " + std::to_string(count) + " elements, expected " + std::to_string(expect));
    void* stage = nullptr;
    cudaError_t e = cudaMalloc(&stage, count * esz);
    if (e != cudaSuccess) return lcuda(h, e, "cudaMalloc(stage)");
    e = cudaMemcpy(stage, ptr, count * esz, cudaMemcpyDefault);
    if (e != cudaSuccess) { cudaFree(stage); return lcuda(h, e, "cudaMemcpy(weight)"); }
    if (kind == T_RAW) {     // folded with its RMSNorm weight at finalize
        if (*raw_slot) cudaFree(*raw_slot);
        *raw_slot = stage;
        *raw_dt = dtype;
        stage = nullptr;
    } else if (kind == T_BF16) {
        if (dtype == AIGV_BF16) cudaMemcpy(dst, stage, count * 2, cudaMemcpyDeviceToDevice);
        else llm_pack_weight_kernel<float><<<1024, 256>>>((const float*)stage, nullptr, (__nv_bfloat16*)dst, 1, count, 1, 0);
    } else {
        if (dtype == AIGV_F32) cudaMemcpy(dst, stage, count * 4, cudaMemcpyDeviceToDevice);
        else llm_to_f32_kernel<__nv_bfloat16><<<256, 256>>>((const __nv_bfloat16*)stage, (float*)dst, count);
    }
    e = cudaDeviceSynchronize();
    if (stage) cudaFree(stage);
    if (e != cudaSuccess) return lcuda(h, e, "weight conversion");
    *flag = true;
    h->finalized = false;
    return AIGV_OK;
}

int aigv_llm_finalize_weights(aigv_llm_handle* h, void* stream) {
    if (!h) return AIGV_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!h->have_norm) return lfail(h, AIGV_ERR_MISSING, "weight not set: model.norm.weight");
    static const char* names[7] = {"attention.wqkv.weight", "attention.wo.weight", "feed_forward.w1.weight", "feed_forward.w2.weight",
                                   "feed_forward.w3.weight", "attention_norm.weight", "ffn_norm.weight"};
    for (int l = 0; l < h->L; ++l)
        for (int f = 0; f < 7; ++f)
            if (!h->layers[l].have[f])
                return lfail(h, AIGV_ERR_MISSING, "weight not set: model.layers." + std::to_string(l) + "." + names[f]);
    const int bb = gemm_box_rows_b(), bc = gemm_box_cols();
    int r = 0;
    for (auto& w : h->layers) {
        // fold the RMSNorm weights: wqkv' = wqkv * diag(attention_norm), w13'[2i] = w1[i] * diag(ffn_norm), w13'[2i+1] = w3[i] * ...
        auto pack = [&](void* raw, int dt, const float* gamma, __nv_bfloat16* dst, size_t rows, int mul, int add) {
            if (!raw) return;   // already packed by an earlier finalize and not re-set since
            if (dt == AIGV_F32) llm_pack_weight_kernel<float><<<2048, 256, 0, st>>>((const float*)raw, gamma, dst, rows, (size_t)h->H, mul, add);
            else llm_pack_weight_kernel<__nv_bfloat16><<<2048, 256, 0, st>>>((const __nv_bfloat16*)raw, gamma, dst, rows, (size_t)h->H, mul, add);
        };
        pack(w.raw_wqkv, w.raw_dtype[0], w.attn_norm, w.wqkv, (size_t)h->QW, 1, 0);
        pack(w.raw_w1, w.raw_dtype[1], w.ffn_norm, w.w13, (size_t)h->I, 2, 0);
        pack(w.raw_w3, w.raw_dtype[2], w.ffn_norm, w.w13, (size_t)h->I, 2, 1);
        r |= make_tmap_bf16_2d(&w.tm_wqkv, w.wqkv, h->QW, h->H, h->H, bb, bc);
        r |= make_tmap_bf16_2d(&w.tm_wo, w.wo, h->H, h->H, h->H, bb, bc);
        r |= make_tmap_bf16_2d(&w.tm_w13, w.w13, 2 * h->I, h->H, h->H, bb, bc);
        r |= make_tmap_bf16_2d(&w.tm_w2, w.w2, h->H, h->I, h->I, bb, bc);
    }
    if (h->have_out) r |= make_tmap_bf16_2d(&h->tm_out_w, h->out_w, h->V, h->H, h->H, bb, bc);
    if (r) return lfail(h, AIGV_ERR_CUDA, "cuTensorMapEncodeTiled failed for a weight matrix");
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return lcuda(h, e, "weight packing");
    for (auto& w : h->layers) {   // the raw copies are no longer needed (re-set all three to re-fold after an update)
        for (void** q : {&w.raw_wqkv, &w.raw_w1, &w.raw_w3}) {
            if (*q) cudaFree(*q);
            *q = nullptr;
        }
    }
    h->finalized = true;
    return AIGV_OK;
}

#define LLM_LAUNCH(expr, what)                                           \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) return lcuda(h, _e, what);                \
        ++h->launches;                                                   \
    } while (0)

int aigv_llm_prefill(aigv_llm_handle* h, const void* inputs_embeds_bf16, int B, int N, void* stream) {
    if (!h) return AIGV_ERR_INVALID;
    if (!h->finalized) return lfail(h, AIGV_ERR_STATE, "aigv_llm_prefill before aigv_llm_finalize_weights");
    if (!inputs_embeds_bf16 || B <= 0 || N <= 0) return lfail(h, AIGV_ERR_INVALID, "aigv_llm_prefill: argument");
    const long long Mll = (long long)B * N;
    if (Mll > h->Mx) return lfail(h, AIGV_ERR_INVALID, "aigv_llm_prefill: B*N exceeds cfg.max_tokens");
    if (h->cfg.max_positions > 0 && N > h->cfg.max_positions)
        return lfail(h, AIGV_ERR_UNSUPPORTED, "sequence longer than max_position_embeddings (dynamic-NTK rescaling is not implemented)");
    if ((reinterpret_cast<uintptr_t>(inputs_embeds_bf16) & 15) != 0) return lfail(h, AIGV_ERR_INVALID, "inputs_embeds must be 16-byte aligned");
    const int M = (int)Mll;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (build_maps(h, M)) return lfail(h, AIGV_ERR_CUDA, "cuTensorMapEncodeTiled failed for an activation");
    const int H = h->H, half = h->hd / 2;
    if (h->rope_n < N) {     // rotary table for positions [0, N); grows with the longest prompt seen
        float2* t = nullptr;
        if (!lalloc(h, &t, (size_t)N * half)) return lfail(h, AIGV_ERR_CUDA, "rope table allocation failed");
        if (h->rope) { cudaStreamSynchronize(st); lfree(h, h->rope); }
        h->rope = t;
        h->rope_n = N;
        LLM_LAUNCH((llm_rope_table_kernel<<<(N * half + 255) / 256, 256, 0, st>>>(h->rope, N, half, h->cfg.rope_theta, (float)h->hd), cudaGetLastError()), "rope table");
    }
    LLM_LAUNCH((llm_seed_residual_kernel<<<(M + 7) / 8, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(inputs_embeds_bf16), h->x, h->xb, h->stats, M, H, h->parts), cudaGetLastError()),
               "seed residual");
    const float scale = 1.0f / sqrtf((float)h->hd);
    const AttnLayout lay{h->QW, h->G, (h->G + 2) * h->hd, h->G * h->hd, 0, h->hd, 1};
    for (int l = 0; l < h->L; ++l) {
        LlmLayer& w = h->layers[l];
        GemmEpilogue cons;   // RMSNorm-fused consumer: out = rstd * acc
        cons.ln_stats = h->stats; cons.ln_nparts = h->parts; cons.ln_wsum = nullptr; cons.ln_inv_d = 1.0f / (float)H; cons.ln_eps = h->cfg.rms_norm_eps;
        GemmEpilogue prod;   // fused-residual producer: x += acc; xb = bf16(x); stats
        prod.out_f32 = 1; prod.fused_residual = 1; prod.resid_in = h->x; prod.stats_out = h->stats;
        if (h->fused_rope) {   // rotary embedding of the q / k heads in the wqkv epilogue (fp32 accumulator, one rounding)
            GemmEpilogue qk = cons;
            qk.rope_tab = h->rope; qk.rope_N = N; qk.rope_group_cols = (h->G + 2) * h->hd;
            LLM_LAUNCH(launch_gemm(h->a_xb, w.tm_wqkv, h->c_qkv, M, h->QW, H, qk, h->num_sms, st), "wqkv + rope");
        } else {
            LLM_LAUNCH(launch_gemm(h->a_xb, w.tm_wqkv, h->c_qkv, M, h->QW, H, cons, h->num_sms, st), "wqkv");
            const long long n = (long long)M * h->nkv * (h->G + 1) * (half / 8);
            LLM_LAUNCH((llm_rope_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->qkv, h->rope, M, N, h->QW, h->nkv, h->G, h->hd), cudaGetLastError()), "rope");
        }
        LLM_LAUNCH(launch_attention2(h->q_map, h->kv_map, h->qkv, h->attn, B, N, h->nh, h->hd, scale, h->num_sms, st, &lay), "causal attention");
        LLM_LAUNCH(launch_gemm(h->a_attn, w.tm_wo, h->c_x, M, H, H, prod, h->num_sms, st, &h->c_xb), "wo");
        if (h->fused_swiglu) {   // silu(gate) * up in the epilogue: the [M, 2I] intermediate never reaches HBM
            GemmEpilogue sw = cons;
            sw.swiglu = 1;
            LLM_LAUNCH(launch_gemm(h->a_xb, w.tm_w13, h->c_hm, M, 2 * h->I, H, sw, h->num_sms, st), "w1|w3 + swiglu");
        } else {
            LLM_LAUNCH(launch_gemm(h->a_xb, w.tm_w13, h->c_h13, M, 2 * h->I, H, cons, h->num_sms, st), "w1|w3");
            const size_t n8 = (size_t)M * h->I / 8;
            LLM_LAUNCH((llm_silu_mul_kernel<<<(unsigned)std::min<size_t>((n8 + 255) / 256, 148 * 16), 256, 0, st>>>(h->h13, h->hm, n8), cudaGetLastError()), "swiglu");
        }
        LLM_LAUNCH(launch_gemm(h->a_hm, w.tm_w2, h->c_x, M, H, h->I, prod, h->num_sms, st, &h->c_xb), "w2");
    }
    h->cur_B = B;
    h->cur_N = N;
    return AIGV_OK;
}

static int stage_rows(aigv_llm_handle* h, const int32_t* rows_host, int n, cudaStream_t st) {
    if (!rows_host || n <= 0 || n > kMaxRows) return lfail(h, AIGV_ERR_INVALID, "rows: 1.." + std::to_string(kMaxRows) + " per call");
    if (h->cur_B == 0) return lfail(h, AIGV_ERR_STATE, "no prefill has run");
    const int M = h->cur_B * h->cur_N;
    for (int i = 0; i < n; ++i)
        if (rows_host[i] < 0 || rows_host[i] >= M) return lfail(h, AIGV_ERR_INVALID, "row index outside the last prefill");
    cudaError_t e = cudaMemcpyAsync(h->rows_dev, rows_host, (size_t)n * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return lcuda(h, e, "rows upload");
    return 0;
}

int aigv_llm_hidden_rows(aigv_llm_handle* h, const int32_t* rows_host, int n, float* out_f32, void* stream) {
    if (!h || !out_f32) return AIGV_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int r = stage_rows(h, rows_host, n, st)) return r;
    LLM_LAUNCH((llm_norm_rows_kernel<<<(n + 7) / 8, 256, 0, st>>>(h->x, h->rows_dev, h->norm_w, out_f32, nullptr, n, h->cur_B * h->cur_N, h->H, h->cfg.rms_norm_eps), cudaGetLastError()),
               "final norm");
    return AIGV_OK;
}

int aigv_llm_logits_rows(aigv_llm_handle* h, const int32_t* rows_host, int n, float* out_f32, void* stream) {
    if (!h || !out_f32) return AIGV_ERR_INVALID;
    if (!h->have_out) return lfail(h, AIGV_ERR_MISSING, "weight not set: output.weight");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int r = stage_rows(h, rows_host, n, st)) return r;
    LLM_LAUNCH((llm_norm_rows_kernel<<<(n + 7) / 8, 256, 0, st>>>(h->x, h->rows_dev, h->norm_w, nullptr, h->sel_b, n, h->cur_B * h->cur_N, h->H, h->cfg.rms_norm_eps), cudaGetLastError()),
               "final norm");
    CUtensorMap ta, tc;
    if (make_tmap_bf16_2d(&ta, h->sel_b, n, h->H, h->H, gemm_box_rows_a(), gemm_box_cols()) ||
        make_tmap_2d(&tc, h->logits, 1, n, h->Vpad, h->Vpad, gemm_box_rows_c(), gemm_box_cols_c(true)))
        return lfail(h, AIGV_ERR_CUDA, "cuTensorMapEncodeTiled failed for the output head");
    GemmEpilogue ep;
    ep.out_f32 = 1;
    LLM_LAUNCH(launch_gemm(ta, h->tm_out_w, tc, n, h->Vpad, h->H, ep, h->num_sms, st), "output head");
    cudaError_t e = cudaMemcpy2DAsync(out_f32, (size_t)h->V * 4, h->logits, (size_t)h->Vpad * 4, (size_t)h->V * 4, n, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return lcuda(h, e, "logits copy");
    return AIGV_OK;
}

int aigv_llm_copy_residual(aigv_llm_handle* h, float* dst, void* stream) {
    if (!h || !dst || h->cur_B == 0) return AIGV_ERR_INVALID;
    cudaError_t e = cudaMemcpyAsync(dst, h->x, (size_t)h->cur_B * h->cur_N * h->H * 4, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? AIGV_OK : lcuda(h, e, "copy residual");
}

/* Per-op entry points (tests). */
int aigv_attention_gqa(const void* qkv_bf16, void* out_bf16, int B, int N, int heads, int kv_heads, int head_dim, float scale,
                       int causal, void* stream) {
    if (!qkv_bf16 || !out_bf16 || B <= 0 || N <= 0 || heads <= 0 || kv_heads <= 0 || heads % kv_heads) return AIGV_ERR_INVALID;
    if (head_dim != 64 && head_dim != 128) return AIGV_ERR_UNSUPPORTED;
    const int G = heads / kv_heads;
    const uint64_t M = (uint64_t)B * N, QW = (uint64_t)(heads + 2 * kv_heads) * head_dim;
    CUtensorMap tq, tk;
    if (make_tmap_bf16_2d(&tq, qkv_bf16, M, QW, QW, attention2_q_box_rows(), 64)) return AIGV_ERR_INVALID;
    if (make_tmap_bf16_2d(&tk, qkv_bf16, M, QW, QW, attention2_kv_box_rows(head_dim), 64)) return AIGV_ERR_INVALID;
    const AttnLayout lay{(int)QW, G, (G + 2) * head_dim, G * head_dim, 0, head_dim, causal ? 1 : 0};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = launch_attention2(tq, tk, static_cast<const __nv_bfloat16*>(qkv_bf16), static_cast<__nv_bfloat16*>(out_bf16), B, N, heads,
                                      head_dim, scale, sms, static_cast<cudaStream_t>(stream), &lay);
    return e == cudaSuccess ? AIGV_OK : (e == cudaErrorInvalidValue ? AIGV_ERR_INVALID : AIGV_ERR_CUDA);
}

int aigv_rope(void* qkv_bf16, int B, int N, int heads, int kv_heads, int head_dim, float theta, void* stream) {
    if (!qkv_bf16 || B <= 0 || N <= 0 || heads <= 0 || kv_heads <= 0 || heads % kv_heads || head_dim % 16) return AIGV_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int G = heads / kv_heads, half = head_dim / 2, M = B * N, QW = (heads + 2 * kv_heads) * head_dim;
    float2* tab = nullptr;
    if (cudaMalloc(&tab, (size_t)N * half * sizeof(float2)) != cudaSuccess) return AIGV_ERR_CUDA;
    llm_rope_table_kernel<<<(N * half + 255) / 256, 256, 0, st>>>(tab, N, half, theta, (float)head_dim);
    const long long n = (long long)M * kv_heads * (G + 1) * (half / 8);
    llm_rope_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<__nv_bfloat16*>(qkv_bf16), tab, M, N, QW, kv_heads, G, head_dim);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(tab);
    return e == cudaSuccess ? AIGV_OK : AIGV_ERR_CUDA;
}

}  // extern "C"
