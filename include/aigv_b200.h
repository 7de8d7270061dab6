/*
 * aigv_b200 — C ABI of the B200-native vision hot path of AIGV-Assessor.
 *
 * The reference (pure Python) has no FFI; its seam for this path is the bound method
 *   InternVLChatModel.extract_feature(self, pixel_values) -> Tensor
 *   (/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:508-531; same method in
 *    internvl_chat/modeling_internvl_chat.py:390-409 and the four other variants)
 * and, in front of it, load_video's per-frame dynamic_preprocess + transform
 *   (/root/reference/internvl/train/internvl/eval/stage1_eval.py:443-457,484-485;
 *    /root/reference/internvl/train/dataset.py:702-738,268-274).
 * Each entry point below names the reference code it replaces. INTEGRATION.md shows the ctypes binding a
 * maintainer adds on the reference side.
 *
 * Conventions: plain C types only; every function returns 0 on success or a negative aigv_status and never throws;
 * all device work is enqueued asynchronously on the given stream (a cudaStream_t passed as void*), with no hidden
 * synchronisation or allocation after aigv_finalize_weights (the workspace is owned by the handle and sized by
 * cfg.max_tiles; larger calls are micro-batched); a handle belongs to the device that was current at aigv_create (make it
 * current for every call; several handles on several devices of one process are fine); a handle is not thread-safe.
 * Pointers are device pointers unless stated otherwise.
 */
#ifndef AIGV_B200_H_
#define AIGV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct aigv_handle aigv_handle;

enum aigv_status {
    AIGV_OK = 0,
    AIGV_ERR_INVALID = -1,     /* bad argument / shape */
    AIGV_ERR_CUDA = -2,        /* a CUDA call failed; see aigv_last_error */
    AIGV_ERR_MISSING = -3,     /* a required weight was never set */
    AIGV_ERR_STATE = -4,       /* call order (e.g. extract before finalize) */
    AIGV_ERR_UNSUPPORTED = -5  /* configuration outside what the kernels cover */
};

enum aigv_dtype { AIGV_F32 = 0, AIGV_BF16 = 1 };
enum aigv_norm { AIGV_LAYER_NORM = 0, AIGV_RMS_NORM = 1 };

/* Mirrors the InternVisionConfig fields the path reads (configuration_intern_vit.py:62-107) + the projector width. */
typedef struct aigv_cfg {
    int32_t hidden_size;        /* D */
    int32_t num_heads;
    int32_t intermediate_size;  /* I */
    int32_t num_layers;
    int32_t image_size;         /* 448 */
    int32_t patch_size;         /* 14 */
    int32_t llm_hidden;         /* C = mlp1 output width */
    int32_t qkv_bias;
    int32_t qk_normalization;
    int32_t norm_type;          /* aigv_norm */
    float layer_norm_eps;
    float mlp1_ln_eps;
    int32_t max_tiles;          /* micro-batch (tiles per launch) the workspace is sized for */
} aigv_cfg;

const char* aigv_version(void);
/* Hash of the sources this binary was built from (the Python loader refuses / rebuilds a stale library). */
const char* aigv_source_hash(void);
const char* aigv_status_string(int status);

/* Lifecycle. Replaces InternVLChatModel.__init__'s vision_model/mlp1 construction (eval2 :218-243) for this path. */
int aigv_create(const aigv_cfg* cfg, aigv_handle** out);
void aigv_destroy(aigv_handle* h);
const char* aigv_last_error(const aigv_handle* h);
size_t aigv_workspace_bytes(const aigv_handle* h, int n_tiles);

/* Weight ABI: `name` is the reference state-dict key ("vision_model.encoder.layers.3.attn.qkv.weight", "mlp1.1.bias",
 * ...). `ptr` may be a host or a device pointer (UVA); the data is copied and re-packed, the caller
 * keeps ownership. Synchronous. */
int aigv_set_weight(aigv_handle* h, const char* name, const void* ptr, int dtype, const int64_t* shape, int rank);
int aigv_finalize_weights(aigv_handle* h, void* stream);

/* extract_feature (eval2 :508-531 -> modeling_intern_vit.py:324-362): pixel_values bf16 [n_tiles,3,S,S] NCHW ->
 * out bf16 [n_tiles, (S/P/2)^2, C]. */
int aigv_extract_feature(aigv_handle* h, const void* pixel_values_bf16, int n_tiles, void* out_bf16, void* stream);
/* Debug/test: copy the fp32 residual stream [n_tiles, 1+g*g, D] (= ViT last_hidden_state) of the most recent
 * micro-batch into dst (device). */
int aigv_copy_vit_hidden(aigv_handle* h, float* dst, int n_tiles, void* stream);

/* Instrumentation (bench.py): when enabled every kernel launch of the path is bracketed by CUDA events on the launching
 * stream and accumulated per kernel class; aigv_profile_read synchronises those events. aigv_launch_count = number of
 * this library's kernels launched through the handle so far. */
int aigv_profile_enable(aigv_handle* h, int on);
int aigv_profile_num_classes(void);
const char* aigv_profile_class_name(int cls);
int aigv_profile_read(aigv_handle* h, double* ms, int64_t* counts, int n, int reset);
int64_t aigv_launch_count(const aigv_handle* h);

/* Front end. aigv_tile_grid = find_closest_aspect_ratio / grid choice of dynamic_preprocess (dataset.py:685-713).
 * aigv_preprocess_u8 = dynamic_preprocess resize+crop(+thumbnail) (dataset.py:715-738) fused with build_transform
 * (dataset.py:268-274) and the bf16 cast. frames: uint8 HWC RGB [n, H, W, 3] with the given byte strides. Writes
 * n * (gi*gj + (thumb && gi*gj>1)) tiles, frame-major, tiles row-major then thumbnail. tiles_u8 (optional, may be NULL)
 * receives the uint8 HWC tiles before normalisation. The first call for a new (H, W, grid) builds and uploads the tap
 * tables synchronously; every later call is asynchronous on `stream`. Sources whose down-scale window does not fit in
 * shared memory (beyond about 8x, i.e. larger than 4K) return AIGV_ERR_UNSUPPORTED. */
int aigv_tile_grid(int width, int height, int min_num, int max_num, int image_size, int* gi, int* gj);
int aigv_preprocess_u8(aigv_handle* h, const uint8_t* frames, int n, int H, int W, int64_t frame_stride,
                       int64_t row_stride, int gi, int gj, int thumb, uint8_t* tiles_u8, void* out_bf16, void* stream);

/* Per-op entry points (unit parity tests call these; the fused path above uses the same launchers). */
/* out = residual + colscale * act(A[M,K] @ W[N,K]^T + bias); exactly one of out_f32/out_bf16 non-NULL. */
int aigv_gemm_epilogue(const void* a_bf16, const void* w_bf16, const float* bias, const float* colscale,
                       const float* residual, float* out_f32, void* out_bf16, int M, int N, int K, int gelu,
                       void* stream);
/* LayerNorm-fused schedule (DESIGN.md "LayerNorm fusion"): the norm in front of qkv / fc1 (modeling_intern_vit.py:
 * 224-226 norm1/norm2) is never materialised. Producers (proj, fc2 and the token assembly) leave the fp32 residual
 * stream, its bf16 copy and per-token (sum, sum of squares) slices; consumers apply mean/rstd in their epilogue with
 * gamma/beta folded into the weights:  LN(x) W^T + b = rstd * (x Wf^T - mean * wsum) + bias_f.
 *   aigv_stats_slices(N)     number of statistics slices a producer with N output columns writes
 *   aigv_fold_norm           Wf = bf16(W * gamma), wsum[n] = sum_k Wf[n,k], bias_f = bias + W beta (beta/bias may be NULL)
 *   aigv_gemm_residual_stats x += colscale * (A W^T + bias) in place; xb = bf16(x); stats[slices][M][2]
 *   aigv_gemm_norm_fused     out = act(rstd * (A Wf^T - mean * wsum) + bias_f); wsum NULL = RMSNorm; D = normalised width
 *   aigv_assemble_tokens_stats  aigv_assemble_tokens + xb + stats (slice 0 carries the sums, the others are zeroed) */
int aigv_stats_slices(int N);
int aigv_fold_norm(const void* w_bf16, const float* gamma, const float* beta, const float* bias, void* wf_bf16,
                   float* wsum, float* bias_f, int N, int K, void* stream);
int aigv_gemm_residual_stats(const void* a_bf16, const void* w_bf16, const float* bias, const float* colscale,
                             float* x_f32, void* xb_bf16, float* stats, int M, int N, int K, void* stream);
int aigv_gemm_norm_fused(const void* a_bf16, const void* wf_bf16, const float* wsum, const float* bias_f,
                         const float* stats, int nparts, int D, float eps, void* out_bf16, int M, int N, int K, int gelu,
                         void* stream);
int aigv_assemble_tokens_stats(const float* pe, const float* cls, const float* pos, float* x, void* xb_bf16,
                               float* stats, int nparts, int n, int P, int D, void* stream);
/* qkv bf16 [frames*S, 3*heads*head_dim] (three, head, d) -> out bf16 [frames*S, heads*head_dim]
 * (flash_attention.py:52-55 / modeling_intern_vit.py:143-157). Any sequence length; head_dim 64 or 128. */
int aigv_attention(const void* qkv_bf16, void* out_bf16, int frames, int S, int heads, int head_dim, float scale,
                   void* stream);
/* Test hook (host only, no GPU work): the order in which the persistent GEMM walks the 256 x 256 output tiles of an [M, K] x [N, K]^T
 * product. *group_m: 0 = N fastest, G = M fastest inside groups of G row tiles (>= the number of row tiles: plain M fastest). With
 * m_blk / n_blk (capacity >= tiles): the (row tile, column tile) of tile 0, 1, ... exactly as the kernel computes them. */
int aigv_gemm_tile_order(int M, int N, int K, int* group_m, int32_t* m_blk, int32_t* n_blk, int capacity);
/* Test hook: y[i] = 2^x[i] exactly as the attention kernel evaluates the share of its exponentials that runs on the
 * FMA pipe instead of MUFU (Cody-Waite split + cubic; x fp32 [n], device pointers). */
int aigv_exp2_poly(const float* x, float* y, int n, void* stream);
/* Above the seam (first pieces). aigv_splice_tokens = InternVLChatModel.forward's
 * input_embeds = tok_embeddings(input_ids) with every <IMG_CONTEXT> slot but the last of a sample overwritten by the next
 * visual token (flattened sample-major order, like the boolean-mask assignment) and the last one by the sample's motion
 * token (modeling_internvl_chat.py eval2 :324,351-378). work: int32 scratch of B*N + 2*B + 1 elements; its last element is
 * an error flag (0 = ok; 1: a token id outside [0,V); 2/4: n_vit does not match the slots -- the case where the reference
 * falls into its except branch). aigv_mlpscore = the MLP head of eval2 :82-94 (ReLU after every layer incl. the last) on
 * x[n, 4096] (row stride ldx), weights[l] fp32 [out_l, in_l] for 4096-1024-256-64-16-1; pre_relu (optional) receives the
 * last layer's value before its ReLU. */
int aigv_splice_tokens(const int64_t* input_ids, const void* tok_embeddings_bf16, const void* vit_embeds_bf16,
                       const void* motion_embeds_bf16, void* out_bf16, int32_t* work, int B, int N, int C, int V, int n_vit,
                       int64_t img_context_token_id, void* stream);
int aigv_mlpscore(const float* x, int ldx, const float* const* weights, const float* const* biases, float* score,
                  float* pre_relu, int n, void* stream);
int aigv_layernorm(const float* x, const float* w, const float* b, void* y_bf16, int rows, int D, float eps, int rms,
                   void* stream);
int aigv_rmsnorm_inplace(void* x_bf16, const float* w, int rows, int D, int ld, int col0, float eps, void* stream);
int aigv_im2col(const void* pixel_values_bf16, void* patches_bf16, int n, int image_size, int patch, int kpad,
                void* stream);
int aigv_assemble_tokens(const float* pe, const float* cls, const float* pos, float* x, int n, int P, int D,
                         void* stream);
int aigv_pixel_shuffle(const float* x, const float* w, const float* b, void* y_bf16, int n, int g, int D, float eps,
                       int apply_ln, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Above the seam, second piece: the InternLM2 prefill of the stage-2 forward
 *   outputs = self.language_model(inputs_embeds=input_embeds, ..., output_hidden_states=True)
 *   (/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:441-450 ->
 *    /root/reference/internvl/model/internlm2/modeling_internlm2.py:1036-1124 -> :868-998 -> :635-695)
 * of which the caller consumes hidden_states[-1][:, -4] (mlpscore, eval2 :469-475) and the argmax of the answer-span
 * logits (eval2 :488, stage1_eval.py:928-929). One handle = one InternLM2 decoder stack on the current device.
 * Scope: a full prefill of B unpadded prompts of N tokens each (positions 0..N-1, causal mask), bf16 with fp32 accumulation
 * and an fp32 residual stream; no KV cache / decode, no padding mask, N <= max_positions (LLM:218-244 dynamic NTK leaves the
 * rotary base unchanged there). */
typedef struct aigv_llm_handle aigv_llm_handle;

/* Mirrors the InternLM2Config fields the prefill reads (configuration_internlm2.py). */
typedef struct aigv_llm_cfg {
    int32_t hidden_size;
    int32_t num_heads;
    int32_t num_kv_heads;
    int32_t intermediate_size;
    int32_t num_layers;
    int32_t vocab_size;
    float rms_norm_eps;
    float rope_theta;
    int32_t max_tokens;     /* workspace: tokens (B*N) per aigv_llm_prefill call */
    int32_t max_positions;  /* max_position_embeddings; longer prompts are refused (0 = do not check) */
} aigv_llm_cfg;

int aigv_llm_create(const aigv_llm_cfg* cfg, aigv_llm_handle** out);
void aigv_llm_destroy(aigv_llm_handle* h);
const char* aigv_llm_last_error(const aigv_llm_handle* h);
int64_t aigv_llm_launch_count(const aigv_llm_handle* h);
/* `name` is the reference state-dict key, with or without the "language_model." prefix:
 * model.layers.{i}.attention.wqkv.weight / .attention.wo.weight / .attention_norm.weight / .ffn_norm.weight /
 * .feed_forward.w1.weight / .w2.weight / .w3.weight, model.norm.weight, output.weight (optional: only aigv_llm_logits_rows
 * needs it). Host or device pointer, fp32 or bf16; copied and re-packed (RMSNorm weights folded into wqkv / w1 / w3, w1 and
 * w3 interleaved) at aigv_llm_finalize_weights, after which the raw copies are dropped: to change a norm weight later, re-set
 * the matrices it feeds as well and finalize again. Synchronous. */
int aigv_llm_set_weight(aigv_llm_handle* h, const char* name, const void* ptr, int dtype, const int64_t* shape, int rank);
int aigv_llm_finalize_weights(aigv_llm_handle* h, void* stream);
/* The decoder stack on inputs_embeds bf16 [B*N, hidden] (sample-major; = the output of aigv_splice_tokens). The residual
 * stream of the last layer stays in the handle for the two readers below. Asynchronous on `stream`. */
int aigv_llm_prefill(aigv_llm_handle* h, const void* inputs_embeds_bf16, int B, int N, void* stream);
/* hidden_states[-1] (= final RMSNorm applied, modeling_internlm2.py:984) of the given flattened rows b*N + pos of the last
 * prefill -> out fp32 [n, hidden]. rows is a HOST array (n <= 512). */
int aigv_llm_hidden_rows(aigv_llm_handle* h, const int32_t* rows_host, int n, float* out_f32, void* stream);
/* logits (output head, modeling_internlm2.py:1095) of the given rows only -> out fp32 [n, vocab_size]. */
int aigv_llm_logits_rows(aigv_llm_handle* h, const int32_t* rows_host, int n, float* out_f32, void* stream);
/* Debug/test: the fp32 residual stream [B*N, hidden] after the last layer, before the final norm. */
int aigv_llm_copy_residual(aigv_llm_handle* h, float* dst, void* stream);
/* Per-op entry points of the prefill (tests). qkv bf16 [B*N, (heads + 2*kv_heads)*head_dim] in InternLM2's wqkv packing
 * (kv group, G query heads + k + v, d) (modeling_internlm2.py:373-389). aigv_rope rotates its q and k columns in place
 * (:247-260, positions 0..N-1; synchronous); aigv_attention_gqa -> out bf16 [B*N, heads*head_dim] with K / V shared per
 * group (:404-405) and, if causal, key <= query position (:411-416 mask). */
int aigv_rope(void* qkv_bf16, int B, int N, int heads, int kv_heads, int head_dim, float theta, void* stream);
int aigv_attention_gqa(const void* qkv_bf16, void* out_bf16, int B, int N, int heads, int kv_heads, int head_dim,
                       float scale, int causal, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Motion branch: per-op entry points from which aigv_assessor_b200/slowfast.py assembles the
 * SlowFast-R50 trunk the reference calls (eval2 modeling_internvl_chat.py:136-193 -> pytorchvideo slowfast_r50 blocks 0-4 +
 * pools). Activations are channels-last bf16 [n, t, h, w, c]; BatchNorm is folded into weights / bias by the host.
 *   aigv_im2col3d     Conv3d's patch gather: out[pos, (dt, dh, dw, c)] (row stride kpad, zero padded); element strides
 *                     of the input are explicit, so it also reads the NCHW pixel_values for the stems
 *   aigv_gemm_ex      out[M, N] (row stride ldc) = act(A[M, K] (row stride lda) W[N, K]^T + bias). W / bias hold n_pad >= N
 *                     rows (zero padded by the caller): with n_pad % 32 == 0, K % 8 == 0, N % 8 == 0 and 16-byte alignment the
 *                     tcgen05 kernel computes n_pad columns and its TMA store is clipped to the N real ones; otherwise a
 *                     CUDA-core kernel runs
 *   aigv_maxpool3d    MaxPool3d((1,kh,kw), (1,sh,sw), (0,ph,pw)) of the stems (n_t = n * t frames), output row stride ldo
 *   aigv_add_relu     residual join: out[r, :C] (row stride ldo) = relu(a[r, :] + b[r, :])
 *   aigv_slowfast_tail  repeat_interleave(rep) along t + AvgPool3d((kt,kh,kw), stride 1) + AdaptiveAvgPool3d(1)
 *                     (eval2 :183-189) -> out fp32 [n, out_ld] columns [out_col, out_col + C) */
int aigv_im2col3d(const void* in_bf16, void* out_bf16, int n, int T, int H, int W, int C, int64_t s_n, int64_t s_t,
                  int64_t s_h, int64_t s_w, int64_t s_c, int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph,
                  int pw, int kpad, void* stream);
int aigv_gemm_ex(const void* a_bf16, int lda, const void* w_bf16, const float* bias, void* out_bf16, int ldc, int M, int N,
                 int n_pad, int K, int relu, void* stream);
int aigv_maxpool3d(const void* in_bf16, void* out_bf16, int n_t, int H, int W, int C, int kh, int kw, int sh, int sw, int ph,
                   int pw, int ldo, void* stream);
int aigv_add_relu(const void* a_bf16, const void* b_bf16, void* out_bf16, int64_t rows, int C, int ldo, void* stream);
int aigv_slowfast_tail(const void* x_bf16, float* out_f32, int n, int T, int H, int W, int C, int ld, int rep, int kt, int kh,
                       int kw, int out_ld, int out_col, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AIGV_B200_H_ */
