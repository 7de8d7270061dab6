// Internal launcher declarations shared by the .cu files and the C-ABI layer (api.cu). Not a public header:
// the public boundary is include/aigv_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace aigv {

struct GemmEpilogue {
    const float* bias = nullptr;      // [N] fp32, optional
    const float* colscale = nullptr;  // [N] fp32, optional (layer scale ls1/ls2)
    int out_f32 = 0;                  // output element type of the C tensor map: fp32 (1) or bf16 (0)
    int accumulate = 0;               // fp32 only: C += result (TMA reduce-add; the in-place residual stream)
    int gelu = 0;                     // exact erf GELU after bias
    int group_m = 0;                  // set by launch_gemm: tile order (0: N fastest; G: M fastest inside groups of G row tiles)
    int relu = 0;                     // ReLU after bias (plain epilogues only)
    int swiglu = 0;                   // bf16 only: W interleaves (gate, up) rows; C[M, N/2] = silu(acc[2i]) * acc[2i+1]
    // LayerNorm fusion, consumer side: A was the un-normalised activation (bf16) and W carries gamma;
    //   out = rstd[m] * (acc - mean[m] * ln_wsum[n]) + bias[n]  with mean/rstd from ln_stats[m] = (sum, sum of squares).
    //   ln_wsum == nullptr selects RMSNorm (no mean).
    //   ln_stats is [ln_nparts][M]: per-row partial sums over column slices, added in slice order.
    const float2* ln_stats = nullptr;
    int ln_nparts = 0;
    const float* ln_wsum = nullptr;
    float ln_inv_d = 0.f, ln_eps = 0.f;
    // LayerNorm fusion, producer side (fp32 C only): C_new = C_old + colscale * (acc + bias) is formed in the SM (C_old
    // tiles arrive by TMA), stored back as fp32 AND as bf16 through the extra map, and (sum, sum of squares) of every
    // row of C_new are written to stats_out[slice][M], one slice per gemm_stats_slice_cols() columns of C.
    int fused_residual = 0;
    const float* resid_in = nullptr;  // == the C matrix (row stride N): lets the epilogue prefetch its rows into L2
    float2* stats_out = nullptr;
    // "Delta" scheme (see gemm_tcgen05.cu, kResid == 3): delta_out = this producer leaves the fp32 master alone: C (bf16)
    // receives d = colscale * (acc + bias), the bf16 copy (extra map) is read and rewritten as bf16(xb + d), statistics of
    // that value; delta_in = a pending d (bf16 [M, N]) that this (fp32, long-K) producer adds on top of x_old.
    int delta_out = 0;
    const __nv_bfloat16* delta_in = nullptr;
    // Rotary embedding in the epilogue (bf16 output, 128-wide heads; InternLM2's wqkv): the columns are groups of rope_group_cols =
    // (G + 2) * 128 -- G query heads, one key head, one value head -- and every 128-column slot but the last of a group is rotated
    // in the rotate-half form with (cos, sin) = rope_tab[(row % rope_N) * 64 + j] (modeling_internlm2.py:247-260), in fp32 on the
    // accumulator, before the single rounding to bf16.
    const float2* rope_tab = nullptr;
    int rope_N = 0, rope_group_cols = 0;
};

// ---- TMA descriptors (host) -------------------------------------------------------------------------------------
// Row-major matrix [rows, cols] (bf16, or fp32 when f32 != 0) with row stride `ld` elements; the box is
// [box_rows, box_cols] with a 128-byte inner extent (64 bf16 / 32 fp32 columns), 128B swizzle. Returns 0 on success.
int make_tmap_2d(CUtensorMap* out, const void* base, int f32, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols);
inline int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_rows, uint32_t box_cols) {
    return make_tmap_2d(out, base, 0, rows, cols, ld, box_rows, box_cols);
}

// ---- GEMM (CTA-pair kernel, gemm_tcgen05.cu) ---------------------------------------------------------------------
// A box = [gemm_box_rows_a(), 64], W box = [gemm_box_rows_b(), 64], C box = [gemm_box_rows_c(), gemm_box_cols_c(f32)].
int gemm_box_rows_a();
int gemm_box_rows_b();
int gemm_box_cols();
int gemm_box_rows_c();
int gemm_box_cols_c(bool out_f32);
int gemm_stats_slice_cols();  // columns per statistics slice of the fused-residual epilogue
// tmX (fused_residual only): bf16 copy of C, box [gemm_box_rows_c(), 64].
cudaError_t launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, int M, int N, int K,
                        const GemmEpilogue& ep, int num_sms, cudaStream_t stream, const CUtensorMap* tmX = nullptr);

// ---- attention -------------------------------------------------------------------------------------------------
// qkv: bf16 [frames*S, 3*D] with channel order (three, head, d); out: bf16 [frames*S, D] (head-major channels).
// Persistent kernel (attention2_tcgen05.cu; the round-1 kernel it replaced is in the history, its measurements in
// profiles/r2_attention_notes.txt): two maps over the qkv matrix, boxes [attention2_q_box_rows(), 64] (Q
// tiles) and [attention2_kv_box_rows(head_dim), 64] (K / V blocks); any sequence length; head_dim 64 or 128; num_sms =
// persistent CTAs to launch.
int attention2_q_box_rows();
int attention2_kv_box_rows(int head_dim);
// Where head h's q / k / v columns sit in a row of the fused projection output, and the mask:
//   q at (h / q_group) * grp_stride + (h % q_group) * head_dim
//   k at (h / q_group) * grp_stride + k_off + (h % q_group) * k_mul,   v at k + v_off
// ViT qkv (three, head, d):            {3*H*d, H, 0, H*d, d, H*d, 0}
// InternLM2 wqkv (kv group, G q + k + v, d), G = heads / kv heads (modeling_internlm2.py:373-389), causal:
//                                       {(H + 2*Hkv)*d, G, (G+2)*d, G*d, 0, d, 1}
struct AttnLayout {
    int row_pitch, q_group, grp_stride, k_off, k_mul, v_off, causal;
};
// layout == nullptr: the ViT layout above. causal: key <= query position inside each sequence of S rows.
cudaError_t launch_attention2(const CUtensorMap& tm_q, const CUtensorMap& tm_kv, const __nv_bfloat16* qkv,
                              __nv_bfloat16* out, int frames, int S, int heads, int head_dim, float softmax_scale,
                              int num_sms, cudaStream_t stream, const AttnLayout* layout = nullptr);

// Test hook: y[i] = the attention kernel's FMA-pipe 2^x (exp2_poly2) for x[i] (softmax arguments are <= 8).
cudaError_t launch_exp2_poly_test(const float* x, float* y, int n, cudaStream_t stream);

// ---- above the seam (llm_glue.cu) ---------------------------------------------------
// Visual-token splice (eval2 :324,351-378). work: int32 [B*N + 2*B + 1] scratch; work[B*N + 2*B] is an error flag the caller
// may read back (1 = token id out of range, 2/4 = number of visual rows does not match the <IMG_CONTEXT> slots).
cudaError_t launch_splice_tokens(const long long* ids, const __nv_bfloat16* tok, const __nv_bfloat16* vit,
                                 const __nv_bfloat16* motion, __nv_bfloat16* out, int* work, int B, int N, int C, int V,
                                 int n_vit, long long ctx_id, cudaStream_t stream);
// mlpscore head (eval2 :82-94): x fp32 [n, >=4096] with row stride ldx; w[l] fp32 [out_l, in_l], b[l] fp32 [out_l].
cudaError_t launch_mlpscore(const float* x, int ldx, const float* const* w, const float* const* b, float* score, float* pre,
                            int n, cudaStream_t stream);

// Opt a kernel in to > 48 KiB of dynamic shared memory once PER DEVICE (function attributes are per device: a second
// engine on another GPU of the same process needs its own opt-in). `mask` is the call site's static bit set.
inline cudaError_t ensure_smem_optin(const void* kernel, int bytes, unsigned long long& mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && ((mask >> dev) & 1ull)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) mask |= 1ull << dev;
    return e;
}

// Output tile -> (row tile, column tile). group_m = 0: N fastest (a wave shares few A row tiles and the whole, L2-resident W);
// group_m = G > 0: M fastest inside groups of G row tiles, the groups walked over all of N one after the other (G >= num_m: plain
// M fastest -- every W tile leaves DRAM once while a small A stays in L2).
__host__ __device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int group_m, int& m_blk, int& n_blk) {
    if (group_m <= 0) {
        m_blk = tile / num_n;
        n_blk = tile - m_blk * num_n;
        return;
    }
    const int per_group = group_m * num_n;
    const int g = tile / per_group;
    const int first_m = g * group_m;
    const int gsz = num_m - first_m < group_m ? num_m - first_m : group_m;
    const int in_g = tile - g * per_group;
    n_blk = in_g / gsz;
    m_blk = first_m + (in_g - n_blk * gsz);
}

// The tile order launch_gemm picks for an [M, K] x [N, K]^T product (0: N fastest, G: groups of G row tiles, >= row tiles: M fastest).
int gemm_group_m(int M, int N, int K);

// Launch `kernel` with programmatic stream serialization (PDL): its prologue may overlap the tail of the previous kernel of the
// stream. ONLY for kernels that execute pdl_wait() (ptx.cuh) before their first global-memory access. AIGV_PDL=0 turns it off.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    static const int enabled = [] { const char* e = getenv("AIGV_PDL"); return (e && e[0] == '0') ? 0 : 1; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = enabled;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- row-wise / bandwidth kernels --------------------------------------------------------------------------------
// y[r,:] = LN(x[r,:]) * w + b (b may be null -> RMSNorm when rms != 0). x fp32 [rows, D] -> y bf16 [rows, D].
cudaError_t launch_norm_f32_to_bf16(const float* x, const float* w, const float* b, __nv_bfloat16* y, int rows, int D,
                                    float eps, int rms, cudaStream_t stream);
// In-place RMSNorm over column block [col0, col0+D) of a bf16 matrix with row stride ld (q/k norm of InternViT-6B).
// With w1: a second column block [col1, col1+D) with its own weight in the same launch (q_norm + k_norm).
cudaError_t launch_rmsnorm_bf16_inplace(__nv_bfloat16* x, const float* w, int rows, int D, int ld, int col0, float eps,
                                        cudaStream_t stream, const float* w1 = nullptr, int col1 = 0);
// pixel_values bf16 [n,3,S,S] -> patches bf16 [n*g*g, Kpad] with (c,ky,kx) flattening, zero padded to Kpad.
cudaError_t launch_im2col(const __nv_bfloat16* pv, __nv_bfloat16* patches, int n, int img, int patch, int kpad,
                          cudaStream_t stream);
// x[n, 0, :] = cls + pos[0]; x[n, 1+p, :] = pe[n*P + p, :] + pos[1+p]   (pe fp32 already has the conv bias)
// With xb/stats (LayerNorm-fused schedule): also the bf16 copy of x and per-token (sum, sum of squares) in slice 0 of
// stats[nparts][n*(P+1)] (remaining slices zeroed).
cudaError_t launch_assemble_tokens(const float* pe, const float* cls, const float* pos, float* x, int n, int P, int D,
                                   cudaStream_t stream, __nv_bfloat16* xb = nullptr, float2* stats = nullptr,
                                   int nparts = 0);
// Wf = bf16(W * gamma) per column, wsum[n] = sum_k Wf[n,k], bf[n] = b[n] + sum_k W[n,k] beta[k]  (beta, b may be null)
cudaError_t launch_fold_norm_into_linear(const __nv_bfloat16* W, const float* gamma, const float* beta, const float* b,
                                         __nv_bfloat16* Wf, float* wsum, float* bf, int N, int K, cudaStream_t stream);
// Drop cls, 2x2 pixel-shuffle (v2) and LayerNorm(4D): x fp32 [n, 1+g*g, D] -> y bf16 [n*(g/2)^2, 4D].
cudaError_t launch_pixel_shuffle_ln(const float* x, const float* w, const float* b, __nv_bfloat16* y, int n, int g,
                                    int D, float eps, int apply_ln, cudaStream_t stream);
// Pillow-exact two-pass bicubic (antialias) resize of uint8 HWC frames into square tiles, fused with
// ToTensor/Normalize (768-entry LUT) and the bf16 cast, written NCHW. Tables live on the device.
struct ResizePlan {
    const int32_t* kx;   // [out_w, ksize_x] fixed-point taps (ksize_* = table row stride, zero padded to a multiple of 4)
    const int32_t* bx;   // [out_w, 2] xmin, count
    const int32_t* ky;   // [out_h, ksize_y]
    const int32_t* by;   // [out_h, 2]
    int ksize_x, ksize_y;
    int out_w, out_h;    // resized image size (multiples of the tile size)
};
struct ResizeGeom {
    int rows_src, rows_tmp, span_words;  // shared-memory window of one 32x32 output block (see preprocess.cu)
};
}  // namespace aigv
#include <vector>
namespace aigv {
// Host: Pillow precompute_coeffs + normalize_coeffs_8bpc for one axis; returns ksize.
int plan_axis(int in_size, int out_size, std::vector<int32_t>& kk, std::vector<int32_t>& bounds);
ResizeGeom plan_geometry(const std::vector<int32_t>& bx, const std::vector<int32_t>& by, int out_w, int out_h,
                         int ksize_x, int ksize_y);
bool preprocess_fits(const ResizeGeom& g, int ksize_x, int ksize_y);  // does a block's window fit in shared memory?
cudaError_t launch_preprocess_impl(const uint8_t* frames, int n, int H, int W, int64_t frame_stride,
                                   int64_t row_stride, const ResizePlan& plan, const ResizeGeom& geo,
                                   const float* lut, int tile, int tiles_x, int tiles_per_frame, int tile_offset,
                                   uint8_t* tiles_u8, __nv_bfloat16* out, cudaStream_t stream);

}  // namespace aigv
