// Motion branch: the plumbing kernels that turn the SlowFast-R50 trunk the reference calls
// (/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:136-193: pytorchvideo's slowfast_r50, its
// first five blocks, the two average pools and the adaptive pool) into GEMMs on this library's tcgen05 kernel:
//   * activations live channels-last, [n, t, h, w, c] bf16 rows of c, so every 1x1x1 convolution IS a GEMM on the activation
//     matrix and every other convolution is an im2col gather (aigv_im2col3d) followed by a GEMM; BatchNorm (eval mode) is
//     folded into the weights and bias on the host, ReLU runs in the GEMM epilogue;
//   * aigv_gemm_ex = the tcgen05 GEMM with leading dimensions (outputs land in a column range of a wider matrix: the
//     fast-to-slow fusion's channel concat costs nothing) when the shape allows (N % 32 == 0, K % 8 == 0), else a plain
//     CUDA-core kernel (the fast pathway's 8- and 16-channel layers: ~5 % of the trunk's FLOPs);
//   * aigv_maxpool3d (the stem pools), aigv_add_relu (residual joins), aigv_slowfast_tail (repeat_interleave(4) along t +
//     AvgPool3d((kt,7,7)) + AdaptiveAvgPool3d(1), :183-189, as one weighted reduction).
// The network itself (which layers, in which order) is host logic: aigv_assessor_b200/slowfast.py.
#include <math.h>

#include "../../include/aigv_b200.h"
#include "kernels.h"
#include "ptx.cuh"

namespace aigv {
namespace {

// out[pos, (dt, dh, dw, c)] = in[n, t*st + dt - pt, h*sh + dh - ph, w*sw + dw - pw, c]  (0 outside), rows padded to kpad.
// Input strides are in elements, so the same kernel reads channels-last activations and the NCHW pixel_values of the stem.
struct Im2colArgs {
    int n, T, H, W, C;
    long long s_n, s_t, s_h, s_w, s_c;
    int kt, kh, kw, st, sh, sw, pt, ph, pw;
    int To, Ho, Wo, kpad;
};

__global__ void __launch_bounds__(256)
im2col3d_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, Im2colArgs a) {
    // one thread = 8 consecutive k of one output row = one 16-byte store; channels-last inputs with C % 8 == 0 fetch them
    // with one 16-byte load, anything else (the stems' 3-channel NCHW frames) element by element
    const int K = a.kt * a.kh * a.kw * a.C;
    const long long rows = static_cast<long long>(a.n) * a.To * a.Ho * a.Wo;
    const bool vec = a.s_c == 1 && (a.C % 8) == 0;
    const int kunits = a.kpad / 8;
    const long long total = rows * kunits;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / kunits;
        const int k0 = static_cast<int>(i - row * kunits) * 8;
        const int wo = static_cast<int>(row % a.Wo);
        long long q = row / a.Wo;
        const int ho = static_cast<int>(q % a.Ho);
        q /= a.Ho;
        const int to = static_cast<int>(q % a.To);
        const long long base = (q / a.To) * a.s_n;
        const int t0 = to * a.st - a.pt, h0 = ho * a.sh - a.ph, w0 = wo * a.sw - a.pw;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        int c = k0 % a.C, tap = k0 / a.C;
        int dw = tap % a.kw, dh = (tap / a.kw) % a.kh, dt = tap / (a.kw * a.kh);
        if (vec) {
            const int t = t0 + dt, hh = h0 + dh, ww = w0 + dw;
            if (k0 < K && t >= 0 && t < a.T && hh >= 0 && hh < a.H && ww >= 0 && ww < a.W)
                v = __ldg(reinterpret_cast<const uint4*>(in + base + t * a.s_t + hh * a.s_h + ww * a.s_w + c));
        } else {
            uint32_t e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = t0 + dt, hh = h0 + dh, ww = w0 + dw;
                uint32_t bits = 0;
                if (k0 + j < K && t >= 0 && t < a.T && hh >= 0 && hh < a.H && ww >= 0 && ww < a.W)
                    bits = __bfloat16_as_ushort(in[base + t * a.s_t + hh * a.s_h + ww * a.s_w + c * a.s_c]);
                e[j] = bits;
                if (++c == a.C) {   // next tap
                    c = 0;
                    if (++dw == a.kw) { dw = 0; if (++dh == a.kh) { dh = 0; ++dt; } }
                }
            }
            v = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
        }
        *reinterpret_cast<uint4*>(out + row * a.kpad + k0) = v;
    }
}

// MaxPool3d over (1, kh, kw) windows, stride (1, sh, sw), padding (0, ph, pw) (padding never wins), channels-last.
__global__ void __launch_bounds__(256)
maxpool_hw_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int nt, int H, int W, int C, int kh,
                  int kw, int sh, int sw, int ph, int pw, int Ho, int Wo, int ldo) {
    const long long total = static_cast<long long>(nt) * Ho * Wo * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long long q = i / C;
        const int wo = static_cast<int>(q % Wo);
        q /= Wo;
        const int ho = static_cast<int>(q % Ho);
        const long long f = q / Ho;
        float m = -INFINITY;
        for (int dh = 0; dh < kh; ++dh)
            for (int dw = 0; dw < kw; ++dw) {
                const int hh = ho * sh + dh - ph, ww = wo * sw + dw - pw;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                    m = fmaxf(m, __bfloat162float(in[((f * H + hh) * W + ww) * C + c]));
            }
        out[(i / C) * ldo + c] = __float2bfloat16_rn(m);
    }
}

// out[r, :C] = relu(a[r, :] + b[r, :]), bf16, 8 elements per thread; a, b contiguous [rows, C], out with row stride ldo
__global__ void __launch_bounds__(256)
add_relu_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out, size_t n8,
                int c8, int ldo8) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(a) + i), y = __ldg(reinterpret_cast<const uint4*>(b) + i);
        const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float l = fmaxf(__uint_as_float(xw[e] << 16) + __uint_as_float(yw[e] << 16), 0.f);
            const float h = fmaxf(__uint_as_float(xw[e] & 0xffff0000u) + __uint_as_float(yw[e] & 0xffff0000u), 0.f);
            o[e] = pack_bf16x2(l, h);
        }
        reinterpret_cast<uint4*>(out)[(i / c8) * ldo8 + (i % c8)] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// repeat_interleave(rep) along t, AvgPool3d((kt, kh, kw), stride 1, no padding), then the mean of the pooled grid
// (AdaptiveAvgPool3d(1)): every input position enters with weight (#windows covering it) / (#windows * window size).
// One block per (sample, 32-channel group); x channels-last [n, T, H, W, C] with row stride ld; out fp32 [n, out_ld] at out_col.
__global__ void __launch_bounds__(256)
slowfast_tail_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int T, int H, int W, int C, int ld, int rep,
                     int kt, int kh, int kw, int out_ld, int out_col) {
    __shared__ float part[8][32];
    const int n = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), grp = threadIdx.x >> 5;
    const int Tr = T * rep, To = Tr - kt + 1, Ho = H - kh + 1, Wo = W - kw + 1;
    float acc = 0.f;
    if (c < C && To > 0 && Ho > 0 && Wo > 0) {
        for (int p = grp; p < T * H * W; p += 8) {
            const int w = p % W, h = (p / W) % H, t = p / (W * H);
            // windows (start index s in [0, n_out)) that cover index i: max(0, i - k + 1) <= s <= min(i, n_out - 1)
            const int cw = min(w, Wo - 1) - max(0, w - kw + 1) + 1;
            const int ch = min(h, Ho - 1) - max(0, h - kh + 1) + 1;
            int ct = 0;
            for (int r = 0; r < rep; ++r) {
                const int tr = t * rep + r;
                ct += max(0, min(tr, To - 1) - max(0, tr - kt + 1) + 1);
            }
            const float wgt = static_cast<float>(cw) * ch * ct;
            acc = fmaf(wgt, __bfloat162float(x[(static_cast<size_t>(n) * T * H * W + p) * ld + c]), acc);
        }
    }
    part[grp][threadIdx.x & 31] = acc;
    __syncthreads();
    if (grp == 0 && c < C) {
        float s = 0.f;
        for (int g = 0; g < 8; ++g) s += part[g][threadIdx.x];
        const float norm = static_cast<float>(To) * Ho * Wo * kt * kh * kw;
        out[static_cast<size_t>(n) * out_ld + out_col + c] = s / norm;
    }
}

// Plain CUDA-core GEMM for the shapes the tensor-core kernel does not take (N % 32 != 0, tiny K):
// out[m, n] = act(sum_k A[m, k] W[n, k] + bias[n]); 32 x 32 output tile per block, 32-wide k slabs in shared memory.
__global__ void __launch_bounds__(256)
gemm_small_kernel(const __nv_bfloat16* __restrict__ A, int lda, const __nv_bfloat16* __restrict__ Wt, const float* __restrict__ bias,
                  __nv_bfloat16* __restrict__ out, int ldc, int M, int N, int K, int relu) {
    __shared__ float sa[32][33], sb[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 row groups x 32 columns
    const int m0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += 32) {
        for (int r = ty; r < 32; r += 8) {
            const int k = k0 + tx;
            sa[r][tx] = (m0 + r < M && k < K) ? __bfloat162float(A[static_cast<size_t>(m0 + r) * lda + k]) : 0.f;
            sb[r][tx] = (n0 + r < N && k < K) ? __bfloat162float(Wt[static_cast<size_t>(n0 + r) * K + k]) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            const float b = sb[tx][kk];
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = fmaf(sa[ty + 8 * e][kk], b, acc[e]);
        }
        __syncthreads();
    }
    const int n = n0 + tx;
    if (n < N) {
        const float bv = bias ? bias[n] : 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int m = m0 + ty + 8 * e;
            if (m < M) {
                float v = acc[e] + bv;
                if (relu) v = fmaxf(v, 0.f);
                out[static_cast<size_t>(m) * ldc + n] = __float2bfloat16_rn(v);
            }
        }
    }
}

int grid_for(long long total) {
    long long b = (total + 255) / 256;
    return static_cast<int>(b < 1 ? 1 : (b > 148 * 32 ? 148 * 32 : b));
}

}  // namespace
}  // namespace aigv

using namespace aigv;

extern "C" {

int aigv_im2col3d(const void* in_bf16, void* out_bf16, int n, int T, int H, int W, int C, int64_t s_n, int64_t s_t, int64_t s_h,
                  int64_t s_w, int64_t s_c, int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw, int kpad,
                  void* stream) {
    if (!in_bf16 || !out_bf16 || n <= 0 || T <= 0 || H <= 0 || W <= 0 || C <= 0 || kt <= 0 || kh <= 0 || kw <= 0 || st <= 0 ||
        sh <= 0 || sw <= 0 || pt < 0 || ph < 0 || pw < 0)
        return AIGV_ERR_INVALID;
    Im2colArgs a{n, T, H, W, C, s_n, s_t, s_h, s_w, s_c, kt, kh, kw, st, sh, sw, pt, ph, pw, 0, 0, 0, kpad};
    a.To = (T + 2 * pt - kt) / st + 1;
    a.Ho = (H + 2 * ph - kh) / sh + 1;
    a.Wo = (W + 2 * pw - kw) / sw + 1;
    if (a.To <= 0 || a.Ho <= 0 || a.Wo <= 0 || kpad < kt * kh * kw * C || (kpad % 8) != 0 || (reinterpret_cast<uintptr_t>(out_bf16) & 15))
        return AIGV_ERR_INVALID;
    const bool vec = s_c == 1 && (C % 8) == 0;
    if (vec && ((s_n | s_t | s_h | s_w) % 8 != 0 || (reinterpret_cast<uintptr_t>(in_bf16) & 15))) return AIGV_ERR_INVALID;
    const long long total = static_cast<long long>(n) * a.To * a.Ho * a.Wo * (kpad / 8);
    im2col3d_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(in_bf16),
                                                                                      static_cast<__nv_bfloat16*>(out_bf16), a);
    return cudaGetLastError() == cudaSuccess ? AIGV_OK : AIGV_ERR_CUDA;
}

int aigv_maxpool3d(const void* in_bf16, void* out_bf16, int n_t, int H, int W, int C, int kh, int kw, int sh, int sw, int ph, int pw,
                   int ldo, void* stream) {
    if (!in_bf16 || !out_bf16 || n_t <= 0 || H <= 0 || W <= 0 || C <= 0 || kh <= 0 || kw <= 0 || sh <= 0 || sw <= 0 || ldo < C)
        return AIGV_ERR_INVALID;
    const int Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;
    if (Ho <= 0 || Wo <= 0) return AIGV_ERR_INVALID;
    const long long total = static_cast<long long>(n_t) * Ho * Wo * C;
    maxpool_hw_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in_bf16), static_cast<__nv_bfloat16*>(out_bf16), n_t, H, W, C, kh, kw, sh, sw, ph, pw, Ho, Wo, ldo);
    return cudaGetLastError() == cudaSuccess ? AIGV_OK : AIGV_ERR_CUDA;
}

int aigv_add_relu(const void* a_bf16, const void* b_bf16, void* out_bf16, int64_t rows, int C, int ldo, void* stream) {
    if (!a_bf16 || !b_bf16 || !out_bf16 || rows <= 0 || C <= 0 || (C % 8) != 0 || (ldo % 8) != 0 || ldo < C) return AIGV_ERR_INVALID;
    const int64_t n = rows * C;
    add_relu_kernel<<<grid_for(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(a_bf16), static_cast<const __nv_bfloat16*>(b_bf16), static_cast<__nv_bfloat16*>(out_bf16),
        static_cast<size_t>(n / 8), C / 8, ldo / 8);
    return cudaGetLastError() == cudaSuccess ? AIGV_OK : AIGV_ERR_CUDA;
}

int aigv_slowfast_tail(const void* x_bf16, float* out_f32, int n, int T, int H, int W, int C, int ld, int rep, int kt, int kh, int kw,
                       int out_ld, int out_col, void* stream) {
    if (!x_bf16 || !out_f32 || n <= 0 || T <= 0 || H < kh || W < kw || T * rep < kt || C <= 0 || ld < C || rep <= 0) return AIGV_ERR_INVALID;
    dim3 grid((C + 31) / 32, n);
    slowfast_tail_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x_bf16), out_f32, T, H, W, C,
                                                                               ld, rep, kt, kh, kw, out_ld, out_col);
    return cudaGetLastError() == cudaSuccess ? AIGV_OK : AIGV_ERR_CUDA;
}

int aigv_gemm_ex(const void* a_bf16, int lda, const void* w_bf16, const float* bias, void* out_bf16, int ldc, int M, int N, int n_pad,
                 int K, int relu, void* stream) {
    if (!a_bf16 || !w_bf16 || !out_bf16 || M <= 0 || N <= 0 || K <= 0 || lda < K || ldc < N || n_pad < N) return AIGV_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // n_pad = rows W (and entries bias) really has, zero padded by the caller: lets a narrow layer (N = 8, 16) run on the
    // tensor cores as an n_pad-wide GEMM whose TMA store is clipped to the N real columns
    const bool tc_ok = (n_pad % 32) == 0 && (K % 8) == 0 && (lda % 8) == 0 && (ldc % 8) == 0 && (N % 8) == 0 &&
                       (reinterpret_cast<uintptr_t>(a_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_bf16) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0;
    if (tc_ok) {
        CUtensorMap ta, tb, tc;
        if (make_tmap_bf16_2d(&ta, a_bf16, M, K, lda, gemm_box_rows_a(), gemm_box_cols()) ||
            make_tmap_bf16_2d(&tb, w_bf16, n_pad, K, K, gemm_box_rows_b(), gemm_box_cols()) ||
            make_tmap_2d(&tc, out_bf16, 0, M, N, ldc, gemm_box_rows_c(), gemm_box_cols_c(false)))
            return AIGV_ERR_INVALID;
        GemmEpilogue ep;
        ep.bias = bias;
        ep.relu = relu;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = launch_gemm(ta, tb, tc, M, n_pad, K, ep, sms, st);
        return e == cudaSuccess ? AIGV_OK : (e == cudaErrorInvalidValue ? AIGV_ERR_INVALID : AIGV_ERR_CUDA);
    }
    dim3 grid((M + 31) / 32, (N + 31) / 32);
    gemm_small_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(a_bf16), lda, static_cast<const __nv_bfloat16*>(w_bf16), bias,
                                            static_cast<__nv_bfloat16*>(out_bf16), ldc, M, N, K, relu);
    return cudaGetLastError() == cudaSuccess ? AIGV_OK : AIGV_ERR_CUDA;
}

}  // extern "C"
