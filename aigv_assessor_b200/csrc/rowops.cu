// Row-wise / bandwidth-bound kernels of the vision path (HBM roofline; vectorised 16-byte accesses, one warp per row).
//   norm_f32_to_bf16      nn.LayerNorm(D, eps) / InternRMSNorm  (modeling_intern_vit.py:208-209,32-43)
//   rmsnorm_bf16_inplace  q_norm / k_norm of InternViT-6B       (modeling_intern_vit.py:135-137,166-170)
//   im2col                patch extraction for the k=s=14 conv  (modeling_intern_vit.py:97)
//   assemble_tokens       cls concat + position embedding       (modeling_intern_vit.py:99-106)
//   pixel_shuffle_ln      drop cls + pixel_shuffle v2 + mlp1[0] LayerNorm (eval2 modeling_internvl_chat.py:492-506,520-529,239)
#include "kernels.h"
#include "ptx.cuh"

namespace aigv {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Statistics over `nseg` segments of `seg_len` floats each (segment s starts at seg_ptr(s)); two-pass variance in fp32.
struct RowStats { float mean, rstd; };

template <typename SegPtr>
__device__ __forceinline__ RowStats row_stats(SegPtr seg_ptr, int nseg, int seg_len, float eps, bool rms, int lane) {
    const int n4 = seg_len >> 2;
    const float inv_n = 1.0f / static_cast<float>(nseg * seg_len);
    float mean = 0.f;
    if (!rms) {
        float s = 0.f;
        for (int g = 0; g < nseg; ++g) {
            const float4* p = reinterpret_cast<const float4*>(seg_ptr(g));
            for (int i = lane; i < n4; i += 32) {
                const float4 t = p[i];
                s += (t.x + t.y) + (t.z + t.w);
            }
        }
        mean = warp_sum(s) * inv_n;
    }
    float q = 0.f;
    for (int g = 0; g < nseg; ++g) {
        const float4* p = reinterpret_cast<const float4*>(seg_ptr(g));
        for (int i = lane; i < n4; i += 32) {
            const float4 t = p[i];
            const float a = t.x - mean, b = t.y - mean, c = t.z - mean, d = t.w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float var = warp_sum(q) * inv_n;
    return {mean, rsqrtf(var + eps)};
}

// One warp per row, the whole row held in registers (VPL float4 per lane, D = 128 * VPL): a single global read per
// element, all loads in flight at once, exact two-pass mean / variance in fp32, 16-byte bf16 stores.
template <int VPL>
__global__ void __launch_bounds__(256)
norm_f32_to_bf16_reg_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                            __nv_bfloat16* __restrict__ y, int rows, float eps, int rms) {
    constexpr int D = 128 * VPL;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
    float4 v[VPL];
    // lane owns float4 index (2*i)*32 + 2*lane + {0,1}: pairs of adjacent float4 -> one 16-byte bf16 store per pair
    if constexpr (VPL == 1) {
        v[0] = __ldcs(xr + lane);
    } else {
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[i] = __ldcs(xr + (i >> 1) * 64 + 2 * lane + (i & 1));  // streaming: read once
    }
    float mean = 0.f;
    if (!rms) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        mean = warp_sum(s) * (1.0f / D);
    }
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    __nv_bfloat16* yr = y + static_cast<size_t>(row) * D;
    auto nrm = [&](float xv, float wv, float bv) { return (xv - mean) * rstd * wv + bv; };
    if constexpr (VPL >= 2) {
#pragma unroll
        for (int i = 0; i < VPL; i += 2) {
            const int col = ((i >> 1) * 64 + 2 * lane) * 4;  // first of 8 consecutive columns
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + col));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + col + 4));
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (b) {
                b0 = __ldg(reinterpret_cast<const float4*>(b + col));
                b1 = __ldg(reinterpret_cast<const float4*>(b + col + 4));
            }
            uint4 o;
            o.x = pack_bf16x2(nrm(v[i].x, w0.x, b0.x), nrm(v[i].y, w0.y, b0.y));
            o.y = pack_bf16x2(nrm(v[i].z, w0.z, b0.z), nrm(v[i].w, w0.w, b0.w));
            o.z = pack_bf16x2(nrm(v[i + 1].x, w1.x, b1.x), nrm(v[i + 1].y, w1.y, b1.y));
            o.w = pack_bf16x2(nrm(v[i + 1].z, w1.z, b1.z), nrm(v[i + 1].w, w1.w, b1.w));
            *reinterpret_cast<uint4*>(yr + col) = o;
        }
    } else {  // D = 128: one float4 per lane (the pair mapping degenerates to xr[lane] handled below)
        const int col = lane * 4;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + col));
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b) b0 = __ldg(reinterpret_cast<const float4*>(b + col));
        uint2 o;
        o.x = pack_bf16x2(nrm(v[0].x, w0.x, b0.x), nrm(v[0].y, w0.y, b0.y));
        o.y = pack_bf16x2(nrm(v[0].z, w0.z, b0.z), nrm(v[0].w, w0.w, b0.w));
        *reinterpret_cast<uint2*>(yr + col) = o;
    }
}

__global__ void __launch_bounds__(256)
norm_f32_to_bf16_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                        __nv_bfloat16* __restrict__ y, int rows, int D, float eps, int rms) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + static_cast<size_t>(row) * D;
    const RowStats st = row_stats([&](int) { return xr; }, 1, D, eps, rms != 0, lane);
    __nv_bfloat16* yr = y + static_cast<size_t>(row) * D;
    for (int i = lane * 8; i < D; i += 256) {
        const float4 a0 = *reinterpret_cast<const float4*>(xr + i);
        const float4 a1 = *reinterpret_cast<const float4*>(xr + i + 4);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + i));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + i + 4));
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (b) {
            b0 = __ldg(reinterpret_cast<const float4*>(b + i));
            b1 = __ldg(reinterpret_cast<const float4*>(b + i + 4));
        }
        uint4 o;
        o.x = pack_bf16x2((a0.x - st.mean) * st.rstd * w0.x + b0.x, (a0.y - st.mean) * st.rstd * w0.y + b0.y);
        o.y = pack_bf16x2((a0.z - st.mean) * st.rstd * w0.z + b0.z, (a0.w - st.mean) * st.rstd * w0.w + b0.w);
        o.z = pack_bf16x2((a1.x - st.mean) * st.rstd * w1.x + b1.x, (a1.y - st.mean) * st.rstd * w1.y + b1.y);
        o.w = pack_bf16x2((a1.z - st.mean) * st.rstd * w1.z + b1.z, (a1.w - st.mean) * st.rstd * w1.w + b1.w);
        *reinterpret_cast<uint4*>(yr + i) = o;
    }
}

// In-place RMSNorm of up to two column blocks of a bf16 matrix in one launch (blockIdx.y selects the block: q_norm and
// k_norm of InternViT-6B act on adjacent column blocks of the qkv matrix). kRegChunks > 0: the row block is held in
// registers (kRegChunks x 16 bytes per lane), so it is read from memory once; 0: two passes for any width.
template <int kRegChunks>
__global__ void __launch_bounds__(256)
rmsnorm_bf16_inplace_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ w0, const float* __restrict__ w1,
                            int rows, int D, int ld, int col0, int col1, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* __restrict__ w = blockIdx.y ? w1 : w0;
    __nv_bfloat16* xr = x + static_cast<size_t>(row) * ld + (blockIdx.y ? col1 : col0);
    float q = 0.f;
    uint4 buf[kRegChunks > 0 ? kRegChunks : 1];
    if (kRegChunks > 0) {
#pragma unroll
        for (int c = 0; c < kRegChunks; ++c) {
            const int i = lane * 8 + c * 256;
            buf[c] = make_uint4(0u, 0u, 0u, 0u);
            if (i < D) buf[c] = *reinterpret_cast<const uint4*>(xr + i);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&buf[c]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __bfloat1622float2(h[k]);
                q += f.x * f.x + f.y * f.y;
            }
        }
    } else {
        for (int i = lane * 8; i < D; i += 256) {
            const uint4 t = *reinterpret_cast<const uint4*>(xr + i);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __bfloat1622float2(h[k]);
                q += f.x * f.x + f.y * f.y;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(D) + eps);
    // InternRMSNorm: weight * (x * rstd)  (modeling_intern_vit.py:39-43), one rounding at the store
    auto scale_store = [&](uint4 t, int i) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
        const float4 wa = __ldg(reinterpret_cast<const float4*>(w + i)), wb = __ldg(reinterpret_cast<const float4*>(w + i + 4));
        const float ws[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(h[k]);
            h[k] = __floats2bfloat162_rn(f.x * rstd * ws[2 * k], f.y * rstd * ws[2 * k + 1]);
        }
        *reinterpret_cast<uint4*>(xr + i) = t;
    };
    if (kRegChunks > 0) {
#pragma unroll
        for (int c = 0; c < kRegChunks; ++c) {
            const int i = lane * 8 + c * 256;
            if (i < D) scale_store(buf[c], i);
        }
    } else {
        for (int i = lane * 8; i < D; i += 256) scale_store(*reinterpret_cast<const uint4*>(xr + i), i);
    }
}

// One thread per (patch, channel, ky) strip of `patch` pixels (28 contiguous bytes at patch 14), plus one thread per patch
// for the zero padding up to kpad: no per-element index arithmetic, destination rows written contiguously.
__global__ void __launch_bounds__(256)
im2col_kernel(const __nv_bfloat16* __restrict__ pv, __nv_bfloat16* __restrict__ patches, int n, int img, int patch,
              int kpad) {
    const int g = img / patch;
    const int strips = 3 * patch + 1;  // + the padding strip
    const int kreal = 3 * patch * patch;
    const size_t total = static_cast<size_t>(n) * g * g * strips;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int s = static_cast<int>(idx % strips);
        const size_t prow = idx / strips;
        uint32_t* dst = reinterpret_cast<uint32_t*>(patches + prow * kpad);
        if (s == strips - 1) {
            for (int k = kreal >> 1; k < (kpad >> 1); ++k) dst[k] = 0u;
            continue;
        }
        const int p = static_cast<int>(prow % (g * g));
        const int f = static_cast<int>(prow / (g * g));
        const int py = p / g, px = p - py * g;
        const int c = s / patch, ky = s - c * patch;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(
            pv + ((static_cast<size_t>(f) * 3 + c) * img + (py * patch + ky)) * img + px * patch);
        dst += (s * patch) >> 1;
        for (int k = 0; k < (patch >> 1); ++k) dst[k] = __ldg(src + k);
    }
}

__global__ void __launch_bounds__(256)
assemble_tokens_kernel(const float* __restrict__ pe, const float* __restrict__ cls, const float* __restrict__ pos,
                       float* __restrict__ x, int n, int P, int D) {
    const int d4 = D >> 2;
    const size_t total = static_cast<size_t>(n) * (P + 1) * d4;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % d4) * 4;
        const size_t tr = idx / d4;
        const int t = static_cast<int>(tr % (P + 1));
        const int f = static_cast<int>(tr / (P + 1));
        const float4 ps = __ldg(reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * D + c));
        float4 v;
        if (t == 0) v = __ldg(reinterpret_cast<const float4*>(cls + c));
        else v = *reinterpret_cast<const float4*>(pe + (static_cast<size_t>(f) * P + (t - 1)) * D + c);
        v.x += ps.x; v.y += ps.y; v.z += ps.z; v.w += ps.w;
        *reinterpret_cast<float4*>(x + tr * D + c) = v;
    }
}

// Warp-per-token variant for the LayerNorm-fused schedule: besides the fp32 residual stream it writes the bf16 copy
// the first qkv GEMM reads, and the token's (sum, sum of squares) into slice 0 of the statistics array (other slices 0).
__global__ void __launch_bounds__(256)
assemble_tokens_stats_kernel(const float* __restrict__ pe, const float* __restrict__ cls, const float* __restrict__ pos,
                             float* __restrict__ x, __nv_bfloat16* __restrict__ xb, float2* __restrict__ stats,
                             int nparts, int n, int P, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int M = n * (P + 1);
    if (row >= M) return;
    const int t = row % (P + 1), f = row / (P + 1);
    const float* src = t == 0 ? cls : pe + (static_cast<size_t>(f) * P + (t - 1)) * D;
    const float* ps = pos + static_cast<size_t>(t) * D;
    float s = 0.f, q = 0.f;
    for (int c = lane * 4; c < D; c += 128) {
        float4 v = t == 0 ? __ldg(reinterpret_cast<const float4*>(src + c)) : __ldcs(reinterpret_cast<const float4*>(src + c));
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(ps + c));
        v.x += p4.x; v.y += p4.y; v.z += p4.z; v.w += p4.w;
        s += (v.x + v.y) + (v.z + v.w);
        q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        *reinterpret_cast<float4*>(x + static_cast<size_t>(row) * D + c) = v;
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y);
        o.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(xb + static_cast<size_t>(row) * D + c) = o;
    }
    s = warp_sum(s);
    q = warp_sum(q);
    for (int p = lane; p < nparts; p += 32)
        stats[static_cast<size_t>(p) * M + row] = p == 0 ? make_float2(s, q) : make_float2(0.f, 0.f);
}

// LayerNorm folded into the following Linear (run once per weight load):
//   Wf[n,k] = bf16(W[n,k] * gamma[k]);  wsum[n] = sum_k Wf[n,k];  bf[n] = b[n] + sum_k W[n,k] * beta[k]
// so that  LN(x) W^T + b  =  rstd * (x Wf^T - mean * wsum) + bf.  One warp per output feature n.
__global__ void __launch_bounds__(256)
fold_norm_into_linear_kernel(const __nv_bfloat16* __restrict__ W, const float* __restrict__ gamma,
                             const float* __restrict__ beta, const float* __restrict__ b,
                             __nv_bfloat16* __restrict__ Wf, float* __restrict__ wsum, float* __restrict__ bf, int N,
                             int K) {
    const int nrow = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (nrow >= N) return;
    float sw = 0.f, sb = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float w = __bfloat162float(W[static_cast<size_t>(nrow) * K + k]);
        const __nv_bfloat16 wf = __float2bfloat16_rn(w * gamma[k]);
        Wf[static_cast<size_t>(nrow) * K + k] = wf;
        sw += __bfloat162float(wf);
        if (beta) sb += w * beta[k];
    }
    sw = warp_sum(sw);
    sb = warp_sum(sb);
    if (lane == 0) {
        wsum[nrow] = sw;
        bf[nrow] = (b ? b[nrow] : 0.f) + sb;
    }
}

__global__ void __launch_bounds__(256)
pixel_shuffle_ln_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                        __nv_bfloat16* __restrict__ y, int n, int g, int D, float eps, int apply_ln) {
    const int g2 = g >> 1;
    const int orow = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (orow >= n * g2 * g2) return;
    const int f = orow / (g2 * g2);
    const int ij = orow % (g2 * g2);
    const int I = ij / g2, J = ij % g2;
    const float* frame = x + static_cast<size_t>(f) * (g * g + 1) * D;
    // out[n,I,J,(2di+dj)*D + c] = in[n, 1 + g*(2I+di) + (2J+dj), c]   (probed on the reference)
    auto seg = [&](int s) { return frame + static_cast<size_t>(1 + g * (2 * I + (s >> 1)) + (2 * J + (s & 1))) * D; };
    RowStats st{0.f, 1.f};
    if (apply_ln) st = row_stats(seg, 4, D, eps, false, lane);
    __nv_bfloat16* yr = y + static_cast<size_t>(orow) * 4 * D;
    for (int s = 0; s < 4; ++s) {
        const float* xs = seg(s);
        for (int i = lane * 8; i < D; i += 256) {
            const float4 a0 = *reinterpret_cast<const float4*>(xs + i);
            const float4 a1 = *reinterpret_cast<const float4*>(xs + i + 4);
            float4 w0 = make_float4(1.f, 1.f, 1.f, 1.f), w1 = w0;
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (apply_ln) {
                w0 = __ldg(reinterpret_cast<const float4*>(w + s * D + i));
                w1 = __ldg(reinterpret_cast<const float4*>(w + s * D + i + 4));
                b0 = __ldg(reinterpret_cast<const float4*>(b + s * D + i));
                b1 = __ldg(reinterpret_cast<const float4*>(b + s * D + i + 4));
            }
            uint4 o;
            o.x = pack_bf16x2((a0.x - st.mean) * st.rstd * w0.x + b0.x, (a0.y - st.mean) * st.rstd * w0.y + b0.y);
            o.y = pack_bf16x2((a0.z - st.mean) * st.rstd * w0.z + b0.z, (a0.w - st.mean) * st.rstd * w0.w + b0.w);
            o.z = pack_bf16x2((a1.x - st.mean) * st.rstd * w1.x + b1.x, (a1.y - st.mean) * st.rstd * w1.y + b1.y);
            o.w = pack_bf16x2((a1.z - st.mean) * st.rstd * w1.z + b1.z, (a1.w - st.mean) * st.rstd * w1.w + b1.w);
            *reinterpret_cast<uint4*>(yr + s * D + i) = o;
        }
    }
}

// Register-resident variant for D = 128 * VPL: the 4 x D gathered row is read from global memory exactly once.
template <int VPL>
__global__ void __launch_bounds__(256)
pixel_shuffle_ln_reg_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                            __nv_bfloat16* __restrict__ y, int n, int g, float eps) {
    constexpr int D = 128 * VPL;
    const int g2 = g >> 1;
    const int orow = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (orow >= n * g2 * g2) return;
    const int f = orow / (g2 * g2);
    const int ij = orow % (g2 * g2);
    const int I = ij / g2, J = ij % g2;
    const float* frame = x + static_cast<size_t>(f) * (g * g + 1) * D;
    float4 v[4][VPL];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const float4* xs = reinterpret_cast<const float4*>(
            frame + static_cast<size_t>(1 + g * (2 * I + (s >> 1)) + (2 * J + (s & 1))) * D);
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[s][i] = xs[(i >> 1) * 64 + 2 * lane + (i & 1)];
    }
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int i = 0; i < VPL; ++i) sum += (v[s][i].x + v[s][i].y) + (v[s][i].z + v[s][i].w);
    const float mean = warp_sum(sum) * (1.0f / (4 * D));
    float q = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float a0 = v[s][i].x - mean, a1 = v[s][i].y - mean, a2 = v[s][i].z - mean, a3 = v[s][i].w - mean;
            q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / (4 * D)) + eps);
    __nv_bfloat16* yr = y + static_cast<size_t>(orow) * 4 * D;
    auto nrm = [&](float xv, float wv, float bv) { return (xv - mean) * rstd * wv + bv; };
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int i = 0; i < VPL; i += 2) {
            const int col = s * D + ((i >> 1) * 64 + 2 * lane) * 4;
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + col));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + col + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(b + col + 4));
            uint4 o;
            o.x = pack_bf16x2(nrm(v[s][i].x, w0.x, b0.x), nrm(v[s][i].y, w0.y, b0.y));
            o.y = pack_bf16x2(nrm(v[s][i].z, w0.z, b0.z), nrm(v[s][i].w, w0.w, b0.w));
            o.z = pack_bf16x2(nrm(v[s][i + 1].x, w1.x, b1.x), nrm(v[s][i + 1].y, w1.y, b1.y));
            o.w = pack_bf16x2(nrm(v[s][i + 1].z, w1.z, b1.z), nrm(v[s][i + 1].w, w1.w, b1.w));
            *reinterpret_cast<uint4*>(yr + col) = o;
        }
}

inline int grid_for(size_t work_items, int threads, int cap = 148 * 16) {
    size_t b = (work_items + threads - 1) / threads;
    if (b > static_cast<size_t>(cap)) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

}  // namespace

cudaError_t launch_norm_f32_to_bf16(const float* x, const float* w, const float* b, __nv_bfloat16* y, int rows, int D,
                                    float eps, int rms, cudaStream_t stream) {
    if (rows <= 0 || D <= 0 || (D % 8) != 0) return cudaErrorInvalidValue;
    const int grid = (rows + 7) / 8;
    switch (D) {
        case 128: norm_f32_to_bf16_reg_kernel<1><<<grid, 256, 0, stream>>>(x, w, b, y, rows, eps, rms); break;
        case 256: norm_f32_to_bf16_reg_kernel<2><<<grid, 256, 0, stream>>>(x, w, b, y, rows, eps, rms); break;
        case 1024: norm_f32_to_bf16_reg_kernel<8><<<grid, 256, 0, stream>>>(x, w, b, y, rows, eps, rms); break;
        default: norm_f32_to_bf16_kernel<<<grid, 256, 0, stream>>>(x, w, b, y, rows, D, eps, rms); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_rmsnorm_bf16_inplace(__nv_bfloat16* x, const float* w, int rows, int D, int ld, int col0, float eps,
                                        cudaStream_t stream, const float* w1, int col1) {
    if (rows <= 0 || (D % 8) != 0 || (ld % 8) != 0 || (col0 % 8) != 0 || (col1 % 8) != 0) return cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(w) & 15) != 0 || (w1 && (reinterpret_cast<uintptr_t>(w1) & 15) != 0))
        return cudaErrorInvalidValue;
    const dim3 grid((rows + 7) / 8, w1 ? 2 : 1);
    if (D <= 4 * 256)
        rmsnorm_bf16_inplace_kernel<4><<<grid, 256, 0, stream>>>(x, w, w1, rows, D, ld, col0, col1, eps);
    else if (D <= 13 * 256)
        rmsnorm_bf16_inplace_kernel<13><<<grid, 256, 0, stream>>>(x, w, w1, rows, D, ld, col0, col1, eps);
    else
        rmsnorm_bf16_inplace_kernel<0><<<grid, 256, 0, stream>>>(x, w, w1, rows, D, ld, col0, col1, eps);
    return cudaGetLastError();
}

cudaError_t launch_im2col(const __nv_bfloat16* pv, __nv_bfloat16* patches, int n, int img, int patch, int kpad,
                          cudaStream_t stream) {
    if (n <= 0 || (patch % 2) != 0 || (kpad % 2) != 0 || kpad < 3 * patch * patch) return cudaErrorInvalidValue;
    const int g = img / patch;
    const size_t total = static_cast<size_t>(n) * g * g * (3 * patch + 1);
    im2col_kernel<<<grid_for(total, 256), 256, 0, stream>>>(pv, patches, n, img, patch, kpad);
    return cudaGetLastError();
}

cudaError_t launch_assemble_tokens(const float* pe, const float* cls, const float* pos, float* x, int n, int P, int D,
                                   cudaStream_t stream, __nv_bfloat16* xb, float2* stats, int nparts) {
    if (n <= 0 || (D % 4) != 0) return cudaErrorInvalidValue;
    if (xb || stats) {
        if (!xb || !stats || nparts < 1) return cudaErrorInvalidValue;
        const int rows = n * (P + 1);
        assemble_tokens_stats_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(pe, cls, pos, x, xb, stats, nparts, n, P, D);
        return cudaGetLastError();
    }
    const size_t total = static_cast<size_t>(n) * (P + 1) * (D / 4);
    assemble_tokens_kernel<<<grid_for(total, 256), 256, 0, stream>>>(pe, cls, pos, x, n, P, D);
    return cudaGetLastError();
}

cudaError_t launch_fold_norm_into_linear(const __nv_bfloat16* W, const float* gamma, const float* beta, const float* b,
                                         __nv_bfloat16* Wf, float* wsum, float* bf, int N, int K, cudaStream_t stream) {
    if (N <= 0 || K <= 0) return cudaErrorInvalidValue;
    fold_norm_into_linear_kernel<<<(N + 7) / 8, 256, 0, stream>>>(W, gamma, beta, b, Wf, wsum, bf, N, K);
    return cudaGetLastError();
}

cudaError_t launch_pixel_shuffle_ln(const float* x, const float* w, const float* b, __nv_bfloat16* y, int n, int g,
                                    int D, float eps, int apply_ln, cudaStream_t stream) {
    if (n <= 0 || (g % 2) != 0 || (D % 8) != 0) return cudaErrorInvalidValue;
    const int rows = n * (g / 2) * (g / 2);
    if (apply_ln && D == 1024)
        pixel_shuffle_ln_reg_kernel<8><<<(rows + 7) / 8, 256, 0, stream>>>(x, w, b, y, n, g, eps);
    else if (apply_ln && D == 256)
        pixel_shuffle_ln_reg_kernel<2><<<(rows + 7) / 8, 256, 0, stream>>>(x, w, b, y, n, g, eps);
    else
        pixel_shuffle_ln_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(x, w, b, y, n, g, D, eps, apply_ln);
    return cudaGetLastError();
}

}  // namespace aigv
