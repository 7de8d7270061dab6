// Frame front end: uint8 HWC frame -> Pillow-exact bicubic (antialiased) resize -> 448x448 tiles -> ToTensor/Normalize
// -> bf16 NCHW, in one kernel per (frame batch, target size).
//
// Replaces, per frame, dynamic_preprocess (internvl/train/dataset.py:702-738: image.resize = Pillow BICUBIC, crop tiles,
// thumbnail) + build_transform eval branch (dataset.py:268-274) + .to(bfloat16) (stage1_eval.py:922).
// Pillow's resampler (src/libImaging/Resample.c, third party) is two passes of 22-bit fixed-point taps with a uint8
// rounding between the passes; this kernel keeps that exact integer arithmetic (horizontal pass -> uint8 tile staged in
// shared memory -> vertical pass), so the uint8 tile is bit-identical and the normalised value comes from a
// 3x256-entry fp32 table computed on the host the way torchvision does ((u/255 - mean)/std in fp32).
//
// The coefficient tables are built on the host in double precision exactly like precompute_coeffs /
// normalize_coeffs_8bpc (plan_resize below). HBM-bound: reads H*W*3 bytes, writes 3*448*448*2 bytes per tile.
#include <math.h>

#include <algorithm>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace aigv {

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;
constexpr int TX = 32, TY = 32;  // output pixels per block
constexpr int TMP_STRIDE = TX + 1;  // words per row of the horizontally resampled stage (odd: conflict-free both ways)

__device__ __forceinline__ uint32_t clip8(int v) {
    return static_cast<uint32_t>(__vimin_s32_relu(v >> PRECISION_BITS, 255));   // max(min(v, 255), 0): one VIMNMX
}
// byte i of w, zero extended: one PRMT (selector nibble 4 = byte 0 of the zero operand)
template <int I>
__device__ __forceinline__ int byte_of(uint32_t w) {
    return static_cast<int>(__byte_perm(w, 0u, 0x4440u | I));
}

// One block = a 32x32 block of output pixels of one frame. All shared-memory traffic is 32-bit or wider:
//   phase A: the source window (rows x bytes this block's taps touch) is copied into shared memory with coalesced
//            16-byte (or word) loads; the block's tap tables and the bf16 normalisation table are staged next to it
//   phase B: horizontal taps. Lane = source ROW (odd row stride: no bank conflicts), so a warp shares one output column,
//            one tap table (broadcast 16-byte loads) and one alignment; the 3 x 4 taps of a chunk come out of three
//            funnel-shifted words with one PRMT per byte and one IMAD per tap -> packed RGBX word per (row, column)
//   phase C: vertical taps on the RGBX words (lane = column), uint8 rounding, table look-up, bf16 NCHW stores
// Tap tables are zero padded to a multiple of 4 taps: padding taps multiply whatever the window holds there (possibly
// never-written shared memory -- integers, so harmless) by zero.
__global__ void __launch_bounds__(256, 5)
preprocess_kernel(const uint8_t* __restrict__ frames, int W, int64_t frame_stride, int64_t row_stride,
                  ResizePlan plan, const float* __restrict__ lut, int tile, int tiles_x, int tiles_per_frame,
                  int tile_offset, uint8_t* __restrict__ tiles_u8, __nv_bfloat16* __restrict__ out, ResizeGeom geo,
                  int aligned) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int ksx = plan.ksize_x, ksy = plan.ksize_y, sw = geo.span_words;
    int32_t* const kxs = reinterpret_cast<int32_t*>(sm);                        // [TX][ksx]
    int32_t* const kys = kxs + TX * ksx;                                        // [TY][ksy]
    uint32_t* const srct = reinterpret_cast<uint32_t*>(kys + TY * ksy);         // [rows_src][sw] source window
    uint32_t* const tmp = srct + geo.rows_src * sw;                             // [rows_tmp][TMP_STRIDE] RGBX, pass 1
    __nv_bfloat16* const lutb = reinterpret_cast<__nv_bfloat16*>(tmp + geo.rows_tmp * TMP_STRIDE);  // [3][256]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int nx = min(TX, plan.out_w - x0), ny = min(TY, plan.out_h - y0);
    const int2* const bx2 = reinterpret_cast<const int2*>(plan.bx) + x0;   // (first source column, taps) per output column
    const int2* const by2 = reinterpret_cast<const int2*>(plan.by) + y0;
    const int ysrc0 = __ldg(by2).x;
    const int2 bl = __ldg(by2 + ny - 1);
    const int nrows = bl.x + bl.y - ysrc0;
    // byte offset (within a source row) of srct column 0: 16-byte aligned when the rows can be read with 16-byte loads
    const int win0 = (__ldg(bx2).x * 3) & (aligned == 2 ? ~15 : ~3);
    const uint8_t* src = frames + f * frame_stride + static_cast<int64_t>(ysrc0) * row_stride + win0;

    // phase A
    for (int i = tid; i < nx * ksx; i += 256) kxs[i] = __ldg(plan.kx + static_cast<size_t>(x0) * ksx + i);
    for (int i = tid; i < ny * ksy; i += 256) kys[i] = __ldg(plan.ky + static_cast<size_t>(y0) * ksy + i);
    for (int i = tid; i < 768; i += 256) lutb[i] = __float2bfloat16_rn(__ldg(lut + i));
    {
        const int2 bxl = __ldg(bx2 + nx - 1);
        const int need = (bxl.x + bxl.y) * 3 - win0;                 // bytes of the window the real taps read
        if (aligned == 2) {
            const int nq = (need + 15) >> 4;                          // 16-byte pieces per row (4 * nq <= sw - 1)
            const int safe = static_cast<int>((row_stride - win0) >> 4);   // pieces that stay inside the row
            for (int rr = warp; rr < nrows; rr += 8) {
                const uint4* g = reinterpret_cast<const uint4*>(src + static_cast<int64_t>(rr) * row_stride);
                uint32_t* d = srct + rr * sw;
                for (int q = lane; q < nq; q += 32) {
                    const uint4 v = q < safe ? __ldg(g + q) : make_uint4(0u, 0u, 0u, 0u);
                    d[4 * q] = v.x; d[4 * q + 1] = v.y; d[4 * q + 2] = v.z; d[4 * q + 3] = v.w;
                }
            }
        } else if (aligned == 1) {
            const int nw = (need + 3) >> 2;
            const int safe = static_cast<int>((row_stride - win0) >> 2);
            for (int rr = warp; rr < nrows; rr += 8) {
                const uint32_t* g = reinterpret_cast<const uint32_t*>(src + static_cast<int64_t>(rr) * row_stride);
                for (int ww = lane; ww < nw; ww += 32) srct[rr * sw + ww] = ww < safe ? __ldg(g + ww) : 0u;
            }
        } else {
            const int nw = (need + 3) >> 2;
            const int nb = W * 3 - win0;
            for (int rr = warp; rr < nrows; rr += 8) {
                const uint8_t* g = src + static_cast<int64_t>(rr) * row_stride;
                for (int ww = lane; ww < nw; ww += 32) {
                    uint32_t v = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (ww * 4 + b < nb) v |= static_cast<uint32_t>(g[ww * 4 + b]) << (8 * b);
                    srct[rr * sw + ww] = v;
                }
            }
        }
    }
    __syncthreads();

    // phase B: horizontal taps. One item = one output column x TWO source rows (rr and rr + nrp): the pair shares the column's
    // bounds, alignment and tap table loads. Items in row-fastest order, 256 apart per step: no division in the loop.
    const int half = 1 << (PRECISION_BITS - 1);
    {
        const int nrp = (nrows + 1) >> 1;
        const int step_x = 256 / nrp, step_r = 256 - step_x * nrp;
        int xx = tid / nrp, rr = tid - xx * nrp;
        while (xx < nx) {
            const int2 b = __ldg(bx2 + xx);
            const int boff = b.x * 3 - win0;
            const uint32_t* p = srct + rr * sw + (boff >> 2);
            const uint32_t* q = p + nrp * sw;                       // second row (past the window for the odd last pair: unused)
            const int sh = (boff & 3) * 8;
            const int4* k4 = reinterpret_cast<const int4*>(kxs + xx * ksx);
            const int chunks = (b.y + 3) >> 2;
            int a0 = half, a1 = half, a2 = half, c0 = half, c1 = half, c2 = half;
            uint32_t w0 = p[0], u0 = q[0];
            for (int c = 0; c < chunks; ++c) {
                const int4 k = k4[c];
                {
                    const uint32_t w1 = p[1], w2 = p[2], w3 = p[3];
                    const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh),
                                   v2 = __funnelshift_r(w2, w3, sh);
                    a0 += byte_of<0>(v0) * k.x + byte_of<3>(v0) * k.y + byte_of<2>(v1) * k.z + byte_of<1>(v2) * k.w;
                    a1 += byte_of<1>(v0) * k.x + byte_of<0>(v1) * k.y + byte_of<3>(v1) * k.z + byte_of<2>(v2) * k.w;
                    a2 += byte_of<2>(v0) * k.x + byte_of<1>(v1) * k.y + byte_of<0>(v2) * k.z + byte_of<3>(v2) * k.w;
                    w0 = w3;
                }
                {
                    const uint32_t u1 = q[1], u2 = q[2], u3 = q[3];
                    const uint32_t v0 = __funnelshift_r(u0, u1, sh), v1 = __funnelshift_r(u1, u2, sh),
                                   v2 = __funnelshift_r(u2, u3, sh);
                    c0 += byte_of<0>(v0) * k.x + byte_of<3>(v0) * k.y + byte_of<2>(v1) * k.z + byte_of<1>(v2) * k.w;
                    c1 += byte_of<1>(v0) * k.x + byte_of<0>(v1) * k.y + byte_of<3>(v1) * k.z + byte_of<2>(v2) * k.w;
                    c2 += byte_of<2>(v0) * k.x + byte_of<1>(v1) * k.y + byte_of<0>(v2) * k.z + byte_of<3>(v2) * k.w;
                    u0 = u3;
                }
                p += 3;
                q += 3;
            }
            tmp[rr * TMP_STRIDE + xx] = clip8(a0) | (clip8(a1) << 8) | (clip8(a2) << 16);
            if (rr + nrp < nrows) tmp[(rr + nrp) * TMP_STRIDE + xx] = clip8(c0) | (clip8(c1) << 8) | (clip8(c2) << 16);
            xx += step_x;
            rr += step_r;
            if (rr >= nrp) { rr -= nrp; ++xx; }
        }
    }
    __syncthreads();

    // phase C: vertical taps, normalise, write bf16 NCHW (and optionally the uint8 HWC tile). Lane = column.
    if (lane >= nx) return;
    const int X = x0 + lane;
    const int tcol = X / tile, lx = X - tcol * tile;
    const size_t plane = static_cast<size_t>(tile) * tile;
    int trow = y0 / tile, ly = y0 - trow * tile + warp;
    for (int yy = warp; yy < ny; yy += 8, ly += 8) {
        while (ly >= tile) { ly -= tile; ++trow; }
        const int2 b = __ldg(by2 + yy);
        const uint32_t* p = tmp + (b.x - ysrc0) * TMP_STRIDE + lane;
        const int4* k4 = reinterpret_cast<const int4*>(kys + yy * ksy);
        const int chunks = (b.y + 3) >> 2;
        int a0 = half, a1 = half, a2 = half;
        for (int c = 0; c < chunks; ++c) {
            const int4 k = k4[c];
            const uint32_t v0 = p[0], v1 = p[TMP_STRIDE], v2 = p[2 * TMP_STRIDE], v3 = p[3 * TMP_STRIDE];
            a0 += byte_of<0>(v0) * k.x + byte_of<0>(v1) * k.y + byte_of<0>(v2) * k.z + byte_of<0>(v3) * k.w;
            a1 += byte_of<1>(v0) * k.x + byte_of<1>(v1) * k.y + byte_of<1>(v2) * k.z + byte_of<1>(v3) * k.w;
            a2 += byte_of<2>(v0) * k.x + byte_of<2>(v1) * k.y + byte_of<2>(v2) * k.z + byte_of<2>(v3) * k.w;
            p += 4 * TMP_STRIDE;
        }
        const uint32_t u0 = clip8(a0), u1 = clip8(a1), u2 = clip8(a2);
        const size_t tile_id = static_cast<size_t>(f) * tiles_per_frame + tile_offset + trow * tiles_x + tcol;
        const size_t pix = static_cast<size_t>(ly) * tile + lx;
        __nv_bfloat16* o = out + tile_id * 3 * plane + pix;
        o[0] = lutb[u0];
        o[plane] = lutb[256 + u1];
        o[2 * plane] = lutb[512 + u2];
        if (tiles_u8) {
            uint8_t* q = tiles_u8 + (tile_id * plane + pix) * 3;
            q[0] = static_cast<uint8_t>(u0);
            q[1] = static_cast<uint8_t>(u1);
            q[2] = static_cast<uint8_t>(u2);
        }
    }
}

double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

}  // namespace

// Host: Pillow's precompute_coeffs + normalize_coeffs_8bpc for one axis (bicubic, support 2.0, box = full axis).
// When in_size == out_size Pillow skips the pass; an identity tap (1<<22) reproduces that bit-exactly.
int plan_axis(int in_size, int out_size, std::vector<int32_t>& kk, std::vector<int32_t>& bounds) {
    if (in_size == out_size) {
        kk.assign(static_cast<size_t>(out_size) * 4, 0);
        bounds.resize(2 * static_cast<size_t>(out_size));
        for (int i = 0; i < out_size; ++i) { kk[4 * i] = 1 << PRECISION_BITS; bounds[2 * i] = i; bounds[2 * i + 1] = 1; }
        return 4;
    }
    const double in0 = 0.0, in1 = static_cast<double>(in_size);
    double scale, filterscale;
    filterscale = scale = (in1 - in0) / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    const int ksize_real = static_cast<int>(ceil(support)) * 2 + 1;
    const int ksize = (ksize_real + 3) & ~3;  // row stride of the table: zero padded to whole 4-tap chunks
    kk.assign(static_cast<size_t>(out_size) * ksize, 0);
    bounds.resize(2 * static_cast<size_t>(out_size));
    std::vector<double> w(ksize_real);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = in0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x) {
            double v = w[x];
            if (ww != 0.0) v /= ww;
            kk[static_cast<size_t>(xx) * ksize + x] = v < 0 ? static_cast<int>(-0.5 + v * (1 << PRECISION_BITS))
                                                            : static_cast<int>(0.5 + v * (1 << PRECISION_BITS));
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    return ksize;
}

// Shared-memory geometry of the kernel for one (source size, target size): the largest source window any 32x32 output
// block touches, including the reach of the zero-padded taps.
ResizeGeom plan_geometry(const std::vector<int32_t>& bx, const std::vector<int32_t>& by, int out_w, int out_h,
                         int ksize_x, int ksize_y) {
    ResizeGeom g{1, 1, 1};
    for (int y0 = 0; y0 < out_h; y0 += TY) {
        const int yl = (y0 + TY < out_h ? y0 + TY : out_h) - 1;
        g.rows_src = std::max(g.rows_src, by[2 * yl] + by[2 * yl + 1] - by[2 * y0]);
        g.rows_tmp = std::max(g.rows_tmp, by[2 * yl] - by[2 * y0] + ksize_y);
    }
    g.rows_tmp = std::max(g.rows_tmp, g.rows_src);
    for (int x0 = 0; x0 < out_w; x0 += TX) {
        const int xl = (x0 + TX < out_w ? x0 + TX : out_w) - 1;
        const int win0 = (bx[2 * x0] * 3) & ~15;  // the most conservative window start the kernel uses
        const int words = ((bx[2 * xl] * 3 - win0) >> 2) + 3 * (ksize_x >> 2) + 1;   // reach of the padded taps
        g.span_words = std::max(g.span_words, words);
    }
    // whole 16-byte pieces plus one word: odd row stride (lanes on consecutive rows hit distinct banks) and room for
    // the last piece phase A stores
    g.span_words = ((g.span_words + 3) & ~3) + 1;
    return g;
}

size_t preprocess_smem_bytes(const ResizeGeom& g, int ksize_x, int ksize_y) {
    return static_cast<size_t>(TX) * ksize_x * 4 + static_cast<size_t>(TY) * ksize_y * 4 +
           static_cast<size_t>(g.rows_src) * g.span_words * 4 + static_cast<size_t>(g.rows_tmp) * TMP_STRIDE * 4 + 768 * 2;
}

bool preprocess_fits(const ResizeGeom& g, int ksize_x, int ksize_y) {
    return preprocess_smem_bytes(g, ksize_x, ksize_y) <= 200 * 1024;
}

cudaError_t launch_preprocess_impl(const uint8_t* frames, int n, int H, int W, int64_t frame_stride,
                                   int64_t row_stride, const ResizePlan& plan, const ResizeGeom& geo,
                                   const float* lut, int tile, int tiles_x, int tiles_per_frame, int tile_offset,
                                   uint8_t* tiles_u8, __nv_bfloat16* out, cudaStream_t stream) {
    const size_t smem = preprocess_smem_bytes(geo, plan.ksize_x, plan.ksize_y);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    const uintptr_t al = reinterpret_cast<uintptr_t>(frames) | static_cast<uintptr_t>(frame_stride) |
                         static_cast<uintptr_t>(row_stride);
    const int aligned = (al & 15) == 0 ? 2 : ((al & 3) == 0 ? 1 : 0);   // 2: 16-byte row loads, 1: words, 0: bytes
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    dim3 grid((plan.out_w + TX - 1) / TX, (plan.out_h + TY - 1) / TY, n);
    preprocess_kernel<<<grid, 256, smem, stream>>>(frames, W, frame_stride, row_stride, plan, lut, tile, tiles_x,
                                                   tiles_per_frame, tile_offset, tiles_u8, out, geo, aligned);
    return cudaGetLastError();
}

}  // namespace aigv
