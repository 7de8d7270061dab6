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

#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace aigv {

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;
constexpr int TX = 32, TY = 32;  // output pixels per block

__device__ __forceinline__ int clip8(int v) {
    v >>= PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// One block = a 32x32 block of output pixels of one frame.
//   phase A: the source rows/columns this block's taps touch are copied into shared memory with coalesced 4-byte loads
//   phase B: horizontal taps from shared memory -> uint8 rows in shared memory (Pillow's imTemp)
//   phase C: vertical taps -> uint8 -> LUT -> bf16 NCHW (and optionally the uint8 HWC tile)
__global__ void __launch_bounds__(256)
preprocess_kernel(const uint8_t* __restrict__ frames, int H, int W, int64_t frame_stride, int64_t row_stride,
                  ResizePlan plan, const float* __restrict__ lut, int tile, int tiles_x, int tiles_per_frame,
                  int tile_offset, uint8_t* __restrict__ tiles_u8, __nv_bfloat16* __restrict__ out, int max_rows,
                  int span_bytes, int aligned4) {
    extern __shared__ __align__(16) uint8_t sm[];
    uint8_t* srct = sm;                               // [max_rows][span_bytes] source window
    uint8_t* tmp = sm + max_rows * span_bytes;        // [max_rows][TX][3] horizontally resampled rows
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int x_last = min(x0 + TX, plan.out_w) - 1;
    const int y_last = min(y0 + TY, plan.out_h) - 1;
    const int ysrc0 = plan.by[2 * y0];
    const int nrows = plan.by[2 * y_last] + plan.by[2 * y_last + 1] - ysrc0;
    const int xsrc0 = plan.bx[2 * x0];
    const int xsrc1 = plan.bx[2 * x_last] + plan.bx[2 * x_last + 1];
    const uint8_t* src = frames + f * frame_stride;
    const int byte0 = xsrc0 * 3;
    const int byte1 = xsrc1 * 3;

    // phase A
    if (aligned4) {
        const int w0 = byte0 & ~3;                     // window starts at a 4-byte boundary of the row
        const int nwords = (byte1 - w0 + 3) >> 2;
        const int row_words = static_cast<int>(row_stride >> 2);
        for (int idx = threadIdx.x; idx < nrows * nwords; idx += blockDim.x) {
            const int rr = idx / nwords, ww = idx - rr * nwords;
            const int word = (w0 >> 2) + ww;
            uint32_t v = 0;
            if (word < row_words)                      // never read past the end of the row
                v = __ldg(reinterpret_cast<const uint32_t*>(src + static_cast<int64_t>(ysrc0 + rr) * row_stride) + word);
            *reinterpret_cast<uint32_t*>(srct + rr * span_bytes + ww * 4) = v;
        }
    } else {
        const int w0 = byte0;
        const int nbytes = byte1 - w0;
        for (int idx = threadIdx.x; idx < nrows * nbytes; idx += blockDim.x) {
            const int rr = idx / nbytes, bb = idx - rr * nbytes;
            srct[rr * span_bytes + bb] = src[static_cast<int64_t>(ysrc0 + rr) * row_stride + w0 + bb];
        }
    }
    const int win0 = aligned4 ? (byte0 & ~3) : byte0;  // byte offset (within the row) of srct column 0
    __syncthreads();

    // phase B: horizontal taps
    for (int idx = threadIdx.x; idx < nrows * TX; idx += blockDim.x) {
        const int rr = idx / TX, xx = idx % TX;
        const int X = x0 + xx;
        int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
        if (X < plan.out_w) {
            const int xmin = plan.bx[2 * X], cnt = plan.bx[2 * X + 1];
            const int32_t* k = plan.kx + static_cast<size_t>(X) * plan.ksize_x;
            const uint8_t* p = srct + rr * span_bytes + (xmin * 3 - win0);
            for (int t = 0; t < cnt; ++t) {
                const int kk = __ldg(k + t);
                a0 += p[3 * t + 0] * kk;
                a1 += p[3 * t + 1] * kk;
                a2 += p[3 * t + 2] * kk;
            }
        }
        uint8_t* d = tmp + (rr * TX + xx) * 3;
        d[0] = static_cast<uint8_t>(clip8(a0));
        d[1] = static_cast<uint8_t>(clip8(a1));
        d[2] = static_cast<uint8_t>(clip8(a2));
    }
    __syncthreads();

    // phase C: vertical taps, normalise, write bf16 NCHW (and optionally the uint8 HWC tile)
    const int tx = threadIdx.x % TX;
    const int X = x0 + tx;
    if (X >= plan.out_w) return;
    for (int yy = threadIdx.x / TX; yy < TY; yy += blockDim.x / TX) {
        const int Y = y0 + yy;
        if (Y >= plan.out_h) break;
        const int ymin = plan.by[2 * Y] - ysrc0, cnt = plan.by[2 * Y + 1];
        const int32_t* k = plan.ky + static_cast<size_t>(Y) * plan.ksize_y;
        int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
        for (int t = 0; t < cnt; ++t) {
            const int kk = __ldg(k + t);
            const uint8_t* p = tmp + ((ymin + t) * TX + tx) * 3;
            a0 += p[0] * kk;
            a1 += p[1] * kk;
            a2 += p[2] * kk;
        }
        const int u0 = clip8(a0), u1 = clip8(a1), u2 = clip8(a2);
        const int t_idx = (Y / tile) * tiles_x + (X / tile);
        const int ly = Y % tile, lx = X % tile;
        const size_t tile_id = static_cast<size_t>(f) * tiles_per_frame + tile_offset + t_idx;
        const size_t plane = static_cast<size_t>(tile) * tile;
        __nv_bfloat16* o = out + tile_id * 3 * plane + static_cast<size_t>(ly) * tile + lx;
        o[0] = __float2bfloat16_rn(__ldg(lut + u0));
        o[plane] = __float2bfloat16_rn(__ldg(lut + 256 + u1));
        o[2 * plane] = __float2bfloat16_rn(__ldg(lut + 512 + u2));
        if (tiles_u8) {
            uint8_t* q = tiles_u8 + (tile_id * plane + static_cast<size_t>(ly) * tile + lx) * 3;
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
        kk.assign(out_size, 1 << PRECISION_BITS);
        bounds.resize(2 * static_cast<size_t>(out_size));
        for (int i = 0; i < out_size; ++i) { bounds[2 * i] = i; bounds[2 * i + 1] = 1; }
        return 1;
    }
    const double in0 = 0.0, in1 = static_cast<double>(in_size);
    double scale, filterscale;
    filterscale = scale = (in1 - in0) / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
    kk.assign(static_cast<size_t>(out_size) * ksize, 0);
    bounds.resize(2 * static_cast<size_t>(out_size));
    std::vector<double> w(ksize);
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

// Largest number of source rows any 32-row output block touches (sizes the shared-memory stage).
int plan_max_rows(const std::vector<int32_t>& by, int out_h) {
    int best = 1;
    for (int y0 = 0; y0 < out_h; y0 += TY) {
        const int yl = (y0 + TY < out_h ? y0 + TY : out_h) - 1;
        const int n = by[2 * yl] + by[2 * yl + 1] - by[2 * y0];
        if (n > best) best = n;
    }
    return best;
}

// Widest source window (in bytes, rounded for 4-byte staging) any 32-column output block touches.
int plan_max_span_bytes(const std::vector<int32_t>& bx, int out_w) {
    int best = 4;
    for (int x0 = 0; x0 < out_w; x0 += TX) {
        const int xl = (x0 + TX < out_w ? x0 + TX : out_w) - 1;
        const int b0 = (bx[2 * x0] * 3) & ~3;
        const int b1 = (bx[2 * xl] + bx[2 * xl + 1]) * 3;
        const int n = ((b1 - b0 + 3) >> 2) << 2;
        if (n > best) best = n;
    }
    return best + 4;
}

bool preprocess_fits(int max_rows, int span_bytes) {
    return static_cast<size_t>(max_rows) * (TX * 3 + span_bytes) <= 200 * 1024;
}

cudaError_t launch_preprocess_impl(const uint8_t* frames, int n, int H, int W, int64_t frame_stride,
                                   int64_t row_stride, const ResizePlan& plan, int max_rows, int span_bytes,
                                   const float* lut, int tile, int tiles_x, int tiles_per_frame, int tile_offset,
                                   uint8_t* tiles_u8, __nv_bfloat16* out, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(max_rows) * (TX * 3 + span_bytes);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    const int aligned4 = ((reinterpret_cast<uintptr_t>(frames) | static_cast<uintptr_t>(frame_stride) |
                           static_cast<uintptr_t>(row_stride)) & 3) == 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    dim3 grid((plan.out_w + TX - 1) / TX, (plan.out_h + TY - 1) / TY, n);
    preprocess_kernel<<<grid, 256, smem, stream>>>(frames, H, W, frame_stride, row_stride, plan, lut, tile, tiles_x,
                                                   tiles_per_frame, tile_offset, tiles_u8, out, max_rows, span_bytes,
                                                   aligned4);
    return cudaGetLastError();
}

}  // namespace aigv
