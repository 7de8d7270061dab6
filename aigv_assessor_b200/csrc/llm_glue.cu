// First pieces ABOVE the seam, behind their own C-ABI entry points:
//   * visual-token splice: input_embeds = tok_embeddings[input_ids]; every <IMG_CONTEXT> slot but the last of a sample
//     takes the next visual token, the last one the sample's motion token
//     (/root/reference/internvl/model/internvl_chat_eval2/modeling_internvl_chat.py:324,351-378);
//   * the mlpscore regression head: 4096 -> 1024 -> 256 -> 64 -> 16 -> 1 with ReLU after EVERY layer, applied to
//     hidden_states[-1][:, -4, :] (eval2 :82-94,469-475).
// Both are data movement / GEMV work, HBM-bound: coalesced 16-byte accesses, one warp per token row.
#include "kernels.h"

namespace aigv {

namespace {

// One block per sample: rank[b, n] = index of position n among the sample's <IMG_CONTEXT> slots (-1 elsewhere), cnt[b] =
// number of slots, last[b] = position of the last slot (-1 if none). Block-wide exclusive scan over the sample's tokens.
__global__ void __launch_bounds__(256)
splice_rank_kernel(const long long* __restrict__ ids, int N, long long ctx_id, int* __restrict__ rank,
                   int* __restrict__ cnt, int* __restrict__ last) {
    __shared__ int warp_sums[8];
    __shared__ int carry_s;
    const int b = blockIdx.x;
    const long long* row = ids + static_cast<size_t>(b) * N;
    int* rrow = rank + static_cast<size_t>(b) * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int last_pos = -1;
    for (int base = 0; base < N; base += 256) {
        const int n = base + threadIdx.x;
        const int sel = (n < N && row[n] == ctx_id) ? 1 : 0;
        const unsigned m = __ballot_sync(0xffffffffu, sel);
        const int in_warp = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) warp_sums[warp] = __popc(m);
        __syncthreads();
        int before = carry_s;
        for (int w = 0; w < warp; ++w) before += warp_sums[w];
        if (n < N) rrow[n] = sel ? before + in_warp : -1;
        if (sel) last_pos = n;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += warp_sums[w];
            carry_s += tot;
        }
        __syncthreads();
    }
    // block max of last_pos
    for (int o = 16; o > 0; o >>= 1) last_pos = max(last_pos, __shfl_xor_sync(0xffffffffu, last_pos, o));
    if (lane == 0) warp_sums[warp] = last_pos;
    __syncthreads();
    if (threadIdx.x == 0) {
        int lp = -1;
        for (int w = 0; w < 8; ++w) lp = max(lp, warp_sums[w]);
        last[b] = lp;
        cnt[b] = carry_s;
    }
}

// One warp per token: copy its C-element row (16 bytes per lane per step) from the embedding table, the visual tokens
// or the motion tokens. Visual rows are consumed in flattened (sample-major) order, like the boolean-mask assignment
// of the reference; a sample's visual-row base is the number of non-last slots of the samples before it.
__global__ void __launch_bounds__(256)
splice_copy_kernel(const long long* __restrict__ ids, const int* __restrict__ rank, const int* __restrict__ cnt,
                   const int* __restrict__ last, const __nv_bfloat16* __restrict__ tok, const __nv_bfloat16* __restrict__ vit,
                   const __nv_bfloat16* __restrict__ motion, __nv_bfloat16* __restrict__ out, int B, int N, int C, int V,
                   int n_vit, int* __restrict__ err) {
    const int tokidx = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tokidx >= B * N) return;
    const int b = tokidx / N, n = tokidx - b * N;
    const int r = rank[tokidx];
    const __nv_bfloat16* src;
    if (r < 0) {
        long long id = ids[tokidx];
        if (id < 0 || id >= V) {  // out-of-range id: flag it, write zeros (nn.Embedding would raise)
            if (lane == 0) atomicOr(err, 1);
            src = nullptr;
        } else {
            src = tok + static_cast<size_t>(id) * C;
        }
    } else if (n == last[b]) {
        src = motion + static_cast<size_t>(b) * C;
    } else {
        int base = 0;
        for (int bb = 0; bb < b; ++bb) base += max(cnt[bb] - 1, 0);
        const int vr = base + r;
        if (vr >= n_vit) {
            if (lane == 0) atomicOr(err, 2);
            src = nullptr;
        } else {
            src = vit + static_cast<size_t>(vr) * C;
        }
    }
    uint4* d = reinterpret_cast<uint4*>(out + static_cast<size_t>(tokidx) * C);
    const int vecs = C / 8;
    if (src) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        for (int i = lane; i < vecs; i += 32) d[i] = __ldg(s4 + i);
    } else {
        for (int i = lane; i < vecs; i += 32) d[i] = make_uint4(0, 0, 0, 0);
    }
    // the last sample's last token also validates the visual-row count (the reference's shape check, eval2 :377)
    if (tokidx == B * N - 1 && lane == 0) {
        int total = 0;
        for (int bb = 0; bb < B; ++bb) total += max(cnt[bb] - 1, 0);
        if (total != n_vit) atomicOr(err, 4);
    }
}

// mlpscore: one block per sample; activations ping-pong through shared memory; one warp per output neuron.
struct MlpScoreW {
    const float* w[5];
    const float* b[5];
};
__global__ void __launch_bounds__(256)
mlpscore_kernel(const float* __restrict__ x, int ldx, MlpScoreW W, float* __restrict__ score, float* __restrict__ pre) {
    __shared__ float a0[4096];
    __shared__ float a1[1024];
    const int dims[6] = {4096, 1024, 256, 64, 16, 1};
    const int s = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += 256) a0[i] = x[static_cast<size_t>(s) * ldx + i];
    __syncthreads();
    float* in = a0;
    float* outp = a1;
    for (int l = 0; l < 5; ++l) {
        const int K = dims[l], Nn = dims[l + 1];
        for (int o = warp; o < Nn; o += 8) {
            const float* wr = W.w[l] + static_cast<size_t>(o) * K;
            float acc = 0.f;
            for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(wr + k), in[k], acc);
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (lane == 0) {
                const float v = acc + __ldg(W.b[l] + o);
                if (l == 4 && pre) pre[s] = v;
                outp[o] = fmaxf(v, 0.f);  // ReLU after every layer, the last one included (eval2 :93)
            }
        }
        __syncthreads();
        float* t = in; in = outp; outp = t;
    }
    if (threadIdx.x == 0) score[s] = in[0];
}

}  // namespace

cudaError_t launch_splice_tokens(const long long* ids, const __nv_bfloat16* tok, const __nv_bfloat16* vit,
                                 const __nv_bfloat16* motion, __nv_bfloat16* out, int* work, int B, int N, int C, int V,
                                 int n_vit, long long ctx_id, cudaStream_t stream) {
    if (B <= 0 || N <= 0 || C <= 0 || (C % 8) != 0) return cudaErrorInvalidValue;
    int* rank = work;                       // [B*N]
    int* cnt = work + static_cast<size_t>(B) * N;   // [B]
    int* last = cnt + B;                    // [B]
    int* err = last + B;                    // [1]
    cudaError_t e = cudaMemsetAsync(err, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    splice_rank_kernel<<<B, 256, 0, stream>>>(ids, N, ctx_id, rank, cnt, last);
    const long long toks = static_cast<long long>(B) * N;
    splice_copy_kernel<<<static_cast<unsigned>((toks + 7) / 8), 256, 0, stream>>>(ids, rank, cnt, last, tok, vit, motion, out,
                                                                                  B, N, C, V, n_vit, err);
    return cudaGetLastError();
}

cudaError_t launch_mlpscore(const float* x, int ldx, const float* const* w, const float* const* b, float* score, float* pre,
                            int n, cudaStream_t stream) {
    if (n <= 0) return cudaErrorInvalidValue;
    MlpScoreW W;
    for (int l = 0; l < 5; ++l) { W.w[l] = w[l]; W.b[l] = b[l]; }
    mlpscore_kernel<<<n, 256, 0, stream>>>(x, ldx, W, score, pre);
    return cudaGetLastError();
}

}  // namespace aigv
