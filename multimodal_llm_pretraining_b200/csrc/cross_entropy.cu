// Vocabulary cross-entropy, forward + backward in ONE pass over the logits (HBM-bound).
// Replaces ForCausalLMLoss / fixed_cross_entropy (HF:loss/loss_utils.py:28-67): upcast to fp32, log-softmax, NLL,
// mean over labels != ignore_index.  The reference materialises fp32 logits (6.6 GB at Pythia-1b mbs 16) and makes
// >= 3 passes; here a 512-thread block keeps one bf16 row (V <= 65536) in registers, reduces max / sum-exp with warp
// shuffles, and overwrites the row in place with dlogits = (softmax - onehot) / n_valid.
// Algorithmic bytes per token: 2*V read + 2*V write.
#include "api.h"
#include "common.cuh"

namespace b200 {

constexpr int CE_THREADS = 512;
constexpr int CE_NV = 16;  // 16-byte vectors per thread -> V <= 512*16*8 = 65536

__device__ __forceinline__ float block_reduce_max(float v, float* sm) {
    v = warp_max(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float r = lane < (CE_THREADS / 32) ? sm[lane] : -INFINITY;
    r = warp_max(r);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* sm) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float r = lane < (CE_THREADS / 32) ? sm[lane] : 0.f;
    r = warp_sum(r);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(CE_THREADS)
ce_kernel(__nv_bfloat16* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ row_loss,
          const int* __restrict__ n_valid, int V, int64_t ld, int64_t ignore_index, int write_grad) {
    __shared__ float sm[CE_THREADS / 32];
    __shared__ float s_xlabel;
    const int row = blockIdx.x;
    const int64_t label = labels[row];
    __nv_bfloat16* rp = logits + static_cast<size_t>(row) * ld;
    const int n_vec = static_cast<int>(ld / 8);
    const bool ignored = (label == ignore_index);

    if (ignored) {  // uniform per block
        if (threadIdx.x == 0) row_loss[row] = 0.f;
        if (write_grad)
            for (int vi = threadIdx.x; vi < n_vec; vi += CE_THREADS) st_v4(rp + vi * 8, make_uint4(0, 0, 0, 0));
        return;
    }
    if (threadIdx.x == 0) s_xlabel = bf_to_f(rp[label]);

    uint4 xv[CE_NV];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < CE_NV; ++i) {
        const int vi = i * CE_THREADS + threadIdx.x;
        if (vi < n_vec) {
            xv[i] = *reinterpret_cast<const uint4*>(rp + vi * 8);
            const uint32_t w[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf2_to_f2(w[j]);
                const int c = vi * 8 + 2 * j;
                if (c < V) mx = fmaxf(mx, f.x);
                if (c + 1 < V) mx = fmaxf(mx, f.y);
            }
        }
    }
    mx = block_reduce_max(mx, sm);
    constexpr float LOG2E = 1.4426950408889634f;
    const float mxs = mx * LOG2E;
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < CE_NV; ++i) {
        const int vi = i * CE_THREADS + threadIdx.x;
        if (vi < n_vec) {
            const uint32_t w[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf2_to_f2(w[j]);
                const int c = vi * 8 + 2 * j;
                if (c < V) se += exp2f(f.x * LOG2E - mxs);
                if (c + 1 < V) se += exp2f(f.y * LOG2E - mxs);
            }
        }
    }
    se = block_reduce_sum(se, sm);  // contains __syncthreads: s_xlabel is visible and read before any write below
    const float lse = mx + logf(se);
    if (threadIdx.x == 0) row_loss[row] = lse - s_xlabel;
    if (!write_grad) return;
    const int nv = *n_valid;
    const float inv_n = 1.0f / static_cast<float>(nv > 0 ? nv : 1);
    const float lses = lse * LOG2E;
#pragma unroll
    for (int i = 0; i < CE_NV; ++i) {
        const int vi = i * CE_THREADS + threadIdx.x;
        if (vi < n_vec) {
            const uint32_t w[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf2_to_f2(w[j]);
                const int c = vi * 8 + 2 * j;
                float g0 = 0.f, g1 = 0.f;
                if (c < V) g0 = (exp2f(f.x * LOG2E - lses) - (c == label ? 1.f : 0.f)) * inv_n;
                if (c + 1 < V) g1 = (exp2f(f.y * LOG2E - lses) - (c + 1 == label ? 1.f : 0.f)) * inv_n;
                o[j] = f2_to_bf2(g0, g1);
            }
            st_v4(rp + vi * 8, make_uint4(o[0], o[1], o[2], o[3]));
        }
    }
}

__global__ void __launch_bounds__(1024) count_valid_kernel(const int64_t* __restrict__ labels, int T, int64_t ignore_index, int* out) {
    __shared__ int sm[32];
    int c = 0;
    for (int i = threadIdx.x; i < T; i += blockDim.x) c += (labels[i] != ignore_index);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
        c = sm[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (threadIdx.x == 0) *out = c;
    }
}
// deterministic: fixed per-thread strided order, then fixed tree
__global__ void __launch_bounds__(1024) mean_loss_kernel(const float* __restrict__ row_loss, const int* __restrict__ n_valid, int T, float* out) {
    __shared__ float sm[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < T; i += blockDim.x) s += row_loss[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = warp_sum(sm[threadIdx.x]);
        if (threadIdx.x == 0) {
            const int nv = *n_valid;
            *out = s / static_cast<float>(nv > 0 ? nv : 1);
        }
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_count_valid(const int64_t* labels, int T, int64_t ignore_index, int* n_valid, b200_stream_t stream) {
    B200_REQUIRE(T > 0, "count_valid: T must be positive");
    count_valid_kernel<<<1, 1024, 0, as_stream(stream)>>>(labels, T, ignore_index, n_valid);
    return check_launch("count_valid");
}
extern "C" int b200_cross_entropy(void* logits, const int64_t* labels, float* row_loss, const int* n_valid, int T, int V,
                                  int64_t ld, int64_t ignore_index, int write_grad, b200_stream_t stream) {
    B200_REQUIRE(T > 0 && V > 0 && ld >= V && ld % 8 == 0, "cross_entropy: need ld >= V and ld %% 8 == 0 (V=%d ld=%lld)", V, (long long)ld);
    B200_REQUIRE(ld <= static_cast<int64_t>(CE_THREADS) * CE_NV * 8, "cross_entropy: ld %lld > %d unsupported", (long long)ld, CE_THREADS * CE_NV * 8);
    B200_REQUIRE(aligned16(logits), "cross_entropy: logits must be 16B aligned");
    ce_kernel<<<T, CE_THREADS, 0, as_stream(stream)>>>(static_cast<__nv_bfloat16*>(logits), labels, row_loss, n_valid, V, ld, ignore_index, write_grad);
    return check_launch("cross_entropy");
}
extern "C" int b200_mean_loss(const float* row_loss, const int* n_valid, int T, float* loss_out, b200_stream_t stream) {
    mean_loss_kernel<<<1, 1024, 0, as_stream(stream)>>>(row_loss, n_valid, T, loss_out);
    return check_launch("mean_loss");
}
