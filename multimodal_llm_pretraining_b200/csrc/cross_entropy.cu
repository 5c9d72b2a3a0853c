// Vocabulary cross-entropy, forward + backward in ONE pass over the logits (HBM-bound).
// Replaces ForCausalLMLoss / fixed_cross_entropy (HF:loss/loss_utils.py:28-67): upcast to fp32, log-softmax, NLL,
// mean over labels != ignore_index.  The reference materialises fp32 logits (6.6 GB at Pythia-1b mbs 16) and makes
// >= 3 passes; here a 512-thread block keeps one bf16 row (V <= 65536) in registers, reduces max / sum-exp with warp
// shuffles, and overwrites the row in place with dlogits = (softmax - onehot) / n_valid.
// Algorithmic bytes per token: 2*V read + 2*V write.
#include <stdlib.h>

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

template <int NT>
__device__ __forceinline__ float block_reduce_max_n(float v, float* sm) {
    v = warp_max(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float r = lane < (NT / 32) ? sm[lane] : -INFINITY;
    r = warp_max(r);
    __syncthreads();
    return r;
}
template <int NT>
__device__ __forceinline__ float block_reduce_sum_n(float v, float* sm) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    float r = lane < (NT / 32) ? sm[lane] : 0.f;
    r = warp_sum(r);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(CE_THREADS)
ce_kernel(elem_t* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ row_loss,
          const int* __restrict__ n_valid, int V, int64_t ld, int64_t ignore_index, int write_grad, const float* __restrict__ grad_scale) {
    pdl_prologue();
    __shared__ float sm[CE_THREADS / 32];
    __shared__ float s_xlabel;
    const int row = blockIdx.x;
    const int64_t label = labels[row];
    elem_t* rp = logits + static_cast<size_t>(row) * ld;
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
    const float inv_n = (grad_scale ? *grad_scale : 1.0f) / static_cast<float>(nv > 0 ? nv : 1);
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


// ---------------------------------------------------------------------------------------------------------------
// v2: the row lives in SHARED memory (one cp.async.bulk in, one out), two CTAs per SM. While one CTA reduces / rewrites its
// row, the other one's bulk copies are in flight, so HBM stays busy; registers no longer hold the row, so the occupancy that
// a 100 KB row used to cost is gone. One exp per element: pass A stores e = exp(x - max) as bf16 over x, pass B turns it
// into (e / sum - onehot) / n_valid in place. Persistent: CTA b handles rows b, b + grid, ...
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}

constexpr int CE2_THREADS = 512;

__global__ void __launch_bounds__(CE2_THREADS, 2)
ce_smem_kernel(elem_t* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ row_loss,
               const int* __restrict__ n_valid, int T, int V, int64_t ld, int64_t ignore_index, int write_grad,
               const float* __restrict__ grad_scale) {
    pdl_prologue();
    extern __shared__ __align__(16) uint8_t ce_smem[];
    __shared__ float sm[CE2_THREADS / 32];
    __shared__ uint64_t bar;
    elem_t* srow = reinterpret_cast<elem_t*>(ce_smem);
    const int n_vec = static_cast<int>(ld / 8);
    const int v_vec = V / 8;          // vectors that are entirely valid
    const uint32_t row_bytes = static_cast<uint32_t>(ld * 2);
    constexpr float LOG2E = 1.4426950408889634f;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    // grad_scale: the fp16 loss scale is folded into dlogits HERE — (softmax - onehot) / n_valid is ~1e-5 and would be flushed
    // to zero / subnormal in fp16 if the scale were applied only by the consumer GEMMs
    const float inv_n = write_grad ? (grad_scale ? *grad_scale : 1.0f) / static_cast<float>(max(*n_valid, 1)) : 0.f;
    uint32_t phase = 0;
    for (int row = blockIdx.x; row < T; row += gridDim.x) {
        const int64_t label = labels[row];
        elem_t* rp = logits + static_cast<size_t>(row) * ld;
        const bool ignored = (label == ignore_index);
        if (threadIdx.x == 0) {
            tma_store_wait_read<0>();  // the previous row's bulk store has finished reading the buffer
            if (!ignored) {
                mbar_expect_tx(&bar, row_bytes);
                bulk_load_1d(srow, rp, row_bytes, &bar);
            }
        }
        if (ignored) {  // uniform per block: zero gradient row, zero loss
            if (threadIdx.x == 0) row_loss[row] = 0.f;
            if (write_grad) {
                __syncthreads();  // buffer free (thread 0 waited above)
                for (int vi = threadIdx.x; vi < n_vec; vi += CE2_THREADS) st_shared_v4(srow + vi * 8, make_uint4(0, 0, 0, 0));
                fence_proxy_async_smem();
                __syncthreads();
                if (threadIdx.x == 0) {
                    bulk_store_1d(rp, srow, row_bytes);
                    tma_store_commit();
                }
            }
            continue;
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        // ---- pass 1: max
        float mx = -INFINITY;
        for (int vi = threadIdx.x; vi < n_vec; vi += CE2_THREADS) {
            const uint4 xv = ld_shared_v4(srow + vi * 8);
            const uint32_t w[4] = {xv.x, xv.y, xv.z, xv.w};
            if (vi < v_vec) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = bf2_to_f2(w[j]);
                    mx = fmaxf(mx, fmaxf(f.x, f.y));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = bf2_to_f2(w[j]);
                    const int c = vi * 8 + 2 * j;
                    if (c < V) mx = fmaxf(mx, f.x);
                    if (c + 1 < V) mx = fmaxf(mx, f.y);
                }
            }
        }
        const float x_label = bf_to_f(srow[label]);  // read before pass 2 overwrites the row (barrier inside the reduce)
        mx = block_reduce_max_n<CE2_THREADS>(mx, sm);
        const float mxs = mx * LOG2E;
        // ---- pass 2: e = exp(x - max) -> bf16 in place; sum in fp32 from the unrounded values
        float se = 0.f;
        for (int vi = threadIdx.x; vi < n_vec; vi += CE2_THREADS) {
            const uint4 xv = ld_shared_v4(srow + vi * 8);
            const uint32_t w[4] = {xv.x, xv.y, xv.z, xv.w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf2_to_f2(w[j]);
                const int c = vi * 8 + 2 * j;
                float e0, e1;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(f.x, LOG2E, -mxs)));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(f.y, LOG2E, -mxs)));
                if (vi >= v_vec) {
                    if (c >= V) e0 = 0.f;
                    if (c + 1 >= V) e1 = 0.f;
                }
                se += e0 + e1;
                o[j] = f2_to_bf2(e0, e1);
            }
            if (write_grad) st_shared_v4(srow + vi * 8, make_uint4(o[0], o[1], o[2], o[3]));
        }
        se = block_reduce_sum_n<CE2_THREADS>(se, sm);
        if (threadIdx.x == 0) row_loss[row] = mx + logf(se) - x_label;
        if (!write_grad) {
            __syncthreads();  // everyone is done reading before the next row's bulk load lands
            continue;
        }
        // ---- pass 3: (e / sum - onehot) / n_valid in place, then one bulk store
        const float scale = inv_n / se;
        const int lvec = static_cast<int>(label >> 3), lsub = static_cast<int>(label & 7);
        for (int vi = threadIdx.x; vi < n_vec; vi += CE2_THREADS) {
            const uint4 xv = ld_shared_v4(srow + vi * 8);
            const uint32_t w[4] = {xv.x, xv.y, xv.z, xv.w};
            float g[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf2_to_f2(w[j]);
                g[2 * j] = f.x * scale, g[2 * j + 1] = f.y * scale;
            }
            if (vi == lvec) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j == lsub) g[j] -= inv_n;
            }
            st_shared_v4(srow + vi * 8, make_uint4(f2_to_bf2(g[0], g[1]), f2_to_bf2(g[2], g[3]), f2_to_bf2(g[4], g[5]), f2_to_bf2(g[6], g[7])));
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store_1d(rp, srow, row_bytes);
            tma_store_commit();
        }
    }
    if (threadIdx.x == 0) tma_store_wait_all<0>();
}

// Also validates the labels (it scans every one anyway): a label that is neither ignore_index nor inside [0, V) would index
// outside the logits row. torch's cross_entropy raises a device-side assert for it; so does this (message + trap).
__global__ void __launch_bounds__(1024) count_valid_kernel(const int64_t* __restrict__ labels, int T, int64_t ignore_index, int V, int* out) {
    pdl_prologue();
    __shared__ int sm[32];
    int c = 0;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const int64_t l = labels[i];
        if (l != ignore_index) {
            ++c;
            if (V > 0 && (l < 0 || l >= V)) {
                printf("b200pt cross_entropy: label %lld at position %d is outside [0, %d) and is not ignore_index\n", (long long)l, i, V);
                __trap();
            }
        }
    }
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
    pdl_prologue();
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

extern "C" int b200_count_valid(const int64_t* labels, int T, int64_t ignore_index, int V, int* n_valid, b200_stream_t stream) {
    B200_REQUIRE(T > 0, "count_valid: T must be positive");
    launch_k(count_valid_kernel, dim3(1), dim3(1024), 0, as_stream(stream), labels, T, ignore_index, V, n_valid);
    return check_launch("count_valid");
}
extern "C" int b200_cross_entropy(void* logits, const int64_t* labels, float* row_loss, const int* n_valid, int T, int V,
                                  int64_t ld, int64_t ignore_index, int write_grad, const float* grad_scale_dev, b200_stream_t stream) {
    B200_REQUIRE(T > 0 && V > 0 && ld >= V && ld % 8 == 0, "cross_entropy: need ld >= V and ld %% 8 == 0 (V=%d ld=%lld)", V, (long long)ld);
    B200_REQUIRE(aligned16(logits), "cross_entropy: logits must be 16B aligned");
    static const bool old_path = getenv("B200_CE_OLD") != nullptr;  // perf triage only
    const size_t row_bytes = static_cast<size_t>(ld) * 2;
    if (!old_path && row_bytes <= 220 * 1024) {
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(ce_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
            if (e != cudaSuccess) return fail(-2, "cross_entropy: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            configured = true;
        }
        const int per_sm = row_bytes <= 110 * 1024 ? 2 : 1;  // two rows per SM when they fit next to each other
        int grid = num_sms() * per_sm;
        if (grid > T) grid = T;
        launch_k(ce_smem_kernel, dim3(grid), dim3(CE2_THREADS), row_bytes, as_stream(stream), static_cast<elem_t*>(logits), labels, row_loss, n_valid, T, V, ld,
                                                                            ignore_index, write_grad, grad_scale_dev);
        return check_launch("cross_entropy");
    }
    B200_REQUIRE(ld <= static_cast<int64_t>(CE_THREADS) * CE_NV * 8, "cross_entropy: ld %lld > %d unsupported", (long long)ld, CE_THREADS * CE_NV * 8);
    launch_k(ce_kernel, dim3(T), dim3(CE_THREADS), 0, as_stream(stream), static_cast<elem_t*>(logits), labels, row_loss, n_valid, V, ld, ignore_index, write_grad, grad_scale_dev);
    return check_launch("cross_entropy");
}
extern "C" int b200_mean_loss(const float* row_loss, const int* n_valid, int T, float* loss_out, b200_stream_t stream) {
    launch_k(mean_loss_kernel, dim3(1), dim3(1024), 0, as_stream(stream), row_loss, n_valid, T, loss_out);
    return check_launch("mean_loss");
}
