// Fused multi-tensor Adam / AdamW + gradient sum-of-squares + clip coefficient (HBM-bound).
// Replaces torch.optim.Adam's foreach path (~10 elementwise passes; fused=True is disabled in the reference,
// src/train.py:75-77) and DeepSpeed FusedAdam (csrc/adam/multi_tensor_adam.cu, selected by src/train.py:79-81,157-167).
// One pass: read g,p,m,v (16 B/param), write p,m,v + bf16 shadow (+ zeroed g) (14-18 B/param).
#include <math.h>

#include "api.h"
#include "common.cuh"

namespace b200 {

struct AdamGroups {
    b200_adam_group g[B200_ADAM_MAX_GROUPS];
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const b200_adam_group& h, float gscale) {
    g *= gscale;
    if (h.adamw_mode) {
        p *= (1.0f - h.lr * h.weight_decay);
    } else {
        g = fmaf(h.weight_decay, p, g);
    }
    m = h.beta1 * m + (1.0f - h.beta1) * g;
    v = h.beta2 * v + (1.0f - h.beta2) * g * g;
    // torch.optim.Adam (single-tensor path): denom = sqrt(v)/sqrt(bc2) + eps ; p -= (lr/bc1) * m/denom
    const float denom = sqrtf(v) / sqrtf(h.bias_corr2) + h.eps;
    p -= (h.lr / h.bias_corr1) * (m / denom);
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            elem_t* __restrict__ p16, int64_t state_base, const int64_t* __restrict__ chunk_start,
            const int32_t* __restrict__ chunk_len, const int32_t* __restrict__ chunk_group,
            const int64_t* __restrict__ chunk_state, const AdamGroups groups, const float* __restrict__ grad_scale,
            int zero_grad, const int* __restrict__ skip_flag, int g_packed, int p_packed) {
    pdl_prologue();
    const int64_t start = chunk_start[blockIdx.x];
    const int len = chunk_len[blockIdx.x];
    const int64_t soff = chunk_state ? chunk_state[blockIdx.x] : (start - state_base);
    // g_packed: the gradients live in a packed shard buffer indexed like the moments (ZeRO-2: a rank only ever holds the reduced
    // gradients of the slices it owns); otherwise g is the full flat buffer indexed like p
    float* gp = g + (g_packed ? soff : start);
    if (skip_flag && *skip_flag) {  // fp16 overflow step: parameters and moments untouched (GradScaler.step semantics)
        if (zero_grad)
            for (int i = threadIdx.x; i < len; i += blockDim.x) gp[i] = 0.f;
        return;
    }
    const b200_adam_group h = groups.g[chunk_group[blockIdx.x]];
    const float gs = grad_scale ? *grad_scale : 1.0f;
    // p_packed: the fp32 master exists only for the slices this rank owns, packed like the moments (true ZeRO: 4 B/param/W)
    float* pp = p + (p_packed ? soff : start);
    float* mp = m + soff;
    float* vp = v + soff;
    const bool vec_ok = ((start & 3) == 0) && ((soff & 3) == 0);
    const int n4 = vec_ok ? len / 4 : 0;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 P = reinterpret_cast<float4*>(pp)[i];
        const float4 G = reinterpret_cast<float4*>(gp)[i];
        float4 M = reinterpret_cast<float4*>(mp)[i];
        float4 V = reinterpret_cast<float4*>(vp)[i];
        adam_update(P.x, G.x, M.x, V.x, h, gs);
        adam_update(P.y, G.y, M.y, V.y, h, gs);
        adam_update(P.z, G.z, M.z, V.z, h, gs);
        adam_update(P.w, G.w, M.w, V.w, h, gs);
        reinterpret_cast<float4*>(pp)[i] = P;
        reinterpret_cast<float4*>(mp)[i] = M;
        reinterpret_cast<float4*>(vp)[i] = V;
        if (zero_grad) reinterpret_cast<float4*>(gp)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p16) reinterpret_cast<uint2*>(p16 + start)[i] = make_uint2(f2_to_bf2(P.x, P.y), f2_to_bf2(P.z, P.w));
    }
    for (int i = n4 * 4 + threadIdx.x; i < len; i += blockDim.x) {
        float P = pp[i], M = mp[i], V = vp[i];
        adam_update(P, gp[i], M, V, h, gs);
        pp[i] = P, mp[i] = M, vp[i] = V;
        if (zero_grad) gp[i] = 0.f;
        if (p16) p16[start + i] = f_to_elem(P);
    }
}

// Sum of squares, DETERMINISTIC: every block reduces a fixed set of elements in a fixed order into partials[blockIdx.x]
// (no atomics), then one block adds the partials in index order. Replicas that hold bit-identical gradients (DDP after the
// all-reduce) therefore compute bit-identical norms and clip coefficients, and stay bit-identical after the optimizer
// step — what torch's clip_grad_norm_ gives HF DDP. (A float atomicAdd per block made the last ulp order-dependent.)
constexpr int SUMSQ_THREADS = 512;
constexpr int SUMSQ_MAX_BLOCKS = 1024;

__device__ __forceinline__ float block_sum_512(float s, float* sm) {
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (SUMSQ_THREADS / 32) ? sm[threadIdx.x] : 0.f;
        r = warp_sum(r);
    }
    return r;  // valid in thread 0
}

__global__ void __launch_bounds__(SUMSQ_THREADS) sumsq_partial_kernel(const float* __restrict__ x, size_t n, float* __restrict__ partials) {
    pdl_prologue();
    __shared__ float sm[SUMSQ_THREADS / 32];
    float s = 0.f;
    const size_t n4 = n / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float4 v = x4[i];
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0)
        for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += x[i] * x[i];
    s = block_sum_512(s, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// one block per chunk of the optimizer's chunk table (the slices a ZeRO rank owns): partials[c] = sum of squares of chunk c
__global__ void __launch_bounds__(SUMSQ_THREADS) sumsq_chunks_kernel(const float* __restrict__ x, const int64_t* __restrict__ chunk_start,
                                                                      const int32_t* __restrict__ chunk_len, float* __restrict__ partials) {
    pdl_prologue();
    __shared__ float sm[SUMSQ_THREADS / 32];
    const int64_t start = chunk_start[blockIdx.x];
    const int len = chunk_len[blockIdx.x];
    const float* xp = x + start;
    float s = 0.f;
    const int n4 = (start & 3) == 0 ? len / 4 : 0;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(xp)[i];
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int i = n4 * 4 + threadIdx.x; i < len; i += blockDim.x) s += xp[i] * xp[i];
    s = block_sum_512(s, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// out[0] += partials[0] + partials[1] + ... in a fixed order (thread t takes indices t, t + 512, ...; fixed shuffle tree)
__global__ void __launch_bounds__(SUMSQ_THREADS) sumsq_finalize_kernel(const float* __restrict__ partials, int n, float* out) {
    pdl_prologue();
    __shared__ float sm[SUMSQ_THREADS / 32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
    s = block_sum_512(s, sm);
    if (threadIdx.x == 0) *out += s;
}

// norm = sqrt(sumsq) / loss_scale; coef = min(1, max_norm / (norm + 1e-6)) / loss_scale (the factor that turns the SCALED
// gradients of an fp16 run into clipped true gradients; loss_scale = NULL means 1). found_inf = 1 if the sum of squares is not
// finite, i.e. some gradient overflowed (the inf/nan check of torch.amp.GradScaler.unscale_ / DeepSpeed's has_overflow, for
// free: one non-finite element makes the sum non-finite).
__global__ void clip_coef_kernel(const float* sumsq, float max_norm, const float* loss_scale, float* norm_out, float* coef_out,
                                 int* found_inf) {
    pdl_prologue();
    const float ss = *sumsq;
    const float inv = loss_scale ? 1.0f / *loss_scale : 1.0f;
    const float nrm = sqrtf(ss) * inv;
    if (norm_out) *norm_out = nrm;
    float c = 1.0f;
    if (max_norm > 0.f) c = fminf(1.0f, max_norm / (nrm + 1e-6f));
    *coef_out = c * inv;
    if (found_inf) *found_inf = (ss == ss && fabsf(ss) != INFINITY) ? 0 : 1;
}

// Dynamic loss scale (torch.amp.GradScaler._amp_update_scale_ with DeepSpeed's hysteresis, src/train.py:143-150):
//   overflow : hysteresis_left -= 1; once it reaches 0 (or hysteresis <= 1) scale = max(scale * backoff, min_scale); tracker = 0
//   clean    : tracker += 1; every growth_interval clean steps scale *= growth, tracker = 0, hysteresis_left = hysteresis
__global__ void loss_scale_update_kernel(float* scale, int* growth_tracker, int* hysteresis_left, const int* found_inf,
                                         float growth_factor, float backoff_factor, int growth_interval, float min_scale,
                                         int hysteresis) {
    pdl_prologue();
    if (*found_inf) {
        int left = *hysteresis_left - 1;
        if (hysteresis <= 1 || left <= 0) {
            *scale = fmaxf(*scale * backoff_factor, min_scale);
            left = hysteresis <= 1 ? hysteresis : 1;  // DeepSpeed keeps cutting on consecutive overflows once hysteresis is spent
        }
        *hysteresis_left = left;
        *growth_tracker = 0;
    } else {
        const int t = *growth_tracker + 1;
        if (t >= growth_interval) {
            const float grown = *scale * growth_factor;
            if (grown == grown && fabsf(grown) != INFINITY) *scale = grown;
            *growth_tracker = 0;
            *hysteresis_left = hysteresis;
        } else {
            *growth_tracker = t;
        }
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_adam_step(float* p, float* g, float* m, float* v, void* p_bf16, int64_t state_base,
                              const int64_t* chunk_start, const int32_t* chunk_len, const int32_t* chunk_group,
                              const int64_t* chunk_state, int n_chunks, const b200_adam_group* groups, int n_groups,
                              const float* grad_scale_dev, int zero_grad, const int* skip_flag_dev, int g_packed, int p_packed, b200_stream_t stream) {
    B200_REQUIRE(n_groups > 0 && n_groups <= B200_ADAM_MAX_GROUPS, "adam_step: n_groups %d out of range", n_groups);
    if (n_chunks == 0) return 0;
    AdamGroups gs;
    for (int i = 0; i < n_groups; ++i) gs.g[i] = groups[i];
    launch_k(adam_kernel, dim3(n_chunks), dim3(256), 0, as_stream(stream), p, g, m, v, static_cast<elem_t*>(p_bf16), state_base,
                                                         chunk_start, chunk_len, chunk_group, chunk_state, gs, grad_scale_dev, zero_grad, skip_flag_dev, g_packed, p_packed);
    return check_launch("adam_step");
}
extern "C" size_t b200_sumsq_workspace_bytes(void) { return SUMSQ_MAX_BLOCKS * sizeof(float); }
extern "C" int b200_sumsq(const float* x, size_t n, float* out, void* workspace, size_t workspace_bytes, b200_stream_t stream) {
    B200_REQUIRE(aligned16(x), "sumsq: x must be 16B aligned");
    B200_REQUIRE(workspace && workspace_bytes >= b200_sumsq_workspace_bytes(), "sumsq: workspace too small (%zu < %zu)", workspace_bytes, b200_sumsq_workspace_bytes());
    if (n == 0) return 0;
    size_t blocks = (n / 4 + SUMSQ_THREADS - 1) / SUMSQ_THREADS;
    size_t cap = static_cast<size_t>(num_sms()) * 4;
    if (cap > SUMSQ_MAX_BLOCKS) cap = SUMSQ_MAX_BLOCKS;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    float* partials = static_cast<float*>(workspace);
    launch_k(sumsq_partial_kernel, dim3(static_cast<int>(blocks)), dim3(SUMSQ_THREADS), 0, as_stream(stream), x, n, partials);
    int rc = check_launch("sumsq");
    if (rc) return rc;
    launch_k(sumsq_finalize_kernel, dim3(1), dim3(SUMSQ_THREADS), 0, as_stream(stream), static_cast<const float*>(partials), static_cast<int>(blocks), out);
    return check_launch("sumsq_finalize");
}
extern "C" int b200_sumsq_chunks(const float* x, const int64_t* chunk_start, const int32_t* chunk_len, int n_chunks, float* out,
                                 float* partials, b200_stream_t stream) {
    if (n_chunks == 0) return 0;
    B200_REQUIRE(partials != nullptr, "sumsq_chunks: partials [n_chunks] is required");
    launch_k(sumsq_chunks_kernel, dim3(n_chunks), dim3(SUMSQ_THREADS), 0, as_stream(stream), x, chunk_start, chunk_len, partials);
    int rc = check_launch("sumsq_chunks");
    if (rc) return rc;
    launch_k(sumsq_finalize_kernel, dim3(1), dim3(SUMSQ_THREADS), 0, as_stream(stream), static_cast<const float*>(partials), n_chunks, out);
    return check_launch("sumsq_finalize");
}
extern "C" int b200_clip_coef(const float* sumsq, float max_norm, const float* loss_scale_dev, float* norm_out, float* coef_out,
                              int* found_inf_out, b200_stream_t stream) {
    launch_k(clip_coef_kernel, dim3(1), dim3(1), 0, as_stream(stream), sumsq, max_norm, loss_scale_dev, norm_out, coef_out, found_inf_out);
    return check_launch("clip_coef");
}
extern "C" int b200_loss_scale_update(float* scale, int* growth_tracker, int* hysteresis_left, const int* found_inf, float growth_factor,
                                      float backoff_factor, int growth_interval, float min_scale, int hysteresis, b200_stream_t stream) {
    B200_REQUIRE(scale && growth_tracker && hysteresis_left && found_inf, "loss_scale_update: null state");
    B200_REQUIRE(growth_factor >= 1.f && backoff_factor > 0.f && backoff_factor <= 1.f && growth_interval > 0, "loss_scale_update: bad factors");
    launch_k(loss_scale_update_kernel, dim3(1), dim3(1), 0, as_stream(stream), scale, growth_tracker, hysteresis_left, found_inf, growth_factor,
             backoff_factor, growth_interval, min_scale, hysteresis);
    return check_launch("loss_scale_update");
}
