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
            __nv_bfloat16* __restrict__ p16, int64_t state_base, const int64_t* __restrict__ chunk_start,
            const int32_t* __restrict__ chunk_len, const int32_t* __restrict__ chunk_group,
            const int64_t* __restrict__ chunk_state, const AdamGroups groups, const float* __restrict__ grad_scale,
            int zero_grad) {
    pdl_prologue();
    const int64_t start = chunk_start[blockIdx.x];
    const int len = chunk_len[blockIdx.x];
    const b200_adam_group h = groups.g[chunk_group[blockIdx.x]];
    const float gs = grad_scale ? *grad_scale : 1.0f;
    float* pp = p + start;
    float* gp = g + start;
    const int64_t soff = chunk_state ? chunk_state[blockIdx.x] : (start - state_base);
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
        if (p16) p16[start + i] = __float2bfloat16_rn(P);
    }
}

__global__ void __launch_bounds__(512) sumsq_kernel(const float* __restrict__ x, size_t n, float* out) {
    pdl_prologue();
    __shared__ float sm[16];
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
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < 16 ? sm[threadIdx.x] : 0.f;
        s = warp_sum(s);
        if (threadIdx.x == 0) atomicAdd(out, s);
    }
}

__global__ void clip_coef_kernel(const float* sumsq, float max_norm, float* norm_out, float* coef_out) {
    pdl_prologue();
    const float nrm = sqrtf(*sumsq);
    if (norm_out) *norm_out = nrm;
    float c = 1.0f;
    if (max_norm > 0.f) c = fminf(1.0f, max_norm / (nrm + 1e-6f));
    *coef_out = c;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_adam_step(float* p, float* g, float* m, float* v, void* p_bf16, int64_t state_base,
                              const int64_t* chunk_start, const int32_t* chunk_len, const int32_t* chunk_group,
                              const int64_t* chunk_state, int n_chunks, const b200_adam_group* groups, int n_groups,
                              const float* grad_scale_dev, int zero_grad, b200_stream_t stream) {
    B200_REQUIRE(n_groups > 0 && n_groups <= B200_ADAM_MAX_GROUPS, "adam_step: n_groups %d out of range", n_groups);
    if (n_chunks == 0) return 0;
    AdamGroups gs;
    for (int i = 0; i < n_groups; ++i) gs.g[i] = groups[i];
    launch_k(adam_kernel, dim3(n_chunks), dim3(256), 0, as_stream(stream), p, g, m, v, static_cast<__nv_bfloat16*>(p_bf16), state_base,
                                                         chunk_start, chunk_len, chunk_group, chunk_state, gs, grad_scale_dev, zero_grad);
    return check_launch("adam_step");
}
extern "C" int b200_sumsq(const float* x, size_t n, float* out, b200_stream_t stream) {
    B200_REQUIRE(aligned16(x), "sumsq: x must be 16B aligned");
    if (n == 0) return 0;
    size_t blocks = (n / 4 + 511) / 512;
    const size_t cap = static_cast<size_t>(num_sms()) * 4;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    launch_k(sumsq_kernel, dim3(static_cast<int>(blocks)), dim3(512), 0, as_stream(stream), x, n, out);
    return check_launch("sumsq");
}
extern "C" int b200_clip_coef(const float* sumsq, float max_norm, float* norm_out, float* coef_out, b200_stream_t stream) {
    launch_k(clip_coef_kernel, dim3(1), dim3(1), 0, as_stream(stream), sumsq, max_norm, norm_out, coef_out);
    return check_launch("clip_coef");
}
