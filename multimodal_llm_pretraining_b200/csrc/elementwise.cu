// HBM-bound elementwise / gather kernels: exact-erf GELU, in-place partial rotary on packed qkv, embedding gather and
// scatter-add, dtype casts.  16-byte vector accesses, grid-stride loops sized to a multiple of the SM count.
#include "api.h"
#include "common.cuh"

namespace b200 {

static inline int ew_grid(size_t n_vec, int threads) {
    size_t blocks = (n_vec + threads - 1) / threads;
    const size_t cap = static_cast<size_t>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    return static_cast<int>(blocks);
}

// ----------------------------------------------------------------------------------------------- GELU
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, size_t n_vec) {
    pdl_prologue();
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_vec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint4 v = ld_nc_v4(x + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = bf2_to_f2(w[j]);
            o[j] = f2_to_bf2(gelu_erf(f.x), gelu_erf(f.y));
        }
        st_v4(y + i, make_uint4(o[0], o[1], o[2], o[3]));
    }
}
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx, size_t n_vec) {
    pdl_prologue();
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_vec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint4 v = ld_nc_v4(x + i);
        const uint4 g = ld_nc_v4(dy + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = bf2_to_f2(w[j]);
            const float2 d = bf2_to_f2(gw[j]);
            o[j] = f2_to_bf2(d.x * gelu_erf_grad(f.x), d.y * gelu_erf_grad(f.y));
        }
        st_v4(dx + i, make_uint4(o[0], o[1], o[2], o[3]));
    }
}

// ----------------------------------------------------------------------------------------------- RoPE
// One thread handles one (token, head, which in {q,k}, pair-of-8) : 8 low dims [i0,i0+8) and their partners at +rot/2.
// When rot/2 is not a multiple of 8 (e.g. rot=20 for pythia-2.8b) falls back to a scalar pair-per-thread kernel.
__global__ void __launch_bounds__(256)
rope_vec_kernel(elem_t* __restrict__ qkv, const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                int T, int S, int nh, int hd, int half, int inverse) {
    pdl_prologue();
    const int vec_per = half / 8;  // vectors per (token, head, q|k)
    const size_t total = static_cast<size_t>(T) * nh * 2 * vec_per;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int vi = static_cast<int>(idx % vec_per);
        size_t r = idx / vec_per;
        const int which = static_cast<int>(r % 2);
        r /= 2;
        const int head = static_cast<int>(r % nh);
        const int t = static_cast<int>(r / nh);
        const int pos = t % S;
        elem_t* base = qkv + (static_cast<size_t>(t) * nh + head) * 3 * hd + which * hd + vi * 8;
        const uint4 lo = *reinterpret_cast<const uint4*>(base);
        const uint4 hi = *reinterpret_cast<const uint4*>(base + half);
        const uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w};
        const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w};
        const float* c = cos_tab + static_cast<size_t>(pos) * half + vi * 8;
        const float* s = sin_tab + static_cast<size_t>(pos) * half + vi * 8;
        uint32_t ol[4], oh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 a = bf2_to_f2(lw[j]);
            const float2 b = bf2_to_f2(hw[j]);
            const float c0 = c[2 * j], c1 = c[2 * j + 1];
            float s0 = s[2 * j], s1 = s[2 * j + 1];
            if (inverse) s0 = -s0, s1 = -s1;
            // out_lo = lo*cos - hi*sin ; out_hi = hi*cos + lo*sin   (HF apply_rotary_pos_emb / rotate_half)
            ol[j] = f2_to_bf2(a.x * c0 - b.x * s0, a.y * c1 - b.y * s1);
            oh[j] = f2_to_bf2(b.x * c0 + a.x * s0, b.y * c1 + a.y * s1);
        }
        *reinterpret_cast<uint4*>(base) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
        *reinterpret_cast<uint4*>(base + half) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    }
}
__global__ void __launch_bounds__(256)
rope_scalar_kernel(elem_t* __restrict__ qkv, const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                   int T, int S, int nh, int hd, int half, int inverse) {
    pdl_prologue();
    const size_t total = static_cast<size_t>(T) * nh * 2 * half;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(idx % half);
        size_t r = idx / half;
        const int which = static_cast<int>(r % 2);
        r /= 2;
        const int head = static_cast<int>(r % nh);
        const int t = static_cast<int>(r / nh);
        const int pos = t % S;
        elem_t* base = qkv + (static_cast<size_t>(t) * nh + head) * 3 * hd + which * hd + i;
        const float a = bf_to_f(base[0]), b = bf_to_f(base[half]);
        const float c = cos_tab[static_cast<size_t>(pos) * half + i];
        float s = sin_tab[static_cast<size_t>(pos) * half + i];
        if (inverse) s = -s;
        base[0] = f_to_elem(a * c - b * s);
        base[half] = f_to_elem(b * c + a * s);
    }
}

// Rotary widths whose half is not a multiple of 8 (Pythia-2.8b: head_dim 80, rot 20): one thread per (token, head, q|k) keeps the
// NV 16-byte vectors that cover the rot leading dims in registers, rotates the HALF pairs and writes the vectors back — 48 bytes per
// thread in full sectors instead of the scalar kernel's two 2-byte accesses per thread (130 us per call, 1.9 % of the 2.8b step).
template <int HALF>
__global__ void __launch_bounds__(256)
rope_row_kernel(elem_t* __restrict__ qkv, const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                int T, int S, int nh, int hd, int inverse) {
    pdl_prologue();
    constexpr int ROT = 2 * HALF, NV = (ROT + 7) / 8;
    const size_t total = static_cast<size_t>(T) * nh * 2;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int which = static_cast<int>(idx & 1);
        const size_t r = idx >> 1;
        const int head = static_cast<int>(r % nh);
        const int t = static_cast<int>(r / nh);
        const int pos = t % S;
        elem_t* base = qkv + (static_cast<size_t>(t) * nh + head) * 3 * hd + which * hd;
        uint4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const uint4*>(base + i * 8);
        float f[NV * 8];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 a = bf2_to_f2(w[j]);
                f[i * 8 + 2 * j] = a.x, f[i * 8 + 2 * j + 1] = a.y;
            }
        }
        const float* c = cos_tab + static_cast<size_t>(pos) * HALF;
        const float* sn = sin_tab + static_cast<size_t>(pos) * HALF;
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
            const float a = f[i], b = f[i + HALF], cc = c[i];
            const float ss = inverse ? -sn[i] : sn[i];
            f[i] = a * cc - b * ss;
            f[i + HALF] = b * cc + a * ss;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i)
            *reinterpret_cast<uint4*>(base + i * 8) = make_uint4(f2_to_bf2(f[i * 8], f[i * 8 + 1]), f2_to_bf2(f[i * 8 + 2], f[i * 8 + 3]),
                                                                f2_to_bf2(f[i * 8 + 4], f[i * 8 + 5]), f2_to_bf2(f[i * 8 + 6], f[i * 8 + 7]));
    }
}

// ----------------------------------------------------------------------------------------------- Embedding
// one warp per token row; 16-byte vectors
__global__ void __launch_bounds__(256)
embedding_fwd_kernel(const int64_t* __restrict__ ids0, const elem_t* __restrict__ t0,
                     const int64_t* __restrict__ ids1, const elem_t* __restrict__ t1,
                     const int64_t* __restrict__ ids2, const elem_t* __restrict__ t2,
                     elem_t* __restrict__ out, int T, int h, int vocab0) {
    pdl_prologue();
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x * warps_per_block + (threadIdx.x >> 5); t < T; t += gridDim.x * warps_per_block) {
        if (vocab0 > 0 && (ids0[t] < 0 || ids0[t] >= vocab0)) {  // torch's embedding raises a device-side assert here
            if (lane == 0) printf("b200pt embedding: token id %lld at position %d is outside [0, %d)\n", (long long)ids0[t], t, vocab0);
            __trap();
        }
        const elem_t* r0 = t0 + static_cast<size_t>(ids0[t]) * h;
        const elem_t* r1 = ids1 ? t1 + static_cast<size_t>(ids1[t]) * h : nullptr;
        const elem_t* r2 = ids2 ? t2 + static_cast<size_t>(ids2[t]) * h : nullptr;
        elem_t* o = out + static_cast<size_t>(t) * h;
        for (int c = lane * 8; c < h; c += 256) {
            uint4 v = *reinterpret_cast<const uint4*>(r0 + c);
            if (r1 || r2) {
                uint32_t w[4] = {v.x, v.y, v.z, v.w};
                float f[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 a = bf2_to_f2(w[j]);
                    f[2 * j] = a.x, f[2 * j + 1] = a.y;
                }
                if (r1) {
                    const uint4 u = *reinterpret_cast<const uint4*>(r1 + c);
                    const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 a = bf2_to_f2(uw[j]);
                        f[2 * j] += a.x, f[2 * j + 1] += a.y;
                    }
                }
                if (r2) {
                    const uint4 u = *reinterpret_cast<const uint4*>(r2 + c);
                    const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 a = bf2_to_f2(uw[j]);
                        f[2 * j] += a.x, f[2 * j + 1] += a.y;
                    }
                }
                v = make_uint4(f2_to_bf2(f[0], f[1]), f2_to_bf2(f[2], f[3]), f2_to_bf2(f[4], f[5]), f2_to_bf2(f[6], f[7]));
            }
            st_v4(o + c, v);
        }
    }
}
// scatter-add into fp32 table gradient. Random ids over a 50k vocab rarely collide, so plain fp32 REDs are cheap.
__global__ void __launch_bounds__(256)
embedding_bwd_kernel(const int64_t* __restrict__ ids, const elem_t* __restrict__ dout, float* __restrict__ dtable,
                     int T, int h, int64_t skip_id) {
    pdl_prologue();
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x * warps_per_block + (threadIdx.x >> 5); t < T; t += gridDim.x * warps_per_block) {
        const int64_t id = ids[t];
        if (id == skip_id) continue;  // nn.Embedding(padding_idx=...): the padding row receives no gradient
        float* drow = dtable + static_cast<size_t>(id) * h;
        const elem_t* g = dout + static_cast<size_t>(t) * h;
        for (int c = lane * 8; c < h; c += 256) {
            const uint4 v = ld_nc_v4(g + c);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 a = bf2_to_f2(w[j]);
                atomicAdd(drow + c + 2 * j, a.x);
                atomicAdd(drow + c + 2 * j + 1, a.y);
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------- casts
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, size_t n_vec4) {
    pdl_prologue();
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_vec4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float4 v = src[i];
        dst[i] = make_uint2(f2_to_bf2(v.x, v.y), f2_to_bf2(v.z, v.w));
    }
}
__global__ void cast_f32_bf16_tail(const float* __restrict__ src, elem_t* __restrict__ dst, size_t start, size_t n) {
    pdl_prologue();
    const size_t i = start + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i < n) dst[i] = f_to_elem(src[i]);
}
__global__ void __launch_bounds__(256)
scale_f32_kernel(float* __restrict__ x, size_t n, const float* __restrict__ scale_dev, float scale_host) {
    pdl_prologue();
    const float s = (scale_dev ? *scale_dev : 1.0f) * scale_host;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        x[i] *= s;
}


// ----------------------------------------------------------------------------------------------- column sums (bias grads)
// partial[rc][c] = sum over rows of chunk rc of x[r][c]; block = 32 column-lanes (8 cols each) x 8 row-lanes.
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const elem_t* __restrict__ x, int rows, int cols, int64_t ld, int rows_per_chunk,
                      float* __restrict__ partial) {
    pdl_prologue();
    __shared__ float sm[8][32][9];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = (blockIdx.x * 32 + cx) * 8;
    const int r_begin = blockIdx.y * rows_per_chunk;
    const int r_end = min(rows, r_begin + rows_per_chunk);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (col < cols) {
        for (int r = r_begin + ry; r < r_end; r += 8) {
            const uint4 v = ld_nc_v4(x + static_cast<size_t>(r) * ld + col);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf2_to_f2(w[j]);
                acc[2 * j] += f.x;
                acc[2 * j + 1] += f.y;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[ry][cx][j] = acc[j];
    __syncthreads();
    if (ry == 0 && col < cols) {
        float* out = partial + static_cast<size_t>(blockIdx.y) * cols + col;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += sm[k][cx][j];
            out[j] = t;
        }
    }
}
__global__ void __launch_bounds__(256)
colsum_finalize_kernel(const float* __restrict__ partial, int n_partial, int cols, float* __restrict__ out, float* __restrict__ out2,
                       const float* __restrict__ scale) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float t = 0.f;
    for (int p = 0; p < n_partial; ++p) t += partial[static_cast<size_t>(p) * cols + c];
    if (scale) t *= *scale;
    out[c] += t;
    if (out2) out2[c] += t;  // two biases that share one upstream gradient (attention.dense and mlp.dense_4h_to_h of a GPT-NeoX block)
}


// ----------------------------------------------------------------------------------------------- RoBERTa position ids
// pos[b, s] = (number of non-pad tokens in ids[b, 0..s]) * (ids[b, s] != pad) + pad   (HF:modeling_roberta.py:146-159).
// One warp per row: inclusive warp scan over 32-token chunks with a running carry. Integer-exact.
__global__ void __launch_bounds__(128)
position_ids_kernel(const int64_t* __restrict__ ids, int64_t* __restrict__ pos, int B, int S, int64_t pad) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B) return;
    const int64_t* r = ids + static_cast<size_t>(row) * S;
    int64_t* o = pos + static_cast<size_t>(row) * S;
    int carry = 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
        const int s = s0 + lane;
        const int m = (s < S && r[s] != pad) ? 1 : 0;
        int inc = m;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (s < S) o[s] = static_cast<int64_t>((carry + inc) * m) + pad;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// ----------------------------------------------------------------------------------------------- dropout
// Counter-based: element i keeps iff the 16-bit lane (i & 3) of mix64(seed, i >> 2) >= thr16, with thr16 = round(p * 65536);
// kept values are scaled by 65536 / (65536 - thr16) so that E[out] = in exactly. The mask is a pure function of
// (seed, element index): the backward pass recomputes it instead of storing it.
__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
// out = dropout(x) (+ residual); 8 elements per thread per step
__global__ void __launch_bounds__(256)
dropout_kernel(const uint4* __restrict__ x, const uint4* __restrict__ residual, uint4* __restrict__ out, size_t n_vec,
               uint32_t thr16, float keep_scale, uint64_t seed) {
    pdl_prologue();
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_vec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint4 v = ld_nc_v4(x + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t rw[4] = {0, 0, 0, 0};
        if (residual != nullptr) {
            const uint4 r = ld_nc_v4(residual + i);
            rw[0] = r.x, rw[1] = r.y, rw[2] = r.z, rw[3] = r.w;
        }
        const uint64_t h0 = mix64(seed + 0x9e3779b97f4a7c15ull * (2 * i + 1)), h1 = mix64(seed + 0x9e3779b97f4a7c15ull * (2 * i + 2));
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t hh = j < 2 ? h0 : h1;
            const uint32_t ra = static_cast<uint32_t>(hh >> (32 * (j & 1))) & 0xffffu, rb = static_cast<uint32_t>(hh >> (32 * (j & 1) + 16)) & 0xffffu;
            const float2 f = bf2_to_f2(w[j]);
            const float2 r = bf2_to_f2(rw[j]);
            o[j] = f2_to_bf2((ra >= thr16 ? f.x * keep_scale : 0.f) + r.x, (rb >= thr16 ? f.y * keep_scale : 0.f) + r.y);
        }
        st_v4(out + i, make_uint4(o[0], o[1], o[2], o[3]));
    }
}
}  // namespace b200

using namespace b200;

extern "C" int b200_gelu_fwd(const void* x, void* y, size_t n, b200_stream_t stream) {
    B200_REQUIRE(n % 8 == 0 && aligned16(x) && aligned16(y), "gelu_fwd: n must be a multiple of 8 and pointers 16B aligned");
    const size_t nv = n / 8;
    launch_k(gelu_fwd_kernel, dim3(ew_grid(nv, 256)), dim3(256), 0, as_stream(stream), static_cast<const uint4*>(x), static_cast<uint4*>(y), nv);
    return check_launch("gelu_fwd");
}
extern "C" int b200_gelu_bwd(const void* x, const void* dy, void* dx, size_t n, b200_stream_t stream) {
    B200_REQUIRE(n % 8 == 0 && aligned16(x) && aligned16(dy) && aligned16(dx), "gelu_bwd: n must be a multiple of 8 and pointers 16B aligned");
    const size_t nv = n / 8;
    launch_k(gelu_bwd_kernel, dim3(ew_grid(nv, 256)), dim3(256), 0, as_stream(stream), static_cast<const uint4*>(x), static_cast<const uint4*>(dy), static_cast<uint4*>(dx), nv);
    return check_launch("gelu_bwd");
}

extern "C" int b200_rope_qk_inplace(void* qkv, const float* cos_tab, const float* sin_tab, int B, int S, int nh, int hd,
                                    int rot, int inverse, b200_stream_t stream) {
    B200_REQUIRE(rot > 0 && rot % 2 == 0 && rot <= hd, "rope: rot (%d) must be even and <= head_dim (%d)", rot, hd);
    const int half = rot / 2;
    const int T = B * S;
    auto p = static_cast<elem_t*>(qkv);
    if (half % 8 == 0 && hd % 8 == 0 && aligned16(qkv)) {
        const size_t total = static_cast<size_t>(T) * nh * 2 * (half / 8);
        launch_k(rope_vec_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, as_stream(stream), p, cos_tab, sin_tab, T, S, nh, hd, half, inverse);
    } else if (half == 10 && hd % 8 == 0 && aligned16(qkv)) {  // Pythia-2.8b: rot 20 of head_dim 80
        const size_t total = static_cast<size_t>(T) * nh * 2;
        launch_k(rope_row_kernel<10>, dim3(ew_grid(total, 256)), dim3(256), 0, as_stream(stream), p, cos_tab, sin_tab, T, S, nh, hd, inverse);
    } else {
        const size_t total = static_cast<size_t>(T) * nh * 2 * half;
        launch_k(rope_scalar_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, as_stream(stream), p, cos_tab, sin_tab, T, S, nh, hd, half, inverse);
    }
    return check_launch("rope_qk_inplace");
}

static int embedding_fwd_impl(const int64_t* ids0, const void* table0, const int64_t* ids1, const void* table1, const int64_t* ids2,
                              const void* table2, void* out, int T, int h, int vocab0, b200_stream_t stream) {
    B200_REQUIRE(h % 8 == 0 && aligned16(table0) && aligned16(out), "embedding_fwd: h must be a multiple of 8, pointers 16B aligned");
    const int blocks = ew_grid(static_cast<size_t>(T) * 32, 256);
    launch_k(embedding_fwd_kernel, dim3(blocks), dim3(256), 0, as_stream(stream),
        ids0, static_cast<const elem_t*>(table0), ids1, static_cast<const elem_t*>(table1), ids2,
        static_cast<const elem_t*>(table2), static_cast<elem_t*>(out), T, h, vocab0);
    return check_launch("embedding_fwd");
}
extern "C" int b200_embedding3_fwd(const int64_t* ids0, const void* table0, int vocab0, const int64_t* ids1, const void* table1,
                                   const int64_t* ids2, const void* table2, void* out, int T, int h,
                                   b200_stream_t stream) {
    return embedding_fwd_impl(ids0, table0, ids1, table1, ids2, table2, out, T, h, vocab0, stream);
}
extern "C" int b200_embedding_fwd(const int64_t* ids, const void* table, void* out, int T, int h, int vocab,
                                  b200_stream_t stream) {
    return embedding_fwd_impl(ids, table, nullptr, nullptr, nullptr, nullptr, out, T, h, vocab, stream);
}
extern "C" int b200_embedding_bwd(const int64_t* ids, const void* dout, float* dtable, int T, int h, int vocab,
                                  b200_stream_t stream) {
    (void)vocab;
    B200_REQUIRE(h % 8 == 0 && aligned16(dout), "embedding_bwd: h must be a multiple of 8, pointers 16B aligned");
    const int blocks = ew_grid(static_cast<size_t>(T) * 32, 256);
    launch_k(embedding_bwd_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), ids, static_cast<const elem_t*>(dout), dtable, T, h, -1);
    return check_launch("embedding_bwd");
}
extern "C" int b200_embedding_bwd_padding(const int64_t* ids, const void* dout, float* dtable, int T, int h, int64_t padding_idx,
                                          b200_stream_t stream) {
    B200_REQUIRE(h % 8 == 0 && aligned16(dout), "embedding_bwd: h must be a multiple of 8, pointers 16B aligned");
    const int blocks = ew_grid(static_cast<size_t>(T) * 32, 256);
    launch_k(embedding_bwd_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), ids, static_cast<const elem_t*>(dout), dtable, T, h, padding_idx);
    return check_launch("embedding_bwd_padding");
}

extern "C" int b200_cast_f32_to_bf16(const float* src, void* dst, size_t n, b200_stream_t stream) {
    B200_REQUIRE(aligned16(src) && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0, "cast: src must be 16B and dst 8B aligned");
    const size_t n4 = n / 4;
    if (n4) launch_k(cast_f32_bf16_kernel, dim3(ew_grid(n4, 256)), dim3(256), 0, as_stream(stream), reinterpret_cast<const float4*>(src), static_cast<uint2*>(dst), n4);
    if (n % 4) launch_k(cast_f32_bf16_tail, dim3(1), dim3(32), 0, as_stream(stream), src, static_cast<elem_t*>(dst), n4 * 4, n);
    return check_launch("cast_f32_to_bf16");
}
extern "C" int b200_scale_f32(float* x, size_t n, const float* scale_dev, float scale_host, b200_stream_t stream) {
    launch_k(scale_f32_kernel, dim3(ew_grid(n, 256)), dim3(256), 0, as_stream(stream), x, n, scale_dev, scale_host);
    return check_launch("scale_f32");
}

extern "C" size_t b200_colsum_workspace_bytes(int cols) { return static_cast<size_t>(64) * cols * sizeof(float); }
extern "C" int b200_colsum_bf16(const void* x, int rows, int cols, int64_t ld, float* out, float* out2, const float* scale_dev,
                                void* workspace, size_t workspace_bytes, b200_stream_t stream) {
    B200_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0 && aligned16(x), "colsum: cols/ld must be multiples of 8 and x 16B aligned");
    B200_REQUIRE(workspace_bytes >= b200_colsum_workspace_bytes(cols), "colsum: workspace too small");
    const int col_blocks = (cols + 255) / 256;
    int chunks = (num_sms() * 4 + col_blocks - 1) / col_blocks;
    if (chunks > 64) chunks = 64;
    if (chunks > (rows + 63) / 64) chunks = (rows + 63) / 64;
    if (chunks < 1) chunks = 1;
    const int rows_per_chunk = (rows + chunks - 1) / chunks;
    chunks = (rows + rows_per_chunk - 1) / rows_per_chunk;
    float* part = static_cast<float*>(workspace);
    launch_k(colsum_partial_kernel, dim3(col_blocks, chunks), dim3(256), 0, as_stream(stream), static_cast<const elem_t*>(x), rows, cols, ld, rows_per_chunk, part);
    int rc = check_launch("colsum_partial");
    if (rc) return rc;
    launch_k(colsum_finalize_kernel, dim3((cols + 255) / 256), dim3(256), 0, as_stream(stream), part, chunks, cols, out, out2, scale_dev);
    return check_launch("colsum_finalize");
}

extern "C" int b200_roberta_position_ids(const int64_t* ids, int64_t* pos, int B, int S, int64_t pad_id, b200_stream_t stream) {
    B200_REQUIRE(B > 0 && S > 0, "position_ids: bad shape");
    launch_k(position_ids_kernel, dim3((B + 3) / 4), dim3(128), 0, as_stream(stream), ids, pos, B, S, pad_id);
    return check_launch("roberta_position_ids");
}

extern "C" int b200_dropout(const void* x, const void* residual, void* out, size_t n, float p, uint64_t seed, b200_stream_t stream) {
    B200_REQUIRE(n % 8 == 0 && aligned16(x) && aligned16(out) && (!residual || aligned16(residual)), "dropout: n %% 8 == 0 and 16B-aligned pointers");
    B200_REQUIRE(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1)");
    const uint32_t thr = static_cast<uint32_t>(p * 65536.0f + 0.5f);
    const float keep = 65536.0f / static_cast<float>(65536u - thr);
    const size_t n_vec = n / 8;
    launch_k(dropout_kernel, dim3(ew_grid(n_vec, 256)), dim3(256), 0, as_stream(stream), static_cast<const uint4*>(x), static_cast<const uint4*>(residual),
                                                                        static_cast<uint4*>(out), n_vec, thr, keep, seed);
    return check_launch("dropout");
}
