// Fused LayerNorm forward/backward (HBM-bound).  Replaces nn.LayerNorm as used by GPT-NeoX
// (HF:models/gpt_neox/modeling_gpt_neox.py:251-252,269,279,376) and RoBERTa (HF:models/roberta/modeling_roberta.py).
//
// Layout: x bf16 [rows, cols] row-major; a row is owned by G consecutive warps (G=1 for cols<=2048), each lane keeps
// NV 16-byte vectors (8 bf16) of the row in registers, so x is read from HBM exactly once.  Statistics in fp32,
// two-pass (mean, then centred sum of squares) on the register copy.  Optional second affine (gamma2/beta2 -> y2) for
// GPT-NeoX's parallel block, which normalises the same x twice: one read, two writes.
//
// Algorithmic bytes: fwd 2*cols (read) + 2*cols*NA (write) per row; bwd 2*cols*(1 + NA [+1 dres]) read + 2*cols write.
#include "api.h"
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ float group_sum(float v, int G, float* red) {
    v = warp_sum(v);
    if (G == 1) return v;
    const int warp = threadIdx.x >> 5;
    if (lane_id() == 0) red[warp] = v;
    __syncthreads();
    const int base = (warp / G) * G;
    float t = 0.f;
    for (int j = 0; j < G; ++j) t += red[base + j];
    __syncthreads();
    return t;
}

template <int NV>
__global__ void __launch_bounds__(128)
ln_fwd_kernel(const elem_t* __restrict__ x, const float* __restrict__ g1, const float* __restrict__ b1,
              elem_t* __restrict__ y1, const float* __restrict__ g2, const float* __restrict__ b2,
              elem_t* __restrict__ y2, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
              int cols, float eps, int G) {
    pdl_prologue();
    __shared__ float red[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows_per_block = 4 / G;
    const int sub = warp % G;
    const int row = blockIdx.x * rows_per_block + warp / G;
    const bool row_ok = row < rows;
    const size_t roff = static_cast<size_t>(row) * cols;

    uint4 xv[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int col = ((i * G + sub) * 32 + lane) * 8;
        if (row_ok && col < cols) {
            xv[i] = ld_nc_v4(x + roff + col);
            const float2 a = bf2_to_f2(xv[i].x), b = bf2_to_f2(xv[i].y), c = bf2_to_f2(xv[i].z), d = bf2_to_f2(xv[i].w);
            sum += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
        } else {
            xv[i] = make_uint4(0, 0, 0, 0);
        }
    }
    const float mean = group_sum(sum, G, red) / cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int col = ((i * G + sub) * 32 + lane) * 8;
        if (col < cols) {
            const uint32_t w[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf2_to_f2(w[j]);
                sq += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
            }
        }
    }
    const float var = group_sum(sq, G, red) / cols;
    const float rstd = rsqrtf(var + eps);
    if (row_ok && sub == 0 && lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
    if (!row_ok) return;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int col = ((i * G + sub) * 32 + lane) * 8;
        if (col < cols) {
            const uint32_t w[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
            float xh[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf2_to_f2(w[j]);
                xh[2 * j] = (f.x - mean) * rstd;
                xh[2 * j + 1] = (f.y - mean) * rstd;
            }
            {
                const float4 ga = __ldg(reinterpret_cast<const float4*>(g1 + col));
                const float4 gb = __ldg(reinterpret_cast<const float4*>(g1 + col + 4));
                const float4 ba = __ldg(reinterpret_cast<const float4*>(b1 + col));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + col + 4));
                uint4 o;
                o.x = f2_to_bf2(xh[0] * ga.x + ba.x, xh[1] * ga.y + ba.y);
                o.y = f2_to_bf2(xh[2] * ga.z + ba.z, xh[3] * ga.w + ba.w);
                o.z = f2_to_bf2(xh[4] * gb.x + bb.x, xh[5] * gb.y + bb.y);
                o.w = f2_to_bf2(xh[6] * gb.z + bb.z, xh[7] * gb.w + bb.w);
                st_v4(y1 + roff + col, o);
            }
            if (g2 != nullptr) {
                const float4 ga = __ldg(reinterpret_cast<const float4*>(g2 + col));
                const float4 gb = __ldg(reinterpret_cast<const float4*>(g2 + col + 4));
                const float4 ba = __ldg(reinterpret_cast<const float4*>(b2 + col));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + col + 4));
                uint4 o;
                o.x = f2_to_bf2(xh[0] * ga.x + ba.x, xh[1] * ga.y + ba.y);
                o.y = f2_to_bf2(xh[2] * ga.z + ba.z, xh[3] * ga.w + ba.w);
                o.z = f2_to_bf2(xh[4] * gb.x + bb.x, xh[5] * gb.y + bb.y);
                o.w = f2_to_bf2(xh[6] * gb.z + bb.z, xh[7] * gb.w + bb.w);
                st_v4(y2 + roff + col, o);
            }
        }
    }
}


// Persistent forward: a row is owned by G consecutive warps, a block holds 8 / G row groups and walks rows
// (row = (it * grid + block) * groups + group). Each lane owns the same NV 16-byte column vectors for every row it visits, so the
// affine parameters of those columns are loaded ONCE into registers; re-reading 4 fp32 parameter vectors per 8 activations
// through L1 for every row cost more LSU cycles than the activation traffic itself.
__device__ __forceinline__ float group_sum_bar(float v, int G, float* red, int grp) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5;
    if (lane_id() == 0) red[warp] = v;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(G * 32) : "memory");
    float t = 0.f;
    for (int j = 0; j < G; ++j) t += red[grp * G + j];
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(G * 32) : "memory");
    return t;
}

template <int NV, int NA>
__global__ void __launch_bounds__(256)
ln_fwd_persist_kernel(const elem_t* __restrict__ x, const float* __restrict__ g1, const float* __restrict__ b1,
                      elem_t* __restrict__ y1, const float* __restrict__ g2, const float* __restrict__ b2,
                      elem_t* __restrict__ y2, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
                      int cols, float eps, int G) {
    pdl_prologue();
    __shared__ float red[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = 8 / G;
    const int grp = warp / G, sub = warp % G;
    const float inv_cols = 1.0f / cols;

    float ga[NA][NV * 8], be[NA][NV * 8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int col = ((i * G + sub) * 32 + lane) * 8;
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            const float* gp = a == 0 ? g1 : g2;
            const float* bp = a == 0 ? b1 : b2;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                ga[a][i * 8 + j] = col < cols ? __ldg(gp + col + j) : 0.f;
                be[a][i * 8 + j] = col < cols ? __ldg(bp + col + j) : 0.f;
            }
        }
    }
    const int n_iter = (rows + groups * gridDim.x - 1) / (groups * gridDim.x);
    uint4 raw[NV];  // next row's vectors: issued one iteration ahead so that their latency overlaps the reductions / stores
    auto load_row = [&](int it) {
        const int row = (it * gridDim.x + blockIdx.x) * groups + grp;
        const size_t roff = static_cast<size_t>(row) * cols;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int col = ((i * G + sub) * 32 + lane) * 8;
            raw[i] = (row < rows && col < cols) ? ld_nc_v4(x + roff + col) : make_uint4(0, 0, 0, 0);
        }
    };
    if (n_iter > 0) load_row(0);
    for (int it = 0; it < n_iter; ++it) {
        const int row = (it * gridDim.x + blockIdx.x) * groups + grp;
        const bool row_ok = row < rows;
        const size_t roff = static_cast<size_t>(row) * cols;
        float xf[NV * 8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = bf2_to_f2(w[j]);
                xf[i * 8 + 2 * j] = f.x, xf[i * 8 + 2 * j + 1] = f.y;
                sum += f.x + f.y;
            }
        }
        if (it + 1 < n_iter) load_row(it + 1);
        const float mean = group_sum_bar(sum, G, red, grp) * inv_cols;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int col = ((i * G + sub) * 32 + lane) * 8;
            if (col < cols) {
#pragma unroll
                for (int j = 0; j < 8; ++j) sq += (xf[i * 8 + j] - mean) * (xf[i * 8 + j] - mean);
            }
        }
        const float rstd = rsqrtf(group_sum_bar(sq, G, red, grp) * inv_cols + eps);
        if (row_ok && sub == 0 && lane == 0) {
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
        if (row_ok) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int col = ((i * G + sub) * 32 + lane) * 8;
                if (col < cols) {
#pragma unroll
                    for (int a = 0; a < NA; ++a) {
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = fmaf((xf[i * 8 + j] - mean) * rstd, ga[a][i * 8 + j], be[a][i * 8 + j]);
                        st_v4((a == 0 ? y1 : y2) + roff + col, make_uint4(f2_to_bf2(o[0], o[1]), f2_to_bf2(o[2], o[3]), f2_to_bf2(o[4], o[5]), f2_to_bf2(o[6], o[7])));
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward.  Persistent: one 256-thread block per SM loops over rows; each lane accumulates dgamma/dbeta for the
// columns it owns in registers, block partials go to the workspace, ln_bwd_finalize sums them deterministically and
// accumulates (+=) into the fp32 parameter-gradient buffers.
// ---------------------------------------------------------------------------------------------------------------
template <int NV, int NA>
__global__ void __launch_bounds__(256, 1)
ln_bwd_kernel(const elem_t* __restrict__ x, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const float* __restrict__ g1, const elem_t* __restrict__ dy1, const float* __restrict__ g2,
              const elem_t* __restrict__ dy2, const elem_t* __restrict__ dres,
              elem_t* __restrict__ dx, float* __restrict__ partial, int rows, int cols, int G) {
    pdl_prologue();
    __shared__ float red[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows_per_block = 8 / G;
    const int sub = warp % G;
    const int slot = warp / G;
    const float inv_cols = 1.0f / cols;

    float acc_g[NA][NV * 8], acc_b[NA][NV * 8];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int i = 0; i < NV * 8; ++i) acc_g[a][i] = 0.f, acc_b[a][i] = 0.f;

    // affine weights of this lane's columns: loaded once (they were re-read through L1 for every row)
    float gam[NA][NV * 8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int col = ((i * G + sub) * 32 + lane) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            gam[0][i * 8 + j] = col < cols ? __ldg(g1 + col + j) : 0.f;
            if (NA == 2) gam[NA - 1][i * 8 + j] = col < cols ? __ldg(g2 + col + j) : 0.f;
        }
    }
    const int n_iter = (rows + rows_per_block * gridDim.x - 1) / (rows_per_block * gridDim.x);
    // software pipeline: the loads of row it+1 are issued before the reductions / stores of row it, which doubles the bytes
    // in flight per SM (one block per SM, and each row is only a few KB per operand)
    uint4 rx[NV], r1[NV], r2[NV], rr[NV];
    auto load_row = [&](int it) {
        const int row = (it * gridDim.x + blockIdx.x) * rows_per_block + slot;
        const size_t roff = static_cast<size_t>(row) * cols;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int col = ((i * G + sub) * 32 + lane) * 8;
            const bool ok = row < rows && col < cols;
            rx[i] = r1[i] = r2[i] = rr[i] = make_uint4(0, 0, 0, 0);
            if (ok) {
                rx[i] = ld_nc_v4(x + roff + col);
                r1[i] = ld_nc_v4(dy1 + roff + col);
                if (NA == 2) r2[i] = ld_nc_v4(dy2 + roff + col);
                if (dres != nullptr) rr[i] = ld_nc_v4(dres + roff + col);
            }
        }
    };
    if (n_iter > 0) load_row(0);
    for (int it = 0; it < n_iter; ++it) {
        const int row = (it * gridDim.x + blockIdx.x) * rows_per_block + slot;
        const bool row_ok = row < rows;
        const size_t roff = static_cast<size_t>(row) * cols;
        const float mean = row_ok ? mean_in[row] : 0.f;
        const float rstd = row_ok ? rstd_in[row] : 0.f;

        float xh[NV * 8], dxh[NV * 8], res[NV * 8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int col = ((i * G + sub) * 32 + lane) * 8;
            const bool ok = row_ok && col < cols;
            const uint32_t xw[4] = {rx[i].x, rx[i].y, rx[i].z, rx[i].w};
            const uint32_t w1[4] = {r1[i].x, r1[i].y, r1[i].z, r1[i].w};
            const uint32_t w2[4] = {r2[i].x, r2[i].y, r2[i].z, r2[i].w};
            const uint32_t wr[4] = {rr[i].x, rr[i].y, rr[i].z, rr[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 xf = bf2_to_f2(xw[j]);
                const float2 a = bf2_to_f2(w1[j]);
                const float2 rf = bf2_to_f2(wr[j]);
                const float h0 = ok ? (xf.x - mean) * rstd : 0.f;
                const float h1 = ok ? (xf.y - mean) * rstd : 0.f;
                float t0 = a.x * gam[0][i * 8 + 2 * j], t1 = a.y * gam[0][i * 8 + 2 * j + 1];
                acc_g[0][i * 8 + 2 * j] += a.x * h0;
                acc_g[0][i * 8 + 2 * j + 1] += a.y * h1;
                acc_b[0][i * 8 + 2 * j] += a.x;
                acc_b[0][i * 8 + 2 * j + 1] += a.y;
                if (NA == 2) {
                    const float2 b = bf2_to_f2(w2[j]);
                    t0 += b.x * gam[NA - 1][i * 8 + 2 * j];
                    t1 += b.y * gam[NA - 1][i * 8 + 2 * j + 1];
                    acc_g[NA - 1][i * 8 + 2 * j] += b.x * h0;
                    acc_g[NA - 1][i * 8 + 2 * j + 1] += b.y * h1;
                    acc_b[NA - 1][i * 8 + 2 * j] += b.x;
                    acc_b[NA - 1][i * 8 + 2 * j + 1] += b.y;
                }
                xh[i * 8 + 2 * j] = h0;
                xh[i * 8 + 2 * j + 1] = h1;
                dxh[i * 8 + 2 * j] = t0;
                dxh[i * 8 + 2 * j + 1] = t1;
                res[i * 8 + 2 * j] = rf.x;
                res[i * 8 + 2 * j + 1] = rf.y;
                s1 += t0 + t1;
                s2 += t0 * h0 + t1 * h1;
            }
        }
        if (it + 1 < n_iter) load_row(it + 1);
        s1 = group_sum(s1, G, red) * inv_cols;
        s2 = group_sum(s2, G, red) * inv_cols;
        if (row_ok) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int col = ((i * G + sub) * 32 + lane) * 8;
                if (col < cols) {
                    float r[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = fmaf(rstd, dxh[i * 8 + j] - s1 - xh[i * 8 + j] * s2, res[i * 8 + j]);
                    uint4 o;
                    o.x = f2_to_bf2(r[0], r[1]);
                    o.y = f2_to_bf2(r[2], r[3]);
                    o.z = f2_to_bf2(r[4], r[5]);
                    o.w = f2_to_bf2(r[6], r[7]);
                    st_v4(dx + roff + col, o);
                }
            }
        }
    }
    // partial layout: [gridDim.x * rows_per_block][NA][2][cols]
    float* prow = partial + (static_cast<size_t>(blockIdx.x) * rows_per_block + slot) * (NA * 2) * cols;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int col = ((i * G + sub) * 32 + lane) * 8;
            if (col < cols) {
                float* pg = prow + (a * 2 + 0) * cols + col;
                float* pb = prow + (a * 2 + 1) * cols + col;
                *reinterpret_cast<float4*>(pg) = make_float4(acc_g[a][i * 8], acc_g[a][i * 8 + 1], acc_g[a][i * 8 + 2], acc_g[a][i * 8 + 3]);
                *reinterpret_cast<float4*>(pg + 4) = make_float4(acc_g[a][i * 8 + 4], acc_g[a][i * 8 + 5], acc_g[a][i * 8 + 6], acc_g[a][i * 8 + 7]);
                *reinterpret_cast<float4*>(pb) = make_float4(acc_b[a][i * 8], acc_b[a][i * 8 + 1], acc_b[a][i * 8 + 2], acc_b[a][i * 8 + 3]);
                *reinterpret_cast<float4*>(pb + 4) = make_float4(acc_b[a][i * 8 + 4], acc_b[a][i * 8 + 5], acc_b[a][i * 8 + 6], acc_b[a][i * 8 + 7]);
            }
        }
    }
}

// out[a][k][c] += sum_p partial[p][a][k][c];  block = 32 columns x 32 row-lanes, four independent loads in flight per thread
// (with 8 row-lanes each thread walked ~150 dependent-latency loads: 29 us for 10 MB, 1.3 % of the RoBERTa step). Fixed order:
// deterministic.
__global__ void __launch_bounds__(1024)
ln_bwd_finalize(const float* __restrict__ partial, int n_partial, int n_arrays, int cols, float* o0, float* o1,
                float* o2, float* o3) {
    pdl_prologue();
    __shared__ float sm[32][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int arr = blockIdx.y;
    const int col = blockIdx.x * 32 + cx;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (col < cols) {
        const float* base = partial + static_cast<size_t>(arr) * cols + col;
        const size_t stride = static_cast<size_t>(n_arrays) * cols;
        int p = ry;
        for (; p + 96 < n_partial; p += 128) {
            s0 += base[p * stride];
            s1 += base[(p + 32) * stride];
            s2 += base[(p + 64) * stride];
            s3 += base[(p + 96) * stride];
        }
        for (; p < n_partial; p += 32) s0 += base[p * stride];
    }
    sm[ry][cx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (ry == 0 && col < cols) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) t += sm[j][cx];
        float* out = arr == 0 ? o0 : arr == 1 ? o1 : arr == 2 ? o2 : o3;
        out[col] += t;
    }
}

static int pick_group(int cols, int budget_cols, int max_g) {
    int g = 1;
    while (g < max_g && cols > budget_cols * g) g *= 2;
    return g;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, const float* gamma2,
                                  const float* beta2, void* y2, float* mean, float* rstd, int rows, int cols, float eps,
                                  b200_stream_t stream) {
    B200_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0, "layernorm_fwd: cols (%d) must be a positive multiple of 8", cols);
    B200_REQUIRE(cols <= 8192, "layernorm_fwd: cols %d > 8192 unsupported", cols);
    B200_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), "layernorm_fwd: pointers must be 16-byte aligned");
    B200_REQUIRE((gamma2 == nullptr) == (y2 == nullptr) && (gamma2 == nullptr) == (beta2 == nullptr), "layernorm_fwd: gamma2/beta2/y2 must be all set or all NULL");
    auto xs = static_cast<const elem_t*>(x);
    auto y1s = static_cast<elem_t*>(y);
    auto y2s = static_cast<elem_t*>(y2);
    cudaStream_t st = as_stream(stream);
    // G warps per row: 4 up to 3072 columns (<= 3 vectors per lane keeps the register-resident parameters small), else 8
    const int G = cols <= 3072 ? 4 : 8;
    const int nv = (cols + 256 * G - 1) / (256 * G);
    const int groups = 8 / G;
    int grid = num_sms() * 2;  // two 256-thread blocks per SM are resident (<= 128 registers per thread)
    const int max_blocks = (rows + groups - 1) / groups;
    if (grid > max_blocks) grid = max_blocks;
#define LAUNCH(NVV)                                                                                                                   \
    do {                                                                                                                              \
        if (gamma2) launch_k(ln_fwd_persist_kernel<NVV, 2>, dim3(grid), dim3(256), 0, st, xs, gamma, beta, y1s, gamma2, beta2, y2s, mean, rstd, rows, cols, eps, G); \
        else launch_k(ln_fwd_persist_kernel<NVV, 1>, dim3(grid), dim3(256), 0, st, xs, gamma, beta, y1s, gamma2, beta2, y2s, mean, rstd, rows, cols, eps, G);        \
    } while (0)
    switch (nv) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 3: LAUNCH(3); break;
        case 4: LAUNCH(4); break;
        default: return fail(-1, "layernorm_fwd: unsupported cols %d", cols);
    }
#undef LAUNCH
    return check_launch("layernorm_fwd");
}

extern "C" size_t b200_layernorm_bwd_workspace_bytes(int cols, int n_affine) {
    return static_cast<size_t>(num_sms()) * 8 * n_affine * 2 * cols * sizeof(float);
}

extern "C" int b200_layernorm_bwd(const void* x, const float* mean, const float* rstd, const float* gamma,
                                  const void* dy, const float* gamma2, const void* dy2, const void* dres, void* dx,
                                  float* dgamma, float* dbeta, float* dgamma2, float* dbeta2, void* workspace,
                                  size_t workspace_bytes, int rows, int cols, b200_stream_t stream) {
    B200_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= 8192, "layernorm_bwd: bad cols %d", cols);
    const int NA = gamma2 ? 2 : 1;
    B200_REQUIRE((gamma2 == nullptr) == (dy2 == nullptr), "layernorm_bwd: gamma2/dy2 must both be set or NULL");
    B200_REQUIRE(workspace_bytes >= b200_layernorm_bwd_workspace_bytes(cols, NA), "layernorm_bwd: workspace too small");
    B200_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(workspace) && aligned16(gamma), "layernorm_bwd: pointers must be 16-byte aligned");
    // registers: NV*NA <= 4 vectors per lane (dgamma/dbeta accumulators + xhat/dxhat copies stay under 255 regs)
    const int G = pick_group(cols * NA, 1024, 8);
    const int nv = (cols + 256 * G - 1) / (256 * G);
    const int rows_per_block = 8 / G;
    int grid = num_sms();
    const int max_blocks = (rows + rows_per_block - 1) / rows_per_block;
    if (grid > max_blocks) grid = max_blocks;
    auto xs = static_cast<const elem_t*>(x);
    auto d1 = static_cast<const elem_t*>(dy);
    auto d2 = static_cast<const elem_t*>(dy2);
    auto dr = static_cast<const elem_t*>(dres);
    auto dxs = static_cast<elem_t*>(dx);
    float* part = static_cast<float*>(workspace);
    cudaStream_t st = as_stream(stream);
#define LAUNCH(NVV, NAA) launch_k(ln_bwd_kernel<NVV, NAA>, dim3(grid), dim3(256), 0, st, xs, mean, rstd, gamma, d1, gamma2, d2, dr, dxs, part, rows, cols, G)
    if (NA == 1) {
        switch (nv) {
            case 1: LAUNCH(1, 1); break;
            case 2: LAUNCH(2, 1); break;
            case 3: LAUNCH(3, 1); break;
            case 4: LAUNCH(4, 1); break;  // pick_group keeps nv <= 4 for every cols <= 8192
            default: return fail(-1, "layernorm_bwd: unsupported cols %d", cols);
        }
    } else {
        switch (nv) {
            case 1: LAUNCH(1, 2); break;
            case 2: LAUNCH(2, 2); break;
            case 3: LAUNCH(3, 2); break;
            case 4: LAUNCH(4, 2); break;
            default: return fail(-1, "layernorm_bwd: unsupported cols %d (dual)", cols);
        }
    }
#undef LAUNCH
    int rc = check_launch("layernorm_bwd");
    if (rc) return rc;
    dim3 fgrid((cols + 31) / 32, NA * 2);
    launch_k(ln_bwd_finalize, dim3(fgrid), dim3(1024), 0, st, part, grid * rows_per_block, NA * 2, cols, dgamma, dbeta, dgamma2, dbeta2);
    return check_launch("layernorm_bwd_finalize");
}
