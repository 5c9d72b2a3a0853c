// Flash attention forward / backward on tcgen05 tensor cores (TMEM accumulators, TMA-fed 128B-swizzled smem).
// Replaces F.scaled_dot_product_attention(is_causal=True) (HF:integrations/sdpa_attention.py:92-101, selected by
// src/models/pythia.py:20) and RoBERTa's eager softmax attention (HF:models/roberta/modeling_roberta.py:162-187).
//
// Data layout: q/k/v are read IN PLACE from the packed projection output ([B*S, H, 3, D] for GPT-NeoX) through 2-D TMA
// tensor maps {row_width, B*S}; head h, dims [64c, 64c+64) is the box at column h*head_stride + 64c.  No transposes.
//
// Forward (one CTA per (b, h, 128 query rows)):
//   warp 4: TMA producer (Q once, K/V ring)      warp 5: single-thread tcgen05.mma issuer
//   warps 0-3: softmax; thread r owns query row r = TMEM lane r, so row max / sum need no shuffles.
//   S = Q K^T is double-buffered in TMEM so QK^T of block j+1 overlaps the softmax of block j; O accumulates in TMEM
//   and is rescaled lazily (only when the running max grows by > 2^8, FlashAttention-4 style).
// Backward = delta kernel + two launches of one templated kernel:
//   DKV=false (per 128 query rows, loops over 64-row K/V tiles): S, dP -> dS -> dQ += dS K
//   DKV=true  (per 128 key rows, loops over 64-row Q/dO tiles):  S^T, dP^T -> P^T, dS^T -> dV += P^T dO, dK += dS^T Q
//   For D = 256 the dK/dV accumulators (2 x 256 columns) exceed TMEM next to S^T/dP^T, so the DKV pass is split over two
//   CTAs that each own 128 of the 256 output columns.
#include <math.h>

#include "api.h"
#include "common.cuh"

namespace b200 {

int make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                      uint32_t box_inner, uint32_t box_outer);  // gemm.cu

constexpr float LOG2E_F = 1.4426950408889634f;
constexpr float LN2_F = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct AttnParams {
    int B, S, H, D;
    int causal;
    float scale;
    int64_t qkv_head_stride;
    __nv_bfloat16* o;
    int64_t o_row_stride, o_head_stride;
    float* lse;
    const float* delta;
    __nv_bfloat16* dq;
    __nv_bfloat16* dk;
    __nv_bfloat16* dv;
    int64_t dqkv_row_stride, dqkv_head_stride;
};

// write 8 consecutive bf16 (one 16-byte chunk, index cc along the row) of row r into a [128 x 64*nsub] K-major
// SWIZZLE_128B operand made of 16 KB sub-tiles (one per 64 columns)
__device__ __forceinline__ void st_operand_chunk(uint8_t* tile, int r, int cc, uint4 v) {
    uint8_t* p = tile + (cc >> 3) * 16384 + sw128_offset(r, cc & 7);
    *reinterpret_cast<uint4*>(p) = v;
}

// =================================================================================================================
// forward
// =================================================================================================================
template <int D, int BN, int STAGES>
struct FwdSmem {
    static constexpr int NSUB = D / 64;
    static constexpr uint32_t Q_BYTES = NSUB * 16384;
    static constexpr uint32_t KV_BYTES = NSUB * BN * 128;  // one K (or V) tile
    static constexpr uint32_t P_BYTES = (BN / 64) * 16384;
    static constexpr uint32_t OFF_K = Q_BYTES;
    static constexpr uint32_t OFF_V = OFF_K + STAGES * KV_BYTES;
    static constexpr uint32_t OFF_P = OFF_V + STAGES * KV_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_P + P_BYTES;
    static constexpr uint32_t TOTAL = OFF_BAR + 256 + 1024;
};

template <int D, int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    using L = FwdSmem<D, BN, STAGES>;
    constexpr int NSUB = L::NSUB;
    constexpr uint32_t TM_S = 0, TM_O = 2 * BN;
    static_assert(2 * BN + D <= 512, "TMEM overflow");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + L::OFF_K;
    uint8_t* sV = smem + L::OFF_V;
    uint8_t* sP = smem + L::OFF_P;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* q_full = bars;                   // 1
    uint64_t* k_full = bars + 1;               // STAGES
    uint64_t* v_full = k_full + STAGES;        // STAGES
    uint64_t* kv_empty = v_full + STAGES;      // STAGES
    uint64_t* s_full = kv_empty + STAGES;      // 2
    uint64_t* p_ready = s_full + 2;            // 1
    uint64_t* o_done = p_ready + 1;            // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = gridDim.x - 1 - blockIdx.x;  // heavy (late) causal blocks first
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = qb * 128;
    const int kv_end = p.causal ? min(p.S, q0 + 128) : p.S;
    const int n_blocks = (kv_end + BN - 1) / BN;
    const int col0 = h * static_cast<int>(p.qkv_head_stride);
    const int row_base = b * p.S;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        mbar_init(p_ready, 4);
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0) {
            mbar_expect_tx(q_full, L::Q_BYTES);
            for (int c = 0; c < NSUB; ++c) tma_load_2d(sQ + c * 16384, &tmQ, q_full, col0 + c * 64, row_base + q0);
        }
        for (int j = 0; j < n_blocks; ++j) {
            const int s = j % STAGES;
            mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
            if (lane == 0) {
                mbar_expect_tx(&k_full[s], L::KV_BYTES);
                for (int c = 0; c < NSUB; ++c)
                    tma_load_2d(sK + s * L::KV_BYTES + c * (BN * 128), &tmK, &k_full[s], col0 + c * 64, row_base + j * BN);
                mbar_expect_tx(&v_full[s], L::KV_BYTES);
                for (int c = 0; c < NSUB; ++c)
                    tma_load_2d(sV + s * L::KV_BYTES + c * (BN * 128), &tmV, &v_full[s], col0 + c * 64, row_base + j * BN);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // ---------------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, BN, false, false);  // S = Q K^T: both K-major
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, D, false, true);    // O = P V  : V is MN-major (d contiguous)
        auto issue_s = [&](int j) {
            const int s = j % STAGES;
            mbar_wait(&k_full[s], (j / STAGES) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + s * L::KV_BYTES);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint64_t ad = umma_desc_sw128(qa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
                    const uint64_t bd = umma_desc_sw128(ka + (kk >> 2) * (BN * 128) + (kk & 3) * 32, 16, 1024);
                    umma_ss(tmem + TM_S + (j & 1) * BN, ad, bd, idesc_s, kk != 0);
                }
                tc_commit(&s_full[j & 1]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_s(0);
        for (int j = 0; j < n_blocks; ++j) {
            const int s = j % STAGES;
            if (j + 1 < n_blocks) issue_s(j + 1);
            mbar_wait(p_ready, j & 1);
            mbar_wait(&v_full[s], (j / STAGES) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t pa = smem_u32(sP), va = smem_u32(sV + s * L::KV_BYTES);
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk) {
                    const uint64_t ad = umma_desc_sw128(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
                    const uint64_t bd = umma_desc_sw128(va + kk * 2048, BN * 128, 1024);
                    umma_ss(tmem + TM_O, ad, bd, idesc_o, (j | kk) != 0);
                }
                tc_commit(&kv_empty[s]);
                tc_commit(o_done);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------------- softmax warps: thread = query row
        const int r = warp * 32 + lane;
        const int q_idx = q0 + r;
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        const float sl2 = p.scale * LOG2E_F;
        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < n_blocks; ++j) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            float x[BN];
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(lane_addr + TM_S + (j & 1) * BN + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) x[c * 32 + i] = __uint_as_float(v[i]) * sl2;
            }
            const int kv0 = j * BN;
            const bool need_mask = (kv0 + BN > p.S) || (p.causal && kv0 + BN - 1 > q0);
            if (need_mask) {
#pragma unroll
                for (int i = 0; i < BN; ++i) {
                    const int kv = kv0 + i;
                    if (kv >= p.S || (p.causal && kv > q_idx)) x[i] = -INFINITY;
                }
            }
            float mx = x[0];
#pragma unroll
            for (int i = 1; i < BN; ++i) mx = fmaxf(mx, x[i]);
            const float m_new = fmaxf(m_used, mx);
            const bool need = m_new > m_used + 8.0f;
            if (j > 0) {
                mbar_wait(o_done, (j - 1) & 1);  // PV_{j-1} retired: O is stable and the P tile may be overwritten
                tc_fence_after();
                if (__any_sync(0xffffffffu, need)) {
                    const float alpha = need ? ex2(m_used - m_new) : 1.0f;
                    l *= alpha;
#pragma unroll 1
                    for (int c = 0; c < D / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32(lane_addr + TM_O + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st_32x32(lane_addr + TM_O + c * 32, v);
                    }
                    tmem_st_wait();
                }
            }
            if (need) m_used = m_new;
            const float m_safe = (m_used == -INFINITY) ? 0.f : m_used;
            float sum = 0.f;
#pragma unroll
            for (int cc = 0; cc < BN / 8; ++cc) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    e[i] = ex2(x[cc * 8 + i] - m_safe);
                    sum += e[i];
                }
                st_operand_chunk(sP, r, cc, make_uint4(f2_to_bf2(e[0], e[1]), f2_to_bf2(e[2], e[3]), f2_to_bf2(e[4], e[5]), f2_to_bf2(e[6], e[7])));
            }
            l += sum;
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
        }
        // ---- epilogue: O / l -> bf16, LSE
        mbar_wait(o_done, (n_blocks - 1) & 1);
        tc_fence_after();
        const float inv_l = l > 0.f ? 1.0f / l : 0.f;
        const bool row_ok = q_idx < p.S;
        __nv_bfloat16* orow = p.o + static_cast<size_t>(row_base + q_idx) * p.o_row_stride + static_cast<size_t>(h) * p.o_head_stride;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + TM_O + c * 32, v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 o;
                    o.x = f2_to_bf2(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
                    o.y = f2_to_bf2(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
                    o.z = f2_to_bf2(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
                    o.w = f2_to_bf2(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
                    st_v4(orow + c * 32 + g * 8, o);
                }
            }
        }
        if (row_ok) p.lse[(static_cast<size_t>(b) * p.H + h) * p.S + q_idx] = (m_used + log2f(l)) * LN2_F;
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// =================================================================================================================
// backward
// =================================================================================================================
// delta[b,h,s] = sum_d dO * O ; one warp per (token, head)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, float* __restrict__ delta,
                  int B, int S, int H, int D, int64_t row_stride, int64_t head_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t total = static_cast<int64_t>(B) * S * H;
    for (int64_t w = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); w < total;
         w += static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5)) {
        const int hh = static_cast<int>(w % H);
        const int64_t t = w / H;
        const __nv_bfloat16* op = o + t * row_stride + hh * head_stride;
        const __nv_bfloat16* dp = d_o + t * row_stride + hh * head_stride;
        float s = 0.f;
        for (int c = lane * 8; c < D; c += 256) {
            const uint4 a = ld_nc_v4(op + c), g = ld_nc_v4(dp + c);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 x = bf2_to_f2(aw[i]), y = bf2_to_f2(gw[i]);
                s += x.x * y.x + x.y * y.y;
            }
        }
        s = warp_sum(s);
        if (lane == 0) {
            const int bb = static_cast<int>(t / S), ss = static_cast<int>(t % S);
            delta[(static_cast<size_t>(bb) * H + hh) * S + ss] = s;
        }
    }
}

template <int D, int STAGES, bool DKV>
struct BwdSmem {
    static constexpr int NSUB = D / 64;
    static constexpr uint32_t R_BYTES = NSUB * 16384;      // resident 128-row tile
    static constexpr uint32_t T_BYTES = NSUB * 8192;       // streamed 64-row tile
    static constexpr uint32_t OFF_R2 = R_BYTES;
    static constexpr uint32_t OFF_T1 = 2 * R_BYTES;
    static constexpr uint32_t OFF_T2 = OFF_T1 + STAGES * T_BYTES;
    static constexpr uint32_t OFF_A1 = OFF_T2 + STAGES * T_BYTES;
    static constexpr uint32_t OFF_A2 = OFF_A1 + 16384;
    static constexpr uint32_t OFF_STAT = OFF_A2 + (DKV ? 16384 : 0);
    static constexpr uint32_t OFF_BAR = OFF_STAT + (DKV ? 512 : 0);
    static constexpr uint32_t TOTAL = OFF_BAR + 128 + 1024;
};

// DKV=false: resident R1=Q_i, R2=dO_i ; streamed T1=K_j, T2=V_j ; out dQ (all D columns).
// DKV=true : resident R1=K_j, R2=V_j  ; streamed T1=Q_i, T2=dO_i; out dV, dK columns [half*DH, half*DH+DH).
template <int D, int DH, int STAGES, bool DKV>
__global__ void __launch_bounds__(192, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmR2,
                const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmT2, const AttnParams p) {
    using L = BwdSmem<D, STAGES, DKV>;
    constexpr int NSUB = L::NSUB;
    constexpr int BT = 64;
    constexpr uint32_t TM_S = 0, TM_DP = 64, TM_ACC1 = 128, TM_ACC2 = 128 + DH;
    static_assert(128 + (DKV ? 2 * DH : D) <= 512, "TMEM overflow");
    constexpr int NSPLIT = D / DH;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sR1 = smem;
    uint8_t* sR2 = smem + L::OFF_R2;
    uint8_t* sT1 = smem + L::OFF_T1;
    uint8_t* sT2 = smem + L::OFF_T2;
    uint8_t* sA1 = smem + L::OFF_A1;
    uint8_t* sA2 = smem + L::OFF_A2;
    float* sStat = reinterpret_cast<float*>(smem + L::OFF_STAT);  // DKV: [2][64] lse*log2e, delta of the streamed q tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* r_full = bars;                  // 1
    uint64_t* t_full = bars + 1;              // STAGES
    uint64_t* t_empty = t_full + STAGES;      // STAGES
    uint64_t* s_full = t_empty + STAGES;      // 1
    uint64_t* a_ready = s_full + 1;           // 1
    uint64_t* acc_done = a_ready + 1;         // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = DKV ? (blockIdx.x / NSPLIT) : (gridDim.x - 1 - blockIdx.x);
    const int half = DKV ? (blockIdx.x % NSPLIT) : 0;
    const int h = blockIdx.y, b = blockIdx.z;
    const int r0 = blk * 128;  // first resident row (query row for dQ, key row for dK/dV)
    const int col0 = h * static_cast<int>(p.qkv_head_stride);
    const int col0_do = h * static_cast<int>(p.o_head_stride);  // dO shares O's layout, not the packed qkv layout
    const int row_base = b * p.S;
    // streamed tile range
    int t_begin, t_end;
    if (!DKV) {
        t_begin = 0;
        t_end = ((p.causal ? min(p.S, r0 + 128) : p.S) + BT - 1) / BT;
    } else {
        t_begin = p.causal ? r0 / BT : 0;
        t_end = (p.S + BT - 1) / BT;
    }
    const int n_tiles = t_end - t_begin;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmR1);
        tma_prefetch_desc(&tmR2);
        tma_prefetch_desc(&tmT1);
        tma_prefetch_desc(&tmT2);
        mbar_init(r_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&t_full[s], 1);
            mbar_init(&t_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(a_ready, 4);
        mbar_init(acc_done, 1);
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0) {
            mbar_expect_tx(r_full, 2 * L::R_BYTES);
            for (int c = 0; c < NSUB; ++c) {
                tma_load_2d(sR1 + c * 16384, &tmR1, r_full, col0 + c * 64, row_base + r0);
                tma_load_2d(sR2 + c * 16384, &tmR2, r_full, (DKV ? col0 : col0_do) + c * 64, row_base + r0);
            }
        }
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t % STAGES;
            mbar_wait(&t_empty[s], ((t / STAGES) & 1) ^ 1);
            if (lane == 0) {
                mbar_expect_tx(&t_full[s], 2 * L::T_BYTES);
                const int row = row_base + (t_begin + t) * BT;
                for (int c = 0; c < NSUB; ++c) {
                    tma_load_2d(sT1 + s * L::T_BYTES + c * 8192, &tmT1, &t_full[s], col0 + c * 64, row);
                    tma_load_2d(sT2 + s * L::T_BYTES + c * 8192, &tmT2, &t_full[s], (DKV ? col0_do : col0) + c * 64, row);
                }
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // ---------------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc_s = umma_idesc_bf16(128, BT, false, false);
        constexpr uint32_t idesc_acc = umma_idesc_bf16(128, DKV ? DH : D, false, true);
        auto issue_scores = [&](int t) {
            const int s = t % STAGES;
            mbar_wait(&t_full[s], (t / STAGES) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t r1 = smem_u32(sR1), r2 = smem_u32(sR2);
                const uint32_t t1 = smem_u32(sT1 + s * L::T_BYTES), t2 = smem_u32(sT2 + s * L::T_BYTES);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint64_t ad = umma_desc_sw128(r1 + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
                    const uint64_t bd = umma_desc_sw128(t1 + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024);
                    umma_ss(tmem + TM_S, ad, bd, idesc_s, kk != 0);
                }
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint64_t ad = umma_desc_sw128(r2 + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
                    const uint64_t bd = umma_desc_sw128(t2 + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024);
                    umma_ss(tmem + TM_DP, ad, bd, idesc_s, kk != 0);
                }
                tc_commit(s_full);
            }
            __syncwarp();
        };
        mbar_wait(r_full, 0);
        if (n_tiles > 0) issue_scores(0);
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t % STAGES;
            mbar_wait(a_ready, t & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a1 = smem_u32(sA1), a2 = smem_u32(sA2);
                const uint32_t t1 = smem_u32(sT1 + s * L::T_BYTES), t2 = smem_u32(sT2 + s * L::T_BYTES);
                const uint32_t boff = half * (DH / 64) * 8192;  // first 64-column sub-tile of this CTA's output half
#pragma unroll
                for (int kk = 0; kk < BT / 16; ++kk) {
                    if (!DKV) {
                        // dQ += dS (K-major A) * K_j (MN-major B)
                        const uint64_t ad = umma_desc_sw128(a1 + kk * 32, 16, 1024);
                        const uint64_t bd = umma_desc_sw128(t1 + kk * 2048, 8192, 1024);
                        umma_ss(tmem + TM_ACC1, ad, bd, idesc_acc, (t | kk) != 0);
                    } else {
                        // dV += P^T * dO_i ; dK += dS^T * Q_i
                        const uint64_t ad1 = umma_desc_sw128(a1 + kk * 32, 16, 1024);
                        const uint64_t bd1 = umma_desc_sw128(t2 + boff + kk * 2048, 8192, 1024);
                        umma_ss(tmem + TM_ACC1, ad1, bd1, idesc_acc, (t | kk) != 0);
                        const uint64_t ad2 = umma_desc_sw128(a2 + kk * 32, 16, 1024);
                        const uint64_t bd2 = umma_desc_sw128(t1 + boff + kk * 2048, 8192, 1024);
                        umma_ss(tmem + TM_ACC2, ad2, bd2, idesc_acc, (t | kk) != 0);
                    }
                }
                tc_commit(&t_empty[s]);
                tc_commit(acc_done);
            }
            __syncwarp();
            if (t + 1 < n_tiles) issue_scores(t + 1);
        }
    } else {
        // ---------------------------------------------------------------- compute warps: thread = resident row
        const int r = warp * 32 + lane;
        const int r_idx = r0 + r;
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        const float sl2 = p.scale * LOG2E_F;
        const size_t stat_base = (static_cast<size_t>(b) * p.H + h) * p.S;
        float my_lse2 = 0.f, my_delta = 0.f;
        if (!DKV && r_idx < p.S) {
            my_lse2 = p.lse[stat_base + r_idx] * LOG2E_F;
            my_delta = p.delta[stat_base + r_idx];
        }
        for (int t = 0; t < n_tiles; ++t) {
            const int c0 = (t_begin + t) * BT;  // first streamed row (kv for dQ, q for dK/dV)
            if (DKV) {
                // stage the streamed q tile's statistics; previous iteration's readers are past a_ready -> s_full chain
                named_bar_sync(1, 128);
                if (r < BT) {
                    const int qi = c0 + r;
                    sStat[r] = qi < p.S ? p.lse[stat_base + qi] * LOG2E_F : 0.f;
                    sStat[64 + r] = qi < p.S ? p.delta[stat_base + qi] : 0.f;
                }
                named_bar_sync(1, 128);
            }
            mbar_wait(s_full, t & 1);
            tc_fence_after();
            uint32_t sv[64], dv[64];
            tmem_ld_32x32(lane_addr + TM_S, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
            tmem_ld_32x32(lane_addr + TM_S + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
            tmem_ld_32x32(lane_addr + TM_DP, *reinterpret_cast<uint32_t(*)[32]>(&dv[0]));
            tmem_ld_32x32(lane_addr + TM_DP + 32, *reinterpret_cast<uint32_t(*)[32]>(&dv[32]));
            tmem_ld_wait();
            // s_full of tile t implies every earlier MMA (incl. the accumulate MMAs of tile t-1) retired: A tiles are free
#pragma unroll
            for (int cc = 0; cc < BT / 8; ++cc) {
                float pv[8], ds[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = cc * 8 + i;
                    const int c_idx = c0 + c;
                    bool keep = (c_idx < p.S) && (r_idx < p.S);
                    float lse2, dlt;
                    if (!DKV) {
                        if (p.causal && c_idx > r_idx) keep = false;  // key after query
                        lse2 = my_lse2, dlt = my_delta;
                    } else {
                        if (p.causal && c_idx < r_idx) keep = false;  // query before key
                        lse2 = sStat[c], dlt = sStat[64 + c];
                    }
                    const float pe = keep ? ex2(__uint_as_float(sv[c]) * sl2 - lse2) : 0.f;
                    pv[i] = pe;
                    ds[i] = pe * (__uint_as_float(dv[c]) - dlt);
                }
                const uint4 dsv = make_uint4(f2_to_bf2(ds[0], ds[1]), f2_to_bf2(ds[2], ds[3]), f2_to_bf2(ds[4], ds[5]), f2_to_bf2(ds[6], ds[7]));
                if (!DKV) {
                    st_operand_chunk(sA1, r, cc, dsv);
                } else {
                    st_operand_chunk(sA1, r, cc, make_uint4(f2_to_bf2(pv[0], pv[1]), f2_to_bf2(pv[2], pv[3]), f2_to_bf2(pv[4], pv[5]), f2_to_bf2(pv[6], pv[7])));
                    st_operand_chunk(sA2, r, cc, dsv);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }
        // ---- epilogue
        if (n_tiles > 0) {
            mbar_wait(acc_done, (n_tiles - 1) & 1);
            tc_fence_after();
        }
        const bool row_ok = r_idx < p.S;
        constexpr int NOUT = DKV ? 2 : 1;
#pragma unroll 1
        for (int which = 0; which < NOUT; ++which) {
            __nv_bfloat16* base;
            float mul;
            uint32_t tcol;
            if (!DKV) base = p.dq, mul = p.scale, tcol = TM_ACC1;
            else if (which == 0) base = p.dv, mul = 1.0f, tcol = TM_ACC1;
            else base = p.dk, mul = p.scale, tcol = TM_ACC2;
            __nv_bfloat16* orow = base + static_cast<size_t>(row_base + r_idx) * p.dqkv_row_stride +
                                  static_cast<size_t>(h) * p.dqkv_head_stride + half * DH;
            constexpr int NC = (DKV ? DH : D) / 32;
#pragma unroll 1
            for (int c = 0; c < NC; ++c) {
                uint32_t v[32];
                if (n_tiles > 0) {
                    tmem_ld_32x32(lane_addr + tcol + c * 32, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0;
                }
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 o;
                        o.x = f2_to_bf2(__uint_as_float(v[g * 8 + 0]) * mul, __uint_as_float(v[g * 8 + 1]) * mul);
                        o.y = f2_to_bf2(__uint_as_float(v[g * 8 + 2]) * mul, __uint_as_float(v[g * 8 + 3]) * mul);
                        o.z = f2_to_bf2(__uint_as_float(v[g * 8 + 4]) * mul, __uint_as_float(v[g * 8 + 5]) * mul);
                        o.w = f2_to_bf2(__uint_as_float(v[g * 8 + 6]) * mul, __uint_as_float(v[g * 8 + 7]) * mul);
                        st_v4(orow + c * 32 + g * 8, o);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
static int check_common(const b200_attn_args* a, const char* who) {
    B200_REQUIRE(a != nullptr, "%s: null args", who);
    B200_REQUIRE(a->D == 64 || a->D == 128 || a->D == 256, "%s: head_dim %d unsupported (64, 128, 256)", who, a->D);
    B200_REQUIRE(a->B > 0 && a->S > 0 && a->H > 0, "%s: bad B/S/H", who);
    B200_REQUIRE(a->qkv_row_stride % 8 == 0 && a->qkv_head_stride % 8 == 0, "%s: q/k/v strides must be multiples of 8 elements", who);
    B200_REQUIRE(a->o_row_stride % 8 == 0 && a->o_head_stride % 8 == 0 && aligned16(a->o), "%s: o must be 16B aligned with strides %% 8 == 0", who);
    B200_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v), "%s: q/k/v must be 16B aligned", who);
    return 0;
}

static int qkv_tmap(CUtensorMap* m, const void* ptr, const b200_attn_args* a, int64_t row_stride, int64_t head_stride, uint32_t box_rows) {
    const uint64_t inner = static_cast<uint64_t>(a->H - 1) * head_stride + a->D;
    return make_tmap_bf16_2d(m, ptr, inner, static_cast<uint64_t>(a->B) * a->S, row_stride, 64, box_rows);
}

static AttnParams make_params(const b200_attn_args* a) {
    AttnParams p;
    p.B = a->B, p.S = a->S, p.H = a->H, p.D = a->D;
    p.causal = a->causal;
    p.scale = a->scale;
    p.qkv_head_stride = a->qkv_head_stride;
    p.o = static_cast<__nv_bfloat16*>(a->o);
    p.o_row_stride = a->o_row_stride, p.o_head_stride = a->o_head_stride;
    p.lse = a->lse;
    p.delta = a->delta;
    p.dq = static_cast<__nv_bfloat16*>(a->dq);
    p.dk = static_cast<__nv_bfloat16*>(a->dk);
    p.dv = static_cast<__nv_bfloat16*>(a->dv);
    p.dqkv_row_stride = a->dqkv_row_stride, p.dqkv_head_stride = a->dqkv_head_stride;
    return p;
}

template <typename KernT>
static int set_smem(KernT kern, size_t bytes, const char* who) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) return fail(-2, "%s: cudaFuncSetAttribute(%zu) failed: %s", who, bytes, cudaGetErrorString(e));
    return 0;
}

template <int D, int BN, int STAGES>
static int launch_fwd(const b200_attn_args* a, cudaStream_t st) {
    using L = FwdSmem<D, BN, STAGES>;
    static_assert(L::TOTAL <= 232448, "forward smem budget");
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = qkv_tmap(&tq, a->q, a, a->qkv_row_stride, a->qkv_head_stride, 128))) return rc;
    if ((rc = qkv_tmap(&tk, a->k, a, a->qkv_row_stride, a->qkv_head_stride, BN))) return rc;
    if ((rc = qkv_tmap(&tv, a->v, a, a->qkv_row_stride, a->qkv_head_stride, BN))) return rc;
    auto kern = attn_fwd_kernel<D, BN, STAGES>;
    if ((rc = set_smem(kern, L::TOTAL, "attention_fwd"))) return rc;
    dim3 grid((a->S + 127) / 128, a->H, a->B);
    kern<<<grid, 192, L::TOTAL, st>>>(tq, tk, tv, make_params(a));
    return check_launch("attention_fwd");
}

template <int D, int DH, int STAGES, bool DKV>
static int launch_bwd(const b200_attn_args* a, cudaStream_t st) {
    using L = BwdSmem<D, STAGES, DKV>;
    static_assert(L::TOTAL <= 232448, "backward smem budget");
    CUtensorMap r1, r2, t1, t2;
    int rc;
    // dO shares O's layout: express it as a "qkv-like" map with O's strides
    b200_attn_args oa = *a;
    if (!DKV) {
        if ((rc = qkv_tmap(&r1, a->q, a, a->qkv_row_stride, a->qkv_head_stride, 128))) return rc;
        if ((rc = qkv_tmap(&r2, a->d_o, &oa, a->o_row_stride, a->o_head_stride, 128))) return rc;
        if ((rc = qkv_tmap(&t1, a->k, a, a->qkv_row_stride, a->qkv_head_stride, 64))) return rc;
        if ((rc = qkv_tmap(&t2, a->v, a, a->qkv_row_stride, a->qkv_head_stride, 64))) return rc;
    } else {
        if ((rc = qkv_tmap(&r1, a->k, a, a->qkv_row_stride, a->qkv_head_stride, 128))) return rc;
        if ((rc = qkv_tmap(&r2, a->v, a, a->qkv_row_stride, a->qkv_head_stride, 128))) return rc;
        if ((rc = qkv_tmap(&t1, a->q, a, a->qkv_row_stride, a->qkv_head_stride, 64))) return rc;
        if ((rc = qkv_tmap(&t2, a->d_o, &oa, a->o_row_stride, a->o_head_stride, 64))) return rc;
    }
    auto kern = attn_bwd_kernel<D, DH, STAGES, DKV>;
    if ((rc = set_smem(kern, L::TOTAL, "attention_bwd"))) return rc;
    dim3 grid(((a->S + 127) / 128) * (DKV ? D / DH : 1), a->H, a->B);
    kern<<<grid, 192, L::TOTAL, st>>>(r1, r2, t1, t2, make_params(a));
    return check_launch(DKV ? "attention_bwd_dkv" : "attention_bwd_dq");
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attention_fwd(const b200_attn_args* a, b200_stream_t stream) {
    int rc = check_common(a, "attention_fwd");
    if (rc) return rc;
    B200_REQUIRE(a->lse != nullptr, "attention_fwd: lse is required");
    cudaStream_t st = as_stream(stream);
    switch (a->D) {
        case 64: return launch_fwd<64, 128, 2>(a, st);
        case 128: return launch_fwd<128, 128, 2>(a, st);
        default: return launch_fwd<256, 64, 2>(a, st);
    }
}

extern "C" int b200_attention_bwd(const b200_attn_args* a, b200_stream_t stream) {
    int rc = check_common(a, "attention_bwd");
    if (rc) return rc;
    B200_REQUIRE(a->lse && a->delta && a->d_o && a->dq && a->dk && a->dv, "attention_bwd: lse, delta, d_o, dq, dk, dv are required");
    B200_REQUIRE(a->dqkv_row_stride % 8 == 0 && a->dqkv_head_stride % 8 == 0 && aligned16(a->dq) && aligned16(a->dk) && aligned16(a->dv) && aligned16(a->d_o),
                 "attention_bwd: gradient buffers must be 16B aligned with strides %% 8 == 0");
    cudaStream_t st = as_stream(stream);
    {
        const int64_t total_warps = static_cast<int64_t>(a->B) * a->S * a->H;
        int64_t blocks = (total_warps + 7) / 8;
        const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
        if (blocks > cap) blocks = cap;
        attn_delta_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(a->o), static_cast<const __nv_bfloat16*>(a->d_o),
                                                                    a->delta, a->B, a->S, a->H, a->D, a->o_row_stride, a->o_head_stride);
        if ((rc = check_launch("attention_delta"))) return rc;
    }
    switch (a->D) {
        case 64:
            if ((rc = launch_bwd<64, 64, 2, false>(a, st))) return rc;
            return launch_bwd<64, 64, 2, true>(a, st);
        case 128:
            if ((rc = launch_bwd<128, 128, 2, false>(a, st))) return rc;
            return launch_bwd<128, 128, 2, true>(a, st);
        default:
            if ((rc = launch_bwd<256, 256, 1, false>(a, st))) return rc;
            return launch_bwd<256, 128, 1, true>(a, st);
    }
}
