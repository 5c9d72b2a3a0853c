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
#include <stdlib.h>

#include "api.h"
#include "common.cuh"

namespace b200 {

int make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                      uint32_t box_inner, uint32_t box_outer);  // gemm.cu
int make_tmap_bf16_3d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1_elems,
                      uint64_t pitch2_elems, uint32_t box0, uint32_t box2);  // gemm.cu

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

// perf triage (B200_ATTN_TRACE=1): per-event SM clock stamps of CTA (0,0,0) of the head_dim-256 score pass
__device__ unsigned long long g_attn_trace[8192];
__device__ __forceinline__ void trace_evt(bool on, int slot) {
    if (on) g_attn_trace[slot] = clock64();
}

struct AttnParams {
    int B, S, H, D;
    int causal;
    float scale;
    int64_t qkv_head_stride;
    elem_t* o;
    int64_t o_row_stride, o_head_stride;
    float* lse;
    const float* delta;
    elem_t* dq;
    elem_t* dk;
    elem_t* dv;
    int64_t dqkv_row_stride, dqkv_head_stride;
    elem_t* p_out;   // dQ pass only (nullable): bf16 P and dS tiles are also written to [B*H, S, S] scratch so that
    int trace;
    // attention-probability dropout (RoBERTa, HF:models/roberta/modeling_roberta.py:209,236): keep iff hash16 >= drop_thr,
    // kept probabilities scaled by drop_scale = 65536 / (65536 - drop_thr); 0 = off. Generic kernels only.
    uint32_t drop_thr;
    float drop_scale;
    unsigned long long drop_seed;
    const elem_t* d_o;  // score pass: delta = rowsum(dO * O) is computed in its prologue (no separate pre-pass)
    int d_real;  // head_dim as stored (80 for Pythia-2.8b); the kernels run on D = d_real rounded up to 64/128/256 with the
    int pad3d;   // tail columns zero-filled by 3-D TMA maps {d, head, token} (pad3d = 1) and clipped again on store
    elem_t* ds_out;  // dV = P^T dO and dK = dS^T Q can run as batched GEMMs (head_dim 256, see file header)
};

// Dropout mask of score element (q, k) of head bh: 16-bit lane (k & 1) of ONE 32-bit counter hash per (query, key pair):
//   idx = ((bh * S + q) * ceil(S / 2) + (k >> 1)) mod 2^32,   x = lowbias32-style mix of idx with the two halves of the seed.
// Forward and the dQ pass (thread = query row, walks keys) pay one hash per two elements, the dK/dV pass (thread = key row, walks
// queries) one per element. Round 1 used a 64-bit splitmix hash per 2 x 2 block: ~30 integer instructions per hash made the
// dropout forward 3x slower than the dropout-free one (RoBERTa-large: 624 vs ~215 us per layer); this one is 9.
struct DropKey {
    uint32_t s0, s1;
};
__device__ __forceinline__ DropKey attn_drop_key(unsigned long long seed) {
    DropKey k;
    k.s0 = static_cast<uint32_t>(seed) * 0x9E3779B1u + 0x85EBCA6Bu;
    k.s1 = static_cast<uint32_t>(seed >> 32) ^ 0xC2B2AE35u;
    return k;
}
__device__ __forceinline__ uint32_t attn_drop_row(int bh, int S, int q) {  // (bh * S + q) * ceil(S / 2), wrapping
    return (static_cast<uint32_t>(bh) * static_cast<uint32_t>(S) + static_cast<uint32_t>(q)) * static_cast<uint32_t>((S + 1) >> 1);
}
__device__ __forceinline__ uint32_t attn_drop_hash(DropKey key, uint32_t row, int k) {
    uint32_t x = (row + static_cast<uint32_t>(k >> 1)) ^ key.s0;
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x += key.s1;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ bool attn_drop_keep(uint32_t hsh, int k, uint32_t thr) {
    return ((hsh >> (16 * (k & 1))) & 0xffffu) >= thr;
}

// 64-column box c of head h, rows [row, row + box_rows): 2-D map {row width, tokens} or, for padded head dims, 3-D map
// {d, head, token} whose out-of-extent columns arrive as zeros
__device__ __forceinline__ void tma_load_head(void* dst, const CUtensorMap* m, uint64_t* bar, int col0, int c, int h, int row, int pad3d) {
    if (pad3d) tma_load_3d(dst, m, bar, c * 64, h, row);
    else tma_load_2d(dst, m, bar, col0 + c * 64, row);
}

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
    static constexpr uint32_t OFF_BAR = OFF_P + 2 * P_BYTES;  // two P tiles: softmax of block j+1 fills one while P V_j reads the other
    static constexpr uint32_t TOTAL = OFF_BAR + 256 + 1024;
};

// When S (double-buffered) and O need <= 256 TMEM columns (head_dim 64 with 64-key blocks) and the tiles fit in < 113 KB, two
// CTAs are resident per SM and cover each other's prologue / epilogue / barrier round trips.
template <int D, int BN, int STAGES, bool DROP>
__global__ void __launch_bounds__(192, (2 * BN + D <= 256) ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    using L = FwdSmem<D, BN, STAGES>;
    constexpr int NSUB = L::NSUB;
    constexpr uint32_t TM_S = 0, TM_O = 2 * BN;
    static_assert(2 * BN + D <= 512, "TMEM overflow");
    constexpr uint32_t TM_COLS = (2 * BN + D <= 256) ? 256 : 512;

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
    uint64_t* p_ready = s_full + 2;            // 2: one per P buffer (see attn_fwd256_kernel)
    uint64_t* o_done = p_ready + 2;            // 2: o_done[b] = the P V that read P buffer b has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

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
        mbar_init(&p_ready[0], 4);
        mbar_init(&p_ready[1], 4);
        mbar_init(&o_done[0], 1);
        mbar_init(&o_done[1], 1);
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, TM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_prologue();  // set-up (barriers, TMEM) overlapped the previous kernel's tail; global memory only from here on

    if (warp == 4) {
        // ---------------------------------------------------------------- TMA producer (one elected thread)
        if (elect_one()) {
            mbar_expect_tx(q_full, L::Q_BYTES);
            for (int c = 0; c < NSUB; ++c) tma_load_head(sQ + c * 16384, &tmQ, q_full, col0, c, h, row_base + q0, p.pad3d);
            for (int j = 0; j < n_blocks; ++j) {
                const int s = j % STAGES;
                mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
                mbar_expect_tx(&k_full[s], L::KV_BYTES);
                for (int c = 0; c < NSUB; ++c)
                    tma_load_head(sK + s * L::KV_BYTES + c * (BN * 128), &tmK, &k_full[s], col0, c, h, row_base + j * BN, p.pad3d);
                mbar_expect_tx(&v_full[s], L::KV_BYTES);
                for (int c = 0; c < NSUB; ++c)
                    tma_load_head(sV + s * L::KV_BYTES + c * (BN * 128), &tmV, &v_full[s], col0, c, h, row_base + j * BN, p.pad3d);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ---------------------------------------------------------------- MMA issuer (one elected thread: the compiler
        // keeps descriptors in uniform registers and issues UTCHMMA back to back; a lane==0 branch costs ~17 instructions
        // of register->uniform broadcast per MMA, more than a 128x64x16 MMA takes to execute)
        if (elect_one()) {
            constexpr uint32_t idesc_s = umma_idesc_f16(128, BN, false, false);  // S = Q K^T: both K-major
            constexpr uint32_t idesc_o = umma_idesc_f16(128, D, false, true);    // O = P V  : V is MN-major (d contiguous)
            const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ), 16, 1024);
            const uint64_t k_desc0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
            const uint64_t p_desc = umma_desc_sw128(smem_u32(sP), 16, 1024);
            const uint64_t v_desc0 = umma_desc_sw128(smem_u32(sV), BN * 128, 1024);
            auto issue_s = [&](int j) {
                const int s = j % STAGES;
                mbar_wait(&k_full[s], (j / STAGES) & 1);
                tc_fence_after();
                const uint64_t kd = k_desc0 + static_cast<uint64_t>((s * L::KV_BYTES) >> 4);
                const uint32_t d_s = tmem + TM_S + (j & 1) * BN;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint64_t ad = q_desc + static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
                    const uint64_t bd = kd + static_cast<uint64_t>(((kk >> 2) * (BN * 128) + (kk & 3) * 32) >> 4);
                    umma_ss(d_s, ad, bd, idesc_s, kk != 0);
                }
                tc_commit(&s_full[j & 1]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_blocks; ++j) {
                const int s = j % STAGES;
                if (j + 1 < n_blocks) issue_s(j + 1);
                mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
                mbar_wait(&v_full[s], (j / STAGES) & 1);
                tc_fence_after();
                const uint64_t vd = v_desc0 + static_cast<uint64_t>((s * L::KV_BYTES) >> 4);
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk) {
                    const uint64_t ad = p_desc + static_cast<uint64_t>(((j & 1) * L::P_BYTES + (kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
                    const uint64_t bd = vd + static_cast<uint64_t>((kk * 2048) >> 4);
                    umma_ss(tmem + TM_O, ad, bd, idesc_o, (j | kk) != 0);
                }
                tc_commit(&kv_empty[s]);
                tc_commit(&o_done[j & 1]);
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- softmax warps: thread = query row
        const int r = warp * 32 + lane;
        const int q_idx = q0 + r;
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        const float sl2 = p.scale * LOG2E_F;
        const DropKey drop_key = attn_drop_key(p.drop_seed);
        const uint32_t drop_row = attn_drop_row(b * p.H + h, p.S, q_idx);
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
            // P buffer j & 1 was last read by P V_{j-2}
            if (j >= 2) mbar_wait(&o_done[j & 1], ((j >> 1) - 1) & 1);
            if (j > 0 && __any_sync(0xffffffffu, need)) {
                // rescaling O races with an in-flight P V: wait for the latest one (rare: the max must grow by > 2^8)
                mbar_wait(&o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
                {
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
                    sum += e[i];  // the softmax denominator is taken before dropout
                }
                if (DROP) {  // compile-time: the dropout-free instantiation keeps its register allocation and schedule
#pragma unroll
                    for (int i2 = 0; i2 < 4; ++i2) {
                        const int kv = kv0 + cc * 8 + 2 * i2;
                        const uint32_t hsh = attn_drop_hash(drop_key, drop_row, kv);
                        e[2 * i2] = attn_drop_keep(hsh, kv, p.drop_thr) ? e[2 * i2] * p.drop_scale : 0.f;
                        e[2 * i2 + 1] = attn_drop_keep(hsh, kv + 1, p.drop_thr) ? e[2 * i2 + 1] * p.drop_scale : 0.f;
                    }
                }
                st_operand_chunk(sP + (j & 1) * L::P_BYTES, r, cc, make_uint4(f2_to_bf2(e[0], e[1]), f2_to_bf2(e[2], e[3]), f2_to_bf2(e[4], e[5]), f2_to_bf2(e[6], e[7])));
            }
            l += sum;
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[j & 1]);
        }
        // ---- epilogue: O / l -> bf16, LSE
        mbar_wait(&o_done[(n_blocks - 1) & 1], ((n_blocks - 1) >> 1) & 1);
        tc_fence_after();
        const float inv_l = l > 0.f ? 1.0f / l : 0.f;
        const bool row_ok = q_idx < p.S;
        elem_t* orow = p.o + static_cast<size_t>(row_base + q_idx) * p.o_row_stride + static_cast<size_t>(h) * p.o_head_stride;
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
                    if (c * 32 + g * 8 < p.d_real) st_v4(orow + c * 32 + g * 8, o);
                }
            }
        }
        if (row_ok) p.lse[(static_cast<size_t>(b) * p.H + h) * p.S + q_idx] = (m_used + log2f(l)) * LN2_F;
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, TM_COLS);
    }
}

// =================================================================================================================
// forward, head_dim 256
// =================================================================================================================
// Same algorithm as attn_fwd_kernel, re-laid-out for the one head size where shared memory is the constraint:
//   * Q (128 x 256 bf16) lives in TMEM (128 columns) and feeds S = Q K^T as the A operand of a TS-mode MMA: it costs no shared
//     memory and the S MMAs read only the 2 KB K slice per step instead of 6 KB (they were shared-memory-bandwidth bound);
//   * the 64 KB that frees up hold 3 K stages + 3 V stages (32 KB each) instead of 2 shared K/V stages, and a K stage is
//     released as soon as S_j retires (V only after P V_j): with ~2000-clock TMA latency the old ring stalled every tile;
//   * producer and MMA issuer poll their inputs and serve whichever is ready.
// TMEM: O [0,256) | S double buffer [256,384) | Q bf16 [384,512).   smem: P 16 KB | K x3 | V x3 = 208 KB.
struct Fwd256Smem {
    static constexpr uint32_t OFF_P = 0;  // two P tiles: softmax of block j+1 writes one while P V_j reads the other
    static constexpr uint32_t OFF_K = 32768;
    static constexpr uint32_t OFF_V = OFF_K + 3 * 32768;
    static constexpr uint32_t OFF_BAR = OFF_V + 3 * 32768;
    static constexpr uint32_t TOTAL = OFF_BAR + 256 + 1024;
};

__global__ void __launch_bounds__(192, 1)
attn_fwd256_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
    using L = Fwd256Smem;
    constexpr int D = 256, BN = 64, NSUB = 4, NST = 3;
    constexpr uint32_t TM_O = 0, TM_S = 256, TM_Q = 384;
    constexpr uint32_t KV_BYTES = 32768;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sP = smem + L::OFF_P;
    uint8_t* sK = smem + L::OFF_K;
    uint8_t* sV = smem + L::OFF_V;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* q_ready = bars;          // 1 (4 warp arrivals)
    uint64_t* k_full = bars + 1;       // 3
    uint64_t* k_empty = bars + 4;      // 3
    uint64_t* v_full = bars + 7;       // 3
    uint64_t* v_empty = bars + 10;     // 3
    uint64_t* s_full = bars + 13;      // 2
    uint64_t* s_free = bars + 15;      // 2 (4 warp arrivals)
    // p_ready[b]: P buffer b is written (4 warp arrivals). One barrier PER BUFFER: with a single barrier whose phase
    // alternates per block, the softmax warps can complete P_{j+1} (its S was issued before P V_j) while the issuer still
    // waits for a late V_j; the barrier is then two phases ahead and a parity wait for phase j never succeeds again (a
    // deadlock seen once in ~10^6 CTAs). P_{j+2} cannot be written before P V_j retired (o_done), so per buffer the
    // waiter is never more than one phase behind.
    uint64_t* p_ready = bars + 17;     // 2
    uint64_t* o_done = bars + 19;      // 2: o_done[b] = the P V that read P buffer b has retired
    uint64_t* q_full = bars + 21;      // 1: the Q tile has landed in its staging area (V stages 1 and 2)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
    uint8_t* sQst = sV + KV_BYTES;     // 64 KB staging of the Q tile on its way to TMEM

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = gridDim.x - 1 - blockIdx.x;  // heavy (late) causal blocks first
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = qb * 128;
    const int kv_end = p.causal ? min(p.S, q0 + 128) : p.S;
    const int n_blocks = (kv_end + BN - 1) / BN;
    const int col0 = h * static_cast<int>(p.qkv_head_stride);
    const int row_base = b * p.S;
    const bool tr = (p.trace == 1 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ||
                    (p.trace == 2 && blockIdx.x == 0 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1);
    if (tr && threadIdx.x == 0) g_attn_trace[8000] = clock64(), g_attn_trace[8006] = globaltimer_ns();

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(q_ready, 4);
        for (int s = 0; s < NST; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_free[s], 4);
        }
        mbar_init(&p_ready[0], 4);
        mbar_init(&p_ready[1], 4);
        mbar_init(&o_done[0], 1);
        mbar_init(&o_done[1], 1);
        mbar_init(q_full, 1);
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
    pdl_prologue();  // set-up (barriers, TMEM) overlapped the previous kernel's tail; global memory only from here on
    if (tr && threadIdx.x == 0) g_attn_trace[8001] = clock64();

    if (warp == 4) {
        // ---------------------------------------------------------------- TMA producer: two in-order rings, served as they free up
        if (elect_one()) {
            // Q tile -> staging (full 128-byte lines through TMA; per-thread row loads of a [rows x 512 B] tile took 20k clocks)
            mbar_expect_tx(q_full, 65536);
            for (int c = 0; c < NSUB; ++c) tma_load_2d(sQst + c * 16384, &tmQ, q_full, col0 + c * 64, row_base + q0);
            bool q_moved = false;  // V stages 1 and 2 are free once the softmax warps have copied Q into TMEM
            int next_k = 0, next_v = 0;
            const uint64_t t_start = globaltimer_ns();
            uint32_t spins = 0;
            while (next_k < n_blocks || next_v < n_blocks) {
                bool progressed = false;
                if (!q_moved) q_moved = mbar_try_wait(q_ready, 0);
                if (next_k < n_blocks && mbar_try_wait(&k_empty[next_k % NST], ((next_k / NST) & 1) ^ 1)) {
                    const int s = next_k % NST;
                    mbar_expect_tx(&k_full[s], KV_BYTES);
                    for (int c = 0; c < NSUB; ++c) tma_load_2d(sK + s * KV_BYTES + c * 8192, &tmK, &k_full[s], col0 + c * 64, row_base + next_k * BN);
                    ++next_k, progressed = true;
                }
                if (next_v < n_blocks && (q_moved || next_v % NST == 0) && mbar_try_wait(&v_empty[next_v % NST], ((next_v / NST) & 1) ^ 1)) {
                    const int s = next_v % NST;
                    mbar_expect_tx(&v_full[s], KV_BYTES);
                    for (int c = 0; c < NSUB; ++c) tma_load_2d(sV + s * KV_BYTES + c * 8192, &tmV, &v_full[s], col0 + c * 64, row_base + next_v * BN);
                    ++next_v, progressed = true;
                }
                if (!progressed && ((++spins) & 0xfff) == 0 && globaltimer_ns() - t_start > 4 * B200_MBAR_TIMEOUT_NS) {
                    printf("b200pt: attention fwd256 producer stalled (block %d,%d,%d k %d v %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, next_k, next_v);
                    __trap();
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ---------------------------------------------------------------- MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc_s = umma_idesc_f16(128, BN, false, false);  // S = Q K^T (A = Q from TMEM)
            constexpr uint32_t idesc_o = umma_idesc_f16(128, D, false, true);    // O += P V (V MN-major)
            const uint64_t k_desc0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
            const uint64_t v_desc0 = umma_desc_sw128(smem_u32(sV), 8192, 1024);
            const uint64_t p_desc = umma_desc_sw128(smem_u32(sP), 16, 1024);
            auto issue_s = [&](int j) {
                const int s = j % NST;
                const uint64_t kd = k_desc0 + static_cast<uint64_t>((s * KV_BYTES) >> 4);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk)
                    umma_ts(tmem + TM_S + (j & 1) * BN, tmem + TM_Q + kk * 8, kd + static_cast<uint64_t>(((kk >> 2) * 8192 + (kk & 3) * 32) >> 4), idesc_s, kk != 0);
                tc_commit(&k_empty[s]);  // K_j is dead as soon as S_j retires
                tc_commit(&s_full[j & 1]);
            };
            auto issue_pv = [&](int j) {
                const int s = j % NST;
                const uint64_t vd = v_desc0 + static_cast<uint64_t>((s * KV_BYTES) >> 4);
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk)
                    umma_ss(tmem + TM_O, p_desc + static_cast<uint64_t>(((j & 1) * 16384 + kk * 32) >> 4), vd + static_cast<uint64_t>((kk * 2048) >> 4), idesc_o,
                            (j | kk) != 0);
                tc_commit(&v_empty[s]);
                tc_commit(&o_done[j & 1]);
            };
            mbar_wait(q_ready, 0);
            tc_fence_after();
            int next_s = 0, next_pv = 0;  // S_j needs K_j and a free S buffer; P V_j needs P_j and V_j
            const uint64_t t_start = globaltimer_ns();
            uint32_t spins = 0;
            while (next_pv < n_blocks) {
                bool progressed = false;
                if (next_s < n_blocks && next_s <= next_pv + 1 && mbar_try_wait(&k_full[next_s % NST], (next_s / NST) & 1) &&
                    (next_s < 2 || mbar_try_wait(&s_free[next_s & 1], ((next_s >> 1) - 1) & 1))) {
                    tc_fence_after();
                    trace_evt(tr, 4096 + 16 * next_s + 0);
                    issue_s(next_s);
                    ++next_s, progressed = true;
                }
                if (next_pv < next_s && mbar_try_wait(&p_ready[next_pv & 1], (next_pv >> 1) & 1) && mbar_try_wait(&v_full[next_pv % NST], (next_pv / NST) & 1)) {
                    tc_fence_after();
                    trace_evt(tr, 4096 + 16 * next_pv + 1);
                    issue_pv(next_pv);
                    ++next_pv, progressed = true;
                }
                if (!progressed && ((++spins) & 0xfff) == 0 && globaltimer_ns() - t_start > 4 * B200_MBAR_TIMEOUT_NS) {
                    printf("b200pt: attention fwd256 issuer stalled (block %d,%d,%d s %d pv %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, next_s, next_pv);
                    __trap();
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- softmax warps: thread = query row = TMEM lane
        const int r = warp * 32 + lane;
        const int q_idx = q0 + r;
        const bool row_ok = q_idx < p.S;
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        const float sl2 = p.scale * LOG2E_F;
        {
            // Q row: staging tile (SWIZZLE_128B, 4 sub-tiles of 64 columns) -> registers -> TMEM (bf16 pairs, 128 columns)
            mbar_wait(q_full, 0);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 u = ld_shared_v4(sQst + c * 16384 + sw128_offset(r, i));
                    v[i * 4 + 0] = u.x, v[i * 4 + 1] = u.y, v[i * 4 + 2] = u.z, v[i * 4 + 3] = u.w;
                }
                tmem_st_32x32(lane_addr + TM_Q + c * 32, v);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_ready);
            if (tr && threadIdx.x == 0) g_attn_trace[8002] = clock64();
        }
        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < n_blocks; ++j) {
            const bool tr0 = tr && threadIdx.x == 0;
            trace_evt(tr0, 4096 + 16 * j + 8);
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            trace_evt(tr0, 4096 + 16 * j + 9);
            uint32_t sv[64];
            tmem_ld_32x32(lane_addr + TM_S + (j & 1) * BN, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
            tmem_ld_32x32(lane_addr + TM_S + (j & 1) * BN + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[j & 1]);
            const int kv0 = j * BN;
            if ((kv0 + BN > p.S) || (p.causal && kv0 + BN - 1 > q0)) {
                const int lim = p.causal ? min(p.S - 1, q_idx) : p.S - 1;
#pragma unroll
                for (int i = 0; i < 64; ++i) sv[i] = (kv0 + i > lim) ? 0xff800000u : sv[i];
            }
            float mx = __uint_as_float(sv[0]);
#pragma unroll
            for (int i = 1; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
            mx *= sl2;
            const float m_new = fmaxf(m_used, mx);
            const bool need = m_new > m_used + 8.0f;  // lazy rescale: only when the running max grew by > 2^8
            // P buffer j & 1 was last read by P V_{j-2}
            trace_evt(tr0, 4096 + 16 * j + 10);
            if (j >= 2) mbar_wait(&o_done[j & 1], ((j >> 1) - 1) & 1);
            trace_evt(tr0, 4096 + 16 * j + 11);
            if (j > 0 && __any_sync(0xffffffffu, need)) {
                // rescaling O races with an in-flight P V: wait for the latest one (rare: the max must grow by > 2^8)
                mbar_wait(&o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
                {
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
            const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
            float sum = 0.f;
#pragma unroll
            for (int cc = 0; cc < BN / 8; ++cc) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    e[i] = ex2(fmaf(__uint_as_float(sv[cc * 8 + i]), sl2, neg_m));
                    sum += e[i];
                }
                st_shared_v4(sP + (j & 1) * 16384 + sw128_offset(r, cc), make_uint4(f2_to_bf2(e[0], e[1]), f2_to_bf2(e[2], e[3]), f2_to_bf2(e[4], e[5]), f2_to_bf2(e[6], e[7])));
            }
            l += sum;
            trace_evt(tr0, 4096 + 16 * j + 12);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[j & 1]);
            trace_evt(tr0, 4096 + 16 * j + 13);
        }
        // ---- epilogue: O / l -> bf16, LSE
        if (tr && threadIdx.x == 0) g_attn_trace[8003] = clock64();
        mbar_wait(&o_done[(n_blocks - 1) & 1], ((n_blocks - 1) >> 1) & 1);
        tc_fence_after();
        const float inv_l = l > 0.f ? 1.0f / l : 0.f;
        const bool tma_out = (p.S % 128) == 0;  // whole 128-row boxes: stage O in the (idle) K ring and TMA-store full lines
        elem_t* orow = p.o + static_cast<size_t>(row_base + q_idx) * p.o_row_stride + static_cast<size_t>(h) * p.o_head_stride;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + TM_O + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 o;
                o.x = f2_to_bf2(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
                o.y = f2_to_bf2(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
                o.z = f2_to_bf2(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
                o.w = f2_to_bf2(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
                if (tma_out) st_shared_v4(sK + (c >> 1) * 16384 + sw128_offset(r, (c & 1) * 4 + g), o);
                else if (row_ok) st_v4(orow + c * 32 + g * 8, o);
            }
        }
        if (tma_out) {
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (threadIdx.x == 0) {
                for (int c = 0; c < NSUB; ++c) tma_store_2d(&tmO, sK + c * 16384, h * static_cast<int>(p.o_head_stride) + c * 64, row_base + q0);
                tma_store_commit();
                tma_store_wait_read<0>();  // shared memory must outlive the bulk stores READING it; the writes drain on their own
            }
        }
        if (row_ok) p.lse[(static_cast<size_t>(b) * p.H + h) * p.S + q_idx] = (m_used + log2f(l)) * LN2_F;
        tc_fence_before();
        if (tr && threadIdx.x == 0) g_attn_trace[8004] = clock64();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
    if (tr && threadIdx.x == 0) g_attn_trace[8005] = clock64(), g_attn_trace[8007] = globaltimer_ns();
}

// ---- attn_fwd256s_kernel: attn_fwd256_kernel with the softmax of every S tile split over TWO warps per TMEM lane quarter (320
// threads: warps 0-7 softmax, 8 TMA producer, 9 MMA issuer). Same TMEM / shared-memory plan plus 1 KB for the exchange.
struct Fwd256SSmem {
    static constexpr uint32_t OFF_P = 0;  // two P tiles: softmax of block j+1 writes one while P V_j reads the other
    static constexpr uint32_t OFF_K = 32768;
    static constexpr uint32_t OFF_V = OFF_K + 3 * 32768;
    static constexpr uint32_t OFF_BAR = OFF_V + 3 * 32768;
    static constexpr uint32_t OFF_X = OFF_BAR + 256;  // 1 KB: partial row maxima of the two column halves, [block parity][half][128] bf16
    static constexpr uint32_t TOTAL = OFF_X + 1024 + 1024;
};

__global__ void __launch_bounds__(320, 1)
attn_fwd256s_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
    using L = Fwd256SSmem;
    constexpr int D = 256, BN = 64, NSUB = 4, NST = 3;
    constexpr uint32_t TM_O = 0, TM_S = 256, TM_Q = 384;
    constexpr uint32_t KV_BYTES = 32768;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sP = smem + L::OFF_P;
    uint8_t* sK = smem + L::OFF_K;
    uint8_t* sV = smem + L::OFF_V;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* q_ready = bars;          // 1 (8 warp arrivals)
    uint64_t* k_full = bars + 1;       // 3
    uint64_t* k_empty = bars + 4;      // 3
    uint64_t* v_full = bars + 7;       // 3
    uint64_t* v_empty = bars + 10;     // 3
    uint64_t* s_full = bars + 13;      // 2
    uint64_t* s_free = bars + 15;      // 2 (4 warp arrivals)
    // p_ready[b]: P buffer b is written (4 warp arrivals). One barrier PER BUFFER: with a single barrier whose phase
    // alternates per block, the softmax warps can complete P_{j+1} (its S was issued before P V_j) while the issuer still
    // waits for a late V_j; the barrier is then two phases ahead and a parity wait for phase j never succeeds again (a
    // deadlock seen once in ~10^6 CTAs). P_{j+2} cannot be written before P V_j retired (o_done), so per buffer the
    // waiter is never more than one phase behind.
    uint64_t* p_ready = bars + 17;     // 2
    uint64_t* o_done = bars + 19;      // 2: o_done[b] = the P V that read P buffer b has retired
    uint64_t* q_full = bars + 21;      // 1: the Q tile has landed in its staging area (V stages 1 and 2)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
    uint8_t* sQst = sV + KV_BYTES;     // 64 KB staging of the Q tile on its way to TMEM
    __nv_bfloat16* xmax = reinterpret_cast<__nv_bfloat16*>(smem + L::OFF_X);  // [2][2][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = gridDim.x - 1 - blockIdx.x;  // heavy (late) causal blocks first
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = qb * 128;
    const int kv_end = p.causal ? min(p.S, q0 + 128) : p.S;
    const int n_blocks = (kv_end + BN - 1) / BN;
    const int col0 = h * static_cast<int>(p.qkv_head_stride);
    const int row_base = b * p.S;
    const bool tr = (p.trace == 1 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ||
                    (p.trace == 2 && blockIdx.x == 0 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1);
    if (tr && threadIdx.x == 0) g_attn_trace[8000] = clock64(), g_attn_trace[8006] = globaltimer_ns();

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(q_ready, 8);
        for (int s = 0; s < NST; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_free[s], 8);
        }
        mbar_init(&p_ready[0], 8);
        mbar_init(&p_ready[1], 8);
        mbar_init(&o_done[0], 1);
        mbar_init(&o_done[1], 1);
        mbar_init(q_full, 1);
        fence_barrier_init();
    }
    if (warp == 9) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_prologue();  // set-up (barriers, TMEM) overlapped the previous kernel's tail; global memory only from here on
    if (tr && threadIdx.x == 0) g_attn_trace[8001] = clock64();

    if (warp == 8) {
        // ---------------------------------------------------------------- TMA producer: two in-order rings, served as they free up
        if (elect_one()) {
            // Q tile -> staging (full 128-byte lines through TMA; per-thread row loads of a [rows x 512 B] tile took 20k clocks)
            mbar_expect_tx(q_full, 65536);
            for (int c = 0; c < NSUB; ++c) tma_load_2d(sQst + c * 16384, &tmQ, q_full, col0 + c * 64, row_base + q0);
            bool q_moved = false;  // V stages 1 and 2 are free once the softmax warps have copied Q into TMEM
            int next_k = 0, next_v = 0;
            const uint64_t t_start = globaltimer_ns();
            uint32_t spins = 0;
            while (next_k < n_blocks || next_v < n_blocks) {
                bool progressed = false;
                if (!q_moved) q_moved = mbar_try_wait(q_ready, 0);
                if (next_k < n_blocks && mbar_try_wait(&k_empty[next_k % NST], ((next_k / NST) & 1) ^ 1)) {
                    const int s = next_k % NST;
                    mbar_expect_tx(&k_full[s], KV_BYTES);
                    for (int c = 0; c < NSUB; ++c) tma_load_2d(sK + s * KV_BYTES + c * 8192, &tmK, &k_full[s], col0 + c * 64, row_base + next_k * BN);
                    ++next_k, progressed = true;
                }
                if (next_v < n_blocks && (q_moved || next_v % NST == 0) && mbar_try_wait(&v_empty[next_v % NST], ((next_v / NST) & 1) ^ 1)) {
                    const int s = next_v % NST;
                    mbar_expect_tx(&v_full[s], KV_BYTES);
                    for (int c = 0; c < NSUB; ++c) tma_load_2d(sV + s * KV_BYTES + c * 8192, &tmV, &v_full[s], col0 + c * 64, row_base + next_v * BN);
                    ++next_v, progressed = true;
                }
                if (!progressed && ((++spins) & 0xfff) == 0 && globaltimer_ns() - t_start > 4 * B200_MBAR_TIMEOUT_NS) {
                    printf("b200pt: attention fwd256s producer stalled (block %d,%d,%d k %d v %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, next_k, next_v);
                    __trap();
                }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ---------------------------------------------------------------- MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc_s = umma_idesc_f16(128, BN, false, false);  // S = Q K^T (A = Q from TMEM)
            constexpr uint32_t idesc_o = umma_idesc_f16(128, D, false, true);    // O += P V (V MN-major)
            const uint64_t k_desc0 = umma_desc_sw128(smem_u32(sK), 16, 1024);
            const uint64_t v_desc0 = umma_desc_sw128(smem_u32(sV), 8192, 1024);
            const uint64_t p_desc = umma_desc_sw128(smem_u32(sP), 16, 1024);
            auto issue_s = [&](int j) {
                const int s = j % NST;
                const uint64_t kd = k_desc0 + static_cast<uint64_t>((s * KV_BYTES) >> 4);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk)
                    umma_ts(tmem + TM_S + (j & 1) * BN, tmem + TM_Q + kk * 8, kd + static_cast<uint64_t>(((kk >> 2) * 8192 + (kk & 3) * 32) >> 4), idesc_s, kk != 0);
                tc_commit(&k_empty[s]);  // K_j is dead as soon as S_j retires
                tc_commit(&s_full[j & 1]);
            };
            auto issue_pv = [&](int j) {
                const int s = j % NST;
                const uint64_t vd = v_desc0 + static_cast<uint64_t>((s * KV_BYTES) >> 4);
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk)
                    umma_ss(tmem + TM_O, p_desc + static_cast<uint64_t>(((j & 1) * 16384 + kk * 32) >> 4), vd + static_cast<uint64_t>((kk * 2048) >> 4), idesc_o,
                            (j | kk) != 0);
                tc_commit(&v_empty[s]);
                tc_commit(&o_done[j & 1]);
            };
            mbar_wait(q_ready, 0);
            tc_fence_after();
            int next_s = 0, next_pv = 0;  // S_j needs K_j and a free S buffer; P V_j needs P_j and V_j
            const uint64_t t_start = globaltimer_ns();
            uint32_t spins = 0;
            while (next_pv < n_blocks) {
                bool progressed = false;
                if (next_s < n_blocks && next_s <= next_pv + 1 && mbar_try_wait(&k_full[next_s % NST], (next_s / NST) & 1) &&
                    (next_s < 2 || mbar_try_wait(&s_free[next_s & 1], ((next_s >> 1) - 1) & 1))) {
                    tc_fence_after();
                    trace_evt(tr, 4096 + 16 * next_s + 0);
                    issue_s(next_s);
                    ++next_s, progressed = true;
                }
                if (next_pv < next_s && mbar_try_wait(&p_ready[next_pv & 1], (next_pv >> 1) & 1) && mbar_try_wait(&v_full[next_pv % NST], (next_pv / NST) & 1)) {
                    tc_fence_after();
                    trace_evt(tr, 4096 + 16 * next_pv + 1);
                    issue_pv(next_pv);
                    ++next_pv, progressed = true;
                }
                if (!progressed && ((++spins) & 0xfff) == 0 && globaltimer_ns() - t_start > 4 * B200_MBAR_TIMEOUT_NS) {
                    printf("b200pt: attention fwd256s issuer stalled (block %d,%d,%d s %d pv %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, next_s, next_pv);
                    __trap();
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- softmax warps 0..7: TWO threads per query row.
        // Warps w and w + 4 sit on the same SM sub-partition and the same TMEM lane quarter (hardware: lanes 32 * (warp % 4)); warp
        // w / 4 = hf owns key columns [32 hf, 32 hf + 32) of every 64-key block. The single softmax warp per sub-partition of
        // attn_fwd256_kernel ran TMEM load -> max -> 64 exp -> pack -> store as ONE dependency chain of ~1900 clocks per block with
        // nothing to overlap it (measured: a lone CTA takes as long per block as a full grid, profiles/r02_attn_scaling.txt), twice
        // the 1024 tensor clocks it feeds. Two warps halve the chain and overlap each other's TMEM / MUFU / store latencies. The two
        // threads of a row must use the SAME reference maximum: each rounds its partial maximum UP to bf16, they swap the 16-bit
        // values through shared memory (one 64-thread named barrier per block) and both take the larger one — any common reference
        // within 2^8 of the true maximum is exact for softmax (lazy rescaling), so the rounding costs nothing.
        const int quad = warp & 3, hf = warp >> 2;
        const int r = quad * 32 + lane;
        const int q_idx = q0 + r;
        const bool row_ok = q_idx < p.S;
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16);
        const float sl2 = p.scale * LOG2E_F;
        const int pair_bar = 2 + quad;
        {
            // Q row: staging tile (SWIZZLE_128B, 4 sub-tiles of 64 columns) -> registers -> TMEM (16-bit pairs, 128 columns); half each
            mbar_wait(q_full, 0);
#pragma unroll
            for (int cq = 0; cq < 2; ++cq) {
                const int c = hf * 2 + cq;
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 u = ld_shared_v4(sQst + c * 16384 + sw128_offset(r, i));
                    v[i * 4 + 0] = u.x, v[i * 4 + 1] = u.y, v[i * 4 + 2] = u.z, v[i * 4 + 3] = u.w;
                }
                tmem_st_32x32(lane_addr + TM_Q + c * 32, v);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_ready);
            if (tr && threadIdx.x == 0) g_attn_trace[8002] = clock64();
        }
        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < n_blocks; ++j) {
            const bool tr0 = tr && threadIdx.x == 0;
            trace_evt(tr0, 4096 + 16 * j + 8);
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            trace_evt(tr0, 4096 + 16 * j + 9);
            uint32_t sv[32];
            tmem_ld_32x32(lane_addr + TM_S + (j & 1) * BN + hf * 32, sv);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[j & 1]);
            const int kv0 = j * BN + hf * 32;  // first key of this thread's columns
            if ((j * BN + BN > p.S) || (p.causal && j * BN + BN - 1 > q0)) {
                const int lim = p.causal ? min(p.S - 1, q_idx) : p.S - 1;
#pragma unroll
                for (int i = 0; i < 32; ++i) sv[i] = (kv0 + i > lim) ? 0xff800000u : sv[i];
            }
            float mx = __uint_as_float(sv[0]);
#pragma unroll
            for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
            // common row maximum of both halves: bf16 values rounded towards +inf, so both threads compute the identical number
            const __nv_bfloat16 mine = __float2bfloat16_ru(mx * sl2);
            xmax[((j & 1) * 2 + hf) * 128 + r] = mine;
            named_bar_sync(pair_bar, 64);
            mx = fmaxf(__bfloat162float(mine), __bfloat162float(xmax[((j & 1) * 2 + (hf ^ 1)) * 128 + r]));
            const float m_new = fmaxf(m_used, mx);
            const bool need = m_new > m_used + 8.0f;  // lazy rescale: only when the running max grew by > 2^8 (same verdict in both halves)
            // P buffer j & 1 was last read by P V_{j-2}
            trace_evt(tr0, 4096 + 16 * j + 10);
            if (j >= 2) mbar_wait(&o_done[j & 1], ((j >> 1) - 1) & 1);
            trace_evt(tr0, 4096 + 16 * j + 11);
            if (j > 0 && __any_sync(0xffffffffu, need)) {
                // rescaling O races with an in-flight P V: wait for the latest one (rare: the max must grow by > 2^8). Each half
                // rescales its four 32-column chunks of the row's accumulator; P V_j is issued only after all eight warps arrive.
                mbar_wait(&o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
                {
                    const float alpha = need ? ex2(m_used - m_new) : 1.0f;
                    l *= alpha;
#pragma unroll 1
                    for (int c = hf * 4; c < hf * 4 + 4; ++c) {
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
            const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
            float sum = 0.f;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    e[i] = ex2(fmaf(__uint_as_float(sv[cc * 8 + i]), sl2, neg_m));
                    sum += e[i];
                }
                st_shared_v4(sP + (j & 1) * 16384 + sw128_offset(r, hf * 4 + cc), make_uint4(f2_to_bf2(e[0], e[1]), f2_to_bf2(e[2], e[3]), f2_to_bf2(e[4], e[5]), f2_to_bf2(e[6], e[7])));
            }
            l += sum;  // partial denominator of this thread's columns; the halves meet in the epilogue
            trace_evt(tr0, 4096 + 16 * j + 12);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[j & 1]);
            trace_evt(tr0, 4096 + 16 * j + 13);
        }
        // ---- epilogue: O / l -> 16-bit, LSE. Every loop read of xmax is over once the last P V has retired (it needs all eight
        // p_ready arrivals), so the same 1 KB now carries the two partial denominators as fp32.
        if (tr && threadIdx.x == 0) g_attn_trace[8003] = clock64();
        mbar_wait(&o_done[(n_blocks - 1) & 1], ((n_blocks - 1) >> 1) & 1);
        tc_fence_after();
        float* xl = reinterpret_cast<float*>(smem + L::OFF_X);  // [2][128]
        xl[hf * 128 + r] = l;
        named_bar_sync(pair_bar, 64);
        l += xl[(hf ^ 1) * 128 + r];
        const float inv_l = l > 0.f ? 1.0f / l : 0.f;
        const bool tma_out = (p.S % 128) == 0;  // whole 128-row boxes: stage O in the (idle) K ring and TMA-store full lines
        elem_t* orow = p.o + static_cast<size_t>(row_base + q_idx) * p.o_row_stride + static_cast<size_t>(h) * p.o_head_stride;
#pragma unroll 1
        for (int c = hf * 4; c < hf * 4 + 4; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + TM_O + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 o;
                o.x = f2_to_bf2(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
                o.y = f2_to_bf2(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
                o.z = f2_to_bf2(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
                o.w = f2_to_bf2(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
                if (tma_out) st_shared_v4(sK + (c >> 1) * 16384 + sw128_offset(r, (c & 1) * 4 + g), o);
                else if (row_ok) st_v4(orow + c * 32 + g * 8, o);
            }
        }
        if (tma_out) {
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (threadIdx.x == 0) {
                for (int c = 0; c < NSUB; ++c) tma_store_2d(&tmO, sK + c * 16384, h * static_cast<int>(p.o_head_stride) + c * 64, row_base + q0);
                tma_store_commit();
                tma_store_wait_read<0>();  // shared memory must outlive the bulk stores READING it; the writes drain on their own
            }
        }
        if (row_ok && hf == 0) p.lse[(static_cast<size_t>(b) * p.H + h) * p.S + q_idx] = (m_used + log2f(l)) * LN2_F;
        tc_fence_before();
        if (tr && threadIdx.x == 0) g_attn_trace[8004] = clock64();
    }
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
    if (tr && threadIdx.x == 0) g_attn_trace[8005] = clock64(), g_attn_trace[8007] = globaltimer_ns();
}

// =================================================================================================================
// backward
// =================================================================================================================
// delta[b,h,s] = sum_d dO * O ; one warp per (token, head)
// LPR lanes per (token, head) row: 8 for head_dim <= 64, 16 up to 128, 32 beyond — a warp covers 32 / LPR consecutive rows, so every
// lane loads 16 bytes per operand whatever the head size (round 1 gave a whole warp to each row: at head_dim 64 only 8 of 32 lanes
// worked and the pass ran at 1.3 TB/s, 2.6 % of the Pythia-410m micro-batch).
template <int LPR>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const elem_t* __restrict__ o, const elem_t* __restrict__ d_o, float* __restrict__ delta,
                  int B, int S, int H, int D, int64_t row_stride, int64_t head_stride) {
    pdl_prologue();
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR, rsel = lane / LPR;
    const int64_t total = static_cast<int64_t>(B) * S * H;
    const int64_t n_groups = (total + RPW - 1) / RPW;
    for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); g < n_groups;
         g += static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5)) {
        const int64_t w = g * RPW + rsel;
        const bool ok = w < total;
        const int hh = ok ? static_cast<int>(w % H) : 0;
        const int64_t t = ok ? w / H : 0;
        const elem_t* op = o + t * row_stride + hh * head_stride;
        const elem_t* dp = d_o + t * row_stride + hh * head_stride;
        float s = 0.f;
        if (ok) {
            for (int c = sub * 8; c < D; c += LPR * 8) {
                const uint4 a = ld_nc_v4(op + c), g4 = ld_nc_v4(dp + c);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 x = bf2_to_f2(aw[i]), y = bf2_to_f2(gw[i]);
                    s += x.x * y.x + x.y * y.y;
                }
            }
        }
#pragma unroll
        for (int off = LPR / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (ok && sub == 0) {
            const int bb = static_cast<int>(t / S), ss = static_cast<int>(t % S);
            delta[(static_cast<size_t>(bb) * H + hh) * S + ss] = s;
        }
    }
}

template <int D, int STAGES, bool DKV, int SBUF = 2>
struct BwdSmem {
    static constexpr int NSUB = D / 64;
    static constexpr uint32_t R_BYTES = NSUB * 16384;      // resident 128-row tile
    static constexpr uint32_t T_BYTES = NSUB * 8192;       // streamed 64-row tile
    static constexpr uint32_t OFF_R2 = R_BYTES;
    static constexpr uint32_t OFF_T1 = 2 * R_BYTES;
    static constexpr uint32_t OFF_T2 = OFF_T1 + STAGES * T_BYTES;
    static constexpr int ABUF = ((D == 256 && DKV) || SBUF == 1) ? 1 : 2;  // operand tiles are double-buffered where they fit
    static constexpr uint32_t OFF_A1 = OFF_T2 + STAGES * T_BYTES;          // dS (dQ pass) / P^T (dK/dV pass) tiles
    static constexpr uint32_t OFF_A2 = OFF_A1 + ABUF * 16384;              // dS^T tiles (dK/dV pass only)
    static constexpr uint32_t OFF_STAT = OFF_A2 + (DKV ? ABUF * 16384 : 0);  // [2][2][64] fp32 -lse*log2e, delta of the q tile
    static constexpr uint32_t OFF_BAR = OFF_STAT + (DKV ? 1024 : 0);
    static constexpr uint32_t TOTAL = OFF_BAR + 256 + 1024;
};

// DKV=false: resident R1=Q_i, R2=dO_i ; streamed T1=K_j, T2=V_j ; out dQ (all D columns).
// DKV=true : resident R1=K_j, R2=V_j  ; streamed T1=Q_i, T2=dO_i; out dV, dK columns [half*DH, half*DH+DH).
// Pipeline (same scheme as the head_dim-256 score pass): the score accumulators S / dP are double-buffered in TMEM and the
// score MMAs of tile t+1 are issued as soon as their operands have landed and the compute warps have pulled tile t-1 out of
// that buffer, so tensor work of tile t+1 overlaps the exp / multiply work of tile t; the bf16 operand tiles the compute
// warps produce are double-buffered too; the issuer polls and serves whichever MMA group is ready.
// SBUF = 1 (head_dim 64): single score buffer, single operand buffers, 256 TMEM columns and < 113 KB of shared memory, so that
// TWO CTAs are resident per SM: CTAs of this size are short (8-32 tiles) and their un-overlapped prologue / epilogue / barrier
// round trips are covered by the co-resident CTA instead of by deeper buffering inside one CTA.
template <int D, int DH, int STAGES, bool DKV, int SBUF, bool DROP>
__global__ void __launch_bounds__(192, SBUF == 1 ? 2 : 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmR2,
                const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmT2, const AttnParams p) {
    using L = BwdSmem<D, STAGES, DKV, SBUF>;
    constexpr int NSUB = L::NSUB;
    constexpr int BT = 64;
    constexpr uint32_t TM_ACC1 = SBUF * 128, TM_ACC2 = SBUF * 128 + DH;  // score buffers: S at b * 128, dP at b * 128 + 64 (b = tile % SBUF)
    constexpr uint32_t TM_USED = SBUF * 128 + (DKV ? 2 * DH : D);
    constexpr uint32_t TM_COLS = TM_USED <= 256 ? 256 : 512;
    static_assert(TM_USED <= 512, "TMEM overflow");
    constexpr int NSPLIT = D / DH;
    constexpr int ABUF = L::ABUF;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sR1 = smem;
    uint8_t* sR2 = smem + L::OFF_R2;
    uint8_t* sT1 = smem + L::OFF_T1;
    uint8_t* sT2 = smem + L::OFF_T2;
    uint8_t* sA1 = smem + L::OFF_A1;
    uint8_t* sA2 = smem + L::OFF_A2;
    float* sStat = reinterpret_cast<float*>(smem + L::OFF_STAT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* r_full = bars;                  // 1
    uint64_t* t_full = bars + 1;              // STAGES
    uint64_t* t_empty = t_full + STAGES;      // STAGES
    uint64_t* s_full = t_empty + STAGES;      // 2
    uint64_t* s_free = s_full + 2;            // 2 (4 warp arrivals)
    uint64_t* a_ready = s_free + 2;           // 2 (4 warp arrivals)
    uint64_t* acc_done = a_ready + 2;         // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 2);
    static_assert((1 + 2 * STAGES + 8) * 8 + 4 <= 256, "barrier block overflow");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = DKV ? (blockIdx.x / NSPLIT) : (gridDim.x - 1 - blockIdx.x);
    const int half = DKV ? (blockIdx.x % NSPLIT) : 0;
    const int h = blockIdx.y, b = blockIdx.z;
    const int r0 = blk * 128;  // first resident row (query row for dQ, key row for dK/dV)
    const int col0 = h * static_cast<int>(p.qkv_head_stride);
    const int col0_do = h * static_cast<int>(p.o_head_stride);  // dO shares O's layout, not the packed qkv layout
    const int row_base = b * p.S;
    int t_begin, t_end;
    if (!DKV) {
        t_begin = 0;
        t_end = ((p.causal ? min(p.S, r0 + 128) : p.S) + BT - 1) / BT;
    } else {
        t_begin = p.causal ? r0 / BT : 0;
        t_end = (p.S + BT - 1) / BT;
    }
    const int n_tiles = t_end - t_begin;
    const bool tr = p.trace == 3 && DKV && blockIdx.x == gridDim.x / 2 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1 && threadIdx.x == 0;
    if (tr) g_attn_trace[8100] = clock64(), g_attn_trace[8108] = globaltimer_ns();

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
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_free[s], 4);
            mbar_init(&a_ready[s], 4);
            mbar_init(&acc_done[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, TM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_prologue();  // set-up (barriers, TMEM) overlapped the previous kernel's tail; global memory only from here on
    if (tr) g_attn_trace[8101] = clock64();

    if (warp == 4) {
        // ---------------------------------------------------------------- TMA producer (one elected thread)
        if (elect_one()) {
            mbar_expect_tx(r_full, 2 * L::R_BYTES);
            for (int c = 0; c < NSUB; ++c) {
                tma_load_head(sR1 + c * 16384, &tmR1, r_full, col0, c, h, row_base + r0, p.pad3d);
                tma_load_head(sR2 + c * 16384, &tmR2, r_full, DKV ? col0 : col0_do, c, h, row_base + r0, p.pad3d);
            }
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % STAGES;
                mbar_wait(&t_empty[s], ((t / STAGES) & 1) ^ 1);
                mbar_expect_tx(&t_full[s], 2 * L::T_BYTES);
                const int row = row_base + (t_begin + t) * BT;
                for (int c = 0; c < NSUB; ++c) {
                    tma_load_head(sT1 + s * L::T_BYTES + c * 8192, &tmT1, &t_full[s], col0, c, h, row, p.pad3d);
                    tma_load_head(sT2 + s * L::T_BYTES + c * 8192, &tmT2, &t_full[s], DKV ? col0_do : col0, c, h, row, p.pad3d);
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ---------------------------------------------------------------- MMA issuer (one elected thread, polling)
        if (elect_one()) {
            constexpr uint32_t idesc_s = umma_idesc_f16(128, BT, false, false);
            constexpr uint32_t idesc_acc = umma_idesc_f16(128, DKV ? DH : D, false, true);
            const uint64_t r1_desc = umma_desc_sw128(smem_u32(sR1), 16, 1024);
            const uint64_t r2_desc = umma_desc_sw128(smem_u32(sR2), 16, 1024);
            const uint64_t t1k_desc0 = umma_desc_sw128(smem_u32(sT1), 16, 1024);    // streamed tiles as K-major B (scores)
            const uint64_t t2k_desc0 = umma_desc_sw128(smem_u32(sT2), 16, 1024);
            const uint64_t t1m_desc0 = umma_desc_sw128(smem_u32(sT1), 8192, 1024);  // ... and as MN-major B (accumulate)
            const uint64_t t2m_desc0 = umma_desc_sw128(smem_u32(sT2), 8192, 1024);
            const uint64_t a1_desc0 = umma_desc_sw128(smem_u32(sA1), 16, 1024);
            const uint64_t a2_desc0 = umma_desc_sw128(smem_u32(sA2), 16, 1024);
            auto issue_scores = [&](int t) {
                const uint64_t soff = static_cast<uint64_t>(((t % STAGES) * L::T_BYTES) >> 4);
                const uint32_t tb = tmem + (t % SBUF) * 128;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk)
                    umma_ss(tb, r1_desc + static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                            t1k_desc0 + soff + static_cast<uint64_t>(((kk >> 2) * 8192 + (kk & 3) * 32) >> 4), idesc_s, kk != 0);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk)
                    umma_ss(tb + 64, r2_desc + static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                            t2k_desc0 + soff + static_cast<uint64_t>(((kk >> 2) * 8192 + (kk & 3) * 32) >> 4), idesc_s, kk != 0);
                tc_commit(&s_full[t % SBUF]);
            };
            auto issue_acc = [&](int t) {
                const int s = t % STAGES;
                const uint64_t soff = static_cast<uint64_t>((s * L::T_BYTES) >> 4);
                const uint64_t boff = soff + static_cast<uint64_t>((half * (DH / 64) * 8192) >> 4);  // this CTA's output half
                const uint64_t aoff = static_cast<uint64_t>(((t % ABUF) * 16384) >> 4);
#pragma unroll
                for (int kk = 0; kk < BT / 16; ++kk) {
                    if (!DKV) {
                        // dQ += dS (K-major A) * K_j (MN-major B)
                        umma_ss(tmem + TM_ACC1, a1_desc0 + aoff + static_cast<uint64_t>((kk * 32) >> 4),
                                t1m_desc0 + soff + static_cast<uint64_t>((kk * 2048) >> 4), idesc_acc, (t | kk) != 0);
                    } else {
                        // dV += P^T * dO_i ; dK += dS^T * Q_i
                        umma_ss(tmem + TM_ACC1, a1_desc0 + aoff + static_cast<uint64_t>((kk * 32) >> 4),
                                t2m_desc0 + boff + static_cast<uint64_t>((kk * 2048) >> 4), idesc_acc, (t | kk) != 0);
                        umma_ss(tmem + TM_ACC2, a2_desc0 + aoff + static_cast<uint64_t>((kk * 32) >> 4),
                                t1m_desc0 + boff + static_cast<uint64_t>((kk * 2048) >> 4), idesc_acc, (t | kk) != 0);
                    }
                }
                tc_commit(&t_empty[s]);
                tc_commit(&acc_done[t % ABUF]);
            };
            mbar_wait(r_full, 0);
            tc_fence_after();
            int next_sc = 0, next_acc = 0;
            const uint64_t t_start = globaltimer_ns();
            uint32_t spins = 0;
            while (next_acc < n_tiles) {
                bool progressed = false;
                // scores of tile t: operands landed, and (t >= 2) the compute warps have pulled tile t-2 out of this buffer
                if (next_sc < n_tiles && next_sc < next_acc + 2 && mbar_try_wait(&t_full[next_sc % STAGES], (next_sc / STAGES) & 1) &&
                    (next_sc < SBUF || mbar_try_wait(&s_free[next_sc % SBUF], ((next_sc / SBUF) - 1) & 1))) {
                    tc_fence_after();
                    issue_scores(next_sc);
                    ++next_sc, progressed = true;
                }
                if (next_acc < next_sc && mbar_try_wait(&a_ready[next_acc % ABUF], (next_acc / ABUF) & 1)) {
                    tc_fence_after();
                    issue_acc(next_acc);
                    ++next_acc, progressed = true;
                }
                if (!progressed && ((++spins) & 0xfff) == 0 && globaltimer_ns() - t_start > 4 * B200_MBAR_TIMEOUT_NS) {
                    printf("b200pt: attention bwd issuer stalled (block %d,%d,%d sc %d acc %d of %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
                           next_sc, next_acc, n_tiles);
                    __trap();
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- compute warps: thread = resident row
        const int r = warp * 32 + lane;
        const int r_idx = r0 + r;
        const bool row_ok = r_idx < p.S;
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        const float sl2 = p.scale * LOG2E_F;
        const size_t stat_base = (static_cast<size_t>(b) * p.H + h) * p.S;
        // dQ pass: this row's statistics; a row beyond S gets lse = +inf so that every P (and dS) of it is exactly 0
        float neg_lse2 = -INFINITY, my_delta = 0.f;
        if (!DKV && row_ok) {
            neg_lse2 = -p.lse[stat_base + r_idx] * LOG2E_F;
            my_delta = p.delta[stat_base + r_idx];
        }
        // dK/dV pass: statistics (-lse*log2e, delta) of the streamed q tile go through smem buffer t & 1. They are fetched one
        // tile ahead into registers (threads r < 64) so that their global-load latency overlaps the previous tile's math:
        // loading them at the top of each iteration stalled all 128 threads for a full memory round trip per tile.
        float nx_l = -INFINITY, nx_d = 0.f;
        auto fetch_stats = [&](int t) {
            const int qi = (t_begin + t) * BT + r;
            nx_l = qi < p.S ? -p.lse[stat_base + qi] * LOG2E_F : -INFINITY;
            nx_d = qi < p.S ? p.delta[stat_base + qi] : 0.f;
        };
        if (DKV && r < BT && n_tiles > 0) {
            fetch_stats(0);
            sStat[r] = nx_l, sStat[64 + r] = nx_d;
        }
        for (int t = 0; t < n_tiles; ++t) {
            const int c0 = (t_begin + t) * BT;  // first streamed row (kv for dQ, q for dK/dV)
            float* st = sStat + (t & 1) * 128;
            if (DKV) {
                named_bar_sync(1, 128);  // buffer t & 1 (written at the end of the previous iteration) is visible
                if (r < BT && t + 1 < n_tiles) fetch_stats(t + 1);
            }
            mbar_wait(&s_full[t % SBUF], (t / SBUF) & 1);
            tc_fence_after();
            if (tr && t == 0) g_attn_trace[8102] = clock64();
            if (tr && t == 1) g_attn_trace[8103] = clock64();
            const uint32_t tb = lane_addr + (t % SBUF) * 128;
            // masking is branch-free inside the tile (score -> -inf => P = dS = 0) and skipped for interior tiles
            bool need_mask;
            if (!DKV) need_mask = (c0 + BT > p.S) || (p.causal && c0 + BT - 1 > r0);
            else need_mask = (c0 + BT > p.S) || !row_ok || (p.causal && c0 < r0 + 127);
            uint8_t* a1 = sA1 + (t % ABUF) * 16384;
            uint8_t* a2 = sA2 + (t % ABUF) * 16384;
            // two CTAs per SM (SBUF == 1) leave 168 registers per thread: the tile is then processed in two 32-column halves
            constexpr int NHALF = SBUF == 1 ? 2 : 1;
            constexpr int HC = 64 / NHALF;  // score columns per pass
#pragma unroll
            for (int hh = 0; hh < NHALF; ++hh) {
                uint32_t sv[HC], dv[HC];
#pragma unroll
                for (int c = 0; c < HC / 32; ++c) {
                    tmem_ld_32x32(tb + hh * HC + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
                    tmem_ld_32x32(tb + 64 + hh * HC + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&dv[c * 32]));
                }
                tmem_ld_wait();
                if (hh == NHALF - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_free[t % SBUF]);  // both halves are in registers: the buffer may be overwritten
                }
                const int ch = c0 + hh * HC;  // streamed index of this pass's first column
                if (need_mask) {
                    if (!DKV) {
                        const int lim = p.causal ? min(p.S - 1, r_idx) : p.S - 1;  // last visible key of this query row
#pragma unroll
                        for (int i = 0; i < HC; ++i) sv[i] = (ch + i > lim) ? 0xff800000u : sv[i];
                    } else {
                        const int first = (p.causal ? r_idx : 0);                  // first query that sees this key row
#pragma unroll
                        for (int i = 0; i < HC; ++i) sv[i] = (!row_ok || ch + i < first || ch + i >= p.S) ? 0xff800000u : sv[i];
                    }
                }
                uint32_t pk[HC / 2], dk[HC / 2];
#pragma unroll
                for (int q4 = 0; q4 < HC / 4; ++q4) {  // 4 score columns per step; the dK/dV pass reads their statistics as float4
                    float4 l4 = make_float4(neg_lse2, neg_lse2, neg_lse2, neg_lse2), d4 = make_float4(my_delta, my_delta, my_delta, my_delta);
                    if (DKV) {
                        l4 = reinterpret_cast<const float4*>(st)[hh * (HC / 4) + q4];
                        d4 = reinterpret_cast<const float4*>(st + 64)[hh * (HC / 4) + q4];
                    }
                    const float pe0 = ex2(fmaf(__uint_as_float(sv[4 * q4 + 0]), sl2, l4.x));
                    const float pe1 = ex2(fmaf(__uint_as_float(sv[4 * q4 + 1]), sl2, l4.y));
                    const float pe2 = ex2(fmaf(__uint_as_float(sv[4 * q4 + 2]), sl2, l4.z));
                    const float pe3 = ex2(fmaf(__uint_as_float(sv[4 * q4 + 3]), sl2, l4.w));
                    float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;  // dropout mask / keep probability
                    if (DROP) {
                        const int cq = ch + 4 * q4;  // streamed index of the first of the 4 elements (a multiple of 4)
                        const int bh = b * p.H + h;
                        const DropKey dkey = attn_drop_key(p.drop_seed);
                        // (query, key) of element e: dQ pass (r_idx, cq + e); dK/dV pass (cq + e, r_idx)
                        if (!DKV) {  // one hash per key pair of this thread's query row
                            const uint32_t row = attn_drop_row(bh, p.S, r_idx);
                            const uint32_t ha = attn_drop_hash(dkey, row, cq), hb = attn_drop_hash(dkey, row, cq + 2);
                            m0 = attn_drop_keep(ha, cq, p.drop_thr) ? p.drop_scale : 0.f;
                            m1 = attn_drop_keep(ha, cq + 1, p.drop_thr) ? p.drop_scale : 0.f;
                            m2 = attn_drop_keep(hb, cq + 2, p.drop_thr) ? p.drop_scale : 0.f;
                            m3 = attn_drop_keep(hb, cq + 3, p.drop_thr) ? p.drop_scale : 0.f;
                        } else {  // this thread's key against four consecutive query rows: one hash each
                            const uint32_t row0 = attn_drop_row(bh, p.S, cq), step = static_cast<uint32_t>((p.S + 1) >> 1);
                            m0 = attn_drop_keep(attn_drop_hash(dkey, row0, r_idx), r_idx, p.drop_thr) ? p.drop_scale : 0.f;
                            m1 = attn_drop_keep(attn_drop_hash(dkey, row0 + step, r_idx), r_idx, p.drop_thr) ? p.drop_scale : 0.f;
                            m2 = attn_drop_keep(attn_drop_hash(dkey, row0 + 2 * step, r_idx), r_idx, p.drop_thr) ? p.drop_scale : 0.f;
                            m3 = attn_drop_keep(attn_drop_hash(dkey, row0 + 3 * step, r_idx), r_idx, p.drop_thr) ? p.drop_scale : 0.f;
                        }
                    }
                    // P^T operand of dV carries the mask; dS = P * (mask * dP - delta)
                    pk[2 * q4] = f2_to_bf2(pe0 * m0, pe1 * m1);
                    pk[2 * q4 + 1] = f2_to_bf2(pe2 * m2, pe3 * m3);
                    dk[2 * q4] = f2_to_bf2(pe0 * (__uint_as_float(dv[4 * q4 + 0]) * m0 - d4.x), pe1 * (__uint_as_float(dv[4 * q4 + 1]) * m1 - d4.y));
                    dk[2 * q4 + 1] = f2_to_bf2(pe2 * (__uint_as_float(dv[4 * q4 + 2]) * m2 - d4.z), pe3 * (__uint_as_float(dv[4 * q4 + 3]) * m3 - d4.w));
                }
                // operand buffer t % ABUF was last read by the accumulate MMAs of tile t - ABUF
                if (hh == 0 && t >= ABUF) mbar_wait(&acc_done[t % ABUF], ((t / ABUF) - 1) & 1);
#pragma unroll
                for (int c4 = 0; c4 < HC / 8; ++c4) {
                    const int cc = hh * (HC / 8) + c4;
                    const uint4 pvec = make_uint4(pk[c4 * 4], pk[c4 * 4 + 1], pk[c4 * 4 + 2], pk[c4 * 4 + 3]);
                    const uint4 dvec = make_uint4(dk[c4 * 4], dk[c4 * 4 + 1], dk[c4 * 4 + 2], dk[c4 * 4 + 3]);
                    if (!DKV) {
                        st_shared_v4(a1 + sw128_offset(r, cc), dvec);
                        if (p.p_out != nullptr && row_ok) {
                            const size_t off = (stat_base + r_idx) * static_cast<size_t>(p.S) + c0 + cc * 8;
                            st_v4(p.p_out + off, pvec);
                            st_v4(p.ds_out + off, dvec);
                        }
                    } else {
                        st_shared_v4(a1 + sw128_offset(r, cc), pvec);
                        st_shared_v4(a2 + sw128_offset(r, cc), dvec);
                    }
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[t % ABUF]);
            if (DKV && r < BT && t + 1 < n_tiles) {
                // buffer (t+1) & 1 was last read in iteration t-1; everyone has passed this iteration's barrier since
                float* nst = sStat + ((t + 1) & 1) * 128;
                nst[r] = nx_l, nst[64 + r] = nx_d;
            }
        }
        if (!DKV && p.p_out != nullptr && p.causal && (blk & 1) == 0 && r_idx < p.S && r0 + 128 < p.S) {
            // (score-scratch mode of this generic kernel, kept for triage: see attn_bwd_dq256_kernel)
            const size_t off = (stat_base + r_idx) * static_cast<size_t>(p.S) + r0 + 128;
            const int n = min(128, p.S - (r0 + 128));
            for (int c = 0; c < n; c += 8) {
                st_v4(p.p_out + off + c, make_uint4(0, 0, 0, 0));
                st_v4(p.ds_out + off + c, make_uint4(0, 0, 0, 0));
            }
        }
        // ---- epilogue
        if (tr) g_attn_trace[8104] = clock64();
        if (n_tiles > 0) {
            mbar_wait(&acc_done[(n_tiles - 1) % ABUF], ((n_tiles - 1) / ABUF) & 1);
            tc_fence_after();
        }
        if (tr) g_attn_trace[8105] = clock64();
        constexpr int NOUT = DKV ? 2 : 1;
#pragma unroll 1
        for (int which = 0; which < NOUT; ++which) {
            elem_t* base;
            float mul;
            uint32_t tcol;
            if (!DKV) base = p.dq, mul = p.scale, tcol = TM_ACC1;
            else if (which == 0) base = p.dv, mul = 1.0f, tcol = TM_ACC1;
            else base = p.dk, mul = p.scale, tcol = TM_ACC2;
            elem_t* orow = base + static_cast<size_t>(row_base + r_idx) * p.dqkv_row_stride +
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
                        if (half * DH + c * 32 + g * 8 < p.d_real) st_v4(orow + c * 32 + g * 8, o);
                    }
                }
            }
        }
        tc_fence_before();
        if (tr) g_attn_trace[8106] = clock64();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem, TM_COLS);
    }
    if (tr) g_attn_trace[8107] = clock64(), g_attn_trace[8109] = globaltimer_ns();
}

// =================================================================================================================
// backward, head_dim 256: score pass
// =================================================================================================================
// Per CTA: 128 query rows of one (b, h); loops over 64-key tiles. Per tile:
//   S = Q K^T (Q bf16 in TMEM as the A operand: no smem, no smem bandwidth), dP = dO V^T, P = exp2(S*c - lse), dS = P (dP - delta)
//   -> bf16 P and dS tiles: TMA-stored to the [B*H, S, S] scratch for the batched dV / dK GEMMs; dS is also the A operand of
//   dQ += dS K (fp32 in TMEM, 256 columns).
// Shared memory (224 KB): dO resident 64 KB | K x2 64 KB | V x2 64 KB | dS 16 KB | P 16 KB. Keeping Q out of shared memory
// is what makes room for double-buffered K/V: with one stage the TMA latency of every tile sat on the critical path.
// TMEM (512 columns): dQ [0,256) | S [256,320) | dP [320,384) | Q bf16 [384,512).
// Pipeline: scores of tile t+1 are issued as soon as the compute warps have pulled tile t into registers, so tensor work
// of tile t+1 overlaps the exp/mul work of tile t; dQ of tile t follows when its dS tile is in shared memory.
struct Dq256Smem {
    static constexpr uint32_t OFF_DO = 0;
    static constexpr uint32_t OFF_K = 65536;
    static constexpr uint32_t OFF_V = OFF_K + 2 * 32768;
    static constexpr uint32_t OFF_DS = OFF_V + 2 * 32768;
    static constexpr uint32_t OFF_P = OFF_DS + 16384;
    static constexpr uint32_t OFF_BAR = OFF_P + 16384;
    static constexpr uint32_t TOTAL = OFF_BAR + 256 + 1024;
};

__global__ void __launch_bounds__(192, 1)
attn_bwd_dq256_kernel(const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP,
                      const __grid_constant__ CUtensorMap tmDS, const __grid_constant__ CUtensorMap tmDQ,
                      const __grid_constant__ CUtensorMap tmQ, const AttnParams p) {
    using L = Dq256Smem;
    constexpr int D = 256, BT = 64, NSUB = 4;
    constexpr uint32_t TM_DQ = 0, TM_S = 256, TM_DP = 320, TM_Q = 384;
    constexpr uint32_t KV_BYTES = 32768;  // one 64-row K (or V) tile: 4 sub-tiles of 64 rows x 128 B

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sDO = smem + L::OFF_DO;
    uint8_t* sK = smem + L::OFF_K;
    uint8_t* sV = smem + L::OFF_V;
    uint8_t* sDS = smem + L::OFF_DS;
    uint8_t* sP = smem + L::OFF_P;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* do_full = bars;        // 1
    uint64_t* q_ready = bars + 1;    // 1 (4 warp arrivals)
    uint64_t* slot_full = bars + 2;   // 4: the K / V ring (see kslot / vslot)
    uint64_t* slot_empty = bars + 6;  // 4
    uint64_t* s_full = bars + 10;    // 1
    uint64_t* s_free = bars + 11;    // 1 (4 warp arrivals)
    uint64_t* a_ready = bars + 12;   // 1
    uint64_t* dq_done = bars + 13;   // 1
    uint64_t* q_full = bars + 14;    // 1: the Q tile has landed in its staging area (ring slots 2 and 3)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = gridDim.x - 1 - blockIdx.x;  // heavy (late) causal blocks first
    const int h = blockIdx.y, b = blockIdx.z;
    const int r0 = blk * 128;
    const int col0 = h * static_cast<int>(p.qkv_head_stride);
    const int col0_do = h * static_cast<int>(p.o_head_stride);
    const int row_base = b * p.S;
    const int n_tiles = ((p.causal ? min(p.S, r0 + 128) : p.S) + BT - 1) / BT;
    const int zrow = (b * p.H + h) * p.S;  // first row of this head's [S, S] score matrix in the scratch
    const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmDO);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmP);
        tma_prefetch_desc(&tmDS);
        tma_prefetch_desc(&tmDQ);
        mbar_init(do_full, 1);
        mbar_init(q_ready, 4);
        for (int s = 0; s < 4; ++s) {
            mbar_init(&slot_full[s], 1);
            mbar_init(&slot_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 4);
        mbar_init(a_ready, 1);
        mbar_init(dq_done, 1);
        mbar_init(q_full, 1);
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
    pdl_prologue();  // set-up (barriers, TMEM) overlapped the previous kernel's tail; global memory only from here on
    // Four 32 KB slots hold the K and V tiles. A V tile is dead once dP of its tile is done (early), a K tile only after
    // dQ (late); a TMA load takes ~2000 clocks. So tile t+2's K goes into the slot V_t just left (2 tiles of lead) and its
    // V into K_t's slot: each slot is used once every two tiles, alternating roles.
    auto kslot = [](int t) { return ((t & 1) << 1) | ((t >> 1) & 1); };  // 0,2,1,3
    auto vslot = [](int t) { return (((t & 1) << 1) | ((t >> 1) & 1)) ^ 1; };  // 1,3,0,2
    uint8_t* sKV = sK;  // slots are contiguous: sV == sK + 2 * KV_BYTES

    if (warp == 4) {
        // ---------------------------------------------------------------- TMA producer
        if (elect_one()) {
            // Q tile -> ring slots 2, 3 (tile 1's K / V slots) on its way to TMEM: full lines through TMA instead of per-thread rows
            mbar_expect_tx(q_full, 65536);
            for (int c = 0; c < NSUB; ++c) tma_load_2d(sKV + 2 * KV_BYTES + c * 16384, &tmQ, q_full, col0 + c * 64, row_base + r0);
            bool q_moved = false;
            mbar_expect_tx(do_full, 65536);
            for (int c = 0; c < NSUB; ++c) tma_load_2d(sDO + c * 16384, &tmDO, do_full, col0_do + c * 64, row_base + r0);
            // K and V tiles are loaded in tile order each, but whichever of the two next slots comes free first is served first
            // (a parity wait cannot tell "use u-1 not even issued" from "use u-1 released": a slot's uses are tried strictly in
            // order, tracked by one 8-bit counter per slot)
            int next_k = 0, next_v = 0;
            uint32_t uses = 0;
            const uint64_t t_start = globaltimer_ns();
            uint32_t spins = 0;
            while (next_k < n_tiles || next_v < n_tiles) {
                bool progressed = false;
                if (!q_moved) q_moved = mbar_try_wait(q_ready, 0);
#pragma unroll
                for (int is_k = 1; is_k >= 0; --is_k) {
                    const int t = is_k ? next_k : next_v;
                    if (t >= n_tiles) continue;
                    const int slot = is_k ? kslot(t) : vslot(t);
                    if (slot >= 2 && !q_moved) continue;  // still holds the Q tile
                    if (((uses >> (8 * slot)) & 0xff) != static_cast<uint32_t>((t >> 1) & 0xff)) continue;
                    if (!mbar_try_wait(&slot_empty[slot], ((t >> 1) & 1) ^ 1)) continue;
                    uses += 1u << (8 * slot);
                    mbar_expect_tx(&slot_full[slot], KV_BYTES);
                    for (int c = 0; c < NSUB; ++c)
                        tma_load_2d(sKV + slot * KV_BYTES + c * 8192, is_k ? &tmK : &tmV, &slot_full[slot], col0 + c * 64, row_base + t * BT);
                    if (is_k) ++next_k; else ++next_v;
                    progressed = true;
                }
                if (!progressed && ((++spins) & 0xfff) == 0 && globaltimer_ns() - t_start > 4 * B200_MBAR_TIMEOUT_NS) {
                    printf("b200pt: attention score pass producer stalled (block %d,%d,%d k %d v %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, next_k, next_v);
                    __trap();
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ---------------------------------------------------------------- MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc_s = umma_idesc_f16(128, BT, false, false);
            constexpr uint32_t idesc_dq = umma_idesc_f16(128, D, false, true);
            const uint64_t do_desc = umma_desc_sw128(smem_u32(sDO), 16, 1024);
            const uint64_t kk_desc0 = umma_desc_sw128(smem_u32(sK), 16, 1024);    // K tile as K-major B (scores)
            const uint64_t km_desc0 = umma_desc_sw128(smem_u32(sK), 8192, 1024);  // K tile as MN-major B (dQ += dS K)
            const uint64_t ds_desc = umma_desc_sw128(smem_u32(sDS), 16, 1024);
            auto issue_s = [&](int t) {  // S_t = Q K_t^T (A = Q from TMEM: 16 bf16 = 8 columns per k step)
                const uint64_t koff = static_cast<uint64_t>((kslot(t) * KV_BYTES) >> 4);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk)
                    umma_ts(tmem + TM_S, tmem + TM_Q + kk * 8, kk_desc0 + koff + static_cast<uint64_t>(((kk >> 2) * 8192 + (kk & 3) * 32) >> 4),
                            idesc_s, kk != 0);
            };
            auto issue_dp = [&](int t) {  // dP_t = dO V_t^T, then the V slot is free and the scores are complete
                const int vs = vslot(t);
                const uint64_t voff = static_cast<uint64_t>((vs * KV_BYTES) >> 4);
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk)
                    umma_ss(tmem + TM_DP, do_desc + static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                            kk_desc0 + voff + static_cast<uint64_t>(((kk >> 2) * 8192 + (kk & 3) * 32) >> 4), idesc_s, kk != 0);
                tc_commit(&slot_empty[vs]);
                tc_commit(s_full);
            };
            auto issue_dq = [&](int t) {  // dQ += dS_t K_t, then the K slot is free
                const uint64_t soff = static_cast<uint64_t>((kslot(t) * KV_BYTES) >> 4);
#pragma unroll
                for (int kk = 0; kk < BT / 16; ++kk)
                    umma_ss(tmem + TM_DQ, ds_desc + static_cast<uint64_t>((kk * 32) >> 4), km_desc0 + soff + static_cast<uint64_t>((kk * 2048) >> 4),
                            idesc_dq, (t | kk) != 0);
                tc_commit(&slot_empty[kslot(t)]);
                tc_commit(dq_done);
            };
            mbar_wait(q_ready, 0);
            mbar_wait(do_full, 0);
            mbar_wait(&slot_full[kslot(0)], 0);
            tc_fence_after();
            issue_s(0);
            mbar_wait(&slot_full[vslot(0)], 0);
            tc_fence_after();
            issue_dp(0);
            for (int t = 0; t < n_tiles; ++t) {
                // Three things to issue, in whatever order their inputs arrive: S_{t+1} (compute warps pulled tile t out of
                // TMEM, K_{t+1} landed), then dP_{t+1} (V_{t+1} landed), and dQ_t (dS_t is in shared memory). Waiting for them
                // in a fixed order delays dQ_t behind a V load, which delays the slot that very load family needs next.
                bool need_s = t + 1 < n_tiles, need_dp = need_s, need_dq = true;
                const uint32_t ph1 = ((t + 1) >> 1) & 1;
                const uint64_t t_start = globaltimer_ns();
                uint32_t spins = 0;
                while (need_s || need_dp || need_dq) {
                    if (((++spins) & 0xfff) == 0 && globaltimer_ns() - t_start > B200_MBAR_TIMEOUT_NS) {
                        printf("b200pt: attention score pass stalled (block %d,%d,%d tile %d: s %d dp %d dq %d)\n", blockIdx.x, blockIdx.y,
                               blockIdx.z, t, need_s, need_dp, need_dq);
                        __trap();
                    }
                    if (need_s && mbar_try_wait(s_free, t & 1) && mbar_try_wait(&slot_full[kslot(t + 1)], ph1)) {
                        tc_fence_after();
                        trace_evt(tr, 16 * t + 0);
                        issue_s(t + 1);
                        need_s = false;
                    }
                    if (!need_s && need_dp && mbar_try_wait(&slot_full[vslot(t + 1)], ph1)) {
                        tc_fence_after();
                        issue_dp(t + 1);
                        trace_evt(tr, 16 * t + 1);
                        need_dp = false;
                    }
                    if (need_dq && mbar_try_wait(a_ready, t & 1)) {
                        tc_fence_after();
                        trace_evt(tr, 16 * t + 2);
                        issue_dq(t);
                        trace_evt(tr, 16 * t + 3);
                        need_dq = false;
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- compute warps: thread = query row = TMEM lane
        const int r = warp * 32 + lane;
        const int r_idx = r0 + r;
        const bool row_ok = r_idx < p.S;
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        const float sl2 = p.scale * LOG2E_F;
        const size_t stat_base = (static_cast<size_t>(b) * p.H + h) * p.S;
        // rows beyond S: lse = +inf makes every P (and dS) of the row exactly 0
        const float neg_lse2 = row_ok ? -p.lse[stat_base + r_idx] * LOG2E_F : -INFINITY;
        // (computing delta = rowsum(dO * O) here instead of in the pre-pass was tried: the extra 64 loads per thread sit on the
        // CTA's un-overlapped prologue and cost 0.2 ms per layer against the 0.06 ms of the separate kernel)
        const float my_delta = row_ok ? p.delta[stat_base + r_idx] : 0.f;
        const int S_ = p.S;
        const bool causal_ = p.causal != 0;
        {
            // Q row: staging tile (SWIZZLE_128B, 4 sub-tiles of 64 columns) -> registers -> TMEM (bf16 pairs, 128 columns)
            mbar_wait(q_full, 0);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 u = ld_shared_v4(sKV + 2 * KV_BYTES + c * 16384 + sw128_offset(r, i));
                    v[i * 4 + 0] = u.x, v[i * 4 + 1] = u.y, v[i * 4 + 2] = u.z, v[i * 4 + 3] = u.w;
                }
                tmem_st_32x32(lane_addr + TM_Q + c * 32, v);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_ready);
        }
        for (int t = 0; t < n_tiles; ++t) {
            const int c0 = t * BT;
            const bool tr0 = tr && threadIdx.x == 0;
            trace_evt(tr0, 16 * t + 8);
            mbar_wait(s_full, t & 1);
            tc_fence_after();
            trace_evt(tr0, 16 * t + 9);
            uint32_t sv[64], dv[64];
            tmem_ld_32x32(lane_addr + TM_S, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
            tmem_ld_32x32(lane_addr + TM_S + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
            tmem_ld_32x32(lane_addr + TM_DP, *reinterpret_cast<uint32_t(*)[32]>(&dv[0]));
            tmem_ld_32x32(lane_addr + TM_DP + 32, *reinterpret_cast<uint32_t(*)[32]>(&dv[32]));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);
            trace_evt(tr0, 16 * t + 10);
            // masking is branch-free inside the tile (scores -> -inf, exp2 -> 0) and skipped for tiles below the diagonal
            const bool need_mask = (c0 + BT > S_) || (causal_ && c0 + BT - 1 > r0);
            if (need_mask) {
                const int lim = causal_ ? min(S_ - 1, r_idx) : S_ - 1;  // last visible key of this row
#pragma unroll
                for (int i = 0; i < 64; ++i) sv[i] = (c0 + i > lim) ? 0xff800000u : sv[i];
            }
            uint32_t pk[32], dk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float pe0 = ex2(fmaf(__uint_as_float(sv[2 * i]), sl2, neg_lse2));
                const float pe1 = ex2(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, neg_lse2));
                pk[i] = f2_to_bf2(pe0, pe1);
                dk[i] = f2_to_bf2(pe0 * (__uint_as_float(dv[2 * i]) - my_delta), pe1 * (__uint_as_float(dv[2 * i + 1]) - my_delta));
            }
            // the P / dS tiles of the previous iteration must have been consumed: dS by the dQ MMAs, both by the TMA stores
            trace_evt(tr0, 16 * t + 11);
            if (t > 0) mbar_wait(dq_done, (t - 1) & 1);
            trace_evt(tr0, 16 * t + 12);
            if (threadIdx.x == 0) tma_store_wait_read<0>();
            trace_evt(tr0, 16 * t + 13);
            named_bar_sync(1, 128);
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                st_shared_v4(sP + sw128_offset(r, cc), make_uint4(pk[cc * 4], pk[cc * 4 + 1], pk[cc * 4 + 2], pk[cc * 4 + 3]));
                st_shared_v4(sDS + sw128_offset(r, cc), make_uint4(dk[cc * 4], dk[cc * 4 + 1], dk[cc * 4 + 2], dk[cc * 4 + 3]));
            }
            fence_proxy_async_smem();
            trace_evt(tr0, 16 * t + 14);
            named_bar_sync(1, 128);
            trace_evt(tr0, 16 * t + 15);
            if (threadIdx.x == 0) {
                mbar_arrive(a_ready);
                tma_store_2d(&tmP, sP, c0, zrow + r0);
                tma_store_2d(&tmDS, sDS, c0, zrow + r0);
                tma_store_commit();
            }
        }
        // ---- the dK/dV GEMMs reduce over queries starting at the first row of their 256-key tile: for the first 128
        // queries of such a tile they read keys [r0+128, r0+256), which this causal pass never visits -> store zeros.
        const bool zero_quad = p.causal && (blk & 1) == 0 && r0 + 128 < p.S;
        if (threadIdx.x == 0) tma_store_wait_read<0>();
        if (n_tiles > 0) {
            mbar_wait(dq_done, (n_tiles - 1) & 1);
            tc_fence_after();
        }
        named_bar_sync(1, 128);
        if (zero_quad) {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) st_shared_v4(sP + sw128_offset(r, cc), make_uint4(0, 0, 0, 0));
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (threadIdx.x == 0) {
                for (int c = r0 + 128; c < min(p.S, r0 + 256); c += 64) {
                    tma_store_2d(&tmP, sP, c, zrow + r0);
                    tma_store_2d(&tmDS, sP, c, zrow + r0);
                }
                tma_store_commit();
            }
        }
        // ---- epilogue: dQ * scale -> bf16, staged through the (now idle) dO tile, TMA-stored
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
            uint32_t v[32];
            if (n_tiles > 0) {
                tmem_ld_32x32(lane_addr + TM_DQ + c * 32, v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 o;
                o.x = f2_to_bf2(__uint_as_float(v[g * 8 + 0]) * p.scale, __uint_as_float(v[g * 8 + 1]) * p.scale);
                o.y = f2_to_bf2(__uint_as_float(v[g * 8 + 2]) * p.scale, __uint_as_float(v[g * 8 + 3]) * p.scale);
                o.z = f2_to_bf2(__uint_as_float(v[g * 8 + 4]) * p.scale, __uint_as_float(v[g * 8 + 5]) * p.scale);
                o.w = f2_to_bf2(__uint_as_float(v[g * 8 + 6]) * p.scale, __uint_as_float(v[g * 8 + 7]) * p.scale);
                st_shared_v4(sDO + (c >> 1) * 16384 + sw128_offset(r, (c & 1) * 4 + g), o);
            }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (threadIdx.x == 0) {
            for (int c = 0; c < NSUB; ++c) tma_store_2d(&tmDQ, sDO + c * 16384, col0 + c * 64, row_base + r0);
            tma_store_commit();
            tma_store_wait_read<0>();  // shared memory must outlive the bulk stores reading it
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
    B200_REQUIRE(a->D == 64 || a->D == 80 || a->D == 128 || a->D == 256, "%s: head_dim %d unsupported (64, 80, 128, 256)", who, a->D);
    B200_REQUIRE(a->B > 0 && a->S > 0 && a->H > 0, "%s: bad B/S/H", who);
    B200_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "%s: dropout_p must be in [0, 1)", who);
    B200_REQUIRE(a->qkv_row_stride % 8 == 0 && a->qkv_head_stride % 8 == 0, "%s: q/k/v strides must be multiples of 8 elements", who);
    B200_REQUIRE(a->o_row_stride % 8 == 0 && a->o_head_stride % 8 == 0 && aligned16(a->o), "%s: o must be 16B aligned with strides %% 8 == 0", who);
    B200_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v), "%s: q/k/v must be 16B aligned", who);
    return 0;
}

static inline bool padded_head(const b200_attn_args* a) { return a->D % 64 != 0; }

static int qkv_tmap(CUtensorMap* m, const void* ptr, const b200_attn_args* a, int64_t row_stride, int64_t head_stride, uint32_t box_rows) {
    if (padded_head(a))
        return make_tmap_bf16_3d(m, ptr, a->D, a->H, static_cast<uint64_t>(a->B) * a->S, head_stride, row_stride, 64, box_rows);
    const uint64_t inner = static_cast<uint64_t>(a->H - 1) * head_stride + a->D;
    return make_tmap_bf16_2d(m, ptr, inner, static_cast<uint64_t>(a->B) * a->S, row_stride, 64, box_rows);
}

static AttnParams make_params(const b200_attn_args* a) {
    AttnParams p;
    p.B = a->B, p.S = a->S, p.H = a->H, p.D = a->D;
    p.causal = a->causal;
    p.scale = a->scale;
    p.qkv_head_stride = a->qkv_head_stride;
    p.o = static_cast<elem_t*>(a->o);
    p.o_row_stride = a->o_row_stride, p.o_head_stride = a->o_head_stride;
    p.lse = a->lse;
    p.delta = a->delta;
    p.dq = static_cast<elem_t*>(a->dq);
    p.dk = static_cast<elem_t*>(a->dk);
    p.dv = static_cast<elem_t*>(a->dv);
    p.dqkv_row_stride = a->dqkv_row_stride, p.dqkv_head_stride = a->dqkv_head_stride;
    p.p_out = nullptr, p.ds_out = nullptr;
    p.drop_thr = static_cast<uint32_t>(a->dropout_p * 65536.0f + 0.5f);
    p.drop_scale = 65536.0f / static_cast<float>(65536u - p.drop_thr);
    p.drop_seed = a->dropout_seed;
    p.d_o = static_cast<const elem_t*>(a->d_o);
    p.d_real = a->D;
    p.pad3d = padded_head(a) ? 1 : 0;
    static const int tr = getenv("B200_ATTN_TRACE") ? atoi(getenv("B200_ATTN_TRACE")) : 0;  // 1: first CTA, 2: a late CTA
    p.trace = tr;
    return p;
}

template <typename KernT>
static int set_smem(KernT kern, size_t bytes, const char* who) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) return fail(-2, "%s: cudaFuncSetAttribute(%zu) failed: %s", who, bytes, cudaGetErrorString(e));
    return 0;
}

template <int D, int BN, int STAGES, bool DROP>
static int launch_fwd_impl(const b200_attn_args* a, cudaStream_t st) {
    using L = FwdSmem<D, BN, STAGES>;
    static_assert(L::TOTAL <= 232448, "forward smem budget");
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = qkv_tmap(&tq, a->q, a, a->qkv_row_stride, a->qkv_head_stride, 128))) return rc;
    if ((rc = qkv_tmap(&tk, a->k, a, a->qkv_row_stride, a->qkv_head_stride, BN))) return rc;
    if ((rc = qkv_tmap(&tv, a->v, a, a->qkv_row_stride, a->qkv_head_stride, BN))) return rc;
    auto kern = attn_fwd_kernel<D, BN, STAGES, DROP>;
    if ((rc = set_smem(kern, L::TOTAL, "attention_fwd"))) return rc;
    dim3 grid((a->S + 127) / 128, a->H, a->B);
    launch_k(kern, dim3(grid), dim3(192), L::TOTAL, st, tq, tk, tv, make_params(a));
    return check_launch("attention_fwd");
}

template <int D, int BN, int STAGES>
static int launch_fwd(const b200_attn_args* a, cudaStream_t st) {
    return a->dropout_p > 0.f ? launch_fwd_impl<D, BN, STAGES, true>(a, st) : launch_fwd_impl<D, BN, STAGES, false>(a, st);
}

static int launch_fwd256(const b200_attn_args* a, cudaStream_t st) {
    using L = Fwd256Smem;
    static_assert(L::TOTAL <= 232448, "fwd256 smem budget");
    CUtensorMap tq, tk, tv, to;
    int rc;
    if ((rc = qkv_tmap(&tq, a->q, a, a->qkv_row_stride, a->qkv_head_stride, 128))) return rc;
    if ((rc = qkv_tmap(&tk, a->k, a, a->qkv_row_stride, a->qkv_head_stride, 64))) return rc;
    if ((rc = qkv_tmap(&tv, a->v, a, a->qkv_row_stride, a->qkv_head_stride, 64))) return rc;
    if ((rc = qkv_tmap(&to, a->o, a, a->o_row_stride, a->o_head_stride, 128))) return rc;
    auto kern = attn_fwd256_kernel;
    if ((rc = set_smem(kern, L::TOTAL, "attention_fwd256"))) return rc;
    dim3 grid((a->S + 127) / 128, a->H, a->B);
    launch_k(kern, dim3(grid), dim3(192), L::TOTAL, st, tq, tk, tv, to, make_params(a));
    return check_launch("attention_fwd256");
}

static int launch_fwd256s(const b200_attn_args* a, cudaStream_t st) {
    using L = Fwd256SSmem;
    static_assert(L::TOTAL <= 232448, "fwd256s smem budget");
    CUtensorMap tq, tk, tv, to;
    int rc;
    if ((rc = qkv_tmap(&tq, a->q, a, a->qkv_row_stride, a->qkv_head_stride, 128))) return rc;
    if ((rc = qkv_tmap(&tk, a->k, a, a->qkv_row_stride, a->qkv_head_stride, 64))) return rc;
    if ((rc = qkv_tmap(&tv, a->v, a, a->qkv_row_stride, a->qkv_head_stride, 64))) return rc;
    if ((rc = qkv_tmap(&to, a->o, a, a->o_row_stride, a->o_head_stride, 128))) return rc;
    auto kern = attn_fwd256s_kernel;
    if ((rc = set_smem(kern, L::TOTAL, "attention_fwd256s"))) return rc;
    dim3 grid((a->S + 127) / 128, a->H, a->B);
    launch_k(kern, dim3(grid), dim3(320), L::TOTAL, st, tq, tk, tv, to, make_params(a));
    return check_launch("attention_fwd256s");
}

template <int D, int DH, int STAGES, bool DKV, int SBUF, bool DROP>
static int launch_bwd_impl(const b200_attn_args* a, cudaStream_t st, bool store_scores) {
    using L = BwdSmem<D, STAGES, DKV, SBUF>;
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
    auto kern = attn_bwd_kernel<D, DH, STAGES, DKV, SBUF, DROP>;
    if ((rc = set_smem(kern, L::TOTAL, "attention_bwd"))) return rc;
    dim3 grid(((a->S + 127) / 128) * (DKV ? D / DH : 1), a->H, a->B);
    AttnParams prm = make_params(a);
    if (store_scores) {
        prm.p_out = static_cast<elem_t*>(a->p_scratch);
        prm.ds_out = static_cast<elem_t*>(a->ds_scratch);
    }
    launch_k(kern, dim3(grid), dim3(192), L::TOTAL, st, r1, r2, t1, t2, prm);
    return check_launch(DKV ? "attention_bwd_dkv" : "attention_bwd_dq");
}

template <int D, int DH, int STAGES, bool DKV, int SBUF = 2>
static int launch_bwd(const b200_attn_args* a, cudaStream_t st, bool store_scores = false) {
    return a->dropout_p > 0.f ? launch_bwd_impl<D, DH, STAGES, DKV, SBUF, true>(a, st, store_scores)
                              : launch_bwd_impl<D, DH, STAGES, DKV, SBUF, false>(a, st, store_scores);
}

static int launch_bwd_dq256(const b200_attn_args* a, cudaStream_t st) {
    using L = Dq256Smem;
    static_assert(L::TOTAL <= 232448, "dq256 smem budget");
    CUtensorMap tdo, tk, tv, tp, tds, tdq;
    int rc;
    b200_attn_args oa = *a;
    if ((rc = qkv_tmap(&tdo, a->d_o, &oa, a->o_row_stride, a->o_head_stride, 128))) return rc;
    if ((rc = qkv_tmap(&tk, a->k, a, a->qkv_row_stride, a->qkv_head_stride, 64))) return rc;
    if ((rc = qkv_tmap(&tv, a->v, a, a->qkv_row_stride, a->qkv_head_stride, 64))) return rc;
    const uint64_t zs = static_cast<uint64_t>(a->B) * a->H * a->S;
    if ((rc = make_tmap_bf16_2d(&tp, a->p_scratch, a->S, zs, a->S, 64, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&tds, a->ds_scratch, a->S, zs, a->S, 64, 128))) return rc;
    if ((rc = qkv_tmap(&tdq, a->dq, a, a->dqkv_row_stride, a->dqkv_head_stride, 128))) return rc;
    CUtensorMap tq;
    if ((rc = qkv_tmap(&tq, a->q, a, a->qkv_row_stride, a->qkv_head_stride, 128))) return rc;
    auto kern = attn_bwd_dq256_kernel;
    if ((rc = set_smem(kern, L::TOTAL, "attention_bwd_dq256"))) return rc;
    dim3 grid((a->S + 127) / 128, a->H, a->B);
    launch_k(kern, dim3(grid), dim3(192), L::TOTAL, st, tdo, tk, tv, tp, tds, tdq, tq, make_params(a));
    return check_launch("attention_bwd_dq256");
}

// dV = P^T dO and dK = scale * dS^T Q over the materialised score tiles: two batched (B*H problems) causal GEMMs on the
// CTA-pair engine (gemm.cu). A = scratch [B*H*S (query i), S (key j)] read as [K, M]; B = dO / Q rows read as [K, N].
int gemm_batched_pair(const b200_gemm_args* a, int nb, int nh, const int (&offs)[12], int causal_k, float alpha,
                      uint64_t a_rows_total, uint64_t b_rows_total, const CUtensorMap& tmC, cudaStream_t st);  // gemm.cu

static int dkv_from_scores(const b200_attn_args* a, cudaStream_t st) {
    const int S = a->S, H = a->H, D = a->D;
    for (int which = 0; which < 2; ++which) {  // 0: dV from P and dO, 1: dK from dS and Q
        b200_gemm_args g = {};
        g.M = S, g.N = D, g.K = S;
        g.A = which == 0 ? a->p_scratch : a->ds_scratch;
        g.lda = S, g.a_mn = 1;
        g.B = which == 0 ? a->d_o : a->q;
        g.ldb = which == 0 ? a->o_row_stride : a->qkv_row_stride;
        g.b_mn = 1;
        g.C = which == 0 ? a->dv : a->dk;
        g.ldc = a->dqkv_row_stride;
        const int b_head = static_cast<int>(which == 0 ? a->o_head_stride : a->qkv_head_stride);
        //                 a_k_b  a_k_h a_m_b a_m_h b_k_b b_k_h b_n_b b_n_h  c_m_b c_m_h c_n_b c_n_h
        const int offs[12] = {H * S, S,    0,    0,    S,    0,    0,    b_head, S,    0,    0,    static_cast<int>(a->dqkv_head_stride)};
        CUtensorMap tmC;
        int rc = make_tmap_bf16_2d(&tmC, g.C, static_cast<uint64_t>(H - 1) * a->dqkv_head_stride + D, static_cast<uint64_t>(a->B) * S,
                                   a->dqkv_row_stride, 64, 128);
        if (rc) return rc;
        rc = gemm_batched_pair(&g, a->B, H, offs, a->causal, which == 0 ? 1.0f : a->scale, static_cast<uint64_t>(a->B) * H * S,
                               static_cast<uint64_t>(a->B) * S, tmC, st);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace b200

using namespace b200;

// debug only (not part of include/b200pt.h): copies the score-pass trace buffer to the host
extern "C" int b200_debug_attn_trace(unsigned long long* host, int n) {
    return cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(unsigned long long) * (n < 8192 ? n : 8192)) == cudaSuccess ? 0 : -2;
}

extern "C" int b200_attention_fwd(const b200_attn_args* a, b200_stream_t stream) {
    int rc = check_common(a, "attention_fwd");
    if (rc) return rc;
    B200_REQUIRE(a->lse != nullptr, "attention_fwd: lse is required");
    cudaStream_t st = as_stream(stream);
    switch (a->D) {
        // deep K/V rings wherever shared memory allows: a TMA load takes ~2000 clocks under load, a tile a few hundred
        case 64: {
            static const bool one_cta = getenv("B200_ATTN_FWD_1CTA") != nullptr;  // perf triage only
            // Without dropout: 64-key blocks, two CTAs per SM. With dropout: the one-CTA layout (128-key blocks). The dropout
            // instantiation of the two-CTA variant (168 registers) raised an illegal-address fault about once per 10^4 CTAs in
            // round 1 (triage log in DESIGN.md); its cause was never found, so that instantiation is NOT built any more
            // (attn_fwd_kernel<64, 64, 3, true> does not exist in the library) rather than shipped behind a switch.
            if (a->dropout_p > 0.f) return launch_fwd_impl<64, 128, 4, true>(a, st);
            if (one_cta) return launch_fwd_impl<64, 128, 4, false>(a, st);
            return launch_fwd_impl<64, 64, 3, false>(a, st);
        }
        case 80:  // zero-padded to 128 by the 3-D tensor maps
        case 128: return launch_fwd<128, 128, 2>(a, st);
        default: {
            static const bool old_fwd = getenv("B200_ATTN_OLD_FWD") != nullptr;  // perf triage only
            // B200_ATTN_FWD256_SPLIT=1: the variant with two softmax warps per TMEM lane quarter (attn_fwd256s_kernel). Measured
            // equal to the one-warp layout (0.306 vs 0.310 ms per layer, step 186.1 k vs 185.6 k tokens/s, profiles/
            // r02_attention_fwd256_analysis.txt): both warps of a pair sit on the same SM sub-partition and share its MUFU unit,
            // and the exp phase (64 ex2 per row per block = 512 MUFU clocks) is 680 of the 1400-clock per-block chain either way.
            static const bool split = getenv("B200_ATTN_FWD256_SPLIT") != nullptr;
            if (old_fwd || a->dropout_p > 0.f) return launch_fwd<256, 64, 2>(a, st);
            return split ? launch_fwd256s(a, st) : launch_fwd256(a, st);
        }
    }
}

extern "C" int b200_attention_bwd(const b200_attn_args* a, b200_stream_t stream) {
    int rc = check_common(a, "attention_bwd");
    if (rc) return rc;
    B200_REQUIRE(a->lse && a->delta && a->d_o && a->dq && a->dk && a->dv, "attention_bwd: lse, delta, d_o, dq, dk, dv are required");
    B200_REQUIRE(a->dqkv_row_stride % 8 == 0 && a->dqkv_head_stride % 8 == 0 && aligned16(a->dq) && aligned16(a->dk) && aligned16(a->dv) && aligned16(a->d_o),
                 "attention_bwd: gradient buffers must be 16B aligned with strides %% 8 == 0");
    cudaStream_t st = as_stream(stream);
    static const bool old_dq = getenv("B200_ATTN_OLD_DQ") != nullptr;  // perf triage only
    const bool score_path = a->D == 256 && a->p_scratch != nullptr && a->ds_scratch != nullptr && a->S % 256 == 0 && a->dropout_p == 0.f;
    {
        const int lpr = a->D <= 64 ? 8 : (a->D <= 128 ? 16 : 32);
        const int64_t total_warps = (static_cast<int64_t>(a->B) * a->S * a->H + (32 / lpr) - 1) / (32 / lpr);
        int64_t blocks = (total_warps + 7) / 8;
        const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
        if (blocks > cap) blocks = cap;
        auto go = [&](auto kern) {
            launch_k(kern, dim3(static_cast<int>(blocks)), dim3(256), 0, st, static_cast<const elem_t*>(a->o), static_cast<const elem_t*>(a->d_o),
                     a->delta, a->B, a->S, a->H, a->D, a->o_row_stride, a->o_head_stride);
        };
        if (lpr == 8) go(attn_delta_kernel<8>);
        else if (lpr == 16) go(attn_delta_kernel<16>);
        else go(attn_delta_kernel<32>);
        if ((rc = check_launch("attention_delta"))) return rc;
    }
    switch (a->D) {
        case 64:
            {
                static const bool one_cta = getenv("B200_ATTN_BWD_1CTA") != nullptr;  // perf triage only
                if (one_cta) {
                    if ((rc = launch_bwd<64, 64, 6, false>(a, st))) return rc;
                    return launch_bwd<64, 64, 6, true>(a, st);
                }
            }
            // two CTAs per SM (single score / operand buffers, 256 TMEM columns, < 113 KB smem each)
            if ((rc = launch_bwd<64, 64, 3, false, 1>(a, st))) return rc;
            return launch_bwd<64, 64, 2, true, 1>(a, st);
        case 80:
        case 128:
            if ((rc = launch_bwd<128, 128, 4, false>(a, st))) return rc;
            return launch_bwd<128, 128, 3, true>(a, st);
        default:
            if (score_path) {
                // head_dim 256: the dQ pass also writes its P / dS tiles; dK and dV become batched causal GEMMs (5 matmul
                // units instead of the 9 a TMEM-limited fused dK/dV pass needs at this head size)
                B200_REQUIRE(aligned16(a->p_scratch) && aligned16(a->ds_scratch), "attention_bwd: score scratch must be 16B aligned");
                B200_REQUIRE(static_cast<int64_t>(a->B) * a->H * a->S < (1ll << 31), "attention_bwd: B*H*S too large for the score scratch path");
                if ((rc = old_dq ? launch_bwd<256, 256, 1, false>(a, st, true) : launch_bwd_dq256(a, st))) return rc;
                return dkv_from_scores(a, st);
            }
            if ((rc = launch_bwd<256, 256, 1, false>(a, st))) return rc;
            return launch_bwd<256, 128, 1, true>(a, st);
    }
}
