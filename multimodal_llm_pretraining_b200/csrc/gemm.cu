// bf16 GEMM on 5th-gen tensor cores: TMA -> 128B-swizzled smem ring -> tcgen05.mma (cta_group::1, 128 x BN x 16) -> fp32
// accumulators in TMEM (double-buffered, 2 x BN columns) -> tcgen05.ld epilogue warps.
// Replaces nn.Linear forward / dgrad / wgrad (cuBLAS in the reference stack; HF:modeling_gpt_neox.py:41-42,200-201,464).
//
// Two tile engines share this file:
//   CG = 2 (default for M > 128, N > 128): CTA PAIRS (cluster of 2, tcgen05.mma.cta_group::2) compute 256 x 256 output
//           tiles; each CTA stages its own 128 rows of A and ONE HALF of the B tile, so B is read from shared memory and
//           fetched from L2 once per pair — 2/3 of the operand bytes per FLOP of the single-CTA engine, which is what
//           bounds a 128 x 256 x 64 stage (96 B/clk of the 128 B/clk shared-memory port).
//   CG = 1: one CTA per 128 x BN tile (narrow / short problems).
// Persistent: one CTA (pair) per SM (TPC) walks output tiles (grouped raster so weight tiles stay hot in the 126 MB L2).
// Batched mode (CG = 2 only): Z independent problems addressed through per-batch coordinate offsets into the same
// tensor maps, optional causal reduction range (k >= first row of the tile) — the dK / dV contractions of the
// attention backward over materialised P / dS tiles (attention.cu).
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..9 = epilogue in two groups of four (TMEM lane quarter = warp_idx % 4; groups alternate 64-column slabs).
// bf16 outputs are staged as 128x64 swizzled slabs in smem and written with TMA stores (full cache lines).  Three mbarrier pipelines: smem full/empty per stage,
// TMEM full/empty per accumulator buffer — the epilogue of tile i overlaps the main loop of tile i+1.
//
// Operand majorness is handled in hardware through the smem descriptors (no transposes in HBM):
//   K-major  X[R, K]: TMA box {64 k, R rows}, rows of 128 B, SBO = 1024.
//   MN-major X[K, R]: TMA boxes {64 r, 64 k} per 64-wide slice of R, LBO = 8192 (slice pitch), SBO = 1024.
#include <stdlib.h>

#include "api.h"
#include "common.cuh"

namespace b200 {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;

struct GemmParams {
    int M, N, K;
    void* C;
    int64_t ldc;
    int c_fp32;
    int accumulate;
    const float* bias;
    const elem_t* residual;
    int64_t ldr;
    int gelu;
    const float* alpha_dev;
    elem_t* aux_out;
    const elem_t* dgelu_in;
    int tiles_m, tiles_n;
    int tma_store;  // bf16 output without accumulate: stage 128x64 slabs in smem and TMA-store them (full-line writes)
    // batched mode: problem z = zb * zH + zh adds (zb * x_b + zh * x_h) to the TMA coordinate x of each operand
    int Z, zH;
    int a_k_b, a_k_h, a_m_b, a_m_h;
    int b_k_b, b_k_h, b_n_b, b_n_h;
    int c_m_b, c_m_h, c_n_b, c_n_h;
    int causal_k;  // reduction starts at k = first output row of the tile (rows of A^T below the diagonal are zero)
    float alpha_host;
    int split_k;  // fp32-accumulate outputs only: the reduction is cut into split_k ranges, each red.add'ed into C
    int debug;  // B200_GEMM_DEBUG (perf triage only): bit0 = skip epilogue math+stores, bit1 = skip TMEM loads as well
    // tile raster: group_m = 0 -> n runs fastest (a wave of CTAs = a few whole A row panels x every B panel: A is fetched from
    // DRAM once); group_m = g > 0 -> groups of g m-tiles, m fastest inside a group (the g A panels stay L2-resident while B streams)
    int group_m;
    unsigned long long hint_a, hint_b;  // L2 eviction-priority hints of the operand loads (common.cuh)
    // fused nn.Dropout on the value BEFORE the residual is added (RoBERTa: LN(dropout(W x + b) + residual)): the same counter-based
    // mask as elementwise.cu: dropout_kernel over the flat element index row * ldc + col, so the stand-alone kernel applied to the
    // gradient with the same seed is its backward. drop_thr = 0: off.
    uint32_t drop_thr;
    float drop_scale;
    unsigned long long drop_seed;
    // dGELU dgrad only: colsum_out[n] += sum_m C[m, n] — the bias gradient of the Linear in front of the GELU, reduced in the
    // epilogue (its warps wait for the tensor pipe most of the time) instead of a separate pass over the 0.5 GB dh1 tensor
    float* colsum_out;
};

__device__ __forceinline__ uint64_t gemm_mix64(uint64_t x) {  // splitmix64 finaliser (== elementwise.cu: mix64)
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int GROUP, int& tm, int& tn) {
    if (GROUP <= 0) {  // n fastest
        tm = tile / tiles_n;
        tn = tile - tm * tiles_n;
        return;
    }
    // groups of GROUP m-tiles; inside a group m runs fastest so GROUP consecutive CTAs share one B tile
    const int group_size = GROUP * tiles_n;
    const int group = tile / group_size;
    const int first_m = group * GROUP;
    const int gm = min(tiles_m - first_m, GROUP);
    const int r = tile - group * group_size;
    tm = first_m + r % gm;
    tn = r / gm;
}

// `side` = the 32-column slice of the residual / dgelu_in row, prefetched one chunk ahead so that its global-load latency
// overlaps the previous chunk's math; `sbias` = this tile's bias staged in shared memory (one coalesced load per tile).
// With p.tma_store the bf16 results go to the 128-row x 64-column SWIZZLE_128B staging slab in shared memory
// (`stage_main`, this chunk being half `half` of the slab, `r` = row inside the tile) instead of global.
// EPI (compile time, keeps the instruction footprint of the hot loop small — a fully generic, fully unrolled epilogue
// was 350 KB of SASS and instruction-fetch bound): 0 = bias/residual/alpha/accumulate, 1 = + exact-erf GELU (+ aux
// pre-activation output), 2 = + multiply by gelu'(dgelu_in).
template <int EPI, int SIDE>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&v)[32], const uint4 (&side)[4],
                                               const float* sbias, int row, int col0, float alpha, uint8_t* stage_main,
                                               uint8_t* stage_aux, int half, int r) {
    if (!p.tma_store && row >= p.M) return;
    float f[32];
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
        const int col = col0 + g8 * 8;
        if (col >= p.N) break;  // N % 8 == 0
        float* x = &f[g8 * 8];
        if (p.bias) {  // acc * alpha + bias in one FFMA per element
            const float4 b0 = *reinterpret_cast<const float4*>(sbias + g8 * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(sbias + g8 * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fmaf(__uint_as_float(v[g8 * 8 + j]), alpha, bb[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[g8 * 8 + j]) * alpha;
        }
        if (EPI == 1 && p.aux_out) {
            const uint4 av = make_uint4(f2_to_bf2(x[0], x[1]), f2_to_bf2(x[2], x[3]), f2_to_bf2(x[4], x[5]), f2_to_bf2(x[6], x[7]));
            if (p.tma_store) st_shared_v4(stage_aux + sw128_offset(r, half * 4 + g8), av);
            else if (row < p.M) st_v4(p.aux_out + static_cast<size_t>(row) * p.ldc + col, av);
        }
        if (EPI == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = gelu_erf(x[j]);
        }
        // SIDE = 1: the residual / dgelu_in tile was TMA-loaded into the output slab itself and is replaced in place
        const uint4 side_v = SIDE ? ld_shared_v4(stage_main + sw128_offset(r, half * 4 + g8)) : side[g8];
        if (EPI == 2) {
            const uint4 hv = side_v;
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 h = bf2_to_f2(hw[j]);
                x[2 * j] *= gelu_erf_grad(h.x);
                x[2 * j + 1] *= gelu_erf_grad(h.y);
            }
            if (p.colsum_out) {  // uniform; rows beyond M hold exact zeros (zero-filled A and side tiles)
                float cs[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float t = x[j];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    cs[j] = t;
                }
                if ((threadIdx.x & 31) == 0) {
                    red_add_v4(p.colsum_out + col, cs[0], cs[1], cs[2], cs[3]);
                    red_add_v4(p.colsum_out + col + 4, cs[4], cs[5], cs[6], cs[7]);
                }
            }
        }
        if (EPI == 0 && SIDE && p.drop_thr) {
            // 8 consecutive elements = vector i of the flat [M, ldc] output: lanes of two 64-bit hashes (see dropout_kernel)
            const uint64_t vi = (static_cast<uint64_t>(row) * static_cast<uint64_t>(p.ldc) + static_cast<uint64_t>(col)) >> 3;
            const uint64_t h0 = gemm_mix64(p.drop_seed + 0x9e3779b97f4a7c15ull * (2 * vi + 1));
            const uint64_t h1 = gemm_mix64(p.drop_seed + 0x9e3779b97f4a7c15ull * (2 * vi + 2));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t hh = j < 2 ? h0 : h1;
                const uint32_t ra = static_cast<uint32_t>(hh >> (32 * (j & 1))) & 0xffffu, rb = static_cast<uint32_t>(hh >> (32 * (j & 1) + 16)) & 0xffffu;
                x[2 * j] = ra >= p.drop_thr ? x[2 * j] * p.drop_scale : 0.f;
                x[2 * j + 1] = rb >= p.drop_thr ? x[2 * j + 1] * p.drop_scale : 0.f;
            }
        }
        if (p.residual) {
            const uint4 rv = side_v;
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 r = bf2_to_f2(rw[j]);
                x[2 * j] += r.x;
                x[2 * j + 1] += r.y;
            }
        }
        if (p.tma_store) {
            st_shared_v4(stage_main + sw128_offset(r, half * 4 + g8),
                         make_uint4(f2_to_bf2(x[0], x[1]), f2_to_bf2(x[2], x[3]), f2_to_bf2(x[4], x[5]), f2_to_bf2(x[6], x[7])));
        } else if (p.c_fp32) {
            float* c = static_cast<float*>(p.C) + static_cast<size_t>(row) * p.ldc + col;
            if (p.accumulate) {
                // C += x without reading C: one vector reduction per 16 B, resolved in L2 (also what makes split-K free
                // of a fix-up pass: every k-range of a tile just adds its partial sum)
                red_add_v4(c, x[0], x[1], x[2], x[3]);
                red_add_v4(c + 4, x[4], x[5], x[6], x[7]);
            } else {
                *reinterpret_cast<float4*>(c) = make_float4(x[0], x[1], x[2], x[3]);
                *reinterpret_cast<float4*>(c + 4) = make_float4(x[4], x[5], x[6], x[7]);
            }
        } else {
            elem_t* c = static_cast<elem_t*>(p.C) + static_cast<size_t>(row) * p.ldc + col;
            if (p.accumulate) {
                const uint4 cv = *reinterpret_cast<const uint4*>(c);
                const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 r = bf2_to_f2(cw[j]);
                    x[2 * j] += r.x;
                    x[2 * j + 1] += r.y;
                }
            }
            st_v4(c, make_uint4(f2_to_bf2(x[0], x[1]), f2_to_bf2(x[2], x[3]), f2_to_bf2(x[4], x[5]), f2_to_bf2(x[6], x[7])));
        }
    }
}

template <int CG>
struct GemmGeom {
    static constexpr int TILE_M = GEMM_BM * CG;  // output rows per tile (per CTA pair when CG = 2)
};

// smem per CTA: ring of STAGES x (A 16 KB + B (BN/CG) x 128 B) + two 16 KB output slabs + barriers + bias [BN] + 1 KB alignment slack
template <int EPI, int SIDE>
constexpr uint32_t gemm_stage_bytes() {
    // per epilogue group: one 16 KB output slab, plus one for the pre-GELU aux output (EPI 1) or to double-buffer the
    // TMA-loaded residual / dgelu_in tiles (SIDE)
    return (EPI == 1 || SIDE) ? 65536u : 32768u;
}
template <int BN, int STAGES, int CG, int EPI, int SIDE>
constexpr size_t gemm_smem_bytes_cg() {
    constexpr size_t need = static_cast<size_t>(STAGES) * (GEMM_BM * GEMM_BK * 2 + (BN / CG) * GEMM_BK * 2) + 256 + BN * 4 + gemm_stage_bytes<EPI, SIDE>();
    return need + 1024 <= 232448 ? need + 1024 : 232448;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI, int CG, int SIDE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux, const GemmParams p) {
    constexpr int BNH = BN / CG;                          // B rows staged by this CTA
    constexpr int TILE_M = GEMM_BM * CG;
    constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
    constexpr uint32_t B_BYTES = BNH * GEMM_BK * 2;
    constexpr uint32_t SLICE_BYTES = 64 * GEMM_BK * 2;   // one 64-wide MN slice of an MN-major operand (8 KB)
    constexpr uint32_t TMEM_COLS = 2 * BN;
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
    static_assert(CG == 1 || CG == 2, "cta_group");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    constexpr uint32_t RING_BYTES = STAGES * (A_BYTES + B_BYTES);
    constexpr uint32_t STG = gemm_stage_bytes<EPI, SIDE>();
    static_assert(!(SIDE && EPI == 1), "the GELU variant has no side operand");
    uint8_t* sstage = smem + RING_BYTES;  // output slabs (1024-aligned: the ring is a multiple of 16 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING_BYTES + STG);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint64_t* side_full = bars + 2 * STAGES + 4;  // [2 groups][2 buffers] (SIDE only)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);
    float* sbias = reinterpret_cast<float*>(smem + RING_BYTES + STG + 256);  // [BN]
    static_assert(RING_BYTES % 1024 == 0, "staging slabs must be 1024B aligned");
    static_assert((2 * STAGES + 9) * 8 <= 256, "barrier block overflow");
    if (threadIdx.x == 0 && (smem + RING_BYTES + STG + 256 + BN * 4) > (smem_raw + gemm_smem_bytes_cg<BN, STAGES, CG, EPI, SIDE>())) {
        printf("b200pt gemm: dynamic smem base misaligned beyond slack\n");
        __trap();
    }

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;  // 0 = leader (issues the MMAs)
    const int worker = CG == 2 ? (blockIdx.x >> 1) : blockIdx.x;
    const int n_workers = CG == 2 ? (gridDim.x >> 1) : gridDim.x;
    const int tiles_per_z = p.tiles_m * p.tiles_n;
    const int num_tiles = tiles_per_z * p.Z * p.split_k;  // work units: (k-range, problem, tile)
    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 8 * CG);  // every epilogue warp of every CTA of the pair
        }
        for (int s = 0; s < 4; ++s) mbar_init(&side_full[s], 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CG == 2) {
            tmem_alloc_pair(tmem_slot, TMEM_COLS);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_slot, TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_prologue();  // set-up (barriers, TMEM) overlapped the previous kernel's tail; global memory only from here on

    // tile -> (z, tm, tn), first k block; identical in every role
    auto decode = [&](int unit, int& z, int& tm, int& tn, int& kb0, int& kb1) {
        const int per_split = tiles_per_z * p.Z;
        const int ks = unit / per_split;
        const int tile = unit - ks * per_split;
        z = tile / tiles_per_z;
        tile_coords(tile - z * tiles_per_z, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
        kb0 = p.causal_k ? (tm * TILE_M) / GEMM_BK : 0;
        kb1 = num_kb;
        if (p.split_k > 1) {
            kb0 = static_cast<int>(static_cast<int64_t>(num_kb) * ks / p.split_k);
            kb1 = static_cast<int>(static_cast<int64_t>(num_kb) * (ks + 1) / p.split_k);
        }
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (every CTA loads its own half)
        // ONE thread runs the whole loop: no warp-wide waits, no reconvergence, no election per iteration.
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = worker; tile < num_tiles; tile += n_workers) {
                int z, tm, tn, kb0, kb1;
                decode(tile, z, tm, tn, kb0, kb1);
                const int zb = z / p.zH, zh = z - zb * p.zH;
                const int m0 = tm * TILE_M + static_cast<int>(rank) * GEMM_BM + zb * p.a_m_b + zh * p.a_m_h;
                const int n0 = tn * BN + static_cast<int>(rank) * BNH + zb * p.b_n_b + zh * p.b_n_h;
                const int ak0 = zb * p.a_k_b + zh * p.a_k_h, bk0 = zb * p.b_k_b + zh * p.b_k_h;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    if (rank == 0) mbar_expect_tx(&full[s], CG * (A_BYTES + B_BYTES));
                    uint8_t* a_dst = sA + s * A_BYTES;
                    uint8_t* b_dst = sB + s * B_BYTES;
                    auto ld = [&](void* dst, const CUtensorMap* m, int c0, int c1, uint64_t hint) {
                        if (CG == 2) tma_load_2d_pair_hint(dst, m, &full[s], c0, c1, hint);
                        else tma_load_2d_hint(dst, m, &full[s], c0, c1, hint);
                    };
                    if (!A_MN) {
                        ld(a_dst, &tmA, ak0 + kb * GEMM_BK, m0, p.hint_a);
                    } else {
#pragma unroll
                        for (int j = 0; j < GEMM_BM / 64; ++j) ld(a_dst + j * SLICE_BYTES, &tmA, m0 + j * 64, ak0 + kb * GEMM_BK, p.hint_a);
                    }
                    if (!B_MN) {
                        ld(b_dst, &tmB, bk0 + kb * GEMM_BK, n0, p.hint_b);
                    } else {
#pragma unroll
                        for (int j = 0; j < BNH / 64; ++j) ld(b_dst + j * SLICE_BYTES, &tmB, n0 + j * 64, bk0 + kb * GEMM_BK, p.hint_b);
                    }
                    if (++s == STAGES) s = 0, ph ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only, ONE thread)
        // The loop body must cost fewer issue cycles than the 4 x 128 tensor cycles it feeds: descriptors are advanced
        // by adding (byte offset >> 4) to precomputed 64-bit templates, and nothing in the loop is warp-collective.
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16(TILE_M, BN, A_MN, B_MN);
            const uint64_t adesc0 = A_MN ? umma_desc_sw128(smem_u32(sA), SLICE_BYTES, 1024) : umma_desc_sw128(smem_u32(sA), 16, 1024);
            const uint64_t bdesc0 = B_MN ? umma_desc_sw128(smem_u32(sB), SLICE_BYTES, 1024) : umma_desc_sw128(smem_u32(sB), 16, 1024);
            constexpr uint32_t A_KSTEP = (A_MN ? 2048u : 32u) >> 4, B_KSTEP = (B_MN ? 2048u : 32u) >> 4;
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            for (int tile = worker; tile < num_tiles; tile += n_workers) {
                int z, tm, tn, kb0, kb1;
                decode(tile, z, tm, tn, kb0, kb1);
                mbar_wait(&tempty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t ad = adesc0 + static_cast<uint64_t>((s * A_BYTES) >> 4);
                    const uint64_t bd = bdesc0 + static_cast<uint64_t>((s * B_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        if (CG == 2) umma_ss_pair(d_tmem, ad + k * A_KSTEP, bd + k * B_KSTEP, idesc, (kb > kb0) || k != 0);
                        else umma_ss(d_tmem, ad + k * A_KSTEP, bd + k * B_KSTEP, idesc, (kb > kb0) || k != 0);
                    }
                    if (CG == 2) {
                        tc_commit_pair(&empty[s], 3);                      // frees the slot in BOTH CTAs
                        if (kb == kb1 - 1) tc_commit_pair(&tfull[as], 3);  // accumulators complete in both CTAs
                    } else {
                        tc_commit(&empty[s]);
                        if (kb == kb1 - 1) tc_commit(&tfull[as]);
                    }
                    if (++s == STAGES) s = 0, ph ^= 1;
                }
                if (++as == 2) as = 0, aph ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9, two groups of four)
        // Group g (warps 2+4g .. 5+4g) owns the 64-column slabs with index = g (mod 2); inside a group warp w reads TMEM
        // lanes [32*(w%4), +32) (hardware restriction) = output rows. Two warps per SM sub-partition hide the TMEM /
        // global-load / erf latencies of each other.
        const int quarter = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int et_all = threadIdx.x - 64;  // 0..255 over both groups
        const int et = et_all & 127;          // 0..127 inside the group
        const int bar_id = 1 + grp;
        const float alpha = (p.alpha_dev ? __ldg(p.alpha_dev) : 1.0f) * p.alpha_host;
        constexpr int NC = BN / 32;
        uint8_t* st_main = sstage + grp * (STG / 2);
        uint8_t* st_aux = st_main + 16384;  // EPI == 1: pre-GELU slab; SIDE: second buffer of the in-place ring
        const int r_in_tile = quarter * 32 + lane;
        // ---- SIDE: this group's slabs form one sequence j = 0,1,2,... over all its tiles; slab j lives in buffer j & 1. The
        // residual / dgelu_in tile of slab j+1 is TMA-loaded while slab j is processed, so its latency is never exposed and
        // the loads are full 128-byte lines (per-thread global loads of a row-major tile touch 32 lines per instruction).
        auto slab_valid = [&](int tile, int k) {
            int z, tm, tn, kb0, kb1;
            decode(tile, z, tm, tn, kb0, kb1);
            return tn * BN + (grp + 2 * k) * 64 < p.N;
        };
        auto next_slab = [&](int& tile, int& k) {  // successor of (tile, k) in this group's sequence; tile >= num_tiles at the end
            do {
                if (k == 0) k = 1;
                else k = 0, tile += n_workers;
            } while (tile < num_tiles && !slab_valid(tile, k));
        };
        auto issue_side = [&](int tile, int k, int buf) {  // one elected thread
            int z, tm, tn, kb0, kb1;
            decode(tile, z, tm, tn, kb0, kb1);
            uint64_t* bar = &side_full[grp * 2 + buf];
            mbar_expect_tx(bar, 16384);
            tma_load_2d(st_main + buf * 16384, &tmAux, bar, tn * BN + (grp + 2 * k) * 64, tm * TILE_M + static_cast<int>(rank) * GEMM_BM);
        };
        int sj = 0;  // running slab count of this group
        if (SIDE) {
            int t0 = worker, k0 = 0;
            if (t0 < num_tiles && !slab_valid(t0, k0)) next_slab(t0, k0);
            if (et == 0 && t0 < num_tiles) issue_side(t0, k0, 0);
        }
        int as = 0;
        uint32_t aph = 0;
        for (int tile = worker; tile < num_tiles; tile += n_workers) {
            int z, tm, tn, kb0, kb1;
            decode(tile, z, tm, tn, kb0, kb1);
            const int zb = z / p.zH, zh = z - zb * p.zH;
            const int row0 = tm * TILE_M + static_cast<int>(rank) * GEMM_BM;  // first row of this CTA's 128-row slab
            const int row = row0 + quarter * 32 + lane;
            const int n0 = tn * BN;
            const int c_m0 = row0 + zb * p.c_m_b + zh * p.c_m_h, c_n0 = n0 + zb * p.c_n_b + zh * p.c_n_h;
            float* sb = sbias;
            if (p.bias) {
                // ONE bias buffer (a second one cost the 1 KB that the alignment slack of the 5-stage fused variants needs): every
                // epilogue warp must be done reading the previous tile's bias before it is overwritten, hence the barrier in front.
                // One coalesced load per tile replaces 32 dependent L1 round trips per thread.
                asm volatile("bar.sync 3, 256;" ::: "memory");
                for (int i = et_all; i < BN; i += 256) sb[i] = (n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
                asm volatile("bar.sync 3, 256;" ::: "memory");
            }
            const int rem = p.N - n0;
            const int nc_valid = rem >= BN ? NC : (rem + 31) / 32;  // chunks that hold at least one valid column
            mbar_wait(&tfull[as], aph);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
            const uint4 no_side[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
#pragma unroll 1
            for (int c0 = 2 * grp; c0 < nc_valid; c0 += 4) {  // rolled: one slab (two chunk bodies) of code
                const bool two = c0 + 1 < nc_valid;
                uint32_t v0[32], v1[32];
                tmem_ld_32x32(t_addr + c0 * 32, v0);
                if (two) tmem_ld_32x32(t_addr + (c0 + 1) * 32, v1);
                uint8_t* slab = st_main;
                if (SIDE) {
                    const int buf = sj & 1;
                    slab = st_main + buf * 16384;
                    if (et == 0) {
                        // the other buffer was last read by the TMA store of slab sj-1: once that has drained, refill it
                        tma_store_wait_read<0>();
                        int nt = tile, nk = (c0 - 2 * grp) >> 2;
                        next_slab(nt, nk);
                        if (nt < num_tiles) issue_side(nt, nk, buf ^ 1);
                    }
                    mbar_wait(&side_full[grp * 2 + buf], (sj >> 1) & 1);
                    ++sj;
                } else if (p.tma_store) {
                    // this group's slab buffer must have been drained by the TMA store that last read it
                    if (et == 0) tma_store_wait_read<0>();
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                }
                tmem_ld_wait();
                if (p.debug & 1) continue;
                epilogue_chunk<EPI, SIDE>(p, v0, no_side, sb + c0 * 32, row, n0 + c0 * 32, alpha, slab, st_aux, 0, r_in_tile);
                if (two) epilogue_chunk<EPI, SIDE>(p, v1, no_side, sb + (c0 + 1) * 32, row, n0 + (c0 + 1) * 32, alpha, slab, st_aux, 1, r_in_tile);
                if (p.tma_store) {
                    fence_proxy_async_smem();
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (et == 0) {
                        tma_store_2d(&tmC, slab, c_n0 + (c0 >> 1) * 64, c_m0);
                        if (EPI == 1 && p.aux_out) tma_store_2d(&tmAux, st_aux, c_n0 + (c0 >> 1) * 64, c_m0);
                        tma_store_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(&tempty[as], 0);  // the leader's barrier gates the next MMA into this buffer
                else mbar_arrive(&tempty[as]);
            }
            if (++as == 2) as = 0, aph ^= 1;
        }
        if (p.tma_store && et == 0) tma_store_wait_all<0>();  // smem must outlive the bulk stores reading it
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();  // nobody leaves while the peer may still signal its barriers
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

int resolve_driver() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
        return fail(-3, "cannot resolve cuTensorMapEncodeTiled: %s", cudaGetErrorString(e));
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    return 0;
}

// 2-D bf16 tensor map, 128B swizzle. inner = contiguous dimension.
int make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                      uint32_t box_inner, uint32_t box_outer) {
    int rc = resolve_driver();
    if (rc) return rc;
    // The driver entry point needs a current context on THIS thread; autograd's backward thread may only have had
    // cudaSetDevice() called. cudaFree(0) binds the primary context (once per thread).
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        cudaFree(0);
        ctx_bound = true;
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15u) || (pitch_elems * 2) % 16 != 0)
        return fail(-1, "tensor map: base must be 16B aligned and pitch a multiple of 8 elements (pitch=%llu)", (unsigned long long)pitch_elems);
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, B200_TMAP_ELEM_TYPE, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu pitch=%llu box=%ux%u", (int)r,
                                       (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems, box_inner, box_outer);
    return 0;
}


// deepest ring <= STAGES that still fits next to the GELU variant's second output slab
// 3-D bf16 map {d, head, token} with 128B swizzle: a head whose width is not a multiple of 64 (head_dim 80) is read as
// zero-padded 64-wide boxes, because the box is clipped at the extent of dimension 0 (attention.cu).
int make_tmap_bf16_3d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1_elems,
                      uint64_t pitch2_elems, uint32_t box0, uint32_t box2) {
    int rc = resolve_driver();
    if (rc) return rc;
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        cudaFree(0);
        ctx_bound = true;
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15u) || (pitch1_elems * 2) % 16 != 0 || (pitch2_elems * 2) % 16 != 0)
        return fail(-1, "tensor map 3d: base must be 16B aligned and pitches multiples of 8 elements");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {pitch1_elems * 2, pitch2_elems * 2};
    cuuint32_t box[3] = {box0, 1, box2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, B200_TMAP_ELEM_TYPE, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
    return 0;
}

// The fused-epilogue variants carry 64 KB of slabs. With CTA pairs (32 KB stages) five stages + slabs + barriers + ONE bias
// buffer come to 230 656 B; with the full 1 KB of alignment slack 231 680 B <= the 232 448 B per-CTA limit, so the layout holds for
// ANY alignment of the dynamic shared window (round 1 kept two bias buffers, had no room for the slack and relied on the window
// starting 1 KB-aligned). Four stages hold only 128 KB in flight, marginal against a ~2000-clock TMA latency at 64 B/clk per SM
// (the 4-stage plain GEMM measured 3 % slower than the 5-stage one).
template <int BN, int STAGES, int CG>
constexpr int gelu_stages() {
    int st = STAGES;
    while (static_cast<size_t>(st) * (GEMM_BM * GEMM_BK * 2 + (BN / CG) * GEMM_BK * 2) + 256 + BN * 4 + gemm_stage_bytes<1, 0>() + 1024 > 232448) --st;
    return st;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI, int CG, int SIDE = 0>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmAux,
                       const GemmParams& p, cudaStream_t st) {
    auto kern = gemm_kernel<BN, STAGES, A_MN, B_MN, EPI, CG, SIDE>;
    constexpr size_t smem = gemm_smem_bytes_cg<BN, STAGES, CG, EPI, SIDE>();
    static bool configured = false;  // benign race: attribute set is idempotent
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(-2, "gemm: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    const int tiles = p.tiles_m * p.tiles_n * p.Z * p.split_k;
    const int workers_max = num_sms() / CG;  // CTA pairs need both SMs of a TPC
    const int workers = tiles < workers_max ? tiles : workers_max;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(workers * CG, 1, 1);
    cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see common.cuh: pdl_trigger / pdl_wait
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmAux, p);
    if (e != cudaSuccess) return fail(-2, "gemm_bf16: launch failed: %s", cudaGetErrorString(e));
    return check_launch("gemm_bf16");
}

template <int BN, int STAGES, int CG>
static int dispatch_major(const b200_gemm_args* a, GemmParams& p, cudaStream_t st, const CUtensorMap* tmC_override = nullptr,
                          uint64_t a_rows_total = 0, uint64_t b_rows_total = 0) {
    constexpr int TILE_M = GEMM_BM * CG;
    constexpr int BNH = BN / CG;
    p.tiles_m = (p.M + TILE_M - 1) / TILE_M;
    p.tiles_n = (p.N + BN - 1) / BN;
    CUtensorMap tmA, tmB;
    int rc;
    // plain problem: the maps are clipped at the logical extents so that edge tiles are zero-filled; batched problems
    // address the whole buffers (their per-problem offsets are added to the coordinates in the kernel)
    const bool batched = p.Z > 1 || a_rows_total || b_rows_total;
    const uint64_t a_in = batched ? static_cast<uint64_t>(a->lda) : static_cast<uint64_t>(a->a_mn ? a->M : a->K);
    const uint64_t a_out = batched ? a_rows_total : static_cast<uint64_t>(a->a_mn ? a->K : a->M);
    const uint64_t b_in = batched ? static_cast<uint64_t>(a->ldb) : static_cast<uint64_t>(a->b_mn ? a->N : a->K);
    const uint64_t b_out = batched ? b_rows_total : static_cast<uint64_t>(a->b_mn ? a->K : a->N);
    if (!a->a_mn) rc = make_tmap_bf16_2d(&tmA, a->A, a_in, a_out, a->lda, GEMM_BK, GEMM_BM);
    else          rc = make_tmap_bf16_2d(&tmA, a->A, a_in, a_out, a->lda, 64, GEMM_BK);
    if (rc) return rc;
    if (!a->b_mn) rc = make_tmap_bf16_2d(&tmB, a->B, b_in, b_out, a->ldb, GEMM_BK, BNH);
    else          rc = make_tmap_bf16_2d(&tmB, a->B, b_in, b_out, a->ldb, 64, GEMM_BK);
    if (rc) return rc;
    CUtensorMap tmC = tmA, tmAux = tmA;  // placeholders when the TMA-store path is off
    if (p.tma_store) {
        if (tmC_override) tmC = *tmC_override;
        else if ((rc = make_tmap_bf16_2d(&tmC, a->C, a->N, a->M, a->ldc, 64, GEMM_BM))) return rc;
        if (a->aux_out && (rc = make_tmap_bf16_2d(&tmAux, a->aux_out, a->N, a->M, a->ldc, 64, GEMM_BM))) return rc;
    }
    if (a->gelu) {
        if (a->a_mn || a->b_mn) return fail(-1, "gemm: the GELU epilogue is built for the forward layout (a_mn=0, b_mn=0) only");
        return launch_gemm<BN, gelu_stages<BN, STAGES, CG>(), false, false, 1, CG>(tmA, tmB, tmC, tmAux, p, st);
    }
    // residual / dgelu_in ("side" operands) are TMA-loaded into the output slabs: bf16 TMA-stored outputs only
    const void* side = a->residual ? a->residual : a->dgelu_in;
    if (side) {
        if (!p.tma_store) return fail(-1, "gemm: residual / dgelu_in need a bf16, non-accumulating output");
        if ((rc = make_tmap_bf16_2d(&tmAux, side, a->N, a->M, a->ldr, 64, GEMM_BM))) return rc;
    }
    if (a->dgelu_in) {
        if (a->a_mn || !a->b_mn) return fail(-1, "gemm: the dGELU epilogue is built for the dgrad layout (a_mn=0, b_mn=1) only");
        return launch_gemm<BN, gelu_stages<BN, STAGES, CG>(), false, true, 2, CG, 1>(tmA, tmB, tmC, tmAux, p, st);
    }
    if (a->residual) {
        if (a->a_mn) return fail(-1, "gemm: the residual epilogue is built for the forward and dgrad layouts (a_mn=0) only");
        if (a->b_mn) return launch_gemm<BN, gelu_stages<BN, STAGES, CG>(), false, true, 0, CG, 1>(tmA, tmB, tmC, tmAux, p, st);
        return launch_gemm<BN, gelu_stages<BN, STAGES, CG>(), false, false, 0, CG, 1>(tmA, tmB, tmC, tmAux, p, st);
    }
    if (!a->a_mn && !a->b_mn) return launch_gemm<BN, STAGES, false, false, 0, CG>(tmA, tmB, tmC, tmAux, p, st);
    if (!a->a_mn && a->b_mn) return launch_gemm<BN, STAGES, false, true, 0, CG>(tmA, tmB, tmC, tmAux, p, st);
    if (a->a_mn && a->b_mn) return launch_gemm<BN, STAGES, true, true, 0, CG>(tmA, tmB, tmC, tmAux, p, st);
    return launch_gemm<BN, STAGES, true, false, 0, CG>(tmA, tmB, tmC, tmAux, p, st);
}

static void fill_params(const b200_gemm_args* a, GemmParams& p) {
    p = GemmParams{};
    p.M = a->M, p.N = a->N, p.K = a->K;
    p.C = a->C, p.ldc = a->ldc, p.c_fp32 = a->c_fp32, p.accumulate = a->accumulate;
    p.bias = a->bias;
    p.residual = static_cast<const elem_t*>(a->residual);
    p.ldr = a->ldr;
    p.gelu = a->gelu;
    p.alpha_dev = a->alpha_dev;
    p.aux_out = static_cast<elem_t*>(a->aux_out);
    p.dgelu_in = static_cast<const elem_t*>(a->dgelu_in);
    p.Z = 1, p.zH = 1;
    p.alpha_host = 1.0f;
    p.split_k = 1;
    p.group_m = 8;
    p.hint_a = L2_EVICT_NORMAL, p.hint_b = L2_EVICT_NORMAL;
    p.drop_thr = static_cast<uint32_t>(a->dropout_p * 65536.0f + 0.5f);
    p.drop_scale = 65536.0f / static_cast<float>(65536u - p.drop_thr);
    p.drop_seed = a->dropout_seed;
    p.colsum_out = a->colsum_out;
}

// Raster and L2 policy of a plain (non-batched) problem. A wave = one tile per CTA (pair), all streaming K in lockstep, so what
// costs DRAM bytes is (a) operand panels that are touched by more than one wave and have left L2 in between, (b) panels shared by
// tiles of different waves because a wave boundary cuts through their group. Round 1 walked groups of 8 m-tiles with m fastest:
// on 74 CTA pairs every group straddles two waves along n, so every A panel was fetched twice and the weights once per wave
// (ncu: 1.96x-2.24x the algorithmic bytes on the K = 8192 shapes). Now: when A (activations / output gradients) is the larger
// operand and the whole B fits in a fraction of L2, n runs fastest — a wave is a few complete A panels times every B panel, A is
// fetched once and marked EVICT_FIRST, B is marked EVICT_LAST and stays resident across waves. Otherwise (LM head: 206 MB of
// weights) groups of m-tiles whose A panels fit in L2 stay resident while B streams. B200_GEMM_GROUP_M / B200_GEMM_HINTS override
// (perf triage only).
static void choose_raster(const b200_gemm_args* a, GemmParams& p, int tile_m) {
    static const int env_group = getenv("B200_GEMM_GROUP_M") ? atoi(getenv("B200_GEMM_GROUP_M")) : -1;
    // hints are OFF by default: measured on the full step they cost 1.5-2 % (EVICT_FIRST lets an A line go before the last of the
    // CTAs sharing it has read it; the LM head with A pinned doubled its DRAM reads) while the raster alone gained 1.2 %
    // (profiles/r02_gemm_raster_ab.txt)
    static const int env_hints = getenv("B200_GEMM_HINTS") ? atoi(getenv("B200_GEMM_HINTS")) : 0;
    const double a_bytes = 2.0 * a->M * a->K, b_bytes = 2.0 * a->N * a->K;
    const double panel_a = 2.0 * tile_m * a->K / (p.split_k > 0 ? p.split_k : 1);
    constexpr double L2_RESIDENT = 40e6;  // what can be expected to stay in the 126 MB L2 next to a streaming operand
    const int tiles_m = (a->M + tile_m - 1) / tile_m;
    if (b_bytes <= L2_RESIDENT || a_bytes >= b_bytes) {
        p.group_m = 0;
        if (env_hints && b_bytes <= L2_RESIDENT) p.hint_a = L2_EVICT_FIRST, p.hint_b = L2_EVICT_LAST;
    } else {
        int g = static_cast<int>(L2_RESIDENT / panel_a);
        if (g < 8) g = 8;
        if (g > tiles_m) g = tiles_m;
        p.group_m = g;
        if (env_hints && g * panel_a <= L2_RESIDENT) p.hint_a = L2_EVICT_LAST, p.hint_b = L2_EVICT_FIRST;
    }
    if (env_group >= 0) p.group_m = env_group;
}

// dOut[z] = alpha * A[z]^T-or-not x B[z] over Z = nb * nh problems sharing the tensor maps (see GemmParams); bf16 TMA-stored
// output only. Used by the attention backward (attention.cu).
int gemm_batched_pair(const b200_gemm_args* a, int nb, int nh, const int (&offs)[12], int causal_k, float alpha,
                      uint64_t a_rows_total, uint64_t b_rows_total, const CUtensorMap& tmC, cudaStream_t st) {
    GemmParams p;
    fill_params(a, p);
    p.tma_store = 1;
    p.Z = nb * nh, p.zH = nh;
    p.a_k_b = offs[0], p.a_k_h = offs[1], p.a_m_b = offs[2], p.a_m_h = offs[3];
    p.b_k_b = offs[4], p.b_k_h = offs[5], p.b_n_b = offs[6], p.b_n_h = offs[7];
    p.c_m_b = offs[8], p.c_m_h = offs[9], p.c_n_b = offs[10], p.c_n_h = offs[11];
    p.causal_k = causal_k;
    p.alpha_host = alpha;
    if (a->M % 256 != 0 || a->N % 64 != 0 || a->K % GEMM_BK != 0) return fail(-1, "batched gemm: M %% 256, N %% 64, K %% 64 must be 0 (%d,%d,%d)", a->M, a->N, a->K);
    if (causal_k && a->K < a->M) return fail(-1, "batched gemm: causal reduction needs K >= M");
    return dispatch_major<256, 5, 2>(a, p, st, &tmC, a_rows_total, b_rows_total);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_gemm_bf16(const b200_gemm_args* a, b200_stream_t stream) {
    B200_REQUIRE(a != nullptr, "gemm: null args");
    B200_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "gemm: M,N,K must be positive (%d,%d,%d)", a->M, a->N, a->K);
    B200_REQUIRE(a->N % 8 == 0, "gemm: N (%d) must be a multiple of 8", a->N);
    B200_REQUIRE(a->ldc % 8 == 0 && aligned16(a->C), "gemm: C must be 16B aligned with ldc %% 8 == 0");
    B200_REQUIRE(!a->residual || (a->ldr % 8 == 0 && aligned16(a->residual)), "gemm: residual must be 16B aligned with ldr %% 8 == 0");
    B200_REQUIRE(!a->dgelu_in || (a->ldr % 8 == 0 && aligned16(a->dgelu_in)), "gemm: dgelu_in must be 16B aligned with ldr %% 8 == 0");
    B200_REQUIRE(!(a->residual && a->dgelu_in), "gemm: residual and dgelu_in are mutually exclusive");
    B200_REQUIRE(!a->bias || aligned16(a->bias), "gemm: bias must be 16B aligned");
    B200_REQUIRE(!a->aux_out || aligned16(a->aux_out), "gemm: aux_out must be 16B aligned");
    GemmParams p;
    fill_params(a, p);
#ifdef B200_PERF_TRIAGE  // bits 0/1 skip epilogue work and give WRONG results: only in a library built for profiling (-DB200_PERF_TRIAGE)
    static const int dbg = getenv("B200_GEMM_DEBUG") ? atoi(getenv("B200_GEMM_DEBUG")) : 0;
#else
    constexpr int dbg = 0;
#endif
    static const int force_cg = getenv("B200_GEMM_CG") ? atoi(getenv("B200_GEMM_CG")) : 0;  // perf triage: 1 = single-CTA engine
    p.debug = dbg;
    p.tma_store = (!a->c_fp32 && !a->accumulate && !(dbg & 64)) ? 1 : 0;
    B200_REQUIRE(!a->aux_out || a->gelu, "gemm: aux_out is the pre-GELU output and needs gelu=1");
    B200_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "gemm: dropout_p must be in [0, 1)");
    B200_REQUIRE(!a->colsum_out || (a->dgelu_in && aligned16(a->colsum_out)), "gemm: colsum_out is built for the dGELU dgrad epilogue (needs dgelu_in) and must be 16B aligned");
    B200_REQUIRE(a->dropout_p == 0.f || (a->residual && !a->gelu && !a->dgelu_in && !a->c_fp32 && !a->accumulate && a->ldc % 8 == 0),
                 "gemm: the fused dropout is built for the bias + dropout + residual epilogue with a 16-bit output");
    cudaStream_t st = as_stream(stream);
    // CTA pairs on 256 x 256 tiles when the problem fills them; 128 x 256 / 128 x 128 single-CTA tiles otherwise
    if (a->N > 128 && a->M > 128 && force_cg != 1) {
        if (a->c_fp32 && a->accumulate && !a->bias && !a->residual && !a->dgelu_in && !a->gelu && !(dbg & 128)) {
            // wgrad: few, very long tiles. Cut the reduction so that the work units fill whole waves of CTA pairs
            // (e.g. 256 tiles on 74 pairs: 3.46 waves -> split 2 -> 6.92); partial sums meet in C through red.add.
            const int tiles = ((a->M + 255) / 256) * ((a->N + 255) / 256);
            const int workers = num_sms() / 2;
            const int num_kb = (a->K + GEMM_BK - 1) / GEMM_BK;
            int best = 1;
            double best_eff = 0.0;
            for (int sk = 1; sk <= 8 && num_kb / sk >= 32; ++sk) {
                const double waves = static_cast<double>(tiles) * sk / workers;
                const double eff = waves / static_cast<double>(static_cast<int64_t>(waves + 0.999999));
                if (eff > best_eff + 0.02) best_eff = eff, best = sk;
            }
            p.split_k = best;
        }
        choose_raster(a, p, 256);
        return dispatch_major<256, 5, 2>(a, p, st);
    }
    choose_raster(a, p, 128);
    if (a->N > 128) return dispatch_major<256, 4, 1>(a, p, st);
    return dispatch_major<128, 6, 1>(a, p, st);
}
