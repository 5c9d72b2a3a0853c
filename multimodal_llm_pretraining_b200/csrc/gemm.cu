// bf16 GEMM on 5th-gen tensor cores: TMA -> 128B-swizzled smem ring -> tcgen05.mma (cta_group::1, 128 x BN x 16) -> fp32
// accumulators in TMEM (double-buffered, 2 x BN columns) -> tcgen05.ld epilogue warps.
// Replaces nn.Linear forward / dgrad / wgrad (cuBLAS in the reference stack; HF:modeling_gpt_neox.py:41-42,200-201,464).
//
// Persistent: one CTA per SM walks output tiles (grouped raster so weight tiles stay hot in the 126 MB L2).
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
    const __nv_bfloat16* residual;
    int64_t ldr;
    int gelu;
    const float* alpha_dev;
    __nv_bfloat16* aux_out;
    const __nv_bfloat16* dgelu_in;
    int tiles_m, tiles_n;
    int tma_store;  // bf16 output without accumulate: stage 128x64 slabs in smem and TMA-store them (full-line writes)
    int debug;  // B200_GEMM_DEBUG (perf triage only): bit0 = skip epilogue math+stores, bit1 = skip TMEM loads as well
};

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& tm, int& tn) {
    // groups of 8 m-tiles; inside a group m runs fastest so 8 consecutive CTAs share one B (weight) tile
    constexpr int GROUP = 8;
    const int group_size = GROUP * tiles_n;
    const int group = tile / group_size;
    const int first_m = group * GROUP;
    const int gm = min(tiles_m - first_m, GROUP);
    const int r = tile - group * group_size;
    tm = first_m + r % gm;
    tn = r / gm;
}

template <int BN, int STAGES>
constexpr size_t gemm_smem_bytes() {
    // ring + barriers (256) + bias [2][BN] fp32 + two 16 KB output slabs + alignment slack; the 227 KB per-CTA limit
    // (232448 B) leaves 256 B of slack less than a full 1 KB for BN=256 — the kernel traps if the base is that unlucky.
    constexpr size_t need = static_cast<size_t>(STAGES) * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + 256 + 2 * BN * 4 + 32768;
    return need + 1024 <= 232448 ? need + 1024 : 232448;
}

// `side` = the 32-column slice of the residual / dgelu_in row, prefetched one chunk ahead so that its global-load latency
// overlaps the previous chunk's math; `sbias` = this tile's bias staged in shared memory (one coalesced load per tile).
// With p.tma_store the bf16 results go to the 128-row x 64-column SWIZZLE_128B staging slab in shared memory
// (`stage_main`, this chunk being half `half` of the slab, `r` = row inside the tile) instead of global.
// EPI (compile time, keeps the instruction footprint of the hot loop small — a fully generic, fully unrolled epilogue
// was 350 KB of SASS and instruction-fetch bound): 0 = bias/residual/alpha/accumulate, 1 = + exact-erf GELU (+ aux
// pre-activation output), 2 = + multiply by gelu'(dgelu_in).
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&v)[32], const uint4 (&side)[4],
                                               const float* sbias, int row, int col0, float alpha, uint8_t* stage_main,
                                               int half, int r) {
    if (!p.tma_store && row >= p.M) return;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * alpha;
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
        const int col = col0 + g8 * 8;
        if (col >= p.N) break;  // N % 8 == 0
        float* x = &f[g8 * 8];
        if (p.bias) {
            const float4 b0 = *reinterpret_cast<const float4*>(sbias + g8 * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(sbias + g8 * 8 + 4);
            x[0] += b0.x, x[1] += b0.y, x[2] += b0.z, x[3] += b0.w;
            x[4] += b1.x, x[5] += b1.y, x[6] += b1.z, x[7] += b1.w;
        }
        if (EPI == 1 && p.aux_out) {
            const uint4 av = make_uint4(f2_to_bf2(x[0], x[1]), f2_to_bf2(x[2], x[3]), f2_to_bf2(x[4], x[5]), f2_to_bf2(x[6], x[7]));
            if (row < p.M) st_v4(p.aux_out + static_cast<size_t>(row) * p.ldc + col, av);
        }
        if (EPI == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = gelu_erf(x[j]);
        }
        if (EPI == 2) {
            const uint4 hv = side[g8];
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 h = bf2_to_f2(hw[j]);
                x[2 * j] *= gelu_erf_grad(h.x);
                x[2 * j + 1] *= gelu_erf_grad(h.y);
            }
        }
        if (p.residual) {
            const uint4 rv = side[g8];
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 r = bf2_to_f2(rw[j]);
                x[2 * j] += r.x;
                x[2 * j + 1] += r.y;
            }
        }
        if (p.tma_store) {
            *reinterpret_cast<uint4*>(stage_main + sw128_offset(r, half * 4 + g8)) =
                make_uint4(f2_to_bf2(x[0], x[1]), f2_to_bf2(x[2], x[3]), f2_to_bf2(x[4], x[5]), f2_to_bf2(x[6], x[7]));
        } else if (p.c_fp32) {
            float* c = static_cast<float*>(p.C) + static_cast<size_t>(row) * p.ldc + col;
            if (p.accumulate) {
                const float4 c0 = *reinterpret_cast<const float4*>(c);
                const float4 c1 = *reinterpret_cast<const float4*>(c + 4);
                x[0] += c0.x, x[1] += c0.y, x[2] += c0.z, x[3] += c0.w;
                x[4] += c1.x, x[5] += c1.y, x[6] += c1.z, x[7] += c1.w;
            }
            *reinterpret_cast<float4*>(c) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4*>(c + 4) = make_float4(x[4], x[5], x[6], x[7]);
        } else {
            __nv_bfloat16* c = static_cast<__nv_bfloat16*>(p.C) + static_cast<size_t>(row) * p.ldc + col;
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

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux, const GemmParams p) {
    constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
    constexpr uint32_t B_BYTES = BN * GEMM_BK * 2;
    constexpr uint32_t SLICE_BYTES = 64 * GEMM_BK * 2;   // one 64-wide MN slice of an MN-major operand (8 KB)
    constexpr uint32_t TMEM_COLS = 2 * BN;
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    constexpr uint32_t RING_BYTES = STAGES * (A_BYTES + B_BYTES);
    uint8_t* sstage = smem + RING_BYTES;  // 2 x 16 KB output slabs (1024-aligned: the ring is a multiple of 16 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING_BYTES + 32768);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* sbias = reinterpret_cast<float*>(smem + RING_BYTES + 32768 + 256);  // [2][BN]
    static_assert(RING_BYTES % 1024 == 0, "staging slabs must be 1024B aligned");
    if (threadIdx.x == 0 && (smem + RING_BYTES + 32768 + 256 + 2 * BN * 4) > (smem_raw + gemm_smem_bytes<BN, STAGES>())) {
        printf("b200pt gemm: dynamic smem base misaligned beyond slack\n");
        __trap();
    }

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_m * p.tiles_n;
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
            mbar_init(&tempty[s], 8);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int tm, tn;
            tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
            const int m0 = tm * GEMM_BM, n0 = tn * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty[s], ph ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                    uint8_t* a_dst = sA + s * A_BYTES;
                    uint8_t* b_dst = sB + s * B_BYTES;
                    if (!A_MN) {
                        tma_load_2d(a_dst, &tmA, &full[s], kb * GEMM_BK, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_2d(a_dst + j * SLICE_BYTES, &tmA, &full[s], m0 + j * 64, kb * GEMM_BK);
                    }
                    if (!B_MN) {
                        tma_load_2d(b_dst, &tmB, &full[s], kb * GEMM_BK, n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * SLICE_BYTES, &tmB, &full[s], n0 + j * 64, kb * GEMM_BK);
                    }
                }
                __syncwarp();
                if (++s == STAGES) s = 0, ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, A_MN, B_MN);
        int s = 0;
        uint32_t ph = 0;
        int as = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[as], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = smem_u32(sA + s * A_BYTES);
                    const uint32_t b_addr = smem_u32(sB + s * B_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        const uint64_t adesc = A_MN ? umma_desc_sw128(a_addr + k * 2048, SLICE_BYTES, 1024)
                                                    : umma_desc_sw128(a_addr + k * 32, 16, 1024);
                        const uint64_t bdesc = B_MN ? umma_desc_sw128(b_addr + k * 2048, SLICE_BYTES, 1024)
                                                    : umma_desc_sw128(b_addr + k * 32, 16, 1024);
                        umma_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
                    }
                    tc_commit(&empty[s]);                       // frees the smem slot when these MMAs retire
                    if (kb == num_kb - 1) tc_commit(&tfull[as]);  // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++s == STAGES) s = 0, ph ^= 1;
            }
            if (++as == 2) as = 0, aph ^= 1;
        }
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
        const float alpha = p.alpha_dev ? __ldg(p.alpha_dev) : 1.0f;
        const __nv_bfloat16* side_ptr = p.residual ? p.residual : p.dgelu_in;
        constexpr int NC = BN / 32;
        uint8_t* st_main = sstage + grp * 16384;
        int as = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int tm, tn;
            tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
            const int row = tm * GEMM_BM + quarter * 32 + lane;
            const int n0 = tn * BN;
            float* sb = sbias + as * BN;
            if (p.bias) {
                // buffer `as` was last read two tiles ago; every epilogue warp has passed the barrier of the tile in
                // between, so it is free. One coalesced load per tile replaces 32 dependent L1 round trips per thread.
                for (int i = et_all; i < BN; i += 256) sb[i] = (n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
                asm volatile("bar.sync 3, 256;" ::: "memory");
            }
            auto load_side = [&](int c, uint4 (&r)[4]) {
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    const int col = n0 + c * 32 + g8 * 8;
                    r[g8] = (side_ptr != nullptr && row < p.M && col < p.N)
                                ? ld_nc_v4(side_ptr + static_cast<size_t>(row) * p.ldr + col)
                                : make_uint4(0, 0, 0, 0);
                }
            };
            const int rem = p.N - n0;
            const int nc_valid = rem >= BN ? NC : (rem + 31) / 32;  // chunks that hold at least one valid column
            const int r_in_tile = quarter * 32 + lane;
            mbar_wait(&tfull[as], aph);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
#pragma unroll 1
            for (int c0 = 2 * grp; c0 < nc_valid; c0 += 4) {  // rolled: one slab (two chunk bodies) of code
                const bool two = c0 + 1 < nc_valid;
                uint32_t v0[32], v1[32];
                uint4 side0[4], side1[4];
                tmem_ld_32x32(t_addr + c0 * 32, v0);
                if (two) tmem_ld_32x32(t_addr + (c0 + 1) * 32, v1);
                load_side(c0, side0);
                if (two) load_side(c0 + 1, side1);
                if (p.tma_store) {
                    // this group's slab buffer must have been drained by the TMA store that last read it
                    if (et == 0) tma_store_wait_read<0>();
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                }
                tmem_ld_wait();
                if (p.debug & 1) continue;
                epilogue_chunk<EPI>(p, v0, side0, sb + c0 * 32, row, n0 + c0 * 32, alpha, st_main, 0, r_in_tile);
                if (two) epilogue_chunk<EPI>(p, v1, side1, sb + (c0 + 1) * 32, row, n0 + (c0 + 1) * 32, alpha, st_main, 1, r_in_tile);
                if (p.tma_store) {
                    fence_proxy_async_smem();
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (et == 0) {
                        tma_store_2d(&tmC, st_main, n0 + (c0 >> 1) * 64, tm * GEMM_BM);
                        tma_store_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
            if (++as == 2) as = 0, aph ^= 1;
        }
        if (p.tma_store && et == 0) tma_store_wait_all<0>();  // smem must outlive the bulk stores reading it
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
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
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu pitch=%llu box=%ux%u", (int)r,
                                       (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems, box_inner, box_outer);
    return 0;
}


template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmAux,
                       const GemmParams& p, cudaStream_t st) {
    auto kern = gemm_kernel<BN, STAGES, A_MN, B_MN, EPI>;
    constexpr size_t smem = gemm_smem_bytes<BN, STAGES>();
    static bool configured = false;  // benign race: attribute set is idempotent
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(-2, "gemm: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    const int tiles = p.tiles_m * p.tiles_n;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    kern<<<grid, GEMM_THREADS, smem, st>>>(tmA, tmB, tmC, tmAux, p);
    return check_launch("gemm_bf16");
}

template <int BN, int STAGES>
static int dispatch_major(const b200_gemm_args* a, GemmParams& p, cudaStream_t st) {
    p.tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
    p.tiles_n = (p.N + BN - 1) / BN;
    CUtensorMap tmA, tmB;
    int rc;
    if (!a->a_mn) rc = make_tmap_bf16_2d(&tmA, a->A, a->K, a->M, a->lda, GEMM_BK, GEMM_BM);
    else          rc = make_tmap_bf16_2d(&tmA, a->A, a->M, a->K, a->lda, 64, GEMM_BK);
    if (rc) return rc;
    if (!a->b_mn) rc = make_tmap_bf16_2d(&tmB, a->B, a->K, a->N, a->ldb, GEMM_BK, BN);
    else          rc = make_tmap_bf16_2d(&tmB, a->B, a->N, a->K, a->ldb, 64, GEMM_BK);
    if (rc) return rc;
    CUtensorMap tmC = tmA, tmAux = tmA;  // placeholders when the TMA-store path is off
    if (p.tma_store) {
        if ((rc = make_tmap_bf16_2d(&tmC, a->C, a->N, a->M, a->ldc, 64, GEMM_BM))) return rc;
        if (a->aux_out && (rc = make_tmap_bf16_2d(&tmAux, a->aux_out, a->N, a->M, a->ldc, 64, GEMM_BM))) return rc;
    }
    if (a->gelu) {
        if (a->a_mn || a->b_mn) return fail(-1, "gemm: the GELU epilogue is built for the forward layout (a_mn=0, b_mn=0) only");
        return launch_gemm<BN, STAGES, false, false, 1>(tmA, tmB, tmC, tmAux, p, st);
    }
    if (a->dgelu_in) {
        if (a->a_mn || !a->b_mn) return fail(-1, "gemm: the dGELU epilogue is built for the dgrad layout (a_mn=0, b_mn=1) only");
        return launch_gemm<BN, STAGES, false, true, 2>(tmA, tmB, tmC, tmAux, p, st);
    }
    if (!a->a_mn && !a->b_mn) return launch_gemm<BN, STAGES, false, false, 0>(tmA, tmB, tmC, tmAux, p, st);
    if (!a->a_mn && a->b_mn) return launch_gemm<BN, STAGES, false, true, 0>(tmA, tmB, tmC, tmAux, p, st);
    if (a->a_mn && a->b_mn) return launch_gemm<BN, STAGES, true, true, 0>(tmA, tmB, tmC, tmAux, p, st);
    return launch_gemm<BN, STAGES, true, false, 0>(tmA, tmB, tmC, tmAux, p, st);
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
    p.M = a->M, p.N = a->N, p.K = a->K;
    p.C = a->C, p.ldc = a->ldc, p.c_fp32 = a->c_fp32, p.accumulate = a->accumulate;
    p.bias = a->bias;
    p.residual = static_cast<const __nv_bfloat16*>(a->residual);
    p.ldr = a->ldr;
    p.gelu = a->gelu;
    p.alpha_dev = a->alpha_dev;
    p.aux_out = static_cast<__nv_bfloat16*>(a->aux_out);
    p.dgelu_in = static_cast<const __nv_bfloat16*>(a->dgelu_in);
    static const int dbg = getenv("B200_GEMM_DEBUG") ? atoi(getenv("B200_GEMM_DEBUG")) : 0;
    p.debug = dbg;
    p.tma_store = (!a->c_fp32 && !a->accumulate && !(dbg & 64)) ? 1 : 0;
    B200_REQUIRE(!a->aux_out || a->gelu, "gemm: aux_out is the pre-GELU output and needs gelu=1");
    cudaStream_t st = as_stream(stream);
    // 128 x 256 tiles when N is wide enough to fill them; 128 x 128 otherwise
    if (a->N > 128) return dispatch_major<256, 4>(a, p, st);
    return dispatch_major<128, 6>(a, p, st);
}
