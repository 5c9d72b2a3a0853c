/* libb200pt — C ABI of the B200 (sm_100a) kernels behind the data-parallel pretraining step.
 *
 * The reference (tttyuntian/multimodal_llm_pretraining) is pure Python: the arithmetic of its hot path lives in
 * transformers / torch / deepspeed. There is therefore no native FFI in the reference to mirror; each
 * entry point below names the third-party op it replaces on the path that `src/benchmarking/utils.py:61-80`
 * (manual_training_step / manual_optimization_step) drives.  "HF:" = transformers/ (v5.5.0 in this image).
 *
 * Conventions:
 *   - every function returns 0 on success, <0 on error; b200_last_error() gives a thread-local message;
 *   - no C++ exceptions cross the ABI; no torch types; plain device pointers + sizes;
 *   - the caller owns every buffer including workspaces; kernels never allocate and never synchronise the device;
 *   - everything is enqueued on the cudaStream_t that is passed in (as void*);
 *   - activations/weights feeding tensor cores are bf16; statistics, biases, LayerNorm affine, gradients of
 *     parameters and optimizer state are fp32.
 */
#ifndef B200PT_H
#define B200PT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PT_ABI_VERSION 8

typedef void* b200_stream_t; /* cudaStream_t */

int b200_abi_version(void);
/* 16-bit element type of the loaded library: 0 = bf16 (libb200pt.so), 1 = IEEE fp16 (libb200pt_fp16.so — the SAME sources and
 * the SAME entry points compiled with -DB200_ELEM_FP16; wherever this header says "bf16" that build reads and writes fp16:
 * the reference trains every Pythia but 1b and RoBERTa in fp16, src/models/pythia.py:33-41, src/models/roberta.py:29-30).
 * fp32 statistics / gradients of parameters / optimizer state are the same in both. */
int b200_elem_dtype(void);
const char* b200_last_error(void);
/* One-time per-process/device setup: resolves cuTensorMapEncodeTiled, raises dynamic-smem limits. Idempotent. */
int b200_init(int device);

/* ---------------------------------------------------------------- LayerNorm
 * Replaces nn.LayerNorm fwd/bwd (HF:models/gpt_neox/modeling_gpt_neox.py:251-252,269,279,376).
 * x,y,dy,dx bf16 [rows, cols]; gamma/beta fp32; mean/rstd fp32 [rows].
 * gamma2/beta2/y2 (nullable): second affine on the SAME normalised x — GPT-NeoX's parallel-residual block applies
 * input_layernorm and post_attention_layernorm to the same tensor (modeling_gpt_neox.py:269,279). */
int b200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, const float* gamma2,
                       const float* beta2, void* y2, float* mean, float* rstd, int rows, int cols, float eps,
                       b200_stream_t stream);
/* dx = LN'(dy[,dy2]) (+ dres); dgamma/dbeta (+ second affine) are ACCUMULATED (+=) in fp32.
 * workspace: b200_layernorm_bwd_workspace_bytes(cols, n_affine) bytes of scratch. */
size_t b200_layernorm_bwd_workspace_bytes(int cols, int n_affine);
int b200_layernorm_bwd(const void* x, const float* mean, const float* rstd, const float* gamma, const void* dy,
                       const float* gamma2, const void* dy2, const void* dres, void* dx, float* dgamma, float* dbeta,
                       float* dgamma2, float* dbeta2, void* workspace, size_t workspace_bytes, int rows, int cols,
                       b200_stream_t stream);

/* ---------------------------------------------------------------- GELU (exact erf; HF:activations.py GELUActivation) */
int b200_gelu_fwd(const void* x, void* y, size_t n, b200_stream_t stream);
int b200_gelu_bwd(const void* x, const void* dy, void* dx, size_t n, b200_stream_t stream);

/* ---------------------------------------------------------------- Rotary embedding
 * In-place NeoX rotate-half on the first `rot` dims of q and k inside a packed qkv buffer
 * [B*S, nh, 3, hd] (HF:modeling_gpt_neox.py:119-159,211-222). cos/sin fp32 [S, rot/2].
 * inverse=1 applies the transpose rotation (backward). */
int b200_rope_qk_inplace(void* qkv, const float* cos_tab, const float* sin_tab, int B, int S, int nh, int hd, int rot,
                         int inverse, b200_stream_t stream);

/* ---------------------------------------------------------------- Embedding
 * out[t,:] = table[ids[t],:] (HF:modeling_gpt_neox.py:345). bwd: dtable[ids[t],:] += dout[t,:] in fp32.
 * An id outside [0, vocab) traps the kernel with a message (torch raises a device-side assert); vocab <= 0 = unchecked. */
int b200_embedding_fwd(const int64_t* ids, const void* table, void* out, int T, int h, int vocab,
                       b200_stream_t stream);
int b200_embedding_bwd(const int64_t* ids, const void* dout, float* dtable, int T, int h, int vocab,
                       b200_stream_t stream);
/* Same, for nn.Embedding(padding_idx=...): tokens whose id == padding_idx contribute nothing (RoBERTa word and position
 * tables, HF:models/roberta/modeling_roberta.py:61,74-76). */
int b200_embedding_bwd_padding(const int64_t* ids, const void* dout, float* dtable, int T, int h, int64_t padding_idx,
                               b200_stream_t stream);
/* out[t,:] = LN-less sum of up to three bf16 table rows (RoBERTa word + position + token-type,
 * HF:models/roberta/modeling_roberta.py:56-144); ids1/ids2 may be NULL. */
int b200_embedding3_fwd(const int64_t* ids0, const void* table0, int vocab0, const int64_t* ids1, const void* table1,
                        const int64_t* ids2, const void* table2, void* out, int T, int h, b200_stream_t stream);

/* RoBERTa position ids: pos = cumsum(ids != pad) * (ids != pad) + pad along each row
 * (create_position_ids_from_input_ids, HF:models/roberta/modeling_roberta.py:146-159). int64 in / out [B, S]. */
int b200_roberta_position_ids(const int64_t* ids, int64_t* pos, int B, int S, int64_t pad_id, b200_stream_t stream);

/* ---------------------------------------------------------------- Dropout (nn.Dropout on the RoBERTa path,
 * HF:models/roberta/modeling_roberta.py:65,339,397): out = x * mask / (1 - p) (+ residual), bf16, n elements (n % 8 == 0).
 * The mask is a pure function of (seed, element index), so backward calls the same entry point on the gradient with the
 * same seed (residual = NULL) instead of storing a mask. p is quantised to 1/65536. */
int b200_dropout(const void* x, const void* residual, void* out, size_t n, float p, uint64_t seed, b200_stream_t stream);

/* ---------------------------------------------------------------- Cross entropy over the vocabulary
 * Replaces ForCausalLMLoss / fixed_cross_entropy (HF:loss/loss_utils.py:28-67): fp32 log-softmax + NLL, mean over
 * labels != ignore_index. logits bf16 [T, ld] (V valid columns) are overwritten IN PLACE by
 * dlogits = (softmax - onehot) / n_valid when write_grad != 0.  row_loss fp32 [T]; n_valid device int (output of
 * b200_count_valid); loss_out device fp32 scalar. */
/* b200_count_valid also validates: a label that is neither ignore_index nor in [0, V) traps the kernel with a message
 * (torch's cross_entropy raises a device-side assert for it); V <= 0 disables the check.
 * grad_scale_dev (nullable device fp32 scalar): dlogits are multiplied by it — the fp16 loss scale is applied HERE, where
 * (softmax - onehot) / n_valid (~1e-5) would otherwise underflow the 16-bit gradient. */
int b200_count_valid(const int64_t* labels, int T, int64_t ignore_index, int V, int* n_valid, b200_stream_t stream);
int b200_cross_entropy(void* logits, const int64_t* labels, float* row_loss, const int* n_valid, int T, int V,
                       int64_t ld, int64_t ignore_index, int write_grad, const float* grad_scale_dev, b200_stream_t stream);
int b200_mean_loss(const float* row_loss, const int* n_valid, int T, float* loss_out, b200_stream_t stream);

/* ---------------------------------------------------------------- Column sums (bias gradients)
 * out[c] += s * sum_r x[r, c]  (x bf16 [rows, ld]); replaces the bias-grad reduction autograd performs for nn.Linear.
 * out2 (nullable, fp32 [cols]) receives the same increment: two biases fed by one upstream gradient (attention.dense and
 * mlp.dense_4h_to_h of a parallel-residual GPT-NeoX block). scale_dev (nullable device fp32 scalar) = s, else 1. */
size_t b200_colsum_workspace_bytes(int cols);
int b200_colsum_bf16(const void* x, int rows, int cols, int64_t ld, float* out, float* out2, const float* scale_dev,
                     void* workspace, size_t workspace_bytes, b200_stream_t stream);

/* ---------------------------------------------------------------- GEMM on tcgen05 / TMEM / TMA
 * C[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )  — replaces nn.Linear fwd / dgrad / wgrad (cuBLAS in the reference
 * stack; HF:modeling_gpt_neox.py:41-42,200-201,464).
 *   a_mn = 0: A stored [M, K] row-major (lda = row pitch in elements);  a_mn = 1: A stored [K, M] row-major.
 *   b_mn = 0: B stored [N, K] row-major;                                b_mn = 1: B stored [K, N] row-major.
 *   fwd   y = x W^T      : a_mn 0, b_mn 0 (W [out,in])
 *   dgrad dx = dy W      : a_mn 0, b_mn 1
 *   wgrad dW = dy^T x    : a_mn 1, b_mn 1, c_fp32 = 1, accumulate = 1
 * epilogue order: acc*alpha -> +bias[n] -> gelu -> dropout -> +residual[m,n] -> (+C if accumulate) -> store (bf16 or fp32).
 * aux_out (nullable, bf16 [M,N], ld = ldc): receives the value BEFORE gelu (pre-activation kept for backward).
 * dgelu_in (nullable, bf16 [M,N], ld = ldr): value is multiplied by gelu'(dgelu_in[m,n]) (fused dGELU for dgrad). */
typedef struct b200_gemm_args {
    int M, N, K;
    const void* A;
    int64_t lda;
    int a_mn;
    const void* B;
    int64_t ldb;
    int b_mn;
    void* C;
    int64_t ldc;
    int c_fp32;
    int accumulate;
    const float* bias;      /* fp32 [N] or NULL */
    const void* residual;   /* bf16 [M, ldr] or NULL */
    int64_t ldr;
    int gelu;
    const float* alpha_dev; /* device fp32 scalar or NULL (=1) */
    void* aux_out;          /* bf16 or NULL */
    const void* dgelu_in;   /* bf16 or NULL */
    /* fused nn.Dropout in front of the residual add (RoBERTa: LN(dropout(dense(x)) + residual), HF:models/roberta/
     * modeling_roberta.py:339-341,397-399): C = dropout(alpha * A B^T + bias) + residual with b200_dropout's mask of
     * (dropout_seed, flat element index m * ldc + n), so b200_dropout(grad, seed) is its backward. Needs `residual`;
     * dropout_p = 0 is off. */
    float dropout_p;
    uint64_t dropout_seed;
    /* dGELU dgrad only (dgelu_in != NULL), nullable fp32 [N]: colsum_out[n] += sum_m C[m, n] (of the fp32 values before the
     * 16-bit rounding) — the bias gradient of the Linear in front of the GELU (mlp.dense_h_to_4h.bias, HF:modeling_gpt_neox.py:
     * 41-47), reduced in the epilogue instead of a separate b200_colsum_bf16 pass over the [M, N] gradient. */
    float* colsum_out;
} b200_gemm_args;
int b200_gemm_bf16(const b200_gemm_args* args, b200_stream_t stream);

/* ---------------------------------------------------------------- Flash attention on tcgen05
 * Replaces F.scaled_dot_product_attention(is_causal=True) (HF:integrations/sdpa_attention.py:92-101) and RoBERTa's
 * eager softmax attention (HF:models/roberta/modeling_roberta.py:162-187) with causal=0.
 * q,k,v,o,do,dq,dk,dv: bf16; token t = b*S+s, head hh element d at  ptr[t*row_stride + hh*head_stride + d].
 * lse, delta: fp32 [B, H, S].  D in {64, 80, 128, 256} (80 = Pythia-2.8b: read through zero-padding 3-D tensor maps and
 * computed as 128).  scale = head_dim^-0.5 of the REAL head_dim. */
typedef struct b200_attn_args {
    int B, S, H, D;
    int causal;
    float scale;
    const void* q;
    const void* k;
    const void* v;
    int64_t qkv_row_stride;   /* elements between tokens in q/k/v buffers */
    int64_t qkv_head_stride;  /* elements between heads */
    void* o;                  /* fwd out / bwd in */
    int64_t o_row_stride;
    int64_t o_head_stride;
    float* lse;               /* fwd out / bwd in */
    /* backward only */
    const void* d_o;          /* same layout as o */
    float* delta;             /* scratch [B,H,S] */
    void* dq;
    void* dk;
    void* dv;                 /* same layout as q/k/v (dqkv_* strides) */
    int64_t dqkv_row_stride;
    int64_t dqkv_head_stride;
    /* backward, head_dim 256 only (nullable): two bf16 scratch buffers [B*H, S, S]. When both are given and S % 256 == 0
     * the dQ pass also stores its P and dS tiles there and dV = P^T dO, dK = scale dS^T Q run as batched causal GEMMs
     * on the CTA-pair tensor-core engine; otherwise dK/dV are recomputed in a second fused pass. The causal path only
     * ever writes the lower block triangle; the buffers need no initialisation. */
    void* p_scratch;
    void* ds_scratch;
    /* attention-probability dropout (RoBERTa: nn.Dropout on the softmax output, HF:models/roberta/modeling_roberta.py:209,236).
     * The mask is a pure function of (dropout_seed, b*H+h, query, key); forward and backward must be given the same
     * (dropout_p, dropout_seed). 0 = off. p is quantised to 1/65536. */
    float dropout_p;
    uint64_t dropout_seed;
} b200_attn_args;
int b200_attention_fwd(const b200_attn_args* args, b200_stream_t stream);
int b200_attention_bwd(const b200_attn_args* args, b200_stream_t stream);

/* ---------------------------------------------------------------- Optimizer (fused multi-tensor Adam / AdamW)
 * Replaces torch.optim.Adam foreach path and DeepSpeed FusedAdam (src/models/pythia.py:43-67, src/train.py:157-167).
 * All parameters live in one flat fp32 buffer; `chunks` partition the part this rank updates into pieces that each
 * belong to one param group. p/g are indexed by absolute element offset; m/v by chunk_state[c] + i when chunk_state is
 * given (a ZeRO-1 rank keeps only the moments of the slices it owns, packed back to back), else by
 * (offset - state_base).
 *   adamw_mode = 0: L2 (torch.optim.Adam):  g += wd*p          adamw_mode = 1: decoupled (p *= 1 - lr*wd)
 *   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
 * grad_scale_dev (nullable): device fp32 scalar multiplied into every gradient first (clip coefficient / unscale).
 * p_bf16 (nullable): bf16 shadow of p written in the same pass. zero_grad != 0 clears g in the same pass. */
typedef struct b200_adam_group {
    float lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2;
    int adamw_mode;
} b200_adam_group;
#define B200_ADAM_MAX_GROUPS 8
int b200_adam_step(float* p, float* g, float* m, float* v, void* p_bf16, int64_t state_base,
                   const int64_t* chunk_start, const int32_t* chunk_len, const int32_t* chunk_group,
                   const int64_t* chunk_state, int n_chunks, const b200_adam_group* groups, int n_groups,
                   const float* grad_scale_dev, int zero_grad, const int* skip_flag_dev, int g_packed, int p_packed,
                   b200_stream_t stream);
/* skip_flag_dev (nullable device int): when *skip_flag_dev != 0 the step leaves p, m, v and the shadow untouched (an fp16
 * overflow step, torch.amp.GradScaler.step semantics); g is still cleared when zero_grad != 0.
 * g_packed != 0: g is a packed shard buffer indexed like m / v (chunk_state[c] + i) instead of the full flat gradient buffer —
 * ZeRO-2 (src/train.py:172-181), where a rank only holds the reduced gradients of the slices it owns.
 * p_packed != 0: p likewise is a packed shard buffer (the fp32 master exists only on the owning rank, 4 B/param/W — DeepSpeed's
 * ZeRO partition of the fp32 weights); p_bf16 stays the full replicated 16-bit copy, indexed by absolute offset. */

/* out[0] += sum(x[i]^2), DETERMINISTIC (fixed per-block partials into `workspace`, then a fixed-order final sum; no atomics):
 * replicas holding identical gradients get bit-identical norms. workspace: b200_sumsq_workspace_bytes() bytes. */
size_t b200_sumsq_workspace_bytes(void);
int b200_sumsq(const float* x, size_t n, float* out, void* workspace, size_t workspace_bytes, b200_stream_t stream);
/* Same over the chunks of an optimizer chunk table (the slices a ZeRO rank owns): one block per chunk writes partials[c]
 * (fp32 [n_chunks], caller-owned), then the fixed-order final sum is added to out[0]. */
int b200_sumsq_chunks(const float* x, const int64_t* chunk_start, const int32_t* chunk_len, int n_chunks, float* out,
                      float* partials, b200_stream_t stream);
/* norm_out = sqrt(sumsq) / loss_scale; coef_out = min(1, max_norm / (norm + 1e-6)) / loss_scale
 * (torch.nn.utils.clip_grad_norm_ semantics on the UNSCALED gradients; max_norm <= 0 gives min(..) = 1).
 * loss_scale_dev (nullable) = the fp16 loss scale the gradients carry (NULL = 1).
 * found_inf_out (nullable device int) = 1 when sumsq is inf / nan, i.e. some gradient overflowed (the overflow check of
 * GradScaler.unscale_ / DeepSpeed has_overflow: a single non-finite element makes the sum non-finite), else 0. */
int b200_clip_coef(const float* sumsq, float max_norm, const float* loss_scale_dev, float* norm_out, float* coef_out,
                   int* found_inf_out, b200_stream_t stream);
/* Dynamic loss scale, all state on the device (no host sync in the step): on overflow the hysteresis counter is decremented
 * and, once spent (or hysteresis <= 1), scale = max(scale * backoff_factor, min_scale); after growth_interval consecutive
 * clean steps scale *= growth_factor and the hysteresis is refilled. torch.amp.GradScaler = (2, 0.5, 2000, -, 1);
 * the reference's DeepSpeed config (src/train.py:143-150) = (2, 0.5, 1000, min 1, hysteresis 2, initial 2^16). */
int b200_loss_scale_update(float* scale, int* growth_tracker, int* hysteresis_left, const int* found_inf,
                           float growth_factor, float backoff_factor, int growth_interval, float min_scale,
                           int hysteresis, b200_stream_t stream);
int b200_cast_f32_to_bf16(const float* src, void* dst, size_t n, b200_stream_t stream);
int b200_scale_f32(float* x, size_t n, const float* scale_dev, float scale_host, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200PT_H */
