/*
 * moegan_sm100.h - C ABI of libmoegan_sm100.so: the B200 (sm_100a) hot path of MoE-GAN's
 * Bayesian-routed Mixture-of-Experts block (forward + backward).
 *
 * The reference (shmurthy08/MoE-GAN_CPSC541) is pure Python/PyTorch and has no FFI; the boundary it
 * offers for this path is the nn.Module `SparseMoE` (moegan/t2i_moe_gan.py:426-491).  Each entry point
 * below replaces one cluster of eager torch ops inside that module; the citation after "replaces:"
 * is the reference code it stands in for (paths relative to the reference root).  INTEGRATION.md shows
 * the ctypes binding and the two-line swap a reference maintainer would add.
 *
 * Conventions
 *  - Plain pointers and sizes only.  Every pointer is a DEVICE pointer unless it says "host".
 *  - The caller owns all memory (including workspaces); the library never allocates or frees device
 *    memory and keeps no state besides the last error string and cached device attributes.
 *  - Every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns without
 *    synchronising.  Return value: 0 = success, negative = moegan_status.  Nothing throws or exits.
 *  - Tokens are "token-major": x is [T, d] row-major with T = B * tokens_per_sample, token order
 *    (b, h, w) - exactly `x.permute(0,2,3,1).contiguous().view(-1, C)` (t2i_moe_gan.py:455).
 *  - precision: 0 = bf16 operands / fp32 accumulate on tcgen05;  1 = "fp32": every operand is split
 *    into bf16 hi + lo halves and each GEMM runs three tcgen05 passes (hi*hi + hi*lo + lo*hi),
 *    accurate to ~2^-16 relative - this is the mode that meets the 1e-4 fp32 parity bar.
 *    "bf16 operand" buffers are [rows, S*cols] bf16 with S = 1 (precision 0) or 2 (hi | lo halves).
 *  - Row-permuted ("dispatched") buffers have m_cap rows (moegan_m_cap()).  Each expert's rows are
 *    contiguous, in ascending token order, and start at a multiple of MOEGAN_ROW_ALIGN (256) rows (one cta_group::2
 *    MMA tile); padding rows are zero.
 */
#ifndef MOEGAN_SM100_H
#define MOEGAN_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; these are its exports */
#endif

#define MOEGAN_ABI_VERSION 6
#define MOEGAN_ROUTER_HID 128 /* t2i_moe_gan.py:279,286 */
#define MOEGAN_MAX_EXPERTS 32
#define MOEGAN_MAX_TOPK 8     /* sparse (non-dense) mode */
#define MOEGAN_CHUNK 32       /* tokens per router CTA / permutation chunk */
#define MOEGAN_ROW_ALIGN 256  /* every expert segment of a dispatched buffer starts at a multiple of this many rows */
#define MOEGAN_KL_PARTS 256   /* scratch partial sums behind kl_out */
#define MOEGAN_KL_OUT_FLOATS (2 + MOEGAN_KL_PARTS)

typedef enum {
    MOEGAN_OK = 0,
    MOEGAN_ERR_BAD_SHAPE = -1,
    MOEGAN_ERR_NULL_POINTER = -2,
    MOEGAN_ERR_MISALIGNED = -3,
    MOEGAN_ERR_UNSUPPORTED = -4,
    MOEGAN_ERR_CUDA = -5,
    MOEGAN_ERR_DRIVER = -6
} moegan_status;

typedef struct {
    int32_t T;                 /* tokens = B * tokens_per_sample */
    int32_t d;                 /* feature dim (multiple of 128) */
    int32_t Dt;                /* text / style dim (multiple of 32) */
    int32_t E;                 /* experts, <= MOEGAN_MAX_EXPERTS */
    int32_t k;                 /* experts per token; k == E in dense mode, 1 in hard mode */
    int32_t B;                 /* samples */
    int32_t tokens_per_sample; /* H * W */
    int32_t precision;         /* 0 bf16, 1 fp32 (3-pass bf16 split) */
    int32_t dense;             /* 1: all experts see all tokens, gate = prob (reference training path, :465-470) */
    int32_t hard;              /* 1: eval path, one-hot probs, top-1, gate 1.0 (:392-400, :471-483) */
    int32_t m_cap;             /* row capacity of dispatched buffers, = moegan_m_cap() */
} moegan_shape;

/* ---- library ------------------------------------------------------------------------------- */
int moegan_version(void);
const char* moegan_last_error_string(void); /* host string, valid until the next failing call on this thread */
/* rows to allocate for dispatched buffers: dense E*ceil256(T); sparse ceil256(T*k) + 256*E */
int64_t moegan_m_cap(int32_t T, int32_t E, int32_t k, int32_t dense);
/* int32 elements of the dispatch plan array */
int64_t moegan_plan_ints(int32_t m_cap);
/* bytes of scratch moegan_router_bwd needs */
int64_t moegan_router_bwd_workspace_bytes(const moegan_shape* s);

/* ---- router --------------------------------------------------------------------------------- */
/* replaces: BayesianRouter.reparameterize x3 (t2i_moe_gan.py:302-333, called :354-356) and
 * BayesianRouter.kl_divergence (:405-423).  sampling=0 copies mu (eval, :357-361).
 * Wf [d,128], Wt [Dt,128], Wc [256,E] fp32.  Wf_op [d, S*128], Wt_op [Dt, S*128] (bf16 operand copies
 * for the backward GEMMs; may be NULL).  kl_out [MOEGAN_KL_OUT_FLOATS]: [0] = nan_to_num+clamp[0,120] value,
 * [1] = raw sum, [2..] scratch partials. */
int moegan_router_sample_kl_fwd(const moegan_shape* s,
                                const float* mu_f, const float* rho_f, const float* eps_f,
                                const float* mu_t, const float* rho_t, const float* eps_t,
                                const float* mu_c, const float* rho_c, const float* eps_c,
                                int32_t sampling, float* Wf, float* Wt, float* Wc,
                                void* Wf_op, void* Wt_op, float* kl_out, void* stream);

/* replaces: text half of the projections (t2i_moe_gan.py:365,368,371) computed ONCE PER SAMPLE instead
 * of on the materialised [T,Dt] expansion (:456).  tp [B,128] = nan_to_num(w) @ Wt;  tl [B,E] = tp @ Wc[128:]. */
int moegan_router_text_fwd(const moegan_shape* s, const float* w, const float* Wt, const float* Wc,
                           float* tp, float* tl, void* stream);

/* replaces: BayesianRouter.forward projections, temperature, clamps, softmax, clamp+renorm, eval one-hot
 * (t2i_moe_gan.py:338-402) plus top-k gating and the per-chunk expert histogram / prob column sums.
 * fp [T,128] (may be NULL when no backward is needed), raw [T,E] logits before temperature,
 * probs [T,E] (one-hot when s->hard), topk_idx [T,k] int32 (descending prob, ties -> lowest index),
 * gates [T,k], chunk_counts [ceil(T/32), E] int32, chunk_colsum [ceil(T/32), E] fp32 (soft probs). */
int moegan_router_fwd(const moegan_shape* s, const float* x, const float* Wf, const float* Wc, const float* tl,
                      const float* temperature, float annealing_factor,
                      float* fp, float* raw, float* probs, int32_t* topk_idx, float* gates,
                      int32_t* chunk_counts, float* chunk_colsum, void* stream);

/* ---- dispatch ------------------------------------------------------------------------------- */
/* replaces: `expert_indices == i` masks / torch.any host syncs (t2i_moe_gan.py:473-479) with a device-side
 * histogram + exclusive scan.  plan: see DESIGN.md (block->expert map, padded expert offsets, counts).
 * chunk_base [ceil(T/32), E] int32: first row (relative to the expert segment) of each chunk.
 * load [E] = column sums of the soft probs (input of moe_balance_loss, t2i_moe_gan.py:973). */
int moegan_dispatch_plan(const moegan_shape* s, const int32_t* chunk_counts, const float* chunk_colsum,
                         int32_t* plan, int32_t* chunk_base, float* load, void* stream);

/* replaces: boolean-mask gather `x_reshaped[mask]` (t2i_moe_gan.py:481).  Stable permutation.
 * x_op [m_cap, S*d] gathered bf16 operand rows; x_copy [T, S*d] un-permuted operand copy (may be NULL);
 * row_token [m_cap] (-1 on padding rows), row_gate [m_cap], token_rows [T,k]. */
int moegan_dispatch_gather(const moegan_shape* s, const float* x, const int32_t* topk_idx, const float* gates,
                           const int32_t* plan, const int32_t* chunk_base,
                           void* x_op, void* x_copy, int32_t* row_token, float* row_gate, int32_t* token_rows,
                           void* stream);

/* dense mode: no permutation, only fp32 -> bf16 operand conversion.  x_copy [T, S*d]. */
int moegan_convert_tokens(const moegan_shape* s, const float* x, void* x_copy, void* stream);

/* replaces: the per-call autocast weight casts of nn.Linear (t2i_moe_gan.py:257,259).  Host arrays of E
 * device pointers to the per-expert fp32 parameters (state_dict layout) -> stacked operands
 * W1_op [E*4d, S*d], W2_op [E*d, S*4d] and stacked fp32 biases b1s [E,4d], b2s [E,d]. */
int moegan_prep_weights(const moegan_shape* s, const float* const* w1_host_ptrs, const float* const* b1_host_ptrs,
                        const float* const* w2_host_ptrs, const float* const* b2_host_ptrs,
                        void* W1_op, float* b1s, void* W2_op, float* b2s, void* stream);

/* ---- experts -------------------------------------------------------------------------------- */
/* replaces: SparseExpertFFN.forward for all experts (t2i_moe_gan.py:262-263, loops :467-468 / :476-482):
 * two grouped GEMMs on tcgen05.  a_op = x_op (sparse) or x_copy (dense: shared by all experts).
 * pre [m_cap,4d]: what the backward needs from the first GEMM - precision 0: bf16 gelu'(x W1^T + b1);
 * precision 1: fp32 pre-activation; NULL for a forward-only call.  h_op [m_cap, S*4d] = gelu(.), y [m_cap,d] (bf16 / fp32). */
int moegan_ffn_fwd(const moegan_shape* s, const int32_t* plan, const void* a_op, const void* W1_op, const float* b1s,
                   const void* W2_op, const float* b2s, void* pre, void* h_op, void* y, void* stream);

/* replaces: `combined_output += probs[:, i:i+1] * expert_output` (t2i_moe_gan.py:462,470) and
 * `combined_output[mask] = expert_output` (:483); with `residual` ([T,d] fp32 or bf16, may be NULL) also the
 * `x_spatial + moe_out` of AttentionBlock (:571) in the same pass.  out [T,d] fp32 or bf16. */
int moegan_combine_fwd(const moegan_shape* s, const void* y, const float* gates, const int32_t* token_rows,
                       const void* residual, int32_t residual_is_bf16, void* out, int32_t out_is_bf16, void* stream);

/* ---- backward ------------------------------------------------------------------------------- */
/* autograd of the combine: dy_op[r] = gate * dout[token(r)] (zero on padding rows),
 * dgate_row[r] = <dout[token(r)], y[r]>.  dout [T,d] fp32 or bf16.  gates: [T,k] (dense: the probs, k == E).
 * dy_slab_sums (may be NULL): [m_cap/32, d] fp32 column sums of every 32 consecutive dy rows - moegan_ffn_bwd turns them
 * into the second Linear's bias gradient without reading dy again. */
int moegan_combine_bwd(const moegan_shape* s, const int32_t* plan, const void* dout, int32_t dout_is_bf16, const void* y,
                       const int32_t* row_token, const float* row_gate, const float* gates,
                       void* dy_op, float* dgate_row, float* dy_slab_sums, void* stream);

/* autograd of the experts (B2): dgrad + wgrad grouped GEMMs and bias gradients (db1 from slab sums emitted by the
 * dgrad epilogue; db2 from dy_slab_sums as written by moegan_combine_bwd, or from a pass over dy_op when it is NULL).
 * dpre_op [m_cap, S*4d], dx_rows [m_cap,d] (bf16 / fp32), dW1 [E,4d,d], db1 [E,4d], dW2 [E,d,4d], db2 [E,d] fp32;
 * scratch: moegan_ffn_bwd_scratch_bytes(), 256-byte aligned. */
int moegan_ffn_bwd(const moegan_shape* s, const int32_t* plan, const void* a_op, const void* W1_op, const void* W2_op,
                   const void* pre, const void* h_op, const void* dy_op, const float* dy_slab_sums, void* dpre_op, void* dx_rows,
                   float* dW1, float* db1, float* dW2, float* db2, void* scratch, int64_t scratch_bytes, void* stream);
/* bytes of scratch moegan_ffn_bwd needs (split-K partials of under-filled weight gradients, slab sums of dpre) */
int64_t moegan_ffn_bwd_scratch_bytes(const moegan_shape* s);

/* un-permute + sum of the k expert contributions per token, plus the router's dx:  dx [T,d] fp32. */
int moegan_uncombine_bwd(const moegan_shape* s, const void* dx_rows, const int32_t* token_rows,
                         const float* dx_router, float* dx, void* stream);

/* autograd of the router (B3-B6): gates/probs -> logits -> projections -> sampled weights -> (mu, rho),
 * temperature, x and w; plus the KL gradient.  dprobs_ext [T,E] and dload [E] may be NULL.
 * grads: dmu/drho x3, dtemperature[1], dx_router [T,d], dw [B,Dt].  x_copy [T,S*d], Wf_op, Wt_op as above. */
int moegan_router_bwd(const moegan_shape* s, const void* x_copy, const float* w,
                      const float* Wf, const float* Wt, const float* Wc, const void* Wf_op, const void* Wt_op,
                      const float* fp, const float* tp, const float* raw, const int32_t* topk_idx,
                      const int32_t* token_rows, const float* dgate_row, const float* dprobs_ext, const float* dload,
                      const float* temperature, float annealing_factor, int32_t sampled,
                      const float* mu_f, const float* rho_f, const float* eps_f,
                      const float* mu_t, const float* rho_t, const float* eps_t,
                      const float* mu_c, const float* rho_c, const float* eps_c,
                      const float* kl_out, const float* dkl,
                      float* dmu_f, float* drho_f, float* dmu_t, float* drho_t, float* dmu_c, float* drho_c,
                      float* dtemperature, float* dx_router, float* dw,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* replaces: AuroraGANLoss.moe_balance_loss on the fused column sums (t2i_moe_gan.py:973-1000).
 * fwd: loss[0] = bw * nan_to_num(clamp(E * std(frac)/(mean(frac)+1e-6), 0, 10)), frac = (load+1e-6)/T.
 * bwd: dload [E] = dloss * d(loss)/d(load). */
int moegan_balance_loss_fwd(const float* load, int32_t E, int32_t T, float balance_weight, float* loss, void* stream);
int moegan_balance_loss_bwd(const float* load, int32_t E, int32_t T, float balance_weight, const float* dloss,
                            float* dload, void* stream);

/* ---- whole block in one call ------------------------------------------------------------------ */
/* The same pipeline as the calls above (sample+KL -> text -> router -> plan -> gather/convert -> experts ->
 * combine, and its backward), enqueued by ONE call per direction with every intermediate carved out of ONE
 * caller-owned workspace: this is what SparseMoE.forward uses, so the host cost per step is two FFI calls.
 * Independent sub-chains (weight conversion in the forward, the router's backward in the backward) are forked onto a
 * library-owned side stream (one set per CUDA device, created lazily) and joined before their results are needed
 * (MOEGAN_DISABLE_SIDE_STREAM=1 turns this off).
 * Byte offsets of the intermediates inside the forward workspace (all 256-byte aligned): */
typedef struct {
    int64_t Wf, Wt, Wc, Wf_op, Wt_op, tp, tl, fp, raw, topk_idx, gates, chunk_counts, chunk_colsum, plan, chunk_base,
            x_copy, a_op, row_token, row_gate, token_rows, pre, h_op, y, eps_f, eps_t, eps_c, total;   /* -1 = not present */
} moegan_block_layout;

typedef struct {
    const float* x;             /* [T,d] */
    const float* w;             /* [B,Dt] */
    const float *mu_f, *rho_f, *eps_f, *mu_t, *rho_t, *eps_t, *mu_c, *rho_c, *eps_c, *temperature;
    const void* W1_op; const float* b1s; const void* W2_op; const float* b2s;   /* stacked expert operands (moegan_prep_weights) */
    /* host arrays of E device pointers, or all NULL: when set, the call (re)builds W1_op/b1s/W2_op/b2s from the per-expert
     * fp32 parameters first - on an internal side stream, overlapped with the router and dispatch kernels */
    const float* const* prep_w1; const float* const* prep_b1; const float* const* prep_w2; const float* const* prep_b2;
    float annealing_factor;
    int32_t sampled;            /* 1: training (weights sampled, KL computed) */
    int32_t need_backward;      /* 1: keep what moegan_block_bwd needs */
    void* out;                  /* [T,d] fp32, or bf16 when out_is_bf16 */
    float* probs;               /* [T,E] */
    float* load;                /* [E] */
    float* kl_out;              /* [MOEGAN_KL_OUT_FLOATS] */
    void* workspace; int64_t workspace_bytes;
    const void* residual;       /* [T,d] fp32 / bf16 or NULL: out = residual + moe(x)  (AttentionBlock :571 fused) */
    int32_t residual_is_bf16;
    int32_t out_is_bf16;
    const void* x_op;           /* optional, bf16 mode only: bf16 [T,d] copy of x that the caller already holds (moegan_layernorm_fwd_dual emits
                                 * it together with the fp32 rows, t2i_moe_gan.py:561) - the block then runs no fp32 -> bf16 conversion pass
                                 * over the tokens; pass the same tensor in moegan_block_bwd_args.x_op */
} moegan_block_fwd_args;

typedef struct {
    const void* fwd_workspace;  /* the forward call's workspace (need_backward = 1), untouched since */
    const float* w;
    const float *mu_f, *rho_f, *eps_f, *mu_t, *rho_t, *eps_t, *mu_c, *rho_c, *eps_c, *temperature;   /* eps_* may be NULL: the forward
                                   call keeps its own copy of the eps it consumed inside fwd_workspace */
    const void* W1_op; const void* W2_op; const float* kl_out;
    const void* dout;           /* [T,d] fp32, or bf16 when dout_is_bf16 */
    const float* dkl;           /* [1] or NULL */
    const float* dprobs;        /* [T,E] or NULL */
    const float* dload;         /* [E] or NULL */
    float annealing_factor;
    int32_t sampled;
    float* dx;                  /* [T,d] */
    float* dw;                  /* [B,Dt]  (NULL in hard mode) */
    float *dmu_f, *drho_f, *dmu_t, *drho_t, *dmu_c, *drho_c, *dtemperature;   /* NULL in hard mode */
    float *dW1, *db1, *dW2, *db2;   /* [E,4d,d] [E,4d] [E,d,4d] [E,d] */
    void* workspace; int64_t workspace_bytes;
    int32_t dout_is_bf16;
    int32_t reserved;
    const void* x_op;           /* the x_op the forward call was given (NULL if none) */
} moegan_block_bwd_args;

int64_t moegan_block_fwd_workspace(const moegan_shape* s, int32_t need_backward, moegan_block_layout* layout /* may be NULL */);
int64_t moegan_block_bwd_workspace(const moegan_shape* s);
int moegan_block_fwd(const moegan_shape* s, const moegan_block_fwd_args* a, void* stream);
int moegan_block_bwd(const moegan_shape* s, const moegan_block_bwd_args* a, void* stream);

/* ---- layers around the block ------------------------------------------ */
/* replaces: nn.LayerNorm(dim) of AttentionBlock.norm1-3 (t2i_moe_gan.py:513-515, applied :541,:553,:561).
 * x [rows,d] fp32 or bf16 (x_is_bf16), y [rows,d] fp32 or bf16 (y_is_bf16: the rounding autocast's consumer would apply
 * anyway); mean / rstd [rows] fp32 are what the backward needs.  d: multiple of 128, <= 1024. */
int moegan_layernorm_fwd(const void* x, int32_t x_is_bf16, const float* gamma, const float* beta, int64_t rows, int32_t d,
                         float eps, void* y, int32_t y_is_bf16, float* mean, float* rstd, void* stream);
/* norm3 in front of the MoE block (t2i_moe_gan.py:561), SURVEY 8(f) N1: the same LayerNorm writing BOTH the fp32 rows the router reads
 * (y [rows,d] fp32) and their bf16 GEMM-operand copy (y_op [rows,d] bf16) in one pass over x. */
int moegan_layernorm_fwd_dual(const void* x, int32_t x_is_bf16, const float* gamma, const float* beta, int64_t rows, int32_t d, float eps,
                              float* y, void* y_op, float* mean, float* rstd, void* stream);
/* The residual add in front of norm2 / norm3 fused into the LayerNorm (t2i_moe_gan.py:545 + :553, :557 + :561:
 * `t = t + attn(...)`, then `norm(t)`): sum_out = x + addend (rounded to their common dtype, as the torch add), y = LayerNorm(sum_out);
 * y_op: optional bf16 copy of y (may be NULL; needs y fp32).  d = 128 / 256 / 512.  One pass over the rows and one launch less. */
int moegan_add_layernorm_fwd(const void* x, const void* addend, int32_t x_is_bf16, const float* gamma, const float* beta, int64_t rows,
                             int32_t d, float eps, void* sum_out, void* y, int32_t y_is_bf16, void* y_op, float* mean, float* rstd,
                             void* stream);
/* moegan_layernorm_bwd with dx = LayerNorm-backward(dy) + gadd: gadd [rows,d] in x's dtype is the gradient reaching the LayerNorm's
 * INPUT along the residual path (the summed rows are also consumed by the next residual add).  d = 128 / 256 / 512. */
int moegan_layernorm_bwd_add(const void* dy, int32_t dy_is_bf16, const void* x, int32_t x_is_bf16, const float* gamma, const float* mean,
                             const float* rstd, int64_t rows, int32_t d, const void* gadd, void* dx, float* dgamma_dbeta, void* scratch,
                             int64_t scratch_bytes, void* stream);
int64_t moegan_layernorm_bwd_scratch_bytes(int64_t rows, int32_t d);
/* dy fp32 or bf16; dx [rows,d] in x's dtype; dgamma_dbeta fp32 [2,d] (row 0 = weight gradient, row 1 = bias gradient),
 * deterministic. */
int moegan_layernorm_bwd(const void* dy, int32_t dy_is_bf16, const void* x, int32_t x_is_bf16, const float* gamma,
                         const float* mean, const float* rstd, int64_t rows, int32_t d, void* dx, float* dgamma_dbeta,
                         void* scratch, int64_t scratch_bytes, void* stream);

/* replaces: the modulate / demodulate / LeakyReLU element-wise steps of ModulatedConv + the activation of
 * ModulatedTransformationModule (t2i_moe_gan.py:158-186, :243-245) in "scale the activations" form:
 * out[b,p,c] = leaky_relu(x[b,p,c] * s[b,c], slope)  (slope = 1: no activation).  x / out: NHWC [B,P,C], fp32 or bf16;
 * s [B,C] fp32 or bf16 (s_is_bf16); C: multiple of 128.  Backward: dx and ds[b,c] = sum_p (in s's dtype) in ONE pass over
 * g and x (deterministic). */
int moegan_channel_scale_fwd(const void* x, int32_t is_bf16, const void* s, int32_t s_is_bf16, int32_t B, int32_t P, int32_t C,
                             float slope, void* out, void* stream);
int64_t moegan_channel_scale_bwd_scratch_bytes(int32_t B, int32_t P, int32_t C);
int moegan_channel_scale_bwd(const void* g, const void* x, int32_t is_bf16, const void* s, int32_t s_is_bf16, int32_t B,
                             int32_t P, int32_t C, float slope, void* dx, void* ds, void* scratch, int64_t scratch_bytes,
                             void* stream);

/* replaces: the bias gradient `grad_output.sum(0)` of the nn.Linear projections applied to all T tokens inside
 * AttentionBlock (nn.MultiheadAttention in_proj / out_proj, t2i_moe_gan.py:516-517, used :545,:557): out[c] = sum_r x[r,c].
 * x [rows, cols] fp32 or bf16, cols % 4 == 0; out fp32 [cols]; deterministic. */
int64_t moegan_colsum_scratch_bytes(int64_t rows, int32_t cols);
int moegan_colsum(const void* x, int32_t is_bf16, int64_t rows, int32_t cols, float* out, void* scratch, int64_t scratch_bytes,
                  void* stream);

/* replaces: the learned-offset resampling of ModulatedTransformationModule (t2i_moe_gan.py:219-240):
 * grid = clamp(meshgrid(linspace(-1,1)) + 0.05 * off, -1, 1);  out = F.grid_sample(x, grid, bilinear, zeros,
 * align_corners=False).  x / out / g / dx: NHWC [B,H,W,C] fp32 or bf16; off / doff: fp32 [B,H,W,2] = (x, y) offsets
 * as produced by offset_net (before the 0.05 scale).  C: multiple of 128; H*W <= 256 for the backward. */
int moegan_offset_sample_fwd(const void* x, int32_t is_bf16, const float* off, int32_t B, int32_t H, int32_t W, int32_t C,
                             void* out, void* stream);
int moegan_offset_sample_bwd(const void* g, const void* x, int32_t is_bf16, const float* off, int32_t B, int32_t H,
                             int32_t W, int32_t C, void* dx, float* doff, void* stream);

/* replaces: F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False) at the head of GenerativeBlock.forward
 * (t2i_moe_gan.py:646-648).  x / dx: NHWC [B,H,W,C], out / g: NHWC [B,2H,2W,C], fp32 or bf16; C: multiple of 8.  The backward is
 * a deterministic gather (no atomics). */
int moegan_upsample2x_fwd(const void* x, int32_t is_bf16, int32_t B, int32_t H, int32_t W, int32_t C, void* out, void* stream);
int moegan_upsample2x_bwd(const void* g, int32_t is_bf16, int32_t B, int32_t H, int32_t W, int32_t C, void* dx, void* stream);

/* replaces: the demodulation coefficients of ModulatedConv.forward (t2i_moe_gan.py:165-170), in the "scale the activations" form the
 * host model uses (the reference scales a per-sample copy of the weight and normalises it; this is the same factor per (sample, output
 * channel)):   wsq[o,i] = sum_{ky,kx} w[o,i,ky,kx]^2,   d[b,o] = rsqrt( sum_i s[b,i]^2 * wsq[o,i] + eps ).
 * s [B,Cin] fp32 or bf16 (the modulation Linear's output); w fp32 [Cout,Cin,KH,KW] with ELEMENT strides w_so / w_si / w_sh / w_sw
 * (contiguous or channels-last); wsq [Cout,Cin] and d [B,Cout] fp32 outputs (wsq is kept for the backward); compute_wsq = 0: wsq is an
 * INPUT (the caller computed it for the same weights earlier - the generator runs twice per train step between optimizer steps).
 * Backward: dd [B,Cout] fp32 -> ds [B,Cin] in s's dtype and dw fp32 in w's strides (every element written).  Deterministic. */
int moegan_demod_coef_fwd(const void* s, int32_t s_is_bf16, const float* w, int64_t w_so, int64_t w_si, int64_t w_sh, int64_t w_sw,
                          int32_t B, int32_t Cin, int32_t Cout, int32_t KH, int32_t KW, float eps, int32_t compute_wsq, float* wsq, float* d,
                          void* stream);
int moegan_demod_coef_bwd(const float* dd, const float* d, const void* s, int32_t s_is_bf16, const float* w, int64_t w_so, int64_t w_si,
                          int64_t w_sh, int64_t w_sw, const float* wsq, int32_t B, int32_t Cin, int32_t Cout, int32_t KH, int32_t KW,
                          void* ds, float* dw, void* stream);

/* ---- test / bench hook ---------------------------------------------------------------------- */
/* Plain (ungrouped) GEMM on the same tcgen05 kernel, fp32 out:  C[M,N] = op(A) * op(B).
 * kind 0: A [M,S*K] K-major, B [N,S*K] K-major.   kind 1: A [M,S*K] K-major, B [K,S*N] (N contiguous).
 * kind 2: A [K,S*M] (M contiguous), B [K,S*N].    precision as above.  Used by tests and bench only. */
int moegan_umma_gemm(int32_t kind, int32_t precision, int32_t M, int32_t N, int32_t K,
                     const void* A, const void* B, float* C, void* stream);
/* id of the CUDA-graph capture `stream` is part of (0: not capturing, -1: query failed): lets the host-side weight cache
 * share one conversion between the forward passes of ONE captured step. */
int64_t moegan_stream_capture_id(void* stream);
/* instrumentation for bench.py: number of kernel launches this library has enqueued so far (all threads) */
int64_t moegan_debug_launch_count(void);
/* Profiling for bench.py: CUDA events recorded by the library around its own launches, on the stream of the launch.
 *   level 0 off; 1 = the tcgen05 GEMMs only; 2 = every kernel launcher (not during CUDA-graph capture).
 * GEMM records carry tag = kind * 1000 + BN + 100000 * cta_pair + 10000 * epilogue (< MOEGAN_PROF_KERNEL_BASE); the other
 * launchers tag = MOEGAN_PROF_KERNEL_BASE + MOEGAN_PROF_<kernel>.  meta[0..3] = {T, d, E, flags} of the block call the launch
 * belongs to (flags: bits 0-7 k, 0x100 dense, 0x200 hard, 0x400 fp32 mode, 0x1000.. kernel-specific dtype bits, see
 * bench.py), or {rows / B, d / P, C, dtype bits} for the layers around the block; meta[4..7] = {N, K (or the A dimension of a
 * weight-gradient GEMM), host-known rows or -1, experts in the group} for GEMM records, 0 otherwise.
 * moegan_debug_profile_read returns the GEMM records only (ms, tag), moegan_debug_profile_read_ex all records (ms, tag,
 * meta[8]) in launch order; both synchronise the events and clear the list. */
#define MOEGAN_PROF_KERNEL_BASE 2000000
enum {
    MOEGAN_PROF_SAMPLE_KL = 1, MOEGAN_PROF_TEXT_FWD = 2, MOEGAN_PROF_ROUTER_FWD = 3, MOEGAN_PROF_PLAN = 4, MOEGAN_PROF_GATHER = 5,
    MOEGAN_PROF_TO_OPERAND = 6, MOEGAN_PROF_PREP_WEIGHTS = 7, MOEGAN_PROF_COMBINE_FWD = 8, MOEGAN_PROF_COMBINE_BWD = 9,
    MOEGAN_PROF_SLAB_SUM = 10, MOEGAN_PROF_UNCOMBINE = 11, MOEGAN_PROF_ROUTER_BWD_LOGITS = 12, MOEGAN_PROF_ROUTER_BWD_SMALL = 13,
    MOEGAN_PROF_ROUTER_BWD_TEXT = 14, MOEGAN_PROF_REPARAM_KL_BWD = 15, MOEGAN_PROF_SUM_PARTIALS = 16, MOEGAN_PROF_COLSUM_GROUPED = 17,
    MOEGAN_PROF_BALANCE = 18,
    MOEGAN_PROF_LAYERNORM_FWD = 20, MOEGAN_PROF_LAYERNORM_BWD = 21, MOEGAN_PROF_CHANNEL_SCALE_FWD = 22, MOEGAN_PROF_CHANNEL_SCALE_BWD = 23,
    MOEGAN_PROF_OFFSET_SAMPLE_FWD = 24, MOEGAN_PROF_OFFSET_SAMPLE_BWD = 25, MOEGAN_PROF_COLSUM = 26,
    MOEGAN_PROF_UPSAMPLE2X_FWD = 27, MOEGAN_PROF_UPSAMPLE2X_BWD = 28, MOEGAN_PROF_DEMOD_COEF_FWD = 29, MOEGAN_PROF_DEMOD_COEF_BWD = 30
};
int moegan_debug_profile_enable(int32_t level);
int moegan_debug_profile_read(float* ms_out, int32_t* tag_out, int32_t max_n);
int moegan_debug_profile_read_ex(float* ms_out, int32_t* tag_out, int32_t* meta_out, int32_t max_n);
/* debug / bench: 0 = keep every kernel of a block call on the caller's stream (per-kernel timing without concurrency), 1 = allow the
 * library's side streams, -1 = default (on, unless MOEGAN_DISABLE_SIDE_STREAM=1) */
int moegan_debug_set_side_streams(int32_t mode);
/* debug / bench / tests: 0 = never use the fused FFN kernel (csrc/ffn_fused.cu) in forward-only block calls, 1 / -1 = default (used in
 * bf16 mode at d = 128 / 256 unless MOEGAN_DISABLE_FUSED_FFN=1).  The forward workspace layout depends on it: query
 * moegan_block_fwd_workspace again after changing it. */
int moegan_debug_set_fused_ffn(int32_t mode);
/* debug: override the MN-major shared-memory descriptor parameters of the GEMM (0 = built-in defaults) */
int moegan_debug_set_mn_desc(int32_t lbo_bytes, int32_t sbo_bytes, int32_t kadv_bytes);
/* fp32 [rows, cols] -> bf16 operand [rows, S*cols] */
int moegan_to_operand(const float* src, int64_t rows, int64_t cols, int32_t precision, void* dst, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MOEGAN_SM100_H */
