// Internal launcher declarations shared by router.cu / dispatch.cu / api.cu.
#pragma once
#include "common.cuh"

namespace moegan {

// ---- router.cu ------------------------------------------------------------------------------------
int launch_sample_kl(const moegan_shape* s, const float* mu_f, const float* rho_f, const float* eps_f,
                     const float* mu_t, const float* rho_t, const float* eps_t,
                     const float* mu_c, const float* rho_c, const float* eps_c, int sampling,
                     float* Wf, float* Wt, float* Wc, void* Wf_op, void* Wt_op, float* kl_out,
                     float* keep_f, float* keep_t, float* keep_c, cudaStream_t st);
int launch_text_fwd(const moegan_shape* s, const float* w, const float* Wt, const float* Wc, float* tp, float* tl, cudaStream_t st);
int launch_router_fwd(const moegan_shape* s, const float* x, const float* Wf, const float* Wc, const float* tl,
                      const float* temperature, float af, float* fp, float* raw, float* probs, int32_t* topk_idx, float* gates,
                      int32_t* chunk_counts, float* chunk_colsum, cudaStream_t st);
int launch_router_bwd_logits(const moegan_shape* s, const float* raw, const float* Wc, const int32_t* topk_idx,
                             const int32_t* token_rows, const float* dgate_row, const float* dprobs_ext, const float* dload,
                             const float* temperature, float af, const float* fp, float* draw, void* dfp_op, float* dtau_part,
                             float* dwc_part, cudaStream_t st);
bool router_bwd_dwc_fused(const moegan_shape* s);
int router_bwd_logits_blocks(const moegan_shape* s);
int router_bwd_dwc_parts(const moegan_shape* s);
int launch_router_bwd_small(const moegan_shape* s, const float* fp, const float* draw, const float* tp, const float* Wc,
                            const float* dtau_part, const float* temperature, float af,
                            float* dtl, float* dwc_part, float* dWc, float* dtp, void* dtp_op, float* dtemperature, cudaStream_t st);
bool router_bwd_text_simt_ok(const moegan_shape* s);
int launch_router_bwd_text(const moegan_shape* s, const float* w, const float* dtp, const float* Wt, float* dw, float* dWt, cudaStream_t st);
int launch_reparam_kl_bwd(const moegan_shape* s, const float* dWf, const float* dWt, const float* dWc,
                          const float* mu_f, const float* rho_f, const float* eps_f,
                          const float* mu_t, const float* rho_t, const float* eps_t,
                          const float* mu_c, const float* rho_c, const float* eps_c, int sampled, int f_splits,
                          const float* kl_out, const float* dkl,
                          float* dmu_f, float* drho_f, float* dmu_t, float* drho_t, float* dmu_c, float* drho_c, cudaStream_t st);
int launch_balance_loss(const float* load, int E, int T, float bw, const float* dloss, float* loss, float* dload, cudaStream_t st);

// ---- dispatch.cu ----------------------------------------------------------------------------------
int launch_plan(const moegan_shape* s, const int32_t* chunk_counts, const float* chunk_colsum, int32_t* plan,
                int32_t* chunk_base, float* load, cudaStream_t st);
int launch_gather(const moegan_shape* s, const float* x, const void* x_bf16, const int32_t* topk_idx, const float* gates, const int32_t* plan,
                  const int32_t* chunk_base, void* x_op, void* x_copy, int32_t* row_token, float* row_gate, int32_t* token_rows,
                  cudaStream_t st);
int launch_to_operand(const float* src, int64_t rows, int64_t cols, int split, int sanitize, void* dst, cudaStream_t st);
int launch_prep_weights(const moegan_shape* s, const float* const* w1, const float* const* b1, const float* const* w2,
                        const float* const* b2, void* W1_op, float* b1s, void* W2_op, float* b2s, cudaStream_t st);
int launch_combine_fwd(const moegan_shape* s, const void* y, const float* gates, const int32_t* token_rows, const void* residual,
                       int residual_bf16, void* out, int out_bf16, cudaStream_t st);
int launch_combine_bwd(const moegan_shape* s, const int32_t* plan, const void* dout, int dout_bf16, const void* y, const int32_t* row_token,
                       const float* row_gate, const float* gates, void* dy_op, float* dgate_row, float* dy_colpart, cudaStream_t st);
int slab_sum_splits(const moegan_shape* s);
int launch_slab_sum(const moegan_shape* s, const int32_t* plan, const float* part, int cols, float* scratch, float* out, cudaStream_t st);
int launch_uncombine(const moegan_shape* s, const void* dx_rows, const int32_t* token_rows, const float* dx_router, float* dx,
                     cudaStream_t st);
// out[e, c] = sum over the real rows of expert e of (hi + lo)[row, c];  src is a bf16 operand [m_cap, S*cols];
// scratch: >= 32 * E * cols floats (row-split partials)
int launch_colsum_grouped(const moegan_shape* s, const int32_t* plan, const void* src_op, int cols, float* scratch, float* out,
                          cudaStream_t st);

// ---- glue.cu (layers around the block: SURVEY 8(f) N1 / N3) ----------------------------------------
int ln_bwd_blocks(long long rows);
int launch_layernorm_fwd(const void* x, int x_bf16, const float* gamma, const float* beta, long long rows, int d, float eps, void* y,
                         int y_bf16, void* y_op, float* mean, float* rstd, cudaStream_t st, const void* addend = nullptr, void* sum_out = nullptr);
int launch_layernorm_bwd(const void* g, int g_bf16, const void* x, int x_bf16, const float* gamma, const float* mean, const float* rstd,
                         long long rows, int d, void* dx, float* dgamma_dbeta, float* scratch, cudaStream_t st, const void* gadd = nullptr);
int channel_scale_psplit(int B, int P, int C);
int launch_channel_scale_fwd(const void* x, int bf16, const void* s, int s_bf16, int B, int P, int C, float slope, void* out, cudaStream_t st);
int launch_channel_scale_bwd(const void* g, const void* x, int bf16, const void* s, int s_bf16, int B, int P, int C, float slope, void* dx,
                             void* ds, float* scratch, cudaStream_t st);
int colsum_splits(long long rows, int cols);
int launch_colsum(const void* x, int bf16, long long rows, int cols, float* out, float* scratch, cudaStream_t st);
int launch_offset_sample_fwd(const void* x, int bf16, const float* off, int B, int H, int W, int C, void* out, cudaStream_t st);
int launch_offset_sample_bwd(const void* g, const void* x, int bf16, const float* off, int B, int H, int W, int C, void* dx, float* doff,
                             cudaStream_t st);

int launch_upsample2x_fwd(const void* x, int bf16, int B, int H, int W, int C, void* out, cudaStream_t st);
int launch_upsample2x_bwd(const void* g, int bf16, int B, int H, int W, int C, void* dx, cudaStream_t st);

int launch_demod_coef_fwd(const void* s, int s_bf16, const float* w, long long so, long long si, long long sh, long long sw, int B, int Cin,
                          int Cout, int KH, int KW, float eps, int compute_wsq, float* wsq, float* d, cudaStream_t st);
int launch_demod_coef_bwd(const float* dd, const float* d, const void* s, int s_bf16, const float* w, long long so, long long si, long long sh,
                          long long sw, const float* wsq, int B, int Cin, int Cout, int KH, int KW, void* ds, float* dw, cudaStream_t st);

int launch_sum_partials(const float* part, long long n, int nsplit, float* out, cudaStream_t st);

}  // namespace moegan
