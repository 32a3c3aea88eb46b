// Host-side interface of the tcgen05 grouped GEMM (implementation in umma_gemm.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace moegan {

enum GemmKind : int {
    GEMM_ROWS_BK = 0,   // C[rows, N] = A[rows, K] * B_e[N, K]^T      A K-major, B K-major      (forward GEMMs)
    GEMM_ROWS_BMN = 1,  // C[rows, N] = A[rows, K] * B_e[K, N]        A K-major, B MN-major     (dgrad)
    GEMM_WGRAD = 2      // C_e[dimA, N] = A[rows_e, dimA]^T * B[rows_e, N]   both MN-major, K = rows of expert e
};

enum GemmEpilogue : int {
    EPI_F32 = 0,        // out0 fp32 = acc
    EPI_BF16 = 1,       // out0 bf16 = acc
    EPI_BIAS_BF16 = 2,  // out0 bf16 = acc + bias[e, n]
    EPI_BIAS_F32 = 3,   // out0 fp32 = acc + bias[e, n]
    EPI_GELU_BF16 = 4,  // pre = acc + bias;  out0 bf16 = gelu(pre);  out1 bf16 = gelu'(pre) (skipped when out1 == NULL)
    EPI_GELU_X3 = 5,    // pre = acc + bias;  out1 fp32 = pre;  out0 (hi | lo) = gelu(pre)
    EPI_MUL_BF16 = 6,   // out0 bf16 = acc * aux bf16        (aux = gelu'(pre) saved by EPI_GELU_BF16)
    EPI_DGELU_X3 = 7    // out0 (hi | lo) = acc * gelu'(aux fp32)
};

struct GemmOperand {      // a 2-D bf16 tensor in global memory
    const void* ptr;
    uint64_t inner;       // contiguous extent (elements)
    uint64_t outer;       // rows
    uint64_t pitch;       // row pitch (elements)
};

struct GemmParams {
    int N;                  // output columns (per expert for WGRAD)
    int K;                  // ROWS kinds: reduction extent per pass (elements, multiple of 64)
    int passes;             // 1 (bf16) or 3 (bf16 split: hi*hi, hi*lo, lo*hi)
    int a_lo_off, b_lo_off; // inner-coordinate offset of the lo half of A / B
    int a_mod;              // ROWS: A row block = block % a_mod (dense mode: all experts read the same tokens); 0 = off
    int b_rel;              // WGRAD: B rows are relative to the expert's first row (dense mode: shared tokens)
    int b_rows_per_expert;  // ROWS_BK: rows of B per expert (= N);  ROWS_BMN: K rows per expert
    int m_blocks_host;      // ROWS: >= 0 -> ungrouped, this many 128-row blocks, expert 0;  < 0 -> read plan
    int k_rows_host;        // WGRAD: >= 0 -> ungrouped, reduction over rows [0, k_rows_host);  < 0 -> read plan
    int k_splits;           // WGRAD: split the reduction over this many CTAs; split s writes its partial to
                            //        out0 + s * E * dimA * N (the consumer sums the partials).  0/1 = off
    int dimA;               // WGRAD: output rows per expert
    int E;                  // WGRAD: experts
    const int* plan;        // device plan array (see common.cuh)
    int epi;                // GemmEpilogue
    void* out0;
    void* out1;
    const void* aux;
    const float* bias;      // [E, N]
    float* colsum_part;     // EPI_MUL_BF16 / EPI_DGELU_X3: if set, [rows / 32, N] fp32 column sums of every 32-row slab of the
                            // (unrounded) output - the bias gradient falls out of the dgrad epilogue instead of a second pass
    long long ld_out0, ld_out1, ld_aux;  // leading dimensions (elements)
    int out_rows_valid;     // rows of the output tensor(s): rows >= this are clipped by the TMA store
    int out_split_off;      // column offset of the lo half in out0 for the *_X3 epilogues
    int mn_lbo, mn_sbo, mn_kadv;  // MN-major smem descriptor parameters (bytes); 0 -> defaults
};

// SWIZZLE_128B tensor map of a bf16 operand with boxes of 64 (inner) x box_outer elements (defined in umma_gemm.cu; CUtensorMap is
// declared by <cuda.h>)
int make_operand_tmap(CUtensorMap* m, const GemmOperand& o, uint32_t box_outer);

// ---- ffn_fused.cu: Linear -> GELU -> Linear of a 256-row tile pair in ONE kernel; the [rows, 4d] hidden activations never leave the SM ----
// mode 0 (grouped rows): y[m_cap, d] bf16 = gelu(a W1_e^T + b1_e) W2_e^T + b2_e, expert e of every 128-row block from the plan
// mode 1 (dense-soft):   out[T, d] = sum_e gates[t, e] * (gelu(x W1_e^T + b1_e) W2_e^T + b2_e) (+ residual), every expert on every token
struct FfnFusedArgs {
    int mode;
    int d, E, T;            // T: token rows (mode 1) / unused (mode 0)
    int rows;               // rows of the A operand tensor (mode 0: m_cap, mode 1: T)
    const void* a_op;       // bf16 [rows, d]
    const void* W1_op;      // bf16 [E * 4d, d]
    const void* W2_op;      // bf16 [E * d, 4d]
    const float* b1s;       // [E, 4d]
    const float* b2s;       // [E, d]
    const int* plan;        // mode 0
    const float* gates;     // mode 1: [T, E]
    const void* residual;   // mode 1, optional: [T, d]
    int residual_bf16;
    void* out;              // mode 0: y bf16 [rows, d]; mode 1: out [T, d] fp32 / bf16
    int out_bf16;
};
bool ffn_fused_supported(int d, int precision);
int launch_ffn_fused(const FfnFusedArgs& a, cudaStream_t stream);

// max_tiles: upper bound on the number of output tiles (sizes the persistent grid).
int launch_umma_gemm(int kind, int BN, const GemmOperand& A, const GemmOperand& B, GemmParams p, long long max_tiles,
                     cudaStream_t stream);

}  // namespace moegan
