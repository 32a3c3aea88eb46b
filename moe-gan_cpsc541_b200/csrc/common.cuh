// Shared host/device helpers for libmoegan_sm100.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/moegan_sm100.h"

namespace moegan {

// ---- plan array layout (int32) -------------------------------------------------------------------
// [0] number of 128-row blocks in use   [1] rows in use (= 128 * [0])   [2] T_pad (dense) or 0
// [PLAN_OFF .. PLAN_OFF+E]  padded row offset of each expert segment (E+1 entries)
// [PLAN_CNT .. PLAN_CNT+E)  real rows per expert
// [PLAN_BLK .. )            expert id of each 128-row block
constexpr int PLAN_OFF = 8;
constexpr int PLAN_CNT = 48;
constexpr int PLAN_BLK = 88;

// ---- error plumbing ------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int num_sms();
void count_launch();

#define MOEGAN_CHECK_CUDA(expr)                                   \
    do {                                                          \
        cudaError_t _e = (expr);                                  \
        if (_e != cudaSuccess) return ::moegan::cuda_fail(_e, #expr); \
    } while (0)

#define MOEGAN_CHECK_LAUNCH(name)                                     \
    do {                                                              \
        cudaError_t _e = cudaGetLastError();                          \
        if (_e != cudaSuccess) return ::moegan::cuda_fail(_e, name);  \
        ::moegan::count_launch();                                     \
    } while (0)

#define MOEGAN_REQUIRE(cond, status, ...)      \
    do {                                       \
        if (!(cond)) {                         \
            ::moegan::set_error(__VA_ARGS__);  \
            return (status);                   \
        }                                      \
    } while (0)

int validate_shape(const moegan_shape* s);

// ---- bench instrumentation (api.cu): CUDA events around launches, recorded on the launching stream ----------
// level 0: off; 1: the tcgen05 GEMMs only; 2: every kernel launcher of the library (MOEGAN_PROF_* ids, include/moegan_sm100.h).
// A record carries 8 ints of metadata (4: shape of the block call / of the layer, 4: GEMM dimensions) so that bench.py can attach
// algorithmic bytes / flops to it.
int profile_level();
void profile_note_shape(const moegan_shape* s);                       // shape of the block call the following launches belong to
bool profile_begin(cudaStream_t st, int tag, const int* meta_lo, const int* meta_hi, void** handle);   // 4 + 4 ints; lo == nullptr: the noted shape
void profile_end(cudaStream_t st, void* handle);
inline int prof_shape_flags(const moegan_shape* s) {
    return (s->k & 0xff) | (s->dense ? 0x100 : 0) | (s->hard ? 0x200 : 0) | (s->precision == 1 ? 0x400 : 0);
}
struct ProfScope {
    void* h = nullptr;
    cudaStream_t st;
    ProfScope(cudaStream_t s, int id, const moegan_shape* sh, int extra_flags = 0) : st(s) {
        if (profile_level() >= 2) {
            const int m[4] = {sh->T, sh->d, sh->E, prof_shape_flags(sh) | extra_flags};
            profile_begin(s, MOEGAN_PROF_KERNEL_BASE + id, m, nullptr, &h);
        }
    }
    ProfScope(cudaStream_t s, int id, long long m0, int m1, int m2, int m3) : st(s) {
        if (profile_level() >= 2) {
            const int m[4] = {(int)m0, m1, m2, m3};
            profile_begin(s, MOEGAN_PROF_KERNEL_BASE + id, m, nullptr, &h);
        }
    }
    ~ProfScope() { if (h != nullptr) profile_end(st, h); }
    ProfScope(const ProfScope&) = delete;
    ProfScope& operator=(const ProfScope&) = delete;
};


inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int row_pad(int rows) { return (rows + MOEGAN_ROW_ALIGN - 1) / MOEGAN_ROW_ALIGN * MOEGAN_ROW_ALIGN; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// torch.nan_to_num(v, nan=0, posinf=1, neginf=-1) - identity on finite input (t2i_moe_gan.py:340,344)
__device__ __forceinline__ float nan_to_num_router(float v) {
    if (v != v) return 0.0f;
    if (isinf(v)) return v > 0.0f ? 1.0f : -1.0f;
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
// v = hi + lo + O(2^-17 |v|), both representable in bf16
__device__ __forceinline__ void split_bf16(float v, float& hi, float& lo) {
    hi = __bfloat162float(__float2bfloat16_rn(v));
    lo = v - hi;
}
__device__ __forceinline__ float bf16_lo_to_float(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_to_float(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

// ---- GELU (exact erf form, nn.GELU(approximate='none'), t2i_moe_gan.py:258) ----------------------------
// Precise variants (fp32 / 3-pass mode): erff / expf.
// Fast variants (bf16 mode; the result is rounded to bf16 = 2^-9 relative anyway): ONE MUFU per erf instead of
// the two (rcp + ex2) of the textbook rational forms, because the SFU pipe (4 lanes/SMSP) is what bounds a
// GELU epilogue running under a tcgen05 main loop:
//     erf(x/sqrt2) ~= tanh(x * (b0 + b1 x^2 + b2 x^4 + b3 x^6)),  |x| clamped to 5.5
// coefficients = weighted least-squares (minimax-refined) fit of atanh(erf(z)); max |erf error| 1.3e-5 before
// the tanh.approx error (2^-11 relative).
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float erf_sqrt2_fast(float x) {      // erf(x / sqrt(2))
    const float xc = clampf(x, -5.5f, 5.5f);
    const float s = xc * xc;
    float p = fmaf(s, -1.3688055686176932e-05f, -0.00019445275122756945f);
    p = fmaf(s, p, 0.036546657541717965f);
    p = fmaf(s, p, 0.797820262700123f);
    return tanh_approx(xc * p);
}
template <bool PRECISE>
__device__ __forceinline__ float gelu_erf(float x) {
    const float e = PRECISE ? erff(x * 0.70710678118654752f) : erf_sqrt2_fast(x);
    const float h = 0.5f * x;
    return fmaf(h, e, h);
}
// gelu'(x) = Phi(x) + x * phi(x)
template <bool PRECISE>
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float e = PRECISE ? erff(x * 0.70710678118654752f) : erf_sqrt2_fast(x);
    const float pdf = (PRECISE ? expf(-0.5f * x * x) : ex2_approx(-0.72134752044448170f * x * x)) * 0.39894228040143268f;
    return fmaf(x, pdf, fmaf(0.5f, e, 0.5f));
}
// Packed-half2 versions for the bf16 GEMM epilogues: the erf argument polynomial, tanh and the Gaussian are evaluated
// for TWO elements per instruction in fp16 (rel. eps 4.9e-4, far below the bf16 quantisation of the stored result:
// max |gelu error| 4.7e-4, max |gelu' error| 8.3e-4 over [-8, 8]); the final 0.5*x*(1 + erf) stays in fp32 so large
// pre-activations cannot overflow.  Halves the instruction count of the epilogue that bounds the K = d GEMM.
__device__ __forceinline__ __half2 h2_tanh_approx(__half2 x) {
    uint32_t r, a = *reinterpret_cast<uint32_t*>(&x);
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a));
    return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ __half2 h2_ex2_approx(__half2 x) {
    uint32_t r, a = *reinterpret_cast<uint32_t*>(&x);
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a));
    return *reinterpret_cast<__half2*>(&r);
}
template <bool WITH_GRAD>
__device__ __forceinline__ void gelu_pair_h2(float x0, float x1, float& g0, float& g1, float& d0, float& d1) {
    const __half2 x = __floats2half2_rn(x0, x1);
    const __half2 xc = __hmin2(__hmax2(x, __float2half2_rn(-5.5f)), __float2half2_rn(5.5f));
    const __half2 s = __hmul2(xc, xc);
    __half2 p = __hfma2(s, __float2half2_rn(-1.3688055686176932e-05f), __float2half2_rn(-0.00019445275122756945f));
    p = __hfma2(s, p, __float2half2_rn(0.036546657541717965f));
    p = __hfma2(s, p, __float2half2_rn(0.797820262700123f));
    const __half2 t = h2_tanh_approx(__hmul2(xc, p));              // erf(x / sqrt 2), two elements per MUFU op
    const float2 tf = __half22float2(t);
    const float h0 = 0.5f * x0, h1 = 0.5f * x1;
    g0 = fmaf(h0, tf.x, h0);
    g1 = fmaf(h1, tf.y, h1);
    if (WITH_GRAD) {
        const __half2 pdf = __hmul2(h2_ex2_approx(__hmul2(s, __float2half2_rn(-0.72134752044448170f))), __float2half2_rn(0.39894228040143268f));
        const __half2 cdf = __hfma2(t, __float2half2_rn(0.5f), __float2half2_rn(0.5f));
        const float2 df = __half22float2(__hfma2(xc, pdf, cdf));   // Phi(x) + x phi(x); |x| > 5.5 -> phi == 0
        d0 = df.x; d1 = df.y;
    }
}
// both at once (shares the erf): used by the forward epilogue, which stores gelu'(pre) for the backward pass
__device__ __forceinline__ void gelu_and_grad_fast(float x, float& g, float& dg) {
    const float e = erf_sqrt2_fast(x);
    const float h = 0.5f * x;
    g = fmaf(h, e, h);
    const float pdf = ex2_approx(-0.72134752044448170f * x * x) * 0.39894228040143268f;
    dg = fmaf(x, pdf, fmaf(0.5f, e, 0.5f));
}
#endif  // __CUDACC__

}  // namespace moegan
