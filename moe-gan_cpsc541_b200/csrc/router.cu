// Bayesian router kernels (forward + backward) and the balance-loss scalar kernels.
//
// Reference semantics: BayesianRouter (moegan/t2i_moe_gan.py:265-423).  All router math is fp32 on the
// CUDA cores on purpose: expert assignment has to be bit-exact against the reference (north_star), so
// the 128-wide projections are NOT sent through the bf16 tensor pipe in the forward pass.
#include "kernels.cuh"

#include <cstdlib>

namespace moegan {

namespace {

constexpr int HID = MOEGAN_ROUTER_HID;

// =====================================================================================================
// K1: sample the three weight matrices from (mu, rho, eps) + KL reduction  (t2i_moe_gan.py:302-333, 405-423)
// =====================================================================================================
__device__ __forceinline__ float softplus_ref(float x) { return log1pf(expf(x)); }   // log1p(exp(x)) as written at :316

__global__ void __launch_bounds__(256)
sample_kl_kernel(const float* __restrict__ mu_f, const float* __restrict__ rho_f, const float* __restrict__ eps_f,
                 const float* __restrict__ mu_t, const float* __restrict__ rho_t, const float* __restrict__ eps_t,
                 const float* __restrict__ mu_c, const float* __restrict__ rho_c, const float* __restrict__ eps_c,
                 int n_f, int n_t, int n_c, int sampling, int split,
                 float* __restrict__ Wf, float* __restrict__ Wt, float* __restrict__ Wc,
                 __nv_bfloat16* __restrict__ Wf_op, __nv_bfloat16* __restrict__ Wt_op, float* __restrict__ kl_part,
                 float* __restrict__ keep_f, float* __restrict__ keep_t, float* __restrict__ keep_c) {
    const int n = n_f + n_t + n_c;
    float kl = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *mu, *rho, *eps; float* W; __nv_bfloat16* op; float* keep; int j;
        if (i < n_f)            { j = i;             mu = mu_f; rho = rho_f; eps = eps_f; W = Wf; op = Wf_op; keep = keep_f; }
        else if (i < n_f + n_t) { j = i - n_f;       mu = mu_t; rho = rho_t; eps = eps_t; W = Wt; op = Wt_op; keep = keep_t; }
        else                    { j = i - n_f - n_t; mu = mu_c; rho = rho_c; eps = eps_c; W = Wc; op = nullptr; keep = keep_c; }
        const float m = mu[j], r = rho[j];
        float w;
        if (sampling) {
            const float sigma = clampf(softplus_ref(clampf(r, -8.0f, 4.0f)), 1e-6f, 10.0f);
            const float ev = eps[j];
            if (keep != nullptr) keep[j] = ev;     // the backward needs the eps this forward consumed (the buffers get redrawn)
            w = clampf(m, -10.0f, 10.0f) + sigma * clampf(ev, -2.0f, 2.0f);
        } else {
            w = m;   // eval: raw mu, no clamp (:357-361)
        }
        W[j] = w;
        if (op != nullptr) {   // bf16 operand copy [rows, S*128] for the backward GEMMs
            const int row = j / HID, col = j - row * HID;
            float hi, lo;
            split_bf16(w, hi, lo);
            if (split) {
                op[(size_t)row * (2 * HID) + col] = __float2bfloat16_rn(hi);
                op[(size_t)row * (2 * HID) + HID + col] = __float2bfloat16_rn(lo);
            } else {
                op[(size_t)row * HID + col] = __float2bfloat16_rn(hi);
            }
        }
        // KL term uses the UNclamped rho (:408-415)
        const float sp = softplus_ref(r);
        const float log_var = 2.0f * logf(sp);
        kl += 0.5f * (expf(log_var) + m * m - 1.0f - log_var);
    }
    __shared__ float red[8];
    kl = warp_sum(kl);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = kl;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int w = 0; w < 8; ++w) s += red[w];
        kl_part[blockIdx.x] = s;
    }
}

__global__ void kl_finalize_kernel(const float* __restrict__ kl_part, int nparts, float* __restrict__ kl_out) {
    // single warp, fixed order -> deterministic
    float s = 0.0f;
    for (int i = threadIdx.x; i < nparts; i += 32) s += kl_part[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) {
        kl_out[1] = s;
        float v = s;
        if (v != v) v = 0.0f;                          // nan_to_num(nan=0, posinf=200, neginf=0)  (:419)
        else if (isinf(v)) v = v > 0.0f ? 200.0f : 0.0f;
        kl_out[0] = clampf(v, 0.0f, 120.0f);           // (:420)
    }
}

// =====================================================================================================
// K2a: text half of the router, once per sample  (t2i_moe_gan.py:365, text rows of :371)
// =====================================================================================================
constexpr int TX_SPLIT = 4;
__global__ void __launch_bounds__(HID * TX_SPLIT)
text_fwd_kernel(const float* __restrict__ w, const float* __restrict__ Wt, const float* __restrict__ Wc,
                int Dt, int E, float* __restrict__ tp, float* __restrict__ tl) {
    extern __shared__ float sm[];
    float* ws = sm;                    // [Dt]
    float* part = sm + Dt;             // [TX_SPLIT][128]
    float* tps = part + TX_SPLIT * HID;   // [128]
    const int b = blockIdx.x, tid = threadIdx.x, j = tid & (HID - 1), q = tid >> 7;
    for (int k = tid; k < Dt; k += HID * TX_SPLIT) ws[k] = nan_to_num_router(w[(size_t)b * Dt + k]);
    __syncthreads();
    const int kper = (Dt + TX_SPLIT - 1) / TX_SPLIT;
    const int k0 = q * kper, k1 = min(k0 + kper, Dt);
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    int k = k0;
    for (; k + 15 < k1; k += 16) {     // 16 independent loads in flight: the kernel is a chain of L2 round trips otherwise
        float wv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) wv[u] = __ldg(Wt + (size_t)(k + u) * HID + j);
#pragma unroll
        for (int u = 0; u < 16; u += 4) {
            a0 = fmaf(ws[k + u], wv[u], a0);
            a1 = fmaf(ws[k + u + 1], wv[u + 1], a1);
            a2 = fmaf(ws[k + u + 2], wv[u + 2], a2);
            a3 = fmaf(ws[k + u + 3], wv[u + 3], a3);
        }
    }
    for (; k < k1; ++k) a0 = fmaf(ws[k], Wt[(size_t)k * HID + j], a0);
    part[q * HID + j] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (q == 0) {
        const float acc = (part[j] + part[HID + j]) + (part[2 * HID + j] + part[3 * HID + j]);
        tp[(size_t)b * HID + j] = acc;
        tps[j] = acc;
    }
    __syncthreads();
    if (tid < E) {
        float sacc = 0.0f;
        for (int i = 0; i < HID; ++i) sacc = fmaf(tps[i], Wc[(size_t)(HID + i) * E + tid], sacc);
        tl[(size_t)b * E + tid] = sacc;
    }
}

// =====================================================================================================
// K2: fused router forward.  CTA = 32 tokens (one permutation chunk), 8 warps: warp (g, q) accumulates
// tokens 8q..8q+7 over the k rows of parity-half g of every smem chunk (split-K inside the CTA, so every
// SMSP has two resident warps even though there is only one CTA per SM at T = 4096); lane owns 4 of the 128
// projection columns.  Wf streams through shared memory in 32-row chunks (cp.async, double buffered).
// Epilogue is warp-per-token: logits, temperature, clamps, softmax, clamp+renorm, top-k, histogram.
// =====================================================================================================
constexpr int RT_TOK = MOEGAN_CHUNK;   // 32 tokens per CTA
constexpr int RT_KC = 32;              // k rows per smem chunk
constexpr int RT_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool valid) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
    const int sz = valid ? 16 : 0;   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Epilogue of the router forward: each warp finishes 4 tokens - logits from the 128 projection columns held 4 per lane, temperature,
// clamps, softmax, clamp + renormalise, top-k by warp argmax (ties -> lowest index), histogram into cnt_s.  P = split-K exchange buffer
// [2 (g)][4 (q)][8 (a)][128].  Returns this lane's share of the probability column sums.
// The 4 tokens are processed SIDE BY SIDE: every stage (logit butterfly, max, two sums, the k argmax rounds) is a chain of dependent warp
// shuffles, ~40 deep per token; one token after the other (round 1: `#pragma unroll 1` over the tokens) that chain - not the
// contraction - was most of a CTA's life (~2 000 cycles per token, 4 tokens per warp, against ~2 500 cycles for the whole [32, d] x
// [d, 128] contraction at d = 128).  Interleaved, the four chains hide each other's shuffle latency.  Per-token arithmetic and the
// order of every reduction are unchanged, so logits, probabilities and routing are bit-identical.
template <int U>
__device__ __forceinline__ void warp_sum_n(float (&v)[U]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] += __shfl_xor_sync(0xffffffffu, v[u], o);
}
template <int U>
__device__ __forceinline__ void warp_max_n(float (&v)[U]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = fmaxf(v[u], __shfl_xor_sync(0xffffffffu, v[u], o));
}
__device__ __forceinline__ float router_epilogue_tokens(const float* __restrict__ P, const float* __restrict__ WcT, const float* __restrict__ tl,
                                                       const float* __restrict__ temperature, float af, int g, int q, int lane, int t0, int T,
                                                       int E, int k, int tps, int dense, int hard, float* __restrict__ fp, float* __restrict__ raw,
                                                       float* __restrict__ probs, int* __restrict__ topk_idx, float* __restrict__ gates,
                                                       int* __restrict__ cnt_s) {
    constexpr int U = 4;
    const float tau = clampf(__ldg(temperature) * af, 0.5f, 5.0f);   // (:375)
    const bool act = lane < E;
    int tt[U]; bool ok[U];
    float f[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int a = g * 4 + u;
        const int t = t0 + q * 8 + a;
        ok[u] = t < T;                       // warp-uniform; tokens past the end are computed on zero rows and never stored
        tt[u] = ok[u] ? t : T - 1;
        const float4 p0 = *reinterpret_cast<const float4*>(P + ((0 * 4 + q) * 8 + a) * HID + lane * 4);
        const float4 p1 = *reinterpret_cast<const float4*>(P + ((1 * 4 + q) * 8 + a) * HID + lane * 4);
        f[u][0] = p0.x + p1.x; f[u][1] = p0.y + p1.y; f[u][2] = p0.z + p1.z; f[u][3] = p0.w + p1.w;
        if (fp != nullptr && ok[u]) *reinterpret_cast<float4*>(fp + (size_t)t * HID + lane * 4) = make_float4(f[u][0], f[u][1], f[u][2], f[u][3]);
    }
    float rw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) rw[u] = 0.0f;
    if (E <= 8) {
        // logits of up to 8 experts with ONE transposing butterfly (10 shuffles) instead of E full warp reductions (5 E): at every
        // stage a lane keeps the half of its values that belongs to its side of the xor pair and hands over the other half.  The
        // pairings are those of warp_sum (xor 16, 8, 4, 2, 1) and a + b == b + a, so every logit is bit-identical to the serial form.
        float v[U][8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float4 wc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < E) wc = *reinterpret_cast<const float4*>(WcT + e * HID + lane * 4);
#pragma unroll
            for (int u = 0; u < U; ++u) v[u][e] = e < E ? fmaf(f[u][3], wc.w, fmaf(f[u][2], wc.z, fmaf(f[u][1], wc.y, f[u][0] * wc.x))) : 0.0f;
        }
#pragma unroll
        for (int st = 0; st < 3; ++st) {
            const int off = 16 >> st, half = 4 >> st;
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float keep = up ? v[u][i + half] : v[u][i];
                    const float send = up ? v[u][i] : v[u][i + half];
                    v[u][i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
        }
        // lanes with (bit4, bit3, bit2) = (e>>2, e>>1, e) & 1 hold expert e after the last two steps; lane e fetches its logit
        const int src = (((lane >> 2) & 1) << 4) | (((lane >> 1) & 1) << 3) | ((lane & 1) << 2);
#pragma unroll
        for (int u = 0; u < U; ++u) v[u][0] += __shfl_xor_sync(0xffffffffu, v[u][0], 2);
#pragma unroll
        for (int u = 0; u < U; ++u) v[u][0] += __shfl_xor_sync(0xffffffffu, v[u][0], 1);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float r = __shfl_sync(0xffffffffu, v[u][0], src);
            rw[u] = act ? r : 0.0f;
        }
    } else {
        for (int e = 0; e < E; ++e) {
            const float4 wc = *reinterpret_cast<const float4*>(WcT + e * HID + lane * 4);
            float part[U];
#pragma unroll
            for (int u = 0; u < U; ++u) part[u] = fmaf(f[u][3], wc.w, fmaf(f[u][2], wc.z, fmaf(f[u][1], wc.y, f[u][0] * wc.x)));
            warp_sum_n<U>(part);
#pragma unroll
            for (int u = 0; u < U; ++u) if (lane == e) rw[u] = part[u];
        }
    }
    float l[U], m[U], ex[U], sm[U], cl[U], cs[U], p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (act) rw[u] += __ldg(tl + (size_t)(tt[u] / tps) * E + lane);
        const float z = rw[u] / tau;                               // (:378)
        l[u] = clampf(z, -20.0f, 20.0f);                           // (:381)
        m[u] = act ? l[u] : -INFINITY;
    }
    warp_max_n<U>(m);
#pragma unroll
    for (int u = 0; u < U; ++u) { ex[u] = act ? expf(l[u] - m[u]) : 0.0f; sm[u] = ex[u]; }
    warp_sum_n<U>(sm);
#pragma unroll
    for (int u = 0; u < U; ++u) { const float sft = ex[u] / sm[u]; cl[u] = act ? clampf(sft, 1e-6f, 1.0f) : 0.0f; cs[u] = cl[u]; }   // (:384, :387)
    warp_sum_n<U>(cs);
    float my_colsum = 0.0f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        p[u] = cl[u] / cs[u];                                      // (:389)
        my_colsum += (act && ok[u]) ? p[u] : 0.0f;
        if (act && ok[u]) raw[(size_t)tt[u] * E + lane] = rw[u];
    }
    if (dense) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (act && ok[u]) {
                probs[(size_t)tt[u] * E + lane] = p[u];
                gates[(size_t)tt[u] * E + lane] = p[u];
                topk_idx[(size_t)tt[u] * E + lane] = lane;
            }
        return my_colsum;
    }
    // top-k: k rounds of warp argmax (ties -> lowest index); lane j keeps the j-th selection of every token
    bool taken[U]; float gsum[U], my_p[U]; int my_i[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { taken[u] = false; gsum[u] = 0.0f; my_p[u] = 0.0f; my_i[u] = 0; }
    for (int j = 0; j < k; ++j) {
        float bv[U]; int bi[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { bv[u] = (act && !taken[u]) ? p[u] : -1.0f; bi[u] = lane; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv[u], o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi[u], o);
                if (ov > bv[u] || (ov == bv[u] && oi < bi[u])) { bv[u] = ov; bi[u] = oi; }
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (lane == bi[u]) taken[u] = true;
            if (lane == j) { my_i[u] = bi[u]; my_p[u] = bv[u]; }
            gsum[u] += bv[u];
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int first = __shfl_sync(0xffffffffu, my_i[u], 0);
        if (ok[u]) {
            if (lane < k) {
                topk_idx[(size_t)tt[u] * k + lane] = my_i[u];
                gates[(size_t)tt[u] * k + lane] = hard ? 1.0f : my_p[u] / gsum[u];
                atomicAdd(&cnt_s[my_i[u]], 1);
            }
            if (act) probs[(size_t)tt[u] * E + lane] = hard ? (lane == first ? 1.0f : 0.0f) : p[u];   // (:392-400)
        }
    }
    return my_colsum;
}

__global__ void __launch_bounds__(RT_THREADS)
router_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Wf, const float* __restrict__ Wc,
                  const float* __restrict__ tl, const float* __restrict__ temperature, float af,
                  int T, int d, int E, int k, int tps, int dense, int hard,
                  float* __restrict__ fp, float* __restrict__ raw, float* __restrict__ probs,
                  int* __restrict__ topk_idx, float* __restrict__ gates,
                  int* __restrict__ chunk_counts, float* __restrict__ chunk_colsum) {
    extern __shared__ __align__(16) uint8_t rsm[];
    float* Ws = reinterpret_cast<float*>(rsm);                    // [2][RT_KC][128]  (reused for the split-K exchange)
    float* xs = Ws + 2 * RT_KC * HID;                             // [2][RT_TOK][RT_KC]
    float* WcT = xs + 2 * RT_TOK * RT_KC;                         // [E][128]
    float* colsum_s = WcT + E * HID;                              // [8][32]
    int* cnt_s = reinterpret_cast<int*>(colsum_s + 8 * 32);       // [32]

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = wid >> 2, q = wid & 3;
    const int chunk = blockIdx.x;
    const int t0 = chunk * RT_TOK;

    for (int i = tid; i < HID * E; i += RT_THREADS) {   // WcT[e][j] = Wc[j][e], top 128 rows only
        const int j = i / E, e = i - j * E;
        WcT[e * HID + j] = Wc[i];
    }
    if (tid < 32) cnt_s[tid] = 0;

    auto load_chunk = [&](int c, int buf) {
        const int k0 = c * RT_KC;
        float* wdst = Ws + buf * RT_KC * HID;
#pragma unroll
        for (int i = 0; i < (RT_KC * HID / 4) / RT_THREADS; ++i) {     // 4 x 16 B per thread
            const int idx = tid + i * RT_THREADS;
            cp_async16(wdst + idx * 4, Wf + (size_t)k0 * HID + idx * 4, true);
        }
        float* xdst = xs + buf * RT_TOK * RT_KC;                       // 1 x 16 B per thread
        const int tok = tid >> 3, q4 = tid & 7;
        const bool ok = (t0 + tok) < T;
        const float* src = x + (size_t)(ok ? (t0 + tok) : 0) * d + k0 + q4 * 4;
        cp_async16(xdst + tok * RT_KC + q4 * 4, src, ok);
    };

    float acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a) { acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.0f; }

    const int nchunks = d / RT_KC;
    load_chunk(0, 0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunks) { load_chunk(c + 1, buf ^ 1); cp_async_commit(); cp_async_wait<1>(); }
        else                 { cp_async_wait<0>(); }
        {   // nan_to_num on the x elements this thread staged (identity on finite input, :338-340)
            float4* p4 = reinterpret_cast<float4*>(xs + buf * RT_TOK * RT_KC + (tid >> 3) * RT_KC + (tid & 7) * 4);
            float4 v = *p4;
            v.x = nan_to_num_router(v.x); v.y = nan_to_num_router(v.y);
            v.z = nan_to_num_router(v.z); v.w = nan_to_num_router(v.w);
            *p4 = v;
        }
        __syncthreads();
        const float* wsrc = Ws + buf * RT_KC * HID;
        const float* xsrc = xs + buf * RT_TOK * RT_KC + (q * 8) * RT_KC;
        // packed fp32 FMA (fma.rn.f32x2, sm_100): two of the lane's four columns per instruction - the same two IEEE fused
        // multiply-adds, bit for bit, in half the issue slots (ncu: this loop was 45 % FFMA and issue / latency bound)
        float2 p01[8], p23[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) { p01[a] = make_float2(0.0f, 0.0f); p23[a] = make_float2(0.0f, 0.0f); }
#pragma unroll
        for (int kq4 = 0; kq4 < RT_KC / 8; ++kq4) {
            const int kq = g * (RT_KC / 8) + kq4;       // this warp group's half of the chunk
            float4 xv[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) xv[a] = *reinterpret_cast<const float4*>(xsrc + a * RT_KC + kq * 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 wv = *reinterpret_cast<const float4*>(wsrc + (kq * 4 + kk) * HID + lane * 4);
                const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const float xk = kk == 0 ? xv[a].x : kk == 1 ? xv[a].y : kk == 2 ? xv[a].z : xv[a].w;
                    const float2 xx = make_float2(xk, xk);
                    p01[a] = __ffma2_rn(xx, w01, p01[a]);
                    p23[a] = __ffma2_rn(xx, w23, p23[a]);
                }
            }
        }
        // blocked summation: per-chunk partials are added once -> error grows with sqrt(#chunks), not K
#pragma unroll
        for (int a = 0; a < 8; ++a) { acc[a][0] += p01[a].x; acc[a][1] += p01[a].y; acc[a][2] += p23[a].x; acc[a][3] += p23[a].y; }
        __syncthreads();
    }

    // ---- split-K exchange through shared memory: P[g][q][a][128] (32 KB, aliases the W staging buffers) ----
    float* P = Ws;
#pragma unroll
    for (int a = 0; a < 8; ++a)
        *reinterpret_cast<float4*>(P + ((g * 4 + q) * 8 + a) * HID + lane * 4) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    __syncthreads();

    const float my_colsum = router_epilogue_tokens(P, WcT, tl, temperature, af, g, q, lane, t0, T, E, k, tps, dense, hard, fp, raw, probs,
                                                   topk_idx, gates, cnt_s);
    colsum_s[wid * 32 + lane] = my_colsum;
    __syncthreads();
    if (tid < E) {
        float cs = 0.0f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) cs += colsum_s[w8 * 32 + tid];   // fixed order -> deterministic
        chunk_colsum[(size_t)chunk * E + tid] = cs;
        chunk_counts[(size_t)chunk * E + tid] = cnt_s[tid];
    }
}

// =====================================================================================================
// B3: gates/probs -> raw logits gradient, dfp operand, dtau partials (warp per token)
// =====================================================================================================
constexpr int RB_THREADS = 256;
constexpr int RB_TOK = 16;   // tokens per CTA (8 warps x 2)

// FUSE_DWC (E <= 8): the CTA also accumulates its share of dWc[0:128] = fp^T draw (lane = 4 of the 128 rows, 8 experts in
// registers) while the draw values are in flight anyway, and writes ONE partial [128][E] per CTA - the separate pass over
// fp / draw (router_bwd_dwc_part_kernel, 30-90 us of the router's backward chain) disappears.  The grid is persistent
// (<= RB_MAX_BLOCKS CTAs walk the 16-token groups) so that the finishing sums stay short however large T is.
constexpr int RB_MAX_BLOCKS = 592;   // 4 CTAs per SM (64 registers): the per-token chain is latency-bound, occupancy is what hides it
constexpr int RB_FUSE_MAX_E = 8;
template <bool FUSE_DWC>
__global__ void __launch_bounds__(RB_THREADS)
router_bwd_logits_kernel(const float* __restrict__ raw, const float* __restrict__ Wc, const int* __restrict__ topk_idx,
                         const int* __restrict__ token_rows, const float* __restrict__ dgate_row,
                         const float* __restrict__ dprobs_ext, const float* __restrict__ dload,
                         const float* __restrict__ temperature, float af,
                         int T, int T_pad, int E, int k, int dense, int split, const float* __restrict__ fp,
                         float* __restrict__ draw, __nv_bfloat16* __restrict__ dfp_op, float* __restrict__ dtau_part,
                         float* __restrict__ dwc_part) {
    extern __shared__ __align__(16) float bsm[];
    float* WcT = bsm;                 // [E][128]
    float* red = bsm + E * HID;       // [8]
    float* wred = red + 8;            // FUSE_DWC: [8 warps][128 * E]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < HID * E; i += RB_THREADS) {
        const int j = i / E, e = i - j * E;
        WcT[e * HID + j] = Wc[i];
    }
    __syncthreads();
    const float tau = clampf(__ldg(temperature) * af, 0.5f, 5.0f);
    const bool act = lane < E;
    const float dl_add = (act && dload != nullptr) ? __ldg(dload + lane) : 0.0f;
    float dtau_acc = 0.0f;
    const int ldop = split ? 2 * HID : HID;
    float4 wacc[RB_FUSE_MAX_E];
#pragma unroll
    for (int e = 0; e < RB_FUSE_MAX_E; ++e) wacc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ngroups = (T + RB_TOK - 1) / RB_TOK;
    const int gpb = (ngroups + gridDim.x - 1) / gridDim.x;
    const int g_end = min(ngroups, (int)(blockIdx.x + 1) * gpb);
    for (int grp = blockIdx.x * gpb; grp < g_end; ++grp)
    for (int a = wid; a < RB_TOK; a += RB_THREADS / 32) {
        const int t = grp * RB_TOK + a;
        if (t >= T) break;
        // recompute the forward chain exactly as router_fwd_kernel does
        const float r = act ? raw[(size_t)t * E + lane] : 0.0f;
        const float z = r / tau;
        const float l = clampf(z, -20.0f, 20.0f);
        const float m = warp_max(act ? l : -INFINITY);
        const float ex = act ? expf(l - m) : 0.0f;
        const float s = ex / warp_sum(ex);
        const float cl = act ? clampf(s, 1e-6f, 1.0f) : 0.0f;
        const float S = warp_sum(cl);
        const float p = cl / S;
        float dp = dl_add;
        if (act && dprobs_ext != nullptr) dp += dprobs_ext[(size_t)t * E + lane];
        if (dense) {
            if (act) dp += dgate_row[(size_t)lane * T_pad + t];         // gate == prob (:470)
        } else {
            bool in = false; float dg = 0.0f;
            for (int j = 0; j < k; ++j) {
                const int e = __ldg(topk_idx + (size_t)t * k + j);
                if (e == lane) { in = true; dg = __ldg(dgate_row + __ldg(token_rows + (size_t)t * k + j)); }
            }
            const float G = warp_sum(in ? p : 0.0f);
            const float g = in ? p / G : 0.0f;
            const float dot = warp_sum(dg * g);
            if (in) dp += (dg - dot) / G;                                // g_i = p_i / sum_topk p
        }
        const float dc = (dp - warp_sum(act ? dp * p : 0.0f)) / S;      // p = c / S
        const float ds = (act && s >= 1e-6f && s <= 1.0f) ? dc : 0.0f;   // clamp mask
        const float dlg = s * (ds - warp_sum(ds * s));                   // softmax
        const float dz = (act && z >= -20.0f && z <= 20.0f) ? dlg : 0.0f;
        const float dr = dz / tau;
        dtau_acc += dz * z;                                              // dtau = -sum(dz*z)/tau
        if (act) draw[(size_t)t * E + lane] = dr;
        // dfp[t, j] = sum_e draw[e] * Wc[j, e]
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (FUSE_DWC) {
            const float4 f4 = __ldg(reinterpret_cast<const float4*>(fp + (size_t)t * HID + lane * 4));
#pragma unroll
            for (int e = 0; e < RB_FUSE_MAX_E; ++e) {
                if (e < E) {
                    const float de = __shfl_sync(0xffffffffu, dr, e);
                    const float4 wc = *reinterpret_cast<const float4*>(WcT + e * HID + lane * 4);
                    o.x = fmaf(de, wc.x, o.x); o.y = fmaf(de, wc.y, o.y); o.z = fmaf(de, wc.z, o.z); o.w = fmaf(de, wc.w, o.w);
                    wacc[e].x = fmaf(f4.x, de, wacc[e].x); wacc[e].y = fmaf(f4.y, de, wacc[e].y);
                    wacc[e].z = fmaf(f4.z, de, wacc[e].z); wacc[e].w = fmaf(f4.w, de, wacc[e].w);
                }
            }
        } else {
            for (int e = 0; e < E; ++e) {
                const float de = __shfl_sync(0xffffffffu, dr, e);
                const float4 wc = *reinterpret_cast<const float4*>(WcT + e * HID + lane * 4);
                o.x = fmaf(de, wc.x, o.x); o.y = fmaf(de, wc.y, o.y); o.z = fmaf(de, wc.z, o.z); o.w = fmaf(de, wc.w, o.w);
            }
        }
        __nv_bfloat16* dst = dfp_op + (size_t)t * ldop + lane * 4;
        float h0, h1, h2, h3, l0, l1, l2, l3;
        split_bf16(o.x, h0, l0); split_bf16(o.y, h1, l1); split_bf16(o.z, h2, l2); split_bf16(o.w, h3, l3);
        *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
        if (split) *reinterpret_cast<uint2*>(dst + HID) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
    }
    dtau_acc = warp_sum(dtau_acc);
    if (lane == 0) red[wid] = dtau_acc;
    __syncthreads();
    if (tid == 0) {
        float s = 0.0f;
        for (int w = 0; w < RB_THREADS / 32; ++w) s += red[w];
        dtau_part[blockIdx.x] = s;
    }
    if (FUSE_DWC) {
        // part[blk][j][e], j = lane * 4 + c: the 8 warps are summed in a fixed order
#pragma unroll
        for (int e = 0; e < RB_FUSE_MAX_E; ++e)
            if (e < E) {
                float* dst = wred + (size_t)wid * HID * E + (lane * 4) * E + e;
                dst[0] = wacc[e].x; dst[E] = wacc[e].y; dst[2 * E] = wacc[e].z; dst[3 * E] = wacc[e].w;
            }
        __syncthreads();
        for (int i = tid; i < HID * E; i += RB_THREADS) {
            float a = wred[i];
#pragma unroll
            for (int w = 1; w < RB_THREADS / 32; ++w) a += wred[(size_t)w * HID * E + i];
            dwc_part[(size_t)blockIdx.x * HID * E + i] = a;
        }
    }
}

// dtl[b, e] = sum over the sample's tokens of draw[t, e]  (one CTA per sample, deterministic)
__global__ void __launch_bounds__(256)
router_bwd_dtl_kernel(const float* __restrict__ draw, int tps, int E, float* __restrict__ dtl) {
    __shared__ float red[8][32];
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float acc = 0.0f;
    if (lane < E)
        for (int i = wid; i < tps; i += 8) acc += draw[((size_t)b * tps + i) * E + lane];
    red[wid][lane] = acc;
    __syncthreads();
    if (wid == 0 && lane < E) {
        float s = 0.0f;
        for (int w = 0; w < 8; ++w) s += red[w][lane];
        dtl[(size_t)b * E + lane] = s;
    }
}

// dWc[0:128] partials: part[blk][j][e] = sum over the block's token range of fp[t, j] * draw[t, e].
// At most WC_MAX_PARTS blocks, so the finishing sum stays short however large T is.
constexpr int WC_TOK = 32;          // tokens per inner chunk (fp values held in registers)
constexpr int WC_MAX_PARTS = 128;
__global__ void __launch_bounds__(HID)
router_bwd_dwc_part_kernel(const float* __restrict__ fp, const float* __restrict__ draw, int T, int E, int tok_per_block,
                           float* __restrict__ part) {
    __shared__ float ds[WC_TOK * MOEGAN_MAX_EXPERTS];
    const int j = threadIdx.x;
    const int tb0 = blockIdx.x * tok_per_block;
    const int tb1 = min(tb0 + tok_per_block, T);
    float acc[MOEGAN_MAX_EXPERTS];
#pragma unroll
    for (int e = 0; e < MOEGAN_MAX_EXPERTS; ++e) acc[e] = 0.0f;
    for (int t0 = tb0; t0 < tb1; t0 += WC_TOK) {
        const int nt = min(WC_TOK, tb1 - t0);
        __syncthreads();
        for (int i = j; i < nt * E; i += HID) ds[i] = draw[(size_t)t0 * E + i];
        float f[WC_TOK];
#pragma unroll
        for (int tt = 0; tt < WC_TOK; ++tt) f[tt] = tt < nt ? fp[(size_t)(t0 + tt) * HID + j] : 0.0f;   // 32 loads in flight
        __syncthreads();
#pragma unroll
        for (int e = 0; e < MOEGAN_MAX_EXPERTS; ++e) {
            if (e < E) {
                float a = acc[e];
#pragma unroll
                for (int tt = 0; tt < WC_TOK; ++tt) a = fmaf(f[tt], tt < nt ? ds[tt * E + e] : 0.0f, a);
                acc[e] = a;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < MOEGAN_MAX_EXPERTS; ++e)
        if (e < E) part[((size_t)blockIdx.x * HID + j) * E + e] = acc[e];
}

// Finishes the small router reductions.  grid = 1 + B + E blocks of 256 threads:
//   block 0            dtemperature
//   blocks 1 .. B      dtp[b] (+ bf16 operand copy)
//   blocks B+1 .. B+E  dWc: 128 outputs each of the top half (sum of the token-block partials) and of the
//                      bottom half (sum_b tp[b, j] * dtl[b, e])
__global__ void __launch_bounds__(256)
router_bwd_small_kernel(const float* __restrict__ dwc_part, int nparts, const float* __restrict__ tp, const float* __restrict__ dtl,
                        const float* __restrict__ Wc, const float* __restrict__ dtau_part, int ntau,
                        const float* __restrict__ temperature, float af, int B, int E, int split,
                        float* __restrict__ dWc, float* __restrict__ dtp, __nv_bfloat16* __restrict__ dtp_op,
                        float* __restrict__ dtemperature) {
    __shared__ float red[256];
    const int tid = threadIdx.x;
    if (blockIdx.x == 0) {
        float s = 0.0f;
        for (int i = tid; i < ntau; i += 256) s += dtau_part[i];
        red[tid] = s;
        __syncthreads();
        if (tid < 32) {
            float v = 0.0f;
            for (int i = tid; i < 256; i += 32) v += red[i];
            v = warp_sum(v);
            if (tid == 0) {
                const float ta = __ldg(temperature) * af;
                const float tau = clampf(ta, 0.5f, 5.0f);
                const float dtau = -v / tau;
                dtemperature[0] = (ta >= 0.5f && ta <= 5.0f) ? dtau * af : 0.0f;   // clamp mask (:375)
            }
        }
    } else if ((int)blockIdx.x <= B) {
        const int b = blockIdx.x - 1;
        if (tid < HID) {
            float s = 0.0f;
            for (int e = 0; e < E; ++e) s = fmaf(dtl[(size_t)b * E + e], Wc[(size_t)(HID + tid) * E + e], s);
            dtp[(size_t)b * HID + tid] = s;
            float hi, lo;
            split_bf16(s, hi, lo);
            const int ld = split ? 2 * HID : HID;
            dtp_op[(size_t)b * ld + tid] = __float2bfloat16_rn(hi);
            if (split) dtp_op[(size_t)b * ld + HID + tid] = __float2bfloat16_rn(lo);
        }
    } else {
        // dWc: CTA = 32 consecutive entries of the flat [128][E] matrix, 8 threads per entry.  Both sums are long and thin (up to
        // 296 partials / B samples per entry): every thread walks its eighth with four independent chains so that the L2 round
        // trips overlap (one serial chain per entry made this kernel ~65 us at every shape), then a fixed-order sum in smem.
        const int oo = tid & 31, h = tid >> 5;
        const int o = (blockIdx.x - 1 - B) * 32 + oo;
        const size_t ps = (size_t)HID * E;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        for (int pz = h; pz < nparts; pz += 32) {
            a0 += dwc_part[(size_t)pz * ps + o];
            if (pz + 8 < nparts)  a1 += dwc_part[(size_t)(pz + 8) * ps + o];
            if (pz + 16 < nparts) a2 += dwc_part[(size_t)(pz + 16) * ps + o];
            if (pz + 24 < nparts) a3 += dwc_part[(size_t)(pz + 24) * ps + o];
        }
        red[tid] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        if (h == 0) {
            float v = red[oo];
#pragma unroll
            for (int w8 = 1; w8 < 8; ++w8) v += red[w8 * 32 + oo];
            dWc[o] = v;
        }
        __syncthreads();
        const int j = o / E, e = o - j * E;                         // bottom rows: dWc[128 + j][e] = sum_b tp[b, j] dtl[b, e]
        a0 = a1 = a2 = a3 = 0.0f;
        for (int b = h; b < B; b += 32) {
            a0 = fmaf(tp[(size_t)b * HID + j], dtl[(size_t)b * E + e], a0);
            if (b + 8 < B)  a1 = fmaf(tp[(size_t)(b + 8) * HID + j], dtl[(size_t)(b + 8) * E + e], a1);
            if (b + 16 < B) a2 = fmaf(tp[(size_t)(b + 16) * HID + j], dtl[(size_t)(b + 16) * E + e], a2);
            if (b + 24 < B) a3 = fmaf(tp[(size_t)(b + 24) * HID + j], dtl[(size_t)(b + 24) * E + e], a3);
        }
        red[tid] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        if (h == 0) {
            float v = red[oo];
#pragma unroll
            for (int w8 = 1; w8 < 8; ++w8) v += red[w8 * 32 + oo];
            dWc[HID * E + o] = v;
        }
    }
}

// Text side of B4 for small batches: dw[b, k] = sum_j dtp[b, j] Wt[k, j]  and  dWt[k, j] = sum_b nan_to_num(w)[b, k] dtp[b, j].
// grid = Dt blocks of 128 threads (thread j): block k owns row k of Wt.  B <= 64 only (registers); larger batches use
// the tcgen05 GEMMs.
constexpr int TXB_MAX_B = 64;
__global__ void __launch_bounds__(HID)
router_bwd_text_kernel(const float* __restrict__ w, const float* __restrict__ dtp, const float* __restrict__ Wt, int B, int Dt,
                       float* __restrict__ dw, float* __restrict__ dWt) {
    __shared__ float red[4][TXB_MAX_B];
    const int k = blockIdx.x, j = threadIdx.x, lane = j & 31, wid = j >> 5;
    const float wt = Wt[(size_t)k * HID + j];
    float acc = 0.0f;
    for (int b = 0; b < B; ++b) {
        const float g = dtp[(size_t)b * HID + j];
        acc = fmaf(nan_to_num_router(__ldg(w + (size_t)b * Dt + k)), g, acc);
        const float p = warp_sum(g * wt);
        if (lane == 0) red[wid][b] = p;
    }
    dWt[(size_t)k * HID + j] = acc;
    __syncthreads();
    if (j < B) dw[(size_t)j * Dt + k] = (red[0][j] + red[1][j]) + (red[2][j] + red[3][j]);
}

// B5 + B6: dW -> (dmu, drho) through the reparameterisation clamps, plus the KL gradient.
__global__ void __launch_bounds__(256)
reparam_kl_bwd_kernel(const float* __restrict__ dWf, const float* __restrict__ dWt, const float* __restrict__ dWc,
                      const float* __restrict__ mu_f, const float* __restrict__ rho_f, const float* __restrict__ eps_f,
                      const float* __restrict__ mu_t, const float* __restrict__ rho_t, const float* __restrict__ eps_t,
                      const float* __restrict__ mu_c, const float* __restrict__ rho_c, const float* __restrict__ eps_c,
                      int n_f, int n_t, int n_c, int f_splits, int sampled, const float* __restrict__ kl_out, const float* __restrict__ dkl,
                      float* __restrict__ dmu_f, float* __restrict__ drho_f, float* __restrict__ dmu_t, float* __restrict__ drho_t,
                      float* __restrict__ dmu_c, float* __restrict__ drho_c) {
    const int n = n_f + n_t + n_c;
    float gkl = 0.0f;
    if (dkl != nullptr && kl_out != nullptr) {
        const float rawkl = kl_out[1];
        // clamp(nan_to_num(raw), 0, 120) passes gradient only for finite raw in [0, 120]
        if (rawkl == rawkl && !isinf(rawkl) && rawkl >= 0.0f && rawkl <= 120.0f) gkl = dkl[0];
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *dW, *mu, *rho, *eps; float *dmu, *drho; int j;
        if (i < n_f)            { j = i;             dW = dWf; mu = mu_f; rho = rho_f; eps = eps_f; dmu = dmu_f; drho = drho_f; }
        else if (i < n_f + n_t) { j = i - n_f;       dW = dWt; mu = mu_t; rho = rho_t; eps = eps_t; dmu = dmu_t; drho = drho_t; }
        else                    { j = i - n_f - n_t; dW = dWc; mu = mu_c; rho = rho_c; eps = eps_c; dmu = dmu_c; drho = drho_c; }
        float g = dW[j];
        if (i < n_f) for (int sp = 1; sp < f_splits; ++sp) g += dWf[(size_t)sp * n_f + j];   // split-K partials of the dWf GEMM
        const float m = mu[j], r = rho[j];
        float gm, gr;
        if (sampled) {
            const float rc = clampf(r, -8.0f, 4.0f);
            const float sp = softplus_ref(rc);
            const float sig = 1.0f / (1.0f + expf(-rc));
            gm = (m >= -10.0f && m <= 10.0f) ? g : 0.0f;
            const bool pass = (r >= -8.0f && r <= 4.0f) && (sp >= 1e-6f && sp <= 10.0f);
            gr = pass ? g * clampf(eps[j], -2.0f, 2.0f) * sig : 0.0f;
        } else {
            gm = g; gr = 0.0f;
        }
        if (gkl != 0.0f) {
            const float sp = softplus_ref(r);
            const float sig = 1.0f / (1.0f + expf(-r));
            gm = fmaf(gkl, m, gm);
            gr = fmaf(gkl, (sp - 1.0f / sp) * sig, gr);
        }
        dmu[j] = gm; drho[j] = gr;
    }
}

// =====================================================================================================
// balance loss on the fused column sums (t2i_moe_gan.py:973-1000): single warp
// =====================================================================================================
__global__ void balance_loss_kernel(const float* __restrict__ load, int E, int T, float bw,
                                    const float* __restrict__ dloss, float* __restrict__ loss, float* __restrict__ dload) {
    const int lane = threadIdx.x;
    const bool act = lane < E;
    const float frac = act ? (load[lane] + 1e-6f) / (float)T : 0.0f;        // (:973,:976)
    const float mean = warp_sum(frac) / (float)E;                            // (:987)
    const float dev = act ? frac - mean : 0.0f;
    const float var = warp_sum(dev * dev) / (float)(E - 1);                  // unbiased (:988); E == 1 -> nan
    const float sd = sqrtf(var);
    const float cv = sd / (mean + 1e-6f);                                    // (:989)
    const float v = (float)E * cv;                                           // (:992)
    if (loss != nullptr && lane == 0) {
        float c = fminf(fmaxf(v, 0.0f), 10.0f);                              // (:995)
        if (v != v) c = 0.0f;                                                // nan_to_num (:998)
        loss[0] = bw * c;
    }
    if (dload != nullptr && act) {
        float g = 0.0f;
        if (v == v && v >= 0.0f && v <= 10.0f && sd > 0.0f) {
            const float dstd = dev / ((float)(E - 1) * sd);
            const float dcv = dstd / (mean + 1e-6f) - sd / ((mean + 1e-6f) * (mean + 1e-6f)) / (float)E;
            g = dloss[0] * bw * (float)E * dcv / (float)T;
        }
        dload[lane] = g;
    }
}

}  // namespace

// =====================================================================================================
// host launchers (called from api.cu)
// =====================================================================================================
int launch_sample_kl(const moegan_shape* s, const float* mu_f, const float* rho_f, const float* eps_f,
                     const float* mu_t, const float* rho_t, const float* eps_t,
                     const float* mu_c, const float* rho_c, const float* eps_c, int sampling,
                     float* Wf, float* Wt, float* Wc, void* Wf_op, void* Wt_op, float* kl_out,
                     float* keep_f, float* keep_t, float* keep_c, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_SAMPLE_KL, s);
    const int n_f = s->d * HID, n_t = s->Dt * HID, n_c = 2 * HID * s->E;
    const int n = n_f + n_t + n_c;
    int blocks = ceil_div(n, 256 * 4);
    if (blocks > MOEGAN_KL_PARTS) blocks = MOEGAN_KL_PARTS;
    sample_kl_kernel<<<blocks, 256, 0, st>>>(mu_f, rho_f, eps_f, mu_t, rho_t, eps_t, mu_c, rho_c, eps_c, n_f, n_t, n_c,
                                             sampling, s->precision == 1, Wf, Wt, Wc,
                                             reinterpret_cast<__nv_bfloat16*>(Wf_op), reinterpret_cast<__nv_bfloat16*>(Wt_op), kl_out + 2,
                                             keep_f, keep_t, keep_c);
    MOEGAN_CHECK_LAUNCH("sample_kl_kernel");
    kl_finalize_kernel<<<1, 32, 0, st>>>(kl_out + 2, blocks, kl_out);
    MOEGAN_CHECK_LAUNCH("kl_finalize_kernel");
    return MOEGAN_OK;
}

int launch_text_fwd(const moegan_shape* s, const float* w, const float* Wt, const float* Wc, float* tp, float* tl, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_TEXT_FWD, s);
    text_fwd_kernel<<<s->B, HID * TX_SPLIT, (s->Dt + (TX_SPLIT + 1) * HID) * sizeof(float), st>>>(w, Wt, Wc, s->Dt, s->E, tp, tl);
    MOEGAN_CHECK_LAUNCH("text_fwd_kernel");
    return MOEGAN_OK;
}

static size_t router_fwd_smem(int E) {
    return (size_t)(2 * RT_KC * HID + 2 * RT_TOK * RT_KC + E * HID + 8 * 32) * sizeof(float) + 32 * sizeof(int);
}

int launch_router_fwd(const moegan_shape* s, const float* x, const float* Wf, const float* Wc, const float* tl,
                      const float* temperature, float af, float* fp, float* raw, float* probs, int32_t* topk_idx, float* gates,
                      int32_t* chunk_counts, float* chunk_colsum, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_ROUTER_FWD, s, fp != nullptr ? 0x1000 : 0);
    static bool attr = false;
    if (!attr) {
        MOEGAN_CHECK_CUDA(cudaFuncSetAttribute(router_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)router_fwd_smem(MOEGAN_MAX_EXPERTS)));
        attr = true;
    }
    const int chunks = ceil_div(s->T, RT_TOK);
    router_fwd_kernel<<<chunks, RT_THREADS, router_fwd_smem(s->E), st>>>(x, Wf, Wc, tl, temperature, af, s->T, s->d, s->E, s->k,
                                                                        s->tokens_per_sample, s->dense, s->hard, fp, raw, probs,
                                                                        topk_idx, gates, chunk_counts, chunk_colsum);
    MOEGAN_CHECK_LAUNCH("router_fwd_kernel");
    return MOEGAN_OK;
}

int launch_router_bwd_logits(const moegan_shape* s, const float* raw, const float* Wc, const int32_t* topk_idx,
                             const int32_t* token_rows, const float* dgate_row, const float* dprobs_ext, const float* dload,
                             const float* temperature, float af, const float* fp, float* draw, void* dfp_op, float* dtau_part,
                             float* dwc_part, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_ROUTER_BWD_LOGITS, s);
    const int blocks = router_bwd_logits_blocks(s);
    const int T_pad = row_pad(s->T);
    if (router_bwd_dwc_fused(s)) {
        const size_t smem = ((size_t)s->E * HID + 8 + (size_t)(RB_THREADS / 32) * HID * s->E) * sizeof(float);
        router_bwd_logits_kernel<true><<<blocks, RB_THREADS, smem, st>>>(
            raw, Wc, topk_idx, token_rows, dgate_row, dprobs_ext, dload, temperature, af, s->T, T_pad, s->E, s->k, s->dense,
            s->precision == 1, fp, draw, reinterpret_cast<__nv_bfloat16*>(dfp_op), dtau_part, dwc_part);
    } else {
        router_bwd_logits_kernel<false><<<blocks, RB_THREADS, (s->E * HID + 8) * sizeof(float), st>>>(
            raw, Wc, topk_idx, token_rows, dgate_row, dprobs_ext, dload, temperature, af, s->T, T_pad, s->E, s->k, s->dense,
            s->precision == 1, fp, draw, reinterpret_cast<__nv_bfloat16*>(dfp_op), dtau_part, nullptr);
    }
    MOEGAN_CHECK_LAUNCH("router_bwd_logits_kernel");
    return MOEGAN_OK;
}
bool router_bwd_dwc_fused(const moegan_shape* s) { return s->E <= RB_FUSE_MAX_E; }
int router_bwd_logits_blocks(const moegan_shape* s) {
    const int groups = ceil_div(s->T, RB_TOK);
    return groups < RB_MAX_BLOCKS ? (groups > 0 ? groups : 1) : RB_MAX_BLOCKS;
}
static int dwc_tok_per_block(const moegan_shape* s) {
    const int chunks = ceil_div(s->T, WC_TOK);
    return ceil_div(chunks, WC_MAX_PARTS) * WC_TOK;
}
int router_bwd_dwc_parts(const moegan_shape* s) {
    return router_bwd_dwc_fused(s) ? router_bwd_logits_blocks(s) : ceil_div(s->T, dwc_tok_per_block(s));
}

int launch_router_bwd_small(const moegan_shape* s, const float* fp, const float* draw, const float* tp, const float* Wc,
                            const float* dtau_part, const float* temperature, float af,
                            float* dtl, float* dwc_part, float* dWc, float* dtp, void* dtp_op, float* dtemperature, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_ROUTER_BWD_SMALL, s);
    router_bwd_dtl_kernel<<<s->B, 256, 0, st>>>(draw, s->tokens_per_sample, s->E, dtl);
    MOEGAN_CHECK_LAUNCH("router_bwd_dtl_kernel");
    const int nparts = router_bwd_dwc_parts(s);
    if (!router_bwd_dwc_fused(s)) {          // E <= 8: the logits kernel has already written the partials
        router_bwd_dwc_part_kernel<<<nparts, HID, 0, st>>>(fp, draw, s->T, s->E, dwc_tok_per_block(s), dwc_part);
        MOEGAN_CHECK_LAUNCH("router_bwd_dwc_part_kernel");
    }
    router_bwd_small_kernel<<<1 + s->B + 4 * s->E, 256, 0, st>>>(dwc_part, nparts, tp, dtl, Wc, dtau_part, router_bwd_logits_blocks(s),
                                                     temperature, af, s->B, s->E, s->precision == 1, dWc, dtp,
                                                     reinterpret_cast<__nv_bfloat16*>(dtp_op), dtemperature);
    MOEGAN_CHECK_LAUNCH("router_bwd_small_kernel");
    return MOEGAN_OK;
}

bool router_bwd_text_simt_ok(const moegan_shape* s) { return s->B <= TXB_MAX_B; }
int launch_router_bwd_text(const moegan_shape* s, const float* w, const float* dtp, const float* Wt, float* dw, float* dWt, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_ROUTER_BWD_TEXT, s);
    router_bwd_text_kernel<<<s->Dt, HID, 0, st>>>(w, dtp, Wt, s->B, s->Dt, dw, dWt);
    MOEGAN_CHECK_LAUNCH("router_bwd_text_kernel");
    return MOEGAN_OK;
}

int launch_reparam_kl_bwd(const moegan_shape* s, const float* dWf, const float* dWt, const float* dWc,
                          const float* mu_f, const float* rho_f, const float* eps_f,
                          const float* mu_t, const float* rho_t, const float* eps_t,
                          const float* mu_c, const float* rho_c, const float* eps_c, int sampled, int f_splits,
                          const float* kl_out, const float* dkl,
                          float* dmu_f, float* drho_f, float* dmu_t, float* drho_t, float* dmu_c, float* drho_c, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_REPARAM_KL_BWD, s);
    const int n_f = s->d * HID, n_t = s->Dt * HID, n_c = 2 * HID * s->E;
    const int blocks = ceil_div(n_f + n_t + n_c, 256 * 2);
    reparam_kl_bwd_kernel<<<blocks, 256, 0, st>>>(dWf, dWt, dWc, mu_f, rho_f, eps_f, mu_t, rho_t, eps_t, mu_c, rho_c, eps_c,
                                                  n_f, n_t, n_c, f_splits, sampled, kl_out, dkl, dmu_f, drho_f, dmu_t, drho_t, dmu_c, drho_c);
    MOEGAN_CHECK_LAUNCH("reparam_kl_bwd_kernel");
    return MOEGAN_OK;
}

int launch_balance_loss(const float* load, int E, int T, float bw, const float* dloss, float* loss, float* dload, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_BALANCE, (long long)T, 0, E, 0);
    balance_loss_kernel<<<1, 32, 0, st>>>(load, E, T, bw, dloss, loss, dload);
    MOEGAN_CHECK_LAUNCH("balance_loss_kernel");
    return MOEGAN_OK;
}

}  // namespace moegan
