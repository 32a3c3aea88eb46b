// Bandwidth-bound kernels for the layers directly around the MoE block:
//   * LayerNorm forward / backward (AttentionBlock.norm1-3, t2i_moe_gan.py:513-515,561) - warp per token row, the
//     weight / bias gradients as deterministic two-stage column sums (ATen's gamma/beta kernel needs 0.5 ms for the
//     [65536, 128] case of the 16x16 block);
//   * channel_scale: out = leaky_relu(x * s[b, c]) on NHWC activations, one pass forward, one pass backward that also
//     reduces ds[b, c] - the modulate / demodulate / activation steps of ModulatedConv (:158-186) written as
//     "scale the activations" instead of materialising per-sample weights;
//   * offset_sample: the learned-offset bilinear resampling of ModulatedTransformationModule (:219-240) on NHWC
//     activations: grid = clamp(base + 0.05 * off, -1, 1), F.grid_sample(bilinear, zeros, align_corners=False).
// All of them are HBM-bound: every tensor is read or written exactly once per pass, with 8- or 16-byte accesses along the
// contiguous channel dimension.
#include <cstdlib>
#include "kernels.cuh"
#include "ptx.cuh"

namespace moegan {

namespace {

// ---- 4 consecutive channels as fp32, from / to fp32 or bf16 storage ---------------------------------------
template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ float4 load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
        return make_float4(bf16_lo_to_float(u.x), bf16_hi_to_float(u.x), bf16_lo_to_float(u.y), bf16_hi_to_float(u.y));
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
        *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
};
template <typename T> __device__ __forceinline__ float load1(const T* p);
template <> __device__ __forceinline__ float load1<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void store1(T* p, float v);
template <> __device__ __forceinline__ void store1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- 8 consecutive channels (16 B of bf16 / 32 B of fp32): what bounds these kernels on B200 is the number of outstanding load
// REQUESTS per SM, so the wider the per-lane access the closer to the HBM roof (ncu on the combine kernels, dispatch.cu) ----
template <typename T> struct Vec8;
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
};
template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
        f[0] = bf16_lo_to_float(u.x); f[1] = bf16_hi_to_float(u.x); f[2] = bf16_lo_to_float(u.y); f[3] = bf16_hi_to_float(u.y);
        f[4] = bf16_lo_to_float(u.z); f[5] = bf16_hi_to_float(u.z); f[6] = bf16_lo_to_float(u.w); f[7] = bf16_hi_to_float(u.w);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
        *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
};

// 8 channels as loaded (no arithmetic): a load whose result is converted right away inside a branch cannot be overlapped with the next
// load - the four tap loads of an item then run one after the other (seen in the SASS of the first version: @!P BRA; LDG.128;
// IMAD.U32 ... per tap).  Taps outside the image read pixel 0 (a valid address) and are zeroed in the conversion.
template <typename T> struct Raw8;
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) { a = __ldg(reinterpret_cast<const float4*>(p)); b = __ldg(reinterpret_cast<const float4*>(p) + 1); }
    __device__ __forceinline__ void get(bool ok, float (&f)[8]) const {
        f[0] = ok ? a.x : 0.f; f[1] = ok ? a.y : 0.f; f[2] = ok ? a.z : 0.f; f[3] = ok ? a.w : 0.f;
        f[4] = ok ? b.x : 0.f; f[5] = ok ? b.y : 0.f; f[6] = ok ? b.z : 0.f; f[7] = ok ? b.w : 0.f;
    }
};
template <> struct Raw8<__nv_bfloat16> {
    uint4 u;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void get(bool ok, float (&f)[8]) const {
        const uint4 v = ok ? u : make_uint4(0u, 0u, 0u, 0u);
        f[0] = bf16_lo_to_float(v.x); f[1] = bf16_hi_to_float(v.x); f[2] = bf16_lo_to_float(v.y); f[3] = bf16_hi_to_float(v.y);
        f[4] = bf16_lo_to_float(v.z); f[5] = bf16_hi_to_float(v.z); f[6] = bf16_lo_to_float(v.w); f[7] = bf16_hi_to_float(v.w);
    }
};

// =====================================================================================================
// LayerNorm over the last dimension (d = 128 * nseg <= 1024), eps inside the sqrt (nn.LayerNorm)
// =====================================================================================================
// NSEG = 128-column segments held per lane (compile time: the per-row state lives in registers); R = rows a warp works on at
// once - for small d a single row is 256-512 bytes, far too little in flight per warp to cover DRAM latency.
template <typename TIN, typename TOUT, int NSEG, int R>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const TIN* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, long long rows, int d,
                     float eps, TOUT* __restrict__ y, __nv_bfloat16* __restrict__ y_op, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    const long long r0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * R;
    const int lane = threadIdx.x & 31;
    if (r0 >= rows) return;
    const int nseg = d / 128;
    float4 v[R][NSEG];
    float sum[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        sum[i] = 0.f;
#pragma unroll
        for (int s = 0; s < NSEG; ++s) {
            v[i][s] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s < nseg && r0 + i < rows) v[i][s] = Vec4<TIN>::load(x + (r0 + i) * d + s * 128 + lane * 4);
            sum[i] += (v[i][s].x + v[i][s].y) + (v[i][s].z + v[i][s].w);
        }
    }
    float mean[R], rstd[R];
#pragma unroll
    for (int i = 0; i < R; ++i) mean[i] = warp_sum(sum[i]) / (float)d;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        float sq = 0.f;
#pragma unroll
        for (int s = 0; s < NSEG; ++s)
            if (s < nseg) {
                const float a = v[i][s].x - mean[i], b = v[i][s].y - mean[i], c = v[i][s].z - mean[i], e = v[i][s].w - mean[i];
                sq += (a * a + b * b) + (c * c + e * e);
            }
        sum[i] = sq;
    }
#pragma unroll
    for (int i = 0; i < R; ++i) rstd[i] = rsqrtf(warp_sum(sum[i]) / (float)d + eps);
#pragma unroll
    for (int s = 0; s < NSEG; ++s)
        if (s < nseg) {
            const int c = s * 128 + lane * 4;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), b = __ldg(reinterpret_cast<const float4*>(beta + c));
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (r0 + i < rows) {
                    float4 o;
                    o.x = (v[i][s].x - mean[i]) * rstd[i] * g.x + b.x; o.y = (v[i][s].y - mean[i]) * rstd[i] * g.y + b.y;
                    o.z = (v[i][s].z - mean[i]) * rstd[i] * g.z + b.z; o.w = (v[i][s].w - mean[i]) * rstd[i] * g.w + b.w;
                    Vec4<TOUT>::store(y + (r0 + i) * d + c, o);
                    // norm3 -> MoE block: the bf16 GEMM-operand copy of the same row leaves in the same pass, so the
                    // block needs no separate fp32 -> bf16 conversion kernel over [T, d]
                    if (y_op != nullptr) Vec4<__nv_bfloat16>::store(y_op + (r0 + i) * d + c, o);
                }
        }
    if (lane < R && r0 + lane < rows) {
        float m = mean[0], q = rstd[0];
#pragma unroll
        for (int i = 1; i < R; ++i) if (lane == i) { m = mean[i]; q = rstd[i]; }
        mean_out[r0 + lane] = m; rstd_out[r0 + lane] = q;
    }
}

// dx = rstd * (gamma*g - mean_d(gamma*g) - xhat * mean_d(gamma*g*xhat)); per-CTA partial column sums of g*xhat and g.
// Each warp walks groups of R rows (group index = warp_global, += total_warps); the CTA's 8 warps are combined through
// shared memory; dx is written in x's storage type.
template <typename TG, typename TIN, int NSEG, int R>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const TG* __restrict__ g, const TIN* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, long long rows, int d,
                     TIN* __restrict__ dx, float* __restrict__ part /* [gridDim.x][2][d] */) {
    __shared__ float4 red[8][NSEG * 32];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nseg = d / 128;
    float4 gam[NSEG], dg[NSEG], db[NSEG];
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
        dg[s] = make_float4(0.f, 0.f, 0.f, 0.f); db[s] = dg[s]; gam[s] = dg[s];
        if (s < nseg) gam[s] = __ldg(reinterpret_cast<const float4*>(gamma + s * 128 + lane * 4));
    }
    const float inv_d = 1.0f / (float)d;
    for (long long r0 = ((long long)blockIdx.x * 8 + wid) * R; r0 < rows; r0 += (long long)gridDim.x * 8 * R) {
        float4 xh[R][NSEG], gg[R][NSEG];
        float s1[R], s2[R], rs[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const bool ok = r0 + i < rows;
            const float mean = ok ? __ldg(mean_in + r0 + i) : 0.f;
            rs[i] = ok ? __ldg(rstd_in + r0 + i) : 0.f;
            s1[i] = 0.f; s2[i] = 0.f;
#pragma unroll
            for (int s = 0; s < NSEG; ++s) {
                xh[i][s] = make_float4(0.f, 0.f, 0.f, 0.f); gg[i][s] = xh[i][s];
                if (s < nseg && ok) {
                    const long long o = (r0 + i) * d + s * 128 + lane * 4;
                    const float4 xv = Vec4<TIN>::load(x + o);
                    const float4 gv = Vec4<TG>::load(g + o);
                    xh[i][s] = make_float4((xv.x - mean) * rs[i], (xv.y - mean) * rs[i], (xv.z - mean) * rs[i], (xv.w - mean) * rs[i]);
                    dg[s].x += gv.x * xh[i][s].x; dg[s].y += gv.y * xh[i][s].y; dg[s].z += gv.z * xh[i][s].z; dg[s].w += gv.w * xh[i][s].w;
                    db[s].x += gv.x; db[s].y += gv.y; db[s].z += gv.z; db[s].w += gv.w;
                    gg[i][s] = make_float4(gv.x * gam[s].x, gv.y * gam[s].y, gv.z * gam[s].z, gv.w * gam[s].w);
                    s1[i] += (gg[i][s].x + gg[i][s].y) + (gg[i][s].z + gg[i][s].w);
                    s2[i] += (gg[i][s].x * xh[i][s].x + gg[i][s].y * xh[i][s].y) + (gg[i][s].z * xh[i][s].z + gg[i][s].w * xh[i][s].w);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < R; ++i) { s1[i] = warp_sum(s1[i]) * inv_d; s2[i] = warp_sum(s2[i]) * inv_d; }
#pragma unroll
        for (int i = 0; i < R; ++i)
            if (r0 + i < rows) {
#pragma unroll
                for (int s = 0; s < NSEG; ++s)
                    if (s < nseg) {
                        float4 o;
                        o.x = rs[i] * (gg[i][s].x - s1[i] - xh[i][s].x * s2[i]); o.y = rs[i] * (gg[i][s].y - s1[i] - xh[i][s].y * s2[i]);
                        o.z = rs[i] * (gg[i][s].z - s1[i] - xh[i][s].z * s2[i]); o.w = rs[i] * (gg[i][s].w - s1[i] - xh[i][s].w * s2[i]);
                        Vec4<TIN>::store(dx + (r0 + i) * d + s * 128 + lane * 4, o);
                    }
            }
    }
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int s = 0; s < NSEG; ++s)
            if (s < nseg) red[wid][s * 32 + lane] = pass == 0 ? dg[s] : db[s];
        __syncthreads();
        for (int j = threadIdx.x; j < nseg * 32; j += 256) {
            float4 a = red[0][j];
#pragma unroll
            for (int w = 1; w < 8; ++w) { const float4 b = red[w][j]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
            reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * 2 + pass) * d)[j] = a;
        }
        __syncthreads();
    }
}
// out[i] = sum_b part[b * n + i]  (fixed order); out fp32 or bf16
template <typename TOUT>
__global__ void __launch_bounds__(256)
partial_sum_kernel(const float* __restrict__ part, int n, int nparts, TOUT* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float s0 = 0.f, s1 = 0.f;
    int b = 0;
    for (; b + 1 < nparts; b += 2) { s0 += part[(size_t)b * n + i]; s1 += part[(size_t)(b + 1) * n + i]; }
    if (b < nparts) s0 += part[(size_t)b * n + i];
    store1<TOUT>(out + i, s0 + s1);
}
// the same for MANY partials (hundreds of CTAs): CTA = 32 columns x 8 row lanes, combined through shared memory in a fixed order
__global__ void __launch_bounds__(256)
partial_sum_tall_kernel(const float* __restrict__ part, int n, int nparts, float* __restrict__ out) {
    __shared__ float red[8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + cx;
    float s0 = 0.f, s1 = 0.f;
    if (i < n) {
        int b = ry;
        for (; b + 8 < nparts; b += 16) { s0 += part[(size_t)b * n + i]; s1 += part[(size_t)(b + 8) * n + i]; }
        if (b < nparts) s0 += part[(size_t)b * n + i];
    }
    red[ry][cx] = s0 + s1;
    __syncthreads();
    if (ry == 0 && i < n) {
        float a = red[0][cx];
#pragma unroll
        for (int w = 1; w < 8; ++w) a += red[w][cx];
        out[i] = a;
    }
}

// =====================================================================================================
// column sums of a tall matrix: out[c] = sum_r x[r, c]  (bias gradient of a Linear applied to [T, d] tokens; ATen's generic
// reduction needs ~80 us for a bf16 [65 536, 128] input).  CTA = 128 columns x one row split; thread = 4 columns, 8 row lanes,
// 4 rows in flight; partials [nsplit][cols] are summed in a fixed order -> deterministic.
// ---- LayerNorm backward, bulk-copy pipelined (d = 128 / 256 / 512) ------------------------------------------------------------------
// The register-pipelined kernel above is latency-bound (ncu, rows = 65 536, d = 128: 24.8 us = 2.1 TB/s, long_scoreboard 7.2 of 11 stalled
// warps per issue, 30 % of the warps resident): a warp issues the 8-byte loads of four rows, waits for them, reduces, stores, and only
// then issues the next loads; 80 registers allow three CTAs per SM.  Here the x / grad rows of a tile (32 / 16 / 8 rows = 8 .. 16 KB per
// tensor, contiguous in memory) are brought in by cp.async.bulk (one elected thread, no tensor map needed) into a 4-stage shared-memory
// ring with full / empty mbarriers, three tiles ahead of the warps that consume them; mean / rstd of the next rows are fetched before the
// wait.  Arithmetic per row and the dgamma / dbeta reduction order are those of the kernel above.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <typename T> __device__ __forceinline__ float4 lds_vec4(const T* p);
template <> __device__ __forceinline__ float4 lds_vec4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 lds_vec4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    return make_float4(bf16_lo_to_float(u.x), bf16_hi_to_float(u.x), bf16_lo_to_float(u.y), bf16_hi_to_float(u.y));
}
template <typename T> __device__ __forceinline__ float round_as(float v);          // v rounded to storage type T, as fp32
template <> __device__ __forceinline__ float round_as<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_as<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
constexpr int LNB_STAGES = 4;
template <typename TG, typename TIN, int NSEG> struct LnBulk {
    static constexpr int D = NSEG * 128;
    static constexpr int TR = 32 / NSEG;                 // rows per tile
    static constexpr int RW = TR / 8;                    // rows per warp and tile (4 / 2 / 1)
    static constexpr uint32_t XB = TR * D * sizeof(TIN), GB = TR * D * sizeof(TG), SB = XB + GB;
    static constexpr uint32_t SMEM = LNB_STAGES * SB + 2 * LNB_STAGES * 8;
};
template <typename TG, typename TIN, int NSEG>
__global__ void __launch_bounds__(256)
layernorm_bwd_bulk_kernel(const TG* __restrict__ g, const TIN* __restrict__ x, const float* __restrict__ gamma,
                          const float* __restrict__ mean_in, const float* __restrict__ rstd_in, long long rows,
                          const TIN* __restrict__ gadd, TIN* __restrict__ dx, float* __restrict__ part /* [gridDim.x][2][d] */) {
    // gadd != nullptr: the gradient that reaches the LayerNorm's INPUT along the residual path (the rows were also consumed by the next
    // residual add) is added into dx here instead of by a separate add kernel; it rides in the ring as a third tile.
    using L = LnBulk<TG, TIN, NSEG>;
    constexpr int D = L::D, TR = L::TR, RW = L::RW, ST = LNB_STAGES;
    extern __shared__ __align__(128) uint8_t lsm[];
    __shared__ float4 red[8][NSEG * 32];
    const uint32_t sm0 = ptx::smem_u32(lsm);
    const uint32_t SBk = gadd != nullptr ? L::SB + L::XB : L::SB;       // stage = [x tile | g tile | gadd tile]
    const uint32_t bar0 = sm0 + ST * SBk;
    auto full = [&](int s_) { return bar0 + 8u * s_; };
    auto empty = [&](int s_) { return bar0 + 8u * (ST + s_); };
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s_ = 0; s_ < ST; ++s_) { ptx::mbar_init(full(s_), 1); ptx::mbar_init(empty(s_), 8); }
        ptx::fence_mbar_init();
    }
    __syncthreads();
    const long long ntiles = (rows + TR - 1) / TR;
    const int my_n = blockIdx.x < ntiles ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    auto tile_of = [&](int i) { return (long long)blockIdx.x + (long long)i * gridDim.x; };
    auto issue = [&](int i) {                                    // thread 0: rows of local tile i -> ring slot i % ST
        const int slot = i % ST;
        if (i >= ST) ptx::mbar_wait(empty(slot), (((uint32_t)(i / ST)) & 1u) ^ 1u);      // all 8 warps have left the slot's previous tile
        const long long r0 = tile_of(i) * TR;
        const uint32_t vr = (uint32_t)((rows - r0) < TR ? (rows - r0) : TR);
        ptx::mbar_arrive_expect_tx(full(slot), vr * D * (uint32_t)(sizeof(TIN) * (gadd != nullptr ? 2 : 1) + sizeof(TG)));
        bulk_g2s(sm0 + slot * SBk, x + r0 * D, vr * D * (uint32_t)sizeof(TIN), full(slot));
        bulk_g2s(sm0 + slot * SBk + L::XB, g + r0 * D, vr * D * (uint32_t)sizeof(TG), full(slot));
        if (gadd != nullptr) bulk_g2s(sm0 + slot * SBk + L::SB, gadd + r0 * D, vr * D * (uint32_t)sizeof(TIN), full(slot));
    };
    if (tid == 0)
        for (int i = 0; i < ST - 1 && i < my_n; ++i) issue(i);

    float4 gam[NSEG], dg[NSEG], db[NSEG];
#pragma unroll
    for (int s_ = 0; s_ < NSEG; ++s_) {
        dg[s_] = make_float4(0.f, 0.f, 0.f, 0.f); db[s_] = dg[s_];
        gam[s_] = __ldg(reinterpret_cast<const float4*>(gamma + s_ * 128 + lane * 4));
    }
    const float inv_d = 1.0f / (float)D;
    for (int i = 0; i < my_n; ++i) {
        const int slot = i % ST;
        if (tid == 0 && i + ST - 1 < my_n) issue(i + ST - 1);
        const long long r0 = tile_of(i) * TR + wid * RW;
        float mean[RW], rs[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) {                           // issued before the wait: their latency hides behind it
            const bool ok = r0 + r < rows;
            mean[r] = ok ? __ldg(mean_in + r0 + r) : 0.f;
            rs[r] = ok ? __ldg(rstd_in + r0 + r) : 0.f;
        }
        ptx::mbar_wait(full(slot), ((uint32_t)(i / ST)) & 1u);
        const TIN* xs_ = reinterpret_cast<const TIN*>(lsm + (size_t)slot * SBk) + (size_t)wid * RW * D;
        const TG* gs_ = reinterpret_cast<const TG*>(lsm + (size_t)slot * SBk + L::XB) + (size_t)wid * RW * D;
        const TIN* as_ = reinterpret_cast<const TIN*>(lsm + (size_t)slot * SBk + L::SB) + (size_t)wid * RW * D;
        float4 xh[RW][NSEG], gg[RW][NSEG];
        float s1[RW], s2[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const bool ok = r0 + r < rows;
            s1[r] = 0.f; s2[r] = 0.f;
#pragma unroll
            for (int s_ = 0; s_ < NSEG; ++s_) {
                xh[r][s_] = make_float4(0.f, 0.f, 0.f, 0.f); gg[r][s_] = xh[r][s_];
                if (ok) {
                    const float4 xv = lds_vec4<TIN>(xs_ + (size_t)r * D + s_ * 128 + lane * 4);
                    const float4 gv = lds_vec4<TG>(gs_ + (size_t)r * D + s_ * 128 + lane * 4);
                    xh[r][s_] = make_float4((xv.x - mean[r]) * rs[r], (xv.y - mean[r]) * rs[r], (xv.z - mean[r]) * rs[r], (xv.w - mean[r]) * rs[r]);
                    dg[s_].x += gv.x * xh[r][s_].x; dg[s_].y += gv.y * xh[r][s_].y; dg[s_].z += gv.z * xh[r][s_].z; dg[s_].w += gv.w * xh[r][s_].w;
                    db[s_].x += gv.x; db[s_].y += gv.y; db[s_].z += gv.z; db[s_].w += gv.w;
                    gg[r][s_] = make_float4(gv.x * gam[s_].x, gv.y * gam[s_].y, gv.z * gam[s_].z, gv.w * gam[s_].w);
                    s1[r] += (gg[r][s_].x + gg[r][s_].y) + (gg[r][s_].z + gg[r][s_].w);
                    s2[r] += (gg[r][s_].x * xh[r][s_].x + gg[r][s_].y * xh[r][s_].y) + (gg[r][s_].z * xh[r][s_].z + gg[r][s_].w * xh[r][s_].w);
                }
            }
        }
        __syncwarp();
        if (gadd == nullptr && lane == 0) ptx::mbar_arrive(empty(slot));      // this warp has read its rows of the slot
#pragma unroll
        for (int r = 0; r < RW; ++r) { s1[r] = warp_sum(s1[r]) * inv_d; s2[r] = warp_sum(s2[r]) * inv_d; }
#pragma unroll
        for (int r = 0; r < RW; ++r)
            if (r0 + r < rows) {
#pragma unroll
                for (int s_ = 0; s_ < NSEG; ++s_) {
                    float4 o;
                    o.x = rs[r] * (gg[r][s_].x - s1[r] - xh[r][s_].x * s2[r]); o.y = rs[r] * (gg[r][s_].y - s1[r] - xh[r][s_].y * s2[r]);
                    o.z = rs[r] * (gg[r][s_].z - s1[r] - xh[r][s_].z * s2[r]); o.w = rs[r] * (gg[r][s_].w - s1[r] - xh[r][s_].w * s2[r]);
                    if (gadd != nullptr) {
                        const float4 av = lds_vec4<TIN>(as_ + (size_t)r * D + s_ * 128 + lane * 4);
                        o.x += av.x; o.y += av.y; o.z += av.z; o.w += av.w;
                    }
                    Vec4<TIN>::store(dx + (r0 + r) * D + s_ * 128 + lane * 4, o);
                }
            }
        if (gadd != nullptr) {                                   // the gadd tile was read last: hand the slot back only now
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(empty(slot));
        }
    }
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int s_ = 0; s_ < NSEG; ++s_) red[wid][s_ * 32 + lane] = pass == 0 ? dg[s_] : db[s_];
        __syncthreads();
        for (int j = threadIdx.x; j < NSEG * 32; j += 256) {
            float4 a = red[0][j];
#pragma unroll
            for (int w = 1; w < 8; ++w) { const float4 b = red[w][j]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
            reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * 2 + pass) * D)[j] = a;
        }
        __syncthreads();
    }
}

// LayerNorm forward with the same ring: x tiles arrive by cp.async.bulk three tiles ahead; per-row arithmetic is that of
// layernorm_fwd_kernel (two-pass mean / variance on the registers of the row).
template <typename TIN, int NSEG> struct LnFwdBulk {
    static constexpr int D = NSEG * 128;
    static constexpr int TR = 32 / NSEG;
    static constexpr int RW = TR / 8;
    static constexpr uint32_t SB = TR * D * sizeof(TIN);
    static constexpr uint32_t SMEM = LNB_STAGES * SB + 2 * LNB_STAGES * 8;
};
template <typename TIN, typename TOUT, int NSEG>
__global__ void __launch_bounds__(256)
layernorm_fwd_bulk_kernel(const TIN* __restrict__ x, const TIN* __restrict__ addend, TIN* __restrict__ sum_out, const float* __restrict__ gamma,
                          const float* __restrict__ beta, long long rows, float eps, TOUT* __restrict__ y, __nv_bfloat16* __restrict__ y_op,
                          float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    // addend != nullptr: the residual add in front of the LayerNorm (t = t + attn(...); n = norm(t), t2i_moe_gan.py:545-561) happens here:
    // v = round_TIN(x + addend) is written to sum_out and normalised - one pass over the rows less than add + LayerNorm, one launch less.
    using L = LnFwdBulk<TIN, NSEG>;
    constexpr int D = L::D, TR = L::TR, RW = L::RW, ST = LNB_STAGES;
    extern __shared__ __align__(128) uint8_t lsm[];
    const uint32_t sm0 = ptx::smem_u32(lsm);
    const uint32_t SBk = addend != nullptr ? 2 * L::SB : L::SB;           // stage = [x tile | addend tile]
    const uint32_t bar0 = sm0 + ST * SBk;
    auto full = [&](int s_) { return bar0 + 8u * s_; };
    auto empty = [&](int s_) { return bar0 + 8u * (ST + s_); };
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s_ = 0; s_ < ST; ++s_) { ptx::mbar_init(full(s_), 1); ptx::mbar_init(empty(s_), 8); }
        ptx::fence_mbar_init();
    }
    __syncthreads();
    const long long ntiles = (rows + TR - 1) / TR;
    const int my_n = blockIdx.x < ntiles ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    auto tile_of = [&](int i) { return (long long)blockIdx.x + (long long)i * gridDim.x; };
    auto issue = [&](int i) {
        const int slot = i % ST;
        if (i >= ST) ptx::mbar_wait(empty(slot), (((uint32_t)(i / ST)) & 1u) ^ 1u);
        const long long r0 = tile_of(i) * TR;
        const uint32_t vr = (uint32_t)((rows - r0) < TR ? (rows - r0) : TR);
        ptx::mbar_arrive_expect_tx(full(slot), vr * D * (uint32_t)sizeof(TIN) * (addend != nullptr ? 2u : 1u));
        bulk_g2s(sm0 + slot * SBk, x + r0 * D, vr * D * (uint32_t)sizeof(TIN), full(slot));
        if (addend != nullptr) bulk_g2s(sm0 + slot * SBk + L::SB, addend + r0 * D, vr * D * (uint32_t)sizeof(TIN), full(slot));
    };
    if (tid == 0)
        for (int i = 0; i < ST - 1 && i < my_n; ++i) issue(i);
    float4 gam[NSEG], bet[NSEG];
#pragma unroll
    for (int s_ = 0; s_ < NSEG; ++s_) {
        gam[s_] = __ldg(reinterpret_cast<const float4*>(gamma + s_ * 128 + lane * 4));
        bet[s_] = __ldg(reinterpret_cast<const float4*>(beta + s_ * 128 + lane * 4));
    }
    const float inv_d = 1.0f / (float)D;
    for (int i = 0; i < my_n; ++i) {
        const int slot = i % ST;
        if (tid == 0 && i + ST - 1 < my_n) issue(i + ST - 1);
        const long long r0 = tile_of(i) * TR + wid * RW;
        ptx::mbar_wait(full(slot), ((uint32_t)(i / ST)) & 1u);
        const TIN* xs_ = reinterpret_cast<const TIN*>(lsm + (size_t)slot * SBk) + (size_t)wid * RW * D;
        const TIN* as_ = reinterpret_cast<const TIN*>(lsm + (size_t)slot * SBk + L::SB) + (size_t)wid * RW * D;
        float4 v[RW][NSEG];
        float sum[RW], mean[RW], rstd[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            sum[r] = 0.f;
#pragma unroll
            for (int s_ = 0; s_ < NSEG; ++s_) {
                v[r][s_] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + r < rows) {
                    v[r][s_] = lds_vec4<TIN>(xs_ + (size_t)r * D + s_ * 128 + lane * 4);
                    if (addend != nullptr) {
                        const float4 av = lds_vec4<TIN>(as_ + (size_t)r * D + s_ * 128 + lane * 4);
                        v[r][s_] = make_float4(round_as<TIN>(v[r][s_].x + av.x), round_as<TIN>(v[r][s_].y + av.y),
                                               round_as<TIN>(v[r][s_].z + av.z), round_as<TIN>(v[r][s_].w + av.w));
                        Vec4<TIN>::store(sum_out + (r0 + r) * D + s_ * 128 + lane * 4, v[r][s_]);
                    }
                }
                sum[r] += (v[r][s_].x + v[r][s_].y) + (v[r][s_].z + v[r][s_].w);
            }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(empty(slot));
#pragma unroll
        for (int r = 0; r < RW; ++r) mean[r] = warp_sum(sum[r]) * inv_d;
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            float sq = 0.f;
#pragma unroll
            for (int s_ = 0; s_ < NSEG; ++s_) {
                const float a = v[r][s_].x - mean[r], b = v[r][s_].y - mean[r], c = v[r][s_].z - mean[r], e = v[r][s_].w - mean[r];
                sq += (a * a + b * b) + (c * c + e * e);
            }
            sum[r] = sq;
        }
#pragma unroll
        for (int r = 0; r < RW; ++r) rstd[r] = rsqrtf(warp_sum(sum[r]) * inv_d + eps);
#pragma unroll
        for (int r = 0; r < RW; ++r)
            if (r0 + r < rows) {
#pragma unroll
                for (int s_ = 0; s_ < NSEG; ++s_) {
                    const int c = s_ * 128 + lane * 4;
                    float4 o;
                    o.x = (v[r][s_].x - mean[r]) * rstd[r] * gam[s_].x + bet[s_].x; o.y = (v[r][s_].y - mean[r]) * rstd[r] * gam[s_].y + bet[s_].y;
                    o.z = (v[r][s_].z - mean[r]) * rstd[r] * gam[s_].z + bet[s_].z; o.w = (v[r][s_].w - mean[r]) * rstd[r] * gam[s_].w + bet[s_].w;
                    Vec4<TOUT>::store(y + (r0 + r) * D + c, o);
                    if (y_op != nullptr) Vec4<__nv_bfloat16>::store(y_op + (r0 + r) * D + c, o);
                }
            }
        if (lane < RW && r0 + lane < rows) {
            float m = mean[0], q = rstd[0];
#pragma unroll
            for (int r = 1; r < RW; ++r) if (lane == r) { m = mean[r]; q = rstd[r]; }
            mean_out[r0 + lane] = m; rstd_out[r0 + lane] = q;
        }
    }
}

// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
colsum_rows_kernel(const T* __restrict__ x, long long rows, int cols, int nsplit, float* __restrict__ part) {
    __shared__ float4 red[8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 128 + cx * 4;
    const long long per = (rows + nsplit - 1) / nsplit;
    const long long r0 = (long long)blockIdx.y * per, r1 = r0 + per < rows ? r0 + per : rows;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    if (c < cols) {
        long long r = r0 + ry;
        for (; r + 24 < r1; r += 32) {
            const float4 v0 = Vec4<T>::load(x + r * cols + c), v1 = Vec4<T>::load(x + (r + 8) * cols + c);
            const float4 v2 = Vec4<T>::load(x + (r + 16) * cols + c), v3 = Vec4<T>::load(x + (r + 24) * cols + c);
            a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;  a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
            a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;  a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
        }
        for (; r < r1; r += 8) { const float4 v0 = Vec4<T>::load(x + r * cols + c); a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w; }
    }
    red[ry][cx] = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w));
    __syncthreads();
    if (ry == 0 && c < cols) {
        float4 a = red[0][cx];
#pragma unroll
        for (int w = 1; w < 8; ++w) { const float4 b = red[w][cx]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
        *reinterpret_cast<float4*>(part + (size_t)blockIdx.y * cols + c) = a;
    }
}

// =====================================================================================================
// channel_scale: out[b, p, c] = lrelu(x[b, p, c] * s[b, c], slope)      x: [B, P, C] (NHWC), s: [B, C] fp32
// =====================================================================================================
// forward: grid.y = sample, flat over the sample's groups of 8 channels, CS_U groups (256 apart) per thread, all loads issued before
// the first use.  Index arithmetic is 32-bit and per sample: the first version decoded (b, p, c) from a flat 64-bit index with two
// 64-bit divisions per group - ~200 instructions next to two 16-byte loads.
constexpr int CS_U = 2;
template <typename T, typename TS>
__global__ void __launch_bounds__(256)
channel_scale_fwd_kernel(const T* __restrict__ x, const TS* __restrict__ s, unsigned n8s /* P * C / 8 */, int C, float slope, T* __restrict__ out) {
    const unsigned c8 = (unsigned)C / 8u;
    const unsigned j0 = blockIdx.x * (256u * CS_U) + threadIdx.x;
    const size_t base = (size_t)blockIdx.y * n8s;
    const TS* sb = s + (size_t)blockIdx.y * C;
    Raw8<T> rx[CS_U];
    Raw8<TS> rs[CS_U];
#pragma unroll
    for (int u = 0; u < CS_U; ++u) {                       // raw loads from clamped indices: no branch, no conversion between them
        const unsigned j = min(j0 + u * 256u, n8s - 1u);
        rx[u].load(x + (base + j) * 8);
        rs[u].load(sb + (j % c8) * 8u);
    }
#pragma unroll
    for (int u = 0; u < CS_U; ++u) {
        const unsigned j = j0 + u * 256u;
        if (j < n8s) {
            float v[CS_U][8], sc[CS_U][8];
            rx[u].get(true, v[u]); rs[u].get(true, sc[u]);
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float t = v[u][e] * sc[u][e]; o[e] = t > 0.f ? t : t * slope; }
            Vec8<T>::store(out + (base + j) * 8, o);
        }
    }
}
// backward: CTA = (sample b, 128-channel group, pixel split); half-warp = one pixel (16 lanes x 8 channels), so the 16 half-warps
// take pixels h, h + 16, ... two at a time.
//   m = g * (x*s > 0 ? 1 : slope);  dx = m * s;  ds[b, c] = sum_p m * x   (partials over the pixel splits, summed in order)
template <typename T, typename TS>
__global__ void __launch_bounds__(256)
channel_scale_bwd_kernel(const T* __restrict__ g, const T* __restrict__ x, const TS* __restrict__ s, int P, int C, float slope,
                         int psplit, T* __restrict__ dx, float* __restrict__ ds_part /* [psplit][B][C] */, int Bn) {
    __shared__ float red[16][128];
    const int b = blockIdx.y, cgp = blockIdx.x, sp = blockIdx.z;
    const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15;
    const int c = cgp * 128 + l16 * 8;
    float sc[8];
    Vec8<TS>::load(s + (size_t)b * C + c, sc);
    const int per = (P + psplit - 1) / psplit;
    const int p0 = sp * per, p1 = min(P, p0 + per);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int p = p0 + hw; p < p1; p += 32) {
        const bool two = p + 16 < p1;
        const size_t o0 = ((size_t)b * P + p) * C + c, o1 = o0 + (size_t)16 * C;
        float g0[8], x0[8], g1[8], x1[8];
        Vec8<T>::load(g + o0, g0); Vec8<T>::load(x + o0, x0);
        if (two) { Vec8<T>::load(g + o1, g1); Vec8<T>::load(x + o1, x1); }
        float d0[8], d1[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float m = x0[e] * sc[e] > 0.f ? g0[e] : g0[e] * slope;
            d0[e] = m * sc[e];
            acc[e] += m * x0[e];
        }
        Vec8<T>::store(dx + o0, d0);
        if (two) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float m = x1[e] * sc[e] > 0.f ? g1[e] : g1[e] * slope;
                d1[e] = m * sc[e];
                acc[e] += m * x1[e];
            }
            Vec8<T>::store(dx + o1, d1);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[hw][l16 * 8 + e] = acc[e];
    __syncthreads();
    if (threadIdx.x < 128) {
        float a = red[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < 16; ++w) a += red[w][threadIdx.x];
        ds_part[((size_t)sp * Bn + b) * C + cgp * 128 + threadIdx.x] = a;
    }
}

// =====================================================================================================
// demod_coef: the demodulation coefficients of ModulatedConv (t2i_moe_gan.py:165-170) in "scale the activations" form,
//     wsq[o, i] = sum_taps w[o, i, ky, kx]^2        d[b, o] = rsqrt( sum_i s[b, i]^2 * wsq[o, i] + eps )
// In torch ops this is ~7 launches forward (float, pow, sum, pow, an fp32 SIMT sgemm, add, rsqrt) and ~12 backward per layer - 32 + 16
// times per train step, ~1.5 ms of 3 us kernels.  Here: one pass over the weights + one small tiled fp32 GEMM forward; two GEMMs
// backward whose epilogues write ds (in s's dtype) and dw (in the weight's own strides, contiguous or channels-last).
//     t[b, o]  = -0.5 * dd[b, o] * d[b, o]^3
//     ds[b, i] = 2 s[b, i] * sum_o t[b, o] wsq[o, i]        dw[o, i, tap] = 2 w[o, i, tap] * sum_b t[b, o] s[b, i]^2
// =====================================================================================================
struct DemodArgs {
    const void* s; int s_bf16;                 // [B, Cin]
    const float* w; long long so, si, sh, sw;  // [Cout, Cin, KH, KW] with element strides
    const float* wsq;                          // [Cout, Cin]
    const float* d; const float* dd;           // [B, Cout]
    float* d_out; void* ds; float* dw;
    int B, Cin, Cout, KH, KW; float eps;
};
__device__ __forceinline__ float demod_s(const DemodArgs& a, int b, int i) {
    return a.s_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.s)[(size_t)b * a.Cin + i])
                    : __ldg(reinterpret_cast<const float*>(a.s) + (size_t)b * a.Cin + i);
}
__global__ void __launch_bounds__(256)
demod_wsq_kernel(const float* __restrict__ w, long long so, long long si, long long sh, long long sw, int Cout, int Cin, int KH, int KW,
                 float* __restrict__ wsq) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= Cout * Cin) return;
    const int o = idx / Cin, i = idx - o * Cin;
    const float* p = w + o * so + i * si;
    float acc = 0.f;
    for (int ky = 0; ky < KH; ++ky)
        for (int kx = 0; kx < KW; ++kx) { const float v = __ldg(p + ky * sh + kx * sw); acc = fmaf(v, v, acc); }
    wsq[idx] = acc;
}
// C[M, N] = A[M, K] * B[K, N], 32 x 32 tiles, K in chunks of 64, 256 threads (4 rows x 1 column each), fp32 FMA in k order.  The
// operands are a few hundred KB (L2-resident) and the grid is 64 - 256 CTAs, so what matters is latency: the next chunk's 16 operand
// values per thread are fetched into registers before the current chunk is multiplied.
//   MODE 0 (forward):  M = B,    N = Cout, K = Cin:  A = s^2, B = wsq^T      -> d  = rsqrt(C + eps)
//   MODE 1 (ds):       M = B,    N = Cin,  K = Cout: A = t,   B = wsq        -> ds = 2 s C
//   MODE 2 (dw):       M = Cout, N = Cin,  K = B:    A = t^T, B = s^2        -> dw = 2 w C   (all taps)
template <int MODE>
__global__ void __launch_bounds__(256)
demod_gemm_kernel(const DemodArgs a) {
    constexpr int BK = 64, LDK = BK + 4;        // row stride 68 floats: 16-byte aligned rows, a quarter-warp's LDS.128 hits 32 distinct banks
    __shared__ __align__(16) float As[32][LDK], Bs[32][LDK];      // both [row][k]: the inner loop reads 4 k values per LDS.128
    const int M = MODE == 2 ? a.Cout : a.B, N = MODE == 0 ? a.Cout : a.Cin, K = MODE == 0 ? a.Cin : (MODE == 1 ? a.Cout : a.B);
    const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // Operand fetch in two halves: LOAD (raw values from clamped, always-valid addresses - no arithmetic, no branches, so the 16 - 24 loads
    // of a chunk are all in flight together) and, one chunk later, CONVERT + store to shared memory (square / t = -0.5 dd d^3, zero outside
    // the matrix).  With the arithmetic inside the fetch (first version) every load was followed by its use and the chunk's loads ran one
    // after the other: 5.7 us per 64-wide chunk at every size.
    const float* sf = reinterpret_cast<const float*>(a.s);
    const unsigned short* sh = reinterpret_cast<const unsigned short*>(a.s);
    auto load_s = [&](int b, int i) -> float {          // raw: the fp32 value, or the bf16 bits in the low half (widened by s_val below)
        const size_t off = (size_t)b * a.Cin + i;
        return a.s_bf16 ? __uint_as_float((unsigned)sh[off]) : __ldg(sf + off);
    };
    auto s_val = [&](float raw) -> float { return a.s_bf16 ? __uint_as_float(__float_as_uint(raw) << 16) : raw; };
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float ra[8], ra2[8], rb[8];
    auto load_chunk = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = threadIdx.x + 256 * j;
            if (MODE == 0) {            // A(b, i) = s^2 (i fastest);  B(i, o) = wsq[o, i] (i fastest)
                const int m = min(m0 + (e >> 6), M - 1), k = min(k0 + (e & 63), K - 1), n = min(n0 + (e >> 6), N - 1);
                ra[j] = load_s(m, k);
                rb[j] = __ldg(a.wsq + (size_t)n * a.Cin + k);
            } else if (MODE == 1) {     // A(b, o) = t (o fastest);  B(o, i) = wsq[o, i] (i fastest)
                const int m = min(m0 + (e >> 6), M - 1), k = min(k0 + (e & 63), K - 1);
                ra[j] = __ldg(a.d + (size_t)m * a.Cout + k); ra2[j] = __ldg(a.dd + (size_t)m * a.Cout + k);
                const int kk = min(k0 + (e >> 5), K - 1), n = min(n0 + (e & 31), N - 1);
                rb[j] = __ldg(a.wsq + (size_t)kk * a.Cin + n);
            } else {                    // A(o, b) = t[b, o] (o fastest);  B(b, i) = s^2 (i fastest)
                const int m = min(m0 + (e & 31), M - 1), kk = min(k0 + (e >> 5), K - 1), n = min(n0 + (e & 31), N - 1);
                ra[j] = __ldg(a.d + (size_t)kk * a.Cout + m); ra2[j] = __ldg(a.dd + (size_t)kk * a.Cout + m);
                rb[j] = load_s(kk, n);
            }
        }
    };
    auto store_chunk = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = threadIdx.x + 256 * j;
            if (MODE == 0) {
                const bool oka = m0 + (e >> 6) < M && k0 + (e & 63) < K, okb = n0 + (e >> 6) < N && k0 + (e & 63) < K;
                const float sv = s_val(ra[j]);
                As[e >> 6][e & 63] = oka ? sv * sv : 0.f;
                Bs[e >> 6][e & 63] = okb ? rb[j] : 0.f;
            } else if (MODE == 1) {
                const bool oka = m0 + (e >> 6) < M && k0 + (e & 63) < K, okb = k0 + (e >> 5) < K && n0 + (e & 31) < N;
                As[e >> 6][e & 63] = oka ? -0.5f * ra2[j] * ra[j] * ra[j] * ra[j] : 0.f;
                Bs[e & 31][e >> 5] = okb ? rb[j] : 0.f;
            } else {
                const bool oka = m0 + (e & 31) < M && k0 + (e >> 5) < K, okb = k0 + (e >> 5) < K && n0 + (e & 31) < N;
                As[e & 31][e >> 5] = oka ? -0.5f * ra2[j] * ra[j] * ra[j] * ra[j] : 0.f;
                const float sv = s_val(rb[j]);
                Bs[e & 31][e >> 5] = okb ? sv * sv : 0.f;
            }
        }
    };
    if (K > 0) load_chunk(0);
    for (int k0 = 0; k0 < K; k0 += BK) {
        store_chunk(k0);
        __syncthreads();
        if (k0 + BK < K) load_chunk(k0 + BK);
        // one shared-memory load per FMA (the first version) made this loop LSU-bound; 5 LDS.128 now feed 16 FMAs
#pragma unroll 4
        for (int kk = 0; kk < BK; kk += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[tx][kk]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[ty + 8 * j][kk]);      // same address in the whole warp: broadcast
                acc[j] = fmaf(a4.w, b4.w, fmaf(a4.z, b4.z, fmaf(a4.y, b4.y, fmaf(a4.x, b4.x, acc[j]))));
            }
        }
        __syncthreads();
    }
    const int n = n0 + tx;
    if (n >= N) return;
    if (MODE == 2 && a.KH * a.KW <= 9) {
        // all weight taps of the thread's 4 outputs are loaded before the first product (a load-multiply-store loop over the taps ran its
        // 36 global loads one after the other: 12 us of the 25 us this launch took at 128 x 128 x 3 x 3)
        const int taps = a.KH * a.KW;
        float wv[4][9];
        long long toff[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) { const int ky = t / a.KW, kx = t - ky * a.KW; toff[t] = t < taps ? ky * a.sh + kx * a.sw : 0; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = min(m0 + ty + 8 * j, M - 1);
#pragma unroll
            for (int t = 0; t < 9; ++t) wv[j][t] = __ldg(a.w + m * a.so + n * a.si + toff[t]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty + 8 * j;
            if (m >= M) continue;
            const float v = 2.0f * acc[j];
#pragma unroll
            for (int t = 0; t < 9; ++t)
                if (t < taps) a.dw[m * a.so + n * a.si + toff[t]] = v * wv[j][t];
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int m = m0 + ty + 8 * j;
        if (m >= M) continue;
        if (MODE == 0) {
            a.d_out[(size_t)m * a.Cout + n] = rsqrtf(acc[j] + a.eps);
        } else if (MODE == 1) {
            const float v = 2.0f * demod_s(a, m, n) * acc[j];
            if (a.s_bf16) reinterpret_cast<__nv_bfloat16*>(a.ds)[(size_t)m * a.Cin + n] = __float2bfloat16_rn(v);
            else          reinterpret_cast<float*>(a.ds)[(size_t)m * a.Cin + n] = v;
        } else {
            const float v = 2.0f * acc[j];
            const long long base = m * a.so + n * a.si;
            for (int ky = 0; ky < a.KH; ++ky)
                for (int kx = 0; kx < a.KW; ++kx) { const long long off = base + ky * a.sh + kx * a.sw; a.dw[off] = v * __ldg(a.w + off); }
        }
    }
}

// =====================================================================================================
// offset_sample (NHWC): out[b, i, j, :] = bilinear(x[b], grid(b, i, j)),  grid = clamp(base + 0.05 * off, -1, 1)
// =====================================================================================================
struct Tap {            // the 4 bilinear taps of one output pixel
    int idx[4];         // pixel index y * W + x inside the sample, -1 = outside (zero padding)
    float w[4];
    float wx, wy;       // fractional parts
    float sx, sy;       // d(off) -> d(ix) chain: 0.05 * W / 2 where the clamp passes gradient, else 0
};
__device__ __forceinline__ Tap make_tap(float offx, float offy, int i, int j, int H, int W) {
    Tap t;
    const float bx = W > 1 ? -1.0f + 2.0f * (float)j / (float)(W - 1) : -1.0f;      // torch.linspace(-1, 1, W)
    const float by = H > 1 ? -1.0f + 2.0f * (float)i / (float)(H - 1) : -1.0f;
    float gx = bx + 0.05f * offx, gy = by + 0.05f * offy;
    t.sx = (gx >= -1.0f && gx <= 1.0f) ? 0.05f * 0.5f * (float)W : 0.0f;
    t.sy = (gy >= -1.0f && gy <= 1.0f) ? 0.05f * 0.5f * (float)H : 0.0f;
    gx = fminf(fmaxf(gx, -1.0f), 1.0f); gy = fminf(fmaxf(gy, -1.0f), 1.0f);
    const float ix = ((gx + 1.0f) * (float)W - 1.0f) * 0.5f;                         // align_corners = False
    const float iy = ((gy + 1.0f) * (float)H - 1.0f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    t.wx = ix - fx; t.wy = iy - fy;
    const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
    t.idx[0] = (vy0 && vx0) ? y0 * W + x0 : -1;           t.w[0] = (1.0f - t.wx) * (1.0f - t.wy);
    t.idx[1] = (vy0 && vx1) ? y0 * W + x0 + 1 : -1;       t.w[1] = t.wx * (1.0f - t.wy);
    t.idx[2] = (vy1 && vx0) ? (y0 + 1) * W + x0 : -1;     t.w[2] = (1.0f - t.wx) * t.wy;
    t.idx[3] = (vy1 && vx1) ? (y0 + 1) * W + x0 + 1 : -1; t.w[3] = t.wx * t.wy;
    return t;
}

// All three kernels were issue-bound in their first form (ncu, B = 256, 16 x 16, C = 128: 61 - 76 % issue-active at 1.3 - 2 TB/s): a
// warp per pixel recomputed the taps in all 32 lanes (~300 warp instructions for 256 bytes of output) and moved 8 bytes per lane.
// Now: CTA = a tile of OS_TP pixels of one sample (grid.y = sample); one THREAD per pixel computes the taps into shared memory, then
// the CTA's threads walk (pixel, 8-channel group) items with 16-byte accesses, two items in flight.
constexpr int OS_TP = 64;
struct TapTile {
    int4 idx[OS_TP];        // pixel index of the 4 taps inside the sample, -1 = outside
    float4 w[OS_TP];        // bilinear weights
    float4 f[OS_TP];        // wx, wy, sx, sy
};
__device__ __forceinline__ int tap_tile_fill(TapTile& tt, const float* __restrict__ off, int b, int p0, int H, int W) {
    const int HW = H * W;
    const int np = min(OS_TP, HW - p0);
    if ((int)threadIdx.x < np) {
        const int p = p0 + threadIdx.x;
        const float2 o = __ldg(reinterpret_cast<const float2*>(off + ((size_t)b * HW + p) * 2));
        const Tap t = make_tap(o.x, o.y, p / W, p % W, H, W);
        tt.idx[threadIdx.x] = make_int4(t.idx[0], t.idx[1], t.idx[2], t.idx[3]);
        tt.w[threadIdx.x] = make_float4(t.w[0], t.w[1], t.w[2], t.w[3]);
        tt.f[threadIdx.x] = make_float4(t.wx, t.wy, t.sx, t.sy);
    }
    __syncthreads();
    return np;
}
template <typename T>
__device__ __forceinline__ void load_taps_raw(const T* __restrict__ xc, int C, const int4 id, Raw8<T> (&r)[4]) {
    r[0].load(xc + (size_t)max(id.x, 0) * C); r[1].load(xc + (size_t)max(id.y, 0) * C);
    r[2].load(xc + (size_t)max(id.z, 0) * C); r[3].load(xc + (size_t)max(id.w, 0) * C);
}
template <typename T>
__device__ __forceinline__ void taps_to_float(const Raw8<T> (&r)[4], const int4 id, float (&v)[4][8]) {
    r[0].get(id.x >= 0, v[0]); r[1].get(id.y >= 0, v[1]); r[2].get(id.z >= 0, v[2]); r[3].get(id.w >= 0, v[3]);
}
// forward: out[p, c] = sum_k w_k * x[tap_k, c]   (fma chain over the taps in order, as before).  U = items in flight per thread:
// U = 2 needs 94 registers (2 CTAs per SM), U = 1 fits 4 CTAs per SM.
template <typename T, int U>
__global__ void __launch_bounds__(256, U == 1 ? 4 : 2)
offset_sample_fwd_kernel(const T* __restrict__ x, const float* __restrict__ off, int H, int W, int C, T* __restrict__ out) {
    __shared__ TapTile tt;
    const int b = blockIdx.y, p0 = blockIdx.x * OS_TP, HW = H * W;
    const int np = tap_tile_fill(tt, off, b, p0, H, W);
    const unsigned c8 = (unsigned)C / 8u, n_items = (unsigned)np * c8;
    const T* xb = x + (size_t)b * HW * C;
    T* ob = out + ((size_t)b * HW + p0) * C;
    for (unsigned it0 = threadIdx.x; it0 < n_items; it0 += 256 * U) {
        Raw8<T> raw[U][4];
        unsigned pl[U], cg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned it = min(it0 + 256u * u, n_items - 1);         // clamped: the loads carry no branch
            pl[u] = it / c8; cg[u] = it - pl[u] * c8;
            load_taps_raw<T>(xb + cg[u] * 8, C, tt.idx[pl[u]], raw[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned it = it0 + 256u * u;
            if (it < n_items) {
                float v[4][8];
                taps_to_float<T>(raw[u], tt.idx[pl[u]], v);
                const float4 w = tt.w[pl[u]];
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = fmaf(w.w, v[3][e], fmaf(w.z, v[2][e], fmaf(w.y, v[1][e], w.x * v[0][e])));
                Vec8<T>::store(ob + (size_t)pl[u] * C + cg[u] * 8, o);
            }
        }
    }
}
// backward w.r.t. the offsets: the 16 lanes that share a pixel (C / 8 is a multiple of 16) reduce their 8-channel partial sums with
// shuffles, the C / 128 partials of a pixel are summed in order by the pixel's thread.
template <typename T>
__global__ void __launch_bounds__(256)
offset_sample_bwd_off_kernel(const T* __restrict__ g, const T* __restrict__ x, const float* __restrict__ off, int H, int W, int C,
                             float* __restrict__ doff) {
    __shared__ TapTile tt;
    __shared__ float2 part[OS_TP][8];                    // [pixel][C / 128 <= 8]
    const int b = blockIdx.y, p0 = blockIdx.x * OS_TP, HW = H * W;
    const int np = tap_tile_fill(tt, off, b, p0, H, W);
    const unsigned c8 = (unsigned)C / 8u, n_items = (unsigned)np * c8;       // a multiple of 16
    const T* xb = x + (size_t)b * HW * C;
    const T* gb = g + ((size_t)b * HW + p0) * C;
    for (unsigned base = 0; base < n_items; base += 256) {                   // warp-uniform trip count (shuffles below)
        const unsigned it = base + threadIdx.x;
        const bool ok = it < n_items;
        const unsigned pl = ok ? it / c8 : 0, cg = ok ? it - pl * c8 : 0;
        float gix = 0.f, giy = 0.f;
        if (ok) {
            float gv[8], v[4][8];
            Raw8<T> graw, raw[4];
            graw.load(gb + (size_t)pl * C + cg * 8);
            load_taps_raw<T>(xb + cg * 8, C, tt.idx[pl], raw);
            graw.get(true, gv);
            taps_to_float<T>(raw, tt.idx[pl], v);
            const float4 f = tt.f[pl];
            const float omx = 1.0f - f.x, omy = 1.0f - f.y;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                gix += gv[e] * ((v[1][e] - v[0][e]) * omy + (v[3][e] - v[2][e]) * f.y);
                giy += gv[e] * ((v[2][e] - v[0][e]) * omx + (v[3][e] - v[1][e]) * f.x);
            }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            gix += __shfl_xor_sync(0xffffffffu, gix, o);
            giy += __shfl_xor_sync(0xffffffffu, giy, o);
        }
        if (ok && (threadIdx.x & 15) == 0) part[pl][cg >> 4] = make_float2(gix, giy);
    }
    __syncthreads();
    if ((int)threadIdx.x < np) {
        const int nparts = C / 128;
        float gix = 0.f, giy = 0.f;
        for (int j = 0; j < nparts; ++j) { gix += part[threadIdx.x][j].x; giy += part[threadIdx.x][j].y; }
        const float4 f = tt.f[threadIdx.x];
        *reinterpret_cast<float2*>(doff + ((size_t)b * HW + p0 + threadIdx.x) * 2) = make_float2(gix * f.z, giy * f.w);
    }
}
// backward w.r.t. the input, as a deterministic GATHER.  The scatter pattern (which output pixels touch which input pixel, with
// which weight) is the same for every channel, so it is inverted in shared memory.  A sample's input pixels are split over S CTAs
// (slices of QS <= 64 pixels):
//   1. every output pixel p of the sample computes its 4 taps (each CTA of the sample repeats this: one thread per pixel);
//   2. thread (q, r) - input pixel q of the slice, quarter r of the output pixels - compares the taps of its <= 64 output pixels
//      with q and keeps the matches as a 256-bit mask in registers; the per-(q, r) counts are scanned into list offsets;
//   3. the same thread expands its mask into the contributor list of q, in (p, tap) order -> fixed accumulation order;
//   4. (input pixel, 8-channel group) items: dx[q, c..c+7] = sum_e w_e * g[p_e, c..c+7]  (16-byte row reads, no atomics).
template <typename T>
__global__ void __launch_bounds__(256)
offset_sample_bwd_in_kernel(const T* __restrict__ g, const float* __restrict__ off, int H, int W, int C, int S, int QS, T* __restrict__ dx) {
    extern __shared__ __align__(16) float smem[];
    const int HW = H * W;                                 // <= 256
    int* tidx = reinterpret_cast<int*>(smem);             // [HW][4]
    float* tw = smem + HW * 4;                            // [HW][4]
    int* ent_p = reinterpret_cast<int*>(smem + HW * 8);   // [4 * HW] (every tap of the sample may land in one slice)
    float* ent_w = smem + HW * 12;                        // [4 * HW]
    __shared__ int cnt[4][64];
    __shared__ int start[65];
    const int b = blockIdx.x / S, q0 = (blockIdx.x % S) * QS, tid = threadIdx.x;
    const int nq = min(QS, HW - q0);                      // input pixels of this slice (<= 64)
    if (tid < HW) {
        const float2 o = __ldg(reinterpret_cast<const float2*>(off + ((size_t)b * HW + tid) * 2));
        const Tap t = make_tap(o.x, o.y, tid / W, tid % W, H, W);
        *reinterpret_cast<int4*>(tidx + tid * 4) = make_int4(t.idx[0], t.idx[1], t.idx[2], t.idx[3]);
        *reinterpret_cast<float4*>(tw + tid * 4) = make_float4(t.w[0], t.w[1], t.w[2], t.w[3]);
    }
    __syncthreads();
    const int ql = tid & 63, r = tid >> 6;                // a warp has one r: the tap reads below are broadcasts
    const int PR = (HW + 3) / 4;                          // output pixels per quarter (<= 64)
    const int pr0 = r * PR;
    const int me = ql < nq ? q0 + ql : -2;                // -2 matches nothing (-1 = outside tap)
    unsigned m[8];                                        // 4 bits per output pixel of the quarter
    int n = 0;
#pragma unroll
    for (int wd = 0; wd < 8; ++wd) {
        m[wd] = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int pp = wd * 8 + j;
            if (pp < PR && pr0 + pp < HW) {               // warp-uniform
                const int4 id = *reinterpret_cast<const int4*>(tidx + (pr0 + pp) * 4);
                const unsigned bits = (id.x == me ? 1u : 0u) | (id.y == me ? 2u : 0u) | (id.z == me ? 4u : 0u) | (id.w == me ? 8u : 0u);
                m[wd] |= bits << (4 * j);
            }
        }
        n += __popc(m[wd]);
    }
    cnt[r][ql] = n;
    __syncthreads();
    if (tid < 32) {                                       // exclusive scan of the 64 per-pixel totals by one warp
        const int t0 = cnt[0][tid] + cnt[1][tid] + cnt[2][tid] + cnt[3][tid];
        const int t1 = cnt[0][tid + 32] + cnt[1][tid + 32] + cnt[2][tid + 32] + cnt[3][tid + 32];
        int i0 = t0, i1 = t1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a0 = __shfl_up_sync(0xffffffffu, i0, o), a1 = __shfl_up_sync(0xffffffffu, i1, o);
            if (tid >= o) { i0 += a0; i1 += a1; }
        }
        const int tot0 = __shfl_sync(0xffffffffu, i0, 31);
        start[tid] = i0 - t0;
        start[tid + 32] = tot0 + i1 - t1;
        if (tid == 31) start[64] = tot0 + i1;
    }
    __syncthreads();
    {
        int pos = start[ql];
        for (int rr = 0; rr < r; ++rr) pos += cnt[rr][ql];
#pragma unroll
        for (int wd = 0; wd < 8; ++wd) {
            unsigned w = m[wd];
            while (w) {
                const int bit = __ffs(w) - 1;
                w &= w - 1;
                const int key = (pr0 + wd * 8) * 4 + bit;         // = p * 4 + tap
                ent_p[pos] = key >> 2;
                ent_w[pos] = tw[key];
                ++pos;
            }
        }
    }
    __syncthreads();
    const unsigned c8 = (unsigned)C / 8u, n_items = (unsigned)nq * c8;
    const T* gb = g + (size_t)b * HW * C;
    T* db = dx + ((size_t)b * HW + q0) * C;
    for (unsigned it = tid; it < n_items; it += 256) {
        const unsigned q = it / c8, cg = it - q * c8;
        const int e0 = start[q], e1 = start[q + 1];
        const T* gc = gb + cg * 8;
        float a[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = 0.f;
        int e = e0;
        for (; e + 3 < e1; e += 4) {                      // four rows in flight; the accumulation order stays e0, e0 + 1, ...
            float v[4][8], w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { Vec8<T>::load(gc + (size_t)ent_p[e + j] * C, v[j]); w[j] = ent_w[e + j]; }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 8; ++c) a[c] = fmaf(w[j], v[j][c], a[c]);
        }
        for (; e < e1; ++e) {
            float v[8];
            Vec8<T>::load(gc + (size_t)ent_p[e] * C, v);
            const float w = ent_w[e];
#pragma unroll
            for (int c = 0; c < 8; ++c) a[c] = fmaf(w, v[c], a[c]);
        }
        Vec8<T>::store(db + (size_t)q * C + cg * 8, a);
    }
}

// =====================================================================================================
// upsample2x (NHWC): F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False) of GenerativeBlock (t2i_moe_gan.py:
// 646-648).  src = 0.5 * (dst + 0.5) - 0.5 clamped at 0, second tap clamped at the last pixel - ATen's formula and evaluation
// order.  Flat over 8-channel chunks of the OUTPUT (forward) / of the INPUT (backward, a 4 x 4 gather of the upstream gradient with
// the transposed weights: deterministic, no atomics).  ATen's NHWC kernels needed 70-140 us per call at batch 256.
// =====================================================================================================
__device__ __forceinline__ void up2_src(int dst, int n_in, int& i0, int& i1, float& l1) {
    float src = 0.5f * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    l1 = src - (float)i0;
}
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const T* __restrict__ x, unsigned n8s /* 4 * H * W * C / 8 */, int H, int W, int C, T* __restrict__ out) {
    const unsigned c8 = (unsigned)C / 8u, OW = 2u * W;
    const unsigned j0 = blockIdx.x * (256u * CS_U) + threadIdx.x;          // grid.y = sample; 32-bit index inside the sample
    const T* xs = x + (size_t)blockIdx.y * H * W * C;
    T* os = out + (size_t)blockIdx.y * n8s * 8;
    float v[CS_U][4][8], ly[CS_U], lx[CS_U];
#pragma unroll
    for (int u = 0; u < CS_U; ++u) {
        const unsigned j = j0 + u * 256u;
        if (j < n8s) {
            const unsigned pix = j / c8;
            const int c = (int)(j - pix * c8) * 8;
            const int oy = (int)(pix / OW), ox = (int)(pix - (unsigned)oy * OW);
            int y0, y1, x0, x1;
            up2_src(oy, H, y0, y1, ly[u]);
            up2_src(ox, W, x0, x1, lx[u]);
            const T* xb = xs + c;
            Vec8<T>::load(xb + ((size_t)y0 * W + x0) * C, v[u][0]);
            Vec8<T>::load(xb + ((size_t)y0 * W + x1) * C, v[u][1]);
            Vec8<T>::load(xb + ((size_t)y1 * W + x0) * C, v[u][2]);
            Vec8<T>::load(xb + ((size_t)y1 * W + x1) * C, v[u][3]);
        }
    }
#pragma unroll
    for (int u = 0; u < CS_U; ++u) {
        const unsigned j = j0 + u * 256u;
        if (j < n8s) {
            const float h1 = ly[u], h0 = 1.f - h1, w1 = lx[u], w0 = 1.f - w1;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = h0 * (w0 * v[u][0][e] + w1 * v[u][1][e]) + h1 * (w0 * v[u][2][e] + w1 * v[u][3][e]);
            Vec8<T>::store(os + (size_t)j * 8, o);
        }
    }
}
// weights with which input index i receives the gradient of output indices 2i-1 .. 2i+2 (the clamps fold the missing
// neighbours of the border pixels into the centre taps)
__device__ __forceinline__ void up2_adj(int i, int n_in, float (&w)[4]) {
    w[0] = i >= 1 ? 0.25f : 0.f;
    w[1] = i >= 1 ? 0.75f : 1.f;
    w[2] = i <= n_in - 2 ? 0.75f : 1.f;
    w[3] = i <= n_in - 2 ? 0.25f : 0.f;
}
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const T* __restrict__ g, unsigned n8s /* H * W * C / 8 */, int H, int W, int C, T* __restrict__ dx) {
    const unsigned c8 = (unsigned)C / 8u;
    const int OW = 2 * W;
    const unsigned j = blockIdx.x * 256u + threadIdx.x;                   // grid.y = sample
    if (j >= n8s) return;
    const unsigned pix = j / c8;
    const int c = (int)(j - pix * c8) * 8;
    const int iy = (int)(pix / (unsigned)W), ix = (int)(pix - (unsigned)iy * (unsigned)W);
    float wy[4], wx[4];
    up2_adj(iy, H, wy);
    up2_adj(ix, W, wx);
    const T* gb = g + (size_t)blockIdx.y * 4 * H * W * C + c;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    const int OH = 2 * H;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        // the 4 loads of a row carry no branch (clamped coordinates; taps with zero weight are zeroed in the conversion)
        Raw8<T> raw[4];
        const int oy = min(max(2 * iy - 1 + a, 0), OH - 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) raw[q].load(gb + ((size_t)oy * OW + min(max(2 * ix - 1 + q, 0), OW - 1)) * C);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool on = wy[a] != 0.f && wx[q] != 0.f;
            float v[8];
            raw[q].get(on, v);
            if (on) {
                const float wgt = wy[a] * wx[q];
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, v[e], acc[e]);
            }
        }
    }
    Vec8<T>::store(dx + ((size_t)blockIdx.y * n8s + j) * 8, acc);
}

inline int grid_cap(long long blocks) {
    const long long cap = (long long)num_sms() * 16;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

// ---- launchers ------------------------------------------------------------------------------------------------
// A/B switches for tools/ln_probe.py (read once): MOEGAN_LN_BULK=0 selects the register-pipelined kernels
static bool ln_bulk_env_off() { static const bool off = [] { const char* e = getenv("MOEGAN_LN_BULK"); return e != nullptr && e[0] == '0'; }(); return off; }
static bool ln_fwd_bulk_disabled() { return ln_bulk_env_off(); }
int ln_bwd_blocks(long long rows) {
    long long b = (rows + 7) / 8;
    const long long cap = (long long)num_sms() * 4;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// (NSEG, R) by segment count: 1 -> (1, 4), 2 -> (2, 2), 3..4 -> (4, 1), 5..8 -> (8, 1)
int launch_layernorm_fwd(const void* x, int x_bf16, const float* gamma, const float* beta, long long rows, int d, float eps, void* y,
                         int y_bf16, void* y_op, float* mean, float* rstd, cudaStream_t st, const void* addend, void* sum_out) {
    ProfScope _ps(st, MOEGAN_PROF_LAYERNORM_FWD, rows, d, 0, (x_bf16 ? 1 : 0) | (y_bf16 ? 2 : 0) | (y_op != nullptr ? 4 : 0) | (addend != nullptr ? 8 : 0));
    if (rows == 0) return MOEGAN_OK;
    typedef __nv_bfloat16 bf;
    const int nseg = d / 128;
    const bool bulk_ok = nseg == 1 || nseg == 2 || nseg == 4;
    MOEGAN_REQUIRE(addend == nullptr || bulk_ok, MOEGAN_ERR_BAD_SHAPE, "add_layernorm_fwd: d = %d must be 128, 256 or 512", d);
    if (bulk_ok && (addend != nullptr || (rows >= 4096 && !ln_fwd_bulk_disabled()))) {
        const uint32_t smem_mul = addend != nullptr ? 2u : 1u;
#define MOEGAN_LN_FBULK3(TI, TO, NS)                                                                                            \
        do {                                                                                                                    \
            typedef LnFwdBulk<TI, NS> LB;                                                                                       \
            static bool attr = false;                                                                                           \
            if (!attr) {                                                                                                        \
                MOEGAN_CHECK_CUDA(cudaFuncSetAttribute(layernorm_fwd_bulk_kernel<TI, TO, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * LNB_STAGES * LB::SB + 2 * LNB_STAGES * 8))); \
                attr = true;                                                                                                    \
            }                                                                                                                   \
            const long long nt = (rows + LB::TR - 1) / LB::TR;                                                                  \
            const long long cap = (long long)num_sms() * 3;                                                                     \
            layernorm_fwd_bulk_kernel<TI, TO, NS><<<(int)(nt < cap ? nt : cap), 256, smem_mul * LNB_STAGES * LB::SB + 2 * LNB_STAGES * 8, st>>>(  \
                reinterpret_cast<const TI*>(x), reinterpret_cast<const TI*>(addend), reinterpret_cast<TI*>(sum_out), gamma, beta, rows, eps,   \
                reinterpret_cast<TO*>(y), reinterpret_cast<__nv_bfloat16*>(y_op), mean, rstd);                                         \
        } while (0)
#define MOEGAN_LN_FBULK(TI, TO) do { if (nseg == 1) MOEGAN_LN_FBULK3(TI, TO, 1); else if (nseg == 2) MOEGAN_LN_FBULK3(TI, TO, 2); else MOEGAN_LN_FBULK3(TI, TO, 4); } while (0)
        if (x_bf16) { if (y_bf16) MOEGAN_LN_FBULK(bf, bf); else MOEGAN_LN_FBULK(bf, float); }
        else        { if (y_bf16) MOEGAN_LN_FBULK(float, bf); else MOEGAN_LN_FBULK(float, float); }
#undef MOEGAN_LN_FBULK
#undef MOEGAN_LN_FBULK3
        MOEGAN_CHECK_LAUNCH("layernorm_fwd_bulk_kernel");
        return MOEGAN_OK;
    }
#define MOEGAN_LN_FWD2(TI, TO, NS, RR)                                                                                         \
    layernorm_fwd_kernel<TI, TO, NS, RR><<<(int)((rows + 8 * RR - 1) / (8 * RR)), 256, 0, st>>>(                              \
        reinterpret_cast<const TI*>(x), gamma, beta, rows, d, eps, reinterpret_cast<TO*>(y), reinterpret_cast<__nv_bfloat16*>(y_op), mean, rstd)
#define MOEGAN_LN_FWD(TI, TO)                                                                                                  \
    do {                                                                                                                       \
        if (nseg == 1) MOEGAN_LN_FWD2(TI, TO, 1, 4); else if (nseg == 2) MOEGAN_LN_FWD2(TI, TO, 2, 2);                         \
        else if (nseg <= 4) MOEGAN_LN_FWD2(TI, TO, 4, 1); else MOEGAN_LN_FWD2(TI, TO, 8, 1);                                   \
    } while (0)
    if (x_bf16) { if (y_bf16) MOEGAN_LN_FWD(bf, bf); else MOEGAN_LN_FWD(bf, float); }
    else        { if (y_bf16) MOEGAN_LN_FWD(float, bf); else MOEGAN_LN_FWD(float, float); }
#undef MOEGAN_LN_FWD
#undef MOEGAN_LN_FWD2
    MOEGAN_CHECK_LAUNCH("layernorm_fwd_kernel");
    return MOEGAN_OK;
}

int launch_layernorm_bwd(const void* g, int g_bf16, const void* x, int x_bf16, const float* gamma, const float* mean, const float* rstd,
                         long long rows, int d, void* dx, float* dgamma_dbeta, float* scratch, cudaStream_t st, const void* gadd) {
    ProfScope _ps(st, MOEGAN_PROF_LAYERNORM_BWD, rows, d, 0, (g_bf16 ? 1 : 0) | (x_bf16 ? 2 : 0) | (gadd != nullptr ? 4 : 0));
    if (rows == 0) {
        MOEGAN_CHECK_CUDA(cudaMemsetAsync(dgamma_dbeta, 0, (size_t)2 * d * 4, st));
        return MOEGAN_OK;
    }
    const int blocks = ln_bwd_blocks(rows);
    typedef __nv_bfloat16 bf;
    const int nseg = d / 128;
    MOEGAN_REQUIRE(gadd == nullptr || nseg == 1 || nseg == 2 || nseg == 4, MOEGAN_ERR_BAD_SHAPE, "layernorm_bwd_add: d = %d must be 128, 256 or 512", d);
    if ((nseg == 1 || nseg == 2 || nseg == 4) && (gadd != nullptr || !ln_bulk_env_off())) {
        // bulk-copy pipelined kernel; same partial layout ([nb][2][d], nb <= blocks) and finishing pass as the register-pipelined one below
        int nb = blocks;
#define MOEGAN_LN_BULK3(TG, TI, NS)                                                                                             \
        do {                                                                                                                    \
            typedef LnBulk<TG, TI, NS> LB;                                                                                      \
            static bool attr = false;                                                                                           \
            if (!attr) {                                                                                                        \
                MOEGAN_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_bulk_kernel<TG, TI, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(LNB_STAGES * (LB::SB + LB::XB) + 2 * LNB_STAGES * 8))); \
                attr = true;                                                                                                    \
            }                                                                                                                   \
            const long long nt = (rows + LB::TR - 1) / LB::TR;                                                                  \
            const uint32_t smem = LNB_STAGES * (LB::SB + (gadd != nullptr ? LB::XB : 0u)) + 2 * LNB_STAGES * 8;                   \
            const long long cap = (long long)num_sms() * (smem > 100 * 1024 ? 1 : 2);        /* persistent: CTAs resident per SM */ \
            nb = (int)(nt < cap ? nt : cap);                                                                                    \
            layernorm_bwd_bulk_kernel<TG, TI, NS><<<nb, 256, smem, st>>>(reinterpret_cast<const TG*>(g), reinterpret_cast<const TI*>(x), gamma, \
                                                                         mean, rstd, rows, reinterpret_cast<const TI*>(gadd), reinterpret_cast<TI*>(dx), scratch); \
        } while (0)
#define MOEGAN_LN_BULK(TG, TI) do { if (nseg == 1) MOEGAN_LN_BULK3(TG, TI, 1); else if (nseg == 2) MOEGAN_LN_BULK3(TG, TI, 2); else MOEGAN_LN_BULK3(TG, TI, 4); } while (0)
        if (g_bf16) { if (x_bf16) MOEGAN_LN_BULK(bf, bf); else MOEGAN_LN_BULK(bf, float); }
        else        { if (x_bf16) MOEGAN_LN_BULK(float, bf); else MOEGAN_LN_BULK(float, float); }
#undef MOEGAN_LN_BULK
#undef MOEGAN_LN_BULK3
        MOEGAN_CHECK_LAUNCH("layernorm_bwd_bulk_kernel");
        partial_sum_tall_kernel<<<ceil_div(2 * d, 32), 256, 0, st>>>(scratch, 2 * d, nb, dgamma_dbeta);
        MOEGAN_CHECK_LAUNCH("partial_sum_tall_kernel");
        return MOEGAN_OK;
    }
#define MOEGAN_LN_BWD2(TG, TI, NS, RR)                                                                                         \
    layernorm_bwd_kernel<TG, TI, NS, RR><<<blocks, 256, 0, st>>>(reinterpret_cast<const TG*>(g), reinterpret_cast<const TI*>(x), gamma, mean, \
                                                                 rstd, rows, d, reinterpret_cast<TI*>(dx), scratch)
#define MOEGAN_LN_BWD(TG, TI)                                                                                                  \
    do {                                                                                                                       \
        if (nseg == 1) MOEGAN_LN_BWD2(TG, TI, 1, 4); else if (nseg == 2) MOEGAN_LN_BWD2(TG, TI, 2, 2);                         \
        else if (nseg <= 4) MOEGAN_LN_BWD2(TG, TI, 4, 1); else MOEGAN_LN_BWD2(TG, TI, 8, 1);                                   \
    } while (0)
    if (g_bf16) { if (x_bf16) MOEGAN_LN_BWD(bf, bf); else MOEGAN_LN_BWD(bf, float); }
    else        { if (x_bf16) MOEGAN_LN_BWD(float, bf); else MOEGAN_LN_BWD(float, float); }
#undef MOEGAN_LN_BWD
#undef MOEGAN_LN_BWD2
    MOEGAN_CHECK_LAUNCH("layernorm_bwd_kernel");
    // scratch = [blocks][2][d] = [blocks][2d]  ->  [dgamma | dbeta]
    partial_sum_tall_kernel<<<ceil_div(2 * d, 32), 256, 0, st>>>(scratch, 2 * d, blocks, dgamma_dbeta);
    MOEGAN_CHECK_LAUNCH("partial_sum_tall_kernel");
    return MOEGAN_OK;
}

int launch_channel_scale_fwd(const void* x, int bf16, const void* s, int s_bf16, int B, int P, int C, float slope, void* out, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_CHANNEL_SCALE_FWD, (long long)B, P, C, (bf16 ? 1 : 0) | (s_bf16 ? 2 : 0));
    const long long n8s = (long long)P * C / 8;
    if (n8s == 0 || B == 0) return MOEGAN_OK;
    MOEGAN_REQUIRE(n8s < (1ll << 31) && B <= 65535, MOEGAN_ERR_BAD_SHAPE, "channel_scale_fwd: P * C / 8 = %lld must be < 2^31 and B = %d <= 65535", n8s, B);
    const dim3 grid((unsigned)((n8s + 256 * CS_U - 1) / (256 * CS_U)), (unsigned)B);
    typedef __nv_bfloat16 bf;
#define MOEGAN_CS_FWD(T, TS) channel_scale_fwd_kernel<T, TS><<<grid, 256, 0, st>>>(reinterpret_cast<const T*>(x), reinterpret_cast<const TS*>(s), (unsigned)n8s, C, slope, reinterpret_cast<T*>(out))
    if (bf16) { if (s_bf16) MOEGAN_CS_FWD(bf, bf); else MOEGAN_CS_FWD(bf, float); }
    else      { if (s_bf16) MOEGAN_CS_FWD(float, bf); else MOEGAN_CS_FWD(float, float); }
#undef MOEGAN_CS_FWD
    MOEGAN_CHECK_LAUNCH("channel_scale_fwd_kernel");
    return MOEGAN_OK;
}

int channel_scale_psplit(int B, int P, int C) {
    // enough CTAs to fill the chip: B * C/128 * psplit >= ~4 per SM, at least 16 pixels (one per half-warp) per split
    long long ctas = (long long)B * (C / 128);
    int ps = 1;
    while (ctas * ps < (long long)num_sms() * 4 && ps * 2 * 16 <= P && ps < 64) ps *= 2;
    return ps;
}

int launch_channel_scale_bwd(const void* g, const void* x, int bf16, const void* s, int s_bf16, int B, int P, int C, float slope, void* dx,
                             void* ds, float* scratch, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_CHANNEL_SCALE_BWD, (long long)B, P, C, (bf16 ? 1 : 0) | (s_bf16 ? 2 : 0));
    if ((long long)B * P * C == 0) return MOEGAN_OK;
    const int ps = channel_scale_psplit(B, P, C);
    dim3 grid(C / 128, B, ps);
    typedef __nv_bfloat16 bf;
#define MOEGAN_CS_BWD(T, TS) channel_scale_bwd_kernel<T, TS><<<grid, 256, 0, st>>>(reinterpret_cast<const T*>(g), reinterpret_cast<const T*>(x), reinterpret_cast<const TS*>(s), P, C, slope, ps, reinterpret_cast<T*>(dx), scratch, B)
    if (bf16) { if (s_bf16) MOEGAN_CS_BWD(bf, bf); else MOEGAN_CS_BWD(bf, float); }
    else      { if (s_bf16) MOEGAN_CS_BWD(float, bf); else MOEGAN_CS_BWD(float, float); }
#undef MOEGAN_CS_BWD
    MOEGAN_CHECK_LAUNCH("channel_scale_bwd_kernel");
    // ds (in s's dtype) = fixed-order sum of the pixel-split partials
    if (s_bf16) partial_sum_kernel<bf><<<ceil_div(B * C, 256), 256, 0, st>>>(scratch, B * C, ps, reinterpret_cast<bf*>(ds));
    else        partial_sum_kernel<float><<<ceil_div(B * C, 256), 256, 0, st>>>(scratch, B * C, ps, reinterpret_cast<float*>(ds));
    MOEGAN_CHECK_LAUNCH("partial_sum_kernel");
    return MOEGAN_OK;
}

int colsum_splits(long long rows, int cols) {
    const long long col_ctas = (cols + 127) / 128;
    long long ns = ((long long)num_sms() * 4 + col_ctas - 1) / col_ctas;     // ~4 CTAs per SM in total
    const long long max_ns = rows / 64 > 1 ? rows / 64 : 1;                  // >= 64 rows per split
    if (ns > max_ns) ns = max_ns;
    return (int)(ns < 1 ? 1 : ns);
}
int launch_colsum(const void* x, int bf16, long long rows, int cols, float* out, float* scratch, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_COLSUM, rows, cols, 0, bf16 ? 1 : 0);
    if (cols == 0) return MOEGAN_OK;
    if (rows == 0) { MOEGAN_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)cols * 4, st)); return MOEGAN_OK; }
    const int ns = colsum_splits(rows, cols);
    dim3 grid((cols + 127) / 128, ns);
    float* part = ns == 1 ? out : scratch;
    if (bf16) colsum_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), rows, cols, ns, part);
    else      colsum_rows_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), rows, cols, ns, part);
    MOEGAN_CHECK_LAUNCH("colsum_rows_kernel");
    if (ns > 16) {
        // hundreds of partials over a few hundred columns: one thread per column walking them serially was most of this launcher's
        // time (28 us for a 17 MB input whose read takes 3 us); 8 row lanes per column, combined in a fixed order
        partial_sum_tall_kernel<<<ceil_div(cols, 32), 256, 0, st>>>(scratch, cols, ns, out);
        MOEGAN_CHECK_LAUNCH("partial_sum_tall_kernel");
    } else if (ns > 1) {
        partial_sum_kernel<float><<<ceil_div(cols, 256), 256, 0, st>>>(scratch, cols, ns, out);
        MOEGAN_CHECK_LAUNCH("partial_sum_kernel");
    }
    return MOEGAN_OK;
}

int launch_upsample2x_fwd(const void* x, int bf16, int B, int H, int W, int C, void* out, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_UPSAMPLE2X_FWD, (long long)B, H * W, C, bf16 ? 1 : 0);
    const long long n8s = (long long)4 * H * W * C / 8;
    if (n8s == 0 || B == 0) return MOEGAN_OK;
    MOEGAN_REQUIRE(n8s < (1ll << 31) && B <= 65535, MOEGAN_ERR_BAD_SHAPE, "upsample2x_fwd: 4 * H * W * C / 8 = %lld must be < 2^31 and B = %d <= 65535", n8s, B);
    const dim3 grid((unsigned)((n8s + 256 * CS_U - 1) / (256 * CS_U)), (unsigned)B);
    if (bf16) upsample2x_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), (unsigned)n8s, H, W, C, reinterpret_cast<__nv_bfloat16*>(out));
    else      upsample2x_fwd_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), (unsigned)n8s, H, W, C, reinterpret_cast<float*>(out));
    MOEGAN_CHECK_LAUNCH("upsample2x_fwd_kernel");
    return MOEGAN_OK;
}
int launch_upsample2x_bwd(const void* g, int bf16, int B, int H, int W, int C, void* dx, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_UPSAMPLE2X_BWD, (long long)B, H * W, C, bf16 ? 1 : 0);
    const long long n8s = (long long)H * W * C / 8;
    if (n8s == 0 || B == 0) return MOEGAN_OK;
    MOEGAN_REQUIRE(n8s < (1ll << 31) && B <= 65535, MOEGAN_ERR_BAD_SHAPE, "upsample2x_bwd: H * W * C / 8 = %lld must be < 2^31 and B = %d <= 65535", n8s, B);
    const dim3 grid((unsigned)((n8s + 255) / 256), (unsigned)B);
    if (bf16) upsample2x_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g), (unsigned)n8s, H, W, C, reinterpret_cast<__nv_bfloat16*>(dx));
    else      upsample2x_bwd_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(g), (unsigned)n8s, H, W, C, reinterpret_cast<float*>(dx));
    MOEGAN_CHECK_LAUNCH("upsample2x_bwd_kernel");
    return MOEGAN_OK;
}

int launch_offset_sample_fwd(const void* x, int bf16, const float* off, int B, int H, int W, int C, void* out, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_OFFSET_SAMPLE_FWD, (long long)B, H * W, C, bf16 ? 1 : 0);
    if ((long long)B * H * W == 0) return MOEGAN_OK;
    MOEGAN_REQUIRE(B <= 65535, MOEGAN_ERR_BAD_SHAPE, "offset_sample: B = %d must be <= 65535", B);
    const dim3 grid((unsigned)((H * W + OS_TP - 1) / OS_TP), (unsigned)B);
    static const bool u2 = [] { const char* e = getenv("MOEGAN_OS_FWD_U"); return e != nullptr && e[0] == '2'; }();     // A/B switch (tools/glue_probe.py)
    if (bf16) {
        if (u2) offset_sample_fwd_kernel<__nv_bfloat16, 2><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), off, H, W, C, reinterpret_cast<__nv_bfloat16*>(out));
        else    offset_sample_fwd_kernel<__nv_bfloat16, 1><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), off, H, W, C, reinterpret_cast<__nv_bfloat16*>(out));
    } else {
        if (u2) offset_sample_fwd_kernel<float, 2><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), off, H, W, C, reinterpret_cast<float*>(out));
        else    offset_sample_fwd_kernel<float, 1><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), off, H, W, C, reinterpret_cast<float*>(out));
    }
    MOEGAN_CHECK_LAUNCH("offset_sample_fwd_kernel");
    return MOEGAN_OK;
}

template <typename T>
static int launch_bwd_in(const T* g, const float* off, int B, int H, int W, int C, T* dx, cudaStream_t st) {
    const int HW = H * W;
    const int S = HW >= 128 ? 4 : (HW >= 32 ? 2 : 1);       // CTAs per sample; slices of QS <= 64 input pixels
    const int QS = (HW + S - 1) / S;
    const size_t smem = (size_t)4 * HW * 16;
    offset_sample_bwd_in_kernel<T><<<B * S, 256, smem, st>>>(g, off, H, W, C, S, QS, dx);
    MOEGAN_CHECK_LAUNCH("offset_sample_bwd_in_kernel");
    return MOEGAN_OK;
}

int launch_offset_sample_bwd(const void* g, const void* x, int bf16, const float* off, int B, int H, int W, int C, void* dx, float* doff,
                             cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_OFFSET_SAMPLE_BWD, (long long)B, H * W, C, bf16 ? 1 : 0);
    if ((long long)B * H * W == 0) return MOEGAN_OK;
    MOEGAN_REQUIRE(B <= 65535 && C <= 1024, MOEGAN_ERR_BAD_SHAPE, "offset_sample_bwd: B = %d must be <= 65535 and C = %d <= 1024", B, C);
    const dim3 grid((unsigned)((H * W + OS_TP - 1) / OS_TP), (unsigned)B);
    if (bf16) {
        offset_sample_bwd_off_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g), reinterpret_cast<const __nv_bfloat16*>(x), off, H, W, C, doff);
        MOEGAN_CHECK_LAUNCH("offset_sample_bwd_off_kernel");
        return launch_bwd_in<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(g), off, B, H, W, C, reinterpret_cast<__nv_bfloat16*>(dx), st);
    }
    offset_sample_bwd_off_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(g), reinterpret_cast<const float*>(x), off, H, W, C, doff);
    MOEGAN_CHECK_LAUNCH("offset_sample_bwd_off_kernel");
    return launch_bwd_in<float>(reinterpret_cast<const float*>(g), off, B, H, W, C, reinterpret_cast<float*>(dx), st);
}

int launch_demod_coef_fwd(const void* sv, int s_bf16, const float* w, long long so, long long si, long long sh, long long sw, int B, int Cin,
                          int Cout, int KH, int KW, float eps, int compute_wsq, float* wsq, float* d, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_DEMOD_COEF_FWD, (long long)B, Cin, Cout, (s_bf16 ? 1 : 0) | (compute_wsq ? 2 : 0) | (KH * KW) << 8);
    if ((long long)Cout * Cin == 0) return MOEGAN_OK;
    if (compute_wsq) {
        demod_wsq_kernel<<<ceil_div(Cout * Cin, 256), 256, 0, st>>>(w, so, si, sh, sw, Cout, Cin, KH, KW, wsq);
        MOEGAN_CHECK_LAUNCH("demod_wsq_kernel");
    }
    if (B == 0) return MOEGAN_OK;
    DemodArgs a;
    memset(&a, 0, sizeof(a));
    a.s = sv; a.s_bf16 = s_bf16; a.w = w; a.so = so; a.si = si; a.sh = sh; a.sw = sw; a.wsq = wsq; a.d_out = d;
    a.B = B; a.Cin = Cin; a.Cout = Cout; a.KH = KH; a.KW = KW; a.eps = eps;
    demod_gemm_kernel<0><<<dim3(ceil_div(Cout, 32), ceil_div(B, 32)), 256, 0, st>>>(a);
    MOEGAN_CHECK_LAUNCH("demod_gemm_kernel<fwd>");
    return MOEGAN_OK;
}
int launch_demod_coef_bwd(const float* dd, const float* d, const void* sv, int s_bf16, const float* w, long long so, long long si, long long sh,
                          long long sw, const float* wsq, int B, int Cin, int Cout, int KH, int KW, void* ds, float* dw, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_DEMOD_COEF_BWD, (long long)B, Cin, Cout, (s_bf16 ? 1 : 0) | (KH * KW) << 8);
    if ((long long)Cout * Cin == 0) return MOEGAN_OK;
    DemodArgs a;
    memset(&a, 0, sizeof(a));
    a.s = sv; a.s_bf16 = s_bf16; a.w = w; a.so = so; a.si = si; a.sh = sh; a.sw = sw; a.wsq = wsq; a.d = d; a.dd = dd; a.ds = ds; a.dw = dw;
    a.B = B; a.Cin = Cin; a.Cout = Cout; a.KH = KH; a.KW = KW;
    if (B > 0) {
        demod_gemm_kernel<1><<<dim3(ceil_div(Cin, 32), ceil_div(B, 32)), 256, 0, st>>>(a);
        MOEGAN_CHECK_LAUNCH("demod_gemm_kernel<ds>");
    }
    demod_gemm_kernel<2><<<dim3(ceil_div(Cin, 32), ceil_div(Cout, 32)), 256, 0, st>>>(a);     // B == 0: K loop empty, dw = 0
    MOEGAN_CHECK_LAUNCH("demod_gemm_kernel<dw>");
    return MOEGAN_OK;
}

}  // namespace moegan
