// Token permutation / expert dispatch, weighted combine and their backward kernels.  All of these are
// HBM-bandwidth-bound byte movers: one warp per token row, 8/16-byte vector accesses, fully coalesced on
// the contiguous side of every gather / scatter.
//
// Reference semantics: SparseMoE.forward expert loops (moegan/t2i_moe_gan.py:462-483).
#include "kernels.cuh"

namespace moegan {

namespace {


// =====================================================================================================
// plan: histogram totals -> padded expert offsets, block->expert map, per-chunk bases, balance load
// =====================================================================================================
__global__ void __launch_bounds__(1024)
plan_kernel(const int* __restrict__ chunk_counts, const float* __restrict__ chunk_colsum, int nchunks, int T, int E, int dense,
            int m_cap, int* __restrict__ plan, int* __restrict__ chunk_base, float* __restrict__ load) {
    __shared__ int tot_e[MOEGAN_MAX_EXPERTS];
    __shared__ int off_e[MOEGAN_MAX_EXPERTS + 1];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;   // warp w owns expert w (E <= 32 warps)

    if (wid < E) {
        // ---- balance load: column sum of the soft probs over chunks, fixed order -> deterministic ----
        const int e = wid;
        float acc = 0.0f;
        for (int c = lane; c < nchunks; c += 32) acc += chunk_colsum[(size_t)c * E + e];
        acc = warp_sum(acc);
        if (lane == 0) load[e] = acc;
        if (!dense) {
            // ---- exclusive scan of this expert's chunk histogram: lane owns a contiguous run of chunks ----
            const int per = (nchunks + 31) / 32;
            const int c0 = lane * per, c1 = min(c0 + per, nchunks);
            int local = 0;
            for (int c = c0; c < c1; ++c) local += chunk_counts[(size_t)c * E + e];
            int incl = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            int run = incl - local;
            for (int c = c0; c < c1; ++c) { chunk_base[(size_t)c * E + e] = run; run += chunk_counts[(size_t)c * E + e]; }
            if (lane == 31) tot_e[e] = incl;
        }
    }
    __syncthreads();

    if (dense) {
        const int T_pad = (T + MOEGAN_ROW_ALIGN - 1) / MOEGAN_ROW_ALIGN * MOEGAN_ROW_ALIGN;
        const int bpe = T_pad / 128;
        if (tid <= E) plan[PLAN_OFF + tid] = tid * T_pad;
        if (tid < E) plan[PLAN_CNT + tid] = T;
        for (int i = tid; i < E * bpe; i += 1024) plan[PLAN_BLK + i] = i / bpe;
        if (tid == 0) { plan[0] = E * bpe; plan[1] = E * T_pad; plan[2] = T_pad; }
        return;
    }
    if (tid == 0) {
        int off = 0;
        for (int e = 0; e < E; ++e) {
            off_e[e] = off;
            plan[PLAN_OFF + e] = off;
            plan[PLAN_CNT + e] = tot_e[e];
            off += (tot_e[e] + MOEGAN_ROW_ALIGN - 1) / MOEGAN_ROW_ALIGN * MOEGAN_ROW_ALIGN;
        }
        if (off > m_cap) off = m_cap;            // cannot happen (m_cap is an upper bound); guards the buffers
        off_e[E] = off;
        plan[PLAN_OFF + E] = off;
        plan[0] = off / 128; plan[1] = off; plan[2] = 0;
    }
    __syncthreads();
    const int nblk = off_e[E] / 128;
    for (int i = tid; i < nblk; i += 1024) {
        int e = 0;
        while (e + 1 < E && i * 128 >= off_e[e + 1]) ++e;
        plan[PLAN_BLK + i] = e;
    }
}

// =====================================================================================================
// gather: stable permutation of token rows into expert-contiguous order, fp32 -> bf16 operand
// =====================================================================================================
__device__ __forceinline__ void store_operand4(__nv_bfloat16* dst, int lo_off, bool split, float4 v) {
    float h0, h1, h2, h3, l0, l1, l2, l3;
    split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
    if (split) *reinterpret_cast<uint2*>(dst + lo_off) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
}

__global__ void __launch_bounds__(256)
gather_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ x_bf16, const int* __restrict__ topk_idx, const float* __restrict__ gates,
              const int* __restrict__ plan, const int* __restrict__ chunk_base, int nchunks, int T, int d, int E, int k, int split,
              __nv_bfloat16* __restrict__ x_op, __nv_bfloat16* __restrict__ x_copy,
              int* __restrict__ row_token, float* __restrict__ row_gate, int* __restrict__ token_rows) {
    __shared__ int dst_row[MOEGAN_CHUNK * MOEGAN_MAX_TOPK];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ld = split ? 2 * d : d;
    if ((int)blockIdx.x >= nchunks) {
        // ---- padding rows of expert e: zero operand rows, row_token = -1 ----
        const int e = blockIdx.x - nchunks;
        const int r0 = plan[PLAN_OFF + e] + plan[PLAN_CNT + e], r1 = plan[PLAN_OFF + e + 1];
        for (int r = r0 + wid; r < r1; r += 8) {
            uint2* row = reinterpret_cast<uint2*>(x_op + (size_t)r * ld);
            for (int i = lane; i < ld / 4; i += 32) row[i] = make_uint2(0u, 0u);
            if (lane == 0) { row_token[r] = -1; row_gate[r] = 0.0f; }
        }
        return;
    }
    const int chunk = blockIdx.x;
    if (wid == 0) {
        const int t = chunk * MOEGAN_CHUNK + lane;
        const bool valid = t < T;
        int idx[MOEGAN_MAX_TOPK], rows[MOEGAN_MAX_TOPK];
#pragma unroll
        for (int j = 0; j < MOEGAN_MAX_TOPK; ++j) { idx[j] = (valid && j < k) ? topk_idx[(size_t)t * k + j] : -1; rows[j] = -1; }
        // lane e holds expert e's first row for this chunk: ONE coalesced load instead of E dependent ones
        const int my_base = lane < E ? plan[PLAN_OFF + lane] + chunk_base[(size_t)chunk * E + lane] : 0;
        for (int e = 0; e < E; ++e) {
            bool has = false;
#pragma unroll
            for (int j = 0; j < MOEGAN_MAX_TOPK; ++j) has |= (idx[j] == e);
            const unsigned b = __ballot_sync(0xffffffffu, has);
            const int base_e = __shfl_sync(0xffffffffu, my_base, e);
            if (has) {
                const int row = base_e + __popc(b & ((1u << lane) - 1u));   // tokens before me in this chunk (stable)
#pragma unroll
                for (int j = 0; j < MOEGAN_MAX_TOPK; ++j) if (idx[j] == e) rows[j] = row;
            }
        }
#pragma unroll
        for (int j = 0; j < MOEGAN_MAX_TOPK; ++j) {
            if (j < k) {
                dst_row[lane * k + j] = rows[j];
                if (valid) {
                    token_rows[(size_t)t * k + j] = rows[j];
                    row_token[rows[j]] = t;
                    row_gate[rows[j]] = gates[(size_t)t * k + j];
                }
            }
        }
    }
    __syncthreads();
    // Row copy, flat over (token, 8-column group) so that all 256 threads carry 16-byte accesses whatever d is (a warp per token
    // with 8-byte lanes left half of every request slot empty: 0.20-0.29 of the HBM roof).  The dispatched row indices come from
    // shared memory (staged above), the source row is read once and written to its k destinations.
    const int g8 = d / 8;
    const int items = MOEGAN_CHUNK * g8;
    for (int it = tid; it < items; it += 256) {
        const int tok = it / g8, col = (it - tok * g8) * 8;
        const int t = chunk * MOEGAN_CHUNK + tok;
        if (t >= T) break;
        if (x_bf16 != nullptr) {              // the caller's bf16 copy of x (norm3 emitted it): a pure 16-byte row copy, half the read bytes
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(x_bf16 + (size_t)t * d + col));
            for (int j = 0; j < k; ++j)
                *reinterpret_cast<uint4*>(x_op + (size_t)dst_row[tok * k + j] * d + col) = u;
            continue;
        }
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + (size_t)t * d + col));
        const float4 b = __ldg(reinterpret_cast<const float4*>(x + (size_t)t * d + col + 4));
        if (!split) {
            const uint4 u = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
            for (int j = 0; j < k; ++j)
                *reinterpret_cast<uint4*>(x_op + (size_t)dst_row[tok * k + j] * d + col) = u;
            if (x_copy != nullptr) *reinterpret_cast<uint4*>(x_copy + (size_t)t * d + col) = u;
        } else {
            for (int j = 0; j < k; ++j) {
                store_operand4(x_op + (size_t)dst_row[tok * k + j] * ld + col, d, true, a);
                store_operand4(x_op + (size_t)dst_row[tok * k + j] * ld + col + 4, d, true, b);
            }
            if (x_copy != nullptr) {
                store_operand4(x_copy + (size_t)t * ld + col, d, true, a);
                store_operand4(x_copy + (size_t)t * ld + col + 4, d, true, b);
            }
        }
    }
}

// fp32 [rows, cols] -> bf16 operand [rows, S*cols]
__global__ void __launch_bounds__(256)
to_operand_kernel(const float* __restrict__ src, long long rows, int cols, int split, int sanitize, __nv_bfloat16* __restrict__ dst) {
    const long long n4 = rows * (cols / 4);
    const int c4 = cols / 4;
    const int ld = split ? 2 * cols : cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c4;
        const int c = (int)(i - r * c4) * 4;
        float4 v = __ldg(reinterpret_cast<const float4*>(src + r * cols + c));
        if (sanitize) { v.x = nan_to_num_router(v.x); v.y = nan_to_num_router(v.y); v.z = nan_to_num_router(v.z); v.w = nan_to_num_router(v.w); }
        store_operand4(dst + r * ld + c, cols, split, v);
    }
}

// per-expert fp32 parameters (state_dict layout) -> stacked operands + stacked biases
struct WeightPtrs {
    const float* w1[MOEGAN_MAX_EXPERTS];
    const float* b1[MOEGAN_MAX_EXPERTS];
    const float* w2[MOEGAN_MAX_EXPERTS];
    const float* b2[MOEGAN_MAX_EXPERTS];
};
// grid (x, 2E): blockIdx.y = (which, expert) selects one [rows, cols] fp32 matrix; it is a plain linear fp32 -> bf16 copy in bf16
// mode (same row-major layout on both sides), so the inner loop has no index arithmetic beyond the grid stride; 4 independent
// 16-byte loads per thread and iteration.  (The first version decoded (which, e, row, col) from a flat index with three 64-bit
// divisions per element and ran at 2.9 TB/s.)
__global__ void __launch_bounds__(256)
prep_weights_kernel(const WeightPtrs ptrs, int E, int d, int split, __nv_bfloat16* __restrict__ W1_op, float* __restrict__ b1s,
                    __nv_bfloat16* __restrict__ W2_op, float* __restrict__ b2s) {
    const int hid = 4 * d;
    const int which = blockIdx.y / E, e = blockIdx.y - which * E;      // 0: W1 [4d, d], 1: W2 [d, 4d]
    const int cols = which == 0 ? d : hid, rows = which == 0 ? hid : d;
    const int per = rows * (cols / 4);                                  // float4 per matrix
    const float4* src = reinterpret_cast<const float4*>(which == 0 ? ptrs.w1[e] : ptrs.w2[e]);
    __nv_bfloat16* dstm = (which == 0 ? W1_op : W2_op) + (size_t)e * rows * (split ? 2 * cols : cols);
    const int stride = gridDim.x * blockDim.x;
    for (int q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < per; q0 += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int q = q0 + u * stride; if (q < per) v[u] = __ldg(src + q); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = q0 + u * stride;
            if (q < per) {
                if (!split) {
                    *reinterpret_cast<uint2*>(dstm + (size_t)q * 4) = make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
                } else {
                    const int c4 = cols / 4;
                    const int row = q / c4, c = (q - row * c4) * 4;
                    store_operand4(dstm + (size_t)row * 2 * cols + c, cols, true, v[u]);
                }
            }
        }
    }
    if (blockIdx.y == 0) {
        const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
        for (int i = gtid; i < E * hid; i += stride) b1s[i] = ptrs.b1[i / hid][i % hid];
        for (int i = gtid; i < E * d; i += stride) b2s[i] = ptrs.b2[i / d][i % d];
    }
}

// =====================================================================================================
// combine forward: out[t] = sum_j gate_j * y[row(t, j)]      (t2i_moe_gan.py:470 / :483)
// =====================================================================================================
template <bool Y_F32>
__device__ __forceinline__ float4 load_row4(const void* base, size_t elem_off) {
    if (Y_F32) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off));
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off));
    return make_float4(bf16_lo_to_float(u.x), bf16_hi_to_float(u.x), bf16_lo_to_float(u.y), bf16_hi_to_float(u.y));
}

// 4 consecutive fp32 / bf16 values -> float4, and back (residual, upstream gradient and output of the combine may be either)
template <bool BF16>
__device__ __forceinline__ float4 load4_any(const void* base, size_t elem_off) { return load_row4<!BF16>(base, elem_off); }
template <bool BF16>
__device__ __forceinline__ void store4_any(void* base, size_t elem_off, float4 v) {
    if (BF16) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    else      *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off) = v;
}

// loads kept in their storage format until use (a pending bf16 load holds 2 registers, not 4)
template <bool F32> struct Raw4 { typedef uint2 type; };
template <> struct Raw4<true> { typedef float4 type; };
template <bool F32>
__device__ __forceinline__ typename Raw4<F32>::type raw_load4(const void* base, size_t elem_off) {
    if constexpr (F32) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off));
    else               return __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off));
}
__device__ __forceinline__ float4 raw_to_float4(float4 v) { return v; }
__device__ __forceinline__ float4 raw_to_float4(uint2 u) {
    return make_float4(bf16_lo_to_float(u.x), bf16_hi_to_float(u.x), bf16_lo_to_float(u.y), bf16_hi_to_float(u.y));
}

// Memory-level parallelism: these kernels only move bytes, so what bounds them is how many loads a warp keeps in flight
// (B200 needs ~30 KB outstanding per SM to saturate HBM).  NSEG (= d / 128) is a template parameter, the slot / row loops are
// processed in batches whose loads are ALL issued before the first use (batch x NSEG = 8 independent 8/16-byte loads per lane),
// and the arithmetic keeps the reference's accumulation order (slot order == expert order), so results do not depend on the batch.
// combine_fwd / uncombine are flat over 8-column chunks (16 B of bf16, 32 B of fp32): thread f owns chunk (t, c) = (f / CPR, f % CPR),
// CPR = d / 8, CHUNKS_PER_THREAD chunks 256 apart, and issues the loads of SLOTS slots of all its chunks before the first use.
// Measured on B200 (ncu, T = 65 536, d = 128): what bounds these kernels is the number of outstanding load REQUESTS per SM, not
// bytes - a warp-per-token version with 8-byte lanes ran at 2.9 TB/s; 16-byte lanes move twice the bytes per request.
constexpr int CHUNKS_PER_THREAD = 2;
template <bool F32> struct Chunk8 { uint4 a; };                       // 8 bf16
template <> struct Chunk8<true> { float4 a, b; };                      // 8 fp32
template <bool F32>
__device__ __forceinline__ Chunk8<F32> chunk_load(const void* base, size_t elem_off) {
    Chunk8<F32> c;
    if constexpr (F32) {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
        c.a = __ldg(p); c.b = __ldg(p + 1);
    } else {
        c.a = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off));
    }
    return c;
}
__device__ __forceinline__ void chunk_unpack(const Chunk8<true>& c, float (&f)[8]) {
    f[0] = c.a.x; f[1] = c.a.y; f[2] = c.a.z; f[3] = c.a.w; f[4] = c.b.x; f[5] = c.b.y; f[6] = c.b.z; f[7] = c.b.w;
}
__device__ __forceinline__ void chunk_unpack(const Chunk8<false>& c, float (&f)[8]) {
    f[0] = bf16_lo_to_float(c.a.x); f[1] = bf16_hi_to_float(c.a.x); f[2] = bf16_lo_to_float(c.a.y); f[3] = bf16_hi_to_float(c.a.y);
    f[4] = bf16_lo_to_float(c.a.z); f[5] = bf16_hi_to_float(c.a.z); f[6] = bf16_lo_to_float(c.a.w); f[7] = bf16_hi_to_float(c.a.w);
}
template <bool BF16>
__device__ __forceinline__ void chunk_store(void* base, size_t elem_off, const float (&f)[8]) {
    if constexpr (BF16) {
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    } else {
        float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off);
        p[0] = make_float4(f[0], f[1], f[2], f[3]); p[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
}

// RES: 0 = no residual, 1 = fp32 residual, 2 = bf16 residual (`x_spatial + moe_out`, t2i_moe_gan.py:571, in the
// combine's epilogue); OUT_BF16: write the result in bf16 (the dtype of the residual stream under autocast)
template <bool Y_F32, int RES, bool OUT_BF16, int NSEG>
__global__ void __launch_bounds__(256)
combine_fwd_kernel(const void* __restrict__ y, const float* __restrict__ gates, const int* __restrict__ token_rows,
                   int T, int T_pad, int d, int k, int dense, const void* __restrict__ residual, void* __restrict__ out) {
    constexpr int CPR = NSEG * 16, U = CHUNKS_PER_THREAD, JB = Y_F32 ? 2 : 4;
    const long long f0 = (long long)blockIdx.x * (256 * U) + threadIdx.x;
    int t[U], col[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long f = f0 + u * 256;
        const int tt = (int)(f / CPR);
        t[u] = tt < T ? tt : -1;
        col[u] = (int)(f - (long long)tt * CPR) * 8;
    }
    Chunk8<RES != 2> res[U];
    if (RES != 0) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (t[u] >= 0) res[u] = chunk_load<RES != 2>(residual, (size_t)t[u] * d + col[u]);
    }
    float acc[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[u][e] = 0.0f;
    for (int j0 = 0; j0 < k; j0 += JB) {              // slot order == expert order in dense mode (:467)
        int row[U][JB]; float g[U][JB];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) {
                const int j = j0 + jj;
                const bool ok = j < k && t[u] >= 0;
                row[u][jj] = ok ? (dense ? j * T_pad + t[u] : __ldg(token_rows + (size_t)t[u] * k + j)) : -1;
                g[u][jj] = ok ? __ldg(gates + (size_t)t[u] * k + j) : 0.0f;
            }
        Chunk8<Y_F32> v[U][JB];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int jj = 0; jj < JB; ++jj)
                if (row[u][jj] >= 0) v[u][jj] = chunk_load<Y_F32>(y, (size_t)row[u][jj] * d + col[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int jj = 0; jj < JB; ++jj)
                if (row[u][jj] >= 0) {
                    float f[8];
                    chunk_unpack(v[u][jj], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[u][e] = fmaf(g[u][jj], f[e], acc[u][e]);
                }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (t[u] >= 0) {
            if (RES != 0) {
                float r[8];
                chunk_unpack(res[u], r);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[u][e] += r[e];
            }
            chunk_store<OUT_BF16>(out, (size_t)t[u] * d + col[u], acc[u]);
        }
}

// combine backward: CTA = 32 consecutive dispatched rows (never straddles experts: segments are 256-row aligned).  Thread = one
// 8-column chunk (c = tid % CPR) of the rows rg, rg + RG, ... (rg = tid / CPR; RG row groups of CPR threads each), 16-byte accesses.
// Also emits the column sums of the CTA's 32 dy rows (colpart [m_cap / 32, d]) - the second Linear's bias gradient is their
// per-expert sum, so dy is never re-read for it.  Every reduction runs in a fixed order (bit-identical reruns).
template <int NSEG> struct CbRowGroups { static constexpr int value = NSEG == 1 ? 16 : (NSEG == 2 ? 8 : (NSEG <= 4 ? 4 : 2)); };
__device__ __forceinline__ void store_operand8(__nv_bfloat16* dst, int lo_off, bool split, const float (&v)[8]) {
    float h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(v[e], h[e], l[e]);
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
    if (split) *reinterpret_cast<uint4*>(dst + lo_off) = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
}

template <bool Y_F32, bool DOUT_BF16, int NSEG>
__global__ void __launch_bounds__(NSEG * 16 * CbRowGroups<NSEG>::value)
combine_bwd_kernel(const int* __restrict__ plan, const void* __restrict__ dout, const void* __restrict__ y,
                   const int* __restrict__ row_token, const float* __restrict__ row_gate, const float* __restrict__ gates,
                   int T, int T_pad, int d, int E, int dense, int split,
                   __nv_bfloat16* __restrict__ dy_op, float* __restrict__ dgate_row, float* __restrict__ colpart) {
    constexpr int CPR = NSEG * 16, RG = CbRowGroups<NSEG>::value, RPT = 32 / RG;     // rows per thread
    constexpr int IB = (Y_F32 || !DOUT_BF16) ? (RPT < 2 ? RPT : 2) : (RPT < 4 ? RPT : 4);   // rows whose loads are in flight together
    __shared__ float red[RG][NSEG * 128];
    __shared__ float dot_part[32][NSEG];
    const int tid = threadIdx.x, c = tid % CPR, rg = tid / CPR, col = c * 8;
    if (blockIdx.x * 32 >= __ldg(plan + 1)) return;    // rows in use (a multiple of 128)
    const int ld = split ? 2 * d : d;
    float cs[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] = 0.0f;
#pragma unroll
    for (int i0 = 0; i0 < RPT; i0 += IB) {
        int tok[IB]; float g[IB];
#pragma unroll
        for (int ii = 0; ii < IB; ++ii) {
            const int r = blockIdx.x * 32 + rg + RG * (i0 + ii);
            int t;
            if (dense) {
                const int e = r / T_pad; t = r - e * T_pad;
                if (t >= T) t = -1;
                g[ii] = t >= 0 ? __ldg(gates + (size_t)t * E + e) : 0.0f;
            } else {
                t = __ldg(row_token + r);
                g[ii] = t >= 0 ? __ldg(row_gate + r) : 0.0f;
            }
            tok[ii] = t;
        }
        Chunk8<!DOUT_BF16> go[IB];
        Chunk8<Y_F32> yv[IB];
#pragma unroll
        for (int ii = 0; ii < IB; ++ii)
            if (tok[ii] >= 0) {
                const int r = blockIdx.x * 32 + rg + RG * (i0 + ii);
                go[ii] = chunk_load<!DOUT_BF16>(dout, (size_t)tok[ii] * d + col);
                yv[ii] = chunk_load<Y_F32>(y, (size_t)r * d + col);
            }
#pragma unroll
        for (int ii = 0; ii < IB; ++ii) {
            const int rl = rg + RG * (i0 + ii);
            const int r = blockIdx.x * 32 + rl;
            __nv_bfloat16* dst = dy_op + (size_t)r * ld + col;
            float dot = 0.0f;
            if (tok[ii] < 0) {                             // padding row: zero (the wgrad GEMMs sum over it)
                *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
                if (split) *reinterpret_cast<uint4*>(dst + d) = make_uint4(0u, 0u, 0u, 0u);
            } else {
                float a[8], b[8], dv[8];
                chunk_unpack(go[ii], a);
                chunk_unpack(yv[ii], b);
#pragma unroll
                for (int e = 7; e >= 0; --e) dot = fmaf(a[e], b[e], dot);
#pragma unroll
                for (int e = 0; e < 8; ++e) { dv[e] = g[ii] * a[e]; cs[e] += dv[e]; }
                store_operand8(dst, d, split != 0, dv);
            }
            // <dout_t, y_r>: 16-lane groups never straddle rows (CPR is a multiple of 16); the NSEG group sums are added in order
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            if ((c & 15) == 0) dot_part[rl][c >> 4] = dot;
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[rg][col + e] = cs[e];
    __syncthreads();
    if (tid < 32) {
        float v = dot_part[tid][0];
#pragma unroll
        for (int q = 1; q < NSEG; ++q) v += dot_part[tid][q];
        dgate_row[blockIdx.x * 32 + tid] = v;
    }
    if (colpart == nullptr) return;
    for (int j = tid; j < NSEG * 128; j += CPR * RG) {
        float a = red[0][j];
#pragma unroll
        for (int w = 1; w < RG; ++w) a += red[w][j];
        colpart[(size_t)blockIdx.x * d + j] = a;
    }
}

// dx[t] = sum_j dx_rows[row(t, j)] (+ dx_router[t])      (same flat chunk layout as combine_fwd)
template <bool R_F32, int NSEG>
__global__ void __launch_bounds__(256)
uncombine_kernel(const void* __restrict__ dx_rows, const int* __restrict__ token_rows, const float* __restrict__ dx_router,
                 int T, int T_pad, int d, int k, int dense, float* __restrict__ dx) {
    constexpr int CPR = NSEG * 16, U = CHUNKS_PER_THREAD, JB = R_F32 ? 2 : 4;
    const long long f0 = (long long)blockIdx.x * (256 * U) + threadIdx.x;
    int t[U], col[U];
    float acc[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long f = f0 + u * 256;
        const int tt = (int)(f / CPR);
        t[u] = tt < T ? tt : -1;
        col[u] = (int)(f - (long long)tt * CPR) * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[u][e] = 0.0f;
        if (dx_router != nullptr && t[u] >= 0) chunk_unpack(chunk_load<true>(dx_router, (size_t)t[u] * d + col[u]), acc[u]);
    }
    for (int j0 = 0; j0 < k; j0 += JB) {
        int row[U][JB];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) {
                const int j = j0 + jj;
                row[u][jj] = (j < k && t[u] >= 0) ? (dense ? j * T_pad + t[u] : __ldg(token_rows + (size_t)t[u] * k + j)) : -1;
            }
        Chunk8<R_F32> v[U][JB];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int jj = 0; jj < JB; ++jj)
                if (row[u][jj] >= 0) v[u][jj] = chunk_load<R_F32>(dx_rows, (size_t)row[u][jj] * d + col[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int jj = 0; jj < JB; ++jj)
                if (row[u][jj] >= 0) {
                    float f[8];
                    chunk_unpack(v[u][jj], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[u][e] += f[e];
                }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (t[u] >= 0) chunk_store<false>(dx, (size_t)t[u] * d + col[u], acc[u]);
}

// bias gradients: out[e, c] = sum over real rows of expert e of (hi + lo)[row, c].
// Stage 1: CTA = (64 columns, expert, row split): 32 warps stride the split's rows, two rows in flight per warp;
//          writes part[split][e][c].  Stage 2 sums the splits in a fixed order (deterministic).
__global__ void __launch_bounds__(1024)
colsum_part_kernel(const int* __restrict__ plan, const __nv_bfloat16* __restrict__ src, int cols, int split, int T, int dense,
                   int E, int nsplit, float* __restrict__ part) {
    __shared__ float red[32][64];
    const int e = blockIdx.y, sp = blockIdx.z;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c = blockIdx.x * 64 + lane * 2;
    const int ld = split ? 2 * cols : cols;
    const int e0 = __ldg(plan + PLAN_OFF + e);
    const int n = dense ? T : __ldg(plan + PLAN_CNT + e);
    const int per = (n + nsplit - 1) / nsplit;
    const int r0 = e0 + min(sp * per, n), r1 = e0 + min((sp + 1) * per, n);
    float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
    auto add_row = [&](int r, float& x0, float& x1) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)r * ld + c));
        x0 += bf16_lo_to_float(u); x1 += bf16_hi_to_float(u);
        if (split) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)r * ld + cols + c));
            x0 += bf16_lo_to_float(v); x1 += bf16_hi_to_float(v);
        }
    };
    int r = r0 + wid;
    for (; r + 32 < r1; r += 64) { add_row(r, a0, a1); add_row(r + 32, b0, b1); }
    if (r < r1) add_row(r, a0, a1);
    red[wid][lane * 2] = a0 + b0;
    red[wid][lane * 2 + 1] = a1 + b1;
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < 32; ++w) s += red[w][threadIdx.x];
        part[((size_t)sp * E + e) * cols + blockIdx.x * 64 + threadIdx.x] = s;
    }
}
// Bias gradients from 32-row slab sums: out[sp][e][c] = sum of part[j][c] over split sp of expert e's slabs j (part =
// [m_cap / 32, cols], written by combine_bwd_kernel or the dgrad epilogue; slabs past the expert's real rows hold zeros
// and are skipped).  Thread = column (coalesced rows), 4 slabs in flight; fixed order -> deterministic.
__global__ void __launch_bounds__(128)
slab_sum_kernel(const int* __restrict__ plan, const float* __restrict__ part, int cols, int T, int dense, int E, int nsplit,
                float* __restrict__ out) {
    const int e = blockIdx.y, sp = blockIdx.z;
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= cols) return;
    const int e0 = __ldg(plan + PLAN_OFF + e);
    const int n = dense ? T : __ldg(plan + PLAN_CNT + e);
    const int j0 = e0 / 32, nj = (n + 31) / 32;
    const int per = (nj + nsplit - 1) / nsplit;
    const int a = j0 + min(sp * per, nj), b = j0 + min((sp + 1) * per, nj);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int j = a;
    for (; j + 3 < b; j += 4) {
        s0 += __ldg(part + (size_t)j * cols + c);       s1 += __ldg(part + (size_t)(j + 1) * cols + c);
        s2 += __ldg(part + (size_t)(j + 2) * cols + c); s3 += __ldg(part + (size_t)(j + 3) * cols + c);
    }
    for (; j < b; ++j) s0 += __ldg(part + (size_t)j * cols + c);
    out[((size_t)sp * E + e) * cols + c] = (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(256)
colsum_finish_kernel(const float* __restrict__ part, int n, int nsplit, float* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float s = 0.0f;
    for (int sp = 0; sp < nsplit; ++sp) s += part[(size_t)sp * n + i];
    out[i] = s;
}

}  // namespace

// =====================================================================================================
// host launchers
// =====================================================================================================
int launch_plan(const moegan_shape* s, const int32_t* chunk_counts, const float* chunk_colsum, int32_t* plan,
                int32_t* chunk_base, float* load, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_PLAN, s);
    plan_kernel<<<1, 1024, 0, st>>>(chunk_counts, chunk_colsum, ceil_div(s->T, MOEGAN_CHUNK), s->T, s->E, s->dense, s->m_cap, plan,
                                    chunk_base, load);
    MOEGAN_CHECK_LAUNCH("plan_kernel");
    return MOEGAN_OK;
}

int launch_gather(const moegan_shape* s, const float* x, const void* x_bf16, const int32_t* topk_idx, const float* gates, const int32_t* plan,
                  const int32_t* chunk_base, void* x_op, void* x_copy, int32_t* row_token, float* row_gate, int32_t* token_rows,
                  cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_GATHER, s, (x_copy != nullptr ? 0x1000 : 0) | (x_bf16 != nullptr ? 0x2000 : 0));
    MOEGAN_REQUIRE(x_bf16 == nullptr || s->precision == 0, MOEGAN_ERR_UNSUPPORTED, "gather: a bf16 source is bf16 mode only");
    const int nchunks = ceil_div(s->T, MOEGAN_CHUNK);
    gather_kernel<<<nchunks + s->E, 256, 0, st>>>(x, reinterpret_cast<const __nv_bfloat16*>(x_bf16), topk_idx, gates, plan, chunk_base, nchunks, s->T, s->d, s->E, s->k,
                                                  s->precision == 1, reinterpret_cast<__nv_bfloat16*>(x_op),
                                                  reinterpret_cast<__nv_bfloat16*>(x_copy), row_token, row_gate, token_rows);
    MOEGAN_CHECK_LAUNCH("gather_kernel");
    return MOEGAN_OK;
}

int launch_to_operand(const float* src, int64_t rows, int64_t cols, int split, int sanitize, void* dst, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_TO_OPERAND, (long long)rows, (int)cols, 0, split ? 0x400 : 0);
    if (rows * cols == 0) return MOEGAN_OK;
    MOEGAN_REQUIRE(cols % 4 == 0, MOEGAN_ERR_BAD_SHAPE, "to_operand: cols (%lld) must be a multiple of 4", (long long)cols);
    const long long n4 = rows * cols / 4;
    long long blocks = ceil_div64(n4, 256);
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    to_operand_kernel<<<(int)blocks, 256, 0, st>>>(src, rows, (int)cols, split, sanitize, reinterpret_cast<__nv_bfloat16*>(dst));
    MOEGAN_CHECK_LAUNCH("to_operand_kernel");
    return MOEGAN_OK;
}

int launch_prep_weights(const moegan_shape* s, const float* const* w1, const float* const* b1, const float* const* w2,
                        const float* const* b2, void* W1_op, float* b1s, void* W2_op, float* b2s, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_PREP_WEIGHTS, s);
    WeightPtrs p;
    for (int e = 0; e < MOEGAN_MAX_EXPERTS; ++e) {
        const int q = e < s->E ? e : 0;
        p.w1[e] = w1[q]; p.b1[e] = b1[q]; p.w2[e] = w2[q]; p.b2[e] = b2[q];
    }
    const long long per = 4LL * s->d * s->d / 4;                        // float4 per matrix
    long long bx = ceil_div64(per, 256 * 4);                            // 4 float4 per thread
    const long long cap = ceil_div64((long long)num_sms() * 8, 2 * s->E);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    prep_weights_kernel<<<dim3((unsigned)bx, (unsigned)(2 * s->E)), 256, 0, st>>>(p, s->E, s->d, s->precision == 1, reinterpret_cast<__nv_bfloat16*>(W1_op), b1s,
                                                     reinterpret_cast<__nv_bfloat16*>(W2_op), b2s);
    MOEGAN_CHECK_LAUNCH("prep_weights_kernel");
    return MOEGAN_OK;
}

// NSEG = d / 128 as a compile-time constant (d is validated as a multiple of 128 in [128, 1024])
#define MOEGAN_NSEG_SWITCH(nseg, CALL) \
    switch (nseg) { case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; case 4: CALL(4); break; \
                    case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; default: CALL(8); break; }

template <bool YF, int RS, bool OB>
static void combine_fwd_dispatch(const moegan_shape* s, const void* y, const float* gates, const int32_t* token_rows, const void* residual,
                                 void* out, cudaStream_t st) {
    const int T_pad = row_pad(s->T);
#define MOEGAN_CF(NS) combine_fwd_kernel<YF, RS, OB, NS><<<(unsigned)ceil_div64((int64_t)s->T * NS * 16, 256 * CHUNKS_PER_THREAD), 256, 0, st>>>(y, gates, token_rows, s->T, T_pad, s->d, s->k, s->dense, residual, out)
    MOEGAN_NSEG_SWITCH(s->d / 128, MOEGAN_CF)
#undef MOEGAN_CF
}

int launch_combine_fwd(const moegan_shape* s, const void* y, const float* gates, const int32_t* token_rows, const void* residual,
                       int residual_bf16, void* out, int out_bf16, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_COMBINE_FWD, s, (residual == nullptr ? 0 : (residual_bf16 ? 0x2000 : 0x1000)) | (out_bf16 ? 0x4000 : 0));
    const int res = residual == nullptr ? 0 : (residual_bf16 ? 2 : 1);
#define MOEGAN_CF_R(YF, OB) do { if (res == 0) combine_fwd_dispatch<YF, 0, OB>(s, y, gates, token_rows, residual, out, st); \
                                 else if (res == 1) combine_fwd_dispatch<YF, 1, OB>(s, y, gates, token_rows, residual, out, st); \
                                 else combine_fwd_dispatch<YF, 2, OB>(s, y, gates, token_rows, residual, out, st); } while (0)
    if (s->precision == 1) { if (out_bf16) MOEGAN_CF_R(true, true); else MOEGAN_CF_R(true, false); }
    else                   { if (out_bf16) MOEGAN_CF_R(false, true); else MOEGAN_CF_R(false, false); }
#undef MOEGAN_CF_R
    MOEGAN_CHECK_LAUNCH("combine_fwd_kernel");
    return MOEGAN_OK;
}

template <bool YF, bool DB>
static void combine_bwd_dispatch(const moegan_shape* s, const int32_t* plan, const void* dout, const void* y, const int32_t* row_token,
                                 const float* row_gate, const float* gates, int split, __nv_bfloat16* dy, float* dgate_row, float* dy_colpart,
                                 cudaStream_t st) {
    const int T_pad = row_pad(s->T);
    const int blocks = s->m_cap / 32;
#define MOEGAN_CB(NS) combine_bwd_kernel<YF, DB, NS><<<blocks, NS * 16 * CbRowGroups<NS>::value, 0, st>>>(plan, dout, y, row_token, row_gate, gates, s->T, T_pad, s->d, s->E, s->dense, split, dy, dgate_row, dy_colpart)
    MOEGAN_NSEG_SWITCH(s->d / 128, MOEGAN_CB)
#undef MOEGAN_CB
}

int launch_combine_bwd(const moegan_shape* s, const int32_t* plan, const void* dout, int dout_bf16, const void* y, const int32_t* row_token,
                       const float* row_gate, const float* gates, void* dy_op, float* dgate_row, float* dy_colpart, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_COMBINE_BWD, s, dout_bf16 ? 0x1000 : 0);
    __nv_bfloat16* dy = reinterpret_cast<__nv_bfloat16*>(dy_op);
    if (s->precision == 1) {
        if (dout_bf16) combine_bwd_dispatch<true, true>(s, plan, dout, y, row_token, row_gate, gates, 1, dy, dgate_row, dy_colpart, st);
        else           combine_bwd_dispatch<true, false>(s, plan, dout, y, row_token, row_gate, gates, 1, dy, dgate_row, dy_colpart, st);
    } else {
        if (dout_bf16) combine_bwd_dispatch<false, true>(s, plan, dout, y, row_token, row_gate, gates, 0, dy, dgate_row, dy_colpart, st);
        else           combine_bwd_dispatch<false, false>(s, plan, dout, y, row_token, row_gate, gates, 0, dy, dgate_row, dy_colpart, st);
    }
    MOEGAN_CHECK_LAUNCH("combine_bwd_kernel");
    return MOEGAN_OK;
}

int slab_sum_splits(const moegan_shape* s) {
    const long long rows = s->dense ? s->T : ((long long)s->T * s->k + s->E - 1) / s->E;   // mean rows per expert
    const long long slabs = rows / 32;
    long long ns = slabs <= 64 ? 1 : slabs / 16;       // few slabs: one pass, no finishing kernel; else >= 16 slabs per split
    if (ns < 1) ns = 1;
    if (ns > 32) ns = 32;
    return (int)ns;
}

// out[e, c] = sum of expert e's slab sums part[m_cap / 32, cols];  scratch: slab_sum_splits() * E * cols floats
int launch_slab_sum(const moegan_shape* s, const int32_t* plan, const float* part, int cols, float* scratch, float* out, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_SLAB_SUM, (long long)s->m_cap, cols, s->E, prof_shape_flags(s));
    const int nsplit = slab_sum_splits(s);
    dim3 grid(ceil_div(cols, 128), s->E, nsplit);
    slab_sum_kernel<<<grid, 128, 0, st>>>(plan, part, cols, s->T, s->dense, s->E, nsplit, nsplit == 1 ? out : scratch);
    MOEGAN_CHECK_LAUNCH("slab_sum_kernel");
    if (nsplit > 1) {
        const int n = s->E * cols;
        colsum_finish_kernel<<<ceil_div(n, 256), 256, 0, st>>>(scratch, n, nsplit, out);
        MOEGAN_CHECK_LAUNCH("colsum_finish_kernel");
    }
    return MOEGAN_OK;
}

int launch_uncombine(const moegan_shape* s, const void* dx_rows, const int32_t* token_rows, const float* dx_router, float* dx,
                     cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_UNCOMBINE, s, dx_router != nullptr ? 0x1000 : 0);
    const int T_pad = row_pad(s->T);
#define MOEGAN_UC32(NS) uncombine_kernel<true, NS><<<(unsigned)ceil_div64((int64_t)s->T * NS * 16, 256 * CHUNKS_PER_THREAD), 256, 0, st>>>(dx_rows, token_rows, dx_router, s->T, T_pad, s->d, s->k, s->dense, dx)
#define MOEGAN_UC16(NS) uncombine_kernel<false, NS><<<(unsigned)ceil_div64((int64_t)s->T * NS * 16, 256 * CHUNKS_PER_THREAD), 256, 0, st>>>(dx_rows, token_rows, dx_router, s->T, T_pad, s->d, s->k, s->dense, dx)
    if (s->precision == 1) { MOEGAN_NSEG_SWITCH(s->d / 128, MOEGAN_UC32) }
    else                   { MOEGAN_NSEG_SWITCH(s->d / 128, MOEGAN_UC16) }
#undef MOEGAN_UC32
#undef MOEGAN_UC16
    MOEGAN_CHECK_LAUNCH("uncombine_kernel");
    return MOEGAN_OK;
}

int launch_colsum_grouped(const moegan_shape* s, const int32_t* plan, const void* src_op, int cols, float* scratch, float* out,
                          cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_COLSUM_GROUPED, (long long)s->m_cap, cols, s->E, prof_shape_flags(s));
    // rows per expert are only known on the device; size the split from the mean (dense: exact)
    const long long rows = s->dense ? s->T : ((long long)s->T * s->k + s->E - 1) / s->E;
    int nsplit = (int)(rows / 1024);
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 32) nsplit = 32;
    dim3 grid(cols / 64, s->E, nsplit);
    colsum_part_kernel<<<grid, 1024, 0, st>>>(plan, reinterpret_cast<const __nv_bfloat16*>(src_op), cols, s->precision == 1, s->T,
                                              s->dense, s->E, nsplit, nsplit == 1 ? out : scratch);
    MOEGAN_CHECK_LAUNCH("colsum_part_kernel");
    if (nsplit > 1) {
        const int n = s->E * cols;
        colsum_finish_kernel<<<ceil_div(n, 256), 256, 0, st>>>(scratch, n, nsplit, out);
        MOEGAN_CHECK_LAUNCH("colsum_finish_kernel");
    }
    return MOEGAN_OK;
}

// out[i] = sum_s part[s * n + i], fixed order (deterministic split-K / split-row finish)
int launch_sum_partials(const float* part, long long n, int nsplit, float* out, cudaStream_t st) {
    ProfScope _ps(st, MOEGAN_PROF_SUM_PARTIALS, n, nsplit, 0, 0);
    if (n <= 0) return MOEGAN_OK;
    colsum_finish_kernel<<<(int)ceil_div64(n, 256), 256, 0, st>>>(part, (int)n, nsplit, out);
    MOEGAN_CHECK_LAUNCH("sum_partials");
    return MOEGAN_OK;
}

}  // namespace moegan
