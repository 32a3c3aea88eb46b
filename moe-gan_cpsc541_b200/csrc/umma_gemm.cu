// Grouped GEMM on 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// One persistent CTA per SM, 10 warps:
//   warps 0-7  epilogue: tcgen05.ld the 128 x BN fp32 accumulator out of TMEM, apply bias / GELU /
//              GELU' and store straight to global (warp w owns TMEM lanes 32*(w%4).. and column half w/4)
//   warp 8     TMA producer (one elected lane): fills a STAGES-deep ring of {A 128x64, B BNx64} bf16 tiles
//   warp 9     TMEM allocator + MMA issuer (one elected lane): 4 x tcgen05.mma (K = 16 each) per stage
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, two
// accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1).
//
// Work decomposition ("kinds", see umma_gemm.cuh):
//   ROWS_*  the dispatched token rows are cut into 128-row blocks; the plan array (built on the device
//           by dispatch_plan) says how many blocks are in use and which expert owns each block, so
//           ragged expert groups need no host synchronisation.  B tiles come from the owning expert's
//           weight matrix (K-major for the forward GEMMs, MN-major for dgrad).
//   WGRAD   one output tile per (expert, 128 x BN block of the weight gradient); the reduction runs
//           over that expert's rows (both operands MN-major, i.e. read "transposed" straight out of the
//           row-major activation buffers - no transposed copies are ever materialised).
// fp32 mode: operands are bf16 (hi | lo) pairs and the main loop runs 3 passes (hi*hi, hi*lo, lo*hi).
#include "umma_gemm.cuh"
#include "ptx.cuh"

#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>

namespace moegan {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                    // 64 bf16 = 128 B = one SWIZZLE_128B span
constexpr uint32_t A_BYTES = BM * BK * 2;  // 16 KB
constexpr uint32_t SMEM_LIMIT = 232448;    // 227 KB opt-in maximum per CTA

// Epilogue warps: the K = d GEMMs (FFN1, dgrad2) have a 2-8 k-block main loop under a GELU / GELU' epilogue, i.e. they
// are bound by how fast the epilogue warps drain TMEM.  With 2 warps per SM sub-partition the tcgen05.ld -> math ->
// st.shared -> fence -> TMA-store chain of each warp is latency bound (IPC 0.3), so the row GEMMs run 16 epilogue warps
// (4 per sub-partition; TMEM lane quarter = warp % 4, column group = warp / 4).  The weight-gradient GEMMs have a long
// reduction and a trivial epilogue: they keep 8 warps and spend the shared memory on pipeline stages instead.
template <int KIND> struct Warps {
    static constexpr int EPI = (KIND == GEMM_WGRAD) ? 8 : 16;
    static constexpr int TMA = EPI;
    static constexpr int MMA = EPI + 1;
    static constexpr int THREADS = (EPI + 2) * 32;
};

// bytes of TMA-store staging per epilogue warp: one 32x32 bf16 tile = 2 KB; two of them (gelu + gelu', hi + lo) or one
// 32x32 fp32 tile = 4 KB
// EPI_MUL_BF16: 2 KB output tile + a two-deep ring of 2 KB gelu' tiles that the warp TMA-loads one tile ahead
__host__ __device__ constexpr uint32_t staging_per_warp(int epi) {
    return epi == EPI_MUL_BF16 ? 6144u : (epi == EPI_BF16 || epi == EPI_BIAS_BF16) ? 2048u : 4096u;
}
__host__ __device__ constexpr bool epi_has_bias(int epi) { return epi >= EPI_BIAS_BF16 && epi <= EPI_GELU_X3; }

// CTA2 = cta_group::2: a CTA pair (cluster of 2) computes a 256 x BN tile; each CTA stages its own 128 rows of A and its
// own HALF of B (BN/2 columns), so the operand bytes entering one SM per k-block drop from 48 KB to 32 KB - at 128x256x64
// MACs per 512 tensor-pipe cycles that is 64 instead of 96 B/clk/SM, which is what the L2->SM path can actually deliver.
template <int BN, int KIND, bool CTA2, int EPI> struct Cfg {
    static constexpr int NEW = Warps<KIND>::EPI;
    static constexpr uint32_t B_BYTES = (CTA2 ? BN / 2 : BN) * BK * 2;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t TMEM_COLS = 2 * BN;
    static constexpr uint32_t STAGING_PER_WARP = staging_per_warp(EPI);
    static constexpr uint32_t STAGING_BYTES = NEW * STAGING_PER_WARP;
    static constexpr uint32_t BIAS_BYTES = NEW * 256;                 // <= 64 fp32 bias values per epilogue warp
    // barriers: 2*STAGES smem full/empty, 4 TMEM full/empty, the TMEM base slot, 2 aux-ring barriers per epilogue warp
    static constexpr uint32_t BAR_BYTES = 512;
    static constexpr uint32_t FIXED = STAGING_BYTES + BIAS_BYTES + BAR_BYTES + 1024 /* manual alignment */;
    static constexpr int FIT = (int)((SMEM_LIMIT - FIXED) / STAGE_BYTES);
    static constexpr int STAGES = FIT > 6 ? 6 : FIT;
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + FIXED;
    static_assert(STAGES >= 2, "pipeline too shallow");
    static_assert(8 * (2 * STAGES + 5 + 2 * NEW) <= BAR_BYTES, "barrier region");
};

struct Tile {
    int mb, nb, e, split;
    int nkb_pp;   // k-blocks per pass
    int row0;     // WGRAD: first reduction row of this tile (expert start + split offset)
    int rel0;     // WGRAD: the same, relative to the expert's first row (dense mode: B rows are shared by all experts)
};

// MB = rows of one scheduled tile: 128, or 256 for a CTA pair (which then splits it 128 + 128 by cluster rank)
template <int BN, int KIND, bool CTA2>
__device__ __forceinline__ int total_tiles(const GemmParams& p) {
    constexpr int MB = CTA2 ? 2 * BM : BM;
    const int n_blocks = (p.N + BN - 1) / BN;
    if (KIND == GEMM_WGRAD) return p.E * ((p.dimA + MB - 1) / MB) * n_blocks * (p.k_splits > 1 ? p.k_splits : 1);
    const int m_blocks = p.m_blocks_host >= 0 ? p.m_blocks_host : __ldg(p.plan + 0);
    return (CTA2 ? m_blocks / 2 : m_blocks) * n_blocks;
}

template <int BN, int KIND, bool CTA2>
__device__ __forceinline__ Tile decode_tile(const GemmParams& p, int tile, int rank) {
    constexpr int MB = CTA2 ? 2 * BM : BM;
    Tile t;
    const int n_blocks = (p.N + BN - 1) / BN;
    t.split = 0;
    if (KIND == GEMM_WGRAD) {
        const int mA = (p.dimA + MB - 1) / MB;
        const int ks = p.k_splits > 1 ? p.k_splits : 1;
        t.split = tile % ks;
        tile /= ks;
        t.e = tile / (mA * n_blocks);
        const int rem = tile - t.e * (mA * n_blocks);
        t.mb = rem / n_blocks;
        t.nb = rem - t.mb * n_blocks;
        int r0, r1;
        if (p.k_rows_host >= 0) { r0 = 0; r1 = p.k_rows_host; }
        else { r0 = __ldg(p.plan + PLAN_OFF + t.e); r1 = __ldg(p.plan + PLAN_OFF + t.e + 1); }
        const int nkb_all = (r1 - r0 + BK - 1) / BK;
        const int per = (nkb_all + ks - 1) / ks;                 // k-blocks per split
        const int kb0 = min(t.split * per, nkb_all);
        t.row0 = r0 + kb0 * BK;
        t.rel0 = kb0 * BK;
        t.nkb_pp = min(per, nkb_all - kb0);
    } else {
        t.mb = tile / n_blocks;
        t.nb = tile - t.mb * n_blocks;
        if (CTA2) t.mb *= 2;     // expert segments are 256-row aligned: both 128-row halves belong to the same expert
        t.e = p.m_blocks_host >= 0 ? 0 : __ldg(p.plan + PLAN_BLK + t.mb);
        t.row0 = 0;
        t.rel0 = 0;
        t.nkb_pp = p.K / BK;
    }
    if (CTA2) { if (KIND == GEMM_WGRAD) t.mb *= 2; t.mb += rank; }   // this CTA's 128-row half of the pair tile
    return t;
}

// ---- epilogue staging helpers --------------------------------------------------------------------------
// One thread = one output row, 32 consecutive columns.  The warp's 32 x 32 tile is written to shared memory in the
// TMA swizzle layout (conflict-free 16-byte stores) and leaves the SM as ONE coalesced bulk tensor store.
//   bf16: rows of 64 B, SWIZZLE_64B  : 16-byte chunk i of row r sits at chunk i ^ ((r >> 1) & 3)
//   fp32: rows of 128 B, SWIZZLE_128B: 16-byte chunk i of row r sits at chunk i ^ (r & 7)
__device__ __forceinline__ void stage32_bf16(uint32_t base, int r, const float (&v)[32]) {
    const uint32_t row = base + r * 64;
    const int sw = (r >> 1) & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        ptx::st_shared_v4(row + ((i ^ sw) << 4), pack_bf16x2(v[8 * i + 0], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                          pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
}
__device__ __forceinline__ void stage32_f32(uint32_t base, int r, const float (&v)[32]) {
    const uint32_t row = base + r * 128;
    const int sw = r & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        ptx::st_shared_v4(row + ((i ^ sw) << 4), __float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]),
                          __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
}
// inverse of stage32_bf16: row r of a 32x32 bf16 tile that a SWIZZLE_64B TMA load wrote to shared memory
__device__ __forceinline__ void unstage32_bf16(uint32_t base, int r, uint4 (&u)[4]) {
    const uint32_t row = base + r * 64;
    const int sw = (r >> 1) & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = ptx::ld_shared_u4(row + ((i ^ sw) << 4));
}
__device__ __forceinline__ void unpack32_bf16(const uint4 (&u)[4], float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[8 * i + 0] = bf16_lo_to_float(u[i].x); v[8 * i + 1] = bf16_hi_to_float(u[i].x);
        v[8 * i + 2] = bf16_lo_to_float(u[i].y); v[8 * i + 3] = bf16_hi_to_float(u[i].y);
        v[8 * i + 4] = bf16_lo_to_float(u[i].z); v[8 * i + 5] = bf16_hi_to_float(u[i].z);
        v[8 * i + 6] = bf16_lo_to_float(u[i].w); v[8 * i + 7] = bf16_hi_to_float(u[i].w);
    }
}
// v[i] of lane r = element (row r, column i) of a 32x32 tile.  Returns, in lane c, the sum of column c over the 32 rows:
// a butterfly that halves the number of live columns per step (31 shuffles).  Destroys v.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}
__device__ __forceinline__ void load32_f32(const float* src, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(src) + i);
        v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
    }
}

template <int BN, int KIND, bool CTA2, int EPI>
__global__ void __launch_bounds__(Warps<KIND>::THREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapO0, const __grid_constant__ CUtensorMap mapO1, const GemmParams p) {
    using C = Cfg<BN, KIND, CTA2, EPI>;
    constexpr int NUM_EPI_WARPS = Warps<KIND>::EPI;
    constexpr int WARP_TMA = Warps<KIND>::TMA;
    constexpr int WARP_MMA = Warps<KIND>::MMA;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_raw_u32 = ptx::smem_u32(smem_raw);
    const uint32_t smem_base = (smem_raw_u32 + 1023u) & ~1023u;   // SWIZZLE_128B atoms need 1024 B alignment
    const uint32_t staging_base = smem_base + C::STAGES * C::STAGE_BYTES;        // 1024-aligned
    const uint32_t bias_base = staging_base + C::STAGING_BYTES;
    const uint32_t bar_base = bias_base + C::BIAS_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_raw_u32));
    auto aux_bar = [&](int w, int b) { return bar_base + 8u * (2 * C::STAGES + 5 + 2 * w + b); };   // EPI_MUL_BF16 only

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = CTA2 ? (int)ptx::cluster_ctarank() : 0;     // 0 = leader CTA of the pair (issues the MMAs)
    const int tile0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_stride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == WARP_TMA && lane == 0) {
        ptx::prefetch_tmap(&mapA);
        ptx::prefetch_tmap(&mapB);
        ptx::prefetch_tmap(&mapO0);
        ptx::prefetch_tmap(&mapO1);
        // CTA2: the leader's full barrier collects both producers' arrive.expect_tx (and both CTAs' TMA bytes); the leader's
        // TMEM-empty barrier collects the epilogue warps of both CTAs
        for (int s = 0; s < C::STAGES; ++s) { ptx::mbar_init(full_bar(s), CTA2 ? 2 : 1); ptx::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(tfull_bar(a), 1); ptx::mbar_init(tempty_bar(a), CTA2 ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS); }
        if (EPI == EPI_MUL_BF16)
            for (int w = 0; w < NUM_EPI_WARPS; ++w) { ptx::mbar_init(aux_bar(w, 0), 1); ptx::mbar_init(aux_bar(w, 1), 1); }
        ptx::fence_mbar_init();
    }
    if (warp == WARP_MMA) {
        if (CTA2) { ptx::tmem_alloc_cta2(tmem_slot, C::TMEM_COLS); ptx::tmem_relinquish_cta2(); }
        else      { ptx::tmem_alloc(tmem_slot, C::TMEM_COLS); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (CTA2) ptx::cluster_sync_all();     // peer barriers are initialised before anybody signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int ntiles = total_tiles<BN, KIND, CTA2>(p);
    const uint32_t mn_lbo = p.mn_lbo ? p.mn_lbo : 8192u;   // bytes between 64-element MN groups (one TMA box)
    const uint32_t mn_sbo = p.mn_sbo ? p.mn_sbo : 1024u;   // bytes between 8-row K groups
    const uint32_t mn_kadv = p.mn_kadv ? p.mn_kadv : 2048u; // bytes per UMMA_K = 16 rows

    if (warp == WARP_TMA) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            constexpr int BNL = CTA2 ? BN / 2 : BN;          // B columns staged by THIS CTA
            const int bn_off = CTA2 ? rank * BNL : 0;
            for (int tile = tile0; tile < ntiles; tile += tile_stride) {
                const Tile t = decode_tile<BN, KIND, CTA2>(p, tile, rank);
                const int nkb = t.nkb_pp * p.passes;
                int pass = 0, kk = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t b_dst = a_dst + A_BYTES;
                    // completion barrier: local, or (CTA2) the LEADER's barrier addressed through the cluster window
                    const uint32_t fbar = CTA2 ? ptx::mapa_cluster(full_bar(stage), 0) : full_bar(stage);
                    if (CTA2) ptx::mbar_arrive_expect_tx_cluster(fbar, C::STAGE_BYTES);
                    else      ptx::mbar_arrive_expect_tx(fbar, C::STAGE_BYTES);
                    auto ld = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1) {
                        if (CTA2) ptx::tma_load_2d_cta2(dst, m, fbar, c0, c1);
                        else      ptx::tma_load_2d(dst, m, fbar, c0, c1);
                    };
                    const int a_lo = (pass == 2) ? p.a_lo_off : 0;
                    const int b_lo = (pass == 1) ? p.b_lo_off : 0;
                    if (KIND == GEMM_WGRAD) {
                        const int krow = t.row0 + kk * BK;
#pragma unroll
                        for (int i = 0; i < 2; ++i)
                            ld(a_dst + i * 8192u, &mapA, a_lo + t.mb * BM + i * 64, krow);
                        const int brow = p.b_rel ? t.rel0 + kk * BK : krow;
#pragma unroll
                        for (int i = 0; i < BNL / 64; ++i)
                            ld(b_dst + i * 8192u, &mapB, b_lo + t.nb * BN + bn_off + i * 64, brow);
                    } else {
                        const int ablk = p.a_mod > 0 ? (t.mb % p.a_mod) : t.mb;
                        ld(a_dst, &mapA, a_lo + kk * BK, ablk * BM);
                        if (KIND == GEMM_ROWS_BK) {
                            ld(b_dst, &mapB, b_lo + kk * BK, t.e * p.b_rows_per_expert + t.nb * BN + bn_off);
                        } else {
                            const int brow = t.e * p.b_rows_per_expert + kk * BK;
#pragma unroll
                            for (int i = 0; i < BNL / 64; ++i)
                                ld(b_dst + i * 8192u, &mapB, b_lo + t.nb * BN + bn_off + i * 64, brow);
                        }
                    }
                    if (++kk == t.nkb_pp) { kk = 0; ++pass; }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // =============================== MMA issuer ===============================
        if (lane == 0 && rank == 0) {                       // CTA2: only the leader issues; it drives both SMs' tensor cores
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(CTA2 ? 2 * BM : BM, BN, KIND == GEMM_WGRAD ? 1 : 0, KIND == GEMM_ROWS_BK ? 0 : 1);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = tile0; tile < ntiles; tile += tile_stride) {
                const Tile t = decode_tile<BN, KIND, CTA2>(p, tile, rank);
                const int nkb = t.nkb_pp * p.passes;
                ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);   // epilogue (of both CTAs) has drained this accumulator
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                if (nkb == 0) {                                     // empty expert: epilogue writes zeros
                    ptx::mbar_arrive(tfull_bar(acc));
                    if (CTA2) ptx::mbar_arrive_cluster(ptx::mapa_cluster(tfull_bar(acc), 1));
                }
                for (int kb = 0; kb < nkb; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = (KIND == GEMM_WGRAD) ? ptx::umma_smem_desc(a_addr + k * mn_kadv, mn_lbo, mn_sbo)
                                                                 : ptx::umma_smem_desc(a_addr + k * 32u, 16u, 1024u);
                        const uint64_t db = (KIND == GEMM_ROWS_BK) ? ptx::umma_smem_desc(b_addr + k * 32u, 16u, 1024u)
                                                                   : ptx::umma_smem_desc(b_addr + k * mn_kadv, mn_lbo, mn_sbo);
                        if (CTA2) ptx::umma_bf16_cta2(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        else      ptx::umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    // smem slot reusable once these MMAs retire (CTA2: signalled in BOTH CTAs - each producer waits locally)
                    if (CTA2) { ptx::umma_commit_cta2(empty_bar(stage), 3); if (kb == nkb - 1) ptx::umma_commit_cta2(tfull_bar(acc), 3); }
                    else      { ptx::umma_commit(empty_bar(stage));         if (kb == nkb - 1) ptx::umma_commit(tfull_bar(acc)); }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
                acc ^= 1; if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // =============================== epilogue warps ===============================
        // TMEM -> registers -> (bias / GELU / GELU') -> swizzled smem staging -> ONE bulk tensor store per 32x32 tile.
        constexpr int CGROUPS = NUM_EPI_WARPS / 4;      // column groups (4 for the row GEMMs, 2 for WGRAD)
        constexpr int CW = BN / CGROUPS;                // columns per warp: 32 .. 128
        constexpr int NCH = CW / 32;                    // 32-column chunks per warp and tile
        constexpr bool HAS_BIAS = epi_has_bias(EPI);
        static_assert(!HAS_BIAS || CW <= 64, "bias slot holds 64 values");
        const int q = warp & 3;            // TMEM lane quarter this warp may read
        const int cg = warp >> 2;          // column group
        const int row_in_tile = q * 32 + lane;
        const uint32_t stage0 = staging_base + warp * C::STAGING_PER_WARP;
        const uint32_t bias_slot = bias_base + warp * 256u;
        int acc = 0; uint32_t acc_phase = 0;
        bool pending = false;              // this warp has a bulk store in flight that still reads its staging buffer
        // free the staging buffer of the previous store, write the new tile(s), hand them to the TMA unit
        auto flush_begin = [&]() {
            if (pending) { if (lane == 0) ptx::bulk_wait_read0(); __syncwarp(); }
        };
        auto flush_end = [&]() {
            ptx::fence_proxy_async_smem();
            __syncwarp();
        };
        const int cbase = cg * CW;                                  // first column of this warp inside the tile
        // EPI_MUL_BF16: the gelu' tile (aux) does not depend on the accumulator.  Each warp TMA-loads its own 32x32 aux
        // tiles into a two-deep ring, two chunks (= one 256-column tile) ahead of their use: exact bytes, DRAM latency off
        // the critical path.  (Plain row-strided loads thrash the ~30 KB of L1 left next to 225 KB of shared memory: ncu
        // showed every 32-byte sector fetched twice.)  chunk sequence number cc -> (tile0 + cc / NCH * stride, cc % NCH).
        uint32_t aux_cc = 0, aux_phase = 0;
        auto aux_issue = [&](uint32_t cc) {                          // lane 0 only
            const int tl = tile0 + (int)(cc / NCH) * tile_stride;
            if (tl >= ntiles) return;
            const Tile tn = decode_tile<BN, KIND, CTA2>(p, tl, rank);
            const int col = tn.nb * BN + cbase + (int)(cc % NCH) * 32;
            const int row = tn.mb * BM + q * 32;
            if (row < p.out_rows_valid && col < p.N) {               // same predicate as the consumer's (slab_ok && col0 < N)
                const uint32_t bar = aux_bar(warp, cc & 1);
                ptx::mbar_arrive_expect_tx(bar, 2048u);
                ptx::tma_load_2d(stage0 + 2048u + (cc & 1u) * 2048u, &mapO1, bar, col, row);
            }
        };
        if (EPI == EPI_MUL_BF16 && lane == 0) { aux_issue(0); aux_issue(1); }
        for (int tile = tile0; tile < ntiles; tile += tile_stride) {
            const Tile t = decode_tile<BN, KIND, CTA2>(p, tile, rank);
            const int nkb = t.nkb_pp * p.passes;
            // bias of this warp's CW (<= 64) columns -> private smem slot, fetched before the accumulator wait and read
            // back as broadcast 16-byte loads (all lanes hold different ROWS of the same columns)
            if (HAS_BIAS) {
                constexpr int PER = CW / 32;                        // 1 or 2 values per lane
                const int bc = t.nb * BN + cbase + lane * PER;
                const float* bsrc = p.bias + (long long)t.e * p.N + bc;
                if (PER == 2) {
                    float2 b2 = make_float2(0.f, 0.f);
                    if (bc < p.N) b2 = __ldg(reinterpret_cast<const float2*>(bsrc));
                    ptx::st_shared_v2(bias_slot + lane * 8u, __float_as_uint(b2.x), __float_as_uint(b2.y));
                } else {
                    ptx::st_shared_b32(bias_slot + lane * 4u, __float_as_uint(bc < p.N ? __ldg(bsrc) : 0.f));
                }
                __syncwarp();
            }
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            const int grow = t.mb * BM + row_in_tile;      // this thread's row inside the (expert's) output
            // first row of this warp's 32-row slab in the output tensor
            int out_row0; bool slab_ok;
            if (KIND == GEMM_WGRAD) { out_row0 = (t.split * p.E + t.e) * p.dimA + t.mb * BM + q * 32; slab_ok = (t.mb * BM + q * 32) < p.dimA; }
            else                    { out_row0 = t.mb * BM + q * 32; slab_ok = out_row0 < p.out_rows_valid; }
            const long long out_row = (long long)out_row0 + lane;
            const bool row_ok = (KIND == GEMM_WGRAD) ? (grow < p.dimA) : (out_row < p.out_rows_valid);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int c = cbase + ch * 32;
                float v[32];
                if (nkb > 0) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c, r);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.0f;
                }
                if (ch == NCH - 1) {
                    // the accumulator is in registers: hand the TMEM stage back to the MMA warp before the math / stores
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (CTA2 && rank != 0) ptx::mbar_arrive_cluster(ptx::mapa_cluster(tempty_bar(acc), 0));   // leader's barrier
                        else                   ptx::mbar_arrive(tempty_bar(acc));
                    }
                }
                const int col0 = t.nb * BN + c;
                if (slab_ok && col0 < p.N) {   // warp-uniform
                    if (HAS_BIAS) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 b4 = ptx::ld_shared_f4(bias_slot + (uint32_t)(ch * 32 + 4 * i) * 4u);
                            v[4 * i + 0] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
                        }
                    }
                    if (EPI == EPI_F32 || EPI == EPI_BIAS_F32) {
                        flush_begin();
                        stage32_f32(stage0, lane, v);
                        flush_end();
                        if (lane == 0) { ptx::tma_store_2d(&mapO0, stage0, col0, out_row0); ptx::bulk_commit(); }
                        pending = true;
                    } else if (EPI == EPI_BF16 || EPI == EPI_BIAS_BF16) {
                        flush_begin();
                        stage32_bf16(stage0, lane, v);
                        flush_end();
                        if (lane == 0) { ptx::tma_store_2d(&mapO0, stage0, col0, out_row0); ptx::bulk_commit(); }
                        pending = true;
                    } else if (EPI == EPI_GELU_BF16) {
                        if (p.out1 != nullptr) {
                            float dg[32];
#pragma unroll
                            for (int i = 0; i < 32; i += 2) gelu_pair_h2<true>(v[i], v[i + 1], v[i], v[i + 1], dg[i], dg[i + 1]);
                            flush_begin();
                            stage32_bf16(stage0 + 2048u, lane, dg);
                        } else {
                            float u0, u1;
#pragma unroll
                            for (int i = 0; i < 32; i += 2) gelu_pair_h2<false>(v[i], v[i + 1], v[i], v[i + 1], u0, u1);
                            flush_begin();
                        }
                        stage32_bf16(stage0, lane, v);
                        flush_end();
                        if (lane == 0) {
                            ptx::tma_store_2d(&mapO0, stage0, col0, out_row0);
                            if (p.out1 != nullptr) ptx::tma_store_2d(&mapO1, stage0 + 2048u, col0, out_row0);
                            ptx::bulk_commit();
                        }
                        pending = true;
                    } else if (EPI == EPI_GELU_X3) {
                        if (p.out1 != nullptr) {   // fp32 pre-activation
                            flush_begin();
                            stage32_f32(stage0, lane, v);
                            flush_end();
                            if (lane == 0) { ptx::tma_store_2d(&mapO1, stage0, col0, out_row0); ptx::bulk_commit(); }
                            pending = true;
                        }
                        float lo[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) { v[i] = gelu_erf<true>(v[i]); float h; split_bf16(v[i], h, lo[i]); }
                        flush_begin();
                        stage32_bf16(stage0, lane, v);            // rounding v to bf16 gives exactly hi
                        stage32_bf16(stage0 + 2048u, lane, lo);
                        flush_end();
                        if (lane == 0) {
                            ptx::tma_store_2d(&mapO0, stage0, col0, out_row0);
                            ptx::tma_store_2d(&mapO0, stage0 + 2048u, col0 + p.out_split_off, out_row0);
                            ptx::bulk_commit();
                        }
                        pending = true;
                    } else if (EPI == EPI_MUL_BF16) {
                        const uint32_t b = aux_cc & 1u;
                        ptx::mbar_wait(aux_bar(warp, b), (aux_phase >> b) & 1u);
                        aux_phase ^= 1u << b;
                        uint4 u[4];
                        unstage32_bf16(stage0 + 2048u + b * 2048u, lane, u);   // rows the TMA clipped arrive as zeros
                        __syncwarp();                                          // every lane has read the ring slot
                        if (lane == 0) aux_issue(aux_cc + 2);
                        float a[32];
                        unpack32_bf16(u, a);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] *= a[i];
                        flush_begin();
                        stage32_bf16(stage0, lane, v);
                        flush_end();
                        if (lane == 0) { ptx::tma_store_2d(&mapO0, stage0, col0, out_row0); ptx::bulk_commit(); }
                        pending = true;
                        if (p.colsum_part != nullptr) {
                            const float cs = warp_colsum32(v, lane);
                            p.colsum_part[(size_t)(t.mb * 4 + q) * p.N + col0 + lane] = cs;
                        }
                    } else if (EPI == EPI_DGELU_X3) {
                        float a[32];
                        if (row_ok) load32_f32(reinterpret_cast<const float*>(p.aux) + out_row * p.ld_aux + col0, a);
                        else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) a[i] = 0.0f;
                        }
                        float lo[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) { v[i] *= gelu_erf_grad<true>(a[i]); float h; split_bf16(v[i], h, lo[i]); }
                        flush_begin();
                        stage32_bf16(stage0, lane, v);
                        stage32_bf16(stage0 + 2048u, lane, lo);
                        flush_end();
                        if (lane == 0) {
                            ptx::tma_store_2d(&mapO0, stage0, col0, out_row0);
                            ptx::tma_store_2d(&mapO0, stage0 + 2048u, col0 + p.out_split_off, out_row0);
                            ptx::bulk_commit();
                        }
                        pending = true;
                        if (p.colsum_part != nullptr) {
                            const float cs = warp_colsum32(v, lane);
                            p.colsum_part[(size_t)(t.mb * 4 + q) * p.N + col0 + lane] = cs;
                        }
                    }
                }
                ++aux_cc;
                __syncwarp();   // tcgen05.ld is warp-collective: reconverge before the next chunk
            }
            acc ^= 1; if (acc == 0) acc_phase ^= 1u;
        }
        if (pending && lane == 0) ptx::bulk_wait0();   // all bulk stores of this warp have been written
        __syncwarp();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (CTA2) ptx::cluster_sync_all();     // nobody leaves (or frees TMEM) while the peer may still signal / read
    if (warp == WARP_MMA) {
        __syncwarp();
        ptx::tc_fence_after();
        if (CTA2) ptx::tmem_dealloc_cta2(tmem_base, C::TMEM_COLS);
        else      ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---- host: tensor maps -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 32 x 32 output tile: bf16 -> 64-byte rows (SWIZZLE_64B), fp32 -> 128-byte rows (SWIZZLE_128B)
int make_out_tmap(CUtensorMap* m, const void* ptr, bool f32, uint64_t cols, uint64_t rows, uint64_t pitch_elems) {
    EncodeTiledFn enc = get_encode_fn();
    MOEGAN_REQUIRE(enc != nullptr, MOEGAN_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    const uint64_t es = f32 ? 4 : 2;
    MOEGAN_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (pitch_elems * es) % 16 == 0, MOEGAN_ERR_MISALIGNED,
                   "GEMM output not 16-byte aligned (pitch %llu elements)", (unsigned long long)pitch_elems);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * es};
    cuuint32_t box[2] = {32u, 32u};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MOEGAN_REQUIRE(r == CUDA_SUCCESS, MOEGAN_ERR_DRIVER, "cuTensorMapEncodeTiled (output) failed (CUresult %d; cols %llu rows %llu pitch %llu)",
                   (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch_elems);
    return MOEGAN_OK;
}

int make_tmap(CUtensorMap* m, const GemmOperand& o, uint32_t box_outer) {
    EncodeTiledFn enc = get_encode_fn();
    MOEGAN_REQUIRE(enc != nullptr, MOEGAN_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    MOEGAN_REQUIRE((reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0, MOEGAN_ERR_MISALIGNED, "GEMM operand not 16-byte aligned");
    MOEGAN_REQUIRE((o.pitch * 2) % 16 == 0, MOEGAN_ERR_MISALIGNED, "GEMM operand row pitch (%llu elements) not a multiple of 16 bytes",
                   (unsigned long long)o.pitch);
    cuuint64_t dims[2] = {o.inner, o.outer};
    cuuint64_t strides[1] = {o.pitch * 2};
    cuuint32_t box[2] = {64u, box_outer};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(o.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MOEGAN_REQUIRE(r == CUDA_SUCCESS, MOEGAN_ERR_DRIVER, "cuTensorMapEncodeTiled failed (CUresult %d; inner %llu outer %llu pitch %llu box %u)",
                   (int)r, (unsigned long long)o.inner, (unsigned long long)o.outer, (unsigned long long)o.pitch, box_outer);
    return MOEGAN_OK;
}

template <int BN, int KIND, bool CTA2, int EPI>
int launch_impl(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo0, const CUtensorMap& mo1, const GemmParams& p, int grid,
                cudaStream_t stream) {
    using C = Cfg<BN, KIND, CTA2, EPI>;
    static bool attr_set = false;
    if (!attr_set) {
        MOEGAN_CHECK_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<BN, KIND, CTA2, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)C::SMEM_BYTES));
        attr_set = true;
    }
    void* prof = nullptr;
    if (profile_level() >= 1) {
        const int dims[4] = {p.N, KIND == GEMM_WGRAD ? p.dimA : p.K, KIND == GEMM_WGRAD ? p.k_rows_host : p.m_blocks_host * 128,
                             p.E | (p.out1 != nullptr ? 0x10000 : 0)};
        profile_begin(stream, KIND * 1000 + BN + (CTA2 ? 100000 : 0) + EPI * 10000, nullptr, dims, &prof);
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(Warps<KIND>::THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTA2 ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, umma_gemm_kernel<BN, KIND, CTA2, EPI>, ma, mb, mo0, mo1, p);
    profile_end(stream, prof);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(umma_gemm_kernel)");
    MOEGAN_CHECK_LAUNCH("umma_gemm_kernel");
    return MOEGAN_OK;
}

// tile width / CTA-pair variants of one (kind, epilogue) combination
template <int KIND, int EPI>
int launch_variant(int BN, bool cta2, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo0, const CUtensorMap& mo1,
                   const GemmParams& p, int grid, cudaStream_t stream) {
    if (cta2) return launch_impl<256, KIND, true, EPI>(ma, mb, mo0, mo1, p, grid, stream);
    if (BN == 256) return launch_impl<256, KIND, false, EPI>(ma, mb, mo0, mo1, p, grid, stream);
    return launch_impl<128, KIND, false, EPI>(ma, mb, mo0, mo1, p, grid, stream);
}

bool cta2_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MOEGAN_DISABLE_CTA2"); v = (e != nullptr && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

}  // namespace

int make_operand_tmap(CUtensorMap* m, const GemmOperand& o, uint32_t box_outer) { return make_tmap(m, o, box_outer); }

int launch_umma_gemm(int kind, int BN, const GemmOperand& A, const GemmOperand& B, GemmParams p, long long max_tiles,
                     cudaStream_t stream) {
    if (max_tiles <= 0) return MOEGAN_OK;
    MOEGAN_REQUIRE(BN == 128 || BN == 256, MOEGAN_ERR_UNSUPPORTED, "BN must be 128 or 256");
    MOEGAN_REQUIRE(p.N % 32 == 0, MOEGAN_ERR_BAD_SHAPE, "GEMM N (%d) must be a multiple of 32", p.N);
    MOEGAN_REQUIRE(kind == GEMM_WGRAD || p.K % BK == 0, MOEGAN_ERR_BAD_SHAPE, "GEMM K (%d) must be a multiple of 64", p.K);
    MOEGAN_REQUIRE(p.passes == 1 || p.passes == 3, MOEGAN_ERR_BAD_SHAPE, "passes must be 1 or 3");
    CUtensorMap ma, mb;
    int rc;
    // CTA pairs (cta_group::2, 256-row tiles) whenever the shape allows: full 256-column blocks, and either the grouped
    // row plan (expert segments are 256-row aligned, see MOEGAN_ROW_ALIGN) or a weight-gradient with 256 | dimA
    const bool cta2 = cta2_enabled() && BN == 256 && p.N % 256 == 0 &&
                      (kind == GEMM_WGRAD ? (p.dimA % 256 == 0) : (p.m_blocks_host < 0 || p.m_blocks_host % 2 == 0));
    const uint32_t a_box = (kind == GEMM_WGRAD) ? 64u : 128u;
    const uint32_t b_box = (kind == GEMM_ROWS_BK) ? (uint32_t)(cta2 ? BN / 2 : BN) : 64u;
    if ((rc = make_tmap(&ma, A, a_box)) != MOEGAN_OK) return rc;
    if ((rc = make_tmap(&mb, B, b_box)) != MOEGAN_OK) return rc;
    // output tensor maps (TMA-store epilogue)
    CUtensorMap mo0, mo1;
    MOEGAN_REQUIRE(p.out0 != nullptr, MOEGAN_ERR_NULL_POINTER, "GEMM output is NULL");
    const bool o0_f32 = (p.epi == EPI_F32 || p.epi == EPI_BIAS_F32);
    const uint64_t o_rows = (kind == GEMM_WGRAD) ? (uint64_t)(p.k_splits > 1 ? p.k_splits : 1) * p.E * p.dimA : (uint64_t)p.out_rows_valid;
    const bool o0_split = (p.epi == EPI_GELU_X3 || p.epi == EPI_DGELU_X3);
    const uint64_t o0_cols = o0_split ? (uint64_t)p.out_split_off + p.N : (uint64_t)p.N;
    if ((rc = make_out_tmap(&mo0, p.out0, o0_f32, o0_cols, o_rows, (uint64_t)p.ld_out0)) != MOEGAN_OK) return rc;
    if (p.epi == EPI_MUL_BF16) {
        // no second output: the slot carries the map of the aux tensor (same 32x32 SWIZZLE_64B boxes as a bf16 output)
        MOEGAN_REQUIRE(p.aux != nullptr, MOEGAN_ERR_NULL_POINTER, "GEMM aux tensor is NULL");
        if ((rc = make_out_tmap(&mo1, p.aux, false, (uint64_t)p.N, o_rows, (uint64_t)p.ld_aux)) != MOEGAN_OK) return rc;
    } else if (p.out1 != nullptr) {
        if ((rc = make_out_tmap(&mo1, p.out1, p.epi == EPI_GELU_X3, (uint64_t)p.N, o_rows, (uint64_t)p.ld_out1)) != MOEGAN_OK) return rc;
    } else {
        mo1 = mo0;
    }
    const int sms = num_sms();
    int grid;
    if (cta2) {
        long long pairs = (max_tiles + 1) / 2;               // max_tiles counts 128-row tiles
        if (pairs > sms / 2) pairs = sms / 2;
        grid = (int)(2 * pairs);
    } else {
        grid = (int)(max_tiles < sms ? max_tiles : sms);
    }
    // the (kind, epilogue) combinations the library uses; the epilogue is a template parameter so that every instance
    // carries only its own code path (register budget of the 18-warp row GEMMs: 112 per thread)
#define MOEGAN_CASE(KIND_, EPI_) \
    if (kind == KIND_ && p.epi == EPI_) return launch_variant<KIND_, EPI_>(BN, cta2, ma, mb, mo0, mo1, p, grid, stream)
    MOEGAN_CASE(GEMM_ROWS_BK, EPI_GELU_BF16);
    MOEGAN_CASE(GEMM_ROWS_BK, EPI_BIAS_BF16);
    MOEGAN_CASE(GEMM_ROWS_BK, EPI_F32);
    MOEGAN_CASE(GEMM_ROWS_BK, EPI_GELU_X3);
    MOEGAN_CASE(GEMM_ROWS_BK, EPI_BIAS_F32);
    MOEGAN_CASE(GEMM_ROWS_BMN, EPI_MUL_BF16);
    MOEGAN_CASE(GEMM_ROWS_BMN, EPI_BF16);
    MOEGAN_CASE(GEMM_ROWS_BMN, EPI_F32);
    MOEGAN_CASE(GEMM_ROWS_BMN, EPI_DGELU_X3);
    MOEGAN_CASE(GEMM_WGRAD, EPI_F32);
#undef MOEGAN_CASE
    set_error("unsupported GEMM kind %d / epilogue %d combination", kind, p.epi);
    return MOEGAN_ERR_UNSUPPORTED;
}

}  // namespace moegan
