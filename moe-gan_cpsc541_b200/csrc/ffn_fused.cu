// Fused expert FFN: Linear(d,4d) -> GELU(erf) -> Linear(4d,d) (moegan/t2i_moe_gan.py:262-263) of a 256-row tile pair in ONE kernel.
// The [rows, 4d] hidden activations live only in TMEM / shared memory: FFN1 accumulates 128 hidden columns at a time into TMEM, the
// epilogue warps apply bias + GELU (and, in dense-soft mode, the token's gate) and write the bf16 tile into shared memory in the
// K-major SWIZZLE_128B operand layout, FFN2 consumes it from there and accumulates the [256, d] output tile in TMEM over all chunks -
// in dense-soft mode over all EXPERTS too, so the weighted combine of :465-470 is the accumulation itself:
//     out[t] = sum_e p[t,e] (gelu(x W1_e^T + b1_e) W2_e^T + b2_e) = (sum_e [p_e * H_e] W2_e^T) + sum_e p_e b2_e
// and neither H, nor the E per-expert outputs y_e, nor a combine kernel touch HBM.  (The separate-GEMM path writes and re-reads
// E*T*4d*2 B of H per forward - at d = 128 / 256 that, not the tensor pipe, bounds it: DESIGN.md section 5.)
//
// CTA pair (cta_group::2, cluster of 2): each CTA owns 128 rows of the 256-row tile and stages HALF of every weight tile, so the weight
// bytes entering one SM per chunk are half of what a single-CTA tile would stream (the L2 -> SM path is what bounds a 128-row tile).
// 18 warps per CTA:
//   warps 0-15  epilogue (TMEM lane quarter = warp % 4, 32-column group = warp / 4): GELU epilogue of every chunk, output epilogue per tile
//   warp 16     TMA producer: the X tile (once per tile), then the weight tiles in the order the MMA warp consumes them
//   warp 17     TMEM allocator; in the leader CTA one lane issues all tcgen05.mma of the pair
// TMEM (512 columns): [0, d) output accumulator | [d, d+128), [d+128, d+256) double-buffered FFN1 accumulator.
// MMA issue order (FFN1 runs two chunks ahead, across tile boundaries, so the tensor pipe never waits for the epilogue of the chunk
// it has just produced):  G1(0) G1(1) | G1(2) G2(0) | G1(3) G2(1) | ...           n = chunks per tile (dense: E * 4d/128)
// mbarrier pipelines (all waits carry a watchdog that traps instead of hanging the GPU):
//     x_full / x_empty          TMA -> MMA, per X stage              w_full[s] / w_empty[s]   TMA -> MMA, weight ring
//     a1_full[b] / a1_empty[b]  MMA -> epilogue, FFN1 accumulator    h_full[b] / h_empty[b]   epilogue -> MMA, GELU tile in smem
//     a2_full / a2_empty        MMA -> epilogue, output accumulator
// "full" barriers that the leader's MMA lane waits on collect arrivals from BOTH CTAs (TMA bytes via cta_group::2 loads, the peer's
// epilogue warps through the cluster window); "empty" barriers are signalled in both CTAs by
// multicast tcgen05.commit.
#include "umma_gemm.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <cstring>

namespace moegan {

namespace {

constexpr int BM = 128;          // rows per CTA
constexpr int NC = 128;          // hidden columns per chunk
constexpr int NEW = 16;          // epilogue warps
constexpr int WARP_TMA = 16;
constexpr int WARP_MMA = 17;
constexpr int THREADS = 18 * 32;

template <int D> struct FCfg {
    static constexpr int XS = (D == 128) ? 2 : 1;                 // X tile stages
    static constexpr int WS = (D == 128) ? 6 : 3;                 // weight ring slots
    static constexpr uint32_t X_BYTES = BM * D * 2;               // 32 / 64 KB
    static constexpr uint32_t W_BYTES = 128u * D;                 // this CTA's half of a weight chunk: 16 / 32 KB (W1: 64 x d, W2: d/2 x 128)
    static constexpr uint32_t H_BYTES = BM * NC * 2;              // 32 KB
    static constexpr uint32_t OFF_X = 0;
    static constexpr uint32_t OFF_W = XS * X_BYTES;
    static constexpr uint32_t OFF_H = OFF_W + WS * W_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_H + 2 * H_BYTES;
    static constexpr uint32_t BAR_BYTES = 256;
    static constexpr uint32_t OFF_BIAS = OFF_BAR + BAR_BYTES;     // per epilogue warp: the 32 FFN1 bias values of its columns (128 B)
    static constexpr uint32_t SMEM_BYTES = OFF_BIAS + NEW * 128;  // the dynamic shared window is declared (and checked) 1024-byte aligned
    static constexpr uint32_t TMEM_COLS = 512;
    static constexpr int CPE = 4 * D / NC;                        // chunks per expert
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(8 * (2 * XS + 2 * WS + 11) <= BAR_BYTES, "barrier region");
    static_assert(D + 2 * NC <= 512, "TMEM budget");
};

struct FusedParams {
    int mode, E, T, rows;
    const int* plan;
    const float* gates;
    const float* b1s;
    const float* b2s;
    const void* residual;
    int residual_bf16;
    void* out;
    int out_bf16;
};

// GELU epilogue arithmetic, two adjacent columns per call, entirely in packed fp16 (the tile is rounded to bf16 = 2^-9 anyway):
//   gelu(x) * gate = (0.5 * gate * x) * (1 + erf(x / sqrt 2)),   erf(x / sqrt 2) ~= tanh(x * P(min(x^2, 5.5^2)))     (common.cuh: max |erf error| 1.3e-5)
// 15 instructions per pair instead of the ~30 of the fp32 form (x0.5, (1 + t), gate and the bf16 pack each as separate fp32 ops): the
// fused kernel's per-chunk time IS this epilogue's issue time (ncu / SASS: ~500 instructions per thread and chunk before), so the
// instruction count is what sets the tensor-pipe utilisation.  Clamping x^2 (not x) keeps tanh's argument on the right side for any
// |x|: P(30.25) > 0, so |x| > 5.5 gives tanh(+-large) = +-1 exactly as the clamped form does.  hg = 0.5 * gate in both halves.
__device__ __forceinline__ uint32_t gelu_gate_pack(float x0, float x1, __half2 hg) {
    const __half2 x = __floats2half2_rn(x0, x1);
    const __half2 s = __hmin2(__hmul2(x, x), __float2half2_rn(30.25f));
    __half2 p = __hfma2(s, __float2half2_rn(-1.3688055686176932e-05f), __float2half2_rn(-0.00019445275122756945f));
    p = __hfma2(s, p, __float2half2_rn(0.036546657541717965f));
    p = __hfma2(s, p, __float2half2_rn(0.797820262700123f));
    const __half2 t = h2_tanh_approx(__hmul2(x, p));
    const __half2 h = __hmul2(x, hg);
    const float2 g = __half22float2(__hfma2(h, t, h));
    return pack_bf16x2(g.x, g.y);
}

template <int D, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW1,
                 const __grid_constant__ CUtensorMap mapW2, const FusedParams p) {
    using C = FCfg<D>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_raw_u32 = ptx::smem_u32(smem_raw);
    const uint32_t smem_base = smem_raw_u32;                      // SWIZZLE_128B atoms need 1024-byte alignment
    if ((smem_base & 1023u) != 0u) { if (threadIdx.x == 0) printf("moegan: ffn_fused: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
    const uint32_t bar_base = smem_base + C::OFF_BAR;
    auto x_full = [&](int s) { return bar_base + 8u * s; };
    auto x_empty = [&](int s) { return bar_base + 8u * (C::XS + s); };
    auto w_full = [&](int s) { return bar_base + 8u * (2 * C::XS + s); };
    auto w_empty = [&](int s) { return bar_base + 8u * (2 * C::XS + C::WS + s); };
    constexpr int B0 = 2 * C::XS + 2 * C::WS;
    auto a1_full = [&](int b) { return bar_base + 8u * (B0 + b); };
    auto a1_empty = [&](int b) { return bar_base + 8u * (B0 + 2 + b); };
    auto h_full = [&](int b) { return bar_base + 8u * (B0 + 4 + b); };
    auto h_empty = [&](int b) { return bar_base + 8u * (B0 + 6 + b); };
    const uint32_t a2_full = bar_base + 8u * (B0 + 8);
    const uint32_t a2_empty = bar_base + 8u * (B0 + 9);
    const uint32_t tmem_slot = bar_base + 8u * (B0 + 10);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_raw_u32));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = (int)ptx::cluster_ctarank();
    const int pair0 = (int)(blockIdx.x >> 1);
    const int pair_stride = (int)(gridDim.x >> 1);

    if (warp == WARP_TMA && lane == 0) {
        ptx::prefetch_tmap(&mapX);
        ptx::prefetch_tmap(&mapW1);
        ptx::prefetch_tmap(&mapW2);
        for (int s = 0; s < C::XS; ++s) { ptx::mbar_init(x_full(s), 2); ptx::mbar_init(x_empty(s), 1); }
        for (int s = 0; s < C::WS; ++s) { ptx::mbar_init(w_full(s), 2); ptx::mbar_init(w_empty(s), 1); }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(a1_full(b), 1); ptx::mbar_init(a1_empty(b), 2 * NEW);
            ptx::mbar_init(h_full(b), 2 * NEW); ptx::mbar_init(h_empty(b), 1);
        }
        ptx::mbar_init(a2_full, 1); ptx::mbar_init(a2_empty, 2 * NEW);
        ptx::fence_mbar_init();
    }
    if (warp == WARP_MMA) { ptx::tmem_alloc_cta2(tmem_slot, C::TMEM_COLS); ptx::tmem_relinquish_cta2(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    constexpr int mode = MODE;
    const int npt = mode == 1 ? (p.T + 2 * BM - 1) / (2 * BM) : __ldg(p.plan + 0) / 2;     // 256-row pair tiles
    const int n = mode == 1 ? p.E * C::CPE : C::CPE;                                         // chunks per tile
    const uint32_t w_base = smem_base + C::OFF_W;

    // The chunks of all tiles this CTA pair works on form ONE stream g = 0 .. total-1 (tile = g / n, chunk in tile = g % n; n is even, so
    // the double-buffer index g & 1 is the same on all sides).  FFN1 runs TWO chunks ahead of FFN2:
    //     G1(0) G1(1) | G1(2) G2(0) | G1(3) G2(1) | ... | G1(g+2) G2(g) | ...
    // G1(g+2) only needs the FFN1 accumulator stage that the epilogue of chunk g has just emptied - it is issued at the START of that
    // epilogue, so its result is waiting long before the epilogue warps get to chunk g+2, across tile boundaries too.  (Issuing G1(g+1)
    // only after H(g-1) was complete, as the first version did, put two barrier round trips and the FFN1 MMA on the critical path of every
    // chunk: ncu showed the epilogue warps polling a1_full 12 times per chunk.)
    const int my_tiles = pair0 < npt ? (npt - pair0 + pair_stride - 1) / pair_stride : 0;
    const int total = my_tiles * n;
    if (warp == WARP_TMA) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            uint32_t wcnt = 0;
            int e_tile = 0;
            auto load_x = [&](int tit) {                                             // tile tit of this pair: 128 rows x d of this CTA
                const int pt = pair0 + tit * pair_stride;
                const int xs = tit % C::XS;
                ptx::mbar_wait(x_empty(xs), ((uint32_t)(tit / C::XS) & 1u) ^ 1u);
                const uint32_t fbx = ptx::mapa_cluster(x_full(xs), 0);
                ptx::mbar_arrive_expect_tx_cluster(fbx, C::X_BYTES);
                const int row0 = pt * 2 * BM + rank * BM;
#pragma unroll
                for (int kb = 0; kb < D / 64; ++kb)
                    ptx::tma_load_2d_cta2(smem_base + C::OFF_X + xs * C::X_BYTES + kb * 16384u, &mapX, fbx, kb * 64, row0);
            };
            auto expert_of = [&](int g) {                                            // expert of stream chunk g
                if (mode == 1) return (g % n) / C::CPE;
                return __ldg(p.plan + PLAN_BLK + (pair0 + (g / n) * pair_stride) * 2);
            };
            auto load_w1 = [&](int g) {
                if (g % n == 0) load_x(g / n);
                const int e = expert_of(g), c = (g % n) % C::CPE;
                const int slot = (int)(wcnt % C::WS);
                ptx::mbar_wait(w_empty(slot), ((wcnt / C::WS) & 1u) ^ 1u);
                const uint32_t fb = ptx::mapa_cluster(w_full(slot), 0);
                ptx::mbar_arrive_expect_tx_cluster(fb, C::W_BYTES);
                const int row = e * 4 * D + c * NC + rank * (NC / 2);               // this CTA's 64 hidden units of the chunk
#pragma unroll
                for (int kb = 0; kb < D / 64; ++kb)
                    ptx::tma_load_2d_cta2(w_base + slot * C::W_BYTES + kb * 8192u, &mapW1, fb, kb * 64, row);
                ++wcnt;
            };
            auto load_w2 = [&](int g) {
                const int e = expert_of(g), c = (g % n) % C::CPE;
                const int slot = (int)(wcnt % C::WS);
                ptx::mbar_wait(w_empty(slot), ((wcnt / C::WS) & 1u) ^ 1u);
                const uint32_t fb = ptx::mapa_cluster(w_full(slot), 0);
                ptx::mbar_arrive_expect_tx_cluster(fb, C::W_BYTES);
                const int row = e * D + rank * (D / 2);                              // this CTA's d/2 output features
#pragma unroll
                for (int kb = 0; kb < NC / 64; ++kb)
                    ptx::tma_load_2d_cta2(w_base + slot * C::W_BYTES + kb * (uint32_t)(D / 2 * 128), &mapW2, fb, c * NC + kb * 64, row);
                ++wcnt;
            };
            (void)e_tile;
            if (total > 0) load_w1(0);
            if (total > 1) load_w1(1);
            for (int g = 0; g < total; ++g) {                                        // the order the MMA lane consumes the tiles in
                if (g + 2 < total) load_w1(g + 2);
                load_w2(g);
            }
        }
    } else if (warp == WARP_MMA) {
        // =============================== MMA issuer (leader CTA) ===============================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc1 = ptx::umma_idesc_bf16(2 * BM, NC, 0, 0);
            constexpr uint32_t idesc2 = ptx::umma_idesc_bf16(2 * BM, D, 0, 0);
            uint32_t wcnt = 0;
            auto issue_g1 = [&](int g) {                                             // acc1[g & 1] = X W1(g)^T
                const int tit = g / n, i = g % n;
                const int xs = tit % C::XS;
                if (i == 0) { ptx::mbar_wait(x_full(xs), (uint32_t)(tit / C::XS) & 1u); }
                const int b = g & 1;
                ptx::mbar_wait(a1_empty(b), (((uint32_t)g >> 1) & 1u) ^ 1u);
                const int slot = (int)(wcnt % C::WS);
                ptx::mbar_wait(w_full(slot), (wcnt / C::WS) & 1u);
                ++wcnt;
                ptx::tc_fence_after();
                const uint32_t x_addr = smem_base + C::OFF_X + xs * C::X_BYTES;
                const uint32_t w_addr = w_base + slot * C::W_BYTES;
                const uint32_t tmem_d = tmem_base + D + b * NC;
#pragma unroll
                for (int kb = 0; kb < D / 64; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_bf16_cta2(tmem_d, ptx::umma_smem_desc(x_addr + kb * 16384u + k * 32u, 16u, 1024u),
                                            ptx::umma_smem_desc(w_addr + kb * 8192u + k * 32u, 16u, 1024u), idesc1, (kb | k) != 0 ? 1u : 0u);
                ptx::umma_commit_cta2(w_empty(slot), 3);
                ptx::umma_commit_cta2(a1_full(b), 3);
                if (i == n - 1) ptx::umma_commit_cta2(x_empty(xs), 3);               // the X stage is free once the tile's last FFN1 chunk has retired
            };
            auto issue_g2 = [&](int g) {                                             // acc2 (+)= H(g) W2(g)^T
                const int tit = g / n, j = g % n;
                const int b = g & 1;
                if (j == 0) ptx::mbar_wait(a2_empty, ((uint32_t)tit & 1u) ^ 1u);     // the previous tile's output has left TMEM
                ptx::mbar_wait(h_full(b), ((uint32_t)g >> 1) & 1u);
                const int slot = (int)(wcnt % C::WS);
                ptx::mbar_wait(w_full(slot), (wcnt / C::WS) & 1u);
                ++wcnt;
                ptx::tc_fence_after();
                const uint32_t w_addr = w_base + slot * C::W_BYTES;
                const uint32_t h_addr = smem_base + C::OFF_H + b * C::H_BYTES;
#pragma unroll
                for (int kb = 0; kb < NC / 64; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_bf16_cta2(tmem_base, ptx::umma_smem_desc(h_addr + kb * 16384u + k * 32u, 16u, 1024u),
                                            ptx::umma_smem_desc(w_addr + kb * (uint32_t)(D / 2 * 128) + k * 32u, 16u, 1024u), idesc2,
                                            (j | kb | k) != 0 ? 1u : 0u);
                ptx::umma_commit_cta2(w_empty(slot), 3);
                ptx::umma_commit_cta2(h_empty(b), 3);
                if (j == n - 1) ptx::umma_commit_cta2(a2_full, 3);
            };
            if (total > 0) issue_g1(0);
            if (total > 1) issue_g1(1);
            for (int g = 0; g < total; ++g) {
                if (g + 2 < total) issue_g1(g + 2);
                issue_g2(g);
            }
        }
    } else {
        // =============================== epilogue warps ===============================
        const int q = warp & 3;             // TMEM lane quarter
        const int cg = warp >> 2;           // 32-column group of a 128-column chunk
        const int row_in = q * 32 + lane;   // this thread's row inside the CTA's 128-row tile
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        // the leader CTA's barrier block through the cluster window (the same barriers, for the leader's own warps too)
        const uint32_t lead_bars = ptx::mapa_cluster(bar_base, 0);
        auto on_leader = [&](uint32_t local_bar) { return lead_bars + (local_bar - bar_base); };
        // this thread's 64 bytes inside a GELU tile: K-block cg / 2, 16-byte chunks (cg & 1) * 4 .. + 3 of row row_in (SWIZZLE_128B)
        const uint32_t h_row_off = (uint32_t)(cg >> 1) * 16384u + (uint32_t)(row_in >> 3) * 1024u + (uint32_t)(row_in & 7) * 128u;
        const int sw = row_in & 7;
        uint32_t a1cnt = 0;
        int tit = 0;
        for (int pt = pair0; pt < npt; pt += pair_stride, ++tit) {
            const long long row = (long long)pt * 2 * BM + rank * BM + row_in;      // token (mode 1) / dispatched row (mode 0)
            const int e_tile = mode == 1 ? 0 : __ldg(p.plan + PLAN_BLK + pt * 2);
            const bool row_ok = row < (mode == 1 ? p.T : p.rows);
            // The lane's FFN1 bias value and the row's gate of chunk i + 1 are fetched while chunk i is processed (ncu on the first version:
            // long_scoreboard - the bias / gate loads issued right before their use - was 3.4 of 8.6 stalled warps per issue).
            auto bias_of = [&](int i) {
                const int e = mode == 1 ? i / C::CPE : e_tile, c = i % C::CPE;
                return __ldg(p.b1s + (size_t)e * 4 * D + c * NC + cg * 32 + lane);
            };
            auto gate_of = [&](int i) { return (mode == 1 && row_ok) ? __ldg(p.gates + row * p.E + i / C::CPE) : (mode == 1 ? 0.0f : 1.0f); };
            float bias_next = bias_of(0), gate_next = gate_of(0);
            const uint32_t bias_slot = smem_base + C::OFF_BIAS + warp * 128u;
            for (int i = 0; i < n; ++i) {
                const int b = (int)(a1cnt & 1u);
                const uint32_t ph = (a1cnt >> 1) & 1u;
                const __half2 hg = __float2half2_rn(0.5f * gate_next);
                ptx::st_shared_b32(bias_slot + lane * 4u, __float_as_uint(bias_next));     // read back below as broadcast 16-byte loads
                if (i + 1 < n) { bias_next = bias_of(i + 1); gate_next = gate_of(i + 1); }
                ptx::mbar_wait(a1_full(b), ph);
                ptx::tc_fence_after();
                uint32_t r[32];
                ptx::tmem_ld_32x32(lane_addr + D + b * NC + cg * 32, r);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(on_leader(a1_empty(b)));      // the FFN1 accumulator stage is free again
                ptx::mbar_wait(h_empty(b), ph ^ 1u);                                  // FFN2 of chunk i-2 has finished reading this buffer
                const uint32_t hrow = smem_base + C::OFF_H + b * C::H_BYTES + h_row_off;
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {                                     // 8 columns -> one 16-byte store
                    const float4 ba = ptx::ld_shared_f4(bias_slot + (uint32_t)(8 * i4) * 4u), bb = ptx::ld_shared_f4(bias_slot + (uint32_t)(8 * i4 + 4) * 4u);
                    const uint32_t o0 = gelu_gate_pack(__uint_as_float(r[8 * i4 + 0]) + ba.x, __uint_as_float(r[8 * i4 + 1]) + ba.y, hg);
                    const uint32_t o1 = gelu_gate_pack(__uint_as_float(r[8 * i4 + 2]) + ba.z, __uint_as_float(r[8 * i4 + 3]) + ba.w, hg);
                    const uint32_t o2 = gelu_gate_pack(__uint_as_float(r[8 * i4 + 4]) + bb.x, __uint_as_float(r[8 * i4 + 5]) + bb.y, hg);
                    const uint32_t o3 = gelu_gate_pack(__uint_as_float(r[8 * i4 + 6]) + bb.z, __uint_as_float(r[8 * i4 + 7]) + bb.w, hg);
                    const int chunk = (cg & 1) * 4 + i4;
                    ptx::st_shared_v4(hrow + ((uint32_t)(chunk ^ sw) << 4), o0, o1, o2, o3);
                }
                ptx::fence_proxy_async_smem();                                        // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                // Plain (release.cta) remote arrive, as CUTLASS's ClusterBarrier::arrive(cta_id) does for cross-CTA buffer hand-offs: the tile
                // is only ever read by THIS SM's tensor core (cta_group::2 reads each CTA's own half), and fence.proxy.async above has made
                // the stores visible to it.  A release.cluster arrive compiles to MEMBAR.ALL.GPU + CGAERRBAR per warp and chunk, and an
                // acquire.cluster wait to CCTL.IVALL (L1 invalidation) per poll: measured, that fixed cost - not the GELU arithmetic - set
                // the per-chunk time of the first version of this kernel (2.6 us per chunk at d = 128 whatever the instruction count).
                if (lane == 0) ptx::mbar_arrive_cluster(on_leader(h_full(b)));
                ++a1cnt;
            }
            // ---- output epilogue of the tile ----
            ptx::mbar_wait(a2_full, (uint32_t)tit & 1u);
            ptx::tc_fence_after();
            constexpr int NCH2 = D / 128;                 // 32-column chunks per warp (warp owns D/4 output columns)
#pragma unroll
            for (int ch = 0; ch < NCH2; ++ch) {
                const int col = cg * (D / 4) + ch * 32;
                uint32_t r[32];
                ptx::tmem_ld_32x32(lane_addr + col, r);
                ptx::tmem_ld_wait();
                if (ch == NCH2 - 1) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(on_leader(a2_empty));
                }
                if (row_ok) {
                    float v[32];
#pragma unroll
                    for (int k2 = 0; k2 < 32; ++k2) v[k2] = __uint_as_float(r[k2]);
                    if (mode == 1) {
                        for (int e = 0; e < p.E; ++e) {                               // + sum_e p[t,e] * b2[e, :]
                            const float g = __ldg(p.gates + row * p.E + e);
                            const float4* b2 = reinterpret_cast<const float4*>(p.b2s + (size_t)e * D + col);
#pragma unroll
                            for (int i4 = 0; i4 < 8; ++i4) {
                                const float4 b4 = __ldg(b2 + i4);
                                v[4 * i4 + 0] = fmaf(g, b4.x, v[4 * i4 + 0]); v[4 * i4 + 1] = fmaf(g, b4.y, v[4 * i4 + 1]);
                                v[4 * i4 + 2] = fmaf(g, b4.z, v[4 * i4 + 2]); v[4 * i4 + 3] = fmaf(g, b4.w, v[4 * i4 + 3]);
                            }
                        }
                        if (p.residual != nullptr) {                                  // x_spatial + moe_out (:571)
                            if (p.residual_bf16) {
                                const uint4* rs = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + row * D + col);
#pragma unroll
                                for (int i4 = 0; i4 < 4; ++i4) {
                                    const uint4 u = __ldg(rs + i4);
                                    v[8 * i4 + 0] += bf16_lo_to_float(u.x); v[8 * i4 + 1] += bf16_hi_to_float(u.x);
                                    v[8 * i4 + 2] += bf16_lo_to_float(u.y); v[8 * i4 + 3] += bf16_hi_to_float(u.y);
                                    v[8 * i4 + 4] += bf16_lo_to_float(u.z); v[8 * i4 + 5] += bf16_hi_to_float(u.z);
                                    v[8 * i4 + 6] += bf16_lo_to_float(u.w); v[8 * i4 + 7] += bf16_hi_to_float(u.w);
                                }
                            } else {
                                const float4* rs = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + row * D + col);
#pragma unroll
                                for (int i4 = 0; i4 < 8; ++i4) {
                                    const float4 f = __ldg(rs + i4);
                                    v[4 * i4 + 0] += f.x; v[4 * i4 + 1] += f.y; v[4 * i4 + 2] += f.z; v[4 * i4 + 3] += f.w;
                                }
                            }
                        }
                    } else {
                        const float4* b2 = reinterpret_cast<const float4*>(p.b2s + (size_t)e_tile * D + col);
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const float4 b4 = __ldg(b2 + i4);
                            v[4 * i4 + 0] += b4.x; v[4 * i4 + 1] += b4.y; v[4 * i4 + 2] += b4.z; v[4 * i4 + 3] += b4.w;
                        }
                    }
                    if (p.out_bf16) {
                        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * D + col);
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4)
                            dst[i4] = make_uint4(pack_bf16x2(v[8 * i4 + 0], v[8 * i4 + 1]), pack_bf16x2(v[8 * i4 + 2], v[8 * i4 + 3]),
                                                 pack_bf16x2(v[8 * i4 + 4], v[8 * i4 + 5]), pack_bf16x2(v[8 * i4 + 6], v[8 * i4 + 7]));
                    } else {
                        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * D + col);
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) dst[i4] = make_float4(v[4 * i4 + 0], v[4 * i4 + 1], v[4 * i4 + 2], v[4 * i4 + 3]);
                    }
                }
                __syncwarp();     // tcgen05.ld is warp-collective: reconverge before the next chunk
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();     // nobody leaves (or frees TMEM) while the peer may still signal / read
    if (warp == WARP_MMA) {
        __syncwarp();
        ptx::tc_fence_after();
        ptx::tmem_dealloc_cta2(tmem_base, C::TMEM_COLS);
    }
}

template <int D, int MODE>
int launch_impl(const CUtensorMap& mx, const CUtensorMap& mw1, const CUtensorMap& mw2, const FusedParams& p, int grid, cudaStream_t stream) {
    using C = FCfg<D>;
    static bool attr_set = false;
    if (!attr_set) {
        MOEGAN_CHECK_CUDA(cudaFuncSetAttribute(ffn_fused_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, ffn_fused_kernel<D, MODE>, mx, mw1, mw2, p);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(ffn_fused_kernel)");
    MOEGAN_CHECK_LAUNCH("ffn_fused_kernel");
    return MOEGAN_OK;
}

}  // namespace

bool ffn_fused_supported(int d, int precision) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("MOEGAN_DISABLE_FUSED_FFN"); enabled = (e != nullptr && e[0] == '1') ? 0 : 1; }
    return enabled == 1 && precision == 0 && (d == 128 || d == 256);
}

int launch_ffn_fused(const FfnFusedArgs& a, cudaStream_t stream) {
    MOEGAN_REQUIRE(a.d == 128 || a.d == 256, MOEGAN_ERR_UNSUPPORTED, "ffn_fused: d (%d) must be 128 or 256", a.d);
    MOEGAN_REQUIRE(a.mode == 0 || a.mode == 1, MOEGAN_ERR_UNSUPPORTED, "ffn_fused: mode %d", a.mode);
    MOEGAN_REQUIRE(a.a_op && a.W1_op && a.W2_op && a.b1s && a.b2s && a.out, MOEGAN_ERR_NULL_POINTER, "ffn_fused: NULL operand");
    MOEGAN_REQUIRE(a.mode == 0 ? a.plan != nullptr : a.gates != nullptr, MOEGAN_ERR_NULL_POINTER, "ffn_fused: plan / gates is NULL");
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.residual) | reinterpret_cast<uintptr_t>(a.b1s) |
                     reinterpret_cast<uintptr_t>(a.b2s)) & 15) == 0, MOEGAN_ERR_MISALIGNED, "ffn_fused: out / residual / biases must be 16-byte aligned");
    if (a.rows <= 0) return MOEGAN_OK;
    const int d = a.d, hid = 4 * a.d;
    CUtensorMap mx, mw1, mw2;
    int rc;
    GemmOperand X{a.a_op, (uint64_t)d, (uint64_t)a.rows, (uint64_t)d};
    GemmOperand W1{a.W1_op, (uint64_t)d, (uint64_t)a.E * hid, (uint64_t)d};
    GemmOperand W2{a.W2_op, (uint64_t)hid, (uint64_t)a.E * d, (uint64_t)hid};
    if ((rc = make_operand_tmap(&mx, X, 128u)) != MOEGAN_OK) return rc;
    if ((rc = make_operand_tmap(&mw1, W1, (uint32_t)(NC / 2))) != MOEGAN_OK) return rc;
    if ((rc = make_operand_tmap(&mw2, W2, (uint32_t)(d / 2))) != MOEGAN_OK) return rc;
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.mode = a.mode; p.E = a.E; p.T = a.T; p.rows = a.rows;
    p.plan = a.plan; p.gates = a.gates; p.b1s = a.b1s; p.b2s = a.b2s;
    p.residual = a.residual; p.residual_bf16 = a.residual_bf16; p.out = a.out; p.out_bf16 = a.out_bf16;
    const long long pairs_max = ((long long)a.rows + 2 * BM - 1) / (2 * BM);
    const int sms = num_sms();
    const int pairs = (int)(pairs_max < sms / 2 ? pairs_max : sms / 2);
    void* prof = nullptr;
    if (profile_level() >= 1) {
        // a GEMM-family record (bench.py kernel_model: kind 3 = fused FFN pair, N = 4d, K = d, two contractions of 2*R*4d*d flop each)
        const int dims[4] = {hid, d, a.mode == 1 ? a.rows * a.E : -1, a.E | (a.mode == 1 ? 0x20000 : 0)};
        profile_begin(stream, 3 * 1000 + NC + 100000 + (8 + a.mode) * 10000, nullptr, dims, &prof);
    }
    if (a.mode == 1) rc = d == 128 ? launch_impl<128, 1>(mx, mw1, mw2, p, 2 * pairs, stream) : launch_impl<256, 1>(mx, mw1, mw2, p, 2 * pairs, stream);
    else             rc = d == 128 ? launch_impl<128, 0>(mx, mw1, mw2, p, 2 * pairs, stream) : launch_impl<256, 0>(mx, mw1, mw2, p, 2 * pairs, stream);
    profile_end(stream, prof);
    return rc;
}

}  // namespace moegan
