// extern "C" boundary of libmoegan_sm100.so (declared in include/moegan_sm100.h).
#include "kernels.cuh"
#include "umma_gemm.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace moegan {

// ---- error plumbing --------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return MOEGAN_ERR_CUDA;
}
int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional profiling (bench.py): see common.cuh ----
struct ProfRec { cudaEvent_t beg, end; int tag; int meta[8]; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::atomic<int> g_prof_level{0};
static thread_local int t_prof_shape[4] = {0, 0, 0, 0};
int profile_level() { return g_prof_level.load(std::memory_order_relaxed); }
void profile_note_shape(const moegan_shape* s) {
    if (g_prof_level.load(std::memory_order_relaxed) == 0 || s == nullptr) return;
    t_prof_shape[0] = s->T; t_prof_shape[1] = s->d; t_prof_shape[2] = s->E; t_prof_shape[3] = prof_shape_flags(s);
}
bool profile_begin(cudaStream_t st, int tag, const int* meta_lo, const int* meta_hi, void** handle) {
    *handle = nullptr;
    if (profile_level() == 0) return false;
    ProfRec* r = new ProfRec;
    r->tag = tag;
    for (int i = 0; i < 4; ++i) {
        r->meta[i] = meta_lo != nullptr ? meta_lo[i] : t_prof_shape[i];
        r->meta[4 + i] = meta_hi != nullptr ? meta_hi[i] : 0;
    }
    if (cudaEventCreate(&r->beg) != cudaSuccess || cudaEventCreate(&r->end) != cudaSuccess) { delete r; return false; }
    cudaEventRecord(r->beg, st);
    *handle = r;
    return true;
}
void profile_end(cudaStream_t st, void* handle) {
    if (handle == nullptr) return;
    ProfRec* r = reinterpret_cast<ProfRec*>(handle);
    cudaEventRecord(r->end, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(*r);
    delete r;
}

int validate_shape(const moegan_shape* s) {
    MOEGAN_REQUIRE(s != nullptr, MOEGAN_ERR_NULL_POINTER, "shape is NULL");
    MOEGAN_REQUIRE(s->T >= 0 && s->B >= 0 && s->tokens_per_sample > 0 && (int64_t)s->B * s->tokens_per_sample == s->T,
                   MOEGAN_ERR_BAD_SHAPE, "T (%d) must equal B (%d) * tokens_per_sample (%d)", s->T, s->B, s->tokens_per_sample);
    MOEGAN_REQUIRE(s->d >= 128 && s->d % 128 == 0 && s->d <= 1024, MOEGAN_ERR_BAD_SHAPE, "d (%d) must be a multiple of 128 in [128, 1024]", s->d);
    MOEGAN_REQUIRE(s->Dt >= 64 && s->Dt % 64 == 0, MOEGAN_ERR_BAD_SHAPE, "Dt (%d) must be a positive multiple of 64", s->Dt);
    MOEGAN_REQUIRE(s->E >= 1 && s->E <= MOEGAN_MAX_EXPERTS, MOEGAN_ERR_BAD_SHAPE, "E (%d) must be in [1, %d]", s->E, MOEGAN_MAX_EXPERTS);
    MOEGAN_REQUIRE(s->precision == 0 || s->precision == 1, MOEGAN_ERR_UNSUPPORTED, "precision (%d) must be 0 (bf16) or 1 (fp32 3-pass)", s->precision);
    if (s->dense) {
        MOEGAN_REQUIRE(s->k == s->E && !s->hard, MOEGAN_ERR_BAD_SHAPE, "dense mode needs k == E (k=%d, E=%d) and hard == 0", s->k, s->E);
    } else {
        MOEGAN_REQUIRE(s->k >= 1 && s->k <= MOEGAN_MAX_TOPK && s->k < s->E + (s->E == 1), MOEGAN_ERR_BAD_SHAPE,
                       "sparse mode needs 1 <= k (%d) <= %d and k < E (%d); use dense mode for k == E", s->k, MOEGAN_MAX_TOPK, s->E);
        MOEGAN_REQUIRE(!s->hard || s->k == 1, MOEGAN_ERR_BAD_SHAPE, "hard (eval) routing is top-1 (k=%d)", s->k);
    }
    MOEGAN_REQUIRE(s->m_cap == moegan_m_cap(s->T, s->E, s->k, s->dense), MOEGAN_ERR_BAD_SHAPE,
                   "m_cap (%d) must equal moegan_m_cap() = %lld", s->m_cap, (long long)moegan_m_cap(s->T, s->E, s->k, s->dense));
    return MOEGAN_OK;
}

static int g_mn_lbo = 0, g_mn_sbo = 0, g_mn_kadv = 0;   // debug overrides (0 = kernel defaults)

static inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// split-K factor for a grouped weight-gradient GEMM whose output has too few tiles to fill the chip (small d, huge T)
static int wgrad_splits(const moegan_shape* s, int dimA, int N) {
    const int BN = N % 256 == 0 ? 256 : 128;
    const long long tiles = (long long)s->E * ceil_div(dimA, 128) * ceil_div(N, BN);
    if (tiles >= 96) return 1;       // >= 0.65 of a wave: the split's partial-sum traffic (ks x E x 8 d^2 x 4 B) costs more than the idle SMs
    const long long rows = s->dense ? s->T : ((long long)s->T * s->k + s->E - 1) / s->E;   // mean rows per expert
    long long ks = (2 * 148 + tiles - 1) / tiles;
    const long long max_ks = rows / 256 > 1 ? rows / 256 : 1;     // keep >= 4 k-blocks per split
    if (ks > max_ks) ks = max_ks;
    if (ks > 64) ks = 64;
    return (int)ks;
}
// scratch of moegan_ffn_bwd: [split-K partials of one weight gradient | 32-row slab sums of dpre | slab-sum split stage]
struct FfnBwdScratch { size_t splitk, dpre_slabs, slab_stage, total; };
static FfnBwdScratch ffn_bwd_scratch_layout(const moegan_shape* s) {
    const size_t d = s->d, hid = 4 * (size_t)s->d;
    const size_t k2 = wgrad_splits(s, s->d, 4 * s->d), k1 = wgrad_splits(s, 4 * s->d, s->d);
    const size_t a = k2 > 1 ? k2 * s->E * d * hid * 4 : 0, b = k1 > 1 ? k1 * s->E * d * hid * 4 : 0;
    FfnBwdScratch L;
    size_t o = 0;
    L.splitk = o;     o = align256(o + (a > b ? a : b));
    L.dpre_slabs = o; o = align256(o + (size_t)(s->m_cap / 32) * hid * 4);
    L.slab_stage = o; o = align256(o + (size_t)slab_sum_splits(s) * s->E * hid * 4);
    L.total = o + 256;
    return L;
}
static size_t ffn_bwd_scratch_bytes_impl(const moegan_shape* s) { return ffn_bwd_scratch_layout(s).total; }

struct RouterBwdWs {
    size_t draw, dfp_op, dtau_part, dtl, dwc_part, dWc, dtp, dtp_op, w_op, dWf, dWt, total;
};
static int dwf_splits(const moegan_shape* s) {   // split-K factor of the dWf = x^T dfp GEMM (few output tiles, long K = T)
    const int nkb = (s->T + 63) / 64;
    int ks = nkb / 4;
    if (ks < 1) ks = 1;
    if (ks > 64) ks = 64;
    return ks;
}
static RouterBwdWs router_bwd_layout(const moegan_shape* s) {
    const size_t S = s->precision == 1 ? 2 : 1;
    const size_t H = MOEGAN_ROUTER_HID;
    RouterBwdWs w;
    size_t o = 0;
    w.draw = o;      o = align256(o + (size_t)s->T * s->E * 4);
    w.dfp_op = o;    o = align256(o + (size_t)s->T * S * H * 2);
    w.dtau_part = o; o = align256(o + (size_t)router_bwd_logits_blocks(s) * 4 + 4);
    w.dtl = o;       o = align256(o + (size_t)s->B * s->E * 4);
    w.dwc_part = o;  o = align256(o + (size_t)router_bwd_dwc_parts(s) * H * s->E * 4 + 4);
    w.dWc = o;       o = align256(o + 2 * H * s->E * 4);
    w.dtp = o;       o = align256(o + (size_t)s->B * H * 4);
    w.dtp_op = o;    o = align256(o + (size_t)s->B * S * H * 2);
    w.w_op = o;      o = align256(o + (size_t)s->B * S * s->Dt * 2);
    w.dWf = o;       o = align256(o + (size_t)dwf_splits(s) * s->d * H * 4);
    w.dWt = o;       o = align256(o + (size_t)s->Dt * H * 4);
    w.total = o;
    return w;
}

}  // namespace moegan

using namespace moegan;

#define REQ_PTR(p) MOEGAN_REQUIRE((p) != nullptr, MOEGAN_ERR_NULL_POINTER, "%s: argument '%s' is NULL", __func__, #p)
#define CHECK_SHAPE(s)                                        \
    do {                                                      \
        int _rc = validate_shape(s);                          \
        if (_rc != MOEGAN_OK) return _rc;                     \
        profile_note_shape(s);                                \
    } while (0)
#define RUN(expr)                                             \
    do {                                                      \
        int _rc = (expr);                                     \
        if (_rc != MOEGAN_OK) return _rc;                     \
    } while (0)

extern "C" {

int moegan_version(void) { return MOEGAN_ABI_VERSION; }
const char* moegan_last_error_string(void) { return g_err; }

int64_t moegan_m_cap(int32_t T, int32_t E, int32_t k, int32_t dense) {
    const int64_t A = MOEGAN_ROW_ALIGN;
    if (dense) return (int64_t)E * (((int64_t)T + A - 1) / A * A);
    return (((int64_t)T * k + A - 1) / A * A) + A * E;
}
int64_t moegan_plan_ints(int32_t m_cap) { return PLAN_BLK + m_cap / 128 + 8; }
int64_t moegan_router_bwd_workspace_bytes(const moegan_shape* s) {
    if (s == nullptr) return 0;
    return (int64_t)router_bwd_layout(s).total;
}
int64_t moegan_ffn_bwd_scratch_bytes(const moegan_shape* s) {
    if (s == nullptr) return 0;
    return (int64_t)ffn_bwd_scratch_bytes_impl(s);
}
int64_t moegan_debug_launch_count(void) { return g_launches.load(); }
int64_t moegan_stream_capture_id(void* stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    if (cudaStreamGetCaptureInfo((cudaStream_t)stream, &st, &id) != cudaSuccess) { cudaGetLastError(); return -1; }
    return st == cudaStreamCaptureStatusActive ? (int64_t)id : 0;
}
int moegan_debug_profile_enable(int32_t level) {
    g_prof_level.store(level < 0 ? 0 : level);
    return MOEGAN_OK;
}
// gemm_only: drop (and free) the records of the other kernels - the contract of the original moegan_debug_profile_read
static int profile_drain(float* ms_out, int32_t* tag_out, int32_t* meta_out, int32_t max_n, bool gemm_only) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int n = 0;
    for (auto& r : g_prof) {
        const bool keep = !gemm_only || r.tag < MOEGAN_PROF_KERNEL_BASE;
        if (keep && n < max_n) {
            cudaEventSynchronize(r.end);
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, r.beg, r.end);
            if (ms_out) ms_out[n] = ms;
            if (tag_out) tag_out[n] = r.tag;
            if (meta_out) for (int i = 0; i < 8; ++i) meta_out[8 * n + i] = r.meta[i];
            ++n;
        }
        cudaEventDestroy(r.beg); cudaEventDestroy(r.end);
    }
    g_prof.clear();
    cudaGetLastError();
    return n;
}
int moegan_debug_profile_read(float* ms_out, int32_t* tag_out, int32_t max_n) { return profile_drain(ms_out, tag_out, nullptr, max_n, true); }
int moegan_debug_profile_read_ex(float* ms_out, int32_t* tag_out, int32_t* meta_out, int32_t max_n) {
    return profile_drain(ms_out, tag_out, meta_out, max_n, false);
}
int moegan_debug_set_mn_desc(int32_t lbo_bytes, int32_t sbo_bytes, int32_t kadv_bytes) {
    g_mn_lbo = lbo_bytes; g_mn_sbo = sbo_bytes; g_mn_kadv = kadv_bytes;
    return MOEGAN_OK;
}

// ---- router ----------------------------------------------------------------------------------------
int moegan_router_sample_kl_fwd(const moegan_shape* s, const float* mu_f, const float* rho_f, const float* eps_f,
                                const float* mu_t, const float* rho_t, const float* eps_t,
                                const float* mu_c, const float* rho_c, const float* eps_c,
                                int32_t sampling, float* Wf, float* Wt, float* Wc, void* Wf_op, void* Wt_op, float* kl_out, void* stream) {
    CHECK_SHAPE(s);
    REQ_PTR(mu_f); REQ_PTR(rho_f); REQ_PTR(mu_t); REQ_PTR(rho_t); REQ_PTR(mu_c); REQ_PTR(rho_c);
    REQ_PTR(Wf); REQ_PTR(Wt); REQ_PTR(Wc); REQ_PTR(kl_out);
    if (sampling) { REQ_PTR(eps_f); REQ_PTR(eps_t); REQ_PTR(eps_c); }
    return launch_sample_kl(s, mu_f, rho_f, eps_f, mu_t, rho_t, eps_t, mu_c, rho_c, eps_c, sampling, Wf, Wt, Wc, Wf_op, Wt_op, kl_out,
                            nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int moegan_router_text_fwd(const moegan_shape* s, const float* w, const float* Wt, const float* Wc, float* tp, float* tl, void* stream) {
    CHECK_SHAPE(s);
    if (s->B == 0) return MOEGAN_OK;
    REQ_PTR(w); REQ_PTR(Wt); REQ_PTR(Wc); REQ_PTR(tp); REQ_PTR(tl);
    return launch_text_fwd(s, w, Wt, Wc, tp, tl, (cudaStream_t)stream);
}

int moegan_router_fwd(const moegan_shape* s, const float* x, const float* Wf, const float* Wc, const float* tl,
                      const float* temperature, float annealing_factor, float* fp, float* raw, float* probs, int32_t* topk_idx,
                      float* gates, int32_t* chunk_counts, float* chunk_colsum, void* stream) {
    CHECK_SHAPE(s);
    if (s->T == 0) return MOEGAN_OK;
    REQ_PTR(x); REQ_PTR(Wf); REQ_PTR(Wc); REQ_PTR(tl); REQ_PTR(temperature); REQ_PTR(raw); REQ_PTR(probs);
    REQ_PTR(topk_idx); REQ_PTR(gates); REQ_PTR(chunk_counts); REQ_PTR(chunk_colsum);
    MOEGAN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wf) & 15) == 0, MOEGAN_ERR_MISALIGNED,
                   "router_fwd: x / Wf must be 16-byte aligned");
    return launch_router_fwd(s, x, Wf, Wc, tl, temperature, annealing_factor, fp, raw, probs, topk_idx, gates, chunk_counts,
                             chunk_colsum, (cudaStream_t)stream);
}

// ---- dispatch --------------------------------------------------------------------------------------
int moegan_dispatch_plan(const moegan_shape* s, const int32_t* chunk_counts, const float* chunk_colsum, int32_t* plan,
                         int32_t* chunk_base, float* load, void* stream) {
    CHECK_SHAPE(s);
    REQ_PTR(chunk_counts); REQ_PTR(chunk_colsum); REQ_PTR(plan); REQ_PTR(chunk_base); REQ_PTR(load);
    return launch_plan(s, chunk_counts, chunk_colsum, plan, chunk_base, load, (cudaStream_t)stream);
}

int moegan_dispatch_gather(const moegan_shape* s, const float* x, const int32_t* topk_idx, const float* gates, const int32_t* plan,
                           const int32_t* chunk_base, void* x_op, void* x_copy, int32_t* row_token, float* row_gate,
                           int32_t* token_rows, void* stream) {
    CHECK_SHAPE(s);
    MOEGAN_REQUIRE(!s->dense, MOEGAN_ERR_UNSUPPORTED, "dispatch_gather is for sparse mode; dense mode uses convert_tokens");
    REQ_PTR(x); REQ_PTR(topk_idx); REQ_PTR(gates); REQ_PTR(plan); REQ_PTR(chunk_base); REQ_PTR(x_op); REQ_PTR(row_token);
    REQ_PTR(row_gate); REQ_PTR(token_rows);
    return launch_gather(s, x, nullptr, topk_idx, gates, plan, chunk_base, x_op, x_copy, row_token, row_gate, token_rows, (cudaStream_t)stream);
}

int moegan_convert_tokens(const moegan_shape* s, const float* x, void* x_copy, void* stream) {
    CHECK_SHAPE(s);
    if (s->T == 0) return MOEGAN_OK;
    REQ_PTR(x); REQ_PTR(x_copy);
    return launch_to_operand(x, s->T, s->d, s->precision == 1, 0, x_copy, (cudaStream_t)stream);
}

int moegan_prep_weights(const moegan_shape* s, const float* const* w1, const float* const* b1, const float* const* w2,
                        const float* const* b2, void* W1_op, float* b1s, void* W2_op, float* b2s, void* stream) {
    CHECK_SHAPE(s);
    REQ_PTR(w1); REQ_PTR(b1); REQ_PTR(w2); REQ_PTR(b2); REQ_PTR(W1_op); REQ_PTR(b1s); REQ_PTR(W2_op); REQ_PTR(b2s);
    for (int e = 0; e < s->E; ++e) {
        MOEGAN_REQUIRE(w1[e] && b1[e] && w2[e] && b2[e], MOEGAN_ERR_NULL_POINTER, "prep_weights: NULL parameter pointer for expert %d", e);
        MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(w1[e]) | reinterpret_cast<uintptr_t>(w2[e])) & 15) == 0, MOEGAN_ERR_MISALIGNED,
                       "prep_weights: expert %d weights not 16-byte aligned", e);
    }
    return launch_prep_weights(s, w1, b1, w2, b2, W1_op, b1s, W2_op, b2s, (cudaStream_t)stream);
}

// ---- experts ---------------------------------------------------------------------------------------
static GemmParams base_params(const moegan_shape* s, const int32_t* plan) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.passes = s->precision == 1 ? 3 : 1;
    p.m_blocks_host = -1;
    p.k_rows_host = -1;
    p.plan = plan;
    p.E = s->E;
    p.out_rows_valid = s->m_cap;
    p.mn_lbo = g_mn_lbo; p.mn_sbo = g_mn_sbo; p.mn_kadv = g_mn_kadv;
    return p;
}

int moegan_ffn_fwd(const moegan_shape* s, const int32_t* plan, const void* a_op, const void* W1_op, const float* b1s,
                   const void* W2_op, const float* b2s, void* pre, void* h_op, void* y, void* stream) {
    CHECK_SHAPE(s);
    if (s->T == 0) return MOEGAN_OK;
    REQ_PTR(plan); REQ_PTR(a_op); REQ_PTR(W1_op); REQ_PTR(b1s); REQ_PTR(W2_op); REQ_PTR(b2s); REQ_PTR(h_op); REQ_PTR(y);
    // pre == NULL: forward-only call, nothing is saved for the backward pass
    const uint64_t S = s->precision == 1 ? 2 : 1;
    const int d = s->d, hid = 4 * s->d;
    const int T_pad = row_pad(s->T);
    const long long m_blocks = s->m_cap / 128;
    cudaStream_t st = (cudaStream_t)stream;
    {   // GEMM1: pre = x W1^T + b1, h = gelu(pre)   (t2i_moe_gan.py:257-258)
        GemmOperand A{a_op, S * d, (uint64_t)(s->dense ? s->T : s->m_cap), S * d};
        GemmOperand B{W1_op, S * d, (uint64_t)s->E * hid, S * d};
        GemmParams p = base_params(s, plan);
        p.N = hid; p.K = d; p.a_lo_off = d; p.b_lo_off = d;
        p.a_mod = s->dense ? T_pad / 128 : 0;
        p.b_rows_per_expert = hid;
        p.epi = s->precision == 1 ? EPI_GELU_X3 : EPI_GELU_BF16;
        p.out0 = h_op; p.ld_out0 = (long long)S * hid; p.out_split_off = hid;
        p.out1 = pre;  p.ld_out1 = hid;
        p.bias = b1s;
        RUN(launch_umma_gemm(GEMM_ROWS_BK, 256, A, B, p, m_blocks * (hid / 256), st));
    }
    {   // GEMM2: y = h W2^T + b2   (:259)
        GemmOperand A{h_op, S * hid, (uint64_t)s->m_cap, S * hid};
        GemmOperand B{W2_op, S * hid, (uint64_t)s->E * d, S * hid};
        GemmParams p = base_params(s, plan);
        const int BN = d % 256 == 0 ? 256 : 128;
        p.N = d; p.K = hid; p.a_lo_off = hid; p.b_lo_off = hid;
        p.b_rows_per_expert = d;
        p.epi = s->precision == 1 ? EPI_BIAS_F32 : EPI_BIAS_BF16;
        p.out0 = y; p.ld_out0 = d;
        p.bias = b2s;
        RUN(launch_umma_gemm(GEMM_ROWS_BK, BN, A, B, p, m_blocks * (d / BN), st));
    }
    return MOEGAN_OK;
}

int moegan_combine_fwd(const moegan_shape* s, const void* y, const float* gates, const int32_t* token_rows, const void* residual,
                       int32_t residual_is_bf16, void* out, int32_t out_is_bf16, void* stream) {
    CHECK_SHAPE(s);
    if (s->T == 0) return MOEGAN_OK;
    REQ_PTR(y); REQ_PTR(gates); REQ_PTR(out);
    if (!s->dense) REQ_PTR(token_rows);
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0, MOEGAN_ERR_MISALIGNED,
                   "combine_fwd: out / residual must be 16-byte aligned");
    return launch_combine_fwd(s, y, gates, token_rows, residual, residual_is_bf16, out, out_is_bf16, (cudaStream_t)stream);
}

// ---- backward --------------------------------------------------------------------------------------
int moegan_combine_bwd(const moegan_shape* s, const int32_t* plan, const void* dout, int32_t dout_is_bf16, const void* y,
                       const int32_t* row_token, const float* row_gate, const float* gates, void* dy_op, float* dgate_row,
                       float* dy_slab_sums, void* stream) {
    CHECK_SHAPE(s);
    if (s->T == 0) return MOEGAN_OK;
    REQ_PTR(plan); REQ_PTR(dout); REQ_PTR(y); REQ_PTR(dy_op); REQ_PTR(dgate_row);
    if (s->dense) { REQ_PTR(gates); } else { REQ_PTR(row_token); REQ_PTR(row_gate); }
    return launch_combine_bwd(s, plan, dout, dout_is_bf16, y, row_token, row_gate, gates, dy_op, dgate_row, dy_slab_sums, (cudaStream_t)stream);
}

// defer_bias: the caller (moegan_block_bwd) runs the two bias-gradient slab sums itself on its side stream, off the critical
// path of the GEMM chain; dgrad2_done is recorded on `stream` right after the dgrad2 launch (the db1 slab sums depend on it).
static int ffn_bwd_core(const moegan_shape* s, const int32_t* plan, const void* a_op, const void* W1_op, const void* W2_op,
                        const void* pre, const void* h_op, const void* dy_op, const float* dy_slab_sums, void* dpre_op, void* dx_rows,
                        float* dW1, float* db1, float* dW2, float* db2, void* scratch, int64_t scratch_bytes, void* stream,
                        bool defer_bias, cudaEvent_t dgrad2_done) {
    CHECK_SHAPE(s);
    const FfnBwdScratch SL = ffn_bwd_scratch_layout(s);
    MOEGAN_REQUIRE(scratch_bytes >= (int64_t)SL.total && scratch != nullptr,
                   MOEGAN_ERR_BAD_SHAPE, "ffn_bwd: scratch %lld < required %lld bytes", (long long)scratch_bytes, (long long)SL.total);
    MOEGAN_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, MOEGAN_ERR_MISALIGNED, "ffn_bwd: scratch must be 256-byte aligned");
    void* splitk = reinterpret_cast<uint8_t*>(scratch) + SL.splitk;
    float* dpre_slabs = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + SL.dpre_slabs);
    float* slab_stage = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + SL.slab_stage);
    REQ_PTR(plan); REQ_PTR(a_op); REQ_PTR(W1_op); REQ_PTR(W2_op); REQ_PTR(pre); REQ_PTR(h_op); REQ_PTR(dy_op); REQ_PTR(dpre_op);
    REQ_PTR(dx_rows); REQ_PTR(dW1); REQ_PTR(db1); REQ_PTR(dW2); REQ_PTR(db2);
    const uint64_t S = s->precision == 1 ? 2 : 1;
    const int d = s->d, hid = 4 * s->d;
    const long long m_blocks = s->m_cap / 128;
    const int BNd = d % 256 == 0 ? 256 : 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (s->T == 0) {
        MOEGAN_CHECK_CUDA(cudaMemsetAsync(dW1, 0, (size_t)s->E * hid * d * 4, st));
        MOEGAN_CHECK_CUDA(cudaMemsetAsync(dW2, 0, (size_t)s->E * hid * d * 4, st));
        MOEGAN_CHECK_CUDA(cudaMemsetAsync(db1, 0, (size_t)s->E * hid * 4, st));
        MOEGAN_CHECK_CUDA(cudaMemsetAsync(db2, 0, (size_t)s->E * d * 4, st));
        return MOEGAN_OK;
    }
    {   // dgrad2 + GELU':  dpre = (dy W2) * gelu'(pre)
        GemmOperand A{dy_op, S * d, (uint64_t)s->m_cap, S * d};
        GemmOperand B{W2_op, S * hid, (uint64_t)s->E * d, S * hid};
        GemmParams p = base_params(s, plan);
        p.N = hid; p.K = d; p.a_lo_off = d; p.b_lo_off = hid;
        p.b_rows_per_expert = d;
        p.epi = s->precision == 1 ? EPI_DGELU_X3 : EPI_MUL_BF16;
        p.aux = pre; p.ld_aux = hid;
        p.out0 = dpre_op; p.ld_out0 = (long long)S * hid; p.out_split_off = hid;
        p.colsum_part = dpre_slabs;                       // db1 falls out of this epilogue
        RUN(launch_umma_gemm(GEMM_ROWS_BMN, 256, A, B, p, m_blocks * (hid / 256), st));
    }
    // bias gradients = per-expert sums of the 32-row slab sums (dy: from combine_bwd; dpre: from the epilogue above)
    if (defer_bias) {
        MOEGAN_CHECK_CUDA(cudaEventRecord(dgrad2_done, st));
    } else {
        RUN(launch_slab_sum(s, plan, dpre_slabs, hid, slab_stage, db1, st));
        if (dy_slab_sums != nullptr) RUN(launch_slab_sum(s, plan, dy_slab_sums, d, slab_stage, db2, st));
        else                         RUN(launch_colsum_grouped(s, plan, dy_op, d, dW2, db2, st));   // partials in dW2 (written further down)
    }
    {   // dgrad1: dx_rows = dpre W1
        GemmOperand A{dpre_op, S * hid, (uint64_t)s->m_cap, S * hid};
        GemmOperand B{W1_op, S * d, (uint64_t)s->E * hid, S * d};
        GemmParams p = base_params(s, plan);
        p.N = d; p.K = hid; p.a_lo_off = hid; p.b_lo_off = d;
        p.b_rows_per_expert = hid;
        p.epi = s->precision == 1 ? EPI_F32 : EPI_BF16;
        p.out0 = dx_rows; p.ld_out0 = d;
        RUN(launch_umma_gemm(GEMM_ROWS_BMN, BNd, A, B, p, m_blocks * (d / BNd), st));
    }
    {   // wgrad2: dW2[e] = dy_e^T h_e   [d, 4d]
        GemmOperand A{dy_op, S * d, (uint64_t)s->m_cap, S * d};
        GemmOperand B{h_op, S * hid, (uint64_t)s->m_cap, S * hid};
        GemmParams p = base_params(s, plan);
        p.N = hid; p.dimA = d; p.a_lo_off = d; p.b_lo_off = hid;
        const int ks = wgrad_splits(s, d, hid);
        p.k_splits = ks;
        p.epi = EPI_F32; p.out0 = ks > 1 ? splitk : (void*)dW2; p.ld_out0 = hid;
        RUN(launch_umma_gemm(GEMM_WGRAD, 256, A, B, p, (long long)s->E * (d / 128) * (hid / 256) * ks, st));
        if (ks > 1) RUN(launch_sum_partials(reinterpret_cast<const float*>(splitk), (long long)s->E * d * hid, ks, dW2, st));
    }
    {   // wgrad1: dW1[e] = dpre_e^T x_e   [4d, d]
        GemmOperand A{dpre_op, S * hid, (uint64_t)s->m_cap, S * hid};
        GemmOperand B{a_op, S * d, (uint64_t)(s->dense ? s->T : s->m_cap), S * d};
        GemmParams p = base_params(s, plan);
        p.N = d; p.dimA = hid; p.a_lo_off = hid; p.b_lo_off = d;
        p.b_rel = s->dense ? 1 : 0;
        const int ks = wgrad_splits(s, hid, d);
        p.k_splits = ks;
        p.epi = EPI_F32; p.out0 = ks > 1 ? splitk : (void*)dW1; p.ld_out0 = d;
        RUN(launch_umma_gemm(GEMM_WGRAD, BNd, A, B, p, (long long)s->E * (hid / 128) * (d / BNd) * ks, st));
        if (ks > 1) RUN(launch_sum_partials(reinterpret_cast<const float*>(splitk), (long long)s->E * d * hid, ks, dW1, st));
    }
    return MOEGAN_OK;
}

int moegan_ffn_bwd(const moegan_shape* s, const int32_t* plan, const void* a_op, const void* W1_op, const void* W2_op,
                   const void* pre, const void* h_op, const void* dy_op, const float* dy_slab_sums, void* dpre_op, void* dx_rows,
                   float* dW1, float* db1, float* dW2, float* db2, void* scratch, int64_t scratch_bytes, void* stream) {
    return ffn_bwd_core(s, plan, a_op, W1_op, W2_op, pre, h_op, dy_op, dy_slab_sums, dpre_op, dx_rows, dW1, db1, dW2, db2, scratch,
                        scratch_bytes, stream, false, nullptr);
}

int moegan_uncombine_bwd(const moegan_shape* s, const void* dx_rows, const int32_t* token_rows, const float* dx_router, float* dx,
                         void* stream) {
    CHECK_SHAPE(s);
    if (s->T == 0) return MOEGAN_OK;
    REQ_PTR(dx_rows); REQ_PTR(dx);
    if (!s->dense) REQ_PTR(token_rows);
    return launch_uncombine(s, dx_rows, token_rows, dx_router, dx, (cudaStream_t)stream);
}

// dx_router_done (optional): recorded on `stream` right after the GEMM that produces dx_router - the un-permute of the block's
// backward waits for that, not for the rest of the chain
static int router_bwd_core(const moegan_shape* s, const void* x_copy, const float* w, const float* Wf, const float* Wt, const float* Wc,
                      const void* Wf_op, const void* Wt_op, const float* fp, const float* tp, const float* raw,
                      const int32_t* topk_idx, const int32_t* token_rows, const float* dgate_row, const float* dprobs_ext,
                      const float* dload, const float* temperature, float annealing_factor, int32_t sampled,
                      const float* mu_f, const float* rho_f, const float* eps_f, const float* mu_t, const float* rho_t,
                      const float* eps_t, const float* mu_c, const float* rho_c, const float* eps_c,
                      const float* kl_out, const float* dkl,
                      float* dmu_f, float* drho_f, float* dmu_t, float* drho_t, float* dmu_c, float* drho_c,
                      float* dtemperature, float* dx_router, float* dw, void* workspace, int64_t workspace_bytes, void* stream,
                           cudaEvent_t dx_router_done) {
    CHECK_SHAPE(s);
    MOEGAN_REQUIRE(!s->hard, MOEGAN_ERR_UNSUPPORTED, "router_bwd: hard (eval) routing has no gradient");
    MOEGAN_REQUIRE(s->T > 0, MOEGAN_ERR_BAD_SHAPE, "router_bwd: T must be positive");
    REQ_PTR(x_copy); REQ_PTR(w); REQ_PTR(Wf); REQ_PTR(Wt); REQ_PTR(Wc); REQ_PTR(Wf_op); REQ_PTR(Wt_op); REQ_PTR(fp); REQ_PTR(tp);
    REQ_PTR(raw); REQ_PTR(dgate_row); REQ_PTR(temperature);
    REQ_PTR(mu_f); REQ_PTR(rho_f); REQ_PTR(mu_t); REQ_PTR(rho_t); REQ_PTR(mu_c); REQ_PTR(rho_c);
    REQ_PTR(dmu_f); REQ_PTR(drho_f); REQ_PTR(dmu_t); REQ_PTR(drho_t); REQ_PTR(dmu_c); REQ_PTR(drho_c);
    REQ_PTR(dtemperature); REQ_PTR(dx_router); REQ_PTR(dw); REQ_PTR(workspace);
    if (!s->dense) { REQ_PTR(topk_idx); REQ_PTR(token_rows); }
    if (sampled) { REQ_PTR(eps_f); REQ_PTR(eps_t); REQ_PTR(eps_c); }
    const RouterBwdWs L = router_bwd_layout(s);
    MOEGAN_REQUIRE(workspace_bytes >= (int64_t)L.total, MOEGAN_ERR_BAD_SHAPE, "router_bwd: workspace %lld < required %lld bytes",
                   (long long)workspace_bytes, (long long)L.total);
    MOEGAN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, MOEGAN_ERR_MISALIGNED, "router_bwd: workspace must be 256-byte aligned");
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    float* draw = reinterpret_cast<float*>(ws + L.draw);
    void* dfp_op = ws + L.dfp_op;
    float* dtau_part = reinterpret_cast<float*>(ws + L.dtau_part);
    float* dtl = reinterpret_cast<float*>(ws + L.dtl);
    float* dwc_part = reinterpret_cast<float*>(ws + L.dwc_part);
    float* dWc = reinterpret_cast<float*>(ws + L.dWc);
    float* dtp = reinterpret_cast<float*>(ws + L.dtp);
    void* dtp_op = ws + L.dtp_op;
    void* w_op = ws + L.w_op;
    float* dWf = reinterpret_cast<float*>(ws + L.dWf);
    float* dWt = reinterpret_cast<float*>(ws + L.dWt);
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t S = s->precision == 1 ? 2 : 1;
    const uint64_t H = MOEGAN_ROUTER_HID;
    const int d = s->d, Dt = s->Dt;

    RUN(launch_router_bwd_logits(s, raw, Wc, topk_idx, token_rows, dgate_row, dprobs_ext, dload, temperature, annealing_factor,
                                 fp, draw, dfp_op, dtau_part, dwc_part, st));
    RUN(launch_router_bwd_small(s, fp, draw, tp, Wc, dtau_part, temperature, annealing_factor, dtl, dwc_part, dWc, dtp, dtp_op,
                                dtemperature, st));
    const bool text_simt = router_bwd_text_simt_ok(s);     // tiny batches: one SIMT kernel instead of two tensor-core GEMMs
    if (!text_simt) RUN(launch_to_operand(w, s->B, Dt, s->precision == 1, 1, w_op, st));

    GemmParams base;
    memset(&base, 0, sizeof(base));
    base.passes = s->precision == 1 ? 3 : 1;
    base.m_blocks_host = -1; base.k_rows_host = -1; base.E = 1; base.epi = EPI_F32;
    base.mn_lbo = g_mn_lbo; base.mn_sbo = g_mn_sbo; base.mn_kadv = g_mn_kadv;
    {   // dx_router [T, d] = dfp [T,128] * Wf[d,128]^T
        GemmOperand A{dfp_op, S * H, (uint64_t)s->T, S * H};
        GemmOperand B{Wf_op, S * H, (uint64_t)d, S * H};
        GemmParams p = base;
        const int BN = d % 256 == 0 ? 256 : 128;
        p.N = d; p.K = (int)H; p.a_lo_off = (int)H; p.b_lo_off = (int)H; p.b_rows_per_expert = d;
        p.m_blocks_host = ceil_div(s->T, 128); p.out_rows_valid = s->T;
        p.out0 = dx_router; p.ld_out0 = d;
        RUN(launch_umma_gemm(GEMM_ROWS_BK, BN, A, B, p, (long long)p.m_blocks_host * (d / BN), st));
        if (dx_router_done != nullptr) MOEGAN_CHECK_CUDA(cudaEventRecord(dx_router_done, st));
    }
    {   // dWf [d, 128] = x^T dfp
        GemmOperand A{x_copy, S * d, (uint64_t)s->T, S * d};
        GemmOperand B{dfp_op, S * H, (uint64_t)s->T, S * H};
        GemmParams p = base;
        p.N = (int)H; p.dimA = d; p.a_lo_off = d; p.b_lo_off = (int)H; p.k_rows_host = s->T;
        p.k_splits = dwf_splits(s);                       // partials summed by reparam_kl_bwd_kernel
        p.out0 = dWf; p.ld_out0 = (long long)H;
        RUN(launch_umma_gemm(GEMM_WGRAD, 128, A, B, p, (long long)(d / 128) * p.k_splits, st));
    }
    if (text_simt) {
        RUN(launch_router_bwd_text(s, w, dtp, Wt, dw, dWt, st));
    } else {
    {   // dw [B, Dt] = dtp [B,128] * Wt[Dt,128]^T
        GemmOperand A{dtp_op, S * H, (uint64_t)s->B, S * H};
        GemmOperand B{Wt_op, S * H, (uint64_t)Dt, S * H};
        GemmParams p = base;
        p.N = Dt; p.K = (int)H; p.a_lo_off = (int)H; p.b_lo_off = (int)H; p.b_rows_per_expert = Dt;
        p.m_blocks_host = ceil_div(s->B, 128); p.out_rows_valid = s->B;
        p.out0 = dw; p.ld_out0 = Dt;
        RUN(launch_umma_gemm(GEMM_ROWS_BK, 128, A, B, p, (long long)p.m_blocks_host * ceil_div(Dt, 128), st));
    }
    {   // dWt [Dt, 128] = w^T dtp
        GemmOperand A{w_op, S * Dt, (uint64_t)s->B, S * Dt};
        GemmOperand B{dtp_op, S * H, (uint64_t)s->B, S * H};
        GemmParams p = base;
        p.N = (int)H; p.dimA = Dt; p.a_lo_off = Dt; p.b_lo_off = (int)H; p.k_rows_host = s->B;
        p.out0 = dWt; p.ld_out0 = (long long)H;
        RUN(launch_umma_gemm(GEMM_WGRAD, 128, A, B, p, (long long)ceil_div(Dt, 128), st));
    }
    }
    RUN(launch_reparam_kl_bwd(s, dWf, dWt, dWc, mu_f, rho_f, eps_f, mu_t, rho_t, eps_t, mu_c, rho_c, eps_c, sampled, dwf_splits(s), kl_out, dkl,
                              dmu_f, drho_f, dmu_t, drho_t, dmu_c, drho_c, st));
    return MOEGAN_OK;
}

int moegan_router_bwd(const moegan_shape* s, const void* x_copy, const float* w, const float* Wf, const float* Wt, const float* Wc,
                      const void* Wf_op, const void* Wt_op, const float* fp, const float* tp, const float* raw,
                      const int32_t* topk_idx, const int32_t* token_rows, const float* dgate_row, const float* dprobs_ext,
                      const float* dload, const float* temperature, float annealing_factor, int32_t sampled,
                      const float* mu_f, const float* rho_f, const float* eps_f, const float* mu_t, const float* rho_t,
                      const float* eps_t, const float* mu_c, const float* rho_c, const float* eps_c,
                      const float* kl_out, const float* dkl,
                      float* dmu_f, float* drho_f, float* dmu_t, float* drho_t, float* dmu_c, float* drho_c,
                      float* dtemperature, float* dx_router, float* dw, void* workspace, int64_t workspace_bytes, void* stream) {
    return router_bwd_core(s, x_copy, w, Wf, Wt, Wc, Wf_op, Wt_op, fp, tp, raw, topk_idx, token_rows, dgate_row, dprobs_ext, dload, temperature,
                           annealing_factor, sampled, mu_f, rho_f, eps_f, mu_t, rho_t, eps_t, mu_c, rho_c, eps_c, kl_out, dkl,
                           dmu_f, drho_f, dmu_t, drho_t, dmu_c, drho_c, dtemperature, dx_router, dw, workspace, workspace_bytes, stream, nullptr);
}

int moegan_balance_loss_fwd(const float* load, int32_t E, int32_t T, float balance_weight, float* loss, void* stream) {
    REQ_PTR(load); REQ_PTR(loss);
    MOEGAN_REQUIRE(E >= 1 && E <= MOEGAN_MAX_EXPERTS && T > 0, MOEGAN_ERR_BAD_SHAPE, "balance_loss: bad E (%d) or T (%d)", E, T);
    return launch_balance_loss(load, E, T, balance_weight, nullptr, loss, nullptr, (cudaStream_t)stream);
}
int moegan_balance_loss_bwd(const float* load, int32_t E, int32_t T, float balance_weight, const float* dloss, float* dload, void* stream) {
    REQ_PTR(load); REQ_PTR(dloss); REQ_PTR(dload);
    MOEGAN_REQUIRE(E >= 1 && E <= MOEGAN_MAX_EXPERTS && T > 0, MOEGAN_ERR_BAD_SHAPE, "balance_loss: bad E (%d) or T (%d)", E, T);
    return launch_balance_loss(load, E, T, balance_weight, dloss, nullptr, dload, (cudaStream_t)stream);
}

// ---- side stream: independent sub-chains of a block call run concurrently with the expert GEMMs ----------
// (the router's backward chain / the weight conversion are many small launches that leave most SMs idle).  Fork and join
// are cudaEventRecord / cudaStreamWaitEvent pairs, so the pattern is legal inside CUDA-graph capture too.
struct SideStream { cudaStream_t s = nullptr, s2 = nullptr; cudaEvent_t fork = nullptr, join = nullptr, mid = nullptr, join2 = nullptr, dxr = nullptr;
                    bool ok = false, tried = false; };
// one set per device (a process that drives several GPUs gets streams / events on each of them, created lazily on the device that is
// current when the block call arrives - the caller's stream belongs to that device)
constexpr int MOEGAN_MAX_DEVICES = 64;
static SideStream g_side_dev[MOEGAN_MAX_DEVICES];
static std::mutex g_side_mu;
static std::atomic<int> g_side_override{-1};     // moegan_debug_set_side_streams: -1 = environment / default
static bool side_enabled() {
    const int o = g_side_override.load(std::memory_order_relaxed);
    if (o >= 0) return o == 1;
    static int v = -1;
    if (v < 0) { const char* e = getenv("MOEGAN_DISABLE_SIDE_STREAM"); v = (e != nullptr && e[0] == '1') ? 0 : 1; }
    return v == 1;
}
int moegan_debug_set_side_streams(int32_t mode) {
    g_side_override.store(mode < 0 ? -1 : (mode != 0 ? 1 : 0));
    return MOEGAN_OK;
}
// Overlap only pays while the expert GEMMs leave SMs idle (configs[1]: 8 192 dispatched rows, 353 -> 323 us).  At the train-step
// shapes (16 K - 262 K rows) every GEMM fills the chip with one 225 KB CTA per SM, nothing can co-reside, and the fork / join
// edges plus the contention cost 0.7 ms per step (measured) - there the block runs on the caller's stream alone.
static bool side_pays(const moegan_shape* s) {
    static int force = -1;
    if (force < 0) { const char* e = getenv("MOEGAN_FORCE_SIDE_STREAM"); force = (e != nullptr && e[0] == '1') ? 1 : 0; }
    if (force == 1) return true;
    const long long rows = s->dense ? (long long)s->T * s->E : (long long)s->T * s->k;
    return rows * s->d <= 6LL * 1024 * 1024;
}
static SideStream* side_stream() {
    if (!side_enabled()) return nullptr;
    std::lock_guard<std::mutex> lk(g_side_mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MOEGAN_MAX_DEVICES) { cudaGetLastError(); return nullptr; }
    SideStream& g_side = g_side_dev[dev];
    if (!g_side.tried) {
        g_side.tried = true;
        g_side.ok = cudaStreamCreateWithFlags(&g_side.s, cudaStreamNonBlocking) == cudaSuccess &&
                    cudaStreamCreateWithFlags(&g_side.s2, cudaStreamNonBlocking) == cudaSuccess &&
                    cudaEventCreateWithFlags(&g_side.join2, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&g_side.dxr, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&g_side.fork, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&g_side.join, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&g_side.mid, cudaEventDisableTiming) == cudaSuccess;
        if (!g_side.ok) cudaGetLastError();
    }
    return g_side.ok ? &g_side : nullptr;
}
static int side_fork(SideStream* ss, cudaStream_t main_st) {
    MOEGAN_CHECK_CUDA(cudaEventRecord(ss->fork, main_st));
    MOEGAN_CHECK_CUDA(cudaStreamWaitEvent(ss->s, ss->fork, 0));
    return MOEGAN_OK;
}
static int side_join(SideStream* ss, cudaStream_t main_st) {
    MOEGAN_CHECK_CUDA(cudaEventRecord(ss->join, ss->s));
    MOEGAN_CHECK_CUDA(cudaStreamWaitEvent(main_st, ss->join, 0));
    return MOEGAN_OK;
}

// ---- whole block in one call -------------------------------------------------------------------------
// Forward-only calls (the D-step's generator pass under no_grad, eval / inference) in bf16 mode at d = 128 / 256 run the fused
// FFN kernel (ffn_fused.cu): dense-soft mode also folds the weighted combine + residual into it, so neither H, y nor a combine
// launch exist; grouped (sparse / hard) mode writes y and keeps the un-permuting combine.
static std::atomic<int> g_fused_override{-1};      // moegan_debug_set_fused_ffn: -1 = default (on where supported)
static bool fused_fwd(const moegan_shape* s, int need_bwd) {
    if (need_bwd) return false;
    const int o = g_fused_override.load(std::memory_order_relaxed);
    if (o == 0) return false;
    return ffn_fused_supported(s->d, s->precision);
}
static void block_layout(const moegan_shape* s, int need_bwd, moegan_block_layout* L) {
    const size_t S = s->precision == 1 ? 2 : 1, H = MOEGAN_ROUTER_HID;
    const size_t T = s->T, d = s->d, Dt = s->Dt, E = s->E, k = s->k, B = s->B, hid = 4 * d, m_cap = s->m_cap;
    const size_t nch = (T + MOEGAN_CHUNK - 1) / MOEGAN_CHUNK > 0 ? (T + MOEGAN_CHUNK - 1) / MOEGAN_CHUNK : 1;
    const size_t act = s->precision == 1 ? 4 : 2;
    size_t o = 0;
    auto take = [&](size_t bytes, bool present) -> int64_t {
        if (!present) return -1;
        const size_t at = o;
        o = align256(o + (bytes > 0 ? bytes : 1));
        return (int64_t)at;
    };
    L->Wf = take(d * H * 4, true);
    L->Wt = take(Dt * H * 4, true);
    L->Wc = take(2 * H * E * 4, true);
    L->Wf_op = take(d * S * H * 2, need_bwd);
    L->Wt_op = take(Dt * S * H * 2, need_bwd);
    L->tp = take(B * H * 4, true);
    L->tl = take(B * E * 4, true);
    L->fp = take(T * H * 4, need_bwd);
    L->raw = take(T * E * 4, true);
    L->topk_idx = take(T * k * 4, true);
    L->gates = take(T * k * 4, true);
    L->chunk_counts = take(nch * E * 4, true);
    L->chunk_colsum = take(nch * E * 4, true);
    L->plan = take((size_t)moegan_plan_ints(s->m_cap) * 4, true);
    L->chunk_base = take(nch * E * 4, true);
    L->x_copy = take(T * S * d * 2, need_bwd || s->dense);
    L->a_op = s->dense ? L->x_copy : take(m_cap * S * d * 2, true);
    L->row_token = take(m_cap * 4, !s->dense);
    L->row_gate = take(m_cap * 4, !s->dense);
    L->token_rows = take(T * k * 4, !s->dense);
    const bool fused = fused_fwd(s, need_bwd);
    L->pre = take(m_cap * hid * act, need_bwd);
    L->h_op = take(m_cap * S * hid * 2, !fused);
    L->y = take(m_cap * d * act, !(fused && s->dense));
    // the eps the forward consumed (the caller's buffers are redrawn by the next training forward)
    L->eps_f = take(d * H * 4, need_bwd);
    L->eps_t = take(Dt * H * 4, need_bwd);
    L->eps_c = take(2 * H * E * 4, need_bwd);
    L->total = (int64_t)o;
}

struct BlockBwdWs { size_t dy_op, dgate_row, dy_slabs, dpre_op, dx_rows, dx_router, router_ws, ffn_scratch, total; };
static BlockBwdWs block_bwd_layout(const moegan_shape* s) {
    const size_t S = s->precision == 1 ? 2 : 1, d = s->d, hid = 4 * (size_t)s->d, m_cap = s->m_cap;
    const size_t act = s->precision == 1 ? 4 : 2;
    BlockBwdWs w;
    size_t o = 0;
    w.dy_op = o;     o = align256(o + m_cap * S * d * 2);
    w.dgate_row = o; o = align256(o + m_cap * 4);
    w.dy_slabs = o;  o = align256(o + (m_cap / 32) * d * 4);
    w.dpre_op = o;   o = align256(o + m_cap * S * hid * 2);
    w.dx_rows = o;   o = align256(o + m_cap * d * act);
    w.dx_router = o; o = align256(o + (size_t)s->T * d * 4);
    w.router_ws = o; o = align256(o + router_bwd_layout(s).total);
    w.ffn_scratch = o; o = align256(o + ffn_bwd_scratch_bytes_impl(s));
    w.total = o;
    return w;
}

int64_t moegan_block_fwd_workspace(const moegan_shape* s, int32_t need_backward, moegan_block_layout* layout) {
    if (validate_shape(s) != MOEGAN_OK) return -1;
    moegan_block_layout L;
    block_layout(s, need_backward, &L);
    if (layout != nullptr) *layout = L;
    return L.total;
}
int64_t moegan_block_bwd_workspace(const moegan_shape* s) {
    if (validate_shape(s) != MOEGAN_OK) return -1;
    return (int64_t)block_bwd_layout(s).total;
}

int moegan_block_fwd(const moegan_shape* s, const moegan_block_fwd_args* a, void* stream) {
    CHECK_SHAPE(s);
    REQ_PTR(a);
    REQ_PTR(a->x); REQ_PTR(a->w); REQ_PTR(a->out); REQ_PTR(a->probs); REQ_PTR(a->load); REQ_PTR(a->kl_out); REQ_PTR(a->workspace);
    moegan_block_layout L;
    block_layout(s, a->need_backward, &L);
    MOEGAN_REQUIRE(a->workspace_bytes >= L.total, MOEGAN_ERR_BAD_SHAPE, "block_fwd: workspace %lld < required %lld bytes",
                   (long long)a->workspace_bytes, (long long)L.total);
    MOEGAN_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, MOEGAN_ERR_MISALIGNED, "block_fwd: workspace must be 256-byte aligned");
    uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
    auto P = [&](int64_t off) -> void* { return off < 0 ? nullptr : ws + off; };
    auto F = [&](int64_t off) -> float* { return reinterpret_cast<float*>(P(off)); };
    auto I = [&](int64_t off) -> int32_t* { return reinterpret_cast<int32_t*>(P(off)); };
    if (s->T == 0) return MOEGAN_OK;
    // Weight conversion (only when the caller's cache is stale).  On small blocks it runs on the side stream, forked AFTER the router
    // kernel has been enqueued: the router's CTAs (one per SM at T = 4096) are dispatched first and the conversion's blocks co-reside with
    // them.  Forked at the top of the call (as in round 1) its many blocks were dispatched ahead of the single-block kl_finalize and the
    // 8-block text kernel of the main chain, which then sat behind it for ~10 us (profiles/r02o_timeline_moe_block.log).
    SideStream* ss = (a->prep_w1 != nullptr && side_pays(s)) ? side_stream() : nullptr;
    auto prep = [&](void* st_) -> int {
        return moegan_prep_weights(s, a->prep_w1, a->prep_b1, a->prep_w2, a->prep_b2, const_cast<void*>(a->W1_op), const_cast<float*>(a->b1s),
                                   const_cast<void*>(a->W2_op), const_cast<float*>(a->b2s), st_);
    };
    if (a->prep_w1 != nullptr && ss == nullptr) RUN(prep(stream));
    REQ_PTR(a->mu_f); REQ_PTR(a->rho_f); REQ_PTR(a->mu_t); REQ_PTR(a->rho_t); REQ_PTR(a->mu_c); REQ_PTR(a->rho_c);
    if (a->sampled) { REQ_PTR(a->eps_f); REQ_PTR(a->eps_t); REQ_PTR(a->eps_c); }
    const bool keep_eps = a->need_backward && a->sampled;      // the eps consumed here, kept inside the workspace for the backward
    RUN(launch_sample_kl(s, a->mu_f, a->rho_f, a->eps_f, a->mu_t, a->rho_t, a->eps_t, a->mu_c, a->rho_c, a->eps_c, a->sampled,
                         F(L.Wf), F(L.Wt), F(L.Wc), P(L.Wf_op), P(L.Wt_op), a->kl_out,
                         keep_eps ? F(L.eps_f) : nullptr, keep_eps ? F(L.eps_t) : nullptr, keep_eps ? F(L.eps_c) : nullptr, (cudaStream_t)stream));
    RUN(moegan_router_text_fwd(s, a->w, F(L.Wt), F(L.Wc), F(L.tp), F(L.tl), stream));
    // the fork EVENT is recorded before the router kernel (so the conversion depends only on what precedes it), the conversion is
    // ENQUEUED after it (so the router's CTAs, one per SM, are dispatched first and the conversion's blocks fill in beside them)
    if (ss != nullptr) MOEGAN_CHECK_CUDA(cudaEventRecord(ss->fork, (cudaStream_t)stream));
    RUN(moegan_router_fwd(s, a->x, F(L.Wf), F(L.Wc), F(L.tl), a->temperature, a->annealing_factor, F(L.fp), F(L.raw), a->probs,
                          I(L.topk_idx), F(L.gates), I(L.chunk_counts), F(L.chunk_colsum), stream));
    if (ss != nullptr) {
        MOEGAN_CHECK_CUDA(cudaStreamWaitEvent(ss->s, ss->fork, 0));
        RUN(prep(ss->s));
    }
    RUN(moegan_dispatch_plan(s, I(L.chunk_counts), F(L.chunk_colsum), I(L.plan), I(L.chunk_base), a->load, stream));
    // a->x_op: the caller already holds the bf16 operand copy of x (emitted by moegan_layernorm_fwd_dual in the same pass as the fp32
    // rows): no conversion kernel here, and the backward pass reads that tensor too (moegan_block_bwd_args.x_op)
    MOEGAN_REQUIRE(a->x_op == nullptr || (s->precision == 0 && (reinterpret_cast<uintptr_t>(a->x_op) & 15) == 0), MOEGAN_ERR_UNSUPPORTED,
                   "block_fwd: x_op is a plain bf16 [T, d] copy of x - bf16 mode only, 16-byte aligned");
    void* const x_copy = a->x_op != nullptr ? const_cast<void*>(a->x_op) : P(L.x_copy);
    void* const a_op = s->dense ? x_copy : P(L.a_op);
    if (s->dense) {
        if (a->x_op == nullptr) RUN(moegan_convert_tokens(s, a->x, x_copy, stream));
    } else {
        RUN(launch_gather(s, a->x, a->x_op, I(L.topk_idx), F(L.gates), I(L.plan), I(L.chunk_base), a_op, a->x_op != nullptr ? nullptr : x_copy,
                          I(L.row_token), F(L.row_gate), I(L.token_rows), (cudaStream_t)stream));
    }
    if (ss != nullptr) RUN(side_join(ss, (cudaStream_t)stream));
    if (fused_fwd(s, a->need_backward)) {
        FfnFusedArgs f;
        memset(&f, 0, sizeof(f));
        f.d = s->d; f.E = s->E; f.T = s->T;
        f.a_op = a_op; f.W1_op = a->W1_op; f.W2_op = a->W2_op; f.b1s = a->b1s; f.b2s = a->b2s;
        if (s->dense) {
            f.mode = 1; f.rows = s->T; f.gates = F(L.gates);
            f.residual = a->residual; f.residual_bf16 = a->residual_is_bf16; f.out = a->out; f.out_bf16 = a->out_is_bf16;
            return launch_ffn_fused(f, (cudaStream_t)stream);
        }
        f.mode = 0; f.rows = s->m_cap; f.plan = I(L.plan); f.out = P(L.y); f.out_bf16 = 1;
        RUN(launch_ffn_fused(f, (cudaStream_t)stream));
    } else {
        RUN(moegan_ffn_fwd(s, I(L.plan), a_op, a->W1_op, a->b1s, a->W2_op, a->b2s, P(L.pre), P(L.h_op), P(L.y), stream));
    }
    RUN(moegan_combine_fwd(s, P(L.y), F(L.gates), I(L.token_rows), a->residual, a->residual_is_bf16, a->out, a->out_is_bf16, stream));
    return MOEGAN_OK;
}

int moegan_debug_set_fused_ffn(int32_t mode) {
    g_fused_override.store(mode < 0 ? -1 : (mode != 0 ? 1 : 0));
    return MOEGAN_OK;
}

int moegan_block_bwd(const moegan_shape* s, const moegan_block_bwd_args* a, void* stream) {
    CHECK_SHAPE(s);
    REQ_PTR(a);
    REQ_PTR(a->fwd_workspace); REQ_PTR(a->dout); REQ_PTR(a->dx); REQ_PTR(a->workspace);
    MOEGAN_REQUIRE(s->T > 0, MOEGAN_ERR_BAD_SHAPE, "block_bwd: T must be positive");
    moegan_block_layout L;
    block_layout(s, 1, &L);
    const BlockBwdWs W = block_bwd_layout(s);
    MOEGAN_REQUIRE(a->workspace_bytes >= (int64_t)W.total, MOEGAN_ERR_BAD_SHAPE, "block_bwd: workspace %lld < required %lld bytes",
                   (long long)a->workspace_bytes, (long long)W.total);
    MOEGAN_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, MOEGAN_ERR_MISALIGNED, "block_bwd: workspace must be 256-byte aligned");
    uint8_t* fw = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(a->fwd_workspace));
    uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
    auto P = [&](int64_t off) -> void* { return off < 0 ? nullptr : fw + off; };
    auto F = [&](int64_t off) -> float* { return reinterpret_cast<float*>(P(off)); };
    auto I = [&](int64_t off) -> int32_t* { return reinterpret_cast<int32_t*>(P(off)); };
    float* dgate_row = reinterpret_cast<float*>(ws + W.dgate_row);
    float* dx_router = s->hard ? nullptr : reinterpret_cast<float*>(ws + W.dx_router);
    float* dy_slabs = reinterpret_cast<float*>(ws + W.dy_slabs);
    MOEGAN_REQUIRE(a->x_op == nullptr || s->precision == 0, MOEGAN_ERR_UNSUPPORTED, "block_bwd: x_op is bf16 mode only");
    void* const x_copy = a->x_op != nullptr ? const_cast<void*>(a->x_op) : P(L.x_copy);
    void* const a_op = s->dense ? x_copy : P(L.a_op);
    RUN(moegan_combine_bwd(s, I(L.plan), a->dout, a->dout_is_bf16, P(L.y), I(L.row_token), F(L.row_gate), F(L.gates), ws + W.dy_op, dgate_row,
                           dy_slabs, stream));
    // Small blocks (side_pays): three concurrent chains after combine_bwd -
    //   caller's stream : dgrad2 -> dgrad1 -> wgrad2 -> wgrad1 -> [dx_router ready] -> uncombine
    //   side stream 1   : the router's backward (it only needs dgate_row); signals `dxr` once dx_router exists
    //   side stream 2   : the two bias-gradient slab sums (db2 at once, db1 once dgrad2 is done) - nothing waits for them
    // so the tail after the last weight-gradient GEMM is one un-permute, not five serial launches.
    SideStream* ss = (s->hard || !side_pays(s)) ? nullptr : side_stream();
    void* rstream = stream;
    if (ss != nullptr) {
        RUN(side_fork(ss, (cudaStream_t)stream));
        MOEGAN_CHECK_CUDA(cudaStreamWaitEvent(ss->s2, ss->fork, 0));
        rstream = ss->s;
    }
    const FfnBwdScratch SL = ffn_bwd_scratch_layout(s);
    float* slab_stage = reinterpret_cast<float*>(ws + W.ffn_scratch + SL.slab_stage);
    RUN(ffn_bwd_core(s, I(L.plan), a_op, a->W1_op, a->W2_op, P(L.pre), P(L.h_op), ws + W.dy_op, dy_slabs, ws + W.dpre_op, ws + W.dx_rows,
                     a->dW1, a->db1, a->dW2, a->db2, ws + W.ffn_scratch, (int64_t)ffn_bwd_scratch_bytes_impl(s), stream, ss != nullptr,
                     ss != nullptr ? ss->mid : nullptr));
    if (!s->hard) {
        RUN(router_bwd_core(s, x_copy, a->w, F(L.Wf), F(L.Wt), F(L.Wc), P(L.Wf_op), P(L.Wt_op), F(L.fp), F(L.tp), F(L.raw),
                            I(L.topk_idx), I(L.token_rows), dgate_row, a->dprobs, a->dload, a->temperature, a->annealing_factor,
                            a->sampled, a->mu_f, a->rho_f, F(L.eps_f), a->mu_t, a->rho_t, F(L.eps_t), a->mu_c, a->rho_c, F(L.eps_c),
                            a->kl_out, a->sampled ? a->dkl : nullptr,
                            a->dmu_f, a->drho_f, a->dmu_t, a->drho_t, a->dmu_c, a->drho_c, a->dtemperature, dx_router, a->dw,
                            ws + W.router_ws, (int64_t)router_bwd_layout(s).total, rstream, ss != nullptr ? ss->dxr : nullptr));
    }
    if (ss != nullptr) {
        RUN(launch_slab_sum(s, I(L.plan), dy_slabs, s->d, slab_stage, a->db2, ss->s2));
        MOEGAN_CHECK_CUDA(cudaStreamWaitEvent(ss->s2, ss->mid, 0));
        RUN(launch_slab_sum(s, I(L.plan), reinterpret_cast<const float*>(ws + W.ffn_scratch + SL.dpre_slabs), 4 * s->d, slab_stage, a->db1, ss->s2));
        MOEGAN_CHECK_CUDA(cudaEventRecord(ss->join2, ss->s2));
        MOEGAN_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ss->dxr, 0));
    }
    RUN(moegan_uncombine_bwd(s, ws + W.dx_rows, I(L.token_rows), dx_router, a->dx, stream));
    if (ss != nullptr) {
        RUN(side_join(ss, (cudaStream_t)stream));
        MOEGAN_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ss->join2, 0));
    }
    return MOEGAN_OK;
}

// ---- layers around the block --------------------------------------------------
static int check_ln(int64_t rows, int32_t d) {
    MOEGAN_REQUIRE(rows >= 0 && d >= 128 && d % 128 == 0 && d <= 1024, MOEGAN_ERR_BAD_SHAPE,
                   "layernorm: d (%d) must be a multiple of 128 in [128, 1024], rows (%lld) >= 0", d, (long long)rows);
    return MOEGAN_OK;
}
int moegan_layernorm_fwd(const void* x, int32_t x_is_bf16, const float* gamma, const float* beta, int64_t rows, int32_t d, float eps,
                         void* y, int32_t y_is_bf16, float* mean, float* rstd, void* stream) {
    RUN(check_ln(rows, d));
    REQ_PTR(x); REQ_PTR(gamma); REQ_PTR(beta); REQ_PTR(y); REQ_PTR(mean); REQ_PTR(rstd);
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                     reinterpret_cast<uintptr_t>(beta)) & 15) == 0, MOEGAN_ERR_MISALIGNED, "layernorm_fwd: pointers must be 16-byte aligned");
    return launch_layernorm_fwd(x, x_is_bf16, gamma, beta, rows, d, eps, y, y_is_bf16, nullptr, mean, rstd, (cudaStream_t)stream);
}
int moegan_layernorm_fwd_dual(const void* x, int32_t x_is_bf16, const float* gamma, const float* beta, int64_t rows, int32_t d, float eps,
                              float* y, void* y_op, float* mean, float* rstd, void* stream) {
    RUN(check_ln(rows, d));
    REQ_PTR(x); REQ_PTR(gamma); REQ_PTR(beta); REQ_PTR(y); REQ_PTR(y_op); REQ_PTR(mean); REQ_PTR(rstd);
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(y_op) | reinterpret_cast<uintptr_t>(gamma) |
                     reinterpret_cast<uintptr_t>(beta)) & 15) == 0, MOEGAN_ERR_MISALIGNED, "layernorm_fwd_dual: pointers must be 16-byte aligned");
    return launch_layernorm_fwd(x, x_is_bf16, gamma, beta, rows, d, eps, y, 0, y_op, mean, rstd, (cudaStream_t)stream);
}
int moegan_add_layernorm_fwd(const void* x, const void* addend, int32_t x_is_bf16, const float* gamma, const float* beta, int64_t rows, int32_t d,
                             float eps, void* sum_out, void* y, int32_t y_is_bf16, void* y_op, float* mean, float* rstd, void* stream) {
    RUN(check_ln(rows, d));
    MOEGAN_REQUIRE(d == 128 || d == 256 || d == 512, MOEGAN_ERR_BAD_SHAPE, "add_layernorm_fwd: d = %d must be 128, 256 or 512", d);
    REQ_PTR(x); REQ_PTR(addend); REQ_PTR(sum_out); REQ_PTR(gamma); REQ_PTR(beta); REQ_PTR(y); REQ_PTR(mean); REQ_PTR(rstd);
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(addend) | reinterpret_cast<uintptr_t>(sum_out) | reinterpret_cast<uintptr_t>(y) |
                     reinterpret_cast<uintptr_t>(y_op) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
                   MOEGAN_ERR_MISALIGNED, "add_layernorm_fwd: pointers must be 16-byte aligned");
    return launch_layernorm_fwd(x, x_is_bf16, gamma, beta, rows, d, eps, y, y_is_bf16, y_op, mean, rstd, (cudaStream_t)stream, addend, sum_out);
}
int moegan_layernorm_bwd_add(const void* dy, int32_t dy_is_bf16, const void* x, int32_t x_is_bf16, const float* gamma, const float* mean,
                             const float* rstd, int64_t rows, int32_t d, const void* gadd, void* dx, float* dgamma_dbeta, void* scratch,
                             int64_t scratch_bytes, void* stream) {
    RUN(check_ln(rows, d));
    MOEGAN_REQUIRE(d == 128 || d == 256 || d == 512, MOEGAN_ERR_BAD_SHAPE, "layernorm_bwd_add: d = %d must be 128, 256 or 512", d);
    REQ_PTR(dy); REQ_PTR(x); REQ_PTR(gamma); REQ_PTR(mean); REQ_PTR(rstd); REQ_PTR(gadd); REQ_PTR(dx); REQ_PTR(dgamma_dbeta); REQ_PTR(scratch);
    MOEGAN_REQUIRE(scratch_bytes >= moegan_layernorm_bwd_scratch_bytes(rows, d), MOEGAN_ERR_BAD_SHAPE, "layernorm_bwd_add: scratch too small");
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(gadd) |
                     reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(scratch)) & 15) == 0, MOEGAN_ERR_MISALIGNED,
                   "layernorm_bwd_add: pointers must be 16-byte aligned");
    return launch_layernorm_bwd(dy, dy_is_bf16, x, x_is_bf16, gamma, mean, rstd, rows, d, dx, dgamma_dbeta, reinterpret_cast<float*>(scratch),
                                (cudaStream_t)stream, gadd);
}
int64_t moegan_layernorm_bwd_scratch_bytes(int64_t rows, int32_t d) { return (int64_t)ln_bwd_blocks(rows) * 2 * d * 4 + 256; }
int moegan_layernorm_bwd(const void* dy, int32_t dy_is_bf16, const void* x, int32_t x_is_bf16, const float* gamma, const float* mean,
                         const float* rstd, int64_t rows, int32_t d, void* dx, float* dgamma_dbeta, void* scratch, int64_t scratch_bytes,
                         void* stream) {
    RUN(check_ln(rows, d));
    REQ_PTR(dy); REQ_PTR(x); REQ_PTR(gamma); REQ_PTR(mean); REQ_PTR(rstd); REQ_PTR(dx); REQ_PTR(dgamma_dbeta); REQ_PTR(scratch);
    MOEGAN_REQUIRE(scratch_bytes >= moegan_layernorm_bwd_scratch_bytes(rows, d), MOEGAN_ERR_BAD_SHAPE, "layernorm_bwd: scratch too small");
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
                     reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(scratch)) & 15) == 0, MOEGAN_ERR_MISALIGNED,
                   "layernorm_bwd: pointers must be 16-byte aligned");
    return launch_layernorm_bwd(dy, dy_is_bf16, x, x_is_bf16, gamma, mean, rstd, rows, d, dx, dgamma_dbeta, reinterpret_cast<float*>(scratch),
                                (cudaStream_t)stream);
}

static int check_nhwc(const char* what, int32_t B, int32_t P, int32_t C) {
    MOEGAN_REQUIRE(B >= 0 && P >= 0 && C >= 128 && C % 128 == 0, MOEGAN_ERR_BAD_SHAPE,
                   "%s: channels (%d) must be a positive multiple of 128 (B %d, pixels %d)", what, C, B, P);
    return MOEGAN_OK;
}
int moegan_channel_scale_fwd(const void* x, int32_t is_bf16, const void* s, int32_t s_is_bf16, int32_t B, int32_t P, int32_t C, float slope,
                             void* out, void* stream) {
    RUN(check_nhwc("channel_scale_fwd", B, P, C));
    REQ_PTR(x); REQ_PTR(s); REQ_PTR(out);
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(s)) & 15) == 0,
                   MOEGAN_ERR_MISALIGNED, "channel_scale_fwd: pointers must be 16-byte aligned");
    return launch_channel_scale_fwd(x, is_bf16, s, s_is_bf16, B, P, C, slope, out, (cudaStream_t)stream);
}
int64_t moegan_channel_scale_bwd_scratch_bytes(int32_t B, int32_t P, int32_t C) {
    if (C < 128 || C % 128 != 0) return 256;
    return (int64_t)channel_scale_psplit(B, P, C) * B * C * 4 + 256;
}
int moegan_channel_scale_bwd(const void* g, const void* x, int32_t is_bf16, const void* s, int32_t s_is_bf16, int32_t B, int32_t P, int32_t C,
                             float slope, void* dx, void* ds, void* scratch, int64_t scratch_bytes, void* stream) {
    RUN(check_nhwc("channel_scale_bwd", B, P, C));
    REQ_PTR(g); REQ_PTR(x); REQ_PTR(s); REQ_PTR(dx); REQ_PTR(ds); REQ_PTR(scratch);
    MOEGAN_REQUIRE(scratch_bytes >= moegan_channel_scale_bwd_scratch_bytes(B, P, C), MOEGAN_ERR_BAD_SHAPE, "channel_scale_bwd: scratch too small");
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dx) |
                     reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(ds) | reinterpret_cast<uintptr_t>(scratch)) & 15) == 0,
                   MOEGAN_ERR_MISALIGNED, "channel_scale_bwd: pointers must be 16-byte aligned");
    return launch_channel_scale_bwd(g, x, is_bf16, s, s_is_bf16, B, P, C, slope, dx, ds, reinterpret_cast<float*>(scratch), (cudaStream_t)stream);
}
int64_t moegan_colsum_scratch_bytes(int64_t rows, int32_t cols) {
    if (rows <= 0 || cols <= 0) return 256;
    return (int64_t)colsum_splits(rows, cols) * cols * 4 + 256;
}
int moegan_colsum(const void* x, int32_t is_bf16, int64_t rows, int32_t cols, float* out, void* scratch, int64_t scratch_bytes, void* stream) {
    MOEGAN_REQUIRE(rows >= 0 && cols >= 0 && cols % 4 == 0, MOEGAN_ERR_BAD_SHAPE, "colsum: cols (%d) must be a multiple of 4, rows (%lld) >= 0",
                   cols, (long long)rows);
    REQ_PTR(x); REQ_PTR(out); REQ_PTR(scratch);
    MOEGAN_REQUIRE(scratch_bytes >= moegan_colsum_scratch_bytes(rows, cols), MOEGAN_ERR_BAD_SHAPE, "colsum: scratch too small");
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(scratch)) & 15) == 0,
                   MOEGAN_ERR_MISALIGNED, "colsum: pointers must be 16-byte aligned");
    return launch_colsum(x, is_bf16, rows, cols, out, reinterpret_cast<float*>(scratch), (cudaStream_t)stream);
}
int moegan_offset_sample_fwd(const void* x, int32_t is_bf16, const float* off, int32_t B, int32_t H, int32_t W, int32_t C, void* out,
                             void* stream) {
    RUN(check_nhwc("offset_sample_fwd", B, H * W, C));
    MOEGAN_REQUIRE(H >= 1 && W >= 1 && H * W <= 1024, MOEGAN_ERR_BAD_SHAPE, "offset_sample: H x W (%d x %d) must be in [1, 1024] pixels", H, W);
    REQ_PTR(x); REQ_PTR(off); REQ_PTR(out);
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, MOEGAN_ERR_MISALIGNED,
                   "offset_sample_fwd: pointers must be 16-byte aligned");
    return launch_offset_sample_fwd(x, is_bf16, off, B, H, W, C, out, (cudaStream_t)stream);
}
int moegan_offset_sample_bwd(const void* g, const void* x, int32_t is_bf16, const float* off, int32_t B, int32_t H, int32_t W, int32_t C,
                             void* dx, float* doff, void* stream) {
    RUN(check_nhwc("offset_sample_bwd", B, H * W, C));
    MOEGAN_REQUIRE(H >= 1 && W >= 1 && H * W <= 256, MOEGAN_ERR_BAD_SHAPE, "offset_sample_bwd: H x W (%d x %d) must be in [1, 256] pixels", H, W);
    REQ_PTR(g); REQ_PTR(x); REQ_PTR(off); REQ_PTR(dx); REQ_PTR(doff);
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
                   MOEGAN_ERR_MISALIGNED, "offset_sample_bwd: pointers must be 16-byte aligned");
    return launch_offset_sample_bwd(g, x, is_bf16, off, B, H, W, C, dx, doff, (cudaStream_t)stream);
}

int moegan_demod_coef_fwd(const void* sv, int32_t s_is_bf16, const float* w, int64_t w_so, int64_t w_si, int64_t w_sh, int64_t w_sw, int32_t B,
                          int32_t Cin, int32_t Cout, int32_t KH, int32_t KW, float eps, int32_t compute_wsq, float* wsq, float* d, void* stream) {
    MOEGAN_REQUIRE(B >= 0 && Cin >= 1 && Cout >= 1 && KH >= 1 && KW >= 1 && (int64_t)Cin * Cout < (1ll << 31), MOEGAN_ERR_BAD_SHAPE,
                   "demod_coef_fwd: bad shape B %d Cin %d Cout %d kernel %d x %d", B, Cin, Cout, KH, KW);
    REQ_PTR(w); REQ_PTR(wsq);
    if (B > 0) { REQ_PTR(sv); REQ_PTR(d); }
    return launch_demod_coef_fwd(sv, s_is_bf16, w, w_so, w_si, w_sh, w_sw, B, Cin, Cout, KH, KW, eps, compute_wsq, wsq, d, (cudaStream_t)stream);
}
int moegan_demod_coef_bwd(const float* dd, const float* d, const void* sv, int32_t s_is_bf16, const float* w, int64_t w_so, int64_t w_si,
                          int64_t w_sh, int64_t w_sw, const float* wsq, int32_t B, int32_t Cin, int32_t Cout, int32_t KH, int32_t KW, void* ds,
                          float* dw, void* stream) {
    MOEGAN_REQUIRE(B >= 0 && Cin >= 1 && Cout >= 1 && KH >= 1 && KW >= 1 && (int64_t)Cin * Cout < (1ll << 31), MOEGAN_ERR_BAD_SHAPE,
                   "demod_coef_bwd: bad shape B %d Cin %d Cout %d kernel %d x %d", B, Cin, Cout, KH, KW);
    REQ_PTR(w); REQ_PTR(wsq); REQ_PTR(dw);
    if (B > 0) { REQ_PTR(dd); REQ_PTR(d); REQ_PTR(sv); REQ_PTR(ds); }
    return launch_demod_coef_bwd(dd, d, sv, s_is_bf16, w, w_so, w_si, w_sh, w_sw, wsq, B, Cin, Cout, KH, KW, ds, dw, (cudaStream_t)stream);
}

int moegan_upsample2x_fwd(const void* x, int32_t is_bf16, int32_t B, int32_t H, int32_t W, int32_t C, void* out, void* stream) {
    MOEGAN_REQUIRE(B >= 0 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, MOEGAN_ERR_BAD_SHAPE,
                   "upsample2x: B (%d) >= 0, H x W (%d x %d) >= 1, C (%d) a positive multiple of 8", B, H, W, C);
    REQ_PTR(x); REQ_PTR(out);
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, MOEGAN_ERR_MISALIGNED,
                   "upsample2x_fwd: pointers must be 16-byte aligned");
    return launch_upsample2x_fwd(x, is_bf16, B, H, W, C, out, (cudaStream_t)stream);
}
int moegan_upsample2x_bwd(const void* g, int32_t is_bf16, int32_t B, int32_t H, int32_t W, int32_t C, void* dx, void* stream) {
    MOEGAN_REQUIRE(B >= 0 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, MOEGAN_ERR_BAD_SHAPE,
                   "upsample2x: B (%d) >= 0, H x W (%d x %d) >= 1, C (%d) a positive multiple of 8", B, H, W, C);
    REQ_PTR(g); REQ_PTR(dx);
    MOEGAN_REQUIRE(((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0, MOEGAN_ERR_MISALIGNED,
                   "upsample2x_bwd: pointers must be 16-byte aligned");
    return launch_upsample2x_bwd(g, is_bf16, B, H, W, C, dx, (cudaStream_t)stream);
}

// ---- test / bench hooks ----------------------------------------------------------------------------
int moegan_umma_gemm(int32_t kind, int32_t precision, int32_t M, int32_t N, int32_t K, const void* A, const void* B, float* C, void* stream) {
    REQ_PTR(A); REQ_PTR(B); REQ_PTR(C);
    MOEGAN_REQUIRE(M > 0 && N > 0 && K > 0 && N % 32 == 0 && M % 8 == 0 && K % 8 == 0, MOEGAN_ERR_BAD_SHAPE, "umma_gemm: bad M/N/K %d/%d/%d", M, N, K);
    const uint64_t S = precision == 1 ? 2 : 1;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.passes = precision == 1 ? 3 : 1;
    p.m_blocks_host = -1; p.k_rows_host = -1; p.E = 1; p.epi = EPI_F32;
    p.out0 = C; p.ld_out0 = N; p.N = N;
    p.mn_lbo = g_mn_lbo; p.mn_sbo = g_mn_sbo; p.mn_kadv = g_mn_kadv;
    const int BN = N % 256 == 0 ? 256 : 128;
    const long long n_blocks = ceil_div(N, BN);
    if (kind == 0 || kind == 1) {
        MOEGAN_REQUIRE(K % 64 == 0, MOEGAN_ERR_BAD_SHAPE, "umma_gemm: K (%d) must be a multiple of 64 for kinds 0/1", K);
        GemmOperand a{A, S * K, (uint64_t)M, S * K};
        GemmOperand b = kind == 0 ? GemmOperand{B, S * K, (uint64_t)N, S * K} : GemmOperand{B, S * N, (uint64_t)K, S * N};
        p.K = K; p.a_lo_off = K; p.b_lo_off = kind == 0 ? K : N;
        p.b_rows_per_expert = kind == 0 ? N : K;
        p.m_blocks_host = ceil_div(M, 128); p.out_rows_valid = M;
        return launch_umma_gemm(kind == 0 ? GEMM_ROWS_BK : GEMM_ROWS_BMN, BN, a, b, p, (long long)p.m_blocks_host * n_blocks, (cudaStream_t)stream);
    }
    if (kind == 2) {
        GemmOperand a{A, S * M, (uint64_t)K, S * M};
        GemmOperand b{B, S * N, (uint64_t)K, S * N};
        p.dimA = M; p.a_lo_off = M; p.b_lo_off = N; p.k_rows_host = K;
        return launch_umma_gemm(GEMM_WGRAD, BN, a, b, p, (long long)ceil_div(M, 128) * n_blocks, (cudaStream_t)stream);
    }
    set_error("umma_gemm: unknown kind %d", kind);
    return MOEGAN_ERR_UNSUPPORTED;
}

int moegan_to_operand(const float* src, int64_t rows, int64_t cols, int32_t precision, void* dst, void* stream) {
    REQ_PTR(src); REQ_PTR(dst);
    return launch_to_operand(src, rows, cols, precision == 1, 0, dst, (cudaStream_t)stream);
}

}  // extern "C"
