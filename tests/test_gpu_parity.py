"""GPU parity tests (B200): CUDA path through the C ABI vs (a) the golden fixtures produced by the unmodified
reference and (b) the numpy oracle on identical weights, inputs and eps.

Bars (north_star): expert assignment and token permutation bit-exact (tokens whose top-k logit gap is below
1e-6 are listed and excluded); outputs and gradients within 1e-4 relative in fp32 mode and 2e-2 in bf16 mode,
"relative" = max|a-b| / max|b| per tensor.
"""
import numpy as np
import pytest
import torch

from cases import CASES, BALANCE_WEIGHT, KL_COEF, BAL_COEF, make_inputs, load_golden
from oracle import moe_oracle as O
import gpu_utils as G

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}


# --------------------------------------------------------------------------------------------------
# the tcgen05 GEMM on its own, all operand-major kinds, both precisions
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("mnk", [(128, 128, 64), (384, 256, 192), (1000, 512, 448), (2048, 2048, 512), (72, 160, 128)])
def test_umma_gemm(kind, precision, mnk):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from gpu_probe import run_gemm
    M, N, K = mnk
    err, nan = run_gemm(kind, precision, M, N, K)
    assert nan == 0
    assert err < (2e-5 if precision == 1 else 1e-5), err   # vs fp64 on the same (rounded) operands


# --------------------------------------------------------------------------------------------------
# golden fixtures from the unmodified reference
# --------------------------------------------------------------------------------------------------
GPU_GOLDEN_TRAIN = [n for n in CASES if CASES[n]["mode"] == "train" and CASES[n]["d"] % 128 == 0]
GPU_GOLDEN_EVAL = [n for n in CASES if CASES[n]["mode"] == "eval" and CASES[n]["d"] % 128 == 0]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", GPU_GOLDEN_TRAIN)
def test_golden_train(name, precision):
    c, params, x, w, dout, gp = make_inputs(name)
    g = load_golden(name)
    eps = {k: g[k] for k in O.EPS_KEYS}
    m = G.build_module(params, c["d"], c["Dt"], c["E"], precision=precision, training=True)
    r = G.run_cuda(m, x, w, eps, c["af"], dout=dout, gp=gp, kl_coef=KL_COEF, bal_coef=BAL_COEF, balance_weight=BALANCE_WEIGHT)
    tol = TOL[precision]
    assert G.rel_err(r["out"], g["out"]) < tol
    assert G.rel_err(r["probs"], g["probs"]) < 1e-4          # router is fp32 in both modes
    assert abs(r["kl"] - float(g["kl"])) <= 1e-4 * max(1.0, abs(float(g["kl"])))
    assert abs(r["balance"] - float(g["balance"])) <= 1e-4 * max(abs(float(g["balance"])), 1e-3)
    B, H, W, d = c["B"], c["H"], c["W"], c["d"]
    assert r["out_strides"] == (H * W * d, 1, W * d, d)      # NCHW view of NHWC memory, like the reference (:486)
    ref = {k: g[k] for k in g}
    G.compare_grads(r, ref, params, tol)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", GPU_GOLDEN_EVAL)
def test_golden_eval(name, precision):
    c, params, x, w, dout, gp = make_inputs(name)
    g = load_golden(name)
    m = G.build_module(params, c["d"], c["Dt"], c["E"], precision=precision, training=False)
    r = G.run_cuda(m, x, w, None, c["af"])
    assert np.array_equal(r["probs"], g["probs"])            # hard top-1 routing: bit-exact one-hot
    assert r["kl"] == 0.0
    assert G.rel_err(r["out"], g["out"]) < TOL[precision]


# --------------------------------------------------------------------------------------------------
# oracle comparisons: sparse top-k (not expressible by the reference), ragged sizes, edge cases
# --------------------------------------------------------------------------------------------------
def _random_case(seed, B, H, W, d, Dt, E, mu_scale=30.0, rho=-4.0, temperature=2.0):
    rng = np.random.RandomState(seed)
    params = O.init_params(rng, d, Dt, E, mu_scale=mu_scale, rho=rho)
    params["router.temperature"][:] = temperature
    x = rng.standard_normal((B, d, H, W)).astype(np.float32)
    w = rng.standard_normal((B, Dt)).astype(np.float32)
    dout = rng.standard_normal((B, d, H, W)).astype(np.float32)
    gp = rng.standard_normal((B * H * W, E)).astype(np.float32)
    eps = O.draw_eps(rng, d, Dt, E)
    return params, x, w, dout, gp, eps


def _check_routing(r, ora, E, k):
    """Bit-exact expert assignment + permutation, excluding (and listing) tokens under the 1e-6 logit gap."""
    cache = ora["cache"]
    gap = O.topk_gap(cache.router.probs_soft, cache.router.logits, k)
    excluded = np.nonzero(gap < 1e-6)[0]
    keep = gap >= 1e-6
    idx = r["aux"]["topk_idx"]
    assert np.array_equal(idx[keep], cache.idx[keep]), f"routing mismatch; excluded tokens: {excluded.tolist()}"
    # the permutation is a pure function of the assignment: stable, expert-contiguous, 256-row aligned
    counts, offs, row_token, token_rows = O.stable_permutation(idx, E, block=256)   # MOEGAN_ROW_ALIGN
    plan = r["aux"]["plan"]
    assert plan[0] * 128 == offs[-1] and plan[1] == offs[-1]
    assert np.array_equal(plan[8:8 + E + 1], offs)
    assert np.array_equal(plan[48:48 + E], counts)
    assert np.array_equal(r["aux"]["token_rows"], token_rows)
    assert np.array_equal(r["aux"]["row_token"][:offs[-1]], row_token)
    blk = plan[88:88 + plan[0]]
    for e in range(E):
        assert np.all(blk[offs[e] // 128: offs[e + 1] // 128] == e)
    return excluded


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [
    dict(B=2, H=8, W=8, d=128, Dt=64, E=8, k=2),
    dict(B=3, H=5, W=5, d=256, Dt=128, E=4, k=2),      # ragged: T=75, tokens_per_sample not a multiple of 32
    dict(B=8, H=4, W=4, d=512, Dt=512, E=8, k=2, mu_scale=8.0),   # block-4 geometry, microbench expert config
    dict(B=8, H=4, W=4, d=512, Dt=512, E=8, k=2, mu_scale=30.0),  # saturated logits: many exact ties at the -20 clamp
    dict(B=2, H=16, W=16, d=128, Dt=512, E=8, k=3, mu_scale=15.0),
])
def test_topk_vs_oracle(shape, precision):
    s = shape
    params, x, w, dout, gp, eps = _random_case(7 + s["d"] + s["k"], s["B"], s["H"], s["W"], s["d"], s["Dt"], s["E"],
                                               mu_scale=s.get("mu_scale", 30.0))
    m = G.build_module(params, s["d"], s["Dt"], s["E"], top_k=s["k"], precision=precision)
    r = G.run_cuda(m, x, w, eps, 1.0, dout=dout, gp=gp, kl_coef=KL_COEF, bal_coef=BAL_COEF)
    ora = G.run_oracle(params, x, w, eps, 1.0, top_k=s["k"], dout=dout, gp=gp, kl_coef=KL_COEF, bal_coef=BAL_COEF)
    excluded = _check_routing(r, ora, s["E"], s["k"])     # bit-exact on every token with a top-k logit gap >= 1e-6
    T = s["B"] * s["H"] * s["W"]
    print(f"tokens excluded (top-k logit gap < 1e-6): {len(excluded)}/{T}: {excluded.tolist()}")
    if s.get("mu_scale", 30.0) < 20.0:
        assert len(excluded) == 0
    # ties are broken towards the lowest expert index on both sides, so even the excluded tokens agree here
    assert np.array_equal(r["aux"]["topk_idx"], ora["cache"].idx)
    tol = TOL[precision]
    assert G.rel_err(r["probs"], ora["probs"]) < 1e-4
    assert np.allclose(r["aux"]["gates"].sum(axis=1), 1.0, atol=1e-5)
    assert G.rel_err(r["out"], ora["out"]) < tol
    G.compare_grads(r, ora, params, tol)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [
    dict(B=3, H=5, W=5, d=128, Dt=64, E=4),            # ragged dense: T=75 (TMA zero-fill + row masks)
    dict(B=4, H=8, W=8, d=256, Dt=512, E=4),           # block-8 geometry
    dict(B=2, H=16, W=16, d=128, Dt=512, E=4),         # block-16 geometry
    dict(B=16, H=16, W=16, d=128, Dt=512, E=4),        # T=4096: multi-split bias-gradient reduction, many M blocks
])
def test_dense_vs_oracle(shape, precision):
    s = shape
    params, x, w, dout, gp, eps = _random_case(11 + s["d"], s["B"], s["H"], s["W"], s["d"], s["Dt"], s["E"])
    m = G.build_module(params, s["d"], s["Dt"], s["E"], precision=precision)
    r = G.run_cuda(m, x, w, eps, 3.0, dout=dout, gp=gp, kl_coef=KL_COEF, bal_coef=BAL_COEF)
    ora = G.run_oracle(params, x, w, eps, 3.0, dout=dout, gp=gp, kl_coef=KL_COEF, bal_coef=BAL_COEF)
    tol = TOL[precision]
    assert G.rel_err(r["probs"], ora["probs"]) < 1e-4
    assert G.rel_err(r["out"], ora["out"]) < tol
    G.compare_grads(r, ora, params, tol)


def test_dense_dprobs_generic_path_matches_fused_load():
    """Balance-loss gradient arriving as a dense [T,E] tensor (reference's own loss code) == fused load path."""
    params, x, w, dout, gp, eps = _random_case(21, 2, 8, 8, 128, 64, 4)
    m = G.build_module(params, 128, 64, 4, precision="fp32")
    a = G.run_cuda(m, x, w, eps, 1.0, dout=dout, bal_coef=50.0, fused_balance=True)
    b = G.run_cuda(m, x, w, eps, 1.0, dout=dout, bal_coef=50.0, fused_balance=False)
    for k in ("dx", "dw", "grad.router.feature_mu", "grad.router.combined_rho"):
        assert G.rel_err(a[k], b[k]) < 1e-5, k


def test_kl_unsaturated_and_eps_buffers():
    """KL below the clamp -> non-zero KL gradient (SURVEY 0.6 'unreachable in practice' branch) and the
    eps=None convention: buffers are redrawn in place (t2i_moe_gan.py:349-351) and reused for parity."""
    params, x, w, dout, gp, eps = _random_case(31, 2, 4, 4, 128, 64, 4, mu_scale=3.0, rho=0.55, temperature=1.5)
    m = G.build_module(params, 128, 64, 4, precision="fp32")
    torch.manual_seed(5)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    wt = torch.from_numpy(w).cuda().requires_grad_(True)
    out, kl, probs = m(xt, wt, 1.0)                     # eps=None -> redraw buffers
    (out.sum() + 0.5 * kl).backward()
    used = {k: getattr(m.router, k).detach().cpu().numpy() for k in O.EPS_KEYS}
    assert all(np.abs(v).max() > 0 for v in used.values())
    ora = G.run_oracle(params, x, w, used, 1.0, dout=np.ones_like(x), kl_coef=0.5)
    assert 0.0 < ora["kl"] < 120.0
    assert abs(float(kl) - ora["kl"]) < 1e-4 * ora["kl"]
    assert G.rel_err(out.detach().cpu().numpy(), ora["out"]) < 1e-4
    for k in ("feature_mu", "feature_rho", "text_rho", "combined_mu"):
        got = getattr(m.router, k).grad.cpu().numpy()
        assert G.rel_err(got, ora["grad.router." + k]) < 1e-4, k


def test_empty_expert_and_eval_topk():
    """An expert that receives no token: zero weight gradient, no NaN; eval path of a top_k module is hard top-1."""
    params, x, w, dout, gp, eps = _random_case(41, 2, 8, 8, 128, 64, 4, mu_scale=5.0)
    # tp[:, 0] = sum_k |w_k| > 0 for every sample; Wc[128, 0] = -50 drives expert 0's logit to the -20 clamp
    w = np.abs(w)
    params["router.text_mu"][:] = 0.0
    params["router.text_mu"][:, 0] = 1.0
    params["router.combined_mu"][:, 0] = 0.0
    params["router.combined_mu"][128, 0] = -50.0
    m = G.build_module(params, 128, 64, 4, top_k=2, precision="fp32")
    r = G.run_cuda(m, x, w, eps, 1.0, dout=dout)
    ora = G.run_oracle(params, x, w, eps, 1.0, top_k=2, dout=dout)
    assert (ora["cache"].idx == 0).sum() == 0
    assert r["aux"]["plan"][48] == 0
    assert np.abs(r["grad.experts.0.net.0.weight"]).max() == 0.0 and np.abs(r["grad.experts.0.net.2.bias"]).max() == 0.0
    assert all(np.isfinite(v).all() for k, v in r.items() if k.startswith("grad."))
    assert G.rel_err(r["out"], ora["out"]) < 1e-4
    G.compare_grads(r, ora, params, 1e-4)
    m.eval()
    re = G.run_cuda(m, x, w, None, 1.0)
    oe = G.run_oracle(params, x, w, None, 1.0, training=False)
    assert np.array_equal(re["probs"], oe["probs"]) and G.rel_err(re["out"], oe["out"]) < 1e-4


# --------------------------------------------------------------------------------------------------
# full BASELINE sizes against the oracle (numpy fp64; a few seconds per case) and size-independent properties
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("mu_scale", [8.0, 30.0])
def test_microbench_full_size_vs_oracle(mu_scale, precision):
    """BASELINE configs[1] as benchmarked: 8 experts top-2, 4096 tokens x 512 dim, fwd+bwd.  Routing / permutation bit-exact against
    the oracle (tokens under the 1e-6 top-k logit gap listed and excluded), out and ALL gradients within the north_star bars.
    mu x8 is the router scale bench.py times; mu x30 is the non-uniform companion line (ragged expert loads)."""
    B, H, W, d, Dt, E, k = 8, 16, 32, 512, 512, 8, 2
    params, x, w, dout, gp, eps = _random_case(61, B, H, W, d, Dt, E, mu_scale=mu_scale)
    m = G.build_module(params, d, Dt, E, top_k=k, precision=precision)
    r = G.run_cuda(m, x, w, eps, 1.0, dout=dout, gp=gp, kl_coef=KL_COEF, bal_coef=BAL_COEF)
    ora = G.run_oracle(params, x, w, eps, 1.0, top_k=k, dout=dout, gp=gp, kl_coef=KL_COEF, bal_coef=BAL_COEF)
    excluded = _check_routing(r, ora, E, k)
    counts = r["aux"]["plan"][48:48 + E]
    print(f"configs[1] mu x{mu_scale:g} {precision}: excluded (top-k logit gap < 1e-6) {len(excluded)}/{B * H * W}: {excluded.tolist()}; "
          f"rows per expert {counts.tolist()}")
    tol = TOL[precision]
    assert G.rel_err(r["probs"], ora["probs"]) < 1e-4
    assert G.rel_err(r["out"], ora["out"]) < tol
    errs = G.compare_grads(r, ora, params, tol)
    assert len(errs) == 2 + 7 + 4 * E                                  # dx, dw, 7 router tensors, 4 per expert
    print(f"  worst gradient error {max(errs.values()):.2e} ({max(errs, key=errs.get)})")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_dense_block16_batch256_vs_oracle(precision):
    """The largest block of the benchmarked train step (configs[2]): block 16x16 at batch 256 = 65 536 tokens x 128, E=4 dense-soft,
    af = 3 (epoch 0: temperature clamp saturated), against the oracle - probabilities, output and all gradients."""
    B, H, W, d, Dt, E = 256, 16, 16, 128, 512, 4
    params, x, w, dout, gp, eps = _random_case(62, B, H, W, d, Dt, E, mu_scale=30.0)
    m = G.build_module(params, d, Dt, E, precision=precision)
    r = G.run_cuda(m, x, w, eps, 3.0, dout=dout, gp=gp, kl_coef=KL_COEF, bal_coef=BAL_COEF)
    ora = G.run_oracle(params, x, w, eps, 3.0, dout=dout, gp=gp, kl_coef=KL_COEF, bal_coef=BAL_COEF)
    tol = TOL[precision]
    assert G.rel_err(r["probs"], ora["probs"]) < 1e-4
    gap = O.topk_gap(ora["cache"].router.probs_soft, ora["cache"].router.logits, 1)
    keep = gap >= 1e-6
    assert np.array_equal(r["probs"].argmax(1)[keep], ora["probs"].argmax(1)[keep])
    assert G.rel_err(r["out"], ora["out"]) < tol
    errs = G.compare_grads(r, ora, params, tol)
    print(f"block16 @ batch 256 {precision}: near-ties excluded {int((~keep).sum())}, worst gradient error {max(errs.values()):.2e} ({max(errs, key=errs.get)})")


def test_microbench_size_properties():
    """configs[1]: 8 experts top-2, 4096 tokens x 512 dim, fwd+bwd (bf16)."""
    B, H, W, d, Dt, E, k = 8, 16, 32, 512, 512, 8, 2
    params, x, w, dout, gp, eps = _random_case(51, B, H, W, d, Dt, E)
    m = G.build_module(params, d, Dt, E, top_k=k, precision="bf16")
    r1 = G.run_cuda(m, x, w, eps, 1.0, dout=dout)
    T = B * H * W
    idx, plan = r1["aux"]["topk_idx"], r1["aux"]["plan"]
    counts, offs, row_token, token_rows = O.stable_permutation(idx, E, block=256)
    assert counts.sum() == T * k and np.array_equal(plan[48:48 + E], counts)            # checksum of the histogram
    assert np.array_equal(r1["aux"]["token_rows"], token_rows)
    assert np.array_equal(r1["aux"]["row_token"][:offs[-1]], row_token)                 # stable + sorted inside experts
    assert np.all(idx[:, 0] != idx[:, 1])
    assert np.allclose(r1["aux"]["gates"].sum(axis=1), 1.0, atol=1e-5)
    assert np.allclose(r1["probs"].sum(axis=1), 1.0, atol=1e-5)
    # routing is a function of (x, w, eps) only: a second run is bit-identical (deterministic kernels)
    r1b = G.run_cuda(m, x, w, eps, 1.0, dout=dout)
    assert np.array_equal(r1["out"], r1b["out"]) and np.array_equal(r1["dx"], r1b["dx"])
    assert np.array_equal(r1["grad.experts.3.net.0.weight"], r1b["grad.experts.3.net.0.weight"])
    # backward is linear in dout: grads(2*dout) == 2*grads(dout) up to bf16 rounding of dy
    r2 = G.run_cuda(m, x, w, eps, 1.0, dout=2.0 * dout)
    for key in ("dx", "grad.experts.0.net.2.weight", "grad.router.feature_mu"):
        assert G.rel_err(r2[key], 2.0 * r1[key]) < 2e-2, key
    # torch fp32 reference of the expert math on the routed tokens (floating-point kernel check at full size)
    xt = torch.from_numpy(O.nchw_to_tokens(x)).cuda()
    out_ref = torch.zeros_like(xt)
    gates = torch.from_numpy(r1["aux"]["gates"]).cuda()
    for e in range(E):
        W1, b1, W2, b2 = (torch.from_numpy(params[kk]).cuda() for kk in O.expert_keys(e))
        for j in range(k):
            sel = torch.from_numpy(np.nonzero(idx[:, j] == e)[0]).cuda()
            if sel.numel():
                h = torch.nn.functional.gelu(xt[sel] @ W1.T + b1)
                out_ref[sel] += gates[sel, j:j + 1] * (h @ W2.T + b2)
    got = torch.from_numpy(O.nchw_to_tokens(r1["out"])).cuda()
    assert float((got - out_ref).abs().max() / out_ref.abs().max()) < 2e-2


@pytest.mark.parametrize("shape", [
    dict(B=2, H=8, W=8, d=128, Dt=64, E=16, k=4, mu_scale=10.0),
    dict(B=2, H=8, W=8, d=128, Dt=64, E=32, k=8, mu_scale=10.0),     # MOEGAN_MAX_EXPERTS / MOEGAN_MAX_TOPK
    dict(B=1, H=8, W=8, d=128, Dt=64, E=32, k=None, mu_scale=10.0),  # dense with 32 experts
    dict(B=2, H=4, W=4, d=1024, Dt=64, E=2, k=1, mu_scale=3.0),      # largest supported dim, top-1 soft training gate
])
def test_expert_count_and_topk_limits(shape):
    s = shape
    params, x, w, dout, gp, eps = _random_case(77 + s["E"], s["B"], s["H"], s["W"], s["d"], s["Dt"], s["E"], mu_scale=s["mu_scale"])
    m = G.build_module(params, s["d"], s["Dt"], s["E"], top_k=s["k"], precision="fp32")
    r = G.run_cuda(m, x, w, eps, 1.0, dout=dout, gp=gp, kl_coef=KL_COEF, bal_coef=BAL_COEF)
    ora = G.run_oracle(params, x, w, eps, 1.0, top_k=s["k"], dout=dout, gp=gp, kl_coef=KL_COEF, bal_coef=BAL_COEF)
    if s["k"] is not None:
        _check_routing(r, ora, s["E"], s["k"])
    assert G.rel_err(r["probs"], ora["probs"]) < 1e-4
    assert G.rel_err(r["out"], ora["out"]) < 1e-4
    G.compare_grads(r, ora, params, 1e-4)


def test_fp16_autocast_and_bf16_input_follow_the_reference_contract():
    """Under autocast (the reference uses fp16 autocast, t2i_moe_gan.py:1278) precision='auto' runs the bf16 tensor-core
    path with an fp32 router; the output has x's dtype and the reference's NCHW-view strides."""
    params, x, w, dout, gp, eps = _random_case(91, 2, 8, 8, 128, 64, 4)
    m = G.build_module(params, 128, 64, 4, precision="auto")
    ora = G.run_oracle(params, x, w, eps, 1.0)
    xt, wt = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
    e = tuple(torch.from_numpy(eps[k]).cuda() for k in O.EPS_KEYS)
    for dt in (torch.float16, torch.bfloat16):
        with torch.autocast("cuda", dtype=dt):
            out, kl, probs = m(xt, wt, 1.0, eps=e)
        assert out.dtype == torch.float32 and probs.dtype == torch.float32 and float(kl) == 120.0
        assert G.rel_err(out.detach().cpu().numpy(), ora["out"]) < 2e-2 and G.rel_err(probs.detach().cpu().numpy(), ora["probs"]) < 1e-4
    out32, _, _ = m(xt, wt, 1.0, eps=e)                       # no autocast -> fp32-accurate 3-pass mode
    assert G.rel_err(out32.detach().cpu().numpy(), ora["out"]) < 1e-4
    outb, _, _ = m(xt.bfloat16(), wt, 1.0, eps=e)             # low-precision input: output follows x.dtype
    assert outb.dtype == torch.bfloat16 and outb.shape == xt.shape and outb.stride() == out32.stride()


# --------------------------------------------------------------------------------------------------
# fused FFN kernel (csrc/ffn_fused.cu): forward-only block calls in bf16 mode at d = 128 / 256
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [
    dict(B=3, H=5, W=5, d=128, Dt=64, E=4),                      # ragged dense: T = 75, one half-empty pair tile
    dict(B=4, H=8, W=8, d=256, Dt=512, E=4),                     # block-8 geometry, T = 256: exactly one pair tile
    dict(B=5, H=8, W=8, d=256, Dt=512, E=4),                     # T = 320: the peer CTA of the second pair has 64 real rows
    dict(B=2, H=16, W=16, d=128, Dt=512, E=8),                   # 8 experts x 4 chunks = 32 chunks per tile
    dict(B=1, H=8, W=8, d=128, Dt=64, E=32, mu_scale=10.0),      # MOEGAN_MAX_EXPERTS: 128 chunks per tile
    dict(B=256, H=16, W=16, d=128, Dt=512, E=4),                 # block 16x16 at batch 256: 256 pair tiles over 74 CTA pairs
    dict(B=256, H=8, W=8, d=256, Dt=512, E=4),                   # block 8x8 at batch 256
    dict(B=2, H=8, W=8, d=128, Dt=64, E=8, k=2),                 # sparse top-k, forward only: grouped mode + combine
    dict(B=8, H=8, W=8, d=256, Dt=512, E=8, k=2, mu_scale=8.0),
    dict(B=64, H=8, W=8, d=256, Dt=512, E=4, hard=True),         # eval: hard top-1 (inference path)
    dict(B=3, H=5, W=5, d=128, Dt=64, E=4, hard=True),
])
def test_fused_ffn_forward(shape):
    """no_grad / eval forward in bf16 mode runs the fused Linear-GELU-Linear(-combine) kernel: against the oracle (2e-2) and against
    the separate FFN1 / FFN2 / combine kernels on the same inputs (both are bf16 pipelines: 1e-2 of each other)."""
    s = shape
    pkg = G.pkg()
    k, hard = s.get("k"), s.get("hard", False)
    params, x, w, dout, gp, eps = _random_case(123 + s["d"] + s["E"], s["B"], s["H"], s["W"], s["d"], s["Dt"], s["E"], mu_scale=s.get("mu_scale", 30.0))
    m = G.build_module(params, s["d"], s["Dt"], s["E"], top_k=k, precision="bf16", training=not hard)
    G.run_cuda(m, x, w, None if hard else eps, 1.0)                 # fills the expert-weight operand cache (its conversion is one more launch)
    n0 = pkg._cabi.lib().moegan_debug_launch_count()
    r = G.run_cuda(m, x, w, None if hard else eps, 1.0)
    n_fused = pkg._cabi.lib().moegan_debug_launch_count() - n0
    try:
        pkg.functional.set_fused_ffn(False)
        n0 = pkg._cabi.lib().moegan_debug_launch_count()
        r_sep = G.run_cuda(m, x, w, None if hard else eps, 1.0)
        n_sep = pkg._cabi.lib().moegan_debug_launch_count() - n0
    finally:
        pkg.functional.set_fused_ffn(None)
    assert n_fused == n_sep - (2 if (k is None and not hard) else 1), (n_fused, n_sep)     # dense: FFN1+FFN2+combine -> 1 launch; grouped: FFN1+FFN2 -> 1
    ora = G.run_oracle(params, x, w, None if hard else eps, 1.0, training=not hard, top_k=k)
    assert np.array_equal(r["probs"], r_sep["probs"])
    e_o, e_s = G.rel_err(r["out"], ora["out"]), G.rel_err(r["out"], r_sep["out"])
    print(f"fused FFN {s}: vs oracle {e_o:.2e}, vs separate kernels {e_s:.2e} (separate vs oracle {G.rel_err(r_sep['out'], ora['out']):.2e})")
    assert e_o < 2e-2 and e_s < 1e-2
    assert np.isfinite(r["out"]).all()


@pytest.mark.parametrize("res_dtype", [torch.float32, torch.bfloat16])
def test_fused_ffn_forward_residual_and_output_dtype(res_dtype):
    """The dense-soft fused kernel adds the attention block's skip connection (t2i_moe_gan.py:571) and writes the residual stream's dtype."""
    pkg = G.pkg()
    params, x, w, dout, gp, eps = _random_case(321, 6, 8, 8, 256, 512, 4)
    m = G.build_module(params, 256, 512, 4, precision="bf16")
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    xt, wt = nhwc(torch.from_numpy(x).cuda()), torch.from_numpy(w).cuda()
    res = nhwc(torch.from_numpy(dout).cuda()).to(res_dtype)
    e = tuple(torch.from_numpy(eps[kk]).cuda() for kk in O.EPS_KEYS)
    with torch.no_grad():
        out_f, _, _ = m(xt, wt, 1.0, eps=e, residual=res)
        plain, _, _ = m(xt, wt, 1.0, eps=e)
    assert out_f.dtype == res_dtype and out_f.shape == xt.shape and out_f.stride() == xt.stride()
    want = res.float() + plain.float()
    assert float((out_f.float() - want).abs().max() / want.abs().max()) < (1e-6 if res_dtype == torch.float32 else 8e-3)
