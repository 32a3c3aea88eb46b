"""Pin the numpy oracle (forward + hand-derived backward) against the golden fixtures produced by
the UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from cases import CASES, BALANCE_WEIGHT, KL_COEF, BAL_COEF, make_inputs, load_golden
from oracle import moe_oracle as O


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max(), 1e-30)
    return np.abs(a - b).max() / den


def run_oracle(name, dtype=np.float64):
    c, params, x, w, dout, gp = make_inputs(name)
    g = load_golden(name)
    eps = {k: g[k] for k in O.EPS_KEYS}
    training = c["mode"] == "train"
    xt = O.nchw_to_tokens(x)
    out, kl, probs, cache = O.moe_forward(xt, w, params, eps, af=c["af"], training=training,
                                          tokens_per_sample=c["H"] * c["W"], dtype=dtype)
    return c, params, eps, g, out, kl, probs, cache, dout, gp


@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_reference(name):
    c, params, eps, g, out, kl, probs, cache, dout, gp = run_oracle(name)
    ref_out = O.nchw_to_tokens(g["out"])
    assert rel_err(out, ref_out) < 2e-5          # reference is fp32, oracle fp64
    assert abs(kl - float(g["kl"])) <= 1e-4 * max(1.0, abs(float(g["kl"])))
    if c["mode"] == "eval":
        # bit-exact routing: one-hot probs identical (no token has a top-1 gap below 1e-6)
        gap = O.topk_gap(cache.router.probs_soft, cache.router.logits, 1)
        keep = gap >= 1e-6
        assert keep.all(), f"tokens under the 1e-6 gap: {np.nonzero(~keep)[0]}"
        assert np.array_equal(probs[keep], g["probs"][keep])
        assert set(np.unique(g["probs"])) <= {0.0, 1.0}
    else:
        assert rel_err(probs, g["probs"]) < 2e-5      # fp32 reference vs fp64 oracle, logits up to +-20
    # returned tensor is an NCHW view of NHWC memory: strides (HWC, 1, WC, C)
    B, H, W, d = c["B"], c["H"], c["W"], c["d"]
    assert tuple(g["out_strides"]) == (H * W * d, 1, W * d, d)


@pytest.mark.parametrize("name", [n for n in CASES if CASES[n]["mode"] == "train"])
def test_backward_matches_reference_autograd(name):
    c, params, eps, g, out, kl, probs, cache, dout, gp = run_oracle(name)
    T = probs.shape[0]
    bal = O.balance_loss(probs, BALANCE_WEIGHT)
    assert abs(bal - float(g["balance"])) < 1e-5 * max(1.0, abs(bal)) + 1e-7
    loss = float((out * O.nchw_to_tokens(dout)).sum() + KL_COEF * kl + BAL_COEF * bal + (probs * gp).sum())
    assert abs(loss - float(g["loss"])) < 5e-4 * max(1.0, abs(loss))
    dprobs = gp.astype(np.float64) + O.balance_loss_bwd(probs, BALANCE_WEIGHT, dloss=BAL_COEF)
    grads = O.moe_backward(cache, params, eps, O.nchw_to_tokens(dout), dprobs=dprobs, dkl=KL_COEF)
    assert rel_err(grads["dx"], O.nchw_to_tokens(g["dx"])) < 1e-4
    assert rel_err(grads["dw"], g["dw"]) < 1e-4
    for k in params:
        ref = g["grad." + k]
        got = grads[k]
        if np.abs(ref).max() == 0.0:
            assert np.abs(got).max() < 1e-12, k
        else:
            assert rel_err(got, ref) < 2e-4, (k, rel_err(got, ref))


def test_kl_saturated_constant_and_zero_grad():
    """at the default init each router's KL is exactly 120.0 with zero gradient."""
    rng = np.random.RandomState(0)
    params = O.init_params(rng, 128, 512, 4)
    kl, raw = O.kl_divergence(params)
    assert kl == 120.0 and raw > 1e5
    g = O.kl_divergence_bwd(params, 1.0)
    assert all(np.abs(v).max() == 0.0 for v in g.values())
    assert O.kl_train_loop_clamp(3 * kl, O.effective_kl_weight(1e-3, 0, 5)) == pytest.approx(50.0 * 1e-8)


def test_kl_unsaturated_gradient():
    c, params, eps, g, *_ = run_oracle("train_klunsat_d128")
    kl, raw = O.kl_divergence(params)
    assert 0.0 < raw < 120.0 and kl == pytest.approx(float(g["kl"]), rel=1e-5)


def test_af3_saturates_temperature_clamp():
    """af=3 with temperature 4 -> tau clamped to 5 -> dtemperature == 0 (SURVEY 7 hard parts)."""
    g = load_golden("train_af3_d128")
    assert float(np.abs(g["grad.router.temperature"]).max()) == 0.0


@pytest.mark.parametrize("name", ["train_d128_E4", "train_E8_d128"])
def test_topk_generalisation_reduces_to_reference_modes(name):
    """k = E reproduces the training path; k = 1 with hard gates reproduces the eval path."""
    c, params, x, w, dout, gp = make_inputs(name)
    g = load_golden(name)
    eps = {k: g[k] for k in O.EPS_KEYS}
    xt = O.nchw_to_tokens(x)
    tps = c["H"] * c["W"]
    o_ref, _, _, _ = O.moe_forward(xt, w, params, eps, af=c["af"], training=True, tokens_per_sample=tps)
    o_kE, _, _, _ = O.moe_forward(xt, w, params, eps, af=c["af"], training=True, top_k=c["E"], tokens_per_sample=tps)
    assert np.array_equal(o_ref, o_kE)
    # k=2: gates renormalised, rows permuted stably
    o_k2, _, p2, cache = O.moe_forward(xt, w, params, eps, af=c["af"], training=True, top_k=2, tokens_per_sample=tps)
    assert np.allclose(cache.gates.sum(axis=1), 1.0)
    counts, offs, row_token, token_rows = O.stable_permutation(cache.idx, c["E"])
    assert counts.sum() == 2 * xt.shape[0]
    for e in range(c["E"]):
        seg = row_token[offs[e]:offs[e] + counts[e]]
        assert np.all(np.diff(seg) > 0)            # stable: ascending token order inside an expert
        assert np.all(row_token[offs[e] + counts[e]:offs[e + 1]] == -1)
        assert offs[e] % 128 == 0
    assert np.array_equal(row_token[token_rows[:, 0]], np.arange(xt.shape[0]))


def test_topk_backward_finite_difference():
    """1 < k < E is not expressible by the reference: pin the oracle's own backward by finite differences."""
    name = "train_E8_d128"
    c, params, x, w, dout, gp = make_inputs(name)
    g = load_golden(name)
    eps = {k: g[k] for k in O.EPS_KEYS}
    xt = O.nchw_to_tokens(x).astype(np.float64)[:16]
    w = w.astype(np.float64)
    do = O.nchw_to_tokens(dout).astype(np.float64)[:16]
    gp = gp[:16].astype(np.float64)
    tps = 16

    def f(xt_, w_, params_):
        out, kl, probs, cache = O.moe_forward(xt_, w_, params_, eps, af=1.0, training=True, top_k=2, tokens_per_sample=tps)
        return float((out * do).sum() + (probs * gp).sum()), cache

    L0, cache = f(xt, w, params)
    grads = O.moe_backward(cache, params, eps, do, dprobs=gp, dkl=0.0)
    rng = np.random.RandomState(7)
    h = 1e-6
    # directional derivative in x, w and a few parameters
    for key in ("dx", "dw", "router.feature_mu", "router.combined_rho", "experts.3.net.0.weight", "experts.5.net.2.bias"):
        if key == "dx":
            d = rng.standard_normal(xt.shape); Lp, _ = f(xt + h * d, w, params); Lm, _ = f(xt - h * d, w, params)
        elif key == "dw":
            d = rng.standard_normal(w.shape); Lp, _ = f(xt, w + h * d, params); Lm, _ = f(xt, w - h * d, params)
        else:
            d = rng.standard_normal(params[key].shape)
            pp = dict(params); pp[key] = params[key].astype(np.float64) + h * d
            pm = dict(params); pm[key] = params[key].astype(np.float64) - h * d
            Lp, _ = f(xt, w, pp); Lm, _ = f(xt, w, pm)
        fd = (Lp - Lm) / (2 * h)
        an = float((grads[key] * d).sum())
        assert abs(fd - an) < 1e-4 * max(1.0, abs(an)), (key, fd, an)


def test_torch_cpu_port_matches_numpy_oracle_and_golden():
    """oracle/torch_bridge.py (the CPU arm bench.py times) against the golden fixture and the numpy oracle."""
    import torch
    from oracle.torch_bridge import TorchSparseMoE, OracleSparseMoE
    name = "train_d128_E4"
    c, params, x, w, dout, gp = make_inputs(name)
    g = load_golden(name)
    for cls in (TorchSparseMoE, OracleSparseMoE):
        m = cls(c["d"], c["Dt"], c["E"])
        sd = {k: torch.from_numpy(v.copy()) for k, v in params.items()}
        for k in O.EPS_KEYS:
            sd["router." + k] = torch.from_numpy(g[k].copy())
        m.load_state_dict(sd)
        m.redraw_eps = False
        m.train()
        xt = torch.from_numpy(x).requires_grad_(True)
        wt = torch.from_numpy(w).requires_grad_(True)
        out, kl, probs = m(xt, wt, c["af"])
        ((out * torch.from_numpy(dout)).sum() + KL_COEF * kl + (probs * torch.from_numpy(gp)).sum()).backward()
        assert rel_err(out.detach().numpy(), g["out"]) < 2e-5
        assert rel_err(probs.detach().numpy(), g["probs"]) < 2e-5
        dprobs_bal = O.balance_loss_bwd(g["probs"].astype(np.float64), BALANCE_WEIGHT, dloss=BAL_COEF)   # golden loss also had the balance term
        assert float(kl) == float(g["kl"])
        for k in ("experts.1.net.0.weight", "experts.2.net.2.bias"):
            assert rel_err(dict(m.named_parameters())[k].grad.numpy(), g["grad." + k]) < 1e-4, k
