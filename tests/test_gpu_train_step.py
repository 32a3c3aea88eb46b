"""GPU tests of the rows around the hot path: the B200 MoE blocks inside the whole generator (drop-in check against the
CPU oracle module inside the same host model), the full train step (eager and CUDA-graph), and eval-mode inference."""
import copy

import numpy as np
import pytest
import torch

import gpu_utils as G_

pytestmark = pytest.mark.gpu


def _fixed_eps(gen_cpu, gen_gpu, seed=3):
    g = torch.Generator().manual_seed(seed)
    for mc, mg in zip(gen_cpu.moe_layers(), gen_gpu.moe_layers()):
        for name in ("epsilon_f", "epsilon_t", "epsilon_c"):
            e = torch.randn(getattr(mc.router, name).shape, generator=g)
            getattr(mc.router, name).copy_(e)
            getattr(mg.router, name).copy_(e.cuda())
        mc.redraw_eps = False
        mg.redraw_eps = False


def test_generator_with_b200_moe_matches_cpu_oracle_generator():
    """Whole generator forward + backward: MoE blocks on the sm_100a kernels (fp32 mode) vs the numpy-oracle MoE inside
    the identical PyTorch host model on CPU, same weights, inputs and eps."""
    from oracle.torch_bridge import OracleSparseMoE
    pkg = G_.pkg()
    hm = pkg.host_model
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    gen_gpu = hm.AuroraGenerator(moe_kwargs=dict(precision="fp32"))
    gen_cpu = copy.deepcopy(gen_gpu)
    for blk in (gen_cpu.gen_block_4, gen_cpu.gen_block_8, gen_cpu.gen_block_16):
        o = OracleSparseMoE(blk.attn_block.moe.dim, 512)
        o.load_state_dict(blk.attn_block.moe.state_dict())
        blk.attn_block.moe = o
    gen_cpu.moe_layers = lambda: [gen_cpu.gen_block_4.attn_block.moe, gen_cpu.gen_block_8.attn_block.moe, gen_cpu.gen_block_16.attn_block.moe]
    gen_gpu.cuda()
    _fixed_eps(gen_cpu, gen_gpu)
    B = 8
    z, text = torch.randn(B, 512), torch.randn(B, 512)
    wimg = torch.randn(B, 3, 16, 16)
    outs = {}
    for name, gen, dev in (("cpu", gen_cpu, "cpu"), ("gpu", gen_gpu, "cuda")):
        gen.train()
        img, img8, kl, probs = gen(z.to(dev), text.to(dev), return_intermediate=True, return_routing=True, annealing_factor=3.0)
        loss = (img * wimg.to(dev)).sum() + 1e-3 * kl + (probs[-1] * probs[-1]).sum()
        loss.backward()
        outs[name] = dict(img=img.detach().cpu(), img8=img8.detach().cpu(), kl=float(kl.detach()), probs=[p.detach().cpu() for p in probs],
                          grads={n: p.grad.detach().cpu() for n, p in gen.named_parameters() if p.grad is not None})
    a, b = outs["gpu"], outs["cpu"]
    assert a["kl"] == b["kl"] == 360.0                                   # 3 routers x 120
    rel = lambda x, y: float((x - y).abs().max() / y.abs().max().clamp_min(1e-12))
    assert rel(a["img"], b["img"]) < 1e-3 and rel(a["img8"], b["img8"]) < 1e-3
    for pa, pb in zip(a["probs"], b["probs"]):
        assert rel(pa, pb) < 1e-4
    checked = 0
    for n, gb in b["grads"].items():
        if gb.abs().max() < 1e-6:      # numerically-zero gradients (e.g. the cross-attention query path, key length 1)
            continue
        assert rel(a["grads"][n], gb) < 5e-3, n
        checked += 1
    assert checked > 100


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_runs(use_graph):
    pkg = G_.pkg()
    hm = pkg.host_model
    torch.manual_seed(0)
    Gm, Dm = hm.AuroraGenerator().cuda(), hm.AuroraDiscriminator().cuda()
    tr = pkg.MoEGANTrainer(Gm, Dm, amp_dtype=torch.bfloat16, use_cuda_graph=use_graph)
    B = 16
    real = torch.randn(B, 3, 64, 64, device="cuda").clamp(-1, 1)
    text = torch.randn(B, 512, device="cuda")
    before = {n: p.detach().clone() for n, p in Gm.named_parameters()}
    for _ in range(3):
        out = tr.step(real, text)
    torch.cuda.synchronize()
    vals = {k: float(v) for k, v in out.items()}
    assert all(np.isfinite(v) for v in vals.values()), vals
    assert vals["kl"] == 50.0                                            # 360 clamped to 50 in the loop (:1369-1370)
    assert 0.0 <= vals["balance"] <= 0.1
    changed = [n for n, p in Gm.named_parameters() if not torch.equal(p.detach(), before[n])]
    assert any("moe.experts" in n for n in changed) and any("router.feature_mu" in n for n in changed)
    h = tr.step_from_host(real.cpu().pin_memory(), text.cpu().pin_memory())
    assert np.isfinite(h["d_loss"]) and np.isfinite(h["g_loss"])


def test_inference_eval_mode_hard_routing():
    pkg = G_.pkg()
    hm = pkg.host_model
    torch.manual_seed(0)
    Gm = hm.AuroraGenerator().cuda()
    z, text = torch.randn(32, 512, device="cuda"), torch.randn(32, 512, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        img = hm.sample_images(Gm, z, text)
    assert img.shape == (32, 3, 16, 16) and torch.isfinite(img).all()
    Gm.eval()
    with torch.no_grad():
        _, _, kl, probs = Gm(z, text, return_intermediate=True, return_routing=True)
    assert float(kl) == 0.0
    for p in probs:                                                      # one-hot routing in eval (:392-400)
        assert set(torch.unique(p).tolist()) <= {0.0, 1.0} and torch.all(p.sum(dim=1) == 1.0)


def test_activation_checkpointing_matches_plain_backward():
    """enable_checkpointing() (t2i_moe_gan.py:712-760, non-reentrant): the recomputation redraws the same eps, the MoE
    autograd node is re-entrant, and the block keeps the eps it consumed - gradients equal the plain run up to the
    non-determinism of the surrounding cuDNN / ATen backward kernels."""
    pkg = G_.pkg()
    hm = pkg.host_model
    torch.manual_seed(0)
    Gm = hm.AuroraGenerator(moe_kwargs=dict(precision="fp32")).cuda()
    z, text = torch.randn(4, 512, device="cuda"), torch.randn(4, 512, device="cuda")
    grads = []
    for ckpt in (False, True):
        Gm.zero_grad(set_to_none=True)
        Gm.enable_checkpointing() if ckpt else Gm.disable_checkpointing()
        torch.manual_seed(11)
        img, kl = Gm(z, text, annealing_factor=2.0)
        (img.square().mean() + 1e-3 * kl).backward()
        grads.append({n: p.grad.detach().clone() for n, p in Gm.named_parameters() if p.grad is not None})
    assert grads[0].keys() == grads[1].keys()
    for n in grads[0]:
        if float(grads[0][n].abs().max()) < 1e-7:      # numerically-zero gradients (cross-attention query path, key length 1)
            continue
        den = grads[0][n].abs().max()
        assert float((grads[0][n] - grads[1][n]).abs().max() / den) < 2e-3, (n, float(den))


def test_two_forwards_then_backward_of_the_first_uses_its_own_eps():
    """The block saves the eps it consumed: a later training forward (which redraws router.epsilon_*) must not change
    the gradients of an earlier one."""
    pkg = G_.pkg()
    torch.manual_seed(0)
    moe = pkg.SparseMoE(128, 64, 4, precision="fp32").cuda()
    x = torch.randn(2, 128, 8, 8, device="cuda", requires_grad=True)
    w = torch.randn(2, 64, device="cuda", requires_grad=True)
    torch.manual_seed(5)
    out1, kl1, p1 = moe(x, w, 1.0)
    eps_used = [getattr(moe.router, n).clone() for n in ("epsilon_f", "epsilon_t", "epsilon_c")]
    out2, _, _ = moe(x.detach(), w.detach(), 1.0)          # redraws the buffers
    out1.square().sum().backward()
    g_late = moe.router.feature_rho.grad.clone()
    moe.zero_grad(set_to_none=True); x.grad = None; w.grad = None
    o, _, _ = moe(x, w, 1.0, eps=tuple(eps_used))
    o.square().sum().backward()
    assert torch.equal(o, out1)
    assert float((moe.router.feature_rho.grad - g_late).abs().max()) == 0.0


@pytest.mark.parametrize("training,af", [(False, 1.5), (True, 1.0), (True, 3.0)])
def test_bayesian_router_forward_vs_oracle(training, af):
    """BayesianRouter.forward(feature [T,d], text [T,Dt]) -> (probs [T,E], logits [T,E] = clamp(raw / tau, +-20)),
    both against the oracle's router (t2i_moe_gan.py:335-402) on the eps the call consumed - not against the block's own output."""
    import numpy as np
    from oracle import moe_oracle as O
    pkg = G_.pkg()
    rng = np.random.RandomState(3)
    d, Dt, E, B, HW = 128, 64, 4, 2, 16
    params = O.init_params(rng, d, Dt, E, mu_scale=30.0)
    params["router.temperature"][:] = 2.0
    m = G_.build_module(params, d, Dt, E, precision="fp32", training=training)
    x = rng.standard_normal((B * HW, d)).astype(np.float32); w = rng.standard_normal((B, Dt)).astype(np.float32)
    feat = torch.from_numpy(x).cuda()
    text = torch.from_numpy(w).cuda().repeat_interleave(HW, dim=0)       # the reference's per-token expansion (:456)
    torch.manual_seed(9)
    probs, logits = m.router(feat, text, sampling=True, annealing_factor=af)
    eps = {k: getattr(m.router, k).detach().cpu().numpy() for k in O.EPS_KEYS}
    p_o, l_o, cache = O.router_forward(x, w, params, eps, af, training, sampling=True, tokens_per_sample=HW)
    assert logits.shape == (B * HW, E) and probs.shape == (B * HW, E)
    assert G_.rel_err(logits.cpu().numpy(), l_o) < 1e-4 and float(logits.abs().max()) <= 20.0
    if training:
        assert G_.rel_err(probs.cpu().numpy(), p_o) < 1e-4
    else:
        assert np.array_equal(probs.cpu().numpy(), p_o)                    # eval: one-hot (:392-400)
    # and the block reports the same probabilities for the same tokens / eps
    m.redraw_eps = False
    with torch.no_grad():
        _, _, probs_blk = m(feat.view(B, HW, 1, d).permute(0, 3, 1, 2), torch.from_numpy(w).cuda(), af)
    assert torch.equal(probs, probs_blk)


@pytest.mark.parametrize("res_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("top_k", [None, 2])
def test_fused_residual_equals_separate_add(res_dtype, top_k):
    """forward(..., residual=r) == r + forward(...) (t2i_moe_gan.py:571), outputs and every gradient; the
    upstream gradient arrives in the residual's dtype and is consumed as such by the combine backward."""
    pkg = G_.pkg()
    torch.manual_seed(0)
    moe = pkg.SparseMoE(128, 64, 4, top_k=top_k, precision="fp32").cuda()
    moe.redraw_eps = False
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    x0 = nhwc(torch.randn(3, 128, 8, 8, device="cuda")); w0 = torch.randn(3, 64, device="cuda")
    r0 = nhwc(torch.randn(3, 128, 8, 8, device="cuda")).to(res_dtype)
    go = nhwc(torch.randn(3, 128, 8, 8, device="cuda")).to(res_dtype)
    res = []
    for fused in (True, False):
        moe.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True); w = w0.clone().requires_grad_(True); r = r0.clone().requires_grad_(True)
        if fused:
            out, kl, probs = moe(x, w, 1.5, residual=r)
        else:
            o, kl, probs = moe(x, w, 1.5)
            out = r + o.to(res_dtype)
        assert out.dtype == res_dtype and out.shape == x.shape
        torch.autograd.backward([out, kl], [go, torch.full((), 1e-3, device="cuda")])
        res.append((out.detach().float(), x.grad.clone(), w.grad.clone(), r.grad.float().clone(),
                    {n: p.grad.clone() for n, p in moe.named_parameters() if p.grad is not None}))
    a, b = res
    tol = 0.0 if res_dtype == torch.float32 else 2e-2
    rel = lambda u, v: float((u - v).abs().max() / v.abs().max().clamp_min(1e-30))
    assert rel(a[0], b[0]) <= tol
    assert torch.equal(a[3], b[3])                               # d residual = upstream gradient, untouched
    assert rel(a[1], b[1]) <= 1e-6 and rel(a[2], b[2]) <= 1e-5   # same dout values reach the block either way
    for n in a[4]:
        assert rel(a[4][n], b[4][n]) <= 1e-5, n


def _small_trainer(pkg, use_graph, seed=0, amp=torch.bfloat16):
    hm = pkg.host_model
    torch.manual_seed(seed)
    Gm, Dm = hm.AuroraGenerator().cuda(), hm.AuroraDiscriminator().cuda()
    return pkg.MoEGANTrainer(Gm, Dm, amp_dtype=amp, use_cuda_graph=use_graph), Gm, Dm


def test_graph_capture_does_not_advance_training():
    """Capturing (or re-capturing after set_epoch) needs three warm-up steps on the live tensors; they are undone in place, so a
    capture costs no optimizer steps: parameters, buffers and AdamW state are what they were (state the warm-up created is zero)."""
    pkg = G_.pkg()
    tr, Gm, Dm = _small_trainer(pkg, True)
    real = torch.randn(8, 3, 64, 64, device="cuda").clamp(-1, 1); text = torch.randn(8, 512, device="cuda")
    before = {n: p.detach().clone() for net in (Gm, Dm) for n, p in net.named_parameters()}
    tr._capture(real, text)
    after = {n: p.detach() for net in (Gm, Dm) for n, p in net.named_parameters()}
    assert all(torch.equal(before[n], after[n]) for n in before)
    for opt in (tr.opt_g, tr.opt_d):
        for st in opt.state.values():
            assert all(float(v.abs().max()) == 0.0 for v in st.values() if torch.is_tensor(v))
    out = tr.step(real, text)                          # first replay = first optimizer step
    torch.cuda.synchronize()
    assert all(float(st["step"]) == 1.0 for st in tr.opt_g.state.values())
    assert any(not torch.equal(before[n], p.detach()) for n, p in Gm.named_parameters())
    assert np.isfinite(float(out["g_loss"]))


def test_weight_cache_is_invalidated_by_graph_replay():
    """a captured AdamW updates the expert weights without bumping Parameter._version; a non-captured forward after
    N replays must not be served the bf16 operand copies converted before them."""
    pkg = G_.pkg()
    hm = pkg.host_model
    tr, Gm, Dm = _small_trainer(pkg, True)
    real = torch.randn(8, 3, 64, 64, device="cuda").clamp(-1, 1); text = torch.randn(8, 512, device="cuda")
    z, t2 = torch.randn(8, 512, device="cuda"), torch.randn(8, 512, device="cuda")

    def sample():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return hm.sample_images(Gm, z, t2).float().clone()
    tr._capture(real, text)
    img0 = sample()                                    # fills every block's WeightCache outside capture
    for _ in range(5):
        tr.step(real, text)                            # replays: the weights move, _version does not
    img1 = sample()
    for m in Gm.moe_layers():
        m._wcache.invalidate()
    img2 = sample()                                    # forced fresh conversion of the current weights
    torch.cuda.synchronize()
    assert torch.equal(img1, img2), "stale expert weights were served after graph replays"
    assert not torch.equal(img0, img1)


@pytest.mark.parametrize("top_k,precision", [(None, "bf16"), (2, "bf16"), (None, "fp32")])
def test_block_graph_replay_equals_eager_bitwise(top_k, precision):
    """The library's launches are the same in capture and eager mode (side streams, cached weight conversion, workspaces):
    a CUDA-graph replay of block forward + backward reproduces the eager outputs and every gradient bit for bit."""
    pkg = G_.pkg()
    torch.manual_seed(0)
    E = 8 if top_k else 4
    moe = pkg.SparseMoE(256, 512, E, top_k=top_k, precision=precision).cuda().train()
    with torch.no_grad():
        for n in ("feature_mu", "text_mu", "combined_mu"):
            getattr(moe.router, n).mul_(20.0)
    x = torch.randn(16, 8, 8, 256, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)
    w = torch.randn(16, 512, device="cuda", requires_grad=True)
    dout = torch.randn(16, 8, 8, 256, device="cuda").permute(0, 3, 1, 2)
    eps = tuple(torch.randn_like(getattr(moe.router, n)) for n in ("epsilon_f", "epsilon_t", "epsilon_c"))
    klw = torch.full((), 0.3, device="cuda")
    params = list(moe.parameters())

    def step():
        for p in params:
            p.grad = None
        x.grad = None; w.grad = None
        out, kl, probs = moe(x, w, 1.0, eps=eps)
        bal = pkg.moe_balance_loss([probs], 0.5)
        torch.autograd.backward([out, kl, bal], [dout, klw, torch.ones((), device="cuda")])
        return out, kl, probs

    # every use of the leaves happens on a side stream: autograd pins a leaf's AccumulateGrad node to the stream of its first use, and
    # a node pinned to the legacy default stream cannot take part in a capture (cudaErrorStreamCaptureImplicit)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        out, kl, probs = step()
        eager = [out.detach().clone(), kl.detach().clone(), probs.detach().clone(), x.grad.clone(), w.grad.clone()] + [p.grad.clone() for p in params]
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out, kl, probs = step()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    replay = [out, kl, probs, x.grad, w.grad] + [p.grad for p in params]
    names = ["out", "kl", "probs", "dx", "dw"] + [n for n, _ in moe.named_parameters()]
    for n, a, b in zip(names, eager, replay):
        assert torch.equal(a, b), f"{n}: graph replay differs from eager (max |diff| {float((a.float() - b.float()).abs().max()):.3e})"


def test_train_step_graph_replay_matches_eager_step():
    """Whole train step: a CUDA-graph replay and an eager step from the same weights, batch and RNG state report the same losses
    (graph-safe Philox: the captured random ops consume the generator exactly like the eager ones)."""
    pkg = G_.pkg()
    tr_e, Ge, De = _small_trainer(pkg, False, seed=4)
    tr_g, Gg, Dg = _small_trainer(pkg, True, seed=4)
    assert all(torch.equal(a, b) for a, b in zip(Ge.parameters(), Gg.parameters()))
    real = torch.randn(16, 3, 64, 64, device="cuda").clamp(-1, 1); text = torch.randn(16, 512, device="cuda")
    tr_g._capture(real, text)
    torch.manual_seed(77); oe = {k: float(v) for k, v in tr_e.step(real, text).items()}
    torch.manual_seed(77); og = {k: float(v) for k, v in tr_g.step(real, text).items()}
    torch.cuda.synchronize()
    print(f"[graph vs eager] eager {oe}  graph {og}")
    for k in oe:
        assert og[k] == pytest.approx(oe[k], rel=5e-3, abs=1e-5), (k, oe[k], og[k])
    # and the update that followed points the same way
    dot = ne = ng = 0.0
    for a, b in zip(Ge.parameters(), Gg.parameters()):
        if a.grad is None or b.grad is None:
            continue
        dot += float((a.grad.double() * b.grad.double()).sum()); ne += float(a.grad.double().pow(2).sum()); ng += float(b.grad.double().pow(2).sum())
    assert dot / max((ne * ng) ** 0.5, 1e-30) > 0.99
