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


def test_bayesian_router_forward_matches_block_probs():
    pkg = G_.pkg()
    torch.manual_seed(0)
    moe = pkg.SparseMoE(128, 64, 4, precision="fp32").cuda().eval()
    x = torch.randn(2, 128, 4, 4, device="cuda"); w = torch.randn(2, 64, device="cuda")
    with torch.no_grad():
        _, _, probs_blk = moe(x, w, 1.5)
    feat = x.permute(0, 2, 3, 1).reshape(-1, 128)
    text = w.repeat_interleave(16, dim=0)                 # the reference's per-token expansion (:456)
    probs, logits = moe.router(feat, text, sampling=False, annealing_factor=1.5)
    assert torch.equal(probs, probs_blk) and logits.shape == (32, 4) and float(logits.abs().max()) <= 20.0


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
