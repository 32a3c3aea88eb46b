"""GPU parity of the kernels around the MoE block against the torch ops they replace in the
reference model: nn.LayerNorm (t2i_moe_gan.py:513-515), the modulate / demodulate / LeakyReLU element-wise steps of
ModulatedConv (:158-186, :243) and the learned-offset grid_sample of ModulatedTransformationModule (:219-240).
The comparison runs on the CPU in fp64 / fp32 torch ops on identical inputs; bars: 1e-5 (fp32 storage), 2e-2 (bf16)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import gpu_utils as G_

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return G_.rel_err(a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy())


@pytest.mark.parametrize("rows,d", [(4099, 128), (1024, 256), (333, 512), (5, 1024), (70001, 128), (4100, 256), (16385, 256), (4097, 512), (9, 384)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm_matches_torch(rows, d, dtype):
    glue = G_.pkg().glue
    g = torch.Generator().manual_seed(rows + d)
    x = (torch.randn(rows, d, generator=g) * 2 + 0.5).to(dtype)
    w = torch.randn(d, generator=g) * 0.5 + 1.0
    b = torch.randn(d, generator=g) * 0.1
    gy = torch.randn(rows, d, generator=g)
    xr = x.double().requires_grad_(True); wr = w.double().requires_grad_(True); br = b.double().requires_grad_(True)
    yr = F.layer_norm(xr, (d,), wr, br, 1e-5)
    yr.backward(gy.double())
    xc = x.cuda().requires_grad_(True); wc = w.cuda().requires_grad_(True); bc = b.cuda().requires_grad_(True)
    yc = glue.layer_norm(xc, wc, bc, 1e-5)
    assert yc.dtype == torch.float32 and yc.shape == x.shape
    yc.backward(gy.cuda())
    tol = 2e-5 if dtype == torch.float32 else 2e-2      # dx is returned in the input's dtype
    assert _rel(yc, yr) < 2e-5
    assert _rel(xc.grad, xr.grad) < tol
    assert _rel(wc.grad, wr.grad) < 2e-5
    assert _rel(bc.grad, br.grad) < 2e-5
    # deterministic: bit-identical on a second run
    xc2 = x.cuda().requires_grad_(True); wc2 = w.cuda().requires_grad_(True); bc2 = b.cuda().requires_grad_(True)
    glue.layer_norm(xc2, wc2, bc2, 1e-5).backward(gy.cuda())
    assert torch.equal(wc2.grad, wc.grad) and torch.equal(bc2.grad, bc.grad) and torch.equal(xc2.grad, xc.grad)


@pytest.mark.parametrize("rows,d", [(4099, 128), (16385, 256), (4097, 512), (33, 128), (70001, 128)])
@pytest.mark.parametrize("dtype,out_dtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32)])
def test_add_layer_norm_matches_torch(rows, d, dtype, out_dtype):
    """The residual add in front of norm2 / norm3 fused into the LayerNorm kernel (t2i_moe_gan.py:545 + :553, :557 + :561): the sum is
    bit-identical to the torch add in the operands' dtype, the normalised rows and all gradients (the sum's consumer and the
    LayerNorm's consumer both send gradient back) match fp64 torch ops; with_operand returns the rounded copy too."""
    glue = G_.pkg().glue
    g = torch.Generator().manual_seed(rows + d)
    x = (torch.randn(rows, d, generator=g) * 2 + 0.5).to(dtype); a = torch.randn(rows, d, generator=g).to(dtype)
    w = torch.randn(d, generator=g) * 0.5 + 1.0; b = torch.randn(d, generator=g) * 0.1
    g_sum = torch.randn(rows, d, generator=g).to(dtype); g_n = torch.randn(rows, d, generator=g).to(out_dtype)
    tr = (x + a).double().requires_grad_(True)              # the reference starts from the SAME rounded sum
    wr = w.double().requires_grad_(True); br = b.double().requires_grad_(True)
    nr = F.layer_norm(tr, (d,), wr, br, 1e-5)
    torch.autograd.backward([tr * 1.0, nr], [g_sum.double(), g_n.double()])
    xc = x.cuda().requires_grad_(True); ac = a.cuda().requires_grad_(True); wc = w.cuda().requires_grad_(True); bc = b.cuda().requires_grad_(True)
    with_op = out_dtype == torch.float32
    outs = glue.add_layer_norm(xc, ac, wc, bc, 1e-5, out_dtype, with_operand=with_op)
    tc, nc = outs[0], outs[1]
    assert tc.dtype == dtype and nc.dtype == out_dtype
    assert torch.equal(tc, (x + a).cuda())
    if with_op:
        assert torch.equal(outs[2], nc.detach().bfloat16())
    torch.autograd.backward([tc, nc], [g_sum.cuda(), g_n.cuda()])
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert _rel(nc, nr) < (2e-5 if out_dtype == torch.float32 else 1e-2)
    assert torch.equal(xc.grad, ac.grad)
    assert _rel(xc.grad, tr.grad) < tol
    assert _rel(wc.grad, wr.grad) < (2e-5 if out_dtype == torch.float32 else 2e-3)
    assert _rel(bc.grad, br.grad) < (2e-5 if out_dtype == torch.float32 else 2e-3)
    # only one of the two outputs used downstream
    xc2 = x.cuda().requires_grad_(True); ac2 = a.cuda().requires_grad_(True)
    t2, n2 = glue.add_layer_norm(xc2, ac2, wc, bc, 1e-5, out_dtype)[:2]
    n2.backward(g_n.cuda())
    tr2 = (x + a).double().requires_grad_(True)
    F.layer_norm(tr2, (d,), w.double(), b.double(), 1e-5).backward(g_n.double())
    assert _rel(xc2.grad, tr2.grad) < tol


def test_layernorm_3d_input_and_autocast():
    glue = G_.pkg().glue
    x = torch.randn(4, 64, 256, device="cuda")
    ln = torch.nn.LayerNorm(256).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = glue.layer_norm(x.bfloat16(), ln.weight, ln.bias, ln.eps)
        ref = ln(x.bfloat16())
    assert y.dtype == ref.dtype == torch.float32
    assert _rel(y, ref) < 1e-5


@pytest.mark.parametrize("B,Cin,Cout,k", [(5, 128, 256, 3), (256, 512, 512, 3), (33, 256, 3, 1), (7, 100, 70, 3), (2, 512, 128, 1), (64, 32, 32, 5)])
@pytest.mark.parametrize("s_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("channels_last", [False, True])
def test_demod_coef_matches_torch(B, Cin, Cout, k, s_dtype, channels_last):
    """Demodulation coefficients of ModulatedConv (t2i_moe_gan.py:165-170), d = rsqrt(s^2 @ sum_taps(w^2)^T + 1e-8), and their
    gradients w.r.t. the modulation and the weight against the same expression in fp64 torch ops; weights contiguous and channels-last
    (dw must come back in the weight's own strides); bf16 modulation as under autocast; ragged sizes (to_rgb: Cout = 3)."""
    glue = G_.pkg().glue
    g = torch.Generator().manual_seed(B + Cin + Cout + k)
    s = (torch.randn(B, Cin, generator=g) * 0.5 + 1.0).to(s_dtype)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    gd = torch.randn(B, Cout, generator=g)
    sr = s.double().requires_grad_(True); wr = w.double().requires_grad_(True)
    dr = torch.rsqrt(sr.pow(2) @ wr.pow(2).sum(dim=(2, 3)).t() + 1e-8)
    dr.backward(gd.double())
    sc = s.cuda().requires_grad_(True)
    wc = w.cuda()
    if channels_last:
        wc = wc.contiguous(memory_format=torch.channels_last)
    wc.requires_grad_(True)
    dc = glue.demod_coef(sc, wc)
    assert dc.dtype == torch.float32 and dc.shape == (B, Cout)
    dc.backward(gd.cuda())
    assert wc.grad.stride() == wc.stride() and sc.grad.dtype == s_dtype
    assert _rel(dc, dr) < 2e-6
    assert _rel(wc.grad, wr.grad) < 1e-5
    assert _rel(sc.grad, sr.grad) < (1e-5 if s_dtype == torch.float32 else 1e-2)
    with torch.no_grad():                                   # the torch-op definition the CPU path uses
        assert _rel(dc, glue.demod_coef_reference(s.cuda(), w.cuda())) < 2e-6
    sc2 = s.cuda().requires_grad_(True); wc2 = wc.detach().clone(memory_format=torch.preserve_format).requires_grad_(True)
    d2 = glue.demod_coef(sc2, wc2); d2.backward(gd.cuda())
    assert torch.equal(d2, dc) and torch.equal(sc2.grad, sc.grad) and torch.equal(wc2.grad, wc.grad)


def test_demod_coef_weight_factor_cache():
    """sum_taps(w^2) is kept between calls while the weights are unchanged (the generator runs twice per train step between optimizer
    steps): the second call launches one kernel less and returns the same coefficients; an in-place weight update or a bumped weight
    generation (CUDA-graph replay of an optimizer step) invalidates it."""
    pkg = G_.pkg(); glue = pkg.glue
    lib = pkg._cabi.lib()
    torch.manual_seed(3)
    w = torch.nn.Parameter(torch.randn(256, 128, 3, 3, device="cuda") * 0.05)
    s = torch.randn(9, 128, device="cuda")
    cache = {}

    def call():
        n0 = lib.moegan_debug_launch_count()
        d = glue.demod_coef(s, w, cache=cache)
        return d, lib.moegan_debug_launch_count() - n0

    d1, n1 = call(); d2, n2 = call()
    assert (n1, n2) == (2, 1) and torch.equal(d1, d2)
    with torch.no_grad():
        w.mul_(1.5)                                          # bumps Parameter._version
    d3, n3 = call()
    assert n3 == 2 and _rel(d3, d1 / 1.5) < 1e-6
    w.data.mul_(2.0)                                         # no version bump: what a captured optimizer step does ...
    pkg.functional.bump_weight_generation()                  # ... and what the trainer calls after every replay
    d4, n4 = call()
    assert n4 == 2 and _rel(d4, d3 / 2.0) < 1e-6
    assert _rel(d4, glue.demod_coef_reference(s, w.detach())) < 2e-6


@pytest.mark.parametrize("B,C,H,W", [(3, 128, 16, 16), (2, 256, 8, 8), (5, 512, 4, 4), (300, 128, 4, 4), (2, 128, 64, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("slope", [1.0, 0.2])
def test_channel_scale_matches_torch(B, C, H, W, dtype, slope):
    glue = G_.pkg().glue
    g = torch.Generator().manual_seed(B * C + H)
    x = torch.randn(B, C, H, W, generator=g).to(dtype)
    s = torch.randn(B, C, generator=g)
    go = torch.randn(B, C, H, W, generator=g).to(dtype)
    xr = x.double().requires_grad_(True); sr = s.double().requires_grad_(True)
    yr = xr * sr[:, :, None, None]
    yr = yr if slope == 1.0 else F.leaky_relu(yr, slope)
    yr.backward(go.double())
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True); sc = s.cuda().requires_grad_(True)
    yc = glue.channel_scale(xc, sc, slope)
    assert yc.dtype == dtype and yc.is_contiguous(memory_format=torch.channels_last)
    yc.backward(go.cuda())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert _rel(yc, yr) < tol
    assert _rel(xc.grad, xr.grad) < tol
    assert _rel(sc.grad, sr.grad) < (1e-5 if dtype == torch.float32 else 2e-2)


def test_channel_scale_nchw_input_and_small_channels():
    """NCHW-contiguous input is converted; channel counts the kernel does not take (to_rgb: 3) use the torch formula."""
    glue = G_.pkg().glue
    x = torch.randn(2, 128, 8, 8, device="cuda"); s = torch.randn(2, 128, device="cuda")
    assert _rel(glue.channel_scale(x, s, 0.2), F.leaky_relu(x * s[:, :, None, None], 0.2)) < 1e-6
    x3 = torch.randn(2, 3, 8, 8, device="cuda"); s3 = torch.randn(2, 3, device="cuda")
    assert _rel(glue.channel_scale(x3, s3), x3 * s3[:, :, None, None]) < 1e-6


@pytest.mark.parametrize("B,C,H,W,scale", [(2, 128, 16, 16, 1.0), (3, 256, 8, 8, 10.0), (2, 512, 4, 4, 30.0), (4, 128, 16, 16, 60.0),
                                           (2, 128, 4, 4, 0.0), (2, 128, 5, 7, 8.0), (3, 384, 3, 3, 10.0), (2, 128, 16, 16, 400.0),
                                           (2, 1024, 2, 2, 5.0), (130, 128, 16, 15, 3.0), (1, 256, 1, 1, 2.0)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_offset_sample_matches_grid_sample(B, C, H, W, scale, dtype):
    """scale stretches the raw offsets so that taps cross pixel borders (10), leave the image (zero padding) and hit the
    [-1, 1] clamp of the grid (30, 60: 0.05 * 60 = 3 grid units)."""
    glue = G_.pkg().glue
    g = torch.Generator().manual_seed(C + H + int(scale))
    x = torch.randn(B, C, H, W, generator=g).to(dtype)
    off = torch.randn(B, 2, H, W, generator=g) * scale
    go = torch.randn(B, C, H, W, generator=g).to(dtype)
    xr = x.double().requires_grad_(True); offr = off.double().requires_grad_(True)
    # reference formula in fp64 (glue.offset_sample_reference builds its base grid in fp32: restated here for fp64)
    gy, gx = torch.meshgrid(torch.linspace(-1, 1, H, dtype=torch.float64), torch.linspace(-1, 1, W, dtype=torch.float64), indexing="ij")
    grid = (torch.stack((gx, gy), dim=2).unsqueeze(0) + offr.permute(0, 2, 3, 1) * 0.05).clamp(-1, 1)
    yr = F.grid_sample(xr, grid, mode="bilinear", align_corners=False)
    yr.backward(go.double())
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True); offc = off.cuda().requires_grad_(True)
    yc = glue.offset_sample(xc, offc)
    assert yc.dtype == dtype and yc.shape == x.shape
    yc.backward(go.cuda())
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert _rel(yc, yr) < tol
    assert _rel(xc.grad, xr.grad) < tol
    assert _rel(offc.grad, offr.grad) < (1e-4 if dtype == torch.float32 else 3e-2)
    # and against the fp32 torch-op restatement of the reference lines
    with torch.no_grad():
        assert _rel(yc, glue.offset_sample_reference(x.cuda().float(), off.cuda())) < tol
    # deterministic scatter
    xc2 = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True); offc2 = off.cuda().requires_grad_(True)
    glue.offset_sample(xc2, offc2).backward(go.cuda())
    assert torch.equal(xc2.grad, xc.grad) and torch.equal(offc2.grad, offc.grad)


def test_modulated_transformation_module_matches_torch_ops(monkeypatch):
    """The host model's MTM / ModulatedConv on the GPU kernels vs the same module evaluated with the torch-op definitions
    (CPU, fp32): forward and all gradients."""
    hm = G_.pkg().host_model
    torch.manual_seed(0)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)      # the cuDNN convolutions in between, in full fp32
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    mtm = hm.ModulatedTransformationModule(128, 256, 3, use_offset=True, resolution=8)
    with torch.no_grad():                               # non-trivial offsets
        for p in mtm.offset_net.parameters():
            p.mul_(8.0)
    x = torch.randn(3, 128, 8, 8); w = torch.randn(3, 512)
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
    yr = mtm(xr, wr); go = torch.randn_like(yr); yr.backward(go)
    ref = {n: p.grad.clone() for n, p in mtm.named_parameters()}
    mtm.zero_grad(); mtm.cuda()
    xc = x.cuda().requires_grad_(True); wc = w.cuda().requires_grad_(True)
    yc = mtm(xc, wc); yc.backward(go.cuda())
    assert _rel(yc, yr) < 1e-4
    assert _rel(xc.grad, xr.grad) < 1e-4 and _rel(wc.grad, wr.grad) < 1e-4
    for n, p in mtm.named_parameters():
        assert _rel(p.grad, ref[n]) < 2e-4, n


@pytest.mark.parametrize("rows", [2050, 8201])            # register-pipelined forward (< 4096 rows) and the bulk-copy ring
def test_layernorm_bf16_output_and_bf16_gradient(rows):
    """norm1 / norm2 under autocast: bf16 in, bf16 out (= the rounding the consuming matmul applies), bf16 upstream gradient."""
    glue = G_.pkg().glue
    g = torch.Generator().manual_seed(7)
    x = torch.randn(rows, 256, generator=g).bfloat16(); w = torch.rand(256, generator=g) + 0.5; b = torch.randn(256, generator=g) * 0.1
    gy = torch.randn(rows, 256, generator=g).bfloat16()
    xr = x.double().requires_grad_(True); wr = w.double().requires_grad_(True); br = b.double().requires_grad_(True)
    yr = F.layer_norm(xr, (256,), wr, br, 1e-5); yr.backward(gy.double())
    xc = x.cuda().requires_grad_(True); wc = w.cuda().requires_grad_(True); bc = b.cuda().requires_grad_(True)
    yc = glue.layer_norm(xc, wc, bc, 1e-5, torch.bfloat16)
    assert yc.dtype == torch.bfloat16
    assert torch.equal(yc, glue.layer_norm(x.cuda(), w.cuda(), b.cuda(), 1e-5).bfloat16())     # exactly the rounded fp32 result
    yc.backward(gy.cuda())
    assert _rel(yc, yr) < 1e-2 and _rel(xc.grad, xr.grad) < 2e-2
    assert _rel(wc.grad, wr.grad) < 1e-4 and _rel(bc.grad, br.grad) < 1e-4


def test_fused_real_pass_matches_reference_formulation_on_gpu(monkeypatch):
    """Trainer D-step with the shared-feature / explicit-adjoint real pass vs the reference's three discriminator calls +
    autograd.grad(create_graph=True): identical discriminator gradients (fp32, TF32 off)."""
    pkg = G_.pkg(); hm = pkg.host_model
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    torch.manual_seed(0)
    D = hm.AuroraDiscriminator().cuda()
    real = torch.randn(4, 3, 64, 64, device="cuda").clamp(-1, 1); text = torch.randn(4, 512, device="cuda"); mis = text[[1, 2, 3, 0]]
    res = []
    for fused in (True, False):
        D.zero_grad(set_to_none=True)
        if fused:
            rp, mp, r1 = hm.discriminator_real_pass(D, real, text, mis, 10.0)
        else:
            r = real.clone().requires_grad_(True)
            rp = D(r, text); r1 = hm.r1_penalty(rp, r, 10.0); mp = D(r.detach(), mis)
        (F.softplus(-rp).mean() + F.softplus(mp).mean() + r1).backward()
        res.append((rp.detach(), mp.detach(), r1.detach(), {n: p.grad.clone() for n, p in D.named_parameters()}))
    a, b = res
    assert _rel(a[0], b[0]) < 1e-5 and _rel(a[1], b[1]) < 1e-5 and _rel(a[2], b[2]) < 1e-4
    for n in a[3]:
        assert _rel(a[3][n], b[3][n]) < 1e-3, n


@pytest.mark.parametrize("rows,cols", [(65536, 128), (16384, 768), (300, 384), (1, 128), (70001, 4), (5000, 1000)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_colsum_matches_torch(rows, cols, dtype):
    glue = G_.pkg().glue
    x = torch.randn(rows, cols, generator=torch.Generator().manual_seed(rows + cols)).to(dtype)
    got = glue.colsum(x.cuda())
    assert got.dtype == torch.float32 and got.shape == (cols,)
    assert _rel(got, x.double().sum(0)) < 2e-5
    assert torch.equal(got, glue.colsum(x.cuda()))          # deterministic


def test_linear_and_multihead_attention_match_torch(monkeypatch):
    """glue.linear == F.linear (values, all gradients); glue.multihead_attention == nn.MultiheadAttention(need_weights=False),
    self- and cross-attention (key length 1, as AttentionBlock uses it), fp32 with TF32 off."""
    glue = G_.pkg().glue
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    torch.manual_seed(0)
    lin = torch.nn.Linear(256, 384).cuda()
    x = torch.randn(5, 77, 256, device="cuda", requires_grad=True); go = torch.randn(5, 77, 384, device="cuda")
    y = glue.linear(x, lin.weight, lin.bias); y.backward(go)
    got = (y.detach(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    x.grad = None; lin.zero_grad()
    yr = torch.nn.functional.linear(x, lin.weight, lin.bias); yr.backward(go)
    for a, b in zip(got, (yr.detach(), x.grad, lin.weight.grad, lin.bias.grad)):
        assert _rel(a, b) < 1e-5
    mha = torch.nn.MultiheadAttention(256, 8, batch_first=True).cuda()
    with torch.no_grad():
        mha.in_proj_bias.normal_(); mha.out_proj.bias.normal_()
    for cross in (False, True):
        q = torch.randn(3, 64, 256, device="cuda", requires_grad=True)
        kv = torch.randn(3, 1, 256, device="cuda", requires_grad=True) if cross else q
        g = torch.randn(3, 64, 256, device="cuda")
        res = []
        for fn in (lambda: glue.multihead_attention(mha, q, kv), lambda: mha(q, kv, kv, need_weights=False)[0]):
            mha.zero_grad(); q.grad = None
            if cross: kv.grad = None
            o = fn(); o.backward(g)
            res.append([o.detach(), q.grad.clone(), kv.grad.clone()] + [p.grad.clone() for p in mha.parameters()])
        for a, b in zip(*res):
            assert _rel(a, b) < 2e-5


@pytest.mark.parametrize("B,C,H,W", [(3, 256, 8, 8), (2, 512, 4, 4), (2, 8, 1, 3), (1, 128, 5, 7), (300, 16, 2, 2)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_upsample2x_matches_interpolate(B, C, H, W, dtype):
    """The 2x bilinear upsampling at the head of GenerativeBlock (t2i_moe_gan.py:646-648) against F.interpolate in fp64:
    border clamps on both sides (H = 1, odd sizes), NCHW and channels-last inputs, the gather backward."""
    glue = G_.pkg().glue
    g = torch.Generator().manual_seed(B * 1000 + C + H)
    x = torch.randn(B, C, H, W, generator=g).to(dtype)
    gy = torch.randn(B, C, 2 * H, 2 * W, generator=g).to(dtype)
    xr = x.double().requires_grad_(True)
    yr = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    yr.backward(gy.double())
    for fmt in (torch.contiguous_format, torch.channels_last):
        xc = x.cuda().contiguous(memory_format=fmt).requires_grad_(True)
        yc = glue.upsample2x_bilinear(xc)
        assert yc.dtype == dtype and yc.shape == yr.shape and yc.is_contiguous(memory_format=torch.channels_last)
        yc.backward(gy.cuda())
        tol = 1e-6 if dtype == torch.float32 else 1e-2
        assert _rel(yc, yr) < tol
        assert _rel(xc.grad, xr.grad) < tol
        y2 = glue.upsample2x_bilinear(xc.detach())
        assert torch.equal(y2, yc.detach())
    if dtype == torch.float32:      # same taps, same evaluation order as ATen's kernel
        ya = F.interpolate(x.cuda().contiguous(memory_format=torch.channels_last), scale_factor=2, mode="bilinear", align_corners=False)
        assert _rel(yc, ya) < 1e-6


@pytest.mark.parametrize("top_k", [None, 2])
def test_layernorm_with_operand_feeds_the_block_without_a_conversion_pass(top_k):
    """norm3 emits the fp32 rows and their bf16 operand copy in one pass; the block given `x_operand=` launches one
    kernel less in dense mode (no to_operand), and output / every gradient are bit-identical to the block converting x itself."""
    import importlib
    pkg = importlib.import_module("moe-gan_cpsc541_b200")
    torch.manual_seed(0)
    B, HW, d = 6, 64, 256
    t = torch.randn(B, HW, d, device="cuda")
    ln = torch.nn.LayerNorm(d).cuda()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5); ln.bias.uniform_(-0.2, 0.2)
    y, y_op = pkg.glue.layer_norm(t, ln.weight, ln.bias, ln.eps, torch.float32, with_operand=True)
    y_ref = pkg.glue.layer_norm(t, ln.weight, ln.bias, ln.eps)
    assert y_op.dtype == torch.bfloat16 and y_op.shape == y.shape and torch.equal(y, y_ref) and torch.equal(y_op, y.to(torch.bfloat16))
    moe = pkg.SparseMoE(d, 512, 8 if top_k else 4, top_k=top_k, precision="bf16").cuda().train()
    moe.redraw_eps = False
    for n in ("epsilon_f", "epsilon_t", "epsilon_c"):
        getattr(moe.router, n).normal_()
    w = torch.randn(B, 512, device="cuda")
    go = torch.randn(B, d, 8, 8, device="cuda")
    res, launches = [], []
    lib = pkg._cabi.lib()
    with torch.no_grad():
        moe(y.reshape(B, 8, 8, d).permute(0, 3, 1, 2), w, 1.0)       # fills the expert-weight operand cache: the counted calls below convert nothing
    for use_op in (False, True):
        moe.zero_grad(set_to_none=True)
        tt = t.clone().requires_grad_(True)
        if use_op:
            yy, yo = pkg.glue.layer_norm(tt, ln.weight, ln.bias, ln.eps, torch.float32, with_operand=True)
        else:
            yy, yo = pkg.glue.layer_norm(tt, ln.weight, ln.bias, ln.eps), None
        x = yy.reshape(B, 8, 8, d).permute(0, 3, 1, 2)
        n0 = lib.moegan_debug_launch_count()
        out, kl, probs = moe(x, w, 1.0, x_operand=yo) if use_op else moe(x, w, 1.0)
        launches.append(lib.moegan_debug_launch_count() - n0)
        (out * go).sum().backward()
        res.append([out.detach().clone(), tt.grad.clone(), ln.weight.grad.clone()] + [p.grad.clone() for p in moe.parameters()])
        ln.zero_grad(set_to_none=True)
    if top_k is None:
        assert launches[1] == launches[0] - 1, launches          # dense mode: the to_operand launch is gone
    for a, b in zip(*res):
        assert torch.equal(a, b)
