"""CPU checks of host-model re-formulations that need neither the reference tree nor a GPU."""
import torch


def test_text_conditioned_head_equals_cat_then_conv(pkg):
    """The discriminator's output layer without the concatenated tensor (host_model.text_conditioned_head) against
    conv2d(cat([f, tiled text])) (t2i_moe_gan.py:899-907): values and every gradient, fp64, both input sizes of the train loop."""
    import torch.nn.functional as F
    hm = pkg.host_model
    torch.manual_seed(0)
    for hw in (16, 4):                                   # features of a 64x64 real image / of a 16x16 generated one
        f = torch.randn(3, 256, hw, hw, dtype=torch.float64, requires_grad=True)
        t = torch.randn(3, 128, dtype=torch.float64, requires_grad=True)
        w = torch.randn(1, 384, 4, 4, dtype=torch.float64, requires_grad=True)
        b = torch.randn(1, dtype=torch.float64, requires_grad=True)
        ref = F.conv2d(torch.cat([f, t[:, :, None, None].expand(-1, -1, hw, hw)], dim=1), w, b)
        gy = torch.randn_like(ref)
        g_ref = torch.autograd.grad(ref, (f, t, w, b), gy)
        out = hm.text_conditioned_head(f, t, w, b)
        assert out.shape == ref.shape and torch.allclose(out, ref, rtol=1e-12, atol=1e-12)
        for a, r in zip(torch.autograd.grad(out, (f, t, w, b), gy), g_ref):
            assert torch.allclose(a, r, rtol=1e-11, atol=1e-11)


def test_glue_cpu_definitions(pkg):
    """On CPU tensors the glue functions ARE the torch formulas they replace (the definition the CPU arm uses)."""
    import torch.nn.functional as F
    glue = pkg.glue
    torch.manual_seed(1)
    x = torch.randn(2, 16, 3, 5)
    assert torch.equal(glue.upsample2x_bilinear(x), F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False))
    s = torch.randn(2, 16)
    assert torch.allclose(glue.channel_scale(x, s, 0.2), F.leaky_relu(x * s[:, :, None, None], 0.2))
    t = torch.randn(7, 128)
    w, b = torch.randn(128), torch.randn(128)
    assert torch.allclose(glue.layer_norm(t, w, b, 1e-5), F.layer_norm(t, (128,), w, b, 1e-5))
    # residual add + LayerNorm (AttentionBlock :545 + :553, :557 + :561)
    a = torch.randn(7, 128)
    t2, n2 = glue.add_layer_norm(t, a, w, b, 1e-5)
    assert torch.equal(t2, t + a) and torch.allclose(n2, F.layer_norm(t + a, (128,), w, b, 1e-5))
    t3, n3, op3 = glue.add_layer_norm(t, a, w, b, 1e-5, with_operand=True)
    assert torch.equal(t3, t2) and torch.equal(n3, n2) and op3.dtype == torch.bfloat16 and torch.equal(op3, n2.bfloat16())


def test_demod_coef_equals_the_weight_space_formula(pkg):
    """glue.demod_coef's definition, d = rsqrt(s^2 @ sum_taps(w^2)^T + 1e-8), is the factor the reference computes in weight space
    (t2i_moe_gan.py:161-166: weight = W[None] * style; d = rsqrt((weight ** 2).sum(dim=(2, 3, 4)) + 1e-8)) - values and both gradients,
    fp64."""
    glue = pkg.glue
    torch.manual_seed(2)
    B, Cin, Cout, k = 3, 6, 5, 3
    s = torch.randn(B, Cin, dtype=torch.float64, requires_grad=True)
    W = torch.randn(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
    gd = torch.randn(B, Cout, dtype=torch.float64)
    weight = W.unsqueeze(0) * s.view(B, 1, Cin, 1, 1)
    d_ref = torch.rsqrt((weight ** 2).sum(dim=(2, 3, 4)) + 1e-8)
    g_ref = torch.autograd.grad(d_ref, (s, W), gd)
    d = torch.rsqrt(s.pow(2) @ W.pow(2).sum(dim=(2, 3)).t() + 1e-8)          # = glue.demod_coef_reference in fp64
    assert torch.allclose(d, d_ref, rtol=1e-12, atol=1e-12)
    for a_, r_ in zip(torch.autograd.grad(d, (s, W), gd), g_ref):
        assert torch.allclose(a_, r_, rtol=1e-10, atol=1e-12)
    # the fp32 CPU path of glue.demod_coef is that expression
    d32 = glue.demod_coef(s.detach().float(), W.detach().float())
    assert d32.dtype == torch.float32 and torch.allclose(d32.double(), d_ref, rtol=1e-5, atol=1e-6)
