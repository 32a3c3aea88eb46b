"""Host model for the full-step configurations (BASELINE configs[2..4]): the generator / discriminator that sit
AROUND the hot path, restated in plain PyTorch so that the train step can run where the reference tree is not
available (the GPU box).  These layers are out of the hot-path scope ("stays
PyTorch/cuDNN"); only the `moe` sub-module of each attention block runs the sm_100a kernels.

Module / parameter names follow the reference (moegan/t2i_moe_gan.py) so that a checkpoint written by
moegan/sagemaker_train.py:297-301 loads here (`load_reference_checkpoint`):
    AuroraGenerator :668-855, GenerativeBlock :622-666, ConvolutionBlock :579-621, ModulatedTransformationModule
    :188-247, ModulatedConv :122-186, AttentionBlock :493-576, AuroraDiscriminator :858-907, losses :917-949.

One deliberate re-formulation: ModulatedConv does not materialise per-sample weights [B,out,in,k,k] and a grouped
convolution (:158-176); it scales the input by the style, runs ONE ordinary convolution with the shared weight
and scales the output by the demodulation factor  d[b,o] = rsqrt(sum_i s[b,i]^2 * sum_k W[o,i,k]^2 + 1e-8),
which is the same function (StyleGAN2's "fused_modconv=False" identity) and is what SURVEY 8(f) N3 asks for.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.parametrizations import weight_norm

from . import glue
from .sparse_moe import SparseMoE, moe_balance_loss

LATENT_DIM = 512          # t2i_moe_gan.py:20
TEXT_EMBEDDING_DIM = 512  # :21


class ModulatedConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, latent_dim=LATENT_DIM, padding=0, demodulate=True):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.padding, self.demodulate = padding, demodulate
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        nn.init.kaiming_normal_(self.weight, a=0, mode="fan_in", nonlinearity="leaky_relu")
        self.modulation = nn.Linear(latent_dim, in_channels)
        self._demod_cache = {}                                            # sum_taps(w^2) between calls with unchanged weights (glue.demod_coef)
        nn.init.normal_(self.modulation.weight, mean=0.0, std=0.02)
        nn.init.zeros_(self.modulation.bias)

    def forward(self, x, w, act_slope=None):
        """act_slope: fuse the LeakyReLU that ModulatedTransformationModule applies to the result (:243-245)."""
        s = self.modulation(w)                                            # [B, in]
        # under autocast F.conv2d casts (and caches) the weight itself; an explicit .to() would re-cast on every call
        wgt = self.weight if (torch.is_autocast_enabled() or self.weight.dtype == x.dtype) else self.weight.to(x.dtype)
        # modulate: x * s  /  demodulate (+ activation): lrelu(y * d) - one fused pass each on the GPU (glue.channel_scale)
        y = F.conv2d(glue.channel_scale(x, s), wgt, padding=self.padding)
        slope = 1.0 if act_slope is None else float(act_slope)
        if self.demodulate:
            # d[b, o] = rsqrt(sum_i s[b,i]^2 * sum_taps w[o,i]^2 + 1e-8), fp32 as in the reference (:165-170); on the GPU two launches
            # forward / two backward (glue.demod_coef) instead of ~7 / ~12 element-wise + SIMT-sgemm launches
            d = glue.demod_coef(s, self.weight, cache=self._demod_cache)  # [B, out]
            return glue.channel_scale(y, d, slope)
        return y if act_slope is None else F.leaky_relu(y, slope)


class ModulatedTransformationModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, latent_dim=LATENT_DIM, use_offset=False, resolution=None):
        super().__init__()
        self.use_offset = bool(use_offset and resolution is not None and resolution <= 16)
        self.modulated_conv = ModulatedConv(in_channels, out_channels, kernel_size, latent_dim=latent_dim, padding=kernel_size // 2)
        if self.use_offset:
            self.offset_net = nn.Sequential(nn.Conv2d(in_channels, 32, 3, padding=1), nn.LeakyReLU(0.2, inplace=True),
                                            nn.Conv2d(32, 2, 3, padding=1))
        self.activation = nn.LeakyReLU(0.2, inplace=True)             # kept for the module tree; applied inside modulated_conv

    def forward(self, x, w):
        if self.use_offset:
            # grid = clamp(meshgrid + 0.05 * offset_net(x), -1, 1); bilinear resampling (glue.offset_sample: NHWC kernels)
            x = glue.offset_sample(x, self.offset_net(x))
        return self.modulated_conv(x, w, act_slope=self.activation.negative_slope)


class ConvolutionBlock(nn.Module):
    def __init__(self, in_channels, out_channels, latent_dim=LATENT_DIM, resolution=None):
        super().__init__()
        self.mtm1 = ModulatedTransformationModule(in_channels, out_channels, 3, latent_dim, True, resolution)
        self.mtm2 = ModulatedTransformationModule(out_channels, out_channels, 3, latent_dim, True, resolution)
        self.skip_proj = ModulatedConv(in_channels, out_channels, 1, latent_dim) if in_channels != out_channels else None

    def forward(self, x, w):
        out = self.mtm2(self.mtm1(x, w), w)
        return out + (x if self.skip_proj is None else self.skip_proj(x, w))


class AttentionBlock(nn.Module):
    """proj_in -> LN -> self-attn -> LN -> cross-attn(text, length 1) -> LN -> MoE -> residual -> proj_out (:525-576)."""

    def __init__(self, dim, text_dim=512, heads=8, moe_kwargs=None):
        super().__init__()
        self.dim, self.heads = dim, heads
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.text_proj = nn.Linear(text_dim, dim)
        self.self_attn = nn.MultiheadAttention(dim, heads, batch_first=True)
        self.cross_attn = nn.MultiheadAttention(dim, heads, batch_first=True)
        self.moe = SparseMoE(dim, text_dim, **(moe_kwargs or {}))          # <- the hot path
        self.proj_in = ModulatedConv(dim, dim, 1)
        self.proj_out = ModulatedConv(dim, dim, 1)

    def forward(self, x, w, text_seq, kl_losses=None, annealing_factor=1.0):
        B, C, H, W = x.shape
        # nn.LayerNorm through the library's kernels; norm1 / norm2 feed autocast matmuls only, which round their input to the
        # autocast dtype first thing - emit that dtype directly.  norm3 feeds the fp32 router: stays fp32.
        mm = torch.get_autocast_dtype("cuda") if (x.is_cuda and torch.is_autocast_enabled()) else torch.float32
        ln = lambda m, v, dt=torch.float32: glue.layer_norm(v, m.weight, m.bias, m.eps, dt)
        t = self.proj_in(x, w).permute(0, 2, 3, 1).reshape(B, H * W, C)
        n = ln(self.norm1, t, mm)
        a1 = glue.multihead_attention(self.self_attn, n, n)               # = self.self_attn(n, n, n, need_weights=False)[0]
        # `t = t + attn; norm(t)` (:545 + :553, :557 + :561): the residual add rides in the LayerNorm kernel (glue.add_layer_norm)
        t, n2 = glue.add_layer_norm(t, a1.to(t.dtype), self.norm2.weight, self.norm2.bias, self.norm2.eps, mm)
        tp = self.text_proj(text_seq)
        a2 = glue.multihead_attention(self.cross_attn, n2, tp)
        # norm3 (:561): the fp32 rows the router reads and - when the block runs its bf16 tensor-core path - their bf16 operand copy
        # from the same pass, so the block has no conversion kernel of its own
        bf16_block = t.is_cuda and getattr(self.moe, "takes_operand", False) and self.moe._precision_code() == 0
        if bf16_block:
            t, x_n, x_op = glue.add_layer_norm(t, a2.to(t.dtype), self.norm3.weight, self.norm3.bias, self.norm3.eps, torch.float32, with_operand=True)
        else:
            (t, x_n), x_op = glue.add_layer_norm(t, a2.to(t.dtype), self.norm3.weight, self.norm3.bias, self.norm3.eps, torch.float32), None
        x_spatial = t.reshape(B, H, W, C).permute(0, 3, 1, 2)
        x_norm = x_n.reshape(B, H, W, C).permute(0, 3, 1, 2)                # NCHW view of NHWC memory (:561)
        if getattr(self.moe, "fuses_residual", False):
            # B200 block: `x_spatial + moe_out` (:571) happens in the combine kernel, in x_spatial's dtype
            kw = {"x_operand": x_op} if x_op is not None else {}
            h, moe_kl, probs = self.moe(x_norm, w, annealing_factor, residual=x_spatial, **kw)
        else:
            moe_out, moe_kl, probs = self.moe(x_norm, w, annealing_factor)
            h = x_spatial + moe_out.to(x_spatial.dtype)
        if kl_losses is not None:
            kl_losses.append(moe_kl)
        return self.proj_out(h, w), probs


class GenerativeBlock(nn.Module):
    def __init__(self, in_channels, out_channels, text_dim=768, upsample=False, resolution=None, moe_kwargs=None):
        super().__init__()
        self.upsample = upsample
        self.conv_block = ConvolutionBlock(in_channels, out_channels, resolution=resolution)
        self.attn_block = AttentionBlock(out_channels, text_dim=text_dim, moe_kwargs=moe_kwargs)

    def forward(self, x, w, text_seq, kl_losses=None, annealing_factor=1.0):
        if self.upsample:
            # NHWC 2x bilinear kernels of the library (ATen: NCHW kernel 21 ms per call at batch 256, its NHWC kernel 70-140 us)
            x = glue.upsample2x_bilinear(x)
        return self.attn_block(self.conv_block(x, w), w, text_seq, kl_losses, annealing_factor)


class AuroraGenerator(nn.Module):
    """z [B,512], text [B,512] -> 16x16 image (+ 8x8 intermediate), summed KL of the 3 MoE routers, routing probs."""

    def __init__(self, latent_dim=LATENT_DIM, text_embedding_dim=TEXT_EMBEDDING_DIM, moe_kwargs=None):
        super().__init__()
        self.latent_dim, self.text_embedding_dim = latent_dim, text_embedding_dim
        td = text_embedding_dim
        self.text_projection = nn.Sequential(nn.Linear(td, td), nn.LayerNorm(td), nn.LeakyReLU(0.2, inplace=True), nn.Linear(td, td))
        self.mapping = nn.Sequential(nn.Linear(latent_dim + td, 512), nn.LeakyReLU(0.2, inplace=True),
                                     nn.Linear(512, 512), nn.LeakyReLU(0.2, inplace=True),
                                     nn.Linear(512, 512), nn.LeakyReLU(0.2, inplace=True), nn.Linear(512, 512))
        self.constant = nn.Parameter(torch.randn(1, 512, 4, 4))
        self.gen_block_4 = GenerativeBlock(512, 512, text_dim=td, upsample=False, resolution=4, moe_kwargs=moe_kwargs)
        self.gen_block_8 = GenerativeBlock(512, 256, text_dim=td, upsample=True, resolution=8, moe_kwargs=moe_kwargs)
        self.gen_block_16 = GenerativeBlock(256, 128, text_dim=td, upsample=True, resolution=16, moe_kwargs=moe_kwargs)
        self.to_rgb_8 = ModulatedConv(256, 3, 1)
        self.to_rgb_16 = ModulatedConv(128, 3, 1)

        self._use_checkpointing = False

    def enable_checkpointing(self):
        """Non-reentrant activation checkpointing of the three generative blocks (t2i_moe_gan.py:712-760).  The RNG
        state is preserved, so the recomputation redraws the same router eps."""
        self._use_checkpointing = True
        return self

    def disable_checkpointing(self):
        self._use_checkpointing = False
        return self

    def _run_block(self, block, x, w, text_seq, kl_losses, annealing_factor):
        if self._use_checkpointing and self.training and torch.is_grad_enabled():
            def fn(x_, w_, t_):
                kl = []
                out, probs = block(x_, w_, t_, kl, annealing_factor)
                return out, probs, kl[0]
            out, probs, kl0 = torch.utils.checkpoint.checkpoint(fn, x, w, text_seq, use_reentrant=False)
            kl_losses.append(kl0)
            return out, probs
        return block(x, w, text_seq, kl_losses, annealing_factor)

    def moe_layers(self) -> List[SparseMoE]:
        return [self.gen_block_4.attn_block.moe, self.gen_block_8.attn_block.moe, self.gen_block_16.attn_block.moe]

    def forward(self, z, text_embeddings, truncation_psi=0.7, return_routing=False, return_intermediate=False, annealing_factor=1.0):
        B = z.size(0)
        if text_embeddings.size(0) != B:
            if text_embeddings.size(0) != 1:
                raise ValueError(f"Batch size mismatch: z has batch size {B}, but text_embeddings has batch size {text_embeddings.size(0)}")
            text_embeddings = text_embeddings.repeat(B, 1)
        text_seq = self.text_projection(text_embeddings).unsqueeze(1)
        w = self.mapping(torch.cat([z, text_embeddings], dim=1))
        if truncation_psi < 1.0:                                            # truncation trick (:799-808)
            with torch.no_grad():
                mean_latent = self.mapping(torch.zeros(1, self.latent_dim + self.text_embedding_dim, device=z.device, dtype=z.dtype))
            w = mean_latent + truncation_psi * (w - mean_latent)
        kl_losses, routing = [], []
        x = self.constant.repeat(B, 1, 1, 1)
        x, p = self._run_block(self.gen_block_4, x, w, text_seq, kl_losses, annealing_factor); routing.append(p)
        x, p = self._run_block(self.gen_block_8, x, w, text_seq, kl_losses, annealing_factor); routing.append(p)
        x_8 = self.to_rgb_8(x, w)
        x, p = self._run_block(self.gen_block_16, x, w, text_seq, kl_losses, annealing_factor); routing.append(p)
        x_16 = self.to_rgb_16(x, w)
        kl = sum(kl_losses) if kl_losses else torch.zeros((), device=z.device)
        if return_routing and return_intermediate:
            return x_16, x_8, kl, routing
        if return_routing:
            return x_16, None, kl, routing
        if return_intermediate:
            return x_16, x_8, kl
        return x_16, kl


class AuroraDiscriminator(nn.Module):
    def __init__(self, text_embedding_dim=TEXT_EMBEDDING_DIM):
        super().__init__()
        self.text_projection = nn.Sequential(weight_norm(nn.Linear(text_embedding_dim, 128)), nn.LeakyReLU(0.2, inplace=True))
        self.conv_layers = nn.Sequential(weight_norm(nn.Conv2d(3, 128, 4, 2, 1)), nn.LeakyReLU(0.2, inplace=True),
                                         weight_norm(nn.Conv2d(128, 256, 4, 2, 1)), nn.LeakyReLU(0.2, inplace=True))
        self.output_layer = nn.Sequential(weight_norm(nn.Conv2d(256 + 128, 1, 4, 1, 0)))

    def features(self, img):
        return self.conv_layers(img)

    def head(self, f, text_embedding):
        c = self.output_layer[0]
        return text_conditioned_head(f, self.text_projection(text_embedding), c.weight, c.bias).reshape(-1)   # 16x16 fake -> B logits; 64x64 real -> 169*B

    def forward(self, img, text_embedding):
        return self.head(self.features(img), text_embedding)


def text_conditioned_head(f, t, weight, bias):
    """`conv2d(cat([f, t tiled over H x W], dim=1), weight, bias)` of the discriminator's output layer (t2i_moe_gan.py:899-907; a
    valid 4x4 convolution, no padding) without materialising the concatenation: the tiled text half is constant over every
    window, so its contribution is one [B,128] x [128,out] product with the window-summed weights, added to the convolution of
    the feature half.  Same function; the [B, 384, H, W] copy, its backward and a third of the convolution's channels disappear."""
    cf = f.shape[1]
    out = F.conv2d(f, weight[:, :cf], bias)
    tt = F.linear(t, weight[:, cf:].sum(dim=(2, 3)))                        # [B, out]
    return out + tt.to(out.dtype)[:, :, None, None]


# ---- losses (t2i_moe_gan.py:917-949) -------------------------------------------------------------------
def generator_adv_loss(fake_pred):
    return F.softplus(-fake_pred).mean()


def discriminator_adv_loss(real_pred, fake_pred, mismatched_pred):
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean() + F.softplus(mismatched_pred).mean()


def r1_penalty(real_pred, real_images, r1_gamma=10.0):
    """(gamma / 2) * E ||d D(real) / d real||^2  (:1282-1286)."""
    (grad_real,) = torch.autograd.grad(outputs=real_pred.sum(), inputs=real_images, create_graph=True)
    return (r1_gamma / 2) * grad_real.reshape(grad_real.size(0), -1).norm(2, dim=1).pow(2).mean()


def discriminator_real_pass(D: "AuroraDiscriminator", real, text, text_mismatched, r1_gamma=10.0):
    """The three quantities the D-step needs from the REAL batch (t2i_moe_gan.py:1278-1306): D(real, text),
    D(real, mismatched text) and the R1 penalty - same values and gradients as
        real_pred = D(real, text); r1 = r1_penalty(real_pred, real); mis = D(real, text_mismatched)
    with two re-formulations:
      * the convolutional features of `real` are computed once and shared by both heads (the reference recomputes them);
      * d real_pred.sum() / d real is written as an explicit adjoint pass of forward ops (transposed convolutions and
        LeakyReLU masks) instead of autograd.grad(create_graph=True): the gradient of the penalty w.r.t. the weights
        then runs through cuDNN's ordinary dgrad / wgrad kernels, not through the generic double-backward convolutions
        (whose [3,B,64,64] x [128,B,32,32] case alone took 1.4 ms at batch 256).  LeakyReLU is piecewise linear, so the
        masks are constants of the differentiation, exactly as in autograd's double backward."""
    c1, c2, c3 = D.conv_layers[0], D.conv_layers[2], D.output_layer[0]
    slope = D.conv_layers[1].negative_slope
    W1, W2, W3 = c1.weight, c2.weight, c3.weight        # weight_norm parametrisations: evaluate once
    a1 = glue.add_channel_bias(F.conv2d(real, W1, None, stride=2, padding=1), c1.bias)     # bias gradient by the library's column sum
    h1 = F.leaky_relu(a1, slope)
    a2 = glue.add_channel_bias(F.conv2d(h1, W2, None, stride=2, padding=1), c2.bias)
    f = F.leaky_relu(a2, slope)

    def head(txt):
        return text_conditioned_head(f, D.text_projection(txt), W3, c3.bias)

    out_real = head(text)
    real_pred, mismatched_pred = out_real.reshape(-1), head(text_mismatched).reshape(-1)
    # adjoint of real_pred.sum(): the seed is all-ones and identical for every sample
    seed = torch.ones(1, 1, out_real.shape[2], out_real.shape[3], device=real.device, dtype=out_real.dtype)
    d_f = F.conv_transpose2d(seed, W3)[:, : f.shape[1]]                          # [1,256,16,16]
    lrelu_adj = torch.ops.aten.leaky_relu_backward                               # grad * (x > 0 ? 1 : slope): one kernel, differentiable
    d_a2 = lrelu_adj(d_f.expand_as(a2), a2, slope, False)                        # [B,256,16,16]
    d_h1 = F.conv_transpose2d(d_a2, W2, stride=2, padding=1)
    d_a1 = lrelu_adj(d_h1, a1, slope, False)
    grad_real = F.conv_transpose2d(d_a1, W1, stride=2, padding=1)                # = d real_pred.sum() / d real
    r1 = (r1_gamma / 2) * grad_real.reshape(grad_real.size(0), -1).norm(2, dim=1).pow(2).mean()
    return real_pred, mismatched_pred, r1


def sample_images(generator, z, text_embeddings, truncation_psi=0.7):
    """sample_aurora_gan with tensor embeddings (:1672-1708): eval mode -> hard top-1 routing through the kernels."""
    was_training = generator.training
    generator.eval()
    with torch.no_grad():
        img, _ = generator(z, text_embeddings, truncation_psi=truncation_psi)
    generator.train(was_training)
    return img


def load_reference_checkpoint(generator: nn.Module, discriminator: Optional[nn.Module], ckpt: dict) -> None:
    """{'generator': sd, 'discriminator': sd} as saved by moegan/sagemaker_train.py:297-301.  The only key
    difference is the discriminator's legacy weight_norm naming (weight_g / weight_v)."""
    generator.load_state_dict(ckpt["generator"])
    if discriminator is not None and "discriminator" in ckpt:
        sd = {}
        for k, v in ckpt["discriminator"].items():
            if k.endswith("weight_g"):
                k = k[:-len("weight_g")] + "parametrizations.weight.original0"
            elif k.endswith("weight_v"):
                k = k[:-len("weight_v")] + "parametrizations.weight.original1"
            sd[k] = v
        discriminator.load_state_dict(sd)
