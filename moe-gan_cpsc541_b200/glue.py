"""Host side of the bandwidth-bound layers directly around the MoE block: autograd wrappers over
the C ABI entry points `moegan_layernorm_*`, `moegan_channel_scale_*`, `moegan_offset_sample_*` (csrc/glue.cu).

Each public function states the reference computation it replaces.  CUDA tensors always go through the library (a missing
`.so` raises); for CPU tensors the function evaluates that reference formula with plain torch ops - this is only the
definition the CPU arm / the pinning tests of the host model use, never a substitute for the kernels on a GPU.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

import warnings

from . import _cabi as C

_WARNED = set()


def _torch_path(fn: str, why: str) -> None:
    """A CUDA tensor whose shape the library's kernel does not cover goes through the plain torch op - never silently: one
    RuntimeWarning per (function, reason), so a model whose layers were expected on the kernels shows where they are not."""
    key = (fn, why)
    if key not in _WARNED:
        _WARNED.add(key)
        warnings.warn(f"moegan glue.{fn}: CUDA tensor handled by the torch op, not the sm_100a kernel ({why})", RuntimeWarning, stacklevel=3)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _is_bf16(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float32:
        return 0
    raise TypeError(f"moegan glue kernels take fp32 or bf16 activations, got {t.dtype}")


# --------------------------------------------------------------------------------------------------
# LayerNorm (t2i_moe_gan.py:513-515 -> :541, :553, :561)
# --------------------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype, with_operand=False):
        d = x.shape[-1]
        x2 = x.reshape(-1, d)
        if x2.dtype not in (torch.float32, torch.bfloat16):
            x2 = x2.float()
        x2 = x2.contiguous()
        rows = x2.shape[0]
        wf, bf = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        y = torch.empty(rows, d, device=x.device, dtype=out_dtype)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        y_op = None
        if with_operand:      # fp32 rows + their bf16 GEMM-operand copy in ONE pass (norm3 -> MoE block)
            y_op = torch.empty(rows, d, device=x.device, dtype=torch.bfloat16)
            C.check(C.lib().moegan_layernorm_fwd_dual(C.ptr(x2), _is_bf16(x2), C.ptr(wf), C.ptr(bf), rows, d, float(eps), C.ptr(y), C.ptr(y_op),
                                                      C.ptr(mean), C.ptr(rstd), _stream()), "moegan_layernorm_fwd_dual")
        else:
            C.check(C.lib().moegan_layernorm_fwd(C.ptr(x2), _is_bf16(x2), C.ptr(wf), C.ptr(bf), rows, d, float(eps), C.ptr(y), _is_bf16(y),
                                                 C.ptr(mean), C.ptr(rstd), _stream()), "moegan_layernorm_fwd")
        ctx.save_for_backward(x2, wf, mean, rstd)
        ctx.x_shape, ctx.x_dtype, ctx.w_dtype = x.shape, x.dtype, weight.dtype
        ctx.with_operand = with_operand
        if with_operand:
            y_op = y_op.view(x.shape)
            ctx.mark_non_differentiable(y_op)     # a rounded copy of y: the gradient flows through y
            return y.view(x.shape), y_op
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, gy, *unused):
        x2, wf, mean, rstd = ctx.saved_tensors
        rows, d = x2.shape
        g = gy.reshape(rows, d)
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.float()
        g = g.contiguous()
        lib = C.lib()
        dx = torch.empty_like(x2)                       # in x's storage dtype, straight from the kernel
        dgb = torch.empty(2, d, device=g.device, dtype=torch.float32)
        nbytes = int(lib.moegan_layernorm_bwd_scratch_bytes(rows, d))
        scratch = torch.empty(nbytes, device=g.device, dtype=torch.uint8)
        C.check(lib.moegan_layernorm_bwd(C.ptr(g), _is_bf16(g), C.ptr(x2), _is_bf16(x2), C.ptr(wf), C.ptr(mean), C.ptr(rstd), rows, d,
                                         C.ptr(dx), C.ptr(dgb), C.ptr(scratch), nbytes, _stream()), "moegan_layernorm_bwd")
        return dx.view(ctx.x_shape).to(ctx.x_dtype), dgb[0].to(ctx.w_dtype), dgb[1].to(ctx.w_dtype), None, None, None


class _AddLayerNormFn(torch.autograd.Function):
    """(t, a) -> (t + a, LayerNorm(t + a)[, bf16 copy]) in one kernel; backward: d(t + a) = LayerNorm-backward(g_n) + g_sum, also fused."""

    @staticmethod
    def forward(ctx, x, addend, weight, bias, eps, out_dtype, with_operand):
        d = x.shape[-1]
        x2, a2 = x.reshape(-1, d).contiguous(), addend.reshape(-1, d).contiguous()
        rows = x2.shape[0]
        wf, bf = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        s = torch.empty_like(x2)
        y = torch.empty(rows, d, device=x.device, dtype=out_dtype)
        y_op = torch.empty(rows, d, device=x.device, dtype=torch.bfloat16) if with_operand else None
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        C.check(C.lib().moegan_add_layernorm_fwd(C.ptr(x2), C.ptr(a2), _is_bf16(x2), C.ptr(wf), C.ptr(bf), rows, d, float(eps), C.ptr(s), C.ptr(y),
                                                 _is_bf16(y), C.ptr(y_op) if with_operand else None, C.ptr(mean), C.ptr(rstd), _stream()),
                "moegan_add_layernorm_fwd")
        ctx.save_for_backward(s, wf, mean, rstd)
        ctx.x_shape, ctx.w_dtype = x.shape, weight.dtype
        outs = (s.view(x.shape), y.view(x.shape))
        if with_operand:
            y_op = y_op.view(x.shape)
            ctx.mark_non_differentiable(y_op)
            outs = outs + (y_op,)
        return outs

    @staticmethod
    def backward(ctx, g_sum, gy, *unused):
        s, wf, mean, rstd = ctx.saved_tensors
        rows, d = s.shape
        lib = C.lib()
        if gy is None:                                   # the normalised rows were not used: only the residual path carries gradient
            return g_sum, g_sum, None, None, None, None, None
        g = gy.reshape(rows, d)
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.float()
        g = g.contiguous()
        dx = torch.empty_like(s)
        dgb = torch.empty(2, d, device=g.device, dtype=torch.float32)
        nbytes = int(lib.moegan_layernorm_bwd_scratch_bytes(rows, d))
        scratch = torch.empty(nbytes, device=g.device, dtype=torch.uint8)
        if g_sum is None:
            C.check(lib.moegan_layernorm_bwd(C.ptr(g), _is_bf16(g), C.ptr(s), _is_bf16(s), C.ptr(wf), C.ptr(mean), C.ptr(rstd), rows, d,
                                             C.ptr(dx), C.ptr(dgb), C.ptr(scratch), nbytes, _stream()), "moegan_layernorm_bwd")
        else:
            ga = g_sum.reshape(rows, d).to(s.dtype).contiguous()
            C.check(lib.moegan_layernorm_bwd_add(C.ptr(g), _is_bf16(g), C.ptr(s), _is_bf16(s), C.ptr(wf), C.ptr(mean), C.ptr(rstd), rows, d,
                                                 C.ptr(ga), C.ptr(dx), C.ptr(dgb), C.ptr(scratch), nbytes, _stream()), "moegan_layernorm_bwd_add")
        dxv = dx.view(ctx.x_shape)
        return dxv, dxv, dgb[0].to(ctx.w_dtype), dgb[1].to(ctx.w_dtype), None, None, None


def add_layer_norm(x: torch.Tensor, addend: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5, out_dtype=torch.float32,
                   with_operand: bool = False):
    """`t = x + addend; n = nn.LayerNorm(d)(t)` -> `(t, n)` (with_operand=True: `(t, n fp32, n rounded to bf16)`): the residual adds in
    front of norm2 / norm3 of AttentionBlock (t2i_moe_gan.py:545 + :553, :557 + :561).  On CUDA, for d = 128 / 256 / 512 and operands of
    one dtype (fp32 / bf16), one kernel forward and one backward; anything else is the add followed by `layer_norm`."""
    d = x.shape[-1]
    if (not _NO_ADD_LN and x.is_cuda and addend.is_cuda and d in (128, 256, 512) and x.dtype == addend.dtype and x.dtype in (torch.float32, torch.bfloat16)
            and x.shape == addend.shape and (not with_operand or out_dtype == torch.float32)):
        return _AddLayerNormFn.apply(x, addend, weight, bias, eps, out_dtype, with_operand)
    t = x + addend
    n = layer_norm(t, weight, bias, eps, out_dtype, with_operand)
    return (t,) + n if with_operand else (t, n)


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5, out_dtype=torch.float32, with_operand: bool = False):
    """`nn.LayerNorm(d)(x)` over the last dimension.  Result fp32 (what autocast's LayerNorm returns), or bf16 when the
    only consumer is an autocast matmul that would round it to bf16 first thing anyway (same values, one pass less).
    with_operand=True (fp32 result only): returns (y fp32, y rounded to bf16) from one pass - what norm3 hands the MoE block."""
    if not x.is_cuda:
        y = F.layer_norm(x.float(), (x.shape[-1],), weight.float(), bias.float(), eps).to(out_dtype)
        return (y, y.detach().to(torch.bfloat16)) if with_operand else y
    if with_operand:
        if out_dtype != torch.float32:
            raise ValueError("layer_norm(with_operand=True) returns the fp32 rows plus their bf16 copy: out_dtype must be float32")
        return _LayerNormFn.apply(x, weight, bias, eps, out_dtype, True)
    return _LayerNormFn.apply(x, weight, bias, eps, out_dtype)


# --------------------------------------------------------------------------------------------------
# per-(sample, channel) scale + LeakyReLU  (ModulatedConv modulate / demodulate, t2i_moe_gan.py:158-186; MTM activation :243)
# --------------------------------------------------------------------------------------------------
def _nhwc(x: torch.Tensor) -> torch.Tensor:
    return x.contiguous(memory_format=torch.channels_last)


class _ChannelScaleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s, slope):
        B, Cn, H, W = x.shape
        xc = _nhwc(x if x.dtype in (torch.float32, torch.bfloat16) else x.float())
        sc = s.detach()
        sc = (sc if sc.dtype in (torch.float32, torch.bfloat16) else sc.float()).contiguous()
        out = torch.empty_like(xc)                      # keeps the channels-last strides
        C.check(C.lib().moegan_channel_scale_fwd(C.ptr(xc), _is_bf16(xc), C.ptr(sc), _is_bf16(sc), B, H * W, Cn, float(slope), C.ptr(out),
                                                 _stream()), "moegan_channel_scale_fwd")
        ctx.save_for_backward(xc, sc)
        ctx.slope, ctx.x_dtype, ctx.s_dtype = float(slope), x.dtype, s.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        xc, sc = ctx.saved_tensors
        B, Cn, H, W = xc.shape
        gc = _nhwc(g.to(xc.dtype))
        lib = C.lib()
        dx = torch.empty_like(xc)
        ds = torch.empty_like(sc)                       # in s's dtype, straight from the kernel
        nbytes = int(lib.moegan_channel_scale_bwd_scratch_bytes(B, H * W, Cn))
        scratch = torch.empty(nbytes, device=xc.device, dtype=torch.uint8)
        C.check(lib.moegan_channel_scale_bwd(C.ptr(gc), C.ptr(xc), _is_bf16(xc), C.ptr(sc), _is_bf16(sc), B, H * W, Cn, ctx.slope, C.ptr(dx),
                                             C.ptr(ds), C.ptr(scratch), nbytes, _stream()), "moegan_channel_scale_bwd")
        return dx.to(ctx.x_dtype), ds.to(ctx.s_dtype), None


def channel_scale(x: torch.Tensor, s: torch.Tensor, slope: float = 1.0) -> torch.Tensor:
    """`F.leaky_relu(x * s[:, :, None, None], slope)` for x [B,C,H,W], s [B,C]; slope = 1.0 means no activation."""
    if x.is_cuda and x.dim() == 4 and x.shape[1] % 128 == 0:
        return _ChannelScaleFn.apply(x, s, slope)
    if x.is_cuda:
        _torch_path("channel_scale", f"channels {x.shape[1]} not a multiple of 128 (e.g. the 3-channel to_rgb layers)")
    y = x * s.to(x.dtype)[:, :, None, None]            # CPU definition; also tiny channel counts (to_rgb: 3 channels)
    return y if slope == 1.0 else F.leaky_relu(y, slope)


# --------------------------------------------------------------------------------------------------
# demodulation coefficients of ModulatedConv (t2i_moe_gan.py:165-170)
# --------------------------------------------------------------------------------------------------
class _DemodCoefFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, weight, eps, cache):
        B, Cin = s.shape
        Cout, _, KH, KW = weight.shape
        sc = s.detach()
        sc = (sc if sc.dtype in (torch.float32, torch.bfloat16) else sc.float()).contiguous()
        w = weight.detach()
        so, si, sh, sw = w.stride()
        # wsq depends on the weights only: the generator runs twice per train step between optimizer steps (the D-step's no_grad pass and
        # the G-step), the second pass reuses the first one's.  Key as for the expert-weight cache (functional.WeightCache): storage,
        # autograd version, and the generation counter the trainer bumps after every CUDA-graph replay.
        # While a CUDA graph is being captured only entries made inside the SAME capture are served (a replay must recompute from the
        # then-current weights; an eager entry would be baked into the graph as a stale pointer), and capture entries are never served
        # to eager calls.
        wsq, compute = None, 1
        if cache is not None:
            from .functional import weight_generation
            capturing = torch.cuda.is_current_stream_capturing()
            cid = int(C.lib().moegan_stream_capture_id(C.c_void_p(torch.cuda.current_stream().cuda_stream))) if capturing else 0
            slot = "capture" if capturing else "eager"
            key = (cid, w.data_ptr(), weight._version, weight_generation(), tuple(w.shape), tuple(w.stride()))
            entry = cache.get(slot)
            if entry is not None and entry[0] == key and (not capturing or cid > 0):
                wsq, compute = entry[1], 0
        if wsq is None:
            wsq = torch.empty((Cout, Cin), device=w.device, dtype=torch.float32)
        d = torch.empty((B, Cout), device=w.device, dtype=torch.float32)
        C.check(C.lib().moegan_demod_coef_fwd(C.ptr(sc), _is_bf16(sc), C.ptr(w), so, si, sh, sw, B, Cin, Cout, KH, KW, float(eps), compute,
                                              C.ptr(wsq), C.ptr(d), _stream()), "moegan_demod_coef_fwd")
        if cache is not None and compute:
            cache[slot] = (key, wsq)
        ctx.save_for_backward(sc, w, wsq, d)
        ctx.s_dtype = s.dtype
        return d

    @staticmethod
    def backward(ctx, dd):
        sc, w, wsq, d = ctx.saved_tensors
        B, Cin = sc.shape
        Cout, _, KH, KW = w.shape
        so, si, sh, sw = w.stride()
        ddc = dd.float().contiguous()
        ds = torch.empty_like(sc)
        dw = torch.empty_like(w)                      # same strides as the weight (contiguous or channels-last); every element is written
        C.check(C.lib().moegan_demod_coef_bwd(C.ptr(ddc), C.ptr(d), C.ptr(sc), _is_bf16(sc), C.ptr(w), so, si, sh, sw, C.ptr(wsq), B, Cin, Cout,
                                              KH, KW, C.ptr(ds), C.ptr(dw), _stream()), "moegan_demod_coef_bwd")
        return ds.to(ctx.s_dtype), dw, None, None


_TORCH_DEMOD = __import__("os").environ.get("MOEGAN_TORCH_DEMOD", "0") == "1"
_NO_ADD_LN = __import__("os").environ.get("MOEGAN_NO_ADD_LN", "0") == "1"      # A/B switch: residual add and LayerNorm as two kernels


def demod_coef_reference(s: torch.Tensor, weight: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """The definition in torch ops (fp32): d[b,o] = rsqrt(sum_i s[b,i]^2 * sum_taps w[o,i]^2 + eps)."""
    wsq = weight.float().pow(2).sum(dim=(2, 3))
    with torch.autocast(s.device.type, enabled=False):
        return torch.rsqrt(s.float().pow(2) @ wsq.t() + eps)


def demod_coef(s: torch.Tensor, weight: torch.Tensor, eps: float = 1e-8, cache: dict = None) -> torch.Tensor:
    """Demodulation coefficients [B, Cout] (fp32) of a ModulatedConv with modulation s [B, Cin] and weight [Cout, Cin, KH, KW]: what the
    reference computes as `rsqrt(sum((weight * s)^2) + 1e-8)` per sample and output channel (t2i_moe_gan.py:165-170).  `cache`: a dict
    owned by the calling module in which the weight-only factor sum_taps(w^2) is kept between calls with unchanged weights."""
    if _TORCH_DEMOD:                     # A/B switch for bench runs: the torch-op definition instead of the kernels
        return demod_coef_reference(s, weight, eps)
    if s.is_cuda and weight.is_cuda and weight.dtype == torch.float32 and weight.dim() == 4 and s.dim() == 2:
        dense = weight.is_contiguous() or weight.is_contiguous(memory_format=torch.channels_last)
        if dense:
            return _DemodCoefFn.apply(s, weight, eps, cache)
        _torch_path("demod_coef", f"weight strides {weight.stride()} are neither contiguous nor channels-last")
    elif s.is_cuda:
        _torch_path("demod_coef", f"weight dtype {weight.dtype} / shape {tuple(weight.shape)}")
    return demod_coef_reference(s, weight, eps)


# --------------------------------------------------------------------------------------------------
# Linear over all tokens with a fast bias gradient; nn.MultiheadAttention evaluated through it (t2i_moe_gan.py:516-517,:545,:557)
# --------------------------------------------------------------------------------------------------
def colsum(x2: torch.Tensor) -> torch.Tensor:
    """fp32 column sums of a contiguous [rows, cols] fp32 / bf16 CUDA matrix (deterministic)."""
    rows, cols = x2.shape
    lib = C.lib()
    out = torch.empty(cols, device=x2.device, dtype=torch.float32)
    nbytes = int(lib.moegan_colsum_scratch_bytes(rows, cols))
    scratch = torch.empty(nbytes, device=x2.device, dtype=torch.uint8)
    C.check(lib.moegan_colsum(C.ptr(x2), _is_bf16(x2), rows, cols, C.ptr(out), C.ptr(scratch), nbytes, _stream()), "moegan_colsum")
    return out


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b as autocast would compute it (operands in the autocast dtype, cuBLAS), with the bias gradient
    `grad_y.sum(0)` - a [T, d] -> [d] reduction that ATen runs at a few percent of HBM bandwidth - by the library's kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
        xc, wc = x.to(dt), weight.to(dt)
        with torch.autocast("cuda", enabled=False):
            y = F.linear(xc, wc, bias.to(dt))
        ctx.save_for_backward(xc, wc)
        ctx.dtypes = (x.dtype, weight.dtype, bias.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        xc, wc = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        if g2.dtype not in (torch.float32, torch.bfloat16):
            g2 = g2.float()
        g2 = g2.contiguous()
        gm = g2 if g2.dtype == wc.dtype else g2.to(wc.dtype)
        x2 = xc.reshape(-1, xc.shape[-1])
        gx = (gm @ wc).view(xc.shape).to(ctx.dtypes[0])
        gw = (gm.t() @ x2).to(ctx.dtypes[1])
        gb = colsum(g2).to(ctx.dtypes[2])
        return gx, gw, gb


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """`F.linear(x, weight, bias)`; on CUDA the bias gradient comes from `moegan_colsum`."""
    if x.is_cuda and bias is not None and weight.shape[0] % 4 == 0 and x.dtype in (torch.float32, torch.bfloat16):
        return _LinearFn.apply(x, weight, bias)
    if x.is_cuda:
        _torch_path("linear", f"bias {'absent' if bias is None else 'present'}, out features {weight.shape[0]}, dtype {x.dtype}")
    return F.linear(x, weight, bias)


class _ChannelBiasFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, bias):
        ctx.bias_dtype = bias.dtype
        return y + bias.to(y.dtype).view(1, -1, 1, 1)

    @staticmethod
    def backward(ctx, g):
        gc = _nhwc(g if g.dtype in (torch.float32, torch.bfloat16) else g.float())
        gb = colsum(gc.permute(0, 2, 3, 1).reshape(-1, gc.shape[1]))          # NHWC memory = a [B*H*W, C] matrix
        return g, gb.to(ctx.bias_dtype)


def add_channel_bias(y: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """`y + bias[None, :, None, None]` (the bias of a convolution evaluated without one); on CUDA its gradient
    `grad.sum((0, 2, 3))` comes from `moegan_colsum` over the NHWC view instead of ATen's strided reduction."""
    if y.is_cuda and y.dim() == 4 and y.shape[1] % 4 == 0:
        return _ChannelBiasFn.apply(y, bias)
    if y.is_cuda:
        _torch_path("add_channel_bias", f"shape {tuple(y.shape)}")
    return y + bias.to(y.dtype).view(1, -1, 1, 1)


def multihead_attention(m: torch.nn.MultiheadAttention, query: torch.Tensor, key: torch.Tensor) -> torch.Tensor:
    """`m(query, key, key, need_weights=False)[0]` for a batch_first nn.MultiheadAttention without masks or dropout - the same
    packed in-projection, scaled_dot_product_attention and out-projection `F.multi_head_attention_forward` performs, with the
    token-wide projections going through `linear` above.  Parameters (and state_dict keys) stay those of the module."""
    d, h = m.embed_dim, m.num_heads
    W, bvec = m.in_proj_weight, m.in_proj_bias
    if query is key:
        q, k, v = linear(query, W, bvec).chunk(3, dim=-1)
    else:
        q = linear(query, W[:d], bvec[:d])
        k, v = F.linear(key, W[d:], bvec[d:]).chunk(2, dim=-1)
    B, L, S = q.shape[0], q.shape[1], k.shape[1]
    q = q.reshape(B, L, h, d // h).transpose(1, 2)
    k = k.reshape(B, S, h, d // h).transpose(1, 2)
    v = v.reshape(B, S, h, d // h).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, d)
    return linear(o, m.out_proj.weight, m.out_proj.bias)


# --------------------------------------------------------------------------------------------------
# 2x bilinear upsampling at the head of a GenerativeBlock (t2i_moe_gan.py:646-648)
# --------------------------------------------------------------------------------------------------
class _Upsample2xFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        B, Cn, H, W = x.shape
        xc = _nhwc(x if x.dtype in (torch.float32, torch.bfloat16) else x.float())
        out = torch.empty((B, Cn, 2 * H, 2 * W), device=x.device, dtype=xc.dtype, memory_format=torch.channels_last)
        C.check(C.lib().moegan_upsample2x_fwd(C.ptr(xc), _is_bf16(xc), B, H, W, Cn, C.ptr(out), _stream()), "moegan_upsample2x_fwd")
        ctx.in_shape, ctx.x_dtype = (B, Cn, H, W), x.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        B, Cn, H, W = ctx.in_shape
        gc = _nhwc(g if g.dtype in (torch.float32, torch.bfloat16) else g.float())
        dx = torch.empty((B, Cn, H, W), device=g.device, dtype=gc.dtype, memory_format=torch.channels_last)
        C.check(C.lib().moegan_upsample2x_bwd(C.ptr(gc), _is_bf16(gc), B, H, W, Cn, C.ptr(dx), _stream()), "moegan_upsample2x_bwd")
        return dx.to(ctx.x_dtype)


def upsample2x_bilinear(x: torch.Tensor) -> torch.Tensor:
    """`F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)` for x [B,C,H,W]; on CUDA (C a multiple of 8) through
    the library's NHWC kernels - same tap formula and evaluation order as ATen, deterministic gather in the backward."""
    if x.is_cuda and x.dim() == 4 and x.shape[1] % 8 == 0 and x.dtype in (torch.float32, torch.bfloat16):
        return _Upsample2xFn.apply(x)
    if x.is_cuda:
        _torch_path("upsample2x_bilinear", f"shape {tuple(x.shape)}, dtype {x.dtype}")
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


# --------------------------------------------------------------------------------------------------
# learned-offset bilinear resampling (ModulatedTransformationModule, t2i_moe_gan.py:219-240)
# --------------------------------------------------------------------------------------------------
class _OffsetSampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, off_raw):
        B, Cn, H, W = x.shape
        xc = _nhwc(x if x.dtype in (torch.float32, torch.bfloat16) else x.float())
        off = off_raw.detach().permute(0, 2, 3, 1).float().contiguous()          # [B,H,W,2]: (x offset, y offset)
        out = torch.empty_like(xc)
        C.check(C.lib().moegan_offset_sample_fwd(C.ptr(xc), _is_bf16(xc), C.ptr(off), B, H, W, Cn, C.ptr(out), _stream()),
                "moegan_offset_sample_fwd")
        ctx.save_for_backward(xc, off)
        ctx.x_dtype, ctx.off_dtype = x.dtype, off_raw.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        xc, off = ctx.saved_tensors
        B, Cn, H, W = xc.shape
        gc = _nhwc(g.to(xc.dtype))
        dx = torch.empty_like(xc)
        doff = torch.empty_like(off)
        C.check(C.lib().moegan_offset_sample_bwd(C.ptr(gc), C.ptr(xc), _is_bf16(xc), C.ptr(off), B, H, W, Cn, C.ptr(dx), C.ptr(doff), _stream()),
                "moegan_offset_sample_bwd")
        return dx.to(ctx.x_dtype), doff.permute(0, 3, 1, 2).to(ctx.off_dtype)


def offset_sample_reference(x: torch.Tensor, off_raw: torch.Tensor) -> torch.Tensor:
    """The reference computation in torch ops (t2i_moe_gan.py:222-240)."""
    B, _, H, W = x.shape
    off = off_raw.permute(0, 2, 3, 1) * 0.05
    gy, gx = torch.meshgrid(torch.linspace(-1, 1, H, device=x.device), torch.linspace(-1, 1, W, device=x.device), indexing="ij")
    grid = (torch.stack((gx, gy), dim=2).unsqueeze(0) + off.float()).clamp(-1, 1)
    return F.grid_sample(x, grid.to(x.dtype), mode="bilinear", align_corners=False)


def offset_sample(x: torch.Tensor, off_raw: torch.Tensor) -> torch.Tensor:
    """x [B,C,H,W] resampled at meshgrid + 0.05 * off_raw (off_raw [B,2,H,W] = offset_net(x): channel 0 = x, 1 = y)."""
    if x.is_cuda and x.shape[1] % 128 == 0 and x.shape[2] * x.shape[3] <= 256:
        return _OffsetSampleFn.apply(x, off_raw)
    if x.is_cuda:
        _torch_path("offset_sample", f"channels {x.shape[1]}, {x.shape[2]}x{x.shape[3]} pixels (kernel: C % 128 == 0, <= 256 pixels)")
    return offset_sample_reference(x, off_raw)
