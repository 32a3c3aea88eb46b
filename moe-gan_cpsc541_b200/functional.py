"""autograd wrappers over the C ABI: the whole MoE block (router -> dispatch -> experts -> combine, plus
KL and balance-load reductions) as ONE torch.autograd.Function, and the balance loss on the fused load.

Reference path replaced: SparseMoE.forward + BayesianRouter.forward / kl_divergence + SparseExpertFFN.forward
(moegan/t2i_moe_gan.py:249-491) and AuroraGANLoss.moe_balance_loss (:951-1000).
PyTorch is used here only for device memory, the current stream and autograd bookkeeping: each direction is
ONE FFI call (moegan_block_fwd / moegan_block_bwd) working out of ONE workspace allocation.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import torch

from . import _cabi as C

HID = C.ROUTER_HID


@dataclass(frozen=True)
class MoEConfig:
    B: int
    tokens_per_sample: int
    d: int
    Dt: int
    E: int
    k: int
    dense: bool          # k == E, gates are the soft probs (reference training path)
    hard: bool           # eval: one-hot probs, top-1, gate 1.0
    precision: int       # 0 bf16, 1 fp32 (3-pass split)
    sampled: bool        # weights sampled from (mu, rho, eps) (training) vs mean weights (eval)
    annealing_factor: float
    need_backward: bool

    @property
    def T(self) -> int:
        return self.B * self.tokens_per_sample

    def shape(self) -> C.Shape:
        return C.make_shape(self.T, self.d, self.Dt, self.E, self.k, self.B, self.tokens_per_sample,
                            self.precision, self.dense, self.hard)


class _Plan:
    """Per-configuration constants (shape struct, workspace sizes, layout), computed once."""
    __slots__ = ("shape", "fwd_bytes", "fwd_bytes_nobwd", "bwd_bytes", "layout", "layout_nobwd")


_PLANS: Dict[tuple, _Plan] = {}


def _plan_for(cfg: MoEConfig) -> _Plan:
    key = (cfg.B, cfg.tokens_per_sample, cfg.d, cfg.Dt, cfg.E, cfg.k, cfg.dense, cfg.hard, cfg.precision)
    p = _PLANS.get(key)
    if p is None:
        lib = C.lib()
        p = _Plan()
        p.shape = cfg.shape()
        p.layout, p.layout_nobwd = C.BlockLayout(), C.BlockLayout()
        p.fwd_bytes = int(lib.moegan_block_fwd_workspace(p.shape, 1, ctypes.byref(p.layout)))
        p.fwd_bytes_nobwd = int(lib.moegan_block_fwd_workspace(p.shape, 0, ctypes.byref(p.layout_nobwd)))
        p.bwd_bytes = int(lib.moegan_block_bwd_workspace(p.shape))
        if min(p.fwd_bytes, p.fwd_bytes_nobwd, p.bwd_bytes) < 0:
            raise RuntimeError("moegan: unsupported shape: " + C.last_error())
        _PLANS[key] = p
    return p


def set_fused_ffn(enabled: Optional[bool]) -> None:
    """Debug / bench / tests: False = forward-only block calls use the separate FFN1 / FFN2 / combine kernels instead of the fused
    FFN kernel; None = library default.  Drops the cached workspace sizes (the forward layout depends on the choice)."""
    C.check(C.lib().moegan_debug_set_fused_ffn(-1 if enabled is None else (1 if enabled else 0)), "moegan_debug_set_fused_ffn")
    _PLANS.clear()


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"moegan: '{name}' must be a CUDA tensor - the B200 path has no CPU fallback")


_WEIGHT_GENERATION = 0


def weight_generation() -> int:
    return _WEIGHT_GENERATION


def bump_weight_generation() -> None:
    """Invalidate every WeightCache.  Call after anything that changes expert weights WITHOUT bumping `Parameter._version` -
    i.e. a CUDA-graph replay whose captured optimizer step writes the parameters from inside the graph (MoEGANTrainer.step does)."""
    global _WEIGHT_GENERATION
    _WEIGHT_GENERATION += 1


class WeightCache:
    """bf16 operand copies of the expert weights, rebuilt only when a parameter changed (optimizer step).

    The reference re-casts every nn.Linear weight on every autocast forward; here both generator forwards
    of a train step (D-step under no_grad, G-step with grad) share one conversion.  While a CUDA graph is
    being captured the conversion is enqueued once per (capture, parameter versions): a replay must re-convert the
    then-current weights, but the forward passes of one captured step still share it.

    Validity key = (precision, d, E, weight generation, (data_ptr, _version) of every parameter).  `_version` does not move
    when a captured optimizer step updates the weights during graph.replay(), hence the generation counter
    (`bump_weight_generation`).  A key is only recorded by `commit()` AFTER the conversion has been enqueued successfully,
    so a failed / empty (T == 0) block call never leaves uninitialised buffers marked valid.
    """

    def __init__(self):
        self.key = None
        self.bufs = None
        self.capture_bufs = None
        self.capture_key = None
        self._pending = None

    def get(self, cfg: MoEConfig, w1: Sequence[torch.Tensor], b1, w2, b2):
        """Returns (buffers, stale).  When `stale`, the caller passes the parameter pointer tables to moegan_block_fwd,
        which converts on a side stream overlapped with the router kernels, and then calls commit()."""
        capturing = torch.cuda.is_current_stream_capturing()
        key = (cfg.precision, cfg.d, cfg.E, _WEIGHT_GENERATION,
               tuple((p.data_ptr(), p._version) for group in (w1, b1, w2, b2) for p in group))
        self._pending = None
        if not capturing and key == self.key and self.bufs is not None:
            return self.bufs, False
        ckey = None
        if capturing:
            cid = int(C.lib().moegan_stream_capture_id(C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            ckey = (cid, key)
            if cid > 0 and ckey == self.capture_key and self.capture_bufs is not None:
                return self.capture_bufs, False
        dev = w1[0].device
        S = 2 if cfg.precision == 1 else 1
        hid = 4 * cfg.d
        bufs = (torch.empty((cfg.E * hid, S * cfg.d), dtype=torch.bfloat16, device=dev),
                torch.empty((cfg.E, hid), dtype=torch.float32, device=dev),
                torch.empty((cfg.E * cfg.d, S * hid), dtype=torch.bfloat16, device=dev),
                torch.empty((cfg.E, cfg.d), dtype=torch.float32, device=dev))
        self._pending = (capturing, ckey if capturing else key, bufs)
        return bufs, True

    def commit(self) -> None:
        """The conversion requested by the last stale get() has been enqueued: its buffers are valid from now on."""
        if self._pending is None:
            return
        capturing, key, bufs = self._pending
        if capturing:
            self.capture_bufs, self.capture_key = bufs, key
        else:
            self.key, self.bufs = key, bufs
        self._pending = None

    def invalidate(self) -> None:
        self.key = self.bufs = self.capture_key = self.capture_bufs = self._pending = None


def _dp(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_I32_VIEWS = ("topk_idx", "chunk_counts", "plan", "chunk_base", "row_token", "token_rows")


def workspace_views(ws: torch.Tensor, layout: "C.BlockLayout", cfg: MoEConfig) -> dict:
    """Typed views of the routing state inside a forward workspace (tests / diagnostics only)."""
    T, E, k, m_cap = cfg.T, cfg.E, cfg.k, C.m_cap(cfg.T, cfg.E, cfg.k, cfg.dense)
    shapes = {"raw": (T, E), "gates": (T, k), "topk_idx": (T, k), "token_rows": (T, k), "row_token": (m_cap,),
              "row_gate": (m_cap,), "plan": (int(C.lib().moegan_plan_ints(m_cap)),)}
    out = {}
    for name, shp in shapes.items():
        off = getattr(layout, name)
        if off < 0:
            out[name] = None
            continue
        n = 1
        for v in shp:
            n *= v
        dt = torch.int32 if name in _I32_VIEWS else torch.float32
        out[name] = ws[off:off + 4 * n].view(dt).view(shp)
    return out


class MoEBlockFunction(torch.autograd.Function):
    """(x_tok [T,d], w [B,Dt], eps x3, router params x7, residual [T,d] | None, x_op bf16 [T,d] | None, expert params x4E)
    -> (out [T,d], kl [], probs [T,E], load [E]).  With a residual, out = residual + moe(x) in the residual's dtype (the
    `x_spatial + moe_out` of AttentionBlock, t2i_moe_gan.py:571, done in the combine kernel's epilogue)."""

    keep_last = False   # tests / diagnostics: keep (workspace, layout, cfg) of the most recent forward alive for last_aux()
    _last = None

    @staticmethod
    def last_aux() -> dict:
        if MoEBlockFunction._last is None:
            raise RuntimeError("MoEBlockFunction.last_aux(): set MoEBlockFunction.keep_last = True before the forward")
        ws, layout, cfg = MoEBlockFunction._last
        return workspace_views(ws, layout, cfg)

    @staticmethod
    def forward(ctx, cfg: MoEConfig, wcache: WeightCache, x_tok, w, eps_f, eps_t, eps_c,
                mu_f, rho_f, mu_t, rho_t, mu_c, rho_c, temperature, residual, x_op, *expert_params):
        lib = C.lib()
        E, d, T = cfg.E, cfg.d, cfg.T
        _require_cuda(x_tok, "x"); _require_cuda(w, "w"); _require_cuda(mu_f, "router.feature_mu")
        dev = x_tok.device
        f32 = torch.float32
        x_tok = x_tok.contiguous().float()
        w = w.contiguous().float()
        # the kernels read parameters through raw pointers as contiguous fp32 on x's device: anything else (a model after
        # .half() / .bfloat16(), a parameter on another GPU, a strided view) would be misread silently - refuse it here
        names = ("router.feature_mu", "router.feature_rho", "router.text_mu", "router.text_rho", "router.combined_mu", "router.combined_rho",
                 "router.temperature")
        for n, p in list(zip(names, (mu_f, rho_f, mu_t, rho_t, mu_c, rho_c, temperature))) + \
                [(f"experts.{i // 4}.net.{(0, 0, 2, 2)[i % 4]}.{('weight', 'bias')[i % 2]}", p) for i, p in enumerate(expert_params)]:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise RuntimeError(f"moegan: parameter '{n}' must be a contiguous float32 tensor on {dev} "
                                   f"(got {p.dtype}, contiguous={p.is_contiguous()}, {p.device}); keep the MoE parameters in fp32 - autocast "
                                   "selects the bf16 tensor-core path, the module's storage stays fp32")
        for n, e_ in (("eps_f", eps_f), ("eps_t", eps_t), ("eps_c", eps_c)):
            if e_ is not None and (e_.dtype != torch.float32 or not e_.is_contiguous() or e_.device != dev):
                raise RuntimeError(f"moegan: '{n}' must be a contiguous float32 tensor on {dev}")
        router_params = [mu_f, rho_f, mu_t, rho_t, mu_c, rho_c, temperature]
        w1 = expert_params[0::4]; b1 = expert_params[1::4]; w2 = expert_params[2::4]; b2 = expert_params[3::4]
        plan = _plan_for(cfg)
        need_bwd = cfg.need_backward
        (W1_op, b1s, W2_op, b2s), stale = wcache.get(cfg, w1, b1, w2, b2)

        ws = torch.empty((plan.fwd_bytes if need_bwd else plan.fwd_bytes_nobwd,), dtype=torch.uint8, device=dev)
        if residual is not None:
            if residual.dtype not in (torch.float32, torch.bfloat16):
                residual = residual.float()
            residual = residual.contiguous()
            if residual.shape != (T, d):
                raise ValueError(f"residual must be [T, d] = [{T}, {d}], got {tuple(residual.shape)}")
        out = torch.empty((T, d), dtype=f32 if residual is None else residual.dtype, device=dev)
        probs = torch.empty((T, E), dtype=f32, device=dev)
        small = torch.empty((E + C.KL_OUT_FLOATS,), dtype=f32, device=dev)     # [load | kl_out]

        a = C.BlockFwdArgs()
        a.x, a.w = x_tok.data_ptr(), w.data_ptr()
        (a.mu_f, a.rho_f, a.mu_t, a.rho_t, a.mu_c, a.rho_c, a.temperature) = [p.data_ptr() for p in router_params]
        a.eps_f, a.eps_t, a.eps_c = _dp(eps_f), _dp(eps_t), _dp(eps_c)
        a.W1_op, a.b1s, a.W2_op, a.b2s = W1_op.data_ptr(), b1s.data_ptr(), W2_op.data_ptr(), b2s.data_ptr()
        if stale:
            tabs = [C.ptr_array(g) for g in (w1, b1, w2, b2)]       # kept alive until the call returns
            a.prep_w1, a.prep_b1, a.prep_w2, a.prep_b2 = [ctypes.cast(t, C.c_void_p) for t in tabs]
        a.annealing_factor = float(cfg.annealing_factor)
        a.sampled = 1 if cfg.sampled else 0
        a.need_backward = 1 if need_bwd else 0
        a.out, a.probs = out.data_ptr(), probs.data_ptr()
        a.load = small.data_ptr()
        a.kl_out = small.data_ptr() + 4 * E
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        a.residual = _dp(residual)
        if x_op is not None:
            if cfg.precision != 0 or x_op.dtype != torch.bfloat16 or x_op.shape != (T, d) or not x_op.is_contiguous() or x_op.device != dev:
                raise ValueError(f"x_operand must be a contiguous bf16 [T, d] = [{T}, {d}] copy of x on {dev} and the block must run in bf16 mode")
            a.x_op = x_op.data_ptr()
        a.residual_is_bf16 = 1 if (residual is not None and residual.dtype == torch.bfloat16) else 0
        a.out_is_bf16 = 1 if out.dtype == torch.bfloat16 else 0
        with torch.cuda.device(dev):     # launches go to the current stream OF x's DEVICE
            C.check(lib.moegan_block_fwd(plan.shape, ctypes.byref(a), C.c_void_p(_stream_ptr())), "moegan_block_fwd")
        if stale and T > 0:
            wcache.commit()              # only now are the converted buffers valid (a failed call raised above)

        load = small[:E]
        kl = small[E] if cfg.sampled else torch.zeros((), dtype=f32, device=dev)   # eval: tensor(0.) (:489)
        if need_bwd:
            ctx.cfg, ctx.plan = cfg, plan
            ctx.ws, ctx.small = ws, small
            ctx.w = w
            ctx.router_params = router_params
            ctx.wops = (W1_op, W2_op)
            ctx.has_residual = residual is not None
            ctx.x_op = x_op
        if MoEBlockFunction.keep_last:
            MoEBlockFunction._last = (ws, plan.layout if need_bwd else plan.layout_nobwd, cfg)
        return out, kl, probs, load

    @staticmethod
    def backward(ctx, dout, dkl, dprobs, dload):
        lib = C.lib()
        cfg, plan = ctx.cfg, ctx.plan
        if getattr(ctx, "ws", None) is None:
            raise RuntimeError("moegan: the MoE block's forward workspace was released by its first backward pass; a second backward "
                               "through the same forward (retain_graph=True) is not supported - run the forward again")
        E, d, Dt, T, B = cfg.E, cfg.d, cfg.Dt, cfg.T, cfg.B
        dev = ctx.w.device
        f32 = torch.float32
        hid = 4 * d
        if dout is None:
            dout = torch.zeros((T, d), dtype=f32, device=dev)
        else:
            dout = (dout if dout.dtype in (f32, torch.bfloat16) else dout.float()).contiguous()   # bf16 is read as such by combine_bwd
        dprobs = None if dprobs is None else dprobs.contiguous().float()
        dload = None if dload is None else dload.contiguous().float()
        dkl = None if dkl is None else dkl.reshape(1).contiguous().float()

        ws_b = torch.empty((plan.bwd_bytes,), dtype=torch.uint8, device=dev)
        dx = torch.empty((T, d), dtype=f32, device=dev)
        dW1 = torch.empty((E, hid, d), dtype=f32, device=dev)
        db1 = torch.empty((E, hid), dtype=f32, device=dev)
        dW2 = torch.empty((E, d, hid), dtype=f32, device=dev)
        db2 = torch.empty((E, d), dtype=f32, device=dev)
        a = C.BlockBwdArgs()
        a.fwd_workspace = ctx.ws.data_ptr()
        a.w = ctx.w.data_ptr()
        (a.mu_f, a.rho_f, a.mu_t, a.rho_t, a.mu_c, a.rho_c, a.temperature) = [p.data_ptr() for p in ctx.router_params]
        a.W1_op, a.W2_op = ctx.wops[0].data_ptr(), ctx.wops[1].data_ptr()
        a.kl_out = ctx.small.data_ptr() + 4 * E
        a.dout, a.dkl, a.dprobs, a.dload = dout.data_ptr(), _dp(dkl), _dp(dprobs), _dp(dload)
        a.annealing_factor = float(cfg.annealing_factor)
        a.sampled = 1 if cfg.sampled else 0
        a.dout_is_bf16 = 1 if dout.dtype == torch.bfloat16 else 0
        a.x_op = _dp(ctx.x_op)
        a.dx = dx.data_ptr()
        a.dW1, a.db1, a.dW2, a.db2 = dW1.data_ptr(), db1.data_ptr(), dW2.data_ptr(), db2.data_ptr()
        router_grads = [None] * 7
        dw = None
        if not cfg.hard:
            sizes = [p.numel() for p in ctx.router_params]
            flat = torch.empty((sum(sizes),), dtype=f32, device=dev)
            router_grads = [g.view(p.shape) for g, p in zip(flat.split(sizes), ctx.router_params)]
            dw = torch.empty((B, Dt), dtype=f32, device=dev)
            a.dw = dw.data_ptr()
            (a.dmu_f, a.drho_f, a.dmu_t, a.drho_t, a.dmu_c, a.drho_c, a.dtemperature) = [g.data_ptr() for g in router_grads]
        a.workspace, a.workspace_bytes = ws_b.data_ptr(), ws_b.numel()
        with torch.cuda.device(dev):
            C.check(lib.moegan_block_bwd(plan.shape, ctypes.byref(a), C.c_void_p(_stream_ptr())), "moegan_block_bwd")
        expert_grads = [t for quad in zip(dW1.unbind(0), db1.unbind(0), dW2.unbind(0), db2.unbind(0)) for t in quad]
        ctx.ws = None
        d_res = dout if ctx.has_residual else None     # d(residual + moe) / d residual = identity
        return (None, None, dx, dw, None, None, None, *router_grads, d_res, None, *expert_grads)


class BalanceLossFunction(torch.autograd.Function):
    """bw * nan_to_num(clamp(E * cv(frac), 0, 10)) on the fused per-expert load (t2i_moe_gan.py:973-1000)."""

    @staticmethod
    def forward(ctx, load, T: int, balance_weight: float):
        _require_cuda(load, "load")
        load = load.contiguous().float()
        loss = torch.empty((1,), dtype=torch.float32, device=load.device)
        C.check(C.lib().moegan_balance_loss_fwd(C.ptr(load), load.numel(), int(T), float(balance_weight), C.ptr(loss),
                                                C.c_void_p(_stream_ptr())), "moegan_balance_loss_fwd")
        ctx.save_for_backward(load)
        ctx.T, ctx.bw = int(T), float(balance_weight)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        (load,) = ctx.saved_tensors
        dloss = dloss.reshape(1).contiguous().float()
        dload = torch.empty_like(load)
        C.check(C.lib().moegan_balance_loss_bwd(C.ptr(load), load.numel(), ctx.T, ctx.bw, C.ptr(dloss), C.ptr(dload),
                                                C.c_void_p(_stream_ptr())), "moegan_balance_loss_bwd")
        return dload, None, None
