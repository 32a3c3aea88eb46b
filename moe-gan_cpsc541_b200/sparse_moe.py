"""Drop-in mirror of the reference's MoE module interface (moegan/t2i_moe_gan.py:249-491).

`SparseMoE(dim, text_dim, num_experts=4)` has the reference's constructor, attributes, parameter names,
buffers, state_dict keys, forward signature and return contract; its forward runs the
hand-written sm_100a kernels through the C ABI instead of eager torch ops.

Swap-in:
    import t2i_moe_gan, importlib
    b200 = importlib.import_module("moe-gan_cpsc541_b200")
    b200.install(t2i_moe_gan)                      # before AuroraGenerator() is constructed
or, post hoc:  blk.attn_block.moe = b200.SparseMoE.from_reference(blk.attn_block.moe)
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .functional import MoEBlockFunction, BalanceLossFunction, MoEConfig, WeightCache

NUM_EXPERTS = 4  # t2i_moe_gan.py:23


class SparseExpertFFN(nn.Module):
    """Parameter container with the reference layout (t2i_moe_gan.py:249-263): net.0 = Linear(d,4d), net.2 = Linear(4d,d).

    The fused block reads these parameters directly; `forward` exists only for API parity when somebody calls
    an expert on its own (plain torch, not on the hot path).
    """

    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, dim * 4), nn.GELU(), nn.Linear(dim * 4, dim))

    def forward(self, x):
        return self.net(x)


class BayesianRouter(nn.Module):
    """Parameters/buffers of the reference router (t2i_moe_gan.py:265-301), same names, shapes and init."""

    def __init__(self, feature_dim: int, text_dim: int, num_experts: int = NUM_EXPERTS):
        super().__init__()
        self.feature_dim, self.text_dim, self.num_experts = feature_dim, text_dim, num_experts
        self.feature_mu = nn.Parameter(torch.empty(feature_dim, 128).normal_(0.0, 0.01))
        self.feature_rho = nn.Parameter(torch.full((feature_dim, 128), -4.0))
        self.text_mu = nn.Parameter(torch.empty(text_dim, 128).normal_(0.0, 0.01))
        self.text_rho = nn.Parameter(torch.full((text_dim, 128), -4.0))
        self.combined_mu = nn.Parameter(torch.empty(256, num_experts).normal_(0.0, 0.01))
        self.combined_rho = nn.Parameter(torch.full((256, num_experts), -4.0))
        self.register_buffer("epsilon_f", torch.zeros(feature_dim, 128))
        self.register_buffer("epsilon_t", torch.zeros(text_dim, 128))
        self.register_buffer("epsilon_c", torch.zeros(256, num_experts))
        self.temperature = nn.Parameter(torch.ones(1) * 4.0)

    def kl_divergence(self) -> torch.Tensor:
        """Same value as the reference method (:405-423).  Off the hot path (SparseMoE.forward gets KL from the
        fused sample+KL kernel); kept so code that calls `moe.router.kl_divergence()` keeps working."""
        def term(mu, rho):
            log_var = 2 * torch.log(torch.log1p(torch.exp(rho)))
            return 0.5 * torch.sum(torch.exp(log_var) + mu.pow(2) - 1 - log_var)
        kl = term(self.feature_mu, self.feature_rho) + term(self.text_mu, self.text_rho) + term(self.combined_mu, self.combined_rho)
        kl = torch.nan_to_num(kl, nan=0.0, posinf=200.0, neginf=0.0)
        return torch.clamp(kl, min=0.0, max=120.0)

    @torch.no_grad()
    def forward(self, feature, text_embedding, sampling=True, annealing_factor=1.0):
        """Same contract as the reference (:335-402): feature [T,dim], text_embedding [T,text_dim] (the per-token expansion
        the reference builds at :456) -> (probs [T,E], logits [T,E]).  Runs the router kernels (sample+KL, text
        projection, fused router); inference-only here - gradients flow through SparseMoE.forward, which is what
        the model calls."""
        from . import _cabi as C
        if not feature.is_cuda:
            raise RuntimeError("BayesianRouter (B200 path) needs CUDA tensors: there is no CPU fallback")
        lib = C.lib()
        T, d = feature.shape
        Dt, E = text_embedding.shape[1], self.num_experts
        dev, f32 = feature.device, torch.float32
        training = bool(sampling and self.training)
        if training:
            self.epsilon_f.normal_(); self.epsilon_t.normal_(); self.epsilon_c.normal_()
        hard = not self.training
        k = 1 if hard else E
        shp = C.make_shape(T, d, Dt, E, k, T, 1, 1, not hard, hard)       # one "sample" per token: text rows are per token here
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        Wf = torch.empty((d, 128), dtype=f32, device=dev); Wt = torch.empty((Dt, 128), dtype=f32, device=dev)
        Wc = torch.empty((256, E), dtype=f32, device=dev); kl = torch.empty((C.KL_OUT_FLOATS,), dtype=f32, device=dev)
        C.check(lib.moegan_router_sample_kl_fwd(shp, C.ptr(self.feature_mu), C.ptr(self.feature_rho), C.ptr(self.epsilon_f),
                                                C.ptr(self.text_mu), C.ptr(self.text_rho), C.ptr(self.epsilon_t),
                                                C.ptr(self.combined_mu), C.ptr(self.combined_rho), C.ptr(self.epsilon_c),
                                                1 if training else 0, C.ptr(Wf), C.ptr(Wt), C.ptr(Wc), None, None, C.ptr(kl), st),
                "moegan_router_sample_kl_fwd")
        tp = torch.empty((T, 128), dtype=f32, device=dev); tl = torch.empty((T, E), dtype=f32, device=dev)
        x = feature.contiguous().float(); w = text_embedding.contiguous().float()
        C.check(lib.moegan_router_text_fwd(shp, C.ptr(w), C.ptr(Wt), C.ptr(Wc), C.ptr(tp), C.ptr(tl), st), "moegan_router_text_fwd")
        nch = max((T + C.CHUNK - 1) // C.CHUNK, 1)
        raw = torch.empty((T, E), dtype=f32, device=dev); probs = torch.empty((T, E), dtype=f32, device=dev)
        idx = torch.empty((T, k), dtype=torch.int32, device=dev); gates = torch.empty((T, k), dtype=f32, device=dev)
        cc = torch.empty((nch, E), dtype=torch.int32, device=dev); cs = torch.empty((nch, E), dtype=f32, device=dev)
        C.check(lib.moegan_router_fwd(shp, C.ptr(x), C.ptr(Wf), C.ptr(Wc), C.ptr(tl), C.ptr(self.temperature), float(annealing_factor),
                                      None, C.ptr(raw), C.ptr(probs), C.ptr(idx), C.ptr(gates), C.ptr(cc), C.ptr(cs), st), "moegan_router_fwd")
        tau = torch.clamp(self.temperature * annealing_factor, min=0.5, max=5.0)
        return probs, torch.clamp(raw / tau, min=-20.0, max=20.0)


class SparseMoE(nn.Module):
    """B200-native SparseMoE.  forward(x[B,d,H,W], w[B,text_dim], annealing_factor=1.0) -> (out, kl, probs[T,E]).

    Extra keyword-only options (not in the reference):
      top_k      None -> reference behaviour (all E experts with soft probs in training, hard top-1 in eval);
                 1 <= top_k < E -> renormalised top-k gating in training.
      precision  "auto" (bf16 under autocast, else fp32), "bf16" or "fp32".
      eps        forward(..., eps=(ef, et, ec)) consumes caller-supplied noise instead of redrawing the buffers.
    """

    def __init__(self, dim: int, text_dim: int, num_experts: int = NUM_EXPERTS, *, top_k: Optional[int] = None,
                 precision: str = "auto"):
        super().__init__()
        if precision not in ("auto", "bf16", "fp32"):
            raise ValueError("precision must be 'auto', 'bf16' or 'fp32'")
        if top_k is not None and not (1 <= top_k <= num_experts):
            raise ValueError("top_k must be in [1, num_experts]")
        self.dim = dim
        self.num_experts = num_experts
        self.top_k = top_k
        self.precision = precision
        self.experts = nn.ModuleList([SparseExpertFFN(dim) for _ in range(num_experts)])
        self.router = BayesianRouter(dim, text_dim, num_experts)
        self._wcache = WeightCache()
        self.fuses_residual = True   # forward(..., residual=) adds the skip connection inside the combine kernel
        self.takes_operand = True    # forward(..., x_operand=) consumes the bf16 token copy norm3 emitted (no conversion pass)
        self.redraw_eps = True    # False = parity mode "use the epsilon_* buffers as they are" 
        self.last_load: Optional[torch.Tensor] = None   # fused per-expert prob column sums of the last forward

    @classmethod
    def from_reference(cls, ref_moe: nn.Module, **kw) -> "SparseMoE":
        """Build from a reference SparseMoE instance (same weights, same device)."""
        text_dim = ref_moe.router.text_mu.shape[0]
        m = cls(ref_moe.dim, text_dim, ref_moe.num_experts, **kw)
        m.load_state_dict(ref_moe.state_dict())
        dev = next(ref_moe.parameters()).device
        m.to(dev)
        m.train(ref_moe.training)
        return m

    def _precision_code(self) -> int:
        if self.precision == "bf16":
            return 0
        if self.precision == "fp32":
            return 1
        return 0 if torch.is_autocast_enabled() else 1

    def forward(self, x: torch.Tensor, w: torch.Tensor, annealing_factor: float = 1.0, *,
                eps: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None, residual: Optional[torch.Tensor] = None,
                x_operand: Optional[torch.Tensor] = None):
        """Reference contract (t2i_moe_gan.py:444-491).  `residual` (optional, [B,d,H,W], same memory layout as x): return
        `residual + moe(x)` in the residual's dtype - the skip connection of AttentionBlock (:571) fused into the combine
        kernel.  `x_operand` (optional, bf16 mode): the bf16 copy of x's tokens that `glue.layer_norm(...,
        with_operand=True)` emitted in the same pass as x (norm3, :561) - the block then skips its own conversion pass."""
        if not x.is_cuda:
            raise RuntimeError("SparseMoE (B200 path) needs CUDA tensors: there is no CPU fallback")
        batch_size, channels, height, width = x.shape
        r = self.router
        training = self.training
        E = self.num_experts
        hard = not training
        k = (1 if hard else E) if self.top_k is None else (1 if hard else self.top_k)
        dense = (k == E) and not hard
        precision = self._precision_code()
        if training:   # sampling=self.training (:459)
            if eps is None:
                if self.redraw_eps:
                    r.epsilon_f.normal_(); r.epsilon_t.normal_(); r.epsilon_c.normal_()   # (:349-351), same RNG order
                eps = (r.epsilon_f, r.epsilon_t, r.epsilon_c)
            else:
                eps = tuple(e.to(device=x.device, dtype=torch.float32).contiguous() for e in eps)
        else:
            eps = (None, None, None)
        # tokens are (b, h, w)-major: x is normally an NCHW view of NHWC memory, so this is a view (:455)
        x_tok = x.permute(0, 2, 3, 1).contiguous().view(-1, channels)
        params = [r.feature_mu, r.feature_rho, r.text_mu, r.text_rho, r.combined_mu, r.combined_rho, r.temperature]
        expert_params = []
        for ex in self.experts:
            expert_params += [ex.net[0].weight, ex.net[0].bias, ex.net[2].weight, ex.net[2].bias]
        need_bwd = torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or any(p.requires_grad for p in params)
                                                or any(p.requires_grad for p in expert_params))
        af = float(annealing_factor)
        cfg = MoEConfig(B=batch_size, tokens_per_sample=height * width, d=channels, Dt=w.shape[1], E=E, k=k,
                        dense=dense, hard=hard, precision=precision, sampled=training, annealing_factor=af,
                        need_backward=need_bwd)
        res_tok = None
        if residual is not None:
            if residual.shape != x.shape:
                raise ValueError(f"residual {tuple(residual.shape)} must have the shape of x {tuple(x.shape)}")
            res_tok = residual.permute(0, 2, 3, 1).contiguous().view(-1, channels)      # a view for NHWC memory
        with torch.autocast(device_type="cuda", enabled=False):
            x_op = None
            if x_operand is not None and precision == 0:
                x_op = x_operand.reshape(-1, channels)
            out, kl, probs, load = MoEBlockFunction.apply(cfg, self._wcache, x_tok, w, eps[0], eps[1], eps[2], *params, res_tok, x_op, *expert_params)
        self.last_load = load
        if not hard:
            # lets moe_balance_loss() use the fused column sums.  Not in eval mode: there `probs` is one-hot (:392-400) while the
            # fused load is the soft-probability column sum, and the reference's balance loss is defined on what it is handed
            probs._moegan_load = load
        if residual is None and out.dtype != x.dtype:
            out = out.to(x.dtype)
        output = out.reshape(batch_size, height, width, channels).permute(0, 3, 1, 2)   # (:486)
        return output, kl, probs


def moe_balance_loss(routing_probs, balance_weight: float = 0.01) -> torch.Tensor:
    """Drop-in for AuroraGANLoss.moe_balance_loss (t2i_moe_gan.py:951-1000).

    Uses the per-expert load the router kernel already reduced (no second pass over [T,E]); for probs that
    did not come from this package it first column-sums them.
    """
    if not routing_probs:
        return torch.tensor(0.0)
    last = routing_probs[-1]
    if last is None or last.numel() == 0:
        return torch.tensor(0.0, device=last.device if last is not None else None)
    load = getattr(last, "_moegan_load", None)
    # The fused load is only a valid stand-in for last.sum(0) when it is on the same autograd graph as `last`.  Under reentrant
    # activation checkpointing (torch.utils.checkpoint's default when use_reentrant is not given) the block runs under no_grad
    # and the checkpoint re-attaches a grad_fn to `probs` only - the attribute survives but points at a load without history,
    # and the balance loss would silently lose its gradient to the router.  Then: column-sum the probs that DO carry the graph.
    if load is None or load.requires_grad != last.requires_grad or load.shape[0] != last.shape[1]:
        load = last.float().sum(dim=0)
    return BalanceLossFunction.apply(load, last.shape[0], float(balance_weight))


def install(reference_module) -> None:
    """Point the reference module's `SparseMoE` (and its balance loss) at the B200 implementation."""
    reference_module.SparseMoE = SparseMoE
    if hasattr(reference_module, "AuroraGANLoss"):
        reference_module.AuroraGANLoss.moe_balance_loss = lambda self, routing_probs, balance_weight=0.01: \
            moe_balance_loss(routing_probs, balance_weight)
