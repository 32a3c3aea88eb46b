"""CPU reference arm for the FULL train step: the numpy oracle of the MoE block wrapped as a torch module so that the
host model (generator / discriminator in plain PyTorch) can run a complete D-step + G-step on the host cores.

TEST / BENCH INFRASTRUCTURE ONLY (see moe_oracle.py header): used by bench.py's `cpu_baseline` leg and
`--impl reference`, and by tests.  The product package never imports this file.

Semantics = the reference's SparseMoE (moegan/t2i_moe_gan.py:426-491): same constructor, parameter names, buffers,
forward contract; eps is redrawn into the buffers on every training forward (:349-351).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import moe_oracle as O


class _OracleMoEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x_tok, w, af, training, tps, *params):
        names = mod._param_names
        P = {n: p.detach().numpy() for n, p in zip(names, params)}
        eps = {k: getattr(mod.router, k).numpy() for k in O.EPS_KEYS}
        out, kl, probs, cache = O.moe_forward(x_tok.detach().numpy(), w.detach().numpy(), P, eps, af=af, training=training,
                                              tokens_per_sample=tps, dtype=np.float32)
        ctx.cache, ctx.P, ctx.eps, ctx.names = cache, P, eps, names
        return (torch.from_numpy(np.ascontiguousarray(out, dtype=np.float32)), torch.tensor(kl, dtype=torch.float32),
                torch.from_numpy(np.ascontiguousarray(probs, dtype=np.float32)))

    @staticmethod
    def backward(ctx, dout, dkl, dprobs):
        g = O.moe_backward(ctx.cache, ctx.P, ctx.eps, dout.numpy(), dprobs=None if dprobs is None else dprobs.numpy(),
                           dkl=0.0 if dkl is None else float(dkl), dtype=np.float32)
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        return (None, f(g["dx"]), f(g["dw"]), None, None, None, *[f(g[n]).reshape(ctx.P[n].shape) for n in ctx.names])


class OracleSparseMoE(nn.Module):
    def __init__(self, dim, text_dim, num_experts=4):
        super().__init__()
        self.dim, self.num_experts = dim, num_experts

        class _FFN(nn.Module):
            def __init__(s):
                super().__init__()
                s.net = nn.Sequential(nn.Linear(dim, dim * 4), nn.GELU(), nn.Linear(dim * 4, dim))
        self.experts = nn.ModuleList([_FFN() for _ in range(num_experts)])

        class _Router(nn.Module):
            def __init__(s):
                super().__init__()
                s.feature_mu = nn.Parameter(torch.empty(dim, 128).normal_(0, 0.01)); s.feature_rho = nn.Parameter(torch.full((dim, 128), -4.0))
                s.text_mu = nn.Parameter(torch.empty(text_dim, 128).normal_(0, 0.01)); s.text_rho = nn.Parameter(torch.full((text_dim, 128), -4.0))
                s.combined_mu = nn.Parameter(torch.empty(256, num_experts).normal_(0, 0.01)); s.combined_rho = nn.Parameter(torch.full((256, num_experts), -4.0))
                s.register_buffer("epsilon_f", torch.zeros(dim, 128)); s.register_buffer("epsilon_t", torch.zeros(text_dim, 128))
                s.register_buffer("epsilon_c", torch.zeros(256, num_experts))
                s.temperature = nn.Parameter(torch.ones(1) * 4.0)
        self.router = _Router()
        self.redraw_eps = True
        self._param_names = [n for n, _ in self.named_parameters()]

    def forward(self, x, w, annealing_factor=1.0):
        B, C, H, W = x.shape
        if self.training and self.redraw_eps:
            self.router.epsilon_f.normal_(); self.router.epsilon_t.normal_(); self.router.epsilon_c.normal_()
        x_tok = x.permute(0, 2, 3, 1).contiguous().view(-1, C).float()
        out, kl, probs = _OracleMoEFn.apply(self, x_tok, w.float(), float(annealing_factor), self.training, H * W,
                                            *[p for _, p in self.named_parameters()])
        return out.reshape(B, H, W, C).permute(0, 3, 1, 2).to(x.dtype), kl, probs


class TorchSparseMoE(OracleSparseMoE):
    """Same module, but the block is evaluated with plain torch CPU ops + torch autograd instead of numpy: this is
    the closest stand-in for "the reference's own CPU implementation" (dense-soft training path :465-470, hard
    top-1 eval path :471-483) and is what bench.py times as the reference arm.  Checked against the numpy oracle in
    tests/test_oracle_golden.py."""

    def forward(self, x, w, annealing_factor=1.0):
        B, C, H, W = x.shape
        r = self.router
        xt = x.permute(0, 2, 3, 1).reshape(-1, C)
        xr = torch.nan_to_num(xt, nan=0.0, posinf=1.0, neginf=-1.0)
        wr = torch.nan_to_num(w, nan=0.0, posinf=1.0, neginf=-1.0)
        if self.training:
            if self.redraw_eps:
                r.epsilon_f.normal_(); r.epsilon_t.normal_(); r.epsilon_c.normal_()

            def sample(mu, rho, eps):
                sigma = torch.log1p(torch.exp(rho.clamp(-8.0, 4.0))).clamp(1e-6, 10.0)
                return mu.clamp(-10.0, 10.0) + sigma * eps.clamp(-2.0, 2.0)
            Wf, Wt, Wc = sample(r.feature_mu, r.feature_rho, r.epsilon_f), sample(r.text_mu, r.text_rho, r.epsilon_t), \
                sample(r.combined_mu, r.combined_rho, r.epsilon_c)
        else:
            Wf, Wt, Wc = r.feature_mu, r.text_mu, r.combined_mu
        tp = (wr @ Wt).repeat_interleave(H * W, dim=0)
        logits = torch.cat([xr @ Wf, tp], dim=1) @ Wc
        logits = (logits / torch.clamp(r.temperature * annealing_factor, 0.5, 5.0)).clamp(-20.0, 20.0)
        probs = torch.softmax(logits, dim=1).clamp(1e-6, 1.0)
        probs = probs / probs.sum(dim=1, keepdim=True)
        out = torch.zeros_like(xt)
        if self.training:
            for i, ex in enumerate(self.experts):
                out = out + probs[:, i:i + 1] * ex.net(xt)

            def kl_term(mu, rho):
                lv = 2 * torch.log(torch.log1p(torch.exp(rho)))
                return 0.5 * torch.sum(torch.exp(lv) + mu.pow(2) - 1 - lv)
            kl = kl_term(r.feature_mu, r.feature_rho) + kl_term(r.text_mu, r.text_rho) + kl_term(r.combined_mu, r.combined_rho)
            kl = torch.nan_to_num(kl, nan=0.0, posinf=200.0, neginf=0.0).clamp(0.0, 120.0)
        else:
            top = probs.argmax(dim=1)
            probs = torch.zeros_like(probs).scatter_(1, top[:, None], 1.0)
            for i, ex in enumerate(self.experts):
                m = top == i
                if m.any():
                    out[m] = ex.net(xt[m])
            kl = torch.zeros((), device=x.device)
        return out.reshape(B, H, W, C).permute(0, 3, 1, 2), kl, probs
