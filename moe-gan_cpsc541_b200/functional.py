"""autograd wrappers over the C ABI: the whole MoE block (router -> dispatch -> experts -> combine, plus
KL and balance-load reductions) as ONE torch.autograd.Function, and the balance loss on the fused load.

Reference path replaced: SparseMoE.forward + BayesianRouter.forward / kl_divergence + SparseExpertFFN.forward
(moegan/t2i_moe_gan.py:249-491) and AuroraGANLoss.moe_balance_loss (:951-1000).
PyTorch is used here only for device memory, the current stream and autograd bookkeeping.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

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


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"moegan: '{name}' must be a CUDA tensor - the B200 path has no CPU fallback")


class WeightCache:
    """bf16 operand copies of the expert weights, rebuilt only when a parameter changed (optimizer step).

    The reference re-casts every nn.Linear weight on every autocast forward; here both generator forwards
    of a train step (D-step under no_grad, G-step with grad) share one conversion.
    """

    def __init__(self):
        self.key = None
        self.bufs = None

    def get(self, cfg: MoEConfig, shape: C.Shape, w1: Sequence[torch.Tensor], b1, w2, b2):
        key = (cfg.precision, cfg.d, cfg.E,
               tuple((p.data_ptr(), p._version) for group in (w1, b1, w2, b2) for p in group))
        if key == self.key and self.bufs is not None:
            return self.bufs
        dev = w1[0].device
        S = 2 if cfg.precision == 1 else 1
        hid = 4 * cfg.d
        W1_op = torch.empty((cfg.E * hid, S * cfg.d), dtype=torch.bfloat16, device=dev)
        W2_op = torch.empty((cfg.E * cfg.d, S * hid), dtype=torch.bfloat16, device=dev)
        b1s = torch.empty((cfg.E, hid), dtype=torch.float32, device=dev)
        b2s = torch.empty((cfg.E, cfg.d), dtype=torch.float32, device=dev)
        rc = C.lib().moegan_prep_weights(shape, C.ptr_array(w1), C.ptr_array(b1), C.ptr_array(w2), C.ptr_array(b2),
                                         C.ptr(W1_op), C.ptr(b1s), C.ptr(W2_op), C.ptr(b2s), _stream())
        C.check(rc, "moegan_prep_weights")
        self.key, self.bufs = key, (W1_op, b1s, W2_op, b2s)
        return self.bufs


class _Saved:
    """Intermediate device buffers kept for backward (plain attributes, never autograd-tracked)."""
    pass


class MoEBlockFunction(torch.autograd.Function):
    """(x_tok [T,d], w [B,Dt], eps x3, router params x7, expert params x4E) -> (out [T,d], kl [], probs [T,E], load [E])."""

    @staticmethod
    def forward(ctx, cfg: MoEConfig, wcache: WeightCache, x_tok, w, eps_f, eps_t, eps_c,
                mu_f, rho_f, mu_t, rho_t, mu_c, rho_c, temperature, *expert_params):
        lib = C.lib()
        E, d, Dt, T, B, k = cfg.E, cfg.d, cfg.Dt, cfg.T, cfg.B, cfg.k
        for name, t in (("x", x_tok), ("w", w), ("feature_mu", mu_f), ("temperature", temperature)):
            _require_cuda(t, name)
        dev = x_tok.device
        f32, bf16, i32 = torch.float32, torch.bfloat16, torch.int32
        x_tok = x_tok.contiguous().float()
        w = w.contiguous().float()
        router_params = [p.contiguous() for p in (mu_f, rho_f, mu_t, rho_t, mu_c, rho_c, temperature)]
        mu_f, rho_f, mu_t, rho_t, mu_c, rho_c, temperature = router_params
        w1 = [expert_params[4 * e + 0] for e in range(E)]
        b1 = [expert_params[4 * e + 1] for e in range(E)]
        w2 = [expert_params[4 * e + 2] for e in range(E)]
        b2 = [expert_params[4 * e + 3] for e in range(E)]
        shape = cfg.shape()
        st = _stream()
        S = 2 if cfg.precision == 1 else 1
        hid = 4 * d
        m_cap = shape.m_cap
        nchunks = (T + C.CHUNK - 1) // C.CHUNK
        need_bwd = cfg.need_backward
        act_dt = f32 if cfg.precision == 1 else bf16

        sv = _Saved()
        # ---- K1: sampled router weights + KL --------------------------------------------------------
        Wf = torch.empty((d, HID), dtype=f32, device=dev)
        Wt = torch.empty((Dt, HID), dtype=f32, device=dev)
        Wc = torch.empty((2 * HID, E), dtype=f32, device=dev)
        Wf_op = torch.empty((d, S * HID), dtype=bf16, device=dev) if need_bwd else None
        Wt_op = torch.empty((Dt, S * HID), dtype=bf16, device=dev) if need_bwd else None
        kl_out = torch.empty((C.KL_OUT_FLOATS,), dtype=f32, device=dev)
        C.check(lib.moegan_router_sample_kl_fwd(shape, C.ptr(mu_f), C.ptr(rho_f), C.ptr(eps_f), C.ptr(mu_t), C.ptr(rho_t),
                                                C.ptr(eps_t), C.ptr(mu_c), C.ptr(rho_c), C.ptr(eps_c), 1 if cfg.sampled else 0,
                                                C.ptr(Wf), C.ptr(Wt), C.ptr(Wc), C.ptr(Wf_op), C.ptr(Wt_op), C.ptr(kl_out), st),
                "moegan_router_sample_kl_fwd")
        # ---- K2: router -----------------------------------------------------------------------------
        tp = torch.empty((B, HID), dtype=f32, device=dev)
        tl = torch.empty((B, E), dtype=f32, device=dev)
        C.check(lib.moegan_router_text_fwd(shape, C.ptr(w), C.ptr(Wt), C.ptr(Wc), C.ptr(tp), C.ptr(tl), st), "moegan_router_text_fwd")
        fp = torch.empty((T, HID), dtype=f32, device=dev) if need_bwd else None
        raw = torch.empty((T, E), dtype=f32, device=dev)
        probs = torch.empty((T, E), dtype=f32, device=dev)
        topk_idx = torch.empty((T, k), dtype=i32, device=dev)
        gates = torch.empty((T, k), dtype=f32, device=dev)
        chunk_counts = torch.empty((max(nchunks, 1), E), dtype=i32, device=dev)
        chunk_colsum = torch.empty((max(nchunks, 1), E), dtype=f32, device=dev)
        C.check(lib.moegan_router_fwd(shape, C.ptr(x_tok), C.ptr(Wf), C.ptr(Wc), C.ptr(tl), C.ptr(temperature),
                                      float(cfg.annealing_factor), C.ptr(fp), C.ptr(raw), C.ptr(probs), C.ptr(topk_idx), C.ptr(gates),
                                      C.ptr(chunk_counts), C.ptr(chunk_colsum), st), "moegan_router_fwd")
        # ---- K3: dispatch ---------------------------------------------------------------------------
        plan = torch.empty((int(lib.moegan_plan_ints(m_cap)),), dtype=i32, device=dev)
        chunk_base = torch.empty((max(nchunks, 1), E), dtype=i32, device=dev)
        load = torch.empty((E,), dtype=f32, device=dev)
        C.check(lib.moegan_dispatch_plan(shape, C.ptr(chunk_counts), C.ptr(chunk_colsum), C.ptr(plan), C.ptr(chunk_base), C.ptr(load), st),
                "moegan_dispatch_plan")
        x_copy = torch.empty((T, S * d), dtype=bf16, device=dev) if (need_bwd or cfg.dense) else None
        row_token = row_gate = token_rows = None
        if cfg.dense:
            C.check(lib.moegan_convert_tokens(shape, C.ptr(x_tok), C.ptr(x_copy), st), "moegan_convert_tokens")
            a_op = x_copy
        else:
            a_op = torch.empty((m_cap, S * d), dtype=bf16, device=dev)
            row_token = torch.empty((m_cap,), dtype=i32, device=dev)
            row_gate = torch.empty((m_cap,), dtype=f32, device=dev)
            token_rows = torch.empty((T, k), dtype=i32, device=dev)
            C.check(lib.moegan_dispatch_gather(shape, C.ptr(x_tok), C.ptr(topk_idx), C.ptr(gates), C.ptr(plan), C.ptr(chunk_base),
                                               C.ptr(a_op), C.ptr(x_copy), C.ptr(row_token), C.ptr(row_gate), C.ptr(token_rows), st),
                    "moegan_dispatch_gather")
        # ---- K4: experts ----------------------------------------------------------------------------
        W1_op, b1s, W2_op, b2s = wcache.get(cfg, shape, w1, b1, w2, b2)
        pre = torch.empty((m_cap, hid), dtype=act_dt, device=dev) if need_bwd else None   # gelu'(pre) (bf16) / pre (fp32)
        h_op = torch.empty((m_cap, S * hid), dtype=bf16, device=dev)
        y = torch.empty((m_cap, d), dtype=act_dt, device=dev)
        C.check(lib.moegan_ffn_fwd(shape, C.ptr(plan), C.ptr(a_op), C.ptr(W1_op), C.ptr(b1s), C.ptr(W2_op), C.ptr(b2s),
                                   C.ptr(pre), C.ptr(h_op), C.ptr(y), st), "moegan_ffn_fwd")
        # ---- K5: combine ----------------------------------------------------------------------------
        out = torch.empty((T, d), dtype=f32, device=dev)
        C.check(lib.moegan_combine_fwd(shape, C.ptr(y), C.ptr(gates), C.ptr(token_rows), C.ptr(out), st), "moegan_combine_fwd")

        kl = kl_out[0].clone() if cfg.sampled else torch.zeros((), dtype=f32, device=dev)   # eval: tensor(0.) (:489)

        if need_bwd:
            sv.shape, sv.cfg = shape, cfg
            sv.x_copy, sv.w, sv.Wf, sv.Wt, sv.Wc, sv.Wf_op, sv.Wt_op = x_copy, w, Wf, Wt, Wc, Wf_op, Wt_op
            sv.fp, sv.tp, sv.raw, sv.topk_idx, sv.gates = fp, tp, raw, topk_idx, gates
            sv.plan, sv.a_op, sv.row_token, sv.row_gate, sv.token_rows = plan, a_op, row_token, row_gate, token_rows
            sv.W1_op, sv.W2_op, sv.pre, sv.h_op, sv.y, sv.kl_out = W1_op, W2_op, pre, h_op, y, kl_out
            sv.eps = (eps_f, eps_t, eps_c)
            sv.router_params = router_params
            ctx.sv = sv
        # internal routing state for tests / diagnostics (not part of the reference interface)
        ctx_aux = {"topk_idx": topk_idx, "gates": gates, "plan": plan, "token_rows": token_rows, "row_token": row_token, "raw": raw}
        MoEBlockFunction.last_aux = ctx_aux
        return out, kl, probs, load

    @staticmethod
    def backward(ctx, dout, dkl, dprobs, dload):
        lib = C.lib()
        sv = ctx.sv
        cfg, shape = sv.cfg, sv.shape
        E, d, Dt, T, B = cfg.E, cfg.d, cfg.Dt, cfg.T, cfg.B
        dev = sv.w.device
        f32, bf16 = torch.float32, torch.bfloat16
        S = 2 if cfg.precision == 1 else 1
        hid = 4 * d
        m_cap = shape.m_cap
        act_dt = f32 if cfg.precision == 1 else bf16
        st = _stream()
        dout = torch.zeros((T, d), dtype=f32, device=dev) if dout is None else dout.contiguous().float()
        dprobs = None if dprobs is None else dprobs.contiguous().float()
        dload = None if dload is None else dload.contiguous().float()
        dkl = None if dkl is None else dkl.reshape(1).contiguous().float()

        # B1: combine backward
        dy_op = torch.empty((m_cap, S * d), dtype=bf16, device=dev)
        dgate_row = torch.empty((m_cap,), dtype=f32, device=dev)
        C.check(lib.moegan_combine_bwd(shape, C.ptr(sv.plan), C.ptr(dout), C.ptr(sv.y), C.ptr(sv.row_token), C.ptr(sv.row_gate),
                                       C.ptr(sv.gates), C.ptr(dy_op), C.ptr(dgate_row), st), "moegan_combine_bwd")
        # B2: experts backward
        dpre_op = torch.empty((m_cap, S * hid), dtype=bf16, device=dev)
        dx_rows = torch.empty((m_cap, d), dtype=act_dt, device=dev)
        dW1 = torch.empty((E, hid, d), dtype=f32, device=dev)
        db1 = torch.empty((E, hid), dtype=f32, device=dev)
        dW2 = torch.empty((E, d, hid), dtype=f32, device=dev)
        db2 = torch.empty((E, d), dtype=f32, device=dev)
        C.check(lib.moegan_ffn_bwd(shape, C.ptr(sv.plan), C.ptr(sv.a_op), C.ptr(sv.W1_op), C.ptr(sv.W2_op), C.ptr(sv.pre), C.ptr(sv.h_op),
                                   C.ptr(dy_op), C.ptr(dpre_op), C.ptr(dx_rows), C.ptr(dW1), C.ptr(db1), C.ptr(dW2), C.ptr(db2), st),
                "moegan_ffn_bwd")
        # B3-B6: router backward
        router_grads = [None] * 7
        dx_router = None
        dw = None
        if not cfg.hard:
            mu_f, rho_f, mu_t, rho_t, mu_c, rho_c, temperature = sv.router_params
            g = [torch.empty_like(p) for p in sv.router_params]
            dx_router = torch.empty((T, d), dtype=f32, device=dev)
            dw = torch.empty((B, Dt), dtype=f32, device=dev)
            ws_bytes = int(lib.moegan_router_bwd_workspace_bytes(shape))
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            eps_f, eps_t, eps_c = sv.eps
            C.check(lib.moegan_router_bwd(
                shape, C.ptr(sv.x_copy), C.ptr(sv.w), C.ptr(sv.Wf), C.ptr(sv.Wt), C.ptr(sv.Wc), C.ptr(sv.Wf_op), C.ptr(sv.Wt_op),
                C.ptr(sv.fp), C.ptr(sv.tp), C.ptr(sv.raw), C.ptr(sv.topk_idx), C.ptr(sv.token_rows), C.ptr(dgate_row),
                C.ptr(dprobs), C.ptr(dload), C.ptr(temperature), float(cfg.annealing_factor), 1 if cfg.sampled else 0,
                C.ptr(mu_f), C.ptr(rho_f), C.ptr(eps_f), C.ptr(mu_t), C.ptr(rho_t), C.ptr(eps_t), C.ptr(mu_c), C.ptr(rho_c), C.ptr(eps_c),
                C.ptr(sv.kl_out), C.ptr(dkl if cfg.sampled else None),
                C.ptr(g[0]), C.ptr(g[1]), C.ptr(g[2]), C.ptr(g[3]), C.ptr(g[4]), C.ptr(g[5]), C.ptr(g[6]),
                C.ptr(dx_router), C.ptr(dw), C.ptr(ws), ws_bytes, st), "moegan_router_bwd")
            router_grads = g
        # un-permute + router dx
        dx = torch.empty((T, d), dtype=f32, device=dev)
        C.check(lib.moegan_uncombine_bwd(shape, C.ptr(dx_rows), C.ptr(sv.token_rows), C.ptr(dx_router), C.ptr(dx), st),
                "moegan_uncombine_bwd")
        expert_grads = []
        for e in range(E):
            expert_grads += [dW1[e], db1[e], dW2[e], db2[e]]
        ctx.sv = None
        return (None, None, dx, dw, None, None, None, *router_grads, *expert_grads)


MoEBlockFunction.last_aux = None


class BalanceLossFunction(torch.autograd.Function):
    """bw * nan_to_num(clamp(E * cv(frac), 0, 10)) on the fused per-expert load (t2i_moe_gan.py:973-1000)."""

    @staticmethod
    def forward(ctx, load, T: int, balance_weight: float):
        _require_cuda(load, "load")
        load = load.contiguous().float()
        loss = torch.empty((1,), dtype=torch.float32, device=load.device)
        C.check(C.lib().moegan_balance_loss_fwd(C.ptr(load), load.numel(), int(T), float(balance_weight), C.ptr(loss), _stream()),
                "moegan_balance_loss_fwd")
        ctx.save_for_backward(load)
        ctx.T, ctx.bw = int(T), float(balance_weight)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        (load,) = ctx.saved_tensors
        dloss = dloss.reshape(1).contiguous().float()
        dload = torch.empty_like(load)
        C.check(C.lib().moegan_balance_loss_bwd(C.ptr(load), load.numel(), ctx.T, ctx.bw, C.ptr(dloss), C.ptr(dload), _stream()),
                "moegan_balance_loss_bwd")
        return dload, None, None
