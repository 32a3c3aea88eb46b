"""Helpers shared by the GPU parity tests: run the CUDA module and the numpy oracle on identical inputs."""
import importlib

import numpy as np
import torch

from oracle import moe_oracle as O

PKG = "moe-gan_cpsc541_b200"


def pkg():
    p = importlib.import_module(PKG)
    p.MoEBlockFunction.keep_last = True      # the parity tests read the routing state out of the last forward workspace
    return p


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / den


def build_module(params, d, Dt, E, top_k=None, precision="fp32", training=True):
    m = pkg().SparseMoE(d, Dt, E, top_k=top_k, precision=precision)
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()}
    for k in ("epsilon_f", "epsilon_t", "epsilon_c"):
        sd["router." + k] = m.state_dict()["router." + k]
    m.load_state_dict(sd, strict=True)
    m.cuda()
    m.train(training)
    return m


def run_cuda(m, x, w, eps, af, dout=None, gp=None, kl_coef=0.0, bal_coef=0.0, balance_weight=0.01, fused_balance=True):
    """x [B,d,H,W] numpy.  Returns dict of numpy outputs (+ grads when dout is given)."""
    p = pkg()
    xt = torch.from_numpy(x).cuda().requires_grad_(dout is not None)
    wt = torch.from_numpy(w).cuda().requires_grad_(dout is not None)
    e = None
    if eps is not None and m.training:
        e = tuple(torch.from_numpy(eps[k]).cuda() for k in O.EPS_KEYS)
    res = {}
    if dout is None:
        with torch.no_grad():
            out, kl, probs = m(xt, wt, af, eps=e)
    else:
        m.zero_grad(set_to_none=True)
        out, kl, probs = m(xt, wt, af, eps=e)
        L = (out * torch.from_numpy(dout).cuda()).sum()
        if kl_coef:
            L = L + kl_coef * kl
        if gp is not None:
            L = L + (probs * torch.from_numpy(gp).cuda()).sum()
        if bal_coef:
            if fused_balance:
                bal = p.moe_balance_loss([probs], balance_weight)
            else:   # generic path: gradient arrives as a dense [T,E] dprobs
                load = probs.sum(dim=0)
                bal = p.BalanceLossFunction.apply(load, probs.shape[0], balance_weight)
            res["balance"] = float(bal.detach().cpu())
            L = L + bal_coef * bal
        L.backward()
        res["loss"] = float(L.detach().cpu())
        res["dx"] = xt.grad.detach().cpu().numpy()
        res["dw"] = wt.grad.detach().cpu().numpy()
        for k, prm in m.named_parameters():
            res["grad." + k] = prm.grad.detach().cpu().numpy() if prm.grad is not None else np.zeros(tuple(prm.shape), np.float32)
    torch.cuda.synchronize()
    res["out"] = out.detach().contiguous().cpu().numpy()
    res["out_strides"] = tuple(out.stride())
    res["kl"] = float(kl.detach().cpu())
    res["probs"] = probs.detach().cpu().numpy()
    aux = p.MoEBlockFunction.last_aux()
    res["aux"] = {k: (v.detach().cpu().numpy() if v is not None else None) for k, v in aux.items()}
    return res


def run_oracle(params, x, w, eps, af, training=True, top_k=None, dout=None, gp=None, kl_coef=0.0, bal_coef=0.0, balance_weight=0.01):
    B, d, H, W = x.shape
    xt = O.nchw_to_tokens(x)
    out, kl, probs, cache = O.moe_forward(xt, w, params, eps, af=af, training=training, top_k=top_k, tokens_per_sample=H * W)
    res = {"out": O.tokens_to_nchw(out, B, H, W), "kl": kl, "probs": probs, "cache": cache}
    if dout is not None:
        dprobs = np.zeros_like(probs)
        if gp is not None:
            dprobs = dprobs + gp
        if bal_coef:
            res["balance"] = O.balance_loss(probs, balance_weight)
            dprobs = dprobs + O.balance_loss_bwd(probs, balance_weight, dloss=bal_coef)
        grads = O.moe_backward(cache, params, eps, O.nchw_to_tokens(dout), dprobs=dprobs, dkl=kl_coef)
        res["dx"] = O.tokens_to_nchw(grads.pop("dx"), B, H, W)
        res["dw"] = grads.pop("dw")
        for k, v in grads.items():
            res["grad." + k] = v
    return res


def compare_grads(got, ref, params, tol, skip=()):
    errs = {}
    for key in ["dx", "dw"] + ["grad." + k for k in params]:
        if key in skip:
            continue
        r = np.asarray(ref[key]); g = np.asarray(got[key])
        if np.abs(r).max() == 0.0:
            errs[key] = float(np.abs(g).max())
            assert errs[key] < 1e-6, (key, errs[key])
        else:
            errs[key] = rel_err(g, r)
            # a single scalar (router.temperature) is a heavily cancelling sum over all tokens: with bf16
            # activations its relative error is not averaged down like a tensor's max-norm error is
            lim = 3.0 * tol if (r.size == 1 and tol > 1e-3) else tol
            assert errs[key] < lim, (key, errs[key])
    return errs
