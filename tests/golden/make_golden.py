#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (CPU, fp32).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

The reference module imports ``clip`` at file scope (moegan/t2i_moe_gan.py:10); the package is
absent offline and nothing on the MoE path calls it, so an empty stub module is pre-seeded.
The scalar test loss is
    L = sum(out * dout) + KL_COEF * kl + BAL_COEF * moe_balance_loss([probs]) + sum(probs * gp)
so every output of SparseMoE.forward (out, kl, probs) carries gradient, and the reference's own
AuroraGANLoss.moe_balance_loss (t2i_moe_gan.py:951-1000) is on the graph.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from cases import CASES, BALANCE_WEIGHT, KL_COEF, BAL_COEF, make_inputs  # noqa: E402

REF = "/root/reference/moegan"


def import_reference():
    sys.modules.setdefault("clip", types.ModuleType("clip"))
    sys.path.insert(0, REF)
    import t2i_moe_gan as ref  # noqa
    return ref


def main():
    ref = import_reference()
    torch.set_num_threads(4)
    for name, c in CASES.items():
        c, params, x, w, dout, gp = make_inputs(name)
        moe = ref.SparseMoE(c["d"], c["Dt"], num_experts=c["E"])
        sd = {k: torch.from_numpy(v.copy()) for k, v in params.items()}
        for k in ("epsilon_f", "epsilon_t", "epsilon_c"):
            sd["router." + k] = moe.state_dict()["router." + k]
        moe.load_state_dict(sd, strict=True)
        loss_obj = ref.AuroraGANLoss.__new__(ref.AuroraGANLoss)   # __init__ would load CLIP (:915)
        loss_obj.device = torch.device("cpu")
        xt = torch.from_numpy(x).requires_grad_(True)
        wt = torch.from_numpy(w).requires_grad_(True)
        out = {}
        torch.manual_seed(c["seed"])
        if c["mode"] == "train":
            moe.train()
            o, kl, probs = moe(xt, wt, c["af"])
            bal = loss_obj.moe_balance_loss([probs], balance_weight=BALANCE_WEIGHT)
            L = (o * torch.from_numpy(dout)).sum() + KL_COEF * kl + BAL_COEF * bal + (probs * torch.from_numpy(gp)).sum()
            L.backward()
            out["loss"] = L.detach().numpy()
            out["balance"] = bal.detach().numpy()
            out["dx"] = xt.grad.numpy()
            out["dw"] = wt.grad.numpy()
            for k, p in moe.named_parameters():
                out["grad." + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        else:
            moe.eval()
            with torch.no_grad():
                o, kl, probs = moe(xt, wt, c["af"])
        out["out"] = o.detach().contiguous().numpy()
        out["out_strides"] = np.array(o.stride(), dtype=np.int64)
        out["kl"] = np.asarray(kl.detach().numpy(), dtype=np.float32)
        out["probs"] = probs.detach().numpy()
        for k in ("epsilon_f", "epsilon_t", "epsilon_c"):
            out[k] = getattr(moe.router, k).detach().numpy().copy()
        # also the router's own (probs, logits) return, via a direct call with the same eps:
        # not possible without re-drawing eps (forward overwrites the buffers, :349-351), so the
        # logits are recomputed by the oracle and only probs are pinned here.
        np.savez(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k in ("out", "probs", "dx")}, "kl", float(out["kl"]))


if __name__ == "__main__":
    main()
