"""cProfile of the eager public-API step (host overhead hunt). Run on the GPU box."""
import cProfile, pstats, importlib, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module(bench.PKG)
params, x_np, w_np, dout_np, eps_np = bench.make_numpy_case(0)
WL = bench.WL
moe = pkg.SparseMoE(WL["d"], WL["Dt"], WL["E"], top_k=WL["k"], precision="bf16")
sd = {k: torch.from_numpy(v) for k, v in params.items()}
for k in ("epsilon_f", "epsilon_t", "epsilon_c"):
    sd["router." + k] = moe.state_dict()["router." + k]
moe.load_state_dict(sd); moe.cuda().train()
plist = list(moe.parameters())
x = torch.from_numpy(x_np).cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True)
w = torch.from_numpy(w_np).cuda().requires_grad_(True)
dout = torch.from_numpy(dout_np).cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
def step():
    for p in plist: p.grad = None
    out, kl, probs = moe(x, w, 1.0)
    loss = (out * dout).sum() + 1e-3 * kl
    loss.backward()
for _ in range(20): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
