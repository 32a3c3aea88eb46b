#!/usr/bin/env python
"""Timeline (start, duration, stream) of every kernel in ONE CUDA-graph replay of the configs[1] MoE block step, from CUPTI
through torch.profiler - shows what overlaps with what and where the gaps are."""
import importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from torch.profiler import profile, ProfilerActivity
pkg = importlib.import_module("moe-gan_cpsc541_b200")
WL = bench.WL
dev = torch.device("cuda", 0)
params, x_np, w_np, dout_np, eps_np = bench.make_numpy_case(0)
moe = pkg.SparseMoE(WL["d"], WL["Dt"], WL["E"], top_k=WL["k"], precision="bf16")
sd = {k: torch.from_numpy(v) for k, v in params.items()}
for k in ("epsilon_f", "epsilon_t", "epsilon_c"):
    sd["router." + k] = moe.state_dict()["router." + k]
moe.load_state_dict(sd); moe.to(dev).train()
plist = list(moe.parameters())
nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
x = nhwc(torch.from_numpy(x_np).to(dev)).requires_grad_(True); w = torch.from_numpy(w_np).to(dev).requires_grad_(True)
dout = nhwc(torch.from_numpy(dout_np).to(dev))
eps = tuple(torch.from_numpy(eps_np[k]).to(dev) for k in ("epsilon_f", "epsilon_t", "epsilon_c"))
kl_seed = torch.full((), 1e-3, device=dev)
def step():
    for p in plist: p.grad = None
    x.grad = None; w.grad = None
    out, kl, probs = moe(x, w, WL["af"], eps=eps)
    torch.autograd.backward([out, kl], [dout, kl_seed])
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): step()
for _ in range(5): g.replay()
torch.cuda.synchronize()
NREP = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NREP): g.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
per = len(evs) // NREP
evs = evs[(NREP - 1) * per:]                 # the last of NREP back-to-back replays (no profiler start-up effects)
t0 = evs[0].time_range.start
print(f"{len(evs)} kernels, span {(max(e.time_range.end for e in evs) - t0):.1f} us")
for e in evs:
    print(f"{e.time_range.start - t0:8.1f} +{e.time_range.end - e.time_range.start:6.1f}  {e.name[:90]}")
