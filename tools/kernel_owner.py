#!/usr/bin/env python
"""Which aten op (with input shapes) launches a given kernel: python tools/kernel_owner.py <substring> [<substring> ...]"""
import importlib, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
pkg = importlib.import_module("moe-gan_cpsc541_b200")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
G = pkg.host_model.AuroraGenerator().cuda().to(memory_format=torch.channels_last); D = pkg.host_model.AuroraDiscriminator().cuda().to(memory_format=torch.channels_last)
tr = pkg.MoEGANTrainer(G, D, amp_dtype=torch.bfloat16, use_cuda_graph=False)
B = 256
real = torch.randn(B, 3, 64, 64, device="cuda").clamp(-1, 1).contiguous(memory_format=torch.channels_last); text = torch.randn(B, 512, device="cuda")
for _ in range(3): tr.step(real, text)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
    tr.step(real, text); torch.cuda.synchronize()
pats = sys.argv[1:] or ["nchwToNhwc", "nhwcToNchw"]
for pat in pats:
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CPU or not ev.kernels:
            continue
        if ev.cpu_children and any(c.kernels for c in ev.cpu_children):
            continue                                   # attribute to the innermost op
        for k in ev.kernels:
            if pat in k.name:
                key = (ev.name, str(ev.input_shapes)[:150])
                agg[key][0] += 1; agg[key][1] += k.duration
    print(f"== kernels matching '{pat}'")
    for (name, shapes), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"  {us:8.1f} us {n:4d}x  {name}  {shapes}")
