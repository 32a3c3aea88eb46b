#!/usr/bin/env python
"""Kernel-level breakdown of one full train step (eager, no CUDA graph): every kernel aggregated by name, plus the
time under each top-level module of the generator / discriminator (torch.profiler record_function ranges)."""
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
real = torch.randn(B, 3, 64, 64, device="cuda").clamp(-1, 1).contiguous(memory_format=torch.channels_last); text = torch.randn(B, 512, device="cuda")
for _ in range(3): tr.step(real, text)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=False) as prof:
    z = torch.randn(B, 512, device="cuda")
    with torch.profiler.record_function("D_STEP"): tr.d_step(real, text, z)
    with torch.profiler.record_function("G_STEP"): tr.g_step(text, z)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
tot = 0.0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ev.name not in ("D_STEP", "G_STEP"):
        agg[ev.name][0] += ev.device_time; agg[ev.name][1] += 1; tot += ev.device_time
print(f"total kernel time {tot / 1e3:.2f} ms, {sum(v[1] for v in agg.values())} launches, {len(agg)} distinct kernels")
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOPK", "70"))]:
    print(f"{us:9.1f} us {100 * us / tot:5.1f}% {n:5d}x {us / n:8.1f}  {name[:150]}")
