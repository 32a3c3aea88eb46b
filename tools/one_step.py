#!/usr/bin/env python
"""One eager train step (batch 256, bf16) between cudaProfilerStart / Stop, after three warm-up steps:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/one_step.py
gives the launch list of exactly one step (every kernel, library and ATen / cuDNN)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("moe-gan_cpsc541_b200")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
G = pkg.host_model.AuroraGenerator().cuda().to(memory_format=torch.channels_last); D = pkg.host_model.AuroraDiscriminator().cuda().to(memory_format=torch.channels_last)
tr = pkg.MoEGANTrainer(G, D, amp_dtype=torch.bfloat16, use_cuda_graph=False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
real = torch.randn(B, 3, 64, 64, device="cuda").clamp(-1, 1).contiguous(memory_format=torch.channels_last); text = torch.randn(B, 512, device="cuda")
for _ in range(3): tr.step(real, text)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(real, text)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
