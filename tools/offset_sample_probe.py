#!/usr/bin/env python
"""offset_sample forward + backward at the 16x16 / C = 128 layer of the train step, twice - the target of
    ncu --set full -k regex:offset_sample -s 3 -c 3 -o out python tools/offset_sample_probe.py"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("moe-gan_cpsc541_b200")
glue = pkg.glue
B, C, H = 256, 128, 16
x = torch.randn(B, C, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
g = torch.randn(B, C, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
off = torch.randn(B, 2, H, H, device="cuda").requires_grad_(True)
for _ in range(2):
    y = glue.offset_sample(x, off); y.backward(g); torch.cuda.synchronize()
