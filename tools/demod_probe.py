#!/usr/bin/env python
"""demod_coef (library kernels) vs the torch-op chain it replaces, forward + backward, at the train step's layer shapes.
Run under `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/demod_probe.py`: the marker
kernels (fill of a 1-element tensor with 111 / 222) separate the two variants in the launch list."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("moe-gan_cpsc541_b200")
glue = pkg.glue
mark = torch.zeros(1, device="cuda")
for B, Cin, Cout, k in ((256, 512, 512, 3), (256, 256, 256, 1), (256, 128, 128, 3)):
    w = (torch.randn(Cout, Cin, k, k, device="cuda") * 0.05).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    s = torch.randn(B, Cin, device="cuda").bfloat16().requires_grad_(True)
    gd = torch.randn(B, Cout, device="cuda")
    for rep in range(2):
        mark.fill_(111.0)
        d = glue.demod_coef(s, w); d.backward(gd)
        mark.fill_(222.0)
        d = glue.demod_coef_reference(s, w); d.backward(gd)
        mark.fill_(333.0)
        s.grad = None; w.grad = None
torch.cuda.synchronize()
