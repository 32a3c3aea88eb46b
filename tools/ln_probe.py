#!/usr/bin/env python
"""LayerNorm forward / backward kernels of the library at the train step's shapes: CUDA-event times (L2 flushed) and GB/s.
    python tools/ln_probe.py            (ncu: ncu --set full -k regex:layernorm_bwd -c 2 python tools/ln_probe.py)"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("moe-gan_cpsc541_b200")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rows, d in ((65536, 128), (16384, 256), (4096, 512)):
    for xdt, odt in ((torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32)):
        ln = torch.nn.LayerNorm(d).cuda()
        x = torch.randn(rows, d, device="cuda").to(xdt).requires_grad_(True)
        g = torch.randn(rows, d, device="cuda").to(odt)
        tf, tb = [], []
        for _ in range(8):
            flush.zero_(); torch.cuda._sleep(1000000)
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record(); y = pkg.glue.layer_norm(x, ln.weight, ln.bias, ln.eps, odt); b.record()
            flush.zero_(); torch.cuda._sleep(1000000)
            b2 = torch.cuda.Event(enable_timing=True); b2.record()
            y.backward(g); c.record(); torch.cuda.synchronize()
            tf.append(a.elapsed_time(b) * 1e3); tb.append(b2.elapsed_time(c) * 1e3)
            x.grad = None; ln.zero_grad()
        tf.sort(); tb.sort()
        es_x, es_o = x.element_size(), g.element_size()
        bf, bb = rows * d * (es_x + es_o), rows * d * (es_o + 2 * es_x)
        print(f"rows={rows} d={d} x={xdt} y/g={odt}: fwd {tf[len(tf)//2]:.1f} us ({bf / tf[len(tf)//2] / 1e3:.0f} GB/s)  bwd {tb[len(tb)//2]:.1f} us ({bb / tb[len(tb)//2] / 1e3:.0f} GB/s, incl. autograd overhead + the dgamma/dbeta finishing launch)")
