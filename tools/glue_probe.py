#!/usr/bin/env python
"""The element-wise / resampling kernels around the MoE block (channel_scale, offset_sample, upsample2x) at the train step's shapes,
forward + backward, a few repetitions each with the L2 flushed in between.  Meant to run under
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/glue_probe.py
so that each launch gets its own duration (the bench's per-kernel table brackets launcher calls with events: ~2 us of launch gap
in every figure, and offset_sample_bwd = two kernels).  Without ncu it prints CUDA-event times of forward and backward."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("moe-gan_cpsc541_b200")
glue = pkg.glue
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
REPS = int(os.environ.get("REPS", "3"))


def timed(name, shape, fwd, inputs, gout):
    tf, tb = [], []
    for _ in range(REPS):
        for t in inputs:
            t.grad = None
        flush.zero_(); torch.cuda._sleep(200000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); y = fwd(); b.record()
        flush.zero_(); torch.cuda._sleep(200000)
        c, d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c.record(); y.backward(gout); d.record(); torch.cuda.synchronize()
        tf.append(a.elapsed_time(b) * 1e3); tb.append(c.elapsed_time(d) * 1e3)
    tf.sort(); tb.sort()
    print(f"{name:16s} {shape}: fwd {tf[len(tf) // 2]:6.1f} us   bwd {tb[len(tb) // 2]:6.1f} us (events around the autograd calls)")


for B, C, H in ((256, 128, 16), (256, 256, 16), (256, 256, 8), (256, 512, 8), (256, 512, 4)):
    x = torch.randn(B, C, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = torch.randn(B, C, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    s = (torch.randn(B, C, device="cuda") * 0.5 + 1.0).requires_grad_(True)
    off = torch.randn(B, 2, H, H, device="cuda").requires_grad_(True)
    timed("channel_scale", (B, C, H, H), lambda: glue.channel_scale(x, s, 0.2), (x, s), g)
    timed("offset_sample", (B, C, H, H), lambda: glue.offset_sample(x, off), (x, off), g)
    if H <= 8:
        g2 = torch.randn(B, C, 2 * H, 2 * H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
        timed("upsample2x", (B, C, H, H), lambda: glue.upsample2x_bilinear(x), (x,), g2)
