#!/usr/bin/env python
"""aten-op level breakdown (with input shapes) of one eager train step: which host-model ops own the elementwise time."""
import importlib, os, sys
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    z = torch.randn(B, 512, device="cuda")
    tr.d_step(real, text, z); tr.g_step(text, z)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=int(os.environ.get("TOPK", "90")), max_name_column_width=40, max_shapes_column_width=90))
