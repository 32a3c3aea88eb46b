"""Kernel-time breakdown of one eager train step (torch.profiler / CUPTI). Run on the GPU box."""
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=False) as prof:
    z = torch.randn(B, 512, device="cuda")
    with torch.profiler.record_function("D_STEP"): tr.d_step(real, text, z)
    with torch.profiler.record_function("G_STEP"): tr.g_step(text, z)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
