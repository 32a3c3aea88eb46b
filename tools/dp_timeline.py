#!/usr/bin/env python
"""Per-stream CUDA-event timeline of ONE data-parallel train step (eager launches): when each gradient bucket became ready on the
compute stream, when its NCCL all-reduce ran on the communication stream, when backward ended and when the step could continue -
i.e. which collective is exposed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dp_timeline.py [batch]
"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
pkg = importlib.import_module("moe-gan_cpsc541_b200")
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.backends.cudnn.benchmark = True
G, D = bench.build_models(pkg, dev)
tr = pkg.MoEGANTrainer(G, D, amp_dtype=torch.bfloat16, use_cuda_graph=False)
torch.manual_seed(1 + rank)
real = torch.randn(B, 3, 64, 64, device=dev).clamp(-1, 1).contiguous(memory_format=torch.channels_last)
text = torch.randn(B, 512, device=dev)
for _ in range(4):
    tr.step(real, text)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ev = lambda: torch.cuda.Event(enable_timing=True)
marks = {}


def mark(name):
    marks[name] = ev(); marks[name].record()


tr.red_g.trace, tr.red_d.trace = [], []
z = torch.randn(B, 512, device=dev)
torch.cuda._sleep(30000000)        # queue the whole step behind a spin kernel: launch-bound eager code would otherwise distort the picture
mark("step_begin")
tr.d_step(real, text, z)
mark("d_step_end")
# G-step with marks around backward / finish (the body of MoEGANTrainer.g_step)
tr.opt_g.zero_grad(set_to_none=True)
with tr._autocast():
    fake16, fake8, kl, routing = G(z, text, return_intermediate=True, return_routing=True, annealing_factor=tr.annealing_factor)
    kl = torch.nan_to_num(torch.clamp(kl, max=50.0), nan=0.0, posinf=50.0, neginf=0.0)
    g_gan = pkg.host_model.generator_adv_loss(D(fake16, text))
    last = routing[-1]
    load = pkg.dp_allreduce_load(last._moegan_load, tr.group)
    balance = pkg.BalanceLossFunction.apply(load, last.shape[0] * world, tr.balance_weight)
    g_loss = g_gan + balance + tr.eff_kl_weight * kl
mark("g_forward_end")
tr.red_g.begin()
g_loss.backward()
mark("g_backward_end")
tr.red_g.finish()
mark("g_reduce_joined")
torch.nn.utils.clip_grad_norm_(G.parameters(), max_norm=0.8)
tr.opt_g.step()
mark("step_end")
torch.cuda.synchronize()
t0 = marks["step_begin"]
rel = lambda e: t0.elapsed_time(e)
if rank == 0:
    print(f"## one eager data-parallel train step, {world} GPUs, batch {B}/GPU, rank 0, times in ms from step begin (CUDA events)")
    for k, v in marks.items():
        print(f"  {k:<18} {rel(v):8.3f}")
    for name, red in (("D", tr.red_d), ("G", tr.red_g)):
        print(f"  -- {name} gradient buckets ({len(red.buckets)}): bytes, ready (compute stream), all-reduce begin / end (comm stream)")
        for r in red.trace:
            print(f"     bucket {r['bucket']:>2}  {r['bytes'] / 1e6:7.2f} MB  ready {rel(r['ready']):8.3f}  allreduce {rel(r['begin']):8.3f} -> {rel(r['end']):8.3f}"
                  f"  ({r['begin'].elapsed_time(r['end']) * 1e3:7.1f} us, {r['bytes'] / max(r['begin'].elapsed_time(r['end']), 1e-6) / 1e6:6.1f} GB/s)")
    if tr.red_g.trace:
        last_end = max(rel(r["end"]) for r in tr.red_g.trace)
        print(f"  exposed G all-reduce tail = last all-reduce end - backward end = {last_end - rel(marks['g_backward_end']):.3f} ms")
if world > 1:
    dist.destroy_process_group()
