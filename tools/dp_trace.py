#!/usr/bin/env python
"""Kernel-level timeline (CUPTI through torch.profiler) of ONE CUDA-graph replay of the train step, on 1 or N GPUs: per stream, when
every kernel started and how long it ran.  Printed: step span, the NCCL kernels (start, duration, what the compute stream was doing
meanwhile), time the compute stream sat idle, and the per-family kernel time - the same numbers for world = 1 and world = N say where
the data-parallel overhead is.

    python tools/dp_trace.py                                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_trace.py
"""
import importlib, json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
pkg = importlib.import_module("moe-gan_cpsc541_b200")
import bench
B = 256
torch.backends.cudnn.benchmark = True
G, D = bench.build_models(pkg, dev)
tr = pkg.MoEGANTrainer(G, D, amp_dtype=torch.bfloat16, use_cuda_graph=True)
torch.manual_seed(1 + rank)
real = torch.randn(B, 3, 64, 64, device=dev).clamp(-1, 1).contiguous(memory_format=torch.channels_last)
text = torch.randn(B, 512, device=dev)
for _ in range(6):
    tr.step(real, text)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
out = os.path.join(ROOT, "gpurun_out", f"dp_trace_w{world}_r{rank}.json")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.step(real, text)
    torch.cuda.synchronize()
prof.export_chrome_trace(out)
if world > 1:
    dist.barrier()
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel" and e.get("ph") == "X"]
os.remove(out)
ev.sort(key=lambda e: e["ts"])
# three identical replays were traced: per stream, the middle third of its kernels (by count) is the middle replay
streams = collections.Counter(e["args"].get("stream") for e in ev)
main = streams.most_common(1)[0][0]
step = []
for sid in streams:
    se = [e for e in ev if e["args"].get("stream") == sid]
    n = len(se) // 3
    step += se[n:2 * n]
step.sort(key=lambda e: e["ts"])
t0 = min(e["ts"] for e in step); t1 = max(e["ts"] + e["dur"] for e in step)
if rank == 0:
    print(f"## world {world}: one graph replay = {len(step)} kernels on {len(set(e['args'].get('stream') for e in step))} streams, span {(t1 - t0) / 1e3:.3f} ms")
    mstep = [e for e in step if e["args"].get("stream") == main]
    busy = sum(e["dur"] for e in mstep)
    idle = 0.0; last = t0
    for e in mstep:
        if e["ts"] > last: idle += e["ts"] - last
        last = max(last, e["ts"] + e["dur"])
    print(f"   compute stream {main}: {len(mstep)} kernels, busy {busy / 1e3:.3f} ms, idle gaps {idle / 1e3:.3f} ms")
    for e in step:
        if "nccl" in e["name"].lower():
            a, b = e["ts"], e["ts"] + e["dur"]
            ov = sum(max(0.0, min(b, m["ts"] + m["dur"]) - max(a, m["ts"])) for m in mstep)
            print(f"   NCCL  start {(a - t0) / 1e3:8.3f} ms  dur {e['dur']:8.1f} us  stream {e['args'].get('stream')}  grid {e['args'].get('grid')}  compute-stream kernels running during it: {ov:8.1f} us   {e['name'][:60]}")
    fam = collections.defaultdict(lambda: [0, 0.0])
    for e in step:
        n = e["name"].split("<")[0].split("(")[0][-60:]
        fam[n][0] += 1; fam[n][1] += e["dur"]
    print("   top kernel families (count, total us):")
    for n, (c, d) in sorted(fam.items(), key=lambda x: -x[1][1])[:25]:
        print(f"     {c:5d} {d:9.1f}  {n}")
    print(f"   sum of all kernel durations {sum(e['dur'] for e in step) / 1e3:.3f} ms")
# no destroy_process_group(): tearing NCCL down while a captured graph still holds its kernels hangs (see tools/nccl_probe.py)
sys.stdout.flush()
os._exit(0)
