#!/usr/bin/env python
"""Stand-alone NCCL all-reduce timing on this box (no compute running): what the gradient all-reduce of the train step can cost at best.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29515 tools/nccl_probe.py
"""
import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb, dt in ((111.1, torch.float32), (55.5, torch.bfloat16), (25.0, torch.float32), (2.4, torch.float32), (0.000016, torch.float32)):
    n = max(int(mb * 1e6 / torch.tensor([], dtype=dt).element_size()), 4)
    x = torch.ones(n, dtype=dt, device="cuda")
    for _ in range(5):
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        dist.all_reduce(x, op=dist.ReduceOp.AVG)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    if rank == 0:
        byt = n * x.element_size()
        print(f"all_reduce {byt / 1e6:9.3f} MB {str(dt):<15} world {world}: {us:9.1f} us  algbw {byt / us / 1e3:7.1f} GB/s  busbw {byt / us / 1e3 * 2 * (world - 1) / world:7.1f} GB/s")
# same, captured in a CUDA graph (how the train step issues it)
x = torch.ones(int(111.1e6 / 4), device="cuda")
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
for _ in range(3):
    g.replay()
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    g.replay()
b.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"all_reduce 111.1 MB fp32 inside a CUDA graph: {a.elapsed_time(b) / 20 * 1e3:.1f} us per replay", flush=True)
# a captured graph that holds NCCL kernels must be gone before the communicator is torn down (destroy_process_group hung otherwise);
# run this tool under `timeout` in any case
del g
torch.cuda.synchronize()
os._exit(0)
