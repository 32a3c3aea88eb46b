#!/usr/bin/env python
"""bench.py - the driver's measurement contract for the MoE-GAN hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload moe_block]

One "step" = one forward + backward pass of the Bayesian-routed MoE block over one batch of synthetic tokens
(BASELINE.json configs[1]: 8 experts, top-2, 4096 tokens x 512 dim, fixed eps).  For N > 1 the driver starts one
rank per GPU with torch.distributed.run; every rank processes its own batch (weak scaling) and the parameter
gradients are all-reduced with NCCL inside the step.

Prints ONE JSON line (rank 0).  `value` = tokens/s with inputs resident in HBM; `e2e` = the same metric through
the public module call with HOST (pinned) inputs, H2D copy and a D2H loss read inside the timed region;
`roofline` = the tcgen05 grouped GEMM (dominant kernel) against the measured bf16 peak; `cpu_baseline` = the CPU
oracle port of the same step on the box's host cores.  `--impl reference` times that CPU port as the reference arm.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "moe-gan_cpsc541_b200"

# configs[1] of BASELINE.json
WL = dict(B=8, H=16, W=32, d=512, Dt=512, E=8, k=2, mu_scale=8.0, af=1.0)
METRIC = "moe_block_fwd_bwd_tokens_per_s"
UNIT = "tokens/s"


def workload_name():
    return (f"MoE block microbench (BASELINE configs[1]): E={WL['E']} top-{WL['k']}, T={WL['B'] * WL['H'] * WL['W']} tokens x "
            f"d={WL['d']}, fwd+bwd, fixed-shape synthetic tokens, router mu x{WL['mu_scale']:g}")


def algorithmic_flops():
    """experts fwd+bwd = 48 * T * k * d^2 (six equal grouped GEMMs)."""
    T = WL["B"] * WL["H"] * WL["W"]
    return 48.0 * T * WL["k"] * WL["d"] ** 2


def make_numpy_case(seed):
    import numpy as np
    from oracle import moe_oracle as O   # parameter *shapes/init* helper only; the oracle is not on the measured path
    rng = np.random.RandomState(seed)
    params = O.init_params(rng, WL["d"], WL["Dt"], WL["E"], mu_scale=WL["mu_scale"])
    x = rng.standard_normal((WL["B"], WL["d"], WL["H"], WL["W"])).astype(np.float32)
    w = rng.standard_normal((WL["B"], WL["Dt"])).astype(np.float32)
    dout = rng.standard_normal((WL["B"], WL["d"], WL["H"], WL["W"])).astype(np.float32)
    eps = O.draw_eps(rng, WL["d"], WL["Dt"], WL["E"])
    return params, x, w, dout, eps


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi fields via NVML), runs during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the same step (reference semantics, numpy fp32, all host threads)
# --------------------------------------------------------------------------------------------------
def cpu_step_time(reps, sample_tokens=None):
    import numpy as np
    from oracle import moe_oracle as O
    params, x, w, dout, eps = make_numpy_case(0)
    xt = O.nchw_to_tokens(x)
    do = O.nchw_to_tokens(dout)
    tps = WL["H"] * WL["W"]
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out, kl, probs, cache = O.moe_forward(xt, w, params, eps, WL["af"], True, top_k=WL["k"], tokens_per_sample=tps, dtype=np.float32)
        O.moe_backward(cache, params, eps, do, dtype=np.float32)
        times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2], xt.shape[0]


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    sec, T = cpu_step_time(max(1, args.steps), None) if args.steps else (None, None)
    for _ in range(max(0, args.warmup - 1)):
        pass
    val = T / sec
    cores = cpu_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name()},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"full step (T={T}) x {args.steps} reps, median; numpy fp32 oracle port of t2i_moe_gan.py:249-491 "
                                   "(the reference is pure Python/PyTorch and does not travel to the GPU box)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200_arm(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    pkg = importlib.import_module(PKG)
    C = pkg._cabi
    lib = C.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    params, x_np, w_np, dout_np, eps_np = make_numpy_case(1000 + rank)          # rank-offset data ...
    params, _, _, _, _ = make_numpy_case(0)                                      # ... identical weights on every rank
    moe = pkg.SparseMoE(WL["d"], WL["Dt"], WL["E"], top_k=WL["k"], precision="bf16")
    sd = {k: torch.from_numpy(v) for k, v in params.items()}
    for k in ("epsilon_f", "epsilon_t", "epsilon_c"):
        sd["router." + k] = moe.state_dict()["router." + k]
    moe.load_state_dict(sd)
    moe.to(dev).train()
    plist = [p for p in moe.parameters()]
    T = WL["B"] * WL["H"] * WL["W"]

    x_host = torch.from_numpy(x_np).pin_memory()
    w_host = torch.from_numpy(w_np).pin_memory()
    x_dev = x_host.to(dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True)   # NHWC memory, as norm3 produces (:561)
    w_dev = w_host.to(dev).requires_grad_(True)
    dout = torch.from_numpy(dout_np).to(dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    eps = tuple(torch.from_numpy(eps_np[k]).to(dev) for k in ("epsilon_f", "epsilon_t", "epsilon_c"))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)        # > 126 MB L2
    flat = torch.zeros(sum(p.numel() for p in plist), dtype=torch.float32, device=dev) if world > 1 else None

    def allreduce_grads():
        if world == 1:
            return
        off = 0
        for p in plist:
            n = p.numel()
            flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(flat)
        flat.mul_(1.0 / world)
        off = 0
        for p in plist:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p))
            off += n

    def step_device():
        for p in plist:
            p.grad = None
        x_dev.grad = None
        w_dev.grad = None
        out, kl, probs = moe(x_dev, w_dev, WL["af"], eps=eps)
        loss = (out * dout).sum() + 1e-3 * kl
        loss.backward()
        allreduce_grads()
        return loss

    def step_e2e():
        for p in plist:
            p.grad = None
        xd = x_host.to(dev, non_blocking=True).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True)
        wd = w_host.to(dev, non_blocking=True).requires_grad_(True)
        out, kl, probs = moe(xd, wd, WL["af"])                    # public call: eps redrawn on the device (:349-351)
        loss = (out * dout).sum() + 1e-3 * kl
        loss.backward()
        allreduce_grads()
        return float(loss.item())                                 # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        n0 = lib.moegan_debug_launch_count()
        host_t0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()                                          # L2 flush between timed iterations (not timed)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        host_ms = (time.perf_counter() - host_t0) * 1e3      # CPU time to enqueue the steps (no sync inside)
        barrier()
        n1 = lib.moegan_debug_launch_count()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        timed.host_ms = host_ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), int(n1 - n0)

    # The device-resident number replays ONE CUDA graph of the whole step (forward, backward, gradient all-reduce):
    # the step is ~30 launches of 2-80 us each, so eager Python launch latency would otherwise be what is measured.
    launch_mode = "eager"
    step_value = step_device
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step_device()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n_before = lib.moegan_debug_launch_count()
            with torch.cuda.graph(graph):
                step_device()
            launches_per_graph = int(lib.moegan_debug_launch_count() - n_before)
            step_value = graph.replay
            launch_mode = "cuda_graph"
        except Exception as e:  # noqa
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            torch.cuda.synchronize()
    with ClockSampler(local_rank) as clk:
        total_ms, launches = timed(step_value, args.steps, max(args.warmup, 3))
    if launch_mode == "cuda_graph":
        launches = launches_per_graph * args.steps
    clocks = clk.summary()
    host_enqueue_ms = timed.host_ms / args.steps
    e2e_ms, _ = timed(step_e2e, args.steps, max(2, args.warmup // 2))

    # ---- live GEMM timing for the roofline (CUDA events recorded by the library on the launching stream) ----
    lib.moegan_debug_profile_enable(1)
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        step_device()
    barrier()
    n_rec = 16 * args.steps
    ms_buf = (C.c_float * n_rec)()
    tag_buf = (C.c_int32 * n_rec)()
    n = lib.moegan_debug_profile_read(ms_buf, tag_buf, n_rec)
    lib.moegan_debug_profile_enable(0)
    per_step = n // args.steps if args.steps else 0
    gemm_ms = []
    for s in range(args.steps):
        gemm_ms += [ms_buf[s * per_step + i] for i in range(min(6, per_step))]     # the six expert GEMMs lead each step's router GEMMs
    expert_order = ["ffn1_fwd", "ffn2_fwd", "dgrad2", "dgrad1", "wgrad2", "wgrad1"]
    by_kernel = {}
    for s in range(args.steps):
        for i, name in enumerate(expert_order[:per_step]):
            by_kernel.setdefault(name, []).append(ms_buf[s * per_step + i])
    mean_gemm_ms = sum(gemm_ms) / max(len(gemm_ms), 1)
    flops_per_launch = algorithmic_flops() / 6.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    achieved_tf = flops_per_launch / (mean_gemm_ms * 1e-3) / 1e12 if mean_gemm_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("umma_gemm_dram_bytes_per_launch")
    except Exception:
        pass

    if rank == 0:
        cpu_sec, cpu_T = cpu_step_time(5)
        ms_per_step = total_ms / args.steps
        value = T * world / (ms_per_step * 1e-3)
        e2e_val = T * world / (e2e_ms / args.steps * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(), "tokens_per_gpu": T, "l2": "256 MiB buffer zeroed between timed iterations",
                       "parallelism": f"dp{world}" if world > 1 else "single", "launch": launch_mode, "router": "fp32", "experts": "bf16 in / fp32 acc (tcgen05)"},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(x_host.numel() * 4 + w_host.numel() * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "umma_gemm_kernel (six grouped expert GEMMs per step)",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; kernels timed individually)" if peaks else "fallback 1590",
                         "traffic": traffic, "flops_per_launch": flops_per_launch, "mean_launch_us": mean_gemm_ms * 1e3,
                         "per_kernel_us": {k: 1e3 * sum(v) / len(v) for k, v in by_kernel.items()}},
            "cpu_baseline": {"value": cpu_T / cpu_sec, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
                             "sample": f"full step (T={cpu_T}) x 5 reps, median {cpu_sec * 1e3:.0f} ms; numpy fp32 oracle port"},
            "moe_block_fwd_bwd_us": ms_per_step * 1e3,
            "host_enqueue_us_per_step": host_enqueue_ms * 1e3,
        }
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="moe_block")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200_arm(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
