#!/usr/bin/env python
"""bench.py - the driver's measurement contract for the MoE-GAN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload train_step|moe_block] [--no-graph]

Default workload `train_step` (BASELINE.json configs[2]; at N GPUs it is configs[3]'s data-parallel step):
one "step" = one full MoE-GAN train iteration - D-step (adv + matching-aware + R1) and G-step (adv + KL + balance),
both AdamW updates - on a batch of 256 synthetic 64x64 images + synthetic CLIP text embeddings PER GPU, bf16
autocast, fp32 router, generator MoE blocks on the sm_100a kernels.  Weak scaling: 256 images per GPU, gradients
all-reduced with NCCL (bucketed, overlapped with backward).  metric = train images/s (whole job).

`moe_block` = BASELINE configs[1] (8 experts, top-2, 4096 tokens x 512, fwd+bwd).  It and configs[4] (inference:
generate 1024 images, eval-mode hard top-1 routing) are also run briefly inside the default workload and reported
under "moe_block" / "inference" in the same JSON line.

JSON keys follow the contract: `value` has inputs resident in HBM; `e2e` goes through the public call with pinned
HOST buffers (H2D inside, D2H loss read inside); `roofline` is the tcgen05 grouped-GEMM family timed live with CUDA
events on its launch stream; `cpu_baseline` is the CPU port (torch-CPU restatement of the reference MoE block inside the same PyTorch host
model) on a bounded sample; `--impl reference` prints that CPU arm as its own line.
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

# BASELINE configs[1]
WL = dict(B=8, H=16, W=32, d=512, Dt=512, E=8, k=2, mu_scale=8.0, af=1.0)
TRAIN_BATCH = 256          # per GPU (configs[2]); 8 GPUs -> global 2048 (configs[3])
CPU_BATCH = 8              # configs[0]: the reference's own CPU-runnable case
INFER_BATCH = 1024         # configs[4]
GLOBAL_BATCH_DP = 2048      # configs[3]
# how the parity the numbers rest on is defined (tests/, -m gpu): stated in the bench line so a reader need not open the tests
PARITY_NOTE = {"metric": "max|a-b| / max|b| per tensor (max-norm relative error)", "fp32_mode_bar": 1e-4, "bf16_mode_bar": 2e-2,
               "routing": "top-k indices / permutation bit-exact; tokens whose top-k logit gap is < 1e-6 are listed and excluded",
               "exception": "scalar router.temperature gradient is allowed 3x the bar in bf16 mode (a cancelling sum over all tokens)",
               "against": "unmodified reference on the same GPU (tests/test_gpu_reference_dropin.py), golden fixtures from the reference (tests/golden), numpy oracle"}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def make_numpy_case(seed, WL=None):
    import numpy as np
    from oracle import moe_oracle as O   # parameter shapes / init helper only; the oracle is never on the measured GPU path
    WL = WL or globals()["WL"]
    rng = np.random.RandomState(seed)
    params = O.init_params(rng, WL["d"], WL["Dt"], WL["E"], mu_scale=WL["mu_scale"])
    x = rng.standard_normal((WL["B"], WL["d"], WL["H"], WL["W"])).astype(np.float32)
    w = rng.standard_normal((WL["B"], WL["Dt"])).astype(np.float32)
    dout = rng.standard_normal((WL["B"], WL["d"], WL["H"], WL["W"])).astype(np.float32)
    eps = O.draw_eps(rng, WL["d"], WL["Dt"], WL["E"])
    return params, x, w, dout, eps


# --------------------------------------------------------------------------------------------------
# clocks sampler (the nvidia-smi fields, via NVML), runs during the timed region
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
            self.h = self._handle(pynvml, index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    @staticmethod
    def _handle(pynvml, index):
        """NVML handle of CUDA device `index` of THIS process: NVML numbers the physical GPUs, CUDA numbers the ones CUDA_VISIBLE_DEVICES
        leaves visible - look the device up by UUID (else through CUDA_VISIBLE_DEVICES, else by the same index)."""
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            return pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        except Exception:
            pass
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if ids and ids[index].isdigit():
                return pynvml.nvmlDeviceGetHandleByIndex(int(ids[index]))
        except Exception:
            pass
        return pynvml.nvmlDeviceGetHandleByIndex(index)

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


def cpu_threads():
    try:
        import torch
        return int(torch.get_num_threads())
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py) through one iteration of its own batch loop
# (oracle/ref_step.py restates t2i_moe_gan.py:1262-1421 around the reference's own modules); the torch-CPU port is only the
# fallback for a tree that was never built next to /root/reference
# --------------------------------------------------------------------------------------------------
def reference_available():
    try:
        from oracle import ref_loader
        return ref_loader.available()
    except Exception:
        return False


def pick_cpu_threads(make_step, candidates=None):
    """The reference arm gets the thread count that is FASTEST for it on this host: batch-8 fp32 ops do not scale to hundreds of
    hardware threads, so "all the host threads it can use" is found by timing one step at a few settings (launchers such as
    torch.distributed.run export OMP_NUM_THREADS=1 - the count is always set explicitly here)."""
    import torch
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (candidates or (ncpu, ncpu // 2, 64, 32, 16, 8)) if 1 <= c <= ncpu}, reverse=True)
    best, best_t, tried = cands[-1], float("inf"), {}
    step = make_step()
    torch.set_num_threads(cands[0]); step()                # warm-up (allocator, lazy inits)
    for c in cands:
        torch.set_num_threads(c)
        step()
        t0 = time.perf_counter(); step(); t = time.perf_counter() - t0
        tried[c] = round(t, 3)
        if t < best_t:
            best, best_t = c, t
    torch.set_num_threads(best)
    return best, tried, step


def cpu_train_step_time(reps, batch=CPU_BATCH, warmup=1):
    """One iteration of t2i_moe_gan.py:1262-1421 (no CLIP terms) on the host cores, fp32, batch 8 (configs[0]).
    Returns (median seconds, batch, info dict with kind / cores / threads tried)."""
    import torch
    if reference_available():
        from oracle.ref_step import ReferenceTrainStep
        real = torch.randn(batch, 3, 64, 64).clamp(-1, 1)
        text = torch.randn(batch, 512)

        def make_step():
            tr = ReferenceTrainStep("cpu", amp_dtype=None, checkpoint_activation=False, seed=0)
            return lambda: tr.step(real, text)
        cores, tried, step = pick_cpu_threads(make_step)
        kind = "reference"
        what = ("the UNMODIFIED reference (oracle/_ref/t2i_moe_gan.py: AuroraGenerator / AuroraDiscriminator / AuroraGANLoss) through one iteration of "
                "its own batch loop (t2i_moe_gan.py:1262-1421 restated in oracle/ref_step.py, CLIP terms :1384-1387 omitted as in the B200 arm), "
                "activation checkpointing off")
    else:
        step = _port_train_step(batch)
        cores, tried, kind = cpu_threads(), {}, "port"
        what = ("oracle/_ref not staged in this tree: torch-CPU PORT of the reference step (host model + oracle/torch_bridge.py), not the reference itself")
    for _ in range(warmup):
        step()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2], batch, {"kind": kind, "cores": cores, "cores_available": os.cpu_count(), "threads_tried_s_per_step": tried, "what": what}


def _port_train_step(batch):
    import torch
    pkg = importlib.import_module(PKG)
    hm = pkg.host_model
    from oracle.torch_bridge import TorchSparseMoE
    torch.manual_seed(0)
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    G, D = hm.AuroraGenerator(), hm.AuroraDiscriminator()
    for blk in (G.gen_block_4, G.gen_block_8, G.gen_block_16):
        o = TorchSparseMoE(blk.attn_block.moe.dim, 512)
        o.load_state_dict(blk.attn_block.moe.state_dict())
        blk.attn_block.moe = o
    opt_g = torch.optim.AdamW(G.parameters(), lr=2e-4, betas=(0.5, 0.999), weight_decay=0.01)
    opt_d = torch.optim.AdamW(D.parameters(), lr=2e-4, betas=(0.5, 0.999), weight_decay=0.01)
    real = torch.randn(batch, 3, 64, 64).clamp(-1, 1)
    text = torch.randn(batch, 512)
    af, eff_kl = 3.0, 1e-3 * 1e-5

    def balance(probs, bw=0.01):
        load = probs.sum(0) + 1e-6
        frac = load / probs.shape[0]
        cv = frac.std() / (frac.mean() + 1e-6)
        return bw * torch.nan_to_num(torch.clamp(probs.shape[1] * cv, 0.0, 10.0), nan=0.0)

    def step():
        z = torch.randn(batch, 512)
        opt_d.zero_grad()
        r = real.clone().requires_grad_(True)
        real_pred = D(r, text)
        r1 = hm.r1_penalty(real_pred, r, 10.0)
        with torch.no_grad():
            fake16, _, _, _ = G(z, text, return_intermediate=True, return_routing=True, annealing_factor=af)
        d_loss = hm.discriminator_adv_loss(real_pred, D(fake16.detach(), text), D(r.detach(), text[torch.randperm(batch)])) + r1
        d_loss.backward()
        torch.nn.utils.clip_grad_norm_(D.parameters(), 0.7)
        opt_d.step()
        opt_g.zero_grad()
        fake16, fake8, kl, routing = G(z, text, return_intermediate=True, return_routing=True, annealing_factor=af)
        kl = torch.clamp(kl, max=50.0)
        g_loss = hm.generator_adv_loss(D(fake16, text)) + balance(routing[-1]) + eff_kl * kl
        g_loss.backward()
        torch.nn.utils.clip_grad_norm_(G.parameters(), 0.8)
        opt_g.step()
    return step


def cpu_baseline_record(sec, units, info, reps):
    return {"value": units / sec, "unit": "imgs/s", "cores": info["cores"], "cores_available": info["cores_available"], "kind": info["kind"],
            "threads_tried_s_per_step": info["threads_tried_s_per_step"],
            "sample": (f"batch {units} (BASELINE configs[0]) x {reps} reps, median {sec * 1e3:.0f} ms/step, fp32, {info['cores']} host threads "
                       f"(fastest of those tried): {info['what']}")}


def cpu_moe_block_time(reps):
    import numpy as np
    from oracle import moe_oracle as O
    params, x, w, dout, eps = make_numpy_case(0)
    xt, do = O.nchw_to_tokens(x), O.nchw_to_tokens(dout)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out, kl, probs, cache = O.moe_forward(xt, w, params, eps, WL["af"], True, top_k=WL["k"], tokens_per_sample=WL["H"] * WL["W"], dtype=np.float32)
        O.moe_backward(cache, params, eps, do, dtype=np.float32)
        times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2], xt.shape[0]


def moe_block_workload_name(WL=None):
    WL = WL or globals()["WL"]
    route = f"top-{WL['k']}" if WL.get("k") else "dense-soft k=E (reference training path)"
    return (f"MoE block microbench ({WL.get('tag', 'BASELINE configs[1]')}): E={WL['E']} {route}, T={WL['B'] * WL['H'] * WL['W']} tokens x "
            f"d={WL['d']}, fwd+bwd, synthetic tokens, router mu x{WL['mu_scale']:g}")


def train_workload_name(world):
    return (f"MoE-GAN full train step (BASELINE configs[2]{'/[3]' if world > 1 else ''}): generator+discriminator, adv + R1 + KL + balance losses, "
            f"64x64 synthetic images + synthetic CLIP text embeddings, batch {TRAIN_BATCH}/GPU, bf16 autocast, fp32 router, AdamW x2")


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores (rank 0 only; the other
    ranks of a torchrun launch exit 0 without work).  Same metric / unit / workload family as the B200 arm; each step is a BOUNDED
    SAMPLE of that workload (batch 8 = BASELINE configs[0], fp32) so the run ends within a few minutes - stated in `config`."""
    if rank != 0:
        return
    if args.workload == "moe_block":
        sec, units = cpu_moe_block_time(max(1, args.steps))
        metric, unit = "moe_block_fwd_bwd_tokens_per_s", "tokens/s"
        config = {"workload": moe_block_workload_name(), "reference_arm": "numpy fp32 oracle port of t2i_moe_gan.py:249-491 (the reference has no top-k)",
                  "same_config": True}
        cpu = {"value": units / sec, "unit": unit, "cores": cpu_threads(), "kind": "port",
               "sample": f"full step (T={units}) x {args.steps} reps, median; numpy fp32 oracle port of t2i_moe_gan.py:249-491"}
    else:
        reps = max(1, args.steps)
        sec, units, info = cpu_train_step_time(reps, warmup=max(1, min(args.warmup, 2)))
        metric, unit = "train_imgs_per_s", "imgs/s"
        cpu = cpu_baseline_record(sec, units, info, reps)
        config = {"workload": train_workload_name(world),
                  "reference_arm": (f"host CPU, fp32, batch {units} per step (BASELINE configs[0]) = a bounded sample of the B200 arm's batch-{TRAIN_BATCH} bf16 workload; "
                                    f"kind '{info['kind']}', {info['cores']} threads on rank 0 only"),
                  "same_config": False, "batch_per_step": units, "device": "cpu", "precision": "fp32"}
    val = units / sec
    line = {
        "impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": cpu,
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# shared timing helpers (GPU)
# --------------------------------------------------------------------------------------------------
class Timer:
    def __init__(self, torch, dist, world, dev, lib):
        self.torch, self.dist, self.world, self.dev, self.lib = torch, dist, world, dev, lib
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn, steps, warmup):
        """W untimed + K timed steps; per-step CUDA events on the launching stream, L2 flushed between steps; MAX over ranks."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        evs = []
        n0 = self.lib.moegan_debug_launch_count()
        h0 = time.perf_counter()
        for _ in range(steps):
            self.flush.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        host_ms = (time.perf_counter() - h0) * 1e3
        self.barrier()
        n1 = self.lib.moegan_debug_launch_count()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item()), int(n1 - n0), host_ms

    def lead(self, lead_ms):
        """Keep the GPU busy for ~lead_ms (a spin kernel) so that the host enqueues the profiled step AHEAD of the device: eager
        launches are host-bound here (~5 us of Python per launch), and an event pair around a kernel that the GPU has to wait for
        measures the wait, not the kernel.  With the queue full, begin -> end is the kernel plus the inter-kernel gap only."""
        if lead_ms <= 0:
            return
        try:
            khz = self.torch.cuda.get_device_properties(self.dev).clock_rate
        except Exception:
            khz = 1965000
        self.torch.cuda._sleep(int(lead_ms * khz))

    def gemm_profile(self, fn, steps, lead_ms=0.0):
        """Durations (ms) of every tcgen05 GEMM launch in `steps` eager steps, measured by CUDA events the library
        records around the launch on its own stream."""
        C = importlib.import_module(PKG)._cabi
        self.lib.moegan_debug_profile_enable(1)
        self.lib.moegan_debug_set_side_streams(0)      # each GEMM alone on the stream: no router kernels competing for its SMs
        self.barrier()
        for _ in range(steps):
            self.flush.zero_()
            self.lead(lead_ms)
            fn()
        self.barrier()
        cap = 4096 * max(steps, 1)
        ms_buf = (C.c_float * cap)()
        tag_buf = (C.c_int32 * cap)()
        n = self.lib.moegan_debug_profile_read(ms_buf, tag_buf, cap)
        self.lib.moegan_debug_profile_enable(0)
        self.lib.moegan_debug_set_side_streams(-1)
        return [ms_buf[i] for i in range(n)], [tag_buf[i] for i in range(n)]

    def kernel_profile(self, fn, steps, lead_ms=0.0, serial=True):
        """[(ms, tag, meta8)] of EVERY launch of the library in `steps` eager steps (level-2 profiling: CUDA events recorded by
        the library on the launching stream around each kernel)."""
        C = importlib.import_module(PKG)._cabi
        self.lib.moegan_debug_profile_enable(2)
        if serial:      # one stream: a bracket around a side-stream kernel would include its wait for SMs held by a GEMM
            self.lib.moegan_debug_set_side_streams(0)
        self.barrier()
        for _ in range(steps):
            self.flush.zero_()
            self.lead(lead_ms)
            fn()
        self.barrier()
        cap = 8192 * max(steps, 1)
        ms_buf, tag_buf, meta_buf = (C.c_float * cap)(), (C.c_int32 * cap)(), (C.c_int32 * (8 * cap))()
        n = self.lib.moegan_debug_profile_read_ex(ms_buf, tag_buf, meta_buf, cap)
        self.lib.moegan_debug_profile_enable(0)
        self.lib.moegan_debug_set_side_streams(-1)
        return [(ms_buf[i], tag_buf[i], tuple(meta_buf[8 * i: 8 * i + 8])) for i in range(n)]


# --------------------------------------------------------------------------------------------------
# per-kernel roofline table: algorithmic bytes / flops of every launcher of the library, from the
# shape metadata the library attaches to its profiling records (include/moegan_sm100.h, MOEGAN_PROF_*)
# --------------------------------------------------------------------------------------------------
PROF_BASE = 2000000
PROF_NAMES = {1: "sample_kl(+kl_finalize)", 2: "text_fwd", 3: "router_fwd", 4: "plan", 5: "gather", 6: "to_operand", 7: "prep_weights",
              8: "combine_fwd", 9: "combine_bwd", 10: "slab_sum", 11: "uncombine", 12: "router_bwd_logits(+dtl)", 13: "router_bwd_small",
              14: "router_bwd_text", 15: "reparam_kl_bwd", 16: "sum_partials", 17: "colsum_grouped", 18: "balance_loss",
              20: "layernorm_fwd", 21: "layernorm_bwd", 22: "channel_scale_fwd", 23: "channel_scale_bwd", 24: "offset_sample_fwd",
              25: "offset_sample_bwd", 26: "colsum", 27: "upsample2x_fwd", 28: "upsample2x_bwd", 29: "demod_coef_fwd", 30: "demod_coef_bwd"}
GEMM_KINDS = {0: "rows x W^T (K-major B)", 1: "rows x W (MN-major B)", 2: "wgrad A^T B"}
GEMM_EPI = {0: "f32", 1: "bf16", 2: "bias+bf16", 3: "bias+f32", 4: "bias+GELU", 5: "bias+GELU fp32x3", 6: "xGELU'", 7: "xGELU' fp32x3"}
NOMINAL_HBM_GBS = 8000.0     # the figure north_star quotes


def fp32_fma_peak_tflops(pk):
    """148 SMs x 128 fp32 lanes x 2 flop x max SM clock: the CUDA-core bound of the fp32 router projection (nominal, not measured)."""
    return 148 * 128 * 2 * float(pk.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12


def kernel_model(tag, meta, Dt=512):
    """(name, shape label, algorithmic bytes, algorithmic flops, bound) of one profiling record; bytes / flops None = latency-bound."""
    m = list(meta)
    H = 128
    if tag < PROF_BASE:                                          # tcgen05 GEMM: meta[4..7] = N, K | dimA, host rows | -1, group size
        kind, BN, epi, cta2 = (tag % 10000) // 1000, tag % 1000, (tag // 10000) % 10, tag // 100000
        T, d, E, fl = m[0], m[1], m[2], m[3]
        k, dense = fl & 0xff, bool(fl & 0x100)
        N, K, rows, grp, two_out = m[4], m[5], m[6], m[7] & 0xffff, bool(m[7] & 0x10000)
        R = rows if rows >= 0 else (T * E if dense else T * k)   # grouped GEMMs: the dispatched rows (padding rows are not work)
        flops = 2.0 * R * N * K
        S = 2 if fl & 0x400 else 1
        if kind == 3:
            # fused FFN pair (csrc/ffn_fused.cu): Linear(K -> N) + GELU + Linear(N -> K) per row, H stays on chip.  Algorithmic bytes: the
            # token rows once (dense-soft: shared by all experts), the expert weights once, the output rows once
            dense_soft = bool(m[7] & 0x20000)
            tok = R // max(grp, 1) if dense_soft else R
            byt = tok * K * 2 + max(grp, 1) * 2 * N * K * 2 + tok * K * (4 if dense_soft else 2)
            name = f"ffn_fused<{'dense-soft +combine' if dense_soft else 'grouped'}> Linear+GELU+Linear"
            return name, f"R={R} hidden={N} d={K} experts={grp}", float(byt), 2.0 * flops, "tensor"
        # operand bytes: A rows x K (or rows x dimA for wgrad) + B (weights, or rows x N for wgrad) + the output
        if kind == 2:
            byt = R * K * 2 * S + R * N * 2 * S + max(grp, 1) * K * N * 4
        else:
            out_b = {0: 4, 1: 2, 2: 2, 3: 4, 4: 2 * (2 if two_out else 1), 5: 8, 6: 2 + 2, 7: 4 + 4}.get(epi, 4)   # outputs (+ the aux tile read)
            byt = R * K * 2 * S + max(grp, 1) * N * K * 2 * S + R * N * out_b + (R // 32 * N * 4 if epi in (6, 7) else 0)
        name = f"umma_gemm<{BN},{kind},{'pair' if cta2 else '1cta'},{epi}> {GEMM_EPI.get(epi, epi)}{' +GELU\' out' if (epi == 4 and two_out) else ''}"
        return name, f"R={R} N={N} K={K} groups={grp}", float(byt), flops, "tensor"
    kid = tag - PROF_BASE
    name = PROF_NAMES.get(kid, f"kernel{kid}")
    if kid >= 20 or kid in (6, 10, 16, 17, 18):                  # layers around the block and helpers with their own metadata
        a, b, c, bits = m[0], m[1], m[2], m[3]
        sz = lambda bit: 2 if bits & bit else 4
        if kid == 20: return name, f"rows={a} d={b}", float(a * b * (sz(1) + sz(2)) + a * 8), None, "hbm"
        if kid == 21: return name, f"rows={a} d={b}", float(a * b * (sz(1) + 2 * sz(2)) + a * 8), None, "hbm"
        if kid == 22: return name, f"B={a} P={b} C={c}", float(a * b * c * 2 * sz(1) + a * c * sz(2)), None, "hbm"
        if kid == 23: return name, f"B={a} P={b} C={c}", float(a * b * c * 3 * sz(1) + 2 * a * c * sz(2)), None, "hbm"
        if kid == 24: return name, f"B={a} P={b} C={c}", float(a * b * c * 2 * sz(1) + a * b * 8), None, "hbm"
        if kid == 25: return name, f"B={a} P={b} C={c}", float(a * b * c * 3 * sz(1) + 2 * a * b * 8), None, "hbm"
        if kid == 26: return name, f"rows={a} cols={b}", float(a * b * sz(1)), None, "hbm"
        if kid in (27, 28): return name, f"B={a} P={b} C={c}", float(5 * a * b * c * sz(1)), None, "hbm"
        if kid in (29, 30):                                      # a = B, b = Cin, c = Cout, taps in bits 8..: weights (+ dw) dominate
            taps = bits >> 8
            return name, f"B={a} Cin={b} Cout={c} taps={taps}", float(b * c * taps * 4 * (1 if kid == 29 else 2) + b * c * 4 * 2 + a * (b + c) * 4), None, "latency"
        if kid == 6:  return name, f"rows={a} cols={b}", float(a * b * (4 + (4 if bits & 0x400 else 2))), None, "hbm"
        if kid == 10: return name, f"m_cap={a} cols={b}", float(a // 32 * b * 4 + c * b * 4), None, "hbm"
        if kid == 16: return name, f"n={a} splits={b}", float(a * (b + 1) * 4), None, "hbm"
        if kid == 17: return name, f"m_cap={a} cols={b}", float(a * b * 2), None, "hbm"
        return name, "", None, None, "latency"
    T, d, E, fl = m[0], m[1], m[2], m[3]
    k, dense, fp32 = fl & 0xff, bool(fl & 0x100), bool(fl & 0x400)
    R = T * E if dense else T * k
    kk = E if dense else k
    S = 2 if fp32 else 1
    act = 4 if fp32 else 2                                       # y / dx rows
    nrt = H * (d + Dt) + 2 * H * E                               # router weight elements
    label = f"T={T} d={d} E={E} k={kk}{' dense' if dense else ''}"
    if kid == 1:  return name, label, float(5 * nrt * 4), None, "hbm"
    if kid == 3:
        byt = T * d * 4 + H * d * 4 + T * (2 * E + 2 * kk) * 4 + (T * H * 4 if fl & 0x1000 else 0)
        return name, label, float(byt), 2.0 * T * d * H + 2.0 * T * H * E, "fp32"
    if kid == 5:  return name, label, float(T * d * (2 if fl & 0x2000 else 4) + T * k * 8 + R * d * 2 * S + (T * d * 2 * S if fl & 0x1000 else 0) + R * 8 + T * k * 4), None, "hbm"
    if kid == 7:  return name, label, float(E * 8 * d * d * (4 + 2 * S) + 2 * E * 5 * d * 4), None, "hbm"
    if kid == 8:
        res = 4 if fl & 0x1000 else (2 if fl & 0x2000 else 0)
        return name, label, float(R * d * act + T * kk * 4 * (1 if dense else 2) + T * d * res + T * d * (2 if fl & 0x4000 else 4)), None, "hbm"
    if kid == 9:  return name, label, float(T * d * (2 if fl & 0x1000 else 4) + R * d * act + R * d * 2 * S + R * 8 + R // 32 * d * 4), None, "hbm"
    if kid == 11: return name, label, float(R * d * act + (T * d * 4 if fl & 0x1000 else 0) + T * d * 4), None, "hbm"
    if kid == 12: return name, label, float(T * (H * 4 + H * 2 * S + 3 * E * 4) + R * 4), None, "hbm"
    if kid == 15: return name, label, float(7 * nrt * 4), None, "hbm"
    return name, label, None, None, "latency"                    # text_fwd, plan, router_bwd_small / text: a few KB, latency-bound


def kernel_table(records, steps, pk):
    """records: [(ms, tag, meta8)] of `steps` eager steps -> rows sorted by time per step.  GB/s against the measured copy
    bandwidth (MEASURED_PEAKS.json) and against the 8 TB/s nominal; TFLOP/s against the measured bf16 peak (tensor) or the
    nominal fp32 FMA peak (router projection).  Event-bracketed durations include the ~1-2 us gap between queued launches."""
    groups = {}
    for ms, tag, meta in records:
        groups.setdefault((tag, tuple(meta)), []).append(ms)
    hbm = float(pk.get("hbm_gbs", 6453.7))
    tens = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1590.0)))
    f32 = fp32_fma_peak_tflops(pk)
    rows = []
    for (tag, meta), v in groups.items():
        name, label, byt, flops, bound = kernel_model(tag, meta)
        us = 1e3 * sum(v) / len(v)
        row = {"kernel": name, "shape": label, "launches_per_step": round(len(v) / max(steps, 1), 2), "us": round(us, 2),
               "us_per_step": round(1e3 * sum(v) / max(steps, 1), 1), "bound": bound}
        if byt is not None and us > 0:
            gbs = byt / (us * 1e-6) / 1e9
            row.update({"bytes": int(byt), "GBps": round(gbs, 1), "frac_hbm_measured": round(gbs / hbm, 3), "frac_hbm_8TBps": round(gbs / NOMINAL_HBM_GBS, 3)})
        if flops is not None and us > 0:
            tf = flops / (us * 1e-6) / 1e12
            row.update({"flops": flops, "TFLOPps": round(tf, 2), "frac_compute": round(tf / (tens if bound == "tensor" else f32), 3)})
        if bound == "tensor" and byt is not None and flops is not None:
            # which roof is lower for this instance: time at the measured tensor peak vs time at the measured HBM bandwidth
            t_t, t_m = flops / (tens * 1e12), byt / (hbm * 1e9)
            row["roof"] = "hbm" if t_m > t_t else "tensor"
            row["frac_of_roof"] = round(max(t_t, t_m) / (us * 1e-6), 3)
        elif bound == "hbm" and "frac_hbm_measured" in row:
            row["roof"], row["frac_of_roof"] = "hbm", row["frac_hbm_measured"]
        elif bound == "fp32" and "frac_compute" in row:
            row["roof"], row["frac_of_roof"] = "fp32_fma", row["frac_compute"]
        rows.append(row)
    rows.sort(key=lambda r: -r["us_per_step"])
    return rows


def gemm_roof_summary(rows):
    """Time-weighted position of the tcgen05 GEMM instances against the LOWER of their two roofs (per-instance `roof` / `frac_of_roof` of
    kernel_table): sum of roof times / sum of measured times, and the share of GEMM time spent in HBM-bound instances."""
    g = [r for r in rows if r["bound"] == "tensor" and "frac_of_roof" in r]
    tot = sum(r["us_per_step"] for r in g)
    if tot <= 0:
        return None
    return {"frac_of_lower_roof": round(sum(r["us_per_step"] * r["frac_of_roof"] for r in g) / tot, 3),
            "hbm_bound_share_of_gemm_time": round(sum(r["us_per_step"] for r in g if r.get("roof") == "hbm") / tot, 3),
            "note": "per instance: max(flops / measured bf16 peak, algorithmic bytes / measured HBM bandwidth) / event-bracketed duration"}


def write_kernel_table(path, title, rows, append=False):
    if not path:
        return
    with open(path, "a" if append else "w") as f:
        f.write(f"### per-kernel roofline: {title}\n\n")
        f.write("CUDA events recorded by the library around each launch on its stream (eager pass, L2 flushed between steps; ~1-2 us of "
                "launch gap inside every figure).  bytes / flops = ALGORITHMIC (DESIGN.md section 4); GB/s against the measured copy bandwidth "
                "of MEASURED_PEAKS.json and the 8 TB/s nominal; `frac of roof` = time at the lower of the two roofs / measured time.\n\n")
        f.write("| kernel | shape | n/step | us | us/step | MB | GB/s | of measured HBM | of 8 TB/s | TFLOP/s | roof | frac of roof |\n")
        f.write("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|\n")
        g = lambda r, k, fmt="{}": fmt.format(r[k]) if k in r else ""
        for r in rows:
            f.write(f"| `{r['kernel']}` | {r['shape']} | {r['launches_per_step']} | {r['us']} | {r['us_per_step']} | "
                    f"{(r['bytes'] / 1e6 if 'bytes' in r else 0):.1f} | {g(r, 'GBps')} | {g(r, 'frac_hbm_measured')} | {g(r, 'frac_hbm_8TBps')} | "
                    f"{g(r, 'TFLOPps')} | {r.get('roof', r['bound'])} | {g(r, 'frac_of_roof')} |\n")
        f.write("\n")


def traffic_from_profiles(key):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(key)
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------
# workload: MoE block microbench (configs[1])
# --------------------------------------------------------------------------------------------------
def bench_moe_block(args, T_, rank, world, dev, steps, warmup, use_graph, WL=None, light=False):
    """light: block time + e2e only (the companion lines of SURVEY 8(d) config 2), no per-kernel profiling passes."""
    import torch
    WL = WL or globals()["WL"]
    pkg = importlib.import_module(PKG)
    lib = pkg._cabi.lib()
    params, x_np, w_np, dout_np, eps_np = make_numpy_case(1000 + rank, WL)      # rank-offset data ...
    params, _, _, _, _ = make_numpy_case(0, WL)                                  # ... identical weights on every rank
    moe = pkg.SparseMoE(WL["d"], WL["Dt"], WL["E"], top_k=WL["k"], precision="bf16")
    sd = {k: torch.from_numpy(v) for k, v in params.items()}
    for k in ("epsilon_f", "epsilon_t", "epsilon_c"):
        sd["router." + k] = moe.state_dict()["router." + k]
    moe.load_state_dict(sd)
    moe.to(dev).train()
    plist = list(moe.parameters())
    T = WL["B"] * WL["H"] * WL["W"]
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)       # NHWC memory, as norm3 produces (:561)
    x_host = torch.from_numpy(x_np).pin_memory()
    w_host = torch.from_numpy(w_np).pin_memory()
    x_dev = nhwc(x_host.to(dev)).requires_grad_(True)
    w_dev = w_host.to(dev).requires_grad_(True)
    dout = nhwc(torch.from_numpy(dout_np).to(dev))
    eps = tuple(torch.from_numpy(eps_np[k]).to(dev) for k in ("epsilon_f", "epsilon_t", "epsilon_c"))
    red = pkg.GradReducer(plist)
    kl_seed = torch.full((), 1e-3, device=dev)

    def step_device():
        for p in plist:
            p.grad = None
        x_dev.grad = None
        w_dev.grad = None
        out, kl, probs = moe(x_dev, w_dev, WL["af"], eps=eps)
        # fwd + bwd with a given upstream gradient ("upstream grad = randn_like(out)") and d loss / d kl = 1e-3
        red.begin(); torch.autograd.backward([out, kl], [dout, kl_seed]); red.finish()
        return out

    def step_e2e():
        for p in plist:
            p.grad = None
        xd = nhwc(x_host.to(dev, non_blocking=True)).requires_grad_(True)
        wd = w_host.to(dev, non_blocking=True).requires_grad_(True)
        out, kl, probs = moe(xd, wd, WL["af"])                    # public call: eps redrawn on the device (:349-351)
        loss = (out * dout).sum() + 1e-3 * kl
        red.begin(); loss.backward(); red.finish()
        return float(loss.item())                                 # D2H read of the step's result

    launch_mode, step_value, launches_per_graph = "eager", step_device, None
    if use_graph:
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
            with torch.cuda.graph(graph, stream=side):        # the stream the leaves' AccumulateGrad nodes were created on
                step_device()
            launches_per_graph = int(lib.moegan_debug_launch_count() - n_before)
            step_value, launch_mode = graph.replay, "cuda_graph"
        except Exception as e:  # noqa
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            torch.cuda.synchronize()
    with ClockSampler(dev.index) as clk:
        total_ms, launches, host_ms = T_.run(step_value, steps, warmup)
    if launches_per_graph is not None:
        launches = launches_per_graph * steps
    e2e_ms, _, _ = T_.run(step_e2e, steps, max(2, warmup // 2))
    if light:
        ms_per_step = total_ms / steps
        return {"ms_per_step": ms_per_step, "fwd_bwd_us": ms_per_step * 1e3, "tokens_per_s": T * world / (ms_per_step * 1e-3), "launch": launch_mode,
                "gpu_launches": launches, "tokens": T, "workload": moe_block_workload_name(WL),
                "e2e_tokens_per_s": T * world / (e2e_ms / steps * 1e-3),
                "tensor_floor_us": 48.0 * T * (WL["k"] or WL["E"]) * WL["d"] ** 2 / (float(peaks().get("bf16_tflops", 1590.0)) * 1e12) * 1e6}
    ms_list, _ = T_.gemm_profile(step_device, steps, lead_ms=2.0)
    nk = max(2, min(steps, 10))
    kernels = kernel_table(T_.kernel_profile(step_device, nk, lead_ms=3.0), nk, peaks())
    per_step = len(ms_list) // steps if steps else 0
    names = ["ffn1_fwd", "ffn2_fwd", "dgrad2", "dgrad1", "wgrad2", "wgrad1"]     # the six expert GEMMs lead each step's router GEMMs
    by_kernel = {}
    for s in range(steps):
        for i, nm in enumerate(names[:per_step]):
            by_kernel.setdefault(nm, []).append(ms_list[s * per_step + i])
    all6 = [v for vs in by_kernel.values() for v in vs]
    mean_ms = sum(all6) / max(len(all6), 1)
    flops_per_launch = 48.0 * T * WL["k"] * WL["d"] ** 2 / 6.0                     # 48*T*k*d^2 over six equal GEMMs
    pk = peaks()
    peak_tf = float(pk.get("bf16_tflops", 1590.0))
    ach = flops_per_launch / (mean_ms * 1e-3) / 1e12 if mean_ms > 0 else 0.0
    ms_per_step = total_ms / steps
    return {
        "ms_per_step": ms_per_step, "tokens_per_s": T * world / (ms_per_step * 1e-3), "launch": launch_mode,
        "clocks": clk.summary(), "gpu_launches": launches, "host_enqueue_us_per_step": host_ms / steps * 1e3,
        "e2e": {"value": T * world / (e2e_ms / steps * 1e-3), "unit": "tokens/s", "ms_per_step": e2e_ms / steps,
                "h2d_bytes_per_step": int(x_host.numel() * 4 + w_host.numel() * 4), "d2h_bytes_per_step": 4},
        "roofline": {"bound": "tensor", "kernel": "umma_gemm_kernel (six grouped expert GEMMs per step)", "achieved": ach, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": ach / peak_tf,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: kernels timed individually)" if pk else "fallback 1590 (B200_PROFILING.md)",
                     "traffic": traffic_from_profiles("umma_gemm_moe_block"), "flops_per_launch": flops_per_launch,
                     "mean_launch_us": mean_ms * 1e3, "per_kernel_us": {k: 1e3 * sum(v) / len(v) for k, v in by_kernel.items()},
                     "by_instance": gemm_roof_summary(kernels)},
        "tokens": T, "kernels": kernels,
    }


# --------------------------------------------------------------------------------------------------
# workload: full train step (configs[2] / [3]) and inference (configs[4])
# --------------------------------------------------------------------------------------------------
def build_models(pkg, dev, seed=0):
    import torch
    torch.manual_seed(seed)                       # identical initial weights on every rank
    G = pkg.host_model.AuroraGenerator().to(dev)
    D = pkg.host_model.AuroraDiscriminator().to(dev)
    if os.environ.get("MOEGAN_BENCH_CHANNELS_LAST", "1") == "1":
        # NHWC conv stacks: cuDNN's tensor-core kernels are NHWC-native (no nchw<->nhwc transposes around every conv) and
        # the MoE blocks consume NHWC tokens anyway
        G = G.to(memory_format=torch.channels_last)
        D = D.to(memory_format=torch.channels_last)
    return G, D


def bench_train_step(args, T_, rank, world, dev, steps, warmup, use_graph):
    import torch
    pkg = importlib.import_module(PKG)
    lib = pkg._cabi.lib()
    # autotune the (fixed-shape) convolutions of the host model, incl. R1's double backward.  MOEGAN_BENCH_CUDNN_BENCHMARK=0: cuDNN's
    # heuristic choice instead (identical on every rank and every run; the autotuner's pick varies from process to process)
    torch.backends.cudnn.benchmark = os.environ.get("MOEGAN_BENCH_CUDNN_BENCHMARK", "1") == "1"
    G, D = build_models(pkg, dev)
    B = TRAIN_BATCH
    torch.manual_seed(1234 + rank)                # rank-offset data
    real_host = torch.randn(B, 3, 64, 64).clamp(-1, 1).pin_memory()
    text_host = torch.randn(B, 512).pin_memory()
    real, text = real_host.to(dev), text_host.to(dev)
    if os.environ.get("MOEGAN_BENCH_CHANNELS_LAST", "1") == "1":
        real = real.contiguous(memory_format=torch.channels_last)

    bucket_mb = float(os.environ.get("MOEGAN_BENCH_BUCKET_MB", "25"))

    def make_trainer(graph):
        return pkg.MoEGANTrainer(G, D, amp_dtype=torch.bfloat16, use_cuda_graph=graph, bucket_bytes=int(bucket_mb * (1 << 20)),
                                 exact_dp_balance=os.environ.get("MOEGAN_BENCH_LOCAL_BALANCE", "0") != "1")     # profiling switch only

    launch_mode = "eager"
    trainer = make_trainer(False)
    launches_per_graph = None
    if use_graph:
        try:
            tg = make_trainer(True)
            n_before = lib.moegan_debug_launch_count()
            tg._capture(real, text)
            # launches counted during capture = 3 warm-up steps + the captured one
            launches_per_graph = int(lib.moegan_debug_launch_count() - n_before) // 4
            trainer, launch_mode = tg, "cuda_graph"
        except Exception as e:  # noqa
            if rank == 0:
                print(f"[bench] CUDA graph capture of the train step failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            torch.cuda.synchronize()
            trainer = make_trainer(False)

    with ClockSampler(dev.index) as clk:
        total_ms, launches, host_ms = T_.run(lambda: trainer.step(real, text), steps, warmup)
    if launches_per_graph is not None:
        launches = launches_per_graph * steps
    e2e_ms, _, _ = T_.run(lambda: trainer.step_from_host(real_host, text_host), steps, max(2, warmup // 2))

    # roofline of the tcgen05 grouped GEMM family inside the step (eager pass; events recorded by the library)
    eager = trainer if launch_mode == "eager" else make_trainer(False)
    nprof = max(2, min(steps, 4))
    ms_list, tags = T_.gemm_profile(lambda: eager._step_impl(real, text), nprof, lead_ms=15.0)
    kernels = kernel_table(T_.kernel_profile(lambda: eager._step_impl(real, text), nprof, lead_ms=25.0), nprof, peaks())
    per_step = len(ms_list) // nprof if nprof else 0
    # every expert GEMM of every block has the same algorithmic flops: 2*M*N*K with M = E*T, {N,K} = {4d, d} -> 8*E*T*d^2,
    # and T*d^2 = 16*B*512^2 is the same for the three blocks.  Per step: 3 blocks x (2 forwards x 2 GEMMs +
    # 4 backward GEMMs) = 24 expert launches (+ 12 small router-backward GEMMs, < 2% of the flops, whose time is included).
    E = 4
    flops_per_launch = 8.0 * E * (16 * B) * 512 ** 2
    step_gemm_ms = sum(ms_list) / nprof if nprof else 0.0
    expert_launches = 24
    pk = peaks()
    peak_tf = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1590.0)))
    ach = expert_launches * flops_per_launch / (step_gemm_ms * 1e-3) / 1e12 if step_gemm_ms > 0 else 0.0
    ms_per_step = total_ms / steps
    res = {
        "ms_per_step": ms_per_step, "imgs_per_s": B * world / (ms_per_step * 1e-3), "launch": launch_mode, "clocks": clk.summary(),
        "gpu_launches": launches, "host_enqueue_us_per_step": host_ms / steps * 1e3,
        "e2e": {"value": B * world / (e2e_ms / steps * 1e-3), "unit": "imgs/s", "ms_per_step": e2e_ms / steps,
                "h2d_bytes_per_step": int(real_host.numel() * 4 + text_host.numel() * 4), "d2h_bytes_per_step": 8},
        "roofline": {"bound": "tensor", "kernel": "umma_gemm_kernel (24 grouped expert GEMMs + 12 router-backward GEMMs per train step)",
                     "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernels inside a long step)" if pk else "fallback 1590 (B200_PROFILING.md)",
                     "traffic": traffic_from_profiles("umma_gemm_train_step"), "flops_per_launch": flops_per_launch,
                     "gemm_ms_per_step": step_gemm_ms, "gemm_launches_per_step": per_step,
                     "gemm_share_of_step": step_gemm_ms / ms_per_step if ms_per_step else None,
                     "by_instance": gemm_roof_summary(kernels)},
        "kernels": kernels,
    }
    # ---- inference (configs[4]): generate_images semantics, eval-mode hard top-1 routing ----
    try:
        zi = torch.randn(INFER_BATCH, 512, device=dev)
        ti = torch.randn(INFER_BATCH, 512, device=dev)
        nrep = max(3, steps // 2)

        def infer():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return pkg.host_model.sample_images(G, zi, ti, truncation_psi=0.7)
        inf_ms, _, _ = T_.run(infer, nrep, 3)
        res["inference"] = {"imgs_per_s": INFER_BATCH * world / (inf_ms / nrep * 1e-3), "batch": INFER_BATCH,
                            "ms_per_batch": inf_ms / nrep, "routing": "hard top-1 (eval)", "dtype": "bf16"}
    except Exception as e:  # noqa
        res["inference"] = {"error": f"{type(e).__name__}: {e}"}
    return res


# --------------------------------------------------------------------------------------------------
# the UNMODIFIED reference on the SAME B200 (SURVEY 8(d) "reference on GPU" companion rows): PyTorch-CUDA eager, stock code path
# --------------------------------------------------------------------------------------------------
# companion microbench shapes of SURVEY 8(d) config 2: the reference-native E=4, k=E blocks at batch 256, and configs[1] with mu x30
NATIVE_WLS = [dict(tag="reference-native block 4x4 @ batch 256", B=256, H=4, W=4, d=512, Dt=512, E=4, k=None, mu_scale=8.0, af=1.0),
              dict(tag="reference-native block 8x8 @ batch 256", B=256, H=8, W=8, d=256, Dt=512, E=4, k=None, mu_scale=8.0, af=1.0),
              dict(tag="reference-native block 16x16 @ batch 256", B=256, H=16, W=16, d=128, Dt=512, E=4, k=None, mu_scale=8.0, af=1.0)]
WL_MU30 = dict(WL, mu_scale=30.0, tag="BASELINE configs[1], non-uniform routing")


def reference_gpu_train_rows(dev, batch, steps, warmup=2):
    """Full train step of the unmodified reference modules (oracle/_ref) on this GPU, fp32 and bf16 autocast, batch 256:
    the number the B200 arm's `value` should be read against (same device, same batch, same losses)."""
    import torch
    from oracle.ref_step import ReferenceTrainStep
    rows = {}
    torch.manual_seed(4321)
    real = torch.randn(batch, 3, 64, 64, device=dev).clamp(-1, 1)
    text = torch.randn(batch, 512, device=dev)
    for name, dt in (("bf16_autocast", torch.bfloat16), ("fp32", None)):
        try:
            tr = ReferenceTrainStep(dev, amp_dtype=dt, checkpoint_activation=False, seed=0)
            for _ in range(warmup):
                tr.step(real, text)
            torch.cuda.synchronize()
            with ClockSampler(dev.index) as clk:
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                h0 = time.perf_counter()
                a.record()
                for _ in range(steps):
                    out = tr.step(real, text)
                b.record()
                torch.cuda.synchronize()
                wall = (time.perf_counter() - h0) * 1e3 / steps
            ms = a.elapsed_time(b) / steps
            rows[name] = {"imgs_per_s": batch / (ms * 1e-3), "ms_per_step": ms, "host_wall_ms_per_step": wall, "batch": batch, "steps": steps,
                          "clocks": clk.summary(), "peak_mem_GB": round(torch.cuda.max_memory_allocated(dev) / 1e9, 1),
                          "losses": {k: float(v) for k, v in out.items()}}
            del tr, out
        except Exception as e:  # noqa
            rows[name] = {"error": f"{type(e).__name__}: {e}"}
        import gc
        gc.collect(); torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats(dev)
    rows["what"] = ("unmodified reference (oracle/_ref/t2i_moe_gan.py) AuroraGenerator/AuroraDiscriminator, PyTorch-CUDA eager on this same GPU, one iteration of "
                    "its batch loop (oracle/ref_step.py = t2i_moe_gan.py:1262-1421 without the CLIP terms), checkpointing off, CUDA events around "
                    f"{steps} steps after {warmup} warm-ups; the reference's host syncs (.item()) are inside")
    return rows


def reference_gpu_moe_block_us(dev, wl, reps=10, warmup=3):
    """fwd+bwd of the unmodified reference SparseMoE (dense-soft, E=4) on this GPU at one reference-native shape; fp32 and bf16 autocast."""
    import torch
    from oracle import ref_loader
    ref = ref_loader.load_reference()
    torch.manual_seed(0)
    moe = ref.SparseMoE(wl["d"], wl["Dt"], num_experts=wl["E"]).to(dev).train()
    with torch.no_grad():
        for n in ("feature_mu", "text_mu", "combined_mu"):
            getattr(moe.router, n).mul_(wl["mu_scale"])
    x = torch.randn(wl["B"], wl["H"], wl["W"], wl["d"], device=dev).permute(0, 3, 1, 2).requires_grad_(True)
    w = torch.randn(wl["B"], wl["Dt"], device=dev, requires_grad=True)
    dout = torch.randn(wl["B"], wl["H"], wl["W"], wl["d"], device=dev).permute(0, 3, 1, 2)
    res = {}
    for name, dt in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
        def step():
            moe.zero_grad(set_to_none=True); x.grad = None; w.grad = None
            with torch.autocast("cuda", dtype=dt, enabled=dt is not None):
                out, kl, probs = moe(x, w, wl["af"])
            torch.autograd.backward([out, kl], [dout.to(out.dtype), torch.full((), 1e-3, device=dev)])
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            step()
        b.record(); torch.cuda.synchronize()
        res[name + "_us"] = a.elapsed_time(b) / reps * 1e3
    return res


def bench_strong_scaling(args, rank, world, dev, T_, steps):
    """configs[3] as written: GLOBAL batch 2048 split over the ranks (1024 / 512 / 256 per GPU at N = 2 / 4 / 8)."""
    import torch
    pkg = importlib.import_module(PKG)
    per = GLOBAL_BATCH_DP // world
    G, D = build_models(pkg, dev)
    torch.manual_seed(99 + rank)
    real = torch.randn(per, 3, 64, 64, device=dev).clamp(-1, 1).contiguous(memory_format=torch.channels_last)
    text = torch.randn(per, 512, device=dev)
    tr = pkg.MoEGANTrainer(G, D, amp_dtype=torch.bfloat16, use_cuda_graph=not args.no_graph)
    total_ms, _, _ = T_.run(lambda: tr.step(real, text), steps, 3)
    ms = total_ms / steps
    del tr, G, D
    torch.cuda.empty_cache()
    return {"global_batch": GLOBAL_BATCH_DP, "batch_per_gpu": per, "ms_per_step": ms, "imgs_per_s": GLOBAL_BATCH_DP / (ms * 1e-3), "scaling": "strong",
            "workload": f"BASELINE configs[3] as written: global batch {GLOBAL_BATCH_DP} sharded over {world} GPUs, NCCL gradient all-reduce"}


# --------------------------------------------------------------------------------------------------
def run_b200_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    lib = pkg._cabi.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    T_ = Timer(torch, dist, world, dev, lib)
    warmup = max(args.warmup, 3)
    use_graph = not args.no_graph

    if args.workload == "moe_block":
        r = bench_moe_block(args, T_, rank, world, dev, args.steps, warmup, use_graph)
        if rank != 0:
            return
        sec, units = cpu_moe_block_time(5)
        line = {
            "metric": "moe_block_fwd_bwd_tokens_per_s", "value": r["tokens_per_s"], "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": moe_block_workload_name(), "tokens_per_gpu": r["tokens"], "l2": "256 MiB buffer zeroed between timed iterations",
                       "parallelism": f"dp{world}" if world > 1 else "single", "launch": r["launch"], "router": "fp32",
                       "experts": "bf16 in / fp32 acc (tcgen05)"},
            "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "roofline": r["roofline"],
            "cpu_baseline": {"value": units / sec, "unit": "tokens/s", "cores": cpu_threads(), "kind": "port",
                             "sample": f"full step (T={units}) x 5 reps, median {sec * 1e3:.0f} ms; numpy fp32 oracle port"},
            "moe_block_fwd_bwd_us": r["ms_per_step"] * 1e3, "host_enqueue_us_per_step": r["host_enqueue_us_per_step"],
            "kernels": r["kernels"],
        }
        write_kernel_table(args.kernels_out, line["config"]["workload"], r["kernels"])
        print(json.dumps(line), flush=True)
        return

    r = bench_train_step(args, T_, rank, world, dev, args.steps, warmup, use_graph)
    mb = None
    if not args.skip_extras:
        try:
            mb = bench_moe_block(args, T_, rank, world, dev, max(10, args.steps), 5, use_graph)
        except Exception as e:  # noqa
            mb = {"error": f"{type(e).__name__}: {e}"}
    companions, strong = None, None
    if not args.skip_extras:
        # SURVEY 8(d) config 2 companion lines: configs[1] with non-uniform routing (mu x30) and the reference-native E=4, k=E blocks
        companions = []
        for wl in [WL_MU30] + NATIVE_WLS:
            try:
                c = bench_moe_block(args, T_, rank, world, dev, 10, 3, use_graph, WL=wl, light=True)
                if world == 1 and wl.get("k") is None and reference_available():
                    c["reference_gpu_eager"] = reference_gpu_moe_block_us(dev, wl)      # the unmodified SparseMoE on this GPU, same shape
                    c["speedup_vs_reference_gpu_bf16"] = c["reference_gpu_eager"]["bf16_autocast_us"] / c["fwd_bwd_us"]
                companions.append(c)
            except Exception as e:  # noqa
                companions.append({"workload": moe_block_workload_name(wl), "error": f"{type(e).__name__}: {e}"})
            torch.cuda.empty_cache()
        if world in (2, 4):
            try:
                strong = bench_strong_scaling(args, rank, world, dev, T_, max(5, args.steps // 2))
            except Exception as e:  # noqa
                strong = {"error": f"{type(e).__name__}: {e}"}
    if rank != 0:
        return
    ref_gpu = None
    if world == 1 and not args.skip_extras and reference_available():
        torch.cuda.empty_cache()
        ref_gpu = reference_gpu_train_rows(dev, TRAIN_BATCH, steps=3, warmup=2)
    cpu = None
    if not args.skip_cpu:
        sec, units, info = cpu_train_step_time(3, warmup=1)
        cpu = cpu_baseline_record(sec, units, info, 3)
    line = {
        "metric": "train_imgs_per_s", "value": r["imgs_per_s"], "unit": "imgs/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": train_workload_name(world), "batch_per_gpu": TRAIN_BATCH, "global_batch": TRAIN_BATCH * world,
                   "image": "real 64x64 -> D; G emits 16x16 (+8x8)", "l2": "256 MiB buffer zeroed between timed iterations",
                   "parallelism": f"dp{world}" if world > 1 else "single", "launch": r["launch"],
                   "allreduce": f"NCCL, {os.environ.get('MOEGAN_BENCH_BUCKET_MB', '25')} MB buckets from grad-ready hooks on a comm stream" if world > 1 else None},
        "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "roofline": r["roofline"], "cpu_baseline": cpu,
        "host_enqueue_us_per_step": r["host_enqueue_us_per_step"], "inference": r.get("inference"),
        # per-kernel roofline of the library's launches inside the step (eager pass, library-recorded CUDA events): the 24 largest
        "kernels": r["kernels"][:24],
        "library_us_per_step": round(sum(k["us_per_step"] for k in r["kernels"]), 1),
    }
    write_kernel_table(args.kernels_out, line["config"]["workload"], r["kernels"])
    if ref_gpu is not None:
        line["reference_gpu_eager"] = ref_gpu
        bf = ref_gpu.get("bf16_autocast", {})
        if "imgs_per_s" in bf:
            line["speedup_vs_reference_gpu_eager_bf16"] = r["e2e"]["value"] / bf["imgs_per_s"]
    if companions is not None:
        line["moe_block_companions"] = companions
    if strong is not None:
        line["config3_strong_scaling"] = strong
    line["parity"] = PARITY_NOTE
    if mb is not None and "error" not in mb:
        line["moe_block"] = {"fwd_bwd_us": mb["ms_per_step"] * 1e3, "tokens_per_s": mb["tokens_per_s"], "launch": mb["launch"],
                             "workload": moe_block_workload_name(), "roofline": mb["roofline"], "e2e": mb["e2e"], "kernels": mb["kernels"]}
        write_kernel_table(args.kernels_out, moe_block_workload_name(), mb["kernels"], append=True)
    elif mb is not None:
        line["moe_block"] = mb
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train_step", choices=["train_step", "moe_block"])
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--skip-extras", action="store_true", help="skip the embedded moe_block microbench")
    ap.add_argument("--kernels-out", default=None, help="also write the per-kernel roofline table(s) as markdown to this file")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world > 1:
        # NCCL writes its version / warning lines to stdout when NCCL_DEBUG is set in the environment: keep stdout for the JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
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
