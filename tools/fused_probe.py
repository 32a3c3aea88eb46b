#!/usr/bin/env python
"""Forward-only MoE block call (the D-step's generator pass under no_grad / inference) with the fused FFN kernel on and off:
whole-call time (CUDA events, L2 flushed) and the per-kernel breakdown recorded by the library (profile level 2).

    python tools/fused_probe.py [b8 b16 micro8 eval8 eval16]
"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pkg = importlib.import_module("moe-gan_cpsc541_b200")
C = pkg._cabi; lib = C.lib()
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
# name: (B, H, W, d, E, top_k, eval)
SHAPES = {"b8": (256, 8, 8, 256, 4, None, False), "b16": (256, 16, 16, 128, 4, None, False), "micro8": (64, 8, 8, 256, 8, 2, False),
          "eval8": (1024, 8, 8, 256, 4, None, True), "eval16": (1024, 16, 16, 128, 4, None, True)}
PEAK = float(bench.peaks().get("bf16_tflops", 1643.8))


def probe(name):
    B, H, W, d, E, k, ev = SHAPES[name]
    torch.manual_seed(0)
    moe = pkg.SparseMoE(d, 512, E, top_k=k, precision="bf16").to(dev)
    moe.train(not ev)
    moe.redraw_eps = False                # same eps for the fused and the separate run
    for n in ("epsilon_f", "epsilon_t", "epsilon_c"):
        getattr(moe.router, n).normal_()
    with torch.no_grad():
        for n in ("feature_mu", "text_mu", "combined_mu"):
            getattr(moe.router, n).mul_(8.0)
    x = torch.randn(B, H, W, d, device=dev).permute(0, 3, 1, 2); w = torch.randn(B, 512, device=dev)
    res = torch.randn(B, H, W, d, device=dev).bfloat16().permute(0, 3, 1, 2)
    T = B * H * W
    kk = 1 if ev else (E if k is None else k)
    flops = 16.0 * T * kk * d * d

    def call():
        with torch.no_grad():
            return moe(x, w, 1.0, residual=res)[0]
    out = {}
    for fused in (True, False):
        pkg.functional.set_fused_ffn(fused)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_(); torch.cuda._sleep(2000000)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); o = call(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        out[fused] = o.float()
        lib.moegan_debug_profile_enable(2); lib.moegan_debug_set_side_streams(0)
        n = 5
        for _ in range(n):
            flush.zero_(); torch.cuda._sleep(4000000); call()
        torch.cuda.synchronize()
        cap = 4096
        ms, tg, meta = (C.c_float * cap)(), (C.c_int32 * cap)(), (C.c_int32 * (8 * cap))()
        cnt = lib.moegan_debug_profile_read_ex(ms, tg, meta, cap)
        lib.moegan_debug_profile_enable(0); lib.moegan_debug_set_side_streams(-1)
        rows = bench.kernel_table([(ms[i], tg[i], tuple(meta[8 * i:8 * i + 8])) for i in range(cnt)], n, bench.peaks())
        ffn_us = sum(r["us_per_step"] for r in rows if r["bound"] == "tensor" or r["kernel"].startswith("combine_fwd"))
        print(f"\n## {name}: T={T} d={d} E={E} k={kk} {'eval' if ev else 'train-mode no_grad'}  fused={fused}: call {ts[len(ts) // 2]:.1f} us (median of 10, events); "
              f"FFN(+combine) kernels {ffn_us:.1f} us = {flops / (ffn_us * 1e-6) / 1e12:.0f} TFLOP/s = {flops / (ffn_us * 1e-6) / 1e12 / PEAK:.3f} of the bf16 peak")
        for r in rows:
            print(f"   {r['kernel']:<62} {r['shape']:<34} n={r['launches_per_step']:<4} {r['us']:8.1f} us  roof={r.get('roof', r['bound'])} frac={r.get('frac_of_roof', '')}")
    pkg.functional.set_fused_ffn(None)
    d_ = (out[True] - out[False]).abs().max() / out[False].abs().max()
    print(f"   fused vs separate kernels: max rel diff {float(d_):.2e}")


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(SHAPES)):
        probe(nm)
