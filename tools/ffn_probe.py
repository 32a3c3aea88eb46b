#!/usr/bin/env python
"""Per-GEMM timings of moegan_ffn_fwd / moegan_ffn_bwd (library CUDA events around every tcgen05 launch) on the
configs[1] microbench shape and on the three reference-native block shapes of the batch-256 train step.

    python tools/ffn_probe.py [micro b4 b8 b16]
"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("moe-gan_cpsc541_b200")
C = pkg._cabi; lib = C.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
dev = "cuda"
SHAPES = {"micro": (4096, 512, 8, 2, False), "b4": (4096, 512, 4, 4, True), "b8": (16384, 256, 4, 4, True), "b16": (65536, 128, 4, 4, True)}
PEAK = 1643.8


def read_prof():
    torch.cuda.synchronize()
    ms = (C.c_float * 8192)(); tg = (C.c_int32 * 8192)()
    cnt = lib.moegan_debug_profile_read(ms, tg, 8192)
    return [ms[i] * 1e3 for i in range(cnt)]


def probe(name):
    T, d, E, k, dense = SHAPES[name]
    shp = C.make_shape(T, d, 512, E, k, 8, T // 8, 0, dense, False)
    m_cap = shp.m_cap; hid = 4 * d
    plan = torch.zeros(int(lib.moegan_plan_ints(m_cap)), dtype=torch.int32)
    rows = T if dense else T * k // E
    off = 0
    for e in range(E):
        plan[8 + e] = off; plan[48 + e] = rows
        seg = (rows + 255) // 256 * 256
        for b in range(off // 128, (off + seg) // 128): plan[88 + b] = e
        off += seg
    plan[8 + E] = off; plan[0] = off // 128; plan[1] = off
    if dense: plan[2] = (T + 255) // 256 * 256
    plan = plan.to(dev)
    a_rows = T if dense else m_cap
    a_op = torch.randn(a_rows, d, device=dev).bfloat16()
    W1 = (torch.randn(E * hid, d, device=dev) * 0.05).bfloat16(); W2 = (torch.randn(E * d, hid, device=dev) * 0.05).bfloat16()
    b1 = torch.randn(E, hid, device=dev) * 0.1; b2 = torch.randn(E, d, device=dev) * 0.1
    pre = torch.empty(m_cap, hid, device=dev, dtype=torch.bfloat16); h = torch.empty(m_cap, hid, device=dev, dtype=torch.bfloat16)
    y = torch.empty(m_cap, d, device=dev, dtype=torch.bfloat16)
    dy = (torch.randn(m_cap, d, device=dev) * 0.1).bfloat16()
    dpre = torch.empty(m_cap, hid, device=dev, dtype=torch.bfloat16); dxr = torch.empty(m_cap, d, device=dev, dtype=torch.bfloat16)
    dW1 = torch.empty(E, hid, d, device=dev); dW2 = torch.empty(E, d, hid, device=dev)
    db1 = torch.empty(E, hid, device=dev); db2 = torch.empty(E, d, device=dev)
    sb = int(lib.moegan_ffn_bwd_scratch_bytes(shp)); scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    flops = 2.0 * (plan[1].item() if not dense else E * T) * hid * d
    n = 10

    def fwd(with_g):
        C.check(lib.moegan_ffn_fwd(shp, C.ptr(plan), C.ptr(a_op), C.ptr(W1), C.ptr(b1), C.ptr(W2), C.ptr(b2), C.ptr(pre) if with_g else None,
                                   C.ptr(h), C.ptr(y), st), "ffn_fwd")

    def bwd():
        C.check(lib.moegan_ffn_bwd(shp, C.ptr(plan), C.ptr(a_op), C.ptr(W1), C.ptr(W2), C.ptr(pre), C.ptr(h), C.ptr(dy), None, C.ptr(dpre), C.ptr(dxr),
                                   C.ptr(dW1), C.ptr(db1), C.ptr(dW2), C.ptr(db2), C.ptr(scratch), sb, st), "ffn_bwd")

    def run(fn, per):
        for _ in range(3): fn()
        read_prof()
        for _ in range(n):
            flush.zero_(); fn()
        v = read_prof()
        assert len(v) == per * n, (len(v), per, n)
        return [sum(v[i::per]) / n for i in range(per)]

    if os.environ.get("FFN_PROBE_ONCE") == "1":      # one launch of each of the six GEMMs (for ncu)
        fwd(True); bwd(); torch.cuda.synchronize(); return
    lib.moegan_debug_profile_enable(1)
    f_g = run(lambda: fwd(True), 2); f_n = run(lambda: fwd(False), 2); b = run(bwd, 4)
    lib.moegan_debug_profile_enable(0)
    # whole-call times (no per-launch events)
    def whole(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize(); tot = 0.0
        for _ in range(n):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
        return tot / n * 1e3
    wf = whole(lambda: fwd(True)); wb = whole(bwd)
    tf = lambda us: flops / us / 1e6
    print(f"[{name}] T={T} d={d} E={E} k={k} dense={dense} rows={int(plan[1])} GF/GEMM={flops / 1e9:.1f}")
    print(f"   ffn1+gelu' {f_g[0]:7.1f} us ({tf(f_g[0]):6.0f} TF)  ffn1 {f_n[0]:7.1f} us ({tf(f_n[0]):6.0f} TF)  ffn2 {f_g[1]:7.1f} us ({tf(f_g[1]):6.0f} TF)")
    print(f"   dgrad2 {b[0]:7.1f} us ({tf(b[0]):6.0f} TF)  dgrad1 {b[1]:7.1f} us ({tf(b[1]):6.0f} TF)  wgrad2 {b[2]:7.1f} us ({tf(b[2]):6.0f} TF)  wgrad1 {b[3]:7.1f} us ({tf(b[3]):6.0f} TF)")
    six = f_g[0] + f_g[1] + sum(b)
    print(f"   six GEMMs {six:7.1f} us = {6 * flops / six / 1e6:6.0f} TF/s = {6 * flops / six / 1e6 / PEAK:.3f} of {PEAK} | ffn_fwd call {wf:7.1f} us, ffn_bwd call (incl. colsum, split-K sums) {wb:7.1f} us")


for nm in (sys.argv[1:] or ["micro", "b4", "b8", "b16"]):
    probe(nm)
