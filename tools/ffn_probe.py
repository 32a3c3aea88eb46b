#!/usr/bin/env python
"""Isolate epilogue cost: moegan_ffn_fwd (two grouped GEMMs) with / without the gelu' output, at configs[1] size."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("moe-gan_cpsc541_b200")
C = pkg._cabi; lib = C.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
T, d, Dt, E, k = 4096, 512, 512, 8, 2
shp = C.make_shape(T, d, Dt, E, k, 8, 512, 0, False, False)
m_cap = shp.m_cap; hid = 4 * d
dev = "cuda"
# uniform routing plan built on the host: every expert gets T*k/E rows
plan = torch.zeros(int(lib.moegan_plan_ints(m_cap)), dtype=torch.int32)
rows = T * k // E
off = 0
for e in range(E):
    plan[8 + e] = off; plan[48 + e] = rows
    for b in range(off // 128, (off + rows) // 128): plan[88 + b] = e
    off += (rows + 255) // 256 * 256
plan[8 + E] = off; plan[0] = off // 128; plan[1] = off
plan = plan.to(dev)
a_op = torch.randn(m_cap, d, device=dev).bfloat16()
W1 = (torch.randn(E * hid, d, device=dev) * 0.05).bfloat16(); W2 = (torch.randn(E * d, hid, device=dev) * 0.05).bfloat16()
b1 = torch.zeros(E, hid, device=dev); b2 = torch.zeros(E, d, device=dev)
pre = torch.empty(m_cap, hid, device=dev, dtype=torch.bfloat16); h = torch.empty(m_cap, hid, device=dev, dtype=torch.bfloat16)
y = torch.empty(m_cap, d, device=dev, dtype=torch.bfloat16)
lib.moegan_debug_profile_enable(1)
def run(with_g, n=20):
    for i in range(n + 3):
        flush.zero_()
        C.check(lib.moegan_ffn_fwd(shp, C.ptr(plan), C.ptr(a_op), C.ptr(W1), C.ptr(b1), C.ptr(W2), C.ptr(b2), C.ptr(pre) if with_g else None, C.ptr(h), C.ptr(y), st), "ffn_fwd")
    torch.cuda.synchronize()
    ms = (C.c_float * 4096)(); tg = (C.c_int32 * 4096)()
    cnt = lib.moegan_debug_profile_read(ms, tg, 4096)
    v = [ms[i] for i in range(cnt)][6:]
    return 1e3 * sum(v[0::2]) / len(v[0::2]), 1e3 * sum(v[1::2]) / len(v[1::2])
print("blocks in use", int(plan[0]))
print("ffn1 with gelu' (2 outputs): %.1f us   ffn2: %.1f us" % run(True))
print("ffn1 without gelu' (1 output): %.1f us   ffn2: %.1f us" % run(False))
