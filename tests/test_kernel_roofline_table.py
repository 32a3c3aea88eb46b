"""The per-kernel roofline table of bench.py: the algorithmic byte / flop model on CPU (synthetic profiling records), and on the
GPU that level-2 profiling really brackets every launcher of one block call and that the table attributes them to the right shape."""
import ctypes

import pytest

import bench


def _flags(k, dense=False, extra=0):
    return k | (0x100 if dense else 0) | extra


def test_kernel_model_bytes_and_flops():
    T, d, E = 65536, 128, 4
    R = T * E
    # grouped expert GEMM of the 16x16 block: 2 * R * N * K flops; A rows + stacked weights + two bf16 outputs
    name, label, byt, flops, bound = bench.kernel_model(100000 + 256 + 4 * 10000, (T, d, E, _flags(4, True), 4 * d, d, -1, E | 0x10000))
    assert bound == "tensor" and flops == 2.0 * R * 4 * d * d == 8.0 * R * d * d
    assert byt == R * d * 2 + E * 4 * d * d * 2 + R * 4 * d * 2 * 2
    assert "GELU" in name and "R=262144" in label
    # ungrouped router GEMM: the host-known row count wins over the block shape
    _, _, _, flops, _ = bench.kernel_model(256, (T, d, E, _flags(4, True), d, 128, T, 1))
    assert flops == 2.0 * T * d * 128
    # combine: reads every dispatched row once, the gates, the bf16 residual; writes bf16
    _, _, byt, flops, bound = bench.kernel_model(bench.PROF_BASE + 8, (T, d, E, _flags(4, True, 0x2000 | 0x4000), 0, 0, 0, 0))
    assert bound == "hbm" and flops is None and byt == R * d * 2 + T * E * 4 + T * d * 2 + T * d * 2
    # sparse top-2 combine (configs[1]): gates + token_rows, fp32 out, no residual
    _, _, byt, _, _ = bench.kernel_model(bench.PROF_BASE + 8, (4096, 512, 8, _flags(2), 0, 0, 0, 0))
    assert byt == 8192 * 512 * 2 + 4096 * 2 * 4 * 2 + 4096 * 512 * 4
    # router forward is bound by the fp32 FMA pipe, not by bytes
    _, _, byt, flops, bound = bench.kernel_model(bench.PROF_BASE + 3, (4096, 512, 8, _flags(2, False, 0x1000), 0, 0, 0, 0))
    assert bound == "fp32" and flops == 2.0 * 4096 * 512 * 128 + 2.0 * 4096 * 128 * 8 and byt > 4096 * 512 * 4
    # tiny kernels carry no roofline
    assert bench.kernel_model(bench.PROF_BASE + 4, (4096, 512, 8, _flags(2), 0, 0, 0, 0))[4] == "latency"
    # layers around the block: LayerNorm bf16 in / fp32 out
    _, _, byt, _, _ = bench.kernel_model(bench.PROF_BASE + 20, (1000, 256, 0, 1, 0, 0, 0, 0))
    assert byt == 1000 * 256 * (2 + 4) + 1000 * 8


def test_kernel_table_groups_and_fractions():
    pk = {"hbm_gbs": 6400.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}
    meta = (65536, 128, 4, _flags(4, True, 0x1000), 0, 0, 0, 0)
    recs = [(0.040, bench.PROF_BASE + 11, meta), (0.060, bench.PROF_BASE + 11, meta), (0.004, bench.PROF_BASE + 4, meta)]
    rows = bench.kernel_table(recs, 2, pk)
    assert [r["kernel"] for r in rows] == ["uncombine", "plan"]
    u = rows[0]
    assert u["launches_per_step"] == 1.0 and u["us"] == 50.0 and u["us_per_step"] == 50.0
    byt = 65536 * 4 * 128 * 2 + 2 * 65536 * 128 * 4
    assert u["bytes"] == byt and abs(u["GBps"] - byt / 50e-6 / 1e9) < 0.1
    assert abs(u["frac_hbm_measured"] - u["GBps"] / 6400.0) < 1e-3 and abs(u["frac_hbm_8TBps"] - u["GBps"] / 8000.0) < 1e-3
    assert u["roof"] == "hbm" and u["frac_of_roof"] == u["frac_hbm_measured"]
    assert "GBps" not in rows[1] and rows[1]["bound"] == "latency"
    # a K = d GEMM of the 16x16 block is bounded by HBM (its [R, 4d] outputs), not by the tensor pipe
    g = bench.kernel_table([(0.100, 100000 + 256 + 4 * 10000, (65536, 128, 4, _flags(4, True), 512, 128, -1, 4 | 0x10000))], 1, pk)[0]
    assert g["roof"] == "hbm" and 0.9 < g["frac_of_roof"] < 1.0 and g["frac_compute"] < 0.3


@pytest.mark.gpu
def test_level2_profiling_brackets_every_launcher(pkg):
    import torch
    C = pkg._cabi
    lib = C.lib()
    d, Dt, E, k, B, H, W = 256, 512, 8, 2, 4, 8, 8
    torch.manual_seed(0)
    m = pkg.SparseMoE(d, Dt, E, top_k=k, precision="bf16").cuda().train()
    x = torch.randn(B, H, W, d, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)
    w = torch.randn(B, Dt, device="cuda", requires_grad=True)
    for level in (1, 2):
        n0 = lib.moegan_debug_launch_count()
        lib.moegan_debug_profile_enable(level)
        out, kl, probs = m(x, w, 1.0)
        (out.sum() + kl).backward()
        torch.cuda.synchronize()
        n1 = lib.moegan_debug_launch_count()
        cap = 1024
        ms, tag, meta = (ctypes.c_float * cap)(), (ctypes.c_int32 * cap)(), (ctypes.c_int32 * (8 * cap))()
        n = lib.moegan_debug_profile_read_ex(ms, tag, meta, cap)
        lib.moegan_debug_profile_enable(0)
        recs = [(ms[i], tag[i], tuple(meta[8 * i: 8 * i + 8])) for i in range(n)]
        gemms = [r for r in recs if r[1] < bench.PROF_BASE]
        assert len(gemms) >= 6 and all(r[0] > 0 for r in recs)
        if level == 1:
            assert len(gemms) == n
            continue
        ids = {r[1] - bench.PROF_BASE for r in recs if r[1] >= bench.PROF_BASE}
        assert {1, 2, 3, 4, 5, 8, 9, 11, 12, 13, 15} <= ids, ids            # router, dispatch, combine, their backwards
        # one record per launcher call; a launcher may enqueue more than one kernel (slab_sum + finish, sample_kl + finalize)
        assert n <= n1 - n0
        T = B * H * W
        blk = [r for r in recs if r[1] - bench.PROF_BASE in (3, 5, 8, 9, 11)]
        assert all(r[2][0] == T and r[2][1] == d and r[2][2] == E and (r[2][3] & 0xff) == k for r in blk)
        expert = [r for r in gemms if (r[2][7] & 0xffff) == E]
        assert len(expert) == 6 and all(r[2][4] * r[2][5] == 4 * d * d for r in expert)
        rows = bench.kernel_table(recs, 1, bench.peaks())
        assert any(r["kernel"] == "combine_fwd" and r["bytes"] == T * k * d * 2 + T * k * 8 + T * d * 4 for r in rows)
    # nothing left behind
    assert lib.moegan_debug_profile_read_ex(ms, tag, meta, cap) == 0


def test_gemm_roof_summary_weights_by_time():
    rows = [{"bound": "tensor", "roof": "hbm", "frac_of_roof": 0.8, "us_per_step": 300.0},
            {"bound": "tensor", "roof": "tensor", "frac_of_roof": 0.5, "us_per_step": 100.0},
            {"bound": "hbm", "roof": "hbm", "frac_of_roof": 0.3, "us_per_step": 1000.0}]
    s = bench.gemm_roof_summary(rows)
    assert s["frac_of_lower_roof"] == 0.725 and s["hbm_bound_share_of_gemm_time"] == 0.75
    assert bench.gemm_roof_summary([rows[2]]) is None
