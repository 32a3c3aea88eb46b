"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2 (no kernels involved):
bucketed gradient all-reduce (GradReducer) and the exact-under-DP balance loss (dp_allreduce_load)."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "moe-gan_cpsc541_b200"


def _balance_torch(load, T, bw=0.01):
    """t2i_moe_gan.py:973-1000 on a load vector."""
    E = load.shape[0]
    frac = (load + 1e-6) / T
    cv = frac.std() / (frac.mean() + 1e-6)
    return bw * torch.nan_to_num(torch.clamp(E * cv, 0.0, 10.0), nan=0.0)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pkg = importlib.import_module(PKG)
        torch.manual_seed(0)                                    # identical weights on every rank
        model = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.GELU(),
                                    torch.nn.Linear(64, 8))
        unused = torch.nn.Parameter(torch.ones(5))             # a parameter that never gets a gradient
        params = list(model.parameters()) + [unused]
        red = pkg.GradReducer(params, bucket_bytes=4096)       # forces several buckets
        assert len(red.buckets) >= 3
        torch.manual_seed(100 + rank)                           # rank-offset data
        x = torch.randn(12, 16)
        red.begin()
        model(x).square().mean().backward()
        red.finish()
        got = [p.grad.clone() for p in model.parameters()]
        # reference: average of the per-rank gradients computed by hand
        ref = []
        for r in range(world):
            torch.manual_seed(0)
            m2 = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.GELU(),
                                     torch.nn.Linear(64, 8))
            torch.manual_seed(100 + r)
            m2(torch.randn(12, 16)).square().mean().backward()
            ref.append([p.grad for p in m2.parameters()])
        for i, g in enumerate(got):
            want = sum(ref[r][i] for r in range(world)) / world
            assert torch.allclose(g, want, atol=1e-6), i
        assert torch.count_nonzero(unused.grad) == 0
        # the reduced gradients are VIEWS of the bucket buffers (no copy back), and the last bucket is the short tail
        spans = [(f.data_ptr(), f.data_ptr() + f.numel() * 4) for f in red.flat]
        assert all(any(a <= p.grad.data_ptr() < b for a, b in spans) for p in params)
        assert sum(p.numel() for p in red.buckets[-1]) * 4 <= 8 << 20
        # a second step: fresh gradients (set_to_none) are packed again; gradients left in place are accumulated into and re-averaged
        for p in params:
            p.grad = None
        red.begin(); model(x).square().mean().backward(); red.finish()
        for i, (g, p) in enumerate(zip(got, model.parameters())):
            assert torch.allclose(p.grad, g, atol=1e-6), i
        red.begin(); model(x).square().mean().backward(); red.finish()          # no zero_grad: avg(avg_prev + local) = avg_prev + avg_new
        for i, (g, p) in enumerate(zip(got, model.parameters())):
            assert torch.allclose(p.grad, 2 * g, atol=2e-6), i

        # channels-last conv weights (the bench's model layout): gradients are dense but not row-major contiguous; they
        # are packed / unpacked through storage-order views, never through per-parameter copies
        def conv_net():
            torch.manual_seed(1)
            return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.LeakyReLU(0.2),
                                       torch.nn.Conv2d(8, 4, 3, padding=1)).to(memory_format=torch.channels_last)
        net = conv_net()
        redc = pkg.GradReducer(list(net.parameters()), bucket_bytes=512)
        torch.manual_seed(200 + rank)
        xi = torch.randn(2, 3, 6, 6).contiguous(memory_format=torch.channels_last)
        redc.begin(); net(xi).square().mean().backward(); redc.finish()
        assert any(not p.grad.is_contiguous() for p in net.parameters()), "test needs a channels-last gradient"
        for i, p in enumerate(net.parameters()):
            want = 0
            for r in range(world):
                n2 = conv_net()
                torch.manual_seed(200 + r)
                n2(torch.randn(2, 3, 6, 6).contiguous(memory_format=torch.channels_last)).square().mean().backward()
                want = want + list(n2.parameters())[i].grad
            assert torch.allclose(p.grad, want / world, atol=1e-6), i
        v = pkg.trainer._flat_storage_view(torch.empty(4, 3, 2, 2).to(memory_format=torch.channels_last))
        assert v is not None and v.numel() == 48 and pkg.trainer._flat_storage_view(torch.empty(4, 6)[:, ::2]) is None

        # exact balance loss under DP: loss(global load) and its gradient after DP averaging
        torch.manual_seed(7)
        probs_all = torch.softmax(torch.randn(world * 10, 4) * 2, dim=1)
        full = probs_all.clone().requires_grad_(True)
        _balance_torch(full.sum(0), full.shape[0]).backward()
        mine = probs_all[rank * 10:(rank + 1) * 10].clone().requires_grad_(True)
        load = pkg.dp_allreduce_load(mine.sum(0))
        loss = _balance_torch(load, world * 10)
        loss.backward()
        assert torch.allclose(load, probs_all.sum(0), atol=1e-6)
        # DP averages gradients over ranks: (1/world) * (world * dL/dp) == single-process gradient
        assert torch.allclose(mine.grad / world, full.grad[rank * 10:(rank + 1) * 10], atol=1e-7)
        q.put((rank, "ok"))
    except Exception as e:  # noqa
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_grad_reducer_and_dp_balance_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_schedules_match_reference_formulas(pkg):
    t = pkg.trainer
    assert t.temperature_factor(0) == 3.0 and t.temperature_factor(25) == 1.0          # :1178
    assert t.effective_kl_weight(1e-3, 0) == pytest.approx(1e-8)                        # :1168-1174
    assert t.effective_kl_weight(1e-3, 5) == pytest.approx(1e-3)
