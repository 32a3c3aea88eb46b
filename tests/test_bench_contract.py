"""The JSON contract of bench.py's reference arm (the one leg that runs without a GPU): one line on stdout with the keys the
driver reads, the CPU arm (unmodified reference, or the port as fallback) described in `cpu_baseline`, `e2e` repeating the line's own value with zero copy bytes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "train_imgs_per_s" and d["unit"] == "imgs/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    # the unmodified reference whenever oracle/_ref (or /root/reference) is present; the torch-CPU port is only the fallback
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    assert cb["kind"] == ("reference" if ref_loader.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and "sample" in cb
    # the reference arm must not claim the B200 arm's configuration: batch 8 / fp32 / CPU is stated, same_config is false
    assert d["config"]["same_config"] is False and d["config"]["batch_per_step"] == 8 and d["config"]["precision"] == "fp32"
    assert "bounded sample" in d["config"]["reference_arm"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_defaults_and_workload_names():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.TRAIN_BATCH == 256 and bench.CPU_BATCH == 8 and bench.INFER_BATCH == 1024
    assert bench.WL["E"] == 8 and bench.WL["k"] == 2 and bench.WL["B"] * bench.WL["H"] * bench.WL["W"] == 4096 and bench.WL["d"] == 512
    assert "configs[1]" in bench.moe_block_workload_name() and "configs[2]" in bench.train_workload_name(1)


def test_reference_arm_sets_threads_under_a_launcher_env():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the reference arm must still use the host's cores (round-1 SCALE runs timed it on 1)."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    if (os.cpu_count() or 1) > 1:
        assert d["cpu_baseline"]["cores"] > 1, d["cpu_baseline"]
    # the other ranks print nothing and exit 0
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_clock_sampler_maps_cuda_device_to_nvml_handle(monkeypatch):
    """NVML numbers the physical GPUs, CUDA the visible ones: with CUDA_VISIBLE_DEVICES=1 the sampler must not read GPU 0 (an idle GPU's
    120 MHz would look like a stuck clock to the driver).  Lookup order: UUID of the torch device, then CUDA_VISIBLE_DEVICES, then index."""
    import bench

    class FakeNvml:
        def __init__(self, by_uuid=True):
            self.calls, self.by_uuid = [], by_uuid

        def nvmlDeviceGetHandleByUUID(self, u):
            if not self.by_uuid:
                raise RuntimeError("no such uuid")
            self.calls.append(("uuid", u)); return "h-" + u

        def nvmlDeviceGetHandleByIndex(self, i):
            self.calls.append(("index", i)); return f"h-idx{i}"

    import torch

    class Props:
        uuid = "12345678-aaaa-bbbb-cccc-1234567890ab"

    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: Props())
    nv = FakeNvml()
    assert bench.ClockSampler._handle(nv, 0) == "h-GPU-" + Props.uuid and nv.calls[0][0] == "uuid"
    nv = FakeNvml(by_uuid=False)
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "3,1")
    assert bench.ClockSampler._handle(nv, 1) == "h-idx1" and bench.ClockSampler._handle(nv, 0) == "h-idx3"
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "GPU-deadbeef")            # UUID form: fall through to the plain index
    assert bench.ClockSampler._handle(nv, 0) == "h-idx0"
