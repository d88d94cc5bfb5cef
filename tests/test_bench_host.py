"""CPU checks of bench.py's host-side helpers (anything that can break the one JSON line without a GPU)."""
import importlib.util
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench", ROOT / "bench.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_helpers():
    import numpy as np
    import torch
    b = _bench()
    pk = b.peaks()
    assert pk["hbm"] > 1000 and pk["bf16_sustained"] > 100
    assert b.DEFAULT_WORKLOAD == "llama2-7b-linears-decode-bs512" and b.WORKLOADS[b.DEFAULT_WORKLOAD]["M"] == 512   # BASELINE.json's metric config
    for wl in b.WORKLOADS.values():
        for _, N, K, mode, key in wl["linears"]:
            s, src = b.act_scales(wl["scales"], key, K)
            assert s.shape == (K,) and src.startswith("reference")                     # the committed act_scales fixture covers every shape
            for tp in (1, 2, 4, 8):
                n, k = (N // tp, K) if mode == "column" else (N, K // tp)
                assert n % 8 == 0 and k % 16 == 0 and k >= 128, (N, K, mode, tp)       # kernel requirements per shard
    # the device packer / sharder of the bench are the product's host packer and tp.py (CPU tensors here)
    from mixq_tensorrt_llm_b200 import checkpoint, tp
    g = torch.Generator().manual_seed(3)
    W = (torch.randn(96, 512, generator=g) * 0.02).half()
    sc = torch.rand(512, generator=g)
    want = checkpoint.pack_linear_weights(W, sc)
    W8, sb, fw, ind = b.pack_gpu(torch, W.clone(), sc)
    assert torch.equal(W8, want["W8"]) and torch.equal(sb, want["scale_b"]) and torch.equal(fw, want["fp_weight"]) and torch.equal(ind, want["ind"])
    packed = {k: v.numpy() for k, v in want.items()}
    for mode in ("column", "row"):
        for r in range(4):
            ref = tp.shard_linear(packed, mode, 4, r)
            w8, s_, f_, i_, (lo, hi) = b.shard_packed(torch, W8, sb, fw, ind, mode, 4, r)
            assert np.array_equal(w8.numpy(), ref["W8"]) and np.array_equal(s_.numpy(), ref["scale_b"])
            assert np.array_equal(f_.numpy(), ref["fp_weight"]) and np.array_equal(i_.numpy(), ref["ind"]) and (lo, hi) == tuple(ref["k_range"])
    assert b.linear_bytes(512, 12288, 4096) == 2 * 512 * 4096 + 12288 * 4096 + 256 * 12288 + 2 * 12288 + 512 + 2 * 512 * 12288   # 70.3 MB
    s = b.ClockSampler(0)
    s.how, s.sm_max = "nvml", 1965.0
    s.samples = [(1.0, 1500.0, 900.5, 0x4), (5.0, 300.0, 100.0, 0)]
    s.t = type("T", (), {"join": lambda self, timeout=None: None})()
    s.t0, s.t1 = 0.5, 2.0
    r = s.stop()
    assert r["sm_mhz"] == 1500.0 and r["reasons"] == ["sw_power_cap"] and r["samples"] == 1


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-tokens", "8"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["config"]["workload"] == "llama2-7b-linears-decode-bs512" and d["config"]["tokens_per_step"] == 512
    assert set(d["config"]) == {"workload", "tokens_per_step", "layers_per_step", "linears", "parallelism"}


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU arm; the others print nothing and exit 0 (and do not touch the build)."""
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == "", (out.stdout[-200:], out.stderr[-200:])
