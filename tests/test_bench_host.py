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
    b = _bench()
    t = b.traffic_from_profile()
    assert t is None or (t["bytes"] > 0 and t["algorithmic_bytes"] > 0 and len(t["shape"]) == 3)
    pk = b.peaks()
    assert pk["hbm"] > 1000 and pk["bf16_sustained"] > 100
    assert b.shard(12288, 4096, "column", 8) == (1536, 4096) and b.shard(4096, 11008, "row", 8) == (4096, 1376)
    for M, lins in b.WORKLOADS.values():
        for _, N, K, mode in lins:
            for tp in (1, 2, 4, 8):
                n, k = b.shard(N, K, mode, tp)
                assert n % 8 == 0 and k % 16 == 0 and k >= 128, (N, K, mode, tp)   # kernel requirements per shard
    s = b.ClockSampler(0)
    s.lines = [(1.0, "1500, 1965, 900.5, Not Active, Not Active, Not Active, Active\n"), (5.0, "300, 1965, 100, Not Active, Not Active, Not Active, Not Active\n")]
    s.proc, s.t = type("P", (), {"terminate": lambda self: None})(), type("T", (), {"join": lambda self, timeout=None: None})()
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


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU arm; the others print nothing and exit 0 (and do not touch the build)."""
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == "", (out.stdout[-200:], out.stderr[-200:])
