"""Multi-rank host logic on CPU: world_size-2 gloo process group, the sharding of
mixq_tensorrt_llm_b200/tp.py applied to an oracle-computed linear.  Column-parallel must
reproduce its output slice bit for bit with no collective; row-parallel needs exactly one
all-reduce and agrees with the single-rank result within rel-Frobenius 2e-3."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import tp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ["OMP_NUM_THREADS"] = "2"
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, K, M = 256, 1024, 48
        lin = O.synth_linear(N, K, seed=11)                      # every rank builds the same full linear
        A = O.synth_activations(M, lin["act_scale"], seed=12)
        sh = tp.shard_linear(lin, mode, world, rank)
        A_r = tp.shard_activations(A, sh)
        y = O.forward(A_r, sh["W8"], sh["scale_b"], sh["fp_weight"], sh["ind"])
        collectives = 0
        if mode == "column":
            parts = [torch.empty(M, N // world, dtype=torch.float16) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(y))          # only to CHECK: the path itself needs no collective
            full = torch.cat(parts, dim=1).numpy()
        else:
            t = torch.from_numpy(y.astype(np.float32))           # gloo has no fp16 sum; fp32 sum of 2 fp16 + RN == fp16 add
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            collectives += 1
            full = t.numpy().astype(np.float16)
        if rank == 0:
            ref = O.forward(A, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"])
            dense = A.astype(np.float64) @ lin["W"].astype(np.float64).T
            q.put((full, ref, collectives, sh.get("n_outliers_local", 128), dense))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_column_parallel_is_bit_exact_without_collective():
    full, ref, collectives, _, _ = _run("column")
    assert collectives == 0
    assert np.array_equal(full.view(np.uint16), ref.view(np.uint16))


def test_row_parallel_one_allreduce_within_tolerance():
    full, ref, collectives, n_local, dense = _run("row")
    assert collectives == 1
    assert 0 <= n_local <= 128
    # K-sharding changes the per-token scales (each rank quantises its own slice), so the sharded and
    # the single-rank results are two different W8A8 roundings of the same product: compare both with
    # the unquantised product.  The sharded one must be no less accurate (its scales are finer), and
    # the two must differ by no more than their quantisation noise.
    nrm = np.linalg.norm(dense)
    err_tp = np.linalg.norm(full.astype(np.float64) - dense) / nrm
    err_1 = np.linalg.norm(ref.astype(np.float64) - dense) / nrm
    assert err_tp <= 1.05 * err_1, (err_tp, err_1)
    assert np.linalg.norm(full.astype(np.float64) - ref.astype(np.float64)) / nrm <= err_tp + err_1


def test_row_shard_outlier_bookkeeping():
    from oracle import oracle as O
    lin = O.synth_linear(64, 512, seed=3)
    seen = []
    for r in range(4):
        sh = tp.shard_row(lin, 4, r)
        lo, hi = sh["k_range"]
        n = sh["n_outliers_local"]
        assert sh["W8"].shape == (64, 128) and sh["fp_weight"].shape == (64, 128) and sh["ind"].shape == (128,)
        assert ((sh["ind"][:n] >= 0) & (sh["ind"][:n] < hi - lo)).all()
        assert (sh["ind"][n:] == 0).all() and (sh["fp_weight"][:, n:] == 0).all()     # padding contributes 0
        assert (sh["W8"][:, sh["ind"][:n]] == 0).all()                                # outlier columns stay zero in W8
        seen += (sh["ind"][:n] + lo).tolist()
    assert sorted(seen) == sorted(lin["ind"].tolist())                               # every outlier lands on exactly one rank
    with pytest.raises(ValueError):
        tp.shard_row(lin, 3, 0)


def test_mixqlinear_tp_shapes():
    """MixQLinear mirrors the reference's parameter shapes (plugin.py:99-123) per shard; no GPU needed to build it."""
    from mixq_tensorrt_llm_b200.plugin import MixQLinear
    col = MixQLinear(4096, 12288, tp_size=8, parallel_mode="column")
    assert col.weight.shape == (1536, 2048) and col.fp_weight.shape == (1536, 128) and col.fp_ind.shape == (256,)
    assert col.weights_scaling_factor.shape == (1536,) and col.weight.dtype == torch.float16
    row = MixQLinear(11008, 4096, tp_size=8, parallel_mode="row")
    assert row.weight.shape == (4096, 688) and row.in_features == 1376 and row.out_features == 4096


def test_qweight_shards_equal_processing_the_sharded_codes():
    """The EETQ-processed weight-only copy can be cut without undoing its layout: a column (N) or row (K) shard of the
    processed tensor equals processing the matching slice of the plain codes, and the weight-only GEMV over the shards
    reproduces the slice / sums to the unsharded product."""
    import numpy as np
    from oracle import oracle as O
    from mixq_tensorrt_llm_b200 import tp
    rng = np.random.default_rng(0)
    K, N, world = 512, 256, 2
    codes = rng.integers(-128, 128, (K, N), dtype=np.int8)
    proc = O.eetq_preprocess(codes)
    scales = (rng.random(N) * 1e-3 + 1e-4).astype(np.float16)
    A = rng.standard_normal((2, K)).astype(np.float16)
    full = O.gemv_w8a16(A, proc, scales)
    packed = dict(W8=np.zeros((N, K), np.int8), scale_b=scales, fp_weight=np.zeros((N, 128), np.float16),
                  ind=np.arange(128, dtype=np.int32), qweight=proc, scales=scales)
    for r in range(world):
        col = tp.shard_linear(packed, "column", world, r)
        n0, n1 = col["n_range"]
        assert np.array_equal(col["qweight"], O.eetq_preprocess(codes[:, n0:n1]))
        assert np.array_equal(O.gemv_w8a16(A, col["qweight"], col["scales"]).view(np.uint16), full[:, n0:n1].view(np.uint16))
        row = tp.shard_linear(packed, "row", world, r)
        k0, k1 = row["k_range"]
        assert np.array_equal(row["qweight"], O.eetq_preprocess(codes[k0:k1]))
    parts = [O.gemv_w8a16(A[:, k0:k1], tp.shard_qweight_row(proc, world, r), scales).astype(np.float32)
             for r, (k0, k1) in enumerate([(0, K // 2), (K // 2, K)])]
    ref = A.astype(np.float64) @ (codes.astype(np.float64) * scales.astype(np.float64)[None, :])
    assert np.abs(parts[0] + parts[1] - ref).max() <= 4e-3 * np.abs(ref).max() + 2e-3


def _bias_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mixq_tensorrt_llm_b200.plugin import MixQLinear
        N, K, M = 64, 128, 5
        g = torch.Generator().manual_seed(7)
        y_full = torch.randn(M, N, generator=g).half()           # stands for the plugin's column-parallel outputs
        bias_full = torch.randn(N, generator=g).half()
        mod = MixQLinear(K, N, bias=True, tp_group=dist.group.WORLD, tp_size=world, parallel_mode="column")
        lo, hi = rank * (N // world), (rank + 1) * (N // world)
        mod.bias.copy_(bias_full[lo:hi])                          # the bias buffer is this rank's [N / tp] shard
        assert mod.bias.shape == (N // world,)
        x = y_full[:, lo:hi] + mod.bias                           # what forward() does before the gather
        out = mod._gather_columns(x)
        mod.gather_output = False
        assert mod._gather_columns(x) is x
        if rank == 0:
            q.put((out, (y_full + bias_full[None, :])))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_column_parallel_bias_is_added_before_the_gather():
    """tp > 1 with bias (Qwen2 qkv): each rank adds ITS bias shard to its own output columns, then the shards are
    gathered -- the gathered [.., N] tensor equals the unsharded linear's output + the full bias."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bias_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out.shape == want.shape and torch.equal(out, want)
