"""Tensor-parallel sharding of a packed MixQ linear (host logic, numpy; no GPU needed).

The reference shards MixQ linears along N only and then -- wrongly -- all-reduces the sharded
outputs (reference plugin.py:97,155-156; quantization/quantize.py:331-337), and forbids
row-parallel outright (quantize.py:342 ``assert module.tp_size == 1``).  This module defines
the sharding the B200 build uses instead:

column-parallel (qkv, gate, up)   split N.  Rank r keeps W8[N_r, K], scale_b[N_r],
    fp_weight[N_r, 128] and the full ind[128]; activations are replicated; the output
    [M, N_r] stays sharded -- NO collective.  Bit-identical to the matching slice of the
    unsharded result.

row-parallel (o_proj, down_proj)  split K.  Rank r keeps W8[N, K_r], the full scale_b[N],
    and the subset of the 128 global outlier columns that fall inside its K-slice, re-based to
    the slice and padded to 128 entries.  Padding entries point at local column 0 and carry
    all-zero fp_weight columns, so they contribute exactly 0 and the kernel keeps its one
    shape.  The input A[:, K_r] is already sharded (its producer was column-parallel).  Each
    rank quantises its slice with its own per-token scale and emits a full-size fp16 partial
    [M, N]; ONE all-reduce (sum) over the tensor-parallel group finishes the linear.  Not
    bit-identical to the single-GPU result: the per-token scales are taken per K-slice, so it
    is a different (slightly finer) W8A8 rounding of the same product.  tests/test_tp.py checks
    that its error against the unquantised product is no larger than the single-GPU path's.
"""
from __future__ import annotations

import numpy as np

NUM_OUTLIERS = 128


def shard_bounds(total: int, world: int, rank: int, multiple: int = 1):
    """[lo, hi) of `rank`'s contiguous share of `total`, both multiples of `multiple`."""
    if total % (world * multiple) != 0:
        raise ValueError(f"{total} is not divisible into {world} shards of a multiple of {multiple}")
    step = total // world
    return rank * step, (rank + 1) * step


def shard_column(packed: dict, world: int, rank: int) -> dict:
    """Column-parallel shard: split the output channels (N), keep K and the outlier list whole."""
    N = packed["W8"].shape[0]
    lo, hi = shard_bounds(N, world, rank, multiple=8)       # N_r % 8 == 0: 16-byte output rows
    return dict(W8=np.ascontiguousarray(packed["W8"][lo:hi]), scale_b=np.ascontiguousarray(packed["scale_b"][lo:hi]),
                fp_weight=np.ascontiguousarray(packed["fp_weight"][lo:hi]), ind=packed["ind"].copy(),
                n_range=(lo, hi), k_range=(0, packed["W8"].shape[1]), mode="column")


def shard_row(packed: dict, world: int, rank: int) -> dict:
    """Row-parallel shard: split the input channels (K); outlier columns follow their K-slice."""
    N, K = packed["W8"].shape
    lo, hi = shard_bounds(K, world, rank, multiple=16)      # K_r % 16 == 0: TMA row pitch
    ind = packed["ind"].astype(np.int64)
    mine = np.nonzero((ind >= lo) & (ind < hi))[0]           # positions in the global outlier list
    loc_ind = np.zeros(NUM_OUTLIERS, dtype=np.int32)         # padding -> local column 0 ...
    fp_w = np.zeros((N, NUM_OUTLIERS), dtype=np.float16)     # ... with zero weight: contributes 0
    loc_ind[: mine.size] = (ind[mine] - lo).astype(np.int32)
    fp_w[:, : mine.size] = packed["fp_weight"][:, mine]
    return dict(W8=np.ascontiguousarray(packed["W8"][:, lo:hi]), scale_b=packed["scale_b"].copy(), fp_weight=fp_w,
                ind=loc_ind, n_outliers_local=int(mine.size), n_range=(0, N), k_range=(lo, hi), mode="row")


# ---- the weight-only copy (M <= 4 branch): EETQ-processed int8 [K, N] (checkpoint.eetq_preprocess).  Its bytes are rows
# of 2K codes per PAIR of output channels, each row made of runs of 64 codes that alternate between the two channels
# and cover 64 consecutive (permuted-in-16s) input channels: it can be cut along N at even channels and along K at
# multiples of 64 without undoing the layout.
def shard_qweight_column(qweight: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Column-parallel shard of the processed qweight: output channels [rank*N/world, (rank+1)*N/world)."""
    K, N = qweight.shape
    lo, hi = shard_bounds(N, world, rank, multiple=64)        # 64: what the reference packer accepts per shard
    rows = qweight.reshape(N // 2, 2 * K)
    return np.ascontiguousarray(rows[lo // 2: hi // 2]).reshape(K, hi - lo)


def shard_qweight_row(qweight: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Row-parallel shard of the processed qweight: input channels [rank*K/world, (rank+1)*K/world)."""
    K, N = qweight.shape
    lo, hi = shard_bounds(K, world, rank, multiple=64)
    runs = qweight.reshape(N // 2, K // 64, 128)              # per channel pair: K/64 blocks of (64 codes of ch 0 | 64 of ch 1)
    return np.ascontiguousarray(runs[:, lo // 64: hi // 64]).reshape(hi - lo, N)


def shard_linear(packed: dict, mode: str, world: int, rank: int) -> dict:
    if world == 1:
        d = dict(packed)
        d.update(n_range=(0, packed["W8"].shape[0]), k_range=(0, packed["W8"].shape[1]), mode=mode)
        return d
    if mode not in ("column", "row"):
        raise ValueError("mode must be 'column' or 'row'")
    d = shard_column(packed, world, rank) if mode == "column" else shard_row(packed, world, rank)
    if "qweight" in packed:                                   # the weight-only copy follows the same cut
        d["qweight"] = (shard_qweight_column if mode == "column" else shard_qweight_row)(packed["qweight"], world, rank)
        if "scales" in packed:
            lo, hi = d["n_range"]
            d["scales"] = np.ascontiguousarray(packed["scales"][lo:hi])
    return d


def shard_activations(A: np.ndarray, shard: dict) -> np.ndarray:
    """The slice of the activations a rank's linear consumes (all of K for column-parallel)."""
    lo, hi = shard["k_range"]
    return np.ascontiguousarray(A[:, lo:hi])
