"""First-contact diagnostics for the GPU box: layout probes for the tcgen05 kernel (identity
weights, one-hot rows) with mismatch-pattern summaries, plus quick CUDA-event timings of both
stages and of the reference's kernels.  Writes gpurun_out/diag.json.  Not a test."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import refgpu  # noqa: E402
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402
from oracle import oracle as O  # noqa: E402

DEV = "cuda"
res = {}
CFGS = (1, 5, 6, 8)


_gws = []


def GWS():
    if not _gws:
        _gws.append(torch.zeros(B.load().mixq_decode_workspace_size(1024, 28672), dtype=torch.uint8, device=DEV))
    return _gws[0]


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def pattern(got, ref):
    bad = got.view(np.uint16) != ref.view(np.uint16)
    if not bad.any():
        return "ok"
    r, c = np.nonzero(bad)
    return dict(n_bad=int(bad.sum()), frac=float(bad.mean()), rows_mod8=np.bincount(r % 8, minlength=8).tolist(),
                cols_mod8=np.bincount(c % 8, minlength=8).tolist(), row_range=[int(r.min()), int(r.max())],
                col_range=[int(c.min()), int(c.max())], first=[[int(a), int(b), float(got[a, b]), float(ref[a, b])]
                                                              for a, b in list(zip(r, c))[:6]])


def run_gemm(cfg, q, w, sa, sb, fpA=None, fpW=None):
    lib = B.load()
    M, N = q.shape[0], w.shape[0]
    out = torch.full((M, N), float("nan"), dtype=torch.float16, device=DEV)
    B.gemm_dequant(t(q), t(w), t(sa), t(sb), None if fpA is None else t(fpA), None if fpW is None else t(fpW), out,
                   workspace=GWS(), config=cfg)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def layout_probes():
    rng = np.random.default_rng(0)
    for cfg in CFGS:
        for (M, N, K) in [(128, 128, 128), (128, 256, 256), (256, 128, 512), (100, 136, 144), (384, 512, 1040)]:
            key = f"cfg{cfg}_{M}x{N}x{K}"
            try:
                q = rng.integers(-127, 128, (M, K), dtype=np.int8)
                w = np.zeros((N, K), dtype=np.int8)
                for n in range(min(N, K)):
                    w[n, n] = 1
                one = np.ones(M, dtype=np.float16)
                oneN = np.ones(N, dtype=np.float16)
                got = run_gemm(cfg, q, w, one, oneN)
                ref = O.epilogue(O.igemm(q, w), one, oneN, None)
                res[key + "_identity"] = pattern(got, ref)
                w = rng.integers(-128, 128, (N, K), dtype=np.int8)
                sa = (rng.random(M) * 0.05 + 1e-3).astype(np.float16)
                sb = (rng.random(N) * 0.002 + 1e-4).astype(np.float16)
                got = run_gemm(cfg, q, w, sa, sb)
                res[key + "_random"] = pattern(got, O.epilogue(O.igemm(q, w), sa, sb, None))
                fpA = (rng.standard_normal((M, 128))).astype(np.float16)
                fpW = (rng.standard_normal((N, 128)) * 0.1).astype(np.float16)
                got = run_gemm(cfg, np.zeros_like(q), w, sa, sb, fpA, fpW)
                ref = O.outlier_gemm(fpA, fpW)
                d = np.abs(got.astype(np.float32) - ref.astype(np.float32))
                res[key + "_outlier_only"] = dict(max_abs=float(np.nanmax(d)), n_nan=int(np.isnan(got.astype(np.float32)).sum()),
                                                  bit_mismatch=float((got.view(np.uint16) != ref.view(np.uint16)).mean()))
            except Exception as e:  # keep going: we want the whole picture from one GPU call
                res[key + "_error"] = repr(e)
            print(key, {k: v for k, v in res.items() if k.startswith(key)}, flush=True)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def timings():
    lib = B.load()
    for (M, N, K) in [(512, 12288, 4096), (32, 12288, 4096), (128, 12288, 4096), (1024, 12288, 4096), (2048, 12288, 4096),
                      (4096, 12288, 4096), (8192, 12288, 4096), (512, 4096, 4096), (512, 4096, 11008), (32, 4096, 11008)]:
        key = f"time_{M}x{N}x{K}"
        try:
            A = (torch.randn(M, K, device=DEV) * 0.5).half()
            W8 = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=DEV)
            sb = (torch.rand(N, device=DEV) * 0.002 + 1e-4).half()
            fw = (torch.randn(N, 128, device=DEV) * 0.02).half()
            ind = torch.randperm(K, device=DEV)[:128].int()
            out = torch.empty(M, N, dtype=torch.float16, device=DEV)
            ws = torch.empty(B.workspace_size(M, N, K), dtype=torch.uint8, device=DEV)
            A8 = torch.empty(M, K, dtype=torch.int8, device=DEV)
            sa = torch.empty(M, dtype=torch.float16, device=DEV)
            fpA = torch.empty(M, 128, dtype=torch.float16, device=DEV)
            r = {}
            r["quant_us"] = timeit(lambda: B.quant_extract(A, ind, A8, sa, fpA))
            r["quant_GBps"] = (3 * M * K + 258 * M) / r["quant_us"] / 1e3
            for cfg in CFGS:
                us = timeit(lambda: B.gemm_dequant(A8, W8, sa, sb, fpA, fw, out, workspace=GWS(), config=cfg))
                r[f"gemm_cfg{cfg}_us"] = us
                r[f"gemm_cfg{cfg}_TOPS"] = 2.0 * M * N * K / us / 1e6
                us = timeit(lambda: B.gemm_dequant(A8, W8, sa, sb, None, None, out, workspace=GWS(), config=cfg))
                r[f"gemm_cfg{cfg}_noout_TOPS"] = 2.0 * M * N * K / us / 1e6
            r["enqueue_us"] = timeit(lambda: B.enqueue(A, W8, sb, fw, ind, out, ws))
            r["enqueue_TOPS"] = 2.0 * M * N * K / r["enqueue_us"] / 1e6
            if refgpu.available():
                rws = torch.empty(refgpu.load().ref_workspace_size(M, N, K), dtype=torch.uint8, device=DEV)
                rout = torch.empty_like(out)
                r["ref_enqueue_us"] = timeit(lambda: refgpu.enqueue(A, W8, sb, fw, ind, rout, rws), iters=5, warm=2)
                r["ref_enqueue_TOPS"] = 2.0 * M * N * K / r["ref_enqueue_us"] / 1e6
                r["ref_quant_us"] = timeit(lambda: refgpu.int8quant(A), iters=5, warm=2)
            if M <= 8192:
                a8 = A8.clone()
                r["torch_int_mm_TOPS"] = 2.0 * M * N * K / timeit(lambda: torch._int_mm(a8, W8.t())) / 1e6
            res[key] = r
        except Exception as e:
            res[key + "_error"] = repr(e)
        print(key, res.get(key, res.get(key + "_error")), flush=True)
        del A, W8, out, ws, A8
        torch.cuda.empty_cache()
    try:
        a = torch.randint(-128, 128, (8192, 8192), dtype=torch.int8, device=DEV)
        b = torch.randint(-128, 128, (8192, 8192), dtype=torch.int8, device=DEV)
        res["int8_peak_int_mm_8192_TOPS"] = 2.0 * 8192 ** 3 / timeit(lambda: torch._int_mm(a, b.t()), iters=10) / 1e6
        x = torch.randn(8192, 8192, device=DEV).bfloat16()
        res["bf16_mm_8192_TFLOPS"] = 2.0 * 8192 ** 3 / timeit(lambda: x @ x, iters=10) / 1e6
    except Exception as e:
        res["peak_error"] = repr(e)
    print({k: v for k, v in res.items() if "peak" in k or "bf16" in k}, flush=True)


if __name__ == "__main__":
    B.require_device()
    print(B.load().mixq_version().decode(), torch.cuda.get_device_name(0), flush=True)
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    what = sys.argv[1:] or ["layout", "time"]
    if "layout" in what:
        layout_probes()
        (out / "diag.json").write_text(json.dumps(res, indent=1))
    if "time" in what:
        timings()
    (out / "diag.json").write_text(json.dumps(res, indent=1))
