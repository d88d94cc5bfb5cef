"""Capture golden vectors from the REFERENCE'S OWN CUDA kernels on a B200.

    gpurun -- python tests/golden/make_gpu_golden.py        # writes gpurun_out/golden/*
    cp gpurun_out/golden/* tests/golden/                    # then commit

Runs oracle/_ref (reference kernel/i8gemm.cu compiled unmodified for sm_100a + our replay of
TsinghuaMixQPlugin.cpp:518-532) on seeded inputs and stores:
  rcp_approx_f16.bin    rcp.approx.ftz.f32 for all 65536 fp16 inputs (what device __hdiv uses)
  ref_kernels_b200.npz  inputs + outputs of FindRowScaleKernel<256>, the outlier gather and the
                        whole enqueue (gather -> cuBLAS fp16 -> int8quant -> CUTLASS GemmDequant)
The CPU tests (tests/test_oracle.py) pin the oracle to these files.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import refgpu  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    out = ROOT / "gpurun_out" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    tab = refgpu.rcp_table()
    tab.tofile(out / "rcp_approx_f16.bin")

    a = np.load(ROOT / "tests/golden/act_scales_l0.npz")
    scale = a["Llama-2-7b/self_attn.q_proj"][:512].copy()
    lin = O.synth_linear(64, 512, scale, seed=7)
    A = O.synth_activations(48, lin["act_scale"], seed=8)
    A[3] = 0                                   # all-zero token
    A[4, :] = np.float16(6e-8)                 # scale underflows to 0
    A[5, 17] = np.float16(65504.0)
    A[6, 100] = np.float16(np.inf)
    A[7, 200] = np.float16(np.nan)
    A[8] = (np.arange(512) % 255 - 127).astype(np.float16) * np.float16(0.5)   # many exact .5 ties
    dev = "cuda"
    tA = torch.from_numpy(A).to(dev)
    tind = torch.from_numpy(lin["ind"]).to(dev)
    q, sa = refgpu.int8quant(tA)
    fpA = refgpu.extract(tA, tind)
    ref_out = refgpu.enqueue(tA, torch.from_numpy(lin["W8"]).to(dev), torch.from_numpy(lin["scale_b"]).to(dev),
                             torch.from_numpy(lin["fp_weight"]).to(dev), tind)
    torch.cuda.synchronize()
    np.savez_compressed(out / "ref_kernels_b200.npz", A=A, ind=lin["ind"], W8=lin["W8"], scale_b=lin["scale_b"],
                        fp_weight=lin["fp_weight"], ref_q=q.cpu().numpy(), ref_sa=sa.cpu().numpy(),
                        ref_fpA=fpA.cpu().numpy(), ref_out=ref_out.cpu().numpy())
    # quick self-report: does the oracle (with the fresh table) reproduce the reference kernels?
    tabp = np.ascontiguousarray(tab)
    O._rcp_table, O._rcp_loaded = tabp, True
    oq, osa = O.quant(A)
    print("rcp table: max ulp vs 1/x =", int(np.abs(tab.view(np.int32).astype(np.int64)[1:0x7C00] -
          (np.float32(1) / np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32)[1:0x7C00]).view(np.int32).astype(np.int64)).max()))
    print("oracle vs reference FindRowScaleKernel: sa equal", np.array_equal(osa.view(np.uint16), sa.cpu().numpy().view(np.uint16)),
          "q mismatches", int((oq != q.cpu().numpy()).sum()), "of", oq.size)
    oo = O.forward(A, lin["W8"], lin["scale_b"], lin["fp_weight"], lin["ind"])
    ro = ref_out.cpu().numpy()
    fin = np.isfinite(ro.astype(np.float32)) & np.isfinite(oo.astype(np.float32))
    print("oracle vs reference enqueue: bit mismatches", int((oo.view(np.uint16) != ro.view(np.uint16)).sum()), "of", oo.size,
          "max abs diff", float(np.abs(oo.astype(np.float32) - ro.astype(np.float32))[fin].max()))


if __name__ == "__main__":
    main()
