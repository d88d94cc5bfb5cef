"""Generate the CPU-side golden fixtures from the reference itself.  Run in the dev
container (needs /root/reference); the outputs are committed so the tests run anywhere.

  python tests/golden/make_cpu_golden.py

1. quantized_weight.npz -- the reference's own ``to_quantized_weight``
   (modelopt/torch/export/model_config_utils.py:298-308), extracted from its source file by
   AST (the modelopt package itself is not importable here: SURVEY.md 8c) and executed with
   torch on seeded weights, plus the lines of ``pack_linear_weights`` that need no
   mixlib/EETQ (:429-430 scale, :448 index selection, :452-453 split) executed verbatim in
   torch.  Pins oracle.pack_linear_weights.
2. act_scales_l0.npz -- layer-0 rows of the reference's act_scales/*.pt (real fixtures shipped
   with the reference) for the projections the packer reads, so tests and bench can build
   realistic outlier columns without /root/reference.
"""
import ast
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def load_to_quantized_weight():
    src = (REF / "modelopt/torch/export/model_config_utils.py").read_text()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "to_quantized_weight")
    ns = {"torch": torch, "QUANTIZATION_INT8_MIX": "int8_mix", "QUANTIZATION_FP8": "fp8",
          "QUANTIZATION_INT4_AWQ": "int4_awq", "QUANTIZATION_W4A8_AWQ": "w4a8_awq",
          "QUANTIZATION_NVFP4": "nvfp4", "QUANTIZATION_INT8_SQ": "int8_sq"}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "model_config_utils.py", "exec"), ns)
    return ns["to_quantized_weight"]


def main():
    tqw = load_to_quantized_weight()
    g = torch.Generator().manual_seed(1234)
    N, K, F = 96, 512, 128
    W = (torch.randn(N, K, generator=g) * 0.02).half()
    # a few adversarial rows: exact ties, a huge entry, an all-zero row
    W[1, :8] = torch.tensor([0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 3.5, -3.5]).half() * (W[1].abs().max() / 127)
    W[2, 3] = 3.0
    W[5] = 0
    act = torch.rand(K, generator=g) * 2
    act[torch.randperm(K, generator=g)[:40]] *= 30

    # pack_linear_weights lines, verbatim semantics (torch on CPU, fp16 weight tensor)
    weight = W.clone()
    sb = (torch.max(torch.abs(weight), dim=1)[0].unsqueeze(1) / (127)).to(torch.float16).reshape((weight.shape[0],))
    fp_ind = torch.sort(act)[1][-F:]
    fp_weight = weight[:, fp_ind].clone()
    weight[:, fp_ind] *= 0
    W8 = tqw(weight, sb, "int8_mix")
    np.savez_compressed(OUT / "quantized_weight.npz", W=W.numpy(), act_scale=act.numpy(),
                        scale_b=sb.numpy(), ind=fp_ind.to(torch.int32).numpy(),
                        fp_weight=fp_weight.numpy(), W8=W8.numpy())
    print("quantized_weight.npz", W8.shape, W8.dtype, int(W8.min()), int(W8.max()))

    # real activation-scale fixtures, layer 0 only
    want = {
        "Llama-2-7b": ["self_attn.q_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"],
        "Llama-2-70b": ["self_attn.q_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"],
        "qwen2-7b-instruct": ["self_attn.q_proj", "self_attn.o_proj", "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"],
    }
    out = {}
    for model, projs in want.items():
        d = torch.load(REF / "act_scales" / f"{model}.pt")
        for p in projs:
            v = d[f"model.layers.0.{p}"].float().numpy()
            out[f"{model}/{p}"] = v
            # reference index selection on the real fixture (model_config_utils.py:448)
            out[f"{model}/{p}/ind"] = torch.sort(d[f"model.layers.0.{p}"])[1][-F:].to(torch.int32).numpy()
    np.savez_compressed(OUT / "act_scales_l0.npz", **out)
    print("act_scales_l0.npz", {k: v.shape for k, v in out.items() if not k.endswith("/ind")})


if __name__ == "__main__":
    sys.exit(main())
