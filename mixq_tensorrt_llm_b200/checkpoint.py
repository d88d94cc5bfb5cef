"""The reference's ``int8_mix`` checkpoint layout: packer, writer and reader (host logic, torch).

SURVEY.md 8f "next #3".  The reference produces the layout in
``pack_linear_weights`` / ``to_quantized_weight``
(modelopt/torch/export/model_config_utils.py:378-466, 298-308) and writes it with
``save_file(weights, rank{r}.safetensors)`` + ``config.json``
(modelopt/torch/export/model_config_export.py:467-498).  Per MixQ linear
``transformer.layers.{i}.{attention.qkv|mlp.gate|mlp.proj}`` the tensors are, all typed float16:

    weight                  [N, K/2]   int8 codes, outlier columns zero, two codes per fp16 slot
    weights_scaling_factor  [N]        max_k |W[n,k]| / 127, taken BEFORE the outlier columns are zeroed
    fp_weight               [N, 128]   the outlier columns of W
    fp_ind                  [256]      128 int32 column indices as raw bytes
    qweight                 [K, N/2]   EETQ weight-only copy of the UN-zeroed weight for the M <= 4 branch: int8 codes
                                       of W^T in the interleaved layout of cutlass_preprocessors.cc:497-533 (:437-440)
    scales                  [N]        its scales, max_k |W[n,k]| / 128 (:441; the plugin is fed weights_scaling_factor
                                       instead, plugin.py:149 -- kept as the reference has it)
    bias                    [N]        optional (Qwen2's qkv): added after the plugin, plugin.py:158-160

This module needs neither mixlib nor EETQ: their ``int8_matrix_to_half`` / ``int_to_half`` helpers are
byte reinterpretations (``Tensor.view(torch.float16)``).  The reference hard-codes its activation
scale file (``act_scales/Qwen2-72B.pt``, :391); here it is an argument.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, Iterable, Mapping, Optional

import torch

NUM_OUTLIERS = 128
MIXQ_LINEARS = ("attention.qkv", "mlp.gate", "mlp.proj")          # model_config_utils.py:409-415
# which activation-scale row picks the outlier columns of which linear (:393-401, 423-425):
ACT_SCALE_KEY = {"attention.qkv": "self_attn.q_proj", "mlp.gate": "mlp.gate_proj", "mlp.proj": "mlp.up_proj"}


_PERM16 = (0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15)     # permute_B_rows_for_mixed_gemm, int8


@torch.no_grad()
def eetq_preprocess(q_kn: torch.Tensor) -> torch.Tensor:
    """int8 codes of W^T, row-major [K, N] -> the processed tensor EETQ hands to its kernels (same nominal shape):
    preprocess_weights_for_mixed_gemm with the Sm80 int8 layout (cutlass_preprocessors.cc:497-533): K permuted in
    groups of 16 (:137-199), transposed (:322-335), channel pairs interleaved in runs of 64 codes (:432-495), +128
    bias and bytes 1 / 2 of every four swapped (:337-358)."""
    K, N = q_kn.shape
    if K % 64 or N % 64:
        raise ValueError("EETQ layout needs K % 64 == 0 and N % 64 == 0 (cutlass_preprocessors.cc:230,455)")
    q = q_kn.to(torch.int8).cpu().reshape(K // 16, 16, N)[:, list(_PERM16), :].reshape(K, N)
    t = q.t().contiguous().reshape(N // 2, 2, K // 64, 64).permute(0, 2, 1, 3).reshape(N // 2, 2 * K)
    u = (t.to(torch.int16) + 128).to(torch.uint8).reshape(-1, 4)[:, [0, 2, 1, 3]]
    return u.contiguous().reshape(K, N).view(torch.int8)


@torch.no_grad()
def eetq_quant_weights(weight_t: torch.Tensor):
    """EETQ.quant_weights(W^T, torch.int8, False) (symmetric_quantize<half, half>, cutlass_preprocessors.cc:581-678):
    per output channel s = max|w| / 128 in fp32, code = clamp(round_half_away(w / s), -128, 127), then the layout
    transform.  Returns (processed int8 [K, N], scales fp16 [N])."""
    w = weight_t.detach().to(torch.float16).cpu().float()
    s = w.abs().amax(dim=0) * (1.0 / 128.0)
    x = w / s[None, :]
    r = torch.where(x >= 0, torch.floor(x + 0.5), torch.ceil(x - 0.5))
    q = torch.nan_to_num(r.clamp(-128, 127), nan=127.0).to(torch.int8)     # std::min(127.f, NaN) keeps 127.f (:639-640)
    return eetq_preprocess(q), s.to(torch.float16)


@torch.no_grad()
def pack_linear_weights(weight: torch.Tensor, act_scale: torch.Tensor, fp_features: int = NUM_OUTLIERS,
                        with_qweight: bool = False) -> Dict[str, torch.Tensor]:
    """One linear, reference order of operations (weight is fp16 [N, K]); returns typed tensors
    (W8 int8 [N,K], scale_b fp16 [N], fp_weight fp16 [N,128], ind int32 [128]) and, with_qweight, the weight-only
    pair (qweight int8 [K,N] processed, scales fp16 [N]) taken from the un-zeroed weight (:437-441 run before :453)."""
    w = weight.detach().to(torch.float16).cpu().clone()
    scale_b = (torch.max(torch.abs(w), dim=1)[0].unsqueeze(1) / 127).to(torch.float16).reshape(w.shape[0])
    extra = {}
    if with_qweight:
        extra["qweight"], extra["scales"] = eetq_quant_weights(w.t().contiguous())
    ind = torch.sort(act_scale.detach().float().cpu(), stable=True)[1][-fp_features:]
    fp_weight = w[:, ind].contiguous()
    w[:, ind] = 0
    # CPU fp16 divide (computed in fp32, rounded to fp16), round-half-even, clamp: to_quantized_weight :303-308
    W8 = (w / scale_b[:, None]).round().clamp(-128, 127).nan_to_num(0).to(torch.int8)
    return {"W8": W8, "scale_b": scale_b, "fp_weight": fp_weight, "ind": ind.to(torch.int32), **extra}


def to_checkpoint_tensors(packed: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Typed tensors -> the fp16-typed containers the reference stores (plugin.py:99-111)."""
    out = {
        "weight": packed["W8"].contiguous().view(torch.float16),
        "weights_scaling_factor": packed["scale_b"].contiguous(),
        "fp_weight": packed["fp_weight"].contiguous(),
        "fp_ind": packed["ind"].to(torch.int32).contiguous().view(torch.float16),
    }
    if "qweight" in packed:
        out["qweight"] = packed["qweight"].contiguous().view(torch.float16)      # [K, N/2], mixlib.int8_matrix_to_half
        out["scales"] = packed["scales"].contiguous()
    if packed.get("bias") is not None:                                           # Qwen2's qkv bias: MixQLinear.bias, plugin.py:131-134
        out["bias"] = packed["bias"].to(torch.float16).contiguous()
    return out


def from_checkpoint_tensors(t: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """fp16-typed containers -> typed tensors."""
    out = {
        "W8": t["weight"].contiguous().view(torch.int8),
        "scale_b": t["weights_scaling_factor"].reshape(-1),
        "fp_weight": t["fp_weight"],
        "ind": t["fp_ind"].contiguous().view(torch.int32),
    }
    if "qweight" in t:
        out["qweight"] = t["qweight"].contiguous().view(torch.int8)
        if "scales" in t:
            out["scales"] = t["scales"].reshape(-1)
    if "bias" in t:
        out["bias"] = t["bias"].reshape(-1)
    return out


def save_checkpoint(path, layers: Iterable[Mapping[str, Mapping[str, torch.Tensor]]], config: Optional[dict] = None,
                    rank: int = 0) -> None:
    """Write ``rank{r}.safetensors`` + ``config.json`` with the reference's key names.
    ``layers[i][linear]`` is a packed dict (see pack_linear_weights)."""
    from safetensors.torch import save_file
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    out = {}
    for i, layer in enumerate(layers):
        for lin, packed in layer.items():
            if lin not in MIXQ_LINEARS:
                raise ValueError(f"{lin} is not a MixQ linear (expected one of {MIXQ_LINEARS})")
            for k, v in to_checkpoint_tensors(packed).items():
                out[f"transformer.layers.{i}.{lin}.{k}"] = v.cpu()
    save_file(out, str(path / f"rank{rank}.safetensors"))
    cfg = dict(config or {})
    cfg.setdefault("quantization", {"quant_algo": "int8_mix"})      # QuantAlgo.int8_mix, quantization/mode.py:37
    (path / "config.json").write_text(json.dumps(cfg, indent=1))


def load_checkpoint(path, rank: int = 0) -> Dict[int, Dict[str, Dict[str, torch.Tensor]]]:
    """Read a reference-layout checkpoint: {layer: {linear: typed packed dict}}."""
    from safetensors import safe_open
    layers: Dict[int, Dict[str, Dict[str, torch.Tensor]]] = {}
    raw: Dict[tuple, Dict[str, torch.Tensor]] = {}
    with safe_open(str(Path(path) / f"rank{rank}.safetensors"), framework="pt") as f:
        for key in f.keys():
            parts = key.split(".")
            if len(parts) < 6 or parts[0] != "transformer" or parts[1] != "layers":
                continue
            lin, name = ".".join(parts[3:5]), parts[5]
            if lin in MIXQ_LINEARS and name in ("weight", "weights_scaling_factor", "fp_weight", "fp_ind", "qweight", "scales", "bias"):
                raw.setdefault((int(parts[2]), lin), {})[name] = f.get_tensor(key)
    for (i, lin), t in raw.items():
        if all(k in t for k in ("weight", "weights_scaling_factor", "fp_weight", "fp_ind")):
            layers.setdefault(i, {})[lin] = from_checkpoint_tensors(t)
    return layers


def load_into(module, packed: Mapping[str, torch.Tensor]):
    """Fill a ``MixQLinear`` from a typed packed dict."""
    dev = module.weight.device
    qw, bias = packed.get("qweight"), packed.get("bias")
    return module.load_packed(packed["W8"].to(dev), packed["scale_b"].to(dev), packed["fp_weight"].to(dev),
                              packed["ind"].to(dev), bias=bias.to(dev) if bias is not None else None,
                              qweight=qw.to(dev) if qw is not None else None)
