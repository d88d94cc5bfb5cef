"""CPU tests of the int8_mix checkpoint layout code: the product packer against the
golden vectors made from the reference's own to_quantized_weight, container round trips, and a
save/load round trip through safetensors with the reference's key names."""
from pathlib import Path

import numpy as np
import torch

from mixq_tensorrt_llm_b200 import checkpoint as ck

GOLD = Path(__file__).resolve().parent / "golden"


def test_product_packer_matches_reference_golden():
    z = np.load(GOLD / "quantized_weight.npz")
    p = ck.pack_linear_weights(torch.from_numpy(z["W"]), torch.from_numpy(z["act_scale"]))
    assert np.array_equal(p["W8"].numpy(), z["W8"])
    assert np.array_equal(p["scale_b"].numpy().view(np.uint16), z["scale_b"].view(np.uint16))
    assert np.array_equal(p["fp_weight"].numpy().view(np.uint16), z["fp_weight"].view(np.uint16))
    assert np.array_equal(p["ind"].numpy(), z["ind"])


def test_product_packer_matches_oracle_packer(oracle):
    rng = np.random.default_rng(3)
    W = (rng.standard_normal((40, 384)) * 0.02).astype(np.float16)
    act = oracle.synth_act_scale(384, 9)
    p = ck.pack_linear_weights(torch.from_numpy(W), torch.from_numpy(act))
    o = oracle.pack_linear_weights(W, act)
    for k in ("W8", "fp_weight", "ind"):
        assert np.array_equal(p[k].numpy(), o[k]), k
    assert np.array_equal(p["scale_b"].numpy().view(np.uint16), o["scale_b"].view(np.uint16))


def test_containers_roundtrip_and_shapes():
    z = np.load(GOLD / "quantized_weight.npz")
    p = ck.pack_linear_weights(torch.from_numpy(z["W"]), torch.from_numpy(z["act_scale"]))
    c = ck.to_checkpoint_tensors(p)
    N, K = z["W"].shape
    assert c["weight"].shape == (N, K // 2) and c["fp_ind"].shape == (256,) and c["fp_weight"].shape == (N, 128)
    assert all(t.dtype == torch.float16 for t in c.values())                 # plugin.py:99-123: everything is half
    back = ck.from_checkpoint_tensors(c)
    for k in ("W8", "scale_b", "fp_weight", "ind"):
        assert torch.equal(back[k], p[k]), k


def test_save_load_reference_layout(tmp_path):
    g = torch.Generator().manual_seed(0)
    layers = []
    for i in range(2):
        layer = {}
        for lin, (N, K) in {"attention.qkv": (48, 256), "mlp.gate": (64, 256), "mlp.proj": (32, 512)}.items():
            W = (torch.randn(N, K, generator=g) * 0.02).half()
            layer[lin] = ck.pack_linear_weights(W, torch.rand(K, generator=g))
        layers.append(layer)
    ck.save_checkpoint(tmp_path, layers, {"architecture": "LlamaForCausalLM"})
    from safetensors import safe_open
    with safe_open(str(tmp_path / "rank0.safetensors"), framework="pt") as f:
        keys = set(f.keys())
    assert "transformer.layers.1.mlp.proj.fp_ind" in keys and "transformer.layers.0.attention.qkv.weight" in keys
    assert len(keys) == 2 * 3 * 4
    loaded = ck.load_checkpoint(tmp_path)
    for i, layer in enumerate(layers):
        for lin, p in layer.items():
            for k in ("W8", "scale_b", "fp_weight", "ind"):
                assert torch.equal(loaded[i][lin][k], p[k]), (i, lin, k)
    import json
    assert json.loads((tmp_path / "config.json").read_text())["quantization"]["quant_algo"] == "int8_mix"


def test_load_into_module():
    from mixq_tensorrt_llm_b200.plugin import MixQLinear
    W = (torch.randn(24, 256) * 0.02).half()
    p = ck.pack_linear_weights(W, torch.rand(256))
    m = ck.load_into(MixQLinear(256, 24), p)
    assert torch.equal(m.weight.view(torch.int8).view(24, 256), p["W8"])
    assert torch.equal(m.fp_ind.view(torch.int32), p["ind"])


def test_eetq_pair_matches_reference_packer_and_roundtrips(tmp_path, oracle):
    """checkpoint.eetq_preprocess / eetq_quant_weights (torch, product side) vs the reference packer's golden vectors
    (tests/golden/eetq_layout.npz) and vs the oracle's numpy restatement; qweight/scales survive save -> load."""
    import numpy as np
    from pathlib import Path
    from mixq_tensorrt_llm_b200 import checkpoint as C
    g = np.load(Path(__file__).resolve().parent / "golden" / "eetq_layout.npz")
    assert np.array_equal(C.eetq_preprocess(torch.from_numpy(g["q_kn"])).numpy(), g["processed"])
    qw, sc = C.eetq_quant_weights(torch.from_numpy(g["W_t"]))
    assert np.array_equal(qw.numpy(), g["processed_codes"])
    assert np.array_equal(sc.numpy().view(np.uint16), g["scales"].view(np.uint16))
    torch.manual_seed(0)
    W = (torch.randn(128, 256) * 0.02).half()
    act = torch.rand(256)
    p = C.pack_linear_weights(W, act, with_qweight=True)
    oq, osc = oracle.eetq_quant_weights(W.t().contiguous().numpy())
    assert np.array_equal(p["qweight"].numpy(), oq) and np.array_equal(p["scales"].numpy().view(np.uint16), osc.view(np.uint16))
    C.save_checkpoint(tmp_path, [{"attention.qkv": p}])
    back = C.load_checkpoint(tmp_path)[0]["attention.qkv"]
    assert torch.equal(back["qweight"].reshape(-1), p["qweight"].reshape(-1)) and torch.equal(back["scales"], p["scales"])
    assert torch.equal(back["W8"].reshape(-1), p["W8"].reshape(-1))


def test_qwen2_shaped_layer_with_qkv_bias_roundtrips(tmp_path):
    """configs[3]: a Qwen2-7B-shaped layer (hidden 3584, qkv 4608 WITH bias, FFN 18944; scaled down 8x here) packed,
    written in the reference's key layout, read back and sharded; the bias travels with the qkv linear."""
    from mixq_tensorrt_llm_b200 import checkpoint as C
    torch.manual_seed(1)
    H, QKV, FFN = 3584 // 8, 4608 // 8, 18944 // 8       # 448, 576, 2368: K % 64 == 0, N % 64 == 0
    act = {k: torch.rand(H) for k in ("qkv", "gate")}
    layer = {
        "attention.qkv": dict(C.pack_linear_weights((torch.randn(QKV, H) * 0.02).half(), act["qkv"], with_qweight=True),
                              bias=(torch.randn(QKV) * 0.1).half()),
        "mlp.gate": C.pack_linear_weights((torch.randn(FFN, H) * 0.02).half(), act["gate"]),
        "mlp.proj": C.pack_linear_weights((torch.randn(H, FFN) * 0.02).half(), torch.rand(FFN)),
    }
    C.save_checkpoint(tmp_path, [layer], config={"architecture": "Qwen2ForCausalLM"})
    back = C.load_checkpoint(tmp_path)[0]
    assert set(back) == set(layer)
    assert torch.equal(back["attention.qkv"]["bias"], layer["attention.qkv"]["bias"])
    assert "bias" not in back["mlp.gate"]
    for lin in layer:
        for k in ("W8", "scale_b", "fp_weight", "ind"):
            assert torch.equal(back[lin][k].reshape(-1), layer[lin][k].reshape(-1)), (lin, k)
    assert back["attention.qkv"]["qweight"].numel() == QKV * H
