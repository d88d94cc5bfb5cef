"""Golden vectors for the M <= 4 weight-only branch, from the REFERENCE'S OWN code.

  (a) CPU, dev container:   python tests/golden/make_gemv_golden.py cpu
      runs the reference packer (weightonlykernel/cutlass_kernels/cutlass_preprocessors.cc compiled unmodified into
      oracle/_ref/libref_preprocess.so) -> tests/golden/eetq_layout.npz
  (b) B200, via gpurun:     python tests/golden/make_gemv_golden.py gpu     (writes gpurun_out/golden/ref_gemv_b200.npz)
      runs the reference GEMV kernels (weightOnlyBatchedGemv/*.cu compiled unmodified for sm_100a into
      oracle/_ref/libref_gemv.so) on seeded inputs; copy the file to tests/golden/ and commit.
Inputs are regenerated from the recorded seeds by gemv_case() below; the fixture stores their CRC and the outputs.
"""
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GEMV_CASES = [  # (M, N, K, seed)
    (1, 64, 4096, 11), (2, 64, 4096, 12), (3, 64, 4096, 13), (4, 64, 4096, 14),
    (4, 128, 11008, 15),      # Llama-2-7B down_proj width: partial last pass over the 256 slots
    (2, 64, 1024, 16),        # K < 2048: half of the slots never load
    (1, 192, 3584, 17),       # Qwen2-7B hidden size
]


def gemv_case(M, N, K, seed):
    """Seeded inputs of one case: activations fp16 [M,K] with a few outlier channels, W^T fp16 [K,N]."""
    rng = np.random.default_rng(seed)
    W_t = (rng.standard_normal((K, N)) * 0.02).astype(np.float16)
    A = rng.standard_normal((M, K)).astype(np.float16)
    cols = rng.choice(K, 16, replace=False)
    A[:, cols] *= np.float16(20.0)
    return A, W_t


def crc(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return np.uint32(c)


def main():
    import refgpu
    mode = sys.argv[1] if len(sys.argv) > 1 else "cpu"
    if mode == "cpu":
        rng = np.random.default_rng(5)
        q = rng.integers(-128, 128, (192, 64), dtype=np.int8)
        W_t = (rng.standard_normal((128, 64)) * 0.05).astype(np.float16)
        W_t[7, 3] = np.float16(0.0)
        W_t[:, 9] = np.float16(0.0)             # all-zero channel: scale 0, codes from 0/0
        codes, scales = refgpu.symmetric_quantize_half(W_t)
        np.savez_compressed(ROOT / "tests/golden/eetq_layout.npz", q_kn=q, processed=refgpu.preprocess_int8(q, 80),
                            W_t=W_t, codes=codes, scales=scales, processed_codes=refgpu.preprocess_int8(codes, 80))
        print("wrote tests/golden/eetq_layout.npz")
        return
    import torch
    from oracle import oracle as O
    out = ROOT / "gpurun_out" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    d = {}
    for i, (M, N, K, seed) in enumerate(GEMV_CASES):
        A, W_t = gemv_case(M, N, K, seed)
        qw, sc = O.eetq_quant_weights(W_t)
        ref = refgpu.gemv(torch.from_numpy(A).cuda(), torch.from_numpy(qw).cuda(), torch.from_numpy(sc).cuda())
        torch.cuda.synchronize()
        d[f"out{i}"] = ref.cpu().numpy()
        d[f"crc{i}"] = crc(A, qw, sc)
        mine = O.gemv_w8a16(A, qw, sc)
        print(f"case {i} M={M} N={N} K={K}: oracle vs reference kernel: "
              f"{int((mine.view(np.uint16) != d[f'out{i}'].view(np.uint16)).sum())} of {M * N} differ")
    np.savez_compressed(out / "ref_gemv_b200.npz", **d)
    print("wrote", out / "ref_gemv_b200.npz")


if __name__ == "__main__":
    main()
