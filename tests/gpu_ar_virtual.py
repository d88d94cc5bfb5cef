"""Fused GEMM + all-reduce protocol test on ONE GPU: `world` virtual ranks = `world` concurrent launches on separate
streams, each confined to 148/world SMs (mixq_options.sm_limit) so that all of them are co-resident; "peer" pointers are
plain local pointers.  Checks, bit-exactly, that every rank's Out equals fp16(sum over ranks in order of fp32(partial_r))
with partial_r from the unfused kernel, over several launches (counter re-arming) and shapes (edges).
Run as a subprocess by tests/test_gpu_parity.py (a protocol bug traps the context instead of hanging pytest).

    python tests/gpu_ar_virtual.py WORLD [M N K]...
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixq_tensorrt_llm_b200 import binding as B  # noqa: E402


def main():
    world = int(sys.argv[1])
    shapes = [tuple(int(x) for x in s.split("x")) for s in sys.argv[2:]] or [(512, 4096, 1024)]
    lib = B.load()
    B.require_device()
    dev = "cuda"
    nsm = torch.cuda.get_device_properties(0).multi_processor_count
    maxM = max(s[0] for s in shapes)
    maxN = max(s[1] for s in shapes)
    st_bytes = int(lib.mixq_allreduce_staging_size(maxM, maxN, world)) + 1024
    ct_bytes = int(lib.mixq_allreduce_counter_size(maxM, maxN, world))
    outs = [torch.empty(maxM * maxN, dtype=torch.float16, device=dev) for _ in range(world)]
    stag = [torch.empty(st_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    cnts = [torch.zeros(ct_bytes // 4 + 64, dtype=torch.int32, device=dev) for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    g = torch.Generator(device=dev).manual_seed(7)
    lim = (nsm // world) // 2 * 2
    bulk_launches = [0]
    for it, (M, N, K) in enumerate(shapes * 2):
        Kr = K // world
        parts, args = [], []
        for r in range(world):
            A8 = torch.randint(-127, 128, (M, Kr), dtype=torch.int8, device=dev, generator=g)
            W8 = torch.randint(-127, 128, (N, Kr), dtype=torch.int8, device=dev, generator=g)
            sa = (torch.rand(M, device=dev, generator=g) * 0.01 + 1e-3).half()
            sb = (torch.rand(N, device=dev, generator=g) * 0.002 + 1e-4).half()
            fpA = torch.randn(M, 128, device=dev, generator=g).half()
            fw = (torch.randn(N, 128, device=dev, generator=g) * 0.02).half()
            p = torch.empty(M, N, dtype=torch.float16, device=dev)
            B.gemm_dequant(A8, W8, sa, sb, fpA, fw, p, config=9)
            parts.append(p)
            args.append((A8, W8, sa, sb, fpA, fw))
            outs[r].fill_(float("nan"))
        torch.cuda.synchronize()
        ref = parts[0].float()
        for r in range(1, world):
            ref = ref + parts[r].float()
        ref = ref.half()
        # config 9 = the one-kernel path (counters re-armed by the kernel); 0 = what the library picks: small results go GEMM ->
        # pull-reduce kernel (allreduce_pull.cu, epoch-scaled counters) -- the same rank-order fp32 sum either way
        for cfg in (9, 0, 0):
            for r in range(world):
                outs[r].fill_(float("nan"))
            torch.cuda.synchronize()
            for r in range(world):
                grp = B.make_peer_group(world, r, [o.data_ptr() for o in outs], [s.data_ptr() for s in stag],
                                        [c.data_ptr() for c in cnts], st_bytes, ct_bytes)
                with torch.cuda.stream(streams[r]):
                    B.gemm_dequant_allreduce(*args[r], grp, stream=streams[r], sm_limit=lim, config=cfg)
            torch.cuda.synchronize()
            for r in range(world):
                got = outs[r][: M * N].view(M, N)
                same = torch.equal(got.view(torch.int16), ref.view(torch.int16))
                if not same:
                    bad = (got.view(torch.int16) != ref.view(torch.int16))
                    nz = bad.nonzero()
                    print(f"MISMATCH it={it} cfg={cfg} shape={M}x{N}x{K} rank={r}: {int(bad.sum())} of {M*N}; first {nz[:4].tolist()} "
                          f"got {got[bad][:4].tolist()} want {ref[bad][:4].tolist()}")
                    sys.exit(1)
            if cfg == 9:
                bulk_launches[0] += 1
                for c in cnts:   # words: 0 done, 1 launch epoch, 2..3 pushed[parity]
                    assert int(c[0]) == 0 and int(c[2]) == 0 and int(c[3]) == 0 and int(c[1]) == bulk_launches[0] % 2, "counters not re-armed"
            for c in cnts:
                assert int(c[4]) == 0, "a peer wait timed out"
        print(f"ok it={it} world={world} shape={M}x{N}x{K} sm_limit={lim}")
    print("PASS")


if __name__ == "__main__":
    main()
