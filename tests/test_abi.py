"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and the plugin surface (creator lookup, identity strings, format negotiation,
serialisation, workspace sizing, argument validation) behaves like the reference's
(TsinghuaMixQPlugin.cpp).  No kernel is launched here."""
import ctypes
import re
import struct
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_are_exported(lib):
    from mixq_tensorrt_llm_b200 import binding
    hdr = (ROOT / "include" / "mixq_b200.h").read_text()
    declared = set(re.findall(r"\b(mixq_[a-z0-9_]+|initOpenAiTritonPlugins)\s*\(", hdr))
    declared -= {"mixq_status"}
    assert declared == set(binding.EXPORTS), declared ^ set(binding.EXPORTS)
    nm = subprocess.run(["nm", "-D", "--defined-only", str(binding.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in nm.splitlines() if " T " in l}
    missing = declared - exported
    assert not missing, missing
    assert "getPluginRegistry" in exported      # shim build: our own registry stands in for libnvinfer's
    assert b"mixq-b200" in lib.mixq_version()


def test_header_compiles_as_c():
    src = '#include "mixq_b200.h"\nint main(void){ mixq_tensors t; (void)t; return MIXQ_NUM_OUTLIERS == 128 ? 0 : 1; }\n'
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", str(ROOT / "include"),
                        "-x", "c", "-"], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_plugin_identity_and_registry(lib):
    ns = b"tensorrt_llm"
    assert lib.mixq_plugin_create(b"never_registered_ns", 1, 2, 3) is None
    assert lib.initOpenAiTritonPlugins(None, ns) is True
    assert lib.initOpenAiTritonPlugins(None, ns) is True          # idempotent (MixQPlugins.cpp:55-76)
    h = lib.mixq_plugin_create(ns, 512, 12288, 4096)
    assert h
    try:
        assert lib.mixq_plugin_type(h) == b"MixQ"                  # TsinghuaMixQPlugin.cpp:181
        assert lib.mixq_plugin_version(h) == b"1"                  # :180
        assert lib.mixq_plugin_namespace(h) == ns
        assert lib.mixq_plugin_nb_outputs(h) == 1
    finally:
        lib.mixq_plugin_destroy(h)


def test_plugin_format_combination(lib):
    lib.initOpenAiTritonPlugins(None, b"tensorrt_llm")
    h = lib.mixq_plugin_create(b"tensorrt_llm", 1, 1, 1)
    kHALF, kFLOAT, kINT8, kLINEAR, kCHW32 = 1, 0, 2, 0, 5
    try:
        for pos in range(8):                                       # 7 inputs + 1 output, all half/linear (:263-320)
            assert lib.mixq_plugin_supports_format(h, pos, kHALF, kLINEAR) == 1
            assert lib.mixq_plugin_supports_format(h, pos, kFLOAT, kLINEAR) == 0
            assert lib.mixq_plugin_supports_format(h, pos, kINT8, kLINEAR) == 0
            assert lib.mixq_plugin_supports_format(h, pos, kHALF, kCHW32) == 0
        assert lib.mixq_plugin_supports_format(h, 8, kHALF, kLINEAR) == 0
    finally:
        lib.mixq_plugin_destroy(h)


def test_plugin_serialization_roundtrip(lib):
    lib.initOpenAiTritonPlugins(None, b"tensorrt_llm")
    h = lib.mixq_plugin_create(b"tensorrt_llm", 512, 12288, 4096)
    try:
        n = lib.mixq_plugin_serialization_size(h)
        assert n == 12                                             # three int32 (:808-820)
        buf = ctypes.create_string_buffer(n)
        lib.mixq_plugin_serialize(h, buf)
        assert struct.unpack("<iii", buf.raw) == (512, 12288, 4096)
        h2 = lib.mixq_plugin_deserialize(b"tensorrt_llm", buf, n)  # (:227-234, :935-952)
        assert h2
        buf2 = ctypes.create_string_buffer(n)
        lib.mixq_plugin_serialize(h2, buf2)
        assert buf2.raw == buf.raw
        h3 = lib.mixq_plugin_clone(h2)
        buf3 = ctypes.create_string_buffer(n)
        lib.mixq_plugin_serialize(h3, buf3)
        assert buf3.raw == buf.raw and lib.mixq_plugin_namespace(h3) == b"tensorrt_llm"
        lib.mixq_plugin_destroy(h3)
        lib.mixq_plugin_destroy(h2)
        assert lib.mixq_plugin_deserialize(b"tensorrt_llm", buf, 8) is None   # truncated engine blob
    finally:
        lib.mixq_plugin_destroy(h)


def test_workspace_size(lib):
    M, N, K = 512, 12288, 4096
    need = lib.mixq_workspace_size(M, N, K)
    raw = M * K + 2 * M + 256 * M                                  # A8 | scale_a | fp_A  (:406-421)
    sk = lib.mixq_decode_workspace_size(M, N)                      # + split-K scratch: fixed part + fp16 out0 of the decode tiles
    assert lib.mixq_gemm_workspace_size() < sk <= lib.mixq_gemm_workspace_size() + 2 * 512 * 12288 and sk < 48 * 2**20
    assert raw <= need <= raw + 6 * 128                            # the plugin asks TensorRT for exactly what the default path uses
    from mixq_tensorrt_llm_b200 import binding
    lib.mixq_workspace_size_opt.restype = ctypes.c_size_t
    for cfg, extra in ((0, 0), (5, 0), (13, 0), (8, sk), (11, sk), (12, sk)):   # only the opt-in split-K configurations add scratch
        got = lib.mixq_workspace_size_opt(M, N, K, ctypes.byref(binding.Options(cfg, 0)))
        assert need + extra <= got <= need + extra + 128, (cfg, got)
    # non-decreasing in M: the size for a profile's maximum covers every smaller batch
    sizes = [lib.mixq_workspace_size(m, N, K) for m in (1, 32, 512, 1024, 1025, 4096, 65536)]
    assert sizes == sorted(sizes)
    # far below the reference's max(M*K + 2M + 2KN, 16MN) (:342-346), and no int overflow for big M
    assert need < max(M * K + 2 * M + 2 * K * N, 16 * M * N)
    big = lib.mixq_workspace_size(65536, 12288, 11008)
    assert big > 2**29 and big < 2**31 + 2**25
    assert lib.mixq_workspace_size(0, N, K) == 0
    lib.initOpenAiTritonPlugins(None, b"tensorrt_llm")
    h = lib.mixq_plugin_create(b"tensorrt_llm", M, N, K)
    try:
        dims = (ctypes.c_int64 * 3)(32, 16, K)                     # [batch, seq, K] -> M = 512
        assert lib.mixq_plugin_workspace_size(h, dims, 3, N) == need
    finally:
        lib.mixq_plugin_destroy(h)


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected with a status and a message before any CUDA call; nothing throws."""
    from mixq_tensorrt_llm_b200 import binding
    t = binding.Tensors()
    assert lib.mixq_enqueue(None, 1, 8, 16, None, 0, 0, None) == -1
    assert lib.mixq_enqueue(ctypes.byref(t), 8, 8, 16, None, 0, 0, None) == -1          # null tensors
    assert b"null" in lib.mixq_last_error()
    assert lib.mixq_enqueue(ctypes.byref(t), 0, 8, 16, None, 0, 0, None) == 0           # M == 0 is a no-op
    assert lib.mixq_enqueue(ctypes.byref(t), -1, 8, 16, None, 0, 0, None) == -1
    assert lib.mixq_gemm_dequant(1, 1, 1, 1, None, None, 1, 8, 8, 24, None) == -4        # K % 16
    assert lib.mixq_gemm_dequant(16, 16, 16, 16, None, None, 16, 8, 12, 32, None) == -4  # N % 8
    assert lib.mixq_gemm_dequant(16, 16, 16, 16, 16, None, 16, 8, 8, 32, None) == -1     # fp_A without fp_weight
    assert lib.mixq_quant_extract(16, 4, 20, None, 0, 16, 16, None, 0, None) == -1       # K % 8
    from mixq_tensorrt_llm_b200.binding import Options
    bad = Options(99, 0)
    assert lib.mixq_gemm_dequant_opt(16, 16, 16, 16, None, None, 16, 8, 8, 32, None, ctypes.byref(bad), None, 0, None) == -1   # unknown config id
    assert not hasattr(lib, "mixq_set_gemm_config_")  # tuning is per call: no process-wide setters


def test_no_cpu_fallback(lib):
    """Without a B200 the compute entry points fail loudly (status + message), they do not compute."""
    import torch
    from mixq_tensorrt_llm_b200 import binding
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.mixq_device_ok() == 0
    with pytest.raises(binding.MixQError):
        binding.require_device()
    buf = (ctypes.c_uint8 * 4096)()
    p = ctypes.addressof(buf)
    p = (p + 127) // 128 * 128
    rc = lib.mixq_quant_extract(p, 1, 64, None, 0, p + 1024, p + 2048, None, 0, None)
    assert rc == -3 and b"device" in lib.mixq_last_error()
    rc = lib.mixq_gemm_dequant(p, p, p, p, None, None, p, 8, 8, 64, None)
    assert rc == -3
    assert lib.mixq_gemv_w8a16(p, p, p, p, 2, 8, 64, None) == -3                        # the M <= 4 branch likewise
    g = binding.PeerGroup()
    g.world, g.rank = 1, 0
    g.out[0] = g.staging[0] = g.counters[0] = p
    g.staging_bytes, g.counter_bytes = 1 << 20, 4096
    assert lib.mixq_gemm_dequant_allreduce(p, p, p, p, None, None, 8, 8, 64, ctypes.byref(g), None) == -3   # and the fused all-reduce
    # host-buffer calls, synchronous and queued, and the drain
    t = binding.Tensors()
    n = lib.mixq_host_scratch_size(8, 8, 64)
    assert n > 0
    for flags in (0, binding.FLAG_HOST_ASYNC):
        assert lib.mixq_linear_host(ctypes.byref(t), p, p, 8, 8, 64, p, n, flags, None) == -3
    assert lib.mixq_host_drain(None) == -3


def test_product_does_not_touch_oracle():
    """The shipped package must not import, link or execute anything under oracle/."""
    pkg = ROOT / "mixq_tensorrt_llm_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")) + \
            list(pkg.rglob("*.cuh")):
        txt = f.read_text()
        assert "oracle" not in txt.lower() or f.name == "__never__", f
    so = pkg / "libmixq_b200.so"
    if so.exists():
        ldd = subprocess.run(["ldd", str(so)], capture_output=True, text=True).stdout
        assert "oracle" not in ldd and "ref_" not in ldd


def test_allreduce_entry_points_validate_arguments(lib):
    """mixq_enqueue_allreduce / mixq_gemm_dequant_allreduce reject bad groups before touching the GPU."""
    from mixq_tensorrt_llm_b200 import binding
    assert lib.mixq_allreduce_counter_size(512, 8192, 8) >= 16
    # one 256x256 tile slot per tile, rounded up to a multiple of world
    assert lib.mixq_allreduce_staging_size(512, 8192, 8) == 512 * 8192 * 2
    assert lib.mixq_allreduce_staging_size(300, 1000, 2) == 2 * 4 * 256 * 256 * 2
    assert lib.mixq_allreduce_staging_size(0, 8, 2) == 0
    g = binding.PeerGroup()
    g.world, g.rank = 9, 0
    assert lib.mixq_gemm_dequant_allreduce(16, 16, 16, 16, None, None, 8, 8, 16, ctypes.byref(g), None) == -1   # world > 8
    g.world, g.rank = 2, 2
    assert lib.mixq_gemm_dequant_allreduce(16, 16, 16, 16, None, None, 8, 8, 16, ctypes.byref(g), None) == -1   # rank >= world
    assert lib.mixq_gemm_dequant_allreduce(16, 16, 16, 16, None, None, 8, 8, 16, None, None) == -1              # no group
    t = binding.Tensors()
    assert lib.mixq_enqueue_allreduce(ctypes.byref(t), 8, 8, 16, None, 0, ctypes.byref(g), 0, None) == -1       # null tensors
    assert lib.mixq_enqueue_allreduce(ctypes.byref(t), 0, 8, 16, None, 0, ctypes.byref(g), 0, None) == 0        # M == 0: no-op


def test_ctypes_structs_match_the_header(tmp_path):
    """binding.Tensors / PeerGroup / Epilogue must have the size and field offsets of the C structs in
    include/mixq_b200.h (compiled here with gcc and printed)."""
    from mixq_tensorrt_llm_b200 import binding
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "mixq_b200.h"
int main(void) {
    printf("%zu %zu %zu %zu\n", sizeof(mixq_tensors), offsetof(mixq_tensors, q_weight), offsetof(mixq_tensors, Out), (size_t)MIXQ_MAX_RANKS);
    printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(mixq_peer_group), offsetof(mixq_peer_group, out), offsetof(mixq_peer_group, staging),
           offsetof(mixq_peer_group, counters), offsetof(mixq_peer_group, staging_bytes), offsetof(mixq_peer_group, counter_bytes),
           offsetof(mixq_peer_group, out_multicast));
    printf("%zu %zu\n", sizeof(mixq_epilogue), offsetof(mixq_epilogue, activation));
    return 0;
}'''
    c = tmp_path / "s.c"
    c.write_text(src)
    exe = tmp_path / "s"
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", str(ROOT / "include"), str(c), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b, e = (list(map(int, ln.split())) for ln in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    T, P, E = binding.Tensors, binding.PeerGroup, binding.Epilogue
    assert a == [ctypes.sizeof(T), T.q_weight.offset, T.Out.offset, binding.MAX_RANKS]
    assert b == [ctypes.sizeof(P), P.out.offset, P.staging.offset, P.counters.offset, P.staging_bytes.offset, P.counter_bytes.offset,
                 P.out_multicast.offset]
    assert e == [ctypes.sizeof(E), E.activation.offset]


def test_c_program_links_and_calls_the_abi(tmp_path, lib):
    """A plain C translation unit includes the header, links libmixq_b200.so and calls the host-only entry points:
    the boundary really is a C ABI (no C++ / torch types), and failures come back as status codes, not exceptions."""
    from mixq_tensorrt_llm_b200 import binding
    src = r'''
#include <stdio.h>
#include <string.h>
#include "mixq_b200.h"
int main(void) {
    mixq_tensors t; memset(&t, 0, sizeof t);
    mixq_peer_group g; memset(&g, 0, sizeof g);
    mixq_epilogue e = {0, MIXQ_ACT_SILU};
    printf("%s\n", mixq_version());
    printf("%zu %zu %zu\n", mixq_workspace_size(512, 12288, 4096), mixq_allreduce_staging_size(512, 8192, 8), mixq_gemm_workspace_size());
    printf("%d\n", mixq_enqueue(0, 8, 8, 16, 0, 0, 0, 0));                       /* null table  -> MIXQ_ERR_BAD_ARG */
    printf("%d\n", mixq_enqueue(&t, 8, 8, 16, 0, 0, 0, 0));                      /* null tensor -> MIXQ_ERR_BAD_ARG */
    printf("%d\n", mixq_enqueue_ex(&t, 0, 8, 16, 0, 0, &e, 0, 0));               /* M == 0      -> MIXQ_OK          */
    printf("%d\n", mixq_enqueue_allreduce(&t, 8, 8, 16, 0, 0, &g, 0, 0));        /* null tensor -> MIXQ_ERR_BAD_ARG */
    printf("%d\n", mixq_gemv_w8a16(0, 0, 0, 0, 2, 8, 64, 0));                    /* null ptr    -> MIXQ_ERR_BAD_ARG */
    printf("%d\n", (int)initOpenAiTritonPlugins(0, "tensorrt_llm"));
    mixq_plugin_t* p = mixq_plugin_create("tensorrt_llm", 1, 2, 3);
    printf("%s %s %d\n", mixq_plugin_type(p), mixq_plugin_version(p), mixq_plugin_nb_outputs(p));
    mixq_plugin_destroy(p);
    printf("[%s]\n", strlen(mixq_last_error()) ? "has-error-text" : "");
    return 0;
}'''
    c = tmp_path / "abi.c"
    c.write_text(src)
    exe = tmp_path / "abi"
    libdir = binding.LIB_PATH.parent
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(c), "-o", str(exe),
                        "-L", str(libdir), "-lmixq_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    ln = out.stdout.splitlines()
    assert ln[0].startswith("mixq-b200")
    ws, st, gw = map(int, ln[1].split())
    assert ws >= 512 * 4096 + 2 * 512 + 256 * 512 and st == 512 * 8192 * 2 and gw > 0
    assert [int(x) for x in ln[2:7]] == [-1, -1, 0, -1, -1]
    assert ln[7] == "1" and ln[8] == "MixQ 1 1" and ln[9] == "[has-error-text]"


def test_fat_tile_plan_invariants(lib):
    """mixq_debug_fat_plan (host arithmetic of the decode fat-tile schedule, csrc/gemm_i8_tcgen05.cu plan_fat) over the
    benchmarked shapes and random ones: the tile fits TMEM (int32 accumulators + parked fp16 outlier product), the UMMA N
    granularity holds, the column tiles cover N, the ring has at least three stages and fits 227 KB next to the epilogue
    staging, and the wave count is what the tile count says."""
    import random
    rng = random.Random(5)
    shapes = [(512, 12288, 0), (512, 11008, 1), (512, 4096, 0), (512, 18944, 1), (512, 4608, 0), (512, 1280, 0), (512, 3584, 1),
              (1024, 28672, 1), (129, 16, 0), (300, 8, 1)]
    shapes += [(rng.randint(129, 1024), 8 * rng.randint(1, 4096), rng.randint(0, 1)) for _ in range(300)]
    out = (ctypes.c_int * 5)()
    for M, N, gated in shapes:
        for pairs in (74, 37, 20, 1):
            for ew in (8, 12):
                assert lib.mixq_debug_fat_plan(M, N, pairs, gated, ew, out) == 0
                Nt, n_tiles, m_tiles, stages, waves = list(out)
                cols = 2 * N if gated else N                      # accumulator columns of the whole problem
                assert Nt % (32 if gated else 16) == 0 and 16 <= Nt <= 336
                assert Nt + Nt // 2 <= 512
                assert n_tiles * Nt >= cols and (n_tiles - 1) * Nt < cols
                assert m_tiles == (M + 255) // 256
                assert waves == -(-(m_tiles * n_tiles) // pairs)
                stage_bytes = 128 * 128 + (Nt // 2) * 128
                fixed = 3072 + 512 + ew * 2 * 2048
                assert stages >= 3 and stages <= 8
                assert 1024 + stages * stage_bytes + fixed <= 227 * 1024
    assert lib.mixq_debug_fat_plan(512, 4096, 74, 0, 10, out) == -1     # 8 or 12 epilogue warps only
    assert lib.mixq_debug_fat_plan(512, 4096, 0, 0, 8, out) == -1


def test_host_scratch_parts(lib):
    """Queued host-buffer calls (MIXQ_FLAG_HOST_ASYNC): the parts of the scratch that consecutive calls take never run past
    its end, do not overlap for calls of one size, are 128-byte aligned, and there are at most four of them."""
    import random
    rng = random.Random(9)
    for _ in range(400):
        need = rng.randint(1, 1 << 26)
        k_given = rng.choice([1.0, 1.5, 2.0, 2.7, 3.0, 4.0, 7.5])
        total = int(need * k_given) + rng.randint(0, 4096)
        offs = [lib.mixq_debug_host_part_offset(need, total, i) for i in range(9)]
        aligned_need = (need + 127) // 128 * 128
        k = min(4, total // aligned_need)
        distinct = sorted(set(offs))
        assert len(distinct) == (k if k >= 2 else 1)
        assert offs[:len(distinct)] == distinct and offs[len(distinct)] == offs[0]        # taken in turn
        for o in distinct:
            assert o % 128 == 0 and o + 127 + need <= total + 127                        # (+ the alignment slack `need` includes)
        for o1, o2 in zip(distinct, distinct[1:]):
            assert o2 - o1 >= aligned_need
    assert lib.mixq_debug_host_part_offset(100, 50, 0) == -1
