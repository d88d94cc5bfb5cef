"""In-tree build of libmixq_b200.so (sm_100a only).

    python -m mixq_tensorrt_llm_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so stays in the package directory (git-ignored) so it
travels with the tree to the GPU box.  The arch is passed as
``-gencode arch=compute_100a,code=sm_100a``: the ``-arch=sm_100a`` shorthand also emits a
compute_100 PTX pass which ptxas rejects for tcgen05.mma.kind::i8.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libmixq_b200.so"
SOURCES = ["quant_extract.cu", "gemm_i8_tcgen05.cu", "gemv_w8a16.cu", "allreduce_pull.cu", "mixq_api.cu", "mixq_plugin.cpp", "mixq_registry.cpp"]
HEADERS = ["ptx.cuh", "mixq_internal.h", "mixq_plugin.h", "../../include/mixq_b200.h", "../../include/mixq/trt_shim.h"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; set NVCC=/path/to/nvcc")
    return cand


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS] + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    obj_dir = PKG / "build"
    obj_dir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a wrapper without OpenMP specs; nvcc only needs a host g++
    ccbin = shutil.which("g++") or "g++"
    objs = []
    procs = []
    for s in SOURCES:
        o = obj_dir / (Path(s).stem + ".o")
        cmd = [nvcc, "-ccbin", ccbin, *NVCC_FLAGS, "-x", "cu", "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {s}\n{out}", file=sys.stderr if p.returncode else sys.stdout, flush=True)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [nvcc, "-ccbin", ccbin, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout, file=sys.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
