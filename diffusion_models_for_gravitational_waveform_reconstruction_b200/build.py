"""Build the C-ABI CUDA library in-tree (nvcc, sm_100a only).  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libgwb200.so")
LIB_FFT = os.path.join(PKG, "libgwb200_fft.so")
SOURCES = ["forward.cu", "conv_tc.cu", "backward.cu", "optim.cu", "wgrad_tc.cu", "stream_gn.cu", "score.cu", "conv_in_gn.cu",
           "gn_bwd_fused.cu", "conv_in_direct.cu", "generic.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(LIB_FFT):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "gwb200.h"),
                                                              os.path.join(PKG, "..", "include", "gwb200_fft.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    flags += os.environ.get("GW_NVCC_EXTRA", "").split()          # experiments only (e.g. -DCGN_ABLATE)
    objs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(PKG, "build", src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    # whitening / sigma library: separate .so so that the core library has no cuFFT dependency
    cmd = [nvcc, *flags, "-shared", "-o", LIB_FFT, os.path.join(CSRC, "whiten.cu"), "-lcufft", "-Xlinker",
           "-rpath=/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"building {LIB_FFT} failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
