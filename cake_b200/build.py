"""Builds libcake_b200.so in-tree with nvcc for sm_100a (the only target; no fallbacks).

    python -m cake_b200.build            # or: from cake_b200.build import build; build()
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libcake_b200.so")
SRC = os.path.join(HERE, "csrc", "cake_b200.cu")


def _nccl_include() -> str:
    try:
        import nvidia.nccl  # type: ignore
        return os.path.join(os.path.dirname(nvidia.nccl.__file__), "include")
    except Exception:
        for p in sys.path:
            cand = os.path.join(p, "nvidia", "nccl", "include")
            if os.path.exists(os.path.join(cand, "nccl.h")):
                return cand
    return "/usr/include"


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(ROOT, "include", "cake_b200.h")]


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        build_host()
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr",
           "-I", _nccl_include(), "-I", os.path.join(ROOT, "include"),
           "-o", SO, SRC, "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    build_host(force=True)
    return SO


def build_trace(defines=("-DMK_TRACE=1",), name="libcake_b200_trace.so") -> str:
    """A/B builds of the same library with extra -D switches, written next to the product library (selected at run time with
    CAKE_B200_LIB=<path>, bench_tools/ab.sh).  Profiling aid only; nothing in the product uses it."""
    out = os.path.join(HERE, name)
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in sources()):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", *defines,
                           "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr",
                           "-I", _nccl_include(), "-I", os.path.join(ROOT, "include"), "-o", out, SRC, "-ldl"])
    return out


def build_host(force: bool = False) -> str:
    """The C++ host side (cake_b200/host/cake_host.hpp, cake_wire.hpp) + its drivers `cake_run` and `cake_worker`,
    linked against the C ABI."""
    hdrs = [os.path.join(HERE, "host", "cake_host.hpp"), os.path.join(HERE, "host", "cake_wire.hpp"),
            os.path.join(ROOT, "include", "cake_b200.h")]
    for name in ("cake_run", "cake_worker"):
        out = os.path.join(HERE, "host", name)
        src = out + ".cc"
        if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in [src] + hdrs):
            continue
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-o", out, src, "-I", os.path.join(ROOT, "include"),
                               "-L", HERE, "-lcake_b200", "-pthread", "-Wl,-rpath,$ORIGIN/.."])
    return os.path.join(HERE, "host", "cake_run")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
