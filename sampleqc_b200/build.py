"""In-tree build of libsampleqc_cuda.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "sqc_lib.cu")
OUT = os.environ.get("SQC_BUILD_OUT") or os.path.join(_HERE, "libsampleqc_cuda.so")      # SQC_BUILD_OUT: development variants
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xcompiler", "-pthread"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "sampleqc_cuda.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, dims: str | None = None, verbose: bool = False) -> str:
    """dims: e.g. "3,6" restricts the template instantiations (development builds)."""
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS
    if dims:
        cmd.append("-DSQC_FOR_D(X)=" + " ".join(f"X({int(d)})" for d in dims.split(",")))
    if os.environ.get("SQC_EXTRA_NVCC"):
        cmd += os.environ["SQC_EXTRA_NVCC"].split()
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, dims=os.environ.get("SQC_DIMS"), verbose="-v" in sys.argv)
    print(OUT)
