"""Builds libqadc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "qadc_capi.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("qadc_capi.cu", "qadc_device.cuh", "qadc_scan.cuh", "qadc_tables.cuh", "qadc_adc.cuh", "qadc_multi.cuh", "qadc_flatprep.cuh")]
DEPS.append(os.path.join(os.path.dirname(HERE), "include", "qadc_b200.h"))
OUT = os.path.join(HERE, "libqadc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    extra = os.environ.get("QADC_NVCC_EXTRA", "").split()
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC, "-ldl"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
