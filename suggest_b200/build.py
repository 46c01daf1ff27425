"""Build libsuggest_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsuggest_b200.so")
SOURCES = ["sg_api.cu", "sg_kernels.cu", "sg_bitmap.cu", "sg_exchange.cu", "sg_fine.cu", "sg_long.cu", "sg_lm.cu", "sg_gpubuild.cu", "sg_text.cpp", "sg_index.cpp", "sg_disk.cpp", "sg_hostapi.cpp", "sg_batcher.cpp", "sg_submit.cpp"]
DEPS = SOURCES + ["sg_common.cuh", "sg_device.h", "sg_host.h", "sg_kernels.h", "sg_exchange.h", "unicode_lower.inc", "../../include/suggest_b200.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--fmad=false",
         "-Xcompiler", "-fPIC,-O2,-Wall,-ffp-contract=off", "-shared", "-cudart", "static"]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB


def build_variant(name, defines):
    """Tuning builds (e.g. other TMA ring geometries) next to the product library; select one with SUGGEST_B200_LIB."""
    out = os.path.join(HERE, "variants", f"libsuggest_b200_{name}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [NVCC] + FLAGS + [f"-D{d}" for d in defines] + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
