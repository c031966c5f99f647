"""Build libtdr_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m textualdegremoval_b200.csrc.build [--force] [--verbose]

The .so lands next to the package (textualdegremoval_b200/libtdr_sm100.so) so that it travels with the repo
snapshot to the GPU box; it is git-ignored.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libtdr_sm100.so")
SOURCES = ["tdr_runtime.cu", "tdr_conv_gemm.cu", "tdr_mdta.cu", "tdr_pointwise.cu", "tdr_masa.cu", "tdr_vit.cu", "tdr_optim.cu"]
HEADERS = ["tdr_common.cuh", os.path.join("..", "..", "include", "tdr_sm100.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math",
         "-Xcompiler", "-fPIC", "-shared"]


def _stamp() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS + ["build.py"]:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp_file = OUT + ".stamp"
    stamp = _stamp()
    if not force and os.path.exists(OUT) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(HERE, s) for s in SOURCES] + ["-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libtdr_sm100.so")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
