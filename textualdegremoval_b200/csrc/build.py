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
SOURCES = ["tdr_runtime.cu", "tdr_conv_gemm.cu", "tdr_mdta.cu", "tdr_pointwise.cu", "tdr_masa.cu", "tdr_vit.cu", "tdr_vit_attn.cu", "tdr_prompt.cu", "tdr_stencil_generic.cu", "tdr_optim.cu", "tdr_backward.cu", "tdr_input.cu", "tdr_gdfn.cu"]
HEADERS = ["tdr_common.cuh", "tdr_stencil.cuh", os.path.join("..", "..", "include", "tdr_sm100.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


OBJ_DIR = os.path.join(HERE, "build")
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math",
          "-Xcompiler", "-fPIC"]


def _hash(files) -> str:
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile_one(src: str, force: bool, verbose: bool):
    """Compile one .cu to build/<name>.o unless its (source + headers + build.py) hash is unchanged."""
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    stamp = _hash([src] + HEADERS + ["build.py"])
    if not force and os.path.exists(obj) and os.path.exists(obj + ".stamp") and open(obj + ".stamp").read() == stamp:
        return obj, None
    cmd = [NVCC] + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(HERE, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        return obj, res.stdout + res.stderr
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    with open(obj + ".stamp", "w") as fh:
        fh.write(stamp)
    return obj, None


def build(force: bool = False, verbose: bool = False) -> str:
    from concurrent.futures import ThreadPoolExecutor
    stamp_file = OUT + ".stamp"
    stamp = _hash(SOURCES + HEADERS + ["build.py"])
    if not force and os.path.exists(OUT) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return OUT
    os.makedirs(OBJ_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(lambda s: _compile_one(s, force, verbose), SOURCES))
    errs = [e for _, e in results if e]
    if errs:
        sys.stderr.write("\n".join(errs))
        raise RuntimeError("nvcc failed building libtdr_sm100.so")
    res = subprocess.run([NVCC, "-shared", "-o", OUT] + [o for o, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("linking libtdr_sm100.so failed")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
