"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe that vendors the UNMODIFIED reference sources of the hot path into
``oracle/_ref/`` (git-ignored, NOT gpurun-ignored: it travels to the GPU box like the built .so, it never enters history).

    python -m oracle.build_ref            (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python, so "building" it is a file copy: the arch files, the DINOv2 package and the script that
holds the Mapper / CleanMapper classes, byte for byte, under the same relative paths.  ``oracle.ref_loader`` falls back to
this tree when /root/reference is absent, which lets ``bench.py --impl reference`` time the reference's OWN modules on
the GPU box's host cores (``cpu_baseline.kind == "reference"``) instead of the oracle port.
Nothing in the product imports it.
"""
import hashlib
import os
import shutil

SRC = os.environ.get("TDR_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = [
    "models/archs/network_restormer_guided_arch.py",
    "models/archs/network_nafnet_guided_arch.py",
    "models/archs/network_promptir_guided_arch.py",
    "models/archs/network_drsformer_guided_arch.py",
    "models/archs/network_drsformer_guided_arch_200L_SPA.py",
    "models/archs/nafnet_arch_utils.py",
    "models/archs/nafnet_local_arch.py",
    "scripts/train/main_train_tr_mapping.py",
]
DIRS = ["models/dino"]


def build(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "models", "archs")):
        if verbose:
            print(f"oracle.build_ref: {SRC} not present, nothing to do")
        return None
    files = list(FILES)
    for d in DIRS:
        files += [os.path.join(d, f) for f in sorted(os.listdir(os.path.join(SRC, d))) if f.endswith(".py")]
    manifest = []
    for rel in files:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {rel}")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    if verbose:
        print(f"oracle.build_ref: {len(files)} reference files -> {DST}")
    return DST


if __name__ == "__main__":
    build()
