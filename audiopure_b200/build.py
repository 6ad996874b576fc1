"""Builds libaudiopure_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

    python -m audiopure_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the tree.
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libaudiopure_b200.so")
SOURCES = ["capi.cu"]
DEPS = ["capi.cu", "sm100.cuh", "diffwave_kernels.cuh", "mel_kernel.cuh", os.path.join("..", "..", "include", "audiopure_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--compiler-options", "-fPIC",
    "-shared",
]
LINK = ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC=/path/to/nvcc")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False, ablation=False, out=None, defines=()):
    """Compile if missing or older than its sources. Returns the library path.
    ablation=True adds -DAP_ENABLE_ABLATION (the AP_DEBUG switches used for profiles/r0*_ablation.md).
    out / defines: build a VARIANT library next to the product one (A/B experiments, selected with AP_LIB=...)."""
    lib = os.path.join(HERE, out) if out else LIB
    if not force and not out and not stale():
        return LIB
    cmd = ([_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DAP_ENABLE_ABLATION"] if ablation else [])
           + ["-D" + d for d in defines] + ["-o", lib] + SOURCES + LINK)
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr))
    if verbose:
        print(proc.stderr)
    return lib


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv or "--ablation" in sys.argv, verbose="-v" in sys.argv,
                ablation="--ablation" in sys.argv, out=_out, defines=[a[2:] for a in sys.argv if a.startswith("-D")]))
