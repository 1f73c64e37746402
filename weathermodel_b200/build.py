"""Build libwm_b200.so (sm_100a only) in-tree with nvcc.

One translation unit (csrc/wm_lib.cu) -> weathermodel_b200/libwm_b200.so. Runs on a CPU-only box:
nvcc cross-compiles for sm_100a. Invoked by __graft_entry__.build() and `python -m weathermodel_b200.build`.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwm_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
]


def _sources():
    out = [os.path.join(HERE, "..", "include", "wm_b200.h")]
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, name))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libwm_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, "wm_lib.cu"), "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libwm_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


def build_variant(name: str, defines, verbose: bool = False) -> str:
    """A/B build of the same sources with extra -D switches -> tools/_diag/libwm_b200_<name>.so (never the product
    library; select it with WM_B200_LIB=... for one measurement)."""
    out_dir = os.path.join(HERE, "..", "tools", "_diag")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.abspath(os.path.join(out_dir, f"libwm_b200_{name}.so"))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, "wm_lib.cu"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed building {out}")
    if verbose:
        sys.stderr.write(res.stderr)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a[2:] for a in sys.argv[i + 2:] if a.startswith("-D")], verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
