"""Builds librlg_b200.so (the C-ABI library, include/rlg_b200.h) in-tree with nvcc for sm_100a.

    python "gan-rl_3d_b200/build.py" [--force] [--verbose] [--experiments]

The .so lands in gan-rl_3d_b200/lib/ (git-ignored, travels to the GPU box with the snapshot).  nvcc
cross-compiles without a GPU.  No torch C++ extension, no pybind: the boundary is plain C.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(LIB_DIR, "librlg_b200.so")
EXP_LIB_PATH = os.path.join(LIB_DIR, "librlg_b200_exp.so")     # -DRLG_EXPERIMENTS: timing variants for tools/, never shipped
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", INCLUDE_DIR,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_input() -> float:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE_DIR, "rlg_b200.h"), __file__]
    return max(os.path.getmtime(p) for p in deps)


def is_fresh(experiments: bool = False) -> bool:
    path = EXP_LIB_PATH if experiments else LIB_PATH
    return os.path.exists(path) and os.path.getmtime(path) >= _newest_input()


def build_variant(tag: str, defines) -> str:
    """tools/ only: an experiments build with extra -D defines -> lib/librlg_b200_exp_<tag>.so (A/B timing)."""
    nvcc = _nvcc()
    obj_dir = OBJ_DIR + "_exp_" + tag
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-DRLG_EXPERIMENTS"] + [f"-D{d}" for d in defines] + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        objs.append(obj)
    out = os.path.join(LIB_DIR, f"librlg_b200_exp_{tag}.so")
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcuda"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out


def build_library(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link librlg_b200.so.  Returns the library path.
    experiments=True builds librlg_b200_exp.so with -DRLG_EXPERIMENTS instead: knock-out / A-B kernel variants behind
    extra flag bits, used only by tools/ (the product library rejects those bits)."""
    lib_path = EXP_LIB_PATH if experiments else LIB_PATH
    if not force and is_fresh(experiments):
        return lib_path
    nvcc = _nvcc()
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = OBJ_DIR + ("_exp" if experiments else "")
    os.makedirs(obj_dir, exist_ok=True)
    log_lines = []
    extra = ["-DRLG_EXPERIMENTS"] if experiments else []

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log_lines.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = lib_path + ".tmp"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs + ["-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log_lines.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, lib_path)
    with open(os.path.join(obj_dir, "build.log"), "w") as f:
        f.write("\n".join(log_lines))
    if verbose:
        print("\n".join(log_lines))
    return lib_path


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                         experiments="--experiments" in sys.argv)
    print(path)
