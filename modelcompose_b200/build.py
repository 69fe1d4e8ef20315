"""In-tree nvcc build of ``libmodelcompose_b200.so`` (sm_100a only, no other arch, no JIT cache).

``python -m modelcompose_b200.build [--force]``.  Objects go to ``build/``, the library to
``modelcompose_b200/_lib/`` (git-ignored, but it travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libmodelcompose_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]

# (source, object suffix, extra defines)
def translation_units():
    tus = []
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith(".cu"):
            continue
        if f == "mc_merge_inst.cu":
            for k in range(7):
                tus.append((f, f"mc_merge_inst_{k}.o", [f"-DMC_PAIR={k}"]))
        elif f == "mc_ties_inst.cu":
            for k in range(3):
                tus.append((f, f"mc_ties_inst_{k}.o", [f"-DMC_TIES_DT={k}"]))
        else:
            tus.append((f, f[:-3] + ".o", []))
    return tus


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    return exe


def _sources_digest() -> str:
    h = hashlib.sha256()
    paths = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "modelcompose_b200.h")]
    for p in paths:
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "sources.sha256")
    digest = _sources_digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    tus = translation_units()

    def compile_one(tu):
        src, obj, defs = tu
        cmd = [nvcc, *NVCC_FLAGS, *defs, "-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, src),
               "-o", os.path.join(OBJ_DIR, obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, tus))
    for src, obj, r in results:
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src} ({obj}):\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr)
    objs = [os.path.join(OBJ_DIR, obj) for _, obj, _ in tus]
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs, "-lcudart_static",
           "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
