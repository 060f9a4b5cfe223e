"""Build libveto_b200.so (the C-ABI library, include/veto_b200.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting
veto_b200/lib/libveto_b200.so travels to the GPU box with the repo snapshot (it is git-ignored).
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libveto_b200.so")

SOURCES = ["api.cu", "pairs.cu", "roi_gather.cu", "box_stage.cu", "tokens.cu", "encoder_ops.cu",
           "gemm_simt.cu", "gemm_tc.cu", "gemm_tc2.cu", "attention_tc.cu", "attention_split.cu", "postprocess.cu", "train.cu", "train_api.cu", "gemm_tn2.cu", "obj_nms.cu", "relsample.cu", "meet_sample.cu", "sgg_eval.cu", "depth_backbone.cu"]
# bit-exact ROIAlign needs un-fused multiply-adds (see roi_gather.cu)
EXTRA = {"roi_gather.cu": ["--fmad=false"]}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libveto_b200.so cannot be built")


def _deps_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    paths.append(os.path.join(HERE, "..", "include", "veto_b200.h"))
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    return not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < _deps_mtime()


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        # VETO_NVCC_DEFINES: extra -D switches for compile-time experiments (e.g. -DVETO_TC2_RES_STAGES=2)
        cmd = [nvcc] + ARCH + COMMON + EXTRA.get(src, []) + os.environ.get("VETO_NVCC_DEFINES", "").split() + \
            ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    r = subprocess.run([nvcc] + ARCH + ["-shared", "-o", tmp] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


def build_library_locked(force: bool = False, verbose: bool = False) -> str:
    """build_library under an exclusive file lock: concurrent processes (the ranks of a torchrun job on a fresh
    checkout) would otherwise share the object files and the temporary library of one another's half-finished build.
    The first process builds, the others wait and find the library up to date."""
    import fcntl
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return build_library(force=force, verbose=verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
