"""Builds libfm3d.so (hand-written CUDA for sm_100a + the C-ABI) in-tree with nvcc.

    python -m 3dfeaturematcher_b200.build   (or __graft_entry__.build())

The library carries sm_100a SASS only (-gencode arch=compute_100a,code=sm_100a: tcgen05 / TMEM
instructions are rejected for the generic compute_100 target).  The built .so is git-ignored
but travels with the working tree to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libfm3d.so")
SOURCES = ["fm3d_ctx.cu", "fm3d_triangulate.cu", "fm3d_pyramid.cu", "fm3d_normals.cu", "fm3d_normals_fast.cu",
           "fm3d_patches.cu", "fm3d_describe.cu", "fm3d_describe_kp.cu", "fm3d_describe_brisk.cu", "fm3d_describe_orb.cu", "fm3d_detect.cu", "fm3d_detect_sift.cu", "fm3d_detect_orb.cu", "fm3d_match.cu", "fm3d_probe.cu", "fm3d_comm.cu"]
HEADERS = ["fm3d_internal.cuh", "fm3d_orb_pattern.h", "fm3d_lm2.h", "fm3d_normals_common.cuh", "fm3d_match_pieces.h", os.path.join("..", "..", "include", "fm3d.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libfm3d cannot be built")
    return cand


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + hdrs):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + ccbin + NVCC_FLAGS + os.environ.get("FM3D_EXTRA_NVCC", "").split() + ["-c", src, "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
        return p.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for out in ex.map(compile_one, jobs):
            if verbose:
                sys.stderr.write(out)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc] + ccbin + ["-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                               "-cudart", "static", "-ldl"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
