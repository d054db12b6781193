"""In-tree build of libdbnet_b200.so (sm_100a only) with plain nvcc.

`python -m ocr_vi_invoice_b200.build` or `build_library()`; `__graft_entry__.build()` calls the
latter.  The .so is git-ignored but ships to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "libdbnet_b200.so"
SOURCES = ["api.cu", "igemm.cu", "dcn2.cu", "asf_mma.cu", "ref.cu", "elementwise.cu", "postproc.cu", "recprep.cu", "ctc.cu", "svtr.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    deps.append(CSRC.parent.parent / "include" / "dbnet_b200.h")
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = CSRC / (src[:-3] + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if os.environ.get("B200_PROF_BUILD"):  # development: per-role wait counters (B200_DEBUG=16, tools/epi_bench.py)
            cmd.insert(1, "-DB200_PROF")
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src}\n{out}\n")
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libdbnet_b200.so")
    link = [_nvcc(), "-shared", "-o", str(LIB), *objs, "-cudart", "static", "-Xcompiler", "-fPIC"]
    subprocess.run(link, check=True)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
