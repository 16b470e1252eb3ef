"""Build libpdwt_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libpdwt_b200.so")
SOURCES = ["pdwt_generic.cu", "pdwt_elementwise.cu", "pdwt_fused.cu", "pdwt_stream.cu", "pdwt_swt.cu", "pdwt_nonsep.cu", "pdwt_nonsep_swt.cu", "pdwt_rows.cu", "pdwt_rows_all.cu", "pdwt_capi.cu", "pdwt_wavelets.cu", "pdwt_sharded.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
              "-fPIC,-O2,-Wall", "-Xptxas", "-v"] + (["-DPDWT_EXPERIMENTS"] if os.environ.get("PDWT_EXPERIMENTS") else [])


PER_FILE_FLAGS = {}   # extra nvcc flags for single sources (none at present)


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d += [os.path.join(HERE, "..", "include", f) for f in ("pdwt_b200.h", "wt.h")]
    return d


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    newest_dep = max(os.path.getmtime(p) for p in _deps())
    objs, jobs = [], []
    for src in SOURCES:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < newest_dep:
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = ["nvcc", *NVCC_FLAGS, *PER_FILE_FLAGS.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr[-4000:]}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(LIB):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
