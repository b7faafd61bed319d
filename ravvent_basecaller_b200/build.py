"""In-tree build of libravvent_b200.so (nvcc, sm_100a only).

    python ravvent_basecaller_b200/build.py [--force]     (or __graft_entry__.build())

The library is linked against the static CUDA runtime, so it loads on a machine
without a GPU (compute calls then fail with RVB_ERR_CUDA - there is no CPU path).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "libravvent_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr"]
# per-file extra flags: the event scan must not contract a*b+c into FMA (bit-exact float64 contract)
SOURCES = {
    "api.cu": [],
    "event_detect.cu": ["-fmad=false"],
    "lstm_recurrent_tc.cu": [],
    "proj_gemm.cu": [],
    "decoder_wave.cu": [],
    "attention_tc.cu": [],
    "snippets.cu": ["-fmad=false"],
    "merger.cu": ["-fmad=false"],        # float64 alignment scores are compared exactly: no a*b+c contraction
}


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    headers = list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "ravvent_b200.h", Path(__file__)]
    bdir = PKG / "build"
    bdir.mkdir(exist_ok=True)
    jobs = []
    for src, extra in SOURCES.items():
        obj = bdir / (src + ".o")
        if force or _stale(obj, [CSRC / src] + headers):
            jobs.append([NVCC, *ARCH, *COMMON, *extra, "-c", str(CSRC / src), "-o", str(obj)])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)

    with ThreadPoolExecutor(max_workers=min(6, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [str(bdir / (s + ".o")) for s in SOURCES]
    if force or jobs or _stale(OUT, objs):
        run([NVCC, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", str(OUT), *objs, "-cudart", "static"])
    return OUT


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose=True)
    print(p)
