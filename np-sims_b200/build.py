"""Build libnpsims_b200.so (sm_100a only) in-tree:  python np-sims_b200/build.py [-j N] [--force]

Each translation unit is compiled with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
and linked into np-sims_b200/lib/libnpsims_b200.so.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libnpsims_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
]

# (source, object suffix, extra defines)
UNITS = (
    [("scan_inst.cu", f"scan_g{g}", [f"-DNPS_WGROUP={g}"]) for g in range(5)]
    + [(f, f[:-3], []) for f in ("merge.cu", "hamming.cu", "replay.cu", "ref_scan.cu", "hist_select.cu", "rerank.cu", "project_f64.cu", "project_tc.cu", "mma_scan.cu", "searcher.cu", "api.cu")]
)


def _newest_header() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    paths.append(os.path.join(HERE, "..", "include", "npsims_b200.h"))
    return max(os.path.getmtime(p) for p in paths)


def _compile(unit, force: bool, verbose: bool):
    src, name, defs = unit
    src_path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, name + ".o")
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) > max(os.path.getmtime(src_path), _newest_header())):
        return obj, False
    if os.environ.get("NPS_DEV"):      # development build: event trace + cycle accounting in mma_scan.cu
        defs = [*defs, "-DNPS_MMA_DEV"]
    if os.environ.get("NPS_KNOBS"):    # host-side development knobs (NPS_MMA_* environment variables) only
        defs = [*defs, "-DNPS_MMA_KNOBS"]
    cmd = [NVCC, *FLAGS, *defs, "-c", src_path, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src} {defs}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj, True


def build(jobs: int | None = None, force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
        results = list(ex.map(lambda u: _compile(u, force, verbose), UNITS))
    objs = [o for o, _ in results]
    if force or any(changed for _, changed in results) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=None)
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.j, a.force, a.verbose))
