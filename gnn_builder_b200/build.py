"""Build libgnnb_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libgnnb_b200.so"
DEBUG_LIB = PKG / "libgnnb_b200_debug.so"
SOURCES = ["model.cu", "layers.cu", "tables.cu", "agg.cu", "gemm.cu", "gemm_tc.cu", "pool.cu", "fused.cu", "fused_tc.cu"]
DEBUG_SOURCES = ["tc_test.cu"]   # tcgen05 probes: a separate test library, not in the product
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + (["-DGNNB_TC_SUBTIMING"] if os.environ.get("GNNB_TC_SUBTIMING") else []) \
  + os.environ.get("GNNB_NVCC_EXTRA", "").split()


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale() -> bool:
    if not LIB.exists() or not DEBUG_LIB.exists():
        return True
    t = min(LIB.stat().st_mtime, DEBUG_LIB.stat().st_mtime)
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [
        PKG.parent / "include" / "gnnb_b200.h", PKG.parent / "include" / "gnnb_b200_debug.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    objs = []
    obj_dir = PKG / "build"
    obj_dir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES + DEBUG_SOURCES:
        obj = obj_dir / (src + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                            text=True)))
        objs.append(str(obj))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    (obj_dir / "ptxas.log").write_text("\n".join(log))
    main_objs = objs[:len(SOURCES)]
    dbg_objs = objs[len(SOURCES):]
    for lib, lobjs, extra in (
            (LIB, main_objs, []),
            (DEBUG_LIB, dbg_objs, [f"-L{PKG}", "-l:libgnnb_b200.so", "-Xlinker", "-rpath=$ORIGIN"])):
        cmd = [nvcc(), "-shared", "-o", str(lib), *lobjs, "-gencode",
               "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", *extra]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
