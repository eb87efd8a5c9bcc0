"""Which layout does a 16-bit A operand have in tensor memory?  Runs the one-tile probe
(gnnb_debug_tc_bf16_ts) for the three hypotheses and compares with the fp64 product of the
bf16-truncated operands.  Not part of the test suite: run it on a B200 before building on the result
(`gpurun -- timeout 120 python tools/tmem_bf16_probe.py`)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gnn_builder_b200 import _lib  # noqa: E402


def trunc_bf16(a):
    return (a.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)


lib = _lib.load_debug()
rng = np.random.default_rng(0)
names = ["packed pairs (k even low half, k odd high half)", "one element per cell, low 16 bits",
         "one element per cell, high 16 bits"]
for K, N in ((32, 16), (64, 64), (128, 128)):
    A = rng.uniform(-1, 1, (128, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, K)).astype(np.float32)
    ref = trunc_bf16(A).astype(np.float64) @ trunc_bf16(B).astype(np.float64).T
    for variant, name in enumerate(names):
        out = np.full((128, N), np.nan, np.float32)
        _lib.check(lib.gnnb_debug_tc_bf16_ts(C.c_void_p(A.ctypes.data), C.c_void_p(B.ctypes.data),
                                             C.c_void_p(out.ctypes.data), K, N, variant))
        err = float(np.abs(out - ref).max())
        print(f"K={K:3d} N={N:3d} variant {variant} ({name}): max |err| = {err:.3e} "
              f"{'MATCH' if err < 1e-4 else ''}", flush=True)
