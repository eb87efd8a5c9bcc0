"""Print issue / completion cycles per tcgen05.mma for the flavours used by the fused kernel."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gnn_builder_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
names = ["tf32 SS", "tf32 TS (A in TMEM)", "bf16 SS K-major", "bf16 SS, B MN-major"]
for flavour, name in [(f, n) for f, n in enumerate(names)] + [
        (20 + f, n + " [lean warp-uniform issue]") for f, n in enumerate(names)] + [
        (1020 + f, n + " [lean, M = 64]") for f, n in enumerate(names)]:
    for N in (32, 64, 128, 256):
        if flavour % 10 == 3 and N == 256:   # (MN-major B: 128 columns staged)
            continue
        for reps in (512,):
            cyc = np.zeros(2, np.int64)
            _lib.check(lib.gnnb_debug_tc_mma_rate(flavour, N, reps, C.c_void_p(cyc.ctypes.data)))
            print(f"{name:44s} N={N:3d} reps={reps:4d}: issue {cyc[0] / reps:7.1f} cyc/MMA, "
                  f"complete {cyc[1] / reps:7.1f} cyc/MMA", flush=True)
