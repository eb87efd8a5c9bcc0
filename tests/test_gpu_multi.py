"""Multi-GPU parity of the row-partitioned large-graph path: the halo exchange (both transports)
under torchrun on every GPU of the box, checked against the CPU oracle by
tests/run_large_multi_gpu.py.  Skips on a box with fewer than two GPUs."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run(world: int, transport: str, nodes: int, env=None):
    port = 29500 + (os.getpid() + 7 * world + len(transport)) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(ROOT / "tests" / "run_large_multi_gpu.py"), "--nodes", str(nodes),
           "--transport", transport]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT),
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, r.stdout[-2000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_halo_exchange_matches_the_oracle(transport):
    import torch

    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs at least two GPUs")
    world = min(n_gpus, 4)
    res = _run(world, transport, 80000)
    assert res["ok"], res
    assert res["stats"]["transport"] == transport
    assert res["err_out"] < 1e-4 and res["err_emb"] < 1e-4
    assert res["bit_identical_rerun"]
    # a true halo: fewer rows than the full remote set travel
    assert 0 < res["stats"]["halo_rows"] <= (80000 - 80000 % world) * (world - 1) // world


def test_halo_exchange_pipelined_row_blocks():
    """the opt-in block mode (GNNB_HALO_BLOCKS): a layer's rows computed in blocks, every finished
    block's rows pushed to the peers while the next block runs -- same results"""
    import torch

    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs at least two GPUs")
    res = _run(min(n_gpus, 4), "p2p", 80000, env={"GNNB_HALO_BLOCKS": "3"})
    assert res["ok"], res
    assert res["stats"]["send_blocks"] == 3 and res["stats"]["transport"] == "p2p"
    assert res["err_out"] < 1e-4 and res["err_emb"] < 1e-4 and res["bit_identical_rerun"]
