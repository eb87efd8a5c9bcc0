import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (C restatement of the reference) -- the checker, never the product."""
    from oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def lib_tb():
    """The reference's own known-answer vectors (gnn_builder_lib_test/tb_data/*.bin)."""
    d = GOLDEN / "lib_tb"

    class TB:
        n = int(np.fromfile(d / "tb_num_nodes.bin", np.int32)[0])
        e = int(np.fromfile(d / "tb_num_edges.bin", np.int32)[0])

        def f32(self, name, *shape):
            a = np.fromfile(d / f"{name}.bin", np.float32)
            return a.reshape(shape) if shape else a

        def i32(self, name, *shape):
            a = np.fromfile(d / f"{name}.bin", np.int32)
            return a.reshape(shape) if shape else a

    tb = TB()
    tb.coo = tb.i32("tb_coo_matrix", tb.e, 2)
    tb.x = tb.f32("tb_input_node_features", tb.n, 8)
    tb.in_deg = tb.i32("tb_in_degree_table")
    tb.out_deg = tb.i32("tb_out_degree_table")
    tb.offsets = tb.i32("tb_neighbor_table_offsets")
    tb.nbr = tb.i32("tb_neighbor_table")
    tb.eidx = tb.i32("tb_edge_index_table")
    return tb


def load_model_golden(name):
    z = np.load(GOLDEN / "models" / f"{name}.npz")
    from gnn_builder_b200.data import GraphBatch

    batch = GraphBatch(z["x"], z["coo"], z["node_ptr"], z["edge_ptr"])
    params = {k[len("param__"):]: z[k] for k in z.files if k.startswith("param__")}
    return batch, z["out"], params, float(z["checksum"])


def workload_by_name(name):
    from gnn_builder_b200.configs import WORKLOADS
    import dataclasses

    if name.endswith("_small"):
        w = WORKLOADS[name[: -len("_small")]]
        return dataclasses.replace(w, name=name, hidden_dim=12, in_dim=5, mlp_hidden_dim=8,
                                   max_nodes=64, max_edges=256)
    return WORKLOADS[name]


def model_and_params(name):
    """(GNNModel mirror, params dict in flat reference order) with seed-0 weights."""
    from gnn_builder_b200.models import build_model

    w = workload_by_name(name)
    model = build_model(w, pna_delta=w.pna_delta, seed=0)
    return w, model, model.named_parameter_arrays()


def rel_err(a, b):
    """max_abs(delta) / max(1, max_abs(ref))."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max())) if a.size else 0.0


MODEL_NAMES = ["c1_gcn_esol", "c2_gin_qm9", "c3_sage_hiv", "c4_pna_lipo"]
