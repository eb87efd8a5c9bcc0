"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the
header declares, the Python mirror exposes the reference's names, and -- with no GPU here -- the
compute entry points fail loudly instead of falling back to a CPU path."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_are_exported():
    from gnn_builder_b200 import _lib

    lib = _lib.load()
    header = (ROOT / "include" / "gnnb_b200.h").read_text()
    declared = set(re.findall(r"\b(gnnb_[a-z_0-9]+)\s*\(", header))
    declared -= {"gnnb_model_t", "gnnb_model_desc"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gnnb_version() >= 100
    # the tcgen05 probes live in a separate test library with its own header: none of them may
    # leak into the product ABI
    assert not any(n.startswith("gnnb_debug") for n in declared)
    dbg = _lib.load_debug()
    dheader = (ROOT / "include" / "gnnb_b200_debug.h").read_text()
    ddeclared = set(re.findall(r"\b(gnnb_debug_[a-z_0-9]+)\s*\(", dheader))
    assert ddeclared == set(_lib.DEBUG_EXPORTS)
    for name in ddeclared:
        assert hasattr(dbg, name), name
        assert not hasattr(lib, name), f"{name} is exported by the product library"


def test_model_desc_struct_matches_header():
    from gnn_builder_b200 import _lib

    header = (ROOT / "include" / "gnnb_b200.h").read_text()
    body = header[header.index("typedef struct gnnb_model_desc {"):header.index("} gnnb_model_desc;")]
    fields = re.findall(r"^\s*(?:int32_t|float)\s+(\w+)(?:\[\d+\])?;", body, re.M)
    assert fields == [f[0] for f in _lib.ModelDesc._fields_]


def test_activation_ids_match_oracle():
    header = (ROOT / "include" / "gnnb_b200.h").read_text()
    oracle_h = (ROOT / "oracle" / "gnnb_oracle.h").read_text()
    a = dict(re.findall(r"GNNB_ACT_(\w+) = (\d+)", header))
    b = dict(re.findall(r"ORC_ACT_(\w+) = (\d+)", oracle_h))
    assert a == b and len(a) == 13


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gnn_builder_b200 import _lib, layers

    coo = np.array([[0, 1], [1, 0]], np.int32)
    with pytest.raises(_lib.GnnbError):
        layers.compute_degree_tables(coo, 2)
    with pytest.raises(_lib.GnnbError):
        layers.linear(np.ones(4, np.float32), np.ones((2, 4), np.float32), np.zeros(2, np.float32))
    from conftest import model_and_params
    from gnn_builder_b200.engine import Engine

    _, model, _ = model_and_params("c1_gcn_esol_small")
    with pytest.raises(_lib.GnnbError):
        Engine(model)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (a CPU fallback would void parity)."""
    for fp in (ROOT / "gnn_builder_b200").rglob("*"):
        if fp.suffix in (".py", ".cu", ".cuh", ".h") and fp.is_file():
            text = fp.read_text()
            assert "oracle" not in text.lower(), fp


def test_mirror_surface_matches_reference_names():
    import gnn_builder_b200 as gnnb

    for name in ("Project", "GNNModel", "MLP", "GlobalPooling", "GCNConv_GNNB", "GINConv_GNNB",
                 "PNAConv_GNNB", "SAGEConv_GNNB", "compute_average_degree",
                 "compute_average_nodes_and_edges", "compute_max_nodes_and_edges",
                 "compute_median_nodes_and_edges"):
        assert hasattr(gnnb, name), name
    import inspect

    sig = inspect.signature(gnnb.Project.__init__)
    for arg in ("name", "model", "pyg_output_encoding", "vitis_hls_path", "build_dir", "dataset",
                "max_nodes", "max_edges", "num_nodes_guess", "num_edges_guess", "degree_guess",
                "float_or_fixed", "fpx", "clock_speed", "fpga_part", "n_jobs", "cosim_wave_debug"):
        assert arg in sig.parameters, arg
    sig = inspect.signature(gnnb.GNNModel.__init__)
    assert list(sig.parameters)[1:] == [
        "graph_input_feature_dim", "graph_input_edge_dim", "gnn_hidden_dim", "gnn_num_layers",
        "gnn_output_dim", "gnn_conv", "gnn_activation", "gnn_skip_connection", "global_pooling",
        "mlp_head", "output_activation", "gnn_p_in", "gnn_p_hidden", "gnn_p_out"]


def test_parameter_names_match_reference_order():
    """SURVEY appendix B (names verified against the reference's own GNNModel during fixture
    generation: tests/golden/make_golden.py asserts equality with RefModel.param_names)."""
    from conftest import model_and_params

    _, model, params = model_and_params("c2_gin_qm9_small")
    names = list(params)
    assert names[:6] == ["mlp_head_linear_layers_0_weight", "mlp_head_linear_layers_0_bias",
                         "mlp_head_linear_layers_1_weight", "mlp_head_linear_layers_1_bias",
                         "mlp_head_linear_layers_2_weight", "mlp_head_linear_layers_2_bias"]
    assert names[6:10] == ["gnn_convs_0_mlp_linear_0_weight", "gnn_convs_0_mlp_linear_0_bias",
                           "gnn_convs_0_mlp_linear_1_weight", "gnn_convs_0_mlp_linear_1_bias"]
    _, model, params = model_and_params("c1_gcn_esol_small")
    assert list(params)[6:8] == ["gnn_convs_0_conv_bias", "gnn_convs_0_conv_lin_weight"]
    _, model, params = model_and_params("c3_sage_hiv_small")
    assert list(params)[6:9] == ["gnn_convs_0_conv_lin_l_weight", "gnn_convs_0_conv_lin_l_bias",
                                 "gnn_convs_0_conv_lin_r_weight"]
    _, model, params = model_and_params("c4_pna_lipo_small")
    assert list(params)[6:12] == [
        "gnn_convs_0_conv_pre_nns_0_0_weight", "gnn_convs_0_conv_pre_nns_0_0_bias",
        "gnn_convs_0_conv_post_nns_0_0_weight", "gnn_convs_0_conv_post_nns_0_0_bias",
        "gnn_convs_0_conv_lin_weight", "gnn_convs_0_conv_lin_bias"]
    assert params["gnn_convs_0_conv_post_nns_0_0_weight"].shape == (12, 65)


def test_model_validation_errors():
    import torch.nn as nn
    import gnn_builder_b200 as gnnb

    with pytest.raises(ValueError):
        gnnb.GlobalPooling([])
    with pytest.raises(NotImplementedError):
        gnnb.GlobalPooling(["median"])
    with pytest.raises(ValueError):
        gnnb.MLP(4, 2, activation=nn.ELU)
    head = gnnb.MLP(8, 2)
    with pytest.raises(ValueError):
        gnnb.GNNModel(4, None, 8, 2, 8, nn.Linear, nn.ReLU, True, gnnb.GlobalPooling(["add"]),
                      head, None)
    with pytest.raises(ValueError):
        gnnb.GNNModel(4, None, 8, 0, 8, gnnb.GCNConv_GNNB, nn.ReLU, True,
                      gnnb.GlobalPooling(["add"]), head, None)
    with pytest.raises(ValueError):
        gnnb.Project("p", None, "bogus", None, Path("/tmp/x"))


def test_dataset_stats_and_tb_data_roundtrip(tmp_path):
    import gnn_builder_b200 as gnnb
    from gnn_builder_b200.data import in_degree_histogram, read_tb_data, write_tb_data

    b = gnnb.make_molecular_batch(50, 13, 27, 9, seed=3)
    mn, me = gnnb.compute_max_nodes_and_edges(b)
    assert mn == np.diff(b.node_ptr).max() and me == np.diff(b.edge_ptr).max()
    an, ae = gnnb.compute_average_nodes_and_edges(b)
    assert abs(an - 13) <= 2 and abs(ae - 27) <= 4
    assert gnnb.compute_average_degree(b) >= 2
    assert np.array_equal(gnnb.compute_in_deg_histogram(b), in_degree_histogram(b))
    # every node has in-degree >= 1 (PNA's std is NaN otherwise)
    assert gnnb.compute_in_deg_histogram(b)[0] == 0
    params = {"w": np.arange(6, dtype=np.float32).reshape(2, 3)}
    write_tb_data(tmp_path / "tb", params, b.slice(0, 5), golden=np.ones((5, 2), np.float32),
                  out_dim=2)
    p2, b2, g2 = read_tb_data(tmp_path / "tb", 9, {"w": (2, 3)})
    assert np.array_equal(p2["w"], params["w"]) and np.array_equal(g2, np.ones((5, 2)))
    assert np.array_equal(b2.x, b.slice(0, 5).x) and np.array_equal(b2.coo, b.slice(0, 5).coo)


def test_library_is_built_from_the_current_sources():
    """a failed rebuild must not leave an older libgnnb_b200.so behind unnoticed: the in-tree
    library may not be older than csrc/ or the public header (build.py rebuilds it)"""
    from gnn_builder_b200 import build

    if build._stale():
        build.build()          # raises with the compiler output if the sources do not compile
    assert not build._stale()


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """the boundary is a C ABI: include/gnnb_b200.h compiles as strict C99 and a C program links
    against the library (no CUDA device needed for the calls it makes)"""
    import subprocess

    from gnn_builder_b200 import _lib

    _lib.load()
    src = tmp_path / "client.c"
    src.write_text(
        '#include <stdio.h>\n#include "gnnb_b200.h"\n'
        "int main(void)\n{\n    gnnb_model_desc d;\n    int count = -1;\n    (void)d;\n"
        '    printf("version %d\\n", gnnb_version());\n'
        "    if (gnnb_device_count(&count) != GNNB_OK && gnnb_last_error()[0] == 0) return 2;\n"
        "    return 0;\n}\n")
    libdir = Path(_lib.LIB_PATH).parent
    exe = tmp_path / "client"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic",
                        f"-I{ROOT / 'include'}", str(src), f"-L{libdir}", "-lgnnb_b200",
                        f"-Wl,-rpath,{libdir}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("version ")


def test_bench_deadline_guard_prints_the_headline(tmp_path):
    """bench.py's guard around the extras: if they overrun (a hung collective, a dead peer) rank 0
    still prints ONE JSON line -- the headline with the reason under `extra` -- and exits 0"""
    import json
    import subprocess
    import sys

    code = (
        "import sys, time; sys.path.insert(0, %r)\n"
        "import bench\n"
        "line = {'metric': 'graphs_per_sec', 'value': 1.0}\n"
        "extra = {'c4_pna_lipo': {'value': 2.0}}\n"
        "bench.Deadline(0.3, lambda why: dict(line, extra=dict(extra, error=why)), rank=0)\n"
        "time.sleep(30)\n" % str(ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] == 1.0 and d["extra"]["error"] == "deadline exceeded"
    assert d["extra"]["c4_pna_lipo"]["value"] == 2.0
