#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- build the checker.

Two products, both plain ``gcc``/``g++`` on a handful of files (no reference build system):

1. ``oracle/_ref/libgnnb_oracle.so``  -- the C restatement (oracle/gnnb_oracle.c).  Built
   anywhere (here and on the GPU box).
2. ``oracle/_ref/...`` -- the REFERENCE ITSELF, compiled from the sources where they lie under
   /root/reference (only in the container that has them; the GPU box uses the prebuilt files):
     * ``lib_test``                : the reference's 21-test unit testbench (test.cpp)
     * ``libgnnb_ref_layers.so``   : oracle/ref_layers.cpp = extern "C" instantiations of the
                                     reference's templates
     * ``models/<name>/``          : model.h / model.cpp / model_tb.cpp rendered by the
                                     reference's OWN ``gnnbuilder.Project`` (running under the
                                     fake torch_geometric of oracle/fake_pyg.py), compiled to
                                     ``lib<name>.so`` (exports ``<name>_top``) and ``result``
                                     (the reference testbench main), plus ``manifest.json``.

Flags mirror gnnbuilder/templates/makefile_testbench.jinja:22-24: -fPIC -O3 -std=c++14, no
-ffast-math, no -march=native (so no FMA contraction).  ``oracle/_ref/`` is git-ignored.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
REF = Path("/root/reference")
OUT = HERE / "_ref"
CXXFLAGS = ["-O3", "-std=c++14", "-fPIC", "-Wno-unused-result", "-ffp-contract=off"]


def _run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError(f"command failed: {' '.join(map(str, cmd))}\n{r.stdout}\n{r.stderr}")
    return r


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).exists() and Path(s).stat().st_mtime > t for s in sources)


def have_reference() -> bool:
    return (REF / "gnnbuilder" / "gnn_builder_lib" / "gnn_builder_lib.h").exists()


def build_oracle(force: bool = False) -> Path:
    OUT.mkdir(parents=True, exist_ok=True)
    so = OUT / "libgnnb_oracle.so"
    srcs = [HERE / "gnnb_oracle.c", HERE / "gnnb_oracle.h"]
    if force or _stale(so, srcs):
        _run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off",
              str(HERE / "gnnb_oracle.c"), "-o", str(so), "-lm"])
    return so


def build_lib_test(force: bool = False) -> Path:
    exe = OUT / "lib_test"
    test_dir = REF / "gnnbuilder" / "gnn_builder_lib_test"
    if force or _stale(exe, [test_dir / "test.cpp"]):
        _run(["g++", *CXXFLAGS, "-I", str(HERE / "shim"), str(test_dir / "test.cpp"),
              "-o", str(exe)])
    return exe


def run_lib_test() -> str:
    """Run the reference's own unit testbench against its own tb_data (cwd = its directory)."""
    exe = build_lib_test()
    r = _run([str(exe)], cwd=str(REF / "gnnbuilder" / "gnn_builder_lib_test"))
    return r.stdout


def build_ref_layers(force: bool = False) -> Path:
    so = OUT / "libgnnb_ref_layers.so"
    if force or _stale(so, [HERE / "ref_layers.cpp"]):
        _run(["g++", *CXXFLAGS, "-shared", "-I", str(HERE / "shim"),
              "-I", str(REF / "gnnbuilder" / "gnn_builder_lib_test"),
              "-mcmodel=medium", str(HERE / "ref_layers.cpp"), "-o", str(so)])
    return so


def _import_reference():
    sys.path.insert(0, str(HERE))
    import fake_pyg

    fake_pyg.install()
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    import gnnbuilder  # the unmodified reference package

    return gnnbuilder


def reference_model(workload, pna_delta: float = 1.0):
    """Instantiate the reference's own GNNModel for a configs.Workload."""
    import torch.nn as nn

    gnnb = _import_reference()
    conv = {"gcn": gnnb.GCNConv_GNNB, "gin": gnnb.GINConv_GNNB, "sage": gnnb.SAGEConv_GNNB,
            "pna": gnnb.PNAConv_GNNB}[workload.conv]
    act = {"relu": nn.ReLU, "gelu": nn.GELU, "sigmoid": nn.Sigmoid, "tanh": nn.Tanh}[
        workload.activation]
    model = gnnb.GNNModel(
        graph_input_feature_dim=workload.in_dim,
        graph_input_edge_dim=0,
        gnn_hidden_dim=workload.hidden_dim,
        gnn_num_layers=workload.num_layers,
        gnn_output_dim=workload.gnn_output_dim,
        gnn_conv=conv,
        gnn_activation=act,
        gnn_skip_connection=workload.skip,
        global_pooling=gnnb.GlobalPooling(list(workload.pools)),
        mlp_head=gnnb.MLP(in_dim=workload.gnn_output_dim * len(workload.pools),
                          out_dim=workload.out_dim, hidden_dim=workload.mlp_hidden_dim,
                          hidden_layers=workload.mlp_hidden_layers, activation=nn.ReLU),
        output_activation=None,
    )
    import torch

    for c in model.gnn_convs:
        if workload.conv == "gin":
            # GINConv_GNNB.hidden_dim defaults to None and is printed verbatim as a template int
            # (model.cpp.jinja:58); the MLP it really builds has hidden = out (models.py:52-53,90)
            c.hidden_dim = c.out_channels
            c.eps = float(workload.gin_eps)
        if workload.conv == "pna":
            c.conv.aggr_module.avg_deg_log = torch.Tensor([pna_delta])
            c.delta_scaler = c.conv.aggr_module.avg_deg_log.item()
    return gnnb, model


def build_model(workload, name: str | None = None, pna_delta: float = 1.0, max_nodes=None,
                max_edges=None, force: bool = False) -> Path:
    """Render with the reference's own Project and compile.  Returns the model directory."""
    name = name or workload.name
    max_nodes = max_nodes or workload.max_nodes
    max_edges = max_edges or workload.max_edges
    model_dir = OUT / "models" / name
    so = model_dir / f"lib{name}.so"
    manifest_fp = model_dir / "manifest.json"
    if not force and so.exists() and manifest_fp.exists():
        return model_dir
    gnnb, model = reference_model(workload, pna_delta)
    proj = gnnb.Project(name, model, "regression", Path("/nonexistent/vitis_hls"),
                        OUT / "models", dataset=None, max_nodes=max_nodes, max_edges=max_edges,
                        float_or_fixed="float")
    proj.gen_hw_model()
    proj.gen_testbench(gen_testbench_data=False)
    proj.gen_makefile()
    _run(["g++", *CXXFLAGS, "-shared", "model.cpp", "-o", so.name], cwd=str(model_dir))
    _run(["g++", *CXXFLAGS, "model.cpp", "model_tb.cpp", "-o", "result"], cwd=str(model_dir))
    manifest = dict(
        name=name, max_nodes=max_nodes, max_edges=max_edges, in_dim=workload.in_dim,
        out_dim=workload.out_dim, pna_delta=pna_delta,
        param_names=list(model.layer_parameter_names_flat),
        param_shapes=[list(s) for s in model.layer_parameter_shapes_flat],
    )
    manifest_fp.write_text(json.dumps(manifest, indent=1))
    return model_dir


def small_variant(workload, hidden: int = 12, in_dim: int = 5):
    """A reduced-width copy of a workload for fast whole-model parity tests."""
    import dataclasses

    return dataclasses.replace(workload, name=workload.name + "_small", hidden_dim=hidden,
                               in_dim=in_dim, mlp_hidden_dim=8, max_nodes=64, max_edges=256)


def build_all(force: bool = False):
    sys.path.insert(0, str(ROOT))
    from gnn_builder_b200.configs import C1, C2, C3, C4

    built = {"oracle": str(build_oracle(force))}
    if not have_reference():
        built["reference"] = "absent (/root/reference not present): using prebuilt oracle/_ref"
        return built
    built["lib_test"] = str(build_lib_test(force))
    built["ref_layers"] = str(build_ref_layers(force))
    for w in (C1, C2, C3, C4):
        built[w.name] = str(build_model(w, pna_delta=w.pna_delta, force=force))
        sv = small_variant(w)
        built[sv.name] = str(build_model(sv, pna_delta=sv.pna_delta, force=force))
    return built


if __name__ == "__main__":
    sys.path.insert(0, str(HERE))
    out = build_all(force="--force" in sys.argv)
    print(json.dumps(out, indent=1))
    if have_reference():
        print(run_lib_test()[-600:])
