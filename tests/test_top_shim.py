"""The C++ seam: ``Project.gen_hw_model`` renders ``model.cpp`` = ``extern "C" <name>_top`` with the
reference's exact signature (model.h.jinja:67-79) on top of the C-ABI, so the reference's own
``model.h`` + ``model_tb.cpp`` build against the GPU backend unchanged.

not gpu : the generated top compiles, exports ``<name>_top``, and -- where the reference's rendered
          files exist (oracle/_ref) -- compiles against the reference's own declaration and links
          the reference's testbench.
gpu     : calling ``<name>_top`` the way the reference's testbench does reproduces the committed
          outputs of the reference's ``<name>_top``; the reference's testbench BINARY linked
          against the GPU top reports the reference's MAE file format with MAE ~ 0.
"""
import ctypes as C
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_model_golden, model_and_params, rel_err

REF_MODELS = ROOT / "oracle" / "_ref" / "models"
NAMES = ["c1_gcn_esol", "c2_gin_qm9_small", "c3_sage_hiv_small", "c4_pna_lipo_small"]


def _project(name, tmp_path):
    from gnn_builder_b200.code_gen import Project

    w, model, params = model_and_params(name)
    proj = Project(name, model, "regression", None, tmp_path, max_nodes=w.max_nodes,
                   max_edges=w.max_edges)
    proj.model_dir.mkdir(parents=True, exist_ok=True)
    (proj.model_dir / "model.cpp").write_text(proj.render_top())  # (gen_hw_model also binds a GPU)
    return w, model, params, proj


def _ref_dir(name):
    d = REF_MODELS / name
    return d if (d / "model_tb.cpp").exists() and (d / "model.h").exists() else None


@pytest.mark.parametrize("name", NAMES)
def test_generated_top_builds_and_matches_reference_declaration(name, tmp_path):
    w, model, params, proj = _project(name, tmp_path)
    src = (proj.model_dir / "model.cpp").read_text()
    # one trailing array per parameter, in the reference's flat order (models.py:607-624)
    pos = [src.index(f"float {n}_fixed_in[") for n in model.layer_parameter_names_flat]
    assert pos == sorted(pos)
    so = proj.build_top(testbench=_ref_dir(name))
    lib = C.CDLL(str(so))
    assert hasattr(lib, f"{name}_top")
    if _ref_dir(name) is not None:   # the reference's own testbench linked against the GPU top
        assert (proj.model_dir / "result").exists()


def test_signature_mismatch_is_a_compile_error(tmp_path):
    """the reference's model.h really is checked: a top rendered for other shapes must not build"""
    name = "c2_gin_qm9_small"
    if _ref_dir(name) is None:
        pytest.skip("reference build (oracle/_ref) not present")
    import dataclasses

    from gnn_builder_b200.code_gen import Project
    from gnn_builder_b200.models import build_model

    w, model, params, proj = _project(name, tmp_path)
    proj.build_top(testbench=_ref_dir(name))   # the matching model builds ...
    wide = build_model(dataclasses.replace(w, hidden_dim=16), pna_delta=w.pna_delta, seed=0)
    other = Project(name, wide, "regression", None, tmp_path, max_nodes=w.max_nodes,
                    max_edges=w.max_edges)   # ... weight arrays [16][16] against the declared [12][12]
    (other.model_dir / "model.cpp").write_text(other.render_top())
    with pytest.raises(RuntimeError, match="conflicting declaration"):
        other.build_top(testbench=_ref_dir(name))


def _call_top(so, name, w, params_in_order, batch, out_dim):
    """the call model_tb.cpp makes (model_tb.cpp.jinja:157-204): padded static tables, flag = 1 once"""
    lib = C.CDLL(str(so))
    top = getattr(lib, f"{name}_top")
    top.restype = None
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    xbuf = np.zeros((w.max_nodes, w.in_dim), np.float32)
    ebuf = np.zeros((w.max_edges, 2), np.int32)
    obuf = np.zeros(out_dim, np.float32)
    ps = [np.ascontiguousarray(p, np.float32) for p in params_in_order]
    outs = []
    for g in range(batch.n_graphs):
        x, coo = batch.graph(g)
        xbuf[: x.shape[0]] = x
        ebuf[: coo.shape[0]] = coo
        top(xbuf.ctypes.data_as(fp), ebuf.ctypes.data_as(ip), obuf.ctypes.data_as(fp),
            C.c_int(x.shape[0]), C.c_int(coo.shape[0]), C.c_int(1 if g == 0 else 0),
            *[p.ctypes.data_as(fp) for p in ps])
        outs.append(obuf.copy())
    return np.stack(outs)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_top_reproduces_reference_outputs(name, tmp_path):
    batch, gold, _, _ = load_model_golden(name)
    w, model, params, proj = _project(name, tmp_path)
    so = proj.build_top()
    order = [params[n] for n in model.layer_parameter_names_flat]
    n = min(batch.n_graphs, 64)
    out = _call_top(so, name, w, order, batch.slice(0, n), gold.shape[1])
    assert rel_err(out, gold[:n]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2_gin_qm9_small", "c1_gcn_esol"])
def test_reference_testbench_binary_on_the_gpu_top(name, tmp_path):
    if _ref_dir(name) is None:
        pytest.skip("reference build (oracle/_ref) not present")
    from gnn_builder_b200.data import write_tb_data

    batch, gold, _, _ = load_model_golden(name)
    w, model, params, proj = _project(name, tmp_path)
    proj.build_top(testbench=_ref_dir(name))
    n = min(batch.n_graphs, 32)
    write_tb_data(proj.model_dir / "tb_data", params, batch.slice(0, n), gold[:n], gold[:n],
                  gold.shape[1])
    r = subprocess.run(["./result"], cwd=proj.model_dir, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    mae = float((proj.model_dir / "tb_data" / "model_output_mae.txt").read_text().split()[1])
    runtime = float((proj.model_dir / "tb_data" / "model_runtime.txt").read_text().split()[1])
    assert mae < 1e-5 and runtime > 0.0
