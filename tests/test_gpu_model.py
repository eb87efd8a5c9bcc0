"""GPU parity, whole model: gnnb_model_run_batch / run_graph (both kernels) against

  * the committed outputs of the reference's own generated <name>_top (tests/golden/models),
  * the CPU oracle on larger seeded batches,
  * size-independent properties at BASELINE.json's full batch size.
"""
import numpy as np
import pytest

from conftest import MODEL_NAMES, load_model_golden, model_and_params, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4  # north star: <= 1e-4 relative in fp32


@pytest.fixture(scope="module")
def gnnb():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import gnn_builder_b200

    return gnn_builder_b200


def _paths(gnnb, eng, batch):
    """run through every path the engine supports for this batch"""
    outs = {}
    eng.set_path(gnnb.PATH_LAYERWISE)
    outs["layerwise"] = eng.run(batch)
    assert eng.last_path == gnnb.PATH_LAYERWISE and eng.last_launches > 0
    eng.set_path(gnnb.PATH_AUTO)
    outs["auto"] = eng.run(batch)
    if eng.last_path == gnnb.PATH_FUSED:
        assert eng.last_launches > 0
        outs["fused"] = outs["auto"]
    # forced fused path: the tcgen05 kernel where the widths allow it, else the fp32-FMA fused
    # kernel (fused.cu) -- AUTO never picks that one, so it is exercised here
    eng.set_path(gnnb.PATH_FUSED)
    try:
        outs["forced-" + "fused"] = eng.run(batch)
        outs["forced-kernel"] = eng.last_kernel
        assert eng.last_kernel in ("fused-tcgen05", "fused-fma") and eng.last_launches > 0
    except gnnb._lib.GnnbError as e:
        assert "fused path requested" in str(e), e
    eng.set_path(gnnb.PATH_AUTO)
    return outs


@pytest.mark.parametrize("name", [m + s for m in MODEL_NAMES for s in ("_small", "")])
def test_golden_reference_top(gnnb, name):
    batch, gold, _, _ = load_model_golden(name)
    w, model, params = model_and_params(name)
    with gnnb.Engine(model, max_nodes=w.max_nodes, max_edges=w.max_edges) as eng:
        outs = _paths(gnnb, eng, batch)
        kernel = outs.pop("forced-kernel", None)
        if name.endswith("_small") and not name.startswith("c4_pna"):
            assert kernel == "fused-fma", (name, kernel)     # widths 12/5: fused.cu is the one that runs
        for path, out in outs.items():
            assert rel_err(out, gold) < TOL, (name, path, rel_err(out, gold))
        # one graph per call, the way <name>_top is called (model_tb.cpp.jinja:189-204)
        for g in range(min(4, batch.n_graphs)):
            assert rel_err(eng.run_graph(*batch.graph(g)), gold[g]) < TOL


@pytest.mark.parametrize("name", ["c1_gcn_esol", "c2_gin_qm9", "c3_sage_hiv"])
def test_strict_mode_bit_identical_to_reference(gnnb, name):
    """STRICT math: same operation order, no FMA contraction => the reference's bits."""
    for nm in (name + "_small", name):
        batch, gold, _, _ = load_model_golden(nm)
        w, model, _ = model_and_params(nm)
        with gnnb.Engine(model, path=gnnb.PATH_LAYERWISE, math=gnnb.MATH_STRICT) as eng:
            assert np.array_equal(eng.run(batch), gold), nm


@pytest.mark.parametrize("name", MODEL_NAMES)
def test_batch_vs_oracle(gnnb, orc, name):
    w, model, params = model_and_params(name)
    batch = gnnb.make_molecular_batch(600, w.mu_nodes, w.mu_edges, w.in_dim, seed=77 + w.seed)
    ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
    with gnnb.Engine(model) as eng:
        outs = _paths(gnnb, eng, batch)
        kernel = outs.pop("forced-kernel", None)
        for path, out in outs.items():
            assert rel_err(out, ref) < TOL, (name, path, rel_err(out, ref))
        # AUTO picks the tcgen05 fused kernel for all four BASELINE convs (PNA since round 2)
        assert "fused" in outs and kernel == "fused-tcgen05", (name, list(outs), kernel)


def test_edge_cases(gnnb, orc):
    """ragged batch: 1-node graphs without edges, an empty edge list, a graph at the capacity
    limit, duplicate edges and self loops"""
    w, model, params = model_and_params("c2_gin_qm9_small")
    rng = np.random.default_rng(4)
    graphs = []
    graphs.append((rng.uniform(-1, 1, (1, w.in_dim)), np.zeros((0, 2), np.int32)))
    graphs.append((rng.uniform(-1, 1, (5, w.in_dim)), np.zeros((0, 2), np.int32)))
    graphs.append((rng.uniform(-1, 1, (2, w.in_dim)), np.array([[0, 1], [0, 1], [1, 1]], np.int32)))
    n = 60
    graphs.append((rng.uniform(-1, 1, (n, w.in_dim)),
                   np.stack([rng.integers(0, n, 240), rng.integers(0, n, 240)], 1)))
    graphs.append((rng.uniform(-1, 1, (3, w.in_dim)), np.array([[2, 0]], np.int32)))
    batch = gnnb.GraphBatch.from_graphs(graphs)
    ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
    with gnnb.Engine(model, max_nodes=64, max_edges=256) as eng:
        outs = _paths(gnnb, eng, batch)
        assert outs.pop("forced-kernel", None) == "fused-fma"
        for path, out in outs.items():
            assert rel_err(out, ref) < TOL, path
        # capacity violations are errors, not undefined behaviour like the reference
        big = gnnb.GraphBatch.from_graphs([(rng.uniform(-1, 1, (65, w.in_dim)),
                                            np.zeros((0, 2), np.int32))])
        with pytest.raises(gnnb._lib.GnnbError):
            eng.run(big)
        assert eng.run(gnnb.GraphBatch.from_graphs(graphs[:1])).shape == (1, w.out_dim)


def test_pna_zero_in_degree_nan_semantics(gnnb, orc):
    """PNA's std is NaN for zero-in-degree nodes (0/0, lib:702).  Through the whole model the
    reference's relu (`x > 0 ? x : 0`, lib:362-375) turns those rows into 0, with sigmoid they
    stay NaN: the GPU must follow the reference in both cases."""
    import dataclasses
    from gnn_builder_b200.models import build_model

    w, model, params = model_and_params("c4_pna_lipo_small")
    x = np.random.default_rng(1).uniform(-1, 1, (4, w.in_dim)).astype(np.float32)
    coo = np.array([[0, 1], [1, 2], [2, 1]], np.int32)   # nodes 0 and 3 have no in-edges
    ref = orc.model_forward(model.describe(), list(params.values()), x, coo)
    assert np.isfinite(ref).all()
    with gnnb.Engine(model, path=gnnb.PATH_LAYERWISE) as eng:
        assert rel_err(eng.run_graph(x, coo), ref) < TOL
    w2 = dataclasses.replace(w, activation="sigmoid")
    model2 = build_model(w2, pna_delta=w2.pna_delta, seed=0)
    p2 = model2.named_parameter_arrays()
    ref2, emb2 = orc.model_forward(model2.describe(), list(p2.values()), x, coo,
                                   return_node_emb=True)
    assert np.isnan(emb2[[0, 3]]).all()   # (NaN then spreads along edges 0->1->2 layer by layer)
    with gnnb.Engine(model2, path=gnnb.PATH_LAYERWISE) as eng:
        out2 = eng.run_graph(x, coo)
        emb = eng.node_embeddings(4)
        assert np.array_equal(np.isnan(emb), np.isnan(emb2))
        assert rel_err(np.nan_to_num(emb), np.nan_to_num(emb2)) < TOL
        # the head's relu maps the NaN pooled sums to 0 exactly like the reference
        assert np.array_equal(np.isnan(out2), np.isnan(ref2))
        assert rel_err(np.nan_to_num(out2), np.nan_to_num(ref2)) < TOL


def test_medium_graphs_use_layerwise(gnnb, orc):
    """graphs larger than a CTA tile (HIV has molecules of 222 atoms) fall back to kernel (2)"""
    w, model, params = model_and_params("c3_sage_hiv")
    batch = gnnb.make_molecular_batch(6, 300, 640, w.in_dim, seed=5, max_nodes=600)
    ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
    with gnnb.Engine(model) as eng:
        out = eng.run(batch)
        assert eng.last_path == gnnb.PATH_LAYERWISE
        assert rel_err(out, ref) < TOL
        emb = eng.node_embeddings(batch.total_nodes)
        _, ref_emb = orc.model_forward(model.describe(), list(params.values()), *batch.graph(0),
                                       return_node_emb=True)
        assert rel_err(emb[: ref_emb.shape[0]], ref_emb) < TOL


def test_device_pointer_api(gnnb, orc):
    import torch

    w, model, params = model_and_params("c2_gin_qm9")
    batch = gnnb.make_molecular_batch(300, w.mu_nodes, w.mu_edges, w.in_dim, seed=12)
    ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
    dx, dcoo = torch.from_numpy(batch.x).cuda(), torch.from_numpy(batch.coo).cuda()
    dn, de = torch.from_numpy(batch.node_ptr).cuda(), torch.from_numpy(batch.edge_ptr).cuda()
    out = torch.empty((batch.n_graphs, w.out_dim), device="cuda")
    with gnnb.Engine(model, max_nodes=64, max_edges=256) as eng:
        eng.run_device_sync(dx, dcoo, dn, de, out, batch.n_graphs)
        assert rel_err(out.cpu().numpy(), ref) < TOL
        out.zero_()
        eng.run_device(dx, dcoo, dn, de, out, batch.n_graphs, batch.total_nodes,
                       batch.total_edges, sync=True)
        assert rel_err(out.cpu().numpy(), ref) < TOL


def test_full_size_properties_c2(gnnb, orc):
    """BASELINE config 2 at full size (1M QM9-shaped graphs): outputs do not depend on batch
    composition (a graph's result is the same alone, in a slice, or in the full batch), and a
    random sample agrees with the oracle."""
    from gnn_builder_b200.configs import C2

    w, model, params = model_and_params("c2_gin_qm9")
    batch = gnnb.make_molecular_batch(C2.n_graphs, w.mu_nodes, w.mu_edges, w.in_dim, seed=w.seed)
    with gnnb.Engine(model, max_nodes=w.max_nodes, max_edges=w.max_edges) as eng:
        out = eng.run(batch)
        assert out.shape == (C2.n_graphs, w.out_dim) and np.isfinite(out).all()
        path = eng.last_path
        rng = np.random.default_rng(0)
        idx = np.sort(rng.choice(C2.n_graphs, 48, replace=False))
        sample = gnnb.GraphBatch.from_graphs([batch.graph(int(g)) for g in idx])
        ref = orc.model_forward_batch(model.describe(), list(params.values()), sample)
        assert rel_err(out[idx], ref) < TOL
        sl = batch.slice(500_000, 500_200)
        out_sl = eng.run(sl)
        assert eng.last_path == path
        # the tensor-core aggregation sums a row's neighbors in an order that depends on where the
        # graph sits inside its 128-row tile, so "independent of batch composition" holds to fp32
        # rounding (a few ulp), not bit for bit; the layerwise path is bit-stable
        assert rel_err(out_sl, out[500_000:500_200]) < 2e-6
        dup = gnnb.GraphBatch.from_graphs([batch.graph(3)] * 5)
        o = eng.run(dup)
        assert all(rel_err(o[0], o[i]) < 2e-6 for i in range(5))
        eng.set_path(gnnb.PATH_LAYERWISE)
        o = eng.run(dup)
        assert all(np.array_equal(o[0], o[i]) for i in range(5))
        assert np.array_equal(eng.run(sl), eng.run(batch.slice(499_990, 500_300))[10:210])


def test_project_flow(gnnb, tmp_path):
    """the reference's user-facing call sequence (demos/demo.py:102-129)"""
    w, model, params = model_and_params("c1_gcn_esol")
    ds = gnnb.make_molecular_batch(20, w.mu_nodes, w.mu_edges, w.in_dim, seed=8)
    proj = gnnb.Project("gcn_esol", model, "regression", None, tmp_path, dataset=ds, max_nodes=600,
                        max_edges=600, float_or_fixed="float")
    proj.gen_hw_model()
    proj.gen_testbench()
    proj.gen_makefile()
    data = proj.build_and_run_testbench()
    assert set(data) == {"model_output_mae", "model_runtime"}
    assert data["model_output_mae"] < 1e-5 and data["model_runtime"] > 0
    data_b = proj.build_and_run_testbench(batched=True)
    assert data_b["model_output_mae"] < 1e-5
    with pytest.raises(NotImplementedError):
        proj.run_vitis_hls_synthesis()


def test_fused_tc_used_for_gcn_gin_sage(gnnb, orc):
    """the tensor-core fused kernel is the AUTO choice for the three molecular configs it covers"""
    for name in ("c1_gcn_esol", "c2_gin_qm9", "c3_sage_hiv"):
        w, model, params = model_and_params(name)
        batch = gnnb.make_molecular_batch(200, w.mu_nodes, w.mu_edges, w.in_dim, seed=3)
        ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
        with gnnb.Engine(model) as eng:
            out = eng.run(batch)
            assert eng.last_kernel == "fused-tcgen05", (name, eng.last_kernel)
            assert rel_err(out, ref) < TOL, name


def test_tile_packing_edge_cases(gnnb, orc):
    """greedy device-side tile packing: runs of empty graphs, > 128 one-node graphs in a row, graphs
    of exactly 128 nodes, a batch that spans several 2048-graph packing chunks"""
    w, model, params = model_and_params("c2_gin_qm9")
    rng = np.random.default_rng(9)
    graphs = []
    for _ in range(300):
        graphs.append((np.zeros((0, w.in_dim), np.float32), np.zeros((0, 2), np.int32)))
    for _ in range(400):
        graphs.append((rng.uniform(-1, 1, (1, w.in_dim)), np.zeros((0, 2), np.int32)))
    for n in (128, 127, 128, 1, 128):
        graphs.append((rng.uniform(-1, 1, (n, w.in_dim)),
                       np.stack([rng.integers(0, n, 3 * n), rng.integers(0, n, 3 * n)], 1)))
    mol = gnnb.make_molecular_batch(5000, w.mu_nodes, w.mu_edges, w.in_dim, seed=21)
    graphs += [mol.graph(g) for g in range(mol.n_graphs)]
    graphs.append((np.zeros((0, w.in_dim), np.float32), np.zeros((0, 2), np.int32)))
    batch = gnnb.GraphBatch.from_graphs(graphs)
    ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
    with gnnb.Engine(model) as eng:
        out = eng.run(batch)
        assert eng.last_kernel == "fused-tcgen05"
        assert rel_err(out, ref) < TOL


def test_non_finite_inputs_do_not_leak_between_graphs(gnnb, orc):
    """Inside the tensor-core aggregation 0 x Inf would poison the other graphs of a tile; the
    kernel detects non-finite activations and the batch is redone on the layerwise path, so every
    graph's output is what the reference computes for that graph alone."""
    w, model, params = model_and_params("c2_gin_qm9")
    batch = gnnb.make_molecular_batch(64, w.mu_nodes, w.mu_edges, w.in_dim, seed=5)
    x = batch.x.copy()
    r0 = int(batch.node_ptr[7])
    x[r0, 2] = np.inf
    x[int(batch.node_ptr[20]) + 1, 0] = np.nan
    bad = gnnb.GraphBatch(x, batch.coo, batch.node_ptr, batch.edge_ptr)
    ref = orc.model_forward_batch(model.describe(), list(params.values()), bad)
    clean = orc.model_forward_batch(model.describe(), list(params.values()), batch)
    with gnnb.Engine(model) as eng:
        out = eng.run(bad)
        assert eng.last_path == gnnb.PATH_LAYERWISE      # fell back
        ok = np.ones(64, bool)
        ok[[7, 20]] = False
        assert np.isfinite(out[ok]).all()
        assert rel_err(out[ok], clean[ok]) < TOL
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        # an explicit request for the fused path reports the problem instead of falling back
        eng.set_path(gnnb.PATH_FUSED)
        with pytest.raises(gnnb._lib.GnnbError):
            eng.run(bad)
        eng.set_path(gnnb.PATH_AUTO)
        assert rel_err(eng.run(batch), clean) < TOL
        assert eng.last_kernel == "fused-tcgen05"


PNA_VARIANTS = {
    # tensor-memory budget edge (hidden 96), two feature groups of different width (in_dim 40:
    # 32 + 16 features), one group of 16, four layers, non-ReLU activation, no skip, single pool
    "pna_96": dict(hidden_dim=96),
    "pna_wide_input_48": dict(hidden_dim=48, in_dim=40, mlp_hidden_dim=32),
    "pna_16_4layers_tanh": dict(hidden_dim=16, num_layers=4, activation="tanh"),
    "pna_64_noskip_sigmoid_maxpool": dict(hidden_dim=64, skip=False, activation="sigmoid", pools=["max"]),
    "pna_one_layer": dict(num_layers=1, hidden_dim=80),
}


@pytest.mark.parametrize("variant", sorted(PNA_VARIANTS))
def test_fused_tc_pna_variants(gnnb, orc, variant):
    """PNA in the fused tcgen05 kernel (gather statistics in shared memory, three accumulators for
    the degree scalers) against the oracle"""
    import dataclasses

    from conftest import workload_by_name
    from gnn_builder_b200.models import build_model

    w = dataclasses.replace(workload_by_name("c4_pna_lipo"), **PNA_VARIANTS[variant])
    model = build_model(w, pna_delta=w.pna_delta, seed=11)
    params = model.named_parameter_arrays()
    batch = gnnb.make_molecular_batch(700, w.mu_nodes, w.mu_edges, w.in_dim, seed=5)
    ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
    with gnnb.Engine(model) as eng:
        out = eng.run(batch)
        assert eng.last_kernel == "fused-tcgen05", (variant, eng.last_kernel)
        assert rel_err(out, ref) < TOL, (variant, rel_err(out, ref))


def test_pna_beyond_the_fused_limits_runs_layerwise(gnnb, orc):
    """PNA wider than the fused kernel's tensor-memory budget (hidden 128 > 96), or with more input
    features than output features, takes the layerwise kernels under AUTO -- same results"""
    import dataclasses

    from conftest import workload_by_name
    from gnn_builder_b200.models import build_model

    for over in (dict(hidden_dim=128), dict(hidden_dim=32, in_dim=40)):
        w = dataclasses.replace(workload_by_name("c4_pna_lipo"), **over)
        model = build_model(w, pna_delta=w.pna_delta, seed=5)
        params = model.named_parameter_arrays()
        batch = gnnb.make_molecular_batch(300, w.mu_nodes, w.mu_edges, w.in_dim, seed=6)
        ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
        with gnnb.Engine(model) as eng:
            out = eng.run(batch)
            assert eng.last_kernel == "layerwise", (over, eng.last_kernel)
            assert rel_err(out, ref) < TOL, over


def test_fused_pna_zero_in_degree_and_multi_edges(gnnb, orc):
    """the fused PNA path keeps the reference's semantics per graph: std = NaN for in-degree 0
    (0/0, lib:702) -- which ReLU turns into 0 and sigmoid keeps --, duplicate edges count twice in
    mean / std, and a NaN graph does not touch its tile neighbours (PNA never mixes graphs)"""
    import dataclasses

    from conftest import workload_by_name
    from gnn_builder_b200.models import build_model

    for act in ("relu", "sigmoid"):
        w = dataclasses.replace(workload_by_name("c4_pna_lipo"), activation=act)
        model = build_model(w, pna_delta=w.pna_delta, seed=3)
        params = model.named_parameter_arrays()
        rng = np.random.default_rng(2)
        mol = gnnb.make_molecular_batch(40, w.mu_nodes, w.mu_edges, w.in_dim, seed=9)
        graphs = [mol.graph(g) for g in range(mol.n_graphs)]
        graphs.insert(3, (rng.uniform(-1, 1, (4, w.in_dim)).astype(np.float32),
                          np.array([[0, 1], [1, 2], [2, 1], [0, 1], [0, 1]], np.int32)))   # nodes 0, 3: no in-edges
        graphs.insert(17, (rng.uniform(-1, 1, (1, w.in_dim)).astype(np.float32), np.zeros((0, 2), np.int32)))
        star = np.stack([np.arange(1, 14), np.zeros(13, np.int64)], 1)         # node 0: 13 distinct in-neighbors
        graphs.insert(25, (rng.uniform(-1, 1, (14, w.in_dim)).astype(np.float32),
                           np.concatenate([star, star[:, ::-1], star[:3]]).astype(np.int32)))
        batch = gnnb.GraphBatch.from_graphs(graphs)
        ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
        with gnnb.Engine(model) as eng:
            out = eng.run(batch)
            assert eng.last_kernel == "fused-tcgen05"
        assert np.array_equal(np.isnan(out), np.isnan(ref)), act
        assert rel_err(np.nan_to_num(out), np.nan_to_num(ref)) < TOL, act
        assert np.isfinite(ref[[0, 1, 2, 4, 5]]).all()     # (the head's ReLU maps a NaN pooled vector to 0)


VARIANTS = {
    # (base workload, overrides): shapes and options the BASELINE configs do not reach
    "gin_eps": ("c2_gin_qm9", dict(gin_eps=0.25)),                        # eps * x_v added in registers
    "gin_gelu_noskip": ("c2_gin_qm9", dict(activation="gelu", skip=False, hidden_dim=64)),
    "gin_min_width_wide_input": ("c2_gin_qm9", dict(hidden_dim=16, in_dim=40, mlp_hidden_dim=16)),
    "gcn_48_tanh_4layers": ("c1_gcn_esol", dict(hidden_dim=48, num_layers=4, activation="tanh")),
    "sage_96_max_pool": ("c3_sage_hiv", dict(hidden_dim=96, num_layers=2, pools=["max"],
                                               mlp_hidden_layers=1)),
    "sage_sigmoid_5layers": ("c3_sage_hiv", dict(hidden_dim=32, num_layers=5, activation="sigmoid")),
    "gcn_one_layer": ("c1_gcn_esol", dict(num_layers=1, hidden_dim=128, pools=["add", "max"])),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_fused_tc_model_variants(gnnb, orc, variant):
    """layer widths with 1-4 K atoms and partial column blocks, GIN eps, non-ReLU activations, with
    and without skip connections, 1-5 layers, single pools: fused tcgen05 kernel against the oracle"""
    import dataclasses

    from conftest import workload_by_name
    from gnn_builder_b200.models import build_model

    base, over = VARIANTS[variant]
    w = dataclasses.replace(workload_by_name(base), **over)
    model = build_model(w, pna_delta=w.pna_delta, seed=11)
    params = model.named_parameter_arrays()
    batch = gnnb.make_molecular_batch(700, w.mu_nodes, w.mu_edges, w.in_dim, seed=5)
    ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
    with gnnb.Engine(model) as eng:
        out = eng.run(batch)
        assert eng.last_kernel == "fused-tcgen05", (variant, eng.last_kernel)
        assert rel_err(out, ref) < TOL, (variant, rel_err(out, ref))
        eng.set_path(gnnb.PATH_LAYERWISE)
        assert rel_err(eng.run(batch), ref) < TOL


def test_two_handles_from_two_host_threads(gnnb, orc):
    """the boundary contract: handles are independent -- own stream, own weights and
    workspaces, thread-local error state -- so two host threads can drive two models at once
    (the generated top is not re-entrant: file-scope statics, model.cpp.jinja:7-22)"""
    import threading

    jobs = []
    for name, seed in (("c2_gin_qm9", 31), ("c1_gcn_esol", 32), ("c4_pna_lipo", 33)):
        w, model, params = model_and_params(name)
        batch = gnnb.make_molecular_batch(3000, w.mu_nodes, w.mu_edges, w.in_dim, seed=seed)
        ref = orc.model_forward_batch(model.describe(), list(params.values()), batch)
        jobs.append((name, model, batch, ref))
    errors, results = [], {}

    def work(name, model, batch, ref):
        try:
            with gnnb.Engine(model) as eng:
                for _ in range(6):
                    out = eng.run(batch)
                    if rel_err(out, ref) >= TOL:
                        errors.append((name, rel_err(out, ref)))
                results[name] = out
        except Exception as exc:  # noqa: BLE001
            errors.append((name, repr(exc)))

    threads = [threading.Thread(target=work, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert set(results) == {j[0] for j in jobs}



@pytest.mark.parametrize("name", ["c1_gcn_esol", "c2_gin_qm9", "c3_sage_hiv", "c4_pna_lipo"])
def test_fused_tcgen05_run_to_run_bit_identity(gnnb, name):
    """the same batch 20 times through the tcgen05 fused kernel gives the same bits every time:
    every hand-off between the row passes and the MMAs is ordered by mbarriers, so no result
    depends on scheduling (racecheck cannot see tcgen05.commit arrivals; this test is the
    evidence that its reports are false positives)"""
    w, model, _ = model_and_params(name)
    batch = gnnb.make_molecular_batch(5000, w.mu_nodes, w.mu_edges, w.in_dim, seed=31 + w.seed)
    with gnnb.Engine(model, path=gnnb.PATH_FUSED) as eng:
        first = eng.run(batch).copy()
        assert eng.last_kernel == "fused-tcgen05"
        for _ in range(19):
            assert np.array_equal(eng.run(batch), first)


def test_layerwise_path_rejects_out_of_range_edges(gnnb):
    """an edge endpoint outside its graph is an error on EVERY path (the fused kernels report it
    as status 2; the layerwise table kernels flag it), never an out-of-bounds access"""
    from gnn_builder_b200 import layers

    w, model, _ = model_and_params("c2_gin_qm9_small")
    rng = np.random.default_rng(0)
    good = (rng.uniform(-1, 1, (6, w.in_dim)), np.array([[0, 1], [1, 2], [5, 0]], np.int32))
    for bad_edge in ([0, 6], [7, 1], [-1, 2], [2, -5]):
        bad = (rng.uniform(-1, 1, (6, w.in_dim)), np.array([[0, 1], bad_edge], np.int32))
        batch = gnnb.GraphBatch.from_graphs([good, bad, good])
        for path in (gnnb.PATH_LAYERWISE, gnnb.PATH_AUTO, gnnb.PATH_FUSED):
            with gnnb.Engine(model, path=path) as eng:
                with pytest.raises(gnnb._lib.GnnbError, match="outside"):
                    eng.run(batch)
                # the handle stays usable and correct afterwards
                assert np.isfinite(eng.run(gnnb.GraphBatch.from_graphs([good]))).all()
        with pytest.raises(gnnb._lib.GnnbError, match="outside"):
            layers.compute_degree_tables(bad[1], 6)


def test_pinning_caller_buffers_in_place(gnnb):
    """Engine.pin_batch page-locks the caller's numpy arrays (cudaHostRegister): same results, the
    arrays are usable as before, unpin / close release them"""
    import ctypes as C

    w, model, _ = model_and_params("c2_gin_qm9")
    batch = gnnb.make_molecular_batch(3000, w.mu_nodes, w.mu_edges, w.in_dim, seed=12)
    with gnnb.Engine(model) as eng:
        ref = eng.run(batch).copy()
        eng.pin_batch(batch)
        eng.pin_batch(batch)                       # idempotent
        assert len(eng._pinned) == 4
        assert np.array_equal(eng.run(batch), ref)
        batch.x[0, 0] += 1.0                       # still ordinary writable memory
        assert not np.array_equal(eng.run(batch), ref)
        eng.unpin()
        assert not eng._pinned
        batch.x[0, 0] -= 1.0
        assert np.array_equal(eng.run(batch), ref)


def test_layerwise_tables_from_sorted_keys_match_the_atomic_path(gnnb):
    """edge lists of >= 2^20 edges take their in-degrees from run lengths of the sorted destination
    keys, smaller ones from atomics: the layerwise path is bit-stable across batch composition, so
    one 30k-graph batch (1.1M edges) must give the bits its two halves give"""
    w, model, _ = model_and_params("c2_gin_qm9")
    batch = gnnb.make_molecular_batch(30000, w.mu_nodes, w.mu_edges, w.in_dim, seed=17)
    assert batch.total_edges >= (1 << 20) and batch.slice(0, 15000).total_edges < (1 << 20)
    for math in (gnnb.MATH_FAST, gnnb.MATH_STRICT):
        with gnnb.Engine(model, path=gnnb.PATH_LAYERWISE, math=math) as eng:
            whole = eng.run(batch)
            halves = np.concatenate([eng.run(batch.slice(0, 15000)), eng.run(batch.slice(15000, 30000))])
        assert np.array_equal(whole, halves), math
