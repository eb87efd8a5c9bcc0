"""GPU parity for the large-graph path (BASELINE config 5 shape, scaled to what the CPU oracle
finishes in seconds): the layerwise kernels with degree-bucketed aggregation through the model
handle, and the row-partitioned runner (world size 1 here; world size 2 is exercised by
tests/run_large_multi_gpu.py under torchrun on a 2-GPU box)."""
import dataclasses

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def setup():
    import torch

    assert torch.cuda.is_available()
    import gnn_builder_b200 as gnnb
    from gnn_builder_b200.configs import C5

    w = dataclasses.replace(C5, out_dim=16)
    model = gnnb.build_model(w, seed=1)
    n = 40000
    x, coo = gnnb.make_powerlaw_graph(n, 16, w.in_dim, seed=5, max_degree=5000)
    coo[:3000, 1] = 17      # a heavy row that one CTA finishes by itself (256 < in-degree <= 4096)
    coo[3000:9500, 1] = 23  # a hub whose neighbor list is sliced over several CTAs (> 4096)
    return gnnb, w, model, n, x, coo


def test_large_graph_through_model_handle(setup, orc):
    gnnb, w, model, n, x, coo = setup
    params = list(model.named_parameter_arrays().values())
    ref, ref_emb = orc.model_forward(model.describe(), params, x, coo, return_node_emb=True)
    with gnnb.Engine(model) as eng:
        out = eng.run_graph(x, coo)
        assert eng.last_kernel == "layerwise"
        emb = eng.node_embeddings(n)
    assert rel_err(emb, ref_emb) < TOL
    assert rel_err(out, ref) < TOL


def test_row_partition_runner_world_1(setup, orc):
    import torch

    gnnb, w, model, n, x, coo = setup
    from gnn_builder_b200.distributed import LargeGraphGCN, RowPartition

    params = list(model.named_parameter_arrays().values())
    ref, ref_emb = orc.model_forward(model.describe(), params, x, coo, return_node_emb=True)
    runner = LargeGraphGCN(model, n, 0, 1).setup(RowPartition(n, 1).local_edges(coo, 0))
    out, emb = runner.forward(x, return_embeddings=True)
    torch.cuda.synchronize()
    assert rel_err(emb.cpu().numpy(), ref_emb) < TOL
    assert rel_err(out.cpu().numpy(), ref) < TOL
    # tables of the partition slice are the reference's tables, bit for bit
    ind, outd, off, nbr = orc.tables(coo, n)
    t_ind, t_off, t_nbr = runner.tables
    assert np.array_equal(t_ind.cpu().numpy(), ind)
    assert np.array_equal(t_off.cpu().numpy(), off)
    assert np.array_equal(t_nbr.cpu().numpy()[: coo.shape[0]], nbr)
