"""TEST INFRASTRUCTURE ONLY -- in-memory stand-in for ``torch_geometric``.

The reference package imports ``torch_geometric`` at module scope
(/root/reference/gnnbuilder/models.py:7-18, code_gen.py:11, utils.py:6) and PyG is not
installable offline.  The reference's *code generator* (Project.gen_hw_model /
gen_testbench / gen_makefile) only needs the conv wrappers to carry PyG's parameter
names and shapes -- it never runs a PyG forward when ``gen_testbench_data=False``.  This
module registers tiny ``nn.Module`` placeholders under the ``torch_geometric`` names so
that the reference's own, unmodified ``gnnbuilder`` can be imported and its templates
rendered by its own code.  No PyG arithmetic is reproduced here (calling ``forward`` on a
placeholder raises).

Parameter names/shapes follow PyG >= 2.3 (pinned by the reference's templates,
model.cpp.jinja:48-49,75-78,107-112,137-139, and by gen_test_data.py:219-309):
  GCNConv : bias[F_out], lin.weight[F_out][F_in]
  GINConv : nn.<mlp params>           (GINConv_GNNB also registers the mlp as ``mlp``)
  SAGEConv: lin_l.{weight,bias}, lin_r.weight
  PNAConv : pre_nns.0.0.{weight[F][2F],bias}, post_nns.0.0.{weight[F_out][13F],bias},
            lin.{weight[F_out][F_out],bias}; aggr_module.avg_deg_log
"""
import sys
import types

import torch
import torch.nn as nn


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("fake torch_geometric placeholder: no forward arithmetic here")


class GCNConv(_NoForward):
    def __init__(self, in_channels, out_channels, **kw):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin = nn.Linear(in_channels, out_channels, bias=False)


class GINConv(_NoForward):
    def __init__(self, nn_module, eps=0.0, train_eps=False, **kw):
        super().__init__()
        self.nn = nn_module
        self.initial_eps = eps
        self.register_buffer("eps", torch.tensor([float(eps)]))


class GINEConv(_NoForward):
    def __init__(self, nn_module, eps=0.0, train_eps=False, edge_dim=None, **kw):
        super().__init__()
        self.nn = nn_module
        self.register_buffer("eps", torch.tensor([float(eps)]))
        in_channels = getattr(nn_module, "in_features", None)
        self.lin = nn.Linear(edge_dim, in_channels) if edge_dim is not None else None


class SAGEConv(_NoForward):
    def __init__(self, in_channels, out_channels, **kw):
        super().__init__()
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)


class _AggrModule(nn.Module):
    pass


class PNAConv(_NoForward):
    def __init__(self, in_channels, out_channels, aggregators, scalers, deg, **kw):
        super().__init__()
        F = in_channels
        self.aggr_module = _AggrModule()
        self.pre_nns = nn.ModuleList([nn.Sequential(nn.Linear(2 * F, F))])
        self.post_nns = nn.ModuleList(
            [nn.Sequential(nn.Linear((len(aggregators) * len(scalers) + 1) * F, out_channels))]
        )
        self.lin = nn.Linear(out_channels, out_channels)


class GATConv(_NoForward):
    def __init__(self, *a, **k):
        super().__init__()


class LGConv(_NoForward):
    def __init__(self, *a, **k):
        super().__init__()


class SimpleConv(_NoForward):
    def __init__(self, *a, **k):
        super().__init__()


class _MultiAggregation(_NoForward):
    def __init__(self, aggrs, mode="cat", **kw):
        super().__init__()
        self.aggrs, self.mode = list(aggrs), mode


class _Dataset:  # placeholders for isinstance/type annotations only
    pass


def install():
    """Register the placeholders as ``torch_geometric`` (no-op if a real PyG is importable)."""
    if "torch_geometric" in sys.modules:
        return
    try:  # pragma: no cover
        import torch_geometric  # noqa: F401

        return
    except Exception:
        pass
    tg = types.ModuleType("torch_geometric")
    tg.__fake__ = True
    tg_nn = types.ModuleType("torch_geometric.nn")
    for cls in (GCNConv, GINConv, GINEConv, SAGEConv, PNAConv, GATConv, LGConv, SimpleConv):
        setattr(tg_nn, cls.__name__, cls)
    aggr = types.ModuleType("torch_geometric.nn.aggr")
    aggr.MultiAggregation = _MultiAggregation
    tg_nn.aggr = aggr
    tg_typing = types.ModuleType("torch_geometric.typing")
    tg_typing.Adj = torch.Tensor
    tg_data = types.ModuleType("torch_geometric.data")
    tg_data.Dataset = _Dataset
    tg_data.InMemoryDataset = _Dataset
    tg_data.Data = _Dataset
    tg_utils = types.ModuleType("torch_geometric.utils")
    tg_utils.degree = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("fake pyg"))
    tg.nn, tg.typing, tg.data, tg.utils = tg_nn, tg_typing, tg_data, tg_utils
    sys.modules.update(
        {
            "torch_geometric": tg,
            "torch_geometric.nn": tg_nn,
            "torch_geometric.nn.aggr": aggr,
            "torch_geometric.typing": tg_typing,
            "torch_geometric.data": tg_data,
            "torch_geometric.utils": tg_utils,
        }
    )
