"""The five BASELINE.json workloads as data (model shape + synthetic dataset shape).

Model hyper-parameters follow SURVEY.md section 8(d): ReLU, skip connections, MLP head
hidden_dim=64 / hidden_layers=2 (gnnbuilder/models.py:370-371 defaults), pools as in
experiments/build_gnnbuilder_benchmarks.py:70 unless BASELINE.json says otherwise.
"""
from dataclasses import dataclass, field
from typing import List


@dataclass(frozen=True)
class Workload:
    name: str
    conv: str                 # "gcn" | "gin" | "sage" | "pna"
    num_layers: int
    in_dim: int
    hidden_dim: int
    out_dim: int              # MLP head output (task dim)
    pools: List[str] = field(default_factory=lambda: ["add", "mean", "max"])
    skip: bool = True
    activation: str = "relu"
    mlp_hidden_dim: int = 64
    mlp_hidden_layers: int = 2
    gin_eps: float = 0.0
    pna_delta: float = 1.0      # PyG avg_deg['log'] of the (synthetic) dataset
    # dataset shape
    mu_nodes: float = 0.0
    mu_edges: float = 0.0
    n_graphs: int = 0
    max_nodes: int = 600
    max_edges: int = 600
    seed: int = 0
    # large-graph workloads
    large_nodes: int = 0
    large_avg_degree: int = 0

    @property
    def gnn_output_dim(self) -> int:
        return self.hidden_dim


C1 = Workload("c1_gcn_esol", "gcn", 3, 9, 64, 1, pools=["mean"], mu_nodes=13, mu_edges=27,
              n_graphs=1000, seed=1)
C2 = Workload("c2_gin_qm9", "gin", 3, 11, 128, 19, mu_nodes=18, mu_edges=38,
              n_graphs=1_000_000, seed=2)
C3 = Workload("c3_sage_hiv", "sage", 3, 9, 128, 2, mu_nodes=26, mu_edges=55,
              n_graphs=100_000, seed=3)
C4 = Workload("c4_pna_lipo", "pna", 3, 9, 80, 1, mu_nodes=27, mu_edges=59,
              n_graphs=100_000, seed=4, pna_delta=1.1147)
C5 = Workload("c5_gcn_large", "gcn", 2, 128, 128, 128, mu_nodes=0, mu_edges=0, n_graphs=1,
              seed=5, large_nodes=2_000_000, large_avg_degree=16,
              max_nodes=2_000_000, max_edges=40_000_000)

WORKLOADS = {w.name: w for w in (C1, C2, C3, C4, C5)}
