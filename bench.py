#!/usr/bin/env python
"""Headline benchmark: graphs/s of whole-model GNN inference on QM9-shaped graphs.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...            # the reference's CPU C++ testbench

The one JSON line carries the headline (C2) at the top level and, under "extra", the other two legs
of BASELINE.json's metric measured in the same invocation: `c4_pna_lipo` (PNA graphs/s, 100k
Lipophilicity-shaped graphs per GPU) and `c5_gcn_large` (edges/s of the 2M-node power-law GCN:
layerwise kernels at N = 1, 1D row partition + per-layer halo exchange at N > 1; aggregation
against the HBM roofline; full-size parity against the reference's own gcn_conv rows).

Workload = BASELINE.json configs[1]: GIN 3-layer hidden=128 (sum aggregation, skip connections,
add|mean|max pooling, 384->64->64->19 MLP head) on 1M QM9-shaped synthetic graphs (~18 nodes,
~38 directed edges, 11 features) PER GPU (independent graphs shard across ranks with no
collective => weak scaling).  A step = one pass of the whole hot path (tables -> 3 convs ->
pooling -> head) over the rank's batch.

  value      whole-job graphs/s with inputs resident in HBM, CUDA events on the library's stream,
             max over ranks.  The batch (1.1 GB) is larger than L2 (126 MB): no flush needed.
  e2e        the same through the public host-buffer API (Engine.run): pinned host inputs,
             H2D + kernels + D2H inside the timed region.
  roofline   dominant kernel class, timed live with CUDA events (gnnb_model_profile_read).
  cpu_baseline  the reference's own generated <name>_top (oracle/_ref, g++ -O3) on one host
             core over a bounded prefix of the same batch.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "graphs_per_sec"
UNIT = "graphs/s"


def env_int(name, default):
    return int(os.environ.get(name, default))


def bind_to_gpu_numa(device_index: int):
    """Pin this process to the CPU cores next to its GPU (sysfs local_cpulist of the PCI device),
    so that the pinned host buffers it allocates -- and the copies out of them -- stay on the
    GPU's own NUMA node when several ranks share the host.  Best effort: silently skipped when
    the topology cannot be read."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = Path(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist")
        cpus = set()
        for part in path.read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return 0


def algorithmic_flops_per_batch(w, batch) -> float:
    """SURVEY 8(d) formulas, evaluated on the actual batch."""
    n, e, g = batch.total_nodes, batch.total_edges, batch.n_graphs
    fl = 0.0
    for layer in range(w.num_layers):
        fi = w.in_dim if layer == 0 else w.hidden_dim
        fo = w.hidden_dim
        if w.conv == "gcn":
            fl += 2.0 * (e + n) * fi + 2.0 * n * fi * fo
        elif w.conv == "gin":
            fl += 2.0 * e * fi + 2.0 * n * (fi * fo + fo * fo)
        elif w.conv == "sage":
            fl += 2.0 * e * fi + 4.0 * n * fi * fo
        elif w.conv == "pna":
            fl += 2.0 * e * 2 * fi * fi + 10.0 * e * fi + 2.0 * n * (13 * fi * fo + fo * fo)
    fl += float(len(w.pools)) * n * w.hidden_dim
    dims = [w.hidden_dim * len(w.pools)] + [w.mlp_hidden_dim] * w.mlp_hidden_layers + [w.out_dim]
    fl += 2.0 * g * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    return fl


def gemm_flops_per_batch(w, batch) -> float:
    n, g = batch.total_nodes, batch.n_graphs
    fl = 0.0
    for layer in range(w.num_layers):
        fi = w.in_dim if layer == 0 else w.hidden_dim
        fo = w.hidden_dim
        fl += {"gcn": 2.0 * n * fi * fo, "gin": 2.0 * n * (fi * fo + fo * fo),
               "sage": 4.0 * n * fi * fo,
               "pna": 2.0 * n * (2 * fi * fi + 13 * fi * fo + fo * fo)}[w.conv]
    dims = [w.hidden_dim * len(w.pools)] + [w.mlp_hidden_dim] * w.mlp_hidden_layers + [w.out_dim]
    fl += 2.0 * g * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    return fl


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.dev = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.dev)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ts, line in self.lines:
            if t0 is not None and not (t0 - 0.12 <= ts <= t1 + 0.12):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ reference arm
def _ref_worker(args):
    name, params, xs, coos, nptr, eptr, repeat = args
    sys.path.insert(0, str(ROOT / "oracle"))
    from oracle import RefModel
    from gnn_builder_b200.data import GraphBatch

    ref = RefModel(name)
    ref.set_params(params)
    batch = GraphBatch(xs, coos, nptr, eptr)
    ref(*batch.graph(0))  # untimed parameter load, like the reference testbench (tb:170-172)
    t0 = time.perf_counter()
    for _ in range(repeat):
        for g in range(batch.n_graphs):
            ref(*batch.graph(g))
    return time.perf_counter() - t0, batch.n_graphs * repeat


def _port_worker(args):
    name, desc, params, xs, coos, nptr, eptr, repeat = args
    sys.path.insert(0, str(ROOT / "oracle"))
    from oracle import Oracle
    from gnn_builder_b200.data import GraphBatch

    orc = Oracle()
    batch = GraphBatch(xs, coos, nptr, eptr)
    t0 = time.perf_counter()
    for _ in range(repeat):
        orc.model_forward_batch(desc, list(params.values()), batch)
    return time.perf_counter() - t0, batch.n_graphs * repeat


def reference_available(name) -> bool:
    return (ROOT / "oracle" / "_ref" / "models" / name / f"lib{name}.so").exists()


def cpu_reference_rate(w, model, batch, n_procs: int, graphs_per_proc: int, repeat: int = 1):
    """graphs/s of the reference CPU implementation: n_procs independent single-threaded
    processes (the generated top is not re-entrant), each over its own slice of the batch."""
    import multiprocessing as mp

    params = model.named_parameter_arrays()
    kind = "reference" if reference_available(w.name) else "port"
    jobs = []
    for p in range(n_procs):
        g0 = (p * graphs_per_proc) % max(1, batch.n_graphs - graphs_per_proc)
        sl = batch.slice(g0, g0 + graphs_per_proc)
        common = (np.ascontiguousarray(sl.x), np.ascontiguousarray(sl.coo),
                  np.ascontiguousarray(sl.node_ptr), np.ascontiguousarray(sl.edge_ptr), repeat)
        if kind == "reference":
            jobs.append((w.name, params) + common)
        else:
            jobs.append((w.name, model.describe(), params) + common)
    fn = _ref_worker if kind == "reference" else _port_worker
    t0 = time.perf_counter()
    if n_procs == 1:
        res = [fn(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(n_procs) as pool:
            res = pool.map(fn, jobs)
    wall = time.perf_counter() - t0
    graphs = sum(r[1] for r in res)
    busy = max(r[0] for r in res)
    return graphs / busy, kind, graphs, busy, wall


def run_reference_arm_large(args, w):
    """Reference arm of the large-graph workload: the reference's own gcn_conv<2M, 40M, 128, 128>
    template (oracle/_ref) over a bounded prefix of destination rows, one host thread (the
    template is single-threaded and not re-entrant)."""
    import gnn_builder_b200 as gnnb

    sys.path.insert(0, str(ROOT / "oracle"))
    from oracle import ref_available, ref_big_gcn_rate

    if not ref_available():
        print(json.dumps({"impl": "reference", "unavailable":
                          "oracle/_ref (the compiled reference templates) is not present"}))
        return
    n = args.nodes or w.large_nodes
    model = gnnb.build_model(w, seed=0)
    x, coo = gnnb.make_powerlaw_graph(n, w.large_avg_degree, w.in_dim, seed=w.seed)
    P = model.named_parameter_arrays()
    Wm, bm = P["gnn_convs_0_conv_lin_weight"], P["gnn_convs_0_conv_bias"]
    rows = min(n, 40000)
    rates = []
    for step in range(args.warmup + args.steps):
        rate, secs, edges_done, t_tab, _ = ref_big_gcn_rate(x, coo, n, Wm, bm, rows)
        if step >= args.warmup:
            rates.append((edges_done, secs))
    value = sum(e for e, _ in rates) / sum(t for _, t in rates)
    sample = (f"reference gcn_conv<2000000,40000000,128,128> over the first {rows} destination rows "
              f"of the same graph per step (tables rebuilt each step, untimed), 1 host thread")
    E = int(coo.shape[0])
    print(json.dumps({
        "impl": "reference", "metric": "edges_per_sec", "value": value, "unit": "edges/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(t for _, t in rates) / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{w.name}: GCN {w.num_layers}-layer hidden={w.hidden_dim}, power-law "
                               f"graph N={n} E={E} F={w.in_dim}; value = E x layers / step time"},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": 1, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def run_reference_arm(args, w):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    if w.large_nodes:
        run_reference_arm_large(args, w)
        return
    import gnn_builder_b200 as gnnb

    model = gnnb.build_model(w, pna_delta=w.pna_delta, seed=0)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    n_procs = max(1, min(cores, 64))
    per_proc = args.ref_graphs
    batch = gnnb.make_molecular_batch(max(per_proc * 4, per_proc + 1), w.mu_nodes, w.mu_edges,
                                      w.in_dim, seed=w.seed)
    rates = []
    kind = "port"
    for step in range(args.warmup + args.steps):
        rate, kind, graphs, busy, wall = cpu_reference_rate(w, model, batch, n_procs, per_proc)
        if step >= args.warmup:
            rates.append((graphs, busy))
    total_graphs = sum(g for g, _ in rates)
    total_time = sum(t for _, t in rates)
    value = total_graphs / total_time
    sample = (f"{n_procs} single-threaded processes x {per_proc} graphs per step, "
              f"{'reference <name>_top compiled g++ -O3 (oracle/_ref)' if kind == 'reference' else 'oracle C port'}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_time / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, args, per_gpu_graphs=args.graphs),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_procs, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(w, args, per_gpu_graphs):
    shape = {"c1_gcn_esol": "ESOL", "c2_gin_qm9": "QM9", "c3_sage_hiv": "HIV",
             "c4_pna_lipo": "Lipophilicity"}.get(w.name, w.name)
    in_mb = per_gpu_graphs * (w.mu_nodes * w.in_dim * 4 + w.mu_edges * 8 + 16) / 1e6
    return {
        "workload": f"{w.name}: {w.conv.upper()} {w.num_layers}-layer hidden={w.hidden_dim} "
                    f"pools={'|'.join(w.pools)} head {w.hidden_dim * len(w.pools)}->"
                    f"{w.mlp_hidden_dim}x{w.mlp_hidden_layers}->{w.out_dim}, {shape}-shaped graphs "
                    f"(mu_nodes={w.mu_nodes}, mu_edges={w.mu_edges}, feats={w.in_dim})",
        "graphs_per_gpu": per_gpu_graphs,
        "sharding": "independent graphs per rank, no collective",
        "l2_policy": f"inputs (~{in_mb:.0f} MB per GPU) larger than the 126 MB L2; no flush"
                     if in_mb > 126 else
                     f"inputs (~{in_mb:.0f} MB per GPU) fit the 126 MB L2 and are NOT flushed: "
                     f"a non-default batch size, not a valid bench number",
    }


# ------------------------------------------------------------------------------ shared plumbing
class Ctx:
    """process-group plumbing of one bench invocation (one process per GPU)"""

    def __init__(self):
        import torch

        self.torch = torch
        self.rank, self.world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
        self.local_rank = env_int("LOCAL_RANK", 0)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.numa_cpus = bind_to_gpu_numa(self.local_rank) if self.world > 1 else 0
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        peaks_fp = ROOT / "MEASURED_PEAKS.json"
        self.peaks = json.loads(peaks_fp.read_text()) if peaks_fp.exists() else {}
        self.hbm_peak = float(self.peaks.get("hbm_gbs", 6650.0))
        self.tensor_peak = float(self.peaks.get("bf16_tflops_sustained",
                                                self.peaks.get("bf16_tflops", 1590.0)))
        self.peak_src = ("measured (MEASURED_PEAKS.json)" if self.peaks
                         else "fallback (B200_PROFILING.md)")

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, v: float, op) -> float:
        if self.dist is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v: float) -> float:
        return self._reduce(v, self.dist.ReduceOp.MAX) if self.dist else v

    def sum(self, v: float) -> float:
        return self._reduce(v, self.dist.ReduceOp.SUM) if self.dist else v

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def measured_tf32_peak(torch, seconds: float = 1.0):
    """cuBLAS TF32 matmul throughput (TFLOP/s), measured the way MEASURED_PEAKS.json measures bf16:
    8192^3, best of a few back-to-back repetitions.  The denominator of `frac_of_tf32_peak`."""
    try:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        n = 8192
        a = torch.randn(n, n, device="cuda")
        b = torch.randn(n, n, device="cuda")
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 0.0
        t_end = time.time() + seconds
        while time.time() < t_end:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                a @ b
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 4 * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        torch.backends.cuda.matmul.allow_tf32 = prev
        del a, b
        return best
    except Exception:
        return None


# ------------------------------------------------------------------------------ molecular batches
def measure_molecular(ctx: Ctx, w, graphs_per_gpu: int, steps: int, warmup: int, path: str,
                      cpu_graphs: int, with_tf32_peak: bool = False):
    """graphs/s of one molecular workload (C1-C4 shapes): device-resident `value`, end-to-end `e2e`
    (pinned and pageable host buffers), live roofline, CPU baseline (rank 0, N = 1)."""
    torch = ctx.torch
    import gnn_builder_b200 as gnnb

    rank, world, local_rank = ctx.rank, ctx.world, ctx.local_rank
    model = gnnb.build_model(w, pna_delta=w.pna_delta, seed=0)
    batch = gnnb.make_molecular_batch(graphs_per_gpu, w.mu_nodes, w.mu_edges, w.in_dim,
                                      seed=w.seed + 1000 * rank, max_nodes=w.max_nodes)
    G, T, E = batch.n_graphs, batch.total_nodes, batch.total_edges
    max_n = int(np.diff(batch.node_ptr).max())
    max_e = int(np.diff(batch.edge_ptr).max())
    eng = gnnb.Engine(model, max_nodes=max_n, max_edges=max_e, device=local_rank,
                      path={"auto": gnnb.PATH_AUTO, "fused": gnnb.PATH_FUSED,
                            "layerwise": gnnb.PATH_LAYERWISE}[path])

    # pinned host copies (e2e) and device-resident copies (value)
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t

    hx, hcoo = pinned(batch.x), pinned(batch.coo)
    hn, he = pinned(batch.node_ptr), pinned(batch.edge_ptr)
    hout = torch.empty((G, w.out_dim), dtype=torch.float32, pin_memory=True)
    host_batch = gnnb.GraphBatch(hx.numpy(), hcoo.numpy(), hn.numpy(), he.numpy())
    dx, dcoo = hx.cuda(non_blocking=True), hcoo.cuda(non_blocking=True)
    dn, de = hn.cuda(non_blocking=True), he.cuda(non_blocking=True)
    dout = torch.empty((G, w.out_dim), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(eng.stream)

    def step_device():
        eng.run_device(dx, dcoo, dn, de, dout, G, T, E)

    # ---- device-resident throughput (`value`)
    for _ in range(warmup):
        step_device()
    eng.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(1.2)  # nvidia-smi needs about a second before its first sample
    ctx.barrier()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            step_device()
        ev1.record(stream)
    eng.synchronize()
    launches_per_step = eng.last_launches      # (kernels of the last timed step: our own count)
    path_used = eng.last_kernel
    ctx.barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = ctx.max(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / steps
    total_graphs = ctx.sum(float(G))
    value = total_graphs / (ms_per_step * 1e-3)

    # ---- end-to-end through the public host-buffer API (`e2e`): pinned host buffers ...
    for _ in range(min(warmup, 2)):
        eng.run(host_batch, out=hout.numpy())
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.run(host_batch, out=hout.numpy())
    torch.cuda.synchronize()
    e2e_s = ctx.max(time.perf_counter() - t0) / steps
    ctx.barrier()
    h2d = int(batch.x.nbytes + batch.coo.nbytes + batch.node_ptr.nbytes + batch.edge_ptr.nbytes)
    d2h = int(G * w.out_dim * 4)
    e2e_value = total_graphs / e2e_s
    # ... and ordinary (pageable) numpy arrays, what a Project user passes
    out_pageable = np.empty((G, w.out_dim), np.float32)
    eng.run(batch, out=out_pageable)
    ctx.barrier()
    t0 = time.perf_counter()
    pg_steps = max(1, min(steps, 2))
    for _ in range(pg_steps):
        eng.run(batch, out=out_pageable)
    torch.cuda.synchronize()
    e2e_pg_s = ctx.max(time.perf_counter() - t0) / pg_steps
    ctx.barrier()
    # ... and the same numpy arrays after Engine.pin_batch (cudaHostRegister in place, once)
    t0 = time.perf_counter()
    eng.pin_batch(batch)
    eng.pin(out_pageable)
    register_s = time.perf_counter() - t0
    eng.run(batch, out=out_pageable)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(pg_steps):
        eng.run(batch, out=out_pageable)
    torch.cuda.synchronize()
    e2e_reg_s = ctx.max(time.perf_counter() - t0) / pg_steps
    eng.unpin()
    ctx.barrier()

    # ---- live per-kernel-class timing for the roofline
    eng.set_profile(True)
    prof_steps = 2
    for _ in range(prof_steps):
        step_device()
    prof = eng.read_profile()
    eng.set_profile(False)
    dominant = max(prof, key=lambda k: prof[k]["ms"])
    dom_ms = prof[dominant]["ms"] / prof_steps
    alg_bytes = batch.algorithmic_bytes(w.out_dim)
    alg_flops = algorithmic_flops_per_batch(w, batch)
    sm_mhz = clocks.get("sm_mhz") or float(ctx.peaks.get("sm_max_mhz", 1965.0))
    fp32_peak_tflops = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    traffic_db = {}
    tfp = ROOT / "profiles" / "roofline_traffic.json"
    if tfp.exists():
        traffic_db = json.loads(tfp.read_text())
    if dominant in ("gemm", "fused"):
        fl = gemm_flops_per_batch(w, batch) if dominant == "gemm" else alg_flops
        achieved = fl / (dom_ms * 1e-3) / 1e12
        on_tensor = path_used == "fused-tcgen05"
        roofline = {"bound": "tensor", "kernel": dominant, "achieved": achieved,
                    "peak": ctx.tensor_peak, "unit": "TFLOP/s", "frac": achieved / ctx.tensor_peak,
                    "traffic": None, "peak_source": ctx.peak_src + ", bf16 sustained",
                    "kernel_ms": dom_ms,
                    "note": ("molecular graphs are compute/latency bound, not HBM bound (SURVEY 8d). "
                             "achieved = ALGORITHMIC fp32 FLOPs / kernel time; see DESIGN.md 5.1 for "
                             "how many tensor-core MMAs one algorithmic MAC costs at fp32-grade "
                             "accuracy") if on_tensor else
                            ("molecular graphs are compute/latency bound, not HBM bound (SURVEY "
                             "8d); the node transforms of this path are the dominant kernel class"),
                    "fp32_fma_peak_tflops": fp32_peak_tflops,
                    "frac_of_fp32_fma_peak": achieved / fp32_peak_tflops}
        if on_tensor:
            # the fused kernel's node transforms are bf16x2: three kind::f16 MMAs per product
            gemm_fl = gemm_flops_per_batch(w, batch)
            executed = 3.0 * gemm_fl / (dom_ms * 1e-3) / 1e12
            roofline["executed_bf16_mma_tflops"] = executed
            roofline["executed_frac_of_bf16_peak"] = executed / ctx.tensor_peak
            roofline["executed_note"] = ("node-transform MMAs only (3 bf16 MMAs of K = 16 per product); the "
                                         "dense 128x128 tile aggregation MMAs are not counted")
        elif dominant == "gemm":
            # layerwise tcgen05 GEMM: 3xTF32 (three kind::tf32 MMAs per product)
            executed = 3.0 * fl / (dom_ms * 1e-3) / 1e12
            roofline["executed_tf32_mma_tflops"] = executed
            if with_tf32_peak:
                tf32 = measured_tf32_peak(torch)
                if tf32:
                    roofline["tf32_peak_measured_tflops"] = tf32
                    roofline["frac_of_tf32_peak"] = executed / tf32
                    roofline["tf32_peak_source"] = "cuBLAS TF32 8192^3 matmul, measured in this run"
        key = f"{w.name}:{path_used}:{G}"
        if key in traffic_db:
            roofline["traffic"] = traffic_db[key]["dram_bytes_per_launch"]
            roofline["traffic_source"] = "static: " + traffic_db[key]["source"]
            for k2 in ("tensor_pipe_active_pct", "issue_active_pct"):
                if k2 in traffic_db[key]:
                    roofline["static_ncu_" + k2] = traffic_db[key][k2]
    else:
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": ctx.hbm_peak,
                    "unit": "GB/s", "frac": achieved / ctx.hbm_peak, "traffic": None,
                    "peak_source": ctx.peak_src, "kernel_ms": dom_ms}
    roofline["hbm_algorithmic_gbs_whole_step"] = alg_bytes / (ms_per_step * 1e-3) / 1e9
    roofline["class_ms_per_step"] = {k: v["ms"] / prof_steps for k, v in prof.items()}

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and cpu_graphs > 0:
        n_cpu = min(cpu_graphs, G)
        rate, kind, graphs, busy, _ = cpu_reference_rate(w, model, batch, 1, n_cpu)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": 1, "kind": kind,
                        "sample": f"first {n_cpu} graphs of the same batch, one pass, "
                                  f"{busy:.1f} s on one host core"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(w, None, G), path=path_used,
                       nodes_per_gpu=T, edges_per_gpu=E,
                       **({"host_binding": f"each rank pinned to its GPU's NUMA-local cores "
                                           f"({ctx.numa_cpus} on rank 0)"} if ctx.numa_cpus else {})),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                "host_memory": "pinned",
                "h2d_gbs_per_gpu": h2d / e2e_s / 1e9,
                "pageable": {"value": total_graphs / e2e_pg_s, "unit": UNIT,
                             "ms_per_step": e2e_pg_s * 1e3, "h2d_gbs_per_gpu": h2d / e2e_pg_s / 1e9,
                             "note": "the same call on ordinary numpy arrays: the copies go "
                                     "through the driver's staging buffers"},
                "registered": {"value": total_graphs / e2e_reg_s, "unit": UNIT,
                               "ms_per_step": e2e_reg_s * 1e3, "register_ms_once": register_s * 1e3,
                               "note": "the same numpy arrays after Engine.pin_batch "
                                       "(cudaHostRegister in place, paid once)"}},
        "gpu_launches": int(launches_per_step * steps),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
    }
    eng.close()
    del dx, dcoo, dn, de, dout, hx, hcoo, hn, he, hout
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------------------------------------ large graph (C5)
def _large_graph(ctx: Ctx, w, n: int):
    """the synthetic power-law graph through the on-disk cache: rank 0 generates, everyone maps"""
    import gnn_builder_b200 as gnnb

    if ctx.rank == 0:
        gnnb.cached_powerlaw_graph(n, w.large_avg_degree, w.in_dim, w.seed)
    ctx.barrier()
    return gnnb.cached_powerlaw_graph(n, w.large_avg_degree, w.in_dim, w.seed, generate=False)


def measure_large(ctx: Ctx, w, nodes: int, steps: int, warmup: int, cpu_baseline_on: bool,
                  transport: str = "auto"):
    """BASELINE configs[4]: GCN 2-layer hidden=128 on one power-law graph (2M nodes, avg in-degree
    16).  N = 1: the layerwise kernels through the model handle, aggregation timed live for the
    HBM roofline.  N > 1: 1D row partition + per-layer halo exchange overlapped with the
    aggregation of the owned-source edges.  Parity at full size, untimed, on every N: the first
    40 000 rows of conv layer 1 against the reference's own gcn_conv<2000000,40000000,128,128>."""
    torch = ctx.torch
    import dataclasses

    import gnn_builder_b200 as gnnb
    from gnn_builder_b200.distributed import LargeGraphGCN, RowPartition

    rank, world, local_rank, dist = ctx.rank, ctx.world, ctx.local_rank, ctx.dist
    n = nodes or w.large_nodes
    n -= n % world
    model = gnnb.build_model(w, seed=0)
    P = model.named_parameter_arrays()
    x, coo = _large_graph(ctx, w, n)
    E, F, L = int(coo.shape[0]), w.in_dim, w.num_layers
    agg_bytes_layer = E * (4 * F + 8) + n * (4 * F + 4 * F + 8)
    sampler = ClockSampler(local_rank)
    roofline, launches, cpu_baseline, parity, exchange = None, 0, None, None, None
    PAR_ROWS = min(n // world, 40000)

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(np.asarray(a[:0])).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t

    if world == 1:
        eng = gnnb.Engine(model, device=local_rank, path=gnnb.PATH_LAYERWISE)
        hx, hcoo = pinned(x), pinned(coo)
        dx, dcoo = hx.cuda(), hcoo.cuda()
        dn = torch.tensor([0, n], dtype=torch.int64, device="cuda")
        de = torch.tensor([0, E], dtype=torch.int64, device="cuda")
        dout = torch.empty((1, w.out_dim), device="cuda")
        step = lambda: eng.run_device(dx, dcoo, dn, de, dout, 1, n, E)  # noqa: E731
        stream = torch.cuda.ExternalStream(eng.stream)
        sync = eng.synchronize
    else:
        part = RowPartition(n, world)
        r0, r1 = part.rows(rank)
        runner = LargeGraphGCN(model, n, rank, world, dist=dist, transport=transport)
        runner.setup(part.local_edges(np.asarray(coo), rank))
        hx = pinned(x[r0:r1])
        runner.input_view().copy_(hx)
        step = lambda: runner.forward(None)  # noqa: E731
        stream = torch.cuda.current_stream()
        sync = torch.cuda.synchronize
        # our kernels per forward: per layer pack (+ signal/wait on the p2p transport), the two
        # CSR parts (light rows + sliced heavy rows + combine each), weight prep (2) and the node
        # transform; pooling (2) and the head's linears once
        launches = L * (1 + (2 if runner.transport == "p2p" else 0) + 2 * 3 + 3) + 2 + \
            model.describe()["mlp_num_linear"]
    for _ in range(warmup):
        step()
    sync()
    sampler.start()
    time.sleep(1.2)  # nvidia-smi needs about a second before its first sample
    ctx.barrier()
    t0w = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            step()
        ev1.record(stream)
    sync()
    torch.cuda.synchronize()
    if world == 1:
        launches = eng.last_launches
    t1w = time.time()
    clocks = sampler.stop(t0w, t1w)
    ms_per_step = ctx.max(ev0.elapsed_time(ev1) / steps)
    value = E * L / (ms_per_step * 1e-3)
    e2e = None
    reps = max(1, steps // 2)
    if world > 1:
        runner.check_transport()
        # end to end per rank: pinned host feature shard -> device, forward, result back to host
        for _ in range(2):
            runner.input_view().copy_(hx, non_blocking=True)
            runner.forward(None).cpu()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            runner.input_view().copy_(hx, non_blocking=True)
            out_h = runner.forward(None).cpu()
        torch.cuda.synchronize()
        e2e_s = ctx.max((time.perf_counter() - t0) / reps)
        e2e = {"value": E * L / e2e_s, "unit": "edges/s",
               "h2d_bytes_per_step": int(hx.numel() * 4 * world),
               "d2h_bytes_per_step": int(out_h.numel() * 4 * world), "ms_per_step": e2e_s * 1e3,
               "host_memory": "pinned",
               "note": "feature shards H2D from pinned memory on every rank each step; the CSR "
                       "slices / halo plan are built once (setup) and stay resident"}
        # ---- what the exchange costs by itself, and the compute by itself (untimed region)
        st = runner.stats
        Fb = 4 * F
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        iters = 5
        for _ in range(2):
            runner._exchange(0, F); runner._arrived()
        ctx.barrier()
        ev[0].record()
        for _ in range(iters):
            runner._exchange(0, F); runner._arrived()
        ev[1].record()
        torch.cuda.synchronize()
        ctx.barrier()
        B, plan = runner.backend, runner.plan
        W0, b0 = runner.layers[0]
        x_ext = runner.ext[runner._buf(0)][: plan.n_ext * F]
        y_loc = runner.ext[runner._buf(1)][: plan.n_local * F]
        ev[2].record()
        for _ in range(iters):
            B.gcn_layer_halo(x_ext, y_loc, plan, runner.dinv_ext, W0, b0, None, 1, 3, F, F)
        ev[3].record()
        torch.cuda.synchronize()
        xchg_ms = ctx.max(ev[0].elapsed_time(ev[1]) / iters)
        comp_ms = ctx.max(ev[2].elapsed_time(ev[3]) / iters)
        recv_bytes = ctx.max(float(st["halo_rows"] * Fb))
        exchange = {
            "transport": runner.transport, "transport_autotune_ms": st.get("autotune_ms"),
            "halo_rows_max_rank": int(ctx.max(float(st["halo_rows"]))),
            "halo_frac_of_remote_rows": ctx.max(float(st["halo_frac_of_remote_rows"])),
            "recv_bytes_per_layer_max_rank": int(recv_bytes),
            "full_allgather_bytes_per_layer": int((n - n // world) * Fb),
            "exchange_ms_per_layer_alone": xchg_ms,
            "exchange_gbs_in_per_gpu": recv_bytes / (xchg_ms * 1e-3) / 1e9,
            "nvlink_peak_gbs": 770.0, "nvlink_peak_source": "B200_PROFILING.md (measured peer copy)",
            "compute_ms_per_layer_alone": comp_ms,
            "overlap": "exchange on a second stream under the aggregation of the owned-source edges",
            "hub_rows_l2_resident": int(st["hub_rows"]),
            "limiter": ("exchange" if xchg_ms > comp_ms else "compute"),
        }
    else:
        hbatch = gnnb.GraphBatch(hx.numpy(), hcoo.numpy(), np.array([0, n], np.int64),
                                 np.array([0, E], np.int64))
        hout = torch.empty((1, w.out_dim), dtype=torch.float32, pin_memory=True)
        eng.run(hbatch, out=hout.numpy())
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.run(hbatch, out=hout.numpy())
        e2e_s = (time.perf_counter() - t0) / reps
        e2e = {"value": E * L / e2e_s, "unit": "edges/s",
               "h2d_bytes_per_step": int(x.nbytes + coo.nbytes + 32),
               "d2h_bytes_per_step": int(w.out_dim * 4), "ms_per_step": e2e_s * 1e3,
               "host_memory": "pinned", "h2d_gbs": (x.nbytes + coo.nbytes) / e2e_s / 1e9}
        eng.set_profile(True)
        for _ in range(2):
            step()
        prof = eng.read_profile()
        eng.set_profile(False)
        agg_ms = prof["aggregate"]["ms"] / 2 / L
        achieved = agg_bytes_layer / (agg_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm",
                    "kernel": "aggregate (agg_rows_kernel + agg_heavy_chunk_kernel + agg_heavy_combine_kernel)",
                    "achieved": achieved, "peak": ctx.hbm_peak, "unit": "GB/s",
                    "frac": achieved / ctx.hbm_peak, "traffic": None,
                    "kernel_ms": agg_ms, "algorithmic_bytes_per_layer": agg_bytes_layer,
                    "peak_source": ctx.peak_src,
                    "class_ms_per_step": {k: v["ms"] / 2 for k, v in prof.items()},
                    "note": "achieved uses SURVEY 8(d)'s per-edge gather model (no cache reuse), so "
                            "it can exceed the HBM peak when neighbor rows hit in the 126 MB L2; "
                            "`traffic` is what actually crossed the HBM interface (ncu, static)"}
        tfp = ROOT / "profiles" / "roofline_traffic.json"
        if tfp.exists():
            tdb = json.loads(tfp.read_text())
            key = f"{w.name}:layerwise:{n}"
            if key in tdb:
                roofline["traffic"] = tdb[key]["dram_bytes_per_launch"]
                roofline["traffic_source"] = "static: " + tdb[key]["source"]
                roofline["dram_gbs_measured"] = tdb[key]["dram_bytes_per_launch"] / (agg_ms * 1e-3) / 1e9
                roofline["dram_frac_of_peak"] = roofline["dram_gbs_measured"] / ctx.hbm_peak

    # ---- full-size parity (untimed, every N) + CPU baseline: the reference's own gcn_conv
    ref_ok = False
    if rank == 0:
        sys.path.insert(0, str(ROOT / "oracle"))
        from oracle import ref_available, ref_big_gcn_rate

        ref_ok = ref_available()
    if world > 1:
        runner.forward(None, capture=(0, PAR_ROWS))
        torch.cuda.synchronize()
        got = runner.captured.cpu().numpy() if rank == 0 else None
    else:
        # conv layer 1 alone = a 1-layer model with the same parameters
        d1 = dict(model.describe(), num_layers=1)
        p1 = {k: v for k, v in P.items() if not k.startswith("gnn_convs_1_")}
        with gnnb.Engine(desc=d1, params=p1, device=local_rank, path=gnnb.PATH_LAYERWISE) as e1:
            e1.run_device_sync(dx, dcoo, dn, de, dout, 1)
            got = e1.node_embeddings(n)[:PAR_ROWS]
    if rank == 0 and ref_ok:
        Wm, bm = P["gnn_convs_0_conv_lin_weight"], P["gnn_convs_0_conv_bias"]
        rate, secs, edges_done, t_tab, y = ref_big_gcn_rate(np.asarray(x), np.asarray(coo), n, Wm, bm,
                                                           PAR_ROWS)
        ref = np.maximum(y[:PAR_ROWS], 0.0)          # the template applies ReLU after the conv (cpp:164)
        err = float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))
        parity = {"n_checked": int(PAR_ROWS), "max_rel_err": err, "tolerance": 1e-4,
                  "ok": bool(err <= 1e-4),
                  "against": "reference gcn_conv<2000000,40000000,128,128> (oracle/_ref, g++ -O3) + ReLU, "
                             "rows [0, n_checked) of conv layer 1 on the full graph"}
        if world == 1 and cpu_baseline_on:
            cpu_baseline = {"value": rate, "unit": "edges/s", "cores": 1, "kind": "reference",
                            "sample": f"reference gcn_conv<2000000,40000000,128,128> over the "
                                      f"first {PAR_ROWS} destination rows ({edges_done} edges, "
                                      f"{secs:.1f} s) of the same graph; tables {t_tab:.1f} s"}
    elif rank == 0:
        parity = {"n_checked": 0, "max_rel_err": None, "ok": None,
                  "against": "oracle/_ref is not present on this box"}
    ctx.barrier()
    line = {"metric": "edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{w.name}: GCN {L}-layer hidden={w.hidden_dim}, power-law "
                                   f"graph N={n} E={E} F={F}; value = E x layers / step time",
                       "partition": "single GPU" if world == 1 else
                       f"1D row partition over {world} GPUs, per-layer halo exchange "
                       f"({runner.transport})",
                       "l2_policy": "feature matrix (1 GB) larger than L2; no flush"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches * steps),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
            "exchange": exchange}
    if world == 1:
        eng.close()
        del dx, dcoo
    else:
        runner.close()
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------------------------------------ our arm
class Deadline:
    """If the extras overrun their budget (a hung collective, a dead peer), rank 0 still prints the
    headline line -- with the reason under `extra` -- and every rank exits."""

    def __init__(self, seconds: float, line_fn, rank: int):
        self.line_fn, self.rank = line_fn, rank
        self.timer = threading.Timer(seconds, self.fire)
        self.timer.daemon = True
        self.timer.start()

    def fire(self):
        if self.rank == 0:
            print(json.dumps(self.line_fn("deadline exceeded")), flush=True)
        os._exit(0)

    def cancel(self):
        self.timer.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gnnb", choices=["gnnb", "reference"])
    ap.add_argument("--workload", default="c2_gin_qm9")
    ap.add_argument("--graphs", type=int, default=0, help="graphs per GPU (default: the config's)")
    ap.add_argument("--path", default="auto", choices=["auto", "fused", "layerwise"])
    ap.add_argument("--ref-graphs", type=int, default=400,
                    help="graphs per process per step of the reference arm")
    ap.add_argument("--cpu-baseline-graphs", type=int, default=8000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nodes", type=int, default=0, help="large-graph workloads: node count")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline workload only (skip extra.c4_pna_lipo / extra.c5_gcn_large)")
    ap.add_argument("--transport", default="auto", choices=["auto", "p2p", "nccl"],
                    help="halo exchange of the row-partitioned large graph")
    ap.add_argument("--extras-deadline", type=float, default=540.0)
    args = ap.parse_args()

    from gnn_builder_b200.configs import WORKLOADS

    w = WORKLOADS[args.workload]
    if args.graphs <= 0:
        args.graphs = w.n_graphs
    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    ctx = Ctx()
    cpu_graphs = 0 if args.no_cpu_baseline else args.cpu_baseline_graphs
    if w.large_nodes:
        line = measure_large(ctx, w, args.nodes, args.steps, args.warmup, not args.no_cpu_baseline,
                             args.transport)
        if ctx.rank == 0:
            print(json.dumps(line), flush=True)
        ctx.close()
        return

    line = measure_molecular(ctx, w, args.graphs, args.steps, args.warmup, args.path, cpu_graphs,
                             with_tf32_peak=True)
    extra = {}
    if args.workload == "c2_gin_qm9" and not args.no_extras:
        def partial_line(reason):
            e = dict(extra)
            e["error"] = reason
            return dict(line, extra=e)

        guard = Deadline(args.extras_deadline, partial_line, ctx.rank)
        for name, fn in (
                ("c4_pna_lipo", lambda: measure_molecular(
                    ctx, WORKLOADS["c4_pna_lipo"], WORKLOADS["c4_pna_lipo"].n_graphs,
                    args.steps, args.warmup, "auto", min(cpu_graphs, 1000))),
                ("c5_gcn_large", lambda: measure_large(
                    ctx, WORKLOADS["c5_gcn_large"], args.nodes, args.steps, args.warmup,
                    not args.no_cpu_baseline, args.transport))):
            try:
                extra[name] = fn()
            except Exception as e:   # the headline must survive a failing extra
                import traceback

                extra[name] = {"error": f"{type(e).__name__}: {e}",
                               "trace": traceback.format_exc()[-1500:]}
                if ctx.world > 1:   # the other ranks may be inside a collective: do not wait for them
                    guard.fire()
        guard.cancel()
    if ctx.rank == 0:
        line["extra"] = extra
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
