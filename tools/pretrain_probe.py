#!/usr/bin/env python
"""Time the pretraining iteration (BASELINE.json configs[2]: batch 256 synthetic PCQM4Mv2-shaped molecules per GPU).
Prints per-phase wall/device times; used to decide what to fuse next.  `python tools/pretrain_probe.py [B] [steps]`"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from moleculesde_b200 import graph as G  # noqa: E402
from moleculesde_b200.data import Batch, synth_molecules  # noqa: E402
from moleculesde_b200.gnn import GNN  # noqa: E402
from moleculesde_b200.pretrain import PretrainStep  # noqa: E402
from moleculesde_b200.schnet import SchNet  # noqa: E402
from moleculesde_b200.sde_2d_to_3d import SDEModel2Dto3D_02  # noqa: E402
from moleculesde_b200.sde_3d_to_2d import SDEModel3Dto2D_node_adj_dense  # noqa: E402


def build_models(seed=1):
    torch.manual_seed(seed)
    gnn = GNN(5, 300, JK="last", drop_ratio=0.0, gnn_type="GIN")
    sch = SchNet(hidden_channels=300, num_filters=128, num_interactions=6, num_gaussians=51, cutoff=10, readout="mean", node_class=119)
    m23 = SDEModel2Dto3D_02(emb_dim=300, hidden_dim=32, beta_schedule=None, beta_min=0.2, beta_max=1.0, num_diffusion_timesteps=1000,
                            SDE_type="VE", use_extend_graph=True)
    m32 = SDEModel3Dto2D_node_adj_dense(dim3D=300, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4,
                                        emb_dim=300, num_linears=3, beta_min=0.1, beta_max=1.0, num_diffusion_timesteps=1000,
                                        SDE_type="VE", num_class_X=119, noise_on_one_hot=True)
    return gnn, sch, m23, m32


def build_batch(B, seed, dev):
    mols = synth_molecules(B, seed, "pcqm")
    b = Batch.from_data_list(mols).to(dev)
    csr = G.extend_graph(b.edge_index, b.batch, b.num_graphs)
    b.extended_edge_index = csr.edge_index
    return b


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = torch.device("cuda:0")
    ps = PretrainStep(*build_models(), dev)
    b = build_batch(B, 0, dev)
    print(f"B={B} N={b.x.size(0)} E_b={b.edge_index.size(1)} E_x={b.extended_edge_index.size(1)} params={ps.store.numel}")
    if os.environ.get("PROBE_ONCE") == "1":   # under ncu: one warm-up + one step
        ps.step(b)
        torch.cuda.synchronize()
        ps.step(b)
        torch.cuda.synchronize()
        return
    for i in range(3):
        out = ps.step(b)
    torch.cuda.synchronize()
    print("losses", {k: float(v) for k, v in out.items() if k.startswith("loss") or k == "cl_loss"}, "launches/step", ps.launches)
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        ps.step(b)
    ev1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps
    print(f"step: wall {wall * 1e3:.2f} ms, device {ev0.elapsed_time(ev1) / steps:.2f} ms -> {B / wall:.0f} molecules/s")
    # the same iteration captured once into a CUDA graph and replayed (static batch): pure device time
    try:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ps.step(b)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=side):
                ps.step(b)
        torch.cuda.synchronize()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(steps):
            g.replay()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        print(f"graph replay: {ms:.2f} ms/step -> {B / ms * 1e3:.0f} molecules/s")
    except Exception as e:  # noqa: BLE001
        print("graph capture failed:", repr(e)[:400])


if __name__ == "__main__":
    main()
