"""Launch profile helper for the configs[4] encode pass: python tools/stress_probe.py [molecules=512] [passes=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from moleculesde_b200 import graph as G
from moleculesde_b200.data import Batch, synth_molecules
from moleculesde_b200.gnn import GNN
from moleculesde_b200.schnet import SchNet
M = int(sys.argv[1]) if len(sys.argv) > 1 else 512
P = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
hb = Batch.from_data_list(synth_molecules(M, 9000, "drug"))
torch.manual_seed(1)
gnn = GNN(5, 300, JK="last", drop_ratio=0.0, gnn_type="GIN").to(dev).eval()
sch = SchNet(hidden_channels=300, num_filters=128, num_interactions=6, num_gaussians=51, cutoff=10, readout="mean", node_class=119).to(dev).eval()
model = bench.make_model(dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
with torch.no_grad():
    for p in range(P):
        b = hb.to(dev)
        torch.cuda.synchronize()
        ev[0].record()
        csr = G.extend_graph(b.edge_index, b.batch, b.num_graphs)
        b.extended_edge_index, b._molsde_ext_csr = csr.edge_index, csr
        ev[1].record()
        h2d = gnn(b.x, b.edge_index, b.edge_attr)
        ev[2].record()
        out3d, h3d = sch(b.x[:, 0].contiguous(), b.positions, b.batch, return_latent=True)
        ev[3].record()
        t = torch.full((b.positions.size(0),), 0.5, device=dev)
        score = model.get_score(h2d, b, b.positions, None, t)
        ev[4].record()
        torch.cuda.synchronize()
        print(f"pass {p}: extend {ev[0].elapsed_time(ev[1]):.2f} ms | GIN {ev[1].elapsed_time(ev[2]):.2f} | SchNet {ev[2].elapsed_time(ev[3]):.2f} | "
              f"get_score {ev[3].elapsed_time(ev[4]):.2f} | total {ev[0].elapsed_time(ev[4]):.2f}")
