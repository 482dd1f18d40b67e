"""Where does the eager / end-to-end pretraining step spend its host time?  cProfile over 5 e2e steps (batch 256)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from moleculesde_b200 import graph as G
from moleculesde_b200.data import Batch, synth_molecules
from moleculesde_b200.pretrain import PretrainStep

dev = torch.device("cuda:0")
ps = PretrainStep(*bench.build_pretrain_models(), dev)
hb = Batch.from_data_list(synth_molecules(256, 7000, "pcqm"))
host = {k: getattr(hb, k).pin_memory() for k in ("x", "edge_index", "edge_attr", "positions", "batch")}

def stage():
    b = hb.__class__()
    for k, v in host.items():
        setattr(b, k, v.to(dev, non_blocking=True))
    b.num_graphs = hb.num_graphs
    csr = G.extend_graph(b.edge_index, b.batch, b.num_graphs)
    b.extended_edge_index = csr.edge_index
    return b

for _ in range(3):
    ps.step(stage())
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    b = stage(); torch.cuda.synchronize(); t1 = time.perf_counter()
    ps.step(b); torch.cuda.synchronize()
print("e2e ms/step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    ps.step(stage()); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
