"""Diagnostic: run the fused PC kernel for a growing number of reverse steps and report non-finite atoms."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from moleculesde_b200 import graph as G
from moleculesde_b200.sampler import position_PC_generation

dev = torch.device("cuda:0")
mols, hb, group_ptr = bench.build_workload(int(sys.argv[1]) if len(sys.argv) > 1 else 256, 10, 0)
model = bench.make_model(dev)
d = hb.to(dev)
csr = G.extend_graph(d.edge_index, d.batch, d.num_graphs)
d.extended_edge_index = csr.edge_index
d._molsde_ext_csr = csr
g = torch.Generator().manual_seed(100)
n = hb.positions.size(0)
rep = torch.randn(n, 300, generator=g).to(dev)
pos0 = torch.randn(n, 3, generator=g).to(dev)
for steps in (10, 50, 100, 200, 400, 1000):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, pm = position_PC_generation(rep, d, pos0, model, model.sde_pos, group_ptr=group_ptr, seed=1, diffusion_steps=steps)
    e1.record()
    torch.cuda.synchronize()
    bad = (~torch.isfinite(pm)).any(dim=1)
    print(f"steps={steps} ms={e0.elapsed_time(e1):.1f} nonfinite_atoms={int(bad.sum())}/{n} "
          f"groups_hit={int(torch.unique(d.batch[bad] // 10).numel())} max|pos|={pm[~bad].abs().max().item():.3f}", flush=True)
