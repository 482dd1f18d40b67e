"""Time the fused PC kernel on a reduced workload: python tools/pc_time_probe.py [molecules=296] [steps=20] [repeat=10]
Prints ms per launch and microseconds per score evaluation per CTA slot (molecules / 148 groups per CTA, 2 evaluations per step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from moleculesde_b200 import graph as G
from moleculesde_b200.sampler import position_PC_generation
dev = torch.device("cuda:0")
nm = int(sys.argv[1]) if len(sys.argv) > 1 else 296
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rep_n = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mols, hb, group_ptr = bench.build_workload(nm, rep_n, 0)
model = bench.make_model(dev)
d = hb.to(dev)
csr = G.extend_graph(d.edge_index, d.batch, d.num_graphs)
d.extended_edge_index, d._molsde_ext_csr = csr.edge_index, csr
g = torch.Generator().manual_seed(100)
n = hb.positions.size(0)
rep = torch.randn(n, 300, generator=g).to(dev); pos0 = torch.randn(n, 3, generator=g).to(dev)
_, pm = position_PC_generation(rep, d, pos0, model, model.sde_pos, group_ptr=group_ptr, seed=1, diffusion_steps=2)
torch.cuda.synchronize()
prep = model.prepared(d, group_ptr)
print("status", int(prep.status.item()), "finite", bool(torch.isfinite(pm).all()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
times = []
for _ in range(5):   # short launches are noisy (clock ramp): best of 5
    e0.record()
    _, pm = position_PC_generation(rep, d, pos0, model, model.sde_pos, group_ptr=group_ptr, seed=1, diffusion_steps=steps)
    e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = min(times)
print("launch times (ms):", " ".join(f"{t:.2f}" for t in times))
rounds = -(-nm // 148)
print(f"molecules={nm} x{rep_n} steps={steps} tiles={prep.plan.num_tiles} atoms={n} edges={prep.csr.num_edges} ms={ms:.2f} "
      f"us/eval/CTA={1e3 * ms / (rounds * steps * 2):.1f} status={int(prep.status.item())} finite={bool(torch.isfinite(pm).all())}")
print(f"extrapolated conformers/s at 1024 molecules x 1000 steps: {nm * rep_n / (ms / 1e3) * steps / 1000:.0f}")
