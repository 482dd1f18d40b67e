"""Per-phase cycle breakdown of the fused PC kernel (library built with MOLSDE_PROF=1)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from moleculesde_b200 import graph as G, _abi
from moleculesde_b200.sampler import position_PC_generation
dev = torch.device("cuda:0")
nm = int(sys.argv[1]) if len(sys.argv) > 1 else 296
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mols, hb, group_ptr = bench.build_workload(nm, 10, 0)
model = bench.make_model(dev)
d = hb.to(dev)
csr = G.extend_graph(d.edge_index, d.batch, d.num_graphs)
d.extended_edge_index, d._molsde_ext_csr = csr.edge_index, csr
g = torch.Generator().manual_seed(100)
n = hb.positions.size(0)
rep = torch.randn(n, 300, generator=g).to(dev); pos0 = torch.randn(n, 3, generator=g).to(dev)
L = _abi.lib()
buf = (ctypes.c_ulonglong * (148 * 8))()
position_PC_generation(rep, d, pos0, model, model.sde_pos, group_ptr=group_ptr, seed=1, diffusion_steps=2)
L.molsde_debug_read_prof(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
position_PC_generation(rep, d, pos0, model, model.sde_pos, group_ptr=group_ptr, seed=1, diffusion_steps=steps)
e1.record(); torch.cuda.synchronize()
L.molsde_debug_read_prof(buf)
a = np.array(list(buf), dtype=np.float64).reshape(148, 8)
tot = a.sum(0)
names = ["E0 edge features", "stage GAT weights", "node q|k|v", "GAT edge phase", "node update", "basis (tcgen05)", "-", "-"]
prep = model.prepared(d, group_ptr)
ntile_evals = prep.plan.num_tiles * steps * 2
print(f"molecules={nm} steps={steps} kernel_ms={e0.elapsed_time(e1):.1f} tiles={prep.plan.num_tiles} tile-evals={ntile_evals}")
for i in range(6):
    print(f"{names[i]:22s} {100*tot[i]/tot.sum():5.1f}%   {tot[i]/ntile_evals:9.0f} cycles per tile-eval")
print(f"sum {tot.sum()/ntile_evals:.0f} cycles per tile-eval; busy cycles per CTA mean {a.sum(1).mean():.3e} max {a.sum(1).max():.3e}")
