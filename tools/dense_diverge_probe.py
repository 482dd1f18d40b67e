"""Debug: does the 3D->2D sampler (random-init weights, VP) stay finite, eager vs graph, fused vs unfused?  python tools/dense_diverge_probe.py [graphs] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moleculesde_b200.data import synth_batch
from moleculesde_b200.sampler_dense import node_adj_PC_generation
from moleculesde_b200.sde_3d_to_2d import SDEModel3Dto2D_node_adj_dense
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 150
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = SDEModel3Dto2D_node_adj_dense(dim3D=300, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4, emb_dim=300,
                                  num_linears=3, beta_min=0.2, beta_max=1.0, num_diffusion_timesteps=1000, SDE_type="VP", num_class_X=119,
                                  noise_on_one_hot=True).to(dev).eval()
b = synth_batch(B, 3, "padded64").to(dev)
h3d = torch.randn(b.positions.size(0), 300, device=dev)
_, rep, _, _, Nm = m.dense_inputs(h3d, b)
g = torch.Generator().manual_seed(1)
x0 = torch.randn(B, Nm, 119, generator=g); a0 = torch.randn(B, Nm, Nm, generator=g)
for ug in (False, True):
    for n in (10, 40, 80, 120, steps):
        torch.manual_seed(5)
        x, adj, xm, am = node_adj_PC_generation(rep, b, m, B=B, max_num_nodes=Nm, num_class_X=119, n_steps=1, x_init=x0, adj_init=a0,
                                                diffusion_steps=n, use_graph=ug)
        print(f"graph={ug} steps={n}: finite x {bool(torch.isfinite(xm).all())} adj {bool(torch.isfinite(am).all())} "
              f"|x|max {float(xm.abs().max()):.3e} |adj|max {float(am.abs().max()):.3e}", flush=True)
