#!/usr/bin/env python
"""Time the dense 3D->2D predictor-corrector sampler (BASELINE configs[3]: VP, noise_on_one_hot, graphs padded to 64 atoms).
`python tools/dense_sampler_probe.py [graphs] [steps]`"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from moleculesde_b200.data import synth_batch  # noqa: E402
from moleculesde_b200.sampler_dense import node_adj_PC_generation  # noqa: E402
from moleculesde_b200.sde_3d_to_2d import SDEModel3Dto2D_node_adj_dense  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = SDEModel3Dto2D_node_adj_dense(dim3D=300, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4,
                                      emb_dim=300, num_linears=3, beta_min=0.2, beta_max=1.0, num_diffusion_timesteps=1000,
                                      SDE_type="VP", num_class_X=119, noise_on_one_hot=True).to(dev).eval()
    with torch.no_grad():   # an untrained score network drives the reverse SDE to inf within ~100 steps: damp the output heads
        for net in (m.edge_score_network, m.node_score_network):
            net.final.layers[-1].weight.mul_(0.02)
            net.final.layers[-1].bias.mul_(0.02)
    b = synth_batch(B, 3, "padded64").to(dev)
    h3d = torch.randn(b.positions.size(0), 300, device=dev)
    _, rep, _, _, Nm = m.dense_inputs(h3d, b)
    print(f"graphs={B} Nm={Nm} atoms={b.positions.size(0)}")
    # one captured step, replayed `steps` times; the state is re-drawn from the prior every 50 steps (an untrained network diverges,
    # and non-finite states were measured to run slower), the time index keeps running -- the same protocol as bench.py
    pc, x0, adj0 = node_adj_PC_generation(rep, b, m, B=rep.size(0), max_num_nodes=Nm, num_class_X=119, n_steps=1, diffusion_steps=steps,
                                          use_graph=True, return_graph=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep_i in range(2):
        torch.cuda.synchronize()
        e0.record()
        done = 0
        while done < steps:
            n = min(50, steps - done)
            pc.reset(x0, adj0, done)
            pc.run(n)
            done += n
        e1.record()
        torch.cuda.synchronize()
    assert torch.isfinite(pc.x_mean).all() and torch.isfinite(pc.adj_mean).all()
    per = e0.elapsed_time(e1) / steps
    print(f"{steps} PC steps (graph replay, device-timed): {per:.3f} ms/step -> {B / per:.1f} graphs/s for a 1000-step trajectory")


if __name__ == "__main__":
    main()
