"""Where does `inference.generate_samples_from_testset` spend its time?  (one launch of 592 molecules x 10 conformers)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moleculesde_b200 import checkpoint, graph as G
from moleculesde_b200.data import Batch, Data, synth_molecules
from moleculesde_b200.sampler import position_PC_generation

dev = torch.device("cuda:0")
torch.manual_seed(42)
models = checkpoint.build_models()
gnn, m23 = models["model_2D"].to(dev).eval(), models["SDE_2Dto3D_model"].to(dev).eval()
mols = synth_molecules(592, 42, "pcqm")
R = 10

def tick(name, t0):
    torch.cuda.synchronize(); t1 = time.time(); print(f"{name:28s} {t1 - t0:8.3f} s"); return t1

for rnd in range(2):
    print("round", rnd)
    t = time.time()
    copies = []
    for m in mols:
        for _ in range(R):
            d = Data.from_dict({k: m[k] for k in m.keys}); d.pos_gen = torch.ones_like(m.positions); copies.append(d)
    batch = Batch.from_data_list(copies).to(dev)
    t = tick("collate + H2D", t)
    csr = G.extend_graph(batch.edge_index, batch.batch, batch.num_graphs)
    batch.extended_edge_index, batch._molsde_ext_csr = csr.edge_index, csr
    t = tick("extend_graph", t)
    with torch.no_grad():
        rep = gnn(batch.x, batch.edge_index, batch.edge_attr)
    t = tick("GIN forward", t)
    group_ptr = torch.arange(0, len(mols) * R + 1, R, dtype=torch.long)
    prep = m23.prepared(batch, group_ptr)
    t = tick("prepared (CSR, tile plan)", t)
    m23.invariants(rep, prep)
    t = tick("invariants", t)
    pos_init = m23.sde_pos.prior_sampling(batch.positions.shape).to(dev)
    for steps in (10, 100):
        _, pos = position_PC_generation(rep, batch, pos_init, m23, m23.sde_pos, group_ptr=group_ptr, seed=1, diffusion_steps=steps)
        t = tick(f"PC kernel, {steps} steps", t)
    print("finite:", bool(torch.isfinite(pos).all()), "max |pos|", float(pos.abs().max()))
    rep2 = torch.randn_like(rep)
    _, pos = position_PC_generation(rep2, batch, pos_init, m23, m23.sde_pos, group_ptr=group_ptr, seed=1, diffusion_steps=100)
    t = tick("PC kernel, 100 steps, randn rep", t)
