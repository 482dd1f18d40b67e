"""A/B of the pretraining end-to-end step: batch prepared inline vs staged ahead by loader.DeviceLoader (thread + copy stream)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from moleculesde_b200.data import Batch, synth_molecules
from moleculesde_b200.loader import DeviceLoader, pin_batch
from moleculesde_b200.pretrain import PretrainStep

dev = torch.device("cuda:0")
ps = PretrainStep(*bench.build_pretrain_models(), dev)
hb = pin_batch(Batch.from_data_list(synth_molecules(256, 7000, "pcqm")))
mx = int((hb.ptr[1:] - hb.ptr[:-1]).max())
loss_h = torch.empty(1).pin_memory()
STEPS = 20

def inline():
    b = hb.to(dev, non_blocking=True)
    return ps.prepare(b, mx)

def run(name, it):
    torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
    for b in it:
        o = ps.step(b); loss_h.copy_(o["loss_2d3d"].reshape(1), non_blocking=True); torch.cuda.synchronize(); n += 1
    dt = (time.perf_counter() - t0) / n
    print(f"{name:42s} {dt * 1e3:7.2f} ms/step  {256 / dt:8.0f} molecules/s")

for _ in range(3):
    ps.step(inline())
run("inline prepare", (inline() for _ in range(STEPS)))
run("DeviceLoader depth 2", DeviceLoader([hb] * STEPS, dev, prepare=ps.prepare, depth=2))
sys.setswitchinterval(2e-4)
run("DeviceLoader depth 2, switchinterval 0.2ms", DeviceLoader([hb] * STEPS, dev, prepare=ps.prepare, depth=2))
run("DeviceLoader copies only + inline prepare", (ps.prepare(b, mx) for b in DeviceLoader([hb] * STEPS, dev, prepare=None, depth=2)))
sys.setswitchinterval(5e-3)
t0 = time.perf_counter()
for _ in range(STEPS):
    b = inline(); torch.cuda.synchronize()
print(f"prepare alone (with sync): {(time.perf_counter() - t0) / STEPS * 1e3:.2f} ms")
b = inline(); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(STEPS):
    ps.step(b); torch.cuda.synchronize()
print(f"step alone, static batch (with sync): {(time.perf_counter() - t0) / STEPS * 1e3:.2f} ms")
