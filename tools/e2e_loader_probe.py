"""A/B of the pretraining end-to-end step (loss of every step read on the host one step late, as bench.py does):
DeviceLoader thread (GIL switch interval default / 0.2 ms / 1 ms) vs the SAME preparation issued by the consumer thread itself on a
copy stream right after it queued the step (software pipelining without a second Python thread) vs a static prepared batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from moleculesde_b200.data import Batch, synth_molecules
from moleculesde_b200.loader import DeviceLoader, InlineLoader, pin_batch
from moleculesde_b200.pretrain import PretrainStep

dev = torch.device("cuda:0")
ps = PretrainStep(*bench.build_pretrain_models(), dev)
hb = pin_batch(Batch.from_data_list(synth_molecules(256, 7000, "pcqm")))
STEPS = 30
loss_h = [torch.empty(1).pin_memory() for _ in range(2)]
loss_ev = [torch.cuda.Event() for _ in range(2)]


def run(name, it):
    torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
    for k, b in enumerate(it):
        o = ps.step(b)
        loss_h[k & 1].copy_(o["loss_2d3d"].reshape(1), non_blocking=True)
        loss_ev[k & 1].record()
        if k > 0:
            loss_ev[(k - 1) & 1].synchronize()
            float(loss_h[(k - 1) & 1])
        n += 1
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"{name:52s} {dt * 1e3:7.2f} ms/step  {256 / dt:8.0f} molecules/s", flush=True)


for b in DeviceLoader([hb] * 3, dev, prepare=ps.prepare):
    ps.step(b)
for rep in range(2):
    run("DeviceLoader thread, default switch interval", DeviceLoader([hb] * STEPS, dev, prepare=ps.prepare))
    run("DeviceLoader thread, switch interval 1 ms", DeviceLoader([hb] * STEPS, dev, prepare=ps.prepare, switch_interval=1e-3))
    run("DeviceLoader thread, switch interval 0.2 ms", DeviceLoader([hb] * STEPS, dev, prepare=ps.prepare, switch_interval=2e-4))
    run("InlineLoader (consumer thread, copy stream)", InlineLoader([hb] * STEPS, dev, prepare=ps.prepare))
    sb = next(iter(InlineLoader([hb], dev, prepare=ps.prepare)))
    run("static prepared batch", (sb for _ in range(STEPS)))
