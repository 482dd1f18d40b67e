"""Does the eager end-to-end step slow down after a CUDA-graph capture in the same process?  (bench.py measures e2e after it.)"""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from moleculesde_b200.data import Batch, synth_molecules
from moleculesde_b200.loader import DeviceLoader, pin_batch
from moleculesde_b200.pretrain import PretrainStep

dev = torch.device("cuda:0")
ps = PretrainStep(*bench.build_pretrain_models(), dev)
hb = pin_batch(Batch.from_data_list(synth_molecules(256, 7000, "pcqm")))
loss_h = torch.empty(1).pin_memory()

def e2e(tag, steps=20):
    for bw in DeviceLoader([hb] * 3, dev, prepare=ps.prepare):
        ps.step(bw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for b in DeviceLoader([hb] * steps, dev, prepare=ps.prepare):
        o = ps.step(b); loss_h.copy_(o["loss_2d3d"].reshape(1), non_blocking=True); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f"{tag:40s} {dt * 1e3:7.2f} ms/step  gc counts {gc.get_count()} objects {len(gc.get_objects())}")

e2e("before capture")
b = ps.prepare(hb.to(dev), None)
g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    ps.forward_backward(b); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=side):
        ps.forward_backward(b)
torch.cuda.synchronize()
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
e2e("after capture")
gc.collect(); gc.freeze()
e2e("after capture, gc.freeze()")
gc.disable()
e2e("gc disabled")
gc.enable()
del g
torch.cuda.empty_cache()
e2e("graph deleted")
