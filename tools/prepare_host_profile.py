"""Host cost of PretrainStep.prepare (per-batch index structures) and of the e2e loop pieces.  python tools/prepare_host_profile.py [B]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moleculesde_b200.data import Batch, synth_molecules
from moleculesde_b200.loader import DeviceLoader, pin_batch
from moleculesde_b200.pretrain import PretrainStep
from tools.pretrain_probe import build_models
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
ps = PretrainStep(*build_models(), dev)
hb = pin_batch(Batch.from_data_list(synth_molecules(B, 7000, "pcqm")))
for _ in range(3):
    b = hb.to(dev); ps.prepare(b); ps.step(b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    b = hb.to(dev)
    ps.prepare(b)
torch.cuda.synchronize()
print(f"H2D + prepare alone: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms per batch")
t0 = time.perf_counter()
for _ in range(10):
    b = hb.to(dev); ps.prepare(b); o = ps.step(b); float(o["loss_2d3d"])
print(f"serial prepare + step + loss read: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms per batch")
mx = int((hb.ptr[1:] - hb.ptr[:-1]).max())
t0 = time.perf_counter()
for _ in range(10):
    b = hb.to(dev); ps.prepare(b, mx); o = ps.step(b); float(o["loss_2d3d"])
print(f"serial prepare(max_nodes given: no mid-step sync) + step + loss read: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms per batch")
b = hb.to(dev); ps.prepare(b, mx)
t0 = time.perf_counter()
for _ in range(10):
    o = ps.step(b); float(o["loss_2d3d"])
print(f"static prepared batch: step + loss read: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms per batch")
t0 = time.perf_counter()
for _ in range(10):
    b.__dict__.pop("_molsde_train_cache", None)
    ps.prepare(b, mx); o = ps.step(b); float(o["loss_2d3d"])
print(f"same device batch, train cache dropped + prepare + step + loss read: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms per batch")
t0 = time.perf_counter()
for b2 in DeviceLoader([hb] * 20, dev, prepare=ps.prepare):
    o = ps.step(b2); float(o["loss_2d3d"])
print(f"DeviceLoader (prepare one batch ahead in a thread) + step + loss read: {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms per batch")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    b = hb.to(dev)
    ps.prepare(b)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
