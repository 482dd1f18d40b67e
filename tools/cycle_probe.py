"""Which reference cycles does one pretraining step leave behind?  (Cycles delay the release of device tensors until the cyclic
GC runs: the caching allocator then grows and the eager step slows down.)"""
import collections, gc, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from moleculesde_b200.data import Batch, synth_molecules
from moleculesde_b200.pretrain import PretrainStep

dev = torch.device("cuda:0")
ps = PretrainStep(*bench.build_pretrain_models(), dev)
hb = Batch.from_data_list(synth_molecules(64, 7000, "pcqm"))
for _ in range(2):
    ps.step(ps.prepare(hb.to(dev), None))
torch.cuda.synchronize()
gc.collect()
gc.disable()
gc.set_debug(gc.DEBUG_SAVEALL)
b = ps.prepare(hb.to(dev), None)
out = ps.step(b)
del out, b
torch.cuda.synchronize()
n = gc.collect()
print("unreachable objects after one step:", n)
hist = collections.Counter(type(o).__name__ for o in gc.garbage)
print(hist.most_common(12))
fn = collections.Counter(getattr(o, "__qualname__", "?") for o in gc.garbage if type(o).__name__ == "function")
print(fn.most_common(25))
tens = [o for o in gc.garbage if torch.is_tensor(o)]
print("tensors in cycles:", len(tens), "bytes", sum(t.numel() * t.element_size() for t in tens))
