"""Where does the HOST time of one eager pretraining iteration go?  cProfile over 10 iterations (static batch, no loader).
python tools/pretrain_host_profile.py [B]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moleculesde_b200.pretrain import PretrainStep
from tools.pretrain_probe import build_batch, build_models
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
ps = PretrainStep(*build_models(), dev)
b = build_batch(B, 0, dev)
for _ in range(3):
    ps.step(b)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(10):
        ps.step(b)
    t_issue = (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    t_total = (time.perf_counter() - t0) / 10
    print(f"eager step: host issue {t_issue * 1e3:.2f} ms, incl. device drain {t_total * 1e3:.2f} ms, {ps.launches} launches")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    ps.step(b)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumtime").print_stats(40)
