#!/usr/bin/env python
"""Device time of the pretraining iteration per loss branch (CUDA-graph replay, one stream): which branch is the critical
path of the fork/join schedule in PretrainStep.forward_backward.  `python tools/pretrain_branch_probe.py [B]`"""
import os
import sys

os.environ["MOLSDE_SINGLE_STREAM"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from moleculesde_b200.pretrain import PretrainStep  # noqa: E402
from tools.pretrain_probe import build_batch, build_models  # noqa: E402


def replay_ms(ps, b, steps=20):
    for _ in range(2):
        ps.forward_backward(b)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            ps.forward_backward(b)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, ps.launches


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda:0")
    b = build_batch(B, 0, dev)
    for name, c in (("all", (1, 1, 1)), ("encoders + CL", (1, 0, 0)), ("encoders + CL + 2D->3D", (1, 1, 0)),
                    ("encoders + CL + 3D->2D", (1, 0, 1))):
        ps = PretrainStep(*build_models(), dev, coeff_contrastive=c[0], coeff_2Dto3D=c[1], coeff_3Dto2D=c[2])
        ms, n = replay_ms(ps, b)
        print(f"{name:28s} {ms:7.2f} ms  {n} launches")


if __name__ == "__main__":
    main()
