#!/usr/bin/env python
"""Kernel timeline of ONE CUDA-graph replay of the pretraining iteration (torch.profiler / CUPTI, not a bench number):
per-stream busy time, how much of the span has 1 / 2 / 3 streams active, and the kernels that run while nothing overlaps them
(the critical path of the fork/join schedule).  Writes gpurun_out/pretrain_timeline.json.  `python tools/pretrain_timeline.py [B]`"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from moleculesde_b200.pretrain import PretrainStep  # noqa: E402
from tools.pretrain_probe import build_batch, build_models  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda:0")
    ps = PretrainStep(*build_models(), dev)
    b = build_batch(B, 0, dev)
    for _ in range(3):
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
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        g.replay()
        torch.cuda.synchronize()
    os.makedirs("gpurun_out", exist_ok=True)
    path = "gpurun_out/pretrain_trace.json"
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    os.remove(path)
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    rows = [{"name": e["name"][:60], "t": e["ts"] - t0, "dur": e["dur"], "stream": e["args"].get("stream")} for e in ev]
    json.dump(rows, open("gpurun_out/pretrain_timeline.json", "w"))
    span = max(r["t"] + r["dur"] for r in rows)
    print(f"{len(rows)} kernels, span {span / 1e3:.2f} ms")
    streams = sorted({r["stream"] for r in rows})
    for s in streams:
        rs = [r for r in rows if r["stream"] == s]
        print(f"  stream {s}: {len(rs)} kernels, busy {sum(r['dur'] for r in rs) / 1e3:.2f} ms, first {rs[0]['t'] / 1e3:.2f} last end "
              f"{max(r['t'] + r['dur'] for r in rs) / 1e3:.2f} ms")
    # sweep: number of active kernels over time
    pts = []
    for i, r in enumerate(rows):
        pts.append((r["t"], 1, i))
        pts.append((r["t"] + r["dur"], -1, i))
    pts.sort()
    active, last, hist, solo = set(), 0.0, {}, {}
    for t, d, i in pts:
        if t > last:
            hist[len(active)] = hist.get(len(active), 0.0) + (t - last)
            if len(active) == 1:
                n = rows[next(iter(active))]["name"]
                solo[n] = solo.get(n, 0.0) + (t - last)
        last = t
        (active.add if d > 0 else active.discard)(i)
    print("  active kernels -> ms:", {k: round(v / 1e3, 2) for k, v in sorted(hist.items())})
    print("  kernels running alone (ms):")
    for n, v in sorted(solo.items(), key=lambda kv: -kv[1])[:14]:
        print(f"    {v / 1e3:6.2f}  {n}")


if __name__ == "__main__":
    main()
