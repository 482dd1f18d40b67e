"""Correlate an ncu report's per-SASS-instruction samples with CUDA source lines (nvdisasm line info).
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [top_n]
Needs the same libmolsde_b200.so that was profiled (in-tree build)."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kname = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(REPO, "moleculesde_b200", "libmolsde_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = os.path.join(tmp, os.environ.get("NCU_CUBIN", "sde2d3d") + ".sm_100a.cubin")
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], stdout=subprocess.PIPE, text=True).stdout
cur, inside, off2line = None, False, {}
for ln in dis.split("\n"):
    if ln.startswith(".text."):
        inside = kname in ln
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
    if m and inside:
        off2line[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:   # a report with several launches repeats the two header rows: keep the first launch only
    if not r or not re.match(r"^(0x)?[0-9a-fA-F]+$", r[0]):
        break
    data.append(r)
base = int(data[0][0], 16)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg, inst = collections.Counter(), collections.Counter()
stall = collections.defaultdict(collections.Counter)
tot = 0
tot_stall = collections.Counter()
for r in data:
    off = int(r[0], 16) - base
    s = int(r[idx["# Samples"]] or 0); tot += s
    line = off2line.get(off)
    agg[line] += s; inst[line] += int(r[idx["Instructions Executed"]] or 0)
    for c in stall_cols:
        v = int(r[idx[c]] or 0); stall[line][c] += v; tot_stall[c] += v
src = open(os.path.join(REPO, "moleculesde_b200", "csrc", "sde2d3d.cu")).read().split("\n")
print("total samples", tot, " instructions", sum(inst.values()))
print("stalls:", ", ".join(f"{k[6:]}={100*v/tot:.1f}%" for k, v in tot_stall.most_common(10)))
for line, s in agg.most_common(topn):
    txt = src[line[1] - 1].strip()[:95] if line and line[0] == "sde2d3d.cu" else str(line)
    top = ", ".join(f"{k[6:]}={v}" for k, v in stall[line].most_common(3))
    print(f"{str(line):26s} {100*s/tot:5.1f}% inst={inst[line]:>10} [{top}] | {txt}")
