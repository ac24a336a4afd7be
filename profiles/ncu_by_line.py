"""Joins an ncu report's SASS page (instructions executed, stall samples per address) with nvdisasm's inline line
info of the shipped cubin, and aggregates by innermost source line and by the enclosing call chain's outermost line.
    python profiles/ncu_by_line.py <report.ncu-rep> <kernel mangled-name substring> [top N]
(needs the library built from the same sources: spark-sched-sim_b200/_lib/libssb.so)"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(REPO, "spark-sched-sim_b200", "_lib", "libssb.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
seg = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-gi", cub], capture_output=True, text=True).stdout.split("\n")
    start = None
    for i, l in enumerate(dis):
        if l.strip().startswith(".section") and ".text." in l:
            if start is not None:
                seg = dis[start:i]
                break
            if kname in l:
                start = i
    if seg:
        break
assert seg, "kernel not found"
addr2 = {}
first = last = None
for l in seg:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        k = (m.group(1).split("/")[-1], int(m.group(2)))
        if first is None:
            first = k
        last = k
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        if first:
            addr2[int(m.group(1), 16)] = (first, last)
            keep = (first, last)
        elif addr2:
            addr2[int(m.group(1), 16)] = keep
        first = last = None
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.split("\n")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) >= 6 and r[0].startswith("0x")]
base = min(int(r[0], 16) for r in data)
ia, sa = hdr.index("Instructions Executed"), hdr.index("# Samples")
inner_i, inner_s, outer_i, outer_s = (collections.Counter() for _ in range(4))
seen = set()
for r in data:
    a = int(r[0], 16) - base
    if a in seen:
        continue
    seen.add(a)
    f, l = addr2.get(a, (("?", 0), ("?", 0)))
    n, s = float(r[ia] or 0), float(r[sa] or 0)
    inner_i[f] += n; inner_s[f] += s; outer_i[l] += n; outer_s[l] += s
ti, ts = sum(inner_i.values()), sum(inner_s.values())
src = {}


def line(f, ln):
    if f not in src:
        p = glob.glob(os.path.join(REPO, "spark-sched-sim_b200", "csrc", f))
        src[f] = open(p[0]).read().split("\n") if p else []
    return src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""


print(f"total warp instructions {ti:.4g}, stall samples {int(ts)}")
print("--- by outermost line of the inline chain")
for k, v in outer_i.most_common(14):
    print(f"{v / ti * 100:5.1f}% inst {outer_s[k] / ts * 100:5.1f}% smp  {k[0]}:{k[1]}  {line(*k)}")
print("--- by innermost line")
for k, v in inner_s.most_common(top):
    print(f"{inner_i[k] / ti * 100:5.1f}% inst {v / ts * 100:5.1f}% smp  {k[0]}:{k[1]}  {line(*k)}")
# --- by line ranges of ssb_decima_fused.cuh (when profiling the fused policy kernel)
if "--ranges" in sys.argv:
    ranges = [("split3/pack2", 60, 73), ("Blob/spec", 74, 165), ("store_a_row", 172, 194), ("issue_layer", 195, 216),
              ("run_layer", 217, 231), ("act", 232, 238), ("mlp_tile", 239, 292), ("k_mlp_rows", 293, 340),
              ("gather", 384, 478), ("scatter", 479, 488), ("run_tiles", 489, 505), ("run_phase", 506, 524),
              ("kernel prologue", 525, 562), ("group prologue/adapter glue", 563, 616), ("node_of/job_item", 617, 629),
              ("phase calls+preds", 630, 670), ("sampling", 671, 745)]
    agg_i, agg_s = collections.Counter(), collections.Counter()
    for (f, ln), v in inner_i.items():
        name = f
        if f == "ssb_decima_fused.cuh":
            name = next((n for n, a, b in ranges if a <= ln <= b), "fused:other")
        agg_i[name] += v; agg_s[name] += inner_s[(f, ln)]
    print("--- by function group")
    for k, v in agg_i.most_common(30):
        print(f"{v / ti * 100:5.1f}% inst {agg_s[k] / ts * 100:5.1f}% smp  {k}")
