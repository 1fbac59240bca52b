"""Aggregate an ncu source-page CSV (SASS level) by CUDA source line using nvdisasm -g line tables.
usage: python tools/ncu_lines.py <ncu --page source --csv file> <nvdisasm -g -c output> <kernel mangled-name substring> [topN]
Prints per-line: share of warp-stall samples, instructions executed, and the dominant stall reasons."""
import csv, re, sys, collections
src_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
# ---- line table: instruction index -> (file, line), with the inline chain collapsed to the innermost location
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
loc = None
table = []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", l):
        table.append(loc)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
assert abs(len(body) - len(table)) < 4, (len(body), len(table))
agg = collections.defaultdict(lambda: collections.Counter())
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
for r, lc in zip(body, table):
    s = int(r[col["# Samples"]] or 0)
    tot += s
    a = agg[lc]
    a["samples"] += s
    a["inst"] += int(r[col["Instructions Executed"]] or 0)
    for h in stall_cols:
        a[h] += int(r[col[h]] or 0)
print("total samples", tot)
for lc, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((a[h], h[6:]) for h in stall_cols), reverse=True)[:3]
    print("%-18s %5d  %5.2f%%  inst %9d  %s" % (lc[0][:18] if lc else "?", lc[1] if lc else 0, 100.0 * a["samples"] / tot, a["inst"],
                                             " ".join("%s=%d" % (n, v) for v, n in st if v)))
