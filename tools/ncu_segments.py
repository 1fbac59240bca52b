"""Split an ncu SASS source-page CSV at BAR.SYNC instructions and report samples / stall mix / instruction mix per segment.
usage: python tools/ncu_segments.py <csv>"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
segs = []; cur = None
def new():
    return dict(samples=0, inst=0, st=collections.Counter(), ops=collections.Counter(), opsamp=collections.Counter(), first=None)
cur = new()
tot = 0
for r in rows[2:]:
    sass = r[col["Source"]].strip()
    s = int(r[col["# Samples"]] or 0); tot += s
    op = re.sub(r"^@!?U?P\d+\s+", "", sass).split()[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "RED", "ATOM")) else op.split(".")[0]
    if cur["first"] is None: cur["first"] = r[col["Address"]][-5:]
    cur["samples"] += s; cur["inst"] += int(r[col["Instructions Executed"]] or 0)
    cur["ops"][op] += int(r[col["Instructions Executed"]] or 0); cur["opsamp"][op] += s
    for h in stall_cols: cur["st"][h[6:]] += int(r[col[h]] or 0)
    if sass.startswith("BAR") or " BAR." in sass:
        segs.append(cur); cur = new()
segs.append(cur)
for i, sg in enumerate(segs):
    if sg["samples"] < tot * 0.002: continue
    print("seg %2d @%s  %5.2f%% samples  inst %.3g | stalls: %s" % (i, sg["first"], 100.0 * sg["samples"] / tot, sg["inst"],
          " ".join("%s=%.1f%%" % (k, 100.0 * v / max(1, sg["samples"])) for k, v in sg["st"].most_common(5))))
    print("        ops(exec M): %s" % " ".join("%s=%.1f" % (k, v / 1e6) for k, v in sg["ops"].most_common(9)))
    print("        samples by op: %s" % " ".join("%s=%.1f%%" % (k, 100.0 * v / max(1, sg["samples"])) for k, v in sg["opsamp"].most_common(6)))
