"""Top CUDA source lines by warp-stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
data = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") or h.lower().startswith("warp stall")]
        continue
    if hdr is None or len(r) <= si:
        continue
    if r[2] != "-":      # SASS row; the source-line rows carry "-" as address
        continue
    try:
        n = int(r[si])
    except ValueError:
        continue
    data.append((n, r[0], r[1], r))
tot = sum(d[0] for d in data) or 1
print("total samples", tot)
names = hdr
for n, line, src, r in sorted(data, key=lambda x: -x[0])[:top]:
    extra = []
    for i, h in enumerate(names):
        if h.startswith("stall_"):
            try:
                v = int(r[i])
            except ValueError:
                continue
            if v > 0.15 * n:
                extra.append("%s=%d" % (h[6:], v))
    print("%7d %5.1f%%  L%-4s %-100s %s" % (n, 100.0 * n / tot, line, src.strip()[:100], " ".join(extra)))
