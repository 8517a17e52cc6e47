"""Top SASS instructions by stall samples from an `ncu --page source --csv` dump."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
iS, iA, iI = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for n, r in enumerate(rows[2:]):
    try:
        s = int(r[iA])
    except Exception:
        continue
    tot += s
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    data.append((s, n, r[iS].strip(), r[iI], st))
print("total samples", tot)
for s, n, src, ie, st in sorted(data, reverse=True)[:top]:
    print("%6d %5.1f%%  #%4d  %-58s exec=%-10s %s" % (s, 100.0 * s / tot, n, src[:58], ie, " ".join("%s:%d" % (h[6:], v) for v, h in st)))
