"""Per-source-line share of warp-stall samples and of executed instructions from one ncu capture taken with
`--set full --import-source on` (the library is compiled with -lineinfo):

    python tools/ncu_source_lines.py capture.ncu-rep [top N]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur, hdr, out = None, None, []
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif r and hdr and len(r) > 8 and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))

        def f(k):
            try:
                return float(d[k])
            except (KeyError, ValueError):
                return 0.0
        out.append((cur, int(r[0]), r[1][:100], f("# Samples"), f("Instructions Executed"), f("Thread Instructions Executed")))
ts, ti = sum(o[3] for o in out), sum(o[4] for o in out)
print("samples %d   warp instructions %d" % (ts, ti))
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print("%-14s %4d  smp %5.1f%%  inst %5.1f%%  lanes %4.1f  %s" % (o[0], o[1], 100 * o[3] / ts, 100 * o[4] / ti, o[5] / max(o[4], 1), o[2]))
