"""Condenses a `compute-sanitizer --tool racecheck` log: one line per (hazard type, kernel, access pair) with a count."""
import collections
import re
import sys

c = collections.Counter()
cur = None
kernel = "?"
for line in open(sys.argv[1], errors="replace"):
    m = re.search(r"(Potential )?(WAR|RAW|WAW) hazard detected at (__shared__|__global__)", line)
    if m:
        cur = [m.group(2), m.group(3), None, None]
        continue
    if cur is not None:
        m = re.search(r"(Read|Write) Thread .* at (.*?)\+0x[0-9a-f]+ in (\S+)", line)
        if m:
            who = "%s %s @ %s" % (m.group(1), m.group(2).split("(")[0], m.group(3))
            if cur[2] is None:
                cur[2] = who
            else:
                cur[3] = who
        m = re.search(r"Host Frame: (tgr::launch_\w+)", line)
        if m and cur[2] is not None:
            c[(cur[0], cur[1], cur[2], cur[3], m.group(1))] += 1
            cur = None
    m = re.search(r"RACECHECK SUMMARY: (.*)", line)
    if m:
        print("RACECHECK SUMMARY:", m.group(1))
for (typ, space, a, b, launch), n in c.most_common():
    print("%7d  %s %s  %s  |  %s  |  %s" % (n, typ, space, launch, a, b))
