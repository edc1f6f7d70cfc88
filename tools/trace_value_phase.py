"""Summary of a tools/trace_pair.py timeline: UMMA issue time per K-slab in the hidden-layer jobs vs the last Linear's jobs."""
import sys, statistics
ev = []
for l in open(sys.argv[1]):
    p = l.split(None, 1)
    if len(p) == 2 and p[0].isdigit():
        ev.append((int(p[0]), p[1].strip()))
# issuer events only, in order; a job = [slab 0..3 ready] + issued
iss = [(t, n) for t, n in ev if n.startswith("ISSUER")]
jobs, cur = [], []
for t, n in iss:
    cur.append((t, n))
    if n == "ISSUER job issued":
        if len(cur) == 5: jobs.append([c[0] for c in cur])
        cur = []
# tile = 9 jobs: k.L2 k.L3 k.L4 v.L2 v.L3 v.L4 c0 c1 c2 ; align on the shortest job pattern: use modulo from the first TILE START
per = {}
tiles = [t for t, n in ev if n == "rows  TILE START"]
import bisect
for j in jobs:
    i = bisect.bisect_right(tiles, j[4]) - 1
    if i < 1: continue
    per.setdefault(i, []).append(j)
slab = {k: [] for k in range(9)}
for i, js in per.items():
    if len(js) != 9: continue
    for k, j in enumerate(js):
        slab[k] += [j[1] - j[0], j[2] - j[1], j[3] - j[2], j[4] - j[3]]
names = ["k.L3", "k.L4", "v.L2", "v.L3", "v.L4", "L5 c0", "L5 c1", "L5 c2 (N=128)", "next k.L2"]
for k in range(9):
    if slab[k]:
        print(f"{names[k]:16s} cycles per K-slab (12 UMMAs): median {statistics.median(slab[k]):6.0f}  min {min(slab[k]):6d}  max {max(slab[k]):6d}")
