"""Summary of a tools/trace_pair.py timeline: tile period, issue time per job (first operand slab seen -> job issued)."""
import sys, statistics
ev = []
for l in open(sys.argv[1]):
    p = l.split(None, 1)
    if len(p) == 2 and p[0].isdigit():
        ev.append((int(p[0]), p[1].strip()))
tiles = [t for t, n in ev if n == "rows  TILE START"]
per = [b - a for a, b in zip(tiles, tiles[1:])]
jobs = []
t_s0 = None
for t, n in ev:
    if n == "ISSUER slab 0 ready":
        t_s0 = t
    elif n == "ISSUER job issued" and t_s0 is not None:
        jobs.append(t - t_s0); t_s0 = None
print(f"{sys.argv[1]}: tiles {len(tiles)}  period median {statistics.median(per):.0f} (min {min(per)}, max {max(per)})  "
      f"issue time per a_new job median {statistics.median(jobs):.0f} cycles ({statistics.median(jobs) / 48:.0f} per N=256-equivalent UMMA)")
