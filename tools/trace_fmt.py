"""Merge a CM_PLANE_TRACE dump (stderr of tools/plane_trace.py) into one time line per section."""
import re, sys
names = {1: 'P wait-empty done', 2: 'P issued', 3: 'M tmem_empty ok', 4: 'M full ok', 6: 'M committed', 10: 'E unit start',
         11: 'E prefetched', 12: 'E tmem_full ok', 13: 'E drained', 14: 'E bar1', 15: 'E stores done', 16: 'E bar2',
         20: 'D unit start', 21: 'D tmem_full ok', 22: 'D ybuf free', 23: 'D tmem read', 24: 'D ybuf written',
         30: 'S prefetched', 31: 'S ybuf_full ok', 32: 'S rows stored', 33: 'S records done'}
cur = None; ev = {}
for line in open(sys.argv[1]):
    if line.startswith('==='): cur = line.strip(); ev[cur] = []; continue
    m = re.match(r"PL_TRACE (\w+) i=(\d+) code=(\d+) t=(-?\d+) dt=(-?\d+)", line)
    if m: ev[cur].append((int(m.group(4)), m.group(1), int(m.group(3))))
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for k, v in ev.items():
    print(k); v.sort(); prev = {}
    for t, r, c in v[:lim]:
        print(f"  {t:7d}  +{t-prev.get(r,0):6d}  {r:5s} {names[c]}"); prev[r] = t
