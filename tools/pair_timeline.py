"""Decode the clock64 timeline of a two-tiles-in-flight chain kernel (COPE_Q2_TIMELINE=<file>):
    python tools/pair_timeline.py <file> [max_events]"""
import sys, numpy as np
d = np.fromfile(sys.argv[1], dtype=np.int64).reshape(2, 4096)
ev = []
for role in (0, 1):
    n = int(d[role, 0])
    for k in range(1, n):
        v = int(d[role, k]); ev.append((v & 0xFFFFFFFFFFFF, role, v >> 48))
ev.sort(); t0 = ev[0][0]; prev = {0: t0, 1: t0}
names = {1: "MMA start", 2: "MMA commit", 3: "    EPI got acc", 4: "    EPI done", 5: "        EPI panel"}
for t, role, tag in ev[:int(sys.argv[2]) if len(sys.argv) > 2 else 200]:
    k, l, s = tag // 1000, (tag % 1000) // 10, tag % 10
    what = f"panel {s} layer {l}" if k == 5 else f"tile {'AB'[s]} layer {l}"
    print(f"{t-t0:9d} (+{t-prev[role]:6d}) {names[k]} {what}")
    prev[role] = t
