"""Decode the clock64 timeline of CTA 0 of a fused chain kernel (COPE_FZ_TIMELINE=<mode> COPE_FZ_TIMELINE_FILE=<file>):
    python tools/chain_timeline.py <file> [max_events]"""
import sys, struct, collections
import numpy as np
d = np.fromfile(sys.argv[1], dtype=np.int64).reshape(2, 4096)
names = {1: "job", 2: "a_ready", 3: "w_full", 4: "commit", 5: "acc_ok/evt/aux", 6: "panel_done"}
ev = []
for role in (0, 1):
    n = int(d[role, 0])
    for k in range(1, n):
        v = int(d[role, k]); tag = v >> 48; t = v & 0xFFFFFFFFFFFF
        ev.append((t, role, tag))
ev.sort()
t0 = ev[0][0]
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 400
prev = {0: t0, 1: t0}
for t, role, tag in ev[:lim]:
    kind = tag // 100
    label = {1: f"MMA job{tag-100} start", 2: f"MMA got panel {tag-200}", 3: f"MMA got W chunk {tag-300}", 4: f"MMA committed job{tag-400}",
             5: {500: "EPI acc0 ready", 501: "EPI acc1 ready", 510: "EPI a_free ok", 520: "EPI aux ok"}.get(tag, str(tag)), 6: f"EPI panel {tag-600} done"}[kind]
    print(f"{t-t0:9d} (+{t-prev[role]:6d}) {'    ' if role else ''}{label}")
    prev[role] = t
