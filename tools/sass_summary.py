"""SASS evidence for profiles/: mnemonic counts per kernel of libspg_b200.so (cuobjdump -sass, no GPU needed).
usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import re
import subprocess
import sys

LIB = "superpoint_graph_b200/libspg_b200.so"
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "LDTM", "STTM", "SYNCS", "REDUX", "HMMA",
         "LDG.E", "STG.E", "ATOM", "RED.", "MEMBAR", "ACQBULK", "PREEXIT")
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
print("# SASS evidence (cuobjdump -sass %s, sm_100a), mnemonic counts per kernel" % LIB)
print("# tcgen05.mma -> UTCHMMA (kind::tf32/f16), tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG, tcgen05.commit -> UTCBAR,")
print("# tcgen05.alloc -> UTCATOMSWS, mbarrier -> SYNCS, redux.sync -> REDUX, griddepcontrol.wait / .launch_dependents -> ACQBULK / PREEXIT;")
print("# HMMA (legacy mma.sync) must be absent.\n")
fn, counts, n = None, None, 0
out = []


def flush():
    if fn is not None:
        out.append((fn, n, dict(counts)))


for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        fn, counts, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn is not None:
        n += 1
        op = m.group(1)
        for wname in WATCH:
            if op.startswith(wname):
                counts[wname] += 1
flush()
hm = 0
for fn, n, c in out:
    hm += c.get("HMMA", 0)
    print(fn)
    print("    instructions %d | %s" % (n, "  ".join("%s x%d" % (k, v) for k, v in sorted(c.items()))))
print("\n# kernels: %d   legacy HMMA instructions in the library: %d" % (len(out), hm))
