"""SASS evidence for the two headline kernels: opcode histogram + the TMA / mbarrier instructions in context.
usage: python tools/sass_excerpt.py > profiles/r02_sass_headline_kernels.txt   (runs here: cuobjdump needs no GPU)"""
import collections
import re
import subprocess
import sys

LIB = "hpgeom_b200/libhpgb.so"
WANT = [("angle_to_pixel nest lonlat degrees (bench.py headline)", "k_streamI", "OpAng2PixILb1ELb1ELb1ELb0ELb1EEELi3"),
        ("nest_to_ring (bench.py headline)", "k_streamI", "OpReorderILb1E")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur:
        funcs[cur].append(line)
ins = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);")
print("cuobjdump -sass %s : %d sm_100a functions" % (LIB, len(funcs)))
tot = collections.Counter()
for body in funcs.values():
    for l in body:
        m = ins.match(l)
        if m:
            op = re.sub(r"^@!?U?P\w+\s+", "", m.group(2)).split()[0]
            tot[op.split(".")[0]] += 1
print("whole library: UBLKCP (1-D TMA bulk copy) %d, SYNCS (mbarrier) %d, DFMA+DADD+DMUL %d, HMMA/UTCMMA %d (no contraction on this path)\n"
      % (tot["UBLKCP"], tot["SYNCS"], tot["DFMA"] + tot["DADD"] + tot["DMUL"], tot["HMMA"] + tot["UTCHMMA"]))
for title, a, b in WANT:
    names = [n for n in funcs if a in n and b in n]
    if not names:
        print("!! no function matches", a, b)
        continue
    name = names[0]
    lines = [(m.group(1), m.group(2)) for m in (ins.match(l) for l in funcs[name]) if m]
    hist = collections.Counter(re.sub(r"^@!?U?P\w+\s+", "", t).split()[0].split(".")[0] for _, t in lines)
    print("=" * 100)
    print("%s\n  %s\n  %d SASS instructions" % (title, name, len(lines)))
    print("  opcode histogram: " + ", ".join("%s %d" % kv for kv in hist.most_common(24)))
    print("  TMA / mbarrier / streaming-store instructions in place:")
    for i, (addr, t) in enumerate(lines):
        if re.search(r"UBLKCP|SYNCS|STG\.E\.EF|ELECT|FENCE", t):
            print("    /*%s*/  %s" % (addr, t))
