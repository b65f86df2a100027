"""Backward branches (loops) of one kernel with their static length and opcode mix.
usage: python tools/sass_loops.py file.o mangled-kernel-name [min-length]"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", "-fun", sys.argv[2], sys.argv[1]], capture_output=True, text=True).stdout
ins = []
for l in out.splitlines():
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 40
for addr, t in ins:
    m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < addr:
        a = int(m.group(1), 16)
        body = [x for x in ins if a <= x[0] <= addr]
        if len(body) < minlen:
            continue
        c = collections.Counter()
        for _, x in body:
            tt = x.split()
            op = tt[1] if tt[0].startswith("@") else tt[0]
            c[op.split(".")[0]] += 1
        fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
        print("loop %#x..%#x: %d instructions, %d FP64-pipe" % (a, addr, len(body), fp64))
        print("   " + ", ".join("%s %d" % kv for kv in c.most_common(16)))
