"""Static SASS size of every kernel in an object file (cuobjdump -sass), optionally an opcode histogram of one kernel.
usage: python tools/sass_count.py file.o [kernel-substring]"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
name, cnt, ops = None, collections.Counter(), collections.defaultdict(collections.Counter)
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        name = m.group(1); continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m and name:
        t = m.group(1).split()
        op = t[1] if t[0].startswith("@") else t[0]
        cnt[name] += 1; ops[name][op.split(".")[0]] += 1
for n, c in sorted(cnt.items(), key=lambda x: -x[1]):
    if len(sys.argv) < 3 or sys.argv[2] in n:
        print("%6d  %s" % (c, n))
        if len(sys.argv) > 2:
            print("        " + ", ".join("%s %d" % kv for kv in ops[n].most_common(25)))
