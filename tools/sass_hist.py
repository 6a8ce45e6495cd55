"""SASS opcode histogram per kernel: python tools/sass_hist.py <lib.so|file.sass> <kernel-substring> [top]"""
import re, subprocess, sys, collections
src, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
text = open(src).read() if src.endswith(".sass") else subprocess.run(["cuobjdump", "-sass", src], capture_output=True, text=True).stdout
cur, hist = None, collections.defaultdict(collections.Counter)
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur and pat in cur:
        op = m.group(1)
        if op != "NOP":
            hist[cur][op] += 1
for k, h in hist.items():
    tot = sum(h.values())
    print("== %s: %d instructions" % (k, tot))
    for op, c in h.most_common(top):
        print("   %5d %s" % (c, op))
