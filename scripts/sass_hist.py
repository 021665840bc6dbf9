"""Opcode histogram of an address range of a kernel's SASS (no GPU needed): python scripts/sass_hist.py OBJ PATTERN [lo hi]
Without a range: lists backward branches (loop candidates) with their spans."""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
cur, funcs = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
name = [f for f in funcs if pat in f][0]
ins = funcs[name]
print(name, len(ins), "instructions")
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    body = [i for a, i in ins if lo <= a <= hi]
    h = collections.Counter()
    for i in body:
        t = i.split()
        op = t[1] if t[0].startswith('@') else t[0]
        h[op.split('.')[0]] += 1
    print(len(body), h.most_common())
else:
    for a, i in ins:
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?(0x[0-9a-f]+)", i)
        if m and int(m.group(1), 16) < a:
            print(hex(int(m.group(1), 16)), hex(a), (a - int(m.group(1), 16)) // 16 + 1, i)
