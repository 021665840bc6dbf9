"""Static instruction mix of a kernel's hot loop from `cuobjdump -sass` (no GPU needed).

    python scripts/sass_loop_stats.py build/csrc/apply_tiled.o 'ILb1ELb0ELb0ELi32ELi8ELb0ELi0ELb0E'

Picks the function whose mangled name contains the pattern, finds the innermost backward branch that encloses a
BAR.SYNC (the plane loop of the tiled kernel; unrolled by 2 in the diagonal variants) and prints the opcode histogram
of that address range.  Used to compare variants of the plane loop before spending GPU time."""
import collections
import re
import subprocess
import sys


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    cur, funcs = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return funcs


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    funcs = functions(obj)
    names = [n for n in funcs if pat in n]
    assert len(names) == 1, names
    ins = funcs[names[0]]
    bars = [a for a, t in ins if "BAR.SYNC" in t]
    loops = []
    for a, t in ins:
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            nb = sum(lo <= b <= a for b in bars)
            if nb:
                loops.append((a - lo, lo, a, nb))
    loops.sort()
    print(f"{names[0]}: {len(ins)} instructions, {len(bars)} BAR.SYNC; loops enclosing a barrier: "
          f"{[(hex(lo), hex(hi), nb) for _, lo, hi, nb in loops]}")
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    _, lo, hi, nb = loops[which]
    body = [t for a, t in ins if lo <= a <= hi]
    hist = collections.Counter()
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t.split()[0]
        hist[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "BAR", "SYNCS", "UBLKCP")) and "." in op else "")] += 1
    tot = len(body)
    print(f"loop {hex(lo)}..{hex(hi)}: {tot} instructions, {nb} barrier(s) -> {tot / nb:.0f} per plane")
    fp64 = sum(v for k, v in hist.items() if k.startswith(("DFMA", "DMUL", "DADD")))
    print(f"  FP64 {fp64} ({fp64 / nb:.0f}/plane), other {tot - fp64} ({(tot - fp64) / nb:.0f}/plane)")
    for k, v in hist.most_common(40):
        print(f"  {k:14s} {v:4d}  {v / nb:6.1f}/plane")


if __name__ == "__main__":
    main()
