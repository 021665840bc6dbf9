#!/usr/bin/env python
"""registers / spills per instantiation from a build/csrc/*.ptxas.log (nvcc -Xptxas -v)"""
import re
import sys
t = open(sys.argv[1] if len(sys.argv) > 1 else "build/csrc/apply_rowpair.ptxas.log").read()
pat = r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers"
for n, st, ss, sl, r in re.findall(pat, t):
    m = re.search(r"kernelI(.*?)EEvNS", n)
    print((m.group(1) if m else n[-70:]).ljust(50), "stack", st, "spill", ss, sl, "regs", r)
