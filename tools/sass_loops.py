#!/usr/bin/env python
"""Loop-level opcode histogram of one kernel in a built library (cuobjdump -sass), for judging instruction economy
without a GPU:  python tools/sass_loops.py multigrid-petsc_b200/lib/libmgb200.so _Z8k_jfusedILi3ELi0ELi1EEv9FusedArgs"""
import collections
import re
import subprocess
import sys


def main():
    lib, fn = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout
    ins = []
    for line in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: k for k, (a, _) in enumerate(ins)}
    print(f"{fn}: {len(ins)} instructions")
    loops = []
    for k, (a, s) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", s)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], k))
    for lo, hi in loops:
        body = ins[lo:hi + 1]
        h = collections.Counter()
        for _, s in body:
            s = re.sub(r"^@!?U?P\d+\s+", "", s)
            op = s.split()[0].split(".")[0]
            h[op] += 1
        n = len(body)
        if n < 50:
            continue
        print(f"loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {n} instr: " + ", ".join(f"{k}={v}" for k, v in h.most_common(18)))


if __name__ == "__main__":
    main()
