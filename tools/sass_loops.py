#!/usr/bin/env python
"""Static instruction mix of the loops of one kernel (no GPU needed).

usage: cuobjdump -sass some.o > x.sass; python tools/sass_loops.py x.sass <mangled-name-substring>

Finds every backward branch of the kernel, treats [target, branch] as a loop body and prints the opcode histogram of the
innermost ones (FP64 pipe: DFMA/DMUL/DADD/DSETP/MUFU.RSQ64H..., shared loads, local (spill) traffic).  Used to compare the
issued FP64 instructions per primitive quartet with the SURVEY 8(d) FLOP model before spending GPU time."""
import re
import sys
from collections import Counter


def main():
    path, pat = sys.argv[1], sys.argv[2]
    txt = open(path).read()
    funcs = re.split(r"\n\s*Function : ", txt)
    body = None
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if pat in name:
            body = f
            print("kernel", name)
            break
    if body is None:
        sys.exit("no kernel matches " + pat)
    ins = []
    for line in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    addr2idx = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr2idx:
                loops.append((addr2idx[tgt], i))
    print("instructions", len(ins), "loops", len(loops))
    for (b, e) in sorted(loops, key=lambda x: x[1] - x[0]):
        inner = [l for l in loops if l != (b, e) and l[0] >= b and l[1] <= e]
        ops = Counter()
        for _, t in ins[b:e + 1]:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            op = t.split()[0]
            ops[op] += 1
        n = e - b + 1
        fp64 = sum(v for k, v in ops.items() if k.startswith(("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")))
        lds = sum(v for k, v in ops.items() if k.startswith("LDS"))
        ldl = sum(v for k, v in ops.items() if k.startswith(("LDL", "STL")))
        ldg = sum(v for k, v in ops.items() if k.startswith(("LDG", "STG", "LD.", "ST.")))
        print("loop [%d..%d] n=%d nested_loops=%d  fp64=%d (DFMA %d DMUL %d DADD %d)  LDS=%d  local=%d  global=%d  other=%d" % (
            b, e, n, len(inner), fp64, sum(v for k, v in ops.items() if k.startswith("DFMA")),
            sum(v for k, v in ops.items() if k.startswith("DMUL")), sum(v for k, v in ops.items() if k.startswith("DADD")),
            lds, ldl, ldg, n - fp64 - lds - ldl - ldg))
        if len(sys.argv) > 3 and not inner:
            print("   ", dict(ops.most_common(25)))


if __name__ == "__main__":
    main()
