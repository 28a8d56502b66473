#!/usr/bin/env python
"""GPU ms of the C60 whole job for the library selected by CINTB200_LIB (3 runs after warm-up)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libcint_b200 as cb
gb = int(sys.argv[1]) if len(sys.argv) > 1 else 80
name = sys.argv[2] if len(sys.argv) > 2 else "c60_ccpvdz"
atm, bas, env = cb.load_fixture(name)
ctx = cb.Context(atm, bas, env)
ms = []
for k in range(6):
    st = ctx.all_unique(chunk_bytes=gb << 30)
    if k >= 3:
        ms.append(float(st[7]))
print(json.dumps({"lib": os.environ.get("CINTB200_LIB", "default"), "pblocks": os.environ.get("CINTB200_PBLOCKS", "16"), "chunk_gb": gb, "name": name, "graph": os.environ.get("CINTB200_NO_GRAPH", "0") != "1", "ms": ms, "launches": int(st[4])}))
