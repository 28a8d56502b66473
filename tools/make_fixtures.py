#!/usr/bin/env python
"""Generate the atm/bas/env fixtures of the reference's benchmark drivers (run in the build container).

The unmodified reference examples (examples/time_c60.c, examples/time_c2h6.c) are compiled against
tools/fixture_shim.c, which records the arrays they pass to the library; the dumps are stored as
libcint_b200/data/<name>.npz.  env[0:20] (reserved slots, never initialised by the drivers —
SURVEY Appendix A.13) is zeroed.  Needs /root/reference and oracle/_ref/include/cint.h
(`make -C oracle ref`), therefore cannot run on the GPU box; the .npz files are committed.
"""
import os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CINT_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "libcint_b200", "data")


def dump(example, names):
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "ex")
        subprocess.check_call(["/usr/bin/gcc", "-O1", "-fopenmp", "-w",
                               "-I" + os.path.join(ROOT, "oracle", "_ref", "include"),
                               os.path.join(REF, "examples", example),
                               os.path.join(ROOT, "tools", "fixture_shim.c"), "-lm", "-o", exe])
        env = dict(os.environ, FIXTURE_OUT=os.path.join(tmp, "fx"), OMP_NUM_THREADS="1")
        subprocess.check_call([exe], env=env, stdout=subprocess.DEVNULL)
        for n, name in enumerate(names):
            raw = open(os.path.join(tmp, "fx.%d.bin" % n), "rb").read()
            natm, nbas, nenv = np.frombuffer(raw, np.int32, 3)
            o = 12
            atm = np.frombuffer(raw, np.int32, natm * 6, o).reshape(natm, 6).copy(); o += natm * 24
            bas = np.frombuffer(raw, np.int32, nbas * 8, o).reshape(nbas, 8).copy(); o += nbas * 32
            envv = np.frombuffer(raw, np.float64, nenv, o).copy()
            envv[:20] = 0.0
            # the drivers leave unused atm/bas slots uninitialised; keep only the defined ones
            atm[:, [2, 3, 4, 5]] = 0
            bas[:, [4, 7]] = 0
            np.savez_compressed(os.path.join(OUT, name + ".npz"), atm=atm, bas=bas, env=envv)
            print(name, "natm", natm, "nbas", nbas, "nenv", nenv)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    dump("time_c60.c", ["c60_ccpvdz"])
    dump("time_c2h6.c", ["c2h6_631g", "c2h6_6311gss", "c2h6_ccpvdz", "c2h6_ccpvtz", "c2h6_ccpvqz"])
