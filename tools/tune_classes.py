"""Per-class tuning of the register / cooperative kernels (run on the GPU box).

For every library variant (libcint_b200.so = current defaults, libcint_b200_v?.so = one-knob builds of
tools/build_variants.sh) the whole job is run with per-launch CUDA events (serialised) and the time of each kernel class is
recorded; the table of all variants goes to gpurun_out/tune_<job>.json.  `python tools/tune_classes.py pick ...` (run anywhere)
turns such tables into csrc/tune_reg.inc / tune_coop.inc.

  python tools/tune_classes.py run c60_ccpvdz|df|c2h6_ccpvqz [reps]      (GPU)
  python tools/tune_classes.py pick gpurun_out/tune_*.json                (CPU)
"""
import sys, os, json, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {"base": "", "va": "_va", "vb": "_vb", "vc": "_vc", "vd": "_vd", "ve": "_ve", "alt": "", "fc": "_fc"}
VARIANT_ENV = {"alt": {"CINTB200_COOP_ALT": "1"}}       # same library, other orientation of the cooperative kernels (GEN_COOP_ALT=1 builds)
# knob values of each variant: (reg minb, reg unroll, coop minb for nacc <= 16 / <= 32 / above)
KNOBS = {"base": (2, 1, (4, 3, 2)), "va": (3, 1, (3, 2, 2)), "vb": (4, 1, (5, 4, 3)), "vc": (2, 2, (6, 5, 4)),
         "vd": (3, 2, (4, 3, 2)), "ve": (4, 2, (4, 3, 2))}


def worker(job, reps):
    import numpy as np
    import ctypes
    import libcint_b200 as cb
    if job == "df":
        from libcint_b200.basis import c60_df_basis
        atm, bas, env, norb = c60_df_basis()
    else:
        atm, bas, env = cb.load_fixture(job)
    ctx = cb.Context(atm, bas, env)
    CH = 80 << 30 if job in ("df", "c60_ccpvdz") else 8 << 30
    run = (lambda: ctx.int3c2e_all(norb, chunk_bytes=CH)) if job == "df" else (lambda: ctx.all_unique(chunk_bytes=CH))
    run()
    best = {}
    for _ in range(reps):
        ctx.lib.cintb200_debug_profile(ctx.handle, 1)
        run()
        ctx.lib.cintb200_debug_profile(ctx.handle, 0)
        n = ctx.lib.cintb200_debug_profile_rows(ctx.handle, None, 0)
        rows = np.zeros((n, 12))
        ctx.lib.cintb200_debug_profile_rows(ctx.handle, rows.ctypes.data_as(ctypes.c_void_p), n)
        for r in rows:
            k = "%d %d %d %d %d %d %d" % tuple(int(v) for v in r[:7])
            best[k] = min(best.get(k, 1e30), float(r[7]))
    print("TUNE_JSON " + json.dumps(best))


def run(job, reps):
    out = {}
    for name, suffix in VARIANTS.items():
        lib = os.path.join(ROOT, "libcint_b200", "libcint_b200%s.so" % suffix)
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, CINTB200_LIB=lib, **VARIANT_ENV.get(name, {}))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "worker", job, str(reps)], env=env, capture_output=True, text=True, timeout=900)
        m = re.search(r"TUNE_JSON (.*)", p.stdout)
        if not m:
            print(name, "FAILED", p.stdout[-300:], p.stderr[-600:])
            continue
        out[name] = json.loads(m.group(1))
        print(name, "total %.1f ms" % sum(out[name].values()))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_%s%s.json" % (job, os.environ.get("TUNE_TAG", ""))), "w"))
    show(out)


def show(out):
    base = out["base"]
    names = [n for n in out if n != "base"]
    print("%-22s %9s " % ("class (T|U nct ncu kind)", "base ms") + " ".join("%8s" % n for n in names))
    for k in sorted(base, key=lambda k: -base[k]):
        print("%-22s %9.2f " % (k, base[k]) + " ".join("%8.3f" % (out[n].get(k, float("nan")) / base[k]) for n in names))


def coop_template_args():
    """(tla tlb ula ulb nct ncu) -> the cooperative kernel's own template arguments (la lb lc ld ncr ncl), from the generated table"""
    m = {}
    for line in open(os.path.join(ROOT, "libcint_b200", "csrc", "kern_coop_table.cu")):
        g = re.match(r"\s*\{\{(\d+),(\d+),(\d+),(\d+),(\d+),(\d+)\}, eri_coop_kernel<(\d+),(\d+),(\d+),(\d+),(\d+),(\d+),(\d+),(\w+)>", line)
        if g:
            v = [int(x) for x in g.groups()[:13]]
            m[tuple(v[:6])] = tuple(v[6:12])
    return m


def nrange(l0, l1):
    return sum((l + 1) * (l + 2) // 2 for l in range(l0, l1 + 1))


def pick(files, margin=0.03, min_ms=1.0):
    """Choose per class the knob of the fastest variant (when it beats the default by more than `margin`); the first file that
    contains a class decides (give the most important job first)."""
    cmap = coop_template_args()
    reg, coop, gain = {}, {}, 0.0
    for fn in files:
        out = json.load(open(fn))
        base = out["base"]
        for k, tb in base.items():
            la, lb, lc, ld, nct, ncu, kind = (int(x) for x in k.split())
            if kind not in (1, 2):
                continue
            key = (la, lb, lc, ld, nct, ncu) if kind == 1 else cmap.get((la, lb, lc, ld, nct, ncu))
            if key is None or key in reg or key in coop:
                continue
            cands = {n: out[n][k] for n in out if k in out[n]}
            bestn = min(cands, key=cands.get)
            if bestn == "base" or cands[bestn] > (1 - margin) * tb or tb < min_ms:
                (reg if kind == 1 else coop)[key] = None
                continue
            gain += tb - cands[bestn]
            if kind == 1:
                reg[key] = (KNOBS[bestn][0], KNOBS[bestn][1], "%s %.2f -> %.2f ms (%s)" % (os.path.basename(fn), tb, cands[bestn], bestn))
            else:
                nacc = key[4] * key[5] * nrange(key[0], key[0] + key[1])
                mb = KNOBS[bestn][2][0 if nacc <= 16 else 1 if nacc <= 32 else 2]
                dflt = KNOBS["base"][2][0 if nacc <= 16 else 1 if nacc <= 32 else 2]
                if mb == dflt:
                    coop[key] = None
                    continue
                coop[key] = (mb, 1, "%s %.2f -> %.2f ms (%s)" % (os.path.basename(fn), tb, cands[bestn], bestn))
    for name, tab in (("tune_reg.inc", reg), ("tune_coop.inc", coop)):
        with open(os.path.join(ROOT, "libcint_b200", "csrc", name), "w") as f:
            for key in sorted(k for k in tab if tab[k]):
                mb, un, why = tab[key]
                f.write("    {tune_key(%d, %d, %d, %d, %d, %d), %d, %d},      // %s\n" % (key + (mb, un, why)))
    print("entries: reg %d coop %d; summed gain of the first-listed jobs %.1f ms" % (sum(1 for v in reg.values() if v), sum(1 for v in coop.values() if v), gain))


if __name__ == "__main__":
    if sys.argv[1] == "worker":
        worker(sys.argv[2], int(sys.argv[3]))
    elif sys.argv[1] == "run":
        run(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 2)
    elif sys.argv[1] == "show":
        show(json.load(open(sys.argv[2])))
    elif sys.argv[1] == "pick":
        pick(sys.argv[2:])
