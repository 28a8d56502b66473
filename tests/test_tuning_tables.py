"""The per-class tuning tables (csrc/tune_reg.inc, tune_coop.inc) must refer to kernel classes that are actually instantiated
(tools/gen_reg_inst.py writes the instantiation tables; tools/tune_classes.py writes the tuning tables from GPU measurements)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "libcint_b200", "csrc")


def tune_entries(name):
    out = []
    for line in open(os.path.join(CSRC, name)):
        m = re.match(r"\s*\{tune_key\((\d+), (\d+), (\d+), (\d+), (\d+), (\d+)\), (\d+), (\d+)\}", line)
        if m:
            v = [int(x) for x in m.groups()]
            out.append((tuple(v[:6]), v[6], v[7]))
    return out


def test_register_tuning_entries_name_instantiated_classes():
    inst = set()
    for line in open(os.path.join(CSRC, "kern_reg_table.cu")):
        m = re.match(r"\s*\{\{(\d+),(\d+),(\d+),(\d+),(\d+),(\d+)\}, eri_reg_kernel<", line)
        if m:
            inst.add(tuple(int(x) for x in m.groups()))
    entries = tune_entries("tune_reg.inc")
    assert entries, "tune_reg.inc is empty"
    keys = [k for k, _, _ in entries]
    assert len(keys) == len(set(keys)), "duplicate class in tune_reg.inc"
    for key, minb, unroll in entries:
        assert key in inst, "tune_reg.inc names %s, which has no register-kernel instantiation" % (key,)
        assert minb in (1, 2, 3, 4) and unroll in (1, 2)
        assert all(v < 8 for v in key), "tune_key packs 3 bits per field"


def test_cooperative_tuning_entries_name_instantiated_classes():
    inst = set()
    for line in open(os.path.join(CSRC, "kern_coop_table.cu")):
        m = re.match(r"\s*\{\{[\d,]+\}, eri_coop_kernel<(\d+),(\d+),(\d+),(\d+),(\d+),(\d+),(\d+),(\w+)>", line)
        if m:
            inst.add(tuple(int(x) for x in m.groups()[:6]))
    entries = tune_entries("tune_coop.inc")
    keys = [k for k, _, _ in entries]
    assert len(keys) == len(set(keys))
    for key, minb, _ in entries:
        assert key in inst, "tune_coop.inc names %s, which has no cooperative-kernel instantiation" % (key,)
        assert 1 <= minb <= 6


def test_multiply_shift_division_of_the_catch_all_epilogue_is_exact_in_its_range():
    """kern_generic.cu:FastDiv computes n / d as (n * (2^32 / d + 1)) >> 32; generic_plan refuses classes with index x divisor
    >= 2^32.  The formula (restated here) is exact for every n, d with n * d < 2^32 -- checked on the extremes and a random sample."""
    import random
    src = open(os.path.join(CSRC, "kern_generic.cu")).read()
    assert "0x100000000ull / (unsigned)d_) + 1" in src and ">= 4294967296.0) return -1" in src      # the two halves of the contract

    def fastdiv(n, d):
        return (n * ((1 << 32) // d + 1)) >> 32

    rnd = random.Random(7)
    cases = [(0, 1), (1, 1), (614655, 1), (246959, 8820), (194480, 9261), (65535, 65535), ((1 << 32) // 9261 - 1, 9261)]
    for _ in range(20000):
        d = rnd.randint(1, 30000)
        cases.append((rnd.randint(0, ((1 << 32) - 1) // d), d))
    for n, d in cases:
        assert n * d < (1 << 32)
        assert fastdiv(n, d) == n // d, (n, d)
