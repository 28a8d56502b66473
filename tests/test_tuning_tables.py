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
