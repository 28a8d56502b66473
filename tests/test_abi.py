"""CPU test: the C-ABI shared library loads and exports every function include/*.h declares."""
import ctypes
import os
import re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in ("cint.h", "cint_b200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        src = re.sub(r"typedef\s[^;{]*\(\s*\*\s*\w+\s*\)[^;]*;", "", src)       # function-pointer typedefs are not symbols
        for m in re.finditer(r"\b(?:CINTIntegralFunction|CINTOptimizerFunction)\s+(\w+)\s*;", src):
            names.add(m.group(1))
        for m in re.finditer(r"^[A-Za-z_][\w \*]*?\b(\w+)\s*\([^;{]*\)\s*;", src, flags=re.M):
            if m.group(1) not in ("CINTIntegralFunction", "CINTOptimizerFunction"):
                names.add(m.group(1))
    return sorted(names)


def test_headers_declare_the_hot_path():
    names = declared_functions()
    for must in ("int2e_sph", "int2e_cart", "int2e_optimizer", "int3c2e_sph", "cint2e_sph", "CINTdel_optimizer",
                 "cintb200_create", "cintb200_int2e_batch", "cintb200_int3c2e_batch", "cintb200_int2c2e_batch", "int2c2e_sph", "cintb200_int2e_sph_block", "cintb200_int3c2e_sph_block", "cintb200_int2c2e_sph_block", "int2e_ip1_sph", "int3c2e_ip1_sph", "cintb200_int2e_ip1_batch", "CINTgto_norm",
                 "cintb200_int2e_sph_all_unique_tiles", "cintb200_int2e_sph_jk", "cintb200_job_checksums", "cintb200_job_row_map",
                 "CINTcgto_spheric", "CINTtot_cgto_spheric"):
        assert must in names, must


def test_library_exports_every_declared_symbol():
    import libcint_b200
    assert os.path.exists(libcint_b200.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(libcint_b200.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_host_helpers_without_gpu():
    # pure host bookkeeping entry points (src/cint_bas.c, src/misc.c:86) need no device
    import numpy as np
    import libcint_b200
    from libcint_b200.basis import gto_norm
    lib = libcint_b200.load_library()
    atm, bas, env = libcint_b200.load_fixture("c60_ccpvdz")
    assert lib.CINTtot_cgto_spheric(bas.ctypes.data_as(ctypes.c_void_p), 300) == 840
    assert lib.CINTtot_pgto_spheric(bas.ctypes.data_as(ctypes.c_void_p), 300) == 1560
    assert lib.CINTcgto_spheric(4, bas.ctypes.data_as(ctypes.c_void_p)) == 5
    for l, a in ((0, 0.3), (1, 2.5), (2, 0.55), (4, 11.0)):
        assert abs(lib.CINTgto_norm(l, a) / gto_norm(l, a) - 1) < 1e-14
    # CINTgto_norm(0, a) = (2a/pi)^(3/4) * 2 sqrt(pi)  (normalised s Gaussian over r^2 dr)
    assert abs(lib.CINTgto_norm(0, 1.0) - 2 * np.pi ** .5 * (2 / np.pi) ** .75) < 1e-14


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import libcint_b200
    atm, bas, env = libcint_b200.load_fixture("c2h6_631g")
    with pytest.raises(libcint_b200.B200Error):
        libcint_b200.Context(atm, bas, env)
    out, rc = libcint_b200.int2e_sph((0, 0, 0, 0), atm, bas, env)
    assert rc == 0


def test_error_codes_without_context():
    # argument validation happens before any device work: a NULL / foreign context is rejected with CINTB200_EINVAL (-2)
    # and a readable message, for every batched entry point (no GPU needed)
    import numpy as np
    import libcint_b200
    lib = libcint_b200.load_library()
    shls = np.zeros(8, np.int32)
    out = np.zeros(16)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for name in ("cintb200_int2e_batch", "cintb200_int3c2e_batch", "cintb200_int2c2e_batch", "cintb200_int2e_ip1_batch",
                 "cintb200_int3c2e_ip1_batch", "cintb200_int3c2e_ip2_batch", "cintb200_int2c2e_ip1_batch", "cintb200_int2c2e_ip2_batch"):
        f = getattr(lib, name)
        f.restype = ctypes.c_long
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        assert f(None, 0, p(shls), 1, None, p(out), 0, None) == -2, name
        assert b"context" in lib.cintb200_last_error()
    for name in ("cintb200_int2e_sph_block", "cintb200_int3c2e_sph_block", "cintb200_int2c2e_sph_block"):
        f = getattr(lib, name)
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        assert f(None, p(shls), p(out), 0, None) == -2, name
    stats = np.zeros(16)
    lib.cintb200_int2e_sph_all_unique.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
    lib.cintb200_int2e_sph_all_unique.restype = ctypes.c_int
    assert lib.cintb200_int2e_sph_all_unique(None, 0, 1, 0, None, p(stats)) == -2
    # host-only planning validates its arguments too
    atm, bas, env = libcint_b200.load_fixture("c2h6_631g")
    with pytest.raises(libcint_b200.B200Error):
        libcint_b200.plan_summary(atm, bas, env, rank=3, nranks=2)
    with pytest.raises(libcint_b200.B200Error):
        libcint_b200.plan_summary(atm, bas, env, aux_shell0=len(bas) + 5)


def _build_example(tmp_path):
    import subprocess
    import libcint_b200
    exe = os.path.join(str(tmp_path), "fill_block")
    libdir = os.path.dirname(libcint_b200.LIB_PATH)
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "fill_block.c"),
                           "-L" + libdir, "-l:libcint_b200.so", "-lm", "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=libdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)


def test_c_example_links_against_the_headers(tmp_path):
    # examples/fill_block.c is the C-side binding shown in INTEGRATION.md: it must compile against include/*.h with -Wall
    # -Werror and link; without a GPU it has to stop with the library's own error message (no CPU path)
    import torch
    r = _build_example(tmp_path)
    if not torch.cuda.is_available():
        assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_c_example_runs_on_the_gpu(tmp_path):
    r = _build_example(tmp_path)
    assert r.returncode == 0, r.stderr
    assert "nonzero = 1" in r.stdout and "(00|00) =" in r.stdout
