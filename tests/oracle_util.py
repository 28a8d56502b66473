"""Access to the CPU oracle for tests (never imported by the product).

port(): oracle/liberi_oracle.so -- our plain-C restatement (built on demand with `make -C oracle port`).
ref():  oracle/_ref/libcint_ref.so -- the unmodified reference compiled by oracle/Makefile, or None when
        it has not been built (it cannot be built on the GPU box: /root/reference does not exist there,
        but the prebuilt .so travels with the repo snapshot).
"""
import ctypes
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_port = None
_ref = None


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def port():
    global _port
    if _port is None:
        so = os.path.join(ORACLE_DIR, "liberi_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
        _port = ctypes.CDLL(so)
        _port.oracle_gto_norm.restype = ctypes.c_double
        _port.oracle_gto_norm.argtypes = [ctypes.c_int, ctypes.c_double]
    return _port


def ref():
    global _ref
    if _ref is None:
        so = os.path.join(ORACLE_DIR, "_ref", "libcint_ref.so")
        if not os.path.exists(so):
            return None
        _ref = ctypes.CDLL(so)
    return _ref


def best():
    """The strongest oracle available: the reference itself if compiled, else the port."""
    return ("ref", ref()) if ref() is not None else ("port", port())


def dims_of(bas, shls, cart=False):
    out = []
    for s in shls:
        l = int(bas[s, 1])
        out.append(((l + 1) * (l + 2) // 2 if cart else 2 * l + 1) * int(bas[s, 3]))
    return out


def eval_tuple(which, name, shls, atm, bas, env, dims=None):
    """name in {int2e_sph,int2e_cart,int3c2e_sph,int3c2e_cart,int2c2e_*,int2e_ip1_*,int3c2e_ip1_*}; returns (flat F-order
    values -- 3 component blocks back to back for the ip1 derivatives --, ret)."""
    atm = np.ascontiguousarray(atm, np.int32)
    bas = np.ascontiguousarray(bas, np.int32)
    env = np.ascontiguousarray(env, np.float64)
    d = dims_of(bas, shls, name.endswith("cart"))
    if name.endswith("_ssc"):               # spherical i, j; Cartesian last (auxiliary) index
        d = d[:-1] + dims_of(bas, shls[-1:], True)
    n = int(np.prod(dims if dims is not None else d)) * (3 if ("_ip1_" in name or "_ip2_" in name) else 1)
    buf = np.zeros(n)
    cs = (ctypes.c_int * len(shls))(*[int(s) for s in shls])
    cd = (ctypes.c_int * len(shls))(*[int(x) for x in dims]) if dims is not None else None
    if which == "ref":
        r = getattr(ref(), name)(_p(buf), cd, cs, _p(atm), len(atm), _p(bas), len(bas), _p(env), None, None)
    else:
        r = getattr(port(), "oracle_" + name)(_p(buf), cd, cs, _p(atm), len(atm), _p(bas), len(bas), _p(env))
    return buf, r


def eval_many(which, name, shls_list, atm, bas, env):
    vals = [eval_tuple(which, name, s, atm, bas, env)[0] for s in shls_list]
    return vals
