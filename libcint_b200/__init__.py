"""libcint_b200 -- B200 (sm_100a) implementation of libcint's ERI hot path behind the libcint C ABI.

The product is the shared library ``libcint_b200/libcint_b200.so`` (sources in ``csrc/``, C headers in
``include/``).  This Python module is only the ctypes mirror of how the reference's own tests drive
libcint (``testsuite/test_cint.py:16-20``): same function names, same argument order.

There is no CPU fallback: importing works anywhere, but any compute call raises / returns an error
when the library or a CUDA device is missing.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CINTB200_LIB", os.path.join(_HERE, "libcint_b200.so"))     # override: kernel-variant experiments
DATA_DIR = os.path.join(_HERE, "data")

SPH, CART = 0, 1
_lib = None


class B200Error(RuntimeError):
    pass


def load_library():
    """dlopen the CUDA library; fails loudly if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error("%s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cd, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t
    lib.cintb200_create.argtypes = [ctypes.POINTER(vp), vp, ci, vp, ci, vp, ci]
    lib.cintb200_create.restype = ci
    lib.cintb200_destroy.argtypes = [vp]
    lib.cintb200_destroy.restype = None
    for name in ("cintb200_int2e_batch", "cintb200_int3c2e_batch", "cintb200_int2c2e_batch", "cintb200_int2e_ip1_batch",
                 "cintb200_int3c2e_ip1_batch", "cintb200_int3c2e_ip2_batch", "cintb200_int2c2e_ip1_batch", "cintb200_int2c2e_ip2_batch"):
        f = getattr(lib, name)
        f.argtypes = [vp, ci, vp, sz, vp, vp, ci, vp]
        f.restype = ctypes.c_long
    lib.cintb200_block_size.argtypes = [vp, ci, vp, ci]
    lib.cintb200_block_size.restype = sz
    lib.cintb200_last_error.restype = ctypes.c_char_p
    lib.cintb200_fp64_peak.argtypes = [ci, cd, vp]
    lib.cintb200_fp64_peak.restype = ci
    if hasattr(lib, "cintb200_int2e_sph_all_unique"):
        lib.cintb200_int2e_sph_all_unique.argtypes = [vp, ci, ci, sz, vp, vp]
        lib.cintb200_int2e_sph_all_unique.restype = ci
    if hasattr(lib, "cintb200_debug_chunk"):
        lib.cintb200_debug_chunk.argtypes = [vp, ci, vp, sz, vp]
        lib.cintb200_debug_chunk.restype = ci
        lib.cintb200_debug_pair_offsets.argtypes = [vp, ci, ci, vp, vp]
        lib.cintb200_debug_pair_offsets.restype = ci
        lib.cintb200_debug_force_generic.argtypes = [vp, ci]
        lib.cintb200_debug_force_generic.restype = None
        lib.cintb200_debug_profile.argtypes = [vp, ci]
        lib.cintb200_debug_profile.restype = None
        lib.cintb200_debug_profile_rows.argtypes = [vp, vp, ci]
        lib.cintb200_debug_profile_rows.restype = ci
    for name in ("int2e_sph", "int2e_cart", "int3c2e_sph", "int3c2e_cart", "int2c2e_sph", "int2c2e_cart",
                 "int2e_ip1_sph", "int2e_ip1_cart", "int3c2e_ip1_sph", "int3c2e_ip1_cart", "int3c2e_ip2_sph", "int3c2e_ip2_cart",
                 "int2c2e_ip1_sph", "int2c2e_ip2_sph"):
        f = getattr(lib, name)
        f.argtypes = [vp, vp, vp, vp, ci, vp, ci, vp, vp, vp]
        f.restype = ci
    for name in ("cint2e_sph", "cint2e_cart", "cint3c2e_sph", "cint3c2e_cart", "cint2c2e_sph", "cint2c2e_cart"):
        f = getattr(lib, name)
        f.argtypes = [vp, vp, vp, ci, vp, ci, vp, vp]
        f.restype = ci
    lib.CINTgto_norm.argtypes = [ci, cd]
    lib.CINTgto_norm.restype = cd
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class TileInfo(ctypes.Structure):
    """cintb200_tile of include/cint_b200.h."""
    _fields_ = [("chunk", ctypes.c_int), ("nchunks", ctypes.c_int), ("rank", ctypes.c_int), ("nranks", ctypes.c_int),
                ("i0", ctypes.c_int), ("i1", ctypes.c_int), ("row0", ctypes.c_longlong), ("nrows", ctypes.c_longlong),
                ("ncols", ctypes.c_longlong), ("ncols_below", ctypes.c_longlong)]


TILE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(TileInfo), ctypes.POINTER(ctypes.c_double))


def job_weights(nao, maxrb=1):
    """Weights of the whole-job fingerprints (formulas of oracle/ref_golden.c and csrc/digest.cu): h[r] per block row position,
    g[c, d] per ket AO pair, the formula density D[a, b] and the probe matrix U[nao, 8] of the committed J/K goldens."""
    r = np.arange(maxrb)
    a = np.arange(nao)
    h = np.cos(0.91 * r + 0.3)
    g = np.cos(0.37 * a[:, None] + 0.61 * a[None, :] + 0.5)
    D = np.cos(0.37 * (a[:, None] + a[None, :]) + 0.2) + 0.5 * np.cos(0.11 * (a[:, None] - a[None, :]))
    U = np.cos(0.13 * (a[:, None] + 1) * (np.arange(8)[None, :] + 1) + 0.7)
    return h, g, D, U


def _as_basis(atm, bas, env):
    atm = np.ascontiguousarray(atm, dtype=np.int32).reshape(-1, 6)
    bas = np.ascontiguousarray(bas, dtype=np.int32).reshape(-1, 8)
    env = np.ascontiguousarray(env, dtype=np.float64)
    return atm, bas, env


def shell_dims(bas, shls, cart=False):
    bas = np.asarray(bas).reshape(-1, 8)
    out = []
    for s in shls:
        l = int(bas[s, 1])
        out.append(((l + 1) * (l + 2) // 2 if cart else 2 * l + 1) * int(bas[s, 3]))
    return out


class Context:
    """Device-resident basis + shell-pair tables (the object the C API hands out as ``CINTOpt*``)."""

    def __init__(self, atm, bas, env, device=-1):
        self.lib = load_library()
        self.atm, self.bas, self.env = _as_basis(atm, bas, env)
        h = ctypes.c_void_p()
        rc = self.lib.cintb200_create(ctypes.byref(h), _p(self.atm), len(self.atm), _p(self.bas), len(self.bas),
                                      _p(self.env), device)
        if rc != 0:
            raise B200Error("cintb200_create failed (%d): %s" % (rc, self.lib.cintb200_last_error().decode()))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cintb200_destroy(self.handle)
            self.handle = None

    __del__ = close

    def _batch(self, fn, ncenter, shls, kind, out=None, out_off=None, device_ptr=None, ncomp=1):
        shls = np.ascontiguousarray(shls, dtype=np.int32).reshape(-1, ncenter)
        n = len(shls)
        cart = kind == CART
        l, nc = self.bas[:, 1].astype(np.int64), self.bas[:, 3].astype(np.int64)
        dim = ((l + 1) * (l + 2) // 2 if cart else 2 * l + 1) * nc                 # per shell
        sizes = (ncomp * np.prod(dim[shls], axis=1)).astype(np.uint64) if n else np.zeros(0, np.uint64)
        packed = out_off is None
        if packed:              # blocks back to back in input order: the library computes the same offsets itself (out_off = NULL)
            offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64) if n else np.zeros(0, np.uint64)
            total = int(sizes.sum())
        else:
            offs = np.ascontiguousarray(out_off, dtype=np.uint64)
            total = int((offs + sizes).max()) if n else 0
        nz = np.zeros(n, dtype=np.int32)
        if device_ptr is not None:
            rc = fn(self.handle, kind, _p(shls), n, None if packed else _p(offs), ctypes.c_void_p(device_ptr), 1, _p(nz))
            res = None
        else:
            if out is None:
                out = np.zeros(total)
            rc = fn(self.handle, kind, _p(shls), n, None if packed else _p(offs), _p(out), 0, _p(nz))
            res = out
        if rc < 0:
            raise B200Error("batch failed (%d): %s" % (rc, self.lib.cintb200_last_error().decode()))
        return res, offs.astype(np.int64), sizes.astype(np.int64), nz

    def int2e_batch(self, shls, kind=SPH, **kw):
        """Evaluate many shell quartets; returns (packed values, offsets, sizes, nonzero flags)."""
        return self._batch(self.lib.cintb200_int2e_batch, 4, shls, kind, **kw)

    def int3c2e_batch(self, shls, kind=SPH, **kw):
        return self._batch(self.lib.cintb200_int3c2e_batch, 3, shls, kind, **kw)

    def int2c2e_batch(self, shls, kind=SPH, **kw):
        """Shell pairs (i|k): the 2-centre Coulomb metric of density fitting (src/cint2c2e.c:351)."""
        return self._batch(self.lib.cintb200_int2c2e_batch, 2, shls, kind, **kw)

    def int2e_ip1_batch(self, shls, kind=SPH, **kw):
        """( nabla i j | k l ): 3 components per quartet, block layout [comp][l][k][j][i] (src/autocode/grad2.c:19-68)."""
        return self._batch(self.lib.cintb200_int2e_ip1_batch, 4, shls, kind, ncomp=3, **kw)

    def int3c2e_ip1_batch(self, shls, kind=SPH, **kw):
        return self._batch(self.lib.cintb200_int3c2e_ip1_batch, 3, shls, kind, ncomp=3, **kw)

    def int3c2e_ip2_batch(self, shls, kind=SPH, **kw):
        """( i j | nabla k ), src/autocode/int3c2e.c:161."""
        return self._batch(self.lib.cintb200_int3c2e_ip2_batch, 3, shls, kind, ncomp=3, **kw)

    def int2c2e_ip1_batch(self, shls, kind=SPH, **kw):
        return self._batch(self.lib.cintb200_int2c2e_ip1_batch, 2, shls, kind, ncomp=3, **kw)

    def int2c2e_ip2_batch(self, shls, kind=SPH, **kw):
        return self._batch(self.lib.cintb200_int2c2e_ip2_batch, 2, shls, kind, ncomp=3, **kw)

    def all_unique(self, rank=0, nranks=1, chunk_bytes=0, host_sink=None, cart=False):
        """Whole-job driver of examples/time_c60.c:200-219 on this rank's shard; returns the stats array (cart: int2e_cart)."""
        stats = np.zeros(16)
        f = self.lib.cintb200_int2e_cart_all_unique if cart else self.lib.cintb200_int2e_sph_all_unique
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        f.restype = ctypes.c_int
        rc = f(self.handle, rank, nranks, chunk_bytes, ctypes.c_void_p(host_sink) if host_sink else None, _p(stats))
        if rc < 0:
            raise B200Error("all_unique failed (%d): %s" % (rc, self.lib.cintb200_last_error().decode()))
        return stats


    # ---- consumers of the whole-job tiles ----
    def set_checksums(self, on=True):
        """Reduce every finished tile to per-row sums on the device (cintb200_set_checksums); stats[3] = sum of all integrals."""
        self.lib.cintb200_set_checksums.argtypes = [ctypes.c_void_p, ctypes.c_int]
        self.lib.cintb200_set_checksums(self.handle, int(on))

    def job_checksums(self):
        """(S, A, F) per bra shell pair i(i+1)/2+j of the last whole-job run with checksums on: this rank's partial sums."""
        f = self.lib.cintb200_job_checksums
        f.argtypes = [ctypes.c_void_p] * 4
        f.restype = ctypes.c_int
        n = f(self.handle, None, None, None)
        if n < 0:
            raise B200Error(self.lib.cintb200_last_error().decode())
        S, A, F = np.zeros(n), np.zeros(n), np.zeros(n)
        f(self.handle, _p(S), _p(A), _p(F))
        return S, A, F

    def job_geometry(self, chunk=0):
        g = np.zeros(9, dtype=np.int64)
        self.lib.cintb200_job_geometry.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        if self.lib.cintb200_job_geometry(self.handle, chunk, _p(g)) != 0:
            raise B200Error(self.lib.cintb200_last_error().decode())
        return dict(zip(("i0", "i1", "row0", "nrows", "ncols", "nchunks", "total_rows", "total_cols", "ncols_below"), (int(v) for v in g)))

    def job_maps(self, chunk):
        """((sh_i, sh_j, pos) per row, (sh_k, sh_l, pos) per column) of chunk `chunk` of the cached whole-job plan."""
        g = self.job_geometry(chunk)
        rows = [np.zeros(g["nrows"], dtype=np.int32) for _ in range(3)]
        cols = [np.zeros(g["ncols"], dtype=np.int32) for _ in range(3)]
        for fn, arrs in ((self.lib.cintb200_job_row_map, rows), (self.lib.cintb200_job_col_map, cols)):
            fn.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3
            if fn(self.handle, chunk, *[_p(a) for a in arrs]) != 0:
                raise B200Error(self.lib.cintb200_last_error().decode())
        return rows, cols

    def all_unique_tiles(self, sinks, callback=None, rank=0, nranks=1, chunk_bytes=0, aux_shell0=None, cart=False):
        """Whole job with every tile delivered to the host: `sinks` = list of pinned-buffer addresses (ints), `callback(tile
        dict, values ndarray[nrows, ncols] F-order view into the sink)` is called once per tile in chunk order."""
        err = []

        def tramp(user, tile, values):
            try:
                t = tile.contents
                info = {k: getattr(t, k) for k, _ in TileInfo._fields_}
                if callback is not None:
                    n = info["nrows"] * info["ncols"]
                    arr = np.ctypeslib.as_array(values, shape=(n,)).reshape((info["nrows"], info["ncols"]), order="F") if n else np.zeros((0, 0))
                    callback(info, arr)
                return 0
            except Exception as e:          # never unwind through C
                err.append(e)
                return 1
        cb = TILE_FN(tramp)
        arr = (ctypes.c_void_p * len(sinks))(*sinks)
        stats = np.zeros(16)
        if aux_shell0 is None:
            f = self.lib.cintb200_int2e_cart_all_unique_tiles if cart else self.lib.cintb200_int2e_sph_all_unique_tiles
            f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, TILE_FN, ctypes.c_void_p, ctypes.c_void_p]
            f.restype = ctypes.c_int
            rc = f(self.handle, rank, nranks, chunk_bytes, arr, len(sinks), cb, None, _p(stats))
        else:
            f = self.lib.cintb200_int3c2e_sph_all_tiles
            f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, TILE_FN, ctypes.c_void_p, ctypes.c_void_p]
            f.restype = ctypes.c_int
            rc = f(self.handle, aux_shell0, rank, nranks, chunk_bytes, arr, len(sinks), cb, None, _p(stats))
        if err:
            raise err[0]
        if rc < 0:
            raise B200Error("all_unique_tiles failed (%d): %s" % (rc, self.lib.cintb200_last_error().decode()))
        return stats

    def jk(self, dm=None, rank=0, nranks=1, chunk_bytes=0, with_k=True, device_ptrs=None):
        """Coulomb / exchange matrices digested on the device (cintb200_int2e_sph_jk): returns (vj, vk, stats) -- this rank's
        PARTIAL matrices when nranks > 1.  device_ptrs = (dm, vj, vk) device addresses -> results stay on the device."""
        f = self.lib.cintb200_int2e_sph_jk
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        f.restype = ctypes.c_int
        stats = np.zeros(16)
        if device_ptrs is not None:
            d, j, k = device_ptrs
            rc = f(self.handle, rank, nranks, chunk_bytes, ctypes.c_void_p(d), ctypes.c_void_p(j), ctypes.c_void_p(k) if (k and with_k) else None, 1, _p(stats))
            vj = vk = None
        else:
            dm = np.ascontiguousarray(dm, dtype=np.float64)
            vj = np.zeros_like(dm)
            vk = np.zeros_like(dm) if with_k else None
            rc = f(self.handle, rank, nranks, chunk_bytes, _p(dm), _p(vj), _p(vk) if with_k else None, 0, _p(stats))
        if rc < 0:
            raise B200Error("jk failed (%d): %s" % (rc, self.lib.cintb200_last_error().decode()))
        return vj, vk, stats

    def int3c2e_all(self, aux_shell0, rank=0, nranks=1, chunk_bytes=0, host_sink=None):
        """Whole density-fitting job: every (ij|k), orbital shells i >= j < aux_shell0 <= k (cintb200_int3c2e_sph_all)."""
        f = self.lib.cintb200_int3c2e_sph_all
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        f.restype = ctypes.c_int
        stats = np.zeros(16)
        rc = f(self.handle, aux_shell0, rank, nranks, chunk_bytes, ctypes.c_void_p(host_sink) if host_sink else None, _p(stats))
        if rc < 0:
            raise B200Error("int3c2e_all failed (%d): %s" % (rc, self.lib.cintb200_last_error().decode()))
        return stats

    def _block(self, fn, ncenter, shls_slice, device_ptr=None, cart=False):
        sl = np.ascontiguousarray(shls_slice, dtype=np.int32).reshape(-1)
        assert sl.size == 2 * ncenter
        ao = np.concatenate([[0], np.cumsum([((int(b[1]) + 1) * (int(b[1]) + 2) // 2 if cart else 2 * int(b[1]) + 1) * int(b[3]) for b in self.bas])])
        shape = tuple(int(ao[sl[2 * m + 1]] - ao[sl[2 * m]]) for m in range(ncenter))
        stats = np.zeros(16)
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        fn.restype = ctypes.c_int
        if device_ptr is not None:
            rc = fn(self.handle, _p(sl), ctypes.c_void_p(device_ptr), 1, _p(stats))
            out = None
        else:
            out = np.zeros(shape, order="F")
            rc = fn(self.handle, _p(sl), _p(out), 0, _p(stats))
        if rc < 0:
            raise B200Error("block failed (%d): %s" % (rc, self.lib.cintb200_last_error().decode()))
        return out, stats

    def int2e_block(self, shls_slice, device_ptr=None, cart=False):
        """Dense (NI,NJ,NK,NL) tensor of int2e_sph (cart: int2e_cart) over shell slices (i0,i1,j0,j1,k0,k1,l0,l1); returns (array, stats)."""
        return self._block(self.lib.cintb200_int2e_cart_block if cart else self.lib.cintb200_int2e_sph_block, 4, shls_slice, device_ptr, cart)

    def int3c2e_block(self, shls_slice, device_ptr=None, cart=False):
        """Dense (NI,NJ,NK) tensor of int3c2e_sph / int3c2e_cart over shell slices (i0,i1,j0,j1,k0,k1)."""
        return self._block(self.lib.cintb200_int3c2e_cart_block if cart else self.lib.cintb200_int3c2e_sph_block, 3, shls_slice, device_ptr, cart)

    def int2c2e_block(self, shls_slice, device_ptr=None, cart=False):
        """Dense (NI,NK) matrix of int2c2e_sph / int2c2e_cart over shell slices (i0,i1,k0,k1): the density-fitting metric."""
        return self._block(self.lib.cintb200_int2c2e_cart_block if cart else self.lib.cintb200_int2c2e_sph_block, 2, shls_slice, device_ptr, cart)

    def ip1_block(self, shls_slice, kind=SPH, device_ptr=None):
        """( nabla i j | k l ) (8 slice bounds) or ( nabla i j | k ) (6 bounds) over shell slices: array of shape (NI, NJ, NK[, NL], 3)."""
        sl = np.ascontiguousarray(shls_slice, dtype=np.int32).reshape(-1)
        nc = sl.size // 2
        fn = self.lib.cintb200_int2e_ip1_block if nc == 4 else self.lib.cintb200_int3c2e_ip1_block
        fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        fn.restype = ctypes.c_int
        cart = kind == CART
        ao = np.concatenate([[0], np.cumsum([((int(b[1]) + 1) * (int(b[1]) + 2) // 2 if cart else 2 * int(b[1]) + 1) * int(b[3]) for b in self.bas])])
        shape = tuple(int(ao[sl[2 * m + 1]] - ao[sl[2 * m]]) for m in range(nc)) + (3,)
        stats = np.zeros(16)
        if device_ptr is not None:
            rc = fn(self.handle, kind, _p(sl), ctypes.c_void_p(device_ptr), 1, _p(stats))
            out = None
        else:
            out = np.zeros(shape, order="F")
            rc = fn(self.handle, kind, _p(sl), _p(out), 0, _p(stats))
        if rc < 0:
            raise B200Error("ip1 block failed (%d): %s" % (rc, self.lib.cintb200_last_error().decode()))
        return out, stats

    def aux_offset(self, k):
        """This rank's column offset of auxiliary shell k in the tiles of int3c2e_all (-1: owned by another rank)."""
        f = self.lib.cintb200_debug_aux_offset
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        f.restype = ctypes.c_int
        col = ctypes.c_longlong()
        if f(self.handle, k, ctypes.byref(col)) != 0:
            raise B200Error(self.lib.cintb200_last_error().decode())
        return col.value

    # ---- verification helpers (tests only) ----
    def force_generic(self, on=True):
        self.lib.cintb200_debug_force_generic(self.handle, int(on))

    def chunk(self, k):
        """Tile of chunk k after all_unique(): (values[ld, ncols] F-order, geometry dict)."""
        geom = np.zeros(8, dtype=np.int64)
        rc = self.lib.cintb200_debug_chunk(self.handle, k, None, 0, _p(geom))
        if rc < 0:
            raise B200Error(self.lib.cintb200_last_error().decode())
        ld, ncols = int(geom[3]), int(geom[4])
        out = np.zeros(ld * ncols)
        rc = self.lib.cintb200_debug_chunk(self.handle, k, _p(out), out.size, _p(geom))
        if rc < 0:
            raise B200Error(self.lib.cintb200_last_error().decode())
        return out.reshape((ld, ncols), order="F"), dict(i0=int(geom[0]), i1=int(geom[1]), row0=int(geom[2]),
                                                         ld=ld, ncols=ncols, nchunks=int(geom[5]))

    def profile(self, **kw):
        """Run the whole job once with per-launch CUDA events; returns (stats, rows[n,12]) -- see driver.cu."""
        self.lib.cintb200_debug_profile(self.handle, 1)
        try:
            st = self.all_unique(**kw)
        finally:
            self.lib.cintb200_debug_profile(self.handle, 0)
        n = self.lib.cintb200_debug_profile_rows(self.handle, None, 0)
        rows = np.zeros((n, 12))
        self.lib.cintb200_debug_profile_rows(self.handle, _p(rows), n)
        return st, rows

    def set_schwarz_threshold(self, thr):
        self.lib.cintb200_set_schwarz_threshold.argtypes = [ctypes.c_void_p, ctypes.c_double]
        self.lib.cintb200_set_schwarz_threshold(self.handle, thr)

    def schwarz_bounds(self):
        """sqrt(max|(ij|ij)|) per shell pair i >= j (index i(i+1)/2 + j), evaluated on the device."""
        f = self.lib.cintb200_schwarz_bounds
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        f.restype = ctypes.c_int
        n = f(self.handle, None)
        if n < 0:
            raise B200Error(self.lib.cintb200_last_error().decode())
        q = np.zeros(n)
        f(self.handle, _p(q))
        return q

    def block(self, row, col, nrow, ncol):
        """nrow x ncol rectangle (global row offset, this rank's column offset) of the LAST chunk's tile."""
        f = self.lib.cintb200_debug_block
        f.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        f.restype = ctypes.c_int
        out = np.zeros((nrow, ncol), order="F")
        if f(self.handle, row, col, nrow, ncol, _p(out)) != 0:
            raise B200Error(self.lib.cintb200_last_error().decode())
        return out

    def launch_rows(self):
        """Launch list of the cached whole-job plan in execution order (see driver.cu)."""
        f = self.lib.cintb200_debug_launch_rows
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        f.restype = ctypes.c_int
        n = f(self.handle, None, 0)
        rows = np.zeros((max(n, 0), 12))
        if n > 0:
            f(self.handle, _p(rows), n)
        return rows

    def pair_offsets(self, i, j):
        """(global row offset, this rank's column offset or -1) of the block of shell pair (i, j), i >= j."""
        r = ctypes.c_longlong()
        c = ctypes.c_longlong()
        self.lib.cintb200_debug_pair_offsets(self.handle, i, j, ctypes.byref(r), ctypes.byref(c))
        return r.value, c.value


def _call_single(name, ncenter, shls, atm, bas, env, opt=None, dims=None, out=None, cart=False):
    """One libcint-style call through the drop-in symbol `name` (e.g. 'int2e_sph')."""
    lib = load_library()
    atm, bas, env = _as_basis(atm, bas, env)
    d = shell_dims(bas, shls, cart)
    if out is None:
        shape = tuple(dims) if dims is not None else tuple(d)
        out = np.zeros(shape, order="F")
    cshls = (ctypes.c_int * ncenter)(*[int(s) for s in shls])
    cdims = (ctypes.c_int * ncenter)(*[int(x) for x in dims]) if dims is not None else None
    rc = getattr(lib, name)(_p(out), cdims, cshls, _p(atm), len(atm), _p(bas), len(bas), _p(env),
                            opt.handle if isinstance(opt, Context) else opt, None)
    return out, rc


def int2e_sph(shls, atm, bas, env, opt=None, dims=None, out=None):
    return _call_single("int2e_sph", 4, shls, atm, bas, env, opt, dims, out)


def int2e_cart(shls, atm, bas, env, opt=None, dims=None, out=None):
    return _call_single("int2e_cart", 4, shls, atm, bas, env, opt, dims, out, cart=True)


def int3c2e_sph(shls, atm, bas, env, opt=None, dims=None, out=None):
    return _call_single("int3c2e_sph", 3, shls, atm, bas, env, opt, dims, out)


def int3c2e_cart(shls, atm, bas, env, opt=None, dims=None, out=None):
    return _call_single("int3c2e_cart", 3, shls, atm, bas, env, opt, dims, out, cart=True)


def int3c2e_sph_ssc(shls, atm, bas, env, opt=None):
    """Spherical i, j with a Cartesian auxiliary index (src/cint3c2e.c:729)."""
    lib = load_library()
    atm, bas, env = _as_basis(atm, bas, env)
    d = shell_dims(bas, shls[:2]) + shell_dims(bas, shls[2:], cart=True)
    out = np.zeros(tuple(d), order="F")
    cshls = (ctypes.c_int * 3)(*[int(s) for s in shls])
    lib.int3c2e_sph_ssc.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3
    lib.int3c2e_sph_ssc.restype = ctypes.c_int
    rc = lib.int3c2e_sph_ssc(_p(out), None, cshls, _p(atm), len(atm), _p(bas), len(bas), _p(env), opt.handle if isinstance(opt, Context) else opt, None)
    return out, rc


def int2c2e_sph(shls, atm, bas, env, opt=None, dims=None, out=None):
    return _call_single("int2c2e_sph", 2, shls, atm, bas, env, opt, dims, out)


def int2c2e_cart(shls, atm, bas, env, opt=None, dims=None, out=None):
    return _call_single("int2c2e_cart", 2, shls, atm, bas, env, opt, dims, out, cart=True)


def plan_summary(atm, bas, env, rank=0, nranks=1, chunk_bytes=0, aux_shell0=None):
    """Static sharding of the whole job (host only, no GPU): dict of counts for `rank` of `nranks`.
    aux_shell0 given -> the density-fitting job (ij|k) with auxiliary shells [aux_shell0, nbas)."""
    lib = load_library()
    atm, bas, env = _as_basis(atm, bas, env)
    out = np.zeros(16)
    lib.cintb200_plan_summary.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
    lib.cintb200_plan_summary_3c.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
    if aux_shell0 is not None:
        rc = lib.cintb200_plan_summary_3c(_p(atm), len(atm), _p(bas), len(bas), _p(env), aux_shell0, rank, nranks, chunk_bytes, _p(out))
    else:
        rc = lib.cintb200_plan_summary(_p(atm), len(atm), _p(bas), len(bas), _p(env), rank, nranks, chunk_bytes, _p(out))
    if rc != 0:
        raise B200Error(lib.cintb200_last_error().decode())
    keys = ("quartets", "integrals", "prim_quartets", "model_flops", "columns", "rows", "chunks", "launches", "tile_bytes")
    return dict(zip(keys, out[:9]))


def fp64_peak_tflops(device=-1, seconds=0.5):
    """DFMA-chain microbenchmark: the FP64 roofline denominator (TFLOP/s)."""
    lib = load_library()
    v = ctypes.c_double()
    if lib.cintb200_fp64_peak(device, seconds, ctypes.byref(v)) != 0:
        raise B200Error(lib.cintb200_last_error().decode())
    return v.value


def release_cached_memory(device=-1):
    """Return the tile buffers cached in the device memory pool to the driver (cintb200_release_cached_memory)."""
    lib = load_library()
    lib.cintb200_release_cached_memory.argtypes = [ctypes.c_int]
    return lib.cintb200_release_cached_memory(device)


def fp64_peak_theoretical_tflops(device=-1, sm_mhz=0.0):
    """SMs x 64 FP64 lanes x 2 x clock (sm_mhz <= 0: maximum SM clock of the device)."""
    lib = load_library()
    lib.cintb200_fp64_peak_theoretical.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    v = ctypes.c_double()
    if lib.cintb200_fp64_peak_theoretical(device, sm_mhz, ctypes.byref(v)) != 0:
        raise B200Error(lib.cintb200_last_error().decode())
    return v.value


def load_fixture(name):
    """atm, bas, env of a committed benchmark molecule (see tools/make_fixtures.py)."""
    d = np.load(os.path.join(DATA_DIR, name + ".npz"))
    return d["atm"].astype(np.int32), d["bas"].astype(np.int32), d["env"].astype(np.float64)
