"""ctypes binding of oracle/_ref/libqkxtm_ref.so: the REFERENCE'S OWN kernel bodies (lib/code_pieces/*_core.h) compiled
for the CPU from /root/reference by oracle/Makefile (ref_shim/qkxtm_kernels_host.cpp).

TEST INFRASTRUCTURE ONLY.  Used to pin the oracle's restatements and to generate tests/golden/ fixtures; the GPU box has
the prebuilt library only (no /root/reference there)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libqkxtm_ref.so")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.qref_set_geometry.argtypes = [ip, C.c_double]
        L.qref_gauss_step_double.argtypes = [dp, dp, dp]
        L.qref_gauss_step_float.argtypes = [fp, fp, fp]
        L.qref_upload.argtypes = [dp, dp, dp]
        L.qref_download.argtypes = [dp, dp, dp]
        L.qref_scale_vector.argtypes = [dp, C.c_double]
        L.qref_apply_gamma5_double.argtypes = [dp]
        _lib = L
    return _lib


def _dp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Ref:
    """Fields are in the plug-in's DEVICE layouts: vector d[(s*3+c)*V + x][re,im] (lib/qudaQKXTM_Vector.cpp:72-81),
    gauge d[((dir*3+c1)*3+c2)*V + x][re,im] (lib/qudaQKXTM_Gauge.cpp:73-89), x lexicographic."""

    def __init__(self, X, alpha_gauss=0.0):
        self.L = lib()
        self.X = tuple(int(v) for v in X)
        self.V = int(np.prod(self.X))
        self.L.qref_set_geometry((C.c_int * 4)(*self.X), float(alpha_gauss))

    def gauss_step(self, vec, gauge):
        out = np.empty_like(vec)
        if vec.dtype == np.float64:
            self.L.qref_gauss_step_double(_dp(out), _dp(vec), _dp(gauge))
        else:
            self.L.qref_gauss_step_float(_fp(out), _fp(vec), _fp(gauge))
        return out

    def gauss_smear(self, vec, gauge, nsmear):
        """QKXTM_Vector::gaussianSmearing (lib/qudaQKXTM_Vector.cpp:386-421): nsmear ping-pong steps; nsmear = 0 copies
        (the final cudaMemcpy of the even-count branch)."""
        cur = vec
        for _ in range(nsmear):
            cur = self.gauss_step(cur, gauge)
        return cur.copy()

    def upload(self, vec, even=True, odd=True):
        e = np.zeros((12, self.V // 2, 2)) if even else None
        o = np.zeros((12, self.V // 2, 2)) if odd else None
        self.L.qref_upload(_dp(vec), _dp(e) if even else None, _dp(o) if odd else None)
        return e, o

    def download(self, even, odd):
        out = np.full((12, self.V, 2), np.nan)
        self.L.qref_download(_dp(out), _dp(even) if even is not None else None, _dp(odd) if odd is not None else None)
        return out

    def scale(self, vec, a):
        v = vec.copy(); self.L.qref_scale_vector(_dp(v), float(a)); return v

    def gamma5(self, vec):
        v = vec.copy(); self.L.qref_apply_gamma5_double(_dp(v)); return v
