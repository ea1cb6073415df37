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
        L.qref_calculate_plaq.argtypes = [dp]; L.qref_calculate_plaq.restype = C.c_double
        for f in ("qref_conjugate_vector_double", "qref_conjugate_propagator_double", "qref_gamma5_propagator_double"):
            getattr(L, f).argtypes = [dp]
        L.qref_rotate_physical_double.argtypes = [dp, C.c_int]
        L.qref_rotate_physical_float.argtypes = [fp, C.c_int]
        L.qref_set_momenta.argtypes = [ip, C.c_int]
        L.qref_contract_mesons_mom_float.argtypes = [fp, fp, fp, ip]
        L.qref_contract_mesons_mom_double.argtypes = [dp, dp, dp, ip]
        L.qref_contract_mesons_pos_float.argtypes = [fp, fp, fp]
        L.qref_seq_source_part1_double.argtypes = [dp, C.c_int, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.qref_seq_source_part2_double.argtypes = [dp, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.qref_fixsink_local_float.argtypes = [fp, fp, fp, C.c_int, C.c_int, ip]
        L.qref_fixsink_local_double.argtypes = [dp, dp, dp, C.c_int, C.c_int, ip]
        L.qref_fixsink_derivative_double.argtypes = [dp, dp, dp, dp, dp, C.c_int, C.c_int, ip]
        L.qref_get_projector.argtypes = [dp, C.c_int, C.c_int]
        L.qref_get_operator.argtypes = [dp, C.c_int, C.c_int, C.c_int]
        L.qref_contract_baryons_mom_float.argtypes = [fp, fp, fp, ip]
        L.qref_contract_baryons_mom_double.argtypes = [dp, dp, dp, ip]
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

    def plaquette(self, gauge):
        """QKXTM_Gauge::calculatePlaq on the QKXTM device layout [4][3][3][V][2]"""
        return self.L.qref_calculate_plaq(_dp(np.ascontiguousarray(gauge, dtype=np.float64)))

    # ---- propagator kernels / meson contraction; propagator device layout [4 mu][4 nu][3 c1][3 c2][V][re,im] ----------------
    def conjugate_vector(self, vec):
        v = vec.copy(); self.L.qref_conjugate_vector_double(_dp(v)); return v

    def conjugate_propagator(self, prop):
        p = prop.copy(); self.L.qref_conjugate_propagator_double(_dp(p)); return p

    def gamma5_propagator(self, prop):
        p = prop.copy(); self.L.qref_gamma5_propagator_double(_dp(p)); return p

    def rotate_physical(self, prop, sign):
        p = prop.copy()
        if p.dtype == np.float64:
            self.L.qref_rotate_physical_double(_dp(p), int(sign))
        else:
            self.L.qref_rotate_physical_float(_fp(p), int(sign))
        return p

    def contract_mesons_mom(self, prop1, prop2, moms, src):
        """contractMesons, MOMENTUM_SPACE: -> [T][nmoms][2][10][re,im] in the propagators' precision (float = what the
        reference launches; double = the double instantiation of the same kernel body)"""
        m = np.ascontiguousarray(np.asarray(moms, dtype=np.int32).reshape(-1, 3))
        self.L.qref_set_momenta(m.ctypes.data_as(C.POINTER(C.c_int)), len(m))
        out = np.zeros((self.X[3], len(m), 2, 10, 2), dtype=prop1.dtype)
        s = (C.c_int * 3)(*[int(v) for v in src])
        if prop1.dtype == np.float32:
            self.L.qref_contract_mesons_mom_float(_fp(out), _fp(prop1), _fp(prop2), s)
        else:
            self.L.qref_contract_mesons_mom_double(_dp(out), _dp(prop1), _dp(prop2), s)
        return out

    def contract_baryons_mom(self, prop1, prop2, moms, src):
        """contractBaryons, MOMENTUM_SPACE: -> [T][nmoms][2][10][4][4][re,im] (float = what the reference launches)"""
        m = np.ascontiguousarray(np.asarray(moms, dtype=np.int32).reshape(-1, 3))
        self.L.qref_set_momenta(m.ctypes.data_as(C.POINTER(C.c_int)), len(m))
        out = np.zeros((self.X[3], len(m), 2, 10, 4, 4, 2), dtype=prop1.dtype)
        s = (C.c_int * 3)(*[int(v) for v in src])
        if prop1.dtype == np.float32:
            self.L.qref_contract_baryons_mom_float(_fp(out), _fp(prop1), _fp(prop2), s)
        else:
            self.L.qref_contract_baryons_mom_double(_dp(out), _dp(prop1), _dp(prop2), s)
        return out

    # ---- fixed-sink three-point function ------------------------------------------------------------------------------------
    def projector(self, pid, particle):
        a = np.zeros(32); self.L.qref_get_projector(_dp(a), pid, particle); return (a[0::2] + 1j * a[1::2]).reshape(4, 4)

    def operator(self, flag, particle, partflag):
        a = np.zeros(32); self.L.qref_get_operator(_dp(a), flag, particle, partflag); return (a[0::2] + 1j * a[1::2]).reshape(4, 4)

    def seq_source(self, part, timeslice, p3d_1, p3d_2, nu, c2, pid, particle):
        """seqSourceFixSinkPart1 / Part2 (double): 3-d propagators [4][4][3][3][V3][2] -> the 4-d vector [12][V][2] (only the
        time slice is written)"""
        out = np.zeros((12, self.V, 2))
        if part == 1:
            self.L.qref_seq_source_part1_double(_dp(out), timeslice, _dp(p3d_1), _dp(p3d_2), nu, c2, pid, particle)
        else:
            self.L.qref_seq_source_part2_double(_dp(out), timeslice, _dp(p3d_1), nu, c2, pid, particle)
        return out

    def fixsink_local(self, fwd, seq, particle, partflag, moms, src):
        """ultra-local part of contractFixSink, MOMENTUM_SPACE -> [T][nmoms][16][re,im]"""
        m = np.ascontiguousarray(np.asarray(moms, dtype=np.int32).reshape(-1, 3))
        self.L.qref_set_momenta(m.ctypes.data_as(C.POINTER(C.c_int)), len(m))
        out = np.zeros((self.X[3], len(m), 16, 2), dtype=fwd.dtype)
        s = (C.c_int * 3)(*[int(v) for v in src])
        if fwd.dtype == np.float32:
            self.L.qref_fixsink_local_float(_fp(out), _fp(fwd), _fp(seq), particle, partflag, s)
        else:
            self.L.qref_fixsink_local_double(_dp(out), _dp(fwd), _dp(seq), particle, partflag, s)
        return out

    def fixsink_derivative(self, fwd, seq, gauge, particle, partflag, moms, src):
        """Noether and one-derivative parts of contractFixSink (double): -> ([T][nmoms][4][2], [T][nmoms][4 dir][16 iop][2])"""
        m = np.ascontiguousarray(np.asarray(moms, dtype=np.int32).reshape(-1, 3))
        self.L.qref_set_momenta(m.ctypes.data_as(C.POINTER(C.c_int)), len(m))
        n = np.zeros((self.X[3], len(m), 4, 2)); o = np.zeros((self.X[3], len(m), 4, 16, 2))
        self.L.qref_fixsink_derivative_double(_dp(n), _dp(o), _dp(fwd), _dp(seq), _dp(gauge), particle, partflag, (C.c_int * 3)(*[int(v) for v in src]))
        return n, o

    def contract_mesons_pos(self, prop1, prop2):
        """contractMesons, POSITION_SPACE (float): -> [T][V3][2][10][re,im]"""
        out = np.zeros((self.X[3], self.V // self.X[3], 2, 10, 2), dtype=np.float32)
        self.L.qref_contract_mesons_pos_float(_fp(out), _fp(prop1), _fp(prop2))
        return out


# ---- the reference's host utility file qkxtm/QKXTM_util.cpp compiled in place (oracle/_ref/libqkxtm_util_ref.so) ------
_SO_UTIL = os.path.join(_HERE, "_ref", "libqkxtm_util_ref.so")
_util = None


def util_available():
    return os.path.exists(_SO_UTIL)


def util_lib():
    global _util
    if _util is None:
        L = C.CDLL(_SO_UTIL)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        dpp = C.POINTER(dp)
        L.qutil_set_dims.argtypes = [ip]
        L.qutil_full_lattice_index.argtypes = [C.c_int, C.c_int]
        L.qutil_neighbor_index.argtypes = [C.c_int] * 6
        L.qutil_get_odd_bit.argtypes = [C.c_int]
        L.qutil_su3_reconstruct12.argtypes = [dp, C.c_int, C.c_int, C.c_int]
        L.qutil_apply_gauge_field_scaling.argtypes = [dpp, C.c_int]
        L.qutil_construct_gauge_field.argtypes = [dpp, C.c_int, C.c_uint, C.c_int]
        L.qutil_read_lime_gauge.argtypes = [dpp, C.c_char_p, ip, C.c_double, C.c_double]
        _util = L
    return _util


def _g4(g):
    arr = (C.POINTER(C.c_double) * 4)()
    for mu in range(4):
        arr[mu] = g[mu].ctypes.data_as(C.POINTER(C.c_double))
    return arr


class RefUtil:
    """fullLatticeIndex / neighborIndex / getOddBit / su3Reconstruct12 / applyGaugeFieldScaling / construct_gauge_field of
    qkxtm/QKXTM_util.cpp and readLimeGauge of include/QKXTM_read_conf.h, run from the reference's own source."""

    def __init__(self, X):
        self.L = util_lib()
        self.X = tuple(int(v) for v in X)
        self.V = int(np.prod(self.X)); self.Vh = self.V // 2
        self.L.qutil_set_dims((C.c_int * 4)(*self.X))

    def full_index(self, i, odd): return self.L.qutil_full_lattice_index(i, odd)
    def neighbor_index(self, i, odd, dx4, dx3, dx2, dx1): return self.L.qutil_neighbor_index(i, odd, dx4, dx3, dx2, dx1)
    def odd_bit(self, Y): return self.L.qutil_get_odd_bit(Y)

    def reconstruct12(self, mat18, direction, ga_idx, t_boundary):
        m = np.array(mat18, dtype=np.float64).reshape(18).copy()
        self.L.qutil_su3_reconstruct12(_dp(m), direction, ga_idx, t_boundary)
        return m

    def apply_gauge_field_scaling(self, gauge, t_boundary):
        g = np.ascontiguousarray(gauge, dtype=np.float64).copy()
        self.L.qutil_apply_gauge_field_scaling(_g4(g), t_boundary)
        return g

    def construct_gauge_field(self, kind=1, seed=137, t_boundary=-1):
        """kind 0 = unit, 1 = random SU(3) (libc rand() after srand(seed)), QDP even-odd order [4][V][3][3][2]"""
        g = np.zeros((4, self.V, 3, 3, 2), dtype=np.float64)
        self.L.qutil_construct_gauge_field(_g4(g), kind, seed, t_boundary)
        return g

    def read_lime_gauge(self, fname):
        g = np.zeros((4, self.V, 3, 3, 2), dtype=np.float64)
        X = (C.c_int * 4)(*self.X)
        self.L.qutil_read_lime_gauge(_g4(g), fname.encode(), X, 0.0, 0.0)
        return g, tuple(X)
