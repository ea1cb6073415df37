"""ctypes binding of libtmq.so (include/tmq.h) -- the thin Python host layer used by the parity tests and
bench.py.  The product's reference-facing host layer is the C++ QKXTM shim in ../host/; this module only
mirrors the C ABI one-to-one (same names, same argument meaning) so that tests read like upstream's
dslash_test / invert_test drivers.

There is NO fallback: if libtmq.so is missing or no CUDA device is usable, the calls raise.
"""
import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TMQ_LIB_PATH: tuning variants built by `make VARIANT=...` (tools/ only; tests and bench use the product library)
LIB_PATH = os.environ.get("TMQ_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libtmq.so")

PREC_SINGLE, PREC_DOUBLE = 4, 8
OPT_PREFETCH, OPT_HALO_P2P, OPT_BOUNDARY_AT_PCT, OPT_SMEAR_BLOCK_T = 1, 2, 3, 4
OPT_PACK_ASYNC, OPT_CONTRACT_SLICES, OPT_HALO_TIMEOUT_MS, OPT_CG_LAG = 5, 6, 7, 8
OPT_PACK_ASYNC, OPT_CONTRACT_SLICES = 5, 6
PARITY, FULL = 1, 2
MATPC_EVEN_EVEN, MATPC_ODD_ODD, MATPC_EVEN_EVEN_ASYM, MATPC_ODD_ODD_ASYM = 0, 1, 2, 3

# every symbol include/tmq.h declares (checked against the header by tests/test_abi.py)
SYMBOLS = """tmq_last_error tmq_version tmq_device_count tmq_create tmq_destroy tmq_sync tmq_comm_unique_id
tmq_comm_init tmq_force_partition tmq_set_tile tmq_set_option tmq_halo_mode tmq_gauge_load tmq_gauge_free tmq_plaquette tmq_spinor_alloc
tmq_spinor_free tmq_spinor_bytes tmq_spinor_from_qkxtm tmq_spinor_to_qkxtm tmq_spinor_from_host tmq_spinor_to_host
tmq_spinor_even tmq_spinor_odd tmq_op_set tmq_dslash tmq_dslash_twist_xpay tmq_matpc tmq_mdagm tmq_mat_full
tmq_prepare tmq_reconstruct tmq_cg_mdagm tmq_cg_history tmq_zero tmq_copy tmq_ax tmq_axpy tmq_axpby tmq_xpay
tmq_caxpy tmq_cxpaypbz tmq_norm2 tmq_redot tmq_cdot tmq_axpy_norm tmq_xmy_norm tmq_axpy_zpbx tmq_gamma5
tmq_qkxtm_plaquette tmq_qkxtm_scale tmq_qkxtm_cast tmq_qkxtm_gamma5 tmq_qkxtm_absorb tmq_dev_malloc tmq_dev_free tmq_dev_memset
tmq_h2d tmq_d2h tmq_time_kernel tmq_launch_count tmq_poly_mdagm tmq_eigset_alloc tmq_eigset_free tmq_eigset_size
tmq_eigset_vector tmq_eigensolve tmq_deflate tmq_project tmq_qkxtm_gauss_smear tmq_timer_start tmq_timer_stop tmq_clover_load tmq_clover_free
tmq_qkxtm_conjugate tmq_qkxtm_gamma5_prop tmq_qkxtm_rotate_physical tmq_qkxtm_column_copy tmq_qkxtm_contract_mesons tmq_qkxtm_contract_baryons tmq_qkxtm_seq_source tmq_qkxtm_fixsink_local tmq_qkxtm_fixsink_derivative tmq_d2d tmq_barrier tmq_host_prefetch tmq_spinor_from_prefetch tmq_spinor_to_host_async tmq_host_wait tmq_host_alloc_pinned tmq_host_free_pinned tmq_host_register tmq_host_unregister tmq_cg_stats tmq_qkxtm_ghost_sites tmq_qkxtm_exchange_ghost tmq_guard_check tmq_allreduce_host tmq_host_link_probe""".split()


class TmqError(RuntimeError):
    pass


_lib = None


def load():
    """dlopen libtmq.so; raises (never falls back) when the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TmqError("libtmq.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.tmq_last_error.restype = C.c_char_p
    L.tmq_create.restype = vp; L.tmq_create.argtypes = [C.c_int, ip, ip, ip]
    L.tmq_destroy.argtypes = [vp]; L.tmq_sync.argtypes = [vp]; L.tmq_barrier.argtypes = [vp]
    L.tmq_comm_unique_id.argtypes = [C.c_char_p]
    L.tmq_comm_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
    L.tmq_force_partition.argtypes = [vp, ip]
    L.tmq_set_tile.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.tmq_set_option.argtypes = [vp, C.c_int, C.c_int]
    L.tmq_halo_mode.argtypes = [vp]
    L.tmq_gauge_load.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int]
    L.tmq_gauge_free.argtypes = [vp]
    L.tmq_plaquette.argtypes = [vp, dp]
    L.tmq_spinor_alloc.restype = vp; L.tmq_spinor_alloc.argtypes = [vp, C.c_int, C.c_int]
    L.tmq_spinor_free.argtypes = [vp]
    L.tmq_spinor_bytes.restype = C.c_size_t; L.tmq_spinor_bytes.argtypes = [vp]
    L.tmq_spinor_from_qkxtm.argtypes = [vp, vp, C.c_int, C.c_int]
    L.tmq_spinor_to_qkxtm.argtypes = [vp, C.c_int, vp, C.c_int, C.c_double]
    L.tmq_spinor_from_host.argtypes = [vp, dp]; L.tmq_spinor_to_host.argtypes = [dp, vp]
    L.tmq_spinor_even.restype = vp; L.tmq_spinor_even.argtypes = [vp]
    L.tmq_spinor_odd.restype = vp; L.tmq_spinor_odd.argtypes = [vp]
    L.tmq_op_set.argtypes = [vp, C.c_double, C.c_double, C.c_int]
    L.tmq_dslash.argtypes = [vp, vp, C.c_int, C.c_int]
    L.tmq_dslash_twist_xpay.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_double]
    L.tmq_matpc.argtypes = [vp, vp, C.c_int]; L.tmq_mdagm.argtypes = [vp, vp]
    L.tmq_mat_full.argtypes = [vp, vp, C.c_int]
    L.tmq_prepare.argtypes = [vp, vp]; L.tmq_reconstruct.argtypes = [vp, vp, vp]
    L.tmq_cg_mdagm.argtypes = [vp, vp, C.c_double, C.c_int, C.c_double, C.c_int, ip, dp, dp, dp]
    L.tmq_cg_stats.argtypes = [vp, dp, ip]
    L.tmq_cg_history.argtypes = [vp, dp, C.c_int]
    L.tmq_zero.argtypes = [vp]; L.tmq_copy.argtypes = [vp, vp]
    L.tmq_ax.argtypes = [C.c_double, vp]
    L.tmq_axpy.argtypes = [C.c_double, vp, vp]
    L.tmq_axpby.argtypes = [C.c_double, vp, C.c_double, vp]
    L.tmq_xpay.argtypes = [vp, C.c_double, vp]
    L.tmq_caxpy.argtypes = [dp, vp, vp]
    L.tmq_cxpaypbz.argtypes = [vp, dp, vp, dp, vp]
    L.tmq_norm2.argtypes = [vp, dp]; L.tmq_redot.argtypes = [vp, vp, dp]; L.tmq_cdot.argtypes = [vp, vp, dp]
    L.tmq_axpy_norm.argtypes = [C.c_double, vp, vp, dp]; L.tmq_xmy_norm.argtypes = [vp, vp, dp]
    L.tmq_axpy_zpbx.argtypes = [C.c_double, vp, vp, vp, C.c_double]
    L.tmq_gamma5.argtypes = [vp]
    L.tmq_qkxtm_plaquette.argtypes = [vp, vp, C.c_int, dp]
    L.tmq_qkxtm_ghost_sites.argtypes = [vp]; L.tmq_qkxtm_ghost_sites.restype = C.c_size_t
    L.tmq_qkxtm_exchange_ghost.argtypes = [vp, vp, C.c_int, C.c_int]
    L.tmq_qkxtm_scale.argtypes = [vp, vp, C.c_int, C.c_double]
    L.tmq_qkxtm_cast.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.tmq_qkxtm_gamma5.argtypes = [vp, vp, C.c_int]
    L.tmq_qkxtm_absorb.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int]
    L.tmq_dev_malloc.argtypes = [vp, C.POINTER(vp), C.c_size_t]; L.tmq_dev_free.argtypes = [vp, vp]
    L.tmq_dev_memset.argtypes = [vp, vp, C.c_int, C.c_size_t]
    L.tmq_h2d.argtypes = [vp, vp, vp, C.c_size_t]; L.tmq_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.tmq_time_kernel.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, dp, C.POINTER(C.c_longlong)]
    L.tmq_launch_count.restype = C.c_longlong; L.tmq_launch_count.argtypes = [vp]
    L.tmq_poly_mdagm.argtypes = [vp, vp, C.c_int, C.c_double, C.c_double]
    L.tmq_eigset_alloc.restype = vp; L.tmq_eigset_alloc.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.tmq_eigset_free.argtypes = [vp]; L.tmq_eigset_size.argtypes = [vp]
    L.tmq_eigset_vector.restype = vp; L.tmq_eigset_vector.argtypes = [vp, C.c_int]
    L.tmq_eigensolve.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                 C.c_ulonglong, dp, dp, ip, ip, ip]
    L.tmq_deflate.argtypes = [vp, vp, vp, dp, C.c_int]
    L.tmq_project.argtypes = [vp, vp, vp, C.c_int]
    L.tmq_qkxtm_gauss_smear.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_double]
    L.tmq_qkxtm_conjugate.argtypes = [vp, vp, C.c_int, C.c_int]
    L.tmq_qkxtm_gamma5_prop.argtypes = [vp, vp, C.c_int]
    L.tmq_qkxtm_rotate_physical.argtypes = [vp, vp, C.c_int, C.c_int]
    L.tmq_qkxtm_column_copy.argtypes = [vp, vp, C.c_longlong, C.c_longlong, vp, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_int,
                                        C.c_int, C.c_int]
    L.tmq_qkxtm_contract_mesons.argtypes = [vp, vp, vp, C.c_int, ip, C.c_int, ip, dp, dp]
    L.tmq_qkxtm_contract_baryons.argtypes = [vp, vp, vp, C.c_int, ip, C.c_int, ip, dp]
    L.tmq_qkxtm_seq_source.argtypes = [vp, vp, C.c_int, vp, vp] + [C.c_int] * 6
    L.tmq_qkxtm_fixsink_local.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, ip, C.c_int, ip, dp]
    L.tmq_qkxtm_fixsink_derivative.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, ip, C.c_int, ip, dp, dp]
    L.tmq_timer_start.argtypes = [vp]; L.tmq_timer_stop.argtypes = [vp, dp]
    L.tmq_clover_load.argtypes = [vp, C.c_double]; L.tmq_clover_free.argtypes = [vp]
    _lib = L
    return L


def _ck(rc):
    if rc != 0:
        raise TmqError(load().tmq_last_error().decode())


def _i4(v):
    return (C.c_int * 4)(*[int(x) for x in v])


def _dp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def comm_unique_id():
    buf = C.create_string_buffer(128)
    _ck(load().tmq_comm_unique_id(buf))
    return buf.raw


class Spinor:
    def __init__(self, ctx, prec=PREC_DOUBLE, subset=PARITY, _handle=None):
        self.ctx, self.prec, self.subset = ctx, prec, subset
        self._owner = _handle is None
        self.h = _handle if _handle is not None else ctx.L.tmq_spinor_alloc(ctx.h, prec, subset)
        if not self.h:
            raise TmqError(ctx.L.tmq_last_error().decode())
        ctx._spinors.add(self)

    def free(self):
        if self.h and self._owner and self.ctx.h:
            self.ctx.L.tmq_spinor_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def nsites(self):
        return self.ctx.Vh * self.subset

    def set(self, host_eo):
        """host even-odd order [cb][4][3][2] float64 (FULL: [even Vh | odd Vh])"""
        a = np.ascontiguousarray(host_eo, dtype=np.float64)
        assert a.size == self.nsites * 24, (a.shape, self.nsites)
        _ck(self.ctx.L.tmq_spinor_from_host(self.h, _dp(a)))
        return self

    def get(self):
        out = np.empty((self.nsites, 4, 3, 2), dtype=np.float64)
        _ck(self.ctx.L.tmq_spinor_to_host(_dp(out), self.h))
        return out

    def even(self):
        return Spinor(self.ctx, self.prec, PARITY, _handle=self.ctx.L.tmq_spinor_even(self.h))

    def odd(self):
        return Spinor(self.ctx, self.prec, PARITY, _handle=self.ctx.L.tmq_spinor_odd(self.h))


class EigSet:
    """A set of parity vectors resident in HBM (tmq_eigset): Krylov basis / eigenvectors of QKXTM_Deflation."""

    def __init__(self, ctx, nvec, prec=PREC_DOUBLE, subset=PARITY):
        self.ctx, self.prec, self.nvec, self.subset = ctx, prec, nvec, subset
        self.h = ctx.L.tmq_eigset_alloc(ctx.h, nvec, prec, subset)
        if not self.h:
            raise TmqError(ctx.L.tmq_last_error().decode())
        ctx._eigsets.add(self)

    def vector(self, i):
        h = self.ctx.L.tmq_eigset_vector(self.h, i)
        if not h:
            raise TmqError(self.ctx.L.tmq_last_error().decode())
        return Spinor(self.ctx, self.prec, self.subset, _handle=h)

    def free(self):
        if self.h and self.ctx.h:
            self.ctx.L.tmq_eigset_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU / one rank.  Mirrors initQuda + init_qudaQKXTM + loadGaugeQuda + createDirac."""

    def __init__(self, localX, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0), device=0):
        self.L = load()
        self.X = tuple(int(x) for x in localX)
        self.grid, self.coord = tuple(grid), tuple(coord)
        self.V = int(np.prod(self.X)); self.Vh = self.V // 2
        self._spinors = weakref.WeakSet()
        self._eigsets = weakref.WeakSet()
        self.h = self.L.tmq_create(device, _i4(localX), _i4(grid), _i4(coord))
        if not self.h:
            raise TmqError(self.L.tmq_last_error().decode())
        self._gauge_keep = None

    def close(self):
        """tmq_destroy frees every spinor of the context: invalidate their Python handles first"""
        if self.h:
            for e in list(self._eigsets):
                e.free()
            for s in list(self._spinors):
                s.free()
            self.L.tmq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup
    def force_partition(self, part): _ck(self.L.tmq_force_partition(self.h, _i4(part)))
    def set_tile(self, ty, tz, tt): _ck(self.L.tmq_set_tile(self.h, ty, tz, tt))
    def set_option(self, opt, val): _ck(self.L.tmq_set_option(self.h, opt, val))
    def halo_mode(self): return self.L.tmq_halo_mode(self.h)
    def comm_init(self, uid, nranks, rank): _ck(self.L.tmq_comm_init(self.h, uid, nranks, rank))
    def sync(self): _ck(self.L.tmq_sync(self.h))
    def barrier(self): _ck(self.L.tmq_barrier(self.h))

    def load_gauge(self, gauge_qdp, t_boundary=-1, recon=12):
        """gauge_qdp: float64 [4][V][3][3][2], QDP even-odd order, boundary sign already folded in"""
        g = np.ascontiguousarray(gauge_qdp, dtype=np.float64)
        assert g.shape[0] == 4 and g[0].size == self.V * 18
        ptrs = (C.c_void_p * 4)(*[g[mu].ctypes.data for mu in range(4)])
        _ck(self.L.tmq_gauge_load(self.h, ptrs, int(t_boundary), int(recon)))

    def plaquette(self):
        p = C.c_double(0)
        _ck(self.L.tmq_plaquette(self.h, C.byref(p)))
        return p.value

    def set_op(self, kappa, mu, matpc=MATPC_EVEN_EVEN): _ck(self.L.tmq_op_set(self.h, kappa, mu, matpc))
    def clover_load(self, clover_coeff): _ck(self.L.tmq_clover_load(self.h, clover_coeff))     # loadCloverQuda(NULL, NULL, &inv_param)
    def clover_free(self): _ck(self.L.tmq_clover_free(self.h))

    def spinor(self, prec=PREC_DOUBLE, subset=PARITY): return Spinor(self, prec, subset)

    # -- operator
    def dslash(self, out, inp, out_parity, dagger=0): _ck(self.L.tmq_dslash(out.h, inp.h, out_parity, dagger))
    def dslash_twist_xpay(self, out, inp, out_parity, dagger=0, x=None, k=0.0):
        _ck(self.L.tmq_dslash_twist_xpay(out.h, inp.h, out_parity, dagger, x.h if x is not None else None, k))
    def matpc(self, out, inp, dagger=0): _ck(self.L.tmq_matpc(out.h, inp.h, dagger))
    def mdagm(self, out, inp): _ck(self.L.tmq_mdagm(out.h, inp.h))
    def mat_full(self, out, inp, dagger=0): _ck(self.L.tmq_mat_full(out.h, inp.h, dagger))
    def prepare(self, src_pc, b_full): _ck(self.L.tmq_prepare(src_pc.h, b_full.h))
    def reconstruct(self, x_full, x_pc, b_full): _ck(self.L.tmq_reconstruct(x_full.h, x_pc.h, b_full.h))

    def cg_mdagm(self, x, b, tol=1e-7, maxiter=10000, reliable_delta=1e-4, sloppy_prec=PREC_DOUBLE):
        it = C.c_int(0); tr = C.c_double(0); secs = C.c_double(0); gf = C.c_double(0)
        _ck(self.L.tmq_cg_mdagm(x.h, b.h, tol, maxiter, reliable_delta, sloppy_prec, C.byref(it), C.byref(tr),
                                C.byref(secs), C.byref(gf)))
        ls = C.c_double(0); ru = C.c_int(0)
        _ck(self.L.tmq_cg_stats(self.h, C.byref(ls), C.byref(ru)))
        return dict(iter=it.value, true_res=tr.value, secs=secs.value, gflops=gf.value, loop_secs=ls.value, reliable_updates=ru.value)

    def cg_history(self, n):
        h = np.zeros(n)
        _ck(self.L.tmq_cg_history(self.h, _dp(h), n))
        return h

    # -- eigensolver (QKXTM_Deflation)
    def poly_mdagm(self, out, inp, deg, amin, amax): _ck(self.L.tmq_poly_mdagm(out.h, inp.h, deg, amin, amax))

    def eigset(self, nvec, prec=PREC_DOUBLE, subset=PARITY): return EigSet(self, nvec, prec, subset)

    def eigensolve(self, eset, nev, nkv, poly_deg=0, amin=0.0, amax=0.0, tol=1e-10, max_restarts=100, which=0, seed=1):
        ev = np.zeros(nev); rs = np.zeros(nev)
        nc = C.c_int(0); nr = C.c_int(0); nm = C.c_int(0)
        _ck(self.L.tmq_eigensolve(eset.h, nev, nkv, poly_deg, amin, amax, tol, max_restarts, which, seed, _dp(ev), _dp(rs),
                                  C.byref(nc), C.byref(nr), C.byref(nm)))
        return dict(evals=ev, resid=rs, nconv=nc.value, restarts=nr.value, matvecs=nm.value)

    def deflate(self, out, inp, eset, evals, nvec=None):
        ev = np.ascontiguousarray(evals, dtype=np.float64)
        _ck(self.L.tmq_deflate(out.h, inp.h, eset.h, _dp(ev), len(ev) if nvec is None else nvec))

    def project(self, out, inp, eset, nvec): _ck(self.L.tmq_project(out.h, inp.h, eset.h, nvec))

    # -- blas
    def zero(self, x): _ck(self.L.tmq_zero(x.h))
    def copy(self, dst, src): _ck(self.L.tmq_copy(dst.h, src.h))
    def ax(self, a, x): _ck(self.L.tmq_ax(a, x.h))
    def axpy(self, a, x, y): _ck(self.L.tmq_axpy(a, x.h, y.h))
    def axpby(self, a, x, b, y): _ck(self.L.tmq_axpby(a, x.h, b, y.h))
    def xpay(self, x, a, y): _ck(self.L.tmq_xpay(x.h, a, y.h))
    def caxpy(self, a, x, y): _ck(self.L.tmq_caxpy((C.c_double * 2)(a.real, a.imag), x.h, y.h))
    def cxpaypbz(self, x, a, y, b, z):
        _ck(self.L.tmq_cxpaypbz(x.h, (C.c_double * 2)(a.real, a.imag), y.h, (C.c_double * 2)(b.real, b.imag), z.h))
    def norm2(self, x):
        o = C.c_double(0); _ck(self.L.tmq_norm2(x.h, C.byref(o))); return o.value
    def redot(self, x, y):
        o = C.c_double(0); _ck(self.L.tmq_redot(x.h, y.h, C.byref(o))); return o.value
    def cdot(self, x, y):
        o = (C.c_double * 2)(); _ck(self.L.tmq_cdot(x.h, y.h, o)); return complex(o[0], o[1])
    def axpy_norm(self, a, x, y):
        o = C.c_double(0); _ck(self.L.tmq_axpy_norm(a, x.h, y.h, C.byref(o))); return o.value
    def xmy_norm(self, x, y):
        o = C.c_double(0); _ck(self.L.tmq_xmy_norm(x.h, y.h, C.byref(o))); return o.value
    def axpy_zpbx(self, a, x, y, z, b): _ck(self.L.tmq_axpy_zpbx(a, x.h, y.h, z.h, b))
    def gamma5(self, x): _ck(self.L.tmq_gamma5(x.h))

    # -- raw device memory + QKXTM-layout kernels
    def dev_malloc(self, nbytes):
        p = C.c_void_p(0); _ck(self.L.tmq_dev_malloc(self.h, C.byref(p), nbytes)); return p.value
    def dev_free(self, p): _ck(self.L.tmq_dev_free(self.h, p))
    def h2d(self, dptr, arr):
        a = np.ascontiguousarray(arr); _ck(self.L.tmq_h2d(self.h, dptr, a.ctypes.data, a.nbytes))
    def d2h(self, arr, dptr):
        assert arr.flags["C_CONTIGUOUS"]; _ck(self.L.tmq_d2h(self.h, arr.ctypes.data, dptr, arr.nbytes))
    def from_qkxtm(self, dst, dptr, qprec=PREC_DOUBLE, parity=-1): _ck(self.L.tmq_spinor_from_qkxtm(dst.h, dptr, qprec, parity))
    def to_qkxtm(self, dptr, src, qprec=PREC_DOUBLE, parity=-1, scale=1.0):
        _ck(self.L.tmq_spinor_to_qkxtm(dptr, qprec, src.h, parity, scale))
    def qkxtm_plaquette(self, dgauge, prec):
        o = C.c_double(0); _ck(self.L.tmq_qkxtm_plaquette(self.h, dgauge, prec, C.byref(o))); return o.value
    def qkxtm_ghost_sites(self): return int(self.L.tmq_qkxtm_ghost_sites(self.h))
    def qkxtm_exchange_ghost(self, dptr, prec, ncomp): _ck(self.L.tmq_qkxtm_exchange_ghost(self.h, dptr, prec, ncomp))
    def qkxtm_scale(self, dptr, prec, a): _ck(self.L.tmq_qkxtm_scale(self.h, dptr, prec, a))
    def qkxtm_cast(self, dst, dprec, src, sprec): _ck(self.L.tmq_qkxtm_cast(self.h, dst, dprec, src, sprec))
    def qkxtm_gamma5(self, dptr, prec): _ck(self.L.tmq_qkxtm_gamma5(self.h, dptr, prec))
    def qkxtm_absorb(self, dprop, dvec, prec, nu, c2): _ck(self.L.tmq_qkxtm_absorb(self.h, dprop, dvec, prec, nu, c2))

    def qkxtm_conjugate(self, dptr, prec, ncomp): _ck(self.L.tmq_qkxtm_conjugate(self.h, dptr, prec, ncomp))
    def qkxtm_gamma5_prop(self, dprop, prec): _ck(self.L.tmq_qkxtm_gamma5_prop(self.h, dprop, prec))
    def qkxtm_rotate_physical(self, dprop, prec, sign): _ck(self.L.tmq_qkxtm_rotate_physical(self.h, dprop, prec, sign))
    def qkxtm_column_copy(self, dprop, prop_sites, prop_site0, dvec, vec_sites, vec_site0, nsites, prec, nu, c2, to_prop):
        _ck(self.L.tmq_qkxtm_column_copy(self.h, dprop, prop_sites, prop_site0, dvec, vec_sites, vec_site0, nsites, prec, nu, c2, int(to_prop)))

    def qkxtm_contract_mesons(self, dprop1, dprop2, prec, moms=None, src=(0, 0, 0), global_T=None, pos=False):
        """-> (corr_mom [T_global][nmoms][2][10] complex or None, corr_pos [V local][2][10] complex or None)"""
        cm = cp = None
        m = None
        if moms is not None:
            m = np.ascontiguousarray(np.asarray(moms, dtype=np.int32).reshape(-1, 3))
            cm = np.zeros((int(global_T), len(m), 2, 10, 2))
        if pos:
            cp = np.zeros((2 * self.Vh, 2, 10, 2))
        _ck(self.L.tmq_qkxtm_contract_mesons(self.h, dprop1, dprop2, prec, m.ctypes.data_as(C.POINTER(C.c_int)) if m is not None else None,
                                             len(m) if m is not None else 0, _i4(list(src) + [0]), _dp(cm) if cm is not None else None,
                                             _dp(cp) if cp is not None else None))
        c = lambda a: None if a is None else a[..., 0] + 1j * a[..., 1]
        return c(cm), c(cp)

    def qkxtm_contract_baryons(self, dprop1, dprop2, prec, moms, src, global_T):
        """-> corr [T_global][nmoms][2 iu][10 ip][4][4] complex"""
        m = np.ascontiguousarray(np.asarray(moms, dtype=np.int32).reshape(-1, 3))
        out = np.zeros((int(global_T), len(m), 2, 10, 4, 4, 2))
        _ck(self.L.tmq_qkxtm_contract_baryons(self.h, dprop1, dprop2, prec, m.ctypes.data_as(C.POINTER(C.c_int)), len(m),
                                              _i4(list(src) + [0]), _dp(out)))
        return out[..., 0] + 1j * out[..., 1]

    def qkxtm_seq_source(self, dvec, timeslice, dp3d_1, dp3d_2, prec, nu, c2, pid, particle, part):
        _ck(self.L.tmq_qkxtm_seq_source(self.h, dvec, timeslice, dp3d_1, dp3d_2, prec, nu, c2, pid, particle, part))

    def qkxtm_fixsink_local(self, dseq, dfwd, prec, particle, partflag, moms, src, global_T):
        """-> corr [T_global][nmoms][16] complex"""
        m = np.ascontiguousarray(np.asarray(moms, dtype=np.int32).reshape(-1, 3))
        out = np.zeros((int(global_T), len(m), 16, 2))
        _ck(self.L.tmq_qkxtm_fixsink_local(self.h, dseq, dfwd, prec, particle, partflag, m.ctypes.data_as(C.POINTER(C.c_int)), len(m),
                                           _i4(list(src) + [0]), _dp(out)))
        return out[..., 0] + 1j * out[..., 1]

    def qkxtm_fixsink_derivative(self, dseq, dfwd, dgauge, prec, particle, partflag, moms, src):
        """-> (noether [T][nmoms][4], oneD [T][nmoms][4 dir][16 iop]) complex"""
        m = np.ascontiguousarray(np.asarray(moms, dtype=np.int32).reshape(-1, 3))
        T = self.X[3]
        n = np.zeros((T, len(m), 4, 2)); o = np.zeros((T, len(m), 4, 16, 2))
        _ck(self.L.tmq_qkxtm_fixsink_derivative(self.h, dseq, dfwd, dgauge, prec, particle, partflag, m.ctypes.data_as(C.POINTER(C.c_int)), len(m),
                                                _i4(list(src) + [0]), _dp(n), _dp(o)))
        return n[..., 0] + 1j * n[..., 1], o[..., 0] + 1j * o[..., 1]

    def qkxtm_gauss_smear(self, dout, din, dgauge, prec, nsmear, alpha):
        _ck(self.L.tmq_qkxtm_gauss_smear(self.h, dout, din, dgauge, prec, nsmear, alpha))

    # -- measurement
    def timer_start(self): _ck(self.L.tmq_timer_start(self.h))
    def timer_stop(self):
        ms = C.c_double(0); _ck(self.L.tmq_timer_stop(self.h, C.byref(ms))); return ms.value

    def time_kernel(self, kind, prec, reps, inp, flush_l2=0):
        ms = C.c_double(0); nl = C.c_longlong(0)
        _ck(self.L.tmq_time_kernel(self.h, kind, prec, reps, inp.h, flush_l2, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def launch_count(self): return int(self.L.tmq_launch_count(self.h))


# ---- host-side field generators (libqkxtm_tmq.so, include/tmq_host.h) ------------------------------------------
HOST_LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libqkxtm_tmq.so")
_hostlib = None


def load_host():
    global _hostlib
    if _hostlib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise TmqError("libqkxtm_tmq.so not built (%s)" % HOST_LIB_PATH)
        load()   # libtmq.so first (the shim links against it)
        H = C.CDLL(HOST_LIB_PATH)
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        H.tmq_fieldgen_gauge_qdp.argtypes = [C.POINTER(dp), ip, ip, ip, C.c_ulonglong, C.c_int]
        H.tmq_fieldgen_unit_gauge_qdp.argtypes = [C.POINTER(dp), ip, ip, ip, C.c_int]
        H.tmq_fieldgen_spinor_gaussian.argtypes = [dp, ip, ip, ip, C.c_ulonglong, C.c_int]
        H.tmq_fieldgen_spinor_z4.argtypes = [dp, ip, ip, ip, C.c_ulonglong, C.c_int]
        H.tmq_lime_last_error.restype = C.c_char_p
        H.tmq_lime_gauge_info.argtypes = [C.c_char_p, ip, ip, dp, dp]
        H.tmq_lime_read_gauge.argtypes = [C.c_char_p, C.POINTER(dp), ip, ip, ip]
        H.tmq_lime_write_gauge.argtypes = [C.c_char_p, C.POINTER(dp), ip, ip, ip, C.c_double, C.c_double]
        H.tmq_lime_write_vector.argtypes = [C.c_char_p, C.c_void_p, C.c_int, ip, ip, ip]
        H.tmq_lime_read_vector.argtypes = [C.c_char_p, C.c_void_p, C.c_int, ip, ip, ip]
        H.tmq_apply_t_boundary.argtypes = [C.POINTER(dp), ip, ip, ip, C.c_int]
        _hostlib = H
    return _hostlib


def gen_gauge(localX, seed=137, t_boundary=-1, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0), unit=False):
    """float64 [4][V][3][3][2], QDP even-odd order, boundary sign folded in"""
    H = load_host()
    V = int(np.prod(localX))
    g = np.empty((4, V, 3, 3, 2), dtype=np.float64)
    ptrs = (C.POINTER(C.c_double) * 4)(*[g[mu].ctypes.data_as(C.POINTER(C.c_double)) for mu in range(4)])
    if unit:
        H.tmq_fieldgen_unit_gauge_qdp(ptrs, _i4(localX), _i4(grid), _i4(coord), t_boundary)
    else:
        H.tmq_fieldgen_gauge_qdp(ptrs, _i4(localX), _i4(grid), _i4(coord), seed, t_boundary)
    return g


def gen_spinor(localX, kind="gaussian", seed=None, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0), eo_order=True):
    H = load_host()
    V = int(np.prod(localX))
    out = np.empty((V, 4, 3, 2), dtype=np.float64)
    if kind == "gaussian":
        H.tmq_fieldgen_spinor_gaussian(_dp(out), _i4(localX), _i4(grid), _i4(coord), 101 if seed is None else seed, int(eo_order))
    elif kind == "z4":
        H.tmq_fieldgen_spinor_z4(_dp(out), _i4(localX), _i4(grid), _i4(coord), 100 if seed is None else seed, int(eo_order))
    else:
        raise ValueError(kind)
    return out


# ---- on-disk formats (host/tmq_lime.cpp) ---------------------------------------------------------------------------
def _hck(rc):
    if rc != 0:
        raise TmqError(load_host().tmq_lime_last_error().decode())


def _g4(g):
    return (C.POINTER(C.c_double) * 4)(*[g[mu].ctypes.data_as(C.POINTER(C.c_double)) for mu in range(4)])


def lime_gauge_info(fname):
    X = (C.c_int * 4)(); prec = C.c_int(0); k = C.c_double(0); m = C.c_double(0)
    _hck(load_host().tmq_lime_gauge_info(fname.encode(), X, C.byref(prec), C.byref(k), C.byref(m)))
    return dict(X=tuple(X), precision=prec.value, kappa=k.value, mu=m.value)


def lime_read_gauge(fname, localX, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0)):
    g = np.empty((4, int(np.prod(localX)), 3, 3, 2), dtype=np.float64)
    _hck(load_host().tmq_lime_read_gauge(fname.encode(), _g4(g), _i4(localX), _i4(grid), _i4(coord)))
    return g


def lime_write_gauge(fname, gauge_qdp, localX, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0), kappa=0.0, mu=0.0):
    g = np.ascontiguousarray(gauge_qdp, dtype=np.float64)
    _hck(load_host().tmq_lime_write_gauge(fname.encode(), _g4(g), _i4(localX), _i4(grid), _i4(coord), kappa, mu))


def lime_write_vector(fname, aos, localX, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0)):
    a = np.ascontiguousarray(aos)
    prec = 8 if a.dtype == np.float64 else 4
    _hck(load_host().tmq_lime_write_vector(fname.encode(), a.ctypes.data, prec, _i4(localX), _i4(grid), _i4(coord)))


def lime_write_vector_header(fname, localX, grid, prec=8):
    H = load_host()
    H.tmq_lime_write_vector_header.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _hck(H.tmq_lime_write_vector_header(fname.encode(), prec, _i4(localX), _i4(grid)))


def lime_write_vector_block(fname, aos, localX, grid, coord):
    H = load_host()
    H.tmq_lime_write_vector_block.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    a = np.ascontiguousarray(aos)
    _hck(H.tmq_lime_write_vector_block(fname.encode(), a.ctypes.data, 8 if a.dtype == np.float64 else 4, _i4(localX), _i4(grid), _i4(coord)))


def lime_read_vector(fname, localX, dtype=np.float64, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0)):
    out = np.empty((int(np.prod(localX)), 4, 3, 2), dtype=dtype)
    _hck(load_host().tmq_lime_read_vector(fname.encode(), out.ctypes.data, out.dtype.itemsize, _i4(localX), _i4(grid), _i4(coord)))
    return out


def apply_t_boundary(gauge_qdp, localX, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0), t_boundary=-1):
    load_host().tmq_apply_t_boundary(_g4(gauge_qdp), _i4(localX), _i4(grid), _i4(coord), t_boundary)


# ---- noise vectors of calc_loops (host/qkxtm_noise.cpp) -------------------------------------------------------------
def _noise_lib():
    H = load_host()
    if not getattr(H, "_noise_ready", False):
        H.tmq_ranlux_alloc.argtypes = [C.c_ulong]; H.tmq_ranlux_alloc.restype = C.c_void_p
        H.tmq_ranlux_free.argtypes = [C.c_void_p]
        H.tmq_ranlux_get.argtypes = [C.c_void_p]; H.tmq_ranlux_get.restype = C.c_ulong
        H.tmq_ranlux_uniform_int.argtypes = [C.c_void_p, C.c_ulong]; H.tmq_ranlux_uniform_int.restype = C.c_ulong
        H.tmq_noise_z4.argtypes = [C.POINTER(C.c_double), C.c_longlong, C.c_void_p, C.c_int]
        H.tmq_hch_coloring.argtypes = [C.POINTER(C.c_ushort), C.POINTER(C.c_int), C.c_int, C.c_int]
        H.tmq_hadamard_element.argtypes = [C.c_int, C.c_int]
        H._noise_ready = True
    return H


def guard_check():
    """number of libtmq device allocations with a corrupted red zone (TMQ_GUARD_BYTES); raises on a CUDA error"""
    L = load()
    n = int(L.tmq_guard_check())
    if n < 0:
        raise TmqError(L.tmq_last_error().decode())
    return n, (L.tmq_last_error().decode() if n else "")


class Ranlux:
    """gsl_rng_ranlux restated (see include/tmq_host.h)"""

    def __init__(self, seed):
        self.H = _noise_lib()
        self.h = self.H.tmq_ranlux_alloc(seed)

    def get(self): return int(self.H.tmq_ranlux_get(self.h))

    def uniform_int(self, n): return int(self.H.tmq_ranlux_uniform_int(self.h, n))

    def z4(self, ncomplex, unity=False):
        out = np.empty((ncomplex, 2), dtype=np.float64)
        self.H.tmq_noise_z4(_dp(out), ncomplex, self.h, int(unity))
        return out

    def __del__(self):
        try:
            self.H.tmq_ranlux_free(self.h)
        except Exception:
            pass


def hch_coloring(L, k, d=4):
    H = _noise_lib()
    n = int(np.prod(L[:d]))
    out = np.empty(n, dtype=np.uint16)
    rc = H.tmq_hch_coloring(out.ctypes.data_as(C.POINTER(C.c_ushort)), (C.c_int * d)(*L[:d]), k, d)
    if rc:
        raise TmqError("tmq_hch_coloring failed (%d)" % rc)
    return out


def hadamard_element(i, j):
    return int(_noise_lib().tmq_hadamard_element(i, j))
