"""ctypes binding of the CPU oracle (oracle/tm_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/tm_oracle.h.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.  PARITY UNPINNED (the
reference holds no golden vectors for this path; upstream QUDA is not vendored).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtm_oracle.so")

# UKQCD gamma matrices, mu = x,y,z,t  (reference lib/code_pieces/gammas_tm_base.h:21-32)
def gamma_ukqcd():
    g = np.zeros((4, 4, 4), dtype=np.complex128)
    i = 1j
    g[0][0][3] = i; g[0][1][2] = i; g[0][2][1] = -i; g[0][3][0] = -i
    g[1][0][3] = 1; g[1][1][2] = -1; g[1][2][1] = -1; g[1][3][0] = 1
    g[2][0][2] = i; g[2][1][3] = -i; g[2][2][0] = -i; g[2][3][1] = i
    g[3][0][0] = 1; g[3][1][1] = 1; g[3][2][2] = -1; g[3][3][3] = -1
    return g

# DeGrand-Rossi gamma matrices (chiral; gamma5 diagonal) -- used only for the change-of-basis check
def gamma_degrand_rossi():
    g = np.zeros((4, 4, 4), dtype=np.complex128)
    i = 1j
    g[0] = [[0, 0, 0, i], [0, 0, i, 0], [0, -i, 0, 0], [-i, 0, 0, 0]]
    g[1] = [[0, 0, 0, -1], [0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0]]
    g[2] = [[0, 0, i, 0], [0, 0, 0, -i], [-i, 0, 0, 0], [0, i, 0, 0]]
    g[3] = [[0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0]]
    return g


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("tm_oracle.c", "tm_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None

def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        dpp = C.POINTER(dp)
        L.orc_set_lattice.argtypes = [C.POINTER(C.c_int)]
        L.orc_full_lattice_index.argtypes = [C.c_int, C.c_int]
        L.orc_neighbor_index.argtypes = [C.c_int] * 6
        L.orc_set_gamma.argtypes = [dp]
        L.orc_get_gamma5.argtypes = [dp]
        L.orc_apply_t_boundary.argtypes = [dpp, C.c_int]
        L.orc_su3_reconstruct12.argtypes = [dp, C.c_double]
        L.orc_plaquette.argtypes = [dpp]; L.orc_plaquette.restype = C.c_double
        L.orc_dslash.argtypes = [dp, dpp, dp, C.c_int, C.c_int]
        L.orc_twist_gamma5.argtypes = [dp, dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_tm_dslash.argtypes = [dp, dpp, dp, C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_tm_matpc.argtypes = [dp, dpp, dp, C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_tm_mdagm.argtypes = [dp, dpp, dp, C.c_double, C.c_double, C.c_int]
        L.orc_tm_mat.argtypes = [dp, dpp, dp, C.c_double, C.c_double, C.c_int]
        L.orc_prepare.argtypes = [dp, dpp, dp, C.c_double, C.c_double, C.c_int]
        L.orc_reconstruct.argtypes = [dp, dpp, dp, C.c_double, C.c_double, C.c_int]
        L.orc_norm2.argtypes = [dp, C.c_long]; L.orc_norm2.restype = C.c_double
        L.orc_redot.argtypes = [dp, dp, C.c_long]; L.orc_redot.restype = C.c_double
        L.orc_cdot.argtypes = [dp, dp, C.c_long, dp]
        L.orc_cg_mdagm.argtypes = [dp, dpp, dp, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int,
                                   C.c_int, dp, dp]
        L.orc_cg_mdagm.restype = C.c_int
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_clover_compute.argtypes = [dp, dpp, C.c_double]
        L.orc_set_clover.argtypes = [dp]
        L.orc_site_A.argtypes = [dp, dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        _lib = L
    return _lib


def _dp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _gpp(gauge):
    """gauge: float64 array [4][V][3][3][2] in QDP even-odd order -> double*[4]"""
    assert gauge.dtype == np.float64 and gauge.flags["C_CONTIGUOUS"] and gauge.shape[0] == 4
    arr = (C.POINTER(C.c_double) * 4)()
    for mu in range(4):
        arr[mu] = gauge[mu].ctypes.data_as(C.POINTER(C.c_double))
    return arr


class Oracle:
    """Stateful wrapper: one lattice + one gamma basis at a time (the C side uses globals, like
    the reference's Z[]/V/Vh in qkxtm/QKXTM_util.cpp:28-44)."""

    def __init__(self, X, gamma=None):
        self.L = lib()
        self.X = tuple(int(x) for x in X)
        self.V = int(np.prod(self.X)); self.Vh = self.V // 2
        self.L.orc_set_lattice((C.c_int * 4)(*self.X))
        self.L.orc_set_clover(None)           # the C side keeps globals: a new Oracle starts without a clover term
        self.set_gamma(gamma_ukqcd() if gamma is None else gamma)

    def set_gamma(self, g):
        a = np.ascontiguousarray(np.stack([g.real, g.imag], axis=-1), dtype=np.float64)
        self.L.orc_set_gamma(_dp(a))

    def gamma5(self):
        a = np.zeros((4, 4, 2)); self.L.orc_get_gamma5(_dp(a)); return a[..., 0] + 1j * a[..., 1]

    def full_index(self, i, odd): return self.L.orc_full_lattice_index(i, odd)
    def neighbor_index(self, i, odd, dx4, dx3, dx2, dx1): return self.L.orc_neighbor_index(i, odd, dx4, dx3, dx2, dx1)

    def apply_t_boundary(self, gauge, sign=-1): self.L.orc_apply_t_boundary(_gpp(gauge), sign)
    def plaquette(self, gauge): return self.L.orc_plaquette(_gpp(gauge))

    def reconstruct12(self, mat18, u0=1.0):
        m = np.array(mat18, dtype=np.float64).reshape(18).copy(); self.L.orc_su3_reconstruct12(_dp(m), u0); return m

    def _par(self): return np.empty((self.Vh, 4, 3, 2), dtype=np.float64)
    def _full(self): return np.empty((2 * self.Vh, 4, 3, 2), dtype=np.float64)

    def dslash(self, gauge, psi, odd_bit, dagger=0):
        out = self._par(); self.L.orc_dslash(_dp(out), _gpp(gauge), _dp(psi), odd_bit, dagger); return out

    def twist(self, psi, kappa, mu, dagger=0, inverse=0):
        out = np.empty_like(psi); n = psi.size // 24
        self.L.orc_twist_gamma5(_dp(out), _dp(psi), dagger, kappa, mu, inverse, n); return out

    def tm_dslash(self, gauge, psi, kappa, mu, odd_bit, dagger=0):
        out = self._par(); self.L.orc_tm_dslash(_dp(out), _gpp(gauge), _dp(psi), kappa, mu, odd_bit, dagger); return out

    def matpc(self, gauge, psi, kappa, mu, matpc=0, dagger=0):
        out = self._par(); self.L.orc_tm_matpc(_dp(out), _gpp(gauge), _dp(psi), kappa, mu, matpc, dagger); return out

    def mdagm(self, gauge, psi, kappa, mu, matpc=0):
        out = self._par(); self.L.orc_tm_mdagm(_dp(out), _gpp(gauge), _dp(psi), kappa, mu, matpc); return out

    def mat(self, gauge, psi, kappa, mu, dagger=0):
        out = self._full(); self.L.orc_tm_mat(_dp(out), _gpp(gauge), _dp(psi), kappa, mu, dagger); return out

    def prepare(self, gauge, b, kappa, mu, matpc=0):
        out = self._par(); self.L.orc_prepare(_dp(out), _gpp(gauge), _dp(b), kappa, mu, matpc); return out

    def reconstruct(self, gauge, x_full, b, kappa, mu, matpc=0):
        self.L.orc_reconstruct(_dp(x_full), _gpp(gauge), _dp(b), kappa, mu, matpc); return x_full

    def norm2(self, x): return self.L.orc_norm2(_dp(x), x.size)
    def redot(self, x, y): return self.L.orc_redot(_dp(x), _dp(y), x.size)
    def cdot(self, x, y):
        o = np.zeros(2); self.L.orc_cdot(_dp(x), _dp(y), x.size, _dp(o)); return complex(o[0], o[1])

    def cg_mdagm(self, gauge, b, kappa, mu, matpc=0, tol=1e-7, maxiter=10000, pr_beta=0):
        x = self._par(); tr = C.c_double(0.0)
        hist = np.zeros(maxiter + 2)
        it = self.L.orc_cg_mdagm(_dp(x), _gpp(gauge), _dp(b), kappa, mu, matpc, tol, maxiter, pr_beta,
                                 C.byref(tr), _dp(hist))
        return x, it, tr.value, hist[: it + 1]

    def num_threads(self): return self.L.orc_num_threads()
    def set_num_threads(self, n): self.L.orc_set_num_threads(int(n))

    # -- twisted-clover: after set_clover every operator above uses A = C + i a g5 (set_clover(None) switches back)
    def clover_compute(self, gauge, coeff):
        clov = np.empty((2 * self.Vh, 12, 12, 2), dtype=np.float64)
        self.L.orc_clover_compute(_dp(clov), _gpp(gauge), float(coeff))
        return clov

    def set_clover(self, clov):
        self._clov_keep = clov          # the C side keeps the pointer
        self.L.orc_set_clover(_dp(clov) if clov is not None else None)

    def site_A(self, psi, kappa, mu, parity, dagger=0, inverse=0):
        out = self._par(); self.L.orc_site_A(_dp(out), _dp(psi), dagger, kappa, mu, inverse, parity); return out


# ---- eigensolver restatements ----------------------------------------------------------------
def cheb_coefficients(deg, amin, amax):
    """The scalar recurrence of QKXTM_Deflation::polynomialOperator (reference lib/qudaQKXTM_Deflation.cpp:1003-1057):
    returns [(d1, d2, d3)] for steps 1..deg, T_1 = d2 T_0 + d1 A T_0, T_i = d1 A T_{i-1} + d2 T_{i-1} + d3 T_{i-2}."""
    delta, theta = (amax - amin) / 2.0, (amax + amin) / 2.0
    sigma1 = -delta / theta
    out = [(sigma1 / delta, 1.0, 0.0)]
    sigma_old = sigma1
    for _ in range(2, deg + 1):
        sigma = 1.0 / (2.0 / sigma1 - sigma_old)
        d1 = 2.0 * sigma / delta
        out.append((d1, -d1 * theta, -sigma * sigma_old))
        sigma_old = sigma
    return out[:deg]


def poly_operator(apply_A, x, deg, amin, amax):
    """out = p(A) x with the reference's loop structure (copy; MdagM + axpby; then per degree MdagM, ax, cxpaypbz and
    two copies -- Deflation.cpp:1013-1057), A given as a callable on numpy arrays."""
    out = x.copy()
    if deg == 0:
        return out
    coef = cheb_coefficients(deg, amin, amax)
    d1, d2, _ = coef[0]
    out = d2 * x + d1 * apply_A(x)
    if deg == 1:
        return out
    tm1, tm2 = x.copy(), out.copy()
    for (d1, d2, d3) in coef[1:]:
        out = apply_A(tm2)
        tm1 = d3 * tm1
        out = tm1 + d2 * tm2 + d1 * out
        tm1, tm2 = tm2, out.copy()
    return out


def poly_scalar(lam, deg, amin, amax):
    """p(lambda): the same recurrence on a number (what the filter does to an eigenvalue)."""
    return poly_operator(lambda v: lam * v, np.ones(1), deg, amin, amax)[0]


def eigs_reference(apply_A, n_complex, nev, ncv, which="SR", poly=None, tol=0.0, v0=None):
    """ARPACK through scipy (the reference drives the same p?naupd / p?neupd by reverse communication,
    Deflation.cpp:1296-1372): eigenpairs of the hermitian operator A acting on complex vectors of length n_complex.
    poly = (deg, amin, amax) iterates p(A) and asks ARPACK for the opposite end, as the reference's isACC branch."""
    import scipy.sparse.linalg as sla
    def mv(v):
        return apply_A(np.ascontiguousarray(v, dtype=np.complex128))
    if poly is not None:
        deg, amin, amax = poly
        op = sla.LinearOperator((n_complex, n_complex), dtype=np.complex128,
                                matvec=lambda v: poly_operator(mv, np.asarray(v, dtype=np.complex128).ravel(), deg, amin, amax))
        w, U = sla.eigsh(op, k=nev, ncv=ncv, which={"SR": "LA", "LR": "SA"}[which], tol=tol, v0=v0)
    else:
        op = sla.LinearOperator((n_complex, n_complex), dtype=np.complex128, matvec=lambda v: mv(np.asarray(v).ravel()))
        w, U = sla.eigsh(op, k=nev, ncv=ncv, which={"SR": "SA", "LR": "LA"}[which], tol=tol, v0=v0)
    # eigenvalues of the actual operator by Rayleigh quotient, as Deflation.cpp:1426-1439
    lam = np.array([np.vdot(U[:, i], mv(U[:, i])).real for i in range(nev)])
    o = np.argsort(lam)
    return lam[o], U[:, o]


# ---- Gaussian (Wuppertal) smearing restatement ---------------------------------------------------
def gauss_smear_step(vec, gauge, X, alpha):
    """One step of lib/code_pieces/Gauss_core.h in numpy on the plug-in's device layouts:
    vec  [12][V] complex (component s*3+c, x lexicographic x + X(y + Y(z + Z t)))    (Gauss_core.h:197 READVECTOR)
    gauge [4][3][3][V] complex (dir, c1, c2)                                          (Gauss_core.h:76 READGAUGE)
    out = (psi + alpha sum_{mu=0,1,2} [U_mu(x) psi(x+mu) + U_mu(x-mu)^dag psi(x-mu)]) / (1 + 6 alpha)   (:80-215);
    the time direction does not hop; periodic in x, y, z (single rank)."""
    Xd, Yd, Zd, Td = X
    psi = vec.reshape(4, 3, Td, Zd, Yd, Xd)
    U = gauge.reshape(4, 3, 3, Td, Zd, Yd, Xd)
    acc = np.zeros_like(psi)
    for mu, ax in ((0, 5), (1, 4), (2, 3)):
        fwd = np.roll(psi, -1, axis=ax)                                  # psi(x + mu)
        acc += np.einsum("ab...,sb...->sa...", U[mu], fwd)               # apply_U_on_S      (core_def.h:498-513)
        Ub = np.roll(U[mu], 1, axis=ax)                                  # U_mu(x - mu)
        bwd = np.roll(psi, 1, axis=ax)
        acc += np.einsum("ba...,sb...->sa...", Ub.conj(), bwd)           # apply_U_DAG_on_S  (core_def.h:515-530)
    out = (psi + alpha * acc) / (1.0 + 6.0 * alpha)
    return out.reshape(12, -1)


def gauss_smear(vec, gauge, X, alpha, nsmear):
    """QKXTM_Vector::gaussianSmearing (lib/qudaQKXTM_Vector.cpp:386-421)."""
    cur = vec
    for _ in range(nsmear):
        cur = gauss_smear_step(cur, gauge, X, alpha)
    return cur.copy()


# ---- propagator container kernels and the meson two-point contraction (the step after the solves) ---------------------
def rotate_physical(prop, sign):
    """lib/code_pieces/rotateToPhysicalBase_core.h: twisted -> physical basis, P <- 1/2 (1 + i s g5) P (1 + i s g5) on the
    propagator device layout [4 mu][4 nu][3 c1][3 c2][V] complex (gamma5 = UKQCD spin swap)."""
    g5 = _gamma5_ukqcd()
    R = np.eye(4) + 1j * sign * g5
    P = prop.reshape(4, 4, 3, 3, -1)
    return (0.5 * np.einsum("am,mncdx,ng->agcdx", R, P, R)).reshape(prop.shape)


def _gamma5_ukqcd():
    g = gamma_ukqcd()
    return g[0] @ g[1] @ g[2] @ g[3]


def meson_gammas():
    """The ten channels of the reference in its order (lib/qudaQKXTM_interface.cpp:305-314): pseudoscalar, scalar, g5g1..g5g4,
    g1..g4, with the overall sign of its tables (lib/qudaQKXTM_kernels.cu:77-78): +1 for the first six, -1 for g_mu."""
    g = gamma_ukqcd(); g5 = _gamma5_ukqcd()
    G = [g5, np.eye(4, dtype=complex)] + [g5 @ g[m] for m in range(4)] + [g[m] for m in range(4)]
    return G, [1.0] * 6 + [-1.0] * 4


def contract_mesons_site(prop):
    """C_G(x) = s_G tr[ G S(x) G^dag g5 S(x)^dag g5 ] over spin and colour for the ten channels, prop [4][4][3][3][V] complex
    -> [10][V] complex.  (What lib/code_pieces/contractMesons_core.h:20-33 accumulates per site from the index tables.)"""
    G, sG = meson_gammas(); g5 = _gamma5_ukqcd()
    S = prop.reshape(4, 4, 3, 3, -1)
    out = np.empty((10, S.shape[-1]), dtype=np.complex128)
    for ip in range(10):
        A = g5 @ G[ip]                       # (g5 G)[d, a]
        B = G[ip].conj().T @ g5              # (G^dag g5)[b, g]
        out[ip] = sG[ip] * np.einsum("da,abijx,bg,dgijx->x", A, S, B, S.conj())
    return out


def contract_mesons_mom(prop1, prop2, X, moms, src):
    """QKXTM_Contraction::contractMesons, MOMENTUM_SPACE (lib/code_pieces/contractMesons_core.h:37-87): for every time slice
    sum_xvec exp(-2 pi i p.(x - src)/L) C(x) -> [T][nmoms][2][10] complex (single rank)."""
    Xd, Yd, Zd, Td = X
    c = np.stack([contract_mesons_site(prop1), contract_mesons_site(prop2)]).reshape(2, 10, Td, Zd, Yd, Xd)
    x = np.arange(Xd) - src[0]; y = np.arange(Yd) - src[1]; z = np.arange(Zd) - src[2]
    out = np.empty((Td, len(moms), 2, 10), dtype=np.complex128)
    for im, (px, py, pz) in enumerate(moms):
        ph = np.exp(-2j * np.pi * (pz * z[:, None, None] / Zd + py * y[None, :, None] / Yd + px * x[None, None, :] / Xd))
        out[:, im] = np.einsum("uptzyx,zyx->tup", c, ph)
    return out


def create_momenta(q_sq):
    """createMomenta (lib/qudaQKXTM_kernels.cu:98-116): all integer momenta with p^2 <= Q_sq, in the reference's order."""
    m = []
    for iq in range(q_sq + 1):
        for nx in range(iq, -iq - 1, -1):
            for ny in range(iq, -iq - 1, -1):
                for nz in range(iq, -iq - 1, -1):
                    if nx * nx + ny * ny + nz * nz == iq:
                        m.append((nx, ny, nz))
    return m


# ---- baryon two-point contraction (the other half of the two-point step, lib/qudaQKXTM_interface.cpp:1220) ----------------
def _eps3():
    e = np.zeros((3, 3, 3))
    e[0, 1, 2] = e[1, 2, 0] = e[2, 0, 1] = 1.0
    e[2, 1, 0] = e[0, 2, 1] = e[1, 0, 2] = -1.0
    return e


def baryon_channels():
    """The ten channels of lib/code_pieces/contractBaryons_core.h in the reference's order (lib/qudaQKXTM_interface.cpp:294-303:
    nucl_nucl, nucl_roper, roper_nucl, roper_roper, deltapp_deltamm_11/22/33, deltap_deltaz_11/22/33).  Every channel is
        C[g][g'] = sign * sum  Gs[a,b] conj(Gr)[a',b']  Xs[g,d] Xr[g',d']  eps eps'  sum_terms coef * P1[a,s1] P2[b,s2] P3[d,s3]
    with (s1,s2,s3) a permutation of the primed spin slots (a', b', d') carrying the matching primed colours.  The index / value
    tables of lib/qudaQKXTM_kernels.cu:79-88 are exactly the non-zero entries of Gs x conj(Gr) x Xs x Xr (checked numerically):
    nucleon J = eps (u^T C g5 d) u, "roper" J = eps (u^T C d) g5 u, Delta J_k = eps (u^T C g_k u) u.
    A term is (coef, (prop of line 1, 2, 3), slots) with props 'A' = the channel's own propagator (prop1 for iu = 0, prop2 for
    iu = 1) and 'B' = the other one; slots name which primed index each line ends on."""
    g = gamma_ukqcd(); g5 = _gamma5_ukqcd(); one = np.eye(4, dtype=complex)
    Cm = g[3] @ g[1]                                   # charge conjugation C = g4 g2
    nucl = [(+1.0, "ABA", ("a", "b", "d")), (-1.0, "ABA", ("d", "b", "a"))]            # contractBaryons_core.h:68-69
    dpp = [(+1.0, "AAA", ("b", "d", "a")), (-1.0, "AAA", ("d", "b", "a")), (+1.0, "AAA", ("d", "a", "b")),
           (-1.0, "AAA", ("a", "d", "b")), (-1.0, "AAA", ("b", "a", "d")), (+1.0, "AAA", ("a", "b", "d"))]   # :360-366
    t = 1.0 / 3.0
    dp = [(-4 * t, "ABA", ("d", "b", "a")), (+2 * t, "ABA", ("b", "d", "a")), (+2 * t, "AAB", ("d", "a", "b")),
          (-2 * t, "AAB", ("a", "d", "b")), (-2 * t, "ABA", ("a", "d", "b")), (-1 * t, "AAB", ("b", "a", "d")),
          (+1 * t, "AAB", ("a", "b", "d")), (+4 * t, "ABA", ("a", "b", "d"))]                                 # :445-453
    ch = [dict(sign=+1.0, Gs=Cm @ g5, Gr=Cm @ g5, Xs=one, Xr=one, terms=nucl),
          dict(sign=+1.0, Gs=Cm @ g5, Gr=Cm, Xs=one, Xr=g5, terms=nucl),
          dict(sign=+1.0, Gs=Cm, Gr=Cm @ g5, Xs=g5, Xr=one, terms=nucl),
          dict(sign=+1.0, Gs=Cm, Gr=Cm, Xs=g5, Xr=g5, terms=nucl)]
    for k in range(3):
        ch.append(dict(sign=+1.0, Gs=Cm @ g[k], Gr=Cm @ g[k], Xs=one, Xr=one, terms=dpp))
    for k in range(3):
        ch.append(dict(sign=+1.0, Gs=Cm @ g[k], Gr=Cm @ g[k], Xs=one, Xr=one, terms=dp))
    return ch


def contract_baryons_site(prop1, prop2):
    """-> [2 iu][10 ip][4 gamma][4 gamma'][V] complex; prop [4][4][3][3][V] complex (spin sink, spin source, colour sink,
    colour source)"""
    e = _eps3()
    P = [prop1.reshape(4, 4, 3, 3, -1), prop2.reshape(4, 4, 3, 3, -1)]
    V = P[0].shape[-1]
    out = np.zeros((2, 10, 4, 4, V), dtype=np.complex128)
    # einsum letters: lines carry (spin, colour) = (a,i), (b,j), (d,k) at the sink; primed slots (A,I), (B,J), (D,K) at the source
    slot = {"a": ("A", "I"), "b": ("B", "J"), "d": ("D", "K")}
    for ip, c in enumerate(baryon_channels()):
        for iu in range(2):
            acc = np.zeros((4, 4, V), dtype=np.complex128)             # [d][D]
            for coef, props, slots in c["terms"]:
                ops = []
                for line, (sp, co) in enumerate((("a", "i"), ("b", "j"), ("d", "k"))):
                    S, Cc = slot[slots[line]]
                    ops.append("%s%s%s%sx" % (sp, S, co, Cc))
                pr = [P[iu] if ch == "A" else P[1 - iu] for ch in props]
                acc += coef * np.einsum("ab,AB,ijk,IJK,%s,%s,%s->dDx" % tuple(ops), c["Gs"], c["Gr"].conj(), e, e, *pr, optimize=True)
            out[iu, ip] = c["sign"] * np.einsum("gd,GD,dDx->gGx", c["Xs"], c["Xr"], acc)
    return out


def contract_baryons_mom(prop1, prop2, X, moms, src):
    """QKXTM_Contraction::contractBaryons, MOMENTUM_SPACE -> [T][nmoms][2][10][4][4] complex (single rank)"""
    Xd, Yd, Zd, Td = X
    c = contract_baryons_site(prop1, prop2).reshape(2, 10, 4, 4, Td, Zd, Yd, Xd)
    x = np.arange(Xd) - src[0]; y = np.arange(Yd) - src[1]; z = np.arange(Zd) - src[2]
    out = np.empty((Td, len(moms), 2, 10, 4, 4), dtype=np.complex128)
    for im, (px, py, pz) in enumerate(moms):
        ph = np.exp(-2j * np.pi * (pz * z[:, None, None] / Zd + py * y[None, :, None] / Yd + px * x[None, None, :] / Xd))
        out[:, im] = np.einsum("upgGtzyx,zyx->tupgG", c, ph)
    return out


# ---- fixed-sink three-point function: projectors, sequential sources, ultra-local insertion ------------------------------------
PROTON, NEUTRON = 0, 1                        # WHICHPARTICLE (include/qudaQKXTM_utils.h:128)
G4, G5G123, G5G1, G5G2, G5G3 = range(5)       # WHICHPROJECTOR (include/qudaQKXTM_utils.h:129)


def _twist_rotate(M, s):
    """physical -> twisted basis at maximal twist: 1/2 (1 + i s g5) M (1 + i s g5)"""
    R = np.eye(4) + 1j * s * _gamma5_ukqcd()
    return 0.5 * R @ M @ R


def projector_tm(pid, particle):
    """lib/code_pieces/projectors_tm_base.h as a formula: the physical projector 1/4 (1 + g4) [x i g5 g_k, or summed over k]
    rotated to the twisted basis with s = +1 (proton) / -1 (neutron)"""
    g = gamma_ukqcd(); g5 = _gamma5_ukqcd()
    P0 = 0.25 * (np.eye(4) + g[3])
    Pk = [P0 @ (1j * g5 @ g[k]) for k in range(3)]
    phys = {G4: P0, G5G123: Pk[0] + Pk[1] + Pk[2], G5G1: Pk[0], G5G2: Pk[1], G5G3: Pk[2]}[pid]
    return _twist_rotate(phys, +1 if particle == PROTON else -1)


def operator_tm(flag, particle, partflag):
    """lib/code_pieces/gammas_tm_base.h as a formula: the 16 insertions 1, g1..g4, g5, g5g1..g5g4, -i sigma_{12,13,23,41,42,43}
    rotated to the twisted basis; s = +1 for (proton, part 1) and (neutron, part 2), -1 otherwise"""
    g = gamma_ukqcd(); g5 = _gamma5_ukqcd()
    sig = lambda a, b: 0.5 * (g[a] @ g[b] - g[b] @ g[a])
    ops = [np.eye(4, dtype=complex)] + [g[k] for k in range(4)] + [g5] + [g5 @ g[k] for k in range(4)] + \
          [-1j * sig(a, b) for a, b in ((0, 1), (0, 2), (1, 2), (3, 0), (3, 1), (3, 2))]
    s = +1 if (particle == PROTON) == (partflag == 1) else -1
    return _twist_rotate(ops[flag], s)


def seq_source_part1(T1, T2, nu_f, c2_f, pid, particle):
    """lib/code_pieces/seqSourceFixSinkPart1_core.h: sequential source at the sink time slice for the quark line that occurs twice
    in the nucleon; T1, T2 3-d propagators [4][4][3][3][V3] complex -> [4 spin][3 colour][V3].  (G = C g5, P the projector)"""
    g = gamma_ukqcd(); Gm = (g[3] @ g[1]) @ _gamma5_ukqcd()
    P = projector_tm(pid, particle); e = _eps3(); ef = e[:, :, c2_f]
    # common factor  -eps_{c1 c2 c3} eps_{c1' c2' c2f} G[m,g] G[j,k] P[b,a] T2[g,j]^{c1 c1'}  times the four T1 placements
    A = -np.einsum("mg,jk,gjuUx->mkuUx", Gm, Gm, T2, optimize=True)          # [m][k][c1][c1'][x]
    out = np.zeros((4, 3, T1.shape[-1]), dtype=np.complex128)
    for b in range(4):
        for a in range(4):
            if abs(P[b, a]) < 1e-3:
                continue
            # (mu == nu, b == nu_f): T1[a][k]
            if b == nu_f:
                out += P[b, a] * np.einsum("uvw,UV,nkuUx,kvVx->nwx", e, ef, A, T1[a], optimize=True)
                # (a == nu, b == nu_f): T1[m][k], output spin a
                out[a] += P[b, a] * np.einsum("uvw,UV,mkuUx,mkvVx->wx", e, ef, A, T1, optimize=True)
            # (mu == nu, k == nu_f): T1[a][b]
            out += P[b, a] * np.einsum("uvw,UV,nuUx,vVx->nwx", e, ef, A[:, nu_f], T1[a, b], optimize=True)
            # (a == nu, k == nu_f): T1[m][b], output spin a
            out[a] += P[b, a] * np.einsum("uvw,UV,muUx,mvVx->wx", e, ef, A[:, nu_f], T1[:, b], optimize=True)
    return out


def seq_source_part2(T, nu_f, c2_f, pid, particle):
    """lib/code_pieces/seqSourceFixSinkPart2_core.h: sequential source for the quark line that occurs once"""
    g = gamma_ukqcd(); Gm = (g[3] @ g[1]) @ _gamma5_ukqcd()
    P = projector_tm(pid, particle); e = _eps3(); ef = e[:, :, c2_f]
    # S[n][c3] = -eps eps' G[m,n] G[nu_f,k] P[b,a] ( T[m,b]^{c1c1'} T[a,k]^{c2c2'} + T[m,k]^{c1c1'} T[a,b]^{c2c2'} )
    t1 = np.einsum("uvw,UV,mn,k,ba,mbuUx,akvVx->nwx", e, ef, Gm, Gm[nu_f], P, T, T, optimize=True)
    t2 = np.einsum("uvw,UV,mn,k,ba,mkuUx,abvVx->nwx", e, ef, Gm, Gm[nu_f], P, T, T, optimize=True)
    return -(t1 + t2)


def fixsink_local_site(fwd, seq, particle, partflag):
    """lib/code_pieces/fixSinkContractions_local_core.h:38-48: C_iop(x) = sum Gamma_iop[n][r] F[r][m'](b,a') S[n][m'](b,a')
    -> [16][V] complex; fwd / seq propagators [4][4][3][3][V] complex"""
    return np.stack([np.einsum("nr,rmbax,nmbax->x", operator_tm(i, particle, partflag), fwd, seq, optimize=True) for i in range(16)])


def fixsink_local_mom(fwd, seq, X, moms, src, particle, partflag):
    """... projected with exp(+2 pi i p.(x - src)/L) (fixSinkContractions_local_core.h:52-56: expon = cos + i sin)
    -> [T][nmoms][16] complex (single rank)"""
    Xd, Yd, Zd, Td = X
    c = fixsink_local_site(fwd.reshape(4, 4, 3, 3, -1), seq.reshape(4, 4, 3, 3, -1), particle, partflag).reshape(16, Td, Zd, Yd, Xd)
    x = np.arange(Xd) - src[0]; y = np.arange(Yd) - src[1]; z = np.arange(Zd) - src[2]
    out = np.empty((Td, len(moms), 16), dtype=np.complex128)
    for im, (px, py, pz) in enumerate(moms):
        ph = np.exp(+2j * np.pi * (pz * z[:, None, None] / Zd + py * y[None, :, None] / Yd + px * x[None, None, :] / Xd))
        out[:, im] = np.einsum("otzyx,zyx->to", c, ph)
    return out


def fixsink_derivative_site(fwd, seq, gauge, X, particle, partflag):
    """Conserved-current (Noether) and one-derivative insertions of the fixed-sink three-point function
    (lib/code_pieces/fixSinkContractions_noether_core.h:117-147, fixSinkContractions_oneD_core.h:100-134), single rank, periodic.
    With the four hop blocks per direction d (spin matrices [k][l], colours and the source spin summed)
        Af = S(x) U_d(x) F(x+d),   Ab = S(x) U_d(x-d)^dag F(x-d),   Bf = S(x+d) U_d(x)^dag F(x),   Bb = S(x-d) U_d(x-d) F(x)
    noether[d] = 1/4 { tr[(1+g_d)^T (Ab + Bf)] - tr[(1-g_d)^T (Af + Bb)] },   oneD[iop][d] = 1/4 tr[Gamma_iop^T (Af - Ab - Bf + Bb)]
    (the 1/4 is applied when the block sums are stored, noether_core.h:162, oneD_core.h:167).
    fwd / seq [4][4][3][3][V] complex, gauge [4][3][3][V] complex -> noether [4][V], oneD [16][4][V]"""
    Xd, Yd, Zd, Td = X
    g = gamma_ukqcd()
    F = fwd.reshape(4, 4, 3, 3, Td, Zd, Yd, Xd); S = seq.reshape(4, 4, 3, 3, Td, Zd, Yd, Xd); U = gauge.reshape(4, 3, 3, Td, Zd, Yd, Xd)
    noether = np.empty((4,) + F.shape[4:], dtype=np.complex128)
    oneD = np.empty((16, 4) + F.shape[4:], dtype=np.complex128)
    ops = [operator_tm(i, particle, partflag) for i in range(16)]
    for d, ax in enumerate((-1, -2, -3, -4)):                # x, y, z, t axes (counted from the end: the same for every array)
        up = lambda A: np.roll(A, -1, axis=ax)               # A(x + d)
        dn = lambda A: np.roll(A, +1, axis=ax)               # A(x - d)
        Ud = U[d]; Udm = dn(Ud)                              # U_d(x - d)
        Af = np.einsum("kpab...,ac...,lpcb...->kl...", S, Ud, up(F), optimize=True)
        Ab = np.einsum("kpab...,ca...,lpcb...->kl...", S, Udm.conj(), dn(F), optimize=True)
        Bf = np.einsum("kpab...,ca...,lpcb...->kl...", up(S), Ud.conj(), F, optimize=True)
        Bb = np.einsum("kpab...,ac...,lpcb...->kl...", dn(S), Udm, F, optimize=True)
        onePg, oneMg = np.eye(4) + g[d], np.eye(4) - g[d]    # operators 16+d / 20+d of gammas_tm_base.h: not rotated
        noether[d] = 0.25 * (np.einsum("kl,kl...->...", onePg, Ab + Bf) - np.einsum("kl,kl...->...", oneMg, Af + Bb))
        D = 0.25 * (Af - Ab - Bf + Bb)
        for i in range(16):
            oneD[i, d] = np.einsum("kl,kl...->...", ops[i], D)
    return noether.reshape(4, -1), oneD.reshape(16, 4, -1)


def fixsink_derivative_mom(fwd, seq, gauge, X, moms, src, particle, partflag):
    """... projected with exp(+2 pi i p.(x - src)/L) -> noether [T][nmoms][4], oneD [T][nmoms][4 dir][16 iop] complex"""
    Xd, Yd, Zd, Td = X
    n, o = fixsink_derivative_site(fwd, seq, gauge, X, particle, partflag)
    n = n.reshape(4, Td, Zd, Yd, Xd); o = o.reshape(16, 4, Td, Zd, Yd, Xd)
    x = np.arange(Xd) - src[0]; y = np.arange(Yd) - src[1]; z = np.arange(Zd) - src[2]
    out_n = np.empty((Td, len(moms), 4), dtype=np.complex128); out_o = np.empty((Td, len(moms), 4, 16), dtype=np.complex128)
    for im, (px, py, pz) in enumerate(moms):
        ph = np.exp(+2j * np.pi * (pz * z[:, None, None] / Zd + py * y[None, :, None] / Yd + px * x[None, None, :] / Xd))
        out_n[:, im] = np.einsum("dtzyx,zyx->td", n, ph)
        out_o[:, im] = np.einsum("idtzyx,zyx->tdi", o, ph)
    return out_n, out_o
