"""Synthetic lattice inputs and index maps shared by the tests and bench.py.

Restates, in numpy, the in-tree half of the reference's test helpers:
  * random SU(3): rows 1,2 uniform in [0,1) -> normalise, Gram-Schmidt, normalise; row 0 = conj cross
    product (qkxtm/QKXTM_util.cpp:879-955) -- but drawn from a portable counter-based RNG keyed by
    (seed, direction, GLOBAL lexicographic site, component) so that any sharding sees the same
    global field (the reference's libc rand() order-dependence is deliberately not reproduced);
  * QDP even-odd gauge order [even Vh | odd Vh] (qkxtm/QKXTM_util.cpp:840-857);
  * anti-periodic T folded into U_t(T-1) (qkxtm/QKXTM_util.cpp:698-705);
  * Z4 noise source (lib/qudaQKXTM_utils.cpp:148-180);
  * lexicographic <-> even-odd spinor reorder (include/QKXTM_mapping_parity.h:17-110).
The same generator exists in C++ in the product's host layer (host/tmq_fieldgen.cpp); a test checks
that both produce identical bytes.
"""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = z.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        z += np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform(seed, stream, site, comp):
    """site: uint64 array of global lexicographic indices; returns float64 in [0,1)."""
    k = _mix64(np.array([seed], dtype=np.uint64))
    k = _mix64(k ^ np.uint64(stream))
    k = _mix64(k ^ site.astype(np.uint64))
    k = _mix64(k ^ np.uint64(comp))
    return (k >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def coords_lex(X):
    """returns x,y,z,t arrays for lexicographic index i = x + X0(y + X1(z + X2 t))"""
    V = int(np.prod(X))
    i = np.arange(V, dtype=np.int64)
    x = i % X[0]; y = (i // X[0]) % X[1]; z = (i // (X[0] * X[1])) % X[2]; t = i // (X[0] * X[1] * X[2])
    return x, y, z, t


def eo_from_lex(X):
    """perm such that field_eo[k] = field_lex[perm[k]]; even-odd order = [even Vh | odd Vh], each in
    lexicographic order (cb index = lex/2)."""
    x, y, z, t = coords_lex(X)
    par = (x + y + z + t) & 1
    lex = np.arange(len(x), dtype=np.int64)
    return np.concatenate([lex[par == 0], lex[par == 1]])


def global_lex(Xloc, grid, coord):
    """global lexicographic index of each local lexicographic site for a rank at `coord` of `grid`"""
    x, y, z, t = coords_lex(Xloc)
    G = [Xloc[d] * grid[d] for d in range(4)]
    gx = x + coord[0] * Xloc[0]; gy = y + coord[1] * Xloc[1]; gz = z + coord[2] * Xloc[2]; gt = t + coord[3] * Xloc[3]
    return (gx + G[0] * (gy + G[1] * (gz + G[2] * gt))).astype(np.uint64), (gx, gy, gz, gt)


def random_su3_lex(Xloc, seed=137, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0)):
    """[4][V][3][3] complex128, local lexicographic site order"""
    gl, _ = global_lex(Xloc, grid, coord)
    V = len(gl)
    U = np.empty((4, V, 3, 3), dtype=np.complex128)
    for mu in range(4):
        rows = np.empty((V, 2, 3), dtype=np.complex128)
        for m in range(2):
            for n in range(3):
                re = uniform(seed, mu, gl, (m * 3 + n) * 2)
                im = uniform(seed, mu, gl, (m * 3 + n) * 2 + 1)
                rows[:, m, n] = re + 1j * im
        u = rows[:, 0, :]; v = rows[:, 1, :]
        u = u / np.sqrt(np.sum(np.abs(u) ** 2, axis=1, keepdims=True))
        dot = np.sum(np.conj(u) * v, axis=1, keepdims=True)
        v = v - dot * u
        v = v / np.sqrt(np.sum(np.abs(v) ** 2, axis=1, keepdims=True))
        w = np.conj(np.cross(u, v))
        U[mu, :, 0, :] = w; U[mu, :, 1, :] = u; U[mu, :, 2, :] = v
    return U


def gauge_qdp_from_lex(U_lex, Xloc, t_boundary=-1, last_in_t=True):
    """complex [4][V][3][3] lexicographic -> float64 [4][V][3][3][2] QDP even-odd with the T boundary
    condition folded into U_t on the last local time slice of the last rank in T."""
    perm = eo_from_lex(Xloc)
    U = U_lex.copy()
    if t_boundary == -1 and last_in_t:
        x, y, z, t = coords_lex(Xloc)
        U[3, t == Xloc[3] - 1] *= -1.0
    Ueo = U[:, perm]
    return np.ascontiguousarray(np.stack([Ueo.real, Ueo.imag], axis=-1))


def random_gauge_qdp(Xloc, seed=137, t_boundary=-1, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0)):
    last = coord[3] == grid[3] - 1
    return gauge_qdp_from_lex(random_su3_lex(Xloc, seed, grid, coord), Xloc, t_boundary, last)


def unit_gauge_qdp(Xloc, t_boundary=+1):
    V = int(np.prod(Xloc))
    U = np.zeros((4, V, 3, 3), dtype=np.complex128)
    U[:, :, range(3), range(3)] = 1.0
    return gauge_qdp_from_lex(U, Xloc, t_boundary)


def gaussian_spinor_lex(Xloc, seed=101, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0)):
    """dense Gaussian spinor, host order [x_lex][s][c][re,im] (lib/qudaQKXTM_Vector.cpp:72-81)"""
    gl, _ = global_lex(Xloc, grid, coord)
    V = len(gl)
    out = np.empty((V, 24), dtype=np.float64)
    for k in range(24):
        u1 = uniform(seed, 16, gl, 2 * k); u2 = uniform(seed, 16, gl, 2 * k + 1)
        out[:, k] = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
    return out.reshape(V, 4, 3, 2)


def z4_source_lex(Xloc, seed=100, grid=(1, 1, 1, 1), coord=(0, 0, 0, 0)):
    """Z4 noise: 0 -> +1, 1 -> -1, 2 -> +i, 3 -> -i per spin-colour (lib/qudaQKXTM_utils.cpp:153-174)"""
    gl, _ = global_lex(Xloc, grid, coord)
    V = len(gl)
    out = np.zeros((V, 12, 2), dtype=np.float64)
    for k in range(12):
        r = np.floor(uniform(seed, 17, gl, k) * 4.0).astype(np.int64)
        out[r == 0, k, 0] = 1.0; out[r == 1, k, 0] = -1.0
        out[r == 2, k, 1] = 1.0; out[r == 3, k, 1] = -1.0
    return out.reshape(V, 4, 3, 2)


def spinor_eo_from_lex(psi_lex, Xloc):
    """[V][4][3][2] lexicographic -> [even Vh | odd Vh] (mapNormalToEvenOdd)"""
    return np.ascontiguousarray(psi_lex[eo_from_lex(Xloc)])


def spinor_lex_from_eo(psi_eo, Xloc):
    perm = eo_from_lex(Xloc)
    out = np.empty_like(psi_eo)
    out[perm] = psi_eo
    return out


def rel_l2(a, b):
    a = np.asarray(a); b = np.asarray(b)
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = a.astype(dt).ravel(); b = b.astype(dt).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# ---- independent dense restatement on the full lexicographic lattice (numpy, complex) ------------

def dense_hop(U_lex, psi_lex, X, gam, dagger=False):
    """(D psi)(x) = sum_mu (1 - g_mu) U_mu(x) psi(x+mu) + (1 + g_mu) U_mu(x-mu)^dag psi(x-mu).
    U_lex: [4][V][3][3] complex (with any boundary sign already folded in), psi_lex: [V][4][3] complex.
    Built with np.roll on a [t][z][y][x] grid: no checkerboarding, no index helper shared with the C
    oracle."""
    shp = (X[3], X[2], X[1], X[0])
    psi = psi_lex.reshape(shp + (4, 3))
    out = np.zeros_like(psi)
    one = np.eye(4)
    for mu in range(4):
        ax = 3 - mu  # grid axis of direction mu
        U = U_lex[mu].reshape(shp + (3, 3))
        Pm = one - gam[mu]; Pp = one + gam[mu]
        if dagger:
            Pm, Pp = Pp, Pm
        fwd = np.roll(psi, -1, axis=ax)                     # psi(x+mu)
        t1 = np.einsum("...ab,...sb->...sa", U, fwd)
        out += np.einsum("st,...tc->...sc", Pm, t1)
        t2 = np.einsum("...ba,...sb->...sa", np.conj(U), psi)  # U_mu(x)^dag psi(x)
        t2 = np.roll(t2, +1, axis=ax)                        # evaluated at x-mu
        out += np.einsum("st,...tc->...sc", Pp, t2)
    return out.reshape(-1, 4, 3)


def dense_mat(U_lex, psi_lex, X, gam, kappa, mu_tm, dagger=False):
    """M_full psi = (1 + i a g5) psi - kappa D psi, a = 2 kappa mu, g5 = g_x g_y g_z g_t"""
    g5 = gam[0] @ gam[1] @ gam[2] @ gam[3]
    a = 2.0 * kappa * mu_tm * (-1.0 if dagger else 1.0)
    A = np.eye(4) + 1j * a * g5
    return np.einsum("st,xtc->xsc", A, psi_lex) - kappa * dense_hop(U_lex, psi_lex, X, gam, dagger)


def dense_clover(U_lex, X, gam, coeff):
    """C(x) = 1 + i coeff sum_{mu<nu} sigma_munu (x) F_munu(x) as [V][4][3][4][3] complex, sigma_munu = (i/2)[g_mu, g_nu],
    F = (Q - Q^dag)/8 with Q the sum of the four plaquette leaves around x, all from np.roll on a [t][z][y][x] grid
    (independent of the C oracle's index code)."""
    shp = (X[3], X[2], X[1], X[0])
    U = [U_lex[mu].reshape(shp + (3, 3)) for mu in range(4)]
    V = int(np.prod(X))

    def sh(A, mu, d):            # field evaluated at x + d mu
        return np.roll(A, -d, axis=3 - mu)

    def mm(*ms):
        out = ms[0]
        for m in ms[1:]:
            out = np.einsum("...ab,...bc->...ac", out, m)
        return out

    def dg(A):
        return np.conj(np.swapaxes(A, -1, -2))

    C = np.zeros((V, 4, 3, 4, 3), dtype=np.complex128)
    for s in range(4):
        for c in range(3):
            C[:, s, c, s, c] = 1.0
    for mu in range(4):
        for nu in range(mu + 1, 4):
            Um, Un = U[mu], U[nu]
            Q = mm(Um, sh(Un, mu, 1), dg(sh(Um, nu, 1)), dg(Un))
            Q = Q + mm(Un, dg(sh(sh(Um, mu, -1), nu, 1)), dg(sh(Un, mu, -1)), sh(Um, mu, -1))
            Q = Q + mm(dg(sh(Um, mu, -1)), dg(sh(sh(Un, mu, -1), nu, -1)), sh(sh(Um, mu, -1), nu, -1), sh(Un, nu, -1))
            Q = Q + mm(dg(sh(Un, nu, -1)), sh(Um, nu, -1), sh(sh(Un, mu, 1), nu, -1), dg(Um))
            F = (Q - dg(Q)) / 8.0
            sig = 0.5j * (gam[mu] @ gam[nu] - gam[nu] @ gam[mu])
            C += 1j * coeff * np.einsum("st,xab->xsatb", sig, F.reshape(V, 3, 3))
    return C


def dense_mat_clover(U_lex, psi_lex, X, gam, kappa, mu_tm, coeff, dagger=False):
    """M_full psi = (C + i a g5) psi - kappa D psi"""
    g5 = gam[0] @ gam[1] @ gam[2] @ gam[3]
    a = 2.0 * kappa * mu_tm * (-1.0 if dagger else 1.0)
    C = dense_clover(U_lex, X, gam, coeff)
    out = np.einsum("xsatb,xtb->xsa", C, psi_lex) + 1j * a * np.einsum("st,xtc->xsc", g5, psi_lex)
    return out - kappa * dense_hop(U_lex, psi_lex, X, gam, dagger)


def c2r(z):
    return np.ascontiguousarray(np.stack([z.real, z.imag], axis=-1))


def r2c(a):
    return a[..., 0] + 1j * a[..., 1]


# ---- sharding helpers (host logic of the multi-GPU path) ---------------------------------------------

def rank_coord(rank, grid):
    """rank -> process-grid coordinate, t fastest then z (the order libtmq's rank_of() uses)"""
    ct = rank % grid[3]; cz = (rank // grid[3]) % grid[2]; cy = (rank // (grid[3] * grid[2])) % grid[1]
    cx = rank // (grid[3] * grid[2] * grid[1])
    return (cx, cy, cz, ct)


def coord_rank(coord, grid):
    return ((coord[0] * grid[1] + coord[1]) * grid[2] + coord[2]) * grid[3] + coord[3]


def local_from_global_eo(field_global_eo, Xloc, grid, coord):
    """slab of a FULL global even-odd field owned by the rank at `coord`, in the rank's local even-odd order"""
    G = tuple(Xloc[d] * grid[d] for d in range(4))
    gl, _ = global_lex(Xloc, grid, coord)                 # global lex index of each local lex site
    inv = np.empty(int(np.prod(G)), dtype=np.int64)
    inv[eo_from_lex(G)] = np.arange(len(inv))             # global lex -> global eo position
    loc_perm = eo_from_lex(Xloc)                          # local eo -> local lex
    return np.ascontiguousarray(field_global_eo[inv[gl.astype(np.int64)[loc_perm]]])
