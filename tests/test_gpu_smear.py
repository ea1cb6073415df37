"""GPU parity tests of the Gaussian smearing and of the container layout kernels against the
REFERENCE'S OWN kernel bodies: the golden fixture produced from lib/code_pieces/*_core.h compiled for the CPU
(tests/golden/make_golden_ref.py) and, for larger / other shapes, the prebuilt oracle/_ref library run live and the
oracle's numpy restatement.  fp64 tolerance 1e-13 per application (accumulation order differs from the reference's
P1 + P2 sums), fp32 1e-5."""
import os
import sys

import numpy as np
import pytest

import lattice_util as lu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_ref as G  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tmq():
    import tmq as T
    if T.load().tmq_device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    return T


def _c(a):
    return a[..., 0] + 1j * a[..., 1]


class Dev:
    """QKXTM-layout device buffers of one context"""

    def __init__(self, T, X):
        self.T, self.c = T, T.Context(X)
        self.V = int(np.prod(X))
        self.bufs = []

    def put(self, arr):
        a = np.ascontiguousarray(arr)
        p = self.c.dev_malloc(a.nbytes)
        self.c.h2d(p, a)
        self.bufs.append(p)
        return p

    def get(self, p, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        self.c.d2h(out, p)
        return out

    def close(self):
        for p in self.bufs:
            self.c.dev_free(p)
        self.c.close()


def smear(dev, vec, gauge, prec, nsmear, alpha):
    dt = np.float64 if prec == 8 else np.float32
    din, dg = dev.put(vec.astype(dt)), dev.put(gauge.astype(dt))
    dout = dev.put(np.zeros_like(vec, dtype=dt))
    dev.c.qkxtm_gauss_smear(dout, din, dg, prec, nsmear, alpha)
    return dev.get(dout, vec.shape, dt)


def test_gauss_smear_matches_reference_fixture(tmq):
    gold = np.load(G.FIXTURE)
    vec, gauge = G.golden_inputs()
    d = Dev(tmq, G.X)
    g = gauge.reshape(36, -1, 2)
    assert lu.rel_l2(_c(smear(d, vec, g, 8, 1, G.ALPHA)), _c(gold["gauss_step"])) < 1e-14
    assert lu.rel_l2(_c(smear(d, vec, g, 8, G.NSMEAR, G.ALPHA)), _c(gold["gauss_smear3"])) < 1e-14
    assert lu.rel_l2(_c(smear(d, vec, g, 4, 1, G.ALPHA)), _c(gold["gauss_step_f32"])) < 1e-6
    assert np.array_equal(smear(d, vec, g, 8, 0, G.ALPHA), vec)           # nsmear = 0 copies
    d.close()


@pytest.mark.parametrize("X,nsmear", [((8, 6, 4, 10), 5), ((16, 16, 16, 8), 50)])
def test_gauss_smear_matches_reference_library_and_is_block_order_independent(tmq, X, nsmear):
    """the reference's kernel body run live (prebuilt oracle/_ref) when present, else the numpy restatement; the L2-blocked
    sweep order (time slices outer, steps inner) gives the same bits as the plain streaming order"""
    from oracle import ref
    from oracle.oracle import gauss_smear
    rng = np.random.default_rng(3)
    V = int(np.prod(X))
    vec = rng.standard_normal((12, V, 2))
    U = lu.random_su3_lex(X, seed=11)                               # SU(3) links keep 50 steps well conditioned
    # [dir][x_lex][c1][c2] -> QKXTM device layout [dir][c1][c2][x_lex][re,im] (lib/qudaQKXTM_Gauge.cpp:73-89)
    Ut = np.transpose(U, (0, 2, 3, 1))
    gq = np.ascontiguousarray(np.stack([Ut.real, Ut.imag], axis=-1)).reshape(36, V, 2)
    alpha = 4.0
    if ref.available():
        want = _c(ref.Ref(X, alpha_gauss=alpha).gauss_smear(vec, gq.reshape(4, 3, 3, V, 2), nsmear))
    else:
        want = gauss_smear(_c(vec), _c(gq).reshape(4, 3, 3, V), X, alpha, nsmear)
    d = Dev(tmq, X)
    outs = []
    for block in (0, 1, 3, 10 ** 6):
        d.c.set_option(tmq.OPT_SMEAR_BLOCK_T, block)
        outs.append(smear(d, vec, gq, 8, nsmear, alpha))
    assert lu.rel_l2(_c(outs[0]), want) < 1e-13
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])
    d.close()


def test_container_layout_kernels_match_reference_fixture(tmq):
    """uploadToCuda / downloadFromCuda / scaleVector / apply_gamma5 against the reference's kernel
    bodies: the native field read back in host even-odd order equals the reference's even / odd blocks"""
    gold = np.load(G.FIXTURE)
    vec, _ = G.golden_inputs()
    d = Dev(tmq, G.X)
    c, V = d.c, d.V
    Vh = V // 2
    dq = d.put(vec)
    full = c.spinor(8, tmq.FULL)
    c.from_qkxtm(full, dq, 8, -1)
    eo = full.get()                                                     # [even Vh | odd Vh][4][3][2]
    even = np.transpose(eo[:Vh], (1, 2, 0, 3)).reshape(12, Vh, 2)
    odd = np.transpose(eo[Vh:], (1, 2, 0, 3)).reshape(12, Vh, 2)
    assert np.array_equal(even, gold["upload_even"]) and np.array_equal(odd, gold["upload_odd"])
    # download of the even parity only: the odd sites are zero-filled
    par = c.spinor(8)
    c.from_qkxtm(par, dq, 8, 0)
    dz = d.put(np.full_like(vec, 7.0))
    c.to_qkxtm(dz, par, 8, 0, 1.0)
    assert np.array_equal(d.get(dz, vec.shape), gold["download_even_only"])
    c.to_qkxtm(dz, full, 8, -1, 1.0)
    assert np.array_equal(d.get(dz, vec.shape), gold["download_both"])
    # scaleVector and the fused 2 kappa rescale of the download
    c.qkxtm_scale(dq, 8, 2 * 0.1234)
    assert np.array_equal(d.get(dq, vec.shape), gold["scale"])
    c.to_qkxtm(dz, full, 8, -1, 2 * 0.1234)
    assert np.array_equal(d.get(dz, vec.shape), gold["scale"])
    dg5 = d.put(vec)
    c.qkxtm_gamma5(dg5, 8)
    assert np.array_equal(d.get(dg5, vec.shape), gold["gamma5"])
    # QKXTM_Gauge::calculatePlaq on the container's device layout vs the reference's plaquette kernel
    _, gauge = G.golden_inputs()
    dgauge = d.put(gauge)
    want = float(gold["plaquette"][0])
    assert abs(c.qkxtm_plaquette(dgauge, 8) - want) < 1e-12 * abs(want)
    d.close()


@pytest.mark.parametrize("prec", [8, 4])
def test_gauss_smear_on_a_z_partitioned_lattice_is_bit_identical(tmq, prec):
    """a z split (T, then Z) needs the two z faces of the vector before every step and the neighbour's U_z face once
    (the reference's ghost zone, lib/qudaQKXTM_Vector.cpp:172-382): with the ghost path forced onto a single rank
    (tmq_force_partition, the reference's --partition) the exchange wraps onto this rank and every bit must agree with the
    plain periodic sweep, which is pinned to the reference's kernel above"""
    X, nsmear, alpha = (6, 4, 8, 6), 7, 4.0
    rng = np.random.default_rng(23)
    V = int(np.prod(X))
    vec = rng.standard_normal((12, V, 2))
    U = lu.random_su3_lex(X, seed=5)
    Ut = np.transpose(U, (0, 2, 3, 1))
    gq = np.ascontiguousarray(np.stack([Ut.real, Ut.imag], axis=-1)).reshape(36, V, 2)
    d0 = Dev(tmq, X)
    plain = smear(d0, vec, gq, prec, nsmear, alpha)
    d0.close()
    for part in ((0, 0, 1, 0), (0, 0, 1, 1)):
        d1 = Dev(tmq, X)
        d1.c.force_partition(part)
        d1.c.set_option(tmq.OPT_SMEAR_BLOCK_T, 2)             # ignored on a z split
        ghost = smear(d1, vec, gq, prec, nsmear, alpha)
        assert np.array_equal(ghost, plain), part
        assert np.array_equal(smear(d1, vec, gq, prec, 0, alpha), vec.astype(ghost.dtype))
        d1.close()


@pytest.mark.parametrize("part", [(0, 0, 0, 1), (0, 0, 1, 0), (0, 0, 1, 1)])
@pytest.mark.parametrize("ncomp,prec", [(36, 8), (12, 4), (144, 8)])
def test_container_ghost_exchange_layout_and_plaquette(tmq, part, ncomp, prec):
    """the containers' ghost trio (ghostToHost / cpuExchangeGhost / ghostToDevice, lib/qudaQKXTM_Gauge.cpp:143-373) as ONE device-side
    exchange: behind the [ncomp][V] array, per partitioned dimension in ascending order, the plus ghost = the forward neighbour's slice 0
    and the minus ghost = the backward neighbour's slice L-1, each [ncomp][surface] with the lexicographic face index of the other three
    coordinates (lib/qudaQKXTM_kernels.cu:160-170, plaquette_core.h:30-50).  With the partition forced onto one rank the neighbours are
    the rank itself.  The plaquette read through the ghost zones must equal the periodic one."""
    X = (6, 4, 8, 6)
    V = int(np.prod(X))
    dt = np.float64 if prec == 8 else np.float32
    rng = np.random.default_rng(ncomp)
    d = Dev(tmq, X)
    d.c.force_partition(part)
    nghost = d.c.qkxtm_ghost_sites()
    surf = {2: V // X[2], 3: V // X[3]}
    assert nghost == sum(2 * surf[dim] for dim in (2, 3) if part[dim])
    field = rng.standard_normal((ncomp, V, 2)).astype(dt)
    buf = np.concatenate([field.reshape(-1), np.full(nghost * ncomp * 2, -7.0, dtype=dt)])
    p = d.put(buf)
    d.c.qkxtm_exchange_ghost(p, prec, ncomp)
    got = d.get(p, buf.shape, dt)
    assert np.array_equal(got[: field.size], field.reshape(-1))
    f4 = field.reshape(ncomp, X[3], X[2], X[1], X[0], 2)
    off = field.size
    for dim in (2, 3):
        if not part[dim]:
            continue
        lo = f4[:, 0] if dim == 3 else f4[:, :, 0]                 # slice 0 of the forward neighbour (= this rank)
        hi = f4[:, -1] if dim == 3 else f4[:, :, -1]               # slice L-1 of the backward neighbour
        n = ncomp * surf[dim] * 2
        assert np.array_equal(got[off: off + n], np.ascontiguousarray(lo).reshape(-1)), ("plus", dim)
        assert np.array_equal(got[off + n: off + 2 * n], np.ascontiguousarray(hi).reshape(-1)), ("minus", dim)
        off += 2 * n
    assert off == buf.size
    if ncomp == 36:
        U = lu.random_su3_lex(X, seed=5)
        Ut = np.transpose(U, (0, 2, 3, 1))
        gq = np.ascontiguousarray(np.stack([Ut.real, Ut.imag], axis=-1)).reshape(36, V, 2)
        d0 = Dev(tmq, X)
        want = d0.c.qkxtm_plaquette(d0.put(gq), 8)
        d0.close()
        pg = d.put(np.concatenate([gq.reshape(-1), np.zeros(nghost * 36 * 2)]))
        assert abs(d.c.qkxtm_plaquette(pg, 8) - want) < 1e-13 * abs(want)
    d.close()
