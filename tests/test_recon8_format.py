"""CPU check of the 8-real link format's arithmetic (tools/recon8_study.py is the numpy statement of csrc/tmq_site.cuh: reconstruct_from8):
random SU(3) links -- incl. links carrying the anti-periodic boundary sign -- come back from U01, U02, U10, tan(arg U00 / 4),
tan(arg U20 / 4) to rounding; the GPU side is tests/test_gpu_parity.py::test_reconstruct_8_matches_oracle_and_the_12_real_path."""
import os
import sys

import numpy as np

import lattice_util as lu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import recon8_study as R  # noqa: E402


def test_eight_real_format_round_trip():
    U = lu.random_su3_lex((4, 4, 4, 8), seed=137).reshape(-1, 3, 3)
    sign = np.ones(len(U)); sign[::3] = -1
    U = U * sign[:, None, None]
    V = R.unpack(*R.pack(U), sign)
    assert np.linalg.norm(V - U) / np.linalg.norm(U) < 1e-14
    assert np.abs(V - U).max() < 1e-11
    # the result is unitary up to the boundary sign and has the right determinant
    assert np.abs(np.einsum("nab,ncb->nac", V, V.conj()) - np.eye(3)).max() < 1e-11
    assert np.abs(np.linalg.det(V) - sign).max() < 1e-11


def test_phase_parametrisation_has_no_singular_point():
    th = np.concatenate([np.linspace(-np.pi, np.pi, 100001), [np.pi - 1e-12, -np.pi + 1e-12, 0.0]])
    assert np.abs(R.phase(np.tan(th / 4)) - np.exp(1j * th)).max() < 2e-15
