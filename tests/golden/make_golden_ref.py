"""Generates tests/golden/qkxtm_ref_4x4x4x6.npz from the REFERENCE'S OWN kernel bodies (lib/code_pieces/*_core.h) compiled
for the CPU by oracle/Makefile (oracle/_ref/libqkxtm_ref.so, needs /root/reference).  Inputs come from a seeded numpy
generator (golden_inputs below, also used by the tests); the fixture stores the reference's outputs only.
Run:  python tests/golden/make_golden_ref.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

X = (4, 4, 4, 6)
ALPHA = 4.0          # --alphaGauss default of the drivers (qkxtm/QKXTM_util.cpp:1654)
NSMEAR = 3
FIXTURE = os.path.join(HERE, "qkxtm_ref_4x4x4x6.npz")


def golden_inputs():
    """vec [12][V][2], gauge [4][3][3][V][2] (random SU(3) is not needed: the kernels are linear in both)"""
    rng = np.random.Generator(np.random.PCG64(20171007))
    V = int(np.prod(X))
    vec = rng.standard_normal((12, V, 2))
    gauge = rng.standard_normal((4, 3, 3, V, 2))
    return vec, gauge


if __name__ == "__main__":
    from oracle.ref import Ref
    vec, gauge = golden_inputs()
    r = Ref(X, alpha_gauss=ALPHA)
    even, odd = r.upload(vec)
    out = {
        "gauss_step": r.gauss_step(vec, gauge),                                  # Gauss_core.h, one step
        "gauss_smear3": r.gauss_smear(vec, gauge, NSMEAR),                       # lib/qudaQKXTM_Vector.cpp:386-421
        "gauss_step_f32": r.gauss_step(vec.astype(np.float32), gauge.astype(np.float32)),
        "upload_even": even, "upload_odd": odd,                                  # uploadToCuda_core.h
        "download_even_only": r.download(even, None),                            # downloadFromCuda_core.h (absent parity zero-filled)
        "download_both": r.download(even, odd),
        "scale": r.scale(vec, 2 * 0.1234),                                       # scaleVector_core.h
        "gamma5": r.gamma5(vec),                                                 # apply_gamma5_vector_core.h
        "plaquette": np.array([r.plaquette(gauge)]),                             # plaquette_core.h + lib/qudaQKXTM_kernels.cu:910-957
    }
    np.savez_compressed(FIXTURE, **out)
    print("written", FIXTURE, os.path.getsize(FIXTURE), "bytes")
