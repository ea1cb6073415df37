"""Generates tests/golden/qkxtm_ref_contract_4x4x4x6.npz from the REFERENCE'S OWN kernel bodies compiled for the CPU by
oracle/Makefile (oracle/_ref/libqkxtm_ref.so, needs /root/reference): the meson two-point contraction
(lib/code_pieces/contractMesons_core.h, contractMesons_core_PosSpace.h, with the reference's channel tables
lib/qudaQKXTM_kernels.cu:77-78) and the site-local propagator kernels (rotateToPhysicalBase_core.h,
apply_gamma5_propagator_core.h, conjugate_propagator_core.h, conjugate_vector_core.h).  Inputs come from a seeded numpy
generator (contract_inputs below, also used by the tests); the fixture stores the reference's outputs only.
Run:  python tests/golden/make_golden_contract.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

X = (4, 4, 4, 6)
Q_SQ = 3                      # 27 momenta
SRC = (1, 2, 3)               # source position (x0, y0, z0)
SAMPLE = np.array([0, 1, 5, 63, 64, 100, 191, 200, 255, 300, 383])     # sites kept of the site-local outputs
X_SMALL = (2, 2, 2, 2)        # a second, tiny lattice: the numpy restatement of the baryon contraction is slow
SRC_SMALL = (1, 0, 1)
FIXTURE = os.path.join(HERE, "qkxtm_ref_contract_4x4x4x6.npz")


def contract_inputs():
    """prop1, prop2 [4 mu][4 nu][3 c1][3 c2][V][2] (the contraction is a plain function of the numbers: random ones exercise
    every index of the tables)"""
    rng = np.random.Generator(np.random.PCG64(20171008))
    V = int(np.prod(X))
    return rng.standard_normal((4, 4, 3, 3, V, 2)), rng.standard_normal((4, 4, 3, 3, V, 2))


# sequential-source cases kept in the fixture: key -> (part, projector, particle, source spin, source colour)
SEQ_CASES = {"seq1_G4_proton_00": (1, 0, 0, 0, 0), "seq1_G5G123_neutron_21": (1, 1, 1, 2, 1), "seq1_G5G2_proton_32": (1, 3, 0, 3, 2),
             "seq2_G4_proton_11": (2, 0, 0, 1, 1), "seq2_G5G3_neutron_30": (2, 4, 1, 3, 0), "seq2_G5G123_proton_02": (2, 1, 0, 0, 2)}


def momenta():
    from oracle.oracle import create_momenta
    return create_momenta(Q_SQ)


def baryon_momenta():
    return momenta()[::6]     # 5 of the 27: the baryon output is 16 times the meson one


def deriv_gauge():
    """links for the conserved-current / one-derivative insertions [4][3][3][V][2] (the contraction is linear in them)"""
    rng = np.random.Generator(np.random.PCG64(20171010))
    return rng.standard_normal((4, 3, 3, int(np.prod(X)), 2))


def small_inputs():
    rng = np.random.Generator(np.random.PCG64(20171009))
    V = int(np.prod(X_SMALL))
    return rng.standard_normal((4, 4, 3, 3, V, 2)), rng.standard_normal((4, 4, 3, 3, V, 2))


if __name__ == "__main__":
    from oracle.ref import Ref
    p1, p2 = contract_inputs()
    f1, f2 = p1.astype(np.float32), p2.astype(np.float32)
    r = Ref(X)
    moms = momenta()
    vec = np.ascontiguousarray(p1[:, 0, :, 0])                 # [4][3][V][2] as a vector
    out = {
        "mom_float": r.contract_mesons_mom(f1, f2, moms, SRC),                  # what the reference launches (float only)
        "mom_double": r.contract_mesons_mom(p1, p2, moms, SRC),                 # contractMesons_kernel_double, same body
        "pos_float": r.contract_mesons_pos(f1, f2),
        "rotate_plus": r.rotate_physical(p1, +1)[..., SAMPLE, :],
        "rotate_minus": r.rotate_physical(p1, -1)[..., SAMPLE, :],
        "rotate_plus_f32": r.rotate_physical(f1, +1)[..., SAMPLE, :],
        "gamma5_prop": r.gamma5_propagator(p1)[..., SAMPLE, :],
        "conj_prop": r.conjugate_propagator(p1)[..., SAMPLE, :],
        "conj_vec": r.conjugate_vector(vec.reshape(12, -1, 2))[..., SAMPLE, :],
        # contractBaryons_core.h with the tables of lib/qudaQKXTM_kernels.cu:79-88: [T][nmoms][2][10][4][4][2]
        "baryon_mom_float": r.contract_baryons_mom(f1, f2, baryon_momenta(), SRC),      # what the reference launches (float only)
        "baryon_mom_double": r.contract_baryons_mom(p1, p2, baryon_momenta(), SRC),     # the same body instantiated in double
    }
    # fixed-sink three-point function: the projector / operator tables, sequential sources and the ultra-local insertion
    out["proj_tables"] = np.array([[r.projector(pid, part) for part in range(2)] for pid in range(5)])
    out["op_tables"] = np.array([[[r.operator(f, part, pf) for pf in (1, 2)] for part in range(2)] for f in range(16)])
    V3 = int(np.prod(X[:3]))
    t1, t2 = np.ascontiguousarray(p1[..., 2 * V3:3 * V3, :]), np.ascontiguousarray(p2[..., 2 * V3:3 * V3, :])     # 3-d propagators: time slice 2
    for key, (part, pid, particle, nu, c2) in SEQ_CASES.items():
        out[key] = r.seq_source(part, 4, t1, t2, nu, c2, pid, particle)[:, 4 * V3:5 * V3]
    out["thrp_local_double"] = r.fixsink_local(p1, p2, 0, 1, baryon_momenta(), SRC)
    out["thrp_local_float"] = r.fixsink_local(f1, f2, 1, 1, baryon_momenta(), SRC)
    out["thrp_noether_double"], out["thrp_oneD_double"] = r.fixsink_derivative(p1, p2, deriv_gauge(), 1, 2, baryon_momenta(), SRC)
    s1, s2 = small_inputs()
    out["baryon_small_double"] = Ref(X_SMALL).contract_baryons_mom(s1, s2, [(0, 0, 0), (1, 0, -1)], SRC_SMALL)
    np.savez_compressed(FIXTURE, **out)
    print("written", FIXTURE, os.path.getsize(FIXTURE), "bytes")
