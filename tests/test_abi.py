"""CPU-side checks of the drop-in boundary: libtmq.so loads, exports every symbol include/tmq.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200")


def header_symbols(path):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tmq_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def tmq():
    import tmq as T
    if not os.path.exists(T.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return T


def test_every_declared_symbol_is_exported(tmq):
    L = tmq.load()
    declared = header_symbols(os.path.join(ROOT, "include", "tmq.h"))
    assert len(declared) >= 55
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(tmq.SYMBOLS) == declared, set(tmq.SYMBOLS) ^ set(declared)


def test_host_library_exports_fieldgen_and_shim(tmq):
    H = tmq.load_host()
    for s in header_symbols(os.path.join(ROOT, "include", "tmq_host.h")):
        assert hasattr(H, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", tmq.HOST_LIB_PATH], capture_output=True, text=True).stdout
    for name in ("invertQuda", "loadGaugeQuda", "initQuda", "MG_bench", "calc_loops_solve", "ApplyMdagM", "quda::init_qudaQKXTM",
                 "quda::QKXTM_Vector<double>::packVector", "quda::QKXTM_Vector<double>::uploadToCuda",
                 "quda::QKXTM_Vector<double>::downloadFromCuda", "quda::QKXTM_Gauge<double>::calculatePlaq",
                 "quda::QKXTM_Propagator<double>::absorbVectorToDevice", "quda::QKXTM_Vector<double>::gaussianSmearing",
                 "quda::QKXTM_Vector<double>::write", "quda::QKXTM_Deflation<double>::eigenSolver",
                 "quda::QKXTM_Deflation<double>::deflateVector", "quda::QKXTM_Deflation<double>::projectVector",
                 "quda::QKXTM_Deflation<double>::polynomialOperator", "loadCloverQuda", "readLimeGauge", "applyBoundaryCondition",
                 "quda::testGaussSmearing", "quda::testPlaquette"):
        assert name in out, name


def test_no_signature_leaks_torch_or_cxx_types():
    txt = open(os.path.join(ROOT, "include", "tmq.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)          # signatures only, not the prose
    assert "torch" not in txt and "std::" not in txt and "at::" not in txt
    assert 'extern "C"' in txt


def test_product_never_links_or_loads_the_oracle(tmq):
    for lib in (tmq.LIB_PATH, tmq.HOST_LIB_PATH):
        deps = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
        assert "tm_oracle" not in deps
        syms = subprocess.run(["nm", "-D", lib], capture_output=True, text=True).stdout
        assert "orc_" not in syms
    bad = re.compile(r"import\s+oracle|from\s+oracle|tm_oracle|libtm_oracle|orc_[a-z]|libqkxtm_ref|libqkxtm_util_ref|qref_|qutil_")
    for root, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                assert not bad.search(open(os.path.join(root, f)).read()), os.path.join(root, f)


def test_fails_loudly_without_a_gpu(tmq):
    L = tmq.load()
    if L.tmq_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(tmq.TmqError, match="no usable CUDA device"):
        tmq.Context((4, 4, 4, 4))
    drv = os.path.join(PKG, "lib", "qkxtm_invert_test")
    p = subprocess.run([drv, "--dim", "4", "4", "4", "4"], capture_output=True, text=True)
    assert p.returncode != 0 and "no CUDA device" in p.stderr
