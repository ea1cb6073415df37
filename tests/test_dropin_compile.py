"""CPU tests of the C++ drop-in boundary (include/qudaQKXTM_tmq.h, include/quda_tmq.h, include/compat/):

  * qudaQKXTMinfo / qudaQKXTM_arpackInfo / qudaQKXTM_loopInfo are passed BY VALUE by every entry point, so their layout is ABI: sizeof and
    every member's offset / size are static_asserted against the reference header compiled from where it lies
    (/root/reference/include/qudaQKXTM_utils.h:45-124), and checked against a committed golden layout on boxes without the reference;
  * the reference's own drivers -- qkxtm/MG_Bench.cpp, Calc_Loops.cpp, CalcMG_2pt3pt_EvenOdd.cpp, CalcLowModeProjection.cpp -- compile
    UNMODIFIED (whole files, including their parameter-setup blocks Calc_Loops.cpp:187-497,585-791) against include/compat;
  * libqkxtm_tmq.so exports the reference's entry points with exactly the reference's signatures (include/qudaQKXTM.h:484-513);
  * the drivers LINK against the product libraries (oracle/_ref/dropin/*, built by oracle/Makefile) and their own CLI runs."""
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
INC = os.path.join(ROOT, "include")
GOLDEN = os.path.join(ROOT, "tests", "golden", "qkxtm_struct_layout.json")
HOST_LIB = os.path.join(ROOT, "quda-qkxtm-multigrid-plugin_b200", "lib", "libqkxtm_tmq.so")
STRUCTS = ("qudaQKXTMinfo", "qudaQKXTM_arpackInfo", "qudaQKXTM_loopInfo")
have_ref = os.path.exists(os.path.join(REF, "include", "qudaQKXTM_utils.h"))
needs_ref = pytest.mark.skipif(not have_ref, reason="/root/reference is not present on this box")


def struct_members(text, name):
    """member names of `typedef struct{ ... } name;` in declaration order"""
    text = re.sub(r"//[^\n]*", "", text)
    m = re.search(r"typedef\s+struct\s*\{([^{}]*)\}\s*%s\s*;" % name, text)
    assert m, name
    body = m.group(1)
    body = "\n".join(ln for ln in body.splitlines() if not ln.strip().startswith("#"))
    names = []
    for decl in body.split(";"):
        decl = decl.split("=")[0].strip()
        if not decl:
            continue
        mm = re.search(r"(\w+)\s*(\[[^\]]*\]\s*)*$", decl)
        assert mm, decl
        names.append(mm.group(1))
    return names


def ours_layout_program(members):
    lines = ['#include <cstddef>', '#include <cstdio>', '#include "qudaQKXTM_tmq.h"', 'int main() {', '  printf("{");']
    for si, (s, mem) in enumerate(members.items()):
        lines.append('  printf("%s\\"%s\\": {\\"sizeof\\": %%zu, \\"members\\": {", sizeof(quda::%s));' % (", " if si else "", s, s))
        for mi, m in enumerate(mem):
            lines.append('  printf("%s\\"%s\\": [%%zu, %%zu]", offsetof(quda::%s, %s), sizeof(((quda::%s *)0)->%s));' % (", " if mi else "", m, s, m, s, m))
        lines.append('  printf("}}");')
    lines += ['  printf("}\\n");', '  return 0;', '}']
    return "\n".join(lines)


def compile_and_run(tmp_path, src, name, extra=()):
    cpp = tmp_path / (name + ".cpp")
    cpp.write_text(src)
    exe = tmp_path / name
    p = subprocess.run(["g++", "-std=c++11", "-Wno-invalid-offsetof", "-I", INC, *extra, str(cpp), "-o", str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]
    return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout


def test_by_value_struct_layouts_match_the_golden_layout(tmp_path):
    """runs everywhere: our header against the layout recorded from the reference header (tests/golden/qkxtm_struct_layout.json)"""
    golden = json.load(open(GOLDEN))
    members = {s: list(golden[s]["members"].keys()) for s in STRUCTS}
    ours = json.loads(compile_and_run(tmp_path, ours_layout_program(members), "layout_ours"))
    assert ours == golden


@needs_ref
def test_by_value_struct_layouts_match_the_reference_header(tmp_path):
    ref_text = open(os.path.join(REF, "include", "qudaQKXTM_utils.h")).read()
    a = ref_text.index("enum SOURCE_T")
    b = ref_text.index("enum APEDIM{D3,D4};") + len("enum APEDIM{D3,D4};")
    ref_block = ref_text[a:b]
    defines = re.findall(r"^#define\s+(QUDAQKXTM_DIM|MAX_\w+)\s+(\d+)\s*$", ref_text, flags=re.M)
    assert {"QUDAQKXTM_DIM", "MAX_NSOURCES", "MAX_NMOMENTA", "MAX_TSINK", "MAX_DEFLSTEPS", "MAX_PROJS"} <= {d[0] for d in defines}
    members = {s: struct_members(ref_block, s) for s in STRUCTS}
    assert "thrp_type" in members["qudaQKXTMinfo"] and "HighMomForm" in members["qudaQKXTMinfo"] and "deflStep" in members["qudaQKXTM_loopInfo"]
    src = ['#include <cstddef>', '#include "qudaQKXTM_tmq.h"']
    for n, v in defines:   # our macro values first, then the reference's own #define lines take over inside namespace ref
        src.append("static_assert(%s == %s, \"%s differs from the reference\");" % (n, v, n))
        src.append("#undef %s" % n)
        src.append("#define %s %s" % (n, v))
    src += ["#define HAVE_ARPACK", "namespace ref {", "typedef ::QudaPrecision QudaPrecision;", ref_block, "}"]
    for s, mem in members.items():
        src.append("static_assert(sizeof(quda::%s) == sizeof(ref::%s), \"sizeof(%s)\");" % (s, s, s))
        for m in mem:
            src.append("static_assert(offsetof(quda::%s, %s) == offsetof(ref::%s, %s), \"offsetof(%s, %s)\");" % (s, m, s, m, s, m))
            src.append("static_assert(sizeof(((quda::%s *)0)->%s) == sizeof(((ref::%s *)0)->%s), \"sizeof(%s::%s)\");" % (s, m, s, m, s, m))
    for e, vals in (("SOURCE_T", ("UNITY", "RANDOM")), ("CORR_SPACE", ("POSITION_SPACE", "MOMENTUM_SPACE")), ("FILE_WRITE_FORMAT", ("ASCII_FORM", "HDF5_FORM")),
                    ("WHICHSPECTRUM", ("SR", "LR", "SM", "LM", "SI", "LI")), ("ALLOCATION_FLAG", ("NONE", "HOST", "DEVICE", "BOTH", "BOTH_EXTRA")),
                    ("CLASS_ENUM", ("FIELD", "GAUGE", "VECTOR", "PROPAGATOR", "PROPAGATOR3D", "VECTOR3D")), ("WHICHPARTICLE", ("PROTON", "NEUTRON")),
                    ("WHICHPROJECTOR", ("G4", "G5G123", "G5G1", "G5G2", "G5G3"))):
        for v in vals:
            src.append("static_assert((int)quda::%s == (int)ref::%s, \"%s::%s\");" % (v, v, e, v))
    src.append("int main() { return 0; }")
    compile_and_run(tmp_path, "\n".join(src), "layout_vs_ref")
    # and the golden file is what the reference header gives (regenerate with: python tests/test_dropin_compile.py)
    ours = json.loads(compile_and_run(tmp_path, ours_layout_program(members), "layout_ours"))
    assert ours == json.load(open(GOLDEN))


@needs_ref
@pytest.mark.parametrize("driver", ["MG_Bench", "Calc_Loops", "CalcMG_2pt3pt_EvenOdd", "CalcLowModeProjection"])
def test_reference_driver_compiles_unmodified(driver):
    """the whole reference driver, from where it lies, against the compat headers of this repository + the reference's own
    include/QKXTM_util.h; -DHAVE_ARPACK as the reference's build sets it for the eigensolver drivers"""
    p = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-DHAVE_ARPACK", "-I", os.path.join(INC, "compat"), "-I", os.path.join(REF, "include"),
                        os.path.join(REF, "qkxtm", driver + ".cpp")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]


def test_entry_points_have_the_reference_signatures():
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", HOST_LIB], capture_output=True, text=True, check=True).stdout
    for sig in ("MG_bench(void**, void**, QudaGaugeParam_s*, QudaInvertParam_s*, quda::qudaQKXTMinfo)",
                "calcMG_threepTwop_EvenOdd(void**, void**, QudaGaugeParam_s*, QudaInvertParam_s*, quda::qudaQKXTMinfo, char*, char*, quda::WHICHPARTICLE)",
                "calc_loops(void**, QudaInvertParam_s*, QudaInvertParam_s*, QudaGaugeParam_s*, quda::qudaQKXTM_arpackInfo, quda::qudaQKXTM_loopInfo, quda::qudaQKXTMinfo)",
                "calcLowModeProjection(QudaInvertParam_s*, quda::qudaQKXTM_arpackInfo)",
                "quda::init_qudaQKXTM(quda::qudaQKXTMinfo*)", "quda::printf_qudaQKXTM()",
                "quda::QKXTM_Deflation<double>::projectVector(quda::QKXTM_Vector<double>&, quda::QKXTM_Vector<double>&, int, int)",
                "quda::QKXTM_Deflation<double>::MapEvenOddToFull()",
                "readLimeGauge(void**, char*, QudaGaugeParam_s*, QudaInvertParam_s*, int*)", "applyBoundaryCondition(void**, int, QudaGaugeParam_s*)"):
        assert sig in out, sig
    plain = subprocess.run(["nm", "-D", "--defined-only", HOST_LIB], capture_output=True, text=True, check=True).stdout
    for c_name in ("initQuda", "endQuda", "loadGaugeQuda", "freeGaugeQuda", "loadCloverQuda", "freeCloverQuda", "invertQuda", "newQudaGaugeParam",
                   "newQudaInvertParam", "newQudaMultigridParam", "newMultigridQuda", "destroyMultigridQuda", "initCommsGridQuda", "setVerbosityQuda",
                   "comm_rank", "comm_size", "comm_coord", "comm_dim", "comm_dim_partitioned", "comm_barrier", "comm_coords", "getVerbosity"):
        assert re.search(r"\bT %s$" % c_name, plain, flags=re.M), c_name


@pytest.mark.parametrize("driver", ["MG_Bench", "Calc_Loops", "CalcMG_2pt3pt_EvenOdd", "CalcLowModeProjection"])
def test_reference_driver_links_against_the_product_libraries(driver):
    """oracle/Makefile compiles qkxtm/<driver>.cpp + qkxtm/QKXTM_util.cpp + qkxtm/misc.cpp from /root/reference and links them with
    -lqkxtm_tmq -ltmq: every symbol the reference's driver side needs is there.  Its own command-line parser then runs."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin", driver)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin not built (needs /root/reference at build time)")
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libqkxtm_tmq.so" in ldd and "libtmq.so" in ldd and "not found" not in ldd, ldd
    p = subprocess.run([exe, "--help"], capture_output=True, text=True, timeout=60)
    assert "--load-gauge" in p.stdout and "--dslash-type" in p.stdout


if __name__ == "__main__":   # regenerate the golden layout from the reference header
    import tempfile
    import pathlib
    ref_text = open(os.path.join(REF, "include", "qudaQKXTM_utils.h")).read()
    blk = ref_text[ref_text.index("enum SOURCE_T"):ref_text.index("enum APEDIM{D3,D4};")]
    mem = {s: struct_members(blk, s) for s in STRUCTS}
    with tempfile.TemporaryDirectory() as d:
        lay = json.loads(compile_and_run(pathlib.Path(d), ours_layout_program(mem), "layout_ours"))
    json.dump(lay, open(GOLDEN, "w"), indent=1)
    print("wrote", GOLDEN)
