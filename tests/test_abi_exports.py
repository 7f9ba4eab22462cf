"""CPU: the C-ABI library builds (nvcc cross-compiles without a GPU), loads, and exports every symbol
include/ekf_b200.h declares; the ctypes mirrors agree with the C structs.  No compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ekf_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(ekf_[a-zA-Z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_expected_surface():
    names = _declared_functions()
    # the reference's public VSlamFilter members on the path (vslamRansac.hpp:99-140), one entry each
    for must in ("ekf_create", "ekf_destroy", "ekf_capture_frame", "ekf_predict", "ekf_update", "ekf_add_feature",
                 "ekf_remove_feature", "ekf_get_state", "ekf_get_sigma", "ekf_covariance_parameter", "ekf_num_features",
                 "ekf_get_dt", "ekf_get_center", "ekf_convert2xyz_if_linear", "ekf_convert2xyz_if_linear_all",
                 "ekf_match_batch", "ekf_batch_create", "ekf_batch_step"):
        assert must in names, must


def test_library_builds_and_exports_every_declared_symbol(pkg):
    path = pkg.build()
    assert os.path.exists(path)
    L = C.CDLL(path)
    missing = [n for n in _declared_functions() if not hasattr(L, n)]
    assert not missing, f"declared in include/ekf_b200.h but not exported: {missing}"
    from importlib import import_module
    sig = import_module("ekf_b200._lib").SIGNATURES
    unbound = [n for n in _declared_functions() if n not in sig]
    assert not unbound, f"declared but without a ctypes signature in _lib.py: {unbound}"
    info = L.ekf_build_info
    info.restype = C.c_char_p
    assert b"sm_100a" in info()


def test_sass_is_sm100a_with_dmma(pkg):
    """The shipped cubin is sm_100a and the downdate GEMM uses the fp64 tensor pipe (DMMA)."""
    path = pkg.build()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn = sass.split("Function : ")
    gemm = [f for f in fn if f.startswith("_Z13k_gemm_nt_sub")]
    upd = [f for f in fn if f.startswith("_Z14k_batch_update")]
    assert gemm and "DMMA.8x8x4" in gemm[0]
    assert upd and "DMMA.8x8x4" in upd[0]


def test_ctypes_structs_match_c_layout(pkg, tmp_path):
    """sizeof / offsetof of every struct that crosses the boundary, from the C compiler."""
    abi = pkg._abi
    structs = {"ekf_config": abi.EkfConfig, "ekf_feature_info": abi.EkfFeatureInfo, "ekf_step_stats": abi.EkfStepStats,
               "ekf_profile": abi.EkfProfile, "ekf_batch_desc": abi.EkfBatchDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "lay.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "lay"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def test_default_config_matches_reference_defaults(pkg):
    """ConfigVSLAM.cpp:27-47 and camModel.hpp:25-33."""
    c = pkg.default_config()
    assert (c.sigma_vx, c.sigma_wz, c.window_size, c.sigma_pixel, c.rho_0, c.sigma_rho_0) == (0.01, 0.01, 21, 2, 0.1, 0.25)
    assert (c.scale, c.T_camera, c.sigma_size, c.nInitFeatures, c.min_features, c.max_features, c.forsePlane) == (1, 0.5, 2, 5, 30, 100, 0)
    assert (c.ncc_threshold, c.search_clamp, c.ransac_p, c.ransac_nhyp0, c.hi_chi2_threshold) == (0.8, 20.0, 0.99, 10000, 1.0)
    L = C.CDLL(pkg.build())
    d = pkg._abi.EkfConfig()
    L.ekf_config_default.argtypes = [C.POINTER(pkg._abi.EkfConfig)]
    L.ekf_config_default(C.byref(d))
    for f, _ in pkg._abi.EkfConfig._fields_:
        assert getattr(c, f) == getattr(d, f), f


def test_product_has_no_oracle_or_cpu_fallback():
    """The product package must not import, link or call anything under oracle/."""
    pdir = os.path.join(ROOT, "ekf-monoslam_for_3d-reconstruction_b200")
    for dirpath, _, files in os.walk(pdir):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import orc" not in txt and "liborc" not in txt and "ekf_oracle" not in txt, os.path.join(dirpath, f)


def test_missing_library_fails_loudly(pkg, monkeypatch):
    lib_mod = sys.modules["ekf_b200._lib"] if "ekf_b200._lib" in sys.modules else __import__("importlib").import_module("ekf_b200._lib")
    monkeypatch.setattr(lib_mod, "_lib", None)
    monkeypatch.setattr(lib_mod, "LIB_PATH", "/nonexistent/libekf_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lib_mod.lib()
