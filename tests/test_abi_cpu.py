"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/vcg.h declares; error plumbing works; no compute is launched."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "vcg.h")).read()
    return sorted(set(re.findall(r"VCG_API\s+[\w\s\*]+?\b(vcg_\w+)\s*\(", txt)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("vcg_conv_fwd", "vcg_conv_wgrad", "vcg_xform_fwd", "vcg_xform_bwd_gather", "vcg_xform_bwd_norm",
                 "vcg_reparam_fwd", "vcg_reparam_bwd", "vcg_l1_fwd_bwd", "vcg_dhead_fwd", "vcg_dhead_bwd",
                 "vcg_adam_multi", "vcg_version", "vcg_last_error"):
        assert must in syms
    assert len(syms) >= 24


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md is the maintainer-facing map of the boundary: every declared entry point appears in it"""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in declared_symbols() if s not in doc]
    assert not missing, missing


def test_library_exports_every_declared_symbol(vcg):
    from vcg_b200 import lib
    out = subprocess.run(["nm", "-D", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (vcg_\w+)", out))
    declared = set(declared_symbols())
    assert declared <= exported, declared - exported
    assert set(lib.EXPORTS) == declared, set(lib.EXPORTS) ^ declared
    l = lib.load()
    assert l.vcg_version() == 2
    assert l.vcg_launch_count() == 0
    assert isinstance(l.vcg_last_error(), bytes)


def test_struct_layouts_match_header(vcg):
    """ctypes mirrors must have the header's field counts (all int32 unless noted)."""
    import ctypes as C
    from vcg_b200 import lib
    txt = open(os.path.join(ROOT, "include", "vcg.h")).read()

    def nfields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), txt, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        n = 0
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if stmt:
                n += stmt.count(",") + 1
        return n
    assert nfields("vcg_conv_desc") == len(lib.ConvDesc._fields_) == 15
    assert nfields("vcg_wpack_desc") == len(lib.WpackDesc._fields_)
    assert nfields("vcg_xform_desc") == len(lib.XformDesc._fields_)
    assert nfields("vcg_xbwd_desc") == len(lib.XbwdDesc._fields_)
    assert C.sizeof(lib.AdamChunk) == 40 and C.sizeof(lib.GSrc) == 24


def test_sass_is_blackwell_native(vcg):
    """tcgen05 / TMA / TMEM must be in the shipped SASS (UTCHMMA, UTMALDG, LDTM), no legacy HMMA."""
    from vcg_b200 import lib
    r = subprocess.run(["cuobjdump", "-sass", lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        import pytest
        pytest.skip("cuobjdump unavailable")
    sass = r.stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HGMMA" not in sass
    # the warp-level HMMA path is allowed only in the degenerate M=3 weight-gradient kernel (wgrad_thin.cu);
    # every GEMM-shaped convolution must be on tcgen05 (UTCHMMA)
    fn = None
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif " HMMA" in line:
            assert fn is not None and "wgrad_thin" in fn, f"legacy HMMA in {fn}"
    for must in ("conv_tc_kernel", "conv_tc2_kernel", "wgrad_tc_kernel", "wgrad_tc2_kernel", "conv_tc_fold_kernel", "wgrad_fold_kernel"):
        body = sass.split(must, 1)[1].split("Function :", 1)[0] if must in sass else ""
        assert "UTCHMMA" in body, must
