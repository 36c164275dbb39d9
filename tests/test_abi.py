"""CPU-only checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/*.h declares, its headers are valid plain C, the table structs have the reference's
sizes, and the entry points refuse to run (rather than fall back) when no CUDA device exists."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

import cpu_checkers as cc
import ref_tables as rt

ROOT = cc.ROOT
INCLUDE = os.path.join(ROOT, "include")


def declared_functions():
    names = set()
    for hdr in os.listdir(INCLUDE):
        text = open(os.path.join(INCLUDE, hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"^\s*(?:const\s+)?(?:struct\s+)?\w[\w\s\*]*?\b(x264(?:dsp)?_\w+)\s*\(", text, flags=re.M):
            names.add(m.group(1))
    return sorted(n for n in names if not n.endswith("_t"))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.lib()
    names = declared_functions()
    assert len(names) >= 40, names
    for must in ("x264_pixel_init", "x264_dct_init", "x264_zigzag_init", "x264_mc_init", "x264_quant_init",
                 "x264_deblock_init", "x264dsp_lookahead_frame_cost_dev", "x264dsp_me_search_batch_dev",
                 "x264dsp_residual_frame_dev", "x264dsp_deblock_frame_dev"):
        assert must in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_headers_are_plain_c_and_table_sizes_match(tmp_path):
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "x264dsp_b200.h"\n#include "x264dsp_tables.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(x264_pixel_function_t),'
                   'sizeof(x264_dct_function_t), sizeof(x264_zigzag_function_t), sizeof(x264_mc_functions_t),'
                   'sizeof(x264_quant_function_t), sizeof(x264_deblock_function_t), sizeof(x264dsp_me_block_t),'
                   'sizeof(x264dsp_me_result_t), sizeof(x264dsp_geom_t));return 0;}\n')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", INCLUDE, str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    for cls, s in zip(rt.TABLES, sizes):
        assert C.sizeof(cls) == s, cls.__name__
    assert sizes[6] == C.sizeof(cc.MeBlock) and sizes[7] == C.sizeof(cc.MeResult) and sizes[8] == C.sizeof(cc.Geom)
    if cc.ref() is not None:
        ref_sizes = (C.c_int * 8)()
        cc.ref().xref_table_sizes(ref_sizes)
        assert list(ref_sizes)[:6] == sizes[:6], "table layouts differ from the reference build"


def test_no_cpu_fallback_without_a_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present; the refusal path is exercised on the CPU box")
    ctx = C.c_void_p()
    rc = pkg.lib().x264dsp_create(0, C.byref(ctx))
    assert rc == -2 and not ctx.value, "x264dsp_create must fail with X264DSP_E_NOGPU, not fall back"
    with pytest.raises(pkg.X264DspError):
        pkg.Context(0)
    # host-side table helpers still work (they are tables, not compute)
    assert pkg.lib().x264dsp_lambda(26) == 5


def test_product_does_not_reference_the_oracle():
    """nothing under the package may import, link or dlopen oracle/"""
    pkgdir = os.path.join(ROOT, "x264-dsp_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text and "libx264ref" not in text and "xo_" not in text, os.path.join(dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkgdir, "libx264dsp_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "x264ref" not in out
