"""CPU tests of the C-ABI boundary: the shared library builds, loads, and exports exactly the
entry points include/fmm_b200.h (+ the measurement aids of include/fmm_b200_debug.h) declare (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(headers=("fmm_b200.h", "fmm_b200_debug.h")):
    names = set()
    for h in headers:
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(fmm_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_product_header_has_no_debug_entry_points():
    assert not [n for n in declared_symbols(("fmm_b200.h",)) if n.startswith("fmm_debug_")]
    assert sorted(declared_symbols(("fmm_b200_debug.h",))) == ["fmm_debug_mma_probe", "fmm_debug_wait_profile"]


@pytest.fixture(scope="module")
def lib():
    from fall_multimodal_b200 import build

    path = build.build()
    return ctypes.CDLL(str(path))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 26
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_binding_table_matches_header():
    from fall_multimodal_b200 import _lib

    bound = set(_lib._SIGNATURES) | {"fmm_last_error"}
    assert bound == set(declared_symbols())


def test_version_and_error_text(lib):
    lib.fmm_last_error.restype = ctypes.c_char_p
    assert lib.fmm_version() >= 100
    assert isinstance(lib.fmm_last_error(), bytes)


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    """Argument validation happens on the host before any launch: NULL pointers -> status -1."""
    lib.fmm_last_error.restype = ctypes.c_char_p
    st = lib.fmm_colstats(None, None, None, None, 1, 1, 1, 8, 0, None)
    assert st == -1 and b"colstats" in lib.fmm_last_error()
    st = lib.fmm_tapconv_pack(None, None, 64, 64, 64, 64, 0, 64, 0, 1, 0, 1, None, 0, None)
    assert st == -1


def test_product_path_has_no_cpu_fallback():
    import torch

    import fall_multimodal_b200 as fmm

    m = fmm.STGCAN(3, {"layout": "coco_cut", "strategy": "spatial"}, num_class=11)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 3, 8, 14), None)
    c = fmm.CNN1D()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c(torch.zeros(2, 15, 30))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fall_multimodal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower().replace("# oracle", ""), f


def test_descriptor_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors of the POD descriptors (bgemm, cell_fwd, cell_bwd) have the C header's size and field offsets."""
    import ctypes
    import shutil
    import subprocess

    from fall_multimodal_b200 import _lib

    gcc = shutil.which("gcc")
    cuda_inc = "/usr/local/cuda/include"
    if gcc is None or not os.path.isdir(cuda_inc):
        pytest.skip("needs gcc and the CUDA headers")
    checks = {"fmm_bgemm_desc": (_lib.BgemmDesc, ["A", "bias_n", "a_g1", "b_k3", "c_n", "G1", "K3", "alpha", "beta", "c_dtype"]),
              "fmm_cell_fwd_args": (_lib.CellFwdArgs, ["x", "hv", "S", "lu", "hout", "ov", "xc1", "mode", "Cp"]),
              "fmm_cell_bwd_args": (_lib.CellBwdArgs, ["S", "dz", "dx", "dxv", "lg", "dH", "z1", "hv1", "dlin_u", "mode", "Cp"]),
              "fmm_gruscan_args": (_lib.GruScanArgs, ["xb", "hout", "bl", "dhout", "dh_b", "dh_v", "dxu", "LT", "err", "B", "KS", "xb_slot0",
                                                      "tsplit"]),
              "fmm_tattn_args": (_lib.TAttnArgs, ["q", "v", "out", "lse", "dout", "dv", "B", "Tp", "F", "scale", "v_btvc"]),
              "fmm_head_args": (_lib.HeadArgs, ["feat", "dfeat", "width", "nseg", "W", "target", "prob2", "gloss", "dbias", "N", "F",
                                                "pre_softmax", "smoothing"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fmm_b200.h"', "int main(void) {"]
    for cname, (_, fields) in checks.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname} {f} %zu\\n", offsetof({cname}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi_probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi_probe"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([gcc, "-I", os.path.join(root, "include"), "-I", cuda_inc, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    for line in filter(None, out):
        cname, what, val = line.split()
        struct = checks[cname][0]
        if what == "size":
            assert ctypes.sizeof(struct) == int(val), (cname, ctypes.sizeof(struct), val)
        else:
            assert getattr(struct, what).offset == int(val), (cname, what)
