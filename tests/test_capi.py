"""CPU: the C-ABI library loads and exports every symbol include/sbr_b200.h declares; the ctypes
binding table covers exactly that set; creating a model without a GPU fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sbr_b200.h")).read()
    return sorted(set(re.findall(r"SBR_API\s+[\w\s\*]+?\b(sbr_\w+)\s*\(", src)))


def test_build_entry_point_and_symbols():
    import __graft_entry__ as g
    g.build()
    from sbr_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 28
    for s in syms:
        assert hasattr(lib, s), "library does not export " + s
    assert sorted(_capi.SIGNATURES) == syms, "ctypes table and header disagree"
    assert lib.sbr_abi_version() == 2


def test_config_struct_matches_header_layout(tmp_path):
    """sizeof / offsetof of sbr_config as gcc sees the header == the ctypes mirror."""
    import subprocess
    from sbr_b200 import _capi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sbr_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu\\n", sizeof(sbr_config), offsetof(sbr_config, lr),'
                   ' offsetof(sbr_config, n_slots), offsetof(sbr_config, nccl_id));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, o_lr, o_slots, o_id = map(int, subprocess.check_output([str(exe)]).split())
    assert ctypes.sizeof(_capi.SbrConfig) == size
    assert _capi.SbrConfig.lr.offset == o_lr
    assert _capi.SbrConfig.n_slots.offset == o_slots
    assert _capi.SbrConfig.nccl_id.offset == o_id


def test_no_cpu_fallback():
    from sbr_b200 import _capi
    lib = _capi.load_library()
    if lib.sbr_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(_capi.SbrError) as e:
        _capi.Engine(n_items=10)
    assert e.value.code == -6 and "no CPU fallback" in str(e.value)


def test_bad_config_is_rejected_before_touching_the_device():
    from sbr_b200 import _capi
    lib = _capi.load_library()
    cfg = _capi.SbrConfig()
    cfg.struct_size = 12  # ABI guard
    h = ctypes.c_void_p()
    assert lib.sbr_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"struct_size" in lib.sbr_last_error(None)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sequence-based-recommendations_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("numpy oracle", "").lower() or f == "build.py", \
                    "%s mentions the oracle" % os.path.join(dirpath, f)
    for f in ("train.py", "test.py"):
        assert "oracle" not in open(os.path.join(ROOT, f)).read()
