"""The C-ABI shared library: exports and failure behaviour without a GPU (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from nvorbis_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    import __graft_entry__ as g
    if not os.path.exists(capi.DEFAULT_LIB):
        g.build()
    return capi.load_library()


def test_header_symbols_exported():
    lib = _built()
    hdr = open(os.path.join(ROOT, "include", "nvorbis_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(nvb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/nvorbis_b200.h but not exported"
    assert sorted(capi.EXPORTS) == declared
    assert lib.nvb_abi_version() == capi.ABI_VERSION
    assert lib.nvb_strerror(capi.ERR_DATA) == b"malformed data"


def test_struct_sizes_match_header():
    # sizes the C compiler gives the structs of include/nvorbis_b200.h (checked with a tiny C program)
    import subprocess, tempfile
    src = r'''
#include <stdio.h>
#include "nvorbis_b200.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(nvb_codebook), sizeof(nvb_floor1), sizeof(nvb_floor), sizeof(nvb_residue),
  sizeof(nvb_mapping), sizeof(nvb_mode), sizeof(nvb_setup), sizeof(nvb_frame), sizeof(nvb_batch), sizeof(nvb_result)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    want = [C.sizeof(capi.Codebook), C.sizeof(capi.Floor1), C.sizeof(capi.Floor), C.sizeof(capi.Residue), C.sizeof(capi.Mapping),
            C.sizeof(capi.Mode), C.sizeof(capi.SetupStruct), capi.FRAME_DTYPE.itemsize, C.sizeof(capi.BatchStruct), C.sizeof(capi.ResultStruct)]
    assert sizes == want


def test_no_cpu_fallback():
    """Without a GPU the product refuses to run; it never falls back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _built()
    with pytest.raises(capi.NvbError) as e:
        capi.Context(0)
    assert e.value.status == capi.ERR_CUDA


def test_product_does_not_touch_oracle():
    pkg = os.path.join(ROOT, "nvorbis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("import oracle", "from oracle", "oracle/", "libnvorbis_oracle", "orc_"):
                    assert needle not in txt, f"{f} references the oracle ({needle})"
