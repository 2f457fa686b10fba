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


def test_run_flags_match_header():
    # the nvb_run flag values of the header against the Python binding's constants (NVB_RUN_ONE_KERNEL / _TWO_KERNELS are round 2's)
    import subprocess, tempfile
    src = r'''
#include <stdio.h>
#include "nvorbis_b200.h"
int main(void) { printf("%d %d %d %d %d %d %d %d\n", NVB_RUN_DEFAULT, NVB_RUN_EXACT, NVB_RUN_NO_CLIP, NVB_RUN_CONTINUE, NVB_RUN_PCM_S16, NVB_RUN_DEVICE_OUT,
  NVB_RUN_ONE_KERNEL, NVB_RUN_TWO_KERNELS); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        vals = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert vals == [capi.RUN_DEFAULT, capi.RUN_EXACT, capi.RUN_NO_CLIP, capi.RUN_CONTINUE, capi.RUN_PCM_S16, capi.RUN_DEVICE_OUT, capi.RUN_ONE_KERNEL,
                    capi.RUN_TWO_KERNELS]


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


def _fnv1a64(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_blob_import_rejects_tampered_blobs():
    """nvb_setup_blob_import must establish the invariants the kernels rely on -- a damaged body fails the
    checksum, a section offset near 2^64 does not wrap past the range check, and a blob whose derived kernel-selection
    fields were edited (checksum fixed up) is refused because the import recomputes them.  Uses the CPU shim build of the
    library (nvb_setup_blob_import itself is CUDA-free up to the final upload)."""
    import struct
    import helpers as H
    from nvorbis_b200 import capi
    shim = H.build_shim()
    r, pcm, b = H.decoded("3test")
    ctx = capi.Context(0, lib_path=shim)
    ctx.upload_setup(H.setup_from_oracle(r))
    blob = ctx.export_blob()
    HASH_AT, OFF_VQ_AT, OFF_RES_AT, HDR = 352, 72, 88, 384
    assert struct.unpack_from("<Q", blob, HASH_AT)[0] == _fnv1a64(blob[HDR:].tobytes())      # the layout this test assumes

    def refused(buf):
        c2 = capi.Context(0, lib_path=shim)
        with pytest.raises(capi.NvbError) as e:
            c2.import_blob(buf)
        assert e.value.status == capi.ERR_DATA
        c2.close()

    bad = blob.copy(); bad[HDR + 1000] ^= 1                       # body damage
    refused(bad)
    bad = blob.copy(); struct.pack_into("<Q", bad, OFF_VQ_AT, 2 ** 64 - 16)      # off + len would wrap
    refused(bad)
    off_res = struct.unpack_from("<Q", blob, OFF_RES_AT)[0]
    bad = blob.copy()
    assert struct.unpack_from("<i", bad, off_res + 28)[0] == 3   # DevResidue.fast of 3test's first residue
    struct.pack_into("<i", bad, off_res + 24, -1)                # pshift = -1 with the run kernels still selected
    struct.pack_into("<Q", bad, HASH_AT, _fnv1a64(bad[HDR:].tobytes()))
    refused(bad)
    ok = capi.Context(0, lib_path=shim); ok.import_blob(blob); ok.close()
    ctx.close()
