"""The oracle against a SECOND independent witness (tests/spec_decoder.py: numpy float64 from the Vorbis I specification's
formulas, own header parser, direct O(N^2) inverse MDCT) on the same boundary records -- in particular for the territory
no reference fixture reaches (floor 0, residue 0, lookup type 2, sequence_p, 3 / 6 channels, other block sizes), where the
C++ oracle used to be the only witness ("pin parity harder").  CPU only; the GPU twin is in
test_synthetic_setups.py::test_gpu_matches_the_spec_decoder."""
import numpy as np
import pytest

import helpers as H
import spec_decoder as SD
import vorbis_headers as VH
from nvorbis_b200 import hostlib
from oracle import oracle as O
from test_synthetic_setups import SETUPS

# setups with block sizes >= 256 (the reference's transform is not an IMDCT below that)
SPEC_SETUPS = ["stereo_r2", "six_ch_r2_coupled", "three_ch_r0", "stereo_r1", "stereo_512_1024", "mono_r1_dims_1_16",
               "stereo_r2_dims_1_16", "stereo_r2_48_posts", "stereo_floor0", "twelve_ch_r2_40_steps", "nine_ch_r1"]


def spec_case(name, n_frames=10, seed=99):
    d, s, g, f = VH.build_stream(**SETUPS[name])
    reader = O.OracleReader(O.PacketList(d, s, g, f))
    host = hostlib.HostStream(packets=(d, s, g, f))
    desc = H.desc_from_oracle(reader)
    hb = VH.random_records(np.random.default_rng(seed), desc, n_frames, host.post_stride, silent_prob=0.15, floor0_stride=host.floor0_stride)
    hdr = SD.parse_headers(d, s)
    got = SD.synth(hdr, hb, host.post_stride, host.floor0_stride)
    peak = float(np.abs(SD.synth(hdr, hb, host.post_stride, host.floor0_stride, clip=False)).max())
    return reader, host, desc, hb, got, max(1.0, peak)


def test_db_table_formula_matches_the_printed_constants():
    import os, re
    txt = open(os.path.join(H.ROOT, "oracle", "inverse_db_table.inc")).read()
    printed = np.array([int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{8})u", txt)], np.uint32).view(np.float32).astype(np.float64)
    tab = SD.inverse_db_table()
    assert printed.size == 256 and tab[255] == 1.0
    assert np.abs(tab / printed - 1.0).max() < 2e-7


@pytest.mark.parametrize("name", SPEC_SETUPS)
def test_oracle_agrees_with_the_spec_decoder(name):
    reader, host, desc, hb, got, scale = spec_case(name)
    want, _ = H.oracle_decode_records(reader, hb, desc)
    assert want.size == got.size and want.size > 0
    # float32 chain against float64: 1e-5 of the signal scale (pre-clip peaks of the synthetic records reach several units)
    assert float(np.abs(got - want).max()) <= 1e-5 * scale


@pytest.mark.parametrize("name,hi", [("1test", 24), ("3test", 40)])
def test_fixture_prefix_agrees_with_the_spec_decoder(name, hi):
    """Real packets too: the first blocks of two reference test files, from the oracle's boundary records."""
    r, pcm, b = H.decoded(name)
    pk = H.packets(name)
    host = hostlib.HostStream(packets=(pk.data, pk.sizes, pk.granules, pk.flags))
    hb = H.batch_from_boundary(b, host.post_stride, 0, hi)
    want, _ = H.oracle_synth(r, b, 0, hi)
    got = SD.synth(SD.parse_headers(pk.data, pk.sizes), hb, host.post_stride)
    assert got.size == want.size and float(np.abs(got - want).max()) <= 1e-5
