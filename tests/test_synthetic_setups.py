"""Synthetic stream setups the reference's fixtures never reach -- every block size 64..8192 (the exact kernels follow
Mdct.cs for all of them, including its N < 256 behaviour), residue types 0 / 1 / 2, codebook lookup type 2 and
sequence_p, three coupling steps over six channels (channel 0 in two steps: BASELINE configs[3]) -- decoded by the
product and by the oracle from the same seeded boundary records.  "Parity unpinned" territory: no reference output
exists for these, the oracle is the only witness.  CPU: tests/cpu_shim; GPU: the real library."""
import numpy as np
import pytest

import helpers as H
import vorbis_headers as VH
from nvorbis_b200 import capi, hostlib
from oracle import oracle as O

SETUPS = {
    "stereo_r2": dict(channels=2, bs0=256, bs1=2048, residue_type=2, coupling=[(0, 1)]),
    "six_ch_r2_coupled": dict(channels=6, bs0=256, bs1=2048, residue_type=2, coupling=[(0, 2), (3, 4), (0, 1)]),
    "quad_r2": dict(channels=4, bs0=256, bs1=2048, residue_type=2, coupling=[(0, 1), (2, 3)]),
    "tiny_blocks_r1_lookup2_seq": dict(channels=2, bs0=64, bs1=128, residue_type=1, coupling=[(0, 1)], lookup=2, sequence_p=True),
    "three_ch_r0": dict(channels=3, bs0=512, bs1=4096, residue_type=0, coupling=[(1, 2)]),
    "mono_r1_big": dict(channels=1, bs0=1024, bs1=8192, residue_type=1),
    "stereo_r1": dict(channels=2, bs0=256, bs1=2048, residue_type=1, coupling=[(1, 0)], sequence_p=True),
}


def _run(name, n_frames, lib_path, seed=1234):
    d, s, g, f = VH.build_stream(**SETUPS[name])
    reader = O.OracleReader(O.PacketList(d, s, g, f))
    host = hostlib.HostStream(packets=(d, s, g, f))
    desc = H.desc_from_oracle(reader)
    hb = VH.random_records(np.random.default_rng(seed), desc, n_frames, host.post_stride)
    want, clipped = H.oracle_decode_records(reader, hb, desc)
    ctx = capi.Context(0, lib_path=lib_path)
    ctx.upload_setup(host.setup())                       # the product's own header parser feeds the setup
    out, res = ctx.decode_batch(hb, capi.RUN_EXACT)
    np.testing.assert_array_equal(out, want)             # stb dataflow, no FMA: bit-identical for every block size
    assert res.has_clipped == clipped and res.n_floor_range == 0
    ctx.reset()
    out, res = ctx.decode_batch(hb, capi.RUN_DEFAULT)    # fused path for {256, 2048}, exact kernels otherwise
    assert out.size == want.size and float(np.abs(out - want).max()) <= 1e-5
    assert res.has_clipped == clipped
    ctx.close()


@pytest.mark.parametrize("name", list(SETUPS))
def test_synthetic_setup_on_cpu_shim(name):
    n = 10 if SETUPS[name]["bs1"] >= 4096 else 18
    _run(name, n, H.build_shim())


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SETUPS))
def test_synthetic_setup_on_gpu(name):
    _run(name, 600, None, seed=99)


@pytest.mark.gpu
def test_config4_six_channel_8k_frames():
    """BASELINE configs[3]: 6-channel mapping with three coupling steps, N = 2048, 8192 frames."""
    d, s, g, f = VH.build_stream(**SETUPS["six_ch_r2_coupled"])
    reader = O.OracleReader(O.PacketList(d, s, g, f))
    host = hostlib.HostStream(packets=(d, s, g, f))
    desc = H.desc_from_oracle(reader)
    hb = VH.random_records(np.random.default_rng(20240004), desc, 8192, host.post_stride, short_prob=0.0)
    import os
    want, clipped = H.oracle_decode_records(reader, hb, desc, threads=os.cpu_count() or 1)
    ctx = capi.Context(0); ctx.upload_setup(host.setup())
    out, res = ctx.decode_batch(hb)
    assert out.size == want.size == 8191 * 1024 * 6
    assert float(np.abs(out - want).max()) <= 1e-5 and res.has_clipped == clipped
