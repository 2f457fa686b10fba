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
    "oct_r2": dict(channels=8, bs0=256, bs1=1024, residue_type=2, coupling=[(0, 1), (2, 3), (6, 7), (0, 4)]),
    "tiny_blocks_r1_lookup2_seq": dict(channels=2, bs0=64, bs1=128, residue_type=1, coupling=[(0, 1)], lookup=2, sequence_p=True),
    "three_ch_r0": dict(channels=3, bs0=512, bs1=4096, residue_type=0, coupling=[(1, 2)]),
    "mono_r1_big": dict(channels=1, bs0=1024, bs1=8192, residue_type=1),
    "stereo_r1": dict(channels=2, bs0=256, bs1=2048, residue_type=1, coupling=[(1, 0)], sequence_p=True),
    "stereo_512_1024": dict(channels=2, bs0=512, bs1=1024, residue_type=2, coupling=[(0, 1)]),
    "stereo_1024_2048": dict(channels=2, bs0=1024, bs1=2048, residue_type=2, coupling=[(0, 1)]),
    "stereo_512_4096": dict(channels=2, bs0=512, bs1=4096, residue_type=2, coupling=[(0, 1)]),
    "stereo_r2_dims_1_16": dict(channels=2, bs0=256, bs1=2048, residue_type=2, coupling=[(0, 1)], lookup=2, res_dims=(1, 16, 8)),
    "mono_r1_dims_1_16": dict(channels=1, bs0=256, bs1=2048, residue_type=1, lookup=2, res_dims=(16, 1, 4)),
    "stereo_r2_48_posts": dict(channels=2, bs0=256, bs1=2048, residue_type=2, coupling=[(0, 1)], floor_posts=46),
    "stereo_floor0": dict(channels=2, bs0=256, bs1=2048, residue_type=2, coupling=[(0, 1)], floor_type=0),
    "quad_floor0_r1": dict(channels=4, bs0=128, bs1=1024, residue_type=1, coupling=[(0, 1), (2, 3)], floor_type=0),
    # more than 8 channels (round 2: NVB_MAX_CHANNELS 32, NVB_MAX_COUPLING 256; the reference takes up to 255 channels, StreamDecoder.cs:186,
    # and 256 coupling steps, Mapping.cs:28): the general spectrum kernel's many-channel instantiation; even counts take the fused
    # kernel's channel-pair units, odd ones whole-frame units while three slots fit, the rest the exact kernels
    "twelve_ch_r2_40_steps": dict(channels=12, bs0=256, bs1=2048, residue_type=2,
                                  coupling=[(i % 12, (i * 5 + 1 + (i // 12)) % 12) for i in range(40) if i % 12 != (i * 5 + 1 + (i // 12)) % 12]),
    "nine_ch_r1": dict(channels=9, bs0=256, bs1=2048, residue_type=1, coupling=[(0, 1), (8, 2), (3, 7)]),
    "thirty_two_ch_r1": dict(channels=32, bs0=256, bs1=1024, residue_type=1, coupling=[(2 * i, 2 * i + 1) for i in range(16)]),
    "twelve_ch_floor0_r1": dict(channels=12, bs0=128, bs1=1024, residue_type=1, coupling=[(0, 1), (2, 3), (11, 4)], floor_type=0),
}
# type 0 floors end in exp(), sqrt() and cos() of the platform's math library (System.Math in the reference, libm in the
# oracle, CUDA's double-precision functions on the GPU): results agree to the last float bit almost everywhere, not
# everywhere -- the "exact" path is held to a relative 1e-6 there instead of bit equality
FLOOR0 = {"stereo_floor0", "quad_floor0_r1", "twelve_ch_floor0_r1"}


def _run(name, n_frames, lib_path, seed=1234):
    d, s, g, f = VH.build_stream(**SETUPS[name])
    reader = O.OracleReader(O.PacketList(d, s, g, f))
    host = hostlib.HostStream(packets=(d, s, g, f))
    desc = H.desc_from_oracle(reader)
    hb = VH.random_records(np.random.default_rng(seed), desc, n_frames, host.post_stride, floor0_stride=host.floor0_stride)
    want, clipped = H.oracle_decode_records(reader, hb, desc)
    ctx = capi.Context(0, lib_path=lib_path)
    ctx.upload_setup(host.setup())                       # the product's own header parser feeds the setup
    assert ctx.floor0_stride == host.floor0_stride
    out, res = ctx.decode_batch(hb, capi.RUN_EXACT)
    if name in FLOOR0:
        assert out.size == want.size and float(np.abs(out - want).max()) <= 1e-6 * max(1.0, float(np.abs(want).max()))
    else:
        np.testing.assert_array_equal(out, want)         # stb dataflow, no FMA: bit-identical for every block size
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


def _floor0_stream(name, n_frames, seed):
    kw = SETUPS[name]
    d, s, g, f = VH.build_stream(**kw)
    pk = VH.floor0_audio_packets(np.random.default_rng(seed), kw["channels"], kw["bs0"], kw["bs1"], n_frames, kw["residue_type"])
    data = np.concatenate([d, np.frombuffer(b"".join(pk), np.uint8)])
    sizes = np.concatenate([s, np.array([len(p) for p in pk], np.int64)])
    return data, sizes, np.zeros(len(sizes), np.int64), np.zeros(len(sizes), np.uint8)


def _run_floor0_packets(name, n_frames, lib_path, seed=7, gpu_unpack=False):
    """Floor 0 end to end from PACKETS: the oracle's full decode (Floor0.Unpack + Apply restated) against the product's
    host unpacker + synthesis behind the VorbisReader mirror."""
    from nvorbis_b200.reader import VorbisReader
    d, s, g, f = _floor0_stream(name, n_frames, seed)
    want = O.OracleReader(O.PacketList(d, s, g, f)).read_all()
    assert want.size > 0 and np.isfinite(want).all() and float(np.abs(want).max()) > 1e-3
    with VorbisReader((d, s, g, f), batch_packets=11, lib_path=lib_path, gpu_unpack=gpu_unpack) as vr:
        got = vr.read_all(chunk_seconds=0.25)
    assert got.size == want.size and float(np.abs(got - want).max()) <= 1e-5


@pytest.mark.parametrize("name", sorted(FLOOR0))
def test_floor0_packets_on_cpu_shim(name):
    _run_floor0_packets(name, 14, H.build_shim())


@pytest.mark.parametrize("name", sorted(FLOOR0))
def test_floor0_packets_device_unpack_on_cpu_shim(name):
    """Floor0.Unpack on the device (k_unpack, round 2): amplitude, book number, LSP coefficients out of the VQ tables, running sum."""
    _run_floor0_packets(name, 14, H.build_shim(), gpu_unpack=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FLOOR0))
def test_floor0_packets_on_gpu(name):
    _run_floor0_packets(name, 400, None)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FLOOR0))
def test_floor0_packets_device_unpack_on_gpu(name):
    _run_floor0_packets(name, 400, None, gpu_unpack=True)


def test_unaligned_type2_residues_run_on_the_bins_kernel():
    """6 channels with 32-wide partitions (BASELINE configs[3]; Residue2.cs:25-27 truncation case) and 3 / 5 channels are
    covered by k_spectrum_bins: with the general kernel forbidden the decode still works and stays bit-identical."""
    import os, subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import test_synthetic_setups as T, helpers as H\n"
            "T.SETUPS['five_ch_r2'] = dict(channels=5, bs0=256, bs1=2048, residue_type=2, coupling=[(0, 1), (3, 4)])\n"
            "T.SETUPS['three_ch_r2'] = dict(channels=3, bs0=128, bs1=512, residue_type=2, coupling=[(2, 0)])\n"
            "T.SETUPS['six_ch_40_posts'] = dict(channels=6, bs0=256, bs1=1024, residue_type=2, coupling=[(0, 1)], floor_posts=38)\n"
            "for name in ('six_ch_r2_coupled', 'five_ch_r2', 'three_ch_r2', 'six_ch_40_posts'):\n    T._run(name, 8, H.build_shim())\nprint('ok')\n") % (H.ROOT, os.path.join(H.ROOT, "tests"))
    env = dict(os.environ, NVB_SPECTRUM_FORBID_GENERIC="1")
    assert subprocess.check_output([sys.executable, "-c", code], env=env, timeout=600).decode().strip().endswith("ok")


def test_block_sizes_from_256_up_run_on_the_fused_path():
    """Every pair of block sizes >= 256 decodes with two kernels (spectrum + fused IMDCT/window/OLA: k_imdct_fused for
    256/2048, k_imdct_generic otherwise); sizes below 256, where the reference's Mdct is not an IMDCT, take the exact
    kernels (three launches)."""
    shim = H.build_shim()
    for name, want in (("stereo_r2", 2), ("stereo_512_1024", 2), ("stereo_512_4096", 2), ("mono_r1_big", 2), ("three_ch_r0", 2),
                       ("tiny_blocks_r1_lookup2_seq", 3), ("quad_floor0_r1", 3)):
        d, s_, g, f = VH.build_stream(**SETUPS[name])
        host = hostlib.HostStream(packets=(d, s_, g, f))
        desc = H.desc_from_oracle(O.OracleReader(O.PacketList(d, s_, g, f)))
        hb = VH.random_records(np.random.default_rng(3), desc, 6, host.post_stride, floor0_stride=host.floor0_stride)
        ctx = capi.Context(0, lib_path=shim)
        ctx.upload_setup(host.setup())
        db = ctx.create_dbatch(hb, capi.RUN_TWO_KERNELS)                      # (small mono / stereo 256/2048 batches default to ONE kernel: test_gpu_one_kernel.py)
        pcm = np.zeros(db.samples * desc["channels"] + 16, np.float32)
        db.run(pcm.ctypes.data, 0)
        assert db.launches == want, (name, db.launches)
        db.destroy(); ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["stereo_r2", "six_ch_r2_coupled", "three_ch_r0", "stereo_r1", "stereo_r2_dims_1_16", "stereo_floor0", "twelve_ch_r2_40_steps",
                                  "nine_ch_r1"])
def test_gpu_matches_the_spec_decoder(name):
    """The GPU path against the second independent witness (tests/spec_decoder.py: float64, from the Vorbis I specification's
    formulas) -- for floor 0, residue 0, lookup type 2, sequence_p and six channels the oracle is not the only reference."""
    from test_spec_decoder import spec_case
    reader, host, desc, hb, got, scale = spec_case(name)
    ctx = capi.Context(0)
    ctx.upload_setup(host.setup())
    for flags in (capi.RUN_EXACT, capi.RUN_DEFAULT):
        ctx.reset()
        out, _ = ctx.decode_batch(hb, flags)
        assert out.size == got.size and float(np.abs(out - got).max()) <= 1e-5 * scale
    ctx.close()
