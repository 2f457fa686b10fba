"""Parity of the CUDA synthesis path (through the C ABI of libnvorbis_b200.so) with the CPU oracle.

Bar: bit-identical for the exact path (stb dataflow, no FMA contraction); max-abs <= 1e-5 (BASELINE.json
north_star) for the fused fast path.  Run on a B200: python -m pytest tests -m gpu
"""
import os

import numpy as np
import pytest

import helpers as H
from nvorbis_b200 import capi, setupio, workloads

pytestmark = pytest.mark.gpu
TOL = 1e-5          # north_star: "output floats must match the reference C# path ... to within 1e-5 max-abs"


def _ctx(name):
    r, pcm, b = H.decoded(name)
    ctx = capi.Context(0)
    ctx.upload_setup(H.setup_from_oracle(r))
    return r, pcm, b, ctx


def test_library_is_the_real_one():
    lib = capi.load_library()
    assert os.path.samefile(lib._name, capi.DEFAULT_LIB)


@pytest.mark.parametrize("name", H.FIXTURES)
def test_fixture_streams(name, golden):
    r, pcm, b, ctx = _ctx(name)
    hb = H.batch_from_boundary(b, ctx.post_stride)
    out, res = ctx.decode_batch(hb, capi.RUN_EXACT)
    np.testing.assert_array_equal(out, pcm)
    assert res.samples_per_channel == golden[name]["samples_per_channel"] and res.has_clipped == golden[name]["has_clipped"]
    ctx.reset()
    out, res = ctx.decode_batch(hb, capi.RUN_DEFAULT)
    assert out.size == pcm.size
    assert float(np.abs(out - pcm).max()) <= TOL
    assert res.has_clipped == golden[name]["has_clipped"] and res.n_floor_range == 0 and res.n_inconsistent == 0
    ctx.close()


@pytest.mark.parametrize("name", ["1test", "3test"])
def test_unclipped_output(name):
    r, _, b = H.decoded(name)
    want = H.O.OracleReader(H.packets(name), clip=False).read_all()
    ctx = capi.Context(0); ctx.upload_setup(H.setup_from_oracle(r))
    hb = H.batch_from_boundary(b, ctx.post_stride)
    out, res = ctx.decode_batch(hb, capi.RUN_EXACT | capi.RUN_NO_CLIP)
    np.testing.assert_array_equal(out, want)
    ctx.reset()
    out, res = ctx.decode_batch(hb, capi.RUN_NO_CLIP)
    assert float(np.abs(out - want).max()) <= TOL and not res.has_clipped


@pytest.mark.parametrize("flags", [capi.RUN_EXACT, capi.RUN_DEFAULT])
def test_chained_batches(flags):
    """Batches of ragged sizes chained with NVB_RUN_CONTINUE equal one big batch (the tail is carried on the device)."""
    r, pcm, b, ctx = _ctx("3test")
    n = len(b.frames)
    cuts = [1, 2, 9, 10, 64, 65, 200, 333, n]
    parts, pos = [], 0
    for cut in cuts:
        out, _ = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride, pos, cut), flags | (capi.RUN_CONTINUE if pos else 0))
        parts.append(out.copy()); pos = cut
    got = np.concatenate(parts)
    assert got.size == pcm.size
    if flags == capi.RUN_EXACT:
        np.testing.assert_array_equal(got, pcm)
    else:
        assert float(np.abs(got - pcm).max()) <= TOL


def test_empty_and_single_frame_batches():
    r, pcm, b, ctx = _ctx("1test")
    empty = capi.HostBatch(np.zeros(0, capi.FRAME_DTYPE), np.zeros(0, np.int16), np.zeros(0, np.uint8), np.zeros(0, np.uint16))
    out, res = ctx.decode_batch(empty)
    assert out.size == 0 and res.samples_per_channel == 0
    out, res = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride, 0, 1))
    assert out.size == 0                                    # first block of a stream emits nothing (StreamDecoder.cs:446-450)


@pytest.mark.parametrize("flags", [capi.RUN_EXACT, capi.RUN_DEFAULT])
def test_failed_packets_drain(flags):
    r, pcm, b, ctx = _ctx("3test")
    fr = b.frames.copy()
    for i in (0, 7, 8, 100, 101, 102, 250, len(fr) - 1):
        fr["ok"][i] = 0
    b2 = H.O.Boundary(b.channels, fr, b.block_size, b.valid_untrimmed, b.no_exec_mask, b.posts, b.post_counts, b.classes, b.entries)
    want, _ = H.oracle_synth(r, b2)
    out, res = ctx.decode_batch(H.batch_from_boundary(b2, ctx.post_stride), flags)
    assert res.n_failed == 8 and out.size == want.size
    if flags == capi.RUN_EXACT:
        np.testing.assert_array_equal(out, want)
    else:
        assert float(np.abs(out - want).max()) <= TOL


def test_spectrum_stage_matches_oracle():
    """Residue dequant/scatter + inverse coupling + Floor1 curve (k_spectrum) against the oracle's IMDCT input."""
    import torch
    r, pcm, b = H.decoded("3test", dense=True)
    ctx = capi.Context(0); ctx.upload_setup(H.setup_from_oracle(r))
    hb = H.batch_from_boundary(b, ctx.post_stride)
    db = ctx.create_dbatch(hb)
    spec = torch.zeros(db.spectrum_floats, dtype=torch.float32, device="cuda")
    db.run_spectrum(spec.data_ptr(), torch.cuda.current_stream().cuda_stream)
    db.result(torch.cuda.current_stream().cuda_stream)
    got = spec.cpu().numpy()
    off = 0
    for i in range(len(b.frames)):
        s = b.spectrum[i]
        if s is None:
            continue
        np.testing.assert_array_equal(got[off:off + s.size].reshape(s.shape), s, err_msg=f"frame {i}")     # integer-indexed gathers + exact FP32: bit-identical
        off += s.size
    assert off == db.spectrum_floats
    db.destroy()


def _pool():
    desc, z = setupio.load(os.path.join(H.GOLDEN, "3test.boundary.npz"))
    return desc, workloads.FramePool.from_npz(desc, z)


def _oracle_on_batch(hb):
    import bench
    r = H.O.OracleReader(H.packets("3test"))
    fr, posts, pc, cls, ent = bench.oracle_inputs(hb)
    cap = int(hb.frames["total"].astype(np.int64).sum()) + 8192
    return r.synth_batch(fr, posts, pc, cls, ent, cap, threads=os.cpu_count() or 1)


def test_config2_full_size():
    """BASELINE configs[1] at full size: 4096 stereo long-block frames, device-resident API + host API."""
    import torch
    desc, pool = _pool()
    hb = workloads.config2(pool, 4096, 20240002)
    want, clipped = _oracle_on_batch(hb)
    ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
    out, res = ctx.decode_batch(hb)
    assert out.size == want.size == 4095 * 1024 * 2
    assert float(np.abs(out - want).max()) <= TOL and res.has_clipped == clipped
    db = ctx.create_dbatch(hb)
    pcm = torch.zeros(db.samples * 2, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    db.run(pcm.data_ptr(), st)
    r2 = db.result(st)
    np.testing.assert_array_equal(pcm.cpu().numpy(), out)            # same kernels, same result
    assert db.launches == 2 and r2.samples_per_channel == res.samples_per_channel
    out_e, _ = ctx.decode_batch(hb, capi.RUN_EXACT)
    np.testing.assert_array_equal(out_e, want)
    db.destroy()


def test_config3_mixed_windows_full_size():
    """BASELINE configs[2]: 16k frames with short/long transitions (all window shapes)."""
    desc, pool = _pool()
    hb = workloads.config3(pool, 16384, 20240003)
    assert {0, 1, 2, 3} <= set(np.unique(hb.frames["window"]).tolist())
    want, clipped = _oracle_on_batch(hb)
    ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
    out, res = ctx.decode_batch(hb)
    assert out.size == want.size and res.n_inconsistent == 0
    assert float(np.abs(out - want).max()) <= TOL
    out_e, _ = ctx.decode_batch(hb, capi.RUN_EXACT)
    np.testing.assert_array_equal(out_e, want)


def test_large_chained_batches_through_the_chunked_pipeline():
    """Two chained 6000-frame calls (each runs nvb_decode_batch's four-chunk copy/compute pipeline) equal one decode."""
    from nvorbis_b200 import sharding
    desc, pool = _pool()
    hb = workloads.config3(pool, 12000, 77)
    want, clipped = _oracle_on_batch(hb)
    ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
    a, ra = ctx.decode_batch(sharding.slice_batch(hb, 0, 6000, 2))
    a = a.copy()
    b, rb = ctx.decode_batch(sharding.slice_batch(hb, 6000, 12000, 2), capi.RUN_CONTINUE)
    got = np.concatenate([a, b])
    assert got.size == want.size and float(np.abs(got - want).max()) <= TOL
    assert (ra.has_clipped or rb.has_clipped) == clipped


def test_blob_roundtrip():
    r, pcm, b, ctx = _ctx("3test")
    ctx2 = capi.Context(0)
    ctx2.import_blob(ctx.export_blob())
    hb = H.batch_from_boundary(b, ctx.post_stride, 0, 80)
    a, _ = ctx.decode_batch(hb); c, _ = ctx2.decode_batch(hb)
    np.testing.assert_array_equal(a, c)


def test_bad_entry_is_reported():
    r, pcm, b, ctx = _ctx("1test")
    hb = H.batch_from_boundary(b, ctx.post_stride)
    ent = hb.entries.copy(); ent[:] = 65535
    with pytest.raises(capi.NvbError) as e:
        ctx.decode_batch(capi.HostBatch(hb.frames, hb.posts, hb.classes, ent))
    assert e.value.status == capi.ERR_DATA


@pytest.mark.parametrize("name", H.FIXTURES)
def test_vorbis_reader_read_samples(name, golden):
    """BASELINE configs[0]: the stream through the VorbisReader.ReadSamples mirror in 4-second chunks (TestApp/Program.cs:21-26):
    product Ogg/packet unpacker -> C ABI -> GPU, against the oracle's PCM."""
    from nvorbis_b200.reader import VorbisReader
    pl = H.packets(name)
    r, pcm, b = H.decoded(name)
    with VorbisReader((pl.data, pl.sizes, pl.granules, pl.flags), batch_packets=128) as vr:
        got = vr.read_all()
        assert got.size == pcm.size == golden[name]["samples_per_channel"] * vr.channels
        assert float(np.abs(got - pcm).max()) <= TOL
        assert vr.has_clipped == golden[name]["has_clipped"]


def test_generic_spectrum_kernel_on_gpu():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, helpers as H\nfrom nvorbis_b200 import capi\n"
            "for name in ('1test', '3test'):\n"
            "    r, pcm, b = H.decoded(name)\n    ctx = capi.Context(0); ctx.upload_setup(H.setup_from_oracle(r))\n"
            "    out, _ = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride), capi.RUN_EXACT)\n"
            "    assert np.array_equal(out, pcm)\nprint('ok')\n") % (H.ROOT, os.path.join(H.ROOT, "tests"))
    # the general kernel / the per-bin fast kernel / planes / the run kernel / k_spectrum_wf with 1, 2 and 4 warps per frame
    for var, val in (("NVB_SPECTRUM_GENERIC", "1"), ("NVB_SPECTRUM_NO_PLANES", "1"), ("NVB_SPECTRUM_WARP", "1"), ("NVB_SPECTRUM_PLANES", "1"), ("NVB_SPECTRUM_RUN", "1"),
                     ("NVB_SPECTRUM_NT", "256"), ("NVB_WF_WPF", "1"), ("NVB_WF_WPF", "2"), ("NVB_WF_WPF", "4")):
        env = dict(os.environ); env[var] = val
        if var == "NVB_SPECTRUM_NT":
            env["NVB_SPECTRUM_RUN"] = "1"
        assert subprocess.check_output([sys.executable, "-c", code], env=env).decode().strip().endswith("ok")


def _s16(x):
    """The 16-bit form NVB_RUN_PCM_S16 defines: round-to-nearest-even(v * 32768), saturated."""
    return np.clip(np.rint(x.astype(np.float32) * np.float32(32768.0)), -32768, 32767).astype(np.int16)


def test_pcm_s16_epilogue_and_device_output():
    """NVB_RUN_PCM_S16 (16-bit PCM, half the read-back): the exact path equals the oracle's PCM quantised by the same rule bit for
    bit -- 3test clips, so saturation is exercised --, the fused path is at most one step away; through the chunked pipeline at
    configs[1] size too.  NVB_RUN_DEVICE_OUT leaves float / 16-bit PCM in a device buffer of the caller."""
    import torch
    r, pcm, b, ctx = _ctx("3test")
    hb = H.batch_from_boundary(b, ctx.post_stride)
    out, res = ctx.decode_batch(hb, capi.RUN_EXACT | capi.RUN_PCM_S16)
    np.testing.assert_array_equal(out, _s16(pcm))
    assert res.has_clipped and out.max() == 32767 and out.min() == -32768
    ctx.reset()
    out, _ = ctx.decode_batch(hb, capi.RUN_PCM_S16)
    assert np.abs(out.astype(np.int32) - _s16(pcm).astype(np.int32)).max() <= 1
    ctx.reset()
    d = torch.full((pcm.size + 32,), 5.0, dtype=torch.float32, device="cuda")
    res = ctx.decode_batch_ptr(hb, capi.RUN_EXACT | capi.RUN_DEVICE_OUT, d.data_ptr(), d.numel())
    got = d.cpu().numpy()
    np.testing.assert_array_equal(got[: pcm.size], pcm)
    assert (got[pcm.size:] == 5.0).all() and res.samples_per_channel * 2 == pcm.size
    ctx.reset()
    d16 = torch.full((pcm.size + 32,), 5, dtype=torch.int16, device="cuda")
    ctx.decode_batch_ptr(hb, capi.RUN_EXACT | capi.RUN_DEVICE_OUT | capi.RUN_PCM_S16, d16.data_ptr(), d16.numel())
    got = d16.cpu().numpy()
    np.testing.assert_array_equal(got[: pcm.size], _s16(pcm))
    assert (got[pcm.size:] == 5).all()
    with pytest.raises(capi.NvbError) as e:                      # a host pointer is refused, not written through
        ctx.decode_batch_ptr(hb, capi.RUN_DEVICE_OUT, np.zeros(pcm.size, np.float32).ctypes.data, pcm.size)
    assert e.value.status == capi.ERR_ARG
    ctx.close()
    # configs[1] size: four chunks, two batches in flight
    desc, pool = _pool()
    hb = workloads.config2(pool, 4096, 20240002)
    want, _ = _oracle_on_batch(hb)
    ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
    outs = [np.zeros(want.size, np.int16), np.zeros(want.size, np.int16)]
    for i in range(3):
        ctx.decode_batch_begin(hb, capi.RUN_EXACT | capi.RUN_PCM_S16, outs[i & 1].ctypes.data, outs[i & 1].size)
        if i >= 1:
            ctx.decode_batch_end()
    ctx.decode_batch_end()
    np.testing.assert_array_equal(outs[0], _s16(want))
    np.testing.assert_array_equal(outs[1], _s16(want))
    ctx.close()


def test_slot_ring_under_stress():
    """The release/acquire counter protocol of k_imdct_fused's slot ring (racecheck cannot model it): a ring of only THREE
    slots for 16 warps plus pseudo-random pauses between the protocol steps (NVB_FUSED_SLOTS / NVB_FUSED_SKEW) -- warps
    overtake each other in ways normal timing never shows -- must give bit-identical PCM to the undisturbed run, for the
    stereo stream with every window shape as one batch and as chained ragged batches."""
    import subprocess, sys
    code = ("import sys, os; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, helpers as H\nfrom nvorbis_b200 import capi\n"
            "r, pcm, b = H.decoded('3test')\nctx = capi.Context(0); ctx.upload_setup(H.setup_from_oracle(r))\n"
            "hb = H.batch_from_boundary(b, ctx.post_stride)\nout, _ = ctx.decode_batch(hb)\n"
            "assert float(np.abs(out - pcm).max()) <= 1e-5\n"
            "parts = []\nctx.reset()\n"
            "for lo, hi in ((0, 37), (37, 38), (38, 200), (200, len(b.frames))):\n"
            "    o, _ = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride, lo, hi), capi.RUN_CONTINUE if lo else 0)\n    parts.append(o.copy())\n"
            "np.save(sys.argv[1], np.concatenate([out, np.concatenate(parts)]))\nprint('ok')\n") % (H.ROOT, os.path.join(H.ROOT, "tests"))
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        outs = []
        for k, extra in enumerate(({}, {"NVB_FUSED_SLOTS": "3", "NVB_FUSED_SKEW": "1"}, {"NVB_FUSED_SLOTS": "4", "NVB_FUSED_SKEW": "5"},
                                   {"NVB_FUSED_SLOTS": "3", "NVB_FUSED_SKEW": "16"})):
            env = dict(os.environ); env.update(extra)
            path = os.path.join(td, f"o{k}.npy")
            assert subprocess.check_output([sys.executable, "-c", code, path], env=env, timeout=600).decode().strip().endswith("ok")
            outs.append(np.load(path))
        for o in outs[1:]:
            np.testing.assert_array_equal(o, outs[0])


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_begin_end_pipeline_matches_sync_on_gpu():
    """Two batches in flight through nvb_decode_batch_begin/_end (chunked copy/compute pipeline inside each): bit-identical
    to the synchronous call, tails chained from batch to batch."""
    r, pcm, b = H.decoded("3test")
    ctx = capi.Context(0)
    ctx.upload_setup(H.setup_from_oracle(r))
    n = len(b.frames)
    cuts = [0, 90, 91, 200, 290, n]
    hbs = [H.batch_from_boundary(b, ctx.post_stride, cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]
    want = []
    for i, hb in enumerate(hbs):
        out, _ = ctx.decode_batch(hb, capi.RUN_DEFAULT | (capi.RUN_CONTINUE if i else 0))
        want.append(out.copy())
    ctx.reset()
    outs = [np.zeros((capi.sum_output_bound(hb.frames) + 64) * 2, np.float32) for hb in hbs]
    for depth in (2, capi.MAX_IN_FLIGHT):
        ctx.reset()
        got, pending, clipped = [], [], False
        for i, hb in enumerate(hbs):
            ctx.decode_batch_begin(hb, capi.RUN_DEFAULT | (capi.RUN_CONTINUE if i else 0), outs[i].ctypes.data, outs[i].size)
            pending.append(i)
            if len(pending) == depth:
                j = pending.pop(0); res = ctx.decode_batch_end(); clipped |= res.has_clipped
                got.append(outs[j][: res.samples_per_channel * 2].copy())
        while pending:
            j = pending.pop(0); res = ctx.decode_batch_end(); clipped |= res.has_clipped
            got.append(outs[j][: res.samples_per_channel * 2].copy())
        for g, w in zip(got, want):
            np.testing.assert_array_equal(g, w)
        assert clipped and float(np.abs(np.concatenate(got) - pcm).max()) <= TOL
    ctx.close()


def test_floor_curve_outside_the_table_is_clamped_identically_on_gpu():
    import test_cpu_shim
    test_cpu_shim._floor_range_case(None)


@pytest.mark.parametrize("name", H.FIXTURES)
def test_gpu_path_matches_independent_ffmpeg_decoder(name):
    """The product (host unpacker + GPU synthesis behind the VorbisReader mirror) against FFmpeg's native vorbis decoder."""
    import ffmpeg_vorbis
    from nvorbis_b200.reader import VorbisReader
    pl = H.packets(name)
    offs = np.concatenate([[0], np.cumsum(pl.sizes)]).astype(np.int64)
    want = ffmpeg_vorbis.decode([bytes(pl.data[offs[i]:offs[i + 1]]) for i in range(len(pl.sizes))])
    if want is None:
        pytest.skip("no usable libavcodec in this environment")
    with VorbisReader((pl.data, pl.sizes, pl.granules, pl.flags), batch_packets=128) as vr:
        got = vr.read_all().reshape(-1, vr.channels)
    n = min(len(got), len(want))
    assert n >= len(got) - 2048
    assert float(np.abs(got[:n] - np.clip(want[:n], -0.99999994, 0.99999994)).max()) <= TOL


def test_empty_failed_and_single_packet_batches_on_gpu():
    import test_cpu_shim
    test_cpu_shim._edge_batches(None)


def test_config5_64k_frames_in_eight_shards():
    """BASELINE configs[4] at full size: a 65 536-frame stereo corpus cut into 8 contiguous shards (+1 halo frame each,
    nvorbis_b200/sharding.py), each decoded from a fresh decoder state as a rank would (here one after the other on one
    GPU): the concatenation equals the oracle's decode of the whole corpus, so shards need no data-path exchange."""
    import bench
    from nvorbis_b200 import sharding
    desc, z = setupio.load(bench.POOL)
    pool = workloads.FramePool.from_npz(desc, z)
    corpus = workloads.config2(pool, 65536, 20240005)
    ctx = capi.Context(0)
    ctx.upload_setup(setupio.to_setup(desc))
    cuts = sharding.shard_cuts(corpus.frames, 8)
    assert cuts == [8192 * k for k in range(9)]
    parts = []
    for rank in range(8):
        ctx.reset()
        out, res = ctx.decode_batch(sharding.take_shard(corpus, cuts, rank, 2))
        assert res.samples_per_channel == (8192 - (1 if rank == 0 else 0)) * 1024
        parts.append(out.copy())
    got = np.concatenate(parts)
    r = H.O.OracleReader(H.packets("3test"))
    fr, posts, pc, cls, ent = bench.oracle_inputs(corpus)
    want, _ = r.synth_batch(fr, posts, pc, cls, ent, 65536 * 1024 + 8192, threads=os.cpu_count() or 1)
    assert got.size == want.size == 65535 * 1024 * 2
    assert float(np.abs(got - want).max()) <= TOL
    ctx.close()


def test_two_contexts_on_two_devices_in_one_process():
    """One process, two GPUs: the launch configuration of every kernel is set per device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r, pcm, b = H.decoded("3test")
    outs = []
    for dev in (0, 1):
        ctx = capi.Context(dev)
        ctx.upload_setup(H.setup_from_oracle(r))
        out, res = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride))
        outs.append(out.copy())
        ctx.close()
    np.testing.assert_array_equal(outs[0], outs[1])
    assert float(np.abs(outs[0] - pcm).max()) <= TOL
