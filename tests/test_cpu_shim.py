"""The product's real kernel and C-ABI sources, compiled against tests/cpu_shim's CUDA-on-CPU emulation
(one OS thread per CUDA thread) and compared with the oracle.  This is how the kernels are debugged in the
GPU-less build container; the GPU parity tests proper are in test_gpu_parity.py."""
import numpy as np
import pytest

import helpers as H
from nvorbis_b200 import capi


@pytest.fixture(scope="module")
def shim():
    return H.build_shim()


def _ctx(shim, name):
    r, pcm, b = H.decoded(name)
    ctx = capi.Context(0, lib_path=shim)
    ctx.upload_setup(H.setup_from_oracle(r))
    return r, pcm, b, ctx


def test_mono_full_stream_exact_and_fused(shim):
    r, pcm, b, ctx = _ctx(shim, "1test")
    hb = H.batch_from_boundary(b, ctx.post_stride)
    out, res = ctx.decode_batch(hb, capi.RUN_EXACT)
    np.testing.assert_array_equal(out, pcm)                     # stb dataflow, no FMA: bit-identical
    assert res.samples_per_channel == pcm.size and not res.has_clipped and res.n_failed == 0
    ctx.reset()
    out, res = ctx.decode_batch(hb, capi.RUN_DEFAULT)
    assert out.size == pcm.size and np.abs(out - pcm).max() <= 1e-5


def test_stereo_coupled_residue2_with_clipping(shim):
    r, pcm, b, ctx = _ctx(shim, "3test")
    lo, hi = 0, 70
    want, clipped = H.oracle_synth(r, b, lo, hi)
    hb = H.batch_from_boundary(b, ctx.post_stride, lo, hi)
    out, res = ctx.decode_batch(hb, capi.RUN_EXACT)
    np.testing.assert_array_equal(out, want)
    ctx.reset()
    out, res = ctx.decode_batch(hb, capi.RUN_DEFAULT)
    assert out.size == want.size and np.abs(out - want).max() <= 1e-5
    # a stretch that clips (the oracle reports HasClipped for the full stream)
    idx = int(np.argmax(np.abs(H.decoded("3test")[1]))) // 2
    # find the frame run containing the loudest sample: decode frames [a, a+12)
    lens = np.maximum(b.frames["valid"] - b.frames["start"], 0); lens[0] = 0
    a = max(int(np.searchsorted(np.cumsum(lens), idx)) - 6, 1)
    want, clipped = H.oracle_synth(r, b, a, a + 12)
    assert clipped
    ctx.reset()
    out, res = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride, a, a + 12), capi.RUN_DEFAULT)
    assert res.has_clipped and np.abs(out - want).max() <= 1e-5 and np.abs(out).max() <= 0.99999994


def test_chained_batches_carry_the_tail(shim):
    r, pcm, b, ctx = _ctx(shim, "1test")
    n = len(b.frames)
    for flags in (capi.RUN_EXACT, capi.RUN_DEFAULT):
        ctx.reset()
        parts, pos = [], 0
        for cut in (3, 4, 11, n):                      # includes a one-frame batch
            out, _ = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride, pos, cut), flags | (capi.RUN_CONTINUE if pos else 0))
            parts.append(out.copy()); pos = cut
        got = np.concatenate(parts)
        if flags == capi.RUN_EXACT:
            np.testing.assert_array_equal(got, pcm)
        else:
            assert got.size == pcm.size and np.abs(got - pcm).max() <= 1e-5


def test_failed_packets_drain_the_tail(shim):
    r, pcm, b, ctx = _ctx(shim, "1test")
    fr = b.frames.copy()
    fr["ok"][5] = 0; fr["ok"][6] = 0; fr["ok"][12] = 0
    b2 = H.O.Boundary(b.channels, fr, b.block_size, b.valid_untrimmed, b.no_exec_mask, b.posts, b.post_counts, b.classes, b.entries)
    want, _ = H.oracle_synth(r, b2)
    for flags in (capi.RUN_EXACT, capi.RUN_DEFAULT):
        ctx.reset()
        out, res = ctx.decode_batch(H.batch_from_boundary(b2, ctx.post_stride), flags)
        assert res.n_failed == 3 and out.size == want.size
        if flags == capi.RUN_EXACT:
            np.testing.assert_array_equal(out, want)
        else:
            assert np.abs(out - want).max() <= 1e-5
    # drain at the start of a chained batch: the tail comes from the carried block
    ctx.reset()
    a, _ = ctx.decode_batch(H.batch_from_boundary(b2, ctx.post_stride, 0, 5), capi.RUN_DEFAULT)
    c, _ = ctx.decode_batch(H.batch_from_boundary(b2, ctx.post_stride, 5, len(fr)), capi.RUN_DEFAULT | capi.RUN_CONTINUE)
    got = np.concatenate([a, c])
    assert got.size == want.size and np.abs(got - want).max() <= 1e-5


def test_blob_roundtrip_and_errors(shim):
    r, pcm, b, ctx = _ctx(shim, "1test")
    blob = ctx.export_blob()
    ctx2 = capi.Context(0, lib_path=shim)
    ctx2.import_blob(blob)
    hb = H.batch_from_boundary(b, ctx.post_stride)
    out, _ = ctx2.decode_batch(hb, capi.RUN_EXACT)
    np.testing.assert_array_equal(out, pcm)
    bad = blob.copy(); bad[0] ^= 0xFF
    with pytest.raises(capi.NvbError) as e:
        ctx2.import_blob(bad)
    assert e.value.status == capi.ERR_DATA
    # malformed batches are rejected up front with NVB_ERR_DATA (InvalidDataException in the reference)
    fr = hb.frames.copy(); fr["mode"][3] = 9
    with pytest.raises(capi.NvbError) as e:
        ctx.decode_batch(capi.HostBatch(fr, hb.posts, hb.classes, hb.entries))
    assert e.value.status == capi.ERR_DATA
    fr = hb.frames.copy(); fr["entries_off"][4] = 1 << 30
    with pytest.raises(capi.NvbError) as e:
        ctx.decode_batch(capi.HostBatch(fr, hb.posts, hb.classes, hb.entries))
    assert e.value.status == capi.ERR_DATA
    with pytest.raises(capi.NvbError) as e:
        ctx.decode_batch(hb, out=np.zeros(16, np.float32))
    assert e.value.status == capi.ERR_CAPACITY
    # setups outside the envelope
    desc = H.desc_from_oracle(r)
    from nvorbis_b200 import setupio
    d2 = dict(desc); d2["mappings"] = [dict(m, n_submaps=2) for m in desc["mappings"]]
    with pytest.raises(capi.NvbError) as e:
        ctx2.upload_setup(setupio.to_setup(d2))
    assert e.value.status == capi.ERR_UNSUPPORTED
    # type 0 floors: a malformed header is InvalidDataException (Floor0.cs:37); a bark map that indexes past the cos map
    # would throw on every packet (Floor0.cs:166) and is refused
    f0 = dict(type=0, order=0, rate=44100, bark_map_size=64, amp_bits=6, amp_ofs=100)
    d3 = dict(desc); d3["floors"] = [dict(f, **f0) for f in desc["floors"]]
    with pytest.raises(capi.NvbError) as e:
        ctx2.upload_setup(setupio.to_setup(d3))
    assert e.value.status == capi.ERR_DATA
    d4 = dict(desc); d4["floors"] = [dict(f, **dict(f0, order=8, bark_map_size=8192)) for f in desc["floors"]]
    with pytest.raises(capi.NvbError) as e:
        ctx2.upload_setup(setupio.to_setup(d4))
    assert e.value.status == capi.ERR_UNSUPPORTED


def test_setup_tables_match_oracle(shim):
    """Windows and MDCT twiddles the library derives itself equal the oracle's (same float/double expression order)."""
    r, pcm, b, ctx = _ctx(shim, "3test")
    blob = ctx.export_blob()
    hdr_off = {}
    import struct
    # BlobHeader: magic, abi, total(8), channels, rate, bs0, bs1, 5 counts, post_stride, max_items, pad, then u64 offsets
    vals = struct.unpack_from("<II Q 4i 5i 3i", blob, 0)
    offs = struct.unpack_from("<6Q 2Q 8Q 2Q 2Q Q Q", blob, 16 + 16 + 20 + 12)
    off_win_short, off_win_long = offs[6], offs[7]
    a0, a1, b0, b1, c0, c1, br0, br1 = offs[8:16]
    ws = np.frombuffer(blob, np.float32, r.block0, off_win_short)
    np.testing.assert_array_equal(ws, r.mode_window(0, 0))
    for w in range(4):
        wl = np.frombuffer(blob, np.float32, r.block1, off_win_long + 4 * w * r.block1)
        np.testing.assert_array_equal(wl, r.mode_window(1, w))
    for n, (oa, ob, oc, obr) in ((r.block0, (a0, b0, c0, br0)), (r.block1, (a1, b1, c1, br1))):
        A, B, Cc, BR = H.O.mdct_tables(n)
        np.testing.assert_array_equal(np.frombuffer(blob, np.float32, n // 2, oa), A)
        np.testing.assert_array_equal(np.frombuffer(blob, np.float32, n // 2, ob), B)
        np.testing.assert_array_equal(np.frombuffer(blob, np.float32, n // 4, oc), Cc)
        np.testing.assert_array_equal(np.frombuffer(blob, np.uint16, n // 8, obr), BR)


def test_vorbis_reader_end_to_end(shim, golden):
    """The VorbisReader mirror (host unpacker -> C ABI -> ReadSamples) over the emulated device, with batches much
    smaller than the stream so that several chained nvb_decode_batch calls are needed."""
    from nvorbis_b200.reader import VorbisReader
    pl = H.packets("1test")
    r, pcm, b = H.decoded("1test")
    with VorbisReader((pl.data, pl.sizes, pl.granules, pl.flags), batch_packets=7, lib_path=shim) as vr:
        assert (vr.channels, vr.sample_rate) == (1, 44100)
        got = vr.read_all(chunk_seconds=0.05)
        assert got.size == pcm.size == golden["1test"]["samples_per_channel"]
        assert np.abs(got - pcm).max() <= 1e-5 and not vr.has_clipped
        assert vr.read_samples(np.zeros(64, np.float32), 0, 64) == 0 and vr.is_end_of_stream
        vr.seek_to_start()
        buf = np.zeros(1000, np.float32)
        assert vr.read_samples(buf, 0, 1000) == 1000 and np.abs(buf - pcm[:1000]).max() <= 1e-5


def test_generic_spectrum_kernel(shim):
    """NVB_SPECTRUM_GENERIC forces the general residue/floor kernel (used for setups outside the fast path's envelope)."""
    import os, subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, helpers as H\nfrom nvorbis_b200 import capi\n"
            "r, pcm, b = H.decoded('3test')\nctx = capi.Context(0, lib_path=%r); ctx.upload_setup(H.setup_from_oracle(r))\n"
            "want, _ = H.oracle_synth(r, b, 0, 40)\nout, _ = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride, 0, 40), capi.RUN_EXACT)\n"
            "assert np.array_equal(out, want)\nprint('ok')\n") % (H.ROOT, os.path.join(H.ROOT, "tests"), shim)
    # the general kernel / the per-bin fast kernel / planes / the run kernel / k_spectrum_wf with 1, 2 and 4 warps per frame
    for var, val in (("NVB_SPECTRUM_GENERIC", "1"), ("NVB_SPECTRUM_NO_PLANES", "1"), ("NVB_SPECTRUM_WARP", "1"), ("NVB_SPECTRUM_PLANES", "1"), ("NVB_SPECTRUM_RUN", "1"),
                     ("NVB_SPECTRUM_NT", "256"), ("NVB_WF_WPF", "1"), ("NVB_WF_WPF", "2"), ("NVB_WF_WPF", "4")):
        env = dict(os.environ); env[var] = val
        if var == "NVB_SPECTRUM_NT":
            env["NVB_SPECTRUM_RUN"] = "1"
        assert subprocess.check_output([sys.executable, "-c", code], env=env).decode().strip().endswith("ok")


def test_slot_ring_wraps_with_mixed_windows(shim):
    """One emulated SM (NVB_SHIM_SMS=1) makes a single CTA walk the whole batch, so the fused kernel's slot ring is
    reused many times; mixed short/long frames (BASELINE configs[2] generator) exercise every window shape."""
    import os, subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, helpers as H, bench\nfrom nvorbis_b200 import capi, setupio, workloads\n"
            "desc, z = setupio.load(%r)\npool = workloads.FramePool.from_npz(desc, z)\n"
            "hb = workloads.config3(pool, 90, 20240003)\n"
            "r = H.O.OracleReader(H.packets('3test'))\nfr, posts, pc, cls, ent = bench.oracle_inputs(hb)\n"
            "want, _ = r.synth_batch(fr, posts, pc, cls, ent, int(hb.frames['total'].astype(np.int64).sum()) + 8192)\n"
            "ctx = capi.Context(0, lib_path=%r); ctx.upload_setup(setupio.to_setup(desc))\n"
            "out, res = ctx.decode_batch(hb)\nassert out.size == want.size and np.abs(out - want).max() <= 1e-5\nprint('ok')\n"
            ) % (H.ROOT, os.path.join(H.ROOT, "tests"), os.path.join(H.GOLDEN, "3test.boundary.npz"), shim)
    env = dict(os.environ, NVB_SHIM_SMS="1")
    assert subprocess.check_output([sys.executable, "-c", code], env=env, timeout=600).decode().strip().endswith("ok")
    # the same batch through nvb_decode_batch's chunked copy/compute pipeline (four frame ranges, halo across chunk edges)
    env = dict(os.environ, NVB_SHIM_SMS="2", NVB_CHUNK_MIN="16")
    assert subprocess.check_output([sys.executable, "-c", code], env=env, timeout=600).decode().strip().endswith("ok")


def test_begin_end_pipeline_matches_sync(shim):
    """nvb_decode_batch_begin/_end with two and with NVB_MAX_IN_FLIGHT (three) batches in flight: the same PCM as consecutive
    nvb_decode_batch calls, tails chained across the batches; one begin more is refused."""
    r, pcm, b, ctx = _ctx(shim, "1test")
    n = len(b.frames)
    cuts = [0, 5, 9, 16, n]
    hbs = [H.batch_from_boundary(b, ctx.post_stride, cuts[i], cuts[i + 1]) for i in range(4)]
    outs = [np.zeros(capi.sum_output_bound(hb.frames) + 64, np.float32) for hb in hbs]
    for depth in (2, capi.MAX_IN_FLIGHT):
        ctx.reset()
        got, pending = [], []
        for i, hb in enumerate(hbs):
            ctx.decode_batch_begin(hb, capi.RUN_EXACT | (capi.RUN_CONTINUE if i else 0), outs[i].ctypes.data, outs[i].size)
            pending.append(i)
            if len(pending) == depth:
                if depth == capi.MAX_IN_FLIGHT and i == depth - 1:
                    with pytest.raises(capi.NvbError) as e:
                        ctx.decode_batch_begin(hbs[depth], capi.RUN_EXACT | capi.RUN_CONTINUE, outs[depth].ctypes.data, outs[depth].size)
                    assert e.value.status == capi.ERR_STATE
                j = pending.pop(0)
                res = ctx.decode_batch_end()
                got.append(outs[j][: res.samples_per_channel].copy())
        while pending:
            j = pending.pop(0)
            res = ctx.decode_batch_end()
            got.append(outs[j][: res.samples_per_channel].copy())
        np.testing.assert_array_equal(np.concatenate(got), pcm)
    with pytest.raises(capi.NvbError) as e:
        ctx.decode_batch_end()
    assert e.value.status == capi.ERR_STATE


def _floor_range_case(lib_path):
    """Floor posts whose curve leaves inverse_dB_table (the reference throws IndexOutOfRangeException, Floor1.cs:318,338):
    every spectrum kernel clamps the same way and counts the frame in n_floor_range; in-range channels of the same frames
    stay on the multiply-high path."""
    import os, subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, helpers as H\nfrom nvorbis_b200 import capi\n"
            "r, pcm, b = H.decoded('3test')\nctx = capi.Context(0, lib_path=%s); ctx.upload_setup(H.setup_from_oracle(r))\n"
            "hb = H.batch_from_boundary(b, ctx.post_stride, 0, 30)\n"
            "p = hb.posts.reshape(30, 2, ctx.post_stride)\n"
            "for i in (7, 8, 19):\n    p[i, 0, 1] = 200; p[i, 0, 2] = 255\n"          # y = 200 * mult, 255 * mult: far above 255
            "out, res = ctx.decode_batch(hb, capi.RUN_EXACT)\n"
            "np.save(sys.argv[1], out); print(res.n_floor_range)\n") % (H.ROOT, os.path.join(H.ROOT, "tests"), repr(lib_path))
    import tempfile
    outs, counts = [], []
    with tempfile.TemporaryDirectory() as td:
        for k, var in enumerate((None, "NVB_SPECTRUM_GENERIC", "NVB_SPECTRUM_PLANES", "NVB_SPECTRUM_NO_PLANES")):
            env = dict(os.environ)
            if var:
                env[var] = "1"
            path = os.path.join(td, "o%d.npy" % k)
            counts.append(int(subprocess.check_output([sys.executable, "-c", code, path], env=env, timeout=600).decode().split()[-1]))
            outs.append(np.load(path))
    assert counts[0] >= 3 and len(set(counts)) == 1, counts
    for o in outs[1:]:
        np.testing.assert_array_equal(o, outs[0])


def test_floor_curve_outside_the_table_is_clamped_identically(shim):
    _floor_range_case(shim)


def _edge_batches(lib_path):
    """Empty batch, a batch of failed packets only, a single packet: nothing is emitted and nothing hangs (the first block
    of a stream leaves only its tail, StreamDecoder.cs:446-450)."""
    r, pcm, b = H.decoded("1test")
    ctx = capi.Context(0, lib_path=lib_path)
    ctx.upload_setup(H.setup_from_oracle(r))
    out, res = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride, 0, 0))
    assert out.size == 0 and res.samples_per_channel == 0 and res.n_failed == 0
    fr = np.zeros(3, capi.FRAME_DTYPE); fr["status"] = capi.FRAME_FAILED
    out, res = ctx.decode_batch(capi.HostBatch(fr, np.zeros(3 * ctx.post_stride, np.int16), np.zeros(0, np.uint8), np.zeros(0, np.uint16)))
    assert out.size == 0 and res.n_failed == 3
    ctx.decode_batch_begin(H.batch_from_boundary(b, ctx.post_stride, 0, 0), capi.RUN_DEFAULT, 0, 0)
    assert ctx.decode_batch_end().samples_per_channel == 0
    out, res = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride, 0, 1))
    assert out.size == 0
    out, res = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride, 1, len(b.frames)), capi.RUN_CONTINUE | capi.RUN_EXACT)
    np.testing.assert_array_equal(out, pcm)                     # ... and the stream continues from that tail
    ctx.close()


def test_empty_failed_and_single_packet_batches(shim):
    _edge_batches(shim)


def _s16(x):
    """The 16-bit form NVB_RUN_PCM_S16 defines: round-to-nearest-even(v * 32768), saturated."""
    return np.clip(np.rint(x.astype(np.float32) * np.float32(32768.0)), -32768, 32767).astype(np.int16)


def test_pcm_s16_and_device_out(shim):
    """NVB_RUN_PCM_S16: the exact path quantised on the device equals the oracle's PCM quantised by the same rule, bit for bit
    (ragged chunk boundaries included); NVB_RUN_DEVICE_OUT leaves the PCM in the caller's (here: emulated) device buffer."""
    r, pcm, b, ctx = _ctx(shim, "3test")
    lo, hi = 0, 48
    want, _ = H.oracle_synth(r, b, lo, hi)
    hb = H.batch_from_boundary(b, ctx.post_stride, lo, hi)
    out, res = ctx.decode_batch(hb, capi.RUN_EXACT | capi.RUN_PCM_S16)
    assert out.dtype == np.int16 and res.samples_per_channel * 2 == want.size
    np.testing.assert_array_equal(out, _s16(want))
    ctx.reset()
    out, _ = ctx.decode_batch(hb, capi.RUN_DEFAULT | capi.RUN_PCM_S16)
    assert np.abs(out.astype(np.int32) - _s16(want).astype(np.int32)).max() <= 1      # fused path: within 1e-5 of the oracle => at most one step apart
    ctx.reset()
    dev = np.full(want.size + 64, 7.0, np.float32)                  # the emulation's "device memory"
    res = ctx.decode_batch_ptr(hb, capi.RUN_EXACT | capi.RUN_DEVICE_OUT, dev.ctypes.data, dev.size)
    np.testing.assert_array_equal(dev[: want.size], want)
    assert (dev[want.size:] == 7.0).all()
    ctx.reset()
    dev16 = np.full(want.size + 64, 7, np.int16)
    ctx.decode_batch_ptr(hb, capi.RUN_EXACT | capi.RUN_DEVICE_OUT | capi.RUN_PCM_S16, dev16.ctypes.data, dev16.size)
    np.testing.assert_array_equal(dev16[: want.size], _s16(want))
    assert (dev16[want.size:] == 7).all()


def test_wave_writer_mirror(shim, tmp_path):
    """TestApp's pipeline (Program.cs:12-28 + WaveWriter.cs): stream -> ReadSamples -> float WAV; the header fields, the sizes
    patched on close, the reference's offset-44 quirk when asked for, and the 16-bit form."""
    import struct, wave as pywave
    from nvorbis_b200 import wave as W
    pl = H.packets("1test")
    r, pcm, b = H.decoded("1test")
    p = str(tmp_path / "o.wav")
    n = W.decode_to_wav((pl.data, pl.sizes, pl.granules, pl.flags), p, lib_path=shim, batch_packets=11)
    raw = open(p, "rb").read()
    assert n == pcm.size and raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[38:42] == b"data"
    fmt_size, enc, ch, rate, bps, align, bits, extra = struct.unpack("<ihhiihhh", raw[16:38])
    assert (fmt_size, enc, ch, rate, bps, align, bits, extra) == (18, 3, 1, 44100, 4 * 44100, 4, 32, 0)
    assert struct.unpack("<I", raw[4:8])[0] == len(raw) - 8 and struct.unpack("<I", raw[42:46])[0] == len(raw) - 46
    assert np.abs(np.frombuffer(raw[46:], "<f4") - pcm).max() <= 1e-5
    q = str(tmp_path / "q.wav")
    with W.WaveWriter(q, 44100, 1, reference_quirk=True) as ww:
        ww.write_samples(pcm, 0, pcm.size)
    rq = open(q, "rb").read()
    assert struct.unpack("<I", rq[44:48])[0] == len(rq) - 48 and rq[42:44] == b"\0\0"        # WaveWriter.cs:56-57
    s = str(tmp_path / "s.wav")
    W.decode_to_wav((pl.data, pl.sizes, pl.granules, pl.flags), s, sample_format="s16", lib_path=shim)
    with pywave.open(s, "rb") as wf:                                                            # a valid PCM file for any reader
        assert (wf.getnchannels(), wf.getsampwidth(), wf.getframerate(), wf.getnframes()) == (1, 2, 44100, pcm.size)
        got = np.frombuffer(wf.readframes(wf.getnframes()), "<i2")
    assert np.abs(got.astype(np.int32) - _s16(pcm).astype(np.int32)).max() <= 1


def test_one_kernel_synthesis_equals_two_kernels(shim):
    """NVB_RUN_ONE_KERNEL (k_imdct_fused_t<false, C>: spectrum stage inside the fused kernel, no dense spectrum) against the
    two-kernel path and the oracle: mono with short + long blocks, stereo with coupling, chained ragged batches."""
    for name, hi in (("1test", None), ("3test", 40)):
        r, pcm, b, ctx = _ctx(shim, name)
        want, _ = H.oracle_synth(r, b, 0, hi)
        hb = H.batch_from_boundary(b, ctx.post_stride, 0, hi)
        two, _ = ctx.decode_batch(hb, capi.RUN_DEFAULT)
        two = two.copy()
        ctx.reset()
        one, res = ctx.decode_batch(hb, capi.RUN_ONE_KERNEL)
        assert one.size == want.size and np.abs(one - want).max() <= 1e-5
        np.testing.assert_array_equal(one, two)
        ctx.reset()
        n = len(b.frames) if hi is None else hi
        chained = {}
        for mode in (capi.RUN_ONE_KERNEL, capi.RUN_TWO_KERNELS):               # (a carried tail is a windowed block: not the fused TDAC arithmetic)
            ctx.reset()
            parts, pos = [], 0
            for cut in (3, 4, 11, n):
                out, _ = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride, pos, cut), mode | (capi.RUN_CONTINUE if pos else 0))
                parts.append(out.copy()); pos = cut
            chained[mode] = np.concatenate(parts)
        np.testing.assert_array_equal(chained[capi.RUN_ONE_KERNEL], chained[capi.RUN_TWO_KERNELS])
        assert np.abs(chained[capi.RUN_ONE_KERNEL] - want).max() <= 1e-5
        ctx.close()


def test_one_kernel_bad_inputs(shim):
    import test_gpu_one_kernel
    test_gpu_one_kernel._bad_inputs_case(shim)
