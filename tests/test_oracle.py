"""The CPU oracle against the committed golden data and against closed-form identities (no GPU)."""
import hashlib

import numpy as np
import pytest

import helpers as H
from oracle import oracle as O


@pytest.mark.parametrize("name", H.FIXTURES)
def test_golden_pcm(name, golden):
    g = golden[name]
    r = O.OracleReader(H.packets(name))
    pcm = r.read_all()
    assert r.channels == g["channels"] and r.sample_rate == g["sample_rate"] and [r.block0, r.block1] == g["block_sizes"]
    assert pcm.size // r.channels == g["samples_per_channel"]
    assert r.has_clipped == g["has_clipped"]
    assert hashlib.sha256(pcm.astype("<f4").tobytes()).hexdigest() == g["sha256_pcm"]
    for idx, vals in g["spots"].items():
        np.testing.assert_array_equal(pcm[int(idx) * r.channels:(int(idx) + 1) * r.channels], np.array(vals, np.float32))


@pytest.mark.parametrize("name", ["1test", "2test", "3test"])
def test_independent_probe_values(name, golden):
    """values from an independent float64 direct-formula decode of the same files."""
    p = golden[name]["survey_probe"]
    r = O.OracleReader(H.packets(name))
    pcm = r.read_all()
    ch = r.channels
    assert pcm.size // ch == p["samples"]
    assert pcm.size // ch == golden[name]["last_granule"]           # EOS trim lands exactly on the last granule
    if "sum_abs" in p:
        assert abs(float(np.abs(pcm.astype(np.float64)).sum()) - p["sum_abs"]) < 0.02
    for idx, vals in p.get("spots", {}).items():
        np.testing.assert_allclose(pcm[int(idx) * ch:(int(idx) + 1) * ch], vals, atol=2e-6)
    if "clipped" in p:
        assert golden[name]["clipped_samples"] == p["clipped"]


def test_issue6_tail_drain(golden):
    """The EOS page of issue6test holds only a zero-length packet, which the reference drops together with the
    page (Ogg/PageReader.cs:131), so no packet is ever flagged end-of-stream, nothing is trimmed, and the last
    block's tail is drained when the provider runs dry (StreamDecoder.cs:352-356)."""
    g = golden["issue6test"]
    assert g["samples_per_channel"] == 549184 and g["last_granule"] == 548223
    assert g["audio_frames"] == g["frames_ok"] + 1


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192])
def test_imdct_closed_form(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n // 2).astype(np.float32)
    y = O.mdct_reverse(x)
    i = np.arange(n)[:, None]; k = np.arange(n // 2)[None, :]
    ref = (np.cos(2 * np.pi / n * (i + 0.5 + n / 4) * (k + 0.5)) * x.astype(np.float64)[None, :]).sum(1)
    assert np.abs(y - ref).max() < 2e-5 * np.sqrt(n / 256)


@pytest.mark.parametrize("n", [64, 128])
def test_imdct_small_n_quirk(n):
    """Mdct.cs runs iterations 0/1 and the ld654 tail unconditionally, so for N < 256 it does not compute the
    IMDCT.  The oracle keeps the reference's behaviour."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n // 2).astype(np.float32)
    y = O.mdct_reverse(x)
    i = np.arange(n)[:, None]; k = np.arange(n // 2)[None, :]
    ref = (np.cos(2 * np.pi / n * (i + 0.5 + n / 4) * (k + 0.5)) * x.astype(np.float64)[None, :]).sum(1)
    assert np.abs(y - ref).max() > 0.1


def test_tdac_reconstruction():
    """Forward MDCT (float64, direct) -> oracle IMDCT -> Vorbis window -> overlap-add reconstructs the signal."""
    n = 2048
    rng = np.random.default_rng(1)
    sig = rng.standard_normal(n * 3 // 2)
    w = O.calc_window(n, n, n).astype(np.float64)
    i = np.arange(n)[:, None]; k = np.arange(n // 2)[None, :]
    basis = np.cos(2 * np.pi / n * (i + 0.5 + n / 4) * (k + 0.5))
    blocks = []
    for b in range(2):
        seg = sig[b * n // 2: b * n // 2 + n] * w
        X = (basis * seg[:, None]).sum(0)
        blocks.append(O.mdct_reverse(X.astype(np.float32)).astype(np.float64) * w)
    rec = blocks[0][n // 2:] + blocks[1][:n // 2]
    np.testing.assert_allclose(rec / (n / 4), sig[n // 2: n], atol=5e-4)     # unscaled IMDCT: gain N/4


def test_window_shape():
    w = O.calc_window(256, 2048, 2048)
    assert (w[:448] == 0).all() and (w[576:1024] == 1).all() and w[448] > 0
    i = np.arange(1024)
    ideal = np.sin(np.pi / 2 * np.sin((i + 0.5) / 1024 * np.pi / 2) ** 2)
    np.testing.assert_allclose(w[1024:][::-1], ideal, atol=3e-7)
    w2 = O.calc_window(2048, 2048, 2048)
    np.testing.assert_allclose(w2[:1024] ** 2 + w2[1024:] ** 2, 1.0, atol=3e-7)      # power complementarity


def test_inverse_coupling_cases():
    m = np.array([1.0, 1.0, -1.0, -1.0, 0.0, 2.0], np.float32)
    a = np.array([0.5, -0.5, 0.5, -0.5, 3.0, 0.0], np.float32)
    nm, na = O.inverse_couple(m, a)
    np.testing.assert_array_equal(nm, np.array([1.0, 0.5, -1.0, -0.5, 0.0, 2.0], np.float32))
    np.testing.assert_array_equal(na, np.array([0.5, 1.0, -0.5, -1.0, 3.0, 2.0], np.float32))


def test_inverse_db_table():
    t = O.inverse_db_table()
    assert t[255] == 1.0 and abs(t[0] - 1.0649863e-07) < 1e-14 and (np.diff(t) > 0).all()


@pytest.mark.parametrize("name,threads", [("1test", 1), ("3test", 1), ("3test", 4), ("issue6test", 1)])
def test_synthesis_from_boundary_records(name, threads):
    """The stand-alone synthesis entry (what the GPU path is compared with) reproduces the full decoder."""
    r, pcm, b = H.decoded(name)
    if threads > 1:                                   # the threaded split needs drain-free batches
        hi = int(np.nonzero(b.frames["ok"] == 0)[0][0]) if (b.frames["ok"] == 0).any() else len(b.frames)
        out, _ = H.oracle_synth(r, b, 0, hi, threads)
        ref, _ = H.oracle_synth(r, b, 0, hi, 1)
        np.testing.assert_array_equal(out, ref)
    else:
        out, clipped = H.oracle_synth(r, b)
        np.testing.assert_array_equal(out, pcm)
        assert clipped == r.has_clipped


@pytest.mark.parametrize("name", H.FIXTURES)
def test_independent_ffmpeg_decoder_agrees(name):
    """Cross-check against a decoder that shares no code with the oracle or with NVorbis: FFmpeg's native vorbis decoder
    (the libavcodec inside opencv-python-headless, driven through ctypes on the same demuxed packets).  Agreement to float
    rounding over the common prefix proves the oracle decodes the reference's own test files correctly; the lengths differ
    only where NVorbis trims / drains the end of the stream (StreamDecoder.cs:429-437, :352-356)."""
    import ffmpeg_vorbis
    pl = H.packets(name)
    offs = np.concatenate([[0], np.cumsum(pl.sizes)]).astype(np.int64)
    got = ffmpeg_vorbis.decode([bytes(pl.data[offs[i]:offs[i + 1]]) for i in range(len(pl.sizes))])
    if got is None:
        pytest.skip("no usable libavcodec in this environment")
    r, pcm, _ = H.decoded(name)
    want = pcm.reshape(-1, r.channels)
    n = min(len(got), len(want))
    assert got.shape[1] == r.channels and n >= len(want) - 2048 and n >= len(got) - 2048
    clipped = np.clip(got[:n], -0.99999994, 0.99999994)          # Utils.ClipValue; FFmpeg hands out un-clipped floats
    assert float(np.abs(clipped - want[:n]).max()) <= 2e-6
