"""GPU-side packet unpack (k_unpack, nvb_decode_packets): the device must produce the same boundary
records as the host unpacker (libnvorbis_host.so, itself checked against the oracle's integers in test_host_unpack.py) -- bit
for bit, on every packet of the reference's four test files and on synthetic streams -- and the same PCM as the record path.
CPU: the real kernel source under tests/cpu_shim; GPU (-m gpu): the real library."""
import numpy as np
import pytest

import helpers as H
from nvorbis_b200 import capi, hostlib


def _streams(name):
    pk = H.packets(name)
    mk = lambda: hostlib.HostStream(packets=(pk.data, pk.sizes, pk.granules, pk.flags))
    return mk(), mk()


def _compare_records(name, lib_path, count=None):
    ha, hb_ = _streams(name)
    ctx = capi.Context(0, lib_path=lib_path)
    ctx.upload_setup(ha.setup())
    ctx.upload_unpack_tables(ha.unpack_tables())
    n = count or ha.n_audio_packets + 1
    want, eos_a = ha.unpack(n, threads=2)
    pb, eos_b = hb_.packet_batch(n)
    assert eos_a == eos_b and len(pb.frames) == len(want.frames)
    frames, posts, classes, entries = ctx.unpack_packets(pb, ha.post_stride)
    for fld in ("status", "mode", "window", "start", "valid", "total", "exec_mask", "res_decoded", "entry_count"):
        np.testing.assert_array_equal(frames[fld], want.frames[fld], err_msg=fld)
    np.testing.assert_array_equal(posts, want.posts)
    C = ha.channels
    for i, f in enumerate(want.frames):
        if f["status"] != capi.FRAME_OK or not f["res_decoded"]:
            continue
        nxt_c = int(want.frames["classes_off"][i + 1]) if i + 1 < len(want.frames) else want.classes.size
        k = nxt_c - int(f["classes_off"])
        np.testing.assert_array_equal(classes[i, :k], want.classes[int(f["classes_off"]): nxt_c], err_msg=f"classes of packet {i}")
        e0, ec = int(f["entries_off"]), int(f["entry_count"])
        np.testing.assert_array_equal(entries[i, :ec], want.entries[e0: e0 + ec], err_msg=f"entries of packet {i}")
    ctx.close()
    return len(want.frames)


def _compare_pcm(name, lib_path, chunk, flags=capi.RUN_EXACT):
    """Whole stream through nvb_decode_packets in chained batches == the oracle's PCM."""
    r, pcm, b = H.decoded(name)
    ha, _ = _streams(name)
    ctx = capi.Context(0, lib_path=lib_path)
    ctx.upload_setup(ha.setup())
    ctx.upload_unpack_tables(ha.unpack_tables())
    parts, first = [], True
    while True:
        pb, eos = ha.packet_batch(chunk)
        out, res = ctx.decode_packets(pb, flags | (0 if first else capi.RUN_CONTINUE))
        parts.append(out.copy()); first = False
        if eos:
            break
    got = np.concatenate(parts)
    if flags & capi.RUN_EXACT:
        np.testing.assert_array_equal(got, pcm)
    else:
        assert got.size == pcm.size and float(np.abs(got - pcm).max()) <= 1e-5
    ctx.close()


@pytest.mark.parametrize("name", ["1test", "3test"])
def test_device_records_equal_host_records_on_cpu_shim(name):
    _compare_records(name, H.build_shim(), count=40)


def test_packets_to_pcm_on_cpu_shim():
    _compare_pcm("1test", H.build_shim(), 9)


def test_unpack_tables_are_validated():
    ha, _ = _streams("1test")
    ctx = capi.Context(0, lib_path=H.build_shim())
    with pytest.raises(capi.NvbError):                       # no setup yet
        ctx.upload_unpack_tables(ha.unpack_tables())
    ctx.upload_setup(ha.setup())
    blob = ha.unpack_tables()
    bad = blob.copy(); bad[0] ^= 0xFF
    with pytest.raises(capi.NvbError) as e:
        ctx.upload_unpack_tables(bad)
    assert e.value.status == capi.ERR_DATA
    other, _ = _streams("3test")                             # tables of another stream do not fit this setup
    with pytest.raises(capi.NvbError) as e:
        ctx.upload_unpack_tables(other.unpack_tables())
    assert e.value.status == capi.ERR_DATA
    pb, _ = ha.packet_batch(4)
    with pytest.raises(capi.NvbError) as e:                  # packets before the tables
        ctx.decode_packets(pb)
    assert e.value.status == capi.ERR_STATE
    ctx.upload_unpack_tables(blob)
    ctx.decode_packets(pb)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", H.FIXTURES)
def test_device_records_equal_host_records(name):
    assert _compare_records(name, None) > 20


@pytest.mark.gpu
@pytest.mark.parametrize("name", H.FIXTURES)
@pytest.mark.parametrize("flags", [capi.RUN_EXACT, capi.RUN_DEFAULT, capi.RUN_ONE_KERNEL, capi.RUN_TWO_KERNELS])
def test_packets_to_pcm(name, flags):
    _compare_pcm(name, None, 97, flags)


# ---- SeekTo (StreamDecoder.cs:562-628): host cursor one packet early + roll forward -------------------------------------
def _seek_case(name, lib_path, positions, gpu_unpack):
    from nvorbis_b200.reader import VorbisReader
    pl = H.packets(name)
    r, pcm, b = H.decoded(name)
    C = r.channels
    with VorbisReader((pl.data, pl.sizes, pl.granules, pl.flags), batch_packets=13, lib_path=lib_path, gpu_unpack=gpu_unpack) as vr:
        assert vr.total_samples == pcm.size // C
        for pos in positions:
            vr.seek_to(pos)
            assert vr.sample_position == pos
            buf = np.zeros(3000 * C, np.float32)
            n = vr.read_samples(buf, 0, buf.size)
            want = pcm[pos * C: pos * C + buf.size]
            assert n == want.size and float(np.abs(buf[:n] - want).max()) <= 1e-5, (name, pos)
        vr.seek_to(pcm.size // C)                          # the very end: nothing left
        assert vr.read_samples(np.zeros(64 * C, np.float32), 0, 64 * C) == 0
        # SeekOrigin forms and the time properties (VorbisReader.cs:215-244, StreamDecoder.cs:562-579): Current SUBTRACTS, like the reference
        from nvorbis_b200.reader import SEEK_CURRENT, SEEK_END
        total = pcm.size // C
        assert abs(vr.total_time - total / vr.sample_rate) < 1e-12
        vr.seek_to(1000, SEEK_END)
        assert vr.sample_position == total - 1000 and abs(vr.time_position - (total - 1000) / vr.sample_rate) < 1e-12
        buf = np.zeros(500 * C, np.float32)
        assert vr.read_samples(buf, 0, buf.size) == buf.size and float(np.abs(buf - pcm[(total - 1000) * C: (total - 500) * C]).max()) <= 1e-5
        vr.seek_to(300, SEEK_CURRENT)                      # SamplePosition - 300
        assert vr.sample_position == total - 500 - 300
        vr.time_position = 0.01
        assert vr.sample_position == int(vr.sample_rate * 0.01)
        vr.sample_position = 77
        assert vr.sample_position == 77 and vr.read_samples(buf, 0, 2 * C) == 2 * C and float(np.abs(buf[: 2 * C] - pcm[77 * C: 79 * C]).max()) <= 1e-5


def test_seek_on_cpu_shim():
    # mono stream with a short first block; positions inside the first blocks, on a block edge, mid-stream, near the end
    _seek_case("1test", H.build_shim(), [0, 1, 127, 128, 700, 5000, 17318 - 2999], gpu_unpack=False)
    _seek_case("1test", H.build_shim(), [4097, 12345], gpu_unpack=True)


def test_host_seek_arithmetic():
    """nvh_seek against the stream-order bookkeeping itself: for every target, cursor + skip reproduce the position."""
    for name in H.FIXTURES:
        pl = H.packets(name)
        hs = hostlib.HostStream(packets=(pl.data, pl.sizes, pl.granules, pl.flags))
        total = hs.total_samples()
        r, pcm, b = H.decoded(name)
        assert total == pcm.size // r.channels
        hb, _ = hs.unpack(hs.n_audio_packets + 1, threads=1)
        f = hb.frames
        # samples each record emits in a decode from the start (first block nothing, failed records drain)
        emitted, prev_tail = [], 0
        have_prev = False
        for i in range(len(f)):
            if f["status"][i] != capi.FRAME_OK:
                emitted.append(prev_tail if have_prev else 0); prev_tail = 0
                continue
            emitted.append(int(f["valid"][i] - f["start"][i]) if have_prev else 0)
            prev_tail = int(f["total"][i] - f["valid"][i]); have_prev = True
        cum = np.concatenate([[0], np.cumsum(emitted)])
        assert cum[-1] == total
        for pos in (1, total // 3, total // 2 + 7, total - 1):
            skip = hs.seek(pos)
            k = int(np.searchsorted(cum, pos, side="right")) - 1          # the record whose output holds the sample
            if f["status"][k] != capi.FRAME_OK:                           # inside the drained tail: the cursor stays on the last packet
                k -= 1
            assert skip == pos - cum[k] and 0 <= skip
        hs.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["2test", "3test", "issue6test"])
def test_seek(name):
    r, pcm, b = H.decoded(name)
    total = pcm.size // r.channels
    _seek_case(name, None, [0, 1, 1023, 1024, 4097, total // 2, total - 2999], gpu_unpack=True)
    _seek_case(name, None, [777, total // 3], gpu_unpack=False)
