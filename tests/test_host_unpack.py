"""The host half of the split decoder (libnvorbis_host.so) against the oracle's boundary records: the product's
own Ogg demux / header parser / Huffman + floor + residue unpacker must hand the GPU exactly the integers the
reference's bit-reading code produces.  CPU only."""
import numpy as np
import pytest

import helpers as H
from nvorbis_b200 import capi, hostlib


def _stream(name):
    pl = H.packets(name)
    return hostlib.HostStream(packets=(pl.data, pl.sizes, pl.granules, pl.flags))


@pytest.mark.parametrize("name", H.FIXTURES)
@pytest.mark.parametrize("threads", [1, 5])
def test_boundary_records_match_oracle(name, threads):
    r, pcm, b = H.decoded(name)
    hs = _stream(name)
    hb, eos = hs.unpack(10 ** 9, threads=threads)
    ref = H.batch_from_boundary(b, hs.post_stride)
    assert eos and len(hb.frames) == len(ref.frames)
    for k in capi.FRAME_DTYPE.names:
        np.testing.assert_array_equal(hb.frames[k], ref.frames[k], err_msg=k)
    np.testing.assert_array_equal(hb.posts, ref.posts)
    np.testing.assert_array_equal(hb.classes, ref.classes)
    np.testing.assert_array_equal(hb.entries, ref.entries)


@pytest.mark.parametrize("name", H.FIXTURES)
def test_chunked_unpack_keeps_position_bookkeeping(name):
    """Unpacking in ragged chunks gives the same records (the EOS trim needs the running sample position)."""
    hs = _stream(name)
    whole, _ = hs.unpack(10 ** 9, threads=2)
    hs.rewind()
    frames = []
    for n in [1, 2, 7, 64, 100, 10 ** 9]:
        hb, eos = hs.unpack(n, threads=3)
        frames.append(hb.frames)
        if eos:
            break
    got = np.concatenate(frames)
    for k in ("status", "mode", "window", "exec_mask", "start", "valid", "total", "entry_count"):
        np.testing.assert_array_equal(got[k], whole.frames[k], err_msg=k)


@pytest.mark.parametrize("name", H.FIXTURES)
def test_setup_tables_match_oracle(name):
    r, _, _ = H.decoded(name)
    hs = _stream(name)
    view, owner = hs.setup(), H.setup_from_oracle(r)        # keep the owners of the pointed-to arrays alive
    s, want = view.struct, owner.struct
    assert (s.channels, s.sample_rate, s.block_size[0], s.block_size[1]) == (want.channels, want.sample_rate, want.block_size[0], want.block_size[1])
    assert (s.n_books, s.n_floors, s.n_residues, s.n_mappings, s.n_modes) == (want.n_books, want.n_floors, want.n_residues, want.n_mappings, want.n_modes)
    assert s.n_vq_floats == want.n_vq_floats
    np.testing.assert_array_equal(np.ctypeslib.as_array(s.vq_floats, (s.n_vq_floats,)), np.ctypeslib.as_array(want.vq_floats, (want.n_vq_floats,)))
    import ctypes as C
    for i in range(s.n_books):
        a, b = s.books[i], want.books[i]
        assert (a.dims, a.entries, a.map_type, a.table_off) == (b.dims, b.entries, b.map_type, b.table_off)
    for i in range(s.n_floors):
        assert bytes(C.string_at(C.addressof(s.floors[i]), C.sizeof(capi.Floor))) == bytes(C.string_at(C.addressof(want.floors[i]), C.sizeof(capi.Floor)))
    for i in range(s.n_residues):
        a, b = s.residues[i], want.residues[i]
        assert (a.type, a.begin, a.end, a.partition_size, a.classifications, a.max_stages) == (b.type, b.begin, b.end, b.partition_size, b.classifications, b.max_stages)
        for c in range(a.classifications):
            assert a.cascade[c] == b.cascade[c]
            for st in range(a.max_stages):
                if (a.cascade[c] >> st) & 1:
                    assert a.books[c][st] == b.books[c][st]
    for i in range(s.n_mappings):
        a, b = s.mappings[i], want.mappings[i]
        assert (a.n_coupling, a.n_submaps, a.floor, a.residue) == (b.n_coupling, b.n_submaps, b.floor, b.residue)
        assert list(a.magnitude[:a.n_coupling]) == list(b.magnitude[:b.n_coupling]) and list(a.angle[:a.n_coupling]) == list(b.angle[:b.n_coupling])
    for i in range(s.n_modes):
        assert (s.modes[i].block_flag, s.modes[i].mapping) == (want.modes[i].block_flag, want.modes[i].mapping)


def test_ogg_demux_roundtrip():
    """Packets re-muxed into Ogg pages (lacing, continued packets across pages, a trailing empty EOS page) demux
    back to the same packets, granules and end-of-stream flags."""
    pl = H.packets("1test")
    pk = []
    off = 0
    for sz, g, fl in zip(pl.sizes, pl.granules, pl.flags):
        pk.append((bytes(pl.data[off:off + sz]), int(g), int(fl))); off += int(sz)
    ogg = H.mux_ogg(pk, serial=0x1234, page_payload=700)
    hs = hostlib.HostStream(data=ogg)
    assert hs.info.n_packets == len(pk)
    for i, (data, g, fl) in enumerate(pk):
        d2, g2, f2 = hs.packet(i)
        assert d2 == data
    hb, eos = hs.unpack(10 ** 9)
    ref, _ = _stream("1test").unpack(10 ** 9)
    np.testing.assert_array_equal(hb.entries, ref.entries)
    np.testing.assert_array_equal(hb.frames["valid"][:-1], ref.frames["valid"][:-1])


def test_malformed_headers_are_rejected():
    pl = H.packets("1test")
    data = pl.data.copy()
    data[int(pl.sizes[0]) + int(pl.sizes[1]) + 3] ^= 0xFF          # corrupt the setup header signature
    with pytest.raises(hostlib.HostError):
        hostlib.HostStream(packets=(data, pl.sizes, pl.granules, pl.flags))
    with pytest.raises(hostlib.HostError):
        hostlib.HostStream(data=b"not an ogg file at all")


def test_multi_stream_container():
    """Two logical streams in one container (pages of different serial numbers interleaved, as a multiplexed file has
    them): nvh_ogg_stream_count sees both, nvh_open_ogg_stream(i) demuxes stream i alone -- same packets, same unpacked
    records as the stream on its own (VorbisReader.Streams / SwitchStreams, VorbisReader.cs:96-150)."""
    import re
    from nvorbis_b200 import hostlib
    blobs = []
    for name, serial in (("1test", 0x1111), ("3test", 0x2222)):
        pl = H.packets(name)
        off = np.concatenate([[0], np.cumsum(pl.sizes)])
        pk = [(pl.data[off[i]:off[i + 1]].tobytes(), int(pl.granules[i]), int(pl.flags[i])) for i in range(len(pl.sizes))]
        blobs.append(H.mux_ogg(pk, serial=serial, page_payload=1500))
    # split both byte streams into pages and interleave them two to one
    pages = [[m.start() for m in re.finditer(b"OggS", b)] + [len(b)] for b in blobs]
    cut = [[b[p[i]:p[i + 1]] for i in range(len(p) - 1)] for b, p in zip(blobs, pages)]
    mixed, ia, ib = [], 0, 0
    while ia < len(cut[0]) or ib < len(cut[1]):
        mixed += cut[1][ib:ib + 2]; ib += 2
        mixed += cut[0][ia:ia + 1]; ia += 1
    data = b"".join(mixed)
    assert hostlib.ogg_stream_count(data) == 2
    for idx, name in ((1, "1test"), (0, "3test")):                 # 3test's first page comes first in the mix
        alone = hostlib.HostStream(data=blobs[0 if name == "1test" else 1])
        both = hostlib.HostStream(data=data, stream_index=idx)
        assert (both.channels, both.n_audio_packets) == (alone.channels, alone.n_audio_packets)
        a, _ = alone.unpack(alone.n_audio_packets + 1, threads=1)
        b, _ = both.unpack(both.n_audio_packets + 1, threads=1)
        for arr_a, arr_b in ((a.frames, b.frames), (a.posts, b.posts), (a.classes, b.classes), (a.entries, b.entries)):
            np.testing.assert_array_equal(arr_a, arr_b)
    with pytest.raises(hostlib.HostError):
        hostlib.HostStream(data=data, stream_index=2)
    # the reader mirror on the multiplexed container: StreamCount / StreamIndex / SwitchStreams (on the CUDA-on-CPU shim here)
    from nvorbis_b200.reader import VorbisReader
    alone = []
    for blob in blobs:                                              # each remuxed stream on its own (the remux moves the end-of-stream granule: not the original file's trim)
        with VorbisReader(blob, batch_packets=64, lib_path=H.build_shim(), gpu_unpack=False) as one:
            alone.append(one.read_all(chunk_seconds=0.5))
    with VorbisReader(data, batch_packets=64, lib_path=H.build_shim(), gpu_unpack=False) as vr:
        assert vr.stream_count == 2 and vr.stream_index == 0 and vr.channels == 2
        first = vr.read_all(chunk_seconds=0.5)
        np.testing.assert_array_equal(first, alone[1])
        assert float(np.abs(first[:100000] - H.decoded("3test")[1][:100000]).max()) <= 1e-5
        assert vr.switch_streams(1) is True and vr.stream_index == 1 and vr.channels == 1       # stereo 3test -> mono 1test: the properties differ
        second = vr.read_all(chunk_seconds=0.5)
        np.testing.assert_array_equal(second, alone[0])
        assert vr.switch_streams(1) is False
        with pytest.raises(IndexError):
            vr.switch_streams(2)


def test_forward_only_feed_equals_whole_file():
    """nvh_open_forward / nvh_feed: the container arrives in ragged chunks (pages and continued packets cut anywhere);
    unpacking whatever is available after every chunk yields, concatenated, exactly the records of the whole-file open."""
    from nvorbis_b200 import hostlib
    for name, payload in (("1test", 300), ("3test", 1100)):
        pl = H.packets(name)
        off = np.concatenate([[0], np.cumsum(pl.sizes)])
        pk = [(pl.data[off[i]:off[i + 1]].tobytes(), int(pl.granules[i]), int(pl.flags[i])) for i in range(len(pl.sizes))]
        ogg = H.mux_ogg(pk, serial=77, page_payload=payload)          # small pages: many packets continue across pages
        whole = hostlib.HostStream(data=ogg)
        want, _ = whole.unpack(whole.n_audio_packets + 1, threads=1)
        fwd = hostlib.HostStream(forward=True)
        rng = np.random.default_rng(5)
        pos, frames, posts, classes, entries, eos = 0, [], [], [], [], False
        while pos < len(ogg):
            n = int(rng.integers(1, 4000))
            avail = fwd.feed(ogg[pos:pos + n], end_of_input=pos + n >= len(ogg)); pos += n
            if avail == 0:
                continue
            hb, eos = fwd.unpack(1 << 20, threads=1)
            f = hb.frames.copy()
            f["classes_off"] += sum(c.size for c in classes); f["entries_off"] += sum(e.size for e in entries)
            frames.append(f); posts.append(hb.posts); classes.append(hb.classes); entries.append(hb.entries)
        assert eos and fwd.n_audio_packets == whole.n_audio_packets
        np.testing.assert_array_equal(np.concatenate(frames), want.frames)
        np.testing.assert_array_equal(np.concatenate(posts), want.posts)
        np.testing.assert_array_equal(np.concatenate(classes), want.classes)
        np.testing.assert_array_equal(np.concatenate(entries), want.entries)
        with pytest.raises(hostlib.HostError):
            fwd.feed(b"x")                                              # the input has ended


# ---- malformed setup headers: the parser must refuse them, not overflow / hang / divide by zero ----------
def _setup_with_books(book_writer, n_books=1):
    import vorbis_headers as VH
    w = VH.BitWriter()
    for b in b"\x05vorbis":
        w.put(b, 8)
    w.put(n_books - 1, 8)
    book_writer(w)
    return w


def _open_headers(setup_bytes):
    import vorbis_headers as VH
    from nvorbis_b200 import hostlib
    pk = [VH.id_header(2, 44100, 256, 2048), VH.comment_header(), setup_bytes]
    data = np.frombuffer(b"".join(pk), np.uint8).copy()
    return hostlib.HostStream(packets=(data, np.array([len(p) for p in pk], np.int64), np.zeros(3, np.int64), np.zeros(3, np.uint8)))


def test_ordered_codebook_with_lengths_above_32_is_rejected():
    """First length 32, one entry per group: the second group would be 33 bits long (stack overflow in the decoder builder)."""
    from nvorbis_b200 import hostlib

    def book(w):
        w.put(0x564342, 24); w.put(1, 16); w.put(4, 24)
        w.put(1, 1)                                   # ordered
        w.put(31, 5)                                  # first length 32
        w.put(1, 3); w.put(1, 2); w.put(1, 2); w.put(1, 1)      # one entry per length: 32, 33, 34, 35
        w.put(0, 4)
    with pytest.raises(hostlib.HostError, match="length above 32|status -7"):
        _open_headers(_setup_with_books(book).done())


def test_truncated_ordered_codebook_does_not_hang():
    """The packet ends inside the ordered-length list: every count reads as 0 and the reference's loop would never advance."""
    import signal
    from nvorbis_b200 import hostlib

    def book(w):
        w.put(0x564342, 24); w.put(1, 16); w.put(1000, 24)
        w.put(1, 1); w.put(0, 5)                      # ordered, first length 1 ... and nothing more

    def on_alarm(*_):
        raise AssertionError("the header parser hangs on a truncated ordered codebook")
    old = signal.signal(signal.SIGALRM, on_alarm); signal.alarm(10)
    try:
        with pytest.raises(hostlib.HostError):
            _open_headers(_setup_with_books(book).done())
    finally:
        signal.alarm(0); signal.signal(signal.SIGALRM, old)


def test_value_book_without_dimensions_is_rejected():
    """dims == 0 with a lookup table: Residue0.cs:183 would divide by it per packet."""
    from nvorbis_b200 import hostlib
    import vorbis_headers as VH

    def book(w):
        w.put(0x564342, 24); w.put(0, 16); w.put(4, 24)        # dims 0
        w.put(0, 1); w.put(0, 1)
        for _ in range(4):
            w.put(1, 5)
        w.put(2, 4)                                             # lookup type 2
        w.put(VH.float32_pack(-1.0), 32); w.put(VH.float32_pack(0.5), 32); w.put(3, 4); w.put(0, 1)
    with pytest.raises(hostlib.HostError):
        _open_headers(_setup_with_books(book).done())
