"""ctypes binding of the CPU oracle (oracle/libnvorbis_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (nvorbis_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnvorbis_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nvorbis_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class SynthFrame(C.Structure):
    _fields_ = [
        ("ok", C.c_int32), ("mode", C.c_int32), ("windowIndex", C.c_int32),
        ("start", C.c_int32), ("valid", C.c_int32), ("total", C.c_int32),
        ("execMask", C.c_uint32),
        ("resDecoded", C.c_int32), ("resStreams", C.c_int32), ("resPartitions", C.c_int32),
        ("postsOff", C.c_int64), ("postCountOff", C.c_int64), ("classesOff", C.c_int64),
        ("entriesOff", C.c_int64), ("entryCount", C.c_int32), ("pad", C.c_int32),
    ]


SYNTH_FRAME_DTYPE = np.dtype([
    ("ok", "<i4"), ("mode", "<i4"), ("windowIndex", "<i4"), ("start", "<i4"), ("valid", "<i4"), ("total", "<i4"),
    ("execMask", "<u4"), ("resDecoded", "<i4"), ("resStreams", "<i4"), ("resPartitions", "<i4"),
    ("postsOff", "<i8"), ("postCountOff", "<i8"), ("classesOff", "<i8"), ("entriesOff", "<i8"),
    ("entryCount", "<i4"), ("pad", "<i4"),
], align=True)
assert SYNTH_FRAME_DTYPE.itemsize == C.sizeof(SynthFrame)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_open.restype = C.c_void_p
        L.orc_open.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_close.argtypes = [C.c_void_p]
        L.orc_open_packets.restype = C.c_void_p
        L.orc_open_packets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_floor_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_residue_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_mapping_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_mode_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_info.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_options.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_read_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_has_clipped.argtypes = [C.c_void_p]
        L.orc_packet_count.restype = C.c_int64
        L.orc_packet_count.argtypes = [C.c_void_p]
        L.orc_packet_size.restype = C.c_int64
        L.orc_packet_size.argtypes = [C.c_void_p, C.c_int64]
        L.orc_packet_get.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_rec_count.restype = C.c_int64
        L.orc_rec_count.argtypes = [C.c_void_p]
        L.orc_rec_meta.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_rec_floor1.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_rec_residue.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_rec_dense.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_book_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_book_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_book_lengths.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_mode_window.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_mdct_reverse.argtypes = [C.c_void_p, C.c_int]
        L.orc_mdct_tables.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_calc_window.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_floor1_apply.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_inverse_couple.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_clip.restype = C.c_float
        L.orc_clip.argtypes = [C.c_float, C.c_void_p]
        L.orc_inverse_db.restype = C.c_float
        L.orc_inverse_db.argtypes = [C.c_int]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_floor0_payload.argtypes = [C.c_void_p, C.c_int]
        L.orc_synth_batch.restype = C.c_int64
        L.orc_synth_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleError(RuntimeError):
    pass


@dataclass
class Boundary:
    """Per-frame boundary records of one decoded stream, flattened into batch arrays
    (the same information the product's unpacker hands to the GPU)."""
    channels: int
    frames: np.ndarray                     # SYNTH_FRAME_DTYPE
    block_size: np.ndarray                 # int32 per frame
    valid_untrimmed: np.ndarray            # int32 per frame
    no_exec_mask: np.ndarray               # uint32 per frame
    posts: np.ndarray                      # int32 [n_frames*channels*64]
    post_counts: np.ndarray                # int32 [n_frames*channels]
    classes: np.ndarray                    # uint8
    entries: np.ndarray                    # int32
    spectrum: list = field(default_factory=list)   # per frame [ch, N/2] (if dense recording)
    block: list = field(default_factory=list)      # per frame [ch, N]


@dataclass
class PacketList:
    """Demuxed packets of one logical stream (what an IPacketProvider hands out, Contracts/IPacketProvider.cs):
    data = all packet bytes back to back; flags bit0 = has granule, bit1 = end of stream, bit2 = resync."""
    data: np.ndarray       # uint8
    sizes: np.ndarray      # int64
    granules: np.ndarray   # int64
    flags: np.ndarray      # uint8

    @staticmethod
    def from_packets(pk):
        data = np.frombuffer(b"".join(p[0] for p in pk), dtype=np.uint8).copy() if pk else np.zeros(0, np.uint8)
        if data.size == 0:
            data = np.zeros(1, np.uint8)
        return PacketList(data, np.array([len(p[0]) for p in pk], np.int64), np.array([p[2] for p in pk], np.int64),
                          np.array([(1 if p[1] else 0) | (2 if p[3] else 0) | (4 if p[4] else 0) for p in pk], np.uint8))

    def save(self, path):
        np.savez_compressed(path, data=self.data, sizes=self.sizes, granules=self.granules, flags=self.flags)

    @staticmethod
    def load(path):
        z = np.load(path)
        return PacketList(np.ascontiguousarray(z["data"]), np.ascontiguousarray(z["sizes"]), np.ascontiguousarray(z["granules"]),
                          np.ascontiguousarray(z["flags"]))


class OracleReader:
    """Mirror of NVorbis.VorbisReader restricted to what TestApp/Program.cs uses."""

    def __init__(self, data, clip: bool = True, record: bool = False, record_dense: bool = False):
        """data: the bytes of an Ogg file, or a PacketList (already-demuxed packets)."""
        L = lib()
        if isinstance(data, PacketList):
            self._data = data
            self._h = L.orc_open_packets(_ptr(data.data), _ptr(data.sizes), _ptr(data.granules), _ptr(data.flags), len(data.sizes))
        else:
            self._data = bytes(data)
            self._h = L.orc_open(self._data, len(self._data))
        if not self._h:
            raise OracleError(L.orc_last_error().decode())
        info = np.zeros(8, dtype=np.int64)
        L.orc_info(self._h, _ptr(info))
        self.channels, self.sample_rate, self.block0, self.block1 = (int(x) for x in info[:4])
        self.n_packets, self.n_modes, self.n_books, self.mode_field_bits = (int(x) for x in info[4:8])
        L.orc_set_options(self._h, int(clip), int(record or record_dense), int(record_dense))

    def close(self):
        if self._h:
            lib().orc_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # VorbisReader.ReadSamples(float[] buffer, int offset, int count)   VorbisReader.cs:336-345
    def read_samples(self, buffer: np.ndarray, offset: int, count: int) -> int:
        assert buffer.dtype == np.float32 and buffer.flags.c_contiguous
        assert offset >= 0 and offset + count <= buffer.size
        n = lib().orc_read_samples(self._h, _ptr(buffer), offset, count)
        if n < 0:
            raise OracleError(lib().orc_last_error().decode())
        return n

    def read_all(self, chunk_seconds: float = 4.0) -> np.ndarray:
        """TestApp/Program.cs:21-26 loop: 4-second chunks until ReadSamples returns 0."""
        buf = np.zeros(int(self.sample_rate * chunk_seconds) * self.channels, dtype=np.float32)
        out = []
        while True:
            n = self.read_samples(buf, 0, buf.size)
            if n <= 0:
                break
            out.append(buf[:n].copy())
        return np.concatenate(out) if out else np.zeros(0, dtype=np.float32)

    @property
    def has_clipped(self) -> bool:
        return bool(lib().orc_has_clipped(self._h))

    # -- packets ---------------------------------------------------------------------
    def packets(self):
        L = lib()
        out = []
        for i in range(L.orc_packet_count(self._h)):
            n = L.orc_packet_size(self._h, i)
            b = np.zeros(max(n, 1), dtype=np.uint8)
            meta = np.zeros(4, dtype=np.int64)
            L.orc_packet_get(self._h, i, _ptr(b), _ptr(meta))
            out.append((bytes(b[:n]), bool(meta[0]), int(meta[1]), bool(meta[2]), bool(meta[3])))
        return out

    # -- setup tables ------------------------------------------------------------------
    def book(self, b: int):
        info = np.zeros(4, dtype=np.int64)
        if lib().orc_book_info(self._h, b, _ptr(info)) != 0:
            raise IndexError(b)
        table = np.zeros(int(info[3]), dtype=np.float32)
        if table.size:
            lib().orc_book_table(self._h, b, _ptr(table))
        lengths = np.zeros(int(info[1]), dtype=np.int32)
        if lengths.size:
            lib().orc_book_lengths(self._h, b, _ptr(lengths))
        return dict(dims=int(info[0]), entries=int(info[1]), map_type=int(info[2]), table=table, lengths=lengths)

    def counts(self):
        out = np.zeros(4, np.int64)
        lib().orc_counts(self._h, _ptr(out))
        return dict(floors=int(out[0]), residues=int(out[1]), mappings=int(out[2]), modes=int(out[3]))

    def floor(self, i: int):
        o = np.zeros(260, np.int32)
        if lib().orc_floor_info(self._h, i, _ptr(o)) != 0:
            raise IndexError(i)
        n = int(o[1])
        if int(o[0]) == 0:
            return dict(type=0, n_posts=0, order=int(o[4]), rate=int(o[5]), bark_map_size=int(o[6]), amp_bits=int(o[7]), amp_ofs=int(o[8]),
                        books=o[10:10 + int(o[9])].copy())
        return dict(type=int(o[0]), n_posts=n, multiplier=int(o[2]), range=int(o[3]), x_list=o[4:4 + n].copy(), l_neigh=o[68:68 + n].copy(),
                    h_neigh=o[132:132 + n].copy(), sort_idx=o[196:196 + n].copy())

    def residue(self, i: int):
        o = np.zeros(8 + 64 + 512, np.int32)
        if lib().orc_residue_info(self._h, i, _ptr(o)) != 0:
            raise IndexError(i)
        nc = int(o[4])
        return dict(type=int(o[0]), begin=int(o[1]), end=int(o[2]), partition_size=int(o[3]), classifications=nc, max_stages=int(o[5]),
                    class_book=int(o[6]), channels=int(o[7]), cascade=o[8:8 + nc].copy(), books=o[72:72 + 512].reshape(64, 8)[:nc].copy())

    def mapping(self, i: int):
        o = np.zeros(516, np.int32)
        if lib().orc_mapping_info(self._h, i, _ptr(o)) != 0:
            raise IndexError(i)
        n = int(o[0])
        return dict(n_coupling=n, n_submaps=int(o[1]), floor=int(o[2]), residue=int(o[3]), magnitude=o[4:4 + n].copy(), angle=o[260:260 + n].copy())

    def mode(self, i: int):
        o = np.zeros(3, np.int32)
        if lib().orc_mode_info(self._h, i, _ptr(o)) != 0:
            raise IndexError(i)
        return dict(block_flag=int(o[0]), mapping=int(o[1]), block_size=int(o[2]))

    def mode_window(self, m: int, w: int) -> np.ndarray:
        n = lib().orc_mode_window(self._h, m, w, None)
        if n < 0:
            raise IndexError((m, w))
        out = np.zeros(n, dtype=np.float32)
        lib().orc_mode_window(self._h, m, w, _ptr(out))
        return out

    # -- boundary records --------------------------------------------------------------
    def boundary(self) -> Boundary:
        L = lib()
        n = L.orc_rec_count(self._h)
        ch = self.channels
        frames = np.zeros(n, dtype=SYNTH_FRAME_DTYPE)
        block_size = np.zeros(n, dtype=np.int32)
        valid_untrimmed = np.zeros(n, dtype=np.int32)
        noexec = np.zeros(n, dtype=np.uint32)
        posts = np.zeros(n * ch * 64, dtype=np.int32)
        post_counts = np.zeros(n * ch, dtype=np.int32)
        cls_parts, ent_parts, spectrum, block = [], [], [], []
        cls_off = ent_off = 0
        meta = np.zeros(16, dtype=np.int64)
        for i in range(n):
            L.orc_rec_meta(self._h, i, _ptr(meta))
            f = frames[i]
            f["ok"], f["mode"], f["windowIndex"] = meta[0], meta[1], meta[3]
            f["start"], f["valid"], f["total"] = meta[4], meta[7], meta[6]
            block_size[i] = meta[2]
            valid_untrimmed[i] = meta[5]
            f["execMask"] = meta[8]
            noexec[i] = meta[9]
            f["resDecoded"], f["resStreams"], f["resPartitions"] = meta[10], meta[11], meta[12]
            f["postsOff"], f["postCountOff"] = i * ch * 64, i * ch
            L.orc_rec_floor1(self._h, i, _ptr(posts[i * ch * 64:]), _ptr(post_counts[i * ch:]))
            ne, nc = int(meta[13]), int(meta[14])
            c = np.zeros(max(nc, 1), dtype=np.uint8)
            e = np.zeros(max(ne, 1), dtype=np.int32)
            L.orc_rec_residue(self._h, i, _ptr(c), _ptr(e))
            f["classesOff"], f["entriesOff"], f["entryCount"] = cls_off, ent_off, ne
            cls_parts.append(c[:nc]); ent_parts.append(e[:ne])
            cls_off += nc; ent_off += ne
            if meta[15]:
                N = int(meta[2])
                s = np.zeros((ch, N // 2), dtype=np.float32)
                b = np.zeros((ch, N), dtype=np.float32)
                L.orc_rec_dense(self._h, i, _ptr(s), _ptr(b))
                spectrum.append(s); block.append(b)
            else:
                spectrum.append(None); block.append(None)
        classes = np.concatenate(cls_parts) if cls_parts else np.zeros(0, np.uint8)
        entries = np.concatenate(ent_parts) if ent_parts else np.zeros(0, np.int32)
        return Boundary(ch, frames, block_size, valid_untrimmed, noexec, posts, post_counts,
                        np.ascontiguousarray(classes), np.ascontiguousarray(entries), spectrum, block)

    # -- batch synthesis from boundary arrays -----------------------------------------
    def synth_batch(self, frames: np.ndarray, posts: np.ndarray, post_counts: np.ndarray, classes: np.ndarray,
                    entries: np.ndarray, pcm_cap_per_channel: int, threads: int = 1, floor0: np.ndarray | None = None, floor0_stride: int = 0):
        L = lib()
        if floor0 is not None:
            floor0 = np.ascontiguousarray(floor0, dtype=np.float32)
            L.orc_set_floor0_payload(_ptr(floor0), int(floor0_stride))
        else:
            L.orc_set_floor0_payload(None, 0)
        frames = np.ascontiguousarray(frames, dtype=SYNTH_FRAME_DTYPE)
        posts = np.ascontiguousarray(posts, dtype=np.int32)
        post_counts = np.ascontiguousarray(post_counts, dtype=np.int32)
        classes = np.ascontiguousarray(classes, dtype=np.uint8)
        entries = np.ascontiguousarray(entries, dtype=np.int32)
        if classes.size == 0:
            classes = np.zeros(1, np.uint8)
        if entries.size == 0:
            entries = np.zeros(1, np.int32)
        pcm = np.zeros(pcm_cap_per_channel * self.channels, dtype=np.float32)
        clipped = C.c_int(0)
        L.orc_set_threads(threads)
        n = L.orc_synth_batch(self._h, _ptr(frames), len(frames), _ptr(posts), _ptr(post_counts), _ptr(classes),
                              _ptr(entries), _ptr(pcm), pcm_cap_per_channel, C.byref(clipped))
        L.orc_set_threads(1)
        L.orc_set_floor0_payload(None, 0)
        if n < 0:
            raise OracleError(L.orc_last_error().decode())
        return pcm[: n * self.channels], bool(clipped.value)


# ---- stand-alone helpers ---------------------------------------------------------------
def mdct_reverse(x: np.ndarray) -> np.ndarray:
    """Mdct.Reverse: x = N/2 spectral floats -> N time-domain floats (unwindowed)."""
    n = x.size * 2
    buf = np.zeros(n, dtype=np.float32)
    buf[: n // 2] = x
    if lib().orc_mdct_reverse(_ptr(buf), n) != 0:
        raise OracleError(lib().orc_last_error().decode())
    return buf


def mdct_tables(n: int):
    a = np.zeros(n // 2, np.float32); b = np.zeros(n // 2, np.float32)
    c = np.zeros(n // 4, np.float32); br = np.zeros(n // 8, np.uint16)
    lib().orc_mdct_tables(n, _ptr(a), _ptr(b), _ptr(c), _ptr(br))
    return a, b, c, br


def calc_window(prev: int, cur: int, nxt: int) -> np.ndarray:
    out = np.zeros(cur, np.float32)
    lib().orc_calc_window(prev, cur, nxt, _ptr(out))
    return out


def inverse_couple(mag: np.ndarray, ang: np.ndarray):
    m = np.ascontiguousarray(mag, np.float32).copy(); a = np.ascontiguousarray(ang, np.float32).copy()
    lib().orc_inverse_couple(_ptr(m), _ptr(a), m.size)
    return m, a


def inverse_db_table() -> np.ndarray:
    return np.array([lib().orc_inverse_db(i) for i in range(256)], dtype=np.float32)
