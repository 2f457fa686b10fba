"""Shared test helpers: fixtures, oracle -> C-ABI marshalling.  Test infrastructure only."""
from __future__ import annotations

import functools
import os
import subprocess

import numpy as np

from nvorbis_b200 import capi
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURES = ["1test", "2test", "3test", "issue6test"]
SHIM_DIR = os.path.join(ROOT, "tests", "cpu_shim")
SHIM_LIB = os.path.join(SHIM_DIR, "libnvb_cpu_shim.so")


def packets(name: str) -> O.PacketList:
    return O.PacketList.load(os.path.join(GOLDEN, name + ".packets.npz"))


@functools.lru_cache(maxsize=None)
def decoded(name: str, dense: bool = False):
    """(reader, pcm, boundary) of a fixture decoded by the oracle with boundary recording on."""
    r = O.OracleReader(packets(name), record=True, record_dense=dense)
    pcm = r.read_all()
    return r, pcm, r.boundary()


def build_shim() -> str:
    subprocess.check_call(["make", "-C", SHIM_DIR, "-s"])
    return SHIM_LIB


def desc_from_oracle(r: O.OracleReader) -> dict:
    """Setup description (nvorbis_b200.setupio format) from the oracle's parsed headers."""
    cnt = r.counts()
    books = []
    for b in range(r.n_books):
        bk = r.book(b)
        books.append(dict(dims=bk["dims"], entries=bk["entries"], map_type=bk["map_type"], table=bk["table"] if bk["table"].size else None))
    return dict(channels=r.channels, sample_rate=r.sample_rate, block_size=(r.block0, r.block1), books=books,
                floors=[r.floor(i) for i in range(cnt["floors"])], residues=[r.residue(i) for i in range(cnt["residues"])],
                mappings=[r.mapping(i) for i in range(cnt["mappings"])], modes=[r.mode(i) for i in range(cnt["modes"])])


def setup_from_oracle(r: O.OracleReader, **kw) -> capi.Setup:
    """nvb_setup from the oracle's parsed headers (the tables StreamDecoder.LoadBooks leaves behind)."""
    cnt = r.counts()
    books = []
    for b in range(r.n_books):
        bk = r.book(b)
        books.append(dict(dims=bk["dims"], entries=bk["entries"], map_type=bk["map_type"], table=bk["table"]))
    floors = [r.floor(i) for i in range(cnt["floors"])]
    residues = [r.residue(i) for i in range(cnt["residues"])]
    mappings = [r.mapping(i) for i in range(cnt["mappings"])]
    modes = [r.mode(i) for i in range(cnt["modes"])]
    return capi.Setup(r.channels, r.sample_rate, (r.block0, r.block1), books, floors, residues, mappings, modes, **kw)


def batch_from_boundary(b: O.Boundary, post_stride: int, lo: int = 0, hi: int | None = None) -> capi.HostBatch:
    """nvb_batch arrays for boundary records [lo, hi)."""
    hi = len(b.frames) if hi is None else hi
    n, ch = hi - lo, b.channels
    fr = np.zeros(n, capi.FRAME_DTYPE)
    src = b.frames[lo:hi]
    fr["status"] = np.where(src["ok"] != 0, capi.FRAME_OK, capi.FRAME_FAILED)
    fr["mode"] = src["mode"]; fr["window"] = src["windowIndex"]; fr["res_decoded"] = src["resDecoded"]
    fr["exec_mask"] = src["execMask"]; fr["start"] = src["start"]; fr["valid"] = src["valid"]; fr["total"] = src["total"]
    c0 = int(src["classesOff"][0]) if n else 0
    e0 = int(src["entriesOff"][0]) if n else 0
    c1 = int(b.frames["classesOff"][hi]) if hi < len(b.frames) else b.classes.size
    e1 = int(b.frames["entriesOff"][hi]) if hi < len(b.frames) else b.entries.size
    fr["classes_off"] = src["classesOff"] - c0; fr["entries_off"] = src["entriesOff"] - e0; fr["entry_count"] = src["entryCount"]
    posts = np.zeros((n, ch, post_stride), np.int16)
    p64 = b.posts.reshape(-1, ch, 64)[lo:hi]
    posts[:, :, 0] = b.post_counts.reshape(-1, ch)[lo:hi]
    k = min(64, post_stride - 1)
    posts[:, :, 1:1 + k] = p64[:, :, :k]
    assert b.entries.size == 0 or b.entries.max() < 65536
    return capi.HostBatch(fr, posts.reshape(-1), b.classes[c0:c1], b.entries[e0:e1].astype(np.uint16))


def oracle_synth(r: O.OracleReader, b: O.Boundary, lo: int = 0, hi: int | None = None, threads: int = 1):
    """Oracle synthesis of boundary records [lo, hi) from a fresh decoder state."""
    hi = len(b.frames) if hi is None else hi
    fr = b.frames[lo:hi].copy()
    cap = int(fr["total"].astype(np.int64).sum()) + 8192
    return r.synth_batch(fr, b.posts, b.post_counts, b.classes, b.entries, cap, threads=threads)


def _ogg_crc(data: bytes) -> int:
    crc = 0
    for byte in data:
        crc ^= byte << 24
        for _ in range(8):
            crc = ((crc << 1) ^ 0x04c11db7) & 0xFFFFFFFF if crc & 0x80000000 else (crc << 1) & 0xFFFFFFFF
    return crc


def mux_ogg(pk, serial: int = 1, page_payload: int = 4000) -> bytes:
    """A plain Ogg muxer for tests: packets (bytes, granule, flags) -> pages of at most `page_payload` body bytes,
    packets continued across pages when they do not fit.  The granule of a page is that of the last packet completed
    in it (-1 if none); flags bit1 marks the end-of-stream packet."""
    import struct
    pages, seq = [], 0
    segs, body, gran, cont, eos = [], b"", -1, False, False
    last_g = 0

    def flush(continued_next=False):
        nonlocal segs, body, gran, cont, eos, seq
        if not segs:
            return
        hdr = struct.pack("<4sBBqIIIB", b"OggS", 0, (1 if cont else 0) | (2 if seq == 0 else 0) | (4 if eos else 0), gran, serial, seq, 0, len(segs))
        page = bytearray(hdr + bytes(segs) + body)
        page[22:26] = struct.pack("<I", _ogg_crc(bytes(page)))
        pages.append(bytes(page)); seq += 1
        segs, body, gran, cont, eos = [], b"", -1, continued_next, False

    for data, g, fl in pk:
        if fl & 1:
            last_g = g
        pos = 0
        lacing = [255] * (len(data) // 255) + [len(data) % 255]
        for lv in lacing:
            if len(segs) == 255 or len(body) + lv > page_payload:
                flush(continued_next=True)
            segs.append(lv); body += data[pos:pos + lv]; pos += lv
        gran = last_g if (fl & 1) else gran
        if fl & 2:
            eos = True
        if fl & 1:
            flush()
    flush()
    return b"".join(pages)


def oracle_records(hb: capi.HostBatch, desc: dict):
    """HostBatch (C-ABI layout) -> the oracle's SynthFrame arrays, for any setup description."""
    n, C = len(hb.frames), desc["channels"]
    fr = np.zeros(n, O.SYNTH_FRAME_DTYPE)
    f = hb.frames
    stride = hb.posts.size // max(n * C, 1)
    fr["ok"] = f["status"] == 0; fr["mode"] = f["mode"]; fr["windowIndex"] = f["window"]
    fr["start"] = f["start"]; fr["valid"] = f["valid"]; fr["total"] = f["total"]; fr["execMask"] = f["exec_mask"]
    fr["resDecoded"] = f["res_decoded"]
    for i in range(n):
        m = desc["modes"][int(f["mode"][i])]
        rs = desc["residues"][desc["mappings"][m["mapping"]]["residue"]]
        N = desc["block_size"][1 if m["block_flag"] else 0]
        span = N * C // 2 if rs["type"] == 2 else N // 2
        fr["resStreams"][i] = 1 if rs["type"] == 2 else C
        fr["resPartitions"][i] = max(min(rs["end"], span) - rs["begin"], 0) // rs["partition_size"]
    fr["postsOff"] = np.arange(n, dtype=np.int64) * C * 64; fr["postCountOff"] = np.arange(n, dtype=np.int64) * C
    fr["classesOff"] = f["classes_off"]; fr["entriesOff"] = f["entries_off"]; fr["entryCount"] = f["entry_count"]
    p = hb.posts.reshape(n, C, stride)
    posts = np.zeros((n, C, 64), np.int32)
    posts[:, :, :stride - 1] = p[:, :, 1:]
    return fr, posts.reshape(-1), p[:, :, 0].astype(np.int32).reshape(-1), hb.classes, hb.entries.astype(np.int32)


def oracle_decode_records(reader: O.OracleReader, hb: capi.HostBatch, desc: dict, threads: int = 1):
    fr, posts, pc, cls, ent = oracle_records(hb, desc)
    cap = int(hb.frames["total"].astype(np.int64).sum()) + 8192
    f0 = getattr(hb, "floor0", None)
    stride = 0 if f0 is None else f0.size // (len(hb.frames) * desc["channels"])
    return reader.synth_batch(fr, posts, pc, cls, ent, cap, threads=threads, floor0=f0, floor0_stride=stride)
