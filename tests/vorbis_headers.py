"""A small Vorbis I header WRITER for tests: builds identification / comment / setup header packets for synthetic
stream setups (block sizes, channel counts, residue types, codebook lookup types the reference's four fixtures do not
reach).  Both the oracle and the product's host library parse the result; the synthesis inputs are then generated
as boundary records directly (no audio packet encoding is needed).  Test infrastructure only."""
from __future__ import annotations

import numpy as np


class BitWriter:
    """LSB-first bit packer (the inverse of DataPacket.ReadBits, DataPacket.cs:149-283)."""

    def __init__(self):
        self.bytes = bytearray(); self.acc = 0; self.n = 0

    def put(self, value: int, bits: int):
        assert 0 <= value < (1 << bits) or bits == 0, (value, bits)
        self.acc |= value << self.n; self.n += bits
        while self.n >= 8:
            self.bytes.append(self.acc & 0xFF); self.acc >>= 8; self.n -= 8

    def done(self) -> bytes:
        if self.n:
            self.bytes.append(self.acc & 0xFF); self.acc = 0; self.n = 0
        return bytes(self.bytes)


def ilog(x: int) -> int:
    n = 0
    while x > 0:
        n += 1; x >>= 1
    return n


def float32_pack(v: float) -> int:
    """Vorbis float32 (Utils.ConvertFromVorbisFloat32, Utils.cs:45-59): mantissa * 2^(exp-788), 21-bit mantissa."""
    if v == 0:
        return 0
    sign = 0x80000000 if v < 0 else 0
    v = abs(v)
    exp = int(np.floor(np.log2(v))) - 20
    mant = int(round(v / 2.0 ** exp))
    while mant >= (1 << 21):
        mant >>= 1; exp += 1
    return sign | ((exp + 788) << 21) | mant


def id_header(channels: int, rate: int, bs0: int, bs1: int) -> bytes:
    w = BitWriter()
    for b in b"\x01vorbis":
        w.put(b, 8)
    w.put(0, 32); w.put(channels, 8); w.put(rate, 32); w.put(0, 32); w.put(0, 32); w.put(0, 32)
    w.put(ilog(bs0) - 1, 4); w.put(ilog(bs1) - 1, 4); w.put(1, 1)
    return w.done()


def comment_header() -> bytes:
    w = BitWriter()
    for b in b"\x03vorbis":
        w.put(b, 8)
    w.put(0, 32); w.put(0, 32); w.put(1, 1)
    return w.done()


def write_codebook(w: BitWriter, dims: int, entries: int, lookup: int = 0, vmin: float = -1.0, delta: float = 0.25,
                   value_bits: int = 4, sequence_p: bool = False, mults=None, length: int | None = None):
    """All codewords get the same length (an under-full tree is legal); lookup 0 / 1 / 2."""
    w.put(0x564342, 24); w.put(dims, 16); w.put(entries, 24)
    w.put(0, 1)                      # not ordered
    w.put(0, 1)                      # not sparse
    L = length or max(1, ilog(entries - 1))
    for _ in range(entries):
        w.put(L - 1, 5)
    w.put(lookup, 4)
    if lookup:
        w.put(float32_pack(vmin), 32); w.put(float32_pack(delta), 32)
        w.put(value_bits - 1, 4); w.put(1 if sequence_p else 0, 1)
        if lookup == 1:
            r = int(np.floor(np.exp(np.log(entries) / dims)))
            if np.floor((r + 1) ** dims) <= entries:
                r += 1
            count = r
        else:
            count = entries * dims
        mults = list(mults) if mults is not None else [(i * 7 + 3) % (1 << value_bits) for i in range(count)]
        assert len(mults) == count
        for m in mults:
            w.put(int(m), value_bits)


def write_floor1(w: BitWriter, xs, multiplier: int, rangebits: int, class_dims, master_book: int, sub_book: int):
    """Floor 1 with one partition class per group of posts: xs = list of partitions, each a list of x values."""
    w.put(len(xs), 5)
    for ci in range(len(xs)):
        w.put(ci, 4)                                           # partition class = index
    for ci, part in enumerate(xs):
        w.put(len(part) - 1, 3)                                # class dimensions
        w.put(0, 2)                                            # subclasses = 0 -> one book slot
        w.put(sub_book + 1, 8)                                 # subclass book (index + 1)
    w.put(multiplier - 1, 2)
    w.put(rangebits, 4)
    for part in xs:
        for x in part:
            w.put(x, rangebits)


def write_floor0(w: BitWriter, order: int, rate: int, bark_map_size: int, amp_bits: int, amp_ofs: int, books):
    """Floor 0 header (Floor0.Init, Floor0.cs:28-51)."""
    w.put(order, 8); w.put(rate, 16); w.put(bark_map_size, 16); w.put(amp_bits, 6); w.put(amp_ofs, 8)
    w.put(len(books) - 1, 4)
    for b in books:
        w.put(b, 8)


def write_residue(w: BitWriter, rtype: int, begin: int, end: int, psize: int, classbook: int, cascades, books):
    w.put(rtype, 16)
    w.put(begin, 24); w.put(end, 24); w.put(psize - 1, 24); w.put(len(cascades) - 1, 6); w.put(classbook, 8)
    for c in cascades:
        w.put(c & 7, 3)
        if c >> 3:
            w.put(1, 1); w.put(c >> 3, 5)
        else:
            w.put(0, 1)
    for c, row in zip(cascades, books):
        for st in range(8):
            if (c >> st) & 1:
                w.put(row[st], 8)


def write_mapping(w: BitWriter, channels: int, coupling, floor: int, residue: int):
    w.put(0, 16)
    w.put(0, 1)                                                # one submap
    if coupling:
        w.put(1, 1); w.put(len(coupling) - 1, 8)
        bits = ilog(channels - 1)
        for m, a in coupling:
            w.put(m, bits); w.put(a, bits)
    else:
        w.put(0, 1)
    w.put(0, 2)
    w.put(0, 8); w.put(floor, 8); w.put(residue, 8)


def build_stream(channels: int, bs0: int, bs1: int, residue_type: int = 2, coupling=(), lookup: int = 1, sequence_p: bool = False,
                 rate: int = 44100, floor_type: int = 1, floor_posts: int = 9, res_dims=(2, 4, 8)):
    """Header packets of a synthetic stream: two floors / residues / mappings / modes (short, long).
    Books: 0 = class book (dims 2, entries 16 -> 4 classes), 1..3 = residue books (dims 2, 4, 8), 4 = floor book (scalar)."""
    w = BitWriter()
    for b in b"\x05vorbis":
        w.put(b, 8)
    w.put((6 if floor_type == 0 else 5) - 1, 8)
    write_codebook(w, 2, 16, 0)                                                    # 0: class book: 4^2 class words
    if lookup == 1:
        write_codebook(w, 2, 25, 1, -1.0, 0.5, 3, sequence_p)                      # 1: 5 values per dim
        write_codebook(w, 4, 81, 1, -0.75, 0.75, 2, sequence_p)                    # 2: 3 values per dim
        write_codebook(w, 8, 256, 1, -0.5, 1.0, 1, sequence_p)                     # 3: 2 values per dim
    else:
        write_codebook(w, res_dims[0], 32, 2, -1.0, 0.125, 4, sequence_p)
        write_codebook(w, res_dims[1], 64, 2, -0.5, 0.0625, 4, sequence_p)
        write_codebook(w, res_dims[2], 128, 2, -0.25, 0.03125, 4, sequence_p)
    write_codebook(w, 1, 128, 0)                                                   # 4: floor values (scalar, 7 bits)
    if floor_type == 0:
        write_codebook(w, 2, 49, 1, 0.1, 0.1, 3, False)                            # 5: LSP steps for floor 0 (7 values per dim)
    w.put(0, 6); w.put(0, 16)                                                      # time domain transforms: 1 dummy
    w.put(2 - 1, 6)                                                                # floors
    for bs in (bs0, bs1):
        n = bs // 2
        if floor_type == 0:                                                        # odd order for the short floor, even for the long one
            w.put(0, 16)
            write_floor0(w, 7 if bs == bs0 else 8, rate, min(n // 2, 256), 6, 100, [5])
            continue
        rb = ilog(n) - 1 if (1 << (ilog(n) - 1)) == n else ilog(n)                 # x_list[1] = 1 << rangebits = n
        w.put(1, 16)
        if floor_posts <= 9:
            inner = sorted(set(int(v) for v in np.unique(np.round(np.geomspace(2, n - 1, 9)))))
            xs = [inner[:3], inner[3:6], inner[6:]]
        else:                                                                      # more than 32 posts: the 64-bit mask paths of the kernels
            cnt = min(floor_posts, n - 3)
            inner = sorted(set(int(v) for v in np.unique(np.round(np.linspace(2, n - 1, cnt)))))
            order = list(np.random.default_rng(n).permutation(len(inner)))          # posts are NOT stored in x order
            inner = [inner[k] for k in order]
            xs = [inner[k:k + 8] for k in range(0, len(inner), 8)]
        write_floor1(w, xs, 2, rb, None, 0, 4)
    w.put(2 - 1, 6)                                                                # residues
    for bs in (bs0, bs1):
        span = (bs // 2) * (channels if residue_type == 2 else 1)
        psize = 16 if bs <= 256 else 32
        end = (span * 3 // 4) // psize * psize
        cascades = [0, 4, 3, 7]
        books = [[0] * 8, [0, 0, 1] + [0] * 5, [2, 3] + [0] * 6, [1, 2, 3] + [0] * 5]
        write_residue(w, residue_type, 0, end, psize, 0, cascades, books)
    w.put(2 - 1, 6)                                                                # mappings
    for i in range(2):
        write_mapping(w, channels, list(coupling), i, i)
    w.put(2 - 1, 6)                                                                # modes
    for i in range(2):
        w.put(i, 1); w.put(0, 16); w.put(0, 16); w.put(i, 8)
    w.put(1, 1)
    packets = [id_header(channels, rate, bs0, bs1), comment_header(), w.done()]
    data = np.frombuffer(b"".join(packets), np.uint8).copy()
    sizes = np.array([len(p) for p in packets], np.int64)
    return data, sizes, np.zeros(3, np.int64), np.zeros(3, np.uint8)


def floor0_min_sqrt_pq(coeff: np.ndarray, bark_map_size: int) -> float:
    """min over the bark bands of sqrt(p + q) of Floor0.Apply (Floor0.cs:171-196), in float64: sizes the test amplitudes."""
    c = 2.0 * np.cos(coeff.astype(np.float64))
    w = 2.0 * np.cos(np.pi / bark_map_size * np.arange(bark_map_size))
    p = np.full_like(w, 0.5); q = np.full_like(w, 0.5)
    j = 1
    while j < len(c):
        q *= w - c[j - 1]; p *= w - c[j]; j += 2
    if j == len(c):
        q *= w - c[j - 1]; p *= p * (4.0 - w * w); q *= q
    else:
        p *= p * (2.0 - w); q *= q * (2.0 + w)
    return float(np.sqrt(p + q).min())


def random_records(rng: np.random.Generator, desc: dict, n_frames: int, post_stride: int, short_prob: float = 0.3,
                   silent_prob: float = 0.1, floor0_stride: int = 0):
    """Seeded synthetic boundary records (nvb_frame + posts + classes + entries) for a setup description: consistent
    window flags, random posts / classes / entries, energy flags per Mapping.cs:105-119."""
    from nvorbis_b200 import capi
    C = desc["channels"]; bs0, bs1 = desc["block_size"]
    is_long = rng.random(n_frames) >= short_prob
    frames = np.zeros(n_frames, capi.FRAME_DTYPE)
    posts = np.zeros((n_frames, C, post_stride), np.int16)
    floor0 = np.zeros((n_frames, C, floor0_stride), np.float32) if floor0_stride else None
    classes, entries = [], []
    coff = eoff = 0
    for i in range(n_frames):
        lng = bool(is_long[i])
        mode = 1 if lng else 0
        m = desc["modes"][mode]; mp = desc["mappings"][m["mapping"]]
        fl = desc["floors"][mp["floor"]]; rs = desc["residues"][mp["residue"]]
        N = bs1 if lng else bs0
        if lng:
            prev = bool(is_long[i - 1]) if i > 0 else True
            nxt = bool(is_long[i + 1]) if i + 1 < n_frames else True
            window = (1 if prev else 0) + (2 if nxt else 0)
            pn, nn = (bs1 if prev else bs0), (bs1 if nxt else bs0)
            start, total = N // 4 - pn // 4, N // 4 * 3 + nn // 4
            valid = total - nn // 4 * 2
        else:
            window, start, valid, total = 0, 0, N // 2, N
        live = 0
        for c in range(C):
            if rng.random() < silent_prob:
                continue
            if fl["type"] == 0:                                # Floor0.Unpack's result: amplitude + LSP angles in (0, pi)
                # well separated angles; the amplitude is the largest step of the amp grid that keeps the curve
                # exp((amp / sqrt(p + q) - ampOfs) * 0.115) below about -10 dB (the 1e-5 absolute gate assumes |pcm| of order 1)
                o = fl["order"]
                for _ in range(64):
                    co = (0.25 + np.arange(o) * (2.6 / o) + rng.uniform(-0.08, 0.08, o)).astype(np.float32)
                    smin = floor0_min_sqrt_pq(co, fl["bark_map_size"])
                    q = min(int(((1 << fl["amp_bits"]) - 1) * smin * 0.9), (1 << fl["amp_bits"]) - 1)
                    if q >= 1:
                        break
                else:
                    continue
                floor0[i, c, 0] = np.float32(q) / np.float32((1 << fl["amp_bits"]) - 1) * np.float32(fl["amp_ofs"])
                floor0[i, c, 1:1 + o] = co
                live |= 1 << c
                continue
            posts[i, c, 0] = fl["n_posts"]
            posts[i, c, 1:3] = rng.integers(20, 60, 2)
            vals = rng.integers(1, 6, fl["n_posts"] - 2)
            vals[rng.random(fl["n_posts"] - 2) < 0.5] = 0
            posts[i, c, 3:1 + fl["n_posts"]] = vals
            live |= 1 << c
        execm = live
        for mg, an in zip(mp["magnitude"], mp["angle"]):
            mg, an = int(mg), int(an)                                       # (Python ints: 32 channels fill a 32-bit mask)
            if ((execm >> mg) | (execm >> an)) & 1:
                execm |= (1 << mg) | (1 << an)
        f = frames[i]
        f["status"], f["mode"], f["window"] = capi.FRAME_OK, mode, window
        f["exec_mask"], f["start"], f["valid"], f["total"] = execm, start, valid, total
        f["classes_off"], f["entries_off"] = coff, eoff
        if live:
            f["res_decoded"] = 1
            span = N * C // 2 if rs["type"] == 2 else N // 2
            streams = 1 if rs["type"] == 2 else C
            P = (min(rs["end"], span) - rs["begin"]) // rs["partition_size"]
            cls = rng.integers(0, rs["classifications"], streams * P).astype(np.uint8)
            classes.append(cls); coff += cls.size
            cnt = 0
            for st in range(rs["max_stages"]):
                for p in range(P):
                    for s_ in range(streams):
                        cl = int(cls[s_ * P + p])
                        if (int(rs["cascade"][cl]) >> st) & 1 and rs["books"][cl][st] >= 0:
                            bk = desc["books"][int(rs["books"][cl][st])]
                            k = rs["partition_size"] // bk["dims"] if rs["type"] == 0 else -(-rs["partition_size"] // bk["dims"])
                            entries.append(rng.integers(0, bk["entries"], k).astype(np.uint16)); cnt += k
            f["entry_count"] = cnt; eoff += cnt
    cls_all = np.concatenate(classes) if classes else np.zeros(0, np.uint8)
    ent_all = np.concatenate(entries) if entries else np.zeros(0, np.uint16)
    return capi.HostBatch(frames, posts.reshape(-1), cls_all, ent_all, None if floor0 is None else floor0.reshape(-1))


# ---- audio packet writer for the floor_type=0 streams of build_stream ------------------------------------------------
def _codeword(w: BitWriter, entry: int, length: int):
    """Codeword of `entry` in a book whose codewords all have `length` bits: the canonical assignment hands out the
    numbers 0, 1, 2, ... (Huffman.cs:15-76) and the decoder reads them most significant bit first."""
    rev = 0
    for i in range(length):
        rev |= ((entry >> i) & 1) << (length - 1 - i)
    w.put(rev, length)


def floor0_audio_packets(rng: np.random.Generator, channels: int, bs0: int, bs1: int, n_frames: int, residue_type: int = 2,
                         short_prob: float = 0.3, silent_prob: float = 0.15, rate: int = 44100):
    """Encodes `n_frames` audio packets for build_stream(..., floor_type=0, lookup=1): mode / window flags, Floor 0 payload
    (amplitude, book number, LSP codewords: Floor0.cs:98-135) and the residue (class words and VQ entries in the order
    Residue0.Decode reads them, Residue0.cs:119-178).  Returns the list of packet byte strings."""
    LEN = {0: 4, 1: 5, 2: 7, 3: 8, 5: 6}                       # codeword length of every entry of book b (write_codebook)
    DIMS = {1: 2, 2: 4, 3: 8}
    cascades = [0, 4, 3, 7]
    books = [[0] * 8, [0, 0, 1] + [0] * 5, [2, 3] + [0] * 6, [1, 2, 3] + [0] * 5]
    is_long = rng.random(n_frames) >= short_prob
    packets = []
    for i in range(n_frames):
        lng = bool(is_long[i])
        N = bs1 if lng else bs0
        n = N // 2
        w = BitWriter()
        w.put(0, 1); w.put(1 if lng else 0, 1)                 # audio packet, mode number (ilog(2 - 1) = 1 bit)
        if lng:
            w.put(1 if (i == 0 or is_long[i - 1]) else 0, 1)
            w.put(1 if (i + 1 >= n_frames or is_long[i + 1]) else 0, 1)
        order, bms = (8, min(n // 2, 256)) if lng else (7, min(n // 2, 256))
        live = False
        for c in range(channels):
            if rng.random() < silent_prob:
                w.put(0, 6); continue
            for _ in range(64):
                # book 5 is lookup type 1 over 7 values 0.1 + 0.1 mult[m], mult[m] = (7 m + 3) % 8: entry = m0 + 7 m1 carries (v[m0], v[m1]); the "averaging"
                # pass adds the previous codeword's last value to both (Floor0.cs:136-147)
                mult = [(k * 7 + 3) % 8 for k in range(7)]          # write_codebook's default multiplicands: 3 2 1 0 7 6 5
                ents = [int(rng.choice([0, 1])) + 7 * int(rng.choice([4, 5, 6])) for _ in range((order + 1) // 2)]
                co, last = [], 0.0
                for e in ents:
                    v0, v1 = 0.1 + 0.1 * mult[e % 7] + last, 0.1 + 0.1 * mult[e // 7] + last
                    co += [v0, v1]; last = v1
                q = min(int(63 * floor0_min_sqrt_pq(np.array(co[:order], np.float32), bms) * 0.9), 63)
                if q >= 1:
                    break
            else:
                w.put(0, 6); continue
            w.put(q, 6); w.put(0, 1)                           # amplitude, book number (ilog(1) = 1 bit)
            for e in ents:
                _codeword(w, e, LEN[5])
            live = True
        if live:
            span = n * channels if residue_type == 2 else n
            streams = 1 if residue_type == 2 else channels
            psize = 16 if N <= 256 else 32
            end = (span * 3 // 4) // psize * psize
            P = end // psize
            cls = rng.integers(0, 4, (streams, P))
            for stage in range(3):
                for word in range((P + 1) // 2):
                    if stage == 0:
                        for s_ in range(streams):
                            c1 = int(cls[s_, 2 * word + 1]) if 2 * word + 1 < P else 0
                            _codeword(w, int(cls[s_, 2 * word]) * 4 + c1, LEN[0])
                    for d in range(2):
                        p = 2 * word + d
                        if p >= P:
                            break
                        for s_ in range(streams):
                            cl = int(cls[s_, p])
                            if (cascades[cl] >> stage) & 1:
                                bk = books[cl][stage]
                                cnt = psize // DIMS[bk]
                                hi = {1: 25, 2: 81, 3: 256}[bk]
                                for _ in range(cnt):
                                    _codeword(w, int(rng.integers(0, hi)), LEN[bk])
        packets.append(w.done())
    return packets
