// nvh_host.cpp -- libnvorbis_host.so: the host half of the split decoder (include/nvorbis_host.h).
//
// Ogg demux, header parsing and the bit-unpacking half of every audio packet, producing the plain arrays of
// nvb_setup / nvb_batch.  CPU only, multi-threaded over packets.  Written from the behaviour of the reference
// (citations are NVorbis/ file:line); it shares no code with the test oracle and is never linked with it.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "../../../include/nvorbis_host.h"
#include "../nvb_unpack_tables.h"

namespace nvh {

struct DataError : std::runtime_error { using std::runtime_error::runtime_error; };

static int ilog(int x) { int n = 0; while (x > 0) { ++n; x >>= 1; } return n; }             // Utils.cs:5-14
static uint32_t bitrev32(uint32_t n) {
    n = ((n & 0xAAAAAAAAu) >> 1) | ((n & 0x55555555u) << 1);
    n = ((n & 0xCCCCCCCCu) >> 2) | ((n & 0x33333333u) << 2);
    n = ((n & 0xF0F0F0F0u) >> 4) | ((n & 0x0F0F0F0Fu) << 4);
    n = ((n & 0xFF00FF00u) >> 8) | ((n & 0x00FF00FFu) << 8);
    return (n >> 16) | (n << 16);
}

// ---------------------------------------------------------------------------------------------------
// Bit cursor over one packet, LSB first (DataPacket.cs:150-283).  Reading past the end yields zero bits and
// raises `short_`, as the reference's TryPeekBits/SkipBits pair does (IsShort, DataPacket.cs:255-279).
// ---------------------------------------------------------------------------------------------------
struct Bits {
    const uint8_t* p; size_t nbits; size_t pos = 0; bool short_ = false;
    Bits(const uint8_t* d, size_t bytes) : p(d), nbits(bytes * 8) {}
    size_t left() const { return pos < nbits ? nbits - pos : 0; }
    // up to 32 bits at the cursor, zero-padded beyond the end; does not advance
    uint32_t peek32() const {
        const size_t byte = pos >> 3; const unsigned sh = pos & 7;
        const size_t nbytes = nbits >> 3;
        if (byte + 8 <= nbytes) {                                             // common case: one unaligned 64-bit window, no masking (>= 57 bits left)
            uint64_t w; std::memcpy(&w, p + byte, 8);
            return (uint32_t)(w >> sh);
        }
        uint64_t v = 0;
        for (size_t i = 0; i < 5 && byte + i < nbytes; i++) v |= (uint64_t)p[byte + i] << (8 * i);
        v >>= sh;
        const size_t l = left();
        if (l < 32) v &= (l == 0) ? 0 : ((1ull << l) - 1);
        return (uint32_t)v;
    }
    void skip(int n) { if ((size_t)n > left()) { pos = nbits; short_ = true; } else pos += (size_t)n; }
    uint32_t read(int n) {                                                    // n <= 32
        if (n == 0) return 0;
        uint32_t v = peek32();
        if (n < 32) v &= (1u << n) - 1;
        skip(n);
        return v;
    }
    bool bit() { return read(1) != 0; }
};

// ---------------------------------------------------------------------------------------------------
// Codebooks (Codebook.cs:59-322, Huffman.cs:15-76).  Decoding uses a 10-bit root table plus per-root
// chains for longer codewords; it returns what Codebook.DecodeScalar returns for the same bits.
// ---------------------------------------------------------------------------------------------------
struct Book {
    int dims = 0, entries = 0, map_type = 0;
    std::vector<int8_t> len;                       // 0 = unused entry
    std::vector<float> table;
    static constexpr int ROOT_BITS = 10;
    struct Root { int32_t value; uint8_t len; };   // len 0: no short code here; value = head of the chain (-1 none)
    struct Long { uint32_t code; int32_t value; int32_t next; uint8_t len; };
    std::vector<Root> root; std::vector<Long> longs;
    int root_bits = 0; bool decodable = false;

    void parse(Bits& b) {
        if (b.read(24) != 0x564342u) throw DataError("Book header had invalid signature!");          // Codebook.cs:62-63
        dims = (int)b.read(16); entries = (int)b.read(24);
        if (b.short_) throw DataError("codebook header is truncated");
        len.assign((size_t)entries, 0);
        if (b.bit()) {                                                                                 // ordered lengths, Codebook.cs:83-104
            int l = (int)b.read(5) + 1;
            for (int i = 0; i < entries;) {
                // a codeword is at most 32 bits (the length field of the unordered form is 5 bits + 1); past the end of the
                // packet every count reads as 0 and the reference's loop would never end
                if (l > 32) throw DataError("ordered codebook: codeword length above 32");
                int cnt = (int)b.read(ilog(entries - i));
                if (b.short_) throw DataError("ordered codebook: packet is truncated");
                if (i + cnt > entries) throw DataError("ordered codebook overrun");
                for (int k = 0; k < cnt; k++) len[(size_t)i++] = (int8_t)l;
                ++l;
            }
        } else {
            const bool sparse = b.bit();                                                               // Codebook.cs:106-127
            for (int i = 0; i < entries; i++)
                if (!sparse || b.bit()) len[(size_t)i] = (int8_t)(b.read(5) + 1);
        }
        build_decoder();
        parse_lookup(b);
    }

    // Codeword assignment of the Vorbis I spec (section 3.2.1), the tree the reference builds in
    // Codebook.ComputeCodewords (Codebook.cs:172-207): each entry takes the lowest free codeword of its length.
    void build_decoder() {
        uint32_t next_free[33] = {0};               // next_free[l]: lowest unused l-bit prefix (MSB first), if any
        std::vector<uint32_t> code((size_t)entries, 0);
        int maxlen = 0; bool any = false;
        for (int i = 0; i < entries; i++) {
            const int l = len[(size_t)i];
            if (l <= 0) continue;
            if (l > 32) throw DataError("codeword length above 32");
            any = true; if (l > maxlen) maxlen = l;
            uint32_t c = next_free[l];
            if (l < 32 && (c >> l)) throw DataError("codebook is over-specified");
            code[(size_t)i] = c;
            for (int j = l; j > 0; j--) {           // claim the subtree: bump the markers above ...
                if (next_free[j] & 1) { if (j == 1) next_free[1]++; else next_free[j] = next_free[j - 1] << 1; break; }
                next_free[j]++;
            }
            for (int j = l + 1; j < 33; j++) {      // ... and re-point the longer lengths that hung below this code
                if ((next_free[j] >> 1) == c) { c = next_free[j]; next_free[j] = next_free[j - 1] << 1; } else break;
            }
        }
        decodable = any;
        if (!any) return;
        root_bits = maxlen < ROOT_BITS ? maxlen : ROOT_BITS;
        root.assign((size_t)1 << root_bits, Root{-1, 0});
        for (int i = 0; i < entries; i++) {
            const int l = len[(size_t)i];
            if (l <= 0) continue;
            const uint32_t rev = bitrev32(code[(size_t)i]) >> (32 - l);      // first transmitted bit in bit 0
            if (l <= root_bits) {
                for (uint32_t hi = 0; hi < (1u << (root_bits - l)); hi++) root[(hi << l) | rev] = Root{i, (uint8_t)l};
            } else {
                Root& r = root[rev & ((1u << root_bits) - 1)];
                longs.push_back(Long{rev, i, r.value, (uint8_t)l});
                r.value = (int32_t)longs.size() - 1;
            }
        }
    }

    // Codebook.DecodeScalar (Codebook.cs:294-320): -1 when no bit is left or no codeword matches.
    int decode(Bits& b) const {
        if (!decodable || b.left() == 0) return -1;
        const uint32_t v = b.peek32();
        const Root& r = root[v & ((1u << root_bits) - 1)];
        if (r.len) { b.skip(r.len); return r.value; }
        for (int k = r.value; k >= 0; k = longs[(size_t)k].next) {
            const Long& L = longs[(size_t)k];
            const uint32_t mask = L.len >= 32 ? 0xFFFFFFFFu : ((1u << L.len) - 1);
            if ((v & mask) == L.code) { b.skip(L.len); return L.value; }
        }
        return -1;
    }

    static float unpack_float(uint32_t bits) {                                                        // Utils.cs:45-59
        const int32_t sign = (int32_t)bits >> 31;
        const double e = (double)((int)((bits & 0x7fe00000u) >> 21) - 788);
        const float mant = (float)(int32_t)(((bits & 0x1fffffu) ^ (uint32_t)sign) + (uint32_t)(sign & 1));
        return mant * (float)std::pow(2.0, e);
    }

    // VQ lookup table (Codebook.InitLookupTable, Codebook.cs:222-283): float product + float min, then a
    // double running sum when sequence_p is set, stored as float.
    void parse_lookup(Bits& b) {
        map_type = (int)b.read(4);
        if (map_type == 0) return;
        if (map_type > 2) throw DataError("invalid codebook lookup type");
        // a value table needs at least one dimension (Floor0.cs:48 rejects such books; for residues the reference divides by
        // Dimensions, Residue0.cs:183): refuse the setup instead of dividing by zero per packet
        if (dims < 1) throw DataError("codebook with a lookup table has no dimensions");
        const float vmin = unpack_float(b.read(32)), vdelta = unpack_float(b.read(32));
        const int vbits = (int)b.read(4) + 1;
        const bool seq = b.bit();
        size_t nq = (size_t)entries * (size_t)dims;
        table.assign(nq, 0.f);
        if (map_type == 1) {                                                                          // Codebook.cs:285-292
            int r = (int)std::floor(std::exp(std::log((double)entries) / dims));
            if (std::floor(std::pow((double)(r + 1), (double)dims)) <= entries) ++r;
            nq = (size_t)r;
        }
        std::vector<uint32_t> q(nq);
        for (auto& x : q) x = b.read(vbits);
        for (int e = 0; e < entries; e++) {
            double run = 0.0; size_t div = 1;
            for (int d = 0; d < dims; d++) {
                const size_t qi = map_type == 1 ? ((size_t)e / div) % nq : (size_t)e * dims + d;
                float f = (float)q[qi] * vdelta;
                f = f + vmin;
                const double v = (double)f + run;
                table[(size_t)e * dims + d] = (float)v;
                if (seq) run = v;
                if (map_type == 1) div *= nq;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// Setup tables
// ---------------------------------------------------------------------------------------------------
struct Floor1Def {
    std::vector<int> part_class, class_dims, class_subs, class_master;
    std::vector<std::vector<int>> sub_books;
    int mult = 0, range = 0, ybits = 0;
    std::vector<int> x, lo, hi, order;
};
struct Floor0Def {                   // Floor0.Init, Floor0.cs:28-51
    int order = 0, rate = 0, bark_map_size = 0, amp_bits = 0, amp_ofs = 0, amp_div = 0, book_bits = 0;
    std::vector<int> books;
};
struct ResidueDef {
    int type = 0, begin = 0, end = 0, psize = 0, nclass = 0, class_book = 0, stages = 0;
    int streams = 1;                 // Residue0._channels (1 for type 2, Residue2.cs:13)
    int cascade[64]; int books[64][8];
    std::vector<std::vector<uint8_t>> class_digits;     // _decodeMap (Residue0.cs:100-114)
};
struct MappingDef { std::vector<int> mag, ang; int submaps = 1; int floor0 = 0, residue0 = 0; std::vector<int> ch_floor, ch_residue; };
struct ModeDef { bool long_block = false; int mapping = 0; };

struct PacketRef { size_t off; size_t size; int64_t granule; uint8_t flags; };       // flags: 1 granule, 2 EOS, 4 resync

struct UnpackedFrame {                 // one packet's result before concatenation
    nvb_frame f; int n_classes = 0;
    int64_t granule = 0; bool has_granule = false, eos = false, resync = false;
};

// Ogg container: pages -> packets of one logical stream, incrementally (the whole image at once is the one-call case).
// Ogg/PageReaderBase.cs:33-111,227-292 (sync + CRC), Ogg/PageReader.cs:27-93,125-158 (lacing; zero-length packets
// and packet-less pages are dropped), Ogg/StreamPageReader.cs:44-90 (EOS page, sequence gaps = resync),
// Ogg/PacketProvider.cs:324-438 (continuations; granule on the last packet completed in a page; EOS flag).
// scan() turns the bytes seen so far into page records, emit() turns complete pages into packets; with final == false both
// stop where more input could still change the answer (a page or a continued packet that is not complete yet) -- the
// forward-only reader (Ogg/ForwardOnlyPageReader.cs, ForwardOnlyPacketProvider.cs) in feed form.
struct OggDemux {
    struct Piece { size_t off, size; };
    struct PageInfo { int64_t granule; bool resync, continued, continuation, eos; std::vector<Piece> pieces; };
    uint32_t crc_table[256];
    std::vector<PageInfo> pages;
    size_t i = 0; bool lost_sync = false, have_serial = false, done = false; uint32_t serial = 0; int32_t last_seq = 0;
    size_t pi = 0, k = 0;                                                     // packet cursor: (page, piece)
    OggDemux() {
        for (uint32_t n = 0; n < 256; n++) {                                                          // Ogg/Crc.cs:8-21
            uint32_t r = n << 24;
            for (int b = 0; b < 8; b++) r = (r << 1) ^ ((r & 0x80000000u) ? 0x04c11db7u : 0u);
            crc_table[n] = r;
        }
    }
    void want(uint32_t sn) { have_serial = true; serial = sn; }

    void scan(const uint8_t* d, size_t len, bool final, nvh_stream& s);
    void emit(const uint8_t* d, bool final, nvh_stream& s);
};

}  // namespace nvh

using namespace nvh;

struct nvh_stream {
    std::string err;
    std::vector<uint8_t> bytes;                  // packet payloads back to back
    std::vector<PacketRef> packets;
    bool has_eos = false;
    // headers
    int channels = 0, sample_rate = 0, bs[2] = {0, 0}, mode_bits = 0;
    std::vector<Book> books; std::vector<Floor1Def> floors; std::vector<int> floor_type; std::vector<Floor0Def> floors0;
    int f0_stride = 0;                           // floats per (frame, channel) in nvb_batch.floor0, 0 = no type 0 floor
    std::vector<ResidueDef> residues; std::vector<MappingDef> mappings; std::vector<ModeDef> modes;
    // C-ABI view of the setup
    std::vector<nvb_codebook> c_books; std::vector<float> c_vq; std::vector<nvb_floor> c_floors; std::vector<nvb_residue> c_residues;
    std::vector<nvb_mapping> c_mappings; std::vector<nvb_mode> c_modes; nvb_setup c_setup;
    int post_stride = 0;
    // decode cursor (StreamDecoder state that lives on the host)
    size_t first_audio = 3, next_packet = 3;
    bool have_prev = false; int prev_start = 0, prev_end = 0, prev_stop = 0;
    int64_t position = 0; bool has_position = false; bool eos_found = false;
    // last unpacked batch
    std::vector<nvb_frame> o_frames; std::vector<int16_t> o_posts; std::vector<uint8_t> o_classes; std::vector<uint16_t> o_entries;
    std::vector<float> o_floor0;
    // GPU-side unpack: the tables blob (built on demand) and the last packet batch
    std::vector<uint8_t> u_blob; std::vector<uint32_t> o_offsets; std::vector<uint8_t> o_pad;
    // forward-only (feed) mode: the container bytes seen so far, the incremental demuxer, whether the input has ended
    bool forward = false, input_final = true, headers_done = false;
    std::vector<uint8_t> raw; OggDemux demux;
};

namespace nvh {

static thread_local std::string g_open_error;

// ---- Ogg container: pages -> packets of the first logical stream ----------------------------------------
// Ogg/PageReaderBase.cs:33-111,227-292 (sync + CRC), Ogg/PageReader.cs:27-93,125-158 (lacing; zero-length packets
// and packet-less pages are dropped), Ogg/StreamPageReader.cs:44-90 (EOS page, sequence gaps = resync),
// Ogg/PacketProvider.cs:324-438 (continuations; granule on the last packet completed in a page; EOS flag).
// Serial numbers of the logical streams of a container, in the order their first pages appear (multiplexed or chained
// streams: ContainerReader raises NewStreamCallback per serial, Ogg/ContainerReader.cs).  Pages with a bad CRC are skipped as
// the page reader does.
static std::vector<uint32_t> ogg_serials(const uint8_t* d, size_t len) {
    uint32_t crc_table[256];
    for (uint32_t i = 0; i < 256; i++) { uint32_t r = i << 24; for (int k = 0; k < 8; k++) r = (r << 1) ^ ((r & 0x80000000u) ? 0x04c11db7u : 0u); crc_table[i] = r; }
    std::vector<uint32_t> out;
    size_t i = 0;
    while (i + 27 <= len) {
        if (std::memcmp(d + i, "OggS", 4) != 0 || d[i + 4] != 0) { ++i; continue; }
        const int nseg = d[i + 26];
        if (i + 27 + (size_t)nseg > len) { ++i; continue; }
        size_t body = 0; for (int k = 0; k < nseg; k++) body += d[i + 27 + k];
        const size_t total = 27 + (size_t)nseg + body;
        if (i + total > len) { ++i; continue; }
        uint32_t crc = 0;
        for (size_t k = 0; k < total; k++) { const uint8_t byte = (k >= 22 && k < 26) ? 0 : d[i + k]; crc = (crc << 8) ^ crc_table[byte ^ (crc >> 24)]; }
        uint32_t stored; std::memcpy(&stored, d + i + 22, 4);
        if (crc != stored) { ++i; continue; }
        uint32_t serial; std::memcpy(&serial, d + i + 14, 4);
        if (std::find(out.begin(), out.end(), serial) == out.end()) out.push_back(serial);
        i += total;
    }
    return out;
}

void OggDemux::scan(const uint8_t* d, size_t len, bool final, nvh_stream& s) {
    while (!done) {
        if (i + 27 > len) { if (final) i = len; break; }
        if (std::memcmp(d + i, "OggS", 4) != 0 || d[i + 4] != 0) { ++i; lost_sync = true; continue; }
        const int nseg = d[i + 26];
        if (i + 27 + (size_t)nseg > len) { if (!final) break; ++i; lost_sync = true; continue; }
        size_t body = 0; for (int n = 0; n < nseg; n++) body += d[i + 27 + n];
        const size_t total = 27 + (size_t)nseg + body;
        if (i + total > len) { if (!final) break; ++i; lost_sync = true; continue; }
        uint32_t crc = 0;
        for (size_t n = 0; n < total; n++) {
            const uint8_t byte = (n >= 22 && n < 26) ? 0 : d[i + n];
            crc = (crc << 8) ^ crc_table[byte ^ (crc >> 24)];
        }
        uint32_t stored; std::memcpy(&stored, d + i + 22, 4);
        if (crc != stored) { ++i; lost_sync = true; continue; }
        uint32_t pg_serial; std::memcpy(&pg_serial, d + i + 14, 4);
        int32_t seq; std::memcpy(&seq, d + i + 18, 4);
        int64_t granule; std::memcpy(&granule, d + i + 6, 8);
        const uint8_t flags = d[i + 5];
        PageInfo pg; pg.granule = granule; pg.continuation = flags & 1; pg.eos = (flags & 4) != 0; pg.continued = false;
        size_t off = i + 27 + (size_t)nseg, run = 0;
        for (int n = 0; n < nseg; n++) {
            const int seg = d[i + 27 + n]; run += (size_t)seg;
            if (seg < 255) { if (run) { pg.pieces.push_back(Piece{off, run}); off += run; } run = 0; }
        }
        if (run) { pg.pieces.push_back(Piece{off, run}); pg.continued = d[i + 26 + nseg] == 255; }
        i += total;
        const bool was_lost = lost_sync; lost_sync = false;
        if (!have_serial) { have_serial = true; serial = pg_serial; }
        if (pg_serial != serial) continue;                    // other logical streams are not decoded here
        if (pg.pieces.empty()) continue;                      // Ogg/PageReader.cs:131
        pg.resync = was_lost || (last_seq != 0 && last_seq + 1 != seq);
        last_seq = seq;
        if (pg.eos) { s.has_eos = true; done = true; }
        pages.push_back(std::move(pg));
    }
}

// packets: a (page, piece) cursor advanced the way PacketProvider.CreatePacket does (:411-434)
void OggDemux::emit(const uint8_t* d, bool final, nvh_stream& s) {
    const bool complete = final || done;                      // no further page can arrive
    while (pi < pages.size()) {
        const PageInfo& pg = pages[pi];
        PacketRef pr; pr.off = s.bytes.size(); pr.flags = 0; pr.granule = 0;
        s.bytes.insert(s.bytes.end(), d + pg.pieces[k].off, d + pg.pieces[k].off + pg.pieces[k].size);
        const bool last_piece = k + 1 == pg.pieces.size();
        bool resync = pg.resync, last_in_page = last_piece; int64_t granule = pg.granule;
        size_t final_page = pi, final_pieces = pg.pieces.size();
        if (last_piece && pg.continued) {
            bool cont = true; size_t cp = pi;
            while (cont) {
                if (++cp >= pages.size()) {                                                       // the rest has not arrived (or never will: no packet, :346-350)
                    s.bytes.resize(pr.off);
                    if (complete) pi = pages.size();
                    return;
                }
                const PageInfo& nx = pages[cp];
                granule = nx.granule; resync = nx.resync; cont = nx.continued; final_pieces = nx.pieces.size();
                if (!nx.continuation || nx.resync) break;                                         // broken chain: keep what we have (:354-357)
                if (cont && nx.pieces.size() > 1) cont = false;                                   // (:360-363)
                s.bytes.insert(s.bytes.end(), d + nx.pieces[0].off, d + nx.pieces[0].off + nx.pieces[0].size);
            }
            last_in_page = final_pieces == 1;
            final_page = cp;
        }
        pr.size = s.bytes.size() - pr.off;
        if (resync) pr.flags |= 4;
        if (last_in_page) {                                                                       // (:399-407)
            pr.flags |= 1; pr.granule = granule;
            if (s.has_eos && final_page + 1 == pages.size()) pr.flags |= 2;
        }
        s.packets.push_back(pr);
        if (final_page != pi) { pi = final_page; k = 0; }
        if (k + 1 == final_pieces) { ++pi; k = 0; } else ++k;
    }
}

// The whole container image at once.  want_serial: decode the logical stream with this serial number (has_want) or the first one met.
static void demux_ogg(const uint8_t* d, size_t len, nvh_stream& s, bool has_want = false, uint32_t want_serial = 0) {
    OggDemux dm;
    if (has_want) dm.want(want_serial);
    dm.scan(d, len, true, s);
    dm.emit(d, true, s);
}

static void parse_id_header(Bits& b, nvh_stream& s) {                                                 // StreamDecoder.cs:179-204
    static const uint8_t sig[7] = {0x01, 'v', 'o', 'r', 'b', 'i', 's'};
    for (int i = 0; i < 7; i++) if (b.read(8) != sig[i]) throw DataError("Could not find Vorbis data to decode.");
    if (b.read(32) != 0) throw DataError("Could not find Vorbis data to decode.");                  // version
    s.channels = (int)b.read(8);
    s.sample_rate = (int)b.read(32);
    b.read(32); b.read(32); b.read(32);
    s.bs[0] = 1 << b.read(4); s.bs[1] = 1 << b.read(4);
    if (s.channels < 1) throw DataError("stream has no channels");
}

static int ilog_i(int x) { int c = 0; while (x > 0) { ++c; x >>= 1; } return c; }                     // Utils.cs:5-14

static void parse_floor0(Bits& b, const nvh_stream& s, Floor0Def& f) {                                // Floor0.cs:28-51
    f.order = (int)b.read(8); f.rate = (int)b.read(16); f.bark_map_size = (int)b.read(16);
    f.amp_bits = (int)b.read(6); f.amp_ofs = (int)b.read(8);
    f.books.resize((size_t)b.read(4) + 1);
    if (f.order < 1 || f.rate < 1 || f.bark_map_size < 1) throw DataError("floor0: invalid header");
    f.amp_div = (int)((1ll << f.amp_bits) - 1);
    for (auto& bk : f.books) {
        bk = (int)b.read(8);
        if (bk >= (int)s.books.size()) throw DataError("floor0: book out of range");
        if (s.books[(size_t)bk].map_type == 0 || s.books[(size_t)bk].dims < 1) throw DataError("floor0: book without a lookup table");
    }
    f.book_bits = ilog_i((int)f.books.size());
}

static void parse_floor1(Bits& b, const nvh_stream& s, Floor1Def& f) {                                // Floor1.cs:30-133
    const int nparts = (int)b.read(5);
    int max_class = -1;
    f.part_class.resize((size_t)nparts);
    for (auto& c : f.part_class) { c = (int)b.read(4); max_class = std::max(max_class, c); }
    const int nclass = max_class + 1;
    f.class_dims.assign((size_t)nclass, 0); f.class_subs.assign((size_t)nclass, 0); f.class_master.assign((size_t)nclass, 0);
    f.sub_books.assign((size_t)nclass, {});
    for (int c = 0; c < nclass; c++) {
        f.class_dims[(size_t)c] = (int)b.read(3) + 1;
        f.class_subs[(size_t)c] = (int)b.read(2);
        if (f.class_subs[(size_t)c] > 0) {
            f.class_master[(size_t)c] = (int)b.read(8);
            if (f.class_master[(size_t)c] >= (int)s.books.size()) throw DataError("floor1: master book out of range");
        }
        f.sub_books[(size_t)c].resize((size_t)1 << f.class_subs[(size_t)c]);
        for (auto& bk : f.sub_books[(size_t)c]) { bk = (int)b.read(8) - 1; if (bk >= (int)s.books.size()) throw DataError("floor1: subclass book out of range"); }
    }
    static const int ranges[4] = {256, 128, 86, 64}, ybits[4] = {8, 7, 7, 6};                         // Floor1.cs:27-28
    const int m = (int)b.read(2);
    f.mult = m + 1; f.range = ranges[m]; f.ybits = ybits[m];
    const int rbits = (int)b.read(4);
    f.x = {0, 1 << rbits};
    for (int p = 0; p < nparts; p++)
        for (int k = 0; k < f.class_dims[(size_t)f.part_class[(size_t)p]]; k++) f.x.push_back((int)b.read(rbits));
    const int n = (int)f.x.size();
    if (n > 64) throw DataError("floor1: more than 64 posts");                                        // Floor1.Data.Posts is int[64] (Floor1.cs:12)
    f.lo.assign((size_t)n, 0); f.hi.assign((size_t)n, 0); f.order.resize((size_t)n);
    for (int i = 0; i < n; i++) f.order[(size_t)i] = i;
    for (int i = 2; i < n; i++) {                                                                     // nearest neighbours among earlier posts, Floor1.cs:98-115
        int lo = 0, hi = 1;
        for (int j = 2; j < i; j++) {
            if (f.x[(size_t)j] < f.x[(size_t)i]) { if (f.x[(size_t)j] > f.x[(size_t)lo]) lo = j; }
            else if (f.x[(size_t)j] < f.x[(size_t)hi]) hi = j;
        }
        f.lo[(size_t)i] = lo; f.hi[(size_t)i] = hi;
    }
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (f.x[(size_t)i] == f.x[(size_t)j]) throw DataError("floor1: duplicate x");
    std::sort(f.order.begin(), f.order.end(), [&](int a, int c) { return f.x[(size_t)a] < f.x[(size_t)c]; });   // Floor1.cs:118-132 (x values are distinct)
}

static void parse_residue(Bits& b, const nvh_stream& s, int type, ResidueDef& r) {                    // Residue0.cs:35-117, Residue2.cs:10-14
    r.type = type; r.streams = type == 2 ? 1 : s.channels;
    r.begin = (int)b.read(24); r.end = (int)b.read(24); r.psize = (int)b.read(24) + 1;
    r.nclass = (int)b.read(6) + 1; r.class_book = (int)b.read(8);
    if (r.class_book >= (int)s.books.size()) throw DataError("residue: class book out of range");
    int nbooks = 0;
    for (int c = 0; c < 64; c++) { r.cascade[c] = 0; for (int k = 0; k < 8; k++) r.books[c][k] = -1; }
    for (int c = 0; c < r.nclass; c++) {
        const int low = (int)b.read(3);
        r.cascade[c] = b.bit() ? (((int)b.read(5) << 3) | low) : low;
        nbooks += __builtin_popcount((unsigned)r.cascade[c]);
    }
    std::vector<int> list((size_t)nbooks);
    for (auto& bk : list) {
        bk = (int)b.read(8);
        if (bk >= (int)s.books.size()) throw DataError("residue: book out of range");
        if (s.books[(size_t)bk].map_type == 0) throw DataError("residue: book without a lookup table");     // Residue0.cs:66-67
        if (s.books[(size_t)bk].dims < 1) throw DataError("residue: book without dimensions");
    }
    const Book& cb = s.books[(size_t)r.class_book];
    long long partvals = 1;
    for (int k = 0; k < cb.dims; k++) { partvals *= r.nclass; if (partvals > cb.entries) throw DataError("residue: class book too small"); }
    int at = 0; r.stages = 0;
    for (int c = 0; c < r.nclass; c++) {
        const int st = ilog(r.cascade[c]);
        r.stages = std::max(r.stages, st);
        for (int k = 0; k < st; k++) if ((r.cascade[c] >> k) & 1) r.books[c][k] = list[(size_t)at++];
    }
    r.class_digits.assign((size_t)partvals, std::vector<uint8_t>((size_t)cb.dims));
    for (long long v = 0; v < partvals; v++) {                                                        // most significant digit first, Residue0.cs:100-114
        long long rem = v, mult = partvals / r.nclass;
        for (int k = 0; k < cb.dims; k++) { const long long dgt = rem / mult; rem -= dgt * mult; mult /= r.nclass; r.class_digits[(size_t)v][(size_t)k] = (uint8_t)dgt; }
    }
}

static void parse_mapping(Bits& b, const nvh_stream& s, MappingDef& m) {                              // Mapping.cs:16-93
    m.submaps = b.bit() ? 1 + (int)b.read(4) : 1;
    const int steps = b.bit() ? 1 + (int)b.read(8) : 0;
    const int cbits = ilog(s.channels - 1);
    for (int k = 0; k < steps; k++) {
        const int mag = (int)b.read(cbits), ang = (int)b.read(cbits);
        if (mag == ang || mag > s.channels - 1 || ang > s.channels - 1) throw DataError("Invalid magnitude or angle in mapping header!");
        m.mag.push_back(mag); m.ang.push_back(ang);
    }
    if (b.read(2) != 0) throw DataError("Reserved bits not 0 in mapping header.");
    std::vector<int> mux((size_t)s.channels, 0);
    if (m.submaps > 1) for (auto& x : mux) { x = (int)b.read(4); if (x >= m.submaps) throw DataError("Invalid channel mux submap index in mapping header!"); }
    std::vector<int> sf((size_t)m.submaps), sr((size_t)m.submaps);
    for (int k = 0; k < m.submaps; k++) {
        b.read(8);
        sf[(size_t)k] = (int)b.read(8); if (sf[(size_t)k] >= (int)s.floors.size()) throw DataError("Invalid floor number in mapping header!");
        sr[(size_t)k] = (int)b.read(8); if (sr[(size_t)k] >= (int)s.residues.size()) throw DataError("Invalid residue number in mapping header!");
    }
    m.floor0 = sf[0]; m.residue0 = sr[0];
    for (int c = 0; c < s.channels; c++) { m.ch_floor.push_back(sf[(size_t)mux[(size_t)c]]); m.ch_residue.push_back(sr[(size_t)mux[(size_t)c]]); }
}

static void parse_setup_header(Bits& b, nvh_stream& s) {                                              // StreamDecoder.cs:226-289
    static const uint8_t sig[7] = {0x05, 'v', 'o', 'r', 'b', 'i', 's'};
    for (int i = 0; i < 7; i++) if (b.read(8) != sig[i]) throw DataError("setup header signature");
    s.books.resize((size_t)b.read(8) + 1);
    for (auto& bk : s.books) bk.parse(b);
    const int times = (int)b.read(6) + 1;
    b.skip(16 * times);
    const int nfloors = (int)b.read(6) + 1;
    s.floors.resize((size_t)nfloors); s.floor_type.resize((size_t)nfloors); s.floors0.resize((size_t)nfloors);
    for (int i = 0; i < nfloors; i++) {
        const int type = (int)b.read(16);                                                             // Factory.cs:22-31
        s.floor_type[(size_t)i] = type;
        if (type == 1) parse_floor1(b, s, s.floors[(size_t)i]);
        else if (type == 0) parse_floor0(b, s, s.floors0[(size_t)i]);
        else throw DataError("Invalid floor type!");
    }
    s.residues.resize((size_t)b.read(6) + 1);
    for (auto& r : s.residues) {
        const int type = (int)b.read(16);                                                             // Factory.cs:48-58
        if (type > 2) throw DataError("Invalid residue type!");
        parse_residue(b, s, type, r);
    }
    s.mappings.resize((size_t)b.read(6) + 1);
    for (auto& m : s.mappings) { if (b.read(16) != 0) throw DataError("Invalid mapping type!"); parse_mapping(b, s, m); }
    s.modes.resize((size_t)b.read(6) + 1);
    for (auto& m : s.modes) {                                                                         // Mode.cs:24-41
        m.long_block = b.bit();
        if (b.read(32) != 0) throw DataError("Mode header had invalid window or transform type!");
        m.mapping = (int)b.read(8);
        if (m.mapping >= (int)s.mappings.size()) throw DataError("Mode header had invalid mapping index!");
    }
    if (!b.bit()) throw DataError("Book packet did not end on correct bit!");
    if (b.short_) throw DataError("setup header is truncated");
    s.mode_bits = ilog((int)s.modes.size() - 1);
}

static void export_setup(nvh_stream& s) {
    int64_t off = 0;
    for (const Book& b : s.books) {
        nvb_codebook c; c.dims = b.dims; c.entries = b.entries; c.map_type = b.map_type; c.reserved = 0;
        c.table_off = b.table.empty() ? -1 : off;
        s.c_vq.insert(s.c_vq.end(), b.table.begin(), b.table.end()); off += (int64_t)b.table.size();
        s.c_books.push_back(c);
    }
    int max_posts = 2, max_order0 = 0;
    for (size_t i = 0; i < s.floors.size(); i++) {
        const Floor1Def& f = s.floors[i];
        nvb_floor c; std::memset(&c, 0, sizeof c);
        c.type = s.floor_type[i]; c.f1.n_posts = (int)f.x.size(); c.f1.multiplier = f.mult; c.f1.range = f.range;
        if (c.type == 0) {
            const Floor0Def& z = s.floors0[i];
            c.f0.order = z.order; c.f0.rate = z.rate; c.f0.bark_map_size = z.bark_map_size; c.f0.amp_bits = z.amp_bits; c.f0.amp_ofs = z.amp_ofs;
            max_order0 = std::max(max_order0, z.order);
        }
        for (size_t k = 0; k < f.x.size(); k++) { c.f1.x_list[k] = (uint16_t)f.x[k]; c.f1.l_neigh[k] = (uint8_t)f.lo[k]; c.f1.h_neigh[k] = (uint8_t)f.hi[k]; c.f1.sort_idx[k] = (uint8_t)f.order[k]; }
        max_posts = std::max(max_posts, (int)f.x.size());
        s.c_floors.push_back(c);
    }
    s.post_stride = (2 + max_posts + 1) & ~1;
    s.f0_stride = max_order0 > 0 ? (1 + max_order0 + 1) & ~1 : 0;                                       // same rule as nvb_floor0_stride()
    for (const ResidueDef& r : s.residues) {
        nvb_residue c; std::memset(&c, 0, sizeof c);
        c.type = r.type; c.begin = r.begin; c.end = r.end; c.partition_size = r.psize; c.classifications = r.nclass; c.max_stages = r.stages;
        for (int k = 0; k < 64; k++) { c.cascade[k] = r.cascade[k]; for (int st = 0; st < 8; st++) c.books[k][st] = (int16_t)r.books[k][st]; }
        s.c_residues.push_back(c);
    }
    for (const MappingDef& m : s.mappings) {
        nvb_mapping c; std::memset(&c, 0, sizeof c);
        c.n_coupling = (int)m.mag.size(); c.n_submaps = m.submaps; c.floor = m.floor0; c.residue = m.residue0;
        for (size_t k = 0; k < m.mag.size() && k < NVB_MAX_COUPLING; k++) { c.magnitude[k] = (uint8_t)m.mag[k]; c.angle[k] = (uint8_t)m.ang[k]; }
        s.c_mappings.push_back(c);
    }
    for (const ModeDef& m : s.modes) { nvb_mode c; c.block_flag = m.long_block ? 1 : 0; c.mapping = m.mapping; s.c_modes.push_back(c); }
    nvb_setup& S = s.c_setup; std::memset(&S, 0, sizeof S);
    S.abi_version = NVB_ABI_VERSION; S.channels = s.channels; S.sample_rate = s.sample_rate; S.block_size[0] = s.bs[0]; S.block_size[1] = s.bs[1];
    S.n_books = (int)s.c_books.size(); S.n_floors = (int)s.c_floors.size(); S.n_residues = (int)s.c_residues.size();
    S.n_mappings = (int)s.c_mappings.size(); S.n_modes = (int)s.c_modes.size();
    S.books = s.c_books.data(); S.vq_floats = s.c_vq.data(); S.n_vq_floats = (int64_t)s.c_vq.size();
    S.floors = s.c_floors.data(); S.residues = s.c_residues.data(); S.mappings = s.c_mappings.data(); S.modes = s.c_modes.data();
}

static void open_common(nvh_stream& s) {
    if (s.packets.size() < 3) throw DataError("Could not find Vorbis data to decode.");
    { Bits b(s.bytes.data() + s.packets[0].off, s.packets[0].size); parse_id_header(b, s); }
    {   // comment header: only the signature matters here (StreamDecoder.cs:206-224)
        Bits b(s.bytes.data() + s.packets[1].off, s.packets[1].size);
        static const uint8_t sig[7] = {0x03, 'v', 'o', 'r', 'b', 'i', 's'};
        for (int i = 0; i < 7; i++) if (b.read(8) != sig[i]) throw DataError("comment header signature");
    }
    { Bits b(s.bytes.data() + s.packets[2].off, s.packets[2].size); parse_setup_header(b, s); }
    export_setup(s);
    s.first_audio = s.next_packet = 3;
}

// ---- one audio packet -> boundary record ----------------------------------------------------------------
struct Scratch { std::vector<int16_t> posts; std::vector<uint8_t> classes; std::vector<uint16_t> entries; std::vector<float> floor0; std::vector<UnpackedFrame> frames; };

static void unpack_packet(const nvh_stream& s, const PacketRef& pr, Scratch& out) {
    UnpackedFrame uf; std::memset(&uf.f, 0, sizeof uf.f);
    uf.has_granule = pr.flags & 1; uf.granule = pr.granule; uf.eos = (pr.flags & 2) != 0; uf.resync = (pr.flags & 4) != 0;
    const int C = s.channels;
    const size_t posts_at = out.posts.size();
    out.posts.resize(posts_at + (size_t)C * s.post_stride, 0);
    const size_t f0_at = out.floor0.size();
    out.floor0.resize(f0_at + (size_t)C * s.f0_stride, 0.f);
    uf.f.status = NVB_FRAME_FAILED;
    uf.f.classes_off = (uint32_t)out.classes.size(); uf.f.entries_off = (uint32_t)out.entries.size();
    Bits b(s.bytes.data() + pr.off, pr.size);
    auto done = [&]() { out.frames.push_back(uf); };

    if (b.bit()) return done();                                                                       // not an audio packet, StreamDecoder.cs:490
    const int mode_idx = (int)b.read(s.mode_bits);
    if (mode_idx >= (int)s.modes.size()) return done();
    const ModeDef& mode = s.modes[(size_t)mode_idx];
    // Mode.GetPacketInfo (Mode.cs:119-151)
    if (b.short_) return done();
    int window = 0, start, valid, total;
    const int N = s.bs[mode.long_block ? 1 : 0];
    if (mode.long_block) {
        const bool prev = b.bit(), next = b.bit();
        window = (prev ? 1 : 0) + (next ? 2 : 0);
        const int pn = prev ? s.bs[1] : s.bs[0], nn = next ? s.bs[1] : s.bs[0];
        start = N / 4 - pn / 4; total = N / 4 * 3 + nn / 4; valid = total - nn / 4 * 2;                // Mode.cs:102-117
    } else { start = 0; valid = N / 2; total = N; }
    uf.f.status = NVB_FRAME_OK; uf.f.mode = (uint8_t)mode_idx; uf.f.window = (uint8_t)window;
    uf.f.start = start; uf.f.valid = valid; uf.f.total = total;

    const MappingDef& map = s.mappings[(size_t)mode.mapping];
    // floors: Floor1.Unpack per channel (Floor1.cs:135-184)
    uint32_t live = 0;
    for (int c = 0; c < C; c++) {
        const Floor1Def& f = s.floors[(size_t)map.ch_floor[(size_t)c]];
        int16_t* dst = out.posts.data() + posts_at + (size_t)c * s.post_stride;
        if (s.floor_type[(size_t)map.ch_floor[(size_t)c]] == 0) {
            // Floor0.Unpack (Floor0.cs:98-150): amplitude, book number, LSP coefficients as VQ vectors, then the "averaging"
            const Floor0Def& z = s.floors0[(size_t)map.ch_floor[(size_t)c]];
            float* pl = out.floor0.data() + f0_at + (size_t)c * s.f0_stride;
            float amp = (float)b.read(z.amp_bits);
            if (amp > 0.f) {
                amp = amp / (float)z.amp_div * (float)z.amp_ofs;
                const uint32_t book_num = (uint32_t)b.read(z.book_bits);
                if (book_num >= z.books.size()) amp = 0.f;
                else {
                    const Book& bk = s.books[(size_t)z.books[book_num]];
                    for (int i = 0; i < z.order && amp > 0.f;) {
                        const int e = bk.decode(b);
                        if (e < 0) { amp = 0.f; break; }
                        for (int j = 0; i < z.order && j < bk.dims; j++, i++) pl[1 + i] = bk.table[(size_t)e * bk.dims + j];
                    }
                    if (amp > 0.f) {
                        float last = 0.f;
                        for (int j = 0; j < z.order;) {
                            for (int k = 0; j < z.order && k < bk.dims; j++, k++) pl[1 + j] += last;
                            last = pl[j];                                                             // Coeff[j - 1]
                        }
                    }
                }
            }
            pl[0] = amp;
            dst[0] = 0;
            if (amp > 0.f) live |= 1u << c;
            continue;
        }
        int count = 0;
        if (b.bit()) {
            count = 2;
            dst[1] = (int16_t)b.read(f.ybits); dst[2] = (int16_t)b.read(f.ybits);
            bool failed = false;
            for (size_t p = 0; p < f.part_class.size() && !failed; p++) {
                const int cls = f.part_class[p], cdim = f.class_dims[(size_t)cls], cbits = f.class_subs[(size_t)cls];
                uint32_t cval = 0;
                if (cbits > 0) {
                    const int v = s.books[(size_t)f.class_master[(size_t)cls]].decode(b);
                    if (v < 0) { failed = true; break; }
                    cval = (uint32_t)v;
                }
                for (int k = 0; k < cdim; k++) {
                    const int bk = f.sub_books[(size_t)cls][cval & ((1u << cbits) - 1)];
                    cval >>= cbits;
                    int y = 0;
                    if (bk >= 0) { y = s.books[(size_t)bk].decode(b); if (y < 0) { failed = true; break; } }
                    dst[1 + count] = (int16_t)y;
                    ++count;
                }
            }
            if (failed) count = 0;                                                                    // "use nothing", Floor1.cs:155-174
        }
        dst[0] = (int16_t)count;
        if (count > 0) live |= 1u << c;
    }
    // energy flags (Mapping.cs:105-119): noExecute is taken before the coupling propagation
    const uint32_t all_ch = C >= 32 ? 0xffffffffu : (1u << C) - 1u;
    const uint32_t no_exec = ~live & all_ch;
    uint32_t exec = live;
    for (size_t k = 0; k < map.mag.size(); k++)
        if (((exec >> map.ang[k]) | (exec >> map.mag[k])) & 1u) exec |= (1u << map.ang[k]) | (1u << map.mag[k]);
    uf.f.exec_mask = exec;

    // residue (Residue0.Decode, Residue0.cs:119-178): runs when any channel is live, over all streams
    const ResidueDef& r = s.residues[(size_t)map.residue0];
    const int span = (r.type == 2 ? N * C : N) / 2;
    const int nn = std::min(r.end, span) - r.begin;
    if (nn > 0 && no_exec != all_ch) {
        uf.f.res_decoded = 1;
        const int P = nn / r.psize, S = r.streams;
        const Book& cb = s.books[(size_t)r.class_book];
        const size_t cls_at = out.classes.size();
        out.classes.resize(cls_at + (size_t)S * P, 0);
        uf.n_classes = S * P;
        bool stop = false;
        std::vector<uint16_t> pending;
        for (int stage = 0; stage < r.stages && !stop; stage++) {
            for (int p = 0; p < P && !stop;) {
                if (stage == 0) {
                    for (int st = 0; st < S; st++) {
                        const int w = cb.decode(b);
                        if (w < 0 || w >= (int)r.class_digits.size()) { stop = true; break; }
                        for (int k = 0; k < cb.dims && p + k < P; k++) out.classes[cls_at + (size_t)st * P + p + k] = r.class_digits[(size_t)w][(size_t)k];
                    }
                    if (stop) break;
                }
                for (int k = 0; k < cb.dims && p < P && !stop; k++, p++) {
                    for (int st = 0; st < S && !stop; st++) {
                        const int cls = out.classes[cls_at + (size_t)st * P + p];
                        if (!((r.cascade[cls] >> stage) & 1)) continue;
                        const int bk = r.books[cls][stage];
                        if (bk < 0) continue;
                        const Book& vb = s.books[(size_t)bk];
                        if (r.type == 0) {
                            // all of a partition's entries are read before any is used (Residue0.cs:186-192)
                            const int steps = r.psize / vb.dims;
                            pending.clear();
                            for (int q = 0; q < steps; q++) { const int e = vb.decode(b); if (e < 0) { stop = true; break; } pending.push_back((uint16_t)e); }
                            if (!stop) out.entries.insert(out.entries.end(), pending.begin(), pending.end());
                        } else {
                            for (int q = 0; q < r.psize; q += vb.dims) {                              // Residue1.cs:12-23, Residue2.cs:29-44
                                const int e = vb.decode(b);
                                if (e < 0) { stop = true; break; }
                                out.entries.push_back((uint16_t)e);
                            }
                        }
                    }
                }
            }
        }
    }
    uf.f.entry_count = (uint32_t)(out.entries.size() - uf.f.entries_off);
    done();
}


// Stream-order pass shared by nvh_unpack and nvh_packet_batch: the bookkeeping of StreamDecoder.ReadNextPacket / Read that
// lives on the host -- sample position (re-based on the first granule, StreamDecoder.cs:358-363), the EOS trim (:429-437) and
// the drain record when the provider runs dry (:352-356,476-480).  Appends to s->o_frames (and pads o_posts / o_floor0 for
// the drain record).
// One packet of the stream-order bookkeeping; returns the samples per channel the packet makes available.
static int order_step(nvh_stream* s, UnpackedFrame& uf) {
    nvb_frame& f = uf.f;
    if (uf.resync) s->has_position = false;                                                       // StreamDecoder.cs:485-488
    s->eos_found |= uf.eos;
    int emitted;
    if (f.status != NVB_FRAME_OK) {
        s->prev_end = s->prev_stop;                                                               // drain, StreamDecoder.cs:352-356
        emitted = s->have_prev ? std::max(0, s->prev_end - s->prev_start) : 0;
        s->prev_start = s->prev_end;
    } else {
        if (uf.has_granule && uf.eos) {
            const int64_t actual_end = s->position + f.valid - f.start;
            const int diff = (int)(uf.granule - actual_end);
            if (diff < 0) f.valid += diff;
        }
        if (s->prev_end > 0) s->prev_start = f.start;
        else if (!s->have_prev) s->prev_start = f.valid;
        s->prev_end = f.valid; s->prev_stop = f.total; s->have_prev = true;
        emitted = std::max(0, s->prev_end - s->prev_start);
        if (s->prev_end < s->prev_start) s->prev_start = s->prev_end;
        s->prev_start = s->prev_end;
    }
    s->position += emitted;
    if (uf.has_granule && !s->has_position && f.status == NVB_FRAME_OK) { s->has_position = true; s->position = uf.granule; }
    return emitted;
}

static void stream_order_pass(nvh_stream* s, std::vector<UnpackedFrame>& metas, size_t next_packet, bool wanted_more, int32_t* end_of_stream) {
    for (UnpackedFrame& uf : metas) { order_step(s, uf); s->o_frames.push_back(uf.f); }
    s->next_packet = next_packet;
    if (s->next_packet >= s->packets.size() && wanted_more && (!s->forward || s->input_final)) {
        // the provider ran dry: DecodeNextPacket returns null with isEndOfStream = true (StreamDecoder.cs:476-480)
        if (!s->eos_found) {
            nvb_frame f; std::memset(&f, 0, sizeof f); f.status = NVB_FRAME_FAILED;
            f.classes_off = (uint32_t)s->o_classes.size(); f.entries_off = (uint32_t)s->o_entries.size();
            s->o_frames.push_back(f);
            s->o_posts.resize(s->o_posts.size() + (size_t)s->channels * s->post_stride, 0);
            s->o_floor0.resize(s->o_floor0.size() + (size_t)s->channels * s->f0_stride, 0.f);
            s->prev_end = s->prev_stop;
            if (s->have_prev) s->position += std::max(0, s->prev_end - s->prev_start);
            s->prev_start = s->prev_end;
            s->eos_found = true;
        }
        if (end_of_stream) *end_of_stream = 1;
    } else if (s->eos_found && end_of_stream) {
        *end_of_stream = 1;
    }
}

// What Mode.GetPacketInfo reads from the first bits of an audio packet (Mode.cs:119-151) -- the head of unpack_packet, without
// any Huffman decoding: status, mode, window flags, nominal start / valid / total.
static void packet_header(const nvh_stream& s, const PacketRef& pr, UnpackedFrame& uf) {
    std::memset(&uf.f, 0, sizeof uf.f);
    uf.has_granule = pr.flags & 1; uf.granule = pr.granule; uf.eos = (pr.flags & 2) != 0; uf.resync = (pr.flags & 4) != 0;
    uf.f.status = NVB_FRAME_FAILED;
    Bits b(s.bytes.data() + pr.off, pr.size);
    if (b.bit()) return;                                                                              // not an audio packet, StreamDecoder.cs:490
    const int mode_idx = (int)b.read(s.mode_bits);
    if (mode_idx >= (int)s.modes.size()) return;
    const ModeDef& mode = s.modes[(size_t)mode_idx];
    if (b.short_) return;
    int window = 0, start, valid, total;
    const int N = s.bs[mode.long_block ? 1 : 0];
    if (mode.long_block) {
        const bool prev = b.bit(), next = b.bit();
        window = (prev ? 1 : 0) + (next ? 2 : 0);
        const int pn = prev ? s.bs[1] : s.bs[0], nn = next ? s.bs[1] : s.bs[0];
        start = N / 4 - pn / 4; total = N / 4 * 3 + nn / 4; valid = total - nn / 4 * 2;                // Mode.cs:102-117
    } else { start = 0; valid = N / 2; total = N; }
    uf.f.status = NVB_FRAME_OK; uf.f.mode = (uint8_t)mode_idx; uf.f.window = (uint8_t)window;
    uf.f.start = start; uf.f.valid = valid; uf.f.total = total;
}

// The unpack tables of the setup (csrc/nvb_unpack_tables.h).
static int build_unpack_tables(nvh_stream& s) {
    using namespace nvbu;
    for (const MappingDef& m : s.mappings) if (m.submaps != 1) { s.err = "multi-submap mappings are not supported"; return NVB_ERR_UNSUPPORTED; }
    if (s.channels > NVB_MAX_CHANNELS) { s.err = "too many channels"; return NVB_ERR_UNSUPPORTED; }
    std::vector<UBook> books; std::vector<uint32_t> roots; std::vector<ULong> longs;
    for (const Book& b : s.books) {
        UBook u; std::memset(&u, 0, sizeof u);
        u.dims = b.dims; u.entries = b.entries; u.root_bits = b.root_bits; u.decodable = b.decodable ? 1 : 0;
        u.root_off = (uint32_t)roots.size(); u.long_off = (uint32_t)longs.size(); u.n_long = (int32_t)b.longs.size();
        if (b.decodable) {
            if (b.entries > (1 << 24)) { s.err = "codebook too large for the device tables"; return NVB_ERR_UNSUPPORTED; }
            for (const Book::Root& r : b.root) {
                if (r.len) roots.push_back(((uint32_t)r.len << 24) | ((uint32_t)r.value & 0xffffffu));
                else roots.push_back(r.value >= 0 ? (uint32_t)r.value + 1u : 0u);
            }
            for (const Book::Long& l : b.longs) longs.push_back(ULong{l.code, l.value, l.next >= 0 ? l.next + 1 : 0, l.len});
        }
        books.push_back(u);
    }
    std::vector<UFloor1> floors;
    for (size_t fi = 0; fi < s.floors.size(); fi++) {
        const Floor1Def& f = s.floors[fi];
        UFloor1 u; std::memset(&u, 0, sizeof u);
        if (s.floor_type[fi] == 0) {                                        // Floor0.Init (Floor0.cs:28-51): what Floor0.Unpack reads
            const Floor0Def& z = s.floors0[fi];
            if (z.books.size() > 16) { s.err = "floor 0 book list outside the device tables"; return NVB_ERR_UNSUPPORTED; }
            u.type = 0; u.f0.order = z.order; u.f0.amp_bits = z.amp_bits; u.f0.amp_div = z.amp_div; u.f0.amp_ofs = z.amp_ofs; u.f0.book_bits = z.book_bits;
            u.f0.n_books = (int32_t)z.books.size();
            for (size_t k = 0; k < z.books.size(); k++) u.f0.books[k] = (int16_t)z.books[k];
            floors.push_back(u);
            continue;
        }
        u.type = 1; u.n_parts = (int32_t)f.part_class.size(); u.ybits = f.ybits; u.n_posts = (int32_t)f.x.size();
        if (u.n_parts > 32 || f.class_dims.size() > 16) { s.err = "floor 1 structure outside the device tables"; return NVB_ERR_UNSUPPORTED; }
        for (size_t k = 0; k < f.part_class.size(); k++) u.part_class[k] = (uint8_t)f.part_class[k];
        for (size_t c = 0; c < f.class_dims.size(); c++) {
            u.class_dims[c] = (uint8_t)f.class_dims[c]; u.class_subs[c] = (uint8_t)f.class_subs[c]; u.class_master[c] = (int16_t)f.class_master[c];
            for (int k = 0; k < 8; k++) u.sub_books[c][k] = k < (int)f.sub_books[c].size() ? (int16_t)f.sub_books[c][(size_t)k] : (int16_t)-1;
        }
        floors.push_back(u);
    }
    std::vector<UResidue> residues; std::vector<uint8_t> digits;
    for (const ResidueDef& r : s.residues) {
        UResidue u; std::memset(&u, 0, sizeof u);
        const Book& cb = s.books[(size_t)r.class_book];
        u.type = r.type; u.begin = r.begin; u.end = r.end; u.psize = r.psize; u.nclass = r.nclass; u.class_book = r.class_book; u.stages = r.stages;
        u.cdims = cb.dims; u.partvals = (int32_t)r.class_digits.size(); u.digits_off = (uint32_t)digits.size();
        for (const auto& row : r.class_digits) digits.insert(digits.end(), row.begin(), row.end());
        for (int c = 0; c < 64; c++) { u.cascade[c] = r.cascade[c]; for (int k = 0; k < 8; k++) u.books[c][k] = (int16_t)r.books[c][k]; }
        residues.push_back(u);
    }
    // per-frame strides of the device-produced records: the largest any mode can need
    int cls_stride = 1, ent_stride = 1;
    for (const ModeDef& m : s.modes) {
        const MappingDef& map = s.mappings[(size_t)m.mapping];
        const ResidueDef& r = s.residues[(size_t)map.residue0];
        const int N = s.bs[m.long_block ? 1 : 0];
        const int span = (r.type == 2 ? N * s.channels : N) / 2;
        const int nn = std::min(r.end, span) - r.begin;
        const int P = nn > 0 ? nn / r.psize : 0, S = r.streams;
        cls_stride = std::max(cls_stride, S * P);
        long long per_partition = 0;                                       // worst class: entries of all its coded stages
        for (int c = 0; c < r.nclass; c++) {
            long long e = 0;
            for (int st = 0; st < r.stages; st++) {
                if (!((r.cascade[c] >> st) & 1) || r.books[c][st] < 0) continue;
                const int d = s.books[(size_t)r.books[c][st]].dims;
                e += r.type == 0 ? r.psize / d : (r.psize + d - 1) / d;
            }
            per_partition = std::max(per_partition, e);
        }
        const long long need = per_partition * P * S;
        if (need > (1 << 22)) { s.err = "a frame can need more VQ entries than the device unpacker provides for"; return NVB_ERR_UNSUPPORTED; }
        ent_stride = std::max(ent_stride, (int)need);
    }
    cls_stride = (cls_stride + 15) & ~15; ent_stride = (ent_stride + 7) & ~7;
    std::vector<UMapping> mappings;
    for (const MappingDef& m : s.mappings) {
        UMapping u; std::memset(&u, 0, sizeof u);
        if (m.mag.size() > (size_t)UNPACK_MAX_COUPLING) { s.err = "too many coupling steps"; return NVB_ERR_UNSUPPORTED; }
        u.n_coupling = (int32_t)m.mag.size(); u.floor = m.floor0; u.residue = m.residue0;
        for (size_t k = 0; k < m.mag.size(); k++) { u.mag[k] = (uint8_t)m.mag[k]; u.ang[k] = (uint8_t)m.ang[k]; }
        mappings.push_back(u);
    }
    std::vector<UMode> modes;
    for (const ModeDef& m : s.modes) modes.push_back(UMode{m.long_block ? 1 : 0, m.mapping});

    UHeader h; std::memset(&h, 0, sizeof h);
    h.magic = UNPACK_MAGIC; h.version = 2; h.f0_stride = s.f0_stride; h.channels = s.channels; h.bs[0] = s.bs[0]; h.bs[1] = s.bs[1]; h.mode_bits = s.mode_bits;
    h.n_books = (int32_t)books.size(); h.n_floors = (int32_t)floors.size(); h.n_residues = (int32_t)residues.size();
    h.n_mappings = (int32_t)mappings.size(); h.n_modes = (int32_t)modes.size();
    h.post_stride = s.post_stride; h.cls_stride = cls_stride; h.ent_stride = ent_stride;
    h.n_roots = (uint32_t)roots.size(); h.n_longs = (uint32_t)longs.size(); h.n_digits = (uint32_t)digits.size();
    std::vector<uint8_t>& blob = s.u_blob;
    blob.assign(sizeof(UHeader), 0);
    auto put = [&](const void* src, size_t bytes) { const size_t at = (blob.size() + 15) & ~size_t(15); blob.resize(at + bytes, 0); if (bytes) std::memcpy(blob.data() + at, src, bytes); return (uint64_t)at; };
    h.off_books = put(books.data(), books.size() * sizeof(UBook));
    h.off_roots = put(roots.data(), roots.size() * sizeof(uint32_t));
    h.off_longs = put(longs.data(), longs.size() * sizeof(ULong));
    h.off_floors = put(floors.data(), floors.size() * sizeof(UFloor1));
    h.off_residues = put(residues.data(), residues.size() * sizeof(UResidue));
    h.off_digits = put(digits.data(), digits.size());
    h.off_mappings = put(mappings.data(), mappings.size() * sizeof(UMapping));
    h.off_modes = put(modes.data(), modes.size() * sizeof(UMode));
    blob.resize((blob.size() + 15) & ~size_t(15), 0);
    h.total_bytes = blob.size();
    std::memcpy(blob.data(), &h, sizeof h);
    return NVB_OK;
}

}  // namespace nvh

// =====================================================================================================
extern "C" {

const char* nvh_last_error(nvh_stream* s) { return s ? s->err.c_str() : g_open_error.c_str(); }

static int finish_open(std::unique_ptr<nvh_stream>& s, nvh_stream** out) {
    try { open_common(*s); }
    catch (const std::exception& e) { g_open_error = e.what(); return NVB_ERR_DATA; }
    *out = s.release();
    return NVB_OK;
}

int nvh_open_ogg(const uint8_t* data, size_t len, nvh_stream** out) {
    if (!data || !out) { g_open_error = "NULL argument"; return NVB_ERR_ARG; }
    *out = nullptr;
    std::unique_ptr<nvh_stream> s(new nvh_stream());
    try { demux_ogg(data, len, *s); }
    catch (const std::exception& e) { g_open_error = e.what(); return NVB_ERR_DATA; }
    return finish_open(s, out);
}

int nvh_open_forward(nvh_stream** out) {
    if (!out) { g_open_error = "NULL argument"; return NVB_ERR_ARG; }
    nvh_stream* s = new (std::nothrow) nvh_stream();
    if (!s) return NVB_ERR_NOMEM;
    s->forward = true; s->input_final = false; s->headers_done = false;
    *out = s;
    return NVB_OK;
}

int64_t nvh_feed(nvh_stream* s, const uint8_t* data, size_t len, int end_of_input) {
    if (!s || !s->forward || (len > 0 && !data)) return NVB_ERR_ARG;
    if (s->input_final) { s->err = "nvh_feed after the end of the input"; return NVB_ERR_STATE; }
    try {
        s->raw.insert(s->raw.end(), data, data + len);
        if (end_of_input) s->input_final = true;
        s->demux.scan(s->raw.data(), s->raw.size(), s->input_final, *s);
        s->demux.emit(s->raw.data(), s->input_final, *s);
        if (!s->headers_done && s->packets.size() >= 3) { open_common(*s); s->headers_done = true; }
        if (!s->headers_done && s->input_final) { s->err = "Could not find Vorbis data to decode."; return NVB_ERR_DATA; }
        return s->headers_done ? (int64_t)(s->packets.size() - s->first_audio) : 0;
    } catch (const std::bad_alloc&) { s->err = "nvh_feed: out of memory"; return NVB_ERR_NOMEM; }
    catch (const std::exception& e) { s->err = e.what(); return NVB_ERR_DATA; }
}

int nvh_ogg_stream_count(const uint8_t* data, size_t len) {
    if (!data) return NVB_ERR_ARG;
    try { return (int)ogg_serials(data, len).size(); } catch (...) { return NVB_ERR_NOMEM; }
}

int nvh_open_ogg_stream(const uint8_t* data, size_t len, int stream_index, nvh_stream** out) {
    if (!data || !out || stream_index < 0) { g_open_error = "NULL argument / negative stream index"; return NVB_ERR_ARG; }
    *out = nullptr;
    std::unique_ptr<nvh_stream> s(new nvh_stream());
    try {
        const std::vector<uint32_t> serials = ogg_serials(data, len);
        if ((size_t)stream_index >= serials.size()) { g_open_error = "the container has no logical stream with that index"; return NVB_ERR_ARG; }
        demux_ogg(data, len, *s, true, serials[(size_t)stream_index]);
    } catch (const std::exception& e) { g_open_error = e.what(); return NVB_ERR_DATA; }
    return finish_open(s, out);
}

int nvh_open_packets(const uint8_t* data, const int64_t* sizes, const int64_t* granules, const uint8_t* flags, int64_t n, nvh_stream** out) {
    if (!data || !sizes || !granules || !flags || !out || n < 0) { g_open_error = "NULL argument"; return NVB_ERR_ARG; }
    *out = nullptr;
    std::unique_ptr<nvh_stream> s(new nvh_stream());
    size_t off = 0;
    for (int64_t i = 0; i < n; i++) {
        if (sizes[i] < 0) { g_open_error = "negative packet size"; return NVB_ERR_ARG; }
        s->packets.push_back(PacketRef{off, (size_t)sizes[i], granules[i], flags[i]});
        if (flags[i] & 2) s->has_eos = true;
        off += (size_t)sizes[i];
    }
    s->bytes.assign(data, data + off);
    return finish_open(s, out);
}

int nvh_close(nvh_stream* s) { delete s; return NVB_OK; }

int nvh_get_info(nvh_stream* s, nvh_info* o) {
    if (!s || !o) return NVB_ERR_ARG;
    std::memset(o, 0, sizeof *o);
    o->channels = s->channels; o->sample_rate = s->sample_rate; o->block_size[0] = s->bs[0]; o->block_size[1] = s->bs[1];
    o->n_books = (int)s->books.size(); o->n_floors = (int)s->floors.size(); o->n_residues = (int)s->residues.size();
    o->n_mappings = (int)s->mappings.size(); o->n_modes = (int)s->modes.size(); o->post_stride = s->post_stride; o->floor0_stride = s->f0_stride;
    o->n_packets = (int64_t)s->packets.size(); o->n_audio_packets = (int64_t)(s->packets.size() - s->first_audio);
    o->last_granule = -1;
    for (const PacketRef& p : s->packets) if (p.flags & 1) o->last_granule = p.granule;
    o->has_eos = s->has_eos ? 1 : 0;
    return NVB_OK;
}

const nvb_setup* nvh_setup(nvh_stream* s) { return s ? &s->c_setup : nullptr; }

int64_t nvh_packet_size(nvh_stream* s, int64_t i) { return (!s || i < 0 || i >= (int64_t)s->packets.size()) ? -1 : (int64_t)s->packets[(size_t)i].size; }
int nvh_packet_get(nvh_stream* s, int64_t i, uint8_t* dst, int64_t* granule, int32_t* flags) {
    if (!s || i < 0 || i >= (int64_t)s->packets.size()) return NVB_ERR_ARG;
    const PacketRef& p = s->packets[(size_t)i];
    if (dst) std::memcpy(dst, s->bytes.data() + p.off, p.size);
    if (granule) *granule = p.granule;
    if (flags) *flags = p.flags;
    return NVB_OK;
}

int nvh_rewind(nvh_stream* s) {
    if (!s) return NVB_ERR_ARG;
    s->next_packet = s->first_audio;
    s->have_prev = false; s->prev_start = s->prev_end = s->prev_stop = 0;
    s->position = 0; s->has_position = false; s->eos_found = false;
    return NVB_OK;
}

int64_t nvh_unpack(nvh_stream* s, int64_t count, int threads, nvb_batch* out, int32_t* end_of_stream) {
    if (!s || !out || count < 0) return NVB_ERR_ARG;
    if (end_of_stream) *end_of_stream = 0;
    const size_t lo = s->next_packet;
    const size_t avail = s->packets.size() - lo;
    const size_t n = std::min<size_t>((size_t)count, avail);
    int T = std::max(1, threads);
    if ((size_t)T > n) T = (int)std::max<size_t>(1, n);
    // nothing may unwind through the C boundary (or out of a worker thread: std::terminate): the first failure is kept
    // in s->err and reported as a status
    std::atomic<int> failed{NVB_OK};
    std::vector<Scratch> parts;
    try {
    parts.resize((size_t)T);
    auto work = [&](int t) {
        try {
            const size_t a = lo + n * (size_t)t / (size_t)T, b = lo + n * (size_t)(t + 1) / (size_t)T;
            for (size_t i = a; i < b; i++) unpack_packet(*s, s->packets[i], parts[(size_t)t]);
        } catch (const std::bad_alloc&) { failed = NVB_ERR_NOMEM; }
        catch (...) { failed = NVB_ERR_DATA; }
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        try { for (int t = 0; t < T; t++) th.emplace_back(work, t); }
        catch (...) { failed = NVB_ERR_NOMEM; }                                // thread creation failed: join what started
        for (auto& x : th) x.join();
    }
    if (failed != NVB_OK) { s->err = failed == NVB_ERR_NOMEM ? "nvh_unpack: out of memory" : "nvh_unpack: internal error while unpacking"; return failed; }

    s->o_frames.clear(); s->o_posts.clear(); s->o_classes.clear(); s->o_entries.clear(); s->o_floor0.clear();
    std::vector<UnpackedFrame> metas;
    for (Scratch& p : parts) {
        const uint32_t coff = (uint32_t)s->o_classes.size(), eoff = (uint32_t)s->o_entries.size();
        for (UnpackedFrame& uf : p.frames) { uf.f.classes_off += coff; uf.f.entries_off += eoff; metas.push_back(uf); }
        s->o_posts.insert(s->o_posts.end(), p.posts.begin(), p.posts.end());
        s->o_classes.insert(s->o_classes.end(), p.classes.begin(), p.classes.end());
        s->o_entries.insert(s->o_entries.end(), p.entries.begin(), p.entries.end());
        s->o_floor0.insert(s->o_floor0.end(), p.floor0.begin(), p.floor0.end());
    }
    stream_order_pass(s, metas, lo + n, (size_t)count > n, end_of_stream);
    std::memset(out, 0, sizeof *out);
    out->n_frames = (int32_t)s->o_frames.size();
    out->frames = s->o_frames.data(); out->posts = s->o_posts.data();
    out->classes = s->o_classes.data(); out->n_classes = (int64_t)s->o_classes.size();
    out->entries = s->o_entries.data(); out->n_entries = (int64_t)s->o_entries.size();
    out->floor0 = s->o_floor0.empty() ? nullptr : s->o_floor0.data();
    return (int64_t)s->o_frames.size();
    } catch (const std::bad_alloc&) { s->err = "nvh_unpack: out of memory"; return NVB_ERR_NOMEM; }
    catch (const std::exception& e) { s->err = std::string("nvh_unpack: ") + e.what(); return NVB_ERR_DATA; }
}

// Header-only walk of the stream from its start (no Huffman decoding): the samples per channel every audio packet makes
// available when the stream is decoded from the beginning.  Leaves the stream rewound.
static void emitted_per_packet(nvh_stream* s, std::vector<int>& emitted, std::vector<int64_t>& pos_before, std::vector<uint8_t>& has_pos_before, int64_t* tail_drain = nullptr) {
    nvh_rewind(s);
    const size_t n = s->packets.size() - s->first_audio;
    emitted.assign(n, 0); pos_before.assign(n, 0); has_pos_before.assign(n, 0);
    for (size_t i = 0; i < n; i++) {
        UnpackedFrame uf; packet_header(*s, s->packets[s->first_audio + i], uf);
        pos_before[i] = s->position; has_pos_before[i] = s->has_position ? 1 : 0;
        emitted[i] = order_step(s, uf);
    }
    // no end-of-stream packet: the provider runs dry and the last tail is drained (StreamDecoder.cs:352-356,476-480)
    if (tail_drain) *tail_drain = (!s->eos_found && s->have_prev) ? std::max(0, s->prev_stop - s->prev_start) : 0;
    nvh_rewind(s);
}

int64_t nvh_total_samples(nvh_stream* s) {
    if (!s) return NVB_ERR_ARG;
    if (s->next_packet != s->first_audio) return NVB_ERR_STATE;             // only on a rewound stream: the walk resets the decode cursor
    try {
        std::vector<int> e; std::vector<int64_t> p; std::vector<uint8_t> h; int64_t tail = 0;
        emitted_per_packet(s, e, p, h, &tail);
        int64_t total = tail; for (int v : e) total += v;
        return total;
    } catch (...) { return NVB_ERR_NOMEM; }
}

int nvh_seek(nvh_stream* s, int64_t sample_position, int64_t* skip_samples) {
    if (!s || !skip_samples || sample_position < 0) return NVB_ERR_ARG;
    if (s->forward && !s->input_final) { s->err = "a forward-only stream cannot seek before its input has ended"; return NVB_ERR_STATE; }
    try {
        std::vector<int> e; std::vector<int64_t> pos; std::vector<uint8_t> hp;
        emitted_per_packet(s, e, pos, hp);
        *skip_samples = 0;
        if (sample_position == 0 || e.empty()) return NVB_OK;                // the looping case: restart at the first audio packet
        // packet k: the first one whose samples reach past the target
        int64_t acc = 0; size_t k = 0;
        while (k < e.size() && acc + e[k] <= sample_position) { acc += e[k]; ++k; }
        if (k >= e.size()) {
            // at or beyond the last packet: only the drained tail (if any) remains -- position the cursor on the last packet
            k = e.size() - 1; acc -= e[k];
        }
        // pre-roll (StreamDecoder.cs:598-617): decoding restarts one packet early; that block only leaves its tail, the next one
        // overlaps onto it, and `skip` samples of its output are dropped (_prevPacketStart += rollForward)
        const size_t start = k > 0 ? k - 1 : 0;
        int64_t lead = 0;                                                    // samples packets start..k-1 would emit after a restart at `start`
        // (the restart block emits nothing; with start == k-1 nothing lies between it and k)
        s->next_packet = s->first_audio + start;
        s->have_prev = false; s->prev_start = s->prev_end = s->prev_stop = 0; s->eos_found = false;
        // the sample position the continuous decode has when packet k starts to emit: the restart block does not advance it
        s->position = pos[k]; s->has_position = hp[k] != 0;
        if (start == k) {                                                    // k == 0: decoding from the very first packet is the continuous decode
            s->position = 0; s->has_position = false;
        }
        *skip_samples = sample_position - acc + lead;
        return NVB_OK;
    } catch (const std::bad_alloc&) { s->err = "nvh_seek: out of memory"; return NVB_ERR_NOMEM; }
    catch (const std::exception& ex) { s->err = std::string("nvh_seek: ") + ex.what(); return NVB_ERR_DATA; }
}

int nvh_unpack_tables(nvh_stream* s, const void** blob, size_t* bytes) {
    if (!s || !blob || !bytes) return NVB_ERR_ARG;
    try {
        if (s->u_blob.empty()) { const int rc = build_unpack_tables(*s); if (rc != NVB_OK) { s->u_blob.clear(); return rc; } }
    } catch (const std::bad_alloc&) { s->err = "nvh_unpack_tables: out of memory"; return NVB_ERR_NOMEM; }
    catch (const std::exception& e) { s->err = std::string("nvh_unpack_tables: ") + e.what(); return NVB_ERR_DATA; }
    *blob = s->u_blob.data(); *bytes = s->u_blob.size();
    return NVB_OK;
}

int64_t nvh_packet_batch(nvh_stream* s, int64_t count, nvb_packet_batch* out, int32_t* end_of_stream) {
    if (!s || !out || count < 0) return NVB_ERR_ARG;
    if (end_of_stream) *end_of_stream = 0;
    try {
        const size_t lo = s->next_packet;
        const size_t n = std::min<size_t>((size_t)count, s->packets.size() - lo);
        std::vector<UnpackedFrame> metas(n);
        for (size_t i = 0; i < n; i++) packet_header(*s, s->packets[lo + i], metas[i]);
        s->o_frames.clear(); s->o_posts.clear(); s->o_classes.clear(); s->o_entries.clear(); s->o_floor0.clear();
        stream_order_pass(s, metas, lo + n, (size_t)count > n, end_of_stream);
        // packet payloads sit back to back in the stream's byte store: the batch points at them; a drain record (provider ran
        // dry) is an empty packet.  Offsets are relative to the first packet of the batch.
        const size_t base = n ? s->packets[lo].off : 0;
        s->o_offsets.assign(s->o_frames.size() + 1, 0);
        for (size_t i = 0; i < n; i++) s->o_offsets[i + 1] = (uint32_t)(s->packets[lo + i].off + s->packets[lo + i].size - base);
        for (size_t i = n; i < s->o_frames.size(); i++) s->o_offsets[i + 1] = s->o_offsets[i];
        if ((n ? s->packets[lo + n - 1].off + s->packets[lo + n - 1].size - base : 0) > 0xffffffffull) { s->err = "packet batch above 4 GiB"; return NVB_ERR_ARG; }
        std::memset(out, 0, sizeof *out);
        out->n_packets = (int32_t)s->o_frames.size();
        out->frames = s->o_frames.data();
        out->data = s->bytes.data() + base;
        out->offsets = s->o_offsets.data();
        return (int64_t)s->o_frames.size();
    } catch (const std::bad_alloc&) { s->err = "nvh_packet_batch: out of memory"; return NVB_ERR_NOMEM; }
    catch (const std::exception& e) { s->err = std::string("nvh_packet_batch: ") + e.what(); return NVB_ERR_DATA; }
}

}  // extern "C"
