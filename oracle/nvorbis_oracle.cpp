// =====================================================================================
// nvorbis_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A scalar C++17 restatement of the managed NVorbis decode path (reference @ /root/reference,
// C#, cannot be compiled in this image: no dotnet/mono).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
// (nvorbis_b200/) never links, imports or calls anything in oracle/.
//
// PARITY PINNING: the reference ships NO golden vectors, known-answer tests or expected
// outputs (NVorbis.sln lists only NVorbis + TestApp; TestFiles/*.ogg are inputs only).  This
// restatement is therefore pinned by (a) line-against-line review with the cited ranges,
// (b) structural invariants of the four fixtures (sample counts == last granule etc.),
// (c) the IMDCT closed form / TDAC identities, (d) an INDEPENDENT decoder: FFmpeg's native
// vorbis decoder (libavcodec, ctypes) agrees with this restatement to <= 6e-7 max-abs on all
// four fixtures (tests/ffmpeg_vorbis.py), see tests/test_oracle.py.  That proves the decode
// is correct Vorbis, not that it is bit-identical to the C# build: "parity unpinned" in the
// strict sense remains.  Parts no fixture reaches (Floor0, Residue0, lookup type 2,
// sequence_p, >2 channels) are synthetic-only.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared (no -ffast-math): float stays
// float, double stays double, no FMA contraction -- RyuJIT x64 SSE2 semantics.
//
// Every function cites the reference file:line it follows (paths relative to NVorbis/).
// =====================================================================================
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

struct InvalidData : std::runtime_error { using std::runtime_error::runtime_error; };

// ------------------------------------------------------------------------------------
// Utils.cs:5-59
// ------------------------------------------------------------------------------------
static int ilog(int x) { int cnt = 0; while (x > 0) { ++cnt; x >>= 1; } return cnt; }           // Utils.cs:5-14

static uint32_t BitReverse(uint32_t n, int bits = 32) {                                          // Utils.cs:16-28
    n = ((n & 0xAAAAAAAAu) >> 1) | ((n & 0x55555555u) << 1);
    n = ((n & 0xCCCCCCCCu) >> 2) | ((n & 0x33333333u) << 2);
    n = ((n & 0xF0F0F0F0u) >> 4) | ((n & 0x0F0F0F0Fu) << 4);
    n = ((n & 0xFF00FF00u) >> 8) | ((n & 0x00FF00FFu) << 8);
    n = (n >> 16) | (n << 16);
    int sh = 32 - bits;                      // C# masks shift counts to 5 bits (32 -> 0)
    return n >> (sh & 31);
}

static float ClipValue(float value, bool& clipped) {                                             // Utils.cs:30-43
    if (value > .99999994f) { clipped = true; return 0.99999994f; }
    if (value < -.99999994f) { clipped = true; return -0.99999994f; }
    return value;
}

static float ConvertFromVorbisFloat32(uint32_t bits) {                                           // Utils.cs:45-59
    int32_t sign = ((int32_t)bits >> 31);
    double exponent = (double)((int)((bits & 0x7fe00000u) >> 21) - 788);
    float mantissa = (float)(int32_t)(((bits & 0x1fffffu) ^ (uint32_t)sign) + (uint32_t)(sign & 1));
    return mantissa * (float)std::pow(2.0, exponent);
}

// ------------------------------------------------------------------------------------
// Bit reader: DataPacket.cs:150-283 over a contiguous packet (Ogg/Packet.cs:32-53 walks the
// packet's segments byte by byte; concatenating them first is equivalent).
// The 64-bit overflow byte (DataPacket.cs:190-193,233-243) can only engage for counts > 56;
// Vorbis never asks for more than 32 bits, so it is not restated (count is asserted).
// ------------------------------------------------------------------------------------
struct Packet {
    std::vector<uint8_t> data;
    bool isResync = false, isEndOfStream = false;
    bool hasGranule = false; int64_t granule = 0;
    // reader state
    size_t pos = 0; uint64_t bucket = 0; int bitCount = 0; int readBits_ = 0; bool isShort = false;

    int TotalBits() const { return (int)data.size() * 8; }
    int BitsRead() const { return readBits_; }
    int BitsRemaining() const { return TotalBits() - readBits_; }
    int ReadNextByte() { return pos < data.size() ? data[pos++] : -1; }
    void Reset() { pos = 0; bucket = 0; bitCount = 0; readBits_ = 0; }                           // DataPacket.cs:141-147

    uint64_t TryPeekBits(int count, int& bitsRead) {                                             // DataPacket.cs:166-208
        if (count < 0 || count > 56) throw std::out_of_range("TryPeekBits count");
        if (count == 0) { bitsRead = 0; return 0; }
        while (bitCount < count) {
            int val = ReadNextByte();
            if (val == -1) { bitsRead = bitCount; return bucket; }
            bucket = ((uint64_t)(val & 0xFF) << bitCount) | bucket;
            bitCount += 8;
        }
        uint64_t value = bucket & ((1ULL << count) - 1);
        bitsRead = count;
        return value;
    }

    void SkipBits(int count) {                                                                   // DataPacket.cs:214-283
        if (count <= 0) return;
        if (bitCount > count) {
            bucket = (count > 63) ? 0 : (bucket >> count);
            bitCount -= count; readBits_ += count;
        } else if (bitCount == count) {
            bucket = 0; bitCount = 0; readBits_ += count;
        } else {
            count -= bitCount; readBits_ += bitCount; bitCount = 0; bucket = 0;
            while (count > 8) {
                if (ReadNextByte() == -1) { count = 0; isShort = true; break; }
                count -= 8; readBits_ += 8;
            }
            if (count > 0) {
                int temp = ReadNextByte();
                if (temp == -1) { isShort = true; }
                else { bucket = (uint64_t)(temp >> count); bitCount = 8 - count; readBits_ += count; }
            }
        }
    }

    uint64_t ReadBits(int count) {                                                               // DataPacket.cs:149-159
        if (count == 0) return 0;
        int br; uint64_t v = TryPeekBits(count, br);
        SkipBits(count);
        return v;
    }
    bool ReadBit() { return ReadBits(1) == 1; }                                                  // Extensions.cs:60-63
};

// ------------------------------------------------------------------------------------
// Ogg container -> packets (seekable single-stream path only).
//   page sync + CRC:     Ogg/PageReaderBase.cs:33-70,227-292, Ogg/Crc.cs:8-37
//   lacing -> packets:   Ogg/PageReader.cs:27-93  (zero-length packets are dropped, :40-47,:77-85;
//                        pages with no packets are dropped, :131)
//   packet assembly:     Ogg/PacketProvider.cs:324-438 (granule only on the last packet that
//                        completes in a page :399-401; EOS :404-407)
//   EOS page:            Ogg/StreamPageReader.cs:72-75 (pages after it are ignored :47)
// Multi-stream containers, seeking and the forward-only reader are outside the hot path.
// ------------------------------------------------------------------------------------
struct Page {
    int64_t granule; uint8_t flags; int32_t serial; int32_t seq; bool isResync; bool isContinued;
    std::vector<std::vector<uint8_t>> packets;   // non-empty slices only
};

static uint32_t g_crcTable[256]; static bool g_crcInit = false;
static void crcInit() {                                                                          // Ogg/Crc.cs:8-21
    if (g_crcInit) return;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t s = i << 24;
        for (int j = 0; j < 8; ++j) s = (s << 1) ^ (s >= (1U << 31) ? 0x04c11db7u : 0);
        g_crcTable[i] = s;
    }
    g_crcInit = true;
}

static std::vector<Page> ReadPages(const uint8_t* d, size_t len) {
    crcInit();
    std::vector<Page> pages;
    size_t i = 0; bool resync = false; bool haveSerial = false; int32_t serial = 0; bool hasAll = false;
    int32_t lastSeq = 0;
    while (i + 27 <= len && !hasAll) {
        // VerifyHeader: "OggS", version 0 (Ogg/PageReaderBase.cs:72-111)
        if (!(d[i] == 0x4f && d[i + 1] == 0x67 && d[i + 2] == 0x67 && d[i + 3] == 0x53 && d[i + 4] == 0)) { ++i; resync = true; continue; }
        int segCnt = d[i + 26];
        if (i + 27 + segCnt > len) { ++i; resync = true; continue; }
        size_t dataLen = 0; for (int s = 0; s < segCnt; s++) dataLen += d[i + 27 + s];
        size_t pageLen = 27 + segCnt + dataLen;
        if (i + pageLen > len) { ++i; resync = true; continue; }
        uint32_t crc = 0;                                                                        // Ogg/PageReaderBase.cs:55-69
        for (size_t k = 0; k < pageLen; k++) {
            uint8_t b = (k >= 22 && k < 26) ? 0 : d[i + k];
            crc = (crc << 8) ^ g_crcTable[b ^ (crc >> 24)];
        }
        uint32_t want = (uint32_t)d[i + 22] | ((uint32_t)d[i + 23] << 8) | ((uint32_t)d[i + 24] << 16) | ((uint32_t)d[i + 25] << 24);
        if (crc != want) { ++i; resync = true; continue; }

        Page pg; pg.flags = d[i + 5];
        uint64_t g = 0; for (int k = 7; k >= 0; k--) g = (g << 8) | d[i + 6 + k]; pg.granule = (int64_t)g;
        pg.serial = (int32_t)((uint32_t)d[i + 14] | ((uint32_t)d[i + 15] << 8) | ((uint32_t)d[i + 16] << 16) | ((uint32_t)d[i + 17] << 24));
        pg.seq = (int32_t)((uint32_t)d[i + 18] | ((uint32_t)d[i + 19] << 8) | ((uint32_t)d[i + 20] << 16) | ((uint32_t)d[i + 21] << 24));
        // lacing (Ogg/PageReader.cs:27-93)
        const uint8_t* body = d + i + 27 + segCnt; size_t dataIdx = 0; int size = 0; pg.isContinued = false;
        for (int s = 0; s < segCnt; s++) {
            int seg = d[i + 27 + s]; size += seg;
            if (seg < 255) {
                if (size > 0) { pg.packets.emplace_back(body + dataIdx, body + dataIdx + size); dataIdx += size; }
                size = 0;
            }
        }
        if (size > 0) { pg.isContinued = d[i + 26 + segCnt] == 255; pg.packets.emplace_back(body + dataIdx, body + dataIdx + size); }
        i += pageLen;
        if (!haveSerial) { haveSerial = true; serial = pg.serial; }
        if (pg.serial != serial) { resync = false; continue; }          // other logical streams: ignored here
        if (pg.packets.empty()) { resync = false; continue; }           // Ogg/PageReader.cs:131
        pg.isResync = resync || (lastSeq != 0 && lastSeq + 1 != pg.seq); // Ogg/StreamPageReader.cs:77-86
        lastSeq = pg.seq; resync = false;
        if (pg.flags & 0x04) hasAll = true;                              // Ogg/StreamPageReader.cs:72-75
        pages.push_back(std::move(pg));
    }
    return pages;
}

static std::vector<std::unique_ptr<Packet>> BuildPackets(const std::vector<Page>& pages) {      // Ogg/PacketProvider.cs:324-438
    std::vector<std::unique_ptr<Packet>> out;
    size_t pageIndex = 0, packetIndex = 0;
    bool lastPageEos = !pages.empty() && (pages.back().flags & 0x04);
    while (pageIndex < pages.size()) {
        const Page& pg = pages[pageIndex];
        int packetCount = (int)pg.packets.size();
        auto pkt = std::make_unique<Packet>();
        pkt->data = pg.packets[packetIndex];
        int64_t granulePos = pg.granule; bool isResync = pg.isResync; bool isContinued = pg.isContinued;
        bool isLastPacket; size_t finalPage = pageIndex;
        if (isContinued && (int)packetIndex == packetCount - 1) {
            size_t contPageIdx = pageIndex; bool failed = false;
            while (isContinued) {
                if (++contPageIdx >= pages.size()) { failed = true; break; }                     // :346-350 -> no packet
                const Page& cp = pages[contPageIdx];
                granulePos = cp.granule; isResync = cp.isResync; isContinued = cp.isContinued; packetCount = (int)cp.packets.size();
                bool isContinuation = (cp.flags & 0x01) != 0;
                if (!isContinuation || isResync) break;                                          // :354-357
                if (isContinued && packetCount > 1) isContinued = false;                         // :360-363
                pkt->data.insert(pkt->data.end(), cp.packets[0].begin(), cp.packets[0].end());
            }
            if (failed) return out;
            isLastPacket = packetCount == 1;
            finalPage = contPageIdx;
        } else {
            isLastPacket = (int)packetIndex == packetCount - 1;
        }
        pkt->isResync = isResync;
        if (isLastPacket) {
            pkt->hasGranule = true; pkt->granule = granulePos;
            if (lastPageEos && finalPage == pages.size() - 1) pkt->isEndOfStream = true;         // :404-407 (HasAllPages <=> EOS page seen)
        }
        if (finalPage != pageIndex) { pageIndex = finalPage; packetIndex = 0; }                  // :411-434
        if ((int)packetIndex == packetCount - 1) { ++pageIndex; packetIndex = 0; } else { ++packetIndex; }
        out.push_back(std::move(pkt));
    }
    return out;
}

// ------------------------------------------------------------------------------------
// Huffman.cs:15-86 + Codebook.cs:59-322
// ------------------------------------------------------------------------------------
struct HuffNode { int Value, Length, Bits, Mask; };

struct Codebook {
    int Dimensions = 0, Entries = 0, MapType = 0;
    std::vector<int> lengths;
    std::vector<float> lookupTable;
    std::vector<int> prefix;            // index into nodes, -1 = null
    std::vector<int> overflow;          // indices into nodes
    std::vector<HuffNode> nodes;
    int prefixBitLength = 0, maxBits = -1;
    bool hasTree = false;

    void Init(Packet& p) {                                                                       // Codebook.cs:59-74
        if (p.ReadBits(24) != 0x564342ULL) throw InvalidData("Book header had invalid signature!");
        Dimensions = (int)p.ReadBits(16);
        Entries = (int)p.ReadBits(24);
        lengths.assign(Entries, 0);
        InitTree(p);
        InitLookupTable(p);
    }

    void InitTree(Packet& p) {                                                                   // Codebook.cs:76-170
        bool sparse; int total = 0; int maxLen;
        if (p.ReadBit()) {
            int len = (int)p.ReadBits(5) + 1;
            for (int i = 0; i < Entries;) {
                int cnt = (int)p.ReadBits(ilog(Entries - i));
                while (--cnt >= 0) { if (i >= Entries) throw InvalidData("ordered codebook overrun"); lengths[i++] = len; }
                ++len;
            }
            total = 0; sparse = false; maxLen = len;
        } else {
            maxLen = -1; sparse = p.ReadBit();
            for (int i = 0; i < Entries; i++) {
                if (!sparse || p.ReadBit()) { lengths[i] = (int)p.ReadBits(5) + 1; ++total; }
                else lengths[i] = -1;
                if (lengths[i] > maxLen) maxLen = lengths[i];
            }
        }
        if ((maxBits = maxLen) > -1) {
            std::vector<int> codewordLengths; bool haveCwl = false;
            if (sparse && total >= (Entries >> 2)) { codewordLengths = lengths; haveCwl = true; sparse = false; }
            int sortedCount = sparse ? total : 0;
            std::vector<int> values, codewords; bool haveValues = false;
            if (!sparse) codewords.assign(Entries, 0);
            else if (sortedCount != 0) { codewordLengths.assign(sortedCount, 0); haveCwl = true; codewords.assign(sortedCount, 0); values.assign(sortedCount, 0); haveValues = true; }
            if (!ComputeCodewords(sparse, codewords, codewordLengths, lengths, Entries, values)) throw InvalidData("bad codeword lengths");
            const std::vector<int>& lenList = haveCwl ? codewordLengths : lengths;
            GenerateTable(haveValues ? &values : nullptr, lenList, codewords);
            hasTree = true;
        }
    }

    static bool ComputeCodewords(bool sparse, std::vector<int>& codewords, std::vector<int>& codewordLengths,
                                 const std::vector<int>& len, int n, std::vector<int>& values) {  // Codebook.cs:172-207
        int i, k, m = 0; uint32_t available[33] = {0};
        for (k = 0; k < n; ++k) if (len[k] > 0) break;
        if (k == n) return true;
        AddEntry(sparse, codewords, codewordLengths, 0, k, m++, len[k], values);
        for (i = 1; i <= len[k]; ++i) available[i] = 1U << (32 - i);
        for (i = k + 1; i < n; ++i) {
            uint32_t res; int z = len[i], y;
            if (z <= 0) continue;
            while (z > 0 && available[z] == 0) --z;
            if (z == 0) return false;
            res = available[z]; available[z] = 0;
            AddEntry(sparse, codewords, codewordLengths, BitReverse(res), i, m++, len[i], values);
            if (z != len[i]) for (y = len[i]; y > z; --y) available[y] = res + (1U << (32 - y));
        }
        return true;
    }
    static void AddEntry(bool sparse, std::vector<int>& codewords, std::vector<int>& codewordLengths, uint32_t huffCode,
                         int symbol, int count, int len, std::vector<int>& values) {              // Codebook.cs:209-221
        if (sparse) { codewords[count] = (int)huffCode; codewordLengths[count] = len; values[count] = symbol; }
        else codewords[symbol] = (int)huffCode;
    }

    void GenerateTable(const std::vector<int>* values, const std::vector<int>& lengthList, const std::vector<int>& codeList) {  // Huffman.cs:15-76
        const int MAX_TABLE_BITS = 10;
        size_t n = lengthList.size();
        nodes.resize(n);
        int maxLen = 0;
        for (size_t i = 0; i < n; i++) {
            nodes[i].Value = values ? (*values)[i] : (int)i;                                     // FastRange.Get(0,n)[i] == i
            nodes[i].Length = lengthList[i] <= 0 ? 99999 : lengthList[i];
            nodes[i].Bits = codeList[i];
            nodes[i].Mask = (int)((1u << (lengthList[i] & 31)) - 1u);
            if (lengthList[i] > 0 && maxLen < lengthList[i]) maxLen = lengthList[i];
        }
        std::vector<int> order(n); for (size_t i = 0; i < n; i++) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {                         // Huffman.cs:78-86
            int len = nodes[a].Length - nodes[b].Length;
            if (len == 0) return (nodes[a].Bits - nodes[b].Bits) < 0;
            return len < 0;
        });
        int tableBits = maxLen > MAX_TABLE_BITS ? MAX_TABLE_BITS : maxLen;
        prefix.assign((size_t)1 << tableBits, -1);
        overflow.clear();
        for (size_t i = 0; i < n && nodes[order[i]].Length < 99999; i++) {
            int itemBits = nodes[order[i]].Length;
            if (itemBits > tableBits) {
                for (; i < n && nodes[order[i]].Length < 99999; i++) overflow.push_back(order[i]);
            } else {
                int maxVal = 1 << (tableBits - itemBits);
                for (int j = 0; j < maxVal; j++) {
                    int idx = (j << itemBits) | nodes[order[i]].Bits;
                    prefix[idx] = order[i];
                }
            }
        }
        prefixBitLength = tableBits;
    }

    int lookup1_values() const {                                                                 // Codebook.cs:285-292
        int r = (int)std::floor(std::exp(std::log((double)Entries) / Dimensions));
        if (std::floor(std::pow((double)(r + 1), (double)Dimensions)) <= Entries) ++r;
        return r;
    }

    void InitLookupTable(Packet& p) {                                                            // Codebook.cs:222-283
        MapType = (int)p.ReadBits(4);
        if (MapType == 0) return;
        float minValue = ConvertFromVorbisFloat32((uint32_t)p.ReadBits(32));
        float deltaValue = ConvertFromVorbisFloat32((uint32_t)p.ReadBits(32));
        int valueBits = (int)p.ReadBits(4) + 1;
        bool sequence_p = p.ReadBit();
        int lookupValueCount = Entries * Dimensions;
        lookupTable.assign((size_t)lookupValueCount, 0.f);
        if (MapType == 1) lookupValueCount = lookup1_values();
        std::vector<uint32_t> multiplicands((size_t)lookupValueCount);
        for (int i = 0; i < lookupValueCount; i++) multiplicands[i] = (uint32_t)p.ReadBits(valueBits);
        if (MapType == 1) {
            for (int idx = 0; idx < Entries; idx++) {
                double last = 0.0; int idxDiv = 1;
                for (int i = 0; i < Dimensions; i++) {
                    int moff = (idx / idxDiv) % lookupValueCount;
                    // float*float + float evaluated in float, then + double (Codebook.cs:253)
                    float ff = (float)multiplicands[moff] * deltaValue;
                    ff = ff + minValue;
                    double value = (double)ff + last;
                    lookupTable[(size_t)idx * Dimensions + i] = (float)value;
                    if (sequence_p) last = value;
                    idxDiv *= lookupValueCount;
                }
            }
        } else {
            for (int idx = 0; idx < Entries; idx++) {
                double last = 0.0; int moff = idx * Dimensions;
                for (int i = 0; i < Dimensions; i++) {
                    float ff = (float)multiplicands[moff] * deltaValue;                          // uint*float -> float (Codebook.cs:273)
                    ff = ff + minValue;
                    double value = (double)ff + last;
                    lookupTable[(size_t)idx * Dimensions + i] = (float)value;
                    if (sequence_p) last = value;
                    ++moff;
                }
            }
        }
    }

    int DecodeScalar(Packet& p) const {                                                          // Codebook.cs:294-320
        int bitsRead;
        int data = (int)p.TryPeekBits(prefixBitLength, bitsRead);
        if (bitsRead == 0) return -1;
        int ni = prefix[(size_t)data & (prefix.size() - 1)];
        if (ni >= 0) { p.SkipBits(nodes[ni].Length); return nodes[ni].Value; }
        int br2;
        data = (int)p.TryPeekBits(maxBits, br2);
        for (size_t i = 0; i < overflow.size(); i++) {
            const HuffNode& node = nodes[overflow[i]];
            if (node.Bits == (data & node.Mask)) { p.SkipBits(node.Length); return node.Value; }
        }
        return -1;
    }
    float at(int entry, int dim) const { return lookupTable[(size_t)entry * Dimensions + dim]; } // Codebook.cs:322
};

// ------------------------------------------------------------------------------------
// Per-frame boundary record: what the CPU half of the split decoder hands to the synthesis
// half (the GPU in the product).  Filled while decoding when Decoder::record is set.
// ------------------------------------------------------------------------------------
struct FrameRec {
    int mode = 0, blockSize = 0, windowIndex = 0;
    int start = 0, valid = 0, total = 0;          // Mode.GetPacketInfo outputs (valid BEFORE the EOS trim)
    int validTrimmed = 0;                          // after StreamDecoder.cs:429-437
    bool ok = false;                               // false: packet failed to decode -> drain
    uint32_t execMask = 0, noExecMask = 0;         // ExecuteChannel after ForceEnergy / noExecuteChannel before
    std::vector<int> postCount;                    // per channel (floor1) ; floor0: 1 if amp>0
    std::vector<std::vector<int>> posts;           // per channel raw Y values (pre-unwrap)
    std::vector<float> f0amp; std::vector<std::vector<float>> f0coeff;   // floor0 payload
    int resStreams = 0, resPartitions = 0;         // classes laid out [stream][partition]
    std::vector<uint8_t> classes;
    std::vector<int> entries;                      // decode order
    bool resDecoded = false;
    std::vector<float> spectrum;                   // [ch][N/2] after floor apply (== IMDCT input), exec channels only
    std::vector<float> block;                      // [ch][N] after window
};

struct Decoder;

// ------------------------------------------------------------------------------------
// Floors
// ------------------------------------------------------------------------------------
struct FloorData {
    // floor1
    int Posts[64]; int PostCount = 0;
    // floor0
    std::vector<float> Coeff; float Amp = 0.f;
    bool isFloor0 = false;
    bool ForceEnergy = false, ForceNoEnergy = false;
    FloorData() { std::memset(Posts, 0, sizeof(Posts)); }
    bool ExecuteChannel() const {                                                                // Floor1.cs:15, Floor0.cs:16
        bool e = isFloor0 ? (Amp > 0.f) : (PostCount > 0);
        return (ForceEnergy || e) && !ForceNoEnergy;
    }
};

static const uint32_t inverse_dB_bits[256] = {
#include "inverse_db_table.inc"
};
static float inverse_dB(int y) { float f; std::memcpy(&f, &inverse_dB_bits[y], 4); return f; }

struct Floor {
    int type = 1;
    // ---- floor1 (Floor1.cs:21-28)
    std::vector<int> partitionClass, classDimensions, classSubclasses, xList, classMasterBookIndex, hNeigh, lNeigh, sortIdx;
    int multiplier = 0, range = 0, yBits = 0;
    std::vector<std::vector<int>> subclassBookIndex;
    // ---- floor0 (Floor0.cs:22-26)
    int order = 0, rate = 0, bark_map_size = 0, ampBits = 0, ampOfs = 0, ampDiv = 0, bookBits = 0;
    std::vector<int> f0books;
    std::vector<int> barkMap[2]; std::vector<float> wMap[2]; int blockSizes[2] = {0, 0};

    const std::vector<Codebook>* books = nullptr;

    void Init1(Packet& p, const std::vector<Codebook>& codebooks) {                              // Floor1.cs:30-133
        books = &codebooks;
        int maximum_class = -1;
        partitionClass.resize((size_t)p.ReadBits(5));
        for (size_t i = 0; i < partitionClass.size(); i++) {
            partitionClass[i] = (int)p.ReadBits(4);
            if (partitionClass[i] > maximum_class) maximum_class = partitionClass[i];
        }
        ++maximum_class;
        classDimensions.assign(maximum_class, 0); classSubclasses.assign(maximum_class, 0);
        classMasterBookIndex.assign(maximum_class, 0); subclassBookIndex.assign(maximum_class, {});
        for (int i = 0; i < maximum_class; i++) {
            classDimensions[i] = (int)p.ReadBits(3) + 1;
            classSubclasses[i] = (int)p.ReadBits(2);
            if (classSubclasses[i] > 0) {
                classMasterBookIndex[i] = (int)p.ReadBits(8);
                if (classMasterBookIndex[i] >= (int)codebooks.size()) throw InvalidData("floor1 master book");
            }
            subclassBookIndex[i].assign((size_t)1 << classSubclasses[i], -1);
            for (size_t j = 0; j < subclassBookIndex[i].size(); j++) {
                int bookNum = (int)p.ReadBits(8) - 1;
                if (bookNum >= (int)codebooks.size()) throw InvalidData("floor1 subclass book");
                subclassBookIndex[i][j] = bookNum;
            }
        }
        multiplier = (int)p.ReadBits(2);
        static const int rangeLookup[4] = {256, 128, 86, 64};                                    // Floor1.cs:27-28
        static const int yBitsLookup[4] = {8, 7, 7, 6};
        range = rangeLookup[multiplier]; yBits = yBitsLookup[multiplier];
        ++multiplier;
        int rangeBits = (int)p.ReadBits(4);
        xList.clear(); xList.push_back(0); xList.push_back(1 << rangeBits);
        for (size_t i = 0; i < partitionClass.size(); i++) {
            int classNum = partitionClass[i];
            for (int j = 0; j < classDimensions[classNum]; j++) xList.push_back((int)p.ReadBits(rangeBits));
        }
        size_t n = xList.size();
        lNeigh.assign(n, 0); hNeigh.assign(n, 0); sortIdx.assign(n, 0);
        sortIdx[0] = 0; sortIdx[1] = 1;
        for (size_t i = 2; i < n; i++) {                                                         // Floor1.cs:98-115
            lNeigh[i] = 0; hNeigh[i] = 1; sortIdx[i] = (int)i;
            for (size_t j = 2; j < i; j++) {
                int temp = xList[j];
                if (temp < xList[i]) { if (temp > xList[lNeigh[i]]) lNeigh[i] = (int)j; }
                else { if (temp < xList[hNeigh[i]]) hNeigh[i] = (int)j; }
            }
        }
        for (size_t i = 0; i + 1 < n; i++) {                                                     // Floor1.cs:118-132
            for (size_t j = i + 1; j < n; j++) {
                if (xList[i] == xList[j]) throw InvalidData("floor1 duplicate x");
                if (xList[sortIdx[i]] > xList[sortIdx[j]]) std::swap(sortIdx[i], sortIdx[j]);
            }
        }
        if (n > 64) throw InvalidData("floor1: more than 64 posts overruns Posts[64] (Floor1.cs:12)");
    }

    static float toBARK(double lsp) {                                                            // Floor0.cs:81-84
        return (float)(13.1 * std::atan(0.00074 * lsp) + 2.24 * std::atan(0.0000000185 * lsp * lsp) + .0001 * lsp);
    }
    std::vector<int> SynthesizeBarkCurve(int n) const {                                          // Floor0.cs:67-79
        float scale = bark_map_size / toBARK(rate / 2);
        std::vector<int> map((size_t)n + 1, 0);
        for (int i = 0; i < n - 1; i++)
            map[i] = std::min(bark_map_size - 1, (int)std::floor(toBARK((double)((rate / 2.f) / n * i)) * scale));
        map[n] = -1;
        return map;
    }
    std::vector<float> SynthesizeWDelMap(int n) const {                                          // Floor0.cs:86-96
        float wdel = (float)(3.14159265358979323846 / bark_map_size);
        std::vector<float> map((size_t)n);
        for (int i = 0; i < n; i++) map[i] = 2.f * (float)std::cos((double)(wdel * i));
        return map;
    }
    void Init0(Packet& p, int block0Size, int block1Size, const std::vector<Codebook>& codebooks) { // Floor0.cs:28-65
        books = &codebooks;
        order = (int)p.ReadBits(8); rate = (int)p.ReadBits(16); bark_map_size = (int)p.ReadBits(16);
        ampBits = (int)p.ReadBits(6); ampOfs = (int)p.ReadBits(8);
        f0books.resize((size_t)p.ReadBits(4) + 1);
        if (order < 1 || rate < 1 || bark_map_size < 1 || f0books.empty()) throw InvalidData("floor0 header");
        ampDiv = (1 << ampBits) - 1;
        for (size_t i = 0; i < f0books.size(); i++) {
            int num = (int)p.ReadBits(8);
            if (num < 0 || num >= (int)codebooks.size()) throw InvalidData("floor0 book");
            if (codebooks[num].MapType == 0 || codebooks[num].Dimensions < 1) throw InvalidData("floor0 book type");
            f0books[i] = num;
        }
        bookBits = ilog((int)f0books.size());
        blockSizes[0] = block0Size; blockSizes[1] = block1Size;
        barkMap[0] = SynthesizeBarkCurve(block0Size / 2); barkMap[1] = SynthesizeBarkCurve(block1Size / 2);
        wMap[0] = SynthesizeWDelMap(block0Size / 2); wMap[1] = SynthesizeWDelMap(block1Size / 2);
    }

    FloorData Unpack(Packet& p) const { return type == 1 ? Unpack1(p) : Unpack0(p); }

    FloorData Unpack1(Packet& p) const {                                                         // Floor1.cs:135-184
        FloorData data;
        if (p.ReadBit()) {
            int postCount = 2;
            data.Posts[0] = (int)p.ReadBits(yBits);
            data.Posts[1] = (int)p.ReadBits(yBits);
            for (int i = 0; i < (int)partitionClass.size(); i++) {
                int clsNum = partitionClass[i];
                int cdim = classDimensions[clsNum];
                int cbits = classSubclasses[clsNum];
                int csub = (1 << cbits) - 1;
                uint32_t cval = 0;
                if (cbits > 0) {
                    if ((cval = (uint32_t)(*books)[classMasterBookIndex[clsNum]].DecodeScalar(p)) == 0xFFFFFFFFu) { postCount = 0; break; }
                }
                for (int j = 0; j < cdim; j++) {
                    int book = subclassBookIndex[clsNum][cval & csub];
                    cval >>= cbits;
                    if (book >= 0) {
                        if ((data.Posts[postCount] = (*books)[book].DecodeScalar(p)) == -1) { postCount = 0; i = (int)partitionClass.size(); break; }
                    }
                    ++postCount;
                }
            }
            data.PostCount = postCount;
        }
        return data;
    }

    FloorData Unpack0(Packet& p) const {                                                         // Floor0.cs:98-150
        FloorData data; data.isFloor0 = true; data.Coeff.assign((size_t)order + 1, 0.f);
        data.Amp = (float)p.ReadBits(ampBits);
        if (data.Amp > 0.f) {
            data.Amp = data.Amp / ampDiv * ampOfs;
            uint32_t bookNum = (uint32_t)p.ReadBits(bookBits);
            if (bookNum >= f0books.size()) { data.Amp = 0; return data; }
            const Codebook& book = (*books)[f0books[bookNum]];
            for (int i = 0; i < order;) {
                int entry = book.DecodeScalar(p);
                if (entry == -1) { data.Amp = 0; return data; }
                for (int j = 0; i < order && j < book.Dimensions; j++, i++) data.Coeff[i] = book.at(entry, j);
            }
            float last = 0.f;
            for (int j = 0; j < order;) {
                for (int k = 0; j < order && k < book.Dimensions; j++, k++) data.Coeff[j] += last;
                last = data.Coeff[j - 1];
            }
        }
        return data;
    }

    void Apply(FloorData& data, int blockSize, float* residue) const { if (type == 1) Apply1(data, blockSize, residue); else Apply0(data, blockSize, residue); }

    // ---- Floor1.Apply  Floor1.cs:186-222
    void Apply1(FloorData& data, int blockSize, float* residue) const {
        int n = blockSize / 2;
        if (data.PostCount > 0) {
            bool stepFlags[64];
            UnwrapPosts(data, stepFlags);
            int lx = 0; int ly = data.Posts[0] * multiplier;
            for (int i = 1; i < data.PostCount; i++) {
                int idx = sortIdx[i];
                if (stepFlags[idx]) {
                    int hx = xList[idx]; int hy = data.Posts[idx] * multiplier;
                    if (lx < n) RenderLineMulti(lx, ly, std::min(hx, n), hy, residue);
                    lx = hx; ly = hy;
                }
                if (lx >= n) break;
            }
            if (lx < n) RenderLineMulti(lx, ly, n, ly, residue);
        } else {
            std::memset(residue, 0, sizeof(float) * (size_t)n);
        }
    }
    void UnwrapPosts(FloorData& data, bool* stepFlags) const {                                   // Floor1.cs:224-297
        std::fill(stepFlags, stepFlags + 64, false);
        stepFlags[0] = true; stepFlags[1] = true;
        int finalY[64] = {0};
        finalY[0] = data.Posts[0]; finalY[1] = data.Posts[1];
        for (int i = 2; i < data.PostCount; i++) {
            int lowOfs = lNeigh[i], highOfs = hNeigh[i];
            int predicted = RenderPoint(xList[lowOfs], finalY[lowOfs], xList[highOfs], finalY[highOfs], xList[i]);
            int val = data.Posts[i];
            int highroom = range - predicted; int lowroom = predicted; int room;
            if (highroom < lowroom) room = highroom * 2; else room = lowroom * 2;
            if (val != 0) {
                stepFlags[lowOfs] = true; stepFlags[highOfs] = true; stepFlags[i] = true;
                if (val >= room) {
                    if (highroom > lowroom) finalY[i] = val - lowroom + predicted;
                    else finalY[i] = predicted - val + highroom - 1;
                } else {
                    if ((val % 2) == 1) finalY[i] = predicted - ((val + 1) / 2);
                    else finalY[i] = predicted + (val / 2);
                }
            } else { stepFlags[i] = false; finalY[i] = predicted; }
        }
        for (int i = 0; i < data.PostCount; i++) data.Posts[i] = finalY[i];
    }
    static int RenderPoint(int x0, int y0, int x1, int y1, int X) {                              // Floor1.cs:299-314
        int dy = y1 - y0; int adx = x1 - x0; int ady = std::abs(dy);
        int err = ady * (X - x0); int off = err / adx;
        return dy < 0 ? y0 - off : y0 + off;
    }
    static void RenderLineMulti(int x0, int y0, int x1, int y1, float* v) {                      // Floor1.cs:316-341
        int dy = y1 - y0; int adx = x1 - x0; int ady = std::abs(dy);
        int sy = 1 - (((dy >> 31) & 1) * 2);
        int b = dy / adx; int x = x0; int y = y0; int err = -adx;
        if (y0 < 0 || y0 > 255) throw InvalidData("floor1 y out of table range");                // C#: IndexOutOfRangeException
        v[x0] *= inverse_dB(y0);
        ady -= std::abs(b) * adx;
        while (++x < x1) {
            y += b; err += ady;
            if (err >= 0) { err -= adx; y += sy; }
            if (y < 0 || y > 255) throw InvalidData("floor1 y out of table range");
            v[x] *= inverse_dB(y);
        }
    }

    // ---- Floor0.Apply  Floor0.cs:152-212
    void Apply0(FloorData& data, int blockSize, float* residue) const {
        int n = blockSize / 2;
        if (data.Amp > 0.f) {
            int bi = (blockSize == blockSizes[0]) ? 0 : 1;
            const std::vector<int>& bm = barkMap[bi]; const std::vector<float>& wm = wMap[bi];
            int i = 0;
            for (i = 0; i < order; i++) data.Coeff[i] = 2.f * (float)std::cos((double)data.Coeff[i]);
            i = 0;
            while (i < n) {
                int j; int k = bm[i]; float p = .5f, q = .5f;
                if (k < 0 || k >= (int)wm.size()) throw InvalidData("floor0 bark index outside wMap");
                float w = wm[k];
                for (j = 1; j < order; j += 2) { q *= w - data.Coeff[j - 1]; p *= w - data.Coeff[j]; }
                if (j == order) { q *= w - data.Coeff[j - 1]; p *= p * (4.f - w * w); q *= q; }
                else { p *= p * (2.f - w); q *= q * (2.f + w); }
                q = data.Amp / (float)std::sqrt((double)(p + q)) - ampOfs;
                q = (float)std::exp((double)(q * 0.11512925f));
                residue[i] *= q;
                while (bm[++i] == k) residue[i] *= q;
            }
        } else std::memset(residue, 0, sizeof(float) * (size_t)n);
    }
};

// ------------------------------------------------------------------------------------
// Residues: Residue0.cs:35-201, Residue1.cs:8-26, Residue2.cs:10-47
// ------------------------------------------------------------------------------------
struct Residue {
    int type = 0;
    int channels = 0;          // Residue0._channels (1 for type 2)
    int r2channels = 0;        // Residue2._channels
    int begin = 0, end = 0, partitionSize = 0, classifications = 0, maxStages = 0;
    int classBook = 0;
    std::vector<int> cascade;
    std::vector<std::vector<int>> books;   // [class][stage] book index or -1
    std::vector<std::vector<int>> decodeMap;
    const std::vector<Codebook>* cb = nullptr;

    static int icount(int v) { int ret = 0; while (v != 0) { ret += (v & 1); v >>= 1; } return ret; }  // Residue0.cs:10-19

    void Init(Packet& p, int chans, const std::vector<Codebook>& codebooks) {                    // Residue0.cs:35-117, Residue2.cs:10-14
        cb = &codebooks;
        if (type == 2) { r2channels = chans; chans = 1; }
        begin = (int)p.ReadBits(24); end = (int)p.ReadBits(24);
        partitionSize = (int)p.ReadBits(24) + 1;
        classifications = (int)p.ReadBits(6) + 1;
        classBook = (int)p.ReadBits(8);
        if (classBook >= (int)codebooks.size()) throw InvalidData("residue class book");
        cascade.assign(classifications, 0);
        int acc = 0;
        for (int i = 0; i < classifications; i++) {
            int low_bits = (int)p.ReadBits(3);
            if (p.ReadBit()) cascade[i] = ((int)p.ReadBits(5) << 3) | low_bits; else cascade[i] = low_bits;
            acc += icount(cascade[i]);
        }
        std::vector<int> bookNums(acc);
        for (int i = 0; i < acc; i++) {
            bookNums[i] = (int)p.ReadBits(8);
            if (bookNums[i] >= (int)codebooks.size()) throw InvalidData("residue book");
            if (codebooks[bookNums[i]].MapType == 0) throw InvalidData("residue book without lookup");
        }
        int entries = codebooks[classBook].Entries; int dim = codebooks[classBook].Dimensions;
        int partvals = 1;
        while (dim > 0) { partvals *= classifications; if (partvals > entries) throw InvalidData("residue partvals"); --dim; }
        books.assign(classifications, {});
        acc = 0; int maxstage = 0;
        for (int j = 0; j < classifications; j++) {
            int stages = ilog(cascade[j]);
            books[j].assign(stages, -1);
            if (stages > 0) {
                maxstage = std::max(maxstage, stages);
                for (int k = 0; k < stages; k++) if ((cascade[j] & (1 << k)) > 0) books[j][k] = bookNums[acc++];
            }
        }
        maxStages = maxstage;
        int cdim = codebooks[classBook].Dimensions;
        decodeMap.assign(partvals, {});
        for (int j = 0; j < partvals; j++) {
            int val = j; int mult = partvals / classifications;
            decodeMap[j].assign(cdim, 0);
            for (int k = 0; k < cdim; k++) {
                int deco = val / mult; val -= deco * mult; mult /= classifications;
                decodeMap[j][k] = deco;
            }
        }
        channels = chans;
    }

    // Residue0.Decode Residue0.cs:119-178 (Residue2.Decode multiplies blockSize, Residue2.cs:16-21)
    void Decode(Packet& p, const std::vector<bool>& doNotDecodeChannel, int blockSize, std::vector<std::vector<float>>& buffer, FrameRec* rec) const {
        if (type == 2) blockSize *= r2channels;
        const Codebook& classBk = (*cb)[classBook];
        int e = end < blockSize / 2 ? end : blockSize / 2;
        int n = e - begin;
        bool anyLive = std::find(doNotDecodeChannel.begin(), doNotDecodeChannel.end(), false) != doNotDecodeChannel.end();
        if (n > 0 && anyLive) {
            int partitionCount = n / partitionSize;
            int cdim = classBk.Dimensions;
            int partitionWords = (partitionCount + cdim - 1) / cdim;
            std::vector<const std::vector<int>*> partWordCache((size_t)channels * partitionWords, nullptr);
            if (rec) { rec->resDecoded = true; rec->resStreams = channels; rec->resPartitions = partitionCount; rec->classes.assign((size_t)channels * partitionCount, 0); }
            int stageLimit = maxStages;
            for (int stage = 0; stage < stageLimit; stage++) {
                for (int partitionIdx = 0, entryIdx = 0; partitionIdx < partitionCount; entryIdx++) {
                    if (stage == 0) {
                        for (int ch = 0; ch < channels; ch++) {
                            int idx = classBk.DecodeScalar(p);
                            if (idx >= 0 && idx < (int)decodeMap.size()) {
                                partWordCache[(size_t)ch * partitionWords + entryIdx] = &decodeMap[idx];
                                if (rec) for (int d = 0; d < cdim && entryIdx * cdim + d < partitionCount; d++)
                                    rec->classes[(size_t)ch * partitionCount + entryIdx * cdim + d] = (uint8_t)decodeMap[idx][d];
                            } else { partitionIdx = partitionCount; stage = stageLimit; break; }
                        }
                    }
                    for (int dimensionIdx = 0; partitionIdx < partitionCount && dimensionIdx < cdim; dimensionIdx++, partitionIdx++) {
                        int offset = begin + partitionIdx * partitionSize;
                        for (int ch = 0; ch < channels; ch++) {
                            int idx = (*partWordCache[(size_t)ch * partitionWords + entryIdx])[dimensionIdx];
                            if ((cascade[idx] & (1 << stage)) != 0) {
                                int book = books[idx][stage];
                                if (book >= 0) {
                                    if (WriteVectors((*cb)[book], p, buffer, ch, offset, partitionSize, rec)) {
                                        partitionIdx = partitionCount; stage = stageLimit; break;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
    }

    bool WriteVectors(const Codebook& codebook, Packet& p, std::vector<std::vector<float>>& residue, int channel, int offset, int psize, FrameRec* rec) const {
        if (type == 0) {                                                                         // Residue0.cs:180-201
            float* res = residue[channel].data();
            int steps = psize / codebook.Dimensions;
            std::vector<int> entryCache(steps);
            for (int i = 0; i < steps; i++) if ((entryCache[i] = codebook.DecodeScalar(p)) == -1) return true;
            if (rec) rec->entries.insert(rec->entries.end(), entryCache.begin(), entryCache.end());
            for (int dim = 0; dim < codebook.Dimensions; dim++)
                for (int step = 0; step < steps; step++, offset++) res[offset] += codebook.at(entryCache[step], dim);
            return false;
        } else if (type == 1) {                                                                  // Residue1.cs:8-26
            float* res = residue[channel].data();
            for (int i = 0; i < psize;) {
                int entry = codebook.DecodeScalar(p);
                if (entry == -1) return true;
                if (rec) rec->entries.push_back(entry);
                for (int j = 0; j < codebook.Dimensions; i++, j++) res[offset + i] += codebook.at(entry, j);
            }
            return false;
        } else {                                                                                 // Residue2.cs:23-47
            int chPtr = 0;
            offset /= r2channels;
            for (int c = 0; c < psize;) {
                int entry = codebook.DecodeScalar(p);
                if (entry == -1) return true;
                if (rec) rec->entries.push_back(entry);
                for (int d = 0; d < codebook.Dimensions; d++, c++) {
                    residue[chPtr][offset] += codebook.at(entry, d);
                    if (++chPtr == r2channels) { chPtr = 0; offset++; }
                }
            }
            return false;
        }
    }

    // Synthesis half only: replays the adds of Decode/WriteVectors from recorded (classes, entries)
    // in exactly the order the reference performs them.  entryLimit entries are available.
    void Replay(const uint8_t* classes, int nStreams, int nPartitions, const int* entries, int entryCount,
                int blockSize, std::vector<std::vector<float>>& buffer) const {
        if (type == 2) blockSize *= r2channels;
        int e = end < blockSize / 2 ? end : blockSize / 2;
        int n = e - begin; if (n <= 0) return;
        int partitionCount = n / partitionSize;
        if (partitionCount != nPartitions || nStreams != channels) throw InvalidData("replay geometry mismatch");
        int pos = 0;
        for (int stage = 0; stage < maxStages; stage++) {
            for (int partitionIdx = 0; partitionIdx < partitionCount; partitionIdx++) {
                int offset = begin + partitionIdx * partitionSize;
                for (int ch = 0; ch < channels; ch++) {
                    int idx = classes[(size_t)ch * partitionCount + partitionIdx];
                    if ((cascade[idx] & (1 << stage)) == 0) continue;
                    int book = books[idx][stage]; if (book < 0) continue;
                    const Codebook& codebook = (*cb)[book];
                    if (type == 0) {
                        int steps = partitionSize / codebook.Dimensions;
                        if (pos + steps > entryCount) return;
                        float* res = buffer[ch].data(); int o = offset;
                        for (int dim = 0; dim < codebook.Dimensions; dim++)
                            for (int step = 0; step < steps; step++, o++) res[o] += codebook.at(entries[pos + step], dim);
                        pos += steps;
                    } else if (type == 1) {
                        float* res = buffer[ch].data();
                        for (int i = 0; i < partitionSize;) {
                            if (pos >= entryCount) return;
                            int entry = entries[pos++];
                            for (int j = 0; j < codebook.Dimensions; i++, j++) res[offset + i] += codebook.at(entry, j);
                        }
                    } else {
                        int chPtr = 0; int o = offset / r2channels;
                        for (int c = 0; c < partitionSize;) {
                            if (pos >= entryCount) return;
                            int entry = entries[pos++];
                            for (int d = 0; d < codebook.Dimensions; d++, c++) {
                                buffer[chPtr][o] += codebook.at(entry, d);
                                if (++chPtr == r2channels) { chPtr = 0; o++; }
                            }
                        }
                    }
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------
// Mdct.cs:23-535
// ------------------------------------------------------------------------------------
struct MdctImpl {
    int _n, _n2, _n4, _n8, _ld;
    std::vector<float> _a, _b, _c; std::vector<uint16_t> _bitrev;
    std::vector<float> buf2;

    explicit MdctImpl(int n) {                                                                   // Mdct.cs:30-63
        const float M_PI_F = 3.14159265358979323846264f;                                         // Mdct.cs:9
        _n = n; _n2 = n >> 1; _n4 = _n2 >> 1; _n8 = _n4 >> 1;
        _ld = ilog(n) - 1;
        _a.assign(_n2, 0.f); _b.assign(_n2, 0.f); _c.assign(_n4, 0.f);
        int k, k2;
        for (k = k2 = 0; k < _n4; ++k, k2 += 2) {
            // C#: int*float/int evaluated in float32, then widened for Math.Cos/Sin
            float ang_a = (float)(4 * k) * M_PI_F / (float)n;
            _a[k2] = (float)std::cos((double)ang_a);
            _a[k2 + 1] = (float)-std::sin((double)ang_a);
            float ang_b = (float)(k2 + 1) * M_PI_F / (float)n / 2.f;
            _b[k2] = (float)std::cos((double)ang_b) * .5f;
            _b[k2 + 1] = (float)std::sin((double)ang_b) * .5f;
        }
        for (k = k2 = 0; k < _n8; ++k, k2 += 2) {
            float ang_c = (float)(2 * (k2 + 1)) * M_PI_F / (float)n;
            _c[k2] = (float)std::cos((double)ang_c);
            _c[k2 + 1] = (float)-std::sin((double)ang_c);
        }
        _bitrev.assign(_n8, 0);
        for (int i = 0; i < _n8; ++i) _bitrev[i] = (uint16_t)(BitReverse((uint32_t)i, _ld - 3) << 2);
        buf2.assign(_n2, 0.f);
    }

    void CalcReverse(float* buffer) {                                                            // Mdct.cs:65-313
        float* u; float* v;
        std::fill(buf2.begin(), buf2.end(), 0.f);                                                // new float[_n2] (:69)
        float* b2 = buf2.data();
        const float* A = _a.data(); const float* B = _b.data(); const float* C = _c.data();
        {   // step 0 (:74-97)
            int d = _n2 - 2, AA = 0, e = 0, e_stop = _n2;
            while (e != e_stop) {
                b2[d + 1] = (buffer[e] * A[AA] - buffer[e + 2] * A[AA + 1]);
                b2[d] = (buffer[e] * A[AA + 1] + buffer[e + 2] * A[AA]);
                d -= 2; AA += 2; e += 4;
            }
            e = _n2 - 3;
            while (d >= 0) {
                b2[d + 1] = (-buffer[e + 2] * A[AA] - -buffer[e] * A[AA + 1]);
                b2[d] = (-buffer[e + 2] * A[AA + 1] + -buffer[e] * A[AA]);
                d -= 2; AA += 2; e -= 4;
            }
        }
        u = buffer; v = b2;
        {   // step 2 (:105-139)
            int AA = _n2 - 8, e0 = _n4, e1 = 0, d0 = _n4, d1 = 0;
            while (AA >= 0) {
                float v40_20, v41_21;
                v41_21 = v[e0 + 1] - v[e1 + 1]; v40_20 = v[e0] - v[e1];
                u[d0 + 1] = v[e0 + 1] + v[e1 + 1]; u[d0] = v[e0] + v[e1];
                u[d1 + 1] = v41_21 * A[AA + 4] - v40_20 * A[AA + 5];
                u[d1] = v40_20 * A[AA + 4] + v41_21 * A[AA + 5];
                v41_21 = v[e0 + 3] - v[e1 + 3]; v40_20 = v[e0 + 2] - v[e1 + 2];
                u[d0 + 3] = v[e0 + 3] + v[e1 + 3]; u[d0 + 2] = v[e0 + 2] + v[e1 + 2];
                u[d1 + 3] = v41_21 * A[AA] - v40_20 * A[AA + 1];
                u[d1 + 2] = v40_20 * A[AA] + v41_21 * A[AA + 1];
                AA -= 8; d0 += 4; d1 += 4; e0 += 4; e1 += 4;
            }
        }
        // step 3 (:141-186)
        step3_iter0_loop(_n >> 4, u, _n2 - 1 - _n4 * 0, -_n8);
        step3_iter0_loop(_n >> 4, u, _n2 - 1 - _n4 * 1, -_n8);
        step3_inner_r_loop(_n >> 5, u, _n2 - 1 - _n8 * 0, -(_n >> 4), 16);
        step3_inner_r_loop(_n >> 5, u, _n2 - 1 - _n8 * 1, -(_n >> 4), 16);
        step3_inner_r_loop(_n >> 5, u, _n2 - 1 - _n8 * 2, -(_n >> 4), 16);
        step3_inner_r_loop(_n >> 5, u, _n2 - 1 - _n8 * 3, -(_n >> 4), 16);
        int l = 2;
        for (; l < (_ld - 3) >> 1; ++l) {
            int k0 = _n >> (l + 2), k0_2 = k0 >> 1, lim = 1 << (l + 1);
            for (int i = 0; i < lim; ++i) step3_inner_r_loop(_n >> (l + 4), u, _n2 - 1 - k0 * i, -k0_2, 1 << (l + 3));
        }
        for (; l < _ld - 6; ++l) {
            int k0 = _n >> (l + 2), k1 = 1 << (l + 3), k0_2 = k0 >> 1;
            int rlim = _n >> (l + 6), lim = 1 << (l + 1);          // C#: 1 << l + 1 == 1 << (l+1)
            int i_off = _n2 - 1, A0 = 0;
            for (int r = rlim; r > 0; --r) { step3_inner_s_loop(lim, u, i_off, -k0_2, A0, k1, k0); A0 += k1 * 4; i_off -= 8; }
        }
        step3_inner_s_loop_ld654(_n >> 5, u, _n2 - 1, _n);
        {   // steps 4,5,6 (:189-214)
            int bit = 0, d0 = _n4 - 4, d1 = _n2 - 4;
            while (d0 >= 0) {
                int k4 = _bitrev[bit];
                v[d1 + 3] = u[k4]; v[d1 + 2] = u[k4 + 1]; v[d0 + 3] = u[k4 + 2]; v[d0 + 2] = u[k4 + 3];
                k4 = _bitrev[bit + 1];
                v[d1 + 1] = u[k4]; v[d1] = u[k4 + 1]; v[d0 + 1] = u[k4 + 2]; v[d0] = u[k4 + 3];
                d0 -= 4; d1 -= 4; bit += 2;
            }
        }
        {   // step 7 (:217-258)
            int c = 0, d = 0, e = _n2 - 4;
            while (d < e) {
                float a02, a11, b0, b1, b2_, b3;
                a02 = v[d] - v[e + 2]; a11 = v[d + 1] + v[e + 3];
                b0 = C[c + 1] * a02 + C[c] * a11; b1 = C[c + 1] * a11 - C[c] * a02;
                b2_ = v[d] + v[e + 2]; b3 = v[d + 1] - v[e + 3];
                v[d] = b2_ + b0; v[d + 1] = b3 + b1; v[e + 2] = b2_ - b0; v[e + 3] = b1 - b3;
                a02 = v[d + 2] - v[e]; a11 = v[d + 3] + v[e + 1];
                b0 = C[c + 3] * a02 + C[c + 2] * a11; b1 = C[c + 3] * a11 - C[c + 2] * a02;
                b2_ = v[d + 2] + v[e]; b3 = v[d + 3] - v[e + 1];
                v[d + 2] = b2_ + b0; v[d + 3] = b3 + b1; v[e] = b2_ - b0; v[e + 1] = b1 - b3;
                c += 4; d += 4; e -= 4;
            }
        }
        {   // step 8 + decode (:261-312)
            int b = _n2 - 8, e = _n2 - 8, d0 = 0, d1 = _n2 - 4, d2 = _n2, d3 = _n - 4;
            while (e >= 0) {
                float p0, p1, p2, p3;
                p3 = b2[e + 6] * B[b + 7] - b2[e + 7] * B[b + 6];
                p2 = -b2[e + 6] * B[b + 6] - b2[e + 7] * B[b + 7];
                buffer[d0] = p3; buffer[d1 + 3] = -p3; buffer[d2] = p2; buffer[d3 + 3] = p2;
                p1 = b2[e + 4] * B[b + 5] - b2[e + 5] * B[b + 4];
                p0 = -b2[e + 4] * B[b + 4] - b2[e + 5] * B[b + 5];
                buffer[d0 + 1] = p1; buffer[d1 + 2] = -p1; buffer[d2 + 1] = p0; buffer[d3 + 2] = p0;
                p3 = b2[e + 2] * B[b + 3] - b2[e + 3] * B[b + 2];
                p2 = -b2[e + 2] * B[b + 2] - b2[e + 3] * B[b + 3];
                buffer[d0 + 2] = p3; buffer[d1 + 1] = -p3; buffer[d2 + 2] = p2; buffer[d3 + 1] = p2;
                p1 = b2[e] * B[b + 1] - b2[e + 1] * B[b];
                p0 = -b2[e] * B[b] - b2[e + 1] * B[b + 1];
                buffer[d0 + 3] = p1; buffer[d1] = -p1; buffer[d2 + 3] = p0; buffer[d3] = p0;
                b -= 8; e -= 8; d0 += 4; d2 += 4; d1 -= 4; d3 -= 4;
            }
        }
    }

    // One radix-2 butterfly pair with twiddle (a0,a1), shared by the three step-3 loop shapes.
    static inline void bfly(float* e, int i0, int i2, float a0, float a1) {
        float k00 = e[i0] - e[i2];
        float k01 = e[i0 - 1] - e[i2 - 1];
        e[i0] += e[i2];
        e[i0 - 1] += e[i2 - 1];
        e[i2] = k00 * a0 - k01 * a1;
        e[i2 - 1] = k01 * a0 + k00 * a1;
    }
    void step3_iter0_loop(int n, float* e, int i_off, int k_off) {                               // Mdct.cs:315-359
        int ee0 = i_off, ee2 = ee0 + k_off, a = 0;
        for (int i = n >> 2; i > 0; --i) {
            bfly(e, ee0, ee2, _a[a], _a[a + 1]); a += 8;
            bfly(e, ee0 - 2, ee2 - 2, _a[a], _a[a + 1]); a += 8;
            bfly(e, ee0 - 4, ee2 - 4, _a[a], _a[a + 1]); a += 8;
            bfly(e, ee0 - 6, ee2 - 6, _a[a], _a[a + 1]); a += 8;
            ee0 -= 8; ee2 -= 8;
        }
    }
    void step3_inner_r_loop(int lim, float* e, int d0, int k_off, int k1) {                      // Mdct.cs:361-410
        int e0 = d0, e2 = e0 + k_off, a = 0;
        for (int i = lim >> 2; i > 0; --i) {
            bfly(e, e0, e2, _a[a], _a[a + 1]); a += k1;
            bfly(e, e0 - 2, e2 - 2, _a[a], _a[a + 1]); a += k1;
            bfly(e, e0 - 4, e2 - 4, _a[a], _a[a + 1]); a += k1;
            bfly(e, e0 - 6, e2 - 6, _a[a], _a[a + 1]); a += k1;
            e0 -= 8; e2 -= 8;
        }
    }
    void step3_inner_s_loop(int n, float* e, int i_off, int k_off, int a, int a_off, int k0) {   // Mdct.cs:412-461
        float A0 = _a[a], A1 = _a[a + 1], A2 = _a[a + a_off], A3 = _a[a + a_off + 1];
        float A4 = _a[a + a_off * 2], A5 = _a[a + a_off * 2 + 1], A6 = _a[a + a_off * 3], A7 = _a[a + a_off * 3 + 1];
        int ee0 = i_off, ee2 = ee0 + k_off;
        for (int i = n; i > 0; --i) {
            bfly(e, ee0, ee2, A0, A1);
            bfly(e, ee0 - 2, ee2 - 2, A2, A3);
            bfly(e, ee0 - 4, ee2 - 4, A4, A5);
            bfly(e, ee0 - 6, ee2 - 6, A6, A7);
            ee0 -= k0; ee2 -= k0;
        }
    }
    void step3_inner_s_loop_ld654(int n, float* e, int i_off, int base_n) {                      // Mdct.cs:463-507
        int a_off = base_n >> 3; float A2 = _a[a_off];
        int z = i_off; int base = z - 16 * n;
        while (z > base) {
            float k00, k11;
            k00 = e[z] - e[z - 8]; k11 = e[z - 1] - e[z - 9];
            e[z] += e[z - 8]; e[z - 1] += e[z - 9]; e[z - 8] = k00; e[z - 9] = k11;
            k00 = e[z - 2] - e[z - 10]; k11 = e[z - 3] - e[z - 11];
            e[z - 2] += e[z - 10]; e[z - 3] += e[z - 11];
            e[z - 10] = (k00 + k11) * A2; e[z - 11] = (k11 - k00) * A2;
            k00 = e[z - 12] - e[z - 4]; k11 = e[z - 5] - e[z - 13];
            e[z - 4] += e[z - 12]; e[z - 5] += e[z - 13]; e[z - 12] = k11; e[z - 13] = k00;
            k00 = e[z - 14] - e[z - 6]; k11 = e[z - 7] - e[z - 15];
            e[z - 6] += e[z - 14]; e[z - 7] += e[z - 15];
            e[z - 14] = (k00 + k11) * A2; e[z - 15] = (k00 - k11) * A2;
            iter_54(e, z); iter_54(e, z - 8);
            z -= 16;
        }
    }
    static void iter_54(float* e, int z) {                                                       // Mdct.cs:509-535
        float k00, k11, k22, k33, y0, y1, y2, y3;
        k00 = e[z] - e[z - 4]; y0 = e[z] + e[z - 4]; y2 = e[z - 2] + e[z - 6]; k22 = e[z - 2] - e[z - 6];
        e[z] = y0 + y2; e[z - 2] = y0 - y2;
        k33 = e[z - 3] - e[z - 7];
        e[z - 4] = k00 + k33; e[z - 6] = k00 - k33;
        k11 = e[z - 1] - e[z - 5]; y1 = e[z - 1] + e[z - 5]; y3 = e[z - 3] + e[z - 7];
        e[z - 1] = y1 + y3; e[z - 3] = y1 - y3; e[z - 5] = k11 - k22; e[z - 7] = k11 + k22;
    }
};

struct Mdct {                                                                                    // Mdct.cs:11-21
    std::vector<std::unique_ptr<MdctImpl>> cache;
    MdctImpl& get(int n) {
        for (auto& m : cache) if (m->_n == n) return *m;
        cache.push_back(std::make_unique<MdctImpl>(n));
        return *cache.back();
    }
    void Reverse(float* samples, int sampleCount) { get(sampleCount).CalcReverse(samples); }
};

// ------------------------------------------------------------------------------------
// Mapping.cs:16-198
// ------------------------------------------------------------------------------------
struct Mapping {
    std::vector<int> couplingAngle, couplingMagnitude;
    std::vector<int> submapFloor, submapResidue;   // indices
    std::vector<int> channelFloor, channelResidue; // indices
    std::vector<int> mux;

    void Init(Packet& p, int channels, int nFloors, int nResidues) {                             // Mapping.cs:16-93
        int submapCount = 1;
        if (p.ReadBit()) submapCount += (int)p.ReadBits(4);
        int couplingSteps = 0;
        if (p.ReadBit()) couplingSteps = (int)p.ReadBits(8) + 1;
        int couplingBits = ilog(channels - 1);
        couplingAngle.assign(couplingSteps, 0); couplingMagnitude.assign(couplingSteps, 0);
        for (int j = 0; j < couplingSteps; j++) {
            int magnitude = (int)p.ReadBits(couplingBits); int angle = (int)p.ReadBits(couplingBits);
            if (magnitude == angle || magnitude > channels - 1 || angle > channels - 1) throw InvalidData("Invalid magnitude or angle in mapping header!");
            couplingAngle[j] = angle; couplingMagnitude[j] = magnitude;
        }
        if (0 != p.ReadBits(2)) throw InvalidData("Reserved bits not 0 in mapping header.");
        mux.assign(channels, 0);
        if (submapCount > 1) for (int c = 0; c < channels; c++) {
            mux[c] = (int)p.ReadBits(4);
            if (mux[c] > submapCount) throw InvalidData("Invalid channel mux submap index in mapping header!");
            if (mux[c] == submapCount) throw InvalidData("mux == submapCount: IndexOutOfRange in the reference (Mapping.cs:88)");
        }
        submapFloor.assign(submapCount, 0); submapResidue.assign(submapCount, 0);
        for (int j = 0; j < submapCount; j++) {
            p.SkipBits(8);
            int floorNum = (int)p.ReadBits(8); if (floorNum >= nFloors) throw InvalidData("Invalid floor number in mapping header!");
            int residueNum = (int)p.ReadBits(8); if (residueNum >= nResidues) throw InvalidData("Invalid residue number in mapping header!");
            submapFloor[j] = floorNum; submapResidue[j] = residueNum;
        }
        channelFloor.assign(channels, 0); channelResidue.assign(channels, 0);
        for (int c = 0; c < channels; c++) { channelFloor[c] = submapFloor[mux[c]]; channelResidue[c] = submapResidue[mux[c]]; }
    }
};

static inline void InverseCouple(float* magnitude, float* angle, int halfBlockSize) {            // Mapping.cs:145-181
    for (int j = 0; j < halfBlockSize; j++) {
        float newM, newA; float oldM = magnitude[j], oldA = angle[j];
        if (oldM > 0) {
            if (oldA > 0) { newM = oldM; newA = oldM - oldA; } else { newA = oldM; newM = oldM + oldA; }
        } else {
            if (oldA > 0) { newM = oldM; newA = oldM + oldA; } else { newA = oldM; newM = oldM - oldA; }
        }
        magnitude[j] = newM; angle[j] = newA;
    }
}

// ------------------------------------------------------------------------------------
// Mode.cs:24-170
// ------------------------------------------------------------------------------------
struct OverlapInfo { int start, total, valid; };

static std::vector<float> CalcWindow(int prevBlockSize, int blockSize, int nextBlockSize) {      // Mode.cs:69-100
    const float M_PI2 = 3.1415926539f / 2;                                                       // Mode.cs:15
    std::vector<float> array((size_t)blockSize, 0.f);
    int left = prevBlockSize / 2, wnd = blockSize, right = nextBlockSize / 2;
    int leftbegin = wnd / 4 - left / 2;
    int rightbegin = wnd - wnd / 4 - right / 2;
    for (int i = 0; i < left; i++) {
        float x = (float)std::sin((i + .5) / left * (double)M_PI2);
        x *= x;
        array[leftbegin + i] = (float)std::sin((double)(x * M_PI2));
    }
    for (int i = leftbegin + left; i < rightbegin; i++) array[i] = 1.0f;
    for (int i = 0; i < right; i++) {
        float x = (float)std::sin((right - i - .5) / right * (double)M_PI2);
        x *= x;
        array[rightbegin + i] = (float)std::sin((double)(x * M_PI2));
    }
    return array;
}
static OverlapInfo CalcOverlap(int prevBlockSize, int blockSize, int nextBlockSize) {            // Mode.cs:102-117
    int leftOverlapHalfSize = prevBlockSize / 4, rightOverlapHalfSize = nextBlockSize / 4;
    OverlapInfo o;
    o.start = blockSize / 4 - leftOverlapHalfSize;
    o.total = blockSize / 4 * 3 + rightOverlapHalfSize;
    o.valid = o.total - rightOverlapHalfSize * 2;
    return o;
}

struct Mode {
    bool blockFlag = false; int blockSize = 0; int mapping = 0;
    std::vector<std::vector<float>> windows; OverlapInfo overlap[4];
    void Init(Packet& p, int block0Size, int block1Size, int nMappings) {                        // Mode.cs:24-67
        blockFlag = p.ReadBit();
        if (0 != p.ReadBits(32)) throw InvalidData("Mode header had invalid window or transform type!");
        mapping = (int)p.ReadBits(8);
        if (mapping >= nMappings) throw InvalidData("Mode header had invalid mapping index!");
        if (blockFlag) {
            blockSize = block1Size;
            windows = { CalcWindow(block0Size, block1Size, block0Size), CalcWindow(block1Size, block1Size, block0Size),
                        CalcWindow(block0Size, block1Size, block1Size), CalcWindow(block1Size, block1Size, block1Size) };
            overlap[0] = CalcOverlap(block0Size, block1Size, block0Size); overlap[1] = CalcOverlap(block1Size, block1Size, block0Size);
            overlap[2] = CalcOverlap(block0Size, block1Size, block1Size); overlap[3] = CalcOverlap(block1Size, block1Size, block1Size);
        } else {
            blockSize = block0Size;
            windows = { CalcWindow(block0Size, block0Size, block0Size) };
        }
    }
    bool GetPacketInfo(Packet& p, int& windowIndex, int& start, int& valid, int& total) const {  // Mode.cs:119-151
        if (p.isShort) { windowIndex = 0; start = valid = total = 0; return false; }
        if (blockFlag) {
            bool prevFlag = p.ReadBit(); bool nextFlag = p.ReadBit();
            windowIndex = (prevFlag ? 1 : 0) + (nextFlag ? 2 : 0);
            start = overlap[windowIndex].start; valid = overlap[windowIndex].valid; total = overlap[windowIndex].total;
        } else { windowIndex = 0; start = 0; valid = blockSize / 2; total = blockSize; }
        return true;
    }
};

// ------------------------------------------------------------------------------------
// StreamDecoder.cs:107-541 (header load, Read loop, OLA, clip, interleave)
// ------------------------------------------------------------------------------------
struct Decoder {
    std::vector<std::unique_ptr<Packet>> packets; size_t nextPacket = 0;
    int channels = 0, sampleRate = 0, block0Size = 0, block1Size = 0, modeFieldBits = 0;
    std::vector<Codebook> books; std::vector<Floor> floors; std::vector<Residue> residues;
    std::vector<Mapping> mappings; std::vector<Mode> modes;
    Mdct mdct;
    // decode state
    int64_t currentPosition = 0; bool hasClipped = false, hasPosition = false, eosFound = false; bool clipSamples = true;
    std::vector<std::vector<float>> bufA, bufB; std::vector<std::vector<float>>* nextBuf = nullptr; std::vector<std::vector<float>>* prevBuf = nullptr;
    int prevPacketStart = 0, prevPacketEnd = 0, prevPacketStop = 0;
    // recording
    bool record = false; bool recordDense = false; std::vector<FrameRec> recs;
    size_t setupPacketIndex = 0;

    Packet* GetNextPacket() { return nextPacket < packets.size() ? packets[nextPacket++].get() : nullptr; }

    static bool ValidateHeader(Packet& p, const uint8_t* expected, int n) {                      // StreamDecoder.cs:149-159
        for (int i = 0; i < n; i++) if (expected[i] != p.ReadBits(8)) return false;
        return true;
    }
    bool LoadStreamHeader(Packet& p) {                                                           // StreamDecoder.cs:179-204
        static const uint8_t sig[] = {0x01, 0x76, 0x6f, 0x72, 0x62, 0x69, 0x73, 0, 0, 0, 0};
        if (!ValidateHeader(p, sig, 11)) return false;
        channels = (int)(uint8_t)p.ReadBits(8);
        sampleRate = (int)p.ReadBits(32);
        p.ReadBits(32); p.ReadBits(32); p.ReadBits(32);
        block0Size = 1 << (int)p.ReadBits(4);
        block1Size = 1 << (int)p.ReadBits(4);
        return true;
    }
    bool LoadComments(Packet& p) {                                                               // StreamDecoder.cs:206-224 (content unused)
        static const uint8_t sig[] = {0x03, 0x76, 0x6f, 0x72, 0x62, 0x69, 0x73};
        return ValidateHeader(p, sig, 7);
    }
    bool LoadBooks(Packet& p) {                                                                  // StreamDecoder.cs:226-289
        static const uint8_t sig[] = {0x05, 0x76, 0x6f, 0x72, 0x62, 0x69, 0x73};
        if (!ValidateHeader(p, sig, 7)) return false;
        books.resize((size_t)p.ReadBits(8) + 1);
        for (auto& b : books) b.Init(p);
        int times = (int)p.ReadBits(6) + 1;
        p.SkipBits(16 * times);
        floors.resize((size_t)p.ReadBits(6) + 1);
        for (auto& f : floors) {
            int type = (int)p.ReadBits(16);                                                      // Factory.cs:22-31
            if (type == 0) { f.type = 0; f.Init0(p, block0Size, block1Size, books); }
            else if (type == 1) { f.type = 1; f.Init1(p, books); }
            else throw InvalidData("Invalid floor type!");
        }
        residues.resize((size_t)p.ReadBits(6) + 1);
        for (auto& r : residues) {
            int type = (int)p.ReadBits(16);                                                      // Factory.cs:48-58
            if (type < 0 || type > 2) throw InvalidData("Invalid residue type!");
            r.type = type; r.Init(p, channels, books);
        }
        mappings.resize((size_t)p.ReadBits(6) + 1);
        for (auto& m : mappings) {
            if (p.ReadBits(16) != 0) throw InvalidData("Invalid mapping type!");                 // Factory.cs:33-41
            m.Init(p, channels, (int)floors.size(), (int)residues.size());
        }
        modes.resize((size_t)p.ReadBits(6) + 1);
        for (auto& m : modes) m.Init(p, block0Size, block1Size, (int)mappings.size());
        if (!p.ReadBit()) throw InvalidData("Book packet did not end on correct bit!");
        modeFieldBits = ilog((int)modes.size() - 1);
        return true;
    }
    void ResetDecoder() {                                                                        // StreamDecoder.cs:295-305
        prevBuf = nullptr; prevPacketStart = prevPacketEnd = prevPacketStop = 0; nextBuf = nullptr;
        eosFound = false; hasClipped = false; hasPosition = false;
    }
    void Open(const uint8_t* d, size_t len) {                                                    // StreamDecoder.cs:50-68,107-127
        auto pages = ReadPages(d, len);
        packets = BuildPackets(pages);
        Packet* p = GetNextPacket();
        if (!p || !LoadStreamHeader(*p)) throw InvalidData("Could not find Vorbis data to decode.");
        p = GetNextPacket(); if (!p || !LoadComments(*p)) throw InvalidData("bad comment header");
        setupPacketIndex = nextPacket;
        p = GetNextPacket(); if (!p || !LoadBooks(*p)) throw InvalidData("bad setup header");
        currentPosition = 0; ResetDecoder();
    }

    // Same as Open, but from an already-demuxed packet list (an IPacketProvider in the reference's
    // terms, Contracts/IPacketProvider.cs): used with the committed tests/golden/*.packets.npz fixtures.
    void OpenPackets(std::vector<std::unique_ptr<Packet>>&& list) {
        packets = std::move(list);
        Packet* p = GetNextPacket();
        if (!p || !LoadStreamHeader(*p)) throw InvalidData("Could not find Vorbis data to decode.");
        p = GetNextPacket(); if (!p || !LoadComments(*p)) throw InvalidData("bad comment header");
        setupPacketIndex = nextPacket;
        p = GetNextPacket(); if (!p || !LoadBooks(*p)) throw InvalidData("bad setup header");
        currentPosition = 0; ResetDecoder();
    }

    // Mapping.DecodePacket Mapping.cs:95-198
    void MappingDecodePacket(const Mapping& map, Packet& p, int blockSize, std::vector<std::vector<float>>& buffer, FrameRec* rec) {
        int halfBlockSize = blockSize >> 1;
        int nch = (int)map.channelFloor.size();
        std::vector<FloorData> floorData((size_t)nch);
        std::vector<bool> noExecuteChannel((size_t)nch);
        for (int i = 0; i < nch; i++) {
            floorData[i] = floors[map.channelFloor[i]].Unpack(p);
            noExecuteChannel[i] = !floorData[i].ExecuteChannel();
            std::fill(buffer[i].begin(), buffer[i].begin() + halfBlockSize, 0.f);
        }
        if (rec) {
            rec->postCount.assign(nch, 0); rec->posts.assign(nch, {}); rec->f0amp.assign(nch, 0.f); rec->f0coeff.assign(nch, {});
            for (int i = 0; i < nch; i++) {
                if (floorData[i].isFloor0) { rec->f0amp[i] = floorData[i].Amp; rec->f0coeff[i] = floorData[i].Coeff; rec->postCount[i] = floorData[i].Amp > 0.f ? 1 : 0; }
                else { rec->postCount[i] = floorData[i].PostCount; rec->posts[i].assign(floorData[i].Posts, floorData[i].Posts + 64); }
                if (noExecuteChannel[i]) rec->noExecMask |= 1u << i;
            }
        }
        for (size_t i = 0; i < map.couplingAngle.size(); i++) {                                  // Mapping.cs:111-119
            if (floorData[map.couplingAngle[i]].ExecuteChannel() || floorData[map.couplingMagnitude[i]].ExecuteChannel()) {
                floorData[map.couplingAngle[i]].ForceEnergy = true;
                floorData[map.couplingMagnitude[i]].ForceEnergy = true;
            }
        }
        for (size_t i = 0; i < map.submapFloor.size(); i++) {                                    // Mapping.cs:122-134
            for (int j = 0; j < nch; j++)
                if (map.submapFloor[i] != map.channelFloor[j] || map.submapResidue[i] != map.channelResidue[j]) floorData[j].ForceNoEnergy = true;
            residues[map.submapResidue[i]].Decode(p, noExecuteChannel, blockSize, buffer, (rec && i == 0) ? rec : nullptr);
        }
        if (rec) for (int i = 0; i < nch; i++) if (floorData[i].ExecuteChannel()) rec->execMask |= 1u << i;
        for (int i = (int)map.couplingAngle.size() - 1; i >= 0; i--) {                           // Mapping.cs:137-182
            if (floorData[map.couplingAngle[i]].ExecuteChannel() || floorData[map.couplingMagnitude[i]].ExecuteChannel())
                InverseCouple(buffer[map.couplingMagnitude[i]].data(), buffer[map.couplingAngle[i]].data(), halfBlockSize);
        }
        if (rec && recordDense) rec->spectrum.assign((size_t)nch * halfBlockSize, 0.f);
        for (int c = 0; c < nch; c++) {                                                          // Mapping.cs:185-197
            if (floorData[c].ExecuteChannel()) {
                floors[map.channelFloor[c]].Apply(floorData[c], blockSize, buffer[c].data());
                if (rec && recordDense) std::copy(buffer[c].begin(), buffer[c].begin() + halfBlockSize, rec->spectrum.begin() + (size_t)c * halfBlockSize);
                mdct.Reverse(buffer[c].data(), blockSize);
            } else {
                if (rec && recordDense) std::copy(buffer[c].begin(), buffer[c].begin() + halfBlockSize, rec->spectrum.begin() + (size_t)c * halfBlockSize);
                std::fill(buffer[c].begin() + halfBlockSize, buffer[c].begin() + 2 * halfBlockSize, 0.f);
            }
        }
    }

    // Mode.Decode Mode.cs:153-170
    bool ModeDecode(int modeIdx, Packet& p, std::vector<std::vector<float>>& buffer, int& start, int& valid, int& total, FrameRec* rec) {
        const Mode& mode = modes[modeIdx];
        int windowIndex;
        if (mode.GetPacketInfo(p, windowIndex, start, valid, total)) {
            if (rec) { rec->mode = modeIdx; rec->blockSize = mode.blockSize; rec->windowIndex = windowIndex; rec->start = start; rec->valid = valid; rec->total = total; }
            MappingDecodePacket(mappings[mode.mapping], p, mode.blockSize, buffer, rec);
            const std::vector<float>& window = mode.windows[windowIndex];
            for (int i = 0; i < mode.blockSize; i++) for (int ch = 0; ch < channels; ch++) buffer[ch][i] *= window[i];
            if (rec && recordDense) {
                rec->block.resize((size_t)channels * mode.blockSize);
                for (int ch = 0; ch < channels; ch++) std::copy(buffer[ch].begin(), buffer[ch].begin() + mode.blockSize, rec->block.begin() + (size_t)ch * mode.blockSize);
            }
            return true;
        }
        return false;
    }

    // StreamDecoder.DecodeNextPacket StreamDecoder.cs:465-530
    std::vector<std::vector<float>>* DecodeNextPacket(int& start, int& valid, int& total, bool& isEndOfStream, bool& hasSamplePos, int64_t& samplePos, FrameRec* rec) {
        Packet* packet = GetNextPacket();
        hasSamplePos = false; samplePos = 0; start = valid = total = 0;
        if (!packet) { isEndOfStream = true; return nullptr; }
        isEndOfStream = packet->isEndOfStream;
        if (packet->isResync) hasPosition = false;
        if (packet->ReadBit()) return nullptr;
        int modeIdx = (int)packet->ReadBits(modeFieldBits);
        if (modeIdx >= (int)modes.size()) throw InvalidData("mode index out of range (IndexOutOfRangeException in the reference)");
        if (!nextBuf) {
            std::vector<std::vector<float>>& nb = (prevBuf == &bufA) ? bufB : bufA;
            nb.assign((size_t)channels, std::vector<float>((size_t)block1Size, 0.f));
            nextBuf = &nb;
        }
        if (ModeDecode(modeIdx, *packet, *nextBuf, start, valid, total, rec)) {
            hasSamplePos = packet->hasGranule; samplePos = packet->granule;
            return nextBuf;
        }
        return nullptr;
    }

    // StreamDecoder.ReadNextPacket StreamDecoder.cs:417-463
    bool ReadNextPacket(int bufferedSamples, bool& hasSamplePos, int64_t& samplePos) {
        int startIndex, validLen, totalLen; bool isEndOfStream;
        FrameRec* rec = nullptr;
        if (record) { recs.emplace_back(); rec = &recs.back(); }
        auto* curPacket = DecodeNextPacket(startIndex, validLen, totalLen, isEndOfStream, hasSamplePos, samplePos, rec);
        eosFound |= isEndOfStream;
        if (!curPacket) { if (rec) rec->ok = false; return false; }
        if (hasSamplePos && isEndOfStream) {                                                     // :429-437
            int64_t actualEnd = currentPosition + bufferedSamples + validLen - startIndex;
            int diff = (int)(samplePos - actualEnd);
            if (diff < 0) validLen += diff;
        }
        if (rec) { rec->ok = true; rec->validTrimmed = validLen; }
        if (prevPacketEnd > 0) {                                                                 // :440-445, OverlapBuffers :532-541
            int ps = prevPacketStart, ns = startIndex;
            for (; ps < prevPacketStop; ps++, ns++) for (int c = 0; c < channels; c++) (*curPacket)[c][ns] += (*prevBuf)[c][ps];
            prevPacketStart = startIndex;
        } else if (!prevBuf) {
            prevPacketStart = validLen;                                                          // :446-450
        }
        nextBuf = prevBuf;                                                                       // :455-461
        prevPacketEnd = validLen; prevPacketStop = totalLen; prevBuf = curPacket;
        return true;
    }

    // StreamDecoder.Read StreamDecoder.cs:320-389 (+ ClippingCopyBuffer 391-402 / CopyBuffer 404-415)
    int Read(float* buffer, int offset, int count) {
        if (count % channels != 0) throw std::out_of_range("count must be a multiple of Channels");
        if (count == 0) return 0;
        int idx = offset, tgt = offset + count;
        while (idx < tgt) {
            if (prevPacketStart == prevPacketEnd) {
                if (eosFound) { nextBuf = nullptr; prevBuf = nullptr; break; }
                bool hasPos; int64_t pos;
                if (!ReadNextPacket((idx - offset) / channels, hasPos, pos)) prevPacketEnd = prevPacketStop;
                if (hasPos && !hasPosition) {
                    hasPosition = true;
                    currentPosition = pos - (prevPacketEnd - prevPacketStart) - (idx - offset) / channels;
                }
            }
            int copyLen = std::min((tgt - idx) / channels, prevPacketEnd - prevPacketStart);
            if (copyLen <= 0 && prevPacketStart != prevPacketEnd) break;   // reference would spin forever here (EOS trim below start)
            if (copyLen > 0) {
                for (; copyLen > 0; prevPacketStart++, copyLen--)
                    for (int ch = 0; ch < channels; ch++) {
                        float v = (*prevBuf)[ch][prevPacketStart];
                        buffer[idx++] = clipSamples ? ClipValue(v, hasClipped) : v;
                    }
            }
        }
        count = idx - offset;
        currentPosition += count / channels;
        return count;
    }
};

// ------------------------------------------------------------------------------------
// Stand-alone synthesis from boundary records (the CPU baseline of the GPU hot path):
// residue replay -> inverse coupling -> floor apply -> IMDCT -> window -> OLA -> clip/interleave,
// with the same bookkeeping as StreamDecoder.ReadNextPacket/Read.
// ------------------------------------------------------------------------------------
struct SynthFrame {     // flat, ctypes-friendly mirror of FrameRec (see oracle.py)
    int32_t ok, mode, windowIndex, start, valid, total;
    uint32_t execMask;
    int32_t resDecoded, resStreams, resPartitions;
    int64_t postsOff;      // index into posts (int32) -- channels*64 values per frame
    int64_t postCountOff;  // index into postCounts (int32) -- channels values per frame
    int64_t classesOff;    // into classes (uint8)
    int64_t entriesOff; int32_t entryCount;   // into entries (int32)
    int32_t pad;
};

}  // namespace orc

// =====================================================================================
// C API (ctypes)
// =====================================================================================
using namespace orc;

struct OrcHandle { Decoder dec; std::string error; std::vector<uint8_t> file; };

static thread_local std::string g_lastError;

extern "C" {

const char* orc_last_error() { return g_lastError.c_str(); }

void* orc_open(const uint8_t* data, size_t len) {
    auto h = std::make_unique<OrcHandle>();
    try { h->file.assign(data, data + len); h->dec.Open(h->file.data(), h->file.size()); }
    catch (const std::exception& e) { g_lastError = e.what(); return nullptr; }
    return h.release();
}
// packet list: `data` = all packets back to back, sizes[i] bytes each; flags bit0 = hasGranule, bit1 = EOS, bit2 = resync
void* orc_open_packets(const uint8_t* data, const int64_t* sizes, const int64_t* granules, const uint8_t* flags, int64_t n) {
    auto h = std::make_unique<OrcHandle>();
    try {
        std::vector<std::unique_ptr<Packet>> list; size_t off = 0;
        for (int64_t i = 0; i < n; i++) {
            auto p = std::make_unique<Packet>();
            p->data.assign(data + off, data + off + sizes[i]); off += (size_t)sizes[i];
            p->hasGranule = flags[i] & 1; p->granule = granules[i]; p->isEndOfStream = (flags[i] & 2) != 0; p->isResync = (flags[i] & 4) != 0;
            list.push_back(std::move(p));
        }
        h->dec.OpenPackets(std::move(list));
    } catch (const std::exception& e) { g_lastError = e.what(); return nullptr; }
    return h.release();
}
void orc_close(void* h) { delete (OrcHandle*)h; }

// ---- setup introspection for the tables Codebook-level calls do not cover
void orc_counts(void* hh, int64_t* out) {   // nFloors, nResidues, nMappings, nModes
    Decoder& d = ((OrcHandle*)hh)->dec;
    out[0] = (int64_t)d.floors.size(); out[1] = (int64_t)d.residues.size(); out[2] = (int64_t)d.mappings.size(); out[3] = (int64_t)d.modes.size();
}
// out[0..3] = type, n_posts, multiplier, range; then xList[64], lNeigh[64], hNeigh[64], sortIdx[64]
int orc_floor_info(void* hh, int fi, int32_t* out) {
    Decoder& d = ((OrcHandle*)hh)->dec; if (fi < 0 || fi >= (int)d.floors.size()) return -1;
    const Floor& f = d.floors[fi];
    std::memset(out, 0, sizeof(int32_t) * 260);
    out[0] = f.type; out[1] = (int)f.xList.size(); out[2] = f.multiplier; out[3] = f.range;
    if (f.type == 0) {   // floor 0: out[4..] = order, rate, bark_map_size, ampBits, ampOfs, number of books, book numbers
        out[1] = 0; out[4] = f.order; out[5] = f.rate; out[6] = f.bark_map_size; out[7] = f.ampBits; out[8] = f.ampOfs; out[9] = (int)f.f0books.size();
        for (size_t k = 0; k < f.f0books.size() && k < 16; k++) out[10 + k] = f.f0books[k];
        return 0;
    }
    for (size_t k = 0; k < f.xList.size() && k < 64; k++) { out[4 + k] = f.xList[k]; out[68 + k] = f.lNeigh[k]; out[132 + k] = f.hNeigh[k]; out[196 + k] = f.sortIdx[k]; }
    return 0;
}
// out[0..7] = type, begin, end, partitionSize, classifications, maxStages, classBook, channels; cascade[64]; books[64][8] (-1 = none)
int orc_residue_info(void* hh, int ri, int32_t* out) {
    Decoder& d = ((OrcHandle*)hh)->dec; if (ri < 0 || ri >= (int)d.residues.size()) return -1;
    const Residue& r = d.residues[ri];
    for (int i = 0; i < 8 + 64 + 512; i++) out[i] = (i >= 72) ? -1 : 0;
    out[0] = r.type; out[1] = r.begin; out[2] = r.end; out[3] = r.partitionSize; out[4] = r.classifications; out[5] = r.maxStages; out[6] = r.classBook;
    out[7] = r.type == 2 ? r.r2channels : r.channels;
    for (int c = 0; c < r.classifications && c < 64; c++) {
        out[8 + c] = r.cascade[c];
        for (size_t st = 0; st < r.books[c].size() && st < 8; st++) out[72 + c * 8 + st] = r.books[c][st];
    }
    return 0;
}
// out[0..3] = couplingSteps, submaps, submapFloor[0], submapResidue[0]; magnitude[256]; angle[256]
int orc_mapping_info(void* hh, int mi, int32_t* out) {
    Decoder& d = ((OrcHandle*)hh)->dec; if (mi < 0 || mi >= (int)d.mappings.size()) return -1;
    const Mapping& m = d.mappings[mi];
    std::memset(out, 0, sizeof(int32_t) * 516);
    out[0] = (int)m.couplingAngle.size(); out[1] = (int)m.submapFloor.size(); out[2] = m.submapFloor[0]; out[3] = m.submapResidue[0];
    for (size_t k = 0; k < m.couplingAngle.size() && k < 256; k++) { out[4 + k] = m.couplingMagnitude[k]; out[260 + k] = m.couplingAngle[k]; }
    return 0;
}
int orc_mode_info(void* hh, int mi, int32_t* out) {   // blockFlag, mapping, blockSize
    Decoder& d = ((OrcHandle*)hh)->dec; if (mi < 0 || mi >= (int)d.modes.size()) return -1;
    out[0] = d.modes[mi].blockFlag ? 1 : 0; out[1] = d.modes[mi].mapping; out[2] = d.modes[mi].blockSize;
    return 0;
}

// info[0..7] = channels, sampleRate, block0, block1, nPackets(total incl. headers), nModes, nBooks, modeFieldBits
void orc_info(void* hh, int64_t* info) {
    Decoder& d = ((OrcHandle*)hh)->dec;
    info[0] = d.channels; info[1] = d.sampleRate; info[2] = d.block0Size; info[3] = d.block1Size;
    info[4] = (int64_t)d.packets.size(); info[5] = (int64_t)d.modes.size(); info[6] = (int64_t)d.books.size(); info[7] = d.modeFieldBits;
}
void orc_set_options(void* hh, int clip, int record, int recordDense) {
    Decoder& d = ((OrcHandle*)hh)->dec; d.clipSamples = clip != 0; d.record = record != 0; d.recordDense = recordDense != 0;
}
// VorbisReader.ReadSamples VorbisReader.cs:336-345
int orc_read_samples(void* hh, float* buffer, int offset, int count) {
    OrcHandle* h = (OrcHandle*)hh;
    try {
        count -= count % h->dec.channels;
        if (count > 0) return h->dec.Read(buffer, offset, count);
        return 0;
    } catch (const std::exception& e) { g_lastError = e.what(); return -1; }
}
int orc_has_clipped(void* hh) { return ((OrcHandle*)hh)->dec.hasClipped ? 1 : 0; }

// ---- raw packet access (so tests can feed identical packets to the product's unpacker)
int64_t orc_packet_count(void* hh) { return (int64_t)((OrcHandle*)hh)->dec.packets.size(); }
int64_t orc_packet_size(void* hh, int64_t i) { return (int64_t)((OrcHandle*)hh)->dec.packets[(size_t)i]->data.size(); }
void orc_packet_get(void* hh, int64_t i, uint8_t* out, int64_t* meta) {
    Packet& p = *((OrcHandle*)hh)->dec.packets[(size_t)i];
    std::memcpy(out, p.data.data(), p.data.size());
    meta[0] = p.hasGranule ? 1 : 0; meta[1] = p.granule; meta[2] = p.isEndOfStream ? 1 : 0; meta[3] = p.isResync ? 1 : 0;
}

// ---- boundary records
int64_t orc_rec_count(void* hh) { return (int64_t)((OrcHandle*)hh)->dec.recs.size(); }
// meta[0..11]: ok, mode, blockSize, windowIndex, start, valid, total, validTrimmed, execMask, noExecMask, resDecoded, resStreams
// meta[12..15]: resPartitions, nEntries, nClasses, hasDense
void orc_rec_meta(void* hh, int64_t i, int64_t* meta) {
    FrameRec& r = ((OrcHandle*)hh)->dec.recs[(size_t)i];
    meta[0] = r.ok; meta[1] = r.mode; meta[2] = r.blockSize; meta[3] = r.windowIndex; meta[4] = r.start; meta[5] = r.valid; meta[6] = r.total;
    meta[7] = r.validTrimmed; meta[8] = r.execMask; meta[9] = r.noExecMask; meta[10] = r.resDecoded; meta[11] = r.resStreams;
    meta[12] = r.resPartitions; meta[13] = (int64_t)r.entries.size(); meta[14] = (int64_t)r.classes.size(); meta[15] = r.spectrum.empty() ? 0 : 1;
}
// posts: channels*64 int32, postCounts: channels int32
void orc_rec_floor1(void* hh, int64_t i, int32_t* posts, int32_t* postCounts) {
    Decoder& d = ((OrcHandle*)hh)->dec; FrameRec& r = d.recs[(size_t)i];
    for (int c = 0; c < d.channels; c++) {
        postCounts[c] = c < (int)r.postCount.size() ? r.postCount[c] : 0;
        for (int k = 0; k < 64; k++) posts[c * 64 + k] = (c < (int)r.posts.size() && k < (int)r.posts[c].size()) ? r.posts[c][k] : 0;
    }
}
void orc_rec_residue(void* hh, int64_t i, uint8_t* classes, int32_t* entries) {
    FrameRec& r = ((OrcHandle*)hh)->dec.recs[(size_t)i];
    if (!r.classes.empty()) std::memcpy(classes, r.classes.data(), r.classes.size());
    if (!r.entries.empty()) std::memcpy(entries, r.entries.data(), r.entries.size() * sizeof(int32_t));
}
void orc_rec_dense(void* hh, int64_t i, float* spectrum, float* block) {
    FrameRec& r = ((OrcHandle*)hh)->dec.recs[(size_t)i];
    if (spectrum && !r.spectrum.empty()) std::memcpy(spectrum, r.spectrum.data(), r.spectrum.size() * sizeof(float));
    if (block && !r.block.empty()) std::memcpy(block, r.block.data(), r.block.size() * sizeof(float));
}

// ---- setup introspection (tests compare the product's tables against these)
int orc_book_info(void* hh, int b, int64_t* info) {   // dims, entries, mapType, tableLen
    Decoder& d = ((OrcHandle*)hh)->dec; if (b < 0 || b >= (int)d.books.size()) return -1;
    info[0] = d.books[b].Dimensions; info[1] = d.books[b].Entries; info[2] = d.books[b].MapType; info[3] = (int64_t)d.books[b].lookupTable.size();
    return 0;
}
void orc_book_table(void* hh, int b, float* out) {
    Decoder& d = ((OrcHandle*)hh)->dec; std::memcpy(out, d.books[b].lookupTable.data(), d.books[b].lookupTable.size() * sizeof(float));
}
void orc_book_lengths(void* hh, int b, int32_t* out) {
    Decoder& d = ((OrcHandle*)hh)->dec; std::memcpy(out, d.books[b].lengths.data(), d.books[b].lengths.size() * sizeof(int32_t));
}
// window of mode m, index w (0..3 long / 0 short) -> blockSize floats ; returns blockSize
int orc_mode_window(void* hh, int m, int w, float* out) {
    Decoder& d = ((OrcHandle*)hh)->dec; if (m < 0 || m >= (int)d.modes.size()) return -1;
    const Mode& mo = d.modes[m]; if (w < 0 || w >= (int)mo.windows.size()) return -1;
    if (out) std::memcpy(out, mo.windows[w].data(), mo.windows[w].size() * sizeof(float));
    return mo.blockSize;
}

// ---- stand-alone pieces -----------------------------------------------------------
// Mdct.Reverse on a caller buffer of n floats (input in the first n/2).  Mdct.cs:13-21
int orc_mdct_reverse(float* buffer, int n) {
    try { MdctImpl m(n); m.CalcReverse(buffer); return 0; } catch (const std::exception& e) { g_lastError = e.what(); return -1; }
}
// twiddles as the reference builds them: a[n/2], b[n/2], c[n/4], bitrev[n/8]
void orc_mdct_tables(int n, float* a, float* b, float* c, uint16_t* bitrev) {
    MdctImpl m(n);
    std::memcpy(a, m._a.data(), m._a.size() * 4); std::memcpy(b, m._b.data(), m._b.size() * 4);
    std::memcpy(c, m._c.data(), m._c.size() * 4); std::memcpy(bitrev, m._bitrev.data(), m._bitrev.size() * 2);
}
void orc_calc_window(int prev, int cur, int next, float* out) { auto w = CalcWindow(prev, cur, next); std::memcpy(out, w.data(), w.size() * 4); }
// Floor1 curve alone: xList/lNeigh/hNeigh/sortIdx come from a parsed setup (floor index fi)
int orc_floor1_apply(void* hh, int fi, const int32_t* posts, int postCount, int blockSize, float* residue) {
    Decoder& d = ((OrcHandle*)hh)->dec;
    try {
        if (fi < 0 || fi >= (int)d.floors.size() || d.floors[fi].type != 1) throw InvalidData("not a floor1");
        FloorData fd; for (int k = 0; k < 64; k++) fd.Posts[k] = posts[k]; fd.PostCount = postCount;
        d.floors[fi].Apply1(fd, blockSize, residue); return 0;
    } catch (const std::exception& e) { g_lastError = e.what(); return -1; }
}
void orc_inverse_couple(float* mag, float* ang, int n) { InverseCouple(mag, ang, n); }
float orc_clip(float v, int* clipped) { bool c = *clipped != 0; float r = ClipValue(v, c); *clipped = c ? 1 : 0; return r; }
float orc_inverse_db(int y) { return inverse_dB(y & 255); }

// ---- batch synthesis from boundary records (CPU baseline of the GPU hot path) -------
// Runs, for frames[0..nFrames), exactly what Mapping.DecodePacket (after the bit-reading half),
// Mode.Decode's window multiply, StreamDecoder.ReadNextPacket's OLA bookkeeping and
// ClippingCopyBuffer do, on a fresh decoder state (stream start at frame 0).
// pcm receives interleaved samples; returns samples-per-channel written, or -1.
// `clipped` (optional) receives HasClipped.  threads > 1: frames are split into contiguous
// runs, one std::thread each, with a 1-frame halo (frame start-1 recomputed, output dropped).
int64_t orc_synth_batch(void* hh, const SynthFrame* frames, int64_t nFrames, const int32_t* posts, const int32_t* postCounts,
                        const uint8_t* classes, const int32_t* entries, float* pcm, int64_t pcmCapPerChannel, int* clipped);
// Floor 0 payload for the next orc_synth_batch calls: [frame][channel][stride] floats, element 0 = Amp, 1.. = Coeff
// (what Floor0.Unpack leaves behind, Floor0.cs:98-150).  nullptr clears it.
void orc_set_floor0_payload(const float* payload, int stride);

}  // extern "C"

namespace orc {

static const float* g_f0Payload = nullptr; static int g_f0Stride = 0;

struct SynthState {
    Decoder* d; Mdct mdct;
    std::vector<std::vector<float>> bufA, bufB; std::vector<std::vector<float>>* nextBuf = nullptr; std::vector<std::vector<float>>* prevBuf = nullptr;
    int prevStart = 0, prevEnd = 0, prevStop = 0; bool hasClipped = false;
};

// synthesis half of Mapping.DecodePacket + Mode.Decode for one recorded frame
static void SynthBlock(SynthState& s, const SynthFrame& f, int64_t frameIndex, const int32_t* posts, const int32_t* postCounts, const uint8_t* classes, const int32_t* entries,
                       std::vector<std::vector<float>>& buffer) {
    Decoder& d = *s.d; const Mode& mode = d.modes[f.mode]; const Mapping& map = d.mappings[mode.mapping];
    int blockSize = mode.blockSize, half = blockSize >> 1, nch = d.channels;
    for (int i = 0; i < nch; i++) std::fill(buffer[i].begin(), buffer[i].begin() + half, 0.f);
    if (f.resDecoded) d.residues[map.submapResidue[0]].Replay(classes + f.classesOff, f.resStreams, f.resPartitions, entries + f.entriesOff, f.entryCount, blockSize, buffer);
    auto exec = [&](int c) { return (f.execMask >> c) & 1u; };
    for (int i = (int)map.couplingAngle.size() - 1; i >= 0; i--)
        if (exec(map.couplingAngle[i]) || exec(map.couplingMagnitude[i]))
            InverseCouple(buffer[map.couplingMagnitude[i]].data(), buffer[map.couplingAngle[i]].data(), half);
    for (int c = 0; c < nch; c++) {
        if (exec(c)) {
            const Floor& fl = d.floors[map.channelFloor[c]];
            FloorData fd;
            for (int k = 0; k < 64; k++) fd.Posts[k] = posts[f.postsOff + c * 64 + k];
            fd.PostCount = postCounts[f.postCountOff + c];
            if (fl.type != 1) {
                if (!g_f0Payload) throw InvalidData("orc_synth_batch: floor 0 stream without orc_set_floor0_payload");
                const float* pl = g_f0Payload + ((size_t)frameIndex * nch + c) * g_f0Stride;
                fd.isFloor0 = true; fd.Amp = pl[0]; fd.Coeff.assign(pl + 1, pl + 1 + fl.order); fd.Coeff.push_back(0.f);
                fl.Apply0(fd, blockSize, buffer[c].data());
            } else
            fl.Apply1(fd, blockSize, buffer[c].data());
            s.mdct.Reverse(buffer[c].data(), blockSize);
        } else std::fill(buffer[c].begin() + half, buffer[c].begin() + blockSize, 0.f);
    }
    const std::vector<float>& window = mode.windows[f.windowIndex];
    for (int i = 0; i < blockSize; i++) for (int ch = 0; ch < nch; ch++) buffer[ch][i] *= window[i];
}

// Processes frames [lo,hi) writing interleaved PCM at out (already offset); `halo`: frame lo-1 is
// decoded first for its tail only.  Returns samples per channel written.
static int64_t SynthRun(Decoder& d, const SynthFrame* frames, int64_t lo, int64_t hi, bool halo, const int32_t* posts, const int32_t* postCounts,
                        const uint8_t* classes, const int32_t* entries, float* out, int64_t cap, bool& clippedOut) {
    SynthState s; s.d = &d;
    s.bufA.assign((size_t)d.channels, std::vector<float>((size_t)d.block1Size, 0.f)); s.bufB = s.bufA;
    int64_t written = 0; int nch = d.channels;
    auto emit = [&](int from, int to) {
        for (int i = from; i < to; i++) {
            if (written >= cap) throw InvalidData("pcm capacity exceeded");
            for (int ch = 0; ch < nch; ch++) out[written * nch + ch] = ClipValue((*s.prevBuf)[ch][i], s.hasClipped);
            ++written;
        }
    };
    // The run starts from a fresh decoder state (prevBuf == null): its first processed frame -- frame 0
    // of the stream, or the halo frame lo-1 of a shard -- emits nothing and only leaves its tail
    // (StreamDecoder.cs:446-450).
    for (int64_t fi = halo ? lo - 1 : lo; fi < hi; fi++) {
        const SynthFrame& f = frames[fi];
        if (!f.ok) {                                   // failed packet: drain previous tail (StreamDecoder.cs:352-356)
            s.prevEnd = s.prevStop;
            if (s.prevBuf) emit(s.prevStart, s.prevEnd);
            s.prevStart = s.prevEnd;
            continue;
        }
        std::vector<std::vector<float>>& cur = (s.prevBuf == &s.bufA) ? s.bufB : s.bufA;
        SynthBlock(s, f, fi, posts, postCounts, classes, entries, cur);
        int start = f.start, valid = f.valid, total = f.total;      // valid: after the EOS trim
        if (s.prevEnd > 0) {                           // StreamDecoder.cs:440-445 (prevStart == prevEnd here)
            int ps = s.prevStart, ns = start;
            for (; ps < s.prevStop; ps++, ns++) for (int c = 0; c < nch; c++) cur[c][ns] += (*s.prevBuf)[c][ps];
            s.prevStart = start;
        } else if (!s.prevBuf) {
            s.prevStart = valid;                       // StreamDecoder.cs:446-450
        }
        s.prevEnd = valid; s.prevStop = total; s.prevBuf = &cur;
        emit(s.prevStart, s.prevEnd);
        s.prevStart = s.prevEnd;
    }
    clippedOut = s.hasClipped;
    return written;
}

}  // namespace orc

#include <thread>

extern "C" {

static int g_synthThreads = 1;
void orc_set_threads(int n) { g_synthThreads = n < 1 ? 1 : n; }
void orc_set_floor0_payload(const float* payload, int stride) { orc::g_f0Payload = payload; orc::g_f0Stride = stride; }

// samples per channel each frame contributes (same bookkeeping as SynthRun), for sharding
static int64_t FrameOutLen(const SynthFrame* frames, int64_t i) {
    const SynthFrame& f = frames[i];
    // find previous ok-state
    if (!f.ok) {
        // drain: emits prev tail if the previous frame was ok
        if (i > 0 && frames[i - 1].ok) return frames[i - 1].total - frames[i - 1].valid;
        return 0;
    }
    if (i == 0) return 0;
    return f.valid - f.start;
}

int64_t orc_synth_batch(void* hh, const SynthFrame* frames, int64_t nFrames, const int32_t* posts, const int32_t* postCounts,
                        const uint8_t* classes, const int32_t* entries, float* pcm, int64_t pcmCapPerChannel, int* clipped) {
    Decoder& d = ((OrcHandle*)hh)->dec;
    try {
        int T = g_synthThreads; if (T > nFrames) T = (int)std::max<int64_t>(1, nFrames);
        if (T <= 1) {
            bool c = false; int64_t w = SynthRun(d, frames, 0, nFrames, false, posts, postCounts, classes, entries, pcm, pcmCapPerChannel, c);
            if (clipped) *clipped = c;
            return w;
        }
        // multi-threaded: only for batches without drain frames (all ok); offsets by prefix sum
        for (int64_t i = 0; i < nFrames; i++) if (!frames[i].ok) throw InvalidData("threads>1 requires all frames ok");
        std::vector<int64_t> off((size_t)nFrames + 1, 0);
        for (int64_t i = 0; i < nFrames; i++) off[i + 1] = off[i] + FrameOutLen(frames, i);
        if (off[nFrames] > pcmCapPerChannel) throw InvalidData("pcm capacity exceeded");
        std::vector<std::thread> th; std::vector<int> clip((size_t)T, 0); std::vector<std::string> errs((size_t)T);
        for (int t = 0; t < T; t++) {
            int64_t lo = nFrames * t / T, hi = nFrames * (t + 1) / T;
            th.emplace_back([&, t, lo, hi]() {
                try {
                    bool c = false;
                    SynthRun(d, frames, lo, hi, lo > 0, posts, postCounts, classes, entries, pcm + off[lo] * d.channels, off[hi] - off[lo], c);
                    clip[t] = c;
                } catch (const std::exception& e) { errs[t] = e.what(); }
            });
        }
        for (auto& x : th) x.join();
        for (auto& e : errs) if (!e.empty()) throw InvalidData(e);
        if (clipped) { *clipped = 0; for (int c : clip) *clipped |= c; }
        return off[nFrames];
    } catch (const std::exception& e) { g_lastError = e.what(); return -1; }
}

}  // extern "C"
