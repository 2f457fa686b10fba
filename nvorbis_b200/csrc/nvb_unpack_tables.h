// nvb_unpack_tables.h -- layout of the "unpack tables" blob: what the GPU-side packet unpacker (k_unpack, nvb_unpack.cu) needs
// beyond the synthesis setup -- Huffman decode tables of every codebook, the floor 1 partition / class structure, the residue
// class books and cascades, coupling steps -- as plain structs in ONE contiguous allocation.  Built by the host half
// (libnvorbis_host.so: nvh_unpack_tables, from the setup header it parsed) and installed with nvb_upload_unpack_tables.
// Shared by both libraries; plain C++ / POD only.  Reference behaviour it encodes: Codebook.cs:59-220 + Huffman.cs:15-86
// (codeword assignment), Floor1.cs:30-133, Residue0.cs:35-117, Residue2.cs:10-14, Mapping.cs:16-93, Mode.cs:24-41.
#pragma once
#include <cstdint>

namespace nvbu {

constexpr uint32_t UNPACK_MAGIC = 0x3155424eu;   // 'NBU1'
constexpr int ROOT_BITS = 10;                     // codewords up to this length resolve with one table load

// Codebook.DecodeScalar (Codebook.cs:294-320) as a root table over the next ROOT_BITS bits (first transmitted bit = bit 0) plus,
// for longer codewords, a chain per root slot.  roots[root_off + idx]: len << 24 | value for a codeword of len <= root_bits;
// len == 0: value = 1 + index of the first chain element (0 = no codeword starts with these bits).
struct UBook { int32_t dims, entries, root_bits, decodable; uint32_t root_off, long_off; int32_t n_long, pad; };
struct ULong { uint32_t code; int32_t value; int32_t next; uint32_t len; };            // next: 1 + index (0 = end of chain), relative to the book's long_off

struct UFloor0 { int32_t order, amp_bits, amp_div, amp_ofs, book_bits, n_books; int16_t books[16]; int32_t pad[2]; };   // Floor0.Init, Floor0.cs:28-51
struct UFloor1 {                                                                       // Floor1.cs:30-133 (type 1); type 0: only f0 is meaningful
    int32_t type, n_parts, ybits, n_posts;
    uint8_t part_class[32]; uint8_t class_dims[16]; uint8_t class_subs[16];
    int16_t class_master[16]; int16_t sub_books[16][8];
    UFloor0 f0;
};
struct UResidue {                                                                      // Residue0.cs:35-117
    int32_t type, begin, end, psize, nclass, class_book, stages, cdims;               // cdims = dimensions of the class book (partitions per class word)
    int32_t partvals;                                                                  // nclass ^ cdims: class words above this are invalid
    uint32_t digits_off;                                                               // into digits[]: [partvals][cdims] class of each partition of a class word (Residue0.cs:100-114)
    int32_t cascade[64]; int16_t books[64][8];
    int32_t pad[2];
};
constexpr int UNPACK_MAX_COUPLING = 256;          // = NVB_MAX_COUPLING (Mapping.cs:28: 8 bits + 1)
struct UMapping { int32_t n_coupling, floor, residue, pad; uint8_t mag[UNPACK_MAX_COUPLING], ang[UNPACK_MAX_COUPLING]; };
struct UMode { int32_t block_flag, mapping; };

struct UHeader {
    uint32_t magic, version;
    uint64_t total_bytes;
    int32_t channels, bs[2], mode_bits;
    int32_t n_books, n_floors, n_residues, n_mappings, n_modes;
    int32_t post_stride;          // int16 per (frame, channel), the same rule as nvb_post_stride()
    int32_t cls_stride;           // class bytes a frame can need (largest streams * partitions of any mode)
    int32_t ent_stride;           // VQ entries a frame can need (largest sum over stages of partitions * entries per partition)
    uint32_t n_roots, n_longs, n_digits;
    int32_t f0_stride;            // floats per (frame, channel) of the type 0 floor records (amplitude + coefficients), 0 = no type 0 floor
    uint64_t off_books, off_roots, off_longs, off_floors, off_residues, off_digits, off_mappings, off_modes;
};

}  // namespace nvbu
