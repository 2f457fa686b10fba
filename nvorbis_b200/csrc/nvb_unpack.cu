// nvb_unpack.cu -- k_unpack: the bit-reading half of Mapping.DecodePacket on the device.
//
// One warp per audio packet, its lane 0 walking the bits: packets are independent given the setup, and inside a packet every
// field's position depends on the Huffman codewords before it, so a packet is one serial walk (32 walks in one warp would
// serialise on their divergent control flow; with a warp each, thousands of walks hide each other's table-load latency) -- Floor1.Unpack per channel (Floor1.cs:135-184), the energy
// flags (Mapping.cs:105-119), then the class words and VQ entry numbers of Residue0.Decode (Residue0.cs:119-178) through
// Codebook.DecodeScalar (Codebook.cs:294-320).  The thread writes the boundary records the host unpacker would have sent
// (posts / classes / entries at fixed per-frame strides) and patches exec_mask / res_decoded / entry_count into the frame's
// plan record, so k_spectrum_* and k_imdct_* run unchanged.  Codeword decode = one load of a 2^10-entry root table per codeword
// (tables in L1/L2), a short chain for longer codewords.  The walk is latency-bound (a dependent table load per codeword);
// 4096 packets take ~0.1 ms, far above what the host threads reach and below the PCIe time of the PCM they produce.
// Results: the same integers as libnvorbis_host.so's nvh_unpack (tests/test_gpu_unpack.py), which equal the oracle's.
#if !defined(NVB_CPU_SHIM)
#include <cuda_runtime.h>
#endif
#include "nvb_internal.h"
#include "nvb_device_core.h"
#include "nvb_unpack_tables.h"

namespace nvb {

#if defined(NVB_CPU_SHIM)
static inline uint32_t nvb_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31)); }
#else
__device__ __forceinline__ uint32_t nvb_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
#endif

#if defined(NVB_CPU_SHIM)
static inline void prefetch_l1_u(const void*) {}
#else
__device__ __forceinline__ void prefetch_l1_u(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif

// Bit cursor over one packet, LSB first (DataPacket.cs:150-283): reading past the end yields zero bits and raises short_.
// The packet store is padded, so the two aligned words around any position inside a packet can always be loaded.
struct DBits {
    const uint32_t* words; uint32_t bit0;       // aligned word that holds the packet's first byte, bit offset of that byte inside it
    uint32_t nbits, pos; bool short_;
    uint32_t lo, hi, cur;                       // the two words around the cursor, cached: reloaded once per 32 bits, not per codeword
    __device__ __forceinline__ void open(const uint8_t* base, uint32_t bytes) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(base);
        words = reinterpret_cast<const uint32_t*>(addr & ~uintptr_t(3)); bit0 = (uint32_t)(addr & 3) * 8u;
        nbits = bytes * 8u; pos = 0u; short_ = false; cur = 0xffffffffu; lo = hi = 0u;
    }
    __device__ __forceinline__ uint32_t left() const { return pos < nbits ? nbits - pos : 0u; }
    __device__ __forceinline__ uint32_t peek32() {
        const uint32_t at = bit0 + pos, idx = at >> 5;
        if (idx != cur) { lo = words[idx]; hi = words[idx + 1]; cur = idx; }
        uint32_t v = nvb_funnel_r(lo, hi, at & 31u);
        const uint32_t l = left();
        if (l < 32u) v = l ? (v & ((1u << l) - 1u)) : 0u;
        return v;
    }
    __device__ __forceinline__ void skip(uint32_t n) { if (n > left()) { pos = nbits; short_ = true; } else pos += n; }
    __device__ __forceinline__ uint32_t read(uint32_t n) {
        if (n == 0) return 0u;
        uint32_t v = peek32();
        if (n < 32u) v &= (1u << n) - 1u;
        skip(n);
        return v;
    }
};

// Codebook.DecodeScalar (Codebook.cs:294-320): -1 when no bit is left or no codeword matches.  B: the book's record, loaded by
// the caller once per run of codewords of that book.
__device__ __forceinline__ int book_decode(const UnpackTables& T, const nvbu::UBook& B, DBits& b) {
    if (!B.decodable || b.left() == 0u) return -1;
    const uint32_t v = b.peek32();
    const uint32_t r = T.roots[B.root_off + (v & ((1u << B.root_bits) - 1u))];
    const uint32_t len = r >> 24;
    if (len) { b.skip(len); return (int)(r & 0xffffffu); }
    for (uint32_t k = r; k != 0u;) {
        const nvbu::ULong L = T.longs[B.long_off + k - 1u];
        const uint32_t mask = L.len >= 32u ? 0xffffffffu : ((1u << L.len) - 1u);
        if ((v & mask) == L.code) { b.skip(L.len); return L.value; }
        k = (uint32_t)L.next;
    }
    return -1;
}

constexpr int UNPACK_WARPS = 4;                                          // packets per CTA: one warp each

__global__ void __launch_bounds__(UNPACK_WARPS * 32) k_unpack(UnpackArgs a, int smem_cls) {
    NVB_DYN_SMEM(dyn_smem);
    nvb_grid_dep_launch();
    const int fi = blockIdx.x * UNPACK_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    nvb_grid_dep_wait();                                                    // the record buffers may still be read by the previous batch's kernels
    if (fi >= a.n_frames) return;
    DevFrame* df = a.frames + a.frame_lo + fi;
    if (df->kind != 0) return;                                              // a drain: nothing to unpack
    const UnpackTables& T = a.T;
    const int C = T.channels;
    const int api = df->api_index;
    const uint32_t p0 = a.offsets[api], p1 = a.offsets[api + 1];
    const nvbu::UMode mode = T.modes[df->mode];
    const nvbu::UMapping& map = T.mappings[mode.mapping];
    const int N = df->n;
    // A packet is one serial walk: lane 0 does it (a warp of 32 walks would serialise on their divergent control flow); the
    // other lanes pull the packet into L1 and clear the frame's post / class records, then leave.
    {
        const nvbu::UResidue& r0 = T.residues[map.residue];
        const int span0 = (r0.type == 2 ? N * C : N) / 2;
        const int nn0 = (r0.end < span0 ? r0.end : span0) - r0.begin;
        const int ncls = nn0 > 0 ? (nn0 / r0.psize) * (r0.type == 2 ? 1 : C) : 0;
        for (uint32_t o = p0 + 128u * lane; o < p1; o += 128u * 32u) prefetch_l1_u(a.data + o);
        int16_t* posts = a.posts + (size_t)api * C * T.post_stride;
        for (int k = lane; k < C * T.post_stride; k += 32) posts[k] = 0;
        if (a.floor0) for (int k = lane; k < C * a.f0_stride; k += 32) a.floor0[(size_t)api * C * a.f0_stride + k] = 0.f;
        // the class bytes are read back partition by partition while the stages are walked: they live in shared memory during
        // the walk (when the setup's stride fits) and are copied out by the whole warp at the end
        uint8_t* cls0 = smem_cls ? dyn_smem + (size_t)(threadIdx.x >> 5) * T.cls_stride : a.classes + (size_t)api * T.cls_stride;
        for (int k = lane; k < ncls; k += 32) cls0[k] = 0;
    }
    __syncwarp();
    int n_cls_out = 0;
    if (lane == 0) {
    DBits b; b.open(a.data + p0, p1 - p0);
    // the host already read the packet type bit, the mode number and the window flags (Mode.cs:119-151): skip them
    b.skip(1u + (uint32_t)T.mode_bits + (mode.block_flag ? 2u : 0u));

    // floors: Floor1.Unpack per channel (Floor1.cs:135-184)
    const nvbu::UFloor1& f = T.floors[map.floor];
    uint32_t live = 0u;
    for (int c = 0; c < C; c++) {
        int16_t* dst = a.posts + ((size_t)api * C + c) * T.post_stride;
        if (f.type == 0) {
            // Floor0.Unpack (Floor0.cs:98-150): amplitude, book number, the LSP coefficients as VQ vectors, then the running sum
            const nvbu::UFloor0& z = f.f0;
            float* pl = a.floor0 + ((size_t)api * C + c) * a.f0_stride;
            float amp = (float)b.read((uint32_t)z.amp_bits);
            if (amp > 0.f) {
                amp = NVB_FMUL(NVB_FDIV(amp, (float)z.amp_div), (float)z.amp_ofs);
                const uint32_t book_num = b.read((uint32_t)z.book_bits);
                if (book_num >= (uint32_t)z.n_books) amp = 0.f;
                else {
                    const int bk = z.books[book_num];
                    const nvbu::UBook vb = T.books[bk];
                    const float* tab = a.vq + a.dbooks[bk].off;
                    const int dims = vb.dims;
                    for (int i = 0; i < z.order && amp > 0.f;) {
                        const int e = book_decode(T, vb, b);
                        if (e < 0) { amp = 0.f; break; }
                        for (int j = 0; i < z.order && j < dims; j++, i++) pl[1 + i] = tab[(size_t)e * dims + j];
                    }
                    if (amp > 0.f) {
                        float last = 0.f;
                        for (int j = 0; j < z.order;) {
                            for (int k = 0; j < z.order && k < dims; j++, k++) pl[1 + j] = NVB_FADD(pl[1 + j], last);
                            last = pl[j];                                   // Coeff[j - 1]
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
        if (b.read(1u)) {
            count = 2;
            dst[1] = (int16_t)b.read((uint32_t)f.ybits); dst[2] = (int16_t)b.read((uint32_t)f.ybits);
            bool failed = false;
            for (int p = 0; p < f.n_parts && !failed; p++) {
                const int cls = f.part_class[p], cdim = f.class_dims[cls], cbits = f.class_subs[cls];
                uint32_t cval = 0u;
                if (cbits > 0) {
                    const int v = book_decode(T, T.books[f.class_master[cls]], b);
                    if (v < 0) { failed = true; break; }
                    cval = (uint32_t)v;
                }
                for (int k = 0; k < cdim; k++) {
                    const int bk = f.sub_books[cls][cval & ((1u << cbits) - 1u)];
                    cval >>= cbits;
                    int y = 0;
                    if (bk >= 0) { y = book_decode(T, T.books[bk], b); if (y < 0) { failed = true; break; } }
                    dst[1 + count] = (int16_t)y;
                    ++count;
                }
            }
            if (failed) count = 0;                                          // "use nothing", Floor1.cs:155-174
        }
        dst[0] = (int16_t)count;
        if (count > 0) live |= 1u << c;
    }
    // energy flags (Mapping.cs:105-119): noExecute is taken before the coupling propagation
    const uint32_t all = C >= 32 ? 0xffffffffu : (1u << C) - 1u;
    const uint32_t no_exec = ~live & all;
    uint32_t exec = live;
    for (int k = 0; k < map.n_coupling; k++)
        if (((exec >> map.ang[k]) | (exec >> map.mag[k])) & 1u) exec |= (1u << map.ang[k]) | (1u << map.mag[k]);

    // residue (Residue0.Decode, Residue0.cs:119-178): runs when any channel is live, over all streams
    const nvbu::UResidue& r = T.residues[map.residue];
    const int span = (r.type == 2 ? N * C : N) / 2;
    const int nn = (r.end < span ? r.end : span) - r.begin;
    uint32_t n_ent = 0u; int res_decoded = 0;
    if (nn > 0 && no_exec != all) {
        res_decoded = 1;
        const int P = nn / r.psize, S = r.type == 2 ? 1 : C;
        uint8_t* cls = smem_cls ? dyn_smem + (size_t)(threadIdx.x >> 5) * T.cls_stride : a.classes + (size_t)api * T.cls_stride;
        uint16_t* ent = a.entries + (size_t)api * T.ent_stride;
        n_cls_out = S * P;
        const uint8_t* digits = T.digits + r.digits_off;
        const nvbu::UBook cbook = T.books[r.class_book];
        bool stop = false;
        for (int stage = 0; stage < r.stages && !stop; stage++) {
            for (int p = 0; p < P && !stop;) {
                if (stage == 0) {
                    for (int st = 0; st < S; st++) {
                        const int w = book_decode(T, cbook, b);
                        if (w < 0 || w >= r.partvals) { stop = true; break; }
                        for (int k = 0; k < r.cdims && p + k < P; k++) cls[st * P + p + k] = digits[w * r.cdims + k];
                    }
                    if (stop) break;
                }
                for (int k = 0; k < r.cdims && p < P && !stop; k++, p++) {
                    for (int st = 0; st < S && !stop; st++) {
                        const int cl = cls[st * P + p];
                        if (!((r.cascade[cl] >> stage) & 1)) continue;
                        const int bk = r.books[cl][stage];
                        if (bk < 0) continue;
                        const nvbu::UBook vbook = T.books[bk];
                        const int dims = vbook.dims;
                        if (r.type == 0) {
                            // all of a partition's entries are read before any is used (Residue0.cs:186-192): they only count once complete
                            const int steps = r.psize / dims;
                            for (int q = 0; q < steps; q++) { const int e = book_decode(T, vbook, b); if (e < 0) { stop = true; break; } ent[n_ent + q] = (uint16_t)e; }
                            if (!stop) n_ent += (uint32_t)steps;
                        } else {
                            for (int q = 0; q < r.psize; q += dims) {              // Residue1.cs:12-23, Residue2.cs:29-44
                                const int e = book_decode(T, vbook, b);
                                if (e < 0) { stop = true; break; }
                                ent[n_ent++] = (uint16_t)e;
                            }
                        }
                    }
                }
            }
        }
    }
    df->exec_mask = exec; df->res_decoded = (uint8_t)res_decoded; df->entry_count = n_ent;
    }
    __syncwarp();
    if (smem_cls) {
        n_cls_out = __shfl_sync(0xffffffffu, n_cls_out, 0);
        const uint8_t* src = dyn_smem + (size_t)(threadIdx.x >> 5) * T.cls_stride;
        uint8_t* dst = a.classes + (size_t)api * T.cls_stride;
        for (int k = lane; k < n_cls_out; k += 32) dst[k] = src[k];
    }
}

int launch_unpack(const UnpackArgs& a, void* stream) {
    if (a.n_frames <= 0) return 0;
    const size_t smem = (size_t)UNPACK_WARPS * (size_t)a.T.cls_stride;
    const int smem_cls = smem <= 40 * 1024 ? 1 : 0;
    NVB_LAUNCHV(k_unpack, (a.n_frames + UNPACK_WARPS - 1) / UNPACK_WARPS, UNPACK_WARPS * 32, smem_cls ? smem : 0, stream, a, smem_cls);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace nvb
