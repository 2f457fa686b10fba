// nvb_host.cpp -- CUDA-free host logic of the C ABI (see nvb_host.h).
//
// Table construction follows the reference's expression order and float/double mix exactly
// (tables are inputs to the GPU arithmetic, so they must be the reference's numbers):
//   Mdct twiddles  Mdct.cs:30-63      window  Mode.cs:15,69-100      overlap  Mode.cs:102-117
// Batch planning restates the bookkeeping of StreamDecoder.ReadNextPacket (StreamDecoder.cs:417-463)
// and the drain rule of StreamDecoder.Read (StreamDecoder.cs:352-356) as a prefix computation.
#include "nvb_host.h"
#include "nvb_fused_core.h"
#include <cmath>
#include <cstdio>
#include <cstring>

namespace nvb {

static const uint32_t k_inverse_db_bits[256] = {
#include "inverse_db_table.inc"
};

namespace {

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

uint32_t bit_reverse32(uint32_t n) {                                    // Utils.cs:16-28
    n = ((n & 0xAAAAAAAAu) >> 1) | ((n & 0x55555555u) << 1);
    n = ((n & 0xCCCCCCCCu) >> 2) | ((n & 0x33333333u) << 2);
    n = ((n & 0xF0F0F0F0u) >> 4) | ((n & 0x0F0F0F0Fu) << 4);
    n = ((n & 0xFF00FF00u) >> 8) | ((n & 0x00FF00FFu) << 8);
    return (n >> 16) | (n << 16);
}

struct BlobWriter {
    std::vector<unsigned char>& buf;
    explicit BlobWriter(std::vector<unsigned char>& b) : buf(b) {}
    uint64_t reserve(size_t bytes) {
        size_t off = (buf.size() + 15) & ~size_t(15);
        buf.resize(off + bytes, 0);
        return off;
    }
    template <class T> T* at(uint64_t off) { return reinterpret_cast<T*>(buf.data() + off); }
};

// Floor0.toBARK (Floor0.cs:81-84): double arithmetic, result rounded to float
float floor0_to_bark(double lsp) {
    return (float)(13.1 * std::atan(0.00074 * lsp) + 2.24 * std::atan(0.0000000185 * lsp * lsp) + .0001 * lsp);
}

// Rising half of the Vorbis window for an overlap of `len` samples (Mode.cs:80-85): the inner sine in
// double with a float pi/2, squared in float, times the float pi/2 in float, outer sine in double.
void window_slope(int len, float* out) {
    const float pi2 = 3.1415926539f / 2;                                 // Mode.cs:15
    for (int i = 0; i < len; i++) {
        float x = (float)std::sin((i + .5) / len * (double)pi2);
        x *= x;
        out[i] = (float)std::sin((double)(x * pi2));
    }
}

// Mode.CalcWindow (Mode.cs:69-100) assembled from the two slopes.
void assemble_window(const float* slope_l, int left, const float* slope_r, int right, int n, float* out) {
    const int lb = n / 4 - left / 2, rb = n - n / 4 - right / 2;
    for (int i = 0; i < n; i++) out[i] = 0.f;
    for (int i = 0; i < left; i++) out[lb + i] = slope_l[i];
    for (int i = lb + left; i < rb; i++) out[i] = 1.0f;
    for (int i = 0; i < right; i++) out[rb + i] = slope_r[right - 1 - i];
}

void mdct_tables(int n, float* A, float* B, float* C, uint16_t* bitrev) {   // Mdct.cs:30-63
    const float pi = 3.14159265358979323846264f;                         // Mdct.cs:9 (a float constant)
    const int n4 = n >> 2, n8 = n >> 3;
    for (int k = 0; k < n4; k++) {
        const float aa = (float)(4 * k) * pi / (float)n;                 // int*float/int evaluated in float
        A[2 * k] = (float)std::cos((double)aa);
        A[2 * k + 1] = (float)-std::sin((double)aa);
        const float ab = (float)(2 * k + 1) * pi / (float)n / 2.f;
        B[2 * k] = (float)std::cos((double)ab) * .5f;
        B[2 * k + 1] = (float)std::sin((double)ab) * .5f;
    }
    for (int k = 0; k < n8; k++) {
        const float ac = (float)(2 * (2 * k + 1)) * pi / (float)n;
        C[2 * k] = (float)std::cos((double)ac);
        C[2 * k + 1] = (float)-std::sin((double)ac);
    }
    const int ld = ilog_u(n) - 1;
    for (int i = 0; i < n8; i++) {
        const int sh = (32 - (ld - 3)) & 31;                             // C# masks shift counts to 5 bits
        bitrev[i] = (uint16_t)((bit_reverse32((uint32_t)i) >> sh) << 2);
    }
}

void fast_tables(int n, float2* tw, float2* fft) {
    const int M = n >> 1, Q = n >> 2;
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < Q; k++) {
        const double a = -pi * (k + 0.125) / M;
        tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
        const double b = -2.0 * pi * k / Q;
        fft[k] = make_float2((float)std::cos(b), (float)std::sin(b));
    }
}

// Lane tables of the fused kernel (FusedTables, nvb_fused_core.h), N = 2048 / 256, evaluated in double.
void build_fused_tables(const float2* tw0, const float2* w64, const float* slope_long, const float* slope_short, float* out) {
    const double pi = 3.14159265358979323846;
    auto tw = [&](int k) { return -pi * (k + 0.125) / 1024.0; };                // angle of tw[k] = exp(-i pi (k + 1/8) / M), M = 1024
    auto put2 = [&](float* p, double ang) { p[0] = (float)std::cos(ang); p[1] = (float)std::sin(ang); };
    for (int l = 0; l < 32; l++) {
        const int ra = l, rb = 63 - l, k0 = l & 7;
        for (int m2 = 0; m2 < 8; m2++) {
            float* p = out + FusedTables::T2 + (m2 * 32 + l) * 4;
            put2(p, tw(ra) - 2.0 * pi * ((ra * m2) & 511) / 512.0);
            put2(p + 2, tw(rb) - 2.0 * pi * ((rb * m2) & 511) / 512.0);
        }
        for (int j = 0; j < 4; j++) {
            float* p = out + FusedTables::T3 + (j * 32 + l) * 4;
            put2(p, -2.0 * pi * ((k0 * (2 * j + 1)) & 63) / 64.0);
            put2(p + 2, -2.0 * pi * ((k0 * (2 * j + 2)) & 63) / 64.0);
        }
        float* p = out + FusedTables::T4 + l * 4;
        put2(p, tw(fused_na0(l))); put2(p + 2, tw(fused_nb0(l)));
    }
    for (int i = 0; i < 1024; i++) out[FusedTables::WIN + i] = slope_long[i];
    for (int i = 0; i < 128; i++) out[FusedTables::WIN0 + i] = slope_short[i];
    for (int k = 0; k < 64; k++) {
        out[FusedTables::TW0 + 2 * k] = tw0[k].x; out[FusedTables::TW0 + 2 * k + 1] = tw0[k].y;
        out[FusedTables::W64 + 2 * k] = w64[k].x; out[FusedTables::W64 + 2 * k + 1] = w64[k].y;
    }
}

// FNV-1a 64 over the blob body (everything behind the header): a damaged blob is rejected before any offset is trusted
uint64_t blob_body_hash(const unsigned char* blob, size_t bytes) {
    uint64_t hsh = 1469598103934665603ull;
    for (size_t i = sizeof(BlobHeader); i < bytes; i++) { hsh ^= blob[i]; hsh *= 1099511628211ull; }
    return hsh;
}

// The kernel-selection level of one residue (DevResidue.fast) and its per-(class, stage) entry counts, from the plain
// residue fields and the book dimensions.  build_blob stores the result; validate_blob recomputes and compares, so that a
// blob cannot steer a residue onto a kernel whose alignment / power-of-two assumptions it does not meet.
struct ResidueDerived { int pshift, fast; int16_t cnt[NVB_MAX_CLASSES][NVB_MAX_STAGES]; uint8_t coded[NVB_MAX_CLASSES]; };
template <class BookDims>
ResidueDerived derive_residue(int type, int begin, int psize, int nclass, int stages, const int32_t* cascade, const int16_t (*books)[NVB_MAX_STAGES], int C, BookDims dims_of) {
    ResidueDerived d; std::memset(&d, 0, sizeof d);
    d.pshift = is_pow2(psize) ? ilog_u(psize) - 1 : -1;
    bool fast = d.pshift >= 0 && psize <= 8192 && stages >= 1;
    if (C > NVB_FAST_CHANNELS) fast = false;                                // the specialised kernels keep a bin's channels in registers: general kernel beyond 8
    if (type == 2 && (begin % C != 0 || psize % C != 0)) fast = false;     // Residue2.cs:27 truncation case
    for (int c = 0; c < NVB_MAX_CLASSES; c++)
        for (int st = 0; st < NVB_MAX_STAGES; st++) {
            const bool coded = c < nclass && st < stages && ((cascade[c] >> st) & 1) && books[c][st] >= 0;
            if (!coded) continue;
            const int dims = dims_of(books[c][st]);
            const int cnt = type == 0 ? psize / dims : (psize + dims - 1) / dims;   // Residue0.cs:183 / Residue1.cs:12 / Residue2.cs:28
            if (cnt > 32767 || !is_pow2(dims) || dims > psize) fast = false;
            d.cnt[c][st] = (int16_t)(cnt > 32767 ? 32767 : cnt);
            if (cnt > 0) d.coded[c] |= (uint8_t)(1u << st);
        }
    d.fast = fast ? 1 : 0;
    // level 2: the plane kernel (k_spectrum_planes): one interleaved stream (type 2, or a single channel of type 1) whose
    // partitions start on multiples of G = max(4, C) floats
    const int G = C > 4 ? C : 4;
    if (fast && (type == 2 || (type == 1 && C == 1)) && begin % G == 0 && psize % G == 0 && is_pow2(C)) d.fast = 2;
    // level 3: k_spectrum_run works on groups of 8 consecutive values of the stream
    if (d.fast == 2 && begin % 8 == 0 && psize % 8 == 0) d.fast = 3;
    return d;
}

int fail(std::string& err, int code, const char* fmt, long long a = 0, long long b = 0) {
    char tmp[256];
    std::snprintf(tmp, sizeof tmp, fmt, a, b);
    err = tmp;
    return code;
}

}  // namespace

Overlap nominal_overlap(const BlobHeader& h, int block_flag, int window) {
    Overlap o;
    if (!block_flag) { o.start = 0; o.valid = h.bs[0] / 2; o.total = h.bs[0]; return o; }   // Mode.cs:144-147
    const int n = h.bs[1];
    const int prev = (window & 1) ? h.bs[1] : h.bs[0], next = (window & 2) ? h.bs[1] : h.bs[0];
    o.start = n / 4 - prev / 4;                                          // Mode.cs:102-117
    o.total = n / 4 * 3 + next / 4;
    o.valid = o.total - next / 4 * 2;
    return o;
}

void resolve_setup(const unsigned char* base, const BlobHeader& h, DevSetup& S) {
    S.channels = h.channels; S.bs[0] = h.bs[0]; S.bs[1] = h.bs[1];
    S.post_stride = h.post_stride; S.max_items = h.max_items; S.spectrum_fast = h.spectrum_fast;
    S.max_stages = h.max_stages > 0 ? h.max_stages : 1;
    S.ci_total = h.ci_total; S.n_residues = h.n_residues;
    S.books = reinterpret_cast<const DevBook*>(base + h.off_books);
    S.vq = reinterpret_cast<const float*>(base + h.off_vq); S.n_vq = (int64_t)h.n_vq;
    S.floors = reinterpret_cast<const DevFloor1*>(base + h.off_floors);
    S.residues = reinterpret_cast<const DevResidue*>(base + h.off_residues);
    S.mappings = reinterpret_cast<const DevMapping*>(base + h.off_mappings);
    S.modes = reinterpret_cast<const DevMode*>(base + h.off_modes);
    S.win_short = reinterpret_cast<const float*>(base + h.off_win_short);
    S.win_long = reinterpret_cast<const float*>(base + h.off_win_long);
    for (int i = 0; i < 2; i++) {
        S.A[i] = reinterpret_cast<const float*>(base + h.off_mdct_a[i]);
        S.B[i] = reinterpret_cast<const float*>(base + h.off_mdct_b[i]);
        S.C[i] = reinterpret_cast<const float*>(base + h.off_mdct_c[i]);
        S.bitrev[i] = reinterpret_cast<const uint16_t*>(base + h.off_bitrev[i]);
        S.tw[i] = reinterpret_cast<const float2*>(base + h.off_tw[i]);
        S.fft[i] = reinterpret_cast<const float2*>(base + h.off_fft[i]);
    }
    S.db = reinterpret_cast<const float*>(base + h.off_db);
    S.fused_tab = h.off_fused_tab ? reinterpret_cast<const float*>(base + h.off_fused_tab) : nullptr;
    S.ci = reinterpret_cast<const CiRec*>(base + h.off_ci);
    S.bin2k = reinterpret_cast<const uint8_t*>(base + h.off_bin2k);
    S.run_modes = reinterpret_cast<const RunMode*>(base + h.off_run_modes);
    S.floors0 = reinterpret_cast<const DevFloor0*>(base + h.off_floors0);
    S.f0_bark = reinterpret_cast<const int32_t*>(base + h.off_f0_bark);
    S.f0_wmap = reinterpret_cast<const float*>(base + h.off_f0_wmap);
    S.f0_stride = h.f0_stride; S.f0_max_order = h.f0_max_order;
    S.r2cand = reinterpret_cast<const uint32_t*>(base + h.off_r2cand);
    S.r2ob = reinterpret_cast<const uint16_t*>(base + h.off_r2ob);
    S.spectrum_bins = h.spectrum_bins; S.r2_max_p = h.r2_max_p;
    S.magic = reinterpret_cast<const uint32_t*>(base + h.off_magic);
    S.cls_cnt = reinterpret_cast<const unsigned long long*>(base + h.off_cls_cnt);
    S.wf_max_p = h.wf_max_p; S.max_posts = (h.post_stride - 2 > 2) ? h.post_stride - 2 : 2;
}

int build_blob(const nvb_setup* s, std::vector<unsigned char>& blob, std::string& err) {
    if (!s) return fail(err, NVB_ERR_ARG, "setup is NULL");
    if (s->abi_version != NVB_ABI_VERSION) return fail(err, NVB_ERR_ARG, "abi_version %lld, library is %lld", s->abi_version, NVB_ABI_VERSION);
    if (s->channels < 1) return fail(err, NVB_ERR_DATA, "channels %lld", s->channels);
    if (s->channels > NVB_MAX_CHANNELS) return fail(err, NVB_ERR_UNSUPPORTED, "channels %lld > %lld", s->channels, NVB_MAX_CHANNELS);
    for (int i = 0; i < 2; i++)
        if (!is_pow2(s->block_size[i]) || s->block_size[i] < 64 || s->block_size[i] > 8192)       // StreamDecoder.cs:192-193: 1 << 4 bits
            return fail(err, NVB_ERR_DATA, "block_size[%lld] = %lld", i, s->block_size[i]);
    if (s->block_size[0] > s->block_size[1]) return fail(err, NVB_ERR_DATA, "block_size[0] > block_size[1]");
    if (s->n_books < 1 || s->n_floors < 1 || s->n_residues < 1 || s->n_mappings < 1 || s->n_modes < 1 ||
        s->n_books > 256 || s->n_floors > 64 || s->n_residues > 64 || s->n_mappings > 64 || s->n_modes > 64)
        return fail(err, NVB_ERR_DATA, "table counts out of range");
    if (!s->books || !s->floors || !s->residues || !s->mappings || !s->modes || (s->n_vq_floats > 0 && !s->vq_floats) || s->n_vq_floats < 0)
        return fail(err, NVB_ERR_ARG, "NULL table pointer");
    const int C = s->channels;

    for (int i = 0; i < s->n_books; i++) {
        const nvb_codebook& b = s->books[i];
        if (b.dims < 0 || b.entries < 0) return fail(err, NVB_ERR_DATA, "book %lld: negative size", i);
        if (b.map_type != 0 && b.table_off >= 0) {
            if (b.dims < 1) return fail(err, NVB_ERR_DATA, "book %lld: dims < 1", i);
            if (b.table_off + (int64_t)b.entries * b.dims > s->n_vq_floats) return fail(err, NVB_ERR_DATA, "book %lld: table outside vq_floats", i);
        }
    }
    int max_posts = 2, max_order0 = 0;
    for (int i = 0; i < s->n_floors; i++) {
        const nvb_floor& f = s->floors[i];
        if (f.type == 0) {                                                                              // Floor0.cs:28-39
            const nvb_floor0& z = f.f0;
            if (z.order < 1 || z.order > 255 || z.rate < 1 || z.rate > 65535 || z.bark_map_size < 1 || z.bark_map_size > 65535 || z.amp_ofs < 0 || z.amp_ofs > 255)
                return fail(err, NVB_ERR_DATA, "floor %lld: type 0 header fields out of range", i);
            if (z.order > max_order0) max_order0 = z.order;
            continue;
        }
        if (f.type != 1) return fail(err, NVB_ERR_DATA, "floor %lld: invalid type %lld", i, f.type);       // Factory.cs:22-31
        const nvb_floor1& g = f.f1;
        if (g.n_posts < 2 || g.n_posts > NVB_MAX_POSTS) return fail(err, NVB_ERR_DATA, "floor %lld: n_posts %lld", i, g.n_posts);
        if (g.multiplier < 1 || g.multiplier > 4 || g.range < 1 || g.range > 256) return fail(err, NVB_ERR_DATA, "floor %lld: multiplier/range", i);
        if (g.x_list[0] != 0) return fail(err, NVB_ERR_DATA, "floor %lld: x_list[0] != 0", i);
        for (int k = 2; k < g.n_posts; k++) {
            const int lo = g.l_neigh[k], hi = g.h_neigh[k];
            if (lo >= k || hi >= k || !(g.x_list[lo] < g.x_list[k] && g.x_list[k] < g.x_list[hi]))
                return fail(err, NVB_ERR_DATA, "floor %lld: bad neighbours of post %lld", i, k);
        }
        unsigned long long seen = 0;
        for (int k = 0; k < g.n_posts; k++) {
            const int sx = g.sort_idx[k];
            if (sx >= g.n_posts || ((seen >> sx) & 1ull)) return fail(err, NVB_ERR_DATA, "floor %lld: sort_idx is not a permutation", i);
            seen |= 1ull << sx;
            if (k > 0 && !(g.x_list[g.sort_idx[k - 1]] < g.x_list[sx])) return fail(err, NVB_ERR_DATA, "floor %lld: sort_idx not ascending in x", i);
        }
        if (g.sort_idx[0] != 0) return fail(err, NVB_ERR_DATA, "floor %lld: sort_idx[0] != 0", i);
        if (g.n_posts > max_posts) max_posts = g.n_posts;
    }
    for (int i = 0; i < s->n_residues; i++) {
        const nvb_residue& r = s->residues[i];
        if (r.type < 0 || r.type > 2) return fail(err, NVB_ERR_DATA, "residue %lld: invalid type %lld", i, r.type);   // Factory.cs:48-58
        if (r.begin < 0 || r.end < 0 || r.partition_size < 1 || r.classifications < 1 || r.classifications > NVB_MAX_CLASSES ||
            r.max_stages < 0 || r.max_stages > NVB_MAX_STAGES)
            return fail(err, NVB_ERR_DATA, "residue %lld: header fields out of range", i);
        for (int c = 0; c < r.classifications; c++)
            for (int st = 0; st < r.max_stages; st++) {
                if (!((r.cascade[c] >> st) & 1)) continue;
                const int bk = r.books[c][st];
                if (bk < 0) continue;
                if (bk >= s->n_books) return fail(err, NVB_ERR_DATA, "residue %lld: book %lld out of range", i, bk);
                const nvb_codebook& b = s->books[bk];
                if (b.map_type == 0 || b.table_off < 0 || b.dims < 1) return fail(err, NVB_ERR_DATA, "residue %lld: book %lld has no lookup table", i, bk);   // Residue0.cs:66-67
                if (b.entries > 65536) return fail(err, NVB_ERR_UNSUPPORTED, "residue %lld: book %lld has more than 65536 entries", i, bk);
                if (r.type == 0 && r.partition_size / b.dims < 1) return fail(err, NVB_ERR_DATA, "residue %lld: type 0 book %lld wider than a partition", i, bk);
            }
    }
    for (int i = 0; i < s->n_mappings; i++) {
        const nvb_mapping& m = s->mappings[i];
        if (m.n_submaps != 1) return fail(err, NVB_ERR_UNSUPPORTED, "mapping %lld has %lld submaps (Mapping.cs:122-134 mis-decodes those)", i, m.n_submaps);
        if (m.n_coupling < 0 || m.n_coupling > 256) return fail(err, NVB_ERR_DATA, "mapping %lld: coupling steps", i);
        if (m.n_coupling > NVB_MAX_COUPLING) return fail(err, NVB_ERR_UNSUPPORTED, "mapping %lld: %lld coupling steps", i, m.n_coupling);
        for (int k = 0; k < m.n_coupling; k++)
            if (m.magnitude[k] == m.angle[k] || m.magnitude[k] >= C || m.angle[k] >= C)                  // Mapping.cs:36-41
                return fail(err, NVB_ERR_DATA, "mapping %lld: invalid coupling step %lld", i, k);
        if (m.floor < 0 || m.floor >= s->n_floors || m.residue < 0 || m.residue >= s->n_residues)
            return fail(err, NVB_ERR_DATA, "mapping %lld: floor/residue index", i);
    }
    for (int i = 0; i < s->n_modes; i++)
        if (s->modes[i].mapping < 0 || s->modes[i].mapping >= s->n_mappings) return fail(err, NVB_ERR_DATA, "mode %lld: mapping index", i);   // Mode.cs:35-38

    blob.clear();
    BlobWriter w(blob);
    w.reserve(sizeof(BlobHeader));
    BlobHeader h; std::memset(&h, 0, sizeof h);
    h.magic = BLOB_MAGIC; h.abi = NVB_ABI_VERSION;
    h.channels = C; h.sample_rate = s->sample_rate; h.bs[0] = s->block_size[0]; h.bs[1] = s->block_size[1];
    h.n_books = s->n_books; h.n_floors = s->n_floors; h.n_residues = s->n_residues; h.n_mappings = s->n_mappings; h.n_modes = s->n_modes;
    h.post_stride = (2 + max_posts + 1) & ~1;

    // VQ tables are re-packed so that every table starts on a 16-byte boundary (vector loads of whole VQ vectors)
    h.off_books = w.reserve(sizeof(DevBook) * s->n_books);
    std::vector<int64_t> new_off((size_t)s->n_books, -1);
    int64_t vq_total = 0;
    for (int i = 0; i < s->n_books; i++) {
        const nvb_codebook& b = s->books[i];
        if (b.map_type != 0 && b.table_off >= 0) { new_off[(size_t)i] = vq_total; vq_total += ((int64_t)b.entries * b.dims + 3) & ~int64_t(3); }
    }
    h.n_vq = (uint64_t)vq_total;
    for (int i = 0; i < s->n_books; i++) {
        DevBook d; d.dims = s->books[i].dims; d.entries = s->books[i].entries; d.pad = 0;
        d.off = new_off[(size_t)i];
        d.dshift = is_pow2(d.dims) ? ilog_u(d.dims) - 1 : -1;
        w.at<DevBook>(h.off_books)[i] = d;
    }
    h.off_vq = w.reserve(sizeof(float) * (size_t)(vq_total > 0 ? vq_total : 1));
    for (int i = 0; i < s->n_books; i++)
        if (new_off[(size_t)i] >= 0)
            std::memcpy(w.at<float>(h.off_vq) + new_off[(size_t)i], s->vq_floats + s->books[i].table_off, sizeof(float) * (size_t)s->books[i].entries * s->books[i].dims);
    h.off_floors = w.reserve(sizeof(DevFloor1) * s->n_floors);
    for (int i = 0; i < s->n_floors; i++) {
        const nvb_floor1& g = s->floors[i].f1;
        DevFloor1 d; std::memset(&d, 0, sizeof d);
        if (s->floors[i].type != 1) { w.at<DevFloor1>(h.off_floors)[i] = d; continue; }    // type 0: see DevFloor0 below
        d.n_posts = g.n_posts; d.mult = g.multiplier; d.range = g.range;
        for (int k = 0; k < g.n_posts; k++) { d.x[k] = g.x_list[k]; d.lo[k] = g.l_neigh[k]; d.hi[k] = g.h_neigh[k]; d.sort[k] = g.sort_idx[k]; }
        for (int k = 2; k < g.n_posts; k++) {                           // neighbours always precede the post (validated above)
            const int lv = 1 + (d.level[d.lo[k]] > d.level[d.hi[k]] ? d.level[d.lo[k]] : d.level[d.hi[k]]);
            d.level[k] = (uint8_t)lv;
            if (lv > d.max_level) d.max_level = lv;
            d.rcp[k] = 1.0f / (float)((int)d.x[d.hi[k]] - (int)d.x[d.lo[k]]);
            const uint32_t adx = (uint32_t)((int)d.x[d.hi[k]] - (int)d.x[d.lo[k]]);      // >= 2: x[lo] < x[k] < x[hi] (validated above)
            d.magic[k] = adx >= 2 ? 0xffffffffu / adx + 1u : 0u;
        }
        for (int k = 0; k < g.n_posts; k++) d.xs[k] = d.x[d.sort[k]];
        w.at<DevFloor1>(h.off_floors)[i] = d;
    }
    // bin -> sorted position of the last post at or below it (posts are strictly ascending in sort order, x[sort[0]] = 0)
    h.off_bin2k = w.reserve((size_t)s->n_floors * (h.bs[1] / 2));
    for (int i = 0; i < s->n_floors; i++) {
        const DevFloor1& d = w.at<DevFloor1>(h.off_floors)[i];
        uint8_t* tab = w.at<uint8_t>(h.off_bin2k) + (size_t)i * (h.bs[1] / 2);
        int k = 0;
        for (int bin = 0; bin < h.bs[1] / 2; bin++) {
            while (k + 1 < d.n_posts && (int)d.xs[k + 1] <= bin) ++k;
            tab[bin] = (uint8_t)k;
        }
    }
    // k_spectrum_wf: floor(q / adx) as umulhi(q, m) with m = floor(2^32 / adx) + 1 is exact while q * adx < 2^32; a segment
    // evaluates q = (x - x0) |dy| < adx |dy|, so |dy| <= lim = floor((2^32 - 1) / adx^2) suffices
    {
        const int nb = h.bs[1] / 2;
        h.off_magic = w.reserve(sizeof(uint32_t) * 2 * (size_t)(nb + 1));
        uint32_t* mg = w.at<uint32_t>(h.off_magic);
        for (int adx = 0; adx <= nb; adx++) {
            if (adx < 2) { mg[2 * adx] = 1u; mg[2 * adx + 1] = 0xffffffffu; continue; }     // one bin: (x - x0) = 0
            mg[2 * adx] = 0xffffffffu / (uint32_t)adx + 1u;
            mg[2 * adx + 1] = (uint32_t)(0xffffffffull / ((uint64_t)adx * (uint64_t)adx));
        }
    }
    // type 0 floors: the bark map and the 2 cos map of Floor0.Init (Floor0.cs:53-96) for both block sizes
    {
        h.f0_max_order = max_order0;
        h.f0_stride = max_order0 > 0 ? (1 + max_order0 + 1) & ~1 : 0;
        std::vector<int32_t> bark; std::vector<float> wmap;
        std::vector<DevFloor0> f0((size_t)s->n_floors);
        for (int i = 0; i < s->n_floors; i++) {
            DevFloor0& d = f0[(size_t)i]; std::memset(&d, 0, sizeof d);
            d.type = s->floors[i].type;
            if (d.type != 0) continue;
            const nvb_floor0& z = s->floors[i].f0;
            d.order = z.order; d.amp_ofs = z.amp_ofs; d.bark_map_size = z.bark_map_size;
            for (int j = 0; j < 2; j++) {
                const int n = h.bs[j] / 2;
                d.bark_off[j] = (int32_t)bark.size(); d.wmap_off[j] = (int32_t)wmap.size();
                // SynthesizeBarkCurve (Floor0.cs:67-79): the loop stops at n - 2, so map[n - 1] keeps its initial 0
                const float scale = (float)z.bark_map_size / floor0_to_bark((double)(z.rate / 2));
                int kmax = 0;
                for (int b = 0; b < n; b++) {
                    int v = 0;
                    if (b < n - 1) {
                        const float fr = (z.rate / 2.f) / n * b;
                        v = (int)std::floor((double)(floor0_to_bark((double)fr) * scale));
                        if (v > z.bark_map_size - 1) v = z.bark_map_size - 1;
                    }
                    if (v < 0 || v >= n) {
                        // wMap[k] with k >= n: IndexOutOfRangeException on every packet that applies this floor at this block size
                        // (Floor0.cs:166); refused if some mode does that, harmless otherwise (the reference builds the map anyway)
                        bool used = false;
                        for (int m = 0; m < s->n_modes; m++)
                            if ((s->modes[m].block_flag ? 1 : 0) == j && s->mappings[s->modes[m].mapping].floor == i) used = true;
                        if (used) return fail(err, NVB_ERR_UNSUPPORTED, "floor %lld: bark map indexes past the cos map (Floor0.cs:166 would throw)", i);
                        v = 0;
                    }
                    if (v > kmax) kmax = v;
                    bark.push_back(v);
                }
                d.kmax[j] = kmax;
                // SynthesizeWDelMap (Floor0.cs:86-96)
                const float wdel = (float)(3.14159265358979323846 / z.bark_map_size);
                for (int b = 0; b < n; b++) wmap.push_back(2.f * (float)std::cos((double)(wdel * b)));
            }
        }
        h.off_floors0 = w.reserve(sizeof(DevFloor0) * (size_t)s->n_floors);
        std::memcpy(w.at<DevFloor0>(h.off_floors0), f0.data(), sizeof(DevFloor0) * f0.size());
        h.n_f0_bark = bark.size(); h.n_f0_wmap = wmap.size();
        h.off_f0_bark = w.reserve(sizeof(int32_t) * (bark.size() ? bark.size() : 1));
        if (!bark.empty()) std::memcpy(w.at<int32_t>(h.off_f0_bark), bark.data(), sizeof(int32_t) * bark.size());
        h.off_f0_wmap = w.reserve(sizeof(float) * (wmap.size() ? wmap.size() : 1));
        if (!wmap.empty()) std::memcpy(w.at<float>(h.off_f0_wmap), wmap.data(), sizeof(float) * wmap.size());
    }
    h.off_residues = w.reserve(sizeof(DevResidue) * s->n_residues);
    int ci_total = 0;
    for (int i = 0; i < s->n_residues; i++) {
        const nvb_residue& r = s->residues[i];
        DevResidue d; std::memset(&d, 0, sizeof d);
        d.type = r.type; d.begin = r.begin; d.end = r.end; d.psize = r.partition_size; d.nclass = r.classifications; d.stages = r.max_stages;
        for (int c = 0; c < NVB_MAX_CLASSES; c++) {
            d.cascade[c] = c < r.classifications ? r.cascade[c] : 0;
            for (int st = 0; st < NVB_MAX_STAGES; st++) {
                const bool coded = c < r.classifications && st < r.max_stages && ((r.cascade[c] >> st) & 1) && r.books[c][st] >= 0;
                d.books[c][st] = coded ? r.books[c][st] : (int16_t)-1;
            }
        }
        const ResidueDerived dv = derive_residue(d.type, d.begin, d.psize, d.nclass, d.stages, d.cascade, d.books, C, [&](int bk) { return s->books[bk].dims; });
        d.pshift = dv.pshift; d.fast = dv.fast;
        std::memcpy(d.cnt, dv.cnt, sizeof d.cnt); std::memcpy(d.coded, dv.coded, sizeof d.coded);
        d.ci_off = ci_total; ci_total += r.classifications * (r.max_stages > 0 ? r.max_stages : 1);
        w.at<DevResidue>(h.off_residues)[i] = d;
    }
    h.off_ci = w.reserve(sizeof(CiRec) * (size_t)(ci_total > 0 ? ci_total : 1));
    for (int i = 0; i < s->n_residues; i++) {
        const DevResidue& d = w.at<DevResidue>(h.off_residues)[i];
        const int st_n = d.stages > 0 ? d.stages : 1;
        for (int c = 0; c < d.nclass; c++) for (int st = 0; st < st_n; st++) {
            CiRec ci; ci.off = 0; ci.dshift = 0; ci.entries = 0; ci.cnt = 0;
            const int bk = st < d.stages ? d.books[c][st] : -1;
            if (bk >= 0 && new_off[(size_t)bk] >= 0 && new_off[(size_t)bk] < (int64_t(1) << 30)) {
                const DevBook& b = w.at<DevBook>(h.off_books)[bk];
                ci.off = (int32_t)b.off; ci.dshift = b.dshift; ci.entries = b.entries; ci.cnt = d.cnt[c][st];
            }
            w.at<CiRec>(h.off_ci)[d.ci_off + c * st_n + st] = ci;
        }
    }
    // k_spectrum_wf: entries per partition of (class, stage) packed 16 bits per stage, four stages per 64-bit word, so that ONE
    // warp scan yields the entry-stream offsets of four stages (a stage holds at most span <= 32768 entries)
    std::vector<int> cc_off((size_t)s->n_residues, 0);
    {
        int total = 0;
        for (int i = 0; i < s->n_residues; i++) { const DevResidue& d = w.at<DevResidue>(h.off_residues)[i]; cc_off[(size_t)i] = total; total += ((d.stages + 3) / 4 > 0 ? (d.stages + 3) / 4 : 1) * d.nclass; }
        h.cls_cnt_total = total;
        h.off_cls_cnt = w.reserve(sizeof(uint64_t) * (size_t)(total > 0 ? total : 1));
        for (int i = 0; i < s->n_residues; i++) {
            const DevResidue& d = w.at<DevResidue>(h.off_residues)[i];
            const int nw = (d.stages + 3) / 4;
            for (int wd = 0; wd < nw; wd++) for (int c = 0; c < d.nclass; c++) {
                uint64_t v = 0;
                for (int k = 0; k < 4 && 4 * wd + k < d.stages; k++) v |= (uint64_t)(uint16_t)d.cnt[c][4 * wd + k] << (16 * k);
                w.at<uint64_t>(h.off_cls_cnt)[cc_off[(size_t)i] + wd * d.nclass + c] = v;
            }
        }
    }
    h.off_mappings = w.reserve(sizeof(DevMapping) * s->n_mappings);
    for (int i = 0; i < s->n_mappings; i++) {
        const nvb_mapping& m = s->mappings[i];
        DevMapping d; std::memset(&d, 0, sizeof d);
        d.n_coupling = m.n_coupling; d.floor = m.floor; d.residue = m.residue;
        for (int k = 0; k < m.n_coupling; k++) { d.mag[k] = m.magnitude[k]; d.ang[k] = m.angle[k]; }
        w.at<DevMapping>(h.off_mappings)[i] = d;
    }
    h.off_modes = w.reserve(sizeof(DevMode) * s->n_modes);
    for (int i = 0; i < s->n_modes; i++) {
        DevMode d; d.block_flag = s->modes[i].block_flag ? 1 : 0; d.mapping = s->modes[i].mapping;
        w.at<DevMode>(h.off_modes)[i] = d;
    }

    // k_spectrum_bins tables: for a type 2 residue, element e of partition p lands on channel e % C of bin ob(p) + e / C with
    // ob(p) = (begin + p * psize) / C (Residue2.cs:25-43: chPtr restarts per partition, offset /= channels), so a bin is
    // reached by the partitions p with ob(p) <= bin and (bin - ob(p)) * C < psize: consecutive p, at most a few
    std::vector<int> bins_ok((size_t)s->n_residues, 0);
    {
        const int nb = h.bs[1] / 2;
        int max_p = 1;
        for (int i = 0; i < s->n_residues; i++) {
            const nvb_residue& r = s->residues[i];
            if (r.type == 2 && r.partition_size > 0) { const int span = nb * C; const int e = r.end < span ? r.end : span; const int P = e > r.begin ? (e - r.begin) / r.partition_size : 0; if (P > max_p) max_p = P; }
        }
        h.r2_max_p = max_p;
        h.off_r2cand = w.reserve(sizeof(uint32_t) * (size_t)s->n_residues * nb);
        h.off_r2ob = w.reserve(sizeof(uint16_t) * (size_t)s->n_residues * max_p);
        for (int i = 0; i < s->n_residues; i++) {
            const nvb_residue& r = s->residues[i];
            const DevResidue d = w.at<DevResidue>(h.off_residues)[i];
            if (r.type != 2 || d.pshift < 0 || r.max_stages < 1 || C > NVB_FAST_CHANNELS) continue;
            bool ok = true;
            for (int c = 0; c < d.nclass && ok; c++) for (int st = 0; st < d.stages && ok; st++) {
                if (d.books[c][st] < 0) continue;
                const int dims = s->books[d.books[c][st]].dims;
                if (!is_pow2(dims) || dims > r.partition_size || d.cnt[c][st] > 32767) ok = false;
            }
            const int span = nb * C; const int e = r.end < span ? r.end : span; const int P = e > r.begin ? (e - r.begin) / r.partition_size : 0;
            uint16_t* ob = w.at<uint16_t>(h.off_r2ob) + (size_t)i * max_p;
            for (int p = 0; p < P; p++) ob[p] = (uint16_t)((r.begin + p * r.partition_size) / C);
            uint32_t* cand = w.at<uint32_t>(h.off_r2cand) + (size_t)i * nb;
            for (int bin = 0; bin < nb && ok; bin++) {
                int first = -1, cnt = 0;
                for (int p = 0; p < P; p++) {
                    const int o = (r.begin + p * r.partition_size) / C;
                    if (o <= bin && (bin - o) * C < r.partition_size) { if (first < 0) first = p; if (p != first + cnt) ok = false; ++cnt; }
                }
                if (cnt > 255) ok = false;
                cand[bin] = first < 0 ? 0u : ((uint32_t)first | ((uint32_t)cnt << 16));
            }
            bins_ok[(size_t)i] = ok ? 1 : 0;
        }
    }
    h.off_run_modes = w.reserve(sizeof(RunMode) * (size_t)s->n_modes);
    for (int i = 0; i < s->n_modes; i++) {
        RunMode rm; std::memset(&rm, 0, sizeof rm);
        const DevMode md = w.at<DevMode>(h.off_modes)[i];
        const DevMapping mp = w.at<DevMapping>(h.off_mappings)[md.mapping];
        const DevResidue& R = w.at<DevResidue>(h.off_residues)[mp.residue];
        rm.rbegin = R.begin; rm.rend = R.end; rm.pshift = R.pshift; rm.stages = R.stages; rm.nclass = R.nclass; rm.ci_off = R.ci_off;
        rm.residue = mp.residue; rm.floor = mp.floor; rm.n_coupling = mp.n_coupling; rm.mapping = md.mapping; rm.block_flag = md.block_flag; rm.rtype = R.type;
        rm.cand_off = mp.residue * (h.bs[1] / 2); rm.ob_off = mp.residue * h.r2_max_p;
        rm.bins_ok = bins_ok[(size_t)mp.residue] && s->floors[mp.floor].type == 1;
        rm.cc_off = cc_off[(size_t)mp.residue]; rm.base_stride = ((R.stages + 3) & ~3) > 0 ? ((R.stages + 3) & ~3) : 4;
        rm.n_posts = s->floors[mp.floor].type == 1 ? s->floors[mp.floor].f1.n_posts : 0;
        {   // partitions of this mode's residue over its block size (single-stream residues only matter here)
            const int span = (R.type == 2 ? C : 1) * (h.bs[md.block_flag] / 2); const int e = R.end < span ? R.end : span;
            const int P = (e > R.begin && R.psize > 0) ? (e - R.begin) / R.psize : 0;
            if (P > h.wf_max_p) h.wf_max_p = P;
        }
        w.at<RunMode>(h.off_run_modes)[i] = rm;
    }
    h.spectrum_bins = 1;
    for (int i = 0; i < s->n_modes; i++) if (!w.at<RunMode>(h.off_run_modes)[i].bins_ok) h.spectrum_bins = 0;

    // windows (Mode.cs:24-67): one for the short size, four for the long size
    {
        std::vector<float> slope[2];
        for (int i = 0; i < 2; i++) {
            slope[i].resize((size_t)h.bs[i] / 2);
            if (s->window_slope[i]) std::memcpy(slope[i].data(), s->window_slope[i], sizeof(float) * slope[i].size());
            else window_slope(h.bs[i] / 2, slope[i].data());
        }
        h.off_win_short = w.reserve(sizeof(float) * h.bs[0]);
        assemble_window(slope[0].data(), h.bs[0] / 2, slope[0].data(), h.bs[0] / 2, h.bs[0], w.at<float>(h.off_win_short));
        h.off_win_long = w.reserve(sizeof(float) * 4 * (size_t)h.bs[1]);
        for (int wi = 0; wi < 4; wi++) {
            const int l = (wi & 1) ? 1 : 0, r = (wi & 2) ? 1 : 0;            // Mode.cs:44-50
            assemble_window(slope[l].data(), h.bs[l] / 2, slope[r].data(), h.bs[r] / 2, h.bs[1], w.at<float>(h.off_win_long) + (size_t)wi * h.bs[1]);
        }
    }
    for (int i = 0; i < 2; i++) {
        const int n = h.bs[i];
        h.off_mdct_a[i] = w.reserve(sizeof(float) * (n / 2));
        h.off_mdct_b[i] = w.reserve(sizeof(float) * (n / 2));
        h.off_mdct_c[i] = w.reserve(sizeof(float) * (n / 4));
        h.off_bitrev[i] = w.reserve(sizeof(uint16_t) * (n / 8));
        h.off_tw[i] = w.reserve(sizeof(float2) * (n / 4));
        h.off_fft[i] = w.reserve(sizeof(float2) * (n / 4));
        mdct_tables(n, w.at<float>(h.off_mdct_a[i]), w.at<float>(h.off_mdct_b[i]), w.at<float>(h.off_mdct_c[i]), w.at<uint16_t>(h.off_bitrev[i]));
        if (s->mdct_a[i]) std::memcpy(w.at<float>(h.off_mdct_a[i]), s->mdct_a[i], sizeof(float) * (n / 2));
        if (s->mdct_b[i]) std::memcpy(w.at<float>(h.off_mdct_b[i]), s->mdct_b[i], sizeof(float) * (n / 2));
        if (s->mdct_c[i]) std::memcpy(w.at<float>(h.off_mdct_c[i]), s->mdct_c[i], sizeof(float) * (n / 4));
        if (s->mdct_bitrev[i]) std::memcpy(w.at<uint16_t>(h.off_bitrev[i]), s->mdct_bitrev[i], sizeof(uint16_t) * (n / 8));
        fast_tables(n, w.at<float2>(h.off_tw[i]), w.at<float2>(h.off_fft[i]));
    }
    h.off_db = w.reserve(sizeof(float) * 256);
    std::memcpy(w.at<float>(h.off_db), k_inverse_db_bits, sizeof k_inverse_db_bits);
    if (h.bs[0] == FUSED_SHORT_N && h.bs[1] == FUSED_LONG_N) {
        h.off_fused_tab = w.reserve(sizeof(float) * FusedTables::FLOATS);
        // rising slope of window 3 (long block between long blocks): the first 1024 values of that window
        build_fused_tables(w.at<float2>(h.off_tw[0]), w.at<float2>(h.off_fft[0]), w.at<float>(h.off_win_long) + 3 * (size_t)h.bs[1], w.at<float>(h.off_win_short),
                           w.at<float>(h.off_fused_tab));
    }

    // largest residue item table over the modes (k_spectrum's shared-memory prefix array)
    {
        DevSetup S; resolve_setup(blob.data(), h, S);
        int mx = 1;
        for (int i = 0; i < h.n_modes; i++) {
            const DevMode& md = S.modes[i];
            const DevResidue& R = S.residues[S.mappings[md.mapping].residue];
            ResGeom g = residue_geom(R, h.bs[md.block_flag], C);
            if (g.n_items > mx) mx = g.n_items;
        }
        if (mx > 40000) return fail(err, NVB_ERR_UNSUPPORTED, "residue layout needs %lld prefix items (> 40000)", mx);
        h.max_items = mx;
        h.ci_total = ci_total;
        h.max_stages = 1;
        for (int i = 0; i < h.n_residues; i++) if (S.residues[i].stages > h.max_stages) h.max_stages = S.residues[i].stages;
        int fast = 3;
        for (int i = 0; i < h.n_modes; i++) {
            const DevResidue& R = S.residues[S.mappings[S.modes[i].mapping].residue];
            if (R.fast < fast) fast = R.fast;
        }
        // plane kernel: stages planes + floor rows + item list in shared memory (k_spectrum_run needs none of those)
        if (fast == 2 && (size_t)(h.max_stages + 1) * C * (h.bs[1] / 2) * 4 + (size_t)mx * 9 + 64 > 96 * 1024) fast = 1;
        if (fast == 3 && (size_t)mx * 5 + (size_t)ci_total * sizeof(CiRec) + 64 > 160 * 1024) fast = 1;
        for (int i = 0; i < h.n_modes; i++)                                  // Floor0 lives in the general kernel only
            if (s->floors[S.mappings[S.modes[i].mapping].floor].type == 0) fast = 0;
        // general kernel: prefix table + (type 0 floors) per-channel curve and coefficient tables in shared memory
        if (h.f0_stride > 0 && (size_t)mx * 4 + (size_t)C * (h.bs[1] / 2 + 256) * 4 + 64 > 200 * 1024)
            return fail(err, NVB_ERR_UNSUPPORTED, "type 0 floor tables do not fit in shared memory");
        if ((size_t)mx * 4 + (size_t)C * (h.bs[1] / 2) * 4 > 160 * 1024) fast = 0;      // prefix table + floor curve rows must fit in shared memory
        h.spectrum_fast = fast;
    }
    w.reserve(0);
    h.total_bytes = blob.size();
    h.body_hash = blob_body_hash(blob.data(), blob.size());
    std::memcpy(blob.data(), &h, sizeof h);
    return NVB_OK;
}

int validate_blob(const void* data, size_t bytes, std::string& err) {
    if (!data || bytes < sizeof(BlobHeader)) return fail(err, NVB_ERR_ARG, "blob too small");
    BlobHeader h; std::memcpy(&h, data, sizeof h);
    if (h.magic != BLOB_MAGIC || h.abi != NVB_ABI_VERSION) return fail(err, NVB_ERR_DATA, "blob magic/abi mismatch");
    if (h.total_bytes != bytes) return fail(err, NVB_ERR_DATA, "blob size %lld, header says %lld", (long long)bytes, (long long)h.total_bytes);
    if (h.channels < 1 || h.channels > NVB_MAX_CHANNELS || !is_pow2(h.bs[0]) || !is_pow2(h.bs[1]) || h.bs[0] < 64 || h.bs[1] > 8192 || h.bs[0] > h.bs[1])
        return fail(err, NVB_ERR_DATA, "blob header fields out of range");
    // len and off are untrusted 64-bit values: compare without forming off + len (which wraps)
    auto in = [&](uint64_t off, uint64_t len) { return off >= sizeof(BlobHeader) && len <= bytes && off <= bytes - len && (off & 15) == 0; };
    if (h.n_books < 0 || h.n_floors < 0 || h.n_residues < 0 || h.n_mappings < 0 || h.n_modes < 0 || h.n_books > 65536 || h.n_floors > 256 || h.n_residues > 256 ||
        h.n_mappings > 256 || h.n_modes > 256 || h.n_vq > bytes / 4 || h.n_f0_bark > bytes / 4 || h.n_f0_wmap > bytes / 4)
        return fail(err, NVB_ERR_DATA, "blob header counts out of range");
    if (h.body_hash != blob_body_hash(static_cast<const unsigned char*>(data), bytes)) return fail(err, NVB_ERR_DATA, "blob checksum mismatch");
    bool ok = in(h.off_books, sizeof(DevBook) * (uint64_t)h.n_books) && in(h.off_vq, sizeof(float) * h.n_vq) &&
              in(h.off_floors, sizeof(DevFloor1) * (uint64_t)h.n_floors) && in(h.off_residues, sizeof(DevResidue) * (uint64_t)h.n_residues) &&
              in(h.off_mappings, sizeof(DevMapping) * (uint64_t)h.n_mappings) && in(h.off_modes, sizeof(DevMode) * (uint64_t)h.n_modes) &&
              in(h.off_win_short, 4ull * h.bs[0]) && in(h.off_win_long, 16ull * h.bs[1]) && in(h.off_db, 1024) &&
              h.ci_total >= 0 && in(h.off_ci, sizeof(CiRec) * (uint64_t)(h.ci_total > 0 ? h.ci_total : 1)) && in(h.off_bin2k, (uint64_t)h.n_floors * (h.bs[1] / 2)) && in(h.off_run_modes, sizeof(RunMode) * (uint64_t)h.n_modes) && in(h.off_floors0, sizeof(DevFloor0) * (uint64_t)h.n_floors) &&
              in(h.off_f0_bark, 4 * (h.n_f0_bark ? h.n_f0_bark : 1)) && h.r2_max_p >= 1 && h.r2_max_p <= 65536 &&
              in(h.off_r2cand, 4ull * h.n_residues * (h.bs[1] / 2)) && in(h.off_r2ob, 2ull * h.n_residues * h.r2_max_p) && in(h.off_f0_wmap, 4 * (h.n_f0_wmap ? h.n_f0_wmap : 1)) && h.f0_stride >= 0 && h.f0_stride <= 258 &&
              (h.off_fused_tab == 0 || (in(h.off_fused_tab, 4ull * FusedTables::FLOATS) && h.bs[0] == FUSED_SHORT_N && h.bs[1] == FUSED_LONG_N));
    ok = ok && in(h.off_magic, 8ull * (h.bs[1] / 2 + 1)) && h.cls_cnt_total >= 0 && h.cls_cnt_total <= 256 * NVB_MAX_CLASSES * 2 &&
         in(h.off_cls_cnt, 8ull * (uint64_t)(h.cls_cnt_total > 0 ? h.cls_cnt_total : 1)) && h.wf_max_p >= 0 && h.wf_max_p <= 65536;
    for (int i = 0; i < 2 && ok; i++)
        ok = in(h.off_mdct_a[i], 2ull * h.bs[i]) && in(h.off_mdct_b[i], 2ull * h.bs[i]) && in(h.off_mdct_c[i], 1ull * h.bs[i]) &&
             in(h.off_bitrev[i], h.bs[i] / 4ull) && in(h.off_tw[i], 2ull * h.bs[i]) && in(h.off_fft[i], 2ull * h.bs[i]);
    if (!ok) return fail(err, NVB_ERR_DATA, "blob section outside the blob");
    // index ranges the kernels rely on
    const unsigned char* base = static_cast<const unsigned char*>(data);
    DevSetup S; resolve_setup(base, h, S);
    for (int i = 0; i < h.n_modes; i++) if (S.modes[i].mapping < 0 || S.modes[i].mapping >= h.n_mappings) return fail(err, NVB_ERR_DATA, "blob: mode %lld", i);
    for (int i = 0; i < h.n_mappings; i++) {
        const DevMapping& m = S.mappings[i];
        if (m.floor < 0 || m.floor >= h.n_floors || m.residue < 0 || m.residue >= h.n_residues || m.n_coupling < 0 || m.n_coupling > NVB_MAX_COUPLING)
            return fail(err, NVB_ERR_DATA, "blob: mapping %lld", i);
        for (int k = 0; k < m.n_coupling; k++) if (m.mag[k] >= h.channels || m.ang[k] >= h.channels) return fail(err, NVB_ERR_DATA, "blob: mapping %lld coupling", i);
    }
    for (int i = 0; i < h.n_residues; i++) {
        const DevResidue& r = S.residues[i];
        if (r.type < 0 || r.type > 2 || r.psize < 1 || r.nclass < 1 || r.nclass > NVB_MAX_CLASSES || r.stages < 0 || r.stages > NVB_MAX_STAGES || r.begin < 0)
            return fail(err, NVB_ERR_DATA, "blob: residue %lld", i);
        for (int c = 0; c < r.nclass; c++) for (int st = 0; st < r.stages; st++) {
            const int bk = r.books[c][st];
            if (bk < 0) continue;
            if (bk >= h.n_books) return fail(err, NVB_ERR_DATA, "blob: residue %lld book", i);
            const DevBook& b = S.books[bk];
            if (b.dims < 1 || b.off < 0 || b.off + (int64_t)b.entries * b.dims > (int64_t)h.n_vq) return fail(err, NVB_ERR_DATA, "blob: residue %lld book table", i);
            if (r.cnt[c][st] < 0 || (r.fast && (b.dshift < 0 || (1 << b.dshift) != b.dims || r.pshift < 0 || (1 << r.pshift) != r.psize)))
                return fail(err, NVB_ERR_DATA, "blob: residue %lld fast-path fields", i);
        }
        // the derived fields select kernels with alignment / power-of-two assumptions: they must be what build_blob derives
        for (int c = 0; c < NVB_MAX_CLASSES; c++) for (int st = 0; st < NVB_MAX_STAGES; st++)
            if (r.books[c][st] >= h.n_books || (r.books[c][st] >= 0 && (c >= r.nclass || st >= r.stages || !((r.cascade[c] >> st) & 1)))) return fail(err, NVB_ERR_DATA, "blob: residue %lld book table", i);
        const ResidueDerived dv = derive_residue(r.type, r.begin, r.psize, r.nclass, r.stages, r.cascade, r.books, h.channels, [&](int bk) { return S.books[bk].dims; });
        if (dv.pshift != r.pshift || dv.fast != r.fast || std::memcmp(dv.cnt, r.cnt, sizeof dv.cnt) != 0 || std::memcmp(dv.coded, r.coded, sizeof dv.coded) != 0)
            return fail(err, NVB_ERR_DATA, "blob: residue %lld derived fields", i);
    }
    for (int i = 0; i < h.n_books; i++) {
        const DevBook& b = S.books[i];
        if (b.dshift != (is_pow2(b.dims) ? ilog_u(b.dims) - 1 : -1)) return fail(err, NVB_ERR_DATA, "blob: book %lld dshift", i);
    }
    for (int i = 0; i < h.n_floors; i++) {
        const DevFloor0& z = S.floors0[i];
        if (z.type == 0) {
            if (z.order < 1 || z.order > 255 || h.f0_stride < z.order + 1) return fail(err, NVB_ERR_DATA, "blob: floor %lld (type 0)", i);
            for (int j = 0; j < 2; j++) {
                const int n = h.bs[j] / 2;
                if (z.bark_off[j] < 0 || z.wmap_off[j] < 0 || (uint64_t)z.bark_off[j] + n > h.n_f0_bark || (uint64_t)z.wmap_off[j] + n > h.n_f0_wmap || z.kmax[j] < 0 || z.kmax[j] >= n)
                    return fail(err, NVB_ERR_DATA, "blob: floor %lld (type 0) tables", i);
                for (int b = 0; b < n; b++) { const int k = S.f0_bark[z.bark_off[j] + b]; if (k < 0 || k > z.kmax[j]) return fail(err, NVB_ERR_DATA, "blob: floor %lld bark map", i); }
            }
            continue;
        }
        if (z.type != 1) return fail(err, NVB_ERR_DATA, "blob: floor %lld type", i);
        const DevFloor1& f = S.floors[i];
        if (f.n_posts < 2 || f.n_posts > NVB_MAX_POSTS || f.max_level < 0 || f.max_level > NVB_MAX_POSTS) return fail(err, NVB_ERR_DATA, "blob: floor %lld", i);
        for (int k = 2; k < f.n_posts; k++) if (f.level[k] < 1 || f.level[k] > f.max_level) return fail(err, NVB_ERR_DATA, "blob: floor %lld levels", i);
        for (int k = 0; k < f.n_posts; k++) {
            if (f.sort[k] >= f.n_posts) return fail(err, NVB_ERR_DATA, "blob: floor %lld sort", i);
            if (k >= 2 && (f.lo[k] >= k || f.hi[k] >= k || !(f.x[f.lo[k]] < f.x[k] && f.x[k] < f.x[f.hi[k]]))) return fail(err, NVB_ERR_DATA, "blob: floor %lld neighbours", i);
            if (k > 0 && !(f.x[f.sort[k - 1]] < f.x[f.sort[k]])) return fail(err, NVB_ERR_DATA, "blob: floor %lld order", i);
        }
    }
    if (h.post_stride < 4 || h.max_items < 1 || h.max_items > 40000 || h.max_stages < 1 || h.max_stages > NVB_MAX_STAGES) return fail(err, NVB_ERR_DATA, "blob: post_stride/max_items/max_stages");
    for (int i = 0; i < h.n_residues; i++) if (S.residues[i].stages > h.max_stages) return fail(err, NVB_ERR_DATA, "blob: max_stages");
    if (h.spectrum_fast == 2 && (size_t)(h.max_stages + 1) * h.channels * (h.bs[1] / 2) * 4 + (size_t)h.max_items * 9 + 64 > 96 * 1024) return fail(err, NVB_ERR_DATA, "blob: plane kernel does not fit");
    if (h.spectrum_fast < 0 || h.spectrum_fast > 3) return fail(err, NVB_ERR_DATA, "blob: spectrum_fast");
    for (int i = 0; i < h.n_residues; i++) {
        const DevResidue& r = S.residues[i];
        const int st_n = r.stages > 0 ? r.stages : 1;
        if (r.ci_off < 0 || r.ci_off + r.nclass * st_n > h.ci_total) return fail(err, NVB_ERR_DATA, "blob: residue %lld ci table", i);
        for (int k = 0; k < r.nclass * st_n; k++) {
            const CiRec& ci = S.ci[r.ci_off + k];
            if (ci.cnt < 0 || ci.entries < 0 || ci.off < 0 || ci.dshift < -1 || ci.dshift > 16 ||
                (ci.cnt > 0 && ci.dshift >= 0 && (uint64_t)ci.off + ((uint64_t)ci.entries << ci.dshift) > h.n_vq)) return fail(err, NVB_ERR_DATA, "blob: residue %lld ci record", i);
            // a record either repeats its book (offset, log2 dims, entries) and the residue's entry count, or codes nothing
            const int c = k / st_n, st = k - c * st_n;
            const int bk = st < r.stages ? r.books[c][st] : -1;
            if (ci.cnt != 0) {
                if (bk < 0 || ci.cnt != r.cnt[c][st] || ci.off != S.books[bk].off || ci.dshift != S.books[bk].dshift || ci.entries != S.books[bk].entries)
                    return fail(err, NVB_ERR_DATA, "blob: residue %lld ci record does not match its book", i);
            } else if (bk >= 0 && r.cnt[c][st] != 0 && S.books[bk].off >= 0 && S.books[bk].off < (int64_t(1) << 30)) return fail(err, NVB_ERR_DATA, "blob: residue %lld ci record missing", i);
        }
    }
    {   // the kernel-selection levels of the header: never above what the modes' residues / floors allow
        int fast = 3; bool bins = true;
        for (int i = 0; i < h.n_modes; i++) {
            const DevMapping& mp = S.mappings[S.modes[i].mapping];
            if (S.residues[mp.residue].fast < fast) fast = S.residues[mp.residue].fast;
            if (S.floors0[mp.floor].type == 0) fast = 0;
            if (!S.run_modes[i].bins_ok) bins = false;
        }
        if (h.spectrum_fast > fast || (h.spectrum_bins != 0 && !bins) || (h.spectrum_bins != 0 && h.spectrum_bins != 1)) return fail(err, NVB_ERR_DATA, "blob: kernel selection flags");
        if (h.spectrum_fast == 3 && (size_t)h.max_items * 5 + (size_t)h.ci_total * sizeof(CiRec) + 64 > 160 * 1024) return fail(err, NVB_ERR_DATA, "blob: run kernel does not fit");
    }
    for (int i = 0; i < h.n_modes; i++) {
        const RunMode& rm = S.run_modes[i];
        const DevMapping& mp = S.mappings[S.modes[i].mapping];
        const DevResidue& R = S.residues[mp.residue];
        if (rm.mapping != S.modes[i].mapping || rm.residue != mp.residue || rm.floor != mp.floor || rm.n_coupling != mp.n_coupling || rm.rbegin != R.begin || rm.rend != R.end ||
            rm.pshift != R.pshift || rm.stages != R.stages || rm.nclass != R.nclass || rm.ci_off != R.ci_off || rm.rtype != R.type ||
            rm.cand_off != mp.residue * (h.bs[1] / 2) || rm.ob_off != mp.residue * h.r2_max_p) return fail(err, NVB_ERR_DATA, "blob: run mode %lld", i);
        {   // k_spectrum_wf's view of the mode
            const int nw = (R.stages + 3) / 4 > 0 ? (R.stages + 3) / 4 : 1;
            if (rm.base_stride != (((R.stages + 3) & ~3) > 0 ? ((R.stages + 3) & ~3) : 4) || rm.cc_off < 0 || rm.cc_off + nw * R.nclass > h.cls_cnt_total ||
                rm.n_posts != (S.floors0[mp.floor].type == 1 ? S.floors[mp.floor].n_posts : 0)) return fail(err, NVB_ERR_DATA, "blob: run mode %lld (wf fields)", i);
            for (int wd = 0; wd < (R.stages + 3) / 4; wd++) for (int c = 0; c < R.nclass; c++) {
                uint64_t v = 0;
                for (int k = 0; k < 4 && 4 * wd + k < R.stages; k++) v |= (uint64_t)(uint16_t)R.cnt[c][4 * wd + k] << (16 * k);
                if (S.cls_cnt[rm.cc_off + wd * R.nclass + c] != v) return fail(err, NVB_ERR_DATA, "blob: run mode %lld packed entry counts", i);
            }
            const int span = (R.type == 2 ? h.channels : 1) * (h.bs[S.modes[i].block_flag] / 2); const int e = R.end < span ? R.end : span;
            const int P = (e > R.begin && R.psize > 0) ? (e - R.begin) / R.psize : 0;
            if (P > h.wf_max_p) return fail(err, NVB_ERR_DATA, "blob: run mode %lld partition count", i);
        }
        if (rm.bins_ok) {
            if (R.type != 2 || R.pshift < 0 || R.stages < 1 || S.floors0[mp.floor].type != 1 || h.channels > NVB_FAST_CHANNELS) return fail(err, NVB_ERR_DATA, "blob: run mode %lld bins flag", i);
            for (int c = 0; c < R.nclass; c++) for (int st = 0; st < R.stages; st++) {
                if (R.books[c][st] < 0) continue;
                const int dims = S.books[R.books[c][st]].dims;
                if (!is_pow2(dims) || dims > R.psize) return fail(err, NVB_ERR_DATA, "blob: run mode %lld bins flag (book sizes)", i);
            }
            const int nb = h.bs[1] / 2; const int span = nb * h.channels; const int e = R.end < span ? R.end : span; const int P = e > R.begin ? (e - R.begin) >> R.pshift : 0;
            if (P > h.r2_max_p) return fail(err, NVB_ERR_DATA, "blob: run mode %lld partitions", i);
            for (int b = 0; b < nb; b++) { const uint32_t cd = S.r2cand[rm.cand_off + b]; if ((int)(cd & 0xffff) + (int)(cd >> 16) > P) return fail(err, NVB_ERR_DATA, "blob: run mode %lld candidate table", i); }
        }
    }
    for (int i = 0; i < h.n_floors; i++) {
        const DevFloor1& f = S.floors[i];
        if (S.floors0[i].type != 1) continue;
        for (int k = 0; k < f.n_posts; k++) if (f.xs[k] != f.x[f.sort[k]]) return fail(err, NVB_ERR_DATA, "blob: floor %lld sorted x", i);
        for (int k = 2; k < f.n_posts; k++) if (f.magic[k] != 0xffffffffu / (uint32_t)((int)f.x[f.hi[k]] - (int)f.x[f.lo[k]]) + 1u) return fail(err, NVB_ERR_DATA, "blob: floor %lld multipliers", i);
        for (int b = 0; b < h.bs[1] / 2; b++) {
            const int k = S.bin2k[(size_t)i * (h.bs[1] / 2) + b];
            if (k >= f.n_posts || f.xs[k] > b || (k + 1 < f.n_posts && f.xs[k + 1] <= b)) return fail(err, NVB_ERR_DATA, "blob: floor %lld bin table", i);
        }
    }
    for (int adx = 2; adx <= h.bs[1] / 2; adx++)
        if (S.magic[2 * adx] != 0xffffffffu / (uint32_t)adx + 1u || S.magic[2 * adx + 1] != (uint32_t)(0xffffffffull / ((uint64_t)adx * (uint64_t)adx)))
            return fail(err, NVB_ERR_DATA, "blob: multiplier table");
    if (S.magic[0] != 1u || S.magic[2] != 1u) return fail(err, NVB_ERR_DATA, "blob: multiplier table");
    return NVB_OK;
}

int plan_batch(const unsigned char* host_blob, const nvb_batch* b, int flags, const CarryState& in, Plan& out, std::string& err) {
    BlobHeader h; std::memcpy(&h, host_blob, sizeof h);
    DevSetup S; resolve_setup(host_blob, h, S);
    if (!b || b->n_frames < 0) return fail(err, NVB_ERR_ARG, "batch is NULL or n_frames < 0");
    if (b->n_frames > 0 && (!b->frames || !b->posts)) return fail(err, NVB_ERR_ARG, "batch.frames / batch.posts is NULL");
    if (b->n_classes < 0 || b->n_entries < 0 || (b->n_classes > 0 && !b->classes) || (b->n_entries > 0 && !b->entries))
        return fail(err, NVB_ERR_ARG, "batch.classes / batch.entries");
    const int C = h.channels;
    out = Plan();
    out.frames.reserve((size_t)b->n_frames);
    CarryState st = (flags & NVB_RUN_CONTINUE) ? in : CarryState();
    int last_ok = -1;
    int64_t pcm = 0, spec = 0;
    int64_t end_c = 0, end_e = 0;

    for (int i = 0; i < b->n_frames; i++) {
        const nvb_frame& f = b->frames[i];
        if (f.status != NVB_FRAME_OK) {
            // DecodeNextPacket returned null: the rest of the previous block is handed out as it is
            // (StreamDecoder.cs:352-356: _prevPacketEnd = _prevPacketStop)
            ++out.n_failed;
            st.prev_end = st.prev_stop;
            if (st.have_prev && st.prev_end > st.prev_start) {
                DevFrame d; std::memset(&d, 0, sizeof d);
                d.kind = 1; d.n = st.prev_n;
                d.out_begin = st.prev_start; d.out_end = st.prev_end;
                d.prev = last_ok >= 0 ? last_ok : PREV_CARRY;
                if (d.prev == PREV_CARRY) out.uses_carry = true;
                d.api_index = i; d.pcm_off = pcm;
                pcm += d.out_end - d.out_begin;
                out.frames.push_back(d);
            }
            st.prev_start = st.prev_end;
            continue;
        }
        if (f.mode >= h.n_modes) return fail(err, NVB_ERR_DATA, "frame %lld: mode %lld out of range", i, f.mode);   // StreamDecoder.cs:497
        const DevMode& md = S.modes[f.mode];
        const int n = h.bs[md.block_flag];
        if (f.window > 3 || (!md.block_flag && f.window != 0)) return fail(err, NVB_ERR_DATA, "frame %lld: window %lld", i, f.window);
        if (f.start < 0 || f.total > n || f.start > f.total || f.valid > f.total) return fail(err, NVB_ERR_DATA, "frame %lld: start/valid/total outside the block", i);
        if (((uint64_t)f.exec_mask >> C) != 0) return fail(err, NVB_ERR_DATA, "frame %lld: exec_mask has bits above the channel count", i);
        const DevMapping& mp = S.mappings[md.mapping];
        if (S.floors0[mp.floor].type == 0) {
            if (!b->floor0) return fail(err, NVB_ERR_ARG, "frame %lld uses a type 0 floor but batch.floor0 is NULL", i);
            out.uses_floor0 = true;
        }
        if (f.res_decoded) {
            const ResGeom g = residue_geom(S.residues[mp.residue], n, C);
            if ((int64_t)f.classes_off + (int64_t)g.P * g.Sx > b->n_classes) return fail(err, NVB_ERR_DATA, "frame %lld: classes outside the batch", i);
            if ((int64_t)f.entries_off + (int64_t)f.entry_count > b->n_entries) return fail(err, NVB_ERR_DATA, "frame %lld: entries outside the batch", i);
            if ((int64_t)f.classes_off < end_c || (int64_t)f.entries_off < end_e) out.sequential = false;
            end_c = (int64_t)f.classes_off + (int64_t)g.P * g.Sx; end_e = (int64_t)f.entries_off + (int64_t)f.entry_count;
        }
        const Overlap nom = nominal_overlap(h, md.block_flag, f.window);

        DevFrame d; std::memset(&d, 0, sizeof d);
        d.kind = 0; d.mode = f.mode; d.window = f.window; d.res_decoded = f.res_decoded ? 1 : 0;
        d.exec_mask = f.exec_mask; d.n = n; d.start = f.start; d.api_index = i;
        d.classes_off = f.classes_off; d.entries_off = f.entries_off; d.entry_count = f.res_decoded ? f.entry_count : 0;
        d.prev = PREV_NONE; d.ola_len = 0; d.prev_valid = 0;

        if (st.have_prev && st.prev_end > 0) {                               // StreamDecoder.cs:440-445
            int ola = st.prev_stop - st.prev_start;
            if (ola < 0) ola = 0;
            // The reference adds the whole previous tail at [start, start+ola) even when it reaches past
            // this block's own overlap region (into its tail, or past the block).  Well-formed streams
            // never do that; clamp and count.
            int room = nom.valid - f.start; if (room < 0) room = 0;
            if (ola > room) { ola = room; ++out.n_inconsistent; }
            if (ola > 0) {
                d.ola_len = ola; d.prev_valid = st.prev_start;
                d.prev = last_ok >= 0 ? last_ok : PREV_CARRY;
                if (d.prev == PREV_CARRY) out.uses_carry = true;
            }
            st.prev_start = f.start;
        } else if (!st.have_prev) {
            st.prev_start = f.valid;                                         // StreamDecoder.cs:446-450: first block emits nothing
        }
        st.prev_end = f.valid; st.prev_stop = f.total; st.have_prev = true; st.prev_n = n;   // StreamDecoder.cs:455-461
        if (st.prev_end < st.prev_start) {
            // EOS trim below the start index (StreamDecoder.cs:429-437 can do that): the reference's Read
            // loop would never terminate; emit nothing.
            ++out.n_inconsistent; st.prev_start = st.prev_end;
        }
        d.out_begin = st.prev_start; d.out_end = st.prev_end;
        if (d.out_begin < 0) { d.out_begin = 0; if (d.out_end < 0) d.out_end = 0; }
        d.pcm_off = pcm; pcm += d.out_end - d.out_begin;
        st.prev_start = st.prev_end;
        d.spec_off = (uint32_t)spec; spec += (int64_t)C * (n / 2);
        if (spec > 0x7fffffffLL / 2) return fail(err, NVB_ERR_CAPACITY, "batch too large: %lld spectrum floats", spec);
        last_ok = (int)out.frames.size();
        out.frames.push_back(d);
    }
    out.samples = pcm; out.spec_floats = spec; out.last_ok = last_ok; out.end_state = st;
    return NVB_OK;
}

}  // namespace nvb
