// nvb_internal.h -- device-side table layout shared by the kernels and the C-ABI host code.
// Not part of the public ABI (include/nvorbis_b200.h is).
#pragma once
#include <cstdint>
#include <cstddef>
#include <atomic>
#include <mutex>
#include "../../include/nvorbis_b200.h"
#include "nvb_unpack_tables.h"

#if !defined(__CUDACC__) && !defined(NVB_HAVE_FLOAT2)
// host-only compilation of nvb_host.cpp: the CUDA vector type the tables use
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

// Kernel launch / dynamic shared memory spelled as macros so that tests/cpu_shim (test-only) can
// compile the same kernel sources against its thread-per-lane emulation.
// 1: the launch macros allow programmatic dependent launch (the default); 0 inside a NoEarlyStart scope: the kernel starts only
// after everything before it in the stream has completed -- needed when a kernel's pre-wait prologue reads what the PREVIOUS
// KERNEL writes (the spectrum kernels read posts / classes / plan records before griddepcontrol.wait, which is fine when
// those come from host copies and wrong when k_unpack produces them).
inline thread_local int nvb_launch_pdl = 1;
struct NvbNoEarlyStart { int prev; explicit NvbNoEarlyStart(bool on) : prev(nvb_launch_pdl) { if (on) nvb_launch_pdl = 0; } ~NvbNoEarlyStart() { nvb_launch_pdl = prev; } };
#if !defined(NVB_CPU_SHIM)
// Every kernel is launched with programmatic stream serialization (PDL): its blocks may be scheduled, and run their
// prologue (shared-memory setup, table staging), while the previous kernel of the stream drains; nvb_grid_dep_wait()
// then blocks until that kernel has completed and its writes are visible, so stream order is preserved for all data.
#define NVB_LAUNCH(kernel, grid, block, smem, stream, arg)                                                   \
    do {                                                                                                    \
        cudaLaunchConfig_t cfg__ = {};                                                                      \
        cfg__.gridDim = dim3((unsigned)(grid)); cfg__.blockDim = dim3((unsigned)(block));                   \
        cfg__.dynamicSmemBytes = (size_t)(smem); cfg__.stream = (cudaStream_t)(stream);                     \
        cudaLaunchAttribute attr__[1];                                                                      \
        attr__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                  \
        attr__[0].val.programmaticStreamSerializationAllowed = nvb_launch_pdl;                                           \
        cfg__.attrs = attr__; cfg__.numAttrs = 1;                                                           \
        cudaLaunchKernelEx(&cfg__, kernel, arg);                                                            \
    } while (0)
#define NVB_LAUNCHV(kernel, grid, block, smem, stream, ...)                                                 \
    do {                                                                                                    \
        cudaLaunchConfig_t cfg__ = {};                                                                      \
        cfg__.gridDim = dim3((unsigned)(grid)); cfg__.blockDim = dim3((unsigned)(block));                   \
        cfg__.dynamicSmemBytes = (size_t)(smem); cfg__.stream = (cudaStream_t)(stream);                     \
        cudaLaunchAttribute attr__[1];                                                                      \
        attr__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                  \
        attr__[0].val.programmaticStreamSerializationAllowed = nvb_launch_pdl;                                           \
        cfg__.attrs = attr__; cfg__.numAttrs = 1;                                                           \
        cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__);                                                    \
    } while (0)
#define NVB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#if defined(__CUDA_ARCH__)
#define nvb_grid_dep_wait() asm volatile("griddepcontrol.wait;" ::: "memory")
#define nvb_grid_dep_launch() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#else
#define nvb_grid_dep_wait() ((void)0)
#define nvb_grid_dep_launch() ((void)0)
#endif
#endif

// Channel counts up to this run on the specialised spectrum kernels (channels of a bin in registers, per-channel tables in static shared
// memory); setups with more channels (up to NVB_MAX_CHANNELS) take the general kernel's many-channel instantiation.
#define NVB_FAST_CHANNELS 8

namespace nvb {

// Launch-configuration caches, safe for several host threads (a context is single-threaded, but contexts on several
// threads / devices launch the same kernels): the largest dynamic shared memory a kernel has been opted in to on a device,
// and the device's SM count.  Zero-initialised statics of these types are valid.
template <class SetAttr> inline bool nvb_ensure_smem(std::atomic<size_t>& slot, size_t smem, SetAttr set_attr) {
    if (smem <= slot.load(std::memory_order_acquire)) return true;
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    if (smem <= slot.load(std::memory_order_relaxed)) return true;
    if (!set_attr()) return false;
    slot.store(smem, std::memory_order_release);
    return true;
}

// ---- immutable per-stream tables, as they sit in ONE contiguous device allocation ("blob") ----
struct DevBook    { int32_t dims, entries; int64_t off; int32_t dshift, pad; };   // off: float index into vq, -1 = no table; dshift = log2(dims) or -1
struct DevFloor1  {
    int32_t n_posts, mult, range, max_level;
    uint16_t x[NVB_MAX_POSTS]; uint8_t lo[NVB_MAX_POSTS], hi[NVB_MAX_POSTS], sort[NVB_MAX_POSTS];
    uint8_t level[NVB_MAX_POSTS];   // depth of post i in the neighbour dependency tree: 1 + max(level[lo], level[hi]); posts 0, 1 are level 0
    float rcp[NVB_MAX_POSTS];       // 1.0f / (x[hi[i]] - x[lo[i]]): RenderPoint's divisor is a setup constant
    uint16_t xs[NVB_MAX_POSTS];     // x[sort[k]]: the x list in ascending order (k_spectrum_run)
    uint32_t magic[NVB_MAX_POSTS];  // floor(2^32 / adx) + 1, adx = x[hi[i]] - x[lo[i]] >= 2: RenderPoint's division as one multiply-high (k_spectrum_wf)
};
// Floor type 0 (Floor0.cs): one record per floor of the setup (type 1 floors carry type = 1 and nothing else).
// bark / wmap: element offsets into DevSetup.f0_bark (int32, n entries: barkMap[0..n)) and DevSetup.f0_wmap (float, n entries)
// for the short / long block size; kmax = largest bark index that occurs.
struct DevFloor0 { int32_t type, order, amp_ofs, bark_map_size; int32_t bark_off[2], wmap_off[2], kmax[2]; int32_t pad[2]; };
// k_spectrum_run: one record per (class, stage) of a residue, DevResidue.ci_off + class * stages + stage
struct alignas(16) CiRec { int32_t off, dshift, entries, cnt; };   // VQ table float offset, log2(dims), book entries, entries per partition (0 = nothing coded)
struct DevResidue {
    int32_t type, begin, end, psize, nclass, stages;
    int32_t pshift;                 // log2(psize) or -1
    int32_t fast;                   // 1: k_spectrum_fast applies (power-of-two partition/book sizes, type 2 partitions aligned to the channel count); 2: k_spectrum_planes too; 3: k_spectrum_run too
    int32_t cascade[NVB_MAX_CLASSES];
    int16_t books[NVB_MAX_CLASSES][NVB_MAX_STAGES];
    int16_t cnt[NVB_MAX_CLASSES][NVB_MAX_STAGES];   // VQ entries one partition of (class, stage) consumes; 0 = nothing coded
    uint8_t coded[NVB_MAX_CLASSES];                 // per class: bit s set <=> stage s codes entries (cnt > 0)
    int32_t ci_off, pad2;                           // k_spectrum_warp: start of this residue's (class, stage) records in its shared table
};
// k_spectrum_run: everything a frame's CTA needs to know about its mode in one 64-byte record (one load instead of the
// mode -> mapping -> residue -> floor chain of dependent loads)
struct alignas(16) RunMode {
    int32_t rbegin, rend, pshift, stages;
    int32_t nclass, ci_off, residue, floor;
    int32_t n_coupling, mapping, block_flag, rtype;
    int32_t cand_off, ob_off;       // k_spectrum_bins: this residue's slices of DevSetup.r2cand / r2ob
    int32_t bins_ok, pad;           // k_spectrum_bins applies to this mode
    int32_t cc_off, base_stride;    // k_spectrum_wf: this residue's slice of DevSetup.cls_cnt; stages rounded up to a multiple of 4
    int32_t n_posts, pad3;          // posts of this mode's floor (type 1)
};
// k_spectrum_wf: line segment that starts at an active post (x-sorted position k): RenderLineMulti(x0, y0, x1 = min(hx, n), hy)
// (Floor1.cs:206); x01 = x0 | x1 << 16 (0xffff: the flat tail, Floor1.cs:213-216); m as in RunSeg.
struct alignas(16) WfSeg { uint32_t x01; int32_t y0, dy; uint32_t m; };
// k_spectrum_wf smem bytes per frame group (host-computed, nvb_host.cpp: wf_layout)
struct WfLayout { int32_t seg_off, fy_off, base_off, cls_off, total, np_pad, cta_bytes, pad; };   // cta_bytes: the CTA-wide part in front of the groups (CiRec table)
struct DevMapping { int32_t n_coupling, floor, residue, pad; uint8_t mag[NVB_MAX_COUPLING], ang[NVB_MAX_COUPLING]; };
struct DevMode    { int32_t block_flag, mapping; };

// Blob header: byte offsets of every section from the start of the blob.
struct BlobHeader {
    uint32_t magic;            // 'NVB1'
    uint32_t abi;
    uint64_t total_bytes;
    int32_t channels, sample_rate, bs[2];
    int32_t n_books, n_floors, n_residues, n_mappings, n_modes;
    int32_t post_stride;       // int16 elements per (frame, channel) in nvb_batch.posts
    int32_t max_items;         // max over modes of stages*partitions*streams (residue prefix table)
    int32_t spectrum_fast;     // 3: every mode fits k_spectrum_run, 2: k_spectrum_planes, 1: k_spectrum_fast, 0: only the general k_spectrum
    uint64_t off_books, off_vq, off_floors, off_residues, off_mappings, off_modes;
    uint64_t off_win_short;    // bs[0] floats
    uint64_t off_win_long;     // 4 * bs[1] floats (window index = prev?1:0 + next?2:0)
    uint64_t off_mdct_a[2], off_mdct_b[2], off_mdct_c[2], off_bitrev[2];   // reference twiddles (exact path)
    uint64_t off_tw[2];        // fast path: float2[bs/4], exp(-i*pi*(k+1/8)/(bs/2))
    uint64_t off_fft[2];       // fast path: float2[bs/4], exp(-2*pi*i*k/(bs/4))
    uint64_t off_db;           // 256 floats
    uint64_t n_vq;
    uint64_t off_fused_tab;    // FusedTables block (nvb_fused_core.h) when bs == {256, 2048}, else 0
    int32_t max_stages;        // largest residue stage count of the setup
    int32_t ci_total;          // sum over residues of nclass * stages: size of the (class, stage) table
    uint64_t off_ci;           // CiRec[ci_total]
    uint64_t off_bin2k;        // uint8[n_floors][bs[1]/2]: sorted position of the last floor post with x <= bin
    uint64_t off_run_modes;    // RunMode[n_modes]
    uint64_t off_floors0;      // DevFloor0[n_floors]
    uint64_t off_f0_bark, off_f0_wmap;   // pooled int32 / float tables of the type 0 floors
    uint64_t n_f0_bark, n_f0_wmap;
    int32_t f0_stride;         // floats per (frame, channel) in nvb_batch.floor0; 0 = the setup has no type 0 floor
    int32_t f0_max_order;
    // k_spectrum_bins (type 2 residues whose partitions are not aligned to the channel count, Residue2.cs:25-27):
    uint64_t off_r2cand;       // uint32[n_residues][bs[1]/2]: first partition that reaches the bin | (number of such partitions << 16)
    uint64_t off_r2ob;         // uint16[n_residues][r2_max_p]: (begin + p * psize) / channels, the bin a partition starts at
    int32_t r2_max_p;
    int32_t spectrum_bins;     // 1: every mode can run k_spectrum_bins
    uint64_t body_hash;        // FNV-1a 64 of everything behind the header: nvb_setup_blob_import rejects a damaged blob
    uint64_t off_magic;        // uint32[bs[1]/2 + 1][2]: {floor(2^32 / adx) + 1, largest |dy| the multiply-high is exact for} per segment length adx
    uint64_t off_cls_cnt;      // uint64[sum over residues of ceil(stages / 4) * nclass]: entries per partition of (class, stage), 16 bits per stage
    int32_t cls_cnt_total;
    int32_t wf_max_p;          // largest partition count of any mode (k_spectrum_wf's per-frame tables)
};

// Resolved pointers handed to kernels by value.
struct DevSetup {
    int32_t channels, bs[2], post_stride, max_items, spectrum_fast, max_stages, ci_total, n_residues;
    const DevBook* books; const float* vq; int64_t n_vq;
    const DevFloor1* floors; const DevResidue* residues; const DevMapping* mappings; const DevMode* modes;
    const float* win_short; const float* win_long;
    const float* A[2]; const float* B[2]; const float* C[2]; const uint16_t* bitrev[2];
    const float2* tw[2]; const float2* fft[2];
    const float* db;
    const float* fused_tab;    // lane tables of the fused kernel, nullptr when the block sizes are not {256, 2048}
    const CiRec* ci; const uint8_t* bin2k; const RunMode* run_modes;
    const DevFloor0* floors0; const int32_t* f0_bark; const float* f0_wmap; int32_t f0_stride, f0_max_order;
    const uint32_t* r2cand; const uint16_t* r2ob; int32_t spectrum_bins, r2_max_p;
    const uint32_t* magic; const unsigned long long* cls_cnt; int32_t wf_max_p, max_posts;
};

// ---- per-frame plan built on the host from nvb_frame (ok frames only, in order) ----------------
enum { PREV_NONE = -1, PREV_CARRY = -2 };
struct alignas(16) DevFrame {
    uint8_t  mode, window, res_decoded, kind;   // kind 0 = decoded block; 1 = drain: emit [out_begin,out_end) of block `prev` as it is
    uint32_t exec_mask;
    int32_t  n;                 // block size
    int32_t  start;             // packetStartIndex: prev tail is added at [start, start+ola_len)
    int32_t  out_begin, out_end;// emitted region of this block
    int32_t  ola_len;
    int32_t  prev;              // index of the previous DevFrame, PREV_NONE or PREV_CARRY
    int32_t  prev_valid;        // where the tail starts inside the previous block
    int32_t  api_index;         // index into nvb_batch.frames (posts are addressed with it)
    int64_t  pcm_off;           // per-channel sample offset of out_begin in the PCM output
    uint32_t classes_off, entries_off, entry_count;
    uint32_t spec_off;          // float offset of this frame's [C][n/2] spectrum; its [C][n] block sits at 2*spec_off
};

struct Counters { int clipped, floor_range, bad_entry, pad; };

// ---- launchers (nvb_kernels.cu) -----------------------------------------------------------------
struct LaunchArgs {
    DevSetup S;
    const DevFrame* frames;     // device array of the whole batch plan
    int frame_lo, n_frames;     // this launch covers plan frames [frame_lo, frame_lo + n_frames)
    const int16_t* posts; const uint8_t* classes; const uint16_t* entries;
    const float* floor0;        // [frame][channel][f0_stride] or nullptr
    float* spectrum;            // [sum C*n/2]
    float* blocks;              // [sum C*n]   (exact path scratch)
    const float* carry_in;      // [C][bs1] previous batch's last block (or nullptr)
    float* carry_out;           // [C][bs1] receives the windowed block of frame `carry_frame` (or nullptr)
    int carry_frame;            // DevFrame index of the batch's last decoded block, -1 = none
    float* pcm;
    Counters* counters;
    int clip;
    int inputs_from_kernel;     // 1: posts / classes / entries / plan records were written by k_unpack just before (no early start of the spectrum kernel)
};

// ---- GPU-side packet unpack (nvb_unpack.cu) ------------------------------------------------------
struct UnpackTables {           // the unpack-tables blob (nvb_unpack_tables.h) resolved against its device allocation
    const nvbu::UBook* books; const uint32_t* roots; const nvbu::ULong* longs; const nvbu::UFloor1* floors; const nvbu::UResidue* residues;
    const uint8_t* digits; const nvbu::UMapping* mappings; const nvbu::UMode* modes;
    int32_t channels, mode_bits, post_stride, cls_stride, ent_stride;
};
struct UnpackArgs {
    UnpackTables T;
    DevFrame* frames; int frame_lo, n_frames;      // plan frames [frame_lo, frame_lo + n_frames): exec_mask / res_decoded / entry_count are written
    const uint8_t* data; const uint32_t* offsets;  // packets of the batch (by api_index), padded by 8 bytes
    int16_t* posts; uint8_t* classes; uint16_t* entries;
    float* floor0; int f0_stride;                  // type 0 floor records [packet][channel][f0_stride] (nullptr: the setup has none)
    const DevBook* dbooks; const float* vq;        // the synthesis setup's VQ tables: Floor0.Unpack reads its coefficients out of them
};
int launch_unpack(const UnpackArgs& a, void* stream);

int launch_spectrum(const LaunchArgs& a, void* stream);            // picks k_spectrum_fast when every residue of the setup allows it
int launch_spectrum_generic(const LaunchArgs& a, void* stream);
int launch_imdct_exact(const LaunchArgs& a, void* stream);   // spectrum -> windowed blocks
int launch_ola(const LaunchArgs& a, void* stream);           // blocks (+carry) -> interleaved PCM
int launch_pcm_s16(const float* src, int16_t* dst, long long lo, long long hi, void* stream);   // NVB_RUN_PCM_S16: elements [lo, hi)
// Fused fast path: spectrum -> PCM for runs of frames; returns <0 if the batch shape is not covered.
int launch_imdct_fused(const LaunchArgs& a, const DevFrame* host_frames, void* stream);
bool fused_supported(const BlobHeader& h, const DevFrame* host_frames, int n_frames);
// One-kernel synthesis (records -> PCM, k_imdct_fused_t<false, C>): 1 = launched, -2 = the setup / batch is not covered or -- unless
// forced -- the launch is larger than one round of the CTAs' warps (take the two-kernel path), -1 = launch error.
int launch_synth_fused(const LaunchArgs& a, const DevFrame* host_frames, void* stream, bool forced);

}  // namespace nvb
