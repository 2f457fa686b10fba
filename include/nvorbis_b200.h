/* nvorbis_b200.h -- C ABI of the B200 Vorbis synthesis path.
 *
 * What this replaces in the reference (NVorbis, C#; file:line relative to NVorbis/):
 *   the SYNTHESIS half of one audio packet, batched over K packets:
 *     Residue0/1/2.WriteVectors  `res[o] += book[entry,dim]`   Residue0.cs:180-201, Residue1.cs:8-26, Residue2.cs:23-47
 *     inverse channel coupling                                  Mapping.cs:137-182
 *     Floor1.Apply (UnwrapPosts + RenderLineMulti)              Floor1.cs:186-341
 *     Floor0.Apply (LSP -> bark-band curve)                     Floor0.cs:152-212
 *     Mdct.Reverse                                              Mdct.cs:13-21,65-535
 *     window multiply                                           Mode.cs:159-166
 *     OverlapBuffers / ClippingCopyBuffer / CopyBuffer          StreamDecoder.cs:391-415,532-541
 *   i.e. everything IMode.Decode (Contracts/IMode.cs:8) does after the bits have been read, plus
 *   the overlap/clip/interleave loop of StreamDecoder.Read (StreamDecoder.cs:320-389).
 *
 * What stays on the host (the caller): Ogg paging, header parsing, Huffman/codebook bit
 * unpacking (Codebook.DecodeScalar, Floor1.Unpack, the class/entry reads of Residue0.Decode,
 * Mode.GetPacketInfo) and the EOS trim (StreamDecoder.cs:429-437).  The host hands over, per
 * packet, exactly the values those functions produce -- see nvb_frame / nvb_batch.
 *
 * All functions are cdecl, take plain pointers/sizes, never throw, and return an nvb_status.
 * A context is single-threaded (like a reference StreamDecoder); several contexts may coexist.
 */
#ifndef NVORBIS_B200_H
#define NVORBIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVB_ABI_VERSION 3

#define NVB_MAX_CHANNELS   32    /* reference allows 255 (StreamDecoder.cs:186); nvb_frame.exec_mask has 32 bits; larger -> NVB_ERR_UNSUPPORTED.
                                    Up to 8 channels run on the specialised spectrum kernels, 9..32 on the general one. */
#define NVB_MAX_POSTS      64    /* Floor1.Data.Posts = new int[64]            (Floor1.cs:12)   */
#define NVB_MAX_CLASSES    64    /* residue classifications: 6 bits + 1         (Residue0.cs:41) */
#define NVB_MAX_STAGES     8     /* cascade: 3 + 5 bits                          (Residue0.cs:48-56) */
#define NVB_MAX_COUPLING   256   /* 8 bits + 1 (Mapping.cs:28): every step count the reference accepts */
#define NVB_MAX_IN_FLIGHT  3     /* batches between nvb_decode_batch_begin and _end */

typedef enum nvb_status {
    NVB_OK = 0,
    NVB_ERR_ARG = -1,          /* NULL / out-of-range argument (reference: ArgumentException family)     */
    NVB_ERR_CUDA = -2,         /* CUDA runtime failure, see nvb_last_error()                             */
    NVB_ERR_UNSUPPORTED = -3,  /* setup outside the supported envelope (see nvb_upload_setup)            */
    NVB_ERR_NOMEM = -4,
    NVB_ERR_STATE = -5,        /* call order (no setup uploaded, ...)  (reference: ObjectDisposedException) */
    NVB_ERR_CAPACITY = -6,     /* pcm_out too small for the batch                                        */
    NVB_ERR_DATA = -7          /* malformed batch (offsets out of range, ...) (reference: InvalidDataException) */
} nvb_status;

typedef struct nvb_ctx nvb_ctx;
typedef struct nvb_dbatch nvb_dbatch;

/* ---- setup (one per logical stream; built by the host from the setup header) ---------------- */

/* Codebook VQ lookup (ICodebook this[entry,dim], Codebook.cs:322; table built by
 * Codebook.InitLookupTable, Codebook.cs:222-283 -- the host computes the floats, the GPU gathers). */
typedef struct nvb_codebook {
    int32_t dims;          /* Codebook.Dimensions */
    int32_t entries;       /* Codebook.Entries    */
    int32_t map_type;      /* 0 = no lookup table */
    int32_t reserved;
    int64_t table_off;     /* index of [entry*dims + dim] floats inside nvb_setup.vq_floats; -1 if none */
} nvb_codebook;

/* Floor type 1 tables (Floor1.Init, Floor1.cs:30-133) */
typedef struct nvb_floor1 {
    int32_t n_posts;                    /* _xList.Length (<= 64)            */
    int32_t multiplier;                 /* _multiplier (already +1)  :69-74 */
    int32_t range;                      /* _range                    :71    */
    int32_t reserved;
    uint16_t x_list[NVB_MAX_POSTS];     /* _xList                    :78-90 */
    uint8_t  l_neigh[NVB_MAX_POSTS];    /* _lNeigh                   :93-115 */
    uint8_t  h_neigh[NVB_MAX_POSTS];    /* _hNeigh                          */
    uint8_t  sort_idx[NVB_MAX_POSTS];   /* _sortIdx                  :118-132 */
} nvb_floor1;

/* Floor type 0 parameters (Floor0.Init, Floor0.cs:28-65).  The library derives the bark map and the 2*cos map per block
 * size itself (SynthesizeBarkCurve / SynthesizeWDelMap, Floor0.cs:67-96, same float/double expression order). */
typedef struct nvb_floor0 {
    int32_t order;                      /* _order (1..255)  */
    int32_t rate;                       /* _rate            */
    int32_t bark_map_size;              /* _bark_map_size   */
    int32_t amp_bits;                   /* _ampBits (used by the host's Unpack only) */
    int32_t amp_ofs;                    /* _ampOfs          */
    int32_t reserved[3];
} nvb_floor0;

typedef struct nvb_floor {
    int32_t type;                       /* 1 = Floor1 (f1), 0 = Floor0 (f0)  (Factory.cs:22-31) */
    int32_t reserved;
    nvb_floor1 f1;
    nvb_floor0 f0;
} nvb_floor;

/* Residue tables (Residue0.Init, Residue0.cs:35-117) */
typedef struct nvb_residue {
    int32_t type;                       /* 0, 1, 2 (Factory.cs:48-58) */
    int32_t begin, end;                 /* _begin, _end               */
    int32_t partition_size;             /* _partitionSize             */
    int32_t classifications;            /* _classifications           */
    int32_t max_stages;                 /* _maxStages                 */
    int32_t cascade[NVB_MAX_CLASSES];   /* _cascade                   */
    int16_t books[NVB_MAX_CLASSES][NVB_MAX_STAGES];  /* _books[class][stage] as codebook index, -1 = null */
} nvb_residue;

/* Mapping tables (Mapping.Init, Mapping.cs:16-93).  Only single-submap mappings are accepted: with
 * more than one submap the reference forces every channel silent yet decodes each submap's residue
 * over all channels (Mapping.cs:122-134) -- rejected with NVB_ERR_UNSUPPORTED instead of emulated. */
typedef struct nvb_mapping {
    int32_t n_coupling;
    int32_t n_submaps;
    uint8_t magnitude[NVB_MAX_COUPLING];   /* _couplingMangitude */
    uint8_t angle[NVB_MAX_COUPLING];       /* _couplingAngle     */
    int32_t floor;                         /* _submapFloor[0]   as index into nvb_setup.floors   */
    int32_t residue;                       /* _submapResidue[0] as index into nvb_setup.residues */
} nvb_mapping;

typedef struct nvb_mode {               /* Mode.Init, Mode.cs:24-67 */
    int32_t block_flag;                 /* _blockFlag: 1 = long (block_size[1]) */
    int32_t mapping;
} nvb_mode;

typedef struct nvb_setup {
    int32_t abi_version;                /* NVB_ABI_VERSION */
    int32_t channels;                   /* StreamDecoder._channels */
    int32_t sample_rate;
    int32_t block_size[2];              /* _block0Size, _block1Size (StreamDecoder.cs:192-193) */
    int32_t n_books, n_floors, n_residues, n_mappings, n_modes;
    const nvb_codebook* books;
    const float*        vq_floats;      /* concatenated lookup tables */
    int64_t             n_vq_floats;
    const nvb_floor*    floors;
    const nvb_residue*  residues;
    const nvb_mapping*  mappings;
    const nvb_mode*     modes;
    /* Optional host-computed tables.  Pass NULL to let the library compute them with the same
     * float/double expression order as the reference (libm instead of System.Math).
     *   window_slope[i]: block_size[i]/2 floats, the rising slope of Mode.CalcWindow (Mode.cs:80-85)
     *   mdct_a/b/c/bitrev[i]: Mdct twiddles for block_size[i] (Mdct.cs:44-62): n/2, n/2, n/4 floats, n/8 u16 */
    const float*    window_slope[2];
    const float*    mdct_a[2];
    const float*    mdct_b[2];
    const float*    mdct_c[2];
    const uint16_t* mdct_bitrev[2];
} nvb_setup;

/* ---- per-packet boundary record ------------------------------------------------------------ */

enum { NVB_FRAME_OK = 0, NVB_FRAME_FAILED = 1 };

/* One audio packet as seen by StreamDecoder.DecodeNextPacket (StreamDecoder.cs:465-530).
 * status FAILED = DecodeNextPacket returned null (non-audio packet, IsShort, or no packet): the
 * previous block's tail is drained un-overlapped (StreamDecoder.cs:352-356); no other field is read. */
typedef struct nvb_frame {
    uint8_t  status;        /* NVB_FRAME_OK / NVB_FRAME_FAILED */
    uint8_t  mode;          /* index into setup.modes (StreamDecoder.cs:497) */
    uint8_t  window;        /* Mode.GetPacketInfo windowIndex = prev?1:0 + next?2:0; 0 for short blocks (Mode.cs:135) */
    uint8_t  res_decoded;   /* 1 if Residue.Decode ran (some channel live: Residue0.cs:125), else 0 */
    uint32_t exec_mask;     /* bit c = IFloorData.ExecuteChannel of channel c after the ForceEnergy pass (Mapping.cs:111-119) */
    int32_t  start;         /* packetStartIndex                                   (Mode.cs:137-140) */
    int32_t  valid;         /* packetValidLength AFTER the EOS trim               (StreamDecoder.cs:429-437) */
    int32_t  total;         /* packetTotalLength */
    uint32_t classes_off;   /* into nvb_batch.classes: [stream][partition] u8, stream = channel (types 0/1) or 0 (type 2) */
    uint32_t entries_off;   /* into nvb_batch.entries: VQ entry numbers in the order DecodeScalar returned them */
    uint32_t entry_count;   /* entries decoded before the packet ended / failed ("use what we have", Residue0.cs:164-170).
                               Residue type 0 adds a partition only when all of its entries were read
                               (Residue0.cs:186-192): truncate entry_count to that partition boundary. */
} nvb_frame;

/* Floor payload: posts[frame][channel][nvb_post_stride()] int16; element 0 = PostCount (0 = floor
 * unused or unpack failed, Floor1.cs:140,155-174), elements 1.. = raw Posts[] as unpacked (before
 * UnwrapPosts).  nvb_post_stride() = 2 + max n_posts over the setup's floors, rounded up to even. */
/* Floor type 0 payload (only read for packets whose mapping uses a type 0 floor; may be NULL otherwise):
 * floor0[frame][channel][nvb_floor0_stride()] floats; element 0 = Data.Amp as Floor0.Unpack leaves it (Floor0.cs:107-114:
 * raw / ampDiv * ampOfs, or 0 when the floor is unused / the unpack failed), elements 1 .. order = Data.Coeff after the
 * "averaging" pass (Floor0.cs:136-147).  nvb_floor0_stride() = 1 + max order over the setup's type 0 floors, rounded up
 * to even (0 when the setup has none). */
typedef struct nvb_batch {
    int32_t          n_frames;
    int32_t          reserved;
    const nvb_frame* frames;
    const int16_t*   posts;
    const uint8_t*   classes;   int64_t n_classes;
    const uint16_t*  entries;   int64_t n_entries;
    const float*     floor0;
} nvb_batch;

typedef struct nvb_result {
    int64_t samples_per_channel;  /* PCM frames written (interleaved: samples_per_channel * channels floats) */
    int32_t has_clipped;          /* StreamDecoder.HasClipped (Utils.ClipValue, Utils.cs:30-43) */
    int32_t n_failed;             /* frames with status FAILED */
    int32_t n_floor_range;        /* floor curve indices outside inverse_dB_table (reference: IndexOutOfRangeException); clamped */
    int32_t n_inconsistent;       /* frames whose previous tail was longer than their own overlap region (clamped) */
} nvb_result;

/* nvb_run flags */
enum {
    NVB_RUN_DEFAULT   = 0,
    NVB_RUN_EXACT     = 1,   /* stb-dataflow IMDCT without FMA contraction: bit-identical to the reference arithmetic */
    NVB_RUN_NO_CLIP   = 2,   /* StreamDecoder.ClipSamples = false (CopyBuffer instead of ClippingCopyBuffer) */
    NVB_RUN_CONTINUE  = 4,   /* chain onto the previous batch of this context (keeps the overlap tail);
                                without it the batch starts a stream: first block emits nothing (StreamDecoder.cs:446-450) */
    NVB_RUN_PCM_S16   = 8,   /* nvb_decode_batch[_begin]: pcm_out receives 16-bit PCM (int16_t*, pcm_cap counts int16 elements):
                                s = round-to-nearest-even(v * 32768) saturated to [-32768, 32767], v = the float sample the call
                                would otherwise return (after Utils.ClipValue unless NVB_RUN_NO_CLIP).  Halves the read-back;
                                the reference itself only ever writes 32-bit float WAV (TestApp/WaveWriter.cs:18-62). */
    NVB_RUN_DEVICE_OUT = 16, /* nvb_decode_batch[_begin]: pcm_out is a DEVICE pointer on the context's GPU (16-byte aligned): the PCM
                                is left there for an on-device consumer, nothing is copied back (float, or int16 with
                                NVB_RUN_PCM_S16).  The buffer is ready when the call / the batch's _end returns. */
    NVB_RUN_ONE_KERNEL = 32, /* whole synthesis (Mapping.DecodePacket's float half .. OverlapBuffers, Mapping.cs:95-198 -> Mdct.cs:65-313
                                -> Mode.cs:159-166 -> StreamDecoder.cs:532-541) in ONE kernel launch per batch: each warp computes its
                                frame's spectrum into shared memory and transforms it there; no dense spectrum in device memory.
                                Same results as the two-kernel path, bit for bit.  Applies to mono / stereo streams with 256/2048
                                blocks and residues the run-per-lane spectrum stage covers; other streams silently take the two-kernel
                                path (nvb_dbatch_launches tells).  Without this flag the library picks per launch: one kernel for
                                launches of up to 15 frames per SM (one round of its warps: lower latency), two kernels above.
                                NVB_ONE_KERNEL=1 / 0 in the environment forces it on / off. */
    NVB_RUN_TWO_KERNELS = 64 /* the opposite request: spectrum kernel + IMDCT kernel for every launch */
};

/* ---- entry points ---------------------------------------------------------------------------- */

int         nvb_abi_version(void);
const char* nvb_strerror(int status);
const char* nvb_last_error(nvb_ctx* ctx);           /* detail text of the last failure on this context */

/* Creates a context on CUDA device `device`.  Fails with NVB_ERR_CUDA when no usable GPU exists:
 * there is no CPU fallback. */
int nvb_create(int device, nvb_ctx** out);
int nvb_destroy(nvb_ctx* ctx);

/* Page-locked host memory for batch inputs / PCM output (the P/Invoke host keeps its batch arrays in
 * these so that the H2D/D2H copies of nvb_decode_batch run at full PCIe rate).  Any other host
 * pointer is accepted by every call too, just slower. */
int nvb_host_alloc(size_t bytes, void** out);
int nvb_host_free(void* p);

/* Uploads the immutable per-stream tables.  Replaces what StreamDecoder.LoadBooks leaves behind
 * (StreamDecoder.cs:226-289).  NVB_ERR_UNSUPPORTED: channels > NVB_MAX_CHANNELS, multi-submap mappings,
 * > NVB_MAX_COUPLING steps, residue books with > 65536 entries, a type 0 floor whose bark map indexes past its cos map
 * (Floor0.cs:166 would throw IndexOutOfRangeException on every packet). */
int nvb_upload_setup(nvb_ctx* ctx, const nvb_setup* setup);

/* Serialises the uploaded setup's device tables into one blob / installs such a blob, so that one
 * rank can parse the headers and every other rank receives the tables with a single broadcast. */
int nvb_setup_blob_size(nvb_ctx* ctx, size_t* bytes);
int nvb_setup_blob_export(nvb_ctx* ctx, void* dst, size_t bytes);
int nvb_setup_blob_import(nvb_ctx* ctx, const void* src, size_t bytes);

int nvb_post_stride(nvb_ctx* ctx);
int nvb_floor0_stride(nvb_ctx* ctx);

/* ResetDecoder (StreamDecoder.cs:295-305): forget the overlap tail. */
int nvb_reset(nvb_ctx* ctx);

/* Host-buffer call: H2D of the batch, synthesis, D2H of interleaved PCM into pcm_out
 * (capacity pcm_cap floats).  Equivalent to what K iterations of StreamDecoder.Read's packet
 * loop leave in the caller's buffer. */
int nvb_decode_batch(nvb_ctx* ctx, const nvb_batch* batch, int flags, float* pcm_out, size_t pcm_cap, nvb_result* res);

/* The same call split in two, so that a host which unpacks the next run of packets while the GPU works (the batching
 * StreamDecoder.Read loop, StreamDecoder.cs:320-389) keeps PCIe busy in both directions: _begin enqueues the H2D copies,
 * the kernels and the D2H copy into pcm_out and returns; _end waits for the OLDEST batch begun and reports it.  At most
 * NVB_MAX_IN_FLIGHT batches may be in flight (one more _begin returns NVB_ERR_STATE; two keep the float read-back busy, a third
 * pays when the read-back is short: 16-bit PCM, device output); batches complete in the order they were begun and chain
 * their overlap tails exactly like consecutive nvb_decode_batch calls (NVB_RUN_CONTINUE).  The batch arrays and pcm_out of a
 * batch must stay alive and untouched until its _end; keep them in nvb_host_alloc memory for the copies to be asynchronous. */
int nvb_decode_batch_begin(nvb_ctx* ctx, const nvb_batch* batch, int flags, float* pcm_out, size_t pcm_cap);
int nvb_decode_batch_end(nvb_ctx* ctx, nvb_result* res);

/* ---- GPU-side packet unpack: raw audio packets in, PCM out -------------------------------------
 * The bit-reading half of Mapping.DecodePacket -- Floor1.Unpack (Floor1.cs:135-184), the class words and VQ entry numbers of
 * Residue0.Decode (Residue0.cs:119-178) through Codebook.DecodeScalar (Codebook.cs:294-320), the energy flags
 * (Mapping.cs:105-119) -- runs on the device too, one thread per packet (k_unpack), and writes the same boundary records
 * the host would have sent.  The host keeps what is sequential across packets: container paging and, per packet, the values
 * of Mode.GetPacketInfo (Mode.cs:119-151: mode / window flags from the packet's first bits) after the granule bookkeeping
 * and the end-of-stream trim (StreamDecoder.cs:417-463).  H2D shrinks to the raw packets (about 300 B instead of 1.5 KB per
 * stereo frame).  Type 0 floors (Floor0.Unpack, Floor0.cs:98-150) are unpacked on the device as well. */
typedef struct nvb_packet_batch {
    int32_t          n_packets;
    int32_t          reserved;
    const nvb_frame* frames;    /* per packet: status, mode, window, start, valid, total; every other field is ignored (device-produced) */
    const uint8_t*   data;      /* packet bytes back to back */
    const uint32_t*  offsets;   /* n_packets + 1 byte offsets into data */
} nvb_packet_batch;

/* Installs the unpack tables (Huffman decode tables, floor / residue structure) next to the uploaded setup.  The blob is built
 * by the host half from the same setup header (include/nvorbis_host.h: nvh_unpack_tables); layout: csrc/nvb_unpack_tables.h. */
int nvb_upload_unpack_tables(nvb_ctx* ctx, const void* blob, size_t bytes);
/* nvb_decode_batch / _begin for raw packets (same flags, same pipeline, completed by nvb_decode_batch_end). */
int nvb_decode_packets(nvb_ctx* ctx, const nvb_packet_batch* batch, int flags, float* pcm_out, size_t pcm_cap, nvb_result* res);
int nvb_decode_packets_begin(nvb_ctx* ctx, const nvb_packet_batch* batch, int flags, float* pcm_out, size_t pcm_cap);
/* Parity / inspection: unpacks on the device and returns the boundary records in the nvb_batch layout with fixed strides --
 * frames_out[n_packets] (classes_off = i * cls_stride, entries_off = i * ent_stride), posts_out[n_packets][channels][post_stride],
 * classes_out[n_packets * cls_stride], entries_out[n_packets * ent_stride]; strides from nvb_unpack_strides. */
int nvb_unpack_strides(nvb_ctx* ctx, int32_t* cls_stride, int32_t* ent_stride);
int nvb_unpack_packets(nvb_ctx* ctx, const nvb_packet_batch* batch, nvb_frame* frames_out, int16_t* posts_out, uint8_t* classes_out, uint16_t* entries_out);

/* Device-resident variant (pipelining / benchmarking): upload once, run many times.
 * `stream` is a cudaStream_t (NULL = default stream); d_pcm is a device pointer with room for
 * nvb_dbatch_samples()*channels floats.  nvb_dbatch_run only enqueues kernels.  With
 * NVB_RUN_CONTINUE the batch overlaps onto the tail left by the last nvb_decode_batch; running a
 * dbatch never changes that tail. */
int     nvb_dbatch_create(nvb_ctx* ctx, const nvb_batch* batch, int flags, nvb_dbatch** out);
int64_t nvb_dbatch_samples(const nvb_dbatch* b);
int     nvb_dbatch_run(nvb_ctx* ctx, nvb_dbatch* b, float* d_pcm, void* stream);
int     nvb_dbatch_result(nvb_ctx* ctx, nvb_dbatch* b, void* stream, nvb_result* res);   /* synchronises `stream`; has_clipped / n_floor_range
                                                                                            cover the runs since the previous call (then cleared) */
int     nvb_dbatch_destroy(nvb_ctx* ctx, nvb_dbatch* b);

/* Stage entry points (same kernels, exposed for parity tests and for the roofline measurement):
 *   nvb_dbatch_run_spectrum: residue + coupling + floor only; writes the dense spectrum
 *     [frame][channel][N/2] (ok frames in order; per-frame float offsets via nvb_dbatch_spec_offsets).
 *   nvb_dbatch_run_imdct: IMDCT + window + OLA + clip + interleave from a dense spectrum. */
int     nvb_dbatch_run_spectrum(nvb_ctx* ctx, nvb_dbatch* b, float* d_spectrum, void* stream);
int     nvb_dbatch_run_imdct(nvb_ctx* ctx, nvb_dbatch* b, const float* d_spectrum, float* d_pcm, void* stream);
int64_t nvb_dbatch_spectrum_floats(const nvb_dbatch* b);
int     nvb_dbatch_launches(const nvb_dbatch* b);   /* kernels enqueued by one nvb_dbatch_run */

#ifdef __cplusplus
}
#endif
#endif /* NVORBIS_B200_H */
