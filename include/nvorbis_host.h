/* nvorbis_host.h -- C ABI of the host-side half of the split decoder (libnvorbis_host.so).
 *
 * In a deployment the host is NVorbis itself (C#): it keeps VorbisReader / StreamDecoder, does the Ogg
 * paging and all bit unpacking, and P/Invokes include/nvorbis_b200.h with batches.  No .NET toolchain exists
 * in this image, so this library is the C++ stand-in for that host half: it produces exactly what the
 * reference's bit-reading code produces, as the plain arrays of an nvb_setup / nvb_batch:
 *
 *   container -> packets          Ogg/PageReaderBase.cs, Ogg/PageReader.cs, Ogg/PacketProvider.cs
 *   setup headers -> nvb_setup    StreamDecoder.cs:179-289, Codebook.cs:59-292, Floor1.cs:30-133,
 *                                 Residue0.cs:35-117, Residue2.cs:10-14, Mapping.cs:16-93, Mode.cs:24-67
 *   audio packet -> nvb_frame     StreamDecoder.cs:465-530, Mode.cs:119-151 (GetPacketInfo), Floor1.cs:135-184 /
 *                                 Floor0.cs:98-150 (Unpack), Mapping.cs:95-134 (ExecuteChannel / ForceEnergy), Residue0.cs:119-178
 *                                 (class words + VQ entry numbers, Codebook.DecodeScalar Codebook.cs:294-320)
 *   EOS trim of the last block    StreamDecoder.cs:429-437
 *
 * It never touches a GPU and never computes a sample: synthesis is include/nvorbis_b200.h.
 */
#ifndef NVORBIS_HOST_H
#define NVORBIS_HOST_H

#include <stddef.h>
#include <stdint.h>
#include "nvorbis_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nvh_stream nvh_stream;

typedef struct nvh_info {
    int32_t channels, sample_rate;
    int32_t block_size[2];
    int32_t n_books, n_floors, n_residues, n_mappings, n_modes;
    int32_t post_stride;           /* int16 per (frame, channel) in nvb_batch.posts: same rule as nvb_post_stride() */
    int64_t n_packets;             /* all packets of the logical stream, the three headers included */
    int64_t n_audio_packets;       /* packets after the setup header */
    int64_t last_granule;          /* granule position of the last packet that carries one, -1 if none */
    int32_t has_eos;               /* the container flagged an end-of-stream page that was kept (Ogg/StreamPageReader.cs:72-75) */
    int32_t floor0_stride;         /* floats per (frame, channel) in nvb_batch.floor0 (same rule as nvb_floor0_stride()), 0 = no type 0 floor */
} nvh_info;

/* nvh_open_ogg: `VorbisReader(Stream)` on an in-memory seekable stream, first logical stream only.
 * nvh_open_packets: `StreamDecoder(IPacketProvider)` on already-demuxed packets; flags bit0 = packet carries a
 * granule position, bit1 = IsEndOfStream, bit2 = IsResync.  Both parse the three header packets and fail with
 * NVB_ERR_DATA where the reference throws InvalidDataException. */
int nvh_open_ogg(const uint8_t* data, size_t len, nvh_stream** out);
int nvh_open_packets(const uint8_t* data, const int64_t* sizes, const int64_t* granules, const uint8_t* flags, int64_t n, nvh_stream** out);
/* Containers with several logical streams (multiplexed or chained; VorbisReader.Streams / SwitchStreams, VorbisReader.cs:96-150):
 * how many there are (by serial number, in order of appearance), and opening the one with a given index.  A stream that is
 * not Vorbis fails with NVB_ERR_DATA like any other bad header.  nvh_open_ogg = index 0. */
int nvh_ogg_stream_count(const uint8_t* data, size_t len);
int nvh_open_ogg_stream(const uint8_t* data, size_t len, int stream_index, nvh_stream** out);
/* Forward-only input (Ogg/ForwardOnlyPageReader.cs, ForwardOnlyPacketProvider.cs: a stream that cannot seek), in feed form:
 * nvh_open_forward creates an empty stream, nvh_feed appends the next container bytes (any chunking; end_of_input = 1 with
 * the last chunk) and returns the number of audio packets demuxed so far (0 until the three header packets are complete;
 * nvh_get_info / nvh_setup are valid from then on).  Pages and continued packets that are not complete yet wait for more
 * input; nvh_unpack / nvh_packet_batch hand out what is available and only report the end of the stream (and drain the last
 * tail) once the input has ended.  nvh_seek / nvh_rewind are not available on such a stream before its input has ended. */
int nvh_open_forward(nvh_stream** out);
int64_t nvh_feed(nvh_stream* s, const uint8_t* data, size_t len, int end_of_input);
int nvh_close(nvh_stream* s);
const char* nvh_last_error(nvh_stream* s);       /* s may be NULL: error of the last failed open on this thread */

int nvh_get_info(nvh_stream* s, nvh_info* out);
/* The parsed setup, ready for nvb_upload_setup.  Pointers stay valid until nvh_close. */
const nvb_setup* nvh_setup(nvh_stream* s);

/* Packet access (IPacketProvider view): size / flags / granule / bytes of packet i (headers included). */
int64_t nvh_packet_size(nvh_stream* s, int64_t i);
int nvh_packet_get(nvh_stream* s, int64_t i, uint8_t* dst, int64_t* granule, int32_t* flags);

/* Unpacks the next `count` audio packets (fewer at the end of the stream) with up to `threads` host threads
 * into stream-owned arrays described by *out (valid until the next nvh_unpack / nvh_close).  Packets are
 * independent given the setup, so they are unpacked in parallel; the position bookkeeping that feeds the EOS
 * trim runs afterwards in stream order.  When the provider runs dry one NVB_FRAME_FAILED record is appended
 * (GetNextPacket() == null drains the last tail, StreamDecoder.cs:352-356,476-480) and *end_of_stream = 1.
 * Returns the number of records produced (>= 0) or an nvb_status (< 0). */
int64_t nvh_unpack(nvh_stream* s, int64_t count, int threads, nvb_batch* out, int32_t* end_of_stream);
/* StreamDecoder.SeekTo(0) / ResetDecoder: restart at the first audio packet. */
int nvh_rewind(nvh_stream* s);

/* StreamDecoder.SeekTo(samplePosition) (StreamDecoder.cs:562-628) on the unpacking half: positions the packet cursor ONE packet
 * before the packet that holds the sample (the pre-roll packet: decoded for its tail only, as the first block of a stream
 * is, StreamDecoder.cs:446-450) and returns in *skip_samples how many samples per channel of the following output the caller
 * drops (the reference's `_prevPacketStart += rollForward`).  The caller resets the synthesis context (nvb_reset) and decodes
 * on without NVB_RUN_CONTINUE.  Sample positions count the samples a decode from the start of the stream emits (equal to the
 * granule positions of a well-formed stream).  Packet lengths come from the packets' first bits (GetPacketGranules,
 * StreamDecoder.cs:630-647): no Huffman decoding.  nvh_total_samples: what a full decode emits (call on a rewound stream). */
int nvh_seek(nvh_stream* s, int64_t sample_position, int64_t* skip_samples);
int64_t nvh_total_samples(nvh_stream* s);

/* ---- host half of the GPU-side packet unpack (include/nvorbis_b200.h: nvb_decode_packets) ----
 * nvh_unpack_tables: the unpack tables of this stream's setup for nvb_upload_unpack_tables (stream-owned, valid until
 * nvh_close).  NVB_ERR_UNSUPPORTED for setups the device unpacker does not cover (a type 0 floor).
 * nvh_packet_batch: the next `count` audio packets as an nvb_packet_batch -- raw bytes plus, per packet, what
 * Mode.GetPacketInfo reads from its first bits (Mode.cs:119-151) after the same stream-order bookkeeping as nvh_unpack
 * (sample position, end-of-stream trim, the drain record when the provider runs dry).  No Huffman decoding happens here. */
int nvh_unpack_tables(nvh_stream* s, const void** blob, size_t* bytes);
int64_t nvh_packet_batch(nvh_stream* s, int64_t count, nvb_packet_batch* out, int32_t* end_of_stream);

#ifdef __cplusplus
}
#endif
#endif /* NVORBIS_HOST_H */
