// nvb_host.h -- host-side logic behind the C ABI that needs no CUDA call: building the table blob
// from an nvb_setup, validating a blob, and turning an nvb_batch into the per-frame device plan
// (the overlap/emit bookkeeping of StreamDecoder.ReadNextPacket/Read, StreamDecoder.cs:320-463).
#pragma once
#include <string>
#include <vector>
#include "nvb_internal.h"

namespace nvb {

constexpr uint32_t BLOB_MAGIC = 0x3142564eu;   // "NVB1"

// Builds the device blob (BlobHeader + sections) from a caller-provided setup.  nvb_status.
int build_blob(const nvb_setup* s, std::vector<unsigned char>& blob, std::string& err);
// Structural validation of a blob received through nvb_setup_blob_import.  nvb_status.
int validate_blob(const void* data, size_t bytes, std::string& err);
// Resolves section pointers against `base` (host copy or device allocation of the same blob).
void resolve_setup(const unsigned char* base, const BlobHeader& h, DevSetup& S);

// Mode.CalcOverlap (Mode.cs:102-117) / GetPacketInfo (Mode.cs:119-151) for one (mode, window).
struct Overlap { int start, valid, total; };
Overlap nominal_overlap(const BlobHeader& h, int block_flag, int window);

// Decoder state that survives from one batch to the next (StreamDecoder._prevPacket*).
struct CarryState {
    bool have_prev = false;       // _prevPacketBuf != null
    int prev_start = 0;           // _prevPacketStart
    int prev_end = 0;             // _prevPacketEnd
    int prev_stop = 0;            // _prevPacketStop
    int prev_n = 0;               // block size of the carried block
};

struct Plan {
    std::vector<DevFrame> frames; // ok frames and tail drains, in order
    int64_t samples = 0;          // PCM samples per channel the batch emits
    int64_t spec_floats = 0;      // sum over ok frames of C * n/2
    int n_failed = 0, n_inconsistent = 0;
    int last_ok = -1;             // DevFrame index of the batch's last decoded block, -1 if none
    bool uses_carry = false;      // some DevFrame reads the previous batch's block
    bool uses_floor0 = false;     // some frame's mapping has a type 0 floor: nvb_batch.floor0 is read
    bool sequential = true;       // class / entry ranges of successive frames do not overlap and ascend (inputs can be uploaded in chunks)
    CarryState end_state;
};

// host_blob: the host copy of the blob.  nvb_status; on failure `err` names the offending frame.
int plan_batch(const unsigned char* host_blob, const nvb_batch* b, int flags, const CarryState& in, Plan& out, std::string& err);

}  // namespace nvb
