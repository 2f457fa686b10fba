// nvb_api.cu -- the C ABI of include/nvorbis_b200.h: contexts, setup upload, batch upload and launch.
// Device memory, streams and copies only; the kernels are in nvb_kernels.cu / nvb_fused.cu and the
// CUDA-free planning in nvb_host.cpp.  There is no CPU fallback: without a usable GPU nvb_create fails.
#if !defined(NVB_CPU_SHIM)
#include <cuda_runtime.h>
#endif
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "nvb_host.h"

using namespace nvb;

constexpr int NVB_CHUNKS = 4;            // frame ranges a large batch is pipelined in

struct nvb_ctx {
    int device = 0;
    std::string err;
    bool has_setup = false;
    std::vector<unsigned char> host_blob;
    unsigned char* d_blob = nullptr;
    BlobHeader H;
    DevSetup S;
    // overlap tail carried from one nvb_decode_batch to the next (StreamDecoder._prevPacketBuf)
    CarryState carry;
    float* d_carry[2] = {nullptr, nullptr};
    int carry_cur = 0;
    cudaStream_t stream = nullptr;
    // nvb_decode_batch_begin: ctx->stream carries the uploads, chunk_stream[0] the kernels, chunk_stream[1] the PCM read-back
    cudaStream_t chunk_stream[2] = {nullptr, nullptr};
    // nvb_decode_batch_begin/_end: up to two batches in flight, each with its own device staging; a batch's completion
    // (its last PCM read-back + the counters read-back) is one event on the read-back stream
    struct Slot {
        nvb_dbatch* staging = nullptr;
        float* d_pcm = nullptr; size_t pcm_cap = 0;
        int16_t* d_pcm16 = nullptr; size_t pcm16_cap = 0;     // NVB_RUN_PCM_S16 staging
        cudaEvent_t ev_up[NVB_CHUNKS] = {}, ev_k[NVB_CHUNKS] = {}, ev_all = nullptr;
        Counters* h_counters = nullptr;      // pinned
        DevFrame* h_frames = nullptr; size_t h_frames_cap = 0;   // pinned copy of the plan: its upload must not block the host
        int rc = NVB_OK;                     // failure detected while enqueuing (reported by _end)
    } slot[NVB_MAX_IN_FLIGHT];
    int head = 0, in_flight = 0;
    // GPU-side packet unpack: the unpack tables (nvb_upload_unpack_tables)
    unsigned char* d_utab = nullptr; bool has_unpack = false; nvbu::UHeader UH; UnpackTables UT;
};

struct nvb_dbatch {
    Plan plan;
    int flags = 0;
    bool fused = false;
    DevFrame* d_frames = nullptr; size_t cap_frames = 0;
    int16_t* d_posts = nullptr;   size_t cap_posts = 0;
    uint8_t* d_classes = nullptr; size_t cap_classes = 0;
    uint16_t* d_entries = nullptr; size_t cap_entries = 0;
    float* d_floor0 = nullptr;    size_t cap_floor0 = 0;
    float* d_spectrum = nullptr;  size_t cap_spectrum = 0;
    float* d_blocks = nullptr;    size_t cap_blocks = 0;
    Counters* d_counters = nullptr;
    int launches = 0;
    // raw packets of a packet batch (nvb_decode_packets) and the frames handed to the planner
    uint8_t* d_pkt = nullptr; size_t cap_pkt = 0;
    uint32_t* d_pkt_off = nullptr; size_t cap_pkt_off = 0;
    std::vector<nvb_frame> pkt_frames;
    bool from_packets = false;
};

namespace {

thread_local std::string g_err;

#if !defined(NVB_CPU_SHIM)
// NVB_TRACE=1: device-side timeline of nvb_decode_batch_begin/_end (timing events around every copy and kernel group),
// printed to stderr by _end relative to the first event ever recorded.  Debugging aid for the copy/compute pipeline.
struct Trace {
    struct Mark { cudaEvent_t ev; const char* what; int batch, chunk; };
    std::vector<Mark> marks; cudaEvent_t origin = nullptr; int batch = 0;
    static bool on() { static const bool v = std::getenv("NVB_TRACE") != nullptr; return v; }
    void mark(cudaStream_t st, const char* what, int chunk) {
        if (!on()) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
        if (!origin) origin = e;
        marks.push_back({e, what, batch, chunk});
    }
    void dump(int upto_batch) {
        if (!on()) return;
        size_t keep = 0;
        for (size_t i = 0; i < marks.size(); i++) {
            if (marks[i].batch > upto_batch) { marks[keep++] = marks[i]; continue; }
            float ms = 0.f; cudaEventElapsedTime(&ms, origin, marks[i].ev);
            std::fprintf(stderr, "[nvb trace] batch %d chunk %d %-12s %9.3f ms\n", marks[i].batch, marks[i].chunk, marks[i].what, ms);
            if (marks[i].ev != origin) cudaEventDestroy(marks[i].ev);
        }
        marks.resize(keep);
    }
};
Trace g_trace;
#define NVB_TRACE_MARK(st, what, chunk) g_trace.mark((st), (what), (chunk))
#else
#define NVB_TRACE_MARK(st, what, chunk) ((void)0)
#endif

int set_err(nvb_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_err = msg;
    return code;
}
int cuda_fail(nvb_ctx* ctx, cudaError_t e, const char* what) {
    char tmp[256];
    std::snprintf(tmp, sizeof tmp, "%s: %s", what, cudaGetErrorString(e));
    return set_err(ctx, NVB_ERR_CUDA, tmp);
}
#define NVB_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail((ctx), e__, #call); } while (0)

struct DeviceGuard {
    int prev = -1; bool ok = false;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <class T> int grow(nvb_ctx* ctx, T*& p, size_t& cap, size_t need) {
    if (need <= cap && p) return NVB_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    // headroom: the batches of a stream differ by a few per cent, and a reallocation (cudaFree waits for the device) in the
    // middle of a pipelined run costs milliseconds
    size_t n = need + need / 4 + 64;
    cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
    if (e != cudaSuccess) { cudaGetLastError(); n = need ? need : 1; e = cudaMalloc((void**)&p, n * sizeof(T)); }
    if (e != cudaSuccess) { p = nullptr; cudaGetLastError(); return set_err(ctx, NVB_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    cap = n;
    return NVB_OK;
}

void free_dbatch(nvb_dbatch* b) {
    if (!b) return;
    cudaFree(b->d_frames); cudaFree(b->d_posts); cudaFree(b->d_classes); cudaFree(b->d_entries); cudaFree(b->d_floor0);
    cudaFree(b->d_spectrum); cudaFree(b->d_blocks); cudaFree(b->d_counters); cudaFree(b->d_pkt); cudaFree(b->d_pkt_off);
    delete b;
}

int install_blob(nvb_ctx* ctx, std::vector<unsigned char>&& blob) {
    DeviceGuard g(ctx->device);
    unsigned char* d = nullptr;
    cudaError_t e = cudaMalloc((void**)&d, blob.size());
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(ctx, NVB_ERR_NOMEM, std::string("cudaMalloc(blob): ") + cudaGetErrorString(e)); }
    e = cudaMemcpy(d, blob.data(), blob.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); return cuda_fail(ctx, e, "cudaMemcpy(blob)"); }
    BlobHeader h; std::memcpy(&h, blob.data(), sizeof h);
    float* carry[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++) {
        e = cudaMalloc((void**)&carry[i], sizeof(float) * (size_t)h.channels * h.bs[1]);
        if (e == cudaSuccess) e = cudaMemset(carry[i], 0, sizeof(float) * (size_t)h.channels * h.bs[1]);
        if (e != cudaSuccess) { cudaFree(d); cudaFree(carry[0]); cudaFree(carry[1]); cudaGetLastError(); return set_err(ctx, NVB_ERR_NOMEM, "cudaMalloc(carry)"); }
    }
    if (ctx->d_blob) cudaFree(ctx->d_blob);
    cudaFree(ctx->d_carry[0]); cudaFree(ctx->d_carry[1]);
    ctx->d_blob = d; ctx->d_carry[0] = carry[0]; ctx->d_carry[1] = carry[1]; ctx->carry_cur = 0;
    ctx->host_blob = std::move(blob);
    ctx->H = h;
    resolve_setup(ctx->d_blob, ctx->H, ctx->S);
    ctx->carry = CarryState();
    ctx->has_setup = true;
    return NVB_OK;
}

// Uploads a batch into `b` (buffers grow as needed) and plans it.
int upload_batch(nvb_ctx* ctx, nvb_dbatch* b, const nvb_batch* batch, int flags, cudaStream_t st, bool* defer_inputs = nullptr,
                 DevFrame** pinned = nullptr, size_t* pinned_cap = nullptr) {
    std::string err;
    int rc = plan_batch(ctx->host_blob.data(), batch, flags, ctx->carry, b->plan, err);
    if (rc != NVB_OK) return set_err(ctx, rc, err);
    b->flags = flags;
    b->fused = !(flags & NVB_RUN_EXACT) && fused_supported(ctx->H, b->plan.frames.data(), (int)b->plan.frames.size());
    const size_t nf = b->plan.frames.size();
    const size_t n_posts = (size_t)batch->n_frames * ctx->H.channels * ctx->H.post_stride;
    if ((rc = grow(ctx, b->d_frames, b->cap_frames, nf)) != NVB_OK) return rc;
    if ((rc = grow(ctx, b->d_posts, b->cap_posts, n_posts)) != NVB_OK) return rc;
    if ((rc = grow(ctx, b->d_classes, b->cap_classes, (size_t)batch->n_classes)) != NVB_OK) return rc;
    if ((rc = grow(ctx, b->d_entries, b->cap_entries, (size_t)batch->n_entries)) != NVB_OK) return rc;
    const size_t n_floor0 = b->plan.uses_floor0 ? (size_t)batch->n_frames * ctx->H.channels * ctx->H.f0_stride : 0;
    if (n_floor0 && (rc = grow(ctx, b->d_floor0, b->cap_floor0, n_floor0)) != NVB_OK) return rc;
    if ((rc = grow(ctx, b->d_spectrum, b->cap_spectrum, (size_t)b->plan.spec_floats)) != NVB_OK) return rc;
    if (!b->fused && (rc = grow(ctx, b->d_blocks, b->cap_blocks, 2 * (size_t)b->plan.spec_floats)) != NVB_OK) return rc;
    if (!b->d_counters) {
        size_t cap = 0;
        if ((rc = grow(ctx, b->d_counters, cap, 1)) != NVB_OK) return rc;
        NVB_CUDA(ctx, cudaMemsetAsync(b->d_counters, 0, sizeof(Counters), st));
    }
    const DevFrame* plan_src = b->plan.frames.data();
    if (pinned && nf) {                                                     // page-locked staging: the copy is truly asynchronous
        if (*pinned_cap < nf) {
            if (*pinned) { cudaFreeHost(*pinned); *pinned = nullptr; *pinned_cap = 0; }
            if (cudaHostAlloc((void**)pinned, (nf + nf / 4 + 64) * sizeof(DevFrame), cudaHostAllocDefault) == cudaSuccess) *pinned_cap = nf + nf / 4 + 64;
            else { cudaGetLastError(); *pinned = nullptr; }
        }
        if (*pinned) { std::memcpy(*pinned, b->plan.frames.data(), nf * sizeof(DevFrame)); plan_src = *pinned; }
    }
    if (nf) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_frames, plan_src, nf * sizeof(DevFrame), cudaMemcpyHostToDevice, st));
    if (n_floor0) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_floor0, batch->floor0, n_floor0 * sizeof(float), cudaMemcpyHostToDevice, st));
    if (defer_inputs) {
        // nvb_decode_batch uploads posts / classes / entries chunk by chunk when the batch is laid out sequentially
        *defer_inputs = *defer_inputs && b->fused && b->plan.sequential;
        if (*defer_inputs) return NVB_OK;
    }
    if (n_posts) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_posts, batch->posts, n_posts * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    if (batch->n_classes) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_classes, batch->classes, (size_t)batch->n_classes, cudaMemcpyHostToDevice, st));
    if (batch->n_entries) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_entries, batch->entries, (size_t)batch->n_entries * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    return NVB_OK;
}

// A packet batch (nvb_decode_packets): plans it from the host-provided packet headers, uploads the raw packets, sizes the
// device-side record buffers at the tables' fixed per-frame strides.  k_unpack (enqueue_unpack) fills them.
int upload_packets(nvb_ctx* ctx, nvb_dbatch* b, const nvb_packet_batch* pb, int flags, cudaStream_t st, DevFrame** pinned, size_t* pinned_cap) {
    if (!ctx->has_unpack) return set_err(ctx, NVB_ERR_STATE, "no unpack tables uploaded (nvb_upload_unpack_tables)");
    if (!pb || pb->n_packets < 0 || (pb->n_packets > 0 && (!pb->frames || !pb->offsets || !pb->data))) return set_err(ctx, NVB_ERR_ARG, "packet batch is NULL / has NULL arrays");
    const size_t n = (size_t)pb->n_packets;
    const size_t cs = (size_t)ctx->UH.cls_stride, es = (size_t)ctx->UH.ent_stride;
    if (n * cs > 0xffffffffull || n * es > 0xffffffffull) return set_err(ctx, NVB_ERR_ARG, "packet batch too large for 32-bit record offsets: split it");
    for (size_t i = 0; i < n; i++) if (pb->offsets[i + 1] < pb->offsets[i]) return set_err(ctx, NVB_ERR_DATA, "packet offsets must ascend");
    b->pkt_frames.assign(pb->frames, pb->frames + n);
    for (size_t i = 0; i < n; i++) {                                        // device-produced fields: placeholders with the fixed strides
        nvb_frame& f = b->pkt_frames[i];
        f.exec_mask = 0; f.res_decoded = 1; f.entry_count = 0;
        f.classes_off = (uint32_t)(i * cs); f.entries_off = (uint32_t)(i * es);
    }
    static const int16_t dummy_posts = 0;
    nvb_batch tmp; std::memset(&tmp, 0, sizeof tmp);
    tmp.n_frames = pb->n_packets; tmp.frames = b->pkt_frames.data(); tmp.posts = &dummy_posts;
    static const uint8_t dummy_c = 0; static const uint16_t dummy_e = 0; static const float dummy_f = 0.f;
    tmp.floor0 = &dummy_f;                                                  // (type 0 floor records are produced on the device: plan_batch only asks for a pointer)
    tmp.classes = &dummy_c; tmp.n_classes = (int64_t)(n * cs); tmp.entries = &dummy_e; tmp.n_entries = (int64_t)(n * es);
    std::string err;
    int rc = plan_batch(ctx->host_blob.data(), &tmp, flags, ctx->carry, b->plan, err);
    if (rc != NVB_OK) return set_err(ctx, rc, err);
    b->flags = flags; b->from_packets = true;
    b->fused = !(flags & NVB_RUN_EXACT) && fused_supported(ctx->H, b->plan.frames.data(), (int)b->plan.frames.size());
    const size_t nf = b->plan.frames.size();
    const size_t n_posts = n * ctx->H.channels * ctx->H.post_stride;
    const size_t bytes = n ? pb->offsets[n] : 0;
    if ((rc = grow(ctx, b->d_frames, b->cap_frames, nf)) != NVB_OK) return rc;
    if ((rc = grow(ctx, b->d_posts, b->cap_posts, n_posts)) != NVB_OK) return rc;
    if ((rc = grow(ctx, b->d_classes, b->cap_classes, n * cs)) != NVB_OK) return rc;
    if ((rc = grow(ctx, b->d_entries, b->cap_entries, n * es)) != NVB_OK) return rc;
    if (ctx->H.f0_stride > 0 && (rc = grow(ctx, b->d_floor0, b->cap_floor0, n * ctx->H.channels * ctx->H.f0_stride)) != NVB_OK) return rc;   // type 0 floor records, written by k_unpack
    if ((rc = grow(ctx, b->d_pkt, b->cap_pkt, bytes + 16)) != NVB_OK) return rc;
    if ((rc = grow(ctx, b->d_pkt_off, b->cap_pkt_off, n + 1)) != NVB_OK) return rc;
    if ((rc = grow(ctx, b->d_spectrum, b->cap_spectrum, (size_t)b->plan.spec_floats)) != NVB_OK) return rc;
    if (!b->fused && (rc = grow(ctx, b->d_blocks, b->cap_blocks, 2 * (size_t)b->plan.spec_floats)) != NVB_OK) return rc;
    if (!b->d_counters) {
        size_t cap = 0;
        if ((rc = grow(ctx, b->d_counters, cap, 1)) != NVB_OK) return rc;
        NVB_CUDA(ctx, cudaMemsetAsync(b->d_counters, 0, sizeof(Counters), st));
    }
    const DevFrame* plan_src = b->plan.frames.data();
    if (pinned && nf) {
        if (*pinned_cap < nf) {
            if (*pinned) { cudaFreeHost(*pinned); *pinned = nullptr; *pinned_cap = 0; }
            if (cudaHostAlloc((void**)pinned, (nf + nf / 4 + 64) * sizeof(DevFrame), cudaHostAllocDefault) == cudaSuccess) *pinned_cap = nf + nf / 4 + 64;
            else { cudaGetLastError(); *pinned = nullptr; }
        }
        if (*pinned) { std::memcpy(*pinned, b->plan.frames.data(), nf * sizeof(DevFrame)); plan_src = *pinned; }
    }
    if (nf) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_frames, plan_src, nf * sizeof(DevFrame), cudaMemcpyHostToDevice, st));
    if (bytes) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_pkt, pb->data, bytes, cudaMemcpyHostToDevice, st));
    NVB_CUDA(ctx, cudaMemsetAsync(b->d_pkt + bytes, 0, 16, st));             // the bit cursor loads whole aligned words
    NVB_CUDA(ctx, cudaMemcpyAsync(b->d_pkt_off, pb->offsets, (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    return NVB_OK;
}

int enqueue_unpack(nvb_ctx* ctx, nvb_dbatch* b, cudaStream_t st, int frame_lo, int frame_cnt) {
    UnpackArgs u;
    u.T = ctx->UT; u.frames = b->d_frames; u.frame_lo = frame_lo; u.n_frames = frame_cnt;
    u.data = b->d_pkt; u.offsets = b->d_pkt_off; u.posts = b->d_posts; u.classes = b->d_classes; u.entries = b->d_entries;
    u.floor0 = ctx->H.f0_stride > 0 ? b->d_floor0 : nullptr; u.f0_stride = ctx->H.f0_stride; u.dbooks = ctx->S.books; u.vq = ctx->S.vq;
    const int r = launch_unpack(u, st);
    if (r < 0) return cuda_fail(ctx, cudaGetLastError(), "k_unpack launch");
    return NVB_OK;
}

LaunchArgs make_args(nvb_ctx* ctx, nvb_dbatch* b, float* spectrum, float* d_pcm, bool save_carry) {
    LaunchArgs a;
    a.S = ctx->S;
    a.frames = b->d_frames; a.frame_lo = 0; a.n_frames = (int)b->plan.frames.size();
    a.posts = b->d_posts; a.classes = b->d_classes; a.entries = b->d_entries;
    a.floor0 = b->plan.uses_floor0 ? b->d_floor0 : nullptr;
    a.spectrum = spectrum;
    a.blocks = b->d_blocks;
    a.carry_in = ctx->d_carry[ctx->carry_cur];
    a.carry_out = save_carry ? ctx->d_carry[ctx->carry_cur ^ 1] : nullptr;
    a.carry_frame = save_carry ? b->plan.last_ok : -1;
    a.pcm = d_pcm;
    a.counters = b->d_counters;
    a.clip = (b->flags & NVB_RUN_NO_CLIP) ? 0 : 1;
    a.inputs_from_kernel = b->from_packets ? 1 : 0;
    return a;
}

// Enqueues the synthesis of an uploaded batch.  stage: 0 = all, 1 = spectrum only, 2 = IMDCT.. only.
// `spectrum` = dense spectrum buffer written by stage 1 and read by stage 2 (nullptr: the batch's own).
int enqueue(nvb_ctx* ctx, nvb_dbatch* b, int stage, float* spectrum, float* d_pcm, bool save_carry, cudaStream_t st,
            int frame_lo = 0, int frame_cnt = -1, bool reset_counters = true, cudaEvent_t after_spectrum = nullptr, cudaEvent_t before_synth = nullptr) {
    if (b->plan.frames.empty()) { b->launches = 0; return NVB_OK; }
    static const int no_pdl = std::getenv("NVB_NO_PDL") ? std::atoi(std::getenv("NVB_NO_PDL")) : 0;      // experiment hook: 1 = no programmatic early start at all
    NvbNoEarlyStart no_early(no_pdl == 1);
    LaunchArgs a = make_args(ctx, b, spectrum ? spectrum : b->d_spectrum, d_pcm, save_carry);
    if (frame_cnt >= 0) { a.frame_lo = frame_lo; a.n_frames = frame_cnt; }
    int launches = 0, r;
    // the counters are cleared when they are read (fetch_result) or per batch (nvb_decode_batch_begin), not per run: a
    // memset between the kernels of consecutive runs would serialise what programmatic dependent launch overlaps
    (void)reset_counters;
    // one-kernel synthesis (NVB_RUN_ONE_KERNEL / NVB_ONE_KERNEL=1): records -> PCM in one launch where the setup is covered
    // (default: the library chooses per launch -- one kernel where the launch fits one round of the CTAs' warps, see launch_synth_fused)
    static const char* ok_env = std::getenv("NVB_ONE_KERNEL");
    const bool forced = ok_env ? std::atoi(ok_env) != 0 : (b->flags & NVB_RUN_ONE_KERNEL) != 0;
    const bool never = ok_env ? std::atoi(ok_env) == 0 : (b->flags & NVB_RUN_TWO_KERNELS) != 0;
    if (stage == 0 && b->fused && !never) {
        r = launch_synth_fused(a, b->plan.frames.data(), st, forced);
        if (r == -1) return cuda_fail(ctx, cudaGetLastError(), "k_imdct_fused<SYN> launch");
        if (r >= 0) {
            if (after_spectrum) NVB_CUDA(ctx, cudaEventRecord(after_spectrum, st));
            b->launches = (frame_cnt >= 0 && frame_lo > 0) ? b->launches + r : r;
            return NVB_OK;
        }
    }
    if (stage != 2) {
        NvbNoEarlyStart spec_no_early(no_pdl == 2);                           // 2: the spectrum kernel waits for everything before it
        if ((r = launch_spectrum(a, st)) < 0) return cuda_fail(ctx, cudaGetLastError(), "k_spectrum launch");
        launches += r;
        if (after_spectrum) NVB_CUDA(ctx, cudaEventRecord(after_spectrum, st));
    }
    if (stage != 1) {
        if (before_synth) NVB_CUDA(ctx, cudaStreamWaitEvent(st, before_synth, 0));     // the halo block's spectrum comes from the previous chunk
        if (b->fused) {
            NvbNoEarlyStart imdct_no_early(no_pdl == 3);                      // 3: the IMDCT kernel starts only after the spectrum kernel has ended
            if ((r = launch_imdct_fused(a, b->plan.frames.data(), st)) < 0) return cuda_fail(ctx, cudaGetLastError(), "k_imdct_fused launch");
            launches += r;
        } else {
            if ((r = launch_imdct_exact(a, st)) < 0) return cuda_fail(ctx, cudaGetLastError(), "k_imdct_exact launch");
            launches += r;
            if ((r = launch_ola(a, st)) < 0) return cuda_fail(ctx, cudaGetLastError(), "k_ola launch");
            launches += r;
            if (save_carry && b->plan.last_ok >= 0) {
                const DevFrame& lf = b->plan.frames[(size_t)b->plan.last_ok];
                NVB_CUDA(ctx, cudaMemcpy2DAsync(a.carry_out, sizeof(float) * ctx->H.bs[1], b->d_blocks + 2 * (size_t)lf.spec_off, sizeof(float) * lf.n,
                                                sizeof(float) * lf.n, (size_t)ctx->H.channels, cudaMemcpyDeviceToDevice, st));
            }
        }
    }
    b->launches = (frame_cnt >= 0 && frame_lo > 0) ? b->launches + launches : launches;
    return NVB_OK;
}

int fetch_result(nvb_ctx* ctx, nvb_dbatch* b, cudaStream_t st, nvb_result* res) {
    Counters c; std::memset(&c, 0, sizeof c);
    if (!b->plan.frames.empty()) {
        NVB_CUDA(ctx, cudaMemcpyAsync(&c, b->d_counters, sizeof c, cudaMemcpyDeviceToHost, st));
        NVB_CUDA(ctx, cudaMemsetAsync(b->d_counters, 0, sizeof(Counters), st));
    }
    NVB_CUDA(ctx, cudaStreamSynchronize(st));
    if (res) {
        res->samples_per_channel = b->plan.samples;
        res->has_clipped = c.clipped ? 1 : 0;
        res->n_failed = b->plan.n_failed;
        res->n_floor_range = c.floor_range;
        res->n_inconsistent = b->plan.n_inconsistent;
    }
    if (c.bad_entry) return set_err(ctx, NVB_ERR_DATA, "a VQ entry number is outside its codebook (Codebook.cs:322 would throw)");
    return NVB_OK;
}

}  // namespace

extern "C" {

int nvb_abi_version(void) { return NVB_ABI_VERSION; }

const char* nvb_strerror(int status) {
    switch (status) {
        case NVB_OK: return "ok";
        case NVB_ERR_ARG: return "invalid argument";
        case NVB_ERR_CUDA: return "CUDA failure";
        case NVB_ERR_UNSUPPORTED: return "setup outside the supported envelope";
        case NVB_ERR_NOMEM: return "out of memory";
        case NVB_ERR_STATE: return "invalid call order";
        case NVB_ERR_CAPACITY: return "output buffer too small";
        case NVB_ERR_DATA: return "malformed data";
        default: return "unknown status";
    }
}

const char* nvb_last_error(nvb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int nvb_create(int device, nvb_ctx** out) {
    if (!out) return set_err(nullptr, NVB_ERR_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(nullptr, e, "cudaGetDeviceCount (no CPU fallback exists)"); }
    if (device < 0 || device >= count) return set_err(nullptr, NVB_ERR_ARG, "device index out of range");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10) return set_err(nullptr, NVB_ERR_CUDA, "device is not sm_100-class: this library carries sm_100a code only");
    nvb_ctx* ctx = new (std::nothrow) nvb_ctx();
    if (!ctx) return set_err(nullptr, NVB_ERR_NOMEM, "host allocation failed");
    ctx->device = device;
    DeviceGuard g(device);
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaStreamCreateWithFlags(&ctx->chunk_stream[i], cudaStreamNonBlocking);
    for (int i = 0; i < NVB_MAX_IN_FLIGHT && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&ctx->slot[i].ev_all, cudaEventDisableTiming);
        for (int k = 0; k < NVB_CHUNKS && e == cudaSuccess; k++) {
            e = cudaEventCreateWithFlags(&ctx->slot[i].ev_up[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->slot[i].ev_k[k], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->slot[i].h_counters, sizeof(Counters), cudaHostAllocDefault);
    }
    if (e != cudaSuccess) { delete ctx; return cuda_fail(nullptr, e, "cudaStreamCreate / cudaEventCreate"); }
    *out = ctx;
    return NVB_OK;
}

int nvb_destroy(nvb_ctx* ctx) {
    if (!ctx) return NVB_OK;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->d_blob); cudaFree(ctx->d_carry[0]); cudaFree(ctx->d_carry[1]); cudaFree(ctx->d_utab);
    cudaStreamDestroy(ctx->stream);
    for (int i = 0; i < 2; i++) cudaStreamDestroy(ctx->chunk_stream[i]);
    for (int i = 0; i < NVB_MAX_IN_FLIGHT; i++) {
        free_dbatch(ctx->slot[i].staging); cudaFree(ctx->slot[i].d_pcm); cudaFree(ctx->slot[i].d_pcm16);
        cudaEventDestroy(ctx->slot[i].ev_all);
        for (int k = 0; k < NVB_CHUNKS; k++) { cudaEventDestroy(ctx->slot[i].ev_up[k]); cudaEventDestroy(ctx->slot[i].ev_k[k]); }
        if (ctx->slot[i].h_counters) cudaFreeHost(ctx->slot[i].h_counters);
        if (ctx->slot[i].h_frames) cudaFreeHost(ctx->slot[i].h_frames);
    }
    delete ctx;
    return NVB_OK;
}

int nvb_host_alloc(size_t bytes, void** out) {
    if (!out) return set_err(nullptr, NVB_ERR_ARG, "out is NULL");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); *out = nullptr; return cuda_fail(nullptr, e, "cudaHostAlloc"); }
    return NVB_OK;
}
int nvb_host_free(void* p) {
    if (!p) return NVB_OK;
    cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFreeHost");
    return NVB_OK;
}

int nvb_upload_setup(nvb_ctx* ctx, const nvb_setup* setup) {
    if (!ctx) return set_err(nullptr, NVB_ERR_ARG, "ctx is NULL");
    if (ctx->in_flight > 0) return set_err(ctx, NVB_ERR_STATE, "batches in flight");
    std::vector<unsigned char> blob; std::string err;
    int rc = build_blob(setup, blob, err);
    if (rc != NVB_OK) return set_err(ctx, rc, err);
    return install_blob(ctx, std::move(blob));
}

int nvb_setup_blob_size(nvb_ctx* ctx, size_t* bytes) {
    if (!ctx || !bytes) return set_err(ctx, NVB_ERR_ARG, "NULL argument");
    if (!ctx->has_setup) return set_err(ctx, NVB_ERR_STATE, "no setup uploaded");
    *bytes = ctx->host_blob.size();
    return NVB_OK;
}
int nvb_setup_blob_export(nvb_ctx* ctx, void* dst, size_t bytes) {
    if (!ctx || !dst) return set_err(ctx, NVB_ERR_ARG, "NULL argument");
    if (!ctx->has_setup) return set_err(ctx, NVB_ERR_STATE, "no setup uploaded");
    if (bytes < ctx->host_blob.size()) return set_err(ctx, NVB_ERR_CAPACITY, "blob buffer too small");
    std::memcpy(dst, ctx->host_blob.data(), ctx->host_blob.size());
    return NVB_OK;
}
int nvb_setup_blob_import(nvb_ctx* ctx, const void* src, size_t bytes) {
    if (!ctx) return set_err(nullptr, NVB_ERR_ARG, "ctx is NULL");
    if (ctx->in_flight > 0) return set_err(ctx, NVB_ERR_STATE, "batches in flight");
    std::string err;
    int rc = validate_blob(src, bytes, err);
    if (rc != NVB_OK) return set_err(ctx, rc, err);
    std::vector<unsigned char> blob((const unsigned char*)src, (const unsigned char*)src + bytes);
    return install_blob(ctx, std::move(blob));
}

int nvb_post_stride(nvb_ctx* ctx) {
    if (!ctx || !ctx->has_setup) return set_err(ctx, NVB_ERR_STATE, "no setup uploaded");
    return ctx->H.post_stride;
}

int nvb_floor0_stride(nvb_ctx* ctx) {
    if (!ctx || !ctx->has_setup) return set_err(ctx, NVB_ERR_STATE, "no setup uploaded");
    return ctx->H.f0_stride;
}

int nvb_reset(nvb_ctx* ctx) {
    if (!ctx) return set_err(nullptr, NVB_ERR_ARG, "ctx is NULL");
    if (ctx->in_flight > 0) return set_err(ctx, NVB_ERR_STATE, "batches in flight");
    ctx->carry = CarryState();
    return NVB_OK;
}

static int decode_begin_impl(nvb_ctx* ctx, const nvb_batch* batch, const nvb_packet_batch* pbatch, int flags, float* pcm_out, size_t pcm_cap);

int nvb_decode_batch_begin(nvb_ctx* ctx, const nvb_batch* batch, int flags, float* pcm_out, size_t pcm_cap) {
    return decode_begin_impl(ctx, batch, nullptr, flags, pcm_out, pcm_cap);
}
int nvb_decode_packets_begin(nvb_ctx* ctx, const nvb_packet_batch* batch, int flags, float* pcm_out, size_t pcm_cap) {
    if (!batch) return set_err(ctx, NVB_ERR_ARG, "packet batch is NULL");
    return decode_begin_impl(ctx, nullptr, batch, flags, pcm_out, pcm_cap);
}
int nvb_decode_packets(nvb_ctx* ctx, const nvb_packet_batch* batch, int flags, float* pcm_out, size_t pcm_cap, nvb_result* res) {
    if (ctx && ctx->in_flight > 0) return set_err(ctx, NVB_ERR_STATE, "batches begun with nvb_decode_batch_begin are still in flight");
    const int rc = nvb_decode_packets_begin(ctx, batch, flags, pcm_out, pcm_cap);
    if (rc != NVB_OK) return rc;
    return nvb_decode_batch_end(ctx, res);
}

static int decode_begin_impl(nvb_ctx* ctx, const nvb_batch* batch, const nvb_packet_batch* pbatch, int flags, float* pcm_out, size_t pcm_cap) {
    if (!ctx) return set_err(nullptr, NVB_ERR_ARG, "ctx is NULL");
    if (!ctx->has_setup) return set_err(ctx, NVB_ERR_STATE, "no setup uploaded");
    if (ctx->in_flight >= NVB_MAX_IN_FLIGHT) return set_err(ctx, NVB_ERR_STATE, "NVB_MAX_IN_FLIGHT batches are already in flight: call nvb_decode_batch_end first");
    DeviceGuard g(ctx->device);
    nvb_ctx::Slot& sl = ctx->slot[(ctx->head + ctx->in_flight) % NVB_MAX_IN_FLIGHT];
    if (!sl.staging) { sl.staging = new (std::nothrow) nvb_dbatch(); if (!sl.staging) return set_err(ctx, NVB_ERR_NOMEM, "host allocation failed"); }
    nvb_dbatch* b = sl.staging;
    static const int chunk_min = std::getenv("NVB_CHUNK_MIN") ? std::atoi(std::getenv("NVB_CHUNK_MIN")) : 1024;   // test hook
    // Three in-order streams, one per engine: uploads (ctx->stream), kernels, PCM read-back.  A large batch is cut into
    // chunks of frames; chunk k's kernels wait for its inputs, its read-back for its kernels, so the PCM of chunk k crosses
    // PCIe while chunk k+1 computes and the next batch's inputs go up.  Kernels of consecutive chunks and batches run in
    // stream order, which is all the overlap tail (halo block, carried block) needs.
    cudaStream_t st_up = ctx->stream, st_k = ctx->chunk_stream[0], st_down = ctx->chunk_stream[1];
#if !defined(NVB_CPU_SHIM)
    g_trace.batch++;
#endif
    NVB_TRACE_MARK(st_up, "begin", -1);
    bool chunked_inputs = batch && batch->n_frames >= chunk_min && batch->n_frames >= 8;
    int rc;
    if (pbatch) { chunked_inputs = false; rc = upload_packets(ctx, b, pbatch, flags, st_up, &sl.h_frames, &sl.h_frames_cap); }
    else { b->from_packets = false; rc = upload_batch(ctx, b, batch, flags, st_up, &chunked_inputs, &sl.h_frames, &sl.h_frames_cap); }
    if (rc != NVB_OK) { cudaStreamSynchronize(st_up); return rc; }
    const size_t n_out = (size_t)b->plan.samples * ctx->H.channels;
    if (n_out > pcm_cap || (n_out > 0 && !pcm_out)) { cudaStreamSynchronize(st_up); return set_err(ctx, NVB_ERR_CAPACITY, "pcm_out too small for the batch"); }
    // output form: float or 16-bit PCM (NVB_RUN_PCM_S16), copied back to the host or left in the caller's device buffer (NVB_RUN_DEVICE_OUT)
    const bool s16 = (flags & NVB_RUN_PCM_S16) != 0, dev_out = (flags & NVB_RUN_DEVICE_OUT) != 0;
    if (dev_out && n_out > 0) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, pcm_out) != cudaSuccess || pa.type != cudaMemoryTypeDevice || pa.device != ctx->device || (reinterpret_cast<uintptr_t>(pcm_out) & 15)) {
            cudaGetLastError(); cudaStreamSynchronize(st_up);
            return set_err(ctx, NVB_ERR_ARG, "NVB_RUN_DEVICE_OUT: pcm_out must be a 16-byte aligned device pointer on the context's GPU");
        }
    }
    float* d_float = (dev_out && !s16) ? pcm_out : nullptr;                 // where the kernels write float PCM
    if (!d_float) { if ((rc = grow(ctx, sl.d_pcm, sl.pcm_cap, n_out)) != NVB_OK) { cudaStreamSynchronize(st_up); return rc; } d_float = sl.d_pcm; }
    int16_t* d_s16 = nullptr;
    if (s16) {
        if (dev_out) d_s16 = reinterpret_cast<int16_t*>(pcm_out);
        else { if ((rc = grow(ctx, sl.d_pcm16, sl.pcm16_cap, n_out)) != NVB_OK) { cudaStreamSynchronize(st_up); return rc; } d_s16 = sl.d_pcm16; }
    }
    const int nf = (int)b->plan.frames.size();
    // How many chunks: four when this is the only batch in flight (the synchronous nvb_decode_batch: chunk k's PCM crosses PCIe while
    // chunk k + 1 computes).  When the caller keeps a batch in flight, the read-back of the batch before this one already covers this
    // batch's upload and kernels, and every extra copy costs: one chunk for float PCM to the host (4096 stereo frames: 0.609 vs
    // 0.646 ms per step, 6.73 vs 6.34 M frames/s), two for 16-bit PCM (its read-back is short enough to need the finer overlap: 11.3 vs
    // 10.0 M), one when the PCM stays on the device (then the host side of _begin is the bound, and every chunk is a handful of driver
    // calls: 24.1 vs 23.3 M with two batches in flight, 29.7 M with three).  NVB_N_CHUNKS overrides (profiles/e2e_quick.py).
    static const int chunks_cfg = std::getenv("NVB_N_CHUNKS") ? std::atoi(std::getenv("NVB_N_CHUNKS")) : 0;
    int chunks_max = NVB_CHUNKS;
    if (ctx->in_flight >= 1) chunks_max = (s16 && !dev_out) ? 2 : 1;
    if (chunks_cfg >= 1) chunks_max = chunks_cfg > NVB_CHUNKS ? NVB_CHUNKS : chunks_cfg;
    const int n_chunks = ((chunked_inputs || pbatch) && nf >= chunk_min && nf >= 8) ? chunks_max : 1;
    if (n_chunks == 1 && chunked_inputs) {                                  // few decoded frames after all: upload everything now
        const size_t n_posts = (size_t)batch->n_frames * ctx->H.channels * ctx->H.post_stride;
        if (n_posts) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_posts, batch->posts, n_posts * sizeof(int16_t), cudaMemcpyHostToDevice, st_up));
        if (batch->n_classes) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_classes, batch->classes, (size_t)batch->n_classes, cudaMemcpyHostToDevice, st_up));
        if (batch->n_entries) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_entries, batch->entries, (size_t)batch->n_entries * sizeof(uint16_t), cudaMemcpyHostToDevice, st_up));
        chunked_inputs = false;
    }
    NVB_CUDA(ctx, cudaMemsetAsync(b->d_counters, 0, sizeof(Counters), st_k));
    const size_t C = (size_t)ctx->H.channels;
    for (int k = 0; k < n_chunks; k++) {
        const int lo = (int)((long long)nf * k / n_chunks), hi = (int)((long long)nf * (k + 1) / n_chunks);
        NVB_TRACE_MARK(st_up, "h2d_begin", k);
        if (chunked_inputs) {
            // this chunk's share of the inputs: the api frames from the chunk's first decoded block up to the next chunk's
            const int a0 = k == 0 ? 0 : b->plan.frames[(size_t)lo].api_index;
            const int a1 = k + 1 == n_chunks ? batch->n_frames : b->plan.frames[(size_t)hi].api_index;
            auto first_off = [&](int from, int64_t& c_off, int64_t& e_off) {
                c_off = batch->n_classes; e_off = batch->n_entries;
                for (int i = from; i < batch->n_frames; i++)
                    if (batch->frames[i].status == NVB_FRAME_OK && batch->frames[i].res_decoded) { c_off = batch->frames[i].classes_off; e_off = batch->frames[i].entries_off; break; }
            };
            int64_t c0, e0, c1, e1;
            first_off(a0, c0, e0);
            if (k + 1 == n_chunks) { c1 = batch->n_classes; e1 = batch->n_entries; } else first_off(a1, c1, e1);
            if (k == 0) { c0 = 0; e0 = 0; }
            const size_t row = (size_t)ctx->H.channels * ctx->H.post_stride;
            if (a1 > a0) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_posts + (size_t)a0 * row, batch->posts + (size_t)a0 * row, (size_t)(a1 - a0) * row * sizeof(int16_t), cudaMemcpyHostToDevice, st_up));
            if (c1 > c0) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_classes + c0, batch->classes + c0, (size_t)(c1 - c0), cudaMemcpyHostToDevice, st_up));
            if (e1 > e0) NVB_CUDA(ctx, cudaMemcpyAsync(b->d_entries + e0, batch->entries + e0, (size_t)(e1 - e0) * sizeof(uint16_t), cudaMemcpyHostToDevice, st_up));
        }
        NVB_CUDA(ctx, cudaEventRecord(sl.ev_up[k], st_up));
        NVB_CUDA(ctx, cudaStreamWaitEvent(st_k, sl.ev_up[k], 0));
        NVB_TRACE_MARK(st_k, "kernels_begin", k);
        // the whole batch is unpacked by ONE launch ahead of the first chunk: a packet is a serial, latency-bound walk, so the
        // more warps (packets) are resident the better it hides its table loads (4 x 1024 packets took 4 x 0.2 ms)
        if (pbatch && k == 0 && nf > 0 && (rc = enqueue_unpack(ctx, b, st_k, 0, nf)) != NVB_OK) { cudaDeviceSynchronize(); return rc; }
        rc = enqueue(ctx, b, 0, nullptr, d_float, true, st_k, n_chunks == 1 ? 0 : lo, n_chunks == 1 ? -1 : hi - lo, false);
        if (rc != NVB_OK) { cudaDeviceSynchronize(); return rc; }
        size_t s0 = 0, s1 = 0;                                              // this chunk's elements of the interleaved PCM
        if (nf > 0) {
            s0 = (size_t)b->plan.frames[(size_t)lo].pcm_off * C;
            s1 = (n_chunks > 1 && hi < nf) ? (size_t)b->plan.frames[(size_t)hi].pcm_off * C : n_out;
        }
        if (s16 && s1 > s0) {
            if (launch_pcm_s16(d_float, d_s16, (long long)s0, (long long)s1, st_k) < 0) { cudaDeviceSynchronize(); return cuda_fail(ctx, cudaGetLastError(), "k_pcm_s16 launch"); }
            b->launches += 1;
        }
        NVB_CUDA(ctx, cudaEventRecord(sl.ev_k[k], st_k));
        NVB_CUDA(ctx, cudaStreamWaitEvent(st_down, sl.ev_k[k], 0));
        NVB_TRACE_MARK(st_down, "d2h_begin", k);
        if (s1 > s0 && !dev_out) {
            if (s16) NVB_CUDA(ctx, cudaMemcpyAsync(reinterpret_cast<int16_t*>(pcm_out) + s0, d_s16 + s0, (s1 - s0) * sizeof(int16_t), cudaMemcpyDeviceToHost, st_down));
            else NVB_CUDA(ctx, cudaMemcpyAsync(pcm_out + s0, d_float + s0, (s1 - s0) * sizeof(float), cudaMemcpyDeviceToHost, st_down));
        }
        NVB_TRACE_MARK(st_down, "d2h_end", k);
    }
    std::memset(sl.h_counters, 0, sizeof(Counters));
    if (!b->plan.frames.empty()) NVB_CUDA(ctx, cudaMemcpyAsync(sl.h_counters, b->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, st_down));
    NVB_CUDA(ctx, cudaEventRecord(sl.ev_all, st_down));
    // the decoder state advances like StreamDecoder's (also when an entry turns out to be out of range)
    ctx->carry = b->plan.end_state;
    if (b->plan.last_ok >= 0) ctx->carry_cur ^= 1;
    ctx->in_flight++;
    return NVB_OK;
}

int nvb_decode_batch_end(nvb_ctx* ctx, nvb_result* res) {
    if (!ctx) return set_err(nullptr, NVB_ERR_ARG, "ctx is NULL");
    if (ctx->in_flight <= 0) return set_err(ctx, NVB_ERR_STATE, "no batch in flight");
    DeviceGuard g(ctx->device);
    nvb_ctx::Slot& sl = ctx->slot[ctx->head];
    ctx->head = (ctx->head + 1) % NVB_MAX_IN_FLIGHT; ctx->in_flight--;
    NVB_CUDA(ctx, cudaEventSynchronize(sl.ev_all));
#if !defined(NVB_CPU_SHIM)
    g_trace.dump(g_trace.batch - ctx->in_flight);
#endif
    nvb_dbatch* b = sl.staging;
    const Counters c = *sl.h_counters;
    if (res) {
        res->samples_per_channel = b->plan.samples;
        res->has_clipped = c.clipped ? 1 : 0;
        res->n_failed = b->plan.n_failed;
        res->n_floor_range = c.floor_range;
        res->n_inconsistent = b->plan.n_inconsistent;
    }
    if (c.bad_entry) return set_err(ctx, NVB_ERR_DATA, "a VQ entry number is outside its codebook (Codebook.cs:322 would throw)");
    return NVB_OK;
}

int nvb_decode_batch(nvb_ctx* ctx, const nvb_batch* batch, int flags, float* pcm_out, size_t pcm_cap, nvb_result* res) {
    if (ctx && ctx->in_flight > 0) return set_err(ctx, NVB_ERR_STATE, "batches begun with nvb_decode_batch_begin are still in flight");
    const int rc = nvb_decode_batch_begin(ctx, batch, flags, pcm_out, pcm_cap);
    if (rc != NVB_OK) return rc;
    return nvb_decode_batch_end(ctx, res);
}

int nvb_upload_unpack_tables(nvb_ctx* ctx, const void* blob, size_t bytes) {
    if (!ctx || !blob) return set_err(ctx, NVB_ERR_ARG, "NULL argument");
    if (!ctx->has_setup) return set_err(ctx, NVB_ERR_STATE, "no setup uploaded");
    if (ctx->in_flight > 0) return set_err(ctx, NVB_ERR_STATE, "batches in flight");
    using namespace nvbu;
    if (bytes < sizeof(UHeader)) return set_err(ctx, NVB_ERR_DATA, "unpack tables: blob too small");
    UHeader h; std::memcpy(&h, blob, sizeof h);
    if (h.magic != UNPACK_MAGIC || h.version != 2 || h.total_bytes != bytes) return set_err(ctx, NVB_ERR_DATA, "unpack tables: magic / size mismatch");
    if (h.channels != ctx->H.channels || h.bs[0] != ctx->H.bs[0] || h.bs[1] != ctx->H.bs[1] || h.post_stride != ctx->H.post_stride ||
        h.n_books != ctx->H.n_books || h.n_floors != ctx->H.n_floors || h.n_residues != ctx->H.n_residues || h.n_mappings != ctx->H.n_mappings || h.n_modes != ctx->H.n_modes)
        return set_err(ctx, NVB_ERR_DATA, "unpack tables do not belong to the uploaded setup");
    auto in = [&](uint64_t off, uint64_t len) { return off >= sizeof(UHeader) && len <= bytes && off <= bytes - len && (off & 15) == 0; };
    if (h.mode_bits < 0 || h.mode_bits > 8 || h.cls_stride < 1 || h.ent_stride < 1 || h.ent_stride > (1 << 22) + 8 ||
        !in(h.off_books, sizeof(UBook) * (uint64_t)h.n_books) || !in(h.off_roots, 4ull * h.n_roots) || !in(h.off_longs, sizeof(ULong) * (uint64_t)h.n_longs) ||
        !in(h.off_floors, sizeof(UFloor1) * (uint64_t)h.n_floors) || !in(h.off_residues, sizeof(UResidue) * (uint64_t)h.n_residues) || !in(h.off_digits, h.n_digits) ||
        !in(h.off_mappings, sizeof(UMapping) * (uint64_t)h.n_mappings) || !in(h.off_modes, sizeof(UMode) * (uint64_t)h.n_modes))
        return set_err(ctx, NVB_ERR_DATA, "unpack tables: section outside the blob");
    // index ranges the kernel relies on
    const unsigned char* base = static_cast<const unsigned char*>(blob);
    const UBook* books = reinterpret_cast<const UBook*>(base + h.off_books);
    const uint32_t* roots = reinterpret_cast<const uint32_t*>(base + h.off_roots);
    const ULong* longs = reinterpret_cast<const ULong*>(base + h.off_longs);
    for (int i = 0; i < h.n_books; i++) {
        const UBook& b = books[i];
        if (!b.decodable) continue;
        if (b.root_bits < 1 || b.root_bits > ROOT_BITS || b.n_long < 0 || (uint64_t)b.root_off + (1ull << b.root_bits) > h.n_roots || (uint64_t)b.long_off + (uint64_t)b.n_long > h.n_longs || b.entries < 1 || b.entries > (1 << 24))
            return set_err(ctx, NVB_ERR_DATA, "unpack tables: codebook table ranges");
        for (uint32_t k = 0; k < (1u << b.root_bits); k++) {
            const uint32_t r = roots[b.root_off + k], len = r >> 24;
            if (len ? (len > (uint32_t)b.root_bits || (r & 0xffffffu) >= (uint32_t)b.entries) : (r > (uint32_t)b.n_long)) return set_err(ctx, NVB_ERR_DATA, "unpack tables: root table entry");
        }
        for (int k = 0; k < b.n_long; k++) {
            const ULong& l = longs[b.long_off + k];
            // chains run strictly backwards (the builder links each element to an older one): no cycles
            if (l.len < 1 || l.len > 32 || l.value < 0 || l.value >= b.entries || l.next < 0 || l.next > k) return set_err(ctx, NVB_ERR_DATA, "unpack tables: long codeword");
        }
    }
    const UFloor1* floors = reinterpret_cast<const UFloor1*>(base + h.off_floors);
    for (int i = 0; i < h.n_floors; i++) {
        const UFloor1& f = floors[i];
        if (f.type == 0) {                                                  // type 0: the fields Floor0.Unpack reads; the records must fit the setup's stride
            const UFloor0& z = f.f0;
            if (z.order < 1 || z.order + 1 > h.f0_stride || h.f0_stride != ctx->H.f0_stride || z.amp_bits < 0 || z.amp_bits > 32 || z.amp_div < 1 || z.book_bits < 0 ||
                z.book_bits > 5 || z.n_books < 1 || z.n_books > 16) return set_err(ctx, NVB_ERR_DATA, "unpack tables: type 0 floor");
            for (int k = 0; k < z.n_books; k++)
                if (z.books[k] < 0 || z.books[k] >= h.n_books || ctx->S.n_vq <= 0) return set_err(ctx, NVB_ERR_DATA, "unpack tables: type 0 floor book");
            continue;
        }
        if (f.type != 1 || f.n_parts < 0 || f.n_parts > 32 || f.ybits < 1 || f.ybits > 16 || f.n_posts < 2 || f.n_posts > NVB_MAX_POSTS) return set_err(ctx, NVB_ERR_DATA, "unpack tables: floor");
        int posts = 2;
        for (int p = 0; p < f.n_parts; p++) {
            const int c = f.part_class[p];
            if (c >= 16 || f.class_dims[c] < 1 || f.class_dims[c] > 8 || f.class_subs[c] > 3) return set_err(ctx, NVB_ERR_DATA, "unpack tables: floor class");
            if (f.class_subs[c] > 0 && (f.class_master[c] < 0 || f.class_master[c] >= h.n_books)) return set_err(ctx, NVB_ERR_DATA, "unpack tables: floor master book");
            for (int k = 0; k < 8; k++) if (f.sub_books[c][k] >= h.n_books) return set_err(ctx, NVB_ERR_DATA, "unpack tables: floor book");
            posts += f.class_dims[c];
        }
        if (posts != f.n_posts || posts + 1 > h.post_stride) return set_err(ctx, NVB_ERR_DATA, "unpack tables: floor post count");
    }
    const UResidue* residues = reinterpret_cast<const UResidue*>(base + h.off_residues);
    for (int i = 0; i < h.n_residues; i++) {
        const UResidue& r = residues[i];
        if (r.type < 0 || r.type > 2 || r.begin < 0 || r.psize < 1 || r.nclass < 1 || r.nclass > 64 || r.stages < 0 || r.stages > 8 || r.class_book < 0 || r.class_book >= h.n_books ||
            r.cdims < 1 || r.partvals < 1 || (uint64_t)r.digits_off + (uint64_t)r.partvals * (uint64_t)r.cdims > h.n_digits)
            return set_err(ctx, NVB_ERR_DATA, "unpack tables: residue");
        const uint8_t* dg = base + h.off_digits + r.digits_off;
        for (int64_t k = 0; k < (int64_t)r.partvals * r.cdims; k++) if (dg[k] >= r.nclass) return set_err(ctx, NVB_ERR_DATA, "unpack tables: residue class digits");
        for (int c = 0; c < 64; c++) for (int st = 0; st < 8; st++) {
            const int bk = r.books[c][st];
            if (bk >= h.n_books || (bk >= 0 && books[bk].dims < 1)) return set_err(ctx, NVB_ERR_DATA, "unpack tables: residue book");
        }
    }
    const UMapping* mappings = reinterpret_cast<const UMapping*>(base + h.off_mappings);
    for (int i = 0; i < h.n_mappings; i++) {
        const UMapping& m = mappings[i];
        if (m.n_coupling < 0 || m.n_coupling > UNPACK_MAX_COUPLING || m.floor < 0 || m.floor >= h.n_floors || m.residue < 0 || m.residue >= h.n_residues) return set_err(ctx, NVB_ERR_DATA, "unpack tables: mapping");
        for (int k = 0; k < m.n_coupling; k++) if (m.mag[k] >= h.channels || m.ang[k] >= h.channels) return set_err(ctx, NVB_ERR_DATA, "unpack tables: coupling");
    }
    const UMode* modes = reinterpret_cast<const UMode*>(base + h.off_modes);
    for (int i = 0; i < h.n_modes; i++) if (modes[i].mapping < 0 || modes[i].mapping >= h.n_mappings) return set_err(ctx, NVB_ERR_DATA, "unpack tables: mode");
    // the strides must cover the largest frame of any mode (k_unpack writes without further checks)
    for (int i = 0; i < h.n_modes; i++) {
        const UMapping& m = mappings[modes[i].mapping]; const UResidue& r = residues[m.residue];
        const int N = h.bs[modes[i].block_flag ? 1 : 0];
        const int span = (r.type == 2 ? N * h.channels : N) / 2;
        const int nn = (r.end < span ? r.end : span) - r.begin;
        const int64_t P = nn > 0 ? nn / r.psize : 0, S = r.type == 2 ? 1 : h.channels;
        int64_t worst = 0;
        for (int c = 0; c < r.nclass; c++) {
            int64_t e = 0;
            for (int st = 0; st < r.stages; st++) { const int bk = r.books[c][st]; if (((r.cascade[c] >> st) & 1) && bk >= 0) e += r.type == 0 ? r.psize / books[bk].dims : (r.psize + books[bk].dims - 1) / books[bk].dims; }
            if (e > worst) worst = e;
        }
        if (S * P > h.cls_stride || worst * P * S > h.ent_stride) return set_err(ctx, NVB_ERR_DATA, "unpack tables: strides too small for the setup");
    }
    DeviceGuard g(ctx->device);
    unsigned char* d = nullptr;
    cudaError_t e = cudaMalloc((void**)&d, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(ctx, NVB_ERR_NOMEM, std::string("cudaMalloc(unpack tables): ") + cudaGetErrorString(e)); }
    e = cudaMemcpy(d, blob, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); return cuda_fail(ctx, e, "cudaMemcpy(unpack tables)"); }
    cudaFree(ctx->d_utab);
    ctx->d_utab = d; ctx->UH = h; ctx->has_unpack = true;
    UnpackTables& T = ctx->UT;
    T.books = reinterpret_cast<const UBook*>(d + h.off_books); T.roots = reinterpret_cast<const uint32_t*>(d + h.off_roots);
    T.longs = reinterpret_cast<const ULong*>(d + h.off_longs); T.floors = reinterpret_cast<const UFloor1*>(d + h.off_floors);
    T.residues = reinterpret_cast<const UResidue*>(d + h.off_residues); T.digits = d + h.off_digits;
    T.mappings = reinterpret_cast<const UMapping*>(d + h.off_mappings); T.modes = reinterpret_cast<const UMode*>(d + h.off_modes);
    T.channels = h.channels; T.mode_bits = h.mode_bits; T.post_stride = h.post_stride; T.cls_stride = h.cls_stride; T.ent_stride = h.ent_stride;
    return NVB_OK;
}

int nvb_unpack_strides(nvb_ctx* ctx, int32_t* cls_stride, int32_t* ent_stride) {
    if (!ctx || !ctx->has_unpack) return set_err(ctx, NVB_ERR_STATE, "no unpack tables uploaded");
    if (cls_stride) *cls_stride = ctx->UH.cls_stride;
    if (ent_stride) *ent_stride = ctx->UH.ent_stride;
    return NVB_OK;
}

int nvb_unpack_packets(nvb_ctx* ctx, const nvb_packet_batch* pb, nvb_frame* frames_out, int16_t* posts_out, uint8_t* classes_out, uint16_t* entries_out) {
    if (!ctx || !pb || !frames_out || !posts_out || !classes_out || !entries_out) return set_err(ctx, NVB_ERR_ARG, "NULL argument");
    if (!ctx->has_setup) return set_err(ctx, NVB_ERR_STATE, "no setup uploaded");
    if (ctx->in_flight > 0) return set_err(ctx, NVB_ERR_STATE, "batches in flight");
    DeviceGuard g(ctx->device);
    nvb_dbatch* b = new (std::nothrow) nvb_dbatch();
    if (!b) return set_err(ctx, NVB_ERR_NOMEM, "host allocation failed");
    int rc = upload_packets(ctx, b, pb, NVB_RUN_DEFAULT, ctx->stream, nullptr, nullptr);
    const size_t n = (size_t)(pb->n_packets > 0 ? pb->n_packets : 0);
    const size_t nf = b->plan.frames.size();
    if (rc == NVB_OK && nf) {
        // records of packets that are not decoded (failed status) stay zero
        cudaMemsetAsync(b->d_posts, 0, n * ctx->H.channels * ctx->H.post_stride * sizeof(int16_t), ctx->stream);
        cudaMemsetAsync(b->d_classes, 0, n * (size_t)ctx->UH.cls_stride, ctx->stream);
        cudaMemsetAsync(b->d_entries, 0, n * (size_t)ctx->UH.ent_stride * sizeof(uint16_t), ctx->stream);
        rc = enqueue_unpack(ctx, b, ctx->stream, 0, (int)nf);
    }
    std::vector<DevFrame> got(nf);
    if (rc == NVB_OK && nf) {
        cudaError_t e = cudaMemcpyAsync(got.data(), b->d_frames, nf * sizeof(DevFrame), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && n) e = cudaMemcpyAsync(posts_out, b->d_posts, n * ctx->H.channels * ctx->H.post_stride * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && n) e = cudaMemcpyAsync(classes_out, b->d_classes, n * (size_t)ctx->UH.cls_stride, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && n) e = cudaMemcpyAsync(entries_out, b->d_entries, n * (size_t)ctx->UH.ent_stride * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "nvb_unpack_packets read-back");
    } else if (rc == NVB_OK) {
        std::memset(posts_out, 0, n * ctx->H.channels * ctx->H.post_stride * sizeof(int16_t));
        std::memset(classes_out, 0, n * (size_t)ctx->UH.cls_stride); std::memset(entries_out, 0, n * (size_t)ctx->UH.ent_stride * sizeof(uint16_t));
    }
    if (rc == NVB_OK) {
        for (size_t i = 0; i < n; i++) { frames_out[i] = b->pkt_frames[i]; frames_out[i].res_decoded = 0; frames_out[i].exec_mask = 0; frames_out[i].entry_count = 0; }
        for (const DevFrame& d : got) if (d.kind == 0) { nvb_frame& f = frames_out[d.api_index]; f.exec_mask = d.exec_mask; f.res_decoded = d.res_decoded; f.entry_count = d.entry_count; }
    }
    cudaStreamSynchronize(ctx->stream);
    free_dbatch(b);
    return rc;
}

int nvb_dbatch_create(nvb_ctx* ctx, const nvb_batch* batch, int flags, nvb_dbatch** out) {
    if (!ctx || !out) return set_err(ctx, NVB_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (!ctx->has_setup) return set_err(ctx, NVB_ERR_STATE, "no setup uploaded");
    DeviceGuard g(ctx->device);
    nvb_dbatch* b = new (std::nothrow) nvb_dbatch();
    if (!b) return set_err(ctx, NVB_ERR_NOMEM, "host allocation failed");
    int rc = upload_batch(ctx, b, batch, flags, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc == NVB_OK && e != cudaSuccess) rc = cuda_fail(ctx, e, "cudaStreamSynchronize");
    if (rc != NVB_OK) { free_dbatch(b); return rc; }
    *out = b;
    return NVB_OK;
}

int64_t nvb_dbatch_samples(const nvb_dbatch* b) { return b ? b->plan.samples : -1; }
int64_t nvb_dbatch_spectrum_floats(const nvb_dbatch* b) { return b ? b->plan.spec_floats : -1; }
int nvb_dbatch_launches(const nvb_dbatch* b) { return b ? b->launches : -1; }

int nvb_dbatch_run(nvb_ctx* ctx, nvb_dbatch* b, float* d_pcm, void* stream) {
    if (!ctx || !b || (!d_pcm && b->plan.samples > 0)) return set_err(ctx, NVB_ERR_ARG, "NULL argument");
    DeviceGuard g(ctx->device);
    return enqueue(ctx, b, 0, nullptr, d_pcm, false, (cudaStream_t)stream);
}
int nvb_dbatch_run_spectrum(nvb_ctx* ctx, nvb_dbatch* b, float* d_spectrum, void* stream) {
    if (!ctx || !b || (!d_spectrum && b->plan.spec_floats > 0)) return set_err(ctx, NVB_ERR_ARG, "NULL argument");
    if (reinterpret_cast<uintptr_t>(d_spectrum) & 15) return set_err(ctx, NVB_ERR_ARG, "d_spectrum must be 16-byte aligned (bulk copies)");
    DeviceGuard g(ctx->device);
    return enqueue(ctx, b, 1, d_spectrum, nullptr, false, (cudaStream_t)stream);
}
int nvb_dbatch_run_imdct(nvb_ctx* ctx, nvb_dbatch* b, const float* d_spectrum, float* d_pcm, void* stream) {
    if (!ctx || !b || (!d_spectrum && b->plan.spec_floats > 0) || (!d_pcm && b->plan.samples > 0)) return set_err(ctx, NVB_ERR_ARG, "NULL argument");
    if (reinterpret_cast<uintptr_t>(d_spectrum) & 15) return set_err(ctx, NVB_ERR_ARG, "d_spectrum must be 16-byte aligned (bulk copies)");
    DeviceGuard g(ctx->device);
    return enqueue(ctx, b, 2, const_cast<float*>(d_spectrum), d_pcm, false, (cudaStream_t)stream);
}
int nvb_dbatch_result(nvb_ctx* ctx, nvb_dbatch* b, void* stream, nvb_result* res) {
    if (!ctx || !b) return set_err(ctx, NVB_ERR_ARG, "NULL argument");
    DeviceGuard g(ctx->device);
    return fetch_result(ctx, b, (cudaStream_t)stream, res);
}
int nvb_dbatch_destroy(nvb_ctx* ctx, nvb_dbatch* b) {
    if (!b) return NVB_OK;
    if (ctx) { DeviceGuard g(ctx->device); cudaDeviceSynchronize(); free_dbatch(b); }
    else free_dbatch(b);
    return NVB_OK;
}

}  // extern "C"
