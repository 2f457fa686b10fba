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
    } slot[2];
    int head = 0, in_flight = 0;
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
    cudaFree(b->d_spectrum); cudaFree(b->d_blocks); cudaFree(b->d_counters);
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
    return a;
}

// Enqueues the synthesis of an uploaded batch.  stage: 0 = all, 1 = spectrum only, 2 = IMDCT.. only.
// `spectrum` = dense spectrum buffer written by stage 1 and read by stage 2 (nullptr: the batch's own).
int enqueue(nvb_ctx* ctx, nvb_dbatch* b, int stage, float* spectrum, float* d_pcm, bool save_carry, cudaStream_t st,
            int frame_lo = 0, int frame_cnt = -1, bool reset_counters = true, cudaEvent_t after_spectrum = nullptr, cudaEvent_t before_synth = nullptr) {
    if (b->plan.frames.empty()) { b->launches = 0; return NVB_OK; }
    LaunchArgs a = make_args(ctx, b, spectrum ? spectrum : b->d_spectrum, d_pcm, save_carry);
    if (frame_cnt >= 0) { a.frame_lo = frame_lo; a.n_frames = frame_cnt; }
    int launches = 0, r;
    // the counters are cleared when they are read (fetch_result) or per batch (nvb_decode_batch_begin), not per run: a
    // memset between the kernels of consecutive runs would serialise what programmatic dependent launch overlaps
    (void)reset_counters;
    if (stage != 2) {
        if ((r = launch_spectrum(a, st)) < 0) return cuda_fail(ctx, cudaGetLastError(), "k_spectrum launch");
        launches += r;
        if (after_spectrum) NVB_CUDA(ctx, cudaEventRecord(after_spectrum, st));
    }
    if (stage != 1) {
        if (before_synth) NVB_CUDA(ctx, cudaStreamWaitEvent(st, before_synth, 0));     // the halo block's spectrum comes from the previous chunk
        if (b->fused) {
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
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
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
    cudaFree(ctx->d_blob); cudaFree(ctx->d_carry[0]); cudaFree(ctx->d_carry[1]);
    cudaStreamDestroy(ctx->stream);
    for (int i = 0; i < 2; i++) {
        cudaStreamDestroy(ctx->chunk_stream[i]);
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

int nvb_decode_batch_begin(nvb_ctx* ctx, const nvb_batch* batch, int flags, float* pcm_out, size_t pcm_cap) {
    if (!ctx) return set_err(nullptr, NVB_ERR_ARG, "ctx is NULL");
    if (!ctx->has_setup) return set_err(ctx, NVB_ERR_STATE, "no setup uploaded");
    if (ctx->in_flight >= 2) return set_err(ctx, NVB_ERR_STATE, "two batches are already in flight: call nvb_decode_batch_end first");
    DeviceGuard g(ctx->device);
    nvb_ctx::Slot& sl = ctx->slot[(ctx->head + ctx->in_flight) & 1];
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
    int rc = upload_batch(ctx, b, batch, flags, st_up, &chunked_inputs, &sl.h_frames, &sl.h_frames_cap);
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
    const int n_chunks = (chunked_inputs && nf >= chunk_min && nf >= 8) ? NVB_CHUNKS : 1;
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
    ctx->head ^= 1; ctx->in_flight--;
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
