// nvb_fused.cu -- fused IMDCT + window + overlap-add + clip + interleave for sm_100a (K4+K5).
//
// Replaces, per frame and channel: Mdct.Reverse (Mdct.cs:13-21,65-313), the window multiply of
// Mode.Decode (Mode.cs:159-166), StreamDecoder.OverlapBuffers (StreamDecoder.cs:532-541) and
// ClippingCopyBuffer / CopyBuffer (StreamDecoder.cs:391-415).
//
// Mapping onto the chip
//   * persistent grid: one CTA (16 warps) per SM owns a contiguous run of frames; its warps take the frames of the
//     run round-robin.  A warp transforms every channel of its frame (512 complex points as 16 per lane = two
//     radix-8 columns: three passes in registers, two exchanges through the frame's shared-memory slot), leaves
//     the DCT-IV output u in the slot, then writes the frame's PCM from its own u and the previous frame's u;
//   * the kernel is bound by shared-memory wavefronts and instruction issue, not by HBM or FP32 rate, so both are
//     minimised: the pre/post twiddles tw[64 j + r] = tw[r] E[j] are folded into compile-time constants E[j] plus
//     per-lane tables (13 conflict-free LDS.128 of twiddles per transform), the lane tables are staged by one bulk
//     copy (cp.async.bulk, TMA) per CTA, the TDAC symmetry out[i] / out[1023-i] lets every loaded u / window value
//     serve two output samples, exchange strides are padded (72 / 9 float2) and u is XOR-swizzled: no bank conflicts;
//   * no CTA-wide barrier in steady state: slots form a ring guarded by release/acquire counters in shared
//     memory -- full[slot] (u of that frame is complete; awaited by the warp that overlaps onto it) and
//     empty[slot] (both readers of the slot are done; awaited by the warp that reuses it) -- so transform,
//     overlap and store of different frames run concurrently on the four schedulers of the SM.  Counters
//     rather than mbarrier phases: nothing bounds the skew between two warps to one ring revolution, and a
//     parity wait cannot tell "two phases behind" from "done";
//   * the previous block's tail never touches HBM: each spectrum float is read once and each PCM float is
//     written once (16 384 B per stereo long frame); the first block of a run is recomputed as a halo;
//   * output: two samples x two channels per lane as one float4 store, a warp writes 512 contiguous bytes.
//   Round 2: both transforms of a stereo unit interleaved phase by phase (two independent instruction streams between the warp barriers,
//   no register prefetch: 2 x 32 data registers) -- 20.8 vs 17.8 us per 4096 frames and 7.7 vs 7.0 us for the single-frame chain (512
//   frames): at 128 registers ptxas serialises and spills instead of overlapping the streams.
//   Variants measured and dropped (profiles/r01_d, r01_f): TMA staging of the spectrum rows into the slot (the extra
//   LDS of the inputs costs more shared-memory bandwidth than the exposed load latency it saves); 32 warps x 64
//   registers with one warp per channel (more twiddle loads and counter polling, slower).
//   * even channel counts above two run as channel-pair units (k_imdct_fused_t<true>); every other pair of block sizes from
//     256 up runs on k_imdct_generic (radix-4 Stockham warp FFT in the slot) -- both further down in this file.
// No tensor cores: the IMDCT is FFT-structured, not a dense contraction.
#if !defined(NVB_CPU_SHIM)
#include <cuda_runtime.h>
#endif
#include <cstdlib>
#include "nvb_fused_core.h"
#include "nvb_wf_core.h"

namespace nvb {

#ifndef NVB_FUSED_WARPS
#define NVB_FUSED_WARPS 16
#endif
constexpr int FUSED_WARPS = NVB_FUSED_WARPS;
constexpr int FUSED_CTAS_PER_SM = FUSED_WARPS <= 8 ? 2 : 1;      // 8-warp build: two half-size CTAs share an SM
constexpr int FUSED_THREADS = FUSED_WARPS * 32;
#ifndef NVB_FUSED_LB_THREADS
#define NVB_FUSED_LB_THREADS FUSED_THREADS                                  // experiment hook: a larger bound caps the registers of builds with fewer warps
#endif
constexpr size_t FUSED_SMEM_LIMIT = (FUSED_WARPS <= 8 ? 112 : 227) * 1024 - FUSED_WARPS * 64;   // minus the static per-warp plan records

struct FusedParams {
    LaunchArgs a;
    int frames_per_cta;
    int n_slots;
    int skew;                   // stress hook (NVB_FUSED_SKEW): pseudo-random pauses that shuffle the warps' relative progress
    WfLayout wfl;               // one-kernel synthesis (SYN): per-warp scratch of the spectrum stage ...
    int wf_off;                 // ... which starts wf_off bytes into the dynamic shared memory (inverse_dB_table first, then the warps)
};

// Stress hook of the slot-ring protocol: a pause that depends on (warp, unit, site), so that warps overtake and fall behind
// each other in ways normal timing never produces.
__device__ __forceinline__ void fused_skew(int skew, int warp, int v, int site) {
#if !defined(NVB_CPU_SHIM)
    if (skew) __nanosleep((unsigned)((warp * 7919 + v * 104729 + site * 1299709) & 2047) * (unsigned)skew);
#else
    (void)skew; (void)warp; (void)v; (void)site;
#endif
}

#if !defined(NVB_CPU_SHIM)
// ---- mbarrier / bulk-copy primitives (PTX) ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
// Monotonic event counters in shared memory: signal = release-add by one lane (after __syncwarp), wait = acquire-poll.
__device__ __forceinline__ void cnt_signal(int* c) {
    asm volatile("red.release.cta.shared::cta.add.s32 [%0], 1;" ::"r"(smem_u32(c)) : "memory");
}
__device__ __forceinline__ void cnt_wait(const int* c, int need) {
    const uint32_t addr = smem_u32(c);
    int v;
    for (;;) {
        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        if (v >= need) break;
        __nanosleep(32);                                                    // (an exponential back-off measured no different, r02)
    }
}
__device__ __forceinline__ bool cnt_ready(const int* c, int need) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(c)) : "memory");
    return v >= need;
}
// One thread: global -> shared bulk copies (TMA), completion counted in bytes on `bar`.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 16-byte asynchronous global -> shared copy (LDGSTS): no register, no scoreboard; completion via cp_async_wait_all.
__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Orders earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) writes.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

__device__ __forceinline__ float clipf(float v, float& peak) {
    peak = fmaxf(peak, fabsf(v));
    return fminf(fmaxf(v, -0.99999994f), 0.99999994f);                       // Utils.ClipValue, Utils.cs:30-43
}

// Windowed block value z[i] = y[i] * window[i] of the block held in `slot` (Mode.cs:159-166).
__device__ __forceinline__ float slot_z(const DevSetup& S, const DevFrame& f, const float* slot, int c, int i, bool swz = true) {
    const bool exec = (f.exec_mask >> c) & 1u;
    return fused_y(slot, exec, f.n, i, swz) * frame_window(S, f)[i];
}

// Output of one frame for every shape the TDAC fast path of k_imdct_fused does not cover (short blocks, window transitions,
// silent channels, drains, C != 2) and for all of k_imdct_generic: lanes over samples; the index of y[i] inside a slot
// (fused_y_index) and the window values are per-sample, the channels of the unit an inner loop.  slot stride per channel =
// slot_floats.  SWZ (k_imdct_fused): the windows come from the two slopes in shared memory (s_tab; round 2: the global window
// loads put ~400 cycles of latency into every iteration, and a batch with window transitions waits for its slowest frame),
// a lane takes two samples per step, and a two-channel unit stores float2 per sample.
template <bool SWZ>
__device__ __forceinline__ void emit_samples(const LaunchArgs& a, const DevSetup& S, const DevFrame& f, const DevFrame* pf, const float* slots_f,
                                             const float* slots_p, int slot_floats, int lane, float& peak, int cbase, int G, const float* s_tab, int s_lo) {
    const int C = S.channels;                                               // row stride of the interleave; the unit covers channels cbase .. cbase + G - 1
    const int len = f.out_end - f.out_begin;
    const float* wf = (!SWZ && f.kind == 0) ? frame_window(S, f) : nullptr;
    const float* wp = (!SWZ && pf) ? frame_window(S, *pf) : nullptr;
    const float* s_win = SWZ ? s_tab + FusedTables::WIN : nullptr;
    const float* s_win0 = SWZ ? s_tab + FusedTables::WIN0 : nullptr;
    const int nf_ = f.n, np_ = pf ? pf->n : 0;
    const int wxf = f.window, wxp = pf ? pf->window : 0;
    const uint32_t ex_f = f.exec_mask >> cbase, ex_p = pf ? (pf->exec_mask >> cbase) : 0u;
    const bool clip = a.clip != 0;
    const bool cur = f.kind == 0;
    constexpr int SPL = SWZ ? 2 : 1;                                        // samples per lane and step
    // The flat part of a long block next to a short one (Mode.cs:24-67: window == 1.0 between the slopes, no overlap from the
    // previous block): out[i] = y[i] = -u[1535 - i] for 512 <= i < 1536, two samples x two channels per lane and step from two
    // 8-byte loads.  Window shapes 0 / 1 / 2 have 896 / 448 / 448 such samples; the per-sample path below costs ~4x more.
    int fl_lo = 0, fl_hi = 0;                                               // sample numbers s (relative to out_begin)
    if (SWZ && G == 2 && C == 2 && cur && nf_ == FUSED_LONG_N && (ex_f & 3u) == 3u && wxf != 3 && (f.out_begin & 1) == 0) {
        int i_lo = (wxf & 1) ? 1024 : 576, i_hi = (wxf & 2) ? 1024 : 1472;
        const int ola_end = f.start + (f.ola_len > 0 ? f.ola_len : 0);
        if (i_lo < ola_end) i_lo = ola_end;
        if (i_lo < f.out_begin) i_lo = f.out_begin;
        if (i_hi > f.out_end) i_hi = f.out_end;
        i_lo = (i_lo + 1) & ~1;                                             // pairs (i even, i + 1) share one aligned float2 of u
        i_hi &= ~1;
        if (i_hi - i_lo >= 64 && i_lo - f.out_begin >= s_lo) {
            fl_lo = i_lo - f.out_begin; fl_hi = i_hi - f.out_begin;
            const bool a16 = (((size_t)f.pcm_off + fl_lo) & 1) == 0;        // float4 stores need an even sample offset
            for (int i = i_lo + 2 * lane; i < i_hi; i += 64) {
                const int j = u_swz(1534 - i);                              // floats (1534 - i, 1535 - i) = (y[i + 1], y[i]) negated; the swizzle keeps the pair together
                const float2 u0 = *reinterpret_cast<const float2*>(slots_f + j);
                const float2 u1 = *reinterpret_cast<const float2*>(slots_f + slot_floats + j);
                float v0 = -u0.y, v1 = -u1.y, v2 = -u0.x, v3 = -u1.x;       // sample i: (ch0, ch1), sample i + 1: (ch0, ch1)
                if (clip) { v0 = clipf(v0, peak); v1 = clipf(v1, peak); v2 = clipf(v2, peak); v3 = clipf(v3, peak); }
                float* dst = a.pcm + ((size_t)f.pcm_off + (i - f.out_begin)) * 2;
                if (a16) *reinterpret_cast<float4*>(dst) = make_float4(v0, v1, v2, v3);
                else { *reinterpret_cast<float2*>(dst) = make_float2(v0, v1); *reinterpret_cast<float2*>(dst + 2) = make_float2(v2, v3); }
            }
        }
    }
    for (int s0 = s_lo + lane * SPL; s0 < len; s0 += 32 * SPL) {                // s_lo: the samples before it were written by the caller
        if (s0 >= fl_lo && s0 < fl_hi) continue;                            // written above (fl_lo, fl_hi and s0 are even)
        float wv[SPL], wpv[SPL], sf[SPL], sp[SPL]; int jf[SPL], jp[SPL], ii[SPL], ipp[SPL]; bool use_p[SPL], in[SPL];
        #pragma unroll
        for (int q = 0; q < SPL; q++) {
            const int s = s0 + q;
            in[q] = s < len;
            const int i = f.out_begin + (in[q] ? s : 0);
            const int o = i - f.start;
            const bool ola = cur && f.ola_len > 0 && o >= 0 && o < f.ola_len;              // StreamDecoder.cs:532-541
            const int ip = cur ? f.prev_valid + o : i;                                   // sample of the previous block (overlap or drain)
            use_p[q] = ola || !cur;
            ii[q] = i; ipp[q] = ip;
            wv[q] = 0.f; wpv[q] = 0.f; jf[q] = 0; jp[q] = 0; sf[q] = 0.f; sp[q] = 0.f;
            if (cur) { wv[q] = SWZ ? fused_window_at(s_win, s_win0, nf_, wxf, i) : wf[i]; fused_y_index(nf_, i, jf[q], sf[q], SWZ); }
            if (use_p[q] && pf) { wpv[q] = SWZ ? fused_window_at(s_win, s_win0, np_, wxp, ip) : wp[ip]; fused_y_index(np_, ip, jp[q], sp[q], SWZ); }
        }
        if (SWZ && G == 2) {                                                // a stereo unit: both channels of a sample in one float2
            #pragma unroll
            for (int q = 0; q < SPL; q++) {
                if (!in[q]) continue;
                float v[2];
                #pragma unroll
                for (int c = 0; c < 2; c++) {
                    float acc = 0.f;
                    if (cur) {
                        const float* sl = slots_f + (size_t)c * slot_floats;
                        const float y = ((ex_f >> c) & 1u) ? sf[q] * sl[jf[q]] : (ii[q] < (nf_ >> 1) ? sl[ii[q]] : 0.f);
                        acc = y * wv[q];
                    }
                    if (use_p[q]) {
                        if (pf) {
                            const float* sl = slots_p + (size_t)c * slot_floats;
                            const float y = ((ex_p >> c) & 1u) ? sp[q] * sl[jp[q]] : (ipp[q] < (np_ >> 1) ? sl[ipp[q]] : 0.f);
                            acc += y * wpv[q];
                        } else if (f.prev == PREV_CARRY) acc += a.carry_in[(size_t)(cbase + c) * S.bs[1] + ipp[q]];
                    }
                    if (clip) acc = clipf(acc, peak);
                    v[c] = acc;
                }
                *reinterpret_cast<float2*>(a.pcm + ((size_t)f.pcm_off + s0 + q) * C + cbase) = make_float2(v[0], v[1]);
            }
        } else {
            #pragma unroll
            for (int q = 0; q < SPL; q++) {
                if (!in[q]) continue;
                float* dst = a.pcm + ((size_t)f.pcm_off + s0 + q) * C + cbase;
                for (int c = 0; c < G; c++) {
                    float v = 0.f;
                    if (cur) {
                        const float* sl = slots_f + (size_t)c * slot_floats;
                        const float y = ((ex_f >> c) & 1u) ? sf[q] * sl[jf[q]] : (ii[q] < (nf_ >> 1) ? sl[ii[q]] : 0.f);
                        v = y * wv[q];
                    }
                    if (use_p[q]) {
                        if (pf) {
                            const float* sl = slots_p + (size_t)c * slot_floats;
                            const float y = ((ex_p >> c) & 1u) ? sp[q] * sl[jp[q]] : (ipp[q] < (np_ >> 1) ? sl[ipp[q]] : 0.f);
                            v += y * wpv[q];
                        } else if (f.prev == PREV_CARRY) v += a.carry_in[(size_t)(cbase + c) * S.bs[1] + ipp[q]];
                    }
                    if (clip) v = clipf(v, peak);
                    dst[c] = v;
                }
            }
        }
    }
}

// GROUPED (even channel counts above two, no drains in the launch): the unit of work is a CHANNEL PAIR of a frame instead
// of a whole frame -- slots hold two channels (18 of them instead of 7 six-channel slots), every warp stays busy, and the pair
// takes the stereo TDAC output path with two float2 stores per sample pair into the C-channel interleave.  Units of one
// frame are consecutive (v = frame * U + pair), the previous block of a unit is unit v - U.
// SYN (one-kernel synthesis, K1-K5; 0 = off, else the channel count): the warp first computes its frame's spectrum from the boundary
// records -- k_spectrum_wf's frame function (nvb_wf_core.h): residue gather, inverse coupling, floor curve -- straight into the
// frame's slot, and the transforms read their inputs from there: one launch per batch, no dense spectrum in HBM (9.7 KB instead
// of 25.7 KB of algorithmic traffic per stereo long frame).  The halo block's spectrum is recomputed like its transform.
template <bool GROUPED, int SYN = 0>
__global__ void __launch_bounds__(NVB_FUSED_LB_THREADS, FUSED_CTAS_PER_SM) k_imdct_fused_t(FusedParams p) {
    NVB_DYN_SMEM(smem_raw);
    const LaunchArgs& a = p.a;
    const DevSetup& S = a.S;
    const int C = S.channels;
    const int G = GROUPED ? 2 : C;                                          // channels per unit (= per slot)
    const int U = GROUPED ? (C >> 1) : 1;                                   // units per frame
    const int NS = p.n_slots;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    float* s_tab = reinterpret_cast<float*>(smem_raw);
    float* s_slots = s_tab + FusedTables::FLOATS;
    DevFrame* s_fr = reinterpret_cast<DevFrame*>(s_slots + (size_t)NS * G * FUSED_SLOT_FLOATS);
    __shared__ DevFrame s_pre[FUSED_WARPS];                                 // per warp: the plan record of its next unit, requested one unit ahead (cp.async)
    int* s_full = reinterpret_cast<int*>(s_fr + NS);                        // s_full[s]: frames completed in slot s
    int* s_empty = s_full + NS;                                             // s_empty[s]: reader releases of slot s (two per frame)
    uint64_t* s_tabbar = reinterpret_cast<uint64_t*>(s_empty + NS);          // 8 NS bytes past s_full: 8-byte aligned
    float* s_db = reinterpret_cast<float*>(smem_raw + p.wf_off);            // SYN: inverse_dB_table, then one spectrum-stage scratch area per warp
    const CiRec* s_ci = reinterpret_cast<const CiRec*>(smem_raw + p.wf_off + 1024);      // SYN: the setup's (class, stage) records
    unsigned char* s_wf = smem_raw + p.wf_off + 1024 + p.wfl.cta_bytes + (size_t)warp * p.wfl.total;

    const int lo = a.frame_lo + blockIdx.x * p.frames_per_cta;              // plan indices; this launch covers [frame_lo, frame_lo + n_frames)
    int hi = lo + p.frames_per_cta; if (hi > a.frame_lo + a.n_frames) hi = a.frame_lo + a.n_frames;
    nvb_grid_dep_launch();                                                  // the next kernel of the stream may start scheduling its blocks
    if (lo >= hi) return;

    if (tid == 0) {
        for (int s = 0; s < NS; s++) { s_full[s] = 0; s_empty[s] = 0; }
        mbar_init(s_tabbar, 1);
        mbar_fence_init();
    }
    if (SYN) {
        for (int i = tid; i < 256; i += FUSED_THREADS) s_db[i] = S.db[i];
        for (int i = tid; i < S.ci_total; i += FUSED_THREADS) reinterpret_cast<int4*>(smem_raw + p.wf_off + 1024)[i] = reinterpret_cast<const int4*>(S.ci)[i];
    }
    __syncthreads();
    if (tid == 0) {                                                         // the lane tables: one bulk copy (TMA) per CTA
        mbar_arrive_expect_tx(s_tabbar, FusedTables::FLOATS * sizeof(float));
        bulk_g2s(s_tab, S.fused_tab, FusedTables::FLOATS * sizeof(float), s_tabbar);
    }

    nvb_grid_dep_wait();                                                    // everything below reads what earlier work of the stream wrote
    int first = lo;
    {
        const DevFrame f0 = a.frames[lo];
        if (f0.prev >= 0 && (f0.ola_len > 0 || f0.kind != 0)) first = lo - 1;   // halo: previous block's tail is needed
    }
    // the first transform's rows are requested before the lane tables have landed: both latencies overlap
    const int vfirst = first * U, vhi = hi * U;                             // unit indices (== frame indices when not grouped)
    // The plan record of a warp's NEXT unit is requested one unit ahead with cp.async (no register, no scoreboard) into the
    // warp's staging record: the dependent global load of the record used to stall the start of every unit.
    if (vfirst + warp < vhi && lane < 4) {
        const int v0 = vfirst + warp, x0 = GROUPED ? v0 / U : v0;
        cp_async_16(reinterpret_cast<char*>(&s_pre[warp]) + 16 * lane, reinterpret_cast<const char*>(a.frames + x0) + 16 * lane);
    }
    LongIn pre; int pre_x = -1, pre_c = -1;
#if !defined(NVB_FUSED_NO_PREFETCH)
    if (!SYN && vfirst + warp < vhi) {
        const int v0 = vfirst + warp, x0 = GROUPED ? v0 / U : v0, c0 = GROUPED ? (v0 - x0 * U) * 2 : 0;
        const DevFrame* f0 = a.frames + x0;
        if (f0->kind == 0 && f0->n == FUSED_LONG_N && ((f0->exec_mask >> c0) & 1u)) {
            long_phase1_load(lane, reinterpret_cast<const float2*>(a.spectrum + (size_t)f0->spec_off + (size_t)c0 * (FUSED_LONG_N / 2)), pre); pre_x = v0; pre_c = 0;
        }
    }
#endif
    mbar_wait(s_tabbar, 0);

    const float2* s_tw0 = reinterpret_cast<const float2*>(s_tab + FusedTables::TW0);
    const float2* s_w64 = reinterpret_cast<const float2*>(s_tab + FusedTables::W64);
    const float* s_win = s_tab + FusedTables::WIN;
    LaneTw ltw; lane_tw_load(lane, s_tab, ltw);
    // swizzled float offsets of this lane's sample pair inside a run of 64 (forward / mirrored), see the fast path
    const int L2 = 2 * lane;
    const int A0 = L2 ^ ((lane >> 4) << 2), A8 = A0 ^ 8;
    const int Bm = 62 - L2;
    const int B0 = Bm ^ ((Bm >> 5) << 2), B8 = B0 ^ 8;
    float peak = 0.f;
    // Register-level software pipeline: the spectrum rows of the warp's next long transform are loaded right after
    // the current transform has consumed its inputs, so the load latency hides behind passes 2-3 and the output.
#if defined(NVB_FUSED_NO_PREFETCH)
    // build option: no register prefetch (32 fewer live registers; more warps per SM hide the load latency instead)
    auto can_prefetch = [&](int, int, uint32_t, int) { return false; };
#else
    auto can_prefetch = [&](int kind, int n, uint32_t exec_mask, int cc) { return !SYN && kind == 0 && n == FUSED_LONG_N && ((exec_mask >> cc) & 1u); };
#endif

    // Units go to the warps round-robin.  (Claiming them dynamically -- a warp that finishes a short block takes the next unit at
    // once -- measured slower in round 2: 18.3 vs 18.0 us on configs[1], 0.143 vs 0.121 ms on the mixed-window configs[2]: a warp
    // has to claim its next unit early to prefetch it, and then sits on that claim while it finishes a long block.)
    for (int v = vfirst + warp; v < vhi; v += FUSED_WARPS) {
        const int rel = v - vfirst, slot = rel % NS, it = rel / NS;
        const int x = GROUPED ? v / U : v;                                   // frame and first channel of this unit
        const int cbase = GROUPED ? (v - x * U) * 2 : 0;
        fused_skew(p.skew, warp, v, 0);
        cnt_wait(&s_empty[slot], 2 * it);                                    // both readers of every earlier frame of the slot are done
        cp_async_wait_all();                                                 // this unit's plan record (requested one unit ago) has landed
        __syncwarp();
        if (lane < 4) reinterpret_cast<int4*>(&s_fr[slot])[lane] = reinterpret_cast<const int4*>(&s_pre[warp])[lane];
        __syncwarp();
        const DevFrame f = s_fr[slot];
        const int vn = v + FUSED_WARPS;                                      // the warp's next unit: its record is requested now, read after phase 1
        const int xn = GROUPED ? vn / U : vn, cn = GROUPED ? (vn - xn * U) * 2 : 0;
        if (vn < vhi && lane < 4) cp_async_16(reinterpret_cast<char*>(&s_pre[warp]) + 16 * lane, reinterpret_cast<const char*>(a.frames + xn) + 16 * lane);
        int n_kind = 1, n_n = 0; uint32_t n_exec = 0, n_spec = 0; bool n_known = false;
        float* slots_f = s_slots + (size_t)slot * G * FUSED_SLOT_FLOATS;

        // ---------------- SYN: the frame's spectrum, from its boundary records into the slot ----------
        if constexpr (SYN != 0) if (f.kind == 0) {
            int bad_entry = 0, bad_floor = 0;
            wf_frame_to_slot<SYN, false>(a, f, p.wfl, s_wf, reinterpret_cast<int*>(slots_f), s_db, s_ci, wf_smem(slots_f), (uint32_t)(FUSED_SLOT_FLOATS * sizeof(float)), lane, bad_entry, bad_floor);
            // a frame is counted once per kind, by the CTA that emits it (not by the one that recomputes it as a halo)
            const bool be = __any_sync(0xffffffffu, bad_entry != 0), bf = __any_sync(0xffffffffu, bad_floor != 0);
            if (x >= lo && lane == 0) { if (be) atomicAdd(&a.counters->bad_entry, 1); if (bf) atomicAdd(&a.counters->floor_range, 1); }
            __syncwarp();
        }
        // ---------------- transform: every channel of frame x -------------------------------------
        if (f.kind == 0) {
            for (int c = 0; c < G; c++) {
                float* slotc = slots_f + (size_t)c * FUSED_SLOT_FLOATS;
                const int M = f.n >> 1;
                const float* spec = SYN ? slotc : a.spectrum + (size_t)f.spec_off + (size_t)(cbase + c) * M;
                if (!((f.exec_mask >> (cbase + c)) & 1u)) {
                    if (!SYN) { for (int i = lane; i < M; i += 32) slotc[i] = spec[i]; }      // raw residue values (Mapping.cs:192-196)
                    else if (f.n == FUSED_LONG_N) {                          // already there; a long block's 16-byte chunks are permuted: undo
                        float4 t[8];
                        #pragma unroll
                        for (int k = 0; k < 8; k++) { const int j = lane + 32 * k; t[k] = reinterpret_cast<const float4*>(slotc)[j ^ ((j >> 3) & 1)]; }
                        __syncwarp();
                        #pragma unroll
                        for (int k = 0; k < 8; k++) reinterpret_cast<float4*>(slotc)[lane + 32 * k] = t[k];
                    }
                } else if (f.n == FUSED_LONG_N) {
                    float2* ex = reinterpret_cast<float2*>(slotc);
                    LongRegs R;
                    if (SYN) { long_phase1_load_slot(lane, reinterpret_cast<const float2*>(slotc), pre); __syncwarp(); }   // inputs out of the slot before ex overwrites it
                    else if (!(pre_x == v && pre_c == c)) long_phase1_load(lane, reinterpret_cast<const float2*>(spec), pre);   // cold start
                    long_phase1_compute(lane, pre, s_tab, ex);
                    if (c + 1 < G && can_prefetch(0, f.n, f.exec_mask >> cbase, c + 1)) {
                        long_phase1_load(lane, reinterpret_cast<const float2*>(spec + M), pre); pre_x = v; pre_c = c + 1;
                    } else if (c + 1 == G && vn < vhi) {
                        if (!n_known) {                                      // the next unit's record: long since landed
                            cp_async_wait_all(); __syncwarp();
                            const DevFrame* fn = &s_pre[warp];
                            n_kind = fn->kind; n_n = fn->n; n_exec = fn->exec_mask >> cn; n_spec = fn->spec_off + (uint32_t)cn * (uint32_t)(fn->n >> 1); n_known = true;
                        }
                        if (can_prefetch(n_kind, n_n, n_exec, 0)) { long_phase1_load(lane, reinterpret_cast<const float2*>(a.spectrum + (size_t)n_spec), pre); pre_x = vn; pre_c = 0; }
                    }
                    __syncwarp();
                    long_phase2_load(lane, ex, R);
                    __syncwarp();
                    long_phase2_store(lane, ltw, ex, R);
                    __syncwarp();
                    long_phase3_load(lane, ex, R);
                    __syncwarp();
                    long_phase3_store(lane, ltw, ex, R);
                } else {
                    ShortRegs R;
                    short_phase1(lane, spec, s_tw0, s_w64, R);
                    #pragma unroll
                    for (int s = 16; s >= 1; s >>= 1) {
                        cpx pa, pb;
                        pa.x = __shfl_xor_sync(0xffffffffu, R.a.x, s); pa.y = __shfl_xor_sync(0xffffffffu, R.a.y, s);
                        pb.x = __shfl_xor_sync(0xffffffffu, R.b.x, s); pb.y = __shfl_xor_sync(0xffffffffu, R.b.y, s);
                        R.a = short_stage(lane, s, R.a, pa, s_w64);
                        R.b = short_stage(lane, s, R.b, pb, s_w64);
                    }
                    short_phase3_store(lane, s_tw0, slotc, R);
                }
            }
        }
        __syncwarp();
        if (lane == 0) cnt_signal(&s_full[slot]);                            // u of frame x is complete

        // Frame x-1 must have claimed and filled its slot before this warp releases it (a release that overtakes the
        // frame itself would be counted against the slot's next user), whether its tail is needed or not.
        fused_skew(p.skew, warp, v, 1);
        if (rel >= U) cnt_wait(&s_full[(rel - U) % NS], (rel - U) / NS + 1);
        fused_skew(p.skew, warp, v, 2);

        // ---------------- output of frame x (a halo block only leaves its tail) --------------------
        if (x >= lo) {
            const int len = f.out_end - f.out_begin;
            const DevFrame* pf = nullptr; const float* slots_p = nullptr;
            if (f.prev >= 0 && (f.ola_len > 0 || f.kind != 0)) {
                const int pslot = (GROUPED ? (f.prev * U + (v - x * U)) - vfirst : f.prev - first) % NS;
                pf = &s_fr[pslot]; slots_p = s_slots + (size_t)pslot * G * FUSED_SLOT_FLOATS;
            }
            // (window 1 = the next block is short: the first 1024 samples are the same TDAC pairs, 448 flat samples follow)
            const bool fast = (G == 2) && f.kind == 0 && pf && f.n == FUSED_LONG_N && pf->n == FUSED_LONG_N && (f.window & 1) &&
                              (pf->window & 2) && f.start == 0 && f.out_begin == 0 && f.out_end >= 1024 && f.ola_len == 1024 &&
                              f.prev_valid == 1024 && ((f.exec_mask >> cbase) & 3u) == 3u && ((pf->exec_mask >> cbase) & 3u) == 3u && pf->kind == 0 &&
                              (GROUPED || (f.pcm_off & 1) == 0);
            if (fast) {
                // long block after long block, both channels live (Mode.cs:44-50 window 3).  With a = u[512+i] of this
                // block, b = u'[511-i] of the previous one, s = S[i], s' = S[1023-i] (i < 512):
                //     out[i] = s a - s' b          out[1023-i] = -s' a - s b
                // (TDAC: the two samples mirror each other), so every loaded value serves two outputs.
                float4* out = reinterpret_cast<float4*>(a.pcm + (size_t)f.pcm_off * 2);
                const bool clip = a.clip != 0;
                #pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int i = L2 + 64 * k;
                    const float2 wl = *reinterpret_cast<const float2*>(s_win + i);          // (S[i], S[i+1])
                    const float2 wr = *reinterpret_cast<const float2*>(s_win + 1022 - i);   // (S[1022-i], S[1023-i])
                    float lo_[4], hi_[4];
                    #pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const float2 a2 = *reinterpret_cast<const float2*>(slots_f + c * FUSED_SLOT_FLOATS + 512 + 64 * k + ((k & 1) ? A8 : A0));   // u[512+i], u[513+i]
                        const float2 b2 = *reinterpret_cast<const float2*>(slots_p + c * FUSED_SLOT_FLOATS + 448 - 64 * k + ((k & 1) ? B0 : B8));   // u'[510-i], u'[511-i]
                        lo_[c]     = fmaf(wl.x, a2.x, -(wr.y * b2.y));           // out[i]
                        lo_[2 + c] = fmaf(wl.y, a2.y, -(wr.x * b2.x));           // out[i+1]
                        hi_[2 + c] = -fmaf(wr.y, a2.x, wl.x * b2.y);             // out[1023-i]
                        hi_[c]     = -fmaf(wr.x, a2.y, wl.y * b2.x);             // out[1022-i]
                    }
                    if (clip) {
                        // Utils.ClipValue (Utils.cs:30-43): clamping is only needed when some sample of the warp's 256 exceeds the
                        // limit, which is rare: one max reduction + vote instead of three min/max per sample
                        const float m = fmaxf(fmaxf(fmaxf(fabsf(lo_[0]), fabsf(lo_[1])), fmaxf(fabsf(lo_[2]), fabsf(lo_[3]))),
                                              fmaxf(fmaxf(fabsf(hi_[0]), fabsf(hi_[1])), fmaxf(fabsf(hi_[2]), fabsf(hi_[3]))));
                        if (__any_sync(0xffffffffu, m > 0.99999994f)) {
                            #pragma unroll
                            for (int e = 0; e < 4; e++) { lo_[e] = clipf(lo_[e], peak); hi_[e] = clipf(hi_[e], peak); }
                        }
                    }
                    if (!GROUPED) {
                        out[lane + 32 * k] = make_float4(lo_[0], lo_[1], lo_[2], lo_[3]);
                        out[511 - lane - 32 * k] = make_float4(hi_[0], hi_[1], hi_[2], hi_[3]);
                    } else {                                                 // the pair's two samples inside the C-channel interleave
                        float* o = a.pcm + (size_t)f.pcm_off * C + cbase;
                        *reinterpret_cast<float2*>(o + (size_t)i * C) = make_float2(lo_[0], lo_[1]);
                        *reinterpret_cast<float2*>(o + (size_t)(i + 1) * C) = make_float2(lo_[2], lo_[3]);
                        *reinterpret_cast<float2*>(o + (size_t)(1022 - i) * C) = make_float2(hi_[0], hi_[1]);
                        *reinterpret_cast<float2*>(o + (size_t)(1023 - i) * C) = make_float2(hi_[2], hi_[3]);
                    }
                }
                if (f.out_end > 1024) emit_samples<true>(a, S, f, pf, slots_f, slots_p, FUSED_SLOT_FLOATS, lane, peak, cbase, G, s_tab, 1024);
            } else if (len > 0) {
                emit_samples<true>(a, S, f, pf, slots_f, slots_p, FUSED_SLOT_FLOATS, lane, peak, cbase, G, s_tab, 0);
            }
            if (x == a.carry_frame && a.carry_out && f.kind == 0) {
                // keep the last windowed block for the next batch (StreamDecoder.cs:455-461)
                for (int idx = lane; idx < f.n * G; idx += 32) {
                    const int c = idx / f.n, i = idx - c * f.n;
                    a.carry_out[(size_t)(cbase + c) * S.bs[1] + i] = slot_z(S, f, slots_f + c * FUSED_SLOT_FLOATS, cbase + c, i);
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            cnt_signal(&s_empty[slot]);                                      // done with frame x as "current"
            if (rel >= U) cnt_signal(&s_empty[(rel - U) % NS]);              // done with the previous block's unit as "previous"
        }
    }
    if (__any_sync(0xffffffffu, peak > 0.99999994f) && lane == 0) atomicOr(&a.counters->clipped, 1);
}

// ------------------------------------------------------------------------------------------------
// k_imdct_generic -- the same fused IMDCT + window + overlap-add + clip + interleave for every other pair of block sizes
// from 256 up (e.g. 512/1024, 512/4096, 1024/8192; the reference's Mdct does not compute an IMDCT below 256, so those
// sizes stay on the exact kernels).  Same persistent CTAs, slot ring and release/acquire counters as k_imdct_fused; the
// transform is the generic form of the same factorisation: pre-twiddle, radix-4 (+ one radix-2) Stockham FFT of N/4 complex points by
// one warp inside the frame's slot (two ping-pong buffers of N/2 floats), post-twiddle into u[0 .. N/2); twiddles come
// from the setup's per-block-size tables (L1-resident).  Slot = N_long floats per channel.
// ------------------------------------------------------------------------------------------------
struct GenericParams {
    LaunchArgs a;
    int frames_per_cta;
    int n_slots;
    int slot_floats;
};

__device__ __forceinline__ void generic_transform(int lane, int N, const float* spec, const float2* tw, const float2* fft, float* slot) {
    const int M = N >> 1, Q = N >> 2;
    int lq = 0; while ((1 << lq) < Q) ++lq;                                  // log2 Q
    const int stages = (lq + 1) >> 1;                                       // radix-4 passes plus one radix-2 pass when log2 Q is odd
    float2* A = reinterpret_cast<float2*>(slot);
    float2* B = A + Q;
    // the last pass must leave the spectrum in B (the post-twiddle writes u over A)
    float2* x = (stages & 1) ? A : B;
    float2* y = (stages & 1) ? B : A;
    for (int k = lane; k < Q; k += 32) {                                    // c[k] = (X[2k] + i X[M-1-2k]) tw[k]
        cpx c, t; c.x = spec[2 * k]; c.y = spec[M - 1 - 2 * k];
        const float2 w = tw[k]; t.x = w.x; t.y = w.y;
        c = cmul(c, t);
        x[k] = make_float2(c.x, c.y);
    }
    __syncwarp();
    auto ld = [](const float2& v) { cpx r; r.x = v.x; r.y = v.y; return r; };
    for (int lm = 0; lm < lq;) {                                            // Stockham autosort passes, forward DFT; m = 1 << lm sub-transforms done
        const int m = 1 << lm;
        if (lq - lm >= 2) {                                                 // radix 4: l = Q / (4 m)
            const int lmQ = Q >> (lm + 2);                                  // l
            for (int b = lane; b < (Q >> 2); b += 32) {
                const int j = b >> lm, k = b & (m - 1);
                const int i0 = k + (j << lm), st = lmQ << lm;               // stride l * m
                const cpx c0 = ld(x[i0]), c1 = ld(x[i0 + st]), c2 = ld(x[i0 + 2 * st]), c3 = ld(x[i0 + 3 * st]);
                const cpx d0 = cadd(c0, c2), d1 = csub(c0, c2), d2 = cadd(c1, c3), d3 = cmul_mi(csub(c1, c3));
                const int tj = j << lm;                                     // w1 = exp(-2 pi i j m / Q), w2 = w1^2, w3 = w1^3
                const cpx w1 = ld(fft[tj]), w2 = ld(fft[2 * tj]), w3 = ld(fft[3 * tj]);
                const cpx o0 = cadd(d0, d2), o1 = cmul(cadd(d1, d3), w1), o2 = cmul(csub(d0, d2), w2), o3 = cmul(csub(d1, d3), w3);
                const int o = k + (j << (lm + 2));
                y[o] = make_float2(o0.x, o0.y); y[o + m] = make_float2(o1.x, o1.y);
                y[o + 2 * m] = make_float2(o2.x, o2.y); y[o + 3 * m] = make_float2(o3.x, o3.y);
            }
            lm += 2;
        } else {                                                            // radix 2: l = Q / (2 m)
            const int st = (Q >> (lm + 1)) << lm;
            for (int b = lane; b < (Q >> 1); b += 32) {
                const int j = b >> lm, k = b & (m - 1);
                const int i0 = k + (j << lm);
                const cpx c0 = ld(x[i0]), c1 = ld(x[i0 + st]);
                const cpx s0 = cadd(c0, c1), s1 = cmul(csub(c0, c1), ld(fft[j << lm]));
                const int o = k + (j << (lm + 1));
                y[o] = make_float2(s0.x, s0.y); y[o + m] = make_float2(s1.x, s1.y);
            }
            lm += 1;
        }
        __syncwarp();
        float2* tmp = x; x = y; y = tmp;
    }
    for (int n = lane; n < Q; n += 32) {                                    // D[n] = C[n] tw[n]; u[2n] = Re D, u[M-1-2n] = -Im D
        cpx c, t; const float2 v = x[n], w = tw[n];
        c.x = v.x; c.y = v.y; t.x = w.x; t.y = w.y;
        c = cmul(c, t);
        slot[2 * n] = c.x; slot[M - 1 - 2 * n] = -c.y;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(FUSED_THREADS, 1) k_imdct_generic(GenericParams p) {
    NVB_DYN_SMEM(smem_raw);
    const LaunchArgs& a = p.a;
    const DevSetup& S = a.S;
    const int C = S.channels;
    const int NS = p.n_slots, SLOT = p.slot_floats;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    float* s_slots = reinterpret_cast<float*>(smem_raw);
    DevFrame* s_fr = reinterpret_cast<DevFrame*>(s_slots + (size_t)NS * C * SLOT);
    int* s_full = reinterpret_cast<int*>(s_fr + NS);
    int* s_empty = s_full + NS;

    const int lo = a.frame_lo + blockIdx.x * p.frames_per_cta;
    int hi = lo + p.frames_per_cta; if (hi > a.frame_lo + a.n_frames) hi = a.frame_lo + a.n_frames;
    nvb_grid_dep_launch();
    if (lo >= hi) return;
    if (tid == 0) for (int s = 0; s < NS; s++) { s_full[s] = 0; s_empty[s] = 0; }
    __syncthreads();
    nvb_grid_dep_wait();
    int first = lo;
    {
        const DevFrame f0 = a.frames[lo];
        if (f0.prev >= 0 && (f0.ola_len > 0 || f0.kind != 0)) first = lo - 1;   // halo: previous block's tail is needed
    }
    float peak = 0.f;
    for (int x = first + warp; x < hi; x += FUSED_WARPS) {
        const int rel = x - first, slot = rel % NS, it = rel / NS;
        cnt_wait(&s_empty[slot], 2 * it);
        if (lane == 0) s_fr[slot] = a.frames[x];
        __syncwarp();
        const DevFrame f = s_fr[slot];
        float* slots_f = s_slots + (size_t)slot * C * SLOT;
        if (f.kind == 0) {
            const int bi = f.n == S.bs[1] ? 1 : 0;
            const int M = f.n >> 1;
            for (int c = 0; c < C; c++) {
                float* slotc = slots_f + (size_t)c * SLOT;
                const float* spec = a.spectrum + (size_t)f.spec_off + (size_t)c * M;
                if (!((f.exec_mask >> c) & 1u)) { for (int i = lane; i < M; i += 32) slotc[i] = spec[i]; }   // raw residue values (Mapping.cs:192-196)
                else generic_transform(lane, f.n, spec, S.tw[bi], S.fft[bi], slotc);
            }
        }
        __syncwarp();
        if (lane == 0) cnt_signal(&s_full[slot]);
        if (rel >= 1) cnt_wait(&s_full[(rel - 1) % NS], (rel - 1) / NS + 1);
        if (x >= lo) {
            const DevFrame* pf = nullptr; const float* slots_p = nullptr;
            if (f.prev >= 0 && (f.ola_len > 0 || f.kind != 0)) {
                const int pslot = (f.prev - first) % NS;
                pf = &s_fr[pslot]; slots_p = s_slots + (size_t)pslot * C * SLOT;
            }
            if (f.out_end > f.out_begin) emit_samples<false>(a, S, f, pf, slots_f, slots_p, SLOT, lane, peak, 0, C, nullptr, 0);
            if (x == a.carry_frame && a.carry_out && f.kind == 0) {
                for (int idx = lane; idx < f.n * C; idx += 32) {            // keep the last windowed block for the next batch (StreamDecoder.cs:455-461)
                    const int c = idx / f.n, i = idx - c * f.n;
                    a.carry_out[(size_t)c * S.bs[1] + i] = slot_z(S, f, slots_f + (size_t)c * SLOT, c, i, false);
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            cnt_signal(&s_empty[slot]);
            if (rel >= 1) cnt_signal(&s_empty[(rel - 1) % NS]);
        }
    }
    if (__any_sync(0xffffffffu, peak > 0.99999994f) && lane == 0) atomicOr(&a.counters->clipped, 1);
}

static int fused_sm_count(int dev_slot) {
    static std::atomic<int> cache[64];
    int v = cache[dev_slot].load(std::memory_order_relaxed);
    if (v == 0) {
        int dev = 0; cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[dev_slot].store(v, std::memory_order_relaxed);
    }
    return v;
}

static int generic_slots(int C, int bs1) {
    const size_t per = (size_t)C * bs1 * sizeof(float) + sizeof(DevFrame) + 2 * sizeof(int);
    int ns = (int)((227 * 1024 - 64) / per);
    if (ns > FUSED_WARPS + 2) ns = FUSED_WARPS + 2;
    return ns;
}
static bool generic_supported(const BlobHeader& h) {
    return h.bs[0] >= 256 && h.bs[1] <= 8192 && h.channels >= 1 && h.channels <= NVB_MAX_CHANNELS && generic_slots(h.channels, h.bs[1]) >= 3;
}

// ------------------------------------------------------------------------------------------------
// Slots of the ring.  A slot is free again when its unit and the unit that overlaps onto it are done: unit v + 1 for whole frames, unit
// v + U for channel-pair units (U pairs per frame) -- so FUSED_WARPS + U + 1 slots let every warp start its next unit without waiting
// for a unit that is still in flight (round 2: with 18 slots for six channels, 40 % of the instructions k_imdct_fused_t<true> executed
// were polling for a slot, profiles/r02_b_config3_ncu.txt).
static int fused_slots(int C, int units_per_frame = 1) {
    const size_t fixed = FusedTables::FLOATS * sizeof(float) + 64;
    const size_t per = (size_t)C * FUSED_SLOT_FLOATS * sizeof(float) + sizeof(DevFrame) + 2 * sizeof(int);
    int ns = (int)((FUSED_SMEM_LIMIT - fixed) / per);
    static const int extra = std::getenv("NVB_FUSED_EXTRA_SLOTS") ? std::atoi(std::getenv("NVB_FUSED_EXTRA_SLOTS")) : 0;   // experiment hook
    const int want = FUSED_WARPS + units_per_frame + 1 + (extra > 0 ? extra : 0);
    if (ns > want) ns = want;
    return ns;
}
static size_t fused_smem(int C, int NS) {
    return FusedTables::FLOATS * sizeof(float) + (size_t)NS * ((size_t)C * FUSED_SLOT_FLOATS * sizeof(float) + sizeof(DevFrame) + 2 * sizeof(int)) + 32;
}

static bool classic_supported(const BlobHeader& h) {
    return h.bs[1] == FUSED_LONG_N && h.bs[0] == FUSED_SHORT_N && h.off_fused_tab != 0 && h.channels >= 1 && h.channels <= NVB_MAX_CHANNELS &&
           fused_slots(h.channels) >= 3;
}
bool fused_supported(const BlobHeader& h, const DevFrame*, int) {
    static const bool no_generic = std::getenv("NVB_NO_GENERIC_IMDCT") != nullptr;        // test hook: other block sizes on the exact kernels
    return classic_supported(h) || (!no_generic && generic_supported(h));
}

static int launch_imdct_generic(const LaunchArgs& a, void* stream) {
    const int C = a.S.channels;
    GenericParams p; p.a = a; p.n_slots = generic_slots(C, a.S.bs[1]); p.slot_floats = a.S.bs[1];
    const size_t smem = (size_t)p.n_slots * ((size_t)C * p.slot_floats * sizeof(float) + sizeof(DevFrame) + 2 * sizeof(int)) + 32;
    static std::atomic<size_t> configured_by_dev[64];
    int dev_slot = 0; cudaGetDevice(&dev_slot); if (dev_slot < 0 || dev_slot >= 64) dev_slot = 0;
    if (!nvb_ensure_smem(configured_by_dev[dev_slot], smem, [&]() { return cudaFuncSetAttribute(k_imdct_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; })) return -1;
    const int num_sms = fused_sm_count(dev_slot);
    int fpc = (a.n_frames + num_sms - 1) / num_sms;
    if (fpc < 8) fpc = 8;
    p.frames_per_cta = fpc;
    NVB_LAUNCH(k_imdct_generic, (a.n_frames + fpc - 1) / fpc, FUSED_THREADS, smem, stream, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_imdct_fused(const LaunchArgs& a, const DevFrame* host_frames, void* stream) {
    if (a.n_frames <= 0) return 0;
    if (!(a.S.bs[0] == FUSED_SHORT_N && a.S.bs[1] == FUSED_LONG_N && a.S.fused_tab)) return launch_imdct_generic(a, stream);
    const int C = a.S.channels;
    // channel-pair units for even channel counts above two -- unless the launch contains a drain (its successor reaches two
    // blocks back, which the unit ring only covers frame by frame)
    static const bool no_grouped = std::getenv("NVB_FUSED_NO_GROUPED") != nullptr;           // test hook
    bool grouped = C > 2 && (C & 1) == 0 && !no_grouped && host_frames != nullptr;
    if (grouped) for (int i = a.frame_lo; i < a.frame_lo + a.n_frames; i++) if (host_frames[i].kind != 0 || (host_frames[i].prev >= 0 && host_frames[i].prev != i - 1)) { grouped = false; break; }
    const int G = grouped ? 2 : C;
    FusedParams p; p.a = a; p.n_slots = fused_slots(G, grouped ? C / 2 : 1);
    // stress hooks (tests): a ring of as few as three slots, pseudo-random pauses between the protocol steps
    const char* env_slots = std::getenv("NVB_FUSED_SLOTS"); const char* env_skew = std::getenv("NVB_FUSED_SKEW");
    if (env_slots) { const int ns = std::atoi(env_slots); if (ns >= 3 && ns < p.n_slots) p.n_slots = ns; }
    p.skew = env_skew ? std::atoi(env_skew) : 0;
    p.wfl = WfLayout(); p.wf_off = 0;
    const size_t smem = fused_smem(G, p.n_slots);
    static std::atomic<size_t> configured_by_dev[64];
    int dev_slot = 0; cudaGetDevice(&dev_slot); if (dev_slot < 0 || dev_slot >= 64) dev_slot = 0;
    if (!nvb_ensure_smem(configured_by_dev[dev_slot], smem, [&]() {
            return cudaFuncSetAttribute(k_imdct_fused_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                   cudaFuncSetAttribute(k_imdct_fused_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; })) return -1;
    const int num_sms = fused_sm_count(dev_slot);
    // persistent: one CTA per SM, each a contiguous run of frames (at least 8, so that the halo block stays cheap)
    const int ctas = num_sms * FUSED_CTAS_PER_SM;
    int fpc = (a.n_frames + ctas - 1) / ctas;
    if (fpc < 8) fpc = 8;
    p.frames_per_cta = fpc;
    const int grid = (a.n_frames + fpc - 1) / fpc;
    if (grouped) NVB_LAUNCH(k_imdct_fused_t<true>, grid, FUSED_THREADS, smem, stream, p);
    else NVB_LAUNCH(k_imdct_fused_t<false>, grid, FUSED_THREADS, smem, stream, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ------------------------------------------------------------------------------------------------
// One-kernel synthesis: k_imdct_fused_t<false, C> -- boundary records in, PCM out, one launch per batch.
static bool synth_supported(const DevSetup& S) {
    return S.bs[0] == FUSED_SHORT_N && S.bs[1] == FUSED_LONG_N && S.fused_tab && (S.channels == 1 || S.channels == 2) && S.spectrum_fast >= 3 &&
           S.f0_stride == 0 && S.max_posts <= 32 && S.magic && S.cls_cnt && S.run_modes;
}
int launch_synth_fused(const LaunchArgs& a, const DevFrame* host_frames, void* stream, bool forced) {
    (void)host_frames;
    if (a.n_frames <= 0) return 0;
    if (!synth_supported(a.S) || a.floor0) return -2;
    const int C = a.S.channels;
    {
        // Unforced, one kernel takes the launches it wins: those that fit ONE round of the CTA's warps (at most FUSED_WARPS units per
        // CTA including the halo block).  There a warp's frame is a single dependent chain -- spectrum, transforms, output -- and
        // one launch beats two (1024 frames: 22.7 vs 34.2 us, profiles/r02_b_one_kernel_ab.json); a second round costs the one-kernel
        // path more than the two-kernel path (4096 frames: 48.1 vs 44.5 us; 65 536 frames: 0.64 vs 0.52 ms).
        int dev = 0; cudaGetDevice(&dev); if (dev < 0 || dev >= 64) dev = 0;
        if (!forced && a.n_frames > (FUSED_WARPS - 1) * fused_sm_count(dev) * FUSED_CTAS_PER_SM) return -2;
    }
    FusedParams p; p.a = a;
    p.wfl = wf_layout_slot(a.S, C);
    if (p.wfl.cta_bytes > 16 * 1024) return -2;
    const size_t scratch = 1024 + (size_t)p.wfl.cta_bytes + (size_t)FUSED_WARPS * p.wfl.total + 16;
    const size_t fixed = FusedTables::FLOATS * sizeof(float) + 64 + scratch;
    const size_t per = (size_t)C * FUSED_SLOT_FLOATS * sizeof(float) + sizeof(DevFrame) + 2 * sizeof(int);
    if (fixed + 6 * per > FUSED_SMEM_LIMIT) return -2;                       // fewer than six slots: the ring would serialise the warps
    int ns = (int)((FUSED_SMEM_LIMIT - fixed) / per);
    if (ns > FUSED_WARPS + 2) ns = FUSED_WARPS + 2;
    p.n_slots = ns;
    const char* env_slots = std::getenv("NVB_FUSED_SLOTS"); const char* env_skew = std::getenv("NVB_FUSED_SKEW");
    if (env_slots) { const int v = std::atoi(env_slots); if (v >= 3 && v < p.n_slots) p.n_slots = v; }
    p.skew = env_skew ? std::atoi(env_skew) : 0;
    p.wf_off = (int)((fused_smem(C, p.n_slots) + 15) & ~size_t(15));
    const size_t smem = (size_t)p.wf_off + 1024 + (size_t)p.wfl.cta_bytes + (size_t)FUSED_WARPS * p.wfl.total;
    static std::atomic<size_t> configured_by_dev[64];
    int dev_slot = 0; cudaGetDevice(&dev_slot); if (dev_slot < 0 || dev_slot >= 64) dev_slot = 0;
    if (!nvb_ensure_smem(configured_by_dev[dev_slot], smem, [&]() {
            return cudaFuncSetAttribute(k_imdct_fused_t<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                   cudaFuncSetAttribute(k_imdct_fused_t<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; })) return -1;
    const int num_sms = fused_sm_count(dev_slot);
    const int ctas = num_sms * FUSED_CTAS_PER_SM;
    int fpc = (a.n_frames + ctas - 1) / ctas;
    if (fpc < 8) fpc = 8;
    p.frames_per_cta = fpc;
    const int grid = (a.n_frames + fpc - 1) / fpc;
    if (C == 1) NVB_LAUNCH((k_imdct_fused_t<false, 1>), grid, FUSED_THREADS, smem, stream, p);
    else NVB_LAUNCH((k_imdct_fused_t<false, 2>), grid, FUSED_THREADS, smem, stream, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace nvb
