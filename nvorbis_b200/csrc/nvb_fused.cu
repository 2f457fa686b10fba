// nvb_fused.cu -- fused IMDCT + window + overlap-add + clip + interleave for sm_100a (K4+K5).
//
// Replaces, per frame and channel: Mdct.Reverse (Mdct.cs:13-21,65-313), the window multiply of
// Mode.Decode (Mode.cs:159-166), StreamDecoder.OverlapBuffers (StreamDecoder.cs:532-541) and
// ClippingCopyBuffer / CopyBuffer (StreamDecoder.cs:391-415).
//
// Mapping onto the chip
//   * one CTA owns a contiguous run of frames and walks it in groups of G frames; one warp
//     transforms one (frame, channel) block: 512 complex points as 16 per lane, three radix-8
//     passes in registers, two exchanges through a warp-private 4.5 KB shared-memory slot
//     (bank-conflict-free strides 72 / 9), twiddles from shared-memory tables;
//   * the DCT-IV output u of every block stays in shared memory; the previous block's u is still
//     there (ring of G+1 slots), so the overlap-add never touches HBM: each spectrum float is
//     read once and each PCM float written once (16 384 B per stereo long frame);
//   * the first block of a run is recomputed as a halo (its output belongs to the previous CTA);
//   * output: two samples x two channels per thread as one float4 store, coalesced.
// No tensor cores: the IMDCT is FFT-structured, not a dense contraction.
#if !defined(NVB_CPU_SHIM)
#include <cuda_runtime.h>
#endif
#include "nvb_fused_core.h"

namespace nvb {

constexpr int FUSED_THREADS = 256;
constexpr int FUSED_WARPS = FUSED_THREADS / 32;

struct FusedParams {
    LaunchArgs a;
    int frames_per_cta;
    int G;
};

__device__ __forceinline__ float clipf(float v, int& clipped) {
    float c = fminf(fmaxf(v, -0.99999994f), 0.99999994f);
    if (c != v) clipped = 1;
    return c;
}

// Windowed block value z[i] = y[i] * window[i] of the block held in `slot` (Mode.cs:159-166).
__device__ __forceinline__ float slot_z(const DevSetup& S, const DevFrame& f, const float* slot, int c, int i) {
    const bool exec = (f.exec_mask >> c) & 1u;
    return fused_y(slot, exec, f.n, i) * frame_window(S, f)[i];
}

__global__ void __launch_bounds__(FUSED_THREADS, 2) k_imdct_fused(FusedParams p) {
    NVB_DYN_SMEM(smem_raw);
    const LaunchArgs& a = p.a;
    const DevSetup& S = a.S;
    const int C = S.channels;
    const int G = p.G;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    float2* s_tw1  = reinterpret_cast<float2*>(smem_raw);           // 512
    float2* s_w512 = s_tw1 + 512;                                   // 512
    float2* s_tw0  = s_w512 + 512;                                  // 64
    float2* s_w64  = s_tw0 + 64;                                    // 64
    float*  s_win  = reinterpret_cast<float*>(s_w64 + 64);          // 1024: rising long slope
    float*  s_slots = s_win + 1024;                                 // (G+1)*C slots
    DevFrame* s_fr = reinterpret_cast<DevFrame*>(s_slots + (size_t)(G + 1) * C * FUSED_SLOT_FLOATS);

    const int lo = blockIdx.x * p.frames_per_cta;
    int hi = lo + p.frames_per_cta; if (hi > a.n_frames) hi = a.n_frames;
    if (lo >= hi) return;

    for (int i = tid; i < 512; i += FUSED_THREADS) { s_tw1[i] = S.tw[1][i]; s_w512[i] = S.fft[1][i]; }
    if (tid < 64) { s_tw0[tid] = S.tw[0][tid]; s_w64[tid] = S.fft[0][tid]; }
    for (int i = tid; i < 1024; i += FUSED_THREADS) s_win[i] = S.win_long[3 * (size_t)S.bs[1] + i];

    int first = lo;
    {
        const DevFrame f0 = a.frames[lo];
        if (f0.prev >= 0 && (f0.ola_len > 0 || f0.kind != 0)) first = lo - 1;   // halo: previous block's tail is needed
    }
    int clipped = 0;

    for (int base = first; base < hi; base += G) {
        int cnt = hi - base; if (cnt > G) cnt = G;
        __syncthreads();                                   // previous output phase done with s_fr / slots
        if (tid < cnt) s_fr[(base + tid - first) % (G + 1)] = a.frames[base + tid];
        __syncthreads();

        // ---------------- transform phase: one warp per (frame, channel) -------------------------
        for (int x = warp; x < cnt * C; x += FUSED_WARPS) {
            const int fi = x / C, c = x - fi * C;
            const int r = (base + fi - first) % (G + 1);
            const DevFrame& f = s_fr[r];
            if (f.kind != 0) continue;
            float* slot = s_slots + (size_t)(r * C + c) * FUSED_SLOT_FLOATS;
            const int M = f.n >> 1;
            const float* spec = a.spectrum + (size_t)f.spec_off + (size_t)c * M;
            if (!((f.exec_mask >> c) & 1u)) {
                for (int i = lane; i < M; i += 32) slot[i] = spec[i];         // raw residue values (Mapping.cs:192-196)
            } else if (f.n == FUSED_LONG_N) {
                float2* ex = reinterpret_cast<float2*>(slot);
                LongRegs R;
                long_phase1(lane, reinterpret_cast<const float2*>(spec), s_tw1, s_w512, ex);
                __syncwarp();
                long_phase2_load(lane, ex, R);
                __syncwarp();
                long_phase2_store(lane, s_w512, ex, R);
                __syncwarp();
                long_phase3_load(lane, ex, R);
                __syncwarp();
                long_phase3_store(lane, s_tw1, ex, R);
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
                short_phase3_store(lane, s_tw0, slot, R);
            }
        }
        __syncthreads();

        // ---------------- output phase: all threads, frame by frame ------------------------------
        for (int fi = 0; fi < cnt; fi++) {
            const int x = base + fi;
            if (x < lo) continue;                                              // halo block: tail only
            const int r = (x - first) % (G + 1);
            const DevFrame& f = s_fr[r];
            const int len = f.out_end - f.out_begin;
            const float* slots_f = s_slots + (size_t)r * C * FUSED_SLOT_FLOATS;
            const DevFrame* pf = nullptr; const float* slots_p = nullptr;
            if (f.prev >= 0 && (f.ola_len > 0 || f.kind != 0)) {
                const int rp = (f.prev - first) % (G + 1);
                pf = &s_fr[rp]; slots_p = s_slots + (size_t)rp * C * FUSED_SLOT_FLOATS;
            }
            const bool fast = (C == 2) && f.kind == 0 && pf && f.n == FUSED_LONG_N && pf->n == FUSED_LONG_N && f.window == 3 &&
                              (pf->window & 2) && f.start == 0 && f.out_begin == 0 && f.out_end == 1024 && f.ola_len == 1024 &&
                              f.prev_valid == 1024 && f.exec_mask == 3u && pf->exec_mask == 3u && a.clip;
            if (fast) {
                // long block after long block, both channels live (Mode.cs:44-50 window 3):
                // out[i] = S[i]*yL[i] + S[1023-i]*yR[i],  yL from this block's u, yR from the previous block's u
                float* out = a.pcm + (size_t)f.pcm_off * 2;
                const bool al16 = (f.pcm_off & 1) == 0;
                for (int q = tid; q < 512; q += FUSED_THREADS) {
                    const int i = 2 * q;
                    const float2 wl = *reinterpret_cast<const float2*>(s_win + i);
                    const float2 wr = *reinterpret_cast<const float2*>(s_win + 1022 - i);   // (S[1022-i], S[1023-i])
                    float o[4];
                    #pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const float* uf = slots_f + c * FUSED_SLOT_FLOATS;
                        const float* up = slots_p + c * FUSED_SLOT_FLOATS;
                        float yl0, yl1, yr0, yr1;
                        if (i < 512) {
                            const float2 l2 = *reinterpret_cast<const float2*>(uf + 512 + i);
                            const float2 r2 = *reinterpret_cast<const float2*>(up + 510 - i);
                            yl0 = l2.x; yl1 = l2.y; yr0 = -r2.y; yr1 = -r2.x;
                        } else {
                            const float2 l2 = *reinterpret_cast<const float2*>(uf + 1534 - i);
                            const float2 r2 = *reinterpret_cast<const float2*>(up + i - 512);
                            yl0 = -l2.y; yl1 = -l2.x; yr0 = -r2.x; yr1 = -r2.y;
                        }
                        o[c]     = clipf(fmaf(wl.x, yl0, wr.y * yr0), clipped);
                        o[2 + c] = clipf(fmaf(wl.y, yl1, wr.x * yr1), clipped);
                    }
                    if (al16) *reinterpret_cast<float4*>(out + 2 * i) = make_float4(o[0], o[1], o[2], o[3]);
                    else { *reinterpret_cast<float2*>(out + 2 * i) = make_float2(o[0], o[1]); *reinterpret_cast<float2*>(out + 2 * i + 2) = make_float2(o[2], o[3]); }
                }
            } else if (len > 0) {
                const int total = len * C;
                for (int idx = tid; idx < total; idx += FUSED_THREADS) {
                    const int s = idx / C, c = idx - s * C;
                    const int i = f.out_begin + s;
                    float v;
                    if (f.kind == 0) {
                        v = slot_z(S, f, slots_f + c * FUSED_SLOT_FLOATS, c, i);
                        const int o = i - f.start;
                        if (f.ola_len > 0 && o >= 0 && o < f.ola_len) {                       // StreamDecoder.cs:532-541
                            if (pf) v += slot_z(S, *pf, slots_p + c * FUSED_SLOT_FLOATS, c, f.prev_valid + o);
                            else if (f.prev == PREV_CARRY) v += a.carry_in[(size_t)c * S.bs[1] + f.prev_valid + o];
                        }
                    } else {                                                                  // drain, StreamDecoder.cs:352-356
                        v = pf ? slot_z(S, *pf, slots_p + c * FUSED_SLOT_FLOATS, c, i) : a.carry_in[(size_t)c * S.bs[1] + i];
                    }
                    if (a.clip) v = clipf(v, clipped);
                    a.pcm[((size_t)f.pcm_off + s) * C + c] = v;
                }
            }
            if (x == a.carry_frame && a.carry_out && f.kind == 0) {
                // keep the last windowed block for the next batch (StreamDecoder.cs:455-461)
                for (int idx = tid; idx < f.n * C; idx += FUSED_THREADS) {
                    const int c = idx / f.n, i = idx - c * f.n;
                    a.carry_out[(size_t)c * S.bs[1] + i] = slot_z(S, f, slots_f + c * FUSED_SLOT_FLOATS, c, i);
                }
            }
        }
    }
    if (__syncthreads_or(clipped) && tid == 0) atomicOr(&a.counters->clipped, 1);
}

// ------------------------------------------------------------------------------------------------
static int fused_group(int C) { int g = 8 / C; return g < 1 ? 1 : g; }
static size_t fused_smem(int C, int G) {
    return (size_t)(512 + 512 + 64 + 64) * sizeof(float2) + 1024 * sizeof(float) +
           (size_t)(G + 1) * C * FUSED_SLOT_FLOATS * sizeof(float) + (size_t)(G + 1) * sizeof(DevFrame);
}

bool fused_supported(const BlobHeader& h, const DevFrame*, int) {
    return h.bs[1] == FUSED_LONG_N && h.bs[0] == FUSED_SHORT_N && h.channels >= 1 && h.channels <= NVB_MAX_CHANNELS;
}

int launch_imdct_fused(const LaunchArgs& a, const DevFrame*, void* stream) {
    if (a.n_frames <= 0) return 0;
    const int C = a.S.channels;
    FusedParams p; p.a = a; p.G = fused_group(C);
    const size_t smem = fused_smem(C, p.G);
    static size_t configured = 0;
    static int num_sms = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_imdct_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        configured = smem;
    }
    if (num_sms == 0) {
        int dev = 0; cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
    }
    // contiguous run of frames per CTA: enough CTAs for ~2 resident per SM, runs a multiple of G
    const int target_ctas = num_sms * 2;
    int fpc = (a.n_frames + target_ctas - 1) / target_ctas;
    fpc = ((fpc + p.G - 1) / p.G) * p.G;
    if (fpc < p.G) fpc = p.G;
    p.frames_per_cta = fpc;
    const int grid = (a.n_frames + fpc - 1) / fpc;
    NVB_LAUNCH(k_imdct_fused, grid, FUSED_THREADS, smem, stream, p);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace nvb
