// nvb_kernels.cu -- sm_100a kernels of the Vorbis synthesis path (generic / exact path).
//
//   k_spectrum     : residue VQ gather (K1) + inverse coupling (K2) + Floor1 curve multiply (K3)
//                    compact boundary records -> dense spectrum [frame][channel][N/2]
//   k_imdct_exact  : inverse MDCT in the reference's stb dataflow, no FMA contraction (K4, exact)
//                    + window multiply -> windowed blocks [frame][channel][N]
//   k_ola          : overlap-add with the previous block's tail + clip + interleave (K5)
//
// The fused fast path (IMDCT + window + OLA + clip + interleave in one kernel) lives in
// nvb_fused.cu.  All arithmetic is in nvb_device_core.h; these are thread-mapping shells.
#if !defined(NVB_CPU_SHIM)
#include <cuda_runtime.h>
#endif
#include "nvb_device_core.h"

namespace nvb {

constexpr int SPEC_THREADS = 256;
constexpr int MDCT_THREADS = 256;
constexpr int OLA_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// K1+K2+K3: one CTA per frame, one thread per spectral bin (all channels of that bin in registers).
// HBM traffic per frame: compact inputs (classes + entries + posts, ~1-2 KB) + VQ table gathers
// (L2-resident) in, C*N/2 floats out, written coalesced per channel row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SPEC_THREADS) k_spectrum(LaunchArgs a) {
    NVB_DYN_SMEM(dyn_smem);
    __shared__ FloorSegs s_segs[NVB_MAX_CHANNELS];
    __shared__ float s_db[256];
    __shared__ uint32_t s_warp[SPEC_THREADS / 32];
    __shared__ int s_bad[2];

    const DevFrame f = a.frames[blockIdx.x];
    if (f.kind != 0) return;
    const DevSetup& S = a.S;
    const int C = S.channels;
    const int t = threadIdx.x, nt = SPEC_THREADS;
    const DevMode md = S.modes[f.mode];
    const DevMapping& mp = S.mappings[md.mapping];
    const DevResidue& R = S.residues[mp.residue];
    const DevFloor1& F = S.floors[mp.floor];
    const int N = f.n, n = N >> 1;
    uint32_t* prefix = reinterpret_cast<uint32_t*>(dyn_smem);
    const uint8_t* cls = a.classes + f.classes_off;
    const uint16_t* ent = a.entries + f.entries_off;

    ResGeom g; g.P = 0; g.Sx = 1; g.n_items = 0;
    if (f.res_decoded) g = residue_geom(R, N, C);

    if (t < 2) s_bad[t] = 0;
    s_db[t & 255] = S.db[t & 255];

    // ---- entry-stream prefix: where each (stage, partition, stream) item's entries start
    {
        int chunk = (g.n_items + nt - 1) / nt;
        int lo = t * chunk; if (lo > g.n_items) lo = g.n_items;
        int hi = lo + chunk; if (hi > g.n_items) hi = g.n_items;
        uint32_t sum = 0;
        for (int i = lo; i < hi; i++) { uint32_t c = residue_item_count(R, S.books, cls, g, i); prefix[i] = c; sum += c; }
        uint32_t incl = sum;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if ((t & 31) >= d) incl += o; }
        if ((t & 31) == 31) s_warp[t >> 5] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < (t >> 5); w++) base += s_warp[w];
        uint32_t run = base + incl - sum;
        for (int i = lo; i < hi; i++) { uint32_t c = prefix[i]; prefix[i] = run; run += c; }
    }
    // ---- floor curves: one lane per channel (serial over <= 64 posts)
    if (t < C) {
        if ((f.exec_mask >> t) & 1u)
            floor1_build(F, a.posts + ((size_t)f.api_index * C + t) * S.post_stride, n, s_segs[t]);
        else
            s_segs[t].n = 0;
    }
    __syncthreads();

    int bad_entry = 0, bad_floor = 0;
    for (int j = t; j < n; j += nt) {
        float r[NVB_MAX_CHANNELS];
        #pragma unroll
        for (int c = 0; c < NVB_MAX_CHANNELS; c++)
            r[c] = (c < C) ? residue_value(R, S.books, S.vq, cls, ent, f.entry_count, prefix, g, C, c, j, &bad_entry) : 0.f;
        for (int i = mp.n_coupling - 1; i >= 0; --i) {                      // Mapping.cs:137-182
            int m = mp.mag[i], an = mp.ang[i];
            if (((f.exec_mask >> m) | (f.exec_mask >> an)) & 1u) inverse_couple(r[m], r[an]);
        }
        #pragma unroll
        for (int c = 0; c < NVB_MAX_CHANNELS; c++) {
            if (c >= C) break;
            float v = r[c];
            if ((f.exec_mask >> c) & 1u) {                                  // Floor1.Apply, Floor1.cs:186-222
                if (s_segs[c].n > 0) {
                    int y = floor1_y(s_segs[c], j);
                    if (y < 0 || y > 255) { bad_floor = 1; y = y < 0 ? 0 : 255; }
                    v = NVB_FMUL(v, s_db[y]);
                } else v = 0.f;
            }
            a.spectrum[(size_t)f.spec_off + (size_t)c * n + j] = v;
        }
    }
    if (bad_entry) atomicOr(&s_bad[0], 1);
    if (bad_floor) atomicOr(&s_bad[1], 1);
    __syncthreads();
    if (t == 0) {
        if (s_bad[0]) atomicAdd(&a.counters->bad_entry, 1);
        if (s_bad[1]) atomicAdd(&a.counters->floor_range, 1);
    }
}

// ------------------------------------------------------------------------------------------------
// K4 exact: one CTA per (frame, channel); the reference's stb_vorbis IMDCT schedule cut into
// data-parallel steps with a barrier between them, all in shared memory (u[N] + v[N/2]).
// Bit-identical to Mdct.cs for every N (including its N = 64/128 behaviour).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MDCT_THREADS) k_imdct_exact(LaunchArgs a) {
    NVB_DYN_SMEM(sm_raw);
    float* sm = reinterpret_cast<float*>(sm_raw);
    const DevSetup& S = a.S;
    const int C = S.channels;
    const int fi = blockIdx.x / C, c = blockIdx.x - fi * C;
    const DevFrame f = a.frames[fi];
    if (f.kind != 0) return;
    const int t = threadIdx.x, nt = MDCT_THREADS;
    const int N = f.n, n2 = N >> 1;
    const int bi = S.modes[f.mode].block_flag ? 1 : 0;
    float* u = sm; float* v = sm + N;
    const float* spec = a.spectrum + (size_t)f.spec_off + (size_t)c * n2;
    const float* win = frame_window(S, f);
    float* out = a.blocks + 2 * (size_t)f.spec_off + (size_t)c * N;

    if (!((f.exec_mask >> c) & 1u)) {
        // Mapping.cs:192-196: no IMDCT, back half cleared, front half keeps the residue values
        for (int i = t; i < N; i += nt) out[i] = NVB_FMUL(i < n2 ? spec[i] : 0.f, win[i]);
        return;
    }
    for (int i = t; i < n2; i += nt) u[i] = spec[i];
    __syncthreads();
    const float* A = S.A[bi]; const float* B = S.B[bi]; const float* Ct = S.C[bi];
    mdct_step0(u, v, A, N, t, nt);              __syncthreads();
    mdct_step2(u, v, A, N, t, nt);              __syncthreads();
    const int passes = mdct_num_r2_passes(N);
    for (int l = 0; l < passes; l++) { mdct_step3_pass(u, A, N, l, t, nt); __syncthreads(); }
    mdct_ld654(u, A, N, t, nt);                 __syncthreads();
    mdct_step456(u, v, S.bitrev[bi], N, t, nt); __syncthreads();
    mdct_step7(v, Ct, N, t, nt);                __syncthreads();
    mdct_step8(u, v, B, N, t, nt);              __syncthreads();
    for (int i = t; i < N; i += nt) out[i] = NVB_FMUL(u[i], win[i]);   // Mode.cs:159-166
}

// ------------------------------------------------------------------------------------------------
// K5: overlap-add + clip + interleave (StreamDecoder.cs:532-541, 391-415).  One CTA per frame.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OLA_THREADS) k_ola(LaunchArgs a) {
    const DevSetup& S = a.S;
    const int C = S.channels;
    const DevFrame f = a.frames[blockIdx.x];
    const int len = f.out_end - f.out_begin;
    if (len <= 0) return;
    const float* cur; int cstride;
    const float* prev = nullptr; int pstride = 0;
    if (f.kind == 0) {
        cur = a.blocks + 2 * (size_t)f.spec_off; cstride = f.n;
        if (f.ola_len > 0) {
            if (f.prev >= 0) { const DevFrame pf = a.frames[f.prev]; prev = a.blocks + 2 * (size_t)pf.spec_off; pstride = pf.n; }
            else if (f.prev == PREV_CARRY) { prev = a.carry_in; pstride = S.bs[1]; }
        }
    } else if (f.prev >= 0) {                   // drain of a block of this batch (StreamDecoder.cs:352-356)
        const DevFrame pf = a.frames[f.prev]; cur = a.blocks + 2 * (size_t)pf.spec_off; cstride = pf.n;
    } else { cur = a.carry_in; cstride = S.bs[1]; }
    int clipped = 0;
    const int total = len * C;
    for (int idx = threadIdx.x; idx < total; idx += OLA_THREADS) {
        int s = idx / C, c = idx - s * C;
        int i = f.out_begin + s;
        float v = cur[(size_t)c * cstride + i];
        int o = i - f.start;
        if (prev && o >= 0 && o < f.ola_len) v = NVB_FADD(v, prev[(size_t)c * pstride + f.prev_valid + o]);
        if (a.clip) v = clip_value(v, clipped);
        a.pcm[((size_t)f.pcm_off + s) * C + c] = v;
    }
    if (__syncthreads_or(clipped) && threadIdx.x == 0) atomicOr(&a.counters->clipped, 1);
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static size_t spectrum_smem(const DevSetup& S) { return (size_t)(S.max_items > 0 ? S.max_items : 1) * sizeof(uint32_t); }

int launch_spectrum(const LaunchArgs& a, void* stream) {
    if (a.n_frames <= 0) return 0;
    size_t smem = spectrum_smem(a.S);
    static size_t configured = 0;
    if (smem > 48 * 1024 - 8192 && smem > configured) {
        if (cudaFuncSetAttribute(k_spectrum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        configured = smem;
    }
    NVB_LAUNCH(k_spectrum, a.n_frames, SPEC_THREADS, smem, stream, a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_imdct_exact(const LaunchArgs& a, void* stream) {
    if (a.n_frames <= 0) return 0;
    size_t smem = (size_t)(a.S.bs[1] + a.S.bs[1] / 2) * sizeof(float);
    static size_t configured = 0;
    if (smem > 40 * 1024 && smem > configured) {
        if (cudaFuncSetAttribute(k_imdct_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        configured = smem;
    }
    NVB_LAUNCH(k_imdct_exact, a.n_frames * a.S.channels, MDCT_THREADS, smem, stream, a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_ola(const LaunchArgs& a, void* stream) {
    if (a.n_frames <= 0) return 0;
    NVB_LAUNCH(k_ola, a.n_frames, OLA_THREADS, 0, stream, a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace nvb
