// nvb_kernels.cu -- sm_100a kernels of the Vorbis synthesis path: the spectrum stage (K1+K2+K3) and the exact IMDCT path.
//
//   spectrum stage: compact boundary records -> dense spectrum [frame][channel][N/2]
//     k_spectrum_run    default (type 2 / mono type 1 residues, power-of-two sizes, 8-aligned partitions): a thread owns
//                       a run of 8 stream values; per-segment floor records with multiply-high division
//     k_spectrum_bins   type 2 residues with any channel count / partition alignment (Residue2.cs:25-27 truncation case)
//     k_spectrum_planes / k_spectrum_fast / k_spectrum_warp   earlier designs, kept as cross-checks (NVB_SPECTRUM_* hooks)
//     k_spectrum        everything else: odd sizes, type 0 floors (Floor0.cs:152-212)
//   exact path (NVB_RUN_EXACT, block sizes below 256):
//     k_imdct_exact     inverse MDCT in the reference's stb dataflow, no FMA contraction (K4) + window multiply
//     k_ola             overlap-add with the previous block's tail + clip + interleave (K5)
//
// The fused fast path (IMDCT + window + OLA + clip + interleave in one kernel: k_imdct_fused_t / k_imdct_generic) lives in
// nvb_fused.cu.  Shared arithmetic is in nvb_device_core.h.
#if !defined(NVB_CPU_SHIM)
#include <cuda_runtime.h>
#endif
#include <cstdlib>
#include <type_traits>
#include "nvb_device_core.h"
#include "nvb_wf_core.h"

namespace nvb {

#if !defined(NVB_CPU_SHIM)
// Monotonic event counters in shared memory: signal = release-add by one lane (after __syncwarp), wait = acquire-poll.
__device__ __forceinline__ void cnt_signal(int* c) {
    asm volatile("red.release.cta.shared::cta.add.s32 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(c)) : "memory");
}
__device__ __forceinline__ void cnt_wait(const int* c, int need) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(c);
    int v;
    for (;;) {
        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        if (v >= need) break;
        __nanosleep(32);
    }
}
#endif

constexpr int SPEC_THREADS = 256;
constexpr int MDCT_THREADS = 256;
constexpr int OLA_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// K1+K2+K3: one CTA per frame, one thread per spectral bin (all channels of that bin in registers).
// HBM traffic per frame: compact inputs (classes + entries + posts, ~1-2 KB) + VQ table gathers
// (L2-resident) in, C*N/2 floats out, written coalesced per channel row.
// ------------------------------------------------------------------------------------------------
// MAXC: NVB_FAST_CHANNELS for up to 8 channels, NVB_MAX_CHANNELS (the bin's channels then live in local memory) beyond.
template <int MAXC>
__global__ void __launch_bounds__(SPEC_THREADS) k_spectrum(LaunchArgs a) {
    NVB_DYN_SMEM(dyn_smem);
    __shared__ FloorSegs s_segs[MAXC];
    __shared__ float s_db[256];
    __shared__ uint32_t s_warp[SPEC_THREADS / 32];
    __shared__ int s_bad[2];

    nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    const DevFrame f = a.frames[a.frame_lo + blockIdx.x];
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
    // ---- floor curves
    const DevFloor0 F0 = S.floors0[mp.floor];
    float* s_f0q = nullptr;                                                 // type 0: [C][n] curve value per bark index
    if (F0.type == 1) {
        // type 1: one lane per channel (serial over <= 64 posts)
        if (t < C) {
            if ((f.exec_mask >> t) & 1u)
                floor1_build(F, a.posts + ((size_t)f.api_index * C + t) * S.post_stride, n, s_segs[t]);
            else
                s_segs[t].n = 0;
        }
        __syncthreads();
    } else {
        // type 0 (Floor0.Apply, Floor0.cs:152-212): the curve only depends on the bark index k = barkMap[bin], so it is
        // evaluated once per (channel, k) into shared memory; s_segs[c].n doubles as "has a curve" (Amp > 0)
        const int bi = md.block_flag ? 1 : 0;
        s_f0q = reinterpret_cast<float*>(prefix + ((S.max_items + 3) & ~3));
        float* s_f0c = s_f0q + (size_t)C * (S.bs[1] >> 1);                   // [C][256]: 2 cos(coeff)
        const float* wmap = S.f0_wmap + F0.wmap_off[bi];
        const int order = F0.order, kmax = F0.kmax[bi];
        if (t < C) {
            const float amp = a.floor0[((size_t)f.api_index * C + t) * S.f0_stride];
            s_segs[t].n = (((f.exec_mask >> t) & 1u) && amp > 0.f) ? 1 : 0;
        }
        __syncthreads();
        for (int i = t; i < C * order; i += nt) {
            const int c = i / order, j = i - c * order;
            if (s_segs[c].n) s_f0c[c * 256 + j] = NVB_FMUL(2.f, (float)cos((double)a.floor0[((size_t)f.api_index * C + c) * S.f0_stride + 1 + j]));   // Floor0.cs:166-169
        }
        __syncthreads();
        for (int i = t; i < C * (kmax + 1); i += nt) {
            const int c = i / (kmax + 1), k = i - c * (kmax + 1);
            if (!s_segs[c].n) continue;
            const float* cf = s_f0c + c * 256;
            const float amp = a.floor0[((size_t)f.api_index * C + c) * S.f0_stride];
            const float w = wmap[k];
            float p = .5f, q = .5f;
            int j;
            for (j = 1; j < order; j += 2) { q = NVB_FMUL(q, NVB_FSUB(w, cf[j - 1])); p = NVB_FMUL(p, NVB_FSUB(w, cf[j])); }
            if (j == order) {                                               // odd order filter; slightly asymmetric
                q = NVB_FMUL(q, NVB_FSUB(w, cf[j - 1]));
                p = NVB_FMUL(p, NVB_FMUL(p, NVB_FSUB(4.f, NVB_FMUL(w, w))));
                q = NVB_FMUL(q, q);
            } else {                                                        // even order filter; still symmetric
                p = NVB_FMUL(p, NVB_FMUL(p, NVB_FSUB(2.f, w)));
                q = NVB_FMUL(q, NVB_FMUL(q, NVB_FADD(2.f, w)));
            }
            q = NVB_FSUB(NVB_FDIV(amp, (float)sqrt((double)NVB_FADD(p, q))), (float)F0.amp_ofs);      // dB of this bark section
            s_f0q[c * n + k] = (float)exp((double)NVB_FMUL(q, 0.11512925f));                          // linear sample multiplier
        }
        __syncthreads();
    }
    const int32_t* f0_bark = S.f0_bark + F0.bark_off[md.block_flag ? 1 : 0];

    int bad_entry = 0, bad_floor = 0;
    for (int j = t; j < n; j += nt) {
        float r[MAXC];
        #pragma unroll
        for (int c = 0; c < MAXC; c++)
            r[c] = (c < C) ? residue_value(R, S.books, S.vq, cls, ent, f.entry_count, prefix, g, C, c, j, &bad_entry) : 0.f;
        for (int i = mp.n_coupling - 1; i >= 0; --i) {                      // Mapping.cs:137-182
            int m = mp.mag[i], an = mp.ang[i];
            if (((f.exec_mask >> m) | (f.exec_mask >> an)) & 1u) inverse_couple(r[m], r[an]);
        }
        #pragma unroll
        for (int c = 0; c < MAXC; c++) {
            if (c >= C) break;
            float v = r[c];
            if ((f.exec_mask >> c) & 1u) {                                  // Floor1.Apply, Floor1.cs:186-222 / Floor0.Apply, Floor0.cs:152-212
                if (s_segs[c].n > 0) {
                    if (s_f0q) v = NVB_FMUL(v, s_f0q[c * n + f0_bark[j]]);
                    else {
                        int y = floor1_y(s_segs[c], j);
                        if (y < 0 || y > 255) { bad_floor = 1; y = y < 0 ? 0 : 255; }
                        v = NVB_FMUL(v, s_db[y]);
                    }
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
// K1+K2+K3, fast path (DevSetup.spectrum_fast): same results as k_spectrum, organised for the machine.
//   phase A  warp per channel: Floor1 unwrap with the posts of one dependency level computed in parallel
//            (lane = post; Floor1.cs:224-297 is serial only along the lo/hi neighbour chains), step flags
//            by an OR reduction, then the x-sorted walk of Floor1.Apply (Floor1.cs:196-216) as a
//            ballot/popcount compaction into line-segment records;  one warp builds the entry-stream
//            prefix (where each (stage, partition, stream) item's VQ entries start).
//   phase B  (channel, segment) units over all warps, lanes over x: y(x) of RenderLineMulti
//            (Floor1.cs:316-341) in closed form, floor multiplier inverse_dB[y] -> shared memory.
//   phase C  thread per spectral bin: gather the VQ values that land on the bin in stage order (the adds of
//            Residue0/1/2.WriteVectors in the reference's order), inverse coupling, floor multiply, store.
// No integer division in phases B/C: partition and book sizes are powers of two here (host-checked).
// ------------------------------------------------------------------------------------------------
struct SegRec { int16_t x0, x1; int32_t y0, b; int16_t ady, sy; float rcp; };

// floor(num / den) for 0 <= num < 2^22 with rcp = 1.0f / den: float estimate (exact operands), one-step correction.
__device__ __forceinline__ int div_small(int num, int den, float rcp) {
    int q = __float2int_rz(__int2float_rn(num) * rcp);
    const int rem = num - q * den;
    if (rem >= den) ++q; else if (rem < 0) --q;
    return q;
}

// UnwrapPosts (Floor1.cs:224-297) by one warp, lane = post (and post + 32): the posts of one dependency level in parallel.
// Leaves finalY in fy[] and returns the step flags (bit i = stepFlags[i]); `count` is the clamped PostCount (< 2: nothing done).
__device__ __forceinline__ unsigned long long floor1_unwrap_warp(const DevFloor1& F, const int16_t* posts, int lane, int* fy, int& count) {
    count = posts[0];
    if (count > F.n_posts) count = F.n_posts;
    if (count < 2) return 0ull;                                           // PostCount == 0: the spectrum is cleared (Floor1.cs:220)
    const int H = count > 32 ? 2 : 1;                                    // posts lane and lane + 32
    int val[2]; unsigned long long contrib = 0ull;
    #pragma unroll
    for (int h = 0; h < 2; h++) { const int i = lane + 32 * h; val[h] = i < count ? posts[1 + i] : 0; }
    if (lane < 2) fy[lane] = val[0];
    __syncwarp();
    // per-post constants of this lane (setup data): neighbours, RenderPoint's x terms, dependency level
    int p_lo[2], p_hi[2], p_x0[2], p_adx[2], p_dx[2], p_lvl[2]; float p_rcp[2];
    #pragma unroll
    for (int h = 0; h < 2; h++) {
        const int i = lane + 32 * h;
        p_lvl[h] = 0;
        if (h < H && i >= 2 && i < count) {
            p_lo[h] = F.lo[i]; p_hi[h] = F.hi[i]; p_x0[h] = F.x[p_lo[h]]; p_adx[h] = F.x[p_hi[h]] - p_x0[h]; p_dx[h] = F.x[i] - p_x0[h];
            p_rcp[h] = F.rcp[i]; p_lvl[h] = F.level[i];
        }
    }
    for (int lvl = 1; lvl <= F.max_level; lvl++) {
        #pragma unroll
        for (int h = 0; h < 2; h++) {
            if (p_lvl[h] == lvl) {
                const int i = lane + 32 * h;
                const int lo = p_lo[h], hi = p_hi[h];
                // RenderPoint (Floor1.cs:299-314): y0 +- |dy| * (X - x0) / adx, the division by the setup constant adx
                const int adx = p_adx[h], y0 = fy[lo];
                const int dy = fy[hi] - y0, ady = dy < 0 ? -dy : dy;
                const int err = ady * p_dx[h];
                const int off = (unsigned)err < (1u << 22) ? div_small(err, adx, p_rcp[h]) : err / adx;
                const int predicted = dy < 0 ? y0 - off : y0 + off;
                const int v = val[h];
                const int highroom = F.range - predicted, lowroom = predicted;
                const int room = (highroom < lowroom ? highroom : lowroom) * 2;
                int out;
                if (v != 0) {
                    contrib |= (1ull << lo) | (1ull << hi) | (1ull << i);
                    if (v >= room) out = highroom > lowroom ? v - lowroom + predicted : predicted - v + highroom - 1;
                    else out = (v & 1) ? predicted - ((v + 1) >> 1) : predicted + (v >> 1);       // v > 0 here: (v % 2) == 1 <=> v & 1
                } else out = predicted;
                fy[i] = out;
            }
        }
        __syncwarp();
    }
    // stepFlags: 0 and 1 always; i when its own value is non-zero or a later post names it as a neighbour (Floor1.cs:253-257,292)
    unsigned lo32 = (unsigned)contrib, hi32 = (unsigned)(contrib >> 32);
    #pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { lo32 |= __shfl_xor_sync(0xffffffffu, lo32, d); hi32 |= __shfl_xor_sync(0xffffffffu, hi32, d); }
    return (((unsigned long long)hi32 << 32) | lo32) | 3ull;
}

__device__ __forceinline__ void floor1_segments_warp(const DevFloor1& F, const int16_t* posts, int n, int lane, int* fy, SegRec* segs, int* nseg_out) {
    int count;
    const unsigned long long flags = floor1_unwrap_warp(F, posts, lane, fy, count);
    if (count < 2) { if (lane == 0) *nseg_out = 0; return; }
    // sorted walk: position k (lane, lane + 32) holds post sort[k]
    int idx[2]; bool act[2];
    #pragma unroll
    for (int h = 0; h < 2; h++) { const int k = lane + 32 * h; idx[h] = k < count ? F.sort[k] : 0; act[h] = k < count && idx[h] < count && ((flags >> idx[h]) & 1ull); }
    const unsigned m0 = __ballot_sync(0xffffffffu, act[0]), m1 = __ballot_sync(0xffffffffu, act[1]);
    const unsigned long long M = ((unsigned long long)m1 << 32) | m0;      // bit 0 is always set (sort[0] = 0, x = 0)
    bool emit[2] = {false, false};
    #pragma unroll
    for (int h = 0; h < 2; h++) {
        const int k = lane + 32 * h;
        if (act[h] && k >= 1) {
            const unsigned long long below = M & ((1ull << k) - 1ull);
            const int pk = 63 - __clzll((long long)below);
            const int pidx = F.sort[pk];
            const int lx = F.x[pidx];
            if (lx < n) {                                                   // Floor1.cs:204 / :211
                emit[h] = true;
                const int ly = fy[pidx] * F.mult, hx = F.x[idx[h]], hy = fy[idx[h]] * F.mult;
                const int x1 = hx < n ? hx : n;                             // x clamped, y NOT re-interpolated (Floor1.cs:206)
                const int dy = hy - ly, adx = x1 - lx;
                const float rcp = 1.0f / (float)adx;
                const int ady = dy < 0 ? -dy : dy;
                const int ab = (unsigned)ady < (1u << 22) ? div_small(ady, adx, rcp) : ady / adx;      // |dy / adx|, truncated like the C# division
                SegRec r; r.x0 = (int16_t)lx; r.x1 = (int16_t)x1; r.y0 = ly; r.b = dy < 0 ? -ab : ab;
                r.ady = (int16_t)(ady - ab * adx); r.sy = (int16_t)(dy < 0 ? -1 : 1); r.rcp = rcp;
                segs[__popcll(below) - 1] = r;
            }
        }
    }
    const int n_emit = __popc(__ballot_sync(0xffffffffu, emit[0])) + __popc(__ballot_sync(0xffffffffu, emit[1]));
    if (lane == 0) {
        int ns = n_emit;
        const int kl = 63 - __clzll((long long)M);
        const int lidx = F.sort[kl];
        const int lx = F.x[lidx];
        if (lx < n) {                                                       // flat tail, Floor1.cs:213-216
            SegRec r; r.x0 = (int16_t)lx; r.x1 = (int16_t)n; r.y0 = fy[lidx] * F.mult; r.b = 0; r.ady = 0; r.sy = 1; r.rcp = 1.0f;
            segs[ns++] = r;
        }
        *nseg_out = ns;
    }
}

__global__ void __launch_bounds__(SPEC_THREADS) k_spectrum_fast(LaunchArgs a) {
    NVB_DYN_SMEM(dyn_smem);
    __shared__ float s_db[256];
    __shared__ int s_fy[NVB_FAST_CHANNELS][NVB_MAX_POSTS];
    __shared__ SegRec s_seg[NVB_FAST_CHANNELS * (NVB_MAX_POSTS + 1)];
    __shared__ int s_nseg[NVB_FAST_CHANNELS + 1];
    __shared__ int s_bad[2];

    nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    const DevFrame f = a.frames[a.frame_lo + blockIdx.x];
    if (f.kind != 0) return;
    const DevSetup& S = a.S;
    const int C = S.channels;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    constexpr int NW = SPEC_THREADS / 32;
    const DevMode md = S.modes[f.mode];
    const DevMapping& mp = S.mappings[md.mapping];
    const DevResidue& R = S.residues[mp.residue];
    const DevFloor1& F = S.floors[mp.floor];
    const int N = f.n, n = N >> 1;
    float* s_fl = reinterpret_cast<float*>(dyn_smem);                       // [C][n] floor multipliers
    uint32_t* prefix = reinterpret_cast<uint32_t*>(s_fl + (size_t)C * (S.bs[1] >> 1));
    const uint8_t* cls = a.classes + f.classes_off;
    const uint16_t* ent = a.entries + f.entries_off;

    ResGeom g; g.P = 0; g.Sx = 1; g.n_items = 0;
    if (f.res_decoded) g = residue_geom(R, N, C);
    if (t < 2) s_bad[t] = 0;
    s_db[t & 255] = S.db[t & 255];

    // ---- phase A
    if (warp == NW - 1) {
        // entry-stream prefix over the items in decode order (stage, partition, stream): per stage, lanes over
        // (partition, stream), warp scan of the per-item entry counts
        uint32_t run = 0;
        const int per_stage = g.P * g.Sx;
        for (int s = 0; s < R.stages && per_stage > 0; s++) {
            for (int base = 0; base < per_stage; base += 32) {
                const int i = base + lane;
                uint32_t c = 0;
                if (i < per_stage) {
                    int p = i, st = 0;
                    if (g.Sx > 1) { p = i / g.Sx; st = i - p * g.Sx; }
                    const int cl = cls[st * g.P + p];
                    if (cl < R.nclass) c = (uint32_t)R.cnt[cl][s];
                }
                uint32_t incl = c;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
                if (i < per_stage) prefix[s * per_stage + i] = run + incl - c;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
    }
    for (int c = warp; c < C; c += NW) {                                    // the prefix warp also takes a channel only when C == NW
        if ((f.exec_mask >> c) & 1u)
            floor1_segments_warp(F, a.posts + ((size_t)f.api_index * C + c) * S.post_stride, n, lane, s_fy[c], s_seg + c * (NVB_MAX_POSTS + 1), &s_nseg[c]);
        else if (lane == 0) s_nseg[c] = 0;
    }
    __syncthreads();

    // ---- phase B: floor curve rows; (channel, segment) units round-robin over the warps, lanes over x
    int bad_floor = 0;
    {
        int c = 0, first_u = 0, ns = s_nseg[0];
        int total = 0;
        for (int k = 0; k < C; k++) total += s_nseg[k];
        for (int u = warp; u < total; u += NW) {
            while (u - first_u >= ns) { first_u += ns; ++c; ns = s_nseg[c]; }
            const SegRec r = s_seg[c * (NVB_MAX_POSTS + 1) + (u - first_u)];
            const int len = r.x1 - r.x0;
            float* row = s_fl + c * n + r.x0;
            for (int k = lane; k < len; k += 32) {
                const int q = div_small(k * r.ady, len, r.rcp);             // k * ady < 2^20
                int y = r.y0 + k * r.b + r.sy * q;                          // RenderLineMulti in closed form (Floor1.cs:316-341)
                if ((unsigned)y > 255u) { bad_floor = 1; y = y < 0 ? 0 : 255; }
                row[k] = s_db[y];
            }
        }
    }
    __syncthreads();

    // ---- phase C: residue values in the reference's add order, inverse coupling, floor multiply
    int bad_entry = 0;
    const int pshift = R.pshift, pmask = R.psize - 1;
    float* spec_out = a.spectrum + (size_t)f.spec_off;
    if (false && R.type == 2 && g.P > 0) {
        // Residue type 2 (Residue2.cs:23-47): a warp per partition, lane = interleaved element; element j belongs to channel
        // j % C of bin ob + j / C.  C is a power of two here (it divides the power-of-two partition size), so the
        // channels of one bin sit in C adjacent lanes and inverse coupling is a lane exchange.
        const int cshift = ilog_u(C) - 1, cmask = C - 1;
        const int psize = R.psize, bins_per = psize >> cshift, ob0 = R.begin >> cshift;
        for (int p = warp; p < g.P; p += NW) {
            const int cl = cls[p];
            const int casc = cl < R.nclass ? R.cascade[cl] : 0;
            for (int j0 = 0; j0 < psize; j0 += 32) {
                const int j = j0 + lane;
                const bool live = j < psize;
                float acc = 0.f;
                for (int s = 0; s < R.stages; s++) {
                    if (!((casc >> s) & 1)) continue;
                    const int book = R.books[cl][s];
                    if (book < 0) continue;
                    const DevBook b = S.books[book];
                    const uint32_t base = prefix[s * g.P + p];
                    if (live) acc = NVB_FADD(acc, vq_fetch(b, S.vq, ent, base + (uint32_t)(j >> b.dshift), f.entry_count, j & (b.dims - 1), &bad_entry));
                }
                const int c = j & cmask, bin = ob0 + p * bins_per + (j >> cshift);
                const int grp = lane & ~cmask;
                for (int i = mp.n_coupling - 1; i >= 0; --i) {              // Mapping.cs:137-182
                    const int m = mp.mag[i], an = mp.ang[i];
                    if (!(((f.exec_mask >> m) | (f.exec_mask >> an)) & 1u)) continue;
                    float vm = __shfl_sync(0xffffffffu, acc, grp + m), va = __shfl_sync(0xffffffffu, acc, grp + an);
                    inverse_couple(vm, va);
                    if (c == m) acc = vm; else if (c == an) acc = va;
                }
                if (live) {
                    float v = acc;
                    if ((f.exec_mask >> c) & 1u) v = s_nseg[c] > 0 ? NVB_FMUL(v, s_fl[c * n + bin]) : 0.f;      // Floor1.Apply, Floor1.cs:186-222
                    spec_out[(size_t)c * n + bin] = v;
                }
            }
        }
        // bins no partition reaches keep the cleared value (Mapping.cs:108); coupling and the floor leave zero at zero
        const int lo_bin = ob0, hi_bin = ob0 + g.P * bins_per;
        for (int idx = t; idx < C * n; idx += SPEC_THREADS) {
            const int c = idx / n, bin = idx - c * n;
            if (bin < lo_bin || bin >= hi_bin) spec_out[(size_t)c * n + bin] = 0.f;
        }
    } else {
        for (int j = t; j < n; j += SPEC_THREADS) {
            float r[NVB_FAST_CHANNELS];
            #pragma unroll
            for (int c = 0; c < NVB_FAST_CHANNELS; c++) r[c] = 0.f;
            if (R.type == 2) {
                // interleaved position of channel 0 of this bin; all C channels sit in one partition (aligned, host-checked)
                const int q = j * C - R.begin;
                const int p = q >> pshift;
                if (g.P > 0 && q >= 0 && p < g.P) {
                    const int o = q & pmask;
                    const int cl = cls[p];
                    if (cl < R.nclass) {
                        const int casc = R.cascade[cl];
                        for (int s = 0; s < R.stages; s++) {
                            if (!((casc >> s) & 1)) continue;
                            const int book = R.books[cl][s];
                            if (book < 0) continue;
                            const DevBook b = S.books[book];
                            const uint32_t base = prefix[s * g.P + p];
                            #pragma unroll
                            for (int c = 0; c < NVB_FAST_CHANNELS; c++) {
                                if (c >= C) break;
                                const int e = o + c;
                                r[c] = NVB_FADD(r[c], vq_fetch(b, S.vq, ent, base + (uint32_t)(e >> b.dshift), f.entry_count, e & (b.dims - 1), &bad_entry));
                            }
                        }
                    }
                }
            } else {
            const int q = j - R.begin;
            const int p = q >> pshift;
            if (g.P > 0 && q >= 0 && p < g.P) {
                const int o = q & pmask;
                for (int c = 0; c < C; c++) {
                    const int cl = cls[c * g.P + p];
                    if (cl >= R.nclass) continue;
                    const int casc = R.cascade[cl];
                    float acc = 0.f;
                    for (int s = 0; s < R.stages; s++) {
                        if (!((casc >> s) & 1)) continue;
                        const int book = R.books[cl][s];
                        if (book < 0) continue;
                        const DevBook b = S.books[book];
                        const uint32_t base = prefix[(s * g.P + p) * g.Sx + c];
                        if (R.type == 1) acc = NVB_FADD(acc, vq_fetch(b, S.vq, ent, base + (uint32_t)(o >> b.dshift), f.entry_count, o & (b.dims - 1), &bad_entry));
                        else {                                              // type 0: element (dim, step) at dim*steps + step (Residue0.cs:193-199)
                            const int sshift = pshift - b.dshift;
                            acc = NVB_FADD(acc, vq_fetch(b, S.vq, ent, base + (uint32_t)(o & ((1 << sshift) - 1)), f.entry_count, o >> sshift, &bad_entry));
                        }
                    }
                    #pragma unroll
                    for (int k = 0; k < NVB_FAST_CHANNELS; k++) if (k == c) r[k] = acc;
                }
            }
            }
            for (int i = mp.n_coupling - 1; i >= 0; --i) {                  // Mapping.cs:137-182
                const int m = mp.mag[i], an = mp.ang[i];
                if (!(((f.exec_mask >> m) | (f.exec_mask >> an)) & 1u)) continue;
                float vm = 0.f, va = 0.f;
                #pragma unroll
                for (int k = 0; k < NVB_FAST_CHANNELS; k++) { if (k == m) vm = r[k]; if (k == an) va = r[k]; }
                inverse_couple(vm, va);
                #pragma unroll
                for (int k = 0; k < NVB_FAST_CHANNELS; k++) { if (k == m) r[k] = vm; if (k == an) r[k] = va; }
            }
            #pragma unroll
            for (int c = 0; c < NVB_FAST_CHANNELS; c++) {
                if (c >= C) break;
                float v = r[c];
                if ((f.exec_mask >> c) & 1u) v = s_nseg[c] > 0 ? NVB_FMUL(v, s_fl[c * n + j]) : 0.f;      // Floor1.Apply, Floor1.cs:186-222
                spec_out[(size_t)c * n + j] = v;
            }
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
// K1+K2+K3, plane path (the default when DevSetup.spectrum_fast == 2): residues that code ONE interleaved stream -- type 2, or a
// single channel of type 1 -- with power-of-two partition and book sizes.  The VQ vectors are copied whole:
//   phase A  as in k_spectrum_fast (floor unwrap per channel warp); the last warp compacts the coded
//            (stage, partition) items and the start of each item's entries into a list;
//   phase B  floor curve rows (channel-interleaved) ...
//   phase G  ... and, without a barrier in between, the residue: half a warp per item, lane = VQ entry, the entry's
//            `dims` consecutive values move with one vector load/store into the stage's plane.  A (stage, position)
//            is written by exactly one entry, so there are no atomics;
//   phase C  thread per group of G = max(4, C) interleaved values: planes summed in stage order (the float adds of
//            Residue1/2.WriteVectors in the reference's order, starting from the cleared +0), inverse coupling,
//            floor multiply, channel rows written with vector stores.
// ------------------------------------------------------------------------------------------------
struct ItemRec { uint16_t p; uint8_t s, cl; uint32_t base; };

template <int CT>
__global__ void __launch_bounds__(SPEC_THREADS) k_spectrum_planes(LaunchArgs a) {
    constexpr int G = CT > 4 ? CT : 4;
    constexpr int NW = SPEC_THREADS / 32;
    NVB_DYN_SMEM(dyn_smem);
    __shared__ float s_db[256];
    __shared__ int s_fy[CT][NVB_MAX_POSTS];
    __shared__ SegRec s_seg[CT * (NVB_MAX_POSTS + 1)];
    __shared__ int s_nseg[CT + 1];
    __shared__ int s_casc[NVB_MAX_CLASSES];
    __shared__ int s_nitems;
    __shared__ int s_bad[2];
    __shared__ int s_ready[2];                                              // [0]: item list complete, [1]: channels whose floor segments are complete

    nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    const DevFrame f = a.frames[a.frame_lo + blockIdx.x];
    if (f.kind != 0) return;
    const DevSetup& S = a.S;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const DevMode md = S.modes[f.mode];
    const DevMapping& mp = S.mappings[md.mapping];
    const DevResidue& R = S.residues[mp.residue];
    const DevFloor1& F = S.floors[mp.floor];
    const int N = f.n, n = N >> 1, span = CT * n;
    const int max_span = CT * (S.bs[1] >> 1);
    float* s_fl = reinterpret_cast<float*>(dyn_smem);                       // [n][CT] floor multipliers
    float* s_pl = s_fl + max_span;                                          // [stages][n][CT] residue planes
    ItemRec* s_items = reinterpret_cast<ItemRec*>(s_pl + (size_t)S.max_stages * max_span);
    uint8_t* s_cls = reinterpret_cast<uint8_t*>(s_items) + (((size_t)S.max_items * sizeof(ItemRec) + 15) & ~size_t(15));
    int4* s_ci = reinterpret_cast<int4*>(s_cls + ((S.max_items + 15) & ~15));   // per (class, stage): vq offset, dims, entries, entries per partition (16-byte aligned)
    const int st_n = R.stages > 0 ? R.stages : 1;
    const uint8_t* cls = a.classes + f.classes_off;
    const uint16_t* ent = a.entries + f.entries_off;

    ResGeom g; g.P = 0; g.Sx = 1; g.n_items = 0;
    if (f.res_decoded) g = residue_geom(R, N, CT);
    const int P = g.P;
    if (t < 2) { s_bad[t] = 0; s_ready[t] = 0; }
    s_db[t & 255] = S.db[t & 255];
    for (int i = t; i < R.nclass * st_n; i += SPEC_THREADS) {
        const int cl = i / st_n, st = i - cl * st_n;
        const int book = st < R.stages ? R.books[cl][st] : -1;
        int4 ci = make_int4(0, 1, 0, 0);
        if (book >= 0) { const DevBook b = S.books[book]; ci = make_int4((int)b.off, b.dims, b.entries, R.cnt[cl][st]); }
        s_ci[i] = ci;
        if (st == 0) s_casc[cl] = R.cascade[cl];
    }
    for (int p = t; p < P; p += SPEC_THREADS) { const int cl = cls[p]; s_cls[p] = cl < R.nclass ? (uint8_t)cl : (uint8_t)255; }

    __syncthreads();                                                        // tables staged, flags cleared

    // ---- phase A: the item list (last warp) and the floor segments (one warp per channel) are produced concurrently; the
    // other warps start on the residue as soon as the item list is there and render floor rows once the segments are
    // (release/acquire counters instead of a block barrier, so nobody waits for the slowest producer)
    if (warp == NW - 1) {
        uint32_t run = 0; int nitems = 0;
        const uint32_t lt = (1u << lane) - 1u;
        for (int st = 0; st < R.stages && P > 0; st++) {
            for (int base = 0; base < P; base += 32) {
                const int p = base + lane;
                uint32_t c = 0; int cl = 0;
                if (p < P) { cl = cls[p]; if (cl < R.nclass) c = (uint32_t)R.cnt[cl][st]; }
                uint32_t incl = c;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
                const unsigned m = __ballot_sync(0xffffffffu, c > 0);
                if (c > 0) { ItemRec r; r.p = (uint16_t)p; r.s = (uint8_t)st; r.cl = (uint8_t)cl; r.base = run + incl - c; s_items[nitems + __popc(m & lt)] = r; }
                nitems += __popc(m);
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        if (lane == 0) s_nitems = nitems;
        __syncwarp();
        if (lane == 0) cnt_signal(&s_ready[0]);
    }
    for (int c = warp; c < CT; c += NW) {
        if ((f.exec_mask >> c) & 1u)
            floor1_segments_warp(F, a.posts + ((size_t)f.api_index * CT + c) * S.post_stride, n, lane, s_fy[c], s_seg + c * (NVB_MAX_POSTS + 1), &s_nseg[c]);
        else if (lane == 0) s_nseg[c] = 0;
        __syncwarp();
        if (lane == 0) cnt_signal(&s_ready[1]);
    }
    int bad_floor = 0, bad_entry = 0;
    // ---- phase G: VQ vectors into the stage planes; half a warp per item, lane = entry
    {
        cnt_wait(&s_ready[0], 1);
        const int nitems = s_nitems;
        const int hl = lane & 15;
        auto move_entry = [&](const ItemRec& r, const int4& ci, int e) {
            const int dims = ci.y;
            const uint32_t ei = r.base + (uint32_t)e;
            const float* src = nullptr;
            if (ei < f.entry_count) {                                       // else never decoded: contributes nothing (Residue0.cs:164-170)
                const int en = ent[ei];
                if (en < ci.z) src = S.vq + ci.x + (size_t)en * dims; else bad_entry = 1;
            }
            float* d = s_pl + (size_t)r.s * max_span + R.begin + (int)r.p * R.psize + e * dims;
            if (dims == 2) *reinterpret_cast<float2*>(d) = src ? *reinterpret_cast<const float2*>(src) : make_float2(0.f, 0.f);
            else if (dims == 1) *d = src ? *src : 0.f;
            else for (int k = 0; k < dims; k += 4) *reinterpret_cast<float4*>(d + k) = src ? *reinterpret_cast<const float4*>(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        #pragma unroll 1
        for (int idx = warp * 2 + (lane >> 4); idx < nitems; idx += NW * 2) {
            const ItemRec r = s_items[idx];
            const int4 ci = s_ci[r.cl * st_n + r.s];
            if (hl < ci.w) move_entry(r, ci, hl);
            for (int e = hl + 16; e < ci.w; e += 16) move_entry(r, ci, e);   // more than 16 entries per partition: rare
        }
    }
    // ---- phase B: floor curve rows, channel-interleaved
    cnt_wait(&s_ready[1], CT);
    {
        int c = 0, first_u = 0, ns = s_nseg[0];
        int total = 0;
        for (int k = 0; k < CT; k++) total += s_nseg[k];
        for (int u = warp; u < total; u += NW) {
            while (u - first_u >= ns) { first_u += ns; ++c; ns = s_nseg[c]; }
            const SegRec r = s_seg[c * (NVB_MAX_POSTS + 1) + (u - first_u)];
            const int len = r.x1 - r.x0;
            const int dyabs = r.ady + (r.b < 0 ? -r.b : r.b) * len;         // |dy|: y(k) = y0 + sy * floor(k |dy| / adx)  (Floor1.cs:316-341 in closed form)
            float* row = s_fl + r.x0 * CT + c;
            for (int k = lane; k < len; k += 32) {
                const int q = div_small(k * dyabs, len, r.rcp);             // k |dy| < 2^22
                int y = r.y0 + r.sy * q;
                if ((unsigned)y > 255u) { bad_floor = 1; y = y < 0 ? 0 : 255; }
                row[k * CT] = s_db[y];
            }
        }
    }
    __syncthreads();

    // ---- phase C: sum the planes in stage order, inverse coupling, floor multiply, store
    float* spec_out = a.spectrum + (size_t)f.spec_off;
    const int pshift = R.pshift;
    constexpr int CSH = CT == 1 ? 0 : CT == 2 ? 1 : CT == 4 ? 2 : 3;
    bool execc[CT]; bool hasfl[CT];
    #pragma unroll
    for (int c = 0; c < CT; c++) { execc[c] = (f.exec_mask >> c) & 1u; hasfl[c] = s_nseg[c] > 0; }
    for (int gi = t; gi < span / G; gi += SPEC_THREADS) {
        const int pos = gi * G;
        float acc[G];
        #pragma unroll
        for (int k = 0; k < G; k++) acc[k] = 0.f;
        const int q = pos - R.begin, p = q >> pshift;
        if (q >= 0 && p < P) {
            const int cl = s_cls[p];
            if (cl != 255) {
                unsigned casc = (unsigned)s_casc[cl];
                while (casc) {
                    const int st = __ffs(casc) - 1; casc &= casc - 1;
                    if (st >= R.stages || s_ci[cl * st_n + st].w == 0) continue;
                    const float* pl = s_pl + (size_t)st * max_span + pos;
                    #pragma unroll
                    for (int k = 0; k < G; k += 4) {
                        const float4 v = *reinterpret_cast<const float4*>(pl + k);
                        acc[k] = NVB_FADD(acc[k], v.x); acc[k + 1] = NVB_FADD(acc[k + 1], v.y); acc[k + 2] = NVB_FADD(acc[k + 2], v.z); acc[k + 3] = NVB_FADD(acc[k + 3], v.w);
                    }
                }
            }
        }
        for (int i = mp.n_coupling - 1; i >= 0; --i) {                      // Mapping.cs:137-182
            const int m = mp.mag[i], an = mp.ang[i];
            if (!(((f.exec_mask >> m) | (f.exec_mask >> an)) & 1u)) continue;
            #pragma unroll
            for (int b = 0; b < G / CT; b++) {
                float vm = 0.f, va = 0.f;
                #pragma unroll
                for (int k = 0; k < CT; k++) { if (k == m) vm = acc[b * CT + k]; if (k == an) va = acc[b * CT + k]; }
                inverse_couple(vm, va);
                #pragma unroll
                for (int k = 0; k < CT; k++) { if (k == m) acc[b * CT + k] = vm; if (k == an) acc[b * CT + k] = va; }
            }
        }
        #pragma unroll
        for (int k = 0; k < G; k += 4) {
            const float4 fl = *reinterpret_cast<const float4*>(s_fl + pos + k);
            const float flv[4] = {fl.x, fl.y, fl.z, fl.w};
            #pragma unroll
            for (int e = 0; e < 4; e++) {
                const int c = (k + e) & (CT - 1);
                if (execc[c]) acc[k + e] = hasfl[c] ? NVB_FMUL(acc[k + e], flv[e]) : 0.f;      // Floor1.Apply, Floor1.cs:186-222
            }
        }
        const int bin0 = pos >> CSH;
        if (CT == 1) *reinterpret_cast<float4*>(spec_out + bin0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else if (CT == 2) {
            *reinterpret_cast<float2*>(spec_out + bin0) = make_float2(acc[0], acc[2]);
            *reinterpret_cast<float2*>(spec_out + n + bin0) = make_float2(acc[1], acc[3]);
        } else {
            #pragma unroll
            for (int c = 0; c < CT; c++) spec_out[(size_t)c * n + bin0] = acc[c];
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
// K1+K2+K3, run path (the default when DevSetup.spectrum_fast == 3): the same residues as the plane path whose
// partitions start on multiples of 8 values.  No planes, no floor rows, no item list: a thread owns a RUN of 8
// consecutive values of the interleaved stream -- 8 / C consecutive bins of every channel -- start to end:
//   phase A  a warp per channel unwraps the floor posts (floor1_unwrap_warp) and leaves the active-post mask of the
//            x-sorted walk of Floor1.Apply (Floor1.cs:196-216) plus the sorted (x, y) lists; the last warp turns the class
//            bytes into the entry-stream offset of every (stage, partition);
//   main     per thread: the VQ vectors of its run are whole vector loads (8 values lie inside one partition, and
//            inside one entry when dims >= 8), summed in stage order from +0 (the float adds of WriteVectors in the
//            reference's order); inverse coupling pairs sit in the same thread; the floor line is walked along the
//            run with the bin -> post table of the setup, y(x) of RenderLineMulti (Floor1.cs:316-341) in closed form;
//            one vector store per channel.
// ------------------------------------------------------------------------------------------------
#if !defined(NVB_CPU_SHIM)
__device__ __forceinline__ float rcp_estimate(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
static inline float rcp_estimate(float x) { return 1.0f / x; }
#endif

// Segment of the floor line that starts at an active post (x-sorted position k): RenderLineMulti(x0, y0, min(hx, n), hy)
// (Floor1.cs:206).  y(x) = y0 + sign(dy) * floor((x - x0) |dy| / adx); the division by adx is a multiply-high with
// m = floor(2^32 / adx) + 1, exact while (x - x0) |dy| adx < 2^32 (checked per segment: adx^2 |dy| < 2^32); m == 0 sends the
// bin to the plain division.
struct alignas(16) RunSeg { int32_t x0, y0, dy; uint32_t m; };

// Inverse coupling of one bin (Mapping.cs:145-181), branch-free: same compares, one add or subtract.
__device__ __forceinline__ void inverse_couple_sel(float& m, float& a) {
    const float M = m, A = a;
    const bool mp = M > 0.f, ap = A > 0.f;
    const float t = (mp == ap) ? NVB_FSUB(M, A) : NVB_FADD(M, A);
    m = ap ? M : t;
    a = ap ? t : M;
}

// Floor 1 of one channel by one warp for k_spectrum_run / k_spectrum_bins: UnwrapPosts, the active-post mask of Floor1.Apply's
// x-sorted walk (bit k: sorted position k starts a segment; 0 = no curve: the channel is cleared, Floor1.cs:220) and one
// RunSeg per active position.  `careful`: some segment needs the plain division or leaves inverse_dB_table's range.
__device__ __forceinline__ void floor1_run_segments_warp(const DevFloor1& F, const int16_t* posts, int n, int lane, int* fy, int* ys, RunSeg* seg, int* adx_out,
                                                         unsigned long long& mask, int& careful_any) {
    mask = 0ull; careful_any = 0;
    int count;
    const unsigned long long flags = floor1_unwrap_warp(F, posts, lane, fy, count);
    if (count < 2) return;
    // the walk of Floor1.Apply visits sorted positions 1 .. PostCount-1 and steps on flagged posts; position 0 is its start
    unsigned m[2];
    #pragma unroll
    for (int h = 0; h < 2; h++) {
        const int k = lane + 32 * h;
        int idx = 0; bool act = false;
        if (k < count) { idx = F.sort[k]; act = idx < count && ((flags >> idx) & 1ull); ys[k] = fy[idx < count ? idx : 0] * F.mult; }
        m[h] = __ballot_sync(0xffffffffu, act);
    }
    mask = (((unsigned long long)m[1] << 32) | m[0]) | 1ull;
    __syncwarp();
    bool careful = false;
    #pragma unroll
    for (int h = 0; h < 2; h++) {
        const int k = lane + 32 * h;
        if ((mask >> k) & 1ull) {
            RunSeg r; r.x0 = F.xs[k]; r.y0 = ys[k]; r.dy = 0; r.m = 1u;
            int adx = 1;
            const unsigned long long above = mask & ~(((1ull << k) << 1) - 1ull);
            if (above) {                                                    // else the flat tail, Floor1.cs:213-216
                const int hi = __ffsll((long long)above) - 1;
                const int hx = F.xs[hi];
                adx = (hx < n ? hx : n) - r.x0;                             // x clamped, y NOT re-interpolated (Floor1.cs:206)
                r.dy = ys[hi] - r.y0;
                const unsigned long long ady = (unsigned long long)(r.dy < 0 ? -(long long)r.dy : (long long)r.dy);
                if (adx <= 1) { adx = 1; r.m = 1u; }                        // one bin (or a post at or beyond n): (x - x0) = 0
                else r.m = ((unsigned long long)adx * (unsigned long long)adx * ady < (1ull << 32)) ? 0xffffffffu / (unsigned)adx + 1u : 0u;
            }
            seg[k] = r; adx_out[k] = adx;
            // y runs monotonically from y0 to y0 + dy: in range at both ends <=> in range everywhere
            if (r.x0 < n && (r.m == 0u || (unsigned)r.y0 > 255u || (unsigned)(r.y0 + r.dy) > 255u)) careful = true;
        }
    }
    careful_any = __any_sync(0xffffffffu, careful);
}

// MINB caps the registers: 32 (16 CTAs of 128 threads per SM) for one or two channels, 40 otherwise -- measured 39.0 vs 41.0 us
template <int CT, int NT, int MINB = (CT <= 2 ? 2048 : 1536) / NT>
__global__ void __launch_bounds__(NT, MINB) k_spectrum_run(LaunchArgs a) {
    constexpr int NW = NT / 32;
    constexpr int RB = 8 / CT;                                              // bins per channel in one run
    NVB_DYN_SMEM(dyn_smem);
    __shared__ float s_db[256];
    __shared__ int s_fy[CT][NVB_MAX_POSTS];
    __shared__ int s_ys[CT][NVB_MAX_POSTS];                                 // finalY * multiplier in x-sorted order
    __shared__ RunSeg s_seg[CT][NVB_MAX_POSTS];                             // segment that starts at sorted position k
    __shared__ int s_adx[CT][NVB_MAX_POSTS];
    __shared__ unsigned long long s_mask[CT];                               // bit k: sorted position k starts a segment; 0 = no floor
    __shared__ int s_bad[2];
    __shared__ int s_careful[CT];                                           // some segment of the channel needs the plain division / range clamp

    // Programmatic dependent launch: everything up to the first store only reads batch inputs and setup tables, none of
    // which the previous kernel of the stream writes, so it runs while that kernel drains; the wait sits in front of the
    // main loop, whose stores may overwrite a spectrum buffer the previous kernel still reads.
    nvb_grid_dep_launch();
    const DevFrame f = a.frames[a.frame_lo + blockIdx.x];
    if (f.kind != 0) { nvb_grid_dep_wait(); return; }
    const DevSetup& S = a.S;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const RunMode rm = S.run_modes[f.mode];
    const DevMapping& mp = S.mappings[rm.mapping];
    const DevFloor1& F = S.floors[rm.floor];
    const int N = f.n, n = N >> 1, span = CT * n;
    const int stages = rm.stages, st_n = stages > 0 ? stages : 1;
    const int rbegin = rm.rbegin, pshift = rm.pshift, nclass = rm.nclass;
    int P = 0;
    if (f.res_decoded) { const int e = rm.rend < span ? rm.rend : span; const int nn = e - rbegin; P = nn > 0 ? nn >> pshift : 0; }   // Residue0.cs:122-127
    uint32_t* s_base = reinterpret_cast<uint32_t*>(dyn_smem);               // [stage][partition]: where the item's entries start
    const CiRec* ci_tab = S.ci + rm.ci_off;                                 // (class, stage) records and class -> coded stages: setup tables, L1-resident
    const uint8_t* coded = S.residues[rm.residue].coded;
    const uint8_t* cls = a.classes + f.classes_off;
    const uint16_t* ent = a.entries + f.entries_off;
    const uint8_t* bin2k = S.bin2k + (size_t)rm.floor * (S.bs[1] >> 1);

    for (int i = t; i < 256; i += NT) s_db[i] = S.db[i];
    if (t < 2) s_bad[t] = 0;
    if (P > 0) for (uint32_t i = (uint32_t)t * 64u; i < f.entry_count; i += NT * 64u) prefetch_l1(ent + i);   // the frame's entries: 128 bytes per thread

    // ---- phase A
    if (warp == NW - 1 && P > 0) {
        uint32_t run = 0;
        for (int st = 0; st < stages; st++) {
            for (int base = 0; base < P; base += 32) {
                const int p = base + lane;
                uint32_t c = 0;
                if (p < P) { const int cl = cls[p]; if (cl < nclass) c = (uint32_t)ci_tab[cl * st_n + st].cnt; }
                uint32_t incl = c;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
                if (p < P) s_base[st * P + p] = run + incl - c;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
    }
    for (int c = warp; c < CT; c += NW) {
        unsigned long long mask = 0ull; int careful_any = 0;
        if ((f.exec_mask >> c) & 1u)
            floor1_run_segments_warp(F, a.posts + ((size_t)f.api_index * CT + c) * S.post_stride, n, lane, s_fy[c], s_ys[c], s_seg[c], s_adx[c], mask, careful_any);
        if (lane == 0) { s_mask[c] = mask; s_careful[c] = careful_any; }
    }
    __syncthreads();
    nvb_grid_dep_wait();

    // ---- main: one run of 8 stream values per thread
    float* spec_out = a.spectrum + (size_t)f.spec_off;
    const uint32_t ecount = f.entry_count;
    const int pmask = (1 << pshift) - 1;
    const bool posts32 = F.n_posts <= 32;
    unsigned long long fmask[CT]; bool careful[CT];
    #pragma unroll
    for (int c = 0; c < CT; c++) { fmask[c] = s_mask[c]; careful[c] = s_careful[c] != 0; }
    int bad_floor = 0, bad_entry = 0;
    for (int gi = t; gi < (span >> 3); gi += NT) {
        const int pos0 = gi << 3;
        const int bin0 = pos0 / CT;
        // sorted position of the last post at or below each bin of the run (setup table, one load)
        unsigned long long kword;
        if (RB == 8) kword = *reinterpret_cast<const unsigned long long*>(bin2k + bin0);
        else if (RB == 4) kword = *reinterpret_cast<const uint32_t*>(bin2k + bin0);
        else if (RB == 2) kword = *reinterpret_cast<const uint16_t*>(bin2k + bin0);
        else kword = bin2k[bin0];
        float acc[8];
        #pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0.f;
        const int q = pos0 - rbegin, p = q >> pshift;
        if (q >= 0 && p < P) {
            const int cl = cls[p];
            if (cl < nclass) {
                const int o = q & pmask;
                unsigned casc = coded[cl];
                while (casc) {
                    const int st = __ffs(casc) - 1; casc &= casc - 1;
                    const CiRec ci = ci_tab[cl * st_n + st];
                    const uint32_t eb = s_base[st * P + p];
                    const float* tab = S.vq + ci.off;
                    if (ci.dshift >= 3) {                                   // the run lies inside one entry
                        const uint32_t ei = eb + (uint32_t)(o >> ci.dshift);
                        if (ei < ecount) {                                  // else never decoded: contributes nothing (Residue0.cs:164-170)
                            const int en = ent[ei];
                            if (en < ci.entries) {
                                const float4* src = reinterpret_cast<const float4*>(tab + ((size_t)en << ci.dshift) + (o & ((1 << ci.dshift) - 1)));
                                const float4 v0 = src[0], v1 = src[1];
                                acc[0] = NVB_FADD(acc[0], v0.x); acc[1] = NVB_FADD(acc[1], v0.y); acc[2] = NVB_FADD(acc[2], v0.z); acc[3] = NVB_FADD(acc[3], v0.w);
                                acc[4] = NVB_FADD(acc[4], v1.x); acc[5] = NVB_FADD(acc[5], v1.y); acc[6] = NVB_FADD(acc[6], v1.z); acc[7] = NVB_FADD(acc[7], v1.w);
                            } else bad_entry = 1;
                        }
                    } else if (ci.dshift == 2) {
                        #pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const uint32_t ei = eb + (uint32_t)(o >> 2) + h;
                            if (ei < ecount) {
                                const int en = ent[ei];
                                if (en < ci.entries) {
                                    const float4 v = *reinterpret_cast<const float4*>(tab + ((size_t)en << 2));
                                    acc[4 * h] = NVB_FADD(acc[4 * h], v.x); acc[4 * h + 1] = NVB_FADD(acc[4 * h + 1], v.y);
                                    acc[4 * h + 2] = NVB_FADD(acc[4 * h + 2], v.z); acc[4 * h + 3] = NVB_FADD(acc[4 * h + 3], v.w);
                                } else bad_entry = 1;
                            }
                        }
                    } else if (ci.dshift == 1) {
                        #pragma unroll
                        for (int h = 0; h < 4; h++) {
                            const uint32_t ei = eb + (uint32_t)(o >> 1) + h;
                            if (ei < ecount) {
                                const int en = ent[ei];
                                if (en < ci.entries) {
                                    const float2 v = *reinterpret_cast<const float2*>(tab + ((size_t)en << 1));
                                    acc[2 * h] = NVB_FADD(acc[2 * h], v.x); acc[2 * h + 1] = NVB_FADD(acc[2 * h + 1], v.y);
                                } else bad_entry = 1;
                            }
                        }
                    } else {
                        #pragma unroll
                        for (int h = 0; h < 8; h++) {
                            const uint32_t ei = eb + (uint32_t)o + h;
                            if (ei < ecount) {
                                const int en = ent[ei];
                                if (en < ci.entries) acc[h] = NVB_FADD(acc[h], tab[en]); else bad_entry = 1;
                            }
                        }
                    }
                }
            }
        }
        for (int i = rm.n_coupling - 1; i >= 0; --i) {                      // Mapping.cs:137-182
            const int m = mp.mag[i], an = mp.ang[i];
            if (!(((f.exec_mask >> m) | (f.exec_mask >> an)) & 1u)) continue;
            if (CT == 2) {                                                  // (magnitude, angle) is (0, 1) or (1, 0)
                #pragma unroll
                for (int b = 0; b < RB; b++) { if (m == 0) inverse_couple_sel(acc[2 * b], acc[2 * b + 1]); else inverse_couple_sel(acc[2 * b + 1], acc[2 * b]); }
            } else {
                #pragma unroll
                for (int b = 0; b < RB; b++) {
                    float vm = 0.f, va = 0.f;
                    #pragma unroll
                    for (int k = 0; k < CT; k++) { if (k == m) vm = acc[b * CT + k]; if (k == an) va = acc[b * CT + k]; }
                    inverse_couple_sel(vm, va);
                    #pragma unroll
                    for (int k = 0; k < CT; k++) { if (k == m) acc[b * CT + k] = vm; if (k == an) acc[b * CT + k] = va; }
                }
            }
        }
        // per bin: mask of the sorted positions at or below the last post that is not beyond the bin
        unsigned long long below[RB];
        #pragma unroll
        for (int b = 0; b < RB; b++) {
            const unsigned kk = (unsigned)(kword >> (8 * b)) & 0xffu;
            below[b] = posts32 ? (unsigned long long)(0xffffffffu >> (31 - kk)) : (0xffffffffffffffffull >> (63 - kk));
        }
        #pragma unroll
        for (int c = 0; c < CT; c++) {
            if ((f.exec_mask >> c) & 1u) {                                  // Floor1.Apply, Floor1.cs:186-222
                const unsigned long long M = fmask[c];
                if (M == 0ull) {
                    #pragma unroll
                    for (int b = 0; b < RB; b++) acc[b * CT + c] = 0.f;
                } else if (!careful[c]) {
                    #pragma unroll
                    for (int b = 0; b < RB; b++) {
                        // the segment's start: last active position at or below the bin's post (bit 0 is set)
                        const int lo = posts32 ? 31 - __clz((int)((unsigned)M & (unsigned)below[b])) : 63 - __clzll((long long)(M & below[b]));
                        const RunSeg r = s_seg[c][lo];
                        const int sg = r.dy >> 31;                          // 0 or -1
                        const int qq = (int)__umulhi((unsigned)((bin0 + b - r.x0) * ((r.dy ^ sg) - sg)), r.m);
                        acc[b * CT + c] = NVB_FMUL(acc[b * CT + c], s_db[r.y0 + ((qq ^ sg) - sg)]);
                    }
                } else {
                    #pragma unroll
                    for (int b = 0; b < RB; b++) {
                        const int lo = posts32 ? 31 - __clz((int)((unsigned)M & (unsigned)below[b])) : 63 - __clzll((long long)(M & below[b]));
                        const RunSeg r = s_seg[c][lo];
                        const int num = (bin0 + b - r.x0) * (r.dy < 0 ? -r.dy : r.dy);
                        const int qq = r.m != 0u ? (int)__umulhi((unsigned)num, r.m) : num / s_adx[c][lo];
                        int y = r.dy < 0 ? r.y0 - qq : r.y0 + qq;
                        if ((unsigned)y > 255u) { bad_floor = 1; y = y < 0 ? 0 : 255; }
                        acc[b * CT + c] = NVB_FMUL(acc[b * CT + c], s_db[y]);
                    }
                }
            }
            float* dst = spec_out + (size_t)c * n + bin0;
            if (RB == 8) {
                *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
            } else if (RB == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[c], acc[CT + c], acc[2 * CT + c], acc[3 * CT + c]);
            else if (RB == 2) *reinterpret_cast<float2*>(dst) = make_float2(acc[c], acc[CT + c]);
            else *dst = acc[c];
        }
    }
    // rare: count the frame once per kind (the first thread to raise a flag reports it)
    if (bad_entry && atomicOr(&s_bad[0], 1) == 0) atomicAdd(&a.counters->bad_entry, 1);
    if (bad_floor && atomicOr(&s_bad[1], 1) == 0) atomicAdd(&a.counters->floor_range, 1);
}

// ------------------------------------------------------------------------------------------------
// K1+K2+K3, frame-group path (k_spectrum_wf; the default when DevSetup.spectrum_fast == 3): the work of k_spectrum_run with
// less than half of its instructions.  A frame belongs to a GROUP of WPF warps (1: the whole frame in one warp, nothing but
// __syncwarp between its phases -- the form a fused K1-K5 kernel wants; 2: one floor per warp in parallel, the better fit
// for small batches), four warps per CTA.
//   phase A  per channel: floor1_unwrap_mh (UnwrapPosts with RenderPoint's division as one multiply-high by a setup constant)
//            and one WfSeg per active post of Floor1.Apply's x-sorted walk; the multiply-high constant of a segment comes from
//            a table indexed by its length (no division).  Entry-stream offsets: the entries-per-partition of the stages of a
//            class are packed 16 bits each, so ONE 64-bit warp scan over the partitions yields the offsets of four stages.
//   main     a thread owns runs of 8 consecutive stream values (as in k_spectrum_run): class + coded-stage mask from shared
//            memory, ONE code path for every book with dims >= 2 (four float2 loads: lanes of a warp sit in different
//            partitions with different books, so per-dims paths serialise); the floor line is walked along the run -- segment
//            found once per (run, channel), y(x) = y0 +- umulhi((x - x0)|dy|, m) with the product advanced by |dy| per bin
//            and a switch to the next active post when the bin reaches the segment's end.
// Results are bit-identical to k_spectrum_run / the oracle (same float adds in the same order, same integer floor curve).
// ------------------------------------------------------------------------------------------------
// MINB: CTAs per SM the registers are capped for -- 8 (64 registers; the throughput form: 3 % faster on the 65 536-frame launch) or 7 (72
// registers, fewer spills; the form for launches of up to 16 384 frames: the kernel is then a single wave of dependent chains, and under
// programmatic dependent launch the CTAs of the next launch pile up on the SMs that drain first -- a cap of 7 spreads 1024 CTAs over 147
// SMs: 14.3 vs 18.4 us on 1024 frames, 15.1 vs 19.9 us on 2048, a 4096-frame spectrum + IMDCT step on one stream 41.7 vs 44.4 us).
template <int CT, int WPF, bool P64, int MINB = NVB_WF_MINB>
__global__ void __launch_bounds__(WF_WARPS * 32, MINB) k_spectrum_wf(LaunchArgs a, WfLayout L) {
    constexpr int FPC = WF_WARPS / WPF;                                     // frames per CTA
    constexpr int GT = WPF * 32;                                            // threads per frame group
    constexpr int H = P64 ? 2 : 1;
    typedef typename WfFrame<CT, P64>::mask_t mask_t;
    NVB_DYN_SMEM(dyn_smem);
    __shared__ float s_db[256];
    __shared__ int s_bad[2 * WF_WARPS];                                     // per frame group: bad entry, floor out of range

    nvb_grid_dep_launch();
    const DevSetup& S = a.S;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int group = warp / WPF, wg = warp - group * WPF, gt = t - group * GT;
    for (int i = t; i < 256; i += WF_WARPS * 32) s_db[i] = S.db[i];
    if (t < 2 * WF_WARPS) s_bad[t] = 0;
    __syncthreads();

    const int fi = blockIdx.x * FPC + group;
    // warps beyond the batch and drains have nothing to do (a whole group leaves together: its barrier has no other user)
    if (fi >= a.n_frames) { nvb_grid_dep_wait(); return; }
    const DevFrame f = a.frames[a.frame_lo + fi];
    if (f.kind != 0) { nvb_grid_dep_wait(); return; }
    const RunMode rm = S.run_modes[f.mode];
    const DevFloor1& F = S.floors[rm.floor];
    const int n = f.n >> 1;
    WfFrame<CT, P64> x;
    x.n = n; x.span = CT * n; x.np = L.np_pad; x.rbegin = rm.rbegin; x.pshift = rm.pshift; x.ST = rm.base_stride;
    x.st_n = rm.stages > 0 ? rm.stages : 1; x.n_coupling = rm.n_coupling; x.mapping = rm.mapping;
    x.P = 0;
    if (f.res_decoded) { const int e = rm.rend < x.span ? rm.rend : x.span; const int nn = e - rm.rbegin; x.P = nn > 0 ? nn >> rm.pshift : 0; }   // Residue0.cs:122-127
    x.entries_off = f.entries_off; x.ci_off = (uint32_t)rm.ci_off; x.bin2k_off = (uint32_t)(rm.floor * (S.bs[1] >> 1));
    x.ecount = f.entry_count; x.exec_mask = f.exec_mask; x.spec_off = f.spec_off;
    const uint8_t* cls = a.classes + f.classes_off;
    if (x.P > 0) for (uint32_t i = (uint32_t)gt * 64u; i < f.entry_count; i += GT * 64u) prefetch_l1(a.entries + f.entries_off + i);   // the frame's entries: 128 bytes per thread

    unsigned char* gsm = dyn_smem + L.cta_bytes + (size_t)group * L.total;
    WfSeg* s_seg = reinterpret_cast<WfSeg*>(gsm + L.seg_off);               // [CT][np_pad]
    int* s_fy = reinterpret_cast<int*>(gsm + L.fy_off) + wg * 256;          // per warp, for two channels at once: finalY[64], finalY * multiplier in x order [64]
    int* s_flags = reinterpret_cast<int*>(gsm + L.fy_off) + WPF * 256;      // [CT][4]: mask lo, mask hi, careful (groups of more than one warp)
    uint32_t* s_base = reinterpret_cast<uint32_t*>(gsm + L.base_off);       // [partition][ST]: where the (partition, stage) item's entries start
    uint16_t* s_cls = reinterpret_cast<uint16_t*>(gsm + L.cls_off);         // [partition]: class | coded stages << 8
    x.sdb = wf_smem(s_db); x.sseg = wf_smem(s_seg); x.sbase = wf_smem(s_base); x.scls = wf_smem(s_cls); x.sci = 0;

    // ---- phase A: floors (warp wg takes channels wg, wg + WPF, ...) and entry offsets (the group's last warp)
    #pragma unroll
    for (int c = 0; c < CT; c++) { x.fmask[c] = 0; x.careful[c] = false; }
    // channels of this warp in pairs (they interleave in floor1_wf_segments: twice the instruction-level parallelism)
    #pragma unroll
    for (int c = 0; c < CT; c++) {
        if ((c % WPF) != wg || !((f.exec_mask >> c) & 1u)) continue;
        const int c2 = c + WPF;                                             // the warp's next channel
        const bool pair = c2 < CT && ((f.exec_mask >> c2) & 1u);
        const int16_t* pp[2] = {a.posts + ((size_t)f.api_index * CT + c) * S.post_stride, a.posts + ((size_t)f.api_index * CT + (pair ? c2 : c)) * S.post_stride};
        int* fyp[2] = {s_fy, s_fy + 128}; int* ysp[2] = {s_fy + 64, s_fy + 192};
        WfSeg* sg[2] = {s_seg + c * L.np_pad, s_seg + (pair ? c2 : c) * L.np_pad};
        unsigned long long mask[2] = {0ull, 0ull}; int careful_any[2] = {0, 0};
        if (pair) floor1_wf_segments<H, 2>(F, S.magic, pp, n, lane, fyp, ysp, sg, mask, careful_any);
        else floor1_wf_segments<H, 1>(F, S.magic, pp, n, lane, fyp, ysp, sg, mask, careful_any);
        #pragma unroll
        for (int j = 0; j < 2; j++) {
            const int cj = j == 0 ? c : c2;
            if (j == 1 && !pair) break;
            #pragma unroll
            for (int cc = 0; cc < CT; cc++) if (cc == cj) { x.fmask[cc] = (mask_t)mask[j]; x.careful[cc] = careful_any[j] != 0; }
            if (WPF > 1 && lane == 0) { s_flags[4 * cj] = (int)(unsigned)mask[j]; s_flags[4 * cj + 1] = (int)(unsigned)(mask[j] >> 32); s_flags[4 * cj + 2] = careful_any[j]; }
        }
        __syncwarp();                                                       // fy / ys are reused by the warp's next channels
        if (pair) c = c2;                                                   // (the loop's own increment moves on from there)
    }
    if (wg == WPF - 1 && x.P > 0) wf_entry_offsets(S, rm, S.residues[rm.residue].coded, cls, x.P, lane, s_base, s_cls);
    wf_group_sync<WPF>(group);
    if (WPF > 1) {
        #pragma unroll
        for (int c = 0; c < CT; c++) {
            if ((f.exec_mask >> c) & 1u) {
                const unsigned lo = (unsigned)s_flags[4 * c], hi = (unsigned)s_flags[4 * c + 1];
                x.fmask[c] = (mask_t)(((unsigned long long)hi << 32) | lo); x.careful[c] = s_flags[4 * c + 2] != 0;
            }
        }
    }
    // the common frame: every channel executed, with a curve, all segments on the multiply-high path
    bool plain = (f.exec_mask & ((1u << CT) - 1u)) == ((1u << CT) - 1u);
    #pragma unroll
    for (int c = 0; c < CT; c++) plain = plain && x.fmask[c] != 0 && !x.careful[c];
    int cm = 0;
    const DevMapping& mp = S.mappings[rm.mapping];
    for (int i = 0; i < rm.n_coupling; i++) if (((f.exec_mask >> mp.mag[i]) | (f.exec_mask >> mp.ang[i])) & 1u) cm = 3;
    if (CT == 2 && cm == 3 && rm.n_coupling == 1) cm = mp.mag[0] == 0 ? 1 : 2;
    nvb_grid_dep_wait();                                                    // the stores below may overwrite a spectrum an earlier kernel still reads

    int bad_floor = 0, bad_entry = 0;
    if (plain) {
        if (cm == 1) wf_main<CT, GT, P64, true, CT == 2 ? 1 : 3>(a, x, gt, bad_entry, bad_floor);
        else if (cm == 0) wf_main<CT, GT, P64, true, 0>(a, x, gt, bad_entry, bad_floor);
        else if (cm == 2) wf_main<CT, GT, P64, true, CT == 2 ? 2 : 3>(a, x, gt, bad_entry, bad_floor);
        else wf_main<CT, GT, P64, true, 3>(a, x, gt, bad_entry, bad_floor);
    } else wf_main<CT, GT, P64, false, 3>(a, x, gt, bad_entry, bad_floor);
    // rare: count the frame once per kind (the first thread of the frame's group to raise a flag reports it)
    if (bad_entry && atomicOr(&s_bad[2 * group], 1) == 0) atomicAdd(&a.counters->bad_entry, 1);
    if (bad_floor && atomicOr(&s_bad[2 * group + 1], 1) == 0) atomicAdd(&a.counters->floor_range, 1);
}


// ------------------------------------------------------------------------------------------------
// K1+K2+K3, bins path (DevSetup.spectrum_bins; used when the run path does not apply): type 2 residues with ANY channel
// count and partition alignment -- e.g. 6 channels with 32-wide partitions, where (begin + p * psize) is not a multiple of
// the channel count and the reference's Residue2.WriteVectors (Residue2.cs:23-47) restarts the channel pointer at every
// partition and truncates the bin offset: element e of partition p lands on channel e % C of bin ob(p) + e / C, and two
// neighbouring partitions can reach the same (channel, bin).  A thread owns one BIN with all its channels in registers:
//   phase A  as in k_spectrum_run (floor segment records per channel warp, entry-stream offsets per (stage, partition));
//   main     per bin: the partitions that reach it come from a setup table; for every stage, in partition order (the
//            order of the reference's adds), the C values of the bin are C consecutive elements of the partition;
//            inverse coupling over the registers, floor multiply, one coalesced store per channel row.
// ------------------------------------------------------------------------------------------------
// (Round 2 tried the opposite mapping for BASELINE configs[3] -- residue accumulated partition by partition into per-channel planes in
// shared memory, warp-uniform class / book per partition, stages separated by block barriers and the elements a partition shares with
// its predecessor's last bin added in a second pass so that the reference's order of adds per cell is kept; then a thread per bin for
// coupling + floor.  Bit-identical, and slower: 0.66-0.73 ms against 0.51 ms per 8192 six-channel frames, 466 M against 445 M warp
// instructions -- the per-element index arithmetic and the read-modify-write of the planes cost what the divergence of this gather costs.)
template <int CT, int NT>
__global__ void __launch_bounds__(NT) k_spectrum_bins(LaunchArgs a) {
    constexpr int NW = NT / 32;
    NVB_DYN_SMEM(dyn_smem);
    __shared__ float s_db[256];
    __shared__ int s_fy[CT][NVB_MAX_POSTS];
    __shared__ int s_ys[CT][NVB_MAX_POSTS];
    __shared__ RunSeg s_seg[CT][NVB_MAX_POSTS];
    __shared__ int s_adx[CT][NVB_MAX_POSTS];
    __shared__ unsigned long long s_mask[CT];
    __shared__ int s_bad[2];
    __shared__ int s_careful[CT];

    nvb_grid_dep_launch();                                                  // see k_spectrum_run
    const DevFrame f = a.frames[a.frame_lo + blockIdx.x];
    if (f.kind != 0) { nvb_grid_dep_wait(); return; }
    const DevSetup& S = a.S;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const RunMode rm = S.run_modes[f.mode];
    const DevMapping& mp = S.mappings[rm.mapping];
    const DevFloor1& F = S.floors[rm.floor];
    const int N = f.n, n = N >> 1, span = CT * n;
    const int stages = rm.stages, st_n = stages > 0 ? stages : 1;
    const int rbegin = rm.rbegin, pshift = rm.pshift, nclass = rm.nclass, psize = 1 << pshift;
    int P = 0;
    if (f.res_decoded) { const int e = rm.rend < span ? rm.rend : span; const int nn = e - rbegin; P = nn > 0 ? nn >> pshift : 0; }   // Residue2.cs:16-21
    uint32_t* s_base = reinterpret_cast<uint32_t*>(dyn_smem);               // [stage][partition]: where the item's entries start
    const CiRec* ci_tab = S.ci + rm.ci_off;
    const uint8_t* cls = a.classes + f.classes_off;
    const uint16_t* ent = a.entries + f.entries_off;
    const uint8_t* bin2k = S.bin2k + (size_t)rm.floor * (S.bs[1] >> 1);
    const uint32_t* cand = S.r2cand + rm.cand_off;
    const uint16_t* obs = S.r2ob + rm.ob_off;

    for (int i = t; i < 256; i += NT) s_db[i] = S.db[i];
    if (t < 2) s_bad[t] = 0;
    if (P > 0) for (uint32_t i = (uint32_t)t * 64u; i < f.entry_count; i += NT * 64u) prefetch_l1(ent + i);

    // ---- phase A
    if (warp == NW - 1 && P > 0) {
        uint32_t run = 0;
        for (int st = 0; st < stages; st++) {
            for (int base = 0; base < P; base += 32) {
                const int p = base + lane;
                uint32_t c = 0;
                if (p < P) { const int cl = cls[p]; if (cl < nclass) c = (uint32_t)ci_tab[cl * st_n + st].cnt; }
                uint32_t incl = c;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
                if (p < P) s_base[st * P + p] = run + incl - c;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
    }
    for (int c = warp; c < CT; c += NW) {
        unsigned long long mask = 0ull; int careful_any = 0;
        if ((f.exec_mask >> c) & 1u)
            floor1_run_segments_warp(F, a.posts + ((size_t)f.api_index * CT + c) * S.post_stride, n, lane, s_fy[c], s_ys[c], s_seg[c], s_adx[c], mask, careful_any);
        if (lane == 0) { s_mask[c] = mask; s_careful[c] = careful_any; }
    }
    __syncthreads();
    nvb_grid_dep_wait();

    // ---- main: one bin per thread
    float* spec_out = a.spectrum + (size_t)f.spec_off;
    const uint32_t ecount = f.entry_count;
    const bool posts32 = F.n_posts <= 32;
    int bad_floor = 0, bad_entry = 0;
    for (int j = t; j < n; j += NT) {
        float acc[CT];
        #pragma unroll
        for (int c = 0; c < CT; c++) acc[c] = 0.f;
        const uint32_t cd = P > 0 ? cand[j] : 0u;
        const int p0 = (int)(cd & 0xffffu), ncand = (int)(cd >> 16);
        for (int st = 0; st < stages; st++) {
            for (int k = 0; k < ncand; k++) {                               // partition order = the order of the reference's adds
                const int p = p0 + k;
                if (p >= P) break;
                const int cl = cls[p];
                if (cl >= nclass) continue;
                const CiRec ci = ci_tab[cl * st_n + st];
                if (ci.cnt == 0) continue;
                const int e0 = (j - (int)obs[p]) * CT;                      // first element of this bin inside the partition
                const uint32_t eb = s_base[st * P + p];
                const float* tab = S.vq + ci.off;
                const int dmask = (1 << ci.dshift) - 1;
                #pragma unroll
                for (int c = 0; c < CT; c++) {
                    const int e = e0 + c;
                    if (e < psize) {
                        const uint32_t ei = eb + (uint32_t)(e >> ci.dshift);
                        if (ei < ecount) {                                  // else never decoded: contributes nothing (Residue0.cs:164-170)
                            const int en = ent[ei];
                            if (en < ci.entries) acc[c] = NVB_FADD(acc[c], tab[((size_t)en << ci.dshift) + (e & dmask)]); else bad_entry = 1;
                        }
                    }
                }
            }
        }
        for (int i = rm.n_coupling - 1; i >= 0; --i) {                      // Mapping.cs:137-182
            const int m = mp.mag[i], an = mp.ang[i];
            if (!(((f.exec_mask >> m) | (f.exec_mask >> an)) & 1u)) continue;
            float vm = 0.f, va = 0.f;
            #pragma unroll
            for (int k = 0; k < CT; k++) { if (k == m) vm = acc[k]; if (k == an) va = acc[k]; }
            inverse_couple_sel(vm, va);
            #pragma unroll
            for (int k = 0; k < CT; k++) { if (k == m) acc[k] = vm; if (k == an) acc[k] = va; }
        }
        const unsigned kk = bin2k[j];
        const unsigned long long below = posts32 ? (unsigned long long)(0xffffffffu >> (31 - kk)) : (0xffffffffffffffffull >> (63 - kk));
        #pragma unroll
        for (int c = 0; c < CT; c++) {
            if ((f.exec_mask >> c) & 1u) {                                  // Floor1.Apply, Floor1.cs:186-222
                const unsigned long long M = s_mask[c];
                if (M == 0ull) acc[c] = 0.f;
                else {
                    const int lo = posts32 ? 31 - __clz((int)((unsigned)M & (unsigned)below)) : 63 - __clzll((long long)(M & below));
                    const RunSeg r = s_seg[c][lo];
                    const int num = (j - r.x0) * (r.dy < 0 ? -r.dy : r.dy);
                    const int qq = r.m != 0u ? (int)__umulhi((unsigned)num, r.m) : num / s_adx[c][lo];
                    int y = r.dy < 0 ? r.y0 - qq : r.y0 + qq;
                    if ((unsigned)y > 255u) { bad_floor = 1; y = y < 0 ? 0 : 255; }
                    acc[c] = NVB_FMUL(acc[c], s_db[y]);
                }
            }
            spec_out[(size_t)c * n + j] = acc[c];
        }
    }
    if (bad_entry && atomicOr(&s_bad[0], 1) == 0) atomicAdd(&a.counters->bad_entry, 1);
    if (bad_floor && atomicOr(&s_bad[1], 1) == 0) atomicAdd(&a.counters->floor_range, 1);
}

// ------------------------------------------------------------------------------------------------
// K1+K2+K3, warp path (NVB_SPECTRUM_WARP=1; see launch_spectrum for why it is not the default): the same arithmetic as k_spectrum_planes with
// ONE WARP PER FRAME in a persistent grid -- no block-wide barrier, every warp runs its own frame start to end
// while the other warps of the SM hide its latencies:
//   floor unwrap + segment list per channel (lane = post), (stage, partition) item compaction (warp scan),
//   floor curve as 8-bit table indices (lanes over x), residue: half a warp per item, lane = VQ entry, whole VQ
//   vectors added into ONE accumulator plane stage after stage (the reference's add order; a (stage, position)
//   pair has a single writer, so no atomics), then groups of G = max(4, C) interleaved values: inverse coupling,
//   floor multiply, vector stores of the channel rows.
// Per-warp shared memory: accumulator plane C*n floats + C*n floor bytes + segment / item lists.
// ------------------------------------------------------------------------------------------------
struct WarpSmem { size_t acc, fl8, seg, fy, items, cls, misc, total; };
__host__ __device__ inline WarpSmem spectrum_warp_layout(int C, int bs1, int max_items) {
    WarpSmem w; size_t o = 0;
    const size_t span = (size_t)C * (bs1 / 2);
    w.acc = o; o += span * sizeof(float);
    w.fl8 = o; o += (span + 15) & ~size_t(15);
    w.seg = o; o += (((size_t)C * (NVB_MAX_POSTS + 1) * sizeof(SegRec)) + 15) & ~size_t(15);
    w.fy = o; o += (size_t)C * NVB_MAX_POSTS * sizeof(int);
    w.items = o; o += ((size_t)max_items * sizeof(ItemRec) + 15) & ~size_t(15);
    w.cls = o; o += ((size_t)max_items + 15) & ~size_t(15);
    w.misc = o; o += 64;
    w.total = o;
    return w;
}

template <int CT>
__global__ void __launch_bounds__(512) k_spectrum_warp(LaunchArgs a) {
    constexpr int G = CT > 4 ? CT : 4;
    constexpr int CSH = CT == 1 ? 0 : CT == 2 ? 1 : CT == 4 ? 2 : 3;
    NVB_DYN_SMEM(dyn_smem);
    const DevSetup& S = a.S;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, nwarps = blockDim.x >> 5;
    // CTA-shared: dB table + the (class, stage) records of every residue
    float* s_db = reinterpret_cast<float*>(dyn_smem);
    int4* s_ci = reinterpret_cast<int4*>(s_db + 256);
    const WarpSmem L = spectrum_warp_layout(CT, S.bs[1], S.max_items);
    unsigned char* wbase = dyn_smem + 1024 + (((size_t)S.ci_total * sizeof(int4) + 15) & ~size_t(15)) + (size_t)warp * L.total;
    float* s_acc = reinterpret_cast<float*>(wbase + L.acc);
    uint8_t* s_fl8 = wbase + L.fl8;
    SegRec* s_seg = reinterpret_cast<SegRec*>(wbase + L.seg);
    int* s_fy = reinterpret_cast<int*>(wbase + L.fy);
    ItemRec* s_items = reinterpret_cast<ItemRec*>(wbase + L.items);
    uint8_t* s_cls = wbase + L.cls;
    int* s_nseg = reinterpret_cast<int*>(wbase + L.misc);                   // [CT]

    nvb_grid_dep_launch();
    for (int i = t; i < 256; i += blockDim.x) s_db[i] = S.db[i];
    for (int r = 0; r < S.n_residues; r++) {
        const DevResidue& R = S.residues[r];
        const int st_n = R.stages > 0 ? R.stages : 1;
        for (int i = t; i < R.nclass * st_n; i += blockDim.x) {
            const int cl = i / st_n, st = i - cl * st_n;
            const int book = st < R.stages ? R.books[cl][st] : -1;
            int4 ci = make_int4(0, 1, 0, 0);
            if (book >= 0) { const DevBook b = S.books[book]; ci = make_int4((int)b.off, b.dims, b.entries, R.cnt[cl][st]); }
            s_ci[R.ci_off + i] = ci;
        }
    }
    __syncthreads();
    nvb_grid_dep_wait();

    const uint32_t lt = (1u << lane) - 1u;
    const int hl = lane & 15, half = lane >> 4;
    for (int fi = blockIdx.x * nwarps + warp; fi < a.n_frames; fi += gridDim.x * nwarps) {
        const DevFrame f = a.frames[a.frame_lo + fi];
        if (f.kind != 0) continue;
        const DevMode md = S.modes[f.mode];
        const DevMapping& mp = S.mappings[md.mapping];
        const DevResidue& R = S.residues[mp.residue];
        const DevFloor1& F = S.floors[mp.floor];
        const int n = f.n >> 1, span = CT * n;
        const int st_n = R.stages > 0 ? R.stages : 1;
        const uint8_t* cls = a.classes + f.classes_off;
        const uint16_t* ent = a.entries + f.entries_off;
        ResGeom g; g.P = 0; g.Sx = 1; g.n_items = 0;
        if (f.res_decoded) g = residue_geom(R, f.n, CT);
        const int P = g.P;
        int bad_floor = 0, bad_entry = 0;
        __syncwarp();                                                       // the previous frame's readers of this warp's buffers are done

        // ---- floors: unwrap + segments, one channel after the other
        for (int c = 0; c < CT; c++) {
            if ((f.exec_mask >> c) & 1u)
                floor1_segments_warp(F, a.posts + ((size_t)f.api_index * CT + c) * S.post_stride, n, lane, s_fy + c * NVB_MAX_POSTS, s_seg + c * (NVB_MAX_POSTS + 1), &s_nseg[c]);
            else if (lane == 0) s_nseg[c] = 0;
        }
        // ---- residue items: compaction of the coded (stage, partition) pairs, start of each one's entries
        int nitems = 0; int stage_end[NVB_MAX_STAGES];
        {
            uint32_t run = 0;
            for (int st = 0; st < R.stages; st++) {
                for (int base = 0; base < P; base += 32) {
                    const int p = base + lane;
                    uint32_t c = 0; int cl = 255;
                    if (p < P) { cl = cls[p]; if (cl < R.nclass) c = (uint32_t)R.cnt[cl][st]; else cl = 255; if (st == 0) s_cls[p] = (uint8_t)cl; }
                    uint32_t incl = c;
                    #pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
                    const unsigned m = __ballot_sync(0xffffffffu, c > 0);
                    if (c > 0) { ItemRec r; r.p = (uint16_t)p; r.s = (uint8_t)st; r.cl = (uint8_t)cl; r.base = run + incl - c; s_items[nitems + __popc(m & lt)] = r; }
                    nitems += __popc(m);
                    run += __shfl_sync(0xffffffffu, incl, 31);
                }
                stage_end[st] = nitems;
            }
        }
        // ---- clear the accumulator plane (Mapping.cs:108)
        for (int i = lane * 4; i < span; i += 128) *reinterpret_cast<float4*>(s_acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();

        // ---- floor curve as table indices, channel-interleaved (RenderLineMulti in closed form, Floor1.cs:316-341)
        for (int c = 0; c < CT; c++) {
            const int ns = s_nseg[c];
            for (int u = 0; u < ns; u++) {
                const SegRec r = s_seg[c * (NVB_MAX_POSTS + 1) + u];
                const int len = r.x1 - r.x0;
                const int dyabs = r.ady + (r.b < 0 ? -r.b : r.b) * len;     // |dy|: y(k) = y0 + sy * floor(k |dy| / adx)
                uint8_t* row = s_fl8 + r.x0 * CT + c;
                for (int k = lane; k < len; k += 32) {
                    int y = r.y0 + r.sy * div_small(k * dyabs, len, r.rcp);
                    if ((unsigned)y > 255u) { bad_floor = 1; y = y < 0 ? 0 : 255; }
                    row[k * CT] = (uint8_t)y;
                }
            }
        }
        // ---- residue: stage after stage, half a warp per item, lane = entry, whole VQ vectors added into the plane
        {
            int it0 = 0;
            for (int st = 0; st < R.stages; st++) {
                const int it1 = stage_end[st];
                for (int idx = it0 + half; idx < it1; idx += 2) {
                    const ItemRec r = s_items[idx];
                    const int4 ci = s_ci[R.ci_off + r.cl * st_n + st];
                    const int dims = ci.y;
                    float* dst = s_acc + R.begin + (int)r.p * R.psize;
                    const float* tab = S.vq + ci.x;
                    for (int e = hl; e < ci.w; e += 16) {
                        const uint32_t ei = r.base + (uint32_t)e;
                        if (ei >= f.entry_count) continue;                  // never decoded: contributes nothing (Residue0.cs:164-170)
                        const int en = ent[ei];
                        if (en >= ci.z) { bad_entry = 1; continue; }
                        const float* src = tab + (size_t)en * dims;
                        float* d = dst + e * dims;
                        if (dims == 2) {
                            const float2 v = *reinterpret_cast<const float2*>(src); float2 o = *reinterpret_cast<float2*>(d);
                            o.x = NVB_FADD(o.x, v.x); o.y = NVB_FADD(o.y, v.y); *reinterpret_cast<float2*>(d) = o;
                        } else if (dims == 1) *d = NVB_FADD(*d, *src);
                        else for (int k = 0; k < dims; k += 4) {
                            const float4 v = *reinterpret_cast<const float4*>(src + k); float4 o = *reinterpret_cast<float4*>(d + k);
                            o.x = NVB_FADD(o.x, v.x); o.y = NVB_FADD(o.y, v.y); o.z = NVB_FADD(o.z, v.z); o.w = NVB_FADD(o.w, v.w);
                            *reinterpret_cast<float4*>(d + k) = o;
                        }
                    }
                }
                it0 = it1;
                __syncwarp();
            }
        }
        __syncwarp();

        // ---- inverse coupling, floor multiply, store
        float* spec_out = a.spectrum + (size_t)f.spec_off;
        bool execc[CT], hasfl[CT];
        #pragma unroll
        for (int c = 0; c < CT; c++) { execc[c] = (f.exec_mask >> c) & 1u; hasfl[c] = s_nseg[c] > 0; }
        for (int pos = lane * G; pos < span; pos += 32 * G) {
            float acc[G];
            #pragma unroll
            for (int k = 0; k < G; k += 4) {
                const float4 v = *reinterpret_cast<const float4*>(s_acc + pos + k);
                acc[k] = v.x; acc[k + 1] = v.y; acc[k + 2] = v.z; acc[k + 3] = v.w;
            }
            for (int i = mp.n_coupling - 1; i >= 0; --i) {                  // Mapping.cs:137-182
                const int m = mp.mag[i], an = mp.ang[i];
                if (!(((f.exec_mask >> m) | (f.exec_mask >> an)) & 1u)) continue;
                #pragma unroll
                for (int b = 0; b < G / CT; b++) {
                    float vm = 0.f, va = 0.f;
                    #pragma unroll
                    for (int k = 0; k < CT; k++) { if (k == m) vm = acc[b * CT + k]; if (k == an) va = acc[b * CT + k]; }
                    inverse_couple(vm, va);
                    #pragma unroll
                    for (int k = 0; k < CT; k++) { if (k == m) acc[b * CT + k] = vm; if (k == an) acc[b * CT + k] = va; }
                }
            }
            #pragma unroll
            for (int k = 0; k < G; k += 4) {
                const uchar4 y4 = *reinterpret_cast<const uchar4*>(s_fl8 + pos + k);
                const uint8_t yv[4] = {y4.x, y4.y, y4.z, y4.w};
                #pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int c = (k + e) & (CT - 1);
                    if (execc[c]) acc[k + e] = hasfl[c] ? NVB_FMUL(acc[k + e], s_db[yv[e]]) : 0.f;      // Floor1.Apply, Floor1.cs:186-222
                }
            }
            const int bin0 = pos >> CSH;
            if (CT == 1) *reinterpret_cast<float4*>(spec_out + bin0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            else if (CT == 2) {
                *reinterpret_cast<float2*>(spec_out + bin0) = make_float2(acc[0], acc[2]);
                *reinterpret_cast<float2*>(spec_out + n + bin0) = make_float2(acc[1], acc[3]);
            } else {
                #pragma unroll
                for (int c = 0; c < CT; c++) spec_out[(size_t)c * n + bin0] = acc[c];
            }
        }
        if (__any_sync(0xffffffffu, bad_entry) && lane == 0) atomicAdd(&a.counters->bad_entry, 1);
        if (__any_sync(0xffffffffu, bad_floor) && lane == 0) atomicAdd(&a.counters->floor_range, 1);
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
    nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    const DevFrame f = a.frames[a.frame_lo + fi];
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
    nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    const DevFrame f = a.frames[a.frame_lo + blockIdx.x];
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
// NVB_RUN_PCM_S16: float PCM -> 16-bit PCM, s = rni(v * 32768) saturated (cvt.rni.sat.s16.f32), elements [lo, hi) of the
// interleaved buffers.  Memory-bound (6 bytes per sample); halves what crosses PCIe afterwards.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ short pcm_s16(float v) {
#if !defined(NVB_CPU_SHIM)
    short r; asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(r) : "f"(v * 32768.0f)); return r;
#else
    const float x = std::nearbyintf(v * 32768.0f);                          // round to nearest even (default rounding mode)
    return (short)(x > 32767.f ? 32767.f : x < -32768.f ? -32768.f : x);
#endif
}
__global__ void __launch_bounds__(256) k_pcm_s16(const float* __restrict__ src, short* __restrict__ dst, long long lo, long long hi) {
    nvb_grid_dep_launch();
    nvb_grid_dep_wait();
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    // groups of four elements whose float4 / 8-byte short4 are both aligned (the buffers are 16- / 8-byte aligned); ragged ends one by one
    const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
    long long a0 = vec ? ((lo + 3) & ~3ll) : hi, a1 = vec ? (hi & ~3ll) : hi;
    if (a0 > a1) { a0 = hi; a1 = hi; }
    for (long long i = lo + tid; i < (a0 < hi ? a0 : hi); i += nt) dst[i] = pcm_s16(src[i]);
    for (long long g = (a0 >> 2) + tid; g < (a1 >> 2); g += nt) {
        const float4 v = reinterpret_cast<const float4*>(src)[g];
        short4 o; o.x = pcm_s16(v.x); o.y = pcm_s16(v.y); o.z = pcm_s16(v.z); o.w = pcm_s16(v.w);
        reinterpret_cast<short4*>(dst)[g] = o;
    }
    for (long long i = (a1 > a0 ? a1 : hi) + tid; i < hi; i += nt) dst[i] = pcm_s16(src[i]);
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------

// Launch-configuration caches are per device: cudaFuncSetAttribute applies to the current device's context only, and one
// process may hold contexts on several GPUs.
constexpr int NVB_MAX_DEVICES = 64;
static inline int current_device_slot() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < NVB_MAX_DEVICES) ? d : 0; }
static inline int device_sm_count() {
    static std::atomic<int> cache[NVB_MAX_DEVICES];
    const int slot = current_device_slot();
    int v = cache[slot].load(std::memory_order_relaxed);
    if (v == 0) {
        int dev = 0; cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[slot].store(v, std::memory_order_relaxed);
    }
    return v;
}

static size_t spectrum_smem(const DevSetup& S) {
    size_t b = (size_t)(((S.max_items > 0 ? S.max_items : 1) + 3) & ~3) * sizeof(uint32_t);
    if (S.f0_stride > 0) b += (size_t)S.channels * ((S.bs[1] >> 1) + 256) * sizeof(float);     // type 0 floors: curve + coefficient tables
    return b;
}

int launch_spectrum(const LaunchArgs& a, void* stream) {
    if (a.n_frames <= 0) return 0;
    NvbNoEarlyStart serialise(a.inputs_from_kernel != 0);
    static const bool force_generic = std::getenv("NVB_SPECTRUM_GENERIC") != nullptr;      // test hook: exercise the general kernel
    static const bool no_bins = std::getenv("NVB_SPECTRUM_NO_BINS") != nullptr;                // test hook
    if (!a.S.spectrum_fast && a.S.spectrum_bins && !force_generic && !no_bins) {
        const int C = a.S.channels;
        const size_t smem = (size_t)a.S.max_items * sizeof(uint32_t) + 16;
        auto go = [&](auto kernel) -> int {
            // the run / bins kernels carry up to ~15 KB of static shared memory: static + dynamic above 48 KB needs the opt-in
            if (smem > 30 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
            NVB_LAUNCH(kernel, a.n_frames, 128, smem, stream, a);
            return cudaGetLastError() == cudaSuccess ? 1 : -1;
        };
        switch (C) {
            case 1: return go(k_spectrum_bins<1, 128>); case 2: return go(k_spectrum_bins<2, 128>); case 3: return go(k_spectrum_bins<3, 128>);
            case 4: return go(k_spectrum_bins<4, 128>); case 5: return go(k_spectrum_bins<5, 128>); case 6: return go(k_spectrum_bins<6, 128>);
            case 7: return go(k_spectrum_bins<7, 128>); default: return go(k_spectrum_bins<8, 128>);
        }
    }
    if (!a.S.spectrum_fast || force_generic) return launch_spectrum_generic(a, stream);
    static const bool no_planes = std::getenv("NVB_SPECTRUM_NO_PLANES") != nullptr;           // test hook: exercise k_spectrum_fast
    // k_spectrum_warp (one warp per frame, no block barrier) is kept as an option: on B200 it measured slower than
    // k_spectrum_planes (94.8 vs 80.7 us for 4096 stereo frames, profiles/r01_i): 14 warps per SM cannot hide its
    // dependent entry -> VQ-vector loads
    static const bool use_warp = std::getenv("NVB_SPECTRUM_WARP") != nullptr;
    if (a.S.spectrum_fast >= 2 && !no_planes && use_warp) {
        const int C = a.S.channels;
        const WarpSmem L = spectrum_warp_layout(C, a.S.bs[1], a.S.max_items);
        const size_t shared_part = 1024 + (((size_t)a.S.ci_total * sizeof(int4) + 15) & ~size_t(15));
        int nw = (int)((220 * 1024 - shared_part) / L.total);
        if (nw > 16) nw = 16;
        if (nw >= 2) {
            const int num_sms = device_sm_count();
            const size_t smem = shared_part + (size_t)nw * L.total;
            static std::atomic<size_t> configured_w_by_c[NVB_MAX_DEVICES][NVB_FAST_CHANNELS + 1];    // per device and template instantiation
            if (!nvb_ensure_smem(configured_w_by_c[current_device_slot()][C], smem, [&]() {
                    return (C == 1 ? cudaFuncSetAttribute(k_spectrum_warp<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : C == 2 ? cudaFuncSetAttribute(k_spectrum_warp<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : C == 4 ? cudaFuncSetAttribute(k_spectrum_warp<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                   : cudaFuncSetAttribute(k_spectrum_warp<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) == cudaSuccess; })) return -1;
            int grid = (a.n_frames + nw - 1) / nw;
            if (grid > num_sms) grid = num_sms;
            if (C == 1) NVB_LAUNCH(k_spectrum_warp<1>, grid, nw * 32, smem, stream, a);
            else if (C == 2) NVB_LAUNCH(k_spectrum_warp<2>, grid, nw * 32, smem, stream, a);
            else if (C == 4) NVB_LAUNCH(k_spectrum_warp<4>, grid, nw * 32, smem, stream, a);
            else NVB_LAUNCH(k_spectrum_warp<8>, grid, nw * 32, smem, stream, a);
            return cudaGetLastError() == cudaSuccess ? 1 : -1;
        }
    }
    static const bool force_planes = std::getenv("NVB_SPECTRUM_PLANES") != nullptr;           // test hook: exercise k_spectrum_planes
    static const bool force_run = std::getenv("NVB_SPECTRUM_RUN") != nullptr;                 // test hook: exercise k_spectrum_run
    if (a.S.spectrum_fast >= 3 && !no_planes && !force_planes && !force_run) {
        const int C = a.S.channels;
        const int wpf_env = std::getenv("NVB_WF_WPF") ? std::atoi(std::getenv("NVB_WF_WPF")) : 0;   // experiment hook: warps per frame
        const int WPF = (wpf_env == 1 || wpf_env == 2 || wpf_env == 4) ? wpf_env : 1;
        const WfLayout L = wf_layout(a.S, C, WPF);
        const int fpc = WF_WARPS / WPF;
        const size_t smem = (size_t)L.cta_bytes + (size_t)L.total * fpc;
        if (smem <= 200 * 1024 && L.cta_bytes <= 16 * 1024) {
            const bool p64 = L.np_pad > 32;
            static const bool no_small = std::getenv("NVB_WF_NO_SMALL") != nullptr;              // experiment hook: always the 8-CTA form
            const bool small = a.n_frames <= 16384 && !no_small;
            auto go = [&](auto kernel) -> int {
                static std::atomic<size_t> configured[NVB_MAX_DEVICES];       // one per instantiation (a lambda instantiation has its own statics)
                // static + dynamic shared memory above 48 KB needs the opt-in: ask for it whenever the dynamic part is not small
                if (smem > 32 * 1024 && !nvb_ensure_smem(configured[current_device_slot()], smem, [&]() {
                        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; })) return -1;
                NVB_LAUNCHV(kernel, (a.n_frames + fpc - 1) / fpc, WF_WARPS * 32, smem, stream, a, L);
                return cudaGetLastError() == cudaSuccess ? 1 : -1;
            };
#define NVB_WF_CASE(CT_)                                                                                          \
            if (WPF == 1 && small) return p64 ? go(k_spectrum_wf<CT_, 1, true, 7>) : go(k_spectrum_wf<CT_, 1, false, 7>); \
            if (WPF == 1) return p64 ? go(k_spectrum_wf<CT_, 1, true>) : go(k_spectrum_wf<CT_, 1, false>);          \
            if (WPF == 2) return p64 ? go(k_spectrum_wf<CT_, 2, true>) : go(k_spectrum_wf<CT_, 2, false>);          \
            return p64 ? go(k_spectrum_wf<CT_, 4, true>) : go(k_spectrum_wf<CT_, 4, false>);
            if (C == 1) { NVB_WF_CASE(1) }
            if (C == 2) { NVB_WF_CASE(2) }
            if (C == 4) { NVB_WF_CASE(4) }
            if (C == 8) { NVB_WF_CASE(8) }
#undef NVB_WF_CASE
        }
    }
    if (a.S.spectrum_fast >= 3 && !no_planes && !force_planes) {
        const int C = a.S.channels;
        static const int nt = std::getenv("NVB_SPECTRUM_NT") ? std::atoi(std::getenv("NVB_SPECTRUM_NT")) : 128;
        const size_t smem = (size_t)a.S.max_items * sizeof(uint32_t) + 16;
        auto go = [&](auto kernel, int threads) -> int {
            // the run / bins kernels carry up to ~15 KB of static shared memory: static + dynamic above 48 KB needs the opt-in
            if (smem > 30 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
            NVB_LAUNCH(kernel, a.n_frames, threads, smem, stream, a);
            return cudaGetLastError() == cudaSuccess ? 1 : -1;
        };
        if (nt == 96 && C <= 2) return C == 1 ? go(k_spectrum_run<1, 96>, 96) : go(k_spectrum_run<2, 96>, 96);
        if (nt == 64) return C == 1 ? go(k_spectrum_run<1, 64>, 64) : C == 2 ? go(k_spectrum_run<2, 64>, 64) : C == 4 ? go(k_spectrum_run<4, 64>, 64) : go(k_spectrum_run<8, 64>, 64);
        if (nt == 128) return C == 1 ? go(k_spectrum_run<1, 128>, 128) : C == 2 ? go(k_spectrum_run<2, 128>, 128) : C == 4 ? go(k_spectrum_run<4, 128>, 128) : go(k_spectrum_run<8, 128>, 128);
        return C == 1 ? go(k_spectrum_run<1, 256>, 256) : C == 2 ? go(k_spectrum_run<2, 256>, 256) : C == 4 ? go(k_spectrum_run<4, 256>, 256) : go(k_spectrum_run<8, 256>, 256);
    }
    if (a.S.spectrum_fast >= 2 && !no_planes) {
        const int C = a.S.channels;
        // planes for the deepest residue + floor rows + item list + class bytes (host-checked to fit)
        const size_t span = (size_t)C * (a.S.bs[1] / 2) * sizeof(float);
        const size_t smem = span * (size_t)(a.S.max_stages + 1) + (((size_t)a.S.max_items * sizeof(ItemRec) + 15) & ~size_t(15)) +
                            (((size_t)a.S.max_items + 15) & ~size_t(15)) + (size_t)a.S.ci_total * sizeof(int4) + 16;
        static std::atomic<size_t> configured_pl_by_c[NVB_MAX_DEVICES][NVB_FAST_CHANNELS + 1];   // per device and template instantiation
        if (!nvb_ensure_smem(configured_pl_by_c[current_device_slot()][C], smem, [&]() {
                return (C == 1 ? cudaFuncSetAttribute(k_spectrum_planes<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : C == 2 ? cudaFuncSetAttribute(k_spectrum_planes<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : C == 4 ? cudaFuncSetAttribute(k_spectrum_planes<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                               : cudaFuncSetAttribute(k_spectrum_planes<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) == cudaSuccess; })) return -1;
        if (C == 1) NVB_LAUNCH(k_spectrum_planes<1>, a.n_frames, SPEC_THREADS, smem, stream, a);
        else if (C == 2) NVB_LAUNCH(k_spectrum_planes<2>, a.n_frames, SPEC_THREADS, smem, stream, a);
        else if (C == 4) NVB_LAUNCH(k_spectrum_planes<4>, a.n_frames, SPEC_THREADS, smem, stream, a);
        else NVB_LAUNCH(k_spectrum_planes<8>, a.n_frames, SPEC_THREADS, smem, stream, a);
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    const size_t smem = spectrum_smem(a.S) + (size_t)a.S.channels * (a.S.bs[1] / 2) * sizeof(float);
    static std::atomic<size_t> configured_by_dev[NVB_MAX_DEVICES];
    if (smem > 24 * 1024 && !nvb_ensure_smem(configured_by_dev[current_device_slot()], smem, [&]() {
            return cudaFuncSetAttribute(k_spectrum_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; })) return -1;
    NVB_LAUNCH(k_spectrum_fast, a.n_frames, SPEC_THREADS, smem, stream, a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_spectrum_generic(const LaunchArgs& a, void* stream) {
    if (a.n_frames <= 0) return 0;
    NvbNoEarlyStart serialise(a.inputs_from_kernel != 0);
    static const bool forbid = std::getenv("NVB_SPECTRUM_FORBID_GENERIC") != nullptr;           // test hook: prove a faster kernel covers the setup
    if (forbid) return -1;
    size_t smem = spectrum_smem(a.S);
    static std::atomic<size_t> configured_by_dev[NVB_MAX_DEVICES];
    // (static shared memory: 5 KB for up to 8 channels, 18 KB for up to 32 -- ask for the opt-in early)
    if (smem > 48 * 1024 - 20 * 1024 && !nvb_ensure_smem(configured_by_dev[current_device_slot()], smem, [&]() {
            return cudaFuncSetAttribute(k_spectrum<NVB_FAST_CHANNELS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                   cudaFuncSetAttribute(k_spectrum<NVB_MAX_CHANNELS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; })) return -1;
    if (a.S.channels <= NVB_FAST_CHANNELS) NVB_LAUNCH(k_spectrum<NVB_FAST_CHANNELS>, a.n_frames, SPEC_THREADS, smem, stream, a);
    else NVB_LAUNCH(k_spectrum<NVB_MAX_CHANNELS>, a.n_frames, SPEC_THREADS, smem, stream, a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_imdct_exact(const LaunchArgs& a, void* stream) {
    if (a.n_frames <= 0) return 0;
    size_t smem = (size_t)(a.S.bs[1] + a.S.bs[1] / 2) * sizeof(float);
    static std::atomic<size_t> configured_by_dev[NVB_MAX_DEVICES];
    if (smem > 40 * 1024 && !nvb_ensure_smem(configured_by_dev[current_device_slot()], smem, [&]() {
            return cudaFuncSetAttribute(k_imdct_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; })) return -1;
    NVB_LAUNCH(k_imdct_exact, a.n_frames * a.S.channels, MDCT_THREADS, smem, stream, a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_pcm_s16(const float* src, int16_t* dst, long long lo, long long hi, void* stream) {
    if (hi <= lo) return 0;
    const long long groups = (hi - lo + 3) / 4;
    long long blocks = (groups + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    NVB_LAUNCHV(k_pcm_s16, blocks, 256, 0, stream, src, reinterpret_cast<short*>(dst), lo, hi);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_ola(const LaunchArgs& a, void* stream) {
    if (a.n_frames <= 0) return 0;
    NVB_LAUNCH(k_ola, a.n_frames, OLA_THREADS, 0, stream, a);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace nvb
