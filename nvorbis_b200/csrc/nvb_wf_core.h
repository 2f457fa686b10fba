// nvb_wf_core.h -- the warp-per-frame spectrum stage (K1+K2+K3) as device functions: floor 1 unwrap + segment records, entry-stream
// offsets, and the main loop (residue gather, inverse coupling, walked floor line).  Shared by k_spectrum_wf (nvb_kernels.cu: dense
// spectrum in HBM) and by the one-kernel synthesis path of nvb_fused.cu (k_imdct_fused_t<.., SYN>: the spectrum goes straight into the
// frame's shared-memory slot).  Restates Residue0.cs:119-201 / Residue2.cs:23-47 (WriteVectors), Mapping.cs:137-182 (inverse
// coupling), Floor1.cs:186-341 (Apply = UnwrapPosts + RenderLineMulti).
#pragma once
#include <type_traits>
#include "nvb_device_core.h"

namespace nvb {

#if !defined(NVB_CPU_SHIM)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#else
static inline void prefetch_l1(const void*) {}
#endif

#ifndef NVB_WF_WARPS
#define NVB_WF_WARPS 4                                                    // warps per CTA.  (2 / 8 warps per CTA and 7 instead of 8 CTAs per SM
#endif                                                                    //  -- 72 registers -- all measured 21.7-21.9 us on configs[1], round 2)
#ifndef NVB_WF_MINB
#define NVB_WF_MINB 8
#endif
constexpr int WF_WARPS = NVB_WF_WARPS;

// the group's barrier: __syncwarp for one warp per frame, a named barrier for two, __syncthreads for the whole CTA
template <int WPF> __device__ __forceinline__ void wf_group_sync(int group) {
    if (WPF == 1) __syncwarp();
    else if (WPF == WF_WARPS) __syncthreads();
    else {
#if !defined(NVB_CPU_SHIM)
        // ids as immediates: a register id makes ptxas reserve all 16 barriers for the CTA
        if (group == 0) asm volatile("bar.sync 1, %0;" ::"n"(WPF * 32) : "memory");
        else asm volatile("bar.sync 2, %0;" ::"n"(WPF * 32) : "memory");
#else
        cuemu_named_barrier(group + 1, WPF * 32);
#endif
    }
}

#if !defined(NVB_CPU_SHIM)
__device__ __forceinline__ unsigned wf_reduce_or(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
#else
static inline unsigned wf_reduce_or(unsigned v) { for (int d = 16; d >= 1; d >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, d); return v; }
#endif

// UnwrapPosts (Floor1.cs:224-297) by one warp for NC channels of one frame at once (they share the floor, so the per-post
// constants are loaded once and the channels' dependency chains interleave): lane = post (and post + 32 when H == 2), the posts
// of one dependency level in parallel; RenderPoint's `err / adx` (Floor1.cs:299-314) is a multiply-high by the setup constant
// F.magic[i] -- exact for err < 2^20 (adx <= 4096), anything larger (only malformed posts get there) takes the division.
// Leaves finalY of channel j in fy[j][]; flags[j] = step flags (0 when PostCount < 2: the spectrum is cleared, Floor1.cs:220).
// (Round 2 also tried one channel per HALF-warp -- lane = (channel, slot of the level), the level's posts from a per-level table, one
// pass of the body for both channels: 200 fewer warp instructions per stereo frame and no change in time, 21.8 us -- the phase is bound
// by the dependent chain level -> level (shared-memory round trip + barrier per level), not by instruction issue.)
template <int H, int NC>
__device__ __forceinline__ void floor1_unwrap_mh(const DevFloor1& F, const int16_t* const* posts, int lane, int* const* fy, int* count, unsigned long long* flags) {
    int p_lo[H], p_hi[H], p_x0[H], p_dx[H], p_lvl[H]; unsigned p_m[H];
    const int n_posts = F.n_posts;
    #pragma unroll
    for (int h = 0; h < H; h++) {
        const int i = lane + 32 * h;
        p_lvl[h] = 0; p_lo[h] = 0; p_hi[h] = 0; p_x0[h] = 0; p_dx[h] = 0; p_m[h] = 0u;
        if (i >= 2 && i < n_posts) {
            p_lo[h] = F.lo[i]; p_hi[h] = F.hi[i]; p_x0[h] = F.x[p_lo[h]]; p_dx[h] = F.x[i] - p_x0[h];
            p_m[h] = F.magic[i]; p_lvl[h] = F.level[i];
        }
    }
    int val[NC][H]; unsigned clo[NC], chi[NC];
    #pragma unroll
    for (int j = 0; j < NC; j++) {
        int c = posts[j][0];
        if (c > n_posts) c = n_posts;
        if (c < 2) c = 0;
        count[j] = c; clo[j] = 0u; chi[j] = 0u;
        #pragma unroll
        for (int h = 0; h < H; h++) { const int i = lane + 32 * h; val[j][h] = i < c ? posts[j][1 + i] : 0; }
        if (lane < 2) fy[j][lane] = val[j][0];
    }
    __syncwarp();
    const int range = F.range, max_level = F.max_level;
    for (int lvl = 1; lvl <= max_level; lvl++) {
        #pragma unroll
        for (int h = 0; h < H; h++) {
            if (p_lvl[h] == lvl) {
                const int i = lane + 32 * h;
                #pragma unroll
                for (int j = 0; j < NC; j++) {
                    if (i < count[j]) {
                        const int y0 = fy[j][p_lo[h]];
                        const int dy = fy[j][p_hi[h]] - y0, ady = dy < 0 ? -dy : dy;
                        const int err = ady * p_dx[h];
                        const int off = (unsigned)err < (1u << 20) ? (int)__umulhi((unsigned)err, p_m[h]) : err / ((int)F.x[p_hi[h]] - p_x0[h]);
                        const int predicted = dy < 0 ? y0 - off : y0 + off;
                        const int v = val[j][h];
                        const int highroom = range - predicted, lowroom = predicted;
                        const int room = (highroom < lowroom ? highroom : lowroom) * 2;
                        int out = predicted;
                        if (v != 0) {
                            if (H == 1) clo[j] |= (1u << p_lo[h]) | (1u << p_hi[h]) | (1u << i);
                            else {
                                const unsigned long long b = (1ull << p_lo[h]) | (1ull << p_hi[h]) | (1ull << i);
                                clo[j] |= (unsigned)b; chi[j] |= (unsigned)(b >> 32);
                            }
                            if (v >= room) out = highroom > lowroom ? v - lowroom + predicted : predicted - v + highroom - 1;
                            else out = (v & 1) ? predicted - ((v + 1) >> 1) : predicted + (v >> 1);       // v > 0 here: (v % 2) == 1 <=> v & 1
                        }
                        fy[j][i] = out;
                    }
                }
            }
        }
        __syncwarp();
    }
    // stepFlags: 0 and 1 always; i when its own value is non-zero or a later post names it as a neighbour (Floor1.cs:253-257,292)
    #pragma unroll
    for (int j = 0; j < NC; j++) {
        const unsigned lo = wf_reduce_or(clo[j]);
        const unsigned hi = H == 2 ? wf_reduce_or(chi[j]) : 0u;
        flags[j] = count[j] >= 2 ? ((((unsigned long long)hi << 32) | lo) | 3ull) : 0ull;
    }
}

// Floor 1 of NC channels of a frame by one warp: unwrap, active-post mask of the x-sorted walk (bit k: sorted position k starts a
// segment; 0 = no curve, Floor1.cs:220) and one WfSeg per active position.  careful: some segment needs the plain division or
// leaves inverse_dB_table's range.  fy[j] / ys[j]: 64 ints of scratch each; seg[j]: the channel's segment records.
template <int H, int NC>
__device__ __forceinline__ void floor1_wf_segments(const DevFloor1& F, const uint32_t* magic, const int16_t* const* posts, int n, int lane, int* const* fy, int* const* ys,
                                                   WfSeg* const* seg, unsigned long long* mask, int* careful_any) {
    int count[NC]; unsigned long long flags[NC];
    floor1_unwrap_mh<H, NC>(F, posts, lane, fy, count, flags);
    const int mult = F.mult;
    int xs_k[H], idx_k[H];
    #pragma unroll
    for (int h = 0; h < H; h++) { const int k = lane + 32 * h; xs_k[h] = 0; idx_k[h] = 0; if (k < F.n_posts) { xs_k[h] = F.xs[k]; idx_k[h] = F.sort[k]; } }
    #pragma unroll
    for (int j = 0; j < NC; j++) {
        unsigned m[2] = {0u, 0u};
        #pragma unroll
        for (int h = 0; h < H; h++) {
            const int k = lane + 32 * h;
            bool act = false;
            if (k < count[j]) { const int idx = idx_k[h]; act = idx < count[j] && ((flags[j] >> idx) & 1ull); ys[j][k] = fy[j][idx < count[j] ? idx : 0] * mult; }
            m[h] = __ballot_sync(0xffffffffu, act);
        }
        mask[j] = count[j] >= 2 ? ((((unsigned long long)m[1] << 32) | m[0]) | 1ull) : 0ull;
    }
    __syncwarp();
    #pragma unroll
    for (int j = 0; j < NC; j++) {
        bool careful = false;
        #pragma unroll
        for (int h = 0; h < H; h++) {
            const int k = lane + 32 * h;
            if ((mask[j] >> k) & 1ull) {
                WfSeg r; const int x0 = xs_k[h]; r.y0 = ys[j][k]; r.dy = 0; r.m = 1u;
                int x1 = 0xffff;                                            // the flat tail, Floor1.cs:213-216
                const unsigned long long above = mask[j] & ~(((1ull << k) << 1) - 1ull);
                if (above) {
                    const int hi = __ffsll((long long)above) - 1;
                    const int hx = F.xs[hi];
                    x1 = hx < n ? hx : n;                                   // x clamped, y NOT re-interpolated (Floor1.cs:206)
                    const int adx = x1 - x0;
                    r.dy = ys[j][hi] - r.y0;
                    const unsigned ady = (unsigned)(r.dy < 0 ? -r.dy : r.dy);
                    if (adx >= 2) { const uint2 mg = *reinterpret_cast<const uint2*>(magic + 2 * adx); r.m = ady <= mg.y ? mg.x : 0u; }
                }
                r.x01 = (unsigned)x0 | ((unsigned)x1 << 16);
                seg[j][k] = r;
                // y runs monotonically from y0 to y0 + dy: in range at both ends <=> in range everywhere
                if (x0 < n && (r.m == 0u || (unsigned)r.y0 > 255u || (unsigned)(r.y0 + r.dy) > 255u)) careful = true;
            }
        }
        careful_any[j] = __any_sync(0xffffffffu, careful);
    }
}

// Entry-stream offsets of every (partition, stage) of the frame by one warp: base[p * ST + st], plus cls[p] = class | coded
// stage mask << 8 (0 for a class byte outside the residue's range).  Order of the stream: stage-major, partitions ascending.
__device__ __forceinline__ void wf_entry_offsets(const DevSetup& S, const RunMode& rm, const uint8_t* coded, const uint8_t* cls, int P, int lane,
                                                 uint32_t* base, uint16_t* scls) {
    const int stages = rm.stages, ST = rm.base_stride, nclass = rm.nclass;
    const int nw = (stages + 3) >> 2;
    uint32_t stage0 = 0;                                                    // entries of all earlier stages
    for (int w = 0; w < nw; w++) {
        const unsigned long long* cc = S.cls_cnt + rm.cc_off + w * nclass;
        unsigned long long run = 0ull;
        for (int b0 = 0; b0 < P; b0 += 32) {
            const int p = b0 + lane;
            unsigned long long pk = 0ull; int cl = 255;
            if (p < P) { cl = cls[p]; if (cl < nclass) pk = cc[cl]; }
            unsigned lo = (unsigned)pk, hi = (unsigned)(pk >> 32);          // fields never carry into each other: a stage holds <= 32768 entries
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned ol = __shfl_up_sync(0xffffffffu, lo, d), oh = __shfl_up_sync(0xffffffffu, hi, d);
                if (lane >= d) { lo += ol; hi += oh; }
            }
            const unsigned el = (unsigned)run + lo - (unsigned)pk, eh = (unsigned)(run >> 32) + hi - (unsigned)(pk >> 32);   // exclusive, per field
            if (p < P) {
                uint4 v; v.x = el & 0xffffu; v.y = el >> 16; v.z = eh & 0xffffu; v.w = eh >> 16;
                *reinterpret_cast<uint4*>(base + (size_t)p * ST + 4 * w) = v;
                if (w == 0) scls[p] = cl < nclass ? (uint16_t)(cl | ((unsigned)coded[cl] << 8)) : (uint16_t)0;
            }
            const unsigned tl = __shfl_sync(0xffffffffu, lo, 31), th = __shfl_sync(0xffffffffu, hi, 31);
            run += ((unsigned long long)th << 32) | tl;
        }
        __syncwarp();
        // the four stages of this word start behind everything earlier: add the stage bases
        const uint32_t t0 = (unsigned)run & 0xffffu, t1 = (unsigned)run >> 16, t2 = (unsigned)(run >> 32) & 0xffffu, t3 = (unsigned)(run >> 48);
        const uint32_t s0 = stage0, s1 = s0 + t0, s2 = s1 + t1, s3 = s2 + t2;
        for (int p = lane; p < P; p += 32) {
            uint4 v = *reinterpret_cast<uint4*>(base + (size_t)p * ST + 4 * w);
            v.x += s0; v.y += s1; v.z += s2; v.w += s3;
            *reinterpret_cast<uint4*>(base + (size_t)p * ST + 4 * w) = v;
        }
        stage0 = s3 + t3;
    }
}

// ---- shared memory by 32-bit address (the generic-pointer form costs a window conversion per access)
#if !defined(NVB_CPU_SHIM)
typedef uint32_t wf_saddr;
__device__ __forceinline__ wf_saddr wf_smem(const void* p) { return (wf_saddr)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float wf_lds_f32(wf_saddr a) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ WfSeg wf_lds_seg(wf_saddr a) {
    WfSeg r; asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x01), "=r"(r.y0), "=r"(r.dy), "=r"(r.m) : "r"(a)); return r;
}
__device__ __forceinline__ uint32_t wf_lds_u32(wf_saddr a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ CiRec wf_lds_ci(wf_saddr a) {
    CiRec r; asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.off), "=r"(r.dshift), "=r"(r.entries), "=r"(r.cnt) : "r"(a)); return r;
}
__device__ __forceinline__ uint32_t wf_lds_u16(wf_saddr a) { uint16_t v; asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ void wf_sts_v4(wf_saddr a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void wf_sts_v2(wf_saddr a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory"); }
#else
typedef uintptr_t wf_saddr;
static inline wf_saddr wf_smem(const void* p) { return (wf_saddr)p; }
static inline float wf_lds_f32(wf_saddr a) { return *reinterpret_cast<const float*>(a); }
static inline WfSeg wf_lds_seg(wf_saddr a) { return *reinterpret_cast<const WfSeg*>(a); }
static inline uint32_t wf_lds_u32(wf_saddr a) { return *reinterpret_cast<const uint32_t*>(a); }
static inline CiRec wf_lds_ci(wf_saddr a) { return *reinterpret_cast<const CiRec*>(a); }
static inline uint32_t wf_lds_u16(wf_saddr a) { return *reinterpret_cast<const uint16_t*>(a); }
static inline void wf_sts_v4(wf_saddr a, float x, float y, float z, float w) { float* p = reinterpret_cast<float*>(a); p[0] = x; p[1] = y; p[2] = z; p[3] = w; }
static inline void wf_sts_v2(wf_saddr a, float x, float y) { float* p = reinterpret_cast<float*>(a); p[0] = x; p[1] = y; }
#endif

// Inverse coupling of one bin (Mapping.cs:145-181): the four sign cases are one add -- new = M + (same sign ? -A : A) -- and
// two selects (M - A and M + (-A) are the same IEEE operation).
__device__ __forceinline__ void inverse_couple_fast(float& m, float& a) {
    const float M = m, A = a;
    const bool mp = M > 0.f, ap = A > 0.f;
    const float t = NVB_FADD(M, (mp == ap) ? -A : A);
    m = ap ? M : t;
    a = ap ? t : M;
}

// Per-frame values of k_spectrum_wf's main loop (warp-uniform).
template <int CT, bool P64> struct WfFrame {
    typedef typename std::conditional<P64, unsigned long long, unsigned>::type mask_t;
    // offsets into the launch's arrays instead of pointers (the bases sit in the constant bank): half the registers
    uint32_t entries_off, spec_off, ci_off, bin2k_off;
    wf_saddr sdb, sseg, sbase, scls;                                        // inverse_dB_table, segments [CT][np], entry offsets [partition][ST], class words
    wf_saddr sci;                                                           // the setup's (class, stage) records, staged once per CTA
    uint32_t ecount, exec_mask;
    wf_saddr sout; uint32_t sout_stride, sout_swz;                           // SLOT: the frame's shared-memory slot, bytes between its channels, 1 = long block (chunks swizzled)
    int n, span, np, P, rbegin, pshift, ST, st_n, n_coupling, mapping;
    mask_t fmask[CT]; bool careful[CT];
};

// Runs of 16 consecutive stream values (two runs of 8: 16 / CT bins of every channel), stride GT.  PLAIN: every channel is
// executed with a floor curve on the multiply-high path (the common frame); CM: 0 no coupling step applies, 1 / 2 stereo
// with (magnitude, angle) = (0, 1) / (1, 0), 3 the general step list.
// SLOT: the spectrum goes into the frame's shared-memory slot (one-kernel synthesis, nvb_fused.cu) instead of the dense spectrum in
// HBM; 16-byte chunk j of a long block's channel sits at j ^ ((j >> 3) & 1), so that the stores of a quarter-warp (lanes 32 bytes
// apart) hit distinct banks; the transform's phase-1 loads apply the same permutation.
template <int CT, int GT, bool P64, bool PLAIN, int CM, bool SLOT = false>
__device__ __forceinline__ void wf_main(const LaunchArgs& a, const WfFrame<CT, P64>& x, int gt, int& bad_entry, int& bad_floor) {
    const uint16_t* __restrict__ ent = a.entries; const float* __restrict__ vq = a.S.vq; const CiRec* __restrict__ ci_tab = a.S.ci;
    const uint8_t* __restrict__ bin2k = a.S.bin2k;
    typedef typename WfFrame<CT, P64>::mask_t mask_t;
    constexpr int NB = 16 / CT;                                             // bins per channel in a double run
    const int pmask = (1 << x.pshift) - 1;
    for (int d = gt; d < (x.span >> 4); d += GT) {
        float acc[16];
        #pragma unroll
        for (int k = 0; k < 16; k++) acc[k] = 0.f;
        // ---- residue: the VQ vectors of the two runs, stage by stage from +0 (the float adds of WriteVectors in the reference's order).
        // (Issuing both runs' entry loads, then both runs' vector loads, before any add -- so that the two dependent load chains
        // overlap -- measured no different in round 2: 22.1 vs 22.0 us, more spills at 64 registers.)
        #pragma unroll
        for (int r8 = 0; r8 < 2; r8++) {
            float* ac = acc + 8 * r8;
            const int q = (d << 4) + 8 * r8 - x.rbegin, p = q >> x.pshift;
            const bool inr = q >= 0 && p < x.P;
            const unsigned cw = inr ? wf_lds_u16(x.scls + (wf_saddr)(2 * (inr ? p : 0))) : 0u;
            unsigned rest = cw >> 8;
            const int cl = (int)(cw & 0xffu);
            const int o = q & pmask;
            // (L1 policy hints on these loads -- evict_last for the VQ tables, evict_first for the entries -- measured slower in round 2:
            //  23.7 vs 21.8 us)
            while (rest) {                                                  // one iteration for most partitions
                const int st = __ffs(rest) - 1; rest &= rest - 1;
                // the (class, stage) record: from the CTA's shared-memory copy in the one-kernel path (its L1 is a few KB: 220 KB of the
                // SM's 256 are shared memory), through L1 in k_spectrum_wf (staging the table per 4-frame CTA cost more than it saved)
                const CiRec ci = SLOT ? wf_lds_ci(x.sci + (wf_saddr)(16 * (x.ci_off + (uint32_t)(cl * x.st_n + st)))) : ci_tab[x.ci_off + (uint32_t)(cl * x.st_n + st)];
                const uint32_t eb = wf_lds_u32(x.sbase + (wf_saddr)(4 * (p * x.ST + st)));
                if (ci.dshift >= 1) {                                       // every book with an even number of dimensions: four float2
                    const int dmask = (1 << ci.dshift) - 1;
                    uint32_t en[4]; bool ok[4];
                    #pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const uint32_t ei = eb + (uint32_t)((o + 2 * h) >> ci.dshift);
                        ok[h] = ei < x.ecount;
                        en[h] = 0u;
                        if (ok[h]) en[h] = ent[x.entries_off + ei];
                    }
                    #pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const bool good = ok[h] && en[h] < (uint32_t)ci.entries;
                        if (ok[h] && !good) bad_entry = 1;
                        float2 v = make_float2(0.f, 0.f);
                        if (good) v = *reinterpret_cast<const float2*>(vq + (uint32_t)(ci.off + (int)(en[h] << ci.dshift) + ((o + 2 * h) & dmask)));
                        if (good) { ac[2 * h] = NVB_FADD(ac[2 * h], v.x); ac[2 * h + 1] = NVB_FADD(ac[2 * h + 1], v.y); }
                    }
                } else {
                    #pragma unroll
                    for (int h = 0; h < 8; h++) {
                        const uint32_t ei = eb + (uint32_t)o + h;
                        if (ei < x.ecount) {
                            const uint32_t e1 = ent[x.entries_off + ei];
                            if (e1 < (uint32_t)ci.entries) ac[h] = NVB_FADD(ac[h], vq[(uint32_t)ci.off + e1]); else bad_entry = 1;
                        }
                    }
                }
            }
        }
        // ---- inverse coupling, last step first (Mapping.cs:137-182); the pairs of a bin sit in the same thread
        if (CM == 1) {
            #pragma unroll
            for (int b = 0; b < 8; b++) inverse_couple_fast(acc[2 * b], acc[2 * b + 1]);
        } else if (CM == 2) {
            #pragma unroll
            for (int b = 0; b < 8; b++) inverse_couple_fast(acc[2 * b + 1], acc[2 * b]);
        } else if (CM == 3) {
            for (int i = x.n_coupling - 1; i >= 0; --i) {
                const int m = a.S.mappings[x.mapping].mag[i], an = a.S.mappings[x.mapping].ang[i];
                if (!(((x.exec_mask >> m) | (x.exec_mask >> an)) & 1u)) continue;
                #pragma unroll
                for (int b = 0; b < NB; b++) {
                    float vm = 0.f, va = 0.f;
                    #pragma unroll
                    for (int k = 0; k < CT; k++) { if (k == m) vm = acc[b * CT + k]; if (k == an) va = acc[b * CT + k]; }
                    inverse_couple_fast(vm, va);
                    #pragma unroll
                    for (int k = 0; k < CT; k++) { if (k == m) acc[b * CT + k] = vm; if (k == an) acc[b * CT + k] = va; }
                }
            }
        }
        // ---- floor curve (Floor1.Apply, Floor1.cs:186-222) walked along the run, and the store
        const int bin0 = (d << 4) / CT;
        const unsigned kk0 = bin2k[x.bin2k_off + (uint32_t)bin0];           // sorted position of the last post at or below the first bin
        #pragma unroll
        for (int c = 0; c < CT; c++) {
            const mask_t M = x.fmask[c];
            if (PLAIN || (((x.exec_mask >> c) & 1u) && M != 0 && !x.careful[c])) {
                // the segment of the first bin: last active position at or below its post (bit 0 is set)
                int cur = P64 ? 63 - __clzll((long long)(M & (mask_t)(0xffffffffffffffffull >> (63 - kk0)))) : 31 - __clz((int)((unsigned)M & (0xffffffffu >> (31 - kk0))));
                const wf_saddr segc = x.sseg + (wf_saddr)(c * x.np) * (wf_saddr)sizeof(WfSeg);
                WfSeg r = wf_lds_seg(segc + (wf_saddr)cur * (wf_saddr)sizeof(WfSeg));
                int xrel = (int)(r.x01 >> 16) - bin0;                       // bins until the next active post
                unsigned ady = (unsigned)(r.dy < 0 ? -r.dy : r.dy);
                int sgn4 = r.dy < 0 ? -4 : 4;
                unsigned tt = (unsigned)(bin0 - (int)(r.x01 & 0xffffu)) * ady;
                wf_saddr base = x.sdb + (wf_saddr)(4 * r.y0);
                #pragma unroll
                for (int b = 0; b < NB; b++) {
                    if (b > 0 && xrel == b) {                               // the bin reaches the next active post: its segment starts here
                        const mask_t above = M & ~((((mask_t)1 << cur) << 1) - 1);
                        cur = P64 ? __ffsll((long long)above) - 1 : __ffs((int)above) - 1;
                        r = wf_lds_seg(segc + (wf_saddr)cur * (wf_saddr)sizeof(WfSeg));
                        xrel = (int)(r.x01 >> 16) - bin0; ady = (unsigned)(r.dy < 0 ? -r.dy : r.dy); sgn4 = r.dy < 0 ? -4 : 4;
                        tt = 0u; base = x.sdb + (wf_saddr)(4 * r.y0);
                    }
                    const int qq = (int)__umulhi(tt, r.m);
                    acc[b * CT + c] = NVB_FMUL(acc[b * CT + c], wf_lds_f32(base + (wf_saddr)(qq * sgn4)));
                    tt += ady;
                }
            } else if ((x.exec_mask >> c) & 1u) {
                if (M == 0) {                                               // no curve: the channel is cleared (Floor1.cs:220)
                    #pragma unroll
                    for (int b = 0; b < NB; b++) acc[b * CT + c] = 0.f;
                } else {                                                    // some segment needs the plain division or leaves inverse_dB_table's range
                    const wf_saddr segc = x.sseg + (wf_saddr)(c * x.np) * (wf_saddr)sizeof(WfSeg);
                    #pragma unroll
                    for (int b = 0; b < NB; b++) {
                        const unsigned kk = bin2k[x.bin2k_off + (uint32_t)(bin0 + b)];
                        const int lo = P64 ? 63 - __clzll((long long)(M & (mask_t)(0xffffffffffffffffull >> (63 - kk)))) : 31 - __clz((int)((unsigned)M & (0xffffffffu >> (31 - kk))));
                        const WfSeg r = wf_lds_seg(segc + (wf_saddr)lo * (wf_saddr)sizeof(WfSeg));
                        const int x0 = (int)(r.x01 & 0xffffu), adx = (int)(r.x01 >> 16) - x0;
                        const int num = (bin0 + b - x0) * (r.dy < 0 ? -r.dy : r.dy);
                        const int qq = r.m != 0u ? (int)__umulhi((unsigned)num, r.m) : num / adx;
                        int y = r.dy < 0 ? r.y0 - qq : r.y0 + qq;
                        if ((unsigned)y > 255u) { bad_floor = 1; y = y < 0 ? 0 : 255; }
                        acc[b * CT + c] = NVB_FMUL(acc[b * CT + c], wf_lds_f32(x.sdb + (wf_saddr)(4 * y)));
                    }
                }
            }
            if (SLOT) {
                constexpr int NV4 = (NB + 3) / 4;                           // float4 stores per channel (NB >= 4), else one float2
                if (NB >= 4) {
                    #pragma unroll
                    for (int v4 = 0; v4 < NV4; v4++) {
                        const int j = (bin0 >> 2) + v4;
                        wf_sts_v4(x.sout + (wf_saddr)c * x.sout_stride + (wf_saddr)(16 * (j ^ (((j >> 3) & 1) & (int)x.sout_swz))),
                                  acc[(4 * v4) * CT + c], acc[(4 * v4 + 1) * CT + c], acc[(4 * v4 + 2) * CT + c], acc[(4 * v4 + 3) * CT + c]);
                    }
                } else wf_sts_v2(x.sout + (wf_saddr)c * x.sout_stride + (wf_saddr)(4 * bin0), acc[c], acc[CT + c]);
                continue;
            }
            float* dst = a.spectrum + (x.spec_off + (uint32_t)(c * x.n + bin0));
            if (CT == 1) {
                #pragma unroll
                for (int v4 = 0; v4 < 4; v4++) reinterpret_cast<float4*>(dst)[v4] = make_float4(acc[4 * v4], acc[4 * v4 + 1], acc[4 * v4 + 2], acc[4 * v4 + 3]);
            } else if (CT == 2) {
                reinterpret_cast<float4*>(dst)[0] = make_float4(acc[c], acc[2 + c], acc[4 + c], acc[6 + c]);
                reinterpret_cast<float4*>(dst)[1] = make_float4(acc[8 + c], acc[10 + c], acc[12 + c], acc[14 + c]);
            } else if (CT == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[c], acc[4 + c], acc[8 + c], acc[12 + c]);
            else *reinterpret_cast<float2*>(dst) = make_float2(acc[c], acc[8 + c]);
        }
    }
}

// One frame's spectrum by ONE warp into the frame's shared-memory slot: the body of k_spectrum_wf<CT, 1, P64> (nvb_kernels.cu) for
// the one-kernel synthesis path (nvb_fused.cu).  gsm: the warp's scratch (WfLayout of one warp per frame), s_db: inverse_dB_table
// in shared memory, sout / sout_stride: the slot and the bytes between its channels.  The caller provides the warp barrier that
// makes the slot visible to the transform.
template <int CT, bool P64>
__device__ __forceinline__ void wf_frame_to_slot(const LaunchArgs& a, const DevFrame& f, const WfLayout& L, unsigned char* gsm, int* s_fy, const float* s_db, const CiRec* s_ci,
                                                 wf_saddr sout, uint32_t sout_stride, int lane, int& bad_entry, int& bad_floor) {
    constexpr int H = P64 ? 2 : 1;
    typedef typename WfFrame<CT, P64>::mask_t mask_t;
    const DevSetup& S = a.S;
    const RunMode rm = S.run_modes[f.mode];
    const DevFloor1& F = S.floors[rm.floor];
    const int n = f.n >> 1;
    WfFrame<CT, P64> x;
    x.n = n; x.span = CT * n; x.np = L.np_pad; x.rbegin = rm.rbegin; x.pshift = rm.pshift; x.ST = rm.base_stride;
    x.st_n = rm.stages > 0 ? rm.stages : 1; x.n_coupling = rm.n_coupling; x.mapping = rm.mapping;
    x.P = 0;
    if (f.res_decoded) { const int e = rm.rend < x.span ? rm.rend : x.span; const int nn = e - rm.rbegin; x.P = nn > 0 ? nn >> rm.pshift : 0; }   // Residue0.cs:122-127
    x.entries_off = f.entries_off; x.ci_off = (uint32_t)rm.ci_off; x.bin2k_off = (uint32_t)(rm.floor * (S.bs[1] >> 1));
    x.ecount = f.entry_count; x.exec_mask = f.exec_mask; x.spec_off = 0;
    x.sout = sout; x.sout_stride = sout_stride; x.sout_swz = (f.n == 2048) ? 1u : 0u;
    const uint8_t* cls = a.classes + f.classes_off;
    if (x.P > 0) for (uint32_t i = (uint32_t)lane * 64u; i < f.entry_count; i += 32 * 64u) prefetch_l1(a.entries + f.entries_off + i);

    WfSeg* s_seg = reinterpret_cast<WfSeg*>(gsm + L.seg_off);          // s_fy (256 ints of unwrap scratch) is only live in phase A: the caller lends the slot
    uint32_t* s_base = reinterpret_cast<uint32_t*>(gsm + L.base_off);
    uint16_t* s_cls = reinterpret_cast<uint16_t*>(gsm + L.cls_off);
    x.sdb = wf_smem(s_db); x.sseg = wf_smem(s_seg); x.sbase = wf_smem(s_base); x.scls = wf_smem(s_cls); x.sci = wf_smem(s_ci);

    #pragma unroll
    for (int c = 0; c < CT; c++) { x.fmask[c] = 0; x.careful[c] = false; }
    #pragma unroll
    for (int c = 0; c < CT; c++) {
        if (!((f.exec_mask >> c) & 1u)) continue;
        const int c2 = c + 1;
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
        }
        __syncwarp();
        if (pair) c = c2;
    }
    if (x.P > 0) wf_entry_offsets(S, rm, S.residues[rm.residue].coded, cls, x.P, lane, s_base, s_cls);
    __syncwarp();
    bool plain = (f.exec_mask & ((1u << CT) - 1u)) == ((1u << CT) - 1u);
    #pragma unroll
    for (int c = 0; c < CT; c++) plain = plain && x.fmask[c] != 0 && !x.careful[c];
    int cm = 0;
    const DevMapping& mp = S.mappings[rm.mapping];
    for (int i = 0; i < rm.n_coupling; i++) if (((f.exec_mask >> mp.mag[i]) | (f.exec_mask >> mp.ang[i])) & 1u) cm = 3;
    if (CT == 2 && cm == 3 && rm.n_coupling == 1) cm = mp.mag[0] == 0 ? 1 : 2;
    if (plain) {
        if (cm == 1) wf_main<CT, 32, P64, true, CT == 2 ? 1 : 3, true>(a, x, lane, bad_entry, bad_floor);
        else if (cm == 0) wf_main<CT, 32, P64, true, 0, true>(a, x, lane, bad_entry, bad_floor);
        else wf_main<CT, 32, P64, true, 3, true>(a, x, lane, bad_entry, bad_floor);
    } else wf_main<CT, 32, P64, false, 3, true>(a, x, lane, bad_entry, bad_floor);
}

// The one-kernel path's per-warp scratch: segment records, entry offsets, class words (the unwrap scratch lives in the frame's slot).
static WfLayout wf_layout_slot(const DevSetup& S, int CT) {
    WfLayout L;
    L.np_pad = S.max_posts <= 32 ? 32 : 64;
    const int st_max = ((S.max_stages + 3) & ~3) > 0 ? ((S.max_stages + 3) & ~3) : 4;
    const int pmax = S.wf_max_p > 0 ? S.wf_max_p : 1;
    L.seg_off = 0;
    L.fy_off = 0;
    L.base_off = CT * L.np_pad * (int)sizeof(WfSeg);
    L.cls_off = L.base_off + pmax * st_max * (int)sizeof(uint32_t);
    L.total = (L.cls_off + pmax * (int)sizeof(uint16_t) + 15) & ~15;
    L.cta_bytes = (S.ci_total > 0 ? S.ci_total : 1) * (int)sizeof(CiRec);
    return L;
}

static WfLayout wf_layout(const DevSetup& S, int CT, int WPF) {
    WfLayout L;
    L.np_pad = S.max_posts <= 32 ? 32 : 64;
    const int st_max = ((S.max_stages + 3) & ~3) > 0 ? ((S.max_stages + 3) & ~3) : 4;
    const int pmax = S.wf_max_p > 0 ? S.wf_max_p : 1;
    L.seg_off = 0;
    L.fy_off = L.seg_off + CT * L.np_pad * (int)sizeof(WfSeg);
    L.base_off = (L.fy_off + WPF * 256 * (int)sizeof(int) + CT * 4 * (int)sizeof(int) + 15) & ~15;
    L.cls_off = L.base_off + pmax * st_max * (int)sizeof(uint32_t);
    L.total = (L.cls_off + pmax * (int)sizeof(uint16_t) + 15) & ~15;
    L.cta_bytes = 0;
    return L;
}

}  // namespace nvb
