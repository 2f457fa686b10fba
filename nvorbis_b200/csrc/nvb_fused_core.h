// nvb_fused_core.h -- lane-level arithmetic of the fused IMDCT + window + OLA + clip + interleave
// kernel (nvb_fused.cu).  Written as phases of one warp: every phase is a function of
// (lane, registers, shared memory) with a warp barrier between phases.
//
// IMDCT factorisation (M = N/2 coefficients, Q = N/4 complex points; reproduces what Mdct.Reverse
// computes for N >= 256, i.e. y[i] = sum_k X[k] cos(2pi/N (i + 1/2 + N/4)(k + 1/2)), Mdct.cs:65-313):
//   c[k] = (X[2k] + i X[M-1-2k]) * tw[k],   tw[k] = exp(-i pi (k + 1/8) / M)
//   Cf   = FFT_Q(c);   D[n] = Cf[n] * tw[n];   u[2n] = Re D[n],  u[M-1-2n] = -Im D[n]
//   y[i] = u[i+M/2] (i < M/2);  -u[3M/2-1-i] (M/2 <= i < 3M/2);  -u[i-3M/2] (i >= 3M/2)
// This path may contract to FMA; it is held to <= 1e-5 max-abs against the reference arithmetic (the exact
// path in nvb_kernels.cu is the bit-identical one).
//
// Long block (N = 2048, Q = 512 = 8*8*8): one warp, 16 points per lane as two radix-8 columns, three
// passes with two exchanges through shared memory.  Every twiddle a lane needs is the same for every
// transform, so the host lays the twiddles out per lane ("lane tables", FusedTables below): each load is
// a conflict-free LDS.128 of consecutive lanes.
#pragma once
#include "nvb_device_core.h"

namespace nvb {

struct cpx { float x, y; };
NVB_HD cpx cmul(cpx a, cpx b) { cpx r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
#if defined(__CUDA_ARCH__)
// Blackwell packed FP32: one add.rn.f32x2 (SASS FADD2) per complex add -- same IEEE results as two scalar adds, half
// the issue slots; the butterflies are add-dominated and the kernel is issue-bound.
__device__ __forceinline__ cpx cadd2_(float ax, float ay, float bx, float by) {
    unsigned long long ua, ub, ud; cpx r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(ax), "f"(ay));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(bx), "f"(by));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(ud));
    return r;
}
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return cadd2_(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return cadd2_(a.x, a.y, -b.x, -b.y); }
#else
NVB_HD cpx cadd(cpx a, cpx b) { cpx r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
NVB_HD cpx csub(cpx a, cpx b) { cpx r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
#endif
NVB_HD cpx cmul_mi(cpx a) { cpx r; r.x = a.y; r.y = -a.x; return r; }          // a * (-i)

// 8-point forward DFT (W8 = exp(-2 pi i / 8)), in place, natural order in and out.
NVB_HD void fft8(cpx* a) {
    const float h = 0.70710678118654752440f;
    cpx b0 = cadd(a[0], a[4]), b4 = csub(a[0], a[4]);
    cpx b1 = cadd(a[1], a[5]), t5 = csub(a[1], a[5]);
    cpx b2 = cadd(a[2], a[6]), t6 = csub(a[2], a[6]);
    cpx b3 = cadd(a[3], a[7]), t7 = csub(a[3], a[7]);
    cpx b5; b5.x = (t5.x + t5.y) * h; b5.y = (t5.y - t5.x) * h;                 // * (1 - i)/sqrt2
    cpx b6 = cmul_mi(t6);                                                       // * (-i)
    cpx b7; b7.x = (t7.y - t7.x) * h; b7.y = -(t7.x + t7.y) * h;                // * (-1 - i)/sqrt2
    // even outputs: FFT4(b0..b3); odd outputs: FFT4(b4..b7)
    cpx e0 = cadd(b0, b2), e1 = csub(b0, b2), e2 = cadd(b1, b3), e3 = cmul_mi(csub(b1, b3));
    cpx o0 = cadd(b4, b6), o1 = csub(b4, b6), o2 = cadd(b5, b7), o3 = cmul_mi(csub(b5, b7));
    a[0] = cadd(e0, e2); a[4] = csub(e0, e2); a[2] = cadd(e1, e3); a[6] = csub(e1, e3);
    a[1] = cadd(o0, o2); a[5] = csub(o0, o2); a[3] = cadd(o1, o3); a[7] = csub(o1, o3);
}

constexpr int FUSED_SLOT_FLOATS = 1152;        // 8 rows * 72 float2 of exchange space >= 1024 floats of u
constexpr int FUSED_LONG_N = 2048;
constexpr int FUSED_SHORT_N = 256;

// Lane tables for N = 2048 / 256, built on the host (nvb_host.cpp: build_fused_tables), one contiguous block
// so that the kernel stages it with a single bulk copy.  Offsets in floats.
// Twiddle folding: tw[64 j + r] = tw[r] * E[j] with the compile-time constants E[j] = exp(-i pi j / 16), so the
// pre-twiddle of pass 1 becomes 7 constant multiplies on the inputs plus tw[r] merged into the pass-1 output
// twiddles (T2), and the post-twiddle becomes 7 constant multiplies plus one per-lane value (T4).  That is 13
// vector loads of twiddles per transform instead of 27: the kernel is bound by shared-memory wavefronts.
struct FusedTables {
    static constexpr int T2 = 0;        // [8][32] float4: tw[l] W512^(l m2), tw[63-l] W512^((63-l) m2), m2 = 0..7
    static constexpr int T3 = 1024;     // [4][32] float4: W64^(k0 m1) for m1 = 2j+1, 2j+2 (k0 = l & 7)
    static constexpr int T4 = 1536;     // [32] float4: tw[na0], tw[nb0]
    static constexpr int WIN = 1664;    // [1024] rising slope of the long/long window (Mode.cs:80-85)
    static constexpr int TW0 = 2688;    // [64] float2: short-block twiddles exp(-i pi (k + 1/8) / 128)
    static constexpr int W64 = 2816;    // [64] float2: exp(-2 pi i k / 64)
    static constexpr int WIN0 = 2944;   // [128] rising slope of the short window: with WIN, every window shape of Mode.cs:24-67
    static constexpr int FLOATS = 3072;
};

// Window value of sample i of a block (N = 2048 with window index widx = prev?1:0 + next?2:0, Mode.cs:44-50, or N = 256) from
// the two slopes in shared memory: the shapes assemble_window() (nvb_host.cpp) builds -- zeros, a rising slope, ones, the
// mirrored slope, zeros -- value for value.
NVB_HD float fused_window_at(const float* s_win, const float* s_win0, int n, int widx, int i) {
    if (n == FUSED_SHORT_N) return i < 128 ? s_win0[i] : s_win0[255 - i];
    if (i < 1024) {
        if (widx & 1) return s_win[i];
        return i < 448 ? 0.f : (i < 576 ? s_win0[i - 448] : 1.f);
    }
    if (widx & 2) return s_win[2047 - i];
    return i < 1472 ? 1.f : (i < 1600 ? s_win0[1599 - i] : 0.f);
}
NVB_HD int fused_na0(int l) { return (l >> 3) + 8 * (l & 7); }
NVB_HD int fused_nb0(int l) { return (7 - (l >> 3)) + 8 * (7 - (l & 7)); }

// a * E[K], E[K] = exp(-i pi K / 16) = (cos, -sin)(pi K / 16)
template <int K> NVB_HD cpx cmul_e(cpx a) {
    constexpr float co[8] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f, 0.70710678118654752440f,
                             0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
    constexpr float si[8] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f, 0.70710678118654752440f,
                             0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f};
    if (K == 0) return a;
    cpx r; r.x = a.x * co[K] + a.y * si[K]; r.y = a.y * co[K] - a.x * si[K];
    return r;
}
NVB_HD void cmul_e_all(cpx* a) {
    a[1] = cmul_e<1>(a[1]); a[2] = cmul_e<2>(a[2]); a[3] = cmul_e<3>(a[3]); a[4] = cmul_e<4>(a[4]);
    a[5] = cmul_e<5>(a[5]); a[6] = cmul_e<6>(a[6]); a[7] = cmul_e<7>(a[7]);
}

// u of an executed long block is stored swizzled: float2 index n -> n ^ (((n >> 4) & 3) << 1).  That makes the
// phase-3 stores (lanes 8 apart in n) and the output reads (lanes adjacent in n) both bank-conflict free, and
// keeps aligned groups of four floats contiguous.
NVB_HD int u_swz2(int n) { return n ^ (((n >> 4) & 3) << 1); }                  // float2 index
NVB_HD int u_swz(int i) { return i ^ (((i >> 5) & 3) << 2); }                   // float index

NVB_HD cpx ld_cpx(const float4& v, int hi) { cpx r; r.x = hi ? v.z : v.x; r.y = hi ? v.w : v.y; return r; }

// Registers of one lane while it transforms one long block: two radix-8 columns.
struct LongRegs { cpx a[8]; cpx b[8]; };

// ---- phase 1: load spectrum pairs, constant part of the pre-twiddle, radix-8 over k2, twiddle tw[r] W512^(r*m2), store ex1
// spec2: the channel's spectrum as float2[512] (global memory); tab: FusedTables in shared memory; ex: float2[576].
struct LongIn { float2 pa[8], pb[8]; };
NVB_HD void long_phase1_load(int l, const float2* spec2, LongIn& in) {
    const int ra = l, rb = 63 - l;
    #pragma unroll
    for (int k2 = 0; k2 < 8; k2++) { in.pa[k2] = spec2[64 * k2 + ra]; in.pb[k2] = spec2[64 * k2 + rb]; }
}
// The same rows out of the frame's slot (one-kernel synthesis): the spectrum stage stores 16-byte chunk j of a long block at
// j ^ ((j >> 3) & 1), i.e. float2 m at m ^ (((m >> 4) & 1) << 1) -- bit 4 of 64 k2 + l is bit 4 of l, of 64 k2 + 63 - l its complement.
NVB_HD void long_phase1_load_slot(int l, const float2* slot2, LongIn& in) {
    const int ra = l ^ (((l >> 4) & 1) << 1), rb = (63 - l) ^ ((((63 - l) >> 4) & 1) << 1);
    #pragma unroll
    for (int k2 = 0; k2 < 8; k2++) { in.pa[k2] = slot2[64 * k2 + ra]; in.pb[k2] = slot2[64 * k2 + rb]; }
}
NVB_HD void long_phase1_compute(int l, const LongIn& in, const float* tab, float2* ex) {
    const int ra = l, rb = 63 - l;
    const float4* T2 = reinterpret_cast<const float4*>(tab + FusedTables::T2);
    LongRegs R;
    #pragma unroll
    for (int k2 = 0; k2 < 8; k2++) {
        R.a[k2].x = in.pa[k2].x; R.a[k2].y = in.pb[7 - k2].y;    // X[2k] + i X[M-1-2k], k = 64 k2 + l
        R.b[k2].x = in.pb[k2].x; R.b[k2].y = in.pa[7 - k2].y;    // k = 64 k2 + 63 - l
    }
    cmul_e_all(R.a); cmul_e_all(R.b);
    fft8(R.a); fft8(R.b);
    #pragma unroll
    for (int m2 = 0; m2 < 8; m2++) {
        const float4 w = T2[m2 * 32 + l];
        const cpx va = cmul(R.a[m2], ld_cpx(w, 0)), vb = cmul(R.b[m2], ld_cpx(w, 1));
        ex[m2 * 72 + ra] = make_float2(va.x, va.y);
        ex[m2 * 72 + rb] = make_float2(vb.x, vb.y);
    }
}

// ---- phase 2: gather the 8 k1 of (m2, k0) for two m2; radix-8 over k1, twiddle W64^(k0*m1), store ex2
NVB_HD void long_phase2_load(int l, const float2* ex, LongRegs& R) {
    const int m2 = l >> 3, k0 = l & 7;
    #pragma unroll
    for (int k1 = 0; k1 < 8; k1++) {
        float2 va = ex[m2 * 72 + 8 * k1 + k0], vb = ex[(m2 + 4) * 72 + 8 * k1 + k0];
        R.a[k1].x = va.x; R.a[k1].y = va.y; R.b[k1].x = vb.x; R.b[k1].y = vb.y;
    }
}
// The pass-2 twiddles W64^(k0 m1) and the post-twiddle pair of a lane never change: they live in registers.
struct LaneTw { cpx t3[7]; cpx t4a, t4b; };
NVB_HD void lane_tw_load(int l, const float* tab, LaneTw& w) {
    const float4* T3 = reinterpret_cast<const float4*>(tab + FusedTables::T3);
    #pragma unroll
    for (int j = 0; j < 4; j++) {
        const float4 v = T3[j * 32 + l];
        w.t3[2 * j] = ld_cpx(v, 0);
        if (2 * j + 1 < 7) w.t3[2 * j + 1] = ld_cpx(v, 1);
    }
    const float4 v = reinterpret_cast<const float4*>(tab + FusedTables::T4)[l];
    w.t4a = ld_cpx(v, 0); w.t4b = ld_cpx(v, 1);
}
NVB_HD void long_phase2_store(int l, const LaneTw& w, float2* ex, LongRegs& R) {
    const int m2 = l >> 3, k0 = l & 7;
    fft8(R.a); fft8(R.b);
    ex[m2 * 72 + k0] = make_float2(R.a[0].x, R.a[0].y);
    ex[(m2 + 4) * 72 + k0] = make_float2(R.b[0].x, R.b[0].y);
    #pragma unroll
    for (int m1 = 1; m1 < 8; m1++) {
        const cpx va = cmul(R.a[m1], w.t3[m1 - 1]), vb = cmul(R.b[m1], w.t3[m1 - 1]);
        ex[m2 * 72 + m1 * 9 + k0] = make_float2(va.x, va.y);
        ex[(m2 + 4) * 72 + m1 * 9 + k0] = make_float2(vb.x, vb.y);
    }
}

// ---- phase 3: gather the 8 k0 of (m2, m1) and of (7-m2, 7-m1); radix-8 over k0, post-twiddle E[m0] tw[n0],
// write u as float2 pairs (u[2n], u[2n+1]) = (Re D[n], -Im D[511-n]), swizzled (u_swz2).
NVB_HD void long_phase3_load(int l, const float2* ex, LongRegs& R) {
    const int m2 = l >> 3, m1 = l & 7;
    #pragma unroll
    for (int k0 = 0; k0 < 8; k0++) {
        float2 va = ex[m2 * 72 + m1 * 9 + k0], vb = ex[(7 - m2) * 72 + (7 - m1) * 9 + k0];
        R.a[k0].x = va.x; R.a[k0].y = va.y; R.b[k0].x = vb.x; R.b[k0].y = vb.y;
    }
}
NVB_HD void long_phase3_store(int l, const LaneTw& w, float2* u2, LongRegs& R) {
    fft8(R.a); fft8(R.b);
    cmul_e_all(R.a); cmul_e_all(R.b);
    const cpx wa = w.t4a, wb = w.t4b;
    #pragma unroll
    for (int m0 = 0; m0 < 8; m0++) { R.a[m0] = cmul(R.a[m0], wa); R.b[m0] = cmul(R.b[m0], wb); }
    // n = na0 + 64 m0 and its partner 511 - n = nb0 + 64 (7 - m0), nb0 = 63 - na0; the swizzle only looks at
    // bits 4-5 of n, which belong to na0 / nb0
    const int sa = u_swz2(fused_na0(l)), sb = u_swz2(fused_nb0(l));
    #pragma unroll
    for (int m0 = 0; m0 < 8; m0++) {
        u2[sa + 64 * m0] = make_float2(R.a[m0].x, -R.b[7 - m0].y);
        u2[sb + 64 * (7 - m0)] = make_float2(R.b[7 - m0].x, -R.a[m0].y);
    }
}

// ---- short block (N = 256, Q = 64): lane holds points l and l+32.
struct ShortRegs { cpx a, b; };
NVB_HD void short_phase1(int l, const float* spec, const float2* tw64, const float2* w64, ShortRegs& R) {
    cpx ca, cb, t;
    ca.x = spec[2 * l]; ca.y = spec[127 - 2 * l];
    cb.x = spec[2 * l + 64]; cb.y = spec[63 - 2 * l];
    float2 w = tw64[l]; t.x = w.x; t.y = w.y; ca = cmul(ca, t);
    w = tw64[l + 32]; t.x = w.x; t.y = w.y; cb = cmul(cb, t);
    R.a = cadd(ca, cb);
    w = w64[l]; t.x = w.x; t.y = w.y;
    R.b = cmul(csub(ca, cb), t);
}
// One shuffle-DIF stage of the two 32-point FFTs (half-size s): own value v, partner's value p.
NVB_HD cpx short_stage(int l, int s, cpx v, cpx p, const float2* w64) {
    const int j = l & (s - 1);
    if (l & s) {
        float2 w = w64[(j * (32 / s)) & 63];            // W_{2s}^j = W64^(j*32/s)
        cpx t; t.x = w.x; t.y = w.y;
        return cmul(csub(p, v), t);
    }
    return cadd(v, p);
}
NVB_HD int brev5(int l) { return ((l & 1) << 4) | ((l & 2) << 2) | (l & 4) | ((l & 8) >> 2) | ((l & 16) >> 4); }
NVB_HD void short_phase3_store(int l, const float2* tw64, float* u, const ShortRegs& R) {
    const int m = brev5(l);
    const int ne = 2 * m, no = 2 * m + 1;
    float2 w = tw64[ne]; cpx t; t.x = w.x; t.y = w.y; cpx de = cmul(R.a, t);
    w = tw64[no]; t.x = w.x; t.y = w.y; cpx dodd = cmul(R.b, t);
    u[2 * ne] = de.x; u[127 - 2 * ne] = -de.y;
    u[2 * no] = dodd.x; u[127 - 2 * no] = -dodd.y;
}

// ---- output side ------------------------------------------------------------------------------------
// Un-windowed block value y[i] of a slot: u (executed channel) or the raw spectrum (Mapping.cs:192-196).
// Only the u of an executed long block is swizzled.
// Where y[i] of an executed channel sits inside its slot: y[i] = sgn * slot[j].
// swz: the slot holds the swizzled u of the specialised N = 2048 transform (k_imdct_fused); k_imdct_generic stores u plainly.
NVB_HD void fused_y_index(int N, int i, int& j, float& sgn, bool swz = true) {
    const int M = N >> 1, h = M >> 1;
    if (i < h) { j = i + h; sgn = 1.f; }
    else if (i < M + h) { j = M + h - 1 - i; sgn = -1.f; }
    else { j = i - M - h; sgn = -1.f; }
    if (swz && N == FUSED_LONG_N) j = u_swz(j);
}

NVB_HD float fused_y(const float* slot, bool exec, int N, int i, bool swz = true) {
    const int M = N >> 1;
    if (!exec) return i < M ? slot[i] : 0.f;
    int j; float sgn;
    fused_y_index(N, i, j, sgn, swz);
    return sgn * slot[j];
}

}  // namespace nvb
