// nvb_device_core.h -- the arithmetic of the synthesis path as host/device inline functions.
//
// The CUDA kernels in nvb_kernels.cu are thin thread-mapping shells around these functions.
// Every function is written as "work item `t` of `nt`" so the same code can be single-stepped
// on the host by tests/cpu_shim.cpp (no GPU in the build container); that shim is test-only.
//
// Float discipline: everything that must be bit-identical to the reference uses explicit
// NVB_FMUL / NVB_FADD / NVB_FSUB (round-to-nearest, never contracted into FMA), mirroring
// RyuJIT's scalar SSE code.  Reference file:line citations are relative to NVorbis/.
#pragma once
#include "nvb_internal.h"

#if defined(__CUDACC__)
#define NVB_HD __host__ __device__ __forceinline__
#else
#define NVB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define NVB_FMUL(a, b) __fmul_rn((a), (b))
#define NVB_FADD(a, b) __fadd_rn((a), (b))
#define NVB_FSUB(a, b) __fsub_rn((a), (b))
#define NVB_FDIV(a, b) __fdiv_rn((a), (b))
#else
// host build of the shim is compiled with -ffp-contract=off
#define NVB_FMUL(a, b) ((a) * (b))
#define NVB_FADD(a, b) ((a) + (b))
#define NVB_FSUB(a, b) ((a) - (b))
#define NVB_FDIV(a, b) ((a) / (b))
#endif

namespace nvb {

NVB_HD int ilog_u(int x) { int c = 0; while (x > 0) { ++c; x >>= 1; } return c; }   // Utils.cs:5-14

// ------------------------------------------------------------------------------------------------
// Residue geometry of one frame (Residue0.Decode head, Residue0.cs:122-127; Residue2.cs:16-21)
// ------------------------------------------------------------------------------------------------
struct ResGeom { int P, Sx, n_items; };
NVB_HD ResGeom residue_geom(const DevResidue& R, int N, int C) {
    int bs = (R.type == 2) ? N * C : N;
    int e = R.end < bs / 2 ? R.end : bs / 2;
    int nn = e - R.begin;
    ResGeom g;
    g.P = nn > 0 ? nn / R.psize : 0;
    g.Sx = (R.type == 2) ? 1 : C;
    g.n_items = R.stages * g.P * g.Sx;
    return g;
}

// Number of VQ entries the item (stage s, partition p, stream st) consumes from the entry stream.
// Order of items == order in which Residue0.Decode visits them (stage, partition, channel).
NVB_HD uint32_t residue_item_count(const DevResidue& R, const DevBook* books, const uint8_t* cls, const ResGeom& g, int idx) {
    int per_stage = g.P * g.Sx;
    int s = idx / per_stage; int rem = idx - s * per_stage;
    int p = rem / g.Sx; int st = rem - p * g.Sx;
    int c = cls[st * g.P + p];
    if (c >= R.nclass) return 0;
    if (!((R.cascade[c] >> s) & 1)) return 0;
    int book = R.books[c][s];
    if (book < 0) return 0;
    int dims = books[book].dims;
    return (uint32_t)(R.type == 0 ? R.psize / dims : (R.psize + dims - 1) / dims);   // Residue0.cs:183 / Residue1.cs:12 / Residue2.cs:28
}

// One VQ contribution: book[entry, d]   (Codebook.cs:322)
NVB_HD float vq_fetch(const DevBook& b, const float* vq, const uint16_t* ent, uint32_t idx, uint32_t ent_cnt, int d, int* bad) {
    if (idx >= ent_cnt) return 0.f;          // never decoded: packet ended ("use what we have", Residue0.cs:164-170)
    int e = ent[idx];
    if (e >= b.entries || b.off < 0) { if (bad) *bad = 1; return 0.f; }
    return vq[b.off + (int64_t)e * b.dims + d];
}

// Residue value of (channel c, bin j): the sum, in the reference's order (stage-major), of every
// `res[o] += book[entry,dim]` that lands on that bin.  Gather instead of scatter: no atomics and the
// float additions happen in exactly the reference's order, starting from the cleared buffer (+0.0f).
NVB_HD float residue_value(const DevResidue& R, const DevBook* books, const float* vq, const uint8_t* cls, const uint16_t* ent,
                           uint32_t ent_cnt, const uint32_t* prefix, const ResGeom& g, int C, int c, int j, int* bad) {
    float acc = 0.f;
    if (g.P == 0) return acc;
    if (R.type != 2) {
        // types 0/1: stream == channel, bin == position (Residue0.cs:155-159)
        int q = j - R.begin;
        if (q < 0) return acc;
        int p = q / R.psize;
        if (p >= g.P) return acc;
        int o = q - p * R.psize;
        int cl = cls[c * g.P + p];
        if (cl >= R.nclass) return acc;
        int casc = R.cascade[cl];
        for (int s = 0; s < R.stages; s++) {
            if (!((casc >> s) & 1)) continue;
            int book = R.books[cl][s];
            if (book < 0) continue;
            const DevBook& b = books[book];
            uint32_t base = prefix[(s * g.P + p) * g.Sx + c];
            if (R.type == 1) {                                   // Residue1.cs:12-23
                acc = NVB_FADD(acc, vq_fetch(b, vq, ent, base + o / b.dims, ent_cnt, o % b.dims, bad));
            } else {                                             // Residue0.cs:193-199: res[offset + dim*steps + step]
                int steps = R.psize / b.dims;
                if (steps == 0 || o >= steps * b.dims) continue;
                acc = NVB_FADD(acc, vq_fetch(b, vq, ent, base + o % steps, ent_cnt, o / steps, bad));
            }
        }
        return acc;
    }
    // type 2 (Residue2.cs:23-47): partition p starts at interleaved offset i0 = begin + p*psize, the
    // reference restarts chPtr = 0 there and uses bin offset i0 / C; element e of the partition lands on
    // channel e % C, bin i0/C + e/C.  When i0 is not a multiple of C this differs from the spec and two
    // neighbouring partitions can hit the same (channel, bin): both are visited, in partition order.
    long long top = (long long)(j + 1) * C - 1 - R.begin;
    if (top < 0) return acc;
    int p_hi = (int)(top / R.psize);
    if (p_hi > g.P - 1) p_hi = g.P - 1;
    int cand[3]; int ncand = 0;
    for (int p = p_hi; p >= 0 && ncand < 3; --p) {
        int ob = (R.begin + p * R.psize) / C;
        long long e = (long long)(j - ob) * C + c;
        if (e >= R.psize) break;
        cand[ncand++] = p;
    }
    for (int s = 0; s < R.stages; s++) {
        for (int k = ncand - 1; k >= 0; --k) {
            int p = cand[k];
            int cl = cls[p];
            if (cl >= R.nclass) continue;
            if (!((R.cascade[cl] >> s) & 1)) continue;
            int book = R.books[cl][s];
            if (book < 0) continue;
            const DevBook& b = books[book];
            int ob = (R.begin + p * R.psize) / C;
            int e = (j - ob) * C + c;
            uint32_t base = prefix[s * g.P + p];
            acc = NVB_FADD(acc, vq_fetch(b, vq, ent, base + e / b.dims, ent_cnt, e % b.dims, bad));
        }
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------
// Inverse channel coupling of one bin (Mapping.cs:145-181).  Exact: compares + one add/sub.
// ------------------------------------------------------------------------------------------------
NVB_HD void inverse_couple(float& m, float& a) {
    float oldM = m, oldA = a, newM, newA;
    if (oldM > 0) {
        if (oldA > 0) { newM = oldM; newA = NVB_FSUB(oldM, oldA); } else { newA = oldM; newM = NVB_FADD(oldM, oldA); }
    } else {
        if (oldA > 0) { newM = oldM; newA = NVB_FADD(oldM, oldA); } else { newA = oldM; newM = NVB_FSUB(oldM, oldA); }
    }
    m = newM; a = newA;
}

// ------------------------------------------------------------------------------------------------
// Floor 1: UnwrapPosts (Floor1.cs:224-297) + the segment walk of Apply (Floor1.cs:196-216),
// serial over <= 64 posts, run by one lane per (frame, channel).  Produces line segments
// (x0, y0) -> (x1, y1) with x1 already clamped to n but y1 NOT re-interpolated (Floor1.cs:206).
// ------------------------------------------------------------------------------------------------
struct FloorSegs {
    int16_t n;                       // number of segments (0 => channel spectrum is cleared, Floor1.cs:220)
    int16_t x0[NVB_MAX_POSTS + 1];   // ascending
    int16_t x1[NVB_MAX_POSTS + 1];
    int16_t y0[NVB_MAX_POSTS + 1];
    int16_t y1[NVB_MAX_POSTS + 1];
};

NVB_HD int render_point(int x0, int y0, int x1, int y1, int X) {                                 // Floor1.cs:299-314
    int dy = y1 - y0; int adx = x1 - x0; int ady = dy < 0 ? -dy : dy;
    int err = ady * (X - x0); int off = err / adx;
    return dy < 0 ? y0 - off : y0 + off;
}

// posts: element 0 = PostCount, 1.. = raw Y values
NVB_HD void floor1_build(const DevFloor1& F, const int16_t* posts, int n, FloorSegs& out) {
    int count = posts[0];
    out.n = 0;
    if (count <= 0) return;
    if (count > F.n_posts) count = F.n_posts;
    int finalY[NVB_MAX_POSTS]; unsigned long long flags = 3ull;          // stepFlags[0] = stepFlags[1] = true
    finalY[0] = posts[1]; finalY[1] = posts[2];
    for (int i = 2; i < count; i++) {
        int lowOfs = F.lo[i], highOfs = F.hi[i];
        int predicted = render_point(F.x[lowOfs], finalY[lowOfs], F.x[highOfs], finalY[highOfs], F.x[i]);
        int val = posts[1 + i];
        int highroom = F.range - predicted, lowroom = predicted;
        int room = (highroom < lowroom) ? highroom * 2 : lowroom * 2;
        if (val != 0) {
            flags |= (1ull << lowOfs) | (1ull << highOfs) | (1ull << i);
            if (val >= room) {
                if (highroom > lowroom) finalY[i] = val - lowroom + predicted;
                else finalY[i] = predicted - val + highroom - 1;
            } else {
                if ((val % 2) == 1) finalY[i] = predicted - ((val + 1) / 2);
                else finalY[i] = predicted + (val / 2);
            }
        } else {
            flags &= ~(1ull << i);
            finalY[i] = predicted;
        }
    }
    int lx = 0, ly = finalY[0] * F.mult; int k = 0;
    for (int i = 1; i < count; i++) {
        int idx = F.sort[i];
        if ((flags >> idx) & 1ull) {
            int hx = F.x[idx], hy = finalY[idx] * F.mult;
            if (lx < n) { out.x0[k] = (int16_t)lx; out.y0[k] = (int16_t)ly; out.x1[k] = (int16_t)(hx < n ? hx : n); out.y1[k] = (int16_t)hy; ++k; }
            lx = hx; ly = hy;
        }
        if (lx >= n) break;
    }
    if (lx < n) { out.x0[k] = (int16_t)lx; out.y0[k] = (int16_t)ly; out.x1[k] = (int16_t)n; out.y1[k] = (int16_t)ly; ++k; }
    out.n = (int16_t)k;
}

// y of RenderLineMulti (Floor1.cs:316-341) at bin j in closed form:
//   y(x0+k) = y0 + k*b + sy*floor(k*ady'/adx),  b = trunc(dy/adx), ady' = |dy| - |b|*adx, sy = sign(dy)
NVB_HD int floor1_y(const FloorSegs& s, int j) {
    int lo = 0, hi = s.n - 1;                 // largest k with x0[k] <= j
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (s.x0[mid] <= j) lo = mid; else hi = mid - 1; }
    int x0 = s.x0[lo], y0 = s.y0[lo], x1 = s.x1[lo], y1 = s.y1[lo];
    int dy = y1 - y0, adx = x1 - x0;
    int ady = dy < 0 ? -dy : dy;
    int b = dy / adx;
    int ab = b < 0 ? -b : b;
    ady -= ab * adx;
    int sy = dy < 0 ? -1 : 1;
    int k = j - x0;
    return y0 + k * b + sy * ((k * ady) / adx);
}

// ------------------------------------------------------------------------------------------------
// Utils.ClipValue (Utils.cs:30-43)
// ------------------------------------------------------------------------------------------------
NVB_HD float clip_value(float v, int& clipped) {
    if (v > .99999994f) { clipped = 1; return 0.99999994f; }
    if (v < -.99999994f) { clipped = 1; return -0.99999994f; }
    return v;
}

// ------------------------------------------------------------------------------------------------
// Inverse MDCT, exact path: the reference's stb_vorbis dataflow (Mdct.cs:65-313) cut into
// data-parallel steps.  u = buffer[n] (spectrum in the first n/2), v = buf2[n/2].  Each function
// executes work items t, t+nt, ... of its step; the caller puts a barrier between steps.
// Butterfly bookkeeping: the four loop shapes of step 3 (iter0 315-359, inner_r 361-410,
// inner_s 412-461) are the same radix-2 pass with twiddle index b << (l+3):
//     e0 = n2-1 - k0*i - 2b,  e2 = e0 - k0/2,  k0 = n >> (l+2),  i < 2^(l+1),  b < 4*(n >> (l+6))
// ------------------------------------------------------------------------------------------------
NVB_HD void mdct_step0(const float* u, float* v, const float* A, int n, int t, int nt) {        // Mdct.cs:74-97
    int n2 = n >> 1, n4 = n >> 2, n8 = n >> 3;
    for (int w = t; w < n4; w += nt) {
        if (w < n8) {
            int d = n2 - 2 - 2 * w, AA = 2 * w, e = 4 * w;
            v[d + 1] = NVB_FSUB(NVB_FMUL(u[e], A[AA]), NVB_FMUL(u[e + 2], A[AA + 1]));
            v[d]     = NVB_FADD(NVB_FMUL(u[e], A[AA + 1]), NVB_FMUL(u[e + 2], A[AA]));
        } else {
            int q = w - n8;
            int d = n4 - 2 - 2 * q, AA = n4 + 2 * q, e = n2 - 3 - 4 * q;
            float x = -u[e + 2], y = -u[e];
            v[d + 1] = NVB_FSUB(NVB_FMUL(x, A[AA]), NVB_FMUL(y, A[AA + 1]));
            v[d]     = NVB_FADD(NVB_FMUL(x, A[AA + 1]), NVB_FMUL(y, A[AA]));
        }
    }
}

NVB_HD void mdct_step2(float* u, const float* v, const float* A, int n, int t, int nt) {        // Mdct.cs:105-139
    int n2 = n >> 1, n4 = n >> 2;
    int iters = n2 >> 3;
    for (int w = t; w < iters; w += nt) {
        int AA = n2 - 8 - 8 * w, e0 = n4 + 4 * w, e1 = 4 * w, d0 = e0, d1 = e1;
        float v41_21 = NVB_FSUB(v[e0 + 1], v[e1 + 1]);
        float v40_20 = NVB_FSUB(v[e0], v[e1]);
        u[d0 + 1] = NVB_FADD(v[e0 + 1], v[e1 + 1]);
        u[d0]     = NVB_FADD(v[e0], v[e1]);
        u[d1 + 1] = NVB_FSUB(NVB_FMUL(v41_21, A[AA + 4]), NVB_FMUL(v40_20, A[AA + 5]));
        u[d1]     = NVB_FADD(NVB_FMUL(v40_20, A[AA + 4]), NVB_FMUL(v41_21, A[AA + 5]));
        v41_21 = NVB_FSUB(v[e0 + 3], v[e1 + 3]);
        v40_20 = NVB_FSUB(v[e0 + 2], v[e1 + 2]);
        u[d0 + 3] = NVB_FADD(v[e0 + 3], v[e1 + 3]);
        u[d0 + 2] = NVB_FADD(v[e0 + 2], v[e1 + 2]);
        u[d1 + 3] = NVB_FSUB(NVB_FMUL(v41_21, A[AA]), NVB_FMUL(v40_20, A[AA + 1]));
        u[d1 + 2] = NVB_FADD(NVB_FMUL(v40_20, A[AA]), NVB_FMUL(v41_21, A[AA + 1]));
    }
}

NVB_HD int mdct_num_r2_passes(int n) {          // passes handled by mdct_step3_pass: l = 0 .. this-1
    int ld = ilog_u(n) - 1;                      // Mdct.cs:37
    int last = ld - 7;                           // loops 155-183 run l = 2 .. ld-7
    return (last < 1 ? 1 : last) + 1;            // iterations 0 and 1 are unconditional (144-151)
}

NVB_HD void mdct_step3_pass(float* e, const float* A, int n, int l, int t, int nt) {            // Mdct.cs:144-183
    int n2 = n >> 1;
    int k0 = n >> (l + 2), k0_2 = k0 >> 1;
    int lim = 1 << (l + 1);
    int nb = 4 * (n >> (l + 6));                 // butterflies per call (loops run in groups of four)
    int total = lim * nb;
    for (int w = t; w < total; w += nt) {
        int i = w / nb, b = w - i * nb;
        int e0 = n2 - 1 - k0 * i - 2 * b, e2 = e0 - k0_2;
        int a = b << (l + 3);
        float k00 = NVB_FSUB(e[e0], e[e2]);
        float k01 = NVB_FSUB(e[e0 - 1], e[e2 - 1]);
        e[e0]     = NVB_FADD(e[e0], e[e2]);
        e[e0 - 1] = NVB_FADD(e[e0 - 1], e[e2 - 1]);
        e[e2]     = NVB_FSUB(NVB_FMUL(k00, A[a]), NVB_FMUL(k01, A[a + 1]));
        e[e2 - 1] = NVB_FADD(NVB_FMUL(k01, A[a]), NVB_FMUL(k00, A[a + 1]));
    }
}

NVB_HD void mdct_iter54(float* e, int z) {                                                       // Mdct.cs:509-535
    float k00 = NVB_FSUB(e[z], e[z - 4]);
    float y0 = NVB_FADD(e[z], e[z - 4]);
    float y2 = NVB_FADD(e[z - 2], e[z - 6]);
    float k22 = NVB_FSUB(e[z - 2], e[z - 6]);
    e[z] = NVB_FADD(y0, y2); e[z - 2] = NVB_FSUB(y0, y2);
    float k33 = NVB_FSUB(e[z - 3], e[z - 7]);
    e[z - 4] = NVB_FADD(k00, k33); e[z - 6] = NVB_FSUB(k00, k33);
    float k11 = NVB_FSUB(e[z - 1], e[z - 5]);
    float y1 = NVB_FADD(e[z - 1], e[z - 5]);
    float y3 = NVB_FADD(e[z - 3], e[z - 7]);
    e[z - 1] = NVB_FADD(y1, y3); e[z - 3] = NVB_FSUB(y1, y3);
    e[z - 5] = NVB_FSUB(k11, k22); e[z - 7] = NVB_FADD(k11, k22);
}

NVB_HD void mdct_ld654(float* e, const float* A, int n, int t, int nt) {                         // Mdct.cs:463-507
    int n2 = n >> 1;
    float A2 = A[n >> 3];
    int groups = n >> 5;
    for (int w = t; w < groups; w += nt) {
        int z = n2 - 1 - 16 * w;
        float k00, k11;
        k00 = NVB_FSUB(e[z], e[z - 8]); k11 = NVB_FSUB(e[z - 1], e[z - 9]);
        e[z] = NVB_FADD(e[z], e[z - 8]); e[z - 1] = NVB_FADD(e[z - 1], e[z - 9]);
        e[z - 8] = k00; e[z - 9] = k11;
        k00 = NVB_FSUB(e[z - 2], e[z - 10]); k11 = NVB_FSUB(e[z - 3], e[z - 11]);
        e[z - 2] = NVB_FADD(e[z - 2], e[z - 10]); e[z - 3] = NVB_FADD(e[z - 3], e[z - 11]);
        e[z - 10] = NVB_FMUL(NVB_FADD(k00, k11), A2); e[z - 11] = NVB_FMUL(NVB_FSUB(k11, k00), A2);
        k00 = NVB_FSUB(e[z - 12], e[z - 4]); k11 = NVB_FSUB(e[z - 5], e[z - 13]);
        e[z - 4] = NVB_FADD(e[z - 4], e[z - 12]); e[z - 5] = NVB_FADD(e[z - 5], e[z - 13]);
        e[z - 12] = k11; e[z - 13] = k00;
        k00 = NVB_FSUB(e[z - 14], e[z - 6]); k11 = NVB_FSUB(e[z - 7], e[z - 15]);
        e[z - 6] = NVB_FADD(e[z - 6], e[z - 14]); e[z - 7] = NVB_FADD(e[z - 7], e[z - 15]);
        e[z - 14] = NVB_FMUL(NVB_FADD(k00, k11), A2); e[z - 15] = NVB_FMUL(NVB_FSUB(k00, k11), A2);
        mdct_iter54(e, z);
        mdct_iter54(e, z - 8);
    }
}

NVB_HD void mdct_step456(const float* u, float* v, const uint16_t* bitrev, int n, int t, int nt) {  // Mdct.cs:189-214
    int n2 = n >> 1, n4 = n >> 2;
    int iters = n4 >> 2;
    for (int w = t; w < iters; w += nt) {
        int bit = 2 * w, d0 = n4 - 4 - 4 * w, d1 = n2 - 4 - 4 * w;
        int k4 = bitrev[bit];
        v[d1 + 3] = u[k4]; v[d1 + 2] = u[k4 + 1]; v[d0 + 3] = u[k4 + 2]; v[d0 + 2] = u[k4 + 3];
        k4 = bitrev[bit + 1];
        v[d1 + 1] = u[k4]; v[d1] = u[k4 + 1]; v[d0 + 1] = u[k4 + 2]; v[d0] = u[k4 + 3];
    }
}

NVB_HD void mdct_step7(float* v, const float* C, int n, int t, int nt) {                         // Mdct.cs:217-258
    int n2 = n >> 1;
    int iters = n2 >> 3;                          // d = 4w < e = n2-4-4w
    for (int w = t; w < iters; w += nt) {
        int c = 4 * w, d = 4 * w, e = n2 - 4 - 4 * w;
        if (!(d < e)) continue;
        float a02, a11, b0, b1, b2, b3;
        a02 = NVB_FSUB(v[d], v[e + 2]); a11 = NVB_FADD(v[d + 1], v[e + 3]);
        b0 = NVB_FADD(NVB_FMUL(C[c + 1], a02), NVB_FMUL(C[c], a11));
        b1 = NVB_FSUB(NVB_FMUL(C[c + 1], a11), NVB_FMUL(C[c], a02));
        b2 = NVB_FADD(v[d], v[e + 2]); b3 = NVB_FSUB(v[d + 1], v[e + 3]);
        v[d] = NVB_FADD(b2, b0); v[d + 1] = NVB_FADD(b3, b1); v[e + 2] = NVB_FSUB(b2, b0); v[e + 3] = NVB_FSUB(b1, b3);
        a02 = NVB_FSUB(v[d + 2], v[e]); a11 = NVB_FADD(v[d + 3], v[e + 1]);
        b0 = NVB_FADD(NVB_FMUL(C[c + 3], a02), NVB_FMUL(C[c + 2], a11));
        b1 = NVB_FSUB(NVB_FMUL(C[c + 3], a11), NVB_FMUL(C[c + 2], a02));
        b2 = NVB_FADD(v[d + 2], v[e]); b3 = NVB_FSUB(v[d + 3], v[e + 1]);
        v[d + 2] = NVB_FADD(b2, b0); v[d + 3] = NVB_FADD(b3, b1); v[e] = NVB_FSUB(b2, b0); v[e + 1] = NVB_FSUB(b1, b3);
    }
}

NVB_HD void mdct_step8(float* u, const float* v, const float* B, int n, int t, int nt) {         // Mdct.cs:261-312
    int n2 = n >> 1;
    int iters = n2 >> 3;
    for (int w = t; w < iters; w += nt) {
        int b = n2 - 8 - 8 * w, e = b, d0 = 4 * w, d1 = n2 - 4 - 4 * w, d2 = n2 + 4 * w, d3 = n - 4 - 4 * w;
        float p0, p1, p2, p3;
        p3 = NVB_FSUB(NVB_FMUL(v[e + 6], B[b + 7]), NVB_FMUL(v[e + 7], B[b + 6]));
        p2 = NVB_FSUB(NVB_FMUL(-v[e + 6], B[b + 6]), NVB_FMUL(v[e + 7], B[b + 7]));
        u[d0] = p3; u[d1 + 3] = -p3; u[d2] = p2; u[d3 + 3] = p2;
        p1 = NVB_FSUB(NVB_FMUL(v[e + 4], B[b + 5]), NVB_FMUL(v[e + 5], B[b + 4]));
        p0 = NVB_FSUB(NVB_FMUL(-v[e + 4], B[b + 4]), NVB_FMUL(v[e + 5], B[b + 5]));
        u[d0 + 1] = p1; u[d1 + 2] = -p1; u[d2 + 1] = p0; u[d3 + 2] = p0;
        p3 = NVB_FSUB(NVB_FMUL(v[e + 2], B[b + 3]), NVB_FMUL(v[e + 3], B[b + 2]));
        p2 = NVB_FSUB(NVB_FMUL(-v[e + 2], B[b + 2]), NVB_FMUL(v[e + 3], B[b + 3]));
        u[d0 + 2] = p3; u[d1 + 1] = -p3; u[d2 + 2] = p2; u[d3 + 1] = p2;
        p1 = NVB_FSUB(NVB_FMUL(v[e], B[b + 1]), NVB_FMUL(v[e + 1], B[b]));
        p0 = NVB_FSUB(NVB_FMUL(-v[e], B[b]), NVB_FMUL(v[e + 1], B[b + 1]));
        u[d0 + 3] = p1; u[d1] = -p1; u[d2 + 3] = p0; u[d3] = p0;
    }
}

// Window of one frame (Mode.cs:44-50,159-166)
NVB_HD const float* frame_window(const DevSetup& S, const DevFrame& f) {
    return S.modes[f.mode].block_flag ? S.win_long + (size_t)f.window * S.bs[1] : S.win_short;
}

}  // namespace nvb
