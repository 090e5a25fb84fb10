// Per-block / per-macroblock arithmetic of the B200 H.264 encoder.
//
// Everything here is written once and compiled twice: as device code inside the CUDA kernels
// (kernels.cuh) and inside the CPU unit-test harness (host_harness.cpp, which exists so that
// logic bugs are found without a GPU; it is not a product path).  This replaces the fixed-function
// Cedar VE macroblock pipeline that the reference starts at kernel/cedar.c:1176; the reference has
// no source for it.  Bit-exactness target: oracle/h264_golden.c.
#pragma once
#include <stddef.h>
#include <stdint.h>

// Under nvcc the functions are device-only and the tables live in device memory; under a plain
// C++ compiler (the CPU test harness) they are ordinary inline functions and host arrays.
#ifdef __CUDACC__
#define HD __device__ __forceinline__
#define H264_TABLE static __device__ const
#else
#define HD inline
#define H264_TABLE static const __attribute__((unused))
#endif
#include "h264_tables.h"

namespace cedar {

enum { MB_I16x16 = 0, MB_P16x16 = 1, MB_PSKIP = 2, MB_I4x4 = 3 };

// One record per macroblock (16 bytes).  mv / mvd are quarter-pel.
struct MbInfo {
    uint8_t type, i16_mode, chroma_mode, cbp; // cbp = luma | chroma << 4
    int16_t mv[2];
    int16_t mvd[2];
    uint32_t pad;
};
static_assert(sizeof(MbInfo) == 16, "MbInfo layout");

// nnz record: 32 bytes per MB. 0..15 luma (luma4x4BlkIdx order; AC count for I16x16), 16 I16 DC,
// 17..20 Cb AC, 21..24 Cr AC, 25 Cb DC, 26 Cr DC.
enum { NNZ_STRIDE = 32, NNZ_DC16 = 16, NNZ_CB = 17, NNZ_CR = 21, NNZ_CBDC = 25, NNZ_CRDC = 26 };
// coefficient record: 26 blocks x 16 int16 levels in zig-zag order. 0..15 luma, 16 I16 DC,
// 17 chroma DC (Cb 0..3, Cr 4..7), 18..21 Cb AC, 22..25 Cr AC (index 0 unused in AC blocks).
enum { COEF_BLOCKS = 26, COEF_STRIDE = 26 * 16 };

struct Geom {
    int W, H, CW, CH;   // coded luma / chroma plane sizes
    int mbw, mbh, nmb;
    int src_w, src_h, src_format;
    int qp, qpc;
    int R, lambda;      // ME radius and SAD lambda
    int cabac;
    int srows, nslices; // macroblock rows per slice (mbh = the reference's one slice per picture), slices per picture
    int intra4x4;       // extension: Intra4x4 macroblocks in I frames (off: Intra16x16 only)
    int p_intra;        // extension: Intra16x16 macroblocks inside P frames where they beat the motion search
    unsigned long long frame_bytes; // W*H*3/2 : one planar frame (Y, U, V)
};

HD int iabs_(int v) { return v < 0 ? -v : v; }
HD int imin_(int a, int b) { return a < b ? a : b; }
HD int imax_(int a, int b) { return a > b ? a : b; }
HD int clip3_(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
HD int clip255_(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

HD int blk_x(int b) { return (b & 1) | ((b >> 1) & 2); }
HD int blk_y(int b) { return ((b >> 1) & 1) | ((b >> 2) & 2); }
HD int xy2blk(int x, int y) { return (x & 1) | ((y & 1) << 1) | ((x & 2) << 1) | ((y & 2) << 2); }

HD int pos_class(int r) // raster index -> 0 (a), 1 (b), 2 (c)
{
    int odd_r = (r >> 2) & 1, odd_c = r & 1;
    return (odd_r & odd_c) ? 1 : ((odd_r | odd_c) ? 2 : 0);
}

// ---- 4x4 forward core transform of a residual block d (raster) ------------------------------
HD void fdct4x4(const int *d, int *w)
{
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int a = d[i * 4], b = d[i * 4 + 1], c = d[i * 4 + 2], e = d[i * 4 + 3];
        int s03 = a + e, d03 = a - e, s12 = b + c, d12 = b - c;
        t[i * 4 + 0] = s03 + s12;
        t[i * 4 + 1] = 2 * d03 + d12;
        t[i * 4 + 2] = s03 - s12;
        t[i * 4 + 3] = d03 - 2 * d12;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int a = t[i], b = t[4 + i], c = t[8 + i], e = t[12 + i];
        int s03 = a + e, d03 = a - e, s12 = b + c, d12 = b - c;
        w[i] = s03 + s12;
        w[4 + i] = 2 * d03 + d12;
        w[8 + i] = s03 - s12;
        w[12 + i] = d03 - 2 * d12;
    }
}

// ---- inverse transform of scaled coefficients d (raster), residual r = (x + 32) >> 6 ---------
HD void idct4x4(const int *d, int *r)
{
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int d0 = d[i * 4], d1 = d[i * 4 + 1], d2 = d[i * 4 + 2], d3 = d[i * 4 + 3];
        int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
        t[i * 4 + 0] = e0 + e3;
        t[i * 4 + 1] = e1 + e2;
        t[i * 4 + 2] = e1 - e2;
        t[i * 4 + 3] = e0 - e3;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int d0 = t[i], d1 = t[4 + i], d2 = t[8 + i], d3 = t[12 + i];
        int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
        r[i] = (e0 + e3 + 32) >> 6;
        r[4 + i] = (e1 + e2 + 32) >> 6;
        r[8 + i] = (e1 - e2 + 32) >> 6;
        r[12 + i] = (e0 - e3 + 32) >> 6;
    }
}

HD int quant1(int w, int mf, int f, int shift)
{
    int a = (iabs_(w) * mf + f) >> shift;
    return w < 0 ? -a : a;
}

HD int dequant_ac(int c, int qp, int cls)
{
    int ls = 16 * h264_dequant_v[qp % 6][cls];
    if (qp >= 24)
        return (c * ls) * (1 << (qp / 6 - 4));
    return (c * ls + (1 << (3 - qp / 6))) >> (4 - qp / 6);
}

// Quantise transformed block w (raster) into zig-zag levels lev[first..15]; returns nnz.
HD int quant_block(const int *w, int qp, int intra, int first, int16_t *lev)
{
    int qbits = 15 + qp / 6;
    int f = (1 << qbits) / (intra ? 3 : 6);
    int nnz = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        int z = 0;
        if (i >= first) {
            int r = h264_zigzag4x4[i];
            z = quant1(w[r], h264_quant_mf[qp % 6][pos_class(r)], f, qbits);
        }
        lev[i] = (int16_t)z;
        nnz += z != 0;
    }
    return nnz;
}

HD void dequant_block(const int16_t *lev, int qp, int first, int *d)
{
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (i >= first) {
            int r = h264_zigzag4x4[i];
            d[r] = dequant_ac(lev[i], qp, pos_class(r));
        }
    }
}

HD int dequant_luma_dc(int f, int qp)
{
    int ls = 16 * h264_dequant_v[qp % 6][0];
    if (qp >= 36)
        return (f * ls) * (1 << (qp / 6 - 6));
    return (f * ls + (1 << (5 - qp / 6))) >> (6 - qp / 6);
}

HD int dequant_chroma_dc(int f, int qpc)
{
    int ls = 16 * h264_dequant_v[qpc % 6][0];
    return ((f * ls) * (1 << (qpc / 6))) >> 5;
}

// 4x4 Hadamard (unscaled) of a raster 4x4 array
HD void hadamard4x4(const int *in, int *out)
{
    int t[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int a = in[i * 4], b = in[i * 4 + 1], c = in[i * 4 + 2], d = in[i * 4 + 3];
        t[i * 4 + 0] = a + b + c + d;
        t[i * 4 + 1] = a + b - c - d;
        t[i * 4 + 2] = a - b - c + d;
        t[i * 4 + 3] = a - b + c - d;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int a = t[i], b = t[4 + i], c = t[8 + i], d = t[12 + i];
        out[i] = a + b + c + d;
        out[4 + i] = a + b - c - d;
        out[8 + i] = a - b - c + d;
        out[12 + i] = a - b + c - d;
    }
}

// element i (0..3) of the 2x2 Hadamard of (d0 d1 / d2 d3)
HD int hadamard2x2_elem(int i, int d0, int d1, int d2, int d3)
{
    int s1 = (i & 1) ? -1 : 1, s2 = (i & 2) ? -1 : 1;
    return d0 + s1 * d1 + s2 * d2 + s1 * s2 * d3;
}

// ---- motion-vector cost (identical integers in the oracle) -----------------------------------
HD int ilog2_(unsigned v)
{
#ifdef __CUDACC__
    return 31 - __clz(v);
#else
    return 31 - __builtin_clz(v);
#endif
}
HD int mv_bits(int d) { return d == 0 ? 1 : 7 + 2 * ilog2_((unsigned)iabs_(d)); }
HD int me_lambda(int qp) { return 1 << clip3_(0, 5, (qp - 12) / 6); }

HD int median3(int a, int b, int c) { return imax_(imin_(a, b), imin_(imax_(a, b), c)); }

// ---- median MV prediction for 16x16 partitions + P_Skip MV (H.264 8.4.1.1 / 8.4.1.3) ---------
// mbi: macroblock records of the frame (types are only tested for intra / inter).
// srows = macroblock rows per slice: the row above is unavailable when it belongs to another slice.
HD void predict_mv(const MbInfo *mbi, int mbw, int srows, int mbx, int mby, int *mvp, int *skip_mv)
{
    int availA = mbx > 0, availB = (mby % srows) != 0;
    int availC = availB && mbx + 1 < mbw, availD = availB && mbx > 0;
    int a[2] = {0, 0}, b[2] = {0, 0}, c[2] = {0, 0};
    int ra = -1, rb = -1, rc = -1;
    const MbInfo *cur = mbi + mby * mbw + mbx;
    if (availA) {
        const MbInfo *n = cur - 1;
        if (n->type == MB_P16x16 || n->type == MB_PSKIP) {
            a[0] = n->mv[0];
            a[1] = n->mv[1];
            ra = 0;
        }
    }
    if (availB) {
        const MbInfo *n = cur - mbw;
        if (n->type == MB_P16x16 || n->type == MB_PSKIP) {
            b[0] = n->mv[0];
            b[1] = n->mv[1];
            rb = 0;
        }
    }
    if (availC || availD) {
        const MbInfo *n = availC ? cur - mbw + 1 : cur - mbw - 1;
        if (n->type == MB_P16x16 || n->type == MB_PSKIP) {
            c[0] = n->mv[0];
            c[1] = n->mv[1];
            rc = 0;
        }
    }
    int a0x = a[0], a0y = a[1], b0x = b[0], b0y = b[1], ra0 = ra, rb0 = rb;
    if (!availB && !(availC || availD) && availA) {
        b[0] = c[0] = a[0];
        b[1] = c[1] = a[1];
        rb = rc = ra;
    }
    int match = (ra == 0) + (rb == 0) + (rc == 0);
    if (match == 1) {
        const int *m = ra == 0 ? a : (rb == 0 ? b : c);
        mvp[0] = m[0];
        mvp[1] = m[1];
    } else {
        mvp[0] = median3(a[0], b[0], c[0]);
        mvp[1] = median3(a[1], b[1], c[1]);
    }
    if (!availA || !availB || (ra0 == 0 && a0x == 0 && a0y == 0) || (rb0 == 0 && b0x == 0 && b0y == 0)) {
        skip_mv[0] = skip_mv[1] = 0;
    } else {
        skip_mv[0] = mvp[0];
        skip_mv[1] = mvp[1];
    }
}

// ---- deblocking (H.264 8.7) ---------------------------------------------------------------
HD int is_intra(int type) { return type == MB_I16x16 || type == MB_I4x4; }

HD int boundary_strength(const MbInfo &mp, int nnzp, const MbInfo &mq, int nnzq, int mb_edge)
{
    if (is_intra(mp.type) || is_intra(mq.type))
        return mb_edge ? 4 : 3;
    if (nnzp || nnzq)
        return 2;
    if (iabs_(mp.mv[0] - mq.mv[0]) >= 4 || iabs_(mp.mv[1] - mq.mv[1]) >= 4)
        return 1;
    return 0;
}

// Luma edge filter on 8 samples p3 p2 p1 p0 | q0 q1 q2 q3 (in place in v[0..7]).
HD void filter_luma8(int *v, int bS, int alpha, int beta, int tc0)
{
    int p3 = v[0], p2 = v[1], p1 = v[2], p0 = v[3], q0 = v[4], q1 = v[5], q2 = v[6], q3 = v[7];
    if (iabs_(p0 - q0) >= alpha || iabs_(p1 - p0) >= beta || iabs_(q1 - q0) >= beta)
        return;
    int ap = iabs_(p2 - p0), aq = iabs_(q2 - q0);
    if (bS < 4) {
        int tc = tc0 + (ap < beta) + (aq < beta);
        int delta = clip3_(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
        v[3] = clip255_(p0 + delta);
        v[4] = clip255_(q0 - delta);
        if (ap < beta)
            v[2] = p1 + clip3_(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 * 2)) >> 1);
        if (aq < beta)
            v[5] = q1 + clip3_(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 * 2)) >> 1);
    } else {
        int strong = iabs_(p0 - q0) < ((alpha >> 2) + 2);
        if (ap < beta && strong) {
            v[3] = (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3;
            v[2] = (p2 + p1 + p0 + q0 + 2) >> 2;
            v[1] = (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3;
        } else
            v[3] = (2 * p1 + p0 + q1 + 2) >> 2;
        if (aq < beta && strong) {
            v[4] = (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3;
            v[5] = (p0 + q0 + q1 + q2 + 2) >> 2;
            v[6] = (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3;
        } else
            v[4] = (2 * q1 + q0 + p1 + 2) >> 2;
    }
}

// Chroma edge filter on p1 p0 | q0 q1 (in place in v[0..3]).
HD void filter_chroma4(int *v, int bS, int alpha, int beta, int tc0)
{
    int p1 = v[0], p0 = v[1], q0 = v[2], q1 = v[3];
    if (iabs_(p0 - q0) >= alpha || iabs_(p1 - p0) >= beta || iabs_(q1 - q0) >= beta)
        return;
    if (bS < 4) {
        int tc = tc0 + 1;
        int delta = clip3_(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
        v[1] = clip255_(p0 + delta);
        v[2] = clip255_(q0 - delta);
    } else {
        v[1] = (2 * p1 + p0 + q1 + 2) >> 2;
        v[2] = (2 * q1 + q0 + p1 + 2) >> 2;
    }
}

// ---- intra prediction of one 4x4 sub-block of a 16x16 / 8x8 predicted macroblock --------------
// top[0] = top-left, top[1..N] = row above, left[0..N-1] = column to the left.
// Intra16x16: mode 0 V, 1 H, 2 DC, 3 Plane.  (bx, by) = pixel offset of the 4x4 block.
HD void pred16_block(int mode, const uint8_t *top, const uint8_t *left, int has_top, int has_left, int bx, int by,
                     int *pred)
{
    if (mode == 0) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            pred[i] = top[1 + bx + (i & 3)];
    } else if (mode == 1) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            pred[i] = left[by + (i >> 2)];
    } else if (mode == 2) {
        int s = 0, dc;
        if (has_top)
            for (int i = 0; i < 16; i++)
                s += top[1 + i];
        if (has_left)
            for (int i = 0; i < 16; i++)
                s += left[i];
        if (has_top && has_left)
            dc = (s + 16) >> 5;
        else if (has_top || has_left)
            dc = (s + 8) >> 4;
        else
            dc = 128;
#pragma unroll
        for (int i = 0; i < 16; i++)
            pred[i] = dc;
    } else {
        int Hh = 0, Vv = 0;
        for (int i = 0; i < 8; i++) {
            Hh += (i + 1) * (top[1 + 8 + i] - top[1 + 6 - i]); // top[0] is the top-left sample (i == 7)
            Vv += (i + 1) * (left[8 + i] - (i == 7 ? top[0] : left[6 - i]));
        }
        int a = 16 * (left[15] + top[16]);
        int b = (5 * Hh + 32) >> 6, c = (5 * Vv + 32) >> 6;
#pragma unroll
        for (int i = 0; i < 16; i++)
            pred[i] = clip255_((a + b * (bx + (i & 3) - 7) + c * (by + (i >> 2) - 7) + 16) >> 5);
    }
}

// One sample (x, y) of an Intra4x4 prediction (H.264 8.3.1.2): t[0..7] = A..H (above, above-right), l[0..3] = I..L
// (left), m = M (above-left).  Modes: 0 V, 1 H, 2 DC, 3 DDL, 4 DDR, 5 VR, 6 HD, 7 VL, 8 HU.
HD int pred4x4_pixel(int mode, int x, int y, const uint8_t *t, const uint8_t *l, int m, int has_top, int has_left)
{
#define PT_(i) ((i) < 0 ? m : (int)t[i])
#define PL_(i) ((i) < 0 ? m : (int)l[i])
    switch (mode) {
    case 0: return t[x];
    case 1: return l[y];
    case 2: {
        int s = 0;
        if (has_top)
            s += t[0] + t[1] + t[2] + t[3];
        if (has_left)
            s += l[0] + l[1] + l[2] + l[3];
        return (has_top && has_left) ? (s + 4) >> 3 : ((has_top || has_left) ? (s + 2) >> 2 : 128);
    }
    case 3: return (x == 3 && y == 3) ? (t[6] + 3 * t[7] + 2) >> 2 : (t[x + y] + 2 * t[x + y + 1] + t[x + y + 2] + 2) >> 2;
    case 4:
        if (x > y)
            return (PT_(x - y - 2) + 2 * PT_(x - y - 1) + PT_(x - y) + 2) >> 2;
        if (x < y)
            return (PL_(y - x - 2) + 2 * PL_(y - x - 1) + PL_(y - x) + 2) >> 2;
        return (t[0] + 2 * m + l[0] + 2) >> 2;
    case 5: {
        int z = 2 * x - y;
        if (z >= 0 && !(z & 1))
            return (PT_(x - (y >> 1) - 1) + PT_(x - (y >> 1)) + 1) >> 1;
        if (z >= 0)
            return (PT_(x - (y >> 1) - 2) + 2 * PT_(x - (y >> 1) - 1) + PT_(x - (y >> 1)) + 2) >> 2;
        if (z == -1)
            return (l[0] + 2 * m + t[0] + 2) >> 2;
        return (PL_(y - 1) + 2 * PL_(y - 2) + PL_(y - 3) + 2) >> 2;
    }
    case 6: {
        int z = 2 * y - x;
        if (z >= 0 && !(z & 1))
            return (PL_(y - (x >> 1) - 1) + PL_(y - (x >> 1)) + 1) >> 1;
        if (z >= 0)
            return (PL_(y - (x >> 1) - 2) + 2 * PL_(y - (x >> 1) - 1) + PL_(y - (x >> 1)) + 2) >> 2;
        if (z == -1)
            return (l[0] + 2 * m + t[0] + 2) >> 2;
        return (PT_(x - 1) + 2 * PT_(x - 2) + PT_(x - 3) + 2) >> 2;
    }
    case 7:
        return !(y & 1) ? (t[x + (y >> 1)] + t[x + (y >> 1) + 1] + 1) >> 1
                        : (t[x + (y >> 1)] + 2 * t[x + (y >> 1) + 1] + t[x + (y >> 1) + 2] + 2) >> 2;
    default: {
        int z = x + 2 * y;
        if (z > 5)
            return l[3];
        if (z == 5)
            return (l[2] + 3 * l[3] + 2) >> 2;
        if (!(z & 1))
            return (l[y + (x >> 1)] + l[y + (x >> 1) + 1] + 1) >> 1;
        return (l[y + (x >> 1)] + 2 * l[y + (x >> 1) + 1] + l[y + (x >> 1) + 2] + 2) >> 2;
    }
    }
#undef PT_
#undef PL_
}

// Chroma 8x8: mode 0 DC, 1 H, 2 V, 3 Plane.
HD void predc_block(int mode, const uint8_t *top, const uint8_t *left, int has_top, int has_left, int bx, int by,
                    int *pred)
{
    if (mode == 0) {
        int st = 0, sl = 0, dc;
        for (int i = 0; i < 4; i++) {
            st += has_top ? top[1 + bx + i] : 0;
            sl += has_left ? left[by + i] : 0;
        }
        int use_t = has_top, use_l = has_left;
        if (bx == 4 && by == 0 && has_top)
            use_l = 0;
        if (bx == 0 && by == 4 && has_left)
            use_t = 0;
        if (use_t && use_l)
            dc = (st + sl + 4) >> 3;
        else if (use_t)
            dc = (st + 2) >> 2;
        else if (use_l)
            dc = (sl + 2) >> 2;
        else
            dc = 128;
#pragma unroll
        for (int i = 0; i < 16; i++)
            pred[i] = dc;
    } else if (mode == 1) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            pred[i] = left[by + (i >> 2)];
    } else if (mode == 2) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            pred[i] = top[1 + bx + (i & 3)];
    } else {
        int Hh = 0, Vv = 0;
        for (int i = 0; i < 4; i++) {
            Hh += (i + 1) * (top[1 + 4 + i] - top[1 + 2 - i]);
            Vv += (i + 1) * (left[4 + i] - (i == 3 ? top[0] : left[2 - i]));
        }
        int a = 16 * (left[7] + top[8]);
        int b = (34 * Hh + 32) >> 6, c = (34 * Vv + 32) >> 6;
#pragma unroll
        for (int i = 0; i < 16; i++)
            pred[i] = clip255_((a + b * (bx + (i & 3) - 3) + c * (by + (i >> 2) - 3) + 16) >> 5);
    }
}

} // namespace cedar
