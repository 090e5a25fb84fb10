// Entropy coding of one macroblock: CAVLC bit strings and CABAC binarisation, plus the serial
// CABAC arithmetic coder.  Written once, compiled as device code (kernels.cuh) and as host code
// (host_harness.cpp, CPU unit tests only).  Replaces the entropy stage of the Cedar VE that the
// reference selects with PARA0 bit 8 (kernel/cedar.c:1155-1158) and never shows in source.
//
// Every function takes a "sink": a counting sink (pass 1: sizes, then a prefix sum gives each
// macroblock its offset) or a writing sink (pass 2: scatter at that offset).  One code path for
// both passes keeps sizes and contents consistent by construction.
#pragma once
#include "h264_core.cuh"

namespace cedar {

struct FrameSyntax {
    const MbInfo *mbi;
    const uint8_t *nnz;   // [nmb][NNZ_STRIDE]
    const int16_t *coef;  // [nmb][COEF_STRIDE]
    int mbw, mbh;
    int srows; // macroblock rows per slice (== mbh: one slice per picture, the reference's layout)
    const uint8_t *i4; // [nmb][16] Intra4x4PredMode per luma4x4BlkIdx (only read for MB_I4x4 macroblocks; may be null)
};

// Slices are whole macroblock rows, so the left neighbour is always in the same slice; the row above is
// "not available" (H.264 6.4.x) for every entropy context when it belongs to another slice.
HD bool top_avail(const FrameSyntax &fs, int mby) { return (mby % fs.srows) != 0; }

// Entropy work items of a picture: per slice its macroblocks in raster order followed by one terminal item
// (CAVLC: trailing mb_skip_run + rbsp_slice_trailing_bits; CABAC: nothing).  All slices but the last have
// srows * mbw macroblocks, so item -> (slice, macroblock) needs one division.
struct SliceItem {
    int slice, mb, first_mb;
    bool is_end, is_first;
};
HD int slice_items_per(const FrameSyntax &fs) { return fs.srows * fs.mbw + 1; }
HD SliceItem slice_item(const FrameSyntax &fs, int item)
{
    SliceItem it;
    int per = slice_items_per(fs);
    it.slice = item / per;
    int local = item - it.slice * per;
    int row0 = it.slice * fs.srows;
    int count = imin_(fs.srows, fs.mbh - row0) * fs.mbw;
    it.first_mb = row0 * fs.mbw;
    it.mb = it.first_mb + local;
    it.is_end = local == count;
    it.is_first = local == 0;
    return it;
}

// total_coeff of the 4x4 block left of / above block `blk`; -1 when outside the picture.
// kind 0: luma (blk = luma4x4BlkIdx), 1: Cb AC, 2: Cr AC (blk 0..3 raster)
HD int nnz_left(const FrameSyntax &fs, int mbx, int mby, int kind, int blk)
{
    const uint8_t *cur = fs.nnz + (size_t)(mby * fs.mbw + mbx) * NNZ_STRIDE;
    if (kind == 0) {
        int bx = blk_x(blk), by = blk_y(blk);
        if (bx > 0)
            return cur[xy2blk(bx - 1, by)];
        if (mbx == 0)
            return -1;
        return (cur - NNZ_STRIDE)[xy2blk(3, by)];
    }
    int base = kind == 1 ? NNZ_CB : NNZ_CR, bx = blk & 1, by = blk >> 1;
    if (bx > 0)
        return cur[base + by * 2];
    if (mbx == 0)
        return -1;
    return (cur - NNZ_STRIDE)[base + by * 2 + 1];
}

HD int nnz_top(const FrameSyntax &fs, int mbx, int mby, int kind, int blk)
{
    const uint8_t *cur = fs.nnz + (size_t)(mby * fs.mbw + mbx) * NNZ_STRIDE;
    if (kind == 0) {
        int bx = blk_x(blk), by = blk_y(blk);
        if (by > 0)
            return cur[xy2blk(bx, by - 1)];
        if (!top_avail(fs, mby))
            return -1;
        return (cur - (size_t)fs.mbw * NNZ_STRIDE)[xy2blk(bx, 3)];
    }
    int base = kind == 1 ? NNZ_CB : NNZ_CR, bx = blk & 1, by = blk >> 1;
    if (by > 0)
        return cur[base + bx];
    if (!top_avail(fs, mby))
        return -1;
    return (cur - (size_t)fs.mbw * NNZ_STRIDE)[base + 2 + bx];
}

// Intra4x4PredMode of the 4x4 block left of / above block blk for the prediction of the mode (8.3.1.1 with
// constrained_intra_pred_flag = 0): -1 = not available, 2 (DC) when the neighbouring macroblock is not Intra4x4.
HD int i4_neighbour_mode(const FrameSyntax &fs, int mbx, int mby, int blk, int left)
{
    int mb = mby * fs.mbw + mbx, bx = blk_x(blk), by = blk_y(blk), nb;
    if (left) {
        if (bx > 0)
            nb = xy2blk(bx - 1, by);
        else {
            if (mbx == 0)
                return -1;
            mb -= 1;
            nb = xy2blk(3, by);
        }
    } else {
        if (by > 0)
            nb = xy2blk(bx, by - 1);
        else {
            if (!top_avail(fs, mby))
                return -1;
            mb -= fs.mbw;
            nb = xy2blk(bx, 3);
        }
    }
    return fs.mbi[mb].type == MB_I4x4 ? fs.i4[(size_t)mb * 16 + nb] : 2;
}
HD int i4_pred_mode(const FrameSyntax &fs, int mbx, int mby, int blk)
{
    int a = i4_neighbour_mode(fs, mbx, mby, blk, 1), b = i4_neighbour_mode(fs, mbx, mby, blk, 0);
    return (a < 0 || b < 0) ? 2 : imin_(a, b);
}

// =============================================================================================
// CAVLC
// =============================================================================================
struct BitCount {
    unsigned n = 0;
    HD void put(uint32_t, int len) { n += (unsigned)len; }
};

HD uint32_t bswap32_(uint32_t v)
{
#ifdef __CUDACC__
    return __byte_perm(v, 0, 0x0123);
#else
    return __builtin_bswap32(v);
#endif
}

// Writes MSB-first bits into a zero-initialised byte stream held as 32-bit words.  Words that may
// be shared with the neighbouring macroblock (first and last) are OR-ed atomically.
struct BitScatter {
    uint32_t *buf;
    unsigned long long acc = 0; // bits are left aligned in acc
    unsigned word;              // index of the 32-bit word acc's top bits belong to
    int n;                      // valid bits in acc
    bool first = true;
    HD BitScatter(uint32_t *b, unsigned long long bitpos) : buf(b), word((unsigned)(bitpos >> 5)), n((int)(bitpos & 31)) {}
    HD void store(uint32_t w, bool atomic)
    {
        if (!w && atomic)
            return;
        uint32_t v = bswap32_(w);
#ifdef __CUDACC__
        if (atomic)
            atomicOr(buf + word, v);
        else
            buf[word] = v;
#else
        if (atomic)
            buf[word] |= v;
        else
            buf[word] = v;
#endif
    }
    HD void put(uint32_t v, int len) // len <= 32
    {
        if (!len)
            return;
        acc |= (unsigned long long)v << (64 - n - len);
        n += len;
        if (n >= 32) {
            store((uint32_t)(acc >> 32), first);
            first = false;
            word++;
            acc <<= 32;
            n -= 32;
        }
    }
    HD void flush()
    {
        if (n > 0)
            store((uint32_t)(acc >> 32), true);
    }
};

template <class S> HD void put_ue(S &s, unsigned v)
{
    v++;
    s.put(v, (32 - (31 - ilog2_(v))) * 2 - 1);
}
template <class S> HD void put_se(S &s, int v)
{
    v = 2 * v - 1;
    v ^= (v >> 31);
    put_ue(s, (unsigned)v);
}

HD int cavlc_nc(const FrameSyntax &fs, int mbx, int mby, int kind, int blk)
{
    int a = nnz_left(fs, mbx, mby, kind, blk), b = nnz_top(fs, mbx, mby, kind, blk);
    if (a >= 0 && b >= 0)
        return (a + b + 1) >> 1;
    if (a >= 0)
        return a;
    if (b >= 0)
        return b;
    return 0;
}

// residual_block_cavlc of lev[0 .. max_coeff-1] (zig-zag order); nC < 0 selects the chroma DC tables
template <class S> HD void cavlc_block(S &s, const int16_t *lev, int max_coeff, int nC)
{
    int total = 0, last = -1, t1 = 0;
    bool t1_open = true;
    for (int i = max_coeff - 1; i >= 0; i--) {
        int v = lev[i];
        if (v) {
            if (last < 0)
                last = i;
            if (t1_open && t1 < 3 && (v == 1 || v == -1))
                t1++;
            else
                t1_open = false;
            total++;
        }
    }
    if (nC < 0)
        s.put(h264_chroma_dc_coeff_token_bits[4 * total + t1], h264_chroma_dc_coeff_token_len[4 * total + t1]);
    else {
        int tab = nC < 2 ? 0 : (nC < 4 ? 1 : (nC < 8 ? 2 : 3));
        s.put(h264_coeff_token_bits[tab][4 * total + t1], h264_coeff_token_len[tab][4 * total + t1]);
    }
    if (!total)
        return;
    // levels, highest frequency first
    int k = 0, suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
    for (int i = last; i >= 0; i--) {
        int level = lev[i];
        if (!level)
            continue;
        if (k < t1) {
            s.put(level < 0, 1);
            k++;
            continue;
        }
        int code = level > 0 ? 2 * level - 2 : -2 * level - 1;
        if (k == t1 && t1 < 3)
            code -= 2;
        int thresh = suffix_len == 0 ? 14 : (15 << suffix_len);
        if (code < thresh) {
            s.put(1, (code >> suffix_len) + 1);
            if (suffix_len)
                s.put((uint32_t)code & ((1u << suffix_len) - 1), suffix_len);
        } else if (suffix_len == 0 && code < 30) {
            s.put(1, 15);
            s.put((uint32_t)(code - 14), 4);
        } else {
            int c = code - (suffix_len == 0 ? 30 : (15 << suffix_len)), prefix = 15;
            while (c >= (1 << (prefix - 3))) {
                c -= 1 << (prefix - 3);
                prefix++;
            }
            s.put(1, prefix + 1);
            s.put((uint32_t)c, prefix - 3);
        }
        if (suffix_len == 0)
            suffix_len = 1;
        if (iabs_(level) > (3 << (suffix_len - 1)) && suffix_len < 6)
            suffix_len++;
        k++;
    }
    int zeros = last + 1 - total;
    if (total < max_coeff) {
        if (nC < 0)
            s.put(h264_chroma_dc_total_zeros_bits[total - 1][zeros], h264_chroma_dc_total_zeros_len[total - 1][zeros]);
        else
            s.put(h264_total_zeros_bits[total - 1][zeros], h264_total_zeros_len[total - 1][zeros]);
    }
    // run_before for every coefficient but the lowest-frequency one
    int zeros_left = zeros, i = last, done = 0;
    while (zeros_left > 0 && done < total - 1) {
        int run = 0, j = i - 1;
        while (j >= 0 && !lev[j]) {
            run++;
            j--;
        }
        int zl = imin_(zeros_left, 7) - 1;
        s.put(h264_run_before_bits[zl][run], h264_run_before_len[zl][run]);
        zeros_left -= run;
        i = j;
        done++;
    }
}

// Terminal item of a slice: trailing skip run + rbsp_slice_trailing_bits stop bit (alignment zeros come
// for free from the zero-initialised buffer).
template <class S> HD void cavlc_end(S &s, int skip_run)
{
    if (skip_run)
        put_ue(s, (unsigned)skip_run);
    s.put(1, 1);
}

// Number of P_Skip macroblocks directly before macroblock i inside its slice (which starts at first_mb).
HD int skip_run_before(const FrameSyntax &fs, int i, int first_mb)
{
    int run = 0;
    for (int j = i - 1; j >= first_mb && fs.mbi[j].type == MB_PSKIP; j--)
        run++;
    return run;
}

// One macroblock of slice data.  `skip_run` = number of P_Skip macroblocks directly before it.
template <class S> HD void cavlc_mb(S &s, const FrameSyntax &fs, int mb_index, int frame_i, int skip_run)
{
    const MbInfo &mb = fs.mbi[mb_index];
    if (mb.type == MB_PSKIP)
        return;
    int mbx = mb_index % fs.mbw, mby = mb_index / fs.mbw;
    const int16_t *coef = fs.coef + (size_t)mb_index * COEF_STRIDE;
    int cbpl = mb.cbp & 15, cbpc = mb.cbp >> 4;
    if (!frame_i)
        put_ue(s, (unsigned)skip_run);
    if (mb.type == MB_I16x16) {
        int t = 1 + mb.i16_mode + 4 * cbpc + (cbpl ? 12 : 0);
        put_ue(s, (unsigned)(t + (frame_i ? 0 : 5)));
        put_ue(s, mb.chroma_mode);
        put_se(s, 0); // mb_qp_delta
        cavlc_block(s, coef + 16 * 16, 16, cavlc_nc(fs, mbx, mby, 0, 0));
        if (cbpl)
            for (int b = 0; b < 16; b++)
                cavlc_block(s, coef + b * 16 + 1, 15, cavlc_nc(fs, mbx, mby, 0, b));
    } else if (mb.type == MB_I4x4) {
        put_ue(s, frame_i ? 0u : 5u); // I_NxN
        for (int b = 0; b < 16; b++) {
            int pm = i4_pred_mode(fs, mbx, mby, b), mode = fs.i4[(size_t)mb_index * 16 + b];
            if (mode == pm)
                s.put(1, 1); // prev_intra4x4_pred_mode_flag
            else
                s.put((uint32_t)(mode < pm ? mode : mode - 1), 4); // flag 0 + rem_intra4x4_pred_mode
        }
        put_ue(s, mb.chroma_mode);
        put_ue(s, h264_cbp_to_codenum_intra[mb.cbp]);
        if (mb.cbp)
            put_se(s, 0); // mb_qp_delta
        for (int b = 0; b < 16; b++)
            if (cbpl & (1 << (b >> 2)))
                cavlc_block(s, coef + b * 16, 16, cavlc_nc(fs, mbx, mby, 0, b));
    } else {
        put_ue(s, 0); // P_L0_16x16
        put_se(s, mb.mvd[0]);
        put_se(s, mb.mvd[1]);
        put_ue(s, h264_cbp_to_codenum_inter[mb.cbp]);
        if (mb.cbp)
            put_se(s, 0); // mb_qp_delta
        for (int b = 0; b < 16; b++)
            if (cbpl & (1 << (b >> 2)))
                cavlc_block(s, coef + b * 16, 16, cavlc_nc(fs, mbx, mby, 0, b));
    }
    if (cbpc) {
        cavlc_block(s, coef + 17 * 16, 4, -1);
        cavlc_block(s, coef + 17 * 16 + 4, 4, -1);
    }
    if (cbpc == 2)
        for (int c = 0; c < 2; c++)
            for (int b = 0; b < 4; b++)
                cavlc_block(s, coef + (18 + c * 4 + b) * 16 + 1, 15, cavlc_nc(fs, mbx, mby, 1 + c, b));
}

// =============================================================================================
// CABAC binarisation.  A bin is 16 bits: ctxIdx (0..459) in bits 0..9, BIN_BYPASS / BIN_TERM flags,
// value in bit 15.  The serial coder below consumes the stream.
// =============================================================================================
enum { BIN_BYPASS = 1 << 10, BIN_TERM = 1 << 11, BIN_VAL = 1 << 15 };

struct BinCount {
    unsigned n = 0;
    HD void bin(int, int) { n++; }
    HD void bypass(int) { n++; }
    HD void term(int) { n++; }
};

struct BinWrite {
    uint16_t *p;
    HD explicit BinWrite(uint16_t *q) : p(q) {}
    HD void bin(int ctx, int v) { *p++ = (uint16_t)(ctx | (v ? BIN_VAL : 0)); }
    HD void bypass(int v) { *p++ = (uint16_t)(BIN_BYPASS | (v ? BIN_VAL : 0)); }
    HD void term(int v) { *p++ = (uint16_t)(BIN_TERM | (v ? BIN_VAL : 0)); }
};

template <class S> HD void cabac_ueg_bypass(S &s, int k, int v)
{
    while (v >= (1 << k)) {
        s.bypass(1);
        v -= 1 << k;
        k++;
    }
    s.bypass(0);
    while (k--)
        s.bypass((v >> k) & 1);
}

// coded_block_flag of the neighbour (left / top) for ctxIdxInc
HD int cbf_neighbour(const FrameSyntax &fs, int mbx, int mby, int cat, int comp, int blk, int left, int intra)
{
    int navail = left ? mbx > 0 : top_avail(fs, mby);
    int nidx = mby * fs.mbw + mbx - (left ? 1 : fs.mbw);
    if (cat == 0) {
        if (!navail)
            return intra;
        return fs.mbi[nidx].type == MB_I16x16 ? fs.nnz[(size_t)nidx * NNZ_STRIDE + NNZ_DC16] != 0 : 0;
    }
    if (cat == 3) {
        if (!navail)
            return intra;
        return fs.nnz[(size_t)nidx * NNZ_STRIDE + NNZ_CBDC + comp] != 0;
    }
    int kind = cat == 4 ? 1 + comp : 0;
    int v = left ? nnz_left(fs, mbx, mby, kind, blk) : nnz_top(fs, mbx, mby, kind, blk);
    if (v < 0)
        return intra;
    return v != 0;
}

template <class S> HD void cabac_block(S &s, const int16_t *lev, int n, int cat, int cbf_inc)
{
    const int cbf_off[5] = {0, 4, 8, 12, 16}, sig_off[5] = {0, 15, 29, 44, 47}, abs_off[5] = {0, 10, 20, 30, 39};
    int last = -1;
    for (int i = 0; i < n; i++)
        if (lev[i])
            last = i;
    s.bin(85 + cbf_off[cat] + cbf_inc, last >= 0);
    if (last < 0)
        return;
    for (int i = 0; i < n - 1; i++) {
        int inc = cat == 3 ? imin_(i, 2) : i;
        if (lev[i]) {
            s.bin(105 + sig_off[cat] + inc, 1);
            s.bin(166 + sig_off[cat] + inc, i == last);
            if (i == last)
                break;
        } else
            s.bin(105 + sig_off[cat] + inc, 0);
    }
    int eq1 = 0, gt1 = 0;
    for (int i = last; i >= 0; i--) {
        int v = lev[i];
        if (!v)
            continue;
        int a = iabs_(v) - 1;
        int ctx = 227 + abs_off[cat] + (gt1 ? 0 : imin_(4, 1 + eq1));
        if (a == 0) {
            s.bin(ctx, 0);
            eq1++;
        } else {
            s.bin(ctx, 1);
            ctx = 227 + abs_off[cat] + 5 + imin_(4 - (cat == 3), gt1);
            int pre = imin_(a, 14);
            for (int j = 1; j < pre; j++)
                s.bin(ctx, 1);
            if (a < 14)
                s.bin(ctx, 0);
            else
                cabac_ueg_bypass(s, 0, a - 14);
            gt1++;
        }
        s.bypass(v < 0);
    }
}

template <class S> HD void cabac_mvd(S &s, int base, int mvd, int sum_abs)
{
    int inc = sum_abs < 3 ? 0 : (sum_abs > 32 ? 2 : 1);
    int a = iabs_(mvd);
    if (a == 0) {
        s.bin(base + inc, 0);
        return;
    }
    s.bin(base + inc, 1);
    int pre = imin_(a, 9);
    for (int i = 1; i < pre; i++)
        s.bin(base + imin_(2 + i, 6), 1);
    if (a < 9)
        s.bin(base + imin_(2 + pre, 6), 0);
    else
        cabac_ueg_bypass(s, 3, a - 9);
    s.bypass(mvd < 0);
}

// All bins of one macroblock including its end_of_slice_flag.
template <class S> HD void cabac_mb(S &s, const FrameSyntax &fs, int mb_index, int frame_i)
{
    int mbx = mb_index % fs.mbw, mby = mb_index / fs.mbw;
    const MbInfo &mb = fs.mbi[mb_index];
    const int row_in_slice = mby % fs.srows;
    const MbInfo *A = mbx > 0 ? &mb - 1 : nullptr, *B = row_in_slice ? &mb - fs.mbw : nullptr;
    const int16_t *coef = fs.coef + (size_t)mb_index * COEF_STRIDE;
    int intra = is_intra(mb.type);
    int end = mbx == fs.mbw - 1 && (row_in_slice == fs.srows - 1 || mby == fs.mbh - 1); // end_of_slice_flag
    if (!frame_i) {
        int inc = (A && A->type != MB_PSKIP) + (B && B->type != MB_PSKIP);
        s.bin(11 + inc, mb.type == MB_PSKIP);
        if (mb.type == MB_PSKIP) {
            s.term(end);
            return;
        }
    }
    int cbpl = mb.cbp & 15, cbpc = mb.cbp >> 4;
    if (mb.type == MB_I16x16) {
        int c0, c1, c2, c3, c4, c5;
        if (frame_i) {
            int inc = (A && A->type != MB_I4x4) + (B && B->type != MB_I4x4);
            c0 = 3 + inc, c1 = 3 + 3, c2 = 3 + 4, c3 = 3 + 5, c4 = 3 + 6, c5 = 3 + 7;
        } else {
            s.bin(14, 1);
            c0 = 17, c1 = 17 + 1, c2 = 17 + 2, c3 = 17 + 2, c4 = 17 + 3, c5 = 17 + 3;
        }
        s.bin(c0, 1);
        s.term(0);
        s.bin(c1, cbpl != 0);
        if (cbpc == 0)
            s.bin(c2, 0);
        else {
            s.bin(c2, 1);
            s.bin(c3, cbpc >> 1);
        }
        s.bin(c4, mb.i16_mode >> 1);
        s.bin(c5, mb.i16_mode & 1);
    } else if (mb.type == MB_I4x4) {
        if (frame_i)
            s.bin(3 + (A && A->type != MB_I4x4) + (B && B->type != MB_I4x4), 0); // I_NxN
        else {
            s.bin(14, 1);
            s.bin(17, 0);
        }
        for (int b = 0; b < 16; b++) {
            int pm = i4_pred_mode(fs, mbx, mby, b), mode = fs.i4[(size_t)mb_index * 16 + b];
            s.bin(68, mode == pm);
            if (mode != pm) {
                int rem = mode < pm ? mode : mode - 1;
                s.bin(69, rem & 1);
                s.bin(69, (rem >> 1) & 1);
                s.bin(69, (rem >> 2) & 1);
            }
        }
    } else {
        s.bin(14, 0);
        s.bin(15, 0);
        s.bin(16, 0);
    }
    if (intra) {
        int inc = (A && is_intra(A->type) && A->chroma_mode != 0) + (B && is_intra(B->type) && B->chroma_mode != 0);
        int m = mb.chroma_mode;
        s.bin(64 + inc, m > 0);
        if (m > 0) {
            s.bin(64 + 3, m > 1);
            if (m > 1)
                s.bin(64 + 3, m > 2);
        }
    } else {
        for (int k = 0; k < 2; k++) {
            int sa = (A && A->type == MB_P16x16 ? iabs_(A->mvd[k]) : 0) + (B && B->type == MB_P16x16 ? iabs_(B->mvd[k]) : 0);
            cabac_mvd(s, k ? 47 : 40, mb.mvd[k], sa);
        }
    }
    if (mb.type != MB_I16x16) {
        int cbp_a = A ? (A->cbp & 15) : 15, cbp_b = B ? (B->cbp & 15) : 15;
        for (int b8 = 0; b8 < 4; b8++) {
            int la = (b8 & 1) ? (cbpl >> (b8 - 1)) & 1 : (cbp_a >> (b8 + 1)) & 1;
            int lb = (b8 & 2) ? (cbpl >> (b8 - 2)) & 1 : (cbp_b >> (b8 + 2)) & 1;
            s.bin(73 + (!la) + 2 * (!lb), (cbpl >> b8) & 1);
        }
        int ca = A ? (A->cbp >> 4) : 0, cb = B ? (B->cbp >> 4) : 0;
        s.bin(77 + (ca > 0) + 2 * (cb > 0), cbpc > 0);
        if (cbpc > 0)
            s.bin(77 + 4 + (ca == 2) + 2 * (cb == 2), cbpc == 2);
    }
    if (mb.type == MB_I16x16 || mb.cbp)
        s.bin(60, 0); // mb_qp_delta == 0 everywhere => ctxIdxInc 0
    if (mb.type == MB_I16x16) {
        int inc = cbf_neighbour(fs, mbx, mby, 0, 0, 0, 1, intra) + 2 * cbf_neighbour(fs, mbx, mby, 0, 0, 0, 0, intra);
        cabac_block(s, coef + 16 * 16, 16, 0, inc);
        if (cbpl)
            for (int b = 0; b < 16; b++) {
                inc = cbf_neighbour(fs, mbx, mby, 1, 0, b, 1, intra) + 2 * cbf_neighbour(fs, mbx, mby, 1, 0, b, 0, intra);
                cabac_block(s, coef + b * 16 + 1, 15, 1, inc);
            }
    } else {
        for (int b = 0; b < 16; b++)
            if (cbpl & (1 << (b >> 2))) {
                int inc = cbf_neighbour(fs, mbx, mby, 2, 0, b, 1, intra) + 2 * cbf_neighbour(fs, mbx, mby, 2, 0, b, 0, intra);
                cabac_block(s, coef + b * 16, 16, 2, inc);
            }
    }
    if (cbpc)
        for (int comp = 0; comp < 2; comp++) {
            int inc = cbf_neighbour(fs, mbx, mby, 3, comp, 0, 1, intra) + 2 * cbf_neighbour(fs, mbx, mby, 3, comp, 0, 0, intra);
            cabac_block(s, coef + 17 * 16 + comp * 4, 4, 3, inc);
        }
    if (cbpc == 2)
        for (int comp = 0; comp < 2; comp++)
            for (int b = 0; b < 4; b++) {
                int inc = cbf_neighbour(fs, mbx, mby, 4, comp, b, 1, intra) + 2 * cbf_neighbour(fs, mbx, mby, 4, comp, b, 0, intra);
                cabac_block(s, coef + (18 + comp * 4 + b) * 16 + 1, 15, 4, inc);
            }
    s.term(end);
}

// =============================================================================================
// Serial CABAC arithmetic coder (H.264 9.3.4) with byte-wise output and carry propagation.
// `low` keeps the 10-bit coding window in bits 9..0 and queue+8 not-yet-written bits above it.
// =============================================================================================
// Lookup tables of the coder, rebuilt per coder instance (in shared memory on the GPU):
// lpsw[pStateIdx] = the four rangeTabLPS entries packed little-endian by qRangeIdx;
// shw[pStateIdx]  = their renormalisation shifts (3 bits each);
// trans[s] (s = pStateIdx << 1 | valMPS) = next s after an MPS in bits 0..7, after an LPS in bits 8..15.
struct CabacTables {
    uint32_t lpsw[64];
    uint16_t shw[64];
    uint16_t trans[128];
    HD void build(int tid, int nthreads)
    {
        for (int i = tid; i < 64; i += nthreads) {
            uint32_t w = 0, sh = 0;
            for (int q = 0; q < 4; q++) {
                uint32_t lps = h264_range_lps[i][q];
                w |= lps << (8 * q);
                sh |= (uint32_t)(8 - ilog2_(lps)) << (3 * q);
            }
            lpsw[i] = w;
            shw[i] = (uint16_t)sh;
        }
        for (int s = tid; s < 128; s += nthreads) {
            int st = s >> 1, mps = s & 1;
            int after_mps = (h264_next_state_mps[st] << 1) | mps;
            int after_lps = (h264_next_state_lps[st] << 1) | (st == 0 ? mps ^ 1 : mps);
            trans[s] = (uint16_t)(after_mps | (after_lps << 8));
        }
    }
};

// State after coding `bin` in state s (s = pStateIdx << 1 | valMPS).
HD uint32_t cabac_next_state(const CabacTables &t, uint32_t s, int bin)
{
    uint32_t tr = t.trans[s];
    return ((uint32_t)bin != (s & 1)) ? (tr >> 8) : (tr & 0xff);
}
HD uint32_t cabac_init_state(int ctx, int frame_i, int qp)
{
    int m = frame_i ? h264_cabac_init_I[ctx][0] : h264_cabac_init_P0[ctx][0];
    int n = frame_i ? h264_cabac_init_I[ctx][1] : h264_cabac_init_P0[ctx][1];
    int pre = clip3_(1, 126, ((m * clip3_(0, 51, qp)) >> 4) + n);
    return (uint32_t)(pre <= 63 ? (63 - pre) << 1 : (((pre - 64) << 1) | 1));
}

// The arithmetic coder (H.264 9.3.4) is split into three stages so that only a minimal recurrence stays
// serial:
//   1. context-state resolution: the state each regular bin is coded in ("pre-state") depends only on the
//      earlier bins of the same context => all contexts in parallel (cabac_next_state);
//   2. CabacRange: the codIRange recurrence.  Per bin it emits an "interval step" word: what to add to
//      codILow and by how much to shift, which no longer depends on anything but the step itself;
//   3. CabacBytes: the codILow recurrence + byte output with carry propagation.
// Stages 2 and 3 are two independent serial chains that run pipelined on different warps.
//
// Interval step word: bits 0..8 add, 9..11 shift after the add, 12 shift-by-one before the add (bypass),
// 13 end of slice (flush).
enum { STEP_PRE1 = 1 << 12, STEP_FINAL = 1 << 13 };

// meta word of a staged bin: bit 0 = isLPS (regular) or value (bypass / terminate), bit 1 bypass,
// bit 2 terminate, bits 4..15 = LPS renormalisation shifts (shw)
HD uint32_t cabac_stage_meta(uint16_t b, uint32_t pre_state, const CabacTables &t)
{
    uint32_t special = (b >> 10) & 3, v = (b >> 15) & 1;
    return (special ? (v | (special << 1)) : (v ^ (pre_state & 1))) | ((uint32_t)t.shw[(pre_state >> 1) & 63] << 4);
}

struct CabacRange {
    uint32_t range = 510;
    HD uint32_t step(uint32_t lps4, uint32_t meta)
    {
        if (meta & 6) {
            if (meta & 2) // bypass: low = (low << 1) + (bin ? range : 0)
                return ((meta & 1) ? range : 0) | STEP_PRE1;
            range -= 2; // terminate
            if (meta & 1) {
                uint32_t add = range;
                range = 2;
                return add | (7u << 9) | STEP_FINAL;
            }
            uint32_t sh = (range >> 8) ^ 1; // range in [254, 508]
            range <<= sh;
            return sh << 9;
        }
        uint32_t q = (range >> 6) & 3;
#ifdef __CUDACC__
        uint32_t lps = __byte_perm(lps4, 0, 0x4440u | q);
#else
        uint32_t lps = (lps4 >> (q * 8)) & 0xff;
#endif
        uint32_t rm = range - lps; // >= 128: the MPS path renormalises by at most one bit
        if (meta & 1) {
            uint32_t sh = (meta >> (4 + 3 * q)) & 7;
            range = lps << sh;
            return rm | (sh << 9);
        }
        uint32_t sh = (rm >> 8) ^ 1;
        range = rm << sh;
        return sh << 9;
    }
};

// Same recurrence as CabacRange::step with the regular-bin path written without branches (both the MPS and
// the LPS continuation are computed and selected), so that an unrolled loop overlaps the independent work
// of neighbouring bins and only the short range -> range dependency stays serial.
HD uint32_t cabac_range_step_flat(uint32_t &range, uint32_t lps4, uint32_t meta)
{
    const uint32_t q = (range >> 6) & 3;
#ifdef __CUDACC__
    const uint32_t lps = __byte_perm(lps4, 0, 0x4440u | q);
#else
    const uint32_t lps = (lps4 >> (q * 8)) & 0xff;
#endif
    const uint32_t rm = range - lps;
    const uint32_t sh_m = (rm >> 8) ^ 1, sh_l = (meta >> (4 + 3 * q)) & 7;
    const bool isl = meta & 1;
    uint32_t nr = isl ? (lps << sh_l) : (rm << sh_m);
    uint32_t w = isl ? (rm | (sh_l << 9)) : (sh_m << 9);
    if (meta & 6) { // bypass / terminate: rare enough for a branch
        if (meta & 2) {
            nr = range;
            w = (isl ? range : 0) | STEP_PRE1;
        } else {
            const uint32_t r2 = range - 2, sh = (r2 >> 8) ^ 1;
            nr = isl ? 2u : (r2 << sh);
            w = isl ? (r2 | (7u << 9) | STEP_FINAL) : (sh << 9);
        }
    }
    range = nr;
    return w;
}

struct CabacBytes {
    uint8_t *out;      // output bytes (first byte written at out[0])
    unsigned pos = 0;  // bytes written
    uint32_t low = 0;  // the 10-bit coding window in bits 9..0 and queue + 8 not-yet-written bits above it
    int queue = -9, outstanding = 0;
    int last = -1;     // most recent byte, held back until a later carry can no longer reach it

    // Pops the top byte of `low`.  o has 9 bits: a carry (only ever together with a 0x00 byte, because
    // low < 2^(queue+18) + range) and the byte.  0xff bytes are counted, not written, until the next
    // non-0xff byte shows whether a carry ripples through them.
    HD void put_byte()
    {
        uint32_t o = low >> (queue + 10);
        low &= (0x400u << queue) - 1;
        queue -= 8;
        if ((o & 0xff) == 0xff) {
            outstanding++;
            return;
        }
        uint32_t carry = o >> 8;
        if (last >= 0)
            out[pos++] = (uint8_t)(last + carry);
#pragma unroll 1
        while (outstanding > 0) {
            out[pos++] = (uint8_t)(carry - 1);
            outstanding--;
        }
        last = (int)(o & 0xff);
    }
    HD void step_fast(uint32_t w) // any step but the final one
    {
        uint32_t pre1 = (w >> 12) & 1, sh = (w >> 9) & 7;
        low = ((low << pre1) + (w & 0x1ff)) << sh;
        queue += (int)(pre1 + sh);
        if (queue >= 0)
            put_byte();
    }
    HD void step(uint32_t w)
    {
        step_fast(w);
        if (w & STEP_FINAL) {
            // 9.3.4.5: put_bit(low >> 9 & 1); write_bits((low >> 7 & 3) | 1, 2): push window bits 9, 8
            // and the stop bit out, then pad the last byte with zeros.
            low = (low & ~0x7fu) | 0x80u;
            low <<= 3;
            queue += 3;
            if (queue >= 0)
                put_byte();
            if (queue > -8) {
                int k = -queue;
                low <<= k;
                queue += k;
                put_byte();
            }
            if (last >= 0)
                out[pos++] = (uint8_t)last;
            last = -1;
#pragma unroll 1
            while (outstanding > 0) {
                out[pos++] = 0xff;
                outstanding--;
            }
        }
    }
};

// =============================================================================================
// Parallel formulation of the arithmetic coder (what cabac_code_kernel runs; the three-stage serial coder
// above is kept as the CPU-checkable restatement of 9.3.4 it is derived from).
//
// (1) codIRange.  After a regular bin coded as LPS the new range is rangeTabLPS[state][q] << shift: it depends
//     on the range before that bin only through q = (range >> 6) & 3.  So the bin sequence is cut right after
//     LPS bins ("chunks" of about CP_K bins); a chunk is walked for all four q hypotheses of the LPS bin in
//     front of it, which yields a 4 -> 4 map (hypothesis -> q at the chunk's closing LPS bin) and the number
//     of bits the chunk shifts out.  Composing the maps (an associative scan) gives every chunk its true
//     start range; a second walk then produces the true interval steps.  All chunks run in parallel.
// (2) codILow.  With every bin's (add, shift) known, the code word is the big integer
//     sum_i add_i << (T_end - P_i), P_i = bits shifted out before bin i's add: additions commute, so every bin
//     adds its 9-bit value into an array of 16-bit limbs (held in 32-bit words: deferred carries) at stream
//     bit position P_i, and one carry-lookahead pass turns the limbs into bytes.  This replaces put_byte's
//     outstanding-0xff bookkeeping: a carry rippling through 0xff bytes is just a propagated carry.
//     Stream bit p <-> code-word bit T_end + 8 - p; the flush of 9.3.4.5 keeps code-word bits >= 8, then the
//     rbsp stop bit, i.e. T_end + 2 bits in total.
// =============================================================================================
// per-bin record left by the context-state resolver: bit 0 = isLPS (regular) or value (bypass / terminate),
// bit 1 bypass, bit 2 terminate, bits 3..8 pStateIdx of the bin's context when it is coded
enum { META_LPS = 1, META_BYPASS = 2, META_TERM = 4 };
HD uint16_t cabac_meta(uint16_t bin, uint32_t pre_state)
{
    uint32_t special = (bin >> 10) & 3, v = (bin >> 15) & 1;
    return (uint16_t)(special ? (v | (special << 1)) : ((v ^ (pre_state & 1)) | ((pre_state >> 1) << 3)));
}
HD bool cabac_meta_is_lps(uint32_t m) { return (m & 7) == META_LPS; }
HD uint32_t byte_of(uint32_t w, uint32_t q)
{
#ifdef __CUDACC__
    return __byte_perm(w, 0, 0x4440u | q);
#else
    return (w >> (q * 8)) & 0xff;
#endif
}
// range entering the bin after an LPS bin of state `st` that was coded with range quantiser index q
HD uint32_t cabac_range_after_lps(uint32_t lps4, uint32_t shw, uint32_t q) { return byte_of(lps4, q) << ((shw >> (3 * q)) & 7); }

// One step of the range recurrence: new range, the value added to the code word (after `pre1` bits and before
// `sh` more bits are shifted out).  lps4 / shw = CabacTables rows of the bin's state (ignored for bypass /
// terminate).  Branch-free on the regular path.
HD void cabac_rstep(uint32_t &range, uint32_t m, uint32_t lps4, uint32_t shw, uint32_t &add, uint32_t &pre1, uint32_t &sh)
{
    const uint32_t q = (range >> 6) & 3, lps = byte_of(lps4, q), rm = range - lps;
    const uint32_t sh_m = (rm >> 8) ^ 1, sh_l = (shw >> (3 * q)) & 7;
    const bool v = m & 1;
    uint32_t nr = v ? (lps << sh_l) : (rm << sh_m);
    add = v ? rm : 0;
    sh = v ? sh_l : sh_m;
    pre1 = 0;
    if (m & (META_BYPASS | META_TERM)) {
        if (m & META_BYPASS) { // low = (low << 1) + (bin ? range : 0)
            nr = range;
            add = v ? range : 0;
            pre1 = 1;
            sh = 0;
        } else { // terminate: range -= 2; value 1 ends the slice (low += range; range = 2; renorm by 7)
            const uint32_t r2 = range - 2, s2 = (r2 >> 8) ^ 1;
            nr = v ? 256u : (r2 << s2);
            add = v ? r2 : 0;
            sh = v ? 7u : s2;
        }
    }
    range = nr;
}

// Chunk summary for the scan: hypothesis h -> q at the closing LPS bin (2 bits each in qmap) and the bits the
// chunk shifts out under hypothesis h.  Identity = an absent chunk.
struct ChunkMap {
    uint32_t qmap;
    uint32_t s[4];
};
HD ChunkMap chunkmap_identity()
{
    ChunkMap m;
    m.qmap = 0xe4; // 3,2,1,0
    m.s[0] = m.s[1] = m.s[2] = m.s[3] = 0;
    return m;
}
HD uint32_t sel4(const uint32_t *v, uint32_t q) { return q & 2 ? (q & 1 ? v[3] : v[2]) : (q & 1 ? v[1] : v[0]); }
// a then b
HD ChunkMap chunkmap_compose(const ChunkMap &a, const ChunkMap &b)
{
    ChunkMap c;
    c.qmap = 0;
#pragma unroll
    for (int h = 0; h < 4; h++) {
        uint32_t qa = (a.qmap >> (2 * h)) & 3;
        c.qmap |= ((b.qmap >> (2 * qa)) & 3) << (2 * h);
        c.s[h] = a.s[h] + sel4(b.s, qa);
    }
    return c;
}

// Adds the 9-bit value `add` at stream bit position P (its MSB lands on bit P) into 16-bit limbs held in
// 32-bit words: limb j covers stream bits [16 j, 16 j + 16), bit 16 j is the limb's bit 15.
template <class A> HD void limb_add(A &&adder, unsigned long long P, uint32_t add)
{
    unsigned long long j = P >> 4;
    uint32_t off = (uint32_t)P & 15;
    if (off <= 7)
        adder(j, add << (7 - off));
    else {
        adder(j, add >> (off - 7));
        adder(j + 1, (add & ((1u << (off - 7)) - 1)) << (23 - off));
    }
}

// Emulation prevention: an escape byte 0x03 goes before byte i iff byte i <= 3 and the run of zero
// bytes directly before it has an even length >= 2 (equivalent to the sequential rule
// "after 00 00 insert 03 if the next byte <= 3, then restart counting").
HD int epb_needed(int byte, unsigned zero_run) { return byte <= 3 && zero_run >= 2 && !(zero_run & 1); }

} // namespace cedar
