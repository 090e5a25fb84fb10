/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Golden-model DECODER for the H.264 subset the encoder emits (north star, correctness part 3:
 * "decoding the output with an independent conformant decoder -- libavcodec if installed offline, otherwise the
 * golden model's decoder -- must reproduce the encoder's reconstructed frames bit-exactly").  libavcodec is the
 * independent check (tests/avdec.py); this decoder is the fallback that removes that dependency.
 *
 * Written from the decoding side of the standard: Annex-B / emulation-prevention parsing, SPS / PPS / slice header
 * parsing (7.3), CAVLC parsing (9.2: coeff_token, level_prefix / suffix, total_zeros, run_before decoded by matching the
 * code tables), the CABAC decoding engine (9.3.3.2) with the binarisations of 9.3.2, then reconstruction.  It shares
 * with the encoder model (this file includes h264_golden.c) only the sample-domain primitives and neighbour helpers:
 * intra predictors, inverse transform, dequantisation, deblocking, median MV prediction, nC / coded_block_flag
 * neighbour rules.  Subset: Main profile, frame MBs, 4:2:0, one reference, I slices with I_NxN / I_16x16 and P slices
 * with P_L0_16x16 / P_Skip / intra macroblocks, integer-pel luma vectors, constant QP (mb_qp_delta = 0), slices of
 * whole macroblock rows.  Anything else returns an error instead of guessing.
 */
#include "h264_golden.c"
#include "h264_decoder.h"

#include <stdio.h>

/* ---------------------------------------------------------------- bit reader ---------- */
typedef struct {
    const uint8_t *p;
    size_t nbits, pos;
} bitr;

static int br_bit(bitr *b)
{
    if (b->pos >= b->nbits)
        return 0;
    int v = (b->p[b->pos >> 3] >> (7 - (b->pos & 7))) & 1;
    b->pos++;
    return v;
}
static uint32_t br_u(bitr *b, int n)
{
    uint32_t v = 0;
    while (n--)
        v = (v << 1) | (uint32_t)br_bit(b);
    return v;
}
static uint32_t br_peek(bitr *b, int n)
{
    size_t save = b->pos;
    uint32_t v = br_u(b, n);
    b->pos = save;
    return v;
}
static uint32_t br_ue(bitr *b)
{
    int z = 0;
    while (!br_bit(b) && z < 32 && b->pos < b->nbits)
        z++;
    return z ? ((1u << z) - 1) + br_u(b, z) : 0;
}
static int br_se(bitr *b)
{
    uint32_t k = br_ue(b);
    return (k & 1) ? (int)((k + 1) >> 1) : -(int)(k >> 1);
}
/* more_rbsp_data(): anything before the final stop bit? */
static int br_more(const bitr *b)
{
    size_t last = b->nbits;
    while (last > 0 && !((b->p[(last - 1) >> 3] >> (7 - ((last - 1) & 7))) & 1))
        last--;
    return last > 0 && b->pos < last - 1;
}

/* ---------------------------------------------------------------- decoder state ---------- */
struct gd_decoder {
    gm_encoder e; /* geometry, macroblock records, unfiltered picture, two reference pictures */
    int have_sps, have_pps, allocated;
    int cabac, pic_init_qp, chroma_qp_offset, deblock_ctrl, log2_max_frame_num, poc_type;
    int crop_r, crop_b;
    int pic_open, disable_deblock, prev_slice_row;
    uint8_t **frames; /* finished pictures: Y | U | V at coded size */
    int nframes, cap;
    uint8_t *rbsp;
    size_t rbsp_cap;
    char err[128];
};

#define FAIL(d, ...)                                   \
    do {                                               \
        snprintf((d)->err, sizeof((d)->err), __VA_ARGS__); \
        return -1;                                     \
    } while (0)

gd_decoder *gd_open(void) { return (gd_decoder *)calloc(1, sizeof(gd_decoder)); }
const char *gd_error(const gd_decoder *d) { return d->err; }
int gd_width(const gd_decoder *d) { return d->e.W; }
int gd_height(const gd_decoder *d) { return d->e.H; }
int gd_crop_right(const gd_decoder *d) { return d->crop_r; }
int gd_crop_bottom(const gd_decoder *d) { return d->crop_b; }
int gd_frames(const gd_decoder *d) { return d->nframes; }
const uint8_t *gd_frame(const gd_decoder *d, int i, int plane)
{
    if (i < 0 || i >= d->nframes)
        return NULL;
    size_t ysz = (size_t)d->e.W * d->e.H;
    return d->frames[i] + (plane == 0 ? 0 : (plane == 1 ? ysz : ysz + ysz / 4));
}
void gd_close(gd_decoder *d)
{
    if (!d)
        return;
    for (int i = 0; i < d->nframes; i++)
        free(d->frames[i]);
    free(d->frames);
    frame_free(&d->e.rec[0]);
    frame_free(&d->e.rec[1]);
    frame_free(&d->e.unf);
    free(d->e.mbs);
    free(d->rbsp);
    free(d);
}

/* ---------------------------------------------------------------- parameter sets (7.3.2) ---------- */
static int parse_sps(gd_decoder *d, bitr *b)
{
    int profile = (int)br_u(b, 8);
    br_u(b, 8);
    br_u(b, 8); /* level */
    br_ue(b);   /* sps id */
    if (profile >= 100)
        FAIL(d, "high profiles are outside the subset");
    d->log2_max_frame_num = 4 + (int)br_ue(b);
    d->poc_type = (int)br_ue(b);
    if (d->poc_type == 0)
        br_ue(b);
    else if (d->poc_type == 1)
        FAIL(d, "pic_order_cnt_type 1 is outside the subset");
    br_ue(b); /* max_num_ref_frames */
    br_u(b, 1);
    int wmb = (int)br_ue(b) + 1, hmb = (int)br_ue(b) + 1;
    if (!br_u(b, 1))
        FAIL(d, "field coding is outside the subset");
    br_u(b, 1); /* direct_8x8_inference */
    d->crop_r = d->crop_b = 0;
    if (br_u(b, 1)) {
        br_ue(b);
        d->crop_r = 2 * (int)br_ue(b);
        br_ue(b);
        d->crop_b = 2 * (int)br_ue(b);
    }
    if (d->allocated && (d->e.mbw != wmb || d->e.mbh != hmb))
        FAIL(d, "picture size change is outside the subset");
    if (!d->allocated) {
        gm_encoder *e = &d->e;
        e->mbw = wmb, e->mbh = hmb, e->W = wmb * 16, e->H = hmb * 16, e->srows = hmb;
        if (frame_alloc(&e->rec[0], e->W, e->H) | frame_alloc(&e->rec[1], e->W, e->H) | frame_alloc(&e->unf, e->W, e->H))
            FAIL(d, "out of memory");
        e->mbs = (gm_mb *)calloc((size_t)wmb * hmb, sizeof(gm_mb));
        d->allocated = 1;
    }
    d->have_sps = 1;
    return 0;
}

static int parse_pps(gd_decoder *d, bitr *b)
{
    br_ue(b);
    br_ue(b);
    d->cabac = (int)br_u(b, 1);
    br_u(b, 1);
    if (br_ue(b))
        FAIL(d, "slice groups are outside the subset");
    if (br_ue(b) || br_ue(b))
        FAIL(d, "more than one reference picture is outside the subset");
    if (br_u(b, 1) || br_u(b, 2))
        FAIL(d, "weighted prediction is outside the subset");
    d->pic_init_qp = 26 + br_se(b);
    br_se(b);
    d->chroma_qp_offset = br_se(b);
    d->deblock_ctrl = (int)br_u(b, 1);
    if (br_u(b, 1))
        FAIL(d, "constrained_intra_pred is outside the subset");
    if (br_u(b, 1))
        FAIL(d, "redundant pictures are outside the subset");
    d->have_pps = 1;
    return 0;
}

/* ---------------------------------------------------------------- reconstruction ---------- */
static void recon_chroma(gd_decoder *d, gm_mb *mb, int mbx, int mby, const uint8_t pred[2][64])
{
    gm_encoder *e = &d->e;
    int CW = e->W / 2, qpc = e->qpc, cbpc = mb->cbp >> 4;
    for (int c = 0; c < 2; c++) {
        uint8_t *dst = e->unf.p[1 + c] + (size_t)(mby * 8) * CW + mbx * 8;
        int dcq[4] = {0, 0, 0, 0};
        if (cbpc >= 1) { /* 8.5.11.1 / 8.5.11.2: 2x2 inverse transform and scaling of the chroma DC levels */
            const int16_t *z = &mb->coef[17][c * 4];
            int f0 = z[0] + z[1] + z[2] + z[3], f1 = z[0] - z[1] + z[2] - z[3];
            int f2 = z[0] + z[1] - z[2] - z[3], f3 = z[0] - z[1] - z[2] + z[3];
            int ls = 16 * h264_dequant_v[qpc % 6][0];
            dcq[0] = ((f0 * ls) * (1 << (qpc / 6))) >> 5;
            dcq[1] = ((f1 * ls) * (1 << (qpc / 6))) >> 5;
            dcq[2] = ((f2 * ls) * (1 << (qpc / 6))) >> 5;
            dcq[3] = ((f3 * ls) * (1 << (qpc / 6))) >> 5;
        }
        for (int b = 0; b < 4; b++) {
            int bx = (b & 1) * 4, by = (b >> 1) * 4, dd[16] = {0};
            if (cbpc == 2)
                dequant_block(mb->coef[18 + c * 4 + b], qpc, 1, dd);
            dd[0] = dcq[b];
            idct4x4_add(dd, dst + by * CW + bx, CW, pred[c] + by * 8 + bx, 8);
        }
    }
}

static void recon_intra(gd_decoder *d, gm_mb *mb, int mbx, int mby)
{
    gm_encoder *e = &d->e;
    int W = e->W, CW = W / 2, qp = e->qp;
    int has_top = top_avail(e, mby), has_left = mbx > 0;
    uint8_t *dst = e->unf.p[0] + (size_t)(mby * 16) * W + mbx * 16;
    if (mb->type == GM_MB_I16x16) {
        uint8_t top[16] = {0}, left[16] = {0}, pred[256];
        int tl = 0;
        if (has_top)
            memcpy(top, dst - W, 16);
        if (has_left)
            for (int y = 0; y < 16; y++)
                left[y] = dst[y * W - 1];
        if (has_top && has_left)
            tl = dst[-W - 1];
        pred16x16(mb->i16_mode, top, left, tl, has_top, has_left, pred);
        /* 8.5.10: inverse Hadamard of the DC levels, then scaling */
        int c[16], t[16], fdc[16];
        for (int i = 0; i < 16; i++)
            c[h264_zigzag4x4[i]] = mb->coef[16][i];
        for (int i = 0; i < 4; i++) {
            int a = c[i * 4], b = c[i * 4 + 1], cc = c[i * 4 + 2], dd = c[i * 4 + 3];
            t[i * 4 + 0] = a + b + cc + dd, t[i * 4 + 1] = a + b - cc - dd, t[i * 4 + 2] = a - b - cc + dd, t[i * 4 + 3] = a - b + cc - dd;
        }
        for (int i = 0; i < 4; i++) {
            int a = t[i], b = t[4 + i], cc = t[8 + i], dd = t[12 + i];
            fdc[i] = a + b + cc + dd, fdc[4 + i] = a + b - cc - dd, fdc[8 + i] = a - b - cc + dd, fdc[12 + i] = a - b + cc - dd;
        }
        int ls = 16 * h264_dequant_v[qp % 6][0];
        for (int b = 0; b < 16; b++) {
            int bx = blk_x[b] * 4, by = blk_y[b] * 4, dd[16] = {0};
            if (mb->cbp & 15)
                dequant_block(mb->coef[b], qp, 1, dd);
            int fv = fdc[blk_y[b] * 4 + blk_x[b]];
            dd[0] = qp >= 36 ? (fv * ls) * (1 << (qp / 6 - 6)) : (fv * ls + (1 << (5 - qp / 6))) >> (6 - qp / 6);
            idct4x4_add(dd, dst + by * W + bx, W, pred + by * 16 + bx, 16);
        }
    } else { /* Intra4x4: block by block, each predicted from the samples reconstructed so far (8.3.1.2) */
        for (int b = 0; b < 16; b++) {
            int bx = blk_x[b], by = blk_y[b];
            int ht = by > 0 || has_top, hl = bx > 0 || has_left, htl = (bx > 0 || has_left) && (by > 0 || has_top);
            int htr = by == 0 ? (has_top && (bx < 3 || mbx + 1 < e->mbw)) : (bx < 3 && xy2blk[by - 1][bx + 1] < b);
            uint8_t *d4 = dst + by * 4 * W + bx * 4, pred[16];
            int t[8] = {0}, l[4] = {0}, m = 0;
            if (ht)
                for (int i = 0; i < 8; i++)
                    t[i] = d4[-W + (i < 4 || htr ? i : 3)];
            if (hl)
                for (int i = 0; i < 4; i++)
                    l[i] = d4[i * W - 1];
            if (htl)
                m = d4[-W - 1];
            pred4x4(mb->i4_mode[b], t, l, m, ht, hl, pred);
            int dd[16] = {0};
            if (mb->nnz[b])
                dequant_block(mb->coef[b], qp, 0, dd);
            idct4x4_add(dd, d4, W, pred, 4);
        }
    }
    uint8_t cpred[2][64];
    for (int c = 0; c < 2; c++) {
        uint8_t *cd = e->unf.p[1 + c] + (size_t)(mby * 8) * CW + mbx * 8, ctop[8] = {0}, cleft[8] = {0};
        int ctl = 0;
        if (has_top)
            memcpy(ctop, cd - CW, 8);
        if (has_left)
            for (int y = 0; y < 8; y++)
                cleft[y] = cd[y * CW - 1];
        if (has_top && has_left)
            ctl = cd[-CW - 1];
        pred_chroma8x8(mb->chroma_mode, ctop, cleft, ctl, has_top, has_left, cpred[c]);
    }
    recon_chroma(d, mb, mbx, mby, cpred);
}

static int recon_inter(gd_decoder *d, gm_mb *mb, int mbx, int mby)
{
    gm_encoder *e = &d->e;
    const frame_t *ref = &e->rec[e->cur ^ 1];
    int W = e->W, H = e->H, CW = W / 2, CH = H / 2, qp = e->qp;
    if ((mb->mv[0] & 3) || (mb->mv[1] & 3))
        FAIL(d, "fractional luma vectors are outside the subset");
    int dx = mb->mv[0] >> 2, dy = mb->mv[1] >> 2;
    uint8_t pred[256], *dst = e->unf.p[0] + (size_t)(mby * 16) * W + mbx * 16;
    for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++)
            pred[y * 16 + x] = ref->p[0][(size_t)CLIP3(0, H - 1, mby * 16 + y + dy) * W + CLIP3(0, W - 1, mbx * 16 + x + dx)];
    for (int b = 0; b < 16; b++) {
        int bx = blk_x[b] * 4, by = blk_y[b] * 4, dd[16] = {0};
        if (mb->nnz[b])
            dequant_block(mb->coef[b], qp, 0, dd);
        idct4x4_add(dd, dst + by * W + bx, W, pred + by * 16 + bx, 16);
    }
    uint8_t cpred[2][64]; /* 8.4.2.2.2: chroma vectors have eighth-sample accuracy */
    int xi = mb->mv[0] >> 3, yi = mb->mv[1] >> 3, xf = mb->mv[0] & 7, yf = mb->mv[1] & 7;
    for (int c = 0; c < 2; c++)
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) {
                int y0 = CLIP3(0, CH - 1, mby * 8 + y + yi), y1 = CLIP3(0, CH - 1, mby * 8 + y + yi + 1);
                int x0 = CLIP3(0, CW - 1, mbx * 8 + x + xi), x1 = CLIP3(0, CW - 1, mbx * 8 + x + xi + 1);
                const uint8_t *rp = ref->p[1 + c];
                cpred[c][y * 8 + x] = (uint8_t)(((8 - xf) * (8 - yf) * rp[(size_t)y0 * CW + x0] + xf * (8 - yf) * rp[(size_t)y0 * CW + x1] +
                                                 (8 - xf) * yf * rp[(size_t)y1 * CW + x0] + xf * yf * rp[(size_t)y1 * CW + x1] + 32) >> 6);
            }
    recon_chroma(d, mb, mbx, mby, cpred);
    return 0;
}

/* ---------------------------------------------------------------- CAVLC parsing (9.2) ---------- */
static int vlc_match(bitr *b, const uint8_t *len, const uint8_t *bits, int n)
{
    uint32_t pk = br_peek(b, 16);
    for (int i = 0; i < n; i++)
        if (len[i] && (pk >> (16 - len[i])) == bits[i]) {
            b->pos += len[i];
            return i;
        }
    return -1;
}

/* residual_block_cavlc: levels into lev[0 .. max_coeff - 1]; returns total_coeff or -1 */
static int cavlc_read_block(bitr *b, int16_t *lev, int max_coeff, int nC)
{
    int idx;
    if (nC < 0)
        idx = vlc_match(b, h264_chroma_dc_coeff_token_len, h264_chroma_dc_coeff_token_bits, 20);
    else {
        int tab = nC < 2 ? 0 : (nC < 4 ? 1 : (nC < 8 ? 2 : 3));
        idx = vlc_match(b, h264_coeff_token_len[tab], h264_coeff_token_bits[tab], 68);
    }
    if (idx < 0)
        return -1;
    int total = idx >> 2, t1 = idx & 3, level[16], run[16];
    for (int i = 0; i < max_coeff; i++)
        lev[i] = 0;
    if (!total)
        return 0;
    if (total > max_coeff)
        return -1;
    int suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
    for (int i = 0; i < total; i++) {
        if (i < t1) {
            level[i] = 1 - 2 * (int)br_u(b, 1);
            continue;
        }
        int prefix = 0;
        while (!br_bit(b) && prefix < 32)
            prefix++;
        int ssize = (prefix == 14 && suffix_len == 0) ? 4 : (prefix >= 15 ? prefix - 3 : suffix_len);
        int code = (imin(15, prefix) << suffix_len) + (ssize ? (int)br_u(b, ssize) : 0);
        if (prefix >= 15 && suffix_len == 0)
            code += 15;
        if (prefix >= 16)
            code += (1 << (prefix - 3)) - 4096;
        if (i == t1 && t1 < 3)
            code += 2;
        level[i] = (code & 1) ? (-code - 1) >> 1 : (code + 2) >> 1;
        if (suffix_len == 0)
            suffix_len = 1;
        if (iabs(level[i]) > (3 << (suffix_len - 1)) && suffix_len < 6)
            suffix_len++;
    }
    int zeros_left = 0;
    if (total < max_coeff) {
        zeros_left = nC < 0 ? vlc_match(b, h264_chroma_dc_total_zeros_len[total - 1], h264_chroma_dc_total_zeros_bits[total - 1], 4)
                            : vlc_match(b, h264_total_zeros_len[total - 1], h264_total_zeros_bits[total - 1], 16);
        if (zeros_left < 0)
            return -1;
    }
    for (int i = 0; i < total - 1; i++) {
        run[i] = 0;
        if (zeros_left > 0) {
            int zl = imin(zeros_left, 7) - 1;
            run[i] = vlc_match(b, h264_run_before_len[zl], h264_run_before_bits[zl], 16);
            if (run[i] < 0)
                return -1;
        }
        zeros_left -= run[i];
    }
    if (zeros_left < 0)
        return -1;
    run[total - 1] = zeros_left;
    int pos = -1;
    for (int i = total - 1; i >= 0; i--) {
        pos += run[i] + 1;
        if (pos >= max_coeff)
            return -1;
        lev[pos] = (int16_t)level[i];
    }
    return total;
}

static int cavlc_read_residual(gd_decoder *d, bitr *b, gm_mb *mb, int mbx, int mby)
{
    gm_encoder *e = &d->e;
    int cbpl = mb->cbp & 15, cbpc = mb->cbp >> 4, n;
    if (mb->type == GM_MB_I16x16) {
        if ((n = cavlc_read_block(b, mb->coef[16], 16, calc_nc(e, mbx, mby, 0, 0))) < 0)
            FAIL(d, "bad Intra16x16 DC block");
        mb->nnz[16] = (uint8_t)n;
    }
    for (int blk = 0; blk < 16; blk++) {
        if (!(cbpl & (1 << (blk >> 2))))
            continue;
        int ac = mb->type == GM_MB_I16x16;
        if ((n = cavlc_read_block(b, mb->coef[blk] + ac, 16 - ac, calc_nc(e, mbx, mby, 0, blk))) < 0)
            FAIL(d, "bad luma block");
        mb->nnz[blk] = (uint8_t)n;
    }
    if (cbpc)
        for (int c = 0; c < 2; c++) {
            if ((n = cavlc_read_block(b, mb->coef[17] + c * 4, 4, -1)) < 0)
                FAIL(d, "bad chroma DC block");
            mb->nnz[25 + c] = (uint8_t)n;
        }
    if (cbpc == 2)
        for (int c = 0; c < 2; c++)
            for (int blk = 0; blk < 4; blk++) {
                if ((n = cavlc_read_block(b, mb->coef[18 + c * 4 + blk] + 1, 15, calc_nc(e, mbx, mby, 1 + c, blk))) < 0)
                    FAIL(d, "bad chroma AC block");
                mb->nnz[17 + c * 4 + blk] = (uint8_t)n;
            }
    return 0;
}

/* I macroblock types (table 7-11) -> fields of the record; mb_type 0 = I_NxN, 1..24 = I_16x16 variants */
static int set_intra_type(gd_decoder *d, gm_mb *mb, int t)
{
    if (t == 0) {
        mb->type = GM_MB_I4x4;
        return 0;
    }
    if (t > 24)
        FAIL(d, "I_PCM is outside the subset");
    mb->type = GM_MB_I16x16;
    mb->i16_mode = (uint8_t)((t - 1) & 3);
    mb->cbp = (uint8_t)((((t - 1) >> 2) % 3) << 4 | (t >= 13 ? 15 : 0));
    return 0;
}

static int finish_p_mv(gd_decoder *d, gm_mb *mb, int mbx, int mby, int skip)
{
    int16_t mvp[2], smv[2];
    predict_mv(&d->e, mbx, mby, mvp, smv);
    if (skip) {
        mb->mv[0] = smv[0], mb->mv[1] = smv[1];
        mb->mvd[0] = mb->mvd[1] = 0;
    } else {
        mb->mv[0] = (int16_t)(mvp[0] + mb->mvd[0]);
        mb->mv[1] = (int16_t)(mvp[1] + mb->mvd[1]);
    }
    return 0;
}

static int cavlc_slice_decode(gd_decoder *d, bitr *b, int first_mb, int slice_i)
{
    gm_encoder *e = &d->e;
    int nmb = e->mbw * e->mbh, addr = first_mb;
    while (addr < nmb && br_more(b)) {
        if (!slice_i) {
            uint32_t run = br_ue(b);
            for (; run && addr < nmb; run--, addr++) {
                gm_mb *mb = &e->mbs[addr];
                memset(mb, 0, sizeof(*mb));
                mb->type = GM_MB_PSKIP;
                finish_p_mv(d, mb, addr % e->mbw, addr / e->mbw, 1);
                if (recon_inter(d, mb, addr % e->mbw, addr / e->mbw))
                    return -1;
            }
            if (run)
                FAIL(d, "mb_skip_run beyond the picture");
            if (addr >= nmb || !br_more(b))
                break;
        }
        gm_mb *mb = &e->mbs[addr];
        int mbx = addr % e->mbw, mby = addr / e->mbw;
        memset(mb, 0, sizeof(*mb));
        int t = (int)br_ue(b);
        if (!slice_i && t < 5) {
            if (t != 0)
                FAIL(d, "P partitions other than 16x16 are outside the subset");
            mb->type = GM_MB_P16x16;
            mb->mvd[0] = (int16_t)br_se(b);
            mb->mvd[1] = (int16_t)br_se(b);
            finish_p_mv(d, mb, mbx, mby, 0);
        } else {
            if (set_intra_type(d, mb, slice_i ? t : t - 5))
                return -1;
            if (mb->type == GM_MB_I4x4)
                for (int blk = 0; blk < 16; blk++) {
                    int pm = i4_pred_mode(e, mbx, mby, blk);
                    if (br_u(b, 1))
                        mb->i4_mode[blk] = (uint8_t)pm;
                    else {
                        int rem = (int)br_u(b, 3);
                        mb->i4_mode[blk] = (uint8_t)(rem < pm ? rem : rem + 1);
                    }
                }
            mb->chroma_mode = (uint8_t)br_ue(b);
            if (mb->chroma_mode > 3)
                FAIL(d, "bad intra_chroma_pred_mode");
        }
        if (mb->type != GM_MB_I16x16) {
            uint32_t code = br_ue(b);
            const uint8_t *map = mb_is_intra(mb) ? h264_cbp_to_codenum_intra : h264_cbp_to_codenum_inter;
            int cbp = -1;
            for (int k = 0; k < 48; k++)
                if (map[k] == code)
                    cbp = k;
            if (cbp < 0)
                FAIL(d, "bad coded_block_pattern");
            mb->cbp = (uint8_t)cbp;
        }
        if (mb->type == GM_MB_I16x16 || mb->cbp)
            if (br_se(b))
                FAIL(d, "mb_qp_delta != 0 is outside the subset");
        if (cavlc_read_residual(d, b, mb, mbx, mby))
            return -1;
        if (mb_is_intra(mb))
            recon_intra(d, mb, mbx, mby);
        else if (recon_inter(d, mb, mbx, mby))
            return -1;
        addr++;
    }
    return addr;
}

/* ---------------------------------------------------------------- CABAC parsing (9.3) ---------- */
typedef struct {
    bitr *b;
    uint32_t range, offset;
    uint8_t state[1024]; /* pStateIdx << 1 | valMPS */
} cabd;

static void cabd_init(cabd *c, bitr *b, int slice_i, int qp)
{
    c->b = b;
    for (int i = 0; i < 460; i++) { /* 9.3.1.1 */
        int m = slice_i ? h264_cabac_init_I[i][0] : h264_cabac_init_P0[i][0];
        int n = slice_i ? h264_cabac_init_I[i][1] : h264_cabac_init_P0[i][1];
        int pre = CLIP3(1, 126, ((m * CLIP3(0, 51, qp)) >> 4) + n);
        c->state[i] = (uint8_t)(pre <= 63 ? (63 - pre) << 1 : ((pre - 64) << 1) | 1);
    }
    c->range = 510;
    c->offset = br_u(b, 9);
}
static int cabd_decision(cabd *c, int ctx) /* 9.3.3.2.1 */
{
    int st = c->state[ctx] >> 1, mps = c->state[ctx] & 1, bin;
    uint32_t lps = h264_range_lps[st][(c->range >> 6) & 3];
    c->range -= lps;
    if (c->offset >= c->range) {
        bin = !mps;
        c->offset -= c->range;
        c->range = lps;
        if (st == 0)
            mps = 1 - mps;
        st = h264_next_state_lps[st];
    } else {
        bin = mps;
        st = h264_next_state_mps[st];
    }
    c->state[ctx] = (uint8_t)((st << 1) | mps);
    while (c->range < 256) {
        c->range <<= 1;
        c->offset = (c->offset << 1) | (uint32_t)br_bit(c->b);
    }
    return bin;
}
static int cabd_bypass(cabd *c) /* 9.3.3.2.3 */
{
    c->offset = (c->offset << 1) | (uint32_t)br_bit(c->b);
    if (c->offset >= c->range) {
        c->offset -= c->range;
        return 1;
    }
    return 0;
}
static int cabd_terminate(cabd *c) /* 9.3.3.2.2.3 */
{
    c->range -= 2;
    if (c->offset >= c->range)
        return 1;
    while (c->range < 256) {
        c->range <<= 1;
        c->offset = (c->offset << 1) | (uint32_t)br_bit(c->b);
    }
    return 0;
}
static int cabd_ueg_suffix(cabd *c, int k)
{
    int v = 0;
    while (cabd_bypass(c) && k < 30) {
        v += 1 << k;
        k++;
    }
    while (k--)
        v += cabd_bypass(c) << k;
    return v;
}

static int cabd_read_block(cabd *c, int16_t *lev, int n, int cat, int cbf_inc)
{
    for (int i = 0; i < n; i++)
        lev[i] = 0;
    if (!cabd_decision(c, 85 + cat_cbf_off[cat] + cbf_inc))
        return 0;
    int sig[16], nsig = 0, i;
    for (i = 0; i < n - 1; i++) {
        int inc = cat == 3 ? imin(i, 2) : i;
        if (cabd_decision(c, 105 + cat_sig_off[cat] + inc)) {
            sig[nsig++] = i;
            if (cabd_decision(c, 166 + cat_sig_off[cat] + inc))
                break;
        }
    }
    if (i == n - 1)
        sig[nsig++] = n - 1; /* the last coefficient is inferred to be significant */
    int eq1 = 0, gt1 = 0;
    for (int k = nsig - 1; k >= 0; k--) {
        int ctx = 227 + cat_abs_off[cat] + (gt1 ? 0 : imin(4, 1 + eq1)), a = 0;
        if (cabd_decision(c, ctx)) {
            ctx = 227 + cat_abs_off[cat] + 5 + imin(4 - (cat == 3), gt1);
            a = 1;
            while (a < 14 && cabd_decision(c, ctx))
                a++;
            if (a == 14)
                a += cabd_ueg_suffix(c, 0);
            gt1++;
        } else
            eq1++;
        lev[sig[k]] = (int16_t)(cabd_bypass(c) ? -(a + 1) : a + 1);
    }
    return nsig;
}

static int cabd_mvd(cabd *c, int base, int sum_abs)
{
    int inc = sum_abs < 3 ? 0 : (sum_abs > 32 ? 2 : 1);
    if (!cabd_decision(c, base + inc))
        return 0;
    int a = 1;
    while (a < 9 && cabd_decision(c, base + imin(2 + a, 6)))
        a++;
    if (a == 9)
        a += cabd_ueg_suffix(c, 3);
    return cabd_bypass(c) ? -a : a;
}

/* mb_type of an intra macroblock after its first bin (prefix) has said "not I_NxN"; c1..c5 as in 9.3.3.1.2 */
static int cabd_i16_type(cabd *c, int c1, int c2, int c3, int c4, int c5)
{
    if (cabd_terminate(c))
        return 25; /* I_PCM */
    int t = 1 + 12 * cabd_decision(c, c1);
    if (cabd_decision(c, c2))
        t += 4 + 4 * cabd_decision(c, c3);
    t += 2 * cabd_decision(c, c4);
    t += cabd_decision(c, c5);
    return t;
}

static int cabac_slice_decode(gd_decoder *d, bitr *b, int first_mb, int slice_i)
{
    gm_encoder *e = &d->e;
    while (b->pos & 7)
        if (!br_bit(b))
            FAIL(d, "cabac_alignment_one_bit is zero");
    cabd c;
    cabd_init(&c, b, slice_i, e->qp);
    int nmb = e->mbw * e->mbh, addr = first_mb;
    for (; addr < nmb;) {
        gm_mb *mb = &e->mbs[addr];
        int mbx = addr % e->mbw, mby = addr / e->mbw;
        const gm_mb *A = mbx > 0 ? mb - 1 : NULL, *B = top_avail(e, mby) ? mb - e->mbw : NULL;
        memset(mb, 0, sizeof(*mb));
        int skip = 0;
        if (!slice_i)
            skip = cabd_decision(&c, 11 + (A && A->type != GM_MB_PSKIP) + (B && B->type != GM_MB_PSKIP));
        if (skip) {
            mb->type = GM_MB_PSKIP;
            finish_p_mv(d, mb, mbx, mby, 1);
            if (recon_inter(d, mb, mbx, mby))
                return -1;
        } else {
            int t;
            if (slice_i) {
                int inc = (A && A->type != GM_MB_I4x4) + (B && B->type != GM_MB_I4x4);
                t = cabd_decision(&c, 3 + inc) ? cabd_i16_type(&c, 6, 7, 8, 9, 10) : 0;
            } else if (!cabd_decision(&c, 14)) {
                if (cabd_decision(&c, 15) || cabd_decision(&c, 16))
                    FAIL(d, "P partitions other than 16x16 are outside the subset");
                t = -1;
            } else
                t = cabd_decision(&c, 17) ? cabd_i16_type(&c, 18, 19, 19, 20, 20) : 0;
            if (t < 0)
                mb->type = GM_MB_P16x16;
            else if (set_intra_type(d, mb, t))
                return -1;
            int intra = mb_is_intra(mb);
            if (mb->type == GM_MB_I4x4)
                for (int blk = 0; blk < 16; blk++) {
                    int pm = i4_pred_mode(e, mbx, mby, blk);
                    if (cabd_decision(&c, 68))
                        mb->i4_mode[blk] = (uint8_t)pm;
                    else {
                        int rem = cabd_decision(&c, 69);
                        rem |= cabd_decision(&c, 69) << 1;
                        rem |= cabd_decision(&c, 69) << 2;
                        mb->i4_mode[blk] = (uint8_t)(rem < pm ? rem : rem + 1);
                    }
                }
            if (intra) {
                int inc = (A && mb_is_intra(A) && A->chroma_mode != 0) + (B && mb_is_intra(B) && B->chroma_mode != 0), m = 0;
                if (cabd_decision(&c, 64 + inc)) {
                    m = 1;
                    if (cabd_decision(&c, 67))
                        m = 2 + cabd_decision(&c, 67);
                }
                mb->chroma_mode = (uint8_t)m;
            } else {
                for (int k = 0; k < 2; k++) {
                    int sa = (A && A->type == GM_MB_P16x16 ? iabs(A->mvd[k]) : 0) + (B && B->type == GM_MB_P16x16 ? iabs(B->mvd[k]) : 0);
                    mb->mvd[k] = (int16_t)cabd_mvd(&c, k ? 47 : 40, sa);
                }
                finish_p_mv(d, mb, mbx, mby, 0);
            }
            if (mb->type != GM_MB_I16x16) {
                int cbp_a = A ? (A->cbp & 15) : 15, cbp_b = B ? (B->cbp & 15) : 15, cbpl = 0;
                if (A && A->type == GM_MB_PSKIP)
                    cbp_a = 0;
                if (B && B->type == GM_MB_PSKIP)
                    cbp_b = 0;
                for (int b8 = 0; b8 < 4; b8++) {
                    int la = (b8 & 1) ? (cbpl >> (b8 - 1)) & 1 : (cbp_a >> (b8 + 1)) & 1;
                    int lb = (b8 & 2) ? (cbpl >> (b8 - 2)) & 1 : (cbp_b >> (b8 + 2)) & 1;
                    cbpl |= cabd_decision(&c, 73 + (!la) + 2 * (!lb)) << b8;
                }
                int ca = A ? (A->cbp >> 4) : 0, cb = B ? (B->cbp >> 4) : 0, cbpc = 0;
                if (cabd_decision(&c, 77 + (ca > 0) + 2 * (cb > 0)))
                    cbpc = 1 + cabd_decision(&c, 77 + 4 + (ca == 2) + 2 * (cb == 2));
                mb->cbp = (uint8_t)(cbpl | (cbpc << 4));
            }
            if (mb->type == GM_MB_I16x16 || mb->cbp)
                if (cabd_decision(&c, 60))
                    FAIL(d, "mb_qp_delta != 0 is outside the subset");
            int cbpl = mb->cbp & 15, cbpc = mb->cbp >> 4;
            if (mb->type == GM_MB_I16x16) {
                int inc = cbf_neighbour(e, mbx, mby, 0, 0, 0, 1, intra) + 2 * cbf_neighbour(e, mbx, mby, 0, 0, 0, 0, intra);
                mb->nnz[16] = (uint8_t)cabd_read_block(&c, mb->coef[16], 16, 0, inc);
            }
            for (int blk = 0; blk < 16; blk++)
                if (cbpl & (1 << (blk >> 2))) {
                    int ac = mb->type == GM_MB_I16x16, cat = ac ? 1 : 2;
                    int inc = cbf_neighbour(e, mbx, mby, cat, 0, blk, 1, intra) + 2 * cbf_neighbour(e, mbx, mby, cat, 0, blk, 0, intra);
                    mb->nnz[blk] = (uint8_t)cabd_read_block(&c, mb->coef[blk] + ac, 16 - ac, cat, inc);
                }
            if (cbpc)
                for (int comp = 0; comp < 2; comp++) {
                    int inc = cbf_neighbour(e, mbx, mby, 3, comp, 0, 1, intra) + 2 * cbf_neighbour(e, mbx, mby, 3, comp, 0, 0, intra);
                    mb->nnz[25 + comp] = (uint8_t)cabd_read_block(&c, mb->coef[17] + comp * 4, 4, 3, inc);
                }
            if (cbpc == 2)
                for (int comp = 0; comp < 2; comp++)
                    for (int blk = 0; blk < 4; blk++) {
                        int inc = cbf_neighbour(e, mbx, mby, 4, comp, blk, 1, intra) + 2 * cbf_neighbour(e, mbx, mby, 4, comp, blk, 0, intra);
                        mb->nnz[17 + comp * 4 + blk] = (uint8_t)cabd_read_block(&c, mb->coef[18 + comp * 4 + blk] + 1, 15, 4, inc);
                    }
            if (intra)
                recon_intra(d, mb, mbx, mby);
            else if (recon_inter(d, mb, mbx, mby))
                return -1;
        }
        addr++;
        if (cabd_terminate(&c)) /* end_of_slice_flag */
            break;
    }
    return addr;
}

/* ---------------------------------------------------------------- pictures ---------- */
static int finish_picture(gd_decoder *d)
{
    gm_encoder *e = &d->e;
    if (!d->pic_open)
        return 0;
    frame_t *cur = &e->rec[e->cur];
    size_t ysz = (size_t)e->W * e->H;
    for (int p = 0; p < 3; p++)
        memcpy(cur->p[p], e->unf.p[p], p ? ysz / 4 : ysz);
    if (!d->disable_deblock)
        for (int mby = 0; mby < e->mbh; mby++)
            for (int mbx = 0; mbx < e->mbw; mbx++)
                deblock_mb(e, cur, mbx, mby);
    if (d->nframes == d->cap) {
        d->cap = d->cap ? d->cap * 2 : 16;
        d->frames = (uint8_t **)realloc(d->frames, sizeof(uint8_t *) * (size_t)d->cap);
    }
    uint8_t *f = (uint8_t *)malloc(ysz * 3 / 2);
    memcpy(f, cur->p[0], ysz);
    memcpy(f + ysz, cur->p[1], ysz / 4);
    memcpy(f + ysz + ysz / 4, cur->p[2], ysz / 4);
    d->frames[d->nframes++] = f;
    e->cur ^= 1;
    d->pic_open = 0;
    return 0;
}

static int decode_slice(gd_decoder *d, bitr *b, int nal_type, int ref_idc)
{
    gm_encoder *e = &d->e;
    if (!d->have_sps || !d->have_pps)
        FAIL(d, "slice before SPS / PPS");
    int first_mb = (int)br_ue(b), slice_type = (int)br_ue(b) % 5;
    br_ue(b); /* pps id */
    br_u(b, d->log2_max_frame_num);
    if (slice_type != 0 && slice_type != 2)
        FAIL(d, "slice_type %d is outside the subset", slice_type);
    int slice_i = slice_type == 2;
    if (nal_type == 5)
        br_ue(b); /* idr_pic_id */
    if (d->poc_type == 0)
        FAIL(d, "pic_order_cnt_type 0 slices are outside the subset");
    if (!slice_i) {
        if (br_u(b, 1))
            FAIL(d, "num_ref_idx_active_override is outside the subset");
        if (br_u(b, 1))
            FAIL(d, "reference list modification is outside the subset");
    }
    if (ref_idc) { /* dec_ref_pic_marking() */
        if (nal_type == 5)
            br_u(b, 2);
        else if (br_u(b, 1))
            FAIL(d, "adaptive reference marking is outside the subset");
    }
    if (d->cabac && !slice_i && br_ue(b))
        FAIL(d, "cabac_init_idc != 0 is outside the subset");
    int qp = d->pic_init_qp + br_se(b);
    d->disable_deblock = 0;
    if (d->deblock_ctrl) {
        int idc = (int)br_ue(b);
        if (idc == 2)
            FAIL(d, "disable_deblocking_filter_idc 2 is outside the subset");
        d->disable_deblock = idc == 1;
        if (idc != 1 && (br_se(b) || br_se(b)))
            FAIL(d, "deblocking offsets are outside the subset");
    }
    if (first_mb % e->mbw)
        FAIL(d, "slices that do not start a macroblock row are outside the subset");
    int row = first_mb / e->mbw;
    if (first_mb == 0) {
        finish_picture(d);
        d->pic_open = 1;
        e->srows = e->mbh;
        e->qp = qp;
        e->qpc = h264_chroma_qp[CLIP3(0, 51, qp + d->chroma_qp_offset)];
        if (nal_type != 5 && d->nframes == 0)
            FAIL(d, "stream does not start with an IDR picture");
    } else {
        if (!d->pic_open || qp != e->qp)
            FAIL(d, "slice without a picture / QP change inside a picture");
        /* neighbour availability across slice edges: slices are rows_per_slice rows each (the last may be shorter) */
        int rows = row - d->prev_slice_row;
        if (e->srows != e->mbh && rows != e->srows)
            FAIL(d, "slices of unequal height are outside the subset");
        e->srows = rows;
    }
    d->prev_slice_row = row;
    int end = d->cabac ? cabac_slice_decode(d, b, first_mb, slice_i) : cavlc_slice_decode(d, b, first_mb, slice_i);
    if (end < 0)
        return -1;
    if (end % e->mbw)
        FAIL(d, "slice ends inside a macroblock row (%d)", end);
    return 0;
}

int gd_decode(gd_decoder *d, const uint8_t *s, size_t n)
{
    size_t i = 0;
    d->err[0] = 0;
    while (i + 3 < n) {
        if (!(s[i] == 0 && s[i + 1] == 0 && s[i + 2] == 1)) {
            i++;
            continue;
        }
        size_t start = i + 3, end = start;
        while (end + 2 < n && !(s[end] == 0 && s[end + 1] == 0 && (s[end + 2] == 1 || (s[end + 2] == 0 && end + 3 < n && s[end + 3] == 1))))
            end++;
        if (end + 2 >= n)
            end = n;
        if (end - start + 8 > d->rbsp_cap) {
            d->rbsp_cap = (end - start) * 2 + 64;
            d->rbsp = (uint8_t *)realloc(d->rbsp, d->rbsp_cap);
        }
        int hdr = s[start], zeros = 0;
        size_t m = 0;
        for (size_t k = start + 1; k < end; k++) { /* 7.4.1: drop emulation_prevention_three_byte */
            if (zeros >= 2 && s[k] == 3) {
                zeros = 0;
                continue;
            }
            d->rbsp[m++] = s[k];
            zeros = s[k] ? 0 : zeros + 1;
        }
        bitr b = {d->rbsp, m * 8, 0};
        int type = hdr & 31, r = 0;
        if (type == 7)
            r = parse_sps(d, &b);
        else if (type == 8)
            r = parse_pps(d, &b);
        else if (type == 1 || type == 5)
            r = decode_slice(d, &b, type, (hdr >> 5) & 3);
        if (r)
            return -1;
        i = end;
    }
    finish_picture(d);
    return d->nframes;
}
