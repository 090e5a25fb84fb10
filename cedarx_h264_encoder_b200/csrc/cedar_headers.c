/*
 * Host-C SPS / PPS / slice-header writer of the B200 encoder.
 *
 * Restates the only H.264 logic the reference has in source, /root/reference/kernel/cedar.c:
 *   :187-223  put-bits port, ue(v), se(v)          -> struct hbits below (software bit buffer)
 *   :868-881  start code (written with EPB disabled) + NAL header byte
 *   :883-890  rbsp_trailing_bits including its quirk (a whole extra 0x00 when 7 bits are used)
 *   :892-937  SPS    :939-982  PPS    :984-1030  slice header
 * The VE hardware inserts emulation-prevention bytes; here it is done in software.
 * The output must be byte-identical to what the reference emits for the same config.
 */
#include "cedar_headers.h"

#include <errno.h>
#include <string.h>

struct hbits {
    uint8_t buf[64];
    int nbits;
};

static void hb_put(struct hbits *b, uint32_t data, int size) /* cedar.c:187-207: size & 0x1f bits, MSB first */
{
    size &= 0x1f;
    while (size--) {
        int bit = (data >> size) & 1;
        if ((b->nbits >> 3) < (int)sizeof(b->buf))
            b->buf[b->nbits >> 3] |= (uint8_t)(bit << (7 - (b->nbits & 7)));
        b->nbits++;
    }
}

static void hb_ue(struct hbits *b, uint32_t data) /* cedar.c:209-215 */
{
    data++;
    hb_put(b, data, (32 - __builtin_clz(data)) * 2 - 1);
}

static void hb_se(struct hbits *b, int32_t data) /* cedar.c:217-223 */
{
    data = (2 * data) - 1;
    data ^= (data >> 31);
    hb_ue(b, (uint32_t)data);
}

static void hb_trailing(struct hbits *b) /* cedar.c:883-890, STMLEN & 7 == nbits & 7 (NALs start aligned) */
{
    uint32_t len = (uint32_t)b->nbits;
    int pad = 8 - ((len + 1) & 0x7);
    /* cedar_bytestream_write masks the size with 0x1f; pad + 1 <= 9 so nothing is lost */
    hb_put(b, 1u << pad, pad + 1);
}

static int emit_nal(uint8_t *out, int cap, int ref_idc, int type, const struct hbits *b)
{
    int n = 0, zeros = 0, bytes = b->nbits >> 3;
    if (cap < 5)
        return -ENOMEM;
    out[n++] = 0; /* cedar.c:875: 24 zero bits */
    out[n++] = 0;
    out[n++] = 0;
    out[n++] = 1; /* cedar.c:876-878: 0x100 | ref_idc << 5 | type in 16 bits */
    out[n++] = (uint8_t)(((ref_idc & 3) << 5) | (type & 0x1F));
    for (int i = 0; i < bytes; i++) {
        if (zeros >= 2 && b->buf[i] <= 3) {
            if (n >= cap)
                return -ENOMEM;
            out[n++] = 3;
            zeros = 0;
        }
        if (n >= cap)
            return -ENOMEM;
        out[n++] = b->buf[i];
        zeros = b->buf[i] ? 0 : zeros + 1;
    }
    return n;
}

int cedar_hdr_min_level(int mbs)
{
    static const int max_fs[][2] = {{99, 10},   {396, 11},  {792, 21},   {1620, 22},  {3600, 31},  {5120, 32},
                                    {8192, 40}, {8704, 42}, {22080, 50}, {36864, 51}, {139264, 60}};
    for (unsigned i = 0; i < sizeof(max_fs) / sizeof(max_fs[0]); i++)
        if (mbs <= max_fs[i][0])
            return max_fs[i][1];
    return 62;
}

int cedar_hdr_sps(int profile, int level, int width_mb, int height_mb, int crop_right, int crop_bottom, uint8_t *out,
                  int cap)
{
    struct hbits b;
    memset(&b, 0, sizeof(b));
    hb_put(&b, (uint32_t)profile, 8); /* :897 */
    hb_put(&b, 0, 8);                 /* :899 constraints */
    hb_put(&b, (uint32_t)level, 8);   /* :900 */
    hb_ue(&b, 0);                     /* :903 seq_parameter_set_id */
    hb_ue(&b, 0);                     /* :906 log2_max_frame_num_minus4 */
    hb_ue(&b, 2);                     /* :908 pic_order_cnt_type */
    hb_ue(&b, 1);                     /* :911 max_num_ref_frames */
    hb_put(&b, 0, 1);                 /* :913 gaps_in_frame_num_value_allowed_flag */
    hb_ue(&b, (uint32_t)(width_mb - 1));  /* :915 */
    hb_ue(&b, (uint32_t)(height_mb - 1)); /* :916 */
    hb_put(&b, 1, 1);                 /* :919 frame_mbs_only_flag */
    hb_put(&b, 0, 1);                 /* :922 direct_8x8_inference_flag */
    if (crop_right || crop_bottom) {  /* :924-929, dead in the reference (:756-761); the `sps_crop` extension */
        hb_put(&b, 1, 1);             /* frame_cropping_flag */
        hb_ue(&b, 0);                 /* frame_crop_left_offset */
        hb_ue(&b, (uint32_t)crop_right);
        hb_ue(&b, 0);                 /* frame_crop_top_offset */
        hb_ue(&b, (uint32_t)crop_bottom);
    } else
        hb_put(&b, 0, 1);             /* :931 frame_cropping_flag (crop is always 0: :756-761) */
    hb_put(&b, 0, 1);                 /* :934 vui_parameters_present_flag */
    hb_trailing(&b);                  /* :936 */
    return emit_nal(out, cap, 3, 7, &b);
}

int cedar_hdr_pps(int qp, int cabac, uint8_t *out, int cap)
{
    struct hbits b;
    memset(&b, 0, sizeof(b));
    hb_ue(&b, 0);                 /* :944 pic_parameter_set_id */
    hb_ue(&b, 0);                 /* :946 seq_parameter_set_id */
    hb_put(&b, cabac ? 1 : 0, 1); /* :948-951 entropy_coding_mode_flag */
    hb_put(&b, 0, 1);             /* :954 bottom_field_pic_order_in_frame_present_flag */
    hb_ue(&b, 0);                 /* :956 num_slice_groups_minus1 */
    hb_ue(&b, 0);                 /* :959 num_ref_idx_l0_default_active_minus1 */
    hb_ue(&b, 0);                 /* :961 num_ref_idx_l1_default_active_minus1 */
    hb_put(&b, 0, 1);             /* :964 weighted_pred_flag */
    hb_put(&b, 0, 2);             /* :966 weighted_bipred_idc */
    hb_se(&b, qp - 26);           /* :968 pic_init_qp_minus26 */
    hb_se(&b, qp - 26);           /* :969 pic_init_qs_minus26 */
    hb_se(&b, 4);                 /* :971 chroma_qp_index_offset */
    hb_put(&b, 1, 1);             /* :974 deblocking_filter_control_present_flag */
    hb_put(&b, 0, 1);             /* :976 constrained_intra_pred_flag */
    hb_put(&b, 0, 1);             /* :978 redundant_pic_cnt_present_flag */
    hb_trailing(&b);              /* :981 */
    return emit_nal(out, cap, 3, 8, &b);
}

int cedar_hdr_slice_mb(int frame_i, int frame_p_count, int cabac, int first_mb, uint64_t *bits, int *nbits)
{
    struct hbits b;
    memset(&b, 0, sizeof(b));
    hb_ue(&b, (uint32_t)first_mb); /* :993 first_mb_in_slice (0 in the reference) */
    hb_ue(&b, frame_i ? 2 : 0);  /* :994-997 slice_type */
    hb_ue(&b, 0);                /* :999 pic_parameter_set_id */
    hb_put(&b, (uint32_t)frame_p_count & 0x0F, 4); /* :1001 frame_num */
    if (frame_i) {
        hb_ue(&b, 0);     /* :1005 idr_pic_id */
        hb_put(&b, 0, 1); /* :1007 no_output_of_prior_pics_flag */
        hb_put(&b, 0, 1); /* :1009 long_term_reference_flag */
    } else {
        hb_put(&b, 0, 1); /* :1012 num_ref_idx_active_override_flag */
        hb_put(&b, 0, 1); /* :1014 ref_pic_list_modification_flag_l0 */
        hb_put(&b, 0, 1); /* :1016 adaptive_ref_pic_marking_mode_flag */
        if (cabac)
            hb_ue(&b, 0); /* :1017-1018 cabac_init_idc */
    }
    hb_se(&b, 0); /* :1022 slice_qp_delta */
    hb_ue(&b, 0); /* :1025 disable_deblocking_filter_idc */
    hb_se(&b, 0); /* :1027 slice_alpha_c0_offset_div2 */
    hb_se(&b, 0); /* :1029 slice_beta_offset_div2 */
    if (b.nbits > 64)
        return -EINVAL;
    *nbits = b.nbits;
    uint64_t v = 0;
    for (int i = 0; i < 8; i++)
        v = (v << 8) | b.buf[i];
    *bits = v >> (64 - b.nbits);
    return 0;
}

int cedar_hdr_slice(int frame_i, int frame_p_count, int cabac, uint32_t *bits, int *nbits)
{
    uint64_t v = 0;
    int r = cedar_hdr_slice_mb(frame_i, frame_p_count, cabac, 0, &v, nbits);
    if (r)
        return r;
    *bits = (uint32_t)v;
    return 0;
}
