/* Host-C header writer (restates /root/reference/kernel/cedar.c:187-223, 868-1030). */
#ifndef CEDAR_HEADERS_H
#define CEDAR_HEADERS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Each returns the number of bytes written (start code + NAL header + escaped RBSP) or -errno. */
/* crop_right / crop_bottom: frame_crop_*_offset in crop units (2 luma samples for 4:2:0 frames); both 0 = no cropping
 * syntax, which is all the reference can emit (kernel/cedar.c:756-761 makes its cropping branch :924-931 dead). */
int cedar_hdr_sps(int profile, int level, int width_mb, int height_mb, int crop_right, int crop_bottom, uint8_t *out,
                  int cap);
/* Lowest level_idc whose MaxFS holds a picture of `mbs` macroblocks (H.264 table A-1); the reference writes the
 * configured level unchecked (kernel/cedar.c:900). */
int cedar_hdr_min_level(int mbs);
int cedar_hdr_pps(int qp, int cabac, uint8_t *out, int cap);
/* Slice header bits that follow the NAL header byte, right aligned in *bits (<= 32 bits). */
int cedar_hdr_slice(int frame_i, int frame_p_count, int cabac, uint32_t *bits, int *nbits);
/* The same with first_mb_in_slice = first_mb (the reference always writes 0, cedar.c:992-993; non-zero only with
 * the slice_rows extension); right aligned in *bits (<= 64 bits). */
int cedar_hdr_slice_mb(int frame_i, int frame_p_count, int cabac, int first_mb, uint64_t *bits, int *nbits);
#ifdef __cplusplus
}
#endif
#endif
