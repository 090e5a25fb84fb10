/* Host-C header writer (restates /root/reference/kernel/cedar.c:187-223, 868-1030). */
#ifndef CEDAR_HEADERS_H
#define CEDAR_HEADERS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Each returns the number of bytes written (start code + NAL header + escaped RBSP) or -errno. */
int cedar_hdr_sps(int profile, int level, int width_mb, int height_mb, uint8_t *out, int cap);
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
