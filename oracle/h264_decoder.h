/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Golden-model decoder for the H.264 subset the encoder emits (h264_decoder.c).
 */
#ifndef H264_DECODER_H
#define H264_DECODER_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct gd_decoder gd_decoder;
gd_decoder *gd_open(void);
/* Decodes a complete Annex-B stream (may be called once per handle); returns the number of pictures or -1 (gd_error). */
int gd_decode(gd_decoder *d, const uint8_t *stream, size_t n);
const char *gd_error(const gd_decoder *d);
int gd_frames(const gd_decoder *d);
int gd_width(const gd_decoder *d);  /* coded size */
int gd_height(const gd_decoder *d);
int gd_crop_right(const gd_decoder *d);  /* luma samples cropped by the SPS */
int gd_crop_bottom(const gd_decoder *d);
const uint8_t *gd_frame(const gd_decoder *d, int i, int plane); /* Y, U, V at the coded size, after deblocking */
void gd_close(gd_decoder *d);
#ifdef __cplusplus
}
#endif
#endif
