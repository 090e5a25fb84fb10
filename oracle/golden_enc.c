/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  CLI over the CPU golden model with the same positional
 * arguments as the reference's userspace/h264enc.c:141-147 (config 1 of BASELINE.json runs here).
 * Optional trailing flags: --qp N --gop N --cavlc --nv16 --me-range N --synth NFRAMES
 */
#define _GNU_SOURCE
#include "h264_golden.h"
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define ALIGN(x, a) (((x) + ((a)-1)) & ~((a)-1))

static int read_frame(int fd, uint8_t *buffer, int size) /* userspace/h264enc.c:119-132 */
{
    int total = 0;
    while (total < size) {
        int len = (int)read(fd, buffer + total, (size_t)(size - total));
        if (len <= 0)
            return -1;
        total += len;
    }
    return total;
}

int main(int argc, char **argv)
{
    if (argc < 5) {
        printf("Usage: %s <infile> <width> <height> <outfile>\n", argv[0]);
        return -1;
    }
    int width = atoi(argv[2]), height = atoi(argv[3]);
    gm_config cfg = {0};
    cfg.src_width = width;
    cfg.src_height = height;
    cfg.src_format = GM_FORMAT_NV12;
    cfg.dst_width = ALIGN(width, 16);
    cfg.dst_height = ALIGN(height, 16);
    cfg.profile = 77;
    cfg.level = 41;
    cfg.qp = 24;
    cfg.keyframe_interval = 25;
    cfg.entropy_coding_mode = GM_ENTROPY_CABAC;
    cfg.me_range = 16;
    cfg.relax_gop = 1;
    int synth = 0;
    for (int i = 5; i < argc; i++) {
        if (!strcmp(argv[i], "--qp") && i + 1 < argc)
            cfg.qp = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--gop") && i + 1 < argc)
            cfg.keyframe_interval = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--me-range") && i + 1 < argc)
            cfg.me_range = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--synth") && i + 1 < argc)
            synth = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--slice-rows") && i + 1 < argc)
            cfg.slice_rows = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--cavlc"))
            cfg.entropy_coding_mode = GM_ENTROPY_CAVLC;
        else if (!strcmp(argv[i], "--nv16"))
            cfg.src_format = GM_FORMAT_NV16;
    }
    int fd_in = 0;
    if (!synth && strcmp(argv[1], "-")) {
        fd_in = open(argv[1], O_RDONLY);
        if (fd_in == -1) {
            fprintf(stderr, "Failed to open input file %s\n", argv[1]);
            return -1;
        }
    }
    int fd_out = open(argv[4], O_CREAT | O_RDWR | O_TRUNC, 0644);
    if (fd_out == -1) {
        fprintf(stderr, "Failed to open output file %s\n", argv[4]);
        return -1;
    }
    gm_encoder *e;
    int ret = gm_open(&cfg, &e);
    if (ret) {
        fprintf(stderr, "config failed: %d\n", ret);
        return ret;
    }
    int luma_size = width * height;
    int chroma_size = cfg.src_format == GM_FORMAT_NV16 ? luma_size : luma_size / 2;
    int cap = cfg.dst_width * cfg.dst_height * 3 + 65536;
    uint8_t *luma = malloc((size_t)luma_size), *chroma = malloc((size_t)chroma_size), *out = malloc((size_t)cap);
    int frame_count = 0;
    double sse = 0;
    long bytes = 0;
    while (1) {
        if (synth) {
            if (frame_count >= synth)
                break;
            gm_synth_frame(width, height, cfg.src_format, frame_count, luma, chroma);
        } else {
            if (read_frame(fd_in, luma, luma_size) != luma_size)
                break;
            if (read_frame(fd_in, chroma, chroma_size) != chroma_size)
                break;
        }
        ret = gm_encode_frame(e, luma, chroma, out, cap);
        if (ret < 0)
            fprintf(stderr, "%d: encode failed: %d\n", frame_count, ret);
        else {
            printf("\rFrame %5d: %5dbytes", frame_count, ret);
            if (write(fd_out, out, (size_t)ret) != ret)
                return -1;
            sse += gm_last_sse_y(e);
            bytes += ret;
            frame_count++;
        }
    }
    printf("\n");
    if (frame_count) {
        double mse = sse / ((double)frame_count * cfg.dst_width * cfg.dst_height);
        fprintf(stderr, "frames %d, %.2f kbit/frame, Y-MSE %.4f\n", frame_count, bytes * 8.0 / 1000.0 / frame_count, mse);
    }
    gm_close(e);
    return 0;
}
