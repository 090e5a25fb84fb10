/*
 * h264enc -- drop-in for the reference CLI (/root/reference/userspace/h264enc.c).
 *
 *   h264enc <infile or -> <width> <height> <outfile>        (userspace/h264enc.c:141-147)
 *
 * Same positional arguments, same defaults (NV12, dst = ALIGN16, profile 77, level 41, QP 24,
 * keyframe interval 25, CABAC -- userspace/h264enc.c:50-66), same I/O pattern: read w*h luma then
 * w*h/2 interleaved chroma per frame (:178-187), one write() of the returned byte count per frame
 * (:195), progress line "\rFrame %5d: %5dbytes" (:194), stop at the first short read, exit 0 (:200).
 * The ioctl/mmap calls on /dev/cedar_dev become the C ABI of include/cedar_b200.h.
 * Optional trailing flags (extensions): --qp N --gop N --cavlc --nv16 --me-range N --slice-rows N --crop --auto-level
 * --repeat-headers --intra4x4 --p-intra --batch-gops N --device N --stats
 * --batch-gops N: read N GOPs at a time and encode them GOP-parallel (clip mode of the C ABI); the bytes written, one
 * write() per frame, are identical to the frame-at-a-time default, only later.
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64

#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "cedar_b200.h"

#define ALIGN(x, a) (((x) + ((a)-1)) & ~((a)-1))

static struct cedar_b200_io io;
static cedar_b200_handle *enc;

static int ve_config(struct cedar_b200_config *config) /* userspace/h264enc.c:47-117 */
{
    int ret = cedar_b200_open(config, &io, &enc);
    if (ret) {
        fprintf(stderr, "%s(): cedar_b200_open failed: %s\n", __func__, strerror(-ret));
        return ret;
    }
    printf("Input Y: %dbytes at %p\n", io.input_luma_size, io.input_luma);
    printf("Input C: %dbytes at %p\n", io.input_chroma_size, io.input_chroma);
    printf("Bytestream: %dbytes at %p\n", io.bytestream_size, io.bytestream);
    return 0;
}

static int read_frame(int fd, void *buffer, int size) /* userspace/h264enc.c:119-132 */
{
    int total = 0, len;

    while (total < size) {
        len = (int)read(fd, (char *)buffer + total, (size_t)(size - total));
        if (len <= 0)
            return -1;
        total += len;
    }
    return total;
}

int main(int argc, char **argv)
{
    uint32_t frame_count = 0;
    int width, height, fd_in, fd_out, luma_size, chroma_size, ret, stats = 0, batch_gops = 0;
    struct cedar_b200_config config;
    double sse_total = 0, bytes_total = 0;

    if (argc < 5 || (argc > 5 && strncmp(argv[5], "--", 2))) {
        printf("Usage: %s <infile> <width> <height> <outfile>\n", argv[0]);
        return -1;
    }
    width = atoi(argv[2]);
    height = atoi(argv[3]);

    memset(&config, 0, sizeof(config));
    config.src_width = width;
    config.src_height = height;
    config.src_format = CEDAR_B200_FORMAT_NV12;
    config.dst_width = ALIGN(width, 16);
    config.dst_height = ALIGN(height, 16);
    config.profile = 77;
    config.level = 41;
    config.qp = 24;
    config.keyframe_interval = 25;
    config.entropy_coding_mode = CEDAR_B200_ENTROPY_CABAC;
    for (int i = 5; i < argc; i++) {
        if (!strcmp(argv[i], "--qp") && i + 1 < argc)
            config.qp = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--gop") && i + 1 < argc) {
            config.keyframe_interval = atoi(argv[++i]);
            config.relax_gop = 1;
        } else if (!strcmp(argv[i], "--me-range") && i + 1 < argc)
            config.me_range = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--slice-rows") && i + 1 < argc)
            config.slice_rows = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc)
            config.device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--crop"))
            config.sps_crop = 1;
        else if (!strcmp(argv[i], "--auto-level"))
            config.auto_level = 1;
        else if (!strcmp(argv[i], "--repeat-headers"))
            config.repeat_headers = 1;
        else if (!strcmp(argv[i], "--intra4x4"))
            config.intra4x4 = 1;
        else if (!strcmp(argv[i], "--p-intra"))
            config.p_intra = 1;
        else if (!strcmp(argv[i], "--cavlc"))
            config.entropy_coding_mode = CEDAR_B200_ENTROPY_CAVLC;
        else if (!strcmp(argv[i], "--nv16"))
            config.src_format = CEDAR_B200_FORMAT_NV16;
        else if (!strcmp(argv[i], "--batch-gops") && i + 1 < argc)
            batch_gops = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--stats"))
            stats = 1;
        else {
            fprintf(stderr, "%s: unknown option %s\n", argv[0], argv[i]);
            return -1;
        }
    }

    if (strcmp(argv[1], "-")) {
        fd_in = open(argv[1], O_RDONLY);
        if (fd_in == -1) {
            fprintf(stderr, "%s(): Failed to open input file %s: %s\n", __func__, argv[1], strerror(errno));
            return -1;
        }
    } else
        fd_in = 0;

    fd_out = open(argv[4], O_CREAT | O_RDWR | O_TRUNC, S_IRUSR | S_IWUSR | S_IRGRP | S_IROTH);
    if (fd_out == -1) {
        fprintf(stderr, "%s(): Failed to open output file %s\n", __func__, argv[4]);
        return -1;
    }

    if (batch_gops > 0)
        config.max_clip_frames = batch_gops * config.keyframe_interval;
    ret = ve_config(&config);
    if (ret)
        return ret;

    luma_size = width * height;
    chroma_size = config.src_format == CEDAR_B200_FORMAT_NV16 ? luma_size : luma_size / 2;

    while (batch_gops > 0) { /* GOP-parallel: whole batches through the clip calls of the C ABI */
        size_t frame_bytes = 0;
        uint8_t *in = (uint8_t *)cedar_b200_clip_input(enc, &frame_bytes);
        const uint8_t *out = NULL;
        int n = 0, cap = config.max_clip_frames;
        int *sizes = (int *)malloc(sizeof(int) * (size_t)cap);
        double *sse = (double *)malloc(sizeof(double) * (size_t)cap);
        while (n < cap && read_frame(fd_in, in + (size_t)n * frame_bytes, (int)frame_bytes) == (int)frame_bytes)
            n++;
        if (n > 0) {
            long long total = -1;
            ret = cedar_b200_clip_upload(enc, n);
            if (!ret)
                ret = cedar_b200_clip_encode(enc, n, (int)frame_count);
            if (!ret)
                total = cedar_b200_clip_download(enc, &out, sizes);
            if (ret || total < 0)
                fprintf(stderr, "%s(): %d: clip encode failed: %s\n", __func__, frame_count, strerror(ret ? -ret : (int)-total));
            else {
                if (stats)
                    cedar_b200_stats(enc, sse, n);
                for (int i = 0; i < n; i++) {
                    printf("\rFrame %5d: %5dbytes", frame_count, sizes[i]);
                    if (write(fd_out, out, (size_t)sizes[i]) != sizes[i])
                        fprintf(stderr, "%s(): short write\n", __func__);
                    out += sizes[i];
                    if (stats) {
                        sse_total += sse[i];
                        bytes_total += sizes[i];
                    }
                    frame_count++;
                }
            }
        }
        free(sizes);
        free(sse);
        if (n < cap)
            break;
    }
    while (batch_gops <= 0) {
        ret = read_frame(fd_in, io.input_luma, luma_size);
        if (ret != luma_size)
            break;
        ret = read_frame(fd_in, io.input_chroma, chroma_size);
        if (ret != chroma_size)
            break;

        ret = cedar_b200_encode_frame(enc);
        if (ret < 0)
            fprintf(stderr, "%s(): %d: cedar_b200_encode_frame failed: %s\n", __func__, frame_count, strerror(-ret));
        else {
            printf("\rFrame %5d: %5dbytes", frame_count, ret);
            if (write(fd_out, io.bytestream, (size_t)ret) != ret)
                fprintf(stderr, "%s(): short write\n", __func__);
            if (stats) {
                double sse = 0;
                cedar_b200_stats(enc, &sse, 1);
                sse_total += sse;
                bytes_total += ret;
            }
            frame_count++;
        }
    }
    printf("\n");
    if (stats && frame_count) {
        double mse = sse_total / ((double)frame_count * config.dst_width * config.dst_height);
        fprintf(stderr, "frames %u, %.2f kbit/frame, Y-PSNR %.2f dB\n", frame_count,
                bytes_total * 8.0 / 1000.0 / frame_count, mse > 0 ? 10.0 * log10(255.0 * 255.0 / mse) : 99.0);
    }
    cedar_b200_close(enc);
    return 0;
}
