/*
 * h264enc -- drop-in for the reference CLI (/root/reference/userspace/h264enc.c).
 *
 *   h264enc <infile or -> <width> <height> <outfile>        (userspace/h264enc.c:141-147)
 *
 * Same positional arguments, same defaults (NV12, dst = ALIGN16, profile 77, level 41, QP 24,
 * keyframe interval 25, CABAC -- userspace/h264enc.c:50-66), same I/O pattern: read w*h luma then
 * w*h/2 interleaved chroma per frame (:178-187), one write() of the returned byte count per frame
 * (:195), progress line "\rFrame %5d: %5dbytes" (:194), the three buffer lines (:86,100,113), stop at
 * the first short read, exit 0 (:200).  The ioctl/mmap calls on /dev/cedar_dev become the C ABI of
 * include/cedar_b200.h.
 *
 * Optional trailing flags (extensions): --qp N --gop N --cavlc --nv16 --me-range N --slice-rows N --crop --auto-level
 * --repeat-headers --intra4x4 --p-intra --device N --stats, and the GOP-parallel modes, all of which write exactly the
 * bytes of the frame-at-a-time default, one write() per frame, only later:
 *   --queue-gops N   the reference's loop unchanged (fill buffers, encode_frame, write ret bytes) on a queued handle:
 *                    encode_frame returns frame t - 2N GOPs while batches of N GOPs are encoded behind it; flush drains
 *   --batch-gops N   a reader thread fills batches of N GOPs while earlier batches are encoded and written
 *                    (cedar_b200_pipe_*); --handles M encoder handles per GPU (default 2)
 *   --gpus N         the same across GPUs 0..N-1 of this box (batch b -> worker b mod (N*M)); --devices a,b,c picks them
 *   --reader-threads N  a batch of a regular input file is read by N threads with pread() (default 4; a pipe is read by
 *                    one thread, as is N = 1): one thread copies 7.7 GB/s out of the page cache, 2 500 frames/s of 1080p
 */
#define _GNU_SOURCE
#define _FILE_OFFSET_BITS 64

#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "cedar_b200.h"

#define ALIGN(x, a) (((x) + ((a)-1)) & ~((a)-1))

static struct cedar_b200_io io;
static cedar_b200_handle *enc;
static cedar_b200_pipe *g_pipe;

static int ve_config(struct cedar_b200_config *config) /* userspace/h264enc.c:47-117 */
{
    int ret = cedar_b200_open(config, &io, &enc);
    if (ret) {
        fprintf(stderr, "%s(): cedar_b200_open failed: %s\n", __func__, strerror(-ret));
        return ret;
    }
    /* the reference prints "<size>bytes at 0x<dma address> -> <mapping>" (:86,100,113); the buffers are pinned host
     * memory and have no bus address: 0 is printed in its place */
    printf("Input Y: %dbytes at 0x%08X -> %p\n", io.input_luma_size, 0u, io.input_luma);
    printf("Input C: %dbytes at 0x%08X -> %p\n", io.input_chroma_size, 0u, io.input_chroma);
    printf("Bytestream: %dbytes at 0x%08X -> %p\n", io.bytestream_size, 0u, io.bytestream);
    return 0;
}

/* userspace/h264enc.c:119-132.  The reference only stops on end of file (len == 0) and would spin on a read error
 * (len < 0 is added to total); here an error ends the stream like end of file does. */
static int read_frame(int fd, void *buffer, int size)
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

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static double sse_total, bytes_total;
static uint32_t frame_count;

static void emit_frame(int fd_out, const void *data, int bytes, int stats, double sse)
{
    printf("\rFrame %5d: %5dbytes", frame_count, bytes);
    if (write(fd_out, data, (size_t)bytes) != bytes)
        fprintf(stderr, "%s(): short write\n", __func__);
    if (stats) {
        sse_total += sse;
        bytes_total += bytes;
    }
    frame_count++;
}

/* ---- a batch of a regular file, read by several threads ---- */
struct pread_job {
    int fd;
    uint8_t *dst;
    off_t off;
    size_t frame_bytes;
    int nframes, done; /* done: whole frames read, from the first one of the job */
};

static void *pread_main(void *arg)
{
    struct pread_job *j = (struct pread_job *)arg;
    for (j->done = 0; j->done < j->nframes; j->done++) {
        size_t got = 0;
        while (got < j->frame_bytes) {
            ssize_t len = pread(j->fd, j->dst + (size_t)j->done * j->frame_bytes + got, j->frame_bytes - got,
                                j->off + (off_t)((size_t)j->done * j->frame_bytes + got));
            if (len <= 0)
                return NULL; /* end of file inside a frame, or an error: the stream ends before this frame */
            got += (size_t)len;
        }
    }
    return NULL;
}

/* Reads up to `cap` whole frames from the current position of the regular file `fd` into `dst` with `threads` threads
 * and moves the position behind them.  Returns the number of frames read: fewer than `cap` at the end of the file (a
 * trailing partial frame is dropped, like the read loop of userspace/h264enc.c:181-187 does). */
static int read_batch_parallel(int fd, uint8_t *dst, size_t frame_bytes, int cap, int threads)
{
    struct pread_job job[16];
    pthread_t th[16];
    int started[16], n = 0, nt, per, stop = 0;
    struct stat sb;
    off_t pos = lseek(fd, 0, SEEK_CUR);
    if (pos < 0 || fstat(fd, &sb) || cap <= 0)
        return 0;
    if (sb.st_size > pos && (off_t)((sb.st_size - pos) / (off_t)frame_bytes) < (off_t)cap)
        cap = (int)((sb.st_size - pos) / (off_t)frame_bytes);
    else if (sb.st_size <= pos)
        cap = 0;
    if (cap == 0)
        return 0;
    nt = threads > 16 ? 16 : threads;
    if (nt > cap)
        nt = cap;
    per = (cap + nt - 1) / nt;
    for (int i = 0; i < nt; i++) {
        int first = i * per, cnt = cap - first < per ? cap - first : per;
        job[i].fd = fd, job[i].dst = dst + (size_t)first * frame_bytes, job[i].off = pos + (off_t)((size_t)first * frame_bytes);
        job[i].frame_bytes = frame_bytes, job[i].nframes = cnt > 0 ? cnt : 0, job[i].done = 0;
        started[i] = job[i].nframes > 0 && i > 0 && !pthread_create(&th[i], NULL, pread_main, &job[i]);
    }
    pread_main(&job[0]); /* this thread takes the first share */
    for (int i = 1; i < nt; i++) {
        if (started[i])
            pthread_join(th[i], NULL);
        else if (job[i].nframes > 0)
            pread_main(&job[i]); /* no thread to be had: read the share here */
    }
    for (int i = 0; i < nt && !stop; i++) { /* the frames that are there without a gap (a file truncated meanwhile) */
        n += job[i].done;
        stop = job[i].done < job[i].nframes;
    }
    lseek(fd, pos + (off_t)((size_t)n * frame_bytes), SEEK_SET);
    return n;
}

#ifdef H264ENC_READER_SELFTEST
/* tests/test_cli.py: h264enc_reader_selftest <file> <frame_bytes> <cap> <threads> -> the batches' bytes on stdout */
int main(int argc, char **argv)
{
    if (argc != 5)
        return 2;
    size_t fb = (size_t)atol(argv[2]);
    int cap = atoi(argv[3]), threads = atoi(argv[4]), fd = open(argv[1], O_RDONLY), n;
    uint8_t *buf = malloc(fb * (size_t)cap);
    if (fd < 0 || !buf)
        return 3;
    do {
        n = read_batch_parallel(fd, buf, fb, cap, threads);
        if (n > 0 && fwrite(buf, fb, (size_t)n, stdout) != (size_t)n)
            return 4;
        fprintf(stderr, "%d\n", n);
    } while (n == cap);
    return 0;
}
#else

/* ---- --batch-gops / --gpus: reader thread -> pipeline workers -> writer (this thread) ---- */
struct reader_ctx {
    cedar_b200_pipe *pipe;
    int fd_in, threads; /* threads > 1: fd_in is a regular file and a batch is read with pread() by that many threads */
};

static void *reader_main(void *arg)
{
    struct reader_ctx *rc = (struct reader_ctx *)arg;
    const int trace = getenv("CEDAR_B200_TRACE") != NULL;
    for (;;) {
        size_t frame_bytes = 0;
        int cap = 0, n = 0;
        double t0 = now_s();
        uint8_t *in = (uint8_t *)cedar_b200_pipe_acquire(rc->pipe, &frame_bytes, &cap);
        double t1 = now_s();
        if (!in)
            break;
        if (rc->threads > 1)
            n = read_batch_parallel(rc->fd_in, in, frame_bytes, cap, rc->threads);
        else
            while (n < cap && read_frame(rc->fd_in, in + (size_t)n * frame_bytes, (int)frame_bytes) == (int)frame_bytes)
                n++;
        if (trace)
            fprintf(stderr, "[trace] reader: waited %.1f ms for a staging buffer, read %d frames in %.1f ms\n",
                    1e3 * (t1 - t0), n, 1e3 * (now_s() - t1));
        if (cedar_b200_pipe_submit(rc->pipe, n) || n < cap)
            break;
    }
    cedar_b200_pipe_finish(rc->pipe);
    return NULL;
}

static int run_pipeline(struct cedar_b200_config *config, const int *devices, int ndevices, int handles, int batch_gops,
                        int fd_in, int fd_out, int stats, double *t_opened, int reader_threads)
{
    cedar_b200_pipe *pipe = NULL;
    struct reader_ctx rc;
    pthread_t reader;
    int ret = cedar_b200_pipe_open(config, devices, ndevices, handles, batch_gops, &pipe);
    if (ret) {
        fprintf(stderr, "%s(): cedar_b200_pipe_open failed: %s\n", __func__, strerror(-ret));
        return ret;
    }
    rc.pipe = pipe;
    rc.fd_in = fd_in;
    {
        struct stat sb;
        rc.threads = (reader_threads > 1 && !fstat(fd_in, &sb) && S_ISREG(sb.st_mode)) ? reader_threads : 1;
    }
    *t_opened = now_s();
    if (pthread_create(&reader, NULL, reader_main, &rc)) {
        cedar_b200_pipe_close(pipe);
        return -EAGAIN;
    }
    for (;;) {
        const uint8_t *out = NULL;
        const int *sizes = NULL;
        const double *sse = NULL;
        int n = 0;
        long long total = cedar_b200_pipe_next(pipe, &out, &sizes, &n, &sse, 1);
        if (total == 0)
            break;
        if (total < 0) {
            fprintf(stderr, "%s(): %d: batch encode failed: %s\n", __func__, frame_count, strerror((int)-total));
            continue;
        }
        for (int i = 0; i < n; i++) {
            emit_frame(fd_out, out, sizes[i], stats, sse[i]);
            out += sizes[i];
        }
    }
    pthread_join(reader, NULL);
    g_pipe = pipe; /* closed by main, after the timing line */
    return 0;
}

int main(int argc, char **argv)
{
    const double t_start = now_s();
    double t_open = 0, t_stream = 0;
    int width, height, fd_in, fd_out, luma_size, chroma_size, ret, stats = 0, batch_gops = 0, queue_gops = 0;
    int devices[64], ndevices = 0, handles = 0, one_device = -1, reader_threads = 4;
    struct cedar_b200_config config;

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
            one_device = config.device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) {
            int n = atoi(argv[++i]);
            if (n < 1 || n > 64) {
                fprintf(stderr, "%s: --gpus %d out of range\n", argv[0], n);
                return -1;
            }
            for (ndevices = 0; ndevices < n; ndevices++)
                devices[ndevices] = ndevices;
        } else if (!strcmp(argv[i], "--devices") && i + 1 < argc) {
            char *s = argv[++i];
            for (ndevices = 0; *s && ndevices < 64; ndevices++) {
                devices[ndevices] = (int)strtol(s, &s, 10);
                if (*s == ',')
                    s++;
            }
        } else if (!strcmp(argv[i], "--handles") && i + 1 < argc)
            handles = atoi(argv[++i]);
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
        else if (!strcmp(argv[i], "--queue-gops") && i + 1 < argc)
            queue_gops = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--reader-threads") && i + 1 < argc)
            reader_threads = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--stats"))
            stats = 1;
        else {
            fprintf(stderr, "%s: unknown option %s\n", argv[0], argv[i]);
            return -1;
        }
    }
    if (batch_gops < 0 || queue_gops < 0 || handles < 0 || (batch_gops && queue_gops) || reader_threads < 1 || reader_threads > 16) {
        fprintf(stderr, "%s: invalid --batch-gops / --queue-gops / --handles / --reader-threads\n", argv[0]);
        return -1;
    }
    if (ndevices > 0 && !batch_gops)
        batch_gops = 4; /* several GPUs only make sense GOP-parallel */
    /* the serial CABAC stages of successive frames overlap on side streams: they need their own hardware queues, and
     * the variable only counts before CUDA initialises in this process */
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);

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

    luma_size = width * height;
    chroma_size = config.src_format == CEDAR_B200_FORMAT_NV16 ? luma_size : luma_size / 2;

    if (batch_gops > 0) { /* GOP-parallel: reader thread, pipeline workers (per GPU, per handle), ordered writes */
        if (ndevices == 0 && one_device >= 0)
            devices[ndevices++] = one_device;
        ret = run_pipeline(&config, ndevices ? devices : NULL, ndevices, handles, batch_gops, fd_in, fd_out, stats, &t_open,
                           reader_threads);
        if (ret)
            return ret;
        t_stream = now_s();
    } else {
        config.queue_gops = queue_gops;
        ret = ve_config(&config);
        if (ret)
            return ret;
        t_open = now_s();
        while (1) { /* userspace/h264enc.c:181-198, unchanged in shape */
            ret = read_frame(fd_in, io.input_luma, luma_size);
            if (ret != luma_size)
                break;
            ret = read_frame(fd_in, io.input_chroma, chroma_size);
            if (ret != chroma_size)
                break;

            ret = cedar_b200_encode_frame(enc);
            if (ret < 0)
                fprintf(stderr, "%s(): %d: cedar_b200_encode_frame failed: %s\n", __func__, frame_count, strerror(-ret));
            else if (ret > 0 || !queue_gops) {
                double sse = 0;
                if (stats)
                    cedar_b200_stats(enc, &sse, 1);
                emit_frame(fd_out, io.bytestream, ret, stats, sse);
            }
        }
        while (queue_gops && (ret = cedar_b200_flush(enc)) != 0) { /* queued handle: the frames still in flight */
            double sse = 0;
            if (ret < 0) {
                fprintf(stderr, "%s(): %d: cedar_b200_flush failed: %s\n", __func__, frame_count, strerror(-ret));
                break;
            }
            if (stats)
                cedar_b200_stats(enc, &sse, 1);
            emit_frame(fd_out, io.bytestream, ret, stats, sse);
        }
    }
    if (!t_stream)
        t_stream = now_s();
    /* the reference ends without a newline after the last progress line (userspace/h264enc.c:194-200) */
    fflush(stdout);
    if (stats && frame_count) {
        double mse = sse_total / ((double)frame_count * config.dst_width * config.dst_height);
        fprintf(stderr, "frames %u, %.2f kbit/frame, Y-PSNR %.2f dB\n", frame_count,
                bytes_total * 8.0 / 1000.0 / frame_count, mse > 0 ? 10.0 * log10(255.0 * 255.0 / mse) : 99.0);
    }
    if (enc)
        cedar_b200_close(enc);
    if (g_pipe)
        cedar_b200_pipe_close(g_pipe);
    if (stats) /* where the wall time went: CUDA start-up and buffer allocation, the stream itself, teardown */
        fprintf(stderr, "timing: open %.3f s, stream %.3f s (%.1f frames/s), close %.3f s\n", t_open - t_start,
                t_stream - t_open, frame_count / (t_stream - t_open > 0 ? t_stream - t_open : 1), now_s() - t_stream);
    return 0;
}
#endif /* H264ENC_READER_SELFTEST */
