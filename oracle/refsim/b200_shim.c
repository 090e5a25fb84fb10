/*
 * TEST INFRASTRUCTURE (drop-in proof): the reference's own userspace/h264enc.c, compiled unmodified and from where it
 * lies with user_shim.h force-included, linked against THE PRODUCT (libcedar_b200.so) instead of the simulated driver:
 * oracle/_ref/h264enc_b200.  Its open / ioctl / mmap calls on /dev/cedar_dev land here and become the C ABI of
 * include/cedar_b200.h -- the binding INTEGRATION.md section 2 describes, as running code:
 *
 *   open("/dev/cedar_dev")            userspace/h264enc.c:149      -> a placeholder descriptor
 *   ioctl(fd, CEDAR_IOCTL_CONFIG, c)  :68, kernel/cedar.c:732-866  -> cedar_b200_open(); sizes and mmap tokens written back
 *   mmap(.., fd, c.*_dma_addr)        :76-106                      -> the pinned host buffers of struct cedar_b200_io
 *   ioctl(fd, CEDAR_IOCTL_ENCODE)     :189, kernel/cedar.c:1032-1209 -> cedar_b200_encode_frame(): byte count or -1 + errno
 *   process exit                      kernel/cedar.c:706-730       -> cedar_b200_close()
 *
 * struct cedar_ioctl_config and the ioctl numbers come from the reference's kernel/cedar_ioctl.h (-I, not copied).
 */
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/ioctl.h>
#include <sys/mman.h>
#include <unistd.h>

#include "cedar_ioctl.h" /* the reference's */
#include "cedar_b200.h"

static int g_fd = -1;
static cedar_b200_handle *g_enc;
static struct cedar_b200_io g_io;
/* what the "DMA addresses" of the three buffers look like to the caller: page-aligned tokens mmap() recognises */
enum { TOKEN_LUMA = 0x10000000, TOKEN_CHROMA = 0x20000000, TOKEN_BYTESTREAM = 0x30000000 };

static void release_at_exit(void) /* the kernel releases the device when the process exits (h264enc never closes it) */
{
    if (g_enc)
        cedar_b200_close(g_enc);
    g_enc = NULL;
}

int refsim_u_open(const char *path, int flags, ...)
{
    va_list ap;
    va_start(ap, flags);
    int mode = (flags & O_CREAT) ? va_arg(ap, int) : 0;
    va_end(ap);
    if (strcmp(path, CEDAR_DEVICE_PATH))
        return open(path, flags, mode);
    if (g_fd >= 0) { /* one opener at a time, kernel/cedar.c:457-474 */
        errno = EBUSY;
        return -1;
    }
    /* the library leaves its host's environment alone; the side streams of its CABAC stage want their own queues */
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    g_fd = open("/dev/null", O_RDWR); /* reserves a descriptor number */
    atexit(release_at_exit);
    return g_fd;
}

static int config(struct cedar_ioctl_config *c)
{
    struct cedar_b200_config cfg;
    if (g_enc)
        return -EINVAL; /* kernel/cedar.c:744-747: configured once */
    memset(&cfg, 0, sizeof(cfg));
    cfg.src_width = c->src_width, cfg.src_height = c->src_height, cfg.src_format = c->src_format;
    cfg.dst_width = c->dst_width, cfg.dst_height = c->dst_height;
    cfg.profile = c->profile, cfg.level = c->level, cfg.qp = c->qp, cfg.keyframe_interval = c->keyframe_interval;
    cfg.thumbnail = c->thumbnail, cfg.thumbnail_downscale = c->thumbnail_downscale;
    cfg.entropy_coding_mode = c->entropy_coding_mode;
    int r = cedar_b200_open(&cfg, &g_io, &g_enc);
    if (r)
        return r;
    c->input_luma_dma_addr = TOKEN_LUMA, c->input_luma_size = g_io.input_luma_size;
    c->input_chroma_dma_addr = TOKEN_CHROMA, c->input_chroma_size = g_io.input_chroma_size;
    c->bytestream_dma_addr = TOKEN_BYTESTREAM, c->bytestream_size = g_io.bytestream_size;
    c->thumbnail = 0; /* the ISP thumbnail scaler is out of scope */
    c->thumb_luma_dma_addr = c->thumb_chroma_dma_addr = 0, c->thumb_luma_size = c->thumb_chroma_size = 0;
    return 0;
}

int refsim_u_ioctl(int fd, unsigned long request, ...)
{
    va_list ap;
    va_start(ap, request);
    void *arg = va_arg(ap, void *);
    va_end(ap);
    if (fd != g_fd || g_fd < 0)
        return ioctl(fd, request, arg);
    int r;
    switch (request) {
    case CEDAR_IOCTL_CONFIG:
        r = arg ? config((struct cedar_ioctl_config *)arg) : -EFAULT;
        break;
    case CEDAR_IOCTL_ENCODE:
        r = g_enc ? cedar_b200_encode_frame(g_enc) : -EINVAL; /* kernel/cedar.c:1039-1043: not configured */
        break;
    default:
        r = -EPERM; /* kernel/cedar.c:1222-1225 returns -1 */
    }
    if (r < 0) {
        errno = -r;
        return -1;
    }
    return r;
}

void *refsim_u_mmap(void *addr, size_t length, int prot, int flags, int fd, off_t offset)
{
    if (fd != g_fd || g_fd < 0)
        return mmap(addr, length, prot, flags, fd, offset);
    if (g_enc && offset == TOKEN_LUMA && length <= (size_t)g_io.input_luma_size)
        return g_io.input_luma;
    if (g_enc && offset == TOKEN_CHROMA && length <= (size_t)g_io.input_chroma_size)
        return g_io.input_chroma;
    if (g_enc && offset == TOKEN_BYTESTREAM && length <= (size_t)g_io.bytestream_size)
        return g_io.bytestream;
    errno = EINVAL; /* kernel/cedar.c mmap handler: unknown offset */
    return MAP_FAILED;
}
